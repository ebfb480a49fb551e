/*
 * libfhp — C-ABI of the B200-native batched foothold planner.
 *
 * This header is the drop-in boundary for ONE hot path of lukechencqu/quadrupedal_foothold_planner: the
 * per-gait-cycle, per-leg foothold search behind the `plan_global_footholds` service
 * (foothold_planner/src/FootholdPlanner.cpp, "FP.cpp" below).  ROS and grid_map stay host-side; a ROS node
 * keeps its `foothold_planner::FootholdPlanner` class and calls these entry points from
 *   - FootholdPlanner::gridmapCallback   (FP.cpp:504-536)  -> fhp_set_map_msg (the message itself) or fhp_set_map
 *   - FootholdPlanner::globalFootholdPlan (FP.cpp:539-1602) -> fhp_plan / fhp_plan_ex (n_queries = 1)
 * (binding shown in INTEGRATION.md).  Batch callers use n_queries >> 1, fhp_plan_compact[_packed] for a
 * service-shaped variable-length response, fhp_plan_device for device-resident buffers.
 *
 * Conventions: plain pointers and sizes, no exceptions cross the boundary, every function returns an int
 * status (FHP_OK == 0); fhp_last_error(ctx) returns a human-readable message for the last failure.
 * There is NO CPU fallback: every compute entry point fails with FHP_ERR_CUDA when no sm_100 device works.
 *
 * Frames and layouts are grid_map's: a layer is a column-major float32 matrix (Eigen::MatrixXf), element
 * (row,col) at col*rows + row, row <-> x index, col <-> y index; the circular-buffer start index is kept.
 */
#ifndef FHP_H_
#define FHP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FHP_OK 0
#define FHP_ERR_INVALID_ARG 1   /* null pointer, negative size, gait_cycles > 255, ... */
#define FHP_ERR_NO_MAP 2        /* fhp_plan before fhp_set_map */
#define FHP_ERR_MISSING_LAYER 3 /* "traversability" or "elevation" absent (GridMap::at would throw, FP.cpp:2057,2533) */
#define FHP_ERR_CUDA 4          /* CUDA runtime failure or no usable device */
#define FHP_ERR_UNSUPPORTED 5   /* geometry outside what the kernel tables cover (see fhp_plan) */
#define FHP_ERR_NCCL 6

/* foot ids of the response, FP.cpp:685-699 / :1382-1396 */
enum { FHP_RF = 0, FHP_RH = 1, FHP_LH = 2, FHP_LF = 3 };
/* how a response row was produced */
enum {
  FHP_KIND_STANCE = 0,    /* initial stance row, FP.cpp:685-699 */
  FHP_KIND_DEFAULT = 1,   /* checkDefaultFoothold accepted the centre, FP.cpp:2012 */
  FHP_KIND_CANDIDATE = 2  /* checkCandidateFoothold chose a spiral cell, FP.cpp:2022 */
};
/* gait presets: per-leg bias tables relative to the stance centroid */
enum {
  FHP_GAIT_TROT = 0,   /* the reference's isosceles-trapezoid table, FP.cpp:403-421 */
  FHP_GAIT_WALK = 1,   /* extension: x phase offsets (LH -3/8, LF -1/8, RH +1/8, RF +3/8)*stepLength on +-L/2, y = +-W/2 */
  FHP_GAIT_CUSTOM = 2  /* extension: bias[][] below is used verbatim */
};
/* grid_map::SpiralIterator ignores the map's start index */
enum {
  FHP_SPIRAL_WRAP_CORRECT = 0, /* spiral walks unwrapped indices; identical to the reference when start index == (0,0) */
  FHP_SPIRAL_FAITHFUL = 1      /* bug-compatible with upstream for start index != (0,0) */
};

/* Geometry of a grid_map::GridMap (what gridmapCallback caches, FP.cpp:508-515, plus the start index). */
typedef struct fhp_map_desc {
  int32_t rows, cols;            /* GridMap::getSize()      */
  double resolution;             /* GridMap::getResolution() */
  double length_x, length_y;     /* GridMap::getLength()    */
  double pos_x, pos_y;           /* GridMap::getPosition()  */
  int32_t start_row, start_col;  /* GridMap::getStartIndex() */
  int32_t n_layers;
  const char* const* layer_names; /* n_layers names; must contain "traversability" and "elevation" */
} fhp_map_desc;

/* The rosparam values the path reads (FP.cpp:249-311) with the reference's C++ types
 * (FootholdPlanner.hpp:609-697: float for radii/thresholds/step/kinematics, double for h_). */
typedef struct fhp_params {
  float foot_radius;        /* footRadius                 */
  float search_radius;      /* searchRadius               */
  float step_length;        /* stepLength                 */
  float skew;               /* laikago_kinematics/skewLength (isos_.skew) */
  float thr_default;        /* defaultFootholdThreshold   */
  float thr_candidate;      /* candidateFootholdThreshold */
  float kin_length;         /* laikago_kinematics/length  */
  float kin_width;          /* laikago_kinematics/width   */
  float kin_l1;             /* laikago_kinematics/l1      */
  int32_t rf_first;         /* RF_FIRST                   */
  int32_t gait;             /* FHP_GAIT_*                 */
  int32_t spiral_mode;      /* FHP_SPIRAL_*               */
  double h;                 /* h_ = 0.01, FP.cpp:336      */
  double drift_per_cycle;   /* -0.007, FP.cpp:1578        */
  double bias[4][2];        /* RF,RH,LH,LF (x,y); read only when gait == FHP_GAIT_CUSTOM */
  double heading;           /* extension: body heading in the map frame, radians; 0 = the reference (which always advances
                               along +x, FP.cpp:2199-2202).  Rotates the stance, the step vector, the bias table and the
                               search rectangle; the lateral pin initialPose.y + adjusted_y (FP.cpp:2201) becomes a pin
                               along the rotated normal.  Exact definition: DESIGN.md §8 */
} fhp_params;

/* One independent plan request: the robot's initial stance centre (rosparam initial_position/{x,y,z},
 * FP.cpp:293-295) and which parameter set it uses. */
typedef struct fhp_query {
  double x, y, z;
  int32_t param_id; /* index into the params array passed to fhp_plan */
  int32_t reserved;
} fhp_query;

/* One response row: foothold_planner_msgs/Foothold {point, foot_id, gait_cycle_id} (msg/Foothold.msg:1-3)
 * plus the grid cell and how it was chosen.  40 bytes. */
typedef struct fhp_foothold {
  double x, y, z;        /* point */
  int32_t row, col;      /* buffer index of the cell containing (x,y); (-1,-1) when outside the map */
  uint8_t foot_id;       /* FHP_RF.. */
  uint8_t gait_cycle_id; /* FP.cpp:1378 */
  uint8_t kind;          /* FHP_KIND_* */
  uint8_t reserved0;
  uint32_t reserved1;
} fhp_foothold;

/* Header fields of foothold_planner_msgs/GlobalFootholds (msg/GlobalFootholds.msg:1-5). 8 bytes. */
typedef struct fhp_query_status {
  uint8_t success;             /* outcome of the LAST cycle, FP.cpp:1380 / :1574 */
  uint8_t gait_cycles;         /* the request, FP.cpp:592-593 */
  uint8_t gait_cycles_succeed; /* index+1 of the last successful cycle, FP.cpp:1379 */
  uint8_t reserved;            /* flags: FHP_STATUS_* below */
  int32_t n_rows;              /* rows written for this query: 4 + 4*(successful cycles) */
} fhp_query_status;

/* fhp_query_status.reserved: the low 7 bits are kernel error bits (never set in a call that returned FHP_OK).
 * FHP_STATUS_OPT_SUBMAP_FAIL: the reference would have ABORTED this service call at gait cycle 0 (`return false`,
 * FP.cpp:931-934) because the gait-cycle submap of its third, NLopt-driven "opt" chain (getGaitCycleSearchGridMap
 * FP.cpp:2307-2350, rect = isos_.length x isos_.width around the next stance centre) does not fit the map.  libfhp does
 * not run that chain; at cycle 0 its state is still the first-gait stance, so the predicate is exact there (later cycles
 * depend on NLopt's output and are not modelled).  A host shim that wants the reference's bool returns false when the
 * bit is set (csrc/FootholdPlannerGpu.hpp does); the rows are planned regardless. */
#define FHP_STATUS_ERR_MASK 0x7f
#define FHP_STATUS_OPT_SUBMAP_FAIL 0x80

typedef struct fhp_ctx fhp_ctx;

/* Rows reserved per query in the output arrays: 4 stance rows + 4 per cycle. */
#define FHP_ROWS_PER_QUERY(gait_cycles) (4 + 4 * (gait_cycles))

/* Library / device ------------------------------------------------------------------------------------ */
/* Creates a planner bound to CUDA device `device` (its own stream).  Replaces nothing in the reference
 * (the reference has no device); lifetime = the FootholdPlanner object (FP.cpp:181-245). */
int fhp_create(int device, fhp_ctx** out);
void fhp_destroy(fhp_ctx* ctx);
const char* fhp_last_error(const fhp_ctx* ctx); /* ctx may be NULL: message of the last failed fhp_create */
/* Fills `p` with foothold_planner/config/foothold_planner.yaml:17-44 (trot, RF_FIRST false).
 * NOTE spiral_mode = FHP_SPIRAL_WRAP_CORRECT: identical to the reference while the map's start index is (0,0); on a map
 * whose circular buffer has moved the REFERENCE's footholds need FHP_SPIRAL_FAITHFUL (upstream's SpiralIterator ignores
 * the start index).  The reference-shaped host classes (FootholdPlannerGpu.hpp, planner.FootholdPlanner) set FAITHFUL. */
void fhp_default_params(fhp_params* p);
const char* fhp_version(void);

/* Map ingest — replaces the storage side of gridmapCallback (FP.cpp:504-536): after
 * GridMapRosConverter::fromMessage the node passes gridmap_[layer].data() for every layer.  Host buffers are
 * copied to the device once per map update (layer-planar, column-major, start index kept); the caller keeps
 * ownership of the host buffers. */
int fhp_set_map(fhp_ctx* ctx, const fhp_map_desc* desc, const float* const* layers_colmajor_host);
/* Message ingest — replaces GridMapRosConverter::fromMessage + the copy above in one step (FP.cpp:506; SURVEY §8f
 * rank 1).  grid_map_ros is third-party and absent from the reference tree; the semantics restated here are those of
 * grid_map_ros 1.6.x fromMessage / multiArrayMessageCopyToMatrixEigen / GridMap::setGeometry:
 *   size = (int)round(length / resolution) per axis, length := size * resolution, position = info.pose.position.xy,
 *   layers.size() must equal data.size(), every data[i] must be column-major (layout.dim[0].label == "column_index",
 *   dim[0].size = cols, dim[1].size = rows; a "row_index" dim[0] is refused like upstream refuses it; any OTHER label is
 *   read as column-major, as upstream's isRowMajor() does after logging — the call returns FHP_OK and fhp_last_error
 *   holds upstream's log text), start index = (outer, inner)_start_index.
 * The float arrays are read IN PLACE (no Eigen copy) and streamed to the device through pinned staging buffers.
 * Unlike the reference (which ignores fromMessage's return value and keeps a half-updated map), a malformed message
 * returns FHP_ERR_INVALID_ARG with upstream's error text and leaves the previous map in place.
 *
 * fhp_gridmap_msg: the fields of a DESERIALISED grid_map_msgs/GridMap (what a roscpp callback holds); strings are
 * (pointer, length) pairs so std::string::data()/size() can be passed without copies. */
typedef struct fhp_str {
  const char* p;
  uint32_t n;
} fhp_str;
typedef struct fhp_multiarray {   /* std_msgs/Float32MultiArray */
  fhp_str dim0_label;             /* layout.dim[0].label */
  uint32_t dim0_size;             /* layout.dim[0].size */
  fhp_str dim1_label;             /* layout.dim[1].label */
  uint32_t dim1_size;             /* layout.dim[1].size */
  uint32_t n_dims;                /* layout.dim.size() (upstream reads dim[0] and dim[1]) */
  uint32_t reserved;
  const float* data;              /* data.data() — may be unaligned when it points into wire bytes */
  uint64_t n;                     /* data.size() */
} fhp_multiarray;
typedef struct fhp_gridmap_msg {
  double resolution, length_x, length_y; /* info.resolution, info.length_x, info.length_y */
  double pos_x, pos_y;                   /* info.pose.position.x / .y */
  int32_t n_layers;                      /* layers.size() */
  int32_t n_data;                        /* data.size() */
  const fhp_str* layers;
  const fhp_multiarray* data;
  uint16_t outer_start_index, inner_start_index;
  uint32_t reserved;
} fhp_gridmap_msg;
int fhp_set_map_msg(fhp_ctx* ctx, const fhp_gridmap_msg* msg);
/* Same from the ROS1 WIRE bytes of a grid_map_msgs/GridMap (what a ShapeShifter / rosbag reader / raw TCPROS
 * subscriber holds): little-endian fields in message-definition order, strings and arrays prefixed by a uint32
 * count.  The bytes are parsed in place; truncated or inconsistent input returns FHP_ERR_INVALID_ARG. */
int fhp_set_map_wire(fhp_ctx* ctx, const uint8_t* bytes, uint64_t len);
/* Geometry of the map currently held (the gridMapInfo_ the reference's callback prints, FP.cpp:508-515);
 * layer_names is set to NULL.  FHP_ERR_NO_MAP before the first ingest. */
int fhp_map_info(const fhp_ctx* ctx, fhp_map_desc* out);
/* Same, but the layers are ALREADY device-resident on ctx's device (e.g. after an NCCL broadcast); the
 * pointers are adopted without a copy and must stay valid until the next fhp_set_map* / fhp_destroy. */
int fhp_set_map_device(fhp_ctx* ctx, const fhp_map_desc* desc, const float* const* layers_colmajor_device);
/* Device pointer of an ingested layer (NULL if absent) — lets a host fill/broadcast the map in place.  After rewriting
 * a layer in place (also a layer adopted through fhp_set_map_device) call fhp_map_changed: the default staging path
 * plans on cell-class planes derived from the traversability layer, which are rebuilt by the next plan only after a
 * fhp_set_map*, fhp_broadcast_map or fhp_map_changed. */
const float* fhp_map_layer_device(const fhp_ctx* ctx, const char* layer_name);
int fhp_map_changed(fhp_ctx* ctx);

/* The plan — replaces the centroid + nominal chains of FootholdPlanner::globalFootholdPlan
 * (FP.cpp:762-1579) for n_queries independent requests in ONE kernel launch.  Blocking; host buffers.
 *   out_nominal  : n_queries * FHP_ROWS_PER_QUERY(gait_cycles) rows; query q owns rows
 *                  [q*R, q*R + status[q].n_rows) — exactly response.footholds.footholds (FP.cpp:1588).
 *   out_centroid : NULL, or the same shape: the centroid chain rows (topic global_footholds_centroid,
 *                  FP.cpp:1448-1462) including their z.
 *   status       : n_queries entries.
 * Parameter sets that differ in foot_radius / search_radius are planned in one launch per radii pair (fhp_plan /
 * fhp_plan_ex only; the device-resident, compact and streamed calls take ONE radii pair per call and return
 * FHP_ERR_INVALID_ARG otherwise).  FHP_ERR_UNSUPPORTED when search_radius/resolution or foot_radius/resolution exceed
 * the kernel's window tables (windows wider than 32 cells, more than 15 spiral rings), or when a foot-disc stencil
 * offset lies within rounding distance of the disc boundary (see DESIGN.md "Refused rather than guessed"). */
int fhp_plan(fhp_ctx* ctx, const fhp_params* params, int32_t n_params, const fhp_query* queries, int64_t n_queries,
             int32_t gait_cycles, fhp_foothold* out_nominal, fhp_foothold* out_centroid, fhp_query_status* status);
/* Extended outputs of the same plan.  default_chain: the reference's "default" chain rows —
 * un-searched nominal footholds with their mean height (getDefaultFootholds FP.cpp:2253-2304, recorded at
 * :666-671 / :1344-1348 and published as global_footholds_default_marker); same shape as the nominal rows. */
typedef struct fhp_plan_outputs {
  fhp_foothold* nominal;       /* required */
  fhp_foothold* centroid;      /* optional */
  fhp_foothold* default_chain; /* optional */
  fhp_query_status* status;    /* required */
} fhp_plan_outputs;
int fhp_plan_ex(fhp_ctx* ctx, const fhp_params* params, int32_t n_params, const fhp_query* queries, int64_t n_queries,
                int32_t gait_cycles, const fhp_plan_outputs* out);
/* The same plan with a COMPACT response, as the service returns it: query q owns rows
 * out_rows[row_offsets[q] .. row_offsets[q+1]) = its 4 stance rows + 4 per successful cycle (nominal chain,
 * response.footholds.footholds of FP.cpp:1588), back to back; row_offsets has n_queries + 1 entries.  capacity_rows is
 * the size of out_rows; n_queries * FHP_ROWS_PER_QUERY(gait_cycles) always suffices.  When it is too small the call
 * returns FHP_ERR_INVALID_ARG with row_offsets[n_queries] = the rows needed.  Only produced rows cross PCIe, and
 * they do so chunk by chunk while the next chunk plans. */
int fhp_plan_compact(fhp_ctx* ctx, const fhp_params* params, int32_t n_params, const fhp_query* queries,
                     int64_t n_queries, int32_t gait_cycles, fhp_foothold* out_rows, int64_t capacity_rows,
                     int64_t* row_offsets, fhp_query_status* status);
/* The compact response with 32-byte rows: the same information as fhp_foothold with the cell index as two uint16
 * (0xFFFF = outside the map) — 20 % fewer bytes over PCIe.  FHP_ERR_UNSUPPORTED for maps with more than 65,534 rows or
 * columns (the grid_map message itself carries uint16 start indices). */
typedef struct fhp_foothold_packed {
  double x, y, z;        /* point */
  uint16_t row, col;     /* buffer index of the cell containing (x,y); 0xFFFF when outside the map */
  uint8_t foot_id;
  uint8_t gait_cycle_id;
  uint8_t kind;
  uint8_t reserved;
} fhp_foothold_packed;
int fhp_plan_compact_packed(fhp_ctx* ctx, const fhp_params* params, int32_t n_params, const fhp_query* queries,
                            int64_t n_queries, int32_t gait_cycles, fhp_foothold_packed* out_rows,
                            int64_t capacity_rows, int64_t* row_offsets, fhp_query_status* status);
/* The compact response with 24-byte rows: exactly the fields of foothold_planner_msgs/Foothold (msg/Foothold.msg:1-3) —
 * point, foot_id, gait_cycle_id — plus `kind`.  z travels as the float it is by construction (getFootholdMeanHeight
 * returns float, FP.cpp:2553; stance rows carry (float)initial z, exact for the reference's z = 0 poses and within
 * float rounding otherwise); the grid cell of a row is a pure function of (x, y) and the map geometry and is not
 * sent.  40 % fewer PCIe bytes than fhp_foothold; no host-side expansion is needed to fill the message. */
typedef struct fhp_foothold24 {
  double x, y;
  float z;
  uint8_t foot_id;
  uint8_t gait_cycle_id;
  uint8_t kind;
  uint8_t reserved;
} fhp_foothold24;
int fhp_plan_compact24(fhp_ctx* ctx, const fhp_params* params, int32_t n_params, const fhp_query* queries,
                       int64_t n_queries, int32_t gait_cycles, fhp_foothold24* out_rows, int64_t capacity_rows,
                       int64_t* row_offsets, fhp_query_status* status);
/* STREAMED batch response: one launch plans the whole batch and WRITES every produced foothold straight into the
 * caller's page-locked host buffer while it runs (zero-copy over PCIe: no staging copy, no compaction pass; the transfer
 * overlaps the planning, failed cycles cost no bytes).  A row is the point alone, 20 bytes {x, y double, z float — it is a
 * float by construction, FP.cpp:2553}; foot_id and gait_cycle_id are positional.
 *   Queries are grouped in runs of 8 (group G = q / 8).  out_blocks holds (gait_cycles + 1) * ceil(n_queries / 8) blocks
 *   of 32 rows (640 bytes).  Block (0, G) holds the stance rows of its queries in query order; block (g + 1, G) holds the
 *   rows of gait cycle g of those queries of the group whose cycle g succeeded (FP.cpp:1323), packed to the front in query
 *   order, 4 rows (RF, RH, LH, LF) each.  masks[(q * FHP_STREAM_MASK_WORDS(gait_cycles) + g / 32) * 4 + f]: bit g % 32 of
 *   .success = cycle g of query q succeeded (same value for the 4 feet), of .candidate = that foot came from the spiral
 *   search (kind FHP_KIND_CANDIDATE instead of FHP_KIND_DEFAULT).  fhp_stream_row() below locates a row.
 * The rows are exactly those of fhp_plan / fhp_plan_compact24 ((float)z is lossless for every planned row; stance rows
 * carry (float)initial z).  out_blocks (128-byte aligned), masks and status must be page-locked and mapped
 * (cudaHostAlloc, cudaHostRegister, torch pin_memory): FHP_ERR_INVALID_ARG otherwise — pageable callers use
 * fhp_plan_compact24. */
#pragma pack(push, 4)
typedef struct fhp_point20 {
  double x, y;
  float z;
} fhp_point20;
#pragma pack(pop)
typedef struct fhp_stream_mask {
  uint32_t success, candidate;
} fhp_stream_mask;
#define FHP_STREAM_MASK_WORDS(gait_cycles) (((gait_cycles) + 31) / 32)
#define FHP_STREAM_BLOCKS(n_queries, gait_cycles) (((int64_t)(gait_cycles) + 1) * (((int64_t)(n_queries) + 7) / 8))
int fhp_plan_stream(fhp_ctx* ctx, const fhp_params* params, int32_t n_params, const fhp_query* queries,
                    int64_t n_queries, int32_t gait_cycles, fhp_point20* out_blocks, fhp_stream_mask* masks,
                    fhp_query_status* status);
/* Row of foot f of query q in gait cycle g (g = -1: stance) inside the streamed blocks, or NULL when that cycle
 * failed: rank = the successful queries before q in its group of 8.  Pure host-side function. */
const fhp_point20* fhp_stream_row(const fhp_point20* blocks, const fhp_stream_mask* masks, int64_t n_queries,
                                  int32_t gait_cycles, int64_t q, int32_t g, int32_t f);
/* Host-side expander of a streamed response into the service-shaped compact response of fhp_plan_compact24: query q owns
 * out_rows[row_offsets[q] .. row_offsets[q+1]) = its 4 stance rows + 4 per successful cycle, with foot_id, gait_cycle_id
 * and kind filled in.  Pure host function (n_threads <= 0: one per hardware thread, at most 16); for consumers that want
 * rows rather than fhp_stream_row lookups.  capacity_rows as in fhp_plan_compact. */
int fhp_stream_expand(const fhp_point20* blocks, const fhp_stream_mask* masks, const fhp_query_status* status,
                      int64_t n_queries, int32_t gait_cycles, fhp_foothold24* out_rows, int64_t capacity_rows,
                      int64_t* row_offsets, int32_t n_threads);
/* KPIs the reference logs per successful cycle (getHipDistance FP.cpp:2571-2584, getCogSpeed :2587-2623): pure
 * host-side functions of ONE query's response rows (nominal or centroid chain).  feet_distance / cog_speed receive
 * 2 values per successful cycle (n_rows/4 - 1 cycles). */
int fhp_cycle_kpi(const fhp_params* params, const fhp_query* query, const fhp_foothold* rows, int32_t n_rows,
                  double* feet_distance, double* cog_speed);
/* Same with DEVICE-resident queries/outputs on ctx's device; asynchronous on ctx's stream unless
 * `cuda_stream` (a cudaStream_t) is non-NULL.  The caller synchronises.  Map updates, parameter changes and
 * fhp_map_changed wait for the whole device before they rewrite state such a plan may still be reading. */
int fhp_plan_device(fhp_ctx* ctx, const fhp_params* params, int32_t n_params, const fhp_query* d_queries,
                    int64_t n_queries, int32_t gait_cycles, fhp_foothold* d_out_nominal, fhp_foothold* d_out_centroid,
                    fhp_query_status* d_status, void* cuda_stream);
int fhp_synchronize(fhp_ctx* ctx);

/* Fine-grained primitives with the semantics of the reference's public members, batched over n independent
 * inputs (host buffers, blocking).  xy arrays hold n (x,y) pairs.
 *   fhp_check_foothold      FootholdPlanner::checkFoothold (FP.cpp:2001-2036): centre, rectangle centre ->
 *                           row {x,y,z,row,col,kind}, valid flag.
 *   fhp_check_centroid      FootholdPlanner::checkFootholdUseCentroidMethod (FP.cpp:1605-1997): point ->
 *                           row {x,y,z,...}; case_out: 0 whole-window valid, 1/2/3 the reference's cases,
 *                           4 no case matched (result 0,0,0), -1 submap failed.
 *   fhp_mean_height         FootholdPlanner::getFootholdMeanHeight (FP.cpp:2520-2554).
 *   fhp_check_default       FootholdPlanner::checkDefaultFoothold (FP.cpp:2039-2082). */
int fhp_check_foothold(fhp_ctx* ctx, const fhp_params* params, const double* centres_xy, const double* rect_centres_xy,
                       int64_t n, fhp_foothold* out, uint8_t* valid_out);
int fhp_check_centroid(fhp_ctx* ctx, const fhp_params* params, const double* points_xy, int64_t n, fhp_foothold* out,
                       int8_t* case_out);
int fhp_mean_height(fhp_ctx* ctx, const fhp_params* params, const double* centres_xy, int64_t n, float* z_out);
int fhp_check_default(fhp_ctx* ctx, const fhp_params* params, const double* centres_xy, int64_t n, uint8_t* valid_out);
/* Host-side geometry members (pure functions of the parameters; no device involved, same expressions as the kernel).
 * fhp_initial_stance: initialize()'s stance RF,RH,LH,LF (FP.cpp:350-378) and setFirstGait (FP.cpp:2679-2699);
 * 4 x (x,y,z) each.
 * fhp_search_geometry: for the four current feet cur_xyz (RF,RH,LH,LF): getPolygonCenter (FP.cpp:2421-2463) ->
 * center_xyz[3]; getFootholdSearchGridMap (FP.cpp:2166-2250) with getDefaultFootholdNext (:2411-2418) -> next_xy[4][2];
 * getSearchPolygon (:2496-2517) -> polygon_xy[4][4][2] in the reference's vertex order.  adjusted_y = ajustedPose_[1]
 * (0 at the first cycle, -0.007 per cycle after). */
int fhp_initial_stance(const fhp_params* params, const fhp_query* query, double* initial_xyz, double* first_gait_xyz);
int fhp_search_geometry(const fhp_params* params, const fhp_query* query, double adjusted_y, const double* cur_xyz,
                        double* center_xyz, double* next_xy, double* polygon_xy);
/* checkCandidateFoothold (FP.cpp:2085-2114): the SpiralIterator search alone (no default test, no height).  out: xy
 * and cell of the first valid cell in spiral order when valid_out is 1, zeros / (-1,-1) otherwise. */
int fhp_check_candidate(fhp_ctx* ctx, const fhp_params* params, const double* centres_xy, const double* rect_centres_xy,
                        int64_t n, fhp_foothold* out, uint8_t* valid_out);
/* checkCirclePolygonFoothold (FP.cpp:2117-2163) at arbitrary foot centres: 1 iff the foot disc visits at least one
 * cell and every finite cell of it is >= candidateFootholdThreshold and inside the search rectangle around
 * rect_centres_xy (getSearchPolygon, FP.cpp:2496-2517). */
int fhp_check_circle_polygon(fhp_ctx* ctx, const fhp_params* params, const double* foot_centres_xy,
                             const double* rect_centres_xy, int64_t n, uint8_t* valid_out);

/* Kernel launch counter (for bench.py's "gpu_launches"): number of libfhp kernels launched on this ctx. */
int64_t fhp_kernel_launches(const fhp_ctx* ctx);
/* Multi-GPU: replicate the map of `ctx` (root's map) to all ranks of an NCCL communicator (ncclComm_t passed
 * as void*).  The only collective of the path; symbols are resolved from the NCCL already loaded in the
 * process (dlsym), so libfhp has no link-time NCCL dependency.  On non-root ranks `desc` gives the geometry
 * (same on every rank) and device buffers are allocated here. */
int fhp_broadcast_map(fhp_ctx* ctx, const fhp_map_desc* desc, void* nccl_comm, int root, int rank);

#ifdef __cplusplus
}
#endif
#endif /* FHP_H_ */
