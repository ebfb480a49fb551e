/*
 * libfhp — C-ABI of the B200-native batched foothold planner.
 *
 * This header is the drop-in boundary for ONE hot path of lukechencqu/quadrupedal_foothold_planner: the
 * per-gait-cycle, per-leg foothold search behind the `plan_global_footholds` service
 * (foothold_planner/src/FootholdPlanner.cpp, "FP.cpp" below).  ROS and grid_map stay host-side; a ROS node
 * keeps its `foothold_planner::FootholdPlanner` class and calls these entry points from
 *   - FootholdPlanner::gridmapCallback   (FP.cpp:504-536)  -> fhp_set_map
 *   - FootholdPlanner::globalFootholdPlan (FP.cpp:539-1602) -> fhp_plan (n_queries = 1)
 * (binding shown in INTEGRATION.md).  Batch callers use n_queries >> 1.
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
  uint8_t reserved;
  int32_t n_rows;              /* rows written for this query: 4 + 4*(successful cycles) */
} fhp_query_status;

typedef struct fhp_ctx fhp_ctx;

/* Rows reserved per query in the output arrays: 4 stance rows + 4 per cycle. */
#define FHP_ROWS_PER_QUERY(gait_cycles) (4 + 4 * (gait_cycles))

/* Library / device ------------------------------------------------------------------------------------ */
/* Creates a planner bound to CUDA device `device` (its own stream).  Replaces nothing in the reference
 * (the reference has no device); lifetime = the FootholdPlanner object (FP.cpp:181-245). */
int fhp_create(int device, fhp_ctx** out);
void fhp_destroy(fhp_ctx* ctx);
const char* fhp_last_error(const fhp_ctx* ctx); /* ctx may be NULL: message of the last failed fhp_create */
/* Fills `p` with foothold_planner/config/foothold_planner.yaml:17-44 (trot, RF_FIRST false). */
void fhp_default_params(fhp_params* p);
const char* fhp_version(void);

/* Map ingest — replaces the storage side of gridmapCallback (FP.cpp:504-536): after
 * GridMapRosConverter::fromMessage the node passes gridmap_[layer].data() for every layer.  Host buffers are
 * copied to the device once per map update (layer-planar, column-major, start index kept); the caller keeps
 * ownership of the host buffers. */
int fhp_set_map(fhp_ctx* ctx, const fhp_map_desc* desc, const float* const* layers_colmajor_host);
/* Same, but the layers are ALREADY device-resident on ctx's device (e.g. after an NCCL broadcast); the
 * pointers are adopted without a copy and must stay valid until the next fhp_set_map* / fhp_destroy. */
int fhp_set_map_device(fhp_ctx* ctx, const fhp_map_desc* desc, const float* const* layers_colmajor_device);
/* Device pointer of an ingested layer (NULL if absent) — lets a host fill/broadcast the map in place. */
const float* fhp_map_layer_device(const fhp_ctx* ctx, const char* layer_name);

/* The plan — replaces the centroid + nominal chains of FootholdPlanner::globalFootholdPlan
 * (FP.cpp:762-1579) for n_queries independent requests in ONE kernel launch.  Blocking; host buffers.
 *   out_nominal  : n_queries * FHP_ROWS_PER_QUERY(gait_cycles) rows; query q owns rows
 *                  [q*R, q*R + status[q].n_rows) — exactly response.footholds.footholds (FP.cpp:1588).
 *   out_centroid : NULL, or the same shape: the centroid chain rows (topic global_footholds_centroid,
 *                  FP.cpp:1448-1462) including their z.
 *   status       : n_queries entries.
 * All parameter sets of one call must share foot_radius and search_radius.  FHP_ERR_UNSUPPORTED when
 * search_radius/resolution or foot_radius/resolution exceed the kernel's window tables, or when a foot-disc
 * stencil offset lies within rounding distance of the disc boundary (see DESIGN.md "stencil classification"). */
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
/* KPIs the reference logs per successful cycle (getHipDistance FP.cpp:2571-2584, getCogSpeed :2587-2623): pure
 * host-side functions of ONE query's response rows (nominal or centroid chain).  feet_distance / cog_speed receive
 * 2 values per successful cycle (n_rows/4 - 1 cycles). */
int fhp_cycle_kpi(const fhp_params* params, const fhp_query* query, const fhp_foothold* rows, int32_t n_rows,
                  double* feet_distance, double* cog_speed);
/* Same with DEVICE-resident queries/outputs on ctx's device; asynchronous on ctx's stream unless
 * `cuda_stream` (a cudaStream_t) is non-NULL.  The caller synchronises. */
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
