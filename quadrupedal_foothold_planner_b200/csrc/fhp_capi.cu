// libfhp C-ABI (include/fhp.h): context, map ingest, parameter derivation, geometry tables, launch plumbing.
// No CPU fallback lives here: every compute entry point launches the sm_100a kernels or fails.
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstddef>
#include <cstring>
#include <dlfcn.h>
#include <atomic>
#include <mutex>
#include <thread>
#include <algorithm>
#include <string>
#include <vector>

#include <cstdlib>

#include "fhp_host.h"

using namespace fhp;

namespace {

std::string g_create_error;

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

}  // namespace

struct fhp_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;  // second lane of the host-buffer pipeline in fhp_plan
  mutable std::string err;
  std::mutex mu;  // one plan at a time per ctx (the reference class is not re-entrant either)
  // map
  bool haveMap = false;
  MapDev M{};
  std::vector<std::string> layerNames;
  std::vector<const float*> layerPtrs;  // device
  std::vector<float*> owned;            // device buffers we allocated
  size_t ownedBytes = 0;                // size of each owned layer buffer (reused when a map update keeps its size)
  // pinned staging of the host->device map upload (two chunks in flight; allocated on the first large upload)
  void* pin[2] = {nullptr, nullptr};
  cudaEvent_t pinEv[2] = {nullptr, nullptr};
  bool pinUsed[2] = {false, false};
  // geometry tables cache
  bool haveGeo = false;
  float geoFoot = 0, geoSearch = 0;
  bool geoTrim = false;
  GeoDev G{};
  // TMA descriptor of the traversability layer (box = one staged window); rebuilt with the geometry tables
  alignas(64) CUtensorMap tmapTrav;
  // preferred window staging path: FHP_STAGING=mask|tma|ldg (default mask; FHP_DISABLE_TMA=1 is the old spelling
  // of ldg).  The tests run every path; calls the preferred path cannot serve fall back mask -> tma -> ldg.
  int stagePref = kStageMask;
  int geoStage = -1;  // staging path the cached geometry tables were laid out for
  // experiment knobs, read once at fhp_create: FHP_CHUNK_DIV (host pipeline chunk = wave / div), FHP_QPW_SHIFT (force
  // log2 queries per warp), FHP_TMA_BOX_BUDGET (bytes of TMA box buffers per warp)
  int chunkDiv = 3;
  long long waveQueries[2] = {0, 0};  // plan_wave_queries(G, wantCen) for the cached geometry (0 = not asked yet)
  int qpwShiftForce = -1;
  int tmaBoxBudget = 8192;
  // FHP_FORCE_HEADING=1 (test knob, read at fhp_create): run the heading expressions even for heading == 0, where they
  // must reduce bit-exactly to the reference's (tests/test_gpu_heading.py)
  bool noSpec = false;  // FHP_NO_SPEC=1 (test knob): never pick a geometry-specialised kernel variant (GeoK)
  bool noOnePar = false;  // FHP_NO_ONEPAR=1 (test knob): single-parameter-set calls read the device table like the others
  bool forceHeading = false;
  bool anyHeading = false;  // some parameter set of the current call has the heading expressions in force
  bool asyncPlans = false;  // a plan was launched on a caller's stream since the last device-wide synchronisation
  // D/C/F mask planes of the traversability layer for one threshold pair (rebuilt lazily after a map update)
  DevBuf dMask;
  bool maskValid = false;
  float maskThrD = 0, maskThrC = 0;
  bool maskNegInf = false;  // the layer holds -inf cells: their raw compare differs from their class -> float paths
  long long maskBuilds = 0;
  // scratch
  ParDev hPar0;  // derived table entry 0 of the last call (travels in the kernel arguments)
  DevBuf dPars, dQ, dOutN, dOutC, dOutD, dSt, dErr, dA, dB, dC2, dD, dCompact, dOffs;
  // zero-copy scratch for latency-bound calls (a robot's single plan): the kernel reads the queries from and writes the
  // rows to mapped pinned host memory, so the call is one launch and one stream sync with no copy engine involved
  void* hSmall = nullptr;
  void* dSmall = nullptr;
  cudaStream_t copyStream = nullptr;   // response rows of fhp_plan_compact travel here while the next chunk plans
  long long* hTotals = nullptr;        // mapped pinned: rows per chunk + 1, written by the scan kernel (fhp_plan_compact)
  long long* dTotals = nullptr;        // device alias
  int hTotalsCap = 0;
  std::vector<cudaEvent_t> chunkEv;
  long long launches = 0;
  unsigned pendingErrCheck = 0;
  // kernel error bits live in zero-copy host memory: reading them after a stream sync costs no memcpy
  unsigned* hErr = nullptr;   // host address
  unsigned* dErrMapped = nullptr;  // device alias
  // last uploaded parameter sets (skips the upload + sync when a caller repeats the same parameters)
  std::vector<fhp_params> lastParams;
};

namespace {

int fail(const fhp_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}
#define CU(call)                                                                                       \
  do {                                                                                                 \
    cudaError_t e__ = (call);                                                                          \
    if (e__ != cudaSuccess) return fail(ctx, FHP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
  } while (0)

void free_owned(fhp_ctx* ctx) {
  for (float* p : ctx->owned) cudaFree(p);
  ctx->owned.clear();
  ctx->ownedBytes = 0;
}

int set_geometry(fhp_ctx* ctx, const fhp_map_desc* d) {
  if (!d || d->rows <= 0 || d->cols <= 0 || !(d->resolution > 0) || d->n_layers < 0 ||
      d->start_row < 0 || d->start_col < 0 || d->start_row >= d->rows || d->start_col >= d->cols)
    return fail(ctx, FHP_ERR_INVALID_ARG, "invalid map description");
  MapDev& M = ctx->M;
  M.rows = d->rows;
  M.cols = d->cols;
  M.start0 = d->start_row;
  M.start1 = d->start_col;
  M.res = d->resolution;
  M.rinv = 1.0 / d->resolution;
  M.lenx = d->length_x;
  M.leny = d->length_y;
  M.posx = d->pos_x;
  M.posy = d->pos_y;
  M.fastOk = (static_cast<long long>(d->rows) * d->cols < (1LL << 31)) ? 1 : 0;
  M.pad0 = 0;
  M.halfx = 0.5 * M.lenx;                        // getVectorToOrigin
  M.halfy = 0.5 * M.leny;
  M.fcx = M.posx + (M.halfx - 0.5 * M.res);      // mapPosition + getVectorToFirstCell
  M.fcy = M.posy + (M.halfy - 0.5 * M.res);
  ctx->haveGeo = false;
  ctx->maskValid = false;  // every map update passes through here
  return FHP_OK;
}


// ---- host -> device map upload ---------------------------------------------------------------------------------------
// Pageable source memory (an Eigen matrix, a std::vector<float> of a ROS message, wire bytes) cannot be DMA'd in
// place: a plain cudaMemcpy stages it through the driver's single-threaded bounce buffer.  Large layers instead go
// through two pinned chunks owned by the context: chunk i+1 is filled by a few host threads while chunk i is in
// flight, so the upload runs at min(host memcpy, PCIe) rather than at their sum.  Already-pinned sources and small
// layers are copied directly.
constexpr size_t kSmallBytes = 256u << 10;  // fhp_plan* calls whose buffers fit here take the zero-copy latency path
constexpr size_t kPinChunk = 8u << 20;
constexpr size_t kDirectBytes = 1u << 20;
constexpr int kCopyThreads = 6;

inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();
#endif
}

int upload_layer(fhp_ctx* ctx, float* dst, const void* src, size_t bytes, cudaStream_t s) {
  bool pinned = false;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, src) == cudaSuccess) pinned = (at.type == cudaMemoryTypeHost);
  cudaGetLastError();  // unregistered host memory reports an error on older drivers; not ours
  if (pinned || bytes <= kDirectBytes) {
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s));
    return FHP_OK;
  }
  for (int b = 0; b < 2; ++b)
    if (!ctx->pin[b]) {
      CU(cudaHostAlloc(&ctx->pin[b], kPinChunk, cudaHostAllocDefault));
      CU(cudaEventCreateWithFlags(&ctx->pinEv[b], cudaEventDisableTiming));
    }
  // Helper threads live for the whole layer (spawning them per chunk costs as much as the copy of a chunk): chunk c is
  // released to them through `go`, each copies its slice into the pinned buffer and bumps `done`.
  const int nChunks = static_cast<int>((bytes + kPinChunk - 1) / kPinChunk);
  const int nThreads = std::max(1, std::min<int>(kCopyThreads, static_cast<int>(std::thread::hardware_concurrency()) / 2));
  std::atomic<int> go{-1}, done{0};
  auto slice = [&](int c, int t) {
    const size_t off = static_cast<size_t>(c) * kPinChunk, n = std::min(kPinChunk, bytes - off);
    const size_t per = (n / nThreads + 63) & ~static_cast<size_t>(63);
    const size_t lo = std::min(n, per * t), hi = (t == nThreads - 1) ? n : std::min(n, per * (t + 1));
    if (hi > lo) std::memcpy(static_cast<char*>(ctx->pin[c & 1]) + lo, static_cast<const char*>(src) + off + lo, hi - lo);
  };
  std::vector<std::thread> helpers;
  for (int t = 1; t < nThreads; ++t)
    helpers.emplace_back([&, t] {
      for (int c = 0; c < nChunks; ++c) {
        while (go.load(std::memory_order_acquire) < c) cpu_relax();
        slice(c, t);
        done.fetch_add(1, std::memory_order_release);
      }
    });
  int rc = FHP_OK;
  for (int c = 0; c < nChunks; ++c) {
    const int b = c & 1;
    const size_t off = static_cast<size_t>(c) * kPinChunk, n = std::min(kPinChunk, bytes - off);
    cudaError_t e = cudaSuccess;
    if (ctx->pinUsed[b]) e = cudaEventSynchronize(ctx->pinEv[b]);  // the DMA that last read this buffer is done
    go.store(c, std::memory_order_release);  // (released even after an error so that the helpers run out)
    slice(c, 0);
    while (done.load(std::memory_order_acquire) < (c + 1) * (nThreads - 1)) cpu_relax();
    if (e == cudaSuccess && rc == FHP_OK) {
      e = cudaMemcpyAsync(reinterpret_cast<char*>(dst) + off, ctx->pin[b], n, cudaMemcpyHostToDevice, s);
      if (e == cudaSuccess) e = cudaEventRecord(ctx->pinEv[b], s);
      ctx->pinUsed[b] = true;
    }
    if (e != cudaSuccess && rc == FHP_OK) rc = fail(ctx, FHP_ERR_CUDA, std::string("map upload: ") + cudaGetErrorString(e));
  }
  for (std::thread& t : helpers) t.join();
  return rc;
}

int bind_layers(fhp_ctx* ctx);

// Common tail of fhp_set_map / fhp_set_map_msg / fhp_set_map_wire: geometry, device buffers (reused when the layer
// count and size are unchanged — the steady state of a robot's periodic map updates), upload, layer binding.
int ingest_host(fhp_ctx* ctx, const fhp_map_desc& d, const std::vector<std::string>& names,
                const std::vector<const void*>& ptrs) {
  for (const void* p : ptrs)
    if (!p) return fail(ctx, FHP_ERR_INVALID_ARG, "null layer");
  if (d.rows <= 0 || d.cols <= 0 || !(d.resolution > 0) || d.n_layers < 0 || d.start_row < 0 || d.start_col < 0 ||
      d.start_row >= d.rows || d.start_col >= d.cols)
    return fail(ctx, FHP_ERR_INVALID_ARG, "invalid map description");
  CU(cudaDeviceSynchronize());  // plans on caller streams (fhp_plan_device) may still read the buffers reused below
  const size_t bytes = static_cast<size_t>(d.rows) * d.cols * sizeof(float);
  // New buffers are allocated BEFORE anything of the current map is touched: an allocation failure leaves the previous
  // map fully in place (same-size updates, the steady state of a robot's periodic map, reuse the buffers).
  std::vector<float*> fresh;
  const bool reuse = ctx->owned.size() == names.size() && ctx->ownedBytes == bytes;
  if (!reuse) {
    for (size_t i = 0; i < names.size(); ++i) {
      float* p = nullptr;
      const cudaError_t e = cudaMalloc(&p, bytes);
      if (e != cudaSuccess) {
        for (float* q : fresh) cudaFree(q);
        return fail(ctx, FHP_ERR_CUDA, std::string("map allocation: ") + cudaGetErrorString(e));
      }
      fresh.push_back(p);
    }
  }
  ctx->haveMap = false;  // from here on the old content is being replaced
  ctx->layerPtrs.clear();
  ctx->M.trav = ctx->M.elev = nullptr;
  if (!reuse) {
    free_owned(ctx);
    ctx->owned = fresh;
    ctx->ownedBytes = bytes;
  }
  int rc = set_geometry(ctx, &d);
  if (rc != FHP_OK) return rc;
  ctx->layerNames = names;
  ctx->layerPtrs.assign(ctx->owned.begin(), ctx->owned.end());
  for (size_t i = 0; i < names.size(); ++i) {
    rc = upload_layer(ctx, ctx->owned[i], ptrs[i], bytes, ctx->stream);
    if (rc != FHP_OK) return rc;
  }
  CU(cudaStreamSynchronize(ctx->stream));
  return bind_layers(ctx);
}

int bind_layers(fhp_ctx* ctx) {
  ctx->M.trav = nullptr;
  ctx->M.elev = nullptr;
  for (size_t i = 0; i < ctx->layerNames.size(); ++i) {
    if (ctx->layerNames[i] == "traversability") ctx->M.trav = ctx->layerPtrs[i];
    if (ctx->layerNames[i] == "elevation") ctx->M.elev = ctx->layerPtrs[i];
  }
  if (!ctx->M.trav || !ctx->M.elev) {
    ctx->haveMap = false;
    return fail(ctx, FHP_ERR_MISSING_LAYER, "map needs layers \"traversability\" and \"elevation\"");
  }
  ctx->haveMap = true;
  return FHP_OK;
}

// Derive what FootholdPlanner::initialize computes from the rosparams (FP.cpp:317-421), with the reference's
// float / double member types (FootholdPlanner.hpp:609-697).
void derive_params(const fhp_params& p, ParDev& o, bool forceHeading = false) {
  const float lengthBase = p.kin_length;                 // FP.cpp:340
  const float widthBase = p.kin_width + p.kin_l1 * 2;    // FP.cpp:341 (float arithmetic)
  const float skew = p.skew;
  o.stancex[0] = lengthBase * 0.5;   o.stancey[0] = -widthBase * 0.5;  // RF  FP.cpp:350-351
  o.stancex[1] = -lengthBase * 0.5;  o.stancey[1] = -widthBase * 0.5;  // RH
  o.stancex[2] = -lengthBase * 0.5;  o.stancey[2] = widthBase * 0.5;   // LH
  o.stancex[3] = lengthBase * 0.5;   o.stancey[3] = widthBase * 0.5;   // LF
  if (p.gait == FHP_GAIT_CUSTOM) {
    for (int f = 0; f < 4; ++f) {
      o.biasx[f] = p.bias[f][0];
      o.biasy[f] = p.bias[f][1];
    }
  } else if (p.gait == FHP_GAIT_WALK) {  // extension, see include/fhp.h
    const double sl = p.step_length;
    o.biasx[0] = 0.5 * lengthBase + 0.375 * sl;   o.biasy[0] = -0.5 * widthBase;
    o.biasx[1] = -0.5 * lengthBase + 0.125 * sl;  o.biasy[1] = -0.5 * widthBase;
    o.biasx[2] = -0.5 * lengthBase - 0.375 * sl;  o.biasy[2] = 0.5 * widthBase;
    o.biasx[3] = 0.5 * lengthBase - 0.125 * sl;   o.biasy[3] = 0.5 * widthBase;
  } else if (p.rf_first) {  // FP.cpp:403-411
    o.biasx[0] = 0.5 * lengthBase + skew;   o.biasy[0] = -0.5 * widthBase;
    o.biasx[1] = -0.5 * lengthBase - skew;  o.biasy[1] = -0.5 * widthBase;
    o.biasx[2] = -0.5 * lengthBase + skew;  o.biasy[2] = 0.5 * widthBase;
    o.biasx[3] = 0.5 * lengthBase - skew;   o.biasy[3] = 0.5 * widthBase;
  } else {  // FP.cpp:412-421
    o.biasx[0] = 0.5 * lengthBase - skew;   o.biasy[0] = -0.5 * widthBase;
    o.biasx[1] = -0.5 * lengthBase + skew;  o.biasy[1] = -0.5 * widthBase;
    o.biasx[2] = -0.5 * lengthBase - skew;  o.biasy[2] = 0.5 * widthBase;
    o.biasx[3] = 0.5 * lengthBase + skew;   o.biasy[3] = 0.5 * widthBase;
  }
  o.stepLength = p.step_length;
  o.halfStep = p.step_length / 2;     // float / int -> float (FP.cpp:2693)
  o.R = p.search_radius;
  o.halfR = 0.5 * p.search_radius;    // 0.5*radius (FP.cpp:2502)
  o.rectLx = p.search_radius * 2;     // float (FP.cpp:1616)
  o.rectLy = p.search_radius;
  o.r = p.foot_radius;
  o.r2 = std::pow(o.r, 2);            // CircleIterator: pow(radius, 2)
  o.R2 = o.R * o.R;                   // SpiralIterator: radius * radius
  o.h = p.h;
  o.drift = p.drift_per_cycle;
  o.thrD = p.thr_default;
  o.thrC = p.thr_candidate;
  o.spiralMode = p.spiral_mode;
  // heading extension: rotate stance, bias table and rectangle vertex offsets (same expressions as the CPU checker)
  o.heading = (p.heading != 0.0 || forceHeading) ? 1 : 0;
  o.ch = std::cos(p.heading);
  o.sh = std::sin(p.heading);
  const double hR = 0.5 * p.search_radius;
  const double vx[4] = {o.R, o.R, -o.R, -o.R}, vy[4] = {hR, -hR, -hR, hR};
  for (int k = 0; k < 4; ++k) {
    o.rectVx[k] = o.heading ? vx[k] * o.ch - vy[k] * o.sh : vx[k];
    o.rectVy[k] = o.heading ? vx[k] * o.sh + vy[k] * o.ch : vy[k];
  }
  {  // isos_ (FP.hpp:677-703): longEdge = lengthBase + skew*2 in float arithmetic, stored in doubles (FP.cpp:384-394)
    const double longEdge = lengthBase + skew * 2, shortEdge = widthBase;
    o.isosLx = longEdge + o.rectLx;
    o.isosLy = shortEdge + o.rectLy;
  }
  for (int k = 0; k < 4; ++k) {  // edge k joins vertex k and vertex k-1 (PNPOLY's i and j)
    const int j = (k + 3) & 3;
    o.rectSlope[k] = (o.rectVx[j] - o.rectVx[k]) / (o.rectVy[j] - o.rectVy[k]);  // inf / nan for horizontal edges: never used
  }
  o.halfStepX = o.halfStep * o.ch;
  o.halfStepY = o.halfStep * o.sh;
  if (o.heading)
    for (int f = 0; f < 4; ++f) {
      const double sx = o.stancex[f], sy = o.stancey[f], bx = o.biasx[f], by = o.biasy[f];
      o.stancex[f] = sx * o.ch - sy * o.sh;
      o.stancey[f] = sx * o.sh + sy * o.ch;
      o.biasx[f] = bx * o.ch - by * o.sh;
      o.biasy[f] = bx * o.sh + by * o.ch;
    }
}

int sgn(int v) { return (v > 0) - (v < 0); }

// 2-D tiled tensor map over the column-major traversability layer: dim0 = rows (contiguous), dim1 = cols; box =
// (32/pack rows) x (nIter*pack cols) = one staged window; out-of-bounds elements are filled with NaN.
// Returns false (=> LDG staging path) when TMA's layout rules are not met: row pitch not a multiple of 16 bytes,
// base pointer not 16-byte aligned, or the driver entry point is unavailable.
bool build_tma(fhp_ctx* ctx) {
  const MapDev& M = ctx->M;
  const GeoDev& G = ctx->G;
  if (!M.trav || (M.rows % 4) != 0 || (reinterpret_cast<uintptr_t>(M.trav) & 15u) != 0) return false;
  if (G.WIN + 3 > 32 / G.pack) return false;  // the row-alignment shift (<= 3) must fit in the box
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || !fn)
      return false;
    encode = reinterpret_cast<encode_fn>(fn);
  }
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(M.rows), static_cast<cuuint64_t>(M.cols)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(M.rows) * sizeof(float)};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(32 / G.pack), static_cast<cuuint32_t>(G.nIter * G.pack)};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(&ctx->tmapTrav, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(M.trav), gdim, gstride, box,
                      estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA);
  return r == CUDA_SUCCESS;
}

// Tables that depend only on (footRadius, searchRadius, resolution): the SpiralIterator visiting order and the
// cell-centred foot-disc stencil.
int build_geo(fhp_ctx* ctx, float footRadius, float searchRadius, bool allowTrim, int stage) {
  if (ctx->haveGeo && ctx->geoFoot == footRadius && ctx->geoSearch == searchRadius && ctx->geoTrim == allowTrim &&
      ctx->geoStage == stage)
    return FHP_OK;
  const MapDev& M = ctx->M;
  GeoDev& G = ctx->G;
  std::memset(&G, 0, sizeof(G));
  const double R = searchRadius, r = footRadius, res = M.res;
  if (!(R >= 0) || !(r >= 0)) return fail(ctx, FHP_ERR_INVALID_ARG, "radii must be non-negative");
  const double nr = std::ceil(R / res);  // SpiralIterator: nRings_ = ceil(radius_/resolution_)
  if (nr > 15) return fail(ctx, FHP_ERR_UNSUPPORTED, "searchRadius/resolution > 15 rings is outside the window tables");
  G.nRings = static_cast<int>(nr);

  // ---- foot-disc stencil around a cell centre: offsets whose centre distance is <= footRadius.
  // The reference decides each cell with (pos(u+s) - pos(u))^2 summed in double; that equals |s|*res up to a
  // rounding error bounded by `tol`.  Offsets further than tol from the boundary are classified here once;
  // an offset inside the band cannot be classified without the per-cell rounding and is refused.
  const double Mmax = std::fmax(std::fabs(M.posx) + M.halfx, std::fabs(M.posy) + M.halfy) + 1.0;
  const double tol = 64.0 * DBL_EPSILON * Mmax * (r + res);
  const int span = static_cast<int>(std::floor(r / res)) + 2;
  int ra = 0, ns = 0;
  for (int sa = -span; sa <= span; ++sa)
    for (int sb = -span; sb <= span; ++sb) {
      const double d2 = static_cast<double>(sa * sa + sb * sb) * (res * res);
      bool in;
      if (sa == 0 && sb == 0) {
        in = true;  // distance is exactly 0
      } else if (d2 <= r * r - tol) {
        in = true;
      } else if (d2 >= r * r + tol) {
        in = false;
      } else {
        return fail(ctx, FHP_ERR_UNSUPPORTED,
                    "footRadius is within rounding distance of a cell-centre distance (stencil offset ambiguous)");
      }
      if (!in) continue;
      if (ns >= kMaxSten) return fail(ctx, FHP_ERR_UNSUPPORTED, "footRadius/resolution too large for the stencil table");
      G.sten[ns][0] = static_cast<int8_t>(sa);
      G.sten[ns][1] = static_cast<int8_t>(sb);
      ++ns;
      ra = std::max(ra, std::max(std::abs(sa), std::abs(sb)));
    }
  G.nSten = ns;
  G.ra = ra;
  // a member cell of the disc around a continuous centre is at most floor(r/res + 1/2) indices from the centre's cell
  // (see disc_window); the 1e-6 guard band covers the rounding of the index conversion at cell edges
  G.discK = static_cast<int>(std::floor(r / res + 0.5 + 1e-6));
  if (2 * G.discK + 1 > 8) G.discK = -1;
  // Rings whose every cell is farther than searchRadius from ANY centre that maps to the centre cell can never pass
  // SpiralIterator's last-two-rings distance filter: ring d holds offsets with |offset| >= d cells and the centre is
  // within sqrt(2)/2 cell of its cell centre (farther, and on the far side, when it lies outside the map).  Not
  // valid for the bug-compatible mode with a shifted start index, whose filter uses unrelated positions.
  int ringsEff = G.nRings;
  if (allowTrim)
    while (ringsEff > 0 && (static_cast<double>(ringsEff) - 0.70711 - 1e-6) * res > R) --ringsEff;
  G.HW = ringsEff + ra;
  if (G.HW < 1) G.HW = 1;
  // the staged window must also contain the foot disc of the (continuous) centre: index distance <= floor(r/res)+1
  if (G.HW < static_cast<int>(std::floor(r / res)) + 1) G.HW = static_cast<int>(std::floor(r / res)) + 1;
  G.WIN = 2 * G.HW + 1;
  if (G.WIN > 32)
    return fail(ctx, FHP_ERR_UNSUPPORTED, "(searchRadius+footRadius)/resolution needs a window wider than 32 cells");
  G.pack = G.WIN <= 8 ? 4 : (G.WIN <= 16 ? 2 : 1);
  if (stage == kStageMask) G.pack = 1;  // one entry per window column
  G.rpShift = G.pack == 4 ? 3 : (G.pack == 2 ? 4 : 5);
  G.nIter = (G.WIN + G.pack - 1) / G.pack;
  G.stride = 4 * G.nIter;
  // mask staging keeps the class-plane pair (e0, e1) per window column plus one word per column for the finished masks
  // (Planes, pack 0): 12 bytes per column, slot padded to a multiple of 8 bytes (WIN is odd)
  if (stage == kStageMask) G.stride = 3 * G.WIN + 1;

  // ---- SpiralIterator order: centre, then ring d = 1..nRings generated by the upstream walk and consumed from
  // the BACK of the ring vector (pointsRing_.back()/pop_back()).
  std::vector<std::pair<int, int>> table;
  table.push_back({0, 0});
  int lastTwo = -1;
  for (int d = 1; d <= ringsEff; ++d) {
    if (lastTwo < 0 && (d == G.nRings - 1 || d == G.nRings)) lastTwo = static_cast<int>(table.size());
    std::vector<std::pair<int, int>> ring;
    int x = d, y = 0;
    do {
      ring.push_back({x, y});
      const int nx = -sgn(y), ny = sgn(x);
      if (nx != 0 && static_cast<int>(std::sqrt(static_cast<double>((x + nx) * (x + nx) + y * y))) == d)
        x += nx;
      else if (ny != 0 && static_cast<int>(std::sqrt(static_cast<double>(x * x + (y + ny) * (y + ny)))) == d)
        y += ny;
      else {
        x += nx;
        y += ny;
      }
    } while (x != d || y != 0);
    for (auto it = ring.rbegin(); it != ring.rend(); ++it) table.push_back(*it);
  }
  if (lastTwo < 0) lastTwo = static_cast<int>(table.size());
  if (table.size() > static_cast<size_t>(kMaxCand)) return fail(ctx, FHP_ERR_UNSUPPORTED, "spiral table overflow");
  G.nCand = static_cast<int>(table.size());
  G.spiralWords = ((G.nCand * 2 + 15) / 16) * 4;
  G.lastTwoStart = lastTwo;
  for (int k = 0; k < G.nCand; ++k) {
    G.spiral[k][0] = static_cast<int8_t>(table[k].first);
    G.spiral[k][1] = static_cast<int8_t>(table[k].second);
    G.candRows[table[k].second + G.HW] |= 1u << (table[k].first + G.HW);
  }
  // ---- shared-memory layout of a plan_kernel block and the TMA descriptor
  const int warps = kPlanThreads / 32;
  const int planesEnd = G.spiralWords + warps * 32 * G.stride;
  G.boxWords = 32 * G.nIter;
  const int boxBudget = ctx->tmaBoxBudget;  // bytes of box buffers per warp (two rounds in flight)
  G.tmaBShift = 0;
  while (G.tmaBShift < 3 && (2 << (G.tmaBShift + 1)) * G.boxWords * 4 <= boxBudget) ++G.tmaBShift;
  G.tmaB = 1 << G.tmaBShift;
  G.mbarOff = planesEnd;                                   // multiple of 4 words -> 8-byte aligned
  G.boxOff = ((G.mbarOff + warps * 4 + 31) / 32) * 32;     // 128-byte aligned
  G.useTma = (stage == kStageTma && build_tma(ctx)) ? 1 : 0;
  G.stageMode = stage == kStageTma ? (G.useTma ? kStageTma : kStageLdg) : stage;
  G.winRows = G.useTma ? std::min(32 / G.pack, G.WIN + 3) : G.WIN;
  G.smemWords = G.useTma ? G.boxOff + warps * 2 * G.tmaB * G.boxWords : planesEnd;
  // foot discs of more than four cells (footRadius > resolution): their elevation cells are fetched into shared memory by
  // cp.async while the window is staged (plan_kernel) — 16 words per lane
  G.elevOff = -1;
  if (G.discK != 1) {
    G.elevOff = (G.smemWords + 3) & ~3;
    G.smemWords = G.elevOff + warps * 32 * 16;
  }
  ctx->geoStage = stage;
  ctx->geoFoot = footRadius;
  ctx->geoSearch = searchRadius;
  ctx->geoTrim = allowTrim;
  ctx->haveGeo = true;
  ctx->waveQueries[0] = ctx->waveQueries[1] = 0;
  return FHP_OK;
}

int ensure_err_word(fhp_ctx* ctx);

int prepare_params(fhp_ctx* ctx, const fhp_params* params, int32_t n_params, cudaStream_t s) {
  if (!params || n_params <= 0) return fail(ctx, FHP_ERR_INVALID_ARG, "params missing");
  for (int i = 1; i < n_params; ++i)
    if (params[i].foot_radius != params[0].foot_radius || params[i].search_radius != params[0].search_radius)
      return fail(ctx, FHP_ERR_INVALID_ARG, "all parameter sets of one call must share foot_radius and search_radius");
  bool allowTrim = true;  // see build_geo: ring trimming is exact unless the bug-compatible spiral meets a shifted start
  for (int i = 0; i < n_params; ++i)
    if (params[i].spiral_mode == FHP_SPIRAL_FAITHFUL && !(ctx->M.start0 == 0 && ctx->M.start1 == 0)) allowTrim = false;
  // the mask planes bake the two thresholds in: usable when every parameter set of the call shares them, the
  // classes nest (thrC <= thrD) and the layer holds no -inf cell (found while building, see mask_build_kernel)
  int stage = ctx->stagePref;
  if (stage == kStageMask) {
    if (!(params[0].thr_candidate <= params[0].thr_default)) stage = kStageTma;
    for (int i = 1; i < n_params; ++i)
      if (params[i].thr_default != params[0].thr_default || params[i].thr_candidate != params[0].thr_candidate)
        stage = kStageTma;
  }
  if (stage == kStageMask &&
      !(ctx->maskValid && ctx->maskThrD == params[0].thr_default && ctx->maskThrC == params[0].thr_candidate)) {
    int rc = ensure_err_word(ctx);
    if (rc != FHP_OK) return rc;
    if (ctx->asyncPlans) {  // a plan launched on a caller's stream (fhp_plan_device) may still be reading the planes
      CU(cudaDeviceSynchronize());
      ctx->asyncPlans = false;
    }
    const int rowWords = (ctx->M.rows + 31) / 32;
    CU(ctx->dMask.reserve(static_cast<size_t>(rowWords) * ctx->M.cols * sizeof(uint4)));
    ctx->hErr[1] = 0u;
    CU(launch_mask_build(ctx->M, params[0].thr_default, params[0].thr_candidate, static_cast<uint4*>(ctx->dMask.p), rowWords,
                         ctx->dErrMapped + 1, s));
    CU(cudaStreamSynchronize(s));  // later calls may run on other streams; the flag word is zero-copy host memory
    ctx->maskNegInf = (*static_cast<volatile unsigned*>(ctx->hErr + 1) & 1u) != 0u;
    ctx->maskThrD = params[0].thr_default;
    ctx->maskThrC = params[0].thr_candidate;
    ctx->maskValid = true;
    ctx->maskBuilds += 1;
    ctx->launches += 1;
  }
  if (stage == kStageMask && ctx->maskNegInf) stage = kStageTma;
  int rc = build_geo(ctx, params[0].foot_radius, params[0].search_radius, allowTrim, stage);
  if (rc != FHP_OK) return rc;
  ctx->anyHeading = ctx->forceHeading;
  for (int i = 0; i < n_params; ++i) {
    if (!std::isfinite(params[i].heading)) return fail(ctx, FHP_ERR_INVALID_ARG, "params.heading is not finite");
    if (params[i].heading != 0.0) ctx->anyHeading = true;
  }
  if (ctx->dPars.p && ctx->lastParams.size() == static_cast<size_t>(n_params) &&
      std::memcmp(ctx->lastParams.data(), params, sizeof(fhp_params) * n_params) == 0)
    return FHP_OK;  // same parameter sets as the previous call: the derived table is already on the device
  if (ctx->asyncPlans) {  // the parameter table is rewritten in place: wait for plans on caller streams
    CU(cudaDeviceSynchronize());
    ctx->asyncPlans = false;
  }
  std::vector<ParDev> h(n_params);
  for (int i = 0; i < n_params; ++i) derive_params(params[i], h[i], ctx->forceHeading);
  ctx->hPar0 = h[0];
  CU(ctx->dPars.reserve(sizeof(ParDev) * n_params));
  CU(cudaMemcpyAsync(ctx->dPars.p, h.data(), sizeof(ParDev) * n_params, cudaMemcpyHostToDevice, s));
  CU(cudaStreamSynchronize(s));  // h goes out of scope
  ctx->lastParams.assign(params, params + n_params);
  return FHP_OK;
}

// Device alias of a page-locked, mapped host pointer (cudaHostAlloc / cudaHostRegister), or NULL for anything else.
// (cudaPointerGetAttributes reports unregistered host memory as such instead of failing.)
void* mapped_device_ptr(const void* h) {
  cudaPointerAttributes at;
  if (!h || cudaPointerGetAttributes(&at, h) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}

// Small batches are validated on the host before anything is launched; large ones (a walk over every query would
// cost as much as a tenth of the plan itself) by the kernel, which reports kErrParam.
bool param_ids_ok(const fhp_query* queries, int64_t n_queries, int32_t n_params) {
  if (n_queries > 4096) return true;
  for (int64_t i = 0; i < n_queries; ++i)
    if (queries[i].param_id < 0 || queries[i].param_id >= n_params) return false;
  return true;
}

int check_dev_err(fhp_ctx* ctx) {
  const unsigned e = ctx->hErr ? *static_cast<volatile unsigned*>(ctx->hErr) : 0u;
  if (e & kErrParam) {
    *ctx->hErr = 0;
    return fail(ctx, FHP_ERR_INVALID_ARG, "query.param_id out of range (reported by the kernel; the response is not valid)");
  }
  if (e) {
    *ctx->hErr = 0;
    char b[160];
    snprintf(b, sizeof b, "kernel reported geometry outside its window tables (bits 0x%x)", e);
    return fail(ctx, FHP_ERR_UNSUPPORTED, b);
  }
  return FHP_OK;
}

int ensure_err_word(fhp_ctx* ctx) {
  if (!ctx->hErr) {
    void* h = nullptr;
    CU(cudaHostAlloc(&h, 2 * sizeof(unsigned), cudaHostAllocMapped));  // [0] kernel error bits, [1] mask-build flags
    ctx->hErr = static_cast<unsigned*>(h);
    ctx->hErr[0] = ctx->hErr[1] = 0;
    void* d = nullptr;
    CU(cudaHostGetDevicePointer(&d, h, 0));
    ctx->dErrMapped = static_cast<unsigned*>(d);
  }
  return FHP_OK;
}

}  // namespace

extern "C" {

const char* fhp_version(void) { return "libfhp 0.1 (sm_100a)"; }

void fhp_default_params(fhp_params* p) {
  if (!p) return;
  std::memset(p, 0, sizeof(*p));
  p->foot_radius = 0.02f;
  p->search_radius = 0.1f;
  p->step_length = 0.18f;
  p->skew = 0.04f;
  p->thr_default = 0.9f;
  p->thr_candidate = 0.7f;
  p->kin_length = 0.4387f;
  p->kin_width = 0.175f;
  p->kin_l1 = 0.037f;
  p->rf_first = 0;
  p->gait = FHP_GAIT_TROT;
  p->spiral_mode = FHP_SPIRAL_WRAP_CORRECT;
  p->h = 0.01;
  p->drift_per_cycle = -0.007;
}

int fhp_create(int device, fhp_ctx** out) {
  if (!out) return FHP_ERR_INVALID_ARG;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    return FHP_ERR_CUDA;
  }
  if (device < 0 || device >= n) {
    g_create_error = "device index out of range";
    return FHP_ERR_INVALID_ARG;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    g_create_error = cudaGetErrorString(e);
    return FHP_ERR_CUDA;
  }
  if (prop.major != 10) {
    g_create_error = "libfhp is built for sm_100a only; device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor);
    return FHP_ERR_CUDA;
  }
  e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    g_create_error = cudaGetErrorString(e);
    return FHP_ERR_CUDA;
  }
  fhp_ctx* ctx = new fhp_ctx();
  ctx->device = device;
  std::memset(&ctx->tmapTrav, 0, sizeof(ctx->tmapTrav));
  if (const char* dt = std::getenv("FHP_DISABLE_TMA"))
    if (dt[0] == '1') ctx->stagePref = kStageLdg;
  if (const char* v = std::getenv("FHP_CHUNK_DIV")) ctx->chunkDiv = std::max(1, atoi(v));
  if (const char* v = std::getenv("FHP_QPW_SHIFT")) ctx->qpwShiftForce = std::max(0, std::min(3, atoi(v)));
  if (const char* v = std::getenv("FHP_TMA_BOX_BUDGET")) ctx->tmaBoxBudget = std::max(1024, atoi(v));
  if (const char* v = std::getenv("FHP_FORCE_HEADING")) ctx->forceHeading = v[0] == '1';
  if (const char* v = std::getenv("FHP_NO_SPEC")) ctx->noSpec = v[0] == '1';
  if (const char* v = std::getenv("FHP_NO_ONEPAR")) ctx->noOnePar = v[0] == '1';
  if (const char* st = std::getenv("FHP_STAGING")) {
    if (!std::strcmp(st, "mask")) ctx->stagePref = kStageMask;
    else if (!std::strcmp(st, "tma")) ctx->stagePref = kStageTma;
    else if (!std::strcmp(st, "ldg")) ctx->stagePref = kStageLdg;
  }
  e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    g_create_error = cudaGetErrorString(e);
    delete ctx;
    return FHP_ERR_CUDA;
  }
  *out = ctx;
  return FHP_OK;
}

void fhp_destroy(fhp_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  free_owned(ctx);
  for (DevBuf* b : {&ctx->dPars, &ctx->dQ, &ctx->dOutN, &ctx->dOutC, &ctx->dOutD, &ctx->dSt, &ctx->dErr, &ctx->dA, &ctx->dB, &ctx->dC2,
                    &ctx->dD, &ctx->dMask, &ctx->dCompact, &ctx->dOffs})
    b->release();
  if (ctx->hErr) cudaFreeHost(ctx->hErr);
  if (ctx->hTotals) cudaFreeHost(ctx->hTotals);
  if (ctx->hSmall) cudaFreeHost(ctx->hSmall);
  for (cudaEvent_t e : ctx->chunkEv) cudaEventDestroy(e);
  if (ctx->copyStream) cudaStreamDestroy(ctx->copyStream);
  for (int b = 0; b < 2; ++b) {
    if (ctx->pin[b]) cudaFreeHost(ctx->pin[b]);
    if (ctx->pinEv[b]) cudaEventDestroy(ctx->pinEv[b]);
  }
  cudaStreamDestroy(ctx->stream);
  if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
  delete ctx;
}

const char* fhp_last_error(const fhp_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int fhp_set_map(fhp_ctx* ctx, const fhp_map_desc* desc, const float* const* layers) {
  if (!ctx) return FHP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  if (!layers || !desc || !desc->layer_names) return fail(ctx, FHP_ERR_INVALID_ARG, "null map");
  std::vector<std::string> names;
  std::vector<const void*> ptrs;
  for (int i = 0; i < desc->n_layers; ++i) {
    if (!layers[i] || !desc->layer_names[i]) return fail(ctx, FHP_ERR_INVALID_ARG, "null layer");
    names.push_back(desc->layer_names[i]);
    ptrs.push_back(layers[i]);
  }
  return ingest_host(ctx, *desc, names, ptrs);
}

// GridMapRosConverter::fromMessage (grid_map_ros 1.6.x) on the message fields, without the Eigen copies.
int fhp_set_map_msg(fhp_ctx* ctx, const fhp_gridmap_msg* m) {
  if (!ctx) return FHP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  if (!m || m->n_layers < 0 || m->n_data < 0 || (m->n_layers > 0 && !m->layers) || (m->n_data > 0 && !m->data))
    return fail(ctx, FHP_ERR_INVALID_ARG, "null grid map message");
  // GridMap::setGeometry: size = round(length / resolution), length = size * resolution, start index reset
  if (!(m->resolution > 0) || !std::isfinite(m->resolution) || !std::isfinite(m->length_x) || !std::isfinite(m->length_y))
    return fail(ctx, FHP_ERR_INVALID_ARG, "grid map message: resolution / length not finite and positive");
  const double s0 = std::round(m->length_x / m->resolution), s1 = std::round(m->length_y / m->resolution);
  if (!(s0 >= 1 && s1 >= 1 && s0 < 2147483647.0 && s1 < 2147483647.0))
    return fail(ctx, FHP_ERR_INVALID_ARG, "grid map message: empty or oversized map");
  fhp_map_desc d;
  std::memset(&d, 0, sizeof(d));
  d.rows = static_cast<int32_t>(s0);
  d.cols = static_cast<int32_t>(s1);
  d.resolution = m->resolution;
  d.length_x = static_cast<double>(d.rows) * m->resolution;
  d.length_y = static_cast<double>(d.cols) * m->resolution;
  d.pos_x = m->pos_x;
  d.pos_y = m->pos_y;
  d.start_row = m->outer_start_index;  // setStartIndex(Index(outer_start_index, inner_start_index))
  d.start_col = m->inner_start_index;
  d.n_layers = m->n_layers;
  if (m->n_layers != m->n_data)
    return fail(ctx, FHP_ERR_INVALID_ARG, "Different number of layers and data in grid map message.");
  std::vector<std::string> names;
  std::vector<const void*> ptrs;
  const std::string colLabel = "column_index", rowLabel = "row_index";
  bool labelWarning = false;
  for (int i = 0; i < m->n_layers; ++i) {
    const fhp_multiarray& a = m->data[i];
    if (!m->layers[i].p && m->layers[i].n) return fail(ctx, FHP_ERR_INVALID_ARG, "null layer name");
    if (a.n_dims < 2 || !a.dim0_label.p || !a.dim1_label.p)
      return fail(ctx, FHP_ERR_INVALID_ARG, "grid map message: a data array has fewer than two layout dimensions");
    const std::string l0(a.dim0_label.p, a.dim0_label.n);
    if (l0 == rowLabel)  // isRowMajor(m) != EigenType::IsRowMajor
      return fail(ctx, FHP_ERR_INVALID_ARG,
                  "multiArrayMessageToMatrixEigen() failed because the storage order is not compatible.");
    // any other label: upstream's isRowMajor() logs "isRowMajor() failed because layout label is not set correctly." and
    // answers false, i.e. the array is then read as column-major like a correctly labelled one — so it is here
    if (l0 != colLabel) labelWarning = true;
    // getCols = dim[0].size, getRows = dim[1].size; GridMap::add asserts the size equals the geometry's
    if (static_cast<int64_t>(a.dim1_size) != d.rows || static_cast<int64_t>(a.dim0_size) != d.cols ||
        a.n != static_cast<uint64_t>(d.rows) * static_cast<uint64_t>(d.cols) || !a.data)
      return fail(ctx, FHP_ERR_INVALID_ARG, "grid map message: layer \"" + std::string(m->layers[i].p, m->layers[i].n) +
                                                "\" does not have the size of the map geometry");
    names.emplace_back(m->layers[i].p ? m->layers[i].p : "", m->layers[i].n);
    ptrs.push_back(a.data);
  }
  if (d.start_row >= d.rows || d.start_col >= d.cols)
    return fail(ctx, FHP_ERR_INVALID_ARG, "grid map message: start index outside the map");
  const int rc = ingest_host(ctx, d, names, ptrs);
  if (rc == FHP_OK && labelWarning)  // not an error (FHP_OK): the text upstream logs, for hosts that want to log it too
    ctx->err = "isRowMajor() failed because layout label is not set correctly.";
  return rc;
}

namespace {
// bounds-checked little-endian reader over ROS1 wire bytes
struct WireReader {
  const uint8_t* p;
  uint64_t n, o = 0;
  bool ok = true;
  WireReader(const uint8_t* p_, uint64_t n_) : p(p_), n(n_) {}
  bool need(uint64_t k) {
    if (!ok || k > n - o) ok = false;
    return ok;
  }
  uint32_t u32() {
    uint32_t v = 0;
    if (need(4)) {
      std::memcpy(&v, p + o, 4);
      o += 4;
    }
    return v;
  }
  uint16_t u16() {
    uint16_t v = 0;
    if (need(2)) {
      std::memcpy(&v, p + o, 2);
      o += 2;
    }
    return v;
  }
  double f64() {
    double v = 0;
    if (need(8)) {
      std::memcpy(&v, p + o, 8);
      o += 8;
    }
    return v;
  }
  fhp_str str() {
    fhp_str s = {nullptr, 0};
    const uint32_t len = u32();
    if (need(len)) {
      s.p = reinterpret_cast<const char*>(p + o);
      s.n = len;
      o += len;
    }
    return s;
  }
  uint32_t count(uint64_t minElemBytes) {  // array length prefix, bounded by the bytes that are actually there
    const uint32_t c = u32();
    if (ok && static_cast<uint64_t>(c) * minElemBytes > n - o) ok = false;
    return ok ? c : 0;
  }
};
}  // namespace

// grid_map_msgs/GridMap, ROS1 serialisation: GridMapInfo info {Header {uint32 seq, time stamp, string frame_id},
// float64 resolution, length_x, length_y, geometry_msgs/Pose pose}, string[] layers, string[] basic_layers,
// std_msgs/Float32MultiArray[] data {MultiArrayLayout {MultiArrayDimension[] dim {string label, uint32 size,
// uint32 stride}, uint32 data_offset}, float32[] data}, uint16 outer_start_index, uint16 inner_start_index.
int fhp_set_map_wire(fhp_ctx* ctx, const uint8_t* bytes, uint64_t len) {
  if (!ctx) return FHP_ERR_INVALID_ARG;
  if (!bytes) return fail(ctx, FHP_ERR_INVALID_ARG, "null message bytes");
  WireReader r(bytes, len);
  fhp_gridmap_msg m;
  std::memset(&m, 0, sizeof(m));
  r.u32();  // header.seq
  r.u32();  // header.stamp.secs
  r.u32();  // header.stamp.nsecs
  r.str();  // header.frame_id
  m.resolution = r.f64();
  m.length_x = r.f64();
  m.length_y = r.f64();
  m.pos_x = r.f64();  // pose.position.x
  m.pos_y = r.f64();
  r.f64();            // pose.position.z
  for (int i = 0; i < 4; ++i) r.f64();  // pose.orientation
  std::vector<fhp_str> layers;
  const uint32_t nl = r.count(4);
  for (uint32_t i = 0; i < nl && r.ok; ++i) layers.push_back(r.str());
  const uint32_t nb = r.count(4);
  for (uint32_t i = 0; i < nb && r.ok; ++i) r.str();  // basic_layers: not used by the path
  std::vector<fhp_multiarray> data;
  const uint32_t nd = r.count(12);
  for (uint32_t i = 0; i < nd && r.ok; ++i) {
    fhp_multiarray a;
    std::memset(&a, 0, sizeof(a));
    const uint32_t ndim = r.count(12);
    a.n_dims = ndim;
    for (uint32_t k = 0; k < ndim && r.ok; ++k) {
      const fhp_str label = r.str();
      const uint32_t size = r.u32();
      r.u32();  // stride
      if (k == 0) {
        a.dim0_label = label;
        a.dim0_size = size;
      } else if (k == 1) {
        a.dim1_label = label;
        a.dim1_size = size;
      }
    }
    r.u32();  // data_offset (upstream maps data.data() from element 0 regardless)
    const uint32_t nf = r.count(4);
    a.n = nf;
    if (r.need(static_cast<uint64_t>(nf) * 4)) {
      a.data = reinterpret_cast<const float*>(r.p + r.o);  // possibly unaligned: only ever memcpy'd
      r.o += static_cast<uint64_t>(nf) * 4;
    }
    data.push_back(a);
  }
  m.outer_start_index = r.u16();
  m.inner_start_index = r.u16();
  if (!r.ok) return fail(ctx, FHP_ERR_INVALID_ARG, "grid map wire message is truncated or inconsistent");
  if (r.o != len) return fail(ctx, FHP_ERR_INVALID_ARG, "grid map wire message has trailing bytes");
  m.n_layers = static_cast<int32_t>(layers.size());
  m.n_data = static_cast<int32_t>(data.size());
  m.layers = layers.data();
  m.data = data.data();
  return fhp_set_map_msg(ctx, &m);
}

int fhp_map_info(const fhp_ctx* ctx, fhp_map_desc* out) {
  if (!ctx || !out) return FHP_ERR_INVALID_ARG;
  if (!ctx->haveMap) return fail(ctx, FHP_ERR_NO_MAP, "no map set");
  std::memset(out, 0, sizeof(*out));
  out->rows = ctx->M.rows;
  out->cols = ctx->M.cols;
  out->resolution = ctx->M.res;
  out->length_x = ctx->M.lenx;
  out->length_y = ctx->M.leny;
  out->pos_x = ctx->M.posx;
  out->pos_y = ctx->M.posy;
  out->start_row = ctx->M.start0;
  out->start_col = ctx->M.start1;
  out->n_layers = static_cast<int32_t>(ctx->layerNames.size());
  out->layer_names = nullptr;
  return FHP_OK;
}

int fhp_set_map_device(fhp_ctx* ctx, const fhp_map_desc* desc, const float* const* layers) {
  if (!ctx) return FHP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  if (!layers || !desc || !desc->layer_names) return fail(ctx, FHP_ERR_INVALID_ARG, "null map");
  int rc = set_geometry(ctx, desc);
  if (rc != FHP_OK) return rc;
  CU(cudaDeviceSynchronize());  // also plans launched on caller streams
  free_owned(ctx);
  ctx->layerNames.clear();
  ctx->layerPtrs.clear();
  ctx->haveMap = false;
  for (int i = 0; i < desc->n_layers; ++i) {
    if (!layers[i] || !desc->layer_names[i]) return fail(ctx, FHP_ERR_INVALID_ARG, "null layer");
    ctx->layerNames.push_back(desc->layer_names[i]);
    ctx->layerPtrs.push_back(layers[i]);
  }
  return bind_layers(ctx);
}

const float* fhp_map_layer_device(const fhp_ctx* ctx, const char* name) {
  if (!ctx || !name) return nullptr;
  for (size_t i = 0; i < ctx->layerNames.size(); ++i)
    if (ctx->layerNames[i] == name) return ctx->layerPtrs[i];
  return nullptr;
}

static long long wave_queries(fhp_ctx* ctx, bool wantCen) {  // occupancy query cached per geometry
  long long& w = ctx->waveQueries[wantCen ? 1 : 0];
  if (w == 0) w = plan_wave_queries(ctx->G, wantCen);
  return w;
}

// Chunk size of the host-buffer pipelines (FHP_CHUNK_DIV overrides the divisor for experiments).
static int64_t chunk_queries(fhp_ctx* ctx, bool wantCen) {
  return std::max<int64_t>(1024, wave_queries(ctx, wantCen) / ctx->chunkDiv);
}

struct StreamOut {  // fhp_plan_stream: device aliases of the caller's mapped host buffers
  fhp_point20* slots;
  uint32_t* mask;
  long long nq, q0;
};
static int plan_launch(fhp_ctx* ctx, int32_t n_params, const fhp_query* d_queries, int64_t n_queries,
                       int32_t gait_cycles, fhp_foothold* d_outN, fhp_foothold* d_outC, fhp_query_status* d_st,
                       cudaStream_t s, fhp_foothold* d_outD = nullptr, const StreamOut* so = nullptr) {
  PlanArgs A;
  A.sOut = so ? so->slots : nullptr;
  A.sMask = so ? so->mask : nullptr;
  A.sNq = so ? so->nq : 0;
  A.sQ0 = so ? so->q0 : 0;
  A.tmap = ctx->tmapTrav;
  A.M = ctx->M;
  A.mask = static_cast<const uint4*>(ctx->dMask.p);
  A.pars = static_cast<const ParDev*>(ctx->dPars.p);
  A.nPars = n_params;
  A.par0 = ctx->hPar0;
  A.q = d_queries;
  A.nq = n_queries;
  A.cycles = gait_cycles;
  {  // queries per warp.  Every warp pays the same instruction stream per gait cycle whatever the number of its busy quads,
    // but a GPU with few resident warps runs each of them faster.  Measured on C3 shards of 1,024 ... 32,768 queries
    // (profiles/r2_experiments.md): fastest is the smallest power of two that keeps the launch at or below HALF a
    // resident wave of warps (2,072 on a B200) — fewer queries per warp only pay while warps are what is scarce.
    const long long waveWarps = std::max<long long>(1, wave_queries(ctx, d_outC != nullptr) / 8);
    const long long limit = std::max<long long>(1, waveWarps / 2);
    const long long need = (n_queries + limit - 1) / limit;
    A.qpwShift = need <= 1 ? 0 : (need <= 2 ? 1 : (need <= 4 ? 2 : 3));
    if (ctx->qpwShiftForce >= 0) A.qpwShift = ctx->qpwShiftForce;
    if (so) A.qpwShift = 3;  // the streamed buffer format groups 8 queries per block
  }
  A.outN = d_outN;
  A.outC = d_outC;
  A.outD = d_outD;
  A.st = d_st;
  A.err = ctx->dErrMapped;
  {
    cudaError_t e = launch_plan(A, ctx->G, d_outC != nullptr, ctx->anyHeading, ctx->noSpec, ctx->noOnePar, s);
    if (e != cudaSuccess) {
      char b[256];
      snprintf(b, sizeof b, "plan kernel launch failed: %s (queries %lld, queries/warp %d, shared memory %d B)",
               cudaGetErrorString(e), static_cast<long long>(n_queries), 1 << A.qpwShift, ctx->G.smemWords * 4 + 128);
      return fail(ctx, FHP_ERR_CUDA, b);
    }
  }
  if (n_queries > 0) ctx->launches += 1;
  return FHP_OK;
}

static int plan_common(fhp_ctx* ctx, const fhp_params* params, int32_t n_params, const fhp_query* d_queries,
                       int64_t n_queries, int32_t gait_cycles, fhp_foothold* d_outN, fhp_foothold* d_outC,
                       fhp_query_status* d_st, cudaStream_t s) {
  int rc = prepare_params(ctx, params, n_params, s);
  if (rc != FHP_OK) return rc;
  rc = ensure_err_word(ctx);
  if (rc != FHP_OK) return rc;
  return plan_launch(ctx, n_params, d_queries, n_queries, gait_cycles, d_outN, d_outC, d_st, s);
}

int fhp_plan(fhp_ctx* ctx, const fhp_params* params, int32_t n_params, const fhp_query* queries, int64_t n_queries,
             int32_t gait_cycles, fhp_foothold* out_nominal, fhp_foothold* out_centroid, fhp_query_status* status) {
  fhp_plan_outputs out;
  out.nominal = out_nominal;
  out.centroid = out_centroid;
  out.default_chain = nullptr;
  out.status = status;
  return fhp_plan_ex(ctx, params, n_params, queries, n_queries, gait_cycles, &out);
}

static int plan_ex_one_geometry(fhp_ctx* ctx, const fhp_params* params, int32_t n_params, const fhp_query* queries,
                                int64_t n_queries, int32_t gait_cycles, const fhp_plan_outputs* outp);

// One call = one window geometry on the device (the tables depend on footRadius, searchRadius and the resolution).
// Parameter sets that differ in their radii are therefore planned group by group: the queries of each (footRadius,
// searchRadius) pair are gathered, planned in their own launch and their rows scattered back — the reference takes the
// radii per call as well.  (Host buffers only; the device-resident, compact and streamed entry points ask for one
// geometry per call.)
int fhp_plan_ex(fhp_ctx* ctx, const fhp_params* params, int32_t n_params, const fhp_query* queries, int64_t n_queries,
                int32_t gait_cycles, const fhp_plan_outputs* outp) {
  if (!ctx) return FHP_ERR_INVALID_ARG;
  bool mixed = false;
  if (params && n_params > 1)
    for (int i = 1; i < n_params; ++i)
      mixed = mixed || params[i].foot_radius != params[0].foot_radius || params[i].search_radius != params[0].search_radius;
  if (!mixed) return plan_ex_one_geometry(ctx, params, n_params, queries, n_queries, gait_cycles, outp);
  if (!outp || !queries || !outp->nominal || !outp->status || n_queries < 0 || gait_cycles < 0 || gait_cycles > 255)
    return fail(ctx, FHP_ERR_INVALID_ARG, "invalid plan arguments");
  for (int64_t i = 0; i < n_queries; ++i)
    if (queries[i].param_id < 0 || queries[i].param_id >= n_params) return fail(ctx, FHP_ERR_INVALID_ARG, "query.param_id out of range");
  const size_t R = FHP_ROWS_PER_QUERY(gait_cycles);
  std::vector<int> group(n_params, -1);  // parameter set -> geometry group
  int nGroups = 0;
  for (int i = 0; i < n_params; ++i) {
    for (int j = 0; j < i && group[i] < 0; ++j)
      if (params[j].foot_radius == params[i].foot_radius && params[j].search_radius == params[i].search_radius) group[i] = group[j];
    if (group[i] < 0) group[i] = nGroups++;
  }
  for (int g = 0; g < nGroups; ++g) {
    std::vector<fhp_params> gp;
    std::vector<int> remap(n_params, -1);
    for (int i = 0; i < n_params; ++i)
      if (group[i] == g) {
        remap[i] = static_cast<int>(gp.size());
        gp.push_back(params[i]);
      }
    std::vector<int64_t> idx;
    std::vector<fhp_query> gq;
    for (int64_t i = 0; i < n_queries; ++i)
      if (group[queries[i].param_id] == g) {
        idx.push_back(i);
        gq.push_back(queries[i]);
        gq.back().param_id = remap[queries[i].param_id];
      }
    if (gq.empty()) continue;
    const size_t n = gq.size();
    std::vector<fhp_foothold> gn(n * R), gc(outp->centroid ? n * R : 0), gd(outp->default_chain ? n * R : 0);
    std::vector<fhp_query_status> gs(n);
    fhp_plan_outputs go;
    go.nominal = gn.data();
    go.centroid = outp->centroid ? gc.data() : nullptr;
    go.default_chain = outp->default_chain ? gd.data() : nullptr;
    go.status = gs.data();
    const int rc = plan_ex_one_geometry(ctx, gp.data(), static_cast<int32_t>(gp.size()), gq.data(), static_cast<int64_t>(n),
                                        gait_cycles, &go);
    if (rc != FHP_OK) return rc;
    for (size_t k = 0; k < n; ++k) {
      const int64_t q = idx[k];
      std::memcpy(outp->nominal + q * R, gn.data() + k * R, sizeof(fhp_foothold) * R);
      if (outp->centroid) std::memcpy(outp->centroid + q * R, gc.data() + k * R, sizeof(fhp_foothold) * R);
      if (outp->default_chain) std::memcpy(outp->default_chain + q * R, gd.data() + k * R, sizeof(fhp_foothold) * R);
      outp->status[q] = gs[k];
    }
  }
  return FHP_OK;
}

static int plan_ex_one_geometry(fhp_ctx* ctx, const fhp_params* params, int32_t n_params, const fhp_query* queries,
                                int64_t n_queries, int32_t gait_cycles, const fhp_plan_outputs* outp) {
  if (!ctx) return FHP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!outp) return fail(ctx, FHP_ERR_INVALID_ARG, "null outputs");
  fhp_foothold* out_nominal = outp->nominal;
  fhp_foothold* out_centroid = outp->centroid;
  fhp_foothold* out_default = outp->default_chain;
  fhp_query_status* status = outp->status;
  if (n_queries < 0 || gait_cycles < 0 || gait_cycles > 255 || (n_queries > 0 && (!queries || !out_nominal || !status)))
    return fail(ctx, FHP_ERR_INVALID_ARG, "invalid plan arguments (gait_cycles is a uint8 in the service request)");
  if (!ctx->haveMap) return fail(ctx, FHP_ERR_NO_MAP, "fhp_plan before fhp_set_map");
  if (!param_ids_ok(queries, n_queries, n_params)) return fail(ctx, FHP_ERR_INVALID_ARG, "query.param_id out of range");
  CU(cudaSetDevice(ctx->device));
  if (n_queries == 0) return FHP_OK;
  cudaStream_t s = ctx->stream;
  const size_t R = FHP_ROWS_PER_QUERY(gait_cycles);
  const size_t qB = sizeof(fhp_query) * n_queries, oB = sizeof(fhp_foothold) * R * n_queries,
               sB = sizeof(fhp_query_status) * n_queries;
  // Latency path: everything fits the mapped scratch -> no memcpy calls, the kernel talks to host memory directly.
  {
    const size_t nOut = 1 + (out_centroid || out_default ? 1 : 0) + (out_default ? 1 : 0);
    const size_t need = qB + nOut * oB + sB;
    if (need <= kSmallBytes) {
      int rc = prepare_params(ctx, params, n_params, s);
      if (rc != FHP_OK) return rc;
      rc = ensure_err_word(ctx);
      if (rc != FHP_OK) return rc;
      if (!ctx->hSmall) {
        CU(cudaHostAlloc(&ctx->hSmall, kSmallBytes, cudaHostAllocMapped));
        CU(cudaHostGetDevicePointer(&ctx->dSmall, ctx->hSmall, 0));
      }
      char* h = static_cast<char*>(ctx->hSmall);
      char* d = static_cast<char*>(ctx->dSmall);
      size_t off = 0;
      const size_t oQ = off;  off += qB;
      const size_t oN = off;  off += oB;
      const size_t oC = off;  off += (nOut >= 2 ? oB : 0);
      const size_t oD = off;  off += (nOut >= 3 ? oB : 0);
      const size_t oS = off;
      std::memcpy(h + oQ, queries, qB);
      rc = plan_launch(ctx, n_params, reinterpret_cast<const fhp_query*>(d + oQ), n_queries, gait_cycles,
                       reinterpret_cast<fhp_foothold*>(d + oN), nOut >= 2 ? reinterpret_cast<fhp_foothold*>(d + oC) : nullptr,
                       reinterpret_cast<fhp_query_status*>(d + oS), s,
                       nOut >= 3 ? reinterpret_cast<fhp_foothold*>(d + oD) : nullptr);
      if (rc != FHP_OK) return rc;
      CU(cudaStreamSynchronize(s));
      std::memcpy(out_nominal, h + oN, oB);
      if (out_centroid) std::memcpy(out_centroid, h + oC, oB);
      if (out_default) std::memcpy(out_default, h + oD, oB);
      std::memcpy(status, h + oS, sB);
      return check_dev_err(ctx);
    }
  }
  CU(ctx->dQ.reserve(qB));
  CU(ctx->dOutN.reserve(oB));
  if (out_centroid || out_default) CU(ctx->dOutC.reserve(oB));  // the default-chain kernel variant also writes centroid rows
  if (out_default) CU(ctx->dOutD.reserve(oB));
  CU(ctx->dSt.reserve(sB));
  // Host-buffer pipeline: the batch is cut into chunks (a third of a wave) that alternate between two streams, so the kernel
  // of chunk i+1 runs while the response rows of chunk i travel back over PCIe (the D2H of the rows is the
  // longest leg of a large batch).  Small batches are a single chunk = one launch.
  int rc = prepare_params(ctx, params, n_params, s);
  if (rc != FHP_OK) return rc;
  rc = ensure_err_word(ctx);
  if (rc != FHP_OK) return rc;
  const int64_t wave = chunk_queries(ctx, out_centroid != nullptr);
  int64_t nChunks = (n_queries + wave - 1) / wave;
  if (nChunks < 1) nChunks = 1;
  const int64_t per = ((n_queries + nChunks - 1) / nChunks + 31) / 32 * 32;  // whole blocks per chunk
  const fhp_query* dq = static_cast<const fhp_query*>(ctx->dQ.p);
  fhp_foothold* dN = static_cast<fhp_foothold*>(ctx->dOutN.p);
  fhp_foothold* dC = (out_centroid || out_default) ? static_cast<fhp_foothold*>(ctx->dOutC.p) : nullptr;
  fhp_foothold* dD = out_default ? static_cast<fhp_foothold*>(ctx->dOutD.p) : nullptr;
  fhp_query_status* dS = static_cast<fhp_query_status*>(ctx->dSt.p);
  int lane = 0;
  for (int64_t lo = 0; lo < n_queries; lo += per, lane ^= 1) {
    const int64_t n = std::min<int64_t>(per, n_queries - lo);
    cudaStream_t cs = lane ? ctx->stream2 : ctx->stream;
    CU(cudaMemcpyAsync(const_cast<fhp_query*>(dq) + lo, queries + lo, sizeof(fhp_query) * n, cudaMemcpyHostToDevice, cs));
    rc = plan_launch(ctx, n_params, dq + lo, n, gait_cycles, dN + lo * R, dC ? dC + lo * R : nullptr, dS + lo, cs,
                     dD ? dD + lo * R : nullptr);
    if (rc != FHP_OK) return rc;
    CU(cudaMemcpyAsync(out_nominal + lo * R, dN + lo * R, sizeof(fhp_foothold) * R * n, cudaMemcpyDeviceToHost, cs));
    if (dC && out_centroid) CU(cudaMemcpyAsync(out_centroid + lo * R, dC + lo * R, sizeof(fhp_foothold) * R * n, cudaMemcpyDeviceToHost, cs));
    if (dD) CU(cudaMemcpyAsync(out_default + lo * R, dD + lo * R, sizeof(fhp_foothold) * R * n, cudaMemcpyDeviceToHost, cs));
    CU(cudaMemcpyAsync(status + lo, dS + lo, sizeof(fhp_query_status) * n, cudaMemcpyDeviceToHost, cs));
  }
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream2));
  return check_dev_err(ctx);
}

// The same plan with a COMPACT response: like the service, every query returns only the rows it produced.
// Chunk i plans, scans its n_rows and packs its rows on one of two streams while the packed rows of chunk i-1 cross
// PCIe on the copy stream; the host learns each chunk's row count from a pinned word before it issues that copy.
static int plan_compact_impl(fhp_ctx* ctx, const fhp_params* params, int32_t n_params, const fhp_query* queries,
                             int64_t n_queries, int32_t gait_cycles, void* out_rows, int fmt, int64_t capacity_rows,
                             int64_t* row_offsets, fhp_query_status* status);

int fhp_plan_compact(fhp_ctx* ctx, const fhp_params* params, int32_t n_params, const fhp_query* queries, int64_t n_queries,
                     int32_t gait_cycles, fhp_foothold* out_rows, int64_t capacity_rows, int64_t* row_offsets,
                     fhp_query_status* status) {
  return plan_compact_impl(ctx, params, n_params, queries, n_queries, gait_cycles, out_rows, 40, capacity_rows,
                           row_offsets, status);
}
int fhp_plan_compact24(fhp_ctx* ctx, const fhp_params* params, int32_t n_params, const fhp_query* queries,
                       int64_t n_queries, int32_t gait_cycles, fhp_foothold24* out_rows, int64_t capacity_rows,
                       int64_t* row_offsets, fhp_query_status* status) {
  return plan_compact_impl(ctx, params, n_params, queries, n_queries, gait_cycles, out_rows, 24, capacity_rows,
                           row_offsets, status);
}
int fhp_plan_compact_packed(fhp_ctx* ctx, const fhp_params* params, int32_t n_params, const fhp_query* queries,
                            int64_t n_queries, int32_t gait_cycles, fhp_foothold_packed* out_rows, int64_t capacity_rows,
                            int64_t* row_offsets, fhp_query_status* status) {
  if (ctx && ctx->haveMap && (ctx->M.rows > 65534 || ctx->M.cols > 65534))
    return fail(ctx, FHP_ERR_UNSUPPORTED, "packed rows hold the cell index as uint16: map has more than 65534 rows or columns");
  return plan_compact_impl(ctx, params, n_params, queries, n_queries, gait_cycles, out_rows, 32, capacity_rows,
                           row_offsets, status);
}

static int plan_compact_impl(fhp_ctx* ctx, const fhp_params* params, int32_t n_params, const fhp_query* queries,
                             int64_t n_queries, int32_t gait_cycles, void* out_rows, int fmt, int64_t capacity_rows,
                             int64_t* row_offsets, fhp_query_status* status) {
  if (!ctx) return FHP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (n_queries < 0 || gait_cycles < 0 || gait_cycles > 255 || capacity_rows < 0 || !row_offsets ||
      (n_queries > 0 && (!queries || !out_rows || !status)))
    return fail(ctx, FHP_ERR_INVALID_ARG, "invalid plan arguments (gait_cycles is a uint8 in the service request)");
  if (!ctx->haveMap) return fail(ctx, FHP_ERR_NO_MAP, "fhp_plan_compact before fhp_set_map");
  if (!param_ids_ok(queries, n_queries, n_params)) return fail(ctx, FHP_ERR_INVALID_ARG, "query.param_id out of range");
  CU(cudaSetDevice(ctx->device));
  row_offsets[0] = 0;
  if (n_queries == 0) return FHP_OK;
  const size_t R = FHP_ROWS_PER_QUERY(gait_cycles);
  int rc = prepare_params(ctx, params, n_params, ctx->stream);
  if (rc != FHP_OK) return rc;
  rc = ensure_err_word(ctx);
  if (rc != FHP_OK) return rc;
  // Chunks of a third of a wave: small enough that the first chunk's kernel (the only one PCIe cannot hide) is short,
  // large enough that launches stay negligible and kernels still overlap the copies (measured: 1/3..1/4 of a wave is
  // the flat optimum; a graduated schedule with a small first chunk is slower).
  const int64_t cap = chunk_queries(ctx, false);
  std::vector<int64_t> chunkLo, chunkN;
  {
    const int64_t k = std::max<int64_t>(1, (n_queries + cap - 1) / cap);
    const int64_t per = ((n_queries + k - 1) / k + 31) / 32 * 32;  // whole blocks per chunk
    for (int64_t lo = 0; lo < n_queries; lo += per) {
      chunkLo.push_back(lo);
      chunkN.push_back(std::min<int64_t>(per, n_queries - lo));
    }
  }
  const int64_t nChunks = static_cast<int64_t>(chunkLo.size());
  CU(ctx->dQ.reserve(sizeof(fhp_query) * n_queries));
  CU(ctx->dOutN.reserve(sizeof(fhp_foothold) * R * n_queries));
  CU(ctx->dCompact.reserve(sizeof(fhp_foothold) * R * n_queries));
  CU(ctx->dSt.reserve(sizeof(fhp_query_status) * n_queries));
  CU(ctx->dOffs.reserve(sizeof(long long) * (n_queries + nChunks)));
  if (!ctx->copyStream) CU(cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
  if (ctx->hTotalsCap < nChunks) {
    if (ctx->hTotals) cudaFreeHost(ctx->hTotals);
    ctx->hTotals = nullptr;
    void* h = nullptr;
    CU(cudaHostAlloc(&h, sizeof(long long) * nChunks, cudaHostAllocMapped));
    ctx->hTotals = static_cast<long long*>(h);
    void* d = nullptr;
    CU(cudaHostGetDevicePointer(&d, h, 0));
    ctx->dTotals = static_cast<long long*>(d);
    ctx->hTotalsCap = static_cast<int>(nChunks);
  }
  while (static_cast<int64_t>(ctx->chunkEv.size()) < nChunks) {
    cudaEvent_t e;
    CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->chunkEv.push_back(e);
  }
  fhp_query* dq = static_cast<fhp_query*>(ctx->dQ.p);
  fhp_foothold* dN = static_cast<fhp_foothold*>(ctx->dOutN.p);
  fhp_foothold* dC = static_cast<fhp_foothold*>(ctx->dCompact.p);
  fhp_query_status* dS = static_cast<fhp_query_status*>(ctx->dSt.p);
  long long* dO = static_cast<long long*>(ctx->dOffs.p);
  // enqueue every chunk's device work up front (two alternating streams)
  for (int64_t i = 0; i < nChunks; ++i) ctx->hTotals[i] = 0;
  for (int64_t c = 0; c < nChunks; ++c) {
    const int64_t lo = chunkLo[static_cast<size_t>(c)], n = chunkN[static_cast<size_t>(c)];
    cudaStream_t cs = (c & 1) ? ctx->stream2 : ctx->stream;
    long long* offs = dO + lo + c;  // n + 1 entries per chunk
    CU(cudaMemcpyAsync(dq + lo, queries + lo, sizeof(fhp_query) * n, cudaMemcpyHostToDevice, cs));
    rc = plan_launch(ctx, n_params, dq + lo, n, gait_cycles, dN + lo * R, nullptr, dS + lo, cs);
    if (rc != FHP_OK) return rc;
    CU(launch_compact(dN + lo * R, dS + lo, n, static_cast<int>(R), offs, dC + lo * R, fmt, ctx->dTotals + c, cs));
    ctx->launches += 2;
    CU(cudaEventRecord(ctx->chunkEv[c], cs));
  }
  // As each chunk's row count appears in the mapped word (written by its scan kernel), queue its copies on the copy
  // stream behind the chunk's event: the host never sleeps on an event, so the copy engine always has the next chunk
  // queued.  Rows land behind those of the chunks before; offsets are chunk-local until the end.
  int64_t base = 0;
  int err = FHP_OK;
  std::vector<int64_t> bases(static_cast<size_t>(nChunks), 0);
  const size_t rowB = fmt == 24 ? sizeof(fhp_foothold24) : (fmt == 32 ? sizeof(fhp_foothold_packed) : sizeof(fhp_foothold));
  for (int64_t c = 0; c < nChunks; ++c) {
    const int64_t lo = chunkLo[static_cast<size_t>(c)], n = chunkN[static_cast<size_t>(c)];
    cudaStream_t cs = (c & 1) ? ctx->stream2 : ctx->stream;
    volatile long long* word = ctx->hTotals + c;
    long long t = 0;
    for (unsigned spins = 0; (t = *word) == 0; ++spins) {  // short: the chunk's kernels are already queued
#if defined(__x86_64__) || defined(__i386__)
      __builtin_ia32_pause();
#endif
      if ((spins & 0x3ff) == 0x3ff) {
        const cudaError_t q = cudaStreamQuery(cs);
        if (q == cudaSuccess) {  // the stream drained: the word must be there now
          t = *word;
          if (t == 0) return fail(ctx, FHP_ERR_CUDA, "compact response: the chunk finished without reporting its row count");
          break;
        }
        if (q != cudaErrorNotReady) return fail(ctx, FHP_ERR_CUDA, std::string("compact response: ") + cudaGetErrorString(q));
      }
    }
    const int64_t total = t - 1;
    bases[static_cast<size_t>(c)] = base;
    CU(cudaStreamWaitEvent(ctx->copyStream, ctx->chunkEv[c], 0));
    if (base + total > capacity_rows) {
      err = FHP_ERR_INVALID_ARG;
    } else if (total > 0 && err == FHP_OK) {
      CU(cudaMemcpyAsync(static_cast<char*>(out_rows) + base * rowB, dC + lo * R, rowB * total, cudaMemcpyDeviceToHost,
                         ctx->copyStream));
    }
    CU(cudaMemcpyAsync(row_offsets + lo, dO + lo + c, sizeof(long long) * n, cudaMemcpyDeviceToHost, ctx->copyStream));
    CU(cudaMemcpyAsync(status + lo, dS + lo, sizeof(fhp_query_status) * n, cudaMemcpyDeviceToHost, ctx->copyStream));
    base += total;
  }
  CU(cudaStreamSynchronize(ctx->copyStream));
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream2));
  for (int64_t c = 0; c < nChunks; ++c) {  // chunk-local -> global offsets
    const int64_t lo = chunkLo[static_cast<size_t>(c)], n = chunkN[static_cast<size_t>(c)], b = bases[static_cast<size_t>(c)];
    if (b != 0)
      for (int64_t i = 0; i < n; ++i) row_offsets[lo + i] += b;
  }
  row_offsets[n_queries] = base;
  if (err != FHP_OK) return fail(ctx, err, "capacity_rows is smaller than the rows the queries produced (row_offsets[n] holds the need)");
  return check_dev_err(ctx);
}

// Streamed response: ONE launch plans the whole batch and writes every produced row straight into the caller's pinned,
// device-mapped host buffer while it runs — PCIe carries the rows as they are produced, there is no staging copy, no
// compaction pass and no chunking; the call ends with one stream synchronisation.
int fhp_plan_stream(fhp_ctx* ctx, const fhp_params* params, int32_t n_params, const fhp_query* queries, int64_t n_queries,
                    int32_t gait_cycles, fhp_point20* out_blocks, fhp_stream_mask* masks, fhp_query_status* status) {
  if (!ctx) return FHP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (n_queries < 0 || gait_cycles < 0 || gait_cycles > 255 ||
      (n_queries > 0 && (!queries || !out_blocks || !status || (gait_cycles > 0 && !masks))))
    return fail(ctx, FHP_ERR_INVALID_ARG, "invalid plan arguments (gait_cycles is a uint8 in the service request)");
  if (!ctx->haveMap) return fail(ctx, FHP_ERR_NO_MAP, "fhp_plan_stream before fhp_set_map");
  if (!param_ids_ok(queries, n_queries, n_params)) return fail(ctx, FHP_ERR_INVALID_ARG, "query.param_id out of range");
  CU(cudaSetDevice(ctx->device));
  if (n_queries == 0) return FHP_OK;
  if ((reinterpret_cast<uintptr_t>(out_blocks) & 127u) != 0 || (reinterpret_cast<uintptr_t>(masks) & 7u) != 0)
    return fail(ctx, FHP_ERR_INVALID_ARG, "fhp_plan_stream: out_blocks must be 128-byte aligned, masks 8-byte aligned");
  void* dSlots = mapped_device_ptr(out_blocks);
  void* dStat = mapped_device_ptr(status);
  void* dMask = masks ? mapped_device_ptr(masks) : nullptr;
  if (!dSlots || !dStat || (masks && !dMask))
    return fail(ctx, FHP_ERR_INVALID_ARG,
                "fhp_plan_stream: out_blocks, masks and status must be page-locked host memory (cudaHostAlloc / "
                "cudaHostRegister): the kernel writes them directly");
  cudaStream_t s = ctx->stream;
  int rc = prepare_params(ctx, params, n_params, s);
  if (rc != FHP_OK) return rc;
  rc = ensure_err_word(ctx);
  if (rc != FHP_OK) return rc;
  // queries: read in place over PCIe when they sit in mapped pinned memory too (32 bytes per query, coalesced per warp;
  // the link is full duplex, so these reads do not compete with the row writes); copied first otherwise
  const fhp_query* dQueries = nullptr;
  {
    void* dq = mapped_device_ptr(queries);
    if (dq) {
      dQueries = static_cast<const fhp_query*>(dq);
    } else {
      CU(ctx->dQ.reserve(sizeof(fhp_query) * n_queries));
      CU(cudaMemcpyAsync(ctx->dQ.p, queries, sizeof(fhp_query) * n_queries, cudaMemcpyHostToDevice, s));
      dQueries = static_cast<const fhp_query*>(ctx->dQ.p);
    }
  }
  StreamOut so;
  so.slots = static_cast<fhp_point20*>(dSlots);
  so.mask = static_cast<uint32_t*>(dMask);
  so.nq = n_queries;
  so.q0 = 0;
  rc = plan_launch(ctx, n_params, dQueries, n_queries, gait_cycles, nullptr, nullptr,
                   static_cast<fhp_query_status*>(dStat), s, nullptr, &so);
  if (rc != FHP_OK) return rc;
  CU(cudaStreamSynchronize(s));
  return check_dev_err(ctx);
}

const fhp_point20* fhp_stream_row(const fhp_point20* blocks, const fhp_stream_mask* masks, int64_t n_queries,
                                  int32_t gait_cycles, int64_t q, int32_t g, int32_t f) {
  if (!blocks || q < 0 || q >= n_queries || g < -1 || g >= gait_cycles || f < 0 || f > 3) return nullptr;
  const int64_t groups = (n_queries + 7) / 8, G = q >> 3;
  int rank = static_cast<int>(q & 7);
  if (g >= 0) {
    if (!masks) return nullptr;
    const int words = FHP_STREAM_MASK_WORDS(gait_cycles);
    const uint32_t bit = 1u << (g & 31);
    if (!(masks[(q * words + (g >> 5)) * 4].success & bit)) return nullptr;
    rank = 0;
    for (int64_t p = G * 8; p < q; ++p) rank += (masks[(p * words + (g >> 5)) * 4].success & bit) ? 1 : 0;
  }
  return blocks + ((static_cast<int64_t>(g + 1) * groups + G) * 32 + rank * 4 + f);
}

// one 24-byte service row from a 20-byte streamed point: the first 20 bytes are the same {x, y, z}, the last four are
// {foot_id, gait_cycle_id, kind, 0} as one word
static inline void put_row(fhp_foothold24* o, const fhp_point20* b, int foot, int cycle, int kind) {
  static_assert(sizeof(fhp_point20) == 20 && sizeof(fhp_foothold24) == 24 && offsetof(fhp_foothold24, z) == 16 &&
                    offsetof(fhp_foothold24, foot_id) == 20 && offsetof(fhp_foothold24, kind) == 22,
                "row layouts");
  std::memcpy(o, b, 20);
  const uint32_t tail = static_cast<uint32_t>(foot) | (static_cast<uint32_t>(cycle) << 8) | (static_cast<uint32_t>(kind) << 16);
  std::memcpy(reinterpret_cast<char*>(o) + 20, &tail, 4);
}
int fhp_stream_expand(const fhp_point20* blocks, const fhp_stream_mask* masks, const fhp_query_status* status,
                      int64_t n_queries, int32_t gait_cycles, fhp_foothold24* out_rows, int64_t capacity_rows,
                      int64_t* row_offsets, int32_t n_threads) {
  if (n_queries < 0 || gait_cycles < 0 || gait_cycles > 255 || !row_offsets ||
      (n_queries > 0 && (!blocks || !status || !out_rows || (gait_cycles > 0 && !masks))))
    return FHP_ERR_INVALID_ARG;
  row_offsets[0] = 0;
  for (int64_t q = 0; q < n_queries; ++q) row_offsets[q + 1] = row_offsets[q] + status[q].n_rows;  // 4 + 4 per success
  if (row_offsets[n_queries] > capacity_rows) return FHP_ERR_INVALID_ARG;
  const int64_t groups = (n_queries + 7) / 8;
  const int words = FHP_STREAM_MASK_WORDS(gait_cycles);
  int nt = n_threads > 0 ? n_threads : static_cast<int>(std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency())));
  nt = static_cast<int>(std::min<int64_t>(nt, std::max<int64_t>(1, groups / 64)));
  auto work = [&](int64_t g0, int64_t g1) {
    for (int64_t G = g0; G < g1; ++G) {
      const int64_t q0 = G * 8, q1 = std::min<int64_t>(n_queries, q0 + 8);
      for (int64_t q = q0; q < q1; ++q) {  // stance rows: block (0, G), query order
        fhp_foothold24* o = out_rows + row_offsets[q];
        const fhp_point20* b = blocks + (G * 32 + (q - q0) * 4);
        for (int f = 0; f < 4; ++f) put_row(o + f, b + f, f, 0, FHP_KIND_STANCE);
      }
      int64_t cursor[8];
      for (int64_t q = q0; q < q1; ++q) cursor[q - q0] = row_offsets[q] + 4;
      for (int g = 0; g < gait_cycles; ++g) {  // cycle blocks: the group's successful queries packed to the front
        const uint32_t bit = 1u << (g & 31);
        const fhp_point20* b = blocks + ((static_cast<int64_t>(g) + 1) * groups + G) * 32;
        for (int64_t q = q0; q < q1; ++q) {
          const fhp_stream_mask* m = masks + (q * words + (g >> 5)) * 4;
          if (!(m[0].success & bit)) continue;
          fhp_foothold24* o = out_rows + cursor[q - q0];
          for (int f = 0; f < 4; ++f)
            put_row(o + f, b + f, f, g, (m[f].candidate & bit) ? FHP_KIND_CANDIDATE : FHP_KIND_DEFAULT);
          cursor[q - q0] += 4;
          b += 4;
        }
      }
    }
  };
  if (nt <= 1) {
    work(0, groups);
  } else {
    std::vector<std::thread> th;
    const int64_t per = (groups + nt - 1) / nt;
    for (int t = 0; t < nt; ++t) th.emplace_back(work, std::min<int64_t>(groups, t * per), std::min<int64_t>(groups, (t + 1) * per));
    for (std::thread& t : th) t.join();
  }
  return FHP_OK;
}

// KPIs of the reference's logs (getHipDistance FP.cpp:2571-2584, getCogSpeed :2587-2623, gaitCycle_ = 1.0 :332) —
// pure functions of one query's response rows, computed on the host like the message assembly.
int fhp_cycle_kpi(const fhp_params* params, const fhp_query* query, const fhp_foothold* rows, int32_t n_rows,
                  double* feet_distance, double* cog_speed) {
  if (!params || !query || !rows || !feet_distance || !cog_speed || n_rows < 4 || (n_rows % 4) != 0)
    return FHP_ERR_INVALID_ARG;
  ParDev d;
  derive_params(*params, d);
  double cur[4];  // x of the chain's current feet before the first success: stance - stepLength/2 (setFirstGait)
  for (int f = 0; f < 4; ++f) cur[f] = (d.stancex[f] + query->x) - d.halfStep;
  const int cycles = n_rows / 4 - 1;
  for (int c = 0; c < cycles; ++c) {
    const fhp_foothold* r = rows + 4 + 4 * c;  // RF, RH, LH, LF
    feet_distance[2 * c] = r[0].x - r[2].x;
    feet_distance[2 * c + 1] = r[3].x - r[1].x;
    double fc1, fc2, fc3;
    if (params->rf_first) {
      fc1 = (cur[0] + cur[2]) / 2;
      fc2 = (r[3].x + r[1].x) / 2;
      fc3 = (r[0].x + r[2].x) / 2;
    } else {
      fc1 = (cur[3] + cur[1]) / 2;
      fc2 = (r[0].x + r[2].x) / 2;
      fc3 = (r[3].x + r[1].x) / 2;
    }
    cog_speed[2 * c] = (fc2 - fc1) / (0.5 * 1.0);
    cog_speed[2 * c + 1] = (fc3 - fc2) / (0.5 * 1.0);
    for (int f = 0; f < 4; ++f) cur[f] = r[f].x;
  }
  return FHP_OK;
}

int fhp_plan_device(fhp_ctx* ctx, const fhp_params* params, int32_t n_params, const fhp_query* d_queries,
                    int64_t n_queries, int32_t gait_cycles, fhp_foothold* d_out_nominal, fhp_foothold* d_out_centroid,
                    fhp_query_status* d_status, void* cuda_stream) {
  if (!ctx) return FHP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (n_queries < 0 || gait_cycles < 0 || gait_cycles > 255 ||
      (n_queries > 0 && (!d_queries || !d_out_nominal || !d_status)))
    return fail(ctx, FHP_ERR_INVALID_ARG, "invalid plan arguments");
  if (!ctx->haveMap) return fail(ctx, FHP_ERR_NO_MAP, "fhp_plan_device before fhp_set_map");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
  ctx->pendingErrCheck = 1;
  if (cuda_stream) ctx->asyncPlans = true;  // shared device state is re-synchronised before it is next rewritten
  return plan_common(ctx, params, n_params, d_queries, n_queries, gait_cycles, d_out_nominal, d_out_centroid, d_status,
                     s);
}

int fhp_synchronize(fhp_ctx* ctx) {
  if (!ctx) return FHP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream2));
  if (ctx->pendingErrCheck && ctx->hErr) {
    ctx->pendingErrCheck = 0;
    CU(cudaDeviceSynchronize());
    return check_dev_err(ctx);
  }
  return FHP_OK;
}

static int run_task(fhp_ctx* ctx, int mode, const fhp_params* params, const double* centres, const double* rects,
                    int64_t n, fhp_foothold* out, uint8_t* valid, int8_t* ccase, float* z) {
  if (!ctx) return FHP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (n < 0 || !params || (n > 0 && !centres)) return fail(ctx, FHP_ERR_INVALID_ARG, "invalid task arguments");
  if (!ctx->haveMap) return fail(ctx, FHP_ERR_NO_MAP, "no map set");
  CU(cudaSetDevice(ctx->device));
  if (n == 0) return FHP_OK;
  cudaStream_t s = ctx->stream;
  int rc = prepare_params(ctx, params, 1, s);
  if (rc != FHP_OK) return rc;
  rc = ensure_err_word(ctx);
  if (rc != FHP_OK) return rc;
  const size_t cB = sizeof(double) * 2 * n;
  CU(ctx->dA.reserve(cB));
  CU(cudaMemcpyAsync(ctx->dA.p, centres, cB, cudaMemcpyHostToDevice, s));
  if (rects) {
    CU(ctx->dB.reserve(cB));
    CU(cudaMemcpyAsync(ctx->dB.p, rects, cB, cudaMemcpyHostToDevice, s));
  }
  if (out) CU(ctx->dOutN.reserve(sizeof(fhp_foothold) * n));
  CU(ctx->dC2.reserve(static_cast<size_t>(n)));      // valid / case bytes
  CU(ctx->dD.reserve(sizeof(float) * n));             // z
  TaskArgs A;
  A.M = ctx->M;
  A.pars = static_cast<const ParDev*>(ctx->dPars.p);
  A.centres = static_cast<const double*>(ctx->dA.p);
  A.rects = rects ? static_cast<const double*>(ctx->dB.p) : nullptr;
  A.n = n;
  A.out = static_cast<fhp_foothold*>(ctx->dOutN.p);
  A.valid = static_cast<uint8_t*>(ctx->dC2.p);
  A.ccase = static_cast<int8_t*>(ctx->dC2.p);
  A.z = static_cast<float*>(ctx->dD.p);
  A.err = ctx->dErrMapped;
  GeoDev Gt = ctx->G;  // the member kernels stage from the float layer into the {D, C, F, pad} layout (4 words per column
  Gt.stride = 4 * Gt.nIter;  // group), whatever plane layout the plan kernels of this context use
  CU(launch_task(mode, A, Gt, s));
  ctx->launches += 1;
  if (out) CU(cudaMemcpyAsync(out, ctx->dOutN.p, sizeof(fhp_foothold) * n, cudaMemcpyDeviceToHost, s));
  if (valid) CU(cudaMemcpyAsync(valid, ctx->dC2.p, static_cast<size_t>(n), cudaMemcpyDeviceToHost, s));
  if (ccase) CU(cudaMemcpyAsync(ccase, ctx->dC2.p, static_cast<size_t>(n), cudaMemcpyDeviceToHost, s));
  if (z) CU(cudaMemcpyAsync(z, ctx->dD.p, sizeof(float) * n, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return check_dev_err(ctx);
}

int fhp_check_foothold(fhp_ctx* ctx, const fhp_params* params, const double* centres_xy, const double* rect_centres_xy,
                       int64_t n, fhp_foothold* out, uint8_t* valid_out) {
  if (n > 0 && (!rect_centres_xy || !out || !valid_out)) return fail(ctx, FHP_ERR_INVALID_ARG, "null output");
  return run_task(ctx, 0, params, centres_xy, rect_centres_xy, n, out, valid_out, nullptr, nullptr);
}
int fhp_check_centroid(fhp_ctx* ctx, const fhp_params* params, const double* points_xy, int64_t n, fhp_foothold* out,
                       int8_t* case_out) {
  if (n > 0 && (!out || !case_out)) return fail(ctx, FHP_ERR_INVALID_ARG, "null output");
  return run_task(ctx, 1, params, points_xy, nullptr, n, out, nullptr, case_out, nullptr);
}
int fhp_mean_height(fhp_ctx* ctx, const fhp_params* params, const double* centres_xy, int64_t n, float* z_out) {
  if (n > 0 && !z_out) return fail(ctx, FHP_ERR_INVALID_ARG, "null output");
  return run_task(ctx, 2, params, centres_xy, nullptr, n, nullptr, nullptr, nullptr, z_out);
}
int fhp_check_default(fhp_ctx* ctx, const fhp_params* params, const double* centres_xy, int64_t n, uint8_t* valid_out) {
  if (n > 0 && !valid_out) return fail(ctx, FHP_ERR_INVALID_ARG, "null output");
  return run_task(ctx, 3, params, centres_xy, nullptr, n, nullptr, valid_out, nullptr, nullptr);
}

// Host-side geometry of the path (pure functions of the parameters, like the message assembly): the same expressions
// the kernel evaluates, for callers of the reference's public geometry members.
int fhp_initial_stance(const fhp_params* params, const fhp_query* query, double* initial_xyz, double* first_gait_xyz) {
  if (!params || !query || !initial_xyz || !first_gait_xyz) return FHP_ERR_INVALID_ARG;
  ParDev d;
  derive_params(*params, d);
  for (int f = 0; f < 4; ++f) {  // initialize() (FP.cpp:350-378), setFirstGait (FP.cpp:2679-2699)
    initial_xyz[3 * f] = d.stancex[f] + query->x;
    initial_xyz[3 * f + 1] = d.stancey[f] + query->y;
    initial_xyz[3 * f + 2] = 0.0 + query->z;
    first_gait_xyz[3 * f] = initial_xyz[3 * f] - (d.heading ? d.halfStepX : d.halfStep);
    first_gait_xyz[3 * f + 1] = d.heading ? initial_xyz[3 * f + 1] - d.halfStepY : initial_xyz[3 * f + 1];
    first_gait_xyz[3 * f + 2] = initial_xyz[3 * f + 2];
  }
  return FHP_OK;
}

int fhp_search_geometry(const fhp_params* params, const fhp_query* query, double adjusted_y, const double* cur_xyz,
                        double* center_xyz, double* next_xy, double* polygon_xy) {
  if (!params || !query || !cur_xyz || !center_xyz || !next_xy || !polygon_xy) return FHP_ERR_INVALID_ARG;
  ParDev d;
  derive_params(*params, d);
  // getPolygonCenter (FP.cpp:2421-2463): area-weighted centroid of the quad RF->RH->LH->LF by triangle fan
  const double x1 = cur_xyz[0], y1 = cur_xyz[1];
  double x2 = cur_xyz[3], y2 = cur_xyz[4];
  double sum_x = 0, sum_y = 0, sum_s = 0;
  for (int i = 2; i <= 3; ++i) {
    const double x3 = cur_xyz[3 * i], y3 = cur_xyz[3 * i + 1];
    const double s = ((x2 - x1) * (y3 - y1) - (x3 - x1) * (y2 - y1)) / 2.0;
    sum_x += (x1 + x2 + x3) * s;
    sum_y += (y1 + y2 + y3) * s;
    sum_s += s;
    x2 = x3;
    y2 = y3;
  }
  center_xyz[0] = sum_x / sum_s / 3.0;
  center_xyz[1] = sum_y / sum_s / 3.0;
  center_xyz[2] = (cur_xyz[2] + cur_xyz[5] + cur_xyz[8] + cur_xyz[11]) / 4.0;
  // getFootholdSearchGridMap (FP.cpp:2166-2250): next centre, getDefaultFootholdNext (:2411-2418), getSearchPolygon
  double ncx = center_xyz[0] + d.stepLength;  // :2199
  double ncy = query->y + adjusted_y;         // :2201
  if (d.heading) {  // heading extension (fhp_params.heading): forward coordinate from the centroid, lateral one pinned
    auto mulz = [](double a, double t) { return t == 0.0 ? 0.0 : a * t; };  // an exactly zero sin drops its term
    const double fwd = (center_xyz[0] * d.ch + mulz(center_xyz[1], d.sh)) + d.stepLength;
    const double lat = (query->y * d.ch - mulz(query->x, d.sh)) + adjusted_y;
    ncx = fwd * d.ch - mulz(lat, d.sh);
    ncy = mulz(fwd, d.sh) + lat * d.ch;
  }
  for (int f = 0; f < 4; ++f) {
    const double nx = ncx + d.biasx[f], ny = ncy + d.biasy[f];
    next_xy[2 * f] = nx;
    next_xy[2 * f + 1] = ny;
    double* v = polygon_xy + 8 * f;  // :2496-2517, vertex order leftup, rightup, rightdown, leftdown
    if (d.heading) {
      for (int k = 0; k < 4; ++k) {
        v[2 * k] = nx + d.rectVx[k];
        v[2 * k + 1] = ny + d.rectVy[k];
      }
      continue;
    }
    v[0] = nx + d.R;  v[1] = ny + 0.5 * d.R;
    v[2] = nx + d.R;  v[3] = ny - 0.5 * d.R;
    v[4] = nx - d.R;  v[5] = ny - 0.5 * d.R;
    v[6] = nx - d.R;  v[7] = ny + 0.5 * d.R;
  }
  return FHP_OK;
}

int fhp_check_candidate(fhp_ctx* ctx, const fhp_params* params, const double* centres_xy, const double* rect_centres_xy,
                        int64_t n, fhp_foothold* out, uint8_t* valid_out) {
  if (n > 0 && (!rect_centres_xy || !out || !valid_out)) return fail(ctx, FHP_ERR_INVALID_ARG, "null output");
  return run_task(ctx, 4, params, centres_xy, rect_centres_xy, n, out, valid_out, nullptr, nullptr);
}
int fhp_check_circle_polygon(fhp_ctx* ctx, const fhp_params* params, const double* foot_centres_xy,
                             const double* rect_centres_xy, int64_t n, uint8_t* valid_out) {
  if (n > 0 && (!rect_centres_xy || !valid_out)) return fail(ctx, FHP_ERR_INVALID_ARG, "null output");
  return run_task(ctx, 5, params, foot_centres_xy, rect_centres_xy, n, nullptr, valid_out, nullptr, nullptr);
}

int64_t fhp_kernel_launches(const fhp_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---- NCCL map replication (symbols resolved at run time from the NCCL already in the process) ------------------
typedef int (*nccl_bcast_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
static void* nccl_sym(const char* name) {
  void* f = dlsym(RTLD_DEFAULT, name);
  if (f) return f;
  static void* h = nullptr;
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW);
  return h ? dlsym(h, name) : nullptr;
}

int fhp_broadcast_map(fhp_ctx* ctx, const fhp_map_desc* desc, void* nccl_comm, int root, int rank) {
  if (!ctx) return FHP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!desc || !nccl_comm || !desc->layer_names) return fail(ctx, FHP_ERR_INVALID_ARG, "null argument");
  CU(cudaSetDevice(ctx->device));
  nccl_bcast_fn bcast = reinterpret_cast<nccl_bcast_fn>(nccl_sym("ncclBroadcast"));
  if (!bcast) return fail(ctx, FHP_ERR_NCCL, "ncclBroadcast not found in the process");
  const size_t count = static_cast<size_t>(desc->rows) * desc->cols;
  if (rank != root) {
    int rc = set_geometry(ctx, desc);
    if (rc != FHP_OK) return rc;
    CU(cudaDeviceSynchronize());
    free_owned(ctx);
    ctx->layerNames.clear();
    ctx->layerPtrs.clear();
    ctx->haveMap = false;
    for (int i = 0; i < desc->n_layers; ++i) {
      float* d = nullptr;
      CU(cudaMalloc(&d, count * sizeof(float)));
      ctx->owned.push_back(d);
      ctx->layerNames.push_back(desc->layer_names[i]);
      ctx->layerPtrs.push_back(d);
    }
  } else {
    if (!ctx->haveMap) return fail(ctx, FHP_ERR_NO_MAP, "root has no map to broadcast");
    if (static_cast<int>(ctx->layerNames.size()) != desc->n_layers || ctx->M.rows != desc->rows ||
        ctx->M.cols != desc->cols)
      return fail(ctx, FHP_ERR_INVALID_ARG, "desc does not match the root's map");
  }
  for (int i = 0; i < desc->n_layers; ++i) {
    // ncclFloat32 == 7
    int r = bcast(ctx->layerPtrs[i], const_cast<float*>(ctx->layerPtrs[i]), count, 7, root, nccl_comm, ctx->stream);
    if (r != 0) return fail(ctx, FHP_ERR_NCCL, "ncclBroadcast failed with code " + std::to_string(r));
  }
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->maskValid = false;  // (the root's layers may have been rewritten in place before the broadcast)
  return bind_layers(ctx);
}

// The caller rewrote an adopted / broadcast layer IN PLACE (fhp_map_layer_device, fhp_set_map_device): every derived
// structure — the cell-class planes of the default staging path — is rebuilt by the next plan.
int fhp_map_changed(fhp_ctx* ctx) {
  if (!ctx) return FHP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!ctx->haveMap) return fail(ctx, FHP_ERR_NO_MAP, "no map set");
  CU(cudaSetDevice(ctx->device));
  CU(cudaDeviceSynchronize());  // no plan may be reading the planes that are about to be rebuilt
  ctx->maskValid = false;
  return FHP_OK;
}

}  // extern "C"
