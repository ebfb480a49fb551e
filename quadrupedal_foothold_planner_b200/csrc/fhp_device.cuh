// libfhp device code (sm_100a).  One LANE owns one (query, leg) — or one fine-grained task — and carries its state in
// registers; the 32 lanes of a warp cooperate only where the work is wide: staging each lane's search window into
// shared memory and scanning a spiral's ranks 32 at a time.
//
// Why this shape: the per-foothold work is ~100 map cells and a few hundred double-precision
// scalar operations that must reproduce grid_map_core's arithmetic bit for bit.  Lane-per-leg makes that scalar
// restatement cost 1/32 of a warp instruction per foothold.  Every map-cell predicate of the path is a per-call
// constant, so the (2*HW+1)^2 traversability window around the search centre lives in shared memory as three
// bit-planes per window column:
//     D: cell < defaultFootholdThreshold      (raw compare; NaN -> 0)          FP.cpp:1653, :1736, :2057
//     C: cell < candidateFootholdThreshold                                      FP.cpp:2138
//     F: isfinite(cell)                        GridMap::isValid                 FP.cpp:2055, :2132
// filled either from the per-map cell-class planes (stage_windows_mask: one 16-byte load per column, the default) or
// from the float layer (stage_windows_tma: 2-D TMA boxes; stage_windows: coalesced loads; both reduced with
// __ballot_sync).  Every decision of the reference (centroid row scan, foot-disc test, spiral candidates, polygon
// test) then becomes bit arithmetic on those planes in the owning lane.
//
// All double arithmetic follows the operand order of grid_map_core / FootholdPlanner.cpp; this translation unit
// is compiled with -fmad=false so no multiply-add is ever contracted (x86-64 reference builds have no FMA); the few
// fused operations (div3_rn) are explicit and proven to round like the expression they replace.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fhp_types.h"

namespace fhp {

// Geometry as the device code sees it.  SPEC 0: every value is read from the GeoDev tables (any geometry).  SPEC > 0:
// the values of ONE concrete geometry are compile-time constants — the staging, plane and scan loops unroll, the generic
// branches (wide discs, multi-cell foot stencils, packed plane layouts) vanish and the register allocator has less to
// keep alive: 15 % fewer cycles on the yaml geometry at 0.02 m (profiles/r2_*).  The host picks a specialised kernel only
// when the tables it built match the constants exactly (geo_spec_of, fhp_capi.cu); specialised variants exist for the
// cell-class (mask) staging path only, so pack = 1 and the plane pitch follows from WIN.
template <int SPEC>
struct GeoK {
  static constexpr bool fixed = false;
  static constexpr int WIN = 0, HW = 0, nSten = 0, discK = 0, nCand = 0, lastTwoStart = 0, nRings = 0, ra = 0;
};
template <>
struct GeoK<1> {  // footRadius 0.02, searchRadius 0.1 at 0.02 m (yaml:17-44), rings trimmed to 5 (build_geo)
  static constexpr bool fixed = true;
  static constexpr int WIN = 11, HW = 5, nSten = 1, discK = 1, nCand = 109, lastTwoStart = 69, nRings = 6, ra = 0;
};
template <>
struct GeoK<2> {  // the same radii at 0.01 m: 3x3 foot stencil, rings trimmed to 10
  static constexpr bool fixed = true;
  static constexpr int WIN = 23, HW = 11, nSten = 9, discK = 2, nCand = 373, lastTwoStart = 305, nRings = 11, ra = 1;
};
template <int SPEC>
struct GeoV {
  using K = GeoK<SPEC>;
  const GeoDev& g;
  __device__ __forceinline__ explicit GeoV(const GeoDev& g_) : g(g_) {}
#define FHP_GEO_K(name) \
  __device__ __forceinline__ int name() const { if constexpr (K::fixed) return K::name; else return g.name; }
  FHP_GEO_K(WIN) FHP_GEO_K(HW) FHP_GEO_K(nSten) FHP_GEO_K(discK) FHP_GEO_K(nCand) FHP_GEO_K(lastTwoStart) FHP_GEO_K(nRings) FHP_GEO_K(ra)
#undef FHP_GEO_K
  __device__ __forceinline__ int nIter() const { if constexpr (K::fixed) return K::WIN; else return g.nIter; }
  __device__ __forceinline__ int stride() const { if constexpr (K::fixed) return 3 * K::WIN + 1; else return g.stride; }
  __device__ __forceinline__ int winRows() const { if constexpr (K::fixed) return K::WIN; else return g.winRows; }
  __device__ __forceinline__ int pack() const { if constexpr (K::fixed) return 1; else return g.pack; }
  __device__ __forceinline__ int rpShift() const { if constexpr (K::fixed) return 5; else return g.rpShift; }
#define FHP_GEO_R(name) __device__ __forceinline__ int name() const { return g.name; }
  FHP_GEO_R(spiralWords) FHP_GEO_R(smemWords) FHP_GEO_R(stageMode) FHP_GEO_R(useTma) FHP_GEO_R(tmaB) FHP_GEO_R(tmaBShift)
  FHP_GEO_R(boxWords) FHP_GEO_R(mbarOff) FHP_GEO_R(boxOff) FHP_GEO_R(elevOff)
#undef FHP_GEO_R
};

// Inlining policy of the small geometry helpers (A/B knobs: the hot loop's instruction footprint vs call overhead)
#ifdef FHP_NI_RAW
#define FHP_INL_RAW __noinline__
#else
#define FHP_INL_RAW __forceinline__
#endif
#ifdef FHP_NI_BOUND
#define FHP_INL_BOUND __noinline__
#else
#define FHP_INL_BOUND __forceinline__
#endif
#ifdef FHP_NI_FB
#define FHP_INL_FB __noinline__
#else
#define FHP_INL_FB __forceinline__
#endif
#ifdef FHP_NI_PC
#define FHP_INL_PC __noinline__
#else
#define FHP_INL_PC __forceinline__
#endif

// ------------------------------------------------------------------------------------------------ grid_map math
// boundPositionToRange (GridMapMath.cpp) for one axis
__device__ FHP_INL_BOUND double bound_axis(double p, double len, double mp, double half) {
  double shifted = (p - mp) + half;
  double eps = 10.0 * 2.220446049250313e-16;
  if (fabs(p) > 1.0) eps *= fabs(p);
  if (shifted <= 0) {
    shifted = eps;
  } else if (shifted >= len) {
    shifted = len - eps;
  }
  return (shifted + mp) - half;
}
// IEEE double division as ONE out-of-line subroutine: the inline expansion is ~35 instructions per site, and the
// instruction footprint of the scalar phases is what limits this kernel (profiles/).
__device__ __noinline__ double div_rn(double a, double b) { return a / b; }
// a / 3.0, correctly rounded, without the division subroutine: y = RN(1/3); a quotient estimate refined once is
// faithful, and Markstein's final step q + RN(a - 3q)*y (residual exact in one FMA) then rounds correctly.  The FMAs
// are explicit (the translation unit is built with -fmad=false); operands outside the guard band (where a residual
// could leave the normal range) take the exact subroutine.  scratch-tested against a/3.0 on 4e8 operands.
__device__ __forceinline__ double div3_rn(double a) {
  if (!(fabs(a) < 1e280)) {
    // infinities and NaNs divide to themselves (a degenerate stance polygon — two feet of the centroid chain at (0,0,0) —
    // has a zero area and a non-finite centroid for the rest of its plan: ~2 % of the lanes on C3); huge finite values
    // take the exact routine
    if (!(fabs(a) <= 1.7976931348623157e308)) return a * 0.33333333333333331482961625624739;
    return div_rn(a, 3.0);
  }
  if (!(fabs(a) > 1e-280)) return div_rn(a, 3.0);
  const double y = 0.33333333333333331482961625624739;  // 0x3FD5555555555555
  const double q0 = a * y;
  const double q1 = __fma_rn(__fma_rn(-3.0, q0, a), y, q0);
  return __fma_rn(__fma_rn(-3.0, q1, a), y, q1);
}

// getIndexFromPosition without the start-index wrap: (int)(-(((p - 0.5*len) - mapPos)/res)).
// Fast path: multiply by the correctly rounded reciprocal; the product differs from the IEEE quotient by at most a
// few ulp, so the truncated integer can only differ when the value sits within that distance of an integer —
// then (and only then) the exact division decides.  `rinv` = 1/res (MapDev::rinv).
__device__ FHP_INL_RAW int raw_index_axis(double p, double half, double mp, double res, double rinv) {
  const double x = (p - half) - mp;
  double t = -(x * rinv);
  // (int) of a NaN / out-of-int-range double is UB upstream (x86 cvttsd2si yields INT_MIN).  Policy shared with the
  // CPU restatement: every such index becomes -2^30 — far outside any map, and index +- window never overflows.
  if (!(fabs(t) < 1073741000.0)) {
    t = -div_rn(x, res);
    if (!(fabs(t) < 1073741824.0)) return -1073741824;
    return static_cast<int>(t);
  }
  const double fr = t - rint(t);
  if (fabs(fr) < 1e-6) t = -div_rn(x, res);  // |t| < 2^30: quotient error < 2^30 * 4 ulp ~ 1e-6 is a safe band
  return static_cast<int>(t);
}
// checkIfPositionWithinMap for one axis
__device__ __forceinline__ bool within_axis(double p, double len, double mp, double half) {
  double t = -((p - mp) - half);
  return t >= 0.0 && t < len;
}
// getPositionFromIndex (unwrapped index): (mapPos + offset) + res * (double)(-u)
__device__ __forceinline__ double pos_axis(double fc, double res, int u) { return fc + res * static_cast<double>(-u); }

__device__ __forceinline__ int wrap_add(int u, int start, int n) {  // unwrapped (in range) -> buffer index
  int b = u + start;
  return b >= n ? b - n : b;
}
__device__ __forceinline__ bool in_map(const MapDev& M, int u0, int u1) {
  return u0 >= 0 && u1 >= 0 && u0 < M.rows && u1 < M.cols;
}
__device__ __forceinline__ float load_cell(const MapDev& M, const float* layer, int u0, int u1) {
  int b0 = wrap_add(u0, M.start0, M.rows), b1 = wrap_add(u1, M.start1, M.cols);
  return __ldg(layer + static_cast<size_t>(b1) * M.rows + b0);
}
// getIndex(position) as the reference's diagnostics use it: buffer index, or (-1,-1) outside the map
__device__ __forceinline__ void cell_of_position(const MapDev& M, double x, double y, int& row, int& col) {
  bool ok = within_axis(x, M.lenx, M.posx, M.halfx) && within_axis(y, M.leny, M.posy, M.halfy);
  if (!ok) {
    row = col = -1;
    return;
  }
  int u0 = raw_index_axis(x, M.halfx, M.posx, M.res, M.rinv), u1 = raw_index_axis(y, M.halfy, M.posy, M.res, M.rinv);
  if (!in_map(M, u0, u1)) {  // unreachable for finite geometry; keep the contract total
    row = col = -1;
    return;
  }
  row = wrap_add(u0, M.start0, M.rows);
  col = wrap_add(u1, M.start1, M.cols);
}

// ------------------------------------------------------------------------------------------------ shared memory
// Explicit shared-space accesses on 32-bit addresses (the compiler otherwise falls back to generic 64-bit
// addressing once a shared pointer has travelled through a struct).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint2 lds_v2(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_v2(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ------------------------------------------------------------------------------------------------ bit planes
struct Planes {  // view of one query's staged window in shared memory
  const uint32_t* w;
  int pack, rpShift, win;
  // from a shared-space byte address: plain loads (the compiler may keep words in registers between barriers).
  // pack 1 / 2 / 4: the float paths' layout, 4 words {D, C, F, pad} per group of `pack` window columns.
  // pack 0: the cell-class (mask) path's layout — per window column the pair (e0, e1) of class bit planes, 8 bytes, then
  // one word per column for the owning lane's finished "bad cell" masks: D = e1, C = e0 & e1, F = e0 | e1 cost one
  // operation at the point of use and the slot shrinks from 16 to 12 bytes per column (occupancy at 0.01 m).
  __device__ __forceinline__ Planes(uint32_t addr, int pack_, int rpShift_, int win_)
      : w(static_cast<const uint32_t*>(__cvta_shared_to_generic(addr))), pack(pack_), rpShift(rpShift_), win(win_) {}
  __device__ __forceinline__ uint32_t col(int plane, int b) const {  // bits over window rows a for column b
    if (pack == 0) {
      if (plane == 3) return w[2 * win + b];
      const uint2 e = *reinterpret_cast<const uint2*>(w + 2 * b);
      return plane == 0 ? e.y : (plane == 1 ? (e.x & e.y) : (e.x | e.y));
    }
    if (pack == 1) return w[b * 4 + plane];
    int k, sh;
    if (pack == 2) {
      k = b >> 1;
      sh = (b & 1) << 4;
      return (w[k * 4 + plane] >> sh) & 0xffffu;
    }
    k = b >> 2;
    sh = (b & 3) << 3;
    return (w[k * 4 + plane] >> sh) & 0xffu;
  }
};
enum { PL_D = 0, PL_C = 1, PL_F = 2 };

// Warp-cooperative staging: for every lane j with has_window, read the WIN x WIN traversability window whose
// unwrapped origin is (w0,w1) of lane j and store its D/C/F bit-planes at smem[j*stride ...] (4 words per load
// iteration: D, C, F, pad — one 16-byte store by lane 0).
// Cells outside the map read as NaN (=> D=C=F=0), which is exactly "not visited" for every iterator.
//
// Interior windows that do not cross the circular-buffer seam (the common case) go through a software pipeline:
// the loads of the NEXT window (up to kStageChunk coalesced column loads per lane) are issued before the ballots
// of the current one, so the L2/HBM latency of a window is covered by the bit-plane reduction of its predecessor.
// Bits of rows/columns beyond WIN then hold real neighbouring cells, which no consumer ever reads.
// Windows touching the map border or the seam take the generic per-cell path.
constexpr int kStageChunk = 8;
#ifndef FHP_STAGE_BATCH
#define FHP_STAGE_BATCH 4
#endif
constexpr int kStageBatch = FHP_STAGE_BATCH;  // class-plane entries a lane has in flight while staging its window

template <class GV>
__device__ __forceinline__ void stage_chunk_ballots(const GV& G, uint32_t dst, int lane, int k0, const float (&v)[kStageChunk],
                                                    float tD, float tC) {
  const float inf = __int_as_float(0x7f800000);
#pragma unroll
  for (int i = 0; i < kStageChunk; ++i) {
    if (k0 + i < G.nIter()) {
      const uint32_t bD = __ballot_sync(kFull, v[i] < tD);
      const uint32_t bC = __ballot_sync(kFull, v[i] < tC);
      const uint32_t bF = __ballot_sync(kFull, fabsf(v[i]) < inf);
      if (lane == 0) sts_v4(dst + static_cast<uint32_t>(k0 + i) * 16u, bD, bC, bF, 0u);
    }
  }
}

// Generic per-cell staging (windows touching the map border or the circular-buffer seam).
template <bool MASKL, class GV>
__device__ __noinline__ void stage_slow(const MapDev& M, const GV& G, uint32_t smemWarp, int lane, unsigned slowMask,
                                        int w0, int w1, float thrD, float thrC) {
  const int RP = 1 << G.rpShift();
  const int a = lane & (RP - 1);
  const int h = lane >> G.rpShift();
  const float inf = __int_as_float(0x7f800000);
  for (unsigned m = slowMask; m; m &= m - 1) {
    const int j = __ffs(m) - 1;
    const int jw0 = __shfl_sync(kFull, w0, j);
    const int jw1 = __shfl_sync(kFull, w1, j);
    const float tD = __shfl_sync(kFull, thrD, j);
    const float tC = __shfl_sync(kFull, thrC, j);
    const uint32_t dst = smemWarp + static_cast<uint32_t>(j * G.stride()) * 4u;
    const int u0 = jw0 + a;
    const bool in0 = u0 >= 0 && u0 < M.rows;  // all RP rows: the TMA path may have shifted the origin down by <= 3
    const int b0 = wrap_add(u0, M.start0, M.rows);
    for (int k = 0; k < G.nIter(); ++k) {
      const int b = k * G.pack() + h;
      const int u1 = jw1 + b;
      const bool in = in0 && (b < G.WIN()) && u1 >= 0 && u1 < M.cols;
      const int b1 = wrap_add(u1, M.start1, M.cols);
      float v = __int_as_float(0x7fc00000);
      if (in) v = __ldg(M.trav + static_cast<size_t>(b1) * M.rows + b0);
      const uint32_t bD = __ballot_sync(kFull, v < tD);
      const uint32_t bC = __ballot_sync(kFull, v < tC);
      const uint32_t bF = __ballot_sync(kFull, fabsf(v) < inf);
      if (MASKL) {  // mask path (pack 1, one column per iteration): the class-plane pair (e0, e1) — no -inf cells and
        // thrC <= thrD on this path, so D and C are subsets of F: e1 = D, e0 = (F & ~D) | C
        if (lane == 0) sts_v2(dst + static_cast<uint32_t>(k) * 8u, (bF & ~bD) | bC, bD);
      } else if (lane == 0) {
        sts_v4(dst + static_cast<uint32_t>(k) * 16u, bD, bC, bF, 0u);
      }
    }
  }
}

template <class GV>
__device__ __forceinline__ void stage_windows(const MapDev& M, const GV& G, uint32_t smemWarp, int lane,
                                              bool has_window, int w0, int w1, float thrD, float thrC) {
  const int RP = 1 << G.rpShift();
  const int a = lane & (RP - 1);
  const int h = lane >> G.rpShift();
  const int colsCovered = G.nIter() * G.pack();
  // every lane classifies ITS OWN window and computes its buffer origin (|w| <= 2^30 + HW: no overflow)
  int bb0 = w0 + M.start0, bb1 = w1 + M.start1;
  if (bb0 >= M.rows) bb0 -= M.rows;
  if (bb1 >= M.cols) bb1 -= M.cols;
  const bool myFast = has_window && M.fastOk && w0 >= 0 && w1 >= 0 && w0 <= M.rows - RP && w1 <= M.cols - colsCovered &&
                      bb0 <= M.rows - RP && bb1 <= M.cols - colsCovered;
  const int org = myFast ? bb1 * M.rows + bb0 : 0;  // rows*cols < 2^31 is part of fastOk
  unsigned fastMask = __ballot_sync(kFull, myFast);
  const unsigned slowMask = __ballot_sync(kFull, has_window && !myFast);
  const unsigned act = fastMask | slowMask;
  if (act == 0) {
    __syncwarp();
    return;
  }
  const int first = __ffs(act) - 1;
  const float tD0 = __shfl_sync(kFull, thrD, first), tC0 = __shfl_sync(kFull, thrC, first);
  const bool uni = __all_sync(kFull, !has_window || (thrD == tD0 && thrC == tC0));
  __syncwarp();

  if (fastMask) {
    const int nChunks = (G.nIter() + kStageChunk - 1) / kStageChunk;
    const long long step = static_cast<long long>(G.pack()) * M.rows;
    const float* laneBase = M.trav + static_cast<long long>(h) * M.rows + a;
    float nxt[kStageChunk], cur[kStageChunk];
    int jn = __ffs(fastMask) - 1, cn = 0;
    {
      const float* p = laneBase + __shfl_sync(kFull, org, jn);
#pragma unroll
      for (int i = 0; i < kStageChunk; ++i) nxt[i] = (i < G.nIter()) ? __ldg(p + i * step) : 0.0f;
    }
    for (;;) {
      const int j = jn, c = cn;
#pragma unroll
      for (int i = 0; i < kStageChunk; ++i) cur[i] = nxt[i];
      bool more = true;
      if (cn + 1 < nChunks) {
        cn += 1;
      } else {
        fastMask &= fastMask - 1;
        if (fastMask) {
          jn = __ffs(fastMask) - 1;
          cn = 0;
        } else {
          more = false;
        }
      }
      if (more) {  // prefetch the next unit before reducing the current one
        const float* p = laneBase + __shfl_sync(kFull, org, jn) + static_cast<long long>(cn) * kStageChunk * step;
        const int k0n = cn * kStageChunk;
#pragma unroll
        for (int i = 0; i < kStageChunk; ++i) nxt[i] = (k0n + i < G.nIter()) ? __ldg(p + i * step) : 0.0f;
      }
      const float tD = uni ? tD0 : __shfl_sync(kFull, thrD, j);
      const float tC = uni ? tC0 : __shfl_sync(kFull, thrC, j);
      stage_chunk_ballots(G, smemWarp + static_cast<uint32_t>(j * G.stride()) * 4u, lane, c * kStageChunk, cur, tD, tC);
      if (!more) break;
    }
  }

  stage_slow<false>(M, G, smemWarp, lane, slowMask, w0, w1, thrD, thrC);
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------ TMA staging
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// One 2-D tiled TMA load: box (boxDim of the tensor map) at element coordinates (c0 = row, c1 = col) of the
// column-major layer -> shared memory; completion is signalled on the mbarrier as transferred bytes.  Elements
// outside the tensor are filled with NaN by the tensor map (CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA).
__device__ __forceinline__ void tma_load_box(uint32_t dst, const void* tmap, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}

// Box (floats in shared memory) -> D/C/F bit-planes.  Load iteration k of the window is box word k*32 + lane: a
// linear, conflict-free read.  Unrolled by 4 without per-iteration predicates (every ballot is a convergence
// point, so a predicated body would cost a branch pair per iteration).
__device__ __forceinline__ void convert_one(uint32_t src, uint32_t dst, int lane, float tD, float tC) {
  const float v = lds_f32(src);
  const uint32_t bD = __ballot_sync(kFull, v < tD);
  const uint32_t bC = __ballot_sync(kFull, v < tC);
  const uint32_t bF = __ballot_sync(kFull, fabsf(v) < __int_as_float(0x7f800000));
  if (lane == 0) sts_v4(dst, bD, bC, bF, 0u);
}
template <int N>
__device__ __forceinline__ void convert_n(uint32_t src, uint32_t dst, int lane, float tD, float tC) {
  float v[N];
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = lds_f32(src + static_cast<uint32_t>(i) * 128u);  // all loads first: ILP
  const float inf = __int_as_float(0x7f800000);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const uint32_t bD = __ballot_sync(kFull, v[i] < tD);
    const uint32_t bC = __ballot_sync(kFull, v[i] < tC);
    const uint32_t bF = __ballot_sync(kFull, fabsf(v[i]) < inf);
    if (lane == 0) sts_v4(dst + static_cast<uint32_t>(i) * 16u, bD, bC, bF, 0u);
  }
}
__device__ __forceinline__ void convert_box(int nIter, uint32_t src, uint32_t dst, int lane, float tD, float tC) {
  switch (nIter) {  // straight-line bodies for the common window sizes (every ballot is a convergence point)
    case 6: convert_n<6>(src, dst, lane, tD, tC); return;
    case 7: convert_n<7>(src, dst, lane, tD, tC); return;
    case 8: convert_n<8>(src, dst, lane, tD, tC); return;
    default: break;
  }
  int k = 0;
  for (; k + 4 <= nIter; k += 4)
    convert_n<4>(src + static_cast<uint32_t>(k) * 128u, dst + static_cast<uint32_t>(k) * 16u, lane, tD, tC);
  const int rem = nIter - k;  // 0..3, straight-line
  if (rem >= 1) convert_one(src + static_cast<uint32_t>(k) * 128u, dst + static_cast<uint32_t>(k) * 16u, lane, tD, tC);
  if (rem >= 2) convert_one(src + static_cast<uint32_t>(k + 1) * 128u, dst + static_cast<uint32_t>(k + 1) * 16u, lane, tD, tC);
  if (rem >= 3) convert_one(src + static_cast<uint32_t>(k + 2) * 128u, dst + static_cast<uint32_t>(k + 2) * 16u, lane, tD, tC);
}

struct TmaWarp {       // per-warp TMA staging state (lives in registers across the whole plan)
  uint32_t bar[2];     // shared-space addresses of the two mbarriers (one per round buffer)
  uint32_t boxAddr;    // shared-space address of the warp's box buffers: [2][tmaB][boxWords] floats
  uint32_t parity[2];
};

// TMA variant of stage_windows (plan_kernel): every eligible window is fetched by ONE cp.async.bulk.tensor (issued by
// an elected lane with the window's broadcast coordinates) into the warp's shared-memory box buffers,
// two rounds of tmaB windows in flight (mbarrier per buffer).  The box lands as [col][row] floats, i.e. exactly
// the order in which the cooperative reduction wants it: load iteration k of the window is box word k*32 + lane,
// a conflict-free linear read.  Eligible = the window does not wrap around the circular-buffer seam and, when the
// start index is non-zero, lies inside the map (with start index (0,0) buffer coordinates equal map coordinates,
// so the tensor map's NaN out-of-bounds fill IS grid_map's "cell outside the map is not visited").
template <class GV>
__device__ __forceinline__ void stage_windows_tma(const MapDev& M, const GV& G, const void* tmap, TmaWarp& T,
                                                  uint32_t smemWarp, int lane, bool has_window, int w0, int w1,
                                                  float thrD, float thrC) {
  const int RP = 1 << G.rpShift();
  const int colsCovered = G.nIter() * G.pack();
  const bool start0 = (M.start0 == 0) && (M.start1 == 0);
  // a window with no cell inside the map stages as all-"not visited": the owning lane just clears its planes
  const bool empty = has_window && (w0 >= M.rows || w1 >= M.cols || w0 <= -RP || w1 <= -colsCovered);
  int c0 = w0, c1 = w1;
  bool elig = has_window && !empty;
  if (!start0) {
    c0 = w0 + M.start0;
    c1 = w1 + M.start1;
    if (c0 >= M.rows) c0 -= M.rows;
    if (c1 >= M.cols) c1 -= M.cols;
    elig = elig && w0 >= 0 && w1 >= 0 && w0 <= M.rows - RP && w1 <= M.cols - colsCovered && c0 <= M.rows - RP &&
           c1 <= M.cols - colsCovered;
  }
  const unsigned fastMask = __ballot_sync(kFull, elig);
  const unsigned slowMask = __ballot_sync(kFull, has_window && !elig && !empty);
  const unsigned act = __ballot_sync(kFull, has_window);
  if (act == 0) {
    __syncwarp();
    return;
  }
  const int first = __ffs(act) - 1;
  const float tD0 = __shfl_sync(kFull, thrD, first), tC0 = __shfl_sync(kFull, thrC, first);
  const bool uni = __all_sync(kFull, !has_window || (thrD == tD0 && thrC == tC0));
  __syncwarp();  // every lane is done reading the planes of the previous leg/cycle
  if (empty) {
    const uint32_t mine = smemWarp + static_cast<uint32_t>(lane * G.stride()) * 4u;
    for (int k = 0; k < G.nIter(); ++k) sts_v4(mine + static_cast<uint32_t>(k) * 16u, 0u, 0u, 0u, 0u);
  }

  const int nFast = __popc(fastMask);
  if (nFast) {
    const int B = 1 << G.tmaBShift();
    const int nRounds = (nFast + B - 1) >> G.tmaBShift();
    const uint32_t boxBytes = static_cast<uint32_t>(G.boxWords()) * 4u;
    const float inf = __int_as_float(0x7f800000);
    // every eligible lane issues the box of ITS window; cp.async.bulk.tensor is a warp-uniform instruction (UTMALDG),
    // so the hardware loop ptxas wraps around it (ELECT / R2UR.BROADCAST / BRA.U.ANY) serialises the active lanes
    const int rank = __popc(fastMask & ((1u << lane) - 1u));
    const int myRound = rank >> G.tmaBShift();
    const uint32_t myBox = T.boxAddr + static_cast<uint32_t>(((myRound & 1) << G.tmaBShift()) + (rank & (B - 1))) * boxBytes;
    auto issue = [&](int r) {
      const int buf = r & 1;
      const int cnt = min(B, nFast - (r << G.tmaBShift()));
      if (lane == 0) mbar_expect_tx(T.bar[buf], static_cast<uint32_t>(cnt) * boxBytes);
      if (elig && myRound == r) tma_load_box(myBox, tmap, c0, c1, T.bar[buf]);
    };
    // the box buffers were last read through the generic proxy; order those reads before the async-proxy writes
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    unsigned rem = fastMask;
    for (int r = -1; r < nRounds; ++r) {  // round r+1 is issued before round r is reduced (one copy of each body)
      if (r + 1 < nRounds) issue(r + 1);
      if (r < 0) continue;
      const int buf = r & 1;
      while (!mbar_try_wait(T.bar[buf], T.parity[buf])) {
      }
      T.parity[buf] ^= 1u;
      const int cnt = min(B, nFast - (r << G.tmaBShift()));
      uint32_t src = T.boxAddr + static_cast<uint32_t>(buf << G.tmaBShift()) * boxBytes + static_cast<uint32_t>(lane) * 4u;
      for (int sidx = 0; sidx < cnt; ++sidx) {
        const int j = __ffs(rem) - 1;
        rem &= rem - 1;
        const float tD = uni ? tD0 : __shfl_sync(kFull, thrD, j);
        const float tC = uni ? tC0 : __shfl_sync(kFull, thrC, j);
        convert_box(G.nIter(), src, smemWarp + static_cast<uint32_t>(j * G.stride()) * 4u, lane, tD, tC);
        src += boxBytes;
      }
      __syncwarp();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
  }
  if (slowMask) stage_slow<false>(M, G, smemWarp, lane, slowMask, w0, w1, thrD, thrC);
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------ mask staging
// The thresholds are per-call constants, so the predicates of a cell are computed ONCE per map update
// (mask_build_kernel) as a 2-bit class per cell — 0 not finite, 1 finite and >= thrD, 2 finite in [thrC, thrD),
// 3 finite and < thrC (thrC <= thrD; maps holding -inf cells, whose raw compare differs from their class, and
// thrC > thrD take the float paths) — split into two bit planes e0/e1 over buffer rows.  Entry (k, c) of `mask` =
// {e0 rows 32k..32k+31, e0 rows 32k+32..32k+63, e1 same, e1 same} of buffer column c, stored k-major (index
// k*cols + c): ONE 16-byte load holds the 32 window rows starting at any row of word k, and neighbouring columns
// are contiguous.  Staging a window column = 1 load, 2 funnel shifts, 1 eight-byte store of the class pair (e0, e1);
// the consumers derive D = e1, C = e0 & e1, F = e0 | e1 where they read (Planes, pack == 0 layout: pair b at words
// 2b, 2b+1 of the lane's slot, one pad word per column behind the pairs for the finished spiral masks).  Every lane
// stages ITS OWN window — WIN independent loads in register-sized batches, no cross-lane traffic; the 8-byte variant
// of the entries (half the footprint, a second load when the window crosses a row word) and 8 lanes per window
// (fewer L1 wavefronts, shuffles to route the words) both measured slower (profiles/r2_experiments.md).
// Fast = the window lies inside the map and does not cross the circular-buffer ROW seam (the column seam is
// handled per column); everything else takes stage_slow on the float layer, exactly like the other paths.
template <class GV>
__device__ __forceinline__ void stage_windows_mask(const MapDev& M, const GV& G, const uint4* __restrict__ mask,
                                                   uint32_t smemWarp, uint32_t mine, int lane, bool has_window, int w0,
                                                   int w1, float thrD, float thrC) {
  const bool empty = has_window && (w0 >= M.rows || w1 >= M.cols || w0 <= -32 || w1 <= -G.WIN());
  int bb0 = w0 + M.start0, bb1 = w1 + M.start1;  // |w| <= 2^30 + HW: no overflow
  if (bb0 >= M.rows) bb0 -= M.rows;
  if (bb1 >= M.cols) bb1 -= M.cols;
  const bool fast = has_window && w0 >= 0 && w1 >= 0 && w0 <= M.rows - G.WIN() && w1 <= M.cols - G.WIN() &&
                    bb0 <= M.rows - G.WIN();
  const unsigned fastMask = __ballot_sync(kFull, fast);
  const unsigned slowMask = __ballot_sync(kFull, has_window && !fast && !empty);
  if (__ballot_sync(kFull, has_window) == 0u) {
    __syncwarp();
    return;
  }
  __syncwarp();  // every lane is done reading the planes of the previous leg/cycle
  if (empty) {
    for (int k = 0; k < G.WIN(); ++k) sts_v2(mine + static_cast<uint32_t>(k) * 8u, 0u, 0u);
  }
  if (fast) {  // every lane stages ITS OWN window: WIN independent 16-byte loads, no cross-lane traffic
    const uint4* src = mask + (bb0 >> 5) * M.cols;
    const int s = bb0 & 31;
    const uint32_t dst = mine;
    int col = bb1;
    for (int c = 0; c < G.WIN(); c += kStageBatch) {
      uint4 e[kStageBatch];
#pragma unroll
      for (int i = 0; i < kStageBatch; ++i) {
        e[i] = __ldg(src + col);  // (columns past WIN-1 read the next entries of the row word: valid memory, unused)
        col += 1;
        if (col >= M.cols) col -= M.cols;
      }
#pragma unroll
      for (int i = 0; i < kStageBatch; ++i)
        if (c + i < G.WIN()) {
          const uint32_t e0 = __funnelshift_r(e[i].x, e[i].y, s), e1 = __funnelshift_r(e[i].z, e[i].w, s);
          sts_v2(dst + static_cast<uint32_t>(c + i) * 8u, e0, e1);
        }
    }
  }
  if (slowMask) stage_slow<true>(M, G, smemWarp, lane, slowMask, w0, w1, thrD, thrC);
  __syncwarp();
}

// The cells a grid_map::CircleIterator(centre, r) visits, decided ONCE per foothold: bit (8*i0 + i1) of `m` <-> cell
// (tl0+i0, tl1+i1) is inside the map and its centre is within r of (cx,cy) (squared-distance test with the
// iterator's arithmetic).  The default test, the height mean and the prefetch all walk the same disc, so none of them
// repeats the double-precision work.
//
// Upstream restricts the walk to the index window [index(bound(c + r)), index(bound(c - r))]
// (CircleIterator::findSubmapParameters).  That window is redundant for membership: a cell whose centre is within r of
// the centre lies at most k = floor(r/res + 1/2) indices from the centre's own cell (uc0,uc1) — a cell d indices away
// has its centre at least (d - 1/2)*res away — and every such in-map cell is inside upstream's window (the window ends
// are the cells containing c +- r; positions outside the map are clamped towards it, cells outside it are never
// visited).  So the disc is tested on the (2k+1)^2 cells around (uc0,uc1) with no index conversion at all; the host
// picks k with a guard band and hands k = -1 (=> the generic, window-based walk below) when 2k+1 > 8.
struct DiscWin {
  int tl0, tl1, br0, br1;
  unsigned long long m;
  int generic;  // 1: m is not filled, consumers walk [tl, br] with the distance test themselves
  __device__ __forceinline__ bool small() const { return generic == 0; }
};
template <int K>
__device__ __forceinline__ unsigned long long disc_members(const MapDev& M, double cx, double cy, int tl0, int tl1,
                                                            double r2) {
  constexpr int N = 2 * K + 1;
  double dy2[N];
  unsigned colsIn = 0;
#pragma unroll
  for (int i1 = 0; i1 < N; ++i1) {
    const int u1 = tl1 + i1;
    const double dy = pos_axis(M.fcy, M.res, u1) - cy;
    dy2[i1] = dy * dy;
    if (u1 >= 0 && u1 < M.cols) colsIn |= 1u << i1;
  }
  unsigned long long m = 0ull;
#pragma unroll
  for (int i0 = 0; i0 < N; ++i0) {  // row-major bit order = the iterator's visiting order
    const int u0 = tl0 + i0;
    const double dx = pos_axis(M.fcx, M.res, u0) - cx;
    const double dx2 = dx * dx;
    unsigned rowBits = 0;
#pragma unroll
    for (int i1 = 0; i1 < N; ++i1)
      if (dx2 + dy2[i1] <= r2) rowBits |= 1u << i1;
    if (u0 >= 0 && u0 < M.rows) m |= static_cast<unsigned long long>(rowBits & colsIn) << (8 * i0);
  }
  return m;
}
__device__ __noinline__ DiscWin disc_window_generic(const MapDev& M, double cx, double cy, double r) {
  DiscWin d;
  const double tlx = bound_axis(cx + r, M.lenx, M.posx, M.halfx), tly = bound_axis(cy + r, M.leny, M.posy, M.halfy);
  const double brx = bound_axis(cx - r, M.lenx, M.posx, M.halfx), bry = bound_axis(cy - r, M.leny, M.posy, M.halfy);
  d.tl0 = raw_index_axis(tlx, M.halfx, M.posx, M.res, M.rinv);
  d.tl1 = raw_index_axis(tly, M.halfy, M.posy, M.res, M.rinv);
  d.br0 = raw_index_axis(brx, M.halfx, M.posx, M.res, M.rinv);
  d.br1 = raw_index_axis(bry, M.halfy, M.posy, M.res, M.rinv);
  d.m = 0ull;
  d.generic = 1;
  return d;
}
template <class GV>
__device__ __forceinline__ DiscWin disc_window(const MapDev& M, const GV& G, double cx, double cy, int uc0, int uc1,
                                               double r, double r2) {
  // wide discs: upstream's window, membership decided by the consumers' generic loops
  if (G.discK() < 0) return disc_window_generic(M, cx, cy, r);
  DiscWin d;
  d.generic = 0;
  const int k = G.discK();
  d.tl0 = uc0 - k;
  d.tl1 = uc1 - k;
  d.br0 = uc0 + k;
  d.br1 = uc1 + k;
  if (k == 1) {  // footRadius < 1.5 cells (the yaml geometry at 0.02 m): 3 x 3, straight-line
    d.m = disc_members<1>(M, cx, cy, d.tl0, d.tl1, r2);
  } else if (k == 2) {  // < 2.5 cells (0.01 m): 5 x 5
    d.m = disc_members<2>(M, cx, cy, d.tl0, d.tl1, r2);
  } else {
    d.m = 0ull;
    for (int i0 = 0; i0 <= 2 * k; ++i0) {  // row-major bit order = the iterator's visiting order
      const int u0 = d.tl0 + i0;
      if (u0 < 0 || u0 >= M.rows) continue;
      const double dx = pos_axis(M.fcx, M.res, u0) - cx;
      const double dx2 = dx * dx;
      unsigned rowBits = 0;
      for (int i1 = 0; i1 <= 2 * k; ++i1) {
        const int u1 = d.tl1 + i1;
        const double dy = pos_axis(M.fcy, M.res, u1) - cy;
        if (dx2 + dy * dy <= r2 && u1 >= 0 && u1 < M.cols) rowBits |= 1u << i1;
      }
      d.m |= static_cast<unsigned long long>(rowBits) << (8 * i0);
    }
  }
  return d;
}

// The elevation cells of a foot disc are the only global loads of the scalar phase whose latency nothing hides (the L1
// beside the shared-memory carve-out is too small for a prefetch to survive until they are summed).  The first four
// member cells — all of them for footRadius <= resolution — are therefore LOADED right after the disc is known, before
// the window staging, and summed after it.
struct DiscElev {
  float e0, e1, e2, e3;
};
__device__ __forceinline__ DiscElev load_disc_elev(const MapDev& M, const DiscWin& dw) {
  DiscElev v = {0.0f, 0.0f, 0.0f, 0.0f};
  unsigned long long t = dw.m;
  if (t) {
    int bit = __ffsll(static_cast<long long>(t)) - 1;
    v.e0 = load_cell(M, M.elev, dw.tl0 + (bit >> 3), dw.tl1 + (bit & 7));
    t &= t - 1;
    if (t) {
      bit = __ffsll(static_cast<long long>(t)) - 1;
      v.e1 = load_cell(M, M.elev, dw.tl0 + (bit >> 3), dw.tl1 + (bit & 7));
      t &= t - 1;
      if (t) {
        bit = __ffsll(static_cast<long long>(t)) - 1;
        v.e2 = load_cell(M, M.elev, dw.tl0 + (bit >> 3), dw.tl1 + (bit & 7));
        t &= t - 1;
        if (t) {
          bit = __ffsll(static_cast<long long>(t)) - 1;
          v.e3 = load_cell(M, M.elev, dw.tl0 + (bit >> 3), dw.tl1 + (bit & 7));
        }
      }
    }
  }
  return v;
}

// ------------------------------------------------------------------------------------------------ a9 mean height
// FootholdPlanner::getFootholdMeanHeight (FP.cpp:2520-2554) over grid_map::CircleIterator (row-major over the
// bounded window in unwrapped index space, cells with squared distance <= r^2).
__device__ __forceinline__ void height_cell(const MapDev& M, int u0, int u1, float& iHeight, float& meanHeight, int& i) {
  const float e = load_cell(M, M.elev, u0, u1);
  iHeight = (fabsf(e) < __int_as_float(0x7f800000)) ? e : 0.0f;
  if (iHeight < 10.0f) {
    i++;
    meanHeight = meanHeight + iHeight;
  }
}
__device__ __forceinline__ void height_value(float e, float& iHeight, float& meanHeight, int& i) {
  iHeight = (fabsf(e) < __int_as_float(0x7f800000)) ? e : 0.0f;
  if (iHeight < 10.0f) {
    i++;
    meanHeight = meanHeight + iHeight;
  }
}
__device__ __noinline__ float mean_height(const MapDev& M, const ParDev& P, double cx, double cy, const DiscWin& dw) {
  const int tl0 = dw.tl0, tl1 = dw.tl1, br0 = dw.br0, br1 = dw.br1;
  float iHeight = 0.0f, meanHeight = 0.0f;
  int i = 0;
  if (dw.small()) {  // membership already decided (disc_window); ascending bit order = the iterator's row-major order.
    // All member cells are LOADED first, sixteen at a time in straight-line code (independent loads in flight: one
    // memory latency per batch instead of one per cell — the height loop was 40 % of a 0.01 m plan), then accumulated.
    unsigned long long t = dw.m;
    while (t) {
      float v[16];
      int n = 0;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        v[k] = 0.0f;
        if (t) {
          const int bit = __ffsll(static_cast<long long>(t)) - 1;
          t &= t - 1;
          v[k] = load_cell(M, M.elev, tl0 + (bit >> 3), tl1 + (bit & 7));
          n = k + 1;
        }
      }
#pragma unroll
      for (int k = 0; k < 16; ++k)
        if (k < n) height_value(v[k], iHeight, meanHeight, i);
    }
  } else {
    for (int u0 = tl0; u0 <= br0; ++u0) {
      const double dx = pos_axis(M.fcx, M.res, u0) - cx;
      const double dx2 = dx * dx;
      for (int u1 = tl1; u1 <= br1; ++u1) {
        const double dy = pos_axis(M.fcy, M.res, u1) - cy;
        if (!(dx2 + dy * dy <= P.r2) || !in_map(M, u0, u1)) continue;
        height_cell(M, u0, u1, iHeight, meanHeight, i);
      }
    }
  }
  if (i != 0) {
    meanHeight = meanHeight / static_cast<float>(i);
  } else {
    meanHeight = iHeight;
  }
  return static_cast<float>(static_cast<double>(meanHeight) + P.h);
}
// getFootholdMeanHeight from the pre-loaded cells (same accumulation order); discs with more than four member cells
// or on the generic walk take the loop above.
__device__ __forceinline__ float mean_height_loaded(const MapDev& M, const ParDev& P, double cx, double cy, const DiscWin& dw,
                                                    const DiscElev& v, bool preloaded) {
  const int n = __popcll(dw.m);
  if (!preloaded || !dw.small() || n > 4) return mean_height(M, P, cx, cy, dw);
  float iHeight = 0.0f, meanHeight = 0.0f;
  int i = 0;
  if (n > 0) height_value(v.e0, iHeight, meanHeight, i);
  if (n > 1) height_value(v.e1, iHeight, meanHeight, i);
  if (n > 2) height_value(v.e2, iHeight, meanHeight, i);
  if (n > 3) height_value(v.e3, iHeight, meanHeight, i);
  if (i != 0) {
    meanHeight = meanHeight / static_cast<float>(i);
  } else {
    meanHeight = iHeight;
  }
  return static_cast<float>(static_cast<double>(meanHeight) + P.h);
}
// Foot discs of more than four cells: the member cells' elevations travel global -> shared memory by cp.async (no
// destination registers, so all of them are in flight at once and stay in flight while the window is staged); slot k of
// a lane is word (k * 32 + lane) of the warp's block.  disc_elev_fetch issues them, disc_elev_mean waits and accumulates
// in the iterator's order.  Discs with more than 16 members (or on the generic walk) take mean_height.
__device__ __forceinline__ void disc_elev_fetch(const MapDev& M, const DiscWin& dw, uint32_t slot0) {
  unsigned long long t = dw.small() ? dw.m : 0ull;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    if (t) {
      const int bit = __ffsll(static_cast<long long>(t)) - 1;
      t &= t - 1;
      const int b0 = wrap_add(dw.tl0 + (bit >> 3), M.start0, M.rows), b1 = wrap_add(dw.tl1 + (bit & 7), M.start1, M.cols);
      const float* src = M.elev + static_cast<size_t>(b1) * M.rows + b0;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(slot0 + static_cast<uint32_t>(k) * 128u), "l"(src) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ float disc_elev_mean(const MapDev& M, const ParDev& P, double cx, double cy, const DiscWin& dw,
                                                uint32_t slot0) {
  asm volatile("cp.async.wait_all;" ::: "memory");
  const int n = __popcll(dw.m);
  if (!dw.small() || n > 16) return mean_height(M, P, cx, cy, dw);
  float iHeight = 0.0f, meanHeight = 0.0f;
  int i = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k)
    if (k < n) height_value(lds_f32(slot0 + static_cast<uint32_t>(k) * 128u), iHeight, meanHeight, i);
  if (i != 0) {
    meanHeight = meanHeight / static_cast<float>(i);
  } else {
    meanHeight = iHeight;
  }
  return static_cast<float>(static_cast<double>(meanHeight) + P.h);
}
template <class GV>
__device__ __forceinline__ float mean_height(const MapDev& M, const GV& G, const ParDev& P, double cx, double cy) {
  const int uc0 = raw_index_axis(cx, M.halfx, M.posx, M.res, M.rinv), uc1 = raw_index_axis(cy, M.halfy, M.posy, M.res, M.rinv);
  return mean_height(M, P, cx, cy, disc_window(M, G, cx, cy, uc0, uc1, P.r, P.r2));
}

// Rare path of the centroid method: whole-window test and edge rows counted from global memory (submap rows
// tl0..tl0+n0-1, columns tl1..tl1+n1-1, all inside the map by construction of getSubmapInformation).
__device__ __noinline__ unsigned long long centroid_rows_slow(const MapDev& M, float thrD, int tl0, int tl1, int n0,
                                                              int n1) {
  uint32_t any = 0, edgeRel = 0;
  for (int i = 0; i < n0; ++i) {
    int cnt = 0;
    for (int c = 0; c < n1; ++c)
      if (in_map(M, tl0 + i, tl1 + c) && load_cell(M, M.trav, tl0 + i, tl1 + c) < thrD) ++cnt;
    if (cnt > 0) any |= 1u << i;
    if (2 * cnt > n1) edgeRel |= 1u << i;  // cnt > n1 * 0.5
  }
  return static_cast<unsigned long long>(any) | (static_cast<unsigned long long>(edgeRel) << 32);
}

// ------------------------------------------------------------------------------------------------ a4 centroid
struct CentroidOut {
  double x, y;  // result xy ((0,0) when the submap failed or no case matched)
  int ccase;    // 0 whole window valid, 1/2/3 reference cases, 4 no case matched, -1 submap failed
};
// FootholdPlanner::checkFootholdUseCentroidMethod (FP.cpp:1605-1997) on the staged planes of this lane.
// (w0,w1): unwrapped origin of the staged window.
// floor and "safely inside its cell" test of a fractional cell coordinate (centroid_eval's fast path)
__device__ __forceinline__ bool cell_split(double t, int& idx) {
  idx = __double2int_rd(t);
  const double fr = t - static_cast<double>(idx);
  return fr > 1e-6 && fr < 1.0 - 1e-6;
}
template <class GV>
__device__ __forceinline__ CentroidOut centroid_eval(const MapDev& M, const GV& G, const ParDev& P, const Planes& pl,
                                                     int w0, int w1, int uc0, int uc1, double px, double py, unsigned& err) {
  CentroidOut o;
  o.x = 0.0;
  o.y = 0.0;
  o.ccase = -1;
  // GridMap::getSubmap -> getSubmapInformation (GridMapMath.cpp): the cells that contain the two corners p +- 0.5*L.
  // FAST PATH, the common case: the window lies well inside the map (no clamping, every range check holds, and the
  // centre is a whole cell away from the submap's edges), so the corner cells follow from the centre's fractional cell
  // coordinate: t(p +- d) = t(p) -+ d/res up to ~1e-12 cell, and the cell index is its floor whenever that is farther
  // than 1e-6 from a cell boundary.  Anything else — near the border, a corner within 1e-6 of a boundary (which also
  // covers the submaps that come out one cell larger) — takes upstream's expressions below.
  int tl0 = 0, tl1 = 0, br0 = 0, br1 = 0;
  bool fast = false;
  {
    const double hx = (0.5 * P.rectLx) * M.rinv, hy = (0.5 * P.rectLy) * M.rinv;  // half extents in cells
    if (hx < 30.0 && hy < 30.0) {
      const int mx = static_cast<int>(hx) + 2, my = static_cast<int>(hy) + 2;
      if (uc0 >= mx && uc0 < M.rows - mx && uc1 >= my && uc1 < M.cols - my) {
        const double tcx = -(((px - M.halfx) - M.posx) * M.rinv), tcy = -(((py - M.halfy) - M.posy) * M.rinv);
        const bool a = cell_split(tcx - hx, tl0), b = cell_split(tcx + hx, br0), c = cell_split(tcy - hy, tl1),
                   d = cell_split(tcy + hy, br1);
        fast = a && b && c && d;
      }
    }
  }
  if (!fast) {
    const double tlx = bound_axis(px + 0.5 * P.rectLx, M.lenx, M.posx, M.halfx);
    const double tly = bound_axis(py + 0.5 * P.rectLy, M.leny, M.posy, M.halfy);
    if (!(within_axis(tlx, M.lenx, M.posx, M.halfx) && within_axis(tly, M.leny, M.posy, M.halfy))) return o;
    tl0 = raw_index_axis(tlx, M.halfx, M.posx, M.res, M.rinv);
    tl1 = raw_index_axis(tly, M.halfy, M.posy, M.res, M.rinv);
    const double brx = bound_axis(px - 0.5 * P.rectLx, M.lenx, M.posx, M.halfx);
    const double bry = bound_axis(py - 0.5 * P.rectLy, M.leny, M.posy, M.halfy);
    if (!(within_axis(brx, M.lenx, M.posx, M.halfx) && within_axis(bry, M.leny, M.posy, M.halfy))) return o;
    br0 = raw_index_axis(brx, M.halfx, M.posx, M.res, M.rinv);
    br1 = raw_index_axis(bry, M.halfy, M.posy, M.res, M.rinv);
    if (!in_map(M, tl0, tl1)) return o;  // getPositionFromIndex(topLeftCorner) range check
  }
  const double cornx = pos_axis(M.fcx, M.res, tl0) + 0.5 * M.res;
  const double corny = pos_axis(M.fcy, M.res, tl1) + 0.5 * M.res;
  const int n0 = br0 - tl0 + 1, n1 = br1 - tl1 + 1;  // submap size (rows, cols)
  const double slx = static_cast<double>(n0) * M.res, sly = static_cast<double>(n1) * M.res;
  const double shx = 0.5 * slx, shy = 0.5 * sly;
  const double spx = cornx - shx, spy = corny - shy;  // submap position
  if (!fast && !(within_axis(px, slx, spx, shx) && within_axis(py, sly, spy, shy))) return o;  // requested centre in submap

  if (n0 < 1 || n1 < 1 || n0 > 32) {
    err |= kErrWindow;
    return o;
  }
  // window coordinates of the submap inside the staged planes
  const int a0 = tl0 - w0, a1 = br0 - w0, c0 = tl1 - w1, c1 = br1 - w1;
  uint32_t any, edgeRel;  // bit i <-> submap row i: any cell below the threshold / row is an "edge"
  if (a0 < 0 || c0 < 0 || a1 >= G.winRows() || c1 >= G.WIN()) {
    // the submap sticks out of the staged window (its size is one cell larger when the corner's lattice phase is
    // within ~1e-7 of a cell boundary): count straight from global memory
    const unsigned long long pk = centroid_rows_slow(M, P.thrD, tl0, tl1, n0, n1);
    any = static_cast<uint32_t>(pk);
    edgeRel = static_cast<uint32_t>(pk >> 32);
  } else {
    const uint32_t rowmask = (n0 >= 32 ? 0xffffffffu : ((1u << n0) - 1u)) << a0;
    // whole-window test (FP.cpp:1649-1658) and per-row counts of cells below the threshold (:1722-1750) as a
    // bit-sliced vertical counter over the window columns (policy C.1: the phantom column is not counted)
    // (counts are at most WIN: four counter bits are enough for windows of up to 15 columns)
    const bool wide = G.WIN() > 15;
    uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
    any = 0;
    for (int c = c0; c <= c1; ++c) {
      uint32_t x = pl.col(PL_D, c) & rowmask;
      any |= x;
      uint32_t t;
      t = s0 & x; s0 ^= x; x = t;
      t = s1 & x; s1 ^= x; x = t;
      t = s2 & x; s2 ^= x; x = t;
      if (wide) {
        t = s3 & x; s3 ^= x; x = t;
        s4 ^= x;
      } else {
        s3 ^= x;
      }
    }
    // row is an "edge" when count > (rightCol+1)*0.5  <=>  count > floor(n1/2) for integers (:1743)
    const int T = n1 >> 1;
    uint32_t gt = 0, eq = 0xffffffffu;
    const uint32_t sv[5] = {s0, s1, s2, s3, s4};
#pragma unroll
    for (int i = 4; i >= 0; --i) {
      if (i == 4 && !wide) continue;
      const uint32_t ti = ((T >> i) & 1) ? 0xffffffffu : 0u;
      gt |= eq & sv[i] & ~ti;
      eq &= ~(sv[i] ^ ti);
    }
    edgeRel = (gt & rowmask) >> a0;
  }
  if (any == 0) {  // wholeRegionValid (:1684-1689)
    o.x = px;
    o.y = py;
    o.ccase = 0;
    return o;
  }
  int minRow = 0, maxRow = 0;
  if (edgeRel) {
    minRow = __ffs(edgeRel) - 1;
    maxRow = 31 - __clz(edgeRel);
  }
  const int topRow = 0, bottomRow = n0 - 1, rightCol = n1 - 1;
  int newRow, newCol;
  if (minRow == topRow && maxRow != bottomRow) {  // case 1 (:1777)
    newRow = (maxRow + bottomRow + 1) >> 1;      // floor((maxRow+bottomRow+1)*0.5)
    newCol = (rightCol + 1) >> 1;                // floor((rightCol+1)*0.5)
    o.ccase = 1;
  } else if (minRow != topRow && maxRow != bottomRow) {  // case 2 (:1843)
    if ((minRow - topRow) >= (bottomRow - maxRow)) {
      newRow = (minRow + 1) >> 1;  // ceil(minRow*0.5)
    } else {
      newRow = (maxRow + bottomRow) >> 1;  // floor((maxRow+bottomRow)*0.5)
    }
    newCol = rightCol >> 1;  // floor(rightCol*0.5)
    o.ccase = 2;
  } else if (minRow != topRow && maxRow == bottomRow) {  // case 3 (:1944)
    newRow = (minRow + 1) >> 1;
    newCol = rightCol >> 1;
    o.ccase = 3;
  } else {  // minRow == topRow && maxRow == bottomRow: no branch in the reference, result stays (0,0,0)
    o.ccase = 4;
    return o;
  }
  // map.getPosition(newIndex) on the submap (start index 0): (subPos + (0.5*subLen - 0.5*res)) + res*(-idx)
  o.x = (spx + (shx - 0.5 * M.res)) + M.res * static_cast<double>(-newRow);
  o.y = (spy + (shy - 0.5 * M.res)) + M.res * static_cast<double>(-newCol);
  return o;
}

// GridMap::getSubmap's success predicate (getSubmapInformation, GridMapMath.cpp) for a request of size (lx, ly) around
// (px, py): what decides whether the reference's opt chain aborts the service call at cycle 0 (FP.cpp:931-934, :2345).
__device__ __noinline__ bool submap_request_ok(const MapDev& M, double px, double py, double lx, double ly) {
  const double tlx = bound_axis(px + 0.5 * lx, M.lenx, M.posx, M.halfx), tly = bound_axis(py + 0.5 * ly, M.leny, M.posy, M.halfy);
  if (!(within_axis(tlx, M.lenx, M.posx, M.halfx) && within_axis(tly, M.leny, M.posy, M.halfy))) return false;
  const int tl0 = raw_index_axis(tlx, M.halfx, M.posx, M.res, M.rinv), tl1 = raw_index_axis(tly, M.halfy, M.posy, M.res, M.rinv);
  const double brx = bound_axis(px - 0.5 * lx, M.lenx, M.posx, M.halfx), bry = bound_axis(py - 0.5 * ly, M.leny, M.posy, M.halfy);
  if (!(within_axis(brx, M.lenx, M.posx, M.halfx) && within_axis(bry, M.leny, M.posy, M.halfy))) return false;
  const int br0 = raw_index_axis(brx, M.halfx, M.posx, M.res, M.rinv), br1 = raw_index_axis(bry, M.halfy, M.posy, M.res, M.rinv);
  if (!in_map(M, tl0, tl1)) return false;
  const double cornx = pos_axis(M.fcx, M.res, tl0) + 0.5 * M.res, corny = pos_axis(M.fcy, M.res, tl1) + 0.5 * M.res;
  const double slx = static_cast<double>(br0 - tl0 + 1) * M.res, sly = static_cast<double>(br1 - tl1 + 1) * M.res;
  const double shx = 0.5 * slx, shy = 0.5 * sly;
  return within_axis(px, slx, cornx - shx, shx) && within_axis(py, sly, corny - shy, shy);
}

// ------------------------------------------------------------------------------------------------ a5..a8 nominal
struct NominalOut {
  double x, y;
  bool ok;
  int kind;
  int row, col;  // buffer index of the chosen cell (or of the centre's cell), (-1,-1) outside the map
};

// Polygon::isInside for the reference's axis-aligned rectangle (getSearchPolygon, FP.cpp:2496-2517) reduces to the
// half-open box  xlo <= x < xhi  &&  ylo <= y < yhi  on the cell centre.
struct Box {
  double xlo, xhi, ylo, yhi;
};
// heading extension: the rectangle is rotated (centre + ParDev::rectV), the slow paths only need its centre, which
// travels in the same struct (xlo = centre x, ylo = centre y) so that the reference-shaped kernels keep their frame
__device__ __forceinline__ Box centre_box(double rx, double ry) {
  Box b;
  b.xlo = rx;
  b.xhi = rx;
  b.ylo = ry;
  b.yhi = ry;
  return b;
}
__device__ __forceinline__ Box search_box(const ParDev& P, double rx, double ry) {
  Box b;
  b.xhi = rx + P.R;
  b.xlo = rx - P.R;
  b.yhi = ry + P.halfR;
  b.ylo = ry - P.halfR;
  return b;
}

// FootholdPlanner::checkDefaultFoothold (FP.cpp:2039-2082): CircleIterator(centre, footRadius) on the staged planes.
template <class GV>
__device__ __noinline__ bool default_eval_generic(const MapDev& M, const GV& G, const ParDev& P, const Planes& pl,
                                                  int w0, int w1, double cx, double cy, const DiscWin& dw, unsigned& err) {
  int visited = 0;
  bool below = false;
  for (int u0 = dw.tl0; u0 <= dw.br0; ++u0) {
    const double dx = pos_axis(M.fcx, M.res, u0) - cx;
    const double dx2 = dx * dx;
    for (int u1 = dw.tl1; u1 <= dw.br1; ++u1) {
      const double dy = pos_axis(M.fcy, M.res, u1) - cy;
      if (!(dx2 + dy * dy <= P.r2) || !in_map(M, u0, u1)) continue;
      ++visited;
      const int a = u0 - w0, b = u1 - w1;
      if (a < 0 || b < 0 || a >= G.winRows() || b >= G.WIN()) {
        err |= kErrWindow;
        continue;
      }
      const uint32_t bits = pl.col(PL_D, b) & pl.col(PL_F, b);
      below |= (bits >> a) & 1u;
    }
  }
  return (visited > 0) && !below;
}
template <class GV>
__device__ __forceinline__ bool default_eval(const MapDev& M, const GV& G, const ParDev& P, const Planes& pl, int w0,
                                             int w1, double cx, double cy, const DiscWin& dw, unsigned& err) {
  if (!dw.small()) return default_eval_generic(M, G, P, pl, w0, w1, cx, cy, dw, err);
  if (dw.m == 0ull) return false;  // the disc visits no cell
  const int a0 = dw.tl0 - w0, b0 = dw.tl1 - w1;
  if (a0 < 0 || b0 < 0 || dw.br0 - w0 >= G.winRows() || dw.br1 - w1 >= G.WIN()) {
    err |= kErrWindow;
    return false;
  }
  uint32_t below = 0;
  for (unsigned long long t = dw.m; t; t &= t - 1) {  // <= 4 member cells for footRadius <= resolution
    const int bit = __ffsll(static_cast<long long>(t)) - 1;
    const int a = a0 + (bit >> 3), b = b0 + (bit & 7);
    below |= ((pl.col(PL_D, b) & pl.col(PL_F, b)) >> a) & 1u;
  }
  return below == 0u;
}

// Which window rows / columns have their cell centre inside the box, with the reference's position arithmetic.
// The cell-centre coordinate pos(org + a) strictly decreases with a, so "pos < bound" is false...false,true...true
// over a = 0..n-1: first_below returns the first a where it holds (n if none), found from a reciprocal estimate and
// verified with the exact predicate on both sides (any estimate gives the right answer; a good one costs two tests).
__device__ FHP_INL_FB int first_below(double fc, double res, double rinv, int org, int n, double bound) {
  // pos(org+a) < bound  <=>  a > t with t = (fc - bound)/res - org.  The estimate below and the exact pos() each
  // carry errors of ~1e-12 cells (|indices| < 2^30 is far from reached: maps have < 2^16 cells per axis), so when
  // t is farther than 1e-6 from an integer both orderings agree and the first integer above t IS the answer; only in
  // that band (and for NaN bounds) the exact predicate decides, walking from the estimate.
  const double t = (fc - bound) * rinv - static_cast<double>(org);
  const int fl = __double2int_rd(t);  // saturating; NaN -> 0 (then not "certain": the exact walk decides)
  const double fr = t - static_cast<double>(fl);
  const bool certain = fr > 1e-6 && fr < 1.0 - 1e-6 && fabs(t) < 1.0e6;
  int a = min(max(fl, -1), n - 1) + 1;  // in [0, n]
  if (!certain) {
    while (a > 0 && pos_axis(fc, res, org + a - 1) < bound) --a;
    while (a < n && !(pos_axis(fc, res, org + a) < bound)) ++a;
  }
  return a;
}
__device__ __forceinline__ uint32_t bits_between(int lo, int hi) {  // bits lo..hi-1 (0 <= lo, hi <= 32)
  if (hi <= lo) return 0u;
  const uint32_t upTo = hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u);
  return upTo & ~((1u << lo) - 1u);
}
template <class GV>
__device__ __forceinline__ void box_masks(const MapDev& M, const GV& G, const Box& bx, int w0, int w1,
                                          uint32_t& rowsBox, uint32_t& colsBox) {
  // inside: lo <= pos < hi  <=>  first_below(hi) <= a < first_below(lo)
  rowsBox = bits_between(first_below(M.fcx, M.res, M.rinv, w0, G.winRows(), bx.xhi),
                         first_below(M.fcx, M.res, M.rinv, w0, G.winRows(), bx.xlo));
  colsBox = bits_between(first_below(M.fcy, M.res, M.rinv, w1, G.WIN(), bx.yhi),
                         first_below(M.fcy, M.res, M.rinv, w1, G.WIN(), bx.ylo));
}

// ---- heading extension: the search rectangle is a ROTATED rectangle, so grid_map::Polygon::isInside is the general
// W. R. Franklin PNPOLY (upstream's operand order; gm::Polygon::isInside in the CPU checker):
//   for every edge (i, j = i-1):  (v[i].y > p.y) != (v[j].y > p.y)  &&  p.x < (v[j].x-v[i].x)*(p.y-v[i].y)/(v[j].y-v[i].y) + v[i].x
// an odd number of such edges <=> inside.  Vertices = rectangle centre + ParDev::rectV (rotated on the host).
__device__ __noinline__ bool pnpoly_point(const ParDev& P, double rx, double ry, double X, double Y) {
  int cross = 0;
  double jx = rx + P.rectVx[3], jy = ry + P.rectVy[3];
#pragma unroll 1
  for (int i = 0; i < 4; ++i) {
    const double ix = rx + P.rectVx[i], iy = ry + P.rectVy[i];
    if ((iy > Y) != (jy > Y)) {
      if (X < div_rn((jx - ix) * (Y - iy), jy - iy) + ix) ++cross;
    }
    jx = ix;
    jy = iy;
  }
  return (cross & 1) != 0;
}
// The window rows of ONE window column (cell index u1 along y) whose cell centre is inside the polygon.  Along a column
// p.y is fixed, so every edge that straddles it contributes ONE threshold X_e (upstream's expression) and the test
// p.x < X_e; cell-centre x strictly decreases with the row, so "p.x < X_e" holds from first_below(X_e) on and the
// parity of the crossings is the XOR of those suffix masks.  Exact for any polygon.
// The division of X_e is only PAID where it can matter: X_e is first estimated from the edge's slope (sl[i] =
// (v[j].x - v[i].x) / (v[j].y - v[i].y) of the vertex offsets, derived once per parameter set on the host); the
// estimate is within a few ulp (~1e-15 m) of upstream's value, so the row index it implies is the exact one unless it
// falls within 1e-6 cell of a row boundary — then (and only then) upstream's expression and the exact walk decide.
struct PolySlopes {
  double s0, s1, s2, s3;  // edge i joins vertex i and vertex i-1 (PNPOLY's j)
};
__device__ __forceinline__ PolySlopes poly_slopes(const ParDev& P) {  // per parameter set: derived on the host
  PolySlopes s;
  s.s0 = P.rectSlope[0];
  s.s1 = P.rectSlope[1];
  s.s2 = P.rectSlope[2];
  s.s3 = P.rectSlope[3];
  return s;
}
__device__ __noinline__ uint32_t poly_col_rows(const MapDev& M, const ParDev& P, double rx, double ry, int w0, int nRows,
                                               int u1, double s0, double s1, double s2, double s3) {
  const double Y = pos_axis(M.fcy, M.res, u1);
  uint32_t m = 0u;
  double jx = rx + P.rectVx[3], jy = ry + P.rectVy[3];
#pragma unroll 1
  for (int i = 0; i < 4; ++i) {
    const double ix = rx + P.rectVx[i], iy = ry + P.rectVy[i];
    if ((iy > Y) != (jy > Y)) {
      const double sl = i == 0 ? s0 : (i == 1 ? s1 : (i == 2 ? s2 : s3));
      const double Xe = sl * (Y - iy) + ix;  // estimate of upstream's threshold
      const double t = (M.fcx - Xe) * M.rinv - static_cast<double>(w0);
      const int fl = __double2int_rd(t);
      const double fr = t - static_cast<double>(fl);
      int a;
      if (fr > 1e-6 && fr < 1.0 - 1e-6 && fabs(t) < 1.0e6) {
        a = min(max(fl, -1), nRows - 1) + 1;
      } else {  // within the guard band of a row boundary (or not finite): upstream's expression, exact predicate
        a = first_below(M.fcx, M.res, M.rinv, w0, nRows, div_rn((jx - ix) * (Y - iy), jy - iy) + ix);
      }
      if (a < 32) m ^= 0xffffffffu << a;
    }
    jx = ix;
    jy = iy;
  }
  return m;
}

// The same row masks for ALL window columns at once, edge by edge (mask staging path): `acc` = shared-space address of
// WIN words, one per column, that receive the masks.  An edge straddles the contiguous run of columns between
// first_below(v[i].y) and first_below(v[j].y) (cell-centre y decreases with the column: the straddle test
// (v[i].y > p.y) != (v[j].y > p.y) flips exactly there), and along that run the threshold's row coordinate advances
// by the edge's slope per column — so a column costs one multiply-add and the integer split of poly_col_rows; only an
// estimate within 1e-6 cell of a row boundary falls back to upstream's expression and the exact walk.
__device__ __noinline__ void poly_rows_all(const MapDev& M, const ParDev& P, double rx, double ry, int w0, int w1, int nRows,
                                           int nCols, uint32_t acc) {
  for (int b = 0; b < nCols; ++b) asm volatile("st.shared.u32 [%0], %1;" ::"r"(acc + static_cast<uint32_t>(b) * 4u), "r"(0u) : "memory");
  double jx = rx + P.rectVx[3], jy = ry + P.rectVy[3];
#pragma unroll 1
  for (int i = 0; i < 4; ++i) {
    const double ix = rx + P.rectVx[i], iy = ry + P.rectVy[i];
    // columns b with (iy > Y_b): b >= bi; with (jy > Y_b): b >= bj
    const int bi = first_below(M.fcy, M.res, M.rinv, w1, nCols, iy), bj = first_below(M.fcy, M.res, M.rinv, w1, nCols, jy);
    const int lo = min(bi, bj), hi = max(bi, bj);
    if (lo < hi) {
      const double sl = P.rectSlope[i];
      const double Y0 = pos_axis(M.fcy, M.res, w1 + lo);
      const double t0 = (M.fcx - (sl * (Y0 - iy) + ix)) * M.rinv - static_cast<double>(w0);  // row coordinate at column lo
      for (int b = lo; b < hi; ++b) {
        const double t = t0 + static_cast<double>(b - lo) * sl;  // Y falls by res per column: X_e by sl*res, t rises by sl
        const int fl = __double2int_rd(t);
        const double fr = t - static_cast<double>(fl);
        int a;
        if (fr > 1e-6 && fr < 1.0 - 1e-6 && fabs(t) < 1.0e6) {
          a = min(max(fl, -1), nRows - 1) + 1;
        } else {  // guard band of a row boundary (or not finite): upstream's expression, exact predicate
          const double Y = pos_axis(M.fcy, M.res, w1 + b);
          a = first_below(M.fcx, M.res, M.rinv, w0, nRows, div_rn((jx - ix) * (Y - iy), jy - iy) + ix);
        }
        if (a < 32) {
          const uint32_t addr = acc + static_cast<uint32_t>(b) * 4u;
          asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(lds_u32(addr) ^ (0xffffffffu << a)) : "memory");
        }
      }
    }
    jx = ix;
    jy = iy;
  }
}

// checkCirclePolygonFoothold (FP.cpp:2117-2163) for the candidate at window coordinates (qa,qb), on the planes:
// every finite cell of the foot disc must be >= candidateFootholdThreshold and inside the box.  HEADING: the owning
// lane has left the finished "bad cell" masks (polygon included) in plane 3 of its slot (nominal_eval).
template <bool HEADING, class GV>
__device__ __forceinline__ bool candidate_on_planes(const GV& G, const Planes& pl, int qa, int qb, uint32_t rowsBox,
                                                    uint32_t colsBox) {
  bool good = true;
  for (int s = 0; s < G.nSten() && good; ++s) {
    const int a = qa + G.g.sten[s][0], b = qb + G.g.sten[s][1];
    // a,b are inside the window by construction (|d| <= nRings, |sten| <= ra); cells outside the map have
    // F = 0 and are skipped like the iterator skips them
    uint32_t bad;
    if (HEADING) {
      bad = pl.col(3, b);
    } else {
      const uint32_t F = pl.col(PL_F, b), C = pl.col(PL_C, b);
      const uint32_t box = ((colsBox >> b) & 1u) ? rowsBox : 0u;
      bad = F & (C | ~box);
    }
    good = !((bad >> a) & 1u);
  }
  return good;
}

// checkCirclePolygonFoothold (FP.cpp:2117-2163) at an ARBITRARY foot centre, straight from global memory (the public
// member; the plan's own candidates sit on cell centres and use the planes): walk CircleIterator's window, every finite
// member cell must be >= candidateFootholdThreshold and inside the box; a disc without member cells is invalid.
__device__ __noinline__ bool circle_polygon_direct(const MapDev& M, const ParDev& P, double cx, double cy, const Box& bx) {
  const double xlo = bx.xlo, xhi = bx.xhi, ylo = bx.ylo, yhi = bx.yhi, rx = bx.xlo, ry = bx.ylo;  // (HEADING: centre)
  const DiscWin dw = disc_window_generic(M, cx, cy, P.r);
  bool validation = false;  // FP.cpp:2122: stays false when the iterator visits nothing
  for (int u0 = dw.tl0; u0 <= dw.br0; ++u0) {
    const double X = pos_axis(M.fcx, M.res, u0), dx = X - cx;
    const double dx2 = dx * dx;
    for (int u1 = dw.tl1; u1 <= dw.br1; ++u1) {
      const double Y = pos_axis(M.fcy, M.res, u1), dy = Y - cy;
      if (!(dx2 + dy * dy <= P.r2) || !in_map(M, u0, u1)) continue;
      validation = true;  // :2159 — every visited cell, valid or not, ends its iteration with validation = true
      const float v = load_cell(M, M.trav, u0, u1);
      if (!(fabsf(v) < __int_as_float(0x7f800000))) continue;  // :2132 invalid cells are not tested
      const bool inside = P.heading ? pnpoly_point(P, rx, ry, X, Y) : ((X >= xlo) && (X < xhi) && (Y >= ylo) && (Y < yhi));
      if (v < P.thrC || !inside) return false;  // :2138-2143 (break with validation = false)
    }
  }
  return validation;
}

// One candidate evaluated straight from global memory (faithful-mode candidates that wrap to the far side of the
// circular buffer): checkCirclePolygonFoothold (FP.cpp:2117-2163).
template <bool HEADING, class GV>
__device__ __noinline__ bool candidate_slow(const MapDev& M, const GV& G, const ParDev& P, int q0, int q1, const Box& bx) {
  const float thrC = P.thrC;
  const double xlo = bx.xlo, xhi = bx.xhi, ylo = bx.ylo, yhi = bx.yhi, rx = bx.xlo, ry = bx.ylo;  // (HEADING: centre)
  for (int s = 0; s < G.nSten(); ++s) {
    const int u0 = q0 + G.g.sten[s][0], u1 = q1 + G.g.sten[s][1];
    if (!in_map(M, u0, u1)) continue;
    const float v = load_cell(M, M.trav, u0, u1);
    if (!(fabsf(v) < __int_as_float(0x7f800000))) continue;
    const double X = pos_axis(M.fcx, M.res, u0), Y = pos_axis(M.fcy, M.res, u1);
    const bool inside = HEADING ? pnpoly_point(P, rx, ry, X, Y) : ((X >= xlo) && (X < xhi) && (Y >= ylo) && (Y < yhi));
    if (v < thrC || !inside) return false;
  }
  return true;
}

__device__ __forceinline__ void spiral_entry(const uint16_t* sSpiral, int k, int& d0, int& d1) {
  const unsigned e = sSpiral[k];
  d0 = static_cast<int>(static_cast<int8_t>(e & 0xffu));
  d1 = static_cast<int>(static_cast<int8_t>(e >> 8));
}

// Bits a (< nBits) with 0 <= org + a < n: the window rows / columns whose cell index lies inside the map.
__device__ __forceinline__ uint32_t in_range_mask(int org, int n, int nBits) {
  const int lo = max(0, -org), hi = min(nBits - 1, n - 1 - org);
  if (hi < lo) return 0u;
  return (hi - lo >= 31 ? 0xffffffffu : ((1u << (hi - lo + 1)) - 1u)) << lo;
}

// The foot-disc stencil around a cell centre is the full 3x3 block (footRadius in [sqrt(2), 2) cells: the yaml radius at
// 0.01 m): 9 offsets within one cell of the centre can only be that block.
template <class GV>
__device__ __forceinline__ bool dilated3x3(const GV& G) {
  return G.nSten() == 9 && G.ra() == 1;
}

// checkCandidateFoothold (FP.cpp:2085-2114) over grid_map::SpiralIterator — warp-cooperative rank scan.
// For every lane j with `need`:
//   1. lane b (< WIN) holds the "bad cell" mask of window column b of j's staged planes: a finite cell is bad when
//      it is below candidateFootholdThreshold or its centre is outside the search rectangle (FP.cpp:2138); NaN and
//      out-of-map cells are never bad (the CircleIterator skips them).  PAD variant (pack == 1 plane layout): lane j
//      has already computed its masks itself and left them in the pad word of each column entry (nominal_eval), so
//      this is one shared-memory load.  The masks are then dilated by the foot-disc stencil, so bit a of column b
//      says "the foot disc centred on cell (a,b) contains a bad cell", and cells whose own index is outside the map
//      are marked too (checkIfIndexInRange in SpiralIterator::generateRing).
//   2. the 32 lanes test candidates rank base..base+31 at once: table lookup, one shuffle to fetch the column mask,
//      one bit test, plus the exact last-two-rings distance filter; the first set bit of the ballot is the first
//      valid cell in spiral order and the scan stops at the first 32-block that has one.
// Handles wrap_correct mode, and faithful mode when the start index is (0,0) (identical there).
template <bool PAD, bool HEADING, class GV>
__device__ __forceinline__ void spiral_scan_coop(const MapDev& M, const GV& G, const uint16_t* sSpiral,
                                                 uint32_t smemWarp, int lane, bool need, int uc0, int uc1,
                                                 int w0, double cx, double cy, double R2, uint32_t rowsBox,
                                                 uint32_t colsBox, uint32_t rowsIn, uint32_t colsIn, bool& found,
                                                 int& fq0, int& fq1) {
  const unsigned needMask = __ballot_sync(kFull, need);
  const int haMine = uc0 - w0;  // window row of the centre cell (HW, plus the TMA alignment shift)
  for (unsigned m = needMask; m; m &= m - 1) {
    const int j = __ffs(m) - 1;
    const int ha = __shfl_sync(kFull, haMine, j);
    const uint32_t jrowsIn = __shfl_sync(kFull, rowsIn, j), jcolsIn = __shfl_sync(kFull, colsIn, j);
    // 1. column masks (lane = window column)
    uint32_t bad = 0;
    if (PAD) {
      if (lane < G.WIN()) bad = lds_u32(smemWarp + static_cast<uint32_t>(j * G.stride() + 2 * G.WIN() + lane) * 4u);
    } else if (HEADING) {  // the owning lane has left the finished masks (rotated rectangle included) in plane 3
      const Planes pj(smemWarp + static_cast<uint32_t>(j * G.stride()) * 4u, G.pack(), G.rpShift(), G.WIN());
      if (lane < G.WIN()) bad = pj.col(3, lane);
    } else {
      const uint32_t jrows = __shfl_sync(kFull, rowsBox, j), jcols = __shfl_sync(kFull, colsBox, j);
      const Planes pj(smemWarp + static_cast<uint32_t>(j * G.stride()) * 4u, G.pack(), G.rpShift(), G.WIN());
      if (lane < G.WIN()) {
        const uint32_t F = pj.col(PL_F, lane), C = pj.col(PL_C, lane);
        const uint32_t box = ((jcols >> lane) & 1u) ? jrows : 0u;
        bad = F & (C | ~box);
      }
    }
    uint32_t dil = bad;
    if (G.nSten() > 1 && !(PAD && dilated3x3(G))) {  // dilate by the foot-disc stencil (nothing to do for a single cell;
                                                      // the owning lane has done it for the 3x3 block, nominal_eval)
      dil = 0;
      for (int s = 0; s < G.nSten(); ++s) {
        const int s0 = G.g.sten[s][0], s1 = G.g.sten[s][1];
        const int srcCol = lane + s1;
        uint32_t nb = __shfl_sync(kFull, bad, srcCol & 31);
        if (srcCol < 0 || srcCol >= G.WIN()) nb = 0;
        dil |= (s0 >= 0) ? (nb >> s0) : (nb << (-s0));
      }
    }
    // candidate cells outside the map: rows (bits) and this lane's column
    dil = ((jcolsIn >> lane) & 1u) ? (dil | ~jrowsIn) : 0xffffffffu;
    // quick reject: no cell of the spiral footprint is acceptable even before the distance filter (a query that is
    // stuck on an untraversable patch re-scans the same window every remaining cycle).  With a single-cell stencil the
    // PAD variant has already made this test in the owning lane.
    if (!PAD || !(G.nSten() == 1 || dilated3x3(G))) {
      const uint32_t footprint = (lane < G.WIN()) ? (G.g.candRows[lane] << (ha - G.HW())) : 0u;
      if (__ballot_sync(kFull, (footprint & ~dil) != 0u) == 0u) continue;
    }
    // 2. rank scan
    bool haveC = false;
    double jcx = 0, jcy = 0, jR2 = 0;
    int juc0 = 0, juc1 = 0;
    for (int base = 0; base < G.nCand(); base += 32) {
      const int k = base + lane;
      bool cand = k < G.nCand();
      int d0, d1;
      spiral_entry(sSpiral, cand ? k : 0, d0, d1);
      const int qa = d0 + ha, qb = d1 + G.HW();
      const uint32_t e = __shfl_sync(kFull, dil, qb);
      cand = cand && !((e >> qa) & 1u);
      if (base + 32 > G.lastTwoStart()) {  // rings nRings-1, nRings: distance filter (uniform branch)
        if (!haveC) {
          jcx = __shfl_sync(kFull, cx, j);
          jcy = __shfl_sync(kFull, cy, j);
          jR2 = __shfl_sync(kFull, R2, j);
          juc0 = __shfl_sync(kFull, uc0, j);
          juc1 = __shfl_sync(kFull, uc1, j);
          haveC = true;
        }
        if (cand && k >= G.lastTwoStart()) {
          const double dx = pos_axis(M.fcx, M.res, juc0 + d0) - jcx, dy = pos_axis(M.fcy, M.res, juc1 + d1) - jcy;
          cand = (dx * dx + dy * dy <= jR2);
        }
      }
      const unsigned g = __ballot_sync(kFull, cand);
      if (g) {
        const int src = __ffs(g) - 1;
        const int wd0 = __shfl_sync(kFull, d0, src), wd1 = __shfl_sync(kFull, d1, src);
        if (lane == j) {
          found = true;
          fq0 = uc0 + wd0;
          fq1 = uc1 + wd1;
        }
        break;
      }
    }
  }
}

// The same scan for FHP_SPIRAL_FAITHFUL with a non-zero start index (bug-compatible with upstream):
// the centre index is the BUFFER index, ring neighbours are taken in buffer space without wrapping, and the
// distance filter evaluates positions as if the start index were 0.  Per-lane scalar loop (compatibility path).
template <bool HEADING, class GV>
__device__ __noinline__ void spiral_scan_faithful(const MapDev& M, const GV& G, const ParDev& P, const uint16_t* sSpiral,
                                                  const Planes& pl, int uc0, int uc1, int w0, double cx,
                                                  double cy, double R2, const Box& bx, uint32_t rowsBox,
                                                  uint32_t colsBox, bool& found, int& fq0, int& fq1) {
  int cb0 = uc0 + M.start0, cb1 = uc1 + M.start1;  // getBufferIndexFromIndex + wrapIndexToRange
  if (cb0 < 0) cb0 += ((-cb0 / M.rows) + 1) * M.rows;
  cb0 %= M.rows;
  if (cb1 < 0) cb1 += ((-cb1 / M.cols) + 1) * M.cols;
  cb1 %= M.cols;
  for (int k = 0; k < G.nCand(); ++k) {
    int d0, d1;
    spiral_entry(sSpiral, k, d0, d1);
    const int bi0 = cb0 + d0, bi1 = cb1 + d1;  // buffer index, NOT wrapped
    if (!in_map(M, bi0, bi1)) continue;
    if (k >= G.lastTwoStart()) {  // upstream evaluates the position as if the start index were 0
      const double dx = pos_axis(M.fcx, M.res, bi0) - cx, dy = pos_axis(M.fcy, M.res, bi1) - cy;
      if (!(dx * dx + dy * dy <= R2)) continue;
    }
    int q0 = bi0 - M.start0, q1 = bi1 - M.start1;  // getIndexFromBufferIndex
    if (q0 < 0) q0 += M.rows;
    if (q1 < 0) q1 += M.cols;
    const bool far = (q0 != uc0 + d0) || (q1 != uc1 + d1);
    const bool good = far ? candidate_slow<HEADING>(M, G, P, q0, q1, bx)
                          : candidate_on_planes<HEADING>(G, pl, d0 + (uc0 - w0), d1 + G.HW(), rowsBox, colsBox);
    if (good) {
      found = true;
      fq0 = q0;
      fq1 = q1;
      return;
    }
  }
}

// FootholdPlanner::checkFoothold (FP.cpp:2001-2036) = checkDefaultFoothold (:2039-2082), then
// checkCandidateFoothold (:2085-2114) over grid_map::SpiralIterator with checkCirclePolygonFoothold (:2117-2163).
// WARP-COLLECTIVE: every lane of the warp must call it (inactive lanes pass active = false).
// (cx,cy): search centre; (rx,ry): centre of the search rectangle; (uc0,uc1): raw unwrapped index of the centre;
// (w0,w1): unwrapped origin of the staged window (uc-HW, minus the TMA row-alignment shift); mine: shared-space
// address of this lane's plane slot (smemWarp + lane * stride words).
template <bool DEFAULT_ONLY, bool PAD, bool HEADING, class GV>
__device__ __forceinline__ NominalOut nominal_eval(const MapDev& M, const GV& G, const ParDev& P,
                                                   const uint16_t* sSpiral, uint32_t smemWarp, uint32_t mine, int lane,
                                                   bool active, int uc0, int uc1, int w0, int w1, double cx, double cy,
                                                   double rx, double ry, bool okDefault, unsigned& err) {
  NominalOut o;
  o.x = cx;  // FP.cpp:2016-2017
  o.y = cy;
  o.ok = false;
  o.kind = FHP_KIND_DEFAULT;
  o.row = o.col = -1;
  const Planes pl(mine, PAD ? 0 : G.pack(), G.rpShift(), G.WIN());  // PAD <=> mask staging: the class-pair layout
  if (active) {
    const bool centre_in_map = within_axis(cx, M.lenx, M.posx, M.halfx) && within_axis(cy, M.leny, M.posy, M.halfy) &&
                               in_map(M, uc0, uc1);
    o.row = centre_in_map ? wrap_add(uc0, M.start0, M.rows) : -1;
    o.col = centre_in_map ? wrap_add(uc1, M.start1, M.cols) : -1;
    o.ok = okDefault;  // checkDefaultFoothold (default_eval), evaluated by the caller
  }
  if (DEFAULT_ONLY) return o;

  const bool need = active && !o.ok;
  const bool faithful = (P.spiralMode == FHP_SPIRAL_FAITHFUL) && !(M.start0 == 0 && M.start1 == 0);
  uint32_t rowsBox = 0, colsBox = 0, rowsIn = 0, colsIn = 0;
  Box bx = {0, 0, 0, 0};
  PolySlopes ps = {0, 0, 0, 0};
  bool found = false;
  int fq0 = 0, fq1 = 0;
  bool pend = need && !faithful;
  if (need) {
    o.kind = FHP_KIND_CANDIDATE;
    if (HEADING) {
      bx = centre_box(rx, ry);
      ps = poly_slopes(P);
    } else {
      bx = search_box(P, rx, ry);
      box_masks(M, G, bx, w0, w1, rowsBox, colsBox);
    }
    rowsIn = in_range_mask(w0, M.rows, 32);
    colsIn = in_range_mask(w1, M.cols, G.WIN());
    if (HEADING && !PAD) {
      // rotated rectangle on the packed plane layouts: the owning lane builds the finished "bad cell" masks
      // (finite & (below the candidate threshold | centre outside the polygon)) into plane 3 of its slot
      const int rp = 1 << G.rpShift();  // rows per packed column (32 / pack)
      const uint32_t rm = rp >= 32 ? 0xffffffffu : ((1u << rp) - 1u);
      for (int k = 0; k < G.nIter(); ++k) {
        uint32_t word = 0u;
        for (int c = 0; c < G.pack(); ++c) {
          const int b = k * G.pack() + c;
          if (b >= G.WIN()) break;
          const uint32_t inside = poly_col_rows(M, P, rx, ry, w0, G.winRows(), w1 + b, ps.s0, ps.s1, ps.s2, ps.s3);
          const uint32_t bad = pl.col(PL_F, b) & (pl.col(PL_C, b) | ~inside) & rm;
          word |= bad << (c * rp);
        }
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(mine + static_cast<uint32_t>(k * 4 + 3) * 4u), "r"(word) : "memory");
      }
    }
    if (PAD && (pend || HEADING)) {
      // The owning lane turns its planes into per-column "bad cell" masks once (left in the pad word of each column
      // entry for the cooperative scan) — already DILATED by the foot-disc stencil when that is a single cell or the
      // full 3x3 block (separable: a vertical OR per column, then the OR of the three neighbouring columns, carried in
      // registers while the loop slides over the window) — and then settles the two most common outcomes itself: no
      // acceptable cell anywhere in the spiral footprint (a stuck query), or the centre cell is valid.
      const int sh = (uc0 - w0) - G.HW();
      // rotated rectangle: the rows inside it, per window column, into the pad words first (each is read below before the
      // finished mask of its column — or, with the 3x3 dilation, of the column before it — overwrites it)
      if (HEADING) poly_rows_all(M, P, rx, ry, w0, w1, G.winRows(), G.WIN(), mine + static_cast<uint32_t>(2 * G.WIN()) * 4u);
      const bool dil3 = dilated3x3(G) && pend;  // (bug-compatible spiral lanes test stencil cells one by one: raw masks)
      uint32_t anyOk = 0, okCentre = 0;
      uint32_t vPrev = 0u, vCur = 0u;  // vertically dilated masks of columns b-2 and b-1
      const int nLoop = G.WIN() + (dilated3x3(G) ? 1 : 0);
      for (int b = 0; b < nLoop; ++b) {
        uint32_t bad = 0u;
        if (b < G.WIN()) {
          const uint2 e = lds_v2(mine + static_cast<uint32_t>(b) * 8u);  // class planes (e0, e1) of column b
          const uint32_t box = HEADING ? lds_u32(mine + static_cast<uint32_t>(2 * G.WIN() + b) * 4u)  // poly_rows_all
                                       : (((colsBox >> b) & 1u) ? rowsBox : 0u);
          bad = (e.x | e.y) & ((e.x & e.y) | ~box);  // finite & (below the candidate threshold | outside the rectangle)
        }
        int col = b;          // the column whose finished mask this iteration emits
        uint32_t out = bad;
        if (dil3) {
          const uint32_t vNext = bad | (bad << 1) | (bad >> 1);
          col = b - 1;
          out = vPrev | vCur | vNext;
          vPrev = vCur;
          vCur = vNext;
        }
        if (col >= 0 && col < G.WIN()) {
          asm volatile("st.shared.u32 [%0], %1;" ::"r"(mine + static_cast<uint32_t>(2 * G.WIN() + col) * 4u), "r"(out) : "memory");
          const uint32_t okb = ((colsIn >> col) & 1u) ? ((G.g.candRows[col] << sh) & ~out & rowsIn) : 0u;
          anyOk |= okb;
          if (col == G.HW()) okCentre = okb;
        }
      }
      if (pend) {
        const bool finished = G.nSten() == 1 || dilated3x3(G);  // the masks above are the dilated ones
        if (finished && G.lastTwoStart() > 0 && ((okCentre >> (uc0 - w0)) & 1u)) {  // rank 0: the centre cell itself
          found = true;
          fq0 = uc0;
          fq1 = uc1;
        }
        if (finished && (found || anyOk == 0u)) pend = false;
      }
    }
  }
  if (PAD || HEADING) __syncwarp();  // pad words visible to the cooperating lanes
  spiral_scan_coop<PAD, HEADING>(M, G, sSpiral, smemWarp, lane, pend, uc0, uc1, w0, cx, cy, P.R2, rowsBox, colsBox, rowsIn,
                                 colsIn, found, fq0, fq1);
  if (need && faithful)
    spiral_scan_faithful<HEADING>(M, G, P, sSpiral, pl, uc0, uc1, w0, cx, cy, P.R2, bx, rowsBox, colsBox, found, fq0, fq1);
  if (found) {
    o.ok = true;
    o.x = pos_axis(M.fcx, M.res, fq0);  // gridmap.getPosition(index) (FP.cpp:2105-2107)
    o.y = pos_axis(M.fcy, M.res, fq1);
    o.row = wrap_add(fq0, M.start0, M.rows);
    o.col = wrap_add(fq1, M.start1, M.cols);
  }
  return o;
}

}  // namespace fhp
