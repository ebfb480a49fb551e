// libfhp kernels (sm_100a).  See fhp_device.cuh for the design of the device functions.
//
// plan_kernel: one LANE per (query, leg); the 4 lanes of a quad own the RF,RH,LH,LF feet of one query and carry
// BOTH chains of the reference's cycle loop (FP.cpp:762-1579) in registers across all gait cycles, so one launch
// plans every cycle of every query (the sequential gait-cycle dependency never leaves the SM).  A warp = up to 8
// queries (fewer when the batch cannot fill the GPU); the only cross-lane traffic is quad shuffles (stance centroid,
// 4-way AND), the warp-cooperative window staging and the cooperative spiral rank scan.
// mask_build_kernel: the per-map cell-class planes the default staging path reads.
// scan_rows_kernel / compact_rows_kernel: the compact (service-shaped) response of fhp_plan_compact.
// task_kernel: the reference's fine-grained public members, one lane per call.
#include "fhp_device.cuh"
#include "fhp_host.h"

namespace fhp {

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(kFull, v, src); }
__device__ __forceinline__ double mulz(double a, double t) { return t == 0.0 ? 0.0 : a * t; }

// FootholdPlanner::getPolygonCenter (FP.cpp:2421-2463), xy only (z never reaches the response).
__device__ FHP_INL_PC void polygon_center(double rfx, double rfy, double rhx, double rhy, double lhx, double lhy,
                                               double lfx, double lfy, double& cx, double& cy) {
  double x1 = rfx, y1 = rfy, x2 = rhx, y2 = rhy, x3, y3;
  double sum_x = 0, sum_y = 0, sum_s = 0;
  x3 = lhx;
  y3 = lhy;
  double s = ((x2 - x1) * (y3 - y1) - (x3 - x1) * (y2 - y1)) * 0.5;  // x/2.0 == x*0.5 exactly
  sum_x += (x1 + x2 + x3) * s;
  sum_y += (y1 + y2 + y3) * s;
  sum_s += s;
  x2 = x3;
  y2 = y3;
  x3 = lfx;
  y3 = lfy;
  s = ((x2 - x1) * (y3 - y1) - (x3 - x1) * (y2 - y1)) * 0.5;
  sum_x += (x1 + x2 + x3) * s;
  sum_y += (y1 + y2 + y3) * s;
  sum_s += s;
  cx = div3_rn(div_rn(sum_x, sum_s));
  cy = div3_rn(div_rn(sum_y, sum_s));
}

// The same centroid for BOTH chains of a quad at once: every lane forms the six sums (cheap), then each of the quad's
// four lanes runs ONE of the four division chains (centroid x / y, nominal x / y) and the quotients are exchanged —
// one pass through the division subroutine per cycle instead of four.  Same operations on the same operands as
// polygon_center, so the results are bit-identical.  WARP-COLLECTIVE.
__device__ __forceinline__ void polygon_sums(double x1, double y1, double x2, double y2, double x3, double y3, double x4,
                                             double y4, double& sum_x, double& sum_y, double& sum_s) {
  const double s1 = ((x2 - x1) * (y3 - y1) - (x3 - x1) * (y2 - y1)) * 0.5;
  const double s2 = ((x3 - x1) * (y4 - y1) - (x4 - x1) * (y3 - y1)) * 0.5;
  sum_x = (0.0 + (x1 + x2 + x3) * s1) + (x1 + x3 + x4) * s2;
  sum_y = (0.0 + (y1 + y2 + y3) * s1) + (y1 + y3 + y4) * s2;
  sum_s = (0.0 + s1) + s2;
}
// `dead` (uniform over the quad, see plan_kernel): the quotients are not needed any more — NaN without the subroutine,
// whose special-operand path (a zero area divides 0/0) is several times longer than the division itself.
// (Without a heading only the x coordinates of the two centres are used, FP.cpp:2199-2201: lanes 0/1 and 2/3 then run
// the same chain and the y sums are never formed.)
template <bool HEADING>
__device__ __forceinline__ void quad_centres(double cenx, double ceny, double nomx, double nomy, int quad0, int f, bool dead,
                                             double& ccx, double& ccy, double& ncx, double& ncy) {
  double csx, csy, css, nsx, nsy, nss;
  polygon_sums(shfl_d(cenx, quad0), shfl_d(ceny, quad0), shfl_d(cenx, quad0 + 1), shfl_d(ceny, quad0 + 1),
               shfl_d(cenx, quad0 + 2), shfl_d(ceny, quad0 + 2), shfl_d(cenx, quad0 + 3), shfl_d(ceny, quad0 + 3), csx, csy, css);
  polygon_sums(shfl_d(nomx, quad0), shfl_d(nomy, quad0), shfl_d(nomx, quad0 + 1), shfl_d(nomy, quad0 + 1),
               shfl_d(nomx, quad0 + 2), shfl_d(nomy, quad0 + 2), shfl_d(nomx, quad0 + 3), shfl_d(nomy, quad0 + 3), nsx, nsy, nss);
  const double num = HEADING ? ((f & 2) ? ((f & 1) ? nsy : nsx) : ((f & 1) ? csy : csx)) : ((f & 2) ? nsx : csx);
  const double den = (f & 2) ? nss : css;
  double r = __longlong_as_double(0x7ff8000000000000ll);
  if (!dead) r = div3_rn(div_rn(num, den));
  ccx = shfl_d(r, quad0);
  ncx = shfl_d(r, quad0 + 2);
  ccy = HEADING ? shfl_d(r, quad0 + 1) : 0.0;
  ncy = HEADING ? shfl_d(r, quad0 + 3) : 0.0;
}

__device__ __forceinline__ void write_row(fhp_foothold* dst, double x, double y, double z, int row, int col, int foot,
                                          int cycle, int kind) {
  double* d = reinterpret_cast<double*>(dst);
  d[0] = x;
  d[1] = y;
  d[2] = z;
  unsigned long long idx = static_cast<unsigned int>(row) | (static_cast<unsigned long long>(static_cast<unsigned int>(col)) << 32);
  unsigned long long ids = static_cast<unsigned long long>(foot & 0xff) | (static_cast<unsigned long long>(cycle & 0xff) << 8) |
                           (static_cast<unsigned long long>(kind & 0xff) << 16);
  reinterpret_cast<unsigned long long*>(dst)[3] = idx;
  reinterpret_cast<unsigned long long*>(dst)[4] = ids;
}

// Streamed response (fhp_plan_stream): the rows a warp produces in one gait cycle — 8 queries x 4 feet x 20 bytes
// {x, y double, z float} — go straight into the caller's mapped HOST buffer.  The quads whose cycle succeeded are packed
// to the front of the warp's 640-byte block (query order) through shared memory, so every store instruction writes 16
// contiguous bytes per lane from a 128-byte aligned base: PCIe carries full write packets while the kernel keeps
// planning, and failed cycles cost no bytes.  The per-query cycle masks tell the reader which quads are there.
// WARP-COLLECTIVE.
__device__ __forceinline__ void stream_rows(uint32_t stg, int lane, bool mine, double x, double y, float z, char* dstWarp) {
  const unsigned okl = __ballot_sync(kFull, mine);
  const unsigned okq = okl & 0x11111111u;  // one bit per quad (its foot 0)
  const int rank = __popc(okq & ((1u << (lane & ~3)) - 1u));
  if (mine) {
    const uint32_t a = stg + static_cast<uint32_t>(rank * 4 + (lane & 3)) * 20u;
    const unsigned long long xb = static_cast<unsigned long long>(__double_as_longlong(x));
    const unsigned long long yb = static_cast<unsigned long long>(__double_as_longlong(y));
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(static_cast<uint32_t>(xb)) : "memory");
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a + 4u), "r"(static_cast<uint32_t>(xb >> 32)) : "memory");
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a + 8u), "r"(static_cast<uint32_t>(yb)) : "memory");
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a + 12u), "r"(static_cast<uint32_t>(yb >> 32)) : "memory");
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a + 16u), "r"(__float_as_uint(z)) : "memory");
  }
  __syncwarp();
  const int nChunks = __popc(okq) * 5;  // 16-byte chunks: 80 bytes per successful quad
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int c = k * 32 + lane;
    if (c < nChunks) {
      const uint4 v = lds_v4(stg + static_cast<uint32_t>(c) * 16u);
      asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dstWarp + c * 16), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                   : "memory");
    }
  }
  __syncwarp();
}

// Block-wide copy of the SpiralIterator table into shared memory (ranks are looked up with a per-lane index in the
// cooperative scan; the constant bank would serialise those).  Returns the table; ends with __syncthreads.
template <class GV>
__device__ __forceinline__ const uint16_t* load_spiral_table(const GV& G, uint32_t* smem) {
  uint16_t* t = reinterpret_cast<uint16_t*>(smem);
  for (int k = threadIdx.x; k < G.nCand(); k += blockDim.x)
    t[k] = static_cast<uint16_t>(static_cast<uint8_t>(G.g.spiral[k][0])) |
           static_cast<uint16_t>(static_cast<uint16_t>(static_cast<uint8_t>(G.g.spiral[k][1])) << 8);
  __syncthreads();
  return t;
}

// Block barriers per gait cycle: 1 = top of the cycle, 2 = + after the staging, 3 = + after the search.  They keep the
// block's warps in the same code region; what that is worth depends on the size of the cycle's code: the 0.02 m
// specialisation (hot loop well inside the instruction cache) is fastest with one, everything else with three
// (profiles/r2_experiments.md, both tables).
#ifndef FHP_SYNC_LEVEL
#define FHP_SYNC_LEVEL 3
#endif
#ifndef FHP_SYNC_LEVEL_S1
#define FHP_SYNC_LEVEL_S1 1
#endif
#ifndef FHP_PLAN_MIN_BLOCKS
#define FHP_PLAN_MIN_BLOCKS 5
#endif
#ifndef FHP_PLAN_MIN_BLOCKS_MASK  // the mask path needs no box buffers: 7 blocks fit in shared memory (72 registers)
#define FHP_PLAN_MIN_BLOCKS_MASK 7
#endif
// The 0.01 m specialisation is compiled for FOUR blocks per SM = 124 registers (shared memory would hold five).  Its
// window is 23 columns: with 124 registers all 23 sixteen-byte column loads of a lane's window are in flight at once (92
// destination registers), with 72-96 the compiler issues them in register-limited groups and a map larger than the L2
// pays its DRAM latency once per group — config 5 (3 GB map): 13.7 ms per launch at 72 registers, 12.7 at 96, 7.7 at 124
// and at 164; the same geometry on an L2-resident map: 2.67 / 2.52 / 2.19 / 2.43 ms.  It is the registers, not the
// number of resident warps: 96 registers at 16 warps per SM (shared memory padded) runs like 96 at 20
// (profiles/r2_experiments.md).
#ifndef FHP_PLAN_MIN_BLOCKS_S2
#define FHP_PLAN_MIN_BLOCKS_S2 4
#endif
template <bool ONEPAR>
__device__ __forceinline__ const ParDev& par_of(const PlanArgs& A, int id) {
  if constexpr (ONEPAR) return A.par0; else return A.pars[id];
}
template <bool WANT_CEN, int STAGE, bool WANT_DEF, bool HEADING, bool STREAM, int SPEC, bool ONEPAR = false>
__global__ void __launch_bounds__(kPlanThreads, STAGE == kStageMask ? (SPEC == 2 ? FHP_PLAN_MIN_BLOCKS_S2 : FHP_PLAN_MIN_BLOCKS_MASK) : FHP_PLAN_MIN_BLOCKS) plan_kernel(const __grid_constant__ PlanArgs A,
                                                            const __grid_constant__ GeoDev Gd) {
  const GeoV<SPEC> G(Gd);  // SPEC > 0: the geometry's scalars are compile-time constants (fhp_device.cuh)
  constexpr int kSync = SPEC == 1 ? FHP_SYNC_LEVEL_S1 : FHP_SYNC_LEVEL;
  extern __shared__ __align__(128) uint32_t smem_raw[];  // (TMA destinations must be 128-byte aligned)
  uint32_t* const smem = smem_raw;
  constexpr bool USE_TMA = STAGE == kStageTma;
  const int lane = threadIdx.x & 31;
  const int warpInBlock = threadIdx.x >> 5;
  TmaWarp T;
  T.bar[0] = smem_u32(smem + G.mbarOff() + warpInBlock * 4);
  T.bar[1] = T.bar[0] + 8;
  T.boxAddr = smem_u32(smem + G.boxOff()) + static_cast<uint32_t>(warpInBlock * 2 * G.tmaB() * G.boxWords()) * 4u;
  T.parity[0] = T.parity[1] = 0;
  if (USE_TMA && lane == 0) {
    mbar_init(T.bar[0], 1);
    mbar_init(T.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint16_t* sSpiral = load_spiral_table(G, smem);  // ends with __syncthreads (also publishes the mbarriers)
  // shared-space addresses of the warp's plane block and of this lane's slot in it.  Opaque to the compiler: under
  // register pressure it otherwise recomputes them at every use — ten instructions each time (thread id, the shared
  // window base from the cluster CTA id, two multiply-adds) instead of one reload
  uint32_t sw = smem_u32(smem + G.spiralWords() + warpInBlock * 32 * G.stride());
  uint32_t mine = sw + static_cast<uint32_t>(lane * G.stride()) * 4u;
#ifndef FHP_NO_OPAQUE
  asm volatile("" : "+r"(sw));
#ifndef FHP_NO_OPAQUE_MINE
  asm volatile("" : "+r"(mine));
#endif
#endif
  const long long gwarp = static_cast<long long>(blockIdx.x) * (kPlanThreads / 32) + warpInBlock;
  // A warp plans 2^qpwShift queries (its first quads).  Eight keeps every lane busy; batches too small to fill the GPU
  // spread over more warps instead, because the warp-cooperative steps (staging passes, spiral scans) serve a warp's
  // lanes one after the other and the gait cycles are sequential: fewer queries per warp = shorter critical path.
  const int quad = lane >> 2;
  const long long q = (gwarp << A.qpwShift) + quad;
  const int f = lane & 3;           // RF,RH,LH,LF
  const int quad0 = lane & ~3;
  const bool active = quad < (1 << A.qpwShift) && q < A.nq;
  // (a whole-idle warp of the last block keeps running with active == false: the block synchronises every cycle)
  const MapDev& M = A.M;

  fhp_query qu;
  qu.x = qu.y = qu.z = 0.0;
  qu.param_id = 0;
  if (active) {
    const fhp_query* src = A.q + q;
    qu.x = src->x;
    qu.y = src->y;
    qu.z = src->z;
    qu.param_id = src->param_id;
  }
  unsigned err = 0;
  if (static_cast<unsigned>(qu.param_id) >= static_cast<unsigned>(A.nPars)) {  // validated here: large batches are not
    err = kErrParam;                                                           // walked on the host
    qu.param_id = 0;
  }
  // ONEPAR (calls with one parameter set, the service's case): the derived table entry travels in the kernel arguments, so
  // its fields are constant-bank operands instead of loads through a pointer held in registers
  const ParDev& P = par_of<ONEPAR>(A, qu.param_id);
  const long long R = 4 + 4 * static_cast<long long>(A.cycles);
  fhp_foothold* outN = A.outN + q * R;
  fhp_foothold* outC = WANT_CEN ? A.outC + q * R : nullptr;
  constexpr bool wantDef = WANT_DEF;  // "default" chain (FP.cpp:774-781, :1338-1348), a by-product of the same state
  fhp_foothold* outD = wantDef ? A.outD + q * R : nullptr;

  // initialize() stance (FP.cpp:350-378) and setFirstGait (FP.cpp:2679-2699) for both chains
  const double stx = P.stancex[f] + qu.x, sty = P.stancey[f] + qu.y, stz = 0.0 + qu.z;
  // (HEADING: half a step back along the heading; include/fhp.h fhp_params.heading)
  double nomx = stx - (HEADING ? P.halfStepX : P.halfStep), nomy = HEADING ? sty - P.halfStepY : sty;
  double cenx = nomx, ceny = nomy;
  double defx = nomx, defy = nomy;  // default chain: setFirstGait too (FP.cpp:562)
  // stream mode: this warp's staging block (behind everything else in shared memory) and its first query's slot
  // (stream mode always runs 8 queries per warp: the group of 8 queries is part of the buffer format)
  const uint32_t sStg = STREAM ? smem_u32(smem + G.smemWords()) + static_cast<uint32_t>(warpInBlock) * kStreamWarpBytes : 0u;
  char* const sBase = STREAM ? reinterpret_cast<char*>(A.sOut) + ((A.sQ0 >> 3) + gwarp) * 640 : nullptr;
  const long long sPitch = STREAM ? ((A.sNq + 7) >> 3) * 640 : 0;  // bytes of one cycle block
  uint32_t succMask = 0u, kindMask = 0u;
  if (STREAM) stream_rows(sStg, lane, active, stx, sty, static_cast<float>(stz), sBase);
  if (active && !STREAM) {  // FP.cpp:685-699 (and :713-727 for the centroid message, :666-671 for the default chain)
    int r0, c0;
    cell_of_position(M, stx, sty, r0, c0);
    write_row(outN + f, stx, sty, stz, r0, c0, f, 0, FHP_KIND_STANCE);
    if (WANT_CEN) write_row(outC + f, stx, sty, stz, r0, c0, f, 0, FHP_KIND_STANCE);
    if (wantDef) write_row(outD + f, stx, sty, stz, r0, c0, f, 0, FHP_KIND_STANCE);
  }
  // would the reference's opt chain abort this service call at cycle 0 (FP.cpp:931-934)?  Its state is
  // still the first-gait stance there, so the submap request is the nominal chain's next centre with isos_ extents.
  bool optFail = false;
  if (A.cycles > 0) {
    double ocx, ocy;
    polygon_center(shfl_d(nomx, quad0), shfl_d(nomy, quad0), shfl_d(nomx, quad0 + 1), shfl_d(nomy, quad0 + 1),
                   shfl_d(nomx, quad0 + 2), shfl_d(nomy, quad0 + 2), shfl_d(nomx, quad0 + 3), shfl_d(nomy, quad0 + 3),
                   ocx, ocy);
    double onx = ocx + P.stepLength, ony = qu.y + 0.0;
    if (HEADING) {
      const double lat0 = (qu.y * P.ch - mulz(qu.x, P.sh)) + 0.0;
      const double fwd = (ocx * P.ch + mulz(ocy, P.sh)) + P.stepLength;
      onx = fwd * P.ch - mulz(lat0, P.sh);
      ony = mulz(fwd, P.sh) + lat0 * P.ch;
    }
    if (active && f == 0) optFail = !submap_request_ok(M, onx, ony, P.isosLx, P.isosLy);
  }
  // A quad whose search centre has stopped being finite is DEAD for the rest of its plan: a degenerate stance polygon
  // (two feet of the centroid chain at the (0,0,0) the centroid method returns without a case) has a zero area, the
  // centroid is 0/0, every foothold test fails on an index of -2^30, the failed cycle leaves the chains as they are, and
  // the next cycle starts from the same polygon (~6 % of the lane-cycles of C3).  Such lanes — and the lanes of a warp
  // that have no query at all, whose all-zero state is the same polygon — keep taking part in the warp-collective steps
  // but skip the work: same statuses, same rows, none of the special-operand division paths.
  // (Not in the bug-compatible spiral mode on a shifted start index: there upstream wraps the centre's index of -2^30
  // into the buffer and the scan goes on from that cell — such a foothold can succeed.)
  const bool canDie = !(P.spiralMode == FHP_SPIRAL_FAITHFUL && !(M.start0 == 0 && M.start1 == 0));
  bool dead = !active;
  double adjY = 0.0;  // ajustedPose_[1] (FP.cpp:759)
  int prevW0 = -0x7fffffff, prevW1 = -0x7fffffff;  // window staged in the previous cycle (still in this lane's slot)
  int nrows = 4, succeed = 0;
  bool success = false;

  for (int g = 0; g < A.cycles; ++g) {
    if constexpr (kSync >= 1) __syncthreads();  // keep the block's warps in the same code region
    // getFootholdSearchGridMap for both chains (FP.cpp:788, :830 -> :2166-2250)
    double ccx, ccy, ncx, ncy;
#ifdef FHP_NO_QUAD_DIV
    polygon_center(shfl_d(cenx, quad0), shfl_d(ceny, quad0), shfl_d(cenx, quad0 + 1), shfl_d(ceny, quad0 + 1),
                   shfl_d(cenx, quad0 + 2), shfl_d(ceny, quad0 + 2), shfl_d(cenx, quad0 + 3), shfl_d(ceny, quad0 + 3),
                   ccx, ccy);
    polygon_center(shfl_d(nomx, quad0), shfl_d(nomy, quad0), shfl_d(nomx, quad0 + 1), shfl_d(nomy, quad0 + 1),
                   shfl_d(nomx, quad0 + 2), shfl_d(nomy, quad0 + 2), shfl_d(nomx, quad0 + 3), shfl_d(nomy, quad0 + 3),
                   ncx, ncy);
#else
    quad_centres<HEADING>(cenx, ceny, nomx, nomy, quad0, f, dead, ccx, ccy, ncx, ncy);
#endif
    const double nextY = qu.y + adjY;                                 // :2201
    // HEADING: next centre = ((c.u) + stepLength) u + ((p0.n) + adjY) n — forward coordinate from the stance centroid,
    // lateral coordinate pinned to the initial pose plus the drift (reduces bit-exactly to :2199-2201 at cos 1, sin 0)
    // (mulz drops a term whose sin factor is exactly zero: theta = 0 then equals the reference for non-finite centroids too)
    const double hLat = HEADING ? (qu.y * P.ch - mulz(qu.x, P.sh)) + adjY : 0.0;
    // (per-leg bias and thresholds are re-read from the parameter table each cycle: L1 hits, fewer live registers)
    const double bx = P.biasx[f], by = P.biasy[f];
    double dnx = 0.0, dny = 0.0, dnz = 0.0;                           // getDefaultFootholds (FP.cpp:2253-2304)
    if (wantDef) {
      double dcx, dcy;
      polygon_center(shfl_d(defx, quad0), shfl_d(defy, quad0), shfl_d(defx, quad0 + 1), shfl_d(defy, quad0 + 1),
                     shfl_d(defx, quad0 + 2), shfl_d(defy, quad0 + 2), shfl_d(defx, quad0 + 3), shfl_d(defy, quad0 + 3),
                     dcx, dcy);
      if (HEADING) {
        const double fwd = (dcx * P.ch + mulz(dcy, P.sh)) + P.stepLength;
        dnx = (fwd * P.ch - mulz(hLat, P.sh)) + bx;
        dny = (mulz(fwd, P.sh) + hLat * P.ch) + by;
      } else {
        dnx = (dcx + P.stepLength) + bx;
        dny = nextY + by;
      }
      if (active) dnz = static_cast<double>(mean_height(M, G, P, dnx, dny));
    }
    double px = (ccx + P.stepLength) + bx, py = nextY + by;           // centroid-chain next foothold (:2199,:2414)
    double rx = (ncx + P.stepLength) + bx, ry = nextY + by;           // nominal-chain next foothold = rectangle centre
    if (HEADING) {
      const double fwdC = (ccx * P.ch + mulz(ccy, P.sh)) + P.stepLength, fwdN = (ncx * P.ch + mulz(ncy, P.sh)) + P.stepLength;
      const double latS = mulz(hLat, P.sh), latC = hLat * P.ch;
      px = (fwdC * P.ch - latS) + bx;
      py = (mulz(fwdC, P.sh) + latC) + by;
      rx = (fwdN * P.ch - latS) + bx;
      ry = (mulz(fwdN, P.sh) + latC) + by;
    }

    dead = dead || (canDie && (!(fabs(px) <= 1.7976931348623157e308) || !(fabs(py) <= 1.7976931348623157e308)));
    const bool live = !dead;  // (implies active)
    // stage the window around the search centre (= centroid-chain point, FP.cpp:861-862)
    int uc0 = -1073741824, uc1 = -1073741824;  // raw_index_axis of a non-finite coordinate
    if (live) {
      uc0 = raw_index_axis(px, M.halfx, M.posx, M.res, M.rinv);
      uc1 = raw_index_axis(py, M.halfy, M.posy, M.res, M.rinv);
    }
    int w0 = uc0 - G.HW();
    const int w1 = uc1 - G.HW();
    if (USE_TMA) {  // TMA needs the box's innermost (row) coordinate 16-byte aligned: shift the origin down to a
      int c0 = w0 + M.start0;  // multiple of 4 buffer rows (rows % 4 == 0, so the wrap keeps the alignment)
      if (c0 >= M.rows) c0 -= M.rows;
      w0 -= c0 & 3;
    }
    DiscWin dw = {0, 0, -1, -1, 0ull, 0};  // foot disc of the centre: default test + height (FP.cpp:2048, :2529)
    if (live) {
      dw = disc_window(M, G, px, py, uc0, uc1, P.r, P.r2);
    }
    // (discs of up to four cells — footRadius <= resolution — are loaded here, in flight during the staging below;
    // larger ones are loaded in one batch where they are summed, mean_height)
    DiscElev de = {0.0f, 0.0f, 0.0f, 0.0f};
    const uint32_t elevSlot = smem_u32(smem + G.elevOff()) + static_cast<uint32_t>(warpInBlock * 512 + lane) * 4u;
    if (G.discK() == 1) {
      de = load_disc_elev(M, dw);
    } else {
      disc_elev_fetch(M, dw, elevSlot);  // cp.async into shared memory, waited for where the mean is taken
    }
    // A query that failed keeps its chains (FP.cpp:1571-1576) and only drifts 7 mm sideways per cycle, so its next
    // window is usually the SAME cells: the bit-planes are still in this lane's slot, nothing to stage.
    const bool restage = live && !(w0 == prevW0 && w1 == prevW1);
    prevW0 = w0;
    prevW1 = w1;
    if (STAGE == kStageMask)
      stage_windows_mask(M, G, A.mask, sw, mine, lane, restage, w0, w1, P.thrD, P.thrC);
    else if (USE_TMA)
      stage_windows_tma(M, G, &A.tmap, T, sw, lane, restage, w0, w1, P.thrD, P.thrC);
    else
      stage_windows(M, G, sw, lane, restage, w0, w1, P.thrD, P.thrC);

    if constexpr (kSync >= 2) __syncthreads();
    CentroidOut co;
    co.x = co.y = 0.0;
    co.ccase = -1;
    double nz = 0.0, cz = 0.0;
    bool okDef = false;
    float zCentre = 0.0f;
    if (live) {
      Planes pl(mine, STAGE == kStageMask ? 0 : G.pack(), G.rpShift(), G.WIN());
      zCentre = G.discK() == 1 ? mean_height_loaded(M, P, px, py, dw, de, true) : disc_elev_mean(M, P, px, py, dw, elevSlot);                // FP.cpp:2029 (used when the leg is valid)
      okDef = default_eval(M, G, P, pl, w0, w1, px, py, dw, err);        // FP.cpp:2012
      co = centroid_eval(M, G, P, pl, w0, w1, uc0, uc1, px, py, err);  // FP.cpp:818-821
    }
    if constexpr (kSync >= 4) __syncthreads();
    const NominalOut no =
        nominal_eval<false, STAGE == kStageMask, HEADING>(M, G, P, sSpiral, sw, mine, lane, live, uc0, uc1, w0, w1, px, py, rx, ry, okDef, err);  // :863-903
    const bool ok = no.ok;
    if constexpr (kSync >= 3) __syncthreads();
    if (active) {
      if (ok) nz = static_cast<double>(zCentre);  // at the CENTRE (FP.cpp:2029)
      if (WANT_CEN && co.ccase >= 0 && co.ccase <= 3) cz = static_cast<double>(mean_height(M, G, P, co.x, co.y));
    }
    // footholdValidation_ = RF & RH & LH & LF (FP.cpp:1323)
    const unsigned okb = __ballot_sync(kFull, ok);
    const bool all_ok = ((okb >> quad0) & 0xfu) == 0xfu;
    if (STREAM) {
      stream_rows(sStg, lane, active && all_ok, no.x, no.y, static_cast<float>(nz),
                  sBase + static_cast<long long>(g + 1) * sPitch);
      if (all_ok) {
        succMask |= 1u << (g & 31);
        if (no.kind == FHP_KIND_CANDIDATE) kindMask |= 1u << (g & 31);
      }
      if ((g & 31) == 31 || g + 1 == A.cycles) {  // per lane: {cycle succeeded, this foot came from the spiral} bits
        if (active)
          reinterpret_cast<uint2*>(A.sMask)[((A.sQ0 + q) * ((A.cycles + 31) >> 5) + (g >> 5)) * 4 + f] = make_uint2(succMask, kindMask);
        succMask = kindMask = 0u;
      }
    }
    if (active) {
      if (all_ok) {
        if (!STREAM) write_row(outN + nrows + f, no.x, no.y, nz, no.row, no.col, f, g, no.kind);  // :1378-1396
        if (WANT_CEN) {
          int r0, c0;
          cell_of_position(M, co.x, co.y, r0, c0);
          write_row(outC + nrows + f, co.x, co.y, cz, r0, c0, f, g, FHP_KIND_DEFAULT);  // :1444-1462
        }
        if (wantDef) {
          int r0, c0;
          cell_of_position(M, dnx, dny, r0, c0);
          write_row(outD + nrows + f, dnx, dny, dnz, r0, c0, f, g, FHP_KIND_DEFAULT);  // :1344-1348
          defx = dnx;  // :1338-1341
          defy = dny;
        }
        nrows += 4;
        succeed = g + 1;   // :1379
        success = true;    // :1380
        nomx = no.x;       // :1413-1416
        nomy = no.y;
        cenx = co.x;       // :1480-1483
        ceny = co.y;
      } else {
        success = false;   // :1574, no break (:1575)
      }
    }
    adjY += P.drift;       // :1578
  }
  if (active && err) atomicOr(A.err, err);
  unsigned e4 = err;  // OR over the quad
  e4 |= __shfl_xor_sync(kFull, e4, 1);
  e4 |= __shfl_xor_sync(kFull, e4, 2);
  if (active && f == 0) {
    fhp_query_status s;
    s.success = success ? 1 : 0;
    s.gait_cycles = static_cast<uint8_t>(A.cycles);
    s.gait_cycles_succeed = static_cast<uint8_t>(succeed);
    s.reserved = static_cast<uint8_t>((e4 & FHP_STATUS_ERR_MASK) | (optFail ? FHP_STATUS_OPT_SUBMAP_FAIL : 0));
    s.n_rows = nrows;
    A.st[q] = s;
  }
}

// ---- fine-grained primitives: one lane per task -------------------------------------------------------------------
// mode 0: checkFoothold, 1: checkFootholdUseCentroidMethod, 2: getFootholdMeanHeight, 3: checkDefaultFoothold,
//      4: checkCandidateFoothold (the spiral search alone), 5: checkCirclePolygonFoothold
template <int MODE>
__global__ void __launch_bounds__(kPlanThreads) task_kernel(const __grid_constant__ TaskArgs A,
                                                            const __grid_constant__ GeoDev Gd) {
  const GeoV<0> G(Gd);
  extern __shared__ __align__(16) uint32_t smem[];
  const int lane = threadIdx.x & 31;
  const int warpInBlock = threadIdx.x >> 5;
  const uint16_t* sSpiral = load_spiral_table(G, smem);
  const uint32_t sw = smem_u32(smem + G.spiralWords() + warpInBlock * 32 * G.stride());
  const long long i = (static_cast<long long>(blockIdx.x) * (kPlanThreads / 32) + warpInBlock) * 32 + lane;
  const bool active = i < A.n;
  const MapDev& M = A.M;
  const ParDev& P = A.pars[0];
  double px = 0, py = 0, rx = 0, ry = 0;
  if (active) {
    px = A.centres[2 * i];
    py = A.centres[2 * i + 1];
    if (MODE == 0 || MODE == 4 || MODE == 5) {
      rx = A.rects[2 * i];
      ry = A.rects[2 * i + 1];
    }
  }
  unsigned err = 0;
  if (MODE == 2) {
    if (active) A.z[i] = mean_height(M, G, P, px, py);
    return;
  }
  if (MODE == 5) {
    if (active) {
      const Box b = P.heading ? centre_box(rx, ry) : search_box(P, rx, ry);
      A.valid[i] = circle_polygon_direct(M, P, px, py, b) ? 1 : 0;
    }
    return;
  }
  const int uc0 = raw_index_axis(px, M.halfx, M.posx, M.res, M.rinv), uc1 = raw_index_axis(py, M.halfy, M.posy, M.res, M.rinv);
  const int w0 = uc0 - G.HW(), w1 = uc1 - G.HW();
  stage_windows(M, G, sw, lane, active, w0, w1, P.thrD, P.thrC);
  Planes pl(sw + static_cast<uint32_t>(lane * G.stride()) * 4u, G.pack(), G.rpShift(), G.WIN());
  if (MODE == 0 || MODE == 3 || MODE == 4) {
    DiscWin dw = {0, 0, -1, -1, 0ull, 0};
    if (active) dw = disc_window(M, G, px, py, uc0, uc1, P.r, P.r2);
    NominalOut no;
    const bool okDef = (MODE != 4) && active && default_eval(M, G, P, pl, w0, w1, px, py, dw, err);
    if (MODE == 3)
      no = nominal_eval<true, false, false>(M, G, P, sSpiral, sw, sw + static_cast<uint32_t>(lane * G.stride()) * 4u, lane, active, uc0, uc1, w0, w1, px, py, rx, ry, okDef, err);
    else if (P.heading)  // uniform: one parameter set per task launch
      no = nominal_eval<false, false, true>(M, G, P, sSpiral, sw, sw + static_cast<uint32_t>(lane * G.stride()) * 4u, lane, active, uc0, uc1, w0, w1, px, py, rx, ry, okDef, err);
    else
      no = nominal_eval<false, false, false>(M, G, P, sSpiral, sw, sw + static_cast<uint32_t>(lane * G.stride()) * 4u, lane, active, uc0, uc1, w0, w1, px, py, rx, ry, okDef, err);
    if (active) {
      if (MODE == 0) {
        double z = 0.0;
        if (no.ok) z = static_cast<double>(mean_height(M, P, px, py, dw));
        write_row(A.out + i, no.x, no.y, z, no.row, no.col, 0, 0, no.kind);
      }
      if (MODE == 4) {  // checkCandidateFoothold sets only the xy of a hit (FP.cpp:2105-2107)
        if (no.ok)
          write_row(A.out + i, no.x, no.y, 0.0, no.row, no.col, 0, 0, FHP_KIND_CANDIDATE);
        else
          write_row(A.out + i, 0.0, 0.0, 0.0, -1, -1, 0, 0, FHP_KIND_CANDIDATE);
      }
      A.valid[i] = no.ok ? 1 : 0;
    }
  }
  if (!active) return;
  if (MODE == 1) {
    CentroidOut co = centroid_eval(M, G, P, pl, w0, w1, uc0, uc1, px, py, err);
    double z = 0.0;
    if (co.ccase >= 0 && co.ccase <= 3) z = static_cast<double>(mean_height(M, G, P, co.x, co.y));
    int r0, c0;
    cell_of_position(M, co.x, co.y, r0, c0);
    write_row(A.out + i, co.x, co.y, z, r0, c0, 0, 0, FHP_KIND_DEFAULT);
    A.ccase[i] = static_cast<int8_t>(co.ccase);
  }
  if (err) atomicOr(A.err, err);
}

// ---- mask planes: one pass over the traversability layer per (map update, threshold pair) ------------------------
// A warp owns 32 consecutive buffer columns of one row word k: 32 coalesced 128-byte column reads, two ballots each
// (the 2-bit cell class, see stage_windows_mask), lane i keeps the words of column i; the words go to entry (k, c)
// as the low halves and to entry (k-1, c) as the high halves.  Rows beyond the map read as NaN (class 0); the high
// halves of the last row word are zeroed by the warp that owns it.  *flags |= 1 when a -inf cell is seen.
__global__ void __launch_bounds__(256) mask_build_kernel(const float* __restrict__ trav, int rows, int cols, int rowWords,
                                                         float thrD, float thrC, uint4* __restrict__ mask,
                                                         unsigned* __restrict__ flags) {
  const int lane = threadIdx.x & 31;
  const long long warp = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int colTiles = (cols + 31) >> 5;
  const int k = static_cast<int>(warp / colTiles), ct = static_cast<int>(warp % colTiles);
  if (k >= rowWords) return;
  const int row = k * 32 + lane;
  const float inf = __int_as_float(0x7f800000);
  uint32_t m0 = 0u, m1 = 0u;
  bool negInf = false;
  for (int i0 = 0; i0 < 32; i0 += 8) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int col = ct * 32 + i0 + i;
      v[i] = __int_as_float(0x7fc00000);
      if (col < cols && row < rows) v[i] = __ldg(trav + static_cast<size_t>(col) * rows + row);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool fin = fabsf(v[i]) < inf;
      const bool bD = fin && v[i] < thrD, bC = fin && v[i] < thrC;
      negInf |= (v[i] == -inf);
      const uint32_t e0 = __ballot_sync(kFull, fin && (bC || !bD));  // classes 1 and 3
      const uint32_t e1 = __ballot_sync(kFull, bD);                  // classes 2 and 3
      if (lane == i0 + i) {
        m0 = e0;
        m1 = e1;
      }
    }
  }
  if (negInf) atomicOr(flags, 1u);
  const int col = ct * 32 + lane;
  if (col < cols) {
    uint32_t* e = reinterpret_cast<uint32_t*>(mask + static_cast<size_t>(k) * cols + col);
    e[0] = m0;
    e[2] = m1;
    if (k + 1 == rowWords) {
      e[1] = 0u;
      e[3] = 0u;
    }
    if (k > 0) {
      uint32_t* p = reinterpret_cast<uint32_t*>(mask + static_cast<size_t>(k - 1) * cols + col);
      p[1] = m0;
      p[3] = m1;
    }
  }
}

cudaError_t launch_mask_build(const MapDev& M, float thrD, float thrC, uint4* mask, int rowWords, unsigned* flags,
                              cudaStream_t stream) {
  const long long warps = static_cast<long long>(rowWords) * ((M.cols + 31) / 32);
  const long long blocks = (warps + 7) / 8;
  mask_build_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(M.trav, M.rows, M.cols, rowWords, thrD, thrC, mask,
                                                                       flags);
  return cudaGetLastError();
}

// ---- compact response: the rows each query actually produced, back to back --------------------------------------
// offs[q] = sum of n_rows of the queries before q (exclusive scan; offs[n] = total).  One block; a wave-sized chunk
// is a few tens of thousands of queries.
__global__ void __launch_bounds__(1024) scan_rows_kernel(const fhp_query_status* __restrict__ st, long long n,
                                                         long long* __restrict__ offs, long long* hostTotal) {
  __shared__ long long warpSum[32];
  __shared__ long long carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (long long base = 0; base < n; base += 1024) {
    const long long i = base + threadIdx.x;
    const long long v = i < n ? st[i].n_rows : 0;
    long long x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const long long y = __shfl_up_sync(kFull, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) warpSum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      long long w = warpSum[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const long long y = __shfl_up_sync(kFull, w, d);
        if (lane >= d) w += y;
      }
      warpSum[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const long long before = carry + (warp ? warpSum[warp - 1] : 0) + (x - v);
    if (i < n) offs[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    offs[n] = carry;
    if (hostTotal) {  // mapped host word: the host sizes the chunk's row copy from it without waiting for an event
      *reinterpret_cast<volatile long long*>(hostTotal) = carry + 1;  // + 1: zero means "not yet"
      __threadfence_system();
    }
  }
}

// One warp per query: its n_rows rows move from the fixed-pitch plan output to out + offs[q] — as they are (FMT 40: five
// 8-byte words each), PACKED to 32 bytes (FMT 32: x, y, z, then row/col as uint16 and the three id bytes in one word) or
// as the 24-byte message row (FMT 24: x, y, then z as float32 — it is one by construction, FP.cpp:2553 — and the id bytes).
template <int FMT>
__global__ void __launch_bounds__(256) compact_rows_kernel(const fhp_foothold* __restrict__ rows,
                                                           const long long* __restrict__ offs, long long n, int R,
                                                           void* __restrict__ out) {
  const long long q = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= n) return;
  const int lane = threadIdx.x & 31;
  const long long o = offs[q];
  const int nr = static_cast<int>(offs[q + 1] - o);
  const unsigned long long* src = reinterpret_cast<const unsigned long long*>(rows + q * R);
  if (FMT == 40) {
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(static_cast<fhp_foothold*>(out) + o);
    for (int w = lane; w < nr * 5; w += 32) dst[w] = src[w];
  } else if (FMT == 32) {
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(static_cast<fhp_foothold_packed*>(out) + o);
    for (int w = lane; w < nr * 4; w += 32) {
      const int r = w >> 2, k = w & 3;
      unsigned long long v;
      if (k < 3) {
        v = src[r * 5 + k];
      } else {
        const unsigned long long idx = src[r * 5 + 3], ids = src[r * 5 + 4];  // row | col << 32 ; foot | cycle << 8 | kind << 16
        v = (idx & 0xffffull) | ((idx >> 32 & 0xffffull) << 16) | ((ids & 0xffffffull) << 32);
      }
      dst[w] = v;
    }
  } else {
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(static_cast<fhp_foothold24*>(out) + o);
    for (int w = lane; w < nr * 3; w += 32) {
      const int r = w / 3, k = w - 3 * r;
      unsigned long long v;
      if (k < 2) {
        v = src[r * 5 + k];
      } else {
        const float z = static_cast<float>(__longlong_as_double(static_cast<long long>(src[r * 5 + 2])));
        v = static_cast<unsigned long long>(__float_as_uint(z)) | ((src[r * 5 + 4] & 0xffffffull) << 32);
      }
      dst[w] = v;
    }
  }
}

cudaError_t launch_compact(const fhp_foothold* rows, const fhp_query_status* st, long long n, int R, long long* offs,
                           void* out, int fmt, long long* hostTotal, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  scan_rows_kernel<<<1, 1024, 0, stream>>>(st, n, offs, hostTotal);
  const unsigned nb = static_cast<unsigned>((n + 7) / 8);
  if (fmt == 32)
    compact_rows_kernel<32><<<nb, 256, 0, stream>>>(rows, offs, n, R, out);
  else if (fmt == 24)
    compact_rows_kernel<24><<<nb, 256, 0, stream>>>(rows, offs, n, R, out);
  else
    compact_rows_kernel<40><<<nb, 256, 0, stream>>>(rows, offs, n, R, out);
  return cudaGetLastError();
}

typedef void (*plan_kernel_fn)(const PlanArgs, const GeoDev);
// the default chain is only compiled into the centroid-output variants (fhp_plan_ex fills all three chains)
template <int STAGE, bool HEADING>
static plan_kernel_fn pick_plan_kernel_s(bool wantCen, bool wantDef, bool stream, int spec, bool onePar) {
  if (STAGE == kStageMask && !wantCen && !wantDef) {  // the response path: geometry-specialised variants (GeoK)
    if (!HEADING && onePar) {  // ... with the single parameter set in the kernel arguments
      if (spec == 1) return stream ? plan_kernel<false, kStageMask, false, false, true, 1, true> : plan_kernel<false, kStageMask, false, false, false, 1, true>;
      if (spec == 2) return stream ? plan_kernel<false, kStageMask, false, false, true, 2, true> : plan_kernel<false, kStageMask, false, false, false, 2, true>;
    }
    if (spec == 1) return stream ? plan_kernel<false, kStageMask, false, HEADING, true, 1> : plan_kernel<false, kStageMask, false, HEADING, false, 1>;
    if (spec == 2) return stream ? plan_kernel<false, kStageMask, false, HEADING, true, 2> : plan_kernel<false, kStageMask, false, HEADING, false, 2>;
  }
  if (stream) return plan_kernel<false, STAGE, false, HEADING, true, 0>;  // nominal rows only
  if (wantDef) return plan_kernel<true, STAGE, true, HEADING, false, 0>;
  return wantCen ? plan_kernel<true, STAGE, false, HEADING, false, 0> : plan_kernel<false, STAGE, false, HEADING, false, 0>;
}
// heading: some parameter set of the call carries a non-zero heading -> the variant with the rotated expressions
// (the reference-shaped variant stays exactly what it was)
static plan_kernel_fn pick_plan_kernel(int stage, bool wantCen, bool wantDef, bool heading, bool stream, int spec, bool onePar) {
#ifdef FHP_BENCH_BUILD  // scratch A/B builds: only the headline variants are instantiated (seconds instead of a minute)
  if (onePar && spec == 2) return plan_kernel<false, kStageMask, false, false, false, 2, true>;
  if (onePar && spec == 1) return plan_kernel<false, kStageMask, false, false, false, 1, true>;
  if (spec == 2) return plan_kernel<false, kStageMask, false, false, false, 2>;
  return spec == 1 ? plan_kernel<false, kStageMask, false, false, false, 1> : plan_kernel<false, kStageMask, false, false, false, 0>;
#else
  if (heading) {
    if (stage == kStageMask) return pick_plan_kernel_s<kStageMask, true>(wantCen, wantDef, stream, spec, onePar);
    if (stage == kStageTma) return pick_plan_kernel_s<kStageTma, true>(wantCen, wantDef, stream, spec, onePar);
    return pick_plan_kernel_s<kStageLdg, true>(wantCen, wantDef, stream, spec, onePar);
  }
  if (stage == kStageMask) return pick_plan_kernel_s<kStageMask, false>(wantCen, wantDef, stream, spec, onePar);
  if (stage == kStageTma) return pick_plan_kernel_s<kStageTma, false>(wantCen, wantDef, stream, spec, onePar);
  return pick_plan_kernel_s<kStageLdg, false>(wantCen, wantDef, stream, spec, onePar);
#endif
}

// The geometry-specialised kernel variant (GeoK) whose constants equal the tables the host built, or 0.
int geo_spec_of(const GeoDev& G) {
  if (G.stageMode != kStageMask || G.pack != 1 || G.rpShift != 5 || G.nIter != G.WIN || G.stride != 3 * G.WIN + 1 || G.winRows != G.WIN)
    return 0;
  auto same = [&](auto k) {
    return G.WIN == k.WIN && G.HW == k.HW && G.nSten == k.nSten && G.discK == k.discK && G.nCand == k.nCand &&
           G.lastTwoStart == k.lastTwoStart && G.nRings == k.nRings && G.ra == k.ra;
  };
  if (same(GeoK<1>{})) return 1;
  if (same(GeoK<2>{})) return 2;
  return 0;
}

// ---- launchers (called from fhp_capi.cu) --------------------------------------------------------------------------
cudaError_t launch_plan(const PlanArgs& A, const GeoDev& G, bool wantCen, bool heading, bool noSpec, bool noOnePar,
                        cudaStream_t stream) {
  if (A.nq <= 0) return cudaSuccess;
  const int warps = kPlanThreads / 32;
  const int qpw = 1 << A.qpwShift;
  const long long nWarps = (A.nq + qpw - 1) / qpw;
  const long long blocks = (nWarps + warps - 1) / warps;
  const bool streamMode = A.sOut != nullptr;
#ifndef FHP_SMEM_PAD  // (A/B knob: extra dynamic shared memory per block, to lower the occupancy without touching registers)
#define FHP_SMEM_PAD 0
#endif
  const size_t shm = static_cast<size_t>(G.smemWords) * sizeof(uint32_t) + 128 + FHP_SMEM_PAD +  // + alignment slack
                     (streamMode ? static_cast<size_t>(warps) * kStreamWarpBytes : 0);
  void (*kern)(const PlanArgs, const GeoDev) = pick_plan_kernel(G.stageMode, wantCen, A.outD != nullptr, heading, streamMode, noSpec ? 0 : geo_spec_of(G), A.nPars == 1 && !noOnePar);
  if (shm > 48 * 1024) {  // res 0.01-class windows: opt in to more than the default 48 KB of dynamic shared memory
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(shm));
    if (e != cudaSuccess) return e;
  }
  kern<<<static_cast<unsigned>(blocks), kPlanThreads, shm, stream>>>(A, G);
  return cudaGetLastError();
}

// Queries one full wave of plan_kernel blocks holds on the current device (occupancy x SM count x 32 queries per block):
// the chunk size of the host-buffer pipeline (a block runs all gait cycles sequentially, so a launch smaller than
// a wave takes as long as a full one).
long long plan_wave_queries(const GeoDev& G, bool wantCen) {
  void (*kern)(const PlanArgs, const GeoDev) = pick_plan_kernel(G.stageMode, wantCen, false, false, false, geo_spec_of(G), false);
  const size_t shm = static_cast<size_t>(G.smemWords) * sizeof(uint32_t) + 128;
  if (shm > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(shm));
  int perSm = 0, dev = 0, sms = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kern, kPlanThreads, shm) != cudaSuccess || perSm < 1) perSm = 1;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = 148;
  return static_cast<long long>(perSm) * sms * (kPlanThreads / 4);
}

cudaError_t launch_task(int mode, const TaskArgs& A, const GeoDev& G, cudaStream_t stream) {
  if (A.n <= 0) return cudaSuccess;
  const int warps = kPlanThreads / 32;
  const long long nWarps = (A.n + 31) / 32;
  const long long blocks = (nWarps + warps - 1) / warps;
  const size_t shm = (static_cast<size_t>(G.spiralWords) + static_cast<size_t>(warps) * 32 * G.stride) * sizeof(uint32_t);
  const unsigned nb = static_cast<unsigned>(blocks);
  if (shm > 48 * 1024) {
    cudaError_t e;
    switch (mode) {
      case 0: e = cudaFuncSetAttribute(task_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(shm)); break;
      case 1: e = cudaFuncSetAttribute(task_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(shm)); break;
      case 2: e = cudaFuncSetAttribute(task_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(shm)); break;
      case 3: e = cudaFuncSetAttribute(task_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(shm)); break;
      case 4: e = cudaFuncSetAttribute(task_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(shm)); break;
      default: e = cudaFuncSetAttribute(task_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(shm)); break;
    }
    if (e != cudaSuccess) return e;
  }
  switch (mode) {
    case 0: task_kernel<0><<<nb, kPlanThreads, shm, stream>>>(A, G); break;
    case 1: task_kernel<1><<<nb, kPlanThreads, shm, stream>>>(A, G); break;
    case 2: task_kernel<2><<<nb, kPlanThreads, shm, stream>>>(A, G); break;
    case 3: task_kernel<3><<<nb, kPlanThreads, shm, stream>>>(A, G); break;
    case 4: task_kernel<4><<<nb, kPlanThreads, shm, stream>>>(A, G); break;
    default: task_kernel<5><<<nb, kPlanThreads, shm, stream>>>(A, G); break;
  }
  return cudaGetLastError();
}

}  // namespace fhp
