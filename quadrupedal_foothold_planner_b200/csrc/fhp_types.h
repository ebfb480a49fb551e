// Plain-data structs shared by the kernels (fhp_device.cuh / fhp_kernels.cu) and the host C-ABI (fhp_capi.cu).
#pragma once
#include <stdint.h>

#include "../../include/fhp.h"

namespace fhp {

constexpr int kMaxCand = 1024;  // spiral table entries (nRings <= 15)
constexpr int kMaxSten = 96;    // foot-disc stencil entries
constexpr unsigned kFull = 0xffffffffu;
// window staging paths of plan_kernel
constexpr int kStageLdg = 0;   // pipelined coalesced loads of the float layer + ballots
constexpr int kStageTma = 1;   // one 2-D TMA box per window + ballots
constexpr int kStageMask = 2;  // funnel-shifted words of the precomputed D/C/F mask planes (no per-cell work)

// error bits reported per query (status.reserved) — any non-zero bit fails the whole call loudly
constexpr unsigned kErrWindow = 1u;   // a cell needed by the reference lies outside the staged window
constexpr unsigned kErrParam = 2u;    // a query's param_id is outside the parameter table (the query planned with set 0)

struct MapDev {
  const float* trav;
  const float* elev;
  int rows, cols, start0, start1;
  int fastOk;  // rows*cols < 2^31: 32-bit linear cell offsets are safe in the pipelined staging path
  int pad0;
  double res;
  double rinv;          // 1/res, correctly rounded (fast path of the position -> index conversion)
  double lenx, leny, posx, posy;
  double halfx, halfy;  // getVectorToOrigin: 0.5*length
  double fcx, fcy;      // mapPosition + (0.5*length - 0.5*res): the constant term of getPositionFromIndex
};

struct GeoDev {  // depends on (foot_radius, search_radius, resolution) only
  int nRings;    // ceil(searchRadius / res)                 SpiralIterator ctor
  int ra;        // foot-disc stencil half width (cells)
  int discK;     // foot disc of a CONTINUOUS centre: member cells lie within discK indices of the centre's cell (-1: generic walk)
  int HW;        // window half width = nRings + ra
  int WIN;       // 2*HW + 1 (<= 32)
  int pack;      // window columns per warp load: 1 (WIN<=32), 2 (WIN<=16), 4 (WIN<=8)
  int rpShift;   // log2(32/pack): lanes per column group
  int nIter;     // ceil(WIN / pack) = words per bit-plane
  int stride;    // shared-memory words per query (4 per load iteration: D, C, F, pad)
  int stageMode; // kStageLdg / kStageTma / kStageMask: how plan_kernel fills the planes
  int winRows;   // window rows a consumer may read: WIN, plus the TMA row-alignment shift (<= 3) in TMA mode
  int spiralWords;  // 32-bit words reserved at the start of shared memory for the spiral table (16-byte multiple)
  // TMA staging (plan_kernel): each window is fetched as one 2-D box of boxWords floats (32 rows*cols per load
  // iteration) into shared memory; tmaB windows per round, two rounds in flight per warp.
  int useTma;       // 1: the tensor map in the kernel arguments is valid for this map/geometry
  int tmaB;         // windows per TMA round (power of two)
  int tmaBShift;    // log2(tmaB)
  int boxWords;     // 32 * nIter
  int mbarOff;      // word offset of the per-warp mbarrier pairs (from the 128-byte aligned shared base)
  int boxOff;       // word offset of the per-warp box buffers (multiple of 32 words)
  int smemWords;    // total dynamic shared memory words for a kPlanThreads block (without alignment slack)
  int elevOff;      // word offset of the per-warp elevation staging (16 words per lane, plan_kernel), -1: none (discK == 1)
  int nCand;     // spiral table length
  int lastTwoStart;  // first rank that belongs to ring nRings-1 (distance-filtered rings)
  int nSten;
  uint32_t candRows[32];  // per window column: bit (d0+HW) set for every spiral-table offset (d0, column-HW)
  int8_t sten[kMaxSten][2];
  int8_t spiral[kMaxCand][2];
};

struct ParDev {  // one per fhp_params, derived on the host exactly as FootholdPlanner::initialize (FP.cpp:317-421)
  double biasx[4], biasy[4];      // RF,RH,LH,LF default bias
  double stancex[4], stancey[4];  // +-lengthBase*0.5, +-widthBase*0.5 (initial stance before adding the pose)
  double stepLength;              // (double)stepLength_
  double halfStep;                // (double)(stepLength_/2)  (float division, FP.cpp:2693)
  double R;                       // (double)searchRadius_
  double halfR;                   // 0.5*R
  double rectLx, rectLy;          // (double)(searchRadius_*2), (double)searchRadius_  (FP.cpp:1616-1617)
  double r, r2, R2;               // (double)footRadius_, pow(r,2), R*R
  double h, drift;
  float thrD, thrC;
  int spiralMode;
  int heading;                    // 1: fhp_params.heading != 0 (the rotated expressions below are in force)
  // heading extension (include/fhp.h); stance and bias above are already rotated.  With heading == 0
  // these hold cos = 1, sin = 0 and the reference's own offsets, and every rotated expression reduces bit-exactly.
  double ch, sh;                  // cos / sin of the heading
  double halfStepX, halfStepY;    // halfStep * ch, halfStep * sh (setFirstGait steps half a step back along u)
  double rectVx[4], rectVy[4];    // rot() of getSearchPolygon's vertex offsets (+R,+R/2) (+R,-R/2) (-R,-R/2) (-R,+R/2)
  double rectSlope[4];            // edge slopes dx/dy of the rotated rectangle (estimates only: poly_col_rows)
  double isosLx, isosLy;          // isos_.length, isos_.width (FP.cpp:384-394): the opt chain's gait-cycle submap request
};

}  // namespace fhp
