// Host <-> kernel argument structs and launchers shared by fhp_kernels.cu and fhp_capi.cu.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "fhp_types.h"

namespace fhp {

#ifndef FHP_PLAN_THREADS
#define FHP_PLAN_THREADS 128
#endif
constexpr int kPlanThreads = FHP_PLAN_THREADS;

struct PlanArgs {
  alignas(64) CUtensorMap tmap;  // traversability layer, box = one staged window (valid iff G.useTma)
  MapDev M;
  const uint4* mask;  // cell-class planes of the traversability layer (valid iff G.stageMode == kStageMask)
  const ParDev* pars;
  int nPars;
  ParDev par0;  // pars[0] by value: calls with ONE parameter set read it from the constant bank (plan_kernel<..., ONEPAR>)
  const fhp_query* q;
  long long nq;
  int cycles;
  int qpwShift;  // log2 of the queries one warp plans (8 = every quad busy; fewer for batches that cannot fill the GPU)
  fhp_foothold* outN;
  fhp_foothold* outC;
  fhp_foothold* outD;  // optional: "default" chain rows
  fhp_query_status* st;
  unsigned* err;  // device word, OR of the kErr* bits of every query
  // streamed response (fhp_plan_stream): rows go straight into the caller's mapped host buffer, cycle-major blocks
  fhp_point20* sOut;     // block (g + 1, q / 8): 32 row slots, successful queries packed to the front
  uint32_t* sMask;       // per (query, 32-cycle word, foot): {success bits, kind bits} (fhp_stream_mask)
  long long sNq;         // queries of the whole call (the slot pitch of a cycle block)
  long long sQ0;         // index of this launch's first query within the call
};

struct TaskArgs {
  MapDev M;
  const ParDev* pars;
  const double* centres;
  const double* rects;
  long long n;
  fhp_foothold* out;
  uint8_t* valid;
  int8_t* ccase;
  float* z;
  unsigned* err;
};

cudaError_t launch_plan(const PlanArgs& A, const GeoDev& G, bool wantCen, bool heading, bool noSpec, bool noOnePar,
                        cudaStream_t stream);
int geo_spec_of(const GeoDev& G);
constexpr int kStreamWarpBytes = 32 * 20;  // shared-memory staging of one warp's rows in stream mode
long long plan_wave_queries(const GeoDev& G, bool wantCen);
cudaError_t launch_mask_build(const MapDev& M, float thrD, float thrC, uint4* mask, int rowWords, unsigned* flags,
                              cudaStream_t stream);
cudaError_t launch_compact(const fhp_foothold* rows, const fhp_query_status* st, long long n, int R, long long* offs,
                           void* out, int fmt, long long* hostTotal, cudaStream_t stream);
cudaError_t launch_task(int mode, const TaskArgs& A, const GeoDev& G, cudaStream_t stream);

}  // namespace fhp
