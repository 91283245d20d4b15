// tcgen05 / TMEM / TMA coarse scoring pass for sm_100a  (stage-2).
#include <math_constants.h>

#include "mol_coarse.cuh"

namespace mol {

bool coarse_supported(const mol_shape_t& s) {
  (void)s;
  return false;
}

void coarse_plan(const mol_shape_t& s, int chunk, Arena& a, CoarseWs* ws) {
  (void)s;
  (void)chunk;
  (void)a;
  (void)ws;
}

int coarse_prepare(const mol_shape_t& s, const mol_weights_t& w, const CoarseWs& ws, cudaStream_t st) {
  (void)s; (void)w; (void)ws; (void)st;
  MOL_CHECK_ARG(false, "tensor-core path not available");
}

int coarse_scores(const mol_shape_t& s, const mol_index_t& ix, const CoarseWs& ws, const float* qsub,
                  const float* gq, int bc, float* scores, cudaStream_t st) {
  (void)s; (void)ix; (void)ws; (void)qsub; (void)gq; (void)bc; (void)scores; (void)st;
  MOL_CHECK_ARG(false, "tensor-core path not available");
}

// One warp per query.
__global__ void safety_flags_kernel(const float* __restrict__ cand, const float* __restrict__ exact,
                                    const float* __restrict__ topk, int bc, int kk, int k,
                                    int32_t* __restrict__ flags) {
  int b = (blockIdx.x * blockDim.x + threadIdx.x) / 32;
  int lane = threadIdx.x % 32;
  if (b >= bc) return;
  float err = 0.f, cmin = CUDART_INF_F;
  for (int j = lane; j < kk; j += 32) {
    float c = cand[(int64_t)b * kk + j], e = exact[(int64_t)b * kk + j];
    if (c > -CUDART_INF_F && e > -CUDART_INF_F) {
      err = fmaxf(err, fabsf(c - e));
      cmin = fminf(cmin, c);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    err = fmaxf(err, __shfl_xor_sync(0xffffffffu, err, o));
    cmin = fminf(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
  }
  if (lane == 0) {
    float sk = topk[(int64_t)b * k + (k - 1)];
    flags[b] = (cmin + 1.5f * err + 1e-3f >= sk) ? 1 : 0;
  }
}

int coarse_safety_flags(const float* cand_scores, const float* exact_scores, const float* topk_scores,
                        int bc, int kk, int k, int32_t* flags, cudaStream_t st) {
  if (bc == 0) return MOL_OK;
  int threads = bc * 32;
  safety_flags_kernel<<<(threads + 255) / 256, 256, 0, st>>>(cand_scores, exact_scores, topk_scores,
                                                             bc, kk, k, flags);
  MOL_LAUNCH_CHECK();
  return MOL_OK;
}

}  // namespace mol
