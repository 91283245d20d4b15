// Interface of the tcgen05 (sm_100a tensor-core) coarse scoring pass; implemented in mol_coarse_sm100.cu.
#pragma once
#include "common.cuh"

namespace mol {

struct CoarseWs {
  __nv_bfloat16* w1_bf16;  // qi-MLP layer-1 weights in the UMMA shared-memory image (see mol_coarse_sm100.cu)
  __nv_bfloat16* w2_bf16;  // qi-MLP layer-2 weights, same
  float* b1h;              // 0.5 * b1   (H)
  float* b2h;              // 0.5 * b2   (L), permuted to the kernel's logit order
  __nv_bfloat16* q_bf16;   // (chunk, ...) query sub-embeddings / tau in the UMMA image
  float* gqh;              // (chunk, L) 0.5 * gq, permuted
};

bool coarse_supported(const mol_shape_t& s);
void coarse_plan(const mol_shape_t& s, int chunk, Arena& a, CoarseWs* ws);
// weight images (once per search call)
int coarse_prepare(const mol_shape_t& s, const mol_weights_t& w, const CoarseWs& ws, cudaStream_t st);
// scores[b, x] ~= MoL score (bf16 operands / fp32 accumulation) for b < bc, x < N; row stride N
int coarse_scores(const mol_shape_t& s, const mol_index_t& ix, const CoarseWs& ws, const float* qsub,
                  const float* gq, int bc, float* scores, cudaStream_t st);
// gi_bf16 of the index in the kernel's logit order l' = m*P_Q + n (called by the index build)
int coarse_gi_image(const mol_shape_t& s, const float* gi_f32, uint16_t* gi_bf16, int64_t n, cudaStream_t st);
// flags[b] = 1 when the coarse candidate set cannot be shown to contain the exact top-k:
//   cand_scores[b, kk-1] + 1.5 * max_j |cand_scores[b,j] - exact_scores[b,j]| + 1e-3 >= topk_scores[b, k-1]
int coarse_safety_flags(const float* cand_scores, const float* exact_scores, const float* topk_scores,
                        int bc, int kk, int k, int32_t* flags, cudaStream_t st);

}  // namespace mol
