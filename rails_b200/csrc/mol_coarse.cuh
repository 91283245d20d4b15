// Interface of the tcgen05 (sm_100a tensor-core) coarse scoring pass; implemented in mol_coarse_sm100.cu.
#pragma once
#include "common.cuh"

namespace mol {

struct CoarseWs {
  uint8_t* w1_img;    // qi-MLP layer-1 weights (+ bias column) in the UMMA shared-memory image, fp16
  uint8_t* w2_img;    // qi-MLP layer-2 weights (+ bias column), same
  uint8_t* q_rec;     // (chunk, record) per query: block-diagonal image of Q_sub / tau | 0.5 * gq, fp16
  int32_t* overflow;  // device flag: a weight / query operand did not fit fp16 -> every query falls back to exact
};

// Where the coarse pass puts its results (any subset), and which item tiles it scores.
struct CoarseOut {
  float* scores;       // (bc, ld): column = (logical tile - tile_begin) * 128 + row; or nullptr
  int64_t ld;          // row stride of `scores`
  int tile_begin;      // LOGICAL item tiles (128 items) [tile_begin, tile_end) are scored; tile_end < 0 = every tile
  int tile_end;
  const int32_t* tile_map;  // device table: physical item tile of logical tile i (coarse_tile_maps), or nullptr = identity
  const float* thr;    // per-query thresholds of the fused candidate filter: thr[b * thr_stride]; or nullptr
  int thr_stride;
  int32_t* cand_cnt;   // (bc) counters (zeroed by the caller)
  float* cand_scores;  // (bc, cand_cap)
  int32_t* cand_idx;   // (bc, cand_cap)
  int cand_cap;
};

// compile-time tuning knobs of this build, e.g. "e2poly=0 e2h2=0x3E ..." (reported through mol_version();
// tests/test_gpu_parity.py configures the CPU numerics model from it when a tuning variant is loaded)
const char* coarse_build_knobs();
void* coarse_trace_buffer();  // debug (MOL_TRACE builds): device buffer for stage timestamps, or nullptr
bool coarse_supported(const mol_shape_t& s);
// takes the per-call buffers (query records, overflow flag) of a chunk of queries from the workspace arena
void coarse_plan(const mol_shape_t& s, int chunk, Arena& a, CoarseWs* ws);
// bytes of the two weight images
void coarse_weight_image_bytes(const mol_shape_t& s, size_t* w1_bytes, size_t* w2_bytes);
// weight images; sets *overflow (not cleared here) when a value does not fit fp16.  They depend on the weights only.
int coarse_prepare_weights(const mol_shape_t& s, const mol_weights_t& w, uint8_t* w1_img, uint8_t* w2_img,
                           int32_t* overflow, cudaStream_t st);
// operand records of queries [0, bc) of a chunk (once per chunk); resets ws.overflow and ORs *weights_overflow into it
int coarse_query_records(const mol_shape_t& s, const CoarseWs& ws, const float* qsub, const float* gq, int bc,
                         const int32_t* weights_overflow, cudaStream_t st);
// One pass over (part of) the corpus for queries [0, bc): fp16 operands / fp32 accumulation scores into `out`.
int coarse_run(const mol_shape_t& s, const mol_index_t& ix, const CoarseWs& ws, int bc, const CoarseOut& out,
               cudaStream_t st);
// scores[b, x] ~= MoL score for b < bc, x < N; row stride N
int coarse_scores(const mol_shape_t& s, const mol_index_t& ix, const CoarseWs& ws, int bc, float* scores,
                  cudaStream_t st);
// Tables of the strided sample: sample_map[j] = j * stride (j < count), main_map = every other physical tile in order
// (tiles - count entries).
int coarse_tile_maps(int32_t* sample_map, int32_t* main_map, int tiles, int stride, int count, cudaStream_t st);
// gi_half of the index in the kernel's logit order l' = m*P_Q + n (called by the index build); sets *overflow
// when a value does not fit fp16
int coarse_gi_image(const mol_shape_t& s, const float* gi_f32, uint16_t* gi_half, int64_t n, int32_t* overflow,
                    cudaStream_t st);
// flags[b] = 1 when the coarse candidate set cannot be accepted as containing the exact top-k:
//   cand_scores[b, kk-1] + 3 * max_j |cand_scores[b,j] - exact_scores[b,j]| + 2e-2 >= topk_scores[b, k-1]
// or when either overflow flag is set (operands did not fit fp16).
// Filter strategy (cnt != nullptr): a query with fewer than kk survivors uses its threshold as the bound on every
// item outside the candidate set; more survivors than `cap` (dropped candidates) flag the query.
// second != 0: the second chance of the filter strategy - only queries whose flag is set are tested again, now with EVERY
// survivor rescored (cand_scores / exact_scores = the (bc, cap) survivor buffers, kk = cap), so the bound on the items outside
// the candidate set is the filter threshold itself.
// stats (optional, 8 x int32, device): [0] += queries that go to the exact fallback (matrix strategy: first test; filter
// strategy: second test), [1] += queries with more than `cap` survivors, [2] = max survivors, [5] += queries accepted by
// the second chance.
int coarse_safety_flags(const float* cand_scores, const float* exact_scores, const float* topk_scores,
                        int bc, int kk, int k, const int32_t* overflow_a, const int32_t* overflow_b,
                        const int32_t* cnt, const float* thr, int thr_stride, int cap, int32_t* flags,
                        int32_t* stats, int second, cudaStream_t st);
// Appends every (score, item) of a (bc, n) matrix with !(score < thr[b]) to the per-query candidate buffers; column c
// of the matrix is item (c / 128) * tile_stride * 128 + c % 128.
int coarse_filter_matrix(const float* scores, int64_t n, int64_t ld, int bc, const float* thr, int thr_stride,
                         int32_t* cnt, float* cand_scores, int32_t* cand_idx, int cap, int tile_stride,
                         cudaStream_t st);

}  // namespace mol
