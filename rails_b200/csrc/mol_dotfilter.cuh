// Streaming dot-product top-k: the exact fp32 top-kk of  Q (R x K) . items^T (N x K)
// per row of Q WITHOUT materialising the (R, N) matrix.  Implemented in mol_dotfilter_sm100.cu.
//
//   1. a tcgen05 kind::tf32 pass over a strided sample of the item tiles writes a small (R, S) matrix; its m-th largest
//      value per row is the row's filter level t_r;
//   2. the same kernel streams EVERY item tile (TMA, fp32 straight from the caller's matrix, no converted copy) and
//      appends (value, position) of every pair with tf32 value >= t_r to the row's survivor buffer in its epilogue;
//   3. the survivors are re-scored in fp32 with the k-ascending fmaf chain of linear_kernel (bit-identical values), the
//      kk best are selected (ties: lower position first);
//   4. proof of completeness per row: |tf32 - fp32| <= E_r = 2^-8 |q_r| max_x |x| (operands keep 11 significant bits,
//      Cauchy-Schwarz), so every item with fp32 value >= t_r + E_r survived; if kk survivors reach t_r + E_r they are the
//      exact top-kk.  Rows that fail the test (or overflowed their buffer) are re-done by a plain fp32 pass + select.
//
// Replaces the (rows, N) fp32 GEMM + radix select behind MIPSBruteForceTopK (rails/indexing/mips_top_k.py:74-81), the
// MoLAvgTopK prefilter (rails/indexing/mol_top_k.py:352-360) and the per-group selections of MoLNaiveTopK / MoLCombTopK
// (mol_top_k.py:239-249 / :501-511).
#pragma once
#include "common.cuh"

namespace mol {

struct DotTopkPlan {
  // sizes
  int64_t N;
  int R, kk;
  int Rc;          // rows of Q per pass (<= 8192)
  int64_t S;       // sampled item rows (multiple of 128)
  int tile_stride; // the sample is every tile_stride-th item tile
  int m;           // sample rank that defines the filter level
  int cap;         // survivor capacity per row
  int expect;      // survivors per row the filter level aims at
  int rows_fb;     // rows of the fallback (rows_fb, N) matrix
  // buffers (carved by dot_topk_plan; `fb_scores` doubles as the (Rc, S) sample matrix)
  float* fb_scores;
  float* samp;          // (Rc, S) sample matrix
  int samp_aliases_fb;
  float *seg_scores, *samp_top, *level, *check, *cval, *cexact, *fb_seg_scores, *xmax;
  int32_t *seg_idx, *cnt, *cidx, *flags, *any_flag, *fb_seg_idx;
};

// True when the streaming path serves this problem (sizes only; pointer alignment is checked by dot_topk_run, which
// reports MOL_ERR_INVALID for a misaligned matrix - callers test dot_topk_aligned first).
bool dot_topk_eligible(int64_t N, int R, int K, int kk);
bool dot_topk_aligned(const float* items, int64_t pitch, int col0, const float* Q, int64_t q_pitch);
// Row-norm bound kept by the caller across calls: cache = 2 device floats {bound or < 0 when not yet computed, accumulator
// (0 on reset)}.  Computes max_x |x[col0 : col0 + K]|_2 into cache[0] when it is negative (one pass over the items), else
// two idle launches.  Pass `cache` as xmax_dev afterwards.
int dot_topk_norm_cache(const float* items, int64_t N, int64_t pitch, int col0, int K, float* cache, cudaStream_t st);
// Carves the buffers of one dot_topk_run out of `a` (measures only when a.base == nullptr).  `fb_scores` / `fb_rows`: the
// caller's existing (rows, N) fp32 score matrix (shared with its non-streaming path), or nullptr to take one here.
void dot_topk_plan(Arena& a, int64_t N, int R, int K, int kk, float* fb_scores, int fb_rows, DotTopkPlan* p);
// items: (N, pitch) fp32 row-major, columns [col0, col0 + K) take part; Q: (R, q_pitch) fp32, columns [0, K).
// xmax_dev: device scalar >= max_x |x[col0 : col0 + K]|_2; nullptr: xmax_host is used when > 0 (e.g. 1 for l2-normalised
// rows), else the bound is computed here (one pass over the items).
// out_scores (R, kk) fp32, out_idx (R, kk) int32 positions or nullptr, out_ids (R, kk) int64 = id_map ? id_map[pos] : pos
// or nullptr.  stats (device, 8 x int32, or nullptr): [0] += rows re-done by the fallback, [1] += rows whose survivor
// buffer overflowed, [2] = max survivors of a row, [3] = 1.
int dot_topk_run(const DotTopkPlan& p, const float* items, int64_t pitch, int col0, int K, const float* xmax_dev,
                 float xmax_host, const float* Q, int64_t q_pitch, float* out_scores, int32_t* out_idx, int64_t* out_ids,
                 const int64_t* id_map, int32_t* stats, cudaStream_t st);

}  // namespace mol
