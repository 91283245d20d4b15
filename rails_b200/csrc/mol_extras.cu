// Callers / siblings of the MoL top-k path.
//
//  * select_valid_kernel — the seen-item masking + back-fill of
//      indexing/candidate_index.py:155-178 (CandidateIndex.get_top_k_outputs) on the GPU, no host sync:
//      from the over-fetched, score-sorted (B, k') list keep, per row, the first k ids that are not in the row's
//      invalid list; rows with fewer than k valid entries are back-filled with their first invalid entries; the
//      output keeps the original rank order of the chosen positions.
//  * MIPS brute force — rails/indexing/mips_top_k.py:74-81: all_logits = q . items^T (fp32), top-k, id gather.
#include <math_constants.h>

#include "common.cuh"
#include "mol_dotfilter.cuh"

namespace mol {

constexpr int SV_THREADS = 256;

__global__ void __launch_bounds__(SV_THREADS)
select_valid_kernel(const float* __restrict__ scores, const int64_t* __restrict__ ids,
                    const int64_t* __restrict__ invalid, int kp, int n0, int k, float* __restrict__ out_scores,
                    int64_t* __restrict__ out_ids, const int32_t* __restrict__ query_flags) {
  extern __shared__ int64_t sv_smem[];
  int64_t* inv = sv_smem;                                   // n0
  unsigned char* seen = reinterpret_cast<unsigned char*>(inv + n0);  // kp
  __shared__ int part[SV_THREADS];
  __shared__ int totals[2];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (query_flags && query_flags[b] == 0) return;
  const int64_t* row_ids = ids + (int64_t)b * kp;
  for (int i = tid; i < n0; i += SV_THREADS) inv[i] = invalid[(int64_t)b * n0 + i];
  __syncthreads();
  for (int j = tid; j < kp; j += SV_THREADS) {
    const int64_t id = row_ids[j];
    unsigned char s = 0;
    for (int i = 0; i < n0; ++i) s |= (inv[i] == id);
    seen[j] = s;
  }
  __syncthreads();
  // contiguous chunk per thread; two block scans (valid entries, then the complement of the kept ones)
  const int per = (kp + SV_THREADS - 1) / SV_THREADS;
  const int lo = tid * per, hi = min(kp, lo + per);
  auto block_exclusive = [&](int mine, int slot) -> int {
    part[tid] = mine;
    __syncthreads();
    if (tid == 0) {
      int run = 0;
      for (int t = 0; t < SV_THREADS; ++t) {
        int v = part[t];
        part[t] = run;
        run += v;
      }
      totals[slot] = run;
    }
    __syncthreads();
    int r = part[tid];
    __syncthreads();
    return r;
  };
  int n_valid = 0;
  for (int j = lo; j < hi; ++j) n_valid += !seen[j];
  const int valid_before = block_exclusive(n_valid, 0);
  const int total_valid = totals[0];
  const int gap = k - min(total_valid, k);  // rows with < k valid entries are back-filled
  // kept-valid flag: valid and among the first k valid;  "invalid" := everything else (candidate_index.py:163)
  int n_inv = 0, run = valid_before;
  for (int j = lo; j < hi; ++j) {
    bool keep = false;
    if (!seen[j]) {
      ++run;
      keep = run <= k;
    }
    n_inv += !keep;
  }
  const int inv_before = block_exclusive(n_inv, 1);
  int out_before;
  {
    // selected = kept-valid or (invalid and among the first `gap` invalid); count selected before this chunk
    int sel = 0, rv = valid_before, ri = inv_before;
    for (int j = lo; j < hi; ++j) {
      bool keep = false;
      if (!seen[j]) {
        ++rv;
        keep = rv <= k;
      }
      if (!keep) {
        ++ri;
        keep = ri <= gap;
      }
      sel += keep;
    }
    out_before = block_exclusive(sel, 0);
  }
  int rv = valid_before, ri = inv_before, o = out_before;
  for (int j = lo; j < hi; ++j) {
    bool keep = false;
    if (!seen[j]) {
      ++rv;
      keep = rv <= k;
    }
    if (!keep) {
      ++ri;
      keep = ri <= gap;
    }
    if (keep && o < k) {
      out_scores[(int64_t)b * k + o] = scores[(int64_t)b * kp + j];
      out_ids[(int64_t)b * k + o] = row_ids[j];
      ++o;
    }
  }
}

int launch_select_valid(const float* scores, const int64_t* ids, const int64_t* invalid_ids, int B, int kp, int n0, int k,
                        float* out_scores, int64_t* out_ids, const int32_t* query_flags, cudaStream_t st) {
  if (B == 0) return MOL_OK;
  const size_t smem = (size_t)n0 * sizeof(int64_t) + (size_t)kp + 16;
  MOL_CHECK_ARG(smem <= 200 * 1024, "k' + invalid list too large for one block (%zu bytes)", smem);
  if (smem > 48 * 1024)
    MOL_CUDA(cudaFuncSetAttribute(select_valid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  select_valid_kernel<<<B, SV_THREADS, smem, st>>>(scores, ids, invalid_ids, kp, n0, k, out_scores, out_ids, query_flags);
  MOL_LAUNCH_CHECK();
  return MOL_OK;
}

}  // namespace mol

using namespace mol;

extern "C" {

int mol_select_valid(const float* scores, const int64_t* ids, const int64_t* invalid_ids, int32_t B,
                     int32_t k_prime, int32_t n_invalid, int32_t k, float* out_scores, int64_t* out_ids,
                     mol_stream_t stream) {
  MOL_CHECK_ARG(B >= 0 && k >= 1 && k_prime >= k && n_invalid >= 0, "bad arguments (need k' >= k >= 1)");
  if (B == 0) return MOL_OK;
  MOL_CHECK_ARG(scores && ids && out_scores && out_ids && (n_invalid == 0 || invalid_ids), "NULL buffer");
  return launch_select_valid(scores, ids, invalid_ids, B, k_prime, n_invalid, k, out_scores, out_ids, nullptr,
                             static_cast<cudaStream_t>(stream));
}

namespace {
// Workspace of mol_mips_search: 8 x int32 stats (mol_search_stats reads them back) | (rows, N) fp32 matrix of the
// materialising path (also the fallback / sample matrix of the streaming path) | select scratch | streaming-path buffers.
struct MipsWs {
  int32_t* stats;
  float* mat;
  char* tk;
  size_t tk_bytes;
  int rows;
  int filter;  // sizes allow the streaming path (D is not known when the workspace is measured: checked again per call)
  mol::DotTopkPlan dp;
  size_t total;
};
int plan_mips(int64_t num_items, int32_t B, int32_t k, void* base, size_t cap, MipsWs* ws) {
  using namespace mol;
  Arena a(base, cap);
  const int64_t n = num_items > 0 ? num_items : 1;
  const int Bq = B > 0 ? B : 1;
  int64_t rows = (int64_t)score_matrix_budget() / (int64_t)(sizeof(float) * (size_t)n);
  if (rows < 1) rows = 1;
  if (rows > Bq) rows = Bq;
  ws->rows = (int)rows;
  ws->stats = a.take<int32_t>(8);
  ws->mat = a.take<float>((size_t)rows * (size_t)n);
  size_t topk;
  MOL_TRY(mol_topk_workspace_bytes(num_items, (int32_t)rows, k, &topk));
  ws->tk_bytes = topk + 256;
  ws->tk = a.take<char>(ws->tk_bytes);
  ws->filter = dot_topk_eligible(num_items, Bq, 32, k) ? 1 : 0;
  if (ws->filter) dot_topk_plan(a, num_items, Bq, 32, k, ws->mat, ws->rows, &ws->dp);
  ws->total = align_up(a.off, 256);
  if (base != nullptr && a.off > cap) {
    set_error("mips workspace too small: need %zu, got %zu", ws->total, cap);
    return MOL_ERR_WORKSPACE;
  }
  return MOL_OK;
}
}  // namespace

int mol_mips_workspace_bytes(int64_t num_items, int32_t B, int32_t k, size_t* bytes) {
  MOL_CHECK_ARG(bytes && num_items >= 0 && B >= 0 && k >= 1, "bad arguments");
  MipsWs ws;
  MOL_TRY(plan_mips(num_items, B, k, nullptr, 0, &ws));
  *bytes = ws.total + 256;
  return MOL_OK;
}

// scores[b, x] = <queries[b], items[x]> (fp32); out (B, N)
int mol_dot_scores(const float* items, const float* queries, int64_t num_items, int32_t D, int32_t B,
                   float* out_scores, mol_stream_t stream) {
  MOL_CHECK_ARG(num_items >= 0 && D >= 1 && B >= 0, "bad arguments");
  if (B == 0 || num_items == 0) return MOL_OK;
  MOL_CHECK_ARG(items && queries && out_scores, "NULL buffer");
  // C[M = B, N = items] = A[B, D] . W^T with W = items (n-major, k contiguous)
  return launch_linear(queries, items, nullptr, out_scores, B, (int)num_items, D, D, 1, ACT_NONE,
                       static_cast<cudaStream_t>(stream));
}

int mol_mips_search(const float* items, const int64_t* item_ids, const float* queries, int64_t num_items,
                    int32_t D, int32_t B, int32_t k, float* out_scores, int64_t* out_ids, void* workspace,
                    size_t workspace_bytes, mol_stream_t stream) {
  return mol_mips_search_cached(items, item_ids, queries, num_items, D, B, k, nullptr, out_scores, out_ids, workspace,
                                workspace_bytes, stream);
}

int mol_mips_search_cached(const float* items, const int64_t* item_ids, const float* queries, int64_t num_items,
                           int32_t D, int32_t B, int32_t k, float* item_norm_cache, float* out_scores,
                           int64_t* out_ids, void* workspace, size_t workspace_bytes, mol_stream_t stream) {
  MOL_CHECK_ARG(num_items >= 0 && D >= 1 && B >= 0 && k >= 1 && k <= MOL_MAX_K, "bad arguments");
  MOL_CHECK_ARG(num_items < (1ll << 31) - 256, "num_items must fit int32");
  if (k > num_items) {
    set_error("selected index k out of range (k=%d > %lld items)", k, (long long)num_items);
    return MOL_ERR_RANGE;
  }
  if (B == 0) return MOL_OK;
  MOL_CHECK_ARG(items && queries && out_scores && out_ids && workspace, "NULL buffer");
  MOL_CHECK_ARG(reinterpret_cast<uintptr_t>(workspace) % 256 == 0, "workspace must be 256-byte aligned");
  MipsWs ws;
  MOL_TRY(plan_mips(num_items, B, k, workspace, workspace_bytes, &ws));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MOL_CUDA(cudaMemsetAsync(ws.stats, 0, 8 * sizeof(int32_t), st));
  // streaming path (f4 "fused GEMM + top-k"): tcgen05 tf32 pass + threshold filter + fp32 rescoring of the survivors, no
  // (B, N) matrix - see mol_dotfilter.cuh
  if (ws.filter && dot_topk_eligible(num_items, B, D, k) && dot_topk_aligned(items, D, 0, queries, D)) {
    if (item_norm_cache) MOL_TRY(dot_topk_norm_cache(items, num_items, D, 0, D, item_norm_cache, st));
    return dot_topk_run(ws.dp, items, D, 0, D, item_norm_cache, 0.f, queries, D, out_scores, nullptr, out_ids, item_ids,
                        ws.stats, st);
  }
  const int64_t rows = ws.rows;
  float* mat = ws.mat;
  char* tk = ws.tk;
  const size_t tk_bytes = ws.tk_bytes;
  for (int b0 = 0; b0 < B; b0 += (int)rows) {
    const int nb = (B - b0 < rows) ? (B - b0) : (int)rows;
    MOL_TRY(mol_dot_scores(items, queries + (size_t)b0 * D, num_items, D, nb, mat, stream));
    MOL_TRY(mol_topk(mat, num_items, num_items, nb, k, item_ids, out_scores + (size_t)b0 * k, out_ids + (size_t)b0 * k,
                     tk, tk_bytes, stream));
  }
  return MOL_OK;
}

}  // extern "C"
