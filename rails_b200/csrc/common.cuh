// Internal helpers shared by the MoL kernels (not part of the C ABI).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/mol_b200.h"

namespace mol {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define MOL_CHECK_ARG(cond, ...)     \
  do {                               \
    if (!(cond)) {                   \
      ::mol::set_error(__VA_ARGS__); \
      return MOL_ERR_INVALID;        \
    }                                \
  } while (0)

#define MOL_CUDA(call)                                                                     \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) {                                                               \
      ::mol::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__,   \
                       __LINE__);                                                          \
      return MOL_ERR_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

#define MOL_LAUNCH_CHECK()                                                               \
  do {                                                                                   \
    ::mol::count_launch();                                                               \
    cudaError_t e_ = cudaGetLastError();                                                 \
    if (e_ != cudaSuccess) {                                                             \
      ::mol::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_),       \
                       __FILE__, __LINE__);                                              \
      return MOL_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define MOL_TRY(call)           \
  do {                          \
    int s_ = (call);            \
    if (s_ != MOL_OK) return s_; \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Byte budget of the (rows, N) fp32 score matrices the prefilter / MIPS / per-group paths materialise per launch (default
// 2 GiB).  MOL_B200_SCORE_MATRIX_BYTES shrinks it so that tests reach the multi-chunk code paths at small N; the
// *_workspace_bytes and the search call read the same value.
inline size_t score_matrix_budget() {
  const char* e = getenv("MOL_B200_SCORE_MATRIX_BYTES");
  if (e) {
    const long long v = atoll(e);
    if (v >= 4096) return (size_t)v;
  }
  return (size_t)2 << 30;
}

// Bump allocator over the caller's workspace.
struct Arena {
  char* base;
  size_t cap;
  size_t off;
  Arena(void* p, size_t n) : base(static_cast<char*>(p)), cap(n), off(0) {}
  // With base == nullptr the arena only measures.
  template <typename T>
  T* take(size_t count) {
    off = align_up(off, 256);
    T* r = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return r;
  }
  bool ok() const { return base == nullptr || off <= cap; }
};

struct Dims {
  int Dq, Dx, d, Pq, Px, L, Hq, Hgq, Hgi, H, u, Pq_proj;
};
inline Dims dims_of(const mol_shape_t& s) {
  Dims r;
  r.Dq = s.query_embedding_dim;
  r.Dx = s.item_embedding_dim;
  r.d = s.dot_product_dimension;
  r.Pq = s.query_dot_product_groups;
  r.Px = s.item_dot_product_groups;
  r.L = r.Pq * r.Px;
  r.Hq = s.query_hidden_dim;
  r.Hgq = s.gating_query_hidden_dim;
  r.Hgi = s.gating_item_hidden_dim;
  r.H = s.gating_qi_hidden_dim;
  r.u = s.num_uid_tables;
  r.Pq_proj = r.Pq - r.u;
  return r;
}

// ---- kernels implemented in the other translation units (host launchers) ----
enum Act { ACT_NONE = 0, ACT_SILU = 1 };

// C[M,N] = act(A[M,K] . W^T + bias); W element (n,k) at w[n*w_sn + k*w_sk]; bias may be null.
int launch_linear(const float* A, const float* W, const float* bias, float* C, int64_t M, int N,
                  int K, int64_t w_sn, int64_t w_sk, Act act, cudaStream_t st);
// h[b, j] = act(pre[b, j]) * pre[b, Hq + j]   (layers.py:36-43 / 67-74)
int launch_glu(const float* pre, float* h, int64_t B, int Hq, int kind, cudaStream_t st);
// rows of `groups` contiguous d-vectors: out = v / max(||v||, eps); optional fp16 copy.
int launch_l2norm_groups(const float* in, float* out_f32, __half* out_half, int64_t rows,
                         int groups, int d, float eps, cudaStream_t st);
// Q_sub assembly: projected groups + uid-embedding groups, then l2 norm (query_embeddings_fns.py:191-253)
int launch_query_assemble(const mol_shape_t& s, const mol_weights_t& w, const float* proj,
                          const int64_t* user_ids, float* qsub, int B, cudaStream_t st);

// Exact fp32 scoring.  cand == nullptr: scores[b, x] for x in [0, N) (ld = N).
// cand != nullptr: scores[b, j] = score(b, cand[b*ld + j]) for j < n_per_query; cand < 0 => -inf.
// query_flags (optional): only queries with flag != 0 are scored.
// w1t (L,H) / w2t (H,L) are the transposed qi-MLP weights (launch_transpose).
int launch_exact_scores(const mol_shape_t& s, const mol_weights_t& w, const mol_index_t& ix,
                        const float* w1t, const float* w2t, const float* qsub, const float* gq,
                        int B, const int32_t* cand, int64_t n_per_query, int64_t ld, float* scores,
                        const int32_t* query_flags, cudaStream_t st);
int launch_transpose(const float* in, float* out, int rows, int cols, cudaStream_t st);

// Top-k selection -----------------------------------------------------------------------------
// Level-1: per row b of `scores` (B rows, n columns, row stride ld) split into S segments; writes the
// segment-local top-kk (unsorted) to cand_scores/cand_idx [(b*S + s)*kk ...] (idx = column, -1 padded).
int launch_select_segments(const float* scores, int64_t n, int64_t ld, int B, int S, int kk,
                           float* cand_scores, int32_t* cand_idx, const int32_t* query_flags,
                           cudaStream_t st);
// Final: per row b, n candidates (row stride ld; payload == nullptr => payload = column index;
// payload < 0 ignored) -> top-kk sorted descending (ties: smaller payload first).  Writes
// out_scores (B,kk), and optionally out_idx (B,kk) int32 payloads and out_ids (B,kk) int64 =
// id_map ? id_map[payload] : payload.
int launch_select_final_i32(const float* scores, const int32_t* payload, int64_t n, int64_t ld,
                            int B, int kk, float* out_scores, int32_t* out_idx, int64_t* out_ids,
                            const int64_t* id_map, const int32_t* query_flags, cudaStream_t st);
// Same with int64 payloads (multi-GPU merge of global ids).
int launch_select_final_i64(const float* scores, const int64_t* payload, int64_t n, int64_t ld,
                            int B, int kk, float* out_scores, int64_t* out_ids, cudaStream_t st);

// Seen-item masking + back-fill of a score-sorted (B, kp) list (mol_select_valid); query_flags (optional): only rows with
// flag != 0 are written.
int launch_select_valid(const float* scores, const int64_t* ids, const int64_t* invalid_ids, int B, int kp, int n0, int k,
                        float* out_scores, int64_t* out_ids, const int32_t* query_flags, cudaStream_t st);

// Index build on the tensor cores (mol_linear_x3_sm100.cu): tf32 x 3 split GEMMs with the l2-norm / silu / fp16-image
// epilogues fused; `supported` looks at the shape, the corpus size and the alignment of raw_items / the biases
// (MOL_B200_INDEX_X3=0 keeps the CUDA-core build).
bool index_build_x3_supported(const mol_shape_t& s, const mol_weights_t& w, const mol_index_t& ix);
size_t index_build_x3_workspace_bytes(const mol_shape_t& s, int64_t N);
int index_build_x3(const mol_shape_t& s, const mol_weights_t& w, const mol_index_t& ix, void* workspace, cudaStream_t st);

int select_num_segments(int64_t n, int B, int kk);
int select_num_segments_streamed(int64_t n, int B, int kk);
int64_t select_streamed_slots(int64_t n, int64_t rows, int kk);

}  // namespace mol
