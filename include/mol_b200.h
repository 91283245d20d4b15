/*
 * mol_b200.h — C ABI of the B200-native Mixture-of-Logits (MoL) brute-force top-k engine.
 *
 * This is the drop-in boundary for ONE hot path of bailuding/rails:
 *     rails/indexing/mol_top_k.py:84-130      MoLBruteForceTopK(.forward)
 *  -> rails/similarities/mol/similarity_fn.py:341-413  MoLSimilarity.forward
 * The reference has no FFI (it is pure PyTorch); the binding a maintainer adds is the ctypes
 * shim in rails_b200/_lib.py (shown in INTEGRATION.md).  Every entry point is `extern "C"`,
 * takes plain pointers and sizes, never allocates or frees device memory behind the caller,
 * never synchronises the device (except the *_host entry, which must), is ordered on the
 * stream passed in, and reports failure through an int status + mol_last_error().
 *
 * All "device" pointers are CUDA device pointers on the current device; weights are fp32,
 * row-major, exactly as they sit in the reference's state dict.
 */
#ifndef MOL_B200_H_
#define MOL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOL_OK 0
#define MOL_ERR_INVALID 1   /* bad argument / unsupported shape  -> Python ValueError  */
#define MOL_ERR_CUDA 2      /* CUDA runtime error                -> Python RuntimeError */
#define MOL_ERR_WORKSPACE 3 /* workspace too small               -> Python RuntimeError */
#define MOL_ERR_RANGE 4     /* k > number of items (torch.topk's "selected index k out of range") -> RuntimeError */

#define MOL_MAX_UID_TABLES 4
#define MOL_MAX_K 8192 /* largest k (and over-fetch k') served by the select kernels */

/* search modes */
#define MOL_MODE_AUTO 0   /* tensor-core coarse pass + fp32 rescoring when the shape supports it */
#define MOL_MODE_EXACT 1  /* fp32 CUDA-core scoring of every (query, item) pair               */
#define MOL_MODE_TENSOR 2 /* force the tcgen05 path (error if unsupported)                     */

typedef void* mol_stream_t; /* a cudaStream_t */

/* Shape / hyper-parameters of one MoL head.  Names follow the kwargs of the reference's
 * create_mol_interaction_module (modeling/similarity_utils.py:41-69). */
typedef struct mol_shape {
  int32_t query_embedding_dim;      /* D_q */
  int32_t item_embedding_dim;       /* D_x */
  int32_t dot_product_dimension;    /* d   */
  int32_t query_dot_product_groups; /* P_Q (including uid-embedding groups) */
  int32_t item_dot_product_groups;  /* P_X */
  int32_t query_hidden_dim;         /* GLU hidden (512) */
  int32_t gating_query_hidden_dim;  /* 128 */
  int32_t gating_item_hidden_dim;   /* 128 */
  int32_t gating_qi_hidden_dim;     /* H = 128 */
  int32_t query_nonlinearity;       /* 0 = geglu (erf GELU), 1 = swiglu */
  int32_t num_uid_tables;           /* u = len(uid_embedding_hash_sizes) */
  int32_t uid_hash_sizes[MOL_MAX_UID_TABLES];
  int32_t softmax_renorm; /* 1 iff softmax_dropout_rate > 0 (similarity_fn.py:43-45 runs in eval) */
  float temperature;      /* tau */
  float eps;              /* l2-norm / renorm clamp */
} mol_shape_t;

/* fp32 device pointers; layouts are the state-dict tensors of the reference MoLSimilarity. */
typedef struct mol_weights {
  const float* q_glu_w; /* _query_embeddings_fn._query_emb_proj_module.1._w   (D_q, 2*Hq)   */
  const float* q_glu_b; /* ...1._b                                            (1, 2*Hq)     */
  const float* q_out_w; /* ...2.weight                                        ((P_Q-u)*d, Hq) */
  const float* q_out_b; /* ...2.bias                                          ((P_Q-u)*d)   */
  const float* uid_emb[MOL_MAX_UID_TABLES]; /* _uid_embeddings_i.weight       (hash_i+1, d) */
  const float* x_w;     /* _item_embeddings_fn._item_emb_proj_module.1.weight (P_X*d, D_x)  */
  const float* x_b;     /* ...1.bias                                          (P_X*d)       */
  const float* gq_w1;   /* _gating_fn._query_only_partial_module.0.weight     (Hgq, D_q)    */
  const float* gq_b1;   /* ...0.bias                                                         */
  const float* gq_w2;   /* ...2.weight (no bias)                              (L, Hgq)      */
  const float* gi_w1;   /* _gating_fn._item_only_partial_module.1.weight      (Hgi, D_x)    */
  const float* gi_b1;   /* ...1.bias                                                         */
  const float* gi_w2;   /* ...3.weight (no bias)                              (L, Hgi)      */
  const float* qi_w1;   /* _gating_fn._qi_partial_module.1.weight             (H, L)        */
  const float* qi_b1;   /* ...1.bias                                          (H)           */
  const float* qi_w2;   /* ...3.weight                                        (L, H)        */
  const float* qi_b2;   /* ...3.bias                                          (L)           */
  /* Optional: device blob filled by mol_weights_prepare() with the operands that depend on the weights alone (the
   * transposed qi-MLP matrices of the fp32 kernels, the fp16 operand images of the tensor-core pass).  NULL: every
   * search call recomputes them inside its workspace (a few small launches). */
  const void* prepared;
} mol_weights_t;

/* The item-side cache ("index").  raw_items / item_ids are BORROWED from the caller exactly as
 * MoLTopKModule keeps references (mol_top_k.py:54-59,75-77); the four caches live inside one
 * caller-allocated blob laid out by mol_index_layout(). */
typedef struct mol_index {
  int64_t num_items;        /* N */
  const float* raw_items;   /* (N, D_x) fp32 */
  const int64_t* item_ids;  /* (N) int64; may be NULL => id == position */
  float* xsub_f32;          /* (N, P_X, d) l2-normalised sub-embeddings  (item_embeddings_fns.py:165-182) */
  float* gi_f32;            /* (N, L)      item-only gating partial      (similarity_fn.py:170-171)       */
  uint16_t* xsub_half;      /* (N_pad, P_X*d) fp16 copy streamed by the tensor-core pass (N_pad = N rounded up to 128; pad rows zero) */
  uint16_t* gi_half;        /* (N_pad, L)  fp16 copy, logit order of the tensor-core pass */
  int32_t* half_overflow;   /* device flag set by the build when a cached value does not fit fp16 (queries then fall back to exact) */
} mol_index_t;

const char* mol_version(void);
const char* mol_last_error(void);

/* 0 if the shape is servable at all (exact path); tensor_ok (may be NULL) is set to 1 when the
 * tcgen05 coarse pass supports it. */
int mol_shape_check(const mol_shape_t* shape, int32_t* tensor_ok);

/* Weight-derived operands, once per weight version: bytes of the blob / fill it (256-byte aligned device memory), then
 * point mol_weights_t::prepared at it. */
int mol_weights_prepared_bytes(const mol_shape_t* shape, size_t* bytes);
int mol_weights_prepare(const mol_shape_t* shape, const mol_weights_t* w, void* blob, size_t blob_bytes,
                        mol_stream_t stream);

/* Bytes of the cache blob for N items. */
int mol_index_bytes(const mol_shape_t* shape, int64_t num_items, size_t* bytes);
/* Carves `blob` into the four caches and fills *index (no device work). */
int mol_index_layout(const mol_shape_t* shape, int64_t num_items, const float* raw_items,
                     const int64_t* item_ids, void* blob, size_t blob_bytes, mol_index_t* index);
/* Scratch bytes needed by mol_index_build. */
int mol_index_build_workspace_bytes(const mol_shape_t* shape, int64_t num_items, size_t* bytes);
/* Item projection + l2-norm + item-only gating MLP over the whole corpus (K4,K5,K9 of SURVEY §2b). */
int mol_index_build(const mol_shape_t* shape, const mol_weights_t* w, const mol_index_t* index,
                    void* workspace, size_t workspace_bytes, mol_stream_t stream);

/* Workspace for mol_search / mol_score_all with at most B queries and top-k k. */
int mol_search_workspace_bytes(const mol_shape_t* shape, int64_t num_items, int32_t B, int32_t k,
                               int32_t mode, size_t* bytes);

/* MoLBruteForceTopK.forward (mol_top_k.py:99-130): queries (B, D_q) fp32 device, user_ids (B) int64
 * device or NULL when num_uid_tables == 0; out_scores (B,k) fp32, out_ids (B,k) int64 device.
 * sorted != 0 => descending by score (ties: lower position first). */
int mol_search(const mol_shape_t* shape, const mol_weights_t* w, const mol_index_t* index,
               const float* queries, const int64_t* user_ids, int32_t B, int32_t k, int32_t sorted,
               int32_t mode, float* out_scores, int64_t* out_ids, void* workspace,
               size_t workspace_bytes, mol_stream_t stream);

/* The same search with a per-query exclusion list (the seen-item filter of
 * CandidateIndex.get_top_k_outputs, indexing/candidate_index.py:144-178, moved inside the search): invalid_ids
 * (B, n_invalid) int64 device; an item whose id is in its query's list never enters the result (ids <= 0 in the list
 * match nothing when item ids are positive, as in the reference).  out = the top-k over the remaining items - what the
 * reference obtains by over-fetching k' = k + n_invalid and masking, without the over-fetch.  Needs k + n_invalid <=
 * min(N, MOL_MAX_K) (MOL_ERR_RANGE / MOL_ERR_INVALID otherwise; callers then over-fetch and call mol_select_valid).
 * Tensor path: the ids are struck from the survivor buffers of the fused filter (or from the coarse top-K') before the
 * fp32 rescoring; exact mode and the per-query exact fallback over-fetch internally. */
int mol_search_excluding_workspace_bytes(const mol_shape_t* shape, int64_t num_items, int32_t B, int32_t k,
                                         int32_t n_invalid, int32_t mode, size_t* bytes);
int mol_search_excluding(const mol_shape_t* shape, const mol_weights_t* w, const mol_index_t* index,
                         const float* queries, const int64_t* user_ids, int32_t B, int32_t k, int32_t sorted,
                         int32_t mode, const int64_t* invalid_ids, int32_t n_invalid, float* out_scores,
                         int64_t* out_ids, void* workspace, size_t workspace_bytes, mol_stream_t stream);

/* Counters of the LAST mol_search / mol_search_host call that used `workspace` (8 x int32, copied to host_stats; this
 * call synchronises the stream): [0] queries re-done by the exact fallback, [1] queries whose candidate filter
 * overflowed, [2] largest survivor count of a query, [3] 1 if the fused-filter strategy ran, [4] 1 if the tensor-core
 * path ran, [5] queries the first acceptance test refused and the second chance (every survivor of the fused filter
 * rescored in fp32) accepted, [6] K', [7] survivor capacity per query. */
int mol_search_stats(const void* workspace, int32_t* host_stats, mol_stream_t stream);

/* Same call with HOST buffers for queries / user_ids / outputs (pinned memory recommended): copies
 * in, searches, copies out and synchronises the stream.  Device staging lives in the workspace
 * (mol_search_workspace_bytes already accounts for it). */
int mol_search_host(const mol_shape_t* shape, const mol_weights_t* w, const mol_index_t* index,
                    const float* host_queries, const int64_t* host_user_ids, int32_t B, int32_t k,
                    int32_t sorted, int32_t mode, float* host_out_scores, int64_t* host_out_ids,
                    void* workspace, size_t workspace_bytes, mol_stream_t stream);

/* MoLSimilarity.forward, B'==1 branch (similarity_fn.py:341-413): all (B, N) scores, fp32. */
int mol_score_all(const mol_shape_t* shape, const mol_weights_t* w, const mol_index_t* index,
                  const float* queries, const int64_t* user_ids, int32_t B, float* out_scores,
                  void* workspace, size_t workspace_bytes, mol_stream_t stream);

/* Diagnostic: the (B, N) output of the tcgen05 coarse pass alone (fp16 operands, fp32 accumulation;
 * approximate — it only ranks candidates for the fp32 rescoring inside mol_search). */
int mol_score_all_coarse(const mol_shape_t* shape, const mol_weights_t* w, const mol_index_t* index,
                         const float* queries, const int64_t* user_ids, int32_t B, float* out_scores,
                         void* workspace, size_t workspace_bytes, mol_stream_t stream);

/* Query prologue only (query_embeddings_fns.py:175-254 + similarity_fn.py:166-169):
 * out_qsub (B, P_Q, d) fp32 l2-normalised, out_gq (B, L) fp32. */
int mol_query_prologue(const mol_shape_t* shape, const mol_weights_t* w, const float* queries,
                       const int64_t* user_ids, int32_t B, float* out_qsub, float* out_gq,
                       void* workspace, size_t workspace_bytes, mol_stream_t stream);

/* Multi-GPU merge: parts (R, B, k) scores + ids from R corpus shards -> global (B, k). */
int mol_merge_topk_workspace_bytes(int32_t R, int32_t B, int32_t k, size_t* bytes);
int mol_merge_topk(const float* part_scores, const int64_t* part_ids, int32_t R, int32_t B,
                   int32_t k, float* out_scores, int64_t* out_ids, void* workspace,
                   size_t workspace_bytes, mol_stream_t stream);

/* The same exchange with ONE buffer per rank: mol_pack_topk packs a rank's (B, k_valid) partial list (k_valid <= k; the
 * remaining k - k_valid entries of each row are marked invalid - a shard with fewer than k items) into (B, k) entries of
 * MOL_PACKED_ENTRY_BYTES {int64 id, float score, int32 valid}; after a single all-gather, mol_merge_topk_packed reads the
 * gathered (R, B, k) entries and writes the global (B, k).  Ids may be any int64 (validity is explicit). */
#define MOL_PACKED_ENTRY_BYTES 16
int mol_pack_topk(const float* scores, const int64_t* ids, int32_t B, int32_t k_valid, int32_t k, void* out_packed,
                  mol_stream_t stream);
int mol_merge_topk_packed_workspace_bytes(int32_t R, int32_t B, int32_t k, size_t* bytes);
int mol_merge_topk_packed(const void* gathered, int32_t R, int32_t B, int32_t k, float* out_scores, int64_t* out_ids,
                          void* workspace, size_t workspace_bytes, mol_stream_t stream);

/* Generic row-wise top-k of a device score matrix (B rows, n columns, row stride ld): the
 * replacement of torch.topk(dim=1, largest=True, sorted=True) at mol_top_k.py:123-129.
 * out_idx (B,k) int64 column indices (or id_map[column] when id_map != NULL). */
int mol_topk_workspace_bytes(int64_t n, int32_t B, int32_t k, size_t* bytes);
int mol_topk(const float* scores, int64_t n, int64_t ld, int32_t B, int32_t k, const int64_t* id_map,
             float* out_scores, int64_t* out_idx, void* workspace, size_t workspace_bytes,
             mol_stream_t stream);

/* ---- callers / siblings of the path ---- */

/* Seen-item masking + back-fill of CandidateIndex.get_top_k_outputs (indexing/candidate_index.py:155-178):
 * scores / ids (B, k') sorted by score (the over-fetched top-k'), invalid_ids (B, n_invalid) int64; keeps per row the
 * first k ids not in the row's invalid list, back-fills short rows with their first invalid entries, preserves
 * rank order.  No workspace, no host sync (the reference syncs in torch.nonzero). */
int mol_select_valid(const float* scores, const int64_t* ids, const int64_t* invalid_ids, int32_t B,
                     int32_t k_prime, int32_t n_invalid, int32_t k, float* out_scores, int64_t* out_ids,
                     mol_stream_t stream);

/* MIPSBruteForceTopK.forward (rails/indexing/mips_top_k.py:74-81): fp32 q . items^T, top-k, id gather.
 * items (N, D) fp32 device, item_ids (N) int64 or NULL, queries (B, D) fp32. */
int mol_mips_workspace_bytes(int64_t num_items, int32_t B, int32_t k, size_t* bytes);
int mol_mips_search(const float* items, const int64_t* item_ids, const float* queries, int64_t num_items,
                    int32_t D, int32_t B, int32_t k, float* out_scores, int64_t* out_ids, void* workspace,
                    size_t workspace_bytes, mol_stream_t stream);
/* The same search with a caller-kept bound on the item norms: item_norm_cache = 2 device floats, {-1, 0} when the items
 * are new or have changed; the first call fills it (one pass over the items), later calls skip that pass.  The streaming
 * path (>= 64k items, D a multiple of 32: tcgen05 tf32 pass + threshold filter + fp32 rescoring, no (B, N) matrix) needs
 * the bound for its per-query completeness test; NULL = recomputed by every call.  Workspace: mol_mips_workspace_bytes;
 * mol_search_stats() on it reports [0] queries re-done by the plain fp32 pass, [1] survivor-buffer overflows, [2] max
 * survivors, [3] 1 when the streaming path ran. */
int mol_mips_search_cached(const float* items, const int64_t* item_ids, const float* queries, int64_t num_items,
                           int32_t D, int32_t B, int32_t k, float* item_norm_cache, float* out_scores,
                           int64_t* out_ids, void* workspace, size_t workspace_bytes, mol_stream_t stream);
/* DotProductSimilarity.forward, (1, X, D) branch (rails/similarities/dot_product_similarity_fn.py:46-51): (B, N) fp32. */
int mol_dot_scores(const float* items, const float* queries, int64_t num_items, int32_t D, int32_t B,
                   float* out_scores, mol_stream_t stream);

/* MoLAvgTopK (rails/indexing/mol_top_k.py:296-429): dot-product prefilter on the
 * group-averaged sub-embeddings (fp32 here; the reference keeps them in bf16), top avg_top_k positions, exact fp32 MoL
 * on those, final top-k.  avg_items (N, d) = mean over P_X of the index's X_sub, filled by mol_index_avg_embeddings. */
int mol_index_avg_embeddings(const mol_shape_t* shape, const mol_index_t* index, float* out_avg, mol_stream_t stream);
int mol_search_avg_workspace_bytes(const mol_shape_t* shape, int64_t num_items, int32_t B, int32_t k,
                                   int32_t avg_top_k, size_t* bytes);
int mol_search_avg(const mol_shape_t* shape, const mol_weights_t* w, const mol_index_t* index, const float* avg_items,
                   const float* queries, const int64_t* user_ids, int32_t B, int32_t k, int32_t avg_top_k,
                   float* out_scores, int64_t* out_ids, void* workspace, size_t workspace_bytes,
                   mol_stream_t stream);

/* MoLNaiveTopK / MoLCombTopK (rails/indexing/mol_top_k.py:133-293 and :432-551): for every
 * (query group n, item group m) the k_per_group items with the largest fp32 <Q_sub[b,n], X_sub[x,m]> (the reference keeps
 * the item operand in bf16), plus - Comb only, avg_top_k > 0 - the avg_top_k items of MoLAvgTopK.topk_ids; the union is
 * sorted by position, scored with exact fp32 MoL, duplicates get the reference's -32767 sentinel, and ALL
 * C = P_Q * P_X * k_per_group + avg_top_k candidates come back sorted by score (the reference overwrites the caller's k
 * with C, :256 / :518).  out_scores (B, C) fp32, out_ids (B, C) int64.  avg_items may be NULL when avg_top_k == 0.
 * FAISS (use_faiss=True) is out of scope. */
int mol_search_groups_workspace_bytes(const mol_shape_t* shape, int64_t num_items, int32_t B, int32_t k_per_group,
                                      int32_t avg_top_k, size_t* bytes);
int mol_search_groups(const mol_shape_t* shape, const mol_weights_t* w, const mol_index_t* index,
                      const float* avg_items, const float* queries, const int64_t* user_ids, int32_t B,
                      int32_t k_per_group, int32_t avg_top_k, float* out_scores, int64_t* out_ids, void* workspace,
                      size_t workspace_bytes, mol_stream_t stream);

/* Optional CUDA-event timing of the dominant scoring kernel (the tcgen05 coarse pass, or the fp32
 * kernel in MOL_MODE_EXACT) on the stream it is launched on: enable, run searches, collect the summed
 * device time and the number of timed launches (collect synchronises on the recorded events). */
void mol_profile_enable(int32_t on);
int mol_profile_collect(double* total_ms, int32_t* launches);

/* Number of kernels this library launched since the last reset (bench.py's gpu_launches). */
int64_t mol_launch_count(void);
void mol_launch_count_reset(void);

#ifdef __cplusplus
}
#endif
#endif /* MOL_B200_H_ */
