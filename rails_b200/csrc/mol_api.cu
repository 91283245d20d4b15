// C-ABI entry points (include/mol_b200.h): argument checking, workspace carving, kernel sequencing.
//
// Path served (reference): rails/indexing/mol_top_k.py:99-130 -> rails/similarities/mol/similarity_fn.py:341-413.
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <utility>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"
#include "mol_coarse.cuh"
#include "mol_dotfilter.cuh"

namespace mol {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- optional timing of the dominant (scoring) kernel, for bench.py's roofline -----------------
static bool g_prof_on = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_events;
static size_t g_prof_used = 0;
static void prof_begin(cudaStream_t st) {
  if (!g_prof_on) return;
  if (g_prof_used == g_prof_events.size()) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    g_prof_events.emplace_back(a, b);
  }
  cudaEventRecord(g_prof_events[g_prof_used].first, st);
}
static void prof_end(cudaStream_t st) {
  if (!g_prof_on) return;
  cudaEventRecord(g_prof_events[g_prof_used].second, st);
  ++g_prof_used;
}

static int check_shape(const mol_shape_t* s) {
  MOL_CHECK_ARG(s != nullptr, "shape is NULL");
  Dims D = dims_of(*s);
  MOL_CHECK_ARG(D.Dq > 0 && D.Dx > 0 && D.d > 0 && D.Pq > 0 && D.Px > 0, "non-positive dimension");
  MOL_CHECK_ARG(D.d % 4 == 0, "dot_product_dimension must be a multiple of 4 (got %d)", D.d);
  MOL_CHECK_ARG(D.u >= 0 && D.u <= MOL_MAX_UID_TABLES && D.u < D.Pq, "bad num_uid_tables %d", D.u);
  MOL_CHECK_ARG(D.Hq > 0, "query_hidden_dim must be > 0 (GLU query projection)");
  MOL_CHECK_ARG(D.Hgq > 0 && D.Hgi > 0, "gating query/item hidden dims must be > 0");
  MOL_CHECK_ARG(D.H == 128, "gating_qi_hidden_dim must be 128 (got %d)", D.H);
  MOL_CHECK_ARG(D.L == 32 || D.L == 64 || D.L == 128 || D.L == 256,
                "P_Q*P_X must be one of 32/64/128/256 (got %d)", D.L);
  MOL_CHECK_ARG(s->query_nonlinearity == 0 || s->query_nonlinearity == 1, "bad query_nonlinearity");
  MOL_CHECK_ARG(s->temperature > 0.f, "temperature must be > 0");
  for (int i = 0; i < D.u; ++i) MOL_CHECK_ARG(s->uid_hash_sizes[i] > 0, "uid hash size must be > 0");
  return MOL_OK;
}

static int check_weights(const mol_shape_t* s, const mol_weights_t* w) {
  MOL_CHECK_ARG(w != nullptr, "weights is NULL");
  MOL_CHECK_ARG(w->q_glu_w && w->q_glu_b && w->q_out_w && w->q_out_b, "query projection weights missing");
  MOL_CHECK_ARG(w->x_w && w->x_b, "item projection weights missing");
  MOL_CHECK_ARG(w->gq_w1 && w->gq_b1 && w->gq_w2, "query-only gating weights missing");
  MOL_CHECK_ARG(w->gi_w1 && w->gi_b1 && w->gi_w2, "item-only gating weights missing");
  MOL_CHECK_ARG(w->qi_w1 && w->qi_b1 && w->qi_w2 && w->qi_b2, "qi gating weights missing");
  for (int i = 0; i < s->num_uid_tables; ++i) MOL_CHECK_ARG(w->uid_emb[i], "uid table %d missing", i);
  return MOL_OK;
}

static inline int64_t pad128(int64_t n) { return (n + 127) / 128 * 128; }

// ---- index -----------------------------------------------------------------------------------
struct IndexLayout {
  size_t xsub_f32, gi_f32, xsub_half, gi_half, half_overflow, total;
};
static IndexLayout index_layout(const Dims& D, int64_t N) {
  IndexLayout l;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    off = align_up(off, 1024);
    size_t r = off;
    off += bytes;
    return r;
  };
  l.xsub_f32 = take((size_t)N * D.Px * D.d * sizeof(float));
  l.gi_f32 = take((size_t)N * D.L * sizeof(float));
  l.xsub_half = take((size_t)pad128(N) * D.Px * D.d * sizeof(uint16_t));
  l.gi_half = take((size_t)pad128(N) * D.L * sizeof(uint16_t));
  l.half_overflow = take(sizeof(int32_t));
  l.total = align_up(off, 1024);
  return l;
}

// ---- search workspace ------------------------------------------------------------------------
// Tensor mode has two candidate-generation strategies:
//   matrix : the coarse pass writes the (chunk, N) score matrix, a radix select keeps the K' best per query
//            (small corpora: the matrix is cheap).
//   filter : a coarse pass over a SAMPLE of the corpus gives each query a score threshold that about
//            `cand_target` items of the whole corpus exceed; the main coarse pass then appends every
//            (score, item) above the threshold to a per-query buffer inside the scoring kernel's epilogue, so
//            no score matrix is written or re-read.  Too many / too few survivors are caught by the safety check.
// Operands derived from the weights alone: the transposed qi-MLP matrices of the fp32 kernels and the fp16 UMMA
// images of the tensor-core pass.  Either inside the caller's mol_weights_t::prepared blob (mol_weights_prepare, once
// per weight version) or, when that is NULL, inside the search workspace and recomputed by every call.
struct Prepared {
  float *w1t, *w2t;          // (L,H) / (H,L)
  uint8_t *w1_img, *w2_img;  // nullptr when the shape has no tensor path
  int32_t* overflow;         // a weight did not fit fp16
  size_t total;
};
static void plan_prepared(const mol_shape_t& s, void* base, size_t cap, Prepared* p) {
  Dims D = dims_of(s);
  Arena a(base, cap);
  p->w1t = a.take<float>((size_t)D.L * D.H);
  p->w2t = a.take<float>((size_t)D.L * D.H);
  p->overflow = a.take<int32_t>(1);
  p->w1_img = p->w2_img = nullptr;
  if (coarse_supported(s)) {
    size_t b1, b2;
    coarse_weight_image_bytes(s, &b1, &b2);
    p->w1_img = a.take<uint8_t>(b1);
    p->w2_img = a.take<uint8_t>(b2);
  }
  p->total = align_up(a.off, 256);
}
static int run_prepare(const mol_shape_t& s, const mol_weights_t& w, const Prepared& p, cudaStream_t st) {
  Dims D = dims_of(s);
  MOL_TRY(launch_transpose(w.qi_w1, p.w1t, D.H, D.L, st));  // (H,L) -> (L,H)
  MOL_TRY(launch_transpose(w.qi_w2, p.w2t, D.L, D.H, st));  // (L,H) -> (H,L)
  MOL_CUDA(cudaMemsetAsync(p.overflow, 0, sizeof(int32_t), st));
  if (p.w1_img) MOL_TRY(coarse_prepare_weights(s, w, p.w1_img, p.w2_img, p.overflow, st));
  return MOL_OK;
}
// the caller's prepared blob if there is one, else `local` (filled now)
static int get_prepared(const mol_shape_t& s, const mol_weights_t& w, const Prepared& local, Prepared* out,
                        cudaStream_t st) {
  if (w.prepared) {
    plan_prepared(s, const_cast<void*>(w.prepared), (size_t)-1, out);
    return MOL_OK;
  }
  *out = local;
  return run_prepare(s, w, local, st);
}

struct NvtxRange {  // visible in nsys / ncu --nvtx; a no-op without a profiler attached
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

constexpr int kNumStats = 8;  // int32 counters at the start of the search workspace (mol_search_stats)

struct SearchWs {
  int32_t* stats;       // [0] queries re-done exactly, [1] filter overflows, [2] max survivors, [3] filter strategy used,
                        // [4] tensor path used, [5] queries accepted by the second chance (all survivors rescored), [6] K', [7] survivor capacity
  // query prologue
  float *pre, *h, *proj, *hq, *qsub, *gq;
  Prepared prep;        // workspace-local copy of the weight-derived operands (used when weights->prepared is NULL)
  // host-entry staging
  float* stage_q;
  int64_t* stage_uid;
  float* stage_out_scores;
  int64_t* stage_out_ids;
  // scoring / selection
  float* scores;        // matrix mode: (chunk, N) coarse or exact scores; filter mode: (chunk, sample) + exact fallback (chunk_fb, N)
  float* seg_scores;    // (chunk, S, kk)
  int32_t* seg_idx;
  float* cand_scores;   // (chunk, K') coarse top-K'
  int32_t* cand_idx;
  float* exact_scores;  // (chunk, K') rescored
  int32_t* flags;       // (chunk) fallback flags
  // filter mode
  float* samp_top;      // (chunk, m) sorted top-m of the sample; thr[b] = samp_top[b*m + m-1]
  int32_t* fcnt;        // (chunk) survivors per query
  float* fscores;       // (chunk, cap)
  int32_t* fidx;        // (chunk, cap)
  float* fexact;        // (chunk, cap) fp32 scores of every survivor (second chance of refused queries)
  int32_t *map_sample, *map_main;  // logical -> physical tile tables of the strided sample / its complement
  CoarseWs coarse;
  int chunk;            // queries per chunk
  int chunk_fb;         // queries per exact-fallback sub-chunk
  int Kp;               // K' (tensor mode)
  int S;                // segments for the (chunk, n) select
  int filter;           // 1 = filter strategy
  int64_t sample;       // sample items (multiple of 128)
  int samp_stride;      // the sample is every samp_stride-th item tile (1: the first `sample` items)
  int m;                // sample rank that defines the threshold
  int cap;              // survivor capacity per query
  // per-query exclusion lists (mol_search_excluding): internal over-fetch of the fp32 matrix paths
  int n0;               // excluded ids per query (0: none)
  float* ex_scores;     // (chunk, k + n0)
  int64_t* ex_ids;      // (chunk, k + n0)
  size_t total;
};

static int coarse_candidates(int k, int64_t N) {
  // K' = max(2k, k + 156) rounded up to 32 (fp16 operands: the exact top-k sits in the coarse top-(k + k/4) in
  // every case tried, tests/sim_coarse.py; rescoring K' pairs costs ~K'/N of the coarse pass), capped by N.
  int64_t kp = 2 * (int64_t)k;
  if (kp < (int64_t)k + 156) kp = (int64_t)k + 156;
  kp = (kp + 31) / 32 * 32;
  if (kp > MOL_MAX_K) kp = MOL_MAX_K;
  if (kp > N) kp = N;
  if (kp < k) kp = k;
  return (int)kp;
}

// MOL_MODE_AUTO below ~32k (query, item) pairs: the fp32 kernel over every pair (~15 us + one select) beats the tensor
// path's fixed sequence - query records, coarse pass, select, rescoring, select, acceptance test, idle fallback launches
// (~75 us; BASELINE config 1: one query over 3883 items, 116 -> 75 us of kernels).  MOL_B200_TENSOR_MIN_PAIRS overrides it
// (the test suite sets 0 so that MODE_AUTO keeps exercising the tensor path on the small reference fixtures).
constexpr int64_t kTensorMinPairs = 1ll << 15;
static int64_t tensor_min_pairs() {
  const char* e = getenv("MOL_B200_TENSOR_MIN_PAIRS");
  if (e) {
    const long long v = atoll(e);
    if (v >= 0) return (int64_t)v;
  }
  return kTensorMinPairs;
}
static bool use_tensor(const mol_shape_t& s, int mode, int64_t B, int64_t N) {
  if (mode == MOL_MODE_EXACT) return false;
  if (mode == MOL_MODE_AUTO && B * N < tensor_min_pairs()) return false;
  return coarse_supported(s);
}

constexpr int64_t kFilterMinItems = 1 << 16;
// ... and only for batches whose (B, N) coarse score matrix would cost more to write and select from than the threshold
// pass does (sample pass + two selections + filter launches ~ 50 us): below 2^24 pairs (64 MB of fp32 scores) the matrix
// strategy is one coarse launch + one segmented select
constexpr int64_t kFilterMinPairs = 1ll << 24;
// (MOL_B200_FILTER_MIN_PAIRS overrides it, so that tests reach the filter strategy with a handful of queries)
static int64_t filter_min_pairs() {
  const char* e = getenv("MOL_B200_FILTER_MIN_PAIRS");
  if (e) {
    const long long v = atoll(e);
    if (v >= 0) return (int64_t)v;
  }
  return kFilterMinPairs;
}

static int plan_search(const mol_shape_t& s, int64_t N, int B, int k, int mode, void* base,
                       size_t cap, SearchWs* ws, int n0 = 0) {
  Dims D = dims_of(s);
  Arena a(base, cap);
  const bool tensor = use_tensor(s, mode, B, N);
  ws->stats = a.take<int32_t>(kNumStats);  // (offset 0 of the workspace: mol_search_stats reads it back)
  ws->pre = a.take<float>((size_t)B * 2 * D.Hq);
  ws->h = a.take<float>((size_t)B * D.Hq);
  ws->proj = a.take<float>((size_t)B * D.Pq_proj * D.d);
  ws->hq = a.take<float>((size_t)B * D.Hgq);
  ws->qsub = a.take<float>((size_t)B * D.Pq * D.d);
  ws->gq = a.take<float>((size_t)B * D.L);
  {
    Prepared measure;
    plan_prepared(s, nullptr, 0, &measure);
    char* blob = a.take<char>(measure.total);
    plan_prepared(s, blob, blob ? measure.total : 0, &ws->prep);
  }
  ws->stage_q = a.take<float>((size_t)B * D.Dq);
  ws->stage_uid = a.take<int64_t>((size_t)B);
  ws->stage_out_scores = a.take<float>((size_t)B * k);
  ws->stage_out_ids = a.take<int64_t>((size_t)B * k);
  ws->Kp = tensor ? coarse_candidates(k, N) : k;
  // filter strategy: capacity 4 K' (>= 4096), aiming at capacity / 4 survivors
  ws->cap = 4 * ws->Kp < 4096 ? 4096 : 4 * ws->Kp;
  ws->filter = (tensor && N >= kFilterMinItems && (int64_t)ws->cap * 16 <= N && (int64_t)B * N >= filter_min_pairs()) ? 1 : 0;
  const int64_t n_rows = N > 0 ? N : 1;
  const int kx = k + n0;  // what the fp32 matrix paths select when ids are excluded (internal over-fetch)
  if (n0 > 0 && tensor && !ws->filter) {
    // matrix strategy: the excluded ids are struck from the coarse top-K', so K' grows by the list length (the filter
    // strategy strikes them from the survivor buffers, before the K' best are chosen)
    int64_t kp = (int64_t)ws->Kp + n0;
    if (kp > MOL_MAX_K) kp = MOL_MAX_K;
    if (kp > N) kp = N;
    ws->Kp = (int)kp;
  }
  if (ws->filter) {
    const int64_t target = ws->cap / 4;
    int64_t sample = 32 * N / target;  // -> threshold rank m >= 32 in the sample
    if (sample < 32768) sample = 32768;
    sample = (sample + 127) / 128 * 128;
    // the sample is every stride-th item tile, so that an ordered corpus (by popularity, id, cluster, ...) still gives
    // a representative threshold; the main pass scores the complement
    const int64_t tiles = (N + 127) / 128;
    int64_t stride = tiles / (sample / 128);
    if (stride < 2) stride = 1;  // (cannot happen with cap * 16 <= N; kept as the contiguous fallback)
    ws->sample = sample;
    ws->samp_stride = (int)stride;
    ws->m = (int)((target * sample + N - 1) / N);
    ws->chunk = B < 1 ? 1 : B;
    int64_t fb = (int64_t)(1ull << 30) / (int64_t)(sizeof(float) * (size_t)n_rows);
    if (fb < 1) fb = 1;
    ws->chunk_fb = (int)(fb < ws->chunk ? fb : ws->chunk);
    // the (chunk, sample) matrix of the threshold pass; query chunks keep it <= 2 GiB
    int64_t max_rows = (int64_t)(2ull << 30) / (int64_t)(sizeof(float) * (size_t)sample);
    if (max_rows < 1) max_rows = 1;
    if (ws->chunk > max_rows) ws->chunk = (int)max_rows;
    if (ws->chunk_fb > ws->chunk) ws->chunk_fb = ws->chunk;
    size_t n_scores = (size_t)ws->chunk * (size_t)sample;
    if ((size_t)ws->chunk_fb * (size_t)n_rows > n_scores) n_scores = (size_t)ws->chunk_fb * (size_t)n_rows;
    ws->scores = a.take<float>(n_scores);
    // segment survivors of any (rows <= chunk, kk <= max(m, k)) select: rows * S(rows) < rows + 2 * 148 + 1
    const int kk_max = ws->m > kx ? ws->m : kx;
    const size_t seg = (size_t)(ws->chunk + 2 * 148 + 1) * kk_max;
    ws->S = 0;
    ws->seg_scores = a.take<float>(seg);
    ws->seg_idx = a.take<int32_t>(seg);
    ws->samp_top = a.take<float>((size_t)ws->chunk * ws->m);
    ws->fcnt = a.take<int32_t>((size_t)ws->chunk);
    ws->fscores = a.take<float>((size_t)ws->chunk * ws->cap);
    ws->fidx = a.take<int32_t>((size_t)ws->chunk * ws->cap);
    ws->fexact = a.take<float>((size_t)ws->chunk * ws->cap);
    ws->map_sample = a.take<int32_t>((size_t)(sample / 128));
    ws->map_main = a.take<int32_t>((size_t)tiles);
  } else {
    // query chunk so that the (chunk, N) score matrix stays <= 4 GiB
    int64_t max_rows = (int64_t)(4ull << 30) / (int64_t)(sizeof(float) * (size_t)n_rows);
    if (max_rows < 1) max_rows = 1;
    int chunk = (int)(B < max_rows ? B : max_rows);
    if (chunk < 1) chunk = 1;
    ws->chunk = chunk;
    ws->chunk_fb = chunk;
    ws->sample = 0;
    ws->samp_stride = 1;
    ws->m = 0;
    ws->S = 0;
    ws->scores = a.take<float>((size_t)chunk * (size_t)N);
    const size_t seg = (size_t)(chunk + 2 * 148 + 1) * (ws->Kp > kx ? ws->Kp : kx);
    ws->seg_scores = a.take<float>(seg);
    ws->seg_idx = a.take<int32_t>(seg);
    ws->samp_top = nullptr;
    ws->fcnt = nullptr;
    ws->fscores = nullptr;
    ws->fidx = nullptr;
    ws->fexact = nullptr;
    ws->map_sample = ws->map_main = nullptr;
  }
  if (tensor) {  // the coarse kernel walks its (tile, query) units in 32-bit arithmetic
    const int64_t tiles = (N + 127) / 128;
    const int64_t max_rows = tiles > 0 ? ((1ll << 31) - 1) / tiles : ws->chunk;
    if (ws->chunk > max_rows) ws->chunk = (int)(max_rows < 1 ? 1 : max_rows);
    if (ws->chunk_fb > ws->chunk) ws->chunk_fb = ws->chunk;
  }
  ws->n0 = n0;
  ws->ex_scores = nullptr;
  ws->ex_ids = nullptr;
  if (n0 > 0) {
    ws->ex_scores = a.take<float>((size_t)ws->chunk * (size_t)(k + n0));
    ws->ex_ids = a.take<int64_t>((size_t)ws->chunk * (size_t)(k + n0));
  }
  ws->cand_scores = a.take<float>((size_t)ws->chunk * ws->Kp);
  ws->cand_idx = a.take<int32_t>((size_t)ws->chunk * ws->Kp);
  ws->exact_scores = a.take<float>((size_t)ws->chunk * ws->Kp);
  ws->flags = a.take<int32_t>((size_t)ws->chunk);
  memset(&ws->coarse, 0, sizeof(ws->coarse));
  if (tensor) coarse_plan(s, ws->chunk, a, &ws->coarse);
  ws->coarse.w1_img = ws->prep.w1_img;
  ws->coarse.w2_img = ws->prep.w2_img;
  ws->total = align_up(a.off, 256);
  if (base != nullptr && a.off > cap) {
    set_error("workspace too small: need %zu bytes, got %zu", ws->total, cap);
    return MOL_ERR_WORKSPACE;
  }
  return MOL_OK;
}

static int run_query_prologue(const mol_shape_t& s, const mol_weights_t& w, const float* queries,
                              const int64_t* user_ids, int B, float* pre, float* h, float* proj,
                              float* hq, float* qsub, float* gq, cudaStream_t st) {
  Dims D = dims_of(s);
  // query_embeddings_fns.py:191-197 : Linear(GLU(q))
  MOL_TRY(launch_linear(queries, w.q_glu_w, w.q_glu_b, pre, B, 2 * D.Hq, D.Dq, /*w_sn=*/1,
                        /*w_sk=*/2 * D.Hq, ACT_NONE, st));
  MOL_TRY(launch_glu(pre, h, B, D.Hq, s.query_nonlinearity, st));
  MOL_TRY(launch_linear(h, w.q_out_w, w.q_out_b, proj, B, D.Pq_proj * D.d, D.Hq, D.Hq, 1, ACT_NONE, st));
  MOL_TRY(launch_query_assemble(s, w, proj, user_ids, qsub, B, st));
  // similarity_fn.py:166-169 : query-only gating partial
  MOL_TRY(launch_linear(queries, w.gq_w1, w.gq_b1, hq, B, D.Hgq, D.Dq, D.Dq, 1, ACT_SILU, st));
  MOL_TRY(launch_linear(hq, w.gq_w2, nullptr, gq, B, D.L, D.Hgq, D.Hgq, 1, ACT_NONE, st));
  return MOL_OK;
}

// top-kk of each row of a (bc, n) device score matrix -> (out_scores, out_idx / out_ids), sorted
static int topk_of_matrix(const SearchWs& ws, const float* scores, int64_t n, int64_t ld, int bc, int kk,
                          float* out_scores, int32_t* out_idx, int64_t* out_ids, const int64_t* id_map,
                          const int32_t* flags, cudaStream_t st) {
  const int S = select_num_segments(n, bc, kk);
  const float* sel_scores = scores;
  const int32_t* sel_payload = nullptr;
  int64_t sel_n = n, sel_ld = ld;
  if (S > 1) {
    MOL_TRY(launch_select_segments(scores, n, ld, bc, S, kk, ws.seg_scores, ws.seg_idx, flags, st));
    sel_scores = ws.seg_scores;
    sel_payload = ws.seg_idx;
    sel_n = (int64_t)S * kk;
    sel_ld = sel_n;
  }
  return launch_select_final_i32(sel_scores, sel_payload, sel_n, sel_ld, bc, kk, out_scores, out_idx, out_ids,
                                 id_map, flags, st);
}

// Strikes the excluded ids from candidate lists: idx (bc, n) int32 item positions (< 0 = empty slot); an entry whose id
// (item_ids[position], or the position itself) is in the query's list invalid (bc, n0) becomes -1 - the rescoring gives
// it -inf and the selects skip it.  One block per query, the list in shared memory.
__global__ void __launch_bounds__(256)
exclude_candidates_kernel(int32_t* __restrict__ idx, int64_t n, const int32_t* __restrict__ cnt,
                          const int64_t* __restrict__ item_ids, const int64_t* __restrict__ invalid, int n0) {
  extern __shared__ int64_t ex_inv[];
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < n0; i += blockDim.x) ex_inv[i] = invalid[(int64_t)b * n0 + i];
  __syncthreads();
  int64_t used = n;
  if (cnt) used = cnt[b] < n ? cnt[b] : n;
  int32_t* row = idx + (int64_t)b * n;
  for (int64_t j = threadIdx.x; j < used; j += blockDim.x) {
    const int32_t p = row[j];
    if (p < 0) continue;
    const int64_t id = item_ids ? item_ids[p] : (int64_t)p;
    bool hit = false;
    for (int i = 0; i < n0; ++i) hit |= (ex_inv[i] == id);
    if (hit) row[j] = -1;
  }
}
static int exclude_candidates(int32_t* idx, int64_t n, const int32_t* cnt, int bc, const int64_t* item_ids,
                              const int64_t* invalid, int n0, cudaStream_t st) {
  if (bc == 0 || n0 == 0) return MOL_OK;
  const size_t smem = (size_t)n0 * sizeof(int64_t);
  if (smem > 48 * 1024)
    MOL_CUDA(cudaFuncSetAttribute(exclude_candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  exclude_candidates_kernel<<<bc, 256, smem, st>>>(idx, n, cnt, item_ids, invalid, n0);
  MOL_LAUNCH_CHECK();
  return MOL_OK;
}

// top-k of an exact (bc, N) score matrix; with an exclusion list: the top-(k + n0) and, from that, the first k entries
// whose id is not excluded (launch_select_valid: the masking of indexing/candidate_index.py:155-178)
static int exact_topk(const SearchWs& ws, const mol_index_t& ix, int bc, int k, const int64_t* invalid, float* o_scores,
                      int64_t* o_ids, const int32_t* flags, cudaStream_t st) {
  const int64_t N = ix.num_items;
  if (ws.n0 == 0)
    return topk_of_matrix(ws, ws.scores, N, N, bc, k, o_scores, nullptr, o_ids, ix.item_ids, flags, st);
  const int kx = k + ws.n0;
  MOL_TRY(topk_of_matrix(ws, ws.scores, N, N, bc, kx, ws.ex_scores, nullptr, ws.ex_ids, ix.item_ids, flags, st));
  return launch_select_valid(ws.ex_scores, ws.ex_ids, invalid, bc, kx, ws.n0, k, o_scores, o_ids, flags, st);
}

// Flagged queries are re-done exactly, in sub-chunks whose (rows, N) fp32 matrix fits the workspace (kernels
// exit immediately for unflagged rows).
static int exact_fallback(const mol_shape_t& s, const mol_weights_t& w, const mol_index_t& ix, const SearchWs& ws,
                          const Prepared& prep, const float* qsub, const float* gq, int bc, int k, const int64_t* invalid,
                          float* o_scores, int64_t* o_ids, cudaStream_t st) {
  Dims D = dims_of(s);
  const int64_t N = ix.num_items;
  for (int b1 = 0; b1 < bc; b1 += ws.chunk_fb) {
    const int nb = (bc - b1 < ws.chunk_fb) ? (bc - b1) : ws.chunk_fb;
    const int32_t* fl = ws.flags + b1;
    MOL_TRY(launch_exact_scores(s, w, ix, prep.w1t, prep.w2t, qsub + (size_t)b1 * D.Pq * D.d, gq + (size_t)b1 * D.L, nb,
                                nullptr, N, N, ws.scores, fl, st));
    MOL_TRY(exact_topk(ws, ix, nb, k, invalid ? invalid + (size_t)b1 * ws.n0 : nullptr, o_scores + (size_t)b1 * k,
                       o_ids + (size_t)b1 * k, fl, st));
  }
  return MOL_OK;
}

__global__ void stats_init_kernel(int32_t* stats, int filter, int tensor, int kp, int cap) {
  if (threadIdx.x < kNumStats) stats[threadIdx.x] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    stats[3] = filter;
    stats[4] = tensor;
    stats[6] = kp;
    stats[7] = cap;
  }
}

static int search_impl(const mol_shape_t& s, const mol_weights_t& w, const mol_index_t& ix,
                       const float* queries, const int64_t* user_ids, int B, int k, int mode,
                       float* out_scores, int64_t* out_ids, const SearchWs& ws_in, cudaStream_t st,
                       const int64_t* invalid_ids = nullptr) {
  Dims D = dims_of(s);
  const int64_t N = ix.num_items;
  const bool tensor = use_tensor(s, mode, B, N);
  SearchWs ws = ws_in;
  Prepared prep;
  {
    NvtxRange r("mol:prologue");
    stats_init_kernel<<<1, 32, 0, st>>>(ws.stats, ws.filter, tensor ? 1 : 0, ws.Kp, ws.cap);
    MOL_LAUNCH_CHECK();
    MOL_TRY(run_query_prologue(s, w, queries, user_ids, B, ws.pre, ws.h, ws.proj, ws.hq, ws.qsub, ws.gq, st));
    MOL_TRY(get_prepared(s, w, ws.prep, &prep, st));
    ws.coarse.w1_img = prep.w1_img;
    ws.coarse.w2_img = prep.w2_img;
    if (tensor && ws.filter && ws.samp_stride > 1)
      MOL_TRY(coarse_tile_maps(ws.map_sample, ws.map_main, (int)((N + 127) / 128), ws.samp_stride, (int)(ws.sample / 128), st));
  }

  for (int b0 = 0; b0 < B; b0 += ws.chunk) {
    const int bc = (B - b0 < ws.chunk) ? (B - b0) : ws.chunk;
    const float* qsub = ws.qsub + (size_t)b0 * D.Pq * D.d;
    const float* gq = ws.gq + (size_t)b0 * D.L;
    float* o_scores = out_scores + (size_t)b0 * k;
    int64_t* o_ids = out_ids + (size_t)b0 * k;
    const int64_t* invalid = (ws.n0 > 0 && invalid_ids) ? invalid_ids + (size_t)b0 * ws.n0 : nullptr;
    if (!tensor) {
      NvtxRange r("mol:exact");
      prof_begin(st);
      MOL_TRY(launch_exact_scores(s, w, ix, prep.w1t, prep.w2t, qsub, gq, bc, nullptr, N, N, ws.scores, nullptr, st));
      prof_end(st);
      MOL_TRY(exact_topk(ws, ix, bc, k, invalid, o_scores, o_ids, nullptr, st));
      continue;
    }
    const int kk = ws.Kp;
    const float* thr = nullptr;
    MOL_TRY(coarse_query_records(s, ws.coarse, qsub, gq, bc, prep.overflow, st));
    if (!ws.filter) {
      // matrix strategy: coarse scores of every pair -> top-K' per query
      {
        NvtxRange r("mol:coarse");
        prof_begin(st);
        MOL_TRY(coarse_scores(s, ix, ws.coarse, bc, ws.scores, st));
        prof_end(st);
      }
      NvtxRange r("mol:select");
      MOL_TRY(topk_of_matrix(ws, ws.scores, N, N, bc, kk, ws.cand_scores, ws.cand_idx, nullptr, nullptr, nullptr, st));
      if (invalid) MOL_TRY(exclude_candidates(ws.cand_idx, kk, nullptr, bc, ix.item_ids, invalid, ws.n0, st));
    } else {
      // filter strategy.  (1) threshold pass over a strided sample of the item tiles
      const int sample_tiles = (int)(ws.sample / 128);
      {
        NvtxRange r("mol:coarse_sample");
        CoarseOut o1{};
        o1.scores = ws.scores;
        o1.ld = ws.sample;
        o1.tile_begin = 0;
        o1.tile_end = sample_tiles;
        o1.tile_map = ws.samp_stride > 1 ? ws.map_sample : nullptr;
        prof_begin(st);
        MOL_TRY(coarse_run(s, ix, ws.coarse, bc, o1, st));
        prof_end(st);
        MOL_TRY(topk_of_matrix(ws, ws.scores, ws.sample, ws.sample, bc, ws.m, ws.samp_top, nullptr, nullptr, nullptr,
                               nullptr, st));
        thr = ws.samp_top + (ws.m - 1);  // stride m
        // (2) survivors of the sample
        MOL_CUDA(cudaMemsetAsync(ws.fcnt, 0, (size_t)bc * sizeof(int32_t), st));
        MOL_CUDA(cudaMemsetAsync(ws.fidx, 0xFF, (size_t)bc * ws.cap * sizeof(int32_t), st));
        MOL_TRY(coarse_filter_matrix(ws.scores, ws.sample, ws.sample, bc, thr, ws.m, ws.fcnt, ws.fscores, ws.fidx,
                                     ws.cap, ws.samp_stride, st));
      }
      {
        // ... then the main pass over every other tile, the filter fused into the scoring kernel's epilogue
        NvtxRange r("mol:coarse_main");
        CoarseOut o2{};
        o2.thr = thr;
        o2.thr_stride = ws.m;
        o2.cand_cnt = ws.fcnt;
        o2.cand_scores = ws.fscores;
        o2.cand_idx = ws.fidx;
        o2.cand_cap = ws.cap;
        const int tiles = (int)((N + 127) / 128);
        if (ws.samp_stride > 1) {
          o2.tile_map = ws.map_main;
          o2.tile_begin = 0;
          o2.tile_end = tiles - sample_tiles;
        } else {
          o2.tile_begin = sample_tiles;
          o2.tile_end = tiles;
        }
        prof_begin(st);
        MOL_TRY(coarse_run(s, ix, ws.coarse, bc, o2, st));
        prof_end(st);
      }
      // (3) the K' best survivors per query (excluded ids struck first: they never take one of the K' places)
      NvtxRange r("mol:select");
      if (invalid) MOL_TRY(exclude_candidates(ws.fidx, ws.cap, ws.fcnt, bc, ix.item_ids, invalid, ws.n0, st));
      MOL_TRY(launch_select_final_i32(ws.fscores, ws.fidx, ws.cap, ws.cap, bc, kk, ws.cand_scores, ws.cand_idx,
                                      nullptr, nullptr, nullptr, st));
    }
    // exact fp32 rescoring of the K' candidates -> final top-k (+ safety check and per-query exact fallback)
    NvtxRange r("mol:rescore");
    MOL_TRY(launch_exact_scores(s, w, ix, prep.w1t, prep.w2t, qsub, gq, bc, ws.cand_idx, kk, kk, ws.exact_scores,
                                nullptr, st));
    MOL_TRY(launch_select_final_i32(ws.exact_scores, ws.cand_idx, kk, kk, bc, k, o_scores, nullptr, o_ids,
                                    ix.item_ids, nullptr, st));
    if (kk < N) {
      MOL_TRY(coarse_safety_flags(ws.cand_scores, ws.exact_scores, o_scores, bc, kk, k, ws.coarse.overflow,
                                  ix.half_overflow, ws.filter ? ws.fcnt : nullptr, thr, ws.m, ws.cap, ws.flags,
                                  ws.stats, 0, st));
      if (ws.filter) {
        {
          const char* e = getenv("MOL_B200_FORCE_SECOND_CHANCE");  // test hook: send every query through the second chance
          if (e && atoi(e) != 0) MOL_CUDA(cudaMemsetAsync(ws.flags, 1, (size_t)bc * sizeof(int32_t), st));
        }
        // second chance for the refused queries, still without touching the corpus again: rescore EVERY survivor (<= cap per
        // query, ~0.15 ms per query instead of the 1.8 ms of a full exact pass); the items outside that set are bounded by
        // the filter threshold, usually far below the K'-th coarse score that the first test had to use.  Dense score
        // distributions (trained models, clustered corpora) that make the first test refuse mostly pass here.  Idle
        // launches when no query is flagged.
        NvtxRange r2("mol:second_chance");
        MOL_TRY(launch_exact_scores(s, w, ix, prep.w1t, prep.w2t, qsub, gq, bc, ws.fidx, ws.cap, ws.cap, ws.fexact, ws.flags,
                                    st));
        MOL_TRY(launch_select_final_i32(ws.fexact, ws.fidx, ws.cap, ws.cap, bc, k, o_scores, nullptr, o_ids, ix.item_ids,
                                        ws.flags, st));
        MOL_TRY(coarse_safety_flags(ws.fscores, ws.fexact, o_scores, bc, ws.cap, k, ws.coarse.overflow, ix.half_overflow,
                                    ws.fcnt, thr, ws.m, ws.cap, ws.flags, ws.stats, 1, st));
      }
      MOL_TRY(exact_fallback(s, w, ix, ws, prep, qsub, gq, bc, k, invalid, o_scores, o_ids, st));
    }
  }
  return MOL_OK;
}

}  // namespace mol

using namespace mol;

extern "C" {

const char* mol_version(void) {
  static const char* const v = [] {  // thread-safe one-time initialisation
    static char buf[192];
    snprintf(buf, sizeof(buf), "rails_b200 0.1 (sm_100a) [%s]", coarse_build_knobs());
    return buf;
  }();
  return v;
}
const char* mol_last_error(void) { return g_err; }
int64_t mol_launch_count(void) { return g_launches.load(); }
void mol_launch_count_reset(void) { g_launches.store(0); }

void mol_profile_enable(int32_t on) {
  g_prof_on = on != 0;
  g_prof_used = 0;
}

int mol_profile_collect(double* total_ms, int32_t* launches) {
  MOL_CHECK_ARG(total_ms && launches, "NULL output");
  double t = 0.0;
  for (size_t i = 0; i < g_prof_used; ++i) {
    MOL_CUDA(cudaEventSynchronize(g_prof_events[i].second));
    float ms = 0.f;
    MOL_CUDA(cudaEventElapsedTime(&ms, g_prof_events[i].first, g_prof_events[i].second));
    t += ms;
  }
  *total_ms = t;
  *launches = (int32_t)g_prof_used;
  g_prof_used = 0;
  return MOL_OK;
}

int mol_shape_check(const mol_shape_t* shape, int32_t* tensor_ok) {
  MOL_TRY(check_shape(shape));
  if (tensor_ok) *tensor_ok = coarse_supported(*shape) ? 1 : 0;
  return MOL_OK;
}

int mol_weights_prepared_bytes(const mol_shape_t* shape, size_t* bytes) {
  MOL_TRY(check_shape(shape));
  MOL_CHECK_ARG(bytes, "bytes is NULL");
  Prepared p;
  plan_prepared(*shape, nullptr, 0, &p);
  *bytes = p.total;
  return MOL_OK;
}

int mol_weights_prepare(const mol_shape_t* shape, const mol_weights_t* w, void* blob, size_t blob_bytes,
                        mol_stream_t stream) {
  MOL_TRY(check_shape(shape));
  MOL_TRY(check_weights(shape, w));
  MOL_CHECK_ARG(blob && (reinterpret_cast<uintptr_t>(blob) & 255) == 0, "prepared blob must be 256-byte aligned");
  Prepared p;
  plan_prepared(*shape, blob, blob_bytes, &p);
  MOL_CHECK_ARG(p.total <= blob_bytes, "prepared blob too small: need %zu, got %zu", p.total, blob_bytes);
  return run_prepare(*shape, *w, p, static_cast<cudaStream_t>(stream));
}

int mol_search_stats(const void* workspace, int32_t* host_stats, mol_stream_t stream) {
  MOL_CHECK_ARG(workspace && host_stats, "NULL buffer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MOL_CUDA(cudaMemcpyAsync(host_stats, workspace, kNumStats * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  MOL_CUDA(cudaStreamSynchronize(st));
  return MOL_OK;
}

int mol_index_bytes(const mol_shape_t* shape, int64_t num_items, size_t* bytes) {
  MOL_TRY(check_shape(shape));
  MOL_CHECK_ARG(num_items >= 0 && bytes, "bad arguments");
  MOL_CHECK_ARG(num_items < (1ll << 31) - 256, "num_items per shard must fit int32");
  *bytes = index_layout(dims_of(*shape), num_items).total;
  return MOL_OK;
}

int mol_index_layout(const mol_shape_t* shape, int64_t num_items, const float* raw_items,
                     const int64_t* item_ids, void* blob, size_t blob_bytes, mol_index_t* index) {
  MOL_TRY(check_shape(shape));
  MOL_CHECK_ARG(index && blob, "index/blob is NULL");
  MOL_CHECK_ARG(num_items >= 0 && num_items < (1ll << 31) - 256, "bad num_items");
  IndexLayout l = index_layout(dims_of(*shape), num_items);
  MOL_CHECK_ARG(blob_bytes >= l.total, "index blob too small: need %zu, got %zu", l.total, blob_bytes);
  MOL_CHECK_ARG((reinterpret_cast<uintptr_t>(blob) & 1023) == 0, "index blob must be 1024-byte aligned");
  char* p = static_cast<char*>(blob);
  index->num_items = num_items;
  index->raw_items = raw_items;
  index->item_ids = item_ids;
  index->xsub_f32 = reinterpret_cast<float*>(p + l.xsub_f32);
  index->gi_f32 = reinterpret_cast<float*>(p + l.gi_f32);
  index->xsub_half = reinterpret_cast<uint16_t*>(p + l.xsub_half);
  index->gi_half = reinterpret_cast<uint16_t*>(p + l.gi_half);
  index->half_overflow = reinterpret_cast<int32_t*>(p + l.half_overflow);
  return MOL_OK;
}

int mol_index_build_workspace_bytes(const mol_shape_t* shape, int64_t num_items, size_t* bytes) {
  MOL_TRY(check_shape(shape));
  MOL_CHECK_ARG(bytes, "bytes is NULL");
  Dims D = dims_of(*shape);
  int cols = D.Px * D.d > D.Hgi ? D.Px * D.d : D.Hgi;
  *bytes = align_up((size_t)num_items * cols * sizeof(float), 256) + 256;
  const size_t x3 = index_build_x3_workspace_bytes(*shape, num_items);  // (the tensor-core build needs less: the max)
  if (x3 > *bytes) *bytes = x3;
  return MOL_OK;
}

int mol_index_build(const mol_shape_t* shape, const mol_weights_t* w, const mol_index_t* index,
                    void* workspace, size_t workspace_bytes, mol_stream_t stream) {
  MOL_TRY(check_shape(shape));
  MOL_TRY(check_weights(shape, w));
  MOL_CHECK_ARG(index && index->raw_items, "index / raw_items is NULL");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Dims D = dims_of(*shape);
  const int64_t N = index->num_items;
  size_t need;
  MOL_TRY(mol_index_build_workspace_bytes(shape, N, &need));
  if (workspace_bytes < need || (N > 0 && !workspace)) {
    set_error("index build workspace too small: need %zu, got %zu", need, workspace_bytes);
    return MOL_ERR_WORKSPACE;
  }
  // tensor-core build: tf32 x 3 split GEMMs with fused l2-norm / silu / fp16-image epilogues (mol_linear_x3_sm100.cu)
  if (N > 0 && reinterpret_cast<uintptr_t>(workspace) % 256 == 0 && index_build_x3_supported(*shape, *w, *index))
    return index_build_x3(*shape, *w, *index, workspace, st);
  float* tmp = static_cast<float*>(workspace);
  const int64_t Np = pad128(N);
  // zero the fp16 pad rows (and everything else) so the tensor-core pass can read whole tiles
  MOL_CUDA(cudaMemsetAsync(index->xsub_half, 0, (size_t)Np * D.Px * D.d * sizeof(uint16_t), st));
  MOL_CUDA(cudaMemsetAsync(index->gi_half, 0, (size_t)Np * D.L * sizeof(uint16_t), st));
  MOL_CUDA(cudaMemsetAsync(index->half_overflow, 0, sizeof(int32_t), st));
  // item_embeddings_fns.py:165-182
  MOL_TRY(launch_linear(index->raw_items, w->x_w, w->x_b, tmp, N, D.Px * D.d, D.Dx, D.Dx, 1, ACT_NONE, st));
  MOL_TRY(launch_l2norm_groups(tmp, index->xsub_f32, reinterpret_cast<__half*>(index->xsub_half),
                               N, D.Px, D.d, shape->eps, st));
  // similarity_fn.py:170-171
  MOL_TRY(launch_linear(index->raw_items, w->gi_w1, w->gi_b1, tmp, N, D.Hgi, D.Dx, D.Dx, 1, ACT_SILU, st));
  MOL_TRY(launch_linear(tmp, w->gi_w2, nullptr, index->gi_f32, N, D.L, D.Hgi, D.Hgi, 1, ACT_NONE, st));
  if (coarse_supported(*shape)) {
    MOL_TRY(coarse_gi_image(*shape, index->gi_f32, index->gi_half, N, index->half_overflow, st));
  }
  return MOL_OK;
}

int mol_search_workspace_bytes(const mol_shape_t* shape, int64_t num_items, int32_t B, int32_t k,
                               int32_t mode, size_t* bytes) {
  MOL_TRY(check_shape(shape));
  MOL_CHECK_ARG(bytes && B >= 0 && k >= 0 && num_items >= 0, "bad arguments");
  MOL_CHECK_ARG(mode != MOL_MODE_TENSOR || coarse_supported(*shape), "shape not supported by the tensor-core path");
  SearchWs ws;
  MOL_TRY(plan_search(*shape, num_items, B > 0 ? B : 1, k > 0 ? k : 1, mode, nullptr, 0, &ws));
  *bytes = ws.total + 256;
  return MOL_OK;
}

static int check_search_args(const mol_shape_t* shape, const mol_weights_t* w, const mol_index_t* index,
                             const void* queries, const void* user_ids, int32_t B, int32_t k,
                             int32_t mode, const void* out_scores, const void* out_ids) {
  MOL_TRY(check_shape(shape));
  MOL_TRY(check_weights(shape, w));
  MOL_CHECK_ARG(index && index->xsub_f32 && index->gi_f32, "index not laid out");
  MOL_CHECK_ARG(B >= 0, "negative batch");
  MOL_CHECK_ARG(k >= 1, "k must be >= 1 (got %d)", k);
  if (k > index->num_items) {
    set_error("selected index k out of range (k=%d > %lld items)", k, (long long)index->num_items);
    return MOL_ERR_RANGE;
  }
  MOL_CHECK_ARG(k <= MOL_MAX_K, "k=%d exceeds MOL_MAX_K=%d", k, MOL_MAX_K);
  MOL_CHECK_ARG(B == 0 || (queries && out_scores && out_ids), "NULL query/output buffer");
  MOL_CHECK_ARG(shape->num_uid_tables == 0 || B == 0 || user_ids, "user_ids required when uid embeddings are configured");
  MOL_CHECK_ARG(mode == MOL_MODE_AUTO || mode == MOL_MODE_EXACT || mode == MOL_MODE_TENSOR, "bad mode %d", mode);
  MOL_CHECK_ARG(mode != MOL_MODE_TENSOR || coarse_supported(*shape), "shape not supported by the tensor-core path");
  return MOL_OK;
}

int mol_search(const mol_shape_t* shape, const mol_weights_t* w, const mol_index_t* index,
               const float* queries, const int64_t* user_ids, int32_t B, int32_t k, int32_t sorted,
               int32_t mode, float* out_scores, int64_t* out_ids, void* workspace,
               size_t workspace_bytes, mol_stream_t stream) {
  (void)sorted;  // results are always sorted; a sorted list is a valid unsorted answer
  MOL_TRY(check_search_args(shape, w, index, queries, user_ids, B, k, mode, out_scores, out_ids));
  if (B == 0) return MOL_OK;
  MOL_CHECK_ARG(workspace, "workspace is NULL");
  SearchWs ws;
  MOL_TRY(plan_search(*shape, index->num_items, B, k, mode, workspace, workspace_bytes, &ws));
  return search_impl(*shape, *w, *index, queries, user_ids, B, k, mode, out_scores, out_ids, ws,
                     static_cast<cudaStream_t>(stream));
}

int mol_search_excluding_workspace_bytes(const mol_shape_t* shape, int64_t num_items, int32_t B, int32_t k,
                                         int32_t n_invalid, int32_t mode, size_t* bytes) {
  MOL_TRY(check_shape(shape));
  MOL_CHECK_ARG(bytes && B >= 0 && k >= 0 && num_items >= 0 && n_invalid >= 0, "bad arguments");
  MOL_CHECK_ARG(mode != MOL_MODE_TENSOR || coarse_supported(*shape), "shape not supported by the tensor-core path");
  SearchWs ws;
  MOL_TRY(plan_search(*shape, num_items, B > 0 ? B : 1, k > 0 ? k : 1, mode, nullptr, 0, &ws, n_invalid));
  *bytes = ws.total + 256;
  return MOL_OK;
}

int mol_search_excluding(const mol_shape_t* shape, const mol_weights_t* w, const mol_index_t* index,
                         const float* queries, const int64_t* user_ids, int32_t B, int32_t k, int32_t sorted,
                         int32_t mode, const int64_t* invalid_ids, int32_t n_invalid, float* out_scores,
                         int64_t* out_ids, void* workspace, size_t workspace_bytes, mol_stream_t stream) {
  (void)sorted;
  MOL_TRY(check_search_args(shape, w, index, queries, user_ids, B, k, mode, out_scores, out_ids));
  MOL_CHECK_ARG(n_invalid >= 0 && (n_invalid == 0 || invalid_ids || B == 0), "exclusion list missing");
  MOL_CHECK_ARG((int64_t)k + n_invalid <= MOL_MAX_K, "k + n_invalid = %lld exceeds MOL_MAX_K=%d", (long long)k + n_invalid,
                MOL_MAX_K);
  if ((int64_t)k + n_invalid > index->num_items) {
    set_error("selected index k out of range (k + n_invalid = %lld > %lld items)", (long long)k + n_invalid,
              (long long)index->num_items);
    return MOL_ERR_RANGE;
  }
  if (B == 0) return MOL_OK;
  MOL_CHECK_ARG(workspace, "workspace is NULL");
  SearchWs ws;
  MOL_TRY(plan_search(*shape, index->num_items, B, k, mode, workspace, workspace_bytes, &ws, n_invalid));
  return search_impl(*shape, *w, *index, queries, user_ids, B, k, mode, out_scores, out_ids, ws,
                     static_cast<cudaStream_t>(stream), n_invalid > 0 ? invalid_ids : nullptr);
}

int mol_search_host(const mol_shape_t* shape, const mol_weights_t* w, const mol_index_t* index,
                    const float* host_queries, const int64_t* host_user_ids, int32_t B, int32_t k,
                    int32_t sorted, int32_t mode, float* host_out_scores, int64_t* host_out_ids,
                    void* workspace, size_t workspace_bytes, mol_stream_t stream) {
  (void)sorted;
  MOL_TRY(check_search_args(shape, w, index, host_queries, host_user_ids, B, k, mode, host_out_scores,
                            host_out_ids));
  if (B == 0) return MOL_OK;
  MOL_CHECK_ARG(workspace, "workspace is NULL");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SearchWs ws;
  MOL_TRY(plan_search(*shape, index->num_items, B, k, mode, workspace, workspace_bytes, &ws));
  Dims D = dims_of(*shape);
  MOL_CUDA(cudaMemcpyAsync(ws.stage_q, host_queries, (size_t)B * D.Dq * sizeof(float),
                           cudaMemcpyHostToDevice, st));
  const int64_t* uid = nullptr;
  if (D.u > 0) {
    MOL_CUDA(cudaMemcpyAsync(ws.stage_uid, host_user_ids, (size_t)B * sizeof(int64_t),
                             cudaMemcpyHostToDevice, st));
    uid = ws.stage_uid;
  }
  MOL_TRY(search_impl(*shape, *w, *index, ws.stage_q, uid, B, k, mode, ws.stage_out_scores,
                      ws.stage_out_ids, ws, st));
  MOL_CUDA(cudaMemcpyAsync(host_out_scores, ws.stage_out_scores, (size_t)B * k * sizeof(float),
                           cudaMemcpyDeviceToHost, st));
  MOL_CUDA(cudaMemcpyAsync(host_out_ids, ws.stage_out_ids, (size_t)B * k * sizeof(int64_t),
                           cudaMemcpyDeviceToHost, st));
  MOL_CUDA(cudaStreamSynchronize(st));
  return MOL_OK;
}

int mol_score_all(const mol_shape_t* shape, const mol_weights_t* w, const mol_index_t* index,
                  const float* queries, const int64_t* user_ids, int32_t B, float* out_scores,
                  void* workspace, size_t workspace_bytes, mol_stream_t stream) {
  MOL_TRY(check_shape(shape));
  MOL_TRY(check_weights(shape, w));
  MOL_CHECK_ARG(index && index->xsub_f32 && index->gi_f32, "index not laid out");
  MOL_CHECK_ARG(B >= 0, "negative batch");
  if (B == 0 || index->num_items == 0) return MOL_OK;
  MOL_CHECK_ARG(queries && out_scores && workspace, "NULL buffer");
  MOL_CHECK_ARG(shape->num_uid_tables == 0 || user_ids, "user_ids required when uid embeddings are configured");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SearchWs ws;
  MOL_TRY(plan_search(*shape, /*N=*/1, B, 1, MOL_MODE_EXACT, workspace, workspace_bytes, &ws));
  Dims D = dims_of(*shape);
  MOL_TRY(run_query_prologue(*shape, *w, queries, user_ids, B, ws.pre, ws.h, ws.proj, ws.hq, ws.qsub,
                             ws.gq, st));
  Prepared prep;
  MOL_TRY(get_prepared(*shape, *w, ws.prep, &prep, st));
  return launch_exact_scores(*shape, *w, *index, prep.w1t, prep.w2t, ws.qsub, ws.gq, B, nullptr,
                             index->num_items, index->num_items, out_scores, nullptr, st);
}

int mol_score_all_coarse(const mol_shape_t* shape, const mol_weights_t* w, const mol_index_t* index,
                         const float* queries, const int64_t* user_ids, int32_t B, float* out_scores,
                         void* workspace, size_t workspace_bytes, mol_stream_t stream) {
  MOL_TRY(check_shape(shape));
  MOL_TRY(check_weights(shape, w));
  MOL_CHECK_ARG(coarse_supported(*shape), "shape not supported by the tensor-core path");
  MOL_CHECK_ARG(index && index->xsub_half && index->gi_half, "index not laid out");
  MOL_CHECK_ARG(B >= 0, "negative batch");
  if (B == 0 || index->num_items == 0) return MOL_OK;
  MOL_CHECK_ARG(queries && out_scores && workspace, "NULL buffer");
  MOL_CHECK_ARG(shape->num_uid_tables == 0 || user_ids, "user_ids required when uid embeddings are configured");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SearchWs ws;
  MOL_TRY(plan_search(*shape, /*N=*/1, B, 1, MOL_MODE_TENSOR, workspace, workspace_bytes, &ws));
  MOL_TRY(run_query_prologue(*shape, *w, queries, user_ids, B, ws.pre, ws.h, ws.proj, ws.hq, ws.qsub,
                             ws.gq, st));
  Prepared prep;
  MOL_TRY(get_prepared(*shape, *w, ws.prep, &prep, st));
  ws.coarse.w1_img = prep.w1_img;
  ws.coarse.w2_img = prep.w2_img;
  MOL_TRY(coarse_query_records(*shape, ws.coarse, ws.qsub, ws.gq, B, prep.overflow, st));
  return coarse_scores(*shape, *index, ws.coarse, B, out_scores, st);
}

}  // extern "C"
namespace mol {
// ---- MoLAvgTopK (rails/indexing/mol_top_k.py:296-429): dot-product prefilter on group-averaged embeddings, then
// exact MoL on the avg_top_k survivors -------------------------------------------
__global__ void avg_groups_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t rows, int groups,
                                  int d, float scale) {  // out[r, :] = scale * sum_g in[r, g, :]
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * d) return;
  const int64_t r = i / d;
  const int c = (int)(i % d);
  float s = 0.f;
  for (int g = 0; g < groups; ++g) s += in[(r * groups + g) * d + c];
  out[i] = s * scale;
}

struct AvgWs {
  int32_t* stats;  // 8 x int32 at offset 0 (mol_search_stats): counters of the streaming prefilter (mol_dotfilter.cuh)
  float *pre, *h, *proj, *hq, *qsub, *gq, *w1t, *w2t, *qsum, *scores, *seg_scores, *exact;
  int32_t *seg_idx, *cand_idx;
  float* cand_scores;
  int rows;
  int filter;      // the streaming dot-product top-k serves the prefilter
  DotTopkPlan dp;
  size_t total;
};

static int plan_avg(const mol_shape_t& s, int64_t N, int B, int k, int avg_top_k, void* base, size_t cap, AvgWs* ws) {
  Dims D = dims_of(s);
  Arena a(base, cap);
  ws->stats = a.take<int32_t>(kNumStats);
  ws->pre = a.take<float>((size_t)B * 2 * D.Hq);
  ws->h = a.take<float>((size_t)B * D.Hq);
  ws->proj = a.take<float>((size_t)B * D.Pq_proj * D.d);
  ws->hq = a.take<float>((size_t)B * D.Hgq);
  ws->qsub = a.take<float>((size_t)B * D.Pq * D.d);
  ws->gq = a.take<float>((size_t)B * D.L);
  ws->w1t = a.take<float>((size_t)D.L * D.H);
  ws->w2t = a.take<float>((size_t)D.L * D.H);
  ws->qsum = a.take<float>((size_t)B * D.d);
  const int64_t n = N > 0 ? N : 1;
  int64_t rows = (int64_t)score_matrix_budget() / (int64_t)(sizeof(float) * (size_t)n);
  if (rows < 1) rows = 1;
  if (rows > B) rows = B > 0 ? B : 1;
  ws->rows = (int)rows;
  ws->scores = a.take<float>((size_t)rows * (size_t)n);
  const size_t seg = (size_t)select_streamed_slots(n, rows, avg_top_k) * (size_t)avg_top_k;
  ws->seg_scores = a.take<float>(seg);
  ws->seg_idx = a.take<int32_t>(seg);
  ws->cand_scores = a.take<float>((size_t)B * avg_top_k);
  ws->cand_idx = a.take<int32_t>((size_t)B * avg_top_k);
  ws->exact = a.take<float>((size_t)B * avg_top_k);
  (void)k;
  ws->filter = dot_topk_eligible(N, B, D.d, avg_top_k) ? 1 : 0;
  if (ws->filter) dot_topk_plan(a, N, B, D.d, avg_top_k, ws->scores, ws->rows, &ws->dp);
  ws->total = align_up(a.off, 256);
  if (base != nullptr && a.off > cap) {
    set_error("workspace too small: need %zu bytes, got %zu", ws->total, cap);
    return MOL_ERR_WORKSPACE;
  }
  return MOL_OK;
}

}  // namespace mol
using namespace mol;
extern "C" {

int mol_index_avg_embeddings(const mol_shape_t* shape, const mol_index_t* index, float* out_avg, mol_stream_t stream) {
  MOL_TRY(check_shape(shape));
  MOL_CHECK_ARG(index && index->xsub_f32 && out_avg, "index not laid out / NULL output");
  Dims D = dims_of(*shape);
  const int64_t total = index->num_items * D.d;
  if (total == 0) return MOL_OK;
  avg_groups_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      index->xsub_f32, out_avg, index->num_items, D.Px, D.d, 1.0f / D.Px);
  MOL_LAUNCH_CHECK();
  return MOL_OK;
}

int mol_search_avg_workspace_bytes(const mol_shape_t* shape, int64_t num_items, int32_t B, int32_t k,
                                   int32_t avg_top_k, size_t* bytes) {
  MOL_TRY(check_shape(shape));
  MOL_CHECK_ARG(bytes && B >= 0 && k >= 1 && avg_top_k >= 1 && num_items >= 0, "bad arguments");
  AvgWs ws;
  MOL_TRY(plan_avg(*shape, num_items, B > 0 ? B : 1, k, avg_top_k, nullptr, 0, &ws));
  *bytes = ws.total + 256;
  return MOL_OK;
}

int mol_search_avg(const mol_shape_t* shape, const mol_weights_t* w, const mol_index_t* index, const float* avg_items,
                   const float* queries, const int64_t* user_ids, int32_t B, int32_t k, int32_t avg_top_k,
                   float* out_scores, int64_t* out_ids, void* workspace, size_t workspace_bytes,
                   mol_stream_t stream) {
  MOL_TRY(check_shape(shape));
  MOL_TRY(check_weights(shape, w));
  MOL_CHECK_ARG(index && index->xsub_f32 && index->gi_f32 && avg_items, "index not laid out / avg embeddings missing");
  MOL_CHECK_ARG(B >= 0 && k >= 1 && avg_top_k >= 1 && avg_top_k <= MOL_MAX_K, "bad arguments");
  MOL_CHECK_ARG(k <= avg_top_k, "avg_top_k (%d) must be larger than k (%d)", avg_top_k, k);
  const int64_t N = index->num_items;
  if (avg_top_k > N) {
    set_error("selected index k out of range (avg_top_k=%d > %lld items)", avg_top_k, (long long)N);
    return MOL_ERR_RANGE;
  }
  if (B == 0) return MOL_OK;
  MOL_CHECK_ARG(queries && out_scores && out_ids && workspace, "NULL buffer");
  MOL_CHECK_ARG(shape->num_uid_tables == 0 || user_ids, "user_ids required when uid embeddings are configured");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Dims D = dims_of(*shape);
  AvgWs ws;
  MOL_TRY(plan_avg(*shape, N, B, k, avg_top_k, workspace, workspace_bytes, &ws));
  MOL_TRY(run_query_prologue(*shape, *w, queries, user_ids, B, ws.pre, ws.h, ws.proj, ws.hq, ws.qsub, ws.gq, st));
  MOL_TRY(launch_transpose(w->qi_w1, ws.w1t, D.H, D.L, st));
  MOL_TRY(launch_transpose(w->qi_w2, ws.w2t, D.L, D.H, st));
  {  // mol_top_k.py:352-356: mol_query_embeddings.sum(1)
    const int64_t total = (int64_t)B * D.d;
    avg_groups_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ws.qsub, ws.qsum, B, D.Pq, D.d, 1.0f);
    MOL_LAUNCH_CHECK();
  }
  MOL_CUDA(cudaMemsetAsync(ws.stats, 0, kNumStats * sizeof(int32_t), st));
  const bool stream_prefilter = ws.filter && dot_topk_aligned(avg_items, D.d, 0, ws.qsum, D.d);
  if (stream_prefilter)  // |avg_items[x]| <= 1: a mean of l2-normalised sub-embeddings
    MOL_TRY(dot_topk_run(ws.dp, avg_items, D.d, 0, D.d, nullptr, 1.0f, ws.qsum, D.d, ws.cand_scores, ws.cand_idx, nullptr,
                         nullptr, ws.stats, st));
  for (int b0 = 0; b0 < B && !stream_prefilter; b0 += ws.rows) {
    const int nb = (B - b0 < ws.rows) ? (B - b0) : ws.rows;
    // avg_sim_values = q_sum . avg_items^T ; top avg_top_k positions per query (:352-360)
    MOL_TRY(launch_linear(ws.qsum + (size_t)b0 * D.d, avg_items, nullptr, ws.scores, nb, (int)N, D.d, D.d, 1, ACT_NONE, st));
    const int S = select_num_segments_streamed(N, nb, avg_top_k);
    const float* sel = ws.scores;
    const int32_t* pay = nullptr;
    int64_t sn = N, sld = N;
    if (S > 1) {
      MOL_TRY(launch_select_segments(ws.scores, N, N, nb, S, avg_top_k, ws.seg_scores, ws.seg_idx, nullptr, st));
      sel = ws.seg_scores;
      pay = ws.seg_idx;
      sn = (int64_t)S * avg_top_k;
      sld = sn;
    }
    MOL_TRY(launch_select_final_i32(sel, pay, sn, sld, nb, avg_top_k, ws.cand_scores + (size_t)b0 * avg_top_k,
                                    ws.cand_idx + (size_t)b0 * avg_top_k, nullptr, nullptr, nullptr, st));
  }
  // exact MoL on the survivors (:368-373), final top-k and id map (:374-385)
  MOL_TRY(launch_exact_scores(*shape, *w, *index, ws.w1t, ws.w2t, ws.qsub, ws.gq, B, ws.cand_idx, avg_top_k, avg_top_k,
                              ws.exact, nullptr, st));
  return launch_select_final_i32(ws.exact, ws.cand_idx, avg_top_k, avg_top_k, B, k, out_scores, nullptr, out_ids,
                                 index->item_ids, nullptr, st);
}

}  // extern "C"
namespace mol {
// ---- MoLNaiveTopK / MoLCombTopK (rails/indexing/mol_top_k.py:133-293, 432-551) ------
// Per (query b, query group n, item group m): the k_per_group items with the largest <Q_sub[b,n], X_sub[x,m]>; the
// union over the L groups (plus, for Comb, the avg_top_k items of the averaged-embedding prefilter) is sorted by
// position, scored with exact fp32 MoL, duplicates are masked to -32767 and ALL C candidates are returned sorted
// by score (the reference overwrites k with C, mol_top_k.py:256 / :518).
constexpr float kDupScore = -32767.0f;  // mol_top_k.py:280 / :542

// One block per query: gathers the C candidate positions (P_X pieces of P_Q * kpg from the per-item-group selections
// + the avg piece) and sorts them ascending in shared memory (torch.sort at mol_top_k.py:252 / :514).
__global__ void __launch_bounds__(256)
union_sort_kernel(const int32_t* __restrict__ group_idx, const int32_t* __restrict__ avg_idx, int B, int Px, int per_m,
                  int avg_top_k, int C, int C2, int32_t* __restrict__ out_sorted) {
  extern __shared__ int32_t us_smem[];
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < C2; i += blockDim.x) {
    int32_t v = 0x7fffffff;
    if (i < Px * per_m) {
      const int m = i / per_m, j = i - m * per_m;
      v = group_idx[((size_t)m * B + b) * per_m + j];
    } else if (i < C) {
      v = avg_idx[(size_t)b * avg_top_k + (i - Px * per_m)];
    }
    us_smem[i] = v;
  }
  __syncthreads();
  for (int size = 2; size <= C2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < C2 / 2; i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool up = (lo & size) == 0;
        const int32_t a = us_smem[lo], c = us_smem[hi];
        if ((a > c) == up) {
          us_smem[lo] = c;
          us_smem[hi] = a;
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < C; i += blockDim.x) out_sorted[(size_t)b * C + i] = us_smem[i];
}

// candidate_is_valid (mol_top_k.py:271-280 / :533-542): a candidate equal to its left neighbour is a duplicate
__global__ void mask_duplicates_kernel(const int32_t* __restrict__ sorted_idx, float* __restrict__ scores, int64_t total,
                                       int C) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  if (i % C != 0 && sorted_idx[i] == sorted_idx[i - 1]) scores[i] = kDupScore;
}

struct GroupsWs {
  int32_t* stats;  // 8 x int32 at offset 0 (mol_search_stats): counters of the streaming selections (mol_dotfilter.cuh)
  float *pre, *h, *proj, *hq, *qsub, *gq, *w1t, *w2t, *qavg, *scores, *seg_scores, *sel_scores, *exact, *dsel;
  int32_t *seg_idx, *group_idx, *avg_idx, *sorted_idx;
  int rows;  // rows of the (rows, N) dot-product matrix scored per launch (a multiple of P_Q)
  int filter_groups, filter_avg;  // the streaming dot-product top-k serves the per-group / the averaged selection
  DotTopkPlan dp_groups, dp_avg;
  size_t total;
};

static int plan_groups(const mol_shape_t& s, int64_t N, int B, int kpg, int avg_top_k, void* base, size_t cap,
                       GroupsWs* ws) {
  Dims D = dims_of(s);
  Arena a(base, cap);
  ws->stats = a.take<int32_t>(kNumStats);
  ws->pre = a.take<float>((size_t)B * 2 * D.Hq);
  ws->h = a.take<float>((size_t)B * D.Hq);
  ws->proj = a.take<float>((size_t)B * D.Pq_proj * D.d);
  ws->hq = a.take<float>((size_t)B * D.Hgq);
  ws->qsub = a.take<float>((size_t)B * D.Pq * D.d);
  ws->gq = a.take<float>((size_t)B * D.L);
  ws->w1t = a.take<float>((size_t)D.L * D.H);
  ws->w2t = a.take<float>((size_t)D.L * D.H);
  ws->qavg = a.take<float>((size_t)B * D.d);
  const int64_t n = N > 0 ? N : 1;
  int64_t rows = (int64_t)score_matrix_budget() / (int64_t)(sizeof(float) * (size_t)n);
  rows = rows / D.Pq * D.Pq;
  if (rows < D.Pq) rows = D.Pq;
  if (rows > (int64_t)B * D.Pq) rows = (int64_t)B * D.Pq;
  ws->rows = (int)rows;
  ws->scores = a.take<float>((size_t)rows * (size_t)n);
  const int kmax = kpg > avg_top_k ? kpg : avg_top_k;
  const size_t seg = (size_t)select_streamed_slots(n, rows, kmax) * (size_t)kmax;
  ws->seg_scores = a.take<float>(seg);
  ws->seg_idx = a.take<int32_t>(seg);
  ws->sel_scores = a.take<float>((size_t)rows * kmax);
  const size_t C = (size_t)D.L * kpg + avg_top_k;
  ws->group_idx = a.take<int32_t>((size_t)B * D.L * kpg);
  ws->avg_idx = a.take<int32_t>((size_t)B * (avg_top_k > 0 ? avg_top_k : 1));
  ws->sorted_idx = a.take<int32_t>((size_t)B * C);
  ws->exact = a.take<float>((size_t)B * C);
  ws->dsel = a.take<float>((size_t)B * (D.Pq * kpg > avg_top_k ? D.Pq * kpg : avg_top_k));
  ws->filter_groups = dot_topk_eligible(N, B * D.Pq, D.d, kpg) ? 1 : 0;
  if (ws->filter_groups) dot_topk_plan(a, N, B * D.Pq, D.d, kpg, ws->scores, ws->rows, &ws->dp_groups);
  ws->filter_avg = (avg_top_k > 0 && dot_topk_eligible(N, B, D.d, avg_top_k)) ? 1 : 0;
  if (ws->filter_avg) dot_topk_plan(a, N, B, D.d, avg_top_k, ws->scores, ws->rows, &ws->dp_avg);
  ws->total = align_up(a.off, 256);
  if (base != nullptr && a.off > cap) {
    set_error("workspace too small: need %zu bytes, got %zu", ws->total, cap);
    return MOL_ERR_WORKSPACE;
  }
  return MOL_OK;
}

// top-kk column positions (int32) of every row of a (nb, N) matrix
static int select_positions(const GroupsWs& ws, int64_t N, int nb, int kk, int32_t* out_idx, cudaStream_t st) {
  const int S = select_num_segments_streamed(N, nb, kk);
  const float* sel = ws.scores;
  const int32_t* pay = nullptr;
  int64_t sn = N;
  if (S > 1) {
    MOL_TRY(launch_select_segments(ws.scores, N, N, nb, S, kk, ws.seg_scores, ws.seg_idx, nullptr, st));
    sel = ws.seg_scores;
    pay = ws.seg_idx;
    sn = (int64_t)S * kk;
  }
  return launch_select_final_i32(sel, pay, sn, sn, nb, kk, ws.sel_scores, out_idx, nullptr, nullptr, nullptr, st);
}

}  // namespace mol
using namespace mol;
extern "C" {

int mol_search_groups_workspace_bytes(const mol_shape_t* shape, int64_t num_items, int32_t B, int32_t k_per_group,
                                      int32_t avg_top_k, size_t* bytes) {
  MOL_TRY(check_shape(shape));
  MOL_CHECK_ARG(bytes && B >= 0 && k_per_group >= 1 && avg_top_k >= 0 && num_items >= 0, "bad arguments");
  GroupsWs ws;
  MOL_TRY(plan_groups(*shape, num_items, B > 0 ? B : 1, k_per_group, avg_top_k, nullptr, 0, &ws));
  *bytes = ws.total + 256;
  return MOL_OK;
}

int mol_search_groups(const mol_shape_t* shape, const mol_weights_t* w, const mol_index_t* index,
                      const float* avg_items, const float* queries, const int64_t* user_ids, int32_t B,
                      int32_t k_per_group, int32_t avg_top_k, float* out_scores, int64_t* out_ids, void* workspace,
                      size_t workspace_bytes, mol_stream_t stream) {
  MOL_TRY(check_shape(shape));
  MOL_TRY(check_weights(shape, w));
  MOL_CHECK_ARG(index && index->xsub_f32 && index->gi_f32, "index not laid out");
  MOL_CHECK_ARG(B >= 0 && k_per_group >= 1 && avg_top_k >= 0, "bad arguments");
  MOL_CHECK_ARG(avg_top_k == 0 || avg_items, "avg embeddings missing (mol_index_avg_embeddings)");
  Dims D = dims_of(*shape);
  const int64_t N = index->num_items;
  const int64_t C64 = (int64_t)D.L * k_per_group + avg_top_k;
  MOL_CHECK_ARG(C64 <= MOL_MAX_K, "P_Q * P_X * k_per_group + avg_top_k = %lld exceeds MOL_MAX_K=%d", (long long)C64,
                MOL_MAX_K);
  if (k_per_group > N || avg_top_k > N) {
    set_error("selected index k out of range (k_per_group=%d / avg_top_k=%d > %lld items)", k_per_group, avg_top_k,
              (long long)N);
    return MOL_ERR_RANGE;
  }
  if (B == 0) return MOL_OK;
  MOL_CHECK_ARG(queries && out_scores && out_ids && workspace, "NULL buffer");
  MOL_CHECK_ARG(shape->num_uid_tables == 0 || user_ids, "user_ids required when uid embeddings are configured");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int C = (int)C64;
  GroupsWs ws;
  MOL_TRY(plan_groups(*shape, N, B, k_per_group, avg_top_k, workspace, workspace_bytes, &ws));
  MOL_TRY(run_query_prologue(*shape, *w, queries, user_ids, B, ws.pre, ws.h, ws.proj, ws.hq, ws.qsub, ws.gq, st));
  MOL_TRY(launch_transpose(w->qi_w1, ws.w1t, D.H, D.L, st));
  MOL_TRY(launch_transpose(w->qi_w2, ws.w2t, D.L, D.H, st));
  // per-group dot products and their top k_per_group positions (mol_top_k.py:239-249 / :501-511): for item group m,
  // rows (b, n) of Q_sub against X_sub[:, m, :] (row stride P_X * d inside the fp32 cache)
  const int per_m = D.Pq * k_per_group;
  const int qrows = ws.rows / D.Pq;  // queries per launch
  MOL_CUDA(cudaMemsetAsync(ws.stats, 0, kNumStats * sizeof(int32_t), st));
  // streaming selection (tcgen05 tf32 pass + threshold filter + fp32 rescoring, no (rows, N) matrix; X_sub rows have
  // norm <= 1) when the sizes allow it, else the materialised matrix + radix select
  const bool stream_groups = ws.filter_groups && dot_topk_aligned(index->xsub_f32, (int64_t)D.Px * D.d, 0, ws.qsub, D.d);
  for (int m = 0; m < D.Px; ++m) {
    if (stream_groups) {
      MOL_TRY(dot_topk_run(ws.dp_groups, index->xsub_f32, (int64_t)D.Px * D.d, m * D.d, D.d, nullptr, 1.0f, ws.qsub, D.d,
                           ws.dsel, ws.group_idx + (size_t)m * B * per_m, nullptr, nullptr, ws.stats, st));
      continue;
    }
    for (int b0 = 0; b0 < B; b0 += qrows) {
      const int nq = (B - b0 < qrows) ? (B - b0) : qrows;
      const int nb = nq * D.Pq;
      MOL_TRY(launch_linear(ws.qsub + (size_t)b0 * D.Pq * D.d, index->xsub_f32 + (size_t)m * D.d, nullptr, ws.scores, nb,
                            (int)N, D.d, (int64_t)D.Px * D.d, 1, ACT_NONE, st));
      MOL_TRY(select_positions(ws, N, nb, k_per_group, ws.group_idx + ((size_t)m * B + b0) * per_m, st));
    }
  }
  if (avg_top_k > 0) {  // MoLAvgTopK.topk_ids (:387-429): mean over the query groups . mean over the item groups
    const int64_t total = (int64_t)B * D.d;
    avg_groups_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ws.qsub, ws.qavg, B, D.Pq, D.d, 1.0f / D.Pq);
    MOL_LAUNCH_CHECK();
    const bool stream_avg = ws.filter_avg && dot_topk_aligned(avg_items, D.d, 0, ws.qavg, D.d);
    if (stream_avg)
      MOL_TRY(dot_topk_run(ws.dp_avg, avg_items, D.d, 0, D.d, nullptr, 1.0f, ws.qavg, D.d, ws.dsel, ws.avg_idx, nullptr,
                           nullptr, ws.stats, st));
    for (int b0 = 0; b0 < B && !stream_avg; b0 += ws.rows) {
      const int nb = (B - b0 < ws.rows) ? (B - b0) : ws.rows;
      MOL_TRY(launch_linear(ws.qavg + (size_t)b0 * D.d, avg_items, nullptr, ws.scores, nb, (int)N, D.d, D.d, 1, ACT_NONE, st));
      MOL_TRY(select_positions(ws, N, nb, avg_top_k, ws.avg_idx + (size_t)b0 * avg_top_k, st));
    }
  }
  int C2 = 1;
  while (C2 < C) C2 <<= 1;
  union_sort_kernel<<<B, 256, (size_t)C2 * sizeof(int32_t), st>>>(ws.group_idx, ws.avg_idx, B, D.Px, per_m, avg_top_k, C,
                                                                  C2, ws.sorted_idx);
  MOL_LAUNCH_CHECK();
  // exact MoL on the union (:257-270), duplicate mask (:271-280), all C candidates sorted by score + id map (:281-292)
  MOL_TRY(launch_exact_scores(*shape, *w, *index, ws.w1t, ws.w2t, ws.qsub, ws.gq, B, ws.sorted_idx, C, C, ws.exact,
                              nullptr, st));
  {
    const int64_t total = (int64_t)B * C;
    mask_duplicates_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ws.sorted_idx, ws.exact, total, C);
    MOL_LAUNCH_CHECK();
  }
  return launch_select_final_i32(ws.exact, ws.sorted_idx, C, C, B, C, out_scores, nullptr, out_ids, index->item_ids,
                                 nullptr, st);
}

int mol_query_prologue(const mol_shape_t* shape, const mol_weights_t* w, const float* queries,
                       const int64_t* user_ids, int32_t B, float* out_qsub, float* out_gq,
                       void* workspace, size_t workspace_bytes, mol_stream_t stream) {
  MOL_TRY(check_shape(shape));
  MOL_TRY(check_weights(shape, w));
  MOL_CHECK_ARG(B >= 0, "negative batch");
  if (B == 0) return MOL_OK;
  MOL_CHECK_ARG(queries && out_qsub && out_gq && workspace, "NULL buffer");
  MOL_CHECK_ARG(shape->num_uid_tables == 0 || user_ids, "user_ids required when uid embeddings are configured");
  SearchWs ws;
  MOL_TRY(plan_search(*shape, 1, B, 1, MOL_MODE_EXACT, workspace, workspace_bytes, &ws));
  return run_query_prologue(*shape, *w, queries, user_ids, B, ws.pre, ws.h, ws.proj, ws.hq, out_qsub,
                            out_gq, static_cast<cudaStream_t>(stream));
}

int mol_topk_workspace_bytes(int64_t n, int32_t B, int32_t k, size_t* bytes) {
  MOL_CHECK_ARG(bytes && n >= 0 && B >= 0 && k >= 1, "bad arguments");
  // slots >= B' * S(B') for every B' <= B: a caller that sizes for its largest chunk can run the smaller ones
  const size_t slots = (size_t)select_streamed_slots(n, B > 0 ? B : 1, k);
  *bytes = 2 * align_up(slots * k * sizeof(float), 256) + 512;
  return MOL_OK;
}

int mol_topk(const float* scores, int64_t n, int64_t ld, int32_t B, int32_t k, const int64_t* id_map,
             float* out_scores, int64_t* out_idx, void* workspace, size_t workspace_bytes,
             mol_stream_t stream) {
  MOL_CHECK_ARG(B >= 0 && k >= 1 && k <= MOL_MAX_K && n >= 0 && ld >= n, "bad arguments");
  MOL_CHECK_ARG(n < (1ll << 31) - 256, "n must fit int32");
  if (k > n) {
    set_error("selected index k out of range (k=%d > %lld columns)", k, (long long)n);
    return MOL_ERR_RANGE;
  }
  if (B == 0) return MOL_OK;
  MOL_CHECK_ARG(scores && out_scores && out_idx, "NULL buffer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int S = select_num_segments_streamed(n, B, k);
  if (S == 1)
    return launch_select_final_i32(scores, nullptr, n, ld, B, k, out_scores, nullptr, out_idx, id_map, nullptr, st);
  size_t need;
  MOL_TRY(mol_topk_workspace_bytes(n, B, k, &need));
  if (!workspace || workspace_bytes < need) {
    set_error("topk workspace too small: need %zu, got %zu", need, workspace_bytes);
    return MOL_ERR_WORKSPACE;
  }
  Arena a(workspace, workspace_bytes);
  float* ss = a.take<float>((size_t)B * S * k);
  int32_t* si = a.take<int32_t>((size_t)B * S * k);
  MOL_TRY(launch_select_segments(scores, n, ld, B, S, k, ss, si, nullptr, st));
  return launch_select_final_i32(ss, si, (int64_t)S * k, (int64_t)S * k, B, k, out_scores, nullptr, out_idx,
                                 id_map, nullptr, st);
}

int mol_merge_topk_workspace_bytes(int32_t R, int32_t B, int32_t k, size_t* bytes) {
  MOL_CHECK_ARG(bytes && R >= 1 && B >= 0 && k >= 1, "bad arguments");
  *bytes = align_up((size_t)R * B * k * sizeof(float), 256) + (size_t)R * B * k * sizeof(int64_t) + 512;
  return MOL_OK;
}

// parts are (R, B, k); the select kernel wants (B, R*k) rows -> handled by a strided view: we launch
// one merge per query row with ld = k and gather the R parts through a small transpose kernel.
__global__ void merge_gather_kernel(const float* __restrict__ ps, const int64_t* __restrict__ pi,
                                    float* __restrict__ os, int64_t* __restrict__ oi, int R, int B, int k) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = (int64_t)R * B * k;
  if (i >= total) return;
  int r = (int)(i / ((int64_t)B * k));
  int64_t rem = i % ((int64_t)B * k);
  int b = (int)(rem / k), j = (int)(rem % k);
  int64_t o = ((int64_t)b * R + r) * k + j;
  os[o] = ps[i];
  oi[o] = pi[i];
}

int mol_merge_topk(const float* part_scores, const int64_t* part_ids, int32_t R, int32_t B,
                   int32_t k, float* out_scores, int64_t* out_ids, void* workspace,
                   size_t workspace_bytes, mol_stream_t stream) {
  MOL_CHECK_ARG(R >= 1 && B >= 0 && k >= 1 && k <= MOL_MAX_K, "bad arguments");
  if (B == 0) return MOL_OK;
  MOL_CHECK_ARG(part_scores && part_ids && out_scores && out_ids, "NULL buffer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  size_t need = align_up((size_t)R * B * k * sizeof(float), 256) + (size_t)R * B * k * sizeof(int64_t) + 256;
  if (!workspace || workspace_bytes < need) {
    set_error("merge workspace too small: need %zu, got %zu", need, workspace_bytes);
    return MOL_ERR_WORKSPACE;
  }
  Arena a(workspace, workspace_bytes);
  float* gs = a.take<float>((size_t)R * B * k);
  int64_t* gi = a.take<int64_t>((size_t)R * B * k);
  int64_t total = (int64_t)R * B * k;
  merge_gather_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(part_scores, part_ids, gs, gi, R, B, k);
  MOL_LAUNCH_CHECK();
  return launch_select_final_i64(gs, gi, (int64_t)R * k, (int64_t)R * k, B, k, out_scores, out_ids, st);
}

}  // extern "C"

namespace mol {
// ---- packed partial top-k lists: the payload of the single all-gather of the multi-GPU paths -------------------------
struct __align__(16) PackedEntry {
  int64_t id;
  float score;
  int32_t valid;
};
static_assert(sizeof(PackedEntry) == MOL_PACKED_ENTRY_BYTES, "packed entry layout");

__global__ void pack_topk_kernel(const float* __restrict__ scores, const int64_t* __restrict__ ids, int B, int k_valid,
                                 int k, PackedEntry* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * k) return;
  const int b = (int)(i / k), j = (int)(i % k);
  PackedEntry e;
  if (j < k_valid) {
    e.id = ids[(int64_t)b * k_valid + j];
    e.score = scores[(int64_t)b * k_valid + j];
    e.valid = 1;
  } else {
    e.id = -1;
    e.score = -__int_as_float(0x7f800000);
    e.valid = 0;
  }
  out[i] = e;
}

// gathered (R, B, k) entries -> per query row (B, R*k): scores, payload = flat position in the row-major (B, R*k) id
// array (or -1 for a padding entry), ids
__global__ void unpack_gathered_kernel(const PackedEntry* __restrict__ g, int R, int B, int k, float* __restrict__ os,
                                       int32_t* __restrict__ op, int64_t* __restrict__ oi) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)R * B * k;
  if (i >= total) return;
  const int r = (int)(i / ((int64_t)B * k));
  const int64_t rem = i % ((int64_t)B * k);
  const int b = (int)(rem / k), j = (int)(rem % k);
  const int64_t o = ((int64_t)b * R + r) * k + j;
  const PackedEntry e = g[i];
  os[o] = e.valid ? e.score : -__int_as_float(0x7f800000);
  op[o] = e.valid ? (int32_t)o : -1;
  oi[o] = e.id;
}
}  // namespace mol

extern "C" {

int mol_pack_topk(const float* scores, const int64_t* ids, int32_t B, int32_t k_valid, int32_t k, void* out_packed,
                  mol_stream_t stream) {
  MOL_CHECK_ARG(B >= 0 && k >= 1 && k_valid >= 0 && k_valid <= k, "bad arguments");
  if (B == 0) return MOL_OK;
  MOL_CHECK_ARG(out_packed && (k_valid == 0 || (scores && ids)), "NULL buffer");
  const int64_t total = (int64_t)B * k;
  pack_topk_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      scores, ids, B, k_valid, k, static_cast<PackedEntry*>(out_packed));
  MOL_LAUNCH_CHECK();
  return MOL_OK;
}

int mol_merge_topk_packed_workspace_bytes(int32_t R, int32_t B, int32_t k, size_t* bytes) {
  MOL_CHECK_ARG(bytes && R >= 1 && B >= 0 && k >= 1, "bad arguments");
  const size_t n = (size_t)R * B * k;
  *bytes = align_up(n * sizeof(float), 256) + align_up(n * sizeof(int32_t), 256) + align_up(n * sizeof(int64_t), 256) + 512;
  return MOL_OK;
}

int mol_merge_topk_packed(const void* gathered, int32_t R, int32_t B, int32_t k, float* out_scores, int64_t* out_ids,
                          void* workspace, size_t workspace_bytes, mol_stream_t stream) {
  MOL_CHECK_ARG(R >= 1 && B >= 0 && k >= 1 && k <= MOL_MAX_K, "bad arguments");
  MOL_CHECK_ARG((int64_t)R * B * k < (1ll << 31), "R * B * k must fit int32");
  if (B == 0) return MOL_OK;
  MOL_CHECK_ARG(gathered && out_scores && out_ids, "NULL buffer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  size_t need;
  MOL_TRY(mol_merge_topk_packed_workspace_bytes(R, B, k, &need));
  if (!workspace || workspace_bytes < need) {
    set_error("merge workspace too small: need %zu, got %zu", need, workspace_bytes);
    return MOL_ERR_WORKSPACE;
  }
  Arena a(workspace, workspace_bytes);
  const int64_t total = (int64_t)R * B * k;
  float* gs = a.take<float>((size_t)total);
  int32_t* gp = a.take<int32_t>((size_t)total);
  int64_t* gi = a.take<int64_t>((size_t)total);
  unpack_gathered_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(static_cast<const PackedEntry*>(gathered), R, B,
                                                                          k, gs, gp, gi);
  MOL_LAUNCH_CHECK();
  // padding entries carry payload -1 (ignored by the select); ties: lower rank, then lower position, first
  return launch_select_final_i32(gs, gp, (int64_t)R * k, (int64_t)R * k, B, k, out_scores, nullptr, out_ids, gi, nullptr,
                                 st);
}

}  // extern "C"
