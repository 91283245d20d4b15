// Per-query top-k selection: 3-pass radix select (11/11/10 bits) + bitonic sort.
//
// Replaces torch.topk(all_logits, dim=1, k, sorted, largest=True) and the id gather
// `self._item_ids.squeeze(0)[top_k_indices]` of rails/indexing/mol_top_k.py:123-130.
// One CTA handles one (segment, query) pair; rows are cut into S segments so that few-query /
// large-corpus calls still fill the 148 SMs, and a second launch of the same kernel merges the S*k
// survivors.  Ordering: descending score, ties broken by the smaller payload (position / id).
#include <math_constants.h>

#include "common.cuh"

namespace mol {

constexpr int SEL_THREADS = 512;
constexpr int SEL_BINS = 2048;

__device__ __forceinline__ uint32_t order_key(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
constexpr uint32_t PAD_KEY = 0x007fffffu;  // order_key(-inf)
__device__ __forceinline__ float key_to_float(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}

template <typename PayloadT>
struct SelectParams {
  const float* vals;        // B rows
  const PayloadT* payload;  // same shape as vals, or nullptr => payload = column index
  int64_t n;                // columns per row
  int64_t ld;               // row stride (elements)
  int S;                    // segments per row
  int kk;                   // survivors per segment
  // outputs: FINAL -> (B, kk) sorted; else (B, S, kk) unsorted, padded with (-inf, -1)
  float* out_vals;
  PayloadT* out_payload;       // nullable when FINAL
  int64_t* out_ids;            // FINAL only, nullable: id_map ? id_map[payload] : payload
  const int64_t* id_map;       // nullable
  const int32_t* query_flags;  // nullable: skip rows whose flag == 0
};

template <typename PayloadT, bool FINAL>
__global__ void __launch_bounds__(SEL_THREADS) select_topk_kernel(SelectParams<PayloadT> P) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  __shared__ int hist[SEL_BINS];
  __shared__ int warp_tot[SEL_THREADS / 32];
  __shared__ int sh_bin, sh_rem, cnt_gt, cnt_eq;

  const int b = blockIdx.y, s = blockIdx.x, tid = threadIdx.x;
  if (P.query_flags && P.query_flags[b] == 0) return;
  const int kk = P.kk;
  int64_t seg = (P.n + P.S - 1) / P.S;
  const int64_t lo = (int64_t)s * seg;
  int64_t hi = lo + seg;
  if (hi > P.n) hi = P.n;
  const int64_t cnt = hi > lo ? hi - lo : 0;
  const float* v = P.vals + (int64_t)b * P.ld;
  const PayloadT* pl = P.payload ? P.payload + (int64_t)b * P.ld : nullptr;

  // FINAL: survivors go to shared memory for sorting; else straight to global.
  int sortP = 1;
  while (sortP < kk) sortP <<= 1;
  uint32_t* skey = reinterpret_cast<uint32_t*>(dyn_smem);
  PayloadT* spay = reinterpret_cast<PayloadT*>(dyn_smem + (((size_t)sortP * sizeof(uint32_t) + 15) & ~(size_t)15));
  float* gvals = FINAL ? nullptr : P.out_vals + ((int64_t)b * P.S + s) * kk;
  PayloadT* gpay = FINAL ? nullptr : P.out_payload + ((int64_t)b * P.S + s) * kk;

  auto emit = [&](int slot, uint32_t key, PayloadT p) {
    if (FINAL) {
      skey[slot] = key;
      spay[slot] = p;
    } else {
      gvals[slot] = key_to_float(key);
      gpay[slot] = p;
    }
  };
  auto valid_at = [&](int64_t i, uint32_t& key, PayloadT& p) -> bool {
    p = pl ? pl[i] : (PayloadT)i;
    key = order_key(v[i]);
    return p >= 0;
  };

  int n_out;  // number of real survivors
  if (cnt <= kk) {
    // take everything (invalid payloads become padding)
    for (int64_t i = tid; i < kk; i += SEL_THREADS) {
      uint32_t key = 0;
      PayloadT p = -1;
      bool ok = (i < cnt) && valid_at(lo + i, key, p);
      emit((int)i, ok ? key : PAD_KEY, ok ? p : (PayloadT)-1);
    }
    n_out = kk;
  } else {
    uint32_t prefix = 0, pmask = 0;
    int remaining = kk;
    const int shifts[3] = {21, 10, 0};
    const int nbits[3] = {11, 11, 10};
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
      const int shift = shifts[pass];
      const uint32_t bmask = (1u << nbits[pass]) - 1u;
      for (int i = tid; i < SEL_BINS; i += SEL_THREADS) hist[i] = 0;
      __syncthreads();
      for (int64_t i = lo + tid; i < hi; i += SEL_THREADS) {
        uint32_t key;
        PayloadT p;
        if (!valid_at(i, key, p)) key = PAD_KEY;
        if ((key & pmask) == prefix) atomicAdd(&hist[(key >> shift) & bmask], 1);
      }
      __syncthreads();
      // descending scan: thread t owns bins (2047-4t) .. (2047-4t-3)
      int c[4], tsum = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        c[j] = hist[SEL_BINS - 1 - (4 * tid + j)];
        tsum += c[j];
      }
      int incl = tsum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((tid & 31) >= o) incl += t;
      }
      if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
      __syncthreads();
      int base = 0;
      for (int w = 0; w < (tid >> 5); ++w) base += warp_tot[w];
      int excl = base + incl - tsum;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (excl < remaining && remaining <= excl + c[j]) {
          sh_bin = SEL_BINS - 1 - (4 * tid + j);
          sh_rem = remaining - excl;
        }
        excl += c[j];
      }
      __syncthreads();
      prefix |= ((uint32_t)sh_bin) << shift;
      pmask |= bmask << shift;
      remaining = sh_rem;
      __syncthreads();
    }
    // prefix = key of the kk-th largest; take all keys > prefix and `remaining` keys == prefix.
    if (tid == 0) {
      cnt_gt = 0;
      cnt_eq = 0;
    }
    __syncthreads();
    const int n_gt = kk - remaining;
    for (int64_t i = lo + tid; i < hi; i += SEL_THREADS) {
      uint32_t key;
      PayloadT p;
      if (!valid_at(i, key, p)) key = PAD_KEY;
      if (key > prefix) {
        int slot = atomicAdd(&cnt_gt, 1);
        emit(slot, key, p);
      } else if (key == prefix) {
        int e = atomicAdd(&cnt_eq, 1);
        if (e < remaining) emit(n_gt + e, key, p >= 0 ? p : (PayloadT)-1);
      }
    }
    n_out = kk;
  }
  if (!FINAL) return;

  // ---- bitonic sort of the survivors (descending key, ascending payload), padded to sortP
  for (int i = n_out + tid; i < sortP; i += SEL_THREADS) {
    skey[i] = PAD_KEY;
    spay[i] = (PayloadT)-1;
  }
  __syncthreads();
  auto before = [](uint32_t ka, PayloadT pa, uint32_t kb, PayloadT pb) -> bool {
    // true if a must come before b;  padding (payload < 0) goes last
    if ((pa < 0) != (pb < 0)) return pb < 0;
    if (ka != kb) return ka > kb;
    return pa < pb;
  };
  for (int size = 2; size <= sortP; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < sortP / 2; t += SEL_THREADS) {
        int i = 2 * t - (t & (stride - 1));
        int j = i + stride;
        bool asc_block = ((i & size) == 0);  // "ascending" here means our order (best first)
        uint32_t ka = skey[i], kb = skey[j];
        PayloadT pa = spay[i], pb = spay[j];
        bool ordered = before(ka, pa, kb, pb) || (ka == kb && pa == pb);
        if (ordered != asc_block) {
          skey[i] = kb;
          skey[j] = ka;
          spay[i] = pb;
          spay[j] = pa;
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < kk; i += SEL_THREADS) {
    PayloadT p = spay[i];
    int64_t o = (int64_t)b * kk + i;
    P.out_vals[o] = p >= 0 ? key_to_float(skey[i]) : -CUDART_INF_F;
    if (P.out_payload) P.out_payload[o] = p;
    if (P.out_ids) P.out_ids[o] = (p >= 0 && P.id_map) ? P.id_map[p] : (int64_t)p;
  }
}

template <typename PayloadT, bool FINAL>
static int launch_select_t(const SelectParams<PayloadT>& P, int B, cudaStream_t st) {
  if (B == 0 || P.kk == 0) return MOL_OK;
  MOL_CHECK_ARG(P.kk <= MOL_MAX_K, "top-k: k=%d exceeds MOL_MAX_K=%d", P.kk, MOL_MAX_K);
  int sortP = 1;
  while (sortP < P.kk) sortP <<= 1;
  size_t smem = FINAL ? (size_t)sortP * (sizeof(uint32_t) + sizeof(PayloadT)) + 16 : 0;
  auto kern = select_topk_kernel<PayloadT, FINAL>;
  if (smem > 48 * 1024)
    MOL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)P.S, (unsigned)B);
  kern<<<grid, SEL_THREADS, smem, st>>>(P);
  MOL_LAUNCH_CHECK();
  return MOL_OK;
}

int select_num_segments(int64_t n, int B, int kk) {
  // enough CTAs to fill the GPU, but segments no shorter than max(16k, 4*kk) elements
  int64_t min_seg = 16384;
  if (min_seg < 4 * (int64_t)kk) min_seg = 4 * (int64_t)kk;
  int64_t s_max = n / min_seg;
  if (s_max < 1) s_max = 1;
  int64_t want = (2 * 148 + B - 1) / B;
  int64_t S = want < s_max ? want : s_max;
  if (S < 1) S = 1;
  if (S > 256) S = 256;
  return (int)S;
}

// Segment count for selects over a MATERIALISED (rows, n) matrix that does not fit the L2 (MIPS, the Avg / Naive / Comb
// prefilters): segments of <= 32k elements (128 KB), so that the four passes of the radix select over a segment (three
// histograms + the emit) read HBM once and the L2 afterwards - ~600 resident CTAs x 128 KB stay below the 126 MB L2.
int select_num_segments_streamed(int64_t n, int B, int kk) {
  int64_t S = select_num_segments(n, B, kk);
  int64_t min_seg = 16384;
  if (min_seg < 4 * (int64_t)kk) min_seg = 4 * (int64_t)kk;
  int64_t s_max = n / min_seg;
  if (s_max < 1) s_max = 1;
  if (s_max > 256) s_max = 256;
  int64_t want = (n + 32767) / 32768;
  if (want > s_max) want = s_max;
  return (int)(want > S ? want : S);
}
// Upper bound of rows' * select_num_segments_streamed(n, rows', kk) over rows' <= rows (workspace sizing).
int64_t select_streamed_slots(int64_t n, int64_t rows, int kk) {
  const int64_t a = rows + 2 * 148 + 1;
  const int64_t b = rows * (int64_t)select_num_segments_streamed(n, 1 << 30, kk);
  return a > b ? a : b;
}

// Public launchers ------------------------------------------------------------------------------
int launch_select_segments(const float* scores, int64_t n, int64_t ld, int B, int S, int kk,
                           float* cand_scores, int32_t* cand_idx, const int32_t* query_flags,
                           cudaStream_t st) {
  SelectParams<int32_t> P{};
  P.vals = scores;
  P.payload = nullptr;
  P.n = n;
  P.ld = ld;
  P.S = S;
  P.kk = kk;
  P.out_vals = cand_scores;
  P.out_payload = cand_idx;
  P.query_flags = query_flags;
  return launch_select_t<int32_t, false>(P, B, st);
}

int launch_select_final_i32(const float* scores, const int32_t* payload, int64_t n, int64_t ld,
                            int B, int kk, float* out_scores, int32_t* out_idx, int64_t* out_ids,
                            const int64_t* id_map, const int32_t* query_flags, cudaStream_t st) {
  SelectParams<int32_t> P{};
  P.vals = scores;
  P.payload = payload;
  P.n = n;
  P.ld = ld;
  P.S = 1;
  P.kk = kk;
  P.out_vals = out_scores;
  P.out_payload = out_idx;
  P.out_ids = out_ids;
  P.id_map = id_map;
  P.query_flags = query_flags;
  return launch_select_t<int32_t, true>(P, B, st);
}

int launch_select_final_i64(const float* scores, const int64_t* payload, int64_t n, int64_t ld,
                            int B, int kk, float* out_scores, int64_t* out_ids, cudaStream_t st) {
  SelectParams<int64_t> P{};
  P.vals = scores;
  P.payload = payload;
  P.n = n;
  P.ld = ld;
  P.S = 1;
  P.kk = kk;
  P.out_vals = out_scores;
  P.out_payload = nullptr;
  P.out_ids = out_ids;
  P.id_map = nullptr;
  P.query_flags = nullptr;
  return launch_select_t<int64_t, true>(P, B, st);
}

}  // namespace mol
