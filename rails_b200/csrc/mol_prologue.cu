// Item-side index build and query prologue kernels (fp32 CUDA cores).
//
// Reference arithmetic followed (see oracle/mol_oracle.py for the CPU restatement):
//   rails/similarities/mol/item_embeddings_fns.py:165-182   Linear -> reshape (P_X, d) -> l2 norm
//   rails/similarities/mol/query_embeddings_fns.py:191-253  GLU-MLP -> reshape -> cat uid emb -> l2 norm
//   rails/similarities/layers.py:36-43, 67-74               GeGLU / SwiGLU
//   rails/similarities/mol/similarity_fn.py:166-171         query-only / item-only gating MLPs
// These run once per corpus (index build) or are O(B) (query prologue); SURVEY.md §8 (f1) lists
// their fused/tensor-core versions as "next".
#include "common.cuh"

namespace mol {

// ------------------------------------------------------------------------------------------
// Tiled fp32 GEMM:  C[M,N] = act(A[M,K] W^T + bias),  64x64 tile, 64-deep K slab, 4x4 per thread.  The next slab is
// fetched into registers while the current one is multiplied: the query prologue runs these GEMMs with a handful
// of blocks (M = B <= 512), where each exposed global-memory round trip costs about a microsecond.
// ------------------------------------------------------------------------------------------
constexpr int LT_M = 64, LT_N = 64, LT_K = 64;
constexpr int LT_LD = LT_M * LT_K / 256;  // elements of each operand tile a thread fetches per slab

template <int ACT>
__global__ void __launch_bounds__(256, 2) linear_kernel(const float* __restrict__ A,
                                                     const float* __restrict__ W,
                                                     const float* __restrict__ bias,
                                                     float* __restrict__ C, int64_t M, int N, int K,
                                                     int64_t w_sn, int64_t w_sk) {
  __shared__ __align__(16) float As[LT_K][LT_M + 4];
  __shared__ __align__(16) float Ws[LT_K][LT_N + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * LT_M;
  const int n0 = blockIdx.y * LT_N;
  const int tx = tid % 16, ty = tid / 16;  // tx -> n, ty -> m
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float ra[LT_LD], rw[LT_LD];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int u = 0; u < LT_LD; ++u) {  // A tile: consecutive threads walk k (contiguous in memory)
      const int e = tid + u * 256;
      const int r = e / LT_K, kk = e % LT_K;
      const int64_t m = m0 + r;
      const int k = k0 + kk;
      ra[u] = (m < M && k < K) ? __ldg(A + m * K + k) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < LT_LD; ++u) {
      const int e = tid + u * 256;
      int r, kk;
      if (w_sk == 1) {  // (n,k) with k contiguous
        r = e / LT_K;
        kk = e % LT_K;
      } else {  // (k,n) with n contiguous
        kk = e / LT_N;
        r = e % LT_N;
      }
      const int n = n0 + r, k = k0 + kk;
      rw[u] = (n < N && k < K) ? __ldg(W + n * w_sn + k * w_sk) : 0.f;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int u = 0; u < LT_LD; ++u) {
      const int e = tid + u * 256;
      As[e % LT_K][e / LT_K] = ra[u];
      if (w_sk == 1)
        Ws[e % LT_K][e / LT_K] = rw[u];
      else
        Ws[e / LT_N][e % LT_N] = rw[u];
    }
  };

  fetch(0);
  for (int k0 = 0; k0 < K; k0 += LT_K) {
    stash();
    __syncthreads();
    if (k0 + LT_K < K) fetch(k0 + LT_K);
    const int kmax = K - k0 < LT_K ? K - k0 : LT_K;  // (the slab's zero padding is not multiplied: K = d = 32 for the prefilters)
#pragma unroll 8
    for (int kk = 0; kk < kmax; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (ACT == ACT_SILU) v = v / (1.f + expf(-v));
      C[m * N + n] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------
// W-resident variant for the index build (M = corpus size, N * K small): persistent blocks keep the whole weight
// matrix in shared memory (k-major, read as float4 along n), walk the 64-row tiles of A with the next tile
// prefetched into registers, and give every thread a 4 x (4 * NJ) register tile (5 LDS.128 per 64 FMAs at NJ = 4).
// Same per-element arithmetic as linear_kernel: acc = 0, fmaf over k ascending, + bias, optional silu.
// ------------------------------------------------------------------------------------------
constexpr int WR_TM = 64;

template <int ACT, int NJ, int K>
__global__ void __launch_bounds__(256, 2) linear_wres_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                             const float* __restrict__ bias, float* __restrict__ C,
                                                             int64_t M, int64_t w_sn, int64_t w_sk) {
  constexpr int N = 64 * NJ;
  extern __shared__ __align__(16) float wr_smem[];
  float* Ws = wr_smem;                // [K][N + 4]
  float* As = Ws + (size_t)K * (N + 4);  // [K][WR_TM + 4]
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  for (int e = tid; e < N * K; e += 256) {
    int n, k;
    if (w_sk == 1) {
      n = e / K;
      k = e % K;
    } else {
      k = e / N;
      n = e % N;
    }
    Ws[k * (N + 4) + n] = __ldg(W + n * w_sn + k * w_sk);
  }
  constexpr int k4 = K / 4;             // float4 per row of A
  constexpr int per_thread = WR_TM * k4 / 256;
  static_assert(WR_TM * k4 % 256 == 0, "A tile must split evenly over the block");
  float4 ra[per_thread];
  const int64_t tiles = (M + WR_TM - 1) / WR_TM;
  auto fetch = [&](int64_t tile) {
#pragma unroll
    for (int u = 0; u < per_thread; ++u) {
      const int e = tid + u * 256;
      const int r = e / k4, c = e % k4;
      const int64_t m = tile * WR_TM + r;
      ra[u] = (m < M) ? __ldg(reinterpret_cast<const float4*>(A + m * K) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int u = 0; u < per_thread; ++u) {
      const int e = tid + u * 256;
      const int r = e / k4, c = e % k4;
      As[(4 * c + 0) * (WR_TM + 4) + r] = ra[u].x;
      As[(4 * c + 1) * (WR_TM + 4) + r] = ra[u].y;
      As[(4 * c + 2) * (WR_TM + 4) + r] = ra[u].z;
      As[(4 * c + 3) * (WR_TM + 4) + r] = ra[u].w;
    }
  };
  int64_t tile = blockIdx.x;
  if (tile < tiles) fetch(tile);
  for (; tile < tiles; tile += gridDim.x) {
    __syncthreads();  // the previous tile's As reads are done (and, first time, Ws is complete)
    stash();
    __syncthreads();
    if (tile + gridDim.x < tiles) fetch(tile + gridDim.x);
    float acc[4][4 * NJ];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4 * NJ; ++j) acc[i][j] = 0.f;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(As + k * (WR_TM + 4) + ty * 4);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const float4 b4 = *reinterpret_cast<const float4*>(Ws + k * (N + 4) + 64 * j + tx * 4);
        const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) acc[i][4 * j + jj] = fmaf(a[i], b[jj], acc[i][4 * j + jj]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t m = tile * WR_TM + ty * 4 + i;
      if (m >= M) continue;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int n = 64 * j + tx * 4;
        float v[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          v[jj] = acc[i][4 * j + jj] + (bias ? bias[n + jj] : 0.f);
          if (ACT == ACT_SILU) v[jj] = v[jj] / (1.f + expf(-v[jj]));
        }
        *reinterpret_cast<float4*>(C + m * N + n) = make_float4(v[0], v[1], v[2], v[3]);
      }
    }
  }
}

template <int ACT>
static bool try_launch_wres(const float* A, const float* W, const float* bias, float* C, int64_t M, int N, int K,
                            int64_t w_sn, int64_t w_sk, cudaStream_t st) {
  // (N = 64 leaves a 4 x 4 register tile per thread: measured slower than the plain tiled kernel, 686 vs 635 us per 1M rows)
  if (M < 16384 || (K != 64 && K != 128) || (N != 128 && N != 256)) return false;
  if (((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(C)) & 15) != 0) return false;
  const size_t smem = ((size_t)K * (N + 4) + (size_t)K * (WR_TM + 4)) * sizeof(float);
  if (smem > 100 * 1024) return false;
  const int64_t tiles = (M + WR_TM - 1) / WR_TM;
  const unsigned grid = (unsigned)(tiles < 296 ? tiles : 296);
#define MOL_WRES_K(NJ, KK)                                                                                          \
  {                                                                                                                   \
    cudaFuncSetAttribute(linear_wres_kernel<ACT, NJ, KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
    linear_wres_kernel<ACT, NJ, KK><<<grid, 256, smem, st>>>(A, W, bias, C, M, w_sn, w_sk);                            \
  }
#define MOL_WRES(NJ)                          \
  {                                           \
    if (K == 64) MOL_WRES_K(NJ, 64) else MOL_WRES_K(NJ, 128) \
  }
  if (N == 128) MOL_WRES(2) else MOL_WRES(4)
#undef MOL_WRES
#undef MOL_WRES_K
  return true;
}

// ------------------------------------------------------------------------------------------
// Small-batch forms (M <= 8 rows and N <= 4096 columns: the query prologue of a latency-bound search).  The tiled kernel
// above runs such a GEMM with N / 64 blocks that each walk K serially behind shared-memory staging (29 us for the
// 512 -> 448 projection of ONE query).  Here every output column gets its own warp when the weights are contiguous along
// K (coalesced reads, lane-strided partial sums, a shuffle reduction) or its own thread when they are contiguous along N.
// Measured: cfg1 (ML-1M, one query) 0.132 -> 0.104 ms per search.  A thread-per-column kernel that keeps the tiled
// kernel's k-ordered fmaf chain (bit-identical sums) was measured too: no faster than the tiled kernel (0.136 ms), the
// 512-long dependent chain is what costs.  So a query's Q_sub may differ in the last bit between a batch of <= 8 and a
// larger one (as PyTorch's own GEMMs do); within a batch size it is deterministic.
// ------------------------------------------------------------------------------------------
constexpr int SM_MAX_M = 8, SM_MAX_N = 4096;

template <int ACT>
__global__ void __launch_bounds__(256) linear_small_kmajor_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                                  const float* __restrict__ bias, float* __restrict__ C,
                                                                  int M, int N, int K, int64_t w_sn) {
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);  // one warp per output column
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  float acc[SM_MAX_M];
#pragma unroll
  for (int m = 0; m < SM_MAX_M; ++m) acc[m] = 0.f;
  const float* w = W + (int64_t)n * w_sn;
  for (int k = lane; k < K; k += 32) {
    const float wv = __ldg(w + k);
#pragma unroll
    for (int m = 0; m < SM_MAX_M; ++m)
      if (m < M) acc[m] = fmaf(__ldg(A + (int64_t)m * K + k), wv, acc[m]);
  }
#pragma unroll
  for (int m = 0; m < SM_MAX_M; ++m) {
    if (m < M) {
      float v = acc[m];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) {
        v += bias ? bias[n] : 0.f;
        if (ACT == ACT_SILU) v = v / (1.f + expf(-v));
        C[(int64_t)m * N + n] = v;
      }
    }
  }
}

template <int ACT>
__global__ void __launch_bounds__(128) linear_small_nmajor_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                                  const float* __restrict__ bias, float* __restrict__ C,
                                                                  int M, int N, int K, int64_t w_sk) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;  // one thread per output column
  if (n >= N) return;
  float acc[SM_MAX_M];
#pragma unroll
  for (int m = 0; m < SM_MAX_M; ++m) acc[m] = 0.f;
#pragma unroll 8
  for (int k = 0; k < K; ++k) {
    const float wv = __ldg(W + (int64_t)k * w_sk + n);
#pragma unroll
    for (int m = 0; m < SM_MAX_M; ++m)
      if (m < M) acc[m] = fmaf(__ldg(A + (int64_t)m * K + k), wv, acc[m]);
  }
#pragma unroll
  for (int m = 0; m < SM_MAX_M; ++m) {
    if (m < M) {
      float v = acc[m] + (bias ? bias[n] : 0.f);
      if (ACT == ACT_SILU) v = v / (1.f + expf(-v));
      C[(int64_t)m * N + n] = v;
    }
  }
}

int launch_linear(const float* A, const float* W, const float* bias, float* C, int64_t M, int N,
                  int K, int64_t w_sn, int64_t w_sk, Act act, cudaStream_t st) {
  if (M == 0 || N == 0) return MOL_OK;
  if (M <= SM_MAX_M && N <= SM_MAX_N && (w_sk == 1 || w_sn == 1)) {
    if (w_sk == 1) {
      const unsigned grid = (unsigned)((N + 7) / 8);
      if (act == ACT_SILU)
        linear_small_kmajor_kernel<ACT_SILU><<<grid, 256, 0, st>>>(A, W, bias, C, (int)M, N, K, w_sn);
      else
        linear_small_kmajor_kernel<ACT_NONE><<<grid, 256, 0, st>>>(A, W, bias, C, (int)M, N, K, w_sn);
    } else {
      const unsigned grid = (unsigned)((N + 127) / 128);
      if (act == ACT_SILU)
        linear_small_nmajor_kernel<ACT_SILU><<<grid, 128, 0, st>>>(A, W, bias, C, (int)M, N, K, w_sk);
      else
        linear_small_nmajor_kernel<ACT_NONE><<<grid, 128, 0, st>>>(A, W, bias, C, (int)M, N, K, w_sk);
    }
    MOL_LAUNCH_CHECK();
    return MOL_OK;
  }
  if (act == ACT_SILU ? try_launch_wres<ACT_SILU>(A, W, bias, C, M, N, K, w_sn, w_sk, st)
                      : try_launch_wres<ACT_NONE>(A, W, bias, C, M, N, K, w_sn, w_sk, st)) {
    MOL_LAUNCH_CHECK();
    return MOL_OK;
  }
  dim3 grid((unsigned)((M + LT_M - 1) / LT_M), (unsigned)((N + LT_N - 1) / LT_N));
  if (act == ACT_SILU)
    linear_kernel<ACT_SILU><<<grid, 256, 0, st>>>(A, W, bias, C, M, N, K, w_sn, w_sk);
  else
    linear_kernel<ACT_NONE><<<grid, 256, 0, st>>>(A, W, bias, C, M, N, K, w_sn, w_sk);
  MOL_LAUNCH_CHECK();
  return MOL_OK;
}

// ------------------------------------------------------------------------------------------
__global__ void glu_kernel(const float* __restrict__ pre, float* __restrict__ h, int64_t total,
                           int Hq, int kind) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int64_t b = i / Hq;
  int j = (int)(i % Hq);
  float lhs = pre[b * 2 * Hq + j], rhs = pre[b * 2 * Hq + Hq + j];
  float a;
  if (kind == 0) {
    a = 0.5f * lhs * (1.f + erff(lhs * 0.70710678118654752440f));  // F.gelu (erf form)
  } else {
    a = lhs / (1.f + expf(-lhs));  // F.silu
  }
  h[i] = a * rhs;
}

int launch_glu(const float* pre, float* h, int64_t B, int Hq, int kind, cudaStream_t st) {
  int64_t total = B * Hq;
  if (total == 0) return MOL_OK;
  glu_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(pre, h, total, Hq, kind);
  MOL_LAUNCH_CHECK();
  return MOL_OK;
}

// ------------------------------------------------------------------------------------------
// One warp per d-vector.
__global__ void l2norm_groups_kernel(const float* __restrict__ in, float* __restrict__ out_f32,
                                     __half* __restrict__ out_half, int64_t n_vec, int d,
                                     float eps) {
  int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 32;
  int lane = threadIdx.x % 32;
  if (v >= n_vec) return;
  const float* p = in + v * d;
  float ss = 0.f;
  for (int i = lane; i < d; i += 32) ss = fmaf(p[i], p[i], ss);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  float nrm = fmaxf(sqrtf(ss), eps);
  for (int i = lane; i < d; i += 32) {
    float y = p[i] / nrm;
    if (out_f32) out_f32[v * d + i] = y;
    if (out_half) out_half[v * d + i] = __float2half_rn(y);  // |y| <= 1: always representable
  }
}

// d / 4 = G lanes (a power of two <= 32) per d-vector, one float4 per lane, UNR vectors in flight per lane group:
// the index build normalises N * P_X vectors (HBM-bound: 4 B read + 6 B written per element).
template <int G, int UNR>
__global__ void __launch_bounds__(256) l2norm_groups_vec_kernel(const float4* __restrict__ in, float4* __restrict__ out_f32,
                                                                uint2* __restrict__ out_half, int64_t n_vec, float eps) {
  const int64_t slot = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;  // lane group index
  const int gl = threadIdx.x % G;
  const int64_t n_slots = (int64_t)gridDim.x * blockDim.x / G;
  const int64_t warp_slot = slot - (threadIdx.x % 32) / G;  // first lane group of this warp: warp-uniform loop bound
  for (int64_t w0 = warp_slot, v0 = slot; w0 < n_vec; w0 += n_slots * UNR, v0 += n_slots * UNR) {
    float4 x[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int64_t v = v0 + (int64_t)u * n_slots;
      x[u] = v < n_vec ? __ldg(in + v * G + gl) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int64_t v = v0 + (int64_t)u * n_slots;
      float ss = x[u].x * x[u].x;
      ss = fmaf(x[u].y, x[u].y, ss);
      ss = fmaf(x[u].z, x[u].z, ss);
      ss = fmaf(x[u].w, x[u].w, ss);
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float nrm = fmaxf(sqrtf(ss), eps);
      const float4 y = make_float4(x[u].x / nrm, x[u].y / nrm, x[u].z / nrm, x[u].w / nrm);
      if (v < n_vec) {
        if (out_f32) out_f32[v * G + gl] = y;
        if (out_half) {  // |y| <= 1: always representable
          const __half2 lo = __floats2half2_rn(y.x, y.y), hi = __floats2half2_rn(y.z, y.w);
          out_half[v * G + gl] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
        }
      }
    }
  }
}

template <int G>
static void launch_l2norm_vec(const float* in, float* out_f32, __half* out_half, int64_t n_vec, float eps, cudaStream_t st) {
  constexpr int UNR = 4;
  const int64_t slots_per_block = 256 / G;
  int64_t blocks = (n_vec + slots_per_block * UNR - 1) / (slots_per_block * UNR);
  if (blocks < 1) blocks = 1;
  l2norm_groups_vec_kernel<G, UNR><<<(unsigned)blocks, 256, 0, st>>>(
      reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out_f32), reinterpret_cast<uint2*>(out_half), n_vec, eps);
}

int launch_l2norm_groups(const float* in, float* out_f32, __half* out_half, int64_t rows,
                         int groups, int d, float eps, cudaStream_t st) {
  int64_t n_vec = rows * groups;
  if (n_vec == 0) return MOL_OK;
  const bool aligned = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out_f32)) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(out_half) & 7) == 0;
  if (aligned && (d == 32 || d == 64 || d == 128) && n_vec >= 4096) {
    if (d == 32) launch_l2norm_vec<8>(in, out_f32, out_half, n_vec, eps, st);
    else if (d == 64) launch_l2norm_vec<16>(in, out_f32, out_half, n_vec, eps, st);
    else launch_l2norm_vec<32>(in, out_f32, out_half, n_vec, eps, st);
    MOL_LAUNCH_CHECK();
    return MOL_OK;
  }
  int64_t threads = n_vec * 32;
  l2norm_groups_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(in, out_f32, out_half,
                                                                          n_vec, d, eps);
  MOL_LAUNCH_CHECK();
  return MOL_OK;
}

// ------------------------------------------------------------------------------------------
struct UidTables {
  const float* emb[MOL_MAX_UID_TABLES];
  int hash[MOL_MAX_UID_TABLES];
};

// One warp per (b, n): n < Pq_proj reads the projected vector, else the uid-embedding row
// (user_ids % hash) + 1  (query_embeddings_fns.py:205-207), then l2-normalises.
__global__ void query_assemble_kernel(const float* __restrict__ proj,
                                      const int64_t* __restrict__ user_ids, UidTables t,
                                      float* __restrict__ qsub, int B, int Pq, int Pq_proj, int d,
                                      float eps) {
  int v = (blockIdx.x * blockDim.x + threadIdx.x) / 32;
  int lane = threadIdx.x % 32;
  if (v >= B * Pq) return;
  int b = v / Pq, n = v % Pq;
  const float* p;
  if (n < Pq_proj) {
    p = proj + ((int64_t)b * Pq_proj + n) * d;
  } else {
    int ti = n - Pq_proj;
    int64_t h = t.hash[ti];
    int64_t r = user_ids[b] % h;
    if (r < 0) r += h;  // python-style modulo
    p = t.emb[ti] + (r + 1) * d;
  }
  float ss = 0.f;
  for (int i = lane; i < d; i += 32) ss = fmaf(p[i], p[i], ss);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  float nrm = fmaxf(sqrtf(ss), eps);
  for (int i = lane; i < d; i += 32) qsub[(int64_t)v * d + i] = p[i] / nrm;
}

int launch_query_assemble(const mol_shape_t& s, const mol_weights_t& w, const float* proj,
                          const int64_t* user_ids, float* qsub, int B, cudaStream_t st) {
  Dims D = dims_of(s);
  if (B == 0) return MOL_OK;
  UidTables t;
  for (int i = 0; i < MOL_MAX_UID_TABLES; ++i) {
    t.emb[i] = i < D.u ? w.uid_emb[i] : nullptr;
    t.hash[i] = i < D.u ? s.uid_hash_sizes[i] : 1;
  }
  int threads = B * D.Pq * 32;
  query_assemble_kernel<<<(threads + 255) / 256, 256, 0, st>>>(proj, user_ids, t, qsub, B, D.Pq,
                                                               D.Pq_proj, D.d, s.eps);
  MOL_LAUNCH_CHECK();
  return MOL_OK;
}

}  // namespace mol
