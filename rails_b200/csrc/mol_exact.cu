// Exact fp32 MoL scoring on CUDA cores.
//
// Follows, operation by operation, the reference's eval-mode arithmetic
//   rails/similarities/mol/similarity_fn.py:389-405   logits = <Q_sub[n], X_sub[m]> / tau
//   rails/similarities/mol/similarity_fn.py:166-179   G = GQ*GI + W2 silu(W1 l + b1) + b2 ; w = G sigmoid(G)
//   rails/similarities/mol/similarity_fn.py:42-46     p = softmax(w); p /= clamp(sum p, eps); score = sum p*l
// It serves three roles: (1) the rescoring pass that turns the tensor-core pass' fp16 candidates into
// reference-exact fp32 scores/order, (2) MOL_MODE_EXACT brute force / MoLSimilarity.forward's (B, N)
// score matrix, (3) the per-query fallback when the coarse pass' safety check fails.
//
// Mapping: blockIdx.y = query; each warp takes groups of PT=8 items of that query.  A lane owns logit
// indices {lane + 32*ll} and hidden units {lane + 32*jj}; the two MLP layers are register-tiled
// (PT x LL / PT x HH accumulators per lane) with the transposed weights broadcast from shared memory.
#include <math_constants.h>

#include "common.cuh"

namespace mol {

constexpr int PT = 8;            // items per warp pass
constexpr int EX_WARPS = 8;      // warps per block
constexpr int EX_THREADS = EX_WARPS * 32;

struct ExactParams {
  const float* qsub;   // (B, Pq, d)
  const float* gq;     // (B, L)
  const float* xsub;   // (N, Px, d)
  const float* gi;     // (N, L)
  const float* w1t;    // (L, H)   W1t[l*H + j] = qi_w1[j, l]
  const float* w2t;    // (H, L)   W2t[j*L + l] = qi_w2[l, j]
  const float* b1;     // (H)
  const float* b2;     // (L)
  const int32_t* cand; // nullable
  const int32_t* query_flags;  // nullable
  float* scores;
  int64_t N, n_per_query, ld;
  int Pq, Px, d, L, H;
  float temperature, eps;
  int renorm;
  int w_in_smem;
};

__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows,
                                 int cols) {  // out[c*rows + r] = in[r*cols + c]
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  int r = i / cols, c = i % cols;
  out[c * rows + r] = in[i];
}

template <int LL, int HH>
__global__ void __launch_bounds__(EX_THREADS) exact_scores_kernel(ExactParams P) {
  extern __shared__ __align__(16) float smem[];
  const int L = LL * 32, H = HH * 32;
  const int b = blockIdx.y;
  if (P.query_flags && P.query_flags[b] == 0) return;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int d = P.d, Pq = P.Pq, Px = P.Px;
  const int qstride = d + 4;

  // ---- carve shared memory
  float* sp = smem;
  const float* W1t;
  const float* W2t;
  if (P.w_in_smem) {
    float* a = sp;
    sp += L * H;
    float* c = sp;
    sp += H * L;
    for (int i = threadIdx.x; i < L * H; i += EX_THREADS) {
      a[i] = P.w1t[i];
      c[i] = P.w2t[i];
    }
    W1t = a;
    W2t = c;
  } else {
    W1t = P.w1t;
    W2t = P.w2t;
  }
  float* b1s = sp;
  sp += H;
  float* b2s = sp;
  sp += L;
  float* gqs = sp;
  sp += L;
  float* Qs = sp;
  sp += Pq * qstride;
  float* logT = sp + warp * (L + H) * PT;  // [L][PT]
  float* hidT = logT + L * PT;             // [H][PT]

  for (int i = threadIdx.x; i < H; i += EX_THREADS) b1s[i] = P.b1[i];
  for (int i = threadIdx.x; i < L; i += EX_THREADS) {
    b2s[i] = P.b2[i];
    gqs[i] = P.gq[(int64_t)b * L + i];
  }
  for (int i = threadIdx.x; i < Pq * d; i += EX_THREADS)
    Qs[(i / d) * qstride + (i % d)] = P.qsub[(int64_t)b * Pq * d + i];
  __syncthreads();

  const int64_t groups = (P.n_per_query + PT - 1) / PT;
  for (int64_t g = (int64_t)blockIdx.x * EX_WARPS + warp; g < groups;
       g += (int64_t)gridDim.x * EX_WARPS) {
    int64_t item[PT];
    bool valid[PT];
#pragma unroll
    for (int p = 0; p < PT; ++p) {
      int64_t j = g * PT + p;
      int64_t x = -1;
      if (j < P.n_per_query) x = P.cand ? (int64_t)P.cand[(int64_t)b * P.ld + j] : j;
      valid[p] = (x >= 0 && x < P.N);
      item[p] = valid[p] ? x : 0;
    }

    // ---- 1. logits  l = n*Px + m  (einsum "bnd,xmd->bxnm" then / tau)
    float lg[PT][LL];
#pragma unroll
    for (int ll = 0; ll < LL; ++ll) {
      const int l = lane + 32 * ll;
      const int n = l / Px, m = l % Px;
      const float4* q4 = reinterpret_cast<const float4*>(Qs + n * qstride);
#pragma unroll
      for (int p = 0; p < PT; ++p) {
        const float4* x4 = reinterpret_cast<const float4*>(P.xsub + (item[p] * Px + m) * d);
        float s = 0.f;
        for (int i = 0; i < d / 4; ++i) {
          float4 a = q4[i], c = __ldg(x4 + i);
          s = fmaf(a.x, c.x, s);
          s = fmaf(a.y, c.y, s);
          s = fmaf(a.z, c.z, s);
          s = fmaf(a.w, c.w, s);
        }
        s = s / P.temperature;
        lg[p][ll] = s;
        logT[l * PT + p] = s;
      }
    }
    __syncwarp();

    // ---- 2. hidden = silu(W1 l + b1)
    {
      float acc[PT][HH];
#pragma unroll
      for (int jj = 0; jj < HH; ++jj) {
        float bv = b1s[lane + 32 * jj];
#pragma unroll
        for (int p = 0; p < PT; ++p) acc[p][jj] = bv;
      }
#pragma unroll 4
      for (int l = 0; l < L; ++l) {
        float wv[HH];
#pragma unroll
        for (int jj = 0; jj < HH; ++jj) wv[jj] = W1t[l * H + lane + 32 * jj];
        float4 x0 = *reinterpret_cast<const float4*>(logT + l * PT);
        float4 x1 = *reinterpret_cast<const float4*>(logT + l * PT + 4);
        float xv[PT] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
        for (int p = 0; p < PT; ++p)
#pragma unroll
          for (int jj = 0; jj < HH; ++jj) acc[p][jj] = fmaf(xv[p], wv[jj], acc[p][jj]);
      }
#pragma unroll
      for (int jj = 0; jj < HH; ++jj)
#pragma unroll
        for (int p = 0; p < PT; ++p) {
          float a = acc[p][jj];
          hidT[(lane + 32 * jj) * PT + p] = a / (1.f + expf(-a));
        }
    }
    __syncwarp();

    // ---- 3. gate pre-activation  G = gq*gi + (W2 h + b2)
    float G[PT][LL];
#pragma unroll
    for (int ll = 0; ll < LL; ++ll) {
      float bv = b2s[lane + 32 * ll];
#pragma unroll
      for (int p = 0; p < PT; ++p) G[p][ll] = bv;
    }
#pragma unroll 4
    for (int j = 0; j < H; ++j) {
      float wv[LL];
#pragma unroll
      for (int ll = 0; ll < LL; ++ll) wv[ll] = W2t[j * L + lane + 32 * ll];
      float4 h0 = *reinterpret_cast<const float4*>(hidT + j * PT);
      float4 h1 = *reinterpret_cast<const float4*>(hidT + j * PT + 4);
      float hv[PT] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
      for (int p = 0; p < PT; ++p)
#pragma unroll
        for (int ll = 0; ll < LL; ++ll) G[p][ll] = fmaf(hv[p], wv[ll], G[p][ll]);
    }
    __syncwarp();

    // ---- 4. w = G*sigmoid(G); softmax over L; renorm; weighted sum
#pragma unroll
    for (int p = 0; p < PT; ++p) {
      float wv[LL];
      float mx = -CUDART_INF_F;
#pragma unroll
      for (int ll = 0; ll < LL; ++ll) {
        const int l = lane + 32 * ll;
        float gi = __ldg(P.gi + item[p] * L + l);
        float gg = gqs[l] * gi + G[p][ll];
        float w = gg * (1.f / (1.f + expf(-gg)));
        wv[ll] = w;
        mx = fmaxf(mx, w);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float sum = 0.f;
#pragma unroll
      for (int ll = 0; ll < LL; ++ll) {
        wv[ll] = expf(wv[ll] - mx);
        sum += wv[ll];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      float psum = 0.f;
#pragma unroll
      for (int ll = 0; ll < LL; ++ll) {
        wv[ll] = wv[ll] / sum;
        psum += wv[ll];
      }
      if (P.renorm) {  // similarity_fn.py:43-45 (dropout is identity in eval, the renorm still runs)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, o);
        float den = fmaxf(psum, P.eps);
#pragma unroll
        for (int ll = 0; ll < LL; ++ll) wv[ll] = wv[ll] / den;
      }
      float sc = 0.f;
#pragma unroll
      for (int ll = 0; ll < LL; ++ll) sc = fmaf(wv[ll], lg[p][ll], sc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, o);
      int64_t j = g * PT + p;
      if (lane == 0 && j < P.n_per_query)
        P.scores[(int64_t)b * P.ld + j] = valid[p] ? sc : -CUDART_INF_F;
    }
    __syncwarp();
  }
}

template <int LL, int HH>
static int launch_exact_t(const ExactParams& P0, int B, cudaStream_t st) {
  ExactParams P = P0;
  const int L = LL * 32, H = HH * 32;
  size_t fixed = (size_t)(H + 2 * L + P.Pq * (P.d + 4) + EX_WARPS * (L + H) * PT) * sizeof(float);
  size_t with_w = fixed + (size_t)2 * L * H * sizeof(float);
  P.w_in_smem = with_w <= 200 * 1024;
  size_t smem = P.w_in_smem ? with_w : fixed;
  auto kern = exact_scores_kernel<LL, HH>;
  MOL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t groups = (P.n_per_query + PT - 1) / PT;
  int64_t max_x = (groups + EX_WARPS - 1) / EX_WARPS;
  int64_t want = (148 * 4 + B - 1) / B;
  int64_t gx = want < 1 ? 1 : want;
  if (gx > max_x) gx = max_x;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)B);
  kern<<<grid, EX_THREADS, smem, st>>>(P);
  MOL_LAUNCH_CHECK();
  return MOL_OK;
}

int launch_transpose(const float* in, float* out, int rows, int cols, cudaStream_t st) {
  int n = rows * cols;
  transpose_kernel<<<(n + 255) / 256, 256, 0, st>>>(in, out, rows, cols);
  MOL_LAUNCH_CHECK();
  return MOL_OK;
}

int launch_exact_scores(const mol_shape_t& s, const mol_weights_t& w, const mol_index_t& ix,
                        const float* w1t, const float* w2t, const float* qsub, const float* gq,
                        int B, const int32_t* cand, int64_t n_per_query, int64_t ld, float* scores,
                        const int32_t* query_flags, cudaStream_t st) {
  if (B == 0 || n_per_query == 0) return MOL_OK;
  Dims D = dims_of(s);
  ExactParams P;
  P.qsub = qsub;
  P.gq = gq;
  P.xsub = ix.xsub_f32;
  P.gi = ix.gi_f32;
  P.w1t = w1t;
  P.w2t = w2t;
  P.b1 = w.qi_b1;
  P.b2 = w.qi_b2;
  P.cand = cand;
  P.query_flags = query_flags;
  P.scores = scores;
  P.N = ix.num_items;
  P.n_per_query = n_per_query;
  P.ld = ld;
  P.Pq = D.Pq;
  P.Px = D.Px;
  P.d = D.d;
  P.L = D.L;
  P.H = D.H;
  P.temperature = s.temperature;
  P.eps = s.eps;
  P.renorm = s.softmax_renorm;
  P.w_in_smem = 0;
  MOL_CHECK_ARG(D.H == 128, "exact kernel supports gating_qi_hidden_dim == 128 (got %d)", D.H);
  switch (D.L) {
    case 32: return launch_exact_t<1, 4>(P, B, st);
    case 64: return launch_exact_t<2, 4>(P, B, st);
    case 128: return launch_exact_t<4, 4>(P, B, st);
    case 256: return launch_exact_t<8, 4>(P, B, st);
    default: MOL_CHECK_ARG(false, "exact kernel supports P_Q*P_X in {32,64,128,256} (got %d)", D.L);
  }
  return MOL_OK;
}

}  // namespace mol
