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
// Mapping: persistent blocks (one per SM) stage the qi-MLP weights in shared memory ONCE; the flat list of work items
// (active query, group of PT=8 items) is split into one contiguous range per warp, so a warp re-stages its query
// (Q_sub, gq) only when the query changes.  A lane owns logit indices {lane + 32*ll} and hidden units {lane + 32*jj};
// the two MLP layers are register-tiled (PT x LL / PT x HH accumulators per lane) with the transposed weights
// broadcast from shared memory.  The item rows of the NEXT work item are prefetched into L2 while the MLP of the
// current one runs (the candidates of the rescoring pass are random rows: without it every pass waits on DRAM).
#include <math_constants.h>

#include "common.cuh"

namespace mol {

constexpr int PT = 8;            // items per warp pass
#ifndef EXACT_NW_SMALL
#define EXACT_NW_SMALL 8      // warps per block for L <= 64 (staged rows: 8 x 18.8 KB + 64 KB of weights)
#endif

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
  int B;
  int Pq, Px, d, L, H;
  float temperature, eps;
  int renorm;
};

__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows,
                                 int cols) {  // out[c*rows + r] = in[r*cols + c]
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  int r = i / cols, c = i % cols;
  out[c * rows + r] = in[i];
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Warp-uniform walk over the ACTIVE queries (query_flags[b] != 0, or all of them): seek(a) positions on the a-th
// active query, next() moves to the following one.  All lanes hold the same state.
struct ActiveQueries {
  const int32_t* flags;
  int B, lane;
  int b;  // current query (B when exhausted)
  __device__ ActiveQueries(const int32_t* f, int B_, int lane_) : flags(f), B(B_), lane(lane_), b(-1) {}
  __device__ int count() const {
    if (!flags) return B;
    int c = 0;
    for (int i = 0; i < B; i += 32) {
      const int q = i + lane;
      c += __popc(__ballot_sync(0xffffffffu, q < B && flags[q] != 0));
    }
    return c;
  }
  __device__ void seek(int a) {
    if (!flags) {
      b = a < B ? a : B;
      return;
    }
    int seen = 0;
    for (int i = 0; i < B; i += 32) {
      const int q = i + lane;
      const unsigned m = __ballot_sync(0xffffffffu, q < B && flags[q] != 0);
      const int c = __popc(m);
      if (seen + c > a) {  // the a-th active query is in this word: its (a - seen)-th set bit
        unsigned mm = m;
        for (int t = 0; t < a - seen; ++t) mm &= mm - 1;
        b = i + __ffs(mm) - 1;
        return;
      }
      seen += c;
    }
    b = B;
  }
  __device__ void next() {
    if (!flags) {
      b = b + 1 < B ? b + 1 : B;
      return;
    }
    int q0 = b + 1;
    for (int i = q0 & ~31; i < B; i += 32) {
      const int q = i + lane;
      const unsigned m = __ballot_sync(0xffffffffu, q >= q0 && q < B && flags[q] != 0);
      if (m) {
        b = i + __ffs(m) - 1;
        return;
      }
    }
    b = B;
  }
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// WS: the transposed qi-MLP weights live in shared memory (else they are read through L1 from global memory).
// ST: the X_sub rows of a work item are staged in shared memory by cp.async, one work item ahead (else __ldg).
template <int LL, int HH, int NW, bool WS, bool ST>
__global__ void __launch_bounds__(NW * 32) exact_scores_kernel(ExactParams P) {
  extern __shared__ __align__(16) float smem[];
  constexpr int L = LL * 32, H = HH * 32;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int d = P.d, Pq = P.Pq, Px = P.Px;
  const int qstride = d + 4;

  // ---- this warp's contiguous range of (active query, item group) work items; blocks without work leave before
  //      staging anything (the fallback launches usually have no active query at all)
  ActiveQueries aq(P.query_flags, P.B, lane);
  const int64_t groups = (P.n_per_query + PT - 1) / PT;
  const int64_t T = (int64_t)aq.count() * groups;
  const int64_t W = (int64_t)gridDim.x * NW, gw = (int64_t)blockIdx.x * NW + warp;
  if (T * ((int64_t)blockIdx.x * NW) / W >= T * ((int64_t)(blockIdx.x + 1) * NW) / W) return;  // block-uniform
  int64_t t = T * gw / W;
  const int64_t t1 = T * (gw + 1) / W;

  // ---- carve shared memory: [W1t | W2t] | b1 | b2 | per warp: gq, Q_sub, logits^T, hidden^T, staged item rows
  float* sp = smem;
  float* W1s = sp;
  float* W2s = sp + (WS ? L * H : 0);
  if (WS) {
    sp += 2 * L * H;
    for (int i = threadIdx.x; i < L * H / 4; i += NW * 32) {
      reinterpret_cast<float4*>(W1s)[i] = __ldg(reinterpret_cast<const float4*>(P.w1t) + i);
      reinterpret_cast<float4*>(W2s)[i] = __ldg(reinterpret_cast<const float4*>(P.w2t) + i);
    }
  }
  float* b1s = sp;
  sp += H;
  float* b2s = sp;
  sp += L;
  const int xrow = Px * qstride;  // staged item row: Px groups of d floats, padded like Q_sub (bank-conflict-free)
  float* wsm = sp + (size_t)warp * (L + Pq * qstride + (L + H) * PT + (ST ? PT * xrow : 0));
  float* gqs = wsm;                    // [L]
  float* Qs = gqs + L;                 // [Pq][qstride]
  float* logT = Qs + Pq * qstride;     // [L][PT]
  float* hidT = logT + L * PT;         // [H][PT]
  float* Xst = hidT + H * PT;          // [PT][Px][qstride]   (ST only)

  for (int i = threadIdx.x; i < H; i += NW * 32) b1s[i] = P.b1[i];
  for (int i = threadIdx.x; i < L; i += NW * 32) b2s[i] = P.b2[i];
  __syncthreads();
  if (t >= t1) return;
  int64_t g = t % groups;
  aq.seek((int)(t / groups));
  int staged = -1;

  auto item_of = [&](int b, int64_t j) -> int64_t {  // -1: padding / out of range
    if (j >= P.n_per_query) return -1;
    const int64_t x = P.cand ? (int64_t)P.cand[(int64_t)b * P.ld + j] : j;
    return (x >= 0 && x < P.N) ? x : -1;
  };
  const int cpr = Px * d / 4;  // 16-byte chunks per item row
  const int d4 = d / 4;
  // lane p < PT holds the item of slot p of a work item; `stage_rows` starts the copy of those rows into Xst
  auto stage_rows = [&](int64_t mine) {
    for (int c0 = 0; c0 < PT * cpr; c0 += 32) {  // warp-uniform trip count (the shuffle needs every lane)
      const int c = c0 + lane;
      const int p = min(c / cpr, PT - 1), r = c - p * cpr;
      const int64_t x = __shfl_sync(0xffffffffu, mine, p);
      const int m = r / d4, i4 = r - m * d4;
      if (c < PT * cpr && x >= 0) cp_async16(Xst + p * xrow + m * qstride + 4 * i4, reinterpret_cast<const float4*>(P.xsub + x * Px * d) + r);
    }
  };
  const int gi_lines = (L * 4 + 127) / 128, row_lines = (Px * d * 4 + 127) / 128;

  int64_t mine = lane < PT ? item_of(aq.b, g * PT + lane) : -1;
  if (ST) stage_rows(mine);

  for (; t < t1; ++t) {
    const int b = aq.b;
    if (b != staged) {  // (re)stage this query: gq and Q_sub
      __syncwarp();
      for (int i = lane; i < L; i += 32) gqs[i] = P.gq[(int64_t)b * L + i];
      for (int i = lane; i < Pq * d; i += 32) Qs[(i / d) * qstride + (i % d)] = P.qsub[(int64_t)b * Pq * d + i];
      staged = b;
    }
    int64_t item[PT];
    bool valid[PT];
#pragma unroll
    for (int p = 0; p < PT; ++p) {
      const int64_t x = __shfl_sync(0xffffffffu, mine, p);
      valid[p] = x >= 0;
      item[p] = valid[p] ? x : 0;
    }
    // next work item (warp-uniform)
    int64_t gn = g + 1;
    int bn = b;
    if (gn == groups) {
      gn = 0;
      aq.next();
      bn = aq.b;
    }
    const bool more = t + 1 < t1 && bn < P.B;
    const int64_t mine_n = (more && lane < PT) ? item_of(bn, gn * PT + lane) : -1;
    if (ST) cp_async_wait_all();
    __syncwarp();

    // ---- 1. logits  l = n*Px + m  (einsum "bnd,xmd->bxnm" then / tau).  k runs outermost with the PT*LL dot products
    //      as independent chains (each still accumulates in k order); when P_X divides 32 the item group m of a
    //      lane is the same for all its logits, so one X_sub fragment serves them all.
    float lg[PT][LL];
    {
      float s[PT][LL];
#pragma unroll
      for (int p = 0; p < PT; ++p)
#pragma unroll
        for (int ll = 0; ll < LL; ++ll) s[p][ll] = 0.f;
      const bool same_m = (32 % Px) == 0;
      const int m0 = lane % Px;
      const float4* qrow[LL];
#pragma unroll
      for (int ll = 0; ll < LL; ++ll) qrow[ll] = reinterpret_cast<const float4*>(Qs + ((lane + 32 * ll) / Px) * qstride);
      if (ST && same_m) {
        const float4* xbase = reinterpret_cast<const float4*>(Xst + m0 * qstride);
#pragma unroll 2
        for (int i = 0; i < d4; ++i) {
          float4 a[LL], c[PT];
#pragma unroll
          for (int ll = 0; ll < LL; ++ll) a[ll] = qrow[ll][i];
#pragma unroll
          for (int p = 0; p < PT; ++p) c[p] = xbase[p * (xrow / 4) + i];
#pragma unroll
          for (int p = 0; p < PT; ++p)
#pragma unroll
            for (int ll = 0; ll < LL; ++ll) {
              s[p][ll] = fmaf(a[ll].x, c[p].x, s[p][ll]);
              s[p][ll] = fmaf(a[ll].y, c[p].y, s[p][ll]);
              s[p][ll] = fmaf(a[ll].z, c[p].z, s[p][ll]);
              s[p][ll] = fmaf(a[ll].w, c[p].w, s[p][ll]);
            }
        }
      } else {
#pragma unroll
        for (int ll = 0; ll < LL; ++ll) {
          const int m = (lane + 32 * ll) % Px;
#pragma unroll 2
          for (int i = 0; i < d4; ++i) {
            const float4 a = qrow[ll][i];
            float4 c[PT];
#pragma unroll
            for (int p = 0; p < PT; ++p)
              c[p] = ST ? reinterpret_cast<const float4*>(Xst + p * xrow + m * qstride)[i]
                        : __ldg(reinterpret_cast<const float4*>(P.xsub + (item[p] * Px + m) * d) + i);
#pragma unroll
            for (int p = 0; p < PT; ++p) {
              s[p][ll] = fmaf(a.x, c[p].x, s[p][ll]);
              s[p][ll] = fmaf(a.y, c[p].y, s[p][ll]);
              s[p][ll] = fmaf(a.z, c[p].z, s[p][ll]);
              s[p][ll] = fmaf(a.w, c[p].w, s[p][ll]);
            }
          }
        }
      }
#pragma unroll
      for (int p = 0; p < PT; ++p)
#pragma unroll
        for (int ll = 0; ll < LL; ++ll) {
          const float v = s[p][ll] / P.temperature;
          lg[p][ll] = v;
          logT[(lane + 32 * ll) * PT + p] = v;
        }
    }
    __syncwarp();

    if (more) {  // next work item: start the copy of its X_sub rows (ST) / prefetch them into L2, prefetch its GI rows
      if (ST) stage_rows(mine_n);
      const int lines = gi_lines + (ST ? 0 : row_lines);
      for (int i0 = 0; i0 < PT * lines; i0 += 32) {
        const int i = i0 + lane;
        const int p = min(i / lines, PT - 1), ln = i - p * lines;
        const int64_t x = __shfl_sync(0xffffffffu, mine_n, p);
        if (i < PT * lines && x >= 0)
          prefetch_l2(ln < gi_lines ? reinterpret_cast<const char*>(P.gi + x * L) + ln * 128
                                    : reinterpret_cast<const char*>(P.xsub + x * Px * d) + (ln - gi_lines) * 128);
      }
    }

    // ---- 2. hidden = silu(W1 l + b1).  The accumulators of two items share one packed fma.rn.f32x2 (same
    //      per-element rounding as fmaf, half the issue slots).
    {
      float2 acc[PT / 2][HH];
#pragma unroll
      for (int jj = 0; jj < HH; ++jj) {
        float bv = b1s[lane + 32 * jj];
#pragma unroll
        for (int q = 0; q < PT / 2; ++q) acc[q][jj] = make_float2(bv, bv);
      }
#pragma unroll 4
      for (int l = 0; l < L; ++l) {
        float2 wv[HH];
#pragma unroll
        for (int jj = 0; jj < HH; ++jj) {
          const float w = WS ? W1s[l * H + lane + 32 * jj] : __ldg(P.w1t + l * H + lane + 32 * jj);
          wv[jj] = make_float2(w, w);
        }
        float4 x0 = *reinterpret_cast<const float4*>(logT + l * PT);
        float4 x1 = *reinterpret_cast<const float4*>(logT + l * PT + 4);
        const float2 xv[PT / 2] = {make_float2(x0.x, x0.y), make_float2(x0.z, x0.w), make_float2(x1.x, x1.y),
                                   make_float2(x1.z, x1.w)};
#pragma unroll
        for (int q = 0; q < PT / 2; ++q)
#pragma unroll
          for (int jj = 0; jj < HH; ++jj) acc[q][jj] = __ffma2_rn(xv[q], wv[jj], acc[q][jj]);
      }
#pragma unroll
      for (int jj = 0; jj < HH; ++jj)
#pragma unroll
        for (int p = 0; p < PT; ++p) {
          float a = (p & 1) ? acc[p / 2][jj].y : acc[p / 2][jj].x;
          hidT[(lane + 32 * jj) * PT + p] = a / (1.f + expf(-a));
        }
    }
    __syncwarp();

    // ---- 3. gate pre-activation  G = gq*gi + (W2 h + b2)   (the GI rows are requested first, used in step 4)
    float giv[PT][LL];
#pragma unroll
    for (int p = 0; p < PT; ++p)
#pragma unroll
      for (int ll = 0; ll < LL; ++ll) giv[p][ll] = __ldg(P.gi + item[p] * L + lane + 32 * ll);
    float2 G2[PT / 2][LL];
#pragma unroll
    for (int ll = 0; ll < LL; ++ll) {
      float bv = b2s[lane + 32 * ll];
#pragma unroll
      for (int q = 0; q < PT / 2; ++q) G2[q][ll] = make_float2(bv, bv);
    }
#pragma unroll 4
    for (int j = 0; j < H; ++j) {
      float2 wv[LL];
#pragma unroll
      for (int ll = 0; ll < LL; ++ll) {
        const float w = WS ? W2s[j * L + lane + 32 * ll] : __ldg(P.w2t + j * L + lane + 32 * ll);
        wv[ll] = make_float2(w, w);
      }
      float4 h0 = *reinterpret_cast<const float4*>(hidT + j * PT);
      float4 h1 = *reinterpret_cast<const float4*>(hidT + j * PT + 4);
      const float2 hv[PT / 2] = {make_float2(h0.x, h0.y), make_float2(h0.z, h0.w), make_float2(h1.x, h1.y),
                                 make_float2(h1.z, h1.w)};
#pragma unroll
      for (int q = 0; q < PT / 2; ++q)
#pragma unroll
        for (int ll = 0; ll < LL; ++ll) G2[q][ll] = __ffma2_rn(hv[q], wv[ll], G2[q][ll]);
    }
    float G[PT][LL];
#pragma unroll
    for (int p = 0; p < PT; ++p)
#pragma unroll
      for (int ll = 0; ll < LL; ++ll) G[p][ll] = (p & 1) ? G2[p / 2][ll].y : G2[p / 2][ll].x;
    __syncwarp();

    // ---- 4. w = G*sigmoid(G); softmax over L; renorm; weighted sum.  Stage-major over the PT items, so the PT
    //      shuffle / division chains of a stage are independent and overlap (same per-item arithmetic).
    {
      float wv[PT][LL], r0[PT], r1[PT];
#pragma unroll
      for (int p = 0; p < PT; ++p) {
        float mx = -CUDART_INF_F;
#pragma unroll
        for (int ll = 0; ll < LL; ++ll) {
          const int l = lane + 32 * ll;
          float gg = gqs[l] * giv[p][ll] + G[p][ll];
          float w = gg * (1.f / (1.f + expf(-gg)));
          wv[p][ll] = w;
          mx = fmaxf(mx, w);
        }
        r0[p] = mx;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int p = 0; p < PT; ++p) r0[p] = fmaxf(r0[p], __shfl_xor_sync(0xffffffffu, r0[p], o));
#pragma unroll
      for (int p = 0; p < PT; ++p) {
        float sum = 0.f;
#pragma unroll
        for (int ll = 0; ll < LL; ++ll) {
          wv[p][ll] = expf(wv[p][ll] - r0[p]);
          sum += wv[p][ll];
        }
        r1[p] = sum;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int p = 0; p < PT; ++p) r1[p] += __shfl_xor_sync(0xffffffffu, r1[p], o);
#pragma unroll
      for (int p = 0; p < PT; ++p) {
        float psum = 0.f;
#pragma unroll
        for (int ll = 0; ll < LL; ++ll) {
          wv[p][ll] = wv[p][ll] / r1[p];
          psum += wv[p][ll];
        }
        r0[p] = psum;
      }
      if (P.renorm) {  // similarity_fn.py:43-45 (dropout is identity in eval, the renorm still runs)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int p = 0; p < PT; ++p) r0[p] += __shfl_xor_sync(0xffffffffu, r0[p], o);
#pragma unroll
        for (int p = 0; p < PT; ++p) {
          const float den = fmaxf(r0[p], P.eps);
#pragma unroll
          for (int ll = 0; ll < LL; ++ll) wv[p][ll] = wv[p][ll] / den;
        }
      }
#pragma unroll
      for (int p = 0; p < PT; ++p) {
        float sc = 0.f;
#pragma unroll
        for (int ll = 0; ll < LL; ++ll) sc = fmaf(wv[p][ll], lg[p][ll], sc);
        r1[p] = sc;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int p = 0; p < PT; ++p) r1[p] += __shfl_xor_sync(0xffffffffu, r1[p], o);
#pragma unroll
      for (int p = 0; p < PT; ++p) {
        const int64_t j = g * PT + p;
        if (lane == 0 && j < P.n_per_query) P.scores[(int64_t)b * P.ld + j] = valid[p] ? r1[p] : -CUDART_INF_F;
      }
    }
    __syncwarp();
    g = gn;  // (aq already points at the next work item's query)
    mine = mine_n;
  }
}

template <int LL, int HH, int NW, bool WS, bool ST>
static int launch_exact_nw(const ExactParams& P, int B, size_t smem, cudaStream_t st) {
  auto kern = exact_scores_kernel<LL, HH, NW, WS, ST>;
  MOL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t groups = (P.n_per_query + PT - 1) / PT;
  const int64_t work = (int64_t)B * groups;  // upper bound (query_flags may deactivate queries)
  // small launches (the rescoring pass of a few queries) spread one work item per block over up to 148 SMs instead of
  // filling the 8 warps of a few blocks: the passes run in parallel and each block's 64 KB weight copy is cheap
  int64_t blocks = work;
  if (blocks > 148) blocks = 148;
  if (blocks < 1) blocks = 1;
  kern<<<(unsigned)blocks, NW * 32, smem, st>>>(P);
  MOL_LAUNCH_CHECK();
  return MOL_OK;
}

template <int LL, int HH>
static int launch_exact_t(const ExactParams& P, int B, cudaStream_t st) {
  const int L = LL * 32, H = HH * 32;
  auto bytes = [&](int nw, bool ws, bool stg) {
    return (size_t)(H + L + nw * (L + P.Pq * (P.d + 4) + (L + H) * PT + (stg ? PT * P.Px * (P.d + 4) : 0)) +
                    (ws ? 2 * L * H : 0)) * sizeof(float);
  };
  const size_t cap = 224 * 1024;
  constexpr int NW = LL <= 2 ? EXACT_NW_SMALL : 8;
  if (bytes(NW, true, true) <= cap) return launch_exact_nw<LL, HH, NW, true, true>(P, B, bytes(NW, true, true), st);
  if (bytes(8, true, false) <= cap) return launch_exact_nw<LL, HH, 8, true, false>(P, B, bytes(8, true, false), st);
  MOL_CHECK_ARG(bytes(8, false, false) <= cap, "exact kernel: shape needs %zu bytes of shared memory", bytes(8, false, false));
  return launch_exact_nw<LL, HH, 8, false, false>(P, B, bytes(8, false, false), st);
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
  P.B = B;
  P.Pq = D.Pq;
  P.Px = D.Px;
  P.d = D.d;
  P.L = D.L;
  P.H = D.H;
  P.temperature = s.temperature;
  P.eps = s.eps;
  P.renorm = s.softmax_renorm;
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
