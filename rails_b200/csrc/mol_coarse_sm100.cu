// tcgen05 / TMEM / TMA coarse scoring pass for sm_100a.
//
// Computes, for every (query b, item x) pair, the MoL score of
//   rails/similarities/mol/similarity_fn.py:389-405 (sub-embedding dot products / tau),
//   :166-179 (gating: GQ*GI + W2 silu(W1 l + b1) + b2, silu), :42-46 (softmax-weighted sum)
// with bf16 tensor-core operands and fp32 accumulation, fused in ONE kernel: the (B, N, L) logits,
// (B, N, H) hidden activations and (B, N, L) gates never leave the SM.  Its (B, N) fp32 output only
// RANKS candidates; final scores come from the fp32 rescoring pass (mol_exact.cu).
//
// Mapping (one persistent CTA per SM, 320 threads):
//   warp 9      TMA producer : item tile (128 items x P_X*d bf16, SWIZZLE_128B boxes) + the tile's GI rows,
//                              double-buffered through full/empty mbarriers.
//   warp 8      MMA issuer   : one elected thread issues every tcgen05.mma and commits to mbarriers.
//   warps 0-3   epilogue warpgroup 0  (TMEM slot 0, even queries of the tile's query range)
//   warps 4-7   epilogue warpgroup 1  (TMEM slot 1, odd queries)
// TMEM lanes = the 128 items of the tile; one epilogue thread owns one (query, item) pair, so every
// reduction over the L logits is thread-local (no shuffles).  Per query and slot:
//   G1 (SS): LOG[128 x L]  = X_tile (smem, K = 2d per item-group pair) . Qimg^T   (block-diagonal zero-padded
//            query image so that N = 16 columns per MMA belong to ONE query; logit order l' = m*P_Q + n)
//   E1     : LOG -> registers (fp32, kept for the final weighted sum) -> bf16 -> A2 (TMEM, aliases LOG)
//   G2 (TS): HID[128 x 128] = A2 . (0.5 W1)^T           E2: u = HID + 0.5 b1; h = u + u tanh(u) -> bf16 -> A3 (TMEM)
//   G3 (TS): GATE[128 x L]  = A3 . (0.5 W2)^T           E3: g = GATE + 0.5 gq*gi + 0.5 b2; w = g + g tanh(g);
//            online softmax over L in chunks of 16 (ex2.approx), score = sum p*l / sum p -> global.
// The two warpgroups run independent query pipelines; the MMA thread serves whichever slot is ready, so
// one group's tensor-core latency hides behind the other group's MUFU/FMA work.
#include <cuda.h>
#include <math_constants.h>

#include "mol_coarse.cuh"
#include "sm100_ptx.cuh"

namespace mol {

using namespace sm100;

constexpr int kPQ = 8;          // query groups supported by this kernel
constexpr int kH = 128;         // gating hidden width
constexpr int kTile = 128;      // items per tile (= TMEM lanes)
constexpr int kThreads = 320;
constexpr float kLog2e = 1.4426950408889634f;

template <int PX, int DD>
struct CoarseCfg {
  static constexpr int L = kPQ * PX;
  static constexpr int MG = 16 / kPQ;            // item groups per G1 MMA (2)
  static constexpr int K1 = MG * DD;             // K of G1
  static constexpr int NG = PX / MG;             // G1 MMA groups per query
  static constexpr int XCOLS = PX * DD;          // bf16 per item row
  static constexpr int XBOXES = XCOLS / 64;      // 128B-swizzle boxes per tile
  static constexpr int X_BYTES = kTile * XCOLS * 2;
  static constexpr int GI_BYTES = kTile * L * 2;
  static constexpr int W_BYTES = kH * L * 2;     // both weight images
  static constexpr int Q_BYTES = 16 * K1 * 2;    // query image (16 rows x K1)
  static constexpr int STAGES = (2 * (X_BYTES + GI_BYTES) + 2 * W_BYTES + 2 * Q_BYTES + 8192 <= 220 * 1024) ? 2 : 1;
  static constexpr int SMEM_BYTES = STAGES * (X_BYTES + GI_BYTES) + 2 * W_BYTES + 2 * Q_BYTES + 4096 + 1024;
  static constexpr uint32_t GI_SWIZZLE_MASK = (L == 64) ? 7u : 3u;  // 128B / 64B swizzle
};

struct CoarseParams {
  const uint8_t* w1_img;
  const uint8_t* w2_img;
  const float* b1h;
  const float* b2h;
  const uint8_t* q_img;  // (bc, Q_BYTES)
  const float* gqh;      // (bc, L)
  float* scores;         // (bc, N)
  int64_t N;
  int n_tiles;
  int bc;
};

// canonical no-swizzle K-major UMMA layout of an R x K bf16 matrix (8x8 core matrices, K-adjacent cores contiguous)
__host__ __device__ inline uint32_t nosw_off(int r, int k, int K) {
  return (uint32_t)((r >> 3) * (K >> 3) * 128 + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2);
}

struct Bars {
  uint64_t full[2], empty[2];
  uint64_t q_ready[2], a2_ready[2], a3_ready[2];
  uint64_t log_full[2], hid_full[2], gate_full[2];
  uint32_t tmem_base;
};

// Walks this CTA's flat range of (tile, query) units tile by tile.
struct TileWalk {
  int64_t f, f1;
  int bc;
  int tile, qa, qb;
  __device__ TileWalk(int64_t f0_, int64_t f1_, int bc_) : f(f0_), f1(f1_), bc(bc_) {}
  __device__ bool next() {
    if (f >= f1) return false;
    tile = (int)(f / bc);
    qa = (int)(f - (int64_t)tile * bc);
    int64_t end = (int64_t)(tile + 1) * bc;
    if (end > f1) end = f1;
    qb = (int)(end - (int64_t)tile * bc);
    f = end;
    return true;
  }
};

template <int PX, int DD>
__global__ void __launch_bounds__(kThreads, 1)
mol_coarse_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmGI,
                  const CoarseParams P) {
  using C = CoarseCfg<PX, DD>;
  constexpr int L = C::L;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* sX = smem;                                   // STAGES x X_BYTES   (1024-aligned boxes)
  unsigned char* sGI = sX + C::STAGES * C::X_BYTES;           // STAGES x GI_BYTES
  unsigned char* sW1 = sGI + C::STAGES * C::GI_BYTES;         // 128 x L  (no-swizzle image)
  unsigned char* sW2 = sW1 + C::W_BYTES;                      // L x 128
  unsigned char* sQ = sW2 + C::W_BYTES;                       // 2 x Q_BYTES
  float* sB1 = reinterpret_cast<float*>(sQ + 2 * C::Q_BYTES);  // 128
  float* sB2 = sB1 + kH;                                      // L
  float* sGQ = sB2 + L;                                       // 2 slots x 2 buffers x L
  Bars* bars = reinterpret_cast<Bars*>(sGQ + 4 * L);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time setup
  for (int i = tid; i < C::W_BYTES / 16; i += kThreads) {
    reinterpret_cast<uint4*>(sW1)[i] = reinterpret_cast<const uint4*>(P.w1_img)[i];
    reinterpret_cast<uint4*>(sW2)[i] = reinterpret_cast<const uint4*>(P.w2_img)[i];
  }
  for (int i = tid; i < kH; i += kThreads) sB1[i] = P.b1h[i];
  for (int i = tid; i < L; i += kThreads) sB2[i] = P.b2h[i];
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 256);
      mbar_init(&bars->q_ready[s], 128);
      mbar_init(&bars->a2_ready[s], 128);
      mbar_init(&bars->a3_ready[s], 128);
      mbar_init(&bars->log_full[s], 1);
      mbar_init(&bars->hid_full[s], 1);
      mbar_init(&bars->gate_full[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc<512>(&bars->tmem_base);
  if (warp == 9 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmGI);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  // this CTA's flat range of (tile, query) units
  const int64_t F = (int64_t)P.n_tiles * P.bc;
  const int64_t f0 = F * blockIdx.x / gridDim.x, f1 = F * (blockIdx.x + 1) / gridDim.x;

  if (warp == 9) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      TileWalk w(f0, f1, P.bc);
      int it = 0;
      while (w.next()) {
        const int s = it % C::STAGES;
        const uint32_t ph = (uint32_t)(it / C::STAGES) & 1u;
        mbar_wait(&bars->empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&bars->full[s], C::X_BYTES + C::GI_BYTES);
#pragma unroll
        for (int bx = 0; bx < C::XBOXES; ++bx)
          tma_load_2d(sX + s * C::X_BYTES + bx * 16384, &tmX, &bars->full[s], bx * 64, w.tile * kTile);
        tma_load_2d(sGI + s * C::GI_BYTES, &tmGI, &bars->full[s], 0, w.tile * kTile);
        ++it;
      }
    }
  } else if (warp == 8) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      constexpr uint32_t idesc1 = make_idesc_bf16(128, 16);
      constexpr uint32_t idesc2 = make_idesc_bf16(128, kH);
      constexpr uint32_t idesc3 = make_idesc_bf16(128, L);
      const uint32_t sW1a = smem_u32(sW1), sW2a = smem_u32(sW2);
      uint32_t cnt[2] = {0, 0};  // queries fully issued (G1) per slot since kernel start -> barrier parity
      uint32_t c2[2] = {0, 0}, c3[2] = {0, 0};
      TileWalk w(f0, f1, P.bc);
      int it = 0;
      while (w.next()) {
        const int s = it % C::STAGES;
        mbar_wait(&bars->full[s], (uint32_t)(it / C::STAGES) & 1u);
        tc_fence_after();
        const uint32_t sXa = smem_u32(sX + s * C::X_BYTES);
        const int nq = w.qb - w.qa;
        const int n[2] = {(nq + 1) / 2, nq / 2};
        int g1[2] = {0, 0}, g2[2] = {0, 0}, g3[2] = {0, 0};
        while (g3[0] < n[0] || g3[1] < n[1]) {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const uint32_t base = tmem + (uint32_t)i * 256u;
            if (g1[i] < n[i] && g1[i] == g2[i] && mbar_try_wait(&bars->q_ready[i], cnt[i] & 1u)) {
              tc_fence_after();
              const uint32_t sQa = smem_u32(sQ + i * C::Q_BYTES);
#pragma unroll
              for (int g = 0; g < C::NG; ++g) {
#pragma unroll
                for (int ks = 0; ks < C::K1 / 16; ++ks) {
                  const int e = g * C::K1 + ks * 16;  // first bf16 column of this K step in the item row
                  const uint64_t da = make_smem_desc(sXa + (e / 64) * 16384 + (e % 64) * 2, 16, 1024, 2);
                  const uint64_t db = make_smem_desc(sQa + ks * 256, 128, (C::K1 / 8) * 128, 0);
                  umma_ss(base + g * 16, da, db, idesc1, ks > 0);
                }
              }
              umma_commit(&bars->log_full[i]);
              ++g1[i];
              ++cnt[i];
            }
            if (g2[i] < g1[i] && mbar_try_wait(&bars->a2_ready[i], c2[i] & 1u)) {
              tc_fence_after();
#pragma unroll
              for (int ks = 0; ks < L / 16; ++ks) {
                const uint64_t db = make_smem_desc(sW1a + ks * 256, 128, (L / 8) * 128, 0);
                umma_ts(base + 64, base + ks * 8, db, idesc2, ks > 0);
              }
              umma_commit(&bars->hid_full[i]);
              ++g2[i];
              ++c2[i];
            }
            if (g3[i] < g2[i] && mbar_try_wait(&bars->a3_ready[i], c3[i] & 1u)) {
              tc_fence_after();
#pragma unroll
              for (int ks = 0; ks < kH / 16; ++ks) {
                const uint64_t db = make_smem_desc(sW2a + ks * 256, 128, (kH / 8) * 128, 0);
                umma_ts(base + 128, base + 64 + ks * 8, db, idesc3, ks > 0);
              }
              umma_commit(&bars->gate_full[i]);
              ++g3[i];
              ++c3[i];
            }
          }
        }
        ++it;
      }
    }
  } else {
    // =============================== epilogue warpgroups ===============================
    const int wg = warp >> 2;                 // slot
    const int r = tid & 127;                  // item row within the tile == TMEM lane
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t base = tmem + (uint32_t)wg * 256u + lane_base;
    unsigned char* sQw = sQ + wg * C::Q_BYTES;
    uint32_t cnt = 0;  // queries processed by this slot
    TileWalk w(f0, f1, P.bc);
    int it = 0;
    while (w.next()) {
      const int s = it % C::STAGES;
      mbar_wait(&bars->full[s], (uint32_t)(it / C::STAGES) & 1u);
      const unsigned char* gi_row = sGI + s * C::GI_BYTES + r * (L * 2);
      const int64_t item = (int64_t)w.tile * kTile + r;
      const int nq = w.qb - w.qa;
      const int n_mine = wg == 0 ? (nq + 1) / 2 : nq / 2;

      auto stage_query = [&](int q, uint32_t buf) {
        const uint4* src = reinterpret_cast<const uint4*>(P.q_img + (size_t)q * C::Q_BYTES);
#pragma unroll
        for (int i = 0; i < C::Q_BYTES / 16 / 128; ++i) reinterpret_cast<uint4*>(sQw)[r + i * 128] = src[r + i * 128];
        if (r < L / 4)
          reinterpret_cast<float4*>(sGQ + (wg * 2 + buf) * L)[r] = reinterpret_cast<const float4*>(P.gqh + (size_t)q * L)[r];
        fence_proxy_async_smem();
        mbar_arrive(&bars->q_ready[wg]);
      };

      if (n_mine > 0) stage_query(w.qa + wg, cnt & 1u);
      for (int j = 0; j < n_mine; ++j) {
        const int q = w.qa + wg + 2 * j;
        const uint32_t par = cnt & 1u;
        // ---------------- E1: logits -> registers, bf16 copy -> A2
        mbar_wait(&bars->log_full[wg], par);
        tc_fence_after();
        uint32_t lg[L];
#pragma unroll
        for (int c = 0; c < L; c += 32) tmem_ld_x32(base + c, lg + c);
        tmem_ld_wait();
        {
          uint32_t pk[L / 2];
#pragma unroll
          for (int j2 = 0; j2 < L / 2; ++j2)
            pk[j2] = pack_bf16x2(__uint_as_float(lg[2 * j2]), __uint_as_float(lg[2 * j2 + 1]));
#pragma unroll
          for (int c = 0; c < L / 2; c += 16) tmem_st_x16(base + c, pk + c);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bars->a2_ready[wg]);
        if (j + 1 < n_mine) stage_query(q + 2, (cnt + 1) & 1u);

        // ---------------- E2: hidden activations -> bf16 -> A3 (in place, first half of HID)
        mbar_wait(&bars->hid_full[wg], par);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld_x32(base + 64 + 32 * c, v);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int j2 = 0; j2 < 16; ++j2) {
            const float2 bb = *reinterpret_cast<const float2*>(sB1 + 32 * c + 2 * j2);
            const float u0 = __uint_as_float(v[2 * j2]) + bb.x;
            const float u1 = __uint_as_float(v[2 * j2 + 1]) + bb.y;
            const float h0 = fmaf(u0, tanh_approx(u0), u0);
            const float h1 = fmaf(u1, tanh_approx(u1), u1);
            pk[j2] = pack_bf16x2(h0, h1);
          }
          tmem_st_x16(base + 64 + 16 * c, pk);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bars->a3_ready[wg]);

        // ---------------- E3: gate -> silu -> online softmax -> weighted sum
        mbar_wait(&bars->gate_full[wg], par);
        tc_fence_after();
        const float* gq = sGQ + (wg * 2 + par) * L;
        float M = -CUDART_INF_F, num = 0.f, den = 0.f;
#pragma unroll
        for (int c = 0; c < L / 16; ++c) {
          uint32_t v[16];
          tmem_ld_x16(base + 128 + 16 * c, v);
          // GI row chunk pair (2c, 2c+1), undoing the TMA swizzle (16-byte chunk index XOR row bits)
          const uint32_t sw = (L == 64) ? (uint32_t)(r & 7) : (uint32_t)((r >> 1) & 3);
          const uint4 ga = *reinterpret_cast<const uint4*>(gi_row + (((uint32_t)(2 * c) ^ sw) << 4));
          const uint4 gb = *reinterpret_cast<const uint4*>(gi_row + (((uint32_t)(2 * c + 1) ^ sw) << 4));
          const uint32_t gw[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
          tmem_ld_wait();
          float wv[16];
          float mc = -CUDART_INF_F;
#pragma unroll
          for (int j2 = 0; j2 < 8; ++j2) {
            const float gi0 = __uint_as_float(gw[j2] << 16);
            const float gi1 = __uint_as_float(gw[j2] & 0xffff0000u);
            const float2 gqv = *reinterpret_cast<const float2*>(gq + 16 * c + 2 * j2);
            const float2 b2v = *reinterpret_cast<const float2*>(sB2 + 16 * c + 2 * j2);
            const float u0 = fmaf(gqv.x, gi0, __uint_as_float(v[2 * j2])) + b2v.x;
            const float u1 = fmaf(gqv.y, gi1, __uint_as_float(v[2 * j2 + 1])) + b2v.y;
            wv[2 * j2] = fmaf(u0, tanh_approx(u0), u0);
            wv[2 * j2 + 1] = fmaf(u1, tanh_approx(u1), u1);
            mc = fmaxf(mc, fmaxf(wv[2 * j2], wv[2 * j2 + 1]));
          }
          const float Mn = fmaxf(M, mc);
          const float sc = ex2_approx((M - Mn) * kLog2e);
          num *= sc;
          den *= sc;
          M = Mn;
          const float off = -Mn * kLog2e;
#pragma unroll
          for (int j2 = 0; j2 < 16; ++j2) {
            const float e = ex2_approx(fmaf(wv[j2], kLog2e, off));
            den += e;
            num = fmaf(e, __uint_as_float(lg[16 * c + j2]), num);
          }
        }
        if (item < P.N) P.scores[(size_t)q * P.N + item] = __fdividef(num, den);
        ++cnt;
      }
      // this warpgroup is done with the tile's smem (GI rows; X was last read by a G1 that completed before E1)
      mbar_arrive(&bars->empty[s]);
      ++it;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------
// Operand images
// ------------------------------------------------------------------------------------------------
// logit order of the kernel: l' = m*PQ + n  <->  reference l = n*PX + m
__device__ __forceinline__ int ref_l(int lp, int PQ, int PX) { return (lp % PQ) * PX + (lp / PQ); }

__global__ void coarse_weight_images_kernel(const float* __restrict__ w1, const float* __restrict__ b1,
                                            const float* __restrict__ w2, const float* __restrict__ b2,
                                            uint8_t* w1_img, uint8_t* w2_img, float* b1h, float* b2h, int PQ,
                                            int PX) {
  const int L = PQ * PX;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kH * L) {
    {  // W1 image: rows h (N = 128), K = L (l')
      const int h = i / L, lp = i % L;
      const float v = 0.5f * w1[h * L + ref_l(lp, PQ, PX)];
      *reinterpret_cast<__nv_bfloat16*>(w1_img + nosw_off(h, lp, L)) = __float2bfloat16_rn(v);
    }
    {  // W2 image: rows l' (N = L), K = 128
      const int lp = i / kH, h = i % kH;
      const float v = 0.5f * w2[ref_l(lp, PQ, PX) * kH + h];
      *reinterpret_cast<__nv_bfloat16*>(w2_img + nosw_off(lp, h, kH)) = __float2bfloat16_rn(v);
    }
  }
  if (i < kH) b1h[i] = 0.5f * b1[i];
  if (i < L) b2h[i] = 0.5f * b2[ref_l(i, PQ, PX)];
}

// Per query: block-diagonal zero-padded image of Q_sub / tau (16 rows x K1) and the permuted 0.5*gq.
__global__ void coarse_query_images_kernel(const float* __restrict__ qsub, const float* __restrict__ gq,
                                           uint8_t* q_img, float* gqh, int bc, int PQ, int PX, int d,
                                           float inv_tau) {
  const int MG = 16 / PQ, K1 = MG * d, L = PQ * PX;
  const int q = blockIdx.x;
  if (q >= bc) return;
  uint8_t* img = q_img + (size_t)q * 16 * K1 * 2;
  for (int i = threadIdx.x; i < 16 * K1; i += blockDim.x) {
    const int row = i / K1, k = i % K1;
    const int mm = row / PQ, n = row % PQ;
    float v = 0.f;
    if (k / d == mm) v = qsub[((size_t)q * PQ + n) * d + (k % d)] * inv_tau;
    *reinterpret_cast<__nv_bfloat16*>(img + nosw_off(row, k, K1)) = __float2bfloat16_rn(v);
  }
  for (int lp = threadIdx.x; lp < L; lp += blockDim.x) gqh[(size_t)q * L + lp] = 0.5f * gq[(size_t)q * L + ref_l(lp, PQ, PX)];
}

// gi_bf16 in the index is stored in the kernel's logit order (written by the index build through this kernel)
__global__ void coarse_gi_image_kernel(const float* __restrict__ gi, __nv_bfloat16* __restrict__ out, int64_t n,
                                       int PQ, int PX) {
  const int L = PQ * PX;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * L) return;
  const int64_t x = i / L;
  const int lp = (int)(i % L);
  out[i] = __float2bfloat16_rn(gi[x * L + ref_l(lp, PQ, PX)]);
}

int coarse_gi_image(const mol_shape_t& s, const float* gi_f32, uint16_t* gi_bf16, int64_t n, cudaStream_t st) {
  Dims D = dims_of(s);
  if (n == 0) return MOL_OK;
  const int64_t total = n * D.L;
  coarse_gi_image_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
      gi_f32, reinterpret_cast<__nv_bfloat16*>(gi_bf16), n, D.Pq, D.Px);
  MOL_LAUNCH_CHECK();
  return MOL_OK;
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
bool coarse_supported(const mol_shape_t& s) {
  Dims D = dims_of(s);
  if (D.Pq != kPQ || D.H != kH) return false;
  if (D.Px == 8 && D.d == 32) return true;
  return false;
}

void coarse_plan(const mol_shape_t& s, int chunk, Arena& a, CoarseWs* ws) {
  Dims D = dims_of(s);
  const int K1 = (16 / D.Pq) * D.d;
  ws->w1_bf16 = reinterpret_cast<__nv_bfloat16*>(a.take<uint8_t>((size_t)kH * D.L * 2));
  ws->w2_bf16 = reinterpret_cast<__nv_bfloat16*>(a.take<uint8_t>((size_t)kH * D.L * 2));
  ws->b1h = a.take<float>(kH);
  ws->b2h = a.take<float>(D.L);
  ws->q_bf16 = reinterpret_cast<__nv_bfloat16*>(a.take<uint8_t>((size_t)chunk * 16 * K1 * 2));
  ws->gqh = a.take<float>((size_t)chunk * D.L);
}

int coarse_prepare(const mol_shape_t& s, const mol_weights_t& w, const CoarseWs& ws, cudaStream_t st) {
  Dims D = dims_of(s);
  const int n = kH * D.L;
  coarse_weight_images_kernel<<<(n + 255) / 256, 256, 0, st>>>(
      w.qi_w1, w.qi_b1, w.qi_w2, w.qi_b2, reinterpret_cast<uint8_t*>(ws.w1_bf16),
      reinterpret_cast<uint8_t*>(ws.w2_bf16), ws.b1h, ws.b2h, D.Pq, D.Px);
  MOL_LAUNCH_CHECK();
  return MOL_OK;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode_2d(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint32_t box_cols,
                     uint32_t box_rows, CUtensorMapSwizzle sw) {
  static PFN_encodeTiled encode = nullptr;
  if (!encode) {
    cudaDriverEntryPointQueryResult qres;
    MOL_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
    MOL_CHECK_ARG(encode != nullptr, "cuTensorMapEncodeTiled not available");
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d)", (int)r);
    return MOL_ERR_CUDA;
  }
  return MOL_OK;
}

template <int PX, int DD>
static int launch_coarse(const mol_shape_t& s, const mol_index_t& ix, const CoarseWs& ws, const float* qsub,
                         const float* gq, int bc, float* scores, cudaStream_t st) {
  using C = CoarseCfg<PX, DD>;
  Dims D = dims_of(s);
  const int64_t N = ix.num_items;
  const int64_t Np = (N + kTile - 1) / kTile * kTile;
  coarse_query_images_kernel<<<bc, 128, 0, st>>>(qsub, gq, reinterpret_cast<uint8_t*>(ws.q_bf16), ws.gqh, bc, D.Pq,
                                                 D.Px, D.d, 1.0f / s.temperature);
  MOL_LAUNCH_CHECK();
  CUtensorMap tmX, tmGI;
  MOL_TRY(encode_2d(&tmX, ix.xsub_bf16, C::XCOLS, (uint64_t)Np, 64, kTile, CU_TENSOR_MAP_SWIZZLE_128B));
  MOL_TRY(encode_2d(&tmGI, ix.gi_bf16, C::L, (uint64_t)Np, C::L, kTile,
                    C::L == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B));
  CoarseParams P;
  P.w1_img = reinterpret_cast<const uint8_t*>(ws.w1_bf16);
  P.w2_img = reinterpret_cast<const uint8_t*>(ws.w2_bf16);
  P.b1h = ws.b1h;
  P.b2h = ws.b2h;
  P.q_img = reinterpret_cast<const uint8_t*>(ws.q_bf16);
  P.gqh = ws.gqh;
  P.scores = scores;
  P.N = N;
  P.n_tiles = (int)(Np / kTile);
  P.bc = bc;
  auto kern = mol_coarse_kernel<PX, DD>;
  MOL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  int dev = 0, sms = 148;
  MOL_CUDA(cudaGetDevice(&dev));
  MOL_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t F = (int64_t)P.n_tiles * bc;
  int grid = (int)(F < sms ? F : sms);
  if (grid < 1) grid = 1;
  kern<<<grid, kThreads, C::SMEM_BYTES, st>>>(tmX, tmGI, P);
  MOL_LAUNCH_CHECK();
  return MOL_OK;
}

int coarse_scores(const mol_shape_t& s, const mol_index_t& ix, const CoarseWs& ws, const float* qsub,
                  const float* gq, int bc, float* scores, cudaStream_t st) {
  Dims D = dims_of(s);
  if (bc == 0 || ix.num_items == 0) return MOL_OK;
  if (D.Px == 8 && D.d == 32) return launch_coarse<8, 32>(s, ix, ws, qsub, gq, bc, scores, st);
  MOL_CHECK_ARG(false, "tensor-core path does not support this shape");
}

// ------------------------------------------------------------------------------------------------
// Safety check: one warp per query.
// ------------------------------------------------------------------------------------------------
__global__ void safety_flags_kernel(const float* __restrict__ cand, const float* __restrict__ exact,
                                    const float* __restrict__ topk, int bc, int kk, int k,
                                    int32_t* __restrict__ flags) {
  int b = (blockIdx.x * blockDim.x + threadIdx.x) / 32;
  int lane = threadIdx.x % 32;
  if (b >= bc) return;
  float err = 0.f, cmin = CUDART_INF_F;
  for (int j = lane; j < kk; j += 32) {
    float c = cand[(int64_t)b * kk + j], e = exact[(int64_t)b * kk + j];
    if (c > -CUDART_INF_F && e > -CUDART_INF_F) {
      err = fmaxf(err, fabsf(c - e));
      cmin = fminf(cmin, c);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    err = fmaxf(err, __shfl_xor_sync(0xffffffffu, err, o));
    cmin = fminf(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
  }
  if (lane == 0) {
    float sk = topk[(int64_t)b * k + (k - 1)];
    // NaN-safe: anything but a provable "no" flags the query for the exact fallback
    flags[b] = (cmin + 1.5f * err + 1e-3f < sk) ? 0 : 1;
  }
}

int coarse_safety_flags(const float* cand_scores, const float* exact_scores, const float* topk_scores,
                        int bc, int kk, int k, int32_t* flags, cudaStream_t st) {
  if (bc == 0) return MOL_OK;
  int threads = bc * 32;
  safety_flags_kernel<<<(threads + 255) / 256, 256, 0, st>>>(cand_scores, exact_scores, topk_scores,
                                                             bc, kk, k, flags);
  MOL_LAUNCH_CHECK();
  return MOL_OK;
}

}  // namespace mol
