// tcgen05 / TMEM / TMA coarse scoring pass for sm_100a (v2: fp16 operands, everything but the
// activations on the tensor pipe).
//
// Computes, for every (query b, item x) pair, the MoL score of
//   rails/similarities/mol/similarity_fn.py:389-405 (sub-embedding dot products / tau),
//   :166-179 (gating: GQ*GI + W2 silu(W1 l + b1) + b2, silu), :42-46 (softmax-weighted sum)
// with fp16 tensor-core operands and fp32 accumulation, fused in ONE kernel: the (B, N, L) logits,
// (B, N, H) hidden activations and (B, N, L) gates never leave the SM.  Its output only RANKS
// candidates; final scores come from the fp32 rescoring pass (mol_exact.cu).
//
// Mapping (one persistent CTA per SM, 640 threads = 5 warpgroups; setmaxnreg splits the register file 160/56/48):
//   per TMEM slot s (2 slots, 256 columns each; slot 0 takes the even queries of a tile's range, slot 1 the odd ones):
//     warps 8s+0..3   E3 warpgroup   : stages query images / diag, gate -> softmax-weighted score -> output (E3)
//     warps 8s+4..7   E1/E2 warpgroup: logits -> fp16 operand (E1), hidden pre-activations -> silu -> fp16 operand (E2)
//   warp 16/17  MMA issuer of slot 0 / 1: the warp runs converged, one elected lane issues tcgen05.mma / commit
//   warp 18     TMA producer: item tile (128 items x P_X*d fp16, SWIZZLE_128B boxes) + the tile's GI rows,
//               double-buffered through full/empty mbarriers.  (warp 19 idle: setmaxnreg needs whole warpgroups.)
// TMEM lanes = the 128 items of the tile; one epilogue thread owns one (query, item) pair, so every
// reduction over the L logits is thread-local (no shuffles).  Per query and slot:
//   G1 (SS): LOG[128 x L]   = X_tile (smem, K = 2d per item-group pair) . Qimg^T   (block-diagonal zero-padded
//            query image so that the N = 16 columns of one MMA belong to ONE query; logit order l' = m*P_Q + n)
//   E1     : LOG -> registers (fp32, kept for the final weighted sum) -> fp16 -> A2 (TMEM, aliases LOG) + a
//            "ones" K-block that folds the bias b1 into G2
//   G2 (TS): HID[128 x 128] = [A2 | 1] . [0.5 W1 | 0.5 b1]^T
//   E2     : u = HID; h = u + u tanh(u) in packed half2 (tanh.approx.f16x2) -> A3 (TMEM) + ones block
//   G3     : GATE[128 x L]  = GI_tile (smem, SS) . diag(0.5 gq)  +  [A3 | 1] (TS) . [0.5 W2 | 0.5 b2]^T, issued in two
//            parts so that it starts while E2 converts the second half of the hidden units
//   E3     : u = GATE (= half the gate pre-activation); w = u + u tanh(u); p = 2^(w log2 e) (no max
//            subtraction: w >= -0.28 and fp32 holds e^w for w < 88; an overflow yields NaN and the
//            caller's safety check falls back); score = sum p*l / sum p.
// Output: the (bc, N) score matrix and/or, fused, a per-query threshold filter that appends
// (score, item) candidates to per-query buffers.
#include <cuda.h>
#include <cuda_fp16.h>
#include <math_constants.h>

#include "mol_coarse.cuh"
#include "sm100_ptx.cuh"

namespace mol {

using namespace sm100;

constexpr int kPQ = 8;          // query groups supported by this kernel
constexpr int kH = 128;         // gating hidden width
constexpr int kTile = 128;      // items per tile (= TMEM lanes)
constexpr int kEpiThreads = 512;               // per TMEM slot: one E1/E3 warpgroup + one E2 warpgroup
constexpr int kCtlWarp0 = kEpiThreads / 32;    // control warpgroup: issuer slot 0, issuer slot 1, TMA producer, (idle)
constexpr int kThreads = kEpiThreads + 4 * 32;
// setmaxnreg split of the 64K-register file (the kernel is compiled for 640 threads -> 96 registers at launch)
// (setmaxnreg only redistributes the CTA's own launch allocation: 640 x 96 = 61440 registers)
#ifndef MOL_HID_F16
#define MOL_HID_F16 1  // measured (round 2): 33.2 -> 32.4 ms per 512 x 1M step, max coarse-vs-exact error 0.023 -> 0.026
#endif
// HID accumulates in fp16 (kind::f16 with an F16 D) and E2 reads it with tcgen05.ld ... .pack::16b: no f32 -> f16x2
// conversions and half the load registers in E2, at the price of five fp16 roundings of the hidden pre-activation
constexpr bool kHidF16 = MOL_HID_F16 != 0;
#ifndef MOL_E2_REGS
#define MOL_E2_REGS 56
#endif
#ifndef MOL_CTL_REGS
#define MOL_CTL_REGS 80  // measured (round 2): 48 made ptxas spill the issuer loops (LDL in front of tcgen05.mma): 35.4 -> 33.2 ms per step
#endif
constexpr int kE2Regs = MOL_E2_REGS, kCtlRegs = MOL_CTL_REGS, kE13Regs = 240 - kE2Regs - kCtlRegs / 2;
static_assert(256 * kE13Regs + 256 * kE2Regs + 128 * kCtlRegs <= kThreads * 96, "register pool over-committed");
constexpr float kLog2e = 1.4426950408889634f;
constexpr int kK3 = kH + 16;    // K of the gate GEMM's TS part: hidden units + the ones block
constexpr int kSmemLimit = 232448;
#ifndef MOL_E2_POLY_MASK
#define MOL_E2_POLY_MASK 0  // (round 1 shipped 0x0E; the half2 form below measured 5.5 % faster)
#endif
// bit c set: chunk c (16 hidden units) of E2 takes tanh from an fp32 odd polynomial on the FMA pipe instead of MUFU.TANH
constexpr unsigned kE2PolyMask = MOL_E2_POLY_MASK;
// finer control: bit (8 * c + j) set = pair j (two hidden units) of chunk c; defaults to whole chunks of MOL_E2_POLY_MASK
#ifdef MOL_E2_POLY_MASK64
constexpr unsigned long long kE2Poly64 = MOL_E2_POLY_MASK64;
#else
constexpr unsigned long long expand_chunk_mask(unsigned m) {
  unsigned long long r = 0;
  for (int c = 0; c < 8; ++c)
    if ((m >> c) & 1u) r |= 0xffull << (8 * c);
  return r;
}
constexpr unsigned long long kE2Poly64 = expand_chunk_mask(kE2PolyMask);
#endif
// tanh(u) ~ c * Q(c^2), c = clamp(u, +-3.75) (weighted least-squares minimax fit, fp32 Horner): |error| <= 6.3e-4
// (MUFU.TANH.F16 itself: ~5e-4)
constexpr float kTanhC = 3.75f;
constexpr float kT0 = 0.996860146522522f, kT1 = -0.3149697482585907f, kT2 = 0.10022653639316559f,
                kT3 = -0.023737963289022446f, kT4 = 0.003820218378677964f, kT5 = -0.0003979474422521889f,
                kT6 = 2.547265285102185e-05f, kT7 = -9.067708219845372e-07f, kT8 = 1.3707315282829313e-08f;

// ---- MUFU-free silu(2u) in packed half2 (design + constants: tools/fit_silu_h2.py; not yet measured on the GPU) ------
//   silu(2u) = (u + |u|) - s(|u|),   s(a) = a (1 - tanh a) in [0, 0.28]  ~  a * w * Q(w),   w = relu(1 - a/A)^4
// u + |u| is exact in fp16 and s is small, so the fp16 evaluation is well conditioned (a polynomial for tanh is not):
// max |error of h| 2.7e-3, rms 3.5e-4 over every fp16 u in [-12, 12] - the MUFU.TANH.F16 path has 2.9e-3 / 2.7e-4 before the
// MUFU's own error.  10 half2 instructions per pair of values after the f32 -> f16x2 pack, none of them on the MUFU.
#ifndef MOL_E2_H2_MASK
#define MOL_E2_H2_MASK 0x3E  // measured (round 2, B200): 33.3 ms vs 35.3 ms per 512 x 1M step, final ids / scores identical
#endif
// bit c set: chunk c (16 hidden units) of E2 uses the half2 form (takes precedence over MOL_E2_POLY_MASK for that chunk)
constexpr unsigned kE2H2Mask = MOL_E2_H2_MASK;
constexpr uint32_t h2x2(unsigned bits16) { return (uint32_t)bits16 * 0x00010001u; }
constexpr uint32_t kH2One = h2x2(0x3C00);
// A = 6, Q of degree 3 (negated coefficients: -0.05252, -0.35986, -1.68262, +1.09277)
constexpr uint32_t kH2NegInvA = h2x2(0xB155);
constexpr uint32_t kH2NC0 = h2x2(0xAAB9), kH2NC1 = h2x2(0xB5C2), kH2NC2 = h2x2(0xBEBB), kH2NC3 = h2x2(0x3C5F);
// a = |u|, aw = a * w, nq = -Q(w)   ->   -s(|u|) = aw * nq
__device__ __forceinline__ void h2_bump_parts(uint32_t u2, uint32_t& a, uint32_t& aw, uint32_t& nq) {
  a = u2 & 0x7fff7fffu;
  uint32_t y = fma_relu_f16x2(a, kH2NegInvA, kH2One);
  y = mul_f16x2(y, y);
  const uint32_t w = mul_f16x2(y, y);
  aw = mul_f16x2(a, w);
  nq = fma_f16x2(kH2NC3, w, kH2NC2);
  nq = fma_f16x2(nq, w, kH2NC1);
  nq = fma_f16x2(nq, w, kH2NC0);
}
// silu(2u) for two packed fp16 values
__device__ __forceinline__ uint32_t silu2_h2(uint32_t u2) {
  uint32_t a, aw, nq;
  h2_bump_parts(u2, a, aw, nq);
  return fma_f16x2(aw, nq, add_f16x2(u2, a));
}
// TMEM column map of one slot (256 columns)
constexpr uint32_t kColLog = 0;     // LOG fp32 [0, L); A2 fp16 aliases [0, L/2) + ones [L/2, L/2 + 8)
constexpr uint32_t kColHid = 64;    // HID fp32 [64, 192); A3 fp16 aliases [64, 128) + ones [128, 136)
constexpr uint32_t kColGate = 192;  // GATE fp32 [192, 192 + L)

template <int PX, int DD>
struct CoarseCfg {
  static constexpr int L = kPQ * PX;
  static constexpr int MG = 16 / kPQ;            // item groups per G1 MMA (2)
  static constexpr int K1 = MG * DD;             // K of G1
  static constexpr int NG = PX / MG;             // G1 MMA groups per query
  static constexpr int K2 = L + 16;              // K of G2: logits + the ones block
  static constexpr int XCOLS = PX * DD;          // fp16 per item row
  static constexpr int XBOXES = XCOLS / 64;      // 128B-swizzle boxes per tile
  static constexpr int X_BYTES = kTile * XCOLS * 2;
  static constexpr int GI_BYTES = kTile * L * 2;
  static constexpr int W1_BYTES = kH * K2 * 2;
  static constexpr int W2_BYTES = L * kK3 * 2;
  static constexpr int Q_BYTES = 16 * K1 * 2;    // query image (16 rows x K1)
  static constexpr int QV = Q_BYTES / 16 / 128;  // uint4 per epilogue thread when staging a query image
  static constexpr int D_BYTES = L * L * 2;      // diag(0.5 gq) image
  static constexpr int QREC_BYTES = Q_BYTES + L * 2;  // per-query record in global memory: image | 0.5 gq (fp16, l' order)
  static constexpr int FIXED = W1_BYTES + W2_BYTES + 2 * Q_BYTES + 2 * D_BYTES + 512 /*barriers*/ + 1024 /*alignment*/;
  static constexpr int STAGES = (2 * (X_BYTES + GI_BYTES) + FIXED <= kSmemLimit) ? 2 : 1;
  static constexpr int SMEM_BYTES = STAGES * (X_BYTES + GI_BYTES) + FIXED;
  static_assert(SMEM_BYTES <= kSmemLimit, "shared memory budget exceeded");
  static_assert(L == 32 || L == 64, "L must be 32 or 64");
  static_assert(QV >= 1, "query image smaller than one uint4 per thread");
};

#ifdef MOL_TRACE
#define TR(role, ev, idx)                                                                        \
  do {                                                                                           \
    if (blockIdx.x == 0 && P.trace && (idx) < 256u && (threadIdx.x & 31) == 0)                    \
      P.trace[((role) * 256 + (idx)) * 8 + (ev)] = clock64();                                     \
  } while (0)
#else
#define TR(role, ev, idx) do {} while (0)
#endif

struct CoarseParams {
  long long* trace;
  const uint8_t* w1_img;
  const uint8_t* w2_img;
  const uint8_t* q_rec;   // (bc, QREC_BYTES)
  float* scores;          // (bc, N) or nullptr
  // fused candidate filter (all nullptr when unused)
  const float* thr;       // per-query threshold thr[q * thr_stride]
  int thr_stride;
  int32_t* cand_cnt;      // (bc) running counters
  float* cand_scores;     // (bc, cand_cap)
  int32_t* cand_idx;      // (bc, cand_cap)
  int cand_cap;
  int64_t N;
  int64_t ld;             // row stride of `scores`
  int tile_begin, tile_end;  // LOGICAL item tiles [tile_begin, tile_end) are scored; scores column = (tile - tile_begin) * 128 + row
  const int32_t* tile_map;   // physical item tile of each logical tile (CoarseOut::tile_map), or nullptr = identity
  int bc;
  __device__ __forceinline__ int phys_tile(int i) const { return tile_map ? __ldg(tile_map + i) : i; }
};

// canonical no-swizzle K-major UMMA layout of an R x K 16-bit matrix (8x8 core matrices, K-adjacent cores contiguous)
__host__ __device__ inline uint32_t nosw_off(int r, int k, int K) {
  return (uint32_t)((r >> 3) * (K >> 3) * 128 + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2);
}

struct Bars {
  uint64_t full[2], empty[2];
  uint64_t q0_ready[2], e1_done[2], a2_read[2], e2a_done[2], e2_done[2], gate_free[2];
  uint64_t log_full[2], hid_full[2], gate_full[2];
  uint32_t tmem_base;
};

// Walks this CTA's flat range of (tile, query) units tile by tile.
struct TileWalk {
  int64_t f, f1;
  int bc;
  int tile, qa, qb;
  bool started;
  __device__ TileWalk(int64_t f0_, int64_t f1_, int bc_)
      : f(f0_), f1(f1_), bc(bc_), tile(0), qa(0), qb(0), started(false) {}
  __device__ bool next() {
    if (f >= f1) return false;
    if (!started) {  // the only division: later tiles start at query 0 of tile + 1
      tile = (int)(f / bc);
      qa = (int)(f - (int64_t)tile * bc);
      started = true;
    } else {
      ++tile;
      qa = 0;
    }
    const int64_t rest = f1 - f;
    const int room = bc - qa;
    qb = rest < (int64_t)room ? qa + (int)rest : bc;
    f += qb - qa;
    return true;
  }
  // queries of slot `wg` in the current tile: qa + wg, qa + wg + 2, ...
  __device__ int n_mine(int wg) const { return (qb - qa + 1 - wg) / 2; }
};

// The flattened (tile, query) sequence of one slot.
struct SlotSeq {
  TileWalk w;
  int wg, q;
  bool in_tile;
  __device__ SlotSeq(int64_t f0, int64_t f1, int bc, int wg_) : w(f0, f1, bc), wg(wg_), q(0), in_tile(false) {}
  __device__ bool next(int& tile, int& query) {
    while (true) {
      if (in_tile && q < w.qb) {
        tile = w.tile;
        query = q;
        q += 2;
        return true;
      }
      if (!w.next()) return false;
      q = w.qa + wg;
      in_tile = true;
    }
  }
};

// One 16-unit chunk of E2: u = 16 fp32 HID values of this thread's item -> h = silu(2u) = u + u tanh(u) as 8 packed fp16 pairs,
// stored to the A3 columns at `taddr`.  bit j2 of `poly`: pair j2 takes tanh from the fp32 polynomial (FMA pipe) instead of
// MUFU.TANH; h2: the whole chunk uses the MUFU-free half2 form.
__device__ __forceinline__ void e2_act_chunk(const uint32_t* v, uint32_t taddr, unsigned poly, bool h2) {
  uint32_t hk[8];
#pragma unroll
  for (int j2 = 0; j2 < 8; ++j2) {
    if (h2) {  // whole chunk MUFU-free in packed half2 (MOL_E2_H2_MASK)
      hk[j2] = silu2_h2(pack_f16x2(__uint_as_float(v[2 * j2]), __uint_as_float(v[2 * j2 + 1])));
    } else if ((poly >> j2) & 1u) {
      const float2 u = make_float2(__uint_as_float(v[2 * j2]), __uint_as_float(v[2 * j2 + 1]));
      const float2 c = make_float2(clamp_sym(u.x, kTanhC), clamp_sym(u.y, kTanhC));
      const float2 s2 = __fmul2_rn(c, c);
      float2 p = __ffma2_rn(make_float2(kT8, kT8), s2, make_float2(kT7, kT7));
      p = __ffma2_rn(p, s2, make_float2(kT6, kT6));
      p = __ffma2_rn(p, s2, make_float2(kT5, kT5));
      p = __ffma2_rn(p, s2, make_float2(kT4, kT4));
      p = __ffma2_rn(p, s2, make_float2(kT3, kT3));
      p = __ffma2_rn(p, s2, make_float2(kT2, kT2));
      p = __ffma2_rn(p, s2, make_float2(kT1, kT1));
      p = __ffma2_rn(p, s2, make_float2(kT0, kT0));
      const float2 t = __fmul2_rn(c, p);
      const float2 h = __ffma2_rn(u, t, u);
      hk[j2] = pack_f16x2(h.x, h.y);
    } else {
      const uint32_t u2 = pack_f16x2(__uint_as_float(v[2 * j2]), __uint_as_float(v[2 * j2 + 1]));
#ifdef MOL_ABLATE_E2
      hk[j2] = fma_f16x2(u2, u2, u2);
#else
      hk[j2] = fma_f16x2(u2, tanh_f16x2(u2), u2);
#endif
    }
  }
  tmem_st_x8(taddr, hk);
}

// The same chunk from eight PACKED fp16 pairs (fp16 HID accumulator read with .pack::16b).
__device__ __forceinline__ void e2_act_chunk_packed(const uint32_t* u2, uint32_t taddr, bool h2) {
  uint32_t hk[8];
#pragma unroll
  for (int j2 = 0; j2 < 8; ++j2) hk[j2] = h2 ? silu2_h2(u2[j2]) : fma_f16x2(u2[j2], tanh_f16x2(u2[j2]), u2[j2]);
  tmem_st_x8(taddr, hk);
}

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
  unsigned char* sW1 = sGI + C::STAGES * C::GI_BYTES;         // 128 x K2  (no-swizzle image)
  unsigned char* sW2 = sW1 + C::W1_BYTES;                     // L x 144
  unsigned char* sQ = sW2 + C::W2_BYTES;                      // 2 x Q_BYTES
  unsigned char* sD = sQ + 2 * C::Q_BYTES;                    // 2 x D_BYTES
  Bars* bars = reinterpret_cast<Bars*>(sD + 2 * C::D_BYTES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time setup
  for (int i = tid; i < C::W1_BYTES / 16; i += kThreads)
    reinterpret_cast<uint4*>(sW1)[i] = reinterpret_cast<const uint4*>(P.w1_img)[i];
  for (int i = tid; i < C::W2_BYTES / 16; i += kThreads)
    reinterpret_cast<uint4*>(sW2)[i] = reinterpret_cast<const uint4*>(P.w2_img)[i];
  for (int i = tid; i < 2 * C::D_BYTES / 16; i += kThreads) reinterpret_cast<uint4*>(sD)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 2);
      mbar_init(&bars->q0_ready[s], 128);
      mbar_init(&bars->e1_done[s], 128);
      mbar_init(&bars->a2_read[s], 128);
      mbar_init(&bars->e2a_done[s], 128);
      mbar_init(&bars->e2_done[s], 128);
      mbar_init(&bars->gate_free[s], 128);
      mbar_init(&bars->log_full[s], 1);
      mbar_init(&bars->hid_full[s], 1);
      mbar_init(&bars->gate_full[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == kCtlWarp0) tmem_alloc<512>(&bars->tmem_base);
  if (warp == kCtlWarp0 + 2 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmGI);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  // this CTA's flat range of (tile, query) units
  const int64_t F = (int64_t)(P.tile_end - P.tile_begin) * P.bc;
  const int64_t f0 = F * blockIdx.x / gridDim.x, f1 = F * (blockIdx.x + 1) / gridDim.x;
  const int t0 = P.tile_begin;

  if (warp >= kCtlWarp0) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kCtlRegs));
  if (warp == kCtlWarp0 + 2) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      TileWalk w(f0, f1, P.bc);
      int it = 0;
      while (w.next()) {
        const int s = it % C::STAGES;
        const uint32_t ph = (uint32_t)(it / C::STAGES) & 1u;
        mbar_wait_sleep(&bars->empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&bars->full[s], C::X_BYTES + C::GI_BYTES);
        const int pt = P.phys_tile(t0 + w.tile);
#pragma unroll
        for (int bx = 0; bx < C::XBOXES; ++bx)
          tma_load_2d(sX + s * C::X_BYTES + bx * 16384, &tmX, &bars->full[s], bx * 64, pt * kTile);
        tma_load_2d(sGI + s * C::GI_BYTES, &tmGI, &bars->full[s], 0, pt * kTile);
        ++it;
      }
    }
  } else if (warp == kCtlWarp0 || warp == kCtlWarp0 + 1) {
    // =============================== MMA issuer of slot `wg` ===============================
    // The whole warp runs this loop converged (all lanes wait, all lanes compute the warp-uniform descriptors);
    // only the tcgen05 instructions themselves are issued by one elected lane.  Issuing from inside an
    // `if (lane == 0)` region makes ptxas wrap every MMA in a uniform-register "waterfall" (~150 clk per MMA,
    // measured), which starved the tensor pipe.
    {
      const int wg = warp - kCtlWarp0;
      constexpr uint32_t idesc1 = make_idesc_f16(128, 16);
      constexpr uint32_t idesc2 = kHidF16 ? make_idesc_f16_acc16(128, kH) : make_idesc_f16(128, kH);
      constexpr uint32_t idesc3 = make_idesc_f16(128, L);
      const uint32_t sW1a = smem_u32(sW1), sW2a = smem_u32(sW2);
      const uint32_t sQa = smem_u32(sQ + wg * C::Q_BYTES), sDa = smem_u32(sD + wg * C::D_BYTES);
      const uint32_t base = tmem + (uint32_t)wg * 256u;
      uint32_t c1 = 0, c2 = 0;  // completed e1_done / e2_done phases of this slot
      bool first = true, pre_g1 = false;

      // G1 of one query from the item tile in stage s, then the commit that announces LOG
      auto issue_g1 = [&](int s) __attribute__((always_inline)) {
        const uint32_t sXa = smem_u32(sX + s * C::X_BYTES);
        if (elect_one_sync()) {
#pragma unroll
          for (int g = 0; g < C::NG; ++g) {
#pragma unroll
#ifdef MOL_ABLATE_G1
            for (int ks = 0; ks < C::K1 / 32; ++ks) {  // (timing experiment: half the K steps)
#else
            for (int ks = 0; ks < C::K1 / 16; ++ks) {
#endif
              const int e = g * C::K1 + ks * 16;  // first fp16 column of this K step in the item row
              const uint64_t da = make_smem_desc(sXa + (e / 64) * 16384 + (e % 64) * 2, 16, 1024, 2);
              const uint64_t db = make_smem_desc(sQa + ks * 256, 128, (C::K1 / 8) * 128, 0);
              umma_ss(base + kColLog + g * 16, da, db, idesc1, ks > 0);
            }
          }
          umma_commit(&bars->log_full[wg]);
        }
        __syncwarp();
      };

      TileWalk w(f0, f1, P.bc);
      int it = 0;
      bool have = w.next();
      while (have) {
        const int s = it % C::STAGES;
        mbar_wait_sleep(&bars->full[s], (uint32_t)(it / C::STAGES) & 1u);
        tc_fence_after();
        const int n = w.n_mine(wg);
        // look ahead: the next tile of this CTA (its first query's G1 is issued early, behind the last G2 of this tile)
        TileWalk wn = w;
        const bool have_next = wn.next();
        const int n_next = have_next ? wn.n_mine(wg) : 0;
        if (n == 0) {
          if (lane == 0) mbar_arrive(&bars->empty[s]);
        } else {
          if (!pre_g1) {
            if (first) {
              mbar_wait_sleep(&bars->q0_ready[wg], 0);
              tc_fence_after();
            }
            issue_g1(s);
          }
          first = false;
          pre_g1 = false;
          const uint32_t sGIa = smem_u32(sGI + s * C::GI_BYTES);
          for (int j = 0; j < n; ++j) {
            // ---- G2 (+ the next query's G1) once E1 has written A2 and staged the next query image.  HID is free:
            //      this warp has already passed e2_done of the previous query.
            if (wg == 0) TR(2, 0, c1);
            mbar_wait_sleep(&bars->e1_done[wg], c1 & 1u);
            if (wg == 0) TR(2, 1, c1);
            ++c1;
            tc_fence_after();
            if (elect_one_sync()) {
#pragma unroll
              for (int ks = 0; ks < C::K2 / 16; ++ks) {
                const uint64_t db = make_smem_desc(sW1a + ks * 256, 128, (C::K2 / 8) * 128, 0);
                umma_ts(base + kColHid, base + kColLog + ks * 8, db, idesc2, ks > 0);
              }
              umma_commit(&bars->hid_full[wg]);
            }
            __syncwarp();
            // the next G1 overwrites LOG / A2 and reads the next query image: the E3 group must have copied the fp16
            // logits of this query out of A2 and staged that image (a2_read)
            if (j + 1 < n || (n_next > 0 && C::STAGES > 1)) mbar_wait_sleep(&bars->a2_read[wg], (c1 - 1u) & 1u);
            int g1_stage = -1;  // smem stage whose item tile the next G1 of this slot reads (-1: none to issue here)
            if (j + 1 < n) {
              g1_stage = s;
            } else if (n_next > 0 && C::STAGES > 1) {  // (single stage: the next tile cannot land before this one is released)
              const int sn = (it + 1) % C::STAGES;
              mbar_wait_sleep(&bars->full[sn], (uint32_t)((it + 1) / C::STAGES) & 1u);
              tc_fence_after();
              g1_stage = sn;
              pre_g1 = true;
            }
            if (g1_stage >= 0) issue_g1(g1_stage);
            if (wg == 0) TR(2, 2, c2);
            // ---- G3, first part: needs the first half of A3 (E2), the diag of this query staged and GATE released
            //      by E3 of the previous query (gate_free; its first phase is arrived by the E1/E3 group's prologue)
            mbar_wait_sleep(&bars->e2a_done[wg], c2 & 1u);
            mbar_wait_sleep(&bars->gate_free[wg], c2 & 1u);
            tc_fence_after();
            if (wg == 0) TR(2, 3, c2);
            if (elect_one_sync()) {
#pragma unroll
#ifdef MOL_ABLATE_DIAG
              for (int ks = 0; ks < 1; ++ks) {  // (timing experiment: one SS k-step only)
#else
              for (int ks = 0; ks < L / 16; ++ks) {  // GATE = GI_tile . diag(0.5 gq)
#endif
                const uint64_t da = (L == 64) ? make_smem_desc(sGIa + ks * 32, 16, 1024, 2)
                                              : make_smem_desc(sGIa + ks * 32, 16, 512, 4);
                const uint64_t db = make_smem_desc(sDa + ks * 256, 128, (L / 8) * 128, 0);
                umma_ss(base + kColGate, da, db, idesc3, ks > 0);
              }
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {  // += A3[:, 0:64] . (0.5 W2[:, 0:64])^T
                const uint64_t db = make_smem_desc(sW2a + ks * 256, 128, (kK3 / 8) * 128, 0);
                umma_ts(base + kColGate, base + kColHid + ks * 8, db, idesc3, 1u);
              }
            }
            __syncwarp();
            if (wg == 0) TR(2, 4, c2);
            // ---- G3, second part, once E2 has written all of A3
            mbar_wait_sleep(&bars->e2_done[wg], c2 & 1u);
            if (wg == 0) TR(2, 5, c2);
            ++c2;
            tc_fence_after();
            if (elect_one_sync()) {
#pragma unroll
              for (int ks = 4; ks < kK3 / 16; ++ks) {  // += [A3[:, 64:128] | 1] . [0.5 W2[:, 64:128] | 0.5 b2]^T
                const uint64_t db = make_smem_desc(sW2a + ks * 256, 128, (kK3 / 8) * 128, 0);
                umma_ts(base + kColGate, base + kColHid + ks * 8, db, idesc3, 1u);
              }
              umma_commit(&bars->gate_full[wg]);
              if (j == n - 1) umma_commit(&bars->empty[s]);  // every MMA of this slot that reads stage s is issued
            }
            __syncwarp();
            if (wg == 0) TR(2, 6, c2 - 1);
          }
        }
        w = wn;
        have = have_next;
        ++it;
      }
    }
  } else if (warp < kCtlWarp0 && ((warp >> 2) & 1) == 1) {
    // =============================== E1 / E2 warpgroup of slot `wg` ===============================
    // E1: LOG (fp32) -> fp16 operand A2 (in place) + ones block.  E2: HID (fp32) -> u -> h = u + u tanh(u) in packed
    // half2 -> A3 (in place) + ones block.
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kE2Regs));
    const int wg = warp >> 3;
    const uint32_t base = tmem + (uint32_t)wg * 256u + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t ones[8];
    ones[0] = 0x00003C00u;  // {1.0h, 0}
#pragma unroll
    for (int i = 1; i < 8; ++i) ones[i] = 0u;
    SlotSeq seq(f0, f1, P.bc, wg);
    int tile = 0, q = 0;
    uint32_t cnt = 0;
    // E1 of the slot's k-th query
    auto do_e1 = [&](uint32_t k) __attribute__((always_inline)) {
      if (warp == 4) TR(1, 4, k);
      mbar_wait_sleep(&bars->log_full[wg], k & 1u);
      tc_fence_after();
      if (warp == 4) TR(1, 5, k);
      {
        uint32_t la[16], lb[16];
        auto conv = [&](const uint32_t* v, uint32_t col) __attribute__((always_inline)) {
          uint32_t pk[8];
#pragma unroll
          for (int j2 = 0; j2 < 8; ++j2)
            pk[j2] = pack_f16x2(__uint_as_float(v[2 * j2]), __uint_as_float(v[2 * j2 + 1]));
          tmem_st_x8(base + col, pk);
        };
        // chunk c (16 fp32 columns) -> A2 columns [8c, 8c + 8): overwrites LOG columns of chunk c / 2 (in registers)
        tmem_ld_x16(base + kColLog, la);
        tmem_ld_x16(base + kColLog + 16, lb);
        tmem_ld_wait_bind16(la);
        tmem_ld_wait_bind16(lb);
        conv(la, kColLog);
        if constexpr (L == 64) tmem_ld_x16(base + kColLog + 32, la);
        conv(lb, kColLog + 8);
        if constexpr (L == 64) {
          tmem_ld_x16(base + kColLog + 48, lb);
          tmem_ld_wait_bind16(la);
          tmem_ld_wait_bind16(lb);
          conv(la, kColLog + 16);
          conv(lb, kColLog + 24);
        }
        tmem_st_x8(base + kColLog + L / 2, ones);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&bars->e1_done[wg]);
      if (warp == 4) TR(1, 0, k);
    };
    bool have = seq.next(tile, q);
    while (have) {
      const bool have_next = seq.next(tile, q);
      do_e1(cnt);
      // ---------------- E2
      mbar_wait_sleep(&bars->hid_full[wg], cnt & 1u);
      tc_fence_after();
      if (warp == 4) TR(1, 1, cnt);
      uint32_t va[16], vb[16];
      auto act = [&](const uint32_t* v, uint32_t col, unsigned poly, bool h2) __attribute__((always_inline)) {
        e2_act_chunk(v, base + col, poly, h2);
      };
      // 8 chunks of 16 hidden units, loads one chunk ahead; A3 chunk c (8 columns) overwrites HID columns that
      // chunk c/2 (already in registers) came from
      if constexpr (kHidF16) {
        static_assert(!kHidF16 || kE2Poly64 == 0, "MOL_HID_F16 supports the MUFU / half2 chunk forms only");
        // packed loads, two chunks ahead (8 registers per chunk)
        uint32_t p0[8], p1[8], p2[8];
        tmem_ld_x8_pack16(base + kColHid, p0);
        tmem_ld_x8_pack16(base + kColHid + 16, p1);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t* cur = (c % 3 == 0) ? p0 : (c % 3 == 1) ? p1 : p2;
          uint32_t* nxt = ((c + 2) % 3 == 0) ? p0 : ((c + 2) % 3 == 1) ? p1 : p2;
          tmem_ld_wait_bind8(cur);
          if (c + 2 < 8) tmem_ld_x8_pack16(base + kColHid + 16 * (c + 2), nxt);
          e2_act_chunk_packed(cur, base + kColHid + 8 * c, ((kE2H2Mask >> c) & 1u) != 0);
          if (c == 3) {
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&bars->e2a_done[wg]);
          }
        }
      } else {
        tmem_ld_x16(base + kColHid, va);
#pragma unroll
        for (int c = 0; c < 8; c += 2) {
          tmem_ld_wait_bind16(va);
          tmem_ld_x16(base + kColHid + 16 * (c + 1), vb);
          act(va, kColHid + 8 * c, (unsigned)((kE2Poly64 >> (8 * c)) & 0xffull), ((kE2H2Mask >> c) & 1u) != 0);
          tmem_ld_wait_bind16(vb);
          if (c + 2 < 8) tmem_ld_x16(base + kColHid + 16 * (c + 2), va);
          act(vb, kColHid + 8 * (c + 1), (unsigned)((kE2Poly64 >> (8 * (c + 1))) & 0xffull),
              ((kE2H2Mask >> (c + 1)) & 1u) != 0);
          if (c == 2) {  // first half of A3 (k < 64) is in TMEM: the issuer may start G3
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&bars->e2a_done[wg]);
            if (warp == 4) TR(1, 2, cnt);
          }
        }
      }
      tmem_st_x8(base + kColHid + 64, ones);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&bars->e2_done[wg]);
      if (warp == 4) TR(1, 3, cnt);
      ++cnt;
      have = have_next;
    }
  } else if (warp < kCtlWarp0) {
    // =============================== E3 warpgroup of slot `wg` (+ query staging) ===============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kE13Regs));
    const int wg = warp >> 3;                 // slot
    const int r = tid & 127;                  // item row within the tile == TMEM lane
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t base = tmem + (uint32_t)wg * 256u + lane_base;
    unsigned char* sQw = sQ + wg * C::Q_BYTES;
    __half* sDw = reinterpret_cast<__half*>(sD + wg * C::D_BYTES + (r < L ? nosw_off(r, r, L) : 0));
    uint32_t cnt = 0;  // queries that went through E1 -> barrier parity

    // prefetched query data: image of the NEXT query to stage, 0.5*gq[r] of the query whose diag is staged next
    uint4 qv[C::QV];
    __half gqv = __float2half(0.f);
    auto load_image = [&](int q) __attribute__((always_inline)) {
      const uint4* src = reinterpret_cast<const uint4*>(P.q_rec + (size_t)q * C::QREC_BYTES);
#pragma unroll
      for (int i = 0; i < C::QV; ++i) qv[i] = __ldg(src + r + i * 128);
    };
    auto load_gq = [&](int q) __attribute__((always_inline)) {
      if (r < L) gqv = reinterpret_cast<const __half*>(P.q_rec + (size_t)q * C::QREC_BYTES + C::Q_BYTES)[r];
    };
    auto store_image = [&]() __attribute__((always_inline)) {
#pragma unroll
      for (int i = 0; i < C::QV; ++i) reinterpret_cast<uint4*>(sQw)[r + i * 128] = qv[i];
    };

    SlotSeq seq(f0, f1, P.bc, wg);
    int tile = 0, q = 0, tile_n = 0, q_n = 0;
    bool have = seq.next(tile, q);
    if (have) {
      load_image(q);
      load_gq(q);
      store_image();
      if (r < L) *sDw = gqv;  // diag of the first query: no G3 has run yet
      fence_proxy_async_smem();
      mbar_arrive(&bars->q0_ready[wg]);
      mbar_arrive(&bars->gate_free[wg]);  // phase 0: GATE is free and the first diag is staged
    }
    bool have_n = have && seq.next(tile_n, q_n);
    if (have_n) {
      load_image(q_n);
      load_gq(q_n);
    }

    uint32_t ones[8];
    ones[0] = 0x00003C00u;  // {1.0h, 0}
#pragma unroll
    for (int i = 1; i < 8; ++i) ones[i] = 0u;
    const float2 l2e2 = make_float2(kLog2e, kLog2e);

    // Stage order of this group: E1(j) -> E3(j-1).  E2(j) runs concurrently in the slot's other warpgroup, the MMAs
    // behind both.  The logits of two queries are live at once, as packed fp16 pairs (pkA / pkB alternate).
    // Start of a step: once E1 (the slot's other warpgroup) has written the fp16 logits of this query to A2, copy them
    // to registers (E3 needs them one step later) and stage the next query image (G1 of this query is complete, so the
    // image buffer is free); a2_read then lets the issuer start the next G1, which overwrites both.
    auto e1 = [&](uint32_t (&pk)[L / 2]) __attribute__((always_inline)) {
      if (warp == 0) TR(0, 0, cnt);
      mbar_wait_sleep(&bars->e1_done[wg], cnt & 1u);
      tc_fence_after();
      if (warp == 0) TR(0, 1, cnt);
      if constexpr (L == 64) {
        tmem_ld_x32(base + kColLog, pk);
      } else {
        tmem_ld_x16(base + kColLog, pk);
      }
      if (have_n) store_image();
      fence_proxy_async_smem();
      if constexpr (L == 64) {
        tmem_ld_wait_bind32(pk);
      } else {
        tmem_ld_wait_bind16(pk);
      }
      tc_fence_before();
      mbar_arrive(&bars->a2_read[wg]);
      if (warp == 0) TR(0, 2, cnt);
    };

    int map_tile = -1, map_phys = 0;
    // E3 of query (tile_p, q_p) (gate_full phase `par`).  Once GATE is in registers it stages the diag of the next
    // query in sequence (if any: its G3 is the next writer of GATE and the next reader of the diag) and releases both.
    // pf_img / pf_gq: queries whose image / 0.5 gq are prefetched from global memory (-1: none).  The loads are issued
    // right BEHIND the gate_free arrive: an mbarrier arrive (release) waits for the thread's outstanding loads, so a
    // prefetch in front of it turns its whole latency into a stall (round-2 stage traces).
    auto e3 = [&](const uint32_t (&pk)[L / 2], int tile_p, int q_p, uint32_t par, bool stage_diag, int pf_img, int pf_gq)
                  __attribute__((always_inline)) {
      float thr_q = -CUDART_INF_F;
      auto prefetch = [&]() __attribute__((always_inline)) {
        if (P.thr) thr_q = __ldg(P.thr + (size_t)q_p * P.thr_stride);  // used after the weighted sum
        if (pf_img >= 0) load_image(pf_img);
        if (pf_gq >= 0) load_gq(pf_gq);
      };
      if (warp == 0) TR(0, 3, cnt - 1u);
      mbar_wait_sleep(&bars->gate_full[wg], par);
      tc_fence_after();
      if (warp == 0) TR(0, 4, cnt - 1u);
      float2 num[4], den[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) num[i] = den[i] = make_float2(0.f, 0.f);
      uint32_t v0[16], v1[16], v2[16];
      auto gate = [&](const uint32_t* v, const uint32_t* lgc) __attribute__((always_inline)) {
#pragma unroll
        for (int j2 = 0; j2 < 8; ++j2) {
          const float2 u = make_float2(__uint_as_float(v[2 * j2]), __uint_as_float(v[2 * j2 + 1]));
          const float2 a = __fmul2_rn(u, l2e2);
#ifdef MOL_ABLATE_E3
          const float2 e = __ffma2_rn(a, u, a);
#else
          const float2 t = make_float2(tanh_approx(u.x), tanh_approx(u.y));
          const float2 x = __ffma2_rn(a, t, a);  // w * log2(e), w = silu(2u), in [-0.41, ...)
          const float2 e = make_float2(ex2_approx(x.x), ex2_approx(x.y));
#endif
          den[j2 & 3] = __fadd2_rn(den[j2 & 3], e);
          num[j2 & 3] = __ffma2_rn(e, __half22float2(*reinterpret_cast<const __half2*>(&lgc[j2])), num[j2 & 3]);
        }
      };
      // GATE in chunks of 16 columns through three buffers; GATE is released once the last chunk has landed
      tmem_ld_x16(base + kColGate, v0);
      tmem_ld_x16(base + kColGate + 16, v1);
      if (stage_diag) {
        if (r < L) *sDw = gqv;  // G3 of this query is complete (gate_full): the diag buffer is free
        fence_proxy_async_smem();
      }
      tmem_ld_wait_bind16(v0);
      tmem_ld_wait_bind16(v1);
      if constexpr (L == 64) {
        tmem_ld_x16(base + kColGate + 32, v2);
        gate(v0, pk);
        tmem_ld_x16(base + kColGate + 48, v0);
        gate(v1, pk + 8);
        tmem_ld_wait_bind16(v2);
        tmem_ld_wait_bind16(v0);
        if (stage_diag) {
          tc_fence_before();
          mbar_arrive(&bars->gate_free[wg]);
        }
        prefetch();
        gate(v2, pk + 16);
        gate(v0, pk + 24);
      } else {
        if (stage_diag) {
          tc_fence_before();
          mbar_arrive(&bars->gate_free[wg]);
        }
        prefetch();
        gate(v0, pk);
        gate(v1, pk + 8);
      }
      const float2 n2 = __fadd2_rn(__fadd2_rn(num[0], num[1]), __fadd2_rn(num[2], num[3]));
      const float2 d2 = __fadd2_rn(__fadd2_rn(den[0], den[1]), __fadd2_rn(den[2], den[3]));
      const float score = __fdividef(n2.x + n2.y, d2.x + d2.y);
      if (warp == 0) TR(0, 5, cnt - 1u);
      if (tile_p != map_tile) {  // (the logical -> physical map has integer divisions: once per tile, not per query)
        map_tile = tile_p;
        map_phys = P.phys_tile(t0 + tile_p);
      }
      const int64_t item = (int64_t)map_phys * kTile + r;
      if (item < P.N) {
        if (P.scores) P.scores[(size_t)q_p * P.ld + ((int64_t)tile_p * kTile + r)] = score;
        if (P.thr && !(score < thr_q)) {  // NaN passes the filter on purpose
          const int pos = atomicAdd(P.cand_cnt + q_p, 1);
          if (pos < P.cand_cap) {
            P.cand_scores[(size_t)q_p * P.cand_cap + pos] = score;
            P.cand_idx[(size_t)q_p * P.cand_cap + pos] = (int32_t)item;
          }
        }
      }
    };

    uint32_t pkA[L / 2], pkB[L / 2];
    int tile_p = 0, q_p = 0;
    bool have_p = false;
    // one step: E1(cur) -> E3(prev); then advance the query window.
    // gqv holds 0.5*gq of the CURRENT query's successor... see the load schedule below:
    //   diag(j) is staged inside E3(j-1) (after gate_full(j-1)), so gq(j) must be in `gqv` at that point: it is
    //   loaded right after diag(j-1) was staged, one step earlier.
    auto step = [&](uint32_t (&pk_cur)[L / 2], const uint32_t (&pk_prev)[L / 2]) __attribute__((always_inline)) {
      e1(pk_cur);
      int tile_nn = 0, q_nn = 0;
      const bool have_nn = have_n && seq.next(tile_nn, q_nn);
      if (have_p) {
        // E3(prev) stages diag(cur) from gqv, which currently holds gq(cur); behind its gate_free arrive it prefetches the
        // image of the query after next (qv was stored to smem inside e1) and gq(next), staged by the next step's E3
        e3(pk_prev, tile_p, q_p, (cnt - 1u) & 1u, true, have_nn ? q_nn : -1, have_n ? q_n : -1);
      } else if (have_nn) {
        load_image(q_nn);
      }
      ++cnt;
      tile_p = tile;
      q_p = q;
      have_p = true;
      tile = tile_n;
      q = q_n;
      have = have_n;
      tile_n = tile_nn;
      q_n = q_nn;
      have_n = have_nn;
    };
    while (have) {
      step(pkA, pkB);
      if (!have) {
        e3(pkA, tile_p, q_p, (cnt - 1u) & 1u, false, -1, -1);
        break;
      }
      step(pkB, pkA);
      if (!have) {
        e3(pkB, tile_p, q_p, (cnt - 1u) & 1u, false, -1, -1);
        break;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kCtlWarp0) tmem_dealloc<512>(tmem);
}


#include "mol_coarse_l256.cuh"

// ------------------------------------------------------------------------------------------------
// Operand images
// ------------------------------------------------------------------------------------------------
// logit order of the kernel: l' = m*PQ + n  <->  reference l = n*PX + m
__device__ __forceinline__ int ref_l(int lp, int PQ, int PX) { return (lp % PQ) * PX + (lp / PQ); }

__device__ __forceinline__ __half to_half_checked(float v, int32_t* overflow) {
  if (!(fabsf(v) <= 65504.f)) atomicOr(overflow, 1);
  return __float2half_rn(v);
}

// W1 image: rows h (N = 128), K = L + 16: [0.5 W1 (l' order) | 0.5 b1 | 0...].
// W2 image: rows l' (N = L), K = 144:     [0.5 W2           | 0.5 b2 | 0...].
__global__ void coarse_weight_images_kernel(const float* __restrict__ w1, const float* __restrict__ b1,
                                            const float* __restrict__ w2, const float* __restrict__ b2,
                                            uint8_t* w1_img, uint8_t* w2_img, int32_t* overflow, int PQ, int PX) {
  const int L = PQ * PX, K2 = L + 16;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kH * K2) {
    const int h = i / K2, k = i % K2;
    float v = 0.f;
    if (k < L) v = 0.5f * w1[h * L + ref_l(k, PQ, PX)];
    else if (k == L) v = 0.5f * b1[h];
    *reinterpret_cast<__half*>(w1_img + nosw_off(h, k, K2)) = to_half_checked(v, overflow);
  }
  if (i < L * kK3) {
    const int lp = i / kK3, k = i % kK3;
    float v = 0.f;
    if (k < kH) v = 0.5f * w2[ref_l(lp, PQ, PX) * kH + k];
    else if (k == kH) v = 0.5f * b2[ref_l(lp, PQ, PX)];
    *reinterpret_cast<__half*>(w2_img + nosw_off(lp, k, kK3)) = to_half_checked(v, overflow);
  }
}

// Per query record: block-diagonal zero-padded image of Q_sub / tau (16 rows x K1) followed by 0.5*gq (fp16, l' order).
__global__ void coarse_query_records_kernel(const float* __restrict__ qsub, const float* __restrict__ gq,
                                            uint8_t* q_rec, int32_t* overflow, const int32_t* __restrict__ weights_overflow,
                                            int bc, int PQ, int PX, int d, float inv_tau) {
  const int MG = 16 / PQ, K1 = MG * d, L = PQ * PX;
  const int q = blockIdx.x;
  if (q >= bc) return;
  if (q == 0 && threadIdx.x == 0 && weights_overflow && *weights_overflow) atomicOr(overflow, 1);
  uint8_t* rec = q_rec + (size_t)q * (16 * K1 * 2 + L * 2);
  for (int i = threadIdx.x; i < 16 * K1; i += blockDim.x) {
    const int row = i / K1, k = i % K1;
    const int mm = row / PQ, n = row % PQ;
    float v = 0.f;
    if (k / d == mm) v = qsub[((size_t)q * PQ + n) * d + (k % d)] * inv_tau;
    *reinterpret_cast<__half*>(rec + nosw_off(row, k, K1)) = to_half_checked(v, overflow);
  }
  __half* g = reinterpret_cast<__half*>(rec + 16 * K1 * 2);
  for (int lp = threadIdx.x; lp < L; lp += blockDim.x)
    g[lp] = to_half_checked(0.5f * gq[(size_t)q * L + ref_l(lp, PQ, PX)], overflow);
}

// gi_half of the index in the kernel's logit order (written by the index build through this kernel)
__global__ void coarse_gi_image_kernel(const float* __restrict__ gi, __half* __restrict__ out, int64_t n,
                                       int32_t* overflow, int PQ, int PX) {
  const int L = PQ * PX;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * L) return;
  const int64_t x = i / L;
  const int lp = (int)(i % L);
  out[i] = to_half_checked(gi[x * L + ref_l(lp, PQ, PX)], overflow);
}

int coarse_gi_image(const mol_shape_t& s, const float* gi_f32, uint16_t* gi_half, int64_t n, int32_t* overflow,
                    cudaStream_t st) {
  Dims D = dims_of(s);
  if (n == 0) return MOL_OK;
  const int64_t total = n * D.L;
  coarse_gi_image_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(gi_f32, reinterpret_cast<__half*>(gi_half),
                                                                          n, overflow, D.Pq, D.Px);
  MOL_LAUNCH_CHECK();
  return MOL_OK;
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
bool coarse_supported(const mol_shape_t& s) {
  Dims D = dims_of(s);
  if (D.H != kH) return false;
  if (D.Pq == 16 && D.Px == 16 && D.d == 64) return true;  // L = 256: mol_coarse256_kernel
  if (D.Pq != kPQ) return false;
  if (D.Px == 8 && D.d == 32) return true;
  if (D.Px == 4 && (D.d == 64 || D.d == 128)) return true;
  return false;
}

static size_t qrec_bytes(const Dims& D) { return (size_t)16 * (16 / D.Pq) * D.d * 2 + (size_t)D.L * 2; }

void coarse_weight_image_bytes(const mol_shape_t& s, size_t* w1_bytes, size_t* w2_bytes) {
  Dims D = dims_of(s);
  *w1_bytes = (size_t)kH * (D.L + 16) * 2;
  *w2_bytes = (size_t)D.L * kK3 * 2;
}

void coarse_plan(const mol_shape_t& s, int chunk, Arena& a, CoarseWs* ws) {
  Dims D = dims_of(s);
  ws->q_rec = a.take<uint8_t>((size_t)chunk * qrec_bytes(D));
  ws->overflow = a.take<int32_t>(1);
}

// weight images (depend on the weights only: once per weight version through mol_weights_prepare, else per call)
int coarse_prepare_weights(const mol_shape_t& s, const mol_weights_t& w, uint8_t* w1_img, uint8_t* w2_img,
                           int32_t* overflow, cudaStream_t st) {
  Dims D = dims_of(s);
  const int n = kH * (D.L + 16) > D.L * kK3 ? kH * (D.L + 16) : D.L * kK3;
  coarse_weight_images_kernel<<<(n + 255) / 256, 256, 0, st>>>(w.qi_w1, w.qi_b1, w.qi_w2, w.qi_b2, w1_img, w2_img,
                                                               overflow, D.Pq, D.Px);
  MOL_LAUNCH_CHECK();
  return MOL_OK;
}

// per-query operand records of a query chunk (once per chunk; every coarse pass over the chunk reads them).  Resets the
// chunk's overflow flag and folds the weight-image overflow flag into it.
int coarse_query_records(const mol_shape_t& s, const CoarseWs& ws, const float* qsub, const float* gq, int bc,
                         const int32_t* weights_overflow, cudaStream_t st) {
  Dims D = dims_of(s);
  if (bc == 0) return MOL_OK;
  MOL_CUDA(cudaMemsetAsync(ws.overflow, 0, sizeof(int32_t), st));
  coarse_query_records_kernel<<<bc, 128, 0, st>>>(qsub, gq, ws.q_rec, ws.overflow, weights_overflow, bc, D.Pq, D.Px, D.d,
                                                  1.0f / s.temperature);
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
  CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d)", (int)r);
    return MOL_ERR_CUDA;
  }
  return MOL_OK;
}

static void* g_trace = nullptr;
#define MOL_STR2(x) #x
#define MOL_STR(x) MOL_STR2(x)
const char* coarse_build_knobs() {
  return "e2poly=" MOL_STR(MOL_E2_POLY_MASK) " e2h2=" MOL_STR(MOL_E2_H2_MASK) " hidf16=" MOL_STR(MOL_HID_F16);
}

void* coarse_trace_buffer() { return g_trace; }
extern "C" void mol_debug_set_trace(void* p) { g_trace = p; }

template <int PX, int DD>
static int launch_coarse(const mol_shape_t& s, const mol_index_t& ix, const CoarseWs& ws, int bc, const CoarseOut& out,
                         cudaStream_t st) {
  using C = CoarseCfg<PX, DD>;
  const int64_t N = ix.num_items;
  const int64_t Np = (N + kTile - 1) / kTile * kTile;
  CUtensorMap tmX, tmGI;
  MOL_TRY(encode_2d(&tmX, ix.xsub_half, C::XCOLS, (uint64_t)Np, 64, kTile, CU_TENSOR_MAP_SWIZZLE_128B));
  MOL_TRY(encode_2d(&tmGI, ix.gi_half, C::L, (uint64_t)Np, C::L, kTile,
                    C::L == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B));
  CoarseParams P;
  P.trace = (long long*)coarse_trace_buffer();
  P.w1_img = ws.w1_img;
  P.w2_img = ws.w2_img;
  P.q_rec = ws.q_rec;
  P.scores = out.scores;
  P.thr = out.thr;
  P.thr_stride = out.thr_stride;
  P.cand_cnt = out.cand_cnt;
  P.cand_scores = out.cand_scores;
  P.cand_idx = out.cand_idx;
  P.cand_cap = out.cand_cap;
  P.N = N;
  P.ld = out.ld;
  P.tile_begin = out.tile_begin;
  P.tile_end = out.tile_end < 0 ? (int)(Np / kTile) : out.tile_end;
  P.tile_map = out.tile_map;
  P.bc = bc;
  if (P.tile_end <= P.tile_begin) return MOL_OK;
  static int sms = 0;  // (set once per process: the attribute call and the device query are not free on a 60 us search)
  if (sms == 0) {
    MOL_CUDA(cudaFuncSetAttribute(mol_coarse_kernel<PX, DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    int dev = 0, n = 148;
    MOL_CUDA(cudaGetDevice(&dev));
    MOL_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    sms = n;
  }
  const int64_t F = (int64_t)(P.tile_end - P.tile_begin) * bc;
  MOL_CHECK_ARG(F < (1ll << 31), "coarse pass: tiles x queries = %lld does not fit 32 bits (use smaller query chunks)", (long long)F);
  int grid = (int)(F < sms ? F : sms);
  if (grid < 1) grid = 1;
  mol_coarse_kernel<PX, DD><<<grid, kThreads, C::SMEM_BYTES, st>>>(tmX, tmGI, P);
  MOL_LAUNCH_CHECK();
  return MOL_OK;
}

// L = 256 (16 x 16 x 64): everything of a (query, tile) unit is streamed, see mol_coarse_l256.cuh
static int launch_coarse256(const mol_index_t& ix, const CoarseWs& ws, int bc, const CoarseOut& out, cudaStream_t st) {
  using C = l256::Cfg256<16, 16, 64>;
  const int64_t N = ix.num_items;
  const int64_t Np = (N + kTile - 1) / kTile * kTile;
  CUtensorMap tmX, tmGI;
  MOL_TRY(encode_2d(&tmX, ix.xsub_half, 16 * 64, (uint64_t)Np, 64, kTile, CU_TENSOR_MAP_SWIZZLE_128B));
  MOL_TRY(encode_2d(&tmGI, ix.gi_half, C::L, (uint64_t)Np, 64, kTile, CU_TENSOR_MAP_SWIZZLE_128B));
  CoarseParams P;
  P.trace = (long long*)coarse_trace_buffer();
  P.w1_img = ws.w1_img;
  P.w2_img = ws.w2_img;
  P.q_rec = ws.q_rec;
  P.scores = out.scores;
  P.thr = out.thr;
  P.thr_stride = out.thr_stride;
  P.cand_cnt = out.cand_cnt;
  P.cand_scores = out.cand_scores;
  P.cand_idx = out.cand_idx;
  P.cand_cap = out.cand_cap;
  P.N = N;
  P.ld = out.ld;
  P.tile_begin = out.tile_begin;
  P.tile_end = out.tile_end < 0 ? (int)(Np / kTile) : out.tile_end;
  P.tile_map = out.tile_map;
  P.bc = bc;
  if (P.tile_end <= P.tile_begin) return MOL_OK;
  static int sms = 0;
  if (sms == 0) {
    MOL_CUDA(cudaFuncSetAttribute(l256::mol_coarse256_kernel<16, 16, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  C::SMEM_BYTES));
    int dev = 0, n = 148;
    MOL_CUDA(cudaGetDevice(&dev));
    MOL_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    sms = n;
  }
  const int64_t F = (int64_t)(P.tile_end - P.tile_begin) * bc;
  MOL_CHECK_ARG(F < (1ll << 31), "coarse pass: tiles x queries = %lld does not fit 32 bits (use smaller query chunks)", (long long)F);
  int grid = (int)(F < sms ? F : sms);
  if (grid < 1) grid = 1;
  l256::mol_coarse256_kernel<16, 16, 64><<<grid, l256::kThreadsL, C::SMEM_BYTES, st>>>(tmX, tmGI, P);
  MOL_LAUNCH_CHECK();
  return MOL_OK;
}

int coarse_run(const mol_shape_t& s, const mol_index_t& ix, const CoarseWs& ws, int bc, const CoarseOut& out,
               cudaStream_t st) {
  Dims D = dims_of(s);
  if (bc == 0 || ix.num_items == 0) return MOL_OK;
  if (D.Pq == 16 && D.Px == 16 && D.d == 64) return launch_coarse256(ix, ws, bc, out, st);
  if (D.Px == 8 && D.d == 32) return launch_coarse<8, 32>(s, ix, ws, bc, out, st);
  if (D.Px == 4 && D.d == 64) return launch_coarse<4, 64>(s, ix, ws, bc, out, st);
  if (D.Px == 4 && D.d == 128) return launch_coarse<4, 128>(s, ix, ws, bc, out, st);
  MOL_CHECK_ARG(false, "tensor-core path does not support this shape");
}

int coarse_scores(const mol_shape_t& s, const mol_index_t& ix, const CoarseWs& ws, int bc, float* scores,
                  cudaStream_t st) {
  CoarseOut out{};
  out.scores = scores;
  out.ld = ix.num_items;
  out.tile_begin = 0;
  out.tile_end = -1;
  return coarse_run(s, ix, ws, bc, out, st);
}

// ------------------------------------------------------------------------------------------------
// Strided sample: logical -> physical tile tables of the threshold pass (tiles 0, stride, 2 stride, ...) and of the
// main pass (every other tile, in order).
// ------------------------------------------------------------------------------------------------
__global__ void tile_maps_kernel(int32_t* __restrict__ sample_map, int32_t* __restrict__ main_map, int tiles, int stride,
                                 int count) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;  // physical tile
  if (t >= tiles) return;
  const bool is_sample = (t % stride == 0) && (t / stride < count);
  if (is_sample) {
    sample_map[t / stride] = t;
  } else {
    const int before = t / stride < count ? t / stride + 1 : count;  // sample tiles in [0, t)
    main_map[t - before] = t;
  }
}

int coarse_tile_maps(int32_t* sample_map, int32_t* main_map, int tiles, int stride, int count, cudaStream_t st) {
  if (tiles <= 0) return MOL_OK;
  tile_maps_kernel<<<(tiles + 255) / 256, 256, 0, st>>>(sample_map, main_map, tiles, stride, count);
  MOL_LAUNCH_CHECK();
  return MOL_OK;
}

// ------------------------------------------------------------------------------------------------
// Survivors of a score matrix (the sample of the threshold pass).
// ------------------------------------------------------------------------------------------------
__global__ void filter_matrix_kernel(const float* __restrict__ scores, int64_t n, int64_t ld, const float* __restrict__ thr,
                                     int thr_stride, int32_t* __restrict__ cnt, float* __restrict__ cand_scores,
                                     int32_t* __restrict__ cand_idx, int cap, int tile_stride) {
  const int b = blockIdx.y;
  const float t = thr[(size_t)b * thr_stride];
  const float* row = scores + (size_t)b * ld;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = row[i];
    if (!(v < t)) {
      const int pos = atomicAdd(cnt + b, 1);
      if (pos < cap) {
        cand_scores[(size_t)b * cap + pos] = v;
        // column -> item: the matrix holds the tiles 0, tile_stride, 2 tile_stride, ... of the corpus (strided sample)
        cand_idx[(size_t)b * cap + pos] = (int32_t)((i >> 7) * tile_stride * 128 + (i & 127));
      }
    }
  }
}

int coarse_filter_matrix(const float* scores, int64_t n, int64_t ld, int bc, const float* thr, int thr_stride,
                         int32_t* cnt, float* cand_scores, int32_t* cand_idx, int cap, int tile_stride,
                         cudaStream_t st) {
  if (bc == 0 || n == 0) return MOL_OK;
  int64_t gx = (n + 1023) / 1024;
  if (gx > 64) gx = 64;
  dim3 grid((unsigned)gx, (unsigned)bc);
  filter_matrix_kernel<<<grid, 256, 0, st>>>(scores, n, ld, thr, thr_stride, cnt, cand_scores, cand_idx, cap,
                                             tile_stride < 1 ? 1 : tile_stride);
  MOL_LAUNCH_CHECK();
  return MOL_OK;
}

// ------------------------------------------------------------------------------------------------
// Safety check: one warp per query.
// ------------------------------------------------------------------------------------------------
// margin of the acceptance test, in units of the largest coarse-vs-exact difference seen on the query's own candidates
// (what this test does and does not prove)
constexpr float kSafetyErrFactor = 3.0f, kSafetyErrAbs = 2e-2f;
__global__ void safety_flags_kernel(const float* __restrict__ cand, const float* __restrict__ exact,
                                    const float* __restrict__ topk, int bc, int kk, int k,
                                    const int32_t* __restrict__ ovf_a, const int32_t* __restrict__ ovf_b,
                                    const int32_t* __restrict__ cnt, const float* __restrict__ thr, int thr_stride,
                                    int cap, int32_t* __restrict__ flags, int32_t* __restrict__ stats, int second) {
  int b = (blockIdx.x * blockDim.x + threadIdx.x) / 32;
  int lane = threadIdx.x % 32;
  if (b >= bc) return;
  if (second && flags[b] == 0) return;  // second chance: only the queries the first test refused
  // candidates that exist: the rest of a row is padding (first test) or unwritten buffer space (second chance)
  int n = kk;
  if (cnt) {
    const int c = cnt[b];
    n = c < kk ? (c < 0 ? 0 : c) : kk;
  }
  float err = 0.f, cmin = CUDART_INF_F;
  for (int j = lane; j < n; j += 32) {
    float c = cand[(int64_t)b * kk + j], e = exact[(int64_t)b * kk + j];
    if (c > -CUDART_INF_F && e > -CUDART_INF_F) {
      err = fmaxf(err, fabsf(c - e));
      cmin = fminf(cmin, c);
    }
    if (c != c) err = CUDART_NAN_F;  // fmaxf would drop a NaN candidate score
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float e2 = __shfl_xor_sync(0xffffffffu, err, o);
    err = (err != err || e2 != e2) ? CUDART_NAN_F : fmaxf(err, e2);
    cmin = fminf(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
  }
  if (lane == 0) {
    float sk = topk[(int64_t)b * k + (k - 1)];
    bool bad = (ovf_a && *ovf_a) || (ovf_b && *ovf_b);
    if (cnt) {
      const int c = cnt[b];
      if (c > cap) bad = true;                              // survivors were dropped
      if (c <= kk) cmin = thr[(size_t)b * thr_stride];      // every survivor is a candidate: outsiders are below thr
      if (stats && !second) {
        if (c > cap) atomicAdd(stats + 1, 1);
        atomicMax(stats + 2, c);
      }
    }
    // NaN-safe: anything but a provable "no" flags the query
    const int f = (!bad && cmin + kSafetyErrFactor * err + kSafetyErrAbs < sk) ? 0 : 1;
    flags[b] = f;
    if (stats) {
      if (second) {
        if (f) atomicAdd(stats + 0, 1);   // still refused: exact fallback
        else atomicAdd(stats + 5, 1);     // accepted on the second chance
      } else if (f && !cnt) {
        atomicAdd(stats + 0, 1);          // matrix strategy: no second chance, straight to the exact fallback
      }
    }
  }
}

int coarse_safety_flags(const float* cand_scores, const float* exact_scores, const float* topk_scores,
                        int bc, int kk, int k, const int32_t* overflow_a, const int32_t* overflow_b,
                        const int32_t* cnt, const float* thr, int thr_stride, int cap, int32_t* flags,
                        int32_t* stats, int second, cudaStream_t st) {
  if (bc == 0) return MOL_OK;
  int threads = bc * 32;
  safety_flags_kernel<<<(threads + 255) / 256, 256, 0, st>>>(cand_scores, exact_scores, topk_scores, bc, kk, k,
                                                             overflow_a, overflow_b, cnt, thr, thr_stride, cap,
                                                             flags, stats, second);
  MOL_LAUNCH_CHECK();
  return MOL_OK;
}

}  // namespace mol
