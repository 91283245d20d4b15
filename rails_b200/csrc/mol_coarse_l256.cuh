// mol_coarse256_kernel: tcgen05 coarse scoring pass for L = 256 logits (P_Q = 16, P_X = 16, d = 64: BASELINE.json config 5,
// "MoL 16x16x64").  Included by mol_coarse_sm100.cu; shares its CoarseParams, operand images and activation code.
//
// Why it is a different kernel.  Per (query, 128-item tile) the L = 64 kernels keep everything of the tile on the SM: the
// item tile (64 KB) in shared memory or TMEM, the logits / hidden units / gates of two queries in TMEM.  At L = 256 the
// item tile is 128 x 1024 fp16 = 256 KB (more than shared memory, all of TMEM), the two weight images are 144 KB, the logits
// alone are 256 fp32 columns.  So:
//   * nothing of the tile is resident.  Its sixteen 64-column boxes (one per item group m) and the four 64-column boxes of
//     its GI rows are streamed through a 4-stage TMA ring for EVERY query (320 KB per (query, tile) from L2: consecutive
//     units of a CTA share the tile, so the corpus itself is read from HBM once per ~7 queries' worth of L2 residency);
//   * the logits are produced and consumed in four QUARTERS of 64 (item groups 4q .. 4q+3): G1 quarter -> E1 (fp16 image
//     A2, 32 columns) -> G2 accumulates the quarter's K = 64 slice into the one 128-column HID; after E2, G3 produces the
//     gate pre-activations quarter by quarter and E3 accumulates sum p and sum p l across the quarters (no max subtraction,
//     as in the L = 64 kernels, so the partial sums simply add);
//   * P_Q = 16 is exactly one MMA's N: G1 needs no zero-padded block-diagonal query image.
// TMEM (512 columns): A2 [0,128) fp16 logits of all four quarters (G2's TS operand, read again by E3) | LOG x2 [128,256)
// | HID [256,384) (fp16 accumulator; A3a in place at +0, ones block at +32, A3b at +64) | GATE x2 [384,512).
// Shared memory: W1 image 68 KB | W2 image 72 KB | ring 4 x 16 KB | query image 2 KB | diag x2 16 KB | 0.5 b1 256 B.
// One chain per SM; within a query the quarters are software-pipelined (LOG / GATE double buffers).  Roles (640 threads):
//   warps 0-3 E3 group, 4-7 E1 + E2 (hidden 0..63), 8-11 E2 (hidden 64..127), 12-15 idle, 16 MMA issuer, 17 TMA producer.
// b1 is added in E2 (half2, from shared memory) because no TMEM column is left for a ones block next to A2; b2 rides on the
// ones block of A3 as in the other kernels.
#pragma once

namespace l256 {

// Walks a CTA's flat range [f0, f1) of (tile, query) units tile by tile, in 32-bit arithmetic (one division in all).
struct Walk32 {
  int f1, bc, tile, qa, qb, qa_next;
  __device__ Walk32(int f0, int f1_, int bc_) : f1(f1_), bc(bc_), qa(0), qb(0) {
    tile = f0 / bc - 1;  // (the only division)
    qa_next = f0 - (tile + 1) * bc;
  }
  __device__ bool next() {
    ++tile;
    qa = qa_next;
    qa_next = 0;
    const int rest = f1 - tile * bc;
    if (rest <= qa) return false;
    qb = rest < bc ? rest : bc;
    return true;
  }
  __device__ int n_mine(int s) const { return (qb - qa + 1 - s) / 2; }
};

constexpr int kThreadsL = 640;
constexpr int kIssuerWarp = 16, kTmaWarp = 17;
constexpr int kRingStages = 4;
constexpr int kBoxBytes = 16384;  // 128 rows x 64 fp16, SWIZZLE_128B

constexpr uint32_t kA2 = 0, kLOG = 128, kHIDc = 256, kGATE = 384;

template <int PQ, int PX, int DD>
struct Cfg256 {
  static_assert(PQ == 16 && PX == 16 && DD == 64, "the L = 256 kernel is written for 16 x 16 x 64");
  static constexpr int L = PQ * PX;             // 256
  static constexpr int K2 = L + 16;             // W1 image row length (the bias block is not multiplied here)
  static constexpr int W1_BYTES = kH * K2 * 2;  // 69 632
  static constexpr int W2_BYTES = L * kK3 * 2;  // 73 728
  static constexpr int Q_BYTES = 16 * DD * 2;   // 2 048: Q_sub / tau of one query, canonical K-major
  static constexpr int D_BYTES = 64 * 64 * 2;   // diag(0.5 gq) of one quarter
  static constexpr int QREC_BYTES = Q_BYTES + L * 2;
  static constexpr int SMEM_BYTES = W1_BYTES + W2_BYTES + kRingStages * kBoxBytes + Q_BYTES + 2 * D_BYTES + 256 + 512 + 1024;
  static_assert(SMEM_BYTES <= kSmemLimit, "shared memory budget exceeded");
};

struct BarsL {
  uint64_t full[kRingStages], empty[kRingStages];
  uint64_t q_ready, g1_done, a2_free, hid_full, e2_done;
  uint64_t log_full[2], e1_done[2], gate_full[2], gate_free[2], diag_ready[2];
  uint32_t tmem_base;
};

template <int PQ, int PX, int DD>
__global__ void __launch_bounds__(kThreadsL, 1)
mol_coarse256_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmGI,
                     const CoarseParams P) {
  using C = Cfg256<PQ, PX, DD>;
  constexpr int L = C::L;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* sRing = smem;                                  // 4 x 16 KB (1024-aligned boxes)
  unsigned char* sW1 = sRing + kRingStages * kBoxBytes;
  unsigned char* sW2 = sW1 + C::W1_BYTES;
  unsigned char* sQ = sW2 + C::W2_BYTES;
  unsigned char* sD = sQ + C::Q_BYTES;                          // 2 x D_BYTES
  uint32_t* sB1 = reinterpret_cast<uint32_t*>(sD + 2 * C::D_BYTES);  // 64 packed pairs of 0.5 b1
  BarsL* bars = reinterpret_cast<BarsL*>(reinterpret_cast<unsigned char*>(sB1) + 256);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time setup
  for (int i = tid; i < C::W1_BYTES / 16; i += kThreadsL)
    reinterpret_cast<uint4*>(sW1)[i] = reinterpret_cast<const uint4*>(P.w1_img)[i];
  for (int i = tid; i < C::W2_BYTES / 16; i += kThreadsL)
    reinterpret_cast<uint4*>(sW2)[i] = reinterpret_cast<const uint4*>(P.w2_img)[i];
  for (int i = tid; i < 2 * C::D_BYTES / 16; i += kThreadsL) reinterpret_cast<uint4*>(sD)[i] = make_uint4(0, 0, 0, 0);
  if (tid < kH)  // 0.5 b1[h] sits in the bias block of the W1 image (column L of row h)
    reinterpret_cast<__half*>(sB1)[tid] = *reinterpret_cast<const __half*>(P.w1_img + nosw_off(tid, L, C::K2));
  if (tid == 0) {
    for (int s = 0; s < kRingStages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    mbar_init(&bars->q_ready, 128);
    mbar_init(&bars->g1_done, 1);
    mbar_init(&bars->a2_free, 128);
    mbar_init(&bars->hid_full, 1);
    mbar_init(&bars->e2_done, 256);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars->log_full[b], 1);
      mbar_init(&bars->e1_done[b], 128);
      mbar_init(&bars->gate_full[b], 1);
      mbar_init(&bars->gate_free[b], 128);
      mbar_init(&bars->diag_ready[b], 128);
    }
    fence_mbar_init();
  }
  if (warp == kIssuerWarp) tmem_alloc<512>(&bars->tmem_base);
  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmGI);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  // this CTA's flat range of (tile, query) units, tile-major
  const int64_t F = (int64_t)(P.tile_end - P.tile_begin) * P.bc;  // (< 2^31: checked by the host)
  const int f0 = (int)(F * blockIdx.x / gridDim.x), f1 = (int)(F * (blockIdx.x + 1) / gridDim.x);
  const int t0 = P.tile_begin;

  if (warp >= 16) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(80));
    if (warp == kTmaWarp) {
      // =============================== TMA producer ===============================
      if (lane == 0) {
        uint32_t i = 0;  // boxes issued
        Walk32 w(f0, f1, P.bc);
        while (w.next()) {
          const int row = P.phys_tile(t0 + w.tile) * kTile;
          for (int q = w.qa; q < w.qb; ++q) {
#pragma unroll 1
            for (int b = 0; b < PX + 4; ++b, ++i) {  // 16 item-group boxes, then the 4 GI quarter boxes
              const uint32_t s = i & (kRingStages - 1);
              mbar_wait_sleep(&bars->empty[s], ((i / kRingStages) & 1u) ^ 1u);
              mbar_arrive_expect_tx(&bars->full[s], kBoxBytes);
              if (b < PX)
                tma_load_2d(sRing + s * kBoxBytes, &tmX, &bars->full[s], b * 64, row);
              else
                tma_load_2d(sRing + s * kBoxBytes, &tmGI, &bars->full[s], (b - PX) * 64, row);
            }
          }
        }
      }
    } else if (warp == kIssuerWarp) {
      // =============================== MMA issuer (converged warp, one elected lane issues) ===============================
      constexpr uint32_t idesc1 = make_idesc_f16(128, 16);
      constexpr uint32_t idesc2 = make_idesc_f16_acc16(128, kH);  // fp16 HID accumulator (read with .pack::16b in E2)
      constexpr uint32_t idesc3 = make_idesc_f16(128, 64);
      const uint32_t sW1a = smem_u32(sW1), sW2a = smem_u32(sW2), sQa = smem_u32(sQ), sDa = smem_u32(sD);
      const uint32_t ring = smem_u32(sRing);
      uint32_t u = 0;   // units done -> parities of the once-per-query barriers
      uint32_t bi = 0;  // ring boxes consumed
      auto g1_quarter = [&](int qtr) __attribute__((always_inline)) {
        const uint32_t lb = tmem + kLOG + (uint32_t)(qtr & 1) * 64u;
        for (int mm = 0; mm < 4; ++mm, ++bi) {
          const uint32_t s = bi & (kRingStages - 1);
          mbar_wait_sleep(&bars->full[s], (bi / kRingStages) & 1u);
          tc_fence_after();
          if (elect_one_sync()) {
#pragma unroll
            for (int ks = 0; ks < DD / 16; ++ks) {
              const uint64_t da = make_smem_desc(ring + s * kBoxBytes + ks * 32, 16, 1024, 2);
              const uint64_t db = make_smem_desc(sQa + ks * 256, 128, (DD / 8) * 128, 0);
              umma_ss(lb + mm * 16, da, db, idesc1, ks > 0);
            }
            umma_commit(&bars->empty[s]);
            if (mm == 3) {
              umma_commit(&bars->log_full[qtr & 1]);
              if (qtr == 3) umma_commit(&bars->g1_done);
            }
          }
          __syncwarp();
        }
      };
      auto g2_quarter = [&](int qtr) __attribute__((always_inline)) {
        mbar_wait_sleep(&bars->e1_done[qtr & 1], (uint32_t)(qtr >> 1) & 1u);
        tc_fence_after();
        if (elect_one_sync()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t db = make_smem_desc(sW1a + (4 * qtr + ks) * 256, 128, (C::K2 / 8) * 128, 0);
            umma_ts(tmem + kHIDc, tmem + kA2 + 32 * qtr + 8 * ks, db, idesc2, (qtr > 0 || ks > 0) ? 1u : 0u);
          }
          if (qtr == 3) umma_commit(&bars->hid_full);
        }
        __syncwarp();
      };
      Walk32 w(f0, f1, P.bc);
      while (w.next()) {
        for (int q = w.qa; q < w.qb; ++q, ++u) {
          const uint32_t pu = u & 1u;
          // ---- phase 1: logits, quarter by quarter (LOG double-buffered: G1(q+1) runs under E1(q))
          mbar_wait_sleep(&bars->q_ready, pu);
          tc_fence_after();
          g1_quarter(0);
          g1_quarter(1);
          g2_quarter(0);
          g1_quarter(2);
          g2_quarter(1);
          g1_quarter(3);
          g2_quarter(2);
          g2_quarter(3);
          // ---- phase 3: gates, quarter by quarter (GATE double-buffered: G3(q+1) runs under E3(q))
          mbar_wait_sleep(&bars->e2_done, pu);
          for (int qtr = 0; qtr < 4; ++qtr, ++bi) {
            const int b = qtr & 1;
            mbar_wait_sleep(&bars->diag_ready[b], (uint32_t)(qtr >> 1) & 1u);
            // GATE[b] must have been read by its previous user: quarter qtr - 2 of this unit, or quarter qtr + 2 of the
            // previous one (two gate_free phases per buffer and unit)
            if (qtr >= 2) {
              mbar_wait_sleep(&bars->gate_free[b], 0u);
            } else if (u > 0) {
              mbar_wait_sleep(&bars->gate_free[b], 1u);
            }
            const uint32_t s = bi & (kRingStages - 1);
            mbar_wait_sleep(&bars->full[s], (bi / kRingStages) & 1u);
            tc_fence_after();
            if (elect_one_sync()) {
              const uint32_t gb = tmem + kGATE + (uint32_t)b * 64u;
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {  // GATE = GI_quarter . diag(0.5 gq_quarter)
                const uint64_t da = make_smem_desc(ring + s * kBoxBytes + ks * 32, 16, 1024, 2);
                const uint64_t db = make_smem_desc(sDa + b * C::D_BYTES + ks * 256, 128, (64 / 8) * 128, 0);
                umma_ss(gb, da, db, idesc3, ks > 0);
              }
              const uint32_t w2q = sW2a + (uint32_t)qtr * 8u * (kK3 / 8) * 128u;  // rows 64 qtr .. of the W2 image
#pragma unroll
              for (int ks = 0; ks < kK3 / 16; ++ks) {  // += [A3a | A3b | 1] . [0.5 W2 | 0.5 b2]^T
                const uint64_t db = make_smem_desc(w2q + ks * 256, 128, (kK3 / 8) * 128, 0);
                // A3a (hidden 0..63) at HID columns [0, 32), A3b at [64, 96), the ones block at [32, 40)
                const uint32_t a3 = ks < 4 ? (uint32_t)ks * 8u : ks < 8 ? 64u + (uint32_t)(ks - 4) * 8u : 32u;
                umma_ts(gb, tmem + kHIDc + a3, db, idesc3, 1u);
              }
              umma_commit(&bars->gate_full[b]);
              umma_commit(&bars->empty[s]);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp >= 12) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(24));  // idle warpgroup
  } else if (warp >= 4) {
    // =============================== E1 (group a only) and E2 (group g: hidden units 64 g ..) ===============================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(72));
    const int g = (warp - 4) >> 2;
    const int r = tid & 127;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tb = tmem + lane_base;
    uint32_t ones[8];
    ones[0] = 0x00003C00u;  // {1.0h, 0}
#pragma unroll
    for (int i = 1; i < 8; ++i) ones[i] = 0u;
    uint4 qv = make_uint4(0, 0, 0, 0);
    uint32_t u = 0;
    bool first = true;
    Walk32 w(f0, f1, P.bc);
    while (w.next()) {
      for (int q = w.qa; q < w.qb; ++q, ++u) {
        const uint32_t pu = u & 1u;
        if (g == 0) {
          // ---- query image of this unit -> shared memory (the previous unit's G1s are complete)
          if (first) qv = __ldg(reinterpret_cast<const uint4*>(P.q_rec + (size_t)q * C::QREC_BYTES) + r);
          if (u > 0) mbar_wait_sleep(&bars->g1_done, pu ^ 1u);
          reinterpret_cast<uint4*>(sQ)[r] = qv;
          fence_proxy_async_smem();
          mbar_arrive(&bars->q_ready);
          {  // prefetch the next unit's image (behind the arrive: see the note in the E3 group)
            int qn = q + 1, tn = w.tile;
            if (qn >= w.qb) {
              qn = 0;
              ++tn;
            }
            if ((int64_t)tn * P.bc + qn < (int64_t)f1) qv = __ldg(reinterpret_cast<const uint4*>(P.q_rec + (size_t)qn * C::QREC_BYTES) + r);
          }
          first = false;
          // ---- E1, quarter by quarter: LOG fp32 -> fp16 -> A2 columns [32 qtr, 32 qtr + 32).  A2 still holds the previous
          //      unit's logits until its E3 has read them all (a2_free)
          if (u > 0) mbar_wait_sleep(&bars->a2_free, pu ^ 1u);
#pragma unroll 1
          for (int qtr = 0; qtr < 4; ++qtr) {
            mbar_wait_sleep(&bars->log_full[qtr & 1], (uint32_t)(qtr >> 1) & 1u);
            tc_fence_after();
            uint32_t la[16], lb[16], pk[16];
            const uint32_t src = tb + kLOG + (uint32_t)(qtr & 1) * 64u;
            tmem_ld_x16(src, la);
            tmem_ld_x16(src + 16, lb);
            tmem_ld_wait_bind16(la);
            tmem_ld_wait_bind16(lb);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              pk[j] = pack_f16x2(__uint_as_float(la[2 * j]), __uint_as_float(la[2 * j + 1]));
              pk[8 + j] = pack_f16x2(__uint_as_float(lb[2 * j]), __uint_as_float(lb[2 * j + 1]));
            }
            tmem_ld_x16(src + 32, la);
            tmem_ld_x16(src + 48, lb);
            tmem_st_x16(tb + kA2 + 32 * qtr, pk);
            tmem_ld_wait_bind16(la);
            tmem_ld_wait_bind16(lb);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              pk[j] = pack_f16x2(__uint_as_float(la[2 * j]), __uint_as_float(la[2 * j + 1]));
              pk[8 + j] = pack_f16x2(__uint_as_float(lb[2 * j]), __uint_as_float(lb[2 * j + 1]));
            }
            tmem_st_x16(tb + kA2 + 32 * qtr + 16, pk);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&bars->e1_done[qtr & 1]);
          }
        }
        // ---- E2 (this group's 64 hidden units): u = HID (fp16, packed) + 0.5 b1 -> silu(2u) -> A3 in place
        mbar_wait_sleep(&bars->hid_full, pu);
        tc_fence_after();
        {
          const uint32_t hid = tb + kHIDc + 64 * g;
          uint32_t p0[8], p1[8];
          tmem_ld_x8_pack16(hid, p0);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t* cur = (c & 1) ? p1 : p0;
            uint32_t* nxt = (c & 1) ? p0 : p1;
            tmem_ld_wait_bind8(cur);
            if (c + 1 < 4) tmem_ld_x8_pack16(hid + 16 * (c + 1), nxt);
#pragma unroll
            for (int j = 0; j < 8; ++j) cur[j] = add_f16x2(cur[j], sB1[32 * g + 8 * c + j]);
            // A3 chunk c of this group -> columns [64 g + 8 c, + 8) of the HID region: columns of the group's own chunk
            // c / 2, already in registers (the two groups run concurrently and never touch each other's columns)
            e2_act_chunk_packed(cur, hid + 8 * c, ((kE2H2Mask >> (4 * g + c)) & 1u) != 0);
          }
          if (g == 0) tmem_st_x8(tb + kHIDc + 32, ones);  // ones block (b2): columns 32..39 = group a's chunk 2, read above
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bars->e2_done);
      }
    }
  } else {
    // =============================== E3 group (+ diag staging) ===============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(168));
    const int r = tid & 127;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tb = tmem + lane_base;
    __half* sDw0 = reinterpret_cast<__half*>(sD + (r < 64 ? nosw_off(r, r, 64) : 0));
    __half* sDw1 = reinterpret_cast<__half*>(sD + C::D_BYTES + (r < 64 ? nosw_off(r, r, 64) : 0));
    const float2 l2e2 = make_float2(kLog2e, kLog2e);
    int map_tile = -1, map_phys = 0;
    uint32_t u = 0;
    __half gq4[4];
    auto load_gq = [&](int q) __attribute__((always_inline)) {
      if (r < 64) {
        const __half* g = reinterpret_cast<const __half*>(P.q_rec + (size_t)q * C::QREC_BYTES + C::Q_BYTES);
#pragma unroll
        for (int i = 0; i < 4; ++i) gq4[i] = g[64 * i + r];
      }
    };
    bool first = true;
    Walk32 w(f0, f1, P.bc);
    while (w.next()) {
      for (int q = w.qa; q < w.qb; ++q, ++u) {
        if (first) load_gq(q);
        first = false;
        // diag of quarters 0 and 1 (both buffers are free: this group has been through gate_full of the previous unit's
        // quarters 2 and 3)
        if (r < 64) {
          *sDw0 = gq4[0];
          *sDw1 = gq4[1];
        }
        fence_proxy_async_smem();
        mbar_arrive(&bars->diag_ready[0]);
        mbar_arrive(&bars->diag_ready[1]);
        const float thr_q = P.thr ? __ldg(P.thr + (size_t)q * P.thr_stride) : -CUDART_INF_F;
        float2 num[4], den[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) num[i] = den[i] = make_float2(0.f, 0.f);
        const __half g2 = gq4[2], g3 = gq4[3];
#pragma unroll 1
        for (int qtr = 0; qtr < 4; ++qtr) {
          const int b = qtr & 1;
          mbar_wait_sleep(&bars->gate_full[b], (uint32_t)(qtr >> 1) & 1u);
          tc_fence_after();
          uint32_t v0[16], v1[16], v2[16], v3[16], pk[32];
          const uint32_t gb = tb + kGATE + (uint32_t)b * 64u;
          tmem_ld_x16(gb, v0);
          tmem_ld_x16(gb + 16, v1);
          tmem_ld_x16(gb + 32, v2);
          tmem_ld_x16(gb + 48, v3);
          tmem_ld_x32(tb + kA2 + 32 * qtr, pk);
          if (qtr < 2) {  // G3 of this quarter is complete: its diag buffer takes quarter qtr + 2
            if (r < 64) *(b ? sDw1 : sDw0) = qtr == 0 ? g2 : g3;
            fence_proxy_async_smem();
          }
          tmem_ld_wait_bind16(v0);
          tmem_ld_wait_bind16(v1);
          tmem_ld_wait_bind16(v2);
          tmem_ld_wait_bind16(v3);
          tmem_ld_wait_bind32(pk);
          tc_fence_before();
          if (qtr < 2) mbar_arrive(&bars->diag_ready[b]);
          mbar_arrive(&bars->gate_free[b]);
          if (qtr == 3) {
            mbar_arrive(&bars->a2_free);
            // prefetch the next unit's 0.5 gq behind the arrives (an mbarrier arrive waits for the thread's outstanding
            // loads: a prefetch in front of one turns its latency into a stall)
            int qn = q + 1, tn = w.tile;
            if (qn >= w.qb) {
              qn = 0;
              ++tn;
            }
            if ((int64_t)tn * P.bc + qn < (int64_t)f1) load_gq(qn);
          }
          auto gate = [&](const uint32_t* v, const uint32_t* lgc) __attribute__((always_inline)) {
#pragma unroll
            for (int j2 = 0; j2 < 8; ++j2) {
              const float2 uu = make_float2(__uint_as_float(v[2 * j2]), __uint_as_float(v[2 * j2 + 1]));
              const float2 a = __fmul2_rn(uu, l2e2);
              const float2 t = make_float2(tanh_approx(uu.x), tanh_approx(uu.y));
              const float2 x = __ffma2_rn(a, t, a);  // w * log2(e), w = silu(2u)
              const float2 e = make_float2(ex2_approx(x.x), ex2_approx(x.y));
              den[j2 & 3] = __fadd2_rn(den[j2 & 3], e);
              num[j2 & 3] = __ffma2_rn(e, __half22float2(*reinterpret_cast<const __half2*>(&lgc[j2])), num[j2 & 3]);
            }
          };
          gate(v0, pk);
          gate(v1, pk + 8);
          gate(v2, pk + 16);
          gate(v3, pk + 24);
        }
        const float2 n2 = __fadd2_rn(__fadd2_rn(num[0], num[1]), __fadd2_rn(num[2], num[3]));
        const float2 d2 = __fadd2_rn(__fadd2_rn(den[0], den[1]), __fadd2_rn(den[2], den[3]));
        const float score = __fdividef(n2.x + n2.y, d2.x + d2.y);
        if (w.tile != map_tile) {
          map_tile = w.tile;
          map_phys = P.phys_tile(t0 + w.tile);
        }
        const int64_t item = (int64_t)map_phys * kTile + r;
        if (item < P.N) {
          if (P.scores) P.scores[(size_t)q * P.ld + ((int64_t)w.tile * kTile + r)] = score;
          if (P.thr && !(score < thr_q)) {  // NaN passes the filter on purpose
            const int pos = atomicAdd(P.cand_cnt + q, 1);
            if (pos < P.cand_cap) {
              P.cand_scores[(size_t)q * P.cand_cap + pos] = score;
              P.cand_idx[(size_t)q * P.cand_cap + pos] = (int32_t)item;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kIssuerWarp) tmem_dealloc<512>(tmem);
}

}  // namespace l256
