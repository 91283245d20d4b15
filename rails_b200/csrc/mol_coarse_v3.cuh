// mol_coarse3_kernel: the X-resident, six-warpgroup form of the tcgen05 coarse scoring pass (included by
// mol_coarse_sm100.cu; shares its operand images, CoarseParams, TileWalk and activation code).
//
// What changed against mol_coarse_kernel, and why (round-1 profile: XU pipe 73 %, tensor pipe 38 % busy, both chains
// of an SM latency-bound; round-2 stage traces: every hand-off between a warp and the tensor pipe costs 300-600 clk):
//   * The item tile lives in TENSOR MEMORY (128 columns, copied once per tile), so G1 is a TS MMA: 16 MMAs of N = 16
//     cost ~160 clk per query instead of ~790 (an SS MMA pays ~40 clk for the 4 KB shared-memory fetch of its A slice
//     whatever N is; a TS MMA costs max(10, N/2)).
//   * Two slots still (TMEM: X 128 | slot 0: LG 64 + HID 128 | slot 1: LG 64 + HID 128 = 512 columns).  LG holds, in
//     turn, the fp32 logits of query c+1 (G1, issued a whole query early) and the gate pre-activations of query c (G3);
//     the fp16 logit operand A2 of G2 lives in SHARED memory (double-buffered per slot), which is what lets the two uses
//     of LG alternate: E1 empties LG long before G3 needs it.  G2 reads A2 as an SS operand (N = 128: same 64 clk per
//     k-step as the TS form).
//   * E2 - the longest stage of a query - is split by COLUMNS over two warpgroups per slot (64 hidden units each; the
//     half is a template parameter so that the per-chunk activation form stays a compile-time choice).  E1 is split the
//     same way and runs in those groups right after E2 of the previous query.  Six epilogue warpgroups (24 warps).
//   * The E3 group no longer carries the logits of two queries in registers: it reads the fp16 logits of its query back
//     from the A2 buffer when the gate arrives.
//   * Control warps keep 80 registers (with 48 ptxas spilled the issuer loops, which cost 5 % on the old kernel).
//
// Warp roles (896 threads):  slot s in {0, 1}:  warps 8s+0..3   E2 group a (X copy, E1 / E2 of hidden units 0..63)
//                                               warps 8s+4..7   E2 group b (X copy, E1 / E2 of hidden units 64..127)
//                                               warps 16+4s..+3 E3 group   (query staging, gate -> score)
//                            warp 24 / 25 MMA issuer of slot 0 / 1, warp 26 TMA producer, warp 27 idle.
// Tensor-pipe order of a slot in steady state:  ... | G3(c-1)  G2(c)  G1(c+1) | G3(c)  G2(c+1)  G1(c+2) | ...
//   G1(c+1) (TS)  LG = X (TMEM) . Qimg^T                   needs gate_free phase c+1: the E3 group has GATE(c-1) in
//                                                          registers and has staged image(c+1) / diag(c)
//   E1(c+1)       LG -> fp16 -> A2[(c+1) & 1] (smem)       after E2(c) in the E2 groups; needs a2_read of that buffer
//   G3(c)         LG = GI_tile . diag(0.5 gq) + [A3 | 1] . W2img^T      needs e2_done(c) and e1_done(c+1) (LG is empty)
//   G2(c+1) (SS)  HID = A2 . W1img^T + ones . b1
//   E2(c+1)       HID half -> silu -> fp16 A3 half (in place; group a adds the ones block)
//   E3(c)         GATE -> registers (gate_free), logits from A2[c & 1] (a2_read) -> softmax-weighted score -> output
#pragma once

namespace v3 {

constexpr int kThreads3 = 6 * 128 + 128;  // 896
constexpr int kCtlWarp = 24;
#ifndef MOL_V3_E2_REGS
#define MOL_V3_E2_REGS 48
#endif
#ifndef MOL_V3_CTL_REGS
#define MOL_V3_CTL_REGS 80
#endif
constexpr int kE2Regs3 = MOL_V3_E2_REGS, kCtlRegs3 = MOL_V3_CTL_REGS;
// pool: 896 threads x 72 registers at launch
constexpr int kE3Regs3 = ((kThreads3 * 72 - 512 * kE2Regs3 - 128 * kCtlRegs3) / 256) / 8 * 8;
static_assert(256 * kE3Regs3 + 512 * kE2Regs3 + 128 * kCtlRegs3 <= kThreads3 * 72, "register pool over-committed");
// setmaxnreg to N registers: .dec below the launch allocation (72 per thread), .inc above it
template <int N>
__device__ __forceinline__ void set_regs() {
  if constexpr (N < 72) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
  } else if constexpr (N > 72) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
  }
}

// Debug build (-DMOL_WATCHDOG): a wait that does not complete within ~2^20 polls records (tag, barrier parity, thread) in
// the trace buffer and raises an abort flag that makes every wait of the grid return, so a deadlocked kernel terminates
// and the stuck waits can be read back (tools/run_watchdog.py).
#ifdef MOL_WATCHDOG
__device__ __forceinline__ void wd_wait(uint64_t* bar, uint32_t parity, int tag, long long* dbg) {
  volatile long long* flag = dbg;
  for (uint32_t spins = 0;; ++spins) {
    if (mbar_try_wait(bar, parity)) return;
    if ((spins & 1023u) == 1023u && *flag != 0) return;
    if (spins > (1u << 20)) {
      const unsigned long long slot = atomicAdd(reinterpret_cast<unsigned long long*>(dbg) + 1, 1ull);
      if (slot < 500) dbg[8 + slot] = ((long long)tag << 48) | ((long long)parity << 40) | ((long long)blockIdx.x << 16) | threadIdx.x;
      *flag = 1;
      __threadfence_system();  // (the buffer is pinned host memory: readable even if the aborted kernel faults later)
      return;
    }
  }
}
#define WAIT(bar, parity, tag) wd_wait(bar, parity, tag, P.trace)
#else
#define WAIT(bar, parity, tag) mbar_wait_sleep(bar, parity)
#endif

constexpr uint32_t kColX = 0;        // item tile, fp16: XCOLS / 2 = 128 columns
constexpr uint32_t kSlot0 = 128;     // slot s at kSlot0 + 192 s
constexpr uint32_t kSlotCols = 192;
constexpr uint32_t kLG = 0;          // within a slot: LOG(c+1) / GATE(c)
constexpr uint32_t kHID = 64;        // HID fp32 [64, 192); A3a [64, 96), ones [96, 104), A3b [128, 160)

struct Bars3 {
  uint64_t xfull[4], xempty[4], gfull[2], gempty[2];
  uint64_t x_free, xt_ready;
  uint64_t log_full[2], e1_done[2], a2_read[2][3], hid_full[2], e2_done[2][2], gate_full[2], lg_free[2], img_ready[2];
  uint32_t tmem_base;
};

template <int PX, int DD>
struct Cfg3 {
  using C = CoarseCfg<PX, DD>;
  static constexpr int L = C::L;
  static constexpr int LH = L / 2;   // fp32 LOG columns converted by one E2 group
  static_assert(C::XCOLS == 256, "the X-resident kernel needs a 128-column item tile (P_X * d == 256)");
  static_assert(C::XBOXES == 4, "one 64-column box of the item tile per E2 warpgroup");
  static constexpr int A2_BYTES = kTile * L * 2;  // fp16 logits of one query, canonical K-major image (K = L)
  static constexpr int ONES_BYTES = kTile * 16 * 2;
  static constexpr int XL_BYTES = 2 * 16384;      // landing zone of the item tile: two 64-column boxes (of four) at a time
  // shared memory: X landing zone | GI (2 stages) | W1 | W2 | Q x2 | D x2 | A2 x6 (three per slot) | ones | barriers
  static constexpr int SMEM_BYTES = XL_BYTES + 2 * C::GI_BYTES + C::W1_BYTES + C::W2_BYTES + 2 * C::Q_BYTES +
                                    2 * C::D_BYTES + 6 * A2_BYTES + ONES_BYTES + 512 + 1024;
  static_assert(SMEM_BYTES <= kSmemLimit, "shared memory budget exceeded");
};

// This CTA's units [f0, f1) of the tile-major (tile, query) sequence, walked tile by tile in 32-bit arithmetic (the host
// keeps tiles x queries below 2^31): queries [qa, qb) of `tile`; slot s takes qa + s, qa + s + 2, ...
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
// the (tile, query) sequence of one slot
struct Seq32 {
  Walk32 w;
  int slot, q;
  __device__ Seq32(int f0, int f1, int bc, int s) : w(f0, f1, bc), slot(s), q(0) {}  // (w.qb == 0: the first next() walks)
  __device__ bool next(int& tile, int& query) {
    q += 2;
    while (q >= w.qb) {
      if (!w.next()) return false;
      q = w.qa + slot;
    }
    tile = w.tile;
    query = q;
    return true;
  }
};

// (c mod 3, c div 3) of a slot's query counter: which of the three A2 buffers, and the phase of its a2_read barrier
struct Tri {
  uint32_t b, u;
  __device__ Tri() : b(0), u(0) {}
  __device__ void inc() {
    if (++b == 3u) {
      b = 0;
      ++u;
    }
  }
};

__device__ __forceinline__ uint4 pack8_f16(const uint32_t* v) {  // 8 fp32 -> 8 fp16 (16 bytes)
  return make_uint4(pack_f16x2(__uint_as_float(v[0]), __uint_as_float(v[1])), pack_f16x2(__uint_as_float(v[2]), __uint_as_float(v[3])),
                    pack_f16x2(__uint_as_float(v[4]), __uint_as_float(v[5])), pack_f16x2(__uint_as_float(v[6]), __uint_as_float(v[7])));
}

// E2 of one column half (G = 0: hidden units 0..63, 1: 64..127): HID fp32 -> silu(2u) -> fp16 A3, in place.  The form
// of each 16-unit chunk (MUFU tanh / fp32 polynomial / half2) is a compile-time choice: chunk 4 G + i of the E2 masks.
// Two phases of two chunks each (the caller converts the next query's logits in between); A3 chunk i (8 columns) lands
// on HID columns whose fp32 values (chunk i / 2) are already in registers.
template <int G, int PHASE>
__device__ __forceinline__ void e2_half(uint32_t hid, const uint32_t* ones) {
  uint32_t va[16], vb[16];
  constexpr int c0 = 2 * PHASE;
  tmem_ld_x16(hid + 16 * c0, va);
  tmem_ld_wait_bind16(va);
  tmem_ld_x16(hid + 16 * (c0 + 1), vb);
  e2_act_chunk(va, hid + 8 * c0, (unsigned)((kE2Poly64 >> (8 * (4 * G + c0))) & 0xffull), ((kE2H2Mask >> (4 * G + c0)) & 1u) != 0);
  tmem_ld_wait_bind16(vb);
  e2_act_chunk(vb, hid + 8 * (c0 + 1), (unsigned)((kE2Poly64 >> (8 * (4 * G + c0 + 1))) & 0xffull),
               ((kE2H2Mask >> (4 * G + c0 + 1)) & 1u) != 0);
  if (G == 0 && PHASE == 1) tmem_st_x8(hid + 32, ones);  // ones block of A3 (b2); HID columns 32..39 were loaded with chunk 2
}

template <int PX, int DD>
__global__ void __launch_bounds__(kThreads3, 1)
mol_coarse3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmGI,
                   const CoarseParams P) {
  using C = CoarseCfg<PX, DD>;
  using V = Cfg3<PX, DD>;
  constexpr int L = C::L, LH = V::LH;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* sX = smem;                                   // 2 x 16 KB: TMA landing zone of the item tile (box b -> buffer b & 1)
  unsigned char* sGI = sX + V::XL_BYTES;                      // 2 x GI_BYTES (SS operand of G3)
  unsigned char* sW1 = sGI + 2 * C::GI_BYTES;                 // 128 x K2 (no-swizzle image; k-step L/16 = the bias block)
  unsigned char* sW2 = sW1 + C::W1_BYTES;                     // L x 144
  unsigned char* sQ = sW2 + C::W2_BYTES;                      // 2 x Q_BYTES (per slot)
  unsigned char* sD = sQ + 2 * C::Q_BYTES;                    // 2 x D_BYTES (per slot)
  unsigned char* sA2 = sD + 2 * C::D_BYTES;                   // [slot][c mod 3] x A2_BYTES
  unsigned char* sOnes = sA2 + 6 * V::A2_BYTES;               // 128 x 16 fp16, column 0 = 1: SS A operand that adds b1
  Bars3* bars = reinterpret_cast<Bars3*>(sOnes + V::ONES_BYTES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time setup
  for (int i = tid; i < C::W1_BYTES / 16; i += kThreads3)
    reinterpret_cast<uint4*>(sW1)[i] = reinterpret_cast<const uint4*>(P.w1_img)[i];
  for (int i = tid; i < C::W2_BYTES / 16; i += kThreads3)
    reinterpret_cast<uint4*>(sW2)[i] = reinterpret_cast<const uint4*>(P.w2_img)[i];
  for (int i = tid; i < 2 * C::D_BYTES / 16; i += kThreads3) reinterpret_cast<uint4*>(sD)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < V::ONES_BYTES / 16; i += kThreads3)  // row r of the ones tile: 16 bytes at (r >> 3) * 256 + (r & 7) * 16
    reinterpret_cast<uint4*>(sOnes)[i] = ((i >> 3) & 1) ? make_uint4(0, 0, 0, 0) : make_uint4(0x00003C00u, 0, 0, 0);
  if (tid == 0) {
    mbar_init(&bars->x_free, 2);
    mbar_init(&bars->xt_ready, 512);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->xfull[s], 1);
      mbar_init(&bars->xfull[s + 2], 1);
      mbar_init(&bars->xempty[s], 128);
      mbar_init(&bars->xempty[s + 2], 128);
      mbar_init(&bars->gfull[s], 1);
      mbar_init(&bars->gempty[s], 2);
      mbar_init(&bars->log_full[s], 1);
      mbar_init(&bars->e1_done[s], 256);
      mbar_init(&bars->a2_read[s][0], 128);
      mbar_init(&bars->a2_read[s][1], 128);
      mbar_init(&bars->a2_read[s][2], 128);
      mbar_init(&bars->hid_full[s], 1);
      mbar_init(&bars->e2_done[s][0], 128);
      mbar_init(&bars->e2_done[s][1], 128);
      mbar_init(&bars->gate_full[s], 1);
      mbar_init(&bars->lg_free[s], 128);
      mbar_init(&bars->img_ready[s], 128);
    }
    fence_mbar_init();
  }
  if (warp == kCtlWarp) tmem_alloc<512>(&bars->tmem_base);
  if (warp == kCtlWarp + 2 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmGI);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  // this CTA's flat range of (tile, query) units
  const int64_t F = (int64_t)(P.tile_end - P.tile_begin) * P.bc;  // (< 2^31: checked by the host)
  const int f0 = (int)(F * blockIdx.x / gridDim.x), f1 = (int)(F * (blockIdx.x + 1) / gridDim.x);
  const int t0 = P.tile_begin;

  if (warp >= kCtlWarp) {
    set_regs<kCtlRegs3>();
    if (warp == kCtlWarp + 2) {
      // =============================== TMA producer ===============================
      if (lane == 0) {
        Walk32 w(f0, f1, P.bc);
        int it = 0;
        while (w.next()) {
          const int s = it & 1;
          const int pt = P.phys_tile(t0 + w.tile);
          WAIT(&bars->gempty[s], ((uint32_t)(it >> 1) & 1u) ^ 1u, 22);
          mbar_arrive_expect_tx(&bars->gfull[s], C::GI_BYTES);
          tma_load_2d(sGI + s * C::GI_BYTES, &tmGI, &bars->gfull[s], 0, pt * kTile);
          // the item tile in four 64-column boxes through two 16 KB landing buffers (box b -> buffer b & 1; one full /
          // empty barrier pair per BOX, one phase per tile): box b is loaded once the box that used its buffer before it
          // (b - 2 of this tile, or b + 2 of the previous tile) has been copied into TMEM
#pragma unroll
          for (int bx = 0; bx < C::XBOXES; ++bx) {
            if (bx >= 2) {
              WAIT(&bars->xempty[bx - 2], (uint32_t)it & 1u, 23);
            } else if (it > 0) {
              WAIT(&bars->xempty[bx + 2], (uint32_t)(it - 1) & 1u, 24);
            }
            mbar_arrive_expect_tx(&bars->xfull[bx], 16384);
            tma_load_2d(sX + (bx & 1) * 16384, &tmX, &bars->xfull[bx], bx * 64, pt * kTile);
          }
          ++it;
        }
      }
    } else if (warp < kCtlWarp + 2) {
      // =============================== MMA issuer of slot `wg` ===============================
      // (whole warp converged; one elected lane issues the tcgen05 instructions)
      const int wg = warp - kCtlWarp;
      constexpr uint32_t idesc1 = make_idesc_f16(128, 16);
      constexpr uint32_t idesc2 = make_idesc_f16(128, kH);
      constexpr uint32_t idesc3 = make_idesc_f16(128, L);
      const uint32_t sW1a = smem_u32(sW1), sW2a = smem_u32(sW2);
      const uint32_t sQa = smem_u32(sQ + wg * C::Q_BYTES), sDa = smem_u32(sD + wg * C::D_BYTES);
      const uint32_t sA2a = smem_u32(sA2 + wg * 3 * V::A2_BYTES);
      const uint64_t dOnes = make_smem_desc(smem_u32(sOnes), 128, 256, 0);
      const uint32_t base = tmem + kSlot0 + (uint32_t)wg * kSlotCols;
      const uint32_t xt = tmem + kColX;
      uint64_t* const log_full = &bars->log_full[wg];
      uint64_t* const e1_done = &bars->e1_done[wg];
      uint64_t* const gate_full = &bars->gate_full[wg];
      uint64_t* const lg_free = &bars->lg_free[wg];
      uint64_t* const img_ready = &bars->img_ready[wg];
      // G1 of the slot's query number `cq`.  Needs its image staged (img_ready phase cq) and LG empty: issued EARLY
      // (ahead of G3 of query cq - 1 in pipe order) the last occupant was GATE(cq - 2); at the top of a tile G3(cq - 1) is
      // already in the pipe and it is GATE(cq - 1).  `last` = last query of this slot in its tile.
      auto issue_g1 = [&](uint32_t cq, bool last, bool top) __attribute__((always_inline)) {
        WAIT(img_ready, cq & 1u, 1);
        if (top) {
          if (cq >= 1) WAIT(lg_free, (cq - 1u) & 1u, 2);
        } else {
          if (cq >= 2) WAIT(lg_free, cq & 1u, 3);
        }
        tc_fence_after();
        if (elect_one_sync()) {
#pragma unroll
          for (int g = 0; g < C::NG; ++g) {
#pragma unroll
            for (int ks = 0; ks < C::K1 / 16; ++ks) {
              const uint64_t db = make_smem_desc(sQa + ks * 256, 128, (C::K1 / 8) * 128, 0);
              umma_ts(base + kLG + g * 16, xt + (uint32_t)((g * C::K1 + ks * 16) / 2), db, idesc1, ks > 0);
            }
          }
          umma_commit(log_full);
          if (last) umma_commit(&bars->x_free);  // every G1 of this slot that reads the tile is issued
        }
        __syncwarp();
      };
      uint32_t c = 0;  // queries of this slot so far -> barrier parities
      Tri t3;          // (c mod 3, c div 3)
      Walk32 w(f0, f1, P.bc);
      int it = 0;
      while (w.next()) {
        const int s = it & 1;
        const int n = w.n_mine(wg);
        // (also with no query of this slot in the tile: the waits keep its hand-offs below in step with the tile - one
        //  arrival per barrier phase)
        WAIT(&bars->gfull[s], (uint32_t)(it >> 1) & 1u, 4);  // GI rows of the tile (G3's SS operand)
        WAIT(&bars->xt_ready, (uint32_t)it & 1u, 5);         // X of the tile is in TMEM
        tc_fence_after();
        if (n == 0) {
          if (lane == 0) {
            mbar_arrive(&bars->x_free);
            mbar_arrive(&bars->gempty[s]);
          }
          __syncwarp();
          ++it;
          continue;
        }
        const uint32_t sGIa = smem_u32(sGI + s * C::GI_BYTES);
        issue_g1(c, n == 1, true);  // first query of the tile: not early (X had to arrive first)
        for (int j = 0; j < n; ++j, ++c, t3.inc()) {
          const uint32_t par = c & 1u;
          // ---- G2(c): A2 is in shared memory (HID is free: G3 of the previous query is ahead in pipe order)
          WAIT(e1_done, par, 6);
          tc_fence_after();
          if (wg == 0) TR(2, 0, c);
          if (elect_one_sync()) {
            const uint32_t a2 = sA2a + t3.b * V::A2_BYTES;
#pragma unroll
            for (int ks = 0; ks < L / 16; ++ks) {
              const uint64_t da = make_smem_desc(a2 + ks * 256, 128, (L / 8) * 128, 0);
              const uint64_t db = make_smem_desc(sW1a + ks * 256, 128, (C::K2 / 8) * 128, 0);
              umma_ss(base + kHID, da, db, idesc2, ks > 0);
            }
            {  // + 0.5 b1: ones tile times the bias k-block of the W1 image
              const uint64_t db = make_smem_desc(sW1a + (L / 16) * 256, 128, (C::K2 / 8) * 128, 0);
              umma_ss(base + kHID, dOnes, db, idesc2, 1u);
            }
            umma_commit(&bars->hid_full[wg]);
          }
          __syncwarp();
          // ---- G1(c+1), a whole query early (same tile only: the next tile's X is not in TMEM yet)
          if (j + 1 < n) issue_g1(c + 1, j + 2 == n, false);
          if (wg == 0) TR(2, 1, c);
          // ---- G3(c), in two parts.  Part 1 (the GI . diag term and the first two 16-unit chunks of each A3 half) goes
          //      out as soon as LG is empty: GATE(c-1) in registers and diag(c) staged (lg_free phase c - 1), LOG(c+1)
          //      converted (e1_done(c+1), which the E2 groups signal BETWEEN the halves of E2(c): their first two chunks
          //      are in TMEM by then).  Part 2 follows e2_done.  Without a next query in the tile both parts wait e2_done.
          if (c >= 1) WAIT(lg_free, par ^ 1u, 7);
          if (j + 1 < n) {
            WAIT(e1_done, par ^ 1u, 8);
          } else {
            WAIT(&bars->e2_done[wg][0], par, 9);
            WAIT(&bars->e2_done[wg][1], par, 10);
          }
          tc_fence_after();
          if (wg == 0) TR(2, 2, c);
          if (elect_one_sync()) {
#pragma unroll
            for (int ks = 0; ks < L / 16; ++ks) {  // GATE = GI_tile . diag(0.5 gq)
              const uint64_t da = (L == 64) ? make_smem_desc(sGIa + ks * 32, 16, 1024, 2)
                                            : make_smem_desc(sGIa + ks * 32, 16, 512, 4);
              const uint64_t db = make_smem_desc(sDa + ks * 256, 128, (L / 8) * 128, 0);
              umma_ss(base + kLG, da, db, idesc3, ks > 0);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
              for (int i = 0; i < 2; ++i) {  // += A3 chunks 0, 1 of half h (hidden units 64 h + 0..31)
                const int ks = 4 * h + i;
                const uint64_t db = make_smem_desc(sW2a + ks * 256, 128, (kK3 / 8) * 128, 0);
                umma_ts(base + kLG, base + kHID + 64 * h + i * 8, db, idesc3, 1u);
              }
            }
          }
          __syncwarp();
          if (j + 1 < n) {
            WAIT(&bars->e2_done[wg][0], par, 11);
            WAIT(&bars->e2_done[wg][1], par, 12);
            tc_fence_after();
          }
          if (elect_one_sync()) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
              for (int i = 2; i < 4; ++i) {  // += A3 chunks 2, 3 of half h (hidden units 64 h + 32..63)
                const int ks = 4 * h + i;
                const uint64_t db = make_smem_desc(sW2a + ks * 256, 128, (kK3 / 8) * 128, 0);
                umma_ts(base + kLG, base + kHID + 64 * h + i * 8, db, idesc3, 1u);
              }
            }
            {  // + 0.5 b2 (ones block written by E2 group a)
              const uint64_t db = make_smem_desc(sW2a + 8 * 256, 128, (kK3 / 8) * 128, 0);
              umma_ts(base + kLG, base + kHID + 32, db, idesc3, 1u);
            }
            umma_commit(gate_full);
            if (j == n - 1) umma_commit(&bars->gempty[s]);  // last MMA of this slot that reads the stage's GI rows
          }
          __syncwarp();
          if (wg == 0) TR(2, 3, c);
        }
        ++it;
      }
    }
  } else if (warp < 16) {
    // =============================== E2 group g of slot `wg`: X copy, E1 half, E2 half ===============================
    // (warps 0..15; the E3 groups take the HIGHER warp ids 16..23: the sub-partition's arbiter prefers the higher warp id,
    //  and the E3 groups are the longer chain)
    set_regs<kE2Regs3>();
    const int wg = warp >> 3;
    const int g = (warp >> 2) & 1;  // column half
    const int box = wg * 2 + g;            // which 64-column box of the item tile this group copies into TMEM
    const int r = tid & 127;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t base = tmem + kSlot0 + (uint32_t)wg * kSlotCols + lane_base;
    const uint32_t xt = tmem + kColX + lane_base;
    uint64_t* const log_full = &bars->log_full[wg];
    uint64_t* const e1_done = &bars->e1_done[wg];
    uint64_t* const hid_full = &bars->hid_full[wg];
    uint64_t* const e2_done = &bars->e2_done[wg][g];
    // this thread's row of the A2 images: 16-byte k-chunk kc at + kc * 128
    unsigned char* const a2row = sA2 + wg * 3 * V::A2_BYTES + (r >> 3) * (L >> 3) * 128 + (r & 7) * 16;
    uint32_t ones[8];
    ones[0] = 0x00003C00u;  // {1.0h, 0}
#pragma unroll
    for (int i = 1; i < 8; ++i) ones[i] = 0u;
    uint32_t c = 0;
    Tri t3n;  // A2 buffer (and a2_read phase) of the query whose E1 comes next
    Walk32 w(f0, f1, P.bc);
    int it = 0;
    while (w.next()) {
      // ---- tile switch: this group's box of the item tile, landing zone -> TMEM.  Every G1 of the previous tile (both
      //      slots) is complete (x_free).
      WAIT(&bars->xfull[box], (uint32_t)it & 1u, 13);
      if (it > 0) {
        WAIT(&bars->x_free, (uint32_t)(it - 1) & 1u, 15);
        tc_fence_after();
      }
      {
        const unsigned char* xrow = sX + (box & 1) * 16384 + r * 128;
        uint32_t v[32];
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {  // undo the 128B TMA swizzle: 16-byte chunk index XOR (row & 7)
          const uint4 t = *reinterpret_cast<const uint4*>(xrow + ((ch ^ (r & 7)) << 4));
          v[4 * ch] = t.x;
          v[4 * ch + 1] = t.y;
          v[4 * ch + 2] = t.z;
          v[4 * ch + 3] = t.w;
        }
        tmem_st_x32(xt + box * 32, v);
        mbar_arrive(&bars->xempty[box]);  // (the values are in registers: the buffer may take another box)
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bars->xt_ready);
      }
      const int n = w.n_mine(wg);
      // E1 (this group's half of the logits) of the slot's query number `cq`: LG fp32 -> fp16 -> A2[cq & 1] in shared memory
      auto e1 = [&](uint32_t cq) __attribute__((always_inline)) {
        const uint32_t pq = cq & 1u;
        if (wg == 0 && (warp & 3) == 0) TR(1 + 2 * g, 0, cq);
        WAIT(log_full, pq, 16);
        if (cq >= 3) WAIT(&bars->a2_read[wg][t3n.b], (t3n.u - 1u) & 1u, 17);  // E3 of query cq - 3 has read the buffer
        tc_fence_after();
        if (wg == 0 && (warp & 3) == 0) TR(1 + 2 * g, 1, cq);
        {
          unsigned char* dst = a2row + t3n.b * V::A2_BYTES + g * (LH / 8) * 128;
          uint32_t la[16];
          tmem_ld_x16(base + kLG + g * LH, la);
          tmem_ld_wait_bind16(la);
          if constexpr (LH == 32) {
            uint32_t lb[16];
            tmem_ld_x16(base + kLG + g * LH + 16, lb);
            *reinterpret_cast<uint4*>(dst) = pack8_f16(la);
            *reinterpret_cast<uint4*>(dst + 128) = pack8_f16(la + 8);
            tmem_ld_wait_bind16(lb);
            *reinterpret_cast<uint4*>(dst + 256) = pack8_f16(lb);
            *reinterpret_cast<uint4*>(dst + 384) = pack8_f16(lb + 8);
          } else {
            *reinterpret_cast<uint4*>(dst) = pack8_f16(la);
            *reinterpret_cast<uint4*>(dst + 128) = pack8_f16(la + 8);
          }
        }
        fence_proxy_async_smem();  // A2 is read by the tensor core (async proxy)
        tmem_st_wait();            // (called between the halves of E2: the first two A3 chunks are complete, G3 part 1 reads them)
        tc_fence_before();
        mbar_arrive(e1_done);
        t3n.inc();
        if (wg == 0 && (warp & 3) == 0) TR(1 + 2 * g, 2, cq);
      };
      if (n > 0) e1(c);  // first query of the tile (its G1 could not be issued early)
      for (int j = 0; j < n; ++j, ++c) {
        const uint32_t par = c & 1u;
        // ---------------- E2 (this group's 64 hidden units), with E1 of the NEXT query between its two halves: LOG(c+1) is
        //                  in LG by then (G1 is issued a query early), and G3(c) - which overwrites LG - can start as
        //                  soon as E2(c) ends
        WAIT(hid_full, par, 18);
        tc_fence_after();
        if (wg == 0 && (warp & 3) == 0) TR(1 + 2 * g, 3, c);
        if (g == 0) {
          e2_half<0, 0>(base + kHID, ones);
        } else {
          e2_half<1, 0>(base + kHID + 64, ones);
        }
        if (j + 1 < n) e1(c + 1);
        if (g == 0) {
          e2_half<0, 1>(base + kHID, ones);
        } else {
          e2_half<1, 1>(base + kHID + 64, ones);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(e2_done);
        if (wg == 0 && (warp & 3) == 0) TR(1 + 2 * g, 4, c);
      }
      ++it;
    }
  } else {
    // =============================== E3 group of slot `wg` (+ query staging) ===============================
    set_regs<kE3Regs3>();
    const int wg = (warp - 16) >> 2;
    const int r = tid & 127;  // item row within the tile == TMEM lane
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t base = tmem + kSlot0 + (uint32_t)wg * kSlotCols + lane_base;
    unsigned char* sQw = sQ + wg * C::Q_BYTES;
    __half* sDw = reinterpret_cast<__half*>(sD + wg * C::D_BYTES + (r < L ? nosw_off(r, r, L) : 0));
    const unsigned char* const a2row = sA2 + wg * 3 * V::A2_BYTES + (r >> 3) * (L >> 3) * 128 + (r & 7) * 16;
    uint64_t* const log_full = &bars->log_full[wg];
    uint64_t* const gate_full = &bars->gate_full[wg];
    uint64_t* const lg_free = &bars->lg_free[wg];
    uint64_t* const img_ready = &bars->img_ready[wg];

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

    // Staging schedule (c = this slot's query counter):
    //   img_ready phase p = image(p) is in the slot's image buffer: phase 0 in the prologue (with diag(0)), phase 1 once
    //     G1(0) is complete (log_full phase 0), phase c + 2 inside E3 of query c once G1(c + 1) is complete.
    //   lg_free phase c (inside E3 of query c) = GATE(c) and the logits of query c are in registers, diag(c + 1) is staged.
    Seq32 seq(f0, f1, P.bc, wg);
    int tile = 0, q = 0, tile_n = 0, q_n = 0, tile_nn = 0, q_nn = 0;
    bool have = seq.next(tile, q);
    if (have) {
      load_image(q);
      load_gq(q);
      store_image();
      if (r < L) *sDw = gqv;
      fence_proxy_async_smem();
      mbar_arrive(img_ready);  // phase 0
    }
    bool have_n = have && seq.next(tile_n, q_n);
    bool have_nn = have_n && seq.next(tile_nn, q_nn);
    if (have) {
      if (have_n) {
        load_image(q_n);
        load_gq(q_n);  // gqv = 0.5 gq of query 1 (its diag is staged inside E3 of query 0)
        WAIT(log_full, 0, 19);  // G1(0) has read image(0)
        store_image();
        fence_proxy_async_smem();
        mbar_arrive(img_ready);  // phase 1
      }
      if (have_nn) load_image(q_nn);
    }
    const float2 l2e2 = make_float2(kLog2e, kLog2e);
    int map_tile = -1, map_phys = 0;
    uint32_t c = 0;
    Tri t3;
    while (have) {
      const uint32_t par = c & 1u;
      if (warp == 16) TR(0, 0, c);
      // ---- gate -> registers; fp16 logits of this query back from A2[par]
      WAIT(gate_full, par, 20);
      tc_fence_after();
      if (warp == 16) TR(0, 1, c);
      float2 num[4], den[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) num[i] = den[i] = make_float2(0.f, 0.f);
      uint32_t v0[16], v1[16], v2[16], v3[16];
      const unsigned char* lsrc = a2row + t3.b * V::A2_BYTES;
      // one 16-logit chunk of the gate: u = GATE, w = silu(2u), p = 2^(w log2 e); the chunk's fp16 logits come back from
      // the A2 buffer in shared memory (two 16-byte loads), so the group keeps no logits in registers between stages
      auto gate = [&](const uint32_t* v, int chunk) __attribute__((always_inline)) {
        const uint4 la = *reinterpret_cast<const uint4*>(lsrc + (2 * chunk) * 128);
        const uint4 lb = *reinterpret_cast<const uint4*>(lsrc + (2 * chunk + 1) * 128);
        const uint32_t lgc[8] = {la.x, la.y, la.z, la.w, lb.x, lb.y, lb.z, lb.w};
#pragma unroll
        for (int j2 = 0; j2 < 8; ++j2) {
          const float2 u = make_float2(__uint_as_float(v[2 * j2]), __uint_as_float(v[2 * j2 + 1]));
          const float2 a = __fmul2_rn(u, l2e2);
          const float2 t = make_float2(tanh_approx(u.x), tanh_approx(u.y));
          const float2 x = __ffma2_rn(a, t, a);  // w * log2(e), w = silu(2u)
          const float2 e = make_float2(ex2_approx(x.x), ex2_approx(x.y));
          den[j2 & 3] = __fadd2_rn(den[j2 & 3], e);
          num[j2 & 3] = __ffma2_rn(e, __half22float2(*reinterpret_cast<const __half2*>(&lgc[j2])), num[j2 & 3]);
        }
      };
      // the whole gate into registers first: LG is the slot's most contended resource (the next G1 waits for it)
      tmem_ld_x16(base + kLG, v0);
      tmem_ld_x16(base + kLG + 16, v1);
      if constexpr (L == 64) {
        tmem_ld_x16(base + kLG + 32, v2);
        tmem_ld_x16(base + kLG + 48, v3);
      }
      tmem_ld_wait_bind16(v0);
      tmem_ld_wait_bind16(v1);
      if constexpr (L == 64) {
        tmem_ld_wait_bind16(v2);
        tmem_ld_wait_bind16(v3);
      }
      // G3 of this query is complete: the diag buffer takes query c + 1
      if (have_n) {
        if (r < L) *sDw = gqv;
        fence_proxy_async_smem();
      }
      tc_fence_before();
      mbar_arrive(lg_free);  // phase c
      if (warp == 16) TR(0, 2, c);
      // The image buffer takes query c + 2 once G1 of query c + 1 has read it (long done inside a tile - that G1 is ahead
      // of G3(c) in pipe order; at a tile boundary it waits for the next tile's X and for the lg_free just arrived)
      if (have_nn) {
        WAIT(log_full, par ^ 1u, 21);
        store_image();
        fence_proxy_async_smem();
        mbar_arrive(img_ready);  // phase c + 2
      }
      // Prefetches from global memory go HERE, behind the arrives above: an mbarrier arrive (release) waits for the
      // thread's outstanding loads, so a load issued just in front of one turns its whole latency into a stall.  The next
      // arrive of this thread is a full query's math away.
      const float thr_q = P.thr ? __ldg(P.thr + (size_t)q * P.thr_stride) : -CUDART_INF_F;
      int tile_n3 = 0, q_n3 = 0;
      const bool have_n3 = have_nn && seq.next(tile_n3, q_n3);
      if (have_nn) load_gq(q_nn);     // 0.5 gq of query c + 2: its diag is staged inside E3 of query c + 1
      if (have_n3) load_image(q_n3);  // image of query c + 3: stored inside E3 of query c + 1
      gate(v0, 0);
      gate(v1, 1);
      if constexpr (L == 64) {
        gate(v2, 2);
        gate(v3, 3);
      }
      mbar_arrive(&bars->a2_read[wg][t3.b]);  // the logits of this query have been read: the buffer may take E1(c + 3)
      const float2 n2 = __fadd2_rn(__fadd2_rn(num[0], num[1]), __fadd2_rn(num[2], num[3]));
      const float2 d2 = __fadd2_rn(__fadd2_rn(den[0], den[1]), __fadd2_rn(den[2], den[3]));
      const float score = __fdividef(n2.x + n2.y, d2.x + d2.y);
      if (warp == 16) TR(0, 3, c);
      if (tile != map_tile) {
        map_tile = tile;
        map_phys = P.phys_tile(t0 + tile);
      }
      const int64_t item = (int64_t)map_phys * kTile + r;
      if (item < P.N) {
        if (P.scores) P.scores[(size_t)q * P.ld + ((int64_t)tile * kTile + r)] = score;
        if (P.thr && !(score < thr_q)) {  // NaN passes the filter on purpose
          const int pos = atomicAdd(P.cand_cnt + q, 1);
          if (pos < P.cand_cap) {
            P.cand_scores[(size_t)q * P.cand_cap + pos] = score;
            P.cand_idx[(size_t)q * P.cand_cap + pos] = (int32_t)item;
          }
        }
      }
      ++c;
      t3.inc();
      tile = tile_n;
      q = q_n;
      have = have_n;
      tile_n = tile_nn;
      q_n = q_nn;
      have_n = have_nn;
      tile_nn = tile_n3;
      q_nn = q_n3;
      have_nn = have_n3;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kCtlWarp) tmem_dealloc<512>(tmem);
}

}  // namespace v3
