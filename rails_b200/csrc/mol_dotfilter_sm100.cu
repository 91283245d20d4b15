// Streaming dot-product top-k on the sm_100a tensor cores (interface and the proof of exactness: mol_dotfilter.cuh).
//
// dot_filter_kernel - one persistent CTA per SM, 320 threads:
//   warp 0      TMA producer: the <= 256 query rows of this launch once (K / 32 boxes of cc rows x 32 fp32, SWIZZLE_128B),
//               then the item tiles (128 rows) box by box through a ring of 16 KB stages;
//   warp 1      MMA issuer: per box 4 x tcgen05.mma kind::tf32 (M = 128 items, N = cc query rows, K = 8) into one of two
//               256-column TMEM accumulators; tcgen05.commit frees the stage, and after the last box of a tile announces
//               the accumulator;
//   warps 2..9  epilogue (two warps per TMEM lane quarter, each half of the columns): TMEM lane = item row, 32 columns
//               per tcgen05.ld; sample mode stores the values (column-major:
//               coalesced over the 32 lanes), filter mode compares against the per-column levels in shared memory and
//               appends the rare survivors (atomic counter per query row).
// The fp32 operands are consumed as they are (tf32 reads the upper 19 bits); HBM traffic = the item matrix once per
// <= 256 query rows.  Roofline at K = 64, 256 query rows: 8 MMAs of 128 clk per tile -> 55 us of tensor pipe per 1M items,
// 256 MB of HBM = 39 us.
#include <cuda.h>
#include <math_constants.h>

#include "mol_dotfilter.cuh"
#include "sm100_ptx.cuh"

namespace mol {
using namespace sm100;

namespace {

constexpr int DF_TILE = 128;
constexpr int DF_EPI_WARPS = 8;  // two per SM sub-partition: one's tcgen05.ld / branch latencies hide behind the other
constexpr int DF_EPI_THREADS = DF_EPI_WARPS * 32;
constexpr int DF_THREADS = 64 + DF_EPI_THREADS;
constexpr int DF_MAX_STAGES = 6;
constexpr int DF_BOX_BYTES = DF_TILE * 128;  // 128 rows x 32 fp32
constexpr int DF_MAX_CC = 256;
constexpr int DF_SMEM_LIMIT = 227 * 1024;
constexpr int DF_STAGE_CAP = 1024;  // survivors staged in shared memory between two drains
constexpr int DF_HIT_PITCH = 36;    // floats per lane of the hit-extraction buffer (16-byte aligned rows, conflict-free float4 stores)
constexpr int DF_HIT_BYTES = 32 * DF_HIT_PITCH * 4;  // per epilogue warp

struct DfParams {
  // sample mode (out != nullptr): out[col * ld + tile * 128 + row in tile] = value
  float* out;
  int64_t ld;
  // filter mode
  const float* level;  // (rc) per query row: keep value >= level
  int32_t* cnt;        // (rc)
  float* cand_val;     // (rc, cap)
  int32_t* cand_idx;   // (rc, cap)
  int cap;
  int rc;           // query rows of this launch (<= cc)
  int cc;           // MMA N: rc rounded up to 64
  int ks;           // K / 32
  int col0;         // first fp32 column of the item rows that takes part
  int64_t N;        // item rows
  int tiles;        // logical tiles of this launch
  int tile_stride;  // physical tile = logical tile * tile_stride
  int stages;
  int drain_every;  // filter mode: the staged survivors go to the per-row buffers every drain_every tiles
  uint32_t idesc;
};

struct DfBars {
  uint64_t full[DF_MAX_STAGES], empty[DF_MAX_STAGES], q_full, acc_full[2], acc_empty[2];
  uint32_t tmem_base;
  int staged;  // entries in the survivor staging buffer (may run past DF_STAGE_CAP: the excess went straight to global)
};

__device__ __forceinline__ void epilogue_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(DF_EPI_THREADS) : "memory"); }
// the per-column levels never change during a launch: a plain (non-volatile) shared-memory load the compiler may reorder
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 r;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
  return r;
}

__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4)                      // D format = F32
         | (2u << 7)                    // A format = TF32
         | (2u << 10)                   // B format = TF32
         | ((uint32_t)(N >> 3) << 17)   // N >> 3
         | ((uint32_t)(M >> 4) << 24);  // M >> 4   (A and B K-major, dense)
}

__device__ __forceinline__ void umma_ss_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// (plain try_wait loop: with the suspend-time hint of mbar_wait_sleep a tile took ~10 us - the waits of this short
// pipeline are rarely satisfied on entry and the suspended threads woke up late)
__device__ __forceinline__ void df_wait(uint64_t* bar, uint32_t parity) {
#ifdef MOL_DF_SLEEP_WAIT
  mbar_wait_sleep(bar, parity);
#else
  mbar_wait(bar, parity);
#endif
}

__global__ void __launch_bounds__(DF_THREADS, 1)
dot_filter_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmQ, const DfParams P) {
  extern __shared__ unsigned char df_smem_raw[];
  unsigned char* smem =
      reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(df_smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t qbox = (uint32_t)P.cc * 128u;  // bytes of one 32-column box of the query rows (cc % 64 == 0: 1024-aligned)
  unsigned char* sQ = smem;
  unsigned char* sA = sQ + (size_t)P.ks * qbox;
  float* sLevel = reinterpret_cast<float*>(sA + (size_t)P.stages * DF_BOX_BYTES);
  // survivor staging: an append costs one shared-memory atomic; the global per-row counters are only touched by the
  // drain, 128 entries at a time (a returning global atomic per survivor inside the epilogue loop serialised a ~1 us
  // round trip per hit: 9 us per tile instead of 0.5)
  float* sStageVal = sLevel + DF_MAX_CC;
  int32_t* sStageRow = reinterpret_cast<int32_t*>(sStageVal + DF_STAGE_CAP);
  int32_t* sStageCol = sStageRow + DF_STAGE_CAP;
  float* sHit = reinterpret_cast<float*>(sStageCol + DF_STAGE_CAP);  // DF_EPI_WARPS x 32 lanes x DF_HIT_PITCH
  DfBars* bars = reinterpret_cast<DfBars*>(reinterpret_cast<unsigned char*>(sHit) + DF_EPI_WARPS * DF_HIT_BYTES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < DF_MAX_STAGES; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    mbar_init(&bars->q_full, 1);
    bars->staged = 0;
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars->acc_full[a], 1);
      mbar_init(&bars->acc_empty[a], DF_EPI_THREADS);
    }
    fence_mbar_init();
  }
  for (int i = tid; i < DF_MAX_CC; i += DF_THREADS)
    sLevel[i] = (P.level != nullptr && i < P.rc) ? P.level[i] : CUDART_INF_F;
  if (warp == 1) tmem_alloc<512>(&bars->tmem_base);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmQ);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      mbar_arrive_expect_tx(&bars->q_full, (uint32_t)P.ks * qbox);
      for (int ks = 0; ks < P.ks; ++ks) tma_load_2d(sQ + (size_t)ks * qbox, &tmQ, &bars->q_full, ks * 32, 0);
      int ib = 0;
      for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
        const int row0 = tile * P.tile_stride * DF_TILE;
        for (int ks = 0; ks < P.ks; ++ks, ++ib) {
          const int s = ib % P.stages;
          const uint32_t ph = (uint32_t)(ib / P.stages) & 1u;
          df_wait(&bars->empty[s], ph ^ 1u);
          mbar_arrive_expect_tx(&bars->full[s], DF_BOX_BYTES);
          tma_load_2d(sA + (size_t)s * DF_BOX_BYTES, &tmA, &bars->full[s], P.col0 + ks * 32, row0);
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    // (the whole warp runs the loop converged; one elected lane issues - see mol_coarse_sm100.cu)
    const uint32_t sAa = smem_u32(sA), sQa = smem_u32(sQ);
    df_wait(&bars->q_full, 0);
    tc_fence_after();
    int ib = 0, it = 0;
    for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      df_wait(&bars->acc_empty[acc], (use & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem + (uint32_t)acc * 256u;
      for (int ks = 0; ks < P.ks; ++ks, ++ib) {
        const int s = ib % P.stages;
        df_wait(&bars->full[s], (uint32_t)(ib / P.stages) & 1u);
        tc_fence_after();
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {  // 4 K steps of 8 tf32 = 32 bytes inside the 128-byte swizzle atom
            const uint64_t da = make_smem_desc(sAa + (uint32_t)s * DF_BOX_BYTES + kk * 32, 16, 1024, 2);
            const uint64_t db = make_smem_desc(sQa + (uint32_t)ks * qbox + kk * 32, 16, 1024, 2);
            umma_ss_tf32(d_tmem, da, db, P.idesc, (ks | kk) != 0);
          }
          umma_commit(&bars->empty[s]);
          if (ks == P.ks - 1) umma_commit(&bars->acc_full[acc]);
        }
        __syncwarp();
      }
    }
  } else {
    // =============================== epilogue (warps 2..9) ===============================
    const int quarter = warp & 3;  // TMEM lanes this warp may read
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const int nch = P.cc / 32;  // even (cc % 64 == 0)
    const int half = (warp - 2) >> 2;  // which half of the column chunks this warp takes
    const int c_lo = half * (nch / 2), c_hi = c_lo + nch / 2;
    const uint32_t sLevel_a = smem_u32(sLevel);
    const int etid = tid - 64;
    auto drain = [&]() {
      epilogue_bar_sync();
      int n = bars->staged;
      if (n > DF_STAGE_CAP) n = DF_STAGE_CAP;
      for (int e = etid; e < n; e += DF_EPI_THREADS) {
        const int col = sStageCol[e];
        const int slot = atomicAdd(P.cnt + col, 1);
        if (slot < P.cap) {
          P.cand_val[(int64_t)col * P.cap + slot] = sStageVal[e];
          P.cand_idx[(int64_t)col * P.cap + slot] = sStageRow[e];
        }
      }
      epilogue_bar_sync();
      if (etid == 0) bars->staged = 0;
      epilogue_bar_sync();
    };
    int it = 0;
    for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      if (P.out == nullptr && it > 0 && it % P.drain_every == 0) drain();
      df_wait(&bars->acc_full[acc], use & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem + (uint32_t)acc * 256u + lane_base;
      const int in_tile = quarter * 32 + lane;
      const int64_t row = (int64_t)tile * P.tile_stride * DF_TILE + in_tile;
      const int64_t out_col = (int64_t)tile * DF_TILE + in_tile;
      auto consume = [&](const uint32_t (&v)[32], int cbase) __attribute__((always_inline)) {
        if (P.out != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (cbase + j < P.rc) P.out[(int64_t)(cbase + j) * P.ld + out_col] = __uint_as_float(v[j]);
        } else {
          // (each epilogue warp is alone on its SM sub-partition: every dependent instruction and every branch costs its
          // full latency, so the test runs as four independent predicate chains and the rare hit is extracted without a
          // branch per column)
          const uint32_t lv = sLevel_a + (uint32_t)cbase * 4u;
          bool a0 = false, a1 = false, a2 = false, a3 = false;
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 t = lds_f4(lv + 16u * j4);
            a0 |= __uint_as_float(v[4 * j4 + 0]) >= t.x;
            a1 |= __uint_as_float(v[4 * j4 + 1]) >= t.y;
            a2 |= __uint_as_float(v[4 * j4 + 2]) >= t.z;
            a3 |= __uint_as_float(v[4 * j4 + 3]) >= t.w;
          }
          if ((a0 | a1 | a2 | a3) && row < P.N) {
            auto stage = [&](int col, float x) __attribute__((always_inline)) {
              const int e = atomicAdd(&bars->staged, 1);
              if (e < DF_STAGE_CAP) {
                sStageVal[e] = x;
                sStageRow[e] = (int32_t)row;
                sStageCol[e] = col;
              } else {  // staging full (a pathological row): straight to the row's buffer
                const int slot = atomicAdd(P.cnt + col, 1);
                if (slot < P.cap) {
                  P.cand_val[(int64_t)col * P.cap + slot] = x;
                  P.cand_idx[(int64_t)col * P.cap + slot] = (int32_t)row;
                }
              }
            };
            // hit extraction: a bit mask of the passing columns, the chunk's values parked in this lane's own row of a
            // shared-memory buffer (no other lane touches it: no synchronisation), then one staged entry per set bit.
            // (Selecting a register by a run-time column index is not possible; a select chain per hit cost ~260
            // instructions where this costs ~80.)
            float* mine = sHit + ((warp - 2) * 32 + lane) * DF_HIT_PITCH;
            uint32_t mask = 0;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 t = lds_f4(lv + 16u * j4);
              const float x0 = __uint_as_float(v[4 * j4]), x1 = __uint_as_float(v[4 * j4 + 1]);
              const float x2 = __uint_as_float(v[4 * j4 + 2]), x3 = __uint_as_float(v[4 * j4 + 3]);
              mask |= (x0 >= t.x) ? (1u << (4 * j4)) : 0u;
              mask |= (x1 >= t.y) ? (2u << (4 * j4)) : 0u;
              mask |= (x2 >= t.z) ? (4u << (4 * j4)) : 0u;
              mask |= (x3 >= t.w) ? (8u << (4 * j4)) : 0u;
              *reinterpret_cast<float4*>(mine + 4 * j4) = make_float4(x0, x1, x2, x3);
            }
            while (mask != 0) {
              const int j = __ffs((int)mask) - 1;
              mask &= mask - 1;
              stage(cbase + j, mine[j]);
            }
          }
        }
      };
      uint32_t va[32], vb[32];
      tmem_ld_x32(taddr + (uint32_t)c_lo * 32u, va);
      for (int c = c_lo; c < c_hi; c += 2) {  // (every condition is warp-uniform)
        tmem_ld_wait_bind32(va);
        if (c + 1 < c_hi) tmem_ld_x32(taddr + (uint32_t)(c + 1) * 32u, vb);
        consume(va, c * 32);
        if (c + 1 < c_hi) {
          tmem_ld_wait_bind32(vb);
          if (c + 2 < c_hi) tmem_ld_x32(taddr + (uint32_t)(c + 2) * 32u, va);
          consume(vb, (c + 1) * 32);
        }
      }
      tc_fence_before();
      mbar_arrive(&bars->acc_empty[acc]);
    }
    if (P.out == nullptr) drain();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// ---- helpers around the tensor pass -----------------------------------------------------------------------------

// *out = max over rows of |x[col0 : col0 + K]|_2 (non-negative floats order like their bit patterns).  K / 4 <= 32: a group
// of K / 4 lanes owns a row (one float4 each, 512 contiguous bytes per warp when the rows are dense); else a warp per row.
__global__ void __launch_bounds__(256)
row_norm_max_kernel(const float* __restrict__ items, int64_t N, int64_t pitch, int col0, int K, float* __restrict__ out,
                    const float* __restrict__ skip_if_valid) {
  if (skip_if_valid != nullptr && skip_if_valid[0] >= 0.f) return;  // a cached bound exists
  const int lane = threadIdx.x & 31;
  const int k4 = K / 4;
  const int lpr = k4 < 32 ? k4 : 32;  // lanes per row (8, 16, 24, 32; K % 32 == 0)
  const int rpw = 32 / lpr;           // rows per warp and iteration
  const int sub = lane / lpr, li = lane - sub * lpr;
  const int64_t warp_id = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  float best = 0.f;
  for (int64_t r0 = warp_id * rpw; r0 < N; r0 += warps * rpw) {
    const int64_t r = r0 + sub;
    float ss = 0.f;
    if (sub < rpw && r < N) {
      const float4* x = reinterpret_cast<const float4*>(items + r * pitch + col0);
      for (int i = li; i < k4; i += lpr) {
        const float4 v = __ldg(x + i);
        ss = fmaf(v.x, v.x, ss);
        ss = fmaf(v.y, v.y, ss);
        ss = fmaf(v.z, v.z, ss);
        ss = fmaf(v.w, v.w, ss);
      }
    }
    // sum over the lanes of a row group (lpr = 8, 16, 32: a butterfly; 24: rpw = 1, the upper 8 lanes hold zeros)
    const int width = lpr == 24 ? 32 : lpr;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      if (o < width) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    best = fmaxf(best, ss);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
  if (lane == 0 && best > 0.f) atomicMax(reinterpret_cast<int*>(out), __float_as_int(sqrtf(best) * 1.000001f));
}

// Filter level of a row from its S sampled values, one CTA per row: every thread keeps the maximum of its strided slice
// (S / 256 values), the 256 maxima are sorted and the m-th largest is the level.  When two of the row's m largest values
// share a slice the result is the (m+1)-th largest instead - the level is a heuristic (it only steers how many items
// survive; completeness is proven afterwards, mol_dotfilter.cuh), and this is 67 MB read once instead of a (rows, S) radix
// select (82 us for 512 rows).
constexpr int SL_THREADS = 256;
__global__ void __launch_bounds__(SL_THREADS)
sample_level_kernel(const float* __restrict__ samp, int64_t S, int m, float* __restrict__ level_out) {
  __shared__ float sv[SL_THREADS];
  const int r = blockIdx.x, t = threadIdx.x;
  const float4* row = reinterpret_cast<const float4*>(samp + (int64_t)r * S);  // S % 128 == 0, rows 16-byte aligned
  float best = -CUDART_INF_F;
  for (int64_t i = t; i < S / 4; i += SL_THREADS) {
    const float4 v = __ldg(row + i);
    best = fmaxf(best, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
  }
  sv[t] = best;
  __syncthreads();
  for (int size = 2; size <= SL_THREADS; size <<= 1) {  // bitonic sort, descending
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const int p = t ^ stride;
      if (p > t) {
        const bool desc = (t & size) == 0;
        const float a = sv[t], b = sv[p];
        if ((a < b) == desc) {
          sv[t] = b;
          sv[p] = a;
        }
      }
      __syncthreads();
    }
  }
  if (t == 0) level_out[r] = sv[m - 1 < SL_THREADS ? m - 1 : SL_THREADS - 1];
}

__global__ void norm_publish_kernel(float* cache) {
  if (cache[0] < 0.f) cache[0] = cache[1];
}

// level[r] = m-th largest sampled value of row r; check[r] = level[r] + E_r with E_r = 2^-8 |q_r| xmax: every item whose
// fp32 dot product reaches check[r] has a tf32 value >= level[r] and is therefore among the survivors.
__global__ void level_kernel(const float* __restrict__ samp_top, int m, const float* __restrict__ Q, int64_t q_pitch, int K,
                             const float* __restrict__ xmax, float xmax_host, float* __restrict__ level,
                             float* __restrict__ check, int rc) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rc) return;
  const float* q = Q + (int64_t)r * q_pitch;
  float ss = 0.f;
  for (int k = 0; k < K; ++k) ss = fmaf(q[k], q[k], ss);
  const float e = (1.0f / 256.0f) * sqrtf(ss) * (xmax ? xmax[0] : xmax_host) * 1.0001f;
  const float t = samp_top[(int64_t)r * m + (m - 1)];  // (m = 1: one level per row, from sample_level_kernel)
  level[r] = t;
  check[r] = t + e;
}

// fp32 values of the survivors, with the k-ascending fmaf chain of linear_kernel (mol_prologue.cu): bit-identical to
// the materialised-matrix path
__global__ void __launch_bounds__(256)
gather_dot_kernel(const float* __restrict__ items, int64_t pitch, int col0, int K, const float* __restrict__ Q,
                  int64_t q_pitch, const int32_t* __restrict__ cnt, const int32_t* __restrict__ cidx,
                  float* __restrict__ cexact, int cap) {
  extern __shared__ float gd_q[];
  const int r = blockIdx.y;
  int n = cnt[r];
  if (n > cap) n = cap;
  if ((int)(blockIdx.x * blockDim.x) >= n) return;
  for (int k = threadIdx.x; k < K; k += blockDim.x) gd_q[k] = Q[(int64_t)r * q_pitch + k];
  __syncthreads();
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n) return;
  const int32_t idx = cidx[(int64_t)r * cap + slot];
  const float4* x = reinterpret_cast<const float4*>(items + (int64_t)idx * pitch + col0);
  float acc = 0.f;
  for (int k4 = 0; k4 < K / 4; ++k4) {
    const float4 v = __ldg(x + k4);
    acc = fmaf(gd_q[4 * k4 + 0], v.x, acc);
    acc = fmaf(gd_q[4 * k4 + 1], v.y, acc);
    acc = fmaf(gd_q[4 * k4 + 2], v.z, acc);
    acc = fmaf(gd_q[4 * k4 + 3], v.w, acc);
  }
  cexact[(int64_t)r * cap + slot] = acc;
}

__global__ void verify_kernel(const int32_t* __restrict__ cnt, const float* __restrict__ topk, int kk,
                              const float* __restrict__ check, int cap, int rows_fb, int32_t* __restrict__ flags,
                              int32_t* __restrict__ any_flag, int32_t* __restrict__ stats, int rc) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rc) return;
  const int c = cnt[r];
  const float sk = topk[(int64_t)r * kk + (kk - 1)];  // -inf when fewer than kk survivors
  const bool bad = (c > cap) || !(sk >= check[r]);
  flags[r] = bad ? 1 : 0;
  if (bad) atomicOr(any_flag + r / rows_fb, 1);
  if (stats) {
    if (bad) atomicAdd(stats + 0, 1);
    if (c > cap) atomicAdd(stats + 1, 1);
    atomicMax(stats + 2, c);
    stats[3] = 1;
  }
}

// fallback: plain fp32 dot products of the flagged rows against every item (same fmaf chain)
__global__ void __launch_bounds__(256)
dot_rows_flagged_kernel(const float* __restrict__ items, int64_t N, int64_t pitch, int col0, int K,
                        const float* __restrict__ Q, int64_t q_pitch, int nb, const int32_t* __restrict__ flags,
                        const int32_t* __restrict__ any_flag, float* __restrict__ out) {
  extern __shared__ float fr_q[];
  if (any_flag[0] == 0) return;
  for (int r = 0; r < nb; ++r) {
    if (flags[r] == 0) continue;
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) fr_q[k] = Q[(int64_t)r * q_pitch + k];
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
      const float4* x = reinterpret_cast<const float4*>(items + i * pitch + col0);
      float acc = 0.f;
      for (int k4 = 0; k4 < K / 4; ++k4) {
        const float4 v = __ldg(x + k4);
        acc = fmaf(fr_q[4 * k4 + 0], v.x, acc);
        acc = fmaf(fr_q[4 * k4 + 1], v.y, acc);
        acc = fmaf(fr_q[4 * k4 + 2], v.z, acc);
        acc = fmaf(fr_q[4 * k4 + 3], v.w, acc);
      }
      out[(int64_t)r * N + i] = acc;
    }
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// fp32 matrix (rows, cols) with row pitch `pitch` floats; boxes of 32 columns (one 128-byte swizzle atom) x box_rows
int encode_f32(CUtensorMap* m, const float* base, uint64_t cols, uint64_t rows, uint64_t pitch, uint32_t box_rows) {
  static PFN_encodeTiled encode = nullptr;
  if (!encode) {
    cudaDriverEntryPointQueryResult qres;
    MOL_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
    MOL_CHECK_ARG(encode != nullptr, "cuTensorMapEncodeTiled not available");
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {pitch * sizeof(float)};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (fp32) failed (%d)", (int)r);
    return MOL_ERR_CUDA;
  }
  return MOL_OK;
}

int cc_max_of(int K) {
  int cc = (128 * 1024) / (K * 4);  // the query rows of a launch take <= 128 KB of shared memory
  if (cc > DF_MAX_CC) cc = DF_MAX_CC;
  return cc / 64 * 64;
}

// One tensor pass of query rows [0, rc) (rc <= cc_max) over `tiles` logical tiles.
int launch_dot_filter(const float* items, int64_t N, int64_t pitch, int col0, int K, const float* Q, int64_t q_pitch, int rc,
                      int tiles, int tile_stride, float* out, int64_t ld, const float* level, int32_t* cnt,
                      float* cand_val, int32_t* cand_idx, int cap, int expect_per_row, cudaStream_t st) {
  if (rc == 0 || tiles == 0) return MOL_OK;
  DfParams P;
  P.out = out;
  P.ld = ld;
  P.level = level;
  P.cnt = cnt;
  P.cand_val = cand_val;
  P.cand_idx = cand_idx;
  P.cap = cap;
  P.rc = rc;
  P.cc = (rc + 63) / 64 * 64;
  P.ks = K / 32;
  P.col0 = col0;
  P.N = N;
  P.tiles = tiles;
  P.tile_stride = tile_stride;
  P.idesc = make_idesc_tf32(DF_TILE, P.cc);
  {  // drain when about half the staging buffer is expected to be in use
    const double per_tile = (double)rc * (double)expect_per_row * DF_TILE / (double)(N > 0 ? N : 1);
    int every = per_tile > 1.0 ? (int)((DF_STAGE_CAP / 2) / per_tile) : DF_STAGE_CAP / 2;
    P.drain_every = every < 1 ? 1 : (every > 64 ? 64 : every);
  }
  const size_t fixed =
      1024 + (size_t)P.ks * P.cc * 128 + DF_MAX_CC * sizeof(float) + (size_t)DF_STAGE_CAP * 12 +
      (size_t)DF_EPI_WARPS * DF_HIT_BYTES + sizeof(DfBars) + 64;
  int stages = (int)((DF_SMEM_LIMIT - fixed) / DF_BOX_BYTES);
  if (stages > DF_MAX_STAGES) stages = DF_MAX_STAGES;
  MOL_CHECK_ARG(stages >= 2, "dot filter: K=%d with %d query rows does not fit shared memory", K, P.cc);
  P.stages = stages;
  const size_t smem = fixed + (size_t)stages * DF_BOX_BYTES;
  CUtensorMap tmA, tmQ;
  MOL_TRY(encode_f32(&tmA, items, (uint64_t)pitch, (uint64_t)N, (uint64_t)pitch, DF_TILE));
  MOL_TRY(encode_f32(&tmQ, Q, (uint64_t)K, (uint64_t)rc, (uint64_t)q_pitch, (uint32_t)P.cc));
  static int sms = 0;
  if (sms == 0) {
    MOL_CUDA(cudaFuncSetAttribute(dot_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DF_SMEM_LIMIT));
    int dev = 0, n = 148;
    MOL_CUDA(cudaGetDevice(&dev));
    MOL_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    sms = n;
  }
  const int grid = tiles < sms ? tiles : sms;
  dot_filter_kernel<<<grid, DF_THREADS, smem, st>>>(tmA, tmQ, P);
  MOL_LAUNCH_CHECK();
  return MOL_OK;
}

int64_t env_i64(const char* name, int64_t dflt) {
  const char* e = getenv(name);
  if (!e) return dflt;
  return (int64_t)atoll(e);
}

struct Sizes {
  int T, cap, m, nst, stride;
  int64_t S;
};
Sizes sizes_of(int64_t N, int kk) {
  Sizes z;
  z.T = 4 * kk > 256 ? 4 * kk : 256;                  // survivors aimed at
  z.cap = z.T <= 4096 ? 4 * z.T : 2 * z.T;            // survivor capacity (the count's relative spread is ~ 1 / sqrt(m))
  const int64_t tiles = (N + DF_TILE - 1) / DF_TILE;
  int64_t s_target = N / 8 < 32768 ? N / 8 : 32768;
  int64_t nst = s_target / DF_TILE;
  if (nst < 1) nst = 1;
  z.nst = (int)nst;
  z.stride = (int)(tiles / nst);
  if (z.stride < 1) z.stride = 1;
  z.S = nst * DF_TILE;
  int64_t m = ((int64_t)z.T * z.S + N - 1) / N;
  if (m < 8) m = 8;  // (the survivor count spreads like a Gamma(m) variable: m >= 8 keeps "fewer than kk survive" out of reach)
  if (m > z.S) m = z.S;
  z.m = (int)m;
  return z;
}

}  // namespace

bool dot_topk_eligible(int64_t N, int R, int K, int kk) {
  if (env_i64("MOL_B200_DOTFILTER", 1) == 0) return false;
  if (R < 1 || kk < 1 || K < 32 || K > 256 || K % 32 != 0) return false;
  if (N < env_i64("MOL_B200_DOTFILTER_MIN_ITEMS", 65536) || N >= (1ll << 31) - 256) return false;
  const Sizes z = sizes_of(N, kk);
  if ((int64_t)z.cap * 8 > N || z.m > MOL_MAX_K || kk > MOL_MAX_K) return false;
  return true;
}

int dot_topk_norm_cache(const float* items, int64_t N, int64_t pitch, int col0, int K, float* cache, cudaStream_t st) {
  if (N == 0) return MOL_OK;
  int64_t blocks = (N + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  row_norm_max_kernel<<<(unsigned)blocks, 256, 0, st>>>(items, N, pitch, col0, K, cache + 1, cache);
  MOL_LAUNCH_CHECK();
  norm_publish_kernel<<<1, 1, 0, st>>>(cache);
  MOL_LAUNCH_CHECK();
  return MOL_OK;
}

bool dot_topk_aligned(const float* items, int64_t pitch, int col0, const float* Q, int64_t q_pitch) {
  return (reinterpret_cast<uintptr_t>(items) % 16 == 0) && (reinterpret_cast<uintptr_t>(Q) % 16 == 0) && pitch % 4 == 0 &&
         q_pitch % 4 == 0 && col0 % 4 == 0;
}

void dot_topk_plan(Arena& a, int64_t N, int R, int K, int kk, float* fb_scores, int fb_rows, DotTopkPlan* p) {
  (void)K;
  const Sizes z = sizes_of(N, kk);
  p->N = N;
  p->R = R;
  p->kk = kk;
  p->Rc = R < 8192 ? R : 8192;
  p->S = z.S;
  p->tile_stride = z.stride;
  p->m = z.m;
  p->cap = z.cap;
  p->expect = z.T;
  p->rows_fb = fb_rows < 1 ? 1 : (fb_rows > p->Rc ? p->Rc : fb_rows);
  p->fb_scores = fb_scores;
  const size_t rc = (size_t)p->Rc;
  // the (Rc, S) sample matrix lives in the fallback matrix when that is large enough (it is consumed before any fallback)
  p->samp = ((int64_t)fb_rows * N >= (int64_t)rc * z.S) ? fb_scores : a.take<float>(rc * (size_t)z.S);
  p->samp_aliases_fb = ((int64_t)fb_rows * N >= (int64_t)rc * z.S) ? 1 : 0;
  const size_t seg = rc * 8 * (size_t)z.m;  // the sample select runs on 8 segments per row
  p->seg_scores = a.take<float>(seg);
  p->seg_idx = a.take<int32_t>(seg);
  p->samp_top = a.take<float>(rc * (size_t)z.m);
  p->level = a.take<float>(rc);
  p->check = a.take<float>(rc);
  p->cnt = a.take<int32_t>(rc);
  p->cval = a.take<float>(rc * (size_t)z.cap);
  p->cidx = a.take<int32_t>(rc * (size_t)z.cap);
  p->cexact = a.take<float>(rc * (size_t)z.cap);
  p->flags = a.take<int32_t>(rc);
  p->any_flag = a.take<int32_t>(rc / (size_t)p->rows_fb + 2);
  const size_t fseg = ((size_t)p->rows_fb + 2 * 148 + 1) * (size_t)kk;  // rows' * select_num_segments(N, rows', kk) slots
  p->fb_seg_scores = a.take<float>(fseg);
  p->fb_seg_idx = a.take<int32_t>(fseg);
  p->xmax = a.take<float>(1);
}

int dot_topk_run(const DotTopkPlan& p, const float* items, int64_t pitch, int col0, int K, const float* xmax_dev,
                 float xmax_host, const float* Q, int64_t q_pitch, float* out_scores, int32_t* out_idx, int64_t* out_ids,
                 const int64_t* id_map, int32_t* stats, cudaStream_t st) {
  MOL_CHECK_ARG(dot_topk_aligned(items, pitch, col0, Q, q_pitch), "dot filter: operands must be 16-byte aligned");
  MOL_CHECK_ARG(col0 + K <= pitch && K <= q_pitch, "dot filter: column range outside the row");
  const int64_t N = p.N;
  const int kk = p.kk, cap = p.cap, m = p.m;
  const int tiles = (int)((N + DF_TILE - 1) / DF_TILE);
  const int nst = (int)(p.S / DF_TILE);
  const int ccm = cc_max_of(K);
  if (!xmax_dev && !(xmax_host > 0.f)) {
    MOL_CUDA(cudaMemsetAsync(p.xmax, 0, sizeof(float), st));
    int64_t blocks = (N + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    row_norm_max_kernel<<<(unsigned)blocks, 256, 0, st>>>(items, N, pitch, col0, K, p.xmax, nullptr);
    MOL_LAUNCH_CHECK();
    xmax_dev = p.xmax;
  }
  for (int r0 = 0; r0 < p.R; r0 += p.Rc) {
    const int rc = (p.R - r0 < p.Rc) ? (p.R - r0) : p.Rc;
    const float* Qc = Q + (int64_t)r0 * q_pitch;
    float* o_scores = out_scores + (int64_t)r0 * kk;
    int32_t* o_idx = out_idx ? out_idx + (int64_t)r0 * kk : nullptr;
    int64_t* o_ids = out_ids ? out_ids + (int64_t)r0 * kk : nullptr;
    // (1) sample pass -> (rc, S) matrix -> the m best per row
    for (int c0 = 0; c0 < rc; c0 += ccm) {
      const int n = (rc - c0 < ccm) ? (rc - c0) : ccm;
      MOL_TRY(launch_dot_filter(items, N, pitch, col0, K, Qc + (int64_t)c0 * q_pitch, q_pitch, n, nst, p.tile_stride,
                                p.samp + (int64_t)c0 * p.S, p.S, nullptr, nullptr, nullptr, nullptr, 0, 0, st));
    }
    const bool bucket_level = m <= 64 && p.S >= 4 * SL_THREADS * 8;  // small m: maxima of 256 slices per row are enough
    if (bucket_level) {
      sample_level_kernel<<<rc, SL_THREADS, 0, st>>>(p.samp, p.S, m, p.samp_top);
      MOL_LAUNCH_CHECK();
    } else {
      // (segments of ~4k values: the one-CTA-per-row select of 32k values took 68 us for 512 rows)
      int S1 = (int)(p.S / 4096);
      S1 = S1 < 1 ? 1 : (S1 > 8 ? 8 : S1);
      if ((int64_t)S1 * m > p.S) S1 = 1;
      const float* sel = p.samp;
      const int32_t* pay = nullptr;
      int64_t sn = p.S;
      if (S1 > 1) {
        MOL_TRY(launch_select_segments(p.samp, p.S, p.S, rc, S1, m, p.seg_scores, p.seg_idx, nullptr, st));
        sel = p.seg_scores;
        pay = p.seg_idx;
        sn = (int64_t)S1 * m;
      }
      MOL_TRY(launch_select_final_i32(sel, pay, sn, sn, rc, m, p.samp_top, nullptr, nullptr, nullptr, nullptr, st));
    }
    level_kernel<<<(rc + 127) / 128, 128, 0, st>>>(p.samp_top, bucket_level ? 1 : m, Qc, q_pitch, K, xmax_dev, xmax_host, p.level,
                                                   p.check, rc);
    MOL_LAUNCH_CHECK();
    MOL_CUDA(cudaMemsetAsync(p.cnt, 0, (size_t)rc * sizeof(int32_t), st));
    MOL_CUDA(cudaMemsetAsync(p.cidx, 0xFF, (size_t)rc * cap * sizeof(int32_t), st));
    MOL_CUDA(cudaMemsetAsync(p.any_flag, 0, ((size_t)rc / p.rows_fb + 2) * sizeof(int32_t), st));
    // (2) the filter pass over every item tile
    for (int c0 = 0; c0 < rc; c0 += ccm) {
      const int n = (rc - c0 < ccm) ? (rc - c0) : ccm;
      MOL_TRY(launch_dot_filter(items, N, pitch, col0, K, Qc + (int64_t)c0 * q_pitch, q_pitch, n, tiles, 1, nullptr, 0,
                                p.level + c0, p.cnt + c0, p.cval + (int64_t)c0 * cap, p.cidx + (int64_t)c0 * cap, cap, p.expect, st));
    }
    // (3) fp32 values of the survivors, the kk best
    {
      dim3 grid((unsigned)((cap + 255) / 256), (unsigned)rc);
      gather_dot_kernel<<<grid, 256, (size_t)K * sizeof(float), st>>>(items, pitch, col0, K, Qc, q_pitch, p.cnt, p.cidx,
                                                                      p.cexact, cap);
      MOL_LAUNCH_CHECK();
    }
    MOL_TRY(launch_select_final_i32(p.cexact, p.cidx, cap, cap, rc, kk, o_scores, o_idx, o_ids, id_map, nullptr, st));
    // (4) completeness test, fallback for the rows that fail it
    verify_kernel<<<(rc + 127) / 128, 128, 0, st>>>(p.cnt, o_scores, kk, p.check, cap, p.rows_fb, p.flags, p.any_flag,
                                                    stats, rc);
    MOL_LAUNCH_CHECK();
    for (int b1 = 0; b1 < rc; b1 += p.rows_fb) {
      const int nb = (rc - b1 < p.rows_fb) ? (rc - b1) : p.rows_fb;
      const int32_t* fl = p.flags + b1;
      int64_t blocks = (N + 255) / 256;
      if (blocks > 148 * 8) blocks = 148 * 8;
      dot_rows_flagged_kernel<<<(unsigned)blocks, 256, (size_t)K * sizeof(float), st>>>(
          items, N, pitch, col0, K, Qc + (int64_t)b1 * q_pitch, q_pitch, nb, fl, p.any_flag + b1 / p.rows_fb, p.fb_scores);
      MOL_LAUNCH_CHECK();
      const int S2 = select_num_segments(N, nb, kk);  // (few CTAs: this launch is idle unless a row failed its test)
      const float* sel = p.fb_scores;
      const int32_t* pay = nullptr;
      int64_t sn = N;
      if (S2 > 1) {
        MOL_TRY(launch_select_segments(p.fb_scores, N, N, nb, S2, kk, p.fb_seg_scores, p.fb_seg_idx, fl, st));
        sel = p.fb_seg_scores;
        pay = p.fb_seg_idx;
        sn = (int64_t)S2 * kk;
      }
      MOL_TRY(launch_select_final_i32(sel, pay, sn, sn, nb, kk, o_scores + (int64_t)b1 * kk,
                                      o_idx ? o_idx + (int64_t)b1 * kk : nullptr, o_ids ? o_ids + (int64_t)b1 * kk : nullptr,
                                      id_map, fl, st));
    }
  }
  return MOL_OK;
}

}  // namespace mol
