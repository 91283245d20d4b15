// Index build on the sm_100a tensor cores (item projection + l2-norm, item-only gating MLP
// over the whole corpus - rails/similarities/mol/item_embeddings_fns.py:165-182, similarity_fn.py:170-171).
//
// linear_x3_kernel computes  C[M, nc] = A[M, K] . W[n0 : n0 + nc, :]^T  with fp32-grade accuracy on tcgen05 kind::tf32:
// both operands are split  x = hi + lo  (hi = the upper 19 bits, lo = x - hi, exact) and three MMAs per K step
// accumulate  A_hi W_hi + A_lo W_hi + A_hi W_lo  in fp32 (the dropped lo x lo term and the truncation of the lo parts are
// ~2^-22 relative, the size of the rounding noise of an fp32 fmaf chain of the same length).  The exact caches
// (xsub_f32 / gi_f32) therefore agree with the CUDA-core build to ~1e-6 and every parity bound on the scores holds.
//
// One persistent CTA per SM, 512 threads (setmaxnreg: 72 registers for the control warps, 184 for the epilogue warps):
//   warp 0        TMA producer: W_hi / W_lo of this launch once (K / 32 boxes of nc rows each), then the A tiles (128 rows)
//                 box by box into a ring of raw fp32 stages;
//   warp 1        MMA issuer: per box 4 K steps x 3 MMAs into one of two 256-column TMEM accumulators;
//   warps 2..5    splitter: raw box -> (hi, lo) boxes.  The split is elementwise and written at the same offsets, so the
//                 128-byte swizzle of the TMA box is preserved without knowing it;
//   warps 8..15   epilogue (two warps per TMEM lane quarter, each half of the columns), TMEM lane = row:
//                   PROJ  + bias, l2-norm over groups of d columns, fp32 row and fp16 row (pad rows: zeros);
//                   SILU  + bias, v / (1 + exp(-v)), fp32 row (the hidden layer of the gating MLP);
//                   GI    columns [0, L): fp32 row; columns [L, 2L) (the same weights with rows permuted into the coarse
//                         kernel's logit order): fp16 row with the overflow flag.
// HBM traffic of a 1M-item 8x8x32 build: 256 MB raw items x 2, 1 GB + 0.5 GB X_sub, 0.5 GB hidden x 2, 0.25 + 0.125 GB
// GI = 3.4 GB -> ~0.55 ms at 6.5 TB/s; the CUDA-core build took 2.8 ms.
#include <cuda.h>
#include <math_constants.h>

#include "common.cuh"
#include "mol_coarse.cuh"
#include "sm100_ptx.cuh"

namespace mol {
using namespace sm100;

namespace {

constexpr int LX_TILE = 128;
constexpr int LX_SPLIT_WARPS = 4, LX_EPI_WARPS = 8;
// 16 warps = 4 warpgroups: {producer, MMA issuer, splitter x 2}, {splitter x 2, 2 idle}, {epilogue x 4}, {epilogue x 4};
// setmaxnreg moves the registers of the first two (72 each) to the epilogue warps (184 each)
constexpr int LX_THREADS = 512;
constexpr int LX_CTL_REGS = 72, LX_EPI_REGS = 184;
static_assert(256 * LX_CTL_REGS + 256 * LX_EPI_REGS <= 65536, "register pool over-committed");
constexpr int LX_EPI_THREADS = 32 * LX_EPI_WARPS;
constexpr int LX_EPI_WARP0 = 8;
constexpr int LX_MAX_RAW = 6, LX_MAX_SPLIT = 2;  // pipeline stages (how many fit is decided per launch)
constexpr int LX_BOX = LX_TILE * 128;  // 128 rows x 32 fp32
constexpr int LX_SMEM_LIMIT = 227 * 1024;
constexpr int LX_STG_PITCH = 36;                          // floats per staged row (16-byte aligned, conflict-free)
constexpr int LX_STG_BYTES = 32 * LX_STG_PITCH * 4;       // per epilogue warp

enum { LX_PROJ = 0, LX_SILU = 1, LX_GI = 2 };

struct LxParams {
  const float* bias;  // (n_total) or nullptr
  int mode;
  int64_t M;          // rows of A
  int64_t M_pad;      // rows of the fp16 output (pad rows are written as zeros)
  int ks;             // K / 32
  int n0, nc;         // output columns [n0, n0 + nc) of this launch; nc % 32 == 0, <= 256
  int tiles;
  float* out_f32;     // PROJ: (M, pitch_f32) normalised; SILU: hidden; GI: gi_f32
  int64_t pitch_f32;
  uint16_t* out_half; // PROJ: (M_pad, pitch_half); GI: (M_pad, pitch_half) from columns >= n_f32; may be nullptr
  int64_t pitch_half;
  int n_f32;          // GI: columns [0, n_f32) of the whole layer go to out_f32, the rest (minus n_f32) to out_half
  int d;              // PROJ: group length
  float eps;
  int32_t* overflow;  // GI: set when a value does not fit fp16
  int raw_stages, split_stages;
  int staged;         // PROJ / SILU: rows go through a warp-private shared-memory transpose so that global stores are coalesced
  uint32_t idesc;
};

struct LxBars {
  uint64_t raw_full[LX_MAX_RAW], raw_empty[LX_MAX_RAW], split_full[LX_MAX_SPLIT], split_empty[LX_MAX_SPLIT];
  uint64_t w_full, acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

__host__ __device__ constexpr uint32_t lx_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void lx_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void st_global_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}

__global__ void __launch_bounds__(LX_THREADS, 1)
linear_x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmWhi,
                 const __grid_constant__ CUtensorMap tmWlo, const LxParams P) {
  extern __shared__ unsigned char lx_smem_raw[];
  unsigned char* smem =
      reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(lx_smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t wbox = (uint32_t)P.nc * 128u;  // one 32-column box of nc weight rows (nc % 32 == 0: 4 KB multiples)
  unsigned char* sWhi = smem;
  unsigned char* sWlo = sWhi + (size_t)P.ks * wbox;
  unsigned char* sRaw = sWlo + (size_t)P.ks * wbox;
  unsigned char* sSplit = sRaw + (size_t)P.raw_stages * LX_BOX;  // stage s: hi box at 2 s, lo box at 2 s + 1
  unsigned char* sStage = sSplit + (size_t)P.split_stages * 2 * LX_BOX;
  LxBars* bars = reinterpret_cast<LxBars*>(sStage + (P.staged ? LX_EPI_WARPS * LX_STG_BYTES : 0));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < LX_MAX_RAW; ++s) {
      mbar_init(&bars->raw_full[s], 1);
      mbar_init(&bars->raw_empty[s], 32 * LX_SPLIT_WARPS);
    }
    for (int s = 0; s < LX_MAX_SPLIT; ++s) {
      mbar_init(&bars->split_full[s], 32 * LX_SPLIT_WARPS);
      mbar_init(&bars->split_empty[s], 1);
    }
    mbar_init(&bars->w_full, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars->acc_full[a], 1);
      mbar_init(&bars->acc_empty[a], LX_EPI_THREADS);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(&bars->tmem_base);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmWhi);
    tma_prefetch_desc(&tmWlo);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  // (each role changes its register budget inside its own branch, so that ptxas allocates that branch against it)
  if (warp < LX_EPI_WARP0) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(LX_CTL_REGS));
  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      mbar_arrive_expect_tx(&bars->w_full, 2u * (uint32_t)P.ks * wbox);
      for (int ks = 0; ks < P.ks; ++ks) {
        tma_load_2d(sWhi + (size_t)ks * wbox, &tmWhi, &bars->w_full, ks * 32, P.n0);
        tma_load_2d(sWlo + (size_t)ks * wbox, &tmWlo, &bars->w_full, ks * 32, P.n0);
      }
      int ib = 0;
      for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
        for (int ks = 0; ks < P.ks; ++ks, ++ib) {
          const int s = ib % P.raw_stages;
          mbar_wait(&bars->raw_empty[s], ((uint32_t)(ib / P.raw_stages) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&bars->raw_full[s], LX_BOX);
          tma_load_2d(sRaw + (size_t)s * LX_BOX, &tmA, &bars->raw_full[s], ks * 32, tile * LX_TILE);
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    const uint32_t sSa = smem_u32(sSplit), sWhia = smem_u32(sWhi), sWloa = smem_u32(sWlo);
    mbar_wait(&bars->w_full, 0);
    tc_fence_after();
    int ib = 0, it = 0;
    for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      mbar_wait(&bars->acc_empty[acc], (((uint32_t)(it >> 1)) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem + (uint32_t)acc * 256u;
      for (int ks = 0; ks < P.ks; ++ks, ++ib) {
        const int s = ib % P.split_stages;
        mbar_wait(&bars->split_full[s], (uint32_t)(ib / P.split_stages) & 1u);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t a_hi = sSa + (uint32_t)(2 * s) * LX_BOX, a_lo = a_hi + LX_BOX;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t dah = make_smem_desc(a_hi + kk * 32, 16, 1024, 2);
            const uint64_t dal = make_smem_desc(a_lo + kk * 32, 16, 1024, 2);
            const uint64_t dwh = make_smem_desc(sWhia + (uint32_t)ks * wbox + kk * 32, 16, 1024, 2);
            const uint64_t dwl = make_smem_desc(sWloa + (uint32_t)ks * wbox + kk * 32, 16, 1024, 2);
            lx_mma(d_tmem, dal, dwh, P.idesc, (ks | kk) != 0);  // small terms first
            lx_mma(d_tmem, dah, dwl, P.idesc, 1);
            lx_mma(d_tmem, dah, dwh, P.idesc, 1);
          }
          umma_commit(&bars->split_empty[s]);
          if (ks == P.ks - 1) umma_commit(&bars->acc_full[acc]);
        }
        __syncwarp();
      }
    }
  } else if (warp < 2 + LX_SPLIT_WARPS) {
    // =============================== splitter (warps 2..5) ===============================
    const int t = tid - 64;  // 0..127
    int ib = 0;
    for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
      for (int ks = 0; ks < P.ks; ++ks, ++ib) {
        const int s = ib % P.raw_stages, s2 = ib % P.split_stages;
        mbar_wait(&bars->raw_full[s], (uint32_t)(ib / P.raw_stages) & 1u);
        mbar_wait(&bars->split_empty[s2], ((uint32_t)(ib / P.split_stages) & 1u) ^ 1u);
        const uint4* src = reinterpret_cast<const uint4*>(sRaw + (size_t)s * LX_BOX);
        uint4* dhi = reinterpret_cast<uint4*>(sSplit + (size_t)(2 * s2) * LX_BOX);
        uint4* dlo = reinterpret_cast<uint4*>(sSplit + (size_t)(2 * s2 + 1) * LX_BOX);
#pragma unroll
        constexpr int NT = 32 * LX_SPLIT_WARPS;
#pragma unroll 4
        for (int i = 0; i < LX_BOX / 16 / NT; ++i) {
          const uint4 v = src[i * NT + t];
          uint4 h, l;
          h.x = v.x & 0xFFFFE000u;
          h.y = v.y & 0xFFFFE000u;
          h.z = v.z & 0xFFFFE000u;
          h.w = v.w & 0xFFFFE000u;
          l.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(h.x));
          l.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(h.y));
          l.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(h.z));
          l.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(h.w));
          dhi[i * NT + t] = h;
          dlo[i * NT + t] = l;
        }
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the tcgen05 operand reads
        mbar_arrive(&bars->split_full[s2]);
        mbar_arrive(&bars->raw_empty[s]);
      }
    }
  } else if (warp >= LX_EPI_WARP0) {
    // =============================== epilogue (warps 8..15) ===============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(LX_EPI_REGS));
    const int quarter = warp & 3;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const int nch = P.nc / 32;
    const int per_half = (nch + 1) / 2;
    const int half = (warp - LX_EPI_WARP0) >> 2;
    const int c_lo = half * per_half, c_hi = (c_lo + per_half < nch) ? c_lo + per_half : nch;
    int it = 0;
    for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      mbar_wait(&bars->acc_full[acc], ((uint32_t)(it >> 1)) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem + (uint32_t)acc * 256u + lane_base;
      const int64_t row = (int64_t)tile * LX_TILE + quarter * 32 + lane;
      const bool live = row < P.M;
      uint32_t v[32];
      // rows of this warp's 32-row group that exist (fp32 outputs) / that exist in the padded fp16 outputs
      const int64_t row0 = (int64_t)tile * LX_TILE + quarter * 32;
      const int live_rows = (int)(P.M - row0 < 0 ? 0 : (P.M - row0 > 32 ? 32 : P.M - row0));
      const int pad_rows = (int)(P.M_pad - row0 < 0 ? 0 : (P.M_pad - row0 > 32 ? 32 : P.M_pad - row0));
      float* stg = reinterpret_cast<float*>(sStage + (size_t)(warp - LX_EPI_WARP0) * LX_STG_BYTES);
      // 32 x 32 fp32 chunk (lane = row) -> global rows, 128-byte segments per row, through the warp's staging buffer:
      // stg_put4 writes four consecutive values of this lane's row, flush_f32 sends the staged chunk out
      auto stg_put4 = [&](int j4, float a0, float a1, float a2, float a3) __attribute__((always_inline)) {
        *reinterpret_cast<float4*>(stg + lane * LX_STG_PITCH + 4 * j4) = make_float4(a0, a1, a2, a3);
      };
      auto flush_f32 = [&](float* g, int64_t pitch) __attribute__((always_inline)) {
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = (lane >> 3) + 4 * i, c4 = lane & 7;
          const float4 t = *reinterpret_cast<const float4*>(stg + r * LX_STG_PITCH + 4 * c4);
          if (r < live_rows)
            st_global_v4(g + (int64_t)r * pitch + 4 * c4, __float_as_uint(t.x), __float_as_uint(t.y), __float_as_uint(t.z),
                         __float_as_uint(t.w));
        }
        __syncwarp();
      };
      // the same chunk as packed fp16 pairs (h[i] = columns 2i, 2i + 1): 64-byte segments per row (staged rows of 80 bytes)
      auto put_f16 = [&](const uint32_t (&h)[16], uint16_t* g, int64_t pitch) __attribute__((always_inline)) {
        uint4* s4 = reinterpret_cast<uint4*>(stg);
#pragma unroll
        for (int j = 0; j < 4; ++j) s4[lane * 5 + j] = make_uint4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = (lane >> 2) + 8 * i, pc = lane & 3;
          const uint4 t = s4[r * 5 + pc];
          if (r < pad_rows) st_global_v4(g + (int64_t)r * pitch + 8 * pc, t.x, t.y, t.z, t.w);
        }
        __syncwarp();
      };
      if (P.mode == LX_PROJ) {
        const int cpg = P.d / 32;  // chunks per group
        for (int g0 = c_lo; g0 < c_hi; g0 += cpg) {
          float ss = 0.f;
          for (int c = g0; c < g0 + cpg; ++c) {
            tmem_ld_x32(taddr + (uint32_t)c * 32u, v);
            tmem_ld_wait_bind32(v);
            const float4* b4 = reinterpret_cast<const float4*>(P.bias + P.n0 + c * 32);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 t = __ldg(b4 + j4);
              const float x0 = __uint_as_float(v[4 * j4]) + t.x, x1 = __uint_as_float(v[4 * j4 + 1]) + t.y;
              const float x2 = __uint_as_float(v[4 * j4 + 2]) + t.z, x3 = __uint_as_float(v[4 * j4 + 3]) + t.w;
              ss = fmaf(x0, x0, ss);
              ss = fmaf(x1, x1, ss);
              ss = fmaf(x2, x2, ss);
              ss = fmaf(x3, x3, ss);
            }
          }
          const float nrm = fmaxf(sqrtf(ss), P.eps);
          const float r = 1.0f / nrm;
          for (int c = g0; c < g0 + cpg; ++c) {
            tmem_ld_x32(taddr + (uint32_t)c * 32u, v);
            tmem_ld_wait_bind32(v);
            const float4* b4 = reinterpret_cast<const float4*>(P.bias + P.n0 + c * 32);
            uint32_t h[16];
            // x / nrm, correctly rounded: quotient estimate + one Newton step on it
            auto quot = [&](float x) __attribute__((always_inline)) {
              const float q = x * r;
              return live ? fmaf(fmaf(-q, nrm, x), r, q) : 0.f;
            };
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 t = __ldg(b4 + j4);
              const float q0 = quot(__uint_as_float(v[4 * j4]) + t.x), q1 = quot(__uint_as_float(v[4 * j4 + 1]) + t.y);
              const float q2 = quot(__uint_as_float(v[4 * j4 + 2]) + t.z), q3 = quot(__uint_as_float(v[4 * j4 + 3]) + t.w);
              stg_put4(j4, q0, q1, q2, q3);
              h[2 * j4] = pack_half2(q0, q1);
              h[2 * j4 + 1] = pack_half2(q2, q3);
            }
            const int col = P.n0 + c * 32;
            flush_f32(P.out_f32 + row0 * P.pitch_f32 + col, P.pitch_f32);
            if (P.out_half != nullptr) put_f16(h, P.out_half + row0 * P.pitch_half + col, P.pitch_half);
          }
        }
      } else {
        uint32_t v2[32];
        auto chunk = [&](const uint32_t (&u)[32], int c) __attribute__((always_inline)) {
          const int col = P.n0 + c * 32;  // first column of the chunk within the whole layer
          if (P.mode == LX_SILU) {
            // x / (1 + exp(-x)) from ex2.approx / rcp.approx (~2^-21 relative, below the fp32 rounding noise of the layer
            // that follows; expf + the IEEE division cost ~28 dependent instructions per element and made this
            // epilogue the bottleneck of the build: 394 us of 1.19 ms)
            const float4* b4 = reinterpret_cast<const float4*>(P.bias + col);
            auto silu = [&](float x) __attribute__((always_inline)) { return __fdividef(x, 1.f + __expf(-x)); };
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 t = __ldg(b4 + j4);
              stg_put4(j4, silu(__uint_as_float(u[4 * j4]) + t.x), silu(__uint_as_float(u[4 * j4 + 1]) + t.y),
                       silu(__uint_as_float(u[4 * j4 + 2]) + t.z), silu(__uint_as_float(u[4 * j4 + 3]) + t.w));
            }
            flush_f32(P.out_f32 + row0 * P.pitch_f32 + col, P.pitch_f32);
          } else if (col < P.n_f32) {  // GI, fp32 part (direct stores: this launch has no staging buffer)
            if (live) {
              float* o = P.out_f32 + row * P.pitch_f32 + col;
#pragma unroll
              for (int j = 0; j < 32; j += 4) st_global_v4(o + j, u[j], u[j + 1], u[j + 2], u[j + 3]);
            }
          } else if (row < P.M_pad) {  // GI, fp16 image (rows of W permuted into the coarse kernel's logit order)
            uint16_t* o = P.out_half + row * P.pitch_half + (col - P.n_f32);
            bool bad = false;
            auto val = [&](int j) __attribute__((always_inline)) {
              const float y = live ? __uint_as_float(u[j]) : 0.f;
              bad |= !(fabsf(y) <= 65504.f);
              return y;
            };
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const uint32_t p0 = pack_half2(val(j), val(j + 1)), p1 = pack_half2(val(j + 2), val(j + 3));
              const uint32_t p2 = pack_half2(val(j + 4), val(j + 5)), p3 = pack_half2(val(j + 6), val(j + 7));
              st_global_v4(o + j, p0, p1, p2, p3);
            }
            if (bad) atomicOr(P.overflow, 1);
          }
        };
        if (c_lo < c_hi) tmem_ld_x32(taddr + (uint32_t)c_lo * 32u, v);
        for (int c = c_lo; c < c_hi; c += 2) {  // the next chunk's TMEM load is in flight while this one is converted
          tmem_ld_wait_bind32(v);
          if (c + 1 < c_hi) tmem_ld_x32(taddr + (uint32_t)(c + 1) * 32u, v2);
          chunk(v, c);
          if (c + 1 < c_hi) {
            tmem_ld_wait_bind32(v2);
            if (c + 2 < c_hi) tmem_ld_x32(taddr + (uint32_t)(c + 2) * 32u, v);
            chunk(v2, c + 1);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&bars->acc_empty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// W (rows, K) fp32 -> hi / lo; row r of the outputs = row perm(r) of W when PQ > 0 (r = m * PQ + n <- n * PX + m)
__global__ void split_weights_kernel(const float* __restrict__ W, int rows, int K, int PQ, int PX, float* __restrict__ hi,
                                     float* __restrict__ lo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * K) return;
  const int r = i / K, k = i - r * K;
  const int src = PQ > 0 ? (r % PQ) * PX + (r / PQ) : r;
  const float x = W[(size_t)src * K + k];
  const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  hi[i] = h;
  lo[i] = x - h;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int lx_encode(CUtensorMap* m, const float* base, uint64_t cols, uint64_t rows, uint32_t box_rows) {
  static PFN_encodeTiled encode = nullptr;
  if (!encode) {
    cudaDriverEntryPointQueryResult qres;
    MOL_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
    MOL_CHECK_ARG(encode != nullptr, "cuTensorMapEncodeTiled not available");
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * sizeof(float)};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (linear_x3) failed (%d)", (int)r);
    return MOL_ERR_CUDA;
  }
  return MOL_OK;
}

// Shared-memory plan of a launch: weights (hi + lo) of the column chunk | raw stages | split stages | staging | barriers.
// Staged modes (PROJ / SILU) keep the chunk <= 128 columns so that the transposing store buffers and a deeper TMA ring fit.
constexpr size_t LX_FIXED = 2048;  // alignment slack + barriers
int lx_chunk_cols(int K, bool staged) {
  const size_t stage_bytes = staged ? (size_t)LX_EPI_WARPS * LX_STG_BYTES : 0;
  const size_t avail = LX_SMEM_LIMIT - LX_FIXED - stage_bytes - 2 * (size_t)LX_BOX - 2 * (size_t)LX_BOX;  // >= 2 raw + 1 split
  int nc = (int)(avail / ((size_t)K * 4 * 2));
  const int cap = staged ? 128 : 256;
  if (nc > cap) nc = cap;
  return nc / 64 * 64;
}

// One layer: C[:, all n_rows_w columns] in column chunks.  W_hi / W_lo: (n_rows_w, K).
int lx_layer(const float* A, int64_t M, int K, const float* W_hi, const float* W_lo, int n_rows_w, LxParams P,
             cudaStream_t st) {
  static int sms = 0;
  if (sms == 0) {
    MOL_CUDA(cudaFuncSetAttribute(linear_x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LX_SMEM_LIMIT));
    int dev = 0, n = 148;
    MOL_CUDA(cudaGetDevice(&dev));
    MOL_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    sms = n;
  }
  P.staged = (P.mode == LX_PROJ || P.mode == LX_SILU) ? 1 : 0;
  const int chunk = lx_chunk_cols(K, P.staged != 0);
  MOL_CHECK_ARG(chunk >= 64, "linear_x3: K=%d too large", K);
  const int64_t rows = P.M_pad > M ? P.M_pad : M;
  P.M = M;
  P.ks = K / 32;
  P.tiles = (int)((rows + LX_TILE - 1) / LX_TILE);
  CUtensorMap tmA, tmWhi, tmWlo;
  MOL_TRY(lx_encode(&tmA, A, (uint64_t)K, (uint64_t)M, LX_TILE));
  for (int n0 = 0; n0 < n_rows_w; n0 += chunk) {
    const int nc = (n_rows_w - n0 < chunk) ? (n_rows_w - n0) : chunk;
    P.n0 = n0;
    P.nc = nc;
    P.idesc = lx_idesc_tf32(LX_TILE, nc);
    MOL_TRY(lx_encode(&tmWhi, W_hi, (uint64_t)K, (uint64_t)n_rows_w, (uint32_t)nc));
    MOL_TRY(lx_encode(&tmWlo, W_lo, (uint64_t)K, (uint64_t)n_rows_w, (uint32_t)nc));
    // what is left after the weights goes to the pipeline: a second split stage when there is room for >= 2 raw stages
    // besides it, the rest to the TMA ring (bytes in flight are what the HBM-bound layers need)
    const size_t used = LX_FIXED + 2 * (size_t)P.ks * nc * 128 + (P.staged ? (size_t)LX_EPI_WARPS * LX_STG_BYTES : 0);
    const int boxes = (int)(((size_t)LX_SMEM_LIMIT - used) / LX_BOX);
    MOL_CHECK_ARG(boxes >= 4, "linear_x3: shared memory (K=%d, %d columns)", K, nc);
    P.split_stages = boxes >= 6 ? 2 : 1;
    int raw = boxes - 2 * P.split_stages;
    P.raw_stages = raw > LX_MAX_RAW ? LX_MAX_RAW : raw;
    const size_t smem = used + (size_t)(P.raw_stages + 2 * P.split_stages) * LX_BOX;
    const int grid = P.tiles < sms ? P.tiles : sms;
    linear_x3_kernel<<<grid, LX_THREADS, smem, st>>>(tmA, tmWhi, tmWlo, P);
    MOL_LAUNCH_CHECK();
  }
  return MOL_OK;
}

}  // namespace

bool index_build_x3_supported(const mol_shape_t& s, const mol_weights_t& w, const mol_index_t& ix) {
  const char* e = getenv("MOL_B200_INDEX_X3");
  if (e && atoi(e) == 0) return false;
  Dims D = dims_of(s);
  if (!(D.Dx % 32 == 0 && D.Dx <= 128 && D.Hgi % 32 == 0 && D.Hgi <= 128 && D.L % 32 == 0)) return false;
  if (!(D.d == 32 || D.d == 64 || D.d == 128)) return false;
  // every launch of the projection (a column chunk, or the remainder) must give each epilogue half whole l2-norm groups
  const int chunk = lx_chunk_cols(D.Dx, true);
  if (chunk < 2 * D.d || chunk % (2 * D.d) != 0 || (D.Px * D.d) % (2 * D.d) != 0) return false;
  // TMA reads raw_items, the epilogues read the biases as float4
  auto aligned16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  return ix.num_items >= 1024 && aligned16(ix.raw_items) && aligned16(w.x_b) && aligned16(w.gi_b1);
}

size_t index_build_x3_workspace_bytes(const mol_shape_t& s, int64_t N) {
  Dims D = dims_of(s);
  size_t b = align_up((size_t)N * D.Hgi * sizeof(float), 256);                    // hidden layer of the gating MLP
  b += 2 * align_up((size_t)D.Px * D.d * D.Dx * sizeof(float), 256);             // W_x hi / lo
  b += 2 * align_up((size_t)D.Hgi * D.Dx * sizeof(float), 256);                  // W_gi1 hi / lo
  b += 2 * align_up((size_t)2 * D.L * D.Hgi * sizeof(float), 256);               // [W_gi2 ; permuted W_gi2] hi / lo
  return b + 1024;
}

int index_build_x3(const mol_shape_t& s, const mol_weights_t& w, const mol_index_t& ix, void* workspace, cudaStream_t st) {
  Dims D = dims_of(s);
  const int64_t N = ix.num_items, Np = (N + 127) / 128 * 128;
  const bool tensor = coarse_supported(s);
  Arena a(workspace, (size_t)-1);
  float* hidden = a.take<float>((size_t)N * D.Hgi);
  const int nx = D.Px * D.d;
  float* wx_hi = a.take<float>((size_t)nx * D.Dx);
  float* wx_lo = a.take<float>((size_t)nx * D.Dx);
  float* w1_hi = a.take<float>((size_t)D.Hgi * D.Dx);
  float* w1_lo = a.take<float>((size_t)D.Hgi * D.Dx);
  float* w2_hi = a.take<float>((size_t)2 * D.L * D.Hgi);
  float* w2_lo = a.take<float>((size_t)2 * D.L * D.Hgi);
  auto split = [&](const float* W, int rows, int K, int PQ, int PX, float* hi, float* lo) -> int {
    split_weights_kernel<<<(rows * K + 255) / 256, 256, 0, st>>>(W, rows, K, PQ, PX, hi, lo);
    MOL_LAUNCH_CHECK();
    return MOL_OK;
  };
  MOL_TRY(split(w.x_w, nx, D.Dx, 0, 0, wx_hi, wx_lo));
  MOL_TRY(split(w.gi_w1, D.Hgi, D.Dx, 0, 0, w1_hi, w1_lo));
  MOL_TRY(split(w.gi_w2, D.L, D.Hgi, 0, 0, w2_hi, w2_lo));
  if (tensor)
    MOL_TRY(split(w.gi_w2, D.L, D.Hgi, D.Pq, D.Px, w2_hi + (size_t)D.L * D.Hgi, w2_lo + (size_t)D.L * D.Hgi));
  MOL_CUDA(cudaMemsetAsync(ix.half_overflow, 0, sizeof(int32_t), st));
  {  // X_sub = l2norm(reshape(W_x e + b_x))  (item_embeddings_fns.py:165-182)
    LxParams P{};
    P.bias = w.x_b;
    P.mode = LX_PROJ;
    P.M_pad = Np;
    P.out_f32 = ix.xsub_f32;
    P.pitch_f32 = nx;
    P.out_half = ix.xsub_half;
    P.pitch_half = nx;
    P.d = D.d;
    P.eps = s.eps;
    MOL_TRY(lx_layer(ix.raw_items, N, D.Dx, wx_hi, wx_lo, nx, P, st));
  }
  {  // hidden = silu(W_gi1 e + b_gi1)  (similarity_fn.py:170-171)
    LxParams P{};
    P.bias = w.gi_b1;
    P.mode = LX_SILU;
    P.M_pad = 0;
    P.out_f32 = hidden;
    P.pitch_f32 = D.Hgi;
    MOL_TRY(lx_layer(ix.raw_items, N, D.Dx, w1_hi, w1_lo, D.Hgi, P, st));
  }
  {  // GI = W_gi2 hidden (no bias) + its fp16 image in the coarse kernel's logit order
    LxParams P{};
    P.mode = LX_GI;
    P.M_pad = tensor ? Np : 0;
    P.out_f32 = ix.gi_f32;
    P.pitch_f32 = D.L;
    P.out_half = tensor ? ix.gi_half : nullptr;
    P.pitch_half = D.L;
    P.n_f32 = D.L;
    P.overflow = ix.half_overflow;
    MOL_TRY(lx_layer(hidden, N, D.Hgi, w2_hi, w2_lo, tensor ? 2 * D.L : D.L, P, st));
  }
  return MOL_OK;
}

}  // namespace mol
