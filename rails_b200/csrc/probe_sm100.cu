// Hardware probes for the tcgen05 / TMEM / TMA building blocks of the coarse kernel.  TEST-ONLY
// (built into libmol_probe.so, driven by tools/run_probe.py): each probe exercises one assumption
// about descriptor layouts in isolation and returns raw results for comparison on the host.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "sm100_ptx.cuh"

using namespace sm100;

// canonical no-swizzle K-major layout: element (r, k) of an R x K bf16 tile
//   byte offset = (r/8) * (K/8)*128 + (k/8)*128 + (r%8)*16 + (k%8)*2      => LBO = 128, SBO = (K/8)*128
__device__ __forceinline__ uint32_t nosw_off(int r, int k, int K) {
  return (uint32_t)((r >> 3) * (K >> 3) * 128 + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2);
}

// ------------------------------------------------------------------------------------------------
// Probe 1: SS MMA (no-swizzle A,B from st.shared) -> TMEM -> tcgen05.ld ; then TS MMA with A = bf16(D1)
// written back to TMEM via tcgen05.st.
//   A  (128, K1) bf16 row-major, B1 (N1, K1) bf16, B2 (N2, N1) bf16
//   out1 (128, N1) fp32 = A B1^T ; out2 (128, N2) fp32 = bf16(out1) B2^T
// ------------------------------------------------------------------------------------------------
template <int K1, int N1, int N2>
__global__ void __launch_bounds__(128) probe_mma_kernel(const __nv_bfloat16* __restrict__ A,
                                                        const __nv_bfloat16* __restrict__ B1,
                                                        const __nv_bfloat16* __restrict__ B2, float* out1,
                                                        float* out2) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sA = smem;                     // 128 x K1
  unsigned char* sB1 = sA + 128 * K1 * 2;       // N1 x K1
  unsigned char* sB2 = sB1 + N1 * K1 * 2;       // N2 x N1
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid / 32;

  for (int e = tid; e < 128 * K1; e += 128) {
    int r = e / K1, k = e % K1;
    *reinterpret_cast<__nv_bfloat16*>(sA + nosw_off(r, k, K1)) = A[e];
  }
  for (int e = tid; e < N1 * K1; e += 128) {
    int r = e / K1, k = e % K1;
    *reinterpret_cast<__nv_bfloat16*>(sB1 + nosw_off(r, k, K1)) = B1[e];
  }
  for (int e = tid; e < N2 * N1; e += 128) {
    int r = e / N1, k = e % N1;
    *reinterpret_cast<__nv_bfloat16*>(sB2 + nosw_off(r, k, N1)) = B2[e];
  }
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t D1 = tmem, D2 = tmem + 128, A2 = tmem + 256;

  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N1);
#pragma unroll
    for (int ks = 0; ks < K1 / 16; ++ks) {
      uint64_t da = make_smem_desc(smem_u32(sA) + ks * 256, 128, (K1 / 8) * 128, 0);
      uint64_t db = make_smem_desc(smem_u32(sB1) + ks * 256, 128, (K1 / 8) * 128, 0);
      umma_ss(D1, da, db, idesc, ks > 0);
    }
    umma_commit(&bar[0]);
  }
  mbar_wait(&bar[0], 0);
  tc_fence_after();

  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const int row = tid;
  // read D1 row, write out1, pack to bf16 and store as the A operand of the second MMA
  for (int c0 = 0; c0 < N1; c0 += 16) {
    uint32_t v[16];
    tmem_ld_x16(D1 + lane_base + c0, v);
    tmem_ld_wait();
    uint32_t p[8];
#pragma unroll
    for (int j = 0; j < 16; ++j) out1[row * N1 + c0 + j] = __uint_as_float(v[j]);
#pragma unroll
    for (int j = 0; j < 8; ++j) p[j] = pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
    tmem_st_x8(A2 + lane_base + c0 / 2, p);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N2);
#pragma unroll
    for (int ks = 0; ks < N1 / 16; ++ks) {
      uint64_t db = make_smem_desc(smem_u32(sB2) + ks * 256, 128, (N1 / 8) * 128, 0);
      umma_ts(D2, A2 + ks * 8, db, idesc, ks > 0);
    }
    umma_commit(&bar[1]);
  }
  mbar_wait(&bar[1], 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N2; c0 += 16) {
    uint32_t v[16];
    tmem_ld_x16(D2 + lane_base + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) out2[row * N2 + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------
// Probe 2: TMA (SWIZZLE_128B, box 64 cols x 128 rows) + SW128 K-major A descriptors with sub-atom K offsets.
//   X (rows, 256) bf16 global via tensor map; Q (16, 32) bf16; for m in 0..7:
//   out[m] (128, 16) fp32 = X[tile*128 + r, 32m : 32m+32] . Q^T
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) probe_tma_kernel(const __grid_constant__ CUtensorMap tmap,
                                                        const __nv_bfloat16* __restrict__ Q, float* out,
                                                        int tile) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sX = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~(uintptr_t)1023);
  unsigned char* sQ = sX + 4 * 16384;  // 16 x 32 bf16 no-swizzle
  __shared__ uint64_t bar_tma, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid / 32;
  for (int e = tid; e < 16 * 32; e += 128) {
    int r = e / 32, k = e % 32;
    *reinterpret_cast<__nv_bfloat16*>(sQ + nosw_off(r, k, 32)) = Q[e];
  }
  if (tid == 0) {
    mbar_init(&bar_tma, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<128>(&tmem_base_s);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    mbar_arrive_expect_tx(&bar_tma, 4 * 16384);
    for (int bx = 0; bx < 4; ++bx) tma_load_2d(sX + bx * 16384, &tmap, &bar_tma, bx * 64, tile * 128);
    mbar_wait(&bar_tma, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, 16);
    for (int m = 0; m < 8; ++m) {
      for (int ks = 0; ks < 2; ++ks) {
        uint32_t a_addr = smem_u32(sX) + (m / 2) * 16384 + (m % 2) * 64 + ks * 32;
        uint64_t da = make_smem_desc(a_addr, 16, 1024, 2);
        uint64_t db = make_smem_desc(smem_u32(sQ) + ks * 256, 128, (32 / 8) * 128, 0);
        umma_ss(tmem + m * 16, da, db, idesc, ks > 0);
      }
    }
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  for (int m = 0; m < 8; ++m) {
    uint32_t v[16];
    tmem_ld_x16(tmem + lane_base + m * 16, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) out[(m * 128 + tid) * 16 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<128>(tmem);
}

// ------------------------------------------------------------------------------------------------
// Probe 3: MUFU / conversion throughput.  which: 0 tanh.f32, 1 tanh.bf16x2, 2 ex2.f32, 3 ex2.bf16x2,
// 4 FFMA chain (reference), 5 cvt.rn.bf16x2.f32.  out[0] = cycles (max over warps of block 0), out[1] = ops.
// ------------------------------------------------------------------------------------------------
__global__ void probe_mufu_kernel(int which, int iters, float seed, float* sink, long long* cycles) {
  float a0 = seed + threadIdx.x * 1e-3f, a1 = a0 + 0.1f, a2 = a0 + 0.2f, a3 = a0 + 0.3f;
  uint32_t u0 = __float_as_uint(a0), u1 = __float_as_uint(a1), u2 = __float_as_uint(a2), u3 = __float_as_uint(a3);
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (which == 0) {
      a0 = tanh_approx(a0); a1 = tanh_approx(a1); a2 = tanh_approx(a2); a3 = tanh_approx(a3);
    } else if (which == 1) {
      u0 = tanh_bf16x2(u0); u1 = tanh_bf16x2(u1); u2 = tanh_bf16x2(u2); u3 = tanh_bf16x2(u3);
    } else if (which == 2) {
      a0 = ex2_approx(a0); a1 = ex2_approx(a1); a2 = ex2_approx(a2); a3 = ex2_approx(a3);
    } else if (which == 3) {
      u0 = ex2_bf16x2(u0); u1 = ex2_bf16x2(u1); u2 = ex2_bf16x2(u2); u3 = ex2_bf16x2(u3);
    } else if (which == 4) {
      a0 = fmaf(a0, 1.0001f, 0.5f); a1 = fmaf(a1, 1.0001f, 0.5f); a2 = fmaf(a2, 1.0001f, 0.5f); a3 = fmaf(a3, 1.0001f, 0.5f);
    } else {
      u0 = pack_bf16x2(__uint_as_float(u0), a1); u1 = pack_bf16x2(__uint_as_float(u1), a2);
      u2 = pack_bf16x2(__uint_as_float(u2), a3); u3 = pack_bf16x2(__uint_as_float(u3), a0);
    }
  }
  long long t1 = clock64();
  __syncthreads();
  sink[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + __uint_as_float(u0 ^ u1 ^ u2 ^ u3);
  if (blockIdx.x == 0 && threadIdx.x == 0) cycles[0] = t1 - t0;
}


// ------------------------------------------------------------------------------------------------
// Probe 4: issue/pipe throughput of the packed and mixed instruction streams the v2 epilogue relies on.
// 8 independent chains per thread.  which: 0 ffma (f32), 1 ffma2 (fma.rn.f32x2), 2 hfma2 (f16x2),
// 3 cvt.rn.f16x2.f32, 4 min.xorsign.abs.f16x2, 5 tanh.approx.f16x2, 6 mix {1 tanh.f32 + 6 ffma},
// 7 mix {1 ex2.f32 + 3 ffma2}, 8 mix {1 tanh.f16x2 + 4 hfma2}, 9 fadd2.  ops counted = instructions.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hfma2_f16(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t tanh_f16x2_p(uint32_t a) {
  uint32_t d;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}
__device__ __forceinline__ uint32_t clamp_f16x2_p(uint32_t a, uint32_t c) {
  uint32_t d;
  asm("min.xorsign.abs.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t pack_f16x2_p(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__global__ void probe_pipe_kernel(int which, int iters, float seed, float* sink, long long* cycles) {
  float a[8];
  float2 f[8];
  uint32_t u[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a[j] = seed + threadIdx.x * 1e-3f + 0.01f * j;
    f[j] = make_float2(a[j], a[j] + 0.5f);
    u[j] = pack_f16x2_p(a[j], a[j] * 0.5f);
  }
  const float2 m2 = make_float2(0.999f, 1.0001f), c2 = make_float2(0.25f, 0.125f);
  const uint32_t hm = pack_f16x2_p(0.999f, 0.998f), hc = pack_f16x2_p(0.01f, 0.02f), h3 = pack_f16x2_p(3.f, 3.f);
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (which == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = fmaf(a[j], 0.999f, 0.25f);
    } else if (which == 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = __ffma2_rn(f[j], m2, c2);
    } else if (which == 2) {
#pragma unroll
      for (int j = 0; j < 8; ++j) u[j] = hfma2_f16(u[j], hm, hc);
    } else if (which == 3) {
#pragma unroll
      for (int j = 0; j < 8; ++j) u[j] = pack_f16x2_p(__uint_as_float(u[j]), a[j]);
    } else if (which == 4) {
#pragma unroll
      for (int j = 0; j < 8; ++j) u[j] = clamp_f16x2_p(u[j], h3);
    } else if (which == 5) {
#pragma unroll
      for (int j = 0; j < 8; ++j) u[j] = tanh_f16x2_p(u[j]);
    } else if (which == 6) {
      a[0] = tanh_approx(a[0]);
      a[1] = fmaf(a[1], 0.999f, 0.25f); a[2] = fmaf(a[2], 0.999f, 0.25f); a[3] = fmaf(a[3], 0.999f, 0.25f);
      a[4] = fmaf(a[4], 0.999f, 0.25f); a[5] = fmaf(a[5], 0.999f, 0.25f); a[6] = fmaf(a[6], 0.999f, 0.25f);
      a[7] = tanh_approx(a[7]);
      a[1] = fmaf(a[1], 0.999f, 0.25f); a[2] = fmaf(a[2], 0.999f, 0.25f); a[3] = fmaf(a[3], 0.999f, 0.25f);
      a[4] = fmaf(a[4], 0.999f, 0.25f); a[5] = fmaf(a[5], 0.999f, 0.25f); a[6] = fmaf(a[6], 0.999f, 0.25f);
    } else if (which == 7) {
      a[0] = ex2_approx(a[0]);
      f[1] = __ffma2_rn(f[1], m2, c2); f[2] = __ffma2_rn(f[2], m2, c2); f[3] = __ffma2_rn(f[3], m2, c2);
      a[7] = ex2_approx(a[7]);
      f[4] = __ffma2_rn(f[4], m2, c2); f[5] = __ffma2_rn(f[5], m2, c2); f[6] = __ffma2_rn(f[6], m2, c2);
    } else if (which == 8) {
      u[0] = tanh_f16x2_p(u[0]);
      u[1] = hfma2_f16(u[1], hm, hc); u[2] = hfma2_f16(u[2], hm, hc); u[3] = hfma2_f16(u[3], hm, hc); u[4] = hfma2_f16(u[4], hm, hc);
      u[7] = tanh_f16x2_p(u[7]);
      u[5] = hfma2_f16(u[5], hm, hc); u[6] = hfma2_f16(u[6], hm, hc); u[1] = hfma2_f16(u[1], hm, hc); u[2] = hfma2_f16(u[2], hm, hc);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = __fadd2_rn(f[j], c2);
    }
  }
  long long t1 = clock64();
  __syncthreads();
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) acc += a[j] + f[j].x + f[j].y + __uint_as_float(u[j]);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (blockIdx.x == 0 && threadIdx.x == 0) cycles[0] = t1 - t0;
}


// ------------------------------------------------------------------------------------------------
// Probe 5: TMEM load / store throughput.  Every warp of the block issues `iters` back-to-back
// tcgen05.ld (which 0: 32x32b.x32, 1: .x16) or tcgen05.st (2: .x32) on its own lane quarter.
// ------------------------------------------------------------------------------------------------
__global__ void probe_tmem_kernel(int which, int iters, float* sink, long long* cycles) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) & 3) * 64u;
  uint32_t v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = threadIdx.x + i;
  tmem_st_x32(base, v);
  tmem_st_x32(base + 32, v);
  tmem_st_wait();
  __syncthreads();
  long long t0 = clock64();
  uint32_t acc = 0;
  for (int i = 0; i < iters; ++i) {
    if (which == 0) {
      tmem_ld_x32(base + (i & 1) * 32, v);
      tmem_ld_wait();
      acc += v[0] ^ v[31];
    } else if (which == 1) {
      tmem_ld_x16(base + (i & 3) * 16, v);
      tmem_ld_wait();
      acc += v[0] ^ v[15];
    } else if (which == 2) {
      v[0] = acc + i;
      tmem_st_x32(base + (i & 1) * 32, v);
      tmem_st_wait();
    } else {  // 4 loads in flight before one wait
      uint32_t w[32];
      tmem_ld_x16(base, v);
      tmem_ld_x16(base + 16, v + 16);
      tmem_ld_x16(base + 32, w);
      tmem_ld_x16(base + 48, w + 16);
      tmem_ld_wait();
      acc += v[0] ^ v[31] ^ w[0] ^ w[31];
    }
  }
  long long t1 = clock64();
  __syncthreads();
  sink[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
  if (blockIdx.x == 0 && threadIdx.x == 0) cycles[0] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_slot);
}


// ------------------------------------------------------------------------------------------------
// Probe 6: tcgen05.mma issue / execution cost per instruction.  One converged warp issues `iters` MMAs
// (M = 128, K = 16, N = n) back to back, then commits.  mode 0: SS (A, B from smem), 1: TS (A from TMEM).
// out: cycles[0] = issue loop, cycles[1] = until the commit's mbarrier fires.
// ------------------------------------------------------------------------------------------------
template <int N, int mode>
__global__ void probe_mma_rate_kernel(int iters, long long* cycles) {
  extern __shared__ unsigned char smem_raw2[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw2) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + 256 * 32) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N);
    const uint32_t sA = smem_u32(smem), sB = smem_u32(smem + 16384);
    const uint64_t da = make_smem_desc(sA, 16, 1024, 2);
    const uint64_t db = make_smem_desc(sB, 128, 256, 0);
    const uint64_t db2 = make_smem_desc(sB + 256 * 16, 128, 256, 0);
    long long t0 = clock64();
    if (elect_one_sync()) {
      for (int i = 0; i < iters; i += 4) {  // branch-free body: `mode` is a template parameter
        if constexpr (mode == 0) {
#pragma unroll
          for (int r = 0; r < 4; ++r) umma_ss(tmem + 256, da, db, idesc, 1u);
        } else if constexpr (mode == 1) {
#pragma unroll
          for (int r = 0; r < 4; ++r) umma_ts(tmem + 256, tmem, db, idesc, 1u);
        } else if constexpr (mode == 2) {  // pairs sharing A through the collector: fill, then lastuse into another accumulator
          umma_ss_coll<1>(tmem + 256, da, db, idesc, 1u);
          umma_ss_coll<3>(tmem + 256 + N, da, db2, idesc, 1u);
          umma_ss_coll<1>(tmem + 256, da, db, idesc, 1u);
          umma_ss_coll<3>(tmem + 256 + N, da, db2, idesc, 1u);
        } else if constexpr (mode == 3) {  // the same pairs without the hint (two accumulators, same A)
          umma_ss(tmem + 256, da, db, idesc, 1u);
          umma_ss(tmem + 256 + N, da, db2, idesc, 1u);
          umma_ss(tmem + 256, da, db, idesc, 1u);
          umma_ss(tmem + 256 + N, da, db2, idesc, 1u);
        } else {  // groups of four: fill, use, use, lastuse
          umma_ss_coll<1>(tmem + 256, da, db, idesc, 1u);
          umma_ss_coll<2>(tmem + 256 + N, da, db2, idesc, 1u);
          umma_ss_coll<2>(tmem + 256, da, db, idesc, 1u);
          umma_ss_coll<3>(tmem + 256 + N, da, db2, idesc, 1u);
        }
      }
      umma_commit(&bar);
    }
    __syncwarp();
    long long t1 = clock64();
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
      cycles[0] = t1 - t0;
      cycles[1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

extern "C" {

int probe_mma(int variant, const void* A, const void* B1, const void* B2, float* out1, float* out2) {
  cudaError_t e;
#define RUN(K1, N1, N2)                                                                                \
  {                                                                                                    \
    size_t smem = (128 * K1 + N1 * K1 + N2 * N1) * 2 + 1024;                                           \
    cudaFuncSetAttribute(probe_mma_kernel<K1, N1, N2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    probe_mma_kernel<K1, N1, N2><<<1, 128, smem>>>((const __nv_bfloat16*)A, (const __nv_bfloat16*)B1,   \
                                                    (const __nv_bfloat16*)B2, out1, out2);              \
  }
  if (variant == 0) RUN(32, 16, 16)
  else if (variant == 1) RUN(64, 128, 64)
  else if (variant == 2) RUN(128, 64, 128)
  else return -1;
#undef RUN
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    fprintf(stderr, "probe_mma: %s\n", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

int probe_tma(const void* X, long long rows, const void* Q, float* out, int tile) {
  PFN_encodeTiled encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres);
  if (e != cudaSuccess || !encode) {
    fprintf(stderr, "probe_tma: no cuTensorMapEncodeTiled\n");
    return 2;
  }
  CUtensorMap tmap;
  cuuint64_t gdim[2] = {256, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {256 * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(X), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "probe_tma: encode failed %d\n", (int)r);
    return 3;
  }
  size_t smem = 4 * 16384 + 1024 + 1024;
  cudaFuncSetAttribute(probe_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_tma_kernel<<<1, 128, smem>>>(tmap, (const __nv_bfloat16*)Q, out, tile);
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    fprintf(stderr, "probe_tma: %s\n", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

// returns cycles for `iters` iterations of 4 independent ops per thread, with `threads` threads x `blocks` blocks
int probe_mufu(int which, int iters, int threads, int blocks, float* sink, long long* cycles) {
  probe_mufu_kernel<<<blocks, threads>>>(which, iters, 0.25f, sink, cycles);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    fprintf(stderr, "probe_mufu: %s\n", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

// instructions per iteration per thread: 8 (which 0-5, 9), 14 (6), 8 (7), 10 (8)
int probe_pipe(int which, int iters, int threads, int blocks, float* sink, long long* cycles) {
  probe_pipe_kernel<<<blocks, threads>>>(which, iters, 0.25f, sink, cycles);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    fprintf(stderr, "probe_pipe: %s\n", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

int probe_tmem(int which, int iters, int threads, int blocks, float* sink, long long* cycles) {
  probe_tmem_kernel<<<blocks, threads>>>(which, iters, sink, cycles);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    fprintf(stderr, "probe_tmem: %s\n", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

int probe_mma_rate(int n, int mode, int iters, int blocks, long long* cycles) {
  size_t smem = 16384 + 256 * 32 + 2048;
#define RUNM(NN, MM)                                                                                             \
  {                                                                                                              \
    cudaFuncSetAttribute(probe_mma_rate_kernel<NN, MM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    probe_mma_rate_kernel<NN, MM><<<blocks, 128, smem>>>(iters, cycles);                                         \
  }
#define RUNR(NN)                                                                                                  \
  {                                                                                                               \
    if (mode == 0) RUNM(NN, 0) else if (mode == 1) RUNM(NN, 1) else if (mode == 2) RUNM(NN, 2)                    \
    else if (mode == 3) RUNM(NN, 3) else RUNM(NN, 4)                                                              \
  }
  if (n == 16) RUNR(16) else if (n == 32) RUNR(32) else if (n == 64) RUNR(64) else if (n == 128) RUNR(128) else if (n == 256) RUNR(256) else return -1;
#undef RUNR
#undef RUNM
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    fprintf(stderr, "probe_mma_rate: %s\n", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}
}
