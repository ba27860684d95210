// EXPERIMENTAL micro-probes behind the attention kernel's design (not part of the library):
//   1. tcgen05.ld throughput per SM for 4 / 8 / 16 reader warps (32x32b.x32), one and two loads in flight per warp
//   2. the same loop with 32 MUFU.EX2 per loaded chunk (the pass-1 instruction mix)
//   3. mbarrier hand-off latency between two warps (arrive -> try_wait succeeds)
//   4. tcgen05.mma round trip: issue 2 x (128x128x32 i8) + commit -> mbarrier completes
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tmem_probe tmem_probe.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../ptx.cuh"

using namespace qd;

__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int ILP, bool EXP>
__global__ void __launch_bounds__(544) ld_probe(int iters, long long* cycles, float* sink) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&tptr, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tptr;
  const int nw = (blockDim.x >> 5) - 1;
  float acc = 0.f;
  long long t0 = 0, t1 = 0;
  if (warp >= 1) {
    const int w = warp - 1;
    const uint32_t lane_q = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t col = (uint32_t)((w >> 2) * 32 * ILP) & 511u;
    asm volatile("bar.sync 1, %0;" ::"r"(nw * 32) : "memory");
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      uint32_t v[ILP][32];
#pragma unroll
      for (int k = 0; k < ILP; ++k) tmem_ld_32x32(base + lane_q + ((col + 32 * k) & 511u), v[k]);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < ILP; ++k)
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (EXP) acc += ex2a(__uint_as_float(v[k][j] & 0x3fffffffu) * -1e-30f);
          else acc += __uint_as_float(v[k][j] & 0x007fffffu);
        }
    }
    t1 = clock64();
    asm volatile("bar.sync 1, %0;" ::"r"(nw * 32) : "memory");
  }
  if (threadIdx.x == 32) cycles[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(base, 512);
}

// warp 1 arrives at time ta, warp 2 spins on try_wait and stamps tb; repeated, average of tb - ta.
__global__ void mbar_probe(int iters, long long* out) {
  __shared__ uint64_t bar_ab, bar_ba;
  __shared__ long long stamp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar_ab, 1); mbar_init(&bar_ba, 1); fence_mbar_init(); }
  __syncthreads();
  long long sum = 0;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < iters; ++i) {
      stamp = clock64();
      __threadfence_block();
      mbar_arrive(&bar_ab);
      mbar_wait(&bar_ba, i & 1);
    }
  } else if (warp == 1) {
    for (int i = 0; i < iters; ++i) {
      mbar_wait(&bar_ab, i & 1);
      const long long t = clock64();
      __syncwarp();
      if (lane == 0) { sum += t - stamp; mbar_arrive(&bar_ba); }
    }
    if (lane == 0) out[0] = sum;
  }
}

// one thread: issue `nk` UMMAs (128 x N x 32, i8) + commit, wait for the barrier; average cycles per round trip
__global__ void mma_probe(int iters, int nk, int N, long long* out) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const uint32_t a = smem_u32(raw);
  uint8_t* smem = raw + (((a + 1023u) & ~1023u) - a);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 2 * 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tptr, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 32) {
    const uint32_t idesc = make_idesc_i8(128, N, 0, 0);
    const uint64_t da = make_smem_desc_sw128(smem_u32(smem)), db = make_smem_desc_sw128(smem_u32(smem + 16384));
    long long sum = 0;
    for (int i = 0; i < iters; ++i) {
      const long long t0 = clock64();
      for (int j = 0; j < nk; ++j) umma_i8(tptr, da + 2 * j, db + 2 * j, idesc, j ? 1u : 0u);
      umma_commit(&bar);
      mbar_wait(&bar, i & 1);
      sum += clock64() - t0;
    }
    out[0] = sum;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tptr, 512);
}

// the pass-1 arithmetic of the attention kernel on 16 warps (no MMA, no barriers): per iteration every warp loads 32 columns
// in two 16-column pieces and runs {IADD z, max, FADD2, FFMA2, EX2, FADD2}.  VAR bits: 1 = no max chain, 2 = scalar fp32
// instead of packed, 4 = no z LDS, 8 = no exp2 (multiply instead), 16 = x32 pieces
template <int VAR>
__global__ void __launch_bounds__(576) p1_probe(int iters, long long* cycles, float* sink, float c) {
  __shared__ uint32_t tptr;
  __shared__ __align__(16) int ztab[512];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 512; i += blockDim.x) ztab[i] = 0x4B400000 - (i & 63);
  if (warp == 0) { tmem_alloc(&tptr, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tptr;
  float l = 0.f;
  int mi = 0x4B400000;
  long long t0 = 0, t1 = 0;
  if (warp >= 2) {
    const int part = (warp - 2) >> 2;
    const uint32_t lane_q = (uint32_t)((warp & 3) * 32) << 16;
    asm volatile("bar.sync 1, %0;" ::"r"(512) : "memory");
    t0 = clock64();
    constexpr int W = (VAR & 16) ? 32 : 16;
    for (int i = 0; i < iters; ++i) {
#pragma unroll 1
      for (int pc = 0; pc < 32 / W; ++pc) {
        uint32_t v[W];
        const int col0 = part * 32 + pc * W;
        if constexpr (W == 16) tmem_ld_32x16(base + lane_q + ((i & 1) * 128) + col0, *reinterpret_cast<uint32_t(*)[16]>(v));
        else tmem_ld_32x32(base + lane_q + ((i & 1) * 128) + col0, *reinterpret_cast<uint32_t(*)[32]>(v));
        tmem_ld_wait();
        int s[W];
        const int* zr = ztab + ((i * 128 + col0) & 511);
#pragma unroll
        for (int j = 0; j < W; j += 4) {
          int4 z = make_int4(0x4B400000, 0x4B400000, 0x4B400000, 0x4B400000);
          if (!(VAR & 4)) z = *reinterpret_cast<const int4*>(zr + (j & 12));
          s[j] = (int)(v[j] & 0xffff) + z.x; s[j + 1] = (int)(v[j + 1] & 0xffff) + z.y;
          s[j + 2] = (int)(v[j + 2] & 0xffff) + z.z; s[j + 3] = (int)(v[j + 3] & 0xffff) + z.w;
        }
        const float b0 = -(float)(mi - 0x4B400000) * c;
        float acc;
        if (VAR & 2) {
          acc = 0.f;
#pragma unroll
          for (int j = 0; j < W; ++j) {
            const float x = fmaf(__int_as_float(s[j]) - 12582912.0f, c, b0);
            acc += (VAR & 8) ? x * 0.5f : ex2a(x);
          }
        } else {
          const float2 c2 = make_float2(c, c), b2 = make_float2(b0, b0);
          float2 acc2 = make_float2(0.f, 0.f);
#pragma unroll
          for (int j = 0; j < W; j += 2) {
            const float2 x = ffma2(fadd2(make_float2(__int_as_float(s[j]), __int_as_float(s[j + 1])), make_float2(-12582912.0f, -12582912.0f)), c2, b2);
            acc2 = fadd2(acc2, make_float2((VAR & 8) ? x.x * 0.5f : ex2a(x.x), (VAR & 8) ? x.y * 0.5f : ex2a(x.y)));
          }
          acc = acc2.x + acc2.y;
        }
        if (!(VAR & 1)) {
          int tm = s[0];
#pragma unroll
          for (int j = 1; j < W; ++j) tm = max(tm, s[j]);
          l += acc;
          if (tm > mi) { l *= ex2a((float)(mi - tm) * c); mi = tm; }
        } else {
          l += acc;
        }
      }
    }
    t1 = clock64();
  }
  if (threadIdx.x == 64) cycles[blockIdx.x] = t1 - t0;
  if (l == 123.456f) sink[0] = l + mi;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(base, 512);
  (void)lane;
}

template <int VAR>
void run_p1(const char* name) {
  const int iters = 2000, blocks = 148;
  long long* cyc;
  float* sink;
  cudaMalloc(&cyc, blocks * sizeof(long long));
  cudaMalloc(&sink, 4);
  p1_probe<VAR><<<blocks, 576>>>(iters, cyc, sink, 1e-4f);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0;
  for (int i = 0; i < blocks; ++i) c += (double)h[i];
  c /= blocks;
  printf("pass-1 arithmetic, 16 warps, %-44s: %7.1f cycles per 128x128 tile (%s)\n", name, c / iters, cudaGetErrorString(e));
  cudaFree(cyc);
  cudaFree(sink);
}

template <int ILP, bool EXP>
void run_ld(int nw, const char* name) {
  const int iters = 4000, blocks = 148;
  long long* cyc;
  float* sink;
  cudaMalloc(&cyc, blocks * sizeof(long long));
  cudaMalloc(&sink, 4);
  ld_probe<ILP, EXP><<<blocks, 32 * (nw + 1)>>>(iters, cyc, sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0;
  for (int i = 0; i < blocks; ++i) c += (double)h[i];
  c /= blocks;
  const double bytes = (double)iters * ILP * nw * 4096.0;
  printf("%-28s warps=%2d ILP=%d: %8.1f cycles / iteration, %6.1f B/clk/SM, %5.2f elem/clk/SM  (%s)\n", name, nw, ILP,
         c / iters, bytes / c, bytes / 4 / c, cudaGetErrorString(e));
  cudaFree(cyc);
  cudaFree(sink);
}

int main() {
  run_p1<0>("full (x16 pieces)");
  run_p1<16>("full (x32 pieces)");
  run_p1<1>("no max chain");
  run_p1<2>("scalar fp32");
  run_p1<4>("no z LDS");
  run_p1<8>("no exp2");
  run_p1<8 | 2>("no exp2, scalar");
  run_p1<1 | 4>("no max, no z");
  run_p1<1 | 4 | 2>("no max, no z, scalar");
  run_p1<1 | 4 | 8>("no max, no z, no exp2");
  for (int nw : {4, 8, 16}) run_ld<1, false>(nw, "tcgen05.ld x32");
  for (int nw : {4, 8, 16}) run_ld<2, false>(nw, "tcgen05.ld x32");
  for (int nw : {4, 8, 16}) run_ld<1, true>(nw, "tcgen05.ld x32 + 32 ex2");
  for (int nw : {4, 8, 16}) run_ld<2, true>(nw, "tcgen05.ld x32 + 32 ex2");
  long long* out;
  cudaMalloc(&out, 8);
  long long h;
  mbar_probe<<<1, 64>>>(10000, out);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  printf("mbarrier arrive -> waiter wakes: %.1f cycles (%s)\n", (double)h / 10000, cudaGetErrorString(cudaGetLastError()));
  cudaFuncSetAttribute(mma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  for (int N : {64, 128}) {
    for (int nk : {1, 2, 4, 8}) {
      mma_probe<<<1, 64, 34 * 1024>>>(2000, nk, N, out);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
      printf("mma round trip 128x%dx32 x %d + commit + wait: %.1f cycles (%s)\n", N, nk, (double)h / 2000, cudaGetErrorString(e));
    }
  }
  return 0;
}
