// EXPERIMENTAL PROBE - not part of libqdiff_b200.so, not validated on hardware in round 1 (written after the GPU budget
// was spent).  A minimal INT8 GEMM on a CTA PAIR (tcgen05.mma.cta_group::2, 256 x BN tile per
// pair, each CTA loads its own 128 rows of A and HALF of the B tile) to measure how far operand sharing lifts the
// L2->SM-bound K >= 2880 layers before the change goes into gemm_i8.cuh.  Built and run by tools/probe_2cta.py.
//
//   out[m, n] = sum_k a[m, k] * b[n, k]        a: u8 [M, K], b: s8 [N, K], out: int32 [M, N]
//   M % 256 == 0, K % 128 == 0, N % BN == 0, BN % 32 == 0, BN <= 256
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <stdint.h>
#include <stdio.h>

#include "../ptx.cuh"

namespace probe {
using namespace qd;

constexpr int BM = 128, BK = 128, STAGES = 4, THREADS = 256;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address -> leader CTA

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// both CTAs execute it; the transaction bytes are credited to the LEADER's mbarrier
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_i8_2sm(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
// arrive (once all prior MMAs of this thread are done) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}
// plain arrive on the LEADER's copy of a barrier (from either CTA)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_MASK) : "memory");
}

struct Args {
  int M, N, K, BN;
  int32_t* out;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm_i8_2cta_probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Args p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const int HB = p.BN / 2;                                   // B rows held by each CTA
  const int stage_bytes = BM * BK + HB * BK;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * stage_bytes);
  uint64_t* full_bar = bars;                 // [STAGES]  used in the leader only
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]  one copy per CTA (multicast commit arrives on both)
  uint64_t* tmem_full = bars + 2 * STAGES;   // [2]       one copy per CTA
  uint64_t* tmem_empty = tmem_full + 2;      // [2]       leader's copy counts the epilogue warps of BOTH CTAs
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int tiles_m2 = p.M / (2 * BM), tiles_n = p.N / p.BN;
  const int num_tiles = tiles_m2 * tiles_n;
  const int num_kb = p.K / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 8);            // 4 epilogue warps x 2 CTAs (only the leader's copy is waited on)
    }
    fence_mbar_init();
    fence_proxy_async();
  }
  if (warp == 2) tmem_alloc_2sm(tmem_ptr, 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += npairs) {
        const int tm = tile / tiles_n, tn = tile - tm * tiles_n;
        const int m0 = (2 * tm + (int)rank) * BM;              // this CTA's 128 rows of the 256-row tile
        const int n0 = tn * p.BN + (int)rank * HB;             // this CTA's half of the B tile
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2u * (uint32_t)stage_bytes);   // bytes of both CTAs
          tma_load_2d_2sm(sa, &tmA, &full_bar[stage], kb * BK, m0);
          tma_load_2d_2sm(sa + BM * BK, &tmB, &full_bar[stage], kb * BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA, one thread) =====================
    if (leader && lane == 0) {
      const uint32_t idesc = make_idesc_i8(2 * BM, p.BN, /*a_signed=*/0, /*b_signed=*/1);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += npairs) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t da = make_smem_desc_sw128(sa);
          const uint64_t db = make_smem_desc_sw128(sa + BM * BK);
          for (int j = 0; j < 4; ++j)
            umma_i8_2sm(d_tmem, da + (uint64_t)(2 * j), db + (uint64_t)(2 * j), idesc, (kb | j) ? 1u : 0u);
          umma_commit_2sm(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs): raw int32 accumulators, one row per thread =====================
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < num_tiles; tile += npairs) {
      const int tm = tile / tiles_n, tn = tile - tm * tiles_n;
      const int m = (2 * tm + (int)rank) * BM + q * 32 + lane;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256);
      for (int c = 0; c < p.BN; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(t_row + (uint32_t)c, v);
        tmem_ld_wait();
        int4* o = reinterpret_cast<int4*>(p.out + (long long)m * p.N + tn * p.BN + c);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = make_int4((int)v[4 * j], (int)v[4 * j + 1], (int)v[4 * j + 2], (int)v[4 * j + 3]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  if (warp == 2) tmem_dealloc_2sm(tmem_base, 512);
}

}  // namespace probe

static int encode2d(CUtensorMap* m, const void* base, uint64_t inner, uint64_t rows, uint32_t box_rows) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return -1;
  auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  cuuint64_t dims[2] = {inner, rows}, strides[1] = {inner};
  cuuint32_t box[2] = {128, box_rows}, estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : -2;
}

// returns 0 or a negative code; *ms = mean kernel time over `iters` launches (CUDA events on the default stream)
extern "C" int probe_gemm_i8_2cta(const void* a, const void* b, int32_t* out, int M, int N, int K, int BN, int iters,
                                  float* ms) {
  if (M % 256 || K % 128 || BN % 32 || BN > 256 || BN < 32 || N % BN) return -10;
  CUtensorMap tmA, tmB;
  if (encode2d(&tmA, a, (uint64_t)K, (uint64_t)M, 128)) return -11;
  if (encode2d(&tmB, b, (uint64_t)K, (uint64_t)N, (uint32_t)(BN / 2))) return -12;
  const int smem = probe::STAGES * (128 * 128 + (BN / 2) * 128) + 256 + 1024;
  if (cudaFuncSetAttribute(probe::gemm_i8_2cta_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -13;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int tiles = (M / 256) * (N / BN);
  int pairs = sms / 2 < tiles ? sms / 2 : tiles;
  probe::Args args{M, N, K, BN, out};
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int it = 0; it < iters + 1; ++it) {
    if (it == 1) cudaEventRecord(e0);
    probe::gemm_i8_2cta_probe<<<2 * pairs, probe::THREADS, smem>>>(tmA, tmB, args);   // __cluster_dims__(2,1,1)
  }
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { fprintf(stderr, "probe: %s\n", cudaGetErrorString(err)); return -20; }
  float t = 0.f;
  if (iters > 0) cudaEventElapsedTime(&t, e0, e1);     // e0 is only recorded when there is a timed launch
  if (ms) *ms = iters > 0 ? t / iters : 0.f;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return 0;
}
