// Quantised attention core: integer QK^T, fp32 softmax, re-quantised P, integer PV.
//   S_int[i,j] = sum_d (q[i,d]-zq)(k[j,d]-zk)         exact int32 (mma.sync m16n8k32 u8/s8)
//   P = softmax_j(S_int * sim_scale)                    fp32, two passes because the reference
//                                                       quantises P AFTER normalisation with a fixed
//                                                       calibrated step (qdiff/quant_block.py:217)
//   Pq = min(rne(P / delta_w), p_qmax)                  8- or 16-bit codes (zero point 0)
//   O[i,:] = out_scale * (sum_j Pq[i,j] v[j,:] - zv * sum_j Pq[i,j])
// 16-bit codes are contracted exactly as two byte planes (hi, lo) against the 8-bit V codes.
// Round-1 implementation uses the legacy mma.sync tensor path with register-resident P
// (FlashAttention-2 style); the key permutation below lets the S accumulator fragment feed the
// PV A-operand without shuffles: inside each 32-key chunk, MMA k-slot (4t+e) <-> key 8*(e>>1)+2t+(e&1).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/qdiff_b200.h"

namespace qd {

template <bool A_SIGNED, bool B_SIGNED>
__device__ __forceinline__ void mma_i8_16832(int (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  if constexpr (A_SIGNED && B_SIGNED) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  } else if constexpr (A_SIGNED && !B_SIGNED) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  } else if constexpr (!A_SIGNED && B_SIGNED) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  } else {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
}

constexpr int ATT_WARPS = 4;
constexpr int ATT_BM = 16 * ATT_WARPS;  // query rows per CTA
constexpr int ATT_BN = 64;              // keys per tile

template <int DQ, bool SIGNED>
__device__ __forceinline__ int bytesum(uint32_t w) {
  if constexpr (SIGNED) return __dp4a((int)w, 0x01010101, 0);
  else return (int)__dp4a(w, 0x01010101u, 0u);
}

// DQ: reduction length of QK^T padded to a multiple of 32; DV: head dim (multiple of 8).
template <int DQ, int DV, bool QK_SIGNED, bool V_SIGNED, bool SM16>
__global__ void __launch_bounds__(ATT_WARPS * 32)
qattention_kernel(const qd_attention_desc p) {
  constexpr int KP = DQ + 16;       // K tile row pitch (bytes)
  constexpr int VP = ATT_BN + 16;   // V^T tile row pitch (bytes)
  constexpr int NKC = DQ / 32;      // k-chunks for QK^T
  constexpr int NDT = DV / 8;       // n8 tiles of the output
  __shared__ __align__(16) uint8_t sK[ATT_BN * KP];
  __shared__ __align__(16) uint8_t sV[DV * VP];
  __shared__ int sRsk[ATT_BN];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int bh = blockIdx.y;
  const int b = bh / p.heads, h = bh - b * p.heads;
  const int row0 = blockIdx.x * ATT_BM + warp * 16;

  const uint8_t* qbase = reinterpret_cast<const uint8_t*>(p.q) + (long long)b * p.Tq * p.ld_q + p.q_off +
                         h * p.head_stride_q;
  const uint8_t* kbase = reinterpret_cast<const uint8_t*>(p.k) + (long long)b * p.Tk * p.ld_k + p.k_off +
                         h * p.head_stride_k;
  const uint8_t* vbase = reinterpret_cast<const uint8_t*>(p.vt) + (long long)b * p.v_batch_stride +
                         (long long)(p.v_off + h * p.head_stride_v) * p.ld_vt;

  // ---- Q fragments (rows g, g+8 of this warp's 16-row slab), zero-padded beyond d
  uint32_t qf[NKC][4];
  int rsq0 = 0, rsq1 = 0;  // row sums of q codes (for the zk correction)
  {
    const int r0 = min(row0 + g, p.Tq - 1), r1 = min(row0 + g + 8, p.Tq - 1);
    const uint8_t* q0 = qbase + (long long)r0 * p.ld_q;
    const uint8_t* q1 = qbase + (long long)r1 * p.ld_q;
#pragma unroll
    for (int kc = 0; kc < NKC; ++kc) {
      const int c0 = kc * 32 + 4 * t, c1 = c0 + 16;
      qf[kc][0] = c0 < p.d ? *reinterpret_cast<const uint32_t*>(q0 + c0) : 0u;
      qf[kc][1] = c0 < p.d ? *reinterpret_cast<const uint32_t*>(q1 + c0) : 0u;
      qf[kc][2] = c1 < p.d ? *reinterpret_cast<const uint32_t*>(q0 + c1) : 0u;
      qf[kc][3] = c1 < p.d ? *reinterpret_cast<const uint32_t*>(q1 + c1) : 0u;
      rsq0 += bytesum<DQ, QK_SIGNED>(qf[kc][0]) + bytesum<DQ, QK_SIGNED>(qf[kc][2]);
      rsq1 += bytesum<DQ, QK_SIGNED>(qf[kc][1]) + bytesum<DQ, QK_SIGNED>(qf[kc][3]);
    }
    rsq0 += __shfl_xor_sync(0xffffffffu, rsq0, 1);
    rsq0 += __shfl_xor_sync(0xffffffffu, rsq0, 2);
    rsq1 += __shfl_xor_sync(0xffffffffu, rsq1, 1);
    rsq1 += __shfl_xor_sync(0xffffffffu, rsq1, 2);
  }
  const int cc0 = p.d * p.zq * p.zk - p.zk * rsq0;  // constant part of the zero-point correction, row g
  const int cc1 = p.d * p.zq * p.zk - p.zk * rsq1;  // row g+8

  const int ntiles = (p.Tk + ATT_BN - 1) / ATT_BN;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  float inv_l0 = 0.f, inv_l1 = 0.f;
  const float inv_dw = 1.0f / p.delta_w;
  const float pmax = (float)p.p_qmax;

  float of[NDT][4];
#pragma unroll
  for (int i = 0; i < NDT; ++i) { of[i][0] = of[i][1] = of[i][2] = of[i][3] = 0.f; }
  int rsp0 = 0, rsp1 = 0;  // row sums of P codes (for the zv correction)

  for (int pass = 0; pass < 2; ++pass) {
    for (int tile = 0; tile < ntiles; ++tile) {
      const int j0 = tile * ATT_BN;
      __syncthreads();  // previous tile fully consumed
      // ---- K tile -> smem (zero-padded in d and beyond Tk)
      for (int idx = threadIdx.x; idx < ATT_BN * (DQ / 4); idx += blockDim.x) {
        const int r = idx / (DQ / 4), w = idx - r * (DQ / 4);
        uint32_t v = 0u;
        if (j0 + r < p.Tk && 4 * w < p.d) v = *reinterpret_cast<const uint32_t*>(kbase + (long long)(j0 + r) * p.ld_k + 4 * w);
        *reinterpret_cast<uint32_t*>(sK + r * KP + 4 * w) = v;
      }
      if (pass == 1) {
        // ---- V^T tile -> smem
        for (int idx = threadIdx.x; idx < DV * (ATT_BN / 16); idx += blockDim.x) {
          const int r = idx / (ATT_BN / 16), w = idx - r * (ATT_BN / 16);
          uint4 v = make_uint4(0u, 0u, 0u, 0u);
          if (j0 + 16 * w < p.Tk) v = *reinterpret_cast<const uint4*>(vbase + (long long)r * p.ld_vt + j0 + 16 * w);
          *reinterpret_cast<uint4*>(sV + r * VP + 16 * w) = v;
        }
      }
      __syncthreads();
      if (p.zq != 0) {
        if (threadIdx.x < ATT_BN) {
          int s = 0;
          for (int w = 0; w < DQ / 4; ++w) s += bytesum<DQ, QK_SIGNED>(*reinterpret_cast<const uint32_t*>(sK + threadIdx.x * KP + 4 * w));
          sRsk[threadIdx.x] = s;
        }
        __syncthreads();
      }

      // ---- S = Q K^T for this warp: 16 x 64
      int sacc[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        sacc[nt][0] = sacc[nt][1] = sacc[nt][2] = sacc[nt][3] = 0;
#pragma unroll
        for (int kc = 0; kc < NKC; ++kc) {
          uint32_t bf[2];
          const uint8_t* kr = sK + (8 * nt + g) * KP + kc * 32 + 4 * t;
          bf[0] = *reinterpret_cast<const uint32_t*>(kr);
          bf[1] = *reinterpret_cast<const uint32_t*>(kr + 16);
          mma_i8_16832<QK_SIGNED, QK_SIGNED>(sacc[nt], qf[kc], bf);
        }
      }
      // ---- scores in fp32
      float sf[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = 8 * nt + 2 * t + (e & 1);
          int v = sacc[nt][e] + ((e < 2) ? cc0 : cc1);
          if (p.zq != 0) v -= p.zq * sRsk[j];
          sf[nt][e] = (j0 + j < p.Tk) ? (float)v * p.sim_scale : -INFINITY;
        }
      }
      if (pass == 0) {
        float tm0 = -INFINITY, tm1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          tm0 = fmaxf(tm0, fmaxf(sf[nt][0], sf[nt][1]));
          tm1 = fmaxf(tm1, fmaxf(sf[nt][2], sf[nt][3]));
        }
        tm0 = fmaxf(tm0, __shfl_xor_sync(0xffffffffu, tm0, 1));
        tm0 = fmaxf(tm0, __shfl_xor_sync(0xffffffffu, tm0, 2));
        tm1 = fmaxf(tm1, __shfl_xor_sync(0xffffffffu, tm1, 1));
        tm1 = fmaxf(tm1, __shfl_xor_sync(0xffffffffu, tm1, 2));
        const float mn0 = fmaxf(m0, tm0), mn1 = fmaxf(m1, tm1);
        l0 *= expf(m0 - mn0);
        l1 *= expf(m1 - mn1);
        m0 = mn0; m1 = mn1;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          l0 += expf(sf[nt][0] - m0) + expf(sf[nt][1] - m0);
          l1 += expf(sf[nt][2] - m1) + expf(sf[nt][3] - m1);
        }
      } else {
        // ---- P codes, packed straight into PV A-fragments (byte planes)
        uint32_t plo[2][4], phi[2][4];
#pragma unroll
        for (int kc = 0; kc < 2; ++kc) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {      // a0/a1 (keys 0..15 of chunk) then a2/a3 (16..31)
            const int ntA = 4 * kc + 2 * half, ntB = ntA + 1;
            uint32_t c[8];
            const float pv[8] = {sf[ntA][0], sf[ntA][1], sf[ntB][0], sf[ntB][1],
                                 sf[ntA][2], sf[ntA][3], sf[ntB][2], sf[ntB][3]};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float pr = expf(pv[e] - (e < 4 ? m0 : m1)) * (e < 4 ? inv_l0 : inv_l1);
              c[e] = (uint32_t)(int)fminf(rintf(pr * inv_dw), pmax);
            }
            rsp0 += (int)(c[0] + c[1] + c[2] + c[3]);
            rsp1 += (int)(c[4] + c[5] + c[6] + c[7]);
            plo[kc][2 * half] = (c[0] & 0xFF) | ((c[1] & 0xFF) << 8) | ((c[2] & 0xFF) << 16) | ((c[3] & 0xFF) << 24);
            plo[kc][2 * half + 1] = (c[4] & 0xFF) | ((c[5] & 0xFF) << 8) | ((c[6] & 0xFF) << 16) | ((c[7] & 0xFF) << 24);
            if constexpr (SM16) {
              phi[kc][2 * half] = ((c[0] >> 8) & 0xFF) | (((c[1] >> 8) & 0xFF) << 8) | (((c[2] >> 8) & 0xFF) << 16) | (((c[3] >> 8) & 0xFF) << 24);
              phi[kc][2 * half + 1] = ((c[4] >> 8) & 0xFF) | (((c[5] >> 8) & 0xFF) << 8) | (((c[6] >> 8) & 0xFF) << 16) | (((c[7] >> 8) & 0xFF) << 24);
            }
          }
        }
        // ---- O += P V
#pragma unroll
        for (int nd = 0; nd < NDT; ++nd) {
          int alo[4] = {0, 0, 0, 0}, ahi[4] = {0, 0, 0, 0};
#pragma unroll
          for (int kc = 0; kc < 2; ++kc) {
            const uint8_t* vr = sV + (8 * nd + g) * VP + 32 * kc + 2 * t;
            uint32_t bf[2];
            bf[0] = (uint32_t)*reinterpret_cast<const uint16_t*>(vr) | ((uint32_t)*reinterpret_cast<const uint16_t*>(vr + 8) << 16);
            bf[1] = (uint32_t)*reinterpret_cast<const uint16_t*>(vr + 16) | ((uint32_t)*reinterpret_cast<const uint16_t*>(vr + 24) << 16);
            mma_i8_16832<false, V_SIGNED>(alo, plo[kc], bf);
            if constexpr (SM16) mma_i8_16832<false, V_SIGNED>(ahi, phi[kc], bf);
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float v = (float)alo[e];
            if constexpr (SM16) v += 256.0f * (float)ahi[e];
            of[nd][e] += v;
          }
        }
      }
    }
    if (pass == 0) {
      l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
      l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
      inv_l0 = 1.0f / l0;
      inv_l1 = 1.0f / l1;
    }
  }
  rsp0 += __shfl_xor_sync(0xffffffffu, rsp0, 1);
  rsp0 += __shfl_xor_sync(0xffffffffu, rsp0, 2);
  rsp1 += __shfl_xor_sync(0xffffffffu, rsp1, 1);
  rsp1 += __shfl_xor_sync(0xffffffffu, rsp1, 2);

  // ---- write O
  const int r0 = row0 + g, r1 = row0 + g + 8;
  const float z0 = (float)p.zv * (float)rsp0, z1 = (float)p.zv * (float)rsp1;
#pragma unroll
  for (int nd = 0; nd < NDT; ++nd) {
    const int col = h * p.d + 8 * nd + 2 * t;
    if (8 * nd + 2 * t < p.d) {
      if (r0 < p.Tq) {
        float2 o = make_float2((of[nd][0] - z0) * p.out_scale, (of[nd][1] - z0) * p.out_scale);
        *reinterpret_cast<float2*>(p.out + ((long long)b * p.Tq + r0) * p.ld_out + col) = o;
      }
      if (r1 < p.Tq) {
        float2 o = make_float2((of[nd][2] - z1) * p.out_scale, (of[nd][3] - z1) * p.out_scale);
        *reinterpret_cast<float2*>(p.out + ((long long)b * p.Tq + r1) * p.ld_out + col) = o;
      }
    }
  }
}

}  // namespace qd
