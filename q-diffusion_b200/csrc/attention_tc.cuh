// Quantised attention on the 5th-generation tensor cores (tcgen05 + TMEM), head dim d <= 112.
//
// Same arithmetic as attention.cuh (integer QK^T, fp32 softmax in the exp2 domain, P re-quantised after
// normalisation with the calibrated step, integer PV with hi/lo byte planes) but S and O live in TMEM, the
// MMAs are issued by one thread, and each softmax thread owns one query ROW (half of the 128 key columns of
// a tile; now a quarter: 16 softmax warps): no fragment bookkeeping, no quad shuffles, no IMMA / fragment-LDS issue slots -> ~16 scalar
// instructions per score instead of ~24 (profiles/r01_attention_v2.txt).
//
// CTA = 128 query rows of one (batch, head); 10 warps:
//   warp 0      loader : Q / K tiles by TMA from the per-head padded code layout (pitch P = 32/64/128 bytes ==
//                        the swizzle span, written by the to_q / to_k GEMM epilogues), V^T tile (d rows x 128
//                        keys) by TMA, zq*rowsum(k) slice by cp.async; 4-stage ring
//   warp 1      MMA    : S = Q K^T  (M=128, N=128, K=32 x ceil(d/32)) into a double-buffered TMEM slot;
//                        O_lo/O_hi += P_lo/P_hi V^T (M=128, N=NV, K=32 x 4), int32 in TMEM for the whole pass
//   warps 2-17  softmax: thread = (row, 32-column quarter); pass 1: integer row max + sum of exp2; pass 2: codes ->
//                        byte planes written to shared memory as the next MMA's A operand (128B swizzle,
//                        V^T key permutation applied while packing)
// Row sums of the P codes come from an all-ones V^T row (row d of the tile).
#pragma once
#include "attention.cuh"
#include "ptx.cuh"
#include <type_traits>

namespace qd {

// Two configurations of the same kernel (template parameter NSW = softmax warps):
//   NSW = 16: one CTA per SM, 4 softmax warps per TMEM lane quarter (32 key columns of a tile each), S double-buffered in
//             TMEM (512 columns), 4-stage K/V ring.
//   NSW =  8: TWO CTAs per SM (d <= 64): 2 softmax warps per lane quarter (64 key columns each), one S slot + O in 256 TMEM
//             columns, 2-stage K/V ring, <= 102 registers.  The two co-resident CTAs are independent pipelines: one CTA's
//             prologue (TMEM allocation, Q/K loads), barrier round trips and epilogue are covered by the other's
//             arithmetic (profiles/r01_attention_tc_final.txt: 13 % of the warp samples sat in mbarrier spins).
constexpr int ATC_BM = 128, ATC_BN = 128;
#ifndef ATC_PBUFS
#define ATC_PBUFS 2         // P buffers (lo + hi plane each); 3 (when shared memory allows) measured slower: 1541 vs 1461 us
#endif

__host__ __device__ constexpr int atc_threads(int NSW) { return 64 + 32 * NSW; }
__host__ __device__ constexpr int atc_stages(int NSW) { return NSW == 16 ? 4 : 2; }
__host__ __device__ constexpr int atc_sslots(int NSW) { return NSW == 16 ? 2 : 1; }

struct AtcSmem {
  int q_off, k_off, v_off, p_off, zrk_off, stat_off, bar_off, total, v_stage, k_stage;
  int npb;      // P buffers
};
__host__ __device__ inline AtcSmem atc_smem_layout(int NV, int P, int NSW = 16) {
  const int stages = atc_stages(NSW);
  AtcSmem l;
  l.q_off = 0;
  l.k_off = (128 * P + 1023) / 1024 * 1024;  // Q tile: 128 rows x P bytes
  l.k_stage = 128 * P;                       // 128 keys x P bytes (P = swizzle span)
  l.v_stage = NV * 128;
  l.v_off = l.k_off + stages * l.k_stage;
  l.p_off = (l.v_off + stages * l.v_stage + 1023) / 1024 * 1024;
  l.npb = 2;
  if (NSW == 16 && ATC_PBUFS > 2 && l.p_off + ATC_PBUFS * 32768 + stages * 512 + (NSW / 4) * 128 * 8 + 256 + 1024 <= 227 * 1024)
    l.npb = ATC_PBUFS;
  l.zrk_off = l.p_off + l.npb * 2 * 16384;          // [buffer][plane]
  l.stat_off = l.zrk_off + stages * 512;
  l.bar_off = l.stat_off + (NSW / 4) * 128 * 8;
  l.total = l.bar_off + 256 + 1024;
  return l;
}

// K-major smem matrix descriptor for rows of P bytes with the P-byte swizzle (P = 32, 64, 128)
__device__ __forceinline__ uint64_t make_smem_desc_sw(uint32_t smem_addr, int P) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8 * P) >> 4) << 32;                       // stride between 8-row groups
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(P == 128 ? 2 : (P == 64 ? 4 : 6)) << 61;  // SWIZZLE_128B / 64B / 32B
  return d;
}

// What sets the pace (csrc/experimental/tmem_probe.cu, tools/prof_attn.py; B=16 x 8 heads, T=4096, d=40):
//   * the softmax warps are INSTRUCTION bound, not XU bound: replacing every exp2 by a multiply changes 1765 us to 1667 us, and
//     the bare pass-1 arithmetic on 16 warps without any MMA or barrier needs 1274 cycles per 128x128 tile against 1024 for
//     the exp2 alone.  Every instruction per score counts: HZ as a template parameter removes 32 predicated MOVs per chunk
//     (-5 %), the int->float conversion is folded into the scaling FMA (below, -0.5 instruction per score and pass).
//   * the S slot goes back to the MMA warp as soon as a warp holds its scores in registers (not after the exp2 work), and
//     in pass 2 the MMA warp issues S(t + 2) BEFORE P(t) V(t) - both become possible at about the same time and the short S
//     MMA must not queue behind 8 PV MMAs on the in-order tensor pipe.  Together -7.5 % (either alone: -1.4 % / -2.7 %).
//   * tried and measured slower or equal (profiles/r02_attention_variants.txt): two softmax groups on alternate tiles
//     (+9 %), phase-shifting the four warps of a scheduler with nanosleep (0), four parallel max chains / two exp2
//     accumulators (0), software pipelines that load the next 16- or 32-column piece during the arithmetic of the current
//     one, with the exp2 stream referenced to the previous pieces' maximum (integer path +13 %, fp16 path +7 % / +14 %),
//     8-stage K ring + whole-axis zq*rowsum(k) table (0: the loader is not on the critical path), a third P buffer (+5 %),
//     a third or a quarter of the exp2 as a degree-6 polynomial on the FMA pipe (+1..+12 %: with four warps per scheduler
//     the softmax is bound by each warp's serial instruction stream, so every added instruction costs time even on an idle
//     pipe, while removing the exp2 altogether only gains 13 %).
// HZ: q has a zero point (scores need the zq*rowsum(k) correction).
// F16: Q and K arrive as fp16 values (code - zero_point), written that way by the to_q / to_k GEMM epilogues
// (qd_gemm_desc.out_q_f16), and S = Q K^T runs as tcgen05.mma kind::f16 with fp32 accumulation: products and sums of
// integers below 2^24 are exact in fp32, so S is the SAME integer (q - zq).(k - zk) the integer path produces after its
// zero-point correction - but it needs no correction (no zq*rowsum(k) kernel, table, LDS or IADD per score), no int->float
// conversion, and the row maximum is a 3-input FMNMX.  Per score: pass 1 = {FMNMX3/2, FFMA2/2, EX2, FADD2/2}, pass 2 =
// {FFMA2/2, EX2, FFMA2/2, PRMT, STS/16} - about 5.6 instructions over both passes against 9.8 on the integer path.  The
// tensor pipe runs kind::f16 at half the kind::i8 rate, which does not matter at ~12 % utilisation.  d <= 64 (one 128-byte
// swizzle span per row).
template <bool SM16, bool MAGIC, int NSW, bool HZ, bool F16 = false>
__global__ void __launch_bounds__(atc_threads(NSW), NSW == 16 ? 1 : 2)
qattention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const qd_attention_desc p, const int NV, const int P) {
  extern __shared__ uint8_t atc_raw[];
  const uint32_t raw_addr = smem_u32(atc_raw);
  uint8_t* smem = atc_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  constexpr int ATC_SOFTMAX_WARPS = NSW;
  constexpr int ATC_THREADS = atc_threads(NSW);
  constexpr int ATC_STAGES = atc_stages(NSW);
  constexpr int SSLOTS = atc_sslots(NSW);
  constexpr int NPART = NSW / 4;             // softmax warps per TMEM lane quarter
  constexpr int CPT = 4 / NPART;             // 32-column chunks of a tile per softmax thread
  constexpr int WPG = NSW;                   // softmax warps that work on one tile
  constexpr uint32_t TMEM_COLS = NSW == 16 ? 512 : 256;
  const AtcSmem L = atc_smem_layout(NV, P, NSW);
  uint8_t* sQ = smem + L.q_off;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  uint64_t* kv_full = bars;        // [ATC_STAGES]
  uint64_t* kv_empty = bars + 4;   // [ATC_STAGES]
  uint64_t* s_full = bars + 8;     // [2]
  uint64_t* s_empty = bars + 10;   // [2]
  uint64_t* p_full = bars + 12;    // [4]
  uint64_t* p_empty = bars + 16;   // [4]
  uint64_t* o_done = bars + 20;    // [1]
  uint64_t* q_full = bars + 21;    // [1]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 22);
  const int npb = L.npb;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int b = bh / p.heads, h = bh - b * p.heads;
  const int row_base = blockIdx.x * ATC_BM;
  const int d = p.d;
  const int nks = F16 ? (d + 15) >> 4 : (d + 31) >> 5;   // 32-byte K slices of QK^T (32 codes, or 16 fp16 values)
  const int ntiles = (p.Tk + ATC_BN - 1) / ATC_BN;
  // timing experiments (-DATC_DBG_SKIP1 / -DATC_DBG_SKIP2: one pass only, results wrong)
#ifdef ATC_DBG_SKIP1
  const int n1 = 0;
#else
  const int n1 = ntiles;
#endif
#ifdef ATC_DBG_SKIP2
  const int n2 = 0;
#else
  const int n2 = ntiles;
#endif
  constexpr bool has_zq = HZ;

  const int* zrk_g = reinterpret_cast<const int*>(p.ws) + (long long)bh * (long long)att_ws_stride(p.Tk);

  // ---- one-time setup: V^T stages zeroed (rows beyond d feed the MMA), all-ones row d for the row sums
  for (int i = threadIdx.x; i < (ATC_STAGES * L.v_stage) / 16; i += ATC_THREADS)
    reinterpret_cast<uint4*>(smem + L.v_off)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  for (int i = threadIdx.x; i < ATC_STAGES * 8; i += ATC_THREADS)   // a constant row is invariant under the swizzle
    reinterpret_cast<uint4*>(smem + L.v_off + (i >> 3) * L.v_stage + d * 128)[i & 7] =
        make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < ATC_STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], F16 ? 1 : 1 + WPG);    // MMA commit (+ the softmax warps of the tile: they read the zq*rowsum(k) slice of the stage)
    }
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], WPG);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&p_full[i], WPG);
      mbar_init(&p_empty[i], 1);
    }
    mbar_init(o_done, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    if (lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); }
    tmem_alloc(tmem_ptr, TMEM_COLS);
    tmem_relinquish();
  }
  fence_proxy_async();   // generic-proxy writes of Q / constant rows -> visible to the tensor-core (async) proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tm_s = tmem_base;             // S slots: columns [0,128) (and [128,256) when double-buffered)
  const uint32_t tm_olo = tmem_base + SSLOTS * 128;                          // O_lo: NV columns
  const uint32_t tm_ohi = tmem_base + SSLOTS * 128 + (NSW == 16 ? 128 : 64);  // O_hi (NV <= 64 in the 256-column layout)

  if (warp == 0) {
    // ===================== loader =====================
    int st = 0;
    uint32_t ph = 0;
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, (uint32_t)(128 * P));
      tma_load_2d(sQ, &tmQ, q_full, h * P, b * p.Tq + row_base);
    }
    for (int pass = 0; pass < 2; ++pass) {
      for (int t = 0; t < (pass == 0 ? n1 : n2); ++t) {
        const int j0 = t * ATC_BN;
        mbar_wait(&kv_empty[st], ph ^ 1);
        if (has_zq) {
          cp_async16(smem + L.zrk_off + st * 512 + 16 * lane, zrk_g + j0 + 4 * lane);
          cp_async_commit();
          cp_async_wait_all();
        }
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_expect_tx(&kv_full[st], (uint32_t)(128 * P + (pass == 1 ? d * 128 : 0)));
          tma_load_2d(smem + L.k_off + st * L.k_stage, &tmK, &kv_full[st], h * P, b * p.Tk + j0);
          if (pass == 1) tma_load_2d(smem + L.v_off + st * L.v_stage, &tmV, &kv_full[st], j0, bh * d);
        }
        if (++st == ATC_STAGES) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc_s = F16 ? make_idesc_f16(128, 128) : make_idesc_i8(128, 128, p.q_signed, p.k_signed);
      const uint32_t idesc_o = make_idesc_i8(128, NV, 0, p.v_signed);
      const uint64_t dq = make_smem_desc_sw(smem_u32(sQ), P);
      mbar_wait(q_full, 0);
      int st = 0, sb = 0, pb = 0;
      uint32_t ph_kv = 0, ph_s = 0, ph_p = 0;
      // ---- pass 1: S only
      for (int t = 0; t < n1; ++t) {
        mbar_wait(&kv_full[st], ph_kv);
        mbar_wait(&s_empty[sb], ph_s ^ 1);
        tc_fence_after();
        const uint64_t dk = make_smem_desc_sw(smem_u32(smem + L.k_off + st * L.k_stage), P);
        for (int j = 0; j < nks; ++j) {
          if (F16) umma_bf16(tm_s + sb * 128, dq + 2 * j, dk + 2 * j, idesc_s, j ? 1u : 0u);     // kind::f16, fp16 operands
          else umma_i8(tm_s + sb * 128, dq + 2 * j, dk + 2 * j, idesc_s, j ? 1u : 0u);
        }
        umma_commit(&s_full[sb]);
        umma_commit(&kv_empty[st]);
        if (++st == ATC_STAGES) { st = 0; ph_kv ^= 1; }
        if (++sb == SSLOTS) { sb = 0; ph_s ^= 1; }
      }
      // ---- pass 2: SSLOTS score tiles are kept in flight ahead of the PV MMAs (S(t + SSLOTS) is issued BEFORE P(t) V(t))
      auto issue_s = [&](int st_, int sb_) {
        const uint64_t dk = make_smem_desc_sw(smem_u32(smem + L.k_off + st_ * L.k_stage), P);
        for (int j = 0; j < nks; ++j) {
          if (F16) umma_bf16(tm_s + sb_ * 128, dq + 2 * j, dk + 2 * j, idesc_s, j ? 1u : 0u);
          else umma_i8(tm_s + sb_ * 128, dq + 2 * j, dk + 2 * j, idesc_s, j ? 1u : 0u);
        }
      };
      int st_s = st, sb_s = sb;
      uint32_t ph_kv_s = ph_kv, ph_s_s = ph_s;
      if (n2 > 0) {
      mbar_wait(&kv_full[st_s], ph_kv_s);
      mbar_wait(&s_empty[sb_s], ph_s_s ^ 1);
      tc_fence_after();
      issue_s(st_s, sb_s);
      umma_commit(&s_full[sb_s]);
      int st_v = st;            // K/V stage of the tile whose P V is issued next
      auto next_s = [&](int tnext) {
        if (++st_s == ATC_STAGES) { st_s = 0; ph_kv_s ^= 1; }
        if (++sb_s == SSLOTS) { sb_s = 0; ph_s_s ^= 1; }
        if (tnext < n2) {
          mbar_wait(&kv_full[st_s], ph_kv_s);
          mbar_wait(&s_empty[sb_s], ph_s_s ^ 1);
          tc_fence_after();
          issue_s(st_s, sb_s);
          umma_commit(&s_full[sb_s]);
        }
      };
      if (SSLOTS == 2) next_s(1);
      for (int t = 0; t < n2; ++t) {
        const int st_cur = st_v;
        if (++st_v == ATC_STAGES) st_v = 0;
        next_s(t + SSLOTS);
        mbar_wait(&p_full[pb], ph_p);
        tc_fence_after();
        const uint64_t dv = make_smem_desc_sw128(smem_u32(smem + L.v_off + st_cur * L.v_stage));
        const uint64_t dplo = make_smem_desc_sw128(smem_u32(smem + L.p_off + (pb * 2) * 16384));
        for (int j = 0; j < 4; ++j) umma_i8(tm_olo, dplo + 2 * j, dv + 2 * j, idesc_o, (t | j) ? 1u : 0u);
        if (SM16) {
          const uint64_t dphi = make_smem_desc_sw128(smem_u32(smem + L.p_off + (pb * 2 + 1) * 16384));
          for (int j = 0; j < 4; ++j) umma_i8(tm_ohi, dphi + 2 * j, dv + 2 * j, idesc_o, (t | j) ? 1u : 0u);
        }
        umma_commit(&p_empty[pb]);
        umma_commit(&kv_empty[st_cur]);
        if (++pb == npb) { pb = 0; ph_p ^= 1; }
      }
      }
      umma_commit(o_done);
    }
  } else {
    // ===================== softmax warps =====================
    const int q4 = warp & 3;                 // TMEM lane quarter of this warp
    const int part = (warp - 2) >> 2;        // which 32 key columns of a tile (0..3)
    const int row = q4 * 32 + lane;          // query row inside the CTA tile == TMEM lane
    const uint32_t t_lane = (uint32_t)(q4 * 32) << 16;
    const float c = p.sim_scale * 1.4426950408889634f;
    const float pmax = (float)p.p_qmax;
    float2* stat = reinterpret_cast<float2*>(smem + L.stat_off);
    // Tile counters: T = pass * ntiles + t is the global index of a key tile in MMA issue order.  Everything follows from
    // it: S slot T & 1 (phase (T >> 1) & 1; one slot: phase T & 1), K/V stage T % stages, P buffer t & 1 (phase (t >> 1) & 1).
    auto s_slot = [](int T) { return SSLOTS == 2 ? (T & 1) : 0; };
    auto s_phase = [](int T) -> uint32_t { return SSLOTS == 2 ? ((T >> 1) & 1) : (T & 1); };
    // Scores are handled in BIASED integer form t = S_raw - zq*rowsum(k) + BIAS (one IADD against the staged
    // "BIAS - zq*rowsum" table): with BIAS = 0x4B400000 (d <= 64, |S| < 2^22) the same register is the score for the
    // integer row max AND the bit pattern of the float 1.5*2^23 + S, so the int->float conversion is a single FADD.
    // F16: the TMEM word already IS the float score; `mi` then holds float bits and NONE / MASKED are -inf.
    constexpr int BIAS = MAGIC ? 0x4B400000 : 0;
    constexpr int NONE = F16 ? (int)0xFF800000 : INT_MIN;
    constexpr int MASKED = F16 ? (int)0xFF800000 : (MAGIC ? BIAS - (1 << 22) + 1 : INT_MIN / 2);
    constexpr bool FOLD = MAGIC || F16;      // scores are float bit patterns F: exponent = fma(F, c, hi)
    auto s_gt = [](int a, int b) { return F16 ? __int_as_float(a) > __int_as_float(b) : a > b; };
    auto s_max = [](int a, int b) { return F16 ? __float_as_int(fmaxf(__int_as_float(a), __int_as_float(b))) : max(a, b); };
    auto s_diff = [](int a, int b) { return F16 ? __int_as_float(a) - __int_as_float(b) : (float)(a - b); };   // a - b in score units
#if defined(ATC_DBG_NOXU)
#define ex2_approx(x) ((x) * 0.5f)       /* timing experiment: no XU work (results wrong) */
#endif
    auto tof2 = [](int t0, int t1) -> float2 {
      return MAGIC ? fadd2(make_float2(__int_as_float(t0), __int_as_float(t1)), make_float2(-12582912.0f, -12582912.0f))
                   : make_float2((float)t0, (float)t1);
    };
    int mi = NONE;
    float l = 0.f;
    // ---- pass 1
    for (int t = 0; t < n1; ++t) {
      const int T = t;
      const int sb = s_slot(T), st = T % ATC_STAGES;
      mbar_wait(&s_full[sb], s_phase(T));
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < CPT; ++cc) {
        const int col0 = (part * CPT + cc) * 32;          // this thread's 32 key columns of the tile
        const int j0 = t * ATC_BN + col0;
        const int* zr = reinterpret_cast<const int*>(smem + L.zrk_off + st * 512) + col0;
        uint32_t v[32];
        tmem_ld_32x32(tm_s + t_lane + sb * 128 + col0, v);
        tmem_ld_wait();
        if (cc == CPT - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_empty[sb]);
        }
        int s[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          if (F16) {
            s[j] = (int)v[j]; s[j + 1] = (int)v[j + 1]; s[j + 2] = (int)v[j + 2]; s[j + 3] = (int)v[j + 3];
          } else {
            int4 z = make_int4(BIAS, BIAS, BIAS, BIAS);
            if (has_zq) z = *reinterpret_cast<const int4*>(zr + j);
            s[j] = (int)v[j] + z.x; s[j + 1] = (int)v[j + 1] + z.y; s[j + 2] = (int)v[j + 2] + z.z; s[j + 3] = (int)v[j + 3] + z.w;
          }
        }
        bool any_valid = true;
        if (j0 + 32 > p.Tk) {     // ragged last tile: masked keys drop out of the max and of the sum
          any_valid = j0 < p.Tk;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j0 + j >= p.Tk) s[j] = MASKED;
        }
        if (any_valid) {
          int tm = s[0];
#pragma unroll
          for (int j = 1; j < 32; ++j) tm = s_max(tm, s[j]);
          if (s_gt(tm, mi)) { l *= (mi == NONE) ? 0.f : ex2_approx(s_diff(mi, tm) * c); mi = tm; }
          // packed fp32 (FADD2 / FFMA2): same IEEE results as the scalar form, half the issue slots
          const float2 c2 = make_float2(c, c);
          float2 acc2 = make_float2(0.f, 0.f);
          if (FOLD) {
            // F = float pattern of the biased score = K + S (F16: the score itself, K = 0), Fm = K + max: c * (S - max) = fma(F, c, -c * Fm).  -c * Fm is
            // split exactly into hi + lo (lo = the FMA residual), the chunk is summed against hi alone - one FFMA2 per
            // two scores instead of FADD2 + FFMA2 - and the sum is multiplied by 2^lo afterwards (|lo| <= ulp(c * Fm) / 2)
            const float Fm = __int_as_float(mi);
            const float hi = -c * Fm, lo = fmaf(-c, Fm, -hi);
            const float2 h2 = make_float2(hi, hi);
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float2 x = ffma2(make_float2(__int_as_float(s[j]), __int_as_float(s[j + 1])), c2, h2);
              acc2 = fadd2(acc2, make_float2(ex2_approx(x.x), ex2_approx(x.y)));
            }
            const float tl = lo * 0.6931471805599453f;
            l = fmaf(acc2.x + acc2.y, fmaf(tl, fmaf(tl, 0.5f, 1.0f), 1.0f), l);
          } else {
            const float b0 = -(float)(mi - BIAS) * c;
            const float2 b2 = make_float2(b0, b0);
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float2 x = ffma2(tof2(s[j], s[j + 1]), c2, b2);
              acc2 = fadd2(acc2, make_float2(ex2_approx(x.x), ex2_approx(x.y)));
            }
            l += acc2.x + acc2.y;
          }
        }
      }
      __syncwarp();
      if (!F16 && lane == 0) mbar_arrive(&kv_empty[st]);
    }
    // ---- combine the column parts of every row (named barrier over the softmax warps)
    stat[part * 128 + row] = make_float2(__int_as_float(mi), l);
    asm volatile("bar.sync 1, %0;" ::"n"(32 * NSW) : "memory");
    float off;          // code = exp2(c * S + off)
    float off_hi, g_lo; // MAGIC: c * S + off = fma(F, c, off_hi) + lo with F = K + S; g_lo = 2^lo multiplies the exponential
    bool row_clamps;
    {
      int mm = NONE;
#pragma unroll
      for (int k = 0; k < NPART; ++k) mm = s_max(mm, __float_as_int(stat[k * 128 + row].x));
      float lt = 0.f;
#pragma unroll
      for (int k = 0; k < NPART; ++k) {
        const float2 o = stat[k * 128 + row];
        const int mo = __float_as_int(o.x);
        lt += o.y * ((mo == NONE) ? 0.f : ex2_approx(s_diff(mo, mm) * c));
      }
      const float inv = 1.0f / (lt * p.delta_w);     // the row's largest possible P code (score == row max)
      const float lg = log2f(inv);
      off = (F16 ? -__int_as_float(mm) : -(float)(mm - BIAS)) * c + lg;
      {
        // -c * (K + max) + lg = h1 + l1 + lg exactly (FMA residual), then TwoSum(h1, lg) = off_hi + err
        const float Fm = __int_as_float(mm);
        const float h1 = -c * Fm, l1 = fmaf(-c, Fm, -h1);
        off_hi = h1 + lg;
        const float bb = off_hi - h1;
        const float err = (h1 - (off_hi - bb)) + (lg - bb);
        const float tl = (err + l1) * 0.6931471805599453f;
        g_lo = fmaf(tl, fmaf(tl, 0.5f, 1.0f), 1.0f);
      }
      row_clamps = !(inv <= pmax);
    }
    const bool warp_clamps = __any_sync(0xffffffffu, row_clamps);
    // ---- pass 2
    int pb = 0;
    uint32_t ph_p = 0;
    for (int t = 0; t < n2; ++t) {
      const int T = n1 + t;
      const int sb = s_slot(T), st = T % ATC_STAGES;
      mbar_wait(&s_full[sb], s_phase(T));
      mbar_wait(&p_empty[pb], ph_p ^ 1);
      tc_fence_after();
      uint8_t* pl = smem + L.p_off + (pb * 2) * 16384 + row * 128;
      uint8_t* phh = pl + 16384;
#pragma unroll 1
      for (int cc = 0; cc < CPT; ++cc) {
        const int col0 = (part * CPT + cc) * 32;
        const int j0 = t * ATC_BN + col0;
        const int* zr = reinterpret_cast<const int*>(smem + L.zrk_off + st * 512) + col0;
        uint32_t v[32];
        tmem_ld_32x32(tm_s + t_lane + sb * 128 + col0, v);
        tmem_ld_wait();
        if (cc == CPT - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_empty[sb]);
        }
        uint32_t cd[32];
        const float2 c2 = make_float2(c, c), off2 = make_float2(off, off), k2 = make_float2(12582912.0f, 12582912.0f);
        const float2 oh2 = make_float2(off_hi, off_hi), g2 = make_float2(g_lo, g_lo);
        // code = rne(min(exp2(c*s + off), pmax)) through the 1.5*2^23 constant; the min is skipped (warp-uniformly) when no
        // row of this warp can exceed pmax: its largest code is 1 / (l * delta_w), known after pass 1
        auto codes = [&](auto clamp_tag) {
          constexpr bool CLAMP = decltype(clamp_tag)::value;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            int4 z = make_int4(BIAS, BIAS, BIAS, BIAS);
            if (F16) z = make_int4(0, 0, 0, 0);
            else if (has_zq) z = *reinterpret_cast<const int4*>(zr + j);
            float2 r0, r1;
            if (FOLD) {
              // fma(F, c, off_hi) replaces {F - K, fma(., c, off)}; the 2^lo factor rides on the rounding add: fma(e, g, K)
              const float2 x0 = ffma2(make_float2(__int_as_float((int)v[j] + z.x), __int_as_float((int)v[j + 1] + z.y)), c2, oh2);
              const float2 x1 = ffma2(make_float2(__int_as_float((int)v[j + 2] + z.z), __int_as_float((int)v[j + 3] + z.w)), c2, oh2);
              float2 e0 = make_float2(ex2_approx(x0.x), ex2_approx(x0.y));
              float2 e1 = make_float2(ex2_approx(x1.x), ex2_approx(x1.y));
              if (CLAMP) {
                e0.x = fminf(e0.x * g_lo, pmax); e0.y = fminf(e0.y * g_lo, pmax);
                e1.x = fminf(e1.x * g_lo, pmax); e1.y = fminf(e1.y * g_lo, pmax);
                r0 = fadd2(e0, k2); r1 = fadd2(e1, k2);
              } else {
                r0 = ffma2(e0, g2, k2); r1 = ffma2(e1, g2, k2);
              }
            } else {
              const float2 x0 = ffma2(tof2((int)v[j] + z.x, (int)v[j + 1] + z.y), c2, off2);
              const float2 x1 = ffma2(tof2((int)v[j + 2] + z.z, (int)v[j + 3] + z.w), c2, off2);
              float2 e0 = make_float2(ex2_approx(x0.x), ex2_approx(x0.y));
              float2 e1 = make_float2(ex2_approx(x1.x), ex2_approx(x1.y));
              if (CLAMP) {
                e0.x = fminf(e0.x, pmax); e0.y = fminf(e0.y, pmax);
                e1.x = fminf(e1.x, pmax); e1.y = fminf(e1.y, pmax);
              }
              r0 = fadd2(e0, k2); r1 = fadd2(e1, k2);
            }
            cd[j] = __float_as_uint(r0.x); cd[j + 1] = __float_as_uint(r0.y);
            cd[j + 2] = __float_as_uint(r1.x); cd[j + 3] = __float_as_uint(r1.y);
          }
        };
        if (warp_clamps) codes(std::true_type{}); else codes(std::false_type{});
        if (j0 + 32 > p.Tk) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j0 + j >= p.Tk) cd[j] = 0x4B400000u;   // code 0
        }
        // two groups of 16 keys; inside a group key 8a+2b+c goes to byte 4b+2a+c (att_vt_perm)
#pragma unroll
        for (int gq = 0; gq < 2; ++gq) {
          const uint32_t* e = cd + 16 * gq;
          uint32_t lo[4], hi[4];
#pragma unroll
          for (int bq = 0; bq < 4; ++bq) {
            // [a.b0 a.b1 b.b0 b.b1] and [c.b0 c.b1 d.b0 d.b1], then even bytes = lo plane, odd bytes = hi plane
            const uint32_t u = __byte_perm(e[2 * bq], e[2 * bq + 1], 0x5410);
            const uint32_t w = __byte_perm(e[8 + 2 * bq], e[9 + 2 * bq], 0x5410);
            lo[bq] = __byte_perm(u, w, 0x6420);
            if (SM16) hi[bq] = __byte_perm(u, w, 0x7531);
          }
          const int chunk = (col0 >> 4) + gq;                      // 16-byte chunk of the 128-key row
          const int sw = (chunk ^ (row & 7)) << 4;
          *reinterpret_cast<uint4*>(pl + sw) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          if (SM16) *reinterpret_cast<uint4*>(phh + sw) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        }
      }
      fence_proxy_async();       // P bytes (generic proxy) -> visible to the MMA (async proxy)
      __syncwarp();
      if (lane == 0) {
        if (!F16) mbar_arrive(&kv_empty[st]);
        mbar_arrive(&p_full[pb]);
      }
      if (++pb == npb) { pb = 0; ph_p ^= 1; }
    }
    // ---- epilogue: O = (256*hi + lo - zv*rowsum) * out_scale
    mbar_wait(o_done, 0);
    tc_fence_after();
    if (part == 0) {
      const int grow = row_base + row;
      float rs = 0.f;
      {
        uint32_t v[16];
        tmem_ld_32x16(tm_olo + t_lane + (uint32_t)((d >> 4) << 4), v);
        tmem_ld_wait();
        rs = (float)(int)v[d & 15];
        if (SM16) {
          uint32_t w[16];
          tmem_ld_32x16(tm_ohi + t_lane + (uint32_t)((d >> 4) << 4), w);
          tmem_ld_wait();
          rs += 256.0f * (float)(int)w[d & 15];
        }
      }
      const float zc = (float)p.zv * rs;
      const QuantK qk = make_quantk(p.oq);
      for (int c0 = 0; c0 < d; c0 += 16) {
        uint32_t v[16], w[16];
        tmem_ld_32x16(tm_olo + t_lane + (uint32_t)c0, v);
        if (SM16) tmem_ld_32x16(tm_ohi + t_lane + (uint32_t)c0, w);
        tmem_ld_wait();
        if (grow < p.Tq) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            if (c0 + j < d) {    // d % 8 == 0: handle 4 columns at a time
              float y[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float val = (float)(int)v[j + e];
                if (SM16) val += 256.0f * (float)(int)w[j + e];
                y[e] = (val - zc) * p.out_scale;
              }
              const long long o = ((long long)b * p.Tq + grow);
              const int col = h * d + c0 + j;
              if (p.out) *reinterpret_cast<float4*>(p.out + o * p.ld_out + col) = make_float4(y[0], y[1], y[2], y[3]);
              if (p.out_q)
                *reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(p.out_q) + o * p.ld_out_q + col) =
                    quant_code(y[0], qk) | (quant_code(y[1], qk) << 8) | (quant_code(y[2], qk) << 16) | (quant_code(y[3], qk) << 24);
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace qd
