/*
 * gemm_tc.cu — TF32 tensor-core form of the FcLayer contractions on sm_100a (PS_FC_TF32):
 * tcgen05.mma (kind::tf32, fp32 accumulate in TMEM) fed by TMA (cp.async.bulk.tensor, 128 B
 * swizzle) through a 4-stage mbarrier pipeline; 128 threads per CTA:
 *     warp 0 / one lane  : TMA producer
 *     warp 1 / one lane  : MMA issuer (tcgen05.mma + tcgen05.commit)
 *     warps 0-3          : epilogue (tcgen05.ld 32x32b → registers → bias / activation /
 *                          activation derivative → global), each warp its own 32 TMEM lanes
 * One output tile (128 x BLOCK_N) per CTA, optional split over K (wgrad: K = batch).
 *
 * All three contractions are expressed as C[M][N] = sum_k A[m][k] * B[n][k] with BOTH operands
 * K-major (K contiguous), the one operand layout whose 128B-swizzle UMMA descriptor is the
 * plain canonical form; the transposed operand copies this needs (activations^T, delta^T, W^T)
 * are written by the producing epilogues, where the TMEM register layout (one row per lane)
 * makes the transposed store the naturally coalesced one.
 *
 *   forward  Z  [B][out]   : A = act [B][in],        B = W  [out][in]
 *   dgrad    dX [B][in]    : A = delta [B][out],     B = Wt [in][out]
 *   wgrad    G  [out][in+1]: A = delta^T [out][B],   B = [act | 1]^T [in+1][B]   (split over B)
 *
 * Tensor maps carry the LOGICAL extents, so every ragged edge (K not a multiple of 32, N not a
 * multiple of BLOCK_N, batch tail) is zero-filled by TMA; stores are clipped.
 */
#include <cuda.h>

#include <map>
#include <mutex>
#include <tuple>

#include "gemm.cuh"

namespace psb {

namespace {

constexpr int BM = 128, BK = 32, UMMA_K = 8;
/* pipeline depth: 8 stages cover a whole K = 256 operand in ONE L2 round trip; the 3xTF32 stages are twice as large */
/* 3xTF32 tiles up to 64 columns run TWO CTAs per SM (2 stages of 48 KB each, <= 128 registers): a 128-CTA GEMM no longer owns the GPU,
 * so the weight-gradient GEMM on the side stream genuinely overlaps the dgrad chain, and one CTA's TMA latency / epilogue hides
 * behind the other's MMAs */
/* DEEP: the same tiles with FOUR stages and one CTA per SM — for grids that fit one wave at one CTA per SM anyway (a 128-CTA GEMM at
 * batch 4096: the second CTA slot stays empty and two stages leave the TMA latency of 8 k-blocks exposed: cfg2 159.2 -> 151.7 us per step)
 * and for the weight-gradient GEMMs, whose K loop is the batch.  Chosen per launch (dispatch_tc, PS_TC_DEEP / PS_TC_DEEP_WGRAD). */
template <int BLOCK_N, bool SPLIT, bool DEEP> struct Depth {
  static constexpr int STAGES = SPLIT ? (BLOCK_N > 64 ? 3 : (DEEP ? 4 : 2)) : (BLOCK_N > 64 ? 6 : 8);
  static constexpr int MIN_CTAS = (SPLIT && BLOCK_N <= 64 && !DEEP) ? 2 : 1;
};
enum { EPI_FWD = 0, EPI_DGRAD = 1, EPI_WGRAD = 2 };

struct TcParams {
  int M, N, K;
  int kb_per_split;            /* k-blocks (of 32) handled by one blockIdx.z */
  float* C; long ldc; size_t slab;
  float* Ct; long ldct;        /* transposed copy Ct[n][m] or null */
  const float* bias; int act;  /* EPI_FWD */
  const float* Yt; long ldyt;  /* EPI_DGRAD: activation of the layer below, TRANSPOSED [n][m] */
};

/* ---- PTX wrappers ------------------------------------------------------------------ */
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld16_issue(uint32_t taddr, uint32_t* r) {   /* caller waits with tcgen05.wait::ld */
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

/* shared-memory matrix descriptor, K-major, 128 B swizzle: rows are 128 B apart, 8-row groups
 * 1024 B apart (SBO); LBO is unused for swizzled K-major; bit 46 = descriptor version 1 (sm_100);
 * bits 61-63 = 2 (SWIZZLE_128B).                                                             */
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

/* instruction descriptor (kind::tf32): D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2),
 * both K-major (bits 15, 16 = 0), N >> 3 at bits 17-22, M >> 4 at bits 24-28                 */
__host__ __device__ constexpr uint32_t make_idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

/* SPLIT (3xTF32): every stage also holds the residual tiles A_lo = A - tf32(A), B_lo = B - tf32(B) */
template <int BLOCK_N, bool SPLIT, bool DEEP = false>
struct SmemLayout {
  static constexpr uint32_t A_BYTES = BM * BK * 4;
  static constexpr uint32_t B_BYTES = BLOCK_N * BK * 4;
  static constexpr int STAGES = Depth<BLOCK_N, SPLIT, DEEP>::STAGES;
  static constexpr uint32_t STAGE_BYTES = (SPLIT ? 2u : 1u) * (A_BYTES + B_BYTES);
  static constexpr uint32_t BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr uint32_t TOTAL = BAR_OFF + 256 + 1024;   /* + barriers/tmem slot + manual 1024 B alignment slack */
};

/* PRE_B (3xTF32 only): the B operand is a weight matrix whose residual B_lo = B - tf32(B) is kept in HBM by the kernels that write the
 * weights (dense_update / split_lo) and arrives by TMA like B itself; only the A tiles (activations, deltas) are split in the kernel */
template <int BLOCK_N, int EPI, bool SPLIT, bool PRE_B, bool DEEP>
__device__ __forceinline__ void gemm_tf32_body(const CUtensorMap* tmA_, const CUtensorMap* tmB_, const CUtensorMap* tmBlo_, const TcParams& p,
                                               int bx, int by, int bz) {
  const CUtensorMap& tmA = *tmA_; const CUtensorMap& tmB = *tmB_; const CUtensorMap& tmBlo = *tmBlo_;
  pdl_launch_dependents();                     /* the next kernel of the chain may set itself up while this one runs */
  using SL = SmemLayout<BLOCK_N, SPLIT, DEEP>;
  constexpr int STAGES = SL::STAGES;
  constexpr int TMEM_COLS = BLOCK_N <= 32 ? 32 : (BLOCK_N <= 64 ? 64 : (BLOCK_N <= 128 ? 128 : 256));
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  /* stage s: [A | B | A_lo | B_lo] (the residual tiles only with SPLIT); all tile bases stay 1024 B aligned */
  const uint32_t stage0 = base, bars = base + SL::BAR_OFF;
  constexpr uint32_t OFF_B = SL::A_BYTES, OFF_ALO = SL::A_BYTES + SL::B_BYTES, OFF_BLO = 2 * SL::A_BYTES + SL::B_BYTES;
  /* bars: full[STAGES] | empty[STAGES] | split[STAGES] | tmem_full | tmem slot */
  const uint32_t full0 = bars, empty0 = bars + 8 * STAGES, split0 = bars + 16 * STAGES, tfull = bars + 24 * STAGES, tslot = bars + 24 * STAGES + 8;
  volatile uint32_t* tslot_gen = reinterpret_cast<volatile uint32_t*>(gen_base + SL::BAR_OFF + 24 * STAGES + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = by * BM, n0 = bx * BLOCK_N;
  const int nkb_total = (p.K + BK - 1) / BK;
  const int kb_begin = bz * p.kb_per_split;
  const int num_kb = max(0, min(nkb_total, kb_begin + p.kb_per_split) - kb_begin);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); mbar_init(split0 + 8 * s, 192); }
    mbar_init(tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    if (PRE_B) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBlo) : "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tslot), "r"((uint32_t)TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot_gen;
  /* launched as a programmatic dependent of the previous GEMM of the forward / dgrad chain (launch_tc): everything above
   * (barrier init, tensor-map prefetch, TMEM allocation) overlapped that kernel's tail; its results are read from here on */
  pdl_wait();

  if (warp == 0 && lane == 0) {
    /* ===== TMA producer ===== */
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % STAGES;
      mbar_wait(empty0 + 8 * s, ((kb / STAGES) & 1) ^ 1);
      mbar_expect_tx(full0 + 8 * s, SL::A_BYTES + (PRE_B ? 2u : 1u) * SL::B_BYTES);
      const int k = (kb_begin + kb) * BK;
      tma_load_2d(stage0 + s * SL::STAGE_BYTES, &tmA, full0 + 8 * s, k, m0);
      tma_load_2d(stage0 + s * SL::STAGE_BYTES + OFF_B, &tmB, full0 + 8 * s, k, n0);
      if (PRE_B) tma_load_2d(stage0 + s * SL::STAGE_BYTES + OFF_BLO, &tmBlo, full0 + 8 * s, k, n0);
    }
  } else if (warp == 1 && lane == 0) {
    /* ===== MMA issuer ===== */
    constexpr uint32_t idesc = make_idesc_tf32(BM, BLOCK_N);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % STAGES;
      mbar_wait((SPLIT ? split0 : full0) + 8 * s, (kb / STAGES) & 1);
      tc_fence_after();
      const uint32_t st = stage0 + s * SL::STAGE_BYTES;
      const uint64_t ad = make_kmajor_sw128_desc(st);
      const uint64_t bd = make_kmajor_sw128_desc(st + OFF_B);
      if (SPLIT) {
        /* 3xTF32: the tensor core truncates each operand to its top 19 bits, so the raw tile IS the
         * high part; a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi (small terms first), fp32 accumulate */
        const uint64_t al = make_kmajor_sw128_desc(st + OFF_ALO);
        const uint64_t bl = make_kmajor_sw128_desc(st + OFF_BLO);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          tc_mma_tf32(tmem, al + 2 * k, bd + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          tc_mma_tf32(tmem, ad + 2 * k, bl + 2 * k, idesc, 1u);
          tc_mma_tf32(tmem, ad + 2 * k, bd + 2 * k, idesc, 1u);
        }
      } else {
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k)   /* advance 32 B inside the 128 B swizzle atom: +2 in the 16 B-unit address field */
          tc_mma_tf32(tmem, ad + 2 * k, bd + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
      }
      tc_commit(empty0 + 8 * s);               /* frees the stage once these MMAs have read it */
    }
    if (num_kb > 0) tc_commit(tfull);          /* accumulator complete */
  } else if (SPLIT && warp >= 2) {
    /* ===== residual producers (warps 2-7, 192 threads): lo = x - tf32(x), element-wise, so the swizzled tile
     * layout carries over byte for byte; generic-proxy stores are fenced for the async proxy ===== */
    const int t = threadIdx.x - 64;                 /* 0..191 */
    constexpr int CHUNKS = (int)((SL::A_BYTES + (PRE_B ? 0u : SL::B_BYTES)) / 16);   /* [A | B] are contiguous, and so are [A_lo | B_lo] */
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % STAGES;
      mbar_wait(full0 + 8 * s, (kb / STAGES) & 1);
      uint8_t* src = gen_base + s * SL::STAGE_BYTES;
      uint8_t* dst = src + OFF_ALO;
#pragma unroll 4
      for (int c = t; c < CHUNKS; c += 192) {
        const float4 x = *reinterpret_cast<const float4*>(src + 16 * c);
        float4 hi, lo;
        hi.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u); lo.x = x.x - hi.x;
        hi.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u); lo.y = x.y - hi.y;
        hi.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u); lo.z = x.z - hi.z;
        hi.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u); lo.w = x.w - hi.w;
        /* the raw tile stays as it is: kind::tf32 TRUNCATES its operands to the top 19 bits (scripts/tf32_rounding_probe.py on B200:
         * 1 + 3*2^-12 -> 1, -(1 + 3*2^-12) -> -1, 1 + 2^-11 + 2^-20 -> 1), so the tensor core reads exactly `hi` from it */
        *reinterpret_cast<float4*>(dst + 16 * c) = lo;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(split0 + 8 * s) : "memory");
    }
  }
  __syncwarp();

  /* ===== epilogue: warp w owns TMEM lanes [32w, 32w+32) = output rows m0 + 32w + lane =====
   * All TMEM loads of the tile are issued first; the epilogue operands (bias: warp-uniform; the
   * activation of the layer below for its derivative: read from the TRANSPOSED copy, so the 32
   * lanes = 32 consecutive rows read one 128 B line per column) are fetched while they fly.   */
  if (warp < 4) {   /* the four TMEM lane quadrants; with SPLIT warps 4-7 only produced residual tiles */
  /* the tile is drained in passes of at most 64 columns (register budget: 64 accumulators + 64 epilogue operands) */
  constexpr int PN = (SPLIT && BLOCK_N <= 64) ? (BLOCK_N > 32 ? 32 : BLOCK_N) : (BLOCK_N > 64 ? 64 : BLOCK_N);   /* 2 CTAs / SM: half the epilogue registers */
  constexpr int NPASS = BLOCK_N / PN;
  constexpr int NCH = PN / 16;
  const int m = m0 + warp * 32 + lane;
  const bool row_ok = m < p.M;
  float* Cz = p.C + (size_t)bz * p.slab;
  const bool vec_ok = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(Cz) & 15) == 0);
  if (num_kb > 0) {
    mbar_wait(tfull, 0);
    tc_fence_after();
  }
#pragma unroll 1
  for (int pass = 0; pass < NPASS; ++pass) {
  const int np0 = n0 + pass * PN;
  if (np0 >= p.N) break;                         /* warp-uniform */
  uint32_t r[NCH][16];
  if (num_kb > 0) {
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) tc_ld16_issue(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(pass * PN + ch * 16), r[ch]);
  } else {
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
      for (int i = 0; i < 16; ++i) r[ch][i] = 0u;
  }
  float e[NCH][16];
  if (EPI == EPI_FWD) {
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
      for (int i = 0; i < 16; ++i) { const int n = np0 + ch * 16 + i; e[ch][i] = n < p.N ? __ldg(p.bias + n) : 0.f; }
  } else if (EPI == EPI_DGRAD) {
    if (p.act != PS_ACT_NONE) {
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
        for (int i = 0; i < 16; ++i) { const int n = np0 + ch * 16 + i; e[ch][i] = (row_ok && n < p.N) ? __ldg(p.Yt + (long)n * p.ldyt + m) : 0.f; }
    }
  }
  if (num_kb > 0) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int nb = np0 + ch * 16;
    if (nb >= p.N) continue;                   /* warp-uniform */
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[ch][i]);
    if (EPI == EPI_FWD) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = act_forward(p.act, __fadd_rn(v[i], e[ch][i]));
    } else if (EPI == EPI_DGRAD) {
      if (p.act != PS_ACT_NONE) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = act_backward(p.act, v[i], e[ch][i]);
      }
    }
    if (row_ok) {
      float* row = Cz + (long)m * p.ldc + nb;
      if (vec_ok && nb + 15 < p.N) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) st_f4(row + i, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) if (nb + i < p.N) row[i] = v[i];
      }
      if (EPI != EPI_WGRAD && p.Ct) {
#pragma unroll
        for (int i = 0; i < 16; ++i) if (nb + i < p.N) p.Ct[(long)(nb + i) * p.ldct + m] = v[i];
      }
    }
  }
  }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

template <int BLOCK_N, int EPI, bool SPLIT, bool PRE_B, bool DEEP>
__global__ void __launch_bounds__(SPLIT ? 256 : 128, Depth<BLOCK_N, SPLIT, DEEP>::MIN_CTAS) gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                        const __grid_constant__ CUtensorMap tmBlo, const TcParams p) {
  gemm_tf32_body<BLOCK_N, EPI, SPLIT, PRE_B, DEEP>(&tmA, &tmB, &tmBlo, p, blockIdx.x, blockIdx.y, blockIdx.z);
}

/* Every FcLayer's weight-gradient GEMM of a step in ONE launch (FcLayer.java:103-106 for all layers): the contractions are independent
 * once the dgrad chain has produced the deltas, each alone fills barely half the GPU (80-112 CTAs), together they are one wave at two
 * CTAs per SM — and the dgrad chain before them runs without a competitor for the SMs.                                          */
constexpr int kMaxGroup = 8;
struct GroupedTc {
  CUtensorMap tmA[kMaxGroup], tmB[kMaxGroup];
  TcParams p[kMaxGroup];
  int first[kMaxGroup + 1], gx[kMaxGroup], gy[kMaxGroup];
  int n;
};
template <int BLOCK_N, bool SPLIT>
__global__ void __launch_bounds__(SPLIT ? 256 : 128, Depth<BLOCK_N, SPLIT, false>::MIN_CTAS) gemm_tf32_grouped_wgrad_kernel(const __grid_constant__ GroupedTc g) {
  int pi = 0;
  while (pi + 1 < g.n && (int)blockIdx.x >= g.first[pi + 1]) ++pi;
  const int local = (int)blockIdx.x - g.first[pi];
  const int bx = local % g.gx[pi], by = (local / g.gx[pi]) % g.gy[pi], bz = local / (g.gx[pi] * g.gy[pi]);
  gemm_tf32_body<BLOCK_N, EPI_WGRAD, SPLIT, false, false>(&g.tmA[pi], &g.tmB[pi], &g.tmB[pi], g.p[pi], bx, by, bz);
}

/* ---- host: tensor maps ---------------------------------------------------------------- */
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  PS_REQUIRE(fn != nullptr, PS_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  return fn;
}

/* a K-major fp32 operand [rows][k_extent], leading dimension ld floats; box = 32 floats x box_rows */
const CUtensorMap& tensor_map(const float* ptr, int rows, int k_extent, long ld, int box_rows) {
  using Key = std::tuple<const float*, int, int, long, int>;
  static std::map<Key, CUtensorMap> cache;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  const Key key(ptr, rows, k_extent, ld, box_rows);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  PS_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ld % 4 == 0, PS_ERR_ARG, "tf32 gemm: operand must be 16 B aligned with ld % 4 == 0");
  CUtensorMap m;
  const cuuint64_t dims[2] = {(cuuint64_t)k_extent, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PS_REQUIRE(r == CUDA_SUCCESS, PS_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  if (cache.size() > 4096) cache.clear();
  return cache.emplace(key, m).first->second;
}

template <int BLOCK_N, int EPI, bool SPLIT, bool PRE_B, bool DEEP>
void launch_tc(Ctx* ctx, const float* A, long lda, const float* B, long ldb, const float* Blo, TcParams& p, int nsplit) {
  using SL = SmemLayout<BLOCK_N, SPLIT, DEEP>;
  const CUtensorMap ta = tensor_map(A, p.M, p.K, lda, BM);
  const CUtensorMap tb = tensor_map(B, p.N, p.K, ldb, BLOCK_N);
  const CUtensorMap tbl = PRE_B ? tensor_map(Blo, p.N, p.K, ldb, BLOCK_N) : tb;
  const int nkb = (p.K + BK - 1) / BK;
  p.kb_per_split = (nkb + nsplit - 1) / nsplit;
  dim3 grid(ceil_div(p.N, BLOCK_N), ceil_div(p.M, BM), nsplit);
  if (ctx->pdl && ctx->pdl_gemm && EPI != EPI_WGRAD) {      /* forward and dgrad GEMMs follow each other on the main stream */
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(SPLIT ? 256 : 128); cfg.dynamicSmemBytes = SL::TOTAL; cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    PS_CUDA(cudaLaunchKernelEx(&cfg, gemm_tf32_kernel<BLOCK_N, EPI, SPLIT, PRE_B, DEEP>, ta, tb, tbl, p));
  } else {
    gemm_tf32_kernel<BLOCK_N, EPI, SPLIT, PRE_B, DEEP><<<grid, SPLIT ? 256 : 128, SL::TOTAL, ctx->stream>>>(ta, tb, tbl, p);
  }
  PS_LAUNCH_CHECK();
  ctx->launches++;
}

template <int BLOCK_N, int EPI>
void launch_tc_mode(Ctx* ctx, const float* A, long lda, const float* B, long ldb, const float* Blo, TcParams& p, int nsplit) {
  if (ctx->fc_precision != PS_FC_TF32X3) { launch_tc<BLOCK_N, EPI, false, false, false>(ctx, A, lda, B, ldb, nullptr, p, nsplit); return; }
  /* four stages at one CTA per SM, or two stages at two (see Depth) */
  const long ctas = (long)ceil_div(p.N, BLOCK_N) * ceil_div(p.M, BM) * nsplit;
  const int mode = EPI == EPI_WGRAD ? ctx->tc_deep_wgrad : ctx->tc_deep;
  const bool deep = BLOCK_N <= 64 && (mode == 1 || (mode == 2 && ctas <= ctx->num_sms));
  if (Blo != nullptr && EPI != EPI_WGRAD) {
    if (deep) launch_tc<BLOCK_N, EPI, true, true, (BLOCK_N <= 64)>(ctx, A, lda, B, ldb, Blo, p, nsplit);
    else launch_tc<BLOCK_N, EPI, true, true, false>(ctx, A, lda, B, ldb, Blo, p, nsplit);
  } else {
    if (deep) launch_tc<BLOCK_N, EPI, true, false, (BLOCK_N <= 64)>(ctx, A, lda, B, ldb, nullptr, p, nsplit);
    else launch_tc<BLOCK_N, EPI, true, false, false>(ctx, A, lda, B, ldb, nullptr, p, nsplit);
  }
}

template <int EPI>
void dispatch_tc(Ctx* ctx, const float* A, long lda, const float* B, long ldb, const float* Blo, TcParams& p, int nsplit) {
  /* 128-wide tiles when 64-wide ones would not fit one wave (1 CTA per SM): fc0's dgrad at cfg2 is 7 x 32 = 224 CTAs
   * at 64 columns, 4 x 32 = 128 at 128 — and every A tile is read (and split) half as often */
  const int ctas_per_sm = ctx->fc_precision == PS_FC_TF32X3 ? 2 : 1;
  const long ctas64 = (long)ceil_div(p.N, 64) * ceil_div(p.M, BM) * nsplit, ctas128 = (long)ceil_div(p.N, 128) * ceil_div(p.M, BM) * nsplit;
  /* ... and when the 64-wide grid needs the second CTA slot of the SMs while the 128-wide one is a single wave at one CTA per SM, unless
   * the wider tiles pad N by more than an eighth (measured: cfg3's 8192 x 256 x 256 layers 289.9 -> 274.8 us per step; cfg2's fc0 dgrad,
   * N = 413 -> 512 columns, 147.1 -> 150.6: excluded).  PS_TC_WIDE_RULE=0: the first rule only */
  const bool low_pad = (long)ceil_div(p.N, 128) * 128 - p.N <= p.N / 8;
  const bool wide = p.N >= 128 && (ctas64 > (long)ctas_per_sm * ctx->num_sms || (ctx->tc_wide_rule == 1 && low_pad && ctas64 > ctx->num_sms && ctas128 <= ctx->num_sms));
  /* 64-column tiles that leave the second CTA slot of every SM empty (a 128-CTA grid): 32-column tiles fill it */
  const bool narrow = ctx->gemm_narrow && ctas_per_sm == 2 && p.N > 32 && (long)ceil_div(p.N, 64) * ceil_div(p.M, BM) * nsplit <= ctx->num_sms;
  if (wide) launch_tc_mode<128, EPI>(ctx, A, lda, B, ldb, Blo, p, nsplit);
  else if (narrow) launch_tc_mode<32, EPI>(ctx, A, lda, B, ldb, Blo, p, nsplit);
  else if (p.N <= 16) launch_tc_mode<16, EPI>(ctx, A, lda, B, ldb, Blo, p, nsplit);
  else if (p.N <= 32) launch_tc_mode<32, EPI>(ctx, A, lda, B, ldb, Blo, p, nsplit);
  else launch_tc_mode<64, EPI>(ctx, A, lda, B, ldb, Blo, p, nsplit);
}

template <int BLOCK_N, bool SPLIT, bool DEEP>
void set_attr() {
  const int bytes = (int)SmemLayout<BLOCK_N, SPLIT, DEEP>::TOTAL;
  PS_CUDA(cudaFuncSetAttribute(gemm_tf32_kernel<BLOCK_N, EPI_FWD, SPLIT, false, DEEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  PS_CUDA(cudaFuncSetAttribute(gemm_tf32_kernel<BLOCK_N, EPI_DGRAD, SPLIT, false, DEEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  PS_CUDA(cudaFuncSetAttribute(gemm_tf32_kernel<BLOCK_N, EPI_WGRAD, SPLIT, false, DEEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  if (SPLIT) {
    PS_CUDA(cudaFuncSetAttribute(gemm_tf32_kernel<BLOCK_N, EPI_FWD, SPLIT, SPLIT, DEEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    PS_CUDA(cudaFuncSetAttribute(gemm_tf32_kernel<BLOCK_N, EPI_DGRAD, SPLIT, SPLIT, DEEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  }
}

}  // namespace

void fc_tf32_init() {
  static bool done = false;
  if (done) return;
  set_attr<16, false, false>(); set_attr<32, false, false>(); set_attr<64, false, false>(); set_attr<128, false, false>();
  set_attr<16, true, false>(); set_attr<32, true, false>(); set_attr<64, true, false>(); set_attr<128, true, false>();
  set_attr<16, true, true>(); set_attr<32, true, true>(); set_attr<64, true, true>();
  PS_CUDA(cudaFuncSetAttribute(gemm_tf32_grouped_wgrad_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemLayout<64, true, false>::TOTAL));
  PS_CUDA(cudaFuncSetAttribute(gemm_tf32_grouped_wgrad_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemLayout<64, false, false>::TOTAL));
  encode_fn();
  done = true;
}

void fc_forward_tf32(Ctx* ctx, const FcFwdArgs& a) {
  fc_tf32_init();
  TcParams p{};
  p.M = a.B; p.N = a.out; p.K = a.in;
  p.C = a.Z; p.ldc = a.ldz; p.slab = 0; p.Ct = a.Zt; p.ldct = a.ldzt;
  p.bias = a.bias; p.act = a.act;
  dispatch_tc<EPI_FWD>(ctx, a.A, a.lda, a.W, a.ldw, a.Wlo, p, 1);
}

void fc_dgrad_tf32(Ctx* ctx, const FcDgradArgs& a) {
  PS_REQUIRE(a.Wt != nullptr, PS_ERR_ARG, "tf32 dgrad needs the transposed weight copy");
  TcParams p{};
  p.M = a.B; p.N = a.n_cols; p.K = a.out;
  p.C = a.dX; p.ldc = a.ldx; p.slab = 0; p.Ct = a.dXt; p.ldct = a.ldxt;
  p.act = a.act_below; p.Yt = a.Yt; p.ldyt = a.ldyt;
  PS_REQUIRE(p.act == PS_ACT_NONE || p.Yt != nullptr, PS_ERR_ARG, "tf32 dgrad needs the transposed activation of the layer below");
  dispatch_tc<EPI_DGRAD>(ctx, a.dl, a.ldd, a.Wt, a.ldwt, a.Wtlo, p, 1);
}

/* all layers whose [A | 1]^T is at most 64 x ... tiles wide share the 64-column instantiation; returns false when a layer does not fit
 * the group (the caller then launches the layers one by one) */
bool fc_wgrad_grouped_tf32(Ctx* ctx, const FcWgradArgs* a, int n) {
  fc_tf32_init();
  if (n < 2 || n > kMaxGroup) return false;
  GroupedTc g{};
  g.n = n;
  int total = 0;
  for (int i = 0; i < n; ++i) {
    if (a[i].dlT == nullptr || a[i].AT == nullptr) return false;
    TcParams& p = g.p[i];
    p = TcParams{};
    p.M = a[i].out; p.N = a[i].in + 1; p.K = a[i].B;
    if (p.N < 64) return false;                  /* narrower layers use narrower tiles: launched one by one */
    p.C = a[i].G; p.ldc = a[i].ldg; p.slab = a[i].slab; p.Ct = nullptr;
    const int nkb = (p.K + BK - 1) / BK;
    p.kb_per_split = (nkb + a[i].nsplit - 1) / a[i].nsplit;
    g.tmA[i] = tensor_map(a[i].dlT, p.M, p.K, a[i].ldt, BM);
    g.tmB[i] = tensor_map(a[i].AT, p.N, p.K, a[i].ldt, 64);
    g.gx[i] = ceil_div(p.N, 64); g.gy[i] = ceil_div(p.M, BM);
    g.first[i] = total;
    total += g.gx[i] * g.gy[i] * a[i].nsplit;
  }
  g.first[n] = total;
  if (ctx->fc_precision == PS_FC_TF32X3)
    gemm_tf32_grouped_wgrad_kernel<64, true><<<total, 256, SmemLayout<64, true, false>::TOTAL, ctx->stream>>>(g);
  else
    gemm_tf32_grouped_wgrad_kernel<64, false><<<total, 128, SmemLayout<64, false, false>::TOTAL, ctx->stream>>>(g);
  PS_LAUNCH_CHECK();
  ctx->launches++;
  return true;
}

void fc_wgrad_tf32(Ctx* ctx, const FcWgradArgs& a) {
  PS_REQUIRE(a.dlT != nullptr && a.AT != nullptr, PS_ERR_ARG, "tf32 wgrad needs the transposed delta / activation copies");
  TcParams p{};
  p.M = a.out; p.N = a.in + 1; p.K = a.B;
  p.C = a.G; p.ldc = a.ldg; p.slab = a.slab; p.Ct = nullptr;
  dispatch_tc<EPI_WGRAD>(ctx, a.dlT, a.ldt, a.AT, a.ldt, nullptr, p, a.nsplit);
}

}  // namespace psb
