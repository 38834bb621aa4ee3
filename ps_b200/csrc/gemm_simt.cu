/*
 * gemm_simt.cu — fp32 FFMA form of the FcLayer contractions (PS_FC_FP32, the exact mode:
 * same arithmetic type as jblas' sgemm, results agree with the ordered-loop oracle to ~1e-6).
 * 64x64x16 tiles, 256 threads, 4x4 register micro-tiles, 128-bit global and shared accesses.
 * The TF32 tensor-core form lives in gemm_tc.cu.
 */
#include "gemm.cuh"

namespace psb {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, PADS = 4;

enum { EPI_FWD = 0, EPI_DGRAD = 1, EPI_WGRAD = 2 };

struct SimtParams {
  int M, N, K;
  const float* A; long sam, sak;     /* A(m,k) = A[m*sam + k*sak]; exactly one stride is 1 */
  const float* B; long sbn, sbk;     /* B(n,k) = B[n*sbn + k*sbk] */
  int vecA, vecB;                    /* 128-bit loads legal (alignment and leading dimension) */
  float* C; long ldc; size_t slab;   /* C(m,n) = C[z*slab + m*ldc + n] */
  int kchunk;                        /* K range per blockIdx.z */
  const float* bias; int act;        /* EPI_FWD */
  const float* Y; long ldy;          /* EPI_DGRAD */
};

template <bool KC>
__device__ __forceinline__ void load_tile(float (*S)[BM + PADS], const float* __restrict__ P, long s_mn, long s_k, int mn0, int k0, int MN,
                                          int Kend, bool vec, int t) {
  if (KC) {                          /* k contiguous: thread reads 4 consecutive k of one row */
    const int r = t >> 2, kq = (t & 3) << 2;
    const int mn = mn0 + r, k = k0 + kq;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (mn < MN) {
      const float* p = P + (long)mn * s_mn + k;
      if (vec && k + 3 < Kend) { const float4 q = __ldg(reinterpret_cast<const float4*>(p)); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
      else {
#pragma unroll
        for (int i = 0; i < 4; ++i) if (k + i < Kend) v[i] = __ldg(p + i);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) S[kq + i][r] = v[i];
  } else {                           /* m/n contiguous: thread reads 4 consecutive rows at one k */
    const int kk = t >> 4, q4 = (t & 15) << 2;
    const int mn = mn0 + q4, k = k0 + kk;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (k < Kend) {
      const float* p = P + (long)k * s_k + mn;
      if (vec && mn + 3 < MN) { const float4 q = __ldg(reinterpret_cast<const float4*>(p)); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
      else {
#pragma unroll
        for (int i = 0; i < 4; ++i) if (mn + i < MN) v[i] = __ldg(p + i);
      }
    }
    *reinterpret_cast<float4*>(&S[kk][q4]) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

template <bool A_KC, bool B_KC, int EPI>
__global__ void __launch_bounds__(256) gemm_simt_kernel(SimtParams p) {
  __shared__ __align__(16) float As[BK][BM + PADS];
  __shared__ __align__(16) float Bs[BK][BN + PADS];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * p.kchunk;
  const int kend = min(p.K, kbeg + p.kchunk);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    load_tile<A_KC>(As, p.A, p.sam, p.sak, m0, k0, p.M, kend, p.vecA != 0, t);
    load_tile<B_KC>(Bs, p.B, p.sbn, p.sbk, n0, k0, p.N, kend, p.vecB != 0, t);
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  float* Cz = p.C + (size_t)blockIdx.z * p.slab;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (EPI == EPI_FWD) v = act_forward(p.act, __fadd_rn(v, p.bias[n]));
      if (EPI == EPI_DGRAD) v = act_backward(p.act, v, p.Y[(long)m * p.ldy + n]);
      Cz[(long)m * p.ldc + n] = v;
    }
  }
}

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

template <bool A_KC, bool B_KC, int EPI>
void launch(Ctx* ctx, SimtParams& p, int nz) {
  dim3 grid(ceil_div(p.N, BN), ceil_div(p.M, BM), nz);
  gemm_simt_kernel<A_KC, B_KC, EPI><<<grid, 256, 0, ctx->stream>>>(p);
  PS_LAUNCH_CHECK();
  ctx->launches++;
}

}  // namespace

void fc_forward_fp32(Ctx* ctx, const FcFwdArgs& a) {
  SimtParams p{};
  p.M = a.B; p.N = a.out; p.K = a.in;
  p.A = a.A; p.sam = a.lda; p.sak = 1; p.vecA = aligned16(a.A) && a.lda % 4 == 0;
  p.B = a.W; p.sbn = a.ldw; p.sbk = 1; p.vecB = aligned16(a.W) && a.ldw % 4 == 0;
  p.C = a.Z; p.ldc = a.ldz; p.slab = 0; p.kchunk = round_up(a.in, BK);
  p.bias = a.bias; p.act = a.act;
  launch<true, true, EPI_FWD>(ctx, p, 1);
}

void fc_dgrad_fp32(Ctx* ctx, const FcDgradArgs& a) {
  SimtParams p{};
  p.M = a.B; p.N = a.n_cols; p.K = a.out;
  p.A = a.dl; p.sam = a.ldd; p.sak = 1; p.vecA = aligned16(a.dl) && a.ldd % 4 == 0;
  p.B = a.W; p.sbn = 1; p.sbk = a.ldw; p.vecB = aligned16(a.W) && a.ldw % 4 == 0;
  p.C = a.dX; p.ldc = a.ldx; p.slab = 0; p.kchunk = round_up(a.out, BK);
  p.act = a.act_below; p.Y = a.Y; p.ldy = a.ldy;
  launch<true, false, EPI_DGRAD>(ctx, p, 1);
}

void fc_wgrad_fp32(Ctx* ctx, const FcWgradArgs& a) {
  SimtParams p{};
  p.M = a.out; p.N = a.in + 1; p.K = a.B;         /* column `in` of [A | 1] yields the bias gradient */
  p.A = a.dl; p.sam = 1; p.sak = a.ldd; p.vecA = aligned16(a.dl) && a.ldd % 4 == 0;
  p.B = a.A; p.sbn = 1; p.sbk = a.lda; p.vecB = aligned16(a.A) && a.lda % 4 == 0;
  p.C = a.G; p.ldc = a.ldg; p.slab = a.slab;
  p.kchunk = round_up(ceil_div(a.B, a.nsplit), BK);
  PS_REQUIRE((long)p.kchunk * a.nsplit >= a.B, PS_ERR_ARG, "wgrad: split does not cover the batch");
  launch<false, false, EPI_WGRAD>(ctx, p, a.nsplit);
}

}  // namespace psb
