/*
 * gemm.cuh — the three dense contractions of layer/FcLayer.java, behind one interface that the
 * fp32 FFMA kernels (gemm_simt.cu, exact mode) and the TF32 tcgen05 kernels (gemm_tc.cu) share.
 *
 * Activations are stored batch-major: A_l is [B][ld_l] row-major (= the reference's
 * features x N column-major FloatMatrix), weights W_l are [out][ldw] row-major with the input
 * index contiguous (the transpose of the reference's out x in column-major storage; the C ABI
 * converts at get/put).  Column `in` of every activation buffer holds the constant 1.0 and the
 * matching weight column holds 0: the forward product ignores it, while the weight-gradient
 * contraction delta^T * [A | 1] delivers db = rowSums(delta) as column `in` for free.
 *
 *   forward  (FcLayer.java:74-91)   Z[b][o]  = sum_i A[b][i] * W[o][i] + bias[o];  A' = act(Z)
 *   dgrad    (FcLayer.java:108)     dX[b][i] = sum_o dl[b][o] * W[o][i], then the activation
 *                                   derivative of the layer below (FcLayer.java:100-102 of that layer)
 *   wgrad    (FcLayer.java:103-105) G[o][c]  = sum_b dl[b][o] * [A | 1][b][c]   (divided by N when applied)
 */
#pragma once
#include "common.cuh"

namespace psb {

struct FcFwdArgs {
  int B, in, out;
  const float* A; int lda;        /* [B][lda] */
  const float* W; int ldw;        /* [out][ldw] */
  const float* Wlo;               /* optional W - tf32(W), same layout (3xTF32: the residual arrives by TMA instead of being split per tile) */
  const float* bias;              /* [out] */
  int act;                        /* PS_ACT_NONE | RELU | SIGMOID (softmax is applied by the tail kernel) */
  float* Z; int ldz;              /* [B][ldz] */
  float* Zt; int ldzt;            /* optional transposed copy [out][ldzt] (operand of the TF32 wgrad) */
};

struct FcDgradArgs {
  int B, in, out;
  const float* dl; int ldd;       /* [B][ldd] delta of this layer (activation derivative already applied) */
  const float* W; int ldw;        /* [out][ldw] */
  const float* Wt; int ldwt;      /* [in][ldwt] transposed copy (TF32 path; may be null for fp32) */
  const float* Wtlo;              /* optional Wt - tf32(Wt) */
  int act_below;                  /* activation of the layer below, whose output is Y */
  const float* Y; int ldy;        /* [B][ldy] */
  const float* Yt; int ldyt;      /* transposed copy [in][ldyt] (TF32 path: coalesced read in the epilogue) */
  int n_cols;                     /* how many of the `in` columns are needed (F*D for the first layer) */
  float* dX; int ldx;             /* [B][ldx] */
  float* dXt; int ldxt;           /* optional transposed copy [in][ldxt] */
};

struct FcWgradArgs {
  int B, in, out;
  const float* dl; int ldd;       /* [B][ldd] */
  const float* A; int lda;        /* [B][lda], column `in` == 1 */
  const float* dlT; const float* AT; int ldt;   /* transposed copies [out][ldt], [in+1][ldt] (row `in` == 1): TF32 path */
  float* G; int ldg;              /* [nsplit][out][ldg] partial sums over batch chunks */
  size_t slab;                    /* elements between consecutive partial slabs */
  int nsplit;
};

void fc_forward_fp32(Ctx* ctx, const FcFwdArgs& a);
void fc_dgrad_fp32(Ctx* ctx, const FcDgradArgs& a);
void fc_wgrad_fp32(Ctx* ctx, const FcWgradArgs& a);
/* TF32 tcgen05 tensor-core forms (gemm_tc.cu): TMA-staged operands, accumulators in TMEM */
void fc_forward_tf32(Ctx* ctx, const FcFwdArgs& a);
void fc_dgrad_tf32(Ctx* ctx, const FcDgradArgs& a);
void fc_wgrad_tf32(Ctx* ctx, const FcWgradArgs& a);
/* the weight-gradient GEMMs of n layers in one launch (false: not groupable — launch them one by one) */
bool fc_wgrad_grouped_tf32(Ctx* ctx, const FcWgradArgs* a, int n);
void fc_tf32_init();   /* one-time kernel attributes (must not happen inside a stream capture) */

#if defined(__CUDACC__)
/* activations/Sigmoid.java:9-14: (float)(0.001f + (.999f-0.001f) / (1f + Math.exp(-x))) */
__device__ __forceinline__ float sigmoid_clipped(float x) {
  const float c = __fsub_rn(.999f, 0.001f);
  return (float)((double)0.001f + (double)c / ((double)1.0f + exp((double)(-x))));
}
__device__ __forceinline__ float act_forward(int act, float z) {
  if (act == PS_ACT_RELU) return fmaxf(0.0f, z);
  if (act == PS_ACT_SIGMOID) return sigmoid_clipped(z);
  return z;
}
/* Relu.java:14-19 / Sigmoid.java:16-21: dy *= f'(y) expressed through the OUTPUT y */
__device__ __forceinline__ float act_backward(int act, float dy, float y) {
  if (act == PS_ACT_RELU) return __fmul_rn(dy, y > 0.0f ? 1.0f : 0.0f);
  if (act == PS_ACT_SIGMOID) return __fmul_rn(dy, __fmul_rn(y, __fsub_rn(1.0f, y)));
  return dy;
}
#endif

}  // namespace psb
