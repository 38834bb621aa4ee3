/*
 * updaters.cuh — per-element device forms of update.AdamUpdater / FtrlUpdater / SimpleUpdater.
 * The operation ORDER and rounding of the reference are kept (every step is a single
 * correctly-rounded fp32 operation: __f*_rn intrinsics are never contracted into FMAs), so
 * given the same gradient bits the new weight and state bits equal the Java/jblas ones.
 */
#pragma once
#include "common.cuh"

namespace psb {

struct UpdaterDev {
  int kind;
  float a, b, c, d;      /* adam: alfa, beta1, beta2, epsilon | ftrl: alfa, beta, l1, l2 | simple: eta */
  float omb1, omb2;      /* adam: (1 - beta1), (1 - beta2) evaluated in float like the Java code */
  float r_omb1, r_omb2, r_a;   /* fast mode: 1/(1-beta1), 1/(1-beta2), 1/alfa rounded once */
};

inline UpdaterDev make_updater_dev(const ps_updater_spec& s) {
  UpdaterDev u;
  u.kind = s.kind; u.a = s.p[0]; u.b = s.p[1]; u.c = s.p[2]; u.d = s.p[3];
  u.omb1 = 1.0f - u.b; u.omb2 = 1.0f - u.c;
  u.r_omb1 = u.omb1 != 0.f ? 1.0f / u.omb1 : 0.f; u.r_omb2 = u.omb2 != 0.f ? 1.0f / u.omb2 : 0.f; u.r_a = u.a != 0.f ? 1.0f / u.a : 0.f;
  return u;
}

#if defined(__CUDACC__)
/* IEEE division / square root with the trivial cases peeled off: a zero numerator (a ReLU-masked
 * gradient element, an untouched moment) otherwise sends __fdiv_rn down its slow path, which
 * dominated the instruction count of the update kernels.  b is positive wherever these are used,
 * so 0 / b == 0 with the sign of the numerator, exactly what IEEE division returns.          */
__device__ __forceinline__ float pdiv(float a, float b) { return a == 0.0f ? a : __fdiv_rn(a, b); }
__device__ __forceinline__ float psqrt(float a) { return a == 0.0f ? a : __fsqrt_rn(a); }

/* update/AdamUpdater.java:61-69.  m,v are the M/V maps' entries (zero until first touch). */
__device__ __forceinline__ void adam_elem(const UpdaterDev& u, float& w, float& m, float& v, float g) {
  const float m_new = __fadd_rn(__fmul_rn(g, u.omb1), __fmul_rn(m, u.b));                 /* :61 */
  const float v_new = __fadd_rn(__fmul_rn(__fmul_rn(g, g), u.omb2), __fmul_rn(v, u.c));   /* :62 */
  const float Mm = pdiv(m_new, u.omb1);                                                    /* :63 */
  const float Vv = pdiv(v_new, u.omb2);                                                    /* :64 */
  const float den = __fadd_rn(psqrt(Vv), u.d);
  const float stp = __fmul_rn(pdiv(Mm, den), -u.a);                                        /* :69 */
  w = __fadd_rn(w, stp);
  m = m_new; v = v_new;
}

/* update/FtrlUpdater.java:64-74 (the early return of :52 is decided by the caller on g[0]). */
__device__ __forceinline__ void ftrl_elem(const UpdaterDev& u, float& w, float& z, float& n, float g) {
  float wn;
  if (fabsf(z) <= u.c) {
    wn = 0.0f;
  } else {
    const float sign = z >= 0.0f ? 1.0f : -1.0f;
    const float num = -__fsub_rn(z, __fmul_rn(sign, u.c));
    const float den = __fdiv_rn(__fadd_rn(u.d, __fadd_rn(u.b, psqrt(n))), u.a);
    wn = __fdiv_rn(num, den);
  }
  const float g2 = __fmul_rn(g, g);
  const float s = __fsub_rn(psqrt(__fadd_rn(n, g2)), psqrt(pdiv(n, u.a)));                  /* :72, sic */
  z = __fadd_rn(z, __fsub_rn(g, __fmul_rn(s, wn)));                                         /* :73 */
  n = __fadd_rn(n, g2);                                                                     /* :74 */
  w = wn;
}

/* update/SimpleUpdater.java:20-22 */
__device__ __forceinline__ void simple_elem(const UpdaterDev& u, float& w, float g) {
  w = __fadd_rn(w, __fmul_rn(g, -u.a));
}

/* ---- fast mode (embedding rows only; the default of the sparse update kernel) -------------------------------------------
 * The exact forms above spend ~150 instructions per element, almost all of them in the 4-5 IEEE divisions and the square
 * root, which made the sparse update issue-bound instead of HBM-bound.  Here divisions by constants become multiplications
 * by their reciprocals, the remaining quotient is div.approx and the root sqrt.approx (each <= 2 ulp): the updated weight
 * agrees with the exact form to a few ulp of the STEP (|step| <= alfa), i.e. <= 1e-6 relative on the row — SURVEY
 * Appendix B's bound for post-update rows; the scatter's atomics already reorder the gradient sums by as much.
 * PS_EXACT_UPDATERS=1 / ps_ctx_set_exact_updaters selects the exact forms for the sparse update as well.                */
__device__ __forceinline__ float sqrt_approx(float a) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}
__device__ __forceinline__ void adam_elem_fast(const UpdaterDev& u, float& w, float& m, float& v, float g) {
  const float m_new = __fadd_rn(__fmul_rn(g, u.omb1), __fmul_rn(m, u.b));
  const float v_new = __fadd_rn(__fmul_rn(__fmul_rn(g, g), u.omb2), __fmul_rn(v, u.c));
  const float Mm = __fmul_rn(m_new, u.r_omb1);
  const float Vv = __fmul_rn(v_new, u.r_omb2);
  const float den = __fadd_rn(sqrt_approx(Vv), u.d);
  w = __fadd_rn(w, __fmul_rn(__fdividef(Mm, den), -u.a));
  m = m_new; v = v_new;
}
__device__ __forceinline__ void ftrl_elem_fast(const UpdaterDev& u, float& w, float& z, float& n, float g) {
  float wn;
  if (fabsf(z) <= u.c) {
    wn = 0.0f;
  } else {
    const float sign = z >= 0.0f ? 1.0f : -1.0f;
    const float num = -__fsub_rn(z, __fmul_rn(sign, u.c));
    const float den = __fmul_rn(__fadd_rn(u.d, __fadd_rn(u.b, sqrt_approx(n))), u.r_a);
    wn = __fdividef(num, den);
  }
  const float g2 = __fmul_rn(g, g);
  const float s = __fsub_rn(sqrt_approx(__fadd_rn(n, g2)), sqrt_approx(__fmul_rn(n, u.r_a)));   /* :72, sic */
  z = __fadd_rn(z, __fsub_rn(g, __fmul_rn(s, wn)));
  n = __fadd_rn(n, g2);
  w = wn;
}

template <bool EXACT>
__device__ __forceinline__ void apply_elem(const UpdaterDev& u, float& w, float& s1, float& s2, float g) {
  if (u.kind == PS_UPD_ADAM) { if (EXACT) adam_elem(u, w, s1, s2, g); else adam_elem_fast(u, w, s1, s2, g); }
  else if (u.kind == PS_UPD_FTRL) { if (EXACT) ftrl_elem(u, w, s1, s2, g); else ftrl_elem_fast(u, w, s1, s2, g); }
  else simple_elem(u, w, g);
}

__device__ __forceinline__ void apply_elem(const UpdaterDev& u, float& w, float& s1, float& s2, float g) {
  if (u.kind == PS_UPD_ADAM) adam_elem(u, w, s1, s2, g);
  else if (u.kind == PS_UPD_FTRL) ftrl_elem(u, w, s1, s2, g);
  else simple_elem(u, w, g);
}
#endif

}  // namespace psb
