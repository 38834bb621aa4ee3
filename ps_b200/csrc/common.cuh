/*
 * common.cuh — error handling, the device context and small device helpers shared by
 * every translation unit of libps_b200.so (sm_100a only; there is no CPU fallback:
 * every entry point fails with PS_ERR_CUDA when no B200 is present).
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>

#include "../../include/ps_b200.h"
#include "../../include/ps_spec.h"

namespace psb {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string& s);

#define PS_CUDA(expr)                                                                        \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      char b__[512];                                                                         \
      snprintf(b__, sizeof b__, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      throw psb::Error(PS_ERR_CUDA, b__);                                                    \
    }                                                                                        \
  } while (0)

#define PS_REQUIRE(cond, code, msg)                                                          \
  do {                                                                                       \
    if (!(cond)) {                                                                           \
      char b__[512];                                                                         \
      snprintf(b__, sizeof b__, "%s (%s:%d)", msg, __FILE__, __LINE__);                      \
      throw psb::Error(code, b__);                                                           \
    }                                                                                        \
  } while (0)

#define PS_LAUNCH_CHECK() PS_CUDA(cudaGetLastError())

/* One per process and device (the reference's KVStore is a process-wide singleton,
 * store/KVStore.java:36,70; here the context owns what that singleton owned).       */
struct Ctx {
  int device = 0;
  int num_sms = 148;
  uint64_t seed = 0;
  cudaStream_t stream = nullptr;   /* compute stream: every kernel of a step is ordered on it */
  cudaStream_t copy_stream = nullptr; /* H2D staging for the pipelined trainer */
  int fc_precision = PS_FC_TF32X3;  /* FcLayer arithmetic: tcgen05 tensor cores, error-compensated 3xTF32 = fp32-grade (ps_ctx_set_fc_precision) */
  int prio_main = 0, prio_side = 0; /* stream priorities of the critical chain / the side branches */
  int pdl = 1;                     /* programmatic dependent launch for producer -> consumer kernel pairs (PS_PDL=0 disables) */
  int pdl_gemm = 0;                /* ... also along the FcLayer forward / dgrad GEMM chain: measured SLOWER at cfg2 (172.8 vs 164.6 us
                                      per step: early CTAs of the next GEMM take the 20 SMs the side-stream wgrad would use), so off
                                      unless PS_PDL_GEMM=1 */
  int p2p_defer = 2;               /* the sharded step leaves its owner-side embedding update and its dense update to the head of the NEXT sharded step,
                                      where they run beside that step's route_send (any other use of the model runs them first).  2 (default): only while
                                      the step is latency-bound (< 16 MB of embedding rows per rank and batch: 2 GPUs, cfg2 210.0 vs 212.6 us and e2e +5.7 %,
                                      cfg3 365 vs 372 us — but cfg4, 54 MB, 812 vs 792 us: there the update competes with the lookups for HBM);
                                      PS_P2P_DEFER=1 always, =0 never */
  int pdl_exchange = 0;            /* PS_PDL_EXCHANGE=1: the sharded step's exchange kernels (route_send -> owner lookup -> gather, scatter_rows -> owner
                                      update) and the backward's scatter are programmatic dependents of their predecessors */
  int exact_updaters = 0;          /* sparse update with the IEEE divisions / roots of the Java code (PS_EXACT_UPDATERS=1) instead of the fast forms */
  int hot_tma = 1;                 /* the lookup stages rows shared by >= 4 lookups of a warp task in shared memory by TMA bulk copies (PS_HOT_TMA=0: off) */
  int tc_deep = 2, tc_deep_wgrad = 2; /* 3xTF32 GEMM pipeline: 0 two stages x two CTAs per SM, 1 four stages x one CTA per SM, 2 the latter when the grid fits
                                      one wave of SMs (PS_TC_DEEP: forward / dgrad; PS_TC_DEEP_WGRAD: weight gradients) */
  int tc_wide_rule = 1;            /* PS_TC_WIDE_RULE=0: see dispatch_tc */
  int gemm_narrow = 0;             /* PS_GEMM_NARROW=1: 32-column 3xTF32 tiles when 64-column ones leave half the CTA slots empty */
  int group_wgrad = 0;             /* PS_GROUP_WGRAD=1: all weight-gradient GEMMs of a step in one grouped launch after the dgrad chain.  Measured
                                      SLOWER at cfg2 (164.9 vs 158.9 us per step: the 272-CTA launch delays the embedding update it runs beside), so off */
  int update_slab = 1;             /* the sparse update moves records by TMA bulk copies through shared memory (PS_UPDATE_SLAB=0: the register-path kernel) */
  int scatter_slab = 1;            /* the backward scatter stages a task's delta rows in shared memory (PS_SCATTER_SLAB=0: the register-path kernel) */
  int hot_share = 4;               /* ... PS_HOT_SHARE: how many lookups of a task must share the row (1 = every row goes through TMA) */
  unsigned hot_min = 8;            /* occurrences in a batch from which the scatter pre-sums a key per block (PS_HOT_MIN; 0 = never) */
  long launches = 0;               /* kernels launched by this library (bench gpu_launches) */
};

template <class T>
T* dmalloc(size_t n) {
  T* p = nullptr;
  if (n == 0) n = 1;
  PS_CUDA(cudaMalloc(&p, n * sizeof(T)));
  return p;
}
template <class T>
T* dmalloc_zero(size_t n, cudaStream_t s) {
  T* p = dmalloc<T>(n);
  PS_CUDA(cudaMemsetAsync(p, 0, (n ? n : 1) * sizeof(T), s));
  return p;
}
inline void dfree(void* p) { if (p) cudaFree(p); }

/* Launch `kernel` as a programmatic dependent of the kernel enqueued before it on ctx->stream: its blocks may be
 * scheduled once every block of that kernel has executed griddepcontrol.launch_dependents (or exited), and whatever it
 * does before its own griddepcontrol.wait overlaps the producer's tail.  Works inside stream capture (a programmatic
 * edge in the graph).  With ctx->pdl == 0 this is an ordinary launch and the device-side instructions are no-ops.    */
template <class... KArgs, class... Args>
inline void launch_pdl(Ctx* ctx, void (*kernel)(KArgs...), dim3 grid, dim3 block, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = ctx->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = ctx->pdl ? 1 : 0;
  PS_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}

/* the same with dynamic shared memory, and the attribute only when `dependent` */
template <class... KArgs, class... Args>
inline void launch_dep(Ctx* ctx, bool dependent, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = ctx->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = (dependent && ctx->pdl) ? 1 : 0;
  PS_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}

static inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }
static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

/* ---- device helpers ----------------------------------------------------------- */
#if defined(__CUDACC__)
__device__ __forceinline__ float4 ld_f4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st_f4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
/* streaming (read-once / write-once) 128-bit accesses that do not allocate in L1 */
__device__ __forceinline__ float4 ld_f4_stream(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_f4_stream(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
/* programmatic dependent launch (sm_90+): see launch_pdl */
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
/* fire-and-forget 32-bit reductions (no value comes back: the issuing thread does not wait for L2) */
__device__ __forceinline__ void red_add_u32(uint32_t* p, uint32_t v) { asm volatile("red.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void red_max_u32(uint32_t* p, uint32_t v) { asm volatile("red.global.max.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
/* 128-bit vector reduction into global memory (sm_90+): one L2 atomic transaction per 16 B */
__device__ __forceinline__ void red_add_f4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
#endif

}  // namespace psb
