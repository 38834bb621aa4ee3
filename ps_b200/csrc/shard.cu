/*
 * shard.cu — routing kernels of the key-hash sharded embedding table (see shard.cuh).
 * Integer / copy work only; owner = ps_owner_of(packed key, R) of include/ps_spec.h (a
 * net/Router.java:5 implementation; the reference's stock Mod router is restated bit-exactly in
 * ps_spec.h for the parity tests but is undefined for the ~29 % of keys whose hashCode is negative).
 */
#include "shard.cuh"

namespace psb {

__global__ void __launch_bounds__(256) shard_count_kernel(const int64_t* __restrict__ E, int L, int F, int R, int32_t* __restrict__ counts) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = l < L;
  int owner = -1 - (threadIdx.x & 31);
  if (valid) owner = (int)ps_owner_of(ps_pack_key((uint32_t)(l % F), (uint64_t)E[l]), (uint32_t)R);
  const unsigned peers = __match_any_sync(0xffffffffu, owner);
  if (valid && (__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&counts[owner], __popc(peers));
}

__global__ void __launch_bounds__(256) shard_place_kernel(const int64_t* __restrict__ E, int L, int F, int R, const int32_t* __restrict__ counts,
                                                          int32_t* __restrict__ cursor, unsigned long long* __restrict__ send_keys,
                                                          int32_t* __restrict__ send_pos) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool valid = l < L;
  unsigned long long key = 0;
  int owner = -1 - lane;
  if (valid) { key = ps_pack_key((uint32_t)(l % F), (uint64_t)E[l]); owner = (int)ps_owner_of(key, (uint32_t)R); }
  const unsigned peers = __match_any_sync(0xffffffffu, owner);
  const int leader = __ffs(peers) - 1;
  int base = 0;
  if (valid && lane == leader) base = atomicAdd(&cursor[owner], __popc(peers));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (!valid) return;
  int off = 0;
  for (int r = 0; r < owner; ++r) off += counts[r];
  const int pos = off + base + __popc(peers & ((1u << lane) - 1u));
  send_keys[pos] = key;
  send_pos[l] = pos;
}

/* fixed-capacity form: owner r's bucket is send_keys[r*cap, (r+1)*cap), unused entries stay EMPTY (the
 * caller zero-fills); no counts have to reach the host, so the whole exchange can live in a CUDA graph.
 * A bucket that would exceed cap raises *overflow (the host then falls back to the exact-size path). */
__global__ void __launch_bounds__(256) shard_place_padded_kernel(const int64_t* __restrict__ E, int L, int F, int R, int cap, int32_t* __restrict__ cursor,
                                                                 unsigned long long* __restrict__ send_keys, int32_t* __restrict__ send_pos,
                                                                 int32_t* __restrict__ overflow) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool valid = l < L;
  unsigned long long key = 0;
  int owner = -1 - lane;
  if (valid) { key = ps_pack_key((uint32_t)(l % F), (uint64_t)E[l]); owner = (int)ps_owner_of(key, (uint32_t)R); }
  const unsigned peers = __match_any_sync(0xffffffffu, owner);
  const int leader = __ffs(peers) - 1;
  int base = 0;
  if (valid && lane == leader) base = atomicAdd(&cursor[owner], __popc(peers));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (!valid) return;
  const int pos = base + __popc(peers & ((1u << lane) - 1u));
  if (pos < cap) { send_keys[(size_t)owner * cap + pos] = key; send_pos[l] = owner * cap + pos; }
  else { send_pos[l] = -1; *overflow = 1; }
}

__global__ void __launch_bounds__(256) shard_unpack_kernel(const float* __restrict__ rows, const int32_t* __restrict__ send_pos, int L, int F, int D, int Dp,
                                                           float* __restrict__ out, int ldo) {
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long l = g / Dp;
  const int d = (int)(g - l * Dp);
  if (l >= L || d >= D) return;
  const int n = (int)(l / F), j = (int)(l - (long)n * F);
  const int pos = send_pos[l];
  out[(size_t)n * ldo + j * D + d] = pos >= 0 ? rows[(size_t)pos * Dp + d] : 0.f;
}

__global__ void __launch_bounds__(256) shard_pack_grads_kernel(const float* __restrict__ delta, int ldd, const float* __restrict__ act, int lda,
                                                               const int32_t* __restrict__ send_pos, int L, int F, int D, int Dp,
                                                               float* __restrict__ grads) {
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long l = g / Dp;
  const int d = (int)(g - l * Dp);
  if (l >= L) return;
  const int n = (int)(l / F), j = (int)(l - (long)n * F);
  float v = 0.f;
  if (d < D) v = __fmul_rn(delta[(size_t)n * ldd + j * D + d], act[(size_t)n * lda + j * D + d] > 0.f ? 1.f : 0.f);
  const int pos = send_pos[l];
  if (pos >= 0) grads[(size_t)pos * Dp + d] = v;
}

void shard_count(Ctx* ctx, const int64_t* E, int N, int F, int R, int32_t* counts) {
  const int L = N * F;
  shard_count_kernel<<<ceil_div(L, 256), 256, 0, ctx->stream>>>(E, L, F, R, counts);
  PS_LAUNCH_CHECK();
  ctx->launches++;
}
void shard_place(Ctx* ctx, const int64_t* E, int N, int F, int R, const int32_t* counts, int32_t* cursor, uint64_t* send_keys, int32_t* send_pos) {
  const int L = N * F;
  shard_place_kernel<<<ceil_div(L, 256), 256, 0, ctx->stream>>>(E, L, F, R, counts, cursor, reinterpret_cast<unsigned long long*>(send_keys), send_pos);
  PS_LAUNCH_CHECK();
  ctx->launches++;
}
void shard_place_padded(Ctx* ctx, const int64_t* E, int N, int F, int R, int cap, int32_t* cursor, uint64_t* send_keys, int32_t* send_pos,
                        int32_t* overflow) {
  const int L = N * F;
  PS_CUDA(cudaMemsetAsync(send_keys, 0, sizeof(uint64_t) * (size_t)R * cap, ctx->stream));
  PS_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int32_t) * R, ctx->stream));
  shard_place_padded_kernel<<<ceil_div(L, 256), 256, 0, ctx->stream>>>(E, L, F, R, cap, cursor, reinterpret_cast<unsigned long long*>(send_keys), send_pos, overflow);
  PS_LAUNCH_CHECK();
  ctx->launches++;
}
void shard_unpack(Ctx* ctx, const float* rows, const int32_t* send_pos, int N, int F, int D, int Dp, float* out, int ldo) {
  const long T = (long)N * F * Dp;
  shard_unpack_kernel<<<ceil_div(T, 256), 256, 0, ctx->stream>>>(rows, send_pos, N * F, F, D, Dp, out, ldo);
  PS_LAUNCH_CHECK();
  ctx->launches++;
}
void shard_pack_grads(Ctx* ctx, const float* delta, int ldd, const float* act, int lda, const int32_t* send_pos, int N, int F, int D, int Dp,
                      float* grads) {
  const long T = (long)N * F * Dp;
  shard_pack_grads_kernel<<<ceil_div(T, 256), 256, 0, ctx->stream>>>(delta, ldd, act, lda, send_pos, N * F, F, D, Dp, grads);
  PS_LAUNCH_CHECK();
  ctx->launches++;
}

}  // namespace psb
