/*
 * shard.cuh — device side of the key-hash sharded table: what
 * net/PSRouterClient.java:60-122 does with thread pools and gRPC — bucket keys by
 * router.shard(key), one batched getList / updateList per shard, merge — done as pack / unpack
 * kernels around an all-to-all over NVLink.  The collective itself is issued by the host
 * orchestrator (ps_b200/sharded.py, torch.distributed / NCCL).
 */
#pragma once
#include "common.cuh"

namespace psb {

/* counts[r] = number of the batch's L = N*F lookups owned by rank r (ps_owner_of); counts must be zeroed */
void shard_count(Ctx* ctx, const int64_t* E, int N, int F, int R, int32_t* counts);
/* send_keys: packed keys grouped by owner (group r starts at sum(counts[:r])); send_pos[l] = index of
 * lookup l in that buffer; cursor: R zeroed ints of scratch                                           */
void shard_place(Ctx* ctx, const int64_t* E, int N, int F, int R, const int32_t* counts, int32_t* cursor, uint64_t* send_keys, int32_t* send_pos);
/* fixed-capacity routing (graph-capturable: no count reaches the host): bucket r = send_keys[r*cap, (r+1)*cap),
 * unused entries EMPTY; send_pos[l] = r*cap + position or -1 on overflow (then *overflow = 1)      */
void shard_place_padded(Ctx* ctx, const int64_t* E, int N, int F, int R, int cap, int32_t* cursor, uint64_t* send_keys, int32_t* send_pos,
                        int32_t* overflow);
/* out[n*ldo + j*D + d] = rows[send_pos[n*F+j]*Dp + d]  (rows come back already ReLU'd by their owner) */
void shard_unpack(Ctx* ctx, const float* rows, const int32_t* send_pos, int N, int F, int D, int Dp, float* out, int ldo);
/* grads[send_pos[l]*Dp + d] = delta[n*ldd + j*D + d] * (act[n*lda + j*D + d] > 0)   (EmbeddingField.java:91-93) */
void shard_pack_grads(Ctx* ctx, const float* delta, int ldd, const float* act, int lda, const int32_t* send_pos, int N, int F, int D, int Dp,
                      float* grads);

}  // namespace psb
