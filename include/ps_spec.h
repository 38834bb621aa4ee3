/*
 * ps_spec.h — the bit-exact, implementation-independent definitions that the CUDA
 * product (ps_b200/csrc) and the CPU oracle (oracle/) must agree on.  Only things
 * the reference leaves UNDEFINED or expresses through Java strings live here:
 *
 *   - the 64-bit packed form of a reference key string ("emF<j>.<id>.0",
 *     "wide.weights.<id>.0", "fc<i>.weights" ...),
 *   - the deterministic replacement for the reference's UNSEEDED initialiser
 *     (util/MatrixUtil.java:62-74 uses commons-lang3 RandomUtils with no seed;
 *     distribution kept: sign chosen 50/50, magnitude U[0, max)),
 *   - the production key->owner hash (the reference accepts any net/Router.java:5;
 *     its stock router net/Mod.java:13-15 is restated separately, bit-exact, as
 *     ps_java_string_hash / ps_router_mod_java).
 *
 * Everything is integer arithmetic plus ONE int->float conversion and ONE fp32
 * multiply, so host (gcc, -ffp-contract=off) and device (nvcc) produce identical bits.
 */
#ifndef PS_SPEC_H_
#define PS_SPEC_H_

#include <stdint.h>

#if defined(__CUDACC__)
#define PS_HD __host__ __device__ __forceinline__
#else
#define PS_HD static inline
#endif

/* ---- key packing --------------------------------------------------------------
 * Reference keys are strings: name + "." + String.valueOf(float id)
 * (layer/EmbeddingField.java:70-71, layer/LRLayer.java:78).  At the C ABI a key is
 * (namespace, id): namespace j in [0, 2^19) is embedding field "emF<j>", id is the
 * integer the float encoded (exact below 2^24 in the reference; int64 here).
 * Packed key: bits [63:44] = namespace+1 (never 0, so 0 is the EMPTY sentinel),
 * bits [43:0] = id.                                                              */
#define PS_KEY_ID_BITS 44
#define PS_KEY_ID_MASK ((1ull << PS_KEY_ID_BITS) - 1ull)
#define PS_KEY_EMPTY 0ull

PS_HD uint64_t ps_pack_key(uint32_t ns, uint64_t id) {
  return ((uint64_t)(ns + 1u) << PS_KEY_ID_BITS) | (id & PS_KEY_ID_MASK);
}
PS_HD uint32_t ps_key_ns(uint64_t key) { return (uint32_t)(key >> PS_KEY_ID_BITS) - 1u; }
PS_HD uint64_t ps_key_id(uint64_t key) { return key & PS_KEY_ID_MASK; }

/* splitmix64 finaliser: the only hash primitive used by the product. */
PS_HD uint64_t ps_mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}

/* Production router (a net/Router.java:5 implementation): owner = hash(key) mod n.
 * Uses the HIGH 32 bits of the mix so it is independent of the bucket index below. */
PS_HD uint32_t ps_owner_of(uint64_t key, uint32_t n_shards) {
  uint32_t h = (uint32_t)(ps_mix64(key ^ 0x5851f42d4c957f2dull) >> 32);
  return (uint32_t)(((uint64_t)h * (uint64_t)n_shards) >> 32);
}

/* Bucket index inside one shard's table (Lemire range reduction, any bucket count). */
PS_HD uint32_t ps_bucket_of(uint64_t key, uint32_t n_buckets) {
  uint32_t h = (uint32_t)ps_mix64(key);
  return (uint32_t)(((uint64_t)h * (uint64_t)n_buckets) >> 32);
}

/* ---- deterministic initialiser -------------------------------------------------
 * Element j of the parameter stored under `key`, drawn as the reference draws it
 * (MatrixUtil.rand(row,col,max): RandomUtils.nextInt(0,2)==0 ? +U[0,max) : -U[0,max))
 * but from a counter-based hash of (seed, key, j) instead of an unseeded RNG.     */
PS_HD float ps_init_value(uint64_t seed, uint64_t key, uint32_t j, float maxv) {
  uint64_t h = ps_mix64(seed + 0x9e3779b97f4a7c15ull * (uint64_t)(j + 1u) + ps_mix64(key));
  float u = (float)(uint32_t)(h >> 40) * (1.0f / 16777216.0f); /* 24 bits -> [0,1), exact */
  float v = u * maxv;                                         /* one rounding          */
  return (h & 1ull) ? -v : v;
}

/* 64-bit FNV-1a of a parameter name; top bit set so dense keys never collide with
 * packed embedding keys (whose namespace field stays below 2^19).                 */
PS_HD uint64_t ps_name_key(const char* s) {
  uint64_t h = 0xcbf29ce484222325ull;
  for (; *s; ++s) { h ^= (uint64_t)(unsigned char)*s; h *= 0x100000001b3ull; }
  return h | 0x8000000000000000ull;
}

/* java.lang.String.hashCode over an ASCII string: s[0]*31^(n-1) + ... in int32
 * wraparound (used by net/Mod.java:14).                                           */
PS_HD int32_t ps_java_string_hash(const char* s) {
  uint32_t h = 0;
  for (; *s; ++s) h = 31u * h + (uint32_t)(unsigned char)*s;
  return (int32_t)h;
}

/* net/Mod.java:13-15 as written: Java `%` truncates toward zero, so a negative
 * hashCode gives a NEGATIVE shard (the reference then throws in clients.get()).   */
PS_HD int32_t ps_router_mod_java(const char* key, int32_t n) {
  return ps_java_string_hash(key) % n;
}
/* The defined replacement used everywhere else: Math.floorMod.                    */
PS_HD int32_t ps_router_floormod_java(const char* key, int32_t n) {
  int32_t r = ps_java_string_hash(key) % n;
  return r < 0 ? r + n : r;
}

/* Xavier bound the reference uses for every initialiser:
 * (float)(4 * (Math.sqrt(6) / Math.sqrt(fan_in + fan_out)))
 * (layer/EmbeddingField.java:40, layer/FcLayer.java:39,46).  Evaluated on the HOST
 * only (double sqrt), passed to kernels as a float.                               */

#endif /* PS_SPEC_H_ */
