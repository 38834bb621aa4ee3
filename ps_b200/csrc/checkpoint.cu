/*
 * checkpoint.cu — dump / load of the GPU-resident store.
 *
 * The reference keeps every parameter in store.KVStore's maps (KVStore.java:38-44) and has NO checkpoint; PServer holds
 * the same maps for dist mode (PServer.java:102-162 getList / upsertList are its only bulk accessors).  This is the bulk
 * form of those accessors: every key the reference's store would hold — "fc<i>.weights/bias", "wide.bias",
 * "wide.weights.<id>", "emF<j>.<id>" — with the updater state beside it (AdamUpdater.M/V, FtrlUpdater.Z/N maps,
 * AdamUpdater.java:27-28, FtrlUpdater.java:26-27), written to one file and read back into a model of the same shape.
 * Embedding rows are exported in slot-range chunks through a device-side compaction (occupied slots only), so a table
 * sized for 180 GB of HBM streams through a bounded staging buffer; loading re-inserts by key, so capacities may differ.
 */
#include <cstdio>
#include <memory>
#include <vector>

#include "model.cuh"

namespace psb {

namespace {

constexpr uint64_t kMagic = 0x3130303242535000ull;   /* "\0PSB2001" */

struct FileCloser { void operator()(FILE* f) const { if (f) std::fclose(f); } };
using File = std::unique_ptr<FILE, FileCloser>;
struct DevFree { void operator()(void* p) const { dfree(p); } };   /* device staging buffers are released on every exit path */
template <class T> using DevBuf = std::unique_ptr<T, DevFree>;

void wr(FILE* f, const void* p, size_t n) { PS_REQUIRE(std::fwrite(p, 1, n, f) == n, PS_ERR_ARG, "checkpoint: short write"); }
void rd(FILE* f, void* p, size_t n) { PS_REQUIRE(std::fread(p, 1, n, f) == n, PS_ERR_ARG, "checkpoint: short read (truncated or foreign file)"); }

void wr_dev(Ctx* ctx, FILE* f, const void* dev, size_t bytes, std::vector<char>& tmp) {
  tmp.resize(bytes);
  PS_CUDA(cudaMemcpyAsync(tmp.data(), dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  PS_CUDA(cudaStreamSynchronize(ctx->stream));
  wr(f, tmp.data(), bytes);
}
void rd_dev(Ctx* ctx, FILE* f, void* dev, size_t bytes, std::vector<char>& tmp) {
  tmp.resize(bytes);
  rd(f, tmp.data(), bytes);
  PS_CUDA(cudaMemcpyAsync(dev, tmp.data(), bytes, cudaMemcpyHostToDevice, ctx->stream));
  PS_CUDA(cudaStreamSynchronize(ctx->stream));
}

/* occupied slots of [s0, s1) -> compact records: keys[i], rows[which][i][D] */
__global__ void __launch_bounds__(256) emb_export_kernel(const EmbSlot* __restrict__ slots, uint32_t s0, uint32_t s1, const float* __restrict__ w,
                                                         const float* __restrict__ a, const float* __restrict__ b, int Dp, int D,
                                                         unsigned long long* __restrict__ keys, float* __restrict__ rows, uint32_t cap,
                                                         uint32_t* __restrict__ counter) {
  const uint32_t s = s0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= s1) return;
  const unsigned long long k = slots[s].key;
  if (k == PS_KEY_EMPTY) return;
  const uint32_t i = atomicAdd(counter, 1u);
  keys[i] = k;
  for (int d = 0; d < D; ++d) {
    rows[((size_t)0 * cap + i) * D + d] = w[(size_t)s * Dp + d];
    rows[((size_t)1 * cap + i) * D + d] = a[(size_t)s * Dp + d];
    rows[((size_t)2 * cap + i) * D + d] = b[(size_t)s * Dp + d];
  }
}

__global__ void __launch_bounds__(256) emb_import_kernel(EmbSlot* __restrict__ slots, uint32_t C, float* __restrict__ w, float* __restrict__ a,
                                                         float* __restrict__ b, int Dp, int D, const unsigned long long* __restrict__ keys,
                                                         const float* __restrict__ rows, uint32_t n, uint32_t cap, uint32_t* __restrict__ counters) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long key = keys[i];
  uint32_t slot = ps_bucket_of(key, C);
  int found = -1;
  const uint32_t limit = C < (uint32_t)kProbeLimit ? C : (uint32_t)kProbeLimit;   /* the bound the runtime lookups stop at: a key placed beyond it would never be found again */
  for (uint32_t p = 0; p < limit; ++p) {
    const unsigned long long k = *reinterpret_cast<const volatile unsigned long long*>(&slots[slot].key);
    if (k == key) { found = (int)slot; break; }
    if (k == PS_KEY_EMPTY) {
      const unsigned long long old = atomicCAS(&slots[slot].key, (unsigned long long)PS_KEY_EMPTY, key);
      if (old == PS_KEY_EMPTY) { atomicAdd(reinterpret_cast<unsigned long long*>(counters + 2), 1ull); found = (int)slot; break; }
      if (old == key) { found = (int)slot; break; }
    }
    slot = slot + 1 == C ? 0 : slot + 1;
  }
  if (found < 0) { atomicOr(&counters[1], 1u); return; }
  for (int d = 0; d < D; ++d) {
    w[(size_t)found * Dp + d] = rows[((size_t)0 * cap + i) * D + d];
    a[(size_t)found * Dp + d] = rows[((size_t)1 * cap + i) * D + d];
    b[(size_t)found * Dp + d] = rows[((size_t)2 * cap + i) * D + d];
  }
  __threadfence();
  atomicOr(&slots[found].uidx, kRowReady);
}

struct Header {
  uint64_t magic;
  int32_t kind, F, D, Xn, L, has_wide;
  int32_t dims[kMaxDenseLayers + 1];
  int64_t emb_rows, wide_capacity;
};

}  // namespace

void Model::save(const std::string& path) {
  flush_deferred();                              /* a sharded step may have left its update to "the next step" */
  PS_REQUIRE(in_flight == 0, PS_ERR_STATE, "save: steps in flight");
  PS_CUDA(cudaStreamSynchronize(ctx->stream));
  /* written beside the target and renamed into place once every byte is on disk: a failed dump never leaves a truncated checkpoint */
  const std::string tmp_path = path + ".tmp";
  File f(std::fopen(tmp_path.c_str(), "wb"));
  PS_REQUIRE(f != nullptr, PS_ERR_ARG, ("save: cannot open " + tmp_path).c_str());
  Header h{};
  h.magic = kMagic; h.kind = kind; h.F = F; h.D = D; h.Xn = Xn; h.L = L; h.has_wide = has_wide ? 1 : 0;
  for (int l = 0; l <= L; ++l) h.dims[l] = width[l];
  h.emb_rows = has_emb ? emb.size() : 0;
  h.wide_capacity = has_wide ? wide.C : 0;
  wr(f.get(), &h, sizeof h);
  std::vector<char> tmp;
  for (int l = 0; l < L; ++l) {                 /* "fc<l>.weights" / "fc<l>.bias" with their updater state */
    const FcLayer& q = fcs[l];
    const size_t wb = sizeof(float) * (size_t)q.out * q.ldw;
    wr_dev(ctx, f.get(), q.W, wb, tmp); wr_dev(ctx, f.get(), q.sW1, wb, tmp); wr_dev(ctx, f.get(), q.sW2, wb, tmp);
    wr_dev(ctx, f.get(), q.bias, sizeof(float) * q.out, tmp); wr_dev(ctx, f.get(), q.sb1, sizeof(float) * q.out, tmp); wr_dev(ctx, f.get(), q.sb2, sizeof(float) * q.out, tmp);
  }
  if (has_wide) {                               /* "wide.bias" {w, s1, s2} and every "wide.weights.<id>" record */
    wr_dev(ctx, f.get(), wide_bias, sizeof(float) * 4, tmp);
    wr_dev(ctx, f.get(), wide.slots, sizeof(WideSlot) * (size_t)wide.C, tmp);
    wr_dev(ctx, f.get(), wide.counters, sizeof(uint32_t) * 4, tmp);
  }
  if (has_emb) {                                /* "emF<j>.<id>": occupied slots only, in slot-range chunks */
    const uint32_t chunk = (uint32_t)std::min<int64_t>(emb.C, 1 << 20);
    DevBuf<unsigned long long> g_keys(dmalloc<unsigned long long>(chunk));
    DevBuf<float> g_rows(dmalloc<float>((size_t)3 * chunk * D));
    DevBuf<uint32_t> g_cnt(dmalloc_zero<uint32_t>(1, ctx->stream));
    unsigned long long* d_keys = g_keys.get(); float* d_rows = g_rows.get(); uint32_t* d_cnt = g_cnt.get();
    std::vector<unsigned long long> hk(chunk);
    std::vector<float> hr((size_t)3 * chunk * D);
    int64_t written = 0;
    for (int64_t s0 = 0; s0 < emb.C; s0 += chunk) {
      const uint32_t s1 = (uint32_t)std::min<int64_t>(emb.C, s0 + chunk);
      PS_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(uint32_t), ctx->stream));
      emb_export_kernel<<<ceil_div(s1 - (uint32_t)s0, 256), 256, 0, ctx->stream>>>(emb.slots, (uint32_t)s0, s1, emb.w, emb.s1, emb.s2, emb.rs, D, d_keys, d_rows,
                                                                                  chunk, d_cnt);
      PS_LAUNCH_CHECK();
      ctx->launches++;
      uint32_t n = 0;
      PS_CUDA(cudaMemcpyAsync(&n, d_cnt, sizeof n, cudaMemcpyDeviceToHost, ctx->stream));
      PS_CUDA(cudaStreamSynchronize(ctx->stream));
      if (n == 0) continue;
      PS_CUDA(cudaMemcpyAsync(hk.data(), d_keys, sizeof(unsigned long long) * n, cudaMemcpyDeviceToHost, ctx->stream));
      for (int which = 0; which < 3; ++which)
        PS_CUDA(cudaMemcpyAsync(hr.data() + (size_t)which * n * D, d_rows + (size_t)which * chunk * D, sizeof(float) * (size_t)n * D, cudaMemcpyDeviceToHost, ctx->stream));
      PS_CUDA(cudaStreamSynchronize(ctx->stream));
      wr(f.get(), &n, sizeof n);
      wr(f.get(), hk.data(), sizeof(unsigned long long) * n);
      wr(f.get(), hr.data(), sizeof(float) * (size_t)3 * n * D);
      written += n;
    }
    const uint32_t end = 0;
    wr(f.get(), &end, sizeof end);
    PS_REQUIRE(written == h.emb_rows, PS_ERR_STATE, "save: embedding row count changed during the dump");
  }
  PS_REQUIRE(std::fflush(f.get()) == 0, PS_ERR_ARG, "save: flush failed (disk full?)");
  FILE* raw = f.release();
  PS_REQUIRE(std::fclose(raw) == 0, PS_ERR_ARG, "save: close failed (disk full?)");
  PS_REQUIRE(std::rename(tmp_path.c_str(), path.c_str()) == 0, PS_ERR_ARG, ("save: cannot rename into " + path).c_str());
}

void Model::load(const std::string& path) {
  flush_deferred();                              /* a sharded step may have left its update to "the next step" */
  PS_REQUIRE(in_flight == 0, PS_ERR_STATE, "load: steps in flight");
  File f(std::fopen(path.c_str(), "rb"));
  PS_REQUIRE(f != nullptr, PS_NOT_FOUND, ("load: cannot open " + path).c_str());
  Header h{};
  rd(f.get(), &h, sizeof h);
  PS_REQUIRE(h.magic == kMagic, PS_ERR_ARG, "load: not a ps-b200 checkpoint");
  bool same = h.kind == kind && h.F == F && h.D == D && h.Xn == Xn && h.L == L && (h.has_wide != 0) == has_wide;
  for (int l = 0; same && l <= L; ++l) same = h.dims[l] == width[l];
  PS_REQUIRE(same, PS_ERR_ARG, "load: the checkpoint was written by a model of a different shape");
  PS_REQUIRE(!has_emb || emb.size() == 0, PS_ERR_STATE, "load: the embedding table must be empty (load into a fresh model)");
  std::vector<char> tmp;
  for (int l = 0; l < L; ++l) {
    FcLayer& q = fcs[l];
    const size_t wb = sizeof(float) * (size_t)q.out * q.ldw;
    rd_dev(ctx, f.get(), q.W, wb, tmp); rd_dev(ctx, f.get(), q.sW1, wb, tmp); rd_dev(ctx, f.get(), q.sW2, wb, tmp);
    rd_dev(ctx, f.get(), q.bias, sizeof(float) * q.out, tmp); rd_dev(ctx, f.get(), q.sb1, sizeof(float) * q.out, tmp); rd_dev(ctx, f.get(), q.sb2, sizeof(float) * q.out, tmp);
    transpose_copy(ctx, q.W, q.ldw, q.Wt, q.ldwt, q.out, q.in);       /* the K-major copy the dgrad GEMM reads */
    q.refresh_lo(ctx);
  }
  if (has_wide) {
    PS_REQUIRE(h.wide_capacity == wide.C, PS_ERR_ARG, "load: wide table capacity mismatch");
    rd_dev(ctx, f.get(), wide_bias, sizeof(float) * 4, tmp);
    rd_dev(ctx, f.get(), wide.slots, sizeof(WideSlot) * (size_t)wide.C, tmp);
    rd_dev(ctx, f.get(), wide.counters, sizeof(uint32_t) * 4, tmp);
  }
  if (has_emb) {
    PS_REQUIRE(h.emb_rows <= emb.C, PS_ERR_CAPACITY, "load: more embedding rows than this table's capacity");
    const uint32_t chunk = 1 << 20;
    DevBuf<unsigned long long> g_keys(dmalloc<unsigned long long>(chunk));
    DevBuf<float> g_rows(dmalloc<float>((size_t)3 * chunk * D));
    unsigned long long* d_keys = g_keys.get(); float* d_rows = g_rows.get();
    std::vector<unsigned long long> hk;
    std::vector<float> hr;
    while (true) {
      uint32_t n = 0;
      rd(f.get(), &n, sizeof n);
      if (n == 0) break;
      PS_REQUIRE(n <= chunk, PS_ERR_ARG, "load: corrupt chunk header");
      hk.resize(n); hr.resize((size_t)3 * n * D);
      rd(f.get(), hk.data(), sizeof(unsigned long long) * n);
      rd(f.get(), hr.data(), sizeof(float) * (size_t)3 * n * D);
      PS_CUDA(cudaMemcpyAsync(d_keys, hk.data(), sizeof(unsigned long long) * n, cudaMemcpyHostToDevice, ctx->stream));
      for (int which = 0; which < 3; ++which)
        PS_CUDA(cudaMemcpyAsync(d_rows + (size_t)which * chunk * D, hr.data() + (size_t)which * n * D, sizeof(float) * (size_t)n * D, cudaMemcpyHostToDevice, ctx->stream));
      emb_import_kernel<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(emb.slots, (uint32_t)emb.C, emb.w, emb.s1, emb.s2, emb.rs, D, d_keys, d_rows, n, chunk, emb.counters);
      PS_LAUNCH_CHECK();
      ctx->launches++;
      PS_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    emb.check_errors();
    PS_REQUIRE(emb.size() == h.emb_rows, PS_ERR_STATE, "load: embedding row count does not match the header");
  }
  PS_CUDA(cudaStreamSynchronize(ctx->stream));
}

}  // namespace psb
