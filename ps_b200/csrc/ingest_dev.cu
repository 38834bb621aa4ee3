/*
 * ingest_dev.cu — libsvm text parsed ON the GPU (the B200 form of the ingest).
 *
 * The host reader (ingest.cu) turns ~0.6 M lines/s per CPU thread into staging buffers; the training step consumes 25 M
 * samples/s.  Raw text is only ~700 B per line, so PCIe can carry it at step speed: copy the text, parse it here.
 *
 * Same semantics as data/LibsvmParser.java:13-25 + CTR.parseFeature CTR.java:47-68 for the line spellings the host
 * parser's fast path accepts (label and values: [-]digits[.digits] with <= 7 significant digits and <= 10 decimals;
 * indices: <= 18 digits) — bit-identical results there (the value is ONE correctly rounded fp32 division m / 10^k on both
 * sides).  Any other spelling (exponents, suffixes, hex floats, NaN/Infinity, signs on indices, inner empty tokens, ...)
 * is not guessed at: the line gets status 2 and the caller re-parses it with ps_libsvm_parse_line.
 *
 *   newline_count_kernel   256 threads x 16 bytes per block: newlines per 4 KB chunk
 *   chunk_scan_kernel      one block: exclusive scan of the chunk counts (a 64 MB text has 16 K chunks)
 *   newline_index_kernel   line_end[rank] = byte offset of the rank-th newline (order-preserving compaction)
 *   parse_lines_kernel     one thread per line: single pass over its bytes, rows written in the step's staging layout
 * One thread per line is deliberate: a batch is a few thousand lines of ~700 B that sit in L1/L2 after the first touch, the
 * kernel is latency- not bandwidth-bound at that size, and a line's tokens have to be counted in order anyway.
 */
#include "ingest.cuh"

namespace psb {

namespace {

constexpr int kChunk = 4096;           /* bytes per block of the newline passes */
__device__ const float kP10[11] = {1.f, 1e1f, 1e2f, 1e3f, 1e4f, 1e5f, 1e6f, 1e7f, 1e8f, 1e9f, 1e10f};

__global__ void __launch_bounds__(256) newline_count_kernel(const char* __restrict__ text, size_t len, uint32_t* __restrict__ chunk_count) {
  const size_t base = (size_t)blockIdx.x * kChunk + (size_t)threadIdx.x * 16;
  int c = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) c += (base + i < len && text[base + i] == '\n') ? 1 : 0;
  __shared__ int warp_sum[8];
  for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += warp_sum[i];
    chunk_count[blockIdx.x] = (uint32_t)s;
  }
}

/* exclusive scan in place over n chunk counts, total to *total; one block of 1024 threads walks the array in tiles */
__global__ void __launch_bounds__(1024) chunk_scan_kernel(uint32_t* __restrict__ chunk_count, int n, uint32_t* __restrict__ total) {
  __shared__ uint32_t warp_tot[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t0 = 0; t0 < n; t0 += 1024) {
    const int i = t0 + threadIdx.x;
    const uint32_t v = i < n ? chunk_count[i] : 0u;
    uint32_t inc = v;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = warp_tot[lane];
      for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
      warp_tot[lane] = w;                 /* inclusive totals of the warps */
    }
    __syncthreads();
    const uint32_t before = carry + (warp > 0 ? warp_tot[warp - 1] : 0u) + inc - v;
    if (i < n) chunk_count[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(256) newline_index_kernel(const char* __restrict__ text, size_t len, const uint32_t* __restrict__ chunk_base,
                                                            uint32_t* __restrict__ line_end, uint32_t max_lines) {
  const size_t base = (size_t)blockIdx.x * kChunk + (size_t)threadIdx.x * 16;
  uint32_t mask = 0u;
#pragma unroll
  for (int i = 0; i < 16; ++i) mask |= (base + i < len && text[base + i] == '\n') ? (1u << i) : 0u;
  const int c = __popc(mask);
  /* exclusive prefix of c over the block's 256 threads */
  __shared__ int warp_sum[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = c;
  for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
  if (lane == 31) warp_sum[warp] = inc;
  __syncthreads();
  int before = inc - c;
  for (int w = 0; w < warp; ++w) before += warp_sum[w];
  uint32_t rank = chunk_base[blockIdx.x] + (uint32_t)before;
  while (mask) {
    const int i = __ffs(mask) - 1;
    mask &= mask - 1;
    if (rank < max_lines) line_end[rank] = (uint32_t)(base + i);
    ++rank;
  }
}

/* [-]digits[.digits], <= 7 significant digits, <= 10 decimals, terminated by ' ' or the end of the line: the host parser's
 * parse_float_fast.  Returns the new position, or nullptr when the spelling needs the general parser. */
__device__ __forceinline__ const char* dev_decimal(const char* q, const char* e, float* out) {
  bool neg = false;
  if (q < e && *q == '-') { neg = true; ++q; }
  uint32_t m = 0;
  int sig = 0, k = 0, digits = 0;
  bool dot = false;
  for (; q < e && *q != ' '; ++q) {
    const unsigned d = (unsigned)(*q - '0');
    if (d <= 9u) {
      m = m * 10 + d; ++digits;
      if (m != 0 || sig > 0) ++sig;
      if (dot) ++k;
      if (sig > 7 || k > 10) return nullptr;
    } else if (*q == '.' && !dot) {
      dot = true;
    } else {
      return nullptr;
    }
  }
  if (digits == 0) return nullptr;
  const float f = __fdiv_rn((float)m, kP10[k]);   /* both exact in fp32: one correctly rounded division, as on the host */
  *out = neg ? -f : f;
  return q;
}

__global__ void __launch_bounds__(128) parse_lines_kernel(const char* __restrict__ text, size_t len, const uint32_t* __restrict__ line_end, int rows,
                                                          int F, int Xn, long long wide, long long* __restrict__ E, float* __restrict__ X,
                                                          long long* __restrict__ W, float* __restrict__ Y, unsigned char* __restrict__ status,
                                                          const uint32_t* __restrict__ total, uint32_t* __restrict__ bad) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  if (total != nullptr) {                        /* asynchronous form: exactly `rows` lines are expected, nobody looked at the count */
    if (r == 0 && *total != (uint32_t)rows) atomicAdd(bad, 1u);
    if ((uint32_t)r >= *total) { status[r] = 3; return; }
  }
  const char* b = text + (r == 0 ? 0 : (size_t)line_end[r - 1] + 1);
  const char* e = text + line_end[r];
  if (e > b && e[-1] == '\r') --e;
  /* StringUtils.isBlank -> IndexOutOfBounds in parseFeature */
  {
    const char* p = b;
    while (p < e && (unsigned char)*p <= ' ') ++p;
    if (p == e) { status[r] = 1; if (bad != nullptr) atomicAdd(bad, 1u); return; }
  }
  while (e > b && e[-1] == ' ') --e;             /* String.split(" ") drops trailing empty strings */
  const int need = 1 + F + Xn;
  int col = 0;
  const char* p = b;
  unsigned char st = 0;
  while (true) {
    if (col == 0) {
      float y;
      const char* q = dev_decimal(p, e, &y);
      if (q == nullptr) { st = 2; break; }
      Y[r] = y;
      p = q;
    } else {
      unsigned long long v = 0;
      int nd = 0;
      while (p < e && (unsigned)(*p - '0') <= 9u) { v = v * 10 + (unsigned long long)(*p - '0'); ++p; ++nd; }
      if (nd == 0 || nd > 18 || p >= e || *p != ':') { st = 2; break; }
      ++p;
      float val;
      const char* q = dev_decimal(p, e, &val);
      if (q == nullptr) { st = 2; break; }
      p = q;
      if (col <= F) {
        const float idf = (float)(long long)v;          /* E[j-1][i] = cols.get(j).getIdx(): long -> float (CTR.java:57) */
        E[(size_t)r * F + col - 1] = (long long)idf;
        W[(size_t)r * F + col - 1] = (v < (1ull << 24) && wide < (1ll << 24)) ? (long long)(v % (unsigned long long)wide) : (long long)fmodf(idf, (float)wide);
      } else if (col < need) {
        X[(size_t)r * Xn + col - 1 - F] = val;          /* X[j-24][i] = cols.get(j).toF() */
      }
    }
    ++col;
    if (p >= e) break;
    ++p;                                                /* the separating ' ' */
  }
  if (st == 0 && col < need) st = 1;                    /* short line */
  status[r] = st;
  if (st != 0 && bad != nullptr) atomicAdd(bad, 1u);
}

}  // namespace

/* text_dev: `len` bytes of complete lines (each ends in '\n'); returns the number of lines found (at most max_rows are parsed) */
int libsvm_parse_dev(Ctx* ctx, const char* text_dev, size_t len, int F, int Xn, int64_t wide, int max_rows, int64_t* E, float* X, int64_t* W, float* Y,
                     uint8_t* status, uint32_t* ws /* >= len / 4096 + 2 + max_rows uint32 */) {
  PS_REQUIRE(text_dev && E && X && W && Y && status && ws && F >= 0 && Xn >= 0 && wide > 0 && max_rows > 0, PS_ERR_ARG, "libsvm_parse_dev: bad argument");
  PS_REQUIRE(len > 0 && len < (1ull << 32), PS_ERR_ARG, "libsvm_parse_dev: text must be 1 byte .. 4 GiB");
  cudaStream_t s = ctx->stream;
  const int chunks = (int)((len + kChunk - 1) / kChunk);
  uint32_t* chunk_count = ws;
  uint32_t* total = ws + chunks;
  uint32_t* line_end = ws + chunks + 1;
  newline_count_kernel<<<chunks, 256, 0, s>>>(text_dev, len, chunk_count);
  chunk_scan_kernel<<<1, 1024, 0, s>>>(chunk_count, chunks, total);
  newline_index_kernel<<<chunks, 256, 0, s>>>(text_dev, len, chunk_count, line_end, (uint32_t)max_rows);
  PS_LAUNCH_CHECK();
  uint32_t n = 0;
  PS_CUDA(cudaMemcpyAsync(&n, total, sizeof n, cudaMemcpyDeviceToHost, s));
  PS_CUDA(cudaStreamSynchronize(s));
  const int rows = (int)std::min<uint32_t>(n, (uint32_t)max_rows);
  if (rows > 0) {
    parse_lines_kernel<<<ceil_div(rows, 128), 128, 0, s>>>(text_dev, len, line_end, rows, F, Xn, (long long)wide, reinterpret_cast<long long*>(E), X,
                                                          reinterpret_cast<long long*>(W), Y, status, nullptr, nullptr);
    PS_LAUNCH_CHECK();
  }
  ctx->launches += 3 + (rows > 0 ? 1 : 0);
  return rows;
}

/* the same without any host synchronisation: the text is expected to hold exactly `rows` lines; *bad (device memory) ends up as the
 * number of lines that are missing, surplus-flagged, blank, short or outside the fast path's spellings — 0 means E/X/W/Y hold
 * exactly what the host reader would have produced.  Everything is enqueued on ctx->stream.                               */
void libsvm_parse_dev_async(Ctx* ctx, const char* text_dev, size_t len, int F, int Xn, int64_t wide, int rows, int64_t* E, float* X, int64_t* W, float* Y,
                            uint8_t* status, uint32_t* ws, uint32_t* bad) {
  PS_REQUIRE(text_dev && E && X && W && Y && status && ws && bad && F >= 0 && Xn >= 0 && wide > 0 && rows > 0, PS_ERR_ARG, "libsvm_parse_dev_async: bad argument");
  PS_REQUIRE(len > 0 && len < (1ull << 32), PS_ERR_ARG, "libsvm_parse_dev_async: text must be 1 byte .. 4 GiB");
  cudaStream_t s = ctx->stream;
  const int chunks = (int)((len + kChunk - 1) / kChunk);
  uint32_t* chunk_count = ws;
  uint32_t* total = ws + chunks;
  uint32_t* line_end = ws + chunks + 1;
  PS_CUDA(cudaMemsetAsync(bad, 0, sizeof(uint32_t), s));
  newline_count_kernel<<<chunks, 256, 0, s>>>(text_dev, len, chunk_count);
  chunk_scan_kernel<<<1, 1024, 0, s>>>(chunk_count, chunks, total);
  newline_index_kernel<<<chunks, 256, 0, s>>>(text_dev, len, chunk_count, line_end, (uint32_t)rows);
  parse_lines_kernel<<<ceil_div(rows, 128), 128, 0, s>>>(text_dev, len, line_end, rows, F, Xn, (long long)wide, reinterpret_cast<long long*>(E), X,
                                                        reinterpret_cast<long long*>(W), Y, status, total, bad);
  PS_LAUNCH_CHECK();
  ctx->launches += 4;
}

}  // namespace psb
