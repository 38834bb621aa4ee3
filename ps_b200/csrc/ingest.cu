/*
 * ingest.cu — libsvm ingest that feeds the training step (host code; no kernels).
 *
 * Path restated (reference, /root/reference/src/main/java/):
 *   data/LibsvmParser.java:13-25      line -> [Feature(0, label), Feature(idx, value) ...]  (split on ' ', then on ':')
 *   CTR.java:47-68 (parseFeature)     Y = value of column 0; E[j] = (float) idx of columns 1..F; X[x] = value of columns
 *                                     F+1..F+Xn; W = MatrixUtil.hash(E, wideSize) = E % wideSize in float (MatrixUtil.java:27-33)
 *   data/DataSource.java:25-46        a reader sees file lines offset, offset+step, offset+2*step, ...
 *   data/DataSet.java:77-100 (run)    `batch` lines per batch; a short last batch is delivered; an exception while a batch is
 *                                     being assembled is swallowed (`// ignore`) and the lines read so far are LOST:
 *                                       - a line that does not parse (NumberFormatException, missing ':') throws inside
 *                                         parser.parse: the partial batch is dropped and the next batch starts at the next line
 *                                       - a blank or short line (< 1+F+Xn columns) only throws later, inside parseFeature
 *                                         (IndexOutOfBounds): the WHOLE batch is dropped
 *   data/DataSet.java:37-43,61-67     next() == null at end of data; reset() rewinds
 *
 * Here: the file is mapped once, a producer thread keeps `depth` parsed batches ahead of the consumer (the reference's reader
 * thread + ArrayBlockingQueue(thread * 2)), a batch is parsed by `threads` workers over disjoint line ranges (deterministic,
 * unlike the reference's thread > 1 readers, which interleave lines), and ps_reader_next copies straight into the caller's
 * (pinned) E/X/W/Y staging buffers in the layout ps_model_submit takes.
 */
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cctype>
#include <cerrno>
#include <cmath>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "ingest.cuh"

namespace psb {

namespace {

inline bool is_java_ws(char c) { return (unsigned char)c <= ' '; }          /* String.trim() / Character.isWhitespace for ASCII */

/* Long.parseLong: [+-]?digits, no whitespace, overflow is an error */
bool parse_long(const char* b, const char* e, int64_t* out) {
  if (b >= e) return false;
  bool neg = false;
  if (*b == '-' || *b == '+') { neg = *b == '-'; ++b; }
  if (b >= e) return false;
  uint64_t v = 0;
  const uint64_t lim = neg ? (uint64_t)1 << 63 : ((uint64_t)1 << 63) - 1;
  for (; b < e; ++b) {
    if (*b < '0' || *b > '9') return false;
    const uint64_t d = (uint64_t)(*b - '0');
    if (v > (lim - d) / 10) return false;
    v = v * 10 + d;
  }
  *out = neg ? (int64_t)(0 - v) : (int64_t)v;
  return true;
}

/* Float.parseFloat: trims whitespace, then [+-]? ( NaN | Infinity | decimal [eE][+-]?digits [fFdD]? | hex float with a p exponent ).
 * The value is the correctly rounded float of the decimal string, which is what glibc strtof returns. */
bool parse_float(const char* b, const char* e, float* out) {
  while (b < e && is_java_ws(*b)) ++b;
  while (e > b && is_java_ws(e[-1])) --e;
  if (b >= e || e - b > 200) return false;
  const char* p = b;
  if (*p == '+' || *p == '-') ++p;
  const size_t rest = (size_t)(e - p);
  if (rest == 3 && std::memcmp(p, "NaN", 3) == 0) { *out = NAN; return true; }
  if (rest == 8 && std::memcmp(p, "Infinity", 8) == 0) { *out = *b == '-' ? -INFINITY : INFINITY; return true; }
  const char* q = p;
  bool hex = false;
  if (rest > 2 && q[0] == '0' && (q[1] == 'x' || q[1] == 'X')) {
    hex = true; q += 2;
    int nd = 0;
    while (q < e && std::isxdigit((unsigned char)*q)) { ++q; ++nd; }
    if (q < e && *q == '.') { ++q; while (q < e && std::isxdigit((unsigned char)*q)) { ++q; ++nd; } }
    if (nd == 0 || q >= e || (*q != 'p' && *q != 'P')) return false;     /* Java requires the binary exponent */
    ++q;
    if (q < e && (*q == '+' || *q == '-')) ++q;
    int ne = 0;
    while (q < e && *q >= '0' && *q <= '9') { ++q; ++ne; }
    if (ne == 0) return false;
  } else {
    int nd = 0;
    while (q < e && *q >= '0' && *q <= '9') { ++q; ++nd; }
    if (q < e && *q == '.') { ++q; while (q < e && *q >= '0' && *q <= '9') { ++q; ++nd; } }
    if (nd == 0) return false;
    if (q < e && (*q == 'e' || *q == 'E')) {
      ++q;
      if (q < e && (*q == '+' || *q == '-')) ++q;
      int ne = 0;
      while (q < e && *q >= '0' && *q <= '9') { ++q; ++ne; }
      if (ne == 0) return false;
    }
  }
  const char* num_end = q;
  if (q < e && (*q == 'f' || *q == 'F' || *q == 'd' || *q == 'D')) ++q;
  if (q != e) return false;
  (void)hex;
  char buf[208];
  const size_t n = (size_t)(num_end - b);
  std::memcpy(buf, b, n);
  buf[n] = 0;
  char* endp = nullptr;
  *out = std::strtof(buf, &endp);
  return endp == buf + n;
}

/* fast path for the overwhelmingly common spellings ("1", "0.48"): up to 7 significant digits, no exponent — the decimal value
 * m / 10^k with m < 2^24 and k <= 10 is a correctly rounded single division of two exactly representable numbers */
inline bool parse_float_fast(const char* b, const char* e, float* out) {
  static const float p10[11] = {1.f, 1e1f, 1e2f, 1e3f, 1e4f, 1e5f, 1e6f, 1e7f, 1e8f, 1e9f, 1e10f};
  const char* p = b;
  bool neg = false;
  if (p < e && *p == '-') { neg = true; ++p; }
  uint32_t m = 0; int nd = 0, k = 0; bool dot = false;
  if (p >= e) return false;
  for (; p < e; ++p) {
    if (*p == '.') { if (dot) return false; dot = true; continue; }
    if (*p < '0' || *p > '9') return false;
    m = m * 10 + (uint32_t)(*p - '0');
    if (m != 0 || nd > 0) ++nd;                 /* significant digits (leading zeros are free) */
    if (dot) ++k;
    if (nd > 7 || k > 10) return false;
  }
  if (nd == 0 && m == 0 && (e - b) == (neg ? 2 : 1) && b[neg ? 1 : 0] == '.') return false;   /* a lone "." */
  const float v = (float)m / p10[k];             /* both exact in fp32 (m < 10^7 < 2^24; 10^k exact for k <= 10) => one rounding */
  *out = neg ? -v : v;
  return true;
}

/* The common token "digits:decimal" in ONE pass over its bytes (idx up to 18 digits: cannot overflow a long; value as in
 * parse_float_fast).  Leaves p at the token's end on success; on anything unusual returns false with p untouched and the
 * general code below decides. */
inline bool fast_pair(const char*& p, const char* e, int64_t* idx, float* val) {
  static const float p10[11] = {1.f, 1e1f, 1e2f, 1e3f, 1e4f, 1e5f, 1e6f, 1e7f, 1e8f, 1e9f, 1e10f};
  const char* q = p;
  uint64_t v = 0;
  int nd = 0;
  while (q < e && (unsigned)(*q - '0') <= 9u) { v = v * 10 + (uint64_t)(*q - '0'); ++q; ++nd; }
  if (nd == 0 || nd > 18 || q >= e || *q != ':') return false;
  ++q;
  if (q < e && *q == '1' && (q + 1 == e || q[1] == ' ')) {   /* "<id>:1": every categorical column of a CTR line */
    *idx = (int64_t)v; *val = 1.0f; p = q + 1;
    return true;
  }
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
      if (sig > 7 || k > 10) return false;
    } else if (*q == '.' && !dot) {
      dot = true;
    } else {
      return false;
    }
  }
  if (digits == 0) return false;
  const float f = (float)m / p10[k];
  *idx = (int64_t)v;
  *val = neg ? -f : f;
  p = q;
  return true;
}

}  // namespace

/* One line -> one row of E/X/W/Y.  Returns LINE_OK, LINE_SHORT (blank or fewer than 1+F+Xn columns: IndexOutOfBounds in
 * parseFeature) or LINE_BAD (an exception inside LibsvmParser.parse).                                                     */
int parse_ctr_line(const char* b, const char* e, int F, int Xn, int64_t wide, int64_t* E, float* X, int64_t* W, float* Y) {
  /* StringUtils.isBlank -> empty feature list */
  {
    const char* p = b;
    while (p < e && is_java_ws(*p)) ++p;
    if (p == e) return LINE_SHORT;
  }
  /* String.split(" "): trailing empty strings are removed, leading / inner ones are kept (and then fail to parse) */
  while (e > b && e[-1] == ' ') --e;
  int col = 0;
  const int need = 1 + F + Xn;
  bool shortline = false;
  const char* p = b;
  /* idx % wide for 32-bit operands without a division per column (Lemire's fastmod: exact for a, d < 2^32) */
  const bool small_wide = wide > 0 && wide < (1 << 24);
  const uint64_t modM = small_wide ? UINT64_C(0xFFFFFFFFFFFFFFFF) / (uint64_t)wide + 1 : 0;
  while (true) {
    const char* t = p;
    int64_t idx;
    float val;
    bool have_pair = false;
    if (col > 0 && fast_pair(t, e, &idx, &val)) have_pair = true;        /* t now at the token's end */
    else while (t < e && *t != ' ') ++t;         /* token [p, t) */
    if (col == 0) {
      float y;
      if (!parse_float_fast(p, t, &y) && !parse_float(p, t, &y)) return LINE_BAD;
      if (Y) *Y = y;
    } else {
      if (!have_pair) {
      const char* c = p;
      while (c < t && *c != ':') ++c;            /* pair[0] = [p, c) */
      if (c == t) {                              /* no ':' — "abc".split(":") has one element, pair[1] throws; "5:" likewise */
        return LINE_BAD;
      }
      const char* v = c + 1;
      const char* ve = v;
      while (ve < t && *ve != ':') ++ve;         /* pair[1] = [v, ve); further ':' segments are ignored */
      /* "5:" -> split drops the trailing empty string -> pair.length == 1 -> ArrayIndexOutOfBounds */
      bool only_empty_after = true;
      for (const char* z = v; z < t; ++z) if (*z != ':') { only_empty_after = false; break; }
      if (only_empty_after) return LINE_BAD;
      if (!parse_long(p, c, &idx)) return LINE_BAD;
      if (!parse_float_fast(v, ve, &val) && !parse_float(v, ve, &val)) return LINE_BAD;
      }
      if (col <= F) {
        const float idf = (float)idx;            /* E[j-1][i] = cols.get(j).getIdx()  (long -> float, CTR.java:57) */
        if (E) E[col - 1] = (int64_t)idf;
        /* result.data[i] % size in float (MatrixUtil.java:30): exact integer arithmetic while the id survives the float cast */
        if (W) {
          if (small_wide && idx >= 0 && idx < (1 << 24)) {
            const uint64_t low = modM * (uint64_t)idx;
            W[col - 1] = (int64_t)(uint64_t)(((unsigned __int128)low * (uint64_t)wide) >> 64);
          } else {
            W[col - 1] = (idx >= 0 && idx < (1 << 24) && wide < (1 << 24)) ? idx % wide : (int64_t)std::fmod(idf, (float)wide);
          }
        }
      } else if (col < need) {
        if (X) X[col - 1 - F] = val;             /* X[j-24][i] = cols.get(j).toF() */
      }
    }
    ++col;
    if (t >= e) break;
    p = t + 1;
  }
  if (col < need) shortline = true;
  return shortline ? LINE_SHORT : LINE_OK;
}

/* ------------------------------------------------------------------ LibsvmReader */
/* batch storage that is NOT zero-filled on allocation (every element of a delivered batch is written by the parser; a 4096-row batch is 2.3 MB) */
template <class T> struct RawBuf {
  std::unique_ptr<T[]> p;
  void resize(size_t n) { p.reset(new T[n]); }
  T* data() { return p.get(); }
  const T* data() const { return p.get(); }
};
struct LibsvmReader::Batch {
  int rows = 0;
  bool eof = false;
  RawBuf<int64_t> E, W;
  RawBuf<float> X, Y;
};

LibsvmReader::LibsvmReader(const std::string& path, int F_, int Xn_, int64_t wide_, int batch_, int offset_, int step_, int threads_, int depth_)
    : F(F_), Xn(Xn_), wide(wide_), batch(batch_), offset(offset_), step(step_), threads(threads_ < 1 ? 1 : threads_), depth(depth_ < 1 ? 1 : depth_) {
  PS_REQUIRE(F >= 0 && Xn >= 0 && batch > 0 && offset >= 0 && step >= 1 && wide > 0, PS_ERR_ARG, "reader: bad argument");
  fd = ::open(path.c_str(), O_RDONLY);
  PS_REQUIRE(fd >= 0, PS_NOT_FOUND, ("reader: cannot open " + path).c_str());
  struct stat st;
  if (fstat(fd, &st) != 0) { ::close(fd); throw Error(PS_ERR_ARG, "reader: fstat failed"); }
  size = (size_t)st.st_size;
  if (size > 0) {
    void* m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) { ::close(fd); throw Error(PS_ERR_ARG, "reader: mmap failed"); }
    data = static_cast<const char*>(m);
    madvise(const_cast<char*>(data), size, MADV_SEQUENTIAL);
  }
  start();
}

LibsvmReader::~LibsvmReader() {
  stop();
  {
    std::lock_guard<std::mutex> g(wmu);
    wquit = true;
  }
  wcv_job.notify_all();
  for (auto& w : workers) if (w.joinable()) w.join();
  workers.clear();
  if (data) munmap(const_cast<char*>(data), size);
  if (fd >= 0) ::close(fd);
}

void LibsvmReader::start() {
  pos = 0; line_no = 0; quit = false; produced_eof = false;
  if (workers.empty())                             /* the helpers outlive reset(): DataSet.reset between epochs restarts the producer only */
    for (int t = 1; t < threads; ++t) workers.emplace_back([this] { this->worker_loop(); });
  producer = std::thread([this] { this->produce(); });
}

void LibsvmReader::stop() {
  {
    std::lock_guard<std::mutex> g(mu);
    quit = true;
  }
  cv_space.notify_all();
  cv_data.notify_all();
  if (producer.joinable()) producer.join();
  queue.clear();
}

static constexpr int kParseChunk = 64;            /* lines per unit of work */

/* takes chunks of job `gen` until none is left (or the counter has moved on to another job) */
void LibsvmReader::parse_chunks(const ParseJob& j, uint64_t gen, int nchunks) {
  while (true) {
    /* claim by compare-exchange: a helper that is late leaving job g must not take (and lose) a chunk of job g + 1 */
    uint64_t t = chunk_counter.load(std::memory_order_acquire);
    int c;
    while (true) {
      if ((t >> 32) != (gen & 0xFFFFFFFFull)) return;
      c = (int)(t & 0xFFFFFFFFull);
      if (c >= nchunks) return;
      if (chunk_counter.compare_exchange_weak(t, t + 1, std::memory_order_acq_rel, std::memory_order_acquire)) break;
    }
    const int lo = c * kParseChunk, hi = std::min(j.n, lo + kParseChunk);
    for (int i = lo; i < hi; ++i)
      j.status[i] = parse_ctr_line(j.lines[i].first, j.lines[i].second, F, Xn, wide, j.out->E.data() + (size_t)i * F, j.out->X.data() + (size_t)i * Xn,
                                   j.out->W.data() + (size_t)i * F, j.out->Y.data() + i);
    if (chunks_done.fetch_add(1, std::memory_order_acq_rel) + 1 == nchunks) {
      std::lock_guard<std::mutex> g(wmu);
      wcv_done.notify_all();
    }
  }
}

void LibsvmReader::worker_loop() {
  uint64_t seen = 0;
  while (true) {
    ParseJob j; uint64_t gen; int nchunks;
    {
      std::unique_lock<std::mutex> g(wmu);
      wcv_job.wait(g, [&] { return wquit || job_gen != seen; });
      if (wquit) return;
      seen = gen = job_gen; j = job; nchunks = n_chunks;
    }
    parse_chunks(j, gen, nchunks);
  }
}

void LibsvmReader::reset() {                      /* DataSet.reset: shutdownNow, queue.clear, source.reset, start */
  stop();
  start();
}

/* BufferedReader.readLine over the mapping: next line [b, e) without its terminator; false at end of file */
bool LibsvmReader::raw_line(const char** b, const char** e) {
  if (pos >= size) return false;
  const char* p = data + pos;
  const char* nl = static_cast<const char*>(std::memchr(p, '\n', size - pos));
  const char* end = nl ? nl : data + size;
  pos = (size_t)((nl ? nl + 1 : end) - data);
  *b = p;
  *e = (end > p && end[-1] == '\r') ? end - 1 : end;
  return true;
}

/* DataSource.readLine (DataSource.java:25-46): the first call skips `offset` lines, later calls advance by `step` */
bool LibsvmReader::next_line(const char** b, const char** e) {
  const int64_t skip = line_no == 0 ? offset : step - 1;
  for (int64_t i = 0; i < skip; ++i) {
    const char *sb, *se;
    if (!raw_line(&sb, &se)) return false;
  }
  if (!raw_line(b, e)) return false;
  ++line_no;
  return true;
}

/* every line of the batch parsed into `out`, status[i] = LINE_*.  begin_parse publishes the job and wakes the helpers; the producer gathers the
 * NEXT batch's lines meanwhile (the serial part of a batch: one memchr pass over its 2-3 MB), then finish_parse takes chunks itself and
 * returns when all are done */
void LibsvmReader::begin_parse(const std::vector<std::pair<const char*, const char*>>& lines, Batch* out, int* status) {
  ParseJob j;
  j.lines = lines.data(); j.out = out; j.status = status; j.n = (int)lines.size();
  const int nchunks = (j.n + kParseChunk - 1) / kParseChunk;
  {
    std::lock_guard<std::mutex> g(wmu);
    job = j; n_chunks = nchunks; cur_gen = ++job_gen;
    chunks_done.store(0, std::memory_order_release);
    chunk_counter.store((cur_gen & 0xFFFFFFFFull) << 32, std::memory_order_release);
  }
  if (!workers.empty() && nchunks > 1) wcv_job.notify_all();
}
void LibsvmReader::finish_parse() {
  const ParseJob j = job;                        /* (only this thread writes it) */
  const int nchunks = n_chunks;
  parse_chunks(j, cur_gen, nchunks);
  std::unique_lock<std::mutex> g(wmu);
  wcv_done.wait(g, [&] { return chunks_done.load(std::memory_order_acquire) >= nchunks; });
}

/* DataSet.run: gather up to `batch` lines; true when the end of the data was reached while gathering */
bool LibsvmReader::gather(std::vector<std::pair<const char*, const char*>>& lines) {
  lines.clear();
  while ((int)lines.size() < batch) {
    const char *b, *e;
    if (!next_line(&b, &e)) return true;
    lines.emplace_back(b, e);
  }
  return false;
}

void LibsvmReader::produce() {
  std::vector<std::pair<const char*, const char*>> lines, ahead;
  std::vector<int> status;
  bool eof = gather(lines);
  while (!lines.empty()) {
    std::unique_ptr<Batch> out(new Batch);
    const int n = (int)lines.size();
    out->E.resize((size_t)n * F); out->W.resize((size_t)n * F); out->X.resize((size_t)n * Xn); out->Y.resize(n);
    status.assign(n, LINE_OK);
    begin_parse(lines, out.get(), status.data());
    /* look ahead while the helpers parse; the cursor is remembered so that a bad line in THIS batch can take the lookahead back */
    const size_t pos_before = pos;
    const int64_t line_no_before = line_no;
    bool eof_ahead = eof;
    ahead.clear();
    if (!eof) eof_ahead = gather(ahead);
    finish_parse();
    /* the reference's swallowed exceptions: a line that fails inside parser.parse drops what was gathered up to and
     * including it — the lines after it (already consumed here) form the head of the next batch */
    int first_bad = -1;
    for (int i = 0; i < n; ++i) if (status[i] == LINE_BAD) { first_bad = i; break; }
    lines_read += n;
    if (first_bad >= 0) {
      ++dropped;
      pos = pos_before; line_no = line_no_before;  /* undo the lookahead, then push this batch's tail back: the cursor goes to just after the offending line */
      rewind_to_after(lines[first_bad].second, n - 1 - first_bad);
      lines_read -= n - 1 - first_bad;
      eof = gather(lines);
      continue;
    }
    bool any_short = false;
    for (int i = 0; i < n; ++i) any_short |= status[i] == LINE_SHORT;
    if (any_short) ++dropped;                      /* IndexOutOfBounds in parseFeature: the whole batch is lost */
    else {
      out->rows = n;
      ++batches;
      std::unique_lock<std::mutex> g(mu);
      cv_space.wait(g, [this] { return quit || (int)queue.size() < depth; });
      if (quit) return;
      queue.push_back(std::move(out));
      g.unlock();
      cv_data.notify_one();
    }
    lines.swap(ahead);
    eof = eof_ahead;
  }
  std::unique_ptr<Batch> fin(new Batch);
  fin->eof = true;
  std::unique_lock<std::mutex> g(mu);
  cv_space.wait(g, [this] { return quit || (int)queue.size() < depth + 1; });
  if (quit) return;
  queue.push_back(std::move(fin));
  produced_eof = true;
  g.unlock();
  cv_data.notify_all();
}

/* un-read `count` of this reader's lines: the cursor returns to the byte after `line_end`'s terminator */
void LibsvmReader::rewind_to_after(const char* line_end, int count) {
  const char* p = line_end;
  if (p < data + size && *p == '\r') ++p;
  if (p < data + size && *p == '\n') ++p;
  pos = (size_t)(p - data);
  line_no -= count;
}

int LibsvmReader::next(int64_t* E, float* X, int64_t* W, float* Y) {
  std::unique_lock<std::mutex> g(mu);
  cv_data.wait(g, [this] { return !queue.empty(); });
  if (queue.front()->eof) return 0;               /* stays at the front: every later call reports end of data too */
  std::unique_ptr<Batch> b = std::move(queue.front());
  queue.pop_front();
  g.unlock();
  cv_space.notify_one();
  const size_t n = (size_t)b->rows;
  if (E && F) std::memcpy(E, b->E.data(), sizeof(int64_t) * n * F);
  if (W && F) std::memcpy(W, b->W.data(), sizeof(int64_t) * n * F);
  if (X && Xn) std::memcpy(X, b->X.data(), sizeof(float) * n * Xn);
  if (Y) std::memcpy(Y, b->Y.data(), sizeof(float) * n);
  return b->rows;
}

}  // namespace psb
