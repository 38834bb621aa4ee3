/*
 * ingest.cuh — native libsvm reader feeding the step's staging buffers (see ingest.cu).
 * Stands behind data.LibsvmParser.parse (LibsvmParser.java:13-25), CTR.parseFeature (CTR.java:47-68),
 * data.DataSource.readLine (DataSource.java:25-46) and data.DataSet.next/hasNext/reset/run (DataSet.java:37-100).
 */
#pragma once
#include <atomic>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "common.cuh"

namespace psb {

enum { LINE_OK = 0, LINE_SHORT = 1, LINE_BAD = 2 };

/* one text line [b, e) -> one sample: E[F], X[Xn], W[F], Y[1] (any output may be null) */
int parse_ctr_line(const char* b, const char* e, int F, int Xn, int64_t wide, int64_t* E, float* X, int64_t* W, float* Y);

/* the same on the GPU for the spellings the fast path accepts (ingest_dev.cu): device pointers, asynchronous on ctx->stream except
 * for the line count it returns; status 2 = spelling outside the fast path, re-parse that line on the host */
int libsvm_parse_dev(Ctx* ctx, const char* text_dev, size_t len, int F, int Xn, int64_t wide, int max_rows, int64_t* E, float* X, int64_t* W, float* Y,
                     uint8_t* status, uint32_t* ws /* >= len / 4096 + 2 + max_rows uint32 */);

/* fully asynchronous form for Model::submit_text: exactly `rows` lines expected, *bad_dev counts the lines that are not usable */
void libsvm_parse_dev_async(Ctx* ctx, const char* text_dev, size_t len, int F, int Xn, int64_t wide, int rows, int64_t* E, float* X, int64_t* W, float* Y,
                            uint8_t* status, uint32_t* ws /* >= len / 4096 + 2 + rows uint32 */, uint32_t* bad_dev);

struct LibsvmReader {
  struct Batch;
  int F, Xn;
  int64_t wide;
  int batch, offset, step, threads, depth;
  /* the mapped file and this reader's cursor (producer thread only) */
  int fd = -1;
  const char* data = nullptr;
  size_t size = 0, pos = 0;
  int64_t line_no = 0;
  /* statistics */
  std::atomic<int64_t> lines_read{0}, batches{0}, dropped{0};
  /* producer -> consumer queue */
  std::mutex mu;
  std::condition_variable cv_data, cv_space;
  std::deque<std::unique_ptr<Batch>> queue;
  std::thread producer;
  bool quit = false, produced_eof = false;
  /* parse workers: threads - 1 persistent helpers (the producer parses too) take 64-line chunks of the batch being parsed.  The chunk
   * counter carries the job's generation in its upper half, so a helper that is late leaving the previous job cannot touch the next one. */
  struct ParseJob { const std::pair<const char*, const char*>* lines = nullptr; Batch* out = nullptr; int* status = nullptr; int n = 0; };
  std::vector<std::thread> workers;
  std::mutex wmu;
  std::condition_variable wcv_job, wcv_done;
  ParseJob job;
  uint64_t job_gen = 0;
  std::atomic<uint64_t> chunk_counter{0};
  std::atomic<int> chunks_done{0};
  int n_chunks = 0;
  bool wquit = false;
  void parse_chunks(const ParseJob& j, uint64_t gen, int nchunks);
  void worker_loop();
  uint64_t cur_gen = 0;
  void begin_parse(const std::vector<std::pair<const char*, const char*>>& lines, Batch* out, int* status);
  void finish_parse();
  bool gather(std::vector<std::pair<const char*, const char*>>& lines);

  LibsvmReader(const std::string& path, int F, int Xn, int64_t wide, int batch, int offset, int step, int threads, int depth);
  ~LibsvmReader();
  LibsvmReader(const LibsvmReader&) = delete;
  int next(int64_t* E, float* X, int64_t* W, float* Y);   /* rows of the batch; 0 = end of data (DataSet.next() == null) */
  void reset();

 private:
  void start();
  void stop();
  void produce();
  bool raw_line(const char** b, const char** e);
  bool next_line(const char** b, const char** e);
  void rewind_to_after(const char* line_end, int count);
};

}  // namespace psb
