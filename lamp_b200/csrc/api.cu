// api.cu -- the C ABI of include/etgpu.h: contexts, resident data, forests.
#include <stdarg.h>

#include <algorithm>
#include <condition_variable>
#include <memory>
#include <thread>

#include "internal.h"

static thread_local std::string g_last_error;

void et_set_error(const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}

#define ET_API_BEGIN try {
#define ET_API_END                                           \
  }                                                          \
  catch (const EtError &e) { return e.code; }                \
  catch (const std::bad_alloc &) {                           \
    et_set_error("host allocation failed");                  \
    return ET_ENOMEM;                                        \
  }                                                          \
  return ET_OK;

extern "C" int32_t et_abi_version(void) { return ET_ABI_VERSION; }
extern "C" const char *et_last_error(void) { return g_last_error.c_str(); }
extern "C" double et_debug_repeat_add(double c, int64_t h) { return et_repeat_add(c, h); }

// ---- device block cache ----------------------------------------------------------------------
// sizes are rounded up to 8 classes per octave, so blocks of similar size (forests of consecutive builds)
// are interchangeable
static size_t block_class(size_t bytes) {
  bytes = std::max<size_t>(bytes, 256);
  if (bytes <= ((size_t)1 << 20)) return (bytes + 255) / 256 * 256;
  int lg = 63 - __builtin_clzll((unsigned long long)bytes);
  const size_t step = (size_t)1 << (lg - 3);
  return (bytes + step - 1) / step * step;
}

void *et_dev_alloc(et_ctx *ctx, size_t bytes) {
  bytes = block_class(bytes);
  for (size_t i = 0; i < ctx->cache.size(); i++) {
    if (ctx->cache[i].bytes == bytes) {
      void *p = ctx->cache[i].p;
      ctx->cache_bytes -= bytes;
      ctx->cache.erase(ctx->cache.begin() + (long)i);
      return p;
    }
  }
  void *p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) {
    cudaGetLastError();
    // give the cached blocks back and retry once
    for (auto &b : ctx->cache) cudaFree(b.p);
    ctx->cache.clear();
    ctx->cache_bytes = 0;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
  }
  return p;
}

void et_dev_free(et_ctx *ctx, void *p, size_t bytes) {
  if (!p) return;
  bytes = block_class(bytes);
  const size_t kMaxBlocks = 32, kMaxBytes = (size_t)8 << 30;
  if (!ctx || bytes > kMaxBytes) {
    cudaFree(p);
    return;
  }
  while (!ctx->cache.empty() && (ctx->cache.size() >= kMaxBlocks || ctx->cache_bytes + bytes > kMaxBytes)) {
    cudaFree(ctx->cache.front().p);  // oldest first
    ctx->cache_bytes -= ctx->cache.front().bytes;
    ctx->cache.erase(ctx->cache.begin());
  }
  ctx->cache.push_back({p, bytes});
  ctx->cache_bytes += bytes;
}

// ---- uploads from pageable host memory -------------------------------------------------------------------
// (the ingest path of the reference, lamp-saddle SaddleTensorHelpers.scala:156-176: Mat.toArray -> native copy ->
// device; a JVM caller hands over a pageable heap array)
struct HostStager {
  static constexpr size_t PIECE = (size_t)8 << 20;
  void *buf[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  bool used[2] = {false, false};
  // copy threads: a job is one memcpy cut into slices
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv_job, cv_done;
  const char *src = nullptr;
  char *dst = nullptr;
  size_t len = 0;
  int gen = 0, pending = 0;
  bool stop = false;
  void run(int t, int T) {
    int seen = 0;
    for (;;) {
      const char *s;
      char *d;
      size_t n;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_job.wait(lk, [&] { return stop || gen != seen; });
        if (stop) return;
        seen = gen;
        s = src;
        d = dst;
        n = len;
      }
      const size_t a = n * (size_t)t / (size_t)T, b = n * (size_t)(t + 1) / (size_t)T;
      if (b > a) memcpy(d + a, s + a, b - a);
      {
        std::lock_guard<std::mutex> lk(mu);
        if (--pending == 0) cv_done.notify_one();
      }
    }
  }
  void copy(char *d, const char *s, size_t n) {  // parallel memcpy, returns when done
    const int T = (int)workers.size();
    if (T == 0) {
      memcpy(d, s, n);
      return;
    }
    std::unique_lock<std::mutex> lk(mu);
    src = s;
    dst = d;
    len = n;
    pending = T;
    gen++;
    cv_job.notify_all();
    cv_done.wait(lk, [&] { return pending == 0; });
  }
};

void et_stager_free(HostStager *s) {
  if (!s) return;
  {
    std::lock_guard<std::mutex> lk(s->mu);
    s->stop = true;
  }
  s->cv_job.notify_all();
  for (auto &t : s->workers) t.join();
  for (int i = 0; i < 2; i++) {
    if (s->buf[i]) cudaFreeHost(s->buf[i]);
    if (s->ev[i]) cudaEventDestroy(s->ev[i]);
  }
  delete s;
}

void et_h2d(et_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return;
  cudaPointerAttributes at;
  bool pageable = true;
  if (cudaPointerGetAttributes(&at, src_host) == cudaSuccess)
    pageable = (at.type == cudaMemoryTypeUnregistered);
  else
    cudaGetLastError();
  static const bool off = getenv("ETGPU_NO_STAGER") != nullptr && atoi(getenv("ETGPU_NO_STAGER")) != 0;
  if (!pageable || off || bytes < ((size_t)1 << 20)) {
    CUDA_CHECK(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, st));
    return;
  }
  if (!ctx->stager) {
    std::unique_ptr<HostStager> s(new HostStager());
    for (int i = 0; i < 2; i++) {
      if (cudaHostAlloc(&s->buf[i], HostStager::PIECE, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        for (int j = 0; j < i; j++) cudaFreeHost(s->buf[j]);
        s->buf[0] = s->buf[1] = nullptr;
        CUDA_CHECK(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, st));  // no pinned memory: plain path
        return;
      }
      CUDA_CHECK(cudaEventCreateWithFlags(&s->ev[i], cudaEventDisableTiming));
    }
    const int T = (int)std::max(1u, std::min(4u, std::thread::hardware_concurrency() / 2));
    HostStager *raw = s.get();
    for (int t = 0; t < T; t++) raw->workers.emplace_back([raw, t, T] { raw->run(t, T); });
    ctx->stager = s.release();
  }
  HostStager &S = *ctx->stager;
  const char *src = static_cast<const char *>(src_host);
  char *dst = static_cast<char *>(dst_dev);
  int b = 0;
  for (size_t off2 = 0; off2 < bytes; off2 += HostStager::PIECE, b ^= 1) {
    const size_t n = std::min(HostStager::PIECE, bytes - off2);
    if (S.used[b]) CUDA_CHECK(cudaEventSynchronize(S.ev[b]));  // the DMA that last read this buffer is done
    S.copy(static_cast<char *>(S.buf[b]), src + off2, n);      // (the other buffer's DMA runs meanwhile)
    CUDA_CHECK(cudaMemcpyAsync(dst + off2, S.buf[b], n, cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaEventRecord(S.ev[b], st));
    S.used[b] = true;
  }
}

// ---- context --------------------------------------------------------------------------------
extern "C" int et_init(int32_t device, et_ctx **out) {
  ET_API_BEGIN
  if (!out) ET_FAIL(ET_EINVAL, "et_init: out is NULL");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    cudaGetLastError();
    ET_FAIL(ET_ECUDA, "et_init: no CUDA device (%s); libetgpu has no CPU fallback",
            e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
  }
  if (device < 0 || device >= count) ET_FAIL(ET_EINVAL, "et_init: device %d out of range [0,%d)", device, count);
  CUDA_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    ET_FAIL(ET_ECUDA, "et_init: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
            prop.minor);
  et_ctx *c = new et_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  int prio_least = 0, prio_greatest = 0;
  CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
  CUDA_CHECK(cudaStreamCreateWithPriority(&c->own_stream, cudaStreamNonBlocking, prio_greatest));
  c->stream = c->own_stream;
  CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  for (int i = 0; i < et_ctx::N_SIDE; i++) {
    CUDA_CHECK(cudaStreamCreateWithPriority(&c->side[i], cudaStreamNonBlocking, prio_greatest));
    CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming));
  }
  *out = c;
  ET_API_END
}

void et_ctx_acquire(et_ctx *ctx) {
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  ctx->live_handles++;
}

static void ctx_destroy(et_ctx *ctx);

void et_ctx_release(et_ctx *ctx) {
  bool destroy = false;
  {
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    ctx->live_handles--;
    destroy = ctx->shutdown_requested && ctx->live_handles <= 0;
  }
  if (destroy) ctx_destroy(ctx);
}

// With et_data / et_forest handles still alive the teardown is deferred until the last of them is freed.
extern "C" void et_shutdown(et_ctx *ctx) {
  if (!ctx) return;
  {
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (ctx->live_handles > 0) {
      ctx->shutdown_requested = true;
      return;
    }
  }
  ctx_destroy(ctx);
}

static void ctx_destroy(et_ctx *ctx) {
  if (ctx->is_multi()) {
    et_multi_shutdown(ctx);
    delete ctx;
    return;
  }
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  for (int i = 0; i < et_ctx::N_SIDE; i++) {
    if (ctx->side[i]) cudaStreamDestroy(ctx->side[i]);
    if (ctx->ev_join[i]) cudaEventDestroy(ctx->ev_join[i]);
  }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  et_stager_free(ctx->stager);
  if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
  if (ctx->ev_comm) cudaEventDestroy(ctx->ev_comm);
  et_workspace_free(ctx->ws);
  for (auto &b : ctx->cache) cudaFree(b.p);
  delete ctx;
}

extern "C" int et_set_stream(et_ctx *ctx, void *cuda_stream) {
  ET_API_BEGIN
  if (!ctx) ET_FAIL(ET_EINVAL, "et_set_stream: ctx is NULL");
  if (ctx->is_multi()) ET_FAIL(ET_EUNSUPPORTED, "et_set_stream: a multi-GPU context runs on its own streams");
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  ET_API_END
}

extern "C" int et_synchronize(et_ctx *ctx) {
  ET_API_BEGIN
  if (!ctx) ET_FAIL(ET_EINVAL, "et_synchronize: ctx is NULL");
  if (ctx->is_multi()) {
    for (et_ctx *c : ctx->peers) {
      CUDA_CHECK(cudaSetDevice(c->device));
      CUDA_CHECK(cudaStreamSynchronize(c->stream));
    }
    return ET_OK;
  }
  CUDA_CHECK(cudaSetDevice(ctx->device));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  ET_API_END
}

// ---- row-major -> column-major transpose ----------------------------------------------------
// 32x32 tiles through padded shared memory: coalesced 256-byte row reads, coalesced column writes.
__global__ void k_transpose(const double *__restrict__ src, int64_t rows, int32_t d, double *__restrict__ dst,
                            int64_t ld, int64_t row0) {
  __shared__ double tile[32][33];
  int64_t r0 = (int64_t)blockIdx.x * 32;
  int32_t c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int64_t r = r0 + j;
    int32_t c = c0 + threadIdx.x;
    if (r < rows && c < d) tile[j][threadIdx.x] = src[r * d + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int32_t c = c0 + j;
    int64_t r = r0 + threadIdx.x;
    if (r < rows && c < d) dst[(int64_t)c * ld + row0 + r] = tile[threadIdx.x][j];
  }
}

void et_launch_transpose(et_ctx *ctx, const double *src, int64_t rows, int32_t d, double *dst, int64_t ld,
                         int64_t row0) {
  if (rows <= 0 || d <= 0) return;
  dim3 block(32, 8);
  dim3 grid((unsigned)ceil_div(rows, 32), (unsigned)ceil_div(d, 32));
  k_transpose<<<grid, block, 0, ctx->stream>>>(src, rows, d, dst, ld, row0);
  ctx->launches++;
  CUDA_CHECK(cudaGetLastError());
}

// ---- data -----------------------------------------------------------------------------------
et_data *et_data_alloc_internal(et_ctx *ctx, int64_t n, int32_t d);
static et_data *data_alloc(et_ctx *ctx, int64_t n, int32_t d) { return et_data_alloc_internal(ctx, n, d); }

// front handle of a table replicated on every GPU of a multi-GPU context: made from the first GPU's table
static et_data *multi_front_data(et_ctx *front, et_data *d0, bool replicate) {
  et_data *F = new et_data();
  F->ctx = front;
  F->n = d0->n;
  F->d = d0->d;
  F->ld = d0->ld;
  F->shards.push_back(d0);
  if (replicate) {
    try {
      et_multi_replicate(front, F);
    } catch (...) {
      et_data_free(d0);
      delete F;
      throw;
    }
  }
  return F;
}

et_data *et_data_alloc_internal(et_ctx *ctx, int64_t n, int32_t d) {
  if (n < 0 || d < 0) ET_FAIL(ET_EINVAL, "negative table dimensions");
  if (n > 0x7fffffff) ET_FAIL(ET_EUNSUPPORTED, "tables with more than 2^31-1 rows are not supported");
  et_data *D = new et_data();
  D->ctx = ctx;
  D->n = n;
  D->d = d;
  D->ld = ((n + 15) / 16) * 16;
  size_t bytes = (size_t)std::max<int64_t>(D->ld, 16) * (size_t)std::max(d, 1) * sizeof(double);
  D->x = static_cast<double *>(et_dev_alloc(ctx, bytes));
  D->x_bytes = bytes;
  if (!D->x) {
    delete D;
    ET_FAIL(ET_ENOMEM, "cannot allocate %zu bytes of HBM for the %lld x %d table", bytes, (long long)n, d);
  }
  return D;
}

extern "C" int et_data_dense_alloc(et_ctx *ctx, int64_t n, int32_t d, et_data **out) {
  ET_API_BEGIN
  if (!ctx || !out) ET_FAIL(ET_EINVAL, "et_data_dense_alloc: NULL argument");
  if (ctx->is_multi()) {
    et_data *F = new et_data();
    F->ctx = ctx;
    try {
      for (et_ctx *c : ctx->peers) {
        et_data *r = nullptr;
        int rc = et_data_dense_alloc(c, n, d, &r);
        if (rc != ET_OK) throw EtError{rc};
        F->shards.push_back(r);
      }
    } catch (...) {
      for (et_data *r : F->shards) et_data_free(r);
      delete F;
      throw;
    }
    F->n = n;
    F->d = d;
    F->ld = F->shards[0]->ld;
    *out = F;
    return ET_OK;
  }
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  CUDA_CHECK(cudaSetDevice(ctx->device));
  *out = data_alloc(ctx, n, d);
  ET_API_END
}

extern "C" int et_data_dense_colblock(et_ctx *ctx, et_data *D, const double *cols, int32_t first_col,
                                      int32_t n_cols) {
  ET_API_BEGIN
  if (!ctx || !D || (!cols && n_cols > 0)) ET_FAIL(ET_EINVAL, "et_data_dense_colblock: NULL argument");
  if (ctx->is_multi()) {  // host -> first GPU, then NVLink to the others
    if (D->shards.size() != ctx->peers.size()) ET_FAIL(ET_EINVAL, "et_data_dense_colblock: not a table of this context");
    int rc = et_data_dense_colblock(ctx->peers[0], D->shards[0], cols, first_col, n_cols);
    if (rc != ET_OK) return rc;
    et_multi_broadcast_columns(ctx, D, first_col, n_cols);
    return ET_OK;
  }
  if (!D->x) ET_FAIL(ET_EINVAL, "et_data_dense_colblock: not a dense table");
  if (first_col < 0 || n_cols < 0 || first_col + n_cols > D->d)
    ET_FAIL(ET_EINVAL, "et_data_dense_colblock: columns [%d,%d) outside [0,%d)", first_col, first_col + n_cols, D->d);
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  CUDA_CHECK(cudaSetDevice(ctx->device));
  if (n_cols > 0 && D->n > 0)
    CUDA_CHECK(cudaMemcpy2DAsync(D->x + (int64_t)first_col * D->ld, (size_t)D->ld * sizeof(double), cols,
                                 (size_t)D->n * sizeof(double), (size_t)D->n * sizeof(double), (size_t)n_cols,
                                 cudaMemcpyHostToDevice, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  if (D->coded != 0) {  // the coded copy is rebuilt at the next build
    et_data_drop_codes(D);
    D->coded = 0;
  }
  ET_API_END
}

extern "C" int et_data_dense_rowmajor(et_ctx *ctx, const double *x, int64_t n, int32_t d, et_data **out) {
  ET_API_BEGIN
  if (!ctx || !out || (!x && n > 0 && d > 0)) ET_FAIL(ET_EINVAL, "et_data_dense_rowmajor: NULL argument");
  if (ctx->is_multi()) {  // one upload to the first GPU, replicas over NVLink
    et_data *d0 = nullptr;
    int rc = et_data_dense_rowmajor(ctx->peers[0], x, n, d, &d0);
    if (rc != ET_OK) return rc;
    *out = multi_front_data(ctx, d0, true);
    return ET_OK;
  }
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  CUDA_CHECK(cudaSetDevice(ctx->device));
  et_data *D = data_alloc(ctx, n, d);
  if (n > 0 && d > 0) {
    // Staged in row chunks through TWO device buffers: the host->device copy of chunk c + 1 (copy stream) overlaps
    // the transpose of chunk c (compute stream).  Pageable sources go through the pinned stager (et_h2d).
    int64_t chunk = std::max<int64_t>(1, ((int64_t)64 << 20) / ((int64_t)d * 8));
    chunk = std::min(chunk, n);
    const size_t stage_bytes = (size_t)chunk * d * sizeof(double);
    double *stage[2] = {static_cast<double *>(et_dev_alloc(ctx, stage_bytes)), nullptr};
    stage[1] = (chunk < n) ? static_cast<double *>(et_dev_alloc(ctx, stage_bytes)) : stage[0];
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
    auto cleanup = [&]() {
      et_dev_free(ctx, stage[0], stage_bytes);
      if (stage[1] != stage[0]) et_dev_free(ctx, stage[1], stage_bytes);
      for (int i = 0; i < 2; i++) {
        if (ev_copied[i]) cudaEventDestroy(ev_copied[i]);
        if (ev_free[i]) cudaEventDestroy(ev_free[i]);
      }
    };
    if (!stage[0] || !stage[1]) {
      cleanup();
      et_data_free(D);
      ET_FAIL(ET_ENOMEM, "cannot allocate the upload staging buffers");
    }
    try {
      cudaStream_t copy_st = ctx->side[0];
      for (int i = 0; i < 2; i++) {
        CUDA_CHECK(cudaEventCreateWithFlags(&ev_copied[i], cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&ev_free[i], cudaEventDisableTiming));
      }
      CUDA_CHECK(cudaEventRecord(ev_free[0], ctx->stream));  // (earlier work on the compute stream comes first)
      CUDA_CHECK(cudaStreamWaitEvent(copy_st, ev_free[0], 0));
      int b = 0;
      bool used[2] = {false, false};
      for (int64_t r0 = 0; r0 < n; r0 += chunk, b ^= 1) {
        const int64_t rows = std::min(chunk, n - r0);
        if (used[b]) CUDA_CHECK(cudaStreamWaitEvent(copy_st, ev_free[b], 0));  // its last transpose has read it
        et_h2d(ctx, stage[b], x + r0 * d, (size_t)rows * d * sizeof(double), copy_st);
        CUDA_CHECK(cudaEventRecord(ev_copied[b], copy_st));
        CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ev_copied[b], 0));
        et_launch_transpose(ctx, stage[b], rows, d, D->x, D->ld, r0);
        CUDA_CHECK(cudaEventRecord(ev_free[b], ctx->stream));
        used[b] = true;
      }
      et_data_encode(ctx, D);
      CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
      CUDA_CHECK(cudaStreamSynchronize(copy_st));
    } catch (...) {
      cudaStreamSynchronize(ctx->stream);
      cudaStreamSynchronize(ctx->side[0]);
      cleanup();
      et_data_free(D);
      throw;
    }
    cleanup();
  }
  *out = D;
  ET_API_END
}

extern "C" int et_data_dense_rowmajor_device(et_ctx *ctx, const double *x_dev, int64_t n, int32_t d,
                                             et_data **out) {
  ET_API_BEGIN
  if (!ctx || !out || (!x_dev && n > 0 && d > 0)) ET_FAIL(ET_EINVAL, "et_data_dense_rowmajor_device: NULL argument");
  if (ctx->is_multi()) ET_FAIL(ET_EUNSUPPORTED, "device-pointer inputs need a single-GPU context");
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  CUDA_CHECK(cudaSetDevice(ctx->device));
  et_data *D = data_alloc(ctx, n, d);
  try {
    et_launch_transpose(ctx, x_dev, n, d, D->x, D->ld, 0);
    et_data_encode(ctx, D);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  } catch (...) {
    et_data_free(D);
    throw;
  }
  *out = D;
  ET_API_END
}

// one thread per stored entry of a block of columns: x[col][row] = val (the table was zero-filled before)
__global__ void k_csc_scatter(const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx,
                              const double *__restrict__ val, int32_t c0, int32_t ncols, int64_t e0, int64_t ne,
                              double *__restrict__ x, int64_t ld) {
  const int64_t e = e0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= e0 + ne) return;
  int lo = 0, hi = ncols;  // the column of entry e: last c with colptr[c0 + c] <= e
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (colptr[c0 + mid] <= e)
      lo = mid;
    else
      hi = mid;
  }
  x[(int64_t)(c0 + lo) * ld + rowidx[e - e0]] = val[e - e0];
}

// row-major index of a sparse-resident table: entries per row, exclusive scan, fill through per-row cursors
__global__ void k_csr_count(const int32_t *__restrict__ rowidx, int64_t nnz, unsigned long long *__restrict__ cnt) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < nnz) atomicAdd(&cnt[rowidx[e]], 1ull);
}
__global__ void __launch_bounds__(1024) k_csr_scan(unsigned long long *cnt, int64_t n, int64_t *ptr) {
  // one CTA: every thread sums a contiguous slice, the slices are scanned through shared memory
  __shared__ unsigned long long s_part[1024];
  const int64_t per = (n + 1023) / 1024, a = (int64_t)threadIdx.x * per, b = min(n, a + per);
  unsigned long long s = 0;
  for (int64_t i = a; i < b; i++) s += cnt[i];
  s_part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    for (int t = 0; t < 1024; t++) {
      const unsigned long long v = s_part[t];
      s_part[t] = run;
      run += v;
    }
    ptr[n] = (int64_t)run;
  }
  __syncthreads();
  unsigned long long run = s_part[threadIdx.x];
  for (int64_t i = a; i < b; i++) {
    const unsigned long long v = cnt[i];
    ptr[i] = (int64_t)run;
    cnt[i] = run;  // becomes the row's fill cursor
    run += v;
  }
}
__global__ void k_csr_fill(const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx, int32_t d, int64_t nnz,
                           unsigned long long *__restrict__ cursor, int32_t *__restrict__ csr_col) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  int lo = 0, hi = d;  // the column of entry e: last c with colptr[c] <= e
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (colptr[mid] <= e)
      lo = mid;
    else
      hi = mid;
  }
  csr_col[atomicAdd(&cursor[rowidx[e]], 1ull)] = lo;
}

// the row-major index (which features a row stores) of a sparse-resident table: small free-running nodes mark every
// feature none of their rows stores as constant up front (build.cuh sparse_mark_constant)
void et_data_build_csr(et_ctx *ctx, et_data *D) {
  const int64_t n = D->n, nnz = D->csc_nnz;
  unsigned long long *d_cnt = nullptr;
  CUDA_CHECK(cudaMalloc((void **)&D->csr_ptr, ((size_t)n + 1) * sizeof(int64_t)));
  CUDA_CHECK(cudaMalloc((void **)&D->csr_col, std::max<size_t>(1, (size_t)nnz) * sizeof(int32_t)));
  CUDA_CHECK(cudaMalloc((void **)&d_cnt, std::max<size_t>(1, (size_t)n) * sizeof(unsigned long long)));
  cudaMemsetAsync(d_cnt, 0, std::max<size_t>(1, (size_t)n) * sizeof(unsigned long long), ctx->stream);
  if (nnz > 0) k_csr_count<<<(unsigned)ceil_div(nnz, 256), 256, 0, ctx->stream>>>(D->csc_row, nnz, d_cnt);
  k_csr_scan<<<1, 1024, 0, ctx->stream>>>(d_cnt, n, D->csr_ptr);
  if (nnz > 0)
    k_csr_fill<<<(unsigned)ceil_div(nnz, 256), 256, 0, ctx->stream>>>(D->csc_colptr, D->csc_row, D->d, nnz, d_cnt, D->csr_col);
  ctx->launches += 3;
  cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_cnt);
  CUDA_CHECK(e2);
  CUDA_CHECK(cudaGetLastError());
}

// CSC input (BASELINE configs[3]).  The stored entries of a column are kept in ascending row order without
// duplicates (a row listed twice keeps the LATER entry; unsorted input is sorted on the host first), so a kernel
// finds the value of (row, column) by binary search and everything not stored is an implicit 0.0 with dense
// semantics.  Tables whose dense form is small (ETGPU_CSC_DENSE_MAX bytes, default 4 GiB) are expanded into the
// resident column-major FP64 matrix instead: one gather per value is cheaper than a search, and such a table can
// still be byte-coded.  Larger tables stay sparse in HBM: 12 bytes per stored entry.
extern "C" int et_data_csc(et_ctx *ctx, const int64_t *colptr, const int32_t *rowidx, const double *val, int64_t n,
                           int32_t d, et_data **out) {
  ET_API_BEGIN
  if (!ctx || !out || !colptr) ET_FAIL(ET_EINVAL, "et_data_csc: NULL argument");
  if (ctx->is_multi()) {
    et_data *d0 = nullptr;
    int rc = et_data_csc(ctx->peers[0], colptr, rowidx, val, n, d, &d0);
    if (rc != ET_OK) return rc;
    *out = multi_front_data(ctx, d0, true);
    return ET_OK;
  }
  if (n < 0 || d < 0) ET_FAIL(ET_EINVAL, "negative table dimensions");
  if (n > 0x7fffffff) ET_FAIL(ET_EUNSUPPORTED, "tables with more than 2^31-1 rows are not supported");
  if (colptr[0] != 0) ET_FAIL(ET_EINVAL, "et_data_csc: colptr[0] must be 0");
  for (int32_t c = 0; c < d; c++)
    if (colptr[c + 1] < colptr[c]) ET_FAIL(ET_EINVAL, "et_data_csc: colptr decreases at column %d", c);
  int64_t nnz = colptr[d];
  if (nnz > 0 && (!rowidx || !val)) ET_FAIL(ET_EINVAL, "et_data_csc: NULL argument");
  bool clean = true;  // ascending rows without duplicates inside every column
  for (int32_t c = 0; c < d; c++) {
    for (int64_t e = colptr[c]; e < colptr[c + 1]; e++) {
      if (rowidx[e] < 0 || rowidx[e] >= n)
        ET_FAIL(ET_EINVAL, "et_data_csc: row index %d outside [0,%lld)", rowidx[e], (long long)n);
      if (e > colptr[c] && rowidx[e] <= rowidx[e - 1]) clean = false;
    }
  }
  // unsorted columns / duplicate rows: a cleaned copy (stable sort by row; of equal rows the last one stays)
  std::vector<int64_t> cp2;
  std::vector<int32_t> row2;
  std::vector<double> val2;
  if (!clean) {
    cp2.assign((size_t)d + 1, 0);
    row2.reserve((size_t)nnz);
    val2.reserve((size_t)nnz);
    std::vector<int64_t> ord;
    for (int32_t c = 0; c < d; c++) {
      const int64_t a = colptr[c], b = colptr[c + 1];
      ord.resize((size_t)(b - a));
      for (int64_t e = a; e < b; e++) ord[(size_t)(e - a)] = e;
      std::stable_sort(ord.begin(), ord.end(), [&](int64_t x, int64_t y) { return rowidx[x] < rowidx[y]; });
      for (size_t q = 0; q < ord.size(); q++) {
        if (q + 1 < ord.size() && rowidx[ord[q + 1]] == rowidx[ord[q]]) continue;  // a later entry of the same row follows
        row2.push_back(rowidx[ord[q]]);
        val2.push_back(val[ord[q]]);
      }
      cp2[(size_t)c + 1] = (int64_t)row2.size();
    }
    colptr = cp2.data();
    rowidx = row2.data();
    val = val2.data();
    nnz = colptr[d];
  }
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  CUDA_CHECK(cudaSetDevice(ctx->device));
  size_t dense_max = (size_t)4 << 30;
  if (const char *env = getenv("ETGPU_CSC_DENSE_MAX")) dense_max = (size_t)atoll(env);
  const bool expand = (size_t)n * (size_t)std::max(d, 1) * sizeof(double) <= dense_max;
  if (!expand) {
    // ---- the table stays sparse in HBM
    std::unique_ptr<et_data> D(new et_data());
    D->ctx = ctx;
    D->n = n;
    D->d = d;
    D->ld = ((n + 15) / 16) * 16;
    D->coded = -1;
    D->csc_nnz = nnz;
    try {
      CUDA_CHECK(cudaMalloc((void **)&D->csc_colptr, ((size_t)d + 1) * sizeof(int64_t)));
      CUDA_CHECK(cudaMalloc((void **)&D->csc_row, std::max<size_t>(1, (size_t)nnz) * sizeof(int32_t)));
      CUDA_CHECK(cudaMalloc((void **)&D->csc_val, std::max<size_t>(1, (size_t)nnz) * sizeof(double)));
      CUDA_CHECK(cudaMemcpyAsync(D->csc_colptr, colptr, ((size_t)d + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
      if (nnz > 0) {
        CUDA_CHECK(cudaMemcpyAsync(D->csc_row, rowidx, (size_t)nnz * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(D->csc_val, val, (size_t)nnz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
      }
      et_data_build_csr(ctx, D.get());
    } catch (...) {
      cudaFree(D->csc_colptr);
      cudaFree(D->csc_row);
      cudaFree(D->csc_val);
      cudaFree(D->csr_ptr);
      cudaFree(D->csr_col);
      D->csc_colptr = nullptr;
      D->csr_ptr = nullptr;
      throw;
    }
    *out = D.release();
    return ET_OK;
  }
  et_data *D = data_alloc(ctx, n, d);
  int64_t *d_colptr = nullptr;
  int32_t *d_row = nullptr;
  double *d_val = nullptr;
  const int64_t chunk_max = (int64_t)16 << 20;  // stored entries per upload (192 MB of staging)
  try {
    CUDA_CHECK(cudaMemsetAsync(D->x, 0, D->x_bytes, ctx->stream));
    if (nnz > 0 && n > 0) {
      CUDA_CHECK(cudaMalloc((void **)&d_colptr, ((size_t)d + 1) * sizeof(int64_t)));
      CUDA_CHECK(cudaMemcpyAsync(d_colptr, colptr, ((size_t)d + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
      const int64_t cap = std::min(nnz, chunk_max);
      CUDA_CHECK(cudaMalloc((void **)&d_row, (size_t)cap * sizeof(int32_t)));
      CUDA_CHECK(cudaMalloc((void **)&d_val, (size_t)cap * sizeof(double)));
      for (int64_t e0 = 0; e0 < nnz; e0 += cap) {
        const int64_t ne = std::min(cap, nnz - e0);
        // the columns this run of entries touches
        const int32_t c0 = (int32_t)(std::upper_bound(colptr, colptr + d + 1, e0) - colptr) - 1;
        const int32_t c1 = (int32_t)(std::upper_bound(colptr, colptr + d + 1, e0 + ne - 1) - colptr) - 1;
        CUDA_CHECK(cudaMemcpyAsync(d_row, rowidx + e0, (size_t)ne * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(d_val, val + e0, (size_t)ne * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        k_csc_scatter<<<(unsigned)ceil_div(ne, 256), 256, 0, ctx->stream>>>(d_colptr, d_row, d_val, c0, c1 - c0 + 1, e0, ne,
                                                                        D->x, D->ld);
        ctx->launches++;
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));  // (the staging buffers are reused)
      }
    }
    et_data_encode(ctx, D);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    CUDA_CHECK(cudaGetLastError());
  } catch (...) {
    cudaFree(d_colptr);
    cudaFree(d_row);
    cudaFree(d_val);
    et_data_free(D);
    throw;
  }
  cudaFree(d_colptr);
  cudaFree(d_row);
  cudaFree(d_val);
  *out = D;
  ET_API_END
}

template <typename T>
static void upload_vec(et_ctx *ctx, T **dst, const T *src, int64_t n) {
  if (*dst) {
    cudaFree(*dst);
    *dst = nullptr;
  }
  cudaError_t e = cudaMalloc((void **)dst, (size_t)std::max<int64_t>(n, 1) * sizeof(T));
  if (e != cudaSuccess) {
    cudaGetLastError();
    *dst = nullptr;
    ET_FAIL(ET_ENOMEM, "cannot allocate target/weight vector");
  }
  if (n > 0) {
    CUDA_CHECK(cudaMemcpyAsync(*dst, src, (size_t)n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  }
}

extern "C" int et_data_set_target_classification(et_ctx *ctx, et_data *D, const int32_t *y, int64_t n_target,
                                                 int32_t num_classes) {
  ET_API_BEGIN
  if (!ctx || !D || (!y && n_target > 0)) ET_FAIL(ET_EINVAL, "et_data_set_target_classification: NULL argument");
  if (ctx->is_multi()) {
    for (size_t g = 0; g < D->shards.size(); g++) {
      int rc = et_data_set_target_classification(ctx->peers[g], D->shards[g], y, n_target, num_classes);
      if (rc != ET_OK) return rc;
    }
    D->num_classes = num_classes;
    return ET_OK;
  }
  if (n_target != D->n)
    ET_FAIL(ET_EINVAL, "requirement failed: Data.numRows(%lld) != target.length (%lld)", (long long)D->n,
            (long long)n_target);
  if (num_classes <= 0) ET_FAIL(ET_EINVAL, "numClasses must be positive");
  std::vector<int64_t> hist((size_t)num_classes, 0);
  for (int64_t i = 0; i < n_target; i++) {
    if (y[i] < 0 || y[i] >= num_classes)
      ET_FAIL(ET_EINVAL, "target[%lld] = %d outside [0, numClasses=%d) (ArrayIndexOutOfBounds in the reference)",
              (long long)i, y[i], num_classes);
    hist[(size_t)y[i]]++;
  }
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  CUDA_CHECK(cudaSetDevice(ctx->device));
  upload_vec(ctx, &D->y_cls, y, n_target);
  D->num_classes = num_classes;
  D->root_hist = hist;
  ET_API_END
}

extern "C" int et_data_set_target_regression(et_ctx *ctx, et_data *D, const double *y, int64_t n_target) {
  ET_API_BEGIN
  if (!ctx || !D || (!y && n_target > 0)) ET_FAIL(ET_EINVAL, "et_data_set_target_regression: NULL argument");
  if (ctx->is_multi()) {
    for (size_t g = 0; g < D->shards.size(); g++) {
      int rc = et_data_set_target_regression(ctx->peers[g], D->shards[g], y, n_target);
      if (rc != ET_OK) return rc;
    }
    return ET_OK;
  }
  if (n_target != D->n)
    ET_FAIL(ET_EINVAL, "requirement failed: Data.numRows(%lld) != target.length (%lld)", (long long)D->n,
            (long long)n_target);
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  CUDA_CHECK(cudaSetDevice(ctx->device));
  upload_vec(ctx, &D->y_reg, y, n_target);
  ET_API_END
}

extern "C" int et_data_set_weights(et_ctx *ctx, et_data *D, const double *w, int64_t n_weights) {
  ET_API_BEGIN
  if (!ctx || !D) ET_FAIL(ET_EINVAL, "et_data_set_weights: NULL argument");
  if (ctx->is_multi()) {
    for (size_t g = 0; g < D->shards.size(); g++) {
      int rc = et_data_set_weights(ctx->peers[g], D->shards[g], w, n_weights);
      if (rc != ET_OK) return rc;
    }
    return ET_OK;
  }
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  CUDA_CHECK(cudaSetDevice(ctx->device));
  if (!w) {
    if (D->w) cudaFree(D->w);
    D->w = nullptr;
  } else {
    if (n_weights != D->n)
      ET_FAIL(ET_EINVAL, "sampleWeights.length (%lld) != Data.numRows (%lld)", (long long)n_weights, (long long)D->n);
    for (int64_t i = 0; i < n_weights; i++)
      if (w[i] < 0.0) ET_FAIL(ET_EINVAL, "requirement failed: Negative weights not allowed.");
    upload_vec(ctx, &D->w, w, n_weights);
  }
  ET_API_END
}

extern "C" int et_data_dims(const et_data *D, int64_t *n, int32_t *d) {
  if (!D) {
    et_set_error("et_data_dims: NULL data");
    return ET_EINVAL;
  }
  if (n) *n = D->n;
  if (d) *d = D->d;
  return ET_OK;
}

extern "C" void et_data_free(et_data *D) {
  if (!D) return;
  if (!D->shards.empty()) {  // front handle of a multi-GPU table
    for (et_data *r : D->shards) et_data_free(r);
    delete D;
    return;
  }
  if (D->ctx) cudaSetDevice(D->ctx->device);
  {
    std::unique_lock<std::recursive_mutex> lk;
    if (D->ctx) lk = std::unique_lock<std::recursive_mutex>(D->ctx->mu);
    et_dev_free(D->ctx, D->x, D->x_bytes);
    et_data_drop_codes(D);
    if (D->csc_colptr) cudaFree(D->csc_colptr);
    if (D->csc_row) cudaFree(D->csc_row);
    if (D->csc_val) cudaFree(D->csc_val);
    if (D->csr_ptr) cudaFree(D->csr_ptr);
    if (D->csr_col) cudaFree(D->csr_col);
    if (D->y_cls) cudaFree(D->y_cls);
    if (D->y_reg) cudaFree(D->y_reg);
    if (D->w) cudaFree(D->w);
  }
  delete D;  // (outside the lock: releasing the last handle may run the context's deferred shutdown)
}

// ---- build ----------------------------------------------------------------------------------
static void check_build_common(et_ctx *ctx, et_data *D, int32_t k, int32_t m, int32_t best_split, et_forest **out) {
  if (!ctx || !D || !out) ET_FAIL(ET_EINVAL, "build: NULL argument");
  if (D->ctx != ctx) ET_FAIL(ET_EINVAL, "build: data belongs to another context");
  if (m < 0) ET_FAIL(ET_EINVAL, "build: m must be >= 0");
  (void)k;
}

extern "C" int et_build_classification(et_ctx *ctx, et_data *D, const int32_t *target, int64_t n_target,
                                       const double *weights, int32_t num_classes, int32_t n_min, int32_t k,
                                       int32_t m, int32_t parallelism, int32_t best_split, int32_t max_depth,
                                       int64_t seed, const int32_t *tree_ids, const et_replay *replay,
                                       et_forest **out, et_stats *stats) {
  ET_API_BEGIN
  check_build_common(ctx, D, k, m, best_split, out);
  if (ctx->is_multi()) {
    if (target) {
      int rc = et_data_set_target_classification(ctx, D, target, n_target, num_classes);
      if (rc != ET_OK) return rc;
    }
    if (weights) {
      int rc = et_data_set_weights(ctx, D, weights, target ? n_target : D->n);
      if (rc != ET_OK) return rc;
    }
    et_data *d0 = D->shards.empty() ? nullptr : D->shards[0];
    if (!d0 || !d0->y_cls) ET_FAIL(ET_EINVAL, "build: no classification target attached");
    if (d0->num_classes != num_classes) ET_FAIL(ET_EINVAL, "build: numClasses differs from the attached target's");
    if (D->n == 0 && m > 0) ET_FAIL(ET_EINVAL, "build: empty table (the reference fails on targetInSubset.raw(0))");
    BuildArgs a;
    a.task = d0->w ? 1 : 0;
    a.num_classes = num_classes;
    a.n_min = n_min;
    a.k = k;
    a.m = m;
    a.parallelism = parallelism;
    a.best_split = best_split;
    a.max_depth = max_depth;
    a.seed = seed;
    a.tree_ids = tree_ids;
    a.replay = replay;
    et_multi_build(ctx, D, a, num_classes, 0, out, stats);
    return ET_OK;
  }
  // target and weights are independent: each is either uploaded for this call or, when NULL, the one attached
  // to `data` is used (et_data_set_weights(NULL) is the way to detach weights)
  if (target) {
    int rc = et_data_set_target_classification(ctx, D, target, n_target, num_classes);
    if (rc != ET_OK) return rc;
  }
  if (weights) {
    int rc = et_data_set_weights(ctx, D, weights, target ? n_target : D->n);
    if (rc != ET_OK) return rc;
  }
  if (!D->y_cls) ET_FAIL(ET_EINVAL, "build: no classification target attached");
  if (D->num_classes != num_classes) ET_FAIL(ET_EINVAL, "build: numClasses differs from the attached target's");
  if (D->n == 0 && m > 0) ET_FAIL(ET_EINVAL, "build: empty table (the reference fails on targetInSubset.raw(0))");
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  CUDA_CHECK(cudaSetDevice(ctx->device));
  BuildArgs a;
  a.task = D->w ? 1 : 0;
  a.num_classes = num_classes;
  a.n_min = n_min;
  a.k = k;
  a.m = m;
  a.parallelism = parallelism;
  a.best_split = best_split;
  a.max_depth = max_depth;
  a.seed = seed;
  a.tree_ids = tree_ids;
  a.replay = replay;
  et_forest *f = new et_forest();
  f->ctx = ctx;
  f->leaf_width = num_classes;
  f->is_regression = 0;
  f->m = m;
  try {
    et_build_forest(ctx, D, a, f, stats);
  } catch (...) {
    delete f;
    throw;
  }
  *out = f;
  ET_API_END
}

extern "C" int et_build_regression(et_ctx *ctx, et_data *D, const double *target, int64_t n_target, int32_t n_min,
                                   int32_t k, int32_t m, int32_t parallelism, int32_t best_split,
                                   int32_t max_depth, int64_t seed, const int32_t *tree_ids,
                                   const et_replay *replay, et_forest **out, et_stats *stats) {
  ET_API_BEGIN
  check_build_common(ctx, D, k, m, best_split, out);
  if (target) {
    int rc = et_data_set_target_regression(ctx, D, target, n_target);
    if (rc != ET_OK) return rc;
  }
  if (ctx->is_multi()) {
    et_data *d0 = D->shards.empty() ? nullptr : D->shards[0];
    if (!d0 || !d0->y_reg) ET_FAIL(ET_EINVAL, "build: no regression target attached");
    if (D->n == 0 && m > 0) ET_FAIL(ET_EINVAL, "requirement failed (subset.length > 0, pkg:779)");
    BuildArgs a;
    a.task = 2;
    a.num_classes = 1;
    a.n_min = n_min;
    a.k = k;
    a.m = m;
    a.parallelism = parallelism;
    a.best_split = best_split;
    a.max_depth = max_depth;
    a.seed = seed;
    a.tree_ids = tree_ids;
    a.replay = replay;
    et_multi_build(ctx, D, a, 1, 1, out, stats);
    return ET_OK;
  }
  if (!D->y_reg) ET_FAIL(ET_EINVAL, "build: no regression target attached");
  if (D->n == 0 && m > 0) ET_FAIL(ET_EINVAL, "requirement failed (subset.length > 0, pkg:779)");
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  CUDA_CHECK(cudaSetDevice(ctx->device));
  BuildArgs a;
  a.task = 2;
  a.num_classes = 1;
  a.n_min = n_min;
  a.k = k;
  a.m = m;
  a.parallelism = parallelism;
  a.best_split = best_split;
  a.max_depth = max_depth;
  a.seed = seed;
  a.tree_ids = tree_ids;
  a.replay = replay;
  et_forest *f = new et_forest();
  f->ctx = ctx;
  f->leaf_width = 1;
  f->is_regression = 1;
  f->m = m;
  try {
    et_build_forest(ctx, D, a, f, stats);
  } catch (...) {
    delete f;
    throw;
  }
  *out = f;
  ET_API_END
}

// ---- forest ---------------------------------------------------------------------------------
et_forest::~et_forest() {
  if (!shards.empty() || full) {  // front handle of a multi-GPU forest
    for (et_forest *s : shards) delete s;
    delete full;
    return;
  }
  if (ctx) cudaSetDevice(ctx->device);
  if (d_tree_off && !tree_off_bytes) cudaFree(d_tree_off);
  std::unique_lock<std::recursive_mutex> lk;
  if (ctx) lk = std::unique_lock<std::recursive_mutex>(ctx->mu);
  if (tree_off_bytes) et_dev_free(ctx, d_tree_off, tree_off_bytes);
  if (nodes_bytes)
    et_dev_free(ctx, d_nodes, nodes_bytes);
  else if (d_nodes)
    cudaFree(d_nodes);
  if (leaf_bytes)
    et_dev_free(ctx, d_leaf, leaf_bytes);
  else if (d_leaf)
    cudaFree(d_leaf);
}

extern "C" void et_forest_free(et_forest *f) { delete f; }

// The forest lives in HBM; the host copy is fetched on first export.
void et_forest_fetch(et_forest *f) {
  if (f->host_ready) return;
  if (f->full) {  // multi-GPU front handle: the gathered forest on the first GPU
    et_forest_fetch(f->full);
    f->h_nodes = f->full->h_nodes;
    f->h_leaf = f->full->h_leaf;
    f->host_ready = true;
    return;
  }
  std::lock_guard<std::recursive_mutex> lk(f->ctx->mu);
  if (f->host_ready) return;
  CUDA_CHECK(cudaSetDevice(f->ctx->device));
  f->h_nodes.resize((size_t)f->total_nodes);
  f->h_leaf.resize((size_t)f->total_leaves * (size_t)f->leaf_width);
  cudaStream_t st = f->ctx->stream;
  if (f->total_nodes)
    CUDA_CHECK(cudaMemcpyAsync(f->h_nodes.data(), f->d_nodes, f->h_nodes.size() * sizeof(PNode),
                               cudaMemcpyDeviceToHost, st));
  if (!f->h_leaf.empty())
    CUDA_CHECK(cudaMemcpyAsync(f->h_leaf.data(), f->d_leaf, f->h_leaf.size() * sizeof(double), cudaMemcpyDeviceToHost,
                               st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  f->host_ready = true;
}

extern "C" int et_forest_dims(const et_forest *f, int32_t *m, int32_t *leaf_width, int32_t *is_regression,
                              int64_t *total_nodes) {
  if (!f) {
    et_set_error("et_forest_dims: NULL forest");
    return ET_EINVAL;
  }
  if (m) *m = f->m;
  if (leaf_width) *leaf_width = f->leaf_width;
  if (is_regression) *is_regression = f->is_regression;
  if (total_nodes) *total_nodes = f->total_nodes;
  return ET_OK;
}

extern "C" int et_forest_tree_size(const et_forest *f, int32_t t, int32_t *n_nodes) {
  if (!f || !n_nodes || t < 0 || t >= f->m) {
    et_set_error("et_forest_tree_size: bad argument");
    return ET_EINVAL;
  }
  *n_nodes = (int32_t)(f->tree_off[(size_t)t + 1] - f->tree_off[(size_t)t]);
  return ET_OK;
}

static void export_tree(const et_forest *f, int32_t t, int32_t *feature, double *cut, uint8_t *mil, int32_t *left,
                        int32_t *right, double *leaf) {
  const int64_t off = f->tree_off[(size_t)t];
  const int64_t n = f->tree_off[(size_t)t + 1] - off;
  const int lw = f->leaf_width;
  for (int64_t i = 0; i < n; i++) {
    const PNode p = f->h_nodes[(size_t)(off + i)];
    if (p.feat >= 0) {
      feature[i] = p.feat & (ET_MIL_BIT - 1);
      mil[i] = (p.feat & ET_MIL_BIT) ? 1 : 0;
      cut[i] = p.cut;
      left[i] = (int32_t)i + 1;
      right[i] = p.right_or_leaf;
      for (int c = 0; c < lw; c++) leaf[i * lw + c] = 0.0;
    } else {
      feature[i] = -1;
      mil[i] = 0;
      cut[i] = NAN;
      left[i] = -1;
      right[i] = -1;
      const double *lv = f->h_leaf.data() + (size_t)p.right_or_leaf * (size_t)lw;
      for (int c = 0; c < lw; c++) leaf[i * lw + c] = lv[c];
    }
  }
}

extern "C" int et_forest_export(const et_forest *f, int32_t t, int32_t *feature, double *cut, uint8_t *mil,
                                int32_t *left, int32_t *right, double *leaf) {
  ET_API_BEGIN
  if (!f || t < 0 || t >= f->m || !feature || !cut || !mil || !left || !right || !leaf)
    ET_FAIL(ET_EINVAL, "et_forest_export: bad argument");
  et_forest_fetch(const_cast<et_forest *>(f));
  export_tree(f, t, feature, cut, mil, left, right, leaf);
  ET_API_END
}

extern "C" int et_forest_export_all(const et_forest *f, int32_t *tree_sizes, int32_t *feature, double *cut,
                                    uint8_t *mil, int32_t *left, int32_t *right, double *leaf) {
  ET_API_BEGIN
  if (!f || !tree_sizes || !feature || !cut || !mil || !left || !right || !leaf)
    ET_FAIL(ET_EINVAL, "et_forest_export_all: bad argument");
  et_forest_fetch(const_cast<et_forest *>(f));
  for (int32_t t = 0; t < f->m; t++) {
    size_t off = (size_t)f->tree_off[(size_t)t];
    tree_sizes[t] = (int32_t)(f->tree_off[(size_t)t + 1] - f->tree_off[(size_t)t]);
    export_tree(f, t, feature + off, cut + off, mil + off, left + off, right + off,
                leaf + off * (size_t)f->leaf_width);
  }
  ET_API_END
}

extern "C" int et_forest_import(et_ctx *ctx, int32_t m, int32_t leaf_width, int32_t is_regression,
                                const int32_t *tree_sizes, const int32_t *feature, const double *cut,
                                const uint8_t *mil, const int32_t *left, const int32_t *right, const double *leaf,
                                et_forest **out) {
  ET_API_BEGIN
  if (!ctx || !out || m < 0 || leaf_width <= 0) ET_FAIL(ET_EINVAL, "et_forest_import: bad argument");
  if (ctx->is_multi()) {  // host-held trees live on the first GPU (predict then runs there)
    et_forest *f0 = nullptr;
    int rc = et_forest_import(ctx->peers[0], m, leaf_width, is_regression, tree_sizes, feature, cut, mil, left, right, leaf, &f0);
    if (rc != ET_OK) return rc;
    et_forest *F = new et_forest();
    F->ctx = ctx;
    F->leaf_width = leaf_width;
    F->is_regression = is_regression;
    F->m = m;
    F->full = f0;
    F->total_nodes = f0->total_nodes;
    F->total_leaves = f0->total_leaves;
    F->tree_off = f0->tree_off;
    F->d_min = f0->d_min;
    *out = F;
    return ET_OK;
  }
  if (m > 0 && (!tree_sizes || !feature || !cut || !mil || !left || !right || !leaf))
    ET_FAIL(ET_EINVAL, "et_forest_import: NULL array");
  std::unique_ptr<et_forest> f(new et_forest());
  f->ctx = ctx;
  f->leaf_width = leaf_width;
  f->is_regression = is_regression;
  f->m = m;
  f->tree_off.assign((size_t)m + 1, 0);
  for (int32_t t = 0; t < m; t++) {
    if (tree_sizes[t] <= 0) ET_FAIL(ET_EINVAL, "et_forest_import: tree %d is empty", t);
    f->tree_off[(size_t)t + 1] = f->tree_off[(size_t)t] + tree_sizes[t];
  }
  f->total_nodes = f->tree_off[(size_t)m];
  f->h_nodes.resize((size_t)f->total_nodes);
  int64_t n_leaves = 0;
  for (int32_t t = 0; t < m; t++) {
    const int64_t off = f->tree_off[(size_t)t], n = tree_sizes[t];
    for (int64_t i = 0; i < n; i++) {
      PNode p;
      p.cut = cut[off + i];
      if (feature[off + i] >= 0) {
        // pre-order: left child directly follows its parent, right child lies further on
        if (left[off + i] != i + 1 || right[off + i] <= i + 1 || right[off + i] >= n)
          ET_FAIL(ET_EINVAL, "et_forest_import: tree %d node %lld is not in pre-order", t, (long long)i);
        if (feature[off + i] >= ET_MIL_BIT) ET_FAIL(ET_EINVAL, "et_forest_import: feature index too large");
        f->d_min = std::max(f->d_min, feature[off + i] + 1);
        p.feat = feature[off + i] | (mil[off + i] ? ET_MIL_BIT : 0);
        p.right_or_leaf = right[off + i];
      } else {
        if (n_leaves >= 0x7fffffff) ET_FAIL(ET_EUNSUPPORTED, "et_forest_import: more than 2^31 leaves");
        p.feat = -1;
        p.right_or_leaf = (int32_t)n_leaves++;
        for (int c = 0; c < leaf_width; c++) f->h_leaf.push_back(leaf[(off + i) * leaf_width + c]);
      }
      f->h_nodes[(size_t)(off + i)] = p;
    }
  }
  f->total_leaves = n_leaves;
  f->host_ready = true;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  CUDA_CHECK(cudaSetDevice(ctx->device));
  CUDA_CHECK(cudaMalloc((void **)&f->d_tree_off, ((size_t)m + 1) * sizeof(int64_t)));
  CUDA_CHECK(cudaMalloc((void **)&f->d_nodes, std::max<size_t>(1, (size_t)f->total_nodes) * sizeof(PNode)));
  CUDA_CHECK(cudaMalloc((void **)&f->d_leaf, std::max<size_t>(1, f->h_leaf.size()) * sizeof(double)));
  cudaStream_t st = ctx->stream;
  CUDA_CHECK(cudaMemcpyAsync(f->d_tree_off, f->tree_off.data(), ((size_t)m + 1) * sizeof(int64_t),
                             cudaMemcpyHostToDevice, st));
  if (f->total_nodes)
    CUDA_CHECK(cudaMemcpyAsync(f->d_nodes, f->h_nodes.data(), f->h_nodes.size() * sizeof(PNode),
                               cudaMemcpyHostToDevice, st));
  if (!f->h_leaf.empty())
    CUDA_CHECK(cudaMemcpyAsync(f->d_leaf, f->h_leaf.data(), f->h_leaf.size() * sizeof(double), cudaMemcpyHostToDevice,
                               st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  *out = f.release();
  ET_API_END
}

// ---- packed (device-layout) serialization ----------------------------------------------------
extern "C" int et_forest_packed_dims(const et_forest *f, int64_t *total_nodes, int64_t *total_leaves) {
  if (!f) {
    et_set_error("et_forest_packed_dims: NULL forest");
    return ET_EINVAL;
  }
  if (total_nodes) *total_nodes = f->total_nodes;
  if (total_leaves) *total_leaves = f->total_leaves;
  return ET_OK;
}

extern "C" int et_forest_export_packed(et_forest *f, void *nodes_out, double *leaves_out, int64_t *tree_off_out) {
  ET_API_BEGIN
  if (!f || !nodes_out || !leaves_out || !tree_off_out) ET_FAIL(ET_EINVAL, "et_forest_export_packed: NULL argument");
  if (f->full) return et_forest_export_packed(f->full, nodes_out, leaves_out, tree_off_out);
  std::lock_guard<std::recursive_mutex> lk(f->ctx->mu);
  CUDA_CHECK(cudaSetDevice(f->ctx->device));
  cudaStream_t st = f->ctx->stream;
  if (f->total_nodes)
    CUDA_CHECK(cudaMemcpyAsync(nodes_out, f->d_nodes, (size_t)f->total_nodes * sizeof(PNode), cudaMemcpyDeviceToHost, st));
  if (f->total_leaves)
    CUDA_CHECK(cudaMemcpyAsync(leaves_out, f->d_leaf, (size_t)f->total_leaves * (size_t)f->leaf_width * sizeof(double),
                               cudaMemcpyDeviceToHost, st));
  for (int32_t t = 0; t <= f->m; t++) tree_off_out[t] = f->tree_off[(size_t)t];
  CUDA_CHECK(cudaStreamSynchronize(st));
  ET_API_END
}

extern "C" int et_forest_import_packed(et_ctx *ctx, int32_t m, int32_t leaf_width, int32_t is_regression,
                                       int64_t total_nodes, int64_t total_leaves, const void *nodes,
                                       const double *leaves, const int64_t *tree_off, et_forest **out) {
  ET_API_BEGIN
  if (!ctx || !out || m < 0 || leaf_width <= 0 || total_nodes < 0 || total_leaves < 0 || !tree_off)
    ET_FAIL(ET_EINVAL, "et_forest_import_packed: bad argument");
  if (ctx->is_multi()) {
    et_forest *f0 = nullptr;
    int rc = et_forest_import_packed(ctx->peers[0], m, leaf_width, is_regression, total_nodes, total_leaves, nodes, leaves,
                                     tree_off, &f0);
    if (rc != ET_OK) return rc;
    et_forest *F = new et_forest();
    F->ctx = ctx;
    F->leaf_width = leaf_width;
    F->is_regression = is_regression;
    F->m = m;
    F->full = f0;
    F->total_nodes = f0->total_nodes;
    F->total_leaves = f0->total_leaves;
    F->tree_off = f0->tree_off;
    F->d_min = f0->d_min;
    *out = F;
    return ET_OK;
  }
  if ((total_nodes > 0 && !nodes) || (total_leaves > 0 && !leaves))
    ET_FAIL(ET_EINVAL, "et_forest_import_packed: NULL array");
  if (tree_off[0] != 0 || tree_off[m] != total_nodes) ET_FAIL(ET_EINVAL, "et_forest_import_packed: bad tree offsets");
  const PNode *pn = static_cast<const PNode *>(nodes);
  int32_t d_min = 0;
  for (int32_t t = 0; t < m; t++) {
    const int64_t off = tree_off[t], n = tree_off[t + 1] - off;
    if (n <= 0) ET_FAIL(ET_EINVAL, "et_forest_import_packed: tree %d is empty", t);
    for (int64_t i = 0; i < n; i++) {
      const PNode &q = pn[off + i];
      if (q.feat >= 0) {
        if (q.right_or_leaf <= i + 1 || q.right_or_leaf >= n)
          ET_FAIL(ET_EINVAL, "et_forest_import_packed: tree %d node %lld is not in pre-order", t, (long long)i);
        d_min = std::max(d_min, (q.feat & (ET_MIL_BIT - 1)) + 1);
      } else if (q.feat != -1) {
        ET_FAIL(ET_EINVAL, "et_forest_import_packed: tree %d node %lld: bad feature field", t, (long long)i);
      } else if (q.right_or_leaf < 0 || q.right_or_leaf >= total_leaves) {
        ET_FAIL(ET_EINVAL, "et_forest_import_packed: tree %d node %lld: leaf index out of range", t, (long long)i);
      }
    }
  }
  std::unique_ptr<et_forest> f(new et_forest());
  f->ctx = ctx;
  f->leaf_width = leaf_width;
  f->is_regression = is_regression;
  f->m = m;
  f->tree_off.assign(tree_off, tree_off + m + 1);
  f->total_nodes = total_nodes;
  f->total_leaves = total_leaves;
  f->d_min = d_min;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  CUDA_CHECK(cudaSetDevice(ctx->device));
  CUDA_CHECK(cudaMalloc((void **)&f->d_tree_off, ((size_t)m + 1) * sizeof(int64_t)));
  CUDA_CHECK(cudaMalloc((void **)&f->d_nodes, std::max<size_t>(1, (size_t)total_nodes) * sizeof(PNode)));
  CUDA_CHECK(cudaMalloc((void **)&f->d_leaf, std::max<size_t>(1, (size_t)total_leaves * leaf_width) * sizeof(double)));
  cudaStream_t st = ctx->stream;
  CUDA_CHECK(cudaMemcpyAsync(f->d_tree_off, tree_off, ((size_t)m + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  if (total_nodes)
    CUDA_CHECK(cudaMemcpyAsync(f->d_nodes, nodes, (size_t)total_nodes * sizeof(PNode), cudaMemcpyHostToDevice, st));
  if (total_leaves)
    CUDA_CHECK(cudaMemcpyAsync(f->d_leaf, leaves, (size_t)total_leaves * leaf_width * sizeof(double),
                               cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  *out = f.release();
  ET_API_END
}

// ---- predict --------------------------------------------------------------------------------
void et_predict_host_impl(et_ctx *ctx, et_forest *f, const double *x, int64_t n, int32_t d, double *out, int sum_only,
                          int want_regression);
static void predict_host(et_ctx *ctx, et_forest *f, const double *x, int64_t n, int32_t d, double *out,
                         int sum_only, int want_regression) {
  if (!ctx || !f || (!x && n > 0 && d > 0) || (!out && n > 0)) ET_FAIL(ET_EINVAL, "predict: NULL argument");
  if (n < 0 || d < 0) ET_FAIL(ET_EINVAL, "predict: negative dimensions");
  if (ctx->is_multi())
    et_multi_predict(ctx, f, x, n, d, out, sum_only, want_regression);
  else
    et_predict_host_impl(ctx, f, x, n, d, out, sum_only, want_regression);
}

void et_predict_host_impl(et_ctx *ctx, et_forest *f, const double *x, int64_t n, int32_t d, double *out, int sum_only,
                          int want_regression) {
  if (f->is_regression != want_regression) ET_FAIL(ET_EINVAL, "predict: forest kind does not match the call");
  if (n < 0 || d < 0) ET_FAIL(ET_EINVAL, "predict: negative dimensions");
  if (n > 0 && d < f->d_min)  // (ArrayIndexOutOfBounds in the reference, pkg:517)
    ET_FAIL(ET_EINVAL, "predict: samples have %d features, the forest splits on feature %d", d, f->d_min - 1);
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  CUDA_CHECK(cudaSetDevice(ctx->device));
  if (n == 0) return;
  int lw = f->leaf_width;
  // rows are streamed through the GPU in chunks: H2D, traverse, D2H
  int64_t chunk = std::max<int64_t>(1, ((int64_t)512 << 20) / ((int64_t)std::max(d, 1) * 8));
  chunk = std::min(chunk, n);
  const size_t dx_bytes = (size_t)chunk * std::max(d, 1) * sizeof(double), dout_bytes = (size_t)chunk * lw * sizeof(double);
  double *dx = static_cast<double *>(et_dev_alloc(ctx, dx_bytes));
  double *dout = static_cast<double *>(et_dev_alloc(ctx, dout_bytes));
  if (!dx || !dout) {
    et_dev_free(ctx, dx, dx_bytes);
    et_dev_free(ctx, dout, dout_bytes);
    ET_FAIL(ET_ENOMEM, "predict: cannot allocate staging buffers");
  }
  try {
    for (int64_t r0 = 0; r0 < n; r0 += chunk) {
      int64_t rows = std::min(chunk, n - r0);
      if (d > 0) et_h2d(ctx, dx, x + r0 * d, (size_t)rows * d * sizeof(double), ctx->stream);
      et_predict_device_impl(ctx, f, dx, rows, d, dout, sum_only);
      CUDA_CHECK(cudaMemcpyAsync(out + r0 * lw, dout, (size_t)rows * lw * sizeof(double), cudaMemcpyDeviceToHost,
                                 ctx->stream));
      CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
  } catch (...) {
    et_dev_free(ctx, dx, dx_bytes);
    et_dev_free(ctx, dout, dout_bytes);
    throw;
  }
  et_dev_free(ctx, dx, dx_bytes);
  et_dev_free(ctx, dout, dout_bytes);
}

extern "C" int et_predict_classification(et_ctx *ctx, et_forest *f, const double *x, int64_t n, int32_t d,
                                         double *out, int32_t sum_only) {
  ET_API_BEGIN
  predict_host(ctx, f, x, n, d, out, sum_only, 0);
  ET_API_END
}
extern "C" int et_predict_regression(et_ctx *ctx, et_forest *f, const double *x, int64_t n, int32_t d, double *out,
                                     int32_t sum_only) {
  ET_API_BEGIN
  predict_host(ctx, f, x, n, d, out, sum_only, 1);
  ET_API_END
}

static void predict_dev(et_ctx *ctx, et_forest *f, const double *x, int64_t n, int32_t d, double *out, int sum_only,
                        int want_regression) {
  if (!ctx || !f || (!x && n > 0 && d > 0) || (!out && n > 0)) ET_FAIL(ET_EINVAL, "predict: NULL argument");
  if (ctx->is_multi()) ET_FAIL(ET_EUNSUPPORTED, "device-pointer inputs need a single-GPU context");
  if (f->is_regression != want_regression) ET_FAIL(ET_EINVAL, "predict: forest kind does not match the call");
  if (n > 0 && d < f->d_min)
    ET_FAIL(ET_EINVAL, "predict: samples have %d features, the forest splits on feature %d", d, f->d_min - 1);
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  CUDA_CHECK(cudaSetDevice(ctx->device));
  if (n <= 0) return;
  et_predict_device_impl(ctx, f, x, n, d, out, sum_only);
}

extern "C" int et_predict_classification_device(et_ctx *ctx, et_forest *f, const double *x, int64_t n, int32_t d,
                                                double *out, int32_t sum_only) {
  ET_API_BEGIN
  predict_dev(ctx, f, x, n, d, out, sum_only, 0);
  ET_API_END
}
extern "C" int et_predict_regression_device(et_ctx *ctx, et_forest *f, const double *x, int64_t n, int32_t d,
                                            double *out, int32_t sum_only) {
  ET_API_BEGIN
  predict_dev(ctx, f, x, n, d, out, sum_only, 1);
  ET_API_END
}
