// dist.cu -- trees sharded across the GPUs of one box; NCCL over NVLink only to replicate the table, gather the
// serialized trees and sum-reduce prediction partials.
//
// Replaces the JVM thread pool of buildForestClassification / buildForestRegression
// (parTraverseN(parallelism), pkg:653-675; result order = input order, pkg:656-675): trees are independent (a
// tree's random stream depends only on (seed, tree id)), so GPU g builds the trees t with t mod G == g and no
// collective runs during the build.
//
// Two front ends over the same per-rank code:
//   * et_init_multi: ONE process, one host thread per GPU, ncclCommInitAll -- what a JVM binding uses; every call
//     of include/etgpu.h works on the returned context (api.cu dispatches here)
//   * et_comm_init_rank: one PROCESS per GPU (torchrun / MPI), the launcher hands the NCCL unique id around
//
// libnccl.so.2 is loaded lazily with dlopen (the process may already hold a copy, e.g. PyTorch's): single-GPU
// use never needs it, multi-GPU calls fail with ET_ENCCL when it is missing.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <memory>
#include <thread>

#include "internal.h"

namespace {

struct NcclApi {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi &nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    NcclApi a;
    a.h = h;
#define ET_SYM(field, name)                                         \
  a.field = reinterpret_cast<decltype(a.field)>(dlsym(h, name));    \
  if (!a.field) return;
    ET_SYM(GetUniqueId, "ncclGetUniqueId")
    ET_SYM(CommInitRank, "ncclCommInitRank")
    ET_SYM(CommInitAll, "ncclCommInitAll")
    ET_SYM(CommDestroy, "ncclCommDestroy")
    ET_SYM(AllGather, "ncclAllGather")
    ET_SYM(AllReduce, "ncclAllReduce")
    ET_SYM(Broadcast, "ncclBroadcast")
    ET_SYM(GroupStart, "ncclGroupStart")
    ET_SYM(GroupEnd, "ncclGroupEnd")
    ET_SYM(GetErrorString, "ncclGetErrorString")
#undef ET_SYM
    api = a;
  });
  if (!api.h) ET_FAIL(ET_ENCCL, "libnccl.so.2 cannot be loaded (%s): multi-GPU calls need NCCL", dlerror() ? dlerror() : "missing symbol");
  return api;
}

#define NCCL_CHECK(expr)                                                                                    \
  do {                                                                                                      \
    ncclResult_t _r = (expr);                                                                               \
    if (_r != ncclSuccess && _r != ncclInProgress)                                                          \
      ET_FAIL(ET_ENCCL, "NCCL error at %s:%d: %s", __FILE__, __LINE__, nccl().GetErrorString(_r));          \
  } while (0)

static_assert(sizeof(ncclUniqueId) == ET_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");

ncclComm_t comm_of(et_ctx *ctx) {
  if (!ctx->comm) ET_FAIL(ET_EINVAL, "this context has no communicator (et_comm_init_rank / et_init_multi)");
  return static_cast<ncclComm_t>(ctx->comm);
}

void ensure_comm_stream(et_ctx *ctx) {
  if (!ctx->comm_stream) CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
  if (!ctx->ev_comm) CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_comm, cudaEventDisableTiming));
}

// device scratch from the context's block cache (no cudaMalloc / cudaFree in the steady state: a gather per build
// would otherwise pay several synchronising frees), given back on scope exit -- the caller has synchronised by then
struct DevTmp {
  et_ctx *ctx;
  std::vector<std::pair<void *, size_t>> p;
  explicit DevTmp(et_ctx *c) : ctx(c) {}
  template <typename T>
  T *alloc(size_t n) {
    const size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
    T *d = static_cast<T *>(et_dev_alloc(ctx, bytes));
    if (!d) ET_FAIL(ET_ENOMEM, "device allocation of %zu bytes failed", bytes);
    p.push_back({d, bytes});
    return d;
  }
  ~DevTmp() {
    for (auto &q : p) et_dev_free(ctx, q.first, q.second);
  }
};

struct EventPair {
  cudaEvent_t a = nullptr, b = nullptr;
  EventPair() {
    cudaEventCreate(&a);
    cudaEventCreate(&b);
  }
  ~EventPair() {
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
  }
};

// ---- gathered forest: staging (rank-major) -> tree order --------------------------------------------------------
struct TreeMove {
  int64_t src_node, dst_node, src_leaf, dst_leaf;  // offsets in the staging buffers / the gathered forest
  int64_t leaf_local;                              // first leaf of the tree in its shard's own leaf numbering
  int32_t n_nodes, pad_;
};

// one CTA per tree: nodes are copied with their leaf indices moved to the tree's place in the gathered leaf table
__global__ void __launch_bounds__(256) k_forest_permute(const TreeMove *__restrict__ mv, const PNode *__restrict__ snodes,
                                                        const double *__restrict__ sleaves, int lw, PNode *__restrict__ nodes,
                                                        double *__restrict__ leaves) {
  const TreeMove m = mv[blockIdx.x];
  const int64_t dl = m.dst_leaf - m.leaf_local;  // (a leaf node holds its index in the SHARD's leaf table)
  for (int32_t j = threadIdx.x; j < m.n_nodes; j += 256) {
    PNode pn = snodes[m.src_node + j];
    if (pn.feat < 0) pn.right_or_leaf = (int32_t)(pn.right_or_leaf + dl);
    nodes[m.dst_node + j] = pn;
  }
  const int64_t nl = ((int64_t)m.n_nodes + 1) / 2 * lw;  // a binary tree of n nodes has (n + 1) / 2 leaves
  for (int64_t j = threadIdx.x; j < nl; j += 256) leaves[m.dst_leaf * lw + j] = sleaves[m.src_leaf * lw + j];
}

__global__ void k_scale(double *v, int64_t n, double denom) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = ET_DIV(v[i], denom);
}

// per-rank part of the all-gather of serialized trees (collective: every rank of the communicator calls it)
et_forest *forest_allgather_rank(et_ctx *ctx, et_forest *shard) {
  NcclApi &N = nccl();
  ncclComm_t comm = comm_of(ctx);
  const int world = ctx->world, rank = ctx->rank;
  cudaStream_t st = ctx->stream;
  const int lw = shard->leaf_width;
  DevTmp tmp(ctx);
  EventPair ev;
  // ---- sizes: (trees, nodes, leaves, d_min) of every rank
  int64_t hdr[4] = {shard->m, shard->total_nodes, shard->total_leaves, shard->d_min};
  int64_t *d_hdr = tmp.alloc<int64_t>((size_t)4 * world);
  CUDA_CHECK(cudaMemcpyAsync(d_hdr + 4 * rank, hdr, sizeof(hdr), cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaEventRecord(ev.a, st));
  NCCL_CHECK(N.AllGather(d_hdr + 4 * rank, d_hdr, 4, ncclInt64, comm, st));
  std::vector<int64_t> all((size_t)4 * world);
  CUDA_CHECK(cudaMemcpyAsync(all.data(), d_hdr, all.size() * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  int64_t max_m = 0, m_tot = 0, nodes_tot = 0, leaves_tot = 0, d_min = 0, max_nodes = 0, max_leaves = 0;
  for (int r = 0; r < world; r++) {
    max_m = std::max(max_m, all[(size_t)4 * r]);
    m_tot += all[(size_t)4 * r];
    nodes_tot += all[(size_t)4 * r + 1];
    leaves_tot += all[(size_t)4 * r + 2];
    max_nodes = std::max(max_nodes, all[(size_t)4 * r + 1]);
    max_leaves = std::max(max_leaves, all[(size_t)4 * r + 2]);
    d_min = std::max(d_min, all[(size_t)4 * r + 3]);
  }
  // rank r's nodes / leaves land at slot r of the staging buffers (slots padded to the largest shard)
  std::vector<int64_t> node0((size_t)world + 1, 0), leaf0((size_t)world + 1, 0);
  for (int r = 0; r <= world; r++) {
    node0[(size_t)r] = (int64_t)r * max_nodes;
    leaf0[(size_t)r] = (int64_t)r * max_leaves;
  }
  if (m_tot > 0x7fffffff) ET_FAIL(ET_EUNSUPPORTED, "gathered forest exceeds 2^31 trees");
  // ---- per tree: (order key, nodes), padded to the largest shard
  std::vector<int64_t> meta((size_t)2 * std::max<int64_t>(max_m, 1), 0);
  for (int32_t t = 0; t < shard->m; t++) {
    meta[(size_t)2 * t] = (size_t)t < shard->order_key.size() ? shard->order_key[(size_t)t] : (int64_t)t;
    meta[(size_t)2 * t + 1] = shard->tree_off[(size_t)t + 1] - shard->tree_off[(size_t)t];
  }
  int64_t *d_meta = tmp.alloc<int64_t>(meta.size() * world);
  CUDA_CHECK(cudaMemcpyAsync(d_meta + meta.size() * rank, meta.data(), meta.size() * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  NCCL_CHECK(N.AllGather(d_meta + meta.size() * rank, d_meta, meta.size(), ncclInt64, comm, st));
  std::vector<int64_t> all_meta(meta.size() * world);
  CUDA_CHECK(cudaMemcpyAsync(all_meta.data(), d_meta, all_meta.size() * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  // ---- the packed 16-byte nodes and the leaf tables of every rank: two in-place all-gathers over padded slots
  PNode *s_nodes = tmp.alloc<PNode>((size_t)world * (size_t)std::max<int64_t>(max_nodes, 1));
  double *s_leaves = tmp.alloc<double>((size_t)world * (size_t)std::max<int64_t>(max_leaves, 1) * lw);
  if (shard->total_nodes)
    CUDA_CHECK(cudaMemcpyAsync(s_nodes + node0[(size_t)rank], shard->d_nodes, (size_t)shard->total_nodes * sizeof(PNode),
                               cudaMemcpyDeviceToDevice, st));
  if (shard->total_leaves)
    CUDA_CHECK(cudaMemcpyAsync(s_leaves + leaf0[(size_t)rank] * lw, shard->d_leaf,
                               (size_t)shard->total_leaves * lw * sizeof(double), cudaMemcpyDeviceToDevice, st));
  if (max_nodes)
    NCCL_CHECK(N.AllGather(s_nodes + node0[(size_t)rank], s_nodes, (size_t)max_nodes * sizeof(PNode), ncclChar, comm, st));
  if (max_leaves)
    NCCL_CHECK(N.AllGather(s_leaves + leaf0[(size_t)rank] * lw, s_leaves, (size_t)max_leaves * lw * sizeof(double), ncclChar,
                           comm, st));
  CUDA_CHECK(cudaEventRecord(ev.b, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  // ---- plan: trees in order of their key
  struct Src {
    int64_t key, node, leaf, leaf_local;
    int32_t n_nodes;
  };
  std::vector<Src> src;
  src.reserve((size_t)m_tot);
  for (int r = 0; r < world; r++) {
    int64_t no = node0[(size_t)r], lo = leaf0[(size_t)r];
    for (int64_t t = 0; t < all[(size_t)4 * r]; t++) {
      const int64_t key = all_meta[meta.size() * r + (size_t)2 * t], nn = all_meta[meta.size() * r + (size_t)2 * t + 1];
      src.push_back(Src{key, no, lo, lo - leaf0[(size_t)r], (int32_t)nn});
      no += nn;
      lo += (nn + 1) / 2;
    }
  }
  std::stable_sort(src.begin(), src.end(), [](const Src &a, const Src &b) { return a.key < b.key; });
  std::unique_ptr<et_forest> f(new et_forest());
  f->ctx = ctx;
  f->leaf_width = lw;
  f->is_regression = shard->is_regression;
  f->m = (int32_t)m_tot;
  f->d_min = (int32_t)d_min;
  f->total_nodes = nodes_tot;
  f->total_leaves = leaves_tot;
  f->tree_off.assign((size_t)m_tot + 1, 0);
  f->order_key.resize((size_t)m_tot);
  std::vector<TreeMove> mv((size_t)m_tot);
  int64_t dn = 0, dl = 0;
  for (size_t t = 0; t < src.size(); t++) {
    mv[t] = TreeMove{src[t].node, dn, src[t].leaf, dl, src[t].leaf_local, src[t].n_nodes, 0};
    f->tree_off[t] = dn;
    f->order_key[t] = src[t].key;
    dn += src[t].n_nodes;
    dl += (src[t].n_nodes + 1) / 2;
  }
  f->tree_off[(size_t)m_tot] = dn;
  if (dn != nodes_tot || dl != leaves_tot) ET_FAIL(ET_ECUDA, "gathered forest: node / leaf counts do not add up");
  f->nodes_bytes = std::max<size_t>(1, (size_t)nodes_tot) * sizeof(PNode);
  f->leaf_bytes = std::max<size_t>(1, (size_t)leaves_tot * lw) * sizeof(double);
  f->d_nodes = static_cast<PNode *>(et_dev_alloc(ctx, f->nodes_bytes));
  f->d_leaf = static_cast<double *>(et_dev_alloc(ctx, f->leaf_bytes));
  if (!f->d_nodes || !f->d_leaf) ET_FAIL(ET_ENOMEM, "cannot allocate the gathered forest (%lld nodes)", (long long)nodes_tot);
  CUDA_CHECK(cudaMalloc((void **)&f->d_tree_off, ((size_t)m_tot + 1) * sizeof(int64_t)));
  CUDA_CHECK(cudaMemcpyAsync(f->d_tree_off, f->tree_off.data(), ((size_t)m_tot + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  if (m_tot > 0) {
    TreeMove *d_mv = tmp.alloc<TreeMove>(mv.size());
    CUDA_CHECK(cudaMemcpyAsync(d_mv, mv.data(), mv.size() * sizeof(TreeMove), cudaMemcpyHostToDevice, st));
    k_forest_permute<<<(unsigned)m_tot, 256, 0, st>>>(d_mv, s_nodes, s_leaves, lw, f->d_nodes, f->d_leaf);
    ctx->launches++;
  }
  CUDA_CHECK(cudaStreamSynchronize(st));
  CUDA_CHECK(cudaGetLastError());
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ev.a, ev.b);
  ctx->comm_ms = ms;
  return f.release();
}

// per-rank part of the tree-sharded predict on device buffers: partial sums over this rank's trees, all-reduced in
// row chunks on the communication stream while the next chunk is traversed, one division by m_total
void predict_allreduce_rank(et_ctx *ctx, et_forest *shard, const double *x, int64_t n, int32_t d, double *out,
                            int32_t m_total) {
  NcclApi &N = nccl();
  ncclComm_t comm = comm_of(ctx);
  ensure_comm_stream(ctx);
  const int lw = shard->leaf_width;
  EventPair ev;
  std::vector<cudaEvent_t> done;
  const int64_t chunk = std::max<int64_t>(1024, ((int64_t)32 << 20) / ((int64_t)lw * 8));  // ~32 MB of partial sums
  CUDA_CHECK(cudaEventRecord(ctx->ev_comm, ctx->stream));
  CUDA_CHECK(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_comm, 0));  // (earlier work on `out`)
  CUDA_CHECK(cudaEventRecord(ev.a, ctx->comm_stream));
  for (int64_t r0 = 0; r0 < n; r0 += chunk) {
    const int64_t rows = std::min(chunk, n - r0);
    if (shard->m > 0)
      et_predict_device_impl(ctx, shard, x + r0 * d, rows, d, out + r0 * lw, 1);
    else
      CUDA_CHECK(cudaMemsetAsync(out + r0 * lw, 0, (size_t)rows * lw * sizeof(double), ctx->stream));
    cudaEvent_t e;
    CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    done.push_back(e);
    CUDA_CHECK(cudaEventRecord(e, ctx->stream));
    CUDA_CHECK(cudaStreamWaitEvent(ctx->comm_stream, e, 0));
    NCCL_CHECK(N.AllReduce(out + r0 * lw, out + r0 * lw, (size_t)rows * lw, ncclDouble, ncclSum, comm, ctx->comm_stream));
    k_scale<<<(unsigned)ceil_div(rows * lw, 256), 256, 0, ctx->comm_stream>>>(out + r0 * lw, rows * lw, (double)m_total);
    ctx->launches++;
  }
  CUDA_CHECK(cudaEventRecord(ev.b, ctx->comm_stream));
  CUDA_CHECK(cudaEventRecord(ctx->ev_comm, ctx->comm_stream));
  CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_comm, 0));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  for (auto e : done) cudaEventDestroy(e);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ev.a, ev.b);
  ctx->comm_ms = ms;  // (span of the communication stream: all-reduces and their waits for the traversal)
}

// table replica on this rank: the root's column-major FP64 table arrives over NVLink
et_data *data_broadcast_rank(et_ctx *ctx, et_data *data, int32_t root) {
  NcclApi &N = nccl();
  ncclComm_t comm = comm_of(ctx);
  cudaStream_t st = ctx->stream;
  DevTmp tmp(ctx);
  int64_t hdr[3] = {data ? data->n : 0, data ? data->d : 0, (data && !data->x) ? data->csc_nnz : -1};  // nnz < 0: dense
  int64_t *d_hdr = tmp.alloc<int64_t>(3);
  if (ctx->rank == root) {
    if (!data) ET_FAIL(ET_EINVAL, "et_data_broadcast: the root rank must pass its table");
    CUDA_CHECK(cudaMemcpyAsync(d_hdr, hdr, sizeof(hdr), cudaMemcpyHostToDevice, st));
  }
  NCCL_CHECK(N.Broadcast(d_hdr, d_hdr, 3, ncclInt64, root, comm, st));
  CUDA_CHECK(cudaMemcpyAsync(hdr, d_hdr, sizeof(hdr), cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  if (hdr[2] >= 0) {
    // a table kept sparse: the three CSC arrays travel, the row-major index is rebuilt locally
    if (ctx->rank == root) {
      if (hdr[1] > 0) NCCL_CHECK(N.Broadcast(data->csc_colptr, data->csc_colptr, (size_t)(hdr[1] + 1) * 8, ncclChar, root, comm, st));
      if (hdr[2] > 0) {
        NCCL_CHECK(N.Broadcast(data->csc_row, data->csc_row, (size_t)hdr[2] * 4, ncclChar, root, comm, st));
        NCCL_CHECK(N.Broadcast(data->csc_val, data->csc_val, (size_t)hdr[2] * 8, ncclChar, root, comm, st));
      }
      CUDA_CHECK(cudaStreamSynchronize(st));
      return data;
    }
    et_data *S = new et_data();
    S->ctx = ctx;
    S->n = hdr[0];
    S->d = (int32_t)hdr[1];
    S->ld = ((hdr[0] + 15) / 16) * 16;
    S->coded = -1;
    S->csc_nnz = hdr[2];
    try {
      CUDA_CHECK(cudaMalloc((void **)&S->csc_colptr, ((size_t)hdr[1] + 1) * sizeof(int64_t)));
      CUDA_CHECK(cudaMalloc((void **)&S->csc_row, std::max<size_t>(1, (size_t)hdr[2]) * sizeof(int32_t)));
      CUDA_CHECK(cudaMalloc((void **)&S->csc_val, std::max<size_t>(1, (size_t)hdr[2]) * sizeof(double)));
      if (hdr[1] > 0) NCCL_CHECK(N.Broadcast(S->csc_colptr, S->csc_colptr, (size_t)(hdr[1] + 1) * 8, ncclChar, root, comm, st));
      if (hdr[2] > 0) {
        NCCL_CHECK(N.Broadcast(S->csc_row, S->csc_row, (size_t)hdr[2] * 4, ncclChar, root, comm, st));
        NCCL_CHECK(N.Broadcast(S->csc_val, S->csc_val, (size_t)hdr[2] * 8, ncclChar, root, comm, st));
      }
      et_data_build_csr(ctx, S);
    } catch (...) {
      et_data_free(S);
      throw;
    }
    return S;
  }
  et_data *D = (ctx->rank == root) ? data : et_data_alloc_internal(ctx, hdr[0], (int32_t)hdr[1]);
  try {
    const size_t bytes = (size_t)D->ld * (size_t)D->d * sizeof(double);
    if (bytes) NCCL_CHECK(N.Broadcast(D->x, D->x, bytes, ncclChar, root, comm, st));
    if (ctx->rank != root) et_data_encode(ctx, D);
    CUDA_CHECK(cudaStreamSynchronize(st));
  } catch (...) {
    if (ctx->rank != root) et_data_free(D);
    throw;
  }
  return D;
}

// runs fn(g) on one host thread per GPU of the front context; the first failure is re-raised on the caller's thread
template <typename Fn>
void for_each_peer(et_ctx *front, Fn fn) {
  const int G = (int)front->peers.size();
  std::vector<int> code((size_t)G, ET_OK);
  std::vector<std::string> msg((size_t)G);
  std::vector<std::thread> th;
  for (int g = 0; g < G; g++) {
    th.emplace_back([&, g] {
      try {
        et_ctx *c = front->peers[(size_t)g];
        cudaSetDevice(c->device);
        std::lock_guard<std::recursive_mutex> lk(c->mu);
        fn(g);
      } catch (const EtError &e) {
        code[(size_t)g] = e.code;
        msg[(size_t)g] = et_last_error();
      } catch (const std::bad_alloc &) {
        code[(size_t)g] = ET_ENOMEM;
        msg[(size_t)g] = "host allocation failed";
      }
    });
  }
  for (auto &t : th) t.join();
  for (int g = 0; g < G; g++)
    if (code[(size_t)g] != ET_OK) ET_FAIL(code[(size_t)g], "GPU %d: %s", front->peers[(size_t)g]->device, msg[(size_t)g].c_str());
}

}  // namespace

// ---- one process, several GPUs ---------------------------------------------------------------------------------------
extern "C" int et_init_multi(const int32_t *devices, int32_t n_devices, et_ctx **out) {
  try {
    if (!out || !devices || n_devices <= 0) ET_FAIL(ET_EINVAL, "et_init_multi: bad argument");
    for (int i = 0; i < n_devices; i++)
      for (int j = 0; j < i; j++)
        if (devices[i] == devices[j]) ET_FAIL(ET_EINVAL, "et_init_multi: device %d listed twice", devices[i]);
    NcclApi &N = nccl();
    std::unique_ptr<et_ctx> front(new et_ctx());
    front->device = devices[0];
    front->world = n_devices;
    std::vector<int> devs(devices, devices + n_devices);
    try {
      for (int g = 0; g < n_devices; g++) {
        et_ctx *c = nullptr;
        int rc = et_init(devices[g], &c);
        if (rc != ET_OK) throw EtError{rc};
        c->world = n_devices;
        c->rank = g;
        front->peers.push_back(c);
      }
      std::vector<ncclComm_t> comms((size_t)n_devices);
      NCCL_CHECK(N.CommInitAll(comms.data(), n_devices, devs.data()));
      for (int g = 0; g < n_devices; g++) front->peers[(size_t)g]->comm = comms[(size_t)g];
    } catch (...) {
      et_multi_shutdown(front.get());
      throw;
    }
    *out = front.release();
  } catch (const EtError &e) {
    return e.code;
  } catch (const std::bad_alloc &) {
    et_set_error("host allocation failed");
    return ET_ENOMEM;
  }
  return ET_OK;
}

extern "C" int32_t et_device_count(const et_ctx *ctx) { return ctx ? (ctx->is_multi() ? (int32_t)ctx->peers.size() : 1) : 0; }

void et_multi_shutdown(et_ctx *front) {
  for (et_ctx *c : front->peers) {
    if (c->comm) {
      cudaSetDevice(c->device);
      nccl().CommDestroy(static_cast<ncclComm_t>(c->comm));
      c->comm = nullptr;
    }
    et_shutdown(c);
  }
  front->peers.clear();
}

// shards[0] holds the uploaded table: replicate it on the other GPUs (ncclBroadcast of the column-major matrix)
void et_multi_replicate(et_ctx *front, et_data *F) {
  const int G = (int)front->peers.size();
  et_data *d0 = F->shards[0];
  F->shards.resize((size_t)G, nullptr);
  try {
    for_each_peer(front, [&](int g) {
      et_ctx *c = front->peers[(size_t)g];
      et_data *r = data_broadcast_rank(c, g == 0 ? d0 : nullptr, 0);
      F->shards[(size_t)g] = r;
    });
  } catch (...) {
    for (int g = 1; g < G; g++)
      if (F->shards[(size_t)g]) et_data_free(F->shards[(size_t)g]);
    F->shards.resize(1);
    throw;
  }
}

// columns uploaded to the first GPU (et_data_dense_colblock) travel to the others
void et_multi_broadcast_columns(et_ctx *front, et_data *F, int32_t first_col, int32_t n_cols) {
  NcclApi &N = nccl();
  for_each_peer(front, [&](int g) {
    et_ctx *c = front->peers[(size_t)g];
    et_data *D = F->shards[(size_t)g];
    const size_t bytes = (size_t)D->ld * (size_t)n_cols * sizeof(double);
    double *blk = D->x + (int64_t)first_col * D->ld;
    if (bytes) NCCL_CHECK(N.Broadcast(blk, blk, bytes, ncclChar, 0, comm_of(c), c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (g != 0 && D->coded != 0) {  // the coded copy is rebuilt at the next build
      et_data_drop_codes(D);
      D->coded = 0;
    }
  });
}

void et_multi_build(et_ctx *front, et_data *F, const BuildArgs &a, int leaf_width, int is_regression, et_forest **out,
                    et_stats *stats) {
  if (a.replay) ET_FAIL(ET_EUNSUPPORTED, "the replay test hook is single-GPU only");
  const int G = (int)front->peers.size();
  if ((int)F->shards.size() != G) ET_FAIL(ET_EINVAL, "build: the table is not resident on every GPU of this context");
  // tree t of the call -> GPU t mod G (balances the depth lottery better than contiguous blocks)
  std::vector<std::vector<int32_t>> ids((size_t)G);
  std::vector<std::vector<int64_t>> keys((size_t)G);
  for (int32_t t = 0; t < a.m; t++) {
    ids[(size_t)(t % G)].push_back(a.tree_ids ? a.tree_ids[t] : t);
    keys[(size_t)(t % G)].push_back(t);
  }
  std::unique_ptr<et_forest> FF(new et_forest());
  FF->ctx = front;
  FF->leaf_width = leaf_width;
  FF->is_regression = is_regression;
  FF->m = a.m;
  FF->shards.assign((size_t)G, nullptr);
  std::vector<et_stats> st((size_t)G);
  std::vector<et_forest *> fulls((size_t)G, nullptr);
  try {
    for_each_peer(front, [&](int g) {
      et_ctx *c = front->peers[(size_t)g];
      BuildArgs b = a;
      b.m = (int32_t)ids[(size_t)g].size();
      b.tree_ids = ids[(size_t)g].data();
      b.order_keys = keys[(size_t)g].data();
      std::unique_ptr<et_forest> f(new et_forest());
      f->ctx = c;
      f->leaf_width = leaf_width;
      f->is_regression = is_regression;
      f->m = b.m;
      memset(&st[(size_t)g], 0, sizeof(et_stats));
      if (b.m > 0) {
        et_build_forest(c, F->shards[(size_t)g], b, f.get(), &st[(size_t)g]);
      } else {  // more GPUs than trees: an empty shard still takes part in the collectives
        f->tree_off.assign(1, 0);
        CUDA_CHECK(cudaMalloc((void **)&f->d_tree_off, sizeof(int64_t)));
        CUDA_CHECK(cudaMemset(f->d_tree_off, 0, sizeof(int64_t)));
      }
      FF->shards[(size_t)g] = f.release();
      fulls[(size_t)g] = forest_allgather_rank(c, FF->shards[(size_t)g]);
    });
  } catch (...) {
    for (auto *f : fulls) delete f;
    for (auto *f : FF->shards) delete f;
    FF->shards.clear();
    throw;
  }
  FF->full = fulls[0];  // (every GPU received the whole forest; the first GPU's copy serves the export calls)
  for (int g = 1; g < G; g++) delete fulls[(size_t)g];
  FF->total_nodes = FF->full->total_nodes;
  FF->total_leaves = FF->full->total_leaves;
  FF->tree_off = FF->full->tree_off;
  FF->d_min = FF->full->d_min;
  if (stats) {
    et_stats S;
    memset(&S, 0, sizeof(S));
    for (int g = 0; g < G; g++) {
      const et_stats &s = st[(size_t)g];
      S.v_mm += s.v_mm;
      S.v_sc += s.v_sc;
      S.s_rows += s.s_rows;
      S.p_rows += s.p_rows;
      S.draws += s.draws;
      S.const_hits += s.const_hits;
      S.scored += s.scored;
      S.nodes += s.nodes;
      S.levels = std::max(S.levels, s.levels);
      S.rounds += s.rounds;
      S.launches += s.launches;
      S.parallel_sum_nodes += s.parallel_sum_nodes;
      S.ambiguous_splits += s.ambiguous_splits;
      S.gpu_ms = std::max(S.gpu_ms, s.gpu_ms);  // the GPUs run side by side
      S.gpu_ms_split = std::max(S.gpu_ms_split, s.gpu_ms_split);
      S.gpu_ms_partition = std::max(S.gpu_ms_partition, s.gpu_ms_partition);
    }
    *stats = S;
  }
  double cms = 0.0;
  for (et_ctx *c : front->peers) cms = std::max(cms, c->comm_ms);
  front->comm_ms = cms;
  *out = FF.release();
}

void et_multi_predict(et_ctx *front, et_forest *f, const double *x, int64_t n, int32_t d, double *out, int sum_only,
                      int want_regression) {
  if (f->is_regression != want_regression) ET_FAIL(ET_EINVAL, "predict: forest kind does not match the call");
  if (n > 0 && d < f->d_min)
    ET_FAIL(ET_EINVAL, "predict: samples have %d features, the forest splits on feature %d", d, f->d_min - 1);
  const int G = (int)front->peers.size();
  if ((int)f->shards.size() != G) {  // an imported forest lives on the first GPU only
    if (!f->full) ET_FAIL(ET_EINVAL, "predict: empty forest handle");
    et_predict_host_impl(front->peers[0], f->full, x, n, d, out, sum_only, want_regression);
    return;
  }
  if (n <= 0) return;
  NcclApi &N = nccl();
  const int lw = f->leaf_width;
  // rows are streamed in chunks: host -> first GPU, NVLink broadcast, traversal of each GPU's trees, all-reduce
  const int64_t chunk = std::min<int64_t>(n, std::max<int64_t>(1, ((int64_t)256 << 20) / ((int64_t)std::max(d, 1) * 8)));
  double cms = 0.0;
  std::vector<double> peer_ms((size_t)G, 0.0);
  for_each_peer(front, [&](int g) {
    et_ctx *c = front->peers[(size_t)g];
    const size_t xb = (size_t)chunk * std::max(d, 1) * sizeof(double), ob = (size_t)chunk * lw * sizeof(double);
    double *dx = static_cast<double *>(et_dev_alloc(c, xb));
    double *dout = static_cast<double *>(et_dev_alloc(c, ob));
    if (!dx || !dout) {
      et_dev_free(c, dx, xb);
      et_dev_free(c, dout, ob);
      ET_FAIL(ET_ENOMEM, "predict: cannot allocate staging buffers");
    }
    try {
      for (int64_t r0 = 0; r0 < n; r0 += chunk) {
        const int64_t rows = std::min(chunk, n - r0);
        if (g == 0 && d > 0)
          CUDA_CHECK(cudaMemcpyAsync(dx, x + r0 * d, (size_t)rows * d * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        if (d > 0) NCCL_CHECK(N.Broadcast(dx, dx, (size_t)rows * d * sizeof(double), ncclChar, 0, comm_of(c), c->stream));
        predict_allreduce_rank(c, f->shards[(size_t)g], dx, rows, d, dout, sum_only ? 1 : f->m);
        peer_ms[(size_t)g] += c->comm_ms;
        if (g == 0) {
          CUDA_CHECK(cudaMemcpyAsync(out + r0 * lw, dout, (size_t)rows * lw * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
          CUDA_CHECK(cudaStreamSynchronize(c->stream));
        }
      }
    } catch (...) {
      et_dev_free(c, dx, xb);
      et_dev_free(c, dout, ob);
      throw;
    }
    et_dev_free(c, dx, xb);
    et_dev_free(c, dout, ob);
  });
  for (double v : peer_ms) cms = std::max(cms, v);
  front->comm_ms = cms;
}

// ---- one process per GPU ---------------------------------------------------------------------------------------------
#define ET_API_BEGIN try {
#define ET_API_END                                           \
  }                                                          \
  catch (const EtError &e) { return e.code; }                \
  catch (const std::bad_alloc &) {                           \
    et_set_error("host allocation failed");                  \
    return ET_ENOMEM;                                        \
  }                                                          \
  return ET_OK;

extern "C" int et_comm_unique_id(uint8_t *id_out) {
  ET_API_BEGIN
  if (!id_out) ET_FAIL(ET_EINVAL, "et_comm_unique_id: NULL argument");
  ncclUniqueId id;
  NCCL_CHECK(nccl().GetUniqueId(&id));
  memcpy(id_out, &id, ET_COMM_ID_BYTES);
  ET_API_END
}

extern "C" int et_comm_init_rank(et_ctx *ctx, int32_t world, int32_t rank, const uint8_t *id) {
  ET_API_BEGIN
  if (!ctx || !id || world <= 0 || rank < 0 || rank >= world) ET_FAIL(ET_EINVAL, "et_comm_init_rank: bad argument");
  if (ctx->is_multi()) ET_FAIL(ET_EINVAL, "et_comm_init_rank: the context already drives several GPUs");
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  CUDA_CHECK(cudaSetDevice(ctx->device));
  if (ctx->comm) {
    nccl().CommDestroy(static_cast<ncclComm_t>(ctx->comm));
    ctx->comm = nullptr;
  }
  ncclUniqueId uid;
  memcpy(&uid, id, ET_COMM_ID_BYTES);
  ncclComm_t comm;
  NCCL_CHECK(nccl().CommInitRank(&comm, world, uid, rank));
  ctx->comm = comm;
  ctx->world = world;
  ctx->rank = rank;
  ET_API_END
}

extern "C" int et_data_broadcast(et_ctx *ctx, et_data *data, int32_t root, et_data **out) {
  ET_API_BEGIN
  if (!ctx || !out || root < 0 || root >= ctx->world) ET_FAIL(ET_EINVAL, "et_data_broadcast: bad argument");
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  CUDA_CHECK(cudaSetDevice(ctx->device));
  *out = data_broadcast_rank(ctx, data, root);
  ET_API_END
}

extern "C" int et_forest_allgather(et_ctx *ctx, et_forest *shard, et_forest **full) {
  ET_API_BEGIN
  if (!ctx || !shard || !full) ET_FAIL(ET_EINVAL, "et_forest_allgather: NULL argument");
  if (shard->ctx != ctx) ET_FAIL(ET_EINVAL, "et_forest_allgather: the forest belongs to another context");
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  CUDA_CHECK(cudaSetDevice(ctx->device));
  *full = forest_allgather_rank(ctx, shard);
  ET_API_END
}

static int predict_allreduce_api(et_ctx *ctx, et_forest *shard, const double *x, int64_t n, int32_t d, double *out,
                                 int32_t m_total, int want_regression) {
  ET_API_BEGIN
  if (!ctx || !shard || (!x && n > 0 && d > 0) || (!out && n > 0) || m_total <= 0)
    ET_FAIL(ET_EINVAL, "predict: bad argument");
  if (shard->is_regression != want_regression) ET_FAIL(ET_EINVAL, "predict: forest kind does not match the call");
  if (n > 0 && d < shard->d_min)
    ET_FAIL(ET_EINVAL, "predict: samples have %d features, the forest splits on feature %d", d, shard->d_min - 1);
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  CUDA_CHECK(cudaSetDevice(ctx->device));
  if (n > 0) predict_allreduce_rank(ctx, shard, x, n, d, out, m_total);
  ET_API_END
}

extern "C" int et_predict_classification_allreduce(et_ctx *ctx, et_forest *shard, const double *x, int64_t n, int32_t d,
                                                   double *out, int32_t m_total) {
  return predict_allreduce_api(ctx, shard, x, n, d, out, m_total, 0);
}
extern "C" int et_predict_regression_allreduce(et_ctx *ctx, et_forest *shard, const double *x, int64_t n, int32_t d,
                                               double *out, int32_t m_total) {
  return predict_allreduce_api(ctx, shard, x, n, d, out, m_total, 1);
}

extern "C" double et_comm_last_ms(const et_ctx *ctx) { return ctx ? ctx->comm_ms : 0.0; }
