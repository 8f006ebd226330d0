// internal.h -- host-side objects behind the opaque handles of include/etgpu.h
#pragma once
#include "common.cuh"

struct Workspace;  // build.cu: device buffers reused across builds
struct HostStager;  // api.cu: pinned bounce buffers + copy threads for uploads from pageable host memory

struct et_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;  // the one work is queued on (own or caller's)
  int sm_count = 148;
  std::recursive_mutex mu;  // a context serialises its calls (recursive: handles are freed inside locked calls)
  int64_t launches = 0;
  Workspace *ws = nullptr;
  // side streams: the node kernels of one level (one launch per size class) run concurrently
  // freed device blocks of tables / forests kept for the next call of the same shape (exact-size reuse), so a
  // build-after-build loop does no cudaMalloc / cudaFree in the steady state
  struct Block {
    void *p;
    size_t bytes;
  };
  std::vector<Block> cache;
  size_t cache_bytes = 0;
  static constexpr int N_SIDE = 8;
  cudaStream_t side[N_SIDE] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[N_SIDE] = {};
  // multi-GPU (dist.cu): communicator of this GPU (ncclComm_t), its rank, a stream for collectives that overlap
  // with kernels, and -- on the front context returned by et_init_multi -- one child context per GPU
  void *comm = nullptr;
  int world = 1, rank = 0;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_comm = nullptr;
  double comm_ms = 0.0;
  std::vector<et_ctx *> peers;
  bool is_multi() const { return !peers.empty(); }
  HostStager *stager = nullptr;
  // et_data / et_forest handles still alive on this context: et_shutdown with handles outstanding is deferred until
  // the last of them is freed (a handle's destructor needs the context's mutex and block cache)
  int live_handles = 0;
  bool shutdown_requested = false;
};
void et_ctx_acquire(et_ctx *ctx);
void et_ctx_release(et_ctx *ctx);  // may run the deferred shutdown
// the context a handle lives on: counted, so the context outlives the handle
struct CtxHold {
  et_ctx *p = nullptr;
  CtxHold() = default;
  CtxHold(const CtxHold &) = delete;
  CtxHold &operator=(const CtxHold &) = delete;
  CtxHold &operator=(et_ctx *c) {
    if (c) et_ctx_acquire(c);
    if (p) et_ctx_release(p);
    p = c;
    return *this;
  }
  operator et_ctx *() const { return p; }
  et_ctx *operator->() const { return p; }
  ~CtxHold() {
    if (p) et_ctx_release(p);
  }
};
void et_workspace_free(Workspace *ws);
// api.cu: device blocks through the context's cache (null on failure, like cudaMalloc != cudaSuccess)
void *et_dev_alloc(et_ctx *ctx, size_t bytes);
void et_dev_free(et_ctx *ctx, void *p, size_t bytes);

struct et_data {
  CtxHold ctx;
  int64_t n = 0;
  int32_t d = 0;
  int64_t ld = 0;       // column stride in elements (n rounded up to 16)
  double *x = nullptr;  // column-major [d][ld], resident in HBM
  size_t x_bytes = 0;
  // order-preserving byte codes of the same table (encode.cu): byte + coff[col] = 0 for NaN, r + 1 for dict[col][r]
  int coded = 0;           // 0 not prepared yet, 1 codes valid, -1 not codable (> 256 distinct values in a column)
  uint8_t *c8 = nullptr;   // column-major [d][ldc]
  uint8_t *coff = nullptr; // [d] 0 if the column holds a NaN, else 1
  int64_t ldc = 0;         // code column stride in bytes (n rounded up to 128)
  double *dict = nullptr;  // [d][256] ascending distinct values, padded with +inf
  // row-major copy of the byte codes gathered by the lane-per-candidate kernel (k_lane): [n][rs8]; rows are 16-byte
  // multiples so a row moves with full vector loads
  uint8_t *r8 = nullptr;
  int64_t rs8 = 0;   // bytes per coded row
  size_t r8_bytes = 0;
  // CSC table kept sparse in HBM (x == null): ascending rows without duplicates inside every column
  int64_t *csc_colptr = nullptr;  // [d + 1]
  int32_t *csc_row = nullptr;
  double *csc_val = nullptr;
  int64_t csc_nnz = 0;
  // ... and the row-major index of the same entries (which features a row stores; the order inside a row is free)
  int64_t *csr_ptr = nullptr;  // [n + 1]
  int32_t *csr_col = nullptr;
  // attached targets / weights (resident)
  int32_t *y_cls = nullptr;
  int32_t num_classes = 0;
  double *y_reg = nullptr;
  double *w = nullptr;
  std::vector<int64_t> root_hist;  // class counts of the whole table
  std::vector<et_data *> shards;   // multi-GPU front handle: the replica on every GPU (nothing else is set)
};

// Device forest node: all trees concatenated in pre-order, one 16-byte node fetched with a single
// 128-bit load.  Pre-order makes the left child implicit (node + 1), so only the right child is
// stored (tree-local id).  For a leaf, feat == -1 and right_or_leaf is the forest-wide leaf index
// into the compact leaf table (leaf_width doubles per leaf, leaves numbered in pre-order).
// splitMissingIsLess rides in bit 30 of feat.
struct __align__(16) PNode {
  double cut;
  int32_t feat;
  int32_t right_or_leaf;
};
#define ET_MIL_BIT 0x40000000

struct et_forest {
  CtxHold ctx;
  int32_t leaf_width = 1;
  int32_t is_regression = 0;
  int32_t m = 0;
  int64_t total_nodes = 0, total_leaves = 0;
  int32_t d_min = 0;  // features a sample row must have: 1 + the largest split feature (the training d after a build)
  std::vector<int64_t> tree_off;  // m + 1, node offsets (host copy)
  // device-resident forest (what predict traverses)
  int64_t *d_tree_off = nullptr;  // m + 1
  PNode *d_nodes = nullptr;       // total_nodes
  double *d_leaf = nullptr;       // total_leaves x leaf_width
  size_t nodes_bytes = 0, leaf_bytes = 0, tree_off_bytes = 0;  // allocation sizes when the blocks came from et_dev_alloc (else 0)
  // multi-GPU: sort key of every tree when shards are gathered (global tree id / position in the call's forest)
  std::vector<int64_t> order_key;
  // multi-GPU front handle: the per-GPU forests of the trees each GPU built, and the gathered whole forest (on the
  // first GPU) that the export calls read
  std::vector<et_forest *> shards;
  et_forest *full = nullptr;
  // lazily fetched host copy (export)
  bool host_ready = false;
  std::vector<PNode> h_nodes;
  std::vector<double> h_leaf;
  ~et_forest();
};
void et_forest_fetch(et_forest *f);  // device -> host copy for export (api.cu)

struct BuildArgs {
  int task;  // 0 cls unweighted, 1 cls weighted, 2 regression
  int32_t num_classes, n_min, k, m, parallelism, best_split, max_depth;
  int64_t seed;
  const int32_t *tree_ids;
  const et_replay *replay;
  const int64_t *order_keys = nullptr;  // gather order of the trees (default: the global tree ids)
};

// encode.cu
static inline size_t et_coff_bytes(int32_t d) { return ((size_t)d + 15) / 16 * 16 + 16; }
void et_data_encode(et_ctx *ctx, et_data *data);
void et_data_drop_codes(et_data *data);  // gives the coded copy back to the context's block cache
void et_data_rowmajor(et_ctx *ctx, et_data *data);  // row-major copy gathered by k_lane
void et_data_drop_rowmajor(et_data *data);
// build.cu
void et_build_forest(et_ctx *ctx, et_data *data, const BuildArgs &a, et_forest *out, et_stats *stats);
// predict.cu
void et_predict_device_impl(et_ctx *ctx, et_forest *f, const double *x_dev, int64_t n, int32_t d,
                            double *out_dev, int sum_only);
// api.cu
// host -> device copy on `st`: pinned sources go straight to the copy engine; pageable ones (a JVM heap array, a
// numpy array) are staged through two pinned bounce buffers filled by a few copy threads while the previous piece
// is in flight, instead of the driver's single-threaded pageable path.  Returns after the source has been read.
void et_h2d(et_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes, cudaStream_t st);
void et_stager_free(HostStager *s);
et_data *et_data_alloc_internal(et_ctx *ctx, int64_t n, int32_t d);
void et_data_build_csr(et_ctx *ctx, et_data *D);  // row-major index of a sparse-resident (CSC) table
void et_predict_host_impl(et_ctx *ctx, et_forest *f, const double *x, int64_t n, int32_t d, double *out, int sum_only,
                          int want_regression);
// dist.cu: the multi-GPU front context (et_init_multi) behind the ordinary calls
void et_multi_shutdown(et_ctx *front);
void et_multi_replicate(et_ctx *front, et_data *front_data);  // shards[0] is filled: broadcast it to the other GPUs
void et_multi_broadcast_columns(et_ctx *front, et_data *front_data, int32_t first_col, int32_t n_cols);
void et_multi_build(et_ctx *front, et_data *D, const BuildArgs &a, int leaf_width, int is_regression, et_forest **out,
                    et_stats *stats);
void et_multi_predict(et_ctx *front, et_forest *f, const double *x, int64_t n, int32_t d, double *out, int sum_only,
                      int want_regression);
void et_launch_transpose(et_ctx *ctx, const double *src_rowmajor, int64_t rows, int32_t d,
                         double *dst_colmajor, int64_t ld, int64_t row0);
