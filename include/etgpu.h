/*
 * etgpu.h -- C ABI of libetgpu.so: B200-native (sm_100a) replacement for the extratrees hot path of
 * pityka/lamp (package lamp.extratrees).
 *
 * The reference has no FFI on this path: its boundary is the public Scala surface in
 * extratrees/src/main/scala/lamp/forest/package.scala (cited pkg:LINE):
 *     buildForestClassification  pkg:611-681      predictClassification  pkg:542-551
 *     buildForestRegression      pkg:704-764      predictRegression      pkg:577-586
 * and the tree ADTs of extratrees.scala:3-63.  Each entry point below names the reference
 * interface it replaces.  A Scala facade keeps those four signatures and calls this ABI through
 * JNI/Panama (INTEGRATION.md shows the binding); lamp_b200/extratrees.py is the same facade in
 * Python over ctypes.
 *
 * Conventions: plain pointers and sizes only.  Every call returns ET_OK or a negative error code;
 * et_last_error() returns the message of the calling thread's last failure.  Host pointers are
 * borrowed for the duration of the call.  Handles are owned by the library until *_free.  A context
 * from et_init drives ONE GPU; a context from et_init_multi drives several GPUs of one box from one
 * process (trees sharded by tree id, NCCL over NVLink inside the library), behind the same calls.
 * There is NO CPU fallback: without a CUDA device et_init fails.
 */
#ifndef ETGPU_H
#define ETGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ET_OK 0
#define ET_EINVAL (-1)   /* the reference's require(...) -> IllegalArgumentException (pkg:624-633,715-718,779) */
#define ET_ECUDA (-2)    /* CUDA runtime failure / no device */
#define ET_ENOMEM (-3)   /* device or host allocation failed */
#define ET_EREPLAY (-4)  /* replay trace does not fit the data (test hook misuse) */
#define ET_EUNSUPPORTED (-5)
#define ET_ENCCL (-6)    /* NCCL failure / libnccl.so.2 not loadable (multi-GPU calls only) */

typedef struct et_ctx et_ctx;
typedef struct et_data et_data;
typedef struct et_forest et_forest;

/* ABI version of this header (checked by the bindings). */
int32_t et_abi_version(void);
const char *et_last_error(void);

/* ---- context ------------------------------------------------------------------------------ */
/* One context per GPU.  Replaces the reference's hidden cats-effect global runtime (pkg:6,656,675). */
int et_init(int32_t device, et_ctx **out);
/* Several GPUs of one box behind ONE context (replaces parTraverseN(parallelism), pkg:653-675: the fan-out
 * stays hidden inside the call).  One host thread per GPU, ncclCommInitAll over the given devices.  Every call
 * of this header then works on the multi-GPU context:
 *   et_data_*            one host->device upload to the first GPU, ncclBroadcast over NVLink to the others
 *                        (every GPU holds a replica of the table); targets / weights go to every GPU
 *   et_build_*           tree t of the forest is built by GPU t mod G (a tree's random stream depends only on
 *                        (seed, tree id): the forest equals the one a single GPU builds); the serialized trees
 *                        are all-gathered (ncclAllGather of the sizes, then of the packed nodes over padded slots),
 *                        so every GPU -- and the host, through et_forest_export* -- holds the whole forest
 *   et_predict_*         trees stay sharded: every GPU traverses all rows for its trees, the per-row partial
 *                        sums are all-reduced (ncclAllReduce, FP64 sum) and divided by m once.  The sum is
 *                        re-associated across GPUs: results agree with the single-GPU / reference value to
 *                        ~1e-15 relative (inside the 1e-12 bar), not bit for bit.
 * The replay test hook is single-GPU only (ET_EUNSUPPORTED here). */
int et_init_multi(const int32_t *devices, int32_t n_devices, et_ctx **out);
int32_t et_device_count(const et_ctx *ctx);
void et_shutdown(et_ctx *ctx);
/* Run all work of this context on the caller's CUDA stream (cudaStream_t passed as void*); NULL
 * restores the context's own stream.  Lets a host framework time the library with its own events. */
int et_set_stream(et_ctx *ctx, void *cuda_stream);
/* Blocks until everything queued by this context has finished. */
int et_synchronize(et_ctx *ctx);

/* ---- data (replaces `data: Mat[Double]`, pkg:612,705; row-major N x d, pkg:936) -------------- */
/* Upload a saddle-layout (row-major) host matrix; it is transposed on the device into the
 * column-major FP64 matrix the builder gathers from. */
int et_data_dense_rowmajor(et_ctx *ctx, const double *x_host, int64_t n, int32_t d, et_data **out);
/* For tables beyond a JVM array (2^31 elements): allocate, then upload blocks of whole columns
 * (column-major host block: n_cols columns of n doubles each). */
int et_data_dense_alloc(et_ctx *ctx, int64_t n, int32_t d, et_data **out);
int et_data_dense_colblock(et_ctx *ctx, et_data *data, const double *cols_host, int32_t first_col,
                           int32_t n_cols);
/* Same, from a device-resident row-major matrix (no host copy; used to keep inputs in HBM). */
int et_data_dense_rowmajor_device(et_ctx *ctx, const double *x_dev, int64_t n, int32_t d,
                                  et_data **out);
/* Sparse input in compressed-sparse-column form (BASELINE.json configs[3]): column c holds the entries
 * colptr[c] .. colptr[c+1]-1 of (rowidx, val); entries not listed are 0.0 with DENSE semantics (the reference
 * has no sparse Mat: a CSC table builds exactly the forest of its dense expansion, pkg:34-54, 931-941; stored
 * NaNs are missing values like anywhere else).  A table whose dense form is small (ETGPU_CSC_DENSE_MAX bytes,
 * default 4 GiB) is expanded on the device into the resident column-major matrix (and byte-coded like any other
 * table when every column has <= 255 distinct values).  A larger table STAYS SPARSE in HBM -- 12 bytes per stored
 * entry, 1.2 GB for 1M x 10000 at 1 % instead of 80 GB: the kernels find the value of (row, column) by binary
 * search among the column's stored rows, a miss is an implicit zero.  Row indices must lie in [0, n), colptr must
 * be non-decreasing with colptr[0] == 0, else ET_EINVAL.  Columns need not be sorted (they are sorted on the
 * host); a row listed twice in one column keeps the LATER entry. */
int et_data_csc(et_ctx *ctx, const int64_t *colptr_host, const int32_t *rowidx_host, const double *val_host,
                int64_t n, int32_t d, et_data **out);
/* Attach targets / sample weights that stay resident in HBM across builds.  n_target must equal
 * the table's rows (the reference's require at pkg:624-627 / 715-718) else ET_EINVAL.  Weights must
 * be non-negative (pkg:631-633); pass NULL to clear (sampleWeights = None). */
int et_data_set_target_classification(et_ctx *ctx, et_data *data, const int32_t *target_host,
                                      int64_t n_target, int32_t num_classes);
int et_data_set_target_regression(et_ctx *ctx, et_data *data, const double *target_host,
                                  int64_t n_target);
int et_data_set_weights(et_ctx *ctx, et_data *data, const double *weights_host, int64_t n_weights);
int et_data_dims(const et_data *data, int64_t *n, int32_t *d);
void et_data_free(et_data *data);

/* ---- replay test hook ------------------------------------------------------------------------
 * Candidate features and threshold uniforms replayed from the reference's own RNG stream
 * (Cmwc5, consumed in DFS order -- a level-wise builder cannot regenerate it).  Trees are given in
 * pre-order; node i of tree t is entry node_offset[t]+i of the node arrays; its ordered draws are
 * cand_feature/cand_u[cand_begin .. cand_begin+cand_count).  cand_u is the raw nextDouble() in
 * [0,1) (NaN when the reference drew none: constant feature, or bestSplit).  cand_flag: 0 constant,
 * 1 scored, 2 scored-NaN (pkg:283-285) -- only used to cross-check the GPU's own decision. */
typedef struct et_replay {
  int32_t n_trees;
  const int64_t *node_offset; /* n_trees + 1 */
  const int64_t *cand_begin;  /* per node, absolute index into cand_* */
  const int32_t *cand_count;  /* per node */
  const int32_t *left;        /* per node, tree-local pre-order id, -1 for leaves */
  const int32_t *right;
  int64_t n_cand;
  const int32_t *cand_feature;
  const double *cand_u;
  const uint8_t *cand_flag;
} et_replay;

/* ---- per-call statistics (algorithmic-byte counters of SURVEY 8d) -------------------------- */
typedef struct et_stats {
  int64_t v_mm;        /* (sample, feature) visits of min/max incl. constant hits          */
  int64_t v_sc;        /* (sample, feature) visits of threshold scoring                    */
  int64_t s_rows;      /* sum of n over nodes that reach split search                      */
  int64_t p_rows;      /* sum of n over nodes actually split                               */
  int64_t draws;       /* candidate features examined                                      */
  int64_t const_hits;  /* of which constant in the node                                    */
  int64_t scored;      /* of which scored                                                  */
  int64_t nodes;       /* nodes in the built trees                                         */
  int64_t levels;      /* level-wise passes                                                */
  int64_t rounds;      /* candidate rounds (free-running redraws after constant hits)      */
  int64_t launches;    /* CUDA kernels launched by this call                               */
  int64_t replay_mismatches; /* replay only: decisions that contradict the trace           */
  double gpu_ms;       /* device time of the call (CUDA events on the context's stream)    */
  double gpu_ms_split; /* ... of which split-search kernels (min/max + score)              */
  double gpu_ms_partition;
  /* Regression nodes of more than 2048 samples are scored from fixed-shape parallel sums instead of the
   * reference's sequential order (deterministic, ~1e-15 of the node variance from the exactly rounded score):
   * how many such nodes were searched, and in how many of them the best two candidates scored within 1e-9
   * (relative) of each other -- the only places where the chosen split could differ from the reference's. */
  int64_t parallel_sum_nodes;
  int64_t ambiguous_splits;
} et_stats;

/* ---- build ---------------------------------------------------------------------------------
 * Replaces buildForestClassification (pkg:611-681) / buildForestRegression (pkg:704-764).
 * target/weights: host arrays uploaded for this call (and left attached to `data`), or NULL to use the ones
 * attached to `data`; the two are independent (et_data_set_weights(.., NULL, 0) detaches weights).
 * parallelism selects the reference's SEEDING SCHEME (pkg:634 vs 654-655), not the GPU's
 * parallelism.  tree_ids (m entries, NULL = 0..m-1) are the global indices of the trees this GPU
 * builds: a tree's random stream depends only on (seed, tree id), so a forest sharded over G GPUs
 * equals the forest built on one.  replay == NULL: free-running counter-based RNG. */
int et_build_classification(et_ctx *ctx, et_data *data, const int32_t *target, int64_t n_target,
                            const double *weights, int32_t num_classes, int32_t n_min, int32_t k,
                            int32_t m, int32_t parallelism, int32_t best_split, int32_t max_depth,
                            int64_t seed, const int32_t *tree_ids, const et_replay *replay,
                            et_forest **out, et_stats *stats);
int et_build_regression(et_ctx *ctx, et_data *data, const double *target, int64_t n_target,
                        int32_t n_min, int32_t k, int32_t m, int32_t parallelism, int32_t best_split,
                        int32_t max_depth, int64_t seed, const int32_t *tree_ids,
                        const et_replay *replay, et_forest **out, et_stats *stats);

/* ---- forest (replaces Seq[ClassificationTree] / Seq[RegressionTree], extratrees.scala:3-63) -- */
/* leaf_width = numClasses (ClassificationLeaf.targetDistribution) or 1 (RegressionLeaf.targetMean) */
int et_forest_dims(const et_forest *f, int32_t *m, int32_t *leaf_width, int32_t *is_regression,
                   int64_t *total_nodes);
int et_forest_tree_size(const et_forest *f, int32_t t, int32_t *n_nodes);
/* Pre-order export of tree t into caller-allocated arrays: feature (-1 = leaf), cutpoint,
 * splitMissingIsLess, left/right child (tree-local pre-order ids, -1 for leaves), leaf values
 * (n_nodes x leaf_width, rows of non-leaves are zero). */
int et_forest_export(const et_forest *f, int32_t t, int32_t *feature, double *cut,
                     uint8_t *missing_is_less, int32_t *left, int32_t *right, double *leaf_values);
/* Whole forest in one call (arrays concatenated over trees in tree order; sizes from
 * et_forest_tree_size).  This is the serialized form gathered across GPUs. */
int et_forest_export_all(const et_forest *f, int32_t *tree_sizes, int32_t *feature, double *cut,
                         uint8_t *missing_is_less, int32_t *left, int32_t *right,
                         double *leaf_values);
/* JVM/host-held trees -> a forest handle usable by et_predict_*. */
int et_forest_import(et_ctx *ctx, int32_t m, int32_t leaf_width, int32_t is_regression,
                     const int32_t *tree_sizes, const int32_t *feature, const double *cut,
                     const uint8_t *missing_is_less, const int32_t *left, const int32_t *right,
                     const double *leaf_values, et_forest **out);
/* Packed serialized forest = the device layout itself, copied device->host without any per-node
 * work (the caller's buffers may be pinned): `nodes` holds total_nodes records of 16 bytes
 *     { double cutpoint; int32 feature | (splitMissingIsLess << 30), -1 for a leaf;
 *       int32 right child (tree-local pre-order id; the left child is node + 1) | forest-wide leaf index }
 * for all trees concatenated in tree order, each tree in pre-order; `leaves` holds total_leaves x
 * leaf_width doubles (leaf values in pre-order); `tree_off` holds m + 1 node offsets.  This is the
 * persisted / gathered form of Seq[ClassificationTree] / Seq[RegressionTree] (extratrees.scala:3-63). */
int et_forest_packed_dims(const et_forest *f, int64_t *total_nodes, int64_t *total_leaves);
int et_forest_export_packed(et_forest *f, void *nodes_out, double *leaves_out, int64_t *tree_off_out);
int et_forest_import_packed(et_ctx *ctx, int32_t m, int32_t leaf_width, int32_t is_regression,
                            int64_t total_nodes, int64_t total_leaves, const void *nodes,
                            const double *leaves, const int64_t *tree_off, et_forest **out);
void et_forest_free(et_forest *f);

/* ---- predict -------------------------------------------------------------------------------
 * Replaces predictClassification (pkg:542-551; out is n x numClasses row-major) and
 * predictRegression (pkg:577-586; out has n entries).  Output = mean over trees, accumulated in
 * tree order like the reference.  `sum_only` != 0 skips the division by m and returns the plain
 * sum over this forest's trees (the partial a tree-sharded predict all-reduces). */
int et_predict_classification(et_ctx *ctx, et_forest *f, const double *x_rowmajor_host, int64_t n,
                              int32_t d, double *out_host, int32_t sum_only);
int et_predict_regression(et_ctx *ctx, et_forest *f, const double *x_rowmajor_host, int64_t n,
                          int32_t d, double *out_host, int32_t sum_only);
/* Device-resident variants (x and out are device pointers; no host copies). */
int et_predict_classification_device(et_ctx *ctx, et_forest *f, const double *x_rowmajor_dev,
                                     int64_t n, int32_t d, double *out_dev, int32_t sum_only);
int et_predict_regression_device(et_ctx *ctx, et_forest *f, const double *x_rowmajor_dev, int64_t n,
                                 int32_t d, double *out_dev, int32_t sum_only);

/* ---- one process per GPU (torchrun / MPI style launchers) ------------------------------------------------
 * The same collectives for callers that run one PROCESS per GPU: rank 0 creates a 128-byte NCCL unique id, the
 * launcher's own channel (torch.distributed, MPI, a file) hands it to the other ranks, and every rank attaches a
 * communicator to its single-GPU context.  The caller shards the trees itself (tree_ids of et_build_*). */
#define ET_COMM_ID_BYTES 128
int et_comm_unique_id(uint8_t *id_out);
int et_comm_init_rank(et_ctx *ctx, int32_t world, int32_t rank, const uint8_t *id);
/* Replicates a resident table from rank `root` to every rank over NVLink (ncclBroadcast): `data` is the table
 * on the root rank and NULL elsewhere; every rank gets a handle in *out (the root gets `data` itself). */
int et_data_broadcast(et_ctx *ctx, et_data *data, int32_t root, et_data **out);
/* All-gather of the serialized trees: every rank passes the forest of its own trees and gets the whole forest,
 * trees ordered by their global tree id (the tree_ids the shards were built with). */
int et_forest_allgather(et_ctx *ctx, et_forest *shard, et_forest **full);
/* Tree-sharded predict on device buffers: partial sums over this rank's trees, ncclAllReduce(sum) in row chunks
 * overlapped with the traversal, one division by m_total (the forest's tree count over all ranks). */
int et_predict_classification_allreduce(et_ctx *ctx, et_forest *shard, const double *x_rowmajor_dev, int64_t n,
                                        int32_t d, double *out_dev, int32_t m_total);
int et_predict_regression_allreduce(et_ctx *ctx, et_forest *shard, const double *x_rowmajor_dev, int64_t n,
                                    int32_t d, double *out_dev, int32_t m_total);
/* Device time of the collectives queued by the last gather / all-reduce call of this context (ms). */
double et_comm_last_ms(const et_ctx *ctx);

/* ---- test hooks (host-side arithmetic shared with the kernels) ------------------------------ */
/* Value of adding `c` to 0.0 `h` times in round-to-nearest FP64 -- the closed form the kernels use
 * for the reference's `ar(j) += 1d / s` loop (pkg:905-911). */
double et_debug_repeat_add(double c, int64_t h);

#ifdef __cplusplus
}
#endif
#endif /* ETGPU_H */
