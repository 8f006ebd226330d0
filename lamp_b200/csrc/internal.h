// internal.h -- host-side objects behind the opaque handles of include/etgpu.h
#pragma once
#include "common.cuh"

struct et_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;  // the one work is queued on (own or caller's)
  int sm_count = 148;
  std::mutex mu;  // a context serialises its calls
  int64_t launches = 0;
};

struct et_data {
  et_ctx *ctx = nullptr;
  int64_t n = 0;
  int32_t d = 0;
  int64_t ld = 0;       // column stride in elements (n rounded up to 16)
  double *x = nullptr;  // column-major [d][ld], resident in HBM
  // attached targets / weights (resident)
  int32_t *y_cls = nullptr;
  int32_t num_classes = 0;
  double *y_reg = nullptr;
  double *w = nullptr;
  std::vector<int64_t> root_hist;  // class counts of the whole table
};

// One tree, pre-order (the wire format of et_forest_export).
struct HostTree {
  std::vector<int32_t> feature, left, right;
  std::vector<double> cut;
  std::vector<uint8_t> mil;
  std::vector<double> leaf;  // n_nodes x leaf_width
};

struct et_forest {
  et_ctx *ctx = nullptr;
  int32_t leaf_width = 1;
  int32_t is_regression = 0;
  std::vector<HostTree> trees;
  // device copy for predict (built lazily): trees concatenated
  bool dev_ready = false;
  int64_t total_nodes = 0;
  int32_t max_depth = 0;
  int64_t *d_tree_off = nullptr;  // m+1
  int32_t *d_feature = nullptr;
  double *d_cut = nullptr;
  int32_t *d_left = nullptr;   // left child (tree-local); right = stored separately
  int32_t *d_right = nullptr;
  uint8_t *d_mil = nullptr;
  double *d_leaf = nullptr;    // total_nodes x leaf_width
  ~et_forest();
};

struct BuildArgs {
  int task;  // 0 cls unweighted, 1 cls weighted, 2 regression
  int32_t num_classes, n_min, k, m, parallelism, best_split, max_depth;
  int64_t seed;
  const int32_t *tree_ids;
  const et_replay *replay;
};

// build.cu
void et_build_forest(et_ctx *ctx, et_data *data, const BuildArgs &a, et_forest *out, et_stats *stats);
// predict.cu
void et_forest_upload(et_ctx *ctx, et_forest *f);
void et_predict_device_impl(et_ctx *ctx, et_forest *f, const double *x_dev, int64_t n, int32_t d,
                            double *out_dev, int sum_only);
// api.cu
void et_launch_transpose(et_ctx *ctx, const double *src_rowmajor, int64_t rows, int32_t d,
                         double *dst_colmajor, int64_t ld, int64_t row0);
