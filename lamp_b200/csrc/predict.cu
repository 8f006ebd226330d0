// predict.cu -- batched forest traversal (replaces predictClassification pkg:513-551 and
// predictRegression pkg:553-586).
//
// Device forest layout: all trees concatenated in pre-order; one 16-byte node
//   { double cut; int32 feat; int32 right_or_leaf }
// fetched with a single 128-bit load.  Pre-order makes the left child implicit (node+1), so only the
// right child is stored; for a leaf, feat == -1 and right_or_leaf indexes the compact leaf table
// (leaf_width doubles per leaf).  splitMissingIsLess rides in bit 30 of feat.
#include <algorithm>

#include "internal.h"

struct __align__(16) PNode {
  double cut;
  int32_t feat;
  int32_t right_or_leaf;
};

#define ET_MIL_BIT 0x40000000

void et_forest_upload(et_ctx *ctx, et_forest *f) {
  if (f->dev_ready) return;
  size_t m = f->trees.size();
  std::vector<int64_t> off(m + 1, 0);
  for (size_t t = 0; t < m; t++) off[t + 1] = off[t] + (int64_t)f->trees[t].feature.size();
  int64_t total = off[m];
  int lw = f->leaf_width;
  std::vector<PNode> nodes((size_t)total);
  std::vector<double> leaves;
  leaves.reserve((size_t)(total / 2 + 1) * lw);
  int64_t n_leaves = 0;
  for (size_t t = 0; t < m; t++) {
    const HostTree &tr = f->trees[t];
    size_t n = tr.feature.size();
    for (size_t i = 0; i < n; i++) {
      PNode p;
      p.cut = tr.cut[i];
      if (tr.feature[i] >= 0) {
        if (tr.left[i] != (int32_t)i + 1) ET_FAIL(ET_EINVAL, "predict: tree %zu is not in pre-order", t);
        p.feat = tr.feature[i] | (tr.mil[i] ? ET_MIL_BIT : 0);
        p.right_or_leaf = tr.right[i];
      } else {
        p.feat = -1;
        if (n_leaves > 0x7fffffff) ET_FAIL(ET_EUNSUPPORTED, "predict: more than 2^31 leaves");
        p.right_or_leaf = (int32_t)n_leaves++;
        for (int c = 0; c < lw; c++) leaves.push_back(tr.leaf[i * lw + c]);
      }
      nodes[(size_t)off[t] + i] = p;
    }
  }
  f->total_nodes = total;
  CUDA_CHECK(cudaMalloc((void **)&f->d_tree_off, (m + 1) * sizeof(int64_t)));
  CUDA_CHECK(cudaMalloc((void **)&f->d_cut, std::max<size_t>(1, (size_t)total) * sizeof(PNode)));
  CUDA_CHECK(cudaMalloc((void **)&f->d_leaf, std::max<size_t>(1, leaves.size()) * sizeof(double)));
  CUDA_CHECK(cudaMemcpyAsync(f->d_tree_off, off.data(), (m + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
  if (total)
    CUDA_CHECK(cudaMemcpyAsync(f->d_cut, nodes.data(), (size_t)total * sizeof(PNode), cudaMemcpyHostToDevice, ctx->stream));
  if (!leaves.empty())
    CUDA_CHECK(cudaMemcpyAsync(f->d_leaf, leaves.data(), leaves.size() * sizeof(double), cudaMemcpyHostToDevice,
                               ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  f->dev_ready = true;
}

// One thread per row; the row's features are read straight from the row-major matrix (the rows of
// a block are contiguous, so the lines it touches are shared in L1/L2).  Per-class sums are kept in
// shared memory in [class][thread] order and accumulated in tree order, like the reference's
// sequential mean over trees (pkg:549,584).
template <bool REG>
__global__ void __launch_bounds__(128) k_predict(const PNode *__restrict__ nodes, const int64_t *__restrict__ tree_off,
                                                 const double *__restrict__ leaves, int32_t m, int32_t lw,
                                                 const double *__restrict__ x, int64_t n, int32_t d,
                                                 double *__restrict__ out, int sum_only) {
  extern __shared__ double acc[];  // [lw][blockDim.x]
  int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  const double *xr = x + row * (int64_t)d;
  double racc = 0.0;
  if (!REG)
    for (int c = 0; c < lw; c++) acc[c * blockDim.x + threadIdx.x] = 0.0;
  for (int32_t t = 0; t < m; t++) {
    const PNode *tn = nodes + tree_off[t];
    int32_t id = 0;
    PNode p = tn[0];
    while (p.feat >= 0) {
      double v = xr[p.feat & (ET_MIL_BIT - 1)];
      bool left = (v < p.cut) || ((p.feat & ET_MIL_BIT) && (v != v));
      id = left ? id + 1 : p.right_or_leaf;
      p = tn[id];
    }
    const double *lv = leaves + (int64_t)p.right_or_leaf * lw;
    if (REG) {
      racc = ET_ADD(racc, lv[0]);
    } else {
      for (int c = 0; c < lw; c++) {
        double *a = &acc[c * blockDim.x + threadIdx.x];
        *a = ET_ADD(*a, lv[c]);
      }
    }
  }
  double dm = (double)m;
  if (REG) {
    out[row] = sum_only ? racc : ET_DIV(racc, dm);
  } else {
    for (int c = 0; c < lw; c++) {
      double a = acc[c * blockDim.x + threadIdx.x];
      out[row * lw + c] = sum_only ? a : ET_DIV(a, dm);
    }
  }
}

void et_predict_device_impl(et_ctx *ctx, et_forest *f, const double *x, int64_t n, int32_t d, double *out,
                            int sum_only) {
  et_forest_upload(ctx, f);
  int32_t m = (int32_t)f->trees.size();
  int lw = f->leaf_width;
  const int threads = 128;
  size_t smem = f->is_regression ? 0 : (size_t)lw * threads * sizeof(double);
  if (smem > 200 * 1024) ET_FAIL(ET_EUNSUPPORTED, "predict: numClasses=%d needs more shared memory than one SM has", lw);
  unsigned grid = (unsigned)ceil_div(n, threads);
  const PNode *nodes = (const PNode *)f->d_cut;
  if (f->is_regression) {
    k_predict<true><<<grid, threads, 0, ctx->stream>>>(nodes, f->d_tree_off, f->d_leaf, m, lw, x, n, d, out, sum_only);
  } else {
    if (smem > 48 * 1024)
      CUDA_CHECK(cudaFuncSetAttribute(k_predict<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_predict<false><<<grid, threads, smem, ctx->stream>>>(nodes, f->d_tree_off, f->d_leaf, m, lw, x, n, d, out,
                                                           sum_only);
  }
  ctx->launches++;
  CUDA_CHECK(cudaGetLastError());
}
