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
  int32_t m = f->m;
  int lw = f->leaf_width;
  const int threads = 128;
  size_t smem = f->is_regression ? 0 : (size_t)lw * threads * sizeof(double);
  if (smem > 200 * 1024) ET_FAIL(ET_EUNSUPPORTED, "predict: numClasses=%d needs more shared memory than one SM has", lw);
  unsigned grid = (unsigned)ceil_div(n, threads);
  const PNode *nodes = f->d_nodes;
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
