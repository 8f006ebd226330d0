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

// A CTA owns R rows; its 256 threads are R x TS (row, tree-slice) pairs.  Traversal is parallel over
// (row, tree): thread (r, s) walks trees s, s + TS, ... of the current block of trees for row r and
// parks the leaf it reaches in shared memory.  The per-class sums are then accumulated by one thread
// per (row, class) IN TREE ORDER, like the reference's sequential mean over trees (pkg:549,584), so the
// result is bit-identical to the one-thread-per-row walk while 256/R times more walks are in flight.
// Threads of one slice walk the same tree for neighbouring rows, so a tree's top levels stay in L1.
// STAGE: the CTA's rows are staged once in shared memory (coalesced), so every feature lookup of the walks
// is a shared-memory read instead of a scattered 8-byte read of a row-major matrix larger than L2.
template <bool STAGE>
__global__ void __launch_bounds__(256) k_predict(const PNode *__restrict__ nodes, const int64_t *__restrict__ tree_off,
                                                 const double *__restrict__ leaves, int32_t m, int32_t lw,
                                                 const double *__restrict__ x, int64_t n, int32_t d,
                                                 double *__restrict__ out, int sum_only, int R, int TB) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *s_acc = reinterpret_cast<double *>(smem_raw);              // [R][lw]
  double *s_x = s_acc + R * lw;                                      // [R][d] (STAGE)
  int32_t *s_leaf = reinterpret_cast<int32_t *>(s_x + (STAGE ? (size_t)R * d : 0));  // [R][TB]
  const int TS = 256 / R;
  const int r = threadIdx.x / TS, sl = threadIdx.x % TS;
  const int64_t row0 = (int64_t)blockIdx.x * R;
  const int64_t row = row0 + r;
  const bool has_row = row < n;
  const double *xr = STAGE ? (s_x + (size_t)r * d) : (x + (has_row ? row : 0) * (int64_t)d);
  for (int q = threadIdx.x; q < R * lw; q += 256) s_acc[q] = 0.0;
  if (STAGE) {
    const int64_t rows_here = min((int64_t)R, n - row0);
    const double *src = x + row0 * (int64_t)d;  // the CTA's rows are contiguous in the row-major matrix
    for (int64_t q = threadIdx.x; q < rows_here * d; q += 256) s_x[q] = src[q];
    __syncthreads();
  }
  for (int32_t t0 = 0; t0 < m; t0 += TB) {
    const int32_t tb = min(TB, m - t0);
    if (has_row) {
      // four walks per thread in flight: the pointer chases of different trees are independent, so their node
      // fetches overlap (a walk is one dependent L2 access per level)
      for (int32_t t = sl; t < tb; t += 4 * TS) {
        const PNode *tn[4];
        PNode pn[4];
        int32_t id[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int32_t tu = t + u * TS;
          tn[u] = nodes + tree_off[t0 + (tu < tb ? tu : t)];
          id[u] = 0;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          pn[u] = tn[u][0];
          if (t + u * TS >= tb) pn[u].feat = -1;  // no such tree: the slot is idle
        }
        bool more = true;
        while (more) {
          more = false;
#pragma unroll
          for (int u = 0; u < 4; u++) {
            if (pn[u].feat >= 0) {
              const double v = xr[pn[u].feat & (ET_MIL_BIT - 1)];
              const bool left = (v < pn[u].cut) || ((pn[u].feat & ET_MIL_BIT) && (v != v));
              id[u] = left ? id[u] + 1 : pn[u].right_or_leaf;
              pn[u] = tn[u][id[u]];
              more = true;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (t + u * TS < tb) s_leaf[r * TB + t + u * TS] = pn[u].right_or_leaf;
      }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < R * lw; q += 256) {
      const int rr = q / lw, c = q - rr * lw;
      if (row0 + rr < n) {
        double a = s_acc[q];
        const int32_t *lf = s_leaf + rr * TB;
#pragma unroll 8
        for (int32_t t = 0; t < tb; t++) a = ET_ADD(a, leaves[(int64_t)lf[t] * lw + c]);  // (loads run ahead of the adds)
        s_acc[q] = a;
      }
    }
    __syncthreads();
  }
  const double dm = (double)m;
  for (int q = threadIdx.x; q < R * lw; q += 256) {
    const int rr = q / lw, c = q - rr * lw;
    if (row0 + rr < n) {
      const double a = s_acc[q];
      out[(row0 + rr) * lw + c] = sum_only ? a : ET_DIV(a, dm);
    }
  }
}

void et_predict_device_impl(et_ctx *ctx, et_forest *f, const double *x, int64_t n, int32_t d, double *out,
                            int sum_only) {
  const int32_t m = f->m;
  const int lw = f->leaf_width;
  const int TB = std::max(1, std::min(m, 1024));
  // rows per CTA: as many as keep the CTA's shared memory near 48 KB (4+ CTAs per SM); the rows themselves are
  // staged when one row fits in 24 KB
  const bool stage = (size_t)d * 8 <= 24 * 1024;
  const size_t per_row = (size_t)TB * 4 + (size_t)lw * 8 + (stage ? (size_t)d * 8 : 0);
  int R = 16;
  while (R > 1 && (size_t)R * per_row > 48 * 1024) R /= 2;
  const size_t smem = (size_t)R * per_row;
  if (smem > 200 * 1024) ET_FAIL(ET_EUNSUPPORTED, "predict: numClasses=%d needs more shared memory than one SM has", lw);
  const unsigned grid = (unsigned)ceil_div(n, R);
  if (stage) {
    CUDA_CHECK(cudaFuncSetAttribute(k_predict<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_predict<true><<<grid, 256, smem, ctx->stream>>>(f->d_nodes, f->d_tree_off, f->d_leaf, m, lw, x, n, d, out, sum_only,
                                                      R, TB);
  } else {
    CUDA_CHECK(cudaFuncSetAttribute(k_predict<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_predict<false><<<grid, 256, smem, ctx->stream>>>(f->d_nodes, f->d_tree_off, f->d_leaf, m, lw, x, n, d, out,
                                                       sum_only, R, TB);
  }
  ctx->launches++;
  CUDA_CHECK(cudaGetLastError());
}
