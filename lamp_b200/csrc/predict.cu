// predict.cu -- batched forest traversal (replaces predictClassification pkg:513-551 and
// predictRegression pkg:553-586).
//
// Device forest layout: all trees concatenated in pre-order; one 16-byte node
//   { double cut; int32 feat; int32 right_or_leaf }
// fetched with a single 128-bit load.  Pre-order makes the left child implicit (node+1), so only the
// right child is stored; for a leaf, feat == -1 and right_or_leaf indexes the compact leaf table
// (leaf_width doubles per leaf).  splitMissingIsLess rides in bit 30 of feat.
#include <algorithm>

#include "bulk.cuh"
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
                                                 double *__restrict__ out, int sum_only, int R, int TB, int32_t t_begin,
                                                 int32_t t_end, int bulk) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_bar;
  double *s_acc = reinterpret_cast<double *>(smem_raw);              // [R][lw]
  double *s_x = s_acc + R * lw;                                      // [R][d] (STAGE)
  int32_t *s_leaf = reinterpret_cast<int32_t *>(s_x + (STAGE ? (size_t)R * d : 0));  // [R][TB]
  const int TS = 256 / R;
  const int r = threadIdx.x / TS, sl = threadIdx.x % TS;
  const int64_t row0 = (int64_t)blockIdx.x * R;
  const int64_t row = row0 + r;
  const bool has_row = row < n;
  const double *xr = STAGE ? (s_x + (size_t)r * d) : (x + (has_row ? row : 0) * (int64_t)d);
  // trees [t_begin, t_end) of the forest: a launch continues the running sums of the launches before it (same
  // additions in the same order as one pass over all trees)
  for (int q = threadIdx.x; q < R * lw; q += 256) {
    const int rr = q / lw;
    s_acc[q] = (t_begin > 0 && row0 + rr < n) ? out[(row0 + rr) * lw + (q - rr * lw)] : 0.0;
  }
  if (STAGE) {
    const int64_t rows_here = min((int64_t)R, n - row0);
    const double *src = x + row0 * (int64_t)d;  // the CTA's rows are contiguous in the row-major matrix
    const uint32_t bytes = (uint32_t)(rows_here * d * 8);
    if (bulk && (((uintptr_t)src | (uintptr_t)s_x | bytes) & 15u) == 0) {
      // one TMA bulk copy queued by one thread (the block of rows is contiguous and 16-byte aligned)
      if (threadIdx.x == 0) {
        etb::mbar_init(&s_bar, 1);
        etb::mbar_arrive_expect_tx(&s_bar, bytes);
        etb::bulk_copy_g2s(s_x, src, bytes, &s_bar);
      }
      __syncthreads();  // the barrier is initialised before anyone waits on it
      etb::mbar_wait(&s_bar, 0);
    } else {
      for (int64_t q = threadIdx.x; q < rows_here * d; q += 256) s_x[q] = src[q];
    }
    __syncthreads();
  }
  for (int32_t t0 = t_begin; t0 < t_end; t0 += TB) {
    const int32_t tb = min(TB, t_end - t0);
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
      out[(row0 + rr) * lw + c] = (sum_only || t_end < m) ? a : ET_DIV(a, dm);
    }
  }
}

void et_predict_device_impl(et_ctx *ctx, et_forest *f, const double *x, int64_t n, int32_t d, double *out,
                            int sum_only) {
  const int32_t m = f->m;
  const int lw = f->leaf_width;
  // Optional (ETGPU_PREDICT_BLOCK_MB): traverse the trees in blocks whose nodes fit a byte budget, one launch per
  // block over all rows, the per-row sums continuing from launch to launch in tree order (bit-identical to one
  // pass).  Measured on B200 (profiles/r2_predict.txt): no gain -- the 80 MB forest of the 500-tree MNIST-shaped
  // workload already hits in L2 (70 %) and the kernel is bound by L1TEX request throughput (75 % of peak: one
  // 16-byte node = one sector per lane and visit), not by DRAM latency -- so the default is a single pass.
  int64_t budget = (int64_t)1 << 60;
  if (const char *env = getenv("ETGPU_PREDICT_BLOCK_MB")) budget = std::max<int64_t>(1, atoll(env)) << 20;
  const int64_t bytes_per_tree = std::max<int64_t>(1, (f->total_nodes * (int64_t)sizeof(PNode) + f->total_leaves * lw * 8) / std::max(m, 1));
  int32_t block = (int32_t)std::min<int64_t>(m, std::max<int64_t>(64, budget / bytes_per_tree));
  const int TB = std::max(1, std::min(block, 1024));
  // rows per CTA: as many as keep the CTA's shared memory near 48 KB (4+ CTAs per SM); the rows themselves are
  // staged when one row fits in 24 KB
  const bool stage = (size_t)d * 8 <= 24 * 1024;
  const size_t per_row = (size_t)TB * 4 + (size_t)lw * 8 + (stage ? (size_t)d * 8 : 0);
  int R = 16;
  while (R > 1 && (size_t)R * per_row > 48 * 1024) R /= 2;
  const size_t smem = (size_t)R * per_row;
  if (smem > 200 * 1024) ET_FAIL(ET_EUNSUPPORTED, "predict: numClasses=%d needs more shared memory than one SM has", lw);
  const unsigned grid = (unsigned)ceil_div(n, R);
  static const int bulk = (getenv("ETGPU_NO_BULK") != nullptr && atoi(getenv("ETGPU_NO_BULK")) != 0) ? 0 : 1;
  if (stage)
    CUDA_CHECK(cudaFuncSetAttribute(k_predict<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else
    CUDA_CHECK(cudaFuncSetAttribute(k_predict<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int32_t t0 = 0; t0 < std::max(m, 1); t0 += block) {
    const int32_t t1 = std::min(m, t0 + block);
    if (stage)
      k_predict<true><<<grid, 256, smem, ctx->stream>>>(f->d_nodes, f->d_tree_off, f->d_leaf, m, lw, x, n, d, out, sum_only,
                                                        R, TB, t0, t1, bulk);
    else
      k_predict<false><<<grid, 256, smem, ctx->stream>>>(f->d_nodes, f->d_tree_off, f->d_leaf, m, lw, x, n, d, out,
                                                         sum_only, R, TB, t0, t1, bulk);
    ctx->launches++;
  }
  CUDA_CHECK(cudaGetLastError());
}
