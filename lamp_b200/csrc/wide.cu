// wide.cu -- nodes too large for one CTA: the node's row segment is cut into chunks, one CTA per chunk.
//
// The reference recursion (buildTreeClassification pkg:943-1082, buildTreeRegression pkg:766-895) has no node-size
// limit; a level-wise builder that gives a node to ONE team leaves the GPU idle at the top of the tree (a 10M-row
// table with a handful of trees keeps 4-8 of 148 SMs busy).  Here every phase of a node is a grid over
// (node, chunk of rows) work items of all wide nodes of the level, with per-node state in global memory:
//
//   k_wide_plan      one CTA            rows per node, chunks per node, exclusive scan -> first chunk of each node
//   k_wide_regsum    CTA per chunk      (regression) partial sums of the targets, uniform-target test (pkg:813-814)
//   k_wide_regmom    CTA per chunk      (regression) node mean from the partial sums; moments about the mean
//   k_wide_first     warp per node      stop rules (pkg:993-994 / 813-814), node impurity / variance, first batch of
//                                       candidates (replayed trace | counter RNG)
//   rounds, until every node has scored k candidates or run out of features (pkg:232):
//     k_wide_pass1   CTA per chunk      min / max / hasMissing per candidate (pkg:34-54), merged with atomics (exact:
//                                       min / max are order-free)
//     k_wide_decide  warp per node      constant test (pkg:236), cutpoint min + (max - min) * u (pkg:240), threshold
//     k_wide_pass2   CTA per chunk      classification: side class histograms (integers, atomics: exact);
//                                       regression: per-chunk moments of the left side (summed in chunk order)
//     k_wide_score   warp per node      exact score per candidate, candidates consumed in draw order, first best
//                                       with strict `>` (pkg:272-290); draws the next batch if the node needs one
//   k_wide_finish    warp per node      leaf | children (frontier entries, class histograms, inherited constants)
//   k_wide_count     CTA per chunk      side bits of the winning split per row, rows going left per chunk
//   k_wide_scatter   CTA per chunk      stable partition (pkg:1024-1039): a chunk's offsets = sums over the chunks
//                                       before it in the node
//
// Unweighted classification is bit-exact by construction (integer histograms, one thread evaluates the
// reference's Gini expression).  Regression nodes of this size are scored from fixed-shape parallel moment sums
// like the one-CTA path before (et_stats.parallel_sum_nodes / ambiguous_splits); weighted classification keeps the
// one-CTA path (sequential weight sums).
#include "bulk.cuh"
#include "node.cuh"

namespace etb {

constexpr int WT = 256;           // threads of a chunk CTA
constexpr int WCHUNK_MAX = 8192;  // rows per chunk (one side-bit word per thread in the scatter)

struct WNode {  // search state of one wide node
  int32_t fi, n, b, tree;  // frontier index, rows, segment begin (tree-local), tree of the batch
  int32_t flags;           // 1 leaf by stop rule, 2 search finished, 4 split (children made)
  int32_t nb;              // candidates in the current batch (0: none drawn)
  int32_t nsweep;          // 2: a candidate's column holds NaNs in this node (second histogram sweep)
  int32_t visited, nconst, dc, tpos, tcnt;
  int32_t best_feature, best_mil, best_nleft, best_thr, best_K;
  int32_t not_uniform;     // regression: some target differs from the first
  int32_t slot;            // first of the two child frontier slots
  int32_t pad_;
  int64_t tb;
  double total, nsum, leaf_mean, reg_mu, reg_S, reg_Q;
  double best_score, second_score, best_cut;
  unsigned long long st_draws, st_const, st_scored, st_mismatch;
};

struct WCand {  // the batch of candidates of one wide node and their accumulators
  int32_t feat[32], flags[32], thr[32], nleft[32];
  double u[32], cut[32], score[32];
  uint32_t mnT[32], mxB[32], mnB[32];           // byte codes: min of (byte - K), max byte, min byte
  unsigned long long mn[32], mx[32];            // FP64: order-preserving keys of min / max
  uint32_t nanmask;                             // FP64: bit c = candidate c saw a NaN
  uint32_t pad_;
  uint8_t thrb[32], enb[32], Kb[32], nanb[32];  // byte codes: pass-2 parameters (thr - 1, enable, K, NaN enable)
};

struct WState {
  WNode *node;
  WCand *cand;
  int64_t *chunk0;              // [count + 1] first chunk of every node
  int32_t *hist;                // [count][32][2C] side histograms of the batch (classification)
  int32_t *besthl;              // [count][C] left histogram of the best split so far
  uint32_t *cmask, *taken;      // [count][W] known-constant / drawn features (free-running)
  double *part;                 // regression: [chunks][2 sweeps][32][3] = (count, S, Q) of the left side per chunk
  double *ysum;                 // regression: [chunks][2] partial target sums, moments about the mean
  int32_t *cnt_left;            // [chunks] rows going left
  uint32_t *bits;               // [chunks][chunk / 32] side bits of the winning split
  unsigned long long *pending;  // [0] nodes that drew a batch this round
  int32_t chunk;                // rows per chunk
  int32_t count;
  int32_t bulk;                 // stage a chunk's sample-index segment with one TMA bulk copy (else plain loads)
  // FP64 tables: pass 1 parks the values it gathers, [chunk][candidate][row of the chunk], so that pass 2 streams
  // them back (8 bytes per value, coalesced) instead of gathering again (a 32-byte sector per value at depth);
  // nodes whose batch holds more than park_stride candidates gather twice.  null: off
  double *park;
  int32_t park_stride;
  // sparse-resident tables: inv[tree * n + row] = position of the row inside the tree's index segment, written each
  // level for the rows of the wide nodes (k_wide_inv); entries of other rows are stale and are validated on use
  int32_t *inv;
};

// Stages `cnt` sample indices of a node's contiguous segment in shared memory (s_buf: 16-byte aligned, room for
// cnt + 8 entries) and returns where entry 0 landed.  bulk: ONE thread queues a cp.async.bulk of the whole segment
// (start rounded down / size rounded up to 16 bytes: the index buffers carry that much slack) and every thread
// waits on the mbarrier after doing its other set-up work; else every thread copies its share.
__device__ __forceinline__ const int32_t *stage_rows(int32_t *s_buf, const int32_t *src, int32_t cnt, uint64_t *bar,
                                                     int bulk, int tid) {
  if (bulk) {
    const int lead = (int)(((uintptr_t)src & 15u) >> 2);
    const uint32_t bytes = (uint32_t)(((lead + cnt) * 4 + 15) & ~15);
    if (tid == 0) {
      mbar_arrive_expect_tx(bar, bytes);
      bulk_copy_g2s(s_buf, src - lead, bytes, bar);
    }
    return s_buf + lead;
  }
  for (int32_t j = tid; j < cnt; j += WT) s_buf[j] = src[j];
  return s_buf;
}

// order-preserving 64-bit key of a non-NaN double (-0.0 and +0.0 share a key: the reference's `<` / `>` cannot
// tell them apart and the cutpoint min + (max - min) * u does not depend on the sign of a zero bound)
__device__ __forceinline__ unsigned long long dkey(double x) {
  if (x == 0.0) x = 0.0;
  const unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dkey_inv(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

// the wide node and the chunk of it that CTA `g` owns; false when there is none
struct ChunkRef {
  int32_t q;   // wide node
  int32_t j0;  // first row of the chunk inside the node
  int32_t cnt; // rows in the chunk
};
__device__ __forceinline__ bool chunk_ref(const WState &w, int64_t g, ChunkRef &r) {
  if (g >= w.chunk0[w.count]) return false;
  int lo = 0, hi = w.count;  // last q with chunk0[q] <= g
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (w.chunk0[mid] <= g)
      lo = mid;
    else
      hi = mid;
  }
  r.q = lo;
  r.j0 = (int32_t)(g - w.chunk0[lo]) * w.chunk;
  r.cnt = min(w.chunk, w.node[lo].n - r.j0);
  return true;
}

// ---- CSC tables kept sparse: a chunk walks the stored entries of a column that fall into its row range --------
// (both the chunk's rows and a column's stored rows ascend, so the chunk's row range is a range of the column: two
// warp-cooperative searches; the rows of the chunk a column does not store are implicit zeros and enter min / max
// and the side histograms by COUNT -- 12 bytes per stored entry visited instead of one search per (row, candidate).)
// Membership of a stored entry's row in the chunk is one lookup in the inverse index map (validated against the
// index segment: the map is only current for rows of wide nodes) -- a node deep in the tree is scattered over the
// whole table, so nearly every stored entry of a candidate column is tested and nearly none is a member; a search
// among the chunk's rows for each was the instruction-bound part of the whole sparse build.
__device__ __forceinline__ int chunk_pos_inv(const int32_t *inv_tree, const int32_t *idx_tree, int32_t c0, int cnt, int32_t row) {
  const int32_t at = __ldg(inv_tree + row);
  const uint32_t pos = (uint32_t)(at - c0);
  return (pos < (uint32_t)cnt && idx_tree[at] == row) ? (int)pos : -1;
}
// first entry of [a, e) whose row is >= key (or > key with `past`): the 32 lanes of a warp probe 32 points of the
// range at once, so a column of 10^4 stored rows is searched in 3 dependent loads instead of 14.  Converged warps only.
__device__ __forceinline__ int64_t warp_bound(const int32_t *rows, int64_t a, int64_t e, int32_t key, bool past, int lane) {
  int64_t lo = a, len = e - a;
  while (len > 0) {
    const int64_t step = (len + 31) >> 5;
    const int64_t at = lo + (int64_t)lane * step;
    const bool valid = at < lo + len;
    bool before = false;
    if (valid) {
      const int32_t v = __ldg(rows + at);
      before = past ? (v <= key) : (v < key);
    }
    const int nvalid = (int)((len + step - 1) / step);                  // probes inside the range (<= 32)
    const int cnt = __popc(__ballot_sync(0xffffffffu, before));        // sorted rows: the first `cnt` probes are before
    if (cnt == 0) break;                                                // the entry at lo is the bound
    const int64_t nlo = lo + (int64_t)(cnt - 1) * step + 1;
    const int64_t nhi = (cnt < nvalid) ? lo + (int64_t)cnt * step : lo + len;
    lo = nlo;
    len = nhi - nlo;
  }
  return lo;
}
__device__ __forceinline__ void col_range(const P &p, int32_t f, int32_t rmin, int32_t rmax, int64_t &lo_out, int64_t &hi_out) {
  const int64_t a = __ldg(p.csc_colptr + f), e = __ldg(p.csc_colptr + f + 1);
  const int lane = threadIdx.x & 31;
  lo_out = warp_bound(p.csc_row, a, e, rmin, false, lane);
  hi_out = warp_bound(p.csc_row, lo_out, e, rmax, true, lane);
}

// ---- plan ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_wide_plan(P p, WState w) {
  __shared__ int64_t s_warp[32];
  __shared__ int64_t s_carry;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int q0 = 0; q0 < w.count; q0 += 1024) {
    const int q = q0 + tid;
    int64_t nch = 0;
    if (q < w.count) {
      const int i = p.q_cur[Q_WIDE][q];
      const int32_t b = p.cur.begin[i], n = p.cur.end[i] - b;
      WNode &nd = w.node[q];
      nd.fi = i;
      nd.n = n;
      nd.b = b;
      nd.tree = p.cur.tree[i];
      nd.flags = 0;
      nd.nb = 0;
      nd.not_uniform = 0;
      nch = (n + w.chunk - 1) / w.chunk;
    }
    int64_t v = nch;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) s_warp[wid] = v;
    __syncthreads();
    if (wid == 0) {
      int64_t t = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int64_t u2 = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += u2;
      }
      s_warp[lane] = t;
    }
    __syncthreads();
    const int64_t excl = s_carry + (wid > 0 ? s_warp[wid - 1] : 0) + v - nch;
    if (q < w.count) w.chunk0[q] = excl;
    __syncthreads();
    if (tid == 1023) s_carry = excl + nch;
    __syncthreads();
  }
  if (tid == 0) w.chunk0[w.count] = s_carry;
}

// sparse-resident tables: the inverse of the index segment for the rows of the wide nodes (one CTA per chunk)
__global__ void __launch_bounds__(WT) k_wide_inv(P p, WState w) {
  ChunkRef r;
  if (!chunk_ref(w, blockIdx.x, r)) return;
  const WNode &nd = w.node[r.q];
  const int64_t base = (int64_t)nd.tree * p.n;
  const int32_t c0 = nd.b + r.j0;
  for (int32_t j = threadIdx.x; j < r.cnt; j += WT) w.inv[base + p.idx_src[base + c0 + j]] = c0 + j;
}

// ---- regression: node mean and moments from chunked partial sums --------------------------------------------
// fixed-shape block sum (butterfly inside a warp, warps in index order); result valid in thread 0
__device__ __forceinline__ double block_sum_fixed(double v, double *s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = ET_ADD(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  double a = 0.0;
  if (threadIdx.x == 0)
    for (int q = 0; q < WT / 32; q++) a = ET_ADD(a, s_red[q]);
  return a;
}

__global__ void __launch_bounds__(WT) k_wide_regsum(P p, WState w) {
  __shared__ double s_red[WT / 32];
  ChunkRef r;
  if (!chunk_ref(w, blockIdx.x, r)) return;
  const WNode &nd = w.node[r.q];
  const double *yy = p.yr_src + (int64_t)nd.tree * p.n + nd.b;
  const double head = yy[0];
  double s = 0.0;
  bool uni = true;
  for (int32_t j = threadIdx.x; j < r.cnt; j += WT) {
    const double y = yy[r.j0 + j];
    s = ET_ADD(s, y);
    uni &= !(y != head);
  }
  if (!__all_sync(0xffffffffu, uni) && (threadIdx.x & 31) == 0) atomicOr(&w.node[r.q].not_uniform, 1);
  const double tot = block_sum_fixed(s, s_red);
  if (threadIdx.x == 0) w.ysum[(int64_t)blockIdx.x * 2] = tot;
}

// the node's mean: the chunk sums added in chunk order (every CTA of the node computes the same bits)
__device__ __forceinline__ double wide_node_mean(const WState &w, int q, int32_t n) {
  double s = 0.0;
  for (int64_t g = w.chunk0[q]; g < w.chunk0[q + 1]; g++) s = ET_ADD(s, w.ysum[g * 2]);
  return ET_DIV(s, (double)n);
}

__global__ void __launch_bounds__(WT) k_wide_regmom(P p, WState w) {
  __shared__ double s_red[WT / 32];
  __shared__ double s_mu;
  ChunkRef r;
  if (!chunk_ref(w, blockIdx.x, r)) return;
  const WNode &nd = w.node[r.q];
  if (threadIdx.x == 0) s_mu = wide_node_mean(w, r.q, nd.n);
  __syncthreads();
  const double mu = s_mu;
  const double *yy = p.yr_src + (int64_t)nd.tree * p.n + nd.b;
  double s = 0.0, q2 = 0.0;
  for (int32_t j = threadIdx.x; j < r.cnt; j += WT) {
    const double dl = ET_SUB(yy[r.j0 + j], mu);
    s = ET_ADD(s, dl);
    q2 = ET_ADD(q2, ET_MUL(dl, dl));
  }
  const double ts = block_sum_fixed(s, s_red);
  const double tq = block_sum_fixed(q2, s_red);
  if (threadIdx.x == 0) {
    w.part[(int64_t)blockIdx.x * 192] = ts;  // (the moment slots of the chunk are free until the first pass 2)
    w.part[(int64_t)blockIdx.x * 192 + 1] = tq;
  }
}

// ---- candidates ------------------------------------------------------------------------------------------------
// Draws the next batch of a node (one warp, lane == candidate); the same draws as the one-CTA path (k_node), so a
// free-running tree does not depend on which path built a node.
template <int TASK>
__device__ void wide_draw(const P &p, const WState &w, int q, int lane) {
  WNode &nd = w.node[q];
  WCand &cd = w.cand[q];
  const int W = p.W, C = p.C;
  uint32_t *taken = w.taken + (int64_t)q * W;
  int32_t nb;
  const int32_t visited = nd.visited, nconst = nd.nconst;
  const int32_t avail = p.d - nconst - visited;
  if (p.replay) {
    nb = min(p.NB, nd.tcnt - nd.tpos);
  } else {
    const int32_t need = min(p.k - visited, avail);
    int32_t extra = 0;
    const unsigned long long st_draws = nd.st_draws;
    if (need > 0 && st_draws > 0)
      extra = (int32_t)(((long long)need * (long long)(st_draws - (unsigned long long)visited)) / max(visited, 1));
    nb = min(p.NB, min(avail, need + extra));
    if (need <= 0) nb = 0;
  }
  __syncwarp();
  if (nb <= 0) {
    if (lane == 0) {
      nd.nb = 0;
      nd.flags |= 2;
    }
    return;
  }
  int32_t f = -1;
  double u = 0.0;
  int32_t fl = 0;
  if (p.replay) {
    if (lane < nb) {
      const int64_t t = nd.tb + nd.tpos + lane;
      f = p.tr.cand_feature[t];
      u = p.tr.cand_u[t];
      fl = (p.tr.cand_flag[t] + 1) << 4;
    }
  } else {
    const uint64_t key = p.cur.key[nd.fi];
    int32_t pick = -1 - lane;
    if (lane < nb) {
      const uint64_t r = et_draw(key, (uint32_t)(nd.dc + 2 * lane));
      pick = rank_select_clear_fast(taken, W, (int32_t)__umul64hi(r, (uint64_t)avail));
      u = et_u01(et_draw(key, (uint32_t)(nd.dc + 2 * lane + 1)));
    }
    const uint32_t same = __match_any_sync(0xffffffffu, pick);
    if (lane < nb && lane == __ffs(same) - 1) f = pick;  // a duplicate of a lower lane's pick sits the batch out
    __syncwarp();
    if (f >= 0) atomicOr(&taken[f >> 5], 1u << (f & 31));
  }
  cd.feat[lane] = f;
  cd.u[lane] = u;
  cd.flags[lane] = fl;
  cd.mnT[lane] = 0xffffffffu;
  cd.mxB[lane] = 0u;
  cd.mnB[lane] = 0xffffffffu;
  cd.mn[lane] = dkey(1.7976931348623157e308);  // pkg:35-36: the reference's sentinels
  cd.mx[lane] = dkey(-1.7976931348623157e308);
  if (TASK == TASK_CLS) {
    int32_t *h = w.hist + (int64_t)q * 64 * C;
    for (int t = lane; t < 64 * C; t += 32) h[t] = 0;
  }
  __syncwarp();
  if (lane == 0) {
    cd.nanmask = 0u;
    nd.nb = nb;
    nd.nsweep = 1;
    if (p.replay)
      nd.tpos += nb;
    else
      nd.dc += 2 * p.NB;
    atomicAdd(w.pending, 1ull);
  }
}

template <int TASK>
__global__ void __launch_bounds__(128) k_wide_first(P p, WState w) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (q >= w.count) return;
  WNode &nd = w.node[q];
  const int i = nd.fi, n = nd.n, C = p.C, W = p.W;
  const int32_t depth = p.cur.depth[i];
  bool leaf;
  double total = 0.0, leaf_mean = 0.0, reg_mu = 0.0, reg_S = 0.0, reg_Q = 0.0;
  if (TASK == TASK_CLS) {
    const int32_t *hn = p.cur.hist + (int64_t)i * C;
    bool pure_l = false;
    for (int c = lane; c < C; c += 32) pure_l |= (hn[c] == n);
    const bool pure = __any_sync(0xffffffffu, pure_l);
    leaf = (p.n_table < p.n_min) || (depth >= p.max_depth) || pure;  // pkg:993-994
    if (!leaf) {
      // giniImpurity with the reference's repeated `+= 1/s` distribution (pkg:905-911, 1160-1180)
      const double inv = ET_DIV(1.0, (double)n);
      double s = 0.0;
      for (int c0 = 0; c0 < C; c0 += 32) {
        const int c = c0 + lane;
        const double pc = (c < C) ? repeat_add_dev(inv, hn[c]) : 0.0;
        const double sq = ET_MUL(pc, pc);
        for (int l2 = 0; l2 < min(32, C - c0); l2++) s = ET_ADD(s, __shfl_sync(0xffffffffu, sq, l2));
      }
      total = ET_SUB(1.0, s);
    }
  } else {
    leaf = (n < p.n_min) || (depth >= p.max_depth) || !nd.not_uniform;  // pkg:813-814
    reg_mu = wide_node_mean(w, q, n);
    // moments about the mean: the chunk sums in chunk order
    for (int64_t g = w.chunk0[q]; g < w.chunk0[q + 1]; g++) {
      reg_S = ET_ADD(reg_S, w.part[g * 192]);
      reg_Q = ET_ADD(reg_Q, w.part[g * 192 + 1]);
    }
    const double dn = (double)n;
    leaf_mean = reg_mu;
    total = ET_DIV(ET_SUB(reg_Q, ET_DIV(ET_MUL(reg_S, reg_S), dn)), dn);
  }
  __syncwarp();
  if (lane == 0) {
    nd.flags = leaf ? 3 : 0;
    nd.total = total;
    nd.nsum = (double)n;
    nd.leaf_mean = leaf_mean;
    nd.reg_mu = reg_mu;
    nd.reg_S = reg_S;
    nd.reg_Q = reg_Q;
    nd.visited = 0;
    nd.dc = 0;
    nd.tpos = 0;
    nd.tcnt = 0;
    nd.tb = 0;
    nd.best_feature = -1;
    nd.best_mil = 0;
    nd.best_nleft = 0;
    nd.best_thr = 0;
    nd.best_K = 0;
    nd.best_score = -INFINITY;
    nd.second_score = -INFINITY;
    nd.best_cut = NAN;
    nd.st_draws = nd.st_const = nd.st_scored = nd.st_mismatch = 0;
    nd.slot = -1;
    nd.nconst = 0;
    if (p.replay) {
      const int64_t tn = p.cur.trace[i];
      if (tn >= 0) {
        nd.tb = p.tr.cand_begin[tn];
        nd.tcnt = p.tr.cand_count[tn];
      }
    }
  }
  if (!p.replay) {
    int nc = 0;
    for (int ww = lane; ww < W; ww += 32) {
      const uint32_t m = p.cur.mask[(int64_t)i * W + ww];
      w.cmask[(int64_t)q * W + ww] = m;
      w.taken[(int64_t)q * W + ww] = m;
      nc += __popc(m);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nc += __shfl_xor_sync(0xffffffffu, nc, o);
    if (lane == 0) nd.nconst = nc - (W * 32 - p.d);
  }
  __syncwarp();
  __threadfence_block();
  if (!leaf) wide_draw<TASK>(p, w, q, lane);
}

// ---- pass 1: min / max / hasMissing per candidate over the chunk's rows ----------------------------------------
// Shared memory of a chunk CTA: the chunk's rows (and labels), staged once.
template <bool CODED>
__global__ void __launch_bounds__(WT) k_wide_pass1(P p, WState w) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) int64_t s_coloff[32];
  __shared__ uint32_t s_cred[(WT / 32) * 24];
  __shared__ __align__(4) uint8_t s_Kb[32];
  __shared__ double s_mm[(WT / 32) * 8];
  __shared__ uint32_t s_nan[WT / 32];
  __shared__ __align__(8) uint64_t s_bar;
  ChunkRef r;
  if (!chunk_ref(w, blockIdx.x, r)) return;
  const WNode &nd = w.node[r.q];
  const int nb = nd.nb;
  if ((nd.flags & 2) || nb <= 0) return;
  WCand &cd = w.cand[r.q];
  const int tid = threadIdx.x, lane = tid & 31, wit = tid >> 5;
  if (w.bulk) {
    if (tid == 0) mbar_init(&s_bar, 1);
    __syncthreads();
  }
  const int32_t *s_rows = stage_rows(reinterpret_cast<int32_t *>(smem_raw), p.idx_src + (int64_t)nd.tree * p.n + nd.b + r.j0,
                                     r.cnt, &s_bar, w.bulk, tid);
  if (CODED) {
    if (tid < 32) {
      const int32_t f = (tid < nb) ? cd.feat[tid] : -1;
      s_coloff[tid] = (int64_t)(f >= 0 ? f : 0) * p.ldc;
      s_Kb[tid] = (f >= 0 && p.coff[f] == 0) ? 1 : 0;  // wide code - 1 = byte - K (mod 256)
    }
    if (w.bulk) mbar_wait(&s_bar, 0);
    __syncthreads();
    const int ng = (nb + 3) >> 2;
    const uint32_t *s_K4 = reinterpret_cast<const uint32_t *>(s_Kb);
    if (ng <= 2)
      coded_pass1<2, WT>(p.C8, s_coloff, s_K4, s_rows, r.cnt, tid, s_cred, wit, lane, nullptr);
    else if (ng <= 4)
      coded_pass1<4, WT>(p.C8, s_coloff, s_K4, s_rows, r.cnt, tid, s_cred, wit, lane, nullptr);
    else
      coded_pass1<8, WT>(p.C8, s_coloff, s_K4, s_rows, r.cnt, tid, s_cred, wit, lane, nullptr);
    __syncthreads();
    if (wit == 0 && lane < nb && cd.feat[lane] >= 0) {
      const int c = lane, g = c >> 2, sh = 8 * (c & 3);
      uint32_t mnt = 255u, mxb = 0u, mnb = 255u;
      for (int w2 = 0; w2 < WT / 32; w2++) {
        mnt = min(mnt, (s_cred[w2 * 24 + g] >> sh) & 255u);
        mxb = max(mxb, (s_cred[w2 * 24 + 8 + g] >> sh) & 255u);
        mnb = min(mnb, (s_cred[w2 * 24 + 16 + g] >> sh) & 255u);
      }
      atomicMin(&cd.mnT[c], mnt);
      atomicMax(&cd.mxB[c], mxb);
      atomicMin(&cd.mnB[c], mnb);
    }
  } else {
    if (w.bulk) mbar_wait(&s_bar, 0);
    __syncthreads();
    if (p.csc_row) {
      // sparse table: one warp per candidate walks the column's entries inside the chunk's row range
      const int32_t rmin = s_rows[0], rmax = s_rows[r.cnt - 1];
      const int32_t *inv_tree = w.inv + (int64_t)nd.tree * p.n, *idx_tree = p.idx_src + (int64_t)nd.tree * p.n;
      const int32_t c0 = nd.b + r.j0;
      for (int c = wit; c < nb; c += WT / 32) {
        const int32_t f = cd.feat[c];
        if (f < 0) continue;
        int64_t lo, hi;
        col_range(p, f, rmin, rmax, lo, hi);
        double mn = 1.7976931348623157e308, mx = -1.7976931348623157e308;  // pkg:35-36
        int32_t members = 0;
        bool nan = false;
        for (int64_t t = lo + lane; t < hi; t += 64) {  // two entries per trip: their loads overlap
          const bool two = t + 32 < hi;
          const int32_t ra = __ldg(p.csc_row + t), rb = two ? __ldg(p.csc_row + t + 32) : -1;
          const double va = __ldg(p.csc_val + t), vb = two ? __ldg(p.csc_val + t + 32) : 0.0;
          if (chunk_pos_inv(inv_tree, idx_tree, c0, r.cnt, ra) >= 0) {
            members++;
            if (va < mn) mn = va;
            if (va > mx) mx = va;
            nan |= (va != va);
          }
          if (two && chunk_pos_inv(inv_tree, idx_tree, c0, r.cnt, rb) >= 0) {
            members++;
            if (vb < mn) mn = vb;
            if (vb > mx) mx = vb;
            nan |= (vb != vb);
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double omn = __shfl_xor_sync(0xffffffffu, mn, o), omx = __shfl_xor_sync(0xffffffffu, mx, o);
          if (omn < mn) mn = omn;
          if (omx > mx) mx = omx;
          members += __shfl_xor_sync(0xffffffffu, members, o);
        }
        nan = __any_sync(0xffffffffu, nan);
        if (members < r.cnt) {  // the other rows of the chunk are implicit zeros
          if (0.0 < mn) mn = 0.0;
          if (0.0 > mx) mx = 0.0;
        }
        if (lane == 0) {
          atomicMin(&cd.mn[c], dkey(mn));
          atomicMax(&cd.mx[c], dkey(mx));
          if (nan) atomicOr(&cd.nanmask, 1u << c);
        }
      }
      return;
    }
    double *park = (w.park && nb <= w.park_stride) ? w.park + (int64_t)blockIdx.x * w.park_stride * w.chunk : nullptr;
    // groups of 4 candidates, four rows per thread and trip: 16 independent gathers in flight
    for (int g0 = 0; g0 < nb; g0 += 4) {
      Col colp[4];
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const int32_t f = (g0 + c < nb) ? cd.feat[g0 + c] : -1;
        colp[c] = col_of(p, f >= 0 ? f : 0);
      }
      double mn[4], mx[4];
      uint32_t nanm = 0u;
#pragma unroll
      for (int c = 0; c < 4; c++) {
        mn[c] = 1.7976931348623157e308;  // pkg:35-36
        mx[c] = -1.7976931348623157e308;
      }
      for (int32_t j0 = tid; j0 < r.cnt; j0 += 4 * WT) {
        int32_t r4[4];
        double x[4][4];
#pragma unroll
        for (int u2 = 0; u2 < 4; u2++) {
          const int32_t j = j0 + u2 * WT;
          r4[u2] = (j < r.cnt) ? s_rows[j] : -1;
        }
#pragma unroll
        for (int u2 = 0; u2 < 4; u2++)
#pragma unroll
          for (int c = 0; c < 4; c++) x[u2][c] = (r4[u2] >= 0) ? col_at(colp[c], r4[u2]) : NAN;
        if (park) {
#pragma unroll
          for (int u2 = 0; u2 < 4; u2++)
#pragma unroll
            for (int c = 0; c < 4; c++)
              if (r4[u2] >= 0 && g0 + c < nb) park[(int64_t)(g0 + c) * w.chunk + j0 + u2 * WT] = x[u2][c];
        }
#pragma unroll
        for (int u2 = 0; u2 < 4; u2++) {
#pragma unroll
          for (int c = 0; c < 4; c++) {
            if (x[u2][c] < mn[c]) mn[c] = x[u2][c];
            if (x[u2][c] > mx[c]) mx[c] = x[u2][c];
            nanm |= (uint32_t)(r4[u2] >= 0 && x[u2][c] != x[u2][c]) << c;
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 4; c++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double omn = __shfl_xor_sync(0xffffffffu, mn[c], o), omx = __shfl_xor_sync(0xffffffffu, mx[c], o);
          if (omn < mn[c]) mn[c] = omn;
          if (omx > mx[c]) mx[c] = omx;
        }
      }
      nanm = __reduce_or_sync(0xffffffffu, nanm);
      __syncthreads();  // previous users of the scratch are done
      if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
          s_mm[wit * 8 + c] = mn[c];
          s_mm[wit * 8 + 4 + c] = mx[c];
        }
        s_nan[wit] = nanm;
      }
      __syncthreads();
      if (tid < 4 && g0 + tid < nb && cd.feat[g0 + tid] >= 0) {
        const int c = tid, ci = g0 + tid;
        double a = 1.7976931348623157e308, bq = -1.7976931348623157e308;
        uint32_t has_nan = 0u;
        for (int w2 = 0; w2 < WT / 32; w2++) {
          const double v1 = s_mm[w2 * 8 + c], v2 = s_mm[w2 * 8 + 4 + c];
          if (v1 < a) a = v1;
          if (v2 > bq) bq = v2;
          has_nan |= (s_nan[w2] >> c) & 1u;
        }
        atomicMin(&cd.mn[ci], dkey(a));
        atomicMax(&cd.mx[ci], dkey(bq));
        if (has_nan) atomicOr(&cd.nanmask, 1u << ci);
      }
    }
  }
}

// ---- decide: constant test, cutpoint, code threshold (lane == candidate) -----------------------------------------
template <bool CODED>
__global__ void __launch_bounds__(128) k_wide_decide(P p, WState w) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (q >= w.count) return;
  WNode &nd = w.node[q];
  const int nb = nd.nb;
  if ((nd.flags & 2) || nb <= 0) return;
  WCand &cd = w.cand[q];
  const int c = lane;
  const int32_t f = (c < nb) ? cd.feat[c] : -1;
  bool nan_c = false;
  int32_t fl = cd.flags[c];
  uint8_t thrb = 0, enb = 0, nanb = 0, Kb = 0;
  if (f >= 0) {
    double mn = 1.7976931348623157e308, mx = -1.7976931348623157e308;  // pkg:35-36
    bool has_nan;
    uint32_t mnt = 0u, wmax = 0u;
    const double *dc8 = CODED ? p.dict + (int64_t)f * 256 : nullptr;
    if (CODED) {
      const uint32_t K = (p.coff[f] == 0) ? 1u : 0u;
      Kb = (uint8_t)K;
      mnt = min(cd.mnT[c], 255u);
      const uint32_t mxb = cd.mxB[c], mnb = min(cd.mnB[c], 255u);
      wmax = mxb + (1u - K);  // largest wide code (0 = only NaNs)
      has_nan = (K == 1u) && (mnb == 0u);
      if (wmax != 0u) {
        mn = __ldg(dc8 + mnt);
        mx = __ldg(dc8 + (wmax - 1u));
      }
    } else {
      has_nan = (cd.nanmask >> c) & 1u;
      mn = dkey_inv(cd.mn[c]);
      mx = dkey_inv(cd.mx[c]);
    }
    if (mx <= mn && !has_nan) {  // pkg:236
      fl |= CF_CONST;
    } else {
      const double cut = ET_ADD(mn, ET_MUL(ET_SUB(mx, mn), cd.u[c]));  // nextDouble(min, max), pkg:240
      cd.cut[c] = cut;
      if (CODED) {
        uint32_t thr = 0u;  // number of dictionary entries below the cutpoint
        if (wmax != 0u) {
          uint32_t lo = mnt, hi = wmax;
          while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (__ldg(dc8 + mid) < cut)
              lo = mid + 1u;
            else
              hi = mid;
          }
          thr = lo;
        }
        cd.thr[c] = (int32_t)thr;
        thrb = (uint8_t)(thr > 0u ? thr - 1u : 0u);
        enb = thr > 0u ? 0xff : 0;
      }
      if (has_nan) {
        fl |= CF_NAN;
        nanb = 0xff;
        nan_c = true;
      }
    }
  }
  cd.flags[c] = fl;
  if (CODED) {
    cd.thrb[c] = thrb;
    cd.enb[c] = enb;
    cd.nanb[c] = nanb;
    cd.Kb[c] = Kb;
  }
  const bool any_nan = __any_sync(0xffffffffu, nan_c);
  if (lane == 0) nd.nsweep = any_nan ? 2 : 1;
}

// ---- pass 2 ------------------------------------------------------------------------------------------------------
// FP64 classification: like coded_pass2, every thread packs the side bits of its row for all candidates into one
// word, a 32 x 32 bit transpose hands lane c the 32 rows' bits of candidate c, counted against the class ballots.
template <typename LabFn>
__device__ __forceinline__ void fp64_pass2(const P &p, const int32_t *s_feat, const double *s_cut, const uint32_t act,
                                           int sweep, const int32_t *rr, LabFn lab, int32_t n, int C, int wit, int lane,
                                           int32_t *s_hist, int hs, int hoff, int nb, const double *park, int chunk) {
  int32_t acc[32];
#pragma unroll
  for (int k = 0; k < 32; k++) acc[k] = 0;
  for (int32_t j0 = wit * 32; j0 < n; j0 += WT) {
    const int32_t j = j0 + lane;
    const bool valid = j < n;
    const int32_t cls = valid ? lab(j) : -1;
    const int32_t row = valid ? rr[j] : 0;
    uint32_t rowbits = 0u;
    if (park) {  // the values pass 1 gathered: candidate-major, a warp reads 32 consecutive rows of one candidate
#pragma unroll 8
      for (int c = 0; c < nb; c++) {
        const double x = valid ? park[(int64_t)c * chunk + j] : 0.0;
        const bool in = sweep ? (x != x) : (x < s_cut[c]);
        rowbits |= (uint32_t)in << c;
      }
    } else {
#pragma unroll 8
      for (int c = 0; c < nb; c++) {
        const double x = col_at(col_of(p, s_feat[c]), row);
        const bool in = sweep ? (x != x) : (x < s_cut[c]);
        rowbits |= (uint32_t)in << c;
      }
    }
    rowbits &= act;
    if (!valid) rowbits = 0u;
    const uint32_t candbits = warp_transpose32(rowbits, lane);
#pragma unroll
    for (int k = 0; k < 32; k++) {
      if (k >= C) break;
      const uint32_t cmk = __ballot_sync(0xffffffffu, cls == k);
      acc[k] += __popc(candbits & cmk);
    }
  }
  if (lane < nb) {
#pragma unroll
    for (int k = 0; k < 32; k++) {
      if (k >= C) break;
      if (acc[k]) atomicAdd(&s_hist[lane * hs + hoff + k], acc[k]);
    }
  }
}

template <int TASK, bool CODED>
__global__ void __launch_bounds__(WT) k_wide_pass2(P p, WState w) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) int64_t s_coloff[32];
  __shared__ __align__(4) uint8_t s_par[128];  // thrb, enb, Kb, nanb
  __shared__ int32_t s_feat[32];
  __shared__ double s_cut[32];
  __shared__ double s_red[(WT / 32) * 12];
  __shared__ int32_t s_redi[(WT / 32) * 4];
  ChunkRef r;
  if (!chunk_ref(w, blockIdx.x, r)) return;
  const WNode &nd = w.node[r.q];
  const int nb = nd.nb;
  if ((nd.flags & 2) || nb <= 0) return;
  const WCand &cd = w.cand[r.q];
  const int tid = threadIdx.x, lane = tid & 31, wit = tid >> 5, C = p.C;
  const int hs = (2 * C) | 1;
  __shared__ __align__(8) uint64_t s_bar;
  int32_t *s_hist = reinterpret_cast<int32_t *>(smem_raw) + w.chunk + 8;  // [32][hs] (classification)
  uint8_t *s_lab8 = reinterpret_cast<uint8_t *>(s_hist + 32 * hs);        // [chunk]
  const int64_t seg = (int64_t)nd.tree * p.n + nd.b + r.j0;
  if (w.bulk) {
    if (tid == 0) mbar_init(&s_bar, 1);
    __syncthreads();
  }
  const int32_t *s_rows = stage_rows(reinterpret_cast<int32_t *>(smem_raw), p.idx_src + seg, r.cnt, &s_bar, w.bulk, tid);
  if (TASK == TASK_CLS)
    for (int32_t j = tid; j < r.cnt; j += WT) s_lab8[j] = (uint8_t)p.yc_src[seg + j];
  if (TASK == TASK_CLS)
    for (int t = tid; t < nb * hs; t += WT) s_hist[t] = 0;
  uint32_t act = 0u, nanact = 0u;  // candidates to count: scored ones / those with NaNs
  if (tid < 32) {
    const int32_t f = (tid < nb) ? cd.feat[tid] : -1;
    const int32_t fl = (tid < nb) ? cd.flags[tid] : CF_CONST;
    s_feat[tid] = f >= 0 ? f : 0;
    s_cut[tid] = cd.cut[tid];
    if (CODED) {
      s_coloff[tid] = (int64_t)(f >= 0 ? f : 0) * p.ldc;
      s_par[tid] = cd.thrb[tid];
      s_par[32 + tid] = cd.enb[tid];
      s_par[64 + tid] = cd.Kb[tid];
      s_par[96 + tid] = cd.nanb[tid];
    }
    act = __ballot_sync(0xffffffffu, f >= 0 && !(fl & CF_CONST));
    nanact = __ballot_sync(0xffffffffu, f >= 0 && !(fl & CF_CONST) && (fl & CF_NAN));
    if (tid == 0) {
      s_redi[0] = (int32_t)act;
      s_redi[1] = (int32_t)nanact;
    }
  }
  if (w.bulk) mbar_wait(&s_bar, 0);
  __syncthreads();
  act = (uint32_t)s_redi[0];
  nanact = (uint32_t)s_redi[1];
  __syncthreads();
  const int nsweep = nd.nsweep;
  const double *park = (!CODED && w.park && nb <= w.park_stride) ? w.park + (int64_t)blockIdx.x * w.park_stride * w.chunk : nullptr;
  if (TASK == TASK_CLS && !CODED && p.csc_row) {
    // sparse table: class histogram of the chunk once, then per candidate the histograms of the STORED rows (left /
    // NaN / all); the rows the column does not store are zeros and go left iff 0 < cut, by count
    __shared__ int32_t s_hc[32];
    __shared__ int32_t s_wh[WT / 32][3][32];
    if (tid < 32) s_hc[tid] = 0;
    __syncthreads();
    {
      int32_t acc = 0;  // lane k counts class k over this warp's rows
      for (int32_t j0 = wit * 32; j0 < r.cnt; j0 += WT) {
        const int32_t j = j0 + lane;
        const int32_t cls = (j < r.cnt) ? (int32_t)s_lab8[j] : -1;
        for (int k = 0; k < C; k++) {
          const uint32_t m = __ballot_sync(0xffffffffu, cls == k);
          if (lane == k) acc += __popc(m);
        }
      }
      if (lane < C && acc) atomicAdd(&s_hc[lane], acc);
    }
    __syncthreads();
    const int32_t rmin = s_rows[0], rmax = s_rows[r.cnt - 1];
    const int32_t *inv_tree = w.inv + (int64_t)nd.tree * p.n, *idx_tree = p.idx_src + (int64_t)nd.tree * p.n;
    const int32_t c0 = nd.b + r.j0;
    int32_t *gh = w.hist + (int64_t)r.q * 64 * C;
    for (int c = wit; c < nb; c += WT / 32) {
      if (!((act >> c) & 1u)) continue;
      s_wh[wit][0][lane] = 0;
      s_wh[wit][1][lane] = 0;
      s_wh[wit][2][lane] = 0;
      __syncwarp();
      const double cut = s_cut[c];
      int64_t lo, hi;
      col_range(p, s_feat[c], rmin, rmax, lo, hi);
      for (int64_t t = lo + lane; t < hi; t += 64) {  // two entries per trip: their loads overlap
        const bool two = t + 32 < hi;
        const int32_t r2[2] = {__ldg(p.csc_row + t), two ? __ldg(p.csc_row + t + 32) : -1};
        const double v2[2] = {__ldg(p.csc_val + t), two ? __ldg(p.csc_val + t + 32) : 0.0};
#pragma unroll
        for (int u2 = 0; u2 < 2; u2++) {
          const int pos = (u2 == 0 || two) ? chunk_pos_inv(inv_tree, idx_tree, c0, r.cnt, r2[u2]) : -1;
          if (pos >= 0) {
            const double v = v2[u2];
            const int cls = (int)s_lab8[pos];
            atomicAdd(&s_wh[wit][2][cls], 1);
            if (v != v)
              atomicAdd(&s_wh[wit][1][cls], 1);
            else if (v < cut)
              atomicAdd(&s_wh[wit][0][cls], 1);
          }
        }
      }
      __syncwarp();
      if (lane < C) {
        const int32_t hl = s_wh[wit][0][lane] + ((0.0 < cut) ? s_hc[lane] - s_wh[wit][2][lane] : 0);
        const int32_t hn = s_wh[wit][1][lane];
        if (hl) atomicAdd(&gh[c * 2 * C + lane], hl);
        if (hn) atomicAdd(&gh[c * 2 * C + C + lane], hn);
      }
      __syncwarp();
    }
  } else if (TASK == TASK_CLS) {
    auto lab = [&](int32_t j) -> int32_t { return (int32_t)s_lab8[j]; };
    for (int sweep = 0; sweep < nsweep; sweep++) {
      const int hoff = sweep ? C : 0;
      if (CODED) {
        const int ng = (nb + 3) >> 2;
        const uint32_t *s_K4 = reinterpret_cast<const uint32_t *>(s_par + 64);
        const uint32_t *s_t4 = reinterpret_cast<const uint32_t *>(s_par);
        const uint32_t *s_e4 = reinterpret_cast<const uint32_t *>(sweep ? s_par + 96 : s_par + 32);
        if (ng <= 2)
          coded_pass2<2, WT>(p.C8, s_coloff, s_K4, s_t4, s_e4, sweep, s_rows, lab, r.cnt, C, wit, lane, s_hist, hs, hoff, nb,
                             nullptr);
        else if (ng <= 4)
          coded_pass2<4, WT>(p.C8, s_coloff, s_K4, s_t4, s_e4, sweep, s_rows, lab, r.cnt, C, wit, lane, s_hist, hs, hoff, nb,
                             nullptr);
        else
          coded_pass2<8, WT>(p.C8, s_coloff, s_K4, s_t4, s_e4, sweep, s_rows, lab, r.cnt, C, wit, lane, s_hist, hs, hoff, nb,
                             nullptr);
      } else {
        fp64_pass2(p, s_feat, s_cut, sweep ? nanact : act, sweep, s_rows, lab, r.cnt, C, wit, lane, s_hist, hs, hoff, nb,
                   park, w.chunk);
      }
    }
    __syncthreads();
    int32_t *gh = w.hist + (int64_t)r.q * 64 * C;  // [32][2C]
    for (int t = tid; t < nb * 2 * C; t += WT) {
      const int c = t / (2 * C), k = t - c * 2 * C;
      const int32_t v = s_hist[c * hs + k];
      if (v) atomicAdd(&gh[c * 2 * C + k], v);
    }
  } else {
    // regression: per group of 4 candidates the moments (count, S, Q about the node mean) of the rows going left;
    // the second sweep counts the NaN rows (they join the left side when missing-is-less is evaluated)
    const double *yy = p.yr_src + seg;
    const double mu = nd.reg_mu;
    double *part = w.part + (int64_t)blockIdx.x * 192;
    for (int sweep = 0; sweep < nsweep; sweep++) {
      for (int g0 = 0; g0 < nb; g0 += 4) {
        const uint32_t gact = ((sweep ? nanact : act) >> g0) & 15u;
        if (gact == 0u) {
          if (tid < 4) {
            part[(sweep * 32 + g0 + tid) * 3] = 0.0;
            part[(sweep * 32 + g0 + tid) * 3 + 1] = 0.0;
            part[(sweep * 32 + g0 + tid) * 3 + 2] = 0.0;
          }
          continue;
        }
        Col colp[4];
        double cut[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
          colp[c] = col_of(p, s_feat[min(g0 + c, 31)]);
          cut[c] = s_cut[min(g0 + c, 31)];
        }
        int32_t cnt[4];
        double S[4], Q[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
          cnt[c] = 0;
          S[c] = 0.0;
          Q[c] = 0.0;
        }
        for (int32_t j0 = tid; j0 < r.cnt; j0 += 4 * WT) {
          int32_t r4[4];
          double x[4][4], yd[4];
#pragma unroll
          for (int u2 = 0; u2 < 4; u2++) {
            const int32_t j = j0 + u2 * WT;
            r4[u2] = (j < r.cnt) ? s_rows[j] : -1;
            yd[u2] = (j < r.cnt) ? ET_SUB(yy[j], mu) : 0.0;
          }
          if (park) {
#pragma unroll
            for (int u2 = 0; u2 < 4; u2++)
#pragma unroll
              for (int c = 0; c < 4; c++)
                x[u2][c] = (r4[u2] >= 0 && g0 + c < nb) ? park[(int64_t)(g0 + c) * w.chunk + j0 + u2 * WT] : 0.0;
          } else {
#pragma unroll
            for (int u2 = 0; u2 < 4; u2++)
#pragma unroll
              for (int c = 0; c < 4; c++) x[u2][c] = (r4[u2] >= 0) ? col_at(colp[c], r4[u2]) : 0.0;
          }
#pragma unroll
          for (int u2 = 0; u2 < 4; u2++) {
            const double yd2 = ET_MUL(yd[u2], yd[u2]);
#pragma unroll
            for (int c = 0; c < 4; c++) {
              const bool in = (r4[u2] >= 0) && (sweep ? (x[u2][c] != x[u2][c]) : (x[u2][c] < cut[c]));
              cnt[c] += in ? 1 : 0;
              S[c] = ET_ADD(S[c], in ? yd[u2] : 0.0);
              Q[c] = ET_ADD(Q[c], in ? yd2 : 0.0);
            }
          }
        }
#pragma unroll
        for (int c = 0; c < 4; c++) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            cnt[c] += __shfl_xor_sync(0xffffffffu, cnt[c], o);
            S[c] = ET_ADD(S[c], __shfl_xor_sync(0xffffffffu, S[c], o));
            Q[c] = ET_ADD(Q[c], __shfl_xor_sync(0xffffffffu, Q[c], o));
          }
        }
        __syncthreads();
        if (lane == 0) {
#pragma unroll
          for (int c = 0; c < 4; c++) {
            s_red[wit * 12 + c] = S[c];
            s_red[wit * 12 + 4 + c] = Q[c];
            s_redi[wit * 4 + c] = cnt[c];
          }
        }
        __syncthreads();
        if (tid < 4) {
          const int c = tid;
          int32_t ni = 0;
          double Si = 0.0, Qi = 0.0;
          for (int w2 = 0; w2 < WT / 32; w2++) {
            ni += s_redi[w2 * 4 + c];
            Si = ET_ADD(Si, s_red[w2 * 12 + c]);
            Qi = ET_ADD(Qi, s_red[w2 * 12 + 4 + c]);
          }
          const bool on = (gact >> c) & 1u;
          part[(sweep * 32 + g0 + c) * 3] = on ? (double)ni : 0.0;
          part[(sweep * 32 + g0 + c) * 3 + 1] = on ? Si : 0.0;
          part[(sweep * 32 + g0 + c) * 3 + 2] = on ? Qi : 0.0;
        }
      }
    }
  }
}

// ---- score + consume (lane == candidate), then the next batch ----------------------------------------------------
template <int TASK>
__global__ void __launch_bounds__(128) k_wide_score(P p, WState w) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (q >= w.count) return;
  WNode &nd = w.node[q];
  const int nb = nd.nb;
  if ((nd.flags & 2) || nb <= 0) return;
  WCand &cd = w.cand[q];
  const int C = p.C, n = nd.n, i = nd.fi;
  const int32_t f = (lane < nb) ? cd.feat[lane] : -1;
  int32_t fl = cd.flags[lane];
  const bool act0 = f >= 0;
  const bool const0 = act0 && (fl & CF_CONST);
  double s = NAN;
  int32_t nleft = 0;
  if (act0 && !const0) {
    const bool has_nan = (fl & CF_NAN) != 0;
    double sn, sl = NAN;
    int32_t nin_n = 0, nin_l = 0;
    if (TASK == TASK_CLS) {
      const int32_t *hnode = p.cur.hist + (int64_t)i * C;
      const int32_t *hl = w.hist + (int64_t)q * 64 * C + lane * 2 * C, *hn = hl + C;
      sn = gini_score_int(hnode, hl, hn, false, C, n, nd.total, &nin_n);
      if (has_nan) sl = gini_score_int(hnode, hl, hn, true, C, n, nd.total, &nin_l);
    } else {
      // the chunks' moments in chunk order
      double c0 = 0.0, S0 = 0.0, Q0 = 0.0, c1 = 0.0, S1 = 0.0, Q1 = 0.0;
      for (int64_t g = w.chunk0[q]; g < w.chunk0[q + 1]; g++) {
        const double *pt = w.part + g * 192 + lane * 3;
        c0 = ET_ADD(c0, pt[0]);
        S0 = ET_ADD(S0, pt[1]);
        Q0 = ET_ADD(Q0, pt[2]);
        if (has_nan) {
          c1 = ET_ADD(c1, pt[96]);
          S1 = ET_ADD(S1, pt[97]);
          Q1 = ET_ADD(Q1, pt[98]);
        }
      }
      nin_n = (int32_t)c0;
      sn = var_reduction_moments(n, nd.reg_S, nd.reg_Q, nd.total, nin_n, S0, Q0);
      if (has_nan) {
        nin_l = nin_n + (int32_t)c1;
        sl = var_reduction_moments(n, nd.reg_S, nd.reg_Q, nd.total, nin_l, ET_ADD(S0, S1), ET_ADD(Q0, Q1));
      }
    }
    const bool mil = !(sl != sl) && (sl > sn || (sn != sn));  // pkg:272-275
    s = mil ? sl : sn;
    nleft = mil ? nin_l : nin_n;
    if (mil) fl |= CF_MIL;
  }
  // ---- consume the batch in draw (lane) order; the reference stops drawing once k candidates have been scored
  const int32_t visited = nd.visited;
  const bool counted0 = act0 && !const0 && !(s != s);
  const uint32_t m_cnt0 = __ballot_sync(0xffffffffu, counted0);
  const bool act = act0 && (p.replay || __popc(m_cnt0 & ((1u << lane) - 1u)) < p.k - visited);
  const bool is_const = act && const0;
  const bool is_nan = act && !const0 && (s != s);
  const bool counted = act && counted0;
  const uint32_t m_act = __ballot_sync(0xffffffffu, act);
  const uint32_t m_const = __ballot_sync(0xffffffffu, is_const);
  const uint32_t m_nan = __ballot_sync(0xffffffffu, is_nan);
  const uint32_t m_cnt = __ballot_sync(0xffffffffu, counted);
  unsigned long long mism = 0;
  if (p.replay) {
    const int exp = (fl >> 4) & 3;
    const bool bad = act && ((is_const && exp != 1) || (is_nan && exp != 3) || (counted && exp != 2));
    mism = __popc(__ballot_sync(0xffffffffu, bad));
  }
  double bs = counted ? s : -INFINITY;
  int bl = counted ? lane : 64;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double os = __shfl_xor_sync(0xffffffffu, bs, o);
    const int ol = __shfl_xor_sync(0xffffffffu, bl, o);
    if (os > bs || (os == bs && ol < bl)) {
      bs = os;
      bl = ol;
    }
  }
  double best_score = nd.best_score, second = nd.second_score;
  if (TASK == TASK_REG) {  // runner-up over everything seen so far (ambiguity count)
    double b2 = (counted && lane != bl) ? s : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b2 = fmax(b2, __shfl_xor_sync(0xffffffffu, b2, o));
    if (bl < 32) {
      if (bs > best_score)
        second = fmax(best_score, fmax(second, b2));
      else
        second = fmax(second, bs);
    }
  }
  const bool better = bl < 32 && bs > best_score;  // strict >: the first best wins (pkg:277)
  if (better) {
    const int32_t bfl = __shfl_sync(0xffffffffu, fl, bl);
    const int32_t bnl = __shfl_sync(0xffffffffu, nleft, bl);
    if (TASK == TASK_CLS) {
      const int32_t *hl = w.hist + (int64_t)q * 64 * C + bl * 2 * C;
      for (int c = lane; c < C; c += 32) w.besthl[(int64_t)q * C + c] = hl[c] + ((bfl & CF_MIL) ? hl[C + c] : 0);
    }
    if (lane == 0) {
      nd.best_score = bs;
      nd.best_feature = cd.feat[bl];
      nd.best_cut = cd.cut[bl];
      nd.best_mil = (bfl & CF_MIL) ? 1 : 0;
      nd.best_nleft = bnl;
      nd.best_thr = cd.thr[bl];
      nd.best_K = cd.Kb[bl];
    }
  }
  if (!p.replay && (is_const || is_nan))
    atomicOr(&w.cmask[(int64_t)q * p.W + (f >> 5)], 1u << (f & 31));  // pkg:236-238, 283-285: inherited by the children
  __syncwarp();
  if (lane == 0) {
    nd.second_score = second;
    nd.visited = visited + __popc(m_cnt);
    nd.nconst += __popc(m_const) + __popc(m_nan);
    nd.st_draws += __popc(m_act);
    nd.st_const += __popc(m_const);
    nd.st_scored += __popc(m_cnt) + __popc(m_nan);
    nd.st_mismatch += mism;
    nd.nb = 0;
  }
  __syncwarp();
  __threadfence_block();
  wide_draw<TASK>(p, w, q, lane);
}

// ---- finish: leaf | children --------------------------------------------------------------------------------------
template <int TASK>
__global__ void __launch_bounds__(128) k_wide_finish(P p, WState w) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (q >= w.count) return;
  WNode &nd = w.node[q];
  const int i = nd.fi, n = nd.n, C = p.C, W = p.W, b = nd.b;
  const int32_t tree = nd.tree, node = p.cur.node[i], depth = p.cur.depth[i];
  const int64_t tn = p.cur.trace[i];
  const bool leaf = (nd.flags & 1) != 0;
  const bool make_leaf = leaf || nd.best_feature < 0;  // pkg:293-296
  const int lw = (TASK == TASK_REG) ? 1 : C;
  if (lane == 0) {
    if (!leaf) {
      atomicAdd(&p.cnt->st[ST_SROWS], (unsigned long long)n);
      atomicAdd(&p.cnt->st[ST_VMM], (unsigned long long)n * nd.st_draws);
      atomicAdd(&p.cnt->st[ST_VSC], (unsigned long long)n * nd.st_scored);
      atomicAdd(&p.cnt->st[ST_DRAWS], nd.st_draws);
      atomicAdd(&p.cnt->st[ST_CONST], nd.st_const);
      atomicAdd(&p.cnt->st[ST_SCORED], nd.st_scored);
    }
    if (p.replay) {
      unsigned long long mm = nd.st_mismatch;
      const bool trace_split = tn >= 0 && p.tr.left[tn] >= 0;
      if (trace_split == make_leaf) mm++;
      if (mm) atomicAdd(&p.cnt->st[ST_MISMATCH], mm);
    }
    if (TASK == TASK_REG && !leaf) {
      atomicAdd(&p.cnt->st[ST_PARNODES], 1ull);
      if (nd.best_feature >= 0 && nd.second_score > -INFINITY && nd.best_score > nd.second_score &&
          ET_SUB(nd.best_score, nd.second_score) <= 1e-9 * fmax(fabs(nd.best_score), 1e-300))
        atomicAdd(&p.cnt->st[ST_AMBIG], 1ull);
    }
  }
  if (make_leaf) {
    int32_t ls = 0;
    if (lane == 0) {
      ls = atomicAdd(&p.cnt->n_leaves, 1);
      p.o.feat[node] = -1;
      p.o.child[node] = ls;
      p.o.cut[node] = NAN;
      p.o.tree[node] = tree;
    }
    ls = __shfl_sync(0xffffffffu, ls, 0);
    double *lv = p.o.leaf_vals + (int64_t)ls * lw;
    if (TASK == TASK_CLS) {
      const int32_t *hn = p.cur.hist + (int64_t)i * C;
      const double inv = ET_DIV(1.0, (double)n);
      for (int c = lane; c < C; c += 32) lv[c] = repeat_add_dev(inv, hn[c]);  // pkg:960-964
    } else if (lane == 0) {
      // a leaf this large is rare: its value is the reference's sequential mean (pkg:782), exactly
      const double *yy = p.yr_src + (int64_t)tree * p.n + b;
      double sum = 0.0;
      for (int32_t j = 0; j < n; j++) sum = ET_ADD(sum, yy[j]);
      lv[0] = ET_DIV(sum, (double)n);
    }
    return;
  }
  int32_t slot = 0;
  const int32_t nl = nd.best_nleft;
  if (lane == 0) {
    slot = atomicAdd(&p.cnt->next_f, 2);
    nd.slot = slot;
    nd.flags |= 4;
    const int32_t cl = p.node_base_next + slot;
    const uint64_t key = p.cur.key[i];
    p.o.feat[node] = nd.best_feature | (nd.best_mil ? ET_MIL_BIT : 0);
    p.o.child[node] = cl;
    p.o.cut[node] = nd.best_cut;
    p.o.tree[node] = tree;
#pragma unroll
    for (int side = 0; side < 2; side++) {
      const int32_t s2 = slot + side;
      p.nxt.tree[s2] = tree;
      p.nxt.begin[s2] = side ? b + nl : b;
      p.nxt.end[s2] = side ? b + n : b + nl;
      p.nxt.node[s2] = cl + side;
      // pkg:870 / 884 (sic): the regression right child keeps currentDepth; pkg:1055,1071: +1 both
      p.nxt.depth[s2] = (TASK == TASK_REG && side) ? depth : depth + 1;
      p.nxt.key[s2] = et_child_key(key, side);
      int64_t tc = -1;
      if (p.replay && tn >= 0) tc = side ? p.tr.right[tn] : p.tr.left[tn];
      p.nxt.trace[s2] = tc;
      const int32_t cn = side ? (n - nl) : nl;
      const int qc = size_class(p, cn);
      p.q_nxt[qc][atomicAdd(&p.cnt->q_count[qc], 1)] = s2;
      if (qc == Q_WIDE) atomicAdd(&p.cnt->wide_rows, (unsigned long long)cn);
      if (qc >= Q_CTA) atomicAdd(&p.cnt->big_rows, (unsigned long long)cn);
    }
    atomicAdd(&p.cnt->st[ST_PROWS], (unsigned long long)n);
  }
  slot = __shfl_sync(0xffffffffu, slot, 0);
  if (TASK == TASK_CLS) {
    const int32_t *hn = p.cur.hist + (int64_t)i * C;
    int32_t *hl = p.nxt.hist + (int64_t)slot * C, *hr = hl + C;
    for (int c = lane; c < C; c += 32) {
      const int32_t v = w.besthl[(int64_t)q * C + c];
      hl[c] = v;
      hr[c] = hn[c] - v;
    }
  }
  if (!p.replay) {
    uint32_t *ml = p.nxt.mask + (int64_t)slot * W, *mr = ml + W;
    for (int ww = lane; ww < W; ww += 32) {
      const uint32_t v = w.cmask[(int64_t)q * W + ww];
      ml[ww] = v;
      mr[ww] = v;
    }
  }
}

// ---- stable partition (pkg:1024-1039) over the chunks of a node ---------------------------------------------------
template <bool CODED>
__global__ void __launch_bounds__(WT) k_wide_count(P p, WState w) {
  __shared__ int32_t s_cnt[WT / 32];
  ChunkRef r;
  if (!chunk_ref(w, blockIdx.x, r)) return;
  const WNode &nd = w.node[r.q];
  if (!(nd.flags & 4)) return;
  const int tid = threadIdx.x, lane = tid & 31, wit = tid >> 5;
  const int32_t *rr = p.idx_src + (int64_t)nd.tree * p.n + nd.b + r.j0;
  const int32_t bf = nd.best_feature, thr = nd.best_thr, K = nd.best_K;
  const bool mil = nd.best_mil != 0;
  const double cut = nd.best_cut;
  const uint8_t *c8 = CODED ? p.C8 + (int64_t)bf * p.ldc : nullptr;
  const Col col = CODED ? Col{nullptr, nullptr, 0} : col_of(p, bf);
  uint32_t *bits = w.bits + (int64_t)blockIdx.x * (w.chunk / 32);
  int32_t cntl = 0;
  if (!CODED && p.csc_row) {
    // sparse table: every row starts on the side of an implicit zero, the column's stored entries flip their rows
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_w[WCHUNK_MAX / 32];
    int32_t *s_rows = reinterpret_cast<int32_t *>(smem_raw);
    for (int32_t j = tid; j < r.cnt; j += WT) s_rows[j] = rr[j];
    const bool zero_left = (0.0 < cut);
    const int nwords = (r.cnt + 31) >> 5;
    for (int t = tid; t < nwords; t += WT) {
      const int rem = r.cnt - t * 32;
      s_w[t] = zero_left ? (rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u)) : 0u;
    }
    __syncthreads();
    const int32_t *inv_tree = w.inv + (int64_t)nd.tree * p.n, *idx_tree = p.idx_src + (int64_t)nd.tree * p.n;
    const int32_t c0 = nd.b + r.j0;
    int64_t lo, hi;
    col_range(p, bf, s_rows[0], s_rows[r.cnt - 1], lo, hi);
    for (int64_t t = lo + tid; t < hi; t += WT) {
      const int pos = chunk_pos_inv(inv_tree, idx_tree, c0, r.cnt, __ldg(p.csc_row + t));
      if (pos >= 0) {
        const double x = __ldg(p.csc_val + t);
        const bool left = (x < cut) || (mil && (x != x));
        if (left != zero_left) atomicXor(&s_w[pos >> 5], 1u << (pos & 31));  // (a row is stored at most once per column)
      }
    }
    __syncthreads();
    for (int t = tid; t < nwords; t += WT) {
      bits[t] = s_w[t];
      cntl += __popc(s_w[t]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cntl += __shfl_xor_sync(0xffffffffu, cntl, o);
    if (lane == 0) s_cnt[wit] = cntl;
    __syncthreads();
    if (tid == 0) {
      int32_t t = 0;
      for (int w2 = 0; w2 < WT / 32; w2++) t += s_cnt[w2];
      w.cnt_left[blockIdx.x] = t;
    }
    return;
  }
  for (int32_t j0 = wit * 32; j0 < r.cnt; j0 += WT) {
    const int32_t j = j0 + lane;
    bool left = false;
    if (j < r.cnt) {
      const int32_t row = rr[j];
      if (CODED) {
        const int32_t b8 = (int32_t)__ldg(c8 + row);
        const bool isn = (K == 1) && (b8 == 0);
        left = isn ? mil : (((b8 - K) & 255) < thr);
      } else {
        const double x = col_at(col, row);
        left = (x < cut) || (mil && (x != x));
      }
    }
    const uint32_t bl = __ballot_sync(0xffffffffu, left);
    if (lane == 0) bits[j0 >> 5] = bl;
    cntl += __popc(bl);
  }
  if (lane == 0) s_cnt[wit] = cntl;
  __syncthreads();
  if (tid == 0) {
    int32_t t = 0;
    for (int w2 = 0; w2 < WT / 32; w2++) t += s_cnt[w2];
    w.cnt_left[blockIdx.x] = t;
  }
}

template <int TASK>
__global__ void __launch_bounds__(WT) k_wide_scatter(P p, WState w) {
  __shared__ int32_t s_pre[WT + 1];
  __shared__ uint32_t s_bits[WT];
  __shared__ int32_t s_warp[WT / 32];
  __shared__ int32_t s_before;
  ChunkRef r;
  if (!chunk_ref(w, blockIdx.x, r)) return;
  const WNode &nd = w.node[r.q];
  if (!(nd.flags & 4)) return;
  const int tid = threadIdx.x, lane = tid & 31, wit = tid >> 5;
  // rows going left in the chunks before this one
  int32_t part = 0;
  for (int64_t g = w.chunk0[r.q] + tid; g < (int64_t)blockIdx.x; g += WT) part += w.cnt_left[g];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane == 0) s_warp[wit] = part;
  // exclusive prefix of the left counts of the chunk's 32-row words (one word per thread)
  const int nwords = (r.cnt + 31) >> 5;
  const uint32_t *bits = w.bits + (int64_t)blockIdx.x * (w.chunk / 32);
  uint32_t bw = 0u;
  if (tid < nwords) {
    bw = bits[tid];
    const int rem = r.cnt - tid * 32;
    if (rem < 32) bw &= (1u << rem) - 1u;
  }
  s_bits[tid] = bw;
  int32_t v = __popc(bw);
  const int32_t own = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  __syncthreads();
  if (tid == 0) {
    int32_t t = 0;
    for (int w2 = 0; w2 < WT / 32; w2++) t += s_warp[w2];
    s_before = t;
  }
  __syncthreads();
  if (lane == 31) s_warp[wit] = v;
  __syncthreads();
  int32_t wbase = 0;
  for (int w2 = 0; w2 < wit; w2++) wbase += s_warp[w2];
  s_pre[tid] = wbase + v - own;
  __syncthreads();
  const int32_t before = s_before;
  const int64_t base = (int64_t)nd.tree * p.n;
  const int32_t lbase = nd.b + before, rbase = nd.b + nd.best_nleft + (r.j0 - before);
  const int64_t seg = base + nd.b + r.j0;
  for (int32_t j = tid; j < r.cnt; j += WT) {
    const int wd = j >> 5, bit = j & 31;
    const uint32_t m = s_bits[wd];
    const int32_t lb = s_pre[wd] + __popc(m & ((1u << bit) - 1u));  // lefts before row j in the chunk
    const bool left = (m >> bit) & 1u;
    const int32_t dst = left ? lbase + lb : rbase + (j - lb);
    p.idx_dst[base + dst] = p.idx_src[seg + j];
    if (TASK == TASK_REG)
      p.yr_dst[base + dst] = p.yr_src[seg + j];
    else
      p.yc_dst[base + dst] = p.yc_src[seg + j];
  }
}

// ---- host ----------------------------------------------------------------------------------------------------------
struct WideBufs {
  DevBuf<WNode> node;
  DevBuf<WCand> cand;
  DevBuf<int64_t> chunk0;
  DevBuf<int32_t> hist, besthl, cnt_left;
  DevBuf<uint32_t> cmask, taken, bits;
  DevBuf<double> part, ysum, park;
  DevBuf<unsigned long long> pending;
  DevBuf<int32_t> inv;
  bool attr_set = false;
};
WideBufs *wide_bufs_create() { return new WideBufs(); }
void wide_bufs_destroy(WideBufs *wb) { delete wb; }

static int wide_chunk_rows(et_ctx *ctx, int64_t wide_rows, bool sparse_cls) {
  if (const char *env = getenv("ETGPU_WIDE_CHUNK")) {
    int v = atoi(env);
    v = std::max(256, std::min(WCHUNK_MAX, v));
    return v / 32 * 32;
  }
  // Sparse-resident table: a chunk walks the stored entries of a column inside its ROW RANGE, and the rows of a node
  // deep in the tree are scattered over the whole table -- a 3000-row node in one chunk tests all 10^4 entries of
  // every candidate column in one CTA.  Small chunks cut the range (and that latency chain) into pieces.
  if (sparse_cls) {  // ... as small as keeps the level within a few waves of CTAs (each pays two range searches per candidate)
    int chunk = 256;
    while (chunk < 2048 && wide_rows / chunk > (int64_t)ctx->sm_count * 64) chunk <<= 1;
    return chunk;
  }
  // enough chunks to fill the GPU a few times over, large enough to amortise the per-chunk staging
  int chunk = WCHUNK_MAX;
  while (chunk > 2048 && wide_rows / chunk < (int64_t)ctx->sm_count * 8) chunk >>= 1;
  return chunk;
}

template <int TASK>
void wide_level(et_ctx *ctx, const P &p, int32_t count, int64_t wide_rows, const LevelCfg &lc, WideBufs &wb,
                cudaStream_t st) {
  if (count <= 0) return;
  NvtxRange nv("etgpu.wide_level");
  const int C = p.C, W = p.W;
  const bool coded = lc.coded_big;
  const int chunk = wide_chunk_rows(ctx, wide_rows, p.csc_row != nullptr && TASK == TASK_CLS);
  const int64_t max_chunks = wide_rows / chunk + count;  // sum of ceil(n / chunk) <= this
  if (max_chunks > 0x7fffffff) ET_FAIL(ET_EUNSUPPORTED, "too many row chunks in one level");
  wb.node.ensure((size_t)count);
  wb.cand.ensure((size_t)count);
  wb.chunk0.ensure((size_t)count + 1);
  wb.hist.ensure((size_t)count * 64 * (size_t)C);
  wb.besthl.ensure((size_t)count * (size_t)C);
  wb.cmask.ensure((size_t)count * (size_t)W);
  wb.taken.ensure((size_t)count * (size_t)W);
  wb.cnt_left.ensure((size_t)max_chunks);
  wb.bits.ensure((size_t)max_chunks * (size_t)(chunk / 32));
  if (TASK == TASK_REG) {
    wb.part.ensure((size_t)max_chunks * 192);
    wb.ysum.ensure((size_t)max_chunks * 2);
  }
  wb.pending.ensure(2);
  if (p.csc_row) wb.inv.ensure((size_t)p.inv_rows, 1.0);
  // parked values of pass 1 (FP64 tables): [chunks][stride candidates][chunk rows]; kept within a third of the free HBM
  int park_stride = 0;
  if (!coded && !p.csc_row) {
    static const bool no_park = getenv("ETGPU_NO_PARK") != nullptr && atoi(getenv("ETGPU_NO_PARK")) != 0;
    park_stride = no_park ? 0 : std::min(32, ((p.k + std::max(2, p.k / 4)) + 3) / 4 * 4);
    const size_t need = (size_t)max_chunks * (size_t)park_stride * (size_t)chunk;
    if (park_stride > 0 && need > wb.park.cap) {
      size_t fr = 0, tot = 0;
      cudaMemGetInfo(&fr, &tot);
      if (need * sizeof(double) > (fr + wb.park.cap * sizeof(double)) / 3) park_stride = 0;
    }
    if (park_stride > 0) wb.park.ensure(need, 1.0);
  }
  WState w;
  w.node = wb.node.p;
  w.cand = wb.cand.p;
  w.chunk0 = wb.chunk0.p;
  w.hist = wb.hist.p;
  w.besthl = wb.besthl.p;
  w.cmask = wb.cmask.p;
  w.taken = wb.taken.p;
  w.part = wb.part.p;
  w.ysum = wb.ysum.p;
  w.cnt_left = wb.cnt_left.p;
  w.bits = wb.bits.p;
  w.pending = wb.pending.p;
  w.park = park_stride > 0 ? wb.park.p : nullptr;
  w.park_stride = park_stride;
  w.inv = p.csc_row ? wb.inv.p : nullptr;
  w.chunk = chunk;
  w.count = count;
  {
    static const bool no_bulk = getenv("ETGPU_NO_BULK") != nullptr && atoi(getenv("ETGPU_NO_BULK")) != 0;
    w.bulk = no_bulk ? 0 : 1;
  }
  const unsigned gchunks = (unsigned)max_chunks, gnodes = (unsigned)ceil_div(count, 4);
  const size_t smem1 = (size_t)chunk * 4 + 32;
  const size_t smem2 = (size_t)chunk * 4 + 32 + (TASK == TASK_CLS ? (size_t)32 * ((2 * C) | 1) * 4 + (size_t)chunk : 0);
  if (!wb.attr_set) {
    const int mx1 = WCHUNK_MAX * 4 + 32, mx2 = WCHUNK_MAX * 5 + 32 + 32 * 65 * 4;
    CUDA_CHECK(cudaFuncSetAttribute(k_wide_pass1<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx1));
    CUDA_CHECK(cudaFuncSetAttribute(k_wide_pass1<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx1));
    CUDA_CHECK(cudaFuncSetAttribute(k_wide_pass2<TASK_CLS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx2));
    CUDA_CHECK(cudaFuncSetAttribute(k_wide_pass2<TASK_CLS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx2));
    CUDA_CHECK(cudaFuncSetAttribute(k_wide_pass2<TASK_REG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx2));
    wb.attr_set = true;
  }
  CUDA_CHECK(cudaMemsetAsync(w.pending, 0, 2 * sizeof(unsigned long long), st));
  k_wide_plan<<<1, 1024, 0, st>>>(p, w);
  ctx->launches++;
  if (p.csc_row) {
    k_wide_inv<<<gchunks, WT, 0, st>>>(p, w);
    ctx->launches++;
  }
  if (TASK == TASK_REG) {
    k_wide_regsum<<<gchunks, WT, 0, st>>>(p, w);
    k_wide_regmom<<<gchunks, WT, 0, st>>>(p, w);
    ctx->launches += 2;
  }
  k_wide_first<TASK><<<gnodes, 128, 0, st>>>(p, w);
  ctx->launches++;
  for (int round = 0;; round++) {
    if (round > 0) {  // (the first round always runs: a node without candidates makes its kernels return at once)
      unsigned long long pend = 0;
      CUDA_CHECK(cudaMemcpyAsync(&pend, w.pending, sizeof(pend), cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaMemsetAsync(w.pending, 0, sizeof(unsigned long long), st));
      CUDA_CHECK(cudaStreamSynchronize(st));
      if (pend == 0) break;
    } else {
      CUDA_CHECK(cudaMemsetAsync(w.pending, 0, sizeof(unsigned long long), st));
    }
    if (coded) {
      k_wide_pass1<true><<<gchunks, WT, smem1, st>>>(p, w);
      k_wide_decide<true><<<gnodes, 128, 0, st>>>(p, w);
      if constexpr (TASK == TASK_CLS) k_wide_pass2<TASK_CLS, true><<<gchunks, WT, smem2, st>>>(p, w);
    } else {
      k_wide_pass1<false><<<gchunks, WT, smem1, st>>>(p, w);
      k_wide_decide<false><<<gnodes, 128, 0, st>>>(p, w);
      k_wide_pass2<TASK, false><<<gchunks, WT, smem2, st>>>(p, w);
    }
    k_wide_score<TASK><<<gnodes, 128, 0, st>>>(p, w);
    ctx->launches += 4;
  }
  k_wide_finish<TASK><<<gnodes, 128, 0, st>>>(p, w);
  if (coded)
    k_wide_count<true><<<gchunks, WT, 0, st>>>(p, w);
  else
    k_wide_count<false><<<gchunks, WT, p.csc_row ? smem1 : 0, st>>>(p, w);
  k_wide_scatter<TASK><<<gchunks, WT, 0, st>>>(p, w);
  ctx->launches += 3;
  CUDA_CHECK(cudaGetLastError());
}

template void wide_level<TASK_CLS>(et_ctx *, const P &, int32_t, int64_t, const LevelCfg &, WideBufs &, cudaStream_t);
template void wide_level<TASK_REG>(et_ctx *, const P &, int32_t, int64_t, const LevelCfg &, WideBufs &, cudaStream_t);

}  // namespace etb
