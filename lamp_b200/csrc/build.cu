// build.cu -- level-wise extratrees builder over all open nodes of all trees of a batch.
//
// Replaces the recursive JVM builder buildTreeClassification (pkg:943-1082) / buildTreeRegression
// (pkg:766-895) and the split search splitClassification (pkg:203-297) / splitRegression
// (pkg:427-511) of extratrees/src/main/scala/lamp/forest/package.scala.
//
// Data layout in HBM
//   X        column-major FP64 [d][ld]              (the JVM walks a row-major matrix with stride d)
//   idx      int32 [B][n] x2 (ping-pong)            sample rows of every open node, ascending inside a
//                                                   node segment (the reference's filter keeps order)
//   yc/yr/w  labels / targets / weights permuted alongside idx so a node's segment streams
//   frontier SoA of open nodes (tree, begin, end, node id, depth, RNG key | trace node, class hist)
//   cand     SoA of candidate (node, feature) pairs of the current round
//
// One level = classify (stop rules) -> rounds of { draw candidates ; min/max + threshold score } ->
// finalize (first-best argmax already folded into the rounds) -> stable partition.
//
// Exactness: every floating-point expression of the reference is evaluated with individually
// rounded _rn operations in the reference's order.  Unweighted classification reduces integer
// class histograms in parallel (exact) and evaluates the Gini expressions in one thread; weighted
// classification and regression sum in subset order (sequential chains, one thread per candidate,
// massively parallel across candidates/nodes/trees) because FP addition is not associative.
#include <algorithm>
#include <chrono>

#include "internal.h"

namespace {

enum { TASK_CLS = 0, TASK_CLSW = 1, TASK_REG = 2 };
enum { FLAG_SEARCH = 0, FLAG_LEAF = 1, FLAG_DONE = 2 };
enum { CF_CONST = 1, CF_NAN = 2 };

enum {
  ST_VMM = 0,
  ST_VSC,
  ST_SROWS,
  ST_PROWS,
  ST_DRAWS,
  ST_CONST,
  ST_SCORED,
  ST_MISMATCH,
  ST_COUNT
};

struct Counters {
  int32_t n_cand[2];
  int32_t next_f;
  int32_t pad;
  unsigned long long mask_words;
  unsigned long long st[ST_COUNT];
};

struct Level {  // frontier of one level
  int32_t *tree, *begin, *end, *node, *depth;
  int64_t *trace;
  uint64_t *key;
  int32_t *hist;   // [F][C]   (TASK_CLS)
  uint32_t *mask;  // [F][W]   (free-running: known-constant | taken features)
};

struct Search {  // per frontier node, valid within one level
  uint8_t *flag, *best_mil;
  int32_t *visited, *nconst, *dc, *cand_begin, *cand_cnt, *best_feature, *best_nleft, *split_slot;
  double *best_score, *best_cut, *total, *nsum, *mean;
  int32_t *scored;   // [F][k] features scored at this node (their mask bits are not inherited)
  int32_t *best_hl;  // [F][C] class histogram of the best candidate's left side (TASK_CLS)
  double *dist;      // [F][C] weighted class distribution (TASK_CLSW)
};

struct Cand {  // per candidate of one round
  int32_t *node, *feature, *cnt_lt, *cnt_nan;
  double *u, *cut, *score;
  uint8_t *flags, *mil;
  int32_t *hist;      // [cand][2][C]  (TASK_CLS): <cut histogram, NaN histogram
  int64_t *mask_off;  // word offset of the side bitmasks (TASK_CLSW / TASK_REG)
};

struct Out {
  int32_t *feature, *left, *right;
  double *cut, *leaf;
  uint8_t *mil;
};

struct Trace {
  const int64_t *cand_begin;
  const int32_t *cand_count, *left, *right, *cand_feature;
  const double *cand_u;
  const uint8_t *cand_flag;
};

struct P {  // kernel parameters shared by all kernels of a batch
  const double *X;
  int64_t ld, n, n_table;
  int32_t d, C, k, n_min, max_depth, W, task, replay;
  int32_t *idx_src, *idx_dst, *yc_src, *yc_dst;
  double *yr_src, *yr_dst, *w_src, *w_dst;
  Level cur, nxt;
  Search s;
  Cand c[2];
  Out o;
  Trace tr;
  Counters *cnt;
  uint32_t *sidemask;  // bitmask scratch (TASK_CLSW / TASK_REG)
  double *wscratch;    // [cand][2][2][C] weighted histograms
  int32_t node_base_next;
};

__device__ __forceinline__ void stat_add(const P &p, int which, unsigned long long v) {
  atomicAdd(&p.cnt->st[which], v);
}

// ---- roots ----------------------------------------------------------------------------------
__global__ void k_init_samples(int64_t n, int32_t B, int32_t *idx, const int32_t *y_cls, int32_t *yc,
                               const double *y_reg, double *yr, const double *w, double *ws) {
  int64_t total = n * B;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = g % n;
    idx[g] = (int32_t)i;
    if (yc) yc[g] = y_cls[i];
    if (yr) yr[g] = y_reg[i];
    if (ws) ws[g] = w[i];
  }
}

__global__ void k_init_roots(P p, int32_t B, const uint64_t *tree_keys, const int64_t *trace_roots,
                             const int32_t *root_hist) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B) return;
  p.cur.tree[t] = t;
  p.cur.begin[t] = 0;
  p.cur.end[t] = (int32_t)p.n;
  p.cur.node[t] = t;
  p.cur.depth[t] = 0;
  p.cur.trace[t] = trace_roots ? trace_roots[t] : -1;
  p.cur.key[t] = tree_keys[t];
  if (p.task == TASK_CLS)
    for (int c = 0; c < p.C; c++) p.cur.hist[(int64_t)t * p.C + c] = root_hist[c];
  if (!p.replay) {
    for (int w = 0; w < p.W; w++) {
      uint32_t m = 0;
      int lo = w * 32;
      if (lo + 32 > p.d) m = (p.d - lo >= 32) ? 0u : (p.d <= lo ? 0xffffffffu : (0xffffffffu << (p.d - lo)));
      p.cur.mask[(int64_t)t * p.W + w] = m;
    }
  }
}

// ---- classify: the reference's stop rules + node totals ---------------------------------------
__device__ __forceinline__ void search_init(const P &p, int i) {
  p.s.flag[i] = FLAG_SEARCH;
  p.s.visited[i] = 0;
  p.s.dc[i] = 0;
  p.s.cand_begin[i] = 0;
  p.s.cand_cnt[i] = 0;
  p.s.best_feature[i] = -1;
  p.s.best_nleft[i] = 0;
  p.s.best_mil[i] = 0;
  p.s.best_score[i] = -INFINITY;
  p.s.best_cut[i] = NAN;
  p.s.split_slot[i] = -1;
}

// TASK_CLS: stop rules of pkg:993-994 from the node's integer class histogram; Gini total with the
// reference's repeated `+= 1/s` distribution (pkg:905-911, 1160-1180).
__global__ void k_classify_cls(P p, int32_t F) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= F) return;
  const int32_t *h = p.cur.hist + (int64_t)i * p.C;
  int32_t n = p.cur.end[i] - p.cur.begin[i];
  bool pure = false;
  for (int c = 0; c < p.C; c++) pure |= (h[c] == n);
  p.s.split_slot[i] = -1;
  p.s.cand_cnt[i] = 0;
  p.s.best_feature[i] = -1;
  if (p.n_table < p.n_min || p.cur.depth[i] >= p.max_depth || pure) {
    p.s.flag[i] = FLAG_LEAF;
    return;
  }
  search_init(p, i);
  p.s.nconst[i] = 0;
  if (!p.replay) {
    int nc = 0;
    const uint32_t *m = p.cur.mask + (int64_t)i * p.W;
    for (int w = 0; w < p.W; w++) nc += __popc(m[w]);
    p.s.nconst[i] = nc - (p.W * 32 - p.d);
  }
  double inv = ET_DIV(1.0, (double)n);
  double s = 0.0;
  for (int c = 0; c < p.C; c++) {
    double pc = et_repeat_add(inv, h[c]);
    s = ET_ADD(s, ET_MUL(pc, pc));
  }
  p.s.total[i] = ET_SUB(1.0, s);
  p.s.nsum[i] = (double)n;
  stat_add(p, ST_SROWS, (unsigned long long)n);
}

// TASK_REG: pkg:799-814 (targetIsConstant with !=), leaf mean (mean2, pkg:782), varianceNoSplit
// (pkg:436-437) -- all sequential in subset order.
__global__ void k_classify_reg(P p, int32_t F) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= F) return;
  int32_t b = p.cur.begin[i], e = p.cur.end[i];
  int32_t n = e - b;
  const double *y = p.yr_src + (int64_t)p.cur.tree[i] * p.n;
  double head = y[b];
  bool uniform = true;
  double sum = 0.0;
  for (int32_t j = b; j < e; j++) {
    double v = y[j];
    sum = ET_ADD(sum, v);
    uniform &= !(v != head);
  }
  double dn = (double)n;
  double mean = ET_DIV(sum, dn);
  p.s.mean[i] = mean;
  p.s.split_slot[i] = -1;
  p.s.cand_cnt[i] = 0;
  p.s.best_feature[i] = -1;
  if (n < p.n_min || p.cur.depth[i] >= p.max_depth || uniform) {
    p.s.flag[i] = FLAG_LEAF;
    return;
  }
  search_init(p, i);
  p.s.nconst[i] = 0;
  if (!p.replay) {
    int nc = 0;
    const uint32_t *m = p.cur.mask + (int64_t)i * p.W;
    for (int w = 0; w < p.W; w++) nc += __popc(m[w]);
    p.s.nconst[i] = nc - (p.W * 32 - p.d);
  }
  // sampleVariance: two-pass; n == 1 -> 0
  double var;
  if (n == 1) {
    var = 0.0;
  } else {
    double q = 0.0;
    for (int32_t j = b; j < e; j++) {
      double dl = ET_SUB(y[j], mean);
      q = ET_ADD(q, ET_MUL(dl, dl));
    }
    var = ET_DIV(q, ET_SUB(dn, 1.0));
  }
  p.s.total[i] = ET_DIV(ET_MUL(var, ET_SUB(dn, 1.0)), dn);
  p.s.nsum[i] = dn;
  stat_add(p, ST_SROWS, (unsigned long long)n);
}

// TASK_CLSW: weighted distribution (pkg:913-927) summed in subset order; also the leaf value.
__global__ void k_classify_clsw(P p, int32_t F) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= F) return;
  int32_t b = p.cur.begin[i], e = p.cur.end[i];
  int32_t n = e - b;
  int64_t base = (int64_t)p.cur.tree[i] * p.n;
  const int32_t *y = p.yc_src + base;
  const double *w = p.w_src + base;
  double *dist = p.s.dist + (int64_t)i * p.C;
  for (int c = 0; c < p.C; c++) dist[c] = 0.0;
  double s = 0.0;
  int32_t head = y[b];
  bool uniform = true;
  for (int32_t j = b; j < e; j++) {
    int32_t cls = y[j];
    double ww = w[j];
    dist[cls] = ET_ADD(dist[cls], ww);
    s = ET_ADD(s, ww);
    uniform &= (cls == head);
  }
  double sq = 0.0;
  for (int c = 0; c < p.C; c++) {
    double pc = ET_DIV(dist[c], s);
    dist[c] = pc;
    sq = ET_ADD(sq, ET_MUL(pc, pc));
  }
  p.s.split_slot[i] = -1;
  p.s.cand_cnt[i] = 0;
  p.s.best_feature[i] = -1;
  if (p.n_table < p.n_min || p.cur.depth[i] >= p.max_depth || uniform) {
    p.s.flag[i] = FLAG_LEAF;
    return;
  }
  search_init(p, i);
  p.s.nconst[i] = 0;
  if (!p.replay) {
    int nc = 0;
    const uint32_t *m = p.cur.mask + (int64_t)i * p.W;
    for (int w2 = 0; w2 < p.W; w2++) nc += __popc(m[w2]);
    p.s.nconst[i] = nc - (p.W * 32 - p.d);
  }
  p.s.total[i] = ET_SUB(1.0, sq);
  p.s.nsum[i] = s;  // sampleWeights.sum2 over the subset (pkg:1112): same order, same value
  stat_add(p, ST_SROWS, (unsigned long long)n);
}

// ---- Gini score of one candidate from integer histograms (pkg:1101-1158, unweighted) ---------
// hin[c] = hl[c] (+ hn[c] when NaN rows go left); hout = node hist - hin.
__device__ double gini_score_int(const int32_t *hnode, const int32_t *hl, const int32_t *hn, bool nan_left, int C,
                                 int32_t n, double G) {
  int32_t cin_i = 0;
  for (int c = 0; c < C; c++) cin_i += hl[c] + (nan_left ? hn[c] : 0);
  double cin = (double)cin_i, cout = (double)(n - cin_i), N = (double)n;
  double sin_ = 0.0, sout = 0.0;
  for (int c = 0; c < C; c++) {
    int32_t hi = hl[c] + (nan_left ? hn[c] : 0);
    double pi = ET_DIV((double)hi, cin);
    double po = ET_DIV((double)(hnode[c] - hi), cout);
    sin_ = ET_ADD(sin_, ET_MUL(pi, pi));
    sout = ET_ADD(sout, ET_MUL(po, po));
  }
  double gin = ET_SUB(1.0, sin_), gout = ET_SUB(1.0, sout);
  return ET_SUB(ET_SUB(G, ET_DIV(ET_MUL(gin, cin), N)), ET_DIV(ET_MUL(gout, cout), N));
}

// ---- free-running feature draw: uniform over features that are neither known-constant nor taken
__device__ int32_t pick_feature(uint32_t *mask, int32_t d, int32_t W, int32_t avail, uint64_t key, int32_t &dc) {
  for (int t = 0; t < 6; t++) {
    uint64_t r = et_draw(key, (uint32_t)dc++);
    int32_t f = (int32_t)__umul64hi(r, (uint64_t)d);
    uint32_t bit = 1u << (f & 31);
    if (!(mask[f >> 5] & bit)) {
      mask[f >> 5] |= bit;
      return f;
    }
  }
  uint64_t r = et_draw(key, (uint32_t)dc++);
  int32_t rank = (int32_t)__umul64hi(r, (uint64_t)avail);
  for (int w = 0; w < W; w++) {
    uint32_t z = ~mask[w];
    int c = __popc(z);
    if (rank < c) {
      for (int q = 0; q < rank; q++) z &= z - 1;  // drop the lowest set bits
      int pos = __ffs(z) - 1;
      mask[w] |= 1u << pos;
      return w * 32 + pos;
    }
    rank -= c;
  }
  return -1;  // unreachable when avail is consistent
}

// ---- round kernel: consume the previous round's results, then draw what is still needed -------
// One thread per frontier node.  Implements the loop of pkg:232-292 / 453-505: candidates are
// consumed in draw order, constants and NaN-scoring features do not count toward k, strict `>`
// keeps the first best.
__global__ void k_update_draw(P p, int32_t F, int round) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= F) return;
  if (p.s.flag[i] != FLAG_SEARCH) return;
  const int prev = (round & 1) ^ 1, cur = round & 1;
  const Cand &cp = p.c[prev];
  const Cand &cc = p.c[cur];
  const int32_t n = p.cur.end[i] - p.cur.begin[i];
  int32_t visited = p.s.visited[i], nconst = p.s.nconst[i];
  const int C = p.C;
  // consume
  int32_t cb = p.s.cand_begin[i], ccnt = p.s.cand_cnt[i];
  if (ccnt > 0) {
    double best = p.s.best_score[i];
    const double G = p.s.total[i];
    unsigned long long n_const = 0, n_scored = 0;
    for (int j = 0; j < ccnt; j++) {
      int c = cb + j;
      uint8_t fl = cp.flags[c];
      int32_t f = cp.feature[c];
      if (fl & CF_CONST) {
        nconst++;
        n_const++;
        if (p.replay && ((fl >> 4) & 3) != 1) stat_add(p, ST_MISMATCH, 1);
        continue;
      }
      n_scored++;
      double s;
      bool mil = false;
      if (p.task == TASK_CLS) {
        const int32_t *hl = cp.hist + (int64_t)c * 2 * C;
        const int32_t *hn = hl + C;
        const int32_t *hnode = p.cur.hist + (int64_t)i * C;
        double sn = gini_score_int(hnode, hl, hn, false, C, n, G);
        s = sn;
        if (fl & CF_NAN) {
          double sl = gini_score_int(hnode, hl, hn, true, C, n, G);
          mil = !(sl != sl) && (sl > sn || (sn != sn));
          if (mil) s = sl;
        }
      } else {
        s = cp.score[c];
        mil = cp.mil[c] != 0;
      }
      if (s > best) {
        best = s;
        p.s.best_feature[i] = f;
        p.s.best_cut[i] = cp.cut[c];
        p.s.best_mil[i] = mil ? 1 : 0;
        p.s.best_nleft[i] = cp.cnt_lt[c] + (mil ? cp.cnt_nan[c] : 0);
        if (p.task == TASK_CLS) {
          const int32_t *hl = cp.hist + (int64_t)c * 2 * C;
          int32_t *bh = p.s.best_hl + (int64_t)i * C;
          for (int q = 0; q < C; q++) bh[q] = hl[q] + (mil ? hl[C + q] : 0);
        }
      }
      if (s != s) {
        nconst++;  // pkg:283-285: joins the inherited "constant" prefix
        if (p.replay && ((fl >> 4) & 3) != 3) stat_add(p, ST_MISMATCH, 1);
      } else {
        if (!p.replay) p.s.scored[(int64_t)i * p.k + visited] = f;
        visited++;
        if (p.replay && ((fl >> 4) & 3) != 2) stat_add(p, ST_MISMATCH, 1);
      }
    }
    p.s.best_score[i] = best;
    stat_add(p, ST_VMM, (unsigned long long)n * (unsigned long long)ccnt);
    stat_add(p, ST_VSC, (unsigned long long)n * n_scored);
    stat_add(p, ST_DRAWS, (unsigned long long)ccnt);
    stat_add(p, ST_CONST, n_const);
    stat_add(p, ST_SCORED, n_scored);
  }
  // decide
  int32_t need;
  int64_t tn = p.cur.trace[i];
  if (p.replay) {
    need = (round == 0 && tn >= 0) ? p.tr.cand_count[tn] : 0;
  } else {
    int32_t avail = p.d - nconst - visited;
    need = min(p.k - visited, avail);
  }
  p.s.visited[i] = visited;
  p.s.nconst[i] = nconst;
  if (need <= 0) {
    p.s.flag[i] = FLAG_DONE;
    p.s.cand_cnt[i] = 0;
    return;
  }
  int32_t base = atomicAdd(&p.cnt->n_cand[cur], need);
  p.s.cand_begin[i] = base;
  p.s.cand_cnt[i] = need;
  if (p.task != TASK_CLS) {
    // side bitmasks: [<cut words][NaN words], n rounded up to 32 each
    unsigned long long words = 2ull * (unsigned long long)((n + 31) / 32);
    unsigned long long off = atomicAdd(&p.cnt->mask_words, words * (unsigned long long)need);
    for (int j = 0; j < need; j++) cc.mask_off[base + j] = (int64_t)(off + words * j);
  }
  if (p.replay) {
    int64_t tb = p.tr.cand_begin[tn];
    for (int j = 0; j < need; j++) {
      cc.node[base + j] = i;
      cc.feature[base + j] = p.tr.cand_feature[tb + j];
      cc.u[base + j] = p.tr.cand_u[tb + j];
      cc.flags[base + j] = (uint8_t)((p.tr.cand_flag[tb + j] + 1) << 4);
    }
  } else {
    uint32_t *mask = p.cur.mask + (int64_t)i * p.W;
    const uint64_t key = p.cur.key[i];
    int32_t dc = p.s.dc[i];
    int32_t avail = p.d - nconst - visited;
    for (int j = 0; j < need; j++) {
      int32_t f = pick_feature(mask, p.d, p.W, avail - j, key, dc);
      cc.node[base + j] = i;
      cc.feature[base + j] = f;
      cc.u[base + j] = et_u01(et_draw(key, (uint32_t)dc++));
      cc.flags[base + j] = 0;
    }
    p.s.dc[i] = dc;
  }
}

// ---- item kernel: one warp per candidate (node, feature) --------------------------------------
// pass 1: min / max / hasMissing over the node's samples (pkg:34-54; NaN ignored by < and >).
// constant test `max <= min && !hasMissing` (pkg:236); cut = min + (max - min) * u (pkg:240).
// pass 2 (same warp, the column segment is still in L1/L2): <cut side.
//   TASK_CLS : integer class histograms of the <cut rows and of the NaN rows
//   otherwise: side bitmasks for the sequential scorer
template <int TASK>
__global__ void __launch_bounds__(256) k_items_warp(P p, int32_t n_cand, int par) {
  extern __shared__ int32_t sm_hist[];  // [warps][2][C]
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int c = blockIdx.x * (blockDim.x >> 5) + wib;
  if (c >= n_cand) return;
  const Cand &cd = p.c[par];
  const int i = cd.node[c];
  const int32_t f = cd.feature[c];
  const int32_t b = p.cur.begin[i], e = p.cur.end[i];
  const int64_t base = (int64_t)p.cur.tree[i] * p.n;
  const int32_t *idx = p.idx_src + base;
  const double *col = p.X + (int64_t)f * p.ld;
  double mn = 1.7976931348623157e308, mx = -1.7976931348623157e308;
  int has_nan = 0;
  for (int32_t j = b + lane; j < e; j += 32) {
    double x = __ldg(col + idx[j]);
    if (x < mn) mn = x;
    if (x > mx) mx = x;
    has_nan |= (x != x);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double omn = __shfl_xor_sync(0xffffffffu, mn, o);
    double omx = __shfl_xor_sync(0xffffffffu, mx, o);
    if (omn < mn) mn = omn;
    if (omx > mx) mx = omx;
  }
  has_nan = __any_sync(0xffffffffu, has_nan);
  uint8_t fl = cd.flags[c] & 0xf0;
  if (mx <= mn && !has_nan) {
    if (lane == 0) cd.flags[c] = fl | CF_CONST;
    return;
  }
  const double cut = ET_ADD(mn, ET_MUL(ET_SUB(mx, mn), cd.u[c]));
  int32_t cnt_lt = 0, cnt_nan = 0;
  if (TASK == TASK_CLS) {
    const int C = p.C;
    int32_t *hl = sm_hist + wib * 2 * C, *hn = hl + C;
    for (int q = lane; q < 2 * C; q += 32) hl[q] = 0;
    __syncwarp();
    const int32_t *y = p.yc_src + base;
    for (int32_t j = b + lane; j < e; j += 32) {
      double x = __ldg(col + idx[j]);
      int32_t cls = y[j];
      if (x < cut) {
        atomicAdd(&hl[cls], 1);
        cnt_lt++;
      } else if (x != x) {
        atomicAdd(&hn[cls], 1);
        cnt_nan++;
      }
    }
    __syncwarp();
    int32_t *gh = cd.hist + (int64_t)c * 2 * C;
    for (int q = lane; q < 2 * C; q += 32) gh[q] = hl[q];
  } else {
    uint32_t *mlt = p.sidemask + cd.mask_off[c];
    uint32_t *mnan = mlt + (e - b + 31) / 32;
    for (int32_t j0 = b; j0 < e; j0 += 32) {
      int32_t j = j0 + lane;
      bool lt = false, isn = false;
      if (j < e) {
        double x = __ldg(col + idx[j]);
        lt = x < cut;
        isn = x != x;
      }
      uint32_t blt = __ballot_sync(0xffffffffu, lt), bnan = __ballot_sync(0xffffffffu, isn);
      if (lane == 0) {
        mlt[(j0 - b) >> 5] = blt;
        mnan[(j0 - b) >> 5] = bnan;
      }
      cnt_lt += lt;
      cnt_nan += isn;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt_lt += __shfl_xor_sync(0xffffffffu, cnt_lt, o);
    cnt_nan += __shfl_xor_sync(0xffffffffu, cnt_nan, o);
  }
  if (lane == 0) {
    cd.flags[c] = fl | (has_nan ? CF_NAN : 0);
    cd.cut[c] = cut;
    cd.cnt_lt[c] = cnt_lt;
    cd.cnt_nan[c] = cnt_nan;
  }
}

// ---- sequential scorers (one thread per candidate) --------------------------------------------
// computeVarianceReduction (pkg:1196-1218) with saddle's two-pass sampleVariance, in subset order.
__device__ double var_reduction_seq(const double *y, int32_t n, const uint32_t *mlt, const uint32_t *mnan,
                                    bool nan_left, double V) {
  double sin_ = 0.0, sout = 0.0;
  int32_t nin = 0;
  for (int32_t j = 0; j < n; j++) {
    uint32_t w = mlt[j >> 5];
    if (nan_left) w |= mnan[j >> 5];
    double v = y[j];
    if ((w >> (j & 31)) & 1u) {
      sin_ = ET_ADD(sin_, v);
      nin++;
    } else {
      sout = ET_ADD(sout, v);
    }
  }
  int32_t nout = n - nin;
  double dnin = (double)nin, dnout = (double)nout, dn = (double)n;
  double vin, vout;
  {
    double min_ = ET_DIV(sin_, dnin), mout = ET_DIV(sout, dnout);
    double qin = 0.0, qout = 0.0;
    for (int32_t j = 0; j < n; j++) {
      uint32_t w = mlt[j >> 5];
      if (nan_left) w |= mnan[j >> 5];
      double v = y[j];
      if ((w >> (j & 31)) & 1u) {
        double dl = ET_SUB(v, min_);
        qin = ET_ADD(qin, ET_MUL(dl, dl));
      } else {
        double dl = ET_SUB(v, mout);
        qout = ET_ADD(qout, ET_MUL(dl, dl));
      }
    }
    // sampleVariance: n < 1 -> NaN, n == 1 -> 0 (pkg:1204 short-circuits n == 1 as well)
    double svin = nin < 1 ? NAN : (nin == 1 ? 0.0 : ET_DIV(qin, ET_SUB(dnin, 1.0)));
    double svout = nout < 1 ? NAN : (nout == 1 ? 0.0 : ET_DIV(qout, ET_SUB(dnout, 1.0)));
    vin = (nin == 1) ? 0.0 : ET_DIV(ET_MUL(svin, ET_SUB(dnin, 1.0)), dnin);
    vout = (nout == 1) ? 0.0 : ET_DIV(ET_MUL(svout, ET_SUB(dnout, 1.0)), dnout);
  }
  double a = ET_MUL(ET_DIV(dnin, dn), vin);
  double bq = ET_MUL(ET_DIV(dnout, dn), vout);
  return ET_DIV(ET_SUB(ET_SUB(V, a), bq), V);
}

__global__ void k_score_reg(P p, int32_t n_cand, int par) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cand) return;
  const Cand &cd = p.c[par];
  uint8_t fl = cd.flags[c];
  if (fl & CF_CONST) return;
  int i = cd.node[c];
  int32_t b = p.cur.begin[i], n = p.cur.end[i] - b;
  const double *y = p.yr_src + (int64_t)p.cur.tree[i] * p.n + b;
  const uint32_t *mlt = p.sidemask + cd.mask_off[c];
  const uint32_t *mnan = mlt + (n + 31) / 32;
  double V = p.s.total[i];
  double sn = var_reduction_seq(y, n, mlt, mnan, false, V);
  double s = sn;
  bool mil = false;
  if (fl & CF_NAN) {
    double sl = var_reduction_seq(y, n, mlt, mnan, true, V);
    mil = !(sl != sl) && (sl > sn || (sn != sn));
    if (mil) s = sl;
  }
  cd.score[c] = s;
  cd.mil[c] = mil ? 1 : 0;
}

// weighted giniScore (pkg:1132-1157): sequential weighted class sums in subset order.
__device__ double gini_score_w_seq(const int32_t *y, const double *w, int32_t n, const uint32_t *mlt,
                                   const uint32_t *mnan, bool nan_left, int C, double G, double N, double *hin,
                                   double *hout) {
  for (int q = 0; q < C; q++) {
    hin[q] = 0.0;
    hout[q] = 0.0;
  }
  double cin = 0.0, cout = 0.0;
  for (int32_t j = 0; j < n; j++) {
    uint32_t m = mlt[j >> 5];
    if (nan_left) m |= mnan[j >> 5];
    double ww = w[j];
    int32_t cls = y[j];
    if ((m >> (j & 31)) & 1u) {
      cin = ET_ADD(cin, ww);
      hin[cls] = ET_ADD(hin[cls], ww);
    } else {
      cout = ET_ADD(cout, ww);
      hout[cls] = ET_ADD(hout[cls], ww);
    }
  }
  double sin_ = 0.0, sout = 0.0;
  for (int q = 0; q < C; q++) {
    double pi = ET_DIV(hin[q], cin), po = ET_DIV(hout[q], cout);
    sin_ = ET_ADD(sin_, ET_MUL(pi, pi));
    sout = ET_ADD(sout, ET_MUL(po, po));
  }
  double gin = ET_SUB(1.0, sin_), gout = ET_SUB(1.0, sout);
  return ET_SUB(ET_SUB(G, ET_DIV(ET_MUL(gin, cin), N)), ET_DIV(ET_MUL(gout, cout), N));
}

__global__ void k_score_clsw(P p, int32_t n_cand, int par) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cand) return;
  const Cand &cd = p.c[par];
  uint8_t fl = cd.flags[c];
  if (fl & CF_CONST) return;
  int i = cd.node[c];
  int32_t b = p.cur.begin[i], n = p.cur.end[i] - b;
  int64_t base = (int64_t)p.cur.tree[i] * p.n + b;
  const int32_t *y = p.yc_src + base;
  const double *w = p.w_src + base;
  const uint32_t *mlt = p.sidemask + cd.mask_off[c];
  const uint32_t *mnan = mlt + (n + 31) / 32;
  double *hin = p.wscratch + (int64_t)c * 2 * p.C, *hout = hin + p.C;
  double G = p.s.total[i], N = p.s.nsum[i];
  double sn = gini_score_w_seq(y, w, n, mlt, mnan, false, p.C, G, N, hin, hout);
  double s = sn;
  bool mil = false;
  if (fl & CF_NAN) {
    double sl = gini_score_w_seq(y, w, n, mlt, mnan, true, p.C, G, N, hin, hout);
    mil = !(sl != sl) && (sl > sn || (sn != sn));
    if (mil) s = sl;
  }
  cd.score[c] = s;
  cd.mil[c] = mil ? 1 : 0;
}

// ---- finalize: write the output node, open the children ---------------------------------------
__global__ void k_finalize(P p, int32_t F) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= F) return;
  const int32_t node = p.cur.node[i];
  const int C = p.C;
  const int lw = (p.task == TASK_REG) ? 1 : C;
  const int32_t b = p.cur.begin[i], e = p.cur.end[i], n = e - b;
  const int32_t bf = p.s.best_feature[i];
  const bool leaf = (p.s.flag[i] == FLAG_LEAF) || bf < 0;  // pkg:293-296 (visited == 0 || cut NaN) <=> bf < 0
  const int64_t tn = p.cur.trace[i];
  if (leaf) {
    p.o.feature[node] = -1;
    p.o.left[node] = -1;
    p.o.right[node] = -1;
    p.o.cut[node] = NAN;
    p.o.mil[node] = 0;
    double *lv = p.o.leaf + (int64_t)node * lw;
    if (p.task == TASK_CLS) {
      const int32_t *h = p.cur.hist + (int64_t)i * C;
      double inv = ET_DIV(1.0, (double)n);
      for (int c = 0; c < C; c++) lv[c] = et_repeat_add(inv, h[c]);
    } else if (p.task == TASK_CLSW) {
      const double *ds = p.s.dist + (int64_t)i * C;
      for (int c = 0; c < C; c++) lv[c] = ds[c];
    } else {
      lv[0] = p.s.mean[i];
    }
    if (p.replay && tn >= 0 && p.tr.left[tn] >= 0) stat_add(p, ST_MISMATCH, 1);
    return;
  }
  const int32_t nl = p.s.best_nleft[i];
  const int32_t slot = atomicAdd(&p.cnt->next_f, 2);
  p.s.split_slot[i] = slot;
  const int32_t cl = p.node_base_next + slot, cr = cl + 1;
  p.o.feature[node] = bf;
  p.o.cut[node] = p.s.best_cut[i];
  p.o.mil[node] = p.s.best_mil[i];
  p.o.left[node] = cl;
  p.o.right[node] = cr;
  double *lv = p.o.leaf + (int64_t)node * lw;
  for (int c = 0; c < lw; c++) lv[c] = 0.0;
  const int32_t t = p.cur.tree[i], dep = p.cur.depth[i];
  p.nxt.tree[slot] = t;
  p.nxt.tree[slot + 1] = t;
  p.nxt.begin[slot] = b;
  p.nxt.end[slot] = b + nl;
  p.nxt.begin[slot + 1] = b + nl;
  p.nxt.end[slot + 1] = e;
  p.nxt.node[slot] = cl;
  p.nxt.node[slot + 1] = cr;
  p.nxt.depth[slot] = dep + 1;
  p.nxt.depth[slot + 1] = (p.task == TASK_REG) ? dep : dep + 1;  // pkg:884 (sic) vs pkg:1071
  const uint64_t key = p.cur.key[i];
  p.nxt.key[slot] = et_child_key(key, 0);
  p.nxt.key[slot + 1] = et_child_key(key, 1);
  int64_t tl = -1, trr = -1;
  if (p.replay) {
    if (tn >= 0 && p.tr.left[tn] >= 0) {
      // trace child ids are tree-local pre-order ids; tn - (its own local id) is not stored, so
      // the host rewrote left/right to absolute node indices before upload
      tl = p.tr.left[tn];
      trr = p.tr.right[tn];
    } else {
      stat_add(p, ST_MISMATCH, 1);
    }
  }
  p.nxt.trace[slot] = tl;
  p.nxt.trace[slot + 1] = trr;
  if (p.task == TASK_CLS) {
    const int32_t *h = p.cur.hist + (int64_t)i * C;
    const int32_t *bh = p.s.best_hl + (int64_t)i * C;
    int32_t *hl = p.nxt.hist + (int64_t)slot * C, *hr = hl + C;
    for (int c = 0; c < C; c++) {
      hl[c] = bh[c];
      hr[c] = h[c] - bh[c];
    }
  }
  if (!p.replay) {
    // children inherit the known-constant set; features merely scored here are released
    uint32_t *m = p.cur.mask + (int64_t)i * p.W;
    const int32_t *sc = p.s.scored + (int64_t)i * p.k;
    const int32_t vis = p.s.visited[i];
    for (int q = 0; q < vis; q++) m[sc[q] >> 5] &= ~(1u << (sc[q] & 31));
    uint32_t *ml = p.nxt.mask + (int64_t)slot * p.W, *mr = ml + p.W;
    for (int w = 0; w < p.W; w++) {
      uint32_t v = m[w];
      ml[w] = v;
      mr[w] = v;
    }
  }
  stat_add(p, ST_PROWS, (unsigned long long)n);
}

// ---- stable partition (pkg:1024-1039 / 841-856): one warp per split node ----------------------
__global__ void __launch_bounds__(256) k_partition_warp(P p, int32_t F) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= F) return;
  if (p.s.split_slot[i] < 0) return;
  const int32_t b = p.cur.begin[i], e = p.cur.end[i];
  const int64_t base = (int64_t)p.cur.tree[i] * p.n;
  const double *col = p.X + (int64_t)p.s.best_feature[i] * p.ld;
  const double cut = p.s.best_cut[i];
  const bool mil = p.s.best_mil[i] != 0;
  int32_t lpos = b, rpos = b + p.s.best_nleft[i];
  const uint32_t lt_mask = (1u << lane) - 1u;
  for (int32_t j0 = b; j0 < e; j0 += 32) {
    const int32_t j = j0 + lane;
    const bool valid = j < e;
    int32_t r = 0;
    bool left = false;
    if (valid) {
      r = p.idx_src[base + j];
      double x = __ldg(col + r);
      left = (x < cut) || (mil && (x != x));
    }
    const uint32_t bv = __ballot_sync(0xffffffffu, valid);
    const uint32_t bl = __ballot_sync(0xffffffffu, left);
    const uint32_t br = bv & ~bl;
    if (valid) {
      const int32_t dst = left ? lpos + __popc(bl & lt_mask) : rpos + __popc(br & lt_mask);
      p.idx_dst[base + dst] = r;
      if (p.task == TASK_REG) {
        p.yr_dst[base + dst] = p.yr_src[base + j];
      } else {
        p.yc_dst[base + dst] = p.yc_src[base + j];
        if (p.task == TASK_CLSW) p.w_dst[base + dst] = p.w_src[base + j];
      }
    }
    lpos += __popc(bl);
    rpos += __popc(br);
  }
}

// ---- host orchestration -----------------------------------------------------------------------
struct LevelBufs {
  DevBuf<int32_t> tree, begin, end, node, depth, hist;
  DevBuf<int64_t> trace;
  DevBuf<uint64_t> key;
  DevBuf<uint32_t> mask;
  void ensure(size_t F, int C, int W, bool need_hist, bool need_mask) {
    tree.ensure(F);
    begin.ensure(F);
    end.ensure(F);
    node.ensure(F);
    depth.ensure(F);
    trace.ensure(F);
    key.ensure(F);
    if (need_hist) hist.ensure(F * (size_t)C);
    if (need_mask) mask.ensure(F * (size_t)W);
  }
  Level view() { return Level{tree.p, begin.p, end.p, node.p, depth.p, trace.p, key.p, hist.p, mask.p}; }
};

struct SearchBufs {
  DevBuf<uint8_t> flag, best_mil;
  DevBuf<int32_t> visited, nconst, dc, cand_begin, cand_cnt, best_feature, best_nleft, split_slot, scored, best_hl;
  DevBuf<double> best_score, best_cut, total, nsum, mean, dist;
  void ensure(size_t F, int C, int k, int task, bool replay) {
    flag.ensure(F);
    best_mil.ensure(F);
    visited.ensure(F);
    nconst.ensure(F);
    dc.ensure(F);
    cand_begin.ensure(F);
    cand_cnt.ensure(F);
    best_feature.ensure(F);
    best_nleft.ensure(F);
    split_slot.ensure(F);
    best_score.ensure(F);
    best_cut.ensure(F);
    total.ensure(F);
    nsum.ensure(F);
    mean.ensure(F);
    if (!replay) scored.ensure(F * (size_t)std::max(k, 1));
    if (task == TASK_CLS) best_hl.ensure(F * (size_t)C);
    if (task == TASK_CLSW) dist.ensure(F * (size_t)C);
  }
  Search view() {
    return Search{flag.p,       best_mil.p,   visited.p,    nconst.p,     dc.p,         cand_begin.p,
                  cand_cnt.p,   best_feature.p, best_nleft.p, split_slot.p, best_score.p, best_cut.p,
                  total.p,      nsum.p,       mean.p,       scored.p,     best_hl.p,    dist.p};
  }
};

struct CandBufs {
  DevBuf<int32_t> node, feature, cnt_lt, cnt_nan, hist;
  DevBuf<double> u, cut, score;
  DevBuf<uint8_t> flags, mil;
  DevBuf<int64_t> mask_off;
  void ensure(size_t N, int C, int task) {
    node.ensure(N);
    feature.ensure(N);
    cnt_lt.ensure(N);
    cnt_nan.ensure(N);
    u.ensure(N);
    cut.ensure(N);
    flags.ensure(N);
    if (task == TASK_CLS) {
      hist.ensure(N * 2 * (size_t)C);
    } else {
      score.ensure(N);
      mil.ensure(N);
      mask_off.ensure(N);
    }
  }
  Cand view() {
    return Cand{node.p, feature.p, cnt_lt.p, cnt_nan.p, u.p, cut.p, score.p, flags.p, mil.p, hist.p, mask_off.p};
  }
};

struct OutBufs {
  DevBuf<int32_t> feature, left, right;
  DevBuf<double> cut, leaf;
  DevBuf<uint8_t> mil;
  void grow(size_t n, size_t used, int lw, cudaStream_t st) {
    feature.grow_keep(n, used, st);
    left.grow_keep(n, used, st);
    right.grow_keep(n, used, st);
    cut.grow_keep(n, used, st);
    mil.grow_keep(n, used, st);
    leaf.grow_keep(n * (size_t)lw, used * (size_t)lw, st);
  }
  Out view() { return Out{feature.p, left.p, right.p, cut.p, leaf.p, mil.p}; }
};

struct EventTimer {
  std::vector<cudaEvent_t> pool;
  std::vector<std::pair<int, int>> spans[2];  // kind 0 = split search, 1 = partition
  size_t used = 0;
  ~EventTimer() {
    for (auto e : pool) cudaEventDestroy(e);
  }
  int rec(cudaStream_t st) {
    if (used == pool.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      pool.push_back(e);
    }
    cudaEventRecord(pool[used], st);
    return (int)used++;
  }
  // call after a stream sync
  void drain(double *acc) {
    for (int kx = 0; kx < 2; kx++) {
      for (auto &sp : spans[kx]) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, pool[(size_t)sp.first], pool[(size_t)sp.second]);
        acc[kx] += ms;
      }
      spans[kx].clear();
    }
    used = 0;
  }
};

template <typename T>
static T *upload_tmp(const T *h, size_t n, cudaStream_t st) {
  T *d = nullptr;
  if (cudaMalloc((void **)&d, std::max<size_t>(n, 1) * sizeof(T)) != cudaSuccess) {
    cudaGetLastError();
    ET_FAIL(ET_ENOMEM, "device allocation failed");
  }
  if (n) cudaMemcpyAsync(d, h, n * sizeof(T), cudaMemcpyHostToDevice, st);
  return d;
}

// BFS-numbered batch output -> per-tree pre-order HostTree
static void to_preorder(const std::vector<int32_t> &feature, const std::vector<int32_t> &left,
                        const std::vector<int32_t> &right, const std::vector<double> &cut,
                        const std::vector<uint8_t> &mil, const std::vector<double> &leaf, int lw, int32_t root,
                        HostTree &out) {
  std::vector<int32_t> order;   // batch ids in pre-order
  std::vector<int32_t> stack;
  stack.push_back(root);
  while (!stack.empty()) {
    int32_t v = stack.back();
    stack.pop_back();
    order.push_back(v);
    if (feature[(size_t)v] >= 0) {
      stack.push_back(right[(size_t)v]);
      stack.push_back(left[(size_t)v]);
    }
  }
  size_t n = order.size();
  out.feature.resize(n);
  out.left.assign(n, -1);
  out.right.assign(n, -1);
  out.cut.resize(n);
  out.mil.resize(n);
  out.leaf.assign(n * (size_t)lw, 0.0);
  // pre-order position of a right child = position right after the whole left subtree; recover
  // it with a second stack walk that records positions
  std::vector<std::pair<int32_t, int32_t>> st2;  // (batch id, parent pos or -1) ; sign encodes side
  size_t pos = 0;
  struct Fr {
    int32_t v, parent;
    bool is_right;
  };
  std::vector<Fr> st;
  st.push_back(Fr{root, -1, false});
  while (!st.empty()) {
    Fr fr = st.back();
    st.pop_back();
    int32_t me = (int32_t)pos++;
    size_t v = (size_t)fr.v;
    if (fr.parent >= 0) {
      if (fr.is_right)
        out.right[(size_t)fr.parent] = me;
      else
        out.left[(size_t)fr.parent] = me;
    }
    out.feature[(size_t)me] = feature[v];
    out.cut[(size_t)me] = cut[v];
    out.mil[(size_t)me] = mil[v];
    if (feature[v] >= 0) {
      st.push_back(Fr{right[v], me, true});
      st.push_back(Fr{left[v], me, false});
    } else {
      for (int c = 0; c < lw; c++) out.leaf[(size_t)me * lw + c] = leaf[v * (size_t)lw + c];
    }
  }
  (void)st2;
}

}  // namespace

void et_build_forest(et_ctx *ctx, et_data *D, const BuildArgs &a, et_forest *out, et_stats *stats) {
  cudaStream_t st = ctx->stream;
  const int task = a.task;
  const int C = (task == TASK_REG) ? 1 : a.num_classes;
  const int lw = C;
  const int64_t n = D->n;
  const int32_t d = D->d;
  const bool replay = a.replay != nullptr;
  const int W = (d + 31) / 32;
  if (a.k < 0) ET_FAIL(ET_EINVAL, "k must be >= 0");
  if (replay && a.replay->n_trees != a.m) ET_FAIL(ET_EREPLAY, "replay trace holds %d trees, m = %d", a.replay->n_trees, a.m);
  et_stats S;
  memset(&S, 0, sizeof(S));
  const int64_t launches0 = ctx->launches;
  cudaEvent_t ev0, ev1;
  CUDA_CHECK(cudaEventCreate(&ev0));
  CUDA_CHECK(cudaEventCreate(&ev1));
  CUDA_CHECK(cudaEventRecord(ev0, st));
  out->trees.clear();
  out->trees.resize((size_t)a.m);

  // the reference's seeding (pkg:629,654-655) names one stream per tree; the free-running GPU RNG
  // is counter based and keyed by (seed, global tree id)
  std::vector<uint64_t> tree_keys((size_t)std::max(a.m, 1));
  for (int t = 0; t < a.m; t++) {
    uint64_t gid = a.tree_ids ? (uint64_t)(uint32_t)a.tree_ids[t] : (uint64_t)t;
    tree_keys[(size_t)t] = et_tree_key((uint64_t)a.seed, gid);
  }

  // batch size: bound the per-sample state (idx + targets, ping-pong) to ~6 GB
  size_t per_sample = 2 * (4 + (task == TASK_REG ? 8 : 4) + (task == TASK_CLSW ? 8 : 0));
  int64_t max_samples = (int64_t)(((size_t)6 << 30) / per_sample);
  int32_t B = (int32_t)std::max<int64_t>(1, std::min<int64_t>(a.m, max_samples / std::max<int64_t>(n, 1)));
  if (const char *env = getenv("ETGPU_BATCH_TREES")) B = std::max(1, std::min(a.m, atoi(env)));

  DevBuf<int32_t> idx[2], yc[2];
  DevBuf<double> yr[2], ws[2];
  LevelBufs lv[2];
  SearchBufs sb;
  CandBufs cb[2];
  OutBufs ob;
  DevBuf<uint32_t> sidemask;
  DevBuf<double> wscratch;
  DevBuf<Counters> cnt;
  cnt.ensure(1);
  EventTimer timer;
  double tacc[2] = {0, 0};

  // root histogram on the device (TASK_CLS)
  std::vector<int32_t> rh((size_t)std::max(C, 1), 0);
  if (task == TASK_CLS)
    for (int c = 0; c < C; c++) rh[(size_t)c] = (int32_t)D->root_hist[(size_t)c];
  int32_t *d_root_hist = upload_tmp(rh.data(), rh.size(), st);

  // replay trace on the device (child ids rewritten to absolute node indices)
  int64_t *d_tr_cand_begin = nullptr;
  int32_t *d_tr_cand_count = nullptr, *d_tr_left = nullptr, *d_tr_right = nullptr, *d_tr_cand_feature = nullptr;
  double *d_tr_cand_u = nullptr;
  uint8_t *d_tr_cand_flag = nullptr;
  std::vector<int64_t> trace_roots;
  if (replay) {
    const et_replay *R = a.replay;
    int64_t nn = R->node_offset[R->n_trees];
    std::vector<int32_t> l((size_t)nn), r((size_t)nn);
    if (nn > 0x7fffffff) ET_FAIL(ET_EREPLAY, "replay trace too large");
    for (int t = 0; t < R->n_trees; t++) {
      int64_t o = R->node_offset[t];
      for (int64_t q = o; q < R->node_offset[t + 1]; q++) {
        l[(size_t)q] = R->left[q] >= 0 ? (int32_t)(o + R->left[q]) : -1;
        r[(size_t)q] = R->right[q] >= 0 ? (int32_t)(o + R->right[q]) : -1;
        if (R->cand_count[q] < 0 || R->cand_begin[q] < 0 || R->cand_begin[q] + R->cand_count[q] > R->n_cand)
          ET_FAIL(ET_EREPLAY, "replay trace: node %lld has candidates outside the trace", (long long)q);
      }
      trace_roots.push_back(o);
    }
    for (int64_t q = 0; q < R->n_cand; q++)
      if (R->cand_feature[q] < 0 || R->cand_feature[q] >= d) ET_FAIL(ET_EREPLAY, "replay trace: feature out of range");
    d_tr_cand_begin = upload_tmp(R->cand_begin, (size_t)nn, st);
    d_tr_cand_count = upload_tmp(R->cand_count, (size_t)nn, st);
    d_tr_left = upload_tmp(l.data(), (size_t)nn, st);
    d_tr_right = upload_tmp(r.data(), (size_t)nn, st);
    d_tr_cand_feature = upload_tmp(R->cand_feature, (size_t)R->n_cand, st);
    d_tr_cand_u = upload_tmp(R->cand_u, (size_t)R->n_cand, st);
    d_tr_cand_flag = upload_tmp(R->cand_flag, (size_t)R->n_cand, st);
    CUDA_CHECK(cudaStreamSynchronize(st));
  }
  auto free_tmp = [&]() {
    cudaFree(d_root_hist);
    cudaFree(d_tr_cand_begin);
    cudaFree(d_tr_cand_count);
    cudaFree(d_tr_left);
    cudaFree(d_tr_right);
    cudaFree(d_tr_cand_feature);
    cudaFree(d_tr_cand_u);
    cudaFree(d_tr_cand_flag);
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
  };

  try {
    for (int32_t t0 = 0; t0 < a.m; t0 += B) {
      const int32_t Bt = std::min(B, a.m - t0);
      const size_t ns = (size_t)Bt * (size_t)n;
      for (int q = 0; q < 2; q++) {
        idx[q].ensure(ns, 1.0);
        if (task == TASK_REG)
          yr[q].ensure(ns, 1.0);
        else
          yc[q].ensure(ns, 1.0);
        if (task == TASK_CLSW) ws[q].ensure(ns, 1.0);
      }
      P p;
      memset(&p, 0, sizeof(p));
      p.X = D->x;
      p.ld = D->ld;
      p.n = n;
      p.n_table = n;
      p.d = d;
      p.C = C;
      p.k = a.k;
      p.n_min = a.n_min;
      p.max_depth = a.max_depth;
      p.W = W;
      p.task = task;
      p.replay = replay ? 1 : 0;
      p.cnt = cnt.p;
      p.tr = Trace{d_tr_cand_begin, d_tr_cand_count, d_tr_left, d_tr_right, d_tr_cand_feature, d_tr_cand_u,
                   d_tr_cand_flag};
      int srcb = 0;
      {
        unsigned grid = (unsigned)std::min<int64_t>(ceil_div((int64_t)ns, 256), (int64_t)ctx->sm_count * 16);
        k_init_samples<<<std::max(grid, 1u), 256, 0, st>>>(n, Bt, idx[0].p, D->y_cls, task == TASK_REG ? nullptr : yc[0].p,
                                                          D->y_reg, task == TASK_REG ? yr[0].p : nullptr, D->w,
                                                          task == TASK_CLSW ? ws[0].p : nullptr);
        ctx->launches++;
      }
      int32_t F = Bt;
      int cl = 0;  // current level buffer
      lv[cl].ensure((size_t)F, C, W, task == TASK_CLS, !replay);
      p.cur = lv[cl].view();
      uint64_t *d_keys = upload_tmp(tree_keys.data() + t0, (size_t)Bt, st);
      int64_t *d_troots = replay ? upload_tmp(trace_roots.data() + t0, (size_t)Bt, st) : nullptr;
      k_init_roots<<<(unsigned)ceil_div(F, 128), 128, 0, st>>>(p, Bt, d_keys, d_troots, d_root_hist);
      ctx->launches++;
      CUDA_CHECK(cudaStreamSynchronize(st));
      cudaFree(d_keys);
      if (d_troots) cudaFree(d_troots);
      CUDA_CHECK(cudaMemsetAsync(cnt.p, 0, sizeof(Counters), st));

      int64_t n_nodes = Bt;  // output nodes allocated so far (roots)
      ob.grow((size_t)n_nodes, 0, lw, st);
      while (F > 0) {
        S.levels++;
        sb.ensure((size_t)F, C, a.k, task, replay);
        lv[cl ^ 1].ensure((size_t)F * 2, C, W, task == TASK_CLS, !replay);
        ob.grow((size_t)(n_nodes + 2 * (int64_t)F), (size_t)n_nodes, lw, st);
        p.idx_src = idx[srcb].p;
        p.idx_dst = idx[srcb ^ 1].p;
        p.yc_src = yc[srcb].p;
        p.yc_dst = yc[srcb ^ 1].p;
        p.yr_src = yr[srcb].p;
        p.yr_dst = yr[srcb ^ 1].p;
        p.w_src = ws[srcb].p;
        p.w_dst = ws[srcb ^ 1].p;
        p.cur = lv[cl].view();
        p.nxt = lv[cl ^ 1].view();
        p.s = sb.view();
        p.o = ob.view();
        p.node_base_next = (int32_t)n_nodes;
        const unsigned gF = (unsigned)ceil_div(F, 128);
        if (task == TASK_CLS)
          k_classify_cls<<<gF, 128, 0, st>>>(p, F);
        else if (task == TASK_REG)
          k_classify_reg<<<gF, 128, 0, st>>>(p, F);
        else
          k_classify_clsw<<<gF, 128, 0, st>>>(p, F);
        ctx->launches++;
        // candidate rounds
        size_t cand_cap = (size_t)F * (size_t)std::max(a.k, 1);
        if (replay) {
          // a node's trace may hold more than k draws (constant hits): bound by the largest count
          cand_cap = (size_t)a.replay->n_cand;
        }
        for (int q = 0; q < 2; q++) cb[q].ensure(cand_cap, C, task);
        p.c[0] = cb[0].view();
        p.c[1] = cb[1].view();
        for (int round = 0;; round++) {
          const int cur = round & 1;
          Counters hc;
          // reset this round's counters (n_cand[cur], mask_words)
          CUDA_CHECK(cudaMemsetAsync(&cnt.p->n_cand[cur], 0, sizeof(int32_t), st));
          CUDA_CHECK(cudaMemsetAsync(&cnt.p->mask_words, 0, sizeof(unsigned long long), st));
          k_update_draw<<<gF, 128, 0, st>>>(p, F, round);
          ctx->launches++;
          CUDA_CHECK(cudaMemcpyAsync(&hc, cnt.p, sizeof(Counters), cudaMemcpyDeviceToHost, st));
          CUDA_CHECK(cudaStreamSynchronize(st));
          const int32_t nc = hc.n_cand[cur];
          if (nc == 0) break;
          S.rounds++;
          if ((size_t)nc > cand_cap) ET_FAIL(ET_ECUDA, "internal: candidate buffer overflow (%d > %zu)", nc, cand_cap);
          if (task != TASK_CLS) {
            sidemask.ensure((size_t)hc.mask_words + 64);
            p.sidemask = sidemask.p;
            if (task == TASK_CLSW) {
              wscratch.ensure((size_t)nc * 2 * (size_t)C);
              p.wscratch = wscratch.p;
            }
          }
          const int wpb = 8;
          const unsigned gi = (unsigned)ceil_div(nc, wpb);
          int e0 = timer.rec(st);
          if (task == TASK_CLS) {
            size_t smem = (size_t)wpb * 2 * C * sizeof(int32_t);
            if (smem > 48 * 1024)
              CUDA_CHECK(cudaFuncSetAttribute(k_items_warp<TASK_CLS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)smem));
            k_items_warp<TASK_CLS><<<gi, wpb * 32, smem, st>>>(p, nc, cur);
          } else {
            k_items_warp<TASK_REG><<<gi, wpb * 32, 0, st>>>(p, nc, cur);
          }
          ctx->launches++;
          if (task == TASK_REG) {
            k_score_reg<<<(unsigned)ceil_div(nc, 64), 64, 0, st>>>(p, nc, cur);
            ctx->launches++;
          } else if (task == TASK_CLSW) {
            k_score_clsw<<<(unsigned)ceil_div(nc, 64), 64, 0, st>>>(p, nc, cur);
            ctx->launches++;
          }
          int e1 = timer.rec(st);
          timer.spans[0].push_back({e0, e1});
          CUDA_CHECK(cudaGetLastError());
        }
        // finalize + partition
        k_finalize<<<gF, 128, 0, st>>>(p, F);
        ctx->launches++;
        Counters hc;
        CUDA_CHECK(cudaMemcpyAsync(&hc, cnt.p, sizeof(Counters), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        timer.drain(tacc);
        const int32_t nf = hc.next_f;
        if (nf > 0) {
          int e0 = timer.rec(st);
          k_partition_warp<<<(unsigned)ceil_div(F, 8), 256, 0, st>>>(p, F);
          ctx->launches++;
          int e1 = timer.rec(st);
          timer.spans[1].push_back({e0, e1});
          srcb ^= 1;
        }
        CUDA_CHECK(cudaMemsetAsync(&cnt.p->next_f, 0, sizeof(int32_t), st));
        n_nodes += nf;
        if (n_nodes > 0x7ffffff0) ET_FAIL(ET_EUNSUPPORTED, "batch exceeds 2^31 nodes; lower ETGPU_BATCH_TREES");
        F = nf;
        cl ^= 1;
        CUDA_CHECK(cudaGetLastError());
      }
      // batch output -> host, BFS numbering -> per-tree pre-order
      Counters hc;
      CUDA_CHECK(cudaMemcpyAsync(&hc, cnt.p, sizeof(Counters), cudaMemcpyDeviceToHost, st));
      std::vector<int32_t> hf((size_t)n_nodes), hl((size_t)n_nodes), hr((size_t)n_nodes);
      std::vector<double> hcut((size_t)n_nodes), hleaf((size_t)n_nodes * (size_t)lw);
      std::vector<uint8_t> hmil((size_t)n_nodes);
      CUDA_CHECK(cudaMemcpyAsync(hf.data(), ob.feature.p, hf.size() * 4, cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaMemcpyAsync(hl.data(), ob.left.p, hl.size() * 4, cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaMemcpyAsync(hr.data(), ob.right.p, hr.size() * 4, cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaMemcpyAsync(hcut.data(), ob.cut.p, hcut.size() * 8, cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaMemcpyAsync(hmil.data(), ob.mil.p, hmil.size(), cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaMemcpyAsync(hleaf.data(), ob.leaf.p, hleaf.size() * 8, cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaStreamSynchronize(st));
      timer.drain(tacc);
      S.v_mm += (int64_t)hc.st[ST_VMM];
      S.v_sc += (int64_t)hc.st[ST_VSC];
      S.s_rows += (int64_t)hc.st[ST_SROWS];
      S.p_rows += (int64_t)hc.st[ST_PROWS];
      S.draws += (int64_t)hc.st[ST_DRAWS];
      S.const_hits += (int64_t)hc.st[ST_CONST];
      S.scored += (int64_t)hc.st[ST_SCORED];
      S.replay_mismatches += (int64_t)hc.st[ST_MISMATCH];
      S.nodes += n_nodes;
      for (int32_t t = 0; t < Bt; t++)
        to_preorder(hf, hl, hr, hcut, hmil, hleaf, lw, t, out->trees[(size_t)(t0 + t)]);
    }
    CUDA_CHECK(cudaEventRecord(ev1, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    S.gpu_ms = ms;
    S.gpu_ms_split = tacc[0];
    S.gpu_ms_partition = tacc[1];
    S.launches = ctx->launches - launches0;
  } catch (...) {
    cudaStreamSynchronize(st);
    free_tmp();
    throw;
  }
  free_tmp();
  if (stats) *stats = S;
  if (replay && S.replay_mismatches && !stats)  // with stats the caller reads replay_mismatches itself
    ET_FAIL(ET_EREPLAY, "replay: %lld decisions contradict the trace (wrong data for this trace?)",
            (long long)S.replay_mismatches);
}
