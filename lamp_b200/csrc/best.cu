// best.cu -- bestSplit = true.
//
// splitBestClassification (pkg:56-202) / splitBestRegression (pkg:298-426): for every drawn non-constant
// feature EVERY sample value of the node is tried as the cutpoint (score(i) for i = 0 .. n-1 in subset order, the
// first maximum wins with strict `>`, default index 0; pkg:133-145), then the feature competes with that
// cutpoint like a random split does.  The reference is O(n^2) per feature; here one thread owns one cutpoint
// and streams the node once per class (sums stay in subset order, so weighted and regression scores are
// bit-exact), 256 cutpoints per CTA, all (node, candidate, chunk) work items of a level side by side:
//
//   k_best_prep     CTA per node    stop rules (pkg:993-994 / 813-814), node impurity, search state
//   k_best_draw     thread per node next batch of candidate features (replayed trace | counter RNG), work items
//   k_best_eval     CTA per item    min / max / hasMissing, scores of 256 cutpoints, first maximum of the chunk
//   k_best_consume  thread per node candidates in draw order: constants, NaN scores, first-best (pkg:176-195)
//   k_best_finish   CTA per node    leaf | stable partition (pkg:1024-1039) and the two children
//
// Tables are read as FP64 (column-major X); this path is the reference's secondary variant.
#include "build.cuh"

namespace etb {

struct BestState {  // search state per open node of the level (SoA over the frontier index)
  double *total, *nsum, *leaf_mean, *best_score, *best_cut, *dist;
  int32_t *flags;  // bit 0: leaf by stop rule, bit 1: search finished
  int32_t *visited, *nconst, *dc, *tpos, *best_feature, *best_mil, *ncand, *nchunk, *hist;
  int32_t *cand_feat, *cand_expect;  // [F][32]
  long long *item_off;
  uint32_t *taken;                   // [F][W]
  unsigned long long *counters;      // [0] work items of the round, [1] nodes still searching
};
struct BestItem {
  int32_t node_i, cand, chunk;
};
struct __align__(8) BestRes {
  double score, cut;
  int32_t idx, flags;  // flags: bit 0 constant feature, bit 1 missing-is-less, bit 2 the chunk has a winner
};
constexpr int BEST_CTA = 256, BEST_TILE = 512;

template <int TASK>
__global__ void __launch_bounds__(BEST_CTA) k_best_prep(P p, BestState s, int32_t qcount) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int32_t *sh_hist = reinterpret_cast<int32_t *>(smem_raw);
  __shared__ int sh_flag;
  const int q = blockIdx.x, tid = threadIdx.x;
  if (q >= qcount) return;
  const int C = p.C, W = p.W;
  const int i = p.q_cur[Q_CTA][q];
  const int32_t tree = p.cur.tree[i], b = p.cur.begin[i], n = p.cur.end[i] - b, depth = p.cur.depth[i];
  const int64_t base = (int64_t)tree * p.n + b;
  bool leaf;
  double total = 0.0, nsum = (double)n, leaf_mean = 0.0;
  if (tid == 0) sh_flag = 1;
  __syncthreads();
  if (TASK != TASK_REG) {
    for (int c = tid; c < C; c += BEST_CTA) sh_hist[c] = 0;
    __syncthreads();
    for (int32_t j = tid; j < n; j += BEST_CTA) atomicAdd(&sh_hist[p.yc_src[base + j]], 1);
    __syncthreads();
    bool pure = false;
    for (int c = 0; c < C; c++) pure |= (sh_hist[c] == n);
    leaf = (p.n_table < p.n_min) || (depth >= p.max_depth) || pure;  // pkg:993-994
    for (int c = tid; c < C; c += BEST_CTA) s.hist[(int64_t)i * C + c] = sh_hist[c];
    if (TASK == TASK_CLS) {
      const double inv = ET_DIV(1.0, (double)n);
      for (int c = tid; c < C; c += BEST_CTA) s.dist[(int64_t)i * C + c] = repeat_add_dev(inv, sh_hist[c]);
    } else {
      // weighted distribution (pkg:913-927): total and per-class sums, each sequential in subset order
      __shared__ double sh_s;
      if (tid == 0) {
        double a = 0.0;
        for (int32_t j = 0; j < n; j++) a = ET_ADD(a, p.w_src[base + j]);
        sh_s = a;
      }
      __syncthreads();
      for (int c = tid; c < C; c += BEST_CTA) {
        double a = 0.0;
        for (int32_t j = 0; j < n; j++)
          if (p.yc_src[base + j] == c) a = ET_ADD(a, p.w_src[base + j]);
        s.dist[(int64_t)i * C + c] = ET_DIV(a, sh_s);
      }
      nsum = sh_s;
    }
    __syncthreads();
    if (tid == 0) {
      double sq = 0.0;
      for (int c = 0; c < C; c++) sq = ET_ADD(sq, ET_MUL(s.dist[(int64_t)i * C + c], s.dist[(int64_t)i * C + c]));
      total = ET_SUB(1.0, sq);
    }
  } else {
    const double head = p.yr_src[base];
    bool uni_l = true;
    for (int32_t j = tid; j < n; j += BEST_CTA) uni_l &= !(p.yr_src[base + j] != head);
    if (!uni_l) sh_flag = 0;
    __syncthreads();
    const bool uni = sh_flag != 0;
    leaf = (n < p.n_min) || (depth >= p.max_depth) || uni;  // pkg:813-814
    if (tid == 0) {  // mean2 (pkg:782) and varianceNoSplit (pkg:307-308), sequential in subset order
      double sum = 0.0;
      for (int32_t j = 0; j < n; j++) sum = ET_ADD(sum, p.yr_src[base + j]);
      const double dn = (double)n;
      leaf_mean = ET_DIV(sum, dn);
      double var = 0.0;
      if (n > 1) {
        double qq = 0.0;
        for (int32_t j = 0; j < n; j++) {
          const double dl = ET_SUB(p.yr_src[base + j], leaf_mean);
          qq = ET_ADD(qq, ET_MUL(dl, dl));
        }
        var = ET_DIV(qq, ET_SUB(dn, 1.0));
      }
      total = ET_DIV(ET_MUL(var, ET_SUB(dn, 1.0)), dn);
    }
  }
  // the features known constant on the path from the root are taken from the start (free-running)
  __shared__ int sh_nc;
  if (tid == 0) sh_nc = 0;
  __syncthreads();
  if (!p.replay) {
    int nc = 0;
    for (int w = tid; w < W; w += BEST_CTA) {
      const uint32_t m = p.cur.mask[(int64_t)i * W + w];
      s.taken[(int64_t)i * W + w] = m;
      nc += __popc(m);
    }
    if (nc) atomicAdd(&sh_nc, nc);
  }
  __syncthreads();
  if (tid == 0) {
    s.total[i] = total;
    s.nsum[i] = nsum;
    s.leaf_mean[i] = leaf_mean;
    s.best_score[i] = -INFINITY;
    s.best_cut[i] = NAN;
    s.best_feature[i] = -1;
    s.best_mil[i] = 0;
    s.visited[i] = 0;
    s.nconst[i] = p.replay ? 0 : sh_nc - (W * 32 - p.d);
    s.dc[i] = 0;
    s.tpos[i] = 0;
    s.ncand[i] = 0;
    s.flags[i] = leaf ? 3 : 0;
    if (!leaf) atomicAdd(&p.cnt->st[ST_SROWS], (unsigned long long)n);
  }
}

__global__ void k_best_draw(P p, BestState s, int32_t qcount, BestItem *items) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= qcount) return;
  const int i = p.q_cur[Q_CTA][q];
  s.ncand[i] = 0;
  if (s.flags[i] & 2) return;
  const int W = p.W;
  const int32_t n = p.cur.end[i] - p.cur.begin[i];
  int32_t nb;
  const int64_t tn = p.cur.trace[i];
  if (p.replay) {
    const int32_t tcnt = (tn >= 0) ? p.tr.cand_count[tn] : 0;
    nb = min(32, tcnt - s.tpos[i]);
  } else {
    const int32_t avail = p.d - s.nconst[i] - s.visited[i];
    nb = min(32, min(p.k - s.visited[i], avail));
  }
  if (nb <= 0) {
    s.flags[i] |= 2;
    return;
  }
  if (p.replay) {
    const int64_t tb = p.tr.cand_begin[tn] + s.tpos[i];
    for (int c = 0; c < nb; c++) {
      s.cand_feat[(int64_t)i * 32 + c] = p.tr.cand_feature[tb + c];
      s.cand_expect[(int64_t)i * 32 + c] = p.tr.cand_flag[tb + c] + 1;
    }
    s.tpos[i] += nb;
  } else {
    // uniform without replacement over the features not taken yet (sequential: no duplicates)
    uint32_t *taken = s.taken + (int64_t)i * W;
    const uint64_t key = p.cur.key[i];
    int32_t left = p.d - s.nconst[i] - s.visited[i], dc = s.dc[i];
    for (int c = 0; c < nb; c++) {
      const uint64_t r = et_draw(key, (uint32_t)dc++);
      const int32_t f = rank_select_clear_fast(taken, W, (int32_t)__umul64hi(r, (uint64_t)left));
      taken[f >> 5] |= 1u << (f & 31);
      left--;
      s.cand_feat[(int64_t)i * 32 + c] = f;
      s.cand_expect[(int64_t)i * 32 + c] = 0;
    }
    s.dc[i] = dc;
  }
  const int32_t nchunk = (n + BEST_CTA - 1) / BEST_CTA;
  const long long off = (long long)atomicAdd(&s.counters[0], (unsigned long long)nb * (unsigned long long)nchunk);
  s.ncand[i] = nb;
  s.nchunk[i] = nchunk;
  s.item_off[i] = off;
  for (int c = 0; c < nb; c++)
    for (int ch = 0; ch < nchunk; ch++) items[off + (long long)c * nchunk + ch] = BestItem{i, c, ch};
  atomicAdd(&s.counters[1], 1ull);
}

// first maximum of (score, index) over the CTA: NaN never wins, ties go to the smaller index (pkg:137-141)
__device__ __forceinline__ void best_block_argmax(double &sc, int32_t &ix, int &ml, double *shd, int32_t *shi) {
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double os = __shfl_xor_sync(0xffffffffu, sc, o);
    const int32_t oi = __shfl_xor_sync(0xffffffffu, ix, o);
    const int om = __shfl_xor_sync(0xffffffffu, ml, o);
    if (oi >= 0 && (ix < 0 || os > sc || (os == sc && oi < ix))) {
      sc = os;
      ix = oi;
      ml = om;
    }
  }
  if (lane == 0) {
    shd[wp] = sc;
    shi[wp] = ix;
    shi[8 + wp] = ml;
  }
  __syncthreads();
  if (tid == 0) {
    for (int w2 = 1; w2 < BEST_CTA / 32; w2++) {
      const double os = shd[w2];
      const int32_t oi = shi[w2];
      if (oi >= 0 && (ix < 0 || os > sc || (os == sc && oi < ix))) {
        sc = os;
        ix = oi;
        ml = shi[8 + w2];
      }
    }
  }
}

template <int TASK>
__global__ void __launch_bounds__(BEST_CTA) k_best_eval(P p, BestState s, const BestItem *items, BestRes *res) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *xs = reinterpret_cast<double *>(smem_raw);                   // [BEST_TILE] the node's values, by position
  double *ys = xs + BEST_TILE;                                         // [BEST_TILE] targets (REG) | weights (CLSW)
  int32_t *cs = reinterpret_cast<int32_t *>(ys + BEST_TILE);           // [BEST_TILE] labels
  int32_t *sh_hn = cs + BEST_TILE;                                     // [C] NaN rows per class (CLS)
  __shared__ double shd[8];
  __shared__ int32_t shi[16];
  __shared__ double sh_mn[8], sh_mx[8];
  __shared__ int sh_nan, sh_nn;
  const BestItem it = items[blockIdx.x];
  const int i = it.node_i, tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  const int C = p.C;
  const int32_t f = s.cand_feat[(int64_t)i * 32 + it.cand];
  const int32_t tree = p.cur.tree[i], b = p.cur.begin[i], n = p.cur.end[i] - b;
  const int64_t base = (int64_t)tree * p.n + b;
  const int32_t *idx = p.idx_src + base;
  const Col col = col_of(p, f);
  // ---- minmax / hasMissing (pkg:34-54) and the NaN rows per class
  if (tid == 0) {
    sh_nan = 0;
    sh_nn = 0;
  }
  if (TASK == TASK_CLS)
    for (int c = tid; c < C; c += BEST_CTA) sh_hn[c] = 0;
  __syncthreads();
  double mn = 1.7976931348623157e308, mx = -1.7976931348623157e308;
  int nn = 0;
  for (int32_t j = tid; j < n; j += BEST_CTA) {
    const double x = col_at(col, idx[j]);
    if (x < mn) mn = x;
    if (x > mx) mx = x;
    if (x != x) {
      nn++;
      if (TASK == TASK_CLS) atomicAdd(&sh_hn[p.yc_src[base + j]], 1);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if (lane == 0) {
    sh_mn[wp] = mn;
    sh_mx[wp] = mx;
  }
  if (nn) {
    atomicOr(&sh_nan, 1);
    atomicAdd(&sh_nn, nn);
  }
  __syncthreads();
  for (int w2 = 0; w2 < BEST_CTA / 32; w2++) {
    mn = fmin(mn, sh_mn[w2]);
    mx = fmax(mx, sh_mx[w2]);
  }
  const bool has_nan = sh_nan != 0;
  const int32_t nan_total = sh_nn;
  if ((mx <= mn) && !has_nan) {  // constant over the node (pkg:89)
    if (tid == 0) res[blockIdx.x] = BestRes{NAN, NAN, -1, 1};
    return;
  }
  // ---- this thread's cutpoint
  const int32_t ii = it.chunk * BEST_CTA + tid;
  const bool active = ii < n;
  const double cut = active ? col_at(col, idx[ii]) : 0.0;
  const double G = s.total[i];
  double s_not = NAN, s_mil = NAN;
  // streams the node through shared memory; `body(j)` sees xs[j], ys[j], cs[j] of every position in subset order
  auto stream = [&](auto body) {
    for (int32_t t0 = 0; t0 < n; t0 += BEST_TILE) {
      const int32_t tn2 = min(BEST_TILE, n - t0);
      __syncthreads();
      for (int32_t j = tid; j < tn2; j += BEST_CTA) {
        xs[j] = col_at(col, idx[t0 + j]);
        if (TASK == TASK_REG) ys[j] = p.yr_src[base + t0 + j];
        if (TASK == TASK_CLSW) ys[j] = p.w_src[base + t0 + j];
        if (TASK != TASK_REG) cs[j] = p.yc_src[base + t0 + j];
      }
      __syncthreads();
      if (active)
        for (int32_t j = 0; j < tn2; j++) body(j);
    }
  };
  if (TASK == TASK_CLS) {
    int32_t lt = 0;
    stream([&](int32_t j) { lt += (xs[j] < cut) ? 1 : 0; });
    const int32_t cin_n = lt, cin_m = lt + nan_total;
    double sin_n = 0.0, sout_n = 0.0, sin_m = 0.0, sout_m = 0.0;
    const double dcin_n = (double)cin_n, dcout_n = (double)(n - cin_n), dcin_m = (double)cin_m, dcout_m = (double)(n - cin_m);
    for (int c = 0; c < C; c++) {
      const int32_t ht = s.hist[(int64_t)i * C + c];
      if (ht == 0) continue;  // contributes exactly +0.0 (or the NaN every present class gives as well)
      int32_t hi = 0;
      stream([&](int32_t j) { hi += (cs[j] == c && xs[j] < cut) ? 1 : 0; });
      {
        const double pi = ET_DIV((double)hi, dcin_n), po = ET_DIV((double)(ht - hi), dcout_n);
        sin_n = ET_ADD(sin_n, ET_MUL(pi, pi));
        sout_n = ET_ADD(sout_n, ET_MUL(po, po));
      }
      if (has_nan) {
        const int32_t hm = hi + sh_hn[c];
        const double pi = ET_DIV((double)hm, dcin_m), po = ET_DIV((double)(ht - hm), dcout_m);
        sin_m = ET_ADD(sin_m, ET_MUL(pi, pi));
        sout_m = ET_ADD(sout_m, ET_MUL(po, po));
      }
    }
    const double N = (double)n;
    s_not = ET_SUB(ET_SUB(G, ET_DIV(ET_MUL(ET_SUB(1.0, sin_n), dcin_n), N)), ET_DIV(ET_MUL(ET_SUB(1.0, sout_n), dcout_n), N));
    if (has_nan)
      s_mil = ET_SUB(ET_SUB(G, ET_DIV(ET_MUL(ET_SUB(1.0, sin_m), dcin_m), N)), ET_DIV(ET_MUL(ET_SUB(1.0, sout_m), dcout_m), N));
  } else if (TASK == TASK_CLSW) {
    double cin_n = 0.0, cout_n = 0.0, cin_m = 0.0, cout_m = 0.0;
    stream([&](int32_t j) {
      const double x = xs[j], ww = ys[j];
      const bool ln = x < cut, lm = ln || (x != x);
      if (ln) cin_n = ET_ADD(cin_n, ww); else cout_n = ET_ADD(cout_n, ww);
      if (lm) cin_m = ET_ADD(cin_m, ww); else cout_m = ET_ADD(cout_m, ww);
    });
    double sin_n = 0.0, sout_n = 0.0, sin_m = 0.0, sout_m = 0.0;
    for (int c = 0; c < C; c++) {
      if (s.hist[(int64_t)i * C + c] == 0) continue;
      double hin_n = 0.0, hout_n = 0.0, hin_m = 0.0, hout_m = 0.0;
      stream([&](int32_t j) {
        if (cs[j] != c) return;
        const double x = xs[j], ww = ys[j];
        const bool ln = x < cut, lm = ln || (x != x);
        if (ln) hin_n = ET_ADD(hin_n, ww); else hout_n = ET_ADD(hout_n, ww);
        if (lm) hin_m = ET_ADD(hin_m, ww); else hout_m = ET_ADD(hout_m, ww);
      });
      {
        const double pi = ET_DIV(hin_n, cin_n), po = ET_DIV(hout_n, cout_n);
        sin_n = ET_ADD(sin_n, ET_MUL(pi, pi));
        sout_n = ET_ADD(sout_n, ET_MUL(po, po));
      }
      {
        const double pi = ET_DIV(hin_m, cin_m), po = ET_DIV(hout_m, cout_m);
        sin_m = ET_ADD(sin_m, ET_MUL(pi, pi));
        sout_m = ET_ADD(sout_m, ET_MUL(po, po));
      }
    }
    const double N = s.nsum[i];
    s_not = ET_SUB(ET_SUB(G, ET_DIV(ET_MUL(ET_SUB(1.0, sin_n), cin_n), N)), ET_DIV(ET_MUL(ET_SUB(1.0, sout_n), cout_n), N));
    if (has_nan)
      s_mil = ET_SUB(ET_SUB(G, ET_DIV(ET_MUL(ET_SUB(1.0, sin_m), cin_m), N)), ET_DIV(ET_MUL(ET_SUB(1.0, sout_m), cout_m), N));
  } else {
    double sin_n = 0.0, sout_n = 0.0, sin_m = 0.0, sout_m = 0.0;
    int32_t nin_n = 0, nin_m = 0;
    stream([&](int32_t j) {
      const double x = xs[j], v = ys[j];
      const bool ln = x < cut, lm = ln || (x != x);
      if (ln) { sin_n = ET_ADD(sin_n, v); nin_n++; } else sout_n = ET_ADD(sout_n, v);
      if (lm) { sin_m = ET_ADD(sin_m, v); nin_m++; } else sout_m = ET_ADD(sout_m, v);
    });
    const double mean_in_n = ET_DIV(sin_n, (double)nin_n), mean_out_n = ET_DIV(sout_n, (double)(n - nin_n));
    const double mean_in_m = ET_DIV(sin_m, (double)nin_m), mean_out_m = ET_DIV(sout_m, (double)(n - nin_m));
    double qin_n = 0.0, qout_n = 0.0, qin_m = 0.0, qout_m = 0.0;
    stream([&](int32_t j) {
      const double x = xs[j], v = ys[j];
      const bool ln = x < cut, lm = ln || (x != x);
      if (ln) { const double dl = ET_SUB(v, mean_in_n); qin_n = ET_ADD(qin_n, ET_MUL(dl, dl)); }
      else { const double dl = ET_SUB(v, mean_out_n); qout_n = ET_ADD(qout_n, ET_MUL(dl, dl)); }
      if (lm) { const double dl = ET_SUB(v, mean_in_m); qin_m = ET_ADD(qin_m, ET_MUL(dl, dl)); }
      else { const double dl = ET_SUB(v, mean_out_m); qout_m = ET_ADD(qout_m, ET_MUL(dl, dl)); }
    });
    auto vr = [&](int32_t nin, double qin, double qout) {  // computeVarianceReduction, pkg:1196-1218
      const int32_t nout = n - nin;
      const double dnin = (double)nin, dnout = (double)nout, dn = (double)n;
      const double svin = nin < 1 ? NAN : (nin == 1 ? 0.0 : ET_DIV(qin, ET_SUB(dnin, 1.0)));
      const double svout = nout < 1 ? NAN : (nout == 1 ? 0.0 : ET_DIV(qout, ET_SUB(dnout, 1.0)));
      const double vin = (nin == 1) ? 0.0 : ET_DIV(ET_MUL(svin, ET_SUB(dnin, 1.0)), dnin);
      const double vout = (nout == 1) ? 0.0 : ET_DIV(ET_MUL(svout, ET_SUB(dnout, 1.0)), dnout);
      const double a = ET_MUL(ET_DIV(dnin, dn), vin);
      const double bq = ET_MUL(ET_DIV(dnout, dn), vout);
      return ET_DIV(ET_SUB(ET_SUB(G, a), bq), G);
    };
    s_not = vr(nin_n, qin_n, qout_n);
    if (has_nan) s_mil = vr(nin_m, qin_m, qout_m);
  }
  int ml = (!(s_mil != s_mil) && (s_mil > s_not || (s_not != s_not))) ? 1 : 0;  // pkg:121-126
  double sc = ml ? s_mil : s_not;
  int32_t ix = (active && !(sc != sc)) ? ii : -1;
  best_block_argmax(sc, ix, ml, shd, shi);
  if (tid == 0) {
    BestRes r;
    r.score = sc;
    r.idx = ix;
    r.cut = (ix >= 0) ? col_at(col, idx[ix]) : NAN;
    r.flags = (ml ? 2 : 0) | (ix >= 0 ? 4 : 0);
    res[blockIdx.x] = r;
  }
}

__global__ void k_best_consume(P p, BestState s, int32_t qcount, const BestRes *res) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= qcount) return;
  const int i = p.q_cur[Q_CTA][q];
  const int32_t nc = s.ncand[i];
  if (nc <= 0) return;
  const int32_t n = p.cur.end[i] - p.cur.begin[i], nchunk = s.nchunk[i];
  const long long off = s.item_off[i];
  int32_t visited = s.visited[i], nconst = s.nconst[i], best_feature = s.best_feature[i], best_mil = s.best_mil[i];
  double best_score = s.best_score[i], best_cut = s.best_cut[i];
  unsigned long long st_draws = 0, st_const = 0, st_scored = 0, st_mismatch = 0;
  for (int c = 0; c < nc; c++) {
    if (!p.replay && visited >= p.k) break;  // the reference stopped drawing here
    const int32_t f = s.cand_feat[(int64_t)i * 32 + c], expect = s.cand_expect[(int64_t)i * 32 + c];
    const BestRes *r = res + off + (long long)c * nchunk;
    st_draws++;
    if (r[0].flags & 1) {
      nconst++;
      st_const++;
      if (!p.replay) p.cur.mask[(int64_t)i * p.W + (f >> 5)] |= 1u << (f & 31);
      if (p.replay && expect != 1) st_mismatch++;
      continue;
    }
    double smax = -INFINITY, cut = NAN;  // chunks hold ascending sample ranges: strict > keeps the first maximum
    int mil = 0;
    bool found = false;
    for (int ch = 0; ch < nchunk; ch++) {
      if ((r[ch].flags & 4) && r[ch].score > smax) {
        smax = r[ch].score;
        cut = r[ch].cut;
        mil = (r[ch].flags >> 1) & 1;
        found = true;
      }
    }
    st_scored++;
    if (!found) {  // every cutpoint scores NaN (pkg:186-188)
      nconst++;
      if (!p.replay) p.cur.mask[(int64_t)i * p.W + (f >> 5)] |= 1u << (f & 31);
      if (p.replay && expect != 3) st_mismatch++;
      continue;
    }
    if (smax > best_score) {  // strict >: the first best wins (pkg:176)
      best_score = smax;
      best_feature = f;
      best_cut = cut;
      best_mil = mil;
    }
    visited++;
    if (p.replay && expect != 2) st_mismatch++;
  }
  s.visited[i] = visited;
  s.nconst[i] = nconst;
  s.best_feature[i] = best_feature;
  s.best_mil[i] = best_mil;
  s.best_score[i] = best_score;
  s.best_cut[i] = best_cut;
  s.ncand[i] = 0;
  atomicAdd(&p.cnt->st[ST_DRAWS], st_draws);
  atomicAdd(&p.cnt->st[ST_CONST], st_const);
  atomicAdd(&p.cnt->st[ST_SCORED], st_scored);
  atomicAdd(&p.cnt->st[ST_VMM], (unsigned long long)n * st_draws);
  atomicAdd(&p.cnt->st[ST_VSC], (unsigned long long)n * (unsigned long long)(n + 1) * st_scored);
  if (st_mismatch) atomicAdd(&p.cnt->st[ST_MISMATCH], st_mismatch);
}

template <int TASK>
__global__ void __launch_bounds__(BEST_CTA) k_best_finish(P p, BestState s, int32_t qcount) {
  __shared__ int32_t sh_cnt[BEST_CTA / 32];
  __shared__ int32_t sh_misc[4];
  const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  if (q >= qcount) return;
  const int C = p.C, W = p.W;
  const int lw = (TASK == TASK_REG) ? 1 : C;
  const int i = p.q_cur[Q_CTA][q];
  const int32_t tree = p.cur.tree[i], b = p.cur.begin[i], e = p.cur.end[i], n = e - b;
  const int32_t node = p.cur.node[i], depth = p.cur.depth[i];
  const int64_t tn = p.cur.trace[i];
  const uint64_t key = p.cur.key[i];
  const int64_t base = (int64_t)tree * p.n;
  const bool leaf = (s.flags[i] & 1) != 0;
  const int32_t best_feature = s.best_feature[i], best_mil = s.best_mil[i];
  const double best_cut = s.best_cut[i];
  // pkg:197-200: no counted candidate, or a NaN cutpoint, gives feature -1
  const bool make_leaf = leaf || s.visited[i] == 0 || (best_cut != best_cut) || best_feature < 0;
  if (tid == 0 && p.replay) {
    const bool trace_split = tn >= 0 && p.tr.left[tn] >= 0;
    if (trace_split == make_leaf) atomicAdd(&p.cnt->st[ST_MISMATCH], 1ull);
  }
  if (make_leaf) {
    if (tid == 0) {
      const int32_t ls = atomicAdd(&p.cnt->n_leaves, 1);
      p.o.feat[node] = -1;
      p.o.child[node] = ls;
      p.o.cut[node] = NAN;
      p.o.tree[node] = tree;
      sh_misc[0] = ls;
    }
    __syncthreads();
    double *lv = p.o.leaf_vals + (int64_t)sh_misc[0] * lw;
    if (TASK == TASK_REG) {
      if (tid == 0) lv[0] = s.leaf_mean[i];
    } else {
      for (int c = tid; c < C; c += BEST_CTA) lv[c] = s.dist[(int64_t)i * C + c];  // pkg:960-964 / 913-927
    }
    return;
  }
  // ---- stable partition of the node's rows (pkg:1024-1039 / 841-856)
  const int32_t *idx = p.idx_src + base + b;
  const Col col = col_of(p, best_feature);
  auto goes_left = [&](int32_t j) {
    const double x = col_at(col, idx[j]);
    return (x < best_cut) || (best_mil && (x != x));
  };
  int32_t cnt = 0;
  for (int32_t j = tid; j < n; j += BEST_CTA) cnt += goes_left(j) ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) sh_cnt[wp] = cnt;
  __syncthreads();
  int32_t nl = 0;
  for (int w2 = 0; w2 < BEST_CTA / 32; w2++) nl += sh_cnt[w2];
  __syncthreads();
  int32_t lpos = b, rpos = b + nl;
  for (int32_t j0 = 0; j0 < n; j0 += BEST_CTA) {
    const int32_t j = j0 + tid;
    const bool has = j < n;
    const bool left = has && goes_left(j);
    const uint32_t lm = __ballot_sync(0xffffffffu, left), hm = __ballot_sync(0xffffffffu, has);
    const uint32_t rm = hm & ~lm;
    if (lane == 0) {
      sh_cnt[wp] = __popc(lm);
    }
    __syncthreads();
    int32_t lbefore = 0, ltot = 0, hbefore = 0;
    for (int w2 = 0; w2 < BEST_CTA / 32; w2++) {
      const int32_t c2 = sh_cnt[w2];
      if (w2 < wp) lbefore += c2;
      ltot += c2;
    }
    hbefore = min(wp * 32, max(0, n - j0));  // rows of the chunk in the warps before this one
    const int32_t chunk_rows = min(BEST_CTA, n - j0);
    if (has) {
      const uint32_t below = (1u << lane) - 1u;
      const int32_t dst = left ? lpos + lbefore + __popc(lm & below)
                               : rpos + (hbefore - lbefore) + __popc(rm & below);
      p.idx_dst[base + dst] = idx[j];
      if (TASK == TASK_REG) {
        p.yr_dst[base + dst] = p.yr_src[base + b + j];
      } else {
        p.yc_dst[base + dst] = p.yc_src[base + b + j];
        if (TASK == TASK_CLSW) p.w_dst[base + dst] = p.w_src[base + b + j];
      }
    }
    lpos += ltot;
    rpos += chunk_rows - ltot;
    __syncthreads();
  }
  if (tid == 0) {
    const int32_t slot = atomicAdd(&p.cnt->next_f, 2);
    sh_misc[1] = slot;
    const int32_t cl = p.node_base_next + slot;
    p.o.feat[node] = best_feature | (best_mil ? ET_MIL_BIT : 0);
    p.o.child[node] = cl;
    p.o.cut[node] = best_cut;
    p.o.tree[node] = tree;
#pragma unroll
    for (int side = 0; side < 2; side++) {
      const int32_t s2 = slot + side;
      p.nxt.tree[s2] = tree;
      p.nxt.begin[s2] = side ? b + nl : b;
      p.nxt.end[s2] = side ? e : b + nl;
      p.nxt.node[s2] = cl + side;
      // pkg:870 / 884 (sic): the regression right child keeps currentDepth; pkg:1055,1071: +1 both
      p.nxt.depth[s2] = (TASK == TASK_REG && side) ? depth : depth + 1;
      p.nxt.key[s2] = et_child_key(key, side);
      int64_t tc = -1;
      if (p.replay && tn >= 0) tc = side ? p.tr.right[tn] : p.tr.left[tn];
      p.nxt.trace[s2] = tc;
      p.q_nxt[Q_CTA][atomicAdd(&p.cnt->q_count[Q_CTA], 1)] = s2;
    }
    atomicAdd(&p.cnt->st[ST_PROWS], (unsigned long long)n);
  }
  __syncthreads();
  if (!p.replay) {
    const int32_t slot = sh_misc[1];
    uint32_t *ml2 = p.nxt.mask + (int64_t)slot * W, *mr = ml2 + W;
    for (int w = tid; w < W; w += BEST_CTA) {
      const uint32_t v = p.cur.mask[(int64_t)i * W + w];
      ml2[w] = v;
      mr[w] = v;
    }
  }
}

// ---- host side: buffers and the per-level sequence ---------------------------------------------------------
struct BestBufs {
  DevBuf<double> total, nsum, leaf_mean, best_score, best_cut, dist;
  DevBuf<int32_t> flags, visited, nconst, dc, tpos, best_feature, best_mil, ncand, nchunk, hist, cand_feat, cand_expect;
  DevBuf<long long> item_off;
  DevBuf<uint32_t> taken;
  DevBuf<unsigned long long> counters;
  DevBuf<BestItem> items;
  DevBuf<BestRes> res;
  void ensure(size_t F, int C, int W, size_t max_items) {
    total.ensure(F); nsum.ensure(F); leaf_mean.ensure(F); best_score.ensure(F); best_cut.ensure(F);
    dist.ensure(F * (size_t)std::max(C, 1));
    flags.ensure(F); visited.ensure(F); nconst.ensure(F); dc.ensure(F); tpos.ensure(F); best_feature.ensure(F);
    best_mil.ensure(F); ncand.ensure(F); nchunk.ensure(F);
    hist.ensure(F * (size_t)std::max(C, 1));
    cand_feat.ensure(F * 32); cand_expect.ensure(F * 32);
    item_off.ensure(F);
    taken.ensure(F * (size_t)W);
    counters.ensure(2);
    items.ensure(max_items);
    res.ensure(max_items);
  }
  BestState view() {
    BestState s;
    s.total = total.p; s.nsum = nsum.p; s.leaf_mean = leaf_mean.p; s.best_score = best_score.p; s.best_cut = best_cut.p;
    s.dist = dist.p; s.flags = flags.p; s.visited = visited.p; s.nconst = nconst.p; s.dc = dc.p; s.tpos = tpos.p;
    s.best_feature = best_feature.p; s.best_mil = best_mil.p; s.ncand = ncand.p; s.nchunk = nchunk.p; s.hist = hist.p;
    s.cand_feat = cand_feat.p; s.cand_expect = cand_expect.p; s.item_off = item_off.p; s.taken = taken.p;
    s.counters = counters.p;
    return s;
  }
};

// One level of the bestSplit builder: every open node sits in queue Q_CTA.  `rows` = samples of the batch (bounds
// the work items of a round: 32 candidates x (rows / 256 + nodes) chunks).
template <int TASK>
void launch_level_best(et_ctx *ctx, const P &p, int32_t count, BestBufs &bb, int64_t rows) {
  cudaStream_t st = ctx->stream;
  if (count <= 0) return;
  const size_t max_items = (size_t)32 * ((size_t)(rows / BEST_CTA) + (size_t)count + 1);
  bb.ensure((size_t)count * 2 + 16, p.C, p.W, max_items);  // (frontier indices of a level are < 2 * nodes of the level before)
  BestState s = bb.view();
  const size_t smem_prep = (size_t)std::max(p.C, 1) * sizeof(int32_t);
  const size_t smem_eval = (size_t)BEST_TILE * (2 * sizeof(double) + sizeof(int32_t)) + (size_t)std::max(p.C, 1) * sizeof(int32_t);
  k_best_prep<TASK><<<(unsigned)count, BEST_CTA, smem_prep, st>>>(p, s, count);
  ctx->launches++;
  for (;;) {
    CUDA_CHECK(cudaMemsetAsync(s.counters, 0, 2 * sizeof(unsigned long long), st));
    k_best_draw<<<(unsigned)ceil_div(count, 128), 128, 0, st>>>(p, s, count, bb.items.p);
    ctx->launches++;
    unsigned long long hcnt[2] = {0, 0};
    CUDA_CHECK(cudaMemcpyAsync(hcnt, s.counters, sizeof(hcnt), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    CUDA_CHECK(cudaGetLastError());
    if (hcnt[1] == 0) break;
    if (hcnt[0] > max_items) ET_FAIL(ET_ECUDA, "bestSplit: work list overflow (%llu items)", hcnt[0]);
    k_best_eval<TASK><<<(unsigned)hcnt[0], BEST_CTA, smem_eval, st>>>(p, s, bb.items.p, bb.res.p);
    k_best_consume<<<(unsigned)ceil_div(count, 128), 128, 0, st>>>(p, s, count, bb.res.p);
    ctx->launches += 2;
  }
  k_best_finish<TASK><<<(unsigned)count, BEST_CTA, 0, st>>>(p, s, count);
  ctx->launches++;
}

BestBufs *best_bufs_create() { return new BestBufs(); }
void best_bufs_destroy(BestBufs *bb) { delete bb; }

template void launch_level_best<TASK_CLS>(et_ctx *, const P &, int32_t, BestBufs &, int64_t);
template void launch_level_best<TASK_CLSW>(et_ctx *, const P &, int32_t, BestBufs &, int64_t);
template void launch_level_best<TASK_REG>(et_ctx *, const P &, int32_t, BestBufs &, int64_t);

}  // namespace etb
