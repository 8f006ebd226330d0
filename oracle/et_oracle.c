/*
 * et_oracle.c -- CPU ORACLE for lamp's extratrees path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is a plain-C restatement of the reference algorithm in
 *   extratrees/src/main/scala/lamp/forest/package.scala   (cited below as pkg:LINE)
 *   extratrees/src/main/scala/lamp/forest/extratrees.scala (adt:LINE)
 * plus the arithmetic of the un-vendored third-party dependency
 *   io.github.pityka::saddle-core 4.0.0-M11 (build.sbt:83,279):
 *   org.saddle.spire.random.rng.{Lcg64,Cmwc5}, Generator.nextInt/nextDouble,
 *   Vec[Double].sampleVariance / mean2 / sum2
 * whose published algorithm is restated from the spire PRNG sources it vendors.
 *
 * Parity pinning: the restatement is checked against every deterministic
 * known-answer test the reference holds for this path
 * (extratrees/src/test/scala/lamp/forest/extratree.test.scala:7-281,515-554) in
 * tests/test_oracle_kat.py.  Corners no reference test exercises ("parity
 * unpinned"): the rejection branch of nextInt(from,to), NaN handling in
 * mean2/sum2, forest-level output for a fixed seed.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product (libetgpu.so) never does.
 *
 * Build:  gcc -O2 -ffp-contract=off -fPIC -shared -pthread et_oracle.c -o libetoracle.so
 * -ffp-contract=off is REQUIRED: the JVM never fuses a*b+c.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* RNG: spire Lcg64 + Cmwc5 as vendored in saddle-core (org.saddle.spire.random) */
/* ------------------------------------------------------------------------- */
typedef struct {
  uint64_t x, y, z, w, v;
} eo_cmwc5;

/* Lcg64.nextLong: seed = 6364136223846793005*seed + 1442695040888963407 */
static inline uint64_t lcg64_next(uint64_t *seed) {
  *seed = 6364136223846793005ULL * (*seed) + 1442695040888963407ULL;
  return *seed;
}

/* Cmwc5.fromTime(t): five successive Lcg64(t).nextLong() -> x,y,z,w,v
 * (call sites pkg:629,655,720,740) */
void eo_cmwc5_from_time(eo_cmwc5 *g, int64_t time) {
  uint64_t s = (uint64_t)time;
  g->x = lcg64_next(&s);
  g->y = lcg64_next(&s);
  g->z = lcg64_next(&s);
  g->w = lcg64_next(&s);
  g->v = lcg64_next(&s);
}

/* Cmwc5.nextLong */
int64_t eo_cmwc5_next_long(eo_cmwc5 *g) {
  uint64_t t = g->x ^ (g->x >> 7);
  g->x = g->y;
  g->y = g->z;
  g->z = g->w;
  g->w = g->v;
  g->v = (g->v ^ (g->v << 6)) ^ (t ^ (t << 13));
  return (int64_t)((g->y + g->y + 1ULL) * g->v);
}

/* Generator.nextInt() = (nextLong() >>> 32).toInt */
int32_t eo_cmwc5_next_int(eo_cmwc5 *g) {
  return (int32_t)(((uint64_t)eo_cmwc5_next_long(g)) >> 32);
}

/* Generator.nextDouble() = (nextLong() >>> 11) * 2^-53 */
double eo_cmwc5_next_double(eo_cmwc5 *g) {
  return (double)(((uint64_t)eo_cmwc5_next_long(g)) >> 11) * 1.1102230246251565e-16;
}

/* Generator.nextDouble(from, until) = from + (until - from) * nextDouble()   (pkg:240,461) */
double eo_cmwc5_next_double_range(eo_cmwc5 *g, double a, double b) {
  double u = eo_cmwc5_next_double(g);
  double span = b - a;
  double prod = span * u;
  return a + prod;
}

static uint32_t retry_cap(uint32_t width) {
  uint32_t q = 0x80000000u / width;
  uint32_t r = 0x80000000u % width;
  uint32_t n = (q << 1) + (r << 1) / width;
  return n * width;
}

/* Generator.nextInt(from, to), both inclusive (pkg:86,233,325,454) */
int32_t eo_cmwc5_next_int_range(eo_cmwc5 *g, int32_t from, int32_t to) {
  uint32_t width = (uint32_t)to - (uint32_t)from + 1u;
  if (width == 0u) return eo_cmwc5_next_int(g);
  uint32_t cap = (width > 0x80000000u) ? width : retry_cap(width);
  if (cap == 0u) {
    uint32_t x = (uint32_t)eo_cmwc5_next_int(g);
    return (int32_t)((uint32_t)from + (x % width));
  }
  for (;;) {
    uint32_t x = (uint32_t)eo_cmwc5_next_int(g);
    if (x <= cap) return (int32_t)((x % width) + (uint32_t)from);
  }
}

/* ------------------------------------------------------------------------- */
/* statistics primitives                                                     */
/* ------------------------------------------------------------------------- */

/* saddle Vec[Double].sampleVariance: two-pass, sequential (pkg:308,437,1206,1210) */
double eo_sample_variance(const double *v, int64_t n) {
  if (n < 1) return NAN;
  if (n == 1) return 0.0;
  double s = 0.0;
  for (int64_t i = 0; i < n; i++) s += v[i];
  double m = s / (double)n;
  double q = 0.0;
  for (int64_t i = 0; i < n; i++) {
    double dlt = v[i] - m;
    q += dlt * dlt;
  }
  return q / ((double)n - 1.0);
}

/* population variance as written at pkg:307-308 / pkg:1206 */
static double pop_variance(const double *v, int64_t n) {
  return eo_sample_variance(v, n) * ((double)n - 1.0) / (double)n;
}

/* pkg:34-54 */
static void minmax(const double *v, int64_t n, double *mn, double *mx, int *has_missing) {
  double smin = 1.7976931348623157e308;
  double smax = -1.7976931348623157e308;
  int hm = 0;
  for (int64_t i = 0; i < n; i++) {
    double x = v[i];
    if (x < smin) smin = x;
    if (x > smax) smax = x;
    if (x != x && !hm) hm = 1;
  }
  *mn = smin;
  *mx = smax;
  *has_missing = hm;
}

/* pkg:10-32 */
static void less_than_cutpoint(const double *v, int64_t n, double cut, int missing_is_less,
                               uint8_t *out) {
  if (missing_is_less) {
    for (int64_t i = 0; i < n; i++) out[i] = (v[i] != v[i]) || (v[i] < cut);
  } else {
    for (int64_t i = 0; i < n; i++) out[i] = v[i] < cut;
  }
}

/* pkg:897-929 */
static void distribution(const int32_t *t, const double *w, int64_t n, int C, double *ar) {
  for (int j = 0; j < C; j++) ar[j] = 0.0;
  if (!w) {
    double s = (double)n;
    for (int64_t i = 0; i < n; i++) ar[t[i]] += 1.0 / s;
  } else {
    double s = 0.0;
    for (int64_t i = 0; i < n; i++) {
      ar[t[i]] += w[i];
      s += w[i];
    }
    for (int j = 0; j < C; j++) ar[j] /= s;
  }
}

/* pkg:1160-1180 */
double eo_gini_impurity(const int32_t *t, const double *w, int64_t n, int32_t C) {
  double *p = (double *)malloc(sizeof(double) * (size_t)(C > 0 ? C : 1));
  distribution(t, w, n, C, p);
  double s = 0.0;
  for (int j = 0; j < C; j++) s += p[j] * p[j];
  free(p);
  return 1.0 - s;
}

/* pkg:1181-1194 (also zeroes the buffer, like the reference) */
static double gini_from_distribution(double *dist, int C) {
  double s = 0.0;
  for (int i = 0; i < C; i++) {
    double k = dist[i];
    dist[i] = 0.0;
    s += k * k;
  }
  return 1.0 - s;
}

/* pkg:1101-1158.  buf1/buf2 must be zero on entry (they are re-zeroed on exit). */
static double gini_score(const int32_t *t, const double *w, const uint8_t *mask, int64_t n,
                         double g_nosplit, int C, double *buf1, double *buf2) {
  double n_nosplit;
  if (!w) {
    n_nosplit = (double)n;
  } else {
    n_nosplit = 0.0; /* sum2 */
    for (int64_t i = 0; i < n; i++) n_nosplit += w[i];
  }
  double cin = 0.0, cout = 0.0;
  if (!w) {
    for (int64_t i = 0; i < n; i++) {
      if (mask[i]) {
        cin += 1.0;
        buf1[t[i]] += 1.0;
      } else {
        cout += 1.0;
        buf2[t[i]] += 1.0;
      }
    }
  } else {
    for (int64_t i = 0; i < n; i++) {
      double ww = w[i];
      if (mask[i]) {
        cin += ww;
        buf1[t[i]] += ww;
      } else {
        cout += ww;
        buf2[t[i]] += ww;
      }
    }
  }
  for (int i = 0; i < C; i++) {
    buf1[i] /= cin;
    buf2[i] /= cout;
  }
  double gin = gini_from_distribution(buf1, C);
  double gout = gini_from_distribution(buf2, C);
  return g_nosplit - gin * cin / n_nosplit - gout * cout / n_nosplit;
}

double eo_gini_score(const int32_t *t, const double *w, const uint8_t *mask, int64_t n,
                     double g_nosplit, int32_t C) {
  double *b = (double *)calloc((size_t)(2 * (C > 0 ? C : 1)), sizeof(double));
  double r = gini_score(t, w, mask, n, g_nosplit, C, b, b + C);
  free(b);
  return r;
}

/* pkg:1196-1218 (partition pkg:1084-1099 is order preserving) */
static double variance_reduction(const double *t, const uint8_t *mask, int64_t n, double var_nosplit,
                                 double *buf_in, double *buf_out) {
  int64_t nin = 0, nout = 0;
  for (int64_t i = 0; i < n; i++) {
    if (mask[i])
      buf_in[nin++] = t[i];
    else
      buf_out[nout++] = t[i];
  }
  double vin = (nin == 1) ? 0.0 : pop_variance(buf_in, nin);
  double vout = (nout == 1) ? 0.0 : pop_variance(buf_out, nout);
  double nn = (double)n;
  return (var_nosplit - ((double)nin / nn) * vin - ((double)nout / nn) * vout) / var_nosplit;
}

double eo_variance_reduction(const double *t, const uint8_t *mask, int64_t n, double var_nosplit) {
  double *b = (double *)malloc(sizeof(double) * (size_t)(2 * n + 2));
  double r = variance_reduction(t, mask, n, var_nosplit, b, b + n + 1);
  free(b);
  return r;
}

double eo_pop_variance(const double *t, int64_t n) { return pop_variance(t, n); }

/* ------------------------------------------------------------------------- */
/* tree storage (pre-order) + replay trace + counters                        */
/* ------------------------------------------------------------------------- */
typedef struct {
  int64_t v_mm;    /* (sample,feature) visits of min/max, incl. constant hits */
  int64_t v_sc;    /* (sample,feature) visits of scoring                       */
  int64_t s_rows;  /* sum of n over nodes that reach split search              */
  int64_t p_rows;  /* sum of n over nodes actually split                       */
  int64_t draws;   /* nextInt(low,high-1) calls                                */
  int64_t const_hits;
  int64_t scored;
  int64_t nodes;
} eo_stats;

typedef struct {
  int32_t n_nodes, cap_nodes;
  int32_t *feature; /* -1 = leaf */
  double *cut;
  uint8_t *mil;
  int32_t *left, *right;
  double *leaf; /* n_nodes x leaf_width, meaningful for leaves */
  /* replay trace */
  int64_t *cand_begin; /* per node; cand_end = next node's begin (filled at export) */
  int32_t *cand_count;
  int64_t n_cand, cap_cand;
  int32_t *cand_feature;
  double *cand_u;      /* raw nextDouble() uniform; NaN when none was drawn */
  uint8_t *cand_flag;  /* 0 = constant, 1 = scored, 2 = scored NaN (treated as constant) */
  eo_stats st;
} eo_tree;

typedef struct eo_forest {
  int32_t m, leaf_width, is_regression, record_trace;
  eo_tree *trees;
  int64_t next_long_after; /* rng.nextLong() after the build when parallelism<=1 */
} eo_forest;

static int32_t tree_new_node(eo_tree *t, int lw, int trace) {
  if (t->n_nodes == t->cap_nodes) {
    int32_t nc = t->cap_nodes ? t->cap_nodes * 2 : 64;
    t->feature = (int32_t *)realloc(t->feature, sizeof(int32_t) * (size_t)nc);
    t->cut = (double *)realloc(t->cut, sizeof(double) * (size_t)nc);
    t->mil = (uint8_t *)realloc(t->mil, (size_t)nc);
    t->left = (int32_t *)realloc(t->left, sizeof(int32_t) * (size_t)nc);
    t->right = (int32_t *)realloc(t->right, sizeof(int32_t) * (size_t)nc);
    t->leaf = (double *)realloc(t->leaf, sizeof(double) * (size_t)nc * (size_t)lw);
    if (trace) {
      t->cand_begin = (int64_t *)realloc(t->cand_begin, sizeof(int64_t) * (size_t)nc);
      t->cand_count = (int32_t *)realloc(t->cand_count, sizeof(int32_t) * (size_t)nc);
    }
    t->cap_nodes = nc;
  }
  int32_t id = t->n_nodes++;
  t->feature[id] = -1;
  t->cut[id] = NAN;
  t->mil[id] = 0;
  t->left[id] = -1;
  t->right[id] = -1;
  for (int j = 0; j < lw; j++) t->leaf[(size_t)id * lw + j] = 0.0;
  if (trace) {
    t->cand_begin[id] = t->n_cand;
    t->cand_count[id] = 0;
  }
  return id;
}

static void tree_push_cand(eo_tree *t, int32_t node, int32_t feature, double u, uint8_t flag) {
  if (t->n_cand == t->cap_cand) {
    int64_t nc = t->cap_cand ? t->cap_cand * 2 : 256;
    t->cand_feature = (int32_t *)realloc(t->cand_feature, sizeof(int32_t) * (size_t)nc);
    t->cand_u = (double *)realloc(t->cand_u, sizeof(double) * (size_t)nc);
    t->cand_flag = (uint8_t *)realloc(t->cand_flag, (size_t)nc);
    t->cap_cand = nc;
  }
  t->cand_feature[t->n_cand] = feature;
  t->cand_u[t->n_cand] = u;
  t->cand_flag[t->n_cand] = flag;
  t->n_cand++;
  t->cand_count[node]++;
}

static void tree_free(eo_tree *t) {
  free(t->feature);
  free(t->cut);
  free(t->mil);
  free(t->left);
  free(t->right);
  free(t->leaf);
  free(t->cand_begin);
  free(t->cand_count);
  free(t->cand_feature);
  free(t->cand_u);
  free(t->cand_flag);
  memset(t, 0, sizeof(*t));
}

/* ------------------------------------------------------------------------- */
/* split search                                                              */
/* ------------------------------------------------------------------------- */
typedef struct {
  /* immutable problem */
  const double *x; /* row-major n_rows x d  (saddle Mat layout, pkg:936) */
  int64_t n_rows;
  int32_t d;
  const int32_t *y_cls;
  const double *y_reg;
  const double *w;
  int32_t C, n_min, k, best_split, max_depth;
  int record_trace;
  /* per-tree scratch, sized n_rows */
  double *col;
  uint8_t *mask_lt, *mask_nanlt;
  double *buf_in, *buf_out;
  double *hb1, *hb2;
  eo_cmwc5 *rng;
  eo_tree *tree;
  int32_t cur_node; /* node whose candidates are being recorded */
} eo_ctx;

/* takeCol, pkg:931-937: strided gather through the row-major matrix */
static void take_col(const eo_ctx *c, const int32_t *subset, int64_t n, int32_t attr, double *out) {
  const double *x = c->x;
  int64_t d = c->d;
  for (int64_t i = 0; i < n; i++) out[i] = x[(int64_t)subset[i] * d + attr];
}

typedef struct {
  int32_t feature;
  double cut;
  int32_t num_constant;
  int32_t mil;
} eo_split;

/* score of one cutpoint: returns chosen score and sets *mil (pkg:243-275 / 463-487) */
static double score_cut(eo_ctx *c, const double *col, int64_t n, double cut, int has_missing,
                        const int32_t *t_cls, const double *t_reg, const double *w, double total,
                        int *mil_out) {
  double s_l = NAN, s_n;
  if (has_missing) {
    less_than_cutpoint(col, n, cut, 1, c->mask_nanlt);
    s_l = t_cls ? gini_score(t_cls, w, c->mask_nanlt, n, total, c->C, c->hb1, c->hb2)
                : variance_reduction(t_reg, c->mask_nanlt, n, total, c->buf_in, c->buf_out);
  }
  less_than_cutpoint(col, n, cut, 0, c->mask_lt);
  s_n = t_cls ? gini_score(t_cls, w, c->mask_lt, n, total, c->C, c->hb1, c->hb2)
              : variance_reduction(t_reg, c->mask_lt, n, total, c->buf_in, c->buf_out);
  int mil = (!(s_l != s_l)) && (s_l > s_n || (s_n != s_n));
  *mil_out = mil;
  return mil ? s_l : s_n;
}

/* the four split functions share one loop: pkg:56-202, 203-297, 298-426, 427-511 */
static eo_split split_generic(eo_ctx *c, const int32_t *subset, int64_t n, int32_t *attributes,
                              int32_t num_constant, const int32_t *t_cls, const double *t_reg,
                              const double *w) {
  double total = t_cls ? eo_gini_impurity(t_cls, w, n, c->C) : pop_variance(t_reg, n);
  int32_t low = num_constant, high = c->d, N = c->d;
  double best_score = -INFINITY;
  int32_t best_feature = -1;
  double best_cut = NAN;
  int best_mil = 0;
  int32_t visited = 0;
  eo_stats *st = &c->tree->st;
  st->s_rows += n;
  while (N - high < c->k && high - low > 0) {
    int32_t r = eo_cmwc5_next_int_range(c->rng, low, high - 1);
    int32_t attr = attributes[r];
    st->draws++;
    double mn, mx;
    int has_missing;
    take_col(c, subset, n, attr, c->col);
    minmax(c->col, n, &mn, &mx, &has_missing);
    st->v_mm += n;
    if (mx <= mn && !has_missing) {
      int32_t tmp = attributes[r];
      attributes[r] = attributes[low];
      attributes[low] = tmp;
      low++;
      st->const_hits++;
      if (c->record_trace) tree_push_cand(c->tree, c->cur_node, attr, NAN, 0);
    } else {
      double cut, u = NAN;
      int mil = 0;
      double chosen;
      take_col(c, subset, n, attr, c->col); /* second gather, pkg:242 */
      if (!c->best_split) {
        u = eo_cmwc5_next_double(c->rng);
        double span = mx - mn;
        double prod = span * u;
        cut = mn + prod; /* Generator.nextDouble(min,max) */
      } else {
        /* pkg:133-145: every sample value tried, strict >, first wins, default index 0 */
        double smax = -INFINITY;
        int64_t maxi = 0;
        for (int64_t i = 0; i < n; i++) {
          int m2;
          double s = score_cut(c, c->col, n, c->col[i], has_missing, t_cls, t_reg, w, total, &m2);
          st->v_sc += n;
          if (s > smax) {
            smax = s;
            maxi = i;
          }
        }
        cut = c->col[maxi];
      }
      chosen = score_cut(c, c->col, n, cut, has_missing, t_cls, t_reg, w, total, &mil);
      st->v_sc += n;
      st->scored++;
      if (chosen > best_score) {
        best_score = chosen;
        best_feature = attr;
        best_cut = cut;
        best_mil = mil;
      }
      if (chosen != chosen) {
        int32_t tmp = attributes[r];
        attributes[r] = attributes[low];
        attributes[low] = tmp;
        low++;
        if (c->record_trace) tree_push_cand(c->tree, c->cur_node, attr, u, 2);
      } else {
        visited++;
        int32_t tmp = attributes[r];
        attributes[r] = attributes[high - 1];
        attributes[high - 1] = tmp;
        high--;
        if (c->record_trace) tree_push_cand(c->tree, c->cur_node, attr, u, 1);
      }
    }
  }
  eo_split out;
  out.cut = best_cut;
  out.num_constant = low;
  out.mil = best_mil;
  out.feature = (visited == 0 || best_cut != best_cut) ? -1 : best_feature;
  return out;
}

/* ------------------------------------------------------------------------- */
/* tree builders                                                             */
/* ------------------------------------------------------------------------- */

/* pkg:943-1082 */
static int32_t build_tree_cls(eo_ctx *c, const int32_t *subset, int64_t n, int32_t *attributes,
                              int32_t num_constant, int32_t depth) {
  eo_tree *t = c->tree;
  int lw = c->C;
  int32_t id = tree_new_node(t, lw, c->record_trace);
  int32_t *tin = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
  double *win = c->w ? (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1)) : NULL;
  for (int64_t i = 0; i < n; i++) tin[i] = c->y_cls[subset[i]];
  if (win)
    for (int64_t i = 0; i < n; i++) win[i] = c->w[subset[i]];
  int make_leaf = 0;
  if (c->n_rows < c->n_min || depth >= c->max_depth) { /* pkg:993 (whole-table rows, sic) */
    make_leaf = 1;
  } else {
    int uniform = 1;
    for (int64_t i = 1; i < n && uniform; i++)
      if (tin[i] != tin[0]) uniform = 0;
    if (uniform) make_leaf = 1;
  }
  eo_split sp;
  sp.feature = -1;
  if (!make_leaf) {
    c->cur_node = id;
    sp = split_generic(c, subset, n, attributes, num_constant, tin, NULL, win);
    if (sp.feature < 0) make_leaf = 1;
  }
  if (make_leaf) {
    distribution(tin, win, n, c->C, &t->leaf[(size_t)id * lw]);
    free(tin);
    free(win);
    return id;
  }
  free(win);
  free(tin);
  /* pkg:1024-1039: two order-preserving filters */
  int32_t *ls = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
  int32_t *rs = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
  int64_t nl = 0, nr = 0;
  const double *x = c->x;
  int64_t d = c->d;
  for (int64_t i = 0; i < n; i++) {
    double v = x[(int64_t)subset[i] * d + sp.feature];
    int isnan_ = v != v;
    if (sp.mil) {
      if (v < sp.cut || isnan_) ls[nl++] = subset[i];
      if (v >= sp.cut) rs[nr++] = subset[i];
    } else {
      if (v < sp.cut) ls[nl++] = subset[i];
      if (v >= sp.cut || isnan_) rs[nr++] = subset[i];
    }
  }
  t->st.p_rows += n;
  int32_t l = build_tree_cls(c, ls, nl, attributes, sp.num_constant, depth + 1);
  free(ls);
  int32_t r = build_tree_cls(c, rs, nr, attributes, sp.num_constant, depth + 1);
  free(rs);
  t = c->tree;
  t->feature[id] = sp.feature;
  t->cut[id] = sp.cut;
  t->mil[id] = (uint8_t)sp.mil;
  t->left[id] = l;
  t->right[id] = r;
  return id;
}

/* pkg:766-895 */
static int32_t build_tree_reg(eo_ctx *c, const int32_t *subset, int64_t n, int32_t *attributes,
                              int32_t num_constant, int32_t depth) {
  eo_tree *t = c->tree;
  int32_t id = tree_new_node(t, 1, c->record_trace);
  double *tin = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  for (int64_t i = 0; i < n; i++) tin[i] = c->y_reg[subset[i]];
  int make_leaf = 0;
  if (n < c->n_min || depth >= c->max_depth) { /* pkg:813 */
    make_leaf = 1;
  } else {
    int uniform = 1;
    for (int64_t i = 1; i < n && uniform; i++)
      if (tin[i] != tin[0]) uniform = 0;
    if (uniform) make_leaf = 1;
  }
  eo_split sp;
  sp.feature = -1;
  if (!make_leaf) {
    c->cur_node = id;
    sp = split_generic(c, subset, n, attributes, num_constant, NULL, tin, NULL);
    if (sp.feature == -1) make_leaf = 1;
  }
  if (make_leaf) {
    double s = 0.0; /* mean2, pkg:782 */
    for (int64_t i = 0; i < n; i++) s += tin[i];
    t->leaf[id] = s / (double)n;
    free(tin);
    return id;
  }
  free(tin);
  int32_t *ls = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
  int32_t *rs = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
  int64_t nl = 0, nr = 0;
  const double *x = c->x;
  int64_t d = c->d;
  for (int64_t i = 0; i < n; i++) {
    double v = x[(int64_t)subset[i] * d + sp.feature];
    int isnan_ = v != v;
    if (sp.mil) {
      if (v < sp.cut || isnan_) ls[nl++] = subset[i];
      if (v >= sp.cut) rs[nr++] = subset[i];
    } else {
      if (v < sp.cut) ls[nl++] = subset[i];
      if (v >= sp.cut || isnan_) rs[nr++] = subset[i];
    }
  }
  t->st.p_rows += n;
  int32_t l = build_tree_reg(c, ls, nl, attributes, sp.num_constant, depth + 1); /* pkg:870 */
  free(ls);
  int32_t r = build_tree_reg(c, rs, nr, attributes, sp.num_constant, depth); /* pkg:884 (sic) */
  free(rs);
  t = c->tree;
  t->feature[id] = sp.feature;
  t->cut[id] = sp.cut;
  t->mil[id] = (uint8_t)sp.mil;
  t->left[id] = l;
  t->right[id] = r;
  return id;
}

/* ------------------------------------------------------------------------- */
/* forest builders (pkg:611-681, 704-764)                                    */
/* ------------------------------------------------------------------------- */
typedef struct {
  eo_ctx proto;
  eo_forest *forest;
  eo_cmwc5 *rngs; /* one per tree (parallelism > 1) */
  int32_t next_tree;
  pthread_mutex_t mu;
} eo_job;

static void ctx_alloc_scratch(eo_ctx *c) {
  size_t n = (size_t)(c->n_rows > 0 ? c->n_rows : 1);
  c->col = (double *)malloc(sizeof(double) * n);
  c->mask_lt = (uint8_t *)malloc(n);
  c->mask_nanlt = (uint8_t *)malloc(n);
  c->buf_in = (double *)malloc(sizeof(double) * (n + 1));
  c->buf_out = (double *)malloc(sizeof(double) * (n + 1));
  c->hb1 = (double *)calloc((size_t)(c->C > 0 ? c->C : 1), sizeof(double));
  c->hb2 = (double *)calloc((size_t)(c->C > 0 ? c->C : 1), sizeof(double));
}
static void ctx_free_scratch(eo_ctx *c) {
  free(c->col);
  free(c->mask_lt);
  free(c->mask_nanlt);
  free(c->buf_in);
  free(c->buf_out);
  free(c->hb1);
  free(c->hb2);
}

static void build_one_tree(eo_ctx *c, eo_tree *tree, eo_cmwc5 *rng) {
  c->tree = tree;
  c->rng = rng;
  int32_t *attributes = (int32_t *)malloc(sizeof(int32_t) * (size_t)(c->d > 0 ? c->d : 1));
  for (int32_t j = 0; j < c->d; j++) attributes[j] = j; /* array.range(0, numCols) */
  int32_t *subset = (int32_t *)malloc(sizeof(int32_t) * (size_t)(c->n_rows > 0 ? c->n_rows : 1));
  for (int64_t i = 0; i < c->n_rows; i++) subset[i] = (int32_t)i;
  if (c->y_cls)
    build_tree_cls(c, subset, c->n_rows, attributes, 0, 0);
  else
    build_tree_reg(c, subset, c->n_rows, attributes, 0, 0);
  tree->st.nodes = tree->n_nodes;
  free(subset);
  free(attributes);
}

static void *worker(void *arg) {
  eo_job *job = (eo_job *)arg;
  eo_ctx c = job->proto;
  ctx_alloc_scratch(&c);
  for (;;) {
    pthread_mutex_lock(&job->mu);
    int32_t t = job->next_tree++;
    pthread_mutex_unlock(&job->mu);
    if (t >= job->forest->m) break;
    build_one_tree(&c, &job->forest->trees[t], &job->rngs[t]);
  }
  ctx_free_scratch(&c);
  return NULL;
}

typedef struct {
  eo_ctx *c;
  eo_forest *f;
  eo_cmwc5 *rng;
} eo_serial_job;

static void *serial_worker(void *arg) {
  eo_serial_job *j = (eo_serial_job *)arg;
  ctx_alloc_scratch(j->c);
  for (int32_t t = 0; t < j->f->m; t++) build_one_tree(j->c, &j->f->trees[t], j->rng);
  ctx_free_scratch(j->c);
  return NULL;
}

static eo_forest *build_forest(eo_ctx proto, int32_t m, int32_t parallelism, int64_t seed,
                               int32_t n_threads) {
  eo_forest *f = (eo_forest *)calloc(1, sizeof(eo_forest));
  f->m = m;
  f->is_regression = proto.y_reg != NULL;
  f->leaf_width = f->is_regression ? 1 : proto.C;
  f->record_trace = proto.record_trace;
  f->trees = (eo_tree *)calloc((size_t)(m > 0 ? m : 1), sizeof(eo_tree));
  eo_cmwc5 rng;
  eo_cmwc5_from_time(&rng, seed); /* pkg:629,720 */
  pthread_attr_t attr;
  pthread_attr_init(&attr);
  pthread_attr_setstacksize(&attr, (size_t)1 << 30); /* recursion depth = tree depth */
  if (parallelism <= 1) {
    /* pkg:634-651: all trees draw from ONE stream, in order */
    eo_serial_job sj;
    sj.c = &proto;
    sj.f = f;
    sj.rng = &rng;
    pthread_t th;
    pthread_create(&th, &attr, serial_worker, &sj);
    pthread_join(th, NULL);
    f->next_long_after = eo_cmwc5_next_long(&rng);
  } else {
    /* pkg:654-655: rng_t = Cmwc5.fromTime(rng.nextLong()), drawn sequentially */
    eo_job job;
    memset(&job, 0, sizeof(job));
    job.proto = proto;
    job.forest = f;
    job.rngs = (eo_cmwc5 *)malloc(sizeof(eo_cmwc5) * (size_t)(m > 0 ? m : 1));
    for (int32_t t = 0; t < m; t++) eo_cmwc5_from_time(&job.rngs[t], eo_cmwc5_next_long(&rng));
    pthread_mutex_init(&job.mu, NULL);
    int nt = n_threads > 0 ? n_threads : parallelism;
    if (nt > m) nt = m > 0 ? m : 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nt);
    for (int i = 0; i < nt; i++) pthread_create(&th[i], &attr, worker, &job);
    for (int i = 0; i < nt; i++) pthread_join(th[i], NULL);
    free(th);
    free(job.rngs);
    pthread_mutex_destroy(&job.mu);
    f->next_long_after = 0;
  }
  pthread_attr_destroy(&attr);
  return f;
}

/* returns NULL on the reference's require() failures (pkg:624-633) -- caller maps to
 * IllegalArgumentException.  n_rows_target lets the test exercise the length check. */
eo_forest *eo_build_classification(const double *x, int64_t n, int32_t d, const int32_t *y,
                                   int64_t n_target, const double *w, int32_t C, int32_t n_min,
                                   int32_t k, int32_t m, int32_t parallelism, int32_t best_split,
                                   int32_t max_depth, int64_t seed, int32_t record_trace,
                                   int32_t n_threads) {
  if (n != n_target) return NULL;
  if (w)
    for (int64_t i = 0; i < n; i++)
      if (w[i] < 0.0) return NULL;
  if (n <= 0 && m > 0) return NULL; /* targetIsConstant reads col.raw(0): AIOOBE in the reference */
  eo_ctx c;
  memset(&c, 0, sizeof(c));
  c.x = x;
  c.n_rows = n;
  c.d = d;
  c.y_cls = y;
  c.w = w;
  c.C = C;
  c.n_min = n_min;
  c.k = k;
  c.best_split = best_split;
  c.max_depth = max_depth;
  c.record_trace = record_trace;
  return build_forest(c, m, parallelism, seed, n_threads);
}

eo_forest *eo_build_regression(const double *x, int64_t n, int32_t d, const double *y,
                               int64_t n_target, int32_t n_min, int32_t k, int32_t m,
                               int32_t parallelism, int32_t best_split, int32_t max_depth,
                               int64_t seed, int32_t record_trace, int32_t n_threads) {
  if (n != n_target) return NULL;
  if (n <= 0 && m > 0) return NULL; /* require(subset.length > 0), pkg:779 */
  eo_ctx c;
  memset(&c, 0, sizeof(c));
  c.x = x;
  c.n_rows = n;
  c.d = d;
  c.y_reg = y;
  c.C = 1;
  c.n_min = n_min;
  c.k = k;
  c.best_split = best_split;
  c.max_depth = max_depth;
  c.record_trace = record_trace;
  return build_forest(c, m, parallelism, seed, n_threads);
}

void eo_forest_free(eo_forest *f) {
  if (!f) return;
  for (int32_t t = 0; t < f->m; t++) tree_free(&f->trees[t]);
  free(f->trees);
  free(f);
}

int32_t eo_forest_num_trees(const eo_forest *f) { return f->m; }
int32_t eo_forest_leaf_width(const eo_forest *f) { return f->leaf_width; }
int64_t eo_forest_next_long_after(const eo_forest *f) { return f->next_long_after; }
int32_t eo_tree_size(const eo_forest *f, int32_t t) { return f->trees[t].n_nodes; }
int64_t eo_tree_trace_size(const eo_forest *f, int32_t t) { return f->trees[t].n_cand; }

void eo_tree_export(const eo_forest *f, int32_t ti, int32_t *feature, double *cut, uint8_t *mil,
                    int32_t *left, int32_t *right, double *leaf) {
  const eo_tree *t = &f->trees[ti];
  size_t n = (size_t)t->n_nodes;
  memcpy(feature, t->feature, sizeof(int32_t) * n);
  memcpy(cut, t->cut, sizeof(double) * n);
  memcpy(mil, t->mil, n);
  memcpy(left, t->left, sizeof(int32_t) * n);
  memcpy(right, t->right, sizeof(int32_t) * n);
  memcpy(leaf, t->leaf, sizeof(double) * n * (size_t)f->leaf_width);
}

/* replay trace: cand_begin has n_nodes+1 entries (node i owns [begin[i], begin[i+1]) is NOT
 * guaranteed because candidates of a node are contiguous but nodes interleave in DFS order;
 * so both begin and count are exported). */
void eo_tree_trace_export(const eo_forest *f, int32_t ti, int64_t *cand_begin, int32_t *cand_count,
                          int32_t *cand_feature, double *cand_u, uint8_t *cand_flag) {
  const eo_tree *t = &f->trees[ti];
  size_t n = (size_t)t->n_nodes, c = (size_t)t->n_cand;
  memcpy(cand_begin, t->cand_begin, sizeof(int64_t) * n);
  memcpy(cand_count, t->cand_count, sizeof(int32_t) * n);
  memcpy(cand_feature, t->cand_feature, sizeof(int32_t) * c);
  memcpy(cand_u, t->cand_u, sizeof(double) * c);
  memcpy(cand_flag, t->cand_flag, c);
}

/* out[8]: v_mm, v_sc, s_rows, p_rows, draws, const_hits, scored, nodes (summed over trees) */
void eo_forest_stats(const eo_forest *f, int64_t *out) {
  for (int i = 0; i < 8; i++) out[i] = 0;
  for (int32_t t = 0; t < f->m; t++) {
    const eo_stats *s = &f->trees[t].st;
    out[0] += s->v_mm;
    out[1] += s->v_sc;
    out[2] += s->s_rows;
    out[3] += s->p_rows;
    out[4] += s->draws;
    out[5] += s->const_hits;
    out[6] += s->scored;
    out[7] += s->nodes;
  }
}

/* import flat pre-order trees (for predicting with forests produced elsewhere) */
eo_forest *eo_forest_import(int32_t m, int32_t leaf_width, int32_t is_regression,
                            const int32_t *tree_sizes, const int32_t *feature, const double *cut,
                            const uint8_t *mil, const int32_t *left, const int32_t *right,
                            const double *leaf) {
  eo_forest *f = (eo_forest *)calloc(1, sizeof(eo_forest));
  f->m = m;
  f->leaf_width = leaf_width;
  f->is_regression = is_regression;
  f->trees = (eo_tree *)calloc((size_t)(m > 0 ? m : 1), sizeof(eo_tree));
  size_t off = 0;
  for (int32_t t = 0; t < m; t++) {
    eo_tree *tr = &f->trees[t];
    size_t n = (size_t)tree_sizes[t];
    tr->n_nodes = tr->cap_nodes = (int32_t)n;
    tr->feature = (int32_t *)malloc(sizeof(int32_t) * (n ? n : 1));
    tr->cut = (double *)malloc(sizeof(double) * (n ? n : 1));
    tr->mil = (uint8_t *)malloc(n ? n : 1);
    tr->left = (int32_t *)malloc(sizeof(int32_t) * (n ? n : 1));
    tr->right = (int32_t *)malloc(sizeof(int32_t) * (n ? n : 1));
    tr->leaf = (double *)malloc(sizeof(double) * (n ? n : 1) * (size_t)leaf_width);
    memcpy(tr->feature, feature + off, sizeof(int32_t) * n);
    memcpy(tr->cut, cut + off, sizeof(double) * n);
    memcpy(tr->mil, mil + off, n);
    memcpy(tr->left, left + off, sizeof(int32_t) * n);
    memcpy(tr->right, right + off, sizeof(int32_t) * n);
    memcpy(tr->leaf, leaf + off * (size_t)leaf_width, sizeof(double) * n * (size_t)leaf_width);
    off += n;
  }
  return f;
}

/* ------------------------------------------------------------------------- */
/* predict (pkg:513-586)                                                     */
/* ------------------------------------------------------------------------- */
static int32_t traverse(const eo_tree *t, const double *sample) {
  int32_t id = 0;
  while (t->feature[id] >= 0) {
    double v = sample[t->feature[id]];
    if (v < t->cut[id] || (t->mil[id] && v != v))
      id = t->left[id];
    else
      id = t->right[id];
  }
  return id;
}

/* out: n x C row-major.  Per class: sequential mean over trees in tree order (pkg:547-549). */
void eo_predict_classification(const eo_forest *f, const double *x, int64_t n, int32_t d,
                               double *out) {
  int C = f->leaf_width;
  int32_t *leaf_of = (int32_t *)malloc(sizeof(int32_t) * (size_t)(f->m > 0 ? f->m : 1));
  for (int64_t i = 0; i < n; i++) {
    const double *s = x + i * (int64_t)d;
    for (int32_t t = 0; t < f->m; t++) leaf_of[t] = traverse(&f->trees[t], s);
    for (int c = 0; c < C; c++) {
      double acc = 0.0;
      for (int32_t t = 0; t < f->m; t++) acc += f->trees[t].leaf[(size_t)leaf_of[t] * C + c];
      out[i * C + c] = acc / (double)f->m;
    }
  }
  free(leaf_of);
}

void eo_predict_regression(const eo_forest *f, const double *x, int64_t n, int32_t d, double *out) {
  for (int64_t i = 0; i < n; i++) {
    const double *s = x + i * (int64_t)d;
    double acc = 0.0;
    for (int32_t t = 0; t < f->m; t++) {
      const eo_tree *tr = &f->trees[t];
      acc += tr->leaf[traverse(tr, s)];
    }
    out[i] = acc / (double)f->m;
  }
}

/* ------------------------------------------------------------------------- */
/* direct entry points to the split functions for the known-answer tests     */
/* (tst:76-281 call them with explicit subset / attributes / numConstant)    */
/* ------------------------------------------------------------------------- */
void eo_split_kat(const double *x, int64_t n_rows, int32_t d, const int32_t *subset, int64_t n_sub,
              int32_t *attributes, int32_t num_constant, int32_t k, const int32_t *t_cls,
              const double *t_reg, const double *w_at_subset, int32_t C, int32_t best_split,
              int64_t rng_seed, int32_t *out_feature, double *out_cut, int32_t *out_num_constant,
              int32_t *out_mil) {
  eo_ctx c;
  memset(&c, 0, sizeof(c));
  c.x = x;
  c.n_rows = n_rows;
  c.d = d;
  c.C = C > 0 ? C : 1;
  c.k = k;
  c.best_split = best_split;
  eo_cmwc5 rng;
  eo_cmwc5_from_time(&rng, rng_seed);
  c.rng = &rng;
  eo_tree tree;
  memset(&tree, 0, sizeof(tree));
  c.tree = &tree;
  ctx_alloc_scratch(&c);
  eo_split sp = split_generic(&c, subset, n_sub, attributes, num_constant, t_cls, t_reg, w_at_subset);
  ctx_free_scratch(&c);
  *out_feature = sp.feature;
  *out_cut = sp.cut;
  *out_num_constant = sp.num_constant;
  *out_mil = sp.mil;
}
