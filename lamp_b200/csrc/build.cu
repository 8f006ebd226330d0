// build.cu -- level-wise extratrees builder over all open nodes of all trees of a batch.
//
// Replaces the recursive JVM builder buildTreeClassification (pkg:943-1082) / buildTreeRegression
// (pkg:766-895) and the split search splitClassification (pkg:203-297) / splitRegression
// (pkg:427-511) of extratrees/src/main/scala/lamp/forest/package.scala.
//
// Data layout in HBM
//   X         column-major FP64 [d][ld]            (the JVM walks a row-major matrix with stride d)
//   idx       int32 [B][n] x2 (ping-pong)          sample rows of every open node, ascending inside a
//                                                  node segment (the reference's filter keeps order)
//   yc/yr/w   labels / targets / weights permuted alongside idx so a node's segment streams
//   frontier  SoA of the open nodes of one level (tree, begin, end, node id, depth, RNG key | trace
//             node, class histogram, known-constant feature bitmask), x2 (this level / next level)
//   queues    frontier indices bucketed by node size: small nodes -> one warp per node,
//             large nodes -> one CTA per node
//   pool      output nodes in creation (level) order + compact leaf-value pool; converted on the
//             device to per-tree pre-order 16-byte nodes (the layout predict traverses)
//
// One level = one launch per size class.  A team (warp or CTA) owns a node end to end: stop rules,
// split search (batches of candidates: gather + min/max + cutpoint + side histogram with the whole
// team on the samples of one candidate, then one thread per candidate for the exact score), first-
// best argmax, stable partition, children.  The host only reads the next level's queue sizes.
//
// Exactness: every floating-point expression of the reference is evaluated with individually
// rounded _rn operations in the reference's order.  Unweighted classification reduces integer
// class histograms in parallel (exact) and evaluates the Gini expressions in one thread per
// candidate; weighted classification and regression sum in subset order (sequential chains, one
// thread per candidate) because FP addition is not associative.
#include <algorithm>
#include <chrono>
#include <type_traits>

#include "internal.h"

namespace {

enum { TASK_CLS = 0, TASK_CLSW = 1, TASK_REG = 2 };
enum { CF_CONST = 1, CF_NAN = 2, CF_MIL = 4 };
enum { ST_VMM = 0, ST_VSC, ST_SROWS, ST_PROWS, ST_DRAWS, ST_CONST, ST_SCORED, ST_MISMATCH, ST_PARNODES, ST_AMBIG, ST_COUNT };

constexpr int NT_MAX = 32;        // "tiny" nodes: one warp per node, one LANE per candidate
constexpr int NW_MAX = 512;       // nodes up to this many samples are owned by one warp (lanes on samples)
constexpr int BITS_W = NW_MAX / 32;
constexpr int NM_MAX = 2048;      // nodes up to this many samples are owned by a 128-thread CTA
constexpr int MID_TEAM = 128;
constexpr int CTA_TEAM = 512;     // threads of the CTA that owns a larger node
// CTA teams park one candidate's gathered values in shared memory when the node fits: the threshold pass
// then needs no second gather (which would go to DRAM again at the top levels, where L2 is thrashed)
// (8 B value + 4 B row + 1 B label per sample: 2048 -> 26 KB for the 128-thread team, 8192 -> 104 KB for the
// 512-thread team, two of which fit one SM)
__host__ __device__ constexpr int stage_cap(int team) { return team == 32 ? 0 : (team == MID_TEAM ? NM_MAX : 8192); }
#ifndef CBIG_CTAS
#define CBIG_CTAS 2
#endif
constexpr int CBIG_TEAM = 256;    // threads of the CTA that owns a larger node of a byte-coded table
constexpr int WARPS_PER_CTA = 4;  // warp teams per CTA in the small-node kernel

// Size classes of the open nodes (upper bounds in P::cls_max; an empty class repeats its predecessor's bound):
//   0..4  resident subtrees (subtree.cuh) when the row-major copy of the table exists and a useful number of
//         rows fits one SM: n <= rw * {1, 2, 4, 8, 16}, built to the leaves by teams of 1..16 warps (k_sub);
//         otherwise n <= 32, 64, 128, 256, 512: one warp per node, one lane per candidate (k_lane; classes
//         1..4 only on byte-coded tables, each with shared memory sized to its bound)
//         (empty when the subtree builder is off)
//   5..9  n <= 32, 64, 128, 256, 512: one warp per node, one lane per candidate (k_lane; classes 6..9 only on
//         byte-coded tables, each with shared memory sized to its bound); FP64 tables: class 9 = n <= 512, one
//         warp per node with lanes on samples (k_node<32>)
//   10, 11  n <= 2048, larger: one CTA per node (k_node)
constexpr int NQ = 12;
constexpr int Q_LANE0 = 5, Q_WARP = 9, Q_MID = 10, Q_CTA = 11;

struct Counters {
  int32_t next_f;
  int32_t q_count[NQ];
  int32_t sub_rows;  // rows of the nodes queued for the resident subtree builder at the next level
  int32_t n_leaves;
  unsigned int sub_nodes;  // nodes written to the subtree block pool so far
  unsigned long long scratch_words;
  unsigned long long st[ST_COUNT];
};

struct Frontier {
  int32_t *tree, *begin, *end, *node, *depth;
  int64_t *trace;
  uint64_t *key;
  int32_t *hist;   // [F][C]  (TASK_CLS)
  uint32_t *mask;  // [F][W]  (free-running: features known constant on the path from the root)
};

struct Pool {  // output nodes in creation order
  int32_t *tree, *feat, *child;  // feat: -1 leaf | feature + MIL bit; child: left child id | leaf slot
  double *cut;
  double *leaf_vals;  // [leaf slot][lw]
};

struct Trace {
  const int64_t *cand_begin;
  const int32_t *cand_count, *left, *right, *cand_feature;
  const double *cand_u;
  const uint8_t *cand_flag;
};

// shared-memory layout of one team (identical on host and device)
struct Lay {
  int o_u, o_cut, o_score, o_dist, o_redd, o_wh, o_xs, o_ys;                                // doubles
  int o_feat, o_flags, o_nleft, o_hnode, o_besthl, o_hist, o_redi, o_mask, o_bits, o_misc;  // int32
  int o_rows, o_lab, o_cm, o_ord;
  int o_coloff, o_cred, o_cb, o_thr;  // byte-coded CTA teams: column offsets, reduction scratch, per-candidate bytes
  int o_park;                         // ... and (128-thread teams) the parked bytes of the node [sample][32 candidates]
  int hs;      // stride of one candidate's histogram row (odd: conflict-free per-candidate reads)
  int use_cm;  // warp teams: per-chunk class bitmasks fit in shared memory
  int bytes;
};

__host__ __device__ inline Lay make_lay(int task, int team, int C, int NB, int W, bool replay, bool coded = false) {
  const bool warp_team = (team == 32);
  Lay L;
  int o = 0;  // in 8-byte units first
  L.o_coloff = o;
  o += coded ? 32 : 0;
  L.o_park = o;  // 16-byte aligned: o counts 8-byte units and everything before is a multiple of 2
  o += 0;  // (measured: parking costs more occupancy than the second gather pass costs time; kept switchable)
  L.o_u = o;
  o += NB;
  L.o_cut = o;
  o += NB;
  L.o_score = o;
  o += NB + 1;
  L.o_dist = o;
  o += (task == TASK_REG) ? 0 : C;
  L.o_redd = o;
  o += warp_team ? 0 : 64;
  L.o_wh = o;
  o += (task == TASK_CLSW) ? NB * 2 * C : 0;
  L.o_xs = o;  // warp teams stage the node once: rows, labels / targets, and one candidate's values
  o += warp_team ? NW_MAX : (coded ? 0 : stage_cap(team));
  L.o_ys = o;
  o += (warp_team && task != TASK_CLS) ? NW_MAX : 0;
  int oi = o * 2;  // switch to 4-byte units
  L.hs = (2 * C) | 1;
  L.o_feat = oi;
  oi += NB;
  L.o_flags = oi;
  oi += NB;
  L.o_nleft = oi;
  oi += NB;
  L.o_hnode = oi;
  oi += (task == TASK_CLS) ? C : 0;
  L.o_besthl = oi;
  oi += (task == TASK_CLS) ? C : 0;
  L.o_hist = oi;
  oi += (task == TASK_CLS) ? NB * L.hs : 0;
  L.o_redi = oi;
  oi += warp_team ? 0 : 128;
  L.o_mask = oi;
  oi += replay ? 0 : 2 * W;
  L.o_bits = oi;
  oi += (task != TASK_CLS && warp_team) ? NB * 2 * BITS_W : 0;
  L.o_misc = oi;
  oi += 8;
  L.o_ord = oi;
  oi += 32;
  L.o_cred = oi;
  oi += coded ? 16 * 24 : 0;
  L.o_cb = oi;  // 4 byte arrays of 32 candidates: thr - 1, enable, K, nan-enable
  oi += coded ? 32 : 0;
  L.o_thr = oi;
  oi += coded ? 32 : 0;
  L.o_rows = oi;
  oi += warp_team ? NW_MAX : stage_cap(team);
  L.o_lab = oi;  // warp teams: int32 labels; CTA teams: uint8 labels (used when C <= 256)
  oi += (task != TASK_REG) ? (warp_team ? NW_MAX : stage_cap(team) / 4) : 0;
  L.use_cm = (warp_team && task == TASK_CLS && BITS_W * C * 4 <= 8192) ? 1 : 0;
  L.o_cm = oi;
  oi += L.use_cm ? BITS_W * C : 0;
  L.bytes = ((oi + 3) / 4) * 16;
  return L;
}

struct P {
  const double *X;
  int64_t ld, n, n_table;
  int32_t d, C, k, n_min, max_depth, W, task, replay, NB;
  int32_t *idx_src, *idx_dst, *yc_src, *yc_dst;
  double *yr_src, *yr_dst, *w_src, *w_dst;
  Frontier cur, nxt;
  int32_t *q_cur[NQ], *q_nxt[NQ];
  Pool o;
  Trace tr;
  Counters *cnt;
  uint32_t *scratch;  // side bitmasks of CTA-owned nodes (TASK_CLSW / TASK_REG)
  int32_t node_base_next;
  const uint8_t *C8;   // byte codes of the table, column-major [d][ldc] (encode.cu); null on FP64-only tables
  int64_t ldc;
  const double *dict;  // [d][256]
  const uint8_t *coff; // [d] stored byte + coff = wide code (0 NaN, r + 1 for dict[r])
  int32_t c8_small;    // the coded table is smaller than 4 GiB: gathers use 32-bit offsets from C8
  int32_t cls_max[NQ - 1];
  // resident subtree builder (subtree.cuh)
  const uint8_t *R8;   // row-major byte codes [n][r8_stride] (also gathered by k_lane)
  int32_t r8_stride;
  int32_t nc_max;      // k_lane: nodes of up to this many rows draw from their varying-feature set
  const double *XR;    // row-major FP64 [n][sub_rowbytes / 8]
  int32_t sub_ncls;    // size classes 0 .. sub_ncls - 1 are built by k_sub (0: off)
  int32_t sub_rw, sub_rowbytes;  // staged rows per warp, bytes per staged row
  PNode *sub_nodes;    // block pool: pre-order blocks of finished subtrees
};

__host__ __device__ inline int size_class(const P &p, int64_t n) {
  int q = 0;
  while (q < NQ - 1 && n > p.cls_max[q]) q++;
  return q;
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
      int lo = w * 32;
      uint32_t m = (p.d - lo >= 32) ? 0u : (0xffffffffu << (p.d - lo));  // padding bits count as taken
      p.cur.mask[(int64_t)t * p.W + w] = m;
    }
  }
  p.q_cur[size_class(p, p.n)][t] = t;
}

// (kept out of line: the closed form is long and is called from several places of every node kernel)
__device__ __noinline__ double repeat_add_dev(double c, int64_t h) { return et_repeat_add(c, h); }

// ---- team helpers ---------------------------------------------------------------------------
template <int TEAM>
__device__ __forceinline__ void team_sync() {
  if (TEAM == 32)
    __syncwarp();
  else
    __syncthreads();
}

// min / max / any over the team; result valid in every thread
template <int TEAM>
__device__ __forceinline__ void team_minmax(double &mn, double &mx, int &flag, double *redd, int32_t *redi) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double omn = __shfl_xor_sync(0xffffffffu, mn, o);
    double omx = __shfl_xor_sync(0xffffffffu, mx, o);
    if (omn < mn) mn = omn;
    if (omx > mx) mx = omx;
  }
  flag = __any_sync(0xffffffffu, flag);
  if (TEAM > 32) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = TEAM / 32;
    __syncthreads();  // previous users of the scratch are done
    if (lane == 0) {
      redd[w] = mn;
      redd[32 + w] = mx;
      redi[w] = flag;
    }
    __syncthreads();
    double a = lane < nw ? redd[lane] : 1.7976931348623157e308;
    double b = lane < nw ? redd[32 + lane] : -1.7976931348623157e308;
    int f = lane < nw ? redi[lane] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double oa = __shfl_xor_sync(0xffffffffu, a, o);
      double ob = __shfl_xor_sync(0xffffffffu, b, o);
      if (oa < a) a = oa;
      if (ob > b) b = ob;
    }
    mn = a;
    mx = b;
    flag = __any_sync(0xffffffffu, f);
  }
}

template <int TEAM>
__device__ __forceinline__ bool team_all(bool v, int32_t *redi) {
  bool r = __all_sync(0xffffffffu, v);
  if (TEAM > 32) {
    __syncthreads();
    if ((threadIdx.x & 31) == 0) redi[64 + (threadIdx.x >> 5)] = r;
    __syncthreads();
    bool a = true;
    for (int w = 0; w < TEAM / 32; w++) a &= (redi[64 + w] != 0);
    r = a;
  }
  return r;
}

// fixed-shape sum over the team (butterfly inside a warp, then the warps in index order): the same
// inputs always give the same bits, independent of scheduling; result valid in every thread
template <int TEAM>
__device__ __forceinline__ double team_sum(double v, double *redd) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = ET_ADD(v, __shfl_xor_sync(0xffffffffu, v, o));
  if (TEAM > 32) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = TEAM / 32;
    __syncthreads();
    if (lane == 0) redd[w] = v;
    __syncthreads();
    double a = 0.0;
    for (int q = 0; q < nw; q++) a = ET_ADD(a, redd[q]);
    v = a;
  }
  return v;
}

// Variance-reduction score of a side (n_in, S_in, Q_in) = count, sum and sum of squares of (y - mu) over
// the samples going left, mu = the node's mean (pkg:1196-1218 evaluated from moments about the node mean;
// used for nodes too large for the reference's sequential order to be affordable, see k_node REGPAR).
__device__ __forceinline__ double var_reduction_moments(int32_t n, double S_tot, double Q_tot, double V, int32_t ni,
                                                        double Si, double Qi) {
  const int32_t no = n - ni;
  if (ni < 1 || no < 1) return NAN;
  const double So = ET_SUB(S_tot, Si), Qo = ET_SUB(Q_tot, Qi);
  const double dni = (double)ni, dno = (double)no, dn = (double)n;
  const double vi = (ni == 1) ? 0.0 : ET_DIV(ET_SUB(Qi, ET_DIV(ET_MUL(Si, Si), dni)), dni);
  const double vo = (no == 1) ? 0.0 : ET_DIV(ET_SUB(Qo, ET_DIV(ET_MUL(So, So), dno)), dno);
  const double a = ET_MUL(ET_DIV(dni, dn), vi);
  const double bq = ET_MUL(ET_DIV(dno, dn), vo);
  return ET_DIV(ET_SUB(ET_SUB(V, a), bq), V);
}

// ---- exact scores ---------------------------------------------------------------------------
// giniScore (pkg:1101-1158, unweighted) from integer histograms: hin[c] = hl[c] (+ hn[c] when NaN
// rows go left); hout = node histogram - hin.
__device__ __noinline__ double gini_score_int(const int32_t *hnode, const int32_t *hl, const int32_t *hn, bool nan_left, int C,
                                 int32_t n, double G, int32_t *cin_out) {
  int32_t cin_i = 0;
  for (int c = 0; c < C; c++) cin_i += hl[c] + (nan_left ? hn[c] : 0);
  *cin_out = cin_i;
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

// computeVarianceReduction (pkg:1196-1218) with saddle's two-pass sampleVariance, in subset order.
__device__ double var_reduction_seq(const double *y, int32_t n, const uint32_t *mlt, const uint32_t *mnan,
                                    bool nan_left, double V, int32_t *nin_out) {
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
  *nin_out = nin;
  int32_t nout = n - nin;
  double dnin = (double)nin, dnout = (double)nout, dn = (double)n;
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
  double vin = (nin == 1) ? 0.0 : ET_DIV(ET_MUL(svin, ET_SUB(dnin, 1.0)), dnin);
  double vout = (nout == 1) ? 0.0 : ET_DIV(ET_MUL(svout, ET_SUB(dnout, 1.0)), dnout);
  double a = ET_MUL(ET_DIV(dnin, dn), vin);
  double bq = ET_MUL(ET_DIV(dnout, dn), vout);
  return ET_DIV(ET_SUB(ET_SUB(V, a), bq), V);
}

// weighted giniScore (pkg:1132-1157): sequential weighted class sums in subset order.
__device__ double gini_score_w_seq(const int32_t *y, const double *w, int32_t n, const uint32_t *mlt,
                                   const uint32_t *mnan, bool nan_left, int C, double G, double N, double *hin,
                                   double *hout, int32_t *nin_out) {
  for (int q = 0; q < C; q++) {
    hin[q] = 0.0;
    hout[q] = 0.0;
  }
  double cin = 0.0, cout = 0.0;
  int32_t nin = 0;
  for (int32_t j = 0; j < n; j++) {
    uint32_t m = mlt[j >> 5];
    if (nan_left) m |= mnan[j >> 5];
    double ww = w[j];
    int32_t cls = y[j];
    if ((m >> (j & 31)) & 1u) {
      cin = ET_ADD(cin, ww);
      hin[cls] = ET_ADD(hin[cls], ww);
      nin++;
    } else {
      cout = ET_ADD(cout, ww);
      hout[cls] = ET_ADD(hout[cls], ww);
    }
  }
  *nin_out = nin;
  double sin_ = 0.0, sout = 0.0;
  for (int q = 0; q < C; q++) {
    double pi = ET_DIV(hin[q], cin), po = ET_DIV(hout[q], cout);
    sin_ = ET_ADD(sin_, ET_MUL(pi, pi));
    sout = ET_ADD(sout, ET_MUL(po, po));
  }
  double gin = ET_SUB(1.0, sin_), gout = ET_SUB(1.0, sout);
  return ET_SUB(ET_SUB(G, ET_DIV(ET_MUL(gin, cin), N)), ET_DIV(ET_MUL(gout, cout), N));
}

// position of the rank-th clear bit of taken[0..W) (rank < number of clear bits)
__device__ __forceinline__ int32_t rank_select_clear(const uint32_t *taken, int W, int32_t rank) {
  for (int w = 0; w < W; w++) {
    uint32_t z = ~taken[w];
    int c = __popc(z);
    if (rank < c) return w * 32 + (int32_t)__fns(z, 0, rank + 1);
    rank -= c;
  }
  return -1;
}

// ---- the node kernel ------------------------------------------------------------------------
// One team per open node (TEAM == 32: a warp, several per CTA; else one CTA).  Implements
// buildTree* (stop rules, pkg:993-994 / 813-814), split* (pkg:232-296 / 453-509: candidates
// consumed in draw order, constants and NaN scores do not count toward k, strict `>` keeps the
// first best), the child filters (pkg:1024-1039) and child creation.
// position of the r-th set bit of z (r < popc(z))
__device__ __forceinline__ int select_bit32(uint32_t z, int r) {
  int pos = 0;
#pragma unroll
  for (int w = 16; w > 0; w >>= 1) {
    const int c = __popc(z & ((1u << w) - 1u));
    if (r >= c) {
      r -= c;
      z >>= w;
      pos += w;
    }
  }
  return pos;
}

__device__ __forceinline__ int32_t rank_select_clear_fast(const uint32_t *taken, int W, int32_t rank) {
  for (int w = 0; w < W; w++) {
    const uint32_t z = ~taken[w];
    const int c = __popc(z);
    if (rank < c) return w * 32 + select_bit32(z, rank);
    rank -= c;
  }
  return -1;
}


// ---- byte-coded CTA teams: the two streaming passes over a node's samples ----------------------
// NG = groups of 4 candidates read per sample (the batch holds up to 4 * NG candidates).  The loops carry
// no branch, so all 4 * NG byte loads of a sample are in flight together.
template <int NG>
__device__ __forceinline__ void coded_load(const uint8_t *__restrict__ C8, const int64_t *s_coloff, int64_t r,
                                           uint32_t (&b4)[NG]) {
  // (the byte-coded CTA teams only run on coded tables below 4 GiB: one 32-bit add per load)
  const uint32_t *off32 = reinterpret_cast<const uint32_t *>(s_coloff);
  const uint32_t r32 = (uint32_t)r;
  uint32_t b[4 * NG];
#pragma unroll
  for (int c = 0; c < 4 * NG; c++) b[c] = __ldg(C8 + (off32[2 * c] + r32));
#pragma unroll
  for (int g = 0; g < NG; g++) b4[g] = b[4 * g] | (b[4 * g + 1] << 8) | (b[4 * g + 2] << 16) | (b[4 * g + 3] << 24);
}

// pass 1: per-candidate min of (byte - K), max of byte, min of byte, packed 4 candidates per register
// (s_park != null: the packed bytes of every sample are parked in shared memory [sample][NG] for pass 2)
template <int NG, int TEAM>
__device__ __forceinline__ void coded_pass1(const uint8_t *__restrict__ C8, const int64_t *s_coloff,
                                            const uint32_t *s_K4, const int32_t *rr, int32_t n, int tid,
                                            uint32_t *s_cred, int wit, int lane, uint32_t *s_park) {
  // Two candidates per register as 16-bit halves: sm_100a has native 16x2 min / max / add (VIMNMX.U16x2,
  // VIADD.16x2), while the 8x4 video intrinsics are emulated with 7-12 logic instructions each.
  constexpr int NP = 2 * NG;
  uint32_t mnT[NP], mxB[NP], mnB[NP], K2[NP];
#pragma unroll
  for (int q = 0; q < NP; q++) {
    mnT[q] = 0xffffffffu;
    mxB[q] = 0u;
    mnB[q] = 0xffffffffu;
    const uint32_t k4 = s_K4[q >> 1] >> (16 * (q & 1));  // bytes 2q, 2q + 1 of the packed K
    K2[q] = (k4 & 0xffu) | ((k4 & 0xff00u) << 8);
  }
  for (int32_t j = tid; j < n; j += TEAM) {
    const uint32_t r32 = (uint32_t)rr[j];
    const uint32_t *off32 = reinterpret_cast<const uint32_t *>(s_coloff);
    uint32_t b[4 * NG];
#pragma unroll
    for (int c = 0; c < 4 * NG; c++) b[c] = __ldg(C8 + (off32[2 * c] + r32));
    if (s_park) {
      uint32_t b4[NG];
#pragma unroll
      for (int g = 0; g < NG; g++) b4[g] = b[4 * g] | (b[4 * g + 1] << 8) | (b[4 * g + 2] << 16) | (b[4 * g + 3] << 24);
      if (NG >= 4) {
#pragma unroll
        for (int g = 0; g < NG; g += 4)
          *reinterpret_cast<uint4 *>(s_park + (size_t)j * NG + g) = make_uint4(b4[g], b4[g + 1], b4[g + 2], b4[g + 3]);
      } else {
        *reinterpret_cast<uint2 *>(s_park + (size_t)j * NG) = make_uint2(b4[0], b4[1]);
      }
    }
#pragma unroll
    for (int q = 0; q < NP; q++) {
      const uint32_t b2 = b[2 * q] | (b[2 * q + 1] << 16);
      mxB[q] = __vmaxu2(mxB[q], b2);
      mnB[q] = __vminu2(mnB[q], b2);
      mnT[q] = __vminu2(mnT[q], __vsub2(b2, K2[q]));  // NaN (byte 0 of a column with NaNs) wraps to 0xffff
    }
  }
#pragma unroll
  for (int q = 0; q < NP; q++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mxB[q] = __vmaxu2(mxB[q], __shfl_xor_sync(0xffffffffu, mxB[q], o));
      mnB[q] = __vminu2(mnB[q], __shfl_xor_sync(0xffffffffu, mnB[q], o));
      mnT[q] = __vminu2(mnT[q], __shfl_xor_sync(0xffffffffu, mnT[q], o));
    }
    if (lane == 0) {  // back to the byte layout the decode step reads: 4 candidates per word, words g, 8 + g, 16 + g
      uint8_t *cred8 = reinterpret_cast<uint8_t *>(s_cred + wit * 24);
      cred8[2 * q] = (uint8_t)min(mnT[q] & 0xffffu, 255u);
      cred8[2 * q + 1] = (uint8_t)min(mnT[q] >> 16, 255u);
      cred8[32 + 2 * q] = (uint8_t)(mxB[q] & 0xffu);
      cred8[32 + 2 * q + 1] = (uint8_t)((mxB[q] >> 16) & 0xffu);
      cred8[64 + 2 * q] = (uint8_t)min(mnB[q] & 0xffffu, 255u);
      cred8[64 + 2 * q + 1] = (uint8_t)min(mnB[q] >> 16, 255u);
    }
  }
}

// pass 2: side histograms.  Per 32 consecutive samples: one ballot per candidate, counted against the
// class masks with lane == class.  sweep 0 counts x < cut, sweep 1 the NaN samples (pkg:244-248).
// 32 x 32 bit transpose across a warp: lane r gives row word x (bit c), lane c receives bit r of every row
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, int lane) {
#pragma unroll
  for (int sft = 16; sft > 0; sft >>= 1) {
    const uint32_t m = (sft == 16) ? 0x0000ffffu : (sft == 8) ? 0x00ff00ffu : (sft == 4) ? 0x0f0f0f0fu
                     : (sft == 2) ? 0x33333333u : 0x55555555u;  // bits whose index has bit `sft` clear
    const uint32_t y = __shfl_xor_sync(0xffffffffu, x, sft);
    x = (lane & sft) ? ((x & ~m) | ((y & ~m) >> sft)) : ((x & m) | ((y & m) << sft));
  }
  return x;
}

template <int NG, int TEAM, typename LabFn>
__device__ __forceinline__ void coded_pass2(const uint8_t *__restrict__ C8, const int64_t *s_coloff,
                                            const uint32_t *s_K4, const uint32_t *s_t4, const uint32_t *s_e4, int sweep,
                                            const int32_t *rr, LabFn lab, int32_t n, int C, int wit, int lane,
                                            int32_t *s_hist, int hs, int hoff, int nb, const uint32_t *s_park) {
  // Per 32 consecutive samples: every thread packs the side bits of its sample for all candidates into one word,
  // one 32 x 32 bit transpose hands lane c the 32 samples' bits of candidate c, and the class counts are
  // popc(bits & class mask) per class present -- C ballots and C popcounts per 32 samples instead of one ballot and
  // one popcount per candidate and class mask.  lane == candidate; acc[k] counts class k.
  int32_t acc[32];
  uint32_t K4[NG], t4[NG], e4[NG];
#pragma unroll
  for (int k = 0; k < 32; k++) acc[k] = 0;
#pragma unroll
  for (int g = 0; g < NG; g++) {
    K4[g] = s_K4[g];
    t4[g] = s_t4[g];
    e4[g] = s_e4[g];
  }
  for (int32_t j0 = wit * 32; j0 < n; j0 += TEAM) {
    const int32_t j = j0 + lane;
    const bool valid = j < n;
    const int32_t cls = valid ? lab(j) : -1;
    uint32_t b4[NG];
    if (s_park) {
      const int32_t jp = valid ? j : 0;
      if (NG >= 4) {
#pragma unroll
        for (int g = 0; g < NG; g += 4) {
          const uint4 v = *reinterpret_cast<const uint4 *>(s_park + (size_t)jp * NG + g);
          b4[g] = v.x;
          b4[g + 1] = v.y;
          b4[g + 2] = v.z;
          b4[g + 3] = v.w;
        }
      } else {
        const uint2 v = *reinterpret_cast<const uint2 *>(s_park + (size_t)jp * NG);
        b4[0] = v.x;
        b4[1] = v.y;
      }
    } else {
      coded_load<NG>(C8, s_coloff, valid ? (int64_t)rr[j] : 0, b4);
    }
    uint32_t rowbits = 0u;  // bit c: this sample is on the counted side for candidate c
#pragma unroll
    for (int g = 0; g < NG; g++) {
      uint32_t l4 = sweep ? __vcmpeq4(b4[g], 0u) : __vcmpleu4(__vsub4(b4[g], K4[g]), t4[g]);
      l4 &= e4[g];
      rowbits |= (((l4 & 0x01010101u) * 0x01020408u) >> 24) << (4 * g);
    }
    if (!valid) rowbits = 0u;
    const uint32_t candbits = warp_transpose32(rowbits, lane);  // bit r: sample j0 + r, for candidate `lane`
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

template <int TASK, int TEAM, bool CODED>
__global__ void __launch_bounds__(TEAM == 32 ? 32 * WARPS_PER_CTA : TEAM,
                                  TEAM == 32 ? 5
                                             : (TEAM == MID_TEAM ? (CODED ? 4 : 6)
                                                                 : ((TASK == TASK_REG && TEAM == CTA_TEAM) ? 1 : ((CODED && TEAM == CBIG_TEAM) ? CBIG_CTAS : 2))))
    k_node(P p, int32_t qcount, int qi) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr bool WARP = (TEAM == 32);
  static_assert(!CODED || (TASK == TASK_CLS && TEAM > 32), "byte-coded teams: unweighted classification, CTA teams");
  // Regression nodes of more than NM_MAX samples: the reference sums targets sequentially in subset order
  // (saddle's two-pass sampleVariance), which one thread would have to replay for up to a million samples per
  // candidate.  These nodes are scored from fixed-shape parallel sums of moments about the node mean instead:
  // deterministic, within ~1e-15 (relative to the node variance) of the exactly rounded value -- closer to it
  // than the reference's own sequential sum -- but not order-identical, so two candidates whose scores differ by
  // less than that could swap.  Such splits are counted (et_stats.ambiguous_splits; 1e-9 relative) so that a
  // replay run can tell.  Smaller nodes (where exact ties live) keep the exact sequential evaluation.
  constexpr bool REGPAR = (TASK == TASK_REG && TEAM == CTA_TEAM && !CODED);
  const int tic = WARP ? (threadIdx.x >> 5) : 0;
  const int q = WARP ? blockIdx.x * WARPS_PER_CTA + tic : blockIdx.x;
  if (q >= qcount) return;
  const int tid = WARP ? (threadIdx.x & 31) : threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int wit = WARP ? 0 : (threadIdx.x >> 5);  // warp index inside the team
  const int C = p.C, NB = p.NB, W = p.W;
  const Lay L = make_lay(TASK, TEAM, C, NB, W, p.replay != 0, CODED);
  unsigned char *sm = smem_raw + (size_t)tic * L.bytes;
  double *smd = reinterpret_cast<double *>(sm);
  int32_t *smi = reinterpret_cast<int32_t *>(sm);
  double *s_u = smd + L.o_u, *s_cut = smd + L.o_cut, *s_score = smd + L.o_score, *s_dist = smd + L.o_dist;
  double *s_redd = smd + L.o_redd, *s_wh = smd + L.o_wh, *s_xs = smd + L.o_xs, *s_ys = smd + L.o_ys;
  int32_t *s_feat = smi + L.o_feat, *s_flags = smi + L.o_flags, *s_nleft = smi + L.o_nleft;
  int32_t *s_hnode = smi + L.o_hnode, *s_besthl = smi + L.o_besthl, *s_hist = smi + L.o_hist, *s_redi = smi + L.o_redi;
  uint32_t *s_const = reinterpret_cast<uint32_t *>(smi + L.o_mask), *s_taken = s_const + W;
  uint32_t *s_bits = reinterpret_cast<uint32_t *>(smi + L.o_bits);
  int32_t *s_misc = smi + L.o_misc, *s_rows = smi + L.o_rows, *s_lab = smi + L.o_lab, *s_ord = smi + L.o_ord;
  uint32_t *s_cm = reinterpret_cast<uint32_t *>(smi + L.o_cm);
  int64_t *s_coloff = reinterpret_cast<int64_t *>(smd + L.o_coloff);
  uint32_t *s_cred = reinterpret_cast<uint32_t *>(smi + L.o_cred);
  uint8_t *s_thrb = reinterpret_cast<uint8_t *>(smi + L.o_cb), *s_enb = s_thrb + 32, *s_Kb = s_thrb + 64, *s_nanb = s_thrb + 96;
  int32_t *s_thr = smi + L.o_thr;
  uint32_t *s_park = nullptr;  // see make_lay: o_park

  const int i = p.q_cur[qi][q];
  const int32_t tree = p.cur.tree[i], b = p.cur.begin[i], e = p.cur.end[i], n = e - b;
  const int32_t node = p.cur.node[i], depth = p.cur.depth[i];
  const int64_t tn = p.cur.trace[i];
  const uint64_t key = p.cur.key[i];
  const int64_t base = (int64_t)tree * p.n;
  const int lw = (TASK == TASK_REG) ? 1 : C;
  const int nv = (n + 31) >> 5;
  const bool staged = !WARP && n <= stage_cap(TEAM);  // CTA team: one candidate's values fit in shared memory

  // A warp team stages its node once in shared memory (rows, labels, targets / weights); a CTA team
  // streams the node's segment from HBM/L2.  rr/ll/yy/ww are indexed by position inside the node.
  if (WARP) {
    for (int j = lane; j < n; j += 32) {
      s_rows[j] = p.idx_src[base + b + j];
      if (TASK != TASK_REG) s_lab[j] = p.yc_src[base + b + j];
      if (TASK == TASK_REG) s_ys[j] = p.yr_src[base + b + j];
      if (TASK == TASK_CLSW) s_ys[j] = p.w_src[base + b + j];
    }
    __syncwarp();
  }
  const bool staged_lab = staged && TASK != TASK_REG && C <= 256;
  uint8_t *s_lab8 = reinterpret_cast<uint8_t *>(s_lab);
  if (staged) {
    for (int j = tid; j < n; j += TEAM) {
      s_rows[j] = p.idx_src[base + b + j];
      if (staged_lab) s_lab8[j] = (uint8_t)p.yc_src[base + b + j];
    }
    __syncthreads();
  }
  const int32_t *rr = (WARP || staged) ? s_rows : (p.idx_src + base + b);
  const int32_t *ll = (TASK == TASK_REG) ? nullptr : (WARP ? s_lab : (p.yc_src + base + b));
#define LAB(j) (staged_lab ? (int32_t)s_lab8[(j)] : ll[(j)])
  const double *yy = (TASK != TASK_REG) ? nullptr : (WARP ? s_ys : (p.yr_src + base + b));
  const double *ww = (TASK != TASK_CLSW) ? nullptr : (WARP ? s_ys : (p.w_src + base + b));

  // ---------------- stop rules + node totals ----------------
  bool leaf;
  double total = 0.0, nsum = (double)n, leaf_mean = 0.0;
  double reg_mu = 0.0, reg_S = 0.0, reg_Q = 0.0;  // REGPAR: node mean, sum and sum of squares of (y - mean)
  double second_score = -INFINITY;                // REGPAR: runner-up score (ambiguity check)
  if (TASK == TASK_CLS) {
    for (int c = tid; c < C; c += TEAM) s_hnode[c] = p.cur.hist[(int64_t)i * C + c];
    team_sync<TEAM>();
    bool pure = false;
    for (int c = 0; c < C; c++) pure |= (s_hnode[c] == n);
    leaf = (p.n_table < p.n_min) || (depth >= p.max_depth) || pure;  // pkg:993-994
    if (!leaf) {
      // giniImpurity with the reference's repeated `+= 1/s` distribution (pkg:905-911, 1160-1180)
      const double inv = ET_DIV(1.0, (double)n);
      for (int c = tid; c < C; c += TEAM) s_dist[c] = repeat_add_dev(inv, s_hnode[c]);
      team_sync<TEAM>();
      double s = 0.0;
      for (int c = 0; c < C; c++) s = ET_ADD(s, ET_MUL(s_dist[c], s_dist[c]));
      total = ET_SUB(1.0, s);
    }
  } else if (TASK == TASK_REG) {
    const double head = yy[0];
    bool uni = true;
    for (int32_t j = tid; j < n; j += TEAM) uni &= !(yy[j] != head);
    uni = team_all<TEAM>(uni, s_redi);
    leaf = (n < p.n_min) || (depth >= p.max_depth) || uni;  // pkg:813-814
    if (REGPAR) {
      double s1 = 0.0;
      for (int32_t j = tid; j < n; j += TEAM) s1 = ET_ADD(s1, yy[j]);
      const double dn = (double)n;
      reg_mu = ET_DIV(team_sum<TEAM>(s1, s_redd), dn);
      double q1 = 0.0, s2 = 0.0;
      for (int32_t j = tid; j < n; j += TEAM) {
        const double dl = ET_SUB(yy[j], reg_mu);
        s2 = ET_ADD(s2, dl);
        q1 = ET_ADD(q1, ET_MUL(dl, dl));
      }
      reg_S = team_sum<TEAM>(s2, s_redd);
      reg_Q = team_sum<TEAM>(q1, s_redd);
      leaf_mean = reg_mu;
      total = ET_DIV(ET_SUB(reg_Q, ET_DIV(ET_MUL(reg_S, reg_S), dn)), dn);
      __syncthreads();
    } else {
    // mean2 (pkg:782) and varianceNoSplit (pkg:436-437), sequential in subset order
    if (tid == 0) {
      double sum = 0.0;
      for (int32_t j = 0; j < n; j++) sum = ET_ADD(sum, yy[j]);
      const double dn = (double)n;
      const double mean = ET_DIV(sum, dn);
      double V = 0.0;
      if (!leaf) {
        double var = 0.0;
        if (n > 1) {
          double qq = 0.0;
          for (int32_t j = 0; j < n; j++) {
            double dl = ET_SUB(yy[j], mean);
            qq = ET_ADD(qq, ET_MUL(dl, dl));
          }
          var = ET_DIV(qq, ET_SUB(dn, 1.0));
        }
        V = ET_DIV(ET_MUL(var, ET_SUB(dn, 1.0)), dn);
      }
      s_score[0] = mean;
      s_score[NB] = V;
    }
    team_sync<TEAM>();
    leaf_mean = s_score[0];
    total = s_score[NB];
    team_sync<TEAM>();
    }
  } else {
    const int32_t head = ll[0];
    bool uni = true;
    for (int32_t j = tid; j < n; j += TEAM) uni &= (ll[j] == head);
    uni = team_all<TEAM>(uni, s_redi);
    leaf = (p.n_table < p.n_min) || (depth >= p.max_depth) || uni;
    // weighted distribution (pkg:913-927): sequential in subset order; also the leaf value
    if (tid == 0) {
      for (int c = 0; c < C; c++) s_dist[c] = 0.0;
      double s = 0.0;
      for (int32_t j = 0; j < n; j++) {
        const double w1 = ww[j];
        const int32_t cls = ll[j];
        s_dist[cls] = ET_ADD(s_dist[cls], w1);
        s = ET_ADD(s, w1);
      }
      double sq = 0.0;
      for (int c = 0; c < C; c++) {
        const double pc = ET_DIV(s_dist[c], s);
        s_dist[c] = pc;
        sq = ET_ADD(sq, ET_MUL(pc, pc));
      }
      s_score[0] = ET_SUB(1.0, sq);
      s_score[NB] = s;  // sampleWeights.sum2 over the subset (pkg:1112): same order, same value
    }
    team_sync<TEAM>();
    total = s_score[0];
    nsum = s_score[NB];
    team_sync<TEAM>();
  }

  // ---------------- split search ----------------
  int32_t visited = 0, nconst = 0, best_feature = -1, best_nleft = 0, best_mil = 0;
  int32_t best_thr = 0, best_K = 0;  // byte-coded tables: the winning split in code space
  double best_score = -INFINITY, best_cut = NAN;
  unsigned long long st_draws = 0, st_const = 0, st_scored = 0, st_mismatch = 0;
  if (!leaf) {
    int32_t dc = 0, tpos = 0;
    int64_t tb = 0;
    int32_t tcnt = 0;
    if (p.replay) {
      if (tn >= 0) {
        tb = p.tr.cand_begin[tn];
        tcnt = p.tr.cand_count[tn];
      }
    } else {
      for (int w = tid; w < W; w += TEAM) {
        const uint32_t m = p.cur.mask[(int64_t)i * W + w];
        s_const[w] = m;
        s_taken[w] = m;
      }
      team_sync<TEAM>();
      int nc = 0;
      for (int w = 0; w < W; w++) nc += __popc(s_const[w]);
      nconst = nc - (W * 32 - p.d);
    }
    if (WARP && L.use_cm) {
      // per 32-sample chunk, one bitmask per class: counting a side histogram becomes popc(ballot & mask)
      for (int t = lane; t < nv * C; t += 32) s_cm[t] = 0u;
      __syncwarp();
      for (int v = 0; v < nv; v++) {
        const int j = v * 32 + lane;
        if (j < n) atomicOr(&s_cm[v * C + s_lab[j]], 1u << lane);
      }
      __syncwarp();
    }
    uint32_t *g_bits = nullptr;  // CTA teams keep side bitmasks in global scratch
    const int words = nv;
    if (TASK != TASK_CLS && !WARP && !REGPAR) {
      if (tid == 0) {
        unsigned long long off =
            atomicAdd(&p.cnt->scratch_words, (unsigned long long)NB * 2ull * (unsigned long long)words);
        s_misc[0] = (int32_t)(off & 0xffffffffull);
        s_misc[1] = (int32_t)(off >> 32);
      }
      team_sync<TEAM>();
      unsigned long long off = ((unsigned long long)(uint32_t)s_misc[1] << 32) | (uint32_t)s_misc[0];
      g_bits = p.scratch + off;
    }
    for (;;) {
      int32_t nb;
      const int32_t avail = p.d - nconst - visited;
      if (p.replay) {
        nb = min(NB, tcnt - tpos);
      } else {
        // draw what is still needed plus the constants expected among them (observed rate at this
        // node); candidates past the k-th scored one are discarded unexamined, like the reference
        // which stops drawing there
        const int32_t need = min(p.k - visited, avail);
        int32_t extra = 0;
        if (need > 0 && st_draws > 0)
          extra = (int32_t)(((long long)need * (long long)(st_draws - (unsigned long long)visited)) / max(visited, 1));
        nb = min(NB, min(avail, need + extra));
        if (need <= 0) nb = 0;
      }
      if (nb <= 0) break;
      // ---- draw a batch of candidates (one lane per candidate)
      if (wit == 0) {
        if (p.replay) {
          if (lane < nb) {
            s_feat[lane] = p.tr.cand_feature[tb + tpos + lane];
            s_u[lane] = p.tr.cand_u[tb + tpos + lane];
            s_flags[lane] = (p.tr.cand_flag[tb + tpos + lane] + 1) << 4;
          }
        } else {
          // uniform over the features that are neither known-constant nor taken; a lane whose pick
          // collides with a lower lane's pick sits this batch out (= sequential rejection sampling)
          int32_t f = -1 - lane;
          if (lane < nb) {
            const uint64_t r = et_draw(key, (uint32_t)(dc + 2 * lane));
            f = rank_select_clear_fast(s_taken, W, (int32_t)__umul64hi(r, (uint64_t)avail));
          }
          const uint32_t same = __match_any_sync(0xffffffffu, f);
          const bool keep = (lane < nb) && (lane == __ffs(same) - 1);
          if (lane < nb) {
            s_feat[lane] = keep ? f : -1;
            s_u[lane] = et_u01(et_draw(key, (uint32_t)(dc + 2 * lane + 1)));
            s_flags[lane] = 0;
          }
          __syncwarp();
          if (keep) atomicOr(&s_taken[f >> 5], 1u << (f & 31));
        }
      }
      if (p.replay)
        tpos += nb;
      else
        dc += 2 * NB;
      if (wit == 0) {
        // Evaluate the batch in ascending feature order (results are consumed in draw order below):
        // teams that run side by side then walk the column space together, so the handful of columns
        // in flight chip-wide stays resident in L2 instead of every team streaming its own column.
        __syncwarp();
        uint32_t keyv = (lane < nb && s_feat[lane] >= 0) ? (((uint32_t)s_feat[lane] << 5) | (uint32_t)lane)
                                                         : (0xffffffe0u | (uint32_t)lane);
#pragma unroll
        for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
          for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
            const uint32_t other = __shfl_xor_sync(0xffffffffu, keyv, j2);
            const bool up = ((lane & k2) == 0);
            const bool lower = ((lane & j2) == 0);
            const uint32_t lo = min(keyv, other), hi = max(keyv, other);
            keyv = (up == lower) ? lo : hi;
          }
        }
        s_ord[lane] = (keyv >= 0xffffffe0u) ? -1 : (int32_t)(keyv & 31u);
      }
      if (TASK == TASK_CLS)
        for (int t = tid; t < nb * L.hs; t += TEAM) s_hist[t] = 0;
      team_sync<TEAM>();
      if (REGPAR) {
        // ---- large regression node: groups of 4 candidates; per group one pass for min / max and one for the
        //      moments of the left side (a second sweep over the NaN samples only if a candidate has any)
        double *scr = s_xs;  // reduction scratch [warp][12] (the value-staging buffer is unused here)
        constexpr int NWP = TEAM / 32;
        for (int g0 = 0; g0 < nb; g0 += 4) {
          const double *colp[4];
#pragma unroll
          for (int c = 0; c < 4; c++) {
            const int32_t f = (g0 + c < nb) ? s_feat[g0 + c] : -1;
            colp[c] = p.X + (int64_t)(f >= 0 ? f : 0) * p.ld;
          }
          {
            double mn[4], mx[4];
            uint32_t nanm = 0u;
#pragma unroll
            for (int c = 0; c < 4; c++) {
              mn[c] = 1.7976931348623157e308;  // pkg:35-36
              mx[c] = -1.7976931348623157e308;
            }
            // four samples per thread and trip: 16 independent gathers in flight
            for (int32_t j0 = tid; j0 < n; j0 += 4 * TEAM) {
              int32_t r[4];
              double x[4][4];
#pragma unroll
              for (int u2 = 0; u2 < 4; u2++) {
                const int32_t j = j0 + u2 * TEAM;
                r[u2] = (j < n) ? rr[j] : -1;
              }
#pragma unroll
              for (int u2 = 0; u2 < 4; u2++)
#pragma unroll
                for (int c = 0; c < 4; c++) x[u2][c] = (r[u2] >= 0) ? __ldg(colp[c] + r[u2]) : 0.0;
#pragma unroll
              for (int u2 = 0; u2 < 4; u2++) {
                if (r[u2] >= 0) {
#pragma unroll
                  for (int c = 0; c < 4; c++) {
                    if (x[u2][c] < mn[c]) mn[c] = x[u2][c];
                    if (x[u2][c] > mx[c]) mx[c] = x[u2][c];
                    nanm |= (uint32_t)(x[u2][c] != x[u2][c]) << c;
                  }
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
                scr[wit * 12 + c] = mn[c];
                scr[wit * 12 + 4 + c] = mx[c];
              }
              s_redi[wit] = (int32_t)nanm;
            }
          }
          __syncthreads();
          if (tid < 4 && g0 + tid < nb && s_feat[g0 + tid] >= 0) {
            const int c = tid, ci = g0 + tid;
            double a = 1.7976931348623157e308, bq = -1.7976931348623157e308;
            int has_nan = 0;
            for (int w2 = 0; w2 < NWP; w2++) {
              const double v1 = scr[w2 * 12 + c], v2 = scr[w2 * 12 + 4 + c];
              if (v1 < a) a = v1;
              if (v2 > bq) bq = v2;
              has_nan |= (s_redi[w2] >> c) & 1;
            }
            if (bq <= a && !has_nan) {  // pkg:236
              s_flags[ci] |= CF_CONST;
            } else {
              s_cut[ci] = ET_ADD(a, ET_MUL(ET_SUB(bq, a), s_u[ci]));  // nextDouble(min, max), pkg:240
              if (has_nan) s_flags[ci] |= CF_NAN;
            }
          }
          __syncthreads();
          double cut[4];
          int any_nan = 0;
#pragma unroll
          for (int c = 0; c < 4; c++) {
            const int ci = min(g0 + c, nb - 1);
            cut[c] = s_cut[ci];
            any_nan |= (g0 + c < nb) && (s_flags[ci] & CF_NAN) && !(s_flags[ci] & CF_CONST);
          }
          int32_t keep_n = 0;  // thread c < 4: the x < cut side of candidate c
          double keep_S = 0.0, keep_Q = 0.0, res_sn = NAN, res_sl = NAN;
          int32_t nin_n = 0, nin_l = 0;
          for (int sweep = 0; sweep < (any_nan ? 2 : 1); sweep++) {
            int32_t cnt[4];
            double S[4], Q[4];
#pragma unroll
            for (int c = 0; c < 4; c++) {
              cnt[c] = 0;
              S[c] = 0.0;
              Q[c] = 0.0;
            }
            for (int32_t j0 = tid; j0 < n; j0 += 4 * TEAM) {
              int32_t r[4];
              double x[4][4], yd[4];
#pragma unroll
              for (int u2 = 0; u2 < 4; u2++) {
                const int32_t j = j0 + u2 * TEAM;
                r[u2] = (j < n) ? rr[j] : -1;
                yd[u2] = (j < n) ? ET_SUB(yy[j], reg_mu) : 0.0;
              }
#pragma unroll
              for (int u2 = 0; u2 < 4; u2++)
#pragma unroll
                for (int c = 0; c < 4; c++) x[u2][c] = (r[u2] >= 0) ? __ldg(colp[c] + r[u2]) : 0.0;
#pragma unroll
              for (int u2 = 0; u2 < 4; u2++) {
                const double yd2 = ET_MUL(yd[u2], yd[u2]);
#pragma unroll
                for (int c = 0; c < 4; c++) {
                  const bool in = (r[u2] >= 0) && (sweep ? (x[u2][c] != x[u2][c]) : (x[u2][c] < cut[c]));
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
                scr[wit * 12 + c] = S[c];
                scr[wit * 12 + 4 + c] = Q[c];
                s_redi[wit * 4 + c] = cnt[c];
              }
            }
            __syncthreads();
            if (tid < 4) {
              const int c = tid;
              int32_t ni = 0;
              double Si = 0.0, Qi = 0.0;
              for (int w2 = 0; w2 < NWP; w2++) {
                ni += s_redi[w2 * 4 + c];
                Si = ET_ADD(Si, scr[w2 * 12 + c]);
                Qi = ET_ADD(Qi, scr[w2 * 12 + 4 + c]);
              }
              if (sweep == 0) {
                keep_n = ni;
                keep_S = Si;
                keep_Q = Qi;
                nin_n = ni;
                res_sn = var_reduction_moments(n, reg_S, reg_Q, total, ni, Si, Qi);
              } else {
                nin_l = keep_n + ni;
                res_sl = var_reduction_moments(n, reg_S, reg_Q, total, nin_l, ET_ADD(keep_S, Si), ET_ADD(keep_Q, Qi));
              }
            }
          }
          if (tid < 4 && g0 + tid < nb && s_feat[g0 + tid] >= 0 && !(s_flags[g0 + tid] & CF_CONST)) {
            const int ci = g0 + tid;
            const double sn = res_sn, sl = (s_flags[ci] & CF_NAN) ? res_sl : NAN;
            const bool mil = !(sl != sl) && (sl > sn || (sn != sn));  // pkg:272-275
            s_score[ci] = mil ? sl : sn;
            s_nleft[ci] = mil ? nin_l : nin_n;
            if (mil) s_flags[ci] |= CF_MIL;
          }
        }
      } else if (CODED) {
        // ---- phase 1 (byte-coded table): the team streams the node's samples ONCE per pass for the whole
        //      batch.  A thread owns a sample and reads its byte in every candidate's column (a warp reads
        //      32 nearby bytes per column); per-candidate min / max live in packed bytes (4 candidates per
        //      register), so one pass and one reduction serve the whole batch.
        if (tid < 32) {
          const int32_t f = (tid < nb) ? s_feat[tid] : -1;
          s_coloff[tid] = (int64_t)(f >= 0 ? f : 0) * p.ldc;
          s_Kb[tid] = (f >= 0 && p.coff[f] == 0) ? 1 : 0;  // wide code - 1 = byte - K (mod 256)
        }
        __syncthreads();
        const int ng = (nb + 3) >> 2;
        const uint32_t *s_K4 = reinterpret_cast<const uint32_t *>(s_Kb);
        // (slots of the last group past nb read column 0: harmless, never consumed)
        if (ng <= 2)
          coded_pass1<2, TEAM>(p.C8, s_coloff, s_K4, rr, n, tid, s_cred, wit, lane, s_park);
        else if (ng <= 4)
          coded_pass1<4, TEAM>(p.C8, s_coloff, s_K4, rr, n, tid, s_cred, wit, lane, s_park);
        else
          coded_pass1<8, TEAM>(p.C8, s_coloff, s_K4, rr, n, tid, s_cred, wit, lane, s_park);
        __syncthreads();
        // ---- per candidate: decode min / max, constant test, cutpoint, code threshold
        if (wit == 0) {
          bool nan_c = false;
          const int c = lane;
          const int32_t f = (c < nb) ? s_feat[c] : -1;
          s_enb[c] = 0;
          s_nanb[c] = 0;
          s_thrb[c] = 0;
          if (f >= 0) {
            const int g = c >> 2, sh = 8 * (c & 3);
            uint32_t mnt = 255u, mxb = 0u, mnb = 255u;
            for (int w2 = 0; w2 < TEAM / 32; w2++) {
              mnt = min(mnt, (s_cred[w2 * 24 + g] >> sh) & 255u);
              mxb = max(mxb, (s_cred[w2 * 24 + 8 + g] >> sh) & 255u);
              mnb = min(mnb, (s_cred[w2 * 24 + 16 + g] >> sh) & 255u);
            }
            const uint32_t K = s_Kb[c], wmax = mxb + (1u - K);  // largest wide code (0 = only NaNs)
            const bool has_nan = (K == 1u) && (mnb == 0u);
            double mn = 1.7976931348623157e308, mx = -1.7976931348623157e308;  // pkg:35-36
            const double *dc8 = p.dict + (int64_t)f * 256;
            if (wmax != 0u) {
              mn = __ldg(dc8 + mnt);
              mx = __ldg(dc8 + (wmax - 1u));
            }
            if (mx <= mn && !has_nan) {  // pkg:236
              s_flags[c] |= CF_CONST;
            } else {
              const double cut = ET_ADD(mn, ET_MUL(ET_SUB(mx, mn), s_u[c]));  // nextDouble(min, max), pkg:240
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
              s_cut[c] = cut;
              s_thr[c] = (int32_t)thr;
              s_thrb[c] = (uint8_t)(thr > 0u ? thr - 1u : 0u);
              s_enb[c] = thr > 0u ? 0xff : 0;
              if (has_nan) {
                s_flags[c] |= CF_NAN;
                s_nanb[c] = 0xff;
                nan_c = true;
              }
            }
          }
          const bool any_nan = __any_sync(0xffffffffu, nan_c);
          if (lane == 0) s_misc[3] = any_nan ? 1 : 0;
        }
        __syncthreads();
        // ---- pass 2: side histograms (second sweep only if a candidate's column holds NaNs in this node)
        const int nsweep = s_misc[3] ? 2 : 1;
        auto lab = [&](int32_t j) -> int32_t { return LAB(j); };
        for (int sweep = 0; sweep < nsweep; sweep++) {
          const uint32_t *s_t4 = reinterpret_cast<const uint32_t *>(s_thrb);
          const uint32_t *s_e4 = reinterpret_cast<const uint32_t *>(sweep ? s_nanb : s_enb);
          const int hoff = sweep ? C : 0;
          if (ng <= 2)
            coded_pass2<2, TEAM>(p.C8, s_coloff, s_K4, s_t4, s_e4, sweep, rr, lab, n, C, wit, lane, s_hist, L.hs, hoff, nb,
                                 s_park);
          else if (ng <= 4)
            coded_pass2<4, TEAM>(p.C8, s_coloff, s_K4, s_t4, s_e4, sweep, rr, lab, n, C, wit, lane, s_hist, L.hs, hoff, nb,
                                 s_park);
          else
            coded_pass2<8, TEAM>(p.C8, s_coloff, s_K4, s_t4, s_e4, sweep, rr, lab, n, C, wit, lane, s_hist, L.hs, hoff, nb,
                                 s_park);
        }
      } else {
      // ---- phase 1: the whole team on the samples of one candidate at a time
      for (int oi = 0; oi < 32; oi++) {
        const int c = s_ord[oi];
        if (c < 0) break;  // inactive slots sort last
        const int32_t f = s_feat[c];
        const double *col = p.X + (int64_t)f * p.ld;
        double mn = 1.7976931348623157e308, mx = -1.7976931348623157e308;  // pkg:35-36
        int has_nan = 0;
        if (WARP) {
          // one gather per sample: the values are parked in shared memory for the second pass
#pragma unroll 4
          for (int v = 0; v < nv; v++) {
            const int j = v * 32 + lane;
            if (j < n) {
              const double x = __ldg(col + s_rows[j]);
              s_xs[j] = x;
              if (x < mn) mn = x;
              if (x > mx) mx = x;
              has_nan |= (x != x);
            }
          }
        } else {
          for (int32_t j0 = 0; j0 < n; j0 += 4 * TEAM) {
            int32_t r4[4];
            double x4[4];
#pragma unroll
            for (int u2 = 0; u2 < 4; u2++) {
              const int32_t j = j0 + u2 * TEAM + tid;
              r4[u2] = (j < n) ? rr[j] : -1;
            }
#pragma unroll
            for (int u2 = 0; u2 < 4; u2++) x4[u2] = (r4[u2] >= 0) ? __ldg(col + r4[u2]) : 0.0;
#pragma unroll
            for (int u2 = 0; u2 < 4; u2++) {
              if (r4[u2] >= 0) {
                const double x = x4[u2];
                if (staged) s_xs[j0 + u2 * TEAM + tid] = x;
                if (x < mn) mn = x;
                if (x > mx) mx = x;
                has_nan |= (x != x);
              }
            }
          }
        }
        team_minmax<TEAM>(mn, mx, has_nan, s_redd, s_redi);
        if (mx <= mn && !has_nan) {  // pkg:236
          if (tid == 0) s_flags[c] |= CF_CONST;
          continue;
        }
        const double cut = ET_ADD(mn, ET_MUL(ET_SUB(mx, mn), s_u[c]));  // nextDouble(min, max), pkg:240
        if (tid == 0) {
          s_cut[c] = cut;
          if (has_nan) s_flags[c] |= CF_NAN;
        }
        if (TASK == TASK_CLS) {
          int32_t *hl = s_hist + c * L.hs, *hn = hl + C;
          if (WARP && L.use_cm) {
            int32_t al = 0, an = 0;  // lane == class (first 32 classes in registers)
            for (int v = 0; v < nv; v++) {
              const int j = v * 32 + lane;
              const double x = (j < n) ? s_xs[j] : cut;
              const uint32_t blt = __ballot_sync(0xffffffffu, x < cut);
              const uint32_t bnan = has_nan ? __ballot_sync(0xffffffffu, x != x) : 0u;
              if (C <= 32) {
                if (lane < C) {
                  const uint32_t m = s_cm[v * C + lane];
                  al += __popc(blt & m);
                  an += __popc(bnan & m);
                }
              } else {
                for (int cc = lane; cc < C; cc += 32) {
                  const uint32_t m = s_cm[v * C + cc];
                  hl[cc] += __popc(blt & m);
                  hn[cc] += __popc(bnan & m);
                }
              }
            }
            if (C <= 32 && lane < C) {
              hl[lane] = al;
              hn[lane] = an;
            }
          } else if (WARP) {
            for (int j = lane; j < n; j += 32) {
              const double x = s_xs[j];
              if (x < cut)
                atomicAdd(&hl[s_lab[j]], 1);
              else if (x != x)
                atomicAdd(&hn[s_lab[j]], 1);
            }
          } else if (C <= 16 && !has_nan) {
            // per-thread packed 8-bit counters (one field per class), flushed before they can overflow
            unsigned long long a0 = 0ull, a1 = 0ull;
            int it = 0;
            for (int32_t j0 = 0; j0 < n; j0 += 4 * TEAM) {
              int32_t r4[4], c4[4];
              double x4[4];
#pragma unroll
              for (int u2 = 0; u2 < 4; u2++) {
                const int32_t j = j0 + u2 * TEAM + tid;
                r4[u2] = (j < n) ? (staged ? j : rr[j]) : -1;
                c4[u2] = (j < n) ? LAB(j) : 0;
              }
#pragma unroll
              for (int u2 = 0; u2 < 4; u2++)
                x4[u2] = (r4[u2] >= 0) ? (staged ? s_xs[r4[u2]] : __ldg(col + r4[u2])) : cut;
#pragma unroll
              for (int u2 = 0; u2 < 4; u2++) {
                const unsigned long long inc = (x4[u2] < cut) ? 1ull : 0ull;
                if (c4[u2] < 8)
                  a0 += inc << (8 * c4[u2]);
                else
                  a1 += inc << (8 * (c4[u2] - 8));
              }
              it += 4;
              if (it >= 252 || j0 + 4 * TEAM >= n) {  // uniform across the team
                for (int cc = 0; cc < C; cc++) {
                  const unsigned v = (unsigned)(((cc < 8) ? (a0 >> (8 * cc)) : (a1 >> (8 * (cc - 8)))) & 0xffull);
                  const unsigned tot = __reduce_add_sync(0xffffffffu, v);
                  if (lane == 0 && tot) atomicAdd(&hl[cc], (int32_t)tot);
                }
                a0 = 0ull;
                a1 = 0ull;
                it = 0;
              }
            }
          } else if (C <= 32) {
            int32_t al = 0, an = 0;
            for (int32_t j0 = wit * 32; j0 < n; j0 += TEAM) {
              const int32_t j = j0 + lane;
              bool lt = false, isn = false;
              int32_t cls = -1;
              if (j < n) {
                const double x = staged ? s_xs[j] : __ldg(col + rr[j]);
                cls = LAB(j);
                lt = x < cut;
                isn = x != x;
              }
              for (int cc = 0; cc < C; cc++) {
                const uint32_t bl = __ballot_sync(0xffffffffu, lt && cls == cc);
                if (lane == cc) al += __popc(bl);
              }
              if (has_nan) {
                for (int cc = 0; cc < C; cc++) {
                  const uint32_t bn = __ballot_sync(0xffffffffu, isn && cls == cc);
                  if (lane == cc) an += __popc(bn);
                }
              }
            }
            if (lane < C) {
              if (al) atomicAdd(&hl[lane], al);
              if (an) atomicAdd(&hn[lane], an);
            }
          } else {
            for (int32_t j = tid; j < n; j += TEAM) {
              const double x = staged ? s_xs[j] : __ldg(col + rr[j]);
              if (x < cut)
                atomicAdd(&hl[LAB(j)], 1);
              else if (x != x)
                atomicAdd(&hn[LAB(j)], 1);
            }
          }
        } else {
          uint32_t *mlt = WARP ? (s_bits + (size_t)c * 2 * BITS_W) : (g_bits + (size_t)c * 2 * words);
          uint32_t *mnan = mlt + (WARP ? BITS_W : words);
          for (int32_t j0 = wit * 32; j0 < n; j0 += TEAM) {
            const int32_t j = j0 + lane;
            bool lt = false, isn = false;
            if (j < n) {
              const double x = (WARP || staged) ? s_xs[j] : __ldg(col + rr[j]);
              lt = x < cut;
              isn = x != x;
            }
            const uint32_t blt = __ballot_sync(0xffffffffu, lt), bnan = __ballot_sync(0xffffffffu, isn);
            if (lane == 0) {
              mlt[j0 >> 5] = blt;
              mnan[j0 >> 5] = bnan;
            }
          }
        }
      }
      }
      team_sync<TEAM>();
      // ---- phase 2: one thread per candidate evaluates the reference's score expression exactly
      if (!REGPAR && tid < nb && s_feat[tid] >= 0 && !(s_flags[tid] & CF_CONST)) {
        const int c = tid;
        const bool has_nan = (s_flags[c] & CF_NAN) != 0;
        double sn, sl = NAN;
        int32_t nin_n = 0, nin_l = 0;
        if (TASK == TASK_CLS) {
          const int32_t *hl = s_hist + c * L.hs, *hn = hl + C;
          sn = gini_score_int(s_hnode, hl, hn, false, C, n, total, &nin_n);
          if (has_nan) sl = gini_score_int(s_hnode, hl, hn, true, C, n, total, &nin_l);
        } else {
          const uint32_t *mlt = WARP ? (s_bits + (size_t)c * 2 * BITS_W) : (g_bits + (size_t)c * 2 * words);
          const uint32_t *mnan = mlt + (WARP ? BITS_W : words);
          if (TASK == TASK_REG) {
            sn = var_reduction_seq(yy, n, mlt, mnan, false, total, &nin_n);
            if (has_nan) sl = var_reduction_seq(yy, n, mlt, mnan, true, total, &nin_l);
          } else {
            double *hin = s_wh + (size_t)c * 2 * C, *hout = hin + C;
            sn = gini_score_w_seq(ll, ww, n, mlt, mnan, false, C, total, nsum, hin, hout, &nin_n);
            if (has_nan) sl = gini_score_w_seq(ll, ww, n, mlt, mnan, true, C, total, nsum, hin, hout, &nin_l);
          }
        }
        // pkg:272-275
        const bool mil = !(sl != sl) && (sl > sn || (sn != sn));
        s_score[c] = mil ? sl : sn;
        s_nleft[c] = mil ? nin_l : nin_n;
        if (mil) s_flags[c] |= CF_MIL;
      }
      team_sync<TEAM>();
      // ---- consume the batch in draw order (every warp computes the same result; warp 0 of the
      //      team applies the side effects)
      {
        const bool act0 = lane < nb && s_feat[lane] >= 0;
        const int32_t fl = act0 ? s_flags[lane] : 0;
        const bool const0 = act0 && (fl & CF_CONST);
        const double s = (act0 && !const0) ? s_score[lane] : NAN;
        const bool counted0 = act0 && !const0 && !(s != s);
        // the reference stops drawing once k candidates have been scored: lanes past that point
        // were never examined
        const uint32_t m_cnt0 = __ballot_sync(0xffffffffu, counted0);
        const bool act = act0 && (p.replay || __popc(m_cnt0 & ((1u << lane) - 1u)) < p.k - visited);
        const bool is_const = act && const0;
        const bool is_nan = act && !const0 && (s != s);
        const bool counted = act && counted0;
        const uint32_t m_act = __ballot_sync(0xffffffffu, act);
        const uint32_t m_const = __ballot_sync(0xffffffffu, is_const);
        const uint32_t m_nan = __ballot_sync(0xffffffffu, is_nan);
        const uint32_t m_cnt = __ballot_sync(0xffffffffu, counted);
        if (p.replay) {
          const int exp = (fl >> 4) & 3;
          const bool bad = act && ((is_const && exp != 1) || (is_nan && exp != 3) || (counted && exp != 2));
          st_mismatch += __popc(__ballot_sync(0xffffffffu, bad));
        }
        // first maximum in lane order among the counted candidates (NaN never wins, pkg:277)
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
        if (REGPAR) {
          // runner-up over everything seen so far (for the ambiguity count)
          double b2 = (counted && lane != bl) ? s : -INFINITY;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) b2 = fmax(b2, __shfl_xor_sync(0xffffffffu, b2, o));
          if (bl < 32) {
            if (bs > best_score)
              second_score = fmax(best_score, fmax(second_score, b2));
            else
              second_score = fmax(second_score, bs);
          }
        }
        if (bl < 32 && bs > best_score) {
          best_score = bs;
          best_feature = s_feat[bl];
          best_cut = s_cut[bl];
          best_mil = (s_flags[bl] & CF_MIL) ? 1 : 0;
          best_nleft = s_nleft[bl];
          if (CODED) {
            best_thr = s_thr[bl];
            best_K = s_Kb[bl];
          }
          if (TASK == TASK_CLS && wit == 0) {
            const int32_t *hl = s_hist + bl * L.hs;
            for (int c = lane; c < C; c += 32) s_besthl[c] = hl[c] + (best_mil ? hl[C + c] : 0);
          }
        }
        if (!p.replay && wit == 0 && (is_const || is_nan)) {
          const int32_t f = s_feat[lane];
          atomicOr(&s_const[f >> 5], 1u << (f & 31));  // pkg:236-238, 283-285: inherited by the children
        }
        visited += __popc(m_cnt);
        nconst += __popc(m_const) + __popc(m_nan);
        st_draws += __popc(m_act);
        st_const += __popc(m_const);
        st_scored += __popc(m_cnt) + __popc(m_nan);
      }
      team_sync<TEAM>();
    }
  }

  // ---------------- finalize ----------------
  const bool make_leaf = leaf || best_feature < 0;  // pkg:293-296: visited == 0 || cut.isNaN  <=>  no best
  if (tid == 0) {
    if (!leaf) {
      atomicAdd(&p.cnt->st[ST_SROWS], (unsigned long long)n);
      atomicAdd(&p.cnt->st[ST_VMM], (unsigned long long)n * st_draws);
      atomicAdd(&p.cnt->st[ST_VSC], (unsigned long long)n * st_scored);
      atomicAdd(&p.cnt->st[ST_DRAWS], st_draws);
      atomicAdd(&p.cnt->st[ST_CONST], st_const);
      atomicAdd(&p.cnt->st[ST_SCORED], st_scored);
    }
    if (p.replay) {
      const bool trace_split = tn >= 0 && p.tr.left[tn] >= 0;
      if (trace_split == make_leaf) st_mismatch++;
      if (st_mismatch) atomicAdd(&p.cnt->st[ST_MISMATCH], st_mismatch);
    }
    if (REGPAR && !leaf) {
      atomicAdd(&p.cnt->st[ST_PARNODES], 1ull);
      if (best_feature >= 0 && second_score > -INFINITY &&
          best_score > second_score &&  // (an exact tie comes from identical partitions: first wins, like the reference)
          ET_SUB(best_score, second_score) <= 1e-9 * fmax(fabs(best_score), 1e-300))
        atomicAdd(&p.cnt->st[ST_AMBIG], 1ull);
    }
  }
  if (make_leaf) {
    if (tid == 0) {
      s_misc[2] = atomicAdd(&p.cnt->n_leaves, 1);
      p.o.feat[node] = -1;
      p.o.child[node] = s_misc[2];
      p.o.cut[node] = NAN;
      p.o.tree[node] = tree;
    }
    team_sync<TEAM>();
    double *lv = p.o.leaf_vals + (int64_t)s_misc[2] * lw;
    if (TASK == TASK_CLS) {
      const double inv = ET_DIV(1.0, (double)n);
      for (int c = tid; c < C; c += TEAM) lv[c] = repeat_add_dev(inv, s_hnode[c]);  // pkg:960-964
    } else if (TASK == TASK_CLSW) {
      for (int c = tid; c < C; c += TEAM) lv[c] = s_dist[c];
    } else {
      if (tid == 0) {
        if (REGPAR) {  // a large leaf is rare: its value is the reference's sequential mean (pkg:782), exactly
          double sum = 0.0;
          for (int32_t j = 0; j < n; j++) sum = ET_ADD(sum, yy[j]);
          leaf_mean = ET_DIV(sum, (double)n);
        }
        lv[0] = leaf_mean;
      }
    }
    return;
  }
  if (tid == 0) {
    const int32_t slot = atomicAdd(&p.cnt->next_f, 2);
    s_misc[2] = slot;
    const int32_t cl = p.node_base_next + slot;
    p.o.feat[node] = best_feature | (best_mil ? ET_MIL_BIT : 0);
    p.o.child[node] = cl;
    p.o.cut[node] = best_cut;
    p.o.tree[node] = tree;
    const int32_t nl = best_nleft;
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
      const int32_t cn = side ? (n - nl) : nl;
      const int qc = size_class(p, cn);
      p.q_nxt[qc][atomicAdd(&p.cnt->q_count[qc], 1)] = s2;
      if (qc < p.sub_ncls) atomicAdd(&p.cnt->sub_rows, cn);
    }
    atomicAdd(&p.cnt->st[ST_PROWS], (unsigned long long)n);
  }
  team_sync<TEAM>();
  const int32_t slot = s_misc[2];
  if (TASK == TASK_CLS) {
    int32_t *hl = p.nxt.hist + (int64_t)slot * C, *hr = hl + C;
    for (int c = tid; c < C; c += TEAM) {
      hl[c] = s_besthl[c];
      hr[c] = s_hnode[c] - s_besthl[c];
    }
  }
  if (!p.replay) {
    uint32_t *ml = p.nxt.mask + (int64_t)slot * W, *mr = ml + W;
    for (int w = tid; w < W; w += TEAM) {
      const uint32_t v = s_const[w];
      ml[w] = v;
      mr[w] = v;
    }
  }
  // ---- stable partition of the node's segment (pkg:1024-1039)
  {
    const double *col = p.X + (int64_t)best_feature * p.ld;
    const bool mil = best_mil != 0;
    int32_t lpos = b, rpos = b + best_nleft;
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (int32_t j0 = 0; j0 < n; j0 += TEAM) {
      const int32_t j = j0 + tid;
      const bool valid = j < n;
      int32_t r = 0;
      bool left = false;
      if (valid) {
        r = rr[j];
        if (CODED) {
          const int32_t b8 = (int32_t)__ldg(p.C8 + (int64_t)best_feature * p.ldc + r);
          const bool isn = (best_K == 1) && (b8 == 0);
          left = isn ? mil : (((b8 - best_K) & 255) < best_thr);
        } else {
          const double x = __ldg(col + r);
          left = (x < best_cut) || (mil && (x != x));
        }
      }
      const uint32_t bv = __ballot_sync(0xffffffffu, valid);
      const uint32_t bl = __ballot_sync(0xffffffffu, left);
      const uint32_t br = bv & ~bl;
      int32_t lbase = lpos, rbase = rpos, ltot = __popc(bl), rtot = __popc(br);
      if (!WARP) {
        __syncthreads();
        if (lane == 0) {
          s_redi[wit] = ltot;
          s_redi[32 + wit] = rtot;
        }
        __syncthreads();
        ltot = 0;
        rtot = 0;
        for (int w = 0; w < TEAM / 32; w++) {
          const int32_t a = s_redi[w], c2 = s_redi[32 + w];
          if (w < wit) {
            lbase += a;
            rbase += c2;
          }
          ltot += a;
          rtot += c2;
        }
      }
      if (valid) {
        const int32_t dst = left ? lbase + __popc(bl & lt_mask) : rbase + __popc(br & lt_mask);
        p.idx_dst[base + dst] = r;
        if (TASK == TASK_REG) {
          p.yr_dst[base + dst] = yy[j];
        } else {
          p.yc_dst[base + dst] = LAB(j);
          if (TASK == TASK_CLSW) p.w_dst[base + dst] = ww[j];
        }
      }
      lpos += ltot;
      rpos += rtot;
    }
  }
}
#undef LAB

// ---- lane-per-candidate node kernel (n <= 32 * NW): one warp per node ---------------------------
// Every lane owns ONE candidate feature of the batch and walks the node's samples for it: gather
// once (parked in shared memory), min / max, cutpoint, side bitmask over the samples, exact score
// from the bitmask.  No cross-lane reductions at all; the winner's bitmask IS the partition.
// VT = double gathers FP64 values from X (NW == 1 only: 8 KB of parked values per warp);
// VT = uint8_t gathers the order-preserving byte codes of encode.cu (wide code 0 = NaN, r + 1 = dict[r]):
// min / max are integer, decoded through the dictionary, and `x < cut` is `code - 1 < thr` with
// thr = number of dictionary entries below the cutpoint -- bit-identical decisions on 1/8 of the bytes,
// and a node of up to 512 samples parks in 16 KB.  NW (32-sample words per node) is a launch
// parameter: one size class per NW in {1, 2, 4, 8, 16}, shared memory sized to the class.
constexpr int LANE_WARPS = 4;
#ifndef LANE_SMALL_CTAS
#define LANE_SMALL_CTAS 8
#endif

__host__ __device__ inline int lane_smem_bytes(int task, int C, int W, bool replay, int NW, int vbytes) {
  int o = 0;
  o += (task == TASK_CLS) ? 0 : 32 * NW * 8;      // s_y    regression target / weight, by position
  o += (task == TASK_REG) ? 0 : C * 8;            // s_dist
  o += 32 * NW * 32 * vbytes;                     // s_x    parked values [position][lane]
  o += 2 * NW * 32 * 4;                           // s_lt, s_nn  side bitmasks [word][lane]
  o += NW * 4;                                    // s_best winner's bitmask
  o += (task == TASK_REG) ? 0 : C * NW * 4;       // s_cm   per class, bitmask over the positions
  o += (task == TASK_REG) ? 0 : C * 4;            // s_hnode
  o += replay ? 0 : 2 * W * 4;                    // const / taken masks
  return ((o + 15) / 16) * 16;
}

// this lane's side bitmask word w: samples with x < cut, plus the NaN samples when they go left
#define LANE_IN(w) (s_lt[(w) * 32 + lane] | (nan_left ? s_nn[(w) * 32 + lane] : 0u))

// giniScore from a side bitmask (bit j = sample j goes left) and per-class sample bitmasks.
// Classes absent from the node contribute exactly +0.0 to both sums and are skipped; an empty side
// gives 0/0 = NaN exactly like the reference (pkg:1148-1157).
__device__ __noinline__ double gini_score_bits(const uint32_t *s_lt, const uint32_t *s_nn, bool nan_left, int lane,
                                                  const uint32_t *cm, const int32_t *hnode, int C, int32_t n, int nw,
                                                  int NW, double G) {
  int32_t cin_i = 0;
  for (int w = 0; w < nw; w++) cin_i += __popc(LANE_IN(w));
  if (cin_i == 0 || cin_i == n) return NAN;
  const double cin = (double)cin_i, cout = (double)(n - cin_i), N = (double)n;
  double sin_ = 0.0, sout = 0.0;
  for (int c = 0; c < C; c++) {
    const int32_t ht = hnode[c];
    if (ht == 0) continue;
    int32_t hi = 0;
    for (int w = 0; w < nw; w++) hi += __popc(cm[c * NW + w] & LANE_IN(w));
    const int32_t ho = ht - hi;
    const double pi = ET_DIV((double)hi, cin), po = ET_DIV((double)ho, cout);
    sin_ = ET_ADD(sin_, ET_MUL(pi, pi));
    sout = ET_ADD(sout, ET_MUL(po, po));
  }
  const double gin = ET_SUB(1.0, sin_), gout = ET_SUB(1.0, sout);
  return ET_SUB(ET_SUB(G, ET_DIV(ET_MUL(gin, cin), N)), ET_DIV(ET_MUL(gout, cout), N));
}

// weighted giniScore (pkg:1132-1157): per-class and per-side sums in subset order.  Each of the
// reference's accumulators only ever sees its own samples, so walking the samples class by class
// (in subset order inside a class) performs the same additions in the same order.
__device__ __noinline__ double gini_score_w_bits(const uint32_t *s_lt, const uint32_t *s_nn, bool nan_left, int lane,
                                                    const uint32_t *cm, const double *wgt, int C, int32_t n, int nw,
                                                    int NW, double G, double N) {
  double cin = 0.0, cout = 0.0;
  for (int v = 0; v < nw; v++) {
    const uint32_t in = LANE_IN(v);
    const int cnt = min(32, n - v * 32);
    for (int j = 0; j < cnt; j++) {
      if ((in >> j) & 1u)
        cin = ET_ADD(cin, wgt[v * 32 + j]);
      else
        cout = ET_ADD(cout, wgt[v * 32 + j]);
    }
  }
  double sin_ = 0.0, sout = 0.0;
  for (int c = 0; c < C; c++) {
    double hi = 0.0, ho = 0.0;
    for (int v = 0; v < nw; v++) {
      const uint32_t in = LANE_IN(v);
      uint32_t m = cm[c * NW + v];
      while (m) {
        const int j = __ffs(m) - 1;
        m &= m - 1;
        if ((in >> j) & 1u)
          hi = ET_ADD(hi, wgt[v * 32 + j]);
        else
          ho = ET_ADD(ho, wgt[v * 32 + j]);
      }
    }
    const double pi = ET_DIV(hi, cin), po = ET_DIV(ho, cout);
    sin_ = ET_ADD(sin_, ET_MUL(pi, pi));
    sout = ET_ADD(sout, ET_MUL(po, po));
  }
  const double gin = ET_SUB(1.0, sin_), gout = ET_SUB(1.0, sout);
  return ET_SUB(ET_SUB(G, ET_DIV(ET_MUL(gin, cin), N)), ET_DIV(ET_MUL(gout, cout), N));
}

// computeVarianceReduction (pkg:1196-1218) from a side bitmask, sequential in subset order
__device__ __noinline__ double var_reduction_bits(const uint32_t *s_lt, const uint32_t *s_nn, bool nan_left, int lane,
                                                     const double *y, int32_t n, int nw, double V) {
  double sin_ = 0.0, sout = 0.0;
  int32_t nin = 0;
  for (int v = 0; v < nw; v++) {
    const uint32_t in = LANE_IN(v);
    nin += __popc(in);
    const int cnt = min(32, n - v * 32);
    for (int j = 0; j < cnt; j++) {
      if ((in >> j) & 1u)
        sin_ = ET_ADD(sin_, y[v * 32 + j]);
      else
        sout = ET_ADD(sout, y[v * 32 + j]);
    }
  }
  const int32_t nout = n - nin;
  const double dnin = (double)nin, dnout = (double)nout, dn = (double)n;
  const double min_ = ET_DIV(sin_, dnin), mout = ET_DIV(sout, dnout);
  double qin = 0.0, qout = 0.0;
  for (int v = 0; v < nw; v++) {
    const uint32_t in = LANE_IN(v);
    const int cnt = min(32, n - v * 32);
    for (int j = 0; j < cnt; j++) {
      if ((in >> j) & 1u) {
        const double dl = ET_SUB(y[v * 32 + j], min_);
        qin = ET_ADD(qin, ET_MUL(dl, dl));
      } else {
        const double dl = ET_SUB(y[v * 32 + j], mout);
        qout = ET_ADD(qout, ET_MUL(dl, dl));
      }
    }
  }
  const double svin = nin < 1 ? NAN : (nin == 1 ? 0.0 : ET_DIV(qin, ET_SUB(dnin, 1.0)));
  const double svout = nout < 1 ? NAN : (nout == 1 ? 0.0 : ET_DIV(qout, ET_SUB(dnout, 1.0)));
  const double vin = (nin == 1) ? 0.0 : ET_DIV(ET_MUL(svin, ET_SUB(dnin, 1.0)), dnin);
  const double vout = (nout == 1) ? 0.0 : ET_DIV(ET_MUL(svout, ET_SUB(dnout, 1.0)), dnout);
  const double a = ET_MUL(ET_DIV(dnin, dn), vin);
  const double bq = ET_MUL(ET_DIV(dnout, dn), vout);
  return ET_DIV(ET_SUB(ET_SUB(V, a), bq), V);
}

// SMALL: the classes of up to 64 samples run with a tighter register budget (more resident warps; these nodes
// are dominated by fixed per-batch latency), the larger classes are shared-memory bound anyway.
template <int TASK, typename VT, bool SMALL>
__global__ void __launch_bounds__(32 * LANE_WARPS, SMALL ? LANE_SMALL_CTAS : 4) k_lane(P p, int32_t qcount, int qi, int NW) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr bool CODED = (sizeof(VT) != 8);
  constexpr uint32_t FULL = 0xffffffffu;
  const int tic = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * LANE_WARPS + tic;
  if (q >= qcount) return;
  const int C = p.C, W = p.W;
  unsigned char *sm = smem_raw + (size_t)tic * lane_smem_bytes(TASK, C, W, p.replay != 0, NW, (int)sizeof(VT));
  double *s_y = reinterpret_cast<double *>(sm);
  double *s_dist = s_y + ((TASK == TASK_CLS) ? 0 : 32 * NW);
  VT *s_x = reinterpret_cast<VT *>(s_dist + ((TASK == TASK_REG) ? 0 : C));
  uint32_t *s_lt = reinterpret_cast<uint32_t *>(s_x + 32 * NW * 32);
  uint32_t *s_nn = s_lt + NW * 32;
  uint32_t *s_best = s_nn + NW * 32;
  uint32_t *s_cm = s_best + NW;
  int32_t *s_hnode = reinterpret_cast<int32_t *>(s_cm + ((TASK == TASK_REG) ? 0 : C * NW));
  uint32_t *s_const = reinterpret_cast<uint32_t *>(s_hnode + ((TASK == TASK_REG) ? 0 : C)), *s_taken = s_const + W;

  const int i = p.q_cur[qi][q];
  const int32_t tree = p.cur.tree[i], b = p.cur.begin[i], e = p.cur.end[i], n = e - b;
  const int32_t node = p.cur.node[i], depth = p.cur.depth[i];
  const int64_t tn = p.cur.trace[i];
  const uint64_t key = p.cur.key[i];
  const int64_t base = (int64_t)tree * p.n;
  const int lw = (TASK == TASK_REG) ? 1 : C;
  const int nw = (n + 31) >> 5;
  const int32_t *idx = p.idx_src + base + b;

  // ---------------- the node's labels / targets (position j = w * 32 + lane) ----------------
  if (TASK != TASK_REG) {
    for (int t = lane; t < C * NW; t += 32) s_cm[t] = 0u;
    __syncwarp();
    const int32_t *yc = p.yc_src + base + b;
    for (int w = 0; w < nw; w++) {
      const int j = w * 32 + lane;
      const bool has = j < n;
      const int32_t cls = has ? yc[j] : -1;
      const uint32_t grp = __match_any_sync(FULL, cls);
      if (has && lane == __ffs(grp) - 1) s_cm[cls * NW + w] = grp;
    }
  }
  if (TASK == TASK_REG) {
    const double *yr = p.yr_src + base + b;
    for (int j = lane; j < n; j += 32) s_y[j] = yr[j];
  }
  if (TASK == TASK_CLSW) {
    const double *wr = p.w_src + base + b;
    for (int j = lane; j < n; j += 32) s_y[j] = wr[j];
  }
  __syncwarp();

  // ---------------- stop rules + node totals ----------------
  bool leaf;
  double total = 0.0, nsum = (double)n, leaf_mean = 0.0;
  if (TASK != TASK_REG) {
    bool pure_l = false;
    for (int c = lane; c < C; c += 32) {
      int32_t h = 0;
      for (int w = 0; w < nw; w++) h += __popc(s_cm[c * NW + w]);
      s_hnode[c] = h;
      pure_l |= (h == n);
    }
    const bool pure = __any_sync(FULL, pure_l);  // all targets in the subset equal (weights ignored)
    leaf = (p.n_table < p.n_min) || (depth >= p.max_depth) || pure;  // pkg:993-994
    __syncwarp();
  }
  if (TASK == TASK_CLS) {
    if (!leaf) {
      // giniImpurity with the reference's repeated `+= 1/s` distribution (pkg:905-911, 1160-1180)
      const double inv = ET_DIV(1.0, (double)n);
      for (int c = lane; c < C; c += 32) s_dist[c] = repeat_add_dev(inv, s_hnode[c]);
      __syncwarp();
      double s = 0.0;
      for (int c = 0; c < C; c++) s = ET_ADD(s, ET_MUL(s_dist[c], s_dist[c]));
      total = ET_SUB(1.0, s);
    }
  } else if (TASK == TASK_REG) {
    const double head = s_y[0];
    bool uni_l = true;
    for (int j = lane; j < n; j += 32) uni_l &= !(s_y[j] != head);
    const bool uni = __all_sync(FULL, uni_l);
    leaf = (n < p.n_min) || (depth >= p.max_depth) || uni;  // pkg:813-814
    // mean2 (pkg:782) and varianceNoSplit (pkg:436-437), sequential in subset order
    double sum = 0.0;
    for (int j = 0; j < n; j++) sum = ET_ADD(sum, s_y[j]);
    const double dn = (double)n;
    leaf_mean = ET_DIV(sum, dn);
    if (!leaf) {
      double var = 0.0;
      if (n > 1) {
        double qq = 0.0;
        for (int j = 0; j < n; j++) {
          const double dl = ET_SUB(s_y[j], leaf_mean);
          qq = ET_ADD(qq, ET_MUL(dl, dl));
        }
        var = ET_DIV(qq, ET_SUB(dn, 1.0));
      }
      total = ET_DIV(ET_MUL(var, ET_SUB(dn, 1.0)), dn);
    }
  } else {
    // weighted distribution (pkg:913-927): per-class sums and the total, each in subset order
    double s = 0.0;
    for (int j = 0; j < n; j++) s = ET_ADD(s, s_y[j]);
    for (int c = lane; c < C; c += 32) {
      double a = 0.0;
      for (int w = 0; w < nw; w++) {
        uint32_t m = s_cm[c * NW + w];
        while (m) {
          const int j = __ffs(m) - 1;
          m &= m - 1;
          a = ET_ADD(a, s_y[w * 32 + j]);
        }
      }
      s_dist[c] = ET_DIV(a, s);
    }
    __syncwarp();
    double sq = 0.0;
    for (int c = 0; c < C; c++) sq = ET_ADD(sq, ET_MUL(s_dist[c], s_dist[c]));
    total = ET_SUB(1.0, sq);
    nsum = s;
  }

  // ---------------- split search ----------------
  int32_t visited = 0, nconst = 0, best_feature = -1, best_mil = 0;
  double best_score = -INFINITY, best_cut = NAN;
  unsigned long long st_draws = 0, st_const = 0, st_scored = 0, st_mismatch = 0;
  if (!leaf) {
    int32_t dc = 0, tpos = 0, tcnt = 0;
    int64_t tb = 0;
    if (p.replay) {
      if (tn >= 0) {
        tb = p.tr.cand_begin[tn];
        tcnt = p.tr.cand_count[tn];
      }
    } else {
      int nc = 0;
      for (int w = lane; w < W; w += 32) {
        const uint32_t m = p.cur.mask[(int64_t)i * W + w];
        s_const[w] = m;
        s_taken[w] = m;
        nc += __popc(m);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) nc += __shfl_xor_sync(FULL, nc, o);
      nconst = nc - (W * 32 - p.d);
      __syncwarp();
    }
    // Small nodes of a byte-coded table (free-running): which features vary over the node's rows is read off the
    // rows themselves in the row-major copy (n x 784 contiguous bytes: OR of the XORs with the first row, four
    // features per word; a NaN byte makes the feature vary like hasMissing does, pkg:236).  Candidates are then
    // drawn from the varying features only -- the scored candidates of the reference are a uniform sample without
    // replacement of exactly that set (draws that hit a constant feature are discarded, pkg:236-239), so the split
    // has the same distribution, and no gather pass is spent on constant features (48 % of the draws before).
    bool use_nc = false;
    if (SMALL && CODED && !p.replay && p.R8 != nullptr && p.r8_stride <= 1024 && n <= p.nc_max) {
      use_nc = true;
      const int nword = p.r8_stride >> 2;
      const uint32_t *cof4 = reinterpret_cast<const uint32_t *>(p.coff);
      for (int h2 = 0; h2 < 2; h2++) {  // table words lane + 32 i, i = 4 h2 .. 4 h2 + 3
        if (h2 * 128 >= nword) {
          for (int w0 = h2 * 16 + lane; w0 < W; w0 += 32) s_taken[w0] = 0xffffffffu;
          continue;
        }
        uint32_t first[4], acc[4];
        {
          const uint32_t *rp = reinterpret_cast<const uint32_t *>(p.R8 + (int64_t)idx[0] * p.r8_stride);
#pragma unroll
          for (int i4 = 0; i4 < 4; i4++) {
            const int tw = (h2 * 4 + i4) * 32 + lane;
            first[i4] = (tw < nword) ? __ldg(rp + tw) : 0u;
            acc[i4] = 0u;
          }
        }
        for (int w = 0; w < nw; w++) {
          const int j0 = w << 5, cnt = min(32, n - j0);
          const int32_t row = (lane < cnt) ? idx[j0 + lane] : 0;
          for (int jj = (w == 0) ? 1 : 0; jj < cnt; jj++) {
            const uint32_t *rp =
                reinterpret_cast<const uint32_t *>(p.R8 + (int64_t)__shfl_sync(FULL, row, jj) * p.r8_stride);
#pragma unroll
            for (int i4 = 0; i4 < 4; i4++) {
              const int tw = (h2 * 4 + i4) * 32 + lane;
              if (tw < nword) acc[i4] |= __ldg(rp + tw) ^ first[i4];
            }
          }
        }
#pragma unroll
        for (int i4 = 0; i4 < 4; i4++) {
          const int tw = (h2 * 4 + i4) * 32 + lane;
          uint32_t bits = 0u;
          if (tw < nword) {
            const uint32_t k4 = __ldg(cof4 + tw);                                  // 0 = the column holds NaNs
            const uint32_t nan4 = __vcmpeq4(first[i4], 0u) & __vcmpeq4(k4, 0u);    // the first row is NaN there
            const uint32_t ne4 = __vcmpne4(acc[i4], 0u) | nan4;                    // 0xff per varying feature
            bits = ((ne4 & 0x01010101u) * 0x01020408u) >> 24;                      // 4 bits, feature order
          }
          uint32_t word = bits << (4 * (lane & 7));
          word |= __shfl_xor_sync(FULL, word, 1);
          word |= __shfl_xor_sync(FULL, word, 2);
          word |= __shfl_xor_sync(FULL, word, 4);
          const int w0 = (h2 * 4 + i4) * 4 + (lane >> 3);
          if ((lane & 7) == 0 && w0 < W) s_taken[w0] = ~word;  // taken = not varying (padding included)
        }
      }
      __syncwarp();
      int nc = 0;
      for (int w = lane; w < W; w += 32) {
        const uint32_t m = s_taken[w];
        s_const[w] = m;  // every feature constant here is constant in the whole subtree
        nc += __popc(m);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) nc += __shfl_xor_sync(FULL, nc, o);
      nconst = nc - (W * 32 - p.d);
      __syncwarp();
    }
    for (;;) {
      int32_t nb;
      const int32_t avail = p.d - nconst - visited;
      if (p.replay) {
        nb = min(32, tcnt - tpos);
      } else if (use_nc) {
        nb = (min(p.k - visited, avail) > 0) ? 32 : 0;
      } else {
        // over-draw by the share of constant features expected among the draws: observed at this node once a
        // batch has been examined; before that, one in two if constants were found on the path from the root
        // (sparse tables) and none otherwise (continuous tables never waste a gather).  Candidates past the k-th
        // scored one are discarded unexamined below, like the reference which stops drawing there.
        const int32_t need = min(p.k - visited, avail);
        int32_t extra;
        if (st_draws > 0)
          extra = (st_draws > (unsigned long long)visited)
                      ? (int32_t)(((long long)need * (long long)(st_draws - (unsigned long long)visited)) / max(visited, 1)) + 2
                      : 0;
        else
          extra = (nconst > 0) ? need + 4 : 0;
        nb = (need > 0) ? min(32, min(avail, need + extra)) : 0;
      }
      if (nb <= 0) break;
      // ---- draw: lane == candidate
      int32_t f = -1;
      double u = 0.0;
      int expect = 0;
      if (p.replay) {
        if (lane < nb) {
          f = p.tr.cand_feature[tb + tpos + lane];
          u = p.tr.cand_u[tb + tpos + lane];
          expect = p.tr.cand_flag[tb + tpos + lane] + 1;
        }
        tpos += nb;
      } else if (use_nc) {
        // rounds of draws among the varying features not taken yet; duplicates inside a round lose to the earlier
        // lane (rejection keeps the sample uniform) and the next round fills up
        const int32_t want = min(32, p.k - visited);
        int32_t ncol = 0, left = avail;
        while (ncol < want && left > 0) {
          const int32_t nd = min(32, left);
          int32_t pick = -1 - lane;
          if (lane < nd)
            pick = rank_select_clear_fast(s_taken, W, (int32_t)__umul64hi(et_draw(key, (uint32_t)(dc + lane)), (uint64_t)left));
          dc += 32;
          const uint32_t same = __match_any_sync(FULL, pick);
          const bool drawn = lane < nd && lane == __ffs(same) - 1;
          const uint32_t m_dr = __ballot_sync(FULL, drawn);
          const int ord = __popc(m_dr & ((1u << lane) - 1u));
          const bool accp = drawn && ord < want - ncol;
          const uint32_t m_acc = __ballot_sync(FULL, accp);
          if (accp) {
            atomicOr(&s_taken[pick >> 5], 1u << (pick & 31));
            s_lt[ncol + ord] = (uint32_t)pick;  // (scratch: the side bitmasks are written after the draw)
          }
          const int nacc = __popc(m_acc);
          ncol += nacc;
          left -= nacc;
          __syncwarp();
        }
        if (ncol == 0) break;
        if (lane < ncol) {
          f = (int32_t)s_lt[lane];
          u = et_u01(et_draw(key, (uint32_t)(dc + lane)));
        }
        dc += 32;
        __syncwarp();
      } else {
        int32_t pick = -1 - lane;
        if (lane < nb) {
          const uint64_t r = et_draw(key, (uint32_t)(dc + 2 * lane));
          pick = rank_select_clear_fast(s_taken, W, (int32_t)__umul64hi(r, (uint64_t)avail));
          u = et_u01(et_draw(key, (uint32_t)(dc + 2 * lane + 1)));
        }
        const uint32_t same = __match_any_sync(FULL, pick);
        if (lane < nb && lane == __ffs(same) - 1) f = pick;
        __syncwarp();
        if (f >= 0) atomicOr(&s_taken[f >> 5], 1u << (f & 31));
        dc += 64;
      }
      const bool act0 = f >= 0;
      // ---- pass 1: gather the node's samples of this lane's feature; min / max / hasMissing (pkg:34-54)
      const VT *col = CODED ? reinterpret_cast<const VT *>(p.C8) + (int64_t)(act0 ? f : 0) * p.ldc
                            : reinterpret_cast<const VT *>(p.X) + (int64_t)(act0 ? f : 0) * p.ld;
      double mn = 1.7976931348623157e308, mx = -1.7976931348623157e308;  // pkg:35-36
      bool has_nan = false;
      // byte codes: K = 1 in a column that holds NaNs (stored byte 0 = NaN), else 0; t = byte - K is the
      // dictionary rank (NaN wraps to the top and never wins the min); the largest byte gives the max
      const uint32_t K = (CODED && act0 && __ldg(p.coff + f) == 0) ? 1u : 0u;
      const bool nan_cols = CODED ? (__any_sync(FULL, K != 0u) != 0) : true;  // can any lane's column hold a NaN?
      uint32_t mnt = 0xffffffffu, mxb = 0u;
      // (one gather per sample and lane; a full chunk keeps all 32 gathers of a lane in flight)
      // byte codes are parked four positions to a word, [position / 4][lane][position % 4]: pass 2 reads one word
      // per four samples.  The gather address is a 32-bit offset from the table base (host-checked: the coded
      // table is smaller than 4 GiB, else the 64-bit form is used).
      // Byte codes are gathered from the ROW-major copy when it exists: the 32 lanes of a gather read 32 features of
      // ONE row, i.e. 32 bytes inside one 784-byte row (<= 7 cache lines, ~18 sectors) instead of 32 sectors in 32
      // different columns (32 lines).  These nodes sit deep in the tree, where a column-major gather gets no
      // sector reuse between the rows of a node either.
      const bool rowmajor = CODED && p.R8 != nullptr;
      const uint8_t *c8base = CODED ? (rowmajor ? p.R8 : p.C8) : nullptr;
      const uint32_t rstride = rowmajor ? (uint32_t)p.r8_stride : 1u;
      const uint32_t coloff32 = (CODED && p.c8_small)
                                    ? (rowmajor ? (uint32_t)(act0 ? f : 0) : (uint32_t)((int64_t)(act0 ? f : 0) * p.ldc))
                                    : 0u;
      if (rowmajor) col = reinterpret_cast<const VT *>(p.R8) + (act0 ? f : 0);
      uint8_t *s_xb = reinterpret_cast<uint8_t *>(s_x);
      // (inactive lanes gather from column 0: no predicate, no branch, so all gathers of a chunk stay in flight)
      auto visit = [&](auto small_tab, int32_t rj, int pos) {
        if (CODED) {
          const uint32_t b8 = decltype(small_tab)::value
                                  ? (uint32_t)__ldg(c8base + (coloff32 + (uint32_t)rj * rstride))
                                  : (uint32_t)__ldg(reinterpret_cast<const uint8_t *>(col) + (int64_t)rj * rstride);
          s_xb[(((pos >> 2) * 32 + lane) << 2) + (pos & 3)] = (uint8_t)b8;
          mxb = max(mxb, b8);
          mnt = min(mnt, b8 - K);
        } else {
          const double x = act0 ? __ldg(reinterpret_cast<const double *>(col) + rj) : 0.0;
          s_x[pos * 32 + lane] = (VT)x;
          if (x < mn) mn = x;
          if (x > mx) mx = x;
          has_nan |= (x != x);
        }
      };
      auto pass1 = [&](auto small_tab) {
        for (int w = 0; w < nw; w++) {
          const int j0 = w << 5, cnt = min(32, n - j0);
          const int32_t row = (lane < cnt) ? idx[j0 + lane] : 0;
          if (!SMALL && cnt == 32) {  // (the small classes keep their code short: instruction fetch is their top stall)
#pragma unroll
            for (int jj = 0; jj < 32; jj++) visit(small_tab, __shfl_sync(FULL, row, jj), j0 + jj);
          } else {
#pragma unroll 4
            for (int jj = 0; jj < cnt; jj++) visit(small_tab, __shfl_sync(FULL, row, jj), j0 + jj);
          }
        }
      };
      if (CODED && p.c8_small)
        pass1(std::true_type{});
      else
        pass1(std::false_type{});
      uint32_t thr = 0u;
      const uint32_t wmax = CODED ? ((K == 1u) ? mxb : mxb + 1u) : 0u;  // largest wide code; 0 = only NaNs
      if (CODED) {
        if (act0 && wmax != 0u) {
          const double *dc8 = p.dict + (int64_t)f * 256;
          mn = __ldg(dc8 + mnt);
          mx = __ldg(dc8 + (wmax - 1u));
        }
      }
      // ---- pass 2: side bitmasks over the samples from the parked values.  The cutpoint only depends on
      //      min / max (pkg:240); for byte codes the NaN samples are found here (has_nan = any NaN bit).
      const double cut = ET_ADD(mn, ET_MUL(ET_SUB(mx, mn), u));  // pkg:240
      if (CODED) {
        if (act0 && wmax != 0u && !(mx <= mn)) {
          // thr = number of dictionary entries below the cutpoint; all of dict[0, mnt) are, none past wmax - 1
          const double *dc8 = p.dict + (int64_t)f * 256;
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
      }
      for (int w = 0; w < nw; w++) {
        const int j0 = w << 5, cnt = min(32, n - j0);
        uint32_t lt = 0u, nn = 0u;
        if (CODED) {
          // four samples per word: t = byte - K bytewise (NaN -> 255), left iff t <= thr - 1 (thr > 0)
          const uint32_t *s_xw = reinterpret_cast<const uint32_t *>(s_x) + (w * 8) * 32 + lane;
          const uint32_t K4 = K * 0x01010101u, t4 = (thr > 0u ? thr - 1u : 0u) * 0x01010101u;
          const uint32_t en4 = thr > 0u ? 0xffffffffu : 0u, kn4 = K ? 0xffffffffu : 0u;
          const int nq = (cnt + 3) >> 2;
#pragma unroll
          for (int q4 = 0; q4 < 8; q4++) {
            if (q4 < nq) {
              const uint32_t b4 = s_xw[q4 * 32];
              const uint32_t l4 = __vcmpleu4(__vsub4(b4, K4), t4) & en4;
              lt |= (((l4 & 0x01010101u) * 0x01020408u) >> 24) << (4 * q4);
              if (nan_cols) {
                const uint32_t n4 = __vcmpeq4(b4, 0u) & kn4;
                nn |= (((n4 & 0x01010101u) * 0x01020408u) >> 24) << (4 * q4);
              }
            }
          }
          const uint32_t valid = (cnt >= 32) ? FULL : ((1u << cnt) - 1u);  // (the last word may hold stale bytes)
          lt &= valid;
          nn &= valid;
        } else {
#pragma unroll 8
          for (int jj = 0; jj < cnt; jj++) {
            const double x = (double)s_x[(j0 + jj) * 32 + lane];
            lt |= (uint32_t)(x < cut) << jj;
            nn |= (uint32_t)(x != x) << jj;
          }
        }
        s_lt[w * 32 + lane] = lt;
        s_nn[w * 32 + lane] = nn;
        if (CODED) has_nan |= (nn != 0u);
      }
      const bool const0 = act0 && (mx <= mn) && !has_nan;  // pkg:236
      // ---- exact score of this lane's candidate (pkg:250-275)
      double s = NAN;
      bool mil = false;
      if (act0 && !const0) {
        double sn, sl = NAN;
        if (TASK == TASK_CLS)
          sn = gini_score_bits(s_lt, s_nn, false, lane, s_cm, s_hnode, C, n, nw, NW, total);
        else if (TASK == TASK_REG)
          sn = var_reduction_bits(s_lt, s_nn, false, lane, s_y, n, nw, total);
        else
          sn = gini_score_w_bits(s_lt, s_nn, false, lane, s_cm, s_y, C, n, nw, NW, total, nsum);
        if (has_nan) {
          if (TASK == TASK_CLS)
            sl = gini_score_bits(s_lt, s_nn, true, lane, s_cm, s_hnode, C, n, nw, NW, total);
          else if (TASK == TASK_REG)
            sl = var_reduction_bits(s_lt, s_nn, true, lane, s_y, n, nw, total);
          else
            sl = gini_score_w_bits(s_lt, s_nn, true, lane, s_cm, s_y, C, n, nw, NW, total, nsum);
        }
        mil = !(sl != sl) && (sl > sn || (sn != sn));  // pkg:272-275
        s = mil ? sl : sn;
      }
      // ---- consume the batch in draw (lane) order; the reference stops drawing once k candidates
      //      have been scored, so lanes past that point were never examined
      const bool counted0 = act0 && !const0 && !(s != s);
      const uint32_t m_cnt0 = __ballot_sync(FULL, counted0);
      const bool act = act0 && (p.replay || __popc(m_cnt0 & ((1u << lane) - 1u)) < p.k - visited);
      const bool is_const = act && const0;
      const bool is_nan = act && !const0 && (s != s);
      const bool counted = act && counted0;
      const uint32_t m_act = __ballot_sync(FULL, act);
      const uint32_t m_const = __ballot_sync(FULL, is_const);
      const uint32_t m_nan = __ballot_sync(FULL, is_nan);
      const uint32_t m_cnt = __ballot_sync(FULL, counted);
      if (p.replay) {
        const bool bad = act && ((is_const && expect != 1) || (is_nan && expect != 3) || (counted && expect != 2));
        st_mismatch += __popc(__ballot_sync(FULL, bad));
      }
      double bs = counted ? s : -INFINITY;
      int bl = counted ? lane : 64;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double os = __shfl_xor_sync(FULL, bs, o);
        const int ol = __shfl_xor_sync(FULL, bl, o);
        if (os > bs || (os == bs && ol < bl)) {
          bs = os;
          bl = ol;
        }
      }
      if (bl < 32 && bs > best_score) {  // strict >: the first best wins (pkg:277)
        best_score = bs;
        best_feature = __shfl_sync(FULL, f, bl);
        best_cut = __shfl_sync(FULL, cut, bl);
        best_mil = __shfl_sync(FULL, (int)mil, bl);
        __syncwarp();
        for (int w = lane; w < nw; w += 32) s_best[w] = s_lt[w * 32 + bl] | (best_mil ? s_nn[w * 32 + bl] : 0u);
      }
      if (!p.replay && (is_const || is_nan)) atomicOr(&s_const[f >> 5], 1u << (f & 31));
      visited += __popc(m_cnt);
      nconst += __popc(m_const) + __popc(m_nan);
      st_draws += __popc(m_act);
      st_const += __popc(m_const);
      st_scored += __popc(m_cnt) + __popc(m_nan);
      __syncwarp();
    }
  }

  // ---------------- finalize ----------------
  const bool make_leaf = leaf || best_feature < 0;
  if (lane == 0) {
    if (!leaf) {
      atomicAdd(&p.cnt->st[ST_SROWS], (unsigned long long)n);
      atomicAdd(&p.cnt->st[ST_VMM], (unsigned long long)n * st_draws);
      atomicAdd(&p.cnt->st[ST_VSC], (unsigned long long)n * st_scored);
      atomicAdd(&p.cnt->st[ST_DRAWS], st_draws);
      atomicAdd(&p.cnt->st[ST_CONST], st_const);
      atomicAdd(&p.cnt->st[ST_SCORED], st_scored);
    }
    if (p.replay) {
      const bool trace_split = tn >= 0 && p.tr.left[tn] >= 0;
      if (trace_split == make_leaf) st_mismatch++;
      if (st_mismatch) atomicAdd(&p.cnt->st[ST_MISMATCH], st_mismatch);
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
    ls = __shfl_sync(FULL, ls, 0);
    double *lv = p.o.leaf_vals + (int64_t)ls * lw;
    if (TASK == TASK_CLS) {
      const double inv = ET_DIV(1.0, (double)n);
      for (int c = lane; c < C; c += 32) lv[c] = repeat_add_dev(inv, s_hnode[c]);  // pkg:960-964
    } else if (TASK == TASK_CLSW) {
      for (int c = lane; c < C; c += 32) lv[c] = s_dist[c];
    } else {
      if (lane == 0) lv[0] = leaf_mean;
    }
    return;
  }
  int32_t nl = 0;
  for (int w = lane; w < nw; w += 32) nl += __popc(s_best[w]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nl += __shfl_xor_sync(FULL, nl, o);
  int32_t slot = 0;
  if (lane == 0) {
    slot = atomicAdd(&p.cnt->next_f, 2);
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
      const int qc = size_class(p, side ? (n - nl) : nl);  // (these classes compute their own class histogram)
      p.q_nxt[qc][atomicAdd(&p.cnt->q_count[qc], 1)] = s2;
      if (qc < p.sub_ncls) atomicAdd(&p.cnt->sub_rows, side ? (n - nl) : nl);
    }
    atomicAdd(&p.cnt->st[ST_PROWS], (unsigned long long)n);
  }
  slot = __shfl_sync(FULL, slot, 0);
  if (!p.replay) {
    uint32_t *ml = p.nxt.mask + (int64_t)slot * W, *mr = ml + W;
    for (int w = lane; w < W; w += 32) {
      const uint32_t v = s_const[w];
      ml[w] = v;
      mr[w] = v;
    }
  }
  // the winner's side bitmask is the stable partition (pkg:1024-1039)
  {
    int32_t lpos = b, rpos = b + nl;
    const uint32_t below = (1u << lane) - 1u;
    for (int w = 0; w < nw; w++) {
      const int j = w * 32 + lane;
      const bool has = j < n;
      const int cnt = min(32, n - w * 32);
      const uint32_t valid = (cnt >= 32) ? FULL : ((1u << cnt) - 1u);
      const uint32_t bm = s_best[w];
      const uint32_t lm = bm & valid, rm = ~bm & valid;
      if (has) {
        const bool left = (lm >> lane) & 1u;
        const int32_t dst = left ? lpos + __popc(lm & below) : rpos + __popc(rm & below);
        p.idx_dst[base + dst] = idx[j];
        if (TASK == TASK_REG) {
          p.yr_dst[base + dst] = s_y[j];
        } else {
          p.yc_dst[base + dst] = p.yc_src[base + b + j];
          if (TASK == TASK_CLSW) p.w_dst[base + dst] = s_y[j];
        }
      }
      lpos += __popc(lm);
      rpos += __popc(rm);
    }
  }
}
#undef LANE_IN

#include "subtree.cuh"

// ---- pool (creation order) -> per-tree pre-order ----------------------------------------------
__global__ void k_subtree_sizes(Pool o, int32_t lo, int32_t hi, int32_t *size, int32_t *nleaf) {
  int v = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= hi) return;
  if (o.feat[v] == SUB_MARK) {  // a finished resident subtree: (nodes, leaves) ride in the cut field
    const unsigned long long packed = (unsigned long long)__double_as_longlong(o.cut[v]);
    size[v] = (int32_t)(packed & 0xffffffffu);
    nleaf[v] = (int32_t)(packed >> 32);
  } else if (o.feat[v] < 0) {
    size[v] = 1;
    nleaf[v] = 1;
  } else {
    int c = o.child[v];
    size[v] = 1 + size[c] + size[c + 1];
    nleaf[v] = nleaf[c] + nleaf[c + 1];
  }
}

// one thread: exclusive scan over the batch's roots -> per-tree node / leaf offsets (forest-wide)
__global__ void k_root_offsets(int32_t B, const int32_t *size, const int32_t *nleaf, int64_t node_base,
                               int64_t leaf_base, int64_t *tree_off, int64_t *leaf_off, int32_t *pos, int32_t *lpos) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int64_t a = node_base, l = leaf_base;
    for (int t = 0; t < B; t++) {
      tree_off[t] = a;
      leaf_off[t] = l;
      a += size[t];
      l += nleaf[t];
      pos[t] = 0;
      lpos[t] = 0;
    }
    tree_off[B] = a;
    leaf_off[B] = l;
  }
}

__global__ void k_assign_pos(Pool o, int32_t lo, int32_t hi, const int32_t *size, const int32_t *nleaf, int32_t *pos,
                             int32_t *lpos) {
  int v = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= hi) return;
  if (o.feat[v] >= 0) {
    int c = o.child[v];
    pos[c] = pos[v] + 1;
    lpos[c] = lpos[v];
    pos[c + 1] = pos[v] + 1 + size[c];
    lpos[c + 1] = lpos[v] + nleaf[c];
  }
}

__global__ void k_scatter(Pool o, int32_t n_nodes, int lw, const int32_t *pos, const int32_t *lpos,
                          const int64_t *tree_off, const int64_t *leaf_off, int64_t node_base, int64_t leaf_base,
                          PNode *nodes, double *leaves) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_nodes) return;
  if (o.feat[v] == SUB_MARK) return;  // k_scatter_sub places the whole block
  const int t = o.tree[v];
  PNode pn;
  const int64_t g = tree_off[t] - node_base + pos[v];
  if (o.feat[v] >= 0) {
    pn.cut = o.cut[v];
    pn.feat = o.feat[v];
    pn.right_or_leaf = pos[o.child[v] + 1];
  } else {
    const int64_t gl = leaf_off[t] + lpos[v];
    pn.cut = NAN;
    pn.feat = -1;
    pn.right_or_leaf = (int32_t)gl;
    const double *src = o.leaf_vals + (int64_t)o.child[v] * lw;
    double *dst = leaves + (gl - leaf_base) * lw;
    for (int c = 0; c < lw; c++) dst[c] = src[c];
  }
  nodes[g] = pn;
}

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

// The frontier (and its queues) of level L lives in slot L % FR_RING.  The resident subtree kernels of a level run
// asynchronously to the level loop and read their slot when their CTAs start, so a slot is only rewritten
// FR_RING - 1 levels later, after those kernels have finished (event wait in launch_level).
constexpr int FR_RING = 4;
static_assert(FR_RING == et_ctx::N_SUB_RING, "one subtree stream per ring slot");

struct FrontierBufs {
  DevBuf<int32_t> tree, begin, end, node, depth, hist;
  DevBuf<int64_t> trace;
  DevBuf<uint64_t> key;
  DevBuf<uint32_t> mask;
  bool fits(size_t F, int C, int W, bool need_hist, bool need_mask) const {
    return tree.cap >= F && begin.cap >= F && end.cap >= F && node.cap >= F && depth.cap >= F && trace.cap >= F &&
           key.cap >= F && (!need_hist || hist.cap >= F * (size_t)C) && (!need_mask || mask.cap >= F * (size_t)W);
  }
  void ensure(size_t F, int C, int W, bool need_hist, bool need_mask) {
    tree.ensure(F, 1.5);
    begin.ensure(F, 1.5);
    end.ensure(F, 1.5);
    node.ensure(F, 1.5);
    depth.ensure(F, 1.5);
    trace.ensure(F, 1.5);
    key.ensure(F, 1.5);
    if (need_hist) hist.ensure(F * (size_t)C, 1.5);
    if (need_mask) mask.ensure(F * (size_t)W, 1.5);
  }
  Frontier view() { return Frontier{tree.p, begin.p, end.p, node.p, depth.p, trace.p, key.p, hist.p, mask.p}; }
};

struct PoolBufs {
  DevBuf<int32_t> tree, feat, child;
  DevBuf<double> cut, leaf_vals;
  bool fits(size_t n, size_t nleaf, int lw) const {
    return tree.cap >= n && feat.cap >= n && child.cap >= n && cut.cap >= n && leaf_vals.cap >= nleaf * (size_t)lw;
  }
  void grow(size_t n, size_t used, size_t nleaf, size_t leaf_used, int lw, cudaStream_t st) {
    tree.grow_keep(n, used, st);
    feat.grow_keep(n, used, st);
    child.grow_keep(n, used, st);
    cut.grow_keep(n, used, st);
    leaf_vals.grow_keep(nleaf * (size_t)lw, leaf_used * (size_t)lw, st);
  }
  Pool view() { return Pool{tree.p, feat.p, child.p, cut.p, leaf_vals.p}; }
};

}  // namespace

// Device buffers that survive across builds on one context (no cudaMalloc in the steady state).
struct Workspace {
  DevBuf<int32_t> idx[2], yc[2], q[FR_RING][NQ], size, nleaf, pos, lpos;
  DevBuf<double> yr[2], ws[2];
  FrontierBufs fr[FR_RING];
  PoolBufs pool;
  DevBuf<uint32_t> scratch;
  DevBuf<PNode> sub_nodes;  // block pool of the resident subtree builder
  DevBuf<Counters> cnt;
  DevBuf<int64_t> tree_off, leaf_off;
};

void et_workspace_free(Workspace *ws) { delete ws; }

namespace {

struct PhaseTimer {
  bool on = getenv("ETGPU_TIMING") != nullptr;
  bool per_level = on && atoi(getenv("ETGPU_TIMING")) >= 2;
  double last[8] = {0};
  void level_report(int level, const int32_t *qn) {
    if (!per_level) return;
    fprintf(stderr, "[etgpu level %3d] nodes sub=%d,%d,%d,%d,%d n32=%d n64=%d n128=%d n256=%d n512=%d mid=%d cta=%d | ms sub=%.3f warp=%.3f mid=%.3f cta=%.3f\n",
            level, qn[0], qn[1], qn[2], qn[3], qn[4], qn[5], qn[6], qn[7], qn[8], qn[9], qn[10], qn[11], acc[7] - last[7],
            acc[2] - last[2], acc[6] - last[6], acc[3] - last[3]);
    for (int i = 0; i < 8; i++) last[i] = acc[i];
  }
  cudaStream_t st;
  double acc[8] = {0};
  std::chrono::steady_clock::time_point t0;
  void start() {
    if (!on) return;
    cudaStreamSynchronize(st);
    t0 = std::chrono::steady_clock::now();
  }
  void stop(int k) {
    if (!on) return;
    cudaStreamSynchronize(st);
    acc[k] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
  void report() {
    if (!on) return;
    static const char *names[] = {"alloc", "init", "node_warp", "node_cta", "sync", "preorder", "node_mid", "node_sub"};
    fprintf(stderr, "[etgpu timing ms]");
    for (int i = 0; i < 8; i++) fprintf(stderr, " %s=%.1f", names[i], acc[i]);
    fprintf(stderr, "\n");
  }
};

// CUDA-event spans of the node kernels (summed after the per-level sync)
struct EventTimer {
  std::vector<cudaEvent_t> pool;
  std::vector<std::pair<int, int>> spans[2];  // 0 = warp-owned nodes, 1 = CTA-owned nodes
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
  void drain(double *acc) {  // call after a stream sync
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

#include "best.cuh"

struct LevelCfg {
  bool coded;      // nodes of up to 512 samples: k_lane on byte codes
  bool coded_big;  // larger nodes: byte-coded CTA teams (unweighted classification, <= 32 classes)
  size_t smem_warp, smem_mid, smem_cta;  // k_node teams (per team)
  size_t smem_lane[5];                   // k_lane per warp, classes 0..4
  int sub_ncls = 0;                      // classes 0 .. sub_ncls - 1: resident subtrees (k_sub), team width 1 << class
  int sub_rw = 0, sub_rowbytes = 0;
  int sub_from = 0;                      // first level whose nodes may be handed to k_sub
  size_t smem_sub[SUB_NCLS] = {0};       // k_sub per CTA
};

template <int TASK, typename VT>
void launch_sub(et_ctx *ctx, const P &p, int32_t count, int q, const LevelCfg &lc, cudaStream_t st) {
  const unsigned grid = (unsigned)count;  // one CTA (team) per subtree
  switch (q) {
    case 0: k_sub<TASK, VT, 1><<<grid, 32, lc.smem_sub[0], st>>>(p, count, q); break;
    case 1: k_sub<TASK, VT, 2><<<grid, 64, lc.smem_sub[1], st>>>(p, count, q); break;
    case 2: k_sub<TASK, VT, 4><<<grid, 128, lc.smem_sub[2], st>>>(p, count, q); break;
    case 3: k_sub<TASK, VT, 8><<<grid, 256, lc.smem_sub[3], st>>>(p, count, q); break;
    default: k_sub<TASK, VT, 16><<<grid, 512, lc.smem_sub[4], st>>>(p, count, q); break;
  }
  ctx->launches++;
}

template <int TASK, typename VT>
void set_sub_attr(const LevelCfg &lc) {
  CUDA_CHECK(cudaFuncSetAttribute(k_sub<TASK, VT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lc.smem_sub[0]));
  CUDA_CHECK(cudaFuncSetAttribute(k_sub<TASK, VT, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lc.smem_sub[1]));
  CUDA_CHECK(cudaFuncSetAttribute(k_sub<TASK, VT, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lc.smem_sub[2]));
  CUDA_CHECK(cudaFuncSetAttribute(k_sub<TASK, VT, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lc.smem_sub[3]));
  CUDA_CHECK(cudaFuncSetAttribute(k_sub<TASK, VT, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lc.smem_sub[4]));
}

template <int TASK, typename VT>
void launch_lane(et_ctx *ctx, const P &p, int32_t count, int qi, int NW, size_t smem_per_warp, cudaStream_t st) {
  const unsigned grid = (unsigned)ceil_div(count, LANE_WARPS);
  if (NW <= 2)
    k_lane<TASK, VT, true><<<grid, 32 * LANE_WARPS, smem_per_warp * LANE_WARPS, st>>>(p, count, qi, NW);
  else
    k_lane<TASK, VT, false><<<grid, 32 * LANE_WARPS, smem_per_warp * LANE_WARPS, st>>>(p, count, qi, NW);
  ctx->launches++;
}

// the byte-coded CTA teams exist for unweighted classification only
template <int TASK, int TEAM>
void launch_coded_team(const P &p, int32_t count, int qi, size_t smem, cudaStream_t st) {
  if constexpr (TASK == TASK_CLS) k_node<TASK_CLS, TEAM, true><<<(unsigned)count, TEAM, smem, st>>>(p, count, qi);
}

// One level = one launch per non-empty size class.  The classes are independent (disjoint nodes), so
// each runs on its own stream: the few long-running CTAs of the large nodes overlap with the many
// small teams instead of serialising behind them.  (ETGPU_TIMING serialises them to time each.)
struct SubEvents {  // completion events of the asynchronous subtree kernels, per frontier slot and team class
  cudaEvent_t ev[FR_RING][SUB_NCLS] = {};
  bool valid[FR_RING][SUB_NCLS] = {};
  bool inflight = false;
};

template <int TASK>
void launch_level(et_ctx *ctx, const P &p, const int32_t *qn, const LevelCfg &lc, PhaseTimer &pt, EventTimer &et,
                  SubEvents &se, int slot_cur, int slot_nxt) {
  cudaStream_t main_st = ctx->stream;
  const bool fork = !pt.on;
  const bool async_sub = !pt.on && lc.sub_ncls > 0;
  // this level rewrites frontier slot `slot_nxt`: the subtree kernels that read it last must have finished
  for (int q = 0; q < SUB_NCLS; q++) {
    if (se.valid[slot_nxt][q]) {
      cudaStreamWaitEvent(main_st, se.ev[slot_nxt][q], 0);
      se.valid[slot_nxt][q] = false;
    }
  }
  int used = 0;
  for (int q = 0; q < NQ; q++) used += (qn[q] > 0 && !(async_sub && q < lc.sub_ncls));
  const int e0 = et.rec(main_st);
  if (fork && used > 1) cudaEventRecord(ctx->ev_fork, main_st);
  bool joined[NQ] = {false};
  int first = 1;
  auto stream_for = [&](int side) -> cudaStream_t {
    if (!fork || used <= 1 || first) {  // the first (largest) class stays on the main stream
      first = 0;
      return main_st;
    }
    cudaStreamWaitEvent(ctx->side[side], ctx->ev_fork, 0);
    joined[side] = true;
    return ctx->side[side];
  };
  // largest teams first: their CTAs run longest
  if (qn[Q_CTA] > 0) {
    pt.start();
    cudaStream_t st = stream_for(Q_CTA);
    if (lc.coded_big)
      launch_coded_team<TASK, CBIG_TEAM>(p, qn[Q_CTA], Q_CTA, lc.smem_cta, st);
    else
      k_node<TASK, CTA_TEAM, false><<<(unsigned)qn[Q_CTA], CTA_TEAM, lc.smem_cta, st>>>(p, qn[Q_CTA], Q_CTA);
    ctx->launches++;
    pt.stop(3);
  }
  if (qn[Q_MID] > 0) {
    pt.start();
    cudaStream_t st = stream_for(Q_MID);
    if (lc.coded_big)
      launch_coded_team<TASK, MID_TEAM>(p, qn[Q_MID], Q_MID, lc.smem_mid, st);
    else
      k_node<TASK, MID_TEAM, false><<<(unsigned)qn[Q_MID], MID_TEAM, lc.smem_mid, st>>>(p, qn[Q_MID], Q_MID);
    ctx->launches++;
    pt.stop(6);
  }
  for (int q = Q_WARP; q >= 0; q--) {
    if (qn[q] <= 0) continue;
    if (async_sub && q < lc.sub_ncls) {
      // resident subtrees produce nothing the next level needs: their stream is not joined at the level end
      // (everything they read was finished before the host launched this level)
      cudaStream_t sst = ctx->sub_stream[q][slot_cur];
      if (lc.coded)
        launch_sub<TASK, uint8_t>(ctx, p, qn[q], q, lc, sst);
      else
        launch_sub<TASK, double>(ctx, p, qn[q], q, lc, sst);
      cudaEventRecord(se.ev[slot_cur][q], sst);
      se.valid[slot_cur][q] = true;
      se.inflight = true;
      continue;
    }
    pt.start();
    cudaStream_t st = stream_for(q);
    if (q < lc.sub_ncls) {
      if (lc.coded)
        launch_sub<TASK, uint8_t>(ctx, p, qn[q], q, lc, st);
      else
        launch_sub<TASK, double>(ctx, p, qn[q], q, lc, st);
    } else if (lc.coded) {
      const int cq = q - Q_LANE0;
      launch_lane<TASK, uint8_t>(ctx, p, qn[q], q, 1 << cq, lc.smem_lane[cq], st);
    } else if (q == Q_LANE0) {
      launch_lane<TASK, double>(ctx, p, qn[q], q, 1, lc.smem_lane[0], st);
    } else {  // FP64 tables: only class Q_WARP is populated besides class Q_LANE0 (and the subtree classes)
      k_node<TASK, 32, false>
          <<<(unsigned)ceil_div(qn[q], WARPS_PER_CTA), 32 * WARPS_PER_CTA, lc.smem_warp * WARPS_PER_CTA, st>>>(p, qn[q], q);
      ctx->launches++;
    }
    pt.stop(q < lc.sub_ncls ? 7 : 2);
  }
  for (int i = 0; i < NQ; i++) {
    if (joined[i]) {
      cudaEventRecord(ctx->ev_join[i], ctx->side[i]);
      cudaStreamWaitEvent(main_st, ctx->ev_join[i], 0);
    }
  }
  const int e1 = et.rec(main_st);
  et.spans[0].push_back({e0, e1});
  if (qn[Q_MID] + qn[Q_CTA] > 0) et.spans[1].push_back({e0, e1});
}

template <int TASK>
void set_smem_attr(const LevelCfg &lc) {
  if (lc.coded_big) {
    if constexpr (TASK == TASK_CLS) {
      CUDA_CHECK(cudaFuncSetAttribute(k_node<TASK_CLS, MID_TEAM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)lc.smem_mid));
      CUDA_CHECK(cudaFuncSetAttribute(k_node<TASK_CLS, CBIG_TEAM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)lc.smem_cta));
    }
  } else {
    CUDA_CHECK(cudaFuncSetAttribute(k_node<TASK, MID_TEAM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)lc.smem_mid));
    CUDA_CHECK(cudaFuncSetAttribute(k_node<TASK, CTA_TEAM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)lc.smem_cta));
  }
  if (lc.coded) {
    CUDA_CHECK(cudaFuncSetAttribute(k_lane<TASK, uint8_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(std::max(lc.smem_lane[0], lc.smem_lane[1]) * LANE_WARPS)));
    CUDA_CHECK(cudaFuncSetAttribute(k_lane<TASK, uint8_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(std::max(lc.smem_lane[2], std::max(lc.smem_lane[3], lc.smem_lane[4])) * LANE_WARPS)));
  } else {
    CUDA_CHECK(cudaFuncSetAttribute(k_lane<TASK, double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(lc.smem_lane[0] * LANE_WARPS)));
    CUDA_CHECK(cudaFuncSetAttribute(k_node<TASK, 32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(lc.smem_warp * WARPS_PER_CTA)));
  }
  if (lc.sub_ncls > 0) {
    if (lc.coded)
      set_sub_attr<TASK, uint8_t>(lc);
    else
      set_sub_attr<TASK, double>(lc);
  }
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
  if (!ctx->ws) ctx->ws = new Workspace();
  Workspace &ws = *ctx->ws;
  et_stats S;
  memset(&S, 0, sizeof(S));
  const int64_t launches0 = ctx->launches;
  PhaseTimer pt;
  pt.st = st;
  EventTimer evt;
  double tacc[2] = {0, 0};

  // candidates per batch: bounded by 32 lanes and by the team's shared memory
  int NB = 32;
  const size_t smem_budget_warp = 40 * 1024, smem_budget_cta = 110 * 1024;
  while (NB > 1 && ((size_t)make_lay(task, 32, C, NB, W, replay).bytes > smem_budget_warp ||
                    (size_t)make_lay(task, CTA_TEAM, C, NB, W, replay).bytes > smem_budget_cta))
    NB--;
  const Lay lay_w = make_lay(task, 32, C, NB, W, replay), lay_c = make_lay(task, CTA_TEAM, C, NB, W, replay);
  const Lay lay_m = make_lay(task, MID_TEAM, C, NB, W, replay);
  if ((size_t)lay_w.bytes * WARPS_PER_CTA > 200 * 1024 || (size_t)lay_c.bytes > 200 * 1024)
    ET_FAIL(ET_EUNSUPPORTED, "numClasses=%d / %d features need more shared memory per node than one SM has", C, d);
  // byte-coded copy of the table (encode.cu): nodes of up to 512 samples are then searched by k_lane
  if (D->coded == 0) et_data_encode(ctx, D);
  LevelCfg lc;
  lc.coded = (D->coded == 1);
  lc.smem_warp = (size_t)lay_w.bytes;
  lc.smem_mid = (size_t)lay_m.bytes;
  lc.smem_cta = (size_t)lay_c.bytes;
  for (int q = 0; q < 5; q++) lc.smem_lane[q] = (size_t)lane_smem_bytes(task, C, W, replay, 1 << q, 1);
  if (lc.coded && lc.smem_lane[4] * LANE_WARPS > 200 * 1024) lc.coded = false;  // (hundreds of classes)
  lc.coded_big = lc.coded && task == TASK_CLS && C <= 32 && (uint64_t)D->ldc * (uint64_t)d < ((uint64_t)1 << 32);
  if (lc.coded_big) {
    lc.smem_mid = (size_t)make_lay(task, MID_TEAM, C, NB, W, replay, true).bytes;
    lc.smem_cta = (size_t)make_lay(task, CBIG_TEAM, C, NB, W, replay, true).bytes;
  }
  if (!lc.coded) lc.smem_lane[0] = (size_t)lane_smem_bytes(task, C, W, replay, 1, 8);
  // resident subtrees: how many rows of the row-major copy fit one SM next to the teams' scratch
  if (!a.best_split) {
    const bool have_rm = lc.coded ? (D->r8 != nullptr) : (D->xr != nullptr);
    const int64_t rowbytes = lc.coded ? D->rs8 : D->rsd * 8;
    int smem_max = 0;
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device);
    // Off unless ETGPU_SUB_NCLS asks for it: measured on B200 the subtree kernels cut the
    // level loop from 145 to 110 ms on the MNIST-shaped workload but need more SM time for the small nodes than
    // the gathering kernels they replace, so the build as a whole gets slower.
    int ncls = 0, rw_cap = 16;
    if (const char *env = getenv("ETGPU_SUB_NCLS")) ncls = std::max(0, std::min(SUB_NCLS, atoi(env)));
    if (const char *env = getenv("ETGPU_SUB_RW")) rw_cap = std::max(1, std::min(16, atoi(env)));
    if (const char *env = getenv("ETGPU_SUB_FROM_LEVEL")) lc.sub_from = std::max(0, atoi(env));
    if (have_rm && ncls > 0 && (task == TASK_REG || C <= 256) && rowbytes < (1 << 20)) {
      int rw = rw_cap;
      // a team of 1 << q warps is one CTA; SUB_WARPS >> q of them share an SM (1 KB per CTA is reserved)
      const size_t sm_total = (size_t)smem_max + 1024;
      auto fits = [&](int r) {
        for (int q = 0; q < SUB_NCLS; q++)
          if ((sub_smem_bytes(task, C, W, replay, 1 << q, r, (int)rowbytes, lc.coded) + 1024) * (size_t)(SUB_WARPS >> q) >
              sm_total)
            return false;
        return true;
      };
      while (rw >= 4 && !fits(rw)) rw--;
      if (rw >= 4) {
        lc.sub_ncls = ncls;
        lc.sub_rw = rw;
        lc.sub_rowbytes = (int)rowbytes;
        for (int q = 0; q < SUB_NCLS; q++)
          lc.smem_sub[q] = sub_smem_bytes(task, C, W, replay, 1 << q, rw, (int)rowbytes, lc.coded);
      }
    }
  }
  if (lc.smem_lane[0] * LANE_WARPS > 200 * 1024)
    ET_FAIL(ET_EUNSUPPORTED, "numClasses=%d / %d features need more shared memory per node than one SM has", C, d);
  if (task == TASK_CLS)
    set_smem_attr<TASK_CLS>(lc);
  else if (task == TASK_CLSW)
    set_smem_attr<TASK_CLSW>(lc);
  else
    set_smem_attr<TASK_REG>(lc);

  cudaEvent_t ev0, ev1;
  CUDA_CHECK(cudaEventCreate(&ev0));
  CUDA_CHECK(cudaEventCreate(&ev1));
  SubEvents sub_ev;
  BestBufs best_bufs;  // (bestSplit builds only)
  for (int r = 0; r < FR_RING; r++)
    for (int q = 0; q < SUB_NCLS; q++) CUDA_CHECK(cudaEventCreateWithFlags(&sub_ev.ev[r][q], cudaEventDisableTiming));
  CUDA_CHECK(cudaEventRecord(ev0, st));

  // the reference's seeding (pkg:629,654-655) names one stream per tree; the free-running GPU RNG
  // is counter based and keyed by (seed, global tree id)
  std::vector<uint64_t> tree_keys((size_t)std::max(a.m, 1));
  for (int t = 0; t < a.m; t++) {
    uint64_t gid = a.tree_ids ? (uint64_t)(uint32_t)a.tree_ids[t] : (uint64_t)t;
    tree_keys[(size_t)t] = et_tree_key((uint64_t)a.seed, gid);
  }

  // batch size: bound the per-sample state (idx + targets, ping-pong) to ~8 GB
  size_t per_sample = 2 * (4 + (task == TASK_REG ? 8 : 4) + (task == TASK_CLSW ? 8 : 0));
  int64_t max_samples = (int64_t)(((size_t)8 << 30) / per_sample);
  int32_t B = (int32_t)std::max<int64_t>(1, std::min<int64_t>(a.m, max_samples / std::max<int64_t>(n, 1)));
  if (const char *env = getenv("ETGPU_BATCH_TREES")) B = std::max(1, std::min(a.m, atoi(env)));

  ws.cnt.ensure(1);
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
    if (nn > 0x7fffffff) ET_FAIL(ET_EREPLAY, "replay trace too large");
    std::vector<int32_t> l((size_t)nn), r((size_t)nn);
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
  struct Seg {
    PNode *nodes;
    double *leaves;
    int64_t n_nodes, n_leaves;
  };
  std::vector<Seg> segs;
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
    for (int r = 0; r < FR_RING; r++)
      for (int q = 0; q < SUB_NCLS; q++)
        if (sub_ev.ev[r][q]) cudaEventDestroy(sub_ev.ev[r][q]);
  };

  out->m = a.m;
  out->tree_off.assign((size_t)a.m + 1, 0);
  int64_t node_base = 0, leaf_base = 0;  // forest-wide offsets of the current batch

  try {
    for (int32_t t0 = 0; t0 < a.m; t0 += B) {
      const int32_t Bt = std::min(B, a.m - t0);
      const size_t ns = (size_t)Bt * (size_t)n;
      pt.start();
      for (int q = 0; q < 2; q++) {
        ws.idx[q].ensure(ns, 1.0);
        if (task == TASK_REG)
          ws.yr[q].ensure(ns, 1.0);
        else
          ws.yc[q].ensure(ns, 1.0);
        if (task == TASK_CLSW) ws.ws[q].ensure(ns, 1.0);
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
      p.NB = NB;
      p.cnt = ws.cnt.p;
      p.C8 = lc.coded ? D->c8 : nullptr;
      p.ldc = D->ldc;
      p.dict = D->dict;
      p.coff = D->coff;
      p.c8_small = ((uint64_t)D->ldc * (uint64_t)d < ((uint64_t)1 << 32) &&
                    (uint64_t)std::max<int64_t>(D->rs8, 1) * (uint64_t)n < ((uint64_t)1 << 32))
                       ? 1
                       : 0;
      for (int q = 0; q < Q_LANE0; q++) p.cls_max[q] = 0;  // 0: empty class (set per level below)
      p.cls_max[Q_LANE0] = NT_MAX;
      for (int q = 1; q < 4; q++) p.cls_max[Q_LANE0 + q] = lc.coded ? (NT_MAX << q) : NT_MAX;
      p.cls_max[Q_WARP] = NW_MAX;
      p.cls_max[Q_MID] = NM_MAX;
      if (a.best_split)  // bestSplit: one kernel family (best.cuh), every node in the last queue
        for (int q = 0; q < NQ - 1; q++) p.cls_max[q] = 0;
      p.R8 = D->r8;
      p.r8_stride = (int32_t)D->rs8;
      p.nc_max = 64;
      if (const char *env = getenv("ETGPU_NC_MAX")) p.nc_max = std::max(0, atoi(env));
      p.XR = D->xr;
      p.sub_ncls = 0;
      if (lc.sub_ncls > 0 && lc.sub_from <= 0) {  // the roots themselves may be small enough
        p.sub_ncls = lc.sub_ncls;
        for (int q = 0; q < lc.sub_ncls; q++) p.cls_max[q] = lc.sub_rw << q;
      }
      p.sub_rw = lc.sub_rw;
      p.sub_rowbytes = lc.sub_rowbytes;
      p.tr = Trace{d_tr_cand_begin, d_tr_cand_count, d_tr_left, d_tr_right, d_tr_cand_feature, d_tr_cand_u,
                   d_tr_cand_flag};
      int srcb = 0, cl = 0;  // cl: frontier slot of the current level (ring of FR_RING)
      const int level0 = (int)S.levels;  // levels counted before this batch
      int32_t F = Bt;
      ws.fr[0].ensure((size_t)F, C, W, task == TASK_CLS, !replay);
      for (int q = 0; q < NQ; q++) ws.q[0][q].ensure((size_t)F, 1.5);
      pt.stop(0);
      pt.start();
      {
        unsigned grid = (unsigned)std::min<int64_t>(ceil_div((int64_t)ns, 256), (int64_t)ctx->sm_count * 16);
        k_init_samples<<<std::max(grid, 1u), 256, 0, st>>>(n, Bt, ws.idx[0].p, D->y_cls,
                                                          task == TASK_REG ? nullptr : ws.yc[0].p, D->y_reg,
                                                          task == TASK_REG ? ws.yr[0].p : nullptr, D->w,
                                                          task == TASK_CLSW ? ws.ws[0].p : nullptr);
        ctx->launches++;
      }
      p.cur = ws.fr[0].view();
      for (int q = 0; q < NQ; q++) p.q_cur[q] = ws.q[0][q].p;
      uint64_t *d_keys = upload_tmp(tree_keys.data() + t0, (size_t)Bt, st);
      int64_t *d_troots = replay ? upload_tmp(trace_roots.data() + t0, (size_t)Bt, st) : nullptr;
      CUDA_CHECK(cudaMemsetAsync(ws.cnt.p, 0, sizeof(Counters), st));
      k_init_roots<<<(unsigned)ceil_div(F, 128), 128, 0, st>>>(p, Bt, d_keys, d_troots, d_root_hist);
      ctx->launches++;
      CUDA_CHECK(cudaStreamSynchronize(st));
      cudaFree(d_keys);
      if (d_troots) cudaFree(d_troots);
      pt.stop(1);

      int64_t n_nodes = Bt, n_leaves = 0;
      std::vector<int32_t> level_start{0};
      int32_t qn[NQ] = {0};
      qn[size_class(p, n)] = Bt;
      int64_t sub_rows_level = (size_class(p, n) < p.sub_ncls) ? (int64_t)Bt * n : 0;  // rows entering k_sub this level
      int64_t sub_nodes_used = 0;
      int64_t leaf_bound = 0, subnode_bound = 0;  // upper bounds of the leaves / block nodes allocated so far
      Counters hc;
      memset(&hc, 0, sizeof(hc));
      // the asynchronous subtree kernels hold pointers into the pools and the frontier ring: a buffer only moves
      // after they have drained (never in the steady state, where the workspace is already large enough)
      auto quiesce = [&]() {
        if (!sub_ev.inflight) return;
        for (int q = 0; q < SUB_NCLS; q++)
          for (int r = 0; r < FR_RING; r++) CUDA_CHECK(cudaStreamSynchronize(ctx->sub_stream[q][r]));
        sub_ev.inflight = false;
        for (int r = 0; r < FR_RING; r++)
          for (int q = 0; q < SUB_NCLS; q++) sub_ev.valid[r][q] = false;
        Counters hq;
        CUDA_CHECK(cudaMemcpyAsync(&hq, ws.cnt.p, sizeof(Counters), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        n_leaves = hq.n_leaves;
        sub_nodes_used = (int64_t)hq.sub_nodes;
      };
      while (F > 0) {
        S.levels++;
        pt.start();
        const int cn = (cl + 1) % FR_RING;
        leaf_bound += (int64_t)F + sub_rows_level;
        subnode_bound += 2 * sub_rows_level;
        {
          bool fits = ws.fr[cn].fits((size_t)F * 2, C, W, task == TASK_CLS, !replay) &&
                      ws.pool.fits((size_t)(n_nodes + 2 * (int64_t)F), (size_t)leaf_bound, lw) &&
                      ws.sub_nodes.cap >= (size_t)subnode_bound;
          for (int q = 0; q < NQ; q++) fits = fits && ws.q[cn][q].cap >= (size_t)F * 2;
          if (!fits) quiesce();
        }
        ws.fr[cn].ensure((size_t)F * 2, C, W, task == TASK_CLS, !replay);
        for (int q = 0; q < NQ; q++) ws.q[cn][q].ensure((size_t)F * 2, 1.5);
        ws.pool.grow((size_t)(n_nodes + 2 * (int64_t)F), (size_t)n_nodes, (size_t)leaf_bound, (size_t)n_leaves, lw, st);
        if (subnode_bound > 0) ws.sub_nodes.grow_keep((size_t)subnode_bound, (size_t)sub_nodes_used, st);
        if (task != TASK_CLS && qn[Q_MID] + qn[Q_CTA] > 0)
          ws.scratch.ensure((size_t)NB * 2 * ((size_t)Bt * (size_t)n / 32 + (size_t)(qn[Q_MID] + qn[Q_CTA]) + 1) + 64, 1.0);
        pt.stop(0);
        p.idx_src = ws.idx[srcb].p;
        p.idx_dst = ws.idx[srcb ^ 1].p;
        p.yc_src = ws.yc[srcb].p;
        p.yc_dst = ws.yc[srcb ^ 1].p;
        p.yr_src = ws.yr[srcb].p;
        p.yr_dst = ws.yr[srcb ^ 1].p;
        p.w_src = ws.ws[srcb].p;
        p.w_dst = ws.ws[srcb ^ 1].p;
        p.cur = ws.fr[cl].view();
        p.nxt = ws.fr[cn].view();
        for (int q = 0; q < NQ; q++) {
          p.q_cur[q] = ws.q[cl][q].p;
          p.q_nxt[q] = ws.q[cn][q].p;
        }
        p.o = ws.pool.view();
        p.scratch = ws.scratch.p;
        p.sub_nodes = ws.sub_nodes.p;
        p.node_base_next = (int32_t)n_nodes;
        // the children made at this level go to the subtree classes once the level loop has reached sub_from
        // (the wide top levels stay with the gathering kernels; the long thin tail of the tree is where a
        // resident subtree saves a kernel launch and a host round trip per level)
        if (lc.sub_ncls > 0 && (int)S.levels - 1 - level0 + 1 >= lc.sub_from) {
          p.sub_ncls = lc.sub_ncls;
          for (int q = 0; q < lc.sub_ncls; q++) p.cls_max[q] = lc.sub_rw << q;
        }
        if (a.best_split) {
          const int e0 = evt.rec(st);
          if (task == TASK_CLS)
            launch_level_best<TASK_CLS>(ctx, p, qn[Q_CTA], best_bufs, (int64_t)ns);
          else if (task == TASK_CLSW)
            launch_level_best<TASK_CLSW>(ctx, p, qn[Q_CTA], best_bufs, (int64_t)ns);
          else
            launch_level_best<TASK_REG>(ctx, p, qn[Q_CTA], best_bufs, (int64_t)ns);
          evt.spans[0].push_back({e0, evt.rec(st)});
        } else if (task == TASK_CLS)
          launch_level<TASK_CLS>(ctx, p, qn, lc, pt, evt, sub_ev, cl, cn);
        else if (task == TASK_CLSW)
          launch_level<TASK_CLSW>(ctx, p, qn, lc, pt, evt, sub_ev, cl, cn);
        else
          launch_level<TASK_REG>(ctx, p, qn, lc, pt, evt, sub_ev, cl, cn);
        pt.level_report((int)S.levels - 1, qn);
        pt.start();
        CUDA_CHECK(cudaMemcpyAsync(&hc, ws.cnt.p, sizeof(Counters), cudaMemcpyDeviceToHost, st));
        // the next level starts from clean per-level counters (leaf count and stats keep accumulating)
        CUDA_CHECK(cudaMemsetAsync(ws.cnt.p, 0, offsetof(Counters, n_leaves), st));
        CUDA_CHECK(cudaMemsetAsync(&ws.cnt.p->scratch_words, 0, sizeof(unsigned long long), st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        CUDA_CHECK(cudaGetLastError());
        {
          const double before = tacc[0];
          evt.drain(tacc);
          static const bool level_ms = getenv("ETGPU_LEVEL_MS") != nullptr;
          if (level_ms) {
            fprintf(stderr, "[etgpu level %3d] %.3f ms | nodes", (int)S.levels - 1, tacc[0] - before);
            for (int q = 0; q < NQ; q++) fprintf(stderr, " %d", qn[q]);
            fprintf(stderr, "\n");
          }
        }
        pt.stop(4);
        const int32_t nf = hc.next_f;
        n_leaves = hc.n_leaves;  // (a snapshot while subtree kernels are in flight; exact again after quiesce())
        level_start.push_back((int32_t)n_nodes);
        n_nodes += nf;
        if (n_nodes > 0x7ffffff0) ET_FAIL(ET_EUNSUPPORTED, "batch exceeds 2^31 nodes; lower ETGPU_BATCH_TREES");
        for (int q = 0; q < NQ; q++) qn[q] = hc.q_count[q];
        sub_rows_level = hc.sub_rows;
        sub_nodes_used = (int64_t)hc.sub_nodes;
        if (n_nodes + subnode_bound > 0x7ffffff0) ET_FAIL(ET_EUNSUPPORTED, "batch exceeds 2^31 nodes; lower ETGPU_BATCH_TREES");
        if (nf > 0) srcb ^= 1;
        F = nf;
        cl = cn;
      }
      // the subtree kernels still in flight finish the batch
      quiesce();
      CUDA_CHECK(cudaMemcpyAsync(&hc, ws.cnt.p, sizeof(Counters), cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaStreamSynchronize(st));
      CUDA_CHECK(cudaGetLastError());
      n_leaves = hc.n_leaves;
      sub_nodes_used = (int64_t)hc.sub_nodes;
      S.v_mm += (int64_t)hc.st[ST_VMM];
      S.v_sc += (int64_t)hc.st[ST_VSC];
      S.s_rows += (int64_t)hc.st[ST_SROWS];
      S.p_rows += (int64_t)hc.st[ST_PROWS];
      S.draws += (int64_t)hc.st[ST_DRAWS];
      S.const_hits += (int64_t)hc.st[ST_CONST];
      S.scored += (int64_t)hc.st[ST_SCORED];
      S.replay_mismatches += (int64_t)hc.st[ST_MISMATCH];
      S.parallel_sum_nodes += (int64_t)hc.st[ST_PARNODES];
      S.ambiguous_splits += (int64_t)hc.st[ST_AMBIG];
      // ---- creation order -> per-tree pre-order, on the device
      pt.start();
      ws.size.ensure((size_t)n_nodes);
      ws.nleaf.ensure((size_t)n_nodes);
      ws.pos.ensure((size_t)n_nodes);
      ws.lpos.ensure((size_t)n_nodes);
      ws.tree_off.ensure((size_t)Bt + 1);
      ws.leaf_off.ensure((size_t)Bt + 1);
      Pool po = ws.pool.view();
      // level l holds node ids [level_start[l], level_start[l+1]) (the last entry closes the list)
      const int nlev = (int)level_start.size() - 1;
      for (int l = nlev - 1; l >= 0; l--) {
        int32_t lo = level_start[(size_t)l], hi = level_start[(size_t)l + 1];
        if (hi <= lo) continue;
        k_subtree_sizes<<<(unsigned)ceil_div(hi - lo, 256), 256, 0, st>>>(po, lo, hi, ws.size.p, ws.nleaf.p);
        ctx->launches++;
      }
      k_root_offsets<<<1, 32, 0, st>>>(Bt, ws.size.p, ws.nleaf.p, node_base, leaf_base, ws.tree_off.p, ws.leaf_off.p,
                                       ws.pos.p, ws.lpos.p);
      ctx->launches++;
      for (int l = 0; l < nlev; l++) {
        int32_t lo = level_start[(size_t)l], hi = level_start[(size_t)l + 1];
        if (hi <= lo) continue;
        k_assign_pos<<<(unsigned)ceil_div(hi - lo, 256), 256, 0, st>>>(po, lo, hi, ws.size.p, ws.nleaf.p, ws.pos.p,
                                                                      ws.lpos.p);
        ctx->launches++;
      }
      // (marker nodes of resident subtrees stand for whole blocks: the forest's node count comes from the offsets)
      std::vector<int64_t> toff((size_t)Bt + 1);
      CUDA_CHECK(cudaMemcpyAsync(toff.data(), ws.tree_off.p, ((size_t)Bt + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaStreamSynchronize(st));
      const int64_t total_nodes = toff[(size_t)Bt] - node_base;
      S.nodes += total_nodes;
      Seg sg;
      sg.n_nodes = total_nodes;
      sg.n_leaves = n_leaves;
      sg.nodes = nullptr;
      sg.leaves = nullptr;
      sg.nodes = static_cast<PNode *>(et_dev_alloc(ctx, (size_t)total_nodes * sizeof(PNode)));
      sg.leaves = static_cast<double *>(et_dev_alloc(ctx, std::max<size_t>(1, (size_t)n_leaves * lw) * sizeof(double)));
      if (!sg.nodes || !sg.leaves) {
        et_dev_free(ctx, sg.nodes, (size_t)total_nodes * sizeof(PNode));
        et_dev_free(ctx, sg.leaves, std::max<size_t>(1, (size_t)n_leaves * lw) * sizeof(double));
        ET_FAIL(ET_ENOMEM, "cannot allocate the forest (%lld nodes)", (long long)total_nodes);
      }
      segs.push_back(sg);
      k_scatter<<<(unsigned)ceil_div(n_nodes, 256), 256, 0, st>>>(po, (int32_t)n_nodes, lw, ws.pos.p, ws.lpos.p,
                                                                  ws.tree_off.p, ws.leaf_off.p, node_base, leaf_base,
                                                                  sg.nodes, sg.leaves);
      ctx->launches++;
      if (sub_nodes_used > 0) {
        k_scatter_sub<<<(unsigned)ceil_div(n_nodes, 256), 256, 0, st>>>(po, (int32_t)n_nodes, lw, ws.pos.p, ws.lpos.p,
                                                                        ws.tree_off.p, ws.leaf_off.p, node_base, leaf_base,
                                                                        ws.sub_nodes.p, sg.nodes, sg.leaves);
        ctx->launches++;
      }
      CUDA_CHECK(cudaStreamSynchronize(st));
      CUDA_CHECK(cudaGetLastError());
      for (int32_t t = 0; t <= Bt; t++) out->tree_off[(size_t)(t0 + t)] = toff[(size_t)t];
      node_base += total_nodes;
      leaf_base += n_leaves;
      pt.stop(5);
    }
    // ---- the forest stays resident in HBM; batches are concatenated
    out->total_nodes = node_base;
    out->total_leaves = leaf_base;
    auto free_seg = [&](Seg &sg) {
      et_dev_free(ctx, sg.nodes, (size_t)sg.n_nodes * sizeof(PNode));
      et_dev_free(ctx, sg.leaves, std::max<size_t>(1, (size_t)sg.n_leaves * lw) * sizeof(double));
    };
    if (segs.size() == 1) {
      out->d_nodes = segs[0].nodes;
      out->d_leaf = segs[0].leaves;
      out->nodes_bytes = (size_t)segs[0].n_nodes * sizeof(PNode);
      out->leaf_bytes = std::max<size_t>(1, (size_t)segs[0].n_leaves * lw) * sizeof(double);
      segs.clear();
    } else {
      CUDA_CHECK(cudaMalloc((void **)&out->d_nodes, std::max<size_t>(1, (size_t)node_base) * sizeof(PNode)));
      CUDA_CHECK(cudaMalloc((void **)&out->d_leaf, std::max<size_t>(1, (size_t)leaf_base * lw) * sizeof(double)));
      int64_t no = 0, lo = 0;
      for (auto &sg : segs) {
        CUDA_CHECK(cudaMemcpyAsync(out->d_nodes + no, sg.nodes, (size_t)sg.n_nodes * sizeof(PNode),
                                   cudaMemcpyDeviceToDevice, st));
        CUDA_CHECK(cudaMemcpyAsync(out->d_leaf + lo * lw, sg.leaves, (size_t)sg.n_leaves * lw * sizeof(double),
                                   cudaMemcpyDeviceToDevice, st));
        no += sg.n_nodes;
        lo += sg.n_leaves;
      }
      CUDA_CHECK(cudaStreamSynchronize(st));
      for (auto &sg : segs) free_seg(sg);
      segs.clear();
    }
    CUDA_CHECK(cudaMalloc((void **)&out->d_tree_off, ((size_t)a.m + 1) * sizeof(int64_t)));
    CUDA_CHECK(cudaMemcpyAsync(out->d_tree_off, out->tree_off.data(), ((size_t)a.m + 1) * sizeof(int64_t),
                               cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaEventRecord(ev1, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    S.gpu_ms = ms;
    S.gpu_ms_split = tacc[0];      // node kernels (split search + partition fused), all size classes of a level overlapped
    S.gpu_ms_partition = tacc[1];  // ... of which levels that still hold CTA-owned (large) nodes
    S.launches = ctx->launches - launches0;
    pt.report();
  } catch (...) {
    cudaStreamSynchronize(st);
    for (int q = 0; q < SUB_NCLS; q++)
      for (int r = 0; r < FR_RING; r++) cudaStreamSynchronize(ctx->sub_stream[q][r]);
    for (auto &sg : segs) {
      et_dev_free(ctx, sg.nodes, (size_t)sg.n_nodes * sizeof(PNode));
      et_dev_free(ctx, sg.leaves, std::max<size_t>(1, (size_t)sg.n_leaves * lw) * sizeof(double));
    }
    free_tmp();
    throw;
  }
  free_tmp();
  if (stats) *stats = S;
  if (replay && S.replay_mismatches && !stats)  // with stats the caller reads replay_mismatches itself
    ET_FAIL(ET_EREPLAY, "replay: %lld decisions contradict the trace (wrong data for this trace?)",
            (long long)S.replay_mismatches);
}
