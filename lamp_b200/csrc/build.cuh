// build.cuh -- shared declarations of the level-wise extratrees builder (included by every builder TU).
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
//   queues    frontier indices bucketed by node size (see the size classes below)
//   pool      output nodes in creation (level) order + compact leaf-value pool; converted on the
//             device to per-tree pre-order 16-byte nodes (the layout predict traverses)
//
// Translation units: node_inst.cu (k_lane / k_node, compiled once per task), wide.cu (nodes split over
// several CTAs), best.cu (bestSplit = true), build.cu (host orchestration, roots, pre-order conversion).
//
// Exactness: every floating-point expression of the reference is evaluated with individually
// rounded _rn operations in the reference's order.  Unweighted classification reduces integer
// class histograms in parallel (exact) and evaluates the Gini expressions in one thread per
// candidate; weighted classification and regression sum in subset order (sequential chains, one
// thread per candidate) because FP addition is not associative.
#pragma once
#include <algorithm>
#include <chrono>
#include <type_traits>

#include "internal.h"

namespace etb {

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
#ifndef MID_CODED_CTAS
#define MID_CODED_CTAS 4  // resident 128-thread CTAs per SM of the byte-coded team for 513..2048-row nodes
#endif
constexpr int CBIG_TEAM = 256;    // threads of the CTA that owns a larger node of a byte-coded table
constexpr int WARPS_PER_CTA = 4;  // warp teams per CTA in the small-node kernel

// Size classes of the open nodes (upper bounds in P::cls_max; an empty class repeats its predecessor's bound):
//   0..4  n <= 32, 64, 128, 256, 512: one warp per node, one lane per candidate (k_lane; classes 1..4 only on
//         byte-coded tables, each with shared memory sized to its bound); FP64 tables: class 4 = n <= 512, one
//         warp per node with lanes on samples (k_node<32>)
//   5     n <= 2048: one 128-thread CTA per node (k_node)
//   6     n <= P::wide_min: one 256/512-thread CTA per node (k_node)
//   7     larger: the node is cut into chunks of rows, one CTA per chunk (wide.cu)
constexpr int NQ = 8;
constexpr int Q_LANE0 = 0, Q_WARP = 4, Q_MID = 5, Q_CTA = 6, Q_WIDE = 7;

struct Counters {
  int32_t next_f;
  int32_t q_count[NQ];
  int32_t n_leaves;
  int32_t max_feat;              // 1 + the largest split feature of the batch (written by the pre-order conversion)
  unsigned long long wide_rows;  // rows of the nodes queued for the chunked (multi-CTA) path at the next level
  unsigned long long big_rows;   // rows of all nodes of more than NM_MAX rows queued for the next level
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
  const uint8_t *R8;   // row-major byte codes [n][r8_stride] (gathered by k_lane)
  int32_t r8_stride;
  int32_t nc_max;      // k_lane: nodes of up to this many rows draw from their varying-feature set
  int32_t lane_nb;     // k_lane on FP64 tables: candidates per batch (parked values per row), <= 32
  // CSC table (X == null): column f holds the entries csc_colptr[f] .. csc_colptr[f + 1] - 1 of (csc_row, csc_val)
  const int64_t *csc_colptr;
  const int32_t *csc_row;
  const double *csc_val;
  // ... and its row-major index (which features a row stores): csr_col[csr_ptr[r] .. csr_ptr[r + 1])
  const int64_t *csr_ptr;
  const int32_t *csr_col;
  int64_t inv_rows;  // trees of the batch x n (size of the chunked path's inverse index map)
};

// Column f of the table: dense FP64 values, or the stored entries of a CSC column (everything not stored is 0.0
// with dense semantics: the reference has no sparse Mat, a CSC table builds the forest of its dense expansion).
struct Col {
  const double *v;      // dense: X + f * ld;  CSC: the column's stored values
  const int32_t *rows;  // CSC: ascending row ids of the stored entries (null: dense)
  int32_t nnz;
};
__device__ __forceinline__ Col col_of(const P &p, int32_t f) {
  Col c;
  if (p.csc_row) {
    const int64_t a = __ldg(p.csc_colptr + f), e = __ldg(p.csc_colptr + f + 1);
    c.v = p.csc_val + a;
    c.rows = p.csc_row + a;
    c.nnz = (int32_t)(e - a);
  } else {
    c.v = p.X + (int64_t)f * p.ld;
    c.rows = nullptr;
    c.nnz = 0;
  }
  return c;
}
// value of row r: one gather, or a binary search among the column's stored rows (a miss is an implicit zero)
__device__ __forceinline__ double col_at(const Col &c, int32_t r) {
  if (!c.rows) return __ldg(c.v + r);
  int lo = 0, hi = c.nnz;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(c.rows + mid) < r)
      lo = mid + 1;
    else
      hi = mid;
  }
  return (lo < c.nnz && __ldg(c.rows + lo) == r) ? __ldg(c.v + lo) : 0.0;
}

// U rows of one column at once.  On a CSC column the U searches advance in lockstep (the halving steps depend only
// on the column's length), so their loads overlap: a one-team kernel's time on a sparse table is the chain of
// dependent row-id loads, and this cuts it by U.  r[u] < 0 = no row (x[u] = 0.0).
template <int U>
__device__ __forceinline__ void col_atn(const Col &c, const int32_t (&r)[U], double (&x)[U]) {
  if (!c.rows) {
#pragma unroll
    for (int u = 0; u < U; u++) x[u] = (r[u] >= 0) ? __ldg(c.v + r[u]) : 0.0;
    return;
  }
#pragma unroll
  for (int u = 0; u < U; u++) x[u] = 0.0;
  if (c.nnz <= 0) return;
  int32_t b[U];  // offsets into the column's stored rows
#pragma unroll
  for (int u = 0; u < U; u++) b[u] = 0;
  int len = c.nnz;
  while (len > 1) {  // first entry with row >= r[u] lies in [b[u], b[u] + len]
    const int half = len >> 1;
#pragma unroll
    for (int u = 0; u < U; u++) b[u] = (__ldg(c.rows + b[u] + half - 1) < r[u]) ? b[u] + half : b[u];
    len -= half;
  }
#pragma unroll
  for (int u = 0; u < U; u++) {
    const int32_t v0 = __ldg(c.rows + b[u]);
    const int32_t at = (v0 < r[u]) ? b[u] + 1 : b[u];
    if (r[u] >= 0 && at < c.nnz && ((v0 == r[u]) || (v0 < r[u] && __ldg(c.rows + at) == r[u]))) x[u] = __ldg(c.v + at);
  }
}
__device__ __forceinline__ void col_at4(const Col &c, const int32_t (&r)[4], double (&x)[4]) { col_atn<4>(c, r, x); }

__host__ __device__ inline int size_class(const P &p, int64_t n) {
  int q = 0;
  while (q < NQ - 1 && n > p.cls_max[q]) q++;
  return q;
}

// (kept out of line: the closed form is long and is called from several places of every node kernel)
static __device__ __noinline__ double repeat_add_dev(double c, int64_t h) { return et_repeat_add(c, h); }

// ---- team helpers ---------------------------------------------------------------------------
template <int TEAM>
__device__ __forceinline__ void team_sync() {
  if (TEAM == 32)
    __syncwarp();
  else
    __syncthreads();
}

// Sparse-resident table, free-running mode: a feature that none of the node's rows stores is all-zero in the
// node, hence constant (pkg:236) -- marked known-constant up front from the rows' stored features (CSR index)
// instead of being found by drawing it.  On a 1 % dense table a 10-row node stores ~1000 of 10000 features: the
// scored candidates of the reference are a uniform sample without replacement of the varying features either way
// (draws that hit a constant feature are discarded, pkg:236-239), so the split has the same distribution and the
// ~90 % of draws that would hit all-zero features are never made.  `rows` = the node's rows (shared or global).
template <int TEAM>
__device__ __forceinline__ void sparse_mark_constant(const P &p, const int32_t *rows, int32_t n, uint32_t *s_const,
                                                     uint32_t *s_taken, int W, int tid) {
  for (int w = tid; w < W; w += TEAM) s_taken[w] = 0xffffffffu;
  team_sync<TEAM>();
  for (int32_t j = 0; j < n; j++) {
    const int32_t r = rows[j];
    const int64_t a = __ldg(p.csr_ptr + r), e = __ldg(p.csr_ptr + r + 1);
    for (int64_t t = a + tid; t < e; t += TEAM) {
      const int32_t f = __ldg(p.csr_col + t);
      atomicAnd(&s_taken[f >> 5], ~(1u << (f & 31)));
    }
  }
  team_sync<TEAM>();
  for (int w = tid; w < W; w += TEAM) {
    const uint32_t v = s_taken[w] | s_const[w];  // (+ what the path from the root already excluded)
    s_const[w] = v;
    s_taken[w] = v;
  }
  team_sync<TEAM>();
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
static __device__ __noinline__ double gini_score_int(const int32_t *hnode, const int32_t *hl, const int32_t *hn, bool nan_left, int C,
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
static __device__ double var_reduction_seq(const double *y, int32_t n, const uint32_t *mlt, const uint32_t *mnan,
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
static __device__ double gini_score_w_seq(const int32_t *y, const double *w, int32_t n, const uint32_t *mlt,
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




// ---- host side: growing device buffers of the level loop --------------------------------------------------
struct FrontierBufs {
  DevBuf<int32_t> tree, begin, end, node, depth, hist;
  DevBuf<int64_t> trace;
  DevBuf<uint64_t> key;
  DevBuf<uint32_t> mask;
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
  void grow(size_t n, size_t used, size_t nleaf, size_t leaf_used, int lw, cudaStream_t st) {
    tree.grow_keep(n, used, st);
    feat.grow_keep(n, used, st);
    child.grow_keep(n, used, st);
    cut.grow_keep(n, used, st);
    leaf_vals.grow_keep(nleaf * (size_t)lw, leaf_used * (size_t)lw, st);
  }
  Pool view() { return Pool{tree.p, feat.p, child.p, cut.p, leaf_vals.p}; }
};

// ETGPU_TIMING=1|2: wall-clock time per kernel family with the size classes serialised (totals | per level)
struct PhaseTimer {
  enum { ALLOC = 0, INIT, LANE, CTA, SYNC, PREORDER, MID, WIDE, NPH };
  bool on = getenv("ETGPU_TIMING") != nullptr;
  bool per_level = on && atoi(getenv("ETGPU_TIMING")) >= 2;
  double last[NPH] = {0};
  cudaStream_t st = nullptr;
  double acc[NPH] = {0};
  std::chrono::steady_clock::time_point t0;
  void level_report(int level, const int32_t *qn) {
    if (!per_level) return;
    fprintf(stderr,
            "[etgpu level %3d] nodes n32=%d n64=%d n128=%d n256=%d n512=%d mid=%d cta=%d wide=%d | ms lane=%.3f mid=%.3f "
            "cta=%.3f wide=%.3f\n",
            level, qn[0], qn[1], qn[2], qn[3], qn[4], qn[5], qn[6], qn[7], acc[LANE] - last[LANE], acc[MID] - last[MID],
            acc[CTA] - last[CTA], acc[WIDE] - last[WIDE]);
    for (int i = 0; i < NPH; i++) last[i] = acc[i];
  }
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
    static const char *names[] = {"alloc", "init", "node_lane", "node_cta", "sync", "preorder", "node_mid", "node_wide"};
    fprintf(stderr, "[etgpu timing ms]");
    for (int i = 0; i < NPH; i++) fprintf(stderr, " %s=%.1f", names[i], acc[i]);
    fprintf(stderr, "\n");
  }
};

// CUDA-event spans of the node kernels (summed after the per-level sync)
struct EventTimer {
  std::vector<cudaEvent_t> pool;
  std::vector<std::pair<int, int>> spans[2];  // 0 = every level, 1 = levels that hold CTA-owned / chunked nodes
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

struct LevelCfg {
  bool coded = false;      // nodes of up to 512 samples: k_lane on byte codes
  bool coded_big = false;  // larger nodes: byte-coded CTA teams / chunks (unweighted classification, <= 32 classes)
  bool wide = false;       // nodes above P::wide_min rows are cut into chunks (wide.cu)
  size_t smem_warp = 0, smem_mid = 0, smem_cta = 0;  // k_node teams (per team)
  size_t smem_lane[5] = {0};                         // k_lane per warp, classes 0..4
};

// NVTX range on the calling thread (nvtx3, header only); compiled out with -DETGPU_NO_NVTX
struct NvtxRange {
  explicit NvtxRange(const char *name);
  ~NvtxRange();
};

// ---- entry points of the builder's translation units ------------------------------------------------------
struct BestBufs;  // best.cu
BestBufs *best_bufs_create();
void best_bufs_destroy(BestBufs *bb);
struct WideBufs;  // wide.cu
WideBufs *wide_bufs_create();
void wide_bufs_destroy(WideBufs *wb);

// node_inst.cu (one object per task): one level = one launch per non-empty size class, classes on concurrent
// streams; the chunked path of the wide nodes (wide.cu) runs on the main stream meanwhile
template <int TASK>
void launch_level(et_ctx *ctx, const P &p, const int32_t *qn, int64_t wide_rows, const LevelCfg &lc, PhaseTimer &pt,
                  EventTimer &et, WideBufs *wb);
template <int TASK>
void set_smem_attr(const LevelCfg &lc);
// best.cu: one level of the bestSplit builder (every open node sits in queue Q_CTA)
template <int TASK>
void launch_level_best(et_ctx *ctx, const P &p, int32_t count, BestBufs &bb, int64_t rows);
// wide.cu: all phases of the chunked path for the `count` nodes of queue Q_WIDE (TASK_CLS and TASK_REG)
template <int TASK>
void wide_level(et_ctx *ctx, const P &p, int32_t count, int64_t wide_rows, const LevelCfg &lc, WideBufs &wb,
                cudaStream_t st);

}  // namespace etb
