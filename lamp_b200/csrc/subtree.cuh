// subtree.cuh -- resident subtree builder (included by build.cu inside its anonymous namespace).
//
// Once a node is small enough that ALL d features of its rows fit in the shared memory of one SM, the
// whole subtree below it is built without touching the table in HBM / L2 again: the team (1..16 warps)
// copies the node's rows from the ROW-MAJOR copy of the table (byte codes or FP64; full 16-byte vector
// loads of contiguous rows instead of one 32-byte sector per gathered value per level), then its warps
// pull open nodes from a team queue in shared memory and run the reference's split search
// (splitClassification pkg:203-297 / splitRegression pkg:427-511, one lane per candidate feature exactly
// like k_lane) against the staged rows, partition a one-byte local row permutation in place (the child
// filters pkg:1024-1039 / 841-856) and push the children.  When the queue drains the team converts its nodes
// from creation order to pre-order in shared memory and writes one contiguous block of 16-byte PNodes; the
// level-wise builder only sees a marker node (feat == -2) that stands for the whole block.
//
// Synchronisation inside a team is a producer / consumer protocol on the `nd_ready` flags (record and permutation
// writes, __threadfence_block, flag store | flag load, __threadfence_block, __syncwarp, reads) and team barriers
// only between the phases; compute-sanitizer racecheck knows barriers only and reports the flag-ordered accesses as
// hazards (memcheck is clean, and every parity test runs with this builder on).
//
// Table traffic: rows x row bytes ONCE per subtree, against rows x (features examined) x 32-byte sectors
// per LEVEL of the subtree for the gathering kernels.
#pragma once

constexpr int SUB_WARPS = 16;     // warps per CTA; a CTA holds SUB_WARPS / TEAMW teams
constexpr int SUB_MARK = -2;      // Pool::feat of a marker node
constexpr int SUB_NCLS = 5;       // team widths 1, 2, 4, 8, 16 warps
constexpr int SUB_NC_MAX = 64;    // nodes up to this many rows read off which features vary before drawing

struct SubLay {
  int T, NWT, NN;      // rows per team, bitmask words per node, node records per team
  int rowbytes;        // bytes of one staged row (multiple of 16)
  // team region (byte offsets)
  int o_tab, o_key, o_cut, o_ytab, o_wtab, o_be, o_depth, o_feat, o_child, o_mask, o_ctl, o_lvl, o_sz, o_nl, o_pos,
      o_perm, o_lab, o_ready, o_coff, team_bytes;
  // per-warp scratch (byte offsets)
  int w_y, w_dist, w_cu, w_lt, w_nn, w_best, w_cm, w_hnode, w_taken, w_nc, w_cf, warp_bytes;
};

__host__ __device__ inline SubLay make_sublay(int task, int C, int W, bool replay, int teamw, int rw, int rowbytes,
                                              bool coded) {
  SubLay L;
  L.T = teamw * rw;
  L.NWT = (L.T + 31) / 32;
  L.NN = 2 * L.T;  // a subtree over T rows has at most 2T - 1 nodes
  L.rowbytes = rowbytes;
  int o = 0;
  L.o_tab = o;
  o += L.T * rowbytes;  // 16-byte multiple
  L.o_key = o;
  o += L.NN * 8;
  L.o_cut = o;
  o += L.NN * 8;
  L.o_ytab = o;
  o += (task == TASK_REG) ? L.T * 8 : 0;
  L.o_wtab = o;
  o += (task == TASK_CLSW) ? L.T * 8 : 0;
  L.o_be = o;
  o += L.NN * 4;
  L.o_depth = o;
  o += L.NN * 4;
  L.o_feat = o;
  o += L.NN * 4;
  L.o_child = o;
  o += L.NN * 4;
  L.o_mask = o;
  o += replay ? 0 : (L.T / 2 + 1) * W * 4;
  o = (o + 7) / 8 * 8;
  L.o_ctl = o;
  o += 16 * 4;  // head, tail, processed, max level, block base, 8 x 64-bit statistics (from word 4)... see SUBC_*
  o += 8 * 8;
  L.o_lvl = o;
  o += L.NN * 2;
  L.o_sz = o;
  o += L.NN * 2;
  L.o_nl = o;
  o += L.NN * 2;
  L.o_pos = o;
  o += L.NN * 2;
  L.o_perm = o;
  o += L.T;
  L.o_lab = o;
  o += (task == TASK_REG) ? 0 : L.T;
  L.o_ready = o;
  o += L.NN;
  o = (o + 15) / 16 * 16;
  L.o_coff = o;  // per feature: 0 if the column holds NaNs, else 1 (byte-coded tables)
  o += coded ? W * 32 : 0;
  L.team_bytes = ((o + 15) / 16) * 16;
  // per-warp scratch
  int w = 0;
  L.w_y = w;
  w += (task == TASK_CLS) ? 0 : L.NWT * 32 * 8;  // regression targets / weights by position
  L.w_dist = w;
  w += (task == TASK_REG) ? 0 : C * 8;
  L.w_cu = w;  // threshold uniforms of the collected candidates
  w += (coded && !replay) ? 32 * 8 : 0;
  L.w_lt = w;
  w += L.NWT * 32 * 4;
  L.w_nn = w;
  w += L.NWT * 32 * 4;
  L.w_best = w;
  w += L.NWT * 4;
  L.w_cm = w;
  w += (task == TASK_REG) ? 0 : C * L.NWT * 4;
  L.w_hnode = w;
  w += (task == TASK_REG) ? 0 : C * 4;
  L.w_taken = w;
  w += replay ? 0 : W * 4;
  L.w_nc = w;  // features that are NOT constant over the node's rows
  w += (coded && !replay) ? W * 4 : 0;
  L.w_cf = w;  // collected candidate features
  w += (coded && !replay) ? 32 * 4 : 0;
  L.warp_bytes = ((w + 15) / 16) * 16;
  return L;
}

// shared memory of one CTA (= one team) of k_sub for team width `teamw`
__host__ __device__ inline size_t sub_smem_bytes(int task, int C, int W, bool replay, int teamw, int rw, int rowbytes,
                                                 bool coded) {
  const SubLay L = make_sublay(task, C, W, replay, teamw, rw, rowbytes, coded);
  return (size_t)L.team_bytes + (size_t)teamw * (size_t)L.warp_bytes;
}

enum { SUBC_HEAD = 0, SUBC_TAIL, SUBC_DONE, SUBC_MAXLVL, SUBC_BASE, SUBC_LEAVES };
enum { SUBS_SROWS = 0, SUBS_VMM, SUBS_VSC, SUBS_DRAWS, SUBS_CONST, SUBS_SCORED, SUBS_MISMATCH, SUBS_PROWS };

template <int TEAMW>
__device__ __forceinline__ void sub_team_bar(int) {
  if (TEAMW == 1)
    __syncwarp();
  else
    __syncthreads();
}

__device__ __forceinline__ int sub_ld_volatile(const int *p) { return *(const volatile int *)p; }

template <int TASK, typename VT, int TEAMW>
__global__ void __launch_bounds__(32 * TEAMW, SUB_WARPS / TEAMW) k_sub(P p, int32_t qcount, int qi) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr bool CODED = (sizeof(VT) != 8);
  constexpr uint32_t FULL = 0xffffffffu;
  constexpr int TT = TEAMW * 32;  // one CTA = one team = one subtree; several CTAs share an SM
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int team = 0, wit = warp, tid = threadIdx.x;
  const int q = blockIdx.x;
  if (q >= qcount) return;
  const int C = p.C, W = p.W;
  const bool replay = p.replay != 0;
  const SubLay L = make_sublay(TASK, C, W, replay, TEAMW, p.sub_rw, p.sub_rowbytes, CODED);
  const int NW = L.NWT;
  unsigned char *tm = smem_raw;
  unsigned char *wm = smem_raw + (size_t)L.team_bytes + (size_t)warp * L.warp_bytes;
  // team region
  const unsigned char *tab = tm + L.o_tab;
  uint64_t *nd_key = reinterpret_cast<uint64_t *>(tm + L.o_key);
  double *nd_cut = reinterpret_cast<double *>(tm + L.o_cut);
  double *ytab = reinterpret_cast<double *>(tm + L.o_ytab);
  double *wtab = reinterpret_cast<double *>(tm + L.o_wtab);
  uint32_t *nd_be = reinterpret_cast<uint32_t *>(tm + L.o_be);
  int32_t *nd_depth = reinterpret_cast<int32_t *>(tm + L.o_depth);
  int32_t *nd_feat = reinterpret_cast<int32_t *>(tm + L.o_feat);
  int32_t *nd_child = reinterpret_cast<int32_t *>(tm + L.o_child);
  uint32_t *masks = reinterpret_cast<uint32_t *>(tm + L.o_mask);
  int *ctl = reinterpret_cast<int *>(tm + L.o_ctl);
  unsigned long long *tstat = reinterpret_cast<unsigned long long *>(tm + L.o_ctl + 64);
  uint16_t *nd_lvl = reinterpret_cast<uint16_t *>(tm + L.o_lvl);
  uint16_t *nd_sz = reinterpret_cast<uint16_t *>(tm + L.o_sz);
  uint16_t *nd_nl = reinterpret_cast<uint16_t *>(tm + L.o_nl);
  uint16_t *nd_pos = reinterpret_cast<uint16_t *>(tm + L.o_pos);
  uint8_t *perm = tm + L.o_perm;
  uint8_t *lab = tm + L.o_lab;
  volatile uint8_t *nd_ready = tm + L.o_ready;
  uint8_t *s_coff = tm + L.o_coff;
  // warp scratch
  double *s_y = reinterpret_cast<double *>(wm + L.w_y);
  double *s_dist = reinterpret_cast<double *>(wm + L.w_dist);
  uint32_t *s_lt = reinterpret_cast<uint32_t *>(wm + L.w_lt);
  uint32_t *s_nn = reinterpret_cast<uint32_t *>(wm + L.w_nn);
  uint32_t *s_best = reinterpret_cast<uint32_t *>(wm + L.w_best);
  uint32_t *s_cm = reinterpret_cast<uint32_t *>(wm + L.w_cm);
  int32_t *s_hnode = reinterpret_cast<int32_t *>(wm + L.w_hnode);
  uint32_t *s_taken = reinterpret_cast<uint32_t *>(wm + L.w_taken);
  uint32_t *s_nc = reinterpret_cast<uint32_t *>(wm + L.w_nc);
  int32_t *s_cf = reinterpret_cast<int32_t *>(wm + L.w_cf);
  double *s_cu = reinterpret_cast<double *>(wm + L.w_cu);

  const int i = p.q_cur[qi][q];
  const int32_t tree = p.cur.tree[i], b0 = p.cur.begin[i], n0 = p.cur.end[i] - b0;
  const int32_t node = p.cur.node[i];
  const int64_t base = (int64_t)tree * p.n;
  const int lw = (TASK == TASK_REG) ? 1 : C;
  const int rowbytes = L.rowbytes;

  // ---------------- stage the subtree's rows (row-major copy of the table: contiguous 16-byte loads) --------
  {
    const int32_t *idx = p.idx_src + base + b0;
    const unsigned char *src_tab = CODED ? reinterpret_cast<const unsigned char *>(p.R8)
                                         : reinterpret_cast<const unsigned char *>(p.XR);
    const int nvec = rowbytes >> 4;
    for (int r0 = wit; r0 < n0; r0 += 4 * TEAMW) {
      const uint4 *src[4];
      uint4 *dst[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int r = r0 + u * TEAMW;
        const int32_t rid = (r < n0) ? __ldg(idx + r) : -1;
        src[u] = (rid >= 0) ? reinterpret_cast<const uint4 *>(src_tab + (int64_t)rid * rowbytes) : nullptr;
        dst[u] = reinterpret_cast<uint4 *>(tm + L.o_tab + (size_t)r * rowbytes);
      }
      for (int c = lane; c < nvec; c += 32) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (src[u]) v[u] = __ldg(src[u] + c);
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (src[u]) dst[u][c] = v[u];
      }
    }
    for (int r = tid; r < n0; r += TT) {
      perm[r] = (uint8_t)r;
      if (TASK != TASK_REG) lab[r] = (uint8_t)p.yc_src[base + b0 + r];
      if (TASK == TASK_REG) ytab[r] = p.yr_src[base + b0 + r];
      if (TASK == TASK_CLSW) wtab[r] = p.w_src[base + b0 + r];
    }
    for (int r = tid; r < L.NN; r += TT) nd_ready[r] = 0;
    if (CODED)
      for (int f = tid; f < W * 32; f += TT) s_coff[f] = (f < p.d) ? __ldg(p.coff + f) : (uint8_t)1;
    if (tid < 8) tstat[tid] = 0ull;
    if (!replay)
      for (int w = tid; w < W; w += TT) masks[w] = p.cur.mask[(int64_t)i * W + w];
    if (tid == 0) {
      ctl[SUBC_HEAD] = 0;
      ctl[SUBC_TAIL] = 1;
      ctl[SUBC_DONE] = 0;
      ctl[SUBC_MAXLVL] = 0;
      ctl[SUBC_LEAVES] = 0;
      nd_be[0] = (uint32_t)n0 << 16;
      nd_depth[0] = p.cur.depth[i];
      nd_key[0] = replay ? (uint64_t)p.cur.trace[i] : p.cur.key[i];
      nd_lvl[0] = 0;
    }
    sub_team_bar<TEAMW>(team);
    if (tid == 0) nd_ready[0] = 1;
  }

  // ---------------- the team's warps pull open nodes until the subtree is complete ----------------
  for (;;) {
    int h = -1;
    if (lane == 0) {
      for (;;) {
        const int done = sub_ld_volatile(&ctl[SUBC_DONE]);  // (read BEFORE the tail: done == tail then means
        const int tail = sub_ld_volatile(&ctl[SUBC_TAIL]);  //  nothing was in flight, so nothing more can come)
        const int head = sub_ld_volatile(&ctl[SUBC_HEAD]);
        if (head < tail) {
          if (atomicCAS(&ctl[SUBC_HEAD], head, head + 1) == head) {
            h = head;
            break;
          }
          continue;
        }
        if (done == tail) break;
        if (TEAMW > 1) __nanosleep(64);
      }
      if (h >= 0) {
        while (!nd_ready[h]) {
        }
        __threadfence_block();
      }
    }
    h = __shfl_sync(FULL, h, 0);
    if (h < 0) break;
    __syncwarp();  // lane 0's acquire (ready flag + fence) orders the whole warp's reads of the node record

    const uint32_t be = nd_be[h];
    const int b = (int)(be & 0xffffu), e = (int)(be >> 16), n = e - b;
    const int32_t depth = nd_depth[h];
    const uint64_t key = nd_key[h];
    const int64_t tn = (int64_t)key;
    const int nw = (n + 31) >> 5;
    const uint8_t *pb = perm + b;
    uint32_t *s_const = masks + (size_t)(b >> 1) * W;  // the node's inherited known-constant mask (n >= 2)

    // ---- labels / targets by position
    if (TASK != TASK_REG) {
      for (int t = lane; t < C * NW; t += 32) s_cm[t] = 0u;
      __syncwarp();
      for (int w = 0; w < nw; w++) {
        const int j = w * 32 + lane;
        const bool has = j < n;
        const int32_t cls = has ? (int32_t)lab[pb[j]] : -1;
        const uint32_t grp = __match_any_sync(FULL, cls);
        if (has && lane == __ffs(grp) - 1) s_cm[cls * NW + w] = grp;
      }
    }
    if (TASK == TASK_REG)
      for (int j = lane; j < n; j += 32) s_y[j] = ytab[pb[j]];
    if (TASK == TASK_CLSW)
      for (int j = lane; j < n; j += 32) s_y[j] = wtab[pb[j]];
    __syncwarp();

    // ---- stop rules + node totals (as k_lane)
    bool leaf;
    double total = 0.0, nsum = (double)n, leaf_mean = 0.0;
    if (TASK != TASK_REG) {
      bool pure_l = false;
      for (int c = lane; c < C; c += 32) {
        int32_t hh = 0;
        for (int w = 0; w < nw; w++) hh += __popc(s_cm[c * NW + w]);
        s_hnode[c] = hh;
        pure_l |= (hh == n);
      }
      const bool pure = __any_sync(FULL, pure_l);
      leaf = (p.n_table < p.n_min) || (depth >= p.max_depth) || pure;  // pkg:993-994
      __syncwarp();
    }
    if (TASK == TASK_CLS) {
      if (!leaf) {
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

    // ---- split search: one lane per candidate, values read from the staged rows
    int32_t visited = 0, nconst = 0, best_feature = -1, best_mil = 0;
    double best_score = -INFINITY, best_cut = NAN;
    unsigned long long st_draws = 0, st_const = 0, st_scored = 0, st_mismatch = 0;
    if (!leaf) {
      int32_t dc = 0, tpos = 0, tcnt = 0;
      int64_t tb = 0;
      if (replay) {
        if (tn >= 0) {
          tb = p.tr.cand_begin[tn];
          tcnt = p.tr.cand_count[tn];
        }
      } else {
        int nc = 0;
        for (int w = lane; w < W; w += 32) {
          const uint32_t m = s_const[w];
          s_taken[w] = m;
          nc += __popc(m);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nc += __shfl_xor_sync(FULL, nc, o);
        nconst = nc - (W * 32 - p.d);
        __syncwarp();
      }
      // Byte-coded tables, free-running: which features vary over the node's rows is read off the staged rows
      // directly (four features per word: OR of the XORs with the first row; a NaN byte makes the feature
      // non-constant like hasMissing does, pkg:236).  Draws that hit a constant feature are then settled without
      // walking the rows, and only the others are collected for the scoring passes below.
      const bool use_nc = CODED && !replay && n <= SUB_NC_MAX;
      if (use_nc) {
        // lane l owns table words l, l + 32, ... (4 features each); feature word 4 i + (l >> 3) is assembled from
        // the 8 lanes l & ~7 .. (l & ~7) + 7 of round i
        const int nword = rowbytes >> 2;
        const uint32_t *trow0 = reinterpret_cast<const uint32_t *>(tab + (uint32_t)pb[0] * rowbytes);
        for (int i8 = 0; i8 * 4 < W; i8++) {
          const int tw = i8 * 32 + lane;
          uint32_t bits = 0u;
          if (tw < nword) {
            const uint32_t first = trow0[tw];
            uint32_t acc = 0u;
#pragma unroll 4
            for (int j = 1; j < n; j++)
              acc |= reinterpret_cast<const uint32_t *>(tab + (uint32_t)pb[j] * rowbytes)[tw] ^ first;
            const uint32_t k4 = *reinterpret_cast<const uint32_t *>(s_coff + tw * 4);  // 0 = column holds NaNs
            const uint32_t nan4 = __vcmpeq4(first, 0u) & __vcmpeq4(k4, 0u);             // first row is NaN there
            const uint32_t ne4 = __vcmpne4(acc, 0u) | nan4;                            // 0xff per varying feature
            bits = ((ne4 & 0x01010101u) * 0x01020408u) >> 24;                          // 4 bits, feature order
          }
          uint32_t word = bits << (4 * (lane & 7));
          word |= __shfl_xor_sync(FULL, word, 1);
          word |= __shfl_xor_sync(FULL, word, 2);
          word |= __shfl_xor_sync(FULL, word, 4);
          const int w0 = i8 * 4 + (lane >> 3);
          if ((lane & 7) == 0 && w0 < W) s_nc[w0] = word;
        }
        __syncwarp();
      }
      for (;;) {
        int32_t nb;
        const int32_t avail = p.d - nconst - visited;
        if (replay) {
          nb = min(32, tcnt - tpos);
        } else if (use_nc) {
          nb = (min(p.k - visited, avail) > 0) ? 32 : 0;
        } else {
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
        int32_t f = -1;
        double u = 0.0;
        int expect = 0;
        if (replay) {
          if (lane < nb) {
            f = p.tr.cand_feature[tb + tpos + lane];
            u = p.tr.cand_u[tb + tpos + lane];
            expect = p.tr.cand_flag[tb + tpos + lane] + 1;
          }
          tpos += nb;
        } else if (use_nc) {
          // rounds of 32 draws in draw order until enough varying features are collected (or none are left):
          // constant hits before the cutoff are counted and marked like the reference does (pkg:236-239),
          // draws past the cutoff were never made (their features stay available)
          const int32_t want = min(32, p.k - visited);
          int32_t ncol = 0, left = avail;
          while (ncol < want && left > 0) {
            const int32_t nd = min(32, left);
            int32_t pick = -1 - lane;
            double uu = 0.0;
            if (lane < nd) {
              const uint64_t r = et_draw(key, (uint32_t)(dc + 2 * lane));
              pick = rank_select_clear_fast(s_taken, W, (int32_t)__umul64hi(r, (uint64_t)left));
              uu = et_u01(et_draw(key, (uint32_t)(dc + 2 * lane + 1)));
            }
            dc += 64;
            const uint32_t same = __match_any_sync(FULL, pick);
            const bool drawn = lane < nd && lane == __ffs(same) - 1;
            const bool varies = drawn && ((s_nc[pick >> 5] >> (pick & 31)) & 1u);
            const uint32_t m_var = __ballot_sync(FULL, varies);
            const int ord = __popc(m_var & ((1u << lane) - 1u));
            const bool acc = drawn && ord < want - ncol;  // made before the (want - ncol)-th varying draw
            const uint32_t m_acc = __ballot_sync(FULL, acc);
            if (acc) {
              atomicOr(&s_taken[pick >> 5], 1u << (pick & 31));
              if (varies) {
                s_cf[ncol + ord] = pick;
                s_cu[ncol + ord] = uu;
              } else {
                atomicOr(&s_const[pick >> 5], 1u << (pick & 31));
              }
            }
            const int nvar = __popc(m_var & m_acc), nacc = __popc(m_acc);
            ncol += nvar;
            left -= nacc;
            nconst += nacc - nvar;
            st_draws += nacc - nvar;
            st_const += nacc - nvar;
            __syncwarp();
          }
          if (ncol == 0) break;  // every feature left was constant
          if (lane < ncol) {
            f = s_cf[lane];
            u = s_cu[lane];
          }
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
        const int fcol = act0 ? f : 0;
        // ---- pass 1: min / max / hasMissing over the node's rows (pkg:34-54)
        double mn = 1.7976931348623157e308, mx = -1.7976931348623157e308;  // pkg:35-36
        bool has_nan = false;
        const uint32_t K = (CODED && act0 && s_coff[f] == 0) ? 1u : 0u;
        const bool nan_cols = CODED ? (__any_sync(FULL, K != 0u) != 0) : true;
        uint32_t mnt = 0xffffffffu, mxb = 0u;
        const unsigned char *tcol = tab + (CODED ? fcol : fcol * 8);
        if (CODED) {
          int j = 0;
          for (; j + 4 <= n; j += 4) {
            const uint32_t r0 = pb[j], r1 = pb[j + 1], r2 = pb[j + 2], r3 = pb[j + 3];
            const uint32_t v0 = tcol[r0 * rowbytes], v1 = tcol[r1 * rowbytes], v2 = tcol[r2 * rowbytes],
                           v3 = tcol[r3 * rowbytes];
            mxb = max(max(mxb, v0), max(v1, max(v2, v3)));
            mnt = min(min(mnt, v0 - K), min(v1 - K, min(v2 - K, v3 - K)));
          }
          for (; j < n; j++) {
            const uint32_t v0 = tcol[(uint32_t)pb[j] * rowbytes];
            mxb = max(mxb, v0);
            mnt = min(mnt, v0 - K);
          }
        } else {
#pragma unroll 4
          for (int j = 0; j < n; j++) {
            const double x = *reinterpret_cast<const double *>(tcol + (uint32_t)pb[j] * rowbytes);
            if (x < mn) mn = x;
            if (x > mx) mx = x;
            has_nan |= (x != x);
          }
        }
        uint32_t thr = 0u;
        const uint32_t wmax = CODED ? ((K == 1u) ? mxb : mxb + 1u) : 0u;  // largest wide code; 0 = only NaNs
        if (CODED) {
          if (act0 && wmax != 0u) {
            const double *dc8 = p.dict + (int64_t)f * 256;
            mn = __ldg(dc8 + mnt);
            mx = __ldg(dc8 + (wmax - 1u));
          }
        }
        const double cut = ET_ADD(mn, ET_MUL(ET_SUB(mx, mn), u));  // pkg:240
        if (CODED) {
          if (act0 && wmax != 0u && !(mx <= mn)) {
            // thr = number of dictionary entries below the cutpoint (all of dict[0, mnt) are, none past wmax - 1):
            // 8 probes per round keep the dependent trips to L2 at <= 3
            const double *dc8 = p.dict + (int64_t)f * 256;
            uint32_t lo = mnt, hi = wmax;  // invariant: dict[lo - 1] < cut (or lo == mnt), dict[hi] >= cut (or hi == wmax)
            while (lo < hi) {
              const uint32_t span = hi - lo, step = (span + 8u) / 9u;  // probes at lo + step * (1..8) - 1
              double pv[8];
#pragma unroll
              for (int t = 0; t < 8; t++) {
                const uint32_t at = lo + step * (uint32_t)(t + 1) - 1u;
                pv[t] = (at < hi) ? __ldg(dc8 + at) : INFINITY;
              }
              uint32_t below = 0;  // probes are ascending: count those < cut
#pragma unroll
              for (int t = 0; t < 8; t++) below += (pv[t] < cut) ? 1u : 0u;
              const uint32_t nlo = lo + step * below;  // dict[nlo - 1] < cut when below > 0
              const uint32_t nhi = (below < 8u) ? min(hi, lo + step * (below + 1u) - 1u) : hi;
              lo = nlo;
              hi = nhi;
            }
            thr = lo;
          }
        }
        // ---- pass 2: side bitmasks over the positions
        for (int w = 0; w < nw; w++) {
          const int j0 = w << 5, cnt = min(32, n - j0);
          uint32_t lt = 0u, nn = 0u;
          if (CODED) {
            const uint32_t tm1 = thr - 1u;  // thr == 0: nothing is below the cutpoint (en = 0)
            const bool en = thr > 0u;
#pragma unroll 4
            for (int jj = 0; jj < cnt; jj++) {
              const uint32_t v = tcol[(uint32_t)pb[j0 + jj] * rowbytes];
              const uint32_t t8 = (v - K) & 0xffu;
              lt |= (uint32_t)(en && t8 <= tm1) << jj;
              if (nan_cols) nn |= (uint32_t)(K != 0u && v == 0u) << jj;
            }
          } else {
#pragma unroll 4
            for (int jj = 0; jj < cnt; jj++) {
              const double x = *reinterpret_cast<const double *>(tcol + (uint32_t)pb[j0 + jj] * rowbytes);
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
        // ---- consume the batch in draw (lane) order
        const bool counted0 = act0 && !const0 && !(s != s);
        const uint32_t m_cnt0 = __ballot_sync(FULL, counted0);
        const bool act = act0 && (replay || __popc(m_cnt0 & ((1u << lane) - 1u)) < p.k - visited);
        const bool is_const = act && const0;
        const bool is_nan = act && !const0 && (s != s);
        const bool counted = act && counted0;
        const uint32_t m_act = __ballot_sync(FULL, act);
        const uint32_t m_const = __ballot_sync(FULL, is_const);
        const uint32_t m_nan = __ballot_sync(FULL, is_nan);
        const uint32_t m_cnt = __ballot_sync(FULL, counted);
        if (replay) {
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
        if (!replay && (is_const || is_nan)) atomicOr(&s_const[f >> 5], 1u << (f & 31));
        visited += __popc(m_cnt);
        nconst += __popc(m_const) + __popc(m_nan);
        st_draws += __popc(m_act);
        st_const += __popc(m_const);
        st_scored += __popc(m_cnt) + __popc(m_nan);
        __syncwarp();
      }
    }

    // ---- finalize the node
    const bool make_leaf = leaf || best_feature < 0;
    if (lane == 0) {
      if (!leaf) {
        atomicAdd(&tstat[SUBS_SROWS], (unsigned long long)n);
        atomicAdd(&tstat[SUBS_VMM], (unsigned long long)n * st_draws);
        atomicAdd(&tstat[SUBS_VSC], (unsigned long long)n * st_scored);
        atomicAdd(&tstat[SUBS_DRAWS], st_draws);
        atomicAdd(&tstat[SUBS_CONST], st_const);
        atomicAdd(&tstat[SUBS_SCORED], st_scored);
      }
      if (replay) {
        const bool trace_split = tn >= 0 && p.tr.left[tn] >= 0;
        if (trace_split == make_leaf) st_mismatch++;
        if (st_mismatch) atomicAdd(&tstat[SUBS_MISMATCH], st_mismatch);
      }
    }
    if (make_leaf) {
      int32_t ls = 0;
      if (lane == 0) {
        ls = atomicAdd(&p.cnt->n_leaves, 1);
        atomicAdd(&ctl[SUBC_LEAVES], 1);
        nd_feat[h] = -1;
        nd_child[h] = ls;
        nd_cut[h] = NAN;
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
    } else {
      int32_t nl = 0;
      for (int w = lane; w < nw; w += 32) nl += __popc(s_best[w]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) nl += __shfl_xor_sync(FULL, nl, o);
      // the winner's side bitmask is the stable partition of the node's rows (pkg:1024-1039), in place
      {
        uint8_t rv[8];
#pragma unroll
        for (int w = 0; w < 8; w++) {
          const int j = w * 32 + lane;
          rv[w] = (w < nw && j < n) ? pb[j] : (uint8_t)0;
        }
        __syncwarp();
        int32_t lpos = b, rpos = b + nl;
        const uint32_t below = (1u << lane) - 1u;
#pragma unroll
        for (int w = 0; w < 8; w++) {
          if (w < nw) {
            const int j = w * 32 + lane;
            const int cnt = min(32, n - w * 32);
            const uint32_t valid = (cnt >= 32) ? FULL : ((1u << cnt) - 1u);
            const uint32_t bm = s_best[w];
            const uint32_t lm = bm & valid, rm = ~bm & valid;
            if (j < n) {
              const bool left = (lm >> lane) & 1u;
              const int32_t dst = left ? lpos + __popc(lm & below) : rpos + __popc(rm & below);
              perm[dst] = rv[w];
            }
            lpos += __popc(lm);
            rpos += __popc(rm);
          }
        }
      }
      int32_t cid = 0;
      if (lane == 0) {
        cid = atomicAdd(&ctl[SUBC_TAIL], 2);
        nd_feat[h] = best_feature | (best_mil ? ET_MIL_BIT : 0);
        nd_child[h] = cid;
        nd_cut[h] = best_cut;
        const int lvl = (int)nd_lvl[h] + 1;
        atomicMax(&ctl[SUBC_MAXLVL], lvl);
#pragma unroll
        for (int side = 0; side < 2; side++) {
          const int c2 = cid + side;
          nd_be[c2] = side ? ((uint32_t)(b + nl) | ((uint32_t)e << 16)) : ((uint32_t)b | ((uint32_t)(b + nl) << 16));
          // pkg:870 / 884 (sic): the regression right child keeps currentDepth; pkg:1055,1071: +1 both
          nd_depth[c2] = (TASK == TASK_REG && side) ? depth : depth + 1;
          uint64_t ck;
          if (replay)
            ck = (uint64_t)((tn >= 0) ? (int64_t)(side ? p.tr.right[tn] : p.tr.left[tn]) : (int64_t)-1);
          else
            ck = et_child_key(key, side);
          nd_key[c2] = ck;
          nd_lvl[c2] = (uint16_t)lvl;
        }
        atomicAdd(&tstat[SUBS_PROWS], (unsigned long long)n);
      }
      // the right child's copy of the known-constant mask (the left child inherits the parent's slot in place)
      if (!replay && (n - nl) >= 2 && ((b + nl) >> 1) != (b >> 1)) {
        uint32_t *mr = masks + (size_t)((b + nl) >> 1) * W;
        for (int w = lane; w < W; w += 32) mr[w] = s_const[w];
      }
      __threadfence_block();  // perm, masks and the child records are visible before the children are
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();  // ... published
        nd_ready[cid] = 1;
        nd_ready[cid + 1] = 1;
      }
    }
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();
      atomicAdd(&ctl[SUBC_DONE], 1);
    }
  }
  sub_team_bar<TEAMW>(team);

  // ---------------- creation order -> pre-order inside the team, then one contiguous block of PNodes ----------
  const int NNc = ctl[SUBC_TAIL];
  const int maxlvl = ctl[SUBC_MAXLVL];
  for (int v = tid; v < NNc; v += TT) {
    const bool lf = nd_feat[v] < 0;
    nd_sz[v] = lf ? 1 : 0;
    nd_nl[v] = lf ? 1 : 0;
  }
  sub_team_bar<TEAMW>(team);
  for (int lv = maxlvl - 1; lv >= 0; lv--) {  // subtree sizes, deepest internal nodes first
    for (int v = tid; v < NNc; v += TT) {
      if (nd_lvl[v] == lv && nd_feat[v] >= 0) {
        const int c = nd_child[v];
        nd_sz[v] = (uint16_t)(1 + nd_sz[c] + nd_sz[c + 1]);
        nd_nl[v] = (uint16_t)(nd_nl[c] + nd_nl[c + 1]);
      }
    }
    sub_team_bar<TEAMW>(team);
  }
  if (tid == 0) {
    nd_pos[0] = 0;
    ctl[SUBC_BASE] = (int)atomicAdd(&p.cnt->sub_nodes, (unsigned int)NNc);
  }
  sub_team_bar<TEAMW>(team);
  for (int lv = 0; lv < maxlvl; lv++) {  // pre-order positions, top down
    for (int v = tid; v < NNc; v += TT) {
      if (nd_lvl[v] == lv && nd_feat[v] >= 0) {
        const int c = nd_child[v];
        nd_pos[c] = (uint16_t)(nd_pos[v] + 1);
        nd_pos[c + 1] = (uint16_t)(nd_pos[v] + 1 + nd_sz[c]);
      }
    }
    sub_team_bar<TEAMW>(team);
  }
  const int32_t blk = ctl[SUBC_BASE];
  for (int v = tid; v < NNc; v += TT) {
    PNode pn;
    const int32_t ft = nd_feat[v];
    pn.cut = nd_cut[v];
    pn.feat = ft;
    pn.right_or_leaf = (ft >= 0) ? (int32_t)nd_pos[nd_child[v] + 1] : nd_child[v];  // block-relative | leaf pool slot
    p.sub_nodes[(int64_t)blk + nd_pos[v]] = pn;
  }
  if (tid == 0) {
    p.o.feat[node] = SUB_MARK;
    p.o.child[node] = blk;
    p.o.cut[node] = __longlong_as_double((long long)(((unsigned long long)(uint32_t)nd_nl[0] << 32) | (uint32_t)NNc));
    p.o.tree[node] = tree;
  }
  if (tid < 8 && tstat[tid]) {
    static_assert(SUBS_SROWS == 0 && SUBS_PROWS == 7, "statistics order");
    const int map[8] = {ST_SROWS, ST_VMM, ST_VSC, ST_DRAWS, ST_CONST, ST_SCORED, ST_MISMATCH, ST_PROWS};
    atomicAdd(&p.cnt->st[map[tid]], tstat[tid]);
  }
}

// Copies the blocks of the resident subtrees to their place in the pre-order forest: one warp per marker node.
// Block nodes are in pre-order already; right-child positions are relocated by the marker's position, leaves
// get consecutive forest-wide leaf indices (their values move from the leaf pool to the compact leaf table).
__global__ void __launch_bounds__(256) k_scatter_sub(Pool o, int32_t n_nodes, int lw, const int32_t *pos,
                                                      const int32_t *lpos, const int64_t *tree_off,
                                                      const int64_t *leaf_off, int64_t node_base, int64_t leaf_base,
                                                      const PNode *sub_nodes, PNode *nodes, double *leaves) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t v = warp_global * 32 + lane;
  const bool mark = v < n_nodes && o.feat[v] == SUB_MARK;
  uint32_t todo = __ballot_sync(0xffffffffu, mark);
  while (todo) {
    const int src_lane = __ffs(todo) - 1;
    todo &= todo - 1;
    const int64_t mv = warp_global * 32 + src_lane;
    const int t = o.tree[mv];
    const int32_t blk = o.child[mv];
    const unsigned long long packed = (unsigned long long)__double_as_longlong(o.cut[mv]);
    const int32_t cnt = (int32_t)(packed & 0xffffffffu);
    const int32_t p0 = pos[mv];
    const int64_t g0 = tree_off[t] - node_base + p0;
    int64_t gl = leaf_off[t] + lpos[mv];  // forest-wide index of the block's first leaf
    for (int j0 = 0; j0 < cnt; j0 += 32) {
      const int j = j0 + lane;
      PNode pn;
      pn.feat = 0;
      if (j < cnt) pn = sub_nodes[(int64_t)blk + j];
      const bool is_leaf = (j < cnt) && pn.feat < 0;
      const uint32_t lm = __ballot_sync(0xffffffffu, is_leaf);
      if (j < cnt) {
        if (is_leaf) {
          const int64_t my = gl + __popc(lm & ((1u << lane) - 1u));
          const double *src = o.leaf_vals + (int64_t)pn.right_or_leaf * lw;
          double *dst = leaves + (my - leaf_base) * lw;
          for (int c = 0; c < lw; c++) dst[c] = src[c];
          pn.right_or_leaf = (int32_t)my;
        } else {
          pn.right_or_leaf += p0;
        }
        nodes[g0 + j] = pn;
      }
      gl += __popc(lm);
    }
  }
}
