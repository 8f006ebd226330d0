// node.cuh -- the node kernels of the level-wise builder: one team (warp or CTA) owns a node end to end.
//
//   k_node<TASK, TEAM, CODED>   TEAM == 32: a warp, several per CTA; else one CTA per node.  Implements buildTree*
//                               (stop rules, pkg:993-994 / 813-814), split* (pkg:232-296 / 453-509: candidates
//                               consumed in draw order, constants and NaN scores do not count toward k, strict `>`
//                               keeps the first best), the child filters (pkg:1024-1039) and child creation.
//   k_lane<TASK, VT, SMALL>     one warp per node of up to 512 rows, one LANE per candidate.
#pragma once
#include "build.cuh"

namespace etb {

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
                                             : (TEAM == MID_TEAM ? (CODED ? MID_CODED_CTAS : 6)
                                                                 : ((TASK == TASK_REG && TEAM == CTA_TEAM) ? 1 : ((CODED && TEAM == CBIG_TEAM) ? CBIG_CTAS : 2))))
    k_node(P p, int32_t qcount, int qi, int lane_mode = 0) {
  // lane_mode: the queue is one of k_lane's size classes, handed to CTA teams because the level holds too few such
  // nodes to fill the GPU with single warps (launch_level): the node histogram is not in the frontier (k_lane
  // parents do not write it) and the batch sizes follow k_lane's rule, so that the draws -- and the tree -- do not
  // depend on which kernel ran.
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
    if (lane_mode) {
      for (int c = tid; c < C; c += TEAM) s_hnode[c] = 0;
      team_sync<TEAM>();
      for (int32_t j = tid; j < n; j += TEAM) atomicAdd(&s_hnode[LAB(j)], 1);
    } else {
      for (int c = tid; c < C; c += TEAM) s_hnode[c] = p.cur.hist[(int64_t)i * C + c];
    }
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
      if (!CODED && p.csr_ptr) sparse_mark_constant<TEAM>(p, rr, n, s_const, s_taken, W, tid);
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
        if (lane_mode) {  // k_lane's rule, verbatim
          if (st_draws > 0)
            extra = (st_draws > (unsigned long long)visited)
                        ? (int32_t)(((long long)need * (long long)(st_draws - (unsigned long long)visited)) / max(visited, 1)) + 2
                        : 0;
          else
            extra = (nconst > 0) ? need + 4 : 0;
        } else if (need > 0 && st_draws > 0) {
          extra = (int32_t)(((long long)need * (long long)(st_draws - (unsigned long long)visited)) / max(visited, 1));
        }
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
          Col colp[4];
#pragma unroll
          for (int c = 0; c < 4; c++) {
            const int32_t f = (g0 + c < nb) ? s_feat[g0 + c] : -1;
            colp[c] = col_of(p, f >= 0 ? f : 0);
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
                for (int c = 0; c < 4; c++) x[u2][c] = (r[u2] >= 0) ? col_at(colp[c], r[u2]) : 0.0;
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
                for (int c = 0; c < 4; c++) x[u2][c] = (r[u2] >= 0) ? col_at(colp[c], r[u2]) : 0.0;
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
        const Col col = col_of(p, f);
        double mn = 1.7976931348623157e308, mx = -1.7976931348623157e308;  // pkg:35-36
        int has_nan = 0;
        if (WARP) {
          // one gather per sample: the values are parked in shared memory for the second pass
          for (int v0 = 0; v0 < nv; v0 += 4) {
            int32_t r4[4];
            double x4[4];
#pragma unroll
            for (int u2 = 0; u2 < 4; u2++) {
              const int j = (v0 + u2) * 32 + lane;
              r4[u2] = (j < n) ? s_rows[j] : -1;
            }
            col_at4(col, r4, x4);
#pragma unroll
            for (int u2 = 0; u2 < 4; u2++) {
              if (r4[u2] >= 0) {
                const double x = x4[u2];
                s_xs[(v0 + u2) * 32 + lane] = x;
                if (x < mn) mn = x;
                if (x > mx) mx = x;
                has_nan |= (x != x);
              }
            }
          }
        } else {
          auto sweep = [&](auto uu) {  // U rows per thread and trip
            constexpr int U = decltype(uu)::value;
            for (int32_t j0 = 0; j0 < n; j0 += U * TEAM) {
              int32_t r4[U];
              double x4[U];
#pragma unroll
              for (int u2 = 0; u2 < U; u2++) {
                const int32_t j = j0 + u2 * TEAM + tid;
                r4[u2] = (j < n) ? rr[j] : -1;
              }
              col_atn<U>(col, r4, x4);
#pragma unroll
              for (int u2 = 0; u2 < U; u2++) {
                if (r4[u2] >= 0) {
                  const double x = x4[u2];
                  if (staged) s_xs[j0 + u2 * TEAM + tid] = x;
                  if (x < mn) mn = x;
                  if (x > mx) mx = x;
                  has_nan |= (x != x);
                }
              }
            }
          };
          if (col.rows)
            sweep(std::integral_constant<int, 8>{});  // sparse table: eight searches in lockstep
          else
            sweep(std::integral_constant<int, 4>{});
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
                x4[u2] = (r4[u2] >= 0) ? (staged ? s_xs[r4[u2]] : col_at(col, r4[u2])) : cut;
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
                const double x = staged ? s_xs[j] : col_at(col, rr[j]);
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
              const double x = staged ? s_xs[j] : col_at(col, rr[j]);
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
              const double x = (WARP || staged) ? s_xs[j] : col_at(col, rr[j]);
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
      if (qc == Q_WIDE) atomicAdd(&p.cnt->wide_rows, (unsigned long long)cn);
      if (qc >= Q_CTA) atomicAdd(&p.cnt->big_rows, (unsigned long long)cn);
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
    const Col col = CODED ? Col{nullptr, nullptr, 0} : col_of(p, best_feature);
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
          const double x = col_at(col, r);
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

// lnb = candidates per batch whose values are parked (32 on byte-coded tables; on FP64 tables k plus some slack,
// so that the 8-byte values of a 32-row node take k-ish instead of 32 slots per row and more warps fit an SM)
__host__ __device__ inline int lane_smem_bytes(int task, int C, int W, bool replay, int NW, int vbytes, int lnb = 32) {
  int o = 0;
  o += (task == TASK_CLS) ? 0 : 32 * NW * 8;      // s_y    regression target / weight, by position
  o += (task == TASK_REG) ? 0 : C * 8;            // s_dist
  o += 32 * NW * lnb * vbytes;                    // s_x    parked values [position][lane < lnb]
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
static __device__ __noinline__ double gini_score_bits(const uint32_t *s_lt, const uint32_t *s_nn, bool nan_left, int lane,
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
static __device__ __noinline__ double gini_score_w_bits(const uint32_t *s_lt, const uint32_t *s_nn, bool nan_left, int lane,
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
static __device__ __noinline__ double var_reduction_bits(const uint32_t *s_lt, const uint32_t *s_nn, bool nan_left, int lane,
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
  const int LNB = CODED ? 32 : p.lane_nb;
  unsigned char *sm = smem_raw + (size_t)tic * lane_smem_bytes(TASK, C, W, p.replay != 0, NW, (int)sizeof(VT), LNB);
  double *s_y = reinterpret_cast<double *>(sm);
  double *s_dist = s_y + ((TASK == TASK_CLS) ? 0 : 32 * NW);
  VT *s_x = reinterpret_cast<VT *>(s_dist + ((TASK == TASK_REG) ? 0 : C));
  uint32_t *s_lt = reinterpret_cast<uint32_t *>(s_x + 32 * NW * LNB);
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
      if (!CODED && p.csr_ptr) {  // sparse-resident table: features none of the node's rows stores are constant
        sparse_mark_constant<32>(p, idx, n, s_const, s_taken, W, lane);
        nc = 0;
        for (int w = lane; w < W; w += 32) nc += __popc(s_const[w]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nc += __shfl_xor_sync(FULL, nc, o);
        nconst = nc - (W * 32 - p.d);
      }
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
      // 16-byte loads: lane q of a trip owns features 16 q .. 16 q + 15 of the row (two trips cover 1024 features)
      const int nchunk = p.r8_stride >> 4;
      const uint4 *cof16 = reinterpret_cast<const uint4 *>(p.coff);
      uint4 first[2], acc[2];
      {
        const uint4 *rp = reinterpret_cast<const uint4 *>(p.R8 + (int64_t)idx[0] * p.r8_stride);
#pragma unroll
        for (int it = 0; it < 2; it++) {
          const int q16 = it * 32 + lane;
          first[it] = (q16 < nchunk) ? __ldg(rp + q16) : make_uint4(0u, 0u, 0u, 0u);
          acc[it] = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      for (int w = 0; w < nw; w++) {
        const int j0 = w << 5, cnt = min(32, n - j0);
        const int32_t row = (lane < cnt) ? idx[j0 + lane] : 0;
        for (int jj = (w == 0) ? 1 : 0; jj < cnt; jj++) {
          const uint4 *rp = reinterpret_cast<const uint4 *>(p.R8 + (int64_t)__shfl_sync(FULL, row, jj) * p.r8_stride);
#pragma unroll
          for (int it = 0; it < 2; it++) {
            const int q16 = it * 32 + lane;
            if (q16 < nchunk) {
              const uint4 v = __ldg(rp + q16);
              acc[it].x |= v.x ^ first[it].x;
              acc[it].y |= v.y ^ first[it].y;
              acc[it].z |= v.z ^ first[it].z;
              acc[it].w |= v.w ^ first[it].w;
            }
          }
        }
      }
#pragma unroll
      for (int it = 0; it < 2; it++) {
        const int q16 = it * 32 + lane;
        uint32_t bits = 0u;  // 16 bits, feature order: bit b = feature 16 q + b varies over the node's rows
        if (q16 < nchunk) {
          const uint4 k4 = __ldg(cof16 + q16);  // byte 0 = the column holds NaNs
          auto vary4 = [](uint32_t a, uint32_t f, uint32_t k) -> uint32_t {
            const uint32_t nan4 = __vcmpeq4(f, 0u) & __vcmpeq4(k, 0u);  // the first row is NaN there
            const uint32_t ne4 = __vcmpne4(a, 0u) | nan4;               // 0xff per varying feature
            return ((ne4 & 0x01010101u) * 0x01020408u) >> 24;           // 4 bits, feature order
          };
          bits = vary4(acc[it].x, first[it].x, k4.x) | (vary4(acc[it].y, first[it].y, k4.y) << 4) |
                 (vary4(acc[it].z, first[it].z, k4.z) << 8) | (vary4(acc[it].w, first[it].w, k4.w) << 12);
        }
        uint32_t word = bits << (16 * (lane & 1));
        word |= __shfl_xor_sync(FULL, word, 1);
        const int w0 = q16 >> 1;
        // taken = not varying (padding included), or excluded on the path from the root: a feature that scored NaN
        // at an ancestor stays out of the whole subtree (pkg:283-285) even if it varies here
        if ((lane & 1) == 0 && w0 < W) s_taken[w0] = ~word | s_const[w0];
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
        nb = min(LNB, tcnt - tpos);
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
        nb = (need > 0) ? min(LNB, min(avail, need + extra)) : 0;
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
                            : nullptr;
      const Col dcol = CODED ? Col{nullptr, nullptr, 0} : col_of(p, act0 ? f : 0);
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
          const double x = act0 ? col_at(dcol, rj) : 0.0;
          if (lane < LNB) s_x[pos * LNB + lane] = (VT)x;
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
      bool sparse_done = false;
      if constexpr (!CODED) {
       if (p.csc_row) {
        // sparse table: four rows of the lane's column in lockstep (col_at4)
        sparse_done = true;
        for (int w = 0; w < nw; w++) {
          const int j0 = w << 5, cnt = min(32, n - j0);
          const int32_t row = (lane < cnt) ? idx[j0 + lane] : 0;
          for (int jj = 0; jj < cnt; jj += 4) {
            int32_t r4[4];
            double x4[4];
#pragma unroll
            for (int u2 = 0; u2 < 4; u2++) {
              const int32_t rj = __shfl_sync(FULL, row, (jj + u2) & 31);
              r4[u2] = (act0 && jj + u2 < cnt) ? rj : -1;
            }
            col_at4(dcol, r4, x4);
#pragma unroll
            for (int u2 = 0; u2 < 4; u2++) {
              if (jj + u2 < cnt) {
                const double x = x4[u2];
                if (lane < LNB) s_x[(j0 + jj + u2) * LNB + lane] = (VT)x;
                if (x < mn) mn = x;
                if (x > mx) mx = x;
                has_nan |= (x != x);
              }
            }
          }
        }
       }
      }
      if (sparse_done) {
      } else if (CODED && p.c8_small)
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
            const double x = (lane < LNB) ? (double)s_x[(j0 + jj) * LNB + lane] : 0.0;
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

}  // namespace etb
