// node_inst.cu -- instantiates the node kernels (node.cuh) for ONE task and launches a level of them.
// Compiled three times (-DET_TASK=0 unweighted classification, 1 weighted classification, 2 regression): the
// kernels are large, so each task is its own object and the three compile side by side.
#include "node.cuh"

#ifndef ET_TASK
#error "compile with -DET_TASK=0|1|2"
#endif

namespace etb {

namespace {

int lane_small_nw() {
  static const int v = getenv("ETGPU_LANE_SMALL_NW") ? std::max(1, atoi(getenv("ETGPU_LANE_SMALL_NW"))) : 2;
  return v;
}

template <int TASK, typename VT>
void launch_lane(et_ctx *ctx, const P &p, int32_t count, int qi, int NW, size_t smem_per_warp, cudaStream_t st) {
  const unsigned grid = (unsigned)ceil_div(count, LANE_WARPS);
  const int small_nw = lane_small_nw();
  // classes of up to 32 * small_nw rows run the register-lean variant (64 registers, 8 CTAs per SM); larger ones the
  // variant that keeps 32 gathers of a lane in flight (128 registers, 4 CTAs per SM).  ETGPU_LANE_SMALL_NW moves it.
  if (NW <= small_nw)
    k_lane<TASK, VT, true><<<grid, 32 * LANE_WARPS, smem_per_warp * LANE_WARPS, st>>>(p, count, qi, NW);
  else
    k_lane<TASK, VT, false><<<grid, 32 * LANE_WARPS, smem_per_warp * LANE_WARPS, st>>>(p, count, qi, NW);
  ctx->launches++;
}

// the byte-coded CTA teams exist for unweighted classification only
template <int TASK, int TEAM>
void launch_coded_team(const P &p, int32_t count, int qi, size_t smem, cudaStream_t st, int lane_mode = 0) {
  if constexpr (TASK == TASK_CLS)
    k_node<TASK_CLS, TEAM, true><<<(unsigned)count, TEAM, smem, st>>>(p, count, qi, lane_mode);
}

template <int TASK>
void set_coded_team_attr(const LevelCfg &lc) {
  if constexpr (TASK == TASK_CLS) {
    CUDA_CHECK(cudaFuncSetAttribute(k_node<TASK_CLS, MID_TEAM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)lc.smem_mid));
    CUDA_CHECK(cudaFuncSetAttribute(k_node<TASK_CLS, CBIG_TEAM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)lc.smem_cta));
  }
}

}  // namespace

// One level = one launch per non-empty size class.  The classes are independent (disjoint nodes), so each runs
// on its own stream: the few long-running CTAs of the large nodes overlap with the many small teams instead of
// serialising behind them.  The chunked path of the wide nodes is a sequence of kernels with host round trips
// (candidate rounds); it runs on the main stream after the other classes have been queued on theirs.
// (ETGPU_TIMING serialises the classes to time each.)
template <>
void launch_level<ET_TASK>(et_ctx *ctx, const P &p, const int32_t *qn, int64_t wide_rows, const LevelCfg &lc,
                           PhaseTimer &pt, EventTimer &et, WideBufs *wb) {
  constexpr int TASK = ET_TASK;
  cudaStream_t main_st = ctx->stream;
  const bool fork = !pt.on;
  const bool has_wide = qn[Q_WIDE] > 0;
  int used = 0;
  for (int q = 0; q < NQ; q++) used += (qn[q] > 0);
  const int e0 = et.rec(main_st);
  if (fork && used > 1) cudaEventRecord(ctx->ev_fork, main_st);
  bool joined[NQ] = {false};
  int first = has_wide ? 0 : 1;  // the main stream belongs to the wide nodes when there are any
  auto stream_for = [&](int side) -> cudaStream_t {
    if (!fork || used <= 1 || first) {  // else the first (largest) class stays on the main stream
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
    pt.stop(PhaseTimer::CTA);
  }
  if (qn[Q_MID] > 0) {
    pt.start();
    cudaStream_t st = stream_for(Q_MID);
    if (lc.coded_big)
      launch_coded_team<TASK, MID_TEAM>(p, qn[Q_MID], Q_MID, lc.smem_mid, st);
    else
      k_node<TASK, MID_TEAM, false><<<(unsigned)qn[Q_MID], MID_TEAM, lc.smem_mid, st>>>(p, qn[Q_MID], Q_MID);
    ctx->launches++;
    pt.stop(PhaseTimer::MID);
  }
  // A warp walks its node's rows one after the other: a level that holds only a few nodes of a large lane class
  // lasts as long as that chain (0.3 ms for 512 rows) while the GPU idles -- the tail levels of every build, and
  // most levels when the forest is sharded over 8 GPUs.  Such a class goes to the 128-thread CTA teams instead
  // (a quarter of the chain; same draws and same tree, see k_node's lane_mode).  ETGPU_TEAM_MAX moves the bound
  // (nodes per class and level; 0 = never).
  static const int small_nw = lane_small_nw();
  // (sparse-resident tables: always -- a node's time there is its chain of dependent searches, whatever the level
  // holds; dense FP64 tables: off by default -- no gain was measured there, and the two FP64 bench workloads read
  // 6 % lower in the one run that had it on)
  int team_max = p.csc_row ? 0x7fffffff : (lc.coded ? 2 * 148 : 0);
  if (const char *env = getenv("ETGPU_TEAM_MAX")) team_max = std::max(0, atoi(env));
  for (int q = Q_WARP; q >= 0; q--) {
    if (qn[q] <= 0) continue;
    pt.start();
    cudaStream_t st = stream_for(q);
    const bool to_team = q >= 2 && (1 << q) > small_nw && qn[q] <= team_max && p.NB == 32;
    if (lc.coded && to_team) {
      if (lc.coded_big)
        launch_coded_team<TASK, MID_TEAM>(p, qn[q], q, lc.smem_mid, st, 1);
      else
        k_node<TASK, MID_TEAM, false><<<(unsigned)qn[q], MID_TEAM, lc.smem_mid, st>>>(p, qn[q], q, 1);
      ctx->launches++;
    } else if (lc.coded) {
      launch_lane<TASK, uint8_t>(ctx, p, qn[q], q, 1 << q, lc.smem_lane[q], st);
    } else if (q == Q_LANE0) {
      launch_lane<TASK, double>(ctx, p, qn[q], q, 1, lc.smem_lane[0], st);
    } else if (qn[q] <= team_max && p.NB == 32) {  // FP64 tables, few nodes: CTA teams (same kernel, wider team)
      k_node<TASK, MID_TEAM, false><<<(unsigned)qn[q], MID_TEAM, lc.smem_mid, st>>>(p, qn[q], q, 0);
      ctx->launches++;
    } else {  // FP64 tables: only class Q_WARP is populated besides class Q_LANE0
      k_node<TASK, 32, false>
          <<<(unsigned)ceil_div(qn[q], WARPS_PER_CTA), 32 * WARPS_PER_CTA, lc.smem_warp * WARPS_PER_CTA, st>>>(p, qn[q], q);
      ctx->launches++;
    }
    pt.stop(PhaseTimer::LANE);
  }
  if (has_wide) {
    pt.start();
    if constexpr (TASK != TASK_CLSW) wide_level<TASK>(ctx, p, qn[Q_WIDE], wide_rows, lc, *wb, main_st);
    pt.stop(PhaseTimer::WIDE);
  }
  for (int i = 0; i < NQ; i++) {
    if (joined[i]) {
      cudaEventRecord(ctx->ev_join[i], ctx->side[i]);
      cudaStreamWaitEvent(main_st, ctx->ev_join[i], 0);
    }
  }
  const int e1 = et.rec(main_st);
  et.spans[0].push_back({e0, e1});
  if (qn[Q_MID] + qn[Q_CTA] + qn[Q_WIDE] > 0) et.spans[1].push_back({e0, e1});
}

template <>
void set_smem_attr<ET_TASK>(const LevelCfg &lc) {
  constexpr int TASK = ET_TASK;
  if (lc.coded_big) {
    set_coded_team_attr<TASK>(lc);
  } else {
    CUDA_CHECK(cudaFuncSetAttribute(k_node<TASK, MID_TEAM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)lc.smem_mid));
    CUDA_CHECK(cudaFuncSetAttribute(k_node<TASK, CTA_TEAM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)lc.smem_cta));
  }
  if (lc.coded) {
    CUDA_CHECK(cudaFuncSetAttribute(k_lane<TASK, uint8_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(lc.smem_lane[4] * LANE_WARPS)));
    CUDA_CHECK(cudaFuncSetAttribute(k_lane<TASK, uint8_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(lc.smem_lane[4] * LANE_WARPS)));
  } else {
    CUDA_CHECK(cudaFuncSetAttribute(k_lane<TASK, double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(lc.smem_lane[0] * LANE_WARPS)));
    CUDA_CHECK(cudaFuncSetAttribute(k_node<TASK, 32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(lc.smem_warp * WARPS_PER_CTA)));
  }
}

}  // namespace etb
