// build.cu -- host orchestration of the level-wise extratrees builder (structures and layout: build.cuh).
//
// One level = one launch per size class (node_inst.cu) plus the chunked path of the wide nodes (wide.cu).
// The host only reads the next level's queue sizes.  At the end of a batch of trees the pool of nodes
// (creation order) is converted on the device into per-tree pre-order 16-byte nodes.
#include <nvtx3/nvToolsExt.h>

#include "node.cuh"

namespace etb {

#ifdef ETGPU_NO_NVTX
NvtxRange::NvtxRange(const char *) {}
NvtxRange::~NvtxRange() {}
#else
NvtxRange::NvtxRange(const char *name) { nvtxRangePushA(name); }
NvtxRange::~NvtxRange() { nvtxRangePop(); }
#endif

namespace {

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


// ---- pool (creation order) -> per-tree pre-order ----------------------------------------------
__global__ void k_subtree_sizes(Pool o, int32_t lo, int32_t hi, int32_t *size, int32_t *nleaf) {
  int v = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= hi) return;
  if (o.feat[v] < 0) {
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
                          PNode *nodes, double *leaves, int32_t *max_feat) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  const int32_t fv = (v < n_nodes && o.feat[v] >= 0) ? (o.feat[v] & (ET_MIL_BIT - 1)) + 1 : 0;
  const int32_t fm = __reduce_max_sync(0xffffffffu, fv);
  if ((threadIdx.x & 31) == 0 && fm > 0) atomicMax(max_feat, fm);
  if (v >= n_nodes) return;
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


// temporaries of one build call: freed on every exit path
struct BuildTmp {
  et_ctx *ctx = nullptr;
  std::vector<std::pair<void *, size_t>> dev;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  BestBufs *best = nullptr;
  // (blocks come from the context's cache: a small build is not a string of cudaMalloc / cudaFree pairs)
  template <typename T>
  T *upload(const T *h, size_t n, cudaStream_t st) {
    const size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
    T *d = static_cast<T *>(et_dev_alloc(ctx, bytes));
    if (!d) ET_FAIL(ET_ENOMEM, "device allocation failed");
    dev.push_back({d, bytes});
    if (n) CUDA_CHECK(cudaMemcpyAsync(d, h, n * sizeof(T), cudaMemcpyHostToDevice, st));
    return d;
  }
  ~BuildTmp() {
    for (auto &d : dev) et_dev_free(ctx, d.first, d.second);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (best) best_bufs_destroy(best);
  }
};

}  // namespace
}  // namespace etb

using namespace etb;

// Device buffers that survive across builds on one context (no cudaMalloc in the steady state).
struct Workspace {
  DevBuf<int32_t> idx[2], yc[2], q[2][NQ], size, nleaf, pos, lpos;
  DevBuf<double> yr[2], ws[2];
  FrontierBufs fr[2];
  PoolBufs pool;
  DevBuf<uint32_t> scratch;
  DevBuf<Counters> cnt;
  DevBuf<int64_t> tree_off, leaf_off;
  WideBufs *wide = nullptr;
  Counters *h_cnt = nullptr;  // pinned: the per-level readback of the counters is a plain DMA
  ~Workspace() {
    if (wide) wide_bufs_destroy(wide);
    if (h_cnt) cudaFreeHost(h_cnt);
  }
};

void et_workspace_free(Workspace *ws) { delete ws; }

void et_build_forest(et_ctx *ctx, et_data *D, const BuildArgs &a, et_forest *out, et_stats *stats) {
  NvtxRange nv_build("etgpu.build_forest");
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
  if (!ws.wide) ws.wide = wide_bufs_create();
  if (!ws.h_cnt) CUDA_CHECK(cudaHostAlloc((void **)&ws.h_cnt, sizeof(Counters), cudaHostAllocDefault));
  et_stats S;
  memset(&S, 0, sizeof(S));
  const int64_t launches0 = ctx->launches;
  PhaseTimer pt;
  pt.st = st;
  EventTimer evt;
  double tacc[2] = {0, 0};
  BuildTmp tmp;
  tmp.ctx = ctx;

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
  // FP64 tables: k_lane parks k + slack values per row instead of 32 (more warps per SM; batches are capped to it)
  const int lane_nb = std::min(32, (a.k + std::max(2, a.k / 4) + 3) / 4 * 4);
  if (!lc.coded) lc.smem_lane[0] = (size_t)lane_smem_bytes(task, C, W, replay, 1, 8, lane_nb);
  if (lc.smem_lane[0] * LANE_WARPS > 200 * 1024)
    ET_FAIL(ET_EUNSUPPORTED, "numClasses=%d / %d features need more shared memory per node than one SM has", C, d);
  // Nodes above wide_min rows are cut into chunks of rows, one CTA per chunk (wide.cu): unweighted classification
  // with <= 32 classes and regression; weighted classification keeps one CTA per node (sequential weight sums).
  // Default: a node is chunked when it holds more than its share of the level's large-node rows over twice the
  // CTA slots of the GPU -- with hundreds of trees in flight one CTA per node already fills the machine (and has
  // less fixed cost per level), with few trees or very large tables the top of every tree is chunked.
  // ETGPU_WIDE_MIN=n fixes the bound (>= 2^31-1: chunked path off).
  int64_t wide_min_fixed = -1;
  if (const char *env = getenv("ETGPU_WIDE_MIN")) wide_min_fixed = std::max<int64_t>(NM_MAX, atoll(env));
  lc.wide = !a.best_split && NB == 32 && (task == TASK_REG || (task == TASK_CLS && C <= 32)) && wide_min_fixed < 0x7fffffff;
  const int64_t cta_slots = (int64_t)ctx->sm_count * ((task == TASK_REG) ? 1 : 2);
  auto wide_min_for = [&](int64_t big_rows) -> int32_t {
    if (!lc.wide) return 0x7fffffff;
    if (wide_min_fixed >= 0) return (int32_t)wide_min_fixed;
    // (and never below a size one CTA finishes faster than the chunked path's fixed cost of ~0.3 ms per level).
    // Regression: the one-CTA team of a large node holds an SM alone (512 threads, moment sums), the chunked path
    // was measured faster at every level of the 1M-row table, so it takes every node above NM_MAX rows.
    if (task == TASK_REG) return NM_MAX;
    if (D->csc_row) return NM_MAX;  // sparse-resident table: the chunked path walks stored entries instead of rows
    return (int32_t)std::min<int64_t>(0x7fffffff, std::max<int64_t>(16384, big_rows / (4 * cta_slots)));
  };
  if (task == TASK_CLS)
    set_smem_attr<TASK_CLS>(lc);
  else if (task == TASK_CLSW)
    set_smem_attr<TASK_CLSW>(lc);
  else
    set_smem_attr<TASK_REG>(lc);

  CUDA_CHECK(cudaEventCreate(&tmp.ev0));
  CUDA_CHECK(cudaEventCreate(&tmp.ev1));
  if (a.best_split) tmp.best = best_bufs_create();
  CUDA_CHECK(cudaEventRecord(tmp.ev0, st));

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
  if (!replay) {  // ... and the known-constant feature masks of the open nodes (W words per node, two levels) to ~16 GB
    const int64_t mask_per_tree = std::max<int64_t>(1, n / 4) * (int64_t)W * 8;
    B = (int32_t)std::max<int64_t>(1, std::min<int64_t>(B, ((int64_t)16 << 30) / mask_per_tree));
  }
  if (const char *env = getenv("ETGPU_BATCH_TREES")) B = std::max(1, std::min(a.m, atoi(env)));

  ws.cnt.ensure(1);
  std::vector<int32_t> rh((size_t)std::max(C, 1), 0);
  if (task == TASK_CLS)
    for (int c = 0; c < C; c++) rh[(size_t)c] = (int32_t)D->root_hist[(size_t)c];
  int32_t *d_root_hist = tmp.upload(rh.data(), rh.size(), st);

  // replay trace on the device (child ids rewritten to absolute node indices)
  Trace trace{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
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
    trace.cand_begin = tmp.upload(R->cand_begin, (size_t)nn, st);
    trace.cand_count = tmp.upload(R->cand_count, (size_t)nn, st);
    trace.left = tmp.upload(l.data(), (size_t)nn, st);
    trace.right = tmp.upload(r.data(), (size_t)nn, st);
    trace.cand_feature = tmp.upload(R->cand_feature, (size_t)R->n_cand, st);
    trace.cand_u = tmp.upload(R->cand_u, (size_t)R->n_cand, st);
    trace.cand_flag = tmp.upload(R->cand_flag, (size_t)R->n_cand, st);
    CUDA_CHECK(cudaStreamSynchronize(st));  // (l, r are host temporaries)
  }
  struct Seg {
    PNode *nodes;
    double *leaves;
    int64_t n_nodes, n_leaves;
  };
  std::vector<Seg> segs;
  auto free_seg = [&](Seg &sg) {
    et_dev_free(ctx, sg.nodes, (size_t)sg.n_nodes * sizeof(PNode));
    et_dev_free(ctx, sg.leaves, std::max<size_t>(1, (size_t)sg.n_leaves * lw) * sizeof(double));
  };

  out->m = a.m;
  out->d_min = 0;
  out->order_key.resize((size_t)a.m);
  for (int t = 0; t < a.m; t++)
    out->order_key[(size_t)t] = a.order_keys ? a.order_keys[t] : (a.tree_ids ? (int64_t)a.tree_ids[t] : (int64_t)t);
  out->tree_off.assign((size_t)a.m + 1, 0);
  int64_t node_base = 0, leaf_base = 0;  // forest-wide offsets of the current batch

  try {
    for (int32_t t0 = 0; t0 < a.m; t0 += B) {
      const int32_t Bt = std::min(B, a.m - t0);
      const size_t ns = (size_t)Bt * (size_t)n;
      pt.start();
      for (int q = 0; q < 2; q++) {
        ws.idx[q].ensure(ns + 8, 1.0);  // (+ slack: bulk copies of a segment are rounded up to 16 bytes)
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
      p.cls_max[Q_LANE0] = NT_MAX;
      for (int q = 1; q < 4; q++) p.cls_max[Q_LANE0 + q] = lc.coded ? (NT_MAX << q) : NT_MAX;
      p.cls_max[Q_WARP] = NW_MAX;
      p.cls_max[Q_MID] = NM_MAX;
      p.cls_max[Q_CTA] = wide_min_for(n > NM_MAX ? (int64_t)Bt * n : 0);
      // Sparse-resident table: every node above sparse_wide_min rows takes the chunked path, where a chunk walks the
      // stored entries of a candidate's column (wide.cu) -- the one-team kernels would search the column once per
      // (row, candidate) and a single 2048-row node held a level for milliseconds.
      int32_t sparse_wide_min = 0;
      if (D->csc_row && lc.wide) {
        sparse_wide_min = NM_MAX;
        if (const char *env = getenv("ETGPU_SPARSE_WIDE_MIN")) sparse_wide_min = std::max(NT_MAX, atoi(env));
        for (int q = Q_LANE0 + 1; q <= Q_CTA; q++) p.cls_max[q] = std::min(p.cls_max[q], sparse_wide_min);
      }
      if (a.best_split) {  // bestSplit: one kernel family (best.cu), every node in queue Q_CTA
        for (int q = 0; q < Q_CTA; q++) p.cls_max[q] = 0;
        p.cls_max[Q_CTA] = 0x7fffffff;
      }
      p.R8 = D->r8;
      p.r8_stride = (int32_t)D->rs8;
      p.nc_max = 64;
      if (const char *env = getenv("ETGPU_NC_MAX")) p.nc_max = std::max(0, atoi(env));
      p.lane_nb = lane_nb;
      p.csc_colptr = D->csc_colptr;
      p.csc_row = D->csc_row;
      p.csc_val = D->csc_val;
      p.csr_ptr = D->csr_ptr;
      p.csr_col = D->csr_col;
      p.inv_rows = (int64_t)Bt * n;
      p.tr = trace;
      int srcb = 0, cl = 0;  // cl: frontier slot of the current level
      int32_t F = Bt;
      ws.fr[0].ensure((size_t)F, C, W, task == TASK_CLS, !replay);
      for (int q = 0; q < NQ; q++) ws.q[0][q].ensure((size_t)F, 1.5);
      pt.stop(PhaseTimer::ALLOC);
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
      uint64_t *d_keys = tmp.upload(tree_keys.data() + t0, (size_t)Bt, st);
      int64_t *d_troots = replay ? tmp.upload(trace_roots.data() + t0, (size_t)Bt, st) : nullptr;
      CUDA_CHECK(cudaMemsetAsync(ws.cnt.p, 0, sizeof(Counters), st));
      k_init_roots<<<(unsigned)ceil_div(F, 128), 128, 0, st>>>(p, Bt, d_keys, d_troots, d_root_hist);
      ctx->launches++;
      pt.stop(PhaseTimer::INIT);

      int64_t n_nodes = Bt, n_leaves = 0;
      std::vector<int32_t> level_start{0};
      int32_t qn[NQ] = {0};
      qn[size_class(p, n)] = Bt;
      int64_t wide_rows = (size_class(p, n) == Q_WIDE) ? (int64_t)Bt * n : 0;
      int64_t big_rows = (n > NM_MAX) ? (int64_t)Bt * n : 0;
      int64_t leaf_bound = 0;  // upper bound of the leaves allocated so far
      Counters &hc = *ws.h_cnt;
      memset(&hc, 0, sizeof(hc));
      while (F > 0) {
        NvtxRange nv_level("etgpu.level");
        S.levels++;
        pt.start();
        const int cn = cl ^ 1;
        leaf_bound += (int64_t)F;
        ws.fr[cn].ensure((size_t)F * 2, C, W, task == TASK_CLS, !replay);
        for (int q = 0; q < NQ; q++) ws.q[cn][q].ensure((size_t)F * 2, 1.5);
        ws.pool.grow((size_t)(n_nodes + 2 * (int64_t)F), (size_t)n_nodes, (size_t)leaf_bound, (size_t)n_leaves, lw, st);
        {  // side-bit scratch of the CTA teams (weighted / regression); lane classes may be handed to teams (launch_level)
          const size_t teams = (size_t)qn[2] + (size_t)qn[3] + (size_t)qn[Q_WARP] + (size_t)qn[Q_MID] + (size_t)qn[Q_CTA];
          if (task != TASK_CLS && teams > 0)
            ws.scratch.ensure((size_t)NB * 2 * ((size_t)Bt * (size_t)n / 32 + teams + 1) + 64, 1.0);
        }
        pt.stop(PhaseTimer::ALLOC);
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
        p.node_base_next = (int32_t)n_nodes;
        if (!a.best_split) p.cls_max[Q_CTA] = sparse_wide_min ? std::min(sparse_wide_min, wide_min_for(big_rows)) : wide_min_for(big_rows);  // the bound the children are classed by
        if (a.best_split) {
          const int e0 = evt.rec(st);
          if (task == TASK_CLS)
            launch_level_best<TASK_CLS>(ctx, p, qn[Q_CTA], *tmp.best, (int64_t)ns);
          else if (task == TASK_CLSW)
            launch_level_best<TASK_CLSW>(ctx, p, qn[Q_CTA], *tmp.best, (int64_t)ns);
          else
            launch_level_best<TASK_REG>(ctx, p, qn[Q_CTA], *tmp.best, (int64_t)ns);
          evt.spans[0].push_back({e0, evt.rec(st)});
        } else if (task == TASK_CLS)
          launch_level<TASK_CLS>(ctx, p, qn, wide_rows, lc, pt, evt, ws.wide);
        else if (task == TASK_CLSW)
          launch_level<TASK_CLSW>(ctx, p, qn, wide_rows, lc, pt, evt, ws.wide);
        else
          launch_level<TASK_REG>(ctx, p, qn, wide_rows, lc, pt, evt, ws.wide);
        pt.level_report((int)S.levels - 1, qn);
        pt.start();
        CUDA_CHECK(cudaMemcpyAsync(&hc, ws.cnt.p, sizeof(Counters), cudaMemcpyDeviceToHost, st));
        // the next level starts from clean per-level counters (leaf count and stats keep accumulating)
        CUDA_CHECK(cudaMemsetAsync(ws.cnt.p, 0, offsetof(Counters, n_leaves), st));
        CUDA_CHECK(cudaMemsetAsync(&ws.cnt.p->wide_rows, 0, 3 * sizeof(unsigned long long), st));
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
        pt.stop(PhaseTimer::SYNC);
        const int32_t nf = hc.next_f;
        n_leaves = hc.n_leaves;
        level_start.push_back((int32_t)n_nodes);
        n_nodes += nf;
        if (n_nodes > 0x7ffffff0) ET_FAIL(ET_EUNSUPPORTED, "batch exceeds 2^31 nodes; lower ETGPU_BATCH_TREES");
        for (int q = 0; q < NQ; q++) qn[q] = hc.q_count[q];
        wide_rows = (int64_t)hc.wide_rows;
        big_rows = (int64_t)hc.big_rows;
        if (nf > 0) srcb ^= 1;
        F = nf;
        cl = cn;
      }
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
      NvtxRange nv_pre("etgpu.preorder");
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
      std::vector<int64_t> toff((size_t)Bt + 1);
      CUDA_CHECK(cudaMemcpyAsync(toff.data(), ws.tree_off.p, ((size_t)Bt + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaStreamSynchronize(st));
      const int64_t total_nodes = toff[(size_t)Bt] - node_base;
      S.nodes += total_nodes;
      Seg sg;
      sg.n_nodes = total_nodes;
      sg.n_leaves = n_leaves;
      sg.nodes = static_cast<PNode *>(et_dev_alloc(ctx, (size_t)total_nodes * sizeof(PNode)));
      sg.leaves = static_cast<double *>(et_dev_alloc(ctx, std::max<size_t>(1, (size_t)n_leaves * lw) * sizeof(double)));
      if (!sg.nodes || !sg.leaves) {
        free_seg(sg);
        ET_FAIL(ET_ENOMEM, "cannot allocate the forest (%lld nodes)", (long long)total_nodes);
      }
      segs.push_back(sg);
      k_scatter<<<(unsigned)ceil_div(n_nodes, 256), 256, 0, st>>>(po, (int32_t)n_nodes, lw, ws.pos.p, ws.lpos.p,
                                                                  ws.tree_off.p, ws.leaf_off.p, node_base, leaf_base,
                                                                  sg.nodes, sg.leaves, &ws.cnt.p->max_feat);
      ctx->launches++;
      int32_t max_feat = 0;
      CUDA_CHECK(cudaMemcpyAsync(&max_feat, &ws.cnt.p->max_feat, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaStreamSynchronize(st));
      CUDA_CHECK(cudaGetLastError());
      out->d_min = std::max(out->d_min, max_feat);
      for (int32_t t = 0; t <= Bt; t++) out->tree_off[(size_t)(t0 + t)] = toff[(size_t)t];
      node_base += total_nodes;
      leaf_base += n_leaves;
      pt.stop(PhaseTimer::PREORDER);
    }
    // ---- the forest stays resident in HBM; batches are concatenated
    out->total_nodes = node_base;
    out->total_leaves = leaf_base;
    if (segs.size() == 1) {
      out->d_nodes = segs[0].nodes;
      out->d_leaf = segs[0].leaves;
      out->nodes_bytes = (size_t)segs[0].n_nodes * sizeof(PNode);
      out->leaf_bytes = std::max<size_t>(1, (size_t)segs[0].n_leaves * lw) * sizeof(double);
      segs.clear();
    } else {
      // (the batches' segments and the whole forest are alive together for a moment: when HBM is short -- 250 trees
      // of a 10M-row table are 33 GB -- the per-batch workspace is given back first; the next build grows it again)
      auto alloc_forest = [&](void **ptr, size_t bytes) {
        if (cudaMalloc(ptr, bytes) == cudaSuccess) return;
        cudaGetLastError();
        for (int q = 0; q < 2; q++) {
          ws.idx[q].release();
          ws.yc[q].release();
          ws.yr[q].release();
          ws.ws[q].release();
        }
        ws.pool.tree.release();
        ws.pool.feat.release();
        ws.pool.child.release();
        ws.pool.cut.release();
        ws.pool.leaf_vals.release();
        ws.size.release();
        ws.nleaf.release();
        ws.pos.release();
        ws.lpos.release();
        ws.scratch.release();
        for (auto &b : ctx->cache) cudaFree(b.p);
        ctx->cache.clear();
        ctx->cache_bytes = 0;
        CUDA_CHECK(cudaMalloc(ptr, bytes));
      };
      alloc_forest((void **)&out->d_nodes, std::max<size_t>(1, (size_t)node_base) * sizeof(PNode));
      alloc_forest((void **)&out->d_leaf, std::max<size_t>(1, (size_t)leaf_base * lw) * sizeof(double));
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
    out->tree_off_bytes = ((size_t)a.m + 1) * sizeof(int64_t);
    out->d_tree_off = static_cast<int64_t *>(et_dev_alloc(ctx, out->tree_off_bytes));
    if (!out->d_tree_off) ET_FAIL(ET_ENOMEM, "device allocation failed");
    CUDA_CHECK(cudaMemcpyAsync(out->d_tree_off, out->tree_off.data(), ((size_t)a.m + 1) * sizeof(int64_t),
                               cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaEventRecord(tmp.ev1, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, tmp.ev0, tmp.ev1);
    S.gpu_ms = ms;
    S.gpu_ms_split = tacc[0];      // node kernels (split search + partition fused), all size classes of a level overlapped
    S.gpu_ms_partition = tacc[1];  // ... of which levels that still hold CTA-owned / chunked (large) nodes
    S.launches = ctx->launches - launches0;
    pt.report();
  } catch (...) {
    cudaStreamSynchronize(st);
    for (int i = 0; i < et_ctx::N_SIDE; i++) cudaStreamSynchronize(ctx->side[i]);
    for (auto &sg : segs) free_seg(sg);
    throw;
  }
  if (stats) *stats = S;
  if (replay && S.replay_mismatches && !stats)  // with stats the caller reads replay_mismatches itself
    ET_FAIL(ET_EREPLAY, "replay: %lld decisions contradict the trace (wrong data for this trace?)",
            (long long)S.replay_mismatches);
}
