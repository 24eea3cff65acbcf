// rnn_rec.cu — persistent recurrence kernels for sm_100a (K2/K3 forward, K4/K5 backward).
//
// Replaces the T-serial inner loop the reference reaches through torch.nn.GRU / torch.nn.LSTM
// (GRU cell equations torch/nn/modules/rnn.py:1221-1224, LSTM cell :842-847; call sites
// audio_gru_whole.py:105, text_bilstm_whole.py:105, fuse_net_whole.py:347,361).
//
// Design (one launch = every time step of every direction of one layer):
//   * grid = D * nslices thread-block clusters of C CTAs. A cluster owns BS batch rows; CTA `rank` of the
//     cluster owns HS = H/C hidden units and keeps the matching rows of W_hh (forward) or of W_hh^T
//     (backward) resident in shared memory for the whole sequence. The slice is staged once with TMA bulk
//     copies (cp.async.bulk ... mbarrier::complete_tx) straight from the parameter tensor.
//   * per step every warp contracts its rows against the BS state vectors (rnn_core.cuh: K across lanes,
//     transposing shuffle butterfly), the lane that ends up owning (unit, batch) applies the gate
//     non-linearities and the state update in registers, and the new state slice is all-gathered into the
//     C peer CTAs with st.shared::cluster (DSMEM), double buffered, one cluster barrier per step
//     (arrive.release early, wait.acquire after the global stores / next-step prefetch).
//   * batch slices are independent clusters: no grid-wide synchronisation anywhere.
//
// fp32 FFMA by choice: the per-step contraction is [BS x H] x [H x G*HS] with BS = 2..8 rows per CTA —
// far too skinny for tcgen05 tiles, and parity is judged at 1e-5 against an fp32 reference.
#include <mutex>

#include "profile.cuh"
#include "ptx.cuh"
#include "rnn_core.cuh"
#include "rnn_kernels.cuh"

namespace b200rnn {

namespace {

constexpr int MAX_SMEM = 232448;  // 227 KB opt-in limit per CTA on sm_100

template <int MODE, int H, int C, int BS, int UPW>
struct FwdCfg {
  static constexpr int G = (MODE == B200RNN_GRU) ? 3 : 4;
  static constexpr int HS = H / C;
  static constexpr int NW = HS / UPW;
  static constexpr int NT = NW * 32;
  static constexpr int NCOL = G * HS;
  static constexpr size_t W_BYTES = (size_t)NCOL * H * sizeof(float);
  static constexpr size_t V_BYTES = (size_t)2 * BS * H * sizeof(float);
  static constexpr size_t SMEM = W_BYTES + V_BYTES + 16;
  static_assert(HS * C == H && NW * UPW == HS, "bad split");
  static_assert(NT <= 1024 && SMEM <= MAX_SMEM, "config does not fit an SM");
};

template <int MODE, int H, int C, int BS, int UPW>
struct BwdCfg {
  static constexpr int G = (MODE == B200RNN_GRU) ? 3 : 4;
  static constexpr int GH = G * H;
  static constexpr int HS = H / C;
  static constexpr int NW = HS / UPW;
  static constexpr int NT = NW * 32;
  static constexpr size_t W_BYTES = (size_t)HS * GH * sizeof(float);
  static constexpr size_t V_BYTES = (size_t)2 * BS * GH * sizeof(float);
  static constexpr size_t SMEM = W_BYTES + V_BYTES + 16;
  static_assert(HS * C == H && NW * UPW == HS, "bad split");
  static_assert(NT <= 1024 && SMEM <= MAX_SMEM, "config does not fit an SM");
};

// =================================================================================================
// forward
// =================================================================================================
template <int MODE, int H, int C, int BS, int UPW>
__global__ void __launch_bounds__(FwdCfg<MODE, H, C, BS, UPW>::NT, 1)
    rec_fwd_kernel(const RecFwdParams p, const int nslices) {
  using Cfg = FwdCfg<MODE, H, C, BS, UPW>;
  using LM = LaneMap<UPW, BS>;
  constexpr int G = Cfg::G, HS = Cfg::HS, NT = Cfg::NT;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* W_s = reinterpret_cast<float*>(smem_raw);             // [G*HS][H]
  float* h_s = W_s + (size_t)Cfg::NCOL * H;                    // [2][BS][H]
  uint64_t* bar = reinterpret_cast<uint64_t*>(h_s + 2 * BS * H);

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint32_t rank = ptx::cluster_ctarank();
  const int cid = blockIdx.x / C;
  const int dir = cid / nslices;
  const int slice = cid - dir * nslices;
  const int b0 = slice * BS;
  const int j0 = (int)rank * HS;
  const int B = p.B, T = p.T;

  // ---- stage this CTA's rows of W_hh (G blocks of HS contiguous rows) with TMA bulk copies ----------
  if (tid == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    ptx::mbar_arrive_expect_tx(bar, (uint32_t)Cfg::W_BYTES);
    const float* w_hh = p.w_hh[dir];
#pragma unroll
    for (int g = 0; g < G; ++g)
      ptx::tma_bulk_g2s(W_s + (size_t)g * HS * H, w_hh + ((size_t)g * H + j0) * H,
                        (uint32_t)(HS * H * sizeof(float)), bar);
  }
  for (int i = tid; i < 2 * BS * H; i += NT) h_s[i] = 0.f;  // h_0 = 0 (rnn.py:1432-1440)
  ptx::mbar_wait(bar, 0);
  __syncthreads();
  ptx::cluster_sync_all();  // every CTA's state buffers are initialised before any peer writes into them

  // ---- lane identity: after the butterfly this lane owns (unit, batch) ------------------------------
  const int pu = LM::p(lane), qb = LM::q(lane), rep = LM::rep(lane);
  const int j = j0 + w * UPW + pu;  // hidden unit
  const int b = b0 + qb;            // batch row
  const bool valid = b < B;
  const int GH = G * H;
  float* gates = p.gates[dir];
  float* extra = p.extra[dir];
  const float bhn = (MODE == B200RNN_GRU) ? p.b_hh[dir][2 * H + j] : 0.f;

  float h_prev = 0.f, c_prev = 0.f;
  float gi[G];
#pragma unroll
  for (int g = 0; g < G; ++g) gi[g] = 0.f;
  if (valid && T > 0) {
    const int t0 = dir ? T - 1 : 0;
    const float* gp = gates + ((size_t)t0 * B + b) * GH + j;
#pragma unroll
    for (int g = 0; g < G; ++g) gi[g] = gp[g * H];
  }

  for (int step = 0; step < T; ++step) {
    const int t = dir ? (T - 1 - step) : step;
    const float* h_cur = h_s + (step & 1) * BS * H;
    float* h_nxt = h_s + ((step & 1) ^ 1) * BS * H;

    float acc[G][UPW][BS];
    warp_partial_dots<G, UPW, BS, H>(W_s, HS, w * UPW, h_cur, lane, acc);
    warp_transpose_reduce<G, UPW, BS>(acc);

    float hnew, s0, s1, s2, s3 = 0.f, sx;
    if (MODE == B200RNN_GRU) {
      const float r = sigmoid_f(gi[0] + acc[0][0][0]);
      const float z = sigmoid_f(gi[1] + acc[1][0][0]);
      const float hn = acc[2][0][0] + bhn;
      const float n = tanhf(gi[2] + r * hn);
      hnew = n + z * (h_prev - n);
      s0 = r; s1 = z; s2 = n; sx = hn;
    } else {
      const float ig = sigmoid_f(gi[0] + acc[0][0][0]);
      const float fg = sigmoid_f(gi[1] + acc[1][0][0]);
      const float gg = tanhf(gi[2] + acc[2][0][0]);
      const float og = sigmoid_f(gi[G - 1] + acc[G - 1][0][0]);
      const float cnew = fg * c_prev + ig * gg;
      hnew = og * tanhf(cnew);
      c_prev = cnew;
      s0 = ig; s1 = fg; s2 = gg; s3 = og; sx = cnew;
    }
    h_prev = hnew;

    // all-gather the new state into every CTA of the cluster (DSMEM), replicas split the peers
    {
      const uint32_t laddr = ptx::smem_u32(&h_nxt[qb * H + j]);
      for (int rk = rep; rk < C; rk += LM::NREP) ptx::st_cluster_f32(ptx::mapa(laddr, (uint32_t)rk), hnew);
    }
    __syncwarp();
    ptx::cluster_arrive_release();

    // off the critical path: global stores of this step, prefetch of the next step's x-projection
    if (valid && rep == 0) {
      p.y[(long long)t * p.y_st + (long long)b * p.y_sb + dir * H + j] = hnew;
      if (p.training) {
        float* gp = gates + ((size_t)t * B + b) * GH + j;
        gp[0] = s0;
        gp[H] = s1;
        gp[2 * H] = s2;
        if (G == 4) gp[3 * H] = s3;
        extra[((size_t)t * B + b) * H + j] = sx;
      }
      if (step == T - 1) {
        p.h_n[((size_t)dir * B + b) * H + j] = hnew;
        if (MODE == B200RNN_LSTM && p.c_n) p.c_n[((size_t)dir * B + b) * H + j] = c_prev;
      }
    }
    if (valid && step + 1 < T) {
      const int tn = dir ? (T - 2 - step) : (step + 1);
      const float* gp = gates + ((size_t)tn * B + b) * GH + j;
#pragma unroll
      for (int g = 0; g < G; ++g) gi[g] = gp[g * H];
    }
    __syncwarp();
    ptx::cluster_wait_acquire();
  }
}

// =================================================================================================
// backward (BPTT)
// =================================================================================================
template <int MODE, int H, int C, int BS, int UPW>
__global__ void __launch_bounds__(BwdCfg<MODE, H, C, BS, UPW>::NT, 1)
    rec_bwd_kernel(const RecBwdParams p, const int nslices) {
  using Cfg = BwdCfg<MODE, H, C, BS, UPW>;
  using LM = LaneMap<UPW, BS>;
  constexpr int G = Cfg::G, GH = Cfg::GH, HS = Cfg::HS, NT = Cfg::NT;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* W_s = reinterpret_cast<float*>(smem_raw);   // [HS][G*H]  rows of W_hh^T = columns of W_hh
  float* d_s = W_s + (size_t)HS * GH;                // [2][BS][G*H] gate gradients of the whole cluster
  uint64_t* bar = reinterpret_cast<uint64_t*>(d_s + 2 * BS * GH);

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint32_t rank = ptx::cluster_ctarank();
  const int cid = blockIdx.x / C;
  const int dir = cid / nslices;
  const int slice = cid - dir * nslices;
  const int b0 = slice * BS;
  const int j0 = (int)rank * HS;
  const int B = p.B, T = p.T;

  if (tid == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    ptx::mbar_arrive_expect_tx(bar, (uint32_t)Cfg::W_BYTES);
    // HS consecutive rows of W_hh^T are one contiguous block; split so each copy stays well below 2^20 B
    constexpr int NCHUNK = 4;
    constexpr uint32_t CH = (uint32_t)(Cfg::W_BYTES / NCHUNK);
    const char* src = reinterpret_cast<const char*>(p.w_hh_t[dir] + (size_t)j0 * GH);
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c)
      ptx::tma_bulk_g2s(reinterpret_cast<char*>(W_s) + (size_t)c * CH, src + (size_t)c * CH, CH, bar);
  }
  for (int i = tid; i < 2 * BS * GH; i += NT) d_s[i] = 0.f;
  ptx::mbar_wait(bar, 0);
  __syncthreads();
  ptx::cluster_sync_all();

  const int pu = LM::p(lane), qb = LM::q(lane), rep = LM::rep(lane);
  const int j = j0 + w * UPW + pu;
  const int b = b0 + qb;
  const bool valid = b < B;
  const float* gates = p.gates[dir];
  const float* extra = p.extra[dir];
  float* dgates = p.dgates[dir];

  float dh_carry = 0.f, dc_carry = 0.f;
  if (valid) {
    if (p.dh_n) dh_carry = p.dh_n[((size_t)dir * B + b) * H + j];
    if (MODE == B200RNN_LSTM && p.dc_n) dc_carry = p.dc_n[((size_t)dir * B + b) * H + j];
  }
  float bsum[G + 1];
#pragma unroll
  for (int g = 0; g <= G; ++g) bsum[g] = 0.f;

  // operands of the current step (prefetched one step ahead)
  float sv[G], sx = 0.f, hp = 0.f, dyv = 0.f;  // saved gates, hn / c_t, h_{prev} / c_{prev}, dy
#pragma unroll
  for (int g = 0; g < G; ++g) sv[g] = 0.f;
  auto load_step = [&](int step) {
    const int t = dir ? step : (T - 1 - step);
    const bool has_prev = step < T - 1;
    const int tp = dir ? t + 1 : t - 1;
    const float* gp = gates + ((size_t)t * B + b) * GH + j;
#pragma unroll
    for (int g = 0; g < G; ++g) sv[g] = gp[g * H];
    sx = extra[((size_t)t * B + b) * H + j];
    dyv = p.dy[(long long)t * p.dy_st + (long long)b * p.dy_sb + dir * H + j];
    if (MODE == B200RNN_GRU)
      hp = has_prev ? p.y[(long long)tp * p.y_st + (long long)b * p.y_sb + dir * H + j] : 0.f;
    else
      hp = has_prev ? extra[((size_t)tp * B + b) * H + j] : 0.f;
  };
  if (valid && T > 0) load_step(0);

  for (int step = 0; step < T; ++step) {
    const int t = dir ? step : (T - 1 - step);
    float* d_nxt = d_s + (step & 1) * BS * GH;
    const bool last = (step == T - 1);

    // ---- cell backward for (unit j, batch b) ----------------------------------------------------
    const float dh = dh_carry + dyv;
    float dg[G], direct, dhn = 0.f;
    if (MODE == B200RNN_GRU) {
      const float r = sv[0], z = sv[1], n = sv[2], hn = sx;
      const float dn = dh * (1.f - z) * (1.f - n * n);
      const float dz = dh * (hp - n) * z * (1.f - z);
      const float dr = dn * hn * r * (1.f - r);
      dhn = dn * r;
      dg[0] = dr; dg[1] = dz; dg[2] = dn;
      direct = dh * z;
    } else {
      const float ig = sv[0], fg = sv[1], gg = sv[2], og = sv[G - 1];
      const float tc = tanhf(sx);
      const float dout = dh * tc * og * (1.f - og);
      const float dc = dc_carry + dh * og * (1.f - tc * tc);
      dg[0] = dc * gg * ig * (1.f - ig);
      dg[1] = dc * hp * fg * (1.f - fg);
      dg[2] = dc * ig * (1.f - gg * gg);
      dg[G - 1] = dout;
      dc_carry = dc * fg;
      direct = 0.f;
    }
    if (valid) {
#pragma unroll
      for (int g = 0; g < G; ++g) bsum[g] += dg[g];
      bsum[G] += dhn;
    }

    if (!last) {
      // all-gather the recurrent-side gate gradient (GRU: n-gate part is dn*r) into every peer CTA
      const uint32_t laddr = ptx::smem_u32(&d_nxt[qb * GH + j]);
      for (int rk = rep; rk < C; rk += LM::NREP) {
        const uint32_t ra = ptx::mapa(laddr, (uint32_t)rk);
#pragma unroll
        for (int g = 0; g < G; ++g) {
          float v = dg[g];
          if (MODE == B200RNN_GRU && g == 2) v = dhn;
          if (!valid) v = 0.f;
          ptx::st_cluster_f32(ra + (uint32_t)(g * H * sizeof(float)), v);
        }
      }
      __syncwarp();
      ptx::cluster_arrive_release();
    }

    if (valid && rep == 0) {
      float* gp = dgates + ((size_t)t * B + b) * GH + j;
#pragma unroll
      for (int g = 0; g < G; ++g) gp[g * H] = dg[g];
      if (MODE == B200RNN_GRU) p.dghn[dir][((size_t)t * B + b) * H + j] = dhn;
    }
    if (last) break;
    if (valid) load_step(step + 1);
    __syncwarp();
    ptx::cluster_wait_acquire();

    // ---- dh_{prev}[b][j] = direct + sum_col dgh[b][col] * W_hh[col][j] -----------------------------
    float acc[1][UPW][BS];
    warp_partial_dots<1, UPW, BS, GH>(W_s, 0, w * UPW, d_nxt, lane, acc);
    warp_transpose_reduce<1, UPW, BS>(acc);
    dh_carry = direct + acc[0][0][0];
  }

  // ---- per-slice bias-gradient partials: sum over this slice's batch rows (lane bits of q) ----------
  constexpr unsigned FULL = 0xffffffffu;
#pragma unroll
  for (int g = 0; g <= G; ++g) {
    float v = bsum[g];
#pragma unroll
    for (int off = LM::NREP; off < LM::NREP * BS; off <<= 1) v += __shfl_xor_sync(FULL, v, off);
    bsum[g] = v;
  }
  if (qb == 0 && rep == 0) {
    float* out = p.dbias_part[dir] + (size_t)slice * (G + 1) * H;
#pragma unroll
    for (int g = 0; g < G; ++g) out[g * H + j] = bsum[g];
    out[G * H + j] = bsum[G];
  }
}

// =================================================================================================
// launchers
// =================================================================================================
template <typename K>
int prepare_kernel(K kernel, size_t smem) {
  static std::mutex mu;  // forward and autograd-backward threads both launch
  static const void* done[64];
  static int ndone = 0;
  std::lock_guard<std::mutex> lk(mu);
  for (int i = 0; i < ndone; ++i)
    if (done[i] == (const void*)kernel) return B200RNN_OK;
  B200_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (ndone < 64) done[ndone++] = (const void*)kernel;
  return B200RNN_OK;
}

template <typename K, typename P>
int launch_clustered(K kernel, const P& params, int nslices, int nclusters, int C, int NT, size_t smem,
                     cudaStream_t stream, int prof_kind) {
  int rc = prepare_kernel(kernel, smem);
  if (rc) return rc;
  ProfScope prof(prof_kind, stream);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(nclusters * C), 1, 1);
  cfg.blockDim = dim3((unsigned)NT, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  B200_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, params, nslices));
  count_launch();
  return B200RNN_OK;
}

// how many clusters of this kernel can be resident at once (cached per kernel)
template <typename K>
int max_active_clusters(K kernel, int C, int NT, size_t smem) {
  struct Entry {
    const void* k;
    int n;
  };
  static std::mutex mu;
  static Entry cache[64];
  static int ncache = 0;
  std::lock_guard<std::mutex> lk(mu);
  for (int i = 0; i < ncache; ++i)
    if (cache[i].k == (const void*)kernel) return cache[i].n;
  if (prepare_kernel(kernel, smem) != B200RNN_OK) return 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(C * 148), 1, 1);
  cfg.blockDim = dim3((unsigned)NT, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  if (ncache < 64) cache[ncache++] = Entry{(const void*)kernel, n};
  return n;
}

template <int MODE, int H, int C, int BS, int UPW>
bool try_fwd(const RecFwdParams& p, cudaStream_t s, bool force, int* rc) {
  using Cfg = FwdCfg<MODE, H, C, BS, UPW>;
  auto k = rec_fwd_kernel<MODE, H, C, BS, UPW>;
  const int nslices = (p.B + BS - 1) / BS;
  const int nclusters = nslices * p.D;
  if (!force && nclusters > max_active_clusters(k, C, Cfg::NT, Cfg::SMEM)) return false;
  *rc = launch_clustered(k, p, nslices, nclusters, C, Cfg::NT, Cfg::SMEM, s, PROF_REC_FWD);
  return true;
}

template <int MODE, int H, int C, int BS, int UPW>
bool try_bwd(RecBwdParams& p, cudaStream_t s, bool force, int* rc) {
  using Cfg = BwdCfg<MODE, H, C, BS, UPW>;
  auto k = rec_bwd_kernel<MODE, H, C, BS, UPW>;
  const int nslices = (p.B + BS - 1) / BS;
  const int nclusters = nslices * p.D;
  if (!force && nclusters > max_active_clusters(k, C, Cfg::NT, Cfg::SMEM)) return false;
  p.nslices_out = nslices;
  *rc = launch_clustered(k, p, nslices, nclusters, C, Cfg::NT, Cfg::SMEM, s, PROF_REC_BWD);
  return true;
}

}  // namespace

// smallest BS any backward config uses is 2
int rec_bwd_max_slices(int B) { return (B + 1) / 2; }

// Candidates are ordered by batch rows per cluster; the first one whose clusters are all co-resident
// (one wave => every sequence advances in lock step) wins, else the widest one runs in several waves.
int launch_rec_fwd(const RecFwdParams& p, cudaStream_t s) {
  int rc = B200RNN_OK;
  if (p.B <= 0 || p.T <= 0) return rc;
  if (p.mode == B200RNN_GRU && p.H == 256) {
    if (try_fwd<B200RNN_GRU, 256, 4, 2, 4>(p, s, false, &rc)) return rc;
    if (try_fwd<B200RNN_GRU, 256, 4, 4, 4>(p, s, false, &rc)) return rc;
    try_fwd<B200RNN_GRU, 256, 8, 8, 2>(p, s, true, &rc);
    return rc;
  }
  if (p.mode == B200RNN_GRU && p.H == 128) {
    if (try_fwd<B200RNN_GRU, 128, 2, 2, 4>(p, s, false, &rc)) return rc;
    if (try_fwd<B200RNN_GRU, 128, 2, 4, 4>(p, s, false, &rc)) return rc;
    try_fwd<B200RNN_GRU, 128, 4, 8, 2>(p, s, true, &rc);
    return rc;
  }
  if (p.mode == B200RNN_LSTM && p.H == 256) {
    if (try_fwd<B200RNN_LSTM, 256, 8, 2, 4>(p, s, false, &rc)) return rc;
    if (try_fwd<B200RNN_LSTM, 256, 8, 4, 4>(p, s, false, &rc)) return rc;
    try_fwd<B200RNN_LSTM, 256, 8, 8, 2>(p, s, true, &rc);
    return rc;
  }
  if (p.mode == B200RNN_LSTM && p.H == 128) {
    if (try_fwd<B200RNN_LSTM, 128, 2, 2, 4>(p, s, false, &rc)) return rc;
    if (try_fwd<B200RNN_LSTM, 128, 2, 4, 4>(p, s, false, &rc)) return rc;
    try_fwd<B200RNN_LSTM, 128, 4, 8, 2>(p, s, true, &rc);
    return rc;
  }
  set_error("recurrence: unsupported (mode=%d, hidden_size=%d); built for hidden_size 128 and 256", p.mode,
            p.H);
  return B200RNN_ERR_UNSUPPORTED;
}

int launch_rec_bwd(RecBwdParams& p, cudaStream_t s) {
  int rc = B200RNN_OK;
  if (p.B <= 0 || p.T <= 0) return rc;
  if (p.mode == B200RNN_GRU && p.H == 256) {
    if (try_bwd<B200RNN_GRU, 256, 4, 2, 8>(p, s, false, &rc)) return rc;
    if (try_bwd<B200RNN_GRU, 256, 4, 4, 8>(p, s, false, &rc)) return rc;
    try_bwd<B200RNN_GRU, 256, 8, 8, 4>(p, s, true, &rc);
    return rc;
  }
  if (p.mode == B200RNN_GRU && p.H == 128) {
    if (try_bwd<B200RNN_GRU, 128, 2, 2, 8>(p, s, false, &rc)) return rc;
    if (try_bwd<B200RNN_GRU, 128, 2, 4, 8>(p, s, false, &rc)) return rc;
    try_bwd<B200RNN_GRU, 128, 4, 8, 4>(p, s, true, &rc);
    return rc;
  }
  if (p.mode == B200RNN_LSTM && p.H == 256) {
    if (try_bwd<B200RNN_LSTM, 256, 8, 2, 4>(p, s, false, &rc)) return rc;
    if (try_bwd<B200RNN_LSTM, 256, 8, 4, 4>(p, s, false, &rc)) return rc;
    try_bwd<B200RNN_LSTM, 256, 8, 8, 4>(p, s, true, &rc);
    return rc;
  }
  if (p.mode == B200RNN_LSTM && p.H == 128) {
    if (try_bwd<B200RNN_LSTM, 128, 2, 2, 8>(p, s, false, &rc)) return rc;
    if (try_bwd<B200RNN_LSTM, 128, 2, 4, 8>(p, s, false, &rc)) return rc;
    try_bwd<B200RNN_LSTM, 128, 4, 8, 4>(p, s, true, &rc);
    return rc;
  }
  set_error("recurrence backward: unsupported (mode=%d, hidden_size=%d)", p.mode, p.H);
  return B200RNN_ERR_UNSUPPORTED;
}

}  // namespace b200rnn
