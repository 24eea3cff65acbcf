// rnn_rec_tc.cu — tensor-core forward recurrence (K2/K3 on tcgen05) for sm_100a.
//
// Same contract as rec_fwd_kernel (rnn_rec.cu): one launch runs every time step of every direction of one layer of
// the GRU / LSTM the reference reaches through torch.nn.GRU / torch.nn.LSTM (cell equations torch rnn.py:1221-1224,
// :842-847; call sites audio_gru_whole.py:105, text_bilstm_whole.py:105, fuse_net_whole.py:347,361). What changes is
// where the h_{t-1} x W_hh^T contraction runs: the FFMA kernel is bound by the fp32 pipe and the shared-memory
// bandwidth of re-reading W_hh every step (1 536 cycles of each per step, profiles/README.md); here the weights sit
// in shared memory in the tensor core's operand format and the contraction is ~80 tcgen05.mma per step.
//
//   * cluster of C = H/32 CTAs owns NB <= 16 batch rows; CTA `rank` owns hidden units [32 rank, 32 rank + 32) and the
//     G*32 gate rows of W_hh that produce them (row m = g*32 + u), K-major, 128-byte swizzled, resident for the
//     whole sequence:  A_hi = tf32(W) in 32-bit containers, A_lo = bf16(W - A_hi).
//   * per step D[m, n] = sum_k W[m, k] h[n, k] as M128 x N16 MMAs into TMEM:
//         A_hi x tf32(h)          kind::tf32   (K = 8 per instruction)  -> three round-robin accumulators
//         A_hi x (h - tf32(h))    kind::tf32                             -> cross accumulator
//         A_lo x bf16(h)          kind::f16    (K = 16 per instruction)  -> cross accumulator
//     i.e. the exact fp32 product up to the bf16 rounding of the 2^-11-sized correction terms (2^-19 relative);
//     accumulators are cut in four because the tensor core adds into fp32 TMEM with truncation (gemm_tc.cu).
//   * epilogue: warp g reads gate block g from TMEM (lane = unit), the blocks meet through shared memory, every
//     thread then owns (unit, batch row) pairs: non-linearities and state update in registers, as in the FFMA kernel.
//   * exchange: the new state slice goes to all C CTAs as fp32 with st.async (16 B + mbarrier complete_tx, double
//     buffered); the RECEIVER splits it into the three MMA operand forms — one third of the DSMEM bytes of sending
//     the split forms.
//
// STATUS (round 1): parity-green (all -m gpu tests pass with B200RNN_REC_TC=1; y error vs torch CPU 7e-7 at
// B=128, T=120, two layers) but NOT the default: 656 us per GRU layer launch against 231 us for the FFMA kernel.
// clock64 timeline of one step (profiles/r01_tc_rec_trace.txt, B=128 -> NB=9, C=8), 9 700 cycles:
//     wait for h slices 510 | split pass + fence + barrier 765 | MMA issue 4 620 | commit -> mbarrier 160 |
//     tcgen05.ld + smem meet + barrier 580 | cell update + st.async exchange 2 950
// What that says: (1) with both operands in shared memory an M128 x N16 MMA costs ~58 cycles and ~116 when it
// depends on the previous one through its accumulator (skipping half of the MMAs changed nothing while they still
// alternated over the same accumulators) - the weights must live in TMEM (tcgen05.mma with A from tensor memory,
// 384 of the 512 columns for A_hi + A_lo) and the 80 MMAs must rotate over >= 8 accumulators; (2) an all-to-all of
// the fp32 state over C = 8 CTAs moves 8 KB per CTA and step through DSMEM, ~3x the FFMA kernel's C = 4 exchange;
// (3) every phase boundary is a CTA barrier - the phases have to be overlapped across two independent batch groups.
// That redesign is the round-2 lever; this file is the verified starting point (descriptors, swizzle, both MMA kinds,
// TMEM epilogue and the exchange protocol are known-good).
#include <cuda_bf16.h>
#include <mutex>
#include <stdlib.h>

#include "profile.cuh"
#include "ptx.cuh"
#include "rnn_kernels.cuh"

namespace b200rnn {

namespace {

constexpr int TC_MAX_SMEM = 232448;
constexpr int NBMAX = 16;  // MMA N
constexpr int UCTA = 32;   // hidden units per CTA
constexpr int TC_NT = 128;
constexpr unsigned FULLMASK = 0xffffffffu;

template <int MODE, int H>
struct TcCfg {
  static constexpr int G = (MODE == B200RNN_GRU) ? 3 : 4;
  static constexpr int C = H / UCTA;
  static constexpr int MR = G * UCTA;              // real gate rows per CTA (96 / 128); the MMA reads 128
  static constexpr int NKB = H / 32;               // tf32 K-blocks (32 floats = one 128-byte swizzled row)
  static constexpr int NKB2 = H / 64;              // bf16 K-blocks (64 bf16)
  static constexpr int A_KB = MR * 128;            // bytes per K-block of A
  static constexpr int B_KB = NBMAX * 128;         // bytes per K-block of B
  static constexpr int OFF_AHI = 0;
  static constexpr int OFF_ALO = OFF_AHI + NKB * A_KB;
  static constexpr int OFF_BHI = OFF_ALO + NKB2 * A_KB;
  static constexpr int OFF_BLO = OFF_BHI + NKB * B_KB;
  static constexpr int OFF_BBF = OFF_BLO + NKB * B_KB;
  static constexpr int OFF_RECV = OFF_BBF + NKB2 * B_KB;           // [2][NBMAX][H] fp32
  static constexpr int OFF_PRE = OFF_RECV + 2 * NBMAX * H * 4;     // [G][NBMAX][32] fp32
  static constexpr int OFF_BAR = OFF_PRE + G * NBMAX * UCTA * 4;   // recv[2], mma, tmem slot
  static constexpr int SMEM = OFF_BAR + 64 + 1024 /*alignment slack*/;
  static_assert(SMEM <= TC_MAX_SMEM, "tensor-core recurrence does not fit an SM for this shape");
  static_assert(OFF_ALO % 1024 == 0 && OFF_BHI % 1024 == 0 && OFF_BLO % 1024 == 0 && OFF_BBF % 1024 == 0 &&
                    A_KB % 1024 == 0,
                "swizzle atoms must stay 1024-byte aligned");
  static_assert(C <= 8, "portable cluster size");
};

__device__ __forceinline__ void tcr_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcr_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcr_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   ptx::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tcr_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tcr_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tcr_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tcr_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte swizzle, rows of 128 B, 8-row groups 1024 B apart (same encoding as gemm_tc.cu)
__device__ __forceinline__ uint64_t tcr_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// D = f32, both operands K-major, M = 128, N = 16; tf32 x tf32 and bf16 x bf16
constexpr uint32_t IDESC_TF32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NBMAX >> 3) << 17) | (8u << 24);
constexpr uint32_t IDESC_BF16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NBMAX >> 3) << 17) | (8u << 24);

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t t;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(x));
  return t;
}
// byte offset of the 16-byte chunk `c` (0..7) of row `r` inside a 128-byte-swizzled K-block
__device__ __forceinline__ uint32_t sw128(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

// bounded wait: a protocol bug must end in a trap (reported as a CUDA error), never in a hung GPU
__device__ __forceinline__ void tcr_wait(uint64_t* bar, uint32_t parity) {
  if (ptx::mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!ptx::mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) __trap();  // ~2 s
  }
}

__device__ __forceinline__ float sigm(float x) { return sigmoid_f(x); }

template <int MODE, int H>
__global__ void __launch_bounds__(TC_NT, 1) rec_fwd_tc_kernel(const RecFwdParams p, const int nslices, const int NB) {
  using Cfg = TcCfg<MODE, H>;
  constexpr int G = Cfg::G, C = Cfg::C, MR = Cfg::MR, GH = G * H;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base =
      reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* a_hi = base + Cfg::OFF_AHI;
  unsigned char* a_lo = base + Cfg::OFF_ALO;
  unsigned char* b_hi = base + Cfg::OFF_BHI;
  unsigned char* b_lo = base + Cfg::OFF_BLO;
  unsigned char* b_bf = base + Cfg::OFF_BBF;
  float* recv = reinterpret_cast<float*>(base + Cfg::OFF_RECV);  // [2][NBMAX][H]
  float* pre = reinterpret_cast<float*>(base + Cfg::OFF_PRE);    // [G][NBMAX][32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + Cfg::OFF_BAR);  // [0],[1] state buffers, [2] MMA done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = ptx::cluster_ctarank();
  const int cid = blockIdx.x / C;
  const int dir = cid / nslices;
  const int slice = cid - dir * nslices;
  const int b0 = slice * NB;
  const int j0 = (int)rank * UCTA;
  const int B = p.B, T = p.T;
  const float* w_hh = p.w_hh[dir];

  if (tid == 0) {
    ptx::mbar_init(&bars[0], 1);
    ptx::mbar_init(&bars[1], 1);
    ptx::mbar_init(&bars[2], 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(tmem_slot)),
                 "r"(64u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }

  // ---- W_hh rows of this CTA -> A_hi (tf32) / A_lo (bf16 of the remainder), swizzled K-major ------------------
  for (int i = tid; i < MR * (H / 4); i += TC_NT) {
    const int m = i / (H / 4), k = (i - m * (H / 4)) * 4;
    const int g = m / UCTA, u = m - g * UCTA;
    const float4 x = __ldg(reinterpret_cast<const float4*>(w_hh + ((size_t)g * H + j0 + u) * H + k));
    uint4 hi;
    hi.x = to_tf32(x.x); hi.y = to_tf32(x.y); hi.z = to_tf32(x.z); hi.w = to_tf32(x.w);
    const int kb = k >> 5, c = (k & 31) >> 2;
    *reinterpret_cast<uint4*>(a_hi + kb * Cfg::A_KB + sw128(m, c)) = hi;
    const __nv_bfloat162 l01 = __floats2bfloat162_rn(x.x - __uint_as_float(hi.x), x.y - __uint_as_float(hi.y));
    const __nv_bfloat162 l23 = __floats2bfloat162_rn(x.z - __uint_as_float(hi.z), x.w - __uint_as_float(hi.w));
    uint2 lo;
    lo.x = *reinterpret_cast<const uint32_t*>(&l01);
    lo.y = *reinterpret_cast<const uint32_t*>(&l23);
    const int kb2 = k >> 6, c2 = (k & 63) >> 3, half = (k >> 2) & 1;
    *reinterpret_cast<uint2*>(a_lo + kb2 * Cfg::A_KB + sw128(m, c2) + half * 8) = lo;
  }
  // operand rows >= NB are never written by the split pass: keep them finite
  for (int i = tid; i < (Cfg::OFF_RECV - Cfg::OFF_BHI) / 16; i += TC_NT)
    reinterpret_cast<uint4*>(b_hi)[i] = make_uint4(0, 0, 0, 0);
  ptx::fence_proxy_async();
  tcr_fence_before();
  __syncthreads();
  tcr_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0 && T > 1) ptx::mbar_arrive_expect_tx(&bars[1], (uint32_t)(NB * H * sizeof(float)));
  ptx::cluster_sync_all();  // peers' barriers are initialised before anyone stores into them

  // ---- thread identity for the cell update: unit u, batch rows n = q + 4 i ------------------------------------
  const int u = lane, q = warp;
  const int j = j0 + u;
  float* gates = p.gates[dir];
  float* extra = p.extra[dir];
  const float bhn = (MODE == B200RNN_GRU) ? p.b_hh[dir][2 * H + j] : 0.f;
  float h_prev[4], c_prev[4], h_sum[4], gi[4][G];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h_prev[i] = c_prev[i] = h_sum[i] = 0.f;
    const int n = q + 4 * i, b = b0 + n;
#pragma unroll
    for (int g = 0; g < G; ++g) gi[i][g] = 0.f;
    if (n < NB && b < B) {
      const int t0 = dir ? T - 1 : 0;
      const float* gp = gates + ((size_t)t0 * B + b) * GH + j;
#pragma unroll
      for (int g = 0; g < G; ++g) gi[i][g] = gp[g * H];
    }
  }

  const uint64_t d_ahi = tcr_desc(ptx::smem_u32(a_hi)), d_alo = tcr_desc(ptx::smem_u32(a_lo));
  const uint64_t d_bhi = tcr_desc(ptx::smem_u32(b_hi)), d_blo = tcr_desc(ptx::smem_u32(b_lo));
  const uint64_t d_bbf = tcr_desc(ptx::smem_u32(b_bf));

  for (int step = 0; step < T; ++step) {
    const int t = dir ? (T - 1 - step) : step;
    const int cur = step & 1, nxt = cur ^ 1;
#ifdef B200RNN_TRACE
    const bool tr = p.trace != nullptr && blockIdx.x == 0 && tid == 0;
#else
    constexpr bool tr = false;  // -DB200RNN_TRACE: per-phase clock64 timeline (tools/trace_rec_tc.py)
#endif
    long long* trow = p.trace + (size_t)step * 8;
    if (tr) trow[0] = clock64();

    if (step > 0) {
      // ---- h_step has arrived from every CTA of the cluster: split it into the MMA operand forms --------------
      tcr_wait(&bars[cur], (uint32_t)(((step - 1) >> 1) & 1));
      if (tr) trow[1] = clock64();
      const float* hc = recv + cur * NBMAX * H;
      for (int i = tid; i < NB * (H / 4); i += TC_NT) {
        const int n = i / (H / 4), k = (i - n * (H / 4)) * 4;
        const float4 x = *reinterpret_cast<const float4*>(hc + n * H + k);
        uint4 hi;
        hi.x = to_tf32(x.x); hi.y = to_tf32(x.y); hi.z = to_tf32(x.z); hi.w = to_tf32(x.w);
        float4 lo;
        lo.x = x.x - __uint_as_float(hi.x); lo.y = x.y - __uint_as_float(hi.y);
        lo.z = x.z - __uint_as_float(hi.z); lo.w = x.w - __uint_as_float(hi.w);
        const int kb = k >> 5, c = (k & 31) >> 2;
        const uint32_t o = (uint32_t)(kb * Cfg::B_KB) + sw128(n, c);
        *reinterpret_cast<uint4*>(b_hi + o) = hi;
        *reinterpret_cast<float4*>(b_lo + o) = lo;
        const __nv_bfloat162 v01 = __floats2bfloat162_rn(x.x, x.y), v23 = __floats2bfloat162_rn(x.z, x.w);
        uint2 bf;
        bf.x = *reinterpret_cast<const uint32_t*>(&v01);
        bf.y = *reinterpret_cast<const uint32_t*>(&v23);
        const int kb2 = k >> 6, c2 = (k & 63) >> 3, half = (k >> 2) & 1;
        *reinterpret_cast<uint2*>(b_bf + kb2 * Cfg::B_KB + sw128(n, c2) + half * 8) = bf;
      }
      ptx::fence_proxy_async();
      tcr_fence_before();
      __syncthreads();
      if (tr) trow[2] = clock64();
      if (tid == 0) {
        tcr_fence_after();
#pragma unroll
        for (int kb = 0; kb < Cfg::NKB; ++kb) {
          const uint32_t acc_main = tmem_base + (uint32_t)((kb % 3) * NBMAX);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t a = d_ahi + (uint64_t)((kb * Cfg::A_KB + k * 32) >> 4);
            const uint64_t bh = d_bhi + (uint64_t)((kb * Cfg::B_KB + k * 32) >> 4);
            const uint64_t bl = d_blo + (uint64_t)((kb * Cfg::B_KB + k * 32) >> 4);
            tcr_mma_tf32(acc_main, a, bh, IDESC_TF32, (kb >= 3 || k != 0) ? 1u : 0u);
            tcr_mma_tf32(tmem_base + 3 * NBMAX, a, bl, IDESC_TF32, (kb | k) != 0 ? 1u : 0u);
          }
        }
#pragma unroll
        for (int kb2 = 0; kb2 < Cfg::NKB2; ++kb2)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tcr_mma_bf16(tmem_base + 3 * NBMAX, d_alo + (uint64_t)((kb2 * Cfg::A_KB + k * 32) >> 4),
                         d_bbf + (uint64_t)((kb2 * Cfg::B_KB + k * 32) >> 4), IDESC_BF16, 1u);
        tcr_commit(&bars[2]);
        if (tr) trow[3] = clock64();
      }
      __syncwarp();
    }
    // h_step is consumed (copied into the operand buffers): re-arm the other buffer for h_{step+1}... which is
    // bars[nxt]; for step 0 that was done before the loop
    if (tid == 0 && step > 0 && step + 1 < T) ptx::mbar_arrive_expect_tx(&bars[nxt], (uint32_t)(NB * H * sizeof(float)));

    if (step > 0) {
      tcr_wait(&bars[2], (uint32_t)((step - 1) & 1));
      tcr_fence_after();
      if (tr) trow[4] = clock64();
      if (warp < G) {
        const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
        uint32_t r0[16], r1[16], r2[16], r3[16];
        tcr_ld_x16(lane_base + 0 * NBMAX, r0);
        tcr_ld_x16(lane_base + 1 * NBMAX, r1);
        tcr_ld_x16(lane_base + 2 * NBMAX, r2);
        tcr_ld_x16(lane_base + 3 * NBMAX, r3);
        tcr_wait_ld();
        float* dst = pre + (size_t)warp * NBMAX * UCTA + lane;
#pragma unroll
        for (int n = 0; n < NBMAX; ++n)
          if (n < NB)
            dst[n * UCTA] = (__uint_as_float(r0[n]) + __uint_as_float(r1[n])) +
                            (__uint_as_float(r2[n]) + __uint_as_float(r3[n]));
      }
      tcr_fence_before();
      __syncthreads();
      if (tr) trow[5] = clock64();
    }

    // ---- cell update for (unit u, rows q, q+4, ...) ------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = q + 4 * i;
      if (n >= NB) break;
      const int b = b0 + n;
      const bool valid = b < B;
      float a[G];
#pragma unroll
      for (int g = 0; g < G; ++g) a[g] = (step > 0) ? pre[((size_t)g * NBMAX + n) * UCTA + u] : 0.f;
      float hnew, s0, s1, s2, s3 = 0.f, sx;
      if (MODE == B200RNN_GRU) {
        const float r = sigm(gi[i][0] + a[0]);
        const float z = sigm(gi[i][1] + a[1]);
        const float hn = a[2] + bhn;
        const float nn = tanh_f(gi[i][2] + r * hn);
        hnew = nn + z * (h_prev[i] - nn);
        s0 = r; s1 = z; s2 = nn; sx = hn;
      } else {
        const float ig = sigm(gi[i][0] + a[0]);
        const float fg = sigm(gi[i][1] + a[1]);
        const float gg = tanh_f(gi[i][2] + a[2]);
        const float og = sigm(gi[i][G - 1] + a[G - 1]);
        const float cnew = fg * c_prev[i] + ig * gg;
        hnew = og * tanh_f(cnew);
        c_prev[i] = cnew;
        s0 = ig; s1 = fg; s2 = gg; s3 = og; sx = cnew;
      }
      h_prev[i] = hnew;
      h_sum[i] += hnew;

      if (step + 1 < T) {
        // 4 consecutive units -> one 16-byte st.async per destination CTA; lane&3 picks the destinations
        float4 v;
        v.x = __shfl_sync(FULLMASK, hnew, (lane & ~3) + 0);
        v.y = __shfl_sync(FULLMASK, hnew, (lane & ~3) + 1);
        v.z = __shfl_sync(FULLMASK, hnew, (lane & ~3) + 2);
        v.w = __shfl_sync(FULLMASK, hnew, (lane & ~3) + 3);
        const uint32_t dst = ptx::smem_u32(recv + (size_t)nxt * NBMAX * H + n * H + j0 + (lane & ~3));
        const uint32_t bar = ptx::smem_u32(&bars[nxt]);
#pragma unroll
        for (int r = (lane & 3); r < C; r += 4) ptx::st_async_v4(ptx::mapa(dst, (uint32_t)r), v, ptx::mapa(bar, (uint32_t)r));
      }

      if (valid) {
        if (p.y) p.y[(long long)t * p.y_st + (long long)b * p.y_sb + dir * H + j] = hnew;
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
          if (p.y_pool) p.y_pool[(size_t)b * p.D * H + dir * H + j] = h_sum[i];
          if (MODE == B200RNN_LSTM && p.c_n) p.c_n[((size_t)dir * B + b) * H + j] = c_prev[i];
        }
        if (step + 1 < T) {
          const int tn = dir ? (T - 2 - step) : (step + 1);
          const float* gp = gates + ((size_t)tn * B + b) * GH + j;
#pragma unroll
          for (int g = 0; g < G; ++g) gi[i][g] = gp[g * H];
        }
      }
    }
    if (tr) trow[6] = clock64();
  }
  tcr_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcr_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
  }
  ptx::cluster_sync_all();  // nobody exits while a peer could still address its shared memory
}

template <typename K>
int tc_prepare(K kernel, size_t smem) {
  static std::mutex mu;
  static const void* done[16];
  static int ndone = 0;
  std::lock_guard<std::mutex> lk(mu);
  for (int i = 0; i < ndone; ++i)
    if (done[i] == (const void*)kernel) return B200RNN_OK;
  B200_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (ndone < 16) done[ndone++] = (const void*)kernel;
  return B200RNN_OK;
}

template <typename K>
int tc_capacity(K kernel, int C, size_t smem) {
  static std::mutex mu;
  static const void* keys[16];
  static int vals[16];
  static int n = 0;
  {
    std::lock_guard<std::mutex> lk(mu);
    for (int i = 0; i < n; ++i)
      if (keys[i] == (const void*)kernel) return vals[i];
  }
  if (tc_prepare(kernel, smem) != B200RNN_OK) return 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(C * 148), 1, 1);
  cfg.blockDim = dim3((unsigned)TC_NT, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int cap = 0;
  if (cudaOccupancyMaxActiveClusters(&cap, kernel, &cfg) != cudaSuccess) {
    cudaGetLastError();
    cap = 0;
  }
  std::lock_guard<std::mutex> lk(mu);
  if (n < 16) {
    keys[n] = (const void*)kernel;
    vals[n++] = cap;
  }
  return cap;
}

template <int MODE, int H>
bool try_fwd_tc(const RecFwdParams& p, cudaStream_t s, int* rc) {
  using Cfg = TcCfg<MODE, H>;
  auto k = rec_fwd_tc_kernel<MODE, H>;
  const int cap = tc_capacity(k, Cfg::C, Cfg::SMEM);
  const int per_dir = cap / p.D;  // clusters available to one direction
  static const bool debug = getenv("B200RNN_DEBUG") != nullptr;
  if (per_dir < 1) return false;
  const int NB = (p.B + per_dir - 1) / per_dir;  // fewest rows per cluster that still runs in ONE wave
  if (debug) fprintf(stderr, "[b200rnn] tc fwd mode=%d H=%d: capacity %d clusters of %d, NB=%d\n", MODE, H, cap, Cfg::C, NB);
  if (NB > NBMAX) return false;
  const int nslices = (p.B + NB - 1) / NB;
  *rc = tc_prepare(k, Cfg::SMEM);
  if (*rc) return true;
  ProfScope prof(PROF_REC_FWD, s);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(nslices * p.D * Cfg::C), 1, 1);
  cfg.blockDim = dim3((unsigned)TC_NT, 1, 1);
  cfg.dynamicSmemBytes = Cfg::SMEM;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)Cfg::C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, k, p, nslices, NB) != cudaSuccess) {
    set_error("tensor-core recurrence launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    *rc = B200RNN_ERR_CUDA;
    return true;
  }
  count_launch();
  *rc = B200RNN_OK;
  return true;
}

}  // namespace

// Returns true when the tensor-core kernel took the launch (rc set); false = shape not covered, use the FFMA kernel.
bool launch_rec_fwd_tc(const RecFwdParams& p, cudaStream_t s, int* rc) {
  if (p.lengths != nullptr) return false;
  if (p.mode == B200RNN_GRU && p.H == 256) return try_fwd_tc<B200RNN_GRU, 256>(p, s, rc);
  if (p.mode == B200RNN_GRU && p.H == 128) return try_fwd_tc<B200RNN_GRU, 128>(p, s, rc);
  if (p.mode == B200RNN_LSTM && p.H == 128) return try_fwd_tc<B200RNN_LSTM, 128>(p, s, rc);
  return false;
}

}  // namespace b200rnn
