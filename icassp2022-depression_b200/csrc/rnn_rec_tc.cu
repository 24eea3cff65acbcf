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
// STATUS (round 1): parity-green (every -m gpu test passes with B200RNN_REC_TC=1; y error vs torch CPU 7e-7 at
// B=128, T=120, two layers) but opt-in, NOT the default: at the benchmark shape (B=128 per GPU) a GRU layer launch
// takes 354 us against 231 us for the FFMA kernel; it is ahead only for small batches (GRU H=256, B <= 45:
// 228-243 us vs 244-254 us per layer forward incl. the input GEMM, tools/rec_crossover.py).
// clock64 timeline of one step (profiles/r01_tc_rec_trace.txt), B=128 -> NB=9 rows per cluster of 8 CTAs:
//     4 350 cycles = wait for h 90-140 | split pass + proxy fence + barrier 660-700 | MMA issue 1 300 (80 MMAs,
//     16 cycles each, A from TMEM) | commit -> mbarrier 100-170 | tcgen05.ld + smem meet + barrier 265 |
//     cell update + st.async exchange 1 850-3 500
// and 3 020 cycles at B=16 (NB=2: exchange 680). What was learned on the way (all measured):
//   (1) a tcgen05.mma issued under `if (tid == 0)` is wrapped by the compiler in an ELECT / R2UR.BROADCAST /
//       BRA.U.ANY loop (operands are not provably warp-uniform): ~55 cycles per MMA, 4 620 per step. Issuing from
//       `if (warp_uniform == 0) if (elect_one_sync())` keeps operands in uniform registers: 16 cycles per MMA.
//       The same fix in gemm_tc.cu took the BiLSTM H=256 train step from 1.35 to 1.09 ms.
//   (2) with A in shared memory an M128xN16 MMA has to read 4 KB of A: the weights belong in TMEM (done: A_hi).
//   (3) a pointer rounded up through uintptr_t loses its shared state space: the loads become generic LD and queue
//       behind the outstanding global prefetches (~900 cycles per batch row). Align by offset instead.
//   (4) what binds now is DSMEM bandwidth: all-to-all of the fp32 state over C = 8 CTAs is 7 x NB x 128 B out and
//       as much in per CTA and step; at the measured ~17 B/cycle (both directions together) that is ~1 100 cycles
//       for NB = 9, against ~360 for the FFMA kernel's C = 4 / 4-row clusters. C = 8 is forced by the operand
//       bytes (A_hi + A_lo = 6 B per weight). A win at B = 128 needs the exchange hidden behind the MMAs of a second
//       batch group, or fewer bytes per weight so that C = 4 fits - the round-2 lever.
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
  static constexpr int OFF_ALO = 0;                                // A_hi lives in TMEM, A_lo (bf16) in shared memory
  static constexpr int OFF_BHI = OFF_ALO + NKB2 * A_KB + 4096;     // +4 KB: the M=128 MMA reads 32 rows past MR
  static constexpr int OFF_BLO = OFF_BHI + NKB * B_KB;
  static constexpr int OFF_BBF = OFF_BLO + NKB * B_KB;
  static constexpr int OFF_RECV = OFF_BBF + NKB2 * B_KB;           // [2][NBMAX][H] fp32
  static constexpr int OFF_PRE = OFF_RECV + 2 * NBMAX * H * 4;     // [G][NBMAX][32] fp32
  static constexpr int OFF_BAR = OFF_PRE + G * NBMAX * UCTA * 4;   // recv[2], mma, tmem slot
  static constexpr int SMEM = OFF_BAR + 256 + 1024 /*alignment slack*/;
  static_assert(SMEM <= TC_MAX_SMEM, "tensor-core recurrence does not fit an SM for this shape");
  static constexpr int COL_AHI = 0;                                // TMEM columns [0, H): A_hi, lane = gate row
  static constexpr int COL_ACC = 384;                              // 8 accumulators of NBMAX columns
  static constexpr int NACC = 8;
  static_assert(H <= COL_ACC, "A_hi must leave room for the accumulators");
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
// A from tensor memory (lane = row, one 32-bit column per K element), B from shared memory
__device__ __forceinline__ void tcr_mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tcr_st_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
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
  // 1024-byte alignment by OFFSET (not by integer round-trip of the pointer): the compiler keeps the shared state
  // space, so the accesses below are LDS/STS. Generic loads would queue behind the outstanding global prefetches.
  unsigned char* base = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* a_lo = base + Cfg::OFF_ALO;
  unsigned char* b_hi = base + Cfg::OFF_BHI;
  unsigned char* b_lo = base + Cfg::OFF_BLO;
  unsigned char* b_bf = base + Cfg::OFF_BBF;
  float* recv = reinterpret_cast<float*>(base + Cfg::OFF_RECV);  // [2][NBMAX][H]
  float* pre = reinterpret_cast<float*>(base + Cfg::OFF_PRE);    // [G][NBMAX][32]
  // [buf*8 + src]: slice of h from CTA `src` has landed in recv[buf] (one barrier per source: 72 st.async
  // completions each instead of 576 on one word); [16]: the step's MMAs are done
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + Cfg::OFF_BAR);
  uint64_t* mma_bar = bars + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(FULLMASK, tid >> 5, 0);  // warp-uniform by construction
  const uint32_t rank = ptx::cluster_ctarank();
  const int cid = blockIdx.x / C;
  const int dir = cid / nslices;
  const int slice = cid - dir * nslices;
  const int b0 = slice * NB;
  const int j0 = (int)rank * UCTA;
  const int B = p.B, T = p.T;
  const float* w_hh = p.w_hh[dir];

  if (tid == 0) {
    for (int i = 0; i < 17; ++i) ptx::mbar_init(&bars[i], 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcr_fence_before();
  __syncthreads();
  tcr_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- W_hh rows of this CTA: A_hi = tf32(W) -> tensor memory (thread = lane = gate row m, column = k) --------
  {
    const int m = tid, g = m / UCTA, uu = m - g * UCTA;
    const float* wrow = w_hh + ((size_t)(m < MR ? g : 0) * H + j0 + uu) * H;
    const uint32_t trow_addr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)Cfg::COL_AHI;
    for (int k0 = 0; k0 < H; k0 += 16) {
      uint32_t r[16];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(wrow + k0 + 4 * v));
        r[4 * v + 0] = (m < MR) ? to_tf32(x.x) : 0u;
        r[4 * v + 1] = (m < MR) ? to_tf32(x.y) : 0u;
        r[4 * v + 2] = (m < MR) ? to_tf32(x.z) : 0u;
        r[4 * v + 3] = (m < MR) ? to_tf32(x.w) : 0u;
      }
      tcr_st_x16(trow_addr + (uint32_t)k0, r);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  // ---- A_lo = bf16(W - A_hi) -> shared memory, swizzled K-major ------------------------------------------------
  for (int i = tid; i < MR * (H / 4); i += TC_NT) {
    const int m = i / (H / 4), k = (i - m * (H / 4)) * 4;
    const int g = m / UCTA, uu = m - g * UCTA;
    const float4 x = __ldg(reinterpret_cast<const float4*>(w_hh + ((size_t)g * H + j0 + uu) * H + k));
    const __nv_bfloat162 l01 = __floats2bfloat162_rn(x.x - __uint_as_float(to_tf32(x.x)), x.y - __uint_as_float(to_tf32(x.y)));
    const __nv_bfloat162 l23 = __floats2bfloat162_rn(x.z - __uint_as_float(to_tf32(x.z)), x.w - __uint_as_float(to_tf32(x.w)));
    uint2 lo;
    lo.x = *reinterpret_cast<const uint32_t*>(&l01);
    lo.y = *reinterpret_cast<const uint32_t*>(&l23);
    const int kb2 = k >> 6, c2 = (k & 63) >> 3, half = (k >> 2) & 1;
    *reinterpret_cast<uint2*>(a_lo + kb2 * Cfg::A_KB + sw128(m, c2) + half * 8) = lo;
  }
  // operand rows >= NB are never written by the split pass, and the M=128 MMA reads 32 rows past A_lo: keep finite
  for (int i = tid; i < (Cfg::OFF_RECV - Cfg::NKB2 * Cfg::A_KB) / 16; i += TC_NT)
    reinterpret_cast<uint4*>(a_lo + Cfg::NKB2 * Cfg::A_KB)[i] = make_uint4(0, 0, 0, 0);
  ptx::fence_proxy_async();
  tcr_fence_before();
  __syncthreads();
  tcr_fence_after();
  if (tid == 0 && T > 1)
    for (int src = 0; src < C; ++src) ptx::mbar_arrive_expect_tx(&bars[8 + src], (uint32_t)(NB * UCTA * sizeof(float)));
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

  const uint64_t d_alo = tcr_desc(ptx::smem_u32(a_lo));
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
      // warp w converts the K-blocks (= source CTAs) [w*KPW, w*KPW + KPW) of all NB rows and waits only for those
      constexpr int KPW = C / 4;
      const uint32_t par = (uint32_t)(((step - 1) >> 1) & 1);
#pragma unroll
      for (int s2 = 0; s2 < KPW; ++s2) tcr_wait(&bars[cur * 8 + warp * KPW + s2], par);
      if (tr) trow[1] = clock64();
      const float* hc = recv + cur * NBMAX * H;
      const int total = NB * KPW * 8;  // 16-byte chunks this warp converts
#pragma unroll 1
      for (int i0 = 0; i0 < total; i0 += 4 * 32) {
        float4 xv[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int i = i0 + v * 32 + lane;
          const int n = i / (KPW * 8), k = (warp * KPW * 8 + (i - n * (KPW * 8))) * 4;
          xv[v] = (i < total) ? *reinterpret_cast<const float4*>(hc + n * H + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int i = i0 + v * 32 + lane;
          if (i < total) {
            const float4 x = xv[v];
            const int n = i / (KPW * 8), k = (warp * KPW * 8 + (i - n * (KPW * 8))) * 4;
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
        }
      }
      ptx::fence_proxy_async();
      tcr_fence_before();
      __syncthreads();
      if (tr) trow[2] = clock64();
      if (warp == 0) {
       const uint32_t tmem_u = __shfl_sync(FULLMASK, tmem_base, 0);
       if (ptx::elect_one_sync()) {
        tcr_fence_after();
        // 5 * NKB/2 MMAs rotate over NACC accumulators so that none waits for its predecessor's accumulate
        int jm = 0;
#pragma unroll
        for (int kb = 0; kb < Cfg::NKB; ++kb) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t a = tmem_u + (uint32_t)(Cfg::COL_AHI + kb * 32 + k * 8);
            const uint64_t bh = d_bhi + (uint64_t)((kb * Cfg::B_KB + k * 32) >> 4);
            const uint64_t bl = d_blo + (uint64_t)((kb * Cfg::B_KB + k * 32) >> 4);
            tcr_mma_tf32_ts(tmem_u + (uint32_t)(Cfg::COL_ACC + (jm % Cfg::NACC) * NBMAX), a, bh, IDESC_TF32,
                            jm >= Cfg::NACC ? 1u : 0u);
            ++jm;
            tcr_mma_tf32_ts(tmem_u + (uint32_t)(Cfg::COL_ACC + (jm % Cfg::NACC) * NBMAX), a, bl, IDESC_TF32,
                            jm >= Cfg::NACC ? 1u : 0u);
            ++jm;
            if (((kb * 4 + k) & 1) == 1) {
              const int i2 = (kb * 4 + k) >> 1, kb2 = i2 >> 2, kk = i2 & 3;
              tcr_mma_bf16(tmem_u + (uint32_t)(Cfg::COL_ACC + (jm % Cfg::NACC) * NBMAX),
                           d_alo + (uint64_t)((kb2 * Cfg::A_KB + kk * 32) >> 4),
                           d_bbf + (uint64_t)((kb2 * Cfg::B_KB + kk * 32) >> 4), IDESC_BF16, jm >= Cfg::NACC ? 1u : 0u);
              ++jm;
            }
          }
        }
        tcr_commit(mma_bar);
       }
       __syncwarp();
       if (tr) trow[3] = clock64();
      }
    }
    // h_step is consumed (copied into the operand buffers): re-arm the other buffer for h_{step+1}... which is
    // bars[nxt]; for step 0 that was done before the loop
    if (tid == 0 && step > 0 && step + 1 < T)
      for (int src = 0; src < C; ++src)
        ptx::mbar_arrive_expect_tx(&bars[nxt * 8 + src], (uint32_t)(NB * UCTA * sizeof(float)));

    if (step > 0) {
      tcr_wait(mma_bar, (uint32_t)((step - 1) & 1));
      tcr_fence_after();
      if (tr) trow[4] = clock64();
      if (warp < G) {
        const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
        uint32_t r0[16], r1[16], r2[16], r3[16];
        float sum[NBMAX];
        tcr_ld_x16(lane_base + (uint32_t)(Cfg::COL_ACC + 0 * NBMAX), r0);
        tcr_ld_x16(lane_base + (uint32_t)(Cfg::COL_ACC + 1 * NBMAX), r1);
        tcr_ld_x16(lane_base + (uint32_t)(Cfg::COL_ACC + 2 * NBMAX), r2);
        tcr_ld_x16(lane_base + (uint32_t)(Cfg::COL_ACC + 3 * NBMAX), r3);
        tcr_wait_ld();
#pragma unroll
        for (int n = 0; n < NBMAX; ++n)
          sum[n] = (__uint_as_float(r0[n]) + __uint_as_float(r1[n])) + (__uint_as_float(r2[n]) + __uint_as_float(r3[n]));
        tcr_ld_x16(lane_base + (uint32_t)(Cfg::COL_ACC + 4 * NBMAX), r0);
        tcr_ld_x16(lane_base + (uint32_t)(Cfg::COL_ACC + 5 * NBMAX), r1);
        tcr_ld_x16(lane_base + (uint32_t)(Cfg::COL_ACC + 6 * NBMAX), r2);
        tcr_ld_x16(lane_base + (uint32_t)(Cfg::COL_ACC + 7 * NBMAX), r3);
        tcr_wait_ld();
        float* dst = pre + (size_t)warp * NBMAX * UCTA + lane;
#pragma unroll
        for (int n = 0; n < NBMAX; ++n)
          if (n < NB)
            dst[n * UCTA] = sum[n] + ((__uint_as_float(r0[n]) + __uint_as_float(r1[n])) +
                                      (__uint_as_float(r2[n]) + __uint_as_float(r3[n])));
      }
      tcr_fence_before();
      __syncthreads();
      if (tr) trow[5] = clock64();
    }

    // ---- cell update for (unit u, rows q, q+4, ...): all rows' math first, then the exchange, then global traffic --
    float hnew[4], sv[4][4], sx[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = q + 4 * i;
      hnew[i] = 0.f;
      if (n >= NB) break;
      float a[G];
#pragma unroll
      for (int g = 0; g < G; ++g) a[g] = (step > 0) ? pre[((size_t)g * NBMAX + n) * UCTA + u] : 0.f;
      if (MODE == B200RNN_GRU) {
        const float r = sigm(gi[i][0] + a[0]);
        const float z = sigm(gi[i][1] + a[1]);
        const float hn = a[2] + bhn;
        const float nn = tanh_f(gi[i][2] + r * hn);
        hnew[i] = nn + z * (h_prev[i] - nn);
        sv[i][0] = r; sv[i][1] = z; sv[i][2] = nn; sv[i][3] = 0.f; sx[i] = hn;
      } else {
        const float ig = sigm(gi[i][0] + a[0]);
        const float fg = sigm(gi[i][1] + a[1]);
        const float gg = tanh_f(gi[i][2] + a[2]);
        const float og = sigm(gi[i][G - 1] + a[G - 1]);
        const float cnew = fg * c_prev[i] + ig * gg;
        hnew[i] = og * tanh_f(cnew);
        c_prev[i] = cnew;
        sv[i][0] = ig; sv[i][1] = fg; sv[i][2] = gg; sv[i][3] = og; sx[i] = cnew;
      }
      h_prev[i] = hnew[i];
      h_sum[i] += hnew[i];
    }
    if (step + 1 < T) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int n = q + 4 * i;
        if (n >= NB) break;
        // 4 consecutive units -> one 16-byte st.async per destination CTA; lane&3 picks the destinations
        float4 v;
        v.x = __shfl_sync(FULLMASK, hnew[i], (lane & ~3) + 0);
        v.y = __shfl_sync(FULLMASK, hnew[i], (lane & ~3) + 1);
        v.z = __shfl_sync(FULLMASK, hnew[i], (lane & ~3) + 2);
        v.w = __shfl_sync(FULLMASK, hnew[i], (lane & ~3) + 3);
        const uint32_t dst = ptx::smem_u32(recv + (size_t)nxt * NBMAX * H + n * H + j0 + (lane & ~3));
        const uint32_t bar = ptx::smem_u32(&bars[nxt * 8 + (int)rank]);
#pragma unroll
        for (int r = (lane & 3); r < C; r += 4)
          ptx::st_async_v4(ptx::mapa(dst, (uint32_t)r), v, ptx::mapa(bar, (uint32_t)r));
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = q + 4 * i;
      if (n >= NB) break;
      const int b = b0 + n;
      if (b < B) {
        if (p.y) p.y[(long long)t * p.y_st + (long long)b * p.y_sb + dir * H + j] = hnew[i];
        if (p.training) {
          float* gp = gates + ((size_t)t * B + b) * GH + j;
          gp[0] = sv[i][0];
          gp[H] = sv[i][1];
          gp[2 * H] = sv[i][2];
          if (G == 4) gp[3 * H] = sv[i][3];
          extra[((size_t)t * B + b) * H + j] = sx[i];
        }
        if (step == T - 1) {
          p.h_n[((size_t)dir * B + b) * H + j] = hnew[i];
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
  ptx::cluster_sync_all();  // nobody exits while a peer could still address its shared memory
}

template <typename K>
int tc_prepare(K kernel, size_t smem) {
  static std::mutex mu;
  static const void* done[64];
  static int done_dev[64];
  static int ndone = 0;
  const int dev = current_device();
  std::lock_guard<std::mutex> lk(mu);
  for (int i = 0; i < ndone; ++i)
    if (done[i] == (const void*)kernel && done_dev[i] == dev) return B200RNN_OK;
  B200_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (ndone < 64) {
    done[ndone] = (const void*)kernel;
    done_dev[ndone++] = dev;
  }
  return B200RNN_OK;
}

template <typename K>
int tc_capacity(K kernel, int C, size_t smem) {
  static std::mutex mu;
  static const void* keys[64];
  static int vals[64], devs[64];
  static int n = 0;
  const int dev = current_device();
  {
    std::lock_guard<std::mutex> lk(mu);
    for (int i = 0; i < n; ++i)
      if (keys[i] == (const void*)kernel && devs[i] == dev) return vals[i];
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
  if (n < 64) {
    keys[n] = (const void*)kernel;
    devs[n] = dev;
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
