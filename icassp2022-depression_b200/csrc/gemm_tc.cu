// gemm_tc.cu — tensor-core path of the time-parallel input projection (K1):
//     C[M,N] = A[M,K] * W[N,K]^T + bias          fp32 in / fp32 out, 3xTF32 on tcgen05
//
// The hidden x input gate contraction over ALL time steps at once is the one genuinely dense GEMM of the path
// (M = B*T up to 15360, N = G*H, K = I), so it goes to the 5th-generation tensor cores — but the reference
// arithmetic is fp32 (torch rnn.py:1221-1224 / :842-847) and parity is judged at 1e-5, which plain TF32
// (10-bit mantissa) cannot hold. Each operand is therefore split x = hi + lo with hi = rna_tf32(x),
// lo = x - hi (exact), and three MMAs accumulate hi*hi + lo*hi + hi*lo in the fp32 TMEM accumulator
// (the dropped lo*lo term is ~2^-22 relative).
// The tensor core adds into the fp32 TMEM accumulator with truncation (round toward zero), a bias of about half an
// ulp per MMA that grows linearly with the number of MMAs chained on one accumulator (measured: 1.5e-5 abs at
// K=1024 with a single accumulator). So the chain is cut: k-blocks go round-robin to NMAIN=3 accumulators for
// hi*hi, the small cross terms get their own, and the epilogue adds the four with ordinary round-to-nearest fp32.
// Resulting error ~1e-6 relative to the largest output, same order as fp32 FFMA.
//
// Kernel anatomy (one 128x128 output tile per CTA, 256 threads):
//   warp 0   : TMA producer  — cp.async.bulk.tensor 2-D tiles (128 rows x 32 fp32 = 128-byte swizzled rows) of
//              A_hi, A_lo, W_hi, W_lo into a 3-stage shared-memory ring, mbarrier complete_tx
//   warp 1   : MMA issuer    — one elected lane issues tcgen05.mma.cta_group::1.kind::tf32 (M128 N128 K8), 12 per
//              stage, accumulator in TMEM; tcgen05.commit releases the stage / signals the epilogue
//   warp 2   : TMEM allocator (128 columns) / deallocator
//   warps 4-7: epilogue      — tcgen05.ld 32x32b.x16 -> registers, + folded biases, coalesced-per-row stores
#include <cuda.h>  // CUtensorMap types only; the encoder is fetched through cudaGetDriverEntryPoint
#include <mutex>

#include "gemm_f32.cuh"
#include "profile.cuh"
#include "ptx.cuh"

namespace b200rnn {

namespace {

constexpr int BM = 128, BN = 128, BK = 32, STAGES = 3;
constexpr int LNB_BLOCKS = 148 * 2;  // CTAs of the LayerNorm backward (per-CTA column partials, reduced in fixed order)
constexpr int TILE_BYTES = BM * BK * 4;        // 16 KB, both A and W tiles (BM == BN)
constexpr int STAGE_BYTES = 4 * TILE_BYTES;    // A_hi, A_lo, W_hi, W_lo
constexpr int TC_THREADS = 256;
constexpr int TC_SMEM = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 128 /*barriers*/ + 2 * BN * 4 /*bias*/ + 64;
constexpr int NMAIN = 3;                       // hi*hi accumulators (k-blocks round-robin), + 1 for the cross terms
constexpr int TMEM_COLS = (NMAIN + 1) * BN;    // 512: the whole TMEM of the SM

// ---- PTX wrappers specific to this kernel -----------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          ptx::smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   ptx::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
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
__device__ __forceinline__ void tc_ld_x16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor: K-major tile, 128-byte swizzle, rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address   [0,14)
  d |= (uint64_t)1 << 16;                             // leading byte offset (unused for swizzled K-major) = 1
  d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset: 8 rows * 128 B
  d |= (uint64_t)1 << 46;                             // descriptor version 1 (sm_100)
  d |= (uint64_t)2 << 61;                             // layout type: SWIZZLE_128B
  return d;
}

// MN-major tile (the contraction index is the ROW index of the source matrix, e.g. dG [T*B, G*H] as the A operand of
// dW = dG^T X). For 32-bit operands the only MN-major shared-memory layout tcgen05 accepts is "128-byte swizzle with
// 32-byte atoms" (descriptor layout type 1; CUTLASS: Layout_MN_SW128_32B_Atom = Swizzle<2,5,2> on byte addresses):
// rows of 32 floats of M/N (128 B), FOUR K rows per 512-byte atom, inside an atom the 32-byte chunks of a row are
// XOR-ed with (row & 3) - exactly what a TMA box {32 floats, 32 rows} with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
// deposits. Atoms along K are 512 B apart (SBO), the four 32-wide column blocks of a 128-wide tile 4 KB apart (LBO).
// One MMA (K = 8) consumes two atoms along K; the next K step starts 1 KB further.
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(4096 >> 4) << 16;  // leading byte offset: next 32-wide block along M/N
  d |= (uint64_t)(512 >> 4) << 32;   // stride byte offset: next 4 K rows
  d |= (uint64_t)1 << 46;            // descriptor version 1 (sm_100)
  d |= (uint64_t)1 << 61;            // layout type: SWIZZLE_128B_BASE32B
  return d;
}

// instruction descriptor: D=f32, A=B=tf32, both K-major, M=128, N=128, dense, no negate
constexpr uint32_t IDESC_TF32_128x128 =
    (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

struct TcArgs {
  float* C;
  RowMap c_rows;
  int M, N, K;
  const float* bias1;
  const float* bias2;
  int bias2_n;
  int tiles_m, tiles_n;
  int nmain;  // hi*hi accumulators in use (1..NMAIN): k-blocks go round-robin over them
  int nsets;  // TMEM accumulator sets (2 when (nmain+1)*BN*2 <= 512: epilogue of tile i overlaps mainloop of i+1)
  int accumulate;   // C += result (splitk == 1 only)
  int a_mn, b_mn;   // operand is MN-major: source matrix [K rows][M or N contiguous] (wgrad / dgrad without transposes)
  int splitk;       // > 1: work item = (k-split, tile); raw partial sums go to partial[ks][M][N]
  int kb_per_split; // k-blocks per split
  float* partial;
};

// Persistent: each CTA walks tiles blockIdx.x, +gridDim.x, ... (m fastest, so consecutive tiles of a CTA mostly
// share their W tile rows in L2). Warp roles loop independently over the same tile sequence and meet only through
// mbarriers: smem ring full/empty (TMA <-> MMA), TMEM set full/empty (MMA <-> epilogue).
__global__ void __launch_bounds__(TC_THREADS, 1)
    gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                       const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                       const TcArgs args) {
  extern __shared__ unsigned char smem_raw[];
  // aligned by offset so that the compiler keeps the shared state space (LDS/STS instead of generic LD/ST)
  unsigned char* base = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(base + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;   // [2]
  uint64_t* tempty = tfull + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);  // [2][BN] folded bias of the tile, per TMEM set

  // canonical warp index: the shuffle makes it warp-uniform for the compiler, so the single-thread TMA / MMA issue
  // loops below stay on the uniform datapath (no per-instruction R2UR broadcast loop around UTMALDG / UTCHMMA)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int nkb_total = (args.K + BK - 1) / BK;  // TMA zero-fills the K tail
  const int ntiles = args.tiles_m * args.tiles_n;
  const int nitems = ntiles * args.splitk;
  const int set_cols = (args.nmain + 1) * BN;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tfull[s], 1);
      ptx::mbar_init(&tempty[s], 4);  // one arrival per epilogue warp
    }
    ptx::fence_mbar_init();
    prefetch_tmap(&map_a_hi);
    prefetch_tmap(&map_a_lo);
    prefetch_tmap(&map_b_hi);
    prefetch_tmap(&map_b_lo);
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(tmem_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    if (ptx::elect_one_sync()) {
      int it = 0;  // running k-block counter across tiles (ring position)
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int tile = item % ntiles, ks = item / ntiles;
        const int m0 = (tile % args.tiles_m) * BM, n0 = (tile / args.tiles_m) * BN;
        const int kb0 = ks * args.kb_per_split;
        const int kb1 = min(nkb_total, kb0 + args.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          ptx::mbar_wait(&empty[s], ph ^ 1);
          unsigned char* st = base + s * STAGE_BYTES;
          ptx::mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
          if (args.a_mn) {  // four {32 floats of M, BK rows of K} boxes per split half
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              tma_load_2d(st + 0 * TILE_BYTES + j * 4096, &map_a_hi, m0 + 32 * j, kb * BK, &full[s]);
              tma_load_2d(st + 1 * TILE_BYTES + j * 4096, &map_a_lo, m0 + 32 * j, kb * BK, &full[s]);
            }
          } else {
            tma_load_2d(st + 0 * TILE_BYTES, &map_a_hi, kb * BK, m0, &full[s]);
            tma_load_2d(st + 1 * TILE_BYTES, &map_a_lo, kb * BK, m0, &full[s]);
          }
          if (args.b_mn) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              tma_load_2d(st + 2 * TILE_BYTES + j * 4096, &map_b_hi, n0 + 32 * j, kb * BK, &full[s]);
              tma_load_2d(st + 3 * TILE_BYTES + j * 4096, &map_b_lo, n0 + 32 * j, kb * BK, &full[s]);
            }
          } else {
            tma_load_2d(st + 2 * TILE_BYTES, &map_b_hi, kb * BK, n0, &full[s]);
            tma_load_2d(st + 3 * TILE_BYTES, &map_b_lo, kb * BK, n0, &full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one_sync()) {
      int it = 0, ti = 0;
      // instruction descriptor with the operands' major-ness (bit 15: A is MN-major, bit 16: B is MN-major)
      const uint32_t idesc = IDESC_TF32_128x128 | (args.a_mn ? (1u << 15) : 0u) | (args.b_mn ? (1u << 16) : 0u);
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++ti) {
        const int nkb = min(nkb_total, (item / ntiles + 1) * args.kb_per_split) - (item / ntiles) * args.kb_per_split;
        const int set = ti % args.nsets;
        const uint32_t use = (uint32_t)(ti / args.nsets);  // how often this set has been used before
        ptx::mbar_wait(&tempty[set], (use & 1) ^ 1);       // the epilogue has drained this set
        tc_fence_after();
        const uint32_t acc_set = tmem_base + (uint32_t)(set * set_cols);
        const uint32_t acc_cross = acc_set + (uint32_t)(args.nmain * BN);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          ptx::mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t st = ptx::smem_u32(base + s * STAGE_BYTES);
          const uint64_t a_hi = args.a_mn ? make_mnmajor_sw128_desc(st + 0 * TILE_BYTES) : make_kmajor_sw128_desc(st + 0 * TILE_BYTES);
          const uint64_t a_lo = args.a_mn ? make_mnmajor_sw128_desc(st + 1 * TILE_BYTES) : make_kmajor_sw128_desc(st + 1 * TILE_BYTES);
          const uint64_t b_hi = args.b_mn ? make_mnmajor_sw128_desc(st + 2 * TILE_BYTES) : make_kmajor_sw128_desc(st + 2 * TILE_BYTES);
          const uint64_t b_lo = args.b_mn ? make_mnmajor_sw128_desc(st + 3 * TILE_BYTES) : make_kmajor_sw128_desc(st + 3 * TILE_BYTES);
          const uint32_t acc_main = acc_set + (uint32_t)((kb % args.nmain) * BN);
          // K step of 8: K-major = 32 bytes further inside the swizzle atom; MN-major = the next 1 KB atom
          const uint32_t a_step = args.a_mn ? (1024u >> 4) : (32u >> 4), b_step = args.b_mn ? (1024u >> 4) : (32u >> 4);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint64_t adva = (uint64_t)(k * a_step), advb = (uint64_t)(k * b_step);
            tc_mma_tf32(acc_cross, a_lo + adva, b_hi + advb, idesc, (kb | k) != 0 ? 1u : 0u);
            tc_mma_tf32(acc_cross, a_hi + adva, b_lo + advb, idesc, 1u);
            tc_mma_tf32(acc_main, a_hi + adva, b_hi + advb, idesc, (kb >= args.nmain || k != 0) ? 1u : 0u);
          }
          tc_commit(&empty[s]);  // implies tcgen05.fence::before_thread_sync
        }
        tc_commit(&tfull[set]);
      }
    }
  } else if (warp >= 4) {
    const int wq = warp & 3;  // TMEM lane quarter this warp may read
    const int et = threadIdx.x - 128;  // 0..127
    int ti = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++ti) {
      const int tile = item % ntiles, ks = item / ntiles;
      const int nkb = min(nkb_total, (ks + 1) * args.kb_per_split) - ks * args.kb_per_split;
      const int m0 = (tile % args.tiles_m) * BM, n0 = (tile / args.tiles_m) * BN;
      const int set = ti % args.nsets;
      const uint32_t use = (uint32_t)(ti / args.nsets);
      // folded bias of this tile's columns -> smem (issued before waiting for the accumulator)
      {
        const int n = n0 + et;
        float bv = 0.f;
        if (args.splitk == 1) {
          if (args.bias1) bv += __ldg(args.bias1 + n);
          if (args.bias2 && n < args.bias2_n) bv += __ldg(args.bias2 + n);
        }
        bias_s[set * BN + et] = bv;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // epilogue warps only
      ptx::mbar_wait(&tfull[set], use & 1);
      tc_fence_after();
      const int row = m0 + wq * 32 + lane;
      float* crow = (args.splitk > 1)
                        ? args.partial + ((size_t)ks * args.M + (row < args.M ? row : 0)) * args.N + n0
                        : args.C + args.c_rows.off(row < args.M ? row : 0) + n0;
      const uint32_t lane_base = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(set * set_cols);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 16) {
        float v[16];
        tc_ld_x16(lane_base + (uint32_t)(args.nmain * BN + c0), v);  // cross terms
        for (int a = 0; a < args.nmain; ++a) {
          if (a < nkb) {  // accumulator a was written (uniform condition)
            float m[16];
            tc_ld_x16(lane_base + (uint32_t)(a * BN + c0), m);
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] += m[e];
          }
        }
        if (row < args.M) {
          const float4* bs = reinterpret_cast<const float4*>(&bias_s[set * BN + c0]);
          float4* dst = reinterpret_cast<float4*>(crow + c0);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 bq = bs[q];
            float4 o = make_float4(v[4 * q + 0] + bq.x, v[4 * q + 1] + bq.y, v[4 * q + 2] + bq.z, v[4 * q + 3] + bq.w);
            if (args.accumulate && args.splitk == 1) {
              const float4 old = dst[q];
              o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
            }
            dst[q] = o;
          }
        }
      }
      // this warp is done reading the set: release it to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ptx::smem_u32(&tempty[set])) : "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// x = hi + lo with hi = round-to-nearest TF32 (kept in a 32-bit container), lo = x - hi (exact in fp32).
// Reads rows through a RowMap (batch_first / permuted inputs), writes two dense [M,K] matrices.
__global__ void split_tf32_kernel(const float* __restrict__ src, RowMap rows, int M, int K, float* __restrict__ hi,
                                  float* __restrict__ lo, int vec_ok) {
  const size_t nvec = (size_t)M * (K / 4);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / (K / 4));
    const int k = (int)(i - (size_t)m * (K / 4)) * 4;
    const float* p = src + rows.off(m) + k;
    float4 x;
    if (vec_ok) {
      x = __ldg(reinterpret_cast<const float4*>(p));
    } else {
      x = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
    }
    float4 h, l;
    uint32_t t;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(x.x)); h.x = __uint_as_float(t); l.x = x.x - h.x;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(x.y)); h.y = __uint_as_float(t); l.y = x.y - h.y;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(x.z)); h.z = __uint_as_float(t); l.z = x.z - h.z;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(x.w)); h.w = __uint_as_float(t); l.w = x.w - h.w;
    *reinterpret_cast<float4*>(hi + (size_t)m * K + k) = h;
    *reinterpret_cast<float4*>(lo + (size_t)m * K + k) = l;
  }
}

// LayerNorm(row) * gamma + beta, then the TF32 split — the prologue of audio_gru_whole.py:104 / fuse_net_whole.py:360
// folded into the operand preparation of K1. One warp per row; Cc % 128 == 0, Cc <= 1024.
template <int NV>  // float4 per lane
__global__ void layernorm_split_kernel(const float* __restrict__ src, RowMap rows, int R, int Cc,
                                       const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                       float* __restrict__ hi, float* __restrict__ lo, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < R; r += gridDim.x * wpb) {
    const float* p = src + rows.off(r);
    float4 x[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      x[i] = __ldg(reinterpret_cast<const float4*>(p + i * 128 + lane * 4));
      s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)Cc;
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float a = x[i].x - mean, b = x[i].y - mean, c = x[i].z - mean, d = x[i].w - mean;
      v += (a * a + b * b) + (c * c + d * d);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = rsqrtf(v / (float)Cc + eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int k = i * 128 + lane * 4;
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + k));
      const float4 bt = __ldg(reinterpret_cast<const float4*>(beta + k));
      float y[4] = {(x[i].x - mean) * rstd * g.x + bt.x, (x[i].y - mean) * rstd * g.y + bt.y,
                    (x[i].z - mean) * rstd * g.z + bt.z, (x[i].w - mean) * rstd * g.w + bt.w};
      float h[4], l[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        uint32_t t;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(y[e]));
        h[e] = __uint_as_float(t);
        l[e] = y[e] - h[e];
      }
      *reinterpret_cast<float4*>(hi + (size_t)r * Cc + k) = make_float4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<float4*>(lo + (size_t)r * Cc + k) = make_float4(l[0], l[1], l[2], l[3]);
      if (out) *reinterpret_cast<float4*>(out + (size_t)r * Cc + k) = make_float4(y[0], y[1], y[2], y[3]);  // for backward
    }
  }
}

// LayerNorm backward for the folded prologue: dy = gradient w.r.t. LN(x) (dense [R][Cc], produced by the layer-0 dgrad
// GEMM), x read through the caller's row map, statistics recomputed (cheaper than saving them):
//   xhat = (x - mean) rstd;  g = dy * gamma;  dx = rstd (g - mean(g) - xhat mean(g xhat))
//   dgamma += sum_rows dy xhat;  dbeta += sum_rows dy      (per-CTA partials here, fixed-order reduce below)
// One warp per row, each lane keeps the column partials of its 4*NV columns in registers across its rows.
template <int NV>
__global__ void layernorm_bwd_kernel(const float* __restrict__ x, RowMap x_rows, const float* __restrict__ dy, int R,
                                     int Cc, const float* __restrict__ gamma, float eps, float* __restrict__ dx,
                                     RowMap dx_rows, float* __restrict__ part /* [grid][2][Cc] */) {
  extern __shared__ float red[];  // [warps][2][Cc]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  float4 dg[NV], db[NV], gm[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    gm[i] = __ldg(reinterpret_cast<const float4*>(gamma + i * 128 + lane * 4));
  }
  for (int r = blockIdx.x * wpb + warp; r < R; r += gridDim.x * wpb) {
    const float* px = x + x_rows.off(r);
    const float* pd = dy + (size_t)r * Cc;
    float4 xv[NV], dv[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      xv[i] = __ldg(reinterpret_cast<const float4*>(px + i * 128 + lane * 4));
      dv[i] = *reinterpret_cast<const float4*>(pd + i * 128 + lane * 4);
      s += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)Cc;
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      xv[i].x -= mean; xv[i].y -= mean; xv[i].z -= mean; xv[i].w -= mean;
      v += (xv[i].x * xv[i].x + xv[i].y * xv[i].y) + (xv[i].z * xv[i].z + xv[i].w * xv[i].w);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = rsqrtf(v / (float)Cc + eps);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      xv[i].x *= rstd; xv[i].y *= rstd; xv[i].z *= rstd; xv[i].w *= rstd;  // xhat
      dg[i].x += dv[i].x * xv[i].x; dg[i].y += dv[i].y * xv[i].y; dg[i].z += dv[i].z * xv[i].z; dg[i].w += dv[i].w * xv[i].w;
      db[i].x += dv[i].x; db[i].y += dv[i].y; db[i].z += dv[i].z; db[i].w += dv[i].w;
      dv[i].x *= gm[i].x; dv[i].y *= gm[i].y; dv[i].z *= gm[i].z; dv[i].w *= gm[i].w;  // g = dy * gamma
      sg += (dv[i].x + dv[i].y) + (dv[i].z + dv[i].w);
      sgx += (dv[i].x * xv[i].x + dv[i].y * xv[i].y) + (dv[i].z * xv[i].z + dv[i].w * xv[i].w);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      sg += __shfl_xor_sync(0xffffffffu, sg, o);
      sgx += __shfl_xor_sync(0xffffffffu, sgx, o);
    }
    const float mg = sg / (float)Cc, mgx = sgx / (float)Cc;
    if (dx) {
      float* po = dx + dx_rows.off(r);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float4 o4;
        o4.x = rstd * (dv[i].x - mg - xv[i].x * mgx);
        o4.y = rstd * (dv[i].y - mg - xv[i].y * mgx);
        o4.z = rstd * (dv[i].z - mg - xv[i].z * mgx);
        o4.w = rstd * (dv[i].w - mg - xv[i].w * mgx);
        *reinterpret_cast<float4*>(po + i * 128 + lane * 4) = o4;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    *reinterpret_cast<float4*>(red + ((size_t)warp * 2 + 0) * Cc + i * 128 + lane * 4) = dg[i];
    *reinterpret_cast<float4*>(red + ((size_t)warp * 2 + 1) * Cc + i * 128 + lane * 4) = db[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * Cc; c += blockDim.x) {
    float a = 0.f;
    for (int w2 = 0; w2 < wpb; ++w2) a += red[(size_t)w2 * 2 * Cc + c];  // fixed order
    part[(size_t)blockIdx.x * 2 * Cc + c] = a;
  }
}
__global__ void layernorm_bwd_reduce_kernel(const float* __restrict__ part, int nparts, int Cc, float* __restrict__ dgamma,
                                            float* __restrict__ dbeta, int accumulate) {
  // one warp per column: lane l adds partials l, l+32, ... in order, then a fixed shuffle tree => deterministic, and
  // the ~300 partial rows are read 32 at a time instead of one after the other by a single thread
  const int lane = threadIdx.x & 31, c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= 2 * Cc) return;
  float a = 0.f;
  for (int p2 = lane; p2 < nparts; p2 += 32) a += part[(size_t)p2 * 2 * Cc + c];
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) {
    float* dst = c < Cc ? (dgamma ? dgamma + c : nullptr) : (dbeta ? dbeta + (c - Cc) : nullptr);
    if (dst) *dst = accumulate ? *dst + a : a;
  }
}

// ---- host side -------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encoder() {
  static std::mutex mu;
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  std::lock_guard<std::mutex> lk(mu);
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  }
  return fn;
}

// dense row-major [rows, K] fp32 matrix, box = [128 rows, 32 floats], 128-byte swizzle, OOB rows read as zero
bool make_map(CUtensorMap* map, const float* ptr, int rows, int K, long long ld = 0) {
  EncodeTiledFn enc = get_encoder();
  if (!enc) return false;
  if (ld == 0) ld = K;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// MN-major operand: dense row-major [Krows, MN] fp32 matrix (the contraction index is the row), box = [32 K rows,
// 32 floats of M/N], 128-byte swizzle; rows beyond Krows and columns beyond MN read as zero
bool make_map_mn(CUtensorMap* map, const float* ptr, int Krows, int MN, long long ld) {
  EncodeTiledFn enc = get_encoder();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)MN, (cuuint64_t)Krows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {32u, (cuuint32_t)BK};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

bool tc_disabled() {
  static const bool off = getenv("B200RNN_NO_TC") != nullptr;
  return off;
}

}  // namespace

size_t gemm_tc_scratch_bytes(int M, int N, int K) { return (size_t)2 * ((size_t)M + N) * K * sizeof(float) + 1024; }

bool gemm_tc_eligible(const GemmParams& p, size_t ws_bytes) {
  if (tc_disabled()) return false;
  if (p.accumulate) return false;
  if (p.M < 1 || p.K < BK || p.K % BK != 0 || p.N % BN != 0) return false;
  if (!p.a_kcontig && p.M % 4 != 0) return false;  // MN-major split copies are dense [K][M]: rows must stay 16-byte aligned
  if (ws_bytes < gemm_tc_scratch_bytes(p.M, p.N, p.K)) return false;
  // the epilogue stores float4 along n
  if ((reinterpret_cast<uintptr_t>(p.C) & 15u) || (p.c_rows.s_outer % 4) || (p.c_rows.s_inner % 4)) return false;
  return get_encoder() != nullptr;
}

int tc_split(const float* src, const RowMap& rows, int R, int Cc, float* hi, float* lo, cudaStream_t stream) {
  const bool vec = (reinterpret_cast<uintptr_t>(src) & 15u) == 0 && rows.s_outer % 4 == 0 && rows.s_inner % 4 == 0;
  size_t nv = (size_t)R * (Cc / 4);
  int blocks = (int)((nv + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  ProfScope prof(PROF_MISC, stream);
  split_tf32_kernel<<<blocks, 256, 0, stream>>>(src, rows, R, Cc, hi, lo, vec ? 1 : 0);
  B200_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return B200RNN_OK;
}

bool tc_available() { return !tc_disabled() && get_encoder() != nullptr; }

float* tc_a_hi(void* ws) { return reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255); }
float* tc_a_lo(void* ws, int M, int K) { return tc_a_hi(ws) + (size_t)M * K; }

int tc_layernorm_split(const float* src, const RowMap& rows, int R, int Cc, const float* gamma, const float* beta,
                       float eps, float* hi, float* lo, cudaStream_t stream, float* out) {
  const bool vec = (reinterpret_cast<uintptr_t>(src) & 15u) == 0 && rows.s_outer % 4 == 0 && rows.s_inner % 4 == 0 &&
                   (reinterpret_cast<uintptr_t>(gamma) & 15u) == 0 && (reinterpret_cast<uintptr_t>(beta) & 15u) == 0;
  if (!vec || !(Cc == 128 || Cc == 256 || Cc == 512 || Cc == 1024)) {
    set_error("layernorm_split: needs 16-byte aligned rows and a feature width of 128, 256, 512 or 1024");
    return B200RNN_ERR_UNSUPPORTED;
  }
  int blocks = (R + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  ProfScope prof(PROF_MISC, stream);
  switch (Cc / 128) {  // instantiated widths: 128, 256, 512, 1024 (the reference normalises 256-d audio features)
    case 1: layernorm_split_kernel<1><<<blocks, 256, 0, stream>>>(src, rows, R, Cc, gamma, beta, eps, hi, lo, out); break;
    case 2: layernorm_split_kernel<2><<<blocks, 256, 0, stream>>>(src, rows, R, Cc, gamma, beta, eps, hi, lo, out); break;
    case 4: layernorm_split_kernel<4><<<blocks, 256, 0, stream>>>(src, rows, R, Cc, gamma, beta, eps, hi, lo, out); break;
    default: layernorm_split_kernel<8><<<blocks, 256, 0, stream>>>(src, rows, R, Cc, gamma, beta, eps, hi, lo, out); break;
  }
  B200_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return B200RNN_OK;
}

size_t layernorm_bwd_scratch_floats(int Cc) { return (size_t)LNB_BLOCKS * 2 * Cc; }

int launch_layernorm_bwd(const float* x, const RowMap& x_rows, const float* dy, int R, int Cc, const float* gamma,
                         float eps, float* dx, const RowMap& dx_rows, float* dgamma, float* dbeta, int accumulate,
                         float* part, cudaStream_t stream) {
  const bool vec = (reinterpret_cast<uintptr_t>(x) & 15u) == 0 && x_rows.s_outer % 4 == 0 && x_rows.s_inner % 4 == 0 &&
                   (reinterpret_cast<uintptr_t>(gamma) & 15u) == 0 && (reinterpret_cast<uintptr_t>(dy) & 15u) == 0 &&
                   (!dx || ((reinterpret_cast<uintptr_t>(dx) & 15u) == 0 && dx_rows.s_outer % 4 == 0 &&
                            dx_rows.s_inner % 4 == 0));
  if (!vec || !(Cc == 128 || Cc == 256 || Cc == 512 || Cc == 1024)) {
    set_error("layernorm_bwd: needs 16-byte aligned rows and a feature width of 128, 256, 512 or 1024");
    return B200RNN_ERR_UNSUPPORTED;
  }
  int blocks = (R + 7) / 8;
  if (blocks > LNB_BLOCKS) blocks = LNB_BLOCKS;
  const size_t smem = (size_t)8 * 2 * Cc * sizeof(float);
  ProfScope prof(PROF_MISC, stream);
  // 8 warps x 2 x Cc floats of per-warp column partials: 64 KB at Cc = 1024, above the 48 KB default limit
  static bool attr[MAX_DEVICES] = {false};
  if (!attr[current_device()]) {
    B200_CUDA_CHECK(cudaFuncSetAttribute(layernorm_bwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr[current_device()] = true;
  }
#define B200_LNB(NV_) \
  layernorm_bwd_kernel<NV_><<<blocks, 256, smem, stream>>>(x, x_rows, dy, R, Cc, gamma, eps, dx, dx_rows, part)
  switch (Cc / 128) {
    case 1: B200_LNB(1); break;
    case 2: B200_LNB(2); break;
    case 4: B200_LNB(4); break;
    default: B200_LNB(8); break;
  }
#undef B200_LNB
  B200_CUDA_CHECK(cudaGetLastError());
  layernorm_bwd_reduce_kernel<<<(2 * Cc + 7) / 8, 256, 0, stream>>>(part, blocks, Cc, dgamma, dbeta, accumulate);
  B200_CUDA_CHECK(cudaGetLastError());
  count_launch(2);
  return B200RNN_OK;
}

// C[M,N] (+)= A[M,K] * B[N,K]^T (+ biases), operands already split into hi/lo K-major matrices.
int tc_gemm_presplit(const TcOperand& A, const TcOperand& B, int M, int N, int K, float* C, const RowMap& c_rows,
                     const float* bias1, const float* bias2, int bias2_n, int accumulate, void* splitk_ws,
                     size_t splitk_ws_bytes, cudaStream_t stream) {
  if (M < 1 || N % BN != 0 || K < 1) {
    set_error("tc_gemm: unsupported shape M=%d N=%d K=%d", M, N, K);
    return B200RNN_ERR_UNSUPPORTED;
  }
  if ((reinterpret_cast<uintptr_t>(C) & 15u) || (c_rows.s_outer % 4) || (c_rows.s_inner % 4)) {
    set_error("tc_gemm: output must be 16-byte aligned with row strides that are multiples of 4 floats");
    return B200RNN_ERR_UNSUPPORTED;
  }
  CUtensorMap m_ahi, m_alo, m_bhi, m_blo;
  const bool ok_a = A.mn ? (make_map_mn(&m_ahi, A.hi, K, M, A.ld) && make_map_mn(&m_alo, A.lo, K, M, A.ld))
                         : (make_map(&m_ahi, A.hi, M, K, A.ld) && make_map(&m_alo, A.lo, M, K, A.ld));
  const bool ok_b = B.mn ? (make_map_mn(&m_bhi, B.hi, K, N, B.ld) && make_map_mn(&m_blo, B.lo, K, N, B.ld))
                         : (make_map(&m_bhi, B.hi, N, K, B.ld) && make_map(&m_blo, B.lo, N, K, B.ld));
  if (!ok_a || !ok_b) {
    set_error("tc_gemm: cuTensorMapEncodeTiled failed (operands must be 16-byte aligned, ld %% 4 == 0)");
    return B200RNN_ERR_CUDA;
  }
  static std::mutex mu;
  static bool attr_done[MAX_DEVICES] = {false};
  {
    const int dev = current_device();
    std::lock_guard<std::mutex> lk(mu);
    if (!attr_done[dev]) {
      B200_CUDA_CHECK(cudaFuncSetAttribute(gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
      attr_done[dev] = true;
    }
  }
  int sms = 148;
  {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  TcArgs a;
  a.C = C; a.c_rows = c_rows;
  a.M = M; a.N = N; a.K = K;
  a.bias1 = bias1; a.bias2 = bias2; a.bias2_n = bias2_n;
  a.accumulate = accumulate;
  a.a_mn = A.mn ? 1 : 0;
  a.b_mn = B.mn ? 1 : 0;
  a.tiles_m = (M + BM - 1) / BM;
  a.tiles_n = N / BN;
  const int ntiles = a.tiles_m * a.tiles_n;
  const int nkb = (K + BK - 1) / BK;
  // split-K when the tile count cannot fill the chip and K is long (the wgrad shapes: K = T*B)
  int splitk = 1;
  if (splitk_ws && ntiles * 2 <= sms && nkb >= 16) {
    splitk = sms / ntiles;
    if (splitk > nkb / 8) splitk = nkb / 8;
    const size_t per = (size_t)M * N * sizeof(float);
    while (splitk > 1 && per * splitk > splitk_ws_bytes) --splitk;
    if (splitk < 1) splitk = 1;
  }
  a.kb_per_split = (nkb + splitk - 1) / splitk;
  a.splitk = (nkb + a.kb_per_split - 1) / a.kb_per_split;
  a.partial = static_cast<float*>(splitk_ws);
  a.nmain = (a.kb_per_split + 11) / 12;  // <= 12 k-blocks (48 hi*hi MMAs) chained per accumulator
  if (a.nmain < 1) a.nmain = 1;
  if (a.nmain > NMAIN) a.nmain = NMAIN;
  a.nsets = ((a.nmain + 1) * BN * 2 <= TMEM_COLS) ? 2 : 1;
  const int nitems = ntiles * a.splitk;
  dim3 grid(nitems < sms ? nitems : sms, 1, 1);
  {
    ProfScope prof(PROF_GEMM, stream);
    gemm_tf32x3_kernel<<<grid, TC_THREADS, TC_SMEM, stream>>>(m_ahi, m_alo, m_bhi, m_blo, a);
    B200_CUDA_CHECK(cudaGetLastError());
    count_launch();
  }
  if (a.splitk > 1)
    return launch_splitk_reduce(a.partial, a.splitk, M, N, C, c_rows, bias1, bias2, bias2_n, accumulate, stream);
  return B200RNN_OK;
}

int launch_gemm_tc(const GemmParams& p, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (!gemm_tc_eligible(p, ws_bytes)) {
    set_error("gemm_tc: problem not eligible for the tcgen05 path");
    return B200RNN_ERR_UNSUPPORTED;
  }
  float* a_hi = tc_a_hi(ws);
  float* a_lo = tc_a_lo(ws, p.M, p.K);
  const float* b_hi = p.tc_b_hi;
  const float* b_lo = p.tc_b_lo;
  int rc = B200RNN_OK;
  // a_kcontig: A is [M rows][K]; else A is [K rows][M] (MN-major): the split copy keeps the source's orientation
  if (!p.tc_a_presplit)
    rc = p.a_kcontig ? tc_split(p.A, p.a_rows, p.M, p.K, a_hi, a_lo, stream)
                     : tc_split(p.A, p.a_rows, p.K, p.M, a_hi, a_lo, stream);
  if (rc) return rc;
  if (!b_hi || !b_lo) {
    float* w_hi = a_lo + (size_t)p.M * p.K;
    float* w_lo = w_hi + (size_t)p.N * p.K;
    rc = p.b_kcontig ? tc_split(p.B, p.b_rows, p.N, p.K, w_hi, w_lo, stream)
                     : tc_split(p.B, p.b_rows, p.K, p.N, w_hi, w_lo, stream);
    if (rc) return rc;
    b_hi = w_hi;
    b_lo = w_lo;
  }
  TcOperand A{a_hi, a_lo, p.a_kcontig ? p.K : p.M, !p.a_kcontig}, B{b_hi, b_lo, p.b_kcontig ? p.K : p.N, !p.b_kcontig};
  return tc_gemm_presplit(A, B, p.M, p.N, p.K, p.C, p.c_rows, p.bias1, p.bias2, p.bias2_n, 0, nullptr, 0, stream);
}

}  // namespace b200rnn
