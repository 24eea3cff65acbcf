// rnn_core.cuh — the per-step contraction shared by the persistent forward and backward kernels.
//
// One warp owns NR*UPW rows of a shared-memory resident weight slice (row-major, length KLEN) and
// multiplies them with BS vectors of length KLEN that also sit in shared memory. The contraction
// dimension is spread across the 32 lanes (each lane owns 4 consecutive k per 128-wide chunk, so all
// shared-memory reads are conflict-free 16-byte loads), and the cross-lane sum is a transposing
// butterfly: every stage halves the number of live partial sums per lane, so the whole reduction costs
// about one SHFL+FADD per accumulator instead of five.
//
// To keep every register index static, the *data* a register slot holds is permuted per lane:
//   slot (au, ab) of lane L accumulates   unit  = au ^ p(L),  batch = ab ^ q(L)
// with p = the top log2(UPW) lane bits and q = the next log2(BS) lane bits. Row / vector addresses are
// computed per lane (free: rows differ by multiples of KLEN floats, i.e. the same banks), and each
// butterfly stage becomes "slot[i] += shfl_xor(slot[i + half])" with compile-time i.
#pragma once
#include "common.cuh"

namespace b200rnn {

template <int V>
struct Log2 {
  static constexpr int value = 1 + Log2<V / 2>::value;
};
template <>
struct Log2<1> {
  static constexpr int value = 0;
};

template <int UPW, int BS>
struct LaneMap {
  static constexpr int LU = Log2<UPW>::value;
  static constexpr int LB = Log2<BS>::value;
  static_assert((1 << LU) == UPW && (1 << LB) == BS, "UPW and BS must be powers of two");
  static_assert(LU + LB <= 5, "UPW*BS must be <= 32");
  static constexpr int NREP = 1 << (5 - LU - LB);  // lanes holding the same (unit,batch) result
  __device__ static __forceinline__ int p(int lane) { return lane >> (5 - LU); }
  __device__ static __forceinline__ int q(int lane) { return (lane >> (5 - LU - LB)) & (BS - 1); }
  __device__ static __forceinline__ int rep(int lane) { return lane & (NREP - 1); }
};

// acc[r][au][ab] += sum_k W[row(r, au^p)][k] * vec[ab^q][k]   over this lane's k (partial sums)
//   W_s   : weight slice, row-major [.][KLEN]
//   row0  : first row of this warp inside group r is  r*group_stride + row0 + unit
//   vec_s : [BS][KLEN]
template <int NR, int UPW, int BS, int KLEN>
__device__ __forceinline__ void warp_partial_dots(const float* __restrict__ W_s, int group_stride, int row0,
                                                  const float* __restrict__ vec_s, int lane,
                                                  float (&acc)[NR][UPW][BS]) {
  using LM = LaneMap<UPW, BS>;
  static_assert(KLEN % 128 == 0, "contraction length must be a multiple of 128");
  const int p = LM::p(lane), q = LM::q(lane);
#pragma unroll
  for (int r = 0; r < NR; ++r)
#pragma unroll
    for (int au = 0; au < UPW; ++au)
#pragma unroll
      for (int ab = 0; ab < BS; ++ab) acc[r][au][ab] = 0.f;

#pragma unroll
  for (int i = 0; i < KLEN / 128; ++i) {
    const int koff = i * 128 + lane * 4;
    float4 hv[BS];
#pragma unroll
    for (int ab = 0; ab < BS; ++ab)
      hv[ab] = *reinterpret_cast<const float4*>(&vec_s[(ab ^ q) * KLEN + koff]);
#pragma unroll
    for (int r = 0; r < NR; ++r) {
#pragma unroll
      for (int au = 0; au < UPW; ++au) {
        const int row = r * group_stride + row0 + (au ^ p);
        const float4 wv = *reinterpret_cast<const float4*>(&W_s[row * KLEN + koff]);
#pragma unroll
        for (int ab = 0; ab < BS; ++ab) {
          float a = acc[r][au][ab];
          a = fmaf(wv.x, hv[ab].x, a);
          a = fmaf(wv.y, hv[ab].y, a);
          a = fmaf(wv.z, hv[ab].z, a);
          a = fmaf(wv.w, hv[ab].w, a);
          acc[r][au][ab] = a;
        }
      }
    }
  }
}

// Transposing butterfly. On return acc[r][0][0] of lane L holds the full sum for
// unit p(L), batch q(L) (replicated over the NREP low lanes).
template <int NR, int UPW, int BS>
__device__ __forceinline__ void warp_transpose_reduce(float (&acc)[NR][UPW][BS]) {
  constexpr unsigned FULL = 0xffffffffu;
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    int off = 16;
#pragma unroll
    for (int s = UPW / 2; s >= 1; s >>= 1) {
#pragma unroll
      for (int au = 0; au < s; ++au)
#pragma unroll
        for (int ab = 0; ab < BS; ++ab) acc[r][au][ab] += __shfl_xor_sync(FULL, acc[r][au + s][ab], off);
      off >>= 1;
    }
#pragma unroll
    for (int s = BS / 2; s >= 1; s >>= 1) {
#pragma unroll
      for (int ab = 0; ab < s; ++ab) acc[r][0][ab] += __shfl_xor_sync(FULL, acc[r][0][ab + s], off);
      off >>= 1;
    }
#pragma unroll
    for (; off >= 1; off >>= 1) acc[r][0][0] += __shfl_xor_sync(FULL, acc[r][0][0], off);
  }
}

}  // namespace b200rnn
