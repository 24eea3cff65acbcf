// misc_kernels.cuh — small time-parallel helpers around the recurrence (K7 dropout, transposes, bias sums).
#pragma once
#include "common.cuh"

namespace b200rnn {

// Resolve the dropout RNG state of one forward call on the device, so the call is CUDA-graph replayable:
//   hdr[0] = seed, hdr[1] = offset   taken from `state_dev` ([seed, offset], then offset += consume) if it is
//   non-NULL, else from the by-value arguments.
int launch_rng_setup(uint64_t* hdr, uint64_t seed, uint64_t offset, uint64_t* state_dev, uint64_t consume,
                     cudaStream_t stream);

// out[i] = in[i] * mask(i) / (1-p) over n dense elements; mask is Philox4x32-10 keyed by hdr = {seed, offset}
// and the per-layer stream id; in == out is allowed (in place).
int launch_dropout(const float* in, float* out, size_t n, float p, const uint64_t* hdr, uint32_t stream_id,
                   cudaStream_t stream);

// Same, and additionally emits the TF32 hi/lo split of the dropped values (operand of the next layer's K1 GEMM),
// saving a separate pass over the layer output. n must be a multiple of 4, all pointers 16-byte aligned.
int launch_dropout_split(const float* in, float* out, float* hi, float* lo, size_t n, float p, const uint64_t* hdr,
                         uint32_t stream_id, cudaStream_t stream);

// db_ih / db_hh from the per-slice partial sums written by the backward recurrence:
//   part [nslices][(G+1)*H]  (first G*H: sum of dGi columns; tail H: GRU sum of dn*r)
//   GRU : db_ih = sum(part[:, :3H]);  db_hh = (sum part[:, :2H], sum part[:, 3H:4H])
//   LSTM: db_ih = db_hh = sum(part[:, :4H])
int launch_bias_reduce(const float* part, int nslices, int mode, int H, float* db_ih, float* db_hh,
                       int accumulate, cudaStream_t stream);

}  // namespace b200rnn
