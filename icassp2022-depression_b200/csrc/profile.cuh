// profile.cuh — optional per-kernel-class device timing (CUDA events on the launching stream).
// Off by default; bench.py switches it on for its roofline pass. Not used under CUDA-graph capture.
#pragma once
#include <cuda_runtime.h>

namespace b200rnn {

enum ProfKind { PROF_REC_FWD = 0, PROF_REC_BWD = 1, PROF_GEMM = 2, PROF_MISC = 3, PROF_NKINDS = 4 };

struct ProfScope {
  ProfScope(int kind, cudaStream_t s);
  ~ProfScope();
  int slot;
  cudaStream_t stream;
};

}  // namespace b200rnn
