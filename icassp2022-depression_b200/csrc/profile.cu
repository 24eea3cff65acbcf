// profile.cu — event-pair pool behind b200rnn_profile / b200rnn_profile_read (include/b200rnn.h).
#include <mutex>
#include <vector>

#include "common.cuh"
#include "profile.cuh"

namespace b200rnn {

namespace {
struct Slot {
  cudaEvent_t a, b;
  int kind;
};
std::mutex g_mu;
bool g_on = false;
std::vector<Slot> g_slots;   // recorded pairs since the last read
std::vector<Slot> g_free;    // recycled event pairs
constexpr size_t MAX_SLOTS = 1 << 16;
}  // namespace

ProfScope::ProfScope(int kind, cudaStream_t s) : slot(-1), stream(s) {
  if (!g_on) return;
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_on || g_slots.size() >= MAX_SLOTS) return;
  Slot sl;
  if (!g_free.empty()) {
    sl = g_free.back();
    g_free.pop_back();
  } else {
    if (cudaEventCreate(&sl.a) != cudaSuccess || cudaEventCreate(&sl.b) != cudaSuccess) {
      cudaGetLastError();
      return;
    }
  }
  sl.kind = kind;
  cudaEventRecord(sl.a, s);
  g_slots.push_back(sl);
  slot = (int)g_slots.size() - 1;
}

ProfScope::~ProfScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_mu);
  if (slot < (int)g_slots.size()) cudaEventRecord(g_slots[slot].b, stream);
}

}  // namespace b200rnn

using namespace b200rnn;

extern "C" {

B200RNN_API int b200rnn_profile(int enable) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& s : g_slots) g_free.push_back(s);
  g_slots.clear();
  g_on = enable != 0;
  return B200RNN_OK;
}

B200RNN_API int b200rnn_profile_read(int kind, float* total_ms, int* launches) {
  std::lock_guard<std::mutex> lk(g_mu);
  float tot = 0.f;
  int n = 0;
  for (auto& s : g_slots) {
    if (s.kind != kind) continue;
    if (cudaEventSynchronize(s.b) != cudaSuccess) {
      set_error("profile_read: %s", cudaGetErrorString(cudaGetLastError()));
      return B200RNN_ERR_CUDA;
    }
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) {
      tot += ms;
      ++n;
    } else {
      cudaGetLastError();
    }
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = n;
  return B200RNN_OK;
}

}  // extern "C"
