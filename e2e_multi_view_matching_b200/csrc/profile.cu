// Launch counter + optional per-kernel-class CUDA-event profiler (used by bench.py to measure
// the dominant kernel's launch duration live, on the launching stream, inside real steps).
#include "../../include/mvm_b200.h"
#include "common.cuh"
#include <vector>

unsigned long long g_mvm_launches = 0;

namespace {
struct Rec { int tag; cudaEvent_t a, b; };
bool g_enabled = false;
std::vector<Rec> g_recs;
}  // namespace

MvmProfScope::MvmProfScope(int tag, cudaStream_t s) : tag_(tag), s_(s), idx_(-1) {
  if (!g_enabled) return;
  Rec r;
  r.tag = tag;
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  cudaEventRecord(r.a, s);
  idx_ = (int)g_recs.size();
  g_recs.push_back(r);
}
MvmProfScope::~MvmProfScope() {
  if (idx_ >= 0) cudaEventRecord(g_recs[idx_].b, s_);
}

// ---- per-device info / once-per-device function attributes (common.cuh) ------------------------
namespace {
constexpr int kMaxDev = 64;
std::mutex g_attr_mutex;
bool g_attr_done[kMaxDev][MVM_N_ONCE];
MvmDevInfo g_dev[kMaxDev];
bool g_dev_init[kMaxDev];
}  // namespace

std::mutex& mvm_attr_mutex() { return g_attr_mutex; }
bool* mvm_attr_flag(int slot) {
  int dev = 0;
  cudaGetDevice(&dev);
  return &g_attr_done[dev % kMaxDev][slot];
}
const MvmDevInfo& mvm_dev_info() {
  int dev = 0;
  cudaGetDevice(&dev);
  dev %= kMaxDev;
  static std::mutex mu;
  std::lock_guard<std::mutex> g(mu);
  if (!g_dev_init[dev]) {
    MvmDevInfo d;
    d.dev = dev;
    d.n_sm = 0;
    cudaDeviceGetAttribute(&d.n_sm, cudaDevAttrMultiProcessorCount, dev);
    int optin = 0;
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    d.max_smem = (size_t)optin;
    g_dev[dev] = d;
    g_dev_init[dev] = true;
  }
  return g_dev[dev];
}

extern "C" {

unsigned long long mvm_launch_count(void) { return g_mvm_launches; }

void mvm_profile_enable(int on) { g_enabled = on != 0; }

// Sums the elapsed milliseconds and launch-scope counts per tag since the last collect.
int mvm_profile_collect(double* ms_per_tag, int* n_per_tag, int n_tags) {
  for (int t = 0; t < n_tags; ++t) { ms_per_tag[t] = 0.0; n_per_tag[t] = 0; }
  for (auto& r : g_recs) {
    cudaEventSynchronize(r.b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    if (r.tag < n_tags) { ms_per_tag[r.tag] += ms; n_per_tag[r.tag] += 1; }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_recs.clear();
  return MVM_OK;
}

}  // extern "C"
