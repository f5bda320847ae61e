// Process-level state of libgnode_b200: error string, launch counter, engine switch.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <set>
#include <utility>

#include "common.cuh"

namespace gnode {
namespace {
thread_local char g_err[1024] = "";
std::atomic<int64_t> g_launches{0};
std::atomic<int> g_engine{GNODE_ENGINE_AUTO};
std::atomic<int> g_fold{1};
std::atomic<int> g_dopri5_fsal{-1};   // -1: not set yet, take GNODE_DOPRI5_FSAL (default on)
}  // namespace

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int current_engine() { return g_engine.load(std::memory_order_relaxed); }
int current_fold() { return g_fold.load(std::memory_order_relaxed); }
bool lazy_images_poisoned() {
  static const bool on = [] { const char* e = std::getenv("GNODE_POISON_LAZY"); return e && e[0] == '1'; }();
  return on;
}
int current_dopri5_fsal() {
  int v = g_dopri5_fsal.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = std::getenv("GNODE_DOPRI5_FSAL");
    v = (e && e[0] == '0') ? 0 : 1;
    int expect = -1;
    g_dopri5_fsal.compare_exchange_strong(expect, v);
    v = g_dopri5_fsal.load(std::memory_order_relaxed);
  }
  return v;
}

// Per-device one-time setup (cudaFuncSetAttribute and friends are per device, not per process): true exactly once for
// every (key, current device) pair.
bool first_use_on_device(const void* key) {
  static std::mutex mu;
  static std::set<std::pair<const void*, int>> seen;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return true;
  std::lock_guard<std::mutex> lk(mu);
  return seen.insert({key, dev}).second;
}
}  // namespace gnode

extern "C" const char* gnode_last_error(void) { return gnode::g_err; }
extern "C" int gnode_abi_version(void) { return 1; }
extern "C" int gnode_set_engine(int engine) {
  if (engine < GNODE_ENGINE_AUTO || engine > GNODE_ENGINE_TC) return -1;
  return gnode::g_engine.exchange(engine);
}
extern "C" int gnode_set_fold(int fold) { return gnode::g_fold.exchange(fold ? 1 : 0); }
extern "C" int gnode_set_dopri5_fsal(int on) {
  const int prev = gnode::current_dopri5_fsal();
  gnode::g_dopri5_fsal.store(on ? 1 : 0);
  return prev;
}
extern "C" int64_t gnode_launch_count(void) { return gnode::g_launches.load(std::memory_order_relaxed); }
