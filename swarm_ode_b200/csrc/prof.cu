#include "prof.cuh"

#include <atomic>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace gnode {
namespace {

struct Pending { cudaEvent_t a, b; std::string label; double flops, bytes; };
struct Acc { int64_t launches = 0; double ms = 0, flops = 0, bytes = 0; };

std::atomic<bool> g_on{false};
std::mutex g_mu;
std::vector<Pending> g_pending;
std::vector<cudaEvent_t> g_free;
std::map<std::string, Acc> g_acc;

cudaEvent_t get_event() {
  if (!g_free.empty()) { cudaEvent_t e = g_free.back(); g_free.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

void drain_locked() {
  for (auto& p : g_pending) {
    float ms = 0.f;
    if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
      Acc& a = g_acc[p.label];
      a.launches++; a.ms += ms; a.flops += p.flops; a.bytes += p.bytes;
    }
    g_free.push_back(p.a); g_free.push_back(p.b);
  }
  g_pending.clear();
}

}  // namespace

bool prof_enabled() { return g_on.load(std::memory_order_relaxed); }

ProfScope::ProfScope(const char* label, cudaStream_t s, double flops, double bytes) : stream(s) {
  if (!prof_enabled()) return;
  std::lock_guard<std::mutex> lk(g_mu);
  Pending p{get_event(), get_event(), label, flops, bytes};
  if (!p.a || !p.b) return;
  cudaEventRecord(p.a, s);
  g_pending.push_back(p);
  slot = (int)g_pending.size() - 1;
}

ProfScope::~ProfScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_mu);
  if (slot < (int)g_pending.size()) cudaEventRecord(g_pending[slot].b, stream);
}

}  // namespace gnode

using namespace gnode;

extern "C" int gnode_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (on) { drain_locked(); g_acc.clear(); }
  g_on.store(on != 0);
  return GNODE_OK;
}

extern "C" int gnode_prof_read(gnode_prof_entry* out, int cap) {
  std::lock_guard<std::mutex> lk(g_mu);
  drain_locked();
  int n = 0;
  for (auto& kv : g_acc) {
    if (out && n < cap) {
      std::memset(&out[n], 0, sizeof(out[n]));
      std::snprintf(out[n].name, sizeof(out[n].name), "%s", kv.first.c_str());
      out[n].launches = kv.second.launches; out[n].ms = kv.second.ms;
      out[n].flops = kv.second.flops; out[n].bytes = kv.second.bytes;
    }
    ++n;
  }
  return n;
}
