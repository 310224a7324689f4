// Optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline leg).
// Disabled by default: ProfScope is then two predictable branches on the host.
#include <mutex>
#include <vector>

#include "../../include/stgconv_b200.h"
#include "stg_common.cuh"

namespace stg {

namespace {
struct Pair { cudaEvent_t a, b; int slot; };
std::mutex g_mu;
bool g_on = false;
std::vector<Pair> g_pairs;      // recorded pairs since the last reset
std::vector<Pair> g_free;       // recycled events
const char* kNames[kProfSlots] = {
    "k_xmoments", "k_block_fwd", "k_block_fwd_fin", "k_block_bwd_stats", "k_block_bwd", "k_block_bwd_fin",
    "k_encoder<F1>", "k_encoder<F2>", "k_encoder<F3>", "k_encoder<F4>",
    "k_encoder<B1>", "k_encoder<B2>", "k_encoder<B3>", "k_encoder<B4>",
    "k_head_fc1", "k_head_tail", "k_head_bwd1", "k_adam", "k_zero", "k_block_prep"};
constexpr size_t kMaxPairs = 1 << 16;
}  // namespace

ProfScope::ProfScope(int slot, cudaStream_t stream) : idx(-1), s(stream) {
  if (!g_on) return;
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_pairs.size() >= kMaxPairs) return;
  Pair p;
  if (!g_free.empty()) { p = g_free.back(); g_free.pop_back(); }
  else if (cudaEventCreate(&p.a) != cudaSuccess || cudaEventCreate(&p.b) != cudaSuccess) return;
  p.slot = slot;
  cudaEventRecord(p.a, s);
  g_pairs.push_back(p);
  idx = (int)g_pairs.size() - 1;
}
ProfScope::~ProfScope() {
  if (idx < 0) return;
  std::lock_guard<std::mutex> lk(g_mu);
  if (idx < (int)g_pairs.size()) cudaEventRecord(g_pairs[idx].b, s);
}
}  // namespace stg

using namespace stg;

extern "C" {

int stg_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_on = on != 0;
  return STG_OK;
}

int stg_profile_reset(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& p : g_pairs) g_free.push_back(p);
  g_pairs.clear();
  return STG_OK;
}

int stg_profile_slots(void) { return kProfSlots; }

const char* stg_profile_name(int slot) { return (slot >= 0 && slot < kProfSlots) ? kNames[slot] : ""; }

int stg_profile_read(int slot, double* total_ms, int64_t* scopes) {
  if (slot < 0 || slot >= kProfSlots || !total_ms || !scopes) return set_err(STG_ERR_INVALID, "bad profile slot");
  std::lock_guard<std::mutex> lk(g_mu);
  double tot = 0.0;
  int64_t n = 0;
  for (auto& p : g_pairs) {
    if (p.slot != slot) continue;
    if (cudaEventSynchronize(p.b) != cudaSuccess) return set_err(STG_ERR_CUDA, "cudaEventSynchronize failed");
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p.a, p.b) != cudaSuccess) return set_err(STG_ERR_CUDA, "cudaEventElapsedTime failed");
    tot += ms;
    ++n;
  }
  *total_ms = tot;
  *scopes = n;
  return STG_OK;
}

}  // extern "C"
