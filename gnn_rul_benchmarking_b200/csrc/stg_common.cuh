// Shared device helpers for the sm_100a kernels of the FC_STGNN hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#define STG_DEVINL __device__ __forceinline__

namespace stg {

constexpr int kMaxWin = 4;        // largest time_window_size the kernels index (reference uses 2)
constexpr float kLeaky = 0.01f;   // F.leaky_relu default slope (Model_Base.py:60,107)

STG_DEVINL float lrelu(float v) { return v > 0.f ? v : kLeaky * v; }
STG_DEVINL float lrelu_grad(float v) { return v > 0.f ? 1.f : kLeaky; }

STG_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
STG_DEVINL float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Warp totals of NP (power of two, <= 32) per-thread values with a halving butterfly: every exchange step halves the
// number of live values (NP - 1 shuffles for NP = 32 instead of 5 NP for one warp_sum per value).  Afterwards lane l
// holds the total of value l / (32 / NP).
template <int NP>
STG_DEVINL float warp_multi_sum(float (&w)[NP]) {
  const int lane = threadIdx.x & 31;
  int o = 16;
#pragma unroll
  for (int n = NP; n > 1; n >>= 1, o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float send = up ? w[i] : w[i + n / 2];
      const float keep = up ? w[i + n / 2] : w[i];
      w[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  float r = w[0];
#pragma unroll
  for (; o >= 1; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  return r;
}

// ---- prologue arithmetic
// 1 / x and 1 / sqrt(x) in double without the division / square-root sequences (40-60 dependent instructions each; every
// CTA of every phase ran several of them back to back in its prologue, 30-50 % of a forward phase's CTA time): fp32 MUFU
// seed + Newton steps in double, relative error < 1e-14.
STG_DEVINL double inv_d(double x) {
  double r = (double)__frcp_rn((float)x);
  r = r * (2.0 - x * r);
  return r * (2.0 - x * r);
}
STG_DEVINL double rsqrt_d(double x) {
  double r = (double)rsqrtf((float)x);
  r = r * (1.5 - 0.5 * x * r * r);
  return r * (1.5 - 0.5 * x * r * r);
}

// ---- mbarrier + 1-D TMA bulk copy (cp.async.bulk, SASS UBLKCP) ---------------------------
STG_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

STG_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
STG_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
STG_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy; dst/src 16-byte aligned, bytes a multiple of 16.
STG_DEVINL void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- programmatic dependent launch -------------------------------------------------------
// Opt-in (STG_PDL=1; measured on B200 inside the captured step: no gain, 0.2795 ms with vs 0.2756 ms without, so it is
// off by default).  Kernels of the training step are then launched with the programmatic-stream-serialization attribute: a kernel may become
// resident (launch latency, shared-memory / TMEM / barrier set-up) while its predecessor in the stream is still
// running.  pdl_sync() is the point behind which the predecessor's memory is visible; nothing before it may touch
// global memory.  It also lets the NEXT kernel start its own set-up, so at most two kernels of the chain overlap.
// Without the launch attribute both instructions are no-ops.
STG_DEVINL void pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
#ifdef __CUDACC__
inline bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("STG_PDL"); return e && *e == '1'; }();
  return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at = {};
  at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

// Cooperative staging of `nfl` contiguous floats (4-byte aligned source) into shared memory.
// Element i lands at dst[shift + i] with shift = ((uintptr_t)src & 15) / 4 so that the 16-byte
// aligned body can go through one TMA bulk copy; the (<4 float) head and tail use plain loads.
// Returns shift.  All threads of the CTA must call it; `bar` must be initialised with count 1.
// The caller waits with mbar_wait(bar, parity) + __syncthreads().
STG_DEVINL int stage_floats_tma(float* dst16, const float* src, int nfl, uint64_t* bar, int tid) {
  const int shift = (int)(((uintptr_t)src & 15u) >> 2);
  int head = shift ? (4 - shift) : 0;
  if (head > nfl) head = nfl;
  const int body = ((nfl - head) >> 2) << 2;
  const int tail = nfl - head - body;
  if (tid == 0) {
    mbar_expect_tx(bar, (uint32_t)body * 4u);   // body == 0: plain arrive completes the phase
    if (body) tma_bulk_g2s(dst16 + shift + head, src + head, (uint32_t)body * 4u, bar);
  }
  if (tid < head) dst16[shift + tid] = src[tid];
  if (tid < tail) dst16[shift + head + body + tid] = src[head + body + tid];
  return shift;
}

// ---- per-kernel CUDA-event timing (stg_profile_* in the C ABI; off by default) ---------------
enum ProfSlot {
  kProfXmoments = 0, kProfFwdMain, kProfFwdFin, kProfBwdStats, kProfBwdMain, kProfBwdFin,
  kProfEncF1, kProfEncF2, kProfEncF3, kProfEncF4, kProfEncB1, kProfEncB2, kProfEncB3, kProfEncB4,
  kProfHeadFc1, kProfHeadTail, kProfHeadBwd1, kProfAdam, kProfZero, kProfBlkPrep, kProfSlots
};
// Records an event pair around the launches issued while the scope is alive (host-side no-op when
// profiling is disabled).  Events go on the same stream as the kernels.
int set_err(int code, const char* fmt, ...);   // thread-local message for stg_last_error()
int check_cuda(const char* what);

struct ProfScope {
  int idx;
  cudaStream_t s;
  ProfScope(int slot, cudaStream_t stream);
  ~ProfScope();
};

}  // namespace stg
