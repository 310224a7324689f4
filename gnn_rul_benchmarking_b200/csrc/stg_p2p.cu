// Data-parallel gradient exchange fused with the optimizer (sm_100a, NVLink 5 / NVSwitch peer memory):
// ONE kernel per step and rank that
//   1. publishes "my flat gradient is complete" into every peer's flag block (st.release.sys over NVLink),
//   2. waits until every peer has published the same step,
//   3. reads all `world` gradient buffers directly through peer-mapped pointers (one-shot all-reduce:
//      266 KB per rank for FC_STGNN FD004, summed in rank order so that all replicas get identical bits),
//   4. applies torch.optim.Adam semantics (algorithms/algorithms.py:60-64) to the local parameters,
//   5. publishes "done reading" and leaves only when every peer is done, so the next step may overwrite
//      the gradient buffers.
// The epoch of the flag protocol is a counter in the flag block itself (word 34, starts at 1, bumped by the
// last block of every exchange): it only ever grows, independently of the Adam step, which a CUDA-graph
// capture snapshots and restores.  A rank whose wait times out (about 2 s) raises word 33 and SKIPS the
// parameter update; the host checks the word (FC_STGNN.update / check_exchange) and raises.
// It replaces ncclAllReduce + k_adam (latency-bound at this size: ~45 us at 8 GPUs inside the step graph).
// The peer pointers come from torch's symmetric-memory rendezvous on the host side (engine.py).
#include <math.h>

#include "../../include/stgconv_b200.h"
#include "stg_model.cuh"

namespace stg {
namespace {

constexpr int kMaxWorld = 16;
struct P2PArgs {
  const float* grad[kMaxWorld];
  unsigned* flags[kMaxWorld];      // per rank: [0,16) ready, [16,32) done, [32] block counter, [33] timeout flag,
                                   //           [34] epoch of the flag protocol (next exchange)
  int rank, world;
};

STG_DEVINL void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
STG_DEVINL unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// bounded spin (about 2 s): a missing peer must not hang the GPU; the timeout flag is checked on the host
STG_DEVINL void wait_flag(const unsigned* p, unsigned epoch, unsigned* err) {
  const long long t0 = clock64();
  while ((int)(ld_acquire_sys(p) - epoch) < 0) {
    if (clock64() - t0 > 4000000000ll) { *err = 1u; break; }
    __nanosleep(64);
  }
}

__global__ void __launch_bounds__(256) k_allreduce_adam(const P2PArgs pa, float* __restrict__ p, float* __restrict__ m,
                                                        float* __restrict__ v, long long n4,
                                                        const long long* __restrict__ step, float lr, float b1,
                                                        float b2, float eps, float wd) {
  __shared__ float s_bc[2];
  __shared__ int s_last;
  const int tid = threadIdx.x, world = pa.world;
  unsigned* mine = pa.flags[pa.rank];
  const unsigned epoch = *reinterpret_cast<volatile unsigned*>(mine + 34);   // bumped by the last block only
  __shared__ unsigned s_err;
  if (tid == 0) s_err = 0u;
  __syncthreads();
  if (blockIdx.x == 0 && tid < world) {
    __threadfence_system();                                  // gradients of the earlier kernels -> visible to peers
    st_release_sys(pa.flags[tid] + pa.rank, epoch);
  }
  if (tid < world) {
    unsigned err = 0u;
    wait_flag(mine + tid, epoch, &err);
    if (err) { mine[33] = 1u; s_err = 1u; }
  }
  if (tid == 0) {
    const double t = (double)(*step);
    s_bc[0] = (float)(1.0 - pow((double)b1, t));
    s_bc[1] = (float)sqrt(1.0 - pow((double)b2, t));
  }
  __syncthreads();
  const float bc1 = s_bc[0], bc2s = s_bc[1], step_size = lr / bc1, gscale = 1.f / (float)world;
  // a peer that never showed up: leave the parameters alone (the host raises on word 33)
  const bool skip = s_err != 0u || *reinterpret_cast<volatile unsigned*>(mine + 33) != 0u;
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (long long i = (long long)blockIdx.x * blockDim.x + tid; i < n4 && !skip; i += (long long)gridDim.x * blockDim.x) {
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < world; ++r) {                        // rank order: identical sums on every replica
      const float4 q = __ldcv(reinterpret_cast<const float4*>(pa.grad[r]) + i);
      g.x += q.x; g.y += q.y; g.z += q.z; g.w += q.w;
    }
    float4 pv = p4[i], mv = m4[i], vv = v4[i];
    float* gp = &g.x; float* pp = &pv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float gv = fmaf(wd, pp[u], gp[u] * gscale);
      mp[u] = fmaf(1.f - b1, gv - mp[u], mp[u]);
      vp[u] = fmaf(1.f - b2, gv * gv, b2 * vp[u]);
      pp[u] -= step_size * (mp[u] / (sqrtf(vp[u]) / bc2s + eps));
    }
    p4[i] = pv; m4[i] = mv; v4[i] = vv;
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    s_last = (atomicAdd(mine + 32, 1u) == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  if (tid == 0) mine[32] = 0u;
  if (tid < world) {
    __threadfence_system();
    st_release_sys(pa.flags[tid] + 16 + pa.rank, epoch);    // "I am done reading your gradients"
    unsigned err = 0u;
    wait_flag(mine + 16 + tid, epoch, &err);                // nobody still reads mine
    if (err) mine[33] = 1u;
  }
  __syncthreads();
  if (tid == 0) mine[34] = epoch + 1u;                      // next exchange (stream order makes it visible)
}

}  // namespace
}  // namespace stg

using namespace stg;

extern "C" int stg_allreduce_adam(float* param_dev, float* exp_avg_dev, float* exp_avg_sq_dev, int64_t n,
                                  int64_t* step_dev, const float* const* grad_ptrs, uint32_t* const* flag_ptrs,
                                  int rank, int world, float lr, float beta1, float beta2, float eps,
                                  float weight_decay, void* stream) {
  if (!param_dev || !exp_avg_dev || !exp_avg_sq_dev || !step_dev || !grad_ptrs || !flag_ptrs)
    return set_err(STG_ERR_INVALID, "null pointer");
  if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world) return set_err(STG_ERR_INVALID, "bad rank / world");
  if (n < 0 || (n & 3)) return set_err(STG_ERR_INVALID, "n must be a multiple of 4 (flat buffers are padded)");
  P2PArgs pa = {};
  for (int r = 0; r < world; ++r) {
    if (!grad_ptrs[r] || !flag_ptrs[r]) return set_err(STG_ERR_INVALID, "null peer pointer for rank %d", r);
    pa.grad[r] = grad_ptrs[r];
    pa.flags[r] = (unsigned*)flag_ptrs[r];
  }
  pa.rank = rank; pa.world = world;
  cudaStream_t s = (cudaStream_t)stream;
  long long* ctr[1] = {(long long*)step_dev};
  launch_tick(ctr, 1, s);                                    // ++step (Adam bias correction only)
  const long long n4 = n / 4;
  int grid = (int)((n4 + 255) / 256);
  if (grid > 132) grid = 132;                                // co-resident by a wide margin: blocks never wait on each other
  if (grid < 1) grid = 1;
  {
    ProfScope ps(kProfAdam, s);
    k_allreduce_adam<<<grid, 256, 0, s>>>(pa, param_dev, exp_avg_dev, exp_avg_sq_dev, n4, (const long long*)step_dev, lr,
                                          beta1, beta2, eps, weight_decay);
  }
  return check_cuda("stg_allreduce_adam");
}
