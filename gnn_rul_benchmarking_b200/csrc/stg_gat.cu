// Dense graph attention of GAT_LSTM (primitive M5), forward and backward, one CTA per graph
// (sm_100a).  GraphAttentionLayer.forward, models/GAT_LSTM/Model.py:87-109, after its nn.Linear:
//   s_i = a[0:F].Wh_i,  t_j = a[F:2F].Wh_j,  z_ij = s_i + t_j + b          (the [N*N, 2F] concat is never built)
//   P = softmax_j(leaky_relu_alpha(z))           -- over ALL j, the adjacency is applied after the softmax
//   Pd = dropout(P) * adj ;  out = leaky_relu_slope(Pd Wh)
// Wh [G,N,F], adj [N,N] (shared) or [G,N,N], keep [G,N,N] 0/1 dropout mask or NULL, out [G,N,F].
// The backward recomputes s, t, z, P from Wh (cheap) and needs only `out` for the outer ReLU mask.
#include <math.h>

#include "../../include/stgconv_b200.h"
#include "stg_common.cuh"

namespace stg {
namespace {

constexpr int kGatThreads = 256;
constexpr int kGatWarps = kGatThreads / 32;

struct GatArgs {
  const float *Wh, *aw, *ab, *adj, *keep, *out, *dout;
  float *y, *dWh, *daw, *dab;
  int adj_per_graph, N, F;
  float pdrop, alpha, slope;
};

// s, t, P (softmax, before dropout / adjacency), zpos (z > 0 flags) and Pd into shared memory
// Non-zero lists of the mask adj * keep (path graphs: 3 entries per row of 40): nzr[i][.] = columns of row i, nzc[j][.] = rows
// of column j, counts in cnt[0..N) / cnt[N..2N).  One warp per row / column, ballot + prefix popcount.
__device__ void gat_lists(const GatArgs& a, size_t g, unsigned char* nzr, unsigned char* nzc, int* cnt) {
  const int N = a.N, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* adj = a.adj + (a.adj_per_graph ? g * N * N : 0);
  const float* keep = a.keep ? a.keep + g * N * N : nullptr;
  for (int r = warp; r < 2 * N; r += kGatWarps) {
    const bool col = r >= N;
    const int i = col ? r - N : r;
    int n = 0;
    for (int j0 = 0; j0 < N; j0 += 32) {
      const int j = j0 + lane;
      bool nz = false;
      if (j < N) {
        const int e = col ? j * N + i : i * N + j;
        nz = adj[e] != 0.f && (!keep || keep[e] != 0.f);
      }
      const unsigned m = __ballot_sync(0xffffffffu, nz);
      if (nz) (col ? nzc : nzr)[i * N + n + __popc(m & ((1u << lane) - 1))] = (unsigned char)j;
      n += __popc(m);
    }
    if (lane == 0) cnt[r] = n;
  }
}

__device__ void gat_scores(const GatArgs& a, size_t g, const float* whs, int FP, float* s, float* t, float* P, float* Pd,
                           unsigned char* zpos) {
  const int N = a.N, F = a.F, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, NP = N + 1;
  for (int i = warp; i < N; i += kGatWarps) {
    float s1 = 0.f, s2 = 0.f;
    for (int f = lane; f < F; f += 32) {
      const float v = whs[i * FP + f];
      s1 = fmaf(a.aw[f], v, s1);
      s2 = fmaf(a.aw[F + f], v, s2);
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) { s[i] = s1; t[i] = s2; }
  }
  __syncthreads();
  const float b = a.ab[0];
  const float ksc = a.pdrop > 0.f ? 1.f / (1.f - a.pdrop) : 1.f;
  const float* adj = a.adj + (a.adj_per_graph ? g * N * N : 0);
  const float* keep = a.keep ? a.keep + g * N * N : nullptr;
  for (int i = warp; i < N; i += kGatWarps) {        // one warp per row: max, exp, sum
    float mx = -INFINITY;
    for (int j = lane; j < N; j += 32) {
      const float z = s[i] + t[j] + b;
      const float e = z > 0.f ? z : a.alpha * z;
      zpos[i * N + j] = z > 0.f;
      P[i * NP + j] = e;
      mx = fmaxf(mx, e);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < N; j += 32) {
      const float e = __expf(P[i * NP + j] - mx);
      P[i * NP + j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int j = lane; j < N; j += 32) {
      const float p = P[i * NP + j] * inv;
      P[i * NP + j] = p;
      Pd[i * NP + j] = p * adj[i * N + j] * (keep ? keep[i * N + j] * ksc : 1.f);
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kGatThreads) k_gat_fwd(const GatArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int N = a.N, F = a.F, FP = F + 1, NP = N + 1, tid = threadIdx.x;
  float* whs = sm;                       // [N][FP]
  float* s = whs + N * FP;               // [N]
  float* t = s + N;                      // [N]
  float* P = t + N;                      // [N][NP]
  float* Pd = P + N * NP;                // [N][NP]
  unsigned char* zpos = reinterpret_cast<unsigned char*>(Pd + N * NP);
  unsigned char* nzr = zpos + N * N;
  unsigned char* nzc = nzr + N * N;
  int* cnt = reinterpret_cast<int*>(zpos + ((3 * N * N + 3) & ~3));
  const size_t g = blockIdx.x;
  const float* Wg = a.Wh + g * N * F;
  for (int e = tid; e < N * F; e += kGatThreads) whs[(e / F) * FP + e % F] = Wg[e];
  gat_lists(a, g, nzr, nzc, cnt);
  __syncthreads();
  gat_scores(a, g, whs, FP, s, t, P, Pd, zpos);
  float* yg = a.y + g * N * F;
  for (int e = tid; e < N * F; e += kGatThreads) {
    const int i = e / F, f = e - i * F;
    float acc = 0.f;
    for (int q = 0; q < cnt[i]; ++q) {                          // path graphs: 3 non-zeros per row
      const int j = nzr[i * N + q];
      acc = fmaf(Pd[i * NP + j], whs[j * FP + f], acc);
    }
    yg[e] = acc > 0.f ? acc : a.slope * acc;
  }
}

// STAGE: dO = dout * outer-ReLU slope and the combined adjacency / dropout mask are staged in shared memory once (they were
// fetched from global memory inside serial per-(i, j) loops: 119 us per launch at N = 40, F = 300 with 12 % of the warps
// active); the unstaged variant serves graphs whose second [N][F] plane does not fit.
template <bool STAGE>
__global__ void __launch_bounds__(kGatThreads) k_gat_bwd(const GatArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int N = a.N, F = a.F, FP = F + 1, NP = N + 1, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* whs = sm;                       // [N][FP]
  float* s = whs + N * FP;               // [N]  s, later ds
  float* t = s + N;                      // [N]  t, later dt
  float* P = t + N;                      // [N][NP]
  float* Pd = P + N * NP;                // [N][NP]
  float* dZ = Pd + N * NP;               // [N][NP]  dP, then dz
  float* dos = dZ + N * NP;              // STAGE: [N][FP] dO
  unsigned char* zpos = reinterpret_cast<unsigned char*>(dos + (STAGE ? N * FP : 0));
  unsigned char* nzr = zpos + N * N;
  unsigned char* nzc = nzr + N * N;
  int* cnt = reinterpret_cast<int*>(zpos + ((3 * N * N + 3) & ~3));
  const size_t g = blockIdx.x;
  const float *Wg = a.Wh + g * N * F, *og = a.out + g * N * F, *dg = a.dout + g * N * F;
  for (int e = tid; e < N * F; e += kGatThreads) {
    whs[(e / F) * FP + e % F] = Wg[e];
    if (STAGE) dos[(e / F) * FP + e % F] = dg[e] * (og[e] > 0.f ? 1.f : a.slope);
  }
  // dO = dout * outer-ReLU slope; sign(out) = sign(Pd Wh) for slope > 0
  auto dO = [&](int i, int f) { return STAGE ? dos[i * FP + f] : dg[i * F + f] * (og[i * F + f] > 0.f ? 1.f : a.slope); };
  gat_lists(a, g, nzr, nzc, cnt);
  __syncthreads();
  gat_scores(a, g, whs, FP, s, t, P, Pd, zpos);
  const float ksc = a.pdrop > 0.f ? 1.f / (1.f - a.pdrop) : 1.f;
  const float* adj = a.adj + (a.adj_per_graph ? g * N * N : 0);
  const float* keep = a.keep ? a.keep + g * N * N : nullptr;
  // one warp per row i:  dP_ij = (dO_i . Wh_j) adj_ij keepscale_ij  (only where that mask is non-zero),
  // de = P (dP - sum_j P dP),  dz = de * leaky_relu'(z),  ds_i = sum_j dz_ij
  if (STAGE) {                                                      // mask plane into dZ (overwritten row by row below)
    for (int e = tid; e < N * N; e += kGatThreads)
      dZ[(e / N) * NP + e % N] = adj[e] * (keep ? keep[e] * ksc : 1.f);
    __syncthreads();
  }
  for (int i = warp; i < N; i += kGatWarps) {
    float rowdot = 0.f;                                            // identical in every lane
    for (int j = 0; j < N; ++j) {
      const float m = STAGE ? dZ[i * NP + j] : adj[i * N + j] * (keep ? keep[i * N + j] * ksc : 1.f);
      float dp = 0.f;
      if (m != 0.f) {                                              // warp-uniform
        float d = 0.f;
        for (int f = lane; f < F; f += 32) d = fmaf(dO(i, f), whs[j * FP + f], d);
        dp = warp_sum(d) * m;
      }
      __syncwarp();
      if (lane == 0) dZ[i * NP + j] = dp;
      rowdot = fmaf(P[i * NP + j], dp, rowdot);
    }
    __syncwarp();
    float ds = 0.f;
    for (int j = lane; j < N; j += 32) {
      const float dz = P[i * NP + j] * (dZ[i * NP + j] - rowdot) * (zpos[i * N + j] ? 1.f : a.alpha);
      dZ[i * NP + j] = dz;
      ds += dz;
    }
    ds = warp_sum(ds);
    if (lane == 0) s[i] = ds;
  }
  __syncthreads();
  for (int j = tid; j < N; j += kGatThreads) {
    float dt = 0.f;
    for (int i = 0; i < N; ++i) dt += dZ[i * NP + j];
    t[j] = dt;
  }
  if (tid == 0) {
    float db = 0.f;
    for (int i = 0; i < N; ++i) db += s[i];
    atomicAdd(a.dab, db);
  }
  __syncthreads();
  // dWh_j = sum_i Pd_ij dO_i + ds_j a1 + dt_j a2
  float* dWg = a.dWh + g * N * F;
  for (int e = tid; e < N * F; e += kGatThreads) {
    const int j = e / F, f = e - j * F;
    float acc = fmaf(s[j], a.aw[f], t[j] * a.aw[F + f]);
    for (int q = 0; q < cnt[N + j]; ++q) {
      const int i = nzc[j * N + q];
      acc = fmaf(Pd[i * NP + j], dO(i, f), acc);
    }
    dWg[e] = acc;
  }
  // da1 += sum_i ds_i Wh_i,  da2 += sum_j dt_j Wh_j
  for (int f = tid; f < F; f += kGatThreads) {
    float g1 = 0.f, g2 = 0.f;
    for (int i = 0; i < N; ++i) {
      const float v = whs[i * FP + f];
      g1 = fmaf(s[i], v, g1);
      g2 = fmaf(t[i], v, g2);
    }
    atomicAdd(&a.daw[f], g1);
    atomicAdd(&a.daw[F + f], g2);
  }
}

size_t gat_smem(int N, int F, bool bwd, bool stage = false) {
  const size_t NP = N + 1, FP = F + 1;
  size_t fl = N * FP + 2 * N + (bwd ? 3 : 2) * N * NP + (stage ? N * FP : 0);
  return fl * 4 + 3 * (size_t)N * N + 8 * (size_t)N + 32;          // + z > 0 flags, row / column non-zero lists, counts
}

int gat_check(const GatArgs& a, int G, bool bwd) {
  if (G < 1 || a.N < 1 || a.F < 1) return set_err(STG_ERR_INVALID, "gat: non-positive dimension");
  if (a.N > 256) return set_err(STG_ERR_UNSUPPORTED, "gat: %d nodes per graph (node lists are 8-bit)", a.N);
  if (!a.Wh || !a.aw || !a.ab || !a.adj) return set_err(STG_ERR_INVALID, "gat: null pointer");
  if (a.pdrop < 0.f || a.pdrop >= 1.f) return set_err(STG_ERR_INVALID, "gat: dropout must be in [0,1)");
  if (a.slope <= 0.f) return set_err(STG_ERR_INVALID, "gat: the output leaky_relu slope must be positive");
  if (gat_smem(a.N, a.F, bwd) > 200 * 1024)
    return set_err(STG_ERR_UNSUPPORTED, "gat: graph of %d nodes x %d features does not fit shared memory", a.N, a.F);
  return 0;
}

}  // namespace
}  // namespace stg

using namespace stg;

extern "C" int stg_gat_forward(const float* Wh_dev, const float* att_w_dev, const float* att_b_dev, const float* adj_dev,
                               int adj_per_graph, const float* keep_dev, float pdrop, float alpha, float out_slope,
                               int G, int N, int F, float* out_dev, void* stream) {
  GatArgs a = {};
  a.Wh = Wh_dev; a.aw = att_w_dev; a.ab = att_b_dev; a.adj = adj_dev; a.keep = keep_dev; a.y = out_dev;
  a.adj_per_graph = adj_per_graph; a.N = N; a.F = F; a.pdrop = keep_dev ? pdrop : 0.f; a.alpha = alpha; a.slope = out_slope;
  if (int rc = gat_check(a, G, false)) return rc;
  if (!out_dev) return set_err(STG_ERR_INVALID, "gat: null output");
  const size_t smem = gat_smem(N, F, false);
  cudaFuncSetAttribute(k_gat_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k_gat_fwd<<<G, kGatThreads, smem, (cudaStream_t)stream>>>(a);
  return check_cuda("stg_gat_forward");
}

extern "C" int stg_gat_backward(const float* Wh_dev, const float* att_w_dev, const float* att_b_dev, const float* adj_dev,
                                int adj_per_graph, const float* keep_dev, float pdrop, float alpha, float out_slope,
                                int G, int N, int F, const float* out_dev, const float* dout_dev, float* dWh_dev,
                                float* datt_w_dev, float* datt_b_dev, void* stream) {
  GatArgs a = {};
  a.Wh = Wh_dev; a.aw = att_w_dev; a.ab = att_b_dev; a.adj = adj_dev; a.keep = keep_dev;
  a.out = out_dev; a.dout = dout_dev; a.dWh = dWh_dev; a.daw = datt_w_dev; a.dab = datt_b_dev;
  a.adj_per_graph = adj_per_graph; a.N = N; a.F = F; a.pdrop = keep_dev ? pdrop : 0.f; a.alpha = alpha; a.slope = out_slope;
  if (int rc = gat_check(a, G, true)) return rc;
  if (!out_dev || !dout_dev || !dWh_dev || !datt_w_dev || !datt_b_dev) return set_err(STG_ERR_INVALID, "gat: null pointer");
  const bool stage = gat_smem(N, F, true, true) <= 200 * 1024;
  const size_t smem = gat_smem(N, F, true, stage);
  if (stage) {
    cudaFuncSetAttribute(k_gat_bwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k_gat_bwd<true><<<G, kGatThreads, smem, (cudaStream_t)stream>>>(a);
  } else {
    cudaFuncSetAttribute(k_gat_bwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k_gat_bwd<false><<<G, kGatThreads, smem, (cudaStream_t)stream>>>(a);
  }
  return check_cuda("stg_gat_backward");
}
