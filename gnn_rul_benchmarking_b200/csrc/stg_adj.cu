// Dense per-graph adjacency builders shared by the sibling models of the path (SURVEY.md section 2.2,
// primitives A2-A4), forward and backward, one CTA per graph (sm_100a):
//   STG_ADJ_PCC       Pearson correlation of the rows      ST_GCN/Model.py:53-71, ST_Conv:10-28, LOGO:17-35
//   STG_ADJ_COSINE    cosine similarity of the rows        HAGCN/Model.py:122-127, SAGCN:74-79
//   STG_ADJ_GAUSS     exp(-||xi-xj||)                      ASTGCNN/Model.py:184-195 (after its Linear P)
//   STG_ADJ_GAUSS2    exp(-||xi-xj||^2), top-k per row     STGNN/Model.py:8-25
//   STG_ADJ_GRAM      x x^T (outer-product adjacency)      STMSGCN/Model.py:96
// X [G, N, F] -> A [G, N, N].  The graph's rows are staged once in shared memory (pitch F+1), row norms /
// means are warp-shuffle reductions, every (i,j) entry is one thread's dot product.  The backward uses
// the closed forms (u = x/|x|: dx = (dU - (dU.u)u)/|x| with dU_i = sum_j (dA_ij + dA_ji) u_j;
// d||xi-xj||/dxi = (xi-xj)/d, 0 at d = 0 like ATen's cdist backward; the top-k mask carries no gradient).
#include <math.h>

#include "../../include/stgconv_b200.h"
#include "stg_common.cuh"

namespace stg {
namespace {

constexpr int kAdjThreads = 256;

// stage X[g] (optionally centred), compute per-row 1/norm.  xs [N][FP], rn [N]
__device__ void adj_stage(const float* __restrict__ X, int N, int F, int FP, bool centre, bool normalise, float* xs,
                          float* rn) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  for (int i = tid; i < N * F; i += blockDim.x) xs[(i / F) * FP + (i % F)] = X[i];
  __syncthreads();
  for (int r = warp; r < N; r += nw) {
    float* row = xs + r * FP;
    if (centre) {
      float s = 0.f;
      for (int c = lane; c < F; c += 32) s += row[c];
      s = warp_sum(s) / (float)F;
      for (int c = lane; c < F; c += 32) row[c] -= s;
      __syncwarp();
    }
    if (normalise) {
      float q = 0.f;
      for (int c = lane; c < F; c += 32) q = fmaf(row[c], row[c], q);
      q = warp_sum(q);
      if (lane == 0) rn[r] = 1.f / sqrtf(q);          // 1/0 = inf -> NaN entries, exactly like the reference
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kAdjThreads) k_adj_fwd(int kind, const float* __restrict__ X, int N, int F, int top_k,
                                                         float* __restrict__ A, unsigned char* __restrict__ mask) {
  extern __shared__ float sm[];
  const int FP = F + 1;
  float* xs = sm;
  float* rn = xs + N * FP;
  float* S = rn + N;                                   // [N][N] (GAUSS2 top-k only)
  const size_t g = blockIdx.x;
  const bool cosine_like = kind == STG_ADJ_PCC || kind == STG_ADJ_COSINE;
  adj_stage(X + g * N * F, N, F, FP, kind == STG_ADJ_PCC, cosine_like, xs, rn);
  float* Ag = A + g * N * N;
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
    const int i = e / N, j = e - i * N;
    const float *xi = xs + i * FP, *xj = xs + j * FP;
    float v;
    if (cosine_like || kind == STG_ADJ_GRAM) {
      float d = 0.f;
      for (int c = 0; c < F; ++c) d = fmaf(xi[c], xj[c], d);
      v = kind == STG_ADJ_GRAM ? d : d * rn[i] * rn[j];
    } else {
      float d2 = 0.f;
      for (int c = 0; c < F; ++c) { const float t = xi[c] - xj[c]; d2 = fmaf(t, t, d2); }
      v = kind == STG_ADJ_GAUSS ? __expf(-sqrtf(d2)) : __expf(-d2);
    }
    if (kind == STG_ADJ_GAUSS2 && top_k > 0 && top_k < N) S[e] = v; else Ag[e] = v;
  }
  if (kind == STG_ADJ_GAUSS2 && top_k > 0 && top_k < N) {
    __syncthreads();
    // keep the top_k largest of every row: rank by counting (ties: lower column first, like a stable sort)
    for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
      const int i = e / N, j = e - i * N;
      const float v = S[e];
      int rank = 0;
      for (int q = 0; q < N; ++q) {
        const float o = S[i * N + q];
        rank += (o > v) || (o == v && q < j);
      }
      const bool keep = rank < top_k;
      Ag[e] = keep ? v : 0.f;
      if (mask) mask[g * N * N + e] = keep ? 1 : 0;
    }
  }
}

__global__ void __launch_bounds__(kAdjThreads) k_adj_bwd(int kind, const float* __restrict__ X, const float* __restrict__ A,
                                                         const float* __restrict__ dA, int N, int F,
                                                         float* __restrict__ dX) {
  extern __shared__ float sm[];
  const int FP = F + 1;
  float* xs = sm;                  // u (normalised rows) for the cosine kinds, raw rows otherwise
  float* rn = xs + N * FP;
  float* Wg = rn + N;              // [N][N] symmetrised pair weights
  float* dU = Wg + N * N;          // [N][FP]
  const size_t g = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  const bool cosine_like = kind == STG_ADJ_PCC || kind == STG_ADJ_COSINE;
  adj_stage(X + g * N * F, N, F, FP, kind == STG_ADJ_PCC, cosine_like, xs, rn);
  const float* dAg = dA + g * N * N;
  const float* Ag = A + g * N * N;
  if (cosine_like || kind == STG_ADJ_GRAM) {
    if (cosine_like)
      for (int i = tid; i < N * F; i += blockDim.x) xs[(i / F) * FP + (i % F)] *= rn[i / F];   // u = x/|x|
    for (int e = tid; e < N * N; e += blockDim.x) {
      const int i = e / N, j = e - i * N;
      Wg[e] = dAg[e] + dAg[j * N + i];
    }
  } else {
    // A_ij = exp(-d) -> dd = -A dA, times (xi-xj)/d ;  A_ij = exp(-d^2) (masked entries are 0) -> -2 A dA (xi-xj)
    for (int e = tid; e < N * N; e += blockDim.x) {
      const int i = e / N, j = e - i * N;
      const float w = Ag[e] * dAg[e] + Ag[j * N + i] * dAg[j * N + i];
      float v;
      if (kind == STG_ADJ_GAUSS) {
        float d2 = 0.f;
        for (int c = 0; c < F; ++c) { const float t = xs[i * FP + c] - xs[j * FP + c]; d2 = fmaf(t, t, d2); }
        const float d = sqrtf(d2);
        v = (d > 0.f) ? -w / d : 0.f;                    // subgradient 0 at d = 0 (the diagonal), like ATen
      } else {
        v = -2.f * w;
      }
      Wg[e] = v;
    }
  }
  __syncthreads();
  // dU_i = sum_j W_ij * (cosine: u_j ; distance kinds: (x_i - x_j))
  for (int e = tid; e < N * F; e += blockDim.x) {
    const int i = e / F, c = e - i * F;
    float acc = 0.f;
    if (cosine_like || kind == STG_ADJ_GRAM) {
      for (int j = 0; j < N; ++j) acc = fmaf(Wg[i * N + j], xs[j * FP + c], acc);
    } else {
      const float xi = xs[i * FP + c];
      for (int j = 0; j < N; ++j) acc = fmaf(Wg[i * N + j], xi - xs[j * FP + c], acc);
    }
    dU[i * FP + c] = acc;
  }
  __syncthreads();
  float* dXg = dX + g * N * F;
  if (!cosine_like) {                                  // distance kinds and GRAM: dX = dU
    for (int e = tid; e < N * F; e += blockDim.x) dXg[e] = dU[(e / F) * FP + (e % F)];
    return;
  }
  // dx_i = (dU_i - (dU_i.u_i) u_i) / |x_i| ;  PCC: minus its mean over the features (centring)
  for (int r = warp; r < N; r += nw) {
    const float* u = xs + r * FP;
    float* d = dU + r * FP;
    float dot = 0.f;
    for (int c = lane; c < F; c += 32) dot = fmaf(d[c], u[c], dot);
    dot = warp_sum(dot);
    float mean = 0.f;
    for (int c = lane; c < F; c += 32) {
      const float v = (d[c] - dot * u[c]) * rn[r];
      d[c] = v;
      mean += v;
    }
    mean = warp_sum(mean) / (float)F;
    if (kind != STG_ADJ_PCC) mean = 0.f;
    for (int c = lane; c < F; c += 32) dXg[r * F + c] = d[c] - mean;
  }
}

size_t adj_smem(int N, int F, bool bwd) {
  size_t fl = (size_t)N * (F + 1) + N + (size_t)N * N;
  if (bwd) fl += (size_t)N * (F + 1);
  return fl * 4;
}
bool g_adj_attr = false;
void adj_attrs() {
  if (g_adj_attr) return;
  cudaFuncSetAttribute(k_adj_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_adj_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  g_adj_attr = true;
}
int adj_check(int kind, const void* a, const void* b, long long G, int N, int F, bool bwd) {
  if (!a || !b || G < 1 || N < 1 || F < 1) return set_err(STG_ERR_INVALID, "bad argument");
  if (kind < STG_ADJ_PCC || kind > STG_ADJ_GRAM) return set_err(STG_ERR_INVALID, "unknown adjacency kind %d", kind);
  if (adj_smem(N, F, bwd) > 200 * 1024)
    return set_err(STG_ERR_UNSUPPORTED, "graph of %d nodes x %d features does not fit shared memory", N, F);
  return STG_OK;
}

}  // namespace
}  // namespace stg

using namespace stg;

// CTA size by graph size: the kernels stride their loops by blockDim.x, so tiny graphs (STMSGCN: 20 480 graphs of 2 nodes)
// get one or two warps per graph and up to 32 graphs resident per SM instead of 256 mostly idle threads each
static inline int graph_threads(int N, int F) {
  const int work = N * F > N * N ? N * F : N * N;
  int t = (work + 31) / 32 * 32;
  return t < 32 ? 32 : (t > stg::kAdjThreads ? stg::kAdjThreads : t);
}

extern "C" int stg_adj_forward(int kind, const float* x_dev, int64_t G, int N, int F, int top_k, float* adj_dev,
                               unsigned char* mask_dev, void* stream) {
  int rc = adj_check(kind, x_dev, adj_dev, G, N, F, false);
  if (rc) return rc;
  adj_attrs();
  k_adj_fwd<<<(unsigned)G, graph_threads(N, F), adj_smem(N, F, false), (cudaStream_t)stream>>>(kind, x_dev, N, F, top_k, adj_dev,
                                                                                      mask_dev);
  return check_cuda("stg_adj_forward");
}

extern "C" int stg_adj_backward(int kind, const float* x_dev, const float* adj_dev, const float* dadj_dev, int64_t G,
                                int N, int F, float* dx_dev, void* stream) {
  int rc = adj_check(kind, x_dev, dx_dev, G, N, F, true);
  if (rc) return rc;
  if (!adj_dev || !dadj_dev) return set_err(STG_ERR_INVALID, "null adjacency / gradient");
  adj_attrs();
  k_adj_bwd<<<(unsigned)G, graph_threads(N, F), adj_smem(N, F, true), (cudaStream_t)stream>>>(kind, x_dev, adj_dev, dadj_dev, N, F,
                                                                                     dx_dev);
  return check_cuda("stg_adj_backward");
}
