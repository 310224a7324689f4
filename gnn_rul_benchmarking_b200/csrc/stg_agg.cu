// Dense graph aggregation of the sibling models (primitives M2 / M3), forward and
// backward, one CTA per graph (sm_100a).  The learnable projections that follow (nn.Linear / filter
// matmuls) are plain GEMMs and stay with the caller.
//   STG_AGG_GCN    Y = D^-1/2 (A+I) D^-1/2 X,  D = rowsum(A+I)      SAGCN/Model.py:81-95, STMSGCN:34-49, RGCNU:7-21
//   STG_AGG_CHEB3  T = [X, A X, 2 A (A X) - X]  (Chebyshev, raw A)   ASTGCNN/Model.py:212-228, STGNN:43-59, STNet:21-37
//   STG_AGG_AX     Y = A X  (MPNN_mk with k = 1, primitive M1)       ST_GCN/Model.py:80-90, ST_Conv, HierCorrPool, LOGO
// X [G,N,F], A [G,N,N] -> GCN: Y [G,N,F];  CHEB3: T [G,3,N,F].
#include <math.h>

#include "../../include/stgconv_b200.h"
#include "stg_common.cuh"

namespace stg {
namespace {

constexpr int kAggThreads = 256;

// out[i][c] (+)= scale * sum_j M(i,j) * in[j][c]   with M = A or A^T (row-major [N][N] in smem)
template <bool TRANS>
__device__ void mat_apply(const float* As, const float* in, float* out, int N, int F, int FP, float scale, bool accumulate) {
  for (int e = threadIdx.x; e < N * F; e += blockDim.x) {
    const int i = e / F, c = e - i * F;
    float acc = 0.f;
    for (int j = 0; j < N; ++j) acc = fmaf(TRANS ? As[j * N + i] : As[i * N + j], in[j * FP + c], acc);
    out[i * FP + c] = accumulate ? fmaf(scale, acc, out[i * FP + c]) : scale * acc;
  }
}

__global__ void __launch_bounds__(kAggThreads) k_agg_fwd(int kind, const float* __restrict__ X, const float* __restrict__ A,
                                                         int N, int F, float* __restrict__ Y) {
  extern __shared__ float sm[];
  const int FP = F + 1, tid = threadIdx.x;
  float* As = sm;                 // [N][N]  (GCN: A + I)
  float* xs = As + N * N;         // [N][FP]
  float* t1 = xs + N * FP;        // [N][FP]
  float* sc = t1 + N * FP;        // [N] deg^-1/2
  const size_t g = blockIdx.x;
  for (int e = tid; e < N * N; e += blockDim.x) As[e] = A[g * N * N + e] + ((kind == STG_AGG_GCN && e / N == e % N) ? 1.f : 0.f);
  for (int e = tid; e < N * F; e += blockDim.x) xs[(e / F) * FP + e % F] = X[g * N * F + e];
  __syncthreads();
  if (kind == STG_AGG_GCN) {
    for (int i = tid; i < N; i += blockDim.x) {
      float d = 0.f;
      for (int j = 0; j < N; ++j) d += As[i * N + j];
      sc[i] = powf(d, -0.5f);                               // NaN for d < 0, inf for d == 0: reference semantics
    }
    __syncthreads();
    for (int e = tid; e < N * F; e += blockDim.x) t1[(e / F) * FP + e % F] = sc[e / F] * xs[(e / F) * FP + e % F];
    __syncthreads();
    float* Yg = Y + g * N * F;
    for (int e = tid; e < N * F; e += blockDim.x) {
      const int i = e / F, c = e - i * F;
      float acc = 0.f;
      for (int j = 0; j < N; ++j) acc = fmaf(As[i * N + j], t1[j * FP + c], acc);
      Yg[e] = sc[i] * acc;
    }
    return;
  }
  if (kind == STG_AGG_AX) {                                  // MPNN_mk, k = 1: plain A X
    float* Yg = Y + g * N * F;
    for (int e = tid; e < N * F; e += blockDim.x) {
      const int i = e / F, c = e - i * F;
      float acc = 0.f;
      for (int j = 0; j < N; ++j) acc = fmaf(As[i * N + j], xs[j * FP + c], acc);
      Yg[e] = acc;
    }
    return;
  }
  float* Tg = Y + g * 3 * N * F;
  for (int e = tid; e < N * F; e += blockDim.x) Tg[e] = xs[(e / F) * FP + e % F];
  mat_apply<false>(As, xs, t1, N, F, FP, 1.f, false);       // T1 = A X
  __syncthreads();
  for (int e = tid; e < N * F; e += blockDim.x) {
    const int i = e / F, c = e - i * F;
    Tg[N * F + e] = t1[i * FP + c];
    float acc = 0.f;
    for (int j = 0; j < N; ++j) acc = fmaf(As[i * N + j], t1[j * FP + c], acc);
    Tg[2 * N * F + e] = 2.f * acc - xs[i * FP + c];          // T2 = 2 A T1 - T0
  }
}

__global__ void __launch_bounds__(kAggThreads) k_agg_bwd(int kind, const float* __restrict__ X, const float* __restrict__ A,
                                                         const float* __restrict__ dY, int N, int F,
                                                         float* __restrict__ dX, float* __restrict__ dA) {
  extern __shared__ float sm[];
  const int FP = F + 1, tid = threadIdx.x;
  float* As = sm;                 // [N][N]
  float* xs = As + N * N;         // [N][FP]   X
  float* b1 = xs + N * FP;        // [N][FP]
  float* b2 = b1 + N * FP;        // [N][FP]
  float* b3 = b2 + N * FP;        // [N][FP]
  float* sc = b3 + N * FP;        // [N]
  float* dsv = sc + N;            // [N]
  float* Gsq = dsv + N;           // [N][N]
  const size_t g = blockIdx.x;
  for (int e = tid; e < N * N; e += blockDim.x) As[e] = A[g * N * N + e] + ((kind == STG_AGG_GCN && e / N == e % N) ? 1.f : 0.f);
  for (int e = tid; e < N * F; e += blockDim.x) xs[(e / F) * FP + e % F] = X[g * N * F + e];
  __syncthreads();
  float* dXg = dX + g * N * F;
  float* dAg = dA + g * N * N;
  if (kind == STG_AGG_GCN) {
    const float* dYg = dY + g * N * F;
    float* deg = dsv;           // reuse: first the degrees, then d(deg)
    for (int i = tid; i < N; i += blockDim.x) {
      float d = 0.f;
      for (int j = 0; j < N; ++j) d += As[i * N + j];
      sc[i] = powf(d, -0.5f);
      deg[i] = d;
    }
    for (int e = tid; e < N * F; e += blockDim.x) b1[(e / F) * FP + e % F] = dYg[e];       // dY
    __syncthreads();
    // G_ij = dY_i . X_j  -> b-space too large; computed on the fly below
    // dX_j = s_j * sum_i A~_ij s_i dY_i
    for (int e = tid; e < N * F; e += blockDim.x) {
      const int j = e / F, c = e - j * F;
      float acc = 0.f;
      for (int i = 0; i < N; ++i) acc = fmaf(As[i * N + j] * sc[i], b1[i * FP + c], acc);
      dXg[e] = sc[j] * acc;
    }
    // direct term and ds:  dA~_ij = s_i s_j G_ij ;  ds_i = sum_j A~_ij s_j G_ij + sum_j A~_ji s_j G_ji
    float* Gs = Gsq;
    for (int e = tid; e < N * N; e += blockDim.x) {
      const int i = e / N, j = e - i * N;
      float d = 0.f;
      for (int c = 0; c < F; ++c) d = fmaf(b1[i * FP + c], xs[j * FP + c], d);
      Gs[e] = d;
    }
    __syncthreads();
    for (int i = tid; i < N; i += blockDim.x) {
      float ds = 0.f;
      for (int j = 0; j < N; ++j) ds += As[i * N + j] * sc[j] * Gs[i * N + j] + As[j * N + i] * sc[j] * Gs[j * N + i];
      const float d = deg[i];
      b1[i] = -0.5f * powf(d, -1.5f) * ds;      // d(deg_i), parked in b1 row 0 (dY no longer needed)
    }
    __syncthreads();
    for (int e = tid; e < N * N; e += blockDim.x) {
      const int i = e / N, j = e - i * N;
      dAg[e] = sc[i] * sc[j] * Gs[e] + b1[i];
    }
    return;
  }
  if (kind == STG_AGG_AX) {                                  // dX = A^T dY ; dA = dY X^T
    const float* dYg = dY + g * N * F;
    for (int e = tid; e < N * F; e += blockDim.x) b1[(e / F) * FP + e % F] = dYg[e];
    __syncthreads();
    for (int e = tid; e < N * F; e += blockDim.x) {
      const int j = e / F, c = e - j * F;
      float acc = 0.f;
      for (int i = 0; i < N; ++i) acc = fmaf(As[i * N + j], b1[i * FP + c], acc);
      dXg[e] = acc;
    }
    for (int e = tid; e < N * N; e += blockDim.x) {
      const int i = e / N, j = e - i * N;
      float d = 0.f;
      for (int c = 0; c < F; ++c) d = fmaf(b1[i * FP + c], xs[j * FP + c], d);
      dAg[e] = d;
    }
    return;
  }
  // ---- CHEB3: dT [3][N][F];  dT1' = dT1 + 2 A^T dT2 ; dX = dT0 - dT2 + A^T dT1' ; dA = dT1' X^T + 2 dT2 T1^T
  const float* dTg = dY + g * 3 * N * F;
  for (int e = tid; e < N * F; e += blockDim.x) {
    b1[(e / F) * FP + e % F] = dTg[N * F + e];             // dT1
    b2[(e / F) * FP + e % F] = dTg[2 * N * F + e];         // dT2
  }
  mat_apply<false>(As, xs, b3, N, F, FP, 1.f, false);       // T1 = A X
  __syncthreads();
  mat_apply<true>(As, b2, b1, N, F, FP, 2.f, true);         // dT1' = dT1 + 2 A^T dT2
  __syncthreads();
  for (int e = tid; e < N * F; e += blockDim.x) {
    const int i = e / F, c = e - i * F;
    float acc = 0.f;
    for (int j = 0; j < N; ++j) acc = fmaf(As[j * N + i], b1[j * FP + c], acc);
    dXg[e] = dTg[e] - b2[i * FP + c] + acc;
  }
  for (int e = tid; e < N * N; e += blockDim.x) {
    const int i = e / N, j = e - i * N;
    float d = 0.f;
    for (int c = 0; c < F; ++c) d += b1[i * FP + c] * xs[j * FP + c] + 2.f * b2[i * FP + c] * b3[j * FP + c];
    dAg[e] = d;
  }
}

size_t agg_smem(int N, int F, bool bwd) {
  const size_t FP = F + 1;
  size_t fl = (size_t)N * N * (bwd ? 2 : 1) + (bwd ? 4 : 2) * (size_t)N * FP + 2 * N;
  return fl * 4;
}
bool g_agg_attr = false;

}  // namespace
}  // namespace stg

using namespace stg;

static int agg_check(int kind, const void* x, const void* a, const void* y, long long G, int N, int F, bool bwd) {
  if (!x || !a || !y || G < 1 || N < 1 || F < 1) return set_err(STG_ERR_INVALID, "bad argument");
  if (kind != STG_AGG_GCN && kind != STG_AGG_CHEB3 && kind != STG_AGG_AX) return set_err(STG_ERR_INVALID, "unknown aggregation kind %d", kind);
  if (agg_smem(N, F, bwd) > 200 * 1024)
    return set_err(STG_ERR_UNSUPPORTED, "graph of %d nodes x %d features does not fit the aggregation tile", N, F);
  if (!g_agg_attr) {
    cudaFuncSetAttribute(k_agg_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_agg_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    g_agg_attr = true;
  }
  return STG_OK;
}

// CTA size by graph size: the kernels stride their loops by blockDim.x, so tiny graphs (STMSGCN: 20 480 graphs of 2 nodes)
// get one or two warps per graph and up to 32 graphs resident per SM instead of 256 mostly idle threads each
static inline int graph_threads(int N, int F) {
  const int work = N * F > N * N ? N * F : N * N;
  int t = (work + 31) / 32 * 32;
  return t < 32 ? 32 : (t > stg::kAggThreads ? stg::kAggThreads : t);
}

extern "C" int stg_agg_forward(int kind, const float* x_dev, const float* adj_dev, int64_t G, int N, int F,
                               float* out_dev, void* stream) {
  int rc = agg_check(kind, x_dev, adj_dev, out_dev, G, N, F, false);
  if (rc) return rc;
  k_agg_fwd<<<(unsigned)G, graph_threads(N, F), agg_smem(N, F, false), (cudaStream_t)stream>>>(kind, x_dev, adj_dev, N, F, out_dev);
  return check_cuda("stg_agg_forward");
}

extern "C" int stg_agg_backward(int kind, const float* x_dev, const float* adj_dev, const float* dout_dev, int64_t G,
                                int N, int F, float* dx_dev, float* dadj_dev, void* stream) {
  int rc = agg_check(kind, x_dev, adj_dev, dout_dev, G, N, F, true);
  if (rc) return rc;
  if (!dx_dev || !dadj_dev) return set_err(STG_ERR_INVALID, "null gradient output");
  k_agg_bwd<<<(unsigned)G, graph_threads(N, F), agg_smem(N, F, true), (cudaStream_t)stream>>>(kind, x_dev, adj_dev, dout_dev, N, F,
                                                                                     dx_dev, dadj_dev);
  return check_cuda("stg_agg_backward");
}
