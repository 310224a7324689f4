// TemporalConvNet of the sibling models (primitive T1), forward and backward (sm_100a):
//   models/ASTGCNN/Model.py:72-146 (kernel 6), models/ST_GCN/Model.py:99-173 (kernel 2), ST_Conv, STAGNN
//     x0    = ReLU(BN1(causal_conv(in,  W1, dilation 1)))        conv(pad (K-1)d, no bias) -> Chomp1d -> BN -> ReLU
//     out_0 = ReLU(x0 + in)                                      (downsample0 is None: C_in == C_out in every config)
//     x1    = ReLU(BN2(causal_conv(out_0, W2, dilation 2)))
//     out   = ReLU(x1 + out_0)
// in / out [B, C, L].  BatchNorm uses batch statistics over (B, L) in training mode, so both directions are
// phase-structured like the encoder: F1 conv1 moments -> F2 conv2 moments -> F3 output (eval: F3 only);
// B1 BN2 sums -> B2 dW2, BN1 sums -> B3 dW1, d(in).  A CTA owns one sample at a time (all of its [C][L]
// planes live in shared memory, every phase recomputes the cheap forward chain) and keeps its weight-gradient
// and moment partials in shared memory across the samples it processes.
#include <math.h>
#include <string.h>

#include "../../include/stgconv_b200.h"
#include "stg_common.cuh"

namespace stg {
namespace {

constexpr int kTcnThreads = 128;

struct TcnArgs {
  int B, C, L, K;
  const float* x;
  const float *W1, *W2, *g1, *be1, *g2, *be2;
  float *rm1, *rv1, *rm2, *rv2;
  long long *nbt1, *nbt2;
  int training;
  float momentum, eps;
  double* st;           // [2C bn1 fwd][2C bn2 fwd][2C bn2 bwd][2C bn1 bwd]
  float* out;
  const float* dout;
  float* dx;
  float *dW1, *dW2, *dg1, *dbe1, *dg2, *dbe2;
  float* sv;            // optional [B][2][C*L]: raw conv1 / conv2 outputs kept between the phases (null: recomputed)
};

__device__ void bn_coef(float* dst, int C, const double* stats, double cnt, const float* g, const float* be, float* rm,
                        float* rv, float eps, float momentum, bool update) {
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double m, var;
    if (stats) {
      m = stats[c] / cnt;
      var = stats[C + c] / cnt - m * m;
      if (var < 0.0) var = 0.0;
      if (update) {
        const double unb = cnt > 1.0 ? var * cnt / (cnt - 1.0) : var;
        rm[c] = (1.f - momentum) * rm[c] + momentum * (float)m;
        rv[c] = (1.f - momentum) * rv[c] + momentum * (float)unb;
      }
    } else {
      m = rm[c];
      var = rv[c];
    }
    const float r = (float)(1.0 / sqrt(var + (double)eps));
    const float A = g[c] * r;
    dst[c] = A; dst[C + c] = be[c] - A * (float)m; dst[2 * C + c] = (float)m; dst[3 * C + c] = r;
  }
}

// out[co][t] = sum_ci sum_j W[co][ci][j] * in[ci][t - (K-1-j)*dil]
__device__ void causal_conv(const float* W, const float* in, float* out, int C, int L, int LP, int K, int dil) {
  for (int e = threadIdx.x; e < C * L; e += blockDim.x) {
    const int co = e / L, t = e - co * L;
    float acc = 0.f;
    for (int ci = 0; ci < C; ++ci) {
      const float* w = W + (co * C + ci) * K;
      const float* row = in + ci * LP;
      for (int j = 0; j < K; ++j) {
        const int s = t - (K - 1 - j) * dil;
        if (s >= 0) acc = fmaf(w[j], row[s], acc);
      }
    }
    out[co * LP + t] = acc;
  }
}
// din[ci][s] (+)= sum_co sum_j W[co][ci][j] * dc[co][s + (K-1-j)*dil]
__device__ void causal_conv_t(const float* W, const float* dc, float* din, int C, int L, int LP, int K, int dil) {
  for (int e = threadIdx.x; e < C * L; e += blockDim.x) {
    const int ci = e / L, s = e - ci * L;
    float acc = 0.f;
    for (int co = 0; co < C; ++co) {
      const float* w = W + (co * C + ci) * K;
      const float* row = dc + co * LP;
      for (int j = 0; j < K; ++j) {
        const int t = s + (K - 1 - j) * dil;
        if (t < L) acc = fmaf(w[j], row[t], acc);
      }
    }
    din[ci * LP + s] += acc;
  }
}
// accW[co][ci][j] += sum_t dc[co][t] * src[ci][t - (K-1-j)*dil]
__device__ void conv_wgrad(float* accW, const float* dc, const float* src, int C, int L, int LP, int K, int dil) {
  for (int e = threadIdx.x; e < C * C * K; e += blockDim.x) {
    const int j = e % K, ci = (e / K) % C, co = e / (K * C);
    const int sh = (K - 1 - j) * dil;
    float acc = 0.f;
    for (int t = sh; t < L; ++t) acc = fmaf(dc[co * LP + t], src[ci * LP + t - sh], acc);
    accW[e] += acc;
  }
}

// ---- register-blocked variants for the kernel sizes the reference uses (ASTGCNN 6, ST_GCN / ST_Conv / STAGNN 2) ----
// The generic loops above spend ~10 instructions per multiply-add (runtime trip counts, two shared-memory loads and a
// bounds test per term: ncu showed ASTGCNN's six k_tcn phases at 1.06 ms of its 1.35 ms step with 15 % of the warps
// active).  Here a thread owns a strip of TB consecutive time steps of one output channel: per input channel it loads the
// K taps and the TB + (K-1) dil input values of the strip once and issues K * TB multiply-adds from registers.
constexpr int kTB = 10;

template <int K, int DIL>
__device__ void causal_conv_f(const float* __restrict__ W, const float* __restrict__ in, float* __restrict__ out, int C,
                              int L, int LP) {
  constexpr int WIN = kTB + (K - 1) * DIL;
  const int nb = (L + kTB - 1) / kTB;
  for (int e = threadIdx.x; e < C * nb; e += blockDim.x) {
    const int co = e / nb, t0 = (e - co * nb) * kTB;
    float acc[kTB];
#pragma unroll
    for (int u = 0; u < kTB; ++u) acc[u] = 0.f;
    const int s0 = t0 - (K - 1) * DIL;
    for (int ci = 0; ci < C; ++ci) {
      const float* w = W + (co * C + ci) * K;
      const float* row = in + ci * LP;
      float wv[K], win[WIN];
#pragma unroll
      for (int j = 0; j < K; ++j) wv[j] = w[j];
#pragma unroll
      for (int q = 0; q < WIN; ++q) { const int s = s0 + q; win[q] = (s >= 0 && s < L) ? row[s] : 0.f; }
#pragma unroll
      for (int u = 0; u < kTB; ++u)
#pragma unroll
        for (int j = 0; j < K; ++j) acc[u] = fmaf(wv[j], win[u + j * DIL], acc[u]);
    }
#pragma unroll
    for (int u = 0; u < kTB; ++u) if (t0 + u < L) out[co * LP + t0 + u] = acc[u];
  }
}
template <int K, int DIL>
__device__ void causal_conv_t_f(const float* __restrict__ W, const float* __restrict__ dc, float* __restrict__ din, int C,
                                int L, int LP) {
  constexpr int WIN = kTB + (K - 1) * DIL;
  const int nb = (L + kTB - 1) / kTB;
  for (int e = threadIdx.x; e < C * nb; e += blockDim.x) {
    const int ci = e / nb, s0 = (e - ci * nb) * kTB;
    float acc[kTB];
#pragma unroll
    for (int u = 0; u < kTB; ++u) acc[u] = 0.f;
    for (int co = 0; co < C; ++co) {
      const float* w = W + (co * C + ci) * K;
      const float* row = dc + co * LP;
      float wv[K], win[WIN];
#pragma unroll
      for (int j = 0; j < K; ++j) wv[j] = w[j];
#pragma unroll
      for (int q = 0; q < WIN; ++q) { const int t = s0 + q; win[q] = t < L ? row[t] : 0.f; }
#pragma unroll
      for (int u = 0; u < kTB; ++u)
#pragma unroll
        for (int j = 0; j < K; ++j) acc[u] = fmaf(wv[j], win[u + (K - 1 - j) * DIL], acc[u]);
    }
#pragma unroll
    for (int u = 0; u < kTB; ++u) if (s0 + u < L) din[ci * LP + s0 + u] += acc[u];
  }
}
// thread (co, ci) owns the K taps of one filter pair; blocks of 8 time steps from registers
template <int K, int DIL>
__device__ void conv_wgrad_f(float* __restrict__ accW, const float* __restrict__ dc, const float* __restrict__ src, int C,
                             int L, int LP) {
  constexpr int TT = 8, WIN = TT + (K - 1) * DIL;
  for (int e = threadIdx.x; e < C * C; e += blockDim.x) {
    const int co = e / C, ci = e - co * C;
    const float* drow = dc + co * LP;
    const float* srow = src + ci * LP;
    float acc[K];
#pragma unroll
    for (int j = 0; j < K; ++j) acc[j] = 0.f;
    for (int t0 = 0; t0 < L; t0 += TT) {
      float d[TT], win[WIN];
#pragma unroll
      for (int u = 0; u < TT; ++u) d[u] = t0 + u < L ? drow[t0 + u] : 0.f;
#pragma unroll
      for (int q = 0; q < WIN; ++q) { const int s = t0 - (K - 1) * DIL + q; win[q] = (s >= 0 && s < L) ? srow[s] : 0.f; }
#pragma unroll
      for (int u = 0; u < TT; ++u)
#pragma unroll
        for (int j = 0; j < K; ++j) acc[j] = fmaf(d[u], win[u + j * DIL], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < K; ++j) accW[e * K + j] += acc[j];
  }
}
// dispatch on the runtime kernel size / dilation
__device__ void conv_any(const float* W, const float* in, float* out, int C, int L, int LP, int K, int dil) {
  if (K == 6 && dil == 1) causal_conv_f<6, 1>(W, in, out, C, L, LP);
  else if (K == 6 && dil == 2) causal_conv_f<6, 2>(W, in, out, C, L, LP);
  else if (K == 2 && dil == 1) causal_conv_f<2, 1>(W, in, out, C, L, LP);
  else if (K == 2 && dil == 2) causal_conv_f<2, 2>(W, in, out, C, L, LP);
  else causal_conv(W, in, out, C, L, LP, K, dil);
}
__device__ void conv_t_any(const float* W, const float* dc, float* din, int C, int L, int LP, int K, int dil) {
  if (K == 6 && dil == 1) causal_conv_t_f<6, 1>(W, dc, din, C, L, LP);
  else if (K == 6 && dil == 2) causal_conv_t_f<6, 2>(W, dc, din, C, L, LP);
  else if (K == 2 && dil == 1) causal_conv_t_f<2, 1>(W, dc, din, C, L, LP);
  else if (K == 2 && dil == 2) causal_conv_t_f<2, 2>(W, dc, din, C, L, LP);
  else causal_conv_t(W, dc, din, C, L, LP, K, dil);
}
__device__ void wgrad_any(float* accW, const float* dc, const float* src, int C, int L, int LP, int K, int dil) {
  if (K == 6 && dil == 1) conv_wgrad_f<6, 1>(accW, dc, src, C, L, LP);
  else if (K == 6 && dil == 2) conv_wgrad_f<6, 2>(accW, dc, src, C, L, LP);
  else if (K == 2 && dil == 1) conv_wgrad_f<2, 1>(accW, dc, src, C, L, LP);
  else if (K == 2 && dil == 2) conv_wgrad_f<2, 2>(accW, dc, src, C, L, LP);
  else conv_wgrad(accW, dc, src, C, L, LP, K, dil);
}

// PH 0..2 forward phases, 3..5 backward phases
template <int PH>
__global__ void __launch_bounds__(kTcnThreads) k_tcn(const TcnArgs a) {
  extern __shared__ float sm[];
  const int C = a.C, L = a.L, K = a.K, LP = L + 1, tid = threadIdx.x, nt = blockDim.x;
  const int NW = C * C * K, PL = C * LP;
  float* W1 = sm;            float* W2 = W1 + NW;
  float* cf1 = W2 + NW;      float* cf2 = cf1 + 4 * C;      // A, Cc, mu, r
  float* q1 = cf2 + 4 * C;   float* q2 = q1 + 2 * C;        // backward means
  float* sacc = q2 + 2 * C;  // [2C] moment partials of this CTA
  // one weight-gradient accumulator, and only in the phase that produces it (B2: dW2, B3: dW1): 512 samples then fit the
  // device in ONE wave of 4 CTAs per SM instead of 1.15 waves of 3
  float* aW1 = sacc + 2 * C; float* aW2 = aW1;
  float* xin = aW1 + ((PH == 4 || PH == 5) ? NW : 0);
  float* c1 = xin + PL;      float* o0 = c1 + PL;
  float* c2 = o0 + PL;       float* da = c2 + PL;           float* db = da + PL;
  const double cnt = (double)a.B * L;
  const double* S1 = a.st;              const double* S2 = a.st + 2 * C;
  double* Q2 = a.st + 4 * C;            double* Q1 = a.st + 6 * C;
  const bool tr = a.training != 0, first = blockIdx.x == 0;

  for (int i = tid; i < NW; i += nt) { W1[i] = a.W1[i]; W2[i] = a.W2[i]; if (PH == 4 || PH == 5) aW1[i] = 0.f; }
  for (int i = tid; i < 2 * C; i += nt) sacc[i] = 0.f;
  if (PH >= 1) bn_coef(cf1, C, tr ? S1 : nullptr, cnt, a.g1, a.be1, a.rm1, a.rv1, a.eps, a.momentum, tr && first && PH == 1);
  if (PH >= 2) bn_coef(cf2, C, tr ? S2 : nullptr, cnt, a.g2, a.be2, a.rm2, a.rv2, a.eps, a.momentum, tr && first && PH == 2);
  if (PH >= 4)
    for (int c = tid; c < C; c += nt) {
      q2[c] = (float)(Q2[c] / cnt); q2[C + c] = (float)(Q2[C + c] / cnt);
      if (PH == 4 && first) { a.dbe2[c] += (float)Q2[c]; a.dg2[c] += (float)Q2[C + c]; }
    }
  if (PH >= 5)
    for (int c = tid; c < C; c += nt) {
      q1[c] = (float)(Q1[c] / cnt); q1[C + c] = (float)(Q1[C + c] / cnt);
      if (first) { a.dbe1[c] += (float)Q1[c]; a.dg1[c] += (float)Q1[C + c]; }
    }
  if (tr && first && tid == 0) {
    if (PH == 1 && a.nbt1) *a.nbt1 += 1;
    if (PH == 2 && a.nbt2) *a.nbt2 += 1;
  }
  __syncthreads();
  const float *A1 = cf1, *C1 = cf1 + C, *mu1 = cf1 + 2 * C, *r1 = cf1 + 3 * C;
  const float *A2 = cf2, *C2 = cf2 + C, *mu2 = cf2 + 2 * C, *r2 = cf2 + 3 * C;

  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    const float* xb = a.x + (size_t)b * C * L;
    for (int e = tid; e < C * L; e += nt) xin[(e / L) * LP + e % L] = xb[e];
    __syncthreads();
    // raw conv outputs: computed once (F1: conv1, F2: conv2) and kept in `sv` for the later phases when the caller gave it
    float* sv1 = a.sv ? a.sv + (size_t)b * 2 * C * L : nullptr;
    float* sv2 = sv1 ? sv1 + C * L : nullptr;
    if (sv1 && PH >= 1) {
      for (int e = tid; e < C * L; e += nt) c1[(e / L) * LP + e % L] = sv1[e];
    } else {
      conv_any(W1, xin, c1, C, L, LP, K, 1);
      if (sv1) {
        __syncthreads();
        for (int e = tid; e < C * L; e += nt) sv1[e] = c1[(e / L) * LP + e % L];
      }
    }
    __syncthreads();
    if (PH == 0) {
      for (int c = tid; c < C; c += nt) {
        float s = 0.f, ss = 0.f;
        for (int t = 0; t < L; ++t) { const float v = c1[c * LP + t]; s += v; ss = fmaf(v, v, ss); }
        sacc[c] += s; sacc[C + c] += ss;
      }
      __syncthreads();
      continue;
    }
    // out_0 = relu(relu(bn1(c1)) + in)
    for (int e = tid; e < C * L; e += nt) {
      const int c = e / L, i = c * LP + e % L;
      o0[i] = fmaxf(fmaxf(fmaf(A1[c], c1[i], C1[c]), 0.f) + xin[i], 0.f);
    }
    __syncthreads();
    if (sv2 && PH >= 2) {
      for (int e = tid; e < C * L; e += nt) c2[(e / L) * LP + e % L] = sv2[e];
    } else {
      conv_any(W2, o0, c2, C, L, LP, K, 2);
      if (sv2) {
        __syncthreads();
        for (int e = tid; e < C * L; e += nt) sv2[e] = c2[(e / L) * LP + e % L];
      }
    }
    __syncthreads();
    if (PH == 1) {
      for (int c = tid; c < C; c += nt) {
        float s = 0.f, ss = 0.f;
        for (int t = 0; t < L; ++t) { const float v = c2[c * LP + t]; s += v; ss = fmaf(v, v, ss); }
        sacc[c] += s; sacc[C + c] += ss;
      }
      __syncthreads();
      continue;
    }
    if (PH == 2) {
      float* ob = a.out + (size_t)b * C * L;
      for (int e = tid; e < C * L; e += nt) {
        const int c = e / L, i = c * LP + e % L;
        ob[e] = fmaxf(fmaxf(fmaf(A2[c], c2[i], C2[c]), 0.f) + o0[i], 0.f);
      }
      __syncthreads();
      continue;
    }
    // ---------------- backward ----------------
    // db <- ds1 = dout * [x1 + out_0 > 0];  da <- dn2 = ds1 * [bn2(c2) > 0]
    const float* dob = a.dout + (size_t)b * C * L;
    for (int e = tid; e < C * L; e += nt) {
      const int c = e / L, i = c * LP + e % L;
      const float pre = fmaf(A2[c], c2[i], C2[c]);
      const float ds1 = (fmaxf(pre, 0.f) + o0[i] > 0.f) ? dob[e] : 0.f;
      db[i] = ds1;
      da[i] = pre > 0.f ? ds1 : 0.f;
    }
    __syncthreads();
    if (PH == 3) {
      for (int c = tid; c < C; c += nt) {
        float s = 0.f, sh = 0.f;
        for (int t = 0; t < L; ++t) {
          const float dn = da[c * LP + t];
          s += dn;
          sh = fmaf(dn, (c2[c * LP + t] - mu2[c]) * r2[c], sh);
        }
        sacc[c] += s; sacc[C + c] += sh;
      }
      __syncthreads();
      continue;
    }
    // da <- dc2 = A2 (dn2 - q2a - c2hat q2b)
    for (int e = tid; e < C * L; e += nt) {
      const int c = e / L, i = c * LP + e % L;
      da[i] = A2[c] * (da[i] - q2[c] - (c2[i] - mu2[c]) * r2[c] * q2[C + c]);
    }
    __syncthreads();
    if (PH == 4) wgrad_any(aW2, da, o0, C, L, LP, K, 2);
    conv_t_any(W2, da, db, C, L, LP, K, 2);              // db <- d(out_0) = ds1 + conv2^T(dc2)
    __syncthreads();
    // db <- ds0 = d(out_0) * [x0 + in > 0];  da <- dn1 = ds0 * [bn1(c1) > 0]
    for (int e = tid; e < C * L; e += nt) {
      const int c = e / L, i = c * LP + e % L;
      const float pre = fmaf(A1[c], c1[i], C1[c]);
      const float ds0 = (fmaxf(pre, 0.f) + xin[i] > 0.f) ? db[i] : 0.f;
      db[i] = ds0;
      da[i] = pre > 0.f ? ds0 : 0.f;
    }
    __syncthreads();
    if (PH == 4) {
      for (int c = tid; c < C; c += nt) {
        float s = 0.f, sh = 0.f;
        for (int t = 0; t < L; ++t) {
          const float dn = da[c * LP + t];
          s += dn;
          sh = fmaf(dn, (c1[c * LP + t] - mu1[c]) * r1[c], sh);
        }
        sacc[c] += s; sacc[C + c] += sh;
      }
      __syncthreads();
      continue;
    }
    // PH == 5: da <- dc1; dW1; dx = ds0 + conv1^T(dc1)
    for (int e = tid; e < C * L; e += nt) {
      const int c = e / L, i = c * LP + e % L;
      da[i] = A1[c] * (da[i] - q1[c] - (c1[i] - mu1[c]) * r1[c] * q1[C + c]);
    }
    __syncthreads();
    wgrad_any(aW1, da, xin, C, L, LP, K, 1);
    conv_t_any(W1, da, db, C, L, LP, K, 1);
    __syncthreads();
    float* dxb = a.dx + (size_t)b * C * L;
    for (int e = tid; e < C * L; e += nt) dxb[e] = db[(e / L) * LP + e % L];
    __syncthreads();
  }
  __syncthreads();
  if (PH == 0 || PH == 1 || PH == 3 || PH == 4) {
    double* dst = PH == 0 ? a.st : PH == 1 ? a.st + 2 * C : PH == 3 ? Q2 : Q1;
    for (int i = tid; i < 2 * C; i += nt) atomicAdd(&dst[i], (double)sacc[i]);
  }
  if (PH == 4) for (int i = tid; i < NW; i += nt) atomicAdd(&a.dW2[i], aW2[i]);
  if (PH == 5) for (int i = tid; i < NW; i += nt) atomicAdd(&a.dW1[i], aW1[i]);
}

size_t tcn_smem(int C, int L, int K, bool wgrad = true) {
  const size_t NW = (size_t)C * C * K, PL = (size_t)C * (L + 1);
  return ((wgrad ? 3 : 2) * NW + 14 * (size_t)C + 6 * PL) * 4;
}
bool g_tcn_attr = false;
void tcn_attrs() {
  if (g_tcn_attr) return;
  cudaFuncSetAttribute(k_tcn<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_tcn<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_tcn<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_tcn<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_tcn<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_tcn<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  g_tcn_attr = true;
}
int tcn_grid(int B, size_t smem) {
  int per_sm = (int)((226 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  int g = 148 * per_sm;
  return g > B ? B : g;
}

int fill(TcnArgs& a, const float* x, int B, int C, int L, int K, const stg_tcn_params* p, int training, float momentum,
         float eps, double* scratch) {
  if (!x || !p || !scratch || B < 1 || C < 1 || L < 1 || K < 1) return set_err(STG_ERR_INVALID, "bad argument");
  if (!p->conv1_w || !p->conv2_w || !p->bn1.weight || !p->bn1.bias || !p->bn1.running_mean || !p->bn1.running_var ||
      !p->bn2.weight || !p->bn2.bias || !p->bn2.running_mean || !p->bn2.running_var)
    return set_err(STG_ERR_INVALID, "null TCN parameter pointer");
  if (tcn_smem(C, L, K) > 200 * 1024)
    return set_err(STG_ERR_UNSUPPORTED, "TCN tile C=%d L=%d K=%d does not fit shared memory", C, L, K);
  memset(&a, 0, sizeof(a));
  a.B = B; a.C = C; a.L = L; a.K = K; a.x = x;
  a.W1 = p->conv1_w; a.W2 = p->conv2_w;
  a.g1 = p->bn1.weight; a.be1 = p->bn1.bias; a.rm1 = p->bn1.running_mean; a.rv1 = p->bn1.running_var;
  a.g2 = p->bn2.weight; a.be2 = p->bn2.bias; a.rm2 = p->bn2.running_mean; a.rv2 = p->bn2.running_var;
  a.nbt1 = (long long*)p->bn1.num_batches_tracked; a.nbt2 = (long long*)p->bn2.num_batches_tracked;
  a.training = training; a.momentum = momentum; a.eps = eps; a.st = scratch;
  return STG_OK;
}

}  // namespace
}  // namespace stg

using namespace stg;

extern "C" int stg_tcn_forward(const float* x_dev, int B, int C, int L, int K, const stg_tcn_params* params, int training,
                               float momentum, float eps, double* scratch_dev, float* saved_dev, float* out_dev,
                               void* stream) {
  TcnArgs a;
  int rc = fill(a, x_dev, B, C, L, K, params, training, momentum, eps, scratch_dev);
  if (rc) return rc;
  if (!out_dev) return set_err(STG_ERR_INVALID, "null out");
  a.out = out_dev;
  a.sv = training ? saved_dev : nullptr;
  tcn_attrs();
  cudaStream_t s = (cudaStream_t)stream;
  const size_t smem = tcn_smem(C, L, K, false);
  const int grid = tcn_grid(B, smem);
  if (training) {
    cudaMemsetAsync(scratch_dev, 0, sizeof(double) * 8 * C, s);
    k_tcn<0><<<grid, kTcnThreads, smem, s>>>(a);
    k_tcn<1><<<grid, kTcnThreads, smem, s>>>(a);
  }
  k_tcn<2><<<grid, kTcnThreads, smem, s>>>(a);
  return check_cuda("stg_tcn_forward");
}

extern "C" int stg_tcn_backward(const float* x_dev, const float* dout_dev, int B, int C, int L, int K,
                                const stg_tcn_params* params, const stg_tcn_params* grads, float eps,
                                double* scratch_dev, const float* saved_dev, float* dx_dev, void* stream) {
  TcnArgs a;
  int rc = fill(a, x_dev, B, C, L, K, params, 1, 0.f, eps, scratch_dev);
  if (rc) return rc;
  if (!dout_dev || !dx_dev || !grads || !grads->conv1_w || !grads->conv2_w || !grads->bn1.weight || !grads->bn1.bias ||
      !grads->bn2.weight || !grads->bn2.bias)
    return set_err(STG_ERR_INVALID, "null gradient pointer");
  a.dout = dout_dev; a.dx = dx_dev;
  a.sv = const_cast<float*>(saved_dev);
  a.dW1 = grads->conv1_w; a.dW2 = grads->conv2_w;
  a.dg1 = grads->bn1.weight; a.dbe1 = grads->bn1.bias; a.dg2 = grads->bn2.weight; a.dbe2 = grads->bn2.bias;
  tcn_attrs();
  cudaStream_t s = (cudaStream_t)stream;
  const size_t smem = tcn_smem(C, L, K), smem3 = tcn_smem(C, L, K, false);
  const int grid = tcn_grid(B, smem);
  k_tcn<3><<<tcn_grid(B, smem3), kTcnThreads, smem3, s>>>(a);
  k_tcn<4><<<grid, kTcnThreads, smem, s>>>(a);
  k_tcn<5><<<grid, kTcnThreads, smem, s>>>(a);
  return check_cuda("stg_tcn_backward");
}
