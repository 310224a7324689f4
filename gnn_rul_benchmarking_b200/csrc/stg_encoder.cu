// Patch encoder + positional encoding of FC_STGNN_RUL, forward and backward (sm_100a).
//   Feature_extractor_1DCNN_RUL (Model_Base.py:12-41), nonlin_map2 (Model.py:18-22,55-59),
//   PositionalEncoding + Dropout (Model_Base.py:111-134, Model.py:62-68).
//
// One thread owns one row r = (b,t,n) (a patch of P samples of one sensor).  Everything a row needs
// is recomputed from X in every phase (X is 4*N*L bytes per window -- the smallest tensor on the
// path), so nothing but the BatchNorm moments crosses a phase boundary:
//   training forward : F1 conv1 moments -> F2 conv2 moments -> F3 linear moments -> F4 write h
//   eval forward     : F4 only (running statistics)
//   backward         : B1 BN3 sums -> B2 dW3, db3, BN2 sums -> B3 dW2, BN1 sums -> B4 dW1
// Per-row activations live in shared memory as [feature][row] columns (conflict-free for the
// row-owner; odd row pitch makes the weight-gradient outer products conflict-free too), so all
// dimensions are runtime values: every FC_STGNN hyper-parameter set of configs/hparams.py runs.
#include <math.h>
#include <stdio.h>

#include "stg_model.cuh"

namespace stg {
namespace {

constexpr int kThreads = 256;
constexpr size_t kSmemCap = 200 * 1024;

struct Lay {          // offsets in floats from the dynamic smem base
  int W1, W2, W3, b3, coef, q, pe, accW, sacc, xs, c1, c2, z, d2, d1, total;
  int TRP;
};

__host__ __device__ inline int imax(int a, int b) { return a > b ? a : b; }

__host__ __device__ inline Lay make_lay(const EncArgs& a, bool bwd) {
  Lay l;
  const int TRP = a.TR + 1;
  l.TRP = TRP;
  int o = 0;
  l.W1 = o; o += a.EH * a.K;
  l.W2 = o; o += a.E * a.EH * a.K;
  l.W3 = o; o += a.C * a.EL2;
  l.b3 = o; o += a.C;
  l.coef = o; o += 4 * (a.EH + a.E + a.C);       // per BN: A, Cc, mu, r
  l.q = o; o += 2 * (a.EH + a.E + a.C);          // per BN: backward means
  l.pe = o; o += a.T * a.C;
  l.accW = o; o += imax(imax(a.EH * a.K, a.E * a.EH * a.K), a.C * a.EL2 + a.C);
  o = (o + 1) & ~1;
  l.sacc = o; o += 2 * 2 * imax(imax(a.EH, a.E), a.C);   // doubles
  l.xs = o; o += a.P * TRP;
  l.c1 = o; o += a.EH * a.L1 * TRP;
  l.c2 = o; o += a.EL2 * TRP;
  l.z = o; o += a.C * TRP;
  l.d2 = o; l.d1 = o;
  if (bwd) {
    o += a.EL2 * TRP;
    l.d1 = o; o += a.EH * a.L1 * TRP;
  }
  l.total = o;
  return l;
}

STG_DEVINL float keep_scale(const EncArgs& a, size_t idx) {
  if (!a.training || a.pdrop <= 0.f) return 1.f;
  const float sc = 1.f / (1.f - a.pdrop);
  if (a.keep) return a.keep[idx] * sc;
  unsigned long long z = a.seed + (a.seed_ptr ? (unsigned long long)*a.seed_ptr * 0xD1B54A32D192ED03ull : 0ull) +
                         (unsigned long long)idx * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  const float u = (float)(z >> 40) * (1.f / 16777216.f);
  return u >= a.pdrop ? sc : 0.f;
}

// BN forward coefficients of one layer into smem: A = g*r, Cc = beta - A*mu, mu, r.
// Threads [0,n).  stats: sums / sums of squares (training) or nullptr (running statistics).
STG_DEVINL void bn_coefs(float* dst, int n, const double* stats, double count, const float* g, const float* be,
                         float* rm, float* rv, float eps, float momentum, bool update) {
  for (int c = threadIdx.x; c < n; c += blockDim.x) {
    double m, var;
    if (stats) {
      m = stats[c] / count;
      var = stats[n + c] / count - m * m;
      if (var < 0.0) var = 0.0;
      if (update) {
        const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
        rm[c] = (1.f - momentum) * rm[c] + momentum * (float)m;
        rv[c] = (1.f - momentum) * rv[c] + momentum * (float)unb;
      }
    } else {
      m = rm[c];
      var = rv[c];
    }
    const float r = (float)(1.0 / sqrt(var + (double)eps));
    const float A = g[c] * r;
    dst[c] = A;
    dst[n + c] = be[c] - A * (float)m;
    dst[2 * n + c] = (float)m;
    dst[3 * n + c] = r;
  }
}

STG_DEVINL void stat_add(double* sacc, int idx, float v) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) atomicAdd(&sacc[idx], (double)v);
}

// accW[pair] += sum over the tile's rows of f(pair, row); npairs small -> rows are sliced over threads.
template <typename Fn>
STG_DEVINL void pair_reduce(float* accW, int npairs, int TR, Fn f) {
  const int tid = threadIdx.x;
  if (npairs <= kThreads / 2) {
    const int slices = kThreads / npairs;
    if (tid < npairs * slices) {
      const int pair = tid % npairs, sl = tid / npairs;
      float acc = 0.f;
      for (int r = sl; r < TR; r += slices) acc += f(pair, r);
      atomicAdd(&accW[pair], acc);
    }
  } else {
    for (int pair = tid; pair < npairs; pair += kThreads) {
      float acc = 0.f;
      for (int r = 0; r < TR; ++r) acc += f(pair, r);
      accW[pair] += acc;
    }
  }
}

// PH: 0..3 forward phases F1..F4, 4..7 backward phases B1..B4.
template <int PH>
__global__ void __launch_bounds__(kThreads) k_encoder(const EncArgs a) {
  extern __shared__ __align__(16) float sm[];
  constexpr bool BWD = PH >= 4;
  const Lay l = make_lay(a, BWD);
  const int TRP = l.TRP, TR = a.TR, tid = threadIdx.x;
  const int EH = a.EH, E = a.E, C = a.C, K = a.K, P = a.P, L1 = a.L1, L2 = a.L2, EL2 = a.EL2, N = a.N, T = a.T;
  float *W1 = sm + l.W1, *W2 = sm + l.W2, *W3 = sm + l.W3, *b3 = sm + l.b3;
  float *cf1 = sm + l.coef, *cf2 = cf1 + 4 * EH, *cf3 = cf2 + 4 * E;
  float *q1 = sm + l.q, *q2 = q1 + 2 * EH, *q3 = q2 + 2 * E;
  float *pe = sm + l.pe, *accW = sm + l.accW;
  double* sacc = reinterpret_cast<double*>(sm + l.sacc);
  float *xs = sm + l.xs, *c1 = sm + l.c1, *c2 = sm + l.c2, *zz = sm + l.z, *d2 = sm + l.d2, *d1 = sm + l.d1;

  const double cnt1 = (double)a.R * L1, cnt2 = (double)a.R * L2, cnt3 = (double)a.R;
  const double* S1 = a.st;                     // forward moments
  const double* S2 = S1 + 2 * EH;
  const double* S3 = S2 + 2 * E;
  double* Bq3 = a.st + 2 * (EH + E + C);       // backward sums
  double* Bq2 = Bq3 + 2 * C;
  double* Bq1 = Bq2 + 2 * E;

  // ---------------- prologue: weights, BN coefficients ----------------
  for (int i = tid; i < EH * K; i += kThreads) W1[i] = a.W1[i];
  if (PH >= 1) for (int i = tid; i < E * EH * K; i += kThreads) W2[i] = a.W2[i];
  if (PH >= 2) {
    for (int i = tid; i < C * EL2; i += kThreads) W3[i] = a.W3[i];
    for (int i = tid; i < C; i += kThreads) b3[i] = a.b3[i];
  }
  const bool first = blockIdx.x == 0;
  const bool tr = a.training != 0;
  if (PH >= 1) bn_coefs(cf1, EH, tr ? S1 : nullptr, cnt1, a.g1, a.be1, a.rm1, a.rv1, a.eps, a.momentum, tr && first && PH == 1);
  if (PH >= 2) bn_coefs(cf2, E, tr ? S2 : nullptr, cnt2, a.g2, a.be2, a.rm2, a.rv2, a.eps, a.momentum, tr && first && PH == 2);
  if (PH >= 3) bn_coefs(cf3, C, tr ? S3 : nullptr, cnt3, a.g3, a.be3, a.rm3, a.rv3, a.eps, a.momentum, tr && first && PH == 3);
  if (PH == 3) for (int i = tid; i < T * C; i += kThreads) pe[i] = a.pe[i];
  if (PH >= 5) for (int c = tid; c < C; c += kThreads) {
    q3[c] = (float)(Bq3[c] / cnt3);
    q3[C + c] = (float)(Bq3[C + c] / cnt3);
    if (PH == 5 && first) { a.dbe3[c] += (float)Bq3[c]; a.dg3[c] += (float)Bq3[C + c]; }
  }
  if (PH >= 6) for (int c = tid; c < E; c += kThreads) {
    q2[c] = (float)(Bq2[c] / cnt2);
    q2[E + c] = (float)(Bq2[E + c] / cnt2);
    if (PH == 6 && first) { a.dbe2[c] += (float)Bq2[c]; a.dg2[c] += (float)Bq2[E + c]; }
  }
  if (PH >= 7) for (int c = tid; c < EH; c += kThreads) {
    q1[c] = (float)(Bq1[c] / cnt1);
    q1[EH + c] = (float)(Bq1[EH + c] / cnt1);
    if (first) { a.dbe1[c] += (float)Bq1[c]; a.dg1[c] += (float)Bq1[EH + c]; }
  }
  constexpr int kNPairsNone = 0;
  const int npairs = PH == 5 ? C * EL2 + C : PH == 6 ? E * EH * K : PH == 7 ? EH * K : kNPairsNone;
  for (int i = tid; i < npairs; i += kThreads) accW[i] = 0.f;
  const int nstat = PH == 0 ? EH : PH == 1 ? E : PH == 2 ? C : PH == 4 ? C : PH == 5 ? E : PH == 6 ? EH : 0;
  for (int i = tid; i < 2 * nstat; i += kThreads) sacc[i] = 0.0;
  __syncthreads();

  const float *A1 = cf1, *C1 = cf1 + EH, *mu1 = cf1 + 2 * EH, *r1 = cf1 + 3 * EH;
  const float *A2 = cf2, *C2 = cf2 + E, *mu2 = cf2 + 2 * E, *r2 = cf2 + 3 * E;
  const float *A3 = cf3, *C3 = cf3 + C, *mu3 = cf3 + 2 * C, *r3 = cf3 + 3 * C;

  const int ntiles = (a.R + TR - 1) / TR;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int r = tile * TR + tid;
    const bool act = tid < TR && r < a.R;
    int n = 0, t = 0, b = 0;
    if (act) {
      n = r % N;
      const int bt = r / N;
      t = bt % T;
      b = bt / T;
      // ---- x, conv1 (raw, pre-BN) ----
      const float* xp = a.X + ((size_t)(b * N + n) * T + t) * P;
      for (int i = 0; i < P; ++i) xs[i * TRP + tid] = xp[i];
      for (int ch = 0; ch < EH; ++ch)
        for (int p = 0; p < L1; ++p) {
          float acc = 0.f;
          for (int j = 0; j < K; ++j) {
            const int q = p + j - a.pad1;
            if (q >= 0 && q < P) acc = fmaf(W1[ch * K + j], xs[q * TRP + tid], acc);
          }
          c1[(ch * L1 + p) * TRP + tid] = acc;
        }
    }
    if (PH == 0) {
      for (int ch = 0; ch < EH; ++ch) {
        float s = 0.f, ss = 0.f;
        if (act)
          for (int p = 0; p < L1; ++p) {
            const float v = c1[(ch * L1 + p) * TRP + tid];
            s += v;
            ss = fmaf(v, v, ss);
          }
        stat_add(sacc, ch, s);
        stat_add(sacc, EH + ch, ss);
      }
      continue;
    }
    // ---- conv2 (raw) on a1 = relu(BN1(c1)) ----
    if (act)
      for (int e = 0; e < E; ++e)
        for (int p = 0; p < L2; ++p) {
          float acc = 0.f;
          for (int ch = 0; ch < EH; ++ch) {
            const float Ac = A1[ch], Cc = C1[ch];
            for (int j = 0; j < K; ++j) {
              const int q = p + j - 1;
              if (q >= 0 && q < L1) {
                const float av = fmaxf(fmaf(Ac, c1[(ch * L1 + q) * TRP + tid], Cc), 0.f);
                acc = fmaf(W2[(e * EH + ch) * K + j], av, acc);
              }
            }
          }
          c2[(e * L2 + p) * TRP + tid] = acc;
        }
    if (PH == 1) {
      for (int e = 0; e < E; ++e) {
        float s = 0.f, ss = 0.f;
        if (act)
          for (int p = 0; p < L2; ++p) {
            const float v = c2[(e * L2 + p) * TRP + tid];
            s += v;
            ss = fmaf(v, v, ss);
          }
        stat_add(sacc, e, s);
        stat_add(sacc, E + e, ss);
      }
      continue;
    }
    // ---- linear (raw z3) on a2 = relu(BN2(c2)) ----
    if (act)
      for (int cb = 0; cb < C; cb += 4) {
        float acc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] = (cb + u < C) ? b3[cb + u] : 0.f;
        for (int k = 0; k < EL2; ++k) {
          const int e = k / L2;
          const float av = fmaxf(fmaf(A2[e], c2[k * TRP + tid], C2[e]), 0.f);
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (cb + u < C) acc[u] = fmaf(W3[(cb + u) * EL2 + k], av, acc[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (cb + u < C) zz[(cb + u) * TRP + tid] = acc[u];
      }
    if (PH == 2) {
      for (int c = 0; c < C; ++c) {
        const float v = act ? zz[c * TRP + tid] : 0.f;
        stat_add(sacc, c, v);
        stat_add(sacc, C + c, v * v);
      }
      continue;
    }
    if (PH == 3) {
      // ---- BN3 + positional encoding + dropout -> h[b,t,n,:] ----
      if (act) {
        float* hr = a.h + (size_t)r * C;
        const size_t kbase = ((size_t)(b * N + n) * T + t) * C;     // reference dropout layout [B*N,T,C]
        for (int c = 0; c < C; ++c) {
          const float hn = fmaf(A3[c], zz[c * TRP + tid], C3[c]) + pe[t * C + c];
          hr[c] = hn * keep_scale(a, kbase + c);
        }
      }
      continue;
    }
    if (BWD) {
      // ---- dhn = dropout'(dh);  BN3 backward ----
      const size_t kbase = ((size_t)(b * N + n) * T + t) * C;
      for (int c = 0; c < C; ++c) {
        float dhn = 0.f, zh = 0.f;
        if (act) {
          dhn = a.dh[(size_t)r * C + c] * keep_scale(a, kbase + c);
          zh = (zz[c * TRP + tid] - mu3[c]) * r3[c];
        }
        if (PH == 4) {
          stat_add(sacc, c, dhn);
          stat_add(sacc, C + c, dhn * zh);
        } else if (tid < TR) {
          zz[c * TRP + tid] = act ? A3[c] * (dhn - q3[c] - zh * q3[C + c]) : 0.f;   // dz3
        }
      }
      if (PH == 4) continue;
      // ---- linear backward: da2 -> dn2 (ReLU) ----
      if (tid < TR)
        for (int k = 0; k < EL2; ++k) {
          float v = 0.f;
          if (act) {
            const int e = k / L2;
            if (fmaf(A2[e], c2[k * TRP + tid], C2[e]) > 0.f) {
              for (int c = 0; c < C; ++c) v = fmaf(zz[c * TRP + tid], W3[c * EL2 + k], v);
            }
          }
          d2[k * TRP + tid] = v;
        }
      if (PH == 5) {
        for (int e = 0; e < E; ++e) {
          float s = 0.f, sh = 0.f;
          if (act)
            for (int p = 0; p < L2; ++p) {
              const float dn = d2[(e * L2 + p) * TRP + tid];
              s += dn;
              sh = fmaf(dn, (c2[(e * L2 + p) * TRP + tid] - mu2[e]) * r2[e], sh);
            }
          stat_add(sacc, e, s);
          stat_add(sacc, E + e, sh);
        }
        __syncthreads();
        // dW3[c][k] += sum_r dz3[c][r] * a2[k][r];  db3[c] += sum_r dz3[c][r]
        const int rows = min(TR, a.R - tile * TR);
        pair_reduce(accW, C * EL2 + C, rows, [&](int pair, int rr) {
          if (pair >= C * EL2) return zz[(pair - C * EL2) * TRP + rr];
          const int c = pair / EL2, k = pair - c * EL2, e = k / L2;
          const float av = fmaxf(fmaf(A2[e], c2[k * TRP + rr], C2[e]), 0.f);
          return zz[c * TRP + rr] * av;
        });
        __syncthreads();
        continue;
      }
      // ---- BN2 backward (in place) -> dc2 ----
      if (tid < TR)
        for (int k = 0; k < EL2; ++k) {
          const int e = k / L2;
          float v = 0.f;
          if (act) {
            const float ch2 = (c2[k * TRP + tid] - mu2[e]) * r2[e];
            v = A2[e] * (d2[k * TRP + tid] - q2[e] - ch2 * q2[E + e]);
          }
          d2[k * TRP + tid] = v;
        }
      // ---- conv2 backward wrt its input: da1 -> dn1 (ReLU) ----
      if (tid < TR)
        for (int ch = 0; ch < EH; ++ch)
          for (int q = 0; q < L1; ++q) {
            float v = 0.f;
            if (act && fmaf(A1[ch], c1[(ch * L1 + q) * TRP + tid], C1[ch]) > 0.f) {
              for (int e = 0; e < E; ++e)
                for (int j = 0; j < K; ++j) {
                  const int p = q - j + 1;
                  if (p >= 0 && p < L2) v = fmaf(d2[(e * L2 + p) * TRP + tid], W2[(e * EH + ch) * K + j], v);
                }
            }
            d1[(ch * L1 + q) * TRP + tid] = v;
          }
      if (PH == 6) {
        for (int ch = 0; ch < EH; ++ch) {
          float s = 0.f, sh = 0.f;
          if (act)
            for (int p = 0; p < L1; ++p) {
              const float dn = d1[(ch * L1 + p) * TRP + tid];
              s += dn;
              sh = fmaf(dn, (c1[(ch * L1 + p) * TRP + tid] - mu1[ch]) * r1[ch], sh);
            }
          stat_add(sacc, ch, s);
          stat_add(sacc, EH + ch, sh);
        }
        __syncthreads();
        // dW2[e][ch][j] += sum_r sum_p dc2[e][p][r] * a1[ch][p+j-1][r]
        const int rows = min(TR, a.R - tile * TR);
        pair_reduce(accW, E * EH * K, rows, [&](int pair, int rr) {
          const int j = pair % K, ech = pair / K, ch = ech % EH, e = ech / EH;
          float acc = 0.f;
          for (int p = 0; p < L2; ++p) {
            const int q = p + j - 1;
            if (q >= 0 && q < L1) {
              const float av = fmaxf(fmaf(A1[ch], c1[(ch * L1 + q) * TRP + rr], C1[ch]), 0.f);
              acc = fmaf(d2[(e * L2 + p) * TRP + rr], av, acc);
            }
          }
          return acc;
        });
        __syncthreads();
        continue;
      }
      // ---- PH == 7: BN1 backward -> dc1; dW1[ch][j] += sum_r sum_p dc1[ch][p][r] * xpad[p+j-pad1][r]
      if (tid < TR)
        for (int ch = 0; ch < EH; ++ch)
          for (int p = 0; p < L1; ++p) {
            float v = 0.f;
            if (act) {
              const float ch1 = (c1[(ch * L1 + p) * TRP + tid] - mu1[ch]) * r1[ch];
              v = A1[ch] * (d1[(ch * L1 + p) * TRP + tid] - q1[ch] - ch1 * q1[EH + ch]);
            }
            d1[(ch * L1 + p) * TRP + tid] = v;
          }
      __syncthreads();
      const int rows = min(TR, a.R - tile * TR);
      pair_reduce(accW, EH * K, rows, [&](int pair, int rr) {
        const int j = pair % K, ch = pair / K;
        float acc = 0.f;
        for (int p = 0; p < L1; ++p) {
          const int q = p + j - a.pad1;
          if (q >= 0 && q < P) acc = fmaf(d1[(ch * L1 + p) * TRP + rr], xs[q * TRP + rr], acc);
        }
        return acc;
      });
      __syncthreads();
    }
  }
  __syncthreads();
  // ---------------- epilogue: flush CTA accumulators ----------------
  if (nstat) {
    double* dst = PH == 0 ? a.st : PH == 1 ? a.st + 2 * EH : PH == 2 ? a.st + 2 * (EH + E) : PH == 4 ? Bq3 : PH == 5 ? Bq2 : Bq1;
    for (int i = tid; i < 2 * nstat; i += kThreads) atomicAdd(&dst[i], sacc[i]);
  }
  if (PH == 5) {
    for (int i = tid; i < C * EL2; i += kThreads) atomicAdd(&a.dW3[i], accW[i]);
    for (int i = tid; i < C; i += kThreads) atomicAdd(&a.db3[i], accW[C * EL2 + i]);
  } else if (PH == 6) {
    for (int i = tid; i < E * EH * K; i += kThreads) atomicAdd(&a.dW2[i], accW[i]);
  } else if (PH == 7) {
    for (int i = tid; i < EH * K; i += kThreads) atomicAdd(&a.dW1[i], accW[i]);
  }
}

typedef void (*EncKernel)(const EncArgs);
const EncKernel kKernels[8] = {k_encoder<0>, k_encoder<1>, k_encoder<2>, k_encoder<3>,
                               k_encoder<4>, k_encoder<5>, k_encoder<6>, k_encoder<7>};
const int kProfOf[8] = {kProfEncF1, kProfEncF2, kProfEncF3, kProfEncF4, kProfEncB1, kProfEncB2, kProfEncB3, kProfEncB4};
bool g_attr[64] = {};

void set_attrs() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (g_attr[dev]) return;
  for (int i = 0; i < 8; ++i)
    cudaFuncSetAttribute(kKernels[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemCap);
  g_attr[dev] = true;
}

int num_sms() {
  static int sms[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!sms[dev]) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    sms[dev] = v;
  }
  return sms[dev];
}

void launch_phase(int ph, const EncArgs& a, size_t smem, cudaStream_t s) {
  const int ntiles = (a.R + a.TR - 1) / a.TR;
  // persistent CTAs: as many as fit (shared memory bound), never more than tiles
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  int grid = num_sms() * per_sm;
  if (grid > ntiles) grid = ntiles;
  ProfScope ps(kProfOf[ph], s);
  kKernels[ph]<<<grid, kThreads, smem, s>>>(a);
}

}  // namespace

int plan_encoder(EncArgs& a, size_t* smem_fwd, size_t* smem_bwd, char* err, size_t errlen) {
  if (a.K < 1 || a.P < 1 || a.EH < 1 || a.E < 1 || a.C < 1) { snprintf(err, errlen, "encoder: non-positive dimension"); return -1; }
  a.pad1 = a.K / 2;
  a.L1 = a.P + 2 * a.pad1 - a.K + 1;
  a.L2 = a.L1 + 2 - a.K + 1;
  if (a.L1 < 1 || a.L2 < 1) {
    snprintf(err, errlen, "encoder: patch_size %d too short for conv kernel %d", a.P, a.K);
    return -1;
  }
  a.EL2 = a.E * a.L2;
  a.R = a.B * a.T * a.N;
  for (int tr = 256; tr >= 8; tr >>= 1) {
    a.TR = tr;
    const size_t sb = (size_t)make_lay(a, true).total * 4;
    if (sb <= kSmemCap) {
      *smem_fwd = (size_t)make_lay(a, false).total * 4;
      *smem_bwd = sb;
      return 0;
    }
  }
  snprintf(err, errlen, "encoder tile does not fit shared memory (P=%d K=%d EH=%d E=%d C=%d)", a.P, a.K, a.EH, a.E, a.C);
  return -2;
}

int launch_encoder_forward(const EncArgs& a, size_t smem, cudaStream_t s) {
  if (a.c2raw && encoder_fast_available(a)) return launch_encoder_fast(a, false, s);
  set_attrs();
  if (a.training)
    for (int ph = 0; ph < 3; ++ph) launch_phase(ph, a, smem, s);
  launch_phase(3, a, smem, s);
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

int launch_encoder_backward(const EncArgs& a, size_t smem, cudaStream_t s) {
  if (a.c2raw && encoder_fast_available(a)) return launch_encoder_fast(a, true, s);
  set_attrs();
  for (int ph = 4; ph < 8; ++ph) launch_phase(ph, a, smem, s);
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

}  // namespace stg
