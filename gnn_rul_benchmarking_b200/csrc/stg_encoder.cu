// Patch encoder + positional encoding of FC_STGNN_RUL, forward and backward (sm_100a), any dimensions.
//   Feature_extractor_1DCNN_RUL (Model_Base.py:12-41), nonlin_map2 (Model.py:18-22,55-59),
//   PositionalEncoding + Dropout (Model_Base.py:111-134, Model.py:62-68).
//
// A row r = (b,t,n) is a patch of P samples of one sensor.  A CTA works on a tile of TR rows held in
// shared memory as [feature][row] columns (odd pitch), and every stage -- conv1, conv2, the linear map,
// their transposes and the weight-gradient outer products -- is spread over all 256 threads as
// (feature, row) work items, the long contraction of the linear map additionally split over lanes.  So a
// set with few, wide rows (FD001: 7168 rows x 864 conv2 features at batch 256) keeps every thread
// busy, as does one with many narrow rows.  Everything a row needs is recomputed from X in every phase (X is
// 4*N*L bytes per window -- the smallest tensor on the path), so nothing but the BatchNorm moments
// crosses a phase boundary:
//   training forward : F1 conv1 moments -> F2 conv2 moments -> F3 linear moments -> F4 write h
//   eval forward     : F4 only (running statistics)
//   backward         : B1 BN3 sums -> B2 dW3, db3, BN2 sums -> B3 dW2, BN1 sums -> B4 dW1
// All dimensions are runtime values and TR is chosen per phase from what that phase keeps in shared
// memory: every FC_STGNN hyper-parameter set of configs/hparams.py runs.  The sets with register-sized rows
// (FD002/FD003/FD004/S2) take the compile-time path in stg_encoder_fast.cu instead.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "stg_model.cuh"

namespace stg {
namespace {

constexpr int kThreads = 1024;       // one CTA per SM (shared-memory bound): 32 warps hide the LDS->FMA chains
constexpr int kWarps = kThreads / 32;
constexpr size_t kSmemCap = 200 * 1024;
constexpr int kMinTR = 4;

struct Lay {          // offsets in floats from the dynamic smem base
  int W1, W2, W3, b3, coef, q, pe, accW, sacc, meta, xs, c1, a1, c2, z, z2, d2, d1, total;
  int TRP;
};

__host__ __device__ inline int imax(int a, int b) { return a > b ? a : b; }

__host__ __device__ inline int enc_npairs(const EncArgs& a, int ph) {
  return ph == 5 ? a.C * a.EL2 + a.C : ph == 6 ? a.E * a.EH * a.K : ph == 7 ? a.EH * a.K : 0;
}

// What phase `ph` computes / keeps in shared memory.  In training the raw conv2 / linear outputs and the
// masked gradients dn2 / dn1 stay in the workspace between phases (a.c2raw ...), so each heavy stage runs
// once per step; in eval (one phase, running statistics) everything is computed in place.
struct Use { bool x, a1, conv2, c2, lin, z, z2, d2, d1, W2, W3; };

__host__ __device__ inline bool enc_kept(const EncArgs& a) { return a.training && a.c2raw != nullptr; }

__host__ __device__ inline Use enc_use(const EncArgs& a, int ph) {
  const bool k = enc_kept(a);
  Use u;
  u.x = !k || ph <= 1 || ph >= 6;                       // x and conv1 (cheap: always recomputed where needed)
  u.a1 = k ? (ph == 1 || ph == 6) : ph >= 1;
  u.conv2 = k ? ph == 1 : ph >= 1;                      // compute conv2 here
  u.c2 = u.conv2 || (k && (ph == 2 || ph == 5 || ph == 6));     // tile holds conv2 (computed or loaded)
  u.lin = k ? ph == 2 : ph >= 2;                        // compute the linear map here
  u.z = u.lin || (k && ph >= 3 && ph <= 5);
  u.z2 = k ? (ph == 4 || ph == 5) : ph >= 4;            // dz3 and the linear map's transpose happen here
  u.d2 = k ? (ph == 5 || ph == 6) : ph >= 5;
  u.d1 = ph >= 6;
  u.W2 = u.conv2 || (u.d2 && ph >= 6);
  u.W3 = u.lin || (u.z2 && ph >= 5);
  return u;
}

__host__ __device__ inline Lay make_lay(const EncArgs& a, int ph) {
  const Use u = enc_use(a, ph);
  Lay l;
  const int TRP = a.TR + 1;
  l.TRP = TRP;
  int o = 0;
  l.W1 = o; o += a.EH * a.K;
  l.W2 = o; o += u.W2 ? a.E * a.EH * a.K : 0;
  l.W3 = o; o += u.W3 ? a.C * a.EL2 : 0;
  l.b3 = o; o += a.C;
  l.coef = o; o += 4 * (a.EH + a.E + a.C);       // per BN: A, Cc, mu, r
  l.q = o; o += 2 * (a.EH + a.E + a.C);          // per BN: backward means
  l.pe = o; o += ph == 3 ? a.T * a.C : 0;
  l.accW = o; o += enc_npairs(a, ph);
  o = (o + 1) & ~1;
  l.sacc = o; o += 2 * 2 * imax(imax(a.EH, a.E), a.C);   // doubles
  l.meta = o; o += 2 * a.TR;                     // per row: (b*N+n)*T+t, t
  l.xs = o; o += u.x ? a.P * TRP : 0;
  l.c1 = o; o += u.x ? a.EH * a.L1 * TRP : 0;    // raw conv1
  l.a1 = o; o += u.a1 ? a.EH * a.L1 * TRP : 0;   // relu(BN1(conv1))
  l.c2 = o; o += u.c2 ? a.EL2 * TRP : 0;         // raw conv2 (forward phases: relu(BN2) in place)
  l.z = o; o += u.z ? a.C * TRP : 0;             // raw linear, then dz3
  l.z2 = o; o += u.z2 ? a.C * TRP : 0;           // dropout'(dh)
  l.d2 = o; o += u.d2 ? a.EL2 * TRP : 0;         // relu(BN2(conv2)), then dn2, then dc2
  l.d1 = o; o += u.d1 ? a.EH * a.L1 * TRP : 0;   // dn1, then dc1
  l.total = o;
  return l;
}

inline int tile_rows_for(const EncArgs& a, int ph) {
  EncArgs b = a;
  for (int tr = 256; tr >= kMinTR; tr >>= 1) {
    b.TR = tr;
    if ((size_t)make_lay(b, ph).total * 4 <= kSmemCap) return tr;
  }
  return 0;
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

// sums over one tile of f(ch, i) -> (u, v) for i in [0, cnt), one warp per channel (fixed owner: no atomics)
template <typename Fn>
STG_DEVINL void chan_stats(double* sacc, int nch, int cnt, Fn f) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int ch = warp; ch < nch; ch += kWarps) {
    float s = 0.f, ss = 0.f;
    for (int i = lane; i < cnt; i += 32) {
      float u, v;
      f(ch, i, u, v);
      s += u;
      ss += v;
    }
    s = warp_sum(s);
    ss = warp_sum(ss);
    if (lane == 0) {
      sacc[ch] += (double)s;
      sacc[nch + ch] += (double)ss;
    }
  }
}

// accW[pair] += sum over the tile's rows of mk(pair)(row): the pair is decoded once, outside the row loop.
// Few pairs -> the rows are sliced over threads (shared-memory atomics); many -> each thread owns its pairs.
template <typename Mk>
STG_DEVINL void pair_reduce(float* accW, int npairs, int TR, Mk mk) {
  const int tid = threadIdx.x;
  if (npairs <= kThreads / 2) {
    const int slices = kThreads / npairs;
    if (tid < npairs * slices) {
      const int pair = tid % npairs, sl = tid / npairs;
      const auto f = mk(pair);
      float acc = 0.f;
      for (int r = sl; r < TR; r += slices) acc += f(r);
      atomicAdd(&accW[pair], acc);
    }
  } else {
    for (int pair = tid; pair < npairs; pair += kThreads) {
      const auto f = mk(pair);
      float acc = 0.f;
      for (int r = 0; r < TR; ++r) acc += f(r);
      accW[pair] += acc;
    }
  }
}

// PH: 0..3 forward phases F1..F4, 4..7 backward phases B1..B4.
template <int PH>
__global__ void __launch_bounds__(kThreads) k_encoder(const EncArgs a) {
  extern __shared__ __align__(16) float sm[];
  constexpr bool BWD = PH >= 4;
  const Lay l = make_lay(a, PH);
  const int TRP = l.TRP, TR = a.TR, tid = threadIdx.x;
  const int lg = 31 - __clz(TR), rmask = TR - 1;            // TR is a power of two
  const int EH = a.EH, E = a.E, C = a.C, K = a.K, P = a.P, L1 = a.L1, L2 = a.L2, EL2 = a.EL2, N = a.N, T = a.T;
  const int NL1 = EH * L1;
  float *W1 = sm + l.W1, *W2 = sm + l.W2, *W3 = sm + l.W3, *b3 = sm + l.b3;
  float *cf1 = sm + l.coef, *cf2 = cf1 + 4 * EH, *cf3 = cf2 + 4 * E;
  float *q1 = sm + l.q, *q2 = q1 + 2 * EH, *q3 = q2 + 2 * E;
  float *pe = sm + l.pe, *accW = sm + l.accW;
  double* sacc = reinterpret_cast<double*>(sm + l.sacc);
  int *rowk = reinterpret_cast<int*>(sm + l.meta), *rowt = rowk + TR;
  float *xs = sm + l.xs, *c1 = sm + l.c1, *a1 = sm + l.a1, *c2 = sm + l.c2, *zz = sm + l.z, *z2 = sm + l.z2;
  float *d2 = sm + l.d2, *d1 = sm + l.d1;
  float* a2 = PH >= 5 ? d2 : c2;                            // where relu(BN2(conv2)) is materialised

  const double cnt1 = (double)a.R * L1, cnt2 = (double)a.R * L2, cnt3 = (double)a.R;
  const double* S1 = a.st;                     // forward moments
  const double* S2 = S1 + 2 * EH;
  const double* S3 = S2 + 2 * E;
  double* Bq3 = a.st + 2 * (EH + E + C);       // backward sums
  double* Bq2 = Bq3 + 2 * C;
  double* Bq1 = Bq2 + 2 * E;

  // ---------------- prologue: weights, BN coefficients ----------------
  const Use u = enc_use(a, PH);
  const bool kept = enc_kept(a);
  for (int i = tid; i < EH * K; i += kThreads) W1[i] = a.W1[i];
  if (u.W2) for (int i = tid; i < E * EH * K; i += kThreads) W2[i] = a.W2[i];
  if (u.W3) for (int i = tid; i < C * EL2; i += kThreads) W3[i] = a.W3[i];
  if (PH >= 2) for (int i = tid; i < C; i += kThreads) b3[i] = a.b3[i];
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
  const int npairs = enc_npairs(a, PH);
  for (int i = tid; i < npairs; i += kThreads) accW[i] = 0.f;
  const int nstat = PH == 0 ? EH : PH == 1 ? E : PH == 2 ? C : PH == 4 ? C : PH == 5 ? E : PH == 6 ? EH : 0;
  for (int i = tid; i < 2 * nstat; i += kThreads) sacc[i] = 0.0;
  __syncthreads();

  const float *A1 = cf1, *C1 = cf1 + EH, *mu1 = cf1 + 2 * EH, *r1 = cf1 + 3 * EH;
  const float *A2 = cf2, *C2 = cf2 + E, *mu2 = cf2 + 2 * E, *r2 = cf2 + 3 * E;
  const float *A3 = cf3, *C3 = cf3 + C, *mu3 = cf3 + 2 * C, *r3 = cf3 + 3 * C;

  // linear map: (4 outputs, row) items, the EL2-long contraction split over ks lanes of a warp
  const int ncg = (C + 3) >> 2, nitems = ncg * TR;
  int ks = 1;
  while (ks < 32 && nitems * ks * 2 <= kThreads) ks <<= 1;
  const int ipp = kThreads / ks, slice = tid & (ks - 1), islot = tid / ks;
  const int neg = (E + 3) >> 2, nkg = (EL2 + 3) >> 2, nhg = (EH + 3) >> 2;      // 4-wide output groups

  // tile <-> workspace ([row][feature], feature fastest: coalesced; odd smem pitch: conflict-free)
  auto tile_load = [&](float* dst, const float* src, int F, int r0, int rows) {
    for (int o = tid; o < F * TR; o += kThreads) {
      const int rr = o / F, f = o - rr * F;
      dst[f * TRP + rr] = rr < rows ? src[(size_t)(r0 + rr) * F + f] : 0.f;
    }
  };
  auto tile_store = [&](float* dst, const float* src, int F, int r0, int rows) {
    for (int o = tid; o < F * rows; o += kThreads) {
      const int rr = o / F, f = o - rr * F;
      dst[(size_t)(r0 + rr) * F + f] = src[f * TRP + rr];
    }
  };

  const int ntiles = (a.R + TR - 1) / TR;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int r0 = tile * TR;
    const int rows = min(TR, a.R - r0);
    __syncthreads();                                          // the previous tile is fully consumed
    if (tid < TR) {
      int kx = 0, t = 0;
      if (tid < rows) {
        const int r = r0 + tid, n = r % N, bt = r / N;
        t = bt % T;
        kx = ((bt / T) * N + n) * T + t;
      }
      rowk[tid] = kx;
      rowt[tid] = t;
    }
    __syncthreads();
    if (u.x) {
      // ---- x (zero rows past the end) ----
      for (int o = tid; o < P * TR; o += kThreads) {
        const int rr = o / P, p = o - rr * P;
        xs[p * TRP + rr] = rr < rows ? a.X[(size_t)rowk[rr] * P + p] : 0.f;
      }
      __syncthreads();
      // ---- conv1 (raw, pre-BN) and a1 = relu(BN1(c1)) ----
      for (int o = tid; o < NL1 * TR; o += kThreads) {
        const int rr = o & rmask, f = o >> lg, ch = f / L1, p = f - ch * L1;
        float acc = 0.f;
        for (int j = 0; j < K; ++j) {
          const int q = p + j - a.pad1;
          if (q >= 0 && q < P) acc = fmaf(W1[ch * K + j], xs[q * TRP + rr], acc);
        }
        c1[f * TRP + rr] = acc;
        if (u.a1) a1[f * TRP + rr] = fmaxf(fmaf(A1[ch], acc, C1[ch]), 0.f);
      }
      __syncthreads();
    }
    if (PH == 0) {
      chan_stats(sacc, EH, L1 * TR, [&](int ch, int i, float& su, float& sv) {
        const int rr = i & rmask, p = i >> lg;
        const float x = rr < rows ? c1[(ch * L1 + p) * TRP + rr] : 0.f;
        su = x;
        sv = x * x;
      });
      continue;
    }
    if (u.conv2) {
      // ---- conv2 (raw): items (4 output channels, position, row) share each a1 load ----
      for (int o = tid; o < neg * L2 * TR; o += kThreads) {
        const int rr = o & rmask, f = o >> lg, eg = f / L2, p = f - eg * L2, e0 = eg << 2;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int ch = 0; ch < EH; ++ch)
          for (int j = 0; j < K; ++j) {
            const int q = p + j - 1;
            if (q >= 0 && q < L1) {
              const float av = a1[(ch * L1 + q) * TRP + rr];
              const float* w = W2 + (e0 * EH + ch) * K + j;
#pragma unroll
              for (int v = 0; v < 4; ++v)
                if (e0 + v < E) acc[v] = fmaf(w[v * EH * K], av, acc[v]);
            }
          }
#pragma unroll
        for (int v = 0; v < 4; ++v)
          if (e0 + v < E) {
            const int k = (e0 + v) * L2 + p;
            c2[k * TRP + rr] = acc[v];
            if (PH >= 2) a2[k * TRP + rr] = fmaxf(fmaf(A2[e0 + v], acc[v], C2[e0 + v]), 0.f);   // own element
          }
      }
      __syncthreads();
      if (PH == 1 && kept) tile_store(a.c2raw, c2, EL2, r0, rows);
    } else if (u.c2) {
      // ---- raw conv2 from the workspace; a2 = relu(BN2) beside it (B2) or in its place (F3) ----
      for (int o = tid; o < EL2 * TR; o += kThreads) {
        const int rr = o / EL2, k = o - rr * EL2, e = k / L2;
        const float v = rr < rows ? a.c2raw[(size_t)(r0 + rr) * EL2 + k] : 0.f;
        c2[k * TRP + rr] = v;
        if (PH == 2 || PH == 5) a2[k * TRP + rr] = fmaxf(fmaf(A2[e], v, C2[e]), 0.f);
      }
      __syncthreads();
    }
    if (PH == 1) {
      chan_stats(sacc, E, L2 * TR, [&](int e, int i, float& su, float& sv) {
        const int rr = i & rmask, p = i >> lg;
        const float x = rr < rows ? c2[(e * L2 + p) * TRP + rr] : 0.f;
        su = x;
        sv = x * x;
      });
      continue;
    }
    if (u.lin) {
      // ---- linear (raw z3) on a2 ----
      for (int base = 0; base < nitems; base += ipp) {
        const int it = base + islot;
        const bool valid = it < nitems;
        const int rr = it & rmask, cb = (it >> lg) << 2;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        if (valid)
          for (int k = slice; k < EL2; k += ks) {
            const float av = a2[k * TRP + rr];
#pragma unroll
            for (int v = 0; v < 4; ++v)
              if (cb + v < C) acc[v] = fmaf(W3[(cb + v) * EL2 + k], av, acc[v]);
          }
        for (int off = ks >> 1; off; off >>= 1)
#pragma unroll
          for (int v = 0; v < 4; ++v) acc[v] += __shfl_xor_sync(0xffffffffu, acc[v], off);
        if (valid && slice == 0)
#pragma unroll
          for (int v = 0; v < 4; ++v)
            if (cb + v < C) zz[(cb + v) * TRP + rr] = acc[v] + b3[cb + v];
      }
      __syncthreads();
      if (PH == 2 && kept) tile_store(a.z3raw, zz, C, r0, rows);
    } else if (u.z) {
      tile_load(zz, a.z3raw, C, r0, rows);
      __syncthreads();
    }
    if (PH == 2) {
      chan_stats(sacc, C, TR, [&](int c, int rr, float& su, float& sv) {
        const float x = rr < rows ? zz[c * TRP + rr] : 0.f;
        su = x;
        sv = x * x;
      });
      continue;
    }
    if (PH == 3) {
      // ---- BN3 + positional encoding + dropout -> h[b,t,n,:] ----
      for (int o = tid; o < C * rows; o += kThreads) {
        const int rr = o / C, c = o - rr * C;
        const float hn = fmaf(A3[c], zz[c * TRP + rr], C3[c]) + pe[rowt[rr] * C + c];
        // reference dropout layout [B*N,T,C]
        a.h[(size_t)(r0 + rr) * C + c] = hn * drop_scale(a, drop_row_key(a, (size_t)rowk[rr]), (size_t)rowk[rr], C, c);
      }
      continue;
    }
    if (BWD) {
      if (u.z2) {
        // ---- dhn = dropout'(dh) ----
        for (int o = tid; o < C * TR; o += kThreads) {
          const int rr = o / C, c = o - rr * C;
          z2[c * TRP + rr] = rr < rows ? a.dh[(size_t)(r0 + rr) * C + c] * drop_scale(a, drop_row_key(a, (size_t)rowk[rr]), (size_t)rowk[rr], C, c) : 0.f;
        }
        __syncthreads();
        if (PH == 4) {
          chan_stats(sacc, C, TR, [&](int c, int rr, float& su, float& sv) {
            const float dhn = z2[c * TRP + rr];               // 0 past the end
            su = dhn;
            sv = dhn * (zz[c * TRP + rr] - mu3[c]) * r3[c];
          });
          continue;
        }
        // ---- BN3 backward -> dz3 (in zz) ----
        for (int o = tid; o < C * TR; o += kThreads) {
          const int rr = o & rmask, c = o >> lg;
          const float zh = (zz[c * TRP + rr] - mu3[c]) * r3[c];
          zz[c * TRP + rr] = rr < rows ? A3[c] * (z2[c * TRP + rr] - q3[c] - zh * q3[C + c]) : 0.f;
        }
        __syncthreads();
        if (PH == 5) {
          // dW3[c][k] += sum_r dz3[c][r] * a2[k][r]   (before a2 is overwritten): a thread owns (4 c, k) items
          for (int it = tid; it < ncg * EL2; it += kThreads) {
            const int cg = it / EL2, k = it - cg * EL2, cb = cg << 2;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            for (int rr = 0; rr < rows; ++rr) {
              const float av = a2[k * TRP + rr];
#pragma unroll
              for (int v = 0; v < 4; ++v)
                if (cb + v < C) acc[v] = fmaf(zz[(cb + v) * TRP + rr], av, acc[v]);
            }
#pragma unroll
            for (int v = 0; v < 4; ++v)
              if (cb + v < C) accW[(cb + v) * EL2 + k] += acc[v];
          }
          // db3[c] += sum_r dz3[c][r]
          for (int c = tid; c < C; c += kThreads) {
            float acc = 0.f;
            for (int rr = 0; rr < rows; ++rr) acc += zz[c * TRP + rr];
            accW[C * EL2 + c] += acc;
          }
          __syncthreads();
        }
        // ---- linear backward: da2 -> dn2 (ReLU mask = a2 > 0), in place over a2; 4 features per item ----
        for (int o = tid; o < nkg * TR; o += kThreads) {
          const int rr = o & rmask, k0 = (o >> lg) << 2;
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
          if (rr < rows)
            for (int c = 0; c < C; ++c) {
              const float dz = zz[c * TRP + rr];
              const float* w = W3 + c * EL2 + k0;
#pragma unroll
              for (int v = 0; v < 4; ++v)
                if (k0 + v < EL2) acc[v] = fmaf(dz, w[v], acc[v]);
            }
#pragma unroll
          for (int v = 0; v < 4; ++v)
            if (k0 + v < EL2) {
              float* dst = d2 + (k0 + v) * TRP + rr;
              *dst = (rr < rows && *dst > 0.f) ? acc[v] : 0.f;
            }
        }
        __syncthreads();
        if (PH == 5) {
          if (kept) tile_store(a.dn2, d2, EL2, r0, rows);
          chan_stats(sacc, E, L2 * TR, [&](int e, int i, float& su, float& sv) {
            const int rr = i & rmask, k = e * L2 + (i >> lg);
            const float dn = d2[k * TRP + rr];                // 0 past the end
            su = dn;
            sv = dn * (c2[k * TRP + rr] - mu2[e]) * r2[e];
          });
          continue;
        }
      } else if (PH == 6) {
        tile_load(d2, a.dn2, EL2, r0, rows);
        __syncthreads();
      }
      if (u.d2) {
        // ---- BN2 backward (in place) -> dc2 ----
        for (int o = tid; o < EL2 * TR; o += kThreads) {
          const int rr = o & rmask, k = o >> lg, e = k / L2;
          const float ch2 = (c2[k * TRP + rr] - mu2[e]) * r2[e];
          d2[k * TRP + rr] = rr < rows ? A2[e] * (d2[k * TRP + rr] - q2[e] - ch2 * q2[E + e]) : 0.f;
        }
        __syncthreads();
        // ---- conv2 backward wrt its input: da1 -> dn1 (ReLU mask = a1 > 0); 4 channels per item ----
        for (int o = tid; o < nhg * L1 * TR; o += kThreads) {
          const int rr = o & rmask, f = o >> lg, hg = f / L1, q = f - hg * L1, ch0 = hg << 2;
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
          if (rr < rows)
            for (int e = 0; e < E; ++e)
              for (int j = 0; j < K; ++j) {
                const int p = q - j + 1;
                if (p >= 0 && p < L2) {
                  const float dv = d2[(e * L2 + p) * TRP + rr];
                  const float* w = W2 + (e * EH + ch0) * K + j;
#pragma unroll
                  for (int v = 0; v < 4; ++v)
                    if (ch0 + v < EH) acc[v] = fmaf(dv, w[v * K], acc[v]);
                }
              }
#pragma unroll
          for (int v = 0; v < 4; ++v)
            if (ch0 + v < EH) {
              const int g = (ch0 + v) * L1 + q;
              d1[g * TRP + rr] = (rr < rows && a1[g * TRP + rr] > 0.f) ? acc[v] : 0.f;
            }
        }
        __syncthreads();
      } else {
        tile_load(d1, a.dn1, NL1, r0, rows);                  // B4 with kept activations
        __syncthreads();
      }
      if (PH == 6) {
        if (kept) tile_store(a.dn1, d1, NL1, r0, rows);
        chan_stats(sacc, EH, L1 * TR, [&](int ch, int i, float& su, float& sv) {
          const int rr = i & rmask, f = ch * L1 + (i >> lg);
          const float dn = d1[f * TRP + rr];                  // 0 past the end
          su = dn;
          sv = dn * (c1[f * TRP + rr] - mu1[ch]) * r1[ch];
        });
        // dW2[e][ch][j] += sum_r sum_p dc2[e][p][r] * a1[ch][p+j-1][r]
        pair_reduce(accW, E * EH * K, rows, [&](int pair) {
          const int j = pair % K, ech = pair / K, ch = ech % EH, e = ech / EH;
          const int p0 = imax(0, 1 - j), p1 = min(L2, L1 + 1 - j);          // 0 <= p + j - 1 < L1
          const float* dv = d2 + (e * L2) * TRP;
          const float* av = a1 + (ch * L1 + j - 1) * TRP;
          return [=](int rr) {
            float acc = 0.f;
            for (int p = p0; p < p1; ++p) acc = fmaf(dv[p * TRP + rr], av[p * TRP + rr], acc);
            return acc;
          };
        });
        continue;
      }
      // ---- PH == 7: BN1 backward -> dc1; dW1[ch][j] += sum_r sum_p dc1[ch][p][r] * xpad[p+j-pad1][r]
      for (int o = tid; o < NL1 * TR; o += kThreads) {
        const int rr = o & rmask, f = o >> lg, ch = f / L1;
        const float ch1 = (c1[f * TRP + rr] - mu1[ch]) * r1[ch];
        d1[f * TRP + rr] = rr < rows ? A1[ch] * (d1[f * TRP + rr] - q1[ch] - ch1 * q1[EH + ch]) : 0.f;
      }
      __syncthreads();
      pair_reduce(accW, EH * K, rows, [&](int pair) {
        const int j = pair % K, ch = pair / K;
        const int p0 = imax(0, a.pad1 - j), p1 = min(L1, P + a.pad1 - j);   // 0 <= p + j - pad1 < P
        const float* dv = d1 + (ch * L1) * TRP;
        const float* xv = xs + (j - a.pad1) * TRP;
        return [=](int rr) {
          float acc = 0.f;
          for (int p = p0; p < p1; ++p) acc = fmaf(dv[p * TRP + rr], xv[p * TRP + rr], acc);
          return acc;
        };
      });
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

void launch_phase(int ph, const EncArgs& a0, cudaStream_t s) {
  EncArgs a = a0;
  a.TR = tile_rows_for(a0, ph);                   // plan_encoder checked that every phase fits
  const size_t smem = (size_t)make_lay(a, ph).total * 4;
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
  bool fits = true;
  for (int ph = 0; ph < 8 && fits; ++ph) fits = tile_rows_for(a, ph) > 0;
  if (fits) {
    a.TR = tile_rows_for(a, 7);
    *smem_fwd = *smem_bwd = 0;                    // chosen per phase at launch
    return 0;
  }
  snprintf(err, errlen, "encoder tile does not fit shared memory (P=%d K=%d EH=%d E=%d C=%d)", a.P, a.K, a.EH, a.E, a.C);
  return -2;
}

int launch_encoder_forward(const EncArgs& a, size_t smem, cudaStream_t s) {
  if (a.c2raw && encoder_fast_available(a)) return launch_encoder_fast(a, false, s);
  set_attrs();
  if (a.training)
    for (int ph = 0; ph < 3; ++ph) launch_phase(ph, a, s);
  launch_phase(3, a, s);
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

int launch_encoder_backward(const EncArgs& a, size_t smem, cudaStream_t s) {
  if (a.c2raw && encoder_fast_available(a)) return launch_encoder_fast(a, true, s);
  set_attrs();
  for (int ph = 4; ph < 8; ++ph) launch_phase(ph, a, s);
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

}  // namespace stg
