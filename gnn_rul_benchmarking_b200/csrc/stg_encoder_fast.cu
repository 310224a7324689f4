// Patch encoder, compile-time-dimension fast path (sm_100a).  Same math and phase structure as the
// generic kernels in stg_encoder.cu (which remain the fallback for every other hyper-parameter set),
// but: all per-row activations live in registers (loops fully unrolled), and the raw conv2 / linear
// outputs and the ReLU-masked gradients are kept in the workspace between phases in a
// [tile][feature][256 rows] layout (perfectly coalesced for one-thread-per-row), so no phase
// recomputes more than conv1:
//   F1 conv1 moments | F2 conv1,conv2 -> c2raw, moments | F3 c2raw -> linear -> z3raw, moments | F4 z3raw -> h
//   B1 dh,z3raw -> BN3 sums | B2 -> dW3,db3, dn2, BN2 sums | B3 -> dW2, dn1, BN1 sums | B4 -> dW1
#include <math.h>
#include <stdlib.h>

#include "stg_model.cuh"

namespace stg {
namespace {

constexpr int kT = 128;          // threads per CTA == rows per CTA tile
constexpr int kLP = 256;         // rows per workspace tile ([tile][feature][256]): a CTA tile is one half of it
constexpr int kTP = 132;         // smem row pitch of the staged columns (float4-aligned)

template <int P_, int K_, int EH_, int E_, int C_>
struct EncDims {
  static constexpr int P = P_, K = K_, EH = EH_, E = E_, C = C_;
  static constexpr int pad1 = K / 2, L1 = P + 2 * pad1 - K + 1, L2 = L1 + 2 - K + 1;
  static constexpr int EL2 = E * L2, NL1 = EH * L1;
};


// swarp[warp][idx] accumulates this warp's tiles (one owner lane per slot, no atomics); the CTA reduces its warps in
// double at the very end.  v[0..NS) are this thread's contributions to slots 0..NS-1.
template <int NS, int BASE = 0>
STG_DEVINL void stat_flush(float (*swarp)[2 * 48], const float (&v)[NS]) {
  constexpr int n = NS - BASE < 32 ? NS - BASE : 32;          // values of this chunk
  constexpr int NP = n > 16 ? 32 : n > 8 ? 16 : 8;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float w[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) w[i] = i < n ? v[i < n ? BASE + i : 0] : 0.f;
  const float r = warp_multi_sum<NP>(w);
  constexpr int per = 32 / NP;                                 // lanes holding the same total
  if ((lane & (per - 1)) == 0 && lane / per < n) swarp[warp][BASE + lane / per] += r;
  if constexpr (BASE + 32 < NS) stat_flush<NS, BASE + 32>(swarp, v);
}

template <class D>
STG_DEVINL void conv1_row(const float* W1, const float (&x)[D::P], float (&c1)[D::NL1]) {
#pragma unroll
  for (int ch = 0; ch < D::EH; ++ch)
#pragma unroll
    for (int p = 0; p < D::L1; ++p) {
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < D::K; ++j) {
        const int q = p + j - D::pad1;
        if (q >= 0 && q < D::P) acc = fmaf(W1[ch * D::K + j], x[q], acc);
      }
      c1[ch * D::L1 + p] = acc;
    }
}

template <class D>
STG_DEVINL void conv2_row(const float* W2, const float (&a1)[D::NL1], float (&c2)[D::EL2]) {
#pragma unroll
  for (int e = 0; e < D::E; ++e) {
    float acc[D::L2];
#pragma unroll
    for (int p = 0; p < D::L2; ++p) acc[p] = 0.f;
#pragma unroll
    for (int ch = 0; ch < D::EH; ++ch)
#pragma unroll
      for (int j = 0; j < D::K; ++j) {
        const float w = W2[(e * D::EH + ch) * D::K + j];
#pragma unroll
        for (int p = 0; p < D::L2; ++p) {
          const int q = p + j - 1;
          if (q >= 0 && q < D::L1) acc[p] = fmaf(w, a1[ch * D::L1 + q], acc[p]);
        }
      }
#pragma unroll
    for (int p = 0; p < D::L2; ++p) c2[e * D::L2 + p] = acc[p];
  }
}

// pairs over threads, float4 over the 256 staged rows; f4(pair, r4) returns the partial product sum
template <int NP, typename Fn>
STG_DEVINL void pair_reduce_f(float* accW, Fn f) {
  const int tid = threadIdx.x;
  if (NP <= kT / 2) {
    constexpr int slices = NP <= kT / 2 ? kT / (NP > 0 ? NP : 1) : 1;
    if (tid < NP * slices) {
      const int pair = tid % NP, sl = tid / NP;
      float acc = 0.f;
      for (int r4 = sl; r4 < kT / 4; r4 += slices) acc += f(pair, r4 * 4);
      atomicAdd(&accW[pair], acc);
    }
  } else {
    for (int pair = tid; pair < NP; pair += kT) {
      float acc = 0.f;
      for (int r4 = 0; r4 < kT / 4; ++r4) acc += f(pair, r4 * 4);
      accW[pair] += acc;
    }
  }
}

// one row of C contiguous floats (row pitch C): widest aligned vector access (C is even; rows are
// 16-byte aligned when C % 4 == 0, 8-byte aligned otherwise)
template <int C>
STG_DEVINL void load_row(const float* __restrict__ p, float (&v)[C]) {
  if constexpr (C % 4 == 0) {
#pragma unroll
    for (int c = 0; c < C; c += 4) {
      const float4 q = *reinterpret_cast<const float4*>(p + c);
      v[c] = q.x; v[c + 1] = q.y; v[c + 2] = q.z; v[c + 3] = q.w;
    }
  } else {
#pragma unroll
    for (int c = 0; c < C; c += 2) {
      const float2 q = *reinterpret_cast<const float2*>(p + c);
      v[c] = q.x; v[c + 1] = q.y;
    }
  }
}
template <int C>
STG_DEVINL void store_row(float* __restrict__ p, const float (&v)[C]) {
  if constexpr (C % 4 == 0) {
#pragma unroll
    for (int c = 0; c < C; c += 4) *reinterpret_cast<float4*>(p + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
  } else {
#pragma unroll
    for (int c = 0; c < C; c += 2) *reinterpret_cast<float2*>(p + c) = make_float2(v[c], v[c + 1]);
  }
}

#ifdef STG_ENC_TIMING
// debug build only (scripts/enc_cta_times.py): per phase and CTA, globaltimer at entry and clock64 deltas of the
// prologue / tile loop / epilogue
__device__ unsigned long long g_enc_time[9][1024][8];
#define ENC_STAMP(slot)                                                                        \
  if (threadIdx.x == 0 && blockIdx.x < 1024) {                                                 \
    if (slot == 0) {                                                                           \
      unsigned long long tg;                                                                   \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tg));                                   \
      g_enc_time[PH][blockIdx.x][0] = tg;                                                      \
      enc_t0 = clock64();                                                                      \
    } else {                                                                                   \
      g_enc_time[PH][blockIdx.x][slot] = (unsigned long long)(clock64() - enc_t0);             \
    }                                                                                          \
  }
#else
#define ENC_STAMP(slot)
#endif

STG_DEVINL float dot4(const float4 a, const float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
STG_DEVINL float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

template <class D, int PH>
__global__ void __launch_bounds__(kT, PH == 8 ? 1 : (D::C > 16 || D::E > 6) ? 3 : 5) k_enc_fast(const EncArgs a) {
  constexpr int P = D::P, K = D::K, EH = D::EH, E = D::E, C = D::C, L1 = D::L1, L2 = D::L2, EL2 = D::EL2, NL1 = D::NL1;
  constexpr int MAXCH = EH > E ? (EH > C ? EH : C) : (E > C ? E : C);
  constexpr int NPAIR = PH == 5 ? C * EL2 + C : PH == 6 ? E * EH * K : PH == 7 ? EH * K : 1;
  __shared__ __align__(16) float W1[EH * K];
  __shared__ __align__(16) float W2[E * EH * K];
  __shared__ __align__(16) float W3[C * EL2];
  __shared__ float b3[C];
  __shared__ float cf1[4 * EH], cf2[4 * E], cf3[4 * C];
  __shared__ float q1[2 * EH], q2[2 * E], q3[2 * C];
  __shared__ float sacc[kT / 32][2 * 48];
  __shared__ float accW[NPAIR];
  extern __shared__ __align__(16) float stage[];      // [features][kTP] columns for the pair reductions / pe
  const int tid = threadIdx.x;
  const int N = a.N, T = a.T;
#ifdef STG_ENC_TIMING
  long long enc_t0 = 0;
#endif
  ENC_STAMP(0)

  const double cnt1 = (double)a.R * L1, cnt2 = (double)a.R * L2, cnt3 = (double)a.R;
  const double* S1 = a.st;
  const double* S2 = S1 + 2 * EH;
  const double* S3 = S2 + 2 * E;
  double* Bq3 = a.st + 2 * (EH + E + C);
  double* Bq2 = Bq3 + 2 * C;
  double* Bq1 = Bq2 + 2 * E;

  for (int i = tid; i < NPAIR; i += kT) accW[i] = 0.f;
  static_assert(MAXCH <= 48, "stat slots");
  for (int i = tid; i < (kT / 32) * 2 * 48; i += kT) (&sacc[0][0])[i] = 0.f;
  pdl_sync();
  for (int i = tid; i < EH * K; i += kT) W1[i] = a.W1[i];
  for (int i = tid; i < E * EH * K; i += kT) W2[i] = a.W2[i];
  for (int i = tid; i < C * EL2; i += kT) W3[i] = a.W3[i];
  for (int i = tid; i < C; i += kT) b3[i] = a.b3[i];
  const bool first = blockIdx.x == 0, tr = a.training != 0;
  // the (up to) three BatchNorm coefficient sets in ONE pass: thread i owns one (set, channel) -- three calls in a row were
  // three serial rounds of dependent global loads + double arithmetic in every CTA's prologue
  {
    constexpr int n1 = PH >= 1 ? EH : 0, n2 = PH >= 2 ? E : 0, n3 = PH >= 3 ? C : 0;
    for (int i = tid; i < n1 + n2 + n3; i += kT) {
      const int set = i < n1 ? 1 : (i < n1 + n2 ? 2 : 3);
      const int c = set == 1 ? i : (set == 2 ? i - n1 : i - n1 - n2);
      const int n = set == 1 ? EH : (set == 2 ? E : C);
      float* dst = set == 1 ? cf1 : (set == 2 ? cf2 : cf3);
      const double* st = set == 1 ? S1 : (set == 2 ? S2 : S3);
      const double count = set == 1 ? cnt1 : (set == 2 ? cnt2 : cnt3);
      const float* gg = set == 1 ? a.g1 : (set == 2 ? a.g2 : a.g3);
      const float* bb = set == 1 ? a.be1 : (set == 2 ? a.be2 : a.be3);
      float* rm = set == 1 ? a.rm1 : (set == 2 ? a.rm2 : a.rm3);
      float* rv = set == 1 ? a.rv1 : (set == 2 ? a.rv2 : a.rv3);
      const bool update = tr && first && PH == set;
      double m, var;
      if (tr) {
        const double icount = inv_d(count);
        m = st[c] * icount;
        var = st[n + c] * icount - m * m;
        if (var < 0.0) var = 0.0;
        if (update) {
          const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
          rm[c] = (1.f - a.momentum) * rm[c] + a.momentum * (float)m;
          rv[c] = (1.f - a.momentum) * rv[c] + a.momentum * (float)unb;
        }
      } else {
        m = rm[c];
        var = rv[c];
      }
      const float r = (float)rsqrt_d(var + (double)a.eps);
      const float A = gg[c] * r;
      dst[c] = A;
      dst[n + c] = bb[c] - A * (float)m;
      dst[2 * n + c] = (float)m;
      dst[3 * n + c] = r;
    }
  }
  ENC_STAMP(7)
  if (PH == 3 || PH == 8) {                     // positional-encoding rows: one 16-byte load round when the table is aligned
    if ((C & 3) == 0 && ((uintptr_t)a.pe & 15) == 0) {
      for (int i = tid; i < T * C / 4; i += kT) reinterpret_cast<float4*>(stage)[i] = reinterpret_cast<const float4*>(a.pe)[i];
    } else {
      for (int i = tid; i < T * C; i += kT) stage[i] = a.pe[i];
    }
  }
  const double ic1 = inv_d(cnt1), ic2 = inv_d(cnt2), ic3 = inv_d(cnt3);
  if (PH >= 5 && PH <= 7) for (int c = tid; c < C; c += kT) {
    q3[c] = (float)(Bq3[c] * ic3);
    q3[C + c] = (float)(Bq3[C + c] * ic3);
    if (PH == 5 && first) { a.dbe3[c] += (float)Bq3[c]; a.dg3[c] += (float)Bq3[C + c]; }
  }
  if (PH >= 6 && PH <= 7) for (int c = tid; c < E; c += kT) {
    q2[c] = (float)(Bq2[c] * ic2);
    q2[E + c] = (float)(Bq2[E + c] * ic2);
    if (PH == 6 && first) { a.dbe2[c] += (float)Bq2[c]; a.dg2[c] += (float)Bq2[E + c]; }
  }
  if (PH == 7) for (int c = tid; c < EH; c += kT) {
    q1[c] = (float)(Bq1[c] * ic1);
    q1[EH + c] = (float)(Bq1[EH + c] * ic1);
    if (first) { a.dbe1[c] += (float)Bq1[c]; a.dg1[c] += (float)Bq1[EH + c]; }
  }
  if (PH == 4 && a.fin.nblk) {
    // stage: per block  mu0[C] r0[C] m1[C] m2[C] cnt[T]
    for (int z = 0; z < a.fin.nblk; ++z) {
      float* tb = stage + z * (4 * C + T);
      const int Hz = a.fin.H[z];
      const double Rz = (double)a.B * a.fin.L[z] * a.fin.w[z] * N, iRz = inv_d(Rz);
      for (int c = tid; c < C; c += kT) {
        const float r0 = a.fin.tab[z][a.fin.CP + c];
        const double sb = a.fin.stats[z][4 * Hz + c], sg = a.fin.stats[z][4 * Hz + C + c];
        const float g0 = a.fin.g0[z][c];
        tb[c] = a.fin.tab[z][c];
        tb[C + c] = r0;
        tb[2 * C + c] = r0 * (float)(g0 * sb * iRz);
        tb[3 * C + c] = r0 * (float)(g0 * sg * iRz);
        if (first) { a.fin.db0[z][c] += (float)sb; a.fin.dg0[z][c] += (float)sg; }
      }
      for (int t2 = tid; t2 < T; t2 += kT) {
        int cn = 0;
        for (int j = 0; j < a.fin.w[z]; ++j) {
          const int d = t2 - j;
          if (d >= 0 && d % a.fin.stride[z] == 0 && d / a.fin.stride[z] < a.fin.L[z]) ++cn;
        }
        tb[4 * C + t2] = (float)cn;
      }
      if (first)
        for (int h = tid; h < Hz; h += kT) {
          a.fin.db1[z][h] += (float)a.fin.stats[z][2 * Hz + h];
          a.fin.dg1[z][h] += (float)a.fin.stats[z][3 * Hz + h];
        }
    }
  }
  __syncthreads();

  ENC_STAMP(1)
  const float *A1 = cf1, *C1 = cf1 + EH, *mu1 = cf1 + 2 * EH, *r1 = cf1 + 3 * EH;
  const float *A2 = cf2, *C2 = cf2 + E, *mu2 = cf2 + 2 * E, *r2 = cf2 + 3 * E;
  const float *A3 = cf3, *C3 = cf3 + C, *mu3 = cf3 + 2 * C, *r3 = cf3 + 3 * C;

  const int ntiles = (a.R + kT - 1) / kT;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    // compiler barrier: without it the weight loads (loop-invariant shared memory) are hoisted out of the
    // tile loop into several hundred registers / local memory
    asm volatile("" ::: "memory");
    const int r = tile * kT + tid;
    const bool act = r < a.R;
    int n = 0, t = 0, b = 0;
    if (act) {
      n = r % N;
      const int bt = r / N;
      t = bt % T;
      b = bt / T;
    }
    const size_t wt = (size_t)(tile >> 1);                    // workspace tile, this CTA tile's half of its rows
    const int wo = (tile & 1) * kT + tid;
    float* __restrict__ c2t = a.c2raw + wt * EL2 * kLP + wo;      // + k*kLP
    float* __restrict__ z3t = a.z3raw + wt * C * kLP + wo;        // + c*kLP
    float* __restrict__ dn2t = a.dn2 + wt * EL2 * kLP + wo;
    float* __restrict__ dn1t = a.dn1 + wt * NL1 * kLP + wo;

    // ---- x and conv1 where the phase needs them ----
    float x[P], c1[NL1];
    constexpr bool NEED_C1 = PH == 0 || PH == 1 || PH == 6 || PH == 7 || PH == 8;
    if (NEED_C1) {
#pragma unroll
      for (int i = 0; i < P; ++i) x[i] = 0.f;
      if (act) {
        const float* xp = a.X + ((size_t)(b * N + n) * T + t) * P;
#pragma unroll
        for (int i = 0; i < P; ++i) x[i] = xp[i];
      }
      conv1_row<D>(W1, x, c1);
    }

    if constexpr (PH == 0) {
      float st[2 * EH];
#pragma unroll
      for (int ch = 0; ch < EH; ++ch) {
        float s = 0.f, ss = 0.f;
#pragma unroll
        for (int p = 0; p < L1; ++p) {
          const float v = act ? c1[ch * L1 + p] : 0.f;
          s += v;
          ss = fmaf(v, v, ss);
        }
        st[ch] = s;
        st[EH + ch] = ss;
      }
      stat_flush<2 * EH>(sacc, st);
    } else if constexpr (PH == 1) {
      float a1[NL1], c2[EL2];
#pragma unroll
      for (int i = 0; i < NL1; ++i) a1[i] = fmaxf(fmaf(A1[i / L1], c1[i], C1[i / L1]), 0.f);
      conv2_row<D>(W2, a1, c2);
#pragma unroll
      for (int k = 0; k < EL2; ++k) c2t[k * kLP] = c2[k];
      float st[2 * E];
#pragma unroll
      for (int e = 0; e < E; ++e) {
        float s = 0.f, ss = 0.f;
#pragma unroll
        for (int p = 0; p < L2; ++p) {
          const float v = act ? c2[e * L2 + p] : 0.f;
          s += v;
          ss = fmaf(v, v, ss);
        }
        st[e] = s;
        st[E + e] = ss;
      }
      stat_flush<2 * E>(sacc, st);
    } else if constexpr (PH == 2 || PH == 3 || PH == 8) {
      float z[C];
      if constexpr (PH == 2 || PH == 8) {
        float a2[EL2];
        if constexpr (PH == 2) {
          // all loads of the row first (one round trip), then the arithmetic
#pragma unroll
          for (int k = 0; k < EL2; ++k) a2[k] = act ? c2t[k * kLP] : 0.f;
#pragma unroll
          for (int k = 0; k < EL2; ++k) a2[k] = fmaxf(fmaf(A2[k / L2], a2[k], C2[k / L2]), 0.f);
        } else {                 // eval forward: whole chain in one pass (no stored intermediates)
          float a1[NL1], c2[EL2];
#pragma unroll
          for (int i = 0; i < NL1; ++i) a1[i] = fmaxf(fmaf(A1[i / L1], c1[i], C1[i / L1]), 0.f);
          conv2_row<D>(W2, a1, c2);
#pragma unroll
          for (int k = 0; k < EL2; ++k) a2[k] = fmaxf(fmaf(A2[k / L2], c2[k], C2[k / L2]), 0.f);
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
          float acc = b3[c];
#pragma unroll
          for (int k = 0; k < EL2; ++k) acc = fmaf(W3[c * EL2 + k], a2[k], acc);
          z[c] = acc;
        }
      } else {
#pragma unroll
        for (int c = 0; c < C; ++c) z[c] = act ? z3t[c * kLP] : 0.f;
      }
      if constexpr (PH == 2) {
#pragma unroll
        for (int c = 0; c < C; ++c) z3t[c * kLP] = z[c];
        float st[2 * C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float v = act ? z[c] : 0.f;
          st[c] = v;
          st[C + c] = v * v;
        }
        stat_flush<2 * C>(sacc, st);
      } else if (act) {
        float* hr = a.h + (size_t)r * C;
        const size_t krow = (size_t)(b * N + n) * T + t;
      const unsigned long long dkey = drop_row_key(a, krow);
        const float* pe = stage + t * C;
        float hv[C];
#pragma unroll
        for (int c = 0; c < C; ++c) hv[c] = (fmaf(A3[c], z[c], C3[c]) + pe[c]) * drop_scale(a, dkey, krow, C, c);
        store_row<C>(hr, hv);
      }
    } else if constexpr (PH == 4) {
      const size_t krow = (size_t)(b * N + n) * T + t;
      const unsigned long long dkey = drop_row_key(a, krow);
      float dhv[C], z3[C];
#pragma unroll
      for (int c = 0; c < C; ++c) z3[c] = act ? z3t[c * kLP] : 0.f;
      if (act) {
        if (a.fin.nblk) {
          float xv[C];
          load_row<C>(a.h + (size_t)r * C, xv);
#pragma unroll
          for (int c = 0; c < C; ++c) dhv[c] = 0.f;
          if (C <= 16 && a.fin.unfolded && a.fin.nblk == 2 && a.fin.w[0] == 2 && a.fin.w[1] == 2) {
            // the usual shape (two blocks, windows of two time steps): the (up to) four partial rows covering
            // time step t are fetched together, rows (l, j, n) with l*stride + j == t
            float pv[4][C];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int z = q >> 1, j = q & 1;
              const int Lz = a.fin.L[z], sz = a.fin.stride[z];
              const int d = t - j, l = d / sz;
              const bool ok = d >= 0 && l * sz == d && l < Lz;
              if (ok) load_row<C>(a.fin.dxp[z] + ((((size_t)b * Lz + l) * 2 + j) * N + n) * C, pv[q]);
              else {
#pragma unroll
                for (int c = 0; c < C; ++c) pv[q][c] = 0.f;
              }
            }
#pragma unroll
            for (int z = 0; z < 2; ++z) {
              const float* tb = stage + z * (4 * C + T);
              const float cn = tb[4 * C + t];
#pragma unroll
              for (int c = 0; c < C; ++c) {
                const float xh = (xv[c] - tb[c]) * tb[C + c];
                dhv[c] += (pv[2 * z][c] + pv[2 * z + 1][c]) - cn * (tb[2 * C + c] + xh * tb[3 * C + c]);
              }
            }
          } else {
            for (int z = 0; z < a.fin.nblk; ++z) {
              const float* tb = stage + z * (4 * C + T);
              const float cn = tb[4 * C + t];
              float dv[C];
              if (a.fin.unfolded) {
#pragma unroll
                for (int c = 0; c < C; ++c) dv[c] = 0.f;
                const int Lz = a.fin.L[z], sz = a.fin.stride[z], wz = a.fin.w[z];
                for (int j = 0; j < wz; ++j) {
                  const int d = t - j;
                  if (d < 0 || d % sz || d / sz >= Lz) continue;
                  float pv[C];
                  load_row<C>(a.fin.dxp[z] + ((((size_t)b * Lz + d / sz) * wz + j) * N + n) * C, pv);
#pragma unroll
                  for (int c = 0; c < C; ++c) dv[c] += pv[c];
                }
              } else {
                load_row<C>(a.fin.dxp[z] + (size_t)r * C, dv);
              }
#pragma unroll
              for (int c = 0; c < C; ++c) {
                const float xh = (xv[c] - tb[c]) * tb[C + c];
                dhv[c] += dv[c] - cn * (tb[2 * C + c] + xh * tb[3 * C + c]);
              }
            }
          }
          store_row<C>(a.fin.dh_out + (size_t)r * C, dhv);
        } else {
          load_row<C>(a.dh + (size_t)r * C, dhv);
        }
      }
      ENC_STAMP(4)
      float st[2 * C];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float dhn = 0.f, zh = 0.f;
        if (act) {
          dhn = dhv[c] * drop_scale(a, dkey, krow, C, c);
          zh = (z3[c] - mu3[c]) * r3[c];
        }
        st[c] = dhn;
        st[C + c] = dhn * zh;
      }
      ENC_STAMP(5)
      stat_flush<2 * C>(sacc, st);
      ENC_STAMP(6)
    } else if constexpr (PH == 5) {
      float* sdz = stage;                 // [C][kTP]
      float* sa2 = stage + C * kTP;       // [EL2][kTP]
      float dz[C];
      const size_t krow = (size_t)(b * N + n) * T + t;
      const unsigned long long dkey = drop_row_key(a, krow);
      // the row's inputs first (one round trip to memory), arithmetic afterwards
      float c2v[EL2];
      {
        float dhv[C], z3[C];
#pragma unroll
        for (int c = 0; c < C; ++c) { dhv[c] = 0.f; z3[c] = act ? z3t[c * kLP] : 0.f; }
        if (act) load_row<C>(a.dh + (size_t)r * C, dhv);
#pragma unroll
        for (int k = 0; k < EL2; ++k) c2v[k] = act ? c2t[k * kLP] : 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          float v = 0.f;
          if (act) {
            const float dhn = dhv[c] * drop_scale(a, dkey, krow, C, c);
            const float zh = (z3[c] - mu3[c]) * r3[c];
            v = A3[c] * (dhn - q3[c] - zh * q3[C + c]);
          }
          dz[c] = v;
          sdz[c * kTP + tid] = v;
        }
      }
      ENC_STAMP(4)
      float st[2 * E];
#pragma unroll
      for (int e = 0; e < E; ++e) {
        float s = 0.f, sh = 0.f;
#pragma unroll
        for (int p = 0; p < L2; ++p) {
          const int k = e * L2 + p;
          const float a2v = act ? fmaxf(fmaf(A2[e], c2v[k], C2[e]), 0.f) : 0.f;
          sa2[k * kTP + tid] = a2v;
          float v = 0.f;
#pragma unroll
          for (int c = 0; c < C; ++c) v = fmaf(dz[c], W3[c * EL2 + k], v);
          v = a2v > 0.f ? v : 0.f;
          dn2t[k * kLP] = v;
          s += v;
          sh = fmaf(v, (c2v[k] - mu2[e]) * r2[e], sh);
        }
        st[e] = s;
        st[E + e] = sh;
      }
      ENC_STAMP(5)
      stat_flush<2 * E>(sacc, st);
      __syncthreads();
      ENC_STAMP(6)
      pair_reduce_f<C * EL2 + C>(accW, [&](int pair, int r4) {
        if (pair >= C * EL2) {
          const float4 d = ld4(sdz + (pair - C * EL2) * kTP + r4);
          return d.x + d.y + d.z + d.w;
        }
        const int c = pair / EL2, k = pair - c * EL2;
        return dot4(ld4(sdz + c * kTP + r4), ld4(sa2 + k * kTP + r4));
      });
      __syncthreads();
    } else if constexpr (PH == 6) {
      float* sdc = stage;                 // [EL2][kTP]
      float* sa1 = stage + EL2 * kTP;     // [NL1][kTP]
      float dc2[EL2];
      {
        float c2v[EL2], dnv[EL2];
#pragma unroll
        for (int k = 0; k < EL2; ++k) { c2v[k] = act ? c2t[k * kLP] : 0.f; dnv[k] = act ? dn2t[k * kLP] : 0.f; }
#pragma unroll
        for (int k = 0; k < EL2; ++k) {
          const int e = k / L2;
          float v = 0.f;
          if (act) {
            const float ch2 = (c2v[k] - mu2[e]) * r2[e];
            v = A2[e] * (dnv[k] - q2[e] - ch2 * q2[E + e]);
          }
          dc2[k] = v;
          sdc[k * kTP + tid] = v;
        }
      }
      float st[2 * EH];
#pragma unroll
      for (int ch = 0; ch < EH; ++ch) {
        float s = 0.f, sh = 0.f;
#pragma unroll
        for (int q = 0; q < L1; ++q) {
          const float a1v = act ? fmaxf(fmaf(A1[ch], c1[ch * L1 + q], C1[ch]), 0.f) : 0.f;
          sa1[(ch * L1 + q) * kTP + tid] = a1v;
          float v = 0.f;
#pragma unroll
          for (int e = 0; e < E; ++e)
#pragma unroll
            for (int j = 0; j < K; ++j) {
              const int p = q - j + 1;
              if (p >= 0 && p < L2) v = fmaf(dc2[e * L2 + p], W2[(e * EH + ch) * K + j], v);
            }
          v = a1v > 0.f ? v : 0.f;
          dn1t[(ch * L1 + q) * kLP] = v;
          s += v;
          sh = fmaf(v, (c1[ch * L1 + q] - mu1[ch]) * r1[ch], sh);
        }
        st[ch] = s;
        st[EH + ch] = sh;
      }
      stat_flush<2 * EH>(sacc, st);
      __syncthreads();
      pair_reduce_f<E * EH * K>(accW, [&](int pair, int r4) {
        const int j = pair % K, ech = pair / K, ch = ech % EH, e = ech / EH;
        float acc = 0.f;
#pragma unroll
        for (int p = 0; p < L2; ++p) {
          const int q = p + j - 1;
          if (q >= 0 && q < L1) acc += dot4(ld4(sdc + (e * L2 + p) * kTP + r4), ld4(sa1 + (ch * L1 + q) * kTP + r4));
        }
        return acc;
      });
      __syncthreads();
    } else if constexpr (PH == 7) {
      float* sdc = stage;                 // [NL1][kTP]
      float* sx = stage + NL1 * kTP;      // [P][kTP]
      {
        float dnv[NL1];
#pragma unroll
        for (int i = 0; i < NL1; ++i) dnv[i] = act ? dn1t[i * kLP] : 0.f;
#pragma unroll
        for (int i = 0; i < NL1; ++i) {
          const int ch = i / L1;
          float v = 0.f;
          if (act) {
            const float ch1 = (c1[i] - mu1[ch]) * r1[ch];
            v = A1[ch] * (dnv[i] - q1[ch] - ch1 * q1[EH + ch]);
          }
          sdc[i * kTP + tid] = v;
        }
      }
#pragma unroll
      for (int i = 0; i < P; ++i) sx[i * kTP + tid] = x[i];
      __syncthreads();
      pair_reduce_f<EH * K>(accW, [&](int pair, int r4) {
        const int j = pair % K, ch = pair / K;
        float acc = 0.f;
#pragma unroll
        for (int p = 0; p < L1; ++p) {
          const int q = p + j - D::pad1;
          if (q >= 0 && q < P) acc += dot4(ld4(sdc + (ch * L1 + p) * kTP + r4), ld4(sx + q * kTP + r4));
        }
        return acc;
      });
      __syncthreads();
    }
  }
  __syncthreads();
  ENC_STAMP(2)
  constexpr int nstat = PH == 0 ? EH : PH == 1 ? E : PH == 2 ? C : PH == 4 ? C : PH == 5 ? E : PH == 6 ? EH : 0;
  if (nstat) {
    double* dst = PH == 0 ? a.st : PH == 1 ? a.st + 2 * EH : PH == 2 ? a.st + 2 * (EH + E) : PH == 4 ? Bq3 : PH == 5 ? Bq2 : Bq1;
    for (int i = tid; i < 2 * nstat; i += kT) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < kT / 32; ++w) v += (double)sacc[w][i];
      atomicAdd(&dst[i], v);
    }
  }
  if (PH == 5) {
    for (int i = tid; i < C * EL2; i += kT) atomicAdd(&a.dW3[i], accW[i]);
    for (int i = tid; i < C; i += kT) atomicAdd(&a.db3[i], accW[C * EL2 + i]);
  } else if (PH == 6) {
    for (int i = tid; i < E * EH * K; i += kT) atomicAdd(&a.dW2[i], accW[i]);
  } else if (PH == 7) {
    for (int i = tid; i < EH * K; i += kT) atomicAdd(&a.dW1[i], accW[i]);
  }
#ifdef STG_ENC_TIMING
  __syncthreads();
  ENC_STAMP(3)
#endif
}

template <class D, int PH>
size_t stage_bytes(int T) {
  size_t fl = 0;
  if (PH == 3 || PH == 8) fl = (size_t)T * D::C;
  if (PH == 4) fl = (size_t)2 * (4 * D::C + T);
  if (PH == 5) fl = (size_t)(D::C + D::EL2) * kTP;
  if (PH == 6) fl = (size_t)(D::EL2 + D::NL1) * kTP;
  if (PH == 7) fl = (size_t)(D::NL1 + D::P) * kTP;
  return fl * 4;
}

const int kProfOfF[9] = {kProfEncF1, kProfEncF2, kProfEncF3, kProfEncF4, kProfEncB1, kProfEncB2, kProfEncB3, kProfEncB4,
                         kProfEncF4};

int sms_f() {
  static int v = 0;
  if (!v) {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    v = n;
  }
  return v;
}

template <class D, int PH>
void launch_one(const EncArgs& a, cudaStream_t s) {
  const size_t smem = stage_bytes<D, PH>(a.T);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(k_enc_fast<D, PH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    attr_done = true;
  }
  const int ntiles = (a.R + kT - 1) / kT;
  int per_sm = smem > 56 * 1024 ? 2 : smem > 24 * 1024 ? 4 : 8;
  if (const char* e = getenv("STG_ENC_PERSM")) { const int v = atoi(e); if (v >= 1) per_sm = v; }
  int grid = sms_f() * per_sm;
  if (grid > ntiles) grid = ntiles;
  ProfScope ps(kProfOfF[PH], s);
  launch_pdl(k_enc_fast<D, PH>, dim3(grid), dim3(kT), smem, s, a);
}

template <class D>
bool matches(const EncArgs& a) {
  return a.P == D::P && a.K == D::K && a.EH == D::EH && a.E == D::E && a.C == D::C &&
         (size_t)a.T * D::C * 4 <= 100 * 1024;
}

template <class D>
void run(const EncArgs& a, bool backward, cudaStream_t s) {
  if (!backward) {
    if (a.training) {
      launch_one<D, 0>(a, s);
      launch_one<D, 1>(a, s);
      launch_one<D, 2>(a, s);
      launch_one<D, 3>(a, s);
    } else {
      launch_one<D, 8>(a, s);           // eval: whole chain in one pass from running statistics
    }
  } else {
    launch_one<D, 4>(a, s);
    launch_one<D, 5>(a, s);
    launch_one<D, 6>(a, s);
    launch_one<D, 7>(a, s);
  }
}

// configs/hparams.py FC_STGNN sets with register-sized rows (FD001 / N-CMAPSS use the generic kernels)
using D_FD004 = EncDims<2, 2, 8, 6, 16>;   // hparams.py:149-151  (S1)
using D_S2 = EncDims<1, 2, 8, 6, 14>;      // BASELINE synthetic   (S2)
using D_FD002 = EncDims<1, 2, 8, 12, 16>;  // hparams.py:69-71
using D_FD003 = EncDims<1, 2, 8, 6, 48>;   // hparams.py:109-111

}  // namespace

#ifdef STG_ENC_TIMING
extern "C" int stg_debug_enc_cta_times(unsigned long long* out) {
  return cudaMemcpyFromSymbol(out, g_enc_time, sizeof(g_enc_time)) == cudaSuccess ? 0 : -1;
}
#endif

bool encoder_fast_available(const EncArgs& a) {
  static const bool off = getenv("STG_ENC_GENERIC") != nullptr;      // tests: force the any-dimension kernels
  if (off) return false;
  return matches<D_FD004>(a) || matches<D_S2>(a) || matches<D_FD002>(a) || matches<D_FD003>(a);
}

int launch_encoder_fast(const EncArgs& a, bool backward, cudaStream_t s) {
  if (matches<D_FD004>(a)) run<D_FD004>(a, backward, s);
  else if (matches<D_S2>(a)) run<D_S2>(a, backward, s);
  else if (matches<D_FD002>(a)) run<D_FD002>(a, backward, s);
  else if (matches<D_FD003>(a)) run<D_FD003>(a, backward, s);
  else return -2;
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

}  // namespace stg
