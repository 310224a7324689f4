// Graph-conv block kernels (sm_100a) == GraphConvpoolMPNN_block_v6
// (reference models/FC_STGNN/Model_Base.py:175-225; math restated in SURVEY.md section 9).
//
// Re-associations relative to the reference (all exact in real arithmetic):
//   * F = Linear_map(x) and V = BN0(x).Wtheta^T are computed ONCE per (b,t,n) row -- the reference
//     recomputes them for every window that contains the time step (unfold first, :194-201).
//   * BN0 is folded into the theta projection: V = x.W'^T + b',  W' = Wtheta.diag(g0*r0).
//   * theta(A.Xb) = A.(Xb.Wtheta^T) + btheta, so the aggregation runs over H (not C) features.
//   * BN0 batch statistics come from per-time-step moments of x weighted by the number of
//     windows covering each step (no unfolded tensor is ever materialised).
//
// Launch sequence (training): [xmoments] -> fwd_main<TRAIN> -> fwd_fin ;
//                             bwd_stats -> bwd_main -> bwd_fin.     Eval: fwd_main<EVAL> only.
#include "stg_block.cuh"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>

namespace stg {

namespace {

STG_DEVINL int cover_count(int t, int w, int s, int L) {
  int c = 0;
  for (int j = 0; j < w; ++j) {
    const int d = t - j;
    if (d >= 0 && d % s == 0 && d / s < L) ++c;
  }
  return c;
}

// ------------------------------------------------------------------------------------------
// per-(t,c) moments of x over (b,n):  xmom[t*C+c] += sum x, xmom[T*C+t*C+c] += sum x^2
// grid (T, BCH); block 256.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_xmoments(const float* __restrict__ x, int B, int T, int N, int C,
                                                  double* __restrict__ xmom) {
  extern __shared__ float sm[];   // [2*C]
  const int t = blockIdx.x;
  const int nb = gridDim.y;
  const int bper = (B + nb - 1) / nb;
  const int b0 = blockIdx.y * bper, b1 = min(B, b0 + bper);
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int NC = N * C;
  for (int e = threadIdx.x; e < NC; e += blockDim.x) {
    float s1 = 0.f, s2 = 0.f;
    for (int b = b0; b < b1; ++b) {
      const float v = x[((size_t)b * T + t) * NC + e];
      s1 += v;
      s2 += v * v;
    }
    const int c = e % C;
    atomicAdd(&sm[c], s1);
    atomicAdd(&sm[C + c], s2);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicAdd(&xmom[t * C + c], (double)sm[c]);
    atomicAdd(&xmom[(size_t)T * C + t * C + c], (double)sm[C + c]);
  }
}

// ------------------------------------------------------------------------------------------
// shared-memory carve-up common to the forward / backward main kernels
// ------------------------------------------------------------------------------------------
// window scratch; in the backward kernel it is re-used for the [CPH][CP]+[CPH] parameter-gradient tile
// backward kernel: pack the windows' threads back to back instead of one warp-aligned slot per window?
__host__ __device__ inline bool bwd_packed(int M) { return M > 32 && (M & 31) != 0; }

__host__ __device__ inline int sb_floats(int wpc, int slot, int CP, int CPH, bool bwd) {
  int n = wpc * slot;
  if (bwd && n < CPH * CP + CPH) n = CPH * CP + CPH;
  return (n + 3) / 4 * 4;
}

template <int CP, int HP>
struct Carve {
  static constexpr int CPH = CP + HP;
  uint64_t* bar;
  float *mu0, *r0, *a0, *c0, *biasc, *bn1c, *pw, *WcT, *xs, *FV, *Sb;
  // backward extras
  float *WmS, *WtS, *dFV, *invs, *dYs, *red;
  __device__ Carve(unsigned char* base, int rows_max, int C, int wpc, int slot, int M, bool bwd) {
    bar = reinterpret_cast<uint64_t*>(base);
    float* p = reinterpret_cast<float*>(base + 16);
    mu0 = p; p += CP;
    r0 = p; p += CP;
    a0 = p; p += CP;
    c0 = p; p += CP;
    biasc = p; p += CPH;
    bn1c = p; p += 8 * HP;
    pw = p; p += 4;
    WcT = p; p += CP * CPH;
    xs = p; p += ((rows_max * C + 4 + 3) / 4) * 4 + 4;
    FV = p; p += rows_max * CPH;
    Sb = p; p += sb_floats(wpc, slot, CP, CP + HP, bwd);
    if (bwd) {
      WmS = p; p += CP * CP;
      WtS = p; p += HP * CP;
      dFV = p; p += rows_max * CPH;
      invs = p; p += ((wpc * M + 3) / 4) * 4;
      dYs = p; p += wpc * M * HP;
      red = p; p += 2 * CP + 2 * HP;
    }
  }
};

static size_t carve_bytes(int CP, int HP, int rows_max, int C, int wpc, int slot, int M, bool bwd) {
  const int CPH = CP + HP;
  size_t f = 4 * CP + CPH + 8 * HP + 4 + (size_t)CP * CPH + (((size_t)rows_max * C + 4 + 3) / 4) * 4 + 4 +
             (size_t)rows_max * CPH + (size_t)sb_floats(wpc, slot, CP, CPH, bwd);
  if (bwd)
    f += (size_t)CP * CP + (size_t)HP * CP + (size_t)rows_max * CPH + (((size_t)wpc * M + 3) / 4) * 4 +
         (size_t)wpc * M * HP + 2 * CP + 2 * HP;
  return 16 + f * 4;
}

// ------------------------------------------------------------------------------------------
// Per-block coefficient table (training): everything the main kernels' prologues need that depends
// only on the parameters and the BN0 batch moments, computed ONCE per launch sequence instead of
// once per CTA:   mu0[CP] r0[CP] a0[CP] c0[CP] biasc[CPH] pw[4] WcT[CP*CPH] cnt[T]
// grid (nblk), block 256.  Also updates the BN0 running statistics.
// ------------------------------------------------------------------------------------------
__host__ __device__ inline int coef_floats(int CP, int HP, int T) { return 4 * CP + (CP + HP) + 4 + CP * (CP + HP) + T; }

// `tid` of `nt` threads work on block z (the callers give every block its own slice of the CTA and its own scratch,
// so that the tables of both blocks are built side by side); all threads of the CTA must call it (CTA barriers).
STG_DEVINL void block_prep_body(const BlkArgs& a, int z, int CP, int HP, const double* xmom, double* ssum,
                                double* ssq, float* sa0, float* sc0, float* sWt, int tid, int nt) {
  const BlkDev& k = a.b[z];
  const int C = a.C, T = a.T, H = k.H, CPH = CP + HP;
  float* tab = k.coef;
  float* mu0 = tab; float* r0 = mu0 + CP; float* a0 = r0 + CP; float* c0 = a0 + CP;
  float* biasc = c0 + CP; float* pw = biasc + CPH; float* WcT = pw + 4; float* cnt = WcT + CP * CPH;
  if (tid < 48) { ssum[tid] = 0.0; ssq[tid] = 0.0; }
  for (int i = tid; i < H * C; i += nt) sWt[i] = k.Wt[i];      // staged: the bias loop below walks rows of it
  __syncthreads();
  // weighted moments of x over time: thread (c, slice of t)
  {
    const int c = tid % CP, sl = tid / CP, nsl = nt / CP;
    if (c < C && sl < nsl) {
      double s = 0.0, q = 0.0;
      for (int t = sl; t < T; t += nsl) {
        const int cn = cover_count(t, k.w, k.stride, k.L);
        if (cn) {
          s += cn * __ldcg(xmom + t * C + c);            // written by other CTAs' atomics: read at L2
          q += cn * __ldcg(xmom + (size_t)T * C + t * C + c);
        }
      }
      atomicAdd(&ssum[c], s);
      atomicAdd(&ssq[c], q);
    }
  }
  for (int t = tid; t < T; t += nt) cnt[t] = (float)cover_count(t, k.w, k.stride, k.L);
  if (tid < 4) pw[tid] = powf(k.decay, (float)tid);
  __syncthreads();
  if (tid < CP) {
    const int c = tid;
    float mean = 0.f, r = 0.f, av = 0.f, cv = 0.f;
    if (c < C) {
      const double R = (double)a.B * k.L * k.w * a.N;
      const double m = ssum[c] / R;
      double var = ssq[c] / R - m * m;
      if (var < 0.0) var = 0.0;
      const double unb = R > 1.0 ? var * R / (R - 1.0) : var;
      k.rm0[c] = (1.f - a.momentum) * k.rm0[c] + a.momentum * (float)m;
      k.rv0[c] = (1.f - a.momentum) * k.rv0[c] + a.momentum * (float)unb;
      mean = (float)m;
      r = (float)(1.0 / sqrt(var + (double)a.eps));
      av = k.g0[c] * r;
      cv = k.b0[c] - av * mean;
    }
    mu0[c] = mean; r0[c] = r; a0[c] = av; c0[c] = cv;
    sa0[c] = av; sc0[c] = cv;
  }
  __syncthreads();
  for (int idx = tid; idx < CP * CPH; idx += nt) {
    const int c = idx / CPH, o = idx % CPH;
    float v = 0.f;
    if (c < C) {
      if (o < C) v = k.Wm[o * C + c];
      else if (o >= CP && o - CP < H) v = sWt[(o - CP) * C + c] * sa0[c];
    }
    WcT[idx] = v;
  }
  for (int o = tid; o < CPH; o += nt) {
    float v = 0.f;
    if (o < C) v = k.bm[o];
    else if (o >= CP && o - CP < H) {
      const float* wr = sWt + (o - CP) * C;
      for (int c = 0; c < C; ++c) v += wr[c] * sc0[c];
    }
    biasc[o] = v;
  }
}

__global__ void __launch_bounds__(256) k_block_prep(const BlkArgs a, int CP, int HP) {
  __shared__ double ssum[48], ssq[48];
  __shared__ float sa0[48], sc0[48], sWt[24 * 48];
  block_prep_body(a, blockIdx.x, CP, HP, a.xmom, ssum, ssq, sa0, sc0, sWt, threadIdx.x, 256);
}

// both blocks' tables at once by one CTA of 256 threads: block z on threads [128 z, 128 z + 128)
struct PrepScratch {
  double ssum[2][48], ssq[2][48];
  float sa0[2][48], sc0[2][48], sWt[2][24 * 48];
};
STG_DEVINL void block_prep_all(const BlkArgs& a, int CP, int HP, const double* xmom, PrepScratch& ps) {
  if (a.nblk == 2 && CP <= 64) {
    const int z = threadIdx.x >> 7;
    block_prep_body(a, z, CP, HP, xmom, ps.ssum[z], ps.ssq[z], ps.sa0[z], ps.sc0[z], ps.sWt[z], threadIdx.x & 127, 128);
  } else {
    for (int z = 0; z < a.nblk; ++z) {
      block_prep_body(a, z, CP, HP, xmom, ps.ssum[0], ps.ssq[0], ps.sa0[0], ps.sc0[0], ps.sWt[0], threadIdx.x, 256);
      __syncthreads();
    }
  }
}

// x-moments and, in the last CTA to finish, the coefficient tables of every block: one launch
// instead of two on the critical path of the training forward.  grid (T, BCH); block 256.
__global__ void __launch_bounds__(256) k_xmoments_prep(const BlkArgs a, int CP, int HP, double* xmom,
                                                       unsigned* counter) {
  __shared__ float sm[2 * 48];
  __shared__ PrepScratch ps;
  __shared__ int s_last;
  const int B = a.B, T = a.T, N = a.N, C = a.C;
  const int t = blockIdx.x, nb = gridDim.y;
  const int bper = (B + nb - 1) / nb;
  const int b0 = blockIdx.y * bper, b1 = min(B, b0 + bper);
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int NC = N * C;
  for (int e = threadIdx.x; e < NC; e += blockDim.x) {
    float s1 = 0.f, s2 = 0.f;
#pragma unroll 4
    for (int b = b0; b < b1; ++b) {
      const float v = a.x[((size_t)b * T + t) * NC + e];
      s1 += v;
      s2 = fmaf(v, v, s2);
    }
    const int c = e % C;
    atomicAdd(&sm[c], s1);
    atomicAdd(&sm[C + c], s2);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicAdd(&xmom[t * C + c], (double)sm[c]);
    atomicAdd(&xmom[(size_t)T * C + t * C + c], (double)sm[C + c]);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(counter, 1u) == gridDim.x * gridDim.y - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  block_prep_all(a, CP, HP, xmom, ps);
}

// Same job for C == 16 (every reference hyper-parameter set with hidden_dim 8), without shared-memory atomics:
// one warp walks the [N,16] slabs of ONE time step over a slice of the batch; lane l always meets feature l % 16
// (the slab is read as consecutive floats and 32 % 16 == 0), so the moments stay in two registers per lane, lanes l and
// l + 16 are folded with one shuffle and every warp issues 32 double atomics in all.  grid (ceil(T * wpt / 8)), 256.
__global__ void __launch_bounds__(256) k_xmoments_prep16(const BlkArgs a, int CP, int HP, double* xmom,
                                                         unsigned* counter, int nslice) {
  pdl_sync();
  __shared__ PrepScratch ps;
  __shared__ float red[8][32];
  __shared__ int s_last;
  const int B = a.B, T = a.T, N = a.N;
  const int t = blockIdx.x % T, sl = blockIdx.x / T;         // CTA = (time step, batch slice); warp = sample stride
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const int NC = N * 16;
    float s1 = 0.f, s2 = 0.f;
    for (int b = sl * 8 + warp; b < B; b += nslice * 8) {
      const float* p = a.x + ((size_t)b * T + t) * NC;
      // all loads of the slab first (independent), then the sums: the kernel is latency bound
      float v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = (lane + 32 * u < NC) ? __ldg(p + lane + 32 * u) : 0.f;
      for (int e = lane + 512; e < NC; e += 32) {           // N > 32 sensors: the rest of the slab
        const float w = __ldg(p + e);
        s1 += w;
        s2 = fmaf(w, w, s2);
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        s1 += v[u];
        s2 = fmaf(v[u], v[u], s2);
      }
    }
    s1 += __shfl_down_sync(0xffffffffu, s1, 16);
    s2 += __shfl_down_sync(0xffffffffu, s2, 16);
    if (lane < 16) {
      red[warp][lane] = s1;
      red[warp][16 + lane] = s2;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
      const int c = threadIdx.x & 15;
      atomicAdd(&xmom[(threadIdx.x < 16 ? 0 : (size_t)T * 16) + t * 16 + c], (double)v);
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(counter, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  block_prep_all(a, CP, HP, xmom, ps);
}

// BN0 statistics (batch or running), folded projection [Wm | Wtheta.diag(g0 r0)]^T and biases.
template <int CP, int HP, bool TRAIN>
STG_DEVINL void block_prologue(const BlkArgs& a, const BlkDev& k, Carve<CP, HP>& sm) {
  constexpr int CPH = CP + HP;
  const int C = a.C, H = k.H, tid = threadIdx.x, nt = blockDim.x;
  if (TRAIN) {
    // table written by k_block_prep: mu0 r0 a0 c0 biasc | pw | WcT   (smem: ... biasc bn1c pw WcT)
    const float* tab = k.coef;
    for (int i = tid; i < 4 * CP + CPH; i += nt) sm.mu0[i] = tab[i];
    if (tid < 4) sm.pw[tid] = tab[4 * CP + CPH + tid];
    const float4* src = reinterpret_cast<const float4*>(tab + 4 * CP + CPH + 4);
    float4* dst = reinterpret_cast<float4*>(sm.WcT);
    for (int i = tid; i < CP * CPH / 4; i += nt) dst[i] = src[i];
    return;
  }
  if (tid < CP) {
    const int c = tid;
    float mean = 0.f, r = 0.f, av = 0.f, cv = 0.f;
    if (c < C) {
      mean = k.rm0[c];
      r = (float)(1.0 / sqrt((double)k.rv0[c] + (double)a.eps));
      av = k.g0[c] * r;
      cv = k.b0[c] - av * mean;
    }
    sm.mu0[c] = mean;
    sm.r0[c] = r;
    sm.a0[c] = av;
    sm.c0[c] = cv;
  }
  if (tid < 4) sm.pw[tid] = powf(k.decay, (float)tid);
  __syncthreads();
  for (int idx = tid; idx < CP * CPH; idx += nt) {
    const int c = idx / CPH, o = idx % CPH;
    float v = 0.f;
    if (c < C) {
      if (o < C) v = k.Wm[o * C + c];
      else if (o >= CP && o - CP < H) v = k.Wt[(o - CP) * C + c] * sm.a0[c];
    }
    sm.WcT[idx] = v;
  }
  for (int o = tid; o < CPH; o += nt) {
    float v = 0.f;
    if (o < C) v = k.bm[o];
    else if (o >= CP && o - CP < H) {
      const float* wr = k.Wt + (o - CP) * C;
      for (int c = 0; c < C; ++c) v += wr[c] * sm.c0[c];
    }
    sm.biasc[o] = v;
  }
}

// FV[row][0:CP] = F = x.Wm^T + bm ; FV[row][CP:CP+HP] = V = BN0(x).Wtheta^T   (no btheta)
template <int CP, int HP>
STG_DEVINL void compute_fv(const Carve<CP, HP>& sm, const float* xs, int rows, int C) {
  constexpr int CPH = CP + HP, G = CPH / 4;
  for (int item = threadIdx.x; item < rows * G; item += blockDim.x) {
    const int row = item / G, og = item % G;
    float4 acc = *reinterpret_cast<const float4*>(sm.biasc + og * 4);
    const float* xr = xs + row * C;
    const float* wp = sm.WcT + og * 4;
    for (int c = 0; c < C; ++c) {
      const float xv = xr[c];
      const float4 w4 = *reinterpret_cast<const float4*>(wp + c * CPH);
      acc.x = fmaf(xv, w4.x, acc.x);
      acc.y = fmaf(xv, w4.y, acc.y);
      acc.z = fmaf(xv, w4.z, acc.z);
      acc.w = fmaf(xv, w4.w, acc.w);
    }
    *reinterpret_cast<float4*>(sm.FV + row * CPH + og * 4) = acc;
  }
}

template <int CP>
STG_DEVINL float dot_cp(const float (&a)[CP], const float* __restrict__ b) {
  // four independent accumulators: the kernel is latency bound, a single fmaf chain of CP links is not
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int c = 0; c < CP; c += 4) {
    const float4 v = *reinterpret_cast<const float4*>(b + c);
    s0 = fmaf(a[c], v.x, s0);
    s1 = fmaf(a[c + 1], v.y, s1);
    s2 = fmaf(a[c + 2], v.z, s2);
    s3 = fmaf(a[c + 3], v.w, s3);
  }
  return (s0 + s1) + (s2 + s3);
}

// ------------------------------------------------------------------------------------------
// forward main: one CTA = (chunk of windows, sample b, block z)
// ------------------------------------------------------------------------------------------
template <int CP, int HP, bool TRAIN>
__global__ void __launch_bounds__(256) k_block_fwd(const BlkArgs a, int rows_max, int wpc, int slot) {
  constexpr int CPH = CP + HP;
  extern __shared__ __align__(16) unsigned char smraw[];
  const BlkDev& k = a.b[blockIdx.z];
  const int chunk = blockIdx.x;
  if (chunk >= k.nchunk_f) return;
  const int b = blockIdx.y;
  const int N = a.N, C = a.C, T = a.T, H = k.H, w = k.w, s = k.stride, L = k.L, M = w * N;
  const int per = (L + k.nchunk_f - 1) / k.nchunk_f;
  const int l0 = chunk * per, l1 = min(L, l0 + per);
  if (l0 >= l1) return;
  const int t_lo = l0 * s, t_hi = (l1 - 1) * s + w - 1;
  const int rows = (t_hi - t_lo + 1) * N;
  const int tid = threadIdx.x;

  Carve<CP, HP> sm(smraw, rows_max, C, wpc, slot, M, false);
  if (tid == 0) mbar_init(sm.bar, 1);
  __syncthreads();
  const int shift = stage_floats_tma(sm.xs, a.x + ((size_t)b * T + t_lo) * N * C, rows * C, sm.bar, tid);
  const float* xs = sm.xs + shift;

  block_prologue<CP, HP, TRAIN>(a, k, sm);
  // BN1 coefficients (eval only): yn = a1*y + c1
  if (!TRAIN && tid < HP) {
    float a1 = 0.f, c1 = 0.f;
    if (tid < H) {
      const float r1 = (float)(1.0 / sqrt((double)k.rv1[tid] + (double)a.eps));
      a1 = k.g1[tid] * r1;
      c1 = k.b1[tid] - a1 * k.rm1[tid];
    }
    sm.bn1c[tid] = a1;
    sm.bn1c[HP + tid] = c1;
  }
  mbar_wait(sm.bar, 0);
  __syncthreads();
  compute_fv<CP, HP>(sm, xs, rows, C);
  __syncthreads();

  const int slot_id = tid / M, i = tid - slot_id * M;
  const int ji = i / N;
  const bool lane_ok = tid < wpc * M;
  float* Sb = sm.Sb + slot_id * slot;
  float st1[HP], st2[HP];
#pragma unroll
  for (int h = 0; h < HP; ++h) st1[h] = st2[h] = 0.f;
  const float invw = 1.f / (float)w;

  for (int lb = l0; lb < l1; lb += wpc) {
    const int l = lb + slot_id;
    const bool act = lane_ok && l < l1;
    float y[HP];
    if (act) {
      const float* FVw = sm.FV + (size_t)(l * s - t_lo) * N * CPH;
      float Fi[CP];
#pragma unroll
      for (int c = 0; c < CP; c += 4) {
        const float4 v = *reinterpret_cast<const float4*>(FVw + i * CPH + c);
        Fi[c] = v.x; Fi[c + 1] = v.y; Fi[c + 2] = v.z; Fi[c + 3] = v.w;
      }
      float* Srow = Sb + i * (M + 1);
      float mx = -INFINITY;
#pragma unroll 2
      for (int kk = 0; kk < M; ++kk) {
        const float sv = lrelu(dot_cp<CP>(Fi, FVw + kk * CPH));
        Srow[kk] = sv;
        if (kk != i) mx = fmaxf(mx, sv);
      }
      float acc[HP];
#pragma unroll
      for (int h = 0; h < HP; ++h) acc[h] = 0.f;
      float sum = 0.f;
      int kk = 0;
      for (int j2 = 0; j2 < w; ++j2) {
        const float mk = sm.pw[abs(j2 - ji)];
        for (int n2 = 0; n2 < N; ++n2, ++kk) {
          if (kk == i) continue;
          const float e = __expf(Srow[kk] - mx);
          sum += e;
          const float em = e * mk;
          const float* Vk = FVw + kk * CPH + CP;
#pragma unroll
          for (int h = 0; h < HP; h += 4) {
            const float4 v = *reinterpret_cast<const float4*>(Vk + h);
            acc[h] = fmaf(em, v.x, acc[h]);
            acc[h + 1] = fmaf(em, v.y, acc[h + 1]);
            acc[h + 2] = fmaf(em, v.z, acc[h + 2]);
            acc[h + 3] = fmaf(em, v.w, acc[h + 3]);
          }
        }
      }
      const float inv = sum > 0.f ? 1.f / sum : 0.f;
      const float* Vi = FVw + i * CPH + CP;
#pragma unroll
      for (int h = 0; h < HP; ++h) y[h] = (h < H) ? fmaf(acc[h], inv, Vi[h] + k.bt[h]) : 0.f;
      if (TRAIN) {
        float* yrow = k.yp + (((size_t)b * L + l) * M + i) * H;
#pragma unroll
        for (int h = 0; h < HP; ++h)
          if (h < H) {
            yrow[h] = y[h];
            st1[h] += y[h];
            st2[h] = fmaf(y[h], y[h], st2[h]);
          }
      }
    }
    if (!TRAIN) {
      // BN1 (running stats) + leaky_relu + mean over the w rows of each sensor
      __syncthreads();
      if (act) {
#pragma unroll
        for (int h = 0; h < HP; ++h) Sb[i * HP + h] = lrelu(fmaf(sm.bn1c[h], y[h], sm.bn1c[HP + h]));
      }
      __syncthreads();
      if (act) {
        float* orow = k.out + (size_t)b * k.out_bs + (size_t)l * N * H;
        for (int e = i; e < N * H; e += M) {
          const int n = e / H, h = e - n * H;
          float v = 0.f;
          for (int j = 0; j < w; ++j) v += Sb[(j * N + n) * HP + h];
          orow[e] = v * invw;
        }
      }
      __syncthreads();
    }
  }
  if (TRAIN) {
    // CTA-level reduction of the BN1 moments, then one double atomic per feature
    __syncthreads();
    float* red = sm.Sb;   // reuse
    if (tid < 2 * HP) red[tid] = 0.f;
    __syncthreads();
#pragma unroll
    for (int h = 0; h < HP; ++h) {
      const float v1 = warp_sum(st1[h]), v2 = warp_sum(st2[h]);
      if ((tid & 31) == 0) {
        atomicAdd(&red[h], v1);
        atomicAdd(&red[HP + h], v2);
      }
    }
    __syncthreads();
    if (tid < H) {
      atomicAdd(&k.stats[tid], (double)red[tid]);
      atomicAdd(&k.stats[H + tid], (double)red[HP + tid]);
    }
  }
}

// BN1 batch-stat coefficients from the accumulated moments. tid < H computes; result in smem c[]:
//   c[0*HP+h]=a1  c[1*HP+h]=c1  c[2*HP+h]=mu1  c[3*HP+h]=r1
STG_DEVINL void bn1_coeffs(const BlkArgs& a, const BlkDev& k, float* c, int HPs, bool update_running) {
  const int h = threadIdx.x;
  if (h < k.H) {
    const double R = (double)a.B * k.L * k.w * a.N;
    const double m = k.stats[h] / R;
    double var = k.stats[k.H + h] / R - m * m;
    if (var < 0.0) var = 0.0;
    const float r1 = (float)(1.0 / sqrt(var + (double)a.eps));
    const float a1 = k.g1[h] * r1;
    c[h] = a1;
    c[HPs + h] = k.b1[h] - a1 * (float)m;
    c[2 * HPs + h] = (float)m;
    c[3 * HPs + h] = r1;
    if (update_running) {
      const double unb = R > 1.0 ? var * R / (R - 1.0) : var;
      k.rm1[h] = (1.f - a.momentum) * k.rm1[h] + a.momentum * (float)m;
      k.rv1[h] = (1.f - a.momentum) * k.rv1[h] + a.momentum * (float)unb;
    }
  }
}

// ------------------------------------------------------------------------------------------
// forward finalize (training): out = mean_j lrelu(BN1(Y'))        grid (ceil(B*L*N*H/256), 1, nblk)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_block_fwd_fin(const BlkArgs a) {
  __shared__ float c[4 * 64];
  const BlkDev& k = a.b[blockIdx.z];
  const int H = k.H, N = a.N, L = k.L, w = k.w, M = w * N;
  const long long total = (long long)a.B * L * N * H;
  if ((long long)blockIdx.x * blockDim.x >= total) return;
  bn1_coeffs(a, k, c, 64, blockIdx.x == 0);
  __syncthreads();
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int h = (int)(e % H);
  const long long r = e / H;           // (b*L + l)*N + n
  const int n = (int)(r % N);
  const long long bl = r / N;
  const int l = (int)(bl % L);
  const long long b = bl / L;
  const float a1 = c[h], c1 = c[64 + h];
  const float* yp = k.yp + ((size_t)bl * M + n) * H + h;
  float v = 0.f;
  for (int j = 0; j < w; ++j) v += lrelu(fmaf(a1, yp[(size_t)j * N * H], c1));
  k.out[(size_t)b * k.out_bs + ((size_t)l * N + n) * H + h] = v / (float)w;
}

// ------------------------------------------------------------------------------------------
// backward stats: sum_R dYn, sum_R dYn*Yhat per h.   grid (nCTA, 1, nblk), thread per row (b,l,m)
// ------------------------------------------------------------------------------------------
template <int HP>
__global__ void __launch_bounds__(256) k_block_bwd_stats(const BlkArgs a) {
  __shared__ float c[4 * 64];
  __shared__ float red[2 * HP];
  const BlkDev& k = a.b[blockIdx.z];
  const int H = k.H, N = a.N, L = k.L, w = k.w, M = w * N;
  bn1_coeffs(a, k, c, 64, false);
  if (threadIdx.x < 2 * HP) red[threadIdx.x] = 0.f;
  __syncthreads();
  const long long rows = (long long)a.B * L * M;
  float s1[HP], s2[HP];
#pragma unroll
  for (int h = 0; h < HP; ++h) s1[h] = s2[h] = 0.f;
  const float invw = 1.f / (float)w;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows;
       r += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(r % M);
    const long long bl = r / M;
    const int l = (int)(bl % L);
    const long long b = bl / L;
    const int n = m % N;
    const float* y = k.yp + (size_t)r * H;
    const float* d = k.dout + (size_t)b * k.dout_bs + ((size_t)l * N + n) * H;
#pragma unroll
    for (int h = 0; h < HP; ++h)
      if (h < H) {
        const float yv = y[h];
        const float yn = fmaf(c[h], yv, c[64 + h]);
        const float dyn = d[h] * invw * lrelu_grad(yn);
        s1[h] += dyn;
        s2[h] = fmaf(dyn, (yv - c[128 + h]) * c[192 + h], s2[h]);
      }
  }
#pragma unroll
  for (int h = 0; h < HP; ++h) {
    const float v1 = warp_sum(s1[h]), v2 = warp_sum(s2[h]);
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&red[h], v1);
      atomicAdd(&red[HP + h], v2);
    }
  }
  __syncthreads();
  if (threadIdx.x < H) {
    atomicAdd(&k.stats[2 * H + threadIdx.x], (double)red[threadIdx.x]);
    atomicAdd(&k.stats[3 * H + threadIdx.x], (double)red[HP + threadIdx.x]);
  }
}

// ------------------------------------------------------------------------------------------
// backward main: one CTA = (chunk of time steps [ta,tb), sample b, block z)
// ------------------------------------------------------------------------------------------
template <int CP, int HP>
__global__ void __launch_bounds__(256) k_block_bwd(const BlkArgs a, int rows_max, int wpc, int slot) {
  constexpr int CPH = CP + HP;
  extern __shared__ __align__(16) unsigned char smraw[];
  const BlkDev& k = a.b[blockIdx.z];
  const int chunk = blockIdx.x;
  if (chunk >= k.nchunk_b) return;
  const int b = blockIdx.y;
  const int N = a.N, C = a.C, T = a.T, H = k.H, w = k.w, s = k.stride, L = k.L, M = w * N;
  int per = (T + k.nchunk_b - 1) / k.nchunk_b;
  per = ((per + s - 1) / s) * s;                       // chunk boundaries on window starts
  const int ta = chunk * per, tb = min(T, ta + per);
  if (ta >= tb) return;
  // windows touching [ta,tb):  l*s <= tb-1  and  l*s + w-1 >= ta
  int l_lo = ta - (w - 1);
  l_lo = l_lo <= 0 ? 0 : (l_lo + s - 1) / s;
  int l_hi = min(L - 1, (tb - 1) / s);
  const bool any_win = l_lo <= l_hi;
  const int t_lo = any_win ? min(ta, l_lo * s) : ta;
  const int t_hi = any_win ? max(tb - 1, l_hi * s + w - 1) : tb - 1;
  const int rows = (t_hi - t_lo + 1) * N;
  const int tid = threadIdx.x, nt = blockDim.x;

  Carve<CP, HP> sm(smraw, rows_max, C, wpc, slot, M, true);
  if (tid == 0) mbar_init(sm.bar, 1);
  __syncthreads();
  const int shift = stage_floats_tma(sm.xs, a.x + ((size_t)b * T + t_lo) * N * C, rows * C, sm.bar, tid);
  const float* xs = sm.xs + shift;

  block_prologue<CP, HP, true>(a, k, sm);
  // raw Wm [o][c] and Wtheta [h][c] for the transposed products
  for (int idx = tid; idx < CP * CP; idx += nt) {
    const int o = idx / CP, c = idx % CP;
    sm.WmS[idx] = (o < C && c < C) ? k.Wm[o * C + c] : 0.f;
  }
  for (int idx = tid; idx < HP * CP; idx += nt) {
    const int h = idx / CP, c = idx % CP;
    sm.WtS[idx] = (h < H && c < C) ? k.Wt[h * C + c] : 0.f;
  }
  // BN1 backward coefficients: bn1c[0]=a1 [1]=c1 [2]=mu1 [3]=r1 [4]=g1*r1 [5]=q1 [6]=q2
  if (tid < HP) {
    float v[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (tid < H) {
      const int h = tid;
      const double R = (double)a.B * L * M;
      const double m = k.stats[h] / R;
      double var = k.stats[H + h] / R - m * m;
      if (var < 0.0) var = 0.0;
      const float r1 = (float)(1.0 / sqrt(var + (double)a.eps));
      const float g1 = k.g1[h];
      v[0] = g1 * r1;
      v[1] = k.b1[h] - v[0] * (float)m;
      v[2] = (float)m;
      v[3] = r1;
      v[4] = g1 * r1;
      v[5] = (float)(g1 * k.stats[2 * H + h] / R) * r1;
      v[6] = (float)(g1 * k.stats[3 * H + h] / R) * r1;
    }
#pragma unroll
    for (int q = 0; q < 7; ++q) sm.bn1c[q * HP + tid] = v[q];
  }
  for (int idx = tid; idx < rows * CPH; idx += nt) sm.dFV[idx] = 0.f;
  for (int idx = tid; idx < 2 * CP + 2 * HP; idx += nt) sm.red[idx] = 0.f;
  mbar_wait(sm.bar, 0);
  __syncthreads();
  compute_fv<CP, HP>(sm, xs, rows, C);
  __syncthreads();

  // one window per slot of ST threads.  M <= 32 (and multiples of 32): ST = roundup32(M), the phases of a
  // window are separated by slot-local barriers (__syncwarp / named barrier).  Otherwise (M = 40, 42: the 21-
  // and 20-sensor shapes) a 64-thread slot would idle a third of its lanes, so the windows are packed back to
  // back, ST = M, and the phases meet at a CTA barrier instead.
  const bool packed = a.bwd_packed != 0;
  const int ST = packed ? M : ((M + 31) / 32) * 32;
  const int slot_id = tid / ST, i = tid - slot_id * ST;
  const int ji = i / N, ni = i - ji * N;
  const bool lane_ok = slot_id < wpc && i < M;
  const int sslot = slot_id < wpc ? slot_id : 0;          // surplus threads (none by construction) alias slot 0
  float* Sb = sm.Sb + sslot * slot;
  float* invs = sm.invs + sslot * M;
  float* dYs = sm.dYs + (size_t)sslot * M * HP;
  const float invw = 1.f / (float)w;
  auto slot_sync = [&]() {
    if (packed) __syncthreads();
    else if (ST == 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(slot_id + 1), "r"(ST) : "memory");
  };
  float dbt_acc[HP];
#pragma unroll
  for (int h = 0; h < HP; ++h) dbt_acc[h] = 0.f;

  if (any_win)
    for (int lb = l_lo; lb <= l_hi; lb += wpc) {
      const int l = lb + slot_id;
      const bool act = lane_ok && l <= l_hi;
      const float* FVw = sm.FV + (size_t)(act ? (l * s - t_lo) : 0) * N * CPH;
      float* Srow = Sb + i * (M + 1);
      float dY[HP];
      float inv = 0.f, rs = 0.f;
      if (act) {
        const float* yrow = k.yp + (((size_t)b * L + l) * M + i) * H;
        const float* drow = k.dout + (size_t)b * k.dout_bs + ((size_t)l * N + ni) * H;
        const bool own_win = (l * s >= ta) && (l * s < tb);
#pragma unroll
        for (int h = 0; h < HP; ++h) {
          float v = 0.f;
          if (h < H) {
            const float yv = yrow[h];
            const float yn = fmaf(sm.bn1c[h], yv, sm.bn1c[HP + h]);
            const float dyn = drow[h] * invw * lrelu_grad(yn);
            const float yh = (yv - sm.bn1c[2 * HP + h]) * sm.bn1c[3 * HP + h];
            v = sm.bn1c[4 * HP + h] * dyn - sm.bn1c[5 * HP + h] - yh * sm.bn1c[6 * HP + h];
            if (own_win) dbt_acc[h] += v;
          }
          dY[h] = v;
          dYs[i * HP + h] = v;
        }
        float Fi[CP];
#pragma unroll
        for (int c = 0; c < CP; c += 4) {
          const float4 v = *reinterpret_cast<const float4*>(FVw + i * CPH + c);
          Fi[c] = v.x; Fi[c + 1] = v.y; Fi[c + 2] = v.z; Fi[c + 3] = v.w;
        }
        float mx = -INFINITY;
#pragma unroll 2
        for (int kk = 0; kk < M; ++kk) {
          const float sv = lrelu(dot_cp<CP>(Fi, FVw + kk * CPH));
          Srow[kk] = sv;
          if (kk != i) mx = fmaxf(mx, sv);
        }
        float sum = 0.f;
        int kk = 0;
        for (int j2 = 0; j2 < w; ++j2) {
          const float mk = sm.pw[abs(j2 - ji)];
          for (int n2 = 0; n2 < N; ++n2, ++kk) {
            if (kk == i) { Srow[kk] = 0.f; continue; }
            const float sv = Srow[kk];
            const float e = __expf(sv - mx);
            sum += e;
            const float* Vk = FVw + kk * CPH + CP;
            float dA0 = 0.f, dA1 = 0.f;
#pragma unroll
            for (int h = 0; h < HP; h += 2) {
              dA0 = fmaf(dY[h], Vk[h], dA0);
              dA1 = fmaf(dY[h + 1], Vk[h + 1], dA1);
            }
            const float dA = dA0 + dA1;
            rs = fmaf(e * mk, dA, rs);
            Srow[kk] = sv > 0.f ? e : -e;
          }
        }
        inv = sum > 0.f ? 1.f / sum : 0.f;
        rs *= inv;
        invs[i] = inv;
      }
      slot_sync();
      // dV_k = sum_i A[i][k] dY'_i   (this thread plays column k = i)
      float dV[HP];
#pragma unroll
      for (int h = 0; h < HP; ++h) dV[h] = 0.f;
      if (act) {
        int i2 = 0;
        for (int j2 = 0; j2 < w; ++j2) {
          const float mk = sm.pw[abs(j2 - ji)];
          for (int n2 = 0; n2 < N; ++n2, ++i2) {
            float p = fabsf(Sb[i2 * (M + 1) + i]) * invs[i2];
            if (i2 == i) p += 1.f;
            p *= mk;
            const float* dy2 = dYs + i2 * HP;
#pragma unroll
            for (int h = 0; h < HP; ++h) dV[h] = fmaf(p, dy2[h], dV[h]);
          }
        }
      }
      slot_sync();
      // dS in place (row i)
      if (act) {
        int kk = 0;
        for (int j2 = 0; j2 < w; ++j2) {
          const float mk = sm.pw[abs(j2 - ji)];
          for (int n2 = 0; n2 < N; ++n2, ++kk) {
            if (kk == i) continue;
            const float ev = Srow[kk];
            const float P = fabsf(ev) * inv;
            const float* Vk = FVw + kk * CPH + CP;
            float dA0 = 0.f, dA1 = 0.f;
#pragma unroll
            for (int h = 0; h < HP; h += 2) {
              dA0 = fmaf(dY[h], Vk[h], dA0);
              dA1 = fmaf(dY[h + 1], Vk[h + 1], dA1);
            }
            const float dA = dA0 + dA1;
            const float dLam = P * (dA * mk - rs);
            Srow[kk] = dLam * (ev > 0.f ? 1.f : kLeaky);
          }
        }
      }
      slot_sync();
      float dF[CP];
#pragma unroll
      for (int c = 0; c < CP; ++c) dF[c] = 0.f;
      if (act) {
        for (int kk = 0; kk < M; ++kk) {
          const float g = Srow[kk] + Sb[kk * (M + 1) + i];
          const float* Fk = FVw + kk * CPH;
#pragma unroll
          for (int c = 0; c < CP; c += 4) {
            const float4 v = *reinterpret_cast<const float4*>(Fk + c);
            dF[c] = fmaf(g, v.x, dF[c]);
            dF[c + 1] = fmaf(g, v.y, dF[c + 1]);
            dF[c + 2] = fmaf(g, v.z, dF[c + 2]);
            dF[c + 3] = fmaf(g, v.w, dF[c + 3]);
          }
        }
      }
      // fold into the per-time-step accumulators, one window offset at a time: rows of equal offset are
      // distinct across the slots, consecutive offsets are ordered by the CTA barrier (no atomics)
      for (int jj = 0; jj < w; ++jj) {
        if (act && ji == jj) {
          float* dst = sm.dFV + ((size_t)(l * s - t_lo) * N + i) * CPH;
#pragma unroll
          for (int c = 0; c < CP; ++c) dst[c] += dF[c];
#pragma unroll
          for (int h = 0; h < HP; ++h) dst[CP + h] += dV[h];
        }
        __syncthreads();
      }
    }
  __syncthreads();

  // ---- owned rows: dx partial + BN0 backward sums.  thread = (row slot, 4 feature columns):
  //      [dF | dV] row (CPH values, broadcast reads) x [Wm ; Wtheta] columns (float4 reads)
  const int r_beg = (ta - t_lo) * N, r_end = (tb - t_lo) * N;
  {
    constexpr int NQ = CP / 4;
    const int cq = tid % NQ, rsl = tid / NQ, nrs = nt / NQ, c0i = cq * 4;
    float sb[4] = {0.f, 0.f, 0.f, 0.f}, sg[4] = {0.f, 0.f, 0.f, 0.f};
    if (rsl < nrs && c0i < C) {
      float* dxp = k.dxp + ((size_t)b * T + ta) * N * C;
      for (int r = r_beg + rsl; r < r_end; r += nrs) {
        const float* d = sm.dFV + (size_t)r * CPH;
        float pf[4] = {0.f, 0.f, 0.f, 0.f}, dx4[4] = {0.f, 0.f, 0.f, 0.f};
        for (int o = 0; o < C; ++o) {
          const float dv = d[o];
          const float4 w4 = *reinterpret_cast<const float4*>(sm.WmS + o * CP + c0i);
          pf[0] = fmaf(dv, w4.x, pf[0]); pf[1] = fmaf(dv, w4.y, pf[1]);
          pf[2] = fmaf(dv, w4.z, pf[2]); pf[3] = fmaf(dv, w4.w, pf[3]);
        }
#pragma unroll
        for (int h = 0; h < HP; ++h) {
          const float dv = d[CP + h];
          const float4 w4 = *reinterpret_cast<const float4*>(sm.WtS + h * CP + c0i);
          dx4[0] = fmaf(dv, w4.x, dx4[0]); dx4[1] = fmaf(dv, w4.y, dx4[1]);
          dx4[2] = fmaf(dv, w4.z, dx4[2]); dx4[3] = fmaf(dv, w4.w, dx4[3]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = c0i + u;
          if (c < C) {
            const float xh = (xs[r * C + c] - sm.mu0[c]) * sm.r0[c];
            sb[u] += dx4[u];
            sg[u] = fmaf(dx4[u], xh, sg[u]);
            dxp[(size_t)(r - r_beg) * C + c] = fmaf(dx4[u], sm.a0[c], pf[u]);
          }
        }
      }
    }
    // reduce over the row slots of each warp first (lanes NQ apart share the columns), then one
    // shared atomic per warp and column -- not one per thread
    if (NQ <= 32 && (32 % NQ) == 0) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float v1 = sb[u], v2 = sg[u];
        for (int o = NQ; o < 32; o <<= 1) {
          v1 += __shfl_xor_sync(0xffffffffu, v1, o);
          v2 += __shfl_xor_sync(0xffffffffu, v2, o);
        }
        if ((tid & 31) < NQ && c0i + u < C) {
          atomicAdd(&sm.red[c0i + u], v1);
          atomicAdd(&sm.red[CP + c0i + u], v2);
        }
      }
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (rsl < nrs && c0i + u < C) {
          atomicAdd(&sm.red[c0i + u], sb[u]);
          atomicAdd(&sm.red[CP + c0i + u], sg[u]);
        }
    }
  }
#pragma unroll
  for (int h = 0; h < HP; ++h) {
    const float v = warp_sum(dbt_acc[h]);
    if ((tid & 31) == 0 && h < H) atomicAdd(&sm.red[2 * CP + h], v);
  }
  // ---- parameter gradients: G[o][c] = sum_rows dFV[r][o]*x[r][c],  so[o] = sum_rows dFV[r][o]
  //      work item = (o, 4 columns), rows sliced over the remaining threads; partial sums meet in
  //      shared memory (the window scratch is free now), one global atomic per entry per CTA.
  {
    constexpr int NQ = CP / 4;
    // Gs[slice][CPH][CP] + so[slice][CPH]: every (item, slice) has one owner thread -> plain stores
    float* Gs = sm.Sb;
    constexpr int GSZ = CPH * CP + CPH;
    const int items = CPH * NQ;
    int nsl = nt >= items ? nt / items : 1;
    {
      const int fit = sb_floats(wpc, slot, CP, CPH, true) / GSZ;
      if (nsl > fit) nsl = fit;
      if (nsl > 8) nsl = 8;
    }
    for (int it = tid; it < items * nsl; it += nt) {
      const int item = it % items, sl = it / items;
      const int o = item / NQ, c0i = (item - o * NQ) * 4;
      float g4[4] = {0.f, 0.f, 0.f, 0.f}, so = 0.f;
      for (int r = r_beg + sl; r < r_end; r += nsl) {
        const float dv = sm.dFV[(size_t)r * CPH + o];
        const float2 xa = *reinterpret_cast<const float2*>(xs + r * C + c0i);
        const float2 xb = *reinterpret_cast<const float2*>(xs + r * C + c0i + 2);
        g4[0] = fmaf(dv, xa.x, g4[0]); g4[1] = fmaf(dv, xa.y, g4[1]);
        g4[2] = fmaf(dv, xb.x, g4[2]); g4[3] = fmaf(dv, xb.y, g4[3]);
        so += dv;
      }
      float* Gp = Gs + sl * GSZ;
#pragma unroll
      for (int u = 0; u < 4; ++u) Gp[o * CP + c0i + u] = g4[u];
      if (c0i == 0) Gp[CPH * CP + o] = so;
    }
    __syncthreads();
    for (int idx = tid; idx < CPH * C; idx += nt) {
      const int o = idx / C, c = idx - o * C;
      const bool isF = o < C, isV = (o >= CP && o - CP < H);
      if (!isF && !isV) continue;
      float g = 0.f, so = 0.f;
      for (int q = 0; q < nsl; ++q) {
        g += Gs[q * GSZ + o * CP + c];
        so += Gs[q * GSZ + CPH * CP + o];
      }
      if (isF) {
        atomicAdd(&k.dWm[o * C + c], g);
        if (c == 0) atomicAdd(&k.dbm[o], so);
      } else {
        atomicAdd(&k.dWt[(o - CP) * C + c], fmaf(sm.a0[c], g, sm.c0[c] * so));
      }
    }
  }
  __syncthreads();
  if (tid < C) {
    atomicAdd(&k.stats[4 * H + tid], (double)sm.red[tid]);
    atomicAdd(&k.stats[4 * H + C + tid], (double)sm.red[CP + tid]);
  }
  if (tid < H) atomicAdd(&k.dbt[tid], sm.red[2 * CP + tid]);
}

// ------------------------------------------------------------------------------------------
// backward finalize: dx = sum_blk dxp_blk - r0*cnt(t)*(m1 + Xhat*m2); BN affine grads.
// grid (ceil(B*T*N*C/256)); dynamic smem: nblk * (4*C + T) floats
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_block_bwd_fin(const BlkArgs a, int CP) {
  extern __shared__ float smf[];
  const int C = a.C, T = a.T, N = a.N;
  const int per = 4 * C + T;
  for (int z = 0; z < a.nblk; ++z) {
    const BlkDev& k = a.b[z];
    float* tb = smf + z * per;
    const double R = (double)a.B * k.L * k.w * N;
    const float* tab = k.coef;                 // mu0[CP] r0[CP] ...
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float r0 = tab[CP + c];
      const float g0 = k.g0[c];
      const double sb = k.stats[4 * k.H + c], sg = k.stats[4 * k.H + C + c];
      tb[c] = tab[c];
      tb[C + c] = r0;
      tb[2 * C + c] = r0 * (float)(g0 * sb / R);
      tb[3 * C + c] = r0 * (float)(g0 * sg / R);
      if (blockIdx.x == 0) {
        k.db0[c] += (float)sb;
        k.dg0[c] += (float)sg;
      }
    }
    for (int t = threadIdx.x; t < T; t += blockDim.x) tb[4 * C + t] = (float)cover_count(t, k.w, k.stride, k.L);
    if (blockIdx.x == 0)
      for (int h = threadIdx.x; h < k.H; h += blockDim.x) {
        k.db1[h] += (float)k.stats[2 * k.H + h];
        k.dg1[h] += (float)k.stats[3 * k.H + h];
      }
  }
  __syncthreads();
  const long long total = (long long)a.B * T * N * C;
  for (long long e4 = (long long)blockIdx.x * blockDim.x + threadIdx.x; e4 * 4 < total;
       e4 += (long long)gridDim.x * blockDim.x) {
    const long long e0 = e4 * 4;
    float xv[4], out[4];
    const int nv = (int)((total - e0) < 4 ? (total - e0) : 4);
    if (nv == 4) {
      const float4 v = *reinterpret_cast<const float4*>(a.x + e0);
      xv[0] = v.x; xv[1] = v.y; xv[2] = v.z; xv[3] = v.w;
    } else {
      for (int u = 0; u < nv; ++u) xv[u] = a.x[e0 + u];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) out[u] = 0.f;
    for (int z = 0; z < a.nblk; ++z) {
      const float* tb = smf + z * per;
      float dv[4];
      if (a.dxp_unfolded) {
        // one row per (window, node): sum the (at most w) windows covering this time step
        const BlkDev& k = a.b[z];
        for (int u = 0; u < 4; ++u) {
          dv[u] = 0.f;
          if (u >= nv) continue;
          const long long e = e0 + u;
          const int c = (int)(e % C);
          const long long r = e / C;
          const int n = (int)(r % N);
          const int t = (int)((r / N) % T);
          const long long b = r / ((long long)N * T);
          for (int j = 0; j < k.w; ++j) {
            const int d = t - j;
            if (d < 0 || d % k.stride || d / k.stride >= k.L) continue;
            dv[u] += k.dxp[(((size_t)b * k.L + d / k.stride) * (k.w * N) + j * N + n) * C + c];
          }
        }
      } else if (nv == 4) {
        const float4 v = *reinterpret_cast<const float4*>(a.b[z].dxp + e0);
        dv[0] = v.x; dv[1] = v.y; dv[2] = v.z; dv[3] = v.w;
      } else {
        for (int u = 0; u < nv; ++u) dv[u] = a.b[z].dxp[e0 + u];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (u < nv) {
          const long long e = e0 + u;
          const int c = (int)(e % C);
          const int t = (int)((e / ((long long)N * C)) % T);
          const float xh = (xv[u] - tb[c]) * tb[C + c];
          out[u] += dv[u] - tb[4 * C + t] * (tb[2 * C + c] + xh * tb[3 * C + c]);
        }
    }
    if (nv == 4) *reinterpret_cast<float4*>(a.dx + e0) = make_float4(out[0], out[1], out[2], out[3]);
    else for (int u = 0; u < nv; ++u) a.dx[e0 + u] = out[u];
  }
}

// ------------------------------------------------------------------------------------------
// host side: template dispatch + launch plan
// ------------------------------------------------------------------------------------------
struct Variant {
  int CP, HP;
  void (*fwd_train)(const BlkArgs, int, int, int);
  void (*fwd_eval)(const BlkArgs, int, int, int);
  void (*bwd)(const BlkArgs, int, int, int);
  void (*bwd_stats)(const BlkArgs);
};
#define STG_VARIANT(CP, HP) \
  { CP, HP, k_block_fwd<CP, HP, true>, k_block_fwd<CP, HP, false>, k_block_bwd<CP, HP>, k_block_bwd_stats<HP> }
static const Variant kVariants[] = {
    // same (CP, HP) set as the tensor-core path (stg_block_mma.cu): both read one coefficient table
    STG_VARIANT(8, 8), STG_VARIANT(16, 8), STG_VARIANT(32, 16), STG_VARIANT(48, 24),
};
constexpr int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);
constexpr size_t kSmemCap = 200 * 1024;

static const Variant* pick_variant(int C, int H) {
  for (int i = 0; i < kNumVariants; ++i)
    if (kVariants[i].CP >= C && kVariants[i].HP >= H) return &kVariants[i];
  return nullptr;
}

static bool g_attr_done[64] = {};
static void set_smem_attrs() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (g_attr_done[dev]) return;
  for (int i = 0; i < kNumVariants; ++i) {
    cudaFuncSetAttribute(kVariants[i].fwd_train, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemCap);
    cudaFuncSetAttribute(kVariants[i].fwd_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemCap);
    cudaFuncSetAttribute(kVariants[i].bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemCap);
  }
  g_attr_done[dev] = true;
}

}  // namespace

int plan_blocks(BlkArgs& a, BlkPlan& p, char* err, size_t errlen) {
  if (a.nblk < 1 || a.nblk > 2) { snprintf(err, errlen, "nblk must be 1 or 2"); return -1; }
  int Hmax = 0, Mmax = 0;
  for (int z = 0; z < a.nblk; ++z) {
    BlkDev& k = a.b[z];
    if (k.w < 1 || k.w > kMaxWin || k.stride < 1 || a.T < k.w) {
      snprintf(err, errlen, "unsupported window %d / stride %d for T=%d", k.w, k.stride, a.T);
      return -2;
    }
    k.L = (a.T - k.w) / k.stride + 1;
    if (k.H > 64) { snprintf(err, errlen, "hidden_dim %d > 64 unsupported", k.H); return -2; }
    Hmax = k.H > Hmax ? k.H : Hmax;
    Mmax = k.w * a.N > Mmax ? k.w * a.N : Mmax;
  }
  if (a.nblk == 2 && a.b[0].w * a.N != a.b[1].w * a.N) {
    snprintf(err, errlen, "blocks fused in one launch must share the window size"); return -2;
  }
  const Variant* v = pick_variant(a.C, Hmax);
  if (!v) { snprintf(err, errlen, "no kernel variant for C=%d H=%d (max C 48, H 24)", a.C, Hmax); return -2; }
  const int M = Mmax;
  if (M > 256) { snprintf(err, errlen, "w*N=%d nodes per graph > 256 unsupported", M); return -2; }
  p.CP = v->CP; p.HP = v->HP;
  p.mma_f = 0; p.mma_b = 0; p.NT = 0; p.tc = 0; p.tc_wr = 0;
  a.dxp_unfolded = 0;
  const int slot = (M * (M + 1) > M * v->HP ? M * (M + 1) : M * v->HP);
  int wpc = 256 / M; if (wpc < 1) wpc = 1;
  // ---- forward: windows per chunk == windows in flight unless shared memory says otherwise
  for (int wp = wpc; wp >= 1; --wp) {
    int rows_max = 0, gx = 0;
    for (int z = 0; z < a.nblk; ++z) {
      BlkDev& k = a.b[z];
      k.nchunk_f = (k.L + wp - 1) / wp;
      const int per = (k.L + k.nchunk_f - 1) / k.nchunk_f;
      const int rows = ((per - 1) * k.stride + k.w) * a.N;
      rows_max = rows > rows_max ? rows : rows_max;
      gx = k.nchunk_f > gx ? k.nchunk_f : gx;
    }
    const size_t sm = carve_bytes(v->CP, v->HP, rows_max, a.C, wp, slot, M, false);
    if (sm <= kSmemCap || wp == 1) {
      if (sm > kSmemCap) { snprintf(err, errlen, "forward tile does not fit shared memory (%zu B)", sm); return -2; }
      p.wpc_f = wp; p.smem_f = sm; p.grid_x_f = gx; p.threads_f = ((wp * M + 31) / 32) * 32;
      break;
    }
  }
  // tensor-core forward when the graph fits the instantiated tiles (overrides the SIMT chunking)
  if (!getenv("STG_NO_MMA")) plan_blocks_mma_fwd(a, p);
  // ---- backward: time steps per chunk so that the touching windows fit one pass.  Two sweeps: first
  //      look for the largest tile that still lets 3 CTAs share an SM (the kernel is latency bound,
  //      resident warps matter more than tile size), then for anything that fits at all.  Blocks
  //      with a larger stride get fewer steps per chunk so that they do not dictate the slab size.
  {
    bool done = false;
    size_t cap0 = 74 * 1024;
    if (const char* e = getenv("STG_BWD_CAP_KB")) { const int v2 = atoi(e); if (v2 >= 16) cap0 = (size_t)v2 * 1024; }
    const bool packed = bwd_packed(M) && !getenv("STG_BWD_NOPACK");
    a.bwd_packed = packed ? 1 : 0;
    // packed shapes: 6 windows of 40-42 nodes per CTA need ~100 KB; 2 CTAs/SM with full slots beat 3 with
    // 4-window chunks (S2: 1085 -> 1012 us per step)
    if (packed && !getenv("STG_BWD_CAP_KB")) cap0 = 113 * 1024;
    const size_t caps[2] = {cap0, kSmemCap};
    const int ST = packed ? M : ((M + 31) / 32) * 32;     // threads per window slot in the backward kernel
    int wpc_b0 = 256 / ST;
    if (wpc_b0 < 1) wpc_b0 = 1;
    if (wpc_b0 > 15) wpc_b0 = 15;                   // named barriers 1..15
    // window passes per chunk (each pass = wp windows): more passes amortise the per-CTA prologue / epilogue and
    // the duplicated boundary window over more windows, at the price of a larger slab.  Measured: worth it for the
    // packed shapes (S2 1012 -> 945 us per step with 2), neutral or worse for M = 28.
    int passes = packed ? 2 : 1;
    if (const char* e = getenv("STG_BWD_PASSES")) { const int v2 = atoi(e); if (v2 >= 1 && v2 <= 8) passes = v2; }
    for (int sweep = 0; sweep < 2 && !done; ++sweep)
      for (int wp = wpc_b0; wp >= 1 && !done; --wp) {
        int gx = 0;
        const int step_cap = passes * wp + 1;           // staged time steps of a stride-1, w=2 chunk
        for (int z = 0; z < a.nblk; ++z) {
          BlkDev& k = a.b[z];
          int per = passes * wp * k.stride - (k.w - 1);
          if (per > step_cap - (k.w - 1) && step_cap - (k.w - 1) >= k.stride) per = step_cap - (k.w - 1);
          if (per < k.stride) per = k.stride;
          per = (per / k.stride) * k.stride;
          k.nchunk_b = (a.T + per - 1) / per;
          gx = k.nchunk_b > gx ? k.nchunk_b : gx;
        }
        int wmax = 0;
        const int rows_max = bwd_rows_exact(a, &wmax);
        if (wmax > passes * wp) continue;               // a chunk would touch more windows than its passes hold
        const size_t sm = carve_bytes(v->CP, v->HP, rows_max, a.C, wp, slot, M, true);
        if (sm <= caps[sweep]) {
          p.wpc_b = wp; p.smem_b = sm; p.grid_x_b = gx; p.threads_b = ((wp * ST + 31) / 32) * 32;
          if (p.threads_b < 64) p.threads_b = 64;       // the epilogue indexes up to C = 48 threads; surplus threads idle in the window loop
          done = true;
        }
      }
    if (!done) { snprintf(err, errlen, "backward tile does not fit shared memory"); return -2; }
  }
  // Blackwell path: tcgen05.mma + TMEM for both directions whenever the shape fits (overrides the above)
  if (plan_blocks_tc(a, p)) {
    p.mma_f = 0; p.mma_b = 0;
    a.dxp_unfolded = 1;
  }
  return 0;
}

static int rows_max_fwd(const BlkArgs& a) {
  int rm = 0;
  for (int z = 0; z < a.nblk; ++z) {
    const BlkDev& k = a.b[z];
    const int per = (k.L + k.nchunk_f - 1) / k.nchunk_f;
    const int rows = ((per - 1) * k.stride + k.w) * a.N;
    rm = rows > rm ? rows : rm;
  }
  return rm;
}
static int rows_max_bwd(const BlkArgs& a) { return bwd_rows_exact(a); }

int launch_xmoments(const float* x, int B, int T, int N, int C, double* xmom, cudaStream_t s) {
  cudaMemsetAsync(xmom, 0, sizeof(double) * 2 * T * C, s);
  int nb = (B + 15) / 16;
  if (nb > 64) nb = 64;
  {
    ProfScope ps(kProfXmoments, s);
    k_xmoments<<<dim3(T, nb), 256, 2 * C * sizeof(float), s>>>(x, B, T, N, C, xmom);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

// model path: x-moments + coefficient tables in one launch (xmom and counter must be zero on entry)
int launch_xmoments_prep(const BlkArgs& a, const BlkPlan& p, double* xmom, unsigned* counter, cudaStream_t s) {
  int nb = (a.B + 15) / 16;
  if (nb > 64) nb = 64;
  ProfScope ps(kProfXmoments, s);
  if (a.C == 16 && !getenv("STG_XMOM_GENERIC")) {
    // batch slices per time step: ~4 CTAs of 8 warps per SM, at most one sample per warp
    int nslice = (148 * 4 + a.T - 1) / a.T;
    if (nslice * 8 > a.B) nslice = (a.B + 7) / 8;
    if (nslice < 1) nslice = 1;
    launch_pdl(k_xmoments_prep16, dim3(a.T * nslice), dim3(256), 0, s, a, p.CP, p.HP, xmom, counter, nslice);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
  }
  k_xmoments_prep<<<dim3(a.T, nb), 256, 0, s>>>(a, p.CP, p.HP, xmom, counter);
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

int launch_block_forward(const BlkArgs& a, const BlkPlan& p, cudaStream_t s) {
  set_smem_attrs();
  const Variant* v = pick_variant(p.CP, p.HP);
  const int M = a.b[0].w * a.N;
  const int slot = (M * (M + 1) > M * p.HP ? M * (M + 1) : M * p.HP);
  const int rows_max = rows_max_fwd(a);
  dim3 grid(p.grid_x_f, a.B, a.nblk);
  if (a.training) {
    if (!a.prep_done)      // model path: the step's first kernel already cleared the whole scratch region
      for (int z = 0; z < a.nblk; ++z)
        cudaMemsetAsync(a.b[z].stats, 0, sizeof(double) * (4 * a.b[z].H + 2 * a.C), s);
    if (!a.prep_done) {
      ProfScope ps(kProfBlkPrep, s);
      k_block_prep<<<a.nblk, 256, 0, s>>>(a, p.CP, p.HP);
    }
    if (p.tc) {
      launch_block_forward_tc(a, p, s);
    } else if (p.mma_f) {
      launch_block_forward_mma(a, p, s);
    } else {
      ProfScope ps(kProfFwdMain, s);
      v->fwd_train<<<grid, p.threads_f, p.smem_f, s>>>(a, rows_max, p.wpc_f, slot);
    }
    if (!a.head_fused) {
      long long tot = 0;
      for (int z = 0; z < a.nblk; ++z) {
        const long long t = (long long)a.B * a.b[z].L * a.N * a.b[z].H;
        tot = t > tot ? t : tot;
      }
      ProfScope ps(kProfFwdFin, s);
      k_block_fwd_fin<<<dim3((unsigned)((tot + 255) / 256), 1, a.nblk), 256, 0, s>>>(a);
    }
  } else if (p.tc) {
    launch_block_forward_tc(a, p, s);
  } else if (p.mma_f) {
    launch_block_forward_mma(a, p, s);
  } else {
    ProfScope ps(kProfFwdMain, s);
    v->fwd_eval<<<grid, p.threads_f, p.smem_f, s>>>(a, rows_max, p.wpc_f, slot);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

int launch_block_backward(const BlkArgs& a, const BlkPlan& p, cudaStream_t s) {
  set_smem_attrs();
  const Variant* v = pick_variant(p.CP, p.HP);
  const int M = a.b[0].w * a.N;
  const int slot = (M * (M + 1) > M * p.HP ? M * (M + 1) : M * p.HP);
  const int rows_max = rows_max_bwd(a);
  for (int z = 0; z < a.nblk; ++z) {
    if (a.head_fused)      // model path: [2H,4H) was filled by k_head_bwd1, the rest is still zero from the
      continue;            // forward's clear (one backward per training forward)
    else
      cudaMemsetAsync(a.b[z].stats + 2 * a.b[z].H, 0, sizeof(double) * (2 * a.b[z].H + 2 * a.C), s);
  }
  long long rows = 0;
  for (int z = 0; z < a.nblk; ++z) {
    const long long r = (long long)a.B * a.b[z].L * M;
    rows = r > rows ? r : rows;
  }
  int g = (int)((rows + 255) / 256);
  if (g > 592) g = 592;
  if (!a.head_fused) {
    ProfScope ps(kProfBwdStats, s);
    v->bwd_stats<<<dim3(g, 1, a.nblk), 256, 0, s>>>(a);
  }
  if (p.tc) {
    launch_block_backward_tc(a, p, s);
  } else {
    ProfScope ps(kProfBwdMain, s);
    v->bwd<<<dim3(p.grid_x_b, a.B, a.nblk), p.threads_b, p.smem_b, s>>>(a, rows_max, p.wpc_b, slot);
  }
  if (a.fin_elsewhere) return cudaGetLastError() == cudaSuccess ? 0 : -3;
  const long long tot = (long long)a.B * a.T * a.N * a.C;
  ProfScope ps(kProfBwdFin, s);
  long long gfin = (tot / 4 + 255) / 256;
  if (gfin > 148 * 8) gfin = 148 * 8;
  if (gfin < 1) gfin = 1;
  k_block_bwd_fin<<<(unsigned)gfin, 256, a.nblk * (4 * a.C + a.T) * sizeof(float), s>>>(a, p.CP);
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

}  // namespace stg
