// Graph-conv block, tensor-core path (sm_100a): the per-graph products of GraphConvpoolMPNN_block_v6
// (reference models/FC_STGNN/Model_Base.py:175-225; math in SURVEY.md section 9) as warp-level
// m16n8k8 3xTF32 MMAs -- one warp owns one window (graph) at a time:
//   forward   FV = x.[Wm | Wtheta']^T (projection, once per time step, stored pre-split hi/lo)
//             -> S = F.F^T -> row softmax in the accumulator fragments (quad shuffles)
//             -> Y' = ((P+I).mask).V   (P re-used as the A operand through a k-permutation)
//   backward  recompute S,P -> dA = dY'.V^T -> dS in fragments; A~ then dS staged per warp in shared
//             memory -> dV = A~^T.dY', dF = (dS+dS^T).F -> folded per time step with shared atomics
//             -> tail on tensor cores: dx partial = dF.Wm + a0*(dV.Wtheta), BN0 sums, dWm / dWtheta
// Instruction economy (ncu r01 v4: 97 % of the first version's instructions were operand splitting,
// bounds guards and index math): operands are split once into tf32 hi/lo planes in shared memory,
// pad rows make every fragment load unguarded, contraction indices are permuted (k'=t <-> 2t,
// k'=t+4 <-> 2t+1) so that fragment pairs are one 64-bit shared load, per-lane column masks are
// precomputed.  The SIMT kernels in stg_block.cu remain the fallback for w*N > 64 nodes per graph.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "stg_block.cuh"
#include "stg_mma.cuh"

namespace stg {
namespace {

constexpr size_t kSmemCapM = 200 * 1024;

__host__ __device__ inline int fv_pitch(int CPH) { return (CPH % 8 == 4) ? CPH : CPH + 4; }
__host__ __device__ inline int up16(int x) { return (x + 15) & ~15; }

// ---- shared-memory carve-up (float offsets after the 16-byte barrier slot) -------------------
struct MLay {
  int tab, whi, wlo, wmh, wml, wth, wtl, bn1c, red, gs, xs, fhi, flo, dfv, scr, total;
  int FVP, SP, DP, WP, scr_per_warp, rows_alloc;
};

// mode 0: forward train, 1: forward eval, 2: backward
__host__ __device__ inline MLay make_mlay(int CP, int HP, int MP, int M, int C, int rows_max, int nwarps, int mode) {
  MLay l;
  const int CPH = CP + HP;
  l.FVP = fv_pitch(CPH);
  l.SP = MP + 4;
  l.DP = HP + 4;
  l.WP = CP + 4;
  int ra = rows_max + (MP - M);
  if (ra < up16(rows_max) + 16) ra = up16(rows_max) + 16;     // tail m-tiles may start at any owned row
  l.rows_alloc = ra;
  int o = 0;
  l.tab = o; o += 4 * CP + CPH + 4 + CP * CPH;
  l.whi = o; o += CP * l.FVP;
  l.wlo = o; o += CP * l.FVP;
  l.wmh = l.wml = l.wth = l.wtl = o;
  if (mode == 2) {
    l.wmh = o; o += CP * l.WP;
    l.wml = o; o += CP * l.WP;
    l.wth = o; o += HP * l.WP;
    l.wtl = o; o += HP * l.WP;
  }
  l.bn1c = o; o += 8 * HP;
  l.red = o; o += 2 * CP + 4 * HP + 4;
  l.gs = o;
  if (mode == 2) o += (up16(CPH) + 1) * CP + CPH + 4;
  l.xs = o; o += up16(rows_max) * C + 8;
  l.fhi = o; o += ra * l.FVP + 8;
  l.flo = o; o += ra * l.FVP + 8;
  l.dfv = o;
  if (mode == 2) o += ra * l.FVP + 8;
  l.scr = o;
  l.scr_per_warp = mode == 0 ? 0 : mode == 1 ? MP * (HP + 1) : MP * l.SP + 2 * MP * l.DP;
  l.scr_per_warp = (l.scr_per_warp + 3) / 4 * 4;
  o += nwarps * l.scr_per_warp;
  l.total = (o + 3) / 4 * 4;
  return l;
}

STG_DEVINL uint32_t hi_of(float x) { return __float_as_uint(x) & 0xffffe000u; }
STG_DEVINL void split2(float x, uint32_t& hi, uint32_t& lo) {
  hi = hi_of(x);
  lo = __float_as_uint(x - __uint_as_float(hi));      // exact; the MMA ignores its low 13 mantissa bits
}
STG_DEVINL FragA frag_a_split(float a0, float a1, float a2, float a3) {
  FragA f;
  split2(a0, f.hi[0], f.lo[0]); split2(a1, f.hi[1], f.lo[1]);
  split2(a2, f.hi[2], f.lo[2]); split2(a3, f.hi[3], f.lo[3]);
  return f;
}
STG_DEVINL FragB frag_b_split(float b0, float b1) {
  FragB f;
  split2(b0, f.hi[0], f.lo[0]); split2(b1, f.hi[1], f.lo[1]);
  return f;
}
STG_DEVINL uint2 ldu2(const float* p) { return *reinterpret_cast<const uint2*>(p); }
STG_DEVINL float2 ldf2(const float* p) { return *reinterpret_cast<const float2*>(p); }

// A fragment of rows (r0, r0+8), k-tile kk, from pre-split planes (k-permuted: pair = 64-bit load)
STG_DEVINL FragA frag_a_planes(const float* hi, const float* lo, int pitch, int r0, int kcol) {
  FragA f;
  const uint2 h0 = ldu2(hi + r0 * pitch + kcol), h1 = ldu2(hi + (r0 + 8) * pitch + kcol);
  const uint2 l0 = ldu2(lo + r0 * pitch + kcol), l1 = ldu2(lo + (r0 + 8) * pitch + kcol);
  f.hi[0] = h0.x; f.hi[2] = h0.y; f.hi[1] = h1.x; f.hi[3] = h1.y;
  f.lo[0] = l0.x; f.lo[2] = l0.y; f.lo[1] = l1.x; f.lo[3] = l1.y;
  return f;
}
// B fragment whose two k entries are adjacent in memory (row rb, columns kcol, kcol+1)
STG_DEVINL FragB frag_b_pair(const float* hi, const float* lo, int pitch, int rb, int kcol) {
  FragB f;
  const uint2 h = ldu2(hi + rb * pitch + kcol), l = ldu2(lo + rb * pitch + kcol);
  f.hi[0] = h.x; f.hi[1] = h.y; f.lo[0] = l.x; f.lo[1] = l.y;
  return f;
}
// B fragment whose two k entries are consecutive rows (rows ka, ka+1, column col)
STG_DEVINL FragB frag_b_rows(const float* hi, const float* lo, int pitch, int ka, int col) {
  FragB f;
  f.hi[0] = __float_as_uint(hi[ka * pitch + col]); f.hi[1] = __float_as_uint(hi[(ka + 1) * pitch + col]);
  f.lo[0] = __float_as_uint(lo[ka * pitch + col]); f.lo[1] = __float_as_uint(lo[(ka + 1) * pitch + col]);
  return f;
}

// ---- prologue: coefficient table into shared memory, projection weights split into hi/lo -------
template <int CP, int HP, bool TRAIN>
STG_DEVINL void load_table(const BlkArgs& a, const BlkDev& k, float* tab) {
  constexpr int CPH = CP + HP;
  const int C = a.C, H = k.H, tid = threadIdx.x, nt = blockDim.x;
  float* mu0 = tab; float* r0 = mu0 + CP; float* a0 = r0 + CP; float* c0 = a0 + CP;
  float* biasc = c0 + CP; float* pw = biasc + CPH; float* WcT = pw + 4;
  if (TRAIN) {
    const float4* src = reinterpret_cast<const float4*>(k.coef);
    float4* dst = reinterpret_cast<float4*>(tab);
    for (int i = tid; i < (4 * CP + CPH + 4 + CP * CPH) / 4; i += nt) dst[i] = src[i];
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
    mu0[c] = mean; r0[c] = r; a0[c] = av; c0[c] = cv;
  }
  if (tid < 4) pw[tid] = powf(k.decay, (float)tid);
  __syncthreads();
  for (int idx = tid; idx < CP * CPH; idx += nt) {
    const int c = idx / CPH, o = idx % CPH;
    float v = 0.f;
    if (c < C) {
      if (o < C) v = k.Wm[o * C + c];
      else if (o >= CP && o - CP < H) v = k.Wt[(o - CP) * C + c] * a0[c];
    }
    WcT[idx] = v;
  }
  for (int o = tid; o < CPH; o += nt) {
    float v = 0.f;
    if (o < C) v = k.bm[o];
    else if (o >= CP && o - CP < H) {
      const float* wr = k.Wt + (o - CP) * C;
      for (int c = 0; c < C; ++c) v += wr[c] * c0[c];
    }
    biasc[o] = v;
  }
}

template <int CP, int HP>
STG_DEVINL void split_weights(const float* tab, float* whi, float* wlo, int FVP) {
  constexpr int CPH = CP + HP;
  const float* WcT = tab + 4 * CP + CPH + 4;
  for (int idx = threadIdx.x; idx < CP * CPH; idx += blockDim.x) {
    const int c = idx / CPH, o = idx - c * CPH;
    uint32_t h, l;
    split2(WcT[idx], h, l);
    whi[c * FVP + o] = __uint_as_float(h);
    wlo[c * FVP + o] = __uint_as_float(l);
  }
}

// ---- FV[row][0:CP] = F = x.Wm^T + bm ; FV[row][CP:CP+HP] = V = BN0(x).Wtheta^T, written as hi/lo planes.
// xs must be zero-filled up to a multiple of 16 rows; columns >= C hit zero weight rows.
template <int CP, int HP>
STG_DEVINL void project_fv(const float* tab, const float* whi, const float* wlo, const float* xs, float* fhi,
                           float* flo, int FVP, int rows, int C) {
  constexpr int CPH = CP + HP, NTO = CPH / 8, KS = CP / 8;
  const float* biasc = tab + 4 * CP;
  const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int ntiles = (rows + 15) / 16;
  for (int mt = warp; mt < ntiles; mt += nw) {
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    float acc[NTO][4];
#pragma unroll
    for (int n = 0; n < NTO; ++n) {
      const float2 bv = ldf2(biasc + n * 8 + 2 * t);
      acc[n][0] = bv.x; acc[n][1] = bv.y; acc[n][2] = bv.x; acc[n][3] = bv.y;
    }
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
      const int kc = kk * 8 + 2 * t;
      const float2 x0 = ldf2(xs + r0 * C + kc), x1 = ldf2(xs + r1 * C + kc);
      const FragA fa = frag_a_split(x0.x, x1.x, x0.y, x1.y);
#pragma unroll
      for (int n = 0; n < NTO; ++n) mma3(acc[n], fa, frag_b_rows(whi, wlo, FVP, kc, n * 8 + g));
    }
#pragma unroll
    for (int n = 0; n < NTO; ++n) {
      uint32_t h0, l0, h1, l1, h2, l2, h3, l3;
      split2(acc[n][0], h0, l0); split2(acc[n][1], h1, l1); split2(acc[n][2], h2, l2); split2(acc[n][3], h3, l3);
      const int c = n * 8 + 2 * t;
      *reinterpret_cast<uint2*>(fhi + r0 * FVP + c) = make_uint2(h0, h1);
      *reinterpret_cast<uint2*>(flo + r0 * FVP + c) = make_uint2(l0, l1);
      *reinterpret_cast<uint2*>(fhi + r1 * FVP + c) = make_uint2(h2, h3);
      *reinterpret_cast<uint2*>(flo + r1 * FVP + c) = make_uint2(l2, l3);
    }
  }
}

// ---- per-lane constants of the window tiles ----------------------------------------------------
template <int NT>
struct LaneCols {
  int tcol[NT][2];   // time index (col / N) of the lane's two columns of every 8-column tile
  int dj;            // lane holds a diagonal element in tile (2*mt) regs 0/1 resp. (2*mt+1) regs 2/3: which j, or -1
};
template <int NT>
STG_DEVINL LaneCols<NT> make_lane_cols(int g, int t, int N) {
  LaneCols<NT> lc;
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    lc.tcol[n][0] = (n * 8 + 2 * t) / N;
    lc.tcol[n][1] = (n * 8 + 2 * t + 1) / N;
  }
  lc.dj = ((g >> 1) == t) ? (g & 1) : -1;
  return lc;
}

// S tile (rows mt*16+g, +8; NT*8 columns) of one window + masked row softmax in the fragments.
// p <- P (diagonal and columns >= M exactly 0); sgn bit (n*4+u) set where S > 0 (only if SGN).
template <int CP, int NT, int MT, bool SGN>
STG_DEVINL void s_tile_softmax(const float* Fh, const float* Fl, int FVP, int M, int g, int t, int dj,
                               float (&p)[NT][4], unsigned& sgn) {
  constexpr int KS = CP / 8;
  const int r0 = MT * 16 + g;
#pragma unroll
  for (int n = 0; n < NT; ++n) p[n][0] = p[n][1] = p[n][2] = p[n][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < KS; ++kk) {
    const int kc = kk * 8 + 2 * t;
    const FragA fa = frag_a_planes(Fh, Fl, FVP, r0, kc);
#pragma unroll
    for (int n = 0; n < NT; ++n) mma3(p[n], fa, frag_b_pair(Fh, Fl, FVP, n * 8 + g, kc));
  }
  float mx0 = -INFINITY, mx1 = -INFINITY;
  sgn = 0u;
#pragma unroll
  for (int n = 0; n < NT; ++n) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float sv = p[n][u];
      if (SGN && sv > 0.f) sgn |= 1u << (n * 4 + u);
      p[n][u] = fmaxf(sv, kLeaky * sv);
    }
    if (n * 8 + 8 > M) {                      // warp-uniform: tile with padded columns
      if (n * 8 + 2 * t >= M) { p[n][0] = -INFINITY; p[n][2] = -INFINITY; }
      if (n * 8 + 2 * t + 1 >= M) { p[n][1] = -INFINITY; p[n][3] = -INFINITY; }
    }
    if (n == 2 * MT) {                        // diagonal of rows r0
      if (dj == 0) p[n][0] = -INFINITY;
      if (dj == 1) p[n][1] = -INFINITY;
    }
    if (n == 2 * MT + 1) {                    // diagonal of rows r0 + 8
      if (dj == 0) p[n][2] = -INFINITY;
      if (dj == 1) p[n][3] = -INFINITY;
    }
    mx0 = fmaxf(mx0, fmaxf(p[n][0], p[n][1]));
    mx1 = fmaxf(mx1, fmaxf(p[n][2], p[n][3]));
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    p[n][0] = __expf(p[n][0] - mx0); p[n][1] = __expf(p[n][1] - mx0);
    p[n][2] = __expf(p[n][2] - mx1); p[n][3] = __expf(p[n][3] - mx1);
    s0 += p[n][0] + p[n][1];
    s1 += p[n][2] + p[n][3];
  }
  s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
  s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
  s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
  s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
  const float i0 = 1.f / s0, i1 = 1.f / s1;
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    p[n][0] *= i0; p[n][1] *= i0; p[n][2] *= i1; p[n][3] *= i1;
  }
}

// decay^|time(row) - time(col)|  (Mask_Matrix, Model_Base.py:150-170); pw = {1, d, d^2, d^3}
STG_DEVINL float mask_val(const float* pw, int tr, int tc) {
  int d = tr - tc;
  d = d < 0 ? -d : d;
  return pw[d & 3];
}

// ------------------------------------------------------------------------------------------
// forward: one CTA = (chunk of windows, sample b, block z); one warp = one window at a time
// ------------------------------------------------------------------------------------------
template <int CP, int HP, int NT, bool TRAIN>
struct FwdTile {
  static constexpr int NH = HP / 8;
  template <int MT>
  static STG_DEVINL void run(const BlkDev& k, const float* Fh, const float* Fl, int FVP, int M, int N, int H, int g,
                             int t, const LaneCols<NT>& lc, const float* pw, const float (&bt)[NH][2], float* yrow,
                             float* scr, const float* bn1c, float (&st1)[NH][2], float (&st2)[NH][2]) {
    if (MT * 16 >= M) return;
    const int r0 = MT * 16 + g, r1 = r0 + 8;
    float p[NT][4];
    unsigned sgn;
    s_tile_softmax<CP, NT, MT, false>(Fh, Fl, FVP, M, g, t, lc.dj, p, sgn);
    const int tr0 = r0 / N, tr1 = r1 / N;
    float y[NH][4];
#pragma unroll
    for (int q = 0; q < NH; ++q) y[q][0] = y[q][1] = y[q][2] = y[q][3] = 0.f;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      // A~ = (P + I) .* mask, as the A operand (k'=t <-> column 2t, k'=t+4 <-> column 2t+1)
      float a0 = p[n][0] * mask_val(pw, tr0, lc.tcol[n][0]), a2 = p[n][1] * mask_val(pw, tr0, lc.tcol[n][1]);
      float a1 = p[n][2] * mask_val(pw, tr1, lc.tcol[n][0]), a3 = p[n][3] * mask_val(pw, tr1, lc.tcol[n][1]);
      if (n == 2 * MT) { if (lc.dj == 0) a0 += 1.f; if (lc.dj == 1) a2 += 1.f; }
      if (n == 2 * MT + 1) { if (lc.dj == 0) a1 += 1.f; if (lc.dj == 1) a3 += 1.f; }
      const FragA fa = frag_a_split(a0, a1, a2, a3);
#pragma unroll
      for (int q = 0; q < NH; ++q) mma3(y[q], fa, frag_b_rows(Fh, Fl, FVP, n * 8 + 2 * t, CP + q * 8 + g));
    }
#pragma unroll
    for (int q = 0; q < NH; ++q)
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int h = q * 8 + 2 * t + (u & 1), row = (u < 2) ? r0 : r1;
        if (h < H && row < M) {
          const float v = y[q][u] + bt[q][u & 1];
          if (TRAIN) {
            yrow[(size_t)row * H + h] = v;
            st1[q][u & 1] += v;
            st2[q][u & 1] = fmaf(v, v, st2[q][u & 1]);
          } else {
            scr[row * (HP + 1) + h] = lrelu(fmaf(bn1c[h], v, bn1c[HP + h]));
          }
        }
      }
  }
};

template <int CP, int HP, int NT, bool TRAIN>
__global__ void __launch_bounds__(256) k_block_fwd_mma(const BlkArgs a, int rows_max) {
  constexpr int CPH = CP + HP, MP = NT * 8, NH = HP / 8;
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
  const int tid = threadIdx.x, warp = tid >> 5, nw = blockDim.x >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;

  const MLay lay = make_mlay(CP, HP, MP, M, C, rows_max, nw, TRAIN ? 0 : 1);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smraw);
  float* sm = reinterpret_cast<float*>(smraw + 16);
  float* tab = sm + lay.tab;
  float* bn1c = sm + lay.bn1c;
  float* red = sm + lay.red;
  float* fhi = sm + lay.fhi;
  float* flo = sm + lay.flo;
  const int FVP = lay.FVP;
  const float* pw = tab + 4 * CP + CPH;

  if (tid == 0) mbar_init(bar, 1);
  __syncthreads();
  const int shift = stage_floats_tma(sm + lay.xs, a.x + ((size_t)b * T + t_lo) * N * C, rows * C, bar, tid);
  float* xs = sm + lay.xs + shift;
  load_table<CP, HP, TRAIN>(a, k, tab);
  if (!TRAIN && tid < HP) {
    float a1 = 0.f, c1 = 0.f;
    if (tid < H) {
      const float r1 = (float)(1.0 / sqrt((double)k.rv1[tid] + (double)a.eps));
      a1 = k.g1[tid] * r1;
      c1 = k.b1[tid] - a1 * k.rm1[tid];
    }
    bn1c[tid] = a1;
    bn1c[HP + tid] = c1;
  }
  if (tid < 2 * HP) red[tid] = 0.f;
  // pad rows of the planes (read by the last windows' padded tiles) and of x (read by the projection)
  for (int i = rows * FVP + tid; i < lay.rows_alloc * FVP + 8; i += blockDim.x) { fhi[i] = 0.f; flo[i] = 0.f; }
  __syncthreads();                                   // table complete (eval path writes it in two steps)
  split_weights<CP, HP>(tab, sm + lay.whi, sm + lay.wlo, FVP);
  mbar_wait(bar, 0);
  for (int i = rows * C + tid; i < up16(rows) * C; i += blockDim.x) xs[i] = 0.f;
  __syncthreads();
  project_fv<CP, HP>(tab, sm + lay.whi, sm + lay.wlo, xs, fhi, flo, FVP, rows, C);
  __syncthreads();

  float st1[NH][2], st2[NH][2], bt[NH][2];
#pragma unroll
  for (int q = 0; q < NH; ++q)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      st1[q][j] = st2[q][j] = 0.f;
      const int h = q * 8 + 2 * t + j;
      bt[q][j] = h < H ? k.bt[h] : 0.f;
    }
  const LaneCols<NT> lc = make_lane_cols<NT>(g, t, N);
  float* scr = sm + lay.scr + warp * lay.scr_per_warp;     // eval: [MP][HP+1]
  const float invw = 1.f / (float)w;
  using Tile = FwdTile<CP, HP, NT, TRAIN>;

  for (int l = l0 + warp; l < l1; l += nw) {
    const int row0 = (l * s - t_lo) * N;
    const float* Fh = fhi + (size_t)row0 * FVP;
    const float* Fl = flo + (size_t)row0 * FVP;
    float* yrow = TRAIN ? k.yp + ((size_t)b * L + l) * M * H : nullptr;
    Tile::template run<0>(k, Fh, Fl, FVP, M, N, H, g, t, lc, pw, bt, yrow, scr, bn1c, st1, st2);
    if (NT > 2) Tile::template run<1>(k, Fh, Fl, FVP, M, N, H, g, t, lc, pw, bt, yrow, scr, bn1c, st1, st2);
    if (NT > 4) Tile::template run<2>(k, Fh, Fl, FVP, M, N, H, g, t, lc, pw, bt, yrow, scr, bn1c, st1, st2);
    if (NT > 6) Tile::template run<3>(k, Fh, Fl, FVP, M, N, H, g, t, lc, pw, bt, yrow, scr, bn1c, st1, st2);
    if (!TRAIN) {
      __syncwarp();
      float* orow = k.out + (size_t)b * k.out_bs + (size_t)l * N * H;
      for (int e = lane; e < N * H; e += 32) {
        const int n = e / H, h = e - n * H;
        float v = 0.f;
        for (int j = 0; j < w; ++j) v += scr[(j * N + n) * (HP + 1) + h];
        orow[e] = v * invw;
      }
      __syncwarp();
    }
  }
  if (TRAIN) {
#pragma unroll
    for (int q = 0; q < NH; ++q)
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        float v1 = st1[q][u], v2 = st2[q][u];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          v1 += __shfl_xor_sync(0xffffffffu, v1, o);
          v2 += __shfl_xor_sync(0xffffffffu, v2, o);
        }
        if (g == 0) {
          atomicAdd(&red[q * 8 + 2 * t + u], v1);
          atomicAdd(&red[HP + q * 8 + 2 * t + u], v2);
        }
      }
    __syncthreads();
    if (tid < H) {
      atomicAdd(&k.stats[tid], (double)red[tid]);
      atomicAdd(&k.stats[H + tid], (double)red[HP + tid]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// backward: one CTA = (chunk of time steps [ta,tb), sample b, block z); one warp = one window
// ------------------------------------------------------------------------------------------

// ------------------------------------------------------------------------------------------
// dispatch
// ------------------------------------------------------------------------------------------
typedef void (*MK)(const BlkArgs, int);
struct MVariant { int CP, HP, NT; MK fwd_train, fwd_eval; };
#define STG_MV(CP, HP, NT) \
  { CP, HP, NT, k_block_fwd_mma<CP, HP, NT, true>, k_block_fwd_mma<CP, HP, NT, false> }
const MVariant kMV[] = {
    STG_MV(8, 8, 2),   STG_MV(8, 8, 4),   STG_MV(8, 8, 8),
    STG_MV(16, 8, 2),  STG_MV(16, 8, 4),  STG_MV(16, 8, 6),  STG_MV(16, 8, 8),
    STG_MV(32, 16, 4), STG_MV(32, 16, 6), STG_MV(32, 16, 8),
    STG_MV(48, 24, 4), STG_MV(48, 24, 6), STG_MV(48, 24, 8),
};
constexpr int kNMV = sizeof(kMV) / sizeof(kMV[0]);

const MVariant* pick_mv(int C, int H, int M) {
  // smallest tile that fits every dimension: table order is (CP, HP) ascending, then NT
  for (int i = 0; i < kNMV; ++i)
    if (kMV[i].CP >= C && kMV[i].HP >= H && kMV[i].NT * 8 >= M) return &kMV[i];
  return nullptr;
}
const MVariant* find_mv(int CP, int HP, int NT) {
  for (int i = 0; i < kNMV; ++i)
    if (kMV[i].CP == CP && kMV[i].HP == HP && kMV[i].NT == NT) return &kMV[i];
  return nullptr;
}

bool g_mattr[64] = {};
void set_mattrs() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (g_mattr[dev]) return;
  for (int i = 0; i < kNMV; ++i) {
    cudaFuncSetAttribute(kMV[i].fwd_train, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemCapM);
    cudaFuncSetAttribute(kMV[i].fwd_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemCapM);
  }
  g_mattr[dev] = true;
}

int rows_of_fwd(const BlkArgs& a) {
  int rm = 0;
  for (int z = 0; z < a.nblk; ++z) {
    const BlkDev& k = a.b[z];
    const int per = (k.L + k.nchunk_f - 1) / k.nchunk_f;
    const int rows = ((per - 1) * k.stride + k.w) * a.N;
    rm = rows > rm ? rows : rm;
  }
  return rm;
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  if (!e) return dflt;
  const int v = atoi(e);
  return v >= 1 ? v : dflt;
}

}  // namespace

// Plans for the tensor-core path: fill p.mma_f / p.mma_b, p.CP/HP/NT, chunking, shared memory, grid.
// Return false when the shape is outside the instantiated tiles (caller keeps the SIMT plan).
bool plan_blocks_mma_fwd(BlkArgs& a, BlkPlan& p) {
  int Hmax = 0, Mmax = 0;
  for (int z = 0; z < a.nblk; ++z) {
    Hmax = a.b[z].H > Hmax ? a.b[z].H : Hmax;
    Mmax = a.b[z].w * a.N > Mmax ? a.b[z].w * a.N : Mmax;
  }
  const MVariant* v = pick_mv(a.C, Hmax, Mmax);
  if (!v || (a.C & 1) || ((uintptr_t)a.x & 7)) return false;      // 64-bit fragment loads need even C, 8-B aligned x
  const int nwarps = 8;
  int saved[2] = {a.b[0].nchunk_f, a.b[1].nchunk_f};
  // windows per chunk: one per warp unless overridden (STG_FWD_WP, tuning aid); shrink until the slab fits
  for (int wp = env_int("STG_FWD_WP", 8); wp >= 1; --wp) {
    int rows_max = 0, gx = 0;
    for (int z = 0; z < a.nblk; ++z) {
      BlkDev& k = a.b[z];
      k.nchunk_f = (k.L + wp - 1) / wp;
      const int per = (k.L + k.nchunk_f - 1) / k.nchunk_f;
      const int rows = ((per - 1) * k.stride + k.w) * a.N;
      rows_max = rows > rows_max ? rows : rows_max;
      gx = k.nchunk_f > gx ? k.nchunk_f : gx;
    }
    const MLay l = make_mlay(v->CP, v->HP, v->NT * 8, Mmax, a.C, rows_max, nwarps, a.training ? 0 : 1);
    const size_t sm = 16 + (size_t)l.total * 4;
    if (sm <= kSmemCapM) {
      p.mma_f = 1; p.CP = v->CP; p.HP = v->HP; p.NT = v->NT;
      p.smem_f = sm; p.grid_x_f = gx; p.threads_f = nwarps * 32;
      return true;
    }
  }
  a.b[0].nchunk_f = saved[0]; a.b[1].nchunk_f = saved[1];      // keep the SIMT plan intact
  return false;
}

int launch_block_forward_mma(const BlkArgs& a, const BlkPlan& p, cudaStream_t s) {
  set_mattrs();
  const MVariant* v = find_mv(p.CP, p.HP, p.NT);
  if (!v) return -2;
  const int rows_max = rows_of_fwd(a);
  dim3 grid(p.grid_x_f, a.B, a.nblk);
  ProfScope ps(kProfFwdMain, s);
  if (a.training) v->fwd_train<<<grid, p.threads_f, p.smem_f, s>>>(a, rows_max);
  else v->fwd_eval<<<grid, p.threads_f, p.smem_f, s>>>(a, rows_max);
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

}  // namespace stg
