// Recurrent layer of the sibling models (primitive T2): the time recurrence of one nn.LSTM / nn.GRU
// layer (one or two directions, zero initial state), forward and backward, as persistent sm_100a kernels.
//   models/HAGCN/Model.py:33-73    three bidirectional LSTMs, batch = patches (1..5), SEQUENCE = bs*N (thousands of steps)
//   models/GAT_LSTM/Model.py:129   LSTM 100 -> 30 -> 20 over 40 patches
//   models/STGNN/Model.py:72,99    GRU 64 -> 64 over bs*N sequences
//   models/STMSGCN/Model.py:52-60  GRU -> 8 over 160 patches
// The input projection x.W_ih^T + b (and its gradients) is a plain GEMM and stays with the caller; what is here is the
// part no GEMM library can batch: h_t = cell(xg_t + W_hh.h_{t-1}).
//
// Work decomposition.  A CTA owns 8 sequences of one direction for all T steps.  The per-step product
// gates[G*H, 8] = W_hh[G*H, H] . h[H, 8] runs on warp-level tensor-core tiles (mma.sync m16n8k8 TF32 with the
// 3-term error-compensated split, fp32-level accuracy): the A fragments -- W_hh, split into hi / lo -- are loaded ONCE and
// stay in registers for the whole sequence (128 registers per thread at H = 60), the 8 sequences are the N dimension,
// and the only per-step operand traffic is the hidden state: 2 conflict-free LDS.32 per k-tile.  (First version: one
// W_hh row per thread and FFMA2 against broadcast LDS.128 of h -- measured 1.72 us per step at H = 60 because a
// broadcast LDS.128 costs 2.45 cycles per warp on the SM and every FMA pair needed one, scripts/rnn_probe.cu.)
// The gate pre-activations meet in shared memory, where thread (unit, sequence) applies the cell and keeps c_t in a
// register; the next step's xg values are prefetched before the products.  H > 64 (HAGCN's 120-wide layer) does not fit
// one CTA's register file, so the hidden units are split over a cluster of 4 CTAs which exchange the new h through
// distributed shared memory (double-buffered, one cluster barrier per step).
// Backward: the same structure with the transposed matrix (output rows = units of this CTA, contraction over all
// gate rows, split over warps into partial sums), the activations the forward saved, dgates written in the layout
// of xg (it IS d loss / d xg), dW_hh left to the caller as a GEMM of the gate gradients against the shifted outputs.
#include <math.h>

#include "../../include/stgconv_b200.h"
#include "stg_common.cuh"
#include "stg_mma.cuh"

namespace stg {
namespace {

constexpr int NB = 8;     // sequences per CTA = N of the m16n8k8 tiles

struct RnnArgs {
  const float* xg;      // element (b, t, d, r) at b*gsb + t*gst + d*G*H + r
  long long gsb, gst;
  const float* whh;     // [ndir][G*H][H]
  const float* bhn;     // GRU: [ndir][H] (b_hn stays inside r * (...)), else null
  float* out;           // element (b, t, d, j) at b*osb + t*ost + d*H + j
  long long osb, ost;
  float* saved;         // [ndir][ntile][T][S][H*8] or null (inference)
  // backward only
  const float* dout;    // layout of out
  float* dxg;           // layout of xg: d loss / d xg
  float* dhn;           // GRU: layout of out: d loss / d (W_hn h + b_hn)
  int T, B, H, Hc, ntile;
};

// Accurate-to-3e-7 activations on the MUFU pipe, 4-5 instructions each (the libm tanhf / expf sequences were a third
// of a step): sigmoid = rcp(1 + ex2(-x log2 e)), tanh = 2 sigmoid(2x) - 1.  Saturate correctly (ex2 -> inf / 0).
STG_DEVINL float ex2_(float v) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
STG_DEVINL float rcp_(float v) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
STG_DEVINL float sigmoidf_(float v) { return rcp_(1.f + ex2_(-1.4426950408889634f * v)); }
STG_DEVINL float tanhf_(float v) { return fmaf(2.f, rcp_(1.f + ex2_(-2.8853900817779268f * v)), -1.f); }

// Per-step B operand split: hi = value truncated to tf32 (one LOP3), lo = the exact remainder, of which the tensor core
// reads the leading 11 bits (cvt.rna.tf32 is a 4-instruction emulation on sm_100a: 9 instructions per value with it).
STG_DEVINL FragB make_b_trunc(float b0, float b1) {
  FragB f;
  f.hi[0] = __float_as_uint(b0) & 0xffffe000u;
  f.hi[1] = __float_as_uint(b1) & 0xffffe000u;
  f.lo[0] = __float_as_uint(b0 - __uint_as_float(f.hi[0]));
  f.lo[1] = __float_as_uint(b1 - __uint_as_float(f.hi[1]));
  return f;
}

template <int CS>
STG_DEVINL unsigned cluster_rank() {
  if (CS == 1) return 0;
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
template <int CS>
STG_DEVINL void step_barrier() {
  if (CS == 1) {
    __syncthreads();
  } else {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
}
// store v at the same shared-memory offset in every CTA of the cluster
template <int CS>
STG_DEVINL void st_all(float* p, float v) {
  if (CS == 1) {
    *p = v;
  } else {
    const uint32_t a = smem_u32(p);
#pragma unroll
    for (int c = 0; c < CS; ++c) {
      uint32_t ra;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(c));
      asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
    }
  }
}

// Kernel shapes.  KT = k-tiles of 8 covering the hidden size, MT = m-tiles (16 gate rows) per warp in the forward,
// WARPS per CTA, CS = CTAs per cluster (hidden units split), ITEMS = (unit, sequence) cells per thread.
template <int KT_> struct RnnCfg;
template <> struct RnnCfg<1>  { static constexpr int KT = 1,  MT = 1, WARPS = 2, CS = 1, ITEMS = 1; };   // H <= 8
template <> struct RnnCfg<4>  { static constexpr int KT = 4,  MT = 1, WARPS = 8, CS = 1, ITEMS = 1; };   // H <= 32
template <> struct RnnCfg<8>  { static constexpr int KT = 8,  MT = 2, WARPS = 8, CS = 1, ITEMS = 2; };   // H <= 64
template <> struct RnnCfg<16> { static constexpr int KT = 16, MT = 1, WARPS = 8, CS = 4, ITEMS = 1; };   // H <= 128

// ------------------------------------------------------------------------------------------------ forward
// saved activations: [dir][tile][t][unit*8 + sequence][6]  (LSTM: i f g o c c_prev; GRU: r z n hn+b h_prev -)
constexpr int SP = 6;

// Row order of the forward tiles: an m-tile of 16 rows holds the FOUR gate rows of four hidden units, arranged so that a
// thread's two C-fragment rows (g8, g8 + 8) are gates (0, 2) of a unit on even g8 and gates (1, 3) of the same unit on odd
// g8.  Two lane-pair shuffles then give every thread all four gate pre-activations of one (unit, sequence) cell: the cell
// is applied straight from the MMA accumulators -- no shared-memory gate exchange, ONE barrier per step (for h).
// GRU: gate rows (r, z, W_hn h + b_hn, xg_n) -- the fourth row has zero weights and only carries its xg term.
template <int G, int KT>
__global__ void __launch_bounds__(RnnCfg<KT>::WARPS * 32, 1) k_rnn_fwd(const RnnArgs a) {
  using Cfg = RnnCfg<KT>;
  constexpr int CS = Cfg::CS, MT = Cfg::MT, NT = Cfg::WARPS * 32;
  constexpr int KP = KT * 8, KS = KP + 4;                 // h row stride: banks of (sequence, k) pairs distinct
  const int H = a.H, Hc = a.Hc, T = a.T;
  const int crank = (int)cluster_rank<CS>();
  const int tile = blockIdx.x / CS, d = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g8 = lane >> 2, t4 = lane & 3;
  const int b0 = tile * NB;
  const bool odd = (g8 & 1) != 0;

  extern __shared__ float4 sm4[];
  float* h_s = reinterpret_cast<float*>(sm4);             // [2][8][KS]

  FragA wa[MT][KT];
  float cinit[MT][2];                                     // constant C init of a row (GRU: b_hn on the hn row)
  bool useb[MT][2], xok[MT][4], live[MT];
  const float* xp[MT][4];                                 // running pointers into xg (element e: row half e>>1, sequence 2*t4 + (e&1))
  float* op[MT];
  float* svp[MT];
  int hoff[MT];
  float st[MT];                                           // LSTM: c ; GRU: h of this thread's cell
  const size_t HN = (size_t)H * NB;
  const int bc = 2 * t4 + (odd ? 1 : 0);                  // sequence of this thread's cell
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    const int ul = (warp * MT + i) * 4 + (g8 >> 1), j = crank * Hc + ul;
    const bool uok = ul < Hc && j < H;
    int gate[2];
    gate[0] = odd ? 1 : 0;
    gate[1] = odd ? 3 : 2;
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      useb[i][hf] = (G == 3 && gate[hf] == 2);
      cinit[i][hf] = (useb[i][hf] && uok) ? a.bhn[d * H + j] : 0.f;
    }
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int hf = e & 1, k = kt * 8 + t4 + 4 * (e >> 1);
        const bool wrow = uok && k < H && gate[hf] < G;     // GRU: the xg_n row has no recurrent weights
        v[e] = wrow ? a.whh[((size_t)(d * G + gate[hf]) * H + j) * H + k] : 0.f;
      }
      wa[i][kt] = make_a(v[0], v[1], v[2], v[3]);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int hf = e >> 1, b = b0 + 2 * t4 + (e & 1);
      const int xrow = (G == 3 && gate[hf] == 3) ? 2 : gate[hf];      // xg plane of this row
      xok[i][e] = uok && b < a.B && !useb[i][hf];
      xp[i][e] = a.xg + (xok[i][e] ? (size_t)b * a.gsb + (size_t)(d ? T - 1 : 0) * a.gst + (size_t)d * G * H + (size_t)xrow * H + j : 0);
    }
    st[i] = 0.f;
    live[i] = uok && b0 + bc < a.B;
    hoff[i] = bc * KS + (uok ? j : 0);
    op[i] = a.out + (live[i] ? (size_t)(b0 + bc) * a.osb + (size_t)(d ? T - 1 : 0) * a.ost + (size_t)d * H + j : 0);
    svp[i] = (a.saved && live[i])
                 ? a.saved + (((size_t)(d * a.ntile + tile) * T + (d ? T - 1 : 0)) * HN + (size_t)j * NB + bc) * SP
                 : nullptr;
  }
  const long long xstep = d ? -a.gst : a.gst;
  const long long ostep = d ? -a.ost : a.ost;
  const long long sstep = (long long)(d ? -1 : 1) * (long long)HN * SP;

  for (int e = tid; e < 2 * NB * KS; e += NT) h_s[e] = 0.f;
  step_barrier<CS>();

  float xa[MT][4], xb[MT][4];                             // xg of steps tt and tt + 1: loads stay in flight for two steps
  auto load_xg = [&](float (&dst)[MT][4]) {
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        dst[i][e] = xok[i][e] ? __ldg(xp[i][e]) : 0.f;
        xp[i][e] += xstep;
      }
  };
  load_xg(xa);
  if (T > 1) load_xg(xb);

  auto step = [&](const int tt, float (&xv)[MT][4]) {
    float acc[MT][4], acl[MT][4], acm[MT][4];             // three independent MMA chains per m-tile
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc[i][e] = useb[i][e >> 1] ? cinit[i][e >> 1] : xv[i][e];
        acl[i][e] = 0.f;
        acm[i][e] = 0.f;
      }
    const float* hc = h_s + (tt & 1) * NB * KS + g8 * KS + t4;
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
      const FragB hb = make_b_trunc(hc[kt * 8], hc[kt * 8 + 4]);
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        mma_tf32(acl[i], wa[i][kt].lo, hb.hi);
        mma_tf32(acm[i], wa[i][kt].hi, hb.lo);
        mma_tf32(acc[i], wa[i][kt].hi, hb.hi);
      }
    }
    if (tt + 2 < T) load_xg(xv);                          // this register set is free again: fetch step tt + 2
    float* hnx = h_s + ((tt + 1) & 1) * NB * KS;
#pragma unroll
    for (int i = 0; i < MT; ++i) {
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = acc[i][e] + (acl[i][e] + acm[i][e]);
      // lane pair (g8 even, g8 odd) = lanes l, l ^ 4: the even lane takes sequence 2 t4, the odd lane 2 t4 + 1
      const float r0 = __shfl_xor_sync(0xffffffffu, odd ? v[0] : v[1], 4);
      const float r1 = __shfl_xor_sync(0xffffffffu, odd ? v[2] : v[3], 4);
      const float p0 = odd ? r0 : v[0], p1 = odd ? v[1] : r0, p2 = odd ? r1 : v[2], p3 = odd ? v[3] : r1;
      float hnew;
      if (G == 4) {
        const float ig = sigmoidf_(p0), fg = sigmoidf_(p1), gg = tanhf_(p2), og = sigmoidf_(p3);
        const float cp = st[i], c = fmaf(fg, cp, ig * gg);
        st[i] = c;
        hnew = og * tanhf_(c);
        if (svp[i]) {
          float2* s2 = reinterpret_cast<float2*>(svp[i]);
          s2[0] = make_float2(ig, fg); s2[1] = make_float2(gg, og); s2[2] = make_float2(c, cp);
        }
      } else {
        const float rgt = sigmoidf_(p0), zg = sigmoidf_(p1), ng = tanhf_(fmaf(rgt, p2, p3));
        const float hp = st[i];
        hnew = fmaf(zg, hp - ng, ng);                      // (1 - z) n + z h
        st[i] = hnew;
        if (svp[i]) {
          float2* s2 = reinterpret_cast<float2*>(svp[i]);
          s2[0] = make_float2(rgt, zg); s2[1] = make_float2(ng, p2); s2[2] = make_float2(hp, 0.f);
        }
      }
      if (live[i]) {
        *op[i] = hnew;
        st_all<CS>(hnx + hoff[i], hnew);
      }
      op[i] += ostep;
      if (svp[i]) svp[i] += sstep;
    }
    step_barrier<CS>();
  };
  for (int tt = 0; tt < T; tt += 2) {
    step(tt, xa);
    if (tt + 1 < T) step(tt + 1, xb);
  }
}

// ------------------------------------------------------------------------------------------------ backward
// dh_{t-1}[k][b] = sum_{g,j} W_hh[g*H + j][k] * dg[g][j][b]: output rows = units k of this CTA (MTB m-tiles), contraction
// index kk = g*KP + j split over NPART warps per m-tile, KTP k-tiles each.
template <int G, int KT>
struct RnnBwdCfg {
  using Cfg = RnnCfg<KT>;
  static constexpr int KP = KT * 8;
  static constexpr int HCMAX = (KT == 16) ? 32 : KP;                    // units per CTA
  static constexpr int MTB = (HCMAX + 15) / 16;                          // 1, 2, 4, 2
  static constexpr int NPART = Cfg::WARPS / MTB;                         // 2, 4, 2, 4
  static constexpr int KTP = (G * KT + NPART - 1) / NPART;               // k-tiles per warp
  static constexpr int KK = NPART * KTP * 8;                             // padded contraction length
  static constexpr int KS = ((4 * KP > KK ? 4 * KP : KK) + 4);           // row stride of dg_s (4 planes kept)
};

template <int G, int KT>
__global__ void __launch_bounds__(RnnCfg<KT>::WARPS * 32, 1) k_rnn_bwd(const RnnArgs a) {
  using Cfg = RnnCfg<KT>;
  using BC = RnnBwdCfg<G, KT>;
  constexpr int CS = Cfg::CS, ITEMS = Cfg::ITEMS, WARPS = Cfg::WARPS, NT = WARPS * 32;
  constexpr int KP = BC::KP, KS = BC::KS, KTP = BC::KTP, NPART = BC::NPART, MTB = BC::MTB;
  constexpr int PSTRIDE = MTB * 16 * NB;
  const int H = a.H, Hc = a.Hc, T = a.T;
  const int crank = (int)cluster_rank<CS>();
  const int tile = blockIdx.x / CS, d = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g8 = lane >> 2, t4 = lane & 3;
  const int mt = warp % MTB, part = warp / MTB;
  const int b0 = tile * NB;

  extern __shared__ float4 sm4[];
  float* dg_s = reinterpret_cast<float*>(sm4);            // [2][8][KS]: gate gradients of EVERY unit, k index g*KP + j
  float* p_s = dg_s + 2 * NB * KS;                        // [NPART][MTB*16*8] partial products

  FragA wa[KTP];                                          // A[m = unit][kk] = W_hh[g*H + j][unit]
#pragma unroll
  for (int kt = 0; kt < KTP; ++kt) {
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int m = mt * 16 + g8 + 8 * (e & 1), ku = crank * Hc + m;
      const int kk = (part * KTP + kt) * 8 + t4 + 4 * (e >> 1);
      const int g = kk / KP, j = kk - g * KP;
      v[e] = (m < Hc && ku < H && g < G && j < H) ? a.whh[((size_t)(d * G + g) * H + j) * H + ku] : 0.f;
    }
    wa[kt] = make_a(v[0], v[1], v[2], v[3]);
  }
  for (int e = tid; e < 2 * NB * KS; e += NT) dg_s[e] = 0.f;
  step_barrier<CS>();

  const int tl = d ? 0 : T - 1;                           // first step of the backward sweep
  const size_t HN = (size_t)H * NB;
  const size_t GH = (size_t)G * H;

  // ---- cell role
  const int Bt = min(NB, a.B - b0);                       // live sequences of this tile
  float dhrec[ITEMS], dcs[ITEMS];
  const float* svp[ITEMS];
  const float* dop[ITEMS];
  const float* pp[ITEMS];
  int goff[ITEMS];
  bool item[ITEMS], live[ITEMS];
#pragma unroll
  for (int q = 0; q < ITEMS; ++q) {
    const int idx = tid + q * NT, jl = idx / Bt, b = idx - jl * Bt, j = crank * Hc + jl;   // live sequences only
    dhrec[q] = 0.f, dcs[q] = 0.f;
    item[q] = idx < Hc * Bt;
    live[q] = item[q] && j < H;
    goff[q] = b * KS + j;
    pp[q] = p_s + jl * NB + b;
    svp[q] = a.saved + (live[q] ? (((size_t)(d * a.ntile + tile) * T + tl) * HN + (size_t)j * NB + b) * SP : 0);
    dop[q] = a.dout + (live[q] ? (size_t)(b0 + b) * a.osb + (size_t)tl * a.ost + (size_t)d * H + j : 0);
  }
  const long long ostep = d ? a.ost : -a.ost;             // the sweep runs against the forward direction
  const long long sstep = (long long)(d ? 1 : -1) * (long long)HN * SP;

  // ---- copy role: warp w moves the (gate plane, sequence) rows w*RPW .. of dg_s to global memory, lanes over units
  constexpr int RPW = 32 / WARPS;                         // 32 = 4 planes x 8 sequences
  constexpr int JL = (KT == 16) ? 32 : KP;                // units per CTA, padded
  float* cp_g[RPW];
  int cp_s[RPW];
  long long cp_step[RPW];
  bool cp_ok[RPW];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int row = warp * RPW + r, g = row & 3, b = row >> 2;
    cp_ok[r] = b0 + b < a.B;
    cp_s[r] = b * KS + g * KP + crank * Hc;
    const size_t xrow = (size_t)(b0 + b) * a.gsb + (size_t)tl * a.gst + (size_t)d * GH + (size_t)crank * Hc;
    const size_t orow = (size_t)(b0 + b) * a.osb + (size_t)tl * a.ost + (size_t)d * H + (size_t)crank * Hc;
    if (!cp_ok[r]) {
      cp_g[r] = a.dxg; cp_step[r] = 0;
    } else if (G == 4) {
      cp_g[r] = a.dxg + xrow + (size_t)g * H; cp_step[r] = d ? a.gst : -a.gst;
    } else if (g == 2) {                                  // GRU: d / d (W_hn h + b_hn) has its own output
      cp_g[r] = a.dhn + orow; cp_step[r] = ostep;
    } else {
      cp_g[r] = a.dxg + xrow + (size_t)(g == 3 ? 2 : g) * H; cp_step[r] = d ? a.gst : -a.gst;
    }
  }

  // prefetch registers of the cell role: saved activations + upstream gradient, two steps deep
  float2 sa[ITEMS][3], sb[ITEMS][3];
  float ua[ITEMS], ub[ITEMS];
  auto fetch = [&](float2 (&sp)[ITEMS][3], float (&du)[ITEMS]) {
#pragma unroll
    for (int q = 0; q < ITEMS; ++q) {
      if (live[q]) {
        const float2* s2 = reinterpret_cast<const float2*>(svp[q]);
        sp[q][0] = __ldg(s2); sp[q][1] = __ldg(s2 + 1); sp[q][2] = __ldg(s2 + 2);
        du[q] = __ldg(dop[q]);
      } else {
        sp[q][0] = sp[q][1] = sp[q][2] = make_float2(0.f, 0.f);
        du[q] = 0.f;
      }
      svp[q] += sstep;
      dop[q] += ostep;
    }
  };
  fetch(sa, ua);
  if (T > 1) fetch(sb, ub);

  auto step = [&](const int n, float2 (&sp)[ITEMS][3], float (&du)[ITEMS]) {   // n-th step of the sweep
    float* dgc = dg_s + (n & 1) * NB * KS;
    // ---- cell role: gradient of the gate pre-activations
#pragma unroll
    for (int q = 0; q < ITEMS; ++q) {
      if (item[q]) {
        const float dh = du[q] + dhrec[q];
        float d0, d1, d2, d3;
        if (G == 4) {
          const float ig = sp[q][0].x, fg = sp[q][0].y, gg = sp[q][1].x, og = sp[q][1].y, c = sp[q][2].x, cp = sp[q][2].y;
          const float tc = tanhf_(c);
          const float dc = fmaf(dh * og, 1.f - tc * tc, dcs[q]);
          dcs[q] = dc * fg;
          d0 = dc * gg * ig * (1.f - ig);
          d1 = dc * cp * fg * (1.f - fg);
          d2 = dc * ig * (1.f - gg * gg);
          d3 = dh * tc * og * (1.f - og);
          dhrec[q] = 0.f;
        } else {
          const float rgt = sp[q][0].x, zg = sp[q][0].y, ng = sp[q][1].x, hnb = sp[q][1].y, hp = sp[q][2].x;
          const float dnp = dh * (1.f - zg) * (1.f - ng * ng);
          d0 = dnp * hnb * rgt * (1.f - rgt);
          d1 = dh * (hp - ng) * zg * (1.f - zg);
          d2 = dnp * rgt;                                  // d / d (W_hn h + b_hn)
          d3 = dnp;                                        // d / d (xg n-plane); not part of the contraction
          dhrec[q] = dh * zg;                              // direct path h_{t-1} -> h_t
        }
        float* p = dgc + goff[q];
        st_all<CS>(p, d0);
        st_all<CS>(p + KP, d1);
        st_all<CS>(p + 2 * KP, d2);
        if (G == 4) st_all<CS>(p + 3 * KP, d3); else p[3 * KP] = d3;
      }
    }
    if (n + 2 < T) fetch(sp, du);                          // registers free again: loads of step n + 2 fly for two steps
    step_barrier<CS>();
    // ---- product role
    {
      float acc[4] = {0.f, 0.f, 0.f, 0.f}, acl[4] = {0.f, 0.f, 0.f, 0.f}, acm[4] = {0.f, 0.f, 0.f, 0.f};
      const float* dv = dgc + g8 * KS + part * KTP * 8 + t4;
#pragma unroll
      for (int kt = 0; kt < KTP; ++kt) {
        const FragB gb = make_b_trunc(dv[kt * 8], dv[kt * 8 + 4]);
        mma_tf32(acl, wa[kt].lo, gb.hi);
        mma_tf32(acm, wa[kt].hi, gb.lo);
        mma_tf32(acc, wa[kt].hi, gb.hi);
      }
      float* pw = p_s + part * PSTRIDE + (mt * 16 + g8) * NB + 2 * t4;
      *reinterpret_cast<float2*>(pw) = make_float2(acc[0] + (acl[0] + acm[0]), acc[1] + (acl[1] + acm[1]));
      *reinterpret_cast<float2*>(pw + 8 * NB) = make_float2(acc[2] + (acl[2] + acm[2]), acc[3] + (acl[3] + acm[3]));
    }
    // ---- copy role: gate gradients of this CTA's units to global memory, coalesced over the unit index
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
#pragma unroll
      for (int jj = 0; jj < JL; jj += 32) {
        const int jl = jj + lane;
        if (cp_ok[r] && jl < Hc && crank * Hc + jl < H) cp_g[r][jl] = dgc[cp_s[r] + jl];
      }
      cp_g[r] += cp_step[r];
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < ITEMS; ++q) {
      if (item[q]) {
        float s = dhrec[q];
#pragma unroll
        for (int pt = 0; pt < NPART; ++pt) s += pp[q][pt * PSTRIDE];
        dhrec[q] = s;
      }
    }
    // p_s is rewritten only after the next step's barrier, dg_s is double-buffered
  };
  for (int n = 0; n < T; n += 2) {
    step(n, sa, ua);
    if (n + 1 < T) step(n + 1, sb, ub);
  }
}

template <int G, int KT>
int launch_rnn(bool backward, const RnnArgs& a, int ndir, cudaStream_t s) {
  using Cfg = RnnCfg<KT>;
  using BC = RnnBwdCfg<G, KT>;
  constexpr int CS = Cfg::CS;
  if (a.Hc > BC::HCMAX) return set_err(STG_ERR_UNSUPPORTED, "rnn: %d units per CTA exceed %d", a.Hc, BC::HCMAX);
  const size_t smem = backward ? sizeof(float) * ((size_t)2 * NB * BC::KS + (size_t)BC::NPART * BC::MTB * 16 * NB)
                               : sizeof(float) * ((size_t)2 * NB * (KT * 8 + 4));
  auto kern = backward ? k_rnn_bwd<G, KT> : k_rnn_fwd<G, KT>;
  if (smem > 48 * 1024) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return check_cuda("rnn smem attribute");
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(a.ntile * CS), (unsigned)ndir, 1);
  cfg.blockDim = dim3((unsigned)(Cfg::WARPS * 32), 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at = {};
  at.id = cudaLaunchAttributeClusterDimension;
  at.val.clusterDim.x = CS;
  at.val.clusterDim.y = 1;
  at.val.clusterDim.z = 1;
  cfg.attrs = &at;
  cfg.numAttrs = CS > 1 ? 1 : 0;
  if (cudaLaunchKernelEx(&cfg, kern, a) != cudaSuccess) return check_cuda("rnn launch");
  return check_cuda(backward ? "k_rnn_bwd" : "k_rnn_fwd");
}

template <int G>
int dispatch_kt(bool backward, const RnnArgs& a, int ndir, int KT, cudaStream_t s) {
  switch (KT) {
    case 1: return launch_rnn<G, 1>(backward, a, ndir, s);
    case 4: return launch_rnn<G, 4>(backward, a, ndir, s);
    case 8: return launch_rnn<G, 8>(backward, a, ndir, s);
    default: return launch_rnn<G, 16>(backward, a, ndir, s);
  }
}

int rnn_kt(int H) { return H <= 8 ? 1 : H <= 32 ? 4 : H <= 64 ? 8 : 16; }

int rnn_common(int cell, int T, int B, int H, int ndir, RnnArgs* a, int* G, int* KT) {
  if (cell != STG_RNN_LSTM && cell != STG_RNN_GRU) return set_err(STG_ERR_INVALID, "rnn: unknown cell %d", cell);
  if (T < 1 || B < 1 || H < 1) return set_err(STG_ERR_INVALID, "rnn: non-positive dimension");
  if (ndir != 1 && ndir != 2) return set_err(STG_ERR_INVALID, "rnn: ndir must be 1 or 2");
  if (H > 128) return set_err(STG_ERR_UNSUPPORTED, "rnn: hidden size %d > 128", H);
  *G = cell == STG_RNN_LSTM ? 4 : 3;
  *KT = rnn_kt(H);
  const int cs = *KT == 16 ? 4 : 1;
  a->T = T; a->B = B; a->H = H;
  a->Hc = (H + cs - 1) / cs;
  a->ntile = (B + NB - 1) / NB;
  return STG_OK;
}

}  // namespace
}  // namespace stg

using namespace stg;

extern "C" int stg_rnn_batch_tile(int B) { (void)B; return NB; }

extern "C" size_t stg_rnn_saved_floats(int cell, int T, int B, int H, int ndir) {
  if (T < 1 || B < 1 || H < 1) return 0;
  const size_t ntile = (size_t)(B + NB - 1) / NB;
  (void)cell;
  return (size_t)ndir * ntile * T * SP * H * NB;
}

extern "C" int stg_rnn_forward(int cell, const float* xg_dev, int64_t xg_bstride, int64_t xg_tstride,
                               const float* whh_dev, const float* bhn_dev, int T, int B, int H, int ndir,
                               float* out_dev, int64_t out_bstride, int64_t out_tstride, float* saved_dev,
                               void* stream) {
  RnnArgs a = {};
  int G, KT;
  if (int rc = rnn_common(cell, T, B, H, ndir, &a, &G, &KT)) return rc;
  if (!xg_dev || !whh_dev || !out_dev) return set_err(STG_ERR_INVALID, "rnn: null pointer");
  if (cell == STG_RNN_GRU && !bhn_dev) return set_err(STG_ERR_INVALID, "rnn: GRU needs b_hn");
  a.xg = xg_dev; a.gsb = xg_bstride; a.gst = xg_tstride;
  a.whh = whh_dev; a.bhn = bhn_dev;
  a.out = out_dev; a.osb = out_bstride; a.ost = out_tstride;
  a.saved = saved_dev;
  cudaStream_t s = (cudaStream_t)stream;
  return G == 4 ? dispatch_kt<4>(false, a, ndir, KT, s) : dispatch_kt<3>(false, a, ndir, KT, s);
}

extern "C" int stg_rnn_backward(int cell, const float* whh_dev, const float* saved_dev, const float* dout_dev,
                                int64_t out_bstride, int64_t out_tstride, int T, int B, int H, int ndir,
                                float* dxg_dev, int64_t xg_bstride, int64_t xg_tstride, float* dhn_dev, void* stream) {
  RnnArgs a = {};
  int G, KT;
  if (int rc = rnn_common(cell, T, B, H, ndir, &a, &G, &KT)) return rc;
  if (!whh_dev || !saved_dev || !dout_dev || !dxg_dev) return set_err(STG_ERR_INVALID, "rnn: null pointer");
  if (cell == STG_RNN_GRU && !dhn_dev) return set_err(STG_ERR_INVALID, "rnn: GRU backward needs dhn");
  a.whh = whh_dev; a.saved = const_cast<float*>(saved_dev);
  a.dout = dout_dev; a.osb = out_bstride; a.ost = out_tstride;
  a.dxg = dxg_dev; a.gsb = xg_bstride; a.gst = xg_tstride; a.dhn = dhn_dev;
  cudaStream_t s = (cudaStream_t)stream;
  return G == 4 ? dispatch_kt<4>(true, a, ndir, KT, s) : dispatch_kt<3>(true, a, ndir, KT, s);
}
