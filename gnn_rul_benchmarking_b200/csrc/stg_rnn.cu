// Recurrent layer of the sibling models (primitive T2): the time recurrence of one nn.LSTM / nn.GRU
// layer (one or two directions, zero initial state), forward and backward, as persistent sm_100a kernels.
//   models/HAGCN/Model.py:33-73    three bidirectional LSTMs, batch = patches (1..5), SEQUENCE = bs*N (thousands of steps)
//   models/GAT_LSTM/Model.py:129   LSTM 100 -> 30 -> 20 over 40 patches
//   models/STGNN/Model.py:72,99    GRU 64 -> 64 over bs*N sequences
//   models/STMSGCN/Model.py:52-60  GRU -> 8 over 160 patches
// The input projection x.W_ih^T + b (and its gradients) is a plain GEMM and stays with the caller; what is here is the
// part no GEMM library can batch: h_t = cell(xg_t + W_hh.h_{t-1}).
//
// Work decomposition.  A CTA owns NB sequences of one direction for all T steps.  Thread r = (gate g, unit j) keeps its
// row of W_hh in REGISTERS for the whole sequence (KP values), the hidden state lives in shared memory as h[b][k] so that
// one 16-byte broadcast load feeds two packed FFMA2 (pairs over k), the next step's xg row is prefetched before the
// current step's products, and the gate pre-activations meet in shared memory where thread (unit, sequence) applies the
// cell and keeps c_t in a register.  H > 64 (HAGCN's 120-wide layer) does not fit the register file with one row per
// thread at 4H threads, so the hidden units are split over a cluster of 2 CTAs which exchange the new h through
// distributed shared memory (double-buffered, one cluster barrier per step).
// Backward: the same structure with the transposed matrix (thread (g, k) keeps column k of gate g), the activations
// the forward saved, dgates written in the layout of xg (it IS d loss / d xg), dW_hh left to the caller as a GEMM of
// dgates against the shifted outputs.
#include <math.h>

#include "../../include/stgconv_b200.h"
#include "stg_common.cuh"

namespace stg {
namespace {

struct RnnArgs {
  const float* xg;      // element (b, t, d, r) at b*gsb + t*gst + d*G*H + r
  long long gsb, gst;
  const float* whh;     // [ndir][G*H][H]
  const float* bhn;     // GRU: [ndir][H] (b_hn stays inside r * (...)), else null
  float* out;           // element (b, t, d, j) at b*osb + t*ost + d*H + j
  long long osb, ost;
  float* saved;         // [ndir][ntile][T][S][H*NB] or null (inference)
  // backward only
  const float* dout;    // layout of out
  float* dxg;           // layout of xg: LSTM d/d(xg); GRU planes (r, z, hn)
  float* dnx;           // GRU: element (b, t, d, j) at (b*gsb + t*gst)/G + d*H + j : d/d(xg n-plane)
  int T, B, H, Hc, ntile;
};

STG_DEVINL float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

STG_DEVINL void ffma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
STG_DEVINL unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
STG_DEVINL float sum2(unsigned long long v) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
  return lo + hi;
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

// acc[b] += sum_k w[k] * v[b][k]   (v: [NB][KP] floats in shared memory, 16-byte aligned rows)
template <int KP, int NB>
STG_DEVINL void matvec(const unsigned long long (&w2)[KP / 2], const float* v, unsigned long long (&acc)[NB]) {
#pragma unroll
  for (int k4 = 0; k4 < KP / 4; ++k4) {
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const ulonglong2 hv = *reinterpret_cast<const ulonglong2*>(v + b * KP + k4 * 4);
      ffma2(acc[b], w2[2 * k4], hv.x);
      ffma2(acc[b], w2[2 * k4 + 1], hv.y);
    }
  }
}

__host__ __device__ constexpr int rnn_min_blocks(int KP) { return KP <= 8 ? 4 : (KP <= 32 ? 3 : (KP <= 64 ? 2 : 1)); }
__host__ __device__ constexpr int rnn_cluster(int KP) { return KP > 64 ? 2 : 1; }
__host__ __device__ constexpr int rnn_saved_planes(int G) { return G == 4 ? 6 : 5; }

// ------------------------------------------------------------------------------------------------ forward
template <int G, int KP, int NB>
__global__ void __launch_bounds__(256, rnn_min_blocks(KP)) k_rnn_fwd(const RnnArgs a) {
  constexpr int CS = rnn_cluster(KP);
  constexpr int S = rnn_saved_planes(G);
  constexpr int ITEMS = (NB + G - 1) / G;
  const int H = a.H, Hc = a.Hc, T = a.T;
  const int crank = (int)cluster_rank<CS>();
  const int tile = blockIdx.x / CS, d = blockIdx.y;
  const int P = G * Hc, tid = threadIdx.x;
  const int g = tid / Hc, jl = tid - g * Hc, j = crank * Hc + jl;
  const bool row_ok = tid < P && j < H;
  const int b0 = tile * NB;

  extern __shared__ float4 sm4[];
  float* h_s = reinterpret_cast<float*>(sm4);             // [2][NB][KP]
  float* g_s = h_s + 2 * NB * KP;                         // [GP][Hc*NB]

  unsigned long long w2[KP / 2];
  {
    const float* wr = a.whh + ((size_t)(d * G + g) * H + j) * H;
#pragma unroll
    for (int k = 0; k < KP; k += 2)
      w2[k / 2] = pack2((row_ok && k < H) ? wr[k] : 0.f, (row_ok && k + 1 < H) ? wr[k + 1] : 0.f);
  }
  const float bhn = (G == 3 && row_ok && g == 2) ? a.bhn[d * H + j] : 0.f;
  for (int e = tid; e < 2 * NB * KP; e += blockDim.x) h_s[e] = 0.f;
  step_barrier<CS>();

  float st[ITEMS];                                        // LSTM: c ; GRU: h
#pragma unroll
  for (int q = 0; q < ITEMS; ++q) st[q] = 0.f;

  const float* xrow = a.xg + (size_t)d * G * H + (size_t)g * H + j;
  float xv[NB];
  {
    const int t = d ? T - 1 : 0;
#pragma unroll
    for (int b = 0; b < NB; ++b) xv[b] = (row_ok && b0 + b < a.B) ? xrow[(size_t)(b0 + b) * a.gsb + (size_t)t * a.gst] : 0.f;
  }
  const size_t HN = (size_t)H * NB;
  float* sv = a.saved ? a.saved + ((size_t)(d * a.ntile + tile) * T) * S * HN : nullptr;

  for (int tt = 0; tt < T; ++tt) {
    const int t = d ? T - 1 - tt : tt;
    float xn[NB];
    if (tt + 1 < T) {
      const int tn = d ? t - 1 : t + 1;
#pragma unroll
      for (int b = 0; b < NB; ++b)
        xn[b] = (row_ok && b0 + b < a.B) ? xrow[(size_t)(b0 + b) * a.gsb + (size_t)tn * a.gst] : 0.f;
    }
    unsigned long long acc[NB];
    const bool hn_plane = (G == 3 && g == 2);
#pragma unroll
    for (int b = 0; b < NB; ++b) acc[b] = pack2(hn_plane ? bhn : xv[b], 0.f);
    const float* hc = h_s + (tt & 1) * NB * KP;
    matvec<KP, NB>(w2, hc, acc);
    if (tid < P) {
#pragma unroll
      for (int b = 0; b < NB; ++b) g_s[g * Hc * NB + jl * NB + b] = sum2(acc[b]);
      if (hn_plane) {
#pragma unroll
        for (int b = 0; b < NB; ++b) g_s[3 * Hc * NB + jl * NB + b] = xv[b];
      }
    }
    __syncthreads();
    float* hnx = h_s + ((tt + 1) & 1) * NB * KP;
#pragma unroll
    for (int q = 0; q < ITEMS; ++q) {
      const int idx = tid + q * P;
      if (tid < P && idx < Hc * NB) {
        const int jl2 = idx / NB, b = idx - jl2 * NB, j2 = crank * Hc + jl2;
        const float p0 = g_s[idx], p1 = g_s[Hc * NB + idx], p2 = g_s[2 * Hc * NB + idx], p3 = g_s[3 * Hc * NB + idx];
        const bool live = j2 < H && b0 + b < a.B;
        float hnew;
        float* svp = sv ? sv + (size_t)t * S * HN + (size_t)j2 * NB + b : nullptr;
        if (G == 4) {
          const float ig = sigmoidf_(p0), fg = sigmoidf_(p1), gg = tanhf(p2), og = sigmoidf_(p3);
          const float cp = st[q], c = fmaf(fg, cp, ig * gg);
          st[q] = c;
          hnew = og * tanhf(c);
          if (svp && j2 < H) {
            svp[0] = ig; svp[HN] = fg; svp[2 * HN] = gg; svp[3 * HN] = og; svp[4 * HN] = c; svp[5 * HN] = cp;
          }
        } else {
          const float rg = sigmoidf_(p0), zg = sigmoidf_(p1), ng = tanhf(fmaf(rg, p2, p3));
          const float hp = st[q];
          hnew = fmaf(zg, hp - ng, ng);                    // (1 - z) n + z h
          st[q] = hnew;
          if (svp && j2 < H) {
            svp[0] = rg; svp[HN] = zg; svp[2 * HN] = ng; svp[3 * HN] = p2; svp[4 * HN] = hp;
          }
        }
        if (live) a.out[(size_t)(b0 + b) * a.osb + (size_t)t * a.ost + (size_t)d * H + j2] = hnew;
        if (j2 < KP) st_all<CS>(hnx + b * KP + j2, hnew);
      }
    }
    step_barrier<CS>();
#pragma unroll
    for (int b = 0; b < NB; ++b) xv[b] = xn[b];
  }
}

// ------------------------------------------------------------------------------------------------ backward
template <int G, int KP, int NB>
__global__ void __launch_bounds__(256, rnn_min_blocks(KP)) k_rnn_bwd(const RnnArgs a) {
  constexpr int CS = rnn_cluster(KP);
  constexpr int S = rnn_saved_planes(G);
  constexpr int ITEMS = (NB + G - 1) / G;
  constexpr int GP = 4;                                   // gradient planes in shared memory (GRU: r, z, hn, n)
  const int H = a.H, Hc = a.Hc, T = a.T;
  const int crank = (int)cluster_rank<CS>();
  const int tile = blockIdx.x / CS, d = blockIdx.y;
  const int P = G * Hc, tid = threadIdx.x;
  const int g = tid / Hc, kl = tid - g * Hc, k = crank * Hc + kl;
  const bool col_ok = tid < P && k < H;
  const int b0 = tile * NB;

  extern __shared__ float4 sm4[];
  float* dg_s = reinterpret_cast<float*>(sm4);            // [2][GP][NB][KP]   (every unit of the layer)
  float* p_s = dg_s + 2 * GP * NB * KP;                   // [G][Hc*NB]       partial W_g^T dg_g of this CTA's units

  unsigned long long w2[KP / 2];                          // column k of gate g: W_hh[g*H + jj][k]
  {
    const float* wc = a.whh + (size_t)(d * G + g) * H * H + k;
#pragma unroll
    for (int jj = 0; jj < KP; jj += 2)
      w2[jj / 2] = pack2((col_ok && jj < H) ? wc[(size_t)jj * H] : 0.f, (col_ok && jj + 1 < H) ? wc[(size_t)(jj + 1) * H] : 0.f);
  }
  for (int e = tid; e < 2 * GP * NB * KP; e += blockDim.x) dg_s[e] = 0.f;
  step_barrier<CS>();

  float dhrec[ITEMS], dcs[ITEMS];
#pragma unroll
  for (int q = 0; q < ITEMS; ++q) dhrec[q] = 0.f, dcs[q] = 0.f;

  const size_t HN = (size_t)H * NB;
  const float* sv = a.saved + ((size_t)(d * a.ntile + tile) * T) * S * HN;
  const size_t GH = (size_t)G * H;

  // prefetch registers of the pointwise role: saved planes + upstream gradient of the step about to be processed
  float sp[ITEMS][S], du[ITEMS];
  auto fetch = [&](int t) {
#pragma unroll
    for (int q = 0; q < ITEMS; ++q) {
      const int idx = tid + q * P;
      const int jl2 = idx / NB, b = idx - jl2 * NB, j2 = crank * Hc + jl2;
      const bool ok = tid < P && idx < Hc * NB && j2 < H;
#pragma unroll
      for (int s = 0; s < S; ++s) sp[q][s] = ok ? sv[((size_t)t * S + s) * HN + (size_t)j2 * NB + b] : 0.f;
      du[q] = (ok && b0 + b < a.B) ? a.dout[(size_t)(b0 + b) * a.osb + (size_t)t * a.ost + (size_t)d * H + j2] : 0.f;
    }
  };
  fetch(d ? 0 : T - 1);

  for (int tt = T - 1; tt >= 0; --tt) {
    const int t = d ? T - 1 - tt : tt;
    float* dgc = dg_s + (tt & 1) * GP * NB * KP;
    // ---- pointwise role: gradient of the gate pre-activations of step t
#pragma unroll
    for (int q = 0; q < ITEMS; ++q) {
      const int idx = tid + q * P;
      if (tid < P && idx < Hc * NB) {
        const int jl2 = idx / NB, b = idx - jl2 * NB, j2 = crank * Hc + jl2;
        const float dh = du[q] + dhrec[q];
        float d0, d1, d2, d3;
        if (G == 4) {
          const float ig = sp[q][0], fg = sp[q][1], gg = sp[q][2], og = sp[q][3], c = sp[q][4], cp = sp[q][5 % S];
          const float tc = tanhf(c);
          const float dc = fmaf(dh * og, 1.f - tc * tc, dcs[q]);
          dcs[q] = dc * fg;
          d0 = dc * gg * ig * (1.f - ig);
          d1 = dc * cp * fg * (1.f - fg);
          d2 = dc * ig * (1.f - gg * gg);
          d3 = dh * tc * og * (1.f - og);
          dhrec[q] = 0.f;
        } else {
          const float rg = sp[q][0], zg = sp[q][1], ng = sp[q][2], hnb = sp[q][3], hp = sp[q][4];
          const float dnp = dh * (1.f - zg) * (1.f - ng * ng);
          d0 = dnp * hnb * rg * (1.f - rg);
          d1 = dh * (hp - ng) * zg * (1.f - zg);
          d2 = dnp * rg;                                   // d / d (W_hn h + b_hn)
          d3 = dnp;                                        // d / d (xg n-plane)
          dhrec[q] = dh * zg;                              // direct path h_{t-1} -> h_t
        }
        if (j2 < KP) {
          st_all<CS>(dgc + (0 * NB + b) * KP + j2, d0);
          st_all<CS>(dgc + (1 * NB + b) * KP + j2, d1);
          st_all<CS>(dgc + (2 * NB + b) * KP + j2, d2);
          st_all<CS>(dgc + (3 * NB + b) * KP + j2, d3);
        }
      }
    }
    if (tt > 0) fetch(d ? T - tt : tt - 1);                // loads of the next step fly during the products
    step_barrier<CS>();
    // ---- product role: partial_g[k][b] = sum_jj W_hh[g*H + jj][k] * dg_g[jj][b]; dgates to global memory
    {
      unsigned long long acc[NB];
#pragma unroll
      for (int b = 0; b < NB; ++b) acc[b] = 0ull;
      const float* dv = dgc + (size_t)(tid < P ? g : 0) * NB * KP;
      matvec<KP, NB>(w2, dv, acc);
      if (tid < P) {
#pragma unroll
        for (int b = 0; b < NB; ++b) p_s[g * Hc * NB + kl * NB + b] = sum2(acc[b]);
      }
      if (col_ok) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          if (b0 + b < a.B) {
            const size_t row = (size_t)(b0 + b) * a.gsb + (size_t)t * a.gst;
            a.dxg[row + (size_t)d * GH + (size_t)g * H + k] = dv[b * KP + k];
            if (G == 3 && g == 2) a.dnx[row / G + (size_t)d * H + k] = dgc[(3 * NB + b) * KP + k];
          }
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < ITEMS; ++q) {
      const int idx = tid + q * P;
      if (tid < P && idx < Hc * NB) {
        float s = dhrec[q];
#pragma unroll
        for (int gg = 0; gg < G; ++gg) s += p_s[gg * Hc * NB + idx];
        dhrec[q] = s;
      }
    }
    // p_s is rewritten only after the next step's cluster / CTA barrier, dg_s is double-buffered
  }
}

template <int G, int KP, int NB>
int launch_rnn(bool backward, const RnnArgs& a, int ndir, cudaStream_t s) {
  constexpr int CS = rnn_cluster(KP);
  const int P = G * a.Hc;
  const int threads = ((P + 31) / 32) * 32;
  if (threads > 256) return set_err(STG_ERR_UNSUPPORTED, "rnn: %d gate rows per CTA exceed 256 threads", P);
  const size_t smem = backward ? sizeof(float) * ((size_t)2 * 4 * NB * KP + (size_t)G * a.Hc * NB)
                               : sizeof(float) * ((size_t)2 * NB * KP + (size_t)4 * a.Hc * NB);
  auto kern = backward ? k_rnn_bwd<G, KP, NB> : k_rnn_fwd<G, KP, NB>;
  if (smem > 48 * 1024) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return check_cuda("rnn smem attribute");
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(a.ntile * CS), (unsigned)ndir, 1);
  cfg.blockDim = dim3((unsigned)threads, 1, 1);
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

template <int G, int KP>
int dispatch_nb(bool backward, const RnnArgs& a, int ndir, int NB, cudaStream_t s) {
  switch (NB) {
    case 2: return launch_rnn<G, KP, 2>(backward, a, ndir, s);
    case 5: return launch_rnn<G, KP, 5>(backward, a, ndir, s);
    default: return launch_rnn<G, KP, 8>(backward, a, ndir, s);
  }
}
template <int G>
int dispatch_kp(bool backward, const RnnArgs& a, int ndir, int KP, int NB, cudaStream_t s) {
  switch (KP) {
    case 8: return dispatch_nb<G, 8>(backward, a, ndir, NB, s);
    case 32: return dispatch_nb<G, 32>(backward, a, ndir, NB, s);
    case 64: return dispatch_nb<G, 64>(backward, a, ndir, NB, s);
    default: return dispatch_nb<G, 128>(backward, a, ndir, NB, s);
  }
}

int rnn_kp(int H) { return H <= 8 ? 8 : H <= 32 ? 32 : H <= 64 ? 64 : 128; }

int rnn_common(int cell, int T, int B, int H, int ndir, RnnArgs* a, int* G, int* KP, int* NB) {
  if (cell != STG_RNN_LSTM && cell != STG_RNN_GRU) return set_err(STG_ERR_INVALID, "rnn: unknown cell %d", cell);
  if (T < 1 || B < 1 || H < 1) return set_err(STG_ERR_INVALID, "rnn: non-positive dimension");
  if (ndir != 1 && ndir != 2) return set_err(STG_ERR_INVALID, "rnn: ndir must be 1 or 2");
  if (H > 128) return set_err(STG_ERR_UNSUPPORTED, "rnn: hidden size %d > 128", H);
  *G = cell == STG_RNN_LSTM ? 4 : 3;
  *KP = rnn_kp(H);
  *NB = stg_rnn_batch_tile(B);
  const int cs = *KP > 64 ? 2 : 1;
  a->T = T; a->B = B; a->H = H;
  a->Hc = (H + cs - 1) / cs;
  a->ntile = (B + *NB - 1) / *NB;
  return STG_OK;
}

}  // namespace
}  // namespace stg

using namespace stg;

extern "C" int stg_rnn_batch_tile(int B) { return B <= 2 ? 2 : (B <= 5 ? 5 : 8); }

extern "C" size_t stg_rnn_saved_floats(int cell, int T, int B, int H, int ndir) {
  if (T < 1 || B < 1 || H < 1) return 0;
  const int NB = stg_rnn_batch_tile(B);
  const size_t ntile = (size_t)(B + NB - 1) / NB;
  return (size_t)ndir * ntile * T * (cell == STG_RNN_LSTM ? 6 : 5) * H * NB;
}

extern "C" int stg_rnn_forward(int cell, const float* xg_dev, int64_t xg_bstride, int64_t xg_tstride,
                               const float* whh_dev, const float* bhn_dev, int T, int B, int H, int ndir,
                               float* out_dev, int64_t out_bstride, int64_t out_tstride, float* saved_dev,
                               void* stream) {
  RnnArgs a = {};
  int G, KP, NB;
  if (int rc = rnn_common(cell, T, B, H, ndir, &a, &G, &KP, &NB)) return rc;
  if (!xg_dev || !whh_dev || !out_dev) return set_err(STG_ERR_INVALID, "rnn: null pointer");
  if (cell == STG_RNN_GRU && !bhn_dev) return set_err(STG_ERR_INVALID, "rnn: GRU needs b_hn");
  a.xg = xg_dev; a.gsb = xg_bstride; a.gst = xg_tstride;
  a.whh = whh_dev; a.bhn = bhn_dev;
  a.out = out_dev; a.osb = out_bstride; a.ost = out_tstride;
  a.saved = saved_dev;
  cudaStream_t s = (cudaStream_t)stream;
  return G == 4 ? dispatch_kp<4>(false, a, ndir, KP, NB, s) : dispatch_kp<3>(false, a, ndir, KP, NB, s);
}

extern "C" int stg_rnn_backward(int cell, const float* whh_dev, const float* saved_dev, const float* dout_dev,
                                int64_t out_bstride, int64_t out_tstride, int T, int B, int H, int ndir,
                                float* dxg_dev, int64_t xg_bstride, int64_t xg_tstride, float* dnx_dev, void* stream) {
  RnnArgs a = {};
  int G, KP, NB;
  if (int rc = rnn_common(cell, T, B, H, ndir, &a, &G, &KP, &NB)) return rc;
  if (!whh_dev || !saved_dev || !dout_dev || !dxg_dev) return set_err(STG_ERR_INVALID, "rnn: null pointer");
  if (cell == STG_RNN_GRU && !dnx_dev) return set_err(STG_ERR_INVALID, "rnn: GRU backward needs dnx");
  if (cell == STG_RNN_GRU && (xg_bstride % 3 || xg_tstride % 3))
    return set_err(STG_ERR_INVALID, "rnn: GRU xg strides must be multiples of 3 (dnx shares them / 3)");
  a.whh = whh_dev; a.saved = const_cast<float*>(saved_dev);
  a.dout = dout_dev; a.osb = out_bstride; a.ost = out_tstride;
  a.dxg = dxg_dev; a.gsb = xg_bstride; a.gst = xg_tstride; a.dnx = dnx_dev;
  cudaStream_t s = (cudaStream_t)stream;
  return G == 4 ? dispatch_kp<4>(true, a, ndir, KP, NB, s) : dispatch_kp<3>(true, a, ndir, KP, NB, s);
}
