// FC head of FC_STGNN_RUL (Model.py:30-39,83: Linear(F->J)+ReLU, Linear(J->J)+ReLU, Linear(J->H)+ReLU,
// Linear(H->1)), nn.MSELoss (algorithms.py:44,70) and torch.optim.Adam (algorithms.py:60-64), sm_100a.
//
//   k_head_fc1   z1[b][j] = b1[j] + sum_k feat[b][k] W1[j][k]       skinny GEMM, one CTA per SPB samples
//   k_head_tail  per-sample tail (fc2..fc4), optional MSE + backward of the tail -> d1, dW2..dW4, db*
//   k_head_bwd1  dfeat[b][k] = sum_j d1[b][j] W1[j][k];  dW1[j][k] += sum_b d1[b][j] feat[b][k]
//   k_adam       flat multi-tensor Adam with L2 weight decay folded into the gradient
#include <math.h>
#include <stdlib.h>

#include "stg_mma.cuh"
#include "stg_model.cuh"

namespace stg {
namespace {

#ifdef STG_HEAD_TIMING
// debug build only (scripts/head_cta_times.py): clock64 deltas of up to 8 points per CTA, kernel 0 = fc1, 1 = bwd1
__device__ unsigned long long g_head_time[2][1024][8];
#define HEAD_STAMP(kid, slot)                                                                            \
  {                                                                                                      \
    const unsigned cta_ = blockIdx.y * gridDim.x + blockIdx.x;                                           \
    if (threadIdx.x == 0 && cta_ < 1024) {                                                               \
      if (slot == 0) head_t0 = clock64();                                                                \
      g_head_time[kid][cta_][slot] = slot == 0 ? 1ull : (unsigned long long)(clock64() - head_t0);       \
    }                                                                                                    \
  }
#define HEAD_T0 long long head_t0 = 0;
#else
#define HEAD_STAMP(kid, slot)
#define HEAD_T0
#endif

// ------------------------------------------------------------------------------------------
// BN1 coefficients of the graph-conv blocks from their batch moments: c[z][0]=a1 [1]=c1 [2]=mu1 [3]=r1
STG_DEVINL void head_bn1(const HeadArgs& a, float (*c)[4][64], bool update_running) {
  for (int i = threadIdx.x; i < a.nblk * 64; i += blockDim.x) {
    const int z = i >> 6, h = i & 63;
    const HeadBlk& k = a.blk[z];
    if (h >= k.H) continue;
    const double R = (double)a.B * k.L * k.M, iR = inv_d(R);
    const double m = k.stats[h] * iR;
    double var = k.stats[k.H + h] * iR - m * m;
    if (var < 0.0) var = 0.0;
    const float r1 = (float)rsqrt_d(var + (double)a.eps);
    const float a1 = k.g1[h] * r1;
    c[z][0][h] = a1;
    c[z][1][h] = k.b1[h] - a1 * (float)m;
    c[z][2][h] = (float)m;
    c[z][3][h] = r1;
    if (update_running) {
      const double unb = R > 1.0 ? var * R / (R - 1.0) : var;
      k.rm1[h] = (1.f - a.momentum) * k.rm1[h] + a.momentum * (float)m;
      k.rv1[h] = (1.f - a.momentum) * k.rv1[h] + a.momentum * (float)unb;
    }
  }
}

// z1[b][j] += sum_{k in slice} feat[b][k] W1[j][k]  (+ b1[j] from slice 0).  grid (ceil(B/SPB), KSPLIT).
// FUSED: feat[b][k] = mean_j lrelu(BN1(Y'[b,l,j*N+n,h])) is computed here from the blocks' saved Y'
// (GraphConvpoolMPNN_block_v6 tail, Model_Base.py:103-107,212-216) and written to feat_out.
template <int JP, int SPB, bool FUSED>
__global__ void __launch_bounds__(256) k_head_fc1(const HeadArgs a, int kper) {
  pdl_sync();
  __shared__ float red[8][SPB * JP];
  __shared__ float bc[2][4][64];
  const int b0 = blockIdx.x * SPB, tid = threadIdx.x, J = a.J, F = a.F;
  const int k_lo = blockIdx.y * kper, k_hi = min(F, k_lo + kper);
  if (FUSED) {
    head_bn1(a, bc, blockIdx.x == 0 && blockIdx.y == 0);
    __syncthreads();
  }
  float acc[SPB][JP];
#pragma unroll
  for (int s = 0; s < SPB; ++s)
#pragma unroll
    for (int j = 0; j < JP; ++j) acc[s][j] = 0.f;
  for (int k = k_lo + tid; k < k_hi; k += 256) {
    float f[SPB];
    if (FUSED) {
      const int z = (a.nblk > 1 && k >= a.blk[1].foff) ? 1 : 0;
      const HeadBlk& kb = a.blk[z];
      const int e = k - kb.foff, h = e % kb.H, ln = e / kb.H, n = ln % kb.N, l = ln / kb.N;
      const float a1 = bc[z][0][h], c1 = bc[z][1][h], invw = 1.f / (float)kb.w;
      const size_t jst = (size_t)kb.N * kb.H;
#pragma unroll
      for (int s = 0; s < SPB; ++s) {
        float v = 0.f;
        if (b0 + s < a.B) {
          const float* yp = kb.yp + (((size_t)(b0 + s) * kb.L + l) * kb.M + n) * kb.H + h;
          for (int j = 0; j < kb.w; ++j) v += lrelu(fmaf(a1, yp[j * jst], c1));
          v *= invw;
          a.feat_out[(size_t)(b0 + s) * F + k] = v;
        }
        f[s] = v;
      }
    } else {
#pragma unroll
      for (int s = 0; s < SPB; ++s) f[s] = (b0 + s < a.B) ? a.feat[(size_t)(b0 + s) * F + k] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < JP; ++j)
      if (j < J) {
        const float w = a.W1[(size_t)j * F + k];
#pragma unroll
        for (int s = 0; s < SPB; ++s) acc[s][j] = fmaf(f[s], w, acc[s][j]);
      }
  }
  const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
  for (int s = 0; s < SPB; ++s)
#pragma unroll
    for (int j = 0; j < JP; ++j) {
      const float v = warp_sum(acc[s][j]);
      if (lane == 0) red[warp][s * JP + j] = v;
    }
  __syncthreads();
  if (tid < SPB * JP) {
    const int s = tid / JP, j = tid - s * JP;
    if (j < J && b0 + s < a.B) {
      float v = blockIdx.y == 0 ? a.b1[j] : 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) v += red[w][tid];
      atomicAdd(&a.z1[(size_t)(b0 + s) * J + j], v);
    }
  }
}

// ------------------------------------------------------------------------------------------
// MODE 0: forward only (pred).  MODE 1: backward with given dpred.  MODE 2: fused MSE + backward.
template <int MODE>
__global__ void __launch_bounds__(128) k_head_tail(const HeadArgs a, int TB) {
  pdl_sync();
  extern __shared__ __align__(16) float sm[];
  const int J = a.J, H = a.H, tid = threadIdx.x, nt = blockDim.x;
  const int TBP = TB + 1;
  float* W2 = sm;                 // [J][J]
  float* W3 = W2 + J * J;         // [H][J]
  float* W4 = W3 + H * J;         // [H]
  float* bb = W4 + H;             // b2[J] b3[H] b4[1]
  float* acc = bb + J + H + 1;    // dW2[J*J] db2[J] dW3[H*J] db3[H] dW4[H] db4[1] db1[J] loss[1]
  const int nacc = J * J + J + H * J + H + H + 1 + J + 1;
  float* a1 = acc + nacc;         // [J][TBP]   relu(z1)  (<=0 stored as 0; mask == a1 > 0)
  float* a2 = a1 + J * TBP;       // [J][TBP]
  float* a3 = a2 + J * TBP;       // [H][TBP]
  float* dd2 = a3 + H * TBP;      // [J][TBP]   gradient wrt z2
  float* dd3 = dd2 + J * TBP;     // [H][TBP]   gradient wrt z3
  float* dd1 = dd3 + H * TBP;     // [J][TBP]   gradient wrt z1
  float* dps = dd1 + J * TBP;     // [TBP]      dpred
  for (int i = tid; i < J * J; i += nt) W2[i] = a.W2[i];
  for (int i = tid; i < H * J; i += nt) W3[i] = a.W3[i];
  for (int i = tid; i < H; i += nt) W4[i] = a.W4[i];
  for (int i = tid; i < J; i += nt) bb[i] = a.b2[i];
  for (int i = tid; i < H; i += nt) bb[J + i] = a.b3[i];
  if (tid == 0) bb[J + H] = a.b4[0];
  if (MODE) for (int i = tid; i < nacc; i += nt) acc[i] = 0.f;
  __syncthreads();
  // four threads per sample: each owns every 4th output of a layer; the sample's activations live in
  // its shared-memory column, layers are separated by __syncwarp (the 4 threads share a warp)
  const int sl = tid >> 2, part = tid & 3;          // sample slot in the tile, part of the sample
  const int b = blockIdx.x * TB + sl;
  const bool act = sl < TB && b < a.B;
  float pred = 0.f;
  if (act)
    for (int j = part; j < J; j += 4) a1[j * TBP + sl] = fmaxf(a.z1[(size_t)b * J + j], 0.f);
  __syncwarp();
  if (act)
    for (int i = part; i < J; i += 4) {
      float v = bb[i];
      for (int j = 0; j < J; ++j) v = fmaf(W2[i * J + j], a1[j * TBP + sl], v);
      a2[i * TBP + sl] = fmaxf(v, 0.f);
    }
  __syncwarp();
  if (act)
    for (int i = part; i < H; i += 4) {
      float v = bb[J + i];
      for (int j = 0; j < J; ++j) v = fmaf(W3[i * J + j], a2[j * TBP + sl], v);
      a3[i * TBP + sl] = fmaxf(v, 0.f);
    }
  __syncwarp();
  if (act)
    for (int i = part; i < H; i += 4) pred = fmaf(W4[i], a3[i * TBP + sl], pred);
  pred += __shfl_xor_sync(0xffffffffu, pred, 1);
  pred += __shfl_xor_sync(0xffffffffu, pred, 2);
  pred += bb[J + H];
  if (act && part == 0 && a.pred) a.pred[b] = pred;
  if (MODE == 0) return;
  float dp = 0.f, lossv = 0.f;
  if (act) {
    if (MODE == 2) {
      const float e = pred - a.y[b];
      lossv = part == 0 ? e * e / (float)a.B : 0.f;
      dp = 2.f * e / (float)a.B;
    } else {
      dp = a.dpred[b];
    }
    for (int i = part; i < H; i += 4) dd3[i * TBP + sl] = a3[i * TBP + sl] > 0.f ? dp * W4[i] : 0.f;
    if (part == 0) dps[sl] = dp;
  }
  __syncwarp();
  if (act)
    for (int j = part; j < J; j += 4) {
      float v = 0.f;
      if (a2[j * TBP + sl] > 0.f)
        for (int i = 0; i < H; ++i) v = fmaf(dd3[i * TBP + sl], W3[i * J + j], v);
      dd2[j * TBP + sl] = v;
    }
  __syncwarp();
  if (act)
    for (int j = part; j < J; j += 4) {
      float v = 0.f;
      if (a1[j * TBP + sl] > 0.f)
        for (int i = 0; i < J; ++i) v = fmaf(dd2[i * TBP + sl], W2[i * J + j], v);
      dd1[j * TBP + sl] = v;
      a.d1[(size_t)b * J + j] = v;
    }
  if (MODE == 2) {
    lossv = warp_sum(lossv);
    if ((tid & 31) == 0) atomicAdd(&acc[nacc - 1], lossv);
  }
  __syncthreads();
  const int rows = min(TB, a.B - blockIdx.x * TB);
  // weight gradients: one owner thread per entry, loop over the CTA's samples
  float* gW2 = acc; float* gb2 = gW2 + J * J; float* gW3 = gb2 + J; float* gb3 = gW3 + H * J;
  float* gW4 = gb3 + H; float* gb4 = gW4 + H; float* gb1 = gb4 + 1;
  for (int e = tid; e < nacc - 1; e += nt) {
    float v = 0.f;
    if (e < J * J) {
      const int i = e / J, j = e - i * J;
      for (int r = 0; r < rows; ++r) v = fmaf(dd2[i * TBP + r], a1[j * TBP + r], v);
    } else if (e < J * J + J) {
      const int i = e - J * J;
      for (int r = 0; r < rows; ++r) v += dd2[i * TBP + r];
    } else if (e < J * J + J + H * J) {
      const int q = e - (J * J + J), i = q / J, j = q - i * J;
      for (int r = 0; r < rows; ++r) v = fmaf(dd3[i * TBP + r], a2[j * TBP + r], v);
    } else if (e < J * J + J + H * J + H) {
      const int i = e - (J * J + J + H * J);
      for (int r = 0; r < rows; ++r) v += dd3[i * TBP + r];
    } else if (e < J * J + J + H * J + 2 * H) {
      const int i = e - (J * J + J + H * J + H);
      for (int r = 0; r < rows; ++r) v = fmaf(dps[r], a3[i * TBP + r], v);
    } else if (e == J * J + J + H * J + 2 * H) {
      for (int r = 0; r < rows; ++r) v += dps[r];
    } else {
      const int j = e - (J * J + J + H * J + 2 * H + 1);
      for (int r = 0; r < rows; ++r) v += dd1[j * TBP + r];
    }
    acc[e] = v;
  }
  __syncthreads();
  for (int i = tid; i < J * J; i += nt) atomicAdd(&a.dW2[i], gW2[i]);
  for (int i = tid; i < J; i += nt) { atomicAdd(&a.db2[i], gb2[i]); atomicAdd(&a.db1[i], gb1[i]); }
  for (int i = tid; i < H * J; i += nt) atomicAdd(&a.dW3[i], gW3[i]);
  for (int i = tid; i < H; i += nt) { atomicAdd(&a.db3[i], gb3[i]); atomicAdd(&a.dW4[i], gW4[i]); }
  if (tid == 0) {
    atomicAdd(&a.db4[0], gb4[0]);
    if (MODE == 2 && a.loss) atomicAdd(a.loss, acc[nacc - 1]);
  }
}

static size_t tail_smem(int J, int H, int TB) {
  const int TBP = TB + 1;
  const size_t nacc = (size_t)J * J + J + (size_t)H * J + H + H + 1 + J + 1;
  return 4 * ((size_t)J * J + (size_t)H * J + H + J + H + 1 + nacc + (size_t)(4 * J + 2 * H + 1) * TBP);
}

// ------------------------------------------------------------------------------------------
// Tail of the head (fc2..fc4 + loss/backward) for the samples [b_lo, b_lo+nb) of one k_head_bwd1 CTA:
// every CTA of a sample slice recomputes it (a few thousand MACs) instead of waiting for a separate
// k_head_tail launch; only the slice's first CTA (owner) accumulates the tail's parameter gradients,
// the loss and the predictions.  Result: d1s[(b-b_lo)*J + j] = dLoss/dz1.  All threads must call it.
template <bool HAVE_Y>
__device__ void tail_slice(const HeadArgs& a, int b_lo, int nb, float* d1s, float* scr, bool owner) {
  const int J = a.J, H = a.H, tid = threadIdx.x, nt = blockDim.x;
  const int NBP = nb + 1;
  float* W2 = scr;                float* W3 = W2 + J * J;        float* W4 = W3 + H * J;
  float* bb = W4 + H;             float* a1 = bb + J + H + 1;    float* a2 = a1 + J * NBP;
  float* a3 = a2 + J * NBP;       float* dd2 = a3 + H * NBP;     float* dd3 = dd2 + J * NBP;
  float* dps = dd3 + H * NBP;
  for (int i = tid; i < J * J; i += nt) W2[i] = a.W2[i];
  for (int i = tid; i < H * J; i += nt) W3[i] = a.W3[i];
  for (int i = tid; i < H; i += nt) W4[i] = a.W4[i];
  for (int i = tid; i < J; i += nt) bb[i] = a.b2[i];
  for (int i = tid; i < H; i += nt) bb[J + i] = a.b3[i];
  if (tid == 0) bb[J + H] = a.b4[0];
  for (int e = tid; e < nb * J; e += nt) a1[(e % J) * NBP + e / J] = fmaxf(a.z1[(size_t)b_lo * J + e], 0.f);
  __syncthreads();
  for (int e = tid; e < nb * J; e += nt) {                 // fc2 + ReLU
    const int sl = e / J, i = e - sl * J;
    float v = bb[i];
    for (int j = 0; j < J; ++j) v = fmaf(W2[i * J + j], a1[j * NBP + sl], v);
    a2[i * NBP + sl] = fmaxf(v, 0.f);
  }
  __syncthreads();
  for (int e = tid; e < nb * H; e += nt) {                 // fc3 + ReLU
    const int sl = e / H, i = e - sl * H;
    float v = bb[J + i];
    for (int j = 0; j < J; ++j) v = fmaf(W3[i * J + j], a2[j * NBP + sl], v);
    a3[i * NBP + sl] = fmaxf(v, 0.f);
  }
  __syncthreads();
  float lossv = 0.f;
  for (int sl = tid; sl < nb; sl += nt) {                  // fc4, loss, dpred
    float pred = bb[J + H];
    for (int i = 0; i < H; ++i) pred = fmaf(W4[i], a3[i * NBP + sl], pred);
    const int b = b_lo + sl;
    float dp;
    if (HAVE_Y) {
      const float e = pred - a.y[b];
      lossv += e * e / (float)a.B;
      dp = 2.f * e / (float)a.B;
    } else {
      dp = a.dpred[b];
    }
    dps[sl] = dp;
    if (owner && a.pred) a.pred[b] = pred;
  }
  if (HAVE_Y && owner) {
    lossv = warp_sum(lossv);
    if ((tid & 31) == 0 && lossv != 0.f && a.loss) atomicAdd(a.loss, lossv);
  }
  __syncthreads();
  for (int e = tid; e < nb * H; e += nt) {
    const int sl = e / H, i = e - sl * H;
    dd3[i * NBP + sl] = a3[i * NBP + sl] > 0.f ? dps[sl] * W4[i] : 0.f;
  }
  __syncthreads();
  for (int e = tid; e < nb * J; e += nt) {
    const int sl = e / J, j = e - sl * J;
    float v = 0.f;
    if (a2[j * NBP + sl] > 0.f)
      for (int i = 0; i < H; ++i) v = fmaf(dd3[i * NBP + sl], W3[i * J + j], v);
    dd2[j * NBP + sl] = v;
  }
  __syncthreads();
  for (int e = tid; e < nb * J; e += nt) {
    const int sl = e / J, j = e - sl * J;
    float v = 0.f;
    if (a1[j * NBP + sl] > 0.f)
      for (int i = 0; i < J; ++i) v = fmaf(dd2[i * NBP + sl], W2[i * J + j], v);
    d1s[sl * J + j] = v;
  }
  __syncthreads();
  if (!owner) return;
  // parameter gradients of the tail over this slice's samples: one owner thread per entry
  const int nacc = J * J + J + H * J + H + H + 1 + J;
  for (int e = tid; e < nacc; e += nt) {
    float v = 0.f;
    if (e < J * J) {
      const int i = e / J, j = e - i * J;
      for (int r = 0; r < nb; ++r) v = fmaf(dd2[i * NBP + r], a1[j * NBP + r], v);
      atomicAdd(&a.dW2[e], v);
    } else if (e < J * J + J) {
      const int i = e - J * J;
      for (int r = 0; r < nb; ++r) v += dd2[i * NBP + r];
      atomicAdd(&a.db2[i], v);
    } else if (e < J * J + J + H * J) {
      const int q = e - (J * J + J), i = q / J, j = q - i * J;
      for (int r = 0; r < nb; ++r) v = fmaf(dd3[i * NBP + r], a2[j * NBP + r], v);
      atomicAdd(&a.dW3[q], v);
    } else if (e < J * J + J + H * J + H) {
      const int i = e - (J * J + J + H * J);
      for (int r = 0; r < nb; ++r) v += dd3[i * NBP + r];
      atomicAdd(&a.db3[i], v);
    } else if (e < J * J + J + H * J + 2 * H) {
      const int i = e - (J * J + J + H * J + H);
      for (int r = 0; r < nb; ++r) v = fmaf(dps[r], a3[i * NBP + r], v);
      atomicAdd(&a.dW4[i], v);
    } else if (e == J * J + J + H * J + 2 * H) {
      for (int r = 0; r < nb; ++r) v += dps[r];
      atomicAdd(&a.db4[0], v);
    } else {
      const int j = e - (J * J + J + H * J + 2 * H + 1);
      for (int r = 0; r < nb; ++r) v += d1s[r * J + j];
      atomicAdd(&a.db1[j], v);
    }
  }
}

static size_t tail_slice_floats(int J, int H, int nb) {
  const size_t NBP = nb + 1;
  return (size_t)J * J + (size_t)H * J + H + J + H + 1 + (size_t)(3 * J + 2 * H + 1) * NBP + 4;
}

// ------------------------------------------------------------------------------------------
// grid (ceil(F/128), BS); thread = one feature column k, samples [b_lo, b_hi) of slice blockIdx.y
//   dfeat[b][k] = sum_j d1[b][j] W1[j][k];  dW1[j][k] += sum_b d1[b][j] feat[b][k]
// FUSED: also the BatchNorm-1 backward sums of the graph-conv blocks (what k_block_bwd_stats computes):
//   stats[2H+h] += sum dYn, stats[3H+h] += sum dYn*Yhat, with dYn = dfeat/w * lrelu'(BN1(Y')).
template <int JP, bool FUSED, int TAIL>      // TAIL: 0 = d1 given, 1 = tail from dpred, 2 = tail with MSE loss
__global__ void __launch_bounds__(128) k_head_bwd1(const HeadArgs a, int bper) {
  pdl_sync();
  extern __shared__ __align__(16) float sm[];   // d1 slice [bper][J] (+ tail scratch)
  __shared__ float bc[2][4][64];
  __shared__ float sred[2][2][64];
  const int J = a.J, F = a.F, tid = threadIdx.x;
  const int b_lo = blockIdx.y * bper, b_hi = min(a.B, b_lo + bper);
  if (TAIL == 0) {
    for (int i = tid; i < (b_hi - b_lo) * J; i += blockDim.x) sm[i] = a.d1[(size_t)b_lo * J + i];
  } else {
    float* scr = sm + ((bper * J + 3) / 4) * 4;
    if (TAIL == 2) tail_slice<true>(a, b_lo, b_hi - b_lo, sm, scr, blockIdx.x == 0);
    else tail_slice<false>(a, b_lo, b_hi - b_lo, sm, scr, blockIdx.x == 0);
  }
  if (FUSED) {
    head_bn1(a, bc, false);
    for (int i = tid; i < 2 * 2 * 64; i += blockDim.x) (&sred[0][0][0])[i] = 0.f;
  }
  __syncthreads();
  const int k = blockIdx.x * blockDim.x + tid;
  if (k < F) {
    float wc[JP], gw[JP];
#pragma unroll
    for (int j = 0; j < JP; ++j) {
      wc[j] = j < J ? a.W1[(size_t)j * F + k] : 0.f;
      gw[j] = 0.f;
    }
    int z = 0, h = 0, w = 1;
    size_t ybase = 0, jst = 0, bst = 0;
    float a1 = 0.f, c1 = 0.f, mu = 0.f, r1 = 0.f, invw = 1.f, s1 = 0.f, s2 = 0.f;
    const float* yp = nullptr;
    if (FUSED) {
      z = (a.nblk > 1 && k >= a.blk[1].foff) ? 1 : 0;
      const HeadBlk& kb = a.blk[z];
      const int e = k - kb.foff;
      h = e % kb.H;
      const int ln = e / kb.H, n = ln % kb.N, l = ln / kb.N;
      w = kb.w; invw = 1.f / (float)w;
      ybase = ((size_t)l * kb.M + n) * kb.H + h;
      jst = (size_t)kb.N * kb.H;
      bst = (size_t)kb.L * kb.M * kb.H;
      yp = kb.yp;
      a1 = bc[z][0][h]; c1 = bc[z][1][h]; mu = bc[z][2][h]; r1 = bc[z][3][h];
    }
#pragma unroll 4
    for (int b = b_lo; b < b_hi; ++b) {
      const float f = a.feat[(size_t)b * F + k];
      const float* d = sm + (b - b_lo) * J;
      float df = 0.f;
#pragma unroll
      for (int j = 0; j < JP; ++j)
        if (j < J) {
          const float dj = d[j];
          df = fmaf(dj, wc[j], df);
          gw[j] = fmaf(dj, f, gw[j]);
        }
      a.dfeat[(size_t)b * F + k] = df;
      if (FUSED) {
        const float* y = yp + (size_t)b * bst + ybase;
        for (int j = 0; j < w; ++j) {
          const float yv = y[j * jst];
          const float dyn = df * invw * (fmaf(a1, yv, c1) > 0.f ? 1.f : kLeaky);
          s1 += dyn;
          s2 = fmaf(dyn, (yv - mu) * r1, s2);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < JP; ++j)
      if (j < J) atomicAdd(&a.dW1[(size_t)j * F + k], gw[j]);
    if (FUSED) {
      atomicAdd(&sred[z][0][h], s1);
      atomicAdd(&sred[z][1][h], s2);
    }
  }
  // (the two shared atomics above are per thread; columns of one warp map to <= 32/H distinct (z,h)
  //  pairs, contention stays below the global-load latency this kernel is bound by)
  if (FUSED) {
    __syncthreads();
    for (int i = tid; i < a.nblk * 64; i += blockDim.x) {
      const int z = i >> 6, h = i & 63;
      const HeadBlk& kb = a.blk[z];
      if (h < kb.H) {
        atomicAdd(&kb.stats[2 * kb.H + h], (double)sred[z][0][h]);
        atomicAdd(&kb.stats[3 * kb.H + h], (double)sred[z][1][h]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                              float* __restrict__ v, long long n, const long long* __restrict__ step,
                                              float lr, float b1, float b2, float eps, float wd, float gscale) {
  pdl_sync();
  __shared__ float s_bc[2];
  if (threadIdx.x == 0) {                      // bias corrections once per CTA (double pow is slow)
    const double t = (double)(*step);
    s_bc[0] = (float)(1.0 - pow((double)b1, t));
    s_bc[1] = (float)sqrt(1.0 - pow((double)b2, t));
  }
  __syncthreads();
  const float bc1 = s_bc[0], bc2s = s_bc[1];
  const float step_size = lr / bc1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float pv = p[i];
    const float gv = fmaf(wd, pv, g[i] * gscale);
    const float mv = fmaf(1.f - b1, gv - m[i], m[i]);            // torch: lerp(m, g, 1-b1)
    const float vv = fmaf(1.f - b2, gv * gv, b2 * v[i]);
    m[i] = mv;
    v[i] = vv;
    const float denom = sqrtf(vv) / bc2s + eps;
    p[i] = pv - step_size * (mv / denom);
  }
}

struct TickArgs { long long* p[16]; int n; float* zero1; };
__global__ void k_tick(const TickArgs t) {
  pdl_sync();
  if (threadIdx.x < t.n && t.p[threadIdx.x]) *t.p[threadIdx.x] += 1;
}

// first kernel of a step: clears the reduction scratch, increments the step counters
// (num_batches_tracked, dropout counter) and clears one extra word (the caller's loss accumulator)
__global__ void __launch_bounds__(256) k_zero(float4* p, size_t n4, const TickArgs t) {
  pdl_sync();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
    p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (blockIdx.x == 0) {
    if (threadIdx.x < t.n && t.p[threadIdx.x]) *t.p[threadIdx.x] += 1;
    if (threadIdx.x == 0 && t.zero1) *t.zero1 = 0.f;
  }
}


// ------------------------------------------------------------------------------------------
// Tensor-core head for the usual shape (J <= 16, features fused from the blocks' saved Y', windows of two time
// steps, H % 8 == 0 so that every aligned group of 8 feature columns is one (window, sensor) of one block): fc1 and
// its backward as warp-level mma.sync m16n8k8 TF32 products with the 3xTF32 split (fp32-level accuracy).  One warp
// owns a [16 samples x 32 columns] (forward) or [bper samples x 16 columns] (backward) piece; every global load of
// the piece is issued before the first use, the K-split partial results meet in shared memory / float atomics.
// (The one-thread-per-column kernels above spend their time in 16..32 serialized load round trips and in one
// five-shuffle warp sum per accumulator.)
STG_DEVINL void head_group(const HeadArgs& a, int k0, int& z, size_t& off, int& h0) {
  z = (a.nblk > 1 && k0 >= a.blk[1].foff) ? 1 : 0;
  const HeadBlk& kb = a.blk[z];
  const int e = k0 - kb.foff;
  h0 = e % kb.H;
  const int ln = e / kb.H, n = ln % kb.N, l = ln / kb.N;
  off = ((size_t)l * kb.M + n) * kb.H + h0;
}

// grid (ceil(B/16), ceil(F/128)), 4 warps; warp = 32 feature columns (4 k-steps) of 16 samples
__global__ void __launch_bounds__(128) k_head_fc1_mma(const HeadArgs a) {
  pdl_sync();
  __shared__ float bc[2][4][64];
  __shared__ float red[4][16 * 16];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int F = a.F, J = a.J;
  const int b0 = blockIdx.x * 16;
  const int kw0 = (blockIdx.y * 4 + warp) * 32;
  HEAD_T0
  HEAD_STAMP(0, 0)
  head_bn1(a, bc, blockIdx.x == 0 && blockIdx.y == 0);
  __syncthreads();
  HEAD_STAMP(0, 1)
  const bool rok[2] = {b0 + g < a.B, b0 + g + 8 < a.B};
  float y[4][2][2][2];      // [k-step][column t / t+4][row g / g+8][time step of the window]
  float wv[4][2][2];        // [k-step][n-tile][column]
  float ca[4][2], cc[4][2];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int k0 = kw0 + ks * 8;
    const bool kok = k0 < F;
    int z = 0, h0 = 0;
    size_t off = 0;
    if (kok) head_group(a, k0, z, off, h0);
    const HeadBlk& kb = a.blk[z];
    const size_t bst = (size_t)kb.L * kb.M * kb.H, jst = (size_t)kb.N * kb.H;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int col = t + 4 * c;
      ca[ks][c] = bc[z][0][h0 + col];
      cc[ks][c] = bc[z][1][h0 + col];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const bool ok = kok && rok[r];
        const float* p = kb.yp + (size_t)(b0 + g + 8 * r) * bst + off + col;
        y[ks][c][r][0] = ok ? p[0] : 0.f;
        y[ks][c][r][1] = ok ? p[jst] : 0.f;
      }
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int j = nt * 8 + g;
        wv[ks][nt][c] = (kok && j < J) ? a.W1[(size_t)j * F + k0 + col] : 0.f;
      }
    }
  }
  float acc[2][4];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[nt][q] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int k0 = kw0 + ks * 8;
    float fv[2][2];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float v = 0.f;
        if (k0 < F && rok[r]) {
          v = 0.5f * (lrelu(fmaf(ca[ks][c], y[ks][c][r][0], cc[ks][c])) + lrelu(fmaf(ca[ks][c], y[ks][c][r][1], cc[ks][c])));
          a.feat_out[(size_t)(b0 + g + 8 * r) * F + k0 + t + 4 * c] = v;
        }
        fv[c][r] = v;
      }
    const FragA A = make_a(fv[0][0], fv[0][1], fv[1][0], fv[1][1]);
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) mma3(acc[nt], A, make_b(wv[ks][nt][0], wv[ks][nt][1]));
  }
  HEAD_STAMP(0, 2)
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    float* r = red[warp] + nt * 8 + 2 * t;
    r[g * 16] = acc[nt][0];
    r[g * 16 + 1] = acc[nt][1];
    r[(g + 8) * 16] = acc[nt][2];
    r[(g + 8) * 16 + 1] = acc[nt][3];
  }
  __syncthreads();
  for (int e = tid; e < 256; e += 128) {
    const int row = e >> 4, j = e & 15;
    if (j < J && b0 + row < a.B) {
      float v = (red[0][e] + red[1][e]) + (red[2][e] + red[3][e]);
      if (blockIdx.y == 0) v += a.b1[j];
      atomicAdd(&a.z1[(size_t)(b0 + row) * J + j], v);
    }
  }
  HEAD_STAMP(0, 3)
}

// Tail of the head for the (up to) 32 samples of one CTA of 128 threads: fc2..fc4, loss / dpred and the backward of
// the tail.  Four threads per sample (thread = sample * 4 + part, so a sample stays inside one warp and the layers
// only need __syncwarp), each computing a quarter of every layer's outputs from the full input vector in registers;
// the vectors are exchanged through ts ([row][sample], 33-float pitch) and stay there for the parameter gradients.
// J <= 16, H <= 8, zero padded.  wsm (staged by the CTA): W2[16][17] W3[8][17] W4[8] b2[16] b3[8] b4[1].
// Result: d1s[sample * J + j] = dLoss/dz1.
constexpr int kTwP = 17, kTwW3 = 16 * kTwP, kTwW4 = kTwW3 + 8 * kTwP, kTwB2 = kTwW4 + 8, kTwB3 = kTwB2 + 16,
              kTwB4 = kTwB3 + 8, kTailW = kTwB4 + 1;
STG_DEVINL void tail_stage(const HeadArgs& a, float* wsm) {
  const int J = a.J, H = a.H;
  for (int i = threadIdx.x; i < kTailW; i += blockDim.x) {
    float v = 0.f;
    if (i < kTwW3) { const int r = i / kTwP, c = i - r * kTwP; if (r < J && c < J) v = a.W2[r * J + c]; }
    else if (i < kTwW4) { const int q = i - kTwW3, r = q / kTwP, c = q - r * kTwP; if (r < H && c < J) v = a.W3[r * J + c]; }
    else if (i < kTwB2) { if (i - kTwW4 < H) v = a.W4[i - kTwW4]; }
    else if (i < kTwB3) { if (i - kTwB2 < J) v = a.b2[i - kTwB2]; }
    else if (i < kTwB4) { if (i - kTwB3 < H) v = a.b3[i - kTwB3]; }
    else v = a.b4[0];
    wsm[i] = v;
  }
}
constexpr int kTsA1 = 0, kTsA2 = 16, kTsA3 = 32, kTsD3 = 40, kTsD2 = 48, kTsD1 = 64, kTsDp = 80, kTsRows = 81, kTsP = 33;
template <bool HAVE_Y>
STG_DEVINL void tail_cta(const HeadArgs& a, int b_lo, int nb, float* d1s, const float* wsm, float* ts, bool owner) {
  const int J = a.J, tid = threadIdx.x, s = tid >> 2, p = tid & 3, b = b_lo + s;
  const bool act = s < nb;
  const float *W2 = wsm, *W3 = wsm + kTwW3, *W4 = wsm + kTwW4, *b2 = wsm + kTwB2, *b3 = wsm + kTwB3;
  float* tl = ts + s;
  float a1[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) a1[j] = (act && j < J) ? fmaxf(a.z1[(size_t)b * J + j], 0.f) : 0.f;
  if (p == 0) {
#pragma unroll
    for (int j = 0; j < 16; ++j) tl[(kTsA1 + j) * kTsP] = a1[j];
  }
  float a2o[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {                     // fc2 + ReLU: rows 4p .. 4p+3
    const int i = p * 4 + q;
    float v = b2[i];
#pragma unroll
    for (int j = 0; j < 16; ++j) v = fmaf(W2[i * kTwP + j], a1[j], v);
    a2o[q] = fmaxf(v, 0.f);
    tl[(kTsA2 + i) * kTsP] = a2o[q];
  }
  __syncwarp();
  float a2[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) a2[j] = tl[(kTsA2 + j) * kTsP];
#pragma unroll
  for (int q = 0; q < 2; ++q) {                     // fc3 + ReLU: rows 2p, 2p+1
    const int i = p * 2 + q;
    float v = b3[i];
#pragma unroll
    for (int j = 0; j < 16; ++j) v = fmaf(W3[i * kTwP + j], a2[j], v);
    tl[(kTsA3 + i) * kTsP] = fmaxf(v, 0.f);
  }
  __syncwarp();
  float a3[8], pred = wsm[kTwB4];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a3[i] = tl[(kTsA3 + i) * kTsP];
    pred = fmaf(W4[i], a3[i], pred);
  }
  float dp = 0.f, lossv = 0.f;
  if (act) {
    if (HAVE_Y) {
      const float e = pred - a.y[b];
      if (p == 0) lossv = e * e / (float)a.B;
      dp = 2.f * e / (float)a.B;
    } else {
      dp = a.dpred[b];
    }
    if (owner && p == 0 && a.pred) a.pred[b] = pred;
  }
  float dd3[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) dd3[i] = a3[i] > 0.f ? dp * W4[i] : 0.f;
  if (p == 0) {
    tl[kTsDp * kTsP] = dp;
#pragma unroll
    for (int i = 0; i < 8; ++i) tl[(kTsD3 + i) * kTsP] = dd3[i];
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {                     // through fc3 and the ReLU of fc2: columns 4p .. 4p+3
    const int j = p * 4 + q;
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) v = fmaf(dd3[i], W3[i * kTwP + j], v);
    tl[(kTsD2 + j) * kTsP] = a2o[q] > 0.f ? v : 0.f;
  }
  __syncwarp();
  float dd2[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) dd2[j] = tl[(kTsD2 + j) * kTsP];
#pragma unroll
  for (int q = 0; q < 4; ++q) {                     // through fc2 and the ReLU of fc1
    const int j = p * 4 + q;
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) v = fmaf(dd2[i], W2[i * kTwP + j], v);
    v = tl[(kTsA1 + j) * kTsP] > 0.f ? v : 0.f;
    tl[(kTsD1 + j) * kTsP] = v;
    if (act && j < J) d1s[s * J + j] = v;
  }
  if (HAVE_Y && owner) {
    lossv = warp_sum(lossv);
    if ((tid & 31) == 0 && lossv != 0.f && a.loss) atomicAdd(a.loss, lossv);
  }
}

// parameter gradients of the tail over the (up to 32) samples whose vectors tail_cta left in ts: all threads of the
// owner CTA, one entry per thread and pass
STG_DEVINL void tail_param_grads(const HeadArgs& a, const float* ts) {
  const int J = a.J, H = a.H;
  for (int e = threadIdx.x; e < 433; e += blockDim.x) {
    int ra, rb = -1;
    float* dst = nullptr;
    if (e < 256) { const int i = e >> 4, j = e & 15; ra = kTsD2 + i; rb = kTsA1 + j; if (i < J && j < J) dst = &a.dW2[i * J + j]; }
    else if (e < 384) { const int i = (e - 256) >> 4, j = e & 15; ra = kTsD3 + i; rb = kTsA2 + j; if (i < H && j < J) dst = &a.dW3[i * J + j]; }
    else if (e < 400) { ra = kTsD2 + (e - 384); if (e - 384 < J) dst = &a.db2[e - 384]; }
    else if (e < 408) { ra = kTsD3 + (e - 400); if (e - 400 < H) dst = &a.db3[e - 400]; }
    else if (e < 416) { ra = kTsDp; rb = kTsA3 + (e - 408); if (e - 408 < H) dst = &a.dW4[e - 408]; }
    else if (e == 416) { ra = kTsDp; dst = &a.db4[0]; }
    else { ra = kTsD1 + (e - 417); if (e - 417 < J) dst = &a.db1[e - 417]; }
    if (!dst) continue;
    const float* pa = ts + ra * kTsP;
    float v = 0.f;
    if (rb >= 0) {
      const float* pb = ts + rb * kTsP;
#pragma unroll 8
      for (int r = 0; r < 32; ++r) v = fmaf(pa[r], pb[r], v);
    } else {
#pragma unroll 8
      for (int r = 0; r < 32; ++r) v += pa[r];
    }
    atomicAdd(dst, v);
  }
}

// grid (ceil(F/64), ceil(B/bper)), 4 warps; warp = 16 feature columns (two n-tiles) of the slice's bper = 16*MT samples
template <int TAIL, int MT>
__global__ void __launch_bounds__(128) k_head_bwd1_mma(const __grid_constant__ HeadArgs a) {
  pdl_sync();
  constexpr int bper = 16 * MT;
  HEAD_T0
  HEAD_STAMP(1, 0)
  extern __shared__ __align__(16) float sm[];   // d1 slice [bper][J] + tail scratch
  __shared__ float bc[2][4][64];
  __shared__ float sred[2][2][64];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int J = a.J, F = a.F;
  const int b_lo = blockIdx.y * bper, nb = min(a.B - b_lo, bper);
  static_assert(TAIL == 0 || MT == 2, "the in-kernel tail handles 32 samples with 128 threads");
  for (int i = tid; i < bper * J; i += 128) sm[i] = (TAIL == 0 && i < nb * J) ? a.d1[(size_t)b_lo * J + i] : 0.f;
  for (int i = tid; i < 2 * 2 * 64; i += 128) (&sred[0][0][0])[i] = 0.f;
  float* wsm = sm + ((bper * J + 3) / 4) * 4;
  float* ts = wsm + (kTailW + 3) / 4 * 4;
  if (TAIL != 0) tail_stage(a, wsm);
  head_bn1(a, bc, false);
  __syncthreads();
  HEAD_STAMP(1, 1)
  if (TAIL != 0) {
    if (TAIL == 2) tail_cta<true>(a, b_lo, nb, sm, wsm, ts, blockIdx.x == 0);
    else tail_cta<false>(a, b_lo, nb, sm, wsm, ts, blockIdx.x == 0);
    __syncthreads();
    if (blockIdx.x == 0) tail_param_grads(a, ts);
  }
  HEAD_STAMP(1, 2)
  const int kc0 = (blockIdx.x * 4 + warp) * 16;
  // ---- every global load of the piece first
  float wv[2][2][2];              // W1[ks*8 + t (+4)][column g of n-tile]: B operand of dfeat = d1 . W1
  float fe[2 * MT][2][2];         // feat[b_lo + ks*8 + t (+4)][column g of n-tile]: B operand of dW1 = d1^T . feat
  float2 yv[MT][2][2][2];         // Y'[sample g (+8) of m-tile][columns 2t, 2t+1 of n-tile][time step]
  int zc[2], hc[2];
  bool nok[2];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    const int k0 = kc0 + nt * 8;
    nok[nt] = k0 < F;
    int z = 0, h0 = 0;
    size_t off = 0;
    if (nok[nt]) head_group(a, k0, z, off, h0);
    zc[nt] = z;
    hc[nt] = h0;
    const HeadBlk& kb = a.blk[z];
    const size_t bst = (size_t)kb.L * kb.M * kb.H, jst = (size_t)kb.N * kb.H;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int j = ks * 8 + t + 4 * c;
        wv[nt][ks][c] = (nok[nt] && j < J) ? a.W1[(size_t)j * F + k0 + g] : 0.f;
      }
#pragma unroll
    for (int ks = 0; ks < 2 * MT; ++ks)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int b = ks * 8 + t + 4 * c;
        fe[ks][nt][c] = (nok[nt] && b < nb) ? a.feat[(size_t)(b_lo + b) * F + k0 + g] : 0.f;
      }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int b = mt * 16 + g + 8 * r;
        const bool ok = nok[nt] && b < nb;
        const float* p = kb.yp + (size_t)(b_lo + b) * bst + off + 2 * t;
        yv[mt][nt][r][0] = ok ? *reinterpret_cast<const float2*>(p) : make_float2(0.f, 0.f);
        yv[mt][nt][r][1] = ok ? *reinterpret_cast<const float2*>(p + jst) : make_float2(0.f, 0.f);
      }
  }
  // ---- dfeat[b][k] = sum_j d1[b][j] W1[j][k]  and the BatchNorm-1 backward sums of the blocks
  FragB wf[2][2];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) wf[nt][ks] = make_b(wv[nt][ks][0], wv[nt][ks][1]);
  float s1[2][2], s2[2][2];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) s1[nt][0] = s1[nt][1] = s2[nt][0] = s2[nt][1] = 0.f;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    float c[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int q = 0; q < 4; ++q) c[nt][q] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const float* d = sm + (mt * 16 + g) * J + ks * 8 + t;
      const bool c0 = ks * 8 + t < J, c1 = ks * 8 + t + 4 < J;
      const FragA A = make_a(c0 ? d[0] : 0.f, c0 ? d[8 * J] : 0.f, c1 ? d[4] : 0.f, c1 ? d[8 * J + 4] : 0.f);
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) mma3(c[nt], A, wf[nt][ks]);
    }
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      if (!nok[nt]) continue;
      const int k = kc0 + nt * 8 + 2 * t, z = zc[nt], h = hc[nt] + 2 * t;
      const float a1x = bc[z][0][h], c1x = bc[z][1][h], mux = bc[z][2][h], r1x = bc[z][3][h];
      const float a1y = bc[z][0][h + 1], c1y = bc[z][1][h + 1], muy = bc[z][2][h + 1], r1y = bc[z][3][h + 1];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int b = mt * 16 + g + 8 * r;
        if (b >= nb) continue;
        const float dx = c[nt][2 * r], dy = c[nt][2 * r + 1];
        *reinterpret_cast<float2*>(a.dfeat + (size_t)(b_lo + b) * F + k) = make_float2(dx, dy);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const float2 yy = yv[mt][nt][r][j];
          const float gx = dx * 0.5f * (fmaf(a1x, yy.x, c1x) > 0.f ? 1.f : kLeaky);
          const float gy = dy * 0.5f * (fmaf(a1y, yy.y, c1y) > 0.f ? 1.f : kLeaky);
          s1[nt][0] += gx;
          s1[nt][1] += gy;
          s2[nt][0] = fmaf(gx, (yy.x - mux) * r1x, s2[nt][0]);
          s2[nt][1] = fmaf(gy, (yy.y - muy) * r1y, s2[nt][1]);
        }
      }
    }
  }
  HEAD_STAMP(1, 3)
  // ---- dW1[j][k] += sum_b d1[b][j] feat[b][k]
  {
    float gw[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int q = 0; q < 4; ++q) gw[nt][q] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2 * MT; ++ks) {
      const float* d = sm + (ks * 8 + t) * J + g;
      const bool j0 = g < J, j1 = g + 8 < J;
      const FragA A = make_a(j0 ? d[0] : 0.f, j1 ? d[8] : 0.f, j0 ? d[4 * J] : 0.f, j1 ? d[4 * J + 8] : 0.f);
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) mma3(gw[nt], A, make_b(fe[ks][nt][0], fe[ks][nt][1]));
    }
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      if (!nok[nt]) continue;
      const int k = kc0 + nt * 8 + 2 * t;
      if (g < J) {
        atomicAdd(&a.dW1[(size_t)g * F + k], gw[nt][0]);
        atomicAdd(&a.dW1[(size_t)g * F + k + 1], gw[nt][1]);
      }
      if (g + 8 < J) {
        atomicAdd(&a.dW1[(size_t)(g + 8) * F + k], gw[nt][2]);
        atomicAdd(&a.dW1[(size_t)(g + 8) * F + k + 1], gw[nt][3]);
      }
    }
  }
  HEAD_STAMP(1, 4)
  // ---- BN1 backward sums: over the 8 sample lanes of the warp, then shared / global atomics
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      float u = s1[nt][p], v = s2[nt][p];
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        u += __shfl_xor_sync(0xffffffffu, u, o);
        v += __shfl_xor_sync(0xffffffffu, v, o);
      }
      if (g == 0 && nok[nt]) {
        atomicAdd(&sred[zc[nt]][0][hc[nt] + 2 * t + p], u);
        atomicAdd(&sred[zc[nt]][1][hc[nt] + 2 * t + p], v);
      }
    }
  __syncthreads();
  for (int i = tid; i < a.nblk * 64; i += 128) {
    const int z = i >> 6, h = i & 63;
    const HeadBlk& kb = a.blk[z];
    if (h < kb.H) {
      atomicAdd(&kb.stats[2 * kb.H + h], (double)sred[z][0][h]);
      atomicAdd(&kb.stats[3 * kb.H + h], (double)sred[z][1][h]);
    }
  }
  HEAD_STAMP(1, 5)
}

// ------------------------------------------------------------------------------------------
// Wide heads (J > 32, e.g. FD003: F = 24 864, J = 48).  There fc1 and its backward are real GEMMs and the
// one-thread-per-column kernels above run out of registers / re-read W1 once per sample; these versions tile
// them through shared memory with register blocking.
constexpr int kWT = 256;            // threads
constexpr int kWBM = 128;           // samples per CTA tile (fc1)
constexpr int kWKC = 32;            // contraction chunk (fc1)
constexpr int kWFP = kWBM + 4;      // pitch of the transposed feature tile (16-byte aligned rows)

// z1[b][j] += sum_{k in slice} feat[b][k] W1[j][k] (+ b1[j] from slice 0).  grid (ceil(B/128), KSPLIT).
// Thread (tx = tid%16, ty = tid/16): JQ = JP/16 outputs j = tx*JQ.. for the 8 samples ty*8.. of the tile.
template <int JP, bool FUSED>
__global__ void __launch_bounds__(kWT) k_head_fc1_wide(const HeadArgs a, int kper) {
  pdl_sync();
  constexpr int JQ = JP / 16;
  __shared__ __align__(16) float fT[kWKC][kWFP];     // feat tile, [k][sample]
  __shared__ float Ws[JP][kWKC + 1];                 // W1 tile, [j][k]
  __shared__ float bc[2][4][64];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, J = a.J, F = a.F;
  const int b0 = blockIdx.x * kWBM;
  const int k_lo = blockIdx.y * kper, k_hi = min(F, k_lo + kper);
  if (FUSED) head_bn1(a, bc, blockIdx.x == 0 && blockIdx.y == 0);
  float acc[8][JQ];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int u = 0; u < JQ; ++u) acc[i][u] = 0.f;
  const int pk = tid & 31, pb = tid >> 5;            // producer: column pk of the chunk, samples pb + 8*i
  for (int kc = k_lo; kc < k_hi; kc += kWKC) {
    __syncthreads();                                 // previous chunk consumed (and bc ready)
    {
      const int k = kc + pk;
      const bool kin = k < k_hi;
      const float* yp = nullptr;
      size_t jst = 0, bst = 0;
      float a1 = 0.f, c1 = 0.f, invw = 1.f;
      int w = 0;
      if (FUSED && kin) {
        const int z = (a.nblk > 1 && k >= a.blk[1].foff) ? 1 : 0;
        const HeadBlk& kb = a.blk[z];
        const int e = k - kb.foff, h = e % kb.H, ln = e / kb.H, n = ln % kb.N, l = ln / kb.N;
        a1 = bc[z][0][h]; c1 = bc[z][1][h];
        w = kb.w; invw = 1.f / (float)w;
        jst = (size_t)kb.N * kb.H;
        bst = (size_t)kb.L * kb.M * kb.H;
        yp = kb.yp + ((size_t)l * kb.M + n) * kb.H + h;
      }
#pragma unroll 4
      for (int i = 0; i < kWBM / 8; ++i) {
        const int bl = pb + 8 * i, b = b0 + bl;
        float v = 0.f;
        if (kin && b < a.B) {
          if (FUSED) {
            const float* y = yp + (size_t)b * bst;
            for (int j = 0; j < w; ++j) v += lrelu(fmaf(a1, y[j * jst], c1));
            v *= invw;
            a.feat_out[(size_t)b * F + k] = v;
          } else {
            v = a.feat[(size_t)b * F + k];
          }
        }
        fT[pk][bl] = v;
      }
      for (int j = pb; j < JP; j += 8) Ws[j][pk] = (kin && j < J) ? a.W1[(size_t)j * F + k] : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int kk = 0; kk < kWKC; ++kk) {
      const float4 f0 = *reinterpret_cast<const float4*>(&fT[kk][ty * 8]);
      const float4 f1 = *reinterpret_cast<const float4*>(&fT[kk][ty * 8 + 4]);
      const float f[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
#pragma unroll
      for (int u = 0; u < JQ; ++u) {
        const float wv = Ws[tx * JQ + u][kk];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i][u] = fmaf(f[i], wv, acc[i][u]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int b = b0 + ty * 8 + i;
#pragma unroll
    for (int u = 0; u < JQ; ++u) {
      const int j = tx * JQ + u;
      if (b < a.B && j < J) atomicAdd(&a.z1[(size_t)b * J + j], acc[i][u] + (blockIdx.y == 0 ? a.b1[j] : 0.f));
    }
  }
}

// One CTA per 128 feature columns, all samples (d1 given in a.d1 by k_head_tail):
//   dfeat[b][k] = sum_j d1[b][j] W1[j][k]      thread (kx = tid%32 -> 4 columns, by = tid/32 -> 8 samples of a 64-chunk)
//   dW1[j][k]  += sum_b d1[b][j] feat[b][k]    thread (kx -> 4 columns, jy = tid/32 -> JP/8 rows j), no atomics
// FUSED: BatchNorm-1 backward sums of the graph-conv blocks from dfeat and the saved Y' (as k_head_bwd1).
constexpr int kWBC = 64;            // samples per chunk (bwd1)
template <int JP, bool FUSED>
__global__ void __launch_bounds__(kWT, 2) k_head_bwd1_wide(const HeadArgs a) {
  pdl_sync();
  constexpr int JR = JP / 8;
  extern __shared__ __align__(16) float sm[];
  float* Ws = sm;                         // [JP][128]
  float* fs = Ws + JP * 128;              // [64][128]
  float* ds = fs + kWBC * 128;            // [64][JP]
  __shared__ float bc[2][4][64];
  __shared__ float sred[2][2][64];
  const int tid = threadIdx.x, kx = tid & 31, wy = tid >> 5, J = a.J, F = a.F;
  const int k0 = blockIdx.x * 128, kq = k0 + kx * 4;
  if (FUSED) {
    head_bn1(a, bc, false);
    for (int i = tid; i < 2 * 2 * 64; i += kWT) (&sred[0][0][0])[i] = 0.f;
  }
  for (int i = tid; i < JP * 128; i += kWT) {
    const int j = i >> 7, c = i & 127;
    Ws[i] = (j < J && k0 + c < F) ? a.W1[(size_t)j * F + k0 + c] : 0.f;
  }
  __syncthreads();
  // per-column constants of the fused statistics (shared: the register file is needed for the two tiles)
  __shared__ unsigned cyo[128];
  __shared__ int cjst[128], cbst[128], cwz[128];          // cwz = w | z << 8 | h << 16
  __shared__ float ccf[4][128];                           // a1, c1, mu, r1
  if (FUSED && tid < 128) {
    const int k = k0 + tid;
    cyo[tid] = 0; cjst[tid] = cbst[tid] = 0; cwz[tid] = 0;
    ccf[0][tid] = ccf[1][tid] = ccf[2][tid] = ccf[3][tid] = 0.f;
    if (k < F) {
      const int z = (a.nblk > 1 && k >= a.blk[1].foff) ? 1 : 0;
      const HeadBlk& kb = a.blk[z];
      const int e = k - kb.foff, h = e % kb.H, ln = e / kb.H, n = ln % kb.N, l = ln / kb.N;
      cyo[tid] = (unsigned)((l * kb.M + n) * kb.H + h);
      cjst[tid] = kb.N * kb.H;
      cbst[tid] = kb.L * kb.M * kb.H;
      cwz[tid] = kb.w | (z << 8) | (h << 16);
      ccf[0][tid] = bc[z][0][h]; ccf[1][tid] = bc[z][1][h]; ccf[2][tid] = bc[z][2][h]; ccf[3][tid] = bc[z][3][h];
    }
  }
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  float gw[JR][4];
#pragma unroll
  for (int u = 0; u < JR; ++u)
#pragma unroll
    for (int q = 0; q < 4; ++q) gw[u][q] = 0.f;

  for (int bc0 = 0; bc0 < a.B; bc0 += kWBC) {
    const int nb = min(kWBC, a.B - bc0);
    __syncthreads();
    for (int i = tid; i < kWBC * 128; i += kWT) {
      const int bl = i >> 7, c = i & 127;
      fs[i] = (bl < nb && k0 + c < F) ? a.feat[(size_t)(bc0 + bl) * F + k0 + c] : 0.f;
    }
    for (int i = tid; i < kWBC * JP; i += kWT) {
      const int bl = i / JP, j = i - bl * JP;
      ds[i] = (bl < nb && j < J) ? a.d1[(size_t)(bc0 + bl) * J + j] : 0.f;
    }
    __syncthreads();
    // ---- dfeat for samples wy*8 .. wy*8+7 of the chunk, columns kq..kq+3
    {
      float df[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) df[i][q] = 0.f;
#pragma unroll 4
      for (int j = 0; j < JP; ++j) {
        const float4 w4 = *reinterpret_cast<const float4*>(&Ws[j * 128 + kx * 4]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float dj = ds[(wy * 8 + i) * JP + j];
          df[i][0] = fmaf(dj, w4.x, df[i][0]);
          df[i][1] = fmaf(dj, w4.y, df[i][1]);
          df[i][2] = fmaf(dj, w4.z, df[i][2]);
          df[i][3] = fmaf(dj, w4.w, df[i][3]);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int bl = wy * 8 + i, b = bc0 + bl;
        if (bl < nb) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (kq + q < F) {
              a.dfeat[(size_t)b * F + kq + q] = df[i][q];
              if (FUSED) {
                const int c = kx * 4 + q, wz = cwz[c], w = wz & 255;
                const float* y = ((wz >> 8) & 1 ? a.blk[1].yp : a.blk[0].yp) + (size_t)b * cbst[c] + cyo[c];
                const float invw = 1.f / (float)w, ca = ccf[0][c], cc = ccf[1][c], cm = ccf[2][c], cr = ccf[3][c];
                for (int jj = 0; jj < w; ++jj) {
                  const float yv = y[jj * cjst[c]];
                  const float dyn = df[i][q] * invw * (fmaf(ca, yv, cc) > 0.f ? 1.f : kLeaky);
                  s1[q] += dyn;
                  s2[q] = fmaf(dyn, (yv - cm) * cr, s2[q]);
                }
              }
            }
        }
      }
    }
    // ---- dW1 rows wy*JR .. wy*JR+JR-1, columns kq..kq+3, over this chunk's samples
#pragma unroll 4
    for (int bl = 0; bl < kWBC; ++bl) {
      const float4 f4 = *reinterpret_cast<const float4*>(&fs[bl * 128 + kx * 4]);
#pragma unroll
      for (int u = 0; u < JR; ++u) {
        const float dj = ds[bl * JP + wy * JR + u];
        gw[u][0] = fmaf(dj, f4.x, gw[u][0]);
        gw[u][1] = fmaf(dj, f4.y, gw[u][1]);
        gw[u][2] = fmaf(dj, f4.z, gw[u][2]);
        gw[u][3] = fmaf(dj, f4.w, gw[u][3]);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < JR; ++u) {
    const int j = wy * JR + u;
    if (j < J)
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (kq + q < F) a.dW1[(size_t)j * F + kq + q] += gw[u][q];      // this CTA is the column's only writer
  }
  if (FUSED) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (kq + q < F) {
        const int wz = cwz[kx * 4 + q];
        atomicAdd(&sred[(wz >> 8) & 1][0][wz >> 16], s1[q]);
        atomicAdd(&sred[(wz >> 8) & 1][1][wz >> 16], s2[q]);
      }
    __syncthreads();
    for (int i = tid; i < a.nblk * 64; i += kWT) {
      const int z = i >> 6, h = i & 63;
      const HeadBlk& kb = a.blk[z];
      if (h < kb.H) {
        atomicAdd(&kb.stats[2 * kb.H + h], (double)sred[z][0][h]);
        atomicAdd(&kb.stats[3 * kb.H + h], (double)sred[z][1][h]);
      }
    }
  }
}

template <int JP>
void fc1_wide_launch(const HeadArgs& a, cudaStream_t s) {
  const int gx = (a.B + kWBM - 1) / kWBM;
  int ksplit = (2 * 148 + gx - 1) / gx;
  if (const char* e = getenv("STG_FC1_KSPLIT")) { const int v = atoi(e); if (v >= 1) ksplit = v; }
  const int maxsplit = (a.F + 4 * kWKC - 1) / (4 * kWKC);
  if (ksplit > maxsplit) ksplit = maxsplit;
  if (ksplit < 1) ksplit = 1;
  const int kper = (((a.F + ksplit - 1) / ksplit) + kWKC - 1) / kWKC * kWKC;
  ksplit = (a.F + kper - 1) / kper;
  if (a.fused_blocks) launch_pdl(k_head_fc1_wide<JP, true>, dim3(gx, ksplit), dim3(kWT), 0, s, a, kper);
  else launch_pdl(k_head_fc1_wide<JP, false>, dim3(gx, ksplit), dim3(kWT), 0, s, a, kper);
}
template <int JP>
void bwd1_wide_launch(const HeadArgs& a, cudaStream_t s) {
  const size_t smem = (size_t)(JP * 128 + kWBC * 128 + kWBC * JP) * 4;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(k_head_bwd1_wide<JP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(k_head_bwd1_wide<JP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    attr_done = true;
  }
  const int gx = (a.F + 127) / 128;
  if (a.fused_blocks) launch_pdl(k_head_bwd1_wide<JP, true>, dim3(gx), dim3(kWT), smem, s, a);
  else launch_pdl(k_head_bwd1_wide<JP, false>, dim3(gx), dim3(kWT), smem, s, a);
}
inline bool head_wide(const HeadArgs& a) {
  static const bool off = getenv("STG_HEAD_NARROW") != nullptr;
  return !off && a.J > 32 && a.F >= 2048;
}

// tensor-core head: see k_head_fc1_mma
inline bool head_mma(const HeadArgs& a) {
  static const bool off = getenv("STG_HEAD_SIMT") != nullptr;
  if (off || !a.fused_blocks || a.J > 16 || a.H > 8 || a.F % 8 || a.nblk < 1) return false;
  for (int z = 0; z < a.nblk; ++z)
    if (a.blk[z].w != 2 || a.blk[z].H % 8 || a.blk[z].H > 64 || a.blk[z].foff % 8) return false;
  return true;
}
constexpr int kBwd1MT = 2;      // backward sample slice = 32 samples

template <int JP, int SPB>
void fc1_launch(const HeadArgs& a, cudaStream_t s) {
  const int gx = (a.B + SPB - 1) / SPB;
  int ksplit = (2 * 148 + gx - 1) / gx;              // about two CTAs per SM
  if (const char* e = getenv("STG_FC1_KSPLIT")) { const int v = atoi(e); if (v >= 1) ksplit = v; }
  const int maxsplit = (a.F + 255) / 256;
  if (ksplit > maxsplit) ksplit = maxsplit;
  if (ksplit < 1) ksplit = 1;
  const int kper = (((a.F + ksplit - 1) / ksplit) + 3) / 4 * 4;
  ksplit = (a.F + kper - 1) / kper;
  if (a.fused_blocks) launch_pdl(k_head_fc1<JP, SPB, true>, dim3(gx, ksplit), dim3(256), 0, s, a, kper);
  else launch_pdl(k_head_fc1<JP, SPB, false>, dim3(gx, ksplit), dim3(256), 0, s, a, kper);
}
template <int JP>
void bwd1_launch(const HeadArgs& a, cudaStream_t s) {
  const int gx = (a.F + 127) / 128;
  int slices = (512 + gx - 1) / gx;                  // ~3.5 CTAs of 128 threads per SM (measured optimum on S1)
  if (const char* e = getenv("STG_BWD1_SLICES")) { const int v = atoi(e); if (v >= 1) slices = v; }
  if (slices > a.B) slices = a.B;
  int bper = (a.B + slices - 1) / slices;
  if ((size_t)bper * a.J * 4 > 40 * 1024) bper = (int)(40 * 1024 / (a.J * 4));
  if (a.fused_blocks && bper > 32) bper = 32;          // the in-kernel tail keeps ~(3J+2H) floats per sample
  slices = (a.B + bper - 1) / bper;
  const size_t sm_d1 = (size_t)((bper * a.J + 3) / 4) * 4 * 4;
  static bool attr_done = false;                       // per instantiation: dynamic smem may exceed 48 KB
  if (!attr_done) {
    cudaFuncSetAttribute(k_head_bwd1<JP, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(k_head_bwd1<JP, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(k_head_bwd1<JP, false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    attr_done = true;
  }
  if (a.fused_blocks) {
    // model path: the tail (fc2..fc4, loss, its backward) is recomputed per slice inside this kernel
    const size_t smt = sm_d1 + tail_slice_floats(a.J, a.H, bper) * 4;
    if (a.y) launch_pdl(k_head_bwd1<JP, true, 2>, dim3(gx, slices), dim3(128), smt, s, a, bper);
    else launch_pdl(k_head_bwd1<JP, true, 1>, dim3(gx, slices), dim3(128), smt, s, a, bper);
  } else {
    launch_pdl(k_head_bwd1<JP, false, 0>, dim3(gx, slices), dim3(128), sm_d1, s, a, bper);
  }
}

int tail_tb(int J) { (void)J; return 32; }
bool g_tail_attr[64] = {};
void tail_attrs() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (g_tail_attr[dev]) return;
  cudaFuncSetAttribute(k_head_tail<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_head_tail<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_head_tail<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  g_tail_attr[dev] = true;
}

}  // namespace

#ifdef STG_HEAD_TIMING
extern "C" int stg_debug_head_cta_times(unsigned long long* out) {
  return cudaMemcpyFromSymbol(out, g_head_time, sizeof(g_head_time)) == cudaSuccess ? 0 : -1;
}
#endif

int launch_head_forward(const HeadArgs& a, cudaStream_t s) {
  if (a.J > 64 || a.H > 64) return -2;
  {
    ProfScope ps(kProfHeadFc1, s);
    if (head_wide(a)) { if (a.J <= 48) fc1_wide_launch<48>(a, s); else fc1_wide_launch<64>(a, s); }
    else if (head_mma(a)) launch_pdl(k_head_fc1_mma, dim3((a.B + 15) / 16, (a.F + 127) / 128), dim3(128), 0, s, a);
    else if (a.J <= 16) fc1_launch<16, 2>(a, s);
    else if (a.J <= 32) fc1_launch<32, 2>(a, s);
    else if (a.J <= 48) fc1_launch<48, 1>(a, s);
    else fc1_launch<64, 1>(a, s);
  }
  if (a.pred && !a.y && !a.dpred) {      // plain forward: finish the tail now
    tail_attrs();
    const int TB = tail_tb(a.J);
    ProfScope ps(kProfHeadTail, s);
    launch_pdl(k_head_tail<0>, dim3((a.B + TB - 1) / TB), dim3(128), tail_smem(a.J, a.H, TB), s, a, TB);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

int launch_head_backward(const HeadArgs& a, cudaStream_t s) {
  if (a.J > 64 || a.H > 64) return -2;
  tail_attrs();
  const int TB = tail_tb(a.J);
  const bool wide = head_wide(a);
  if (!a.fused_blocks || wide) {         // op path / wide heads: separate tail kernel writes d1
    ProfScope ps(kProfHeadTail, s);
    if (a.y) launch_pdl(k_head_tail<2>, dim3((a.B + TB - 1) / TB), dim3(128), tail_smem(a.J, a.H, TB), s, a, TB);
    else launch_pdl(k_head_tail<1>, dim3((a.B + TB - 1) / TB), dim3(128), tail_smem(a.J, a.H, TB), s, a, TB);
  }
  ProfScope ps(kProfHeadBwd1, s);
  if (wide) { if (a.J <= 48) bwd1_wide_launch<48>(a, s); else bwd1_wide_launch<64>(a, s); }
  else if (head_mma(a)) {
    constexpr int bper = 16 * kBwd1MT;
    const size_t smem = ((size_t)((bper * a.J + 3) / 4) * 4 + (kTailW + 3) / 4 * 4 + kTsRows * kTsP) * 4;
    const dim3 grid((a.F + 63) / 64, (a.B + bper - 1) / bper);
    if (a.y) launch_pdl(k_head_bwd1_mma<2, kBwd1MT>, grid, dim3(128), smem, s, a);
    else launch_pdl(k_head_bwd1_mma<1, kBwd1MT>, grid, dim3(128), smem, s, a);
  }
  else if (a.J <= 16) bwd1_launch<16>(a, s);
  else if (a.J <= 32) bwd1_launch<32>(a, s);
  else if (a.J <= 48) bwd1_launch<48>(a, s);
  else bwd1_launch<64>(a, s);
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

int launch_zero(void* p, size_t bytes, long long* const* counters, int ncounters, float* zero1, cudaStream_t s) {
  // bytes must be a multiple of 16 and p 16-byte aligned (workspace segments are)
  const size_t n4 = bytes / 16;
  int grid = (int)((n4 + 255) / 256);
  if (grid > 1184) grid = 1184;
  if (grid < 1) grid = 1;
  TickArgs t = {};
  t.n = ncounters > 16 ? 16 : ncounters;
  for (int i = 0; i < t.n; ++i) t.p[i] = counters[i];
  t.zero1 = zero1;
  ProfScope ps(kProfZero, s);
  launch_pdl(k_zero, dim3(grid), dim3(256), 0, s, reinterpret_cast<float4*>(p), n4, t);
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

int launch_adam(float* p, const float* g, float* m, float* v, long long n, long long* step, float lr, float b1,
                float b2, float eps, float wd, float gscale, cudaStream_t s) {
  TickArgs t = {};
  t.p[0] = step;
  t.n = 1;
  launch_pdl(k_tick, dim3(1), dim3(32), 0, s, t);
  int grid = (int)((n + 255) / 256);
  if (grid > 1184) grid = 1184;
  ProfScope ps(kProfAdam, s);
  launch_pdl(k_adam, dim3(grid), dim3(256), 0, s, p, g, m, v, n, step, lr, b1, b2, eps, wd, gscale);
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

int launch_tick(long long* const* counters, int n, cudaStream_t s) {
  TickArgs t = {};
  t.n = n > 16 ? 16 : n;
  for (int i = 0; i < t.n; ++i) t.p[i] = counters[i];
  launch_pdl(k_tick, dim3(1), dim3(32), 0, s, t);
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

}  // namespace stg
