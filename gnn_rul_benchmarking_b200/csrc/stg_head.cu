// FC head of FC_STGNN_RUL (Model.py:30-39,83: Linear(F->J)+ReLU, Linear(J->J)+ReLU, Linear(J->H)+ReLU,
// Linear(H->1)), nn.MSELoss (algorithms.py:44,70) and torch.optim.Adam (algorithms.py:60-64), sm_100a.
//
//   k_head_fc1   z1[b][j] = b1[j] + sum_k feat[b][k] W1[j][k]       skinny GEMM, one CTA per SPB samples
//   k_head_tail  per-sample tail (fc2..fc4), optional MSE + backward of the tail -> d1, dW2..dW4, db*
//   k_head_bwd1  dfeat[b][k] = sum_j d1[b][j] W1[j][k];  dW1[j][k] += sum_b d1[b][j] feat[b][k]
//   k_adam       flat multi-tensor Adam with L2 weight decay folded into the gradient
#include <math.h>
#include <stdlib.h>

#include "stg_model.cuh"

namespace stg {
namespace {

// ------------------------------------------------------------------------------------------
// BN1 coefficients of the graph-conv blocks from their batch moments: c[z][0]=a1 [1]=c1 [2]=mu1 [3]=r1
STG_DEVINL void head_bn1(const HeadArgs& a, float (*c)[4][64], bool update_running) {
  for (int i = threadIdx.x; i < a.nblk * 64; i += blockDim.x) {
    const int z = i >> 6, h = i & 63;
    const HeadBlk& k = a.blk[z];
    if (h >= k.H) continue;
    const double R = (double)a.B * k.L * k.M;
    const double m = k.stats[h] / R;
    double var = k.stats[k.H + h] / R - m * m;
    if (var < 0.0) var = 0.0;
    const float r1 = (float)(1.0 / sqrt(var + (double)a.eps));
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

int launch_head_forward(const HeadArgs& a, cudaStream_t s) {
  if (a.J > 64 || a.H > 64) return -2;
  {
    ProfScope ps(kProfHeadFc1, s);
    if (head_wide(a)) { if (a.J <= 48) fc1_wide_launch<48>(a, s); else fc1_wide_launch<64>(a, s); }
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
