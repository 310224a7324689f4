// Kernel-side argument structs of the whole-model engine (encoder, head, Adam) shared between
// stg_encoder.cu, stg_head.cu and the C-ABI layer stg_model_capi.cu.
#pragma once
#include "stg_common.cuh"

namespace stg {

// ---- patch encoder + positional encoding (Model_Base.py:12-41,111-134; Model.py:18-22,45-68) ----
struct EncArgs {
  // dims
  int B, N, T, P, K, EH, E, C;
  int L1, L2, EL2, pad1;      // conv output lengths, E*L2, conv1 padding (K/2)
  int R;                      // rows = B*T*N
  int TR;                     // rows per CTA tile (<= 256)
  // tensors
  const float* X;             // [B, N, T*P]
  const float *W1, *W2, *W3, *b3;
  const float *g1, *be1, *g2, *be2, *g3, *be3;
  float *rm1, *rv1, *rm2, *rv2, *rm3, *rv3;
  const float* pe;            // [>=T, C]
  const float* keep;          // optional [B*N, T, C] 0/1 mask
  unsigned long long seed;
  const long long* seed_ptr;  // optional device-side step counter mixed into the seed (CUDA-graph replays)
  float pdrop;
  int training;
  float momentum, eps;
  double* st;                 // stats scratch, see enc_stat_off()
  // raw activations / masked gradients kept between phases in training: [tile][feature][256] (fast path),
  // [row][feature] (any-dimension kernels)
  float *c2raw, *z3raw, *dn2, *dn1;
  float* h;                   // [R, C] output of the forward
  const float* dh;            // [R, C] gradient wrt h
  float *dW1, *dW2, *dW3, *db3, *dg1, *dbe1, *dg2, *dbe2, *dg3, *dbe3;
  // Fast path only: phase B1 also finishes the graph-conv blocks' backward (what k_block_bwd_fin does):
  //   dh = sum_z dxp_z - r0_z*cnt_z(t)*(m1_z + xhat_z*m2_z), written to dh_out, BN0/BN1 affine gradients.
  struct Fin {
    int nblk, CP;
    int unfolded;               // dxp has one row per (window, node): [B, L, w*N, C] (tcgen05 block backward)
    const float* dxp[2];        // [R, C] per block: dx before the BN0 mean terms
    const float* tab[2];        // block coefficient table: mu0[CP] r0[CP] ...
    const double* stats[2];     // block sums (see stg_block_desc.stats)
    const float* g0[2];
    float *dg0[2], *db0[2], *dg1[2], *db1[2];
    int H[2], w[2], stride[2], L[2];
    float* dh_out;              // [R, C]
  } fin;
};
// stats scratch layout (doubles): forward  [S1: 2*EH][S2: 2*E][S3: 2*C]
//                                 backward [B3: 2*C][B2: 2*E][B1: 2*EH]
inline __host__ __device__ int enc_stats_doubles(int EH, int E, int C) { return 4 * (EH + E + C); }

int plan_encoder(EncArgs& a, size_t* smem_fwd, size_t* smem_bwd, char* err, size_t errlen);
// Dropout of the encoder output without a mask tensor (forward and backward re-derive the same decisions): one
// 64-bit mix per row of (seed, step counter, row), then one 32-bit mix per element of the row.
__device__ __forceinline__ unsigned long long drop_row_key(const EncArgs& a, size_t row) {
  unsigned long long z = a.seed + (a.seed_ptr ? (unsigned long long)*a.seed_ptr * 0xD1B54A32D192ED03ull : 0ull) +
                         (unsigned long long)row * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// keep / (1 - p) scale of element c of the row (C elements per row); 1 outside training
__device__ __forceinline__ float drop_scale(const EncArgs& a, unsigned long long key, size_t row, int C, int c) {
  if (!a.training || a.pdrop <= 0.f) return 1.f;
  const float sc = 1.f / (1.f - a.pdrop);
  if (a.keep) return a.keep[row * C + c] * sc;
  unsigned h = (unsigned)key ^ ((unsigned)(key >> 32) + (unsigned)c * 0x9E3779B9u);
  h ^= h >> 16; h *= 0x85EBCA6Bu;
  h ^= h >> 13; h *= 0xC2B2AE35u;
  h ^= h >> 16;
  const float u = (float)(h >> 8) * (1.f / 16777216.f);
  return u >= a.pdrop ? sc : 0.f;
}

int launch_encoder_forward(const EncArgs& a, size_t smem, cudaStream_t s);
int launch_encoder_backward(const EncArgs& a, size_t smem, cudaStream_t s);
// compile-time-dimension variants (stg_encoder_fast.cu) for the register-sized hyper-parameter sets
bool encoder_fast_available(const EncArgs& a);
int launch_encoder_fast(const EncArgs& a, bool backward, cudaStream_t s);

// ---- FC head (Model.py:30-39,83) + MSE -------------------------------------------------------
// One graph-conv block as the head sees it (training): pre-BN Y' saved by k_block_fwd, its batch
// moments, BN1 parameters.  Lets k_head_fc1 apply BN1 + leaky_relu + window mean on the fly (replaces
// k_block_fwd_fin) and k_head_bwd1 accumulate the BN1 backward sums (replaces k_block_bwd_stats).
struct HeadBlk {
  const float* yp;            // [B, L, w*N, H]
  double* stats;              // [0,H) sum Y'  [H,2H) sum Y'^2  [2H,3H) sum dYn  [3H,4H) sum dYn*Yhat
  const float *g1, *b1;
  float *rm1, *rv1;
  int L, M, H, N, w;
  int foff;                   // first feature column of this block
};
struct HeadArgs {
  int B, F, J, H;             // J = 2H
  int fused_blocks;           // 1: features come from blk[].yp (training), 0: feat is already final
  int nblk;
  HeadBlk blk[2];
  float momentum, eps;
  float* feat_out;            // [B, F] written by k_head_fc1 when fused_blocks
  const float* feat;          // [B, F]
  const float *W1, *b1, *W2, *b2, *W3, *b3, *W4, *b4;
  float* z1;                  // [B, J] pre-activation of fc1 (workspace)
  float* d1;                  // [B, J] gradient wrt z1 (workspace)
  float* pred;                // [B]
  const float* y;             // [B] targets (fused loss) or nullptr
  const float* dpred;         // [B] upstream gradient (autograd path) or nullptr
  float* loss;                // [1] accumulated mean squared error (fused loss)
  float* dfeat;               // [B, F]
  float *dW1, *db1, *dW2, *db2, *dW3, *db3, *dW4, *db4;
};
int launch_head_forward(const HeadArgs& a, cudaStream_t s);               // feat -> z1 -> pred
int launch_head_backward(const HeadArgs& a, cudaStream_t s);              // (y | dpred) -> grads, dfeat

// ---- misc -------------------------------------------------------------------------------------
int launch_zero(void* p, size_t bytes, long long* const* counters, int ncounters, float* zero1, cudaStream_t s);
int launch_adam(float* p, const float* g, float* m, float* v, long long n, long long* step, float lr, float b1,
                float b2, float eps, float wd, float gscale, cudaStream_t s);
int launch_tick(long long* const* counters, int n, cudaStream_t s);

}  // namespace stg
