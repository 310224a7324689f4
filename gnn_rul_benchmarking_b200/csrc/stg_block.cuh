// Kernel-side argument structs of the graph-conv block kernels (stg_block.cu) shared with the
// C-ABI layer (stg_capi.cu) and the whole-model engine (stg_engine.cu).
#pragma once
#include "stg_common.cuh"

namespace stg {

struct BlkDev {
  int H, w, stride, L;
  int nchunk_f;     // forward: chunks of windows per sample
  int nchunk_b;     // backward: chunks of time steps per sample
  float decay;
  const float *Wm, *bm, *g0, *b0, *Wt, *bt, *g1, *b1;
  float *rm0, *rv0, *rm1, *rv1;
  float* out;
  long long out_bs;
  float* yp;
  double* stats;
  float* coef;      // training: coefficient table written by k_block_prep (lives behind stats)
  // backward only
  const float* dout;
  long long dout_bs;
  float *dWm, *dbm, *dg0, *db0, *dWt, *dbt, *dg1, *db1, *dxp;
};

struct BlkArgs {
  BlkDev b[2];
  int nblk;
  const float* x;
  int B, T, N, C;
  const double* xmom;   // [2][T][C] sums / sums of squares over (b,n); training only
  float* dx;            // backward finalize output
  int training;
  float momentum, eps;
  // whole-model engine: the FC-head kernels take over BN1+leaky_relu+pool (k_block_fwd_fin) and the
  // BN1 backward sums (k_block_bwd_stats), see stg_head.cu
  int head_fused;
  int prep_done;        // coefficient tables already written by k_xmoments_prep
  int bwd_packed;       // k_block_bwd: windows' threads packed back to back (M = 40, 42 ...) instead of warp-aligned slots
  int fin_elsewhere;    // k_block_bwd_fin's work is done by the encoder's first backward phase
  int dxp_unfolded;     // tcgen05 backward: dxp holds one row per (window, node) [B, L, w*N, C]; the finalize folds it
};

struct BlkPlan {
  int CP, HP;           // padded feature dims the kernel template is instantiated for
  int mma_f, mma_b;     // legacy mma.sync path selected for forward / backward
  int tc, tc_wr;        // tcgen05 / TMEM path (stg_block_tc.cu) and its rows per window slot (32 or 64)
  int tc_split;         // 3-term TF32 product for the projection / Gram / aggregation (see plan_blocks_tc)
  int NT;               // tensor-core path: 8-column tiles per graph (w*N <= 8*NT)
  int threads_f, threads_b;
  int wpc_f, wpc_b;     // windows processed concurrently per CTA
  size_t smem_f, smem_b;
  int grid_x_f, grid_x_b;
};

// Fills nchunk_* of every block and the launch plan; returns 0 or a negative stg_status
// (message in err).  Pure host arithmetic.
int plan_blocks(BlkArgs& a, BlkPlan& p, char* err, size_t errlen);

// Exact maximum of the rows a backward CTA stages (owned time steps plus the halo of the windows
// touching them) over all chunks of all blocks, with the kernels' own chunk arithmetic; also
// returns the largest number of windows a chunk touches.
inline int bwd_rows_exact(const BlkArgs& a, int* max_windows = nullptr) {
  int rm = 0, wm = 0;
  for (int z = 0; z < a.nblk; ++z) {
    const BlkDev& k = a.b[z];
    const int w = k.w, s = k.stride, L = (a.T - w) / s + 1;
    int per = (a.T + k.nchunk_b - 1) / k.nchunk_b;
    per = ((per + s - 1) / s) * s;
    for (int chunk = 0; chunk < k.nchunk_b; ++chunk) {
      const int ta = chunk * per, tb = (a.T < ta + per) ? a.T : ta + per;
      if (ta >= tb) continue;
      int l_lo = ta - (w - 1);
      l_lo = l_lo <= 0 ? 0 : (l_lo + s - 1) / s;
      const int l_hi = (L - 1 < (tb - 1) / s) ? L - 1 : (tb - 1) / s;
      const bool any = l_lo <= l_hi;
      const int t_lo = any ? (ta < l_lo * s ? ta : l_lo * s) : ta;
      const int t_hi = any ? (tb - 1 > l_hi * s + w - 1 ? tb - 1 : l_hi * s + w - 1) : tb - 1;
      const int rows = (t_hi - t_lo + 1) * a.N;
      rm = rows > rm ? rows : rm;
      if (any && l_hi - l_lo + 1 > wm) wm = l_hi - l_lo + 1;
    }
  }
  if (max_windows) *max_windows = wm;
  return rm;
}

// tensor-core path (stg_block_mma.cu)
bool plan_blocks_mma_fwd(BlkArgs& a, BlkPlan& p);
int launch_block_forward_mma(const BlkArgs& a, const BlkPlan& p, cudaStream_t s);

// tcgen05 / TMEM path (stg_block_tc.cu)
bool plan_blocks_tc(BlkArgs& a, BlkPlan& p);
int launch_block_forward_tc(const BlkArgs& a, const BlkPlan& p, cudaStream_t s);
int launch_block_backward_tc(const BlkArgs& a, const BlkPlan& p, cudaStream_t s);

// Launchers (enqueue only).
int launch_xmoments(const float* x, int B, int T, int N, int C, double* xmom, cudaStream_t s);
int launch_xmoments_prep(const BlkArgs& a, const BlkPlan& p, double* xmom, unsigned* counter, cudaStream_t s);
int launch_block_forward(const BlkArgs& a, const BlkPlan& p, cudaStream_t s);
int launch_block_backward(const BlkArgs& a, const BlkPlan& p, cudaStream_t s);

}  // namespace stg
