// extern "C" surface declared in include/stgconv_b200.h -- op-level entry points.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "../../include/stgconv_b200.h"
#include "stg_block.cuh"

namespace stg {
thread_local char g_err[512] = "";
int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
int check_cuda(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err(STG_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return STG_OK;
}

static int fill_args(BlkArgs& a, const float* x, int B, int T, int N, int C, const stg_block_desc* blk,
                     const stg_block_grads* gr, int nblk, const double* xmom, int training, float momentum,
                     float eps) {
  if (!x || !blk) return set_err(STG_ERR_INVALID, "null x / block descriptor");
  if (B < 1 || T < 1 || N < 1 || C < 1) return set_err(STG_ERR_INVALID, "non-positive dimension");
  if (C & 1) return set_err(STG_ERR_UNSUPPORTED, "input_dim C=%d must be even (the reference uses C = 2*hidden_dim)", C);
  if (((uintptr_t)x & 15) != 0) return set_err(STG_ERR_INVALID, "x must be 16-byte aligned");
  if (nblk < 1 || nblk > STG_MAX_BLOCKS) return set_err(STG_ERR_INVALID, "nblk must be 1..%d", STG_MAX_BLOCKS);
  if (training && !xmom) return set_err(STG_ERR_INVALID, "training mode needs xmom (stg_block_xmoments)");
  memset(&a, 0, sizeof(a));
  a.nblk = nblk; a.x = x; a.B = B; a.T = T; a.N = N; a.C = C;
  a.xmom = xmom; a.training = training; a.momentum = momentum; a.eps = eps;
  for (int z = 0; z < nblk; ++z) {
    const stg_block_desc& d = blk[z];
    BlkDev& k = a.b[z];
    if (!d.Wm || !d.bm || !d.bn0_w || !d.bn0_b || !d.bn0_rm || !d.bn0_rv || !d.Wt || !d.bt || !d.bn1_w ||
        !d.bn1_b || !d.bn1_rm || !d.bn1_rv)
      return set_err(STG_ERR_INVALID, "block %d: null parameter pointer", z);
    if (d.H < 1) return set_err(STG_ERR_INVALID, "block %d: H < 1", z);
    if (training && (!d.yp || !d.stats)) return set_err(STG_ERR_INVALID, "block %d: training needs yp and stats", z);
    k.H = d.H; k.w = d.w; k.stride = d.stride; k.decay = d.decay;
    k.Wm = d.Wm; k.bm = d.bm; k.g0 = d.bn0_w; k.b0 = d.bn0_b; k.rm0 = d.bn0_rm; k.rv0 = d.bn0_rv;
    k.Wt = d.Wt; k.bt = d.bt; k.g1 = d.bn1_w; k.b1 = d.bn1_b; k.rm1 = d.bn1_rm; k.rv1 = d.bn1_rv;
    k.out = d.out; k.out_bs = d.out_bstride; k.yp = d.yp; k.stats = d.stats;
    k.coef = d.stats ? reinterpret_cast<float*>(d.stats + ((STG_BLOCK_SUMS_DOUBLES(C, d.H) + 1) / 2) * 2) : nullptr;
    if (gr) {
      const stg_block_grads& g = gr[z];
      if (!g.dout || !g.dWm || !g.dbm || !g.dbn0_w || !g.dbn0_b || !g.dWt || !g.dbt || !g.dbn1_w || !g.dbn1_b ||
          !g.dxp)
        return set_err(STG_ERR_INVALID, "block %d: null gradient pointer", z);
      k.dout = g.dout; k.dout_bs = g.dout_bstride;
      k.dWm = g.dWm; k.dbm = g.dbm; k.dg0 = g.dbn0_w; k.db0 = g.dbn0_b;
      k.dWt = g.dWt; k.dbt = g.dbt; k.dg1 = g.dbn1_w; k.db1 = g.dbn1_b; k.dxp = g.dxp;
    } else if (!d.out) {
      return set_err(STG_ERR_INVALID, "block %d: null out", z);
    }
  }
  return STG_OK;
}
}  // namespace stg

using namespace stg;

extern "C" {

const char* stg_last_error(void) { return g_err; }
const char* stg_version(void) { return "stgconv_b200 0.1 sm_100a"; }

int stg_block_xmoments(const float* x_dev, int B, int T, int N, int C, double* xmom_dev, void* stream) {
  if (!x_dev || !xmom_dev || B < 1 || T < 1 || N < 1 || C < 1) return set_err(STG_ERR_INVALID, "bad argument");
  if (launch_xmoments(x_dev, B, T, N, C, xmom_dev, (cudaStream_t)stream))
    return set_err(STG_ERR_CUDA, "k_xmoments launch failed");
  return check_cuda("stg_block_xmoments");
}

int stg_block_forward(const float* x_dev, int B, int T, int N, int C, const stg_block_desc* blk, int nblk,
                      const double* xmom_dev, int training, float momentum, float eps, void* stream) {
  BlkArgs a;
  int rc = fill_args(a, x_dev, B, T, N, C, blk, nullptr, nblk, xmom_dev, training, momentum, eps);
  if (rc) return rc;
  BlkPlan p;
  char err[256];
  rc = plan_blocks(a, p, err, sizeof(err));
  if (rc) return set_err(rc == -1 ? STG_ERR_INVALID : STG_ERR_UNSUPPORTED, "%s", err);
  for (int z = 0; z < nblk; ++z)
    if (blk[z].out_bstride < (int64_t)a.b[z].L * N * a.b[z].H)
      return set_err(STG_ERR_INVALID, "block %d: out_bstride smaller than L*N*H", z);
  if (launch_block_forward(a, p, (cudaStream_t)stream)) return set_err(STG_ERR_CUDA, "block forward launch failed");
  return check_cuda("stg_block_forward");
}

int stg_block_backward(const float* x_dev, int B, int T, int N, int C, const stg_block_desc* blk,
                       const stg_block_grads* grads, int nblk, const double* xmom_dev, float eps, float* dx_dev,
                       void* stream) {
  if (!grads || !dx_dev) return set_err(STG_ERR_INVALID, "null grads / dx");
  BlkArgs a;
  int rc = fill_args(a, x_dev, B, T, N, C, blk, grads, nblk, xmom_dev, 1, 0.f, eps);
  if (rc) return rc;
  a.dx = dx_dev;
  BlkPlan p;
  char err[256];
  rc = plan_blocks(a, p, err, sizeof(err));
  if (rc) return set_err(rc == -1 ? STG_ERR_INVALID : STG_ERR_UNSUPPORTED, "%s", err);
  if (launch_block_backward(a, p, (cudaStream_t)stream)) return set_err(STG_ERR_CUDA, "block backward launch failed");
  return check_cuda("stg_block_backward");
}

}  // extern "C"
