// extern "C" surface declared in include/stgconv_b200.h -- whole-model entry points
// (FC_STGNN_RUL.forward / backward, fused MSE step, Adam).
#include <stdio.h>
#include <string.h>

#include "../../include/stgconv_b200.h"
#include "stg_block.cuh"
#include "stg_model.cuh"

namespace stg {
namespace {

struct Ws {            // byte offsets into the caller's workspace
  size_t h, dh, feat, dfeat, yp[2], dxp[2], z1, d1, c2raw, z3raw, dn2, dn1, dbl, xmom, bst[2], est, loss, cnt, dbl_end,
      total;
};
struct Geo {
  int C, J, F, L[2], M[2], R, EL2, NL1;
  size_t fsz[2];
};

size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

int geometry(const stg_model_dims& d, Geo& g) {
  if (d.B < 1 || d.N < 1 || d.T < 1 || d.P < 1 || d.K < 1 || d.EH < 1 || d.E < 1 || d.H < 1) return -1;
  g.C = 2 * d.H;
  g.J = 2 * d.H;
  g.R = d.B * d.T * d.N;
  {
    const int pad1 = d.K / 2, L1 = d.P + 2 * pad1 - d.K + 1, L2 = L1 + 2 - d.K + 1;
    if (L1 < 1 || L2 < 1) return -1;
    g.EL2 = d.E * L2;
    g.NL1 = d.EH * L1;
  }
  g.F = 0;
  for (int z = 0; z < STG_MAX_BLOCKS; ++z) {
    if (d.w[z] < 1 || d.stride[z] < 1 || d.T < d.w[z]) return -1;
    g.L[z] = (d.T - d.w[z]) / d.stride[z] + 1;
    g.M[z] = d.w[z] * d.N;
    g.fsz[z] = (size_t)g.L[z] * d.N * d.H;
    g.F += (int)g.fsz[z];
  }
  return 0;
}

void layout(const stg_model_dims& d, const Geo& g, Ws& w) {
  size_t o = 0;
  const size_t rc = al((size_t)g.R * g.C * 4);
  w.h = o; o += rc;
  w.dh = o; o += rc;
  w.feat = o; o += al((size_t)d.B * g.F * 4);
  w.dfeat = o; o += al((size_t)d.B * g.F * 4);
  for (int z = 0; z < 2; ++z) {
    w.yp[z] = o;
    o += al(STG_BLOCK_SAVED_FLOATS(d.B, d.T, d.N, d.H, d.w[z], d.stride[z]) * 4);
  }
  // dx partials: [B,T,N,C] (folded) or one row per (window, node) [B,L,w*N,C] on the tcgen05 path
  for (int z = 0; z < 2; ++z) {
    const size_t unf = al((size_t)d.B * g.L[z] * g.M[z] * g.C * 4);
    w.dxp[z] = o; o += unf > rc ? unf : rc;
  }
  w.d1 = o; o += al((size_t)d.B * g.J * 4);
  {
    const size_t tiles = ((size_t)g.R + 255) / 256;
    w.c2raw = o; o += al(tiles * g.EL2 * 256 * 4);
    w.z3raw = o; o += al(tiles * g.C * 256 * 4);
    w.dn2 = o; o += al(tiles * g.EL2 * 256 * 4);
    w.dn1 = o; o += al(tiles * g.NL1 * 256 * 4);
  }
  w.dbl = o;                                                   // ---- zeroed at the start of every forward
  w.z1 = o; o += al((size_t)d.B * g.J * 4);                    // fc1 accumulates split-K partial sums
  w.xmom = o; o += al((size_t)2 * d.T * g.C * 8);
  for (int z = 0; z < 2; ++z) { w.bst[z] = o; o += al((size_t)STG_BLOCK_STATS_DOUBLES(g.C, d.H, d.T) * 8); }
  w.est = o; o += al((size_t)enc_stats_doubles(d.EH, d.E, g.C) * 8);
  w.loss = o; o += al(16);
  w.cnt = o; o += al(16);
  w.dbl_end = o;
  w.total = o;
}

struct Ctx {
  Geo g;
  Ws w;
  EncArgs enc;
  size_t enc_smem_f, enc_smem_b;
  BlkArgs blk;
  BlkPlan plan;
  HeadArgs head;
};

int build_ctx(Ctx& c, const stg_model_dims* dp, const stg_model_params* pp, const stg_model_params* gp,
              const float* X, void* ws, size_t ws_bytes, int training, const stg_dropout* drop) {
  if (!dp || !pp || !X || !ws) return set_err(STG_ERR_INVALID, "null dims / params / X / workspace");
  const stg_model_dims& d = *dp;
  if (geometry(d, c.g)) return set_err(STG_ERR_INVALID, "invalid model dimensions");
  layout(d, c.g, c.w);
  if (ws_bytes < c.w.total)
    return set_err(STG_ERR_WORKSPACE, "workspace too small: %zu < %zu bytes", ws_bytes, c.w.total);
  if (((uintptr_t)ws & 255) != 0) return set_err(STG_ERR_INVALID, "workspace must be 256-byte aligned");
  char* base = (char*)ws;
  const Geo& g = c.g;
  const stg_model_params& p = *pp;
  char err[256];

  // ---- encoder
  EncArgs& e = c.enc;
  memset(&e, 0, sizeof(e));
  e.B = d.B; e.N = d.N; e.T = d.T; e.P = d.P; e.K = d.K; e.EH = d.EH; e.E = d.E; e.C = g.C;
  int rc = plan_encoder(e, &c.enc_smem_f, &c.enc_smem_b, err, sizeof(err));
  if (rc) return set_err(rc == -1 ? STG_ERR_INVALID : STG_ERR_UNSUPPORTED, "%s", err);
  e.X = X;
  e.W1 = p.conv1_w; e.W2 = p.conv2_w; e.W3 = p.lin_w; e.b3 = p.lin_b;
  e.g1 = p.bn1.weight; e.be1 = p.bn1.bias; e.rm1 = p.bn1.running_mean; e.rv1 = p.bn1.running_var;
  e.g2 = p.bn2.weight; e.be2 = p.bn2.bias; e.rm2 = p.bn2.running_mean; e.rv2 = p.bn2.running_var;
  e.g3 = p.bn3.weight; e.be3 = p.bn3.bias; e.rm3 = p.bn3.running_mean; e.rv3 = p.bn3.running_var;
  e.pe = p.pe;
  if (!e.W1 || !e.W2 || !e.W3 || !e.b3 || !e.g1 || !e.be1 || !e.rm1 || !e.rv1 || !e.g2 || !e.be2 || !e.rm2 ||
      !e.rv2 || !e.g3 || !e.be3 || !e.rm3 || !e.rv3 || !e.pe)
    return set_err(STG_ERR_INVALID, "null encoder parameter pointer");
  e.keep = drop ? drop->keep : nullptr;
  e.seed = drop ? drop->seed : 0ull;
  e.seed_ptr = drop ? (const long long*)drop->step_dev : nullptr;
  e.pdrop = d.pe_dropout;
  if (e.pdrop < 0.f || e.pdrop >= 1.f) return set_err(STG_ERR_INVALID, "pe_dropout must be in [0,1)");
  e.training = training; e.momentum = d.bn_momentum; e.eps = d.bn_eps;
  e.st = (double*)(base + c.w.est);
  e.h = (float*)(base + c.w.h);
  e.dh = (const float*)(base + c.w.dh);
  e.c2raw = (float*)(base + c.w.c2raw); e.z3raw = (float*)(base + c.w.z3raw);
  e.dn2 = (float*)(base + c.w.dn2); e.dn1 = (float*)(base + c.w.dn1);
  if (gp) {
    e.dW1 = gp->conv1_w; e.dW2 = gp->conv2_w; e.dW3 = gp->lin_w; e.db3 = gp->lin_b;
    e.dg1 = gp->bn1.weight; e.dbe1 = gp->bn1.bias; e.dg2 = gp->bn2.weight; e.dbe2 = gp->bn2.bias;
    e.dg3 = gp->bn3.weight; e.dbe3 = gp->bn3.bias;
    if (!e.dW1 || !e.dW2 || !e.dW3 || !e.db3 || !e.dg1 || !e.dbe1 || !e.dg2 || !e.dbe2 || !e.dg3 || !e.dbe3)
      return set_err(STG_ERR_INVALID, "null encoder gradient pointer");
  }

  // ---- graph-conv blocks (both read h, write straight into the concatenated feature rows)
  BlkArgs& a = c.blk;
  memset(&a, 0, sizeof(a));
  a.nblk = STG_MAX_BLOCKS; a.x = e.h; a.B = d.B; a.T = d.T; a.N = d.N; a.C = g.C;
  a.xmom = (const double*)(base + c.w.xmom);
  a.training = training; a.momentum = d.bn_momentum; a.eps = d.bn_eps;
  a.head_fused = training ? 1 : 0;
  a.dx = (float*)(base + c.w.dh);
  size_t foff = 0;
  for (int z = 0; z < STG_MAX_BLOCKS; ++z) {
    BlkDev& k = a.b[z];
    const stg_model_block& q = p.blk[z];
    k.H = d.H; k.w = d.w[z]; k.stride = d.stride[z]; k.decay = d.decay;
    k.Wm = q.Wm; k.bm = q.bm; k.g0 = q.bn0.weight; k.b0 = q.bn0.bias; k.rm0 = q.bn0.running_mean;
    k.rv0 = q.bn0.running_var; k.Wt = q.Wt; k.bt = q.bt; k.g1 = q.bn1.weight; k.b1 = q.bn1.bias;
    k.rm1 = q.bn1.running_mean; k.rv1 = q.bn1.running_var;
    if (!k.Wm || !k.bm || !k.g0 || !k.b0 || !k.rm0 || !k.rv0 || !k.Wt || !k.bt || !k.g1 || !k.b1 || !k.rm1 || !k.rv1)
      return set_err(STG_ERR_INVALID, "block %d: null parameter pointer", z);
    k.out = (float*)(base + c.w.feat) + foff;
    k.out_bs = g.F;
    k.yp = (float*)(base + c.w.yp[z]);
    k.stats = (double*)(base + c.w.bst[z]);
    k.coef = reinterpret_cast<float*>(k.stats + ((STG_BLOCK_SUMS_DOUBLES(g.C, d.H) + 1) / 2) * 2);
    k.dout = (const float*)(base + c.w.dfeat) + foff;
    k.dout_bs = g.F;
    k.dxp = (float*)(base + c.w.dxp[z]);
    if (gp) {
      const stg_model_block& gq = gp->blk[z];
      k.dWm = gq.Wm; k.dbm = gq.bm; k.dg0 = gq.bn0.weight; k.db0 = gq.bn0.bias;
      k.dWt = gq.Wt; k.dbt = gq.bt; k.dg1 = gq.bn1.weight; k.db1 = gq.bn1.bias;
      if (!k.dWm || !k.dbm || !k.dg0 || !k.db0 || !k.dWt || !k.dbt || !k.dg1 || !k.db1)
        return set_err(STG_ERR_INVALID, "block %d: null gradient pointer", z);
    }
    foff += g.fsz[z];
  }
  rc = plan_blocks(a, c.plan, err, sizeof(err));
  if (rc) return set_err(rc == -1 ? STG_ERR_INVALID : STG_ERR_UNSUPPORTED, "%s", err);
  // fast encoder: its first backward phase also finishes the blocks' backward (k_block_bwd_fin fused away)
  if (training && gp && encoder_fast_available(e)) {
    a.fin_elsewhere = 1;
    EncArgs::Fin& f = e.fin;
    f.nblk = STG_MAX_BLOCKS; f.CP = c.plan.CP; f.unfolded = a.dxp_unfolded;
    for (int z = 0; z < STG_MAX_BLOCKS; ++z) {
      const BlkDev& k = a.b[z];
      f.dxp[z] = k.dxp; f.tab[z] = k.coef; f.stats[z] = k.stats; f.g0[z] = k.g0;
      f.dg0[z] = k.dg0; f.db0[z] = k.db0; f.dg1[z] = k.dg1; f.db1[z] = k.db1;
      f.H[z] = k.H; f.w[z] = k.w; f.stride[z] = k.stride; f.L[z] = k.L;
    }
    f.dh_out = (float*)(base + c.w.dh);
  }

  // ---- head
  HeadArgs& hd = c.head;
  memset(&hd, 0, sizeof(hd));
  hd.B = d.B; hd.F = g.F; hd.J = g.J; hd.H = d.H;
  hd.feat = (const float*)(base + c.w.feat);
  hd.feat_out = (float*)(base + c.w.feat);
  hd.fused_blocks = training ? 1 : 0;
  hd.nblk = STG_MAX_BLOCKS;
  hd.momentum = d.bn_momentum; hd.eps = d.bn_eps;
  {
    int foff2 = 0;
    for (int z = 0; z < STG_MAX_BLOCKS; ++z) {
      HeadBlk& hb = hd.blk[z];
      const BlkDev& k = a.b[z];
      hb.yp = k.yp; hb.stats = k.stats; hb.g1 = k.g1; hb.b1 = k.b1; hb.rm1 = k.rm1; hb.rv1 = k.rv1;
      hb.L = g.L[z]; hb.M = g.M[z]; hb.H = d.H; hb.N = d.N; hb.w = d.w[z]; hb.foff = foff2;
      foff2 += (int)g.fsz[z];
    }
  }
  hd.W1 = p.fc_w[0]; hd.b1 = p.fc_b[0]; hd.W2 = p.fc_w[1]; hd.b2 = p.fc_b[1];
  hd.W3 = p.fc_w[2]; hd.b3 = p.fc_b[2]; hd.W4 = p.fc_w[3]; hd.b4 = p.fc_b[3];
  for (int i = 0; i < 4; ++i)
    if (!p.fc_w[i] || !p.fc_b[i]) return set_err(STG_ERR_INVALID, "null fc%d parameter pointer", i + 1);
  hd.z1 = (float*)(base + c.w.z1);
  hd.d1 = (float*)(base + c.w.d1);
  hd.loss = (float*)(base + c.w.loss);
  hd.dfeat = (float*)(base + c.w.dfeat);
  if (gp) {
    hd.dW1 = gp->fc_w[0]; hd.db1 = gp->fc_b[0]; hd.dW2 = gp->fc_w[1]; hd.db2 = gp->fc_b[1];
    hd.dW3 = gp->fc_w[2]; hd.db3 = gp->fc_b[2]; hd.dW4 = gp->fc_w[3]; hd.db4 = gp->fc_b[3];
    for (int i = 0; i < 4; ++i)
      if (!gp->fc_w[i] || !gp->fc_b[i]) return set_err(STG_ERR_INVALID, "null fc%d gradient pointer", i + 1);
  }
  if (g.J > 64) return set_err(STG_ERR_UNSUPPORTED, "hidden_dim %d > 32 unsupported by the head kernels", d.H);
  return STG_OK;
}

// encoder -> h, x-moments, both blocks -> feat, fc1 -> z1 [-> pred when with_tail]
int run_forward(Ctx& c, const stg_model_params& p, int training, float* pred, bool with_tail, float* zero1,
                cudaStream_t s) {
  char* base = (char*)c.enc.h - c.w.h;
  long long* nbt[8] = {(long long*)p.bn1.num_batches_tracked, (long long*)p.bn2.num_batches_tracked,
                       (long long*)p.bn3.num_batches_tracked, (long long*)p.blk[0].bn0.num_batches_tracked,
                       (long long*)p.blk[0].bn1.num_batches_tracked, (long long*)p.blk[1].bn0.num_batches_tracked,
                       (long long*)p.blk[1].bn1.num_batches_tracked, (long long*)c.enc.seed_ptr};
  // one launch: clear the reduction scratch, tick num_batches_tracked + dropout counter, clear the loss
  int rc = launch_zero(base + c.w.dbl, c.w.dbl_end - c.w.dbl, nbt, training ? 8 : 0, zero1, s);
  if (rc) return set_err(STG_ERR_CUDA, "k_zero launch failed");
  if (launch_encoder_forward(c.enc, c.enc_smem_f, s)) return set_err(STG_ERR_CUDA, "encoder forward launch failed");
  if (training) {
    // block stats are cleared by the step's k_zero; tables come with the x-moments (one launch)
    if (launch_xmoments_prep(c.blk, c.plan, (double*)(base + c.w.xmom), (unsigned*)(base + c.w.cnt), s))
      return set_err(STG_ERR_CUDA, "k_xmoments_prep launch failed");
    c.blk.prep_done = 1;
  }
  if (launch_block_forward(c.blk, c.plan, s)) return set_err(STG_ERR_CUDA, "block forward launch failed");
  HeadArgs hd = c.head;
  hd.pred = with_tail ? pred : nullptr;
  hd.y = nullptr; hd.dpred = nullptr;     // pred == nullptr: fc1 only, the backward tail follows
  if (launch_head_forward(hd, s)) return set_err(STG_ERR_CUDA, "head forward launch failed");
  return check_cuda("stg_model forward");
}

int run_backward(Ctx& c, const float* y, const float* dpred, float* pred, cudaStream_t s) {
  HeadArgs hd = c.head;
  hd.y = y; hd.dpred = dpred; hd.pred = pred;
  // all backward sums were cleared by the forward's first kernel: ONE backward per training forward
  if (launch_head_backward(hd, s)) return set_err(STG_ERR_CUDA, "head backward launch failed");
  if (launch_block_backward(c.blk, c.plan, s)) return set_err(STG_ERR_CUDA, "block backward launch failed");
  if (launch_encoder_backward(c.enc, c.enc_smem_b, s)) return set_err(STG_ERR_CUDA, "encoder backward launch failed");
  return check_cuda("stg_model backward");
}

}  // namespace
}  // namespace stg

using namespace stg;

extern "C" {

size_t stg_model_workspace_bytes(const stg_model_dims* dims) {
  if (!dims) return 0;
  Geo g;
  if (geometry(*dims, g)) return 0;
  Ws w;
  layout(*dims, g, w);
  return w.total;
}

int stg_model_forward(const stg_model_dims* dims, const stg_model_params* params, const float* X_dev,
                      void* workspace, size_t workspace_bytes, int training, const stg_dropout* drop,
                      float* pred_dev, void* stream) {
  if (!pred_dev) return set_err(STG_ERR_INVALID, "null pred");
  Ctx c;
  int rc = build_ctx(c, dims, params, nullptr, X_dev, workspace, workspace_bytes, training, drop);
  if (rc) return rc;
  return run_forward(c, *params, training, pred_dev, true, nullptr, (cudaStream_t)stream);
}

int stg_model_backward(const stg_model_dims* dims, const stg_model_params* params, const stg_model_params* grads,
                       const float* X_dev, void* workspace, size_t workspace_bytes, const stg_dropout* drop,
                       const float* dpred_dev, void* stream) {
  if (!grads || !dpred_dev) return set_err(STG_ERR_INVALID, "null grads / dpred");
  Ctx c;
  int rc = build_ctx(c, dims, params, grads, X_dev, workspace, workspace_bytes, 1, drop);
  if (rc) return rc;
  return run_backward(c, nullptr, dpred_dev, nullptr, (cudaStream_t)stream);
}

int stg_model_loss_backward(const stg_model_dims* dims, const stg_model_params* params,
                            const stg_model_params* grads, const float* X_dev, const float* y_dev,
                            void* workspace, size_t workspace_bytes, const stg_dropout* drop, float* pred_dev,
                            float* loss_dev, void* stream) {
  if (!grads || !y_dev || !loss_dev) return set_err(STG_ERR_INVALID, "null grads / y / loss");
  Ctx c;
  int rc = build_ctx(c, dims, params, grads, X_dev, workspace, workspace_bytes, 1, drop);
  if (rc) return rc;
  c.head.loss = loss_dev;
  cudaStream_t s = (cudaStream_t)stream;
  rc = run_forward(c, *params, 1, nullptr, false, loss_dev, s);
  if (rc) return rc;
  return run_backward(c, y_dev, nullptr, pred_dev, s);
}

int stg_adam_step(float* param_dev, const float* grad_dev, float* exp_avg_dev, float* exp_avg_sq_dev, int64_t n,
                  int64_t* step_dev, float lr, float beta1, float beta2, float eps, float weight_decay,
                  float grad_scale, void* stream) {
  if (!param_dev || !grad_dev || !exp_avg_dev || !exp_avg_sq_dev || !step_dev || n < 0)
    return set_err(STG_ERR_INVALID, "bad argument");
  if (n == 0) return STG_OK;
  launch_adam(param_dev, grad_dev, exp_avg_dev, exp_avg_sq_dev, (long long)n, (long long*)step_dev, lr, beta1, beta2,
              eps, weight_decay, grad_scale, (cudaStream_t)stream);
  return check_cuda("stg_adam_step");
}

}  // extern "C"
