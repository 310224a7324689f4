/* stgconv_b200.h -- C ABI of the B200 (sm_100a) spatio-temporal graph-conv hot path.
 *
 * Drop-in boundary for the FC_STGNN path of Frank-Wang-oss/GNN_RUL_Benchmarking.  The
 * reference is pure Python/PyTorch and has no FFI of its own; the "interface each entry
 * point replaces" is therefore the Python call the reference makes at that place
 * (file:line relative to the reference tree).  INTEGRATION.md shows the ctypes binding a
 * maintainer would add.
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch / C++ types.
 *   - all tensors are contiguous float32 unless stated; "dev" pointers are device memory of
 *     the CUDA device current on the calling thread, "host" pointers are host memory
 *     (pinned memory makes the copies asynchronous).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Device
 *     entry points only enqueue work; they never synchronise.
 *   - every function returns STG_OK (0) or a negative stg_status; stg_last_error() gives the
 *     message of the last failure on the calling thread.  (The reference raises Python
 *     exceptions; the Python host layer turns a non-zero status into RuntimeError/ValueError.)
 */
#ifndef STGCONV_B200_H
#define STGCONV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum stg_status {
  STG_OK = 0,
  STG_ERR_INVALID = -1,      /* bad shape / null pointer / inconsistent sizes   */
  STG_ERR_UNSUPPORTED = -2,  /* shape outside what the kernels are built for    */
  STG_ERR_CUDA = -3,         /* a CUDA runtime call failed                      */
  STG_ERR_WORKSPACE = -4     /* caller-provided workspace too small             */
} stg_status;

const char* stg_last_error(void);
/* library / build identification, e.g. "stgconv_b200 0.1 sm_100a". */
const char* stg_version(void);

/* ------------------------------------------------------------------------------------------
 * Graph-conv block  ==  GraphConvpoolMPNN_block_v6   (models/FC_STGNN/Model_Base.py:175-225)
 *   x [B,T,N,C] -> out [B,L,N,H],  L = (T-w)/stride + 1, pool_choice = 'mean'
 *   = Conv_GraphST unfold (:137-148) -> Dot_Graph_Construction_weights (:44-67) -> * Mask_Matrix
 *     (:150-170,203) -> BatchNorm1d(C) (:183,206-208) -> MPNN_mk_v2 k=1 (:72-107) -> mean over w.
 * Up to STG_MAX_BLOCKS blocks that read the SAME x (FC_STGNN_RUL uses two: Model.py:26-27,74-75)
 * run in one launch sequence.
 * ---------------------------------------------------------------------------------------- */
#define STG_MAX_BLOCKS 2

typedef struct stg_block_desc {
  /* hyper-parameters (Model_Base.py:176-188) */
  int32_t H;          /* output_dim                                  */
  int32_t w;          /* time_window_size                            */
  int32_t stride;     /* stride                                      */
  float decay;        /* Mask_Matrix decay                           */
  /* parameters, device pointers (state_dict names in comments) */
  const float* Wm;    /* graph_construction.mapping.weight [C,C]     */
  const float* bm;    /* graph_construction.mapping.bias   [C]       */
  const float* bn0_w; /* BN.weight [C]                               */
  const float* bn0_b; /* BN.bias   [C]                               */
  float* bn0_rm;      /* BN.running_mean [C]  (updated when training)*/
  float* bn0_rv;      /* BN.running_var  [C]                         */
  const float* Wt;    /* MPNN.theta.0.weight [H,C]                   */
  const float* bt;    /* MPNN.theta.0.bias   [H]                     */
  const float* bn1_w; /* MPNN.bn1.weight [H]                         */
  const float* bn1_b; /* MPNN.bn1.bias   [H]                         */
  float* bn1_rm;      /* MPNN.bn1.running_mean [H]                   */
  float* bn1_rv;      /* MPNN.bn1.running_var  [H]                   */
  /* outputs / saved-for-backward, device pointers */
  float* out;         /* [B, L, N, H] with sample stride out_bstride (floats)              */
  int64_t out_bstride;/* >= L*N*H; lets two blocks write straight into the FC-head input   */
  float* yp;          /* training only: STG_BLOCK_SAVED_FLOATS floats saved for the backward:
                         pre-BN Y' [B,L,w*N,H], then (tcgen05 path) the F | V rows [B*L*w*N, 24]
                         and the softmax numerators [B*L, w*N + 1, w*N]                     */
  double* stats;      /* training only: STG_BLOCK_STATS_DOUBLES(C,H,T) doubles of scratch  */
} stg_block_desc;

#define STG_BLOCK_WINDOWS(T, w, stride) (((T) - (w)) / (stride) + 1)
#define STG_BLOCK_SAVED_FLOATS(B, T, N, H, w, stride)                                            \
  ((((size_t)(B) * STG_BLOCK_WINDOWS(T, w, stride) * (w) * (N) * (H) + 3) / 4) * 4 +             \
   (size_t)(B) * STG_BLOCK_WINDOWS(T, w, stride) * (w) * (N) * (24 + (w) * (N) + 1))

/* layout of stg_block_desc.stats (doubles):
 *   [0,H)        sum_R  Y'            [H,2H)      sum_R Y'^2          (forward)
 *   [2H,3H)      sum_R  dYn           [3H,4H)     sum_R dYn*Yhat      (backward)
 *   [4H,4H+C)    sum_R  dXhat         [4H+C,4H+2C) sum_R dXhat*Xhat   (backward)
 *   then (16-byte aligned) the float coefficient table.                                  */
#define STG_BLOCK_SUMS_DOUBLES(C, H) (4 * (H) + 2 * (C))
/* ... followed by a float table of per-feature coefficients shared by the kernels of one launch
 * sequence (BN0 mean / rstd, folded projection weights, cover counts), sized for the padded dims. */
#define STG_PAD16(x) ((((x) + 15) / 16) * 16)
#define STG_PAD8(x) ((((x) + 7) / 8) * 8)
#define STG_BLOCK_COEF_FLOATS(C, H, T) \
  (4 * STG_PAD16(C) + (STG_PAD16(C) + STG_PAD8(H)) + 4 + STG_PAD16(C) * (STG_PAD16(C) + STG_PAD8(H)) + (T))
#define STG_BLOCK_STATS_DOUBLES(C, H, T) \
  (((STG_BLOCK_SUMS_DOUBLES(C, H) + 1) / 2) * 2 + (STG_BLOCK_COEF_FLOATS(C, H, T) + 1) / 2)

#define STG_BLOCK_DXP_FLOATS(B, T, N, C, w, stride)                                              \
  ((size_t)(B) * (size_t)(C) *                                                                   \
   ((size_t)(T) * (N) > (size_t)(((T) - (w)) / (stride) + 1) * (w) * (N)                         \
        ? (size_t)(T) * (N)                                                                      \
        : (size_t)(((T) - (w)) / (stride) + 1) * (w) * (N)))

/* gradients of one block, device pointers; all are ACCUMULATED into (+=) except dxp. */
typedef struct stg_block_grads {
  const float* dout;  /* [B,L,N,H], sample stride dout_bstride                              */
  int64_t dout_bstride;
  float* dWm; float* dbm; float* dbn0_w; float* dbn0_b;
  float* dWt; float* dbt; float* dbn1_w; float* dbn1_b;
  float* dxp;         /* scratch, STG_BLOCK_DXP_FLOATS floats: this block's dx before the BN-0 mean
                         terms, [B,T,N,C] or one row per (window, node) [B,L,w*N,C] (tcgen05 path)  */
} stg_block_grads;

/* Per-time-step moments of x for the BatchNorm1d(C) of every block reading x:
 *   xmom[t*C+c] = sum_{b,n} x[b,t,n,c],  xmom[T*C + t*C+c] = sum x^2     (2*T*C doubles, zeroed here)
 * (training only; replaces the batch-statistics half of Model_Base.py:206-208). */
int stg_block_xmoments(const float* x_dev, int B, int T, int N, int C, double* xmom_dev, void* stream);

/* Forward of nblk blocks on the same x.  training=0 uses running statistics (model.eval(),
 * trainer.py:155) and needs neither xmom nor yp/stats.  training=1 uses batch statistics,
 * updates running_mean/var with `momentum` (unbiased variance) and fills yp/stats. */
int stg_block_forward(const float* x_dev, int B, int T, int N, int C, const stg_block_desc* blk, int nblk,
                      const double* xmom_dev, int training, float momentum, float eps, void* stream);

/* Backward of the same (training mode).  Writes dx [B,T,N,C] (overwritten, summed over the
 * blocks) and accumulates the parameter gradients. */
int stg_block_backward(const float* x_dev, int B, int T, int N, int C, const stg_block_desc* blk,
                       const stg_block_grads* grads, int nblk, const double* xmom_dev, float eps,
                       float* dx_dev, void* stream);

/* ------------------------------------------------------------------------------------------
 * Whole model  ==  FC_STGNN_RUL  (models/FC_STGNN/Model.py:5-85) and one optimisation step of
 * class FC_STGNN (algorithms/algorithms.py:51-76).
 *
 *   X [B, N, T*P]  --reshape/transpose (Model.py:45-49)-->  rows (b,t,n) of P samples
 *     -> Feature_extractor_1DCNN_RUL (Model_Base.py:12-41): Conv1d(1->EH,K,pad K/2,no bias)+BN+ReLU,
 *        Conv1d(EH->E,K,pad 1,no bias)+BN+ReLU                     -> [rows, E, L2]
 *     -> nonlin_map2 (Model.py:18-22): Linear(E*L2 -> C=2H) + BatchNorm1d(C)
 *     -> PositionalEncoding (Model_Base.py:111-134): + pe[t], Dropout(p)      -> h [B,T,N,C]
 *     -> MPNN1 (w=2,stride 1), MPNN2 (w=2,stride 2) on the same h (Model.py:74-75)
 *     -> cat(flatten) (Model.py:78-81)                                        -> feat [B,F]
 *     -> fc: Linear(F->C)+ReLU, Linear(C->C)+ReLU, Linear(C->H)+ReLU, Linear(H->1)  (Model.py:30-39)
 * ---------------------------------------------------------------------------------------- */
typedef struct stg_model_dims {
  int32_t B;      /* windows in this batch                                   */
  int32_t N;      /* num_node                                                */
  int32_t T;      /* num_patch                                               */
  int32_t P;      /* patch_size                                              */
  int32_t K;      /* encoder_conv_kernel                                     */
  int32_t EH;     /* encoder_hidden_dim                                      */
  int32_t E;      /* encoder_out_dim                                         */
  int32_t H;      /* hidden_dim  (C = 2*H)                                   */
  int32_t w[STG_MAX_BLOCKS];       /* moving_window  (Model.py:13)           */
  int32_t stride[STG_MAX_BLOCKS];  /* stride         (Model.py:14)           */
  float decay;        /* Model.py:12                                         */
  float pe_dropout;   /* Model.py:24 (0.1)                                   */
  float bn_momentum;  /* 0.1                                                 */
  float bn_eps;       /* 1e-5                                                */
} stg_model_dims;

typedef struct stg_bn {           /* one nn.BatchNorm1d                                   */
  float* weight; float* bias; float* running_mean; float* running_var;
  int64_t* num_batches_tracked;   /* incremented by training forwards (may be NULL)       */
} stg_bn;

typedef struct stg_model_block {  /* one GraphConvpoolMPNN_block_v6 (MPNN1 / MPNN2)         */
  float* Wm; float* bm;   /* MPNNk.graph_construction.mapping.{weight [C,C],bias [C]} */
  stg_bn bn0;             /* MPNNk.BN.*             [C]                               */
  float* Wt; float* bt;   /* MPNNk.MPNN.theta.0.{weight [H,C],bias [H]}               */
  stg_bn bn1;             /* MPNNk.MPNN.bn1.*       [H]                               */
} stg_model_block;

typedef struct stg_model_params { /* device pointers, state_dict names in comments        */
  float* conv1_w;  /* nonlin_map.conv_block1.0.weight [EH,1,K]    */
  stg_bn bn1;      /* nonlin_map.conv_block1.1.*      [EH]        */
  float* conv2_w;  /* nonlin_map.conv_block2.0.weight [E,EH,K]    */
  stg_bn bn2;      /* nonlin_map.conv_block2.1.*      [E]         */
  float* lin_w;    /* nonlin_map2.0.weight            [C, E*L2]   */
  float* lin_b;    /* nonlin_map2.0.bias              [C]         */
  stg_bn bn3;      /* nonlin_map2.1.*                 [C]         */
  const float* pe; /* positional_encoding.pe[0]       [>=T, C]    */
  stg_model_block blk[STG_MAX_BLOCKS];
  float* fc_w[4];  /* fc.fc1..fc4.weight  [C,F] [C,C] [H,C] [1,H] */
  float* fc_b[4];  /* fc.fc1..fc4.bias                            */
} stg_model_params;
/* Gradients use the same struct: every weight/bias pointer addresses the gradient buffer of that
 * parameter (ACCUMULATED into, +=); running_* / num_batches_tracked / pe are ignored. */

/* Dropout of the positional encoding in training mode:
 *   keep != NULL : 0/1 float mask laid out like the reference's dropout input [B*N, T, C]
 *                  (Model.py:64-65) -- used to pin the mask in parity tests;
 *   keep == NULL : counter-based generator keyed by (seed, element index); the same (seed) must be
 *                  passed to the backward call.  pe_dropout == 0 disables it.
 *   step_dev     : optional device counter mixed into the seed; training forwards increment it on the
 *                  device, so a captured CUDA graph draws a fresh mask on every replay. */
typedef struct stg_dropout { const float* keep; uint64_t seed; int64_t* step_dev; } stg_dropout;

/* Bytes of caller-provided device workspace for a batch of dims->B windows (activations saved for
 * backward + reduction scratch).  0 on invalid dims. */
size_t stg_model_workspace_bytes(const stg_model_dims* dims);

/* pred[B] = FC_STGNN_RUL.forward(X).  training=1: batch statistics, running stats and
 * num_batches_tracked updated, activations kept in `workspace` for stg_model_backward. */
int stg_model_forward(const stg_model_dims* dims, const stg_model_params* params, const float* X_dev,
                      void* workspace, size_t workspace_bytes, int training, const stg_dropout* drop,
                      float* pred_dev, void* stream);

/* Gradients of every parameter given dpred[B] = dLoss/dpred, after a training forward on the same
 * workspace / X / dropout. */
int stg_model_backward(const stg_model_dims* dims, const stg_model_params* params, const stg_model_params* grads,
                       const float* X_dev, void* workspace, size_t workspace_bytes, const stg_dropout* drop,
                       const float* dpred_dev, void* stream);

/* forward -> nn.MSELoss (mean) -> backward in one call (algorithms.py:68-73 without the optimizer):
 * loss_dev[0] = mean((pred - y)^2), gradients accumulated into `grads`. */
int stg_model_loss_backward(const stg_model_dims* dims, const stg_model_params* params,
                            const stg_model_params* grads, const float* X_dev, const float* y_dev,
                            void* workspace, size_t workspace_bytes, const stg_dropout* drop, float* pred_dev,
                            float* loss_dev, void* stream);

/* torch.optim.Adam(lr, betas, eps, weight_decay) (algorithms.py:60-64) over ONE flat parameter
 * buffer: grad += wd*param; m,v EMA; bias correction with step = ++(*step_dev).  n floats. */
int stg_adam_step(float* param_dev, const float* grad_dev, float* exp_avg_dev, float* exp_avg_sq_dev, int64_t n,
                  int64_t* step_dev, float lr, float beta1, float beta2, float eps, float weight_decay,
                  float grad_scale, void* stream);

/* ------------------------------------------------------------------------------------------
 * Dense per-graph adjacency builders of the sibling models (primitives A2-A4):
 * x [G, N, F] -> adj [G, N, N], forward and backward (dx given dadj).
 *   STG_ADJ_PCC     pcc_graph_construction        models/ST_GCN/Model.py:53-71 (also ST_Conv, LOGO, DVGTformer)
 *   STG_ADJ_COSINE  cosine_distance               models/HAGCN/Model.py:122-127, models/SAGCN/Model.py:74-79
 *   STG_ADJ_GAUSS   exp(-cdist(x, x))             models/ASTGCNN/Model.py:193-194 (its Linear P stays a GEMM)
 *   STG_ADJ_GAUSS2  exp(-cdist^2), top_k per row  models/STGNN/Model.py:8-25 (top_k <= 0 or >= N: dense)
 * mask_dev (optional, GAUSS2): 0/1 bytes of the kept entries.  The backward takes the forward's adj
 * (masked entries are zero and carry no gradient). */
enum { STG_ADJ_PCC = 0, STG_ADJ_COSINE = 1, STG_ADJ_GAUSS = 2, STG_ADJ_GAUSS2 = 3,
       STG_ADJ_GRAM = 4 /* x x^T, models/STMSGCN/Model.py:96 */ };
int stg_adj_forward(int kind, const float* x_dev, int64_t G, int N, int F, int top_k, float* adj_dev,
                    unsigned char* mask_dev, void* stream);
int stg_adj_backward(int kind, const float* x_dev, const float* adj_dev, const float* dadj_dev, int64_t G, int N,
                     int F, float* dx_dev, void* stream);

/* Dense graph aggregation of the sibling models (primitives M2 / M3); the learnable
 * projections that follow are plain GEMMs and stay with the caller.
 *   STG_AGG_GCN    out [G,N,F]   = D^-1/2 (A+I) D^-1/2 X       models/SAGCN/Model.py:81-95, STMSGCN:34-49, RGCNU:7-21
 *   STG_AGG_CHEB3  out [G,3,N,F] = [X, A X, 2 A (A X) - X]     models/ASTGCNN/Model.py:212-228, STGNN:43-59
 * backward: dout (same shape as out) -> dx [G,N,F], dadj [G,N,N]. */
enum { STG_AGG_GCN = 0, STG_AGG_CHEB3 = 1, STG_AGG_AX = 2 /* out = A X: MPNN_mk k=1, models/ST_GCN/Model.py:80-90 */ };
int stg_agg_forward(int kind, const float* x_dev, const float* adj_dev, int64_t G, int N, int F, float* out_dev,
                    void* stream);
int stg_agg_backward(int kind, const float* x_dev, const float* adj_dev, const float* dout_dev, int64_t G, int N,
                     int F, float* dx_dev, float* dadj_dev, void* stream);

/* TemporalConvNet of the sibling models (primitive T1): two causal convolutions (dilation 1
 * and 2, no bias, Chomp1d) each followed by BatchNorm1d + ReLU, with residual ReLUs
 * (models/ASTGCNN/Model.py:72-146 kernel 6; models/ST_GCN/Model.py:99-173 kernel 2; ST_Conv, STAGNN).
 * x / out [B, C, L]; C_in == C_out (downsample0/1 are None in every reference configuration).
 * scratch_dev: 8*C doubles, written by the training forward and reused by the backward of the same batch. */
typedef struct stg_tcn_params {
  float* conv1_w;   /* conv_block1.0.weight [C,C,K] */
  stg_bn bn1;       /* conv_block1.2.*      [C]     */
  float* conv2_w;   /* conv_block2.0.weight [C,C,K] */
  stg_bn bn2;       /* conv_block2.2.*      [C]     */
} stg_tcn_params;
/* saved_dev (optional, 2*B*C*L floats): the training forward keeps the raw outputs of both convolutions there and the
 * later phases / the backward of the same batch read them instead of recomputing the chain (null: recompute). */
int stg_tcn_forward(const float* x_dev, int B, int C, int L, int K, const stg_tcn_params* params, int training,
                    float momentum, float eps, double* scratch_dev, float* saved_dev, float* out_dev, void* stream);
/* grads: same struct, weight / bias pointers address gradient buffers (ACCUMULATED into). */
int stg_tcn_backward(const float* x_dev, const float* dout_dev, int B, int C, int L, int K,
                     const stg_tcn_params* params, const stg_tcn_params* grads, float eps, double* scratch_dev,
                     const float* saved_dev, float* dx_dev, void* stream);

/* Per-patch statistics of the bearing models' parameter-free prefix: segment_and_compute_features
 * (models/ST_GCN/Model.py:7-52).  x [R, P] -> out [R, 10] = max, min, ptp, var, std (unbiased), mean, rms,
 * mean|x|, skewness, excess kurtosis.  Forward only (no parameter upstream). */
int stg_patch_stats(const float* x_dev, int64_t R, int P, float* out_dev, void* stream);
/* extract_features (models/GAT_LSTM/Model.py:6-70): x [R, P] -> out [R, 11] = mean, std (unbiased),
 * (mean sqrt|x|)^2, rms, (max-min)/2, skewness, kurtosis (the reference's sample-size coefficients), crest,
 * clearance, shape and impulse factors.  Forward only. */
int stg_patch_stats11(const float* x_dev, int64_t R, int P, float* out_dev, void* stream);
/* extract_temporal_features (models/SAGCN/Model.py:21-38): x [R, P] -> out [R, 12] = max, min, std, rms, mean,
 * ptp, var, softmax entropy, std(asin(clamp x)), std(atan x), excess kurtosis, skewness.  Forward only. */
int stg_patch_stats12(const float* x_dev, int64_t R, int P, float* out_dev, void* stream);

/* Dense graph attention (primitive M5): GraphAttentionLayer.forward after its nn.Linear
 * (models/GAT_LSTM/Model.py:87-109).  Wh [G,N,F]; att_w [2F], att_b [1] = GraphAttentionLayer.attention;
 * adj [N,N] shared by all graphs (adj_per_graph = 0) or [G,N,N]; keep = 0/1 dropout mask [G,N,N] of the
 * attention matrix (NULL: no dropout), scaled by 1/(1-pdrop) inside; alpha = slope of the score leaky_relu,
 * out_slope (> 0) = slope of the output leaky_relu.
 *   out = leaky_relu_{out_slope}((dropout(softmax_j(leaky_relu_alpha(a1.Wh_i + a2.Wh_j + b))) * adj) Wh)
 * backward: dWh written, datt_w [2F] / datt_b [1] ACCUMULATED into. */
int stg_gat_forward(const float* Wh_dev, const float* att_w_dev, const float* att_b_dev, const float* adj_dev,
                    int adj_per_graph, const float* keep_dev, float pdrop, float alpha, float out_slope,
                    int G, int N, int F, float* out_dev, void* stream);
int stg_gat_backward(const float* Wh_dev, const float* att_w_dev, const float* att_b_dev, const float* adj_dev,
                     int adj_per_graph, const float* keep_dev, float pdrop, float alpha, float out_slope,
                     int G, int N, int F, const float* out_dev, const float* dout_dev, float* dWh_dev,
                     float* datt_w_dev, float* datt_b_dev, void* stream);

/* Recurrent layer of the sibling models (primitive T2): the time recurrence of ONE nn.LSTM / nn.GRU
 * layer with zero initial state, one or two directions (models/HAGCN/Model.py:33-73, models/GAT_LSTM/Model.py:129-132,
 * models/STGNN/Model.py:72,99, models/STMSGCN/Model.py:52-60).  The input projection is a plain GEMM and stays with the
 * caller:  xg = x . W_ih^T + b_ih + b_hh  (GRU: the n-gate third gets b_in only, b_hn is passed separately because it
 * sits inside r * (W_hn h + b_hn)).  Gate order as in torch: LSTM i,f,g,o; GRU r,z,n.
 *   xg   element (b, t, dir, row) at b*xg_bstride + t*xg_tstride + dir*G*H + row      G = 4 (LSTM) / 3 (GRU)
 *   out  element (b, t, dir, j)   at b*out_bstride + t*out_tstride + dir*H + j        dir 1 runs t = T-1 .. 0
 *   whh  [ndir][G*H][H] (weight_hh_l0, weight_hh_l0_reverse), bhn [ndir][H] (GRU only)
 *   saved  stg_rnn_saved_floats() floats written by a training forward, read by the backward (null: inference)
 * backward: dout (layout of out) -> dxg = d loss / d xg (layout of xg) and, GRU only, dhn = d loss / d (W_hn h + b_hn)
 * (layout of out).  dW_hh = sum_t dgates_t (x) h_{t-1} (GRU: n rows from dhn) and db_hn = sum dhn are a GEMM / a
 * reduction over those outputs and are left to the caller.  H <= 128; 8 sequences per CTA; H > 64 runs on clusters of 4. */
enum { STG_RNN_LSTM = 0, STG_RNN_GRU = 1 };
int stg_rnn_batch_tile(int B);
size_t stg_rnn_saved_floats(int cell, int T, int B, int H, int ndir);
int stg_rnn_forward(int cell, const float* xg_dev, int64_t xg_bstride, int64_t xg_tstride, const float* whh_dev,
                    const float* bhn_dev, int T, int B, int H, int ndir, float* out_dev, int64_t out_bstride,
                    int64_t out_tstride, float* saved_dev, void* stream);
int stg_rnn_backward(int cell, const float* whh_dev, const float* saved_dev, const float* dout_dev,
                     int64_t out_bstride, int64_t out_tstride, int T, int B, int H, int ndir, float* dxg_dev,
                     int64_t xg_bstride, int64_t xg_tstride, float* dhn_dev, void* stream);

/* Evaluation metrics (utils.py:136-169, called every epoch from trainer.py:119-121): ACCUMULATES into
 * out4_dev (4 doubles, caller zeroes): [0] sum of Score_v1 terms, [1] sum of Score_v2 terms,
 * [2] sum |pred-real|, [3] sum (pred-real)^2.  Score_v2 average, MAE and RMSE follow as
 * out[1]/n, out[2]/n*max_rul, sqrt(out[3]/n)*max_rul. */
int stg_metrics(const float* pred_dev, const float* real_dev, int64_t n, float max_rul, double* out4_dev, void* stream);

/* Data-parallel step (the reference has no distributed code): one-shot
 * all-reduce of the flat gradient buffers over NVLink peer memory fused with the Adam update.
 * grad_ptrs[r] / flag_ptrs[r] (host arrays of `world` device pointers) address rank r's gradient
 * buffer (n floats) and flag block (>= 64 uint32, all zero except word 34 = 1, the epoch of the flag
 * protocol, which the kernel itself advances and which must never be rewound) and must be peer-mapped
 * on this device (e.g. torch symmetric memory).  Every rank must call it once per step with the same n;
 * the kernel waits (bounded, ~2 s) for all peers.  flag word 33 != 0 afterwards means a peer timed out;
 * the rank that saw the timeout skipped its parameter update for that step. */
int stg_allreduce_adam(float* param_dev, float* exp_avg_dev, float* exp_avg_sq_dev, int64_t n, int64_t* step_dev,
                       const float* const* grad_ptrs, uint32_t* const* flag_ptrs, int rank, int world, float lr,
                       float beta1, float beta2, float eps, float weight_decay, void* stream);

/* ------------------------------------------------------------------------------------------
 * Per-kernel timing (measurement aid, no reference counterpart: the reference has no profiler,
 * SURVEY.md section 5).  When enabled, every kernel launch of this library is bracketed by a
 * cudaEvent pair on the launching stream; stg_profile_read() synchronises on the recorded events
 * of one kernel slot and returns their summed duration and the number of launches since the
 * last stg_profile_reset().
 * ---------------------------------------------------------------------------------------- */
int stg_profile_enable(int on);
int stg_profile_reset(void);
int stg_profile_slots(void);
const char* stg_profile_name(int slot);
int stg_profile_read(int slot, double* total_ms, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* STGCONV_B200_H */
