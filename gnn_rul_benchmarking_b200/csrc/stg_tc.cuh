// Graph-conv block on the Blackwell tensor cores (tcgen05.mma, accumulators in TMEM) -- the default path of
// GraphConvpoolMPNN_block_v6 (reference models/FC_STGNN/Model_Base.py:175-225) for C <= 16, H <= 8, w = 2,
// w*N <= 64 (every C-MAPSS / N-CMAPSS hyper-parameter set with hidden_dim 8 and the BASELINE synthetic shape).
//
// Work unit = a TILE of 128 graph rows = WPT windows x WR rows (WR = 32 or 64 >= M = w*N, pad rows are zero).
// One CTA = 128 threads = one thread per row (TMEM lane == row).  Windows are independent graphs, so a tile
// may mix samples; CTAs walk the tiles of one block with a static stride (persistent, 2 CTAs per SM).
//
// Every contraction of the block runs as single-pass TF32 tcgen05.mma (kind::tf32, M = 128, cta_group::1),
// issued by thread 0, completion signalled through tcgen05.commit -> mbarrier:
//   forward   FV  = x . [Wm | Wtheta.diag(g0 r0)]^T          (A, B K-major in shared memory)
//             S   = F . F^T                                   (per-window diagonal blocks are used)
//             Z_w = A_w . V_w                                 (A = softmax rows written back to TMEM in place)
//   backward  the same FV, S, plus
//             dA  = dY' . V^T
//             dF_w = dS_w . F_w + dS_w^T . F_w ,  dV_w = A_w^T . dY'_w
//             dxp_w = [dF_w | dV_w] . [Wm ; a0 Wtheta]        (A straight from the TMEM accumulators)
//             G    = [dF | dV]^T . [x | 1]                    (parameter-gradient outer products of the tile)
// Shared-memory operand formats (validated on hardware by scripts/umma_probe.cu):
//   K-major, no swizzle   : chunk layout X4[k/4][row][4 floats]; descriptor LBO = bytes between k-chunks,
//                           SBO = 128 (8 rows x 16 B core matrices)
//   MN-major (transposed) : tf32 only exists as SWIZZLE_128B_BASE32B: one 128-byte row per k holding 32
//                           consecutive mn values, 32-byte chunk index ^= (k & 3); LBO = bytes between 32-wide
//                           mn blocks, SBO = 512 (4 rows)
// The thread owning a row does the row softmax and its backward entirely in registers (no shuffles), reading
// its row of S / dA with tcgen05.ld.
#pragma once
#include "stg_block.cuh"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>

namespace stg {
namespace tc {

constexpr int kCP = 16, kHP = 8, kCPH = 24;
constexpr float kLog2e = 1.4426950408889634f;

// ---- tcgen05 wrappers ---------------------------------------------------------------------------------
STG_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
STG_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
STG_DEVINL void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

STG_DEVINL uint64_t sdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = (uint64_t)(layout & 7u) << 61;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// kind::tf32, fp32 accumulate, M = 128
__host__ __device__ constexpr uint32_t idesc_tf32(int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
STG_DEVINL void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
STG_DEVINL void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
STG_DEVINL void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
STG_DEVINL void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
STG_DEVINL void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
STG_DEVINL void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
STG_DEVINL void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
               "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
               "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
STG_DEVINL void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

STG_DEVINL float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// round-to-nearest TF32 of an operand value before it is stored for the tensor core (the MMA itself truncates)
STG_DEVINL float rtf(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
STG_DEVINL float4 rtf4(float a, float b, float c, float d) { return make_float4(rtf(a), rtf(b), rtf(c), rtf(d)); }

// ---- shared-memory carve-up ------------------------------------------------------------------------------
// all operand buffers are 1024-byte aligned (the MN-major swizzle works on absolute address bits 5..8)
struct SmemLayout {
  int xk;      // [4][128][4]   x rows, K-major; re-used for F rows (K-major) after the projection MMA
  int xkl;     // [4][128][4]   SPLIT: tf32 residuals (x - tf32(x), then F - tf32(F)) of the same rows
  int vs;      // forward: V rows [128][8] fp32 (broadcast reads of the SIMT aggregation)
  int wcbl;    // SPLIT: residual of the projection weights
  int ra;      // [128][32]     row records, MN-major: F at columns 0..15, dY' at 16..23 (backward); V at 0..7 (forward)
  int rb;      // [128][32]     backward: x at columns 0..15, 1.0 at column 16
  int yk;      // [2][2][128][4] backward: dY' rows and V rows, K-major
  int t1;      // dS^T operand (MN-major), later the [dF | dV] records
  int t2;      // A^T operand (MN-major)
  int wcb;     // [4][32][4]    projection weights  B[n = o][k = c], K-major
  int wc2;     // [6][16][4]    [Wm ; a0 Wtheta]    B[n = c][k = o], K-major
  int wc2l;    // SPLIT: its residual
  int pf_fv, pf_p;   // backward prefetch staging: the tile's saved F|V blocks and softmax blocks
  int cst;     // constants (floats)
  int total;
};
// constants region (float offsets)
constexpr int kCstBias = 0;        // [32]
constexpr int kCstBn1 = 32;        // [7][8]
constexpr int kCstBt = 88;         // [8]
constexpr int kCstA0 = 96;         // a0[16] c0[16] mu0[16] r0[16]
constexpr int kCstMisc = 160;      // decay
constexpr int kCstRed = 176;       // backward: G[24][17], dbt[8], per-warp partials [4][32]; forward: partials [4][16]
constexpr int kCstFloats = 176 + 24 * 17 + 8 + 128 + 8;

// pfM > 0 (backward only): stage the tile's saved F|V and softmax blocks (w*N = pfM rows per window) in shared memory
// with 1-D TMA bulk copies issued one tile ahead.
__host__ __device__ inline SmemLayout make_layout(int WR, bool bwd, bool split, int pfM = 0) {
  SmemLayout l;
  int o = 0;
  // forward: x / F rows (K-major, + residuals), V rows, projection weights.  backward: row records, dY' | V rows,
  // the two transposed operands (t2 doubles as the K-major [dF | dV] rows + residuals afterwards), dx weights.
  l.xk = o; if (!bwd) o += 4 * 128 * 16;
  l.xkl = o; if (!bwd && split) o += 4 * 128 * 16;
  l.vs = o; if (!bwd) o += 128 * 32;
  l.ra = o; if (bwd) o += 128 * 128;
  l.rb = o; if (bwd) o += 128 * 128;
  l.yk = o; if (bwd) o += 2 * 2 * 128 * 16;
  l.t1 = o; if (bwd) o += 4 * WR * 128;
  l.t2 = o; if (bwd) o += (4 * WR * 128 > 24 * 1024 ? 4 * WR * 128 : 24 * 1024);      // later: K-major rows 12 KB + residuals 12 KB
  l.pf_fv = o; if (bwd && pfM) o += (128 / WR) * kCPH * pfM * 4;
  l.pf_p = o; if (bwd && pfM) o += ((128 / WR) * (pfM + 1) * pfM * 4 + 15) / 16 * 16;
  l.wcb = o; if (!bwd) o += 4 * 32 * 16;
  l.wcbl = o; if (!bwd && split) o += 4 * 32 * 16;
  l.wc2 = o; if (bwd) o += 6 * 16 * 16;
  l.wc2l = o; if (bwd && split) o += 6 * 16 * 16;
  o = (o + 1023) / 1024 * 1024;
  l.cst = o; o += kCstFloats * 4;
  l.total = o + 1024;      // slack for the manual 1024-byte alignment of the dynamic window
  return l;
}

// The training forward saves, behind Y' [B,L,M,H] in the same caller-provided buffer (STG_BLOCK_SAVED_FLOATS), per
// window g and with the row index i of the window as the fastest dimension (the thread owning row i reads / writes
// element [.][i]: every access of a warp is one contiguous run):
//   FV block [24][M]      F = mapped features (0..15), V = BN0(x).Wtheta^T (16..23) of the window's rows
//   P  block [M + 1][M]   un-normalised softmax numerators e[i][k] stored at [k][i], the sign of S[i][k] in the sign
//                         bit (it carries leaky_relu'(S)); row M holds 1 / sum_k e[i][k]
// so that the backward neither repeats the projection / Gram products nor the exponentials.
__host__ __device__ inline size_t saved_off_fv(long long rows, int H) { return (size_t)((rows * H + 3) / 4 * 4); }
__host__ __device__ inline size_t saved_off_p(long long rows, int H) { return saved_off_fv(rows, H) + (size_t)rows * 24; }
__host__ __device__ inline int saved_mp(int M) { return M + 1; }

struct TcCtl {
  uint64_t bar;
  uint64_t bar_pf;       // backward: completion of the prefetched blocks (TMA bulk copies)
  uint64_t bar_w;        // backward: per-window products, one committing thread per window
  uint64_t bar_x;        // backward: dx projection and parameter-gradient product, issued by two threads
  uint32_t tmem_base;
  uint32_t pad;
};

// Prologue shared by forward and backward: projection weights, biases, BN0 coefficients.
// Training: copied from the coefficient table written by k_block_prep / k_xmoments_prep; eval: from running stats.
STG_DEVINL void tc_prologue(const BlkArgs& a, const BlkDev& k, unsigned char* sm, const SmemLayout& L, bool split,
                            bool bwd) {
  float* cst = reinterpret_cast<float*>(sm + L.cst);
  float* wcb = reinterpret_cast<float*>(sm + L.wcb);
  float* wcbl = reinterpret_cast<float*>(sm + L.wcbl);
  float* wc2l = reinterpret_cast<float*>(sm + L.wc2l);
  float* wc2 = reinterpret_cast<float*>(sm + L.wc2);
  const int tid = threadIdx.x, C = a.C, H = k.H;
  if (a.training) {
    const float* tab = k.coef;     // mu0[16] r0[16] a0[16] c0[16] biasc[24] pw[4] WcT[16*24] cnt[T]
    if (tid < 32) cst[kCstBias + tid] = tid < kCPH ? tab[4 * kCP + tid] : 0.f;
    if (tid < 16) {
      cst[kCstA0 + tid] = tab[2 * kCP + tid];
      cst[kCstA0 + 16 + tid] = tab[3 * kCP + tid];
      cst[kCstA0 + 32 + tid] = tab[tid];
      cst[kCstA0 + 48 + tid] = tab[kCP + tid];
    }
    const float* WcT = tab + 4 * kCP + kCPH + 4;
    for (int idx = tid; idx < 32 * 16; idx += 128) {
      const int o = idx >> 4, c = idx & 15;
      const float v = o < kCPH ? WcT[c * kCPH + o] : 0.f;
      const float vh = rtf(v);
      if (!bwd) {
        wcb[((c >> 2) * 32 + o) * 4 + (c & 3)] = vh;
        if (split) wcbl[((c >> 2) * 32 + o) * 4 + (c & 3)] = rtf(v - vh);
      } else if (o < kCPH) {
        wc2[((o >> 2) * 16 + c) * 4 + (o & 3)] = vh;
        if (split) wc2l[((o >> 2) * 16 + c) * 4 + (o & 3)] = rtf(v - vh);
      }
    }
  } else {
    // eval: a0 = g0 / sqrt(rv0 + eps), c0 = b0 - a0 * rm0
    if (tid < 16) {
      float mean = 0.f, r = 0.f, av = 0.f, cv = 0.f;
      if (tid < C) {
        mean = k.rm0[tid];
        r = (float)(1.0 / sqrt((double)k.rv0[tid] + (double)a.eps));
        av = k.g0[tid] * r;
        cv = k.b0[tid] - av * mean;
      }
      cst[kCstA0 + tid] = av; cst[kCstA0 + 16 + tid] = cv; cst[kCstA0 + 32 + tid] = mean; cst[kCstA0 + 48 + tid] = r;
    }
    __syncthreads();
    for (int idx = tid; idx < 32 * 16; idx += 128) {
      const int o = idx >> 4, c = idx & 15;
      float v = 0.f;
      if (c < C) {
        if (o < C) v = k.Wm[o * C + c];
        else if (o >= kCP && o - kCP < H) v = k.Wt[(o - kCP) * C + c] * cst[kCstA0 + c];
      }
      const float vh = rtf(v);
      if (!bwd) {
        wcb[((c >> 2) * 32 + o) * 4 + (c & 3)] = vh;
        if (split) wcbl[((c >> 2) * 32 + o) * 4 + (c & 3)] = rtf(v - vh);
      } else if (o < kCPH) {
        wc2[((o >> 2) * 16 + c) * 4 + (o & 3)] = vh;
        if (split) wc2l[((o >> 2) * 16 + c) * 4 + (o & 3)] = rtf(v - vh);
      }
    }
    if (tid < 32) {
      float v = 0.f;
      if (tid < C) v = k.bm[tid];
      else if (tid >= kCP && tid - kCP < H) {
        const float* wr = k.Wt + (tid - kCP) * C;
        for (int c = 0; c < C; ++c) v += wr[c] * cst[kCstA0 + 16 + c];
      }
      cst[kCstBias + tid] = v;
    }
  }
  if (tid < 8) cst[kCstBt + tid] = tid < H ? k.bt[tid] : 0.f;
  if (tid == 0) cst[kCstMisc] = k.decay;
}

// store 4 values as tf32 operands.  Single pass: rounded to nearest.  SPLIT: hi = the value truncated to tf32
// (exactly what the tensor core would read), lo = value - hi (exact in fp32; the tensor core truncates it to its
// top 11 bits, i.e. the pair carries ~21 bits) -- 2 instructions per value instead of 5.
STG_DEVINL float ttf(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
template <bool SPLIT>
STG_DEVINL void st_hl(float4* hi, float4* lo, float a, float b, float c, float d) {
  if (SPLIT) {
    const float4 h = make_float4(ttf(a), ttf(b), ttf(c), ttf(d));
    *hi = h;
    *lo = make_float4(a - h.x, b - h.y, c - h.z, d - h.w);
  } else {
    *hi = rtf4(a, b, c, d);
  }
}

// packed fp32 pair FMA (sm_100 FFMA2): d = a * b + d on both halves
STG_DEVINL void ffma2(float2& d, const float2 a, const float2 b) {
  unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(dd)
      : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
  d = *reinterpret_cast<float2*>(&dd);
}

// projection FV = x . Wc^T and Gram S = F . F^T with the operands in xk (+ residuals in xkl):
// single pass, or the 3-term error-compensated product  a.b ~= a_hi.b_hi + a_hi.b_lo + a_lo.b_hi
template <bool SPLIT>
STG_DEVINL void issue_proj(uint32_t d_tmem, uint32_t xk_u, uint32_t xkl_u, uint32_t wcb_u, uint32_t wcbl_u) {
  constexpr uint32_t id = idesc_tf32(32, 0, 0);
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
    mma_ss(d_tmem, sdesc(xk_u + ks * 4096, 2048, 128, 0), sdesc(wcb_u + ks * 1024, 512, 128, 0), id, ks);
  if (SPLIT) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      mma_ss(d_tmem, sdesc(xk_u + ks * 4096, 2048, 128, 0), sdesc(wcbl_u + ks * 1024, 512, 128, 0), id, 1);
      mma_ss(d_tmem, sdesc(xkl_u + ks * 4096, 2048, 128, 0), sdesc(wcb_u + ks * 1024, 512, 128, 0), id, 1);
    }
  }
}
template <bool SPLIT>
STG_DEVINL void issue_gram(uint32_t d_tmem, uint32_t xk_u, uint32_t xkl_u) {
  constexpr uint32_t id = idesc_tf32(128, 0, 0);
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
    mma_ss(d_tmem, sdesc(xk_u + ks * 4096, 2048, 128, 0), sdesc(xk_u + ks * 4096, 2048, 128, 0), id, ks);
  if (SPLIT) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      mma_ss(d_tmem, sdesc(xk_u + ks * 4096, 2048, 128, 0), sdesc(xkl_u + ks * 4096, 2048, 128, 0), id, 1);
      mma_ss(d_tmem, sdesc(xkl_u + ks * 4096, 2048, 128, 0), sdesc(xk_u + ks * 4096, 2048, 128, 0), id, 1);
    }
  }
}

template <int CN>
STG_DEVINL void load_row(const float* __restrict__ p, int n, bool vec, float (&v)[CN]) {
  if (vec) {
#pragma unroll
    for (int q = 0; q < CN / 4; ++q) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p) + q);
      v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int c = 0; c < CN; ++c) v[c] = c < n ? __ldg(p + c) : 0.f;
  }
}

// One row of a [128 rows][128 B] MN-major record buffer: 16-byte piece `q16` (0..7) of the row, swizzled.
STG_DEVINL float4* rec_ptr(unsigned char* buf, int row, int q16) {
  const int chunk32 = (q16 >> 1) ^ (row & 3);
  return reinterpret_cast<float4*>(buf + row * 128 + chunk32 * 32 + (q16 & 1) * 16);
}


// ---- descriptors with everything but the start address folded into constants -------------------------------
//   lo word = (addr >> 4) | (LBO >> 4) << 16      hi word = (SBO >> 4) | version 1 << 14 | layout << 29
constexpr uint32_t kHiK = (128u >> 4) | (1u << 14);                      // K-major, no swizzle, SBO 128
constexpr uint32_t kHiMN = (512u >> 4) | (1u << 14) | (1u << 29);        // MN-major, SWIZZLE_128B_BASE32B, SBO 512
STG_DEVINL uint32_t dlo(uint32_t saddr, uint32_t lbo) { return (saddr >> 4) | ((lbo >> 4) << 16); }
STG_DEVINL uint64_t dsc(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

template <bool SPLIT>
STG_DEVINL void issue_proj2(uint32_t d_tmem, uint32_t xk_lo, uint32_t xkl_lo, uint32_t wcb_lo, uint32_t wcbl_lo) {
  constexpr uint32_t id = idesc_tf32(32, 0, 0);
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
    mma_ss(d_tmem, dsc(xk_lo + ks * 256, kHiK), dsc(wcb_lo + ks * 64, kHiK), id, ks);
  if (SPLIT) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      mma_ss(d_tmem, dsc(xk_lo + ks * 256, kHiK), dsc(wcbl_lo + ks * 64, kHiK), id, 1);
      mma_ss(d_tmem, dsc(xkl_lo + ks * 256, kHiK), dsc(wcb_lo + ks * 64, kHiK), id, 1);
    }
  }
}
template <bool SPLIT>
STG_DEVINL void issue_gram2(uint32_t d_tmem, uint32_t xk_lo, uint32_t xkl_lo) {
  constexpr uint32_t id = idesc_tf32(128, 0, 0);
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
    mma_ss(d_tmem, dsc(xk_lo + ks * 256, kHiK), dsc(xk_lo + ks * 256, kHiK), id, ks);
  if (SPLIT) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      mma_ss(d_tmem, dsc(xk_lo + ks * 256, kHiK), dsc(xkl_lo + ks * 256, kHiK), id, 1);
      mma_ss(d_tmem, dsc(xkl_lo + ks * 256, kHiK), dsc(xk_lo + ks * 256, kHiK), id, 1);
    }
  }
}

// CTAs per block: proportional to the block's tiles, ctas_total in all
inline void split_ctas(const BlkArgs& a, int WR, int ctas_total, int* n0, int* total) {
  const int WPT = 128 / WR;
  long long t[2] = {0, 0};
  for (int z = 0; z < a.nblk; ++z) t[z] = ((long long)a.B * a.b[z].L + WPT - 1) / WPT;
  if (a.nblk == 1) {
    *n0 = (int)(t[0] < ctas_total ? t[0] : ctas_total);
    *total = *n0;
    return;
  }
  long long c0 = (t[0] * ctas_total + (t[0] + t[1]) / 2) / (t[0] + t[1]);
  if (c0 < 1) c0 = 1;
  if (c0 > ctas_total - 1) c0 = ctas_total - 1;
  long long c1 = ctas_total - c0;
  if (c0 > t[0]) c0 = t[0];
  if (c1 > t[1]) c1 = t[1];
  *n0 = (int)c0;
  *total = (int)(c0 + c1);
}
inline int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
  }
  return n;
}

#ifdef STG_TC_TIMING
// debug build only (-DSTG_TC_TIMING): clock64 stamps of CTA 0's second tile, threads 0 and 64
static __device__ long long g_tc_stamp[2][16];          // per translation unit: [thread 0 | 64][stamp]
static __device__ unsigned long long g_cta_time[1024][3];  // globaltimer at CTA start / end, SM id
#define STG_STAMP(n)                                                                              \
  if (blockIdx.x == 0 && tile == cta + ncta && (tid == 0 || tid == 64)) g_tc_stamp[tid == 64][n] = clock64();
#else
#define STG_STAMP(n)
#endif

}  // namespace tc
}  // namespace stg
