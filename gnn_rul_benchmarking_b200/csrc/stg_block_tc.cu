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
#include "stg_block.cuh"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>

namespace stg {

namespace {

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
  int akl;     // forward SPLIT: residual of the softmax rows, K-major [WR/4][128][4]
  int wcbl;    // SPLIT: residual of the projection weights
  int ra;      // [128][32]     row records, MN-major: F at columns 0..15, dY' at 16..23 (backward); V at 0..7 (forward)
  int rb;      // [128][32]     backward: x at columns 0..15, 1.0 at column 16
  int yk;      // [2][2][128][4] backward: dY' rows and V rows, K-major
  int t1;      // dS^T operand (MN-major), later the [dF | dV] records
  int t2;      // A^T operand (MN-major)
  int wcb;     // [4][32][4]    projection weights  B[n = o][k = c], K-major
  int wc2;     // [6][16][4]    [Wm ; a0 Wtheta]    B[n = c][k = o], K-major
  int wc2l;    // SPLIT: its residual
  int cst;     // constants (floats)
  int total;
};
// constants region (float offsets)
constexpr int kCstBias = 0;        // [32]
constexpr int kCstBn1 = 32;        // [7][8]
constexpr int kCstBt = 88;         // [8]
constexpr int kCstA0 = 96;         // a0[16] c0[16] mu0[16] r0[16]
constexpr int kCstMisc = 160;      // decay
constexpr int kCstRed = 176;       // reductions at CTA end: up to 24*17 + 16 floats
constexpr int kCstFloats = 176 + 24 * 17 + 32;

__host__ __device__ inline SmemLayout make_layout(int WR, bool bwd, bool split) {
  SmemLayout l;
  int o = 0;
  l.xk = o; o += 4 * 128 * 16;
  l.xkl = o; if (split) o += 4 * 128 * 16;
  l.akl = o; if (split && !bwd) o += (WR / 4) * 128 * 16;
  l.ra = o; o += 128 * 128;
  l.rb = o; if (bwd) o += 128 * 128;
  l.yk = o; if (bwd) o += 2 * 2 * 128 * 16;
  l.t1 = o; if (bwd) o += 4 * WR * 128;
  l.t2 = o; if (bwd) o += 4 * WR * 128;
  l.wcb = o; o += 4 * 32 * 16;
  l.wcbl = o; if (split) o += 4 * 32 * 16;
  l.wc2 = o; o += 6 * 16 * 16;
  l.wc2l = o; if (split) o += 6 * 16 * 16;
  o = (o + 1023) / 1024 * 1024;
  l.cst = o; o += kCstFloats * 4;
  l.total = o + 1024;      // slack for the manual 1024-byte alignment of the dynamic window
  return l;
}

struct TcCtl {
  uint64_t bar;
  uint32_t tmem_base;
  uint32_t pad;
};

// Prologue shared by forward and backward: projection weights, biases, BN0 coefficients.
// Training: copied from the coefficient table written by k_block_prep / k_xmoments_prep; eval: from running stats.
STG_DEVINL void tc_prologue(const BlkArgs& a, const BlkDev& k, unsigned char* sm, const SmemLayout& L, bool split) {
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
      wcb[((c >> 2) * 32 + o) * 4 + (c & 3)] = vh;
      if (split) wcbl[((c >> 2) * 32 + o) * 4 + (c & 3)] = rtf(v - vh);
      if (o < kCPH) {
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
      wcb[((c >> 2) * 32 + o) * 4 + (c & 3)] = vh;
      if (split) wcbl[((c >> 2) * 32 + o) * 4 + (c & 3)] = rtf(v - vh);
      if (o < kCPH) {
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

// store 4 values as a tf32 float4 (and, SPLIT, their tf32 residuals)
template <bool SPLIT>
STG_DEVINL void st_hl(float4* hi, float4* lo, float a, float b, float c, float d) {
  const float4 h = rtf4(a, b, c, d);
  *hi = h;
  if (SPLIT) *lo = rtf4(a - h.x, b - h.y, c - h.z, d - h.w);
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

// ------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------
template <int WR, bool TRAIN, bool SPLIT>
__global__ void __launch_bounds__(128, 2) k_block_fwd_tc(const BlkArgs a, int ncta0) {
  constexpr int WPT = 128 / WR;
  extern __shared__ unsigned char smraw[];
  __shared__ TcCtl ctl;
  unsigned char* sm = reinterpret_cast<unsigned char*>(((uintptr_t)smraw + 1023) & ~(uintptr_t)1023);
  const SmemLayout L = make_layout(WR, false, SPLIT);
  const int z = (int)blockIdx.x < ncta0 ? 0 : 1;
  const BlkDev& k = a.b[z];
  const int cta = z == 0 ? blockIdx.x : blockIdx.x - ncta0;
  const int ncta = z == 0 ? ncta0 : gridDim.x - ncta0;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int N = a.N, C = a.C, T = a.T, H = k.H, s = k.stride, Lw = k.L, M = 2 * N;
  const long long nwin = (long long)a.B * Lw;
  const int ntiles = (int)((nwin + WPT - 1) / WPT);

  if (tid == 0) mbar_init(&ctl.bar, 1);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl.tmem_base)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_prologue(a, k, sm, L, SPLIT);
  float* cst = reinterpret_cast<float*>(sm + L.cst);
  if (!TRAIN && tid < 8) {
    float a1 = 0.f, c1 = 0.f;
    if (tid < H) {
      const float r1 = (float)(1.0 / sqrt((double)k.rv1[tid] + (double)a.eps));
      a1 = k.g1[tid] * r1;
      c1 = k.b1[tid] - a1 * k.rm1[tid];
    }
    cst[kCstBn1 + tid] = a1;
    cst[kCstBn1 + 8 + tid] = c1;
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ctl.tmem_base;
  const uint32_t lane_t = tmem + ((uint32_t)(warp * 32) << 16);      // this warp's 32 TMEM lanes
  const uint32_t xk_u = smem_u32(sm + L.xk), ra_u = smem_u32(sm + L.ra), wcb_u = smem_u32(sm + L.wcb),
                 xkl_u = smem_u32(sm + L.xkl), wcbl_u = smem_u32(sm + L.wcbl), akl_u = smem_u32(sm + L.akl);
  float4* xk4 = reinterpret_cast<float4*>(sm + L.xk);
  float4* xkl4 = reinterpret_cast<float4*>(sm + L.xkl);
  float4* akl4 = reinterpret_cast<float4*>(sm + L.akl);
  unsigned char* ra = sm + L.ra;

  const int wl = tid / WR, i = tid - wl * WR;       // window slot inside the tile, row inside the window
  const bool row_ok = i < M;
  const bool j1 = i >= N;                           // time offset of this row inside its window
  const float decay = cst[kCstMisc];
  const bool xvec = (C == 16);
  const bool yvec = (H == 8);
  float st1[kHP], st2[kHP];
#pragma unroll
  for (int h = 0; h < kHP; ++h) st1[h] = st2[h] = 0.f;
  uint32_t ph = 0;

  for (int tile = cta; tile < ntiles; tile += ncta) {
    const long long g = (long long)tile * WPT + wl;
    const bool valid = row_ok && g < nwin;
    const int b = valid ? (int)(g / Lw) : 0;
    const int l = valid ? (int)(g - (long long)b * Lw) : 0;
    // ---- x rows -> K-major operand
    {
      float xr[16];
      if (valid) load_row<16>(a.x + (((size_t)b * T + (size_t)l * s) * N + i) * C, C, xvec, xr);
      else {
#pragma unroll
        for (int c = 0; c < 16; ++c) xr[c] = 0.f;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        st_hl<SPLIT>(&xk4[q * 128 + tid], &xkl4[q * 128 + tid], xr[4 * q], xr[4 * q + 1], xr[4 * q + 2], xr[4 * q + 3]);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      issue_proj<SPLIT>(tmem, xk_u, xkl_u, wcb_u, wcbl_u);
      mma_commit(&ctl.bar);
    }
    mbar_wait(&ctl.bar, ph); ph ^= 1;
    tc_fence_after();
    // ---- F, V rows (V record: tf32 values at columns 0..7, SPLIT residuals at 8..15)
    {
      float fv[32];
      tmem_ld32(lane_t, fv);
#pragma unroll
      for (int c = 0; c < kCPH; ++c) fv[c] += cst[kCstBias + c];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        st_hl<SPLIT>(&xk4[q * 128 + tid], &xkl4[q * 128 + tid], fv[4 * q], fv[4 * q + 1], fv[4 * q + 2], fv[4 * q + 3]);
      st_hl<SPLIT>(rec_ptr(ra, tid, 0), rec_ptr(ra, tid, 2), fv[16], fv[17], fv[18], fv[19]);
      st_hl<SPLIT>(rec_ptr(ra, tid, 1), rec_ptr(ra, tid, 3), fv[20], fv[21], fv[22], fv[23]);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      issue_gram<SPLIT>(tmem, xk_u, xkl_u);
      mma_commit(&ctl.bar);
    }
    mbar_wait(&ctl.bar, ph); ph ^= 1;
    tc_fence_after();
    // ---- row softmax, A = (P + I) o mask written back over S
    {
      float sv[WR];
#pragma unroll
      for (int q = 0; q < WR / 32; ++q) tmem_ld32(lane_t + wl * WR + q * 32, *reinterpret_cast<float(*)[32]>(&sv[q * 32]));
      float mx = -INFINITY;
#pragma unroll
      for (int kk = 0; kk < WR; ++kk) {
        const float lam = fmaxf(sv[kk], kLeaky * sv[kk]);
        sv[kk] = lam;
        if (kk < M && kk != i) mx = fmaxf(mx, lam);
      }
      const float mxl = mx * kLog2e;
      float sum = 0.f;
#pragma unroll
      for (int kk = 0; kk < WR; ++kk) {
        float e = ex2(fmaf(sv[kk], kLog2e, -mxl));
        if (kk >= M || kk == i) e = 0.f;
        sv[kk] = e;
        sum += e;
      }
      const float inv = (valid && sum > 0.f) ? 1.f / sum : 0.f;
#pragma unroll
      for (int q = 0; q < WR / 8; ++q) {
        float o8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int kk = q * 8 + u;
          const float mk = ((kk >= N) == j1) ? 1.f : decay;
          float p = sv[kk] * inv;
          if (kk == i && valid) p = 1.f;
          p *= mk;
          o8[u] = rtf(p);
          sv[kk] = p - o8[u];                 // residual (SPLIT)
        }
        tmem_st8(lane_t + wl * WR + q * 8, o8);
        if (SPLIT) {
          akl4[(2 * q) * 128 + tid] = rtf4(sv[q * 8], sv[q * 8 + 1], sv[q * 8 + 2], sv[q * 8 + 3]);
          akl4[(2 * q + 1) * 128 + tid] = rtf4(sv[q * 8 + 4], sv[q * 8 + 5], sv[q * 8 + 6], sv[q * 8 + 7]);
        }
      }
      tmem_wait_st();
    }
    if (SPLIT) fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      constexpr uint32_t id = idesc_tf32(16, 0, 1);
#pragma unroll
      for (int w2 = 0; w2 < WPT; ++w2)
#pragma unroll
        for (int ks = 0; ks < WR / 8; ++ks) {
          const uint32_t brow = ra_u + (w2 * WR + ks * 8) * 128;
          mma_ts(tmem + 128 + 16 * w2, tmem + w2 * WR + ks * 8, sdesc(brow, 512, 512, 1), id, ks);
          if (SPLIT) {
            mma_ts(tmem + 128 + 16 * w2, tmem + w2 * WR + ks * 8, sdesc(brow + 32, 512, 512, 1), id, 1);
            mma_ss(tmem + 128 + 16 * w2, sdesc(akl_u + ks * 4096, 2048, 128, 0), sdesc(brow, 512, 512, 1), id, 1);
          }
        }
      mma_commit(&ctl.bar);
    }
    mbar_wait(&ctl.bar, ph); ph ^= 1;
    tc_fence_after();
    // ---- Y' = A.V + btheta
    {
      float y[8];
      tmem_ld8(lane_t + 128 + 16 * wl, y);
#pragma unroll
      for (int h = 0; h < kHP; ++h) y[h] += cst[kCstBt + h];
      if (TRAIN) {
        if (valid) {
          float* yrow = k.yp + ((size_t)g * M + i) * H;
          if (yvec) {
            reinterpret_cast<float4*>(yrow)[0] = make_float4(y[0], y[1], y[2], y[3]);
            reinterpret_cast<float4*>(yrow)[1] = make_float4(y[4], y[5], y[6], y[7]);
          } else {
#pragma unroll
            for (int h = 0; h < kHP; ++h)
              if (h < H) yrow[h] = y[h];
          }
#pragma unroll
          for (int h = 0; h < kHP; ++h) {
            st1[h] += y[h];
            st2[h] = fmaf(y[h], y[h], st2[h]);
          }
        }
      } else {
        // BN1 (running statistics) + leaky_relu, then the mean over the two rows of every sensor (through smem)
        float* ex = reinterpret_cast<float*>(sm + L.xk);      // K-major F rows are consumed; reuse as [128][8]
        __syncthreads();
#pragma unroll
        for (int h = 0; h < kHP; ++h) ex[tid * 8 + h] = lrelu(fmaf(cst[kCstBn1 + h], y[h], cst[kCstBn1 + 8 + h]));
        __syncthreads();
        if (valid && !j1) {
          float* orow = k.out + (size_t)b * k.out_bs + ((size_t)l * N + i) * H;
          const float* p0 = ex + tid * 8;
          const float* p1 = ex + (tid + N) * 8;
          for (int h = 0; h < H; ++h) orow[h] = 0.5f * (p0[h] + p1[h]);
        }
        __syncthreads();
      }
    }
    tc_fence_before();
  }

  if (TRAIN) {
    float* red = cst + kCstRed;
    __syncthreads();
    if (tid < 2 * kHP) red[tid] = 0.f;
    __syncthreads();
#pragma unroll
    for (int h = 0; h < kHP; ++h) {
      const float v1 = warp_sum(st1[h]), v2 = warp_sum(st2[h]);
      if ((tid & 31) == 0) {
        atomicAdd(&red[h], v1);
        atomicAdd(&red[kHP + h], v2);
      }
    }
    __syncthreads();
    if (tid < H) {
      atomicAdd(&k.stats[tid], (double)red[tid]);
      atomicAdd(&k.stats[H + tid], (double)red[kHP + tid]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

// ------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------
template <int WR, bool SPLIT>
__global__ void __launch_bounds__(128, 2) k_block_bwd_tc(const BlkArgs a, int ncta0) {
  constexpr int WPT = 128 / WR;
  extern __shared__ unsigned char smraw[];
  __shared__ TcCtl ctl;
  unsigned char* sm = reinterpret_cast<unsigned char*>(((uintptr_t)smraw + 1023) & ~(uintptr_t)1023);
  const SmemLayout L = make_layout(WR, true, SPLIT);
  const int z = (int)blockIdx.x < ncta0 ? 0 : 1;
  const BlkDev& k = a.b[z];
  const int cta = z == 0 ? blockIdx.x : blockIdx.x - ncta0;
  const int ncta = z == 0 ? ncta0 : gridDim.x - ncta0;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int N = a.N, C = a.C, T = a.T, H = k.H, s = k.stride, Lw = k.L, M = 2 * N;
  const long long nwin = (long long)a.B * Lw;
  const int ntiles = (int)((nwin + WPT - 1) / WPT);

  if (tid == 0) mbar_init(&ctl.bar, 1);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl.tmem_base)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_prologue(a, k, sm, L, SPLIT);
  float* cst = reinterpret_cast<float*>(sm + L.cst);
  // BN1 backward coefficients: [0]=a1 [1]=c1 [2]=mu1 [3]=r1 [4]=g1*r1 [5]=q1 [6]=q2
  if (tid < 8) {
    float v[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (tid < H) {
      const int h = tid;
      const double R = (double)a.B * Lw * M;
      const double m = k.stats[h] / R;
      double var = k.stats[H + h] / R - m * m;
      if (var < 0.0) var = 0.0;
      const float r1 = (float)(1.0 / sqrt(var + (double)a.eps));
      const float g1 = k.g1[h];
      v[0] = g1 * r1;
      v[1] = k.b1[h] - v[0] * (float)m;
      v[2] = (float)m;
      v[3] = r1;
      v[4] = g1 * r1;
      v[5] = (float)(g1 * k.stats[2 * H + h] / R) * r1;
      v[6] = (float)(g1 * k.stats[3 * H + h] / R) * r1;
    }
#pragma unroll
    for (int q = 0; q < 7; ++q) cst[kCstBn1 + q * 8 + tid] = v[q];
  }
  // operand buffers whose pad parts are read by the tensor core but never written per tile
  for (int idx = tid; idx < (L.wcb - L.ra) / 16; idx += 128) reinterpret_cast<float4*>(sm + L.ra)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ctl.tmem_base;
  const uint32_t lane_t = tmem + ((uint32_t)(warp * 32) << 16);
  const uint32_t regA = 0, regB = 128;
  const uint32_t xk_u = smem_u32(sm + L.xk), ra_u = smem_u32(sm + L.ra), rb_u = smem_u32(sm + L.rb),
                 yk_u = smem_u32(sm + L.yk), t1_u = smem_u32(sm + L.t1), t2_u = smem_u32(sm + L.t2),
                 wcb_u = smem_u32(sm + L.wcb), wc2_u = smem_u32(sm + L.wc2), xkl_u = smem_u32(sm + L.xkl),
                 wcbl_u = smem_u32(sm + L.wcbl);
  float4* xk4 = reinterpret_cast<float4*>(sm + L.xk);
  float4* xkl4 = reinterpret_cast<float4*>(sm + L.xkl);
  float4* yk4 = reinterpret_cast<float4*>(sm + L.yk);
  unsigned char* ra = sm + L.ra;
  unsigned char* rb = sm + L.rb;
  unsigned char* t1 = sm + L.t1;
  unsigned char* t2 = sm + L.t2;

  const int wl = tid / WR, i = tid - wl * WR;
  const bool row_ok = i < M;
  const bool j1 = i >= N;
  const int n_i = j1 ? i - N : i;
  const float decay = cst[kCstMisc];
  const bool xvec = (C == 16);
  const bool yvec = (H == 8);
  float dbt_acc[kHP];
#pragma unroll
  for (int h = 0; h < kHP; ++h) dbt_acc[h] = 0.f;
  float gacc[16];          // warp 0, lane o < 24: G[o][c] = sum_rows [dF | dV][o] * x[c]
#pragma unroll
  for (int q = 0; q < 16; ++q) gacc[q] = 0.f;
  // column sums of [dF | dV] in plain fp32 (they feed the bias / BatchNorm-shift gradients, which are small
  // differences of large sums): dF rows are added as they come out of TMEM; sum_rows dV = sum_i rowsum(A_i) dY'_i
  float sof[kCP], sov[kHP];
#pragma unroll
  for (int c = 0; c < kCP; ++c) sof[c] = 0.f;
#pragma unroll
  for (int h = 0; h < kHP; ++h) sov[h] = 0.f;
  const uint32_t wc2l_u = smem_u32(sm + L.wc2l);
  float4* t2k4 = reinterpret_cast<float4*>(t2);       // [dF | dV] rows, K-major, for the dx projection
  uint32_t ph = 0;

  for (int tile = cta; tile < ntiles; tile += ncta) {
    const long long g = (long long)tile * WPT + wl;
    const bool valid = row_ok && g < nwin;
    const int b = valid ? (int)(g / Lw) : 0;
    const int l = valid ? (int)(g - (long long)b * Lw) : 0;
    // ---- step 1: x rows, dY' rows
    float dY[8];
    {
      float xr[16], yv[8], dv[8];
      if (valid) {
        load_row<16>(a.x + (((size_t)b * T + (size_t)l * s) * N + i) * C, C, xvec, xr);
        load_row<8>(k.yp + ((size_t)g * M + i) * H, H, yvec, yv);
        const float* dr = k.dout + (size_t)b * k.dout_bs + ((size_t)l * N + n_i) * H;
        load_row<8>(dr, H, yvec && (((uintptr_t)dr & 15) == 0), dv);
#pragma unroll
        for (int h = 0; h < kHP; ++h) {
          const float yn = fmaf(cst[kCstBn1 + h], yv[h], cst[kCstBn1 + 8 + h]);
          const float dyn = dv[h] * 0.5f * lrelu_grad(yn);
          const float yh = (yv[h] - cst[kCstBn1 + 16 + h]) * cst[kCstBn1 + 24 + h];
          float v = cst[kCstBn1 + 32 + h] * dyn - cst[kCstBn1 + 40 + h] - yh * cst[kCstBn1 + 48 + h];
          if (h >= H) v = 0.f;
          dY[h] = v;
          dbt_acc[h] += v;
        }
      } else {
#pragma unroll
        for (int c = 0; c < 16; ++c) xr[c] = 0.f;
#pragma unroll
        for (int h = 0; h < kHP; ++h) dY[h] = 0.f;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        st_hl<SPLIT>(&xk4[q * 128 + tid], &xkl4[q * 128 + tid], xr[4 * q], xr[4 * q + 1], xr[4 * q + 2], xr[4 * q + 3]);
        *rec_ptr(rb, tid, q) = xk4[q * 128 + tid];
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float4 v = rtf4(dY[4 * q], dY[4 * q + 1], dY[4 * q + 2], dY[4 * q + 3]);
        yk4[q * 128 + tid] = v;
        *rec_ptr(ra, tid, 4 + q) = v;
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      issue_proj<SPLIT>(tmem + regA, xk_u, xkl_u, wcb_u, wcbl_u);
      mma_commit(&ctl.bar);
    }
    mbar_wait(&ctl.bar, ph); ph ^= 1;
    tc_fence_after();
    // ---- step 3: F, V rows
    {
      float fv[32];
      tmem_ld32(lane_t + regA, fv);
#pragma unroll
      for (int c = 0; c < kCPH; ++c) fv[c] += cst[kCstBias + c];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        st_hl<SPLIT>(&xk4[q * 128 + tid], &xkl4[q * 128 + tid], fv[4 * q], fv[4 * q + 1], fv[4 * q + 2], fv[4 * q + 3]);
        *rec_ptr(ra, tid, q) = xk4[q * 128 + tid];
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) yk4[(2 + q) * 128 + tid] = rtf4(fv[16 + 4 * q], fv[17 + 4 * q], fv[18 + 4 * q], fv[19 + 4 * q]);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      constexpr uint32_t id = idesc_tf32(128, 0, 0);
      issue_gram<SPLIT>(tmem + regA, xk_u, xkl_u);
      // dA = dY' . V^T   (K = 8: one instruction)
      mma_ss(tmem + regB, sdesc(yk_u, 2048, 128, 0), sdesc(yk_u + 4096, 2048, 128, 0), id, 0);
      mma_commit(&ctl.bar);
    }
    mbar_wait(&ctl.bar, ph); ph ^= 1;
    tc_fence_after();
    // ---- step 5: softmax backward of the row
    {
      float es[WR], da[WR];
#pragma unroll
      for (int q = 0; q < WR / 32; ++q) {
        tmem_ld32(lane_t + regA + wl * WR + q * 32, *reinterpret_cast<float(*)[32]>(&es[q * 32]));
        tmem_ld32(lane_t + regB + wl * WR + q * 32, *reinterpret_cast<float(*)[32]>(&da[q * 32]));
      }
      float mx = -INFINITY;
#pragma unroll
      for (int kk = 0; kk < WR; ++kk) {
        const float lam = fmaxf(es[kk], kLeaky * es[kk]);
        if (kk < M && kk != i) mx = fmaxf(mx, lam);
      }
      const float mxl = mx * kLog2e;
      float sum = 0.f, rs = 0.f;
#pragma unroll
      for (int kk = 0; kk < WR; ++kk) {
        const float sraw = es[kk];
        const float lam = fmaxf(sraw, kLeaky * sraw);
        float e = ex2(fmaf(lam, kLog2e, -mxl));
        if (kk >= M || kk == i) e = 0.f;
        const float mk = ((kk >= N) == j1) ? 1.f : decay;
        da[kk] *= mk;                         // dP = dA o mask
        sum += e;
        rs = fmaf(e, da[kk], rs);
        es[kk] = sraw > 0.f ? e : -e;         // sign keeps leaky_relu'(S)
      }
      const float inv = (valid && sum > 0.f) ? 1.f / sum : 0.f;
      rs *= inv;
      float rho = 0.f;                          // row sum of A
      // outputs, 8 columns at a time: dS row (TMEM, in place + transposed operand), A row (transposed operand)
#pragma unroll
      for (int q = 0; q < WR / 8; ++q) {
        float d8[8], a8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int kk = q * 8 + u;
          const float ev = es[kk];
          const float P = fabsf(ev) * inv;
          const float mk = ((kk >= N) == j1) ? 1.f : decay;
          const float dLam = P * (da[kk] - rs);
          d8[u] = rtf(dLam * (ev > 0.f ? 1.f : kLeaky));
          float av = P * mk;
          if (kk == i && valid) av = 1.f;
          rho += av;
          a8[u] = rtf(av);
        }
        tmem_st8(lane_t + regA + wl * WR + q * 8, d8);
        // transposed operands: mn block = (wl*WR + kk) / 32, row = i, 32-byte chunk (kk % 32) / 8
        const int col = wl * WR + q * 8;
        const size_t off = (size_t)(col >> 5) * (WR * 128) + (size_t)i * 128 + ((((col & 31) >> 3) ^ (i & 3)) * 32);
        reinterpret_cast<float4*>(t1 + off)[0] = make_float4(d8[0], d8[1], d8[2], d8[3]);
        reinterpret_cast<float4*>(t1 + off)[1] = make_float4(d8[4], d8[5], d8[6], d8[7]);
        reinterpret_cast<float4*>(t2 + off)[0] = make_float4(a8[0], a8[1], a8[2], a8[3]);
        reinterpret_cast<float4*>(t2 + off)[1] = make_float4(a8[4], a8[5], a8[6], a8[7]);
      }
#pragma unroll
      for (int h = 0; h < kHP; ++h) sov[h] = fmaf(rho, dY[h], sov[h]);
      tmem_wait_st();
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      constexpr uint32_t id_ts = idesc_tf32(16, 0, 1);      // A from TMEM (K-major by construction), B MN-major
      constexpr uint32_t id_tt = idesc_tf32(16, 1, 1);      // A MN-major (transposed), B MN-major
#pragma unroll
      for (int w2 = 0; w2 < WPT; ++w2) {
        const uint32_t dF = tmem + regB + 32 * w2, dV = dF + 16;
#pragma unroll
        for (int ks = 0; ks < WR / 8; ++ks) {
          const uint32_t brow = ra_u + (w2 * WR + ks * 8) * 128;
          mma_ts(dF, tmem + regA + w2 * WR + ks * 8, sdesc(brow, 512, 512, 1), id_ts, ks);
        }
#pragma unroll
        for (int ks = 0; ks < WR / 8; ++ks) {
          const uint32_t brow = ra_u + (w2 * WR + ks * 8) * 128;
          mma_ss(dF, sdesc(t1_u + ks * 1024, WR * 128, 512, 1), sdesc(brow, 512, 512, 1), id_tt, 1);
          mma_ss(dV, sdesc(t2_u + ks * 1024, WR * 128, 512, 1), sdesc(brow + 64, 512, 512, 1), id_tt, ks);
        }
      }
      mma_commit(&ctl.bar);
    }
    mbar_wait(&ctl.bar, ph); ph ^= 1;
    tc_fence_after();
    // ---- step 7: [dF | dV] rows -> MN-major records (parameter-gradient product) and K-major rows (+ residuals)
    //      for the dx projection
    {
      float fv[32];
      tmem_ld32(lane_t + regB + 32 * wl, fv);
      if (!valid) {
#pragma unroll
        for (int c = 0; c < kCPH; ++c) fv[c] = 0.f;
      }
#pragma unroll
      for (int c = 0; c < kCP; ++c) sof[c] += fv[c];
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        st_hl<SPLIT>(&t2k4[q * 128 + tid], &xk4[q * 128 + tid], fv[4 * q], fv[4 * q + 1], fv[4 * q + 2], fv[4 * q + 3]);
        *rec_ptr(t1, tid, q) = t2k4[q * 128 + tid];
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      constexpr uint32_t id_x = idesc_tf32(16, 0, 0);       // dx partial: all 128 rows share [Wm ; a0 Wtheta]
      constexpr uint32_t id_g = idesc_tf32(16, 1, 1);       // G: both transposed
#pragma unroll
      for (int ks = 0; ks < 3; ++ks) {
        mma_ss(tmem + regA, sdesc(t2_u + ks * 4096, 2048, 128, 0), sdesc(wc2_u + ks * 512, 256, 128, 0), id_x, ks);
        if (SPLIT) {
          mma_ss(tmem + regA, sdesc(t2_u + ks * 4096, 2048, 128, 0), sdesc(wc2l_u + ks * 512, 256, 128, 0), id_x, 1);
          mma_ss(tmem + regA, sdesc(xk_u + ks * 4096, 2048, 128, 0), sdesc(wc2_u + ks * 512, 256, 128, 0), id_x, 1);
        }
      }
#pragma unroll
      for (int ks = 0; ks < 16; ++ks)
        mma_ss(tmem + regA + 64, sdesc(t1_u + ks * 1024, 512, 512, 1), sdesc(rb_u + ks * 1024, 512, 512, 1), id_g, ks);
      mma_commit(&ctl.bar);
    }
    mbar_wait(&ctl.bar, ph); ph ^= 1;
    tc_fence_after();
    // ---- step 9: unfolded dx partial rows, parameter-gradient accumulators
    {
      float dx[16];
      tmem_ld16(lane_t + regA, dx);
      if (valid) {
        float* dst = k.dxp + ((size_t)g * M + i) * C;
        if (xvec) {
#pragma unroll
          for (int q = 0; q < 4; ++q) reinterpret_cast<float4*>(dst)[q] = make_float4(dx[4 * q], dx[4 * q + 1], dx[4 * q + 2], dx[4 * q + 3]);
        } else {
#pragma unroll
          for (int c = 0; c < 16; ++c)
            if (c < C) dst[c] = dx[c];
        }
      }
      if (warp == 0) {
        float gv[16];
        tmem_ld16(lane_t + regA + 64, gv);
#pragma unroll
        for (int q = 0; q < 16; ++q) gacc[q] += gv[q];
      }
    }
    tc_fence_before();
  }

  // ---- CTA epilogue: parameter gradients, BN0 backward sums, dbtheta
  float* red = cst + kCstRed;         // G[24][17] (column 16 = column sums of [dF | dV]) then dbt[8]
  __syncthreads();
  if (tid < kCPH) {
#pragma unroll
    for (int q = 0; q < 16; ++q) red[tid * 17 + q] = gacc[q];
    red[tid * 17 + 16] = 0.f;
  }
  if (tid < 8) red[24 * 17 + tid] = 0.f;
  __syncthreads();
#pragma unroll
  for (int h = 0; h < kHP; ++h) {
    const float v = warp_sum(dbt_acc[h]), v2 = warp_sum(sov[h]);
    if ((tid & 31) == 0) {
      atomicAdd(&red[24 * 17 + h], v);
      atomicAdd(&red[(kCP + h) * 17 + 16], v2);
    }
  }
#pragma unroll
  for (int c = 0; c < kCP; ++c) {
    const float v = warp_sum(sof[c]);
    if ((tid & 31) == 0) atomicAdd(&red[c * 17 + 16], v);
  }
  __syncthreads();
  const float* a0 = cst + kCstA0;
  const float* c0 = a0 + 16;
  const float* mu0 = a0 + 32;
  const float* r0 = a0 + 48;
  for (int idx = tid; idx < kCPH * 16; idx += 128) {
    const int o = idx >> 4, c = idx & 15;
    if (c >= C) continue;
    const float gv = red[o * 17 + c], so = red[o * 17 + 16];
    if (o < C) {
      atomicAdd(&k.dWm[o * C + c], gv);
      if (c == 0) atomicAdd(&k.dbm[o], so);
    } else if (o >= kCP && o - kCP < H) {
      atomicAdd(&k.dWt[(o - kCP) * C + c], fmaf(a0[c], gv, c0[c] * so));
    }
  }
  if (tid < C) {
    const int c = tid;
    float sb = 0.f, sg = 0.f;
    for (int h = 0; h < H; ++h) {
      const float wv = k.Wt[h * C + c];
      const float gv = red[(kCP + h) * 17 + c], so = red[(kCP + h) * 17 + 16];
      sb = fmaf(wv, so, sb);
      sg = fmaf(wv, gv - mu0[c] * so, sg);
    }
    atomicAdd(&k.stats[4 * H + c], (double)sb);
    atomicAdd(&k.stats[4 * H + C + c], (double)(sg * r0[c]));
  }
  if (tid < H) atomicAdd(&k.dbt[tid], red[24 * 17 + tid]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

bool g_tc_attr[64] = {};
void set_tc_attrs() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (g_tc_attr[dev]) return;
  const int cap = 200 * 1024;
#define STG_TC_ATTR(f) cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, cap)
  STG_TC_ATTR((k_block_fwd_tc<32, true, true>)); STG_TC_ATTR((k_block_fwd_tc<32, false, true>));
  STG_TC_ATTR((k_block_fwd_tc<64, true, true>)); STG_TC_ATTR((k_block_fwd_tc<64, false, true>));
  STG_TC_ATTR((k_block_fwd_tc<32, true, false>)); STG_TC_ATTR((k_block_fwd_tc<32, false, false>));
  STG_TC_ATTR((k_block_fwd_tc<64, true, false>)); STG_TC_ATTR((k_block_fwd_tc<64, false, false>));
  STG_TC_ATTR((k_block_bwd_tc<32, true>)); STG_TC_ATTR((k_block_bwd_tc<64, true>));
  STG_TC_ATTR((k_block_bwd_tc<32, false>)); STG_TC_ATTR((k_block_bwd_tc<64, false>));
#undef STG_TC_ATTR
  g_tc_attr[dev] = true;
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
  }
  return n;
}

// CTAs per block: proportional to the block's tiles, ctas_total in all
void split_ctas(const BlkArgs& a, int WR, int ctas_total, int* n0, int* total) {
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

}  // namespace

bool plan_blocks_tc(BlkArgs& a, BlkPlan& p) {
  if (getenv("STG_NO_TC")) return false;
  if (a.C > kCP || ((uintptr_t)a.x & 15)) return false;
  const int M = 2 * a.N;
  if (M > 64) return false;
  for (int z = 0; z < a.nblk; ++z)
    if (a.b[z].w != 2 || a.b[z].H > kHP) return false;
  p.tc = 1;
  p.tc_wr = M <= 32 ? 32 : 64;
  // STG_TC_SPLIT=0: every product single-pass TF32.  Default 1: the products that feed a non-smooth function
  // (projection -> Gram -> softmax / leaky_relu kinks, aggregation -> BatchNorm + leaky_relu) use the 3-term
  // error-compensated TF32 product, the linear backward products stay single-pass.
  const char* e = getenv("STG_TC_SPLIT");
  p.tc_split = e ? (atoi(e) != 0) : 1;
  p.CP = kCP;
  p.HP = kHP;
  return true;
}

int launch_block_forward_tc(const BlkArgs& a, const BlkPlan& p, cudaStream_t s) {
  set_tc_attrs();
  const SmemLayout L = make_layout(p.tc_wr, false, p.tc_split != 0);
  int n0 = 0, total = 0;
  split_ctas(a, p.tc_wr, 2 * sm_count(), &n0, &total);
  ProfScope ps(kProfFwdMain, s);
#define STG_TC_FWD(WR, TR, SP) k_block_fwd_tc<WR, TR, SP><<<total, 128, L.total, s>>>(a, n0)
  if (p.tc_split) {
    if (p.tc_wr == 32) { if (a.training) STG_TC_FWD(32, true, true); else STG_TC_FWD(32, false, true); }
    else { if (a.training) STG_TC_FWD(64, true, true); else STG_TC_FWD(64, false, true); }
  } else {
    if (p.tc_wr == 32) { if (a.training) STG_TC_FWD(32, true, false); else STG_TC_FWD(32, false, false); }
    else { if (a.training) STG_TC_FWD(64, true, false); else STG_TC_FWD(64, false, false); }
  }
#undef STG_TC_FWD
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

int launch_block_backward_tc(const BlkArgs& a, const BlkPlan& p, cudaStream_t s) {
  set_tc_attrs();
  const SmemLayout L = make_layout(p.tc_wr, true, p.tc_split != 0);
  int n0 = 0, total = 0;
  const int per_sm = (size_t)L.total * 2 <= 220 * 1024 ? 2 : 1;
  split_ctas(a, p.tc_wr, per_sm * sm_count(), &n0, &total);
  ProfScope ps(kProfBwdMain, s);
  if (p.tc_split) {
    if (p.tc_wr == 32) k_block_bwd_tc<32, true><<<total, 128, L.total, s>>>(a, n0);
    else k_block_bwd_tc<64, true><<<total, 128, L.total, s>>>(a, n0);
  } else {
    if (p.tc_wr == 32) k_block_bwd_tc<32, false><<<total, 128, L.total, s>>>(a, n0);
    else k_block_bwd_tc<64, false><<<total, 128, L.total, s>>>(a, n0);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

}  // namespace stg
