// tcgen05 / TMEM graph-conv block, FORWARD kernel (design notes: stg_tc.cuh).
// Reference: GraphConvpoolMPNN_block_v6.forward, models/FC_STGNN/Model_Base.py:190-225.
#define STG_STAMP_KERNEL 0
#include "stg_tc.cuh"

namespace stg {
namespace tc {


// WR rows per window slot (32 / 64), NT = number of sensors N when known at compile time (0: runtime, N <= WR/2),
// TRAIN: write pre-BN Y' and its batch moments (BN1 + leaky_relu + pooling are fused into the FC head) and save the
// F | V rows and the softmax rows for the backward; else apply BN1 with running statistics, leaky_relu and the window
// mean here.  SPLIT: 3-term TF32 products for the projection and the Gram matrix.
// Two tcgen05 phases per tile (projection, Gram); the aggregation Z = A.V over the <= 42 nodes of the thread's own
// window runs in fp32 on the CUDA cores: as tcgen05.mma it needs (rows per window / 8) instructions per window whose
// result is only used on that window's 32 lanes, and an instruction costs >= ~50 cycles however small it is.
template <int WR, int NT, bool TRAIN, bool SPLIT>
__global__ void __launch_bounds__(128, WR == 32 ? 4 : 2) k_block_fwd_tc(const BlkArgs a, int ncta0) {
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
#ifdef STG_TC_TIMING
  if (blockIdx.x == 0 && (tid == 0 || tid == 64)) g_tc_stamp[tid == 64][14] = clock64();
  if (tid == 0 && blockIdx.x < 1024) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_cta_time[blockIdx.x][0] = t;
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    g_cta_time[blockIdx.x][2] = smid;
  }
#endif
  const int N = NT ? NT : a.N, M = 2 * N;
  const int C = a.C, T = a.T, H = k.H, s = k.stride, Lw = k.L;
  const long long nwin = (long long)a.B * Lw;
  const int ntiles = (int)((nwin + WPT - 1) / WPT);
  const int MP = saved_mp(M);
  float* fvs = k.yp + saved_off_fv(nwin * M, H);
  float* ps = k.yp + saved_off_p(nwin * M, H);

  if (tid == 0) mbar_init(&ctl.bar, 1);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl.tmem_base)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  pdl_sync();
  tc_prologue(a, k, sm, L, SPLIT, false);
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
  // descriptor low words (start address + leading-dimension offset); per-MMA offsets are compile-time constants
  const uint32_t xk_lo = dlo(smem_u32(sm + L.xk), 2048), xkl_lo = dlo(smem_u32(sm + L.xkl), 2048),
                 wcb_lo = dlo(smem_u32(sm + L.wcb), 512), wcbl_lo = dlo(smem_u32(sm + L.wcbl), 512);
  float4* xk4 = reinterpret_cast<float4*>(sm + L.xk);
  float4* xkl4 = reinterpret_cast<float4*>(sm + L.xkl);
  float4* vs4 = reinterpret_cast<float4*>(sm + L.vs);

  const int wl = tid / WR, i = tid - wl * WR;       // window slot inside the tile, row inside the window
  const bool row_ok = i < M;
  const bool j1 = i >= N;                           // time offset of this row inside its window
  const float decay = cst[kCstMisc];
  const float mkA = j1 ? decay : 1.f, mkB = j1 ? 1.f : decay;      // mask factor for columns < N / >= N
  const bool xvec = (C == 16);
  const bool yvec = (H == 8);
  float st1[kHP], st2[kHP];
#pragma unroll
  for (int h = 0; h < kHP; ++h) st1[h] = st2[h] = 0.f;
  uint32_t ph = 0;

  // x row of this thread for the first tile (later tiles are prefetched while the tensor core works)
  float xr[16];
  auto fetch_x = [&](int tile) {
    const long long g = (long long)tile * WPT + wl;
    if (tile < ntiles && row_ok && g < nwin) {
      const int b = (int)(g / Lw), l = (int)(g - (long long)b * Lw);
      load_row<16>(a.x + (((size_t)b * T + (size_t)l * s) * N + i) * C, C, xvec, xr);
    } else {
#pragma unroll
      for (int c = 0; c < 16; ++c) xr[c] = 0.f;
    }
  };
  fetch_x(cta);

  for (int tile = cta; tile < ntiles; tile += ncta) {
    const long long g = (long long)tile * WPT + wl;
    const bool valid = row_ok && g < nwin;
    const size_t grow = (size_t)g * M + i;          // row of this thread in the [B*L*M, .] saved tensors
    STG_STAMP(0)
    // ---- x rows -> K-major operand
#pragma unroll
    for (int q = 0; q < 4; ++q)
      st_hl<SPLIT>(&xk4[q * 128 + tid], &xkl4[q * 128 + tid], xr[4 * q], xr[4 * q + 1], xr[4 * q + 2], xr[4 * q + 3]);
    fence_async_smem();
    tc_fence_before();
    STG_STAMP(1)
    __syncthreads();
    STG_STAMP(2)
    if (tid == 0) {
      tc_fence_after();
      issue_proj2<SPLIT>(tmem, xk_lo, xkl_lo, wcb_lo, wcbl_lo);
      mma_commit(&ctl.bar);
    }
    STG_STAMP(3)
    fetch_x(tile + ncta);
    mbar_wait(&ctl.bar, ph); ph ^= 1;
    tc_fence_after();
    STG_STAMP(4)
    // ---- F rows -> K-major operand of the Gram product; V rows -> shared memory (fp32) for the aggregation
    float2 y2[4];
    {
      float fv[32];
      tmem_ld32(lane_t, fv);
      const float4* b4 = reinterpret_cast<const float4*>(cst + kCstBias);
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        const float4 bb = b4[q];
        fv[4 * q] += bb.x; fv[4 * q + 1] += bb.y; fv[4 * q + 2] += bb.z; fv[4 * q + 3] += bb.w;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        st_hl<SPLIT>(&xk4[q * 128 + tid], &xkl4[q * 128 + tid], fv[4 * q], fv[4 * q + 1], fv[4 * q + 2], fv[4 * q + 3]);
      vs4[tid * 2] = make_float4(fv[16], fv[17], fv[18], fv[19]);
      vs4[tid * 2 + 1] = make_float4(fv[20], fv[21], fv[22], fv[23]);
      const float4* t4 = reinterpret_cast<const float4*>(cst + kCstBt);
      const float4 ta = t4[0], tb = t4[1];
      y2[0] = make_float2(fv[16] + ta.x, fv[17] + ta.y);      // the +I term of A and btheta
      y2[1] = make_float2(fv[18] + ta.z, fv[19] + ta.w);
      y2[2] = make_float2(fv[20] + tb.x, fv[21] + tb.y);
      y2[3] = make_float2(fv[22] + tb.z, fv[23] + tb.w);
      if (TRAIN && valid) {
        float* dst = fvs + (size_t)g * kCPH * M + i;
#pragma unroll
        for (int c = 0; c < kCPH; ++c) dst[c * M] = fv[c];
      }
    }
    fence_async_smem();
    tc_fence_before();
    STG_STAMP(5)
    __syncthreads();
    STG_STAMP(6)
    if (tid == 0) {
      tc_fence_after();
      issue_gram2<SPLIT>(tmem, xk_lo, xkl_lo);
      mma_commit(&ctl.bar);
    }
    STG_STAMP(7)
    mbar_wait(&ctl.bar, ph); ph ^= 1;
    tc_fence_after();
    STG_STAMP(8)
    // ---- row softmax over the other nodes of the window, then y += sum_k (P o mask)[k] V_k
    {
      float sv[WR];
#pragma unroll
      for (int q = 0; q < WR / 32; ++q) tmem_ld32(lane_t + wl * WR + q * 32, *reinterpret_cast<float(*)[32]>(&sv[q * 32]));
      float mx = -INFINITY;
#pragma unroll
      for (int kk = 0; kk < WR; ++kk) {
        if (kk < M) {
          const float v = (kk == i) ? -INFINITY : sv[kk];
          const float lam = fmaxf(v, kLeaky * v);     // leaky_relu keeps the sign of S
          sv[kk] = lam;
          mx = fmaxf(mx, lam);
        }
      }
      const float mxl = mx * kLog2e;
      float sum = 0.f;
#pragma unroll
      for (int kk = 0; kk < WR; ++kk) {
        if (kk < M) {
          const float lam = sv[kk];
          const float e = ex2(fmaf(lam, kLog2e, -mxl));
          sum += e;
          // numerator with the sign of S in the sign bit (what the backward needs of S)
          sv[kk] = __uint_as_float(__float_as_uint(e) | (__float_as_uint(lam) & 0x80000000u));
        }
      }
      const float inv = (valid && sum > 0.f) ? 1.f / sum : 0.f;
      const float invA = inv * mkA, invB = inv * mkB;
      const float4* vw = vs4 + (size_t)wl * WR * 2;
#pragma unroll
      for (int kk = 0; kk < WR; ++kk) {
        if (kk < M) {
          const float pa = fabsf(sv[kk]) * (kk < N ? invA : invB);
          const float2 pp = make_float2(pa, pa);
          const float4 v0 = vw[kk * 2], v1 = vw[kk * 2 + 1];
          ffma2(y2[0], pp, make_float2(v0.x, v0.y));
          ffma2(y2[1], pp, make_float2(v0.z, v0.w));
          ffma2(y2[2], pp, make_float2(v1.x, v1.y));
          ffma2(y2[3], pp, make_float2(v1.z, v1.w));
        }
      }
      if (TRAIN && valid) {
        float* dst = ps + (size_t)g * MP * M + i;
#pragma unroll
        for (int kk = 0; kk < WR; ++kk)
          if (kk < M) dst[kk * M] = sv[kk];
        dst[(size_t)M * M] = inv;
      }
    }
    float y[kHP];
#pragma unroll
    for (int h = 0; h < 4; ++h) { y[2 * h] = y2[h].x; y[2 * h + 1] = y2[h].y; }
    STG_STAMP(9)
    // ---- Y' = A.V + btheta
    {
      if (TRAIN) {
        if (valid) {
          float* yrow = k.yp + grow * H;
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
        float* ex = reinterpret_cast<float*>(sm + L.xk);      // K-major F rows are consumed; reuse as [128 + 32][8]
        __syncthreads();
#pragma unroll
        for (int h = 0; h < kHP; ++h) ex[tid * 8 + h] = lrelu(fmaf(cst[kCstBn1 + h], y[h], cst[kCstBn1 + 8 + h]));
        __syncthreads();
        if (valid && !j1) {
          const int b = (int)(g / Lw), l = (int)(g - (long long)b * Lw);
          float* orow = k.out + (size_t)b * k.out_bs + ((size_t)l * N + i) * H;
          const float* p0 = ex + tid * 8;
          const float* p1 = ex + (tid + N) * 8;
          for (int h = 0; h < H; ++h) orow[h] = 0.5f * (p0[h] + p1[h]);
        }
      }
    }
    // (the next tile's vs / xk writes come after its first barrier resp. after this tile's Gram product; only the
    //  eval epilogue above still reads xk here)
    if (!TRAIN) __syncthreads();
    tc_fence_before();
    STG_STAMP(13)
  }

#ifdef STG_TC_TIMING
  if (blockIdx.x == 0 && (tid == 0 || tid == 64)) g_tc_stamp[tid == 64][10] = clock64();
#endif
  if (TRAIN) {
    // per-warp partial moments in their own slots (no shared-memory atomics: float atomics on shared memory are CAS
    // loops on sm_100a), one barrier, then one double atomic per feature and CTA
    float* part = cst + kCstRed;        // [4 warps][16]
#pragma unroll
    for (int h = 0; h < kHP; ++h) {
      const float v1 = warp_sum(st1[h]), v2 = warp_sum(st2[h]);
      if ((tid & 31) == 0) {
        part[warp * 16 + h] = v1;
        part[warp * 16 + kHP + h] = v2;
      }
    }
    __syncthreads();
    if (tid < 2 * kHP) {
      const int h = tid & (kHP - 1);
      if (h < H) {
        const float v = (part[tid] + part[16 + tid]) + (part[32 + tid] + part[48 + tid]);
        atomicAdd(&k.stats[(tid < kHP ? 0 : H) + h], (double)v);
      }
    }
  }
#ifdef STG_TC_TIMING
  if (blockIdx.x == 0 && (tid == 0 || tid == 64)) g_tc_stamp[tid == 64][11] = clock64();
#endif
  tc_fence_before();
  __syncthreads();
#ifdef STG_TC_TIMING
  if (blockIdx.x == 0 && (tid == 0 || tid == 64)) g_tc_stamp[tid == 64][15] = clock64();
  if (tid == 0 && blockIdx.x < 1024) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_cta_time[blockIdx.x][1] = t;
  }
#endif
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

template <int WR, int NT, bool SPLIT>
static void launch_fwd(const BlkArgs& a, int total, int n0, size_t smem, cudaStream_t s) {
  static bool attr[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!attr[dev]) {
    cudaFuncSetAttribute(k_block_fwd_tc<WR, NT, true, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_block_fwd_tc<WR, NT, false, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr[dev] = true;
  }
  if (a.training) launch_pdl(k_block_fwd_tc<WR, NT, true, SPLIT>, dim3(total), dim3(128), smem, s, a, n0);
  else launch_pdl(k_block_fwd_tc<WR, NT, false, SPLIT>, dim3(total), dim3(128), smem, s, a, n0);
}

}  // namespace tc

#ifdef STG_TC_TIMING
extern "C" int stg_debug_tc_cta_times_fwd(unsigned long long* out3072) {
  return cudaMemcpyFromSymbol(out3072, tc::g_cta_time, sizeof(unsigned long long) * 3072) == cudaSuccess ? 0 : -1;
}
extern "C" int stg_debug_tc_stamps_fwd(long long* out32) {
  return cudaMemcpyFromSymbol(out32, tc::g_tc_stamp, sizeof(long long) * 32) == cudaSuccess ? 0 : -1;
}
#endif

bool plan_blocks_tc(BlkArgs& a, BlkPlan& p) {
  if (getenv("STG_NO_TC")) return false;
  if (a.C > tc::kCP || ((uintptr_t)a.x & 15)) return false;
  const int M = 2 * a.N;
  if (M > 64) return false;
  for (int z = 0; z < a.nblk; ++z)
    if (a.b[z].w != 2 || a.b[z].H > tc::kHP) return false;
  p.tc = 1;
  p.tc_wr = M <= 32 ? 32 : 64;
  // STG_TC_SPLIT=0: every product single-pass TF32.  Default 1: the products that feed a non-smooth function
  // (projection -> Gram -> softmax / leaky_relu kinks, aggregation -> BatchNorm + leaky_relu) and the final dx
  // projection use the 3-term error-compensated TF32 product, the other backward products stay single-pass.
  const char* e = getenv("STG_TC_SPLIT");
  p.tc_split = e ? (atoi(e) != 0) : 1;
  p.CP = tc::kCP;
  p.HP = tc::kHP;
  return true;
}

int launch_block_forward_tc(const BlkArgs& a, const BlkPlan& p, cudaStream_t s) {
  using namespace tc;
  const SmemLayout L = make_layout(p.tc_wr, false, p.tc_split != 0);
  int n0 = 0, total = 0;
  split_ctas(a, p.tc_wr, (p.tc_wr == 32 ? 4 : 2) * sm_count(), &n0, &total);
  ProfScope ps(kProfFwdMain, s);
#define STG_TC_FWD(WR, NT)                                                  \
  do {                                                                      \
    if (p.tc_split) launch_fwd<WR, NT, true>(a, total, n0, L.total, s);     \
    else launch_fwd<WR, NT, false>(a, total, n0, L.total, s);               \
  } while (0)
  if (p.tc_wr == 32) {
    if (a.N == 14) STG_TC_FWD(32, 14); else STG_TC_FWD(32, 0);
  } else {
    if (a.N == 21) STG_TC_FWD(64, 21); else if (a.N == 20) STG_TC_FWD(64, 20); else STG_TC_FWD(64, 0);
  }
#undef STG_TC_FWD
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

}  // namespace stg
