// tcgen05 / TMEM graph-conv block, BACKWARD kernel (design notes: stg_tc.cuh; math: SURVEY.md section 9.2).
// Reference: autograd through GraphConvpoolMPNN_block_v6.forward, models/FC_STGNN/Model_Base.py:190-225.
#define STG_STAMP_KERNEL 1
#include "stg_tc.cuh"

namespace stg {
namespace tc {

template <int WR, int NT, bool SPLIT>
__global__ void __launch_bounds__(128, 2) k_block_bwd_tc(const BlkArgs a, int ncta0, int pf) {
  constexpr int WPT = 128 / WR;
  extern __shared__ unsigned char smraw[];
  __shared__ TcCtl ctl;
  unsigned char* sm = reinterpret_cast<unsigned char*>(((uintptr_t)smraw + 1023) & ~(uintptr_t)1023);
  const SmemLayout L = make_layout(WR, true, SPLIT, pf ? 2 * (NT ? NT : a.N) : 0);
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

  // saved forward tensors of the block; with pf != 0 the F|V and softmax blocks of a tile's WPT windows (one contiguous
  // run each) are staged in shared memory by two TMA bulk copies issued one tile ahead
  const int MP = saved_mp(M);
  const float* fvs = k.yp + saved_off_fv(nwin * M, H);
  const float* ps = k.yp + saved_off_p(nwin * M, H);
  const float* pf_fv = reinterpret_cast<const float*>(sm + L.pf_fv);
  const float* pf_p = reinterpret_cast<const float*>(sm + L.pf_p);
  auto issue_pf = [&](int t) {
    const uint32_t bfv = (uint32_t)(WPT * kCPH * M * 4), bp = (uint32_t)(WPT * MP * M * 4);
    mbar_expect_tx(&ctl.bar_pf, bfv + bp);
    tma_bulk_g2s(sm + L.pf_fv, fvs + (size_t)t * (WPT * kCPH * M), bfv, &ctl.bar_pf);
    tma_bulk_g2s(sm + L.pf_p, ps + (size_t)t * (WPT * MP * M), bp, &ctl.bar_pf);
  };
  if (tid == 0) {
    mbar_init(&ctl.bar, 1);
    mbar_init(&ctl.bar_pf, 1);
    mbar_init(&ctl.bar_w, WPT);
    mbar_init(&ctl.bar_x, 2);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl.tmem_base)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  float* cst = reinterpret_cast<float*>(sm + L.cst);
  float* red = cst + kCstRed;         // G[24][17] (column 16 = column sums of [dF | dV]) then dbt[8]
  // operand buffers whose pad parts are read by the tensor core but never written per tile
  for (int idx = tid; idx < (L.pf_fv - L.ra) / 16; idx += 128) reinterpret_cast<float4*>(sm + L.ra)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int idx = tid; idx < 24 * 17 + 8; idx += 128) red[idx] = 0.f;
  pdl_sync();
  if (tid == 0 && pf && cta < ntiles && (long long)(cta + 1) * WPT <= nwin) issue_pf(cta);
  tc_prologue(a, k, sm, L, SPLIT, true);
  // BN1 backward coefficients: [0]=a1 [1]=c1 [2]=mu1 [3]=r1 [4]=g1*r1 [5]=q1 [6]=q2
  if (tid < 8) {
    float v[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (tid < H) {
      const int h = tid;
      const double R = (double)a.B * Lw * M, iR = inv_d(R);
      const double m = k.stats[h] * iR;
      double var = k.stats[H + h] * iR - m * m;
      if (var < 0.0) var = 0.0;
      const float r1 = (float)rsqrt_d(var + (double)a.eps);
      const float g1 = k.g1[h];
      v[0] = g1 * r1;
      v[1] = k.b1[h] - v[0] * (float)m;
      v[2] = (float)m;
      v[3] = r1;
      v[4] = g1 * r1;
      v[5] = (float)(g1 * k.stats[2 * H + h] * iR) * r1;
      v[6] = (float)(g1 * k.stats[3 * H + h] * iR) * r1;
    }
#pragma unroll
    for (int q = 0; q < 7; ++q) cst[kCstBn1 + q * 8 + tid] = v[q];
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ctl.tmem_base;
  const uint32_t lane_t = tmem + ((uint32_t)(warp * 32) << 16);
  constexpr uint32_t regA = 0, regB = 128;
  const uint32_t t2l_lo = dlo(smem_u32(sm + L.t2) + 12 * 1024, 2048),
                 yk_lo = dlo(smem_u32(sm + L.yk), 2048), ra_lo = dlo(smem_u32(sm + L.ra), 512),
                 rb_lo = dlo(smem_u32(sm + L.rb), 512), t1_lo = dlo(smem_u32(sm + L.t1), WR * 128),
                 t2_lo = dlo(smem_u32(sm + L.t2), WR * 128), t1g_lo = dlo(smem_u32(sm + L.t1), 512),
                 t2k_lo = dlo(smem_u32(sm + L.t2), 2048), wc2_lo = dlo(smem_u32(sm + L.wc2), 256),
                 wc2l_lo = dlo(smem_u32(sm + L.wc2l), 256);
  float4* yk4 = reinterpret_cast<float4*>(sm + L.yk);
  unsigned char* ra = sm + L.ra;
  unsigned char* rb = sm + L.rb;
  unsigned char* t1 = sm + L.t1;
  unsigned char* t2 = sm + L.t2;
  float4* t2k4 = reinterpret_cast<float4*>(t2);       // [dF | dV] rows, K-major, for the dx projection ...
  float4* t2l4 = reinterpret_cast<float4*>(t2 + 12 * 1024);      // ... and their tf32 residuals

  const int wl = tid / WR, i = tid - wl * WR;
  const bool row_ok = i < M;
  const bool j1 = i >= N;
  const int n_i = j1 ? i - N : i;
  const float decay = cst[kCstMisc];
  const float mkA = j1 ? decay : 1.f, mkB = j1 ? 1.f : decay;
  const bool xvec = (C == 16);
  const bool yvec = (H == 8);
  // column sums of [dF | dV] and of dY' in plain fp32 (they feed the bias / BatchNorm-shift gradients, which are small
  // differences of large sums): dF rows are added as they come out of TMEM; sum_rows dV = sum_i (1 + rowsum(A_i)) dY'_i
  float sof[kCP], sov[kHP], dbt_acc[kHP];
#pragma unroll
  for (int c = 0; c < kCP; ++c) sof[c] = 0.f;
#pragma unroll
  for (int h = 0; h < kHP; ++h) sov[h] = dbt_acc[h] = 0.f;
  uint32_t ph = 0, ph_pf = 0, ph_w = 0, ph_x = 0;

  for (int tile = cta; tile < ntiles; tile += ncta) {
    const long long g = (long long)tile * WPT + wl;
    const bool valid = row_ok && g < nwin;
    const size_t grow = (size_t)g * M + i;          // row of this thread in the [B*L*M, .] saved tensors
    const bool staged = pf && (long long)(tile + 1) * WPT <= nwin;      // partial last tile: plain loads
    STG_STAMP(0)
    if (staged) { mbar_wait(&ctl.bar_pf, ph_pf); ph_pf ^= 1; }
    // ---- step 1: this row's x, saved F | V, saved softmax row, Y', dout -> dY'; operands of dA = dY' . V^T
    float dY[8];
    float es[WR];                                   // softmax numerators e_k of the row (sign bit: S > 0)
    float inv = 0.f;                                // 1 / sum_k e_k
    {
      float xr[16], fv[kCPH], yv[8], dv[8];
      if (valid) {
        const int b = (int)(g / Lw), l = (int)(g - (long long)b * Lw);
        load_row<16>(a.x + (((size_t)b * T + (size_t)l * s) * N + i) * C, C, xvec, xr);
        if (staged) {
          const float* fsrc = pf_fv + wl * kCPH * M + i;
#pragma unroll
          for (int c = 0; c < kCPH; ++c) fv[c] = fsrc[c * M];
          const float* psrc = pf_p + wl * MP * M + i;
#pragma unroll
          for (int kk = 0; kk < WR; ++kk)
            if (kk < M) es[kk] = psrc[kk * M];
          inv = psrc[M * M];
        } else {
          const float* fsrc = fvs + (size_t)g * kCPH * M + i;
#pragma unroll
          for (int c = 0; c < kCPH; ++c) fv[c] = __ldg(fsrc + c * M);
          const float* psrc = ps + (size_t)g * MP * M + i;
#pragma unroll
          for (int kk = 0; kk < WR; ++kk)
            if (kk < M) es[kk] = __ldg(psrc + kk * M);
          inv = __ldg(psrc + (size_t)M * M);
        }
        load_row<8>(k.yp + grow * H, H, yvec, yv);
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
        for (int c = 0; c < kCPH; ++c) fv[c] = 0.f;
#pragma unroll
        for (int kk = 0; kk < WR; ++kk) es[kk] = 0.f;
#pragma unroll
        for (int h = 0; h < kHP; ++h) dY[h] = 0.f;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        *rec_ptr(rb, tid, q) = rtf4(xr[4 * q], xr[4 * q + 1], xr[4 * q + 2], xr[4 * q + 3]);
        *rec_ptr(ra, tid, q) = rtf4(fv[4 * q], fv[4 * q + 1], fv[4 * q + 2], fv[4 * q + 3]);
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float4 v = rtf4(dY[4 * q], dY[4 * q + 1], dY[4 * q + 2], dY[4 * q + 3]);
        yk4[q * 128 + tid] = v;
        *rec_ptr(ra, tid, 4 + q) = v;
        yk4[(2 + q) * 128 + tid] = rtf4(fv[16 + 4 * q], fv[17 + 4 * q], fv[18 + 4 * q], fv[19 + 4 * q]);
      }
    }
    fence_async_smem();
    tc_fence_before();
    STG_STAMP(1)
    __syncthreads();
    STG_STAMP(2)
    if (tid == 0) {
      tc_fence_after();
      constexpr uint32_t id = idesc_tf32(128, 0, 0);
      // dA = dY' . V^T   (K = 8: one instruction)
      mma_ss(tmem + regB, dsc(yk_lo, kHiK), dsc(yk_lo + 256, kHiK), id, 0);
      mma_commit(&ctl.bar);
      // every thread is past its reads of the staged blocks: fetch the next tile's
      const int nt = tile + ncta;
      if (pf && nt < ntiles && (long long)(nt + 1) * WPT <= nwin) issue_pf(nt);
    }
    STG_STAMP(3)
    mbar_wait(&ctl.bar, ph); ph ^= 1;
    tc_fence_after();
    STG_STAMP(4)
    // ---- step 5: softmax backward of the row (P and the sign of S come from the forward)
    {
      float da[WR];
#pragma unroll
      for (int q = 0; q < WR / 32; ++q)
        tmem_ld32(lane_t + regB + wl * WR + q * 32, *reinterpret_cast<float(*)[32]>(&da[q * 32]));
      float rs = 0.f;
#pragma unroll
      for (int kk = 0; kk < WR; ++kk) {
        if (kk < M) {
          da[kk] *= (kk < N ? mkA : mkB);       // dP = dA o mask
          rs = fmaf(fabsf(es[kk]), da[kk], rs);
        }
      }
      rs *= inv;
      float rho = valid ? 1.f : 0.f;            // 1 (identity part of A) + row sum of P o mask
      // outputs, 8 columns at a time: dS row (TMEM, in place + transposed operand), A row (transposed operand)
#pragma unroll
      for (int q = 0; q < WR / 8; ++q) {
        float d8[8], a8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int kk = q * 8 + u;
          if (kk < M) {
            const float ev = es[kk];
            const float P = fabsf(ev) * inv;
            const float dLam = P * (da[kk] - rs);
            d8[u] = rtf(dLam * (ev > 0.f ? 1.f : kLeaky));
            const float av = P * (kk < N ? mkA : mkB);
            rho += av;
            a8[u] = rtf(av);
          } else {
            d8[u] = 0.f;
            a8[u] = 0.f;
          }
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
    STG_STAMP(5)
    __syncthreads();
    STG_STAMP(6)
    // a tcgen05.mma costs its issuing thread 50-110 cycles however small it is (profiles/r02_tcgen05.md): the windows'
    // products are independent accumulators, so lane 0 of warp w issues -- and commits -- the instructions of window w
    if ((tid & 31) == 0 && (tid >> 5) < WPT) {
      tc_fence_after();
      constexpr uint32_t id_ts = idesc_tf32(16, 0, 1);      // A from TMEM (K-major by construction), B MN-major
      constexpr uint32_t id_tt = idesc_tf32(16, 1, 1);      // A MN-major (transposed), B MN-major
      const int w2 = tid >> 5;
      const uint32_t dF = tmem + regB + 32 * w2, dV = dF + 16;
#pragma unroll
      for (int ks = 0; ks < WR / 8; ++ks) {
        const uint32_t brow = ra_lo + ((w2 * WR + ks * 8) * 128 >> 4);
        mma_ts(dF, tmem + regA + w2 * WR + ks * 8, dsc(brow, kHiMN), id_ts, ks);
        mma_ss(dF, dsc(t1_lo + ks * 64, kHiMN), dsc(brow, kHiMN), id_tt, 1);
        mma_ss(dV, dsc(t2_lo + ks * 64, kHiMN), dsc(brow + 4, kHiMN), id_tt, ks);
      }
      mma_commit(&ctl.bar_w);
    }
    STG_STAMP(7)
    mbar_wait(&ctl.bar_w, ph_w); ph_w ^= 1;
    tc_fence_after();
    STG_STAMP(8)
    // ---- step 7: [dF | dV] rows -> MN-major records (parameter-gradient product) and K-major rows (+ residuals)
    //      for the dx projection
    {
      float fv[32];
      tmem_ld32(lane_t + regB + 32 * wl, fv);
      if (valid) {
#pragma unroll
        for (int h = 0; h < kHP; ++h) fv[16 + h] += dY[h];      // the identity part of A^T
      } else {
#pragma unroll
        for (int c = 0; c < kCPH; ++c) fv[c] = 0.f;
      }
#pragma unroll
      for (int c = 0; c < kCP; ++c) sof[c] += fv[c];
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        st_hl<SPLIT>(&t2k4[q * 128 + tid], &t2l4[q * 128 + tid], fv[4 * q], fv[4 * q + 1], fv[4 * q + 2], fv[4 * q + 3]);
        *rec_ptr(t1, tid, q) = t2k4[q * 128 + tid];
      }
    }
    fence_async_smem();
    tc_fence_before();
    STG_STAMP(9)
    __syncthreads();
    STG_STAMP(10)
    if (tid == 0) {                                         // dx partial: all 128 rows share [Wm ; a0 Wtheta]
      tc_fence_after();
      constexpr uint32_t id_x = idesc_tf32(16, 0, 0);
#pragma unroll
      for (int ks = 0; ks < 3; ++ks) {
        mma_ss(tmem + regA, dsc(t2k_lo + ks * 256, kHiK), dsc(wc2_lo + ks * 32, kHiK), id_x, ks);
        if (SPLIT) {
          mma_ss(tmem + regA, dsc(t2k_lo + ks * 256, kHiK), dsc(wc2l_lo + ks * 32, kHiK), id_x, 1);
          mma_ss(tmem + regA, dsc(t2l_lo + ks * 256, kHiK), dsc(wc2_lo + ks * 32, kHiK), id_x, 1);
        }
      }
      mma_commit(&ctl.bar_x);
    } else if (tid == 32) {                                 // parameter-gradient product G: its own accumulator, its own thread
      tc_fence_after();
      constexpr uint32_t id_g = idesc_tf32(16, 1, 1);       // both transposed
#pragma unroll
      for (int ks = 0; ks < 16; ++ks)
        mma_ss(tmem + regA + 64, dsc(t1g_lo + ks * 64, kHiMN), dsc(rb_lo + ks * 64, kHiMN), id_g, ks);
      mma_commit(&ctl.bar_x);
    }
    STG_STAMP(11)
    mbar_wait(&ctl.bar_x, ph_x); ph_x ^= 1;
    tc_fence_after();
    STG_STAMP(12)
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
        // G[o][c] = sum_rows [dF | dV][o] * x[c] of this tile: lane o < 24 owns row o of the CTA's accumulator
        float gv[16];
        tmem_ld16(lane_t + regA + 64, gv);
        if (tid < kCPH) {
#pragma unroll
          for (int q = 0; q < 16; ++q) red[tid * 17 + q] += gv[q];
        }
      }
    }
    tc_fence_before();
    STG_STAMP(13)
  }

  // ---- CTA epilogue: parameter gradients, BN0 backward sums, dbtheta
  //      per-warp partial column sums go to their own slots (no shared-memory atomics), one barrier, then they are
  //      folded into column 16 of the G accumulator / the dbtheta slots
  float* part = red + 24 * 17 + 8;      // [4 warps][32]: sof[16] | sov[8] | dbt[8]
#pragma unroll
  for (int c = 0; c < kCP; ++c) {
    const float v = warp_sum(sof[c]);
    if ((tid & 31) == 0) part[warp * 32 + c] = v;
  }
#pragma unroll
  for (int h = 0; h < kHP; ++h) {
    const float v = warp_sum(sov[h]), v2 = warp_sum(dbt_acc[h]);
    if ((tid & 31) == 0) {
      part[warp * 32 + 16 + h] = v;
      part[warp * 32 + 24 + h] = v2;
    }
  }
  __syncthreads();
  if (tid < 32) {
    const float v = (part[tid] + part[32 + tid]) + (part[64 + tid] + part[96 + tid]);
    if (tid < 24) red[tid * 17 + 16] = v;          // column sums of [dF | dV]
    else red[24 * 17 + (tid - 24)] = v;            // dbtheta
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
#ifdef STG_TC_TIMING
  if (blockIdx.x == 0 && (tid == 0 || tid == 64)) g_tc_stamp[tid == 64][15] = clock64();
  if (tid == 0 && blockIdx.x < 1024) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_cta_time[blockIdx.x][1] = t;
  }
#endif
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

template <int WR, int NT, bool SPLIT>
static void launch_bwd(const BlkArgs& a, int total, int n0, size_t smem, int pf, cudaStream_t s) {
  static bool attr[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!attr[dev]) {
    cudaFuncSetAttribute(k_block_bwd_tc<WR, NT, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr[dev] = true;
  }
  launch_pdl(k_block_bwd_tc<WR, NT, SPLIT>, dim3(total), dim3(128), smem, s, a, n0, pf);
}

}  // namespace tc

#ifdef STG_TC_TIMING
extern "C" int stg_debug_tc_cta_times_bwd(unsigned long long* out3072) {
  return cudaMemcpyFromSymbol(out3072, tc::g_cta_time, sizeof(unsigned long long) * 3072) == cudaSuccess ? 0 : -1;
}
extern "C" int stg_debug_tc_stamps_bwd(long long* out32) {
  return cudaMemcpyFromSymbol(out32, tc::g_tc_stamp, sizeof(long long) * 32) == cudaSuccess ? 0 : -1;
}
#endif

int launch_block_backward_tc(const BlkArgs& a, const BlkPlan& p, cudaStream_t s) {
  using namespace tc;
  // 227 KB usable per SM, 1 KB reserved per CTA.  The saved blocks are staged by TMA one tile ahead when that keeps the
  // CTAs per SM of the plain-load variant (STG_TC_NO_PREFETCH=1 forces plain loads).
  auto fit = [](const SmemLayout& l) { return 2 * ((size_t)l.total + 1024) <= 227 * 1024 ? 2 : 1; };
  static const bool no_pf = [] { const char* e = getenv("STG_TC_NO_PREFETCH"); return e && *e == '1'; }();
  SmemLayout L = make_layout(p.tc_wr, true, p.tc_split != 0);
  int pf = 0;
  {
    const SmemLayout Lp = make_layout(p.tc_wr, true, p.tc_split != 0, 2 * a.N);
    const bool aligned = (((uintptr_t)a.b[0].yp | (uintptr_t)a.b[1].yp) & 15) == 0;
    if (!no_pf && aligned && fit(Lp) == fit(L)) { L = Lp; pf = 1; }
  }
  int n0 = 0, total = 0;
  const int per_sm = fit(L);
  split_ctas(a, p.tc_wr, per_sm * sm_count(), &n0, &total);
  ProfScope ps(kProfBwdMain, s);
#define STG_TC_BWD(WR, NT)                                                  \
  do {                                                                      \
    if (p.tc_split) launch_bwd<WR, NT, true>(a, total, n0, L.total, pf, s);     \
    else launch_bwd<WR, NT, false>(a, total, n0, L.total, pf, s);               \
  } while (0)
  if (p.tc_wr == 32) {
    if (a.N == 14) STG_TC_BWD(32, 14); else STG_TC_BWD(32, 0);
  } else {
    if (a.N == 21) STG_TC_BWD(64, 21); else if (a.N == 20) STG_TC_BWD(64, 20); else STG_TC_BWD(64, 0);
  }
#undef STG_TC_BWD
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

}  // namespace stg
