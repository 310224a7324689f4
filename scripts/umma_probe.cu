// Probe of the tcgen05 / TMEM conventions the graph-conv block kernels rely on (sm_100a).
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/umma_probe scripts/umma_probe.cu
// Run on a B200 (gpurun).  Every test multiplies small exactly-representable matrices with one
// tcgen05.mma chain (kind::tf32, M=128, operands in shared memory in the no-swizzle "chunk" layout
// X4[q][r][4 floats]) and compares with a host product.  For each operand major-ness both readings of
// the descriptor's (LBO, SBO) fields are tried, so one run settles the convention.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

#define DEVINL __device__ __forceinline__

DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

DEVINL uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout = 0) {
  uint64_t d = (uint64_t)(layout & 7) << 61;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  return d;                 // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE
}

DEVINL uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

DEVINL void mma_tf32_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
DEVINL void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
DEVINL bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
DEVINL bool mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    if (clock64() - t0 > 2000000000LL) return false;
  }
  return true;
}

struct Params {
  int N, K;            // M = 128
  int a_mn, b_mn;      // 0 = K-major, 1 = MN-major
  int a_lbo, a_sbo, a_kstep;   // bytes; kstep = start-address advance per MMA (K = 8)
  int b_lbo, b_sbo, b_kstep;
  int a_bytes, b_bytes;        // operand images in shared memory
  int a_tmem;          // 1: A operand goes through TMEM (tcgen05.st) instead of shared memory
  int a_layout, b_layout;      // descriptor layout type: 0 = no swizzle, 1 = SWIZZLE_128B_BASE32B
  int a_off, b_off;            // byte offset of the operand start inside its image
};

// A image / B image are prepared by the host in their shared-memory byte layout.
__global__ void __launch_bounds__(128) k_probe(Params p, const float* __restrict__ a_img, const float* __restrict__ b_img,
                                               const float* __restrict__ a_rows, float* __restrict__ out, int* status) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  float* sa = reinterpret_cast<float*>(sm);
  float* sb = reinterpret_cast<float*>(sm + ((p.a_bytes + 1023) / 1024) * 1024);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < p.a_bytes / 4; i += 128) sa[i] = a_img[i];
  for (int i = tid; i < p.b_bytes / 4; i += 128) sb[i] = b_img[i];
  if (tid == 0) mbar_init(&bar, 1);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t d_tmem = tmem;              // columns [0, N)
  const uint32_t a_tm = tmem + 256;          // columns [256, 256+K) for the TMEM-A variant

  if (p.a_tmem) {
    // thread = row; write K values of its row (a_rows is row-major [128][K]) to TMEM lanes
    const uint32_t taddr = a_tm + ((uint32_t)(warp * 32) << 16);
    for (int k0 = 0; k0 < p.K; k0 += 8) {
      uint32_t v[8];
      for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(a_rows[tid * p.K + k0 + j]);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr + k0), "r"(v[0]),
                   "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                   : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }

  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, p.N, p.a_mn, p.b_mn);
    const uint32_t abase = smem_u32(sa), bbase = smem_u32(sb);
    for (int ks = 0; ks < p.K / 8; ++ks) {
      const uint64_t bd = make_desc(bbase + p.b_off + ks * p.b_kstep, p.b_lbo, p.b_sbo, p.b_layout);
      if (p.a_tmem) {
        mma_tf32_ts(d_tmem, a_tm + ks * 8, bd, idesc, ks > 0);
      } else {
        const uint64_t ad = make_desc(abase + p.a_off + ks * p.a_kstep, p.a_lbo, p.a_sbo, p.a_layout);
        mma_tf32_ss(d_tmem, ad, bd, idesc, ks > 0);
      }
    }
    umma_commit(&bar);
  }
  __syncwarp();
  const bool ok = mbar_wait_bounded(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!ok) {
    if (tid == 0) *status = 1;
  } else {
    const uint32_t taddr = d_tmem + ((uint32_t)(warp * 32) << 16);
    for (int n0 = 0; n0 < p.N; n0 += 8) {
      uint32_t v[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(taddr + n0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) out[tid * p.N + n0 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}


// ---- host side -------------------------------------------------------------------------------
// operand formats
//   FMT_K_INTER : K-major, no swizzle: chunk layout X4[k/4][mn][4]           (validated: LBO = K-chunk stride, SBO = 128)
//   FMT_MN_SW32B: MN-major, SWIZZLE_128B_BASE32B: rows of 128 B (32 mn values) per k, 4-row atoms (512 B),
//                 32-byte chunk index ^= (k % 4); mn blocks of 32 at stride MNB, k blocks of 4 at stride 512
enum { FMT_K_INTER = 0, FMT_MN_SW32B = 1, FMT_MN_SW32B_NOSWZ = 2 };

struct Operand {
  int fmt, hypo, MN, K;
  int col_off;      // MN-major only: first mn value lives at column col_off of the 128-byte rows
  std::vector<float> img;
  int lbo, sbo, kstep, layout, off, mn_major;
};

static void build_operand(Operand& o, const std::vector<float>& mat) {
  const int MN = o.MN, K = o.K;
  if (o.fmt == FMT_K_INTER) {
    o.img.assign((size_t)(K / 4) * MN * 4, 0.f);
    for (int m = 0; m < MN; ++m)
      for (int k = 0; k < K; ++k) o.img[((size_t)(k / 4) * MN + m) * 4 + (k % 4)] = mat[(size_t)m * K + k];
    const int along_mn = 128, along_k = MN * 16;
    if (o.hypo == 0) { o.lbo = along_k; o.sbo = along_mn; } else { o.lbo = along_mn; o.sbo = along_k; }
    o.kstep = 2 * along_k; o.layout = 0; o.off = 0; o.mn_major = 0;
  } else {
    const int nblk = (o.col_off + MN + 31) / 32;
    const int MNB = K * 128;                       // bytes between 32-wide mn blocks
    o.img.assign((size_t)nblk * MNB / 4 + 4096, 0.f);
    for (int m = 0; m < MN; ++m)
      for (int k = 0; k < K; ++k) {
        const int col = o.col_off + m, blk = col / 32, c = col % 32;
        int chunk = c / 8;
        if (o.fmt == FMT_MN_SW32B) chunk ^= (k % 4);
        const size_t byte = (size_t)blk * MNB + (size_t)(k / 4) * 512 + (k % 4) * 128 + chunk * 32 + (c % 8) * 4;
        o.img[byte / 4] = mat[(size_t)m * K + k];
      }
    if (o.hypo == 0) { o.lbo = MNB; o.sbo = 512; } else { o.lbo = 512; o.sbo = MNB; }
    o.kstep = 2 * 512; o.layout = 1; o.off = o.col_off * 4; o.mn_major = 1;
    // NOTE: with col_off != 0 the swizzle of the start address itself is what is being probed
  }
}

static int run_case(const char* name, int N, int K, int a_fmt, int a_hypo, int b_fmt, int b_hypo, int b_col_off, int a_tmem,
                    bool exact, int ncheck = -1) {
  const int M = 128;
  if (ncheck < 0) ncheck = N;
  std::vector<float> A((size_t)M * K), Bm((size_t)N * K);
  srand(1234);
  for (auto& v : A) v = exact ? (float)((rand() % 17) - 8) * 0.25f : (float)rand() / RAND_MAX - 0.5f;
  for (auto& v : Bm) v = exact ? (float)((rand() % 13) - 6) * 0.5f : (float)rand() / RAND_MAX - 0.5f;
  Operand oa{a_fmt, a_hypo, M, K, 0}, ob{b_fmt, b_hypo, N, K, b_col_off};
  build_operand(oa, A);
  build_operand(ob, Bm);
  Params p{};
  p.N = N; p.K = K; p.a_mn = oa.mn_major; p.b_mn = ob.mn_major; p.a_tmem = a_tmem;
  p.a_bytes = (int)oa.img.size() * 4; p.b_bytes = (int)ob.img.size() * 4;
  p.a_lbo = oa.lbo; p.a_sbo = oa.sbo; p.a_kstep = oa.kstep; p.a_layout = oa.layout; p.a_off = oa.off;
  p.b_lbo = ob.lbo; p.b_sbo = ob.sbo; p.b_kstep = ob.kstep; p.b_layout = ob.layout; p.b_off = ob.off;
  float *da, *db, *dr, *dout; int* dstat;
  cudaMalloc(&da, p.a_bytes); cudaMalloc(&db, p.b_bytes); cudaMalloc(&dr, A.size() * 4);
  cudaMalloc(&dout, (size_t)M * N * 4); cudaMalloc(&dstat, 4);
  cudaMemcpy(da, oa.img.data(), p.a_bytes, cudaMemcpyHostToDevice);
  cudaMemcpy(db, ob.img.data(), p.b_bytes, cudaMemcpyHostToDevice);
  cudaMemcpy(dr, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dout, 0xff, (size_t)M * N * 4); cudaMemset(dstat, 0, 4);
  const size_t smem = 200 * 1024;  // generous: a wrong stride hypothesis must stay inside the allocation
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k_probe<<<1, 128, smem>>>(p, da, db, dr, dout, dstat);
  cudaError_t e = cudaDeviceSynchronize();
  int stat = 0;
  std::vector<float> out((size_t)M * N);
  if (e == cudaSuccess) {
    cudaMemcpy(&stat, dstat, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
  }
  double maxerr = 0, maxref = 0;
  int nbad = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < ncheck; ++n) {
      double r = 0;
      for (int k = 0; k < K; ++k) r += (double)A[(size_t)m * K + k] * Bm[(size_t)n * K + k];
      const double d = fabs(r - out[(size_t)m * N + n]);
      if (!(d < 1e-3)) ++nbad;
      maxerr = fmax(maxerr, d);
      maxref = fmax(maxref, fabs(r));
    }
  const bool pass = (e == cudaSuccess && !stat && maxerr < (exact ? 1e-6 : 2e-3) * fmax(1.0, maxref));
  printf("%-30s N=%3d K=%3d A(fmt %d hypo %d tmem %d) B(fmt %d hypo %d off %2d) : cuda=%s timeout=%d maxerr=%.3e bad=%d/%d out[0..3]=%g %g %g %g %s\n",
         name, N, K, a_fmt, a_hypo, a_tmem, b_fmt, b_hypo, b_col_off, cudaGetErrorString(e), stat, maxerr, nbad, M * ncheck,
         out[0], out[1], out[2], out[3], pass ? "PASS" : "FAIL");
  fflush(stdout);
  cudaFree(da); cudaFree(db); cudaFree(dr); cudaFree(dout); cudaFree(dstat);
  if (e != cudaSuccess) { cudaDeviceReset(); return 2; }
  return 0;
}

int main() {
  cudaDeviceProp pr;
  cudaGetDeviceProperties(&pr, 0);
  printf("device %s sm_%d%d SMs %d\n", pr.name, pr.major, pr.minor, pr.multiProcessorCount);
  run_case("K/K baseline N=128 K=16", 128, 16, FMT_K_INTER, 0, FMT_K_INTER, 0, 0, 0, true);
  run_case("K/K N=16 K=32", 16, 32, FMT_K_INTER, 0, FMT_K_INTER, 0, 0, 0, true);
  run_case("K/K N=32 K=128", 32, 128, FMT_K_INTER, 0, FMT_K_INTER, 0, 0, 0, true);
  for (int fmt = 1; fmt <= 2; ++fmt)
    for (int hb = 0; hb < 2; ++hb) run_case("A=K  B=MN32B", 16, 32, FMT_K_INTER, 0, fmt, hb, 0, 0, true);
  for (int fmt = 1; fmt <= 2; ++fmt)
    for (int hb = 0; hb < 2; ++hb) run_case("A=K  B=MN32B N=32", 32, 32, FMT_K_INTER, 0, fmt, hb, 0, 0, true);
  for (int fmt = 1; fmt <= 2; ++fmt)
    for (int ha = 0; ha < 2; ++ha) run_case("A=MN32B B=K", 16, 32, fmt, ha, FMT_K_INTER, 0, 0, 0, true);
  for (int ha = 0; ha < 2; ++ha)
    for (int hb = 0; hb < 2; ++hb) run_case("A=MN32B B=MN32B", 16, 32, FMT_MN_SW32B, ha, FMT_MN_SW32B, hb, 0, 0, true);
  for (int hb = 0; hb < 2; ++hb) run_case("A=tmem B=MN32B", 16, 32, FMT_K_INTER, 0, FMT_MN_SW32B, hb, 0, 1, true);
  for (int hb = 0; hb < 2; ++hb) run_case("A=K B=MN32B col_off 16", 16, 32, FMT_K_INTER, 0, FMT_MN_SW32B, hb, 16, 0, true);
  for (int hb = 0; hb < 2; ++hb) run_case("A=K B=MN32B col_off 8", 16, 32, FMT_K_INTER, 0, FMT_MN_SW32B, hb, 8, 0, true);
  for (int hb = 0; hb < 2; ++hb) run_case("A=K B=MN32B col_off 24 chk 8", 16, 32, FMT_K_INTER, 0, FMT_MN_SW32B, hb, 24, 0, true, 8);
  run_case("random fp32 K/K", 128, 16, FMT_K_INTER, 0, FMT_K_INTER, 0, 0, 0, false);
  return 0;
}
