// Throughput probe behind the design of csrc/stg_rnn.cu: cycles per CTA-wide round of (a) broadcast LDS.128 / LDS.32,
// (b) mma.sync m16n8k8 tf32, (c) FFMA2, with 1..8 warps resident on one SM.   nvcc -arch=sm_100a -o rnn_probe rnn_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_lds128(const float* in, float* out, long long* cyc, int iters) {
  __shared__ float4 s[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s[i] = make_float4(in[i], 1.f, 2.f, 3.f);
  __syncthreads();
  float4 acc = make_float4(0, 0, 0, 0);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 64; ++j) {
      float4 v = s[(j * 3 + it) & 255];           // warp-uniform address: broadcast
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_lds32(const float* in, float* out, long long* cyc, int iters) {
  __shared__ float s[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = in[i & 255];
  __syncthreads();
  float acc = 0, acc2 = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 64; j += 2) {
      acc += s[(j * 5 + it) & 1023];
      acc2 += s[(j * 7 + it + 1) & 1023];
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = acc + acc2;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_mma(const float* in, float* out, long long* cyc, int iters) {
  unsigned a[4][4], b[2];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) a[i][j] = __float_as_uint(in[(threadIdx.x + i * 4 + j) & 255]);
  b[0] = __float_as_uint(in[threadIdx.x & 255]); b[1] = __float_as_uint(in[(threadIdx.x + 9) & 255]);
  float c[4][4] = {};
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int i = j & 3;
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[i][0]), "r"(a[i][1]), "r"(a[i][2]), "r"(a[i][3]), "r"(b[0]), "r"(b[1]));
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_ffma2(const float* in, float* out, long long* cyc, int iters) {
  unsigned long long acc[8], w, h;
  float f = in[threadIdx.x & 255];
  asm("mov.b64 %0, {%1,%1};" : "=l"(w) : "f"(f));
  asm("mov.b64 %0, {%1,%1};" : "=l"(h) : "f"(f * 0.5f));
  for (int i = 0; i < 8; ++i) acc[i] = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 64; ++j) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[j & 7]) : "l"(w), "l"(h));
  }
  long long t1 = clock64();
  unsigned long long s = 0;
  for (int i = 0; i < 8; ++i) s ^= acc[i];
  out[threadIdx.x] = (float)(s & 0xffff);
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float *in, *out; long long* cyc, h;
  cudaMalloc(&in, 4096); cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
  cudaMemset(in, 0, 4096);
  const int iters = 2000;
  for (int warps = 1; warps <= 8; warps *= 2) {
    double r[4];
    for (int k = 0; k < 4; ++k) {
      for (int rep = 0; rep < 2; ++rep) {
        if (k == 0) k_lds128<<<1, warps * 32>>>(in, out, cyc, iters);
        if (k == 1) k_lds32<<<1, warps * 32>>>(in, out, cyc, iters);
        if (k == 2) k_mma<<<1, warps * 32>>>(in, out, cyc, iters);
        if (k == 3) k_ffma2<<<1, warps * 32>>>(in, out, cyc, iters);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      r[k] = (double)h / iters;
    }
    printf("warps %d: cycles per warp-instruction per SM: LDS.128 bcast %.2f | LDS.32 bcast %.2f | mma.m16n8k8.tf32 %.2f | FFMA2 %.2f\n",
           warps, r[0] / 64 / warps, r[1] / 64 / warps, r[2] / 16 / warps, r[3] / 64 / warps);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
