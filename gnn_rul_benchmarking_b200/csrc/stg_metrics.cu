// Evaluation metrics of the reference (utils.py:136-169) as one reduction kernel (sm_100a):
//   out[0] += sum Score_v1 terms   exp((real-pred)*max_rul/13)-1 if real > pred else exp((pred-real)*max_rul/10)-1
//   out[1] += sum Score_v2 terms   err = (real-pred)/(real+1e-8)*100;  err<=0: exp(-ln(.5)*err/5) else exp(ln(.5)*err/20)
//   out[2] += sum |pred-real|      out[3] += sum (pred-real)^2          (double accumulators)
#include <math.h>

#include "../../include/stgconv_b200.h"
#include "stg_common.cuh"

namespace stg {
namespace {
__global__ void __launch_bounds__(256) k_metrics(const float* __restrict__ pred, const float* __restrict__ real,
                                                 long long n, float max_rul, double* __restrict__ out) {
  __shared__ double red[4][8];
  double s1 = 0.0, s2 = 0.0, sa = 0.0, sq = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double p = pred[i], r = real[i];
    if (r > p) s1 += exp((r * max_rul - p * max_rul) / 13.0) - 1.0;
    else s1 += exp((p * max_rul - r * max_rul) / 10.0) - 1.0;
    const double err = ((r - p) / (r + 1e-8)) * 100.0;
    if (err <= 0.0) s2 += exp(-log(0.5) * (err / 5.0));
    else s2 += exp(log(0.5) * (err / 20.0));
    sa += fabs(p - r);
    sq += (p - r) * (p - r);
  }
  double v[4] = {s1, s2, sa, sq};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
    if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = v[q];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
    atomicAdd(&out[threadIdx.x], t);
  }
}
}  // namespace
}  // namespace stg

using namespace stg;

extern "C" int stg_metrics(const float* pred_dev, const float* real_dev, int64_t n, float max_rul, double* out4_dev,
                           void* stream) {
  if (!pred_dev || !real_dev || !out4_dev || n < 1) return set_err(STG_ERR_INVALID, "bad argument");
  int grid = (int)((n + 255) / 256);
  if (grid > 592) grid = 592;
  k_metrics<<<grid, 256, 0, (cudaStream_t)stream>>>(pred_dev, real_dev, (long long)n, max_rul, out4_dev);
  return check_cuda("stg_metrics");
}
