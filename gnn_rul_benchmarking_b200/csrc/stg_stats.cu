// Hand-crafted per-patch statistics of the bearing models' parameter-free prefix:
// segment_and_compute_features, models/ST_GCN/Model.py:7-52 -- for every patch (row of P samples):
//   max, min, peak-to-peak, variance (unbiased), std (unbiased), mean, rms, mean |x|,
//   skewness mean(((x-mean)/std)^3), excess kurtosis mean(((x-mean)/std)^4) - 3.
// One warp per patch, two passes over the row held in registers / re-read from L1 (sm_100a).
// No learnable parameter lies upstream, so there is no backward.
#include <math.h>

#include "../../include/stgconv_b200.h"
#include "stg_common.cuh"

namespace stg {
namespace {
__global__ void __launch_bounds__(256) k_patch_stats(const float* __restrict__ x, long long R, int P,
                                                     float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  const float* xr = x + row * P;
  float mx = -INFINITY, mn = INFINITY, s = 0.f, sq = 0.f, sa = 0.f;
  for (int i = lane; i < P; i += 32) {
    const float v = xr[i];
    mx = fmaxf(mx, v); mn = fminf(mn, v);
    s += v; sq = fmaf(v, v, sq); sa += fabsf(v);
  }
  mx = warp_max(mx);
  mn = -warp_max(-mn);
  s = warp_sum(s); sq = warp_sum(sq); sa = warp_sum(sa);
  const float mean = s / (float)P;
  float m2 = 0.f;
  for (int i = lane; i < P; i += 32) { const float d = xr[i] - mean; m2 = fmaf(d, d, m2); }
  m2 = warp_sum(m2);
  const float var = m2 / (float)(P - 1), sd = sqrtf(var);       // torch.var / torch.std default: unbiased
  float m3 = 0.f, m4 = 0.f;
  for (int i = lane; i < P; i += 32) {
    const float z = (xr[i] - mean) / sd, z2 = z * z;
    m3 = fmaf(z2, z, m3); m4 = fmaf(z2, z2, m4);
  }
  m3 = warp_sum(m3); m4 = warp_sum(m4);
  if (lane == 0) {
    float* o = out + row * 10;
    o[0] = mx; o[1] = mn; o[2] = mx - mn; o[3] = var; o[4] = sd; o[5] = mean;
    o[6] = sqrtf(sq / (float)P); o[7] = sa / (float)P; o[8] = m3 / (float)P; o[9] = m4 / (float)P - 3.f;
  }
}

// extract_features, models/GAT_LSTM/Model.py:6-70 -- 11 statistics per patch:
//   mean, std (unbiased), (mean sqrt|x|)^2, rms, (max-min)/2, skewness m/((m-1)(m-2)) sum d^3 / std^3,
//   kurtosis (m(m+1)-3(m-1)^3)/((m-1)(m-2)(m-3)) sum d^4 / std^4, crest max|x|/rms, clearance max|x|/(mean sqrt|x|)^2,
//   shape rms/mean|x|, impulse max|x|/mean|x|.
__global__ void __launch_bounds__(256) k_patch_stats11(const float* __restrict__ x, long long R, int P,
                                                       float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  const float* xr = x + row * P;
  float mx = -INFINITY, mn = INFINITY, ma = 0.f, s = 0.f, sq = 0.f, sa = 0.f, sr = 0.f;
  for (int i = lane; i < P; i += 32) {
    const float v = xr[i], av = fabsf(v);
    mx = fmaxf(mx, v); mn = fminf(mn, v); ma = fmaxf(ma, av);
    s += v; sq = fmaf(v, v, sq); sa += av; sr += sqrtf(av);
  }
  mx = warp_max(mx); mn = -warp_max(-mn); ma = warp_max(ma);
  s = warp_sum(s); sq = warp_sum(sq); sa = warp_sum(sa); sr = warp_sum(sr);
  const float m = (float)P, mean = s / m;
  float m2 = 0.f, m3 = 0.f, m4 = 0.f;
  for (int i = lane; i < P; i += 32) {
    const float d = xr[i] - mean, d2 = d * d;
    m2 += d2; m3 = fmaf(d2, d, m3); m4 = fmaf(d2, d2, m4);
  }
  m2 = warp_sum(m2); m3 = warp_sum(m3); m4 = warp_sum(m4);
  if (lane == 0) {
    const float sd = sqrtf(m2 / (m - 1.f)), rms = sqrtf(sq / m), rsa = (sr / m) * (sr / m), mabs = sa / m;
    const float cs = m / ((m - 1.f) * (m - 2.f));
    const float ck = (m * (m + 1.f) - 3.f * (m - 1.f) * (m - 1.f) * (m - 1.f)) / ((m - 1.f) * (m - 2.f) * (m - 3.f));
    float* o = out + row * 11;
    o[0] = mean; o[1] = sd; o[2] = rsa; o[3] = rms; o[4] = 0.5f * (mx - mn);
    o[5] = cs * m3 / (sd * sd * sd); o[6] = ck * m4 / (sd * sd * sd * sd);
    o[7] = ma / rms; o[8] = ma / rsa; o[9] = rms / mabs; o[10] = ma / mabs;
  }
}

// extract_temporal_features, models/SAGCN/Model.py:21-38 (also AGCN_TF) -- 12 statistics per patch:
//   max, min, std (unbiased), rms, mean, ptp, var (unbiased), entropy of softmax(x), std(asin(clamp(x))),
//   std(atan(x)), kurtosis mean(d^4)/std^4 - 3, skewness mean(d^3)/std^3.
__global__ void __launch_bounds__(256) k_patch_stats12(const float* __restrict__ x, long long R, int P,
                                                       float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  const float* xr = x + row * P;
  float mx = -INFINITY, mn = INFINITY, s = 0.f, sq = 0.f, sas = 0.f, sat = 0.f;
  for (int i = lane; i < P; i += 32) {
    const float v = xr[i];
    mx = fmaxf(mx, v); mn = fminf(mn, v);
    s += v; sq = fmaf(v, v, sq);
    sas += asinf(fminf(fmaxf(v, -0.9999999f), 0.9999999f));
    sat += atanf(v);
  }
  mx = warp_max(mx); mn = -warp_max(-mn);
  s = warp_sum(s); sq = warp_sum(sq); sas = warp_sum(sas); sat = warp_sum(sat);
  const float m = (float)P, mean = s / m, mas = sas / m, mat = sat / m;
  float m2 = 0.f, m3 = 0.f, m4 = 0.f, vas = 0.f, vat = 0.f, se = 0.f;
  for (int i = lane; i < P; i += 32) {
    const float v = xr[i], d = v - mean, d2 = d * d;
    m2 += d2; m3 = fmaf(d2, d, m3); m4 = fmaf(d2, d2, m4);
    const float da = asinf(fminf(fmaxf(v, -0.9999999f), 0.9999999f)) - mas, dt = atanf(v) - mat;
    vas = fmaf(da, da, vas); vat = fmaf(dt, dt, vat);
    se += expf(v - mx);
  }
  m2 = warp_sum(m2); m3 = warp_sum(m3); m4 = warp_sum(m4);
  vas = warp_sum(vas); vat = warp_sum(vat); se = warp_sum(se);
  // entropy = -sum p log p with log p_i = x_i - mx - log(se)
  const float lse = logf(se);
  float ent = 0.f;
  for (int i = lane; i < P; i += 32) {
    const float lp = xr[i] - mx - lse;
    ent = fmaf(-expf(lp), lp, ent);
  }
  ent = warp_sum(ent);
  if (lane == 0) {
    const float var = m2 / (m - 1.f), sd = sqrtf(var);
    float* o = out + row * 12;
    o[0] = mx; o[1] = mn; o[2] = sd; o[3] = sqrtf(sq / m); o[4] = mean; o[5] = mx - mn; o[6] = var; o[7] = ent;
    o[8] = sqrtf(vas / (m - 1.f)); o[9] = sqrtf(vat / (m - 1.f));
    o[10] = (m4 / m) / (sd * sd * sd * sd) - 3.f; o[11] = (m3 / m) / (sd * sd * sd);
  }
}
}  // namespace
}  // namespace stg

using namespace stg;

extern "C" int stg_patch_stats12(const float* x_dev, int64_t R, int P, float* out_dev, void* stream) {
  if (!x_dev || !out_dev || R < 1 || P < 2) return set_err(STG_ERR_INVALID, "bad argument (patches need >= 2 samples)");
  const long long grid = (R + 7) / 8;
  k_patch_stats12<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(x_dev, (long long)R, P, out_dev);
  return check_cuda("stg_patch_stats12");
}

extern "C" int stg_patch_stats11(const float* x_dev, int64_t R, int P, float* out_dev, void* stream) {
  if (!x_dev || !out_dev || R < 1 || P < 4) return set_err(STG_ERR_INVALID, "bad argument (patches need >= 4 samples)");
  const long long grid = (R + 7) / 8;
  k_patch_stats11<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(x_dev, (long long)R, P, out_dev);
  return check_cuda("stg_patch_stats11");
}

extern "C" int stg_patch_stats(const float* x_dev, int64_t R, int P, float* out_dev, void* stream) {
  if (!x_dev || !out_dev || R < 1 || P < 2) return set_err(STG_ERR_INVALID, "bad argument (patches need >= 2 samples)");
  const long long grid = (R + 7) / 8;
  k_patch_stats<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(x_dev, (long long)R, P, out_dev);
  return check_cuda("stg_patch_stats");
}
