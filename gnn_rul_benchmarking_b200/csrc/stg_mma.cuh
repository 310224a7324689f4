// Warp-level tensor-core helpers for the graph-conv block kernels: mma.sync m16n8k8 TF32 with the
// 3xTF32 error-compensated split (a = a_hi + a_lo, a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi), which
// keeps fp32-level accuracy (parity tolerance 1e-4 on outputs, BASELINE.json) while the per-graph
// products (Gram, aggregation, their transposes) run on the tensor pipe instead of the FP32 pipe.
//
// Fragment layouts (PTX ISA, m16n8k8 .tf32):  g = lane >> 2, t = lane & 3
//   A (16x8, row):  a0=A[g][t]  a1=A[g+8][t]  a2=A[g][t+4]  a3=A[g+8][t+4]
//   B (8x8,  col):  b0=B[t][g]  b1=B[t+4][g]
//   C (16x8):       c0=C[g][2t] c1=C[g][2t+1] c2=C[g+8][2t] c3=C[g+8][2t+1]
#pragma once
#include "stg_common.cuh"

namespace stg {

STG_DEVINL uint32_t f2tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

STG_DEVINL void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = f2tf32(x);
  lo = f2tf32(x - __uint_as_float(hi));
}

STG_DEVINL void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

struct FragA { uint32_t hi[4], lo[4]; };
struct FragB { uint32_t hi[2], lo[2]; };

STG_DEVINL FragA make_a(float a0, float a1, float a2, float a3) {
  FragA f;
  split_tf32(a0, f.hi[0], f.lo[0]);
  split_tf32(a1, f.hi[1], f.lo[1]);
  split_tf32(a2, f.hi[2], f.lo[2]);
  split_tf32(a3, f.hi[3], f.lo[3]);
  return f;
}
STG_DEVINL FragB make_b(float b0, float b1) {
  FragB f;
  split_tf32(b0, f.hi[0], f.lo[0]);
  split_tf32(b1, f.hi[1], f.lo[1]);
  return f;
}

// d += a * b with 3xTF32 (small terms first)
STG_DEVINL void mma3(float (&d)[4], const FragA& a, const FragB& b) {
  mma_tf32(d, a.lo, b.hi);
  mma_tf32(d, a.hi, b.lo);
  mma_tf32(d, a.hi, b.hi);
}

}  // namespace stg
