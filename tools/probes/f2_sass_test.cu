#include <cuda_runtime.h>
struct P { float2 k[4]; float s; };
__global__ void t(const __grid_constant__ P p, float2* io) {
  float2 a = io[threadIdx.x], b = io[threadIdx.x + 32], c = io[threadIdx.x + 64];
  // sub via neg
  float2 r1 = __fadd2_rn(a, make_float2(-b.x, -b.y));
  // fma with negated multiplicand
  float2 r2 = __ffma2_rn(make_float2(-a.x, -a.y), b, c);
  // constant-bank pair operand
  float2 r3 = __ffma2_rn(a, p.k[1], c);
  // broadcast scalar constant
  float2 r4 = __fmul2_rn(b, make_float2(p.s, p.s));
  // immediate
  float2 r5 = __ffma2_rn(c, make_float2(-0.5f, -0.5f), make_float2(1.f, 1.f));
  // mixed halves: (a.x, b.y)
  float2 r6 = __fmul2_rn(make_float2(a.x, b.y), make_float2(c.y, c.x));
  io[threadIdx.x] = __fadd2_rn(__fadd2_rn(__fadd2_rn(r1, r2), __fadd2_rn(r3, r4)), __fadd2_rn(r5, r6));
}
