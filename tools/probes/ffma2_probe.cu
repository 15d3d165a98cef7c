// ffma2_probe.cu — throughput of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a with distinct register operands
// (the shape of the env-step substep: no shared multiplier, no immediates). Measurement helper, not product code.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu && ./ffma2_probe
#include <cuda_runtime.h>
#include <stdio.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float s) {
  // 8 independent accumulator pairs; multipliers/addends are per-thread distinct registers
  float2 x[8], y[8], z[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    y[i] = make_float2(0.999f + 1e-6f * (threadIdx.x + i) * s, 0.998f + 1e-6f * i * s);
    z[i] = make_float2(1e-3f * i * s, 2e-3f * (i + threadIdx.x) * s);
  }
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) { x[i].x = fmaf(x[i].x, y[i].x, z[i].x); x[i].y = fmaf(x[i].y, y[i].y, z[i].y); }
        else if (MODE == 1) x[i] = __ffma2_rn(x[i], y[i], z[i]);
        else if (MODE == 2) {  // mixed: one FFMA2 + one ALU op (integer) per pair: does issue relief let ALU ride along?
          x[i] = __ffma2_rn(x[i], y[i], z[i]);
          z[i].x = __int_as_float(__float_as_int(z[i].x) ^ (it + i));
        } else {  // scalar + same ALU
          x[i].x = fmaf(x[i].x, y[i].x, z[i].x); x[i].y = fmaf(x[i].y, y[i].y, z[i].y);
          z[i].x = __int_as_float(__float_as_int(z[i].x) ^ (it + i));
        }
      }
    }
  }
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += x[i].x + x[i].y + z[i].x;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
double run(int blocks_per_sm) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * blocks_per_sm, threads = 256, iters = 2048;
  float* out; cudaMalloc(&out, sizeof(float) * blocks * threads);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, iters, 1.0f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double tf = 2.0 * 2 * 8 * 8 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
    if (r && tf > best) best = tf;
  }
  cudaFree(out);
  return best;
}

int main() {
  for (int b = 1; b <= 8; b *= 2) {
    printf("blocks/SM %d (warps/SM %d): FFMA %.1f  FFMA2 %.1f  FFMA2+ALU %.1f  FFMA+ALU %.1f TFLOP/s\n", b, b * 8,
           run<0>(b), run<1>(b), run<2>(b), run<3>(b));
  }
  return 0;
}
