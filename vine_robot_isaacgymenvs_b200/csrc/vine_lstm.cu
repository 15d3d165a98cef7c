// vine_lstm.cu — the pointwise half of the LSTM layer of the reference's actor-critic (Vine5LinkMovingBasePPO.yaml:32-38:
// lstm, 256 units, 1 layer) as two fused, HBM-bound kernels per time step (forward / backward), replacing ~45 elementwise
// torch kernels per step of the truncated-BPTT loop (rl_games LSTMWithDones; torch gate order i, f, g, o).
//
//   forward  (step t):  c_in = c_prev * nd_t ;  i,f,o = sigmoid, g = tanh of the pre-activations G[s, 4H]
//                       c = f c_in + i g ;  h = o tanh(c)
//                       outputs: c f32, h f32 (-> LayerNorm), h_next bf16 = h * nd_{t+1} (operand of the next recurrent GEMM),
//                                acts bf16 [S, 4H] = (i, f, g, o) saved for the backward pass
//   backward (step t):  dh = dh_out + dh_rec * nd_{t+1} ;  dc = dc_next + dh o (1 - tanh^2 c)
//                       dG = (dc g i(1-i), dc c_in f(1-f), dc i (1-g^2), dh tanh(c) o(1-o))  bf16 [S, 4H]
//                       dc_prev = dc f nd_t
// Bytes per (sequence, unit) and step: forward 4*2 (G bf16) + 4 + 4 in, 4 + 4 + 2 + 8 out = 34 B; backward 8 + 4 + 4 + 4 +
// 4 + 4 in, 8 + 4 out = 40 B  -> roofline: HBM.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vine_b200.h"

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float tanhf_(float x) {
  const float e = __expf(-2.f * fabsf(x));
  const float t = (1.f - e) / (1.f + e);
  return x < 0.f ? -t : t;
}
__device__ __forceinline__ float2 ld2(const __nv_bfloat16* p) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}
__device__ __forceinline__ void st2(__nv_bfloat16* p, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}

// one thread = 2 adjacent hidden units of one sequence
__global__ void __launch_bounds__(256) lstm_cell_fwd_kernel(const __nv_bfloat16* __restrict__ G, const float* __restrict__ c_prev,
                                                            const float* __restrict__ nd, const float* __restrict__ nd_next,
                                                            int64_t S, int H, float* __restrict__ c, float* __restrict__ h,
                                                            __nv_bfloat16* __restrict__ h_next, __nv_bfloat16* __restrict__ acts) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int half = H / 2;
  if (idx >= S * half) return;
  const int64_t s = idx / half;
  const int j = (int)(idx % half) * 2;
  const __nv_bfloat16* g = G + s * 4 * H;
  const float2 pi = ld2(g + j), pf = ld2(g + H + j), pg = ld2(g + 2 * H + j), po = ld2(g + 3 * H + j);
  const float m = nd ? nd[s] : 1.f;
  const float2 cp = *reinterpret_cast<const float2*>(c_prev + s * H + j);
  const float i0 = sigmoidf_(pi.x), i1 = sigmoidf_(pi.y), f0 = sigmoidf_(pf.x), f1 = sigmoidf_(pf.y);
  const float g0 = tanhf_(pg.x), g1 = tanhf_(pg.y), o0 = sigmoidf_(po.x), o1 = sigmoidf_(po.y);
  const float c0 = fmaf(f0, cp.x * m, i0 * g0), c1 = fmaf(f1, cp.y * m, i1 * g1);
  const float h0 = o0 * tanhf_(c0), h1 = o1 * tanhf_(c1);
  *reinterpret_cast<float2*>(c + s * H + j) = make_float2(c0, c1);
  *reinterpret_cast<float2*>(h + s * H + j) = make_float2(h0, h1);
  if (h_next) {
    const float mn = nd_next ? nd_next[s] : 1.f;
    st2(h_next + s * H + j, h0 * mn, h1 * mn);
  }
  if (acts) {
    __nv_bfloat16* a = acts + s * 4 * H;
    st2(a + j, i0, i1); st2(a + H + j, f0, f1); st2(a + 2 * H + j, g0, g1); st2(a + 3 * H + j, o0, o1);
  }
}

__global__ void __launch_bounds__(256) lstm_cell_bwd_kernel(const __nv_bfloat16* __restrict__ acts, const float* __restrict__ c_prev,
                                                            const float* __restrict__ c, const float* __restrict__ nd,
                                                            const float* __restrict__ dh_out, const __nv_bfloat16* __restrict__ dh_rec,
                                                            const float* __restrict__ nd_next, const float* __restrict__ dc_next,
                                                            int64_t S, int H, __nv_bfloat16* __restrict__ dG,
                                                            float* __restrict__ dc_prev) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int half = H / 2;
  if (idx >= S * half) return;
  const int64_t s = idx / half;
  const int j = (int)(idx % half) * 2;
  const __nv_bfloat16* a = acts + s * 4 * H;
  const float2 vi = ld2(a + j), vf = ld2(a + H + j), vg = ld2(a + 2 * H + j), vo = ld2(a + 3 * H + j);
  const float m = nd ? nd[s] : 1.f;
  const float2 cp = *reinterpret_cast<const float2*>(c_prev + s * H + j);
  const float2 cc = *reinterpret_cast<const float2*>(c + s * H + j);
  float2 dh = *reinterpret_cast<const float2*>(dh_out + s * H + j);
  if (dh_rec) {
    const float mn = nd_next ? nd_next[s] : 1.f;
    const float2 r = ld2(dh_rec + s * H + j);
    dh.x = fmaf(r.x, mn, dh.x), dh.y = fmaf(r.y, mn, dh.y);
  }
  float2 dc = dc_next ? *reinterpret_cast<const float2*>(dc_next + s * H + j) : make_float2(0.f, 0.f);
  const float t0 = tanhf_(cc.x), t1 = tanhf_(cc.y);
  dc.x = fmaf(dh.x * vo.x, 1.f - t0 * t0, dc.x), dc.y = fmaf(dh.y * vo.y, 1.f - t1 * t1, dc.y);
  __nv_bfloat16* d = dG + s * 4 * H;
  st2(d + j, dc.x * vg.x * vi.x * (1.f - vi.x), dc.y * vg.y * vi.y * (1.f - vi.y));
  st2(d + H + j, dc.x * cp.x * m * vf.x * (1.f - vf.x), dc.y * cp.y * m * vf.y * (1.f - vf.y));
  st2(d + 2 * H + j, dc.x * vi.x * (1.f - vg.x * vg.x), dc.y * vi.y * (1.f - vg.y * vg.y));
  st2(d + 3 * H + j, dh.x * t0 * vo.x * (1.f - vo.x), dh.y * t1 * vo.y * (1.f - vo.y));
  *reinterpret_cast<float2*>(dc_prev + s * H + j) = make_float2(dc.x * vf.x * m, dc.y * vf.y * m);
}

}  // namespace

extern "C" {

int vine_lstm_cell_fwd(const void* gates, const float* c_prev, const float* not_done, const float* not_done_next, int64_t num_seqs,
                       int hidden, float* c, float* h, void* h_next_bf16, void* acts_bf16, void* stream) {
  if (!gates || !c_prev || !c || !h || num_seqs <= 0 || hidden < 2 || (hidden & 1)) return VINE_ERR_INVALID_ARG;
  const int64_t n = num_seqs * (hidden / 2);
  lstm_cell_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)gates, c_prev, not_done, not_done_next, num_seqs, hidden, c, h, (__nv_bfloat16*)h_next_bf16,
      (__nv_bfloat16*)acts_bf16);
  return cudaGetLastError() == cudaSuccess ? VINE_OK : VINE_ERR_CUDA;
}

int vine_lstm_cell_bwd(const void* acts_bf16, const float* c_prev, const float* c, const float* not_done, const float* dh_out,
                       const void* dh_rec_bf16, const float* not_done_next, const float* dc_next, int64_t num_seqs, int hidden,
                       void* dgates_bf16, float* dc_prev, void* stream) {
  if (!acts_bf16 || !c_prev || !c || !dh_out || !dgates_bf16 || !dc_prev || num_seqs <= 0 || hidden < 2 || (hidden & 1))
    return VINE_ERR_INVALID_ARG;
  const int64_t n = num_seqs * (hidden / 2);
  lstm_cell_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)acts_bf16, c_prev, c, not_done, dh_out, (const __nv_bfloat16*)dh_rec_bf16, not_done_next, dc_next, num_seqs,
      hidden, (__nv_bfloat16*)dgates_bf16, dc_prev);
  return cudaGetLastError() == cudaSuccess ? VINE_OK : VINE_ERR_CUDA;
}

}  // extern "C"
