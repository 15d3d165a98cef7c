// vine_mlp.cu — packing of the actor-critic MLP parameters for the tensor cores (sm_100a).
//
// Network of cfg/train/Vine5LinkMovingBasePPO.yaml:10-30 (rl_games `actor_critic`, MLP part):
//   x = clamp((obs - mean) / std, +-5);  h1 = ELU(W1 x + b1) [256], h2 = ELU(W2 h1 + b2) [128], h3 = ELU(W3 h2 + b3) [64]
//   mu = Wmu h3 + bmu [2],  v = Wv h3 + bv [1]
// vine_mlp_pack converts torch-layout f32 weights into ONE block (bf16 weights in the UMMA row-blocked layout of
// vine_umma.cuh + f32 biases) that a CTA pulls into shared memory with a single bulk TMA copy.  The kernels that
// consume it: vine_policy_act / vine_mlp_forward (vine_rollout.cu) and vine_ppo_minibatch (vine_ppo.cu);
// vine_ppo_adam rewrites it in place after every optimiser step.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "vine_mlp_common.cuh"

namespace {
using namespace vine_mlp;

__global__ void vine_mlp_pack_kernel(const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
                                     const float* b3, const float* wmu, const float* bmu, const float* wv, const float* bv,
                                     int num_obs, uint8_t* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  auto put = [&](int off, int row, int k, int K, float v) {
    *reinterpret_cast<__nv_bfloat16*>(out + off + tile_offset(row, k, K)) = __float2bfloat16_rn(v);
  };
  if (i < H1 * K1) { const int r = i / K1, k = i % K1; put(OFF_W1, r, k, K1, k < num_obs ? w1[r * num_obs + k] : 0.f); }
  if (i < H2 * H1) { const int r = i / H1, k = i % H1; put(OFF_W2, r, k, H1, w2[r * H1 + k]); }
  if (i < H3 * H2) { const int r = i / H2, k = i % H2; put(OFF_W3, r, k, H2, w3[r * H2 + k]); }
  if (i < NH * H3) {
    const int r = i / H3, k = i % H3;
    put(OFF_W4, r, k, H3, r < 2 ? wmu[r * H3 + k] : (r == 2 ? wv[k] : 0.f));
  }
  float* bias = reinterpret_cast<float*>(out + OFF_B);
  if (i < H1) bias[i] = b1[i];
  if (i < H2) bias[H1 + i] = b2[i];
  if (i < H3) bias[H1 + H2 + i] = b3[i];
  if (i < NH) bias[H1 + H2 + H3 + i] = i < 2 ? bmu[i] : (i == 2 ? bv[0] : 0.f);
}

}  // namespace

extern "C" {

int vine_mlp_pack(const float* w1, const float* b1, const float* w2, const float* b2, const float* w3, const float* b3,
                  const float* w_mu, const float* b_mu, const float* w_v, const float* b_v, int num_obs, void* packed,
                  void* stream) {
  if (!w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !w_mu || !b_mu || !w_v || !b_v || !packed || num_obs < 1 || num_obs > K1)
    return VINE_ERR_INVALID_ARG;
  const int total = H2 * H1;
  vine_mlp_pack_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w1, b1, w2, b2, w3, b3, w_mu, b_mu, w_v, b_v,
                                                                             num_obs, (uint8_t*)packed);
  return cudaGetLastError() == cudaSuccess ? VINE_OK : VINE_ERR_CUDA;
}

}  // extern "C"
