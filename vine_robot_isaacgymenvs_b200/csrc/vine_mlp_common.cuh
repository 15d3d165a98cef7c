// vine_mlp_common.cuh — what the actor-critic MLP kernels share: the packed parameter block (bf16 weights in the UMMA
// row-blocked layout + f32 biases; written by vine_mlp_pack and vine_ppo_adam) and the per-row TMEM epilogues.
// Network: cfg/train/Vine5LinkMovingBasePPO.yaml:10-30 -- MLP [256,128,64] ELU, heads mu[2] + value[1] (padded to 16).
#pragma once
#include "../../include/vine_b200.h"
#include "vine_umma.cuh"

namespace vine_mlp {
using namespace vine_umma;

constexpr int H1 = 256, H2 = 128, H3 = 64, NH = 16, K1 = 32;
constexpr int TILE = 128, THREADS = 256;
// packed parameter block: identical to vine_mlp.cu (vine_mlp_pack / vine_mlp_forward share it)
constexpr int OFF_W1 = 0;
constexpr int OFF_W2 = OFF_W1 + H1 * K1 * 2;
constexpr int OFF_W3 = OFF_W2 + H2 * H1 * 2;
constexpr int OFF_W4 = OFF_W3 + H3 * H2 * 2;
constexpr int OFF_B = OFF_W4 + NH * H3 * 2;  // f32: b1[256] b2[128] b3[64] bh[16]
constexpr int PACKED_BYTES = OFF_B + (H1 + H2 + H3 + NH) * 4;
static_assert(PACKED_BYTES == VINE_MLP_PACKED_BYTES, "header constant out of date");
// shared-memory map common to the kernels: [packed block][x][h1][h2][h3]
constexpr int OFF_X = 102400;                      // x   [128 x 32]  bf16
constexpr int OFF_A1 = OFF_X + TILE * K1 * 2;      // h1 [128 x 256]
constexpr int OFF_A2 = OFF_A1 + TILE * H1 * 2;     // h2 [128 x 128]
constexpr int OFF_A3 = OFF_A2 + TILE * H2 * 2;     // h3 [128 x 64]
constexpr int OFF_END = OFF_A3 + TILE * H3 * 2;
static_assert(PACKED_BYTES <= OFF_X, "packed block overlaps the activation tiles");

__device__ __forceinline__ float elu(float x) { return x > 0.f ? x : __expf(x) - 1.f; }

// accumulator columns [taddr, taddr+ncols) -> bias + ELU -> bf16 -> tile columns [c_out, c_out+ncols) of row `row`
template <int KL>
__device__ __forceinline__ void fwd_epilogue(uint32_t taddr, int ncols, int c_out, const float* bias, uint8_t* tile, int row) {
#pragma unroll 1
  for (int c0 = 0; c0 < ncols; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(taddr + c0, r);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint32_t w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = c_out + c0 + g * 8 + 2 * i;
        w[i] = pack_bf16(elu(__uint_as_float(r[g * 8 + 2 * i]) + bias[c]), elu(__uint_as_float(r[g * 8 + 2 * i + 1]) + bias[c + 1]));
      }
      *reinterpret_cast<uint4*>(tile + tile_offset(row, c_out + c0 + g * 8, KL)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

// dz = dh * ELU'(h) with ELU'(h) = h > 0 ? 1 : h + 1, written over h in place
template <int KL>
__device__ __forceinline__ void bwd_epilogue(uint32_t taddr, int ncols, int c_out, uint8_t* tile, int row) {
#pragma unroll 1
  for (int c0 = 0; c0 < ncols; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(taddr + c0, r);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint4* p = reinterpret_cast<uint4*>(tile + tile_offset(row, c_out + c0 + g * 8, KL));
      const uint4 hv = *p;
      const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
      uint32_t w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 h = unpack_bf16(hw[i]);
        const float d0 = __uint_as_float(r[g * 8 + 2 * i]) * (h.x > 0.f ? 1.f : h.x + 1.f);
        const float d1 = __uint_as_float(r[g * 8 + 2 * i + 1]) * (h.y > 0.f ? 1.f : h.y + 1.f);
        w[i] = pack_bf16(d0, d1);
      }
      *p = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

__device__ __forceinline__ float bf16_at(const uint8_t* tile, int row, int col, int KL) {
  return __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(tile + tile_offset(row, col, KL)));
}

// normalised observation row -> bf16 x tile (this thread: 16 of the 32 padded columns of its row); optionally copies
// the raw observation out (rollout buffer) and plants the constant 1 in column 31 (bias gradient through dW1)
__device__ __forceinline__ void build_x_tile(uint8_t* x_t, int row, int half, bool valid, const float* __restrict__ obs_row,
                                             const float* __restrict__ mean, const float* __restrict__ inv_std, int O,
                                             bool ones_column, float* __restrict__ obs_copy_row) {
  const int k0 = half * 16;
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int k = k0 + i;
    float v = 0.f;
    if (valid && k < O) {
      const float raw = obs_row[k];
      if (obs_copy_row) obs_copy_row[k] = raw;
      v = fminf(fmaxf((raw - mean[k]) * inv_std[k], -5.f), 5.f);
    }
    if (ones_column && valid && k == K1 - 1) v = 1.f;
    x[i] = v;
  }
#pragma unroll
  for (int g = 0; g < 2; ++g)
    *reinterpret_cast<uint4*>(x_t + tile_offset(row, k0 + g * 8, K1)) =
        make_uint4(pack_bf16(x[g * 8], x[g * 8 + 1]), pack_bf16(x[g * 8 + 2], x[g * 8 + 3]),
                   pack_bf16(x[g * 8 + 4], x[g * 8 + 5]), pack_bf16(x[g * 8 + 6], x[g * 8 + 7]));
}

}  // namespace vine_mlp
