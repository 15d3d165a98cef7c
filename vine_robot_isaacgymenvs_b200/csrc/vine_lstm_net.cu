// vine_lstm_net.cu — the recurrent half of the reference's actor-critic on hand-written sm_100a kernels:
//   u = [MLP(x) (64) | x (32, padded)]  ->  LSTM 256 (gates = u W_ih^T + h W_hh^T + b, torch order i,f,g,o)  ->  LayerNorm
//   ->  mu (2), value (1)                                   (cfg/train/Vine5LinkMovingBasePPO.yaml:10-40, rl_games A2CBuilder)
//
// Data lives in HBM as 128-row tiles in the UMMA row-blocked layout (vine_umma.cuh), so every GEMM operand is one bulk-TMA
// copy and the same bytes serve forward (K-major) and backward/weight-gradient (MN-major) GEMMs:
//   U  [tile]        [128 x 128] bf16: cols 0..63 MLP output, 64..95 normalised observation (col 95 == 1), 96..127 zero
//   HM [tile][half]  [128 x 128] bf16: recurrent input of the step = not_done * h_prev (hidden units 128*half ..)
//   HH [tile][half]  [128 x 128] bf16: h of the step (input of LayerNorm and, masked, of the next step)
//   ACT[tile][piece] [128 x 64]  bf16: activated gates of 16 hidden units: [i(16) f(16) g(16) o(16)]   (saved for backward)
//   C  [n, 256] f32 row-major cell state
// Packed LSTM parameters (vine_lstm_pack): W_ih as 16 pieces [64 x 128] (piece p row g*16+k = gate g of hidden unit 16p+k,
// columns = U's columns), W_hh as [2 halves][16 pieces][64 x 128], then f32: bias (same row order), LayerNorm gamma/beta,
// head weights [3][256] (mu0, mu1, v) and head biases.
//
// vine_lstm_step  : one LSTM time step for every 128-sequence tile; CTA = (tile, slice of 128 gate rows = 32 hidden units):
//                   6 bulk copies (192 KB) -> 22 tcgen05.mma (M128 N128 K16) into TMEM -> cell epilogue per row.
// vine_lstm_head  : LayerNorm + heads per row (one warp per row) + Gaussian sampling / neglogp (rollout).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "vine_device.cuh"
#include "vine_mlp_common.cuh"

namespace {
using namespace vine_mlp;

constexpr int HID = 256, GATES = 4 * HID, UK = 128;            // U tile width
constexpr int PIECE_ROWS = 64, NPIECE = GATES / PIECE_ROWS;    // 16 pieces of 16 hidden units
constexpr int TILE_BYTES = TILE * UK * 2;                      // 32 KB: one [128 x 128] bf16 tile
constexpr int PIECE_BYTES = PIECE_ROWS * UK * 2;               // 16 KB: one [64 x 128] weight piece
constexpr int ACT_BYTES = TILE * PIECE_ROWS * 2;               // 16 KB: one [128 x 64] gate tile
constexpr int LP_WIH = 0;
constexpr int LP_WHH = LP_WIH + NPIECE * PIECE_BYTES;          // [half][piece]
constexpr int LP_BIAS = LP_WHH + 2 * NPIECE * PIECE_BYTES;     // f32 [1024] (piece row order)
constexpr int LP_LNG = LP_BIAS + GATES * 4;
constexpr int LP_LNB = LP_LNG + HID * 4;
constexpr int LP_WH = LP_LNB + HID * 4;                        // f32 [3][256]
constexpr int LP_BH = LP_WH + 3 * HID * 4;                     // f32 [4]
constexpr int LP_BYTES = LP_BH + 16;
static_assert(LP_BYTES == VINE_LSTM_PACKED_BYTES, "header constant out of date");

// packed gate row (piece, r) <-> torch gate row
__host__ __device__ inline int torch_gate_row(int packed_row) {
  const int p = packed_row / PIECE_ROWS, r = packed_row % PIECE_ROWS;
  return (r / 16) * HID + 16 * p + (r % 16);
}

__global__ void vine_lstm_pack_kernel(const float* __restrict__ w_ih, const float* __restrict__ w_hh, const float* __restrict__ b_ih,
                                      const float* __restrict__ b_hh, const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                                      const float* __restrict__ w_mu, const float* __restrict__ b_mu, const float* __restrict__ w_v,
                                      const float* __restrict__ b_v, int O, uint8_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int in_w = H3 + O;   // torch W_ih is [1024, 64 + O]: MLP output first, then the observation (concat_input)
  if (i < GATES * UK) {      // W_ih pieces
    const int R = i / UK, c = i % UK, tr = torch_gate_row(R);
    float v = 0.f;
    if (c < H3) v = w_ih[tr * in_w + c];
    else if (c - H3 < O) v = w_ih[tr * in_w + c];
    *reinterpret_cast<__nv_bfloat16*>(out + LP_WIH + (R / PIECE_ROWS) * PIECE_BYTES + tile_offset(R % PIECE_ROWS, c, UK)) =
        __float2bfloat16_rn(v);
  }
  if (i < GATES * HID) {     // W_hh [half][piece]
    const int R = i / HID, c = i % HID, tr = torch_gate_row(R);
    *reinterpret_cast<__nv_bfloat16*>(out + LP_WHH + ((c / UK) * NPIECE + R / PIECE_ROWS) * PIECE_BYTES +
                                      tile_offset(R % PIECE_ROWS, c % UK, UK)) = __float2bfloat16_rn(w_hh[tr * HID + c]);
  }
  if (i < GATES) {
    const int tr = torch_gate_row(i);
    reinterpret_cast<float*>(out + LP_BIAS)[i] = b_ih[tr] + b_hh[tr];
  }
  if (i < HID) {
    reinterpret_cast<float*>(out + LP_LNG)[i] = ln_g[i];
    reinterpret_cast<float*>(out + LP_LNB)[i] = ln_b[i];
    reinterpret_cast<float*>(out + LP_WH)[i] = w_mu[i];
    reinterpret_cast<float*>(out + LP_WH)[HID + i] = w_mu[HID + i];
    reinterpret_cast<float*>(out + LP_WH)[2 * HID + i] = w_v[i];
  }
  if (i < 4) reinterpret_cast<float*>(out + LP_BH)[i] = i < 2 ? b_mu[i] : (i == 2 ? b_v[0] : 0.f);
}

// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoid_(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_(float x) {
  const float e = __expf(-2.f * fabsf(x));
  const float t = (1.f - e) / (1.f + e);
  return x < 0.f ? -t : t;
}

constexpr int SO_U = 0, SO_HM = TILE_BYTES, SO_WIH = 3 * TILE_BYTES, SO_WHH = 4 * TILE_BYTES, SO_BIAS = 6 * TILE_BYTES;
constexpr int SO_BAR = SO_BIAS + 512;
constexpr int STEP_SMEM = SO_BAR + 64;

__global__ void __launch_bounds__(THREADS, 1) vine_lstm_step_kernel(const VineLstmStep a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, row = tid & 127, half = tid >> 7;
  const int tile = blockIdx.x, slice = blockIdx.y;
  const uint32_t bar_ld = smem_u32(smem + SO_BAR), bar_mma = bar_ld + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SO_BAR + 16);
  const uint8_t* P = reinterpret_cast<const uint8_t*>(a.params);
  if (tid == 0) {
    mbar_init(bar_ld, 1);
    mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar_ld, 6 * TILE_BYTES + 512);
    bulk_g2s(smem_u32(smem + SO_U), reinterpret_cast<const uint8_t*>(a.u) + (size_t)tile * TILE_BYTES, TILE_BYTES, bar_ld);
    bulk_g2s(smem_u32(smem + SO_HM), reinterpret_cast<const uint8_t*>(a.hm) + (size_t)tile * 2 * TILE_BYTES, 2 * TILE_BYTES, bar_ld);
    bulk_g2s(smem_u32(smem + SO_WIH), P + LP_WIH + (size_t)slice * 2 * PIECE_BYTES, 2 * PIECE_BYTES, bar_ld);
    bulk_g2s(smem_u32(smem + SO_WHH), P + LP_WHH + (size_t)(slice * 2) * PIECE_BYTES, 2 * PIECE_BYTES, bar_ld);
    bulk_g2s(smem_u32(smem + SO_WHH + TILE_BYTES), P + LP_WHH + (size_t)(NPIECE + slice * 2) * PIECE_BYTES, 2 * PIECE_BYTES, bar_ld);
    bulk_g2s(smem_u32(smem + SO_BIAS), P + LP_BIAS + (size_t)slice * 512, 512, bar_ld);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  if (tid == 0) {
    mbar_wait(bar_ld, 0);
    fence_after_sync();
    const uint32_t idesc = instr_desc(128, false, false);
    mma_sequence(tmem, k_major(smem_u32(smem + SO_U), UK), k_major(smem_u32(smem + SO_WIH), UK), idesc, 96 / 16, false);
    mma_sequence(tmem, k_major(smem_u32(smem + SO_HM), UK), k_major(smem_u32(smem + SO_WHH), UK), idesc, UK / 16, true);
    mma_sequence(tmem, k_major(smem_u32(smem + SO_HM + TILE_BYTES), UK), k_major(smem_u32(smem + SO_WHH + TILE_BYTES), UK), idesc,
                 UK / 16, true);
    mma_commit(bar_mma);
  }
  mbar_wait(bar_ld, 0);    // the bias slice is read below by every thread
  mbar_wait(bar_mma, 0);
  fence_after_sync();
  // ---- cell epilogue: this thread = one sequence x 16 hidden units (piece = 2*slice + half) ----
  const int64_t s = (int64_t)tile * TILE + row;
  const int piece = 2 * slice + half, unit0 = 16 * piece;
  uint32_t gi[16], gf[16], gg[16], go[16];
  tmem_ld16(lane_base + half * 64, gi);
  tmem_ld16(lane_base + half * 64 + 16, gf);
  tmem_ld16(lane_base + half * 64 + 32, gg);
  tmem_ld16(lane_base + half * 64 + 48, go);
  if (s < a.n) {
    const float* sb = reinterpret_cast<const float*>(smem + SO_BIAS) + half * 64;
    const float m = a.not_done ? a.not_done[s] : 1.f;
    const float mn = a.not_done_next ? a.not_done_next[s] : 1.f;
    float cp[16], cn[16], hn[16], act[64];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 v = *reinterpret_cast<const float4*>(a.c_prev + s * HID + unit0 + 4 * q);
      cp[4 * q] = v.x, cp[4 * q + 1] = v.y, cp[4 * q + 2] = v.z, cp[4 * q + 3] = v.w;
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float i_ = sigmoid_(__uint_as_float(gi[k]) + sb[k]), f_ = sigmoid_(__uint_as_float(gf[k]) + sb[16 + k]);
      const float g_ = tanh_(__uint_as_float(gg[k]) + sb[32 + k]), o_ = sigmoid_(__uint_as_float(go[k]) + sb[48 + k]);
      cn[k] = fmaf(f_, cp[k] * m, i_ * g_);
      hn[k] = o_ * tanh_(cn[k]);
      act[k] = i_, act[16 + k] = f_, act[32 + k] = g_, act[48 + k] = o_;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
      *reinterpret_cast<float4*>(a.c + s * HID + unit0 + 4 * q) = make_float4(cn[4 * q], cn[4 * q + 1], cn[4 * q + 2], cn[4 * q + 3]);
    const size_t hoff = ((size_t)tile * 2 + unit0 / UK) * TILE_BYTES;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int off = tile_offset(row, unit0 % UK + 8 * q, UK);
      *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(a.hh) + hoff + off) =
          make_uint4(pack_bf16(hn[8 * q], hn[8 * q + 1]), pack_bf16(hn[8 * q + 2], hn[8 * q + 3]),
                     pack_bf16(hn[8 * q + 4], hn[8 * q + 5]), pack_bf16(hn[8 * q + 6], hn[8 * q + 7]));
      if (a.hm_next)
        *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(a.hm_next) + hoff + off) =
            make_uint4(pack_bf16(hn[8 * q] * mn, hn[8 * q + 1] * mn), pack_bf16(hn[8 * q + 2] * mn, hn[8 * q + 3] * mn),
                       pack_bf16(hn[8 * q + 4] * mn, hn[8 * q + 5] * mn), pack_bf16(hn[8 * q + 6] * mn, hn[8 * q + 7] * mn));
    }
    if (a.act) {
      uint8_t* at = reinterpret_cast<uint8_t*>(a.act) + ((size_t)tile * NPIECE + piece) * ACT_BYTES;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<uint4*>(at + tile_offset(row, 8 * q, PIECE_ROWS)) =
            make_uint4(pack_bf16(act[8 * q], act[8 * q + 1]), pack_bf16(act[8 * q + 2], act[8 * q + 3]),
                       pack_bf16(act[8 * q + 4], act[8 * q + 5]), pack_bf16(act[8 * q + 6], act[8 * q + 7]));
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

// h tile rows -> masked copy (rollout: the done flag of the env step arrives after the LSTM step has run)
__global__ void vine_lstm_mask_kernel(const uint8_t* __restrict__ hh, const float* __restrict__ not_done, int64_t n, uint8_t* __restrict__ hm) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // one 16-byte chunk (8 hidden units of one row)
  const int64_t chunks = ((n + TILE - 1) / TILE) * 2 * (TILE_BYTES / 16);
  if (i >= chunks) return;
  const int64_t tile = i / (2 * (TILE_BYTES / 16));
  const int within = (int)(i % (TILE_BYTES / 16));
  const int row = (within / 128) * 8 + (within % 8);   // chunk -> row of the row-blocked [128 x 128] tile
  const int64_t s = tile * TILE + row;
  const float m = (s < n) ? not_done[s] : 0.f;
  const uint4 v = reinterpret_cast<const uint4*>(hh)[i];
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  uint32_t o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 f = unpack_bf16(w[k]);
    o[k] = pack_bf16(f.x * m, f.y * m);
  }
  reinterpret_cast<uint4*>(hm)[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm + heads, one warp per row; lane l owns hidden units [8l, 8l+8)
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256) vine_lstm_head_kernel(const VineLstmHead a) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const uint8_t* P = reinterpret_cast<const uint8_t*>(a.params);
  const float* lng = reinterpret_cast<const float*>(P + LP_LNG) + 8 * lane;
  const float* lnb = reinterpret_cast<const float*>(P + LP_LNB) + 8 * lane;
  const float* wh = reinterpret_cast<const float*>(P + LP_WH) + 8 * lane;
  const float* bh = reinterpret_cast<const float*>(P + LP_BH);
  float g[8], b[8], w0[8], w1[8], w2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) g[k] = lng[k], b[k] = lnb[k], w0[k] = wh[k], w1[k] = wh[HID + k], w2[k] = wh[2 * HID + k];
  const bool sample = a.actions != nullptr;
  const uint32_t ctr = (sample && a.rng_counter) ? *a.rng_counter : 0u;
  const float ls0 = sample ? a.logstd[0] : 0.f, ls1 = sample ? a.logstd[1] : 0.f;
  const float vmean = a.value_stats[0], vstd = a.value_stats[1];
  for (int64_t s = warp0; s < a.n; s += nwarps) {
    const int64_t tile = s / TILE;
    const int row = (int)(s % TILE), unit = 8 * lane;
    const uint4 hv = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(a.hh) + ((size_t)tile * 2 + unit / UK) * TILE_BYTES +
                                                     tile_offset(row, unit % UK, UK));
    const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
    float h[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack_bf16(hw[k]);
      h[2 * k] = f.x, h[2 * k + 1] = f.y;
    }
    float sm = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) sm += h[k];
    const float mean = wsum(sm) * (1.f / HID);
    float sq = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) sq += (h[k] - mean) * (h[k] - mean);
    const float rstd = rsqrtf(wsum(sq) * (1.f / HID) + 1e-5f);     // torch.nn.LayerNorm: biased variance, eps 1e-5
    float d0 = 0.f, d1 = 0.f, d2 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float y = fmaf((h[k] - mean) * rstd, g[k], b[k]);
      d0 = fmaf(y, w0[k], d0), d1 = fmaf(y, w1[k], d1), d2 = fmaf(y, w2[k], d2);
    }
    const float mu0 = wsum(d0) + bh[0], mu1 = wsum(d1) + bh[1], v = wsum(d2) + bh[2];
    if (lane == 0) {
      const float value = fminf(fmaxf(v, -5.f), 5.f) * vstd + vmean;
      if (a.mu) *reinterpret_cast<float2*>(a.mu + 2 * s) = make_float2(mu0, mu1);
      if (a.value) a.value[s] = value;
      if (sample) {
        float z[4];
        normal4(philox4x32((uint32_t)a.seed, (uint32_t)(a.seed >> 32), (uint32_t)(a.global_env_offset + s), 16u, ctr, 0u), z);
        const float act0 = fmaf(__expf(ls0), z[0], mu0), act1 = fmaf(__expf(ls1), z[1], mu1);
        *reinterpret_cast<float2*>(a.actions + 2 * s) = make_float2(act0, act1);
        if (a.neglogp) a.neglogp[s] = 0.5f * (z[0] * z[0] + z[1] * z[1]) + 1.8378770664093453f + ls0 + ls1;
        if (a.env_actions)
          *reinterpret_cast<float2*>(a.env_actions + 2 * s) =
              make_float2(fminf(fmaxf(act0, -1.f), 1.f), fminf(fmaxf(act1, -1.f), 1.f));
      }
    }
  }
}

}  // namespace

extern "C" {

int vine_lstm_pack(const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, const float* ln_gamma,
                   const float* ln_beta, const float* w_mu, const float* b_mu, const float* w_v, const float* b_v, int num_obs,
                   void* packed, void* stream) {
  if (!w_ih || !w_hh || !b_ih || !b_hh || !ln_gamma || !ln_beta || !w_mu || !b_mu || !w_v || !b_v || !packed || num_obs < 1 ||
      num_obs >= K1)
    return VINE_ERR_INVALID_ARG;
  const int total = GATES * HID;
  vine_lstm_pack_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w_ih, w_hh, b_ih, b_hh, ln_gamma, ln_beta, w_mu, b_mu,
                                                                             w_v, b_v, num_obs, (uint8_t*)packed);
  return cudaGetLastError() == cudaSuccess ? VINE_OK : VINE_ERR_CUDA;
}

int vine_lstm_step(const VineLstmStep* a, void* stream) {
  if (!a || !a->params || !a->u || !a->hm || !a->c_prev || !a->c || !a->hh || a->n <= 0) return VINE_ERR_INVALID_ARG;
  if ((((uintptr_t)a->params) | ((uintptr_t)a->u) | ((uintptr_t)a->hm) | ((uintptr_t)a->hh) | ((uintptr_t)a->c_prev) |
       ((uintptr_t)a->c)) & 15u)
    return VINE_ERR_INVALID_ARG;
  static int configured = -1;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return VINE_ERR_CUDA;
  if (configured != dev) {
    if (cudaFuncSetAttribute(vine_lstm_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, STEP_SMEM) != cudaSuccess)
      return VINE_ERR_CUDA;
    configured = dev;
  }
  const dim3 grid((unsigned)((a->n + TILE - 1) / TILE), 8);
  vine_lstm_step_kernel<<<grid, THREADS, STEP_SMEM, (cudaStream_t)stream>>>(*a);
  return cudaGetLastError() == cudaSuccess ? VINE_OK : VINE_ERR_CUDA;
}

int vine_lstm_mask(const void* hh, const float* not_done, int64_t n, void* hm, void* stream) {
  if (!hh || !not_done || !hm || n <= 0) return VINE_ERR_INVALID_ARG;
  const int64_t chunks = ((n + TILE - 1) / TILE) * 2 * (TILE_BYTES / 16);
  vine_lstm_mask_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)hh, not_done, n, (uint8_t*)hm);
  return cudaGetLastError() == cudaSuccess ? VINE_OK : VINE_ERR_CUDA;
}

int vine_lstm_head(const VineLstmHead* a, void* stream) {
  if (!a || !a->params || !a->hh || !a->value_stats || a->n <= 0) return VINE_ERR_INVALID_ARG;
  if (a->actions && !a->logstd) return VINE_ERR_INVALID_ARG;
  int64_t blocks = (a->n + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  vine_lstm_head_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*a);
  return cudaGetLastError() == cudaSuccess ? VINE_OK : VINE_ERR_CUDA;
}

}  // extern "C"
