// vine_mlp.cu — fused actor-critic MLP forward on 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
//
// Network of cfg/train/Vine5LinkMovingBasePPO.yaml:10-30 (rl_games `actor_critic`, MLP variant):
//   x = clamp((obs - mean) / std, +-5)                       (rl_games RunningMeanStd, normalize_input)
//   h1 = ELU(W1 x + b1) [256], h2 = ELU(W2 h1 + b2) [128], h3 = ELU(W3 h2 + b3) [64]
//   mu = Wmu h3 + bmu [2],  v = Wv h3 + bv [1]
// One CTA owns a tile of 128 environments (UMMA M = 128) and runs the four layers back to back:
//   * all weights (bf16, pre-packed in the UMMA K-major core-matrix layout by vine_mlp_pack) arrive
//     with ONE TMA bulk copy (cp.async.bulk -> mbarrier complete_tx) and stay resident in shared
//     memory while the CTA loops over its tiles (persistent grid);
//   * each layer = K/16 `tcgen05.mma.cta_group::1.kind::f16` instructions issued by one thread,
//     A (activations) and B (weights) from shared memory, accumulator in TMEM (f32);
//   * epilogue: every thread owns one row, `tcgen05.ld.32x32b` pulls 32 accumulator columns at a
//     time, bias + ELU in registers, bf16 pack, 16-byte stores straight into the next layer's A tile.
// Activations never touch HBM; HBM traffic per env = O*4 bytes in, 12 bytes out.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vine_b200.h"

namespace {

constexpr int H1 = 256, H2 = 128, H3 = 64, NH = 16;  // NH: heads padded to the UMMA N granularity (mu0, mu1, v, 0...)
constexpr int K1 = 32;                               // obs width padded to 2 UMMA K-steps
constexpr int TILE_M = 128;
// byte offsets inside the packed parameter block (== shared-memory image)
constexpr int OFF_W1 = 0;
constexpr int OFF_W2 = OFF_W1 + H1 * K1 * 2;
constexpr int OFF_W3 = OFF_W2 + H2 * H1 * 2;
constexpr int OFF_W4 = OFF_W3 + H3 * H2 * 2;
constexpr int OFF_B = OFF_W4 + NH * H3 * 2;          // f32 biases: b1[256] b2[128] b3[64] b4[16]
constexpr int PACKED_BYTES = OFF_B + (H1 + H2 + H3 + NH) * 4;
constexpr int OFF_A = (PACKED_BYTES + 1023) / 1024 * 1024;   // activation tile, 128 x 256 bf16
constexpr int OFF_BAR = OFF_A + TILE_M * H1 * 2;     // 2 mbarriers + tmem address
constexpr int SMEM_BYTES = OFF_BAR + 64;
static_assert(PACKED_BYTES % 16 == 0, "bulk copy size must be a multiple of 16");
static_assert(PACKED_BYTES == VINE_MLP_PACKED_BYTES, "header constant out of date");

// K-major "interleaved" (no swizzle) UMMA operand layout: 8 x 16-byte core matrices stored contiguously
// (128 B); core matrices adjacent in K are LBO = 128 B apart, 8-row groups are SBO = (K/8)*128 B apart.
__host__ __device__ inline int core_offset(int row, int k, int K) {
  return (row >> 3) * ((K >> 3) * 128) + (k >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, int K) {
  // SmemDescriptor (cute/arch/mma_sm100_desc.hpp): addr>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48)
  const uint64_t lbo = 128 >> 4, sbo = ((K >> 3) * 128) >> 4;
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (lbo << 16) | (sbo << 32) | (1ull << 46);
}

__device__ __forceinline__ uint32_t instr_desc(int N) {
  // InstrDescriptor: D=f32 (1<<4), A=B=bf16 (1<<7, 1<<10), both K-major, N>>3 at [17,23), M>>4 at [24,29)
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float elu(float x) { return x > 0.f ? x : __expf(x) - 1.f; }

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// one layer's MMAs: D[128 x N] (TMEM) = A[128 x K] (smem) * W[N x K]^T (smem); issued by ONE thread
__device__ __forceinline__ void issue_layer(uint32_t tmem, uint32_t a_addr, uint32_t w_addr, int K, int N, uint32_t bar) {
  const uint32_t idesc = instr_desc(N);
  for (int ks = 0; ks < K / 16; ++ks) {
    // one UMMA consumes K = 16 = two 16-byte core-matrix columns = 2 * LBO = 256 bytes further along K
    mma_bf16(tmem, umma_desc(a_addr + ks * 256, K), umma_desc(w_addr + ks * 256, K), idesc, ks > 0);
  }
  // arrives on the mbarrier when every MMA above has completed (implies fence::before_thread_sync)
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// epilogue of a hidden layer: row `row` of D (N columns) -> bias + ELU -> bf16 -> next A tile (K' = N)
template <int N>
__device__ __forceinline__ void hidden_epilogue(uint32_t tmem_row, const float* bias, uint8_t* a_tile, int row) {
#pragma unroll 1
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(tmem_row + c0, r);
#pragma unroll
    for (int g = 0; g < 4; ++g) {  // 8 columns = one 16-byte core-matrix row
      uint32_t w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = c0 + g * 8 + 2 * i;
        w[i] = pack_bf16(elu(__uint_as_float(r[g * 8 + 2 * i]) + bias[c]), elu(__uint_as_float(r[g * 8 + 2 * i + 1]) + bias[c + 1]));
      }
      *reinterpret_cast<uint4*>(a_tile + core_offset(row, c0 + g * 8, N)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

__global__ void __launch_bounds__(TILE_M, 1)
vine_mlp_forward_kernel(const uint8_t* __restrict__ packed, const float* __restrict__ obs, const float* __restrict__ obs_mean,
                        const float* __restrict__ obs_inv_std, int64_t n, int num_obs, const float* __restrict__ value_stats,
                        float* __restrict__ mu_out, float* __restrict__ value_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bar_w = smem_u32(smem + OFF_BAR), bar_mma = bar_w + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 16);
  const float* biases = reinterpret_cast<const float*>(smem + OFF_B);

  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // TMA: one bulk copy brings every layer's weights and biases into shared memory
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_w), "r"(PACKED_BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem)), "l"(packed), "r"(PACKED_BYTES), "r"(bar_w) : "memory");
  }
  if (warp == 0) {  // TMEM: 256 f32 accumulator columns x 128 lanes
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_row = tmem + ((uint32_t)(warp * 32) << 16);  // this warp's 32 TMEM lanes
  mbar_wait(bar_w, 0);

  uint8_t* a_tile = smem + OFF_A;
  const uint32_t a_addr = smem_u32(a_tile), w_base = smem_u32(smem);
  uint32_t phase = 0;
  const int64_t ntiles = (n + TILE_M - 1) / TILE_M;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t e = tile * TILE_M + tid;
    // ---- A1: normalised observation row, bf16, zero-padded to K1 ----
    {
      float x[K1];
#pragma unroll
      for (int k = 0; k < K1; ++k) {
        float v = 0.f;
        if (e < n && k < num_obs) v = fminf(fmaxf((obs[e * num_obs + k] - obs_mean[k]) * obs_inv_std[k], -5.f), 5.f);
        x[k] = v;
      }
#pragma unroll
      for (int g = 0; g < K1 / 8; ++g)
        *reinterpret_cast<uint4*>(a_tile + core_offset(tid, g * 8, K1)) =
            make_uint4(pack_bf16(x[g * 8], x[g * 8 + 1]), pack_bf16(x[g * 8 + 2], x[g * 8 + 3]),
                       pack_bf16(x[g * 8 + 4], x[g * 8 + 5]), pack_bf16(x[g * 8 + 6], x[g * 8 + 7]));
    }
    // layers: {K, N, weight offset, bias offset}
    const int Ks[4] = {K1, H1, H2, H3}, Ns[4] = {H1, H2, H3, NH};
    const int Wo[4] = {OFF_W1, OFF_W2, OFF_W3, OFF_W4}, Bo[4] = {0, H1, H1 + H2, H1 + H2 + H3};
#pragma unroll
    for (int L = 0; L < 4; ++L) {
      // generic-proxy smem writes (the A tile) must be visible to the async proxy that feeds the tensor core
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        issue_layer(tmem, a_addr, w_base + Wo[L], Ks[L], Ns[L], bar_mma);
      }
      mbar_wait(bar_mma, phase);
      phase ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (L == 0) hidden_epilogue<H1>(tmem_row, biases + Bo[0], a_tile, tid);
      else if (L == 1) hidden_epilogue<H2>(tmem_row, biases + Bo[1], a_tile, tid);
      else if (L == 2) hidden_epilogue<H3>(tmem_row, biases + Bo[2], a_tile, tid);
      else {
        uint32_t r[32];
        tmem_ld32(tmem_row, r);   // columns 0..15 hold mu0, mu1, v (the rest of the 32 are stale and ignored)
        if (e < n) {
          const float* b = biases + Bo[3];
          mu_out[2 * e] = __uint_as_float(r[0]) + b[0];
          mu_out[2 * e + 1] = __uint_as_float(r[1]) + b[1];
          const float v = __uint_as_float(r[2]) + b[2];
          value_out[e] = fminf(fmaxf(v, -5.f), 5.f) * value_stats[1] + value_stats[0];   // RunningMeanStd(unnorm=True)
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

__global__ void vine_mlp_pack_kernel(const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
                                     const float* b3, const float* wmu, const float* bmu, const float* wv, const float* bv,
                                     int num_obs, uint8_t* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  auto put = [&](int off, int row, int k, int K, float v) {
    *reinterpret_cast<__nv_bfloat16*>(out + off + core_offset(row, k, K)) = __float2bfloat16_rn(v);
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

int vine_mlp_forward(const void* packed, const float* obs, const float* obs_mean, const float* obs_inv_std, int64_t n,
                     int num_obs, const float* value_stats, float* mu, float* value, void* stream) {
  if (!packed || !obs || !obs_mean || !obs_inv_std || !value_stats || !mu || !value || n <= 0 || num_obs < 1 || num_obs > K1)
    return VINE_ERR_INVALID_ARG;
  if (((uintptr_t)packed) & 15u) return VINE_ERR_INVALID_ARG;
  static int configured = -1;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return VINE_ERR_CUDA;
  if (configured != dev) {
    if (cudaFuncSetAttribute(vine_mlp_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess)
      return VINE_ERR_CUDA;
    configured = dev;
  }
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t ntiles = (n + TILE_M - 1) / TILE_M;
  const int grid = (int)(ntiles < sms ? ntiles : sms);
  vine_mlp_forward_kernel<<<grid, TILE_M, SMEM_BYTES, (cudaStream_t)stream>>>((const uint8_t*)packed, obs, obs_mean, obs_inv_std,
                                                                              n, num_obs, value_stats, mu, value);
  return cudaGetLastError() == cudaSuccess ? VINE_OK : VINE_ERR_CUDA;
}

}  // extern "C"
