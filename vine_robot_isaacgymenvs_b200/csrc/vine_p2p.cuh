// vine_p2p.cuh — one-shot all-reduce of the PPO gradient vector over NVLink peer memory, folded into the kernels on either
// side of it (replaces the per-minibatch NCCL all-reduce; reference call sites: learning/common_agent.py:125-126,219 /
// rl_games a2c_common `dist.all_reduce` of the flattened gradients).
//
// Every rank owns one REGION in its own HBM (cudaMalloc, exported with cudaIpcGetMemHandle, mapped by every peer):
//     [flags: one u32 per peer rank][pad to 256 B][buffer 0: count f32][buffer 1: count f32]
// Per minibatch s (all ranks run the same launches):
//   producer  (vine_ppo_reduce / vine_lstm_reduce): writes the rank's gradient sum into ITS OWN buffer s & 1;
//   consumer  (vine_ppo_adam / vine_lstm_adam), at its start: block 0 stores s + 1 into flag[rank] of EVERY rank
//             (st.release.sys after a system fence), every block waits until all of its own flags reached s + 1
//             (ld.acquire.sys), then each thread adds the W ranks' buffers in rank order for its parameter --
//             identical order on every rank, hence bit-identical parameters -- and goes on with Adam.
// The two buffers make the exchange safe without a second barrier: a rank can overwrite buffer s & 1 (at minibatch s + 2) only
// after it passed the wait of s + 1, which needs every peer's flag s + 2, which a peer stores only at the start of its
// consumer of s + 1, i.e. after its consumer of s has finished reading.  The sequence number lives on the device (the launches
// are captured in CUDA graphs); the last block of the consumer advances it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define VINE_P2P_MAX_RANKS 16
#define VINE_P2P_FLAG_BYTES 256

struct VineP2PChannel {            // lives in device memory; built by vine_p2p_channel_create
  unsigned long long peer_base[VINE_P2P_MAX_RANKS];   // region of every rank as mapped in THIS process (own region included)
  int world, rank;
  long long count;                 // f32 per buffer
  unsigned int seq;                // exchanges completed
  unsigned int ticket;             // consumer blocks finished in the current exchange
  unsigned int producer_ticket;    // producer blocks finished in the current exchange
  unsigned int error;              // set when a wait timed out (a peer died): the caller checks it, nothing hangs
  // timing of the exchanges as seen by block 0 (globaltimer ns, accumulated; read by vine_p2p_channel_status): from the
  // consumer's entry to "my flag is stored at every peer's" and to "peer r's flag arrived here"
  unsigned long long t_signal, t_wait[VINE_P2P_MAX_RANKS], t_n;
  unsigned long long t_prod_begin, t_prev_end, t_mb, t_prod, t_gap;   // producer entry, previous consumer's end; sums of: consumer end -> producer entry
                                                                     // (the kernels in between), producer duration, producer end -> consumer entry
  unsigned long long t_begin, t_kernel, t_loads;   // entry of block 0; sum of (last block's end - entry); block 0: entry -> sums ready
};

__device__ __forceinline__ float* p2p_buffer(const VineP2PChannel* ch, int rank, unsigned seq) {
  return reinterpret_cast<float*>(ch->peer_base[rank] + VINE_P2P_FLAG_BYTES) + (size_t)(seq & 1u) * (size_t)ch->count;
}

// producer side: where this rank's contribution of the current exchange goes
__device__ __forceinline__ float* p2p_local_buffer(const VineP2PChannel* ch) { return p2p_buffer(ch, ch->rank, ch->seq); }

// producer side, every block, after its last store into p2p_local_buffer: the LAST block to get here publishes "this rank's
// buffer of the current exchange is ready" to every rank, so the flags travel while the producer retires and the consumer
// launches (a system fence + a remote store are ~3.5 us when the consumer has to do them first)
__device__ __forceinline__ void p2p_producer_begin(VineP2PChannel* ch) {   // timing only
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long g;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    ch->t_prod_begin = g;
    if (ch->t_prev_end) ch->t_mb += g - ch->t_prev_end;
  }
}

__device__ __forceinline__ void p2p_producer_done(VineP2PChannel* ch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned done = atomicAdd(&ch->producer_ticket, 1u) + 1u;
    if (done == gridDim.x) {
      ch->producer_ticket = 0u;
      __threadfence_system();      // ONE system fence orders every block's buffer stores before the flags ...
      const unsigned seq = ch->seq;
      for (int r = 0; r < ch->world; ++r) {   // ... which then go out back to back (a release store per peer would serialise
        unsigned* flag = reinterpret_cast<unsigned*>(ch->peer_base[r]) + ch->rank;   // the NVLink round trips: 22 us at 8 ranks)
        asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(seq + 1u) : "memory");
      }
      unsigned long long g;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
      ch->t_prod += g - *(volatile unsigned long long*)&ch->t_prod_begin;
      ch->t_prev_end = g;   // reused below as "producer end" until the consumer overwrites it
    }
  }
}

// consumer side, every block, before the first p2p_sum; returns the sequence number of this exchange.
// signal = false when the producer kernel has already published the flag (p2p_producer_done)
__device__ __forceinline__ unsigned p2p_exchange_begin(VineP2PChannel* ch, bool signal = true) {
  const unsigned seq = ch->seq;
  const int W = ch->world;
  if ((int)threadIdx.x < W) {
    unsigned long long g0 = 0;
    if (blockIdx.x == 0) {         // the producer kernel has completed (stream order): publish "my buffer seq is ready"
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
      if (threadIdx.x == 0) { ch->t_begin = g0; if (!signal && ch->t_prev_end) ch->t_gap += g0 - ch->t_prev_end; }
      if (signal) {
        __threadfence_system();
        unsigned* flag = reinterpret_cast<unsigned*>(ch->peer_base[threadIdx.x]) + ch->rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(seq + 1u) : "memory");
      }
      if (threadIdx.x == 0) {
        unsigned long long g1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
        ch->t_signal += g1 - g0;
        ch->t_n += 1;
      }
    }
    const unsigned* mine = reinterpret_cast<const unsigned*>(ch->peer_base[ch->rank]) + threadIdx.x;
    const long long t0 = clock64();
    unsigned v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if ((int)(v - (seq + 1u)) >= 0) break;
      if (clock64() - t0 > 8000000000ll) { ch->error = 1u; break; }   // ~4 s: a peer is gone; do not hang the GPU
    } while (true);
    if (blockIdx.x == 0) {
      unsigned long long g2;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g2));
      ch->t_wait[threadIdx.x] += g2 - g0;
    }
  }
  __syncthreads();
  return seq;
}

// sum over the ranks (rank order) of element idx of the buffers of exchange seq
__device__ __forceinline__ float p2p_sum(const VineP2PChannel* ch, unsigned seq, int idx) {
  const int W = ch->world;
  const size_t off = (size_t)(seq & 1u) * (size_t)ch->count + (size_t)idx;
  float v[VINE_P2P_MAX_RANKS];
#pragma unroll
  for (int r = 0; r < VINE_P2P_MAX_RANKS; ++r)   // all peer loads in flight at once (an NVLink round trip each), then added in rank order
    v[r] = r < W ? __ldcv(reinterpret_cast<const float*>(ch->peer_base[r] + VINE_P2P_FLAG_BYTES) + off) : 0.f;
  float acc = 0.f;
#pragma unroll
  for (int r = 0; r < VINE_P2P_MAX_RANKS; ++r) acc += v[r];   // + 0.f beyond the world size: exact
  return acc;
}

// four consecutive elements at once (the loss statistics behind the gradient): all 4 W loads are issued before the first add --
// a warp issues in order, so four calls of p2p_sum would be four NVLink round trips one after the other
__device__ __forceinline__ void p2p_sum4(const VineP2PChannel* ch, unsigned seq, int idx, float out[4]) {
  const int W = ch->world;
  const size_t off = (size_t)(seq & 1u) * (size_t)ch->count + (size_t)idx;
  float v[VINE_P2P_MAX_RANKS][4];
#pragma unroll
  for (int r = 0; r < VINE_P2P_MAX_RANKS; ++r)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      v[r][j] = r < W ? __ldcv(reinterpret_cast<const float*>(ch->peer_base[r] + VINE_P2P_FLAG_BYTES) + off + j) : 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < VINE_P2P_MAX_RANKS; ++r) acc += v[r][j];
    out[j] = acc;
  }
}

// Block-cooperative form for consumers whose block owns the contiguous elements [base, base + n) (base % 4 == 0,
// n <= 4 * blockDim.x): 128-bit peer loads (a 512-byte request per warp and peer instead of four 128-byte ones: NVLink
// reads are request-bound), rank-order sums left in shared memory.  The buffers are padded to a multiple of 4 floats.
__device__ __forceinline__ void p2p_sum_block(const VineP2PChannel* ch, unsigned seq, int base, int n, float* smem_out) {
  const int W = ch->world;
  const size_t off4 = ((size_t)(seq & 1u) * (size_t)ch->count + (size_t)base) >> 2;
  if ((int)threadIdx.x < ((n + 3) >> 2)) {
    float4 v[VINE_P2P_MAX_RANKS];
#pragma unroll
    for (int r = 0; r < VINE_P2P_MAX_RANKS; ++r)
      v[r] = r < W ? __ldcv(reinterpret_cast<const float4*>(ch->peer_base[r] + VINE_P2P_FLAG_BYTES) + off4 + threadIdx.x)
                   : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < VINE_P2P_MAX_RANKS; ++r) { acc.x += v[r].x; acc.y += v[r].y; acc.z += v[r].z; acc.w += v[r].w; }
    reinterpret_cast<float4*>(smem_out)[threadIdx.x] = acc;
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long g;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    const_cast<VineP2PChannel*>(ch)->t_loads += g - ch->t_begin;
  }
}

// consumer side, every block, after its last p2p_sum
__device__ __forceinline__ void p2p_exchange_end(VineP2PChannel* ch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(&ch->ticket, 1u) + 1u;
    if (done == gridDim.x) {
      unsigned long long g;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
      ch->t_kernel += g - *(volatile unsigned long long*)&ch->t_begin;
      ch->t_prev_end = g;
      ch->ticket = 0u; __threadfence(); ch->seq = ch->seq + 1u;
    }
  }
}
