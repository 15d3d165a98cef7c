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
  unsigned int error;              // set when a wait timed out (a peer died): the caller checks it, nothing hangs
  unsigned int pad;
};

__device__ __forceinline__ float* p2p_buffer(const VineP2PChannel* ch, int rank, unsigned seq) {
  return reinterpret_cast<float*>(ch->peer_base[rank] + VINE_P2P_FLAG_BYTES) + (size_t)(seq & 1u) * (size_t)ch->count;
}

// producer side: where this rank's contribution of the current exchange goes
__device__ __forceinline__ float* p2p_local_buffer(const VineP2PChannel* ch) { return p2p_buffer(ch, ch->rank, ch->seq); }

// consumer side, every block, before the first p2p_sum; returns the sequence number of this exchange
__device__ __forceinline__ unsigned p2p_exchange_begin(VineP2PChannel* ch) {
  const unsigned seq = ch->seq;
  const int W = ch->world;
  if ((int)threadIdx.x < W) {
    if (blockIdx.x == 0) {         // the producer kernel has completed (stream order): publish "my buffer seq is ready"
      __threadfence_system();
      unsigned* flag = reinterpret_cast<unsigned*>(ch->peer_base[threadIdx.x]) + ch->rank;
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(seq + 1u) : "memory");
    }
    const unsigned* mine = reinterpret_cast<const unsigned*>(ch->peer_base[ch->rank]) + threadIdx.x;
    const long long t0 = clock64();
    unsigned v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if ((int)(v - (seq + 1u)) >= 0) break;
      if (clock64() - t0 > 8000000000ll) { ch->error = 1u; break; }   // ~4 s: a peer is gone; do not hang the GPU
    } while (true);
  }
  __syncthreads();
  return seq;
}

// sum over the ranks (rank order) of element idx of the buffers of exchange seq
__device__ __forceinline__ float p2p_sum(const VineP2PChannel* ch, unsigned seq, int idx) {
  float acc = 0.f;
  const int W = ch->world;
#pragma unroll 1
  for (int r = 0; r < W; ++r) acc += __ldcv(p2p_buffer(ch, r, seq) + idx);
  return acc;
}

// consumer side, every block, after its last p2p_sum
__device__ __forceinline__ void p2p_exchange_end(VineP2PChannel* ch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(&ch->ticket, 1u) + 1u;
    if (done == gridDim.x) { ch->ticket = 0u; __threadfence(); ch->seq = ch->seq + 1u; }
  }
}
