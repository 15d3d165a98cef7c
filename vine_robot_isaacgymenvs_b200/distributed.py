"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL on GPUs, gloo in CPU tests).

The env step needs no collective: envs shard by contiguous *global* env-id ranges and the Philox
streams are keyed by the global id, so any sharding reproduces the single-GPU result bit for bit.
Collectives exist only on the PPO side (the reference reaches them through rl_games' Horovod
wrapper: ``hvd.setup_algo / sync_stats / average_value``, learning/common_agent.py:125-138,219):
  * gradient all-reduce (+ the KL scalar packed into the same buffer),
  * running mean/std sufficient statistics (observations, values) and advantage sums.
"""
import os

import torch
import torch.distributed as dist


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_from_env(backend=None, device=None):
    """Initialise the default process group from torchrun's env vars (no-op when WORLD_SIZE == 1)."""
    rank, world, local_rank = rank_world()
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device(device or f"cuda:{local_rank}")
        dist.init_process_group(backend, **kw)
    return rank, world, local_rank


def env_shard(total_envs, rank, world):
    """Contiguous global env-id range [start, start+count) owned by ``rank`` (remainder to low ranks)."""
    base, rem = divmod(int(total_envs), int(world))
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def max_over_ranks(value, device="cpu"):
    """max of a python float over ranks (bench timing: the slowest rank defines the step time)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def allreduce_mean_(flat, extra_scalars=None):
    """In-place mean over ranks of a flat gradient buffer; ``extra_scalars`` (e.g. the KL estimate
    needed by the adaptive-LR schedule, Vine5LinkMovingBasePPO.yaml:64-66) ride in the same
    collective so one minibatch costs exactly one all-reduce."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return flat, extra_scalars
    world = dist.get_world_size()
    if extra_scalars is not None:
        packed = torch.cat([flat.reshape(-1), extra_scalars.reshape(-1).to(flat.dtype)])
        dist.all_reduce(packed)
        packed /= world
        flat.copy_(packed[: flat.numel()].view_as(flat))
        extra_scalars = packed[flat.numel():].clone()
    else:
        dist.all_reduce(flat)
        flat /= world
    return flat, extra_scalars


def merge_moments(count, mean, m2):
    """All-reduce running-mean/std sufficient statistics (Chan et al. parallel update):
    every rank passes its (count, mean[D], M2[D]); returns the statistics of the union."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return count, mean, m2
    c = torch.as_tensor(count, dtype=torch.float64, device=mean.device).reshape(1)
    s1 = mean.double() * c[0]
    packed = torch.cat([c, s1.reshape(-1)])
    dist.all_reduce(packed)
    tot = packed[0]
    gmean = packed[1:].view_as(mean) / tot
    # M2_total = sum_r [ M2_r + n_r (mean_r - gmean)^2 ]
    local = m2.double() + c[0] * (mean.double() - gmean) ** 2   # c[0]: keeps 0-dim statistics 0-dim
    dist.all_reduce(local)
    return tot.to(torch.float64), gmean.to(mean.dtype), local.to(m2.dtype)


class P2PChannel:
    """One channel of the gradient all-reduce over NVLink peer memory (include/vine_b200.h "Gradient all-reduce over peer
    memory", csrc/vine_p2p.cuh): this rank's region (two buffers of ``count`` f32 + flags) is allocated by the library and
    exported as a CUDA IPC handle; the handles are exchanged with ONE all_gather at construction; every peer region is mapped;
    ``ptr`` is the device-side channel the producer / consumer kernels take.  world == 1 works too (the channel then only
    exercises the buffers and the sequence logic).  NCCL is used for nothing but the handle exchange."""

    def __init__(self, lib, count, device):
        import ctypes as C
        self.lib, self.count, self.device = lib, int(count), torch.device(device)
        rank, world, _ = rank_world()
        if not (dist.is_available() and dist.is_initialized()):
            rank, world = 0, 1
        self.rank, self.world = rank, world
        region, handle = C.c_void_p(), (C.c_ubyte * 64)()
        with torch.cuda.device(self.device):
            rc = lib.vine_p2p_alloc(self.count, C.byref(region), handle)
        if rc != 0:
            raise RuntimeError(f"vine_p2p_alloc failed ({rc})")
        self._own = region
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=self.device)
        handles = [torch.empty_like(mine) for _ in range(world)]
        if world > 1:
            dist.all_gather(handles, mine)
        else:
            handles = [mine]
        self._opened = []
        regions = (C.c_void_p * world)()
        failed = None
        for r in range(world):
            if r == rank:
                regions[r] = region.value
                continue
            peer = C.c_void_p()
            raw = (C.c_ubyte * 64)(*handles[r].cpu().tolist())
            with torch.cuda.device(self.device):
                rc = lib.vine_p2p_open(raw, C.byref(peer))
            if rc != 0:
                failed = (r, rc)
                break
            self._opened.append(peer)
            regions[r] = peer.value
        if world > 1:   # every rank learns whether ALL mappings succeeded, so that all raise (or none): nobody is left at a barrier
            ok = torch.tensor([0 if failed else 1], device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok) == 0:
                for peer in self._opened:
                    lib.vine_p2p_close(peer)
                lib.vine_p2p_free(region)
                raise RuntimeError(("vine_p2p_open failed for rank %d (%d)" % failed if failed else "vine_p2p_open failed on another rank") +
                                   ": CUDA IPC / NVLink peer access is required (one process per GPU on ONE node); pass "
                                   "grad_allreduce='nccl' to use NCCL instead")
        ch = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = lib.vine_p2p_channel_create(regions, world, rank, self.count, C.byref(ch))
        if rc != 0:
            raise RuntimeError(f"vine_p2p_channel_create failed ({rc})")
        self.ptr = ch
        if world > 1:
            dist.barrier()          # nobody launches before every rank has mapped every region

    def status(self):
        """(exchanges completed, timed_out) -- synchronises the device."""
        import ctypes as C
        seq, err = C.c_uint32(), C.c_uint32()
        assert self.lib.vine_p2p_channel_status(self.ptr, C.byref(seq), C.byref(err)) == 0
        return int(seq.value), bool(err.value)

    def timing(self):
        """Mean microseconds per exchange since the last call, seen by block 0 of the consumer kernel: time to publish this
        rank's flag, and per rank the time from the kernel's entry until that rank's flag was seen.  Synchronises."""
        import ctypes as C
        out = (C.c_double * 22)()
        assert self.lib.vine_p2p_channel_timing(self.ptr, out, 22) == 0
        return {"signal_us": out[0] / 1e3, "wait_us": [out[1 + r] / 1e3 for r in range(self.world)],
                "sums_ready_us": out[18] / 1e3, "consumer_kernel_us": out[17] / 1e3,
                "between_consumer_end_and_producer_entry_us": out[19] / 1e3, "producer_kernel_us": out[20] / 1e3,
                "producer_end_to_consumer_entry_us": out[21] / 1e3}

    def close(self):
        if getattr(self, "ptr", None) is not None:
            torch.cuda.synchronize(self.device)
            if self.world > 1:
                dist.barrier()      # peers may still be reading this rank's region
            self.lib.vine_p2p_channel_destroy(self.ptr)
            for peer in self._opened:
                self.lib.vine_p2p_close(peer)
            self.lib.vine_p2p_free(self._own)
            self.ptr = None
