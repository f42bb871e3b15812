"""Data-parallel plumbing (one process per GPU, torch.distributed).  The hot path shards by cloud / triplet with
no data-path collective; only the training step exchanges data: ONE all-reduce of the flat
fp32 gradient buffer (107 619 floats for feature_dim 32), averaged over ranks."""
import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init(backend=None):
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def bind_to_gpu_numa_node(local_rank):
    """Restrict this process to the CPUs NVML reports as local to its GPU (nvmlDeviceSetCpuAffinity), so that pinned host buffers
    allocated afterwards are first-touched on the GPU's NUMA node and the H2D / D2H copies of 8 ranks do not all cross one socket.
    Returns the CPU count of the new affinity mask, or None when NVML / the call is unavailable (nothing changes then)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(int(local_rank)))
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


def shard_range(total, rank, world):
    """Contiguous [lo,hi) slice of `total` units owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_mean_(flat):
    """In-place mean over ranks of a flat gradient buffer: one collective per training step."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(dist.get_world_size())
    return flat


def allreduce_sum_(flat):
    """In-place SUM over ranks; pair it with get_train_op(grad_scale=1/world): the division then happens inside the Adam
    kernel (f3d_adam_step) instead of as an extra pass over the gradient buffer."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


def max_over_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def shutdown():
    if dist.is_initialized():
        dist.destroy_process_group()
