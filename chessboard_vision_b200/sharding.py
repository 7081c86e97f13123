"""Multi-GPU layout of the hot path: frames / camera streams are independent units, so they are
sharded across ranks (one process per GPU) with NO data-path collective.  The only
cross-rank traffic is control plane: a barrier, the max-over-ranks of the timed region, and an
optional gather of the small per-square result records to rank 0.
"""
import numpy as np


def shard_range(n_items, rank, world):
    """Contiguous block of `n_items` owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(int(n_items), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def stream_owner(stream_id, world):
    """Camera stream -> rank that keeps its ChangeDetector / PieceDetector state (stream mode)."""
    return int(stream_id) % int(world)


def dist_max(value):
    """max over ranks of a python float (the time of the slowest rank); identity without a group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_records(local, dst=0):
    """Gather per-rank structured result arrays (n_local, n_sq) on `dst`, in rank order (= shard order
    of shard_range).  Returns the concatenated array on dst, None elsewhere."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    bucket = [None] * world if rank == dst else None
    dist.gather_object((local.dtype.descr, local.shape, local.tobytes()), bucket, dst=dst)
    if rank != dst:
        return None
    parts = [np.frombuffer(b, dtype=np.dtype([tuple(d) if isinstance(d, list) else d for d in descr])).reshape(shape)
             for descr, shape, b in bucket]
    return np.concatenate(parts, axis=0)


def bind_to_gpu_numa(device_index):
    """Pin this process to the CPUs NVML reports as local to the GPU, so that page-locked staging
    buffers are allocated on the GPU's own NUMA node (matters when 8 ranks stream frames at once).
    Returns the number of CPUs in the new affinity mask, or 0 when nothing was changed."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        allowed = os.sched_getaffinity(0)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1} & allowed
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def place_rank(local_rank, world, visible_gpus):
    """GPU index for `local_rank` of `world` ranks on a node that shows `visible_gpus` devices.

    Measured on this pool's 8 x B200 boxes (profiles/r02_h2d_probe_8gpu.txt): every GPU copies host -> device at
    55.6 GB/s alone, GPUs 4-7 keep that rate together (218 GB/s), GPUs 0-3 share one 115 GB/s host path (29 GB/s each).
    The host-buffer path is bound by exactly that copy, so runs on fewer GPUs than the node has start from the top:
    rank r -> GPU (visible - world + r).  With all GPUs in use, or a launcher that shows each job only its own GPUs,
    this is the identity."""
    visible_gpus, world, local_rank = int(visible_gpus), int(world), int(local_rank)
    if visible_gpus <= world:
        return local_rank % max(1, visible_gpus)
    return visible_gpus - world + local_rank


def proportional_split(total, rates, multiple=1):
    """Split `total` work items over ranks in proportion to their measured `rates` (items / s), each share a multiple of
    `multiple` except that the shares always sum to `total`.  Ranks whose host path is slower get fewer frames, so that
    all ranks finish together (the host-fed path on a box whose GPUs do not share the host bandwidth equally)."""
    rates = [max(float(r), 1e-9) for r in rates]
    s = sum(rates)
    shares = [int(total * r / s) // multiple * multiple for r in rates]
    rest = total - sum(shares)
    order = sorted(range(len(rates)), key=lambda i: -rates[i])
    i = 0
    while rest > 0:
        step = min(multiple, rest)
        shares[order[i % len(order)]] += step
        rest -= step
        i += 1
    return shares
