"""One process per GPU, batch-sharded (images are independent units of this path).

The reference is single-process (no torch.distributed anywhere); this is the harness the multi-GPU configurations of
BASELINE.json need around its loops:
  * inference / evaluation: each rank runs the model on its contiguous batch shard; the only collective is the gather
    of logits (B x nc) and embeddings (B x g*g) for the classification / t-SNE modes -- all_gather_into_tensor;
  * training: DistributedDataParallel over NCCL (NVLink 5 / NVSwitch) averages the 12.7 M fp32 gradients (51 MB per
    step at truncate_layer 7); the head's gradients are produced first in backward, the encoder's last, so DDP's
    buckets overlap the all-reduce with the cuDNN backward of the encoder.
BatchNorm note: the reference trains with plain BatchNorm in train mode; under data parallelism each rank normalises
with its own shard's statistics (standard DDP behaviour) unless `sync_bn=True` converts the encoder to SyncBatchNorm.
"""
from __future__ import annotations

import os
from typing import Optional, Sequence, Tuple

import torch
import torch.distributed as dist


ORIGINAL_AFFINITY = None          # the process's CPU set before bind_to_gpu_cpus() narrowed it (restore_cpu_affinity())


def restore_cpu_affinity() -> None:
    """Gives the process back every CPU it had at start (host-side work that should use all cores, e.g. a CPU baseline)."""
    if ORIGINAL_AFFINITY is not None:
        os.sched_setaffinity(0, ORIGINAL_AFFINITY)


def bind_to_gpu_cpus(device: torch.device) -> Optional[Sequence[int]]:
    """Pins this process to the CPU cores NVML reports as local to `device` (same NUMA node / PCIe root), so the pinned
    staging buffers it allocates afterwards and its launch thread sit next to the GPU. With several ranks on a
    two-socket box the scheduler otherwise places ranks on the remote socket and host->device copies cross the
    inter-socket link. A placement hint only: returns None (and changes nothing) when NVML cannot answer."""
    try:
        import pynvml
        pynvml.nvmlInit()
        p = torch.cuda.get_device_properties(device)
        bus_id = f"{p.pci_domain_id:08x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus_id.encode())
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        global ORIGINAL_AFFINITY
        if ORIGINAL_AFFINITY is None:
            ORIGINAL_AFFINITY = os.sched_getaffinity(0)
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int, torch.device]:
    """Reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun). Single-process when WORLD_SIZE is absent or 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    use_cuda = torch.cuda.is_available()
    device = torch.device(f"cuda:{local}") if use_cuda else torch.device("cpu")
    if use_cuda:
        torch.cuda.set_device(device)
        # "auto": only with several ranks per box, where the scheduler may otherwise place a rank on the socket far from
        # its GPU; a single rank keeps every core (its launch thread, the pinned-buffer copies of HostCollector and the
        # sampler processes then never compete for a narrowed CPU set).
        want = os.environ.get("GRAMHEAD_CPU_AFFINITY", "auto")
        if want == "1" or (want == "auto" and world > 1):
            bind_to_gpu_cpus(device)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kwargs = {}
        if use_cuda:
            kwargs["device_id"] = device
        dist.init_process_group(backend or ("nccl" if use_cuda else "gloo"), rank=rank, world_size=world, **kwargs)
    return rank, world, local, device


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n items for `rank`; the first n % world ranks get one extra item."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi]


def gather_rows(t: torch.Tensor, total_rows: int, world: int) -> torch.Tensor:
    """Concatenates every rank's rows (dim 0) in rank order. Shards may differ by one row (shard_bounds); they are
    padded to the largest shard for all_gather_into_tensor and trimmed afterwards."""
    if world == 1:
        return t
    rank = dist.get_rank()
    longest = shard_bounds(total_rows, 0, world)[1]
    pad = longest - t.shape[0]
    src = t.contiguous()
    if pad:
        src = torch.cat([src, src.new_zeros((pad,) + tuple(t.shape[1:]))], dim=0)
    out = src.new_empty((world * longest,) + tuple(t.shape[1:]))
    dist.all_gather_into_tensor(out, src)
    if total_rows == world * longest:
        return out
    pieces = []
    for r in range(world):
        lo, hi = shard_bounds(total_rows, r, world)
        pieces.append(out[r * longest: r * longest + (hi - lo)])
    del rank
    return torch.cat(pieces, dim=0)


def wrap_ddp(model: torch.nn.Module, device: torch.device, sync_bn: bool = False, bucket_cap_mb: int = 25,
             broadcast_buffers: bool = True, static_graph: bool = False, grad_compression: str = "none",
             for_graph_capture: bool = False) -> torch.nn.Module:
    """DistributedDataParallel around the drop-in model (construct it with device=f'cuda:{local_rank}').

    bucket_cap_mb: gradient bucket size. The head's 16.8 MB of gradients are ready first in backward (attention and
    classifier sit at the end of the forward), the encoder's 34 MB follow layer by layer while cuDNN is still busy, so
    smaller buckets let the all-reduce start earlier and finish under the encoder's backward.
    static_graph: the set of used parameters does not change between iterations (true for this model): DDP skips its
    per-iteration bookkeeping and may reorder buckets after the first step.
    grad_compression: "bf16" all-reduces bf16 copies of the gradient buckets (half the NVLink bytes; the averaged gradient
    is rounded to bf16 once) -- opt-in, it changes the numerics of the update; "none" keeps fp32 (default).
    for_graph_capture: construct DDP on a side stream, as capturing its step in a CUDA graph requires."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return model
    if sync_bn:
        model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
    ids = [device.index] if device.type == "cuda" else None

    def build():
        return torch.nn.parallel.DistributedDataParallel(model, device_ids=ids, bucket_cap_mb=bucket_cap_mb,
                                                         broadcast_buffers=broadcast_buffers, gradient_as_bucket_view=True,
                                                         static_graph=static_graph)
    if for_graph_capture and device.type == "cuda":
        # torch's rule for capturing a DDP step in a CUDA graph (functions.GraphedTrainStep): DDP is constructed on a side
        # stream, and TORCH_NCCL_ASYNC_ERROR_HANDLING=0 must have been set before init_process_group
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            ddp = build()
        torch.cuda.current_stream(device).wait_stream(side)
    else:
        ddp = build()
    if grad_compression == "bf16":
        from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
        ddp.register_comm_hook(None, default_hooks.bf16_compress_hook)
    elif grad_compression != "none":
        raise ValueError(f"grad_compression must be 'none' or 'bf16', got {grad_compression!r}")
    return ddp


def max_over_ranks(value: float, device: torch.device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(device: torch.device) -> None:
    if dist.is_initialized() and dist.get_world_size() > 1:
        if device.type == "cuda":
            dist.barrier(device_ids=[device.index])
        else:
            dist.barrier()
