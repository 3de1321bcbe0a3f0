"""Row sharding for multi-GPU runs: one process per GPU, cells (rows) partitioned into
contiguous blocks balanced by stored entries; every n_eff-sized object is replicated and all-reduced
inside the library (NCCL).  `torch.distributed` is only the plumbing that hands the NCCL unique id
from rank 0 to the other ranks."""
from __future__ import annotations

import os

import numpy as np


def partition_rows_by_nnz(row_offsets, nparts):
    """Contiguous row ranges [(r0, r1), ...] whose stored-entry counts are as equal as a row-aligned
    cut allows: boundary i is the first row whose offset reaches i * nnz / nparts."""
    off = np.asarray(row_offsets, dtype=np.int64)
    nrows = len(off) - 1
    nnz = int(off[-1])
    if nnz == 0:
        cuts = [(nrows * i) // nparts for i in range(nparts + 1)]
    else:
        targets = (np.arange(1, nparts, dtype=np.float64) * nnz / nparts)
        inner = np.searchsorted(off, targets, side="left")
        cuts = [0] + [int(min(max(c, 0), nrows)) for c in inner] + [nrows]
        for i in range(1, len(cuts)):
            cuts[i] = max(cuts[i], cuts[i - 1])
    return [(cuts[i], cuts[i + 1]) for i in range(nparts)]


def partition_rows_even(nrows, nparts):
    """Equal row counts (used for the device-generated synthetic shards, whose rows are i.i.d.)."""
    return [((nrows * i) // nparts, (nrows * (i + 1)) // nparts) for i in range(nparts)]


def shard_csr(row_offsets, col_indices, values, r0, r1):
    """Host CSR arrays of rows [r0, r1) with offsets rebased to 0."""
    off = np.asarray(row_offsets)
    s, e = int(off[r0]), int(off[r1])
    new_off = (off[r0:r1 + 1].astype(np.int64) - s).astype(off.dtype)
    return new_off, col_indices[s:e], values[s:e]


def broadcast_unique_id(make_id, rank, world):
    """Rank 0 calls `make_id()` (-> 128 bytes); everyone returns the same bytes."""
    import torch
    import torch.distributed as dist
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = torch.frombuffer(bytearray(make_id()), dtype=torch.uint8).clone()
    if world > 1:
        backend = dist.get_backend()
        if "gloo" in str(backend):
            dist.broadcast(buf, src=0)
        else:
            dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
            t = buf.to(dev)
            dist.broadcast(t, src=0)
            buf = t.cpu()
    return bytes(buf.numpy().tobytes())


def bind_to_gpu_numa_node(local_rank):
    """Multi-rank runs: pin this process (and with it the first-touch placement of the pinned host buffers it allocates
    afterwards) to the CPUs NVML reports as closest to its GPU, so that the 8 concurrent host->device uploads of a
    row-sharded fit do not cross the socket interconnect.  Best effort: returns the CPU count bound to, or 0."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local_rank)
        bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return len(os.sched_getaffinity(0))
    except Exception:
        return 0


def init_context_from_env():
    """Context for this process under torchrun (RANK / LOCAL_RANK / WORLD_SIZE); torch.distributed must
    already be initialised when WORLD_SIZE > 1."""
    from .api import Context
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1:
        return Context(local)
    uid = broadcast_unique_id(Context.nccl_unique_id, rank, world)
    return Context(local, rank, world, uid)
