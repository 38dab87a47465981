"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed only carries the 128-byte
NCCL unique id and the timing reductions; the data path (panel broadcasts, slice all-gathers, the
P+1-scalar all-reduce) runs inside libgpb200 on its own NCCL communicator over NVLink."""


def broadcast_bytes(payload_or_none, src=0):
    """rank `src` passes bytes, the others None; everyone gets the bytes (any backend, incl. gloo)."""
    import torch.distributed as dist
    obj = [payload_or_none if dist.get_rank() == src else None]
    dist.broadcast_object_list(obj, src=src)
    return obj[0]


def init_engine_comm(engine, p2p=True):
    """Join `engine` (one per rank) into one NCCL communicator spanning the default process group."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 1, 0
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = broadcast_bytes(engine.nccl_unique_id() if rank == 0 else None, src=0)
    engine.comm_init(world, rank, uid)
    if p2p and hasattr(engine, "ipc_export"):
        exchange_ipc(engine)
    return world, rank


def exchange_ipc(engine):
    """Fused panel broadcast: map every peer's factor buffers over NVLink (CUDA IPC).  Collective; must be repeated
    after a set_data that reallocates the device buffers (the old mappings are closed by the engine)."""
    import torch.distributed as dist
    world = dist.get_world_size()
    blobs = [None] * world
    dist.all_gather_object(blobs, engine.ipc_export())
    engine.ipc_import(blobs)


def max_over_ranks(value):
    """max of a python float over the default process group (device-side for nccl, host for gloo)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def partition_bounds(m, world):
    """contiguous, balanced split of m items over `world` ranks: rank r owns [b[r], b[r+1])"""
    return [(m * r) // world for r in range(world + 1)]


def allgather_concat(local, counts):
    """Every rank passes its float64 vector (length counts[rank]); every rank gets the concatenation in rank order.
    Device-side for nccl, host for gloo (only O(M) result values travel -- the data path stays inside the engine)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    width = max(max(counts), 1)
    buf = torch.zeros(width, dtype=torch.float64, device=dev)
    if local.size:
        buf[:local.size] = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64)).to(dev)
    outs = [torch.empty_like(buf) for _ in counts]
    dist.all_gather(outs, buf)
    return np.concatenate([o[:c].cpu().numpy() for o, c in zip(outs, counts)])
