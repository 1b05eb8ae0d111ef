"""Multi-GPU sharding of the forward run: one process per GPU, basins / ensemble
members / grid cells are independent (`/root/reference/src/bucket.jl:229-233` broadcasts over nodes
with no cross-node term), so every rank runs the stepper on a contiguous slice of the node axis
and there is NO data-path collective.  The only exchange is an optional gather of small per-node
results (objective values, final states) over `torch.distributed` (NCCL on GPUs, gloo in the CPU
tests)."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np


def shard_range(n_nodes: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced `[n0, n1)`: the first `n_nodes % world` ranks get one extra node."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(int(n_nodes), int(world))
    n0 = rank * base + min(rank, rem)
    return n0, n0 + base + (1 if rank < rem else 0)


def shard_params(params: Dict, n0: int, n1: int, n_nodes: int) -> Dict:
    """Slice per-node parameter vectors (length `n_nodes`) to `[n0, n1)`; scalars, shared
    (shorter) vectors and neural-network parameters are replicated."""
    out = {"params": {}}
    for k, v in params["params"].items():
        a = np.asarray(v)
        out["params"][k] = a[n0:n1] if (a.ndim == 1 and a.size == n_nodes) else v
    if "nns" in params:
        out["nns"] = params["nns"]
    return out


def shard_htypes(htypes: np.ndarray, n0: int, n1: int) -> np.ndarray:
    """Local HRU-type vector: per-node types (`1:N`) are renumbered to `1:(n1-n0)`; shared types are kept."""
    h = np.asarray(htypes)
    if np.array_equal(h, np.arange(1, len(h) + 1)):
        return np.arange(1, n1 - n0 + 1)
    return h[n0:n1]


def shard_initstates(initstates, n0: int, n1: int, n_nodes: int, n_states: int):
    if initstates is None:
        return None
    if isinstance(initstates, dict):
        return {k: (np.asarray(v)[n0:n1] if np.ndim(v) == 1 else v) for k, v in initstates.items()}
    v = np.asarray(initstates).reshape(n_states, n_nodes)       # state-major, node-fastest
    return np.ascontiguousarray(v[:, n0:n1]).reshape(-1)


def gather_nodes(local: np.ndarray, n_nodes: int, group=None, device: Optional[str] = None) -> np.ndarray:
    """All-gather a per-node vector (`[n1-n0]` or `[k, n1-n0]`, node axis last) from every rank into the
    full `[..., n_nodes]` array, in rank order.  Uses the default process group's backend: NCCL
    (tensors staged on `device`) or gloo (CPU)."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return np.asarray(local)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    loc = np.ascontiguousarray(local, dtype=np.float64)
    lead = loc.shape[:-1]
    sizes = [shard_range(n_nodes, r, world) for r in range(world)]
    width = max(b - a for a, b in sizes)
    dev = device or ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    pad = np.zeros(lead + (width,), dtype=np.float64)
    pad[..., :loc.shape[-1]] = loc
    t = torch.from_numpy(pad).to(dev)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    full = np.empty(lead + (n_nodes,), dtype=np.float64)
    for r, (a, b) in enumerate(sizes):
        full[..., a:b] = outs[r].cpu().numpy()[..., :b - a]
    return full
