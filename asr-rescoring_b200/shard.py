"""Multi-GPU plumbing: utterances are sharded across ranks (one process per GPU); the only
exchanges are a gather of per-hypothesis scores and a sum of CER counts.
There is no data-path collective inside the scoring kernels — every hypothesis is
independent (MLM_PLL/main.py:106-107 only accumulates per (utt, hyp))."""
from __future__ import annotations

import os
from typing import List, Sequence

import numpy as np


def dist_env():
    """(rank, world_size, local_rank) from the torchrun environment; (0, 1, 0) otherwise."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def utterance_costs(hyp_lengths: Sequence[Sequence[int]]) -> np.ndarray:
    """c(u) = sum_k L_k * (L_k + 2): expanded tokens, proportional to encoder FLOPs."""
    return np.array([sum(int(L) * (int(L) + 2) for L in ls) for ls in hyp_lengths], np.int64)


def lpt_partition(costs: Sequence[int], world_size: int) -> List[np.ndarray]:
    """Greedy longest-processing-time partition, deterministic in the input order: items by
    descending cost (stable), each to the currently lightest bin (lowest rank on ties).
    Returns, per rank, the ascending original indices it owns."""
    costs = np.asarray(costs, np.int64)
    order = np.argsort(-costs, kind="stable")
    load = np.zeros(world_size, np.int64)
    bins: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = int(np.argmin(load))
        bins[r].append(int(i))
        load[r] += costs[i]
    return [np.array(sorted(b), np.int64) for b in bins]


def active_world():
    """(rank, world_size) of the initialised default process group; (0, 1) without one."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def block_range(n: int, rank: int, world: int):
    """Contiguous [lo, hi) share of n items for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_hyp_index(utt_sel: Sequence[int], n_best: int) -> np.ndarray:
    """Position of every hypothesis of the selected utterances in the utterance-major list of the
    WHOLE workload (utterance u, candidate k -> u * n_best + k): where a rank's scores land after
    gather_scores, whatever the partition."""
    utt_sel = np.asarray(utt_sel, np.int64)
    return np.repeat(utt_sel, n_best) * n_best + np.tile(np.arange(n_best, dtype=np.int64), len(utt_sel))


def init_process_group(backend: str | None = None):
    import torch
    import torch.distributed as dist
    rank, world, local = dist_env()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def gather_scores(local_values: np.ndarray, local_index: np.ndarray, n_total: int, device=None) -> np.ndarray:
    """All ranks contribute (index, value) pairs; every rank gets the dense float64[n_total]
    array.  Values travel bit-exactly (no arithmetic), so results are identical for any
    world size."""
    import torch
    import torch.distributed as dist
    local_values = np.ascontiguousarray(local_values, np.float64)
    local_index = np.ascontiguousarray(local_index, np.int64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out = np.zeros(n_total, np.float64)
        out[local_index] = local_values
        return out
    world = dist.get_world_size()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    n_local = torch.tensor([len(local_values)], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local)
    n_max = max(int(c.item()) for c in counts)
    v = torch.zeros(n_max, dtype=torch.float64, device=device)
    ix = torch.full((n_max,), -1, dtype=torch.int64, device=device)
    v[:len(local_values)] = torch.from_numpy(local_values).to(device)
    ix[:len(local_index)] = torch.from_numpy(local_index).to(device)
    vs = [torch.zeros_like(v) for _ in range(world)]
    ixs = [torch.zeros_like(ix) for _ in range(world)]
    dist.all_gather(vs, v)
    dist.all_gather(ixs, ix)
    out = np.zeros(n_total, np.float64)
    for r in range(world):
        n = int(counts[r].item())
        out[ixs[r][:n].cpu().numpy()] = vs[r][:n].cpu().numpy()
    return out


def gather_blocks(local: np.ndarray, n_total: int, axis: int = 0) -> np.ndarray:
    """Inverse of block_range: every rank contributes its contiguous block along `axis`; every
    rank gets the concatenation (length n_total along that axis).  Bit-exact transport."""
    import torch
    import torch.distributed as dist
    rank, world = active_world()
    local = np.ascontiguousarray(np.moveaxis(local, axis, 0))
    if world == 1:
        return np.moveaxis(local, 0, axis)
    device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    sizes = [block_range(n_total, r, world)[1] - block_range(n_total, r, world)[0] for r in range(world)]
    assert local.shape[0] == sizes[rank], (local.shape, sizes, rank)
    n_max = max(sizes)
    buf = torch.zeros((n_max,) + local.shape[1:], dtype=torch.from_numpy(local[:0]).dtype, device=device)
    buf[:local.shape[0]] = torch.from_numpy(local).to(device)
    outs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    full = np.concatenate([outs[r][:sizes[r]].cpu().numpy() for r in range(world)], axis=0)
    return np.moveaxis(full, 0, axis)


def reduce_counts(counts: np.ndarray, device=None) -> np.ndarray:
    """Element-wise integer SUM over ranks (edit counts per weight + reference length)."""
    import torch
    import torch.distributed as dist
    counts = np.ascontiguousarray(counts, np.int64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return counts.copy()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.from_numpy(counts).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()
