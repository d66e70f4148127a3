"""Snapshot sharding for the multi-GPU fit.

The fit shards naturally over snapshots: G = sum_r Px_r' Px_r, C = sum_r Px_r' Py_r.  Rank r owns a
contiguous slice of the snapshot pairs, accumulates its partial Gram tiles on its own GPU, and ONE
all-reduce (sum) of the packed accumulator makes every rank hold the full (G, C); the solve is then
replicated (deterministic, no broadcast).  There is no other data-path collective.

The collective goes through `torch.distributed` (NCCL on GPUs; gloo in the CPU tests of the host logic).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(M, rank, world):
    """Contiguous, balanced slice [lo, hi) of M snapshot pairs owned by `rank` (first M % world ranks get one more)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(int(M), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def dealt_columns(P, world):
    """Column order of the lasso sweep split inside the library (csrc/qp.cu: kf_qp_deal_cols_kernel): position p of the dealt
    order holds original column dealt_columns(P, world)[p]; rank r owns the positions shard_bounds(P, r, world), i.e. the original
    columns r, r + world, r + 2 world, ... — a round-robin deal, so that column ranges with systematically different support
    sizes (config 3a: the psi and the u psi halves of K) spread over all ranks."""
    order = []
    for r in range(int(world)):
        order.extend(range(r, int(P), int(world)))
    return order


def allreduce_sum_(tensor, group=None):
    """In-place sum over ranks of the packed partial-Gram accumulator; no-op for a single process."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor


def budgets_for_rank(nt, rank, world):
    """Round-robin split of a lasso vector over ranks (indices of the budgets rank solves)."""
    return np.arange(rank, nt, world)


class DeviceArrayView:
    """Zero-copy view of a raw device pointer for torch.as_tensor (via __cuda_array_interface__)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def fit_sharded(fitter, basis, model_type, alpha_dev, beta_dev, u_dev, P=None, budgets=None, group=None, split="budgets", pc_cols=0,
                **solve_kw):
    """One rank's part of a snapshot-sharded fit (one process per GPU).

    alpha_dev / beta_dev / u_dev: this rank's shard as CUDA tensors of shape (nzeta, M_r) / (m, M_r), float64,
    contiguous (= column-major M_r x nzeta).  Steps: local lift + Gram (kf_accumulate_dev), ONE all-reduce of the
    packed partial Grams, then either the replicated LS solve or — for a lasso vector — this rank's round-robin
    share of the budgets (all ranks hold the same G, C), gathered so that every rank returns all candidates.
    split = "columns": the exact active-set solver instead, every rank solving a block of COLUMNS of K for all budgets
    (the right split for the regularisation-path solver: a budget split would make every rank walk the whole path).
    Returns the dict of Fitter.solve_dev with K_all ordered like `budgets`.
    """
    import torch
    import torch.distributed as dist

    nzeta, M = alpha_dev.shape
    m = u_dev.shape[0]
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0

    def data_pass(reset):
        fitter.accumulate_dev(basis, model_type, M, nzeta, m, alpha_dev.data_ptr(), beta_dev.data_ptr(), u_dev.data_ptr(),
                              reset=reset, pc_cols=pc_cols)
        if world > 1:
            ptr, n = fitter.accum_buffer()
            fitter.sync()
            allreduce_sum_(torch.as_tensor(DeviceArrayView(ptr, n), device=alpha_dev.device), group)
            torch.cuda.synchronize(alpha_dev.device)

    data_pass(True)
    P = fitter._acc_dims[0]      # the regressor width the library accumulated with (a caller-supplied P is only checked)
    if budgets is None:
        # least squares: an ill-conditioned regressor makes the library ask for extra data passes (solve_dev returns None:
        # KF_EAGAIN, Gram-route refinement); every rank takes the same decision, so the collectives stay matched
        while True:
            res = fitter.solve_dev(P, **solve_kw)
            if res is not None:
                return res
            data_pass(False)
    budgets = np.atleast_1d(np.asarray(budgets, dtype=np.float64))
    if world > 1 and split == "columns":
        return _solve_column_split(fitter, P, budgets, rank, world, alpha_dev.device, group, **solve_kw)
    mine = budgets_for_rank(budgets.size, rank, world)
    res = fitter.solve_dev(P, least_squares=False, t=budgets[mine] if mine.size else budgets[:1], **solve_kw)
    if world == 1:
        return res
    parts = [None] * world
    dist.all_gather_object(parts, (mine, res["K_all"][:, :, :mine.size], res["objective"][:mine.size], res["l1norm"][:mine.size],
                                   res["qp_gap"][:mine.size]), group=group)
    K_all = np.zeros((P, P, budgets.size), order="F")
    obj, l1, gap = np.zeros(budgets.size), np.zeros(budgets.size), np.zeros(budgets.size)
    for idx, Kp, ob, ln, gp in parts:
        for j, i in enumerate(idx):
            K_all[:, :, i], obj[i], l1[i], gap[i] = Kp[:, :, j], ob[j], ln[j], gp[j]
    res.update(K_all=K_all, K=K_all[:, :, 0], objective=obj, l1norm=l1, qp_gap=gap)
    return res


def column_bounds(P, rank, world):
    """Contiguous block [lo, hi) of the P columns of K that `rank` solves in a column-split lasso sweep."""
    return shard_bounds(P, rank, world)


def make_allreduce(device, group=None):
    """The reduction hook of kf_set_qp_partition on top of torch.distributed: values (a small float64 numpy view)
    are reduced in place over the ranks; op 0 = sum, 1 = max.  NCCL needs device tensors, gloo takes host ones."""
    import torch
    import torch.distributed as dist
    on_gpu = dist.get_backend(group) == "nccl"

    buf = torch.zeros(64, dtype=torch.float64, device=device if on_gpu else "cpu")     # reused: the hook runs a few times per step

    def hook(values, op):
        n = values.size
        tns = buf[:n] if n <= buf.numel() else torch.zeros(n, dtype=torch.float64, device=buf.device)
        tns.copy_(torch.from_numpy(values))
        dist.all_reduce(tns, op=dist.ReduceOp.MAX if op == 1 else dist.ReduceOp.SUM, group=group)
        values[:] = tns.cpu().numpy()
    return hook


def _solve_column_split(fitter, P, budgets, rank, world, device, group, **solve_kw):
    """Lasso sweep split across the GPUs by COLUMNS of K (exact active-set solver): every rank holds the same
    (G, C) after the all-reduce, factors only its own columns for all budgets and exchanges a few scalars per step;
    the column blocks of K are gathered at the end."""
    import torch
    import torch.distributed as dist
    lo, hi = column_bounds(P, rank, world)
    # the cases the column-split solver refuses are the same on every rank: fail BEFORE the first collective, everywhere
    if P > 4096 or solve_kw.get("delay_constraint"):
        raise ValueError("column-split lasso sweep: P <= 4096 and no pinned delay columns (linear model with delays) required")
    prev_method = getattr(fitter, "_qp_method", 0)
    fitter.set_option("qp_method", 2)
    fitter.set_qp_partition(lo, hi, make_allreduce(device, group))
    err = None
    try:
        res = fitter.solve_dev(P, least_squares=False, t=budgets, **solve_kw)
    except Exception as exc:                    # a rank-local failure must not leave the others blocked in a collective
        err, res = exc, None
    finally:
        fitter.set_qp_partition(0, 0)
        fitter.set_option("qp_method", prev_method)
    on_gpu0 = dist.get_backend(group) == "nccl"
    flag = torch.tensor([1.0 if err is not None else 0.0], dtype=torch.float64, device=device if on_gpu0 else "cpu")
    dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
    if float(flag.item()) > 0:
        raise RuntimeError(f"column-split lasso sweep failed on at least one rank (this rank: {err!r})")
    nt = budgets.size
    wmax = max(column_bounds(P, r, world)[1] - column_bounds(P, r, world)[0] for r in range(world))
    mine = np.zeros((nt, wmax, P))                       # [budget][column][row]: contiguous column blocks
    mine[:, :hi - lo, :] = np.transpose(res["K_all"][:, lo:hi, :], (2, 1, 0))
    on_gpu = dist.get_backend(group) == "nccl"
    src = torch.from_numpy(mine).to(device) if on_gpu else torch.from_numpy(mine)
    parts = [torch.empty_like(src) for _ in range(world)]
    dist.all_gather(parts, src, group=group)
    K_all = np.zeros((P, P, nt), order="F")
    for r, part in enumerate(parts):
        a, b = column_bounds(P, r, world)
        K_all[:, a:b, :] = np.transpose(part.cpu().numpy()[:, :b - a, :], (2, 1, 0))
    res.update(K_all=K_all, K=K_all[:, :, 0])
    return res
