"""CPU tests of the multi-GPU host logic with a world_size-2 gloo group: shard bounds, the single sum
all-reduce of the partial Grams, and the replicated solve."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as O
from koopfit.sharding import allreduce_sum_, budgets_for_rank, shard_bounds


def test_shard_bounds_cover_everything():
    for M, W in ((10, 3), (100000000, 8), (7, 8), (11999, 2)):
        spans = [shard_bounds(M, r, W) for r in range(W)]
        assert spans[0][0] == 0 and spans[-1][1] == M
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 3, 3)
    assert sorted(np.concatenate([budgets_for_rank(64, r, 8) for r in range(8)]).tolist()) == list(range(64))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)                      # same data on every rank
    M, n, m = 2001, 3, 2
    alpha = 2 * rng.random((M, n)) - 1
    u = 2 * rng.random((M, m)) - 1
    beta = np.clip(alpha @ (0.9 * np.eye(n)) + 0.1 * u @ rng.standard_normal((m, n)), -1, 1)
    prog = O.build_program(["poly"], [2], n)
    lo, hi = shard_bounds(M, rank, world)
    Px, Py = O.build_regressors("bilinear", prog, alpha[lo:hi], beta[lo:hi], u[lo:hi])
    G, C = O.gram(Px, Py)
    packed = torch.from_numpy(np.concatenate([G.ravel(), C.ravel()]))
    allreduce_sum_(packed)                              # the ONE collective of the fit
    P = G.shape[0]
    Gs, Cs = packed[:P * P].numpy().reshape(P, P), packed[P * P:].numpy().reshape(P, P)
    K = np.linalg.solve(Gs, Cs)                         # replicated solve
    q.put((rank, Gs, Cs, K))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_single_pass():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    M, n, m = 2001, 3, 2
    alpha = 2 * rng.random((M, n)) - 1
    u = 2 * rng.random((M, m)) - 1
    beta = np.clip(alpha @ (0.9 * np.eye(n)) + 0.1 * u @ rng.standard_normal((m, n)), -1, 1)
    prog = O.build_program(["poly"], [2], n)
    Px, Py = O.build_regressors("bilinear", prog, alpha, beta, u)
    G, C = O.gram(Px, Py)
    for rank, Gs, Cs, K in out:
        assert np.linalg.norm(Gs - G) <= 1e-13 * np.linalg.norm(G)
        assert np.linalg.norm(Cs - C) <= 1e-13 * np.linalg.norm(C)
    assert np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][3], out[1][3])   # ranks agree bit for bit


def _hook_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from koopfit.sharding import make_allreduce
    hook = make_allreduce(torch.device("cpu"))
    v = np.array([1.0 + rank, 10.0 * (rank + 1), -3.0])
    hook(v, 0)                                          # sum over ranks, in place
    w = np.array([float(rank), -float(rank)])
    hook(w, 1)                                          # max over ranks
    q.put((rank, v, w))
    dist.barrier()
    dist.destroy_process_group()


def test_qp_partition_hook_and_column_bounds():
    """The reduction hook the column-split lasso sweep hands to kf_set_qp_partition (sum / max, in place) on a
    world_size-2 gloo group, and the column blocks covering K."""
    from koopfit.sharding import column_bounds
    spans = [column_bounds(1464, r, 8) for r in range(8)]
    assert spans[0][0] == 0 and spans[-1][1] == 1464 and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_hook_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, v, w in out:
        assert np.array_equal(v, [3.0, 30.0, -6.0]) and np.array_equal(w, [1.0, 0.0])


class _FakeFitter:
    """Stands in for koopfit.Fitter in the column-split path: records the partition, calls the reduction hook the way the
    library does, and 'solves' by filling only its own columns of a known K (K[i, j, b] = 1000 b + i + j / 1000)."""

    def __init__(self):
        self.options, self.lo, self.hi, self.hook = {}, 0, 0, None

    def set_option(self, name, value):
        self.options[name] = value

    def set_qp_partition(self, lo, hi, allreduce=None):
        self.lo, self.hi, self.hook = lo, hi, allreduce

    def solve_dev(self, P, least_squares=False, t=None, **kw):
        assert self.hi > self.lo and self.options.get("qp_method") == 2 and not least_squares
        nt = len(t)
        v = np.array([float(self.hi - self.lo), 1.0])
        self.hook(v, 0)                                   # sum over ranks: total number of columns, number of ranks
        assert v[0] == P
        w = np.array([float(self.lo)])
        self.hook(w, 1)                                   # max over ranks
        K = np.zeros((P, P, nt), order="F")
        i, j, b = np.meshgrid(np.arange(P), np.arange(self.lo, self.hi), np.arange(nt), indexing="ij")
        K[:, self.lo:self.hi, :] = 1000.0 * b + i + j / 1000.0
        return {"K_all": K, "K": K[:, :, 0], "objective": np.arange(nt, dtype=float), "l1norm": np.ones(nt), "qp_gap": np.zeros(nt),
                "qp_iters": np.ones(nt, dtype=np.int32), "info": {"qp_capped": 0, "nranks": int(v[1]), "max_lo": float(w[0])}}


def _split_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from koopfit.sharding import _solve_column_split, column_bounds
    P, budgets = 11, np.array([1.0, 2.0, 3.0])
    f = _FakeFitter()
    res = _solve_column_split(f, P, budgets, rank, world, torch.device("cpu"), None)
    q.put((rank, res["K_all"], res["info"], (f.lo, f.hi), f.options.get("qp_method"), column_bounds(P, rank, world)))
    dist.barrier()
    dist.destroy_process_group()


def test_column_split_gathers_the_column_blocks():
    """sharding._solve_column_split on a world_size-2 gloo group with a fake fitter: every rank solves its own column block
    (uneven: 6 + 5 of 11 columns), the hook reduces across ranks, and every rank ends up with the complete K for all budgets;
    the partition and the solver choice are switched off again afterwards."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_split_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    P, nt = 11, 3
    i, j, b = np.meshgrid(np.arange(P), np.arange(P), np.arange(nt), indexing="ij")
    want = 1000.0 * b + i + j / 1000.0
    for rank, K_all, info, part, method, bounds in out:
        assert np.array_equal(K_all, want)
        assert info["nranks"] == 2 and info["max_lo"] == 6.0
        assert part == (0, 0) and method == 0              # restored
        assert bounds == ((0, 6) if rank == 0 else (6, 11))


def test_dealt_columns_match_the_library_kernel_formula():
    """The round-robin deal of the columns of K over the ranks (kf_qp_deal_cols_kernel, csrc/qp.cu) restated from its index
    arithmetic: a permutation whose rank blocks are exactly shard_bounds(P, r, world) and hold the columns r, r + world, ..."""
    from koopfit.sharding import dealt_columns, shard_bounds
    for P, R in ((1464, 8), (1464, 2), (17, 4), (5, 8), (252, 3)):
        base, extra = divmod(P, R)
        big = extra * (base + 1)
        kernel = []
        for p in range(P):                                  # the kernel's mapping position -> original column
            r = p // (base + 1) if p < big else extra + (p - big) // max(base, 1)
            k = p - (r * base + min(r, extra))
            kernel.append(r + k * R)
        assert kernel == dealt_columns(P, R)
        assert sorted(kernel) == list(range(P))
        for r in range(R):
            lo, hi = shard_bounds(P, r, R)
            assert kernel[lo:hi] == list(range(r, P, R))
