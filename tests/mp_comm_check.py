"""Run under torchrun with 2 ranks (tests/test_gpu_multi.py): one process per GPU, the NCCL communicator owned by
libkoopfit.so (kf_comm_unique_id on rank 0 -> broadcast -> kf_comm_init_rank); every rank passes ITS shard to kf_fit and
receives the K of the whole data set.  torch.distributed (gloo) is used only to broadcast the 128-byte id."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import koopfit
    import oracle as O
    from koopfit.sharding import shard_bounds

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    fit = koopfit.Fitter(device=local)
    ids = [koopfit.Fitter.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    fit.comm_init(world, rank, ids[0])
    assert fit.comm_info()["nranks"] == world

    rng = np.random.default_rng(0)
    M, n, m = 20011, 4, 2
    alpha = 2 * rng.random((M, n)) - 1
    u = 2 * rng.random((M, m)) - 1
    beta = np.clip(alpha @ (0.9 * np.linalg.qr(rng.standard_normal((n, n)))[0]).T + 0.1 * alpha * u[:, :1] + 0.01 * rng.standard_normal((M, n)), -1, 1)
    basis = koopfit.Basis(["poly"], [3], n)
    prog = O.build_program(["poly"], [3], n)
    Px, Py = O.build_regressors("bilinear", prog, alpha, beta, u)
    lo, hi = shard_bounds(M, rank, world)
    res = fit.fit(basis, "bilinear", alpha[lo:hi], beta[lo:hi], u[lo:hi], want_gram=True)     # this rank's shard
    G = Px.T @ Px
    assert np.linalg.norm(res["G"] - G) / np.linalg.norm(G) < 1e-13
    Ko = O.mldivide(Px, Py)
    assert np.linalg.norm(res["K"] - Ko) / np.linalg.norm(Ko) < 1e-9
    # device-resident shard through kf_fit_dev
    ta, tb, tu = (torch.tensor(np.ascontiguousarray(x[lo:hi].T), device=f"cuda:{local}") for x in (alpha, beta, u))
    torch.cuda.synchronize()
    r2 = fit.fit_dev(basis, "bilinear", hi - lo, n, m, ta.data_ptr(), tb.data_ptr(), tu.data_ptr())
    assert np.linalg.norm(r2["K"] - res["K"]) / np.linalg.norm(res["K"]) < 1e-12
    # identical on every rank (replicated deterministic solve)
    ks = [None] * world
    dist.all_gather_object(ks, res["K"])
    assert all(np.array_equal(ks[0], k) for k in ks)
    fit.close()
    dist.destroy_process_group()
    print("RANK_OK", rank, flush=True)


if __name__ == "__main__":
    main()
