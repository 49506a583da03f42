"""world_size-2 gloo tests (CPU) of the multi-rank host logic: contiguous sample sharding, the single gather
and the assembly on rank 0.  The CUDA compute / summary entry points are replaced by the oracle here (this
box has no GPU); the real kernels are covered by the ``-m gpu`` tests."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs(S, n=12, g=5, c=5):
    rng = np.random.default_rng(1)
    coords = rng.random((n, 2))
    grid = rng.random((g, 2))
    model = {"phi": rng.uniform(3, 9, S), "sig2": rng.uniform(0.5, 2, S), "z": rng.standard_normal((S, n))}
    eps = rng.standard_normal((S, g, c))
    return coords, grid, model, eps


def _oracle_compute(coords_d, grid_d, cov_type, phi_d, sig2_d, z_d, eps_d):
    from oracle.gradient import gradient_draws
    d = gradient_draws(coords_d.numpy().T, grid_d.numpy().T, cov_type, phi_d.numpy(), sig2_d.numpy(), z_d.numpy(),
                       eps_d.numpy())
    return torch.from_numpy(np.ascontiguousarray(d.transpose(0, 2, 1)))      # [S][c][G]


def _oracle_summarise(draws, nbatch, prob):
    from oracle.summary import summarise_draws
    d = draws.numpy()
    return torch.from_numpy(np.stack([summarise_draws(d[:, k, :], nbatch, prob).T for k in range(d.shape[1])]))


def _to_r(draws):
    return draws.permute(1, 2, 0).contiguous()


def _worker(rank, world, port, S, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from spwombling_b200 import device as dev
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    coords, grid, model, eps = _inputs(S)
    out = dev.spatial_gradient_sharded(coords, model, "matern2", grid, nbatch=3, eps=eps, device=torch.device("cpu"),
                                       rounded=False, compute=_oracle_compute, summarise=_oracle_summarise,
                                       to_r_layout=_to_r)
    if rank == 0:
        q.put({k: np.asarray(v) for k, v in out.items()})
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("S", [12, 13])      # even shards (one gather collective) and ragged shards (send/recv)
def test_sharded_equals_single_process(S):
    import oracle
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, S, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    coords, grid, model, eps = _inputs(S)
    ref = oracle.spatial_gradient(coords, model, "matern2", grid, nbatch=3, eps=eps, rounded=False)
    assert list(got) == list(ref)
    for k in ref:
        assert np.array_equal(got[k], ref[k]), k      # sharding by sample changes no arithmetic


def test_shard_bounds_cover_all_samples():
    from spwombling_b200.device import shard_bounds
    for S in (1, 7, 5000, 5001):
        for world in (1, 2, 4, 8):
            edges = [shard_bounds(S, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == S
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
