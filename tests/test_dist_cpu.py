"""world_size-2 gloo test (CPU) of the host-side logic of the latent-sharded step: latent partitioning and the in-place
all-gather layout of the [Q][ldB] moment arrays + the scalar ELBO all-reduce.  (The kernels need a GPU; the collective
plumbing and index arithmetic do not.)"""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, Q, ld, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import agp_b200 as agp

    class M:  # the attributes _latent_range / _allgather_rows use
        n_latent = Q
        _latent_range = agp.api.AbstractGPModel._latent_range

    m = M()
    m.rank, m.world = rank, world
    q0, ql = m._latent_range()
    mean = torch.full((Q * ld,), -1.0, dtype=torch.float64)
    var = torch.full((Q * ld,), -1.0, dtype=torch.float64)
    for q in range(q0, q0 + ql):  # what agp_step_moments_async writes: the owned rows
        mean[q * ld : (q + 1) * ld] = q + 0.25 * torch.arange(ld, dtype=torch.float64)
        var[q * ld : (q + 1) * ld] = 100 + q
    agp.api._allgather_rows((mean, var), q0, ql, ld)
    kl = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(kl)
    ok = all(torch.equal(mean[q * ld : (q + 1) * ld], q + 0.25 * torch.arange(ld, dtype=torch.float64)) for q in range(Q))
    ok = ok and all(bool((var[q * ld : (q + 1) * ld] == 100 + q).all()) for q in range(Q))
    ok = ok and float(kl) == sum(range(1, world + 1)) and (q0, ql) == (rank * Q // world, Q // world)
    out[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("Q", [2, 8])
def test_sharded_allgather_gloo(Q):
    world = 2
    port = 29600 + (os.getpid() + Q) % 300
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, Q, 12, out), nprocs=world, join=True)
        assert out[0] and out[1]


def test_latent_partition_errors():
    import agp_b200 as agp

    Z = np.random.randn(4, 2)
    m = agp.SVGP(agp.SqExponentialKernel(), agp.LogisticSoftMaxLikelihood(3), agp.AnalyticSVI(8), Z, shard=(0, 2))
    with pytest.raises(ValueError):
        m._latent_range()
    m = agp.SVGP(agp.SqExponentialKernel(), agp.LogisticSoftMaxLikelihood(4), agp.AnalyticSVI(8), Z, shard=(1, 2))
    assert m._latent_range() == (2, 2)
