"""Latent-sharded run (torchrun, one rank per GPU) against the same model on one GPU: posterior / ELBO must agree.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_parity.py
Covers the NVLink peer-memory exchange (default) and, with AGP_NO_PEER=1, the NCCL all-gather fallback."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import agp_b200 as agp
from problems import rel_fro


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    rng = np.random.default_rng(11)
    n, D, m, B, iters = 20000, 16, 256, 2048, 6
    X = rng.standard_normal((n, D)).astype(np.float32)
    sc = 1.0 / np.sqrt(D)
    mbs = [rng.choice(n, B, replace=False).astype(np.int64) for _ in range(iters)]
    ok = True
    for case in ("mosvgp", "logisticsoftmax"):
        Q = 2 * world
        if case == "mosvgp":
            W = rng.standard_normal((D, Q))
            ys = [np.where(X @ W[:, t] + 0.1 * rng.standard_normal(n) >= 0, 1.0, -1.0) for t in range(Q)]
            Zs = [X[rng.permutation(n)[:m]].astype(np.float64) for _ in range(Q)]
            A = rng.standard_normal((Q, Q)); A /= np.linalg.norm(A, axis=1, keepdims=True)
            mk = lambda **kw: agp.MOSVGP(agp.SqExponentialKernel() @ agp.ScaleTransform(sc), [agp.LogisticLikelihood() for _ in range(Q)],
                                         agp.AnalyticSVI(B), Zs, A=A, precision="tf32x3", device=lr, **kw)
            y = ys
        else:
            y = np.argmax(X @ rng.standard_normal((D, Q)), axis=1) + 1
            Z = X[rng.permutation(n)[:m]].astype(np.float64)
            mk = lambda **kw: agp.SVGP(agp.SqExponentialKernel() @ agp.ScaleTransform(sc), agp.LogisticSoftMaxLikelihood(Q), agp.AnalyticSVI(B), Z,
                                       precision="tf32x3", device=lr, **kw)
        ms = mk(shard=(rank, world))
        ms, ss = agp.train(ms, X, y, iters, minibatches=mbs)
        e_s = agp.ELBO(ms, ss)
        yp_s = agp.predict_y(ms, X[:500])                      # collective: rows of the owned latents gathered on the host
        q0, ql = ms._latent_range()
        mine = [ms.posterior(q) for q in range(ql)]
        if rank == 0:
            m1 = mk()
            m1, s1 = agp.train(m1, X, y, iters, minibatches=mbs)
            e_1 = agp.ELBO(m1, s1)
            errs = [max(rel_fro(mine[q][0], m1.posterior(q0 + q)[0]), rel_fro(mine[q][1], m1.posterior(q0 + q)[1])) for q in range(ql)]
            yp_1 = agp.predict_y(m1, X[:500])
            same_pred = all(np.array_equal(a, b) for a, b in zip(yp_s, yp_1)) if isinstance(yp_1, list) else np.array_equal(yp_s, yp_1)
            good = max(errs) < 1e-9 and abs(e_s - e_1) <= 1e-9 * abs(e_1) and same_pred
            ok = ok and good
            print(f"[{case}] world={world} peer={getattr(ms, '_peer', False)}: max rel err (mu, Sigma) {max(errs):.2e}, ELBO {e_s:.6f} vs {e_1:.6f} -> {'OK' if good else 'MISMATCH'}", flush=True)
        dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
