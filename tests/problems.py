"""Seeded synthetic problems shared by the oracle tests, the GPU parity tests, smoke() and bench.py.
Both the oracle (oracle/agp_oracle.py) and the engine (agp_b200) are built from the same spec."""
from __future__ import annotations

import numpy as np

KINDS = {"sqexp": 0, "matern32": 1, "matern52": 2}


def make_data(lik: str, n: int, D: int, m: int, B: int, iters: int, seed: int = 0, n_class: int = 3, n_task: int = 1):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, D))
    w = rng.standard_normal((D, max(n_class, n_task, 1)))
    F = np.sin(X @ w / np.sqrt(D) * 1.5) + 0.3 * (X @ w) / np.sqrt(D)
    if lik == "logistic":
        y = np.sign(F[:, 0] + 0.1 * rng.standard_normal(n))
        y[y == 0] = 1.0
    elif lik == "gaussian":
        y = F[:, 0] + 0.1 * rng.standard_normal(n)
    elif lik == "studentt":
        y = F[:, 0] + 0.1 * rng.standard_t(3.0, n)
    elif lik == "logisticsoftmax":
        y = np.argmax(F[:, :n_class] + 0.1 * rng.standard_normal((n, n_class)), axis=1) + 1
    elif lik == "laplace":
        y = F[:, 0] + rng.laplace(0.0, 0.2, n)
    elif lik == "bayesiansvm":
        y = np.sign(F[:, 0] + 0.1 * rng.standard_normal(n))
        y[y == 0] = 1.0
    elif lik == "negbinomial":
        y = rng.negative_binomial(5, 1.0 / (1.0 + np.exp(F[:, 0]))).astype(np.int64)   # p(success) = sigma(-f)
    elif lik == "poisson":
        y = rng.poisson(4.0 / (1.0 + np.exp(-F[:, 0]))).astype(np.int64)
    elif lik == "heteroscedastic":
        y = F[:, 0] + rng.standard_normal(n) * 0.2 * np.exp(0.5 * F[:, 1])
    elif lik == "mo":
        y = None
    else:
        raise ValueError(lik)
    Z = X[rng.permutation(n)[:m]].copy()
    mbs = [rng.choice(n, B, replace=False).astype(np.int64) for _ in range(iters)]
    return X, y, Z, mbs, F, rng


def oracle_lik(O, lik: str, n_class: int = 3):
    return {
        "logistic": lambda: O.LogisticLikelihood(),
        "gaussian": lambda: O.GaussianLikelihood(1e-2),
        "studentt": lambda: O.StudentTLikelihood(3.0, 1.0),
        "logisticsoftmax": lambda: O.LogisticSoftMaxLikelihood(n_class),
        "laplace": lambda: O.LaplaceLikelihood(0.5),
        "bayesiansvm": lambda: O.BayesianSVM(),
        "negbinomial": lambda: O.NegBinomialLikelihood(5),
        "poisson": lambda: O.PoissonLikelihood(3.0),
        "heteroscedastic": lambda: O.HeteroscedasticLikelihood(2.0),
    }[lik]()


def engine_lik(agp, lik: str, n_class: int = 3):
    return {
        "logistic": lambda: agp.LogisticLikelihood(),
        "gaussian": lambda: agp.GaussianLikelihood(1e-2),
        "studentt": lambda: agp.StudentTLikelihood(3.0, 1.0),
        "logisticsoftmax": lambda: agp.LogisticSoftMaxLikelihood(n_class),
        "laplace": lambda: agp.LaplaceLikelihood(0.5),
        "bayesiansvm": lambda: agp.BayesianSVM(),
        "negbinomial": lambda: agp.NegBinomialLikelihood(5),
        "poisson": lambda: agp.PoissonLikelihood(3.0),
        "heteroscedastic": lambda: agp.HeteroscedasticLikelihood(2.0),
    }[lik]()


def oracle_kernel(O, kind: str, scale: float, variance: float):
    return O.Kernel(kind, scale=scale, variance=variance)


def engine_kernel(agp, kind: str, scale: float, variance: float):
    return agp.Kernel(KINDS[kind], scale, variance)


def rel_fro(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
