"""Generate the golden fixtures under tests/golden/ with the fp64 oracle (oracle/agp_oracle.py).

    python tests/golden/make_golden.py

PARITY UNPINNED: the reference (Julia) cannot run in this image and ships no golden mu/Sigma/ELBO vectors, so
these fixtures pin the ORACLE (a restatement by code reading), not the Julia package.  They guard against
regressions of the oracle and give the GPU tests a file-based target that does not depend on importing it.
One .npz per case: inputs (X, y, Z, minibatches, kernel/likelihood parameters) and outputs after `iters`
AnalyticSVI / AnalyticVI iterations (mu, Sigma, eta1, eta2 per latent, ELBO on the last minibatch, predictive
mean/variance on the first 64 rows).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import agp_oracle as O  # noqa: E402
from problems import make_data, oracle_kernel, oracle_lik  # noqa: E402

CASES = [
    # name, lik, n, D, m, B, iters, kernel kind, scale, variance, stochastic, n_class
    ("c1_gaussian_svi", "gaussian_c1", 1000, 2, 16, 100, 20, "sqexp", 1.0 / np.sqrt(2.0), 1.0, True, 0),
    ("logistic_svi", "logistic", 800, 4, 32, 128, 10, "sqexp", 0.5, 2.0, True, 0),
    ("studentt_matern32_svi", "studentt", 800, 4, 32, 128, 10, "matern32", 0.5, 1.0, True, 0),
    ("logisticsoftmax_svi", "logisticsoftmax", 800, 4, 32, 128, 10, "sqexp", 0.5, 1.0, True, 4),
    ("logistic_avi", "logistic", 300, 3, 24, 300, 5, "matern52", 0.6, 1.0, False, 0),
    ("tf32_logistic_svi", "logistic", 4096, 8, 128, 256, 6, "sqexp", 1.0 / np.sqrt(8.0), 1.0, True, 0),
    # SURVEY 8 f2 likelihoods
    ("laplace_svi", "laplace", 600, 3, 24, 128, 8, "sqexp", 0.6, 1.0, True, 0),
    ("bayesiansvm_svi", "bayesiansvm", 600, 3, 24, 128, 8, "matern32", 0.6, 1.0, True, 0),
    ("negbinomial_svi", "negbinomial", 600, 3, 24, 128, 8, "sqexp", 0.6, 1.0, True, 0),
    ("poisson_svi", "poisson", 600, 3, 24, 128, 8, "sqexp", 0.6, 1.5, True, 0),
    ("heteroscedastic_svi", "heteroscedastic", 600, 3, 24, 128, 8, "sqexp", 0.6, 1.0, True, 0),
]


def run_case(name, lik, n, D, m, B, iters, kind, scale, variance, stoch, n_class):
    base = "gaussian" if lik == "gaussian_c1" else lik
    X, y, Z, mbs, F, rng = make_data(base, n, D, m, B, iters, seed=abs(hash(name)) % 1000 if False else len(name), n_class=max(n_class, 3))
    likelihood = O.GaussianLikelihood(1e-3) if lik == "gaussian_c1" else oracle_lik(O, lik, max(n_class, 3))
    inf = O.AnalyticSVI(B) if stoch else O.AnalyticVI()
    model = O.SVGP(oracle_kernel(O, kind, scale, variance), likelihood, inf, Z)
    model, state = O.train(model, X, y, iters, minibatches=mbs)
    elbo = model.ELBO(state, state["y_batch"])
    mu_p, var_p = O.predict_f(model, X[:64], cov=True)
    out = dict(
        X=X, y=np.asarray(y), Z=Z, minibatches=np.stack(mbs), lik=lik, kind=kind, scale=scale, variance=variance, stoch=stoch,
        n_class=n_class, iters=iters, B=B, elbo=elbo, pred_mu=mu_p, pred_var=var_p,
        mu=np.stack([gp.mu for gp in model.f]), Sigma=np.stack([gp.Sigma for gp in model.f]),
        eta1=np.stack([gp.eta1 for gp in model.f]), eta2=np.stack([gp.eta2 for gp in model.f]),
    )
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: ELBO {elbo:.6f}")


if __name__ == "__main__":
    only = sys.argv[1:]
    for c in CASES:
        if not only or c[0] in only:
            run_case(*c)
