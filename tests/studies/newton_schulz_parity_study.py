"""Parity impact (CPU, oracle only) of replacing the fp64 Cholesky tail by an fp32 Newton-Schulz refinement of the
previous whitened covariance:  after `t0` exact iterations,  Y <- Y + Y (I - P_v Y)  (k_ns times, all fp32; the
tensor-core path would run it as 3xTF32 GEMMs),  P_v = L^T (-2 eta2) L,  and the NEXT step's moments use
Sigma = L Y L^T  and  mu = Sigma eta1.  The natural parameters stay exact functions of what the steps produced, so an
inexact Y only perturbs the local variables of the following step.  Prints, after `iters` iterations of the C2 workload
(n = 1e6, D = 32, m = 512, B = 8192, Logistic), the deviation of mu, Sigma (both recomputed exactly from eta at the
end, as the getters do) and of the ELBO from the exact oracle run.

    python tests/studies/newton_schulz_parity_study.py [iters] [t0] [k_ns]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import agp_oracle as O  # noqa: E402
from bench import make_problem  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
t0 = int(sys.argv[2]) if len(sys.argv) > 2 else 8
k_ns = int(sys.argv[3]) if len(sys.argv) > 3 else 3
n, D, m, B = 1_000_000, 32, 512, 8192
X, ys, Z, mbs, _ = make_problem(n, D, m, B, iters, seed=0)
X64 = X.astype(np.float64)


def run(ns: bool):
    model = O.SVGP(O.Kernel("sqexp", scale=1.0 / np.sqrt(D)), O.LogisticLikelihood(), O.AnalyticSVI(B), Z)
    exact = O.global_update
    st = dict(t=0, Y=None, L=None, worst=0.0, rad=0.0)

    def patched(gp):
        st["t"] += 1
        exact(gp)
        if not ns:
            return
        L = st["L"]
        if L is None:
            return
        Pv = L.T @ (-2.0 * gp.eta2) @ L
        if st["t"] <= t0 or st["Y"] is None:
            st["Y"] = np.linalg.inv(Pv).astype(np.float32)       # Cholesky tail (exact) while the step size is large
            return
        P32, Y, I32 = Pv.astype(np.float32), st["Y"], np.eye(m, dtype=np.float32)
        st["rad"] = max(st["rad"], float(np.max(np.abs(1.0 - np.linalg.eigvals(Pv @ Y.astype(np.float64))))))
        for _ in range(k_ns):
            Y = Y + Y @ (I32 - P32 @ Y)
        st["Y"] = Y
        Sig = L @ Y.astype(np.float64) @ L.T
        st["worst"] = max(st["worst"], np.linalg.norm(Sig - gp.Sigma) / np.linalg.norm(gp.Sigma))
        gp.Sigma = Sig                                          # what the next step's mean_f / var_f will see
        gp.mu = Sig @ gp.eta1

    O.global_update = patched
    state = None
    try:
        for t in range(iters):
            model, state = O.train(model, X64, ys[0], 1, minibatches=[mbs[t]], state=state)
            if st["L"] is None:
                st["L"] = state["kernel_matrices"][0]["L"]
    finally:
        O.global_update = exact
    gp = model.f[0]
    exact(gp)                                                   # getters: exact Sigma, mu from the final natural parameters
    elbo = model.ELBO(state, ys[0][mbs[iters - 1]])
    return gp.mu.copy(), gp.Sigma.copy(), elbo, st


t_start = time.time()
mu0, S0, e0, _ = run(False)
mu1, S1, e1, st = run(True)
rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
print(f"C2, {iters} iterations, exact tail for the first {t0}, then {k_ns} fp32 Newton-Schulz iterations per step")
print(f"  largest spectral radius of I - P_v Y_old at the start of a refinement : {st['rad']:.3f}   (must be < 1)")
print(f"  worst per-step rel. Frobenius error of the refined Sigma : {st['worst']:.2e}")
print(f"  final mu    rel. Frobenius deviation from the exact run  : {rel(mu1, mu0):.2e}")
print(f"  final Sigma rel. Frobenius deviation from the exact run  : {rel(S1, S0):.2e}")
print(f"  final ELBO  {e1:.6f} vs {e0:.6f}  rel. {abs(e1 - e0) / abs(e0):.2e}   (tf32x3 parity tolerance: 5e-4)")
print(f"{time.time() - t_start:.0f} s")
