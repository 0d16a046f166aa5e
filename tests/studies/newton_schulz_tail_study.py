"""Feasibility study (CPU, oracle only): could the fp64 Cholesky+inverse tail (57 % of the C2 step) be replaced by a
Newton-Schulz refinement of the previous Sigma?   Sigma_new = -1/2 eta2_new^-1,  P = -2 eta2,
    Y <- Y + Y (I - P Y),  Y0 = Sigma_old                  (all GEMM-shaped, no pivot chain)
Converges iff the spectral radius of E0 = I - P_new Sigma_old is < 1; the error squares each iteration.
For every SVI iteration of a C2-shaped run this prints rho(E0), the Robbins-Monro step, and the iterations needed
for |I - P Y|_F / sqrt(m) < 1e-9 with the residual in fp64 and the correction product Y E in fp32.

    python tests/studies/newton_schulz_tail_study.py [iters] [C2|C3]
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

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
cfg = sys.argv[2] if len(sys.argv) > 2 else "C2"
if cfg == "C3":      # StudentT(3), Matern-3/2, D = 64, m = 1024, B = 16384 (BASELINE configs[2]); labels: f(X) + t_3 noise
    n, D, m, B = 200_000, 64, 1024, 16384
    X, ys, Z, mbs, rng = make_problem(n, D, m, B, iters, seed=1)
    y = X @ rng.standard_normal(D) / np.sqrt(D) + 0.3 * rng.standard_t(3, n)
    model = O.SVGP(O.Kernel("matern32", scale=1.0 / np.sqrt(D)), O.StudentTLikelihood(3.0), O.AnalyticSVI(B), Z)
    ys = [y]
else:
    n, D, m, B = 200_000, 32, 512, 8192
    X, ys, Z, mbs, _ = make_problem(n, D, m, B, iters, seed=1)
    model = O.SVGP(O.Kernel("sqexp", scale=1.0 / np.sqrt(D)), O.LogisticLikelihood(), O.AnalyticSVI(B), Z)
X64 = X.astype(np.float64)

rows = []
orig = O.global_update


def patched(gp):
    Sold = gp.Sigma.copy()
    orig(gp)
    P = -2.0 * gp.eta2
    E0 = np.eye(m) - P @ Sold
    rad = np.max(np.abs(np.linalg.eigvals(E0)))
    k_needed, Y = None, Sold.copy()
    if rad < 1.0:
        for k in range(1, 30):
            E = np.eye(m) - P @ Y                                   # fp64 residual
            Y = Y + (Y.astype(np.float32) @ E.astype(np.float32)).astype(np.float64)   # fp32 correction
            Y = (Y + Y.T) / 2
            res = np.linalg.norm(np.eye(m) - P @ Y) / np.sqrt(m)
            if res < 1e-9:
                k_needed = k
                break
    err = np.linalg.norm(Y - gp.Sigma) / np.linalg.norm(gp.Sigma) if k_needed else float("nan")
    rows.append((rad, k_needed, err, np.linalg.cond(P)))


O.global_update = patched
t0 = time.time()
state = None
for t in range(iters):
    model, state = O.train(model, X64, ys[0], 1, minibatches=[mbs[t]], state=state)
    rad, k, err, cond = rows[-1]
    lr = (1.0 + (t + 1)) ** -0.51
    print(f"iter {t + 1:3d}  lr {lr:.3f}  rho(I - P_new Sigma_old) {rad:9.3e}  NS iterations to 1e-9: {k}  rel err vs exact {err:.2e}  cond(P) {cond:.2e}",
          flush=True)
print(f"{time.time() - t0:.0f} s")
