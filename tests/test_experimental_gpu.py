"""GPU checks of EXPERIMENTAL code that is not on the product path (the opt-in second-generation tcgen05 GEMM and the
Newton-Schulz refinement built on it).  Written after the round's GPU budget was spent, hence never run on a B200 yet:
skipped unless AGP_EXPERIMENTAL=1 so that an untested kernel cannot take the default `-m gpu` suite down with it
(first contact: `bash tools/umma_v2_check.sh` under a gpurun timeout)."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("AGP_EXPERIMENTAL") != "1", reason="experimental kernels: set AGP_EXPERIMENTAL=1")]


def _spd(m, cond, rng):
    # identity + decaying spectrum: the shape of the whitened precision P_v = I + rho V^T theta V
    Q, _ = np.linalg.qr(rng.standard_normal((m, m)))
    ev = 1 + (cond - 1) * np.exp(-np.linspace(0.0, 0.1 * m, m))
    return (Q * ev) @ Q.T


# (m, cond(P), mode, tolerance after 3 iterations); expectations from profiles/r1/studies/newton_schulz_precision_study.txt:
# fp64 residual (mode 3): ~4e-7; 3xTF32 residual (mode 2): floors at ~eps_fp32 * cond (7e-4 at 1e4), fine at cond 1e2
@pytest.mark.parametrize("m,cond,mode,tol", [(128, 1e2, 2, 5e-5), (512, 1e2, 3, 5e-6), (512, 1e4, 3, 5e-6), (1024, 5e4, 3, 5e-6), (512, 1e4, 2, 5e-3)])
def test_newton_schulz_refine_matches_fp64_inverse(agp, m, cond, mode, tol):
    """Y <- Y + Y (I - P Y) from a 20 %-perturbed inverse: correction product on the tensor cores (3xTF32, second-generation
    GEMM kernel), residual on DMMA (mode bit 0) or 3xTF32, symmetrised at the end."""
    from agp_b200 import _lib as L

    lib = L.load()
    rng = np.random.default_rng(0)
    P = _spd(m, cond, rng)
    P = (P + P.T) / 2
    S = np.linalg.inv(P)
    # previous-step covariance: Y = S^1/2 (I + 0.2 N) S^1/2 with |N|_2 = 1, so I - P Y has spectral radius 0.2 (what a
    # Robbins-Monro step of ~0.2 leaves behind, profiles/r1/studies/newton_schulz_tail_study.txt)
    w, Q = np.linalg.eigh(S)
    Sh = (Q * np.sqrt(w)) @ Q.T
    N = rng.standard_normal((m, m))
    N = (N + N.T) / 2
    N /= np.linalg.norm(N, 2)
    Y = np.ascontiguousarray(Sh @ (np.eye(m) + 0.2 * N) @ Sh)
    assert np.max(np.abs(1.0 - np.linalg.eigvals(P @ Y))) < 0.25
    ctx = C.c_void_p()
    assert lib.agp_ctx_create(0, None, C.byref(ctx)) == L.AGP_OK
    try:
        resid = np.zeros(3)
        ms = C.c_double(0.0)
        Pc = np.ascontiguousarray(P)
        rc = lib.agp_experimental_ns_refine(ctx, m, Pc.ctypes.data_as(L.c_double_p), Y.ctypes.data_as(L.c_double_p), 3, mode,
                                            resid.ctypes.data_as(L.c_double_p), C.byref(ms))
        L.check(ctx, rc)
    finally:
        lib.agp_ctx_destroy(ctx)
    err = np.linalg.norm(Y - S) / np.linalg.norm(S)
    print(f"m={m} cond={cond:g} mode={mode}: residuals {resid}, rel err {err:.2e}, {ms.value * 1e3:.1f} us for 3 iterations")
    assert resid[0] > resid[1]
    assert err < tol


@pytest.mark.parametrize("lik", ["logistic", "studentt"])
def test_newton_schulz_tail_parity(agp, lik, monkeypatch):
    """AnalyticSVI with the experimental tail (AGP_UMMA_V2=1, AGP_TAIL_NS=4: Cholesky for the first 8 steps, then four
    Newton-Schulz refinements of the previous Sigma_v per step, statistics against the full covariance) against the fp64
    oracle, same tolerance as the product tf32x3 path.  The environment is read when the engine is created."""
    from problems import make_data, oracle_lik, engine_lik, oracle_kernel, engine_kernel, rel_fro
    import agp_oracle as O

    monkeypatch.setenv("AGP_UMMA_V2", "1")
    # B = 1024 is noisy: the spectral radius of I - P_new Sigma_old is still 0.3 (logistic) / 0.5 (studentt) after step 8
    # (CPU check with the oracle), so 4 refinements per step and a loose Frobenius acceptance threshold
    monkeypatch.setenv("AGP_TAIL_NS", "4")
    monkeypatch.setenv("AGP_TAIL_NS_TOL", "1.0")
    n, D, m, B, iters = 8192, 8, 256, 1024, 24
    scale = 1.0 / np.sqrt(D)
    X, y, Z, mbs, F, rng = make_data(lik, n, D, m, B, iters, seed=0)
    mo = O.SVGP(oracle_kernel(O, "sqexp", scale, 1.0), oracle_lik(O, lik), O.AnalyticSVI(B), Z)
    mo, so = O.train(mo, X, y, iters, minibatches=mbs)
    me = agp.SVGP(engine_kernel(agp, "sqexp", scale, 1.0), engine_lik(agp, lik), agp.AnalyticSVI(B), Z, precision="tf32x3")
    me, se = agp.train(me, X, y, iters, minibatches=mbs)
    mu, S, e1, e2 = me.posterior(0)
    gp = mo.f[0]
    print(lik, "mu", rel_fro(mu, gp.mu), "Sigma", rel_fro(S, gp.Sigma))
    assert rel_fro(mu, gp.mu) < 5e-4
    assert rel_fro(S, gp.Sigma) < 5e-4
    elbo_o, elbo_e = mo.ELBO(so, so["y_batch"]), agp.ELBO(me, se)
    assert abs(elbo_e - elbo_o) <= 5e-4 * max(1.0, abs(elbo_o)) * 5, (elbo_e, elbo_o)
