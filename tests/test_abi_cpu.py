"""CPU tests of the boundary: the library loads, exports every symbol include/agp_b200.h declares, fails loudly
without a GPU (no CPU fallback), and the host-side API validates its arguments like the reference does."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "agp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(agp_[a-zA-Z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(agp):
    lib = agp._lib.load()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"libagp_b200.so does not export {n}"
        assert n in agp._lib.SIGNATURES, f"ctypes binding lacks {n}"
    assert lib.agp_abi_version() == 1


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "augmentedgaussianprocesses.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "agp_oracle" not in txt and "oracle/" not in txt, f


def _has_gpu(agp):
    ctx = C.c_void_p()
    rc = agp._lib.load().agp_ctx_create(0, None, C.byref(ctx))
    if rc == 0:
        agp._lib.load().agp_ctx_destroy(ctx)
    return rc == 0


def test_no_cpu_fallback(agp):
    if _has_gpu(agp):
        pytest.skip("a GPU is present")
    m = agp.SVGP(agp.SqExponentialKernel(), agp.LogisticLikelihood(), agp.AnalyticSVI(10), np.random.randn(5, 2))
    with pytest.raises(agp.AGPError) as e:
        agp.train(m, np.random.randn(50, 2), np.sign(np.random.randn(50)), 2)
    assert e.value.code == agp._lib.AGP_ERR_CUDA
    with pytest.raises(agp.AGPError):
        agp.predict_y(m, np.random.randn(5, 2))


def test_constructor_and_argument_errors(agp):
    Z = np.random.randn(5, 2)
    k = 2.0 * agp.SqExponentialKernel() @ agp.ScaleTransform(10.0)
    assert (k.kind, k.scale, k.variance) == (0, 10.0, 2.0)
    assert agp.with_lengthscale(agp.Matern32Kernel(), 4.0).scale == 0.25
    with pytest.raises(TypeError):
        agp.SVGP(k, agp.LogisticLikelihood(), "not an inference", Z)
    mopt = agp.SVGP(k, agp.LogisticLikelihood(), agp.AnalyticVI(), Z, optimiser=True)     # SVGP.jl:33-44: `true` -> ADAM(0.01)
    assert isinstance(mopt.optimiser, agp.ADAM) and mopt.optimiser.eta == 0.01 and mopt.Zoptimiser is None
    with pytest.raises(NotImplementedError):
        agp.SVGP(k, agp.LogisticLikelihood(), agp.AnalyticVI(), Z, optimiser="descent")
    with pytest.raises(ValueError):
        agp.StudentTLikelihood(0.3)
    with pytest.raises(ValueError):
        agp.RobbinsMonro(0.4)
    i = agp.AnalyticVI()
    assert repr(i) == "Analytic Variational Inference" and i.rho == 1.0 and not agp.is_stochastic(i)
    i = agp.AnalyticSVI(5)
    assert repr(i) == "Analytic Stochastic Variational Inference" and agp.is_stochastic(i)
    m = agp.SVGP(k, agp.LogisticLikelihood(), agp.AnalyticSVI(100), Z)
    with pytest.raises(ValueError):  # training/training.jl:27-29
        agp.train(m, np.random.randn(50, 2), np.sign(np.random.randn(50)), 3)
    with pytest.raises(ValueError):
        agp.train(m, np.random.randn(50, 2), np.sign(np.random.randn(50)), 0)
    with pytest.raises(ValueError):
        agp.treat_labels(np.array([0, 1, 2]), agp.LogisticLikelihood())
    assert np.array_equal(agp.treat_labels(np.array([0, 1, 1]), agp.LogisticLikelihood()), [-1.0, 1.0, 1.0])
    l = agp.LogisticSoftMaxLikelihood(3)
    assert np.array_equal(agp.treat_labels(["b", "a", "c", "a"], l), [0, 1, 2, 1]) and l.class_mapping == ["b", "a", "c"]
    with pytest.raises(ValueError):
        agp.treat_labels([1, 2, 3, 4], agp.LogisticSoftMaxLikelihood(3))
    mu, S, e1, e2 = m.posterior(0)  # posterior.jl:29-37 before any training
    assert np.all(mu == 0) and np.array_equal(S, np.eye(5)) and np.array_equal(e2, -0.5 * np.eye(5))
    with pytest.raises(ValueError):   # VGP is full-batch (models/VGP.jl)
        agp.VGP(np.random.randn(6, 2), np.ones(6), k, agp.GaussianLikelihood(), agp.AnalyticSVI(3))
    with pytest.raises(ValueError):   # sample-count check (data/utils.jl)
        agp.VGP(np.random.randn(6, 2), np.ones(5), k, agp.GaussianLikelihood(), agp.AnalyticVI())


def test_precision_policy_tiers(agp):
    """precision="auto": the path kept for an error amplification sqrt(variance ||K_mm^-1||_inf) (api.AMPLIFICATION_LIMIT):
    tensor cores up to 30, fp32 CUDA cores up to 100, fp64 above; never a faster path than the shapes selected."""
    from agp_b200.api import AMPLIFICATION_LIMIT, precision_for_amplification as pick
    assert [p for p, _ in AMPLIFICATION_LIMIT] == ["tf32x3", "f32", "f64"]
    lims = dict(AMPLIFICATION_LIMIT)
    assert lims["tf32x3"] < lims["f32"] < lims["f64"] == float("inf")
    assert pick(1.0, "tf32x3") == "tf32x3" and pick(lims["tf32x3"], "tf32x3") == "tf32x3"
    assert pick(lims["tf32x3"] * 1.01, "tf32x3") == "f32" and pick(lims["f32"], "tf32x3") == "f32"
    assert pick(lims["f32"] * 1.01, "tf32x3") == "f64" and pick(1e9, "f32") == "f64"
    assert pick(1.0, "f32") == "f32"          # a small model (m < 128) is not moved onto the tensor cores
    assert pick(1.0, "f64") == "f64" and pick(50.0, "f64") == "f64"
