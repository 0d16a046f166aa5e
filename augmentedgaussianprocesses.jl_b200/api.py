"""Host-side mirror of the reference API for the AnalyticVI / AnalyticSVI path.

Same names, argument meaning and error behaviour as theogf/AugmentedGaussianProcesses.jl
(`SVGP`, `MOSVGP`, `AnalyticVI`, `AnalyticSVI`, `RobbinsMonro`, `train!` -> `train`, `predict_f`,
`predict_y`, `proba_y`, `ELBO`), with every numerical operation delegated to the CUDA engine behind the
C ABI of include/agp_b200.h.  No arithmetic of the hot path happens in Python and there is no CPU
fallback.  Reference citations are paths under /root/reference/src.
"""
from __future__ import annotations

import ctypes as C
import os
import math
import warnings
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _lib as L

JITTER = {np.float64: 1e-4, np.float32: 1e-3, np.float16: 1e-2}  # functions/utils.jl:8-10


# --------------------------------------------------------------------------------------------------
# KernelFunctions.jl surface used by the reference call sites
# --------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class ScaleTransform:
    s: float = 1.0


@dataclass(frozen=True)
class Kernel:
    """variance * base(s*x, s*z);  `2.0 * SqExponentialKernel() @ ScaleTransform(10.0)` mirrors
    `2.0 * SqExponentialKernel() ∘ ScaleTransform(10.0)` (test/likelihood/logistic.jl:3)."""

    kind: int = L.KERNEL_SQEXP
    scale: float = 1.0
    variance: float = 1.0

    def __rmul__(self, v):
        return Kernel(self.kind, self.scale, self.variance * float(v))

    __mul__ = __rmul__

    def __matmul__(self, t: ScaleTransform):
        return Kernel(self.kind, self.scale * float(t.s), self.variance)


def SqExponentialKernel():
    return Kernel(L.KERNEL_SQEXP)


def Matern32Kernel():
    return Kernel(L.KERNEL_MATERN32)


def Matern52Kernel():
    return Kernel(L.KERNEL_MATERN52)


def transform(k: Kernel, t: ScaleTransform):
    return k @ t


def with_lengthscale(k: Kernel, lengthscale: float):
    return k @ ScaleTransform(1.0 / float(lengthscale))


# --------------------------------------------------------------------------------------------------
# Likelihoods (AnalyticVI-capable ones of likelihood/*.jl)
# --------------------------------------------------------------------------------------------------
class AbstractLikelihood:
    kind: int
    p0 = 0.0
    p1 = 0.0
    n_latent = 1


class GaussianLikelihood(AbstractLikelihood):
    """likelihood/gaussian.jl:10-24"""

    kind = L.LIK_GAUSSIAN

    def __init__(self, sigma2: float = 1e-3, opt_noise=False):
        if isinstance(opt_noise, bool) or opt_noise is None:   # gaussian.jl:18-24
            opt_noise = ADAM(0.05) if opt_noise else None
        if opt_noise is not None and not isinstance(opt_noise, ADAM):
            raise NotImplementedError("only ADAM is implemented for the noise optimiser")
        self.opt_noise = opt_noise
        self.sigma2 = float(sigma2)

    @property
    def p0(self):
        return self.sigma2

    def __repr__(self):
        return f"Gaussian likelihood (σ² = {self.sigma2})"


class LogisticLikelihood(AbstractLikelihood):
    """likelihood/logistic.jl:19 (BernoulliLikelihood(LogisticLink()))"""

    kind = L.LIK_LOGISTIC

    def __repr__(self):
        return "Bernoulli Likelihood with Logistic Link"


class StudentTLikelihood(AbstractLikelihood):
    """likelihood/studentt.jl:23-35"""

    kind = L.LIK_STUDENTT

    def __init__(self, nu: float, sigma: float = 1.0):
        if not nu > 0.5:
            raise ValueError("ν should be greater than 0.5")
        self.nu, self.sigma = float(nu), float(sigma)
        self.p0, self.p1 = self.nu, self.sigma

    def __repr__(self):
        return f"Student-t likelihood (ν={self.nu}, σ={self.sigma})"


class LaplaceLikelihood(AbstractLikelihood):
    """likelihood/laplace.jl:17-28 (a = beta^-2, p = 1/2)"""

    kind = L.LIK_LAPLACE

    def __init__(self, beta: float = 1.0):
        self.beta = float(beta)
        self.p0 = self.beta

    def __repr__(self):
        return f"Laplace likelihood (β={self.beta})"


class BayesianSVM(AbstractLikelihood):
    """likelihood/bayesiansvm.jl:19 (BernoulliLikelihood(SVMLink()))"""

    kind = L.LIK_BAYESIANSVM

    def __repr__(self):
        return "Bernoulli Likelihood with SVM Link"


class NegBinomialLikelihood(AbstractLikelihood):
    """likelihood/negativebinomial.jl:22-27"""

    kind = L.LIK_NEGBINOMIAL

    def __init__(self, r):
        self.r = r
        self.p0 = float(r)

    def __repr__(self):
        return f"Negative Binomial Likelihood (r = {self.r})"


class PoissonLikelihood(AbstractLikelihood):
    """likelihood/poisson.jl:16-26.  `lam` (l.invlink.λ[1]) is re-estimated by every local update (poisson.jl:80); the
    device holds the live value, this attribute is refreshed at the end of `train`."""

    kind = L.LIK_POISSON

    def __init__(self, lam: float = 1.0):
        self.lam = float(lam)

    @property
    def p0(self):
        return self.lam

    def __repr__(self):
        return f"Poisson Likelihood (λ = {self.lam})"


class HeteroscedasticLikelihood(AbstractLikelihood):
    """likelihood/heteroscedastic.jl:17-48: N(y | f, (λ σ(g))^-1), two latent GPs; λ re-estimated (heteroscedastic.jl:98)."""

    kind = L.LIK_HETEROSCEDASTIC
    n_latent = 2

    def __init__(self, lam: float = 1.0):
        self.lam = float(lam)

    @property
    def p0(self):
        return self.lam

    def __repr__(self):
        return "Gaussian likelihood with heteroscedastic noise"


class LogisticSoftMaxLikelihood(AbstractLikelihood):
    """likelihood/logisticsoftmax.jl:23 + likelihood/multiclass.jl:1-24"""

    kind = L.LIK_LOGISTICSOFTMAX

    def __init__(self, x):
        if isinstance(x, (int, np.integer)):
            self.n_class = int(x)
            self.class_mapping = None
            self.ind_mapping = None
        else:
            self.class_mapping = list(x)
            self.n_class = len(self.class_mapping)
            self.ind_mapping = {v: i for i, v in enumerate(self.class_mapping)}

    @property
    def n_latent(self):
        return self.n_class

    def __repr__(self):
        return f"Multiclass Likelihood ({self.n_class} classes, Logistic-SoftMax Link )"


def create_mapping(l: LogisticSoftMaxLikelihood, y):
    """likelihood/multiclass.jl:62-78"""
    K = l.n_latent
    if l.class_mapping is None:
        seen = []
        for v in y:
            if v not in seen:
                seen.append(v)
        l.class_mapping = seen
        ints = all(isinstance(v, (int, np.integer)) for v in seen)
        if len(seen) <= K and ints and set(seen) <= set(range(1, K + 1)):
            l.class_mapping = list(range(1, K + 1))
        elif len(seen) > K:
            raise ValueError(
                f"The number of unique labels in the data : {seen} is not of the same size then the predefined class number ; {K}"
            )
    l.ind_mapping = {v: i for i, v in enumerate(l.class_mapping)}
    return l.ind_mapping


def _class_indices(l: LogisticSoftMaxLikelihood, y) -> np.ndarray:
    """treat_labels! + create_one_hot (multiclass.jl:40-44, 81-94) as 0-based int32 class indices."""
    y = y.tolist() if isinstance(y, np.ndarray) else list(y)
    if l.ind_mapping is None:
        create_mapping(l, y)
    try:
        return np.fromiter((l.ind_mapping[v] for v in y), dtype=np.int32, count=len(y))
    except KeyError:
        raise ValueError("Some labels of y are not part of the expect labels") from None


def treat_labels(y, lik) -> np.ndarray:
    """likelihood/classification.jl:29-45, regression.jl:10-15, multiclass.jl:40-44"""
    if lik.kind in (L.LIK_POISSON, L.LIK_NEGBINOMIAL):  # likelihood/event.jl:7-13
        y = np.asarray(y)
        if not np.issubdtype(y.dtype, np.integer):
            raise TypeError("For event count target(s) should be integers")
        return np.ascontiguousarray(y, dtype=np.float64)
    if lik.kind in (L.LIK_LOGISTIC, L.LIK_BAYESIANSVM):
        y = np.asarray(y)
        if not (np.issubdtype(y.dtype, np.number) or y.dtype == bool):
            raise TypeError("For classification target(s) should be real valued (Bool, Integer or Float)")
        labels = sorted(int(v) for v in np.unique(y))
        if labels == [0, 1]:
            return np.sign(y.astype(np.float64) - 0.5)
        if labels == [-1, 1]:
            return np.ascontiguousarray(y, dtype=np.float64)
        raise ValueError("Labels of y should be binary {-1,1} or {0,1}")
    if lik.kind == L.LIK_LOGISTICSOFTMAX:
        return _class_indices(lik, y)
    y = np.asarray(y)
    if not np.issubdtype(y.dtype, np.number):
        raise TypeError("For regression target(s) should be real valued")
    return np.ascontiguousarray(y, dtype=np.float64)


# --------------------------------------------------------------------------------------------------
# Inference objects (inference/analyticVI.jl:1-52, inference/optimisers.jl:1-19)
# --------------------------------------------------------------------------------------------------
class RobbinsMonro:
    def __init__(self, kappa: float = 0.51, tau: float = 1.0):
        if not (0.5 < kappa <= 1):
            raise ValueError("κ should be in the interval (0.5,1]")
        if not tau > 0:
            raise ValueError("τ should be positive")
        self.kappa, self.tau = float(kappa), float(tau)


class Descent:
    """Optimisers.jl Descent(eta): constant step for the stochastic natural-gradient update (`AnalyticSVI(B; optimiser=Descent(0.1))`)."""

    def __init__(self, eta: float = 0.1):
        if not 0.0 < eta <= 1.0:
            raise ValueError("eta should be in (0, 1]")
        self.eta = float(eta)
        self.kappa, self.tau = 0.51, 1.0   # unused by the device when a constant step is set


class ADAM:
    """Optimisers.jl ADAM(eta = 0.001, beta = (0.9, 0.999)) as accepted by `MOSVGP(...; Aoptimiser)` (MOSVGP.jl:51)."""

    def __init__(self, eta: float = 0.001, beta=(0.9, 0.999), epsilon: float = 1e-8):
        self.eta, self.beta, self.epsilon = float(eta), (float(beta[0]), float(beta[1])), float(epsilon)

    # Optimisers.init / Optimisers.apply (host side: used for the handful of kernel parameters and the m x D inducing points)
    def init(self, x):
        return dict(mt=np.zeros_like(x, dtype=np.float64), vt=np.zeros_like(x, dtype=np.float64), bt=np.array(self.beta))

    def apply(self, st, g):
        b1, b2 = self.beta
        st["mt"] = b1 * st["mt"] + (1.0 - b1) * g
        st["vt"] = b2 * st["vt"] + (1.0 - b2) * g**2
        step = st["mt"] / (1.0 - st["bt"][0]) / (np.sqrt(st["vt"] / (1.0 - st["bt"][1])) + self.epsilon) * self.eta
        st["bt"] = st["bt"] * np.array(self.beta)
        return st, step


class AnalyticVI:
    """inference/analyticVI.jl:1-14.  `AnalyticVI()` = full batch, `AnalyticSVI(B)` = stochastic."""

    def __init__(self, eps: float = 1e-5, *, _optimiser=None, _batchsize: int = 0, _stoch: bool = False):
        self.eps = eps
        self.n_iter = 0
        self.stoch = _stoch
        self.batchsize = int(_batchsize)
        self.rho = 1.0
        self.HyperParametersUpdated = True
        self.optimiser = _optimiser

    def __repr__(self):
        return "Analytic" + (" Stochastic" if self.stoch else "") + " Variational Inference"


def AnalyticSVI(nMinibatch: int, eps: float = 1e-5, optimiser: Optional[RobbinsMonro] = None):
    """inference/analyticVI.jl:48-52"""
    optimiser = optimiser if optimiser is not None else RobbinsMonro()
    if not isinstance(optimiser, (RobbinsMonro, Descent)):
        raise NotImplementedError("only the RobbinsMonro and Descent variational optimisers are accelerated")
    return AnalyticVI(eps, _optimiser=optimiser, _batchsize=int(nMinibatch), _stoch=True)


def is_stochastic(i: AnalyticVI) -> bool:
    return i.stoch


# --------------------------------------------------------------------------------------------------
# engine handle
# --------------------------------------------------------------------------------------------------
class _Engine:
    """Owns one agp_ctx + agp_model.  Created lazily (the batch capacity is only known at train time)."""

    def __init__(self, desc_kwargs: dict, device: int, stream):
        lib = L.load()
        self.lib = lib
        self.ctx = C.c_void_p()
        # stream: None -> the library creates its own; 0 (the framework's default stream) -> cudaStreamLegacy (0x1)
        sp = None if stream is None else C.c_void_p(int(stream) if int(stream) != 0 else 1)
        rc = lib.agp_ctx_create(int(device), sp, C.byref(self.ctx))
        if rc != L.AGP_OK:
            raise L.AGPError(rc, "agp_ctx_create failed: no usable CUDA device (there is no CPU fallback)")
        self._keep = []
        d = L.ModelDesc()
        for k, v in desc_kwargs.items():
            if isinstance(v, np.ndarray):
                self._keep.append(v)
                if v.dtype == np.int32:
                    v = v.ctypes.data_as(L.c_int32_p)
                else:
                    v = v.ctypes.data_as(L.c_double_p)
            setattr(d, k, v)
        self.desc = d
        self.model = C.c_void_p()
        rc = lib.agp_model_create(self.ctx, C.byref(d), C.byref(self.model))
        try:
            L.check(self.ctx, rc)
        except Exception:
            lib.agp_ctx_destroy(self.ctx)
            self.ctx = None
            raise
        self.capacity = int(d.batch_capacity)

    def ck(self, rc):
        L.check(self.ctx, rc)

    def close(self):
        if getattr(self, "model", None):
            self.lib.agp_model_destroy(self.model)
            self.model = None
        if getattr(self, "ctx", None):
            self.lib.agp_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _x_args(X):
    """(array kept alive, void*, dtype code, layout code, n, D) for a host matrix, without copying when possible."""
    X = np.asarray(X)
    if X.ndim == 1:
        X = X[:, None]
    if X.dtype not in (np.float64, np.float32):
        X = X.astype(np.float64)
    if X.flags.c_contiguous:
        layout = L.LAYOUT_ROWMAJOR
    elif X.flags.f_contiguous:
        layout = L.LAYOUT_COLMAJOR
    else:
        X = np.ascontiguousarray(X)
        layout = L.LAYOUT_ROWMAJOR
    dt = L.DTYPE_F64 if X.dtype == np.float64 else L.DTYPE_F32
    return X, X.ctypes.data_as(C.c_void_p), dt, layout, X.shape[0], X.shape[1]


# --------------------------------------------------------------------------------------------------
# Models
# --------------------------------------------------------------------------------------------------
class AbstractGPModel:
    model_kind = L.MODEL_SVGP

    def _common_init(self, inference, verbose, atfrequency, optimiser, Zoptimiser, T, precision, device, stream, shard):
        if not isinstance(inference, AnalyticVI):
            raise TypeError(
                "The inference object should be of type `VariationalInference` : either `AnalyticVI` or `NumericalVI`"
                " (only AnalyticVI / AnalyticSVI are accelerated)"
            )
        # optimiser / Zoptimiser (SVGP.jl:33-44): `true` means ADAM(0.01) like the reference; only ADAM is implemented
        def _opt(o):
            if o is None or o is False:
                return None
            if o is True:
                return ADAM(0.01)
            if isinstance(o, ADAM):
                return o
            raise NotImplementedError("only ADAM is implemented for the hyper-parameter optimisers")
        self.optimiser, self.Zoptimiser = _opt(optimiser), _opt(Zoptimiser)
        self._hyperopt_state = None
        if T not in JITTER:
            raise TypeError("T must be np.float64 / np.float32 / np.float16")
        if precision != "auto" and precision not in L.PRECISIONS:
            raise ValueError(f"precision must be 'auto' or one of {sorted(L.PRECISIONS)}")
        self.inference = inference
        self.verbose = verbose
        self.atfrequency = atfrequency
        self.trained = False
        self.T = T
        self.jitter = JITTER[T]
        self.precision_requested = precision   # "auto": tcgen05 (tf32x3) when the shapes allow it, fp32 SIMT otherwise -- resolved per engine
        self.precision = precision
        self.device = device
        self.stream = stream
        self.rank, self.world = shard if shard is not None else (0, 1)
        if self.world > 1 and self.stream is None:
            # a sharded model shares the host framework's stream, so that its collectives (NCCL fallback, ELBO all-reduce)
            # are ordered with the engine's kernels without extra synchronisation
            try:
                import torch

                self.stream = torch.cuda.current_stream(device).cuda_stream
            except Exception:
                pass
        self._eng: Optional[_Engine] = None
        self._data_key = None
        self._n = 0

    # ---- lazily (re)create the engine with enough batch capacity, carrying the posterior over
    def _engine(self, capacity: int) -> _Engine:
        # capacity 0: whatever engine exists (prediction: the rows are chunked by the engine, no batch shape to honour)
        if self._eng is not None and capacity == 0:
            return self._eng
        if self._eng is not None and capacity <= self._eng.capacity:
            return self._eng
        old = self._eng
        saved = None
        if old is not None:
            saved = [self._get_posterior_raw(q) for q in range(self.n_latent_local)]
            cnt = self.counters()
            old.close()
        cap = int(capacity)
        if self.precision_requested == "auto":
            # the tcgen05 kernels tile m and B by 128; the engine pads both itself (Engine::mk, Engine::rowsK), so every model with at
            # least 128 inducing points takes the tensor-core path; smaller ones (e.g. the reference's 10-inducing-point tests) would
            # mostly multiply padding and take the fp32 SIMT path
            self.precision = "tf32x3" if self.m >= 128 else "f32"
            by_cond = getattr(self, "_precision_by_condition", None)
            if by_cond is not None:         # train() found K_mm too ill conditioned for the faster path (amplification())
                self.precision = by_cond
        if self.precision == "tf32x3":
            cap = (cap + 127) // 128 * 128
        self._eng = _Engine(self._desc(cap), self.device, self.stream)
        self._data_key = None
        if self.inference.stoch and isinstance(self.inference.optimiser, Descent):
            self._eng.ck(self._eng.lib.agp_set_step_size(self._eng.model, self.inference.optimiser.eta))
        for t, l in enumerate(self.likelihoods):
            on = getattr(l, "opt_noise", None)
            if on is not None:
                self._eng.ck(self._eng.lib.agp_set_noise_optimiser(self._eng.model, t, 1, on.eta, on.beta[0], on.beta[1], on.epsilon))
        ao = getattr(self, "A_opt", None)
        if ao is not None:
            self._eng.ck(self._eng.lib.agp_set_A_optimiser(self._eng.model, 1, ao.eta, ao.beta[0], ao.beta[1], ao.epsilon))
        self._peer = False
        if self.world > 1 and os.environ.get("AGP_NO_PEER", "0") != "1":
            self._peer = _attach_peers(self, self._eng)
        if any(l.kind in (L.LIK_POISSON, L.LIK_NEGBINOMIAL, L.LIK_BAYESIANSVM) for l in self.likelihoods):
            nodes, weights = _pred_nodes()   # `expectation` (functions/utils.jl:16-19) / compute_proba rule
            self._eng.ck(self._eng.lib.agp_set_quadrature(self._eng.model, L.dptr(nodes), L.dptr(weights), len(nodes)))
        if saved is not None:
            for q, (mu, S, e1, e2) in enumerate(saved):
                self._eng.ck(self._eng.lib.agp_set_posterior(self._eng.model, q, L.dptr(e1), L.dptr(e2)))
            self._eng.ck(self._eng.lib.agp_set_counters(self._eng.model, cnt[0], cnt[1]))
            self.inference.HyperParametersUpdated = True   # the new engine has no K_mm factor yet: the next train() call refreshes it
        return self._eng

    def _latent_range(self):
        Q = self.n_latent
        if Q % self.world:
            raise ValueError("the number of latent GPs must be divisible by the number of ranks")
        per = Q // self.world
        return self.rank * per, per

    @property
    def n_latent_local(self):
        return self._latent_range()[1]

    def _get_posterior_raw(self, q):
        e = self._eng
        m = self.m
        mu, e1 = np.empty(m), np.empty(m)
        S, e2 = np.empty((m, m)), np.empty((m, m))
        e.ck(e.lib.agp_get_posterior(e.model, q, L.dptr(mu), L.dptr(S), L.dptr(e1), L.dptr(e2)))
        return mu, S, e1, e2

    # ---- accessors mirroring mean(gp), cov(gp), nat1(gp), nat2(gp) (gpblocks/latentgp.jl:163-168)
    def posterior(self, q: int = 0):
        """(mu, Sigma, eta1, eta2) of the q-th OWNED latent."""
        if self._eng is None:
            m = self.m
            return np.zeros(m), np.eye(m), np.zeros(m), -0.5 * np.eye(m)  # posterior.jl:29-37
        return self._get_posterior_raw(q)

    def counters(self):
        e = self._eng
        t, c = C.c_int64(), C.c_int64()
        e.ck(e.lib.agp_get_counters(e.model, C.byref(t), C.byref(c)))
        return t.value, c.value

    def amplification(self) -> float:
        """sqrt(max_q variance_q * ||K_q^-1||_inf) over the owned latents, read from the engine's fp64 K_mm^-1 (agp_get_Kinv; needs
        agp_refresh_K).  V = K_nm L^-T multiplies the rounding error of the fp32-class K_nm entries (relative to the kernel variance) by
        ||L^-1||_2 = sqrt(||K^-1||_2) <= sqrt(||K^-1||_inf): the whitened iteration loses about log10 of this figure in digits on the
        fp32 / 3xTF32 paths, nothing in fp64."""
        e, m, amp = self._eng, self.m, 0.0
        Kinv, ld = np.empty((m, m)), C.c_double()
        for q in range(self.n_latent_local):
            e.ck(e.lib.agp_get_Kinv(e.model, q, L.dptr(Kinv), C.byref(ld)))
            k = self.kernels[self._latent_range()[0] + q]
            amp = max(amp, math.sqrt(float(k.variance) * float(np.abs(Kinv).sum(axis=1).max())))
        return amp

    def launch_count(self) -> int:
        return int(self._eng.lib.agp_launch_count(self._eng.model)) if self._eng else 0


class SVGP(AbstractGPModel):
    """models/SVGP.jl:22-80.  `SVGP(kernel, likelihood, inference, Z; optimiser=false, Zoptimiser=false)`.

    precision: "auto" (default: "tf32x3" when m >= 128, else "f32"), "f64" (fp64 SIMT, exact
    mode), "f32" (fp32 SIMT) or "tf32x3" (tcgen05 tensor cores).
    shard=(rank, world): own n_latent/world latents (LogisticSoftMax classes) on this process.
    """

    model_kind = L.MODEL_SVGP

    def __init__(self, kernel: Kernel, likelihood: AbstractLikelihood, inference: AnalyticVI, Z, *, verbose: int = 0,
                 optimiser=False, atfrequency: int = 1, mean=None, Zoptimiser=False, T=np.float64, precision: str = "auto",
                 device: int = 0, stream=None, shard=None):
        if not isinstance(likelihood, AbstractLikelihood):
            raise TypeError(f"The {likelihood} is not compatible or implemented with the {inference}")
        self._common_init(inference, verbose, atfrequency, optimiser, Zoptimiser, T, precision, device, stream, shard)
        self.likelihood = likelihood
        self.likelihoods = [likelihood]
        self.kernel = kernel
        Z = np.ascontiguousarray(np.asarray(Z, dtype=np.float64))
        if Z.ndim == 1:
            Z = Z[:, None]
        self.Z = Z
        self.m, self.D = Z.shape
        self.n_latent = likelihood.n_latent
        self.kernels = [kernel] * self.n_latent
        self.Zs = [Z] * self.n_latent
        if mean is None:
            self.mu0 = None
        elif np.isscalar(mean):
            self.mu0 = np.full(self.m, float(mean))  # ConstantMean (mean/constantmean.jl)
        else:
            raise NotImplementedError("only ZeroMean / ConstantMean priors cross the boundary (mu0 evaluated at Z)")
        self.A = None

    def _desc(self, capacity: int) -> dict:
        return _make_desc(self, capacity)

    def __repr__(self):
        return f"Sparse Variational Gaussian Process with a {self.likelihood} infered by {self.inference} "


class MOSVGP(AbstractGPModel):
    """models/MOSVGP.jl:22-115 with single-latent task likelihoods; A is T x Q (rows normalised like
    MOSVGP.jl:100-103).  `Aoptimiser=True` (ADAM(0.01), the reference default) runs update_A! on the device; default here: fixed A."""

    model_kind = L.MODEL_MOSVGP

    def __init__(self, kernel, likelihoods: Sequence[AbstractLikelihood], inference: AnalyticVI, Zs: Sequence, *, A=None,
                 verbose: int = 0, atfrequency: int = 1, mean=None, optimiser=False, Aoptimiser=False, Zoptimiser=False,
                 T=np.float64, precision: str = "auto", device: int = 0, stream=None, shard=None, rng=None):
        self._common_init(inference, verbose, atfrequency, optimiser, Zoptimiser, T, precision, device, stream, shard)
        if isinstance(Aoptimiser, bool) or Aoptimiser is None:   # MOSVGP.jl:79-81
            Aoptimiser = ADAM(0.01) if Aoptimiser else None
        if Aoptimiser is not None and not isinstance(Aoptimiser, ADAM):
            raise NotImplementedError("only ADAM (the reference default) is implemented for the mixing-matrix optimiser")
        self.A_opt = Aoptimiser
        if mean is not None:
            raise NotImplementedError("MOSVGP with a non-zero prior mean is not supported")
        self.likelihoods = list(likelihoods)
        for l in self.likelihoods:
            if not isinstance(l, AbstractLikelihood) or l.n_latent != 1:
                raise TypeError(f"One (or more) of the likelihoods {likelihoods} are not compatible or implemented with the {inference}")
        self.likelihood = self.likelihoods
        kernels = [kernel] if isinstance(kernel, Kernel) else list(kernel)
        self.Zs = [np.ascontiguousarray(np.asarray(z, dtype=np.float64)) for z in Zs]
        self.n_latent = len(self.Zs)
        self.n_task = len(self.likelihoods)
        if not isinstance(kernel, Kernel) and len(kernels) != self.n_task:
            raise ValueError("Number of kernels should be equal to the number of tasks")
        self.kernels = [kernels[i % len(kernels)] for i in range(self.n_latent)]
        self.m, self.D = self.Zs[0].shape
        if any(z.shape != (self.m, self.D) for z in self.Zs):
            raise ValueError("all latent GPs must share the same number of inducing points and input dimension")
        if A is None:
            rng = rng or np.random.default_rng()
            A = rng.standard_normal((self.n_task, self.n_latent))
            A /= np.linalg.norm(A, axis=1, keepdims=True)
        self.A = np.array(A, dtype=np.float64, order="C")   # own copy: update_A! refreshes it in place
        if self.A.shape != (self.n_task, self.n_latent):
            raise ValueError("A must be (n_task, n_latent)")
        self.mu0 = None

    def _desc(self, capacity: int) -> dict:
        return _make_desc(self, capacity)

    def __repr__(self):
        return f"Multioutput Sparse Variational Gaussian Process with the likelihoods {self.likelihoods} infered by {self.inference} "


class VGP(AbstractGPModel):
    """models/VGP.jl:22-75.  `VGP(X, y, kernel, likelihood, inference; optimiser=false)`: the full variational GP, trained with
    `train(model, iterations)` on its own data.  On the device it is the SVGP algebra with Z = X, κ = I, K̃ = 0
    (natural_gradient!(::VarLatent), analyticVI.jl:126-140); O(n³), so meant for n up to a few thousand."""

    model_kind = L.MODEL_VGP

    def __init__(self, X, y, kernel: Kernel, likelihood: AbstractLikelihood, inference: AnalyticVI, *, verbose: int = 0, optimiser=False,
                 atfrequency: int = 1, mean=None, obsdim: int = 1, T=np.float64, precision: str = "auto", device: int = 0, stream=None):
        if not isinstance(likelihood, AbstractLikelihood):
            raise TypeError(f"The {likelihood} is not compatible or implemented with the {inference}")
        self._common_init(inference, verbose, atfrequency, optimiser, False, T, precision, device, stream, None)
        if self.optimiser is not None:
            # update_hyperparameters! of full models (autotuning.jl:48-84) differentiates an ELBO whose kernel dependence is the
            # GaussianKL only (mean_f = mu, var_f = diag(Sigma)); agp_hyper_grads implements the sparse ELBO gradient
            raise NotImplementedError("hyper-parameter optimisation of full (non-sparse) models is outside the accelerated path")
        if inference.stoch:
            raise ValueError("VGP is a full-batch model: use AnalyticVI()")
        X = np.asarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X[:, None]
        if X.ndim == 2 and obsdim == 2:
            X = X.T
        self.X = np.ascontiguousarray(X)
        self.y = y
        if len(y) != len(self.X):
            raise ValueError(f"There is not the same number of samples in X ({len(self.X)}) and y ({len(y)})")
        self.likelihood = likelihood
        self.likelihoods = [likelihood]
        self.kernel = kernel
        self.Z = self.X
        self.m, self.D = self.X.shape
        self.n_latent = likelihood.n_latent
        self.kernels = [kernel] * self.n_latent
        self.Zs = [self.X] * self.n_latent
        if mean is None:
            self.mu0 = None
        elif np.isscalar(mean):
            self.mu0 = np.full(self.m, float(mean))
        else:
            raise NotImplementedError("only ZeroMean / ConstantMean priors cross the boundary (mu0 evaluated at X)")
        self.A = None

    def _desc(self, capacity: int) -> dict:
        return _make_desc(self, capacity)

    def __repr__(self):
        return f"Variational Gaussian Process with a {self.likelihood} infered by {self.inference} "


class MOVGP(AbstractGPModel):
    """models/MOVGP.jl:22-120.  `MOVGP(X, ys, kernel, likelihoods, inference, num_latent; Aoptimiser)`: multi-output full GP, the
    MOSVGP algebra with Z = X, κ = I, K̃ = 0 for every latent (update_parameters!(::MOVGP), training/training.jl:146-151)."""

    model_kind = L.MODEL_MOVGP

    def __init__(self, X, ys, kernel, likelihoods: Sequence[AbstractLikelihood], inference: AnalyticVI, num_latent: int, *, A=None,
                 verbose: int = 0, atfrequency: int = 1, optimiser=False, Aoptimiser=False, obsdim: int = 1, T=np.float64,
                 precision: str = "auto", device: int = 0, stream=None, rng=None):
        self._common_init(inference, verbose, atfrequency, optimiser, False, T, precision, device, stream, None)
        if self.optimiser is not None:
            raise NotImplementedError("hyper-parameter optimisation of full (non-sparse) models is outside the accelerated path")
        if inference.stoch:
            raise ValueError("MOVGP is a full-batch model: use AnalyticVI()")
        if isinstance(Aoptimiser, bool) or Aoptimiser is None:
            Aoptimiser = ADAM(0.01) if Aoptimiser else None
        if Aoptimiser is not None and not isinstance(Aoptimiser, ADAM):
            raise NotImplementedError("only ADAM (the reference default) is implemented for the mixing-matrix optimiser")
        self.A_opt = Aoptimiser
        X = np.asarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X[:, None]
        if X.ndim == 2 and obsdim == 2:
            X = X.T
        self.X = np.ascontiguousarray(X)
        self.y = list(ys)
        self.likelihoods = list(likelihoods)
        for l in self.likelihoods:
            if not isinstance(l, AbstractLikelihood) or l.n_latent != 1:
                raise TypeError(f"One (or more) of the likelihoods {likelihoods} are not compatible or implemented with the {inference}")
        if len(self.y) != len(self.likelihoods) or any(len(yt) != len(self.X) for yt in self.y):
            raise ValueError("one target vector of length n per task is required")
        self.likelihood = self.likelihoods
        kernels = [kernel] if isinstance(kernel, Kernel) else list(kernel)
        self.n_latent, self.n_task = int(num_latent), len(self.likelihoods)
        self.kernels = [kernels[i % len(kernels)] for i in range(self.n_latent)]
        self.Zs = [self.X] * self.n_latent
        self.m, self.D = self.X.shape
        if A is None:
            rng = rng or np.random.default_rng()
            A = rng.standard_normal((self.n_task, self.n_latent))
            A /= np.linalg.norm(A, axis=1, keepdims=True)
        self.A = np.array(A, dtype=np.float64, order="C")
        if self.A.shape != (self.n_task, self.n_latent):
            raise ValueError("A must be (n_task, n_latent)")
        self.mu0 = None

    def _desc(self, capacity: int) -> dict:
        return _make_desc(self, capacity)

    def __repr__(self):
        return f"Multioutput Variational Gaussian Process with the likelihoods {self.likelihoods} infered by {self.inference} "


class OnlineSVGP(SVGP):
    """models/OnlineSVGP.jl:1-78: streaming sparse variational GP, AnalyticVI only, `optimiser = nothing` semantics.

    `OnlineSVGP(kernel, likelihood, inference)`: the model starts without inducing points; every `train_online` call brings one
    batch (X, y) and the inducing set to use for it.  The reference chooses that set with the un-vendored InducingPoints.jl
    (`inducingpoints(Zalg, x)`, `updateZ`, `remove_point(Random.GLOBAL_RNG, ...)`, training/onlinetraining.jl:157,175,193) from
    Julia's global RNG, so -- like the minibatch indices of `train` -- it is an argument here instead of a `Zalg`.
    A new inducing set is a new device model of that size; the previous one contributes through agp_online_carry."""

    def __init__(self, kernel: Kernel, likelihood: AbstractLikelihood, inference: AnalyticVI, *, verbose: int = 0, mean=None,
                 T=np.float64, precision: str = "auto", device: int = 0, stream=None):
        if not isinstance(inference, AnalyticVI) or inference.stoch:
            raise ValueError("The inference object should be of type `AnalyticVI`")      # OnlineSVGP.jl:46
        if not isinstance(likelihood, AbstractLikelihood):
            raise TypeError(f"The {likelihood} is not compatible or implemented with the {inference}")
        if likelihood.kind in (L.LIK_LOGISTICSOFTMAX, L.LIK_POISSON, L.LIK_HETEROSCEDASTIC) or getattr(likelihood, "opt_noise", None) is not None:
            raise NotImplementedError("OnlineSVGP is accelerated for likelihoods whose local updates carry no state between batches")
        self._common_init(inference, verbose, 1, False, False, T, precision, device, stream, None)
        self.likelihood = likelihood
        self.likelihoods = [likelihood]
        self.kernel = kernel
        self.n_latent = likelihood.n_latent
        self.kernels = [kernel] * self.n_latent
        if mean is not None and not np.isscalar(mean):
            raise NotImplementedError("only ZeroMean / ConstantMean priors cross the boundary")
        self._mean = None if mean is None else float(mean)
        self.mu0 = None
        self.A = None
        self.Z = None
        self.Zs = None
        self.m, self.D = 0, 0

    def __repr__(self):
        return f"Online Variational Gaussian Process with a {self.likelihood} infered by {self.inference}"


def train_online(model: OnlineSVGP, X, y, Z, state: Optional["State"] = None, iterations: int = 20):
    """`train!(m::OnlineSVGP, X, y, state; iterations)` (training/onlinetraining.jl:36-144) for one batch, with the inducing set `Z`
    to use for it (see OnlineSVGP).  First iteration: the local updates run under the PREVIOUS inducing set and posterior, the
    natural gradient under the new one (onlinetraining.jl:75-104); later iterations are ordinary AnalyticVI steps whose natural
    gradient carries the two previous-set terms (analyticVI.jl:183-203)."""
    if iterations <= 0:
        raise ValueError("Number of iterations should be positive")
    X = np.asarray(X)
    if X.ndim == 1:
        X = X[:, None]
    Z = np.ascontiguousarray(np.asarray(Z, dtype=np.float64))
    if Z.ndim == 1:
        Z = Z[:, None]
    n = X.shape[0]
    inf = model.inference
    inf.batchsize = n
    inf.rho = 1.0
    ys = _wrap_y(model, y)
    full = np.arange(n, dtype=np.int64)
    ip = full.ctypes.data_as(L.c_int64_p)
    Q = model.n_latent
    old = model._eng
    carry, grads = None, None
    if inf.n_iter > 0:
        if old is None or state is None:
            raise ValueError("a trained OnlineSVGP needs the state returned by the previous train_online call")
        if n > old.capacity:
            raise ValueError("the batch is larger than the previous one (the reference's in-place local updates need equal sizes)")
        # save_old_parameters! (onlinetraining.jl:164-183): invD_a = -2 eta2 - K^-1, prev_eta1, prev_L per latent
        ma = model.m
        carry = []
        for q in range(Q):
            mu, S, e1, e2 = model._get_posterior_raw(q)
            Kinv, ldK = np.empty((ma, ma)), C.c_double()
            old.ck(old.lib.agp_get_Kinv(old.model, q, L.dptr(Kinv), C.byref(ldK)))
            invD = np.triu(-2.0 * e2 - Kinv)
            invD = invD + np.triu(invD, 1).T                                  # Symmetric(...) reads the upper triangle
            prevL = float((-np.linalg.slogdet(S)[1] + ldK.value - mu @ e1) / 2.0)
            carry.append((np.ascontiguousarray(model.Zs[q]), np.ascontiguousarray(invD), np.ascontiguousarray(e1), prevL))
        # local updates of the previous model on the new batch (onlinetraining.jl:79-92)
        model._data_key = None
        _upload(model, old, X, ys, ("online-prev", inf.n_iter))
        old.ck(old.lib.agp_step_moments_async(old.model, ip, n, 0))
        old.ck(old.lib.agp_local_updates_async(old.model))
        gmu, gS = np.empty((Q, n)), np.empty((Q, n))
        for q in range(Q):
            old.ck(old.lib.agp_get_local(old.model, b"grad_mu", q, L.dptr(gmu[q]), n))
            old.ck(old.lib.agp_get_local(old.model, b"grad_Sigma", q, L.dptr(gS[q]), n))
        grads = (np.ascontiguousarray(gmu), np.ascontiguousarray(gS))
        old.close()
        model._eng = None
    # updateZs! / init_online_model: the device model of the new inducing set (fresh posterior: posterior.jl:47-55)
    model.Z = Z
    model.m, model.D = Z.shape
    model.Zs = [Z] * Q
    model.mu0 = None if model._mean is None else np.full(model.m, model._mean)
    model.precision = model.precision_requested
    model._data_key = None
    eng = model._engine(n)
    lib = eng.lib
    _upload(model, eng, X, ys, ("online", inf.n_iter))
    model._data_refs = (X, ys)
    eng.ck(lib.agp_state_reset(eng.model))
    eng.ck(lib.agp_refresh_K(eng.model))
    for q in range(Q):
        if carry is None:
            eng.ck(lib.agp_online_carry(eng.model, q, None, 0, None, None, 0.0))
        else:
            Za, invD, e1, prevL = carry[q]
            eng.ck(lib.agp_online_carry(eng.model, q, L.dptr(Za), Za.shape[0], L.dptr(invD), L.dptr(e1), prevL))
    state = State(model)
    state.B = n
    for local_iter in range(1, iterations + 1):
        if local_iter == 1 and grads is not None:
            eng.ck(lib.agp_step_with_gradients(eng.model, ip, n, 0, L.dptr(grads[0]), L.dptr(grads[1])))
            eng.ck(lib.agp_sync(eng.model))
        else:
            eng.ck(lib.agp_step(eng.model, ip, n, 0, 1.0))
        model.trained = True
        inf.n_iter += 1
    return model, state


def _make_desc(model, capacity: int) -> dict:
    q0, ql = model._latent_range()
    inf = model.inference
    opt = inf.optimiser if inf.stoch else None
    liks = model.likelihoods
    Z = np.ascontiguousarray(np.stack(model.Zs[q0 : q0 + ql]))
    d = dict(
        model_kind=model.model_kind,
        n_latent_global=model.n_latent,
        latent_begin=q0,
        n_latent_local=ql,
        m=model.m,
        D=model.D,
        batch_capacity=int(capacity),
        precision=L.PRECISIONS[model.precision],
        stochastic=1 if inf.stoch else 0,
        rm_kappa=opt.kappa if opt else 0.51,
        rm_tau=opt.tau if opt else 1.0,
        jitter=model.jitter,
        n_task=len(liks),
        lik_kind=np.array([l.kind for l in liks], dtype=np.int32),
        lik_p0=np.array([l.p0 for l in liks], dtype=np.float64),
        lik_p1=np.array([l.p1 for l in liks], dtype=np.float64),
        kernel_kind=np.array([k.kind for k in model.kernels[q0 : q0 + ql]], dtype=np.int32),
        kernel_scale=np.array([k.scale for k in model.kernels[q0 : q0 + ql]], dtype=np.float64),
        kernel_variance=np.array([k.variance for k in model.kernels[q0 : q0 + ql]], dtype=np.float64),
        Z=Z,
    )
    if model.A is not None:
        d["A"] = model.A
    if model.mu0 is not None:
        d["mu0"] = np.ascontiguousarray(np.tile(model.mu0, (ql, 1)))
    return d


# --------------------------------------------------------------------------------------------------
# state (training/states.jl): the device holds it; this object exposes it
# --------------------------------------------------------------------------------------------------
class State:
    """Handle on the device-resident training state of the last step (local_vars, opt_state,
    kernel_matrices of the reference's `state` NamedTuple)."""

    def __init__(self, model):
        self.model = model
        self.B = 0

    def local(self, name: str, row: int = 0) -> np.ndarray:
        e = self.model._eng
        out = np.empty(self.B)
        e.ck(e.lib.agp_get_local(e.model, name.encode(), row, L.dptr(out), self.B))
        return out

    def kernel_matrices(self, q: int = 0):
        e = self.model._eng
        m = self.model.m
        Knm, kappa = np.empty((self.B, m)), np.empty((self.B, m))
        e.ck(e.lib.agp_get_kernel_matrices(e.model, q, L.dptr(Knm), L.dptr(kappa), self.B))
        return dict(Knm=Knm, kappa=kappa, Ktilde=self.local("Ktilde", q))

    @property
    def opt_state(self):
        t, _ = self.model.counters()
        return dict(state_eta1=t, state_eta2=t)


# --------------------------------------------------------------------------------------------------
# train!  (training/training.jl:13-111)
# --------------------------------------------------------------------------------------------------
# precision="auto": the largest error amplification sqrt(variance ||K_mm^-1||_inf) (AbstractGPModel.amplification) each path keeps; a fresh
# train() call moves a model above it to the next path.  Calibrated with tools/shape_sweep.py on a B200 (profiles/r2/shape_sweep/): relative
# error of mu / Sigma / ELBO / predictions against the fp64 oracle <= 2e-5 x amplification on the 3xTF32 tensor-core path (the tensor
# cores' truncating fp32 accumulation over long sums of cancelling terms), <= 2e-6 x on the fp32 CUDA-core path; targets 5e-4 / 2e-4 as
# in the parity tests.
AMPLIFICATION_LIMIT = (("tf32x3", 30.0), ("f32", 100.0), ("f64", float("inf")))


def precision_for_amplification(amp: float, current: str) -> str:
    """The path precision="auto" keeps for an error amplification `amp`: the first entry of AMPLIFICATION_LIMIT that admits it, never a
    faster path than the one the shapes selected (`current`: a model below 128 inducing points stays off the tensor cores)."""
    order = [p for p, _ in AMPLIFICATION_LIMIT]
    for p, lim in AMPLIFICATION_LIMIT:
        if amp <= lim and order.index(p) >= order.index(current):
            return p
    return order[-1]


def _wrap_y(model, y):
    if isinstance(model, (MOSVGP, MOVGP)):
        if len(y) != model.n_task:
            raise ValueError("one target vector per task is required")
        return [treat_labels(yt, l) for yt, l in zip(y, model.likelihoods)]
    return [treat_labels(y, model.likelihood)]


def _upload(model, eng, X, ys, key):
    if model._data_key == key:
        return
    Xk, xp, dt, layout, n, D = _x_args(X)
    if D != model.D:
        raise ValueError("input dimension of X does not match the inducing points")
    for yt in ys:
        if len(yt) != n:
            raise ValueError(f"There is not the same number of samples in X ({n}) and y ({len(yt)})")
    is_class = ys[0].dtype == np.int32
    arr = (C.c_void_p * len(ys))(*[yt.ctypes.data_as(C.c_void_p) for yt in ys])
    eng.ck(eng.lib.agp_data_upload(eng.model, xp, dt, layout, n, arr, L.Y_CLASS if is_class else L.Y_REAL))
    model._data_key = key
    model._n = n


def train(model: AbstractGPModel, X=None, y=None, iterations: int = 100, *, callback=None, convergence=None, state: Optional[State] = None,
          obsdim: int = 1, minibatches: Optional[Sequence[np.ndarray]] = None, rng=None, check_every: int = 1,
          refresh_K_after_hyper: bool = False, reupload: bool = False):
    """`train!(model, X, y, iterations; callback, state)`.

    refresh_K_after_hyper: False (default) reproduces the reference (quirk Q3): after `update_hyperparameters!` the K_mm
    factor of this call's first iteration stays in use until the next `train!` call without a state (autotuning.jl:45 is
    commented out; training.jl:187-208 clears the flag), while K_nm follows the new kernel / Z.  True = refactorise K_mm after
    every hyper-parameter update (the conscious fix).
    reupload: the resident copy of (X, y) is keyed on the identity of the arrays; pass True after refilling them in place.

    minibatches: optional list of 0-based index arrays, one per iteration (the reference draws them with
    StatsBase.sample on Julia's global RNG, training.jl:51-53, which cannot be reproduced; parity runs
    inject the lists).  check_every: read the device status back every k iterations (1 = after every
    step, like the reference's immediate error).
    """
    if isinstance(model, (VGP, MOVGP)):   # train!(model::VGP, iterations): the model carries its data (models/VGP.jl, MOVGP.jl)
        if isinstance(X, (int, np.integer)) and y is None:
            iterations, X = int(X), None
        if X is None:
            X, y = model.X, model.y
    elif X is None or y is None:
        raise TypeError("train(model, X, y, iterations) needs the data for sparse models")
    if iterations <= 0:
        raise ValueError("Number of iterations should be positive")
    X = np.asarray(X)
    if X.ndim == 2 and obsdim == 2:
        X = X.T
    ys = _wrap_y(model, y)
    n = X.shape[0]
    inf = model.inference
    if inf.stoch:
        if not (0 < inf.batchsize <= n):
            raise ValueError(
                f"The size of mini-batch {inf.batchsize} is incorrect (negative or bigger than number of samples), "
                "please set `batchsize` correctly in the inference object"
            )
        inf.rho = n / inf.batchsize
    else:
        inf.batchsize = n
    B = inf.batchsize
    eng = model._engine(B)
    if reupload:
        model._data_key = None
    _upload(model, eng, X, ys, (id(X), tuple(id(v) for v in ys), X.shape))
    model._data_refs = (X, ys)   # keep the keyed arrays alive: a freed array's id() can be reused by a new one
    lib = eng.lib
    eng.ck(lib.agp_keep_stale_K(eng.model, 0 if refresh_K_after_hyper else 1))
    fresh = state is None
    if state is None:
        inf.HyperParametersUpdated = True
        eng.ck(lib.agp_state_reset(eng.model))
        state = State(model)
    if inf.HyperParametersUpdated:
        eng.ck(lib.agp_refresh_K(eng.model))  # compute_K, once per train! call (training.jl:41-43, Q3)
        inf.HyperParametersUpdated = False
        # (full models form V = chol(K) directly - no K_nm L^-T product, nothing to amplify)
        if fresh and model.world == 1 and model.precision_requested == "auto" and model.precision != "f64" \
                and not isinstance(model, (VGP, MOVGP)) and os.environ.get("AGP_COND_SWITCH", "1") != "0":
            amp = model.amplification()
            want = precision_for_amplification(amp, model.precision)
            if want != model.precision:
                # precision="auto" promises the reference's (fp64) results to the tolerances of the parity tests: the fp32-class paths lose
                # ~log10(amp) digits in V = K_nm L^-T, the tensor-core accumulation more than the CUDA-core one
                warnings.warn(f"K_mm is ill conditioned (error amplification {amp:.1e}): precision='auto' moves this model from the "
                              f"{model.precision} to the {want} path; pass precision='{model.precision}' to keep the faster one",
                              RuntimeWarning, stacklevel=2)
                model._precision_by_condition = want
                model._eng.close()
                model._eng = None
                eng = model._engine(B)
                lib = eng.lib
                _upload(model, eng, X, ys, (id(X), tuple(id(v) for v in ys), X.shape))
                eng.ck(lib.agp_keep_stale_K(eng.model, 0 if refresh_K_after_hyper else 1))
                eng.ck(lib.agp_state_reset(eng.model))
                state = State(model)
                eng.ck(lib.agp_refresh_K(eng.model))
    state.B = B
    rng = rng or np.random.default_rng()
    full = np.arange(n, dtype=np.int64) if not inf.stoch else None
    for it in range(iterations):
        if inf.stoch:
            idx = minibatches[it] if minibatches is not None else rng.choice(n, B, replace=False)
            idx = np.ascontiguousarray(idx, dtype=np.int64)
            if idx.shape != (B,):
                raise ValueError("minibatch index list has the wrong length")
        else:
            idx = full
        ip = idx.ctypes.data_as(L.c_int64_p)
        if model.world > 1:
            _sharded_step(model, eng, ip, B, inf.rho)
            if (it + 1) % check_every == 0 or it == iterations - 1:
                eng.ck(lib.agp_sync(eng.model))
        elif (it + 1) % check_every == 0 or it == iterations - 1:
            eng.ck(lib.agp_step(eng.model, ip, B, 0, inf.rho))
        else:
            eng.ck(lib.agp_step_async(eng.model, ip, B, 0, inf.rho))
        model.trained = True
        if callback is not None:
            callback(model, state, inf.n_iter)
        # training/training.jl:65-69 : hyper-parameters every `atfrequency` iterations, from the 4th on, never on the last one
        if (model.optimiser is not None or model.Zoptimiser is not None) and inf.n_iter % model.atfrequency == 0 and inf.n_iter >= 3 \
                and it != iterations - 1:
            update_hyperparameters(model, eng)
            if refresh_K_after_hyper:
                eng.ck(lib.agp_refresh_K(eng.model))   # the fix: compute_kernel_matrices as if HPupdated had been raised
        inf.n_iter += 1
    _refresh_lik_params(model, eng)
    return model, state


def hyper_grads(model, eng=None):
    """ELBO gradients w.r.t. kernel scale / variance / inducing points of the owned latents on the last minibatch (agp_hyper_grads).
    Returns a list of dict(scale=, variance=, Z=(m, D))."""
    eng = eng or model._eng
    ql = model.n_latent_local
    ds, dv, dZ = np.zeros(ql), np.zeros(ql), np.zeros((ql, model.m, model.D))
    eng.ck(eng.lib.agp_hyper_grads(eng.model, float(model.inference.rho), L.dptr(ds), L.dptr(dv), L.dptr(dZ)))
    return [dict(scale=float(ds[q]), variance=float(dv[q]), Z=dZ[q]) for q in range(ql)]


def update_hyperparameters(model, eng):
    """update_hyperparameters! for sparse models (hyperparameter/autotuning.jl:86-140): gradients from the device, ADAM on the log
    of the positive kernel parameters (update_kernel!, autotuning_utils.jl:63-67) and on the inducing points (update_Z!, :78-82)
    on the host, new values pushed with agp_set_kernel / agp_set_Z."""
    grads = hyper_grads(model, eng)
    q0, ql = model._latent_range()
    if model._hyperopt_state is None:
        model._hyperopt_state = [None] * ql
    for q, g in enumerate(grads):
        k = model.kernels[q0 + q]
        st = model._hyperopt_state[q]
        if st is None:
            st = model._hyperopt_state[q] = dict(
                scale=model.optimiser.init(np.zeros(1)) if model.optimiser else None,
                variance=model.optimiser.init(np.zeros(1)) if model.optimiser else None,
                Z=model.Zoptimiser.init(np.zeros((model.m, model.D))) if model.Zoptimiser else None)
        new_scale, new_var = k.scale, k.variance
        if model.optimiser is not None:
            st["variance"], d = model.optimiser.apply(st["variance"], np.array([k.variance * g["variance"]]))
            new_var = float(np.exp(np.log(k.variance) + d[0]))
            st["scale"], d = model.optimiser.apply(st["scale"], np.array([k.scale * g["scale"]]))
            new_scale = float(np.exp(np.log(k.scale) + d[0]))
            model.kernels = list(model.kernels)
            model.kernels[q0 + q] = Kernel(k.kind, new_scale, new_var)
            eng.ck(eng.lib.agp_set_kernel(eng.model, q, k.kind, new_scale, new_var))
        if model.Zoptimiser is not None:
            st["Z"], dZ = model.Zoptimiser.apply(st["Z"], g["Z"])
            model.Zs = list(model.Zs)
            model.Zs[q0 + q] = np.ascontiguousarray(model.Zs[q0 + q] + dZ)
            eng.ck(eng.lib.agp_set_Z(eng.model, q, L.dptr(model.Zs[q0 + q])))
    if isinstance(model, SVGP):
        model.kernel, model.Z = model.kernels[0], model.Zs[0]


def _refresh_lik_params(model, eng):
    """read the re-estimated link parameters (λ of Poisson / Heteroscedastic) back into the likelihood objects, and the
    mixing matrix when update_A! is on"""
    if getattr(model, "A_opt", None) is not None:
        eng.ck(eng.lib.agp_get_A(eng.model, L.dptr(model.A)))
    for t, l in enumerate(model.likelihoods):
        if l.kind in (L.LIK_POISSON, L.LIK_HETEROSCEDASTIC):
            v = C.c_double(0.0)
            eng.ck(eng.lib.agp_get_lik_param(eng.model, t, C.byref(v)))
            l.lam = v.value
        if l.kind == L.LIK_GAUSSIAN and getattr(l, "opt_noise", None) is not None:
            v = C.c_double(0.0)
            eng.ck(eng.lib.agp_get_lik_param(eng.model, t, C.byref(v)))
            l.sigma2 = v.value


train_ = train  # `train!`


def _moment_views(model, eng):
    """torch views (zero-copy) on the [Q][ldB] device moment arrays, for the NCCL all-gather."""
    import torch

    if getattr(model, "_mviews", None) is not None and model._mviews[0] is eng:
        return model._mviews[1]
    views = []
    for which in (0, 1):
        ld = C.c_int64()
        p = eng.lib.agp_moments_devptr(eng.model, which, C.byref(ld))

        class _Arr:
            __cuda_array_interface__ = dict(shape=(model.n_latent * ld.value,), typestr="<f8", data=(int(p), False), version=2)

        views.append(torch.as_tensor(_Arr(), device=f"cuda:{model.device}"))
    model._mviews = (eng, (views, ld.value))
    return model._mviews[1]


def _allgather_rows(views, q0: int, ql: int, ld: int):
    """In-place all-gather of the owned rows [q0, q0+ql) of flat [Q*ld] arrays (every rank owns ql rows)."""
    import torch.distributed as dist

    for v in views:
        dist.all_gather_into_tensor(v, v[q0 * ld : (q0 + ql) * ld])


def _allgather_moments(model, eng):
    views, ld = _moment_views(model, eng)
    q0, ql = model._latent_range()
    _allgather_rows(views, q0, ql, ld)


def _attach_peers(model, eng) -> bool:
    """Exchange the CUDA IPC handles of every rank's moment block over the host process group and attach them
    (agp_peer_export / agp_peer_attach): afterwards the per-step exchange of (mean_f, var_f) rows runs device-side
    over NVLink peer memory inside the step.  Returns False (NCCL all-gather fallback) when no process group exists."""
    try:
        import torch.distributed as dist
    except Exception:
        return False
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() != model.world:
        return False
    h = C.create_string_buffer(64)
    ok = eng.lib.agp_peer_export(eng.model, h) == L.AGP_OK
    hs = [None] * model.world
    dist.all_gather_object(hs, bytes(h.raw) if ok else b"")
    attached = False
    if all(len(x) == 64 for x in hs):
        blob = C.create_string_buffer(b"".join(hs), 64 * model.world)
        attached = eng.lib.agp_peer_attach(eng.model, model.world, model.rank, blob) == L.AGP_OK
    # every rank must end up in the same mode: one failure (IPC not permitted, no P2P path) sends everybody to the NCCL fallback
    votes = [None] * model.world
    dist.all_gather_object(votes, bool(attached))
    if not all(votes):
        if attached:
            eng.ck(eng.lib.agp_peer_detach(eng.model))
        return False
    return True


def _sharded_step(model, eng, ip, B, rho):
    """one-latent(-group)-per-rank step.  Peer mode: one agp_step_async, the moment rows cross NVLink inside the step.
    Fallback: moments -> NCCL all-gather of (mean_f, var_f) rows -> update."""
    if getattr(model, "_peer", False):
        eng.ck(eng.lib.agp_step_async(eng.model, ip, B, 0, rho))
        return
    eng.ck(eng.lib.agp_step_moments_async(eng.model, ip, B, 0))
    _allgather_moments(model, eng)
    eng.ck(eng.lib.agp_step_update_async(eng.model, rho))


# --------------------------------------------------------------------------------------------------
# ELBO / objective (inference/analyticVI.jl:255-297)
# --------------------------------------------------------------------------------------------------
def ELBO(model: AbstractGPModel, state: Optional[State] = None, y=None) -> float:
    """ELBO(model, state, y) on the last minibatch (y is implied by the state: the labels of that batch)."""
    eng = model._eng
    if eng is None:
        raise RuntimeError("the model has not been trained yet")
    out = np.zeros(3)
    rho = model.inference.rho
    if model.world > 1:
        import torch
        import torch.distributed as dist

        eng.ck(eng.lib.agp_elbo_moments_async(eng.model))
        if not getattr(model, "_peer", False):
            _allgather_moments(model, eng)
        eng.ck(eng.lib.agp_elbo(eng.model, rho, L.dptr(out)))
        kl = torch.tensor([out[1]], dtype=torch.float64, device=f"cuda:{model.device}")
        dist.all_reduce(kl)  # the single scalar all-reduce of the north star
        out[1] = float(kl.item())
    else:
        eng.ck(eng.lib.agp_elbo(eng.model, rho, L.dptr(out)))
    extra = 0.0
    if isinstance(model, OnlineSVGP):   # extraKL (functions/KLdivergences.jl:37-54)
        ek = C.c_double()
        eng.ck(eng.lib.agp_online_extra_kl(eng.model, C.byref(ek)))
        extra = ek.value
    return float(out[0] - out[1] - out[2] - extra)


objective = ELBO


# --------------------------------------------------------------------------------------------------
# predictions (training/predictions.jl)
# --------------------------------------------------------------------------------------------------
_GH = None


def _pred_nodes():
    """predictions.jl:4 : 100-point Gauss-Hermite nodes * sqrt2, weights / sqrt(pi)"""
    global _GH
    if _GH is None:
        x, w = np.polynomial.hermite.hermgauss(100)
        _GH = (np.ascontiguousarray(x * math.sqrt(2.0)), np.ascontiguousarray(w / math.sqrt(math.pi)))
    return _GH


def _predict_f(model, X_test, cov: bool):
    """(Q_local, N*) latent moments of the owned latents, then multi-output mixing on the host."""
    eng = model._engine(max(model.inference.batchsize, 1) if model._eng is None else 0)
    Xk, xp, dt, layout, nt, D = _x_args(X_test)
    if D != model.D:
        raise ValueError("input dimension of X_test does not match the inducing points")
    ql = model.n_latent_local
    mu = np.empty((ql, nt))
    var = np.empty((ql, nt)) if cov else None
    eng.ck(eng.lib.agp_predict_f(eng.model, xp, dt, layout, nt, 1 if cov else 0, L.dptr(mu), L.dptr(var) if cov else None))
    if model.world > 1:   # latent-sharded model: every rank predicts its own latents, the rows are gathered on the host (collective call)
        import torch.distributed as dist

        parts = [None] * model.world
        dist.all_gather_object(parts, (mu, var))
        mu = np.concatenate([p[0] for p in parts], axis=0)
        var = np.concatenate([p[1] for p in parts], axis=0) if cov else None
    if isinstance(model, (MOSVGP, MOVGP)):  # predictions.jl:63-84
        mu_t = model.A @ mu
        return (mu_t, (model.A**2) @ var) if cov else (mu_t, None)
    return mu, var


def predict_f(model, X_test, state=None, *, cov: bool = False, diag: bool = True, obsdim: int = 1):
    """predictions.jl:136-163"""
    X_test = np.asarray(X_test)
    if X_test.ndim == 2 and obsdim == 2:
        X_test = X_test.T
    if cov and not diag:    # predictions.jl:45-49: full covariance (fp64 on the device, nt x nt per latent)
        if model.world > 1:
            raise NotImplementedError("full predictive covariance of a latent-sharded model")
        Xd = np.ascontiguousarray(X_test if X_test.ndim == 2 else X_test[:, None], dtype=np.float64)
        nt = Xd.shape[0]
        eng = model._engine(max(model.inference.batchsize, 1) if model._eng is None else 0)
        ql = model.n_latent_local
        mu, S = np.empty((ql, nt)), np.empty((ql, nt, nt))
        eng.ck(eng.lib.agp_predict_f_cov(eng.model, L.dptr(Xd), nt, L.dptr(mu), L.dptr(S)))
        if isinstance(model, (MOSVGP, MOVGP)):  # predictions.jl:63-84
            mu, S = model.A @ mu, np.einsum("tq,qij->tij", model.A**2, S)
        if mu.shape[0] == 1:
            return mu[0], S[0]
        return tuple(mu), tuple(S)
    mu, var = _predict_f(model, X_test, cov)
    if mu.shape[0] == 1:
        return (mu[0], var[0]) if cov else mu[0]
    return (tuple(mu), tuple(var)) if cov else tuple(mu)


def _predict_y_lik(lik, mu):
    if lik.kind in (L.LIK_LOGISTIC, L.LIK_BAYESIANSVM):  # classification.jl:47
        return mu[0] > 0
    if lik.kind == L.LIK_POISSON:  # poisson.jl:39-41 (predict_y = expec_count = λ σ(μ))
        return lik.lam / (1.0 + np.exp(-mu[0]))
    if lik.kind == L.LIK_NEGBINOMIAL:  # negativebinomial.jl (predict_y = r σ(μ) / (1 - σ(μ)) = r e^μ)
        return lik.r * np.exp(mu[0])
    if lik.kind == L.LIK_LOGISTICSOFTMAX:  # predictions.jl:196-198
        am = np.argmax(mu, axis=0)
        return np.array([lik.class_mapping[i] for i in am])
    return mu[0]  # regression.jl:17


def predict_y(model, X_test, state=None, *, obsdim: int = 1):
    """predictions.jl:178-198"""
    X_test = np.asarray(X_test)
    if X_test.ndim == 2 and obsdim == 2:
        X_test = X_test.T
    mu, _ = _predict_f(model, X_test, False)
    if isinstance(model, (MOSVGP, MOVGP)):
        return [_predict_y_lik(l, mu[t : t + 1]) for t, l in enumerate(model.likelihoods)]
    return _predict_y_lik(model.likelihood, mu)


def _compute_proba(model, lik, mu, var):
    if lik.kind == L.LIK_LOGISTIC:  # classification.jl:14-26, Gauss-Hermite on the device
        eng = model._eng
        nodes, weights = _pred_nodes()
        m_ = np.ascontiguousarray(mu[0])
        v_ = np.ascontiguousarray(var[0])
        p, pv = np.empty_like(m_), np.empty_like(m_)
        eng.ck(eng.lib.agp_proba_logistic(eng.model, L.dptr(m_), L.dptr(v_), len(m_), L.dptr(nodes), L.dptr(weights),
                                          len(nodes), L.dptr(p), L.dptr(pv)))
        return p, pv
    if lik.kind in (L.LIK_POISSON, L.LIK_NEGBINOMIAL, L.LIK_BAYESIANSVM):  # poisson.jl:43-55, negativebinomial.jl:47-62
        eng = model._eng
        link, p0 = {L.LIK_POISSON: (1, lik.p0), L.LIK_NEGBINOMIAL: (2, lik.p0), L.LIK_BAYESIANSVM: (3, 0.0)}[lik.kind]
        m_ = np.ascontiguousarray(mu[0])
        v_ = np.ascontiguousarray(var[0])
        p, pv = np.empty_like(m_), np.empty_like(m_)
        eng.ck(eng.lib.agp_proba_link(eng.model, link, float(p0), L.dptr(m_), L.dptr(v_), len(m_), L.dptr(p), L.dptr(pv)))
        return p, pv
    if lik.kind == L.LIK_LAPLACE:  # laplace.jl:48-52
        return mu[0], np.maximum(var[0], 0.0) + 2.0 * lik.beta**2
    if lik.kind == L.LIK_HETEROSCEDASTIC:  # heteroscedastic.jl:64-70 : (μ_f, σ²_f + 1 / (λ σ(μ_g)))
        return mu[0], var[0] + (1.0 + np.exp(-mu[1])) / lik.lam
    if lik.kind == L.LIK_GAUSSIAN:  # gaussian.jl:41-45
        return mu[0], var[0] + lik.sigma2
    if lik.kind == L.LIK_STUDENTT:  # studentt.jl:57-61
        return mu[0], np.maximum(var[0], 0.0) + lik.nu * lik.sigma**2 / (2.0 * (lik.nu / 2.0 - 1.0))
    if lik.kind == L.LIK_LOGISTICSOFTMAX:  # multiclass.jl:96-117, logisticsoftmax.jl:28-30
        s = 1.0 / (1.0 + np.exp(-mu))
        return s / np.sum(s, axis=0, keepdims=True)
    raise ValueError("unknown likelihood")


def proba_y(model, X_test, state=None, *, obsdim: int = 1):
    """predictions.jl:231-246"""
    X_test = np.asarray(X_test)
    if X_test.ndim == 2 and obsdim == 2:
        X_test = X_test.T
    mu, var = _predict_f(model, X_test, True)
    if isinstance(model, (MOSVGP, MOVGP)):
        return [_compute_proba(model, l, mu[t : t + 1], var[t : t + 1]) for t, l in enumerate(model.likelihoods)]
    return _compute_proba(model, model.likelihood, mu, var)
