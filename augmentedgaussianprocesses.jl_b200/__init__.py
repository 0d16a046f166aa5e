"""B200-native engine for the AnalyticVI / AnalyticSVI hot path of AugmentedGaussianProcesses.jl.

The directory name contains a dot, so import it through the top-level alias module `agp_b200`
(`import agp_b200 as agp`).  Everything numerical runs in libagp_b200.so (hand-written CUDA for
sm_100a behind the C ABI of include/agp_b200.h); there is no CPU fallback.
"""
from . import _lib
from ._lib import AGPError, KtildeError, PosDefException
from .api import (
    ADAM,
    Descent,
    AnalyticSVI,
    AnalyticVI,
    ELBO,
    BayesianSVM,
    GaussianLikelihood,
    HeteroscedasticLikelihood,
    LaplaceLikelihood,
    NegBinomialLikelihood,
    PoissonLikelihood,
    Kernel,
    LogisticLikelihood,
    LogisticSoftMaxLikelihood,
    Matern32Kernel,
    Matern52Kernel,
    MOSVGP,
    MOVGP,
    OnlineSVGP,
    RobbinsMonro,
    ScaleTransform,
    SqExponentialKernel,
    State,
    StudentTLikelihood,
    SVGP,
    VGP,
    create_mapping,
    hyper_grads,
    is_stochastic,
    objective,
    predict_f,
    predict_y,
    proba_y,
    train,
    train_online,
    transform,
    treat_labels,
    with_lengthscale,
)

__all__ = [n for n in dir() if not n.startswith("_")]
