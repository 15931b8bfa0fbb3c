"""fortuna_b200: B200 (sm_100a) implementation of awslabs/fortuna's posterior-sampling and
Bayesian-model-averaging hot path behind Fortuna's own module layout.

Only the path of SURVEY.md section 8 lives here (SG-MCMC sampler step, last-layer BMA predictive,
[S,B,C] reductions, conformal / calibration scoring).  All arithmetic runs in hand-written CUDA kernels
reached through the C ABI of include/fortuna_b200.h; there is no CPU fallback.
"""

__version__ = "0.1.0"
