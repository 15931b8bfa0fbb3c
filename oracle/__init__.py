"""CPU oracle for the Fortuna SG-MCMC / BMA-predictive / scoring hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

It is a numpy restatement (float32 op-by-op, uint32 bit-exact) of the arithmetic
of awslabs/fortuna's hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and
only as the checker or as the timed CPU baseline.  Nothing under ``fortuna_b200/``
imports it; the product path fails loudly when the CUDA library is missing.

Why a restatement: the reference is pure Python on jax 0.4.30 / flax 0.8.5 /
optax 0.1.9 (poetry.lock:2187-2188, 1328-1329, 3984-3985).  None of those are
installed in the build container or on the GPU box, and there are no wheels, so
the reference cannot be run on real jax.  The arithmetic of jax's PRIMITIVES
lives in those un-vendored dependencies (threefry2x32 PRNG, ``lax.erf_inv``,
``jax.nn.softmax``, ``jsp.special.logsumexp``, ``jnp.argsort`` / ``cumsum`` /
``quantile`` / ``linspace``); their published algorithms are restated here and
cited function by function.

Parity status (also in DESIGN.md section 4):
  * pinned against the reference's OWN SOURCE, executed here
            - tests/golden/make_ref_vectors.py imports the reference's files
              for this path from /root/reference (predictive/base.py,
              classification.py, regression.py, likelihood/*, prob_output_layer/*,
              sgmcmc_posterior.py, sghmc_integrator.py, cyclical_sgld_integrator.py,
              sgmcmc_preconditioner.py, sgmcmc_step_schedule.py, sgmcmc_diagnostic.py,
              utils/random.py, utils/training.py, conformal/classification/*,
              conformal/regression/*, conformal/multivalid/*, metric/classification.py,
              loss/*, output_calib_model/predictive/*) and runs them over a
              numpy stand-in for jax (tests/golden/jax_standin.py: float32
              promotions derived by rule, jax.random bound to the KAT-pinned
              threefry).  tests/test_ref_vectors.py: sampler trajectories
              (parameters, momentum, second moments, key chain), predictive
              mean / mode / variances / std / log-probs, stored-sample draws,
              APS and simple scores, conformal and credible sets, ECE bins,
              MaxCov patches, conformal-regression quantiles and intervals, the
              output-calibration predictive maps and target draws are BIT-EQUAL
              to the reference source's results;
              entropies, Brier / ECE values, KSD / ESS, multivalid runs agree to a
              few float32 ulps (a jnp.sum over many terms has no single order).
              This pins everything written in the reference's .py files: control
              flow, formulas, key schedule, dtype promotion, indexing.
  * pinned against external known answers
            - threefry2x32 / split / uniform / normal against the Random123 and
              JAX-documented values; step schedules,
              output-layer log-probs and the simple-score conformal case against
              the reference's own tests.
  * parity unpinned
            - the numerics of XLA primitives that no vector here can reach without
              jax itself: the association order of ``jnp.cumsum`` (restated as
              lax.associative_scan's, oracle/scoring.py) and of ``jnp.sum`` /
              ``mean``, XLA's exp / log / erf_inv polynomials (libm here; the
              tolerances cover them), ``jax.random.choice`` with ``p=`` and the
              sort-based permutation (restated from jax/_src/random.py), the
              flax-msgpack checkpoint bytes, the calibration-model loss heads
              and Adam (checked against float64 derivatives and optax's formulas).
"""

from . import prng, sampler, predictive, scoring  # noqa: F401
