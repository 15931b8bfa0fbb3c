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
the reference cannot be imported or run.  The arithmetic itself lives in those
un-vendored dependencies (threefry2x32 PRNG, ``lax.erf_inv``, ``jax.nn.softmax``,
``jsp.special.logsumexp``, ``jnp.argsort`` / ``quantile`` / ``linspace``); their
published algorithms are restated here and cited function by function.

Parity status (also in DESIGN.md):
  * pinned  - threefry2x32 / split / uniform / normal against the Random123 and
              JAX-documented known answers; step
              schedules, output-layer log-probs and the simple-score conformal
              case against the reference's own tests.
  * parity unpinned - SGHMC / cyclical-SGLD trajectories, predictive
              mean/variance/entropy values, APS scores, ECE/MCE, quantile,
              ``randint``/``choice``: the reference's tests only check shapes
              there, so the oracle is guarded by analytic identities and
              float64 re-evaluation instead (tests/test_oracle_*.py).
"""

from . import prng, sampler, predictive, scoring  # noqa: F401
