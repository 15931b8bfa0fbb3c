"""Which SG-MCMC steps are stored, and the store itself.

Mirrors fortuna/prob_model/posterior/sgmcmc/sgmcmc_sampling_callback.py:7-47 (the callback), and the two
store rules of sghmc/sghmc_callback.py:49-53 and cyclical_sgld/cyclical_sgld_callback.py:55-59.  The rules
are host integer logic (plain functions below, shared with the oracle tests); the store is a device row
copy into the sample bank (fb200_bank_put_f32) - no device_get, no msgpack.
"""

from ....training.callback import Callback


def sghmc_stores_step(step: int, burnin_length: int, n_thinning: int) -> bool:
    """1-based `step` is stored once the burn-in is over, every `n_thinning`-th step."""
    past = step - burnin_length
    return past > 0 and past % n_thinning == 0


def cyclical_sgld_stores_step(step: int, cycle_length: int, exploration_ratio: float, n_thinning: int) -> bool:
    """Stored only in the sampling part of a cycle (position / length >= ratio, true division like the
    reference), every `n_thinning`-th position of the cycle."""
    pos = step % cycle_length
    return (pos / cycle_length) >= exploration_ratio and pos % n_thinning == 0


def count_stored(rule, n_steps: int) -> int:
    return sum(1 for step in range(1, n_steps + 1) if rule(step))


class SGMCMCSamplingCallback(Callback):
    """`rule(step)` decides; at most `n_samples` states are put into `state_repository` rows 0..n_samples-1."""

    def __init__(self, trainer, state_repository, keep_top_n_checkpoints: int, rule=None, n_samples=None):
        self._trainer = trainer
        self._state_repository = state_repository
        self._keep_top_n_checkpoints = keep_top_n_checkpoints
        self._rule = rule
        self._n_samples = n_samples
        self._current_step = 0
        self._samples_count = 0

    def _do_sample(self, current_step: int, samples_count: int) -> bool:
        if self._rule is None:
            raise NotImplementedError
        return samples_count < self._n_samples and self._rule(current_step)

    def _require(self, n_total_steps: int, hint: str):
        got = count_stored(self._rule, n_total_steps)
        if got < self._n_samples:
            raise ValueError(
                f"The number of desired samples `n_samples` is {self._n_samples}. However, only {got} samples "
                f"will be collected. Consider adjusting {hint}."
            )

    def training_step_end(self, state):
        self._current_step += 1
        if self._do_sample(self._current_step, self._samples_count):
            synced = self._trainer._sync_state(state) if self._trainer is not None else state
            self._state_repository.put(state=synced, i=self._samples_count, keep=self._keep_top_n_checkpoints)
            self._samples_count += 1
        return state
