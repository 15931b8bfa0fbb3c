"""Step schedules of the SG-MCMC samplers - same names, arguments and errors as
fortuna/prob_model/posterior/sgmcmc/sgmcmc_step_schedule.py:11-149.

A schedule maps the int32 step count to a float32 step size.  It is scalar host logic evaluated once per
step (numpy float32, same operation order as the jnp expressions) and handed to the sampler kernel as
``step_size``; the kernel never re-derives it.
"""

from typing import Callable

import numpy as np

StepSchedule = Callable[[int], np.float32]
_F = np.float32


def constant_schedule(init_step_size: float) -> StepSchedule:
    if not init_step_size >= 0:
        raise ValueError("`init_step_size` should be >= 0.")

    def schedule(_step):
        return _F(init_step_size)

    return schedule


def cosine_schedule(init_step_size: float, total_steps: int) -> StepSchedule:
    if not init_step_size >= 0:
        raise ValueError("`init_step_size` should be >= 0.")
    if not total_steps > 0:
        raise ValueError("`total_steps` should be > 0.")

    def schedule(step):
        t = _F(np.int32(step)) / _F(total_steps)
        return _F(0.5 * init_step_size) * (_F(1) + np.cos(t * _F(np.pi), dtype=_F))

    return schedule


def polynomial_schedule(a: float = 1.0, b: float = 1.0, gamma: float = 0.55) -> StepSchedule:
    if not 0.5 < gamma <= 1.0:
        raise ValueError("`gamma` should be in (0.5, 1.0] range.")

    def schedule(step):
        return _F(a) * np.power(_F(b) + _F(np.int32(step)), _F(-gamma), dtype=_F)

    return schedule


def constant_schedule_with_cosine_burnin(
    init_step_size: float, final_step_size: float, burnin_steps: int
) -> StepSchedule:
    if not init_step_size >= 0:
        raise ValueError("`init_step_size` should be >= 0.")
    if not final_step_size >= 0:
        raise ValueError("`final_step_size` should be >= 0.")
    if not burnin_steps >= 0:
        raise ValueError("`burnin_steps` should be >= 0.")

    def schedule(step):
        with np.errstate(divide="ignore", invalid="ignore"):
            t = np.minimum(_F(np.int32(step)) / _F(burnin_steps), _F(1.0))
        coef = (_F(1) + np.cos(t * _F(np.pi), dtype=_F)) * _F(0.5)
        return coef * _F(init_step_size) + (_F(1) - coef) * _F(final_step_size)

    return schedule


def cyclical_cosine_schedule_with_const_burnin(
    init_step_size: float, burnin_steps: int, cycle_length: int
) -> StepSchedule:
    if not init_step_size >= 0:
        raise ValueError("`init_step_size` should be >= 0.")
    if not burnin_steps >= 0:
        raise ValueError("`burnin_steps` should be >= 0.")
    if not cycle_length >= 0:
        raise ValueError("`cycle_length` should be >= 0.")

    def schedule(step):
        t = np.maximum(_F(np.int32(step) - np.int32(burnin_steps) - np.int32(1)), _F(0.0))
        t = np.fmod(t, _F(cycle_length)) / _F(cycle_length)
        return _F(0.5 * init_step_size) * (_F(1) + np.cos(t * _F(np.pi), dtype=_F))

    return schedule
