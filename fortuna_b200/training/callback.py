"""Callback base class - fortuna/training/callback.py:4-68."""


class Callback:
    def training_epoch_start(self, state):
        return state

    def training_epoch_end(self, state):
        return state

    def training_step_end(self, state):
        return state
