"""``BatchMVPConformalRegressor`` - fortuna/conformal/regression/batch_mvp.py:5-21."""
from ..multivalid.iterative.batch_mvp import BatchMVPConformalMethod


class BatchMVPConformalRegressor(BatchMVPConformalMethod):
    def __init__(self, seed: int = 0):
        super().__init__(seed=seed)
