"""Minimal stand-in for `empymod` so the reference package can be imported to
generate golden vectors (the real package is not installed; it is not on the
multigrid hot path). Used only by tests/golden/make_golden.py."""
import numpy as np

__version__ = '0.0.stub'


class EMArray(np.ndarray):
    def __new__(cls, data, dtype=None):
        return np.asarray(data, dtype=dtype).view(cls)

    def amp(self):
        return np.abs(self.view(np.ndarray))

    def pha(self, deg=False, unwrap=True, lag=True):
        return np.angle(self.view(np.ndarray))
