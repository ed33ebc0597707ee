"""A sample container with the surface of `uravu.distribution.Distribution` that kinisi's users touch
(kinisi/diffusion.py:18, 392-395, 436, 456, 482; docs notebooks: .samples, .size, .n, .con_int()).
There is no arithmetic of the hot path here."""
import numpy as np


class Distribution:
    """:param samples: Sample values.  :param name: A label.  :param ci_points: percentiles of the interval."""

    def __init__(self, samples, name='Distribution', ci_points=None):
        self.name = name
        self.samples = np.array(samples, dtype=float).flatten()
        self.ci_points = np.array([2.5, 97.5]) if ci_points is None else np.array(ci_points)

    @property
    def size(self):
        return self.samples.size

    @property
    def n(self):
        """Median of the samples."""
        return float(np.percentile(self.samples, 50))

    @property
    def s(self):
        return float(np.std(self.samples, ddof=1)) if self.samples.size > 1 else 0.0

    @property
    def v(self):
        return float(np.var(self.samples, ddof=1)) if self.samples.size > 1 else 0.0

    @property
    def min(self):
        return float(self.samples.min())

    @property
    def max(self):
        return float(self.samples.max())

    def con_int(self):
        return np.percentile(self.samples, self.ci_points)

    ci = con_int

    def __array__(self, dtype=None, copy=None):
        return self.samples if dtype is None else self.samples.astype(dtype)

    def __len__(self):
        return self.samples.size

    def to_dict(self):
        return {'name': self.name, 'samples': self.samples, 'ci_points': self.ci_points}

    @classmethod
    def from_dict(cls, d):
        return cls(d['samples'], d.get('name', 'Distribution'), d.get('ci_points'))

    def __repr__(self):
        lo, hi = self.con_int() if self.size else (np.nan, np.nan)
        return f'Distribution(n={self.n if self.size else np.nan:.6g}, ci=[{lo:.6g}, {hi:.6g}], size={self.size})'
