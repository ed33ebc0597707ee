"""Minimal `uravu.distribution.Distribution` stand-in: a container of samples.

Only what kinisi touches: `.samples`, `.size`, `.n` (median), `.s`, `.con_int()`/`.ci()`,
`.to_dict()/.from_dict()`. No normality test, no KDE.
"""
import numpy as np


class Distribution:

    def __init__(self, samples, name='Distribution', ci_points=None):
        self.name = name
        self.samples = np.array(samples, dtype=float).flatten()
        self.ci_points = np.array([2.5, 97.5]) if ci_points is None else np.array(ci_points)

    @property
    def size(self):
        return self.samples.size

    @property
    def n(self):
        return float(np.percentile(self.samples, 50))

    @property
    def s(self):
        return float(np.std(self.samples, ddof=1)) if self.samples.size > 1 else 0.0

    @property
    def v(self):
        return float(np.var(self.samples, ddof=1)) if self.samples.size > 1 else 0.0

    def con_int(self):
        return np.percentile(self.samples, self.ci_points)

    ci = con_int

    def to_dict(self):
        return {'name': self.name, 'samples': self.samples, 'ci_points': self.ci_points}

    @classmethod
    def from_dict(cls, d):
        return cls(d['samples'], d.get('name', 'Distribution'), d.get('ci_points'))
