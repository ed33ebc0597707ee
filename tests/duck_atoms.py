"""Duck-typed stand-in for `ase.atoms.Atoms` (ase is not installed in this image).

Exposes exactly what `ASEParser` touches (kinisi/parser.py:295-356): `get_scaled_positions()`,
`.cell[:]`, `get_volume()` and iteration yielding objects with `.symbol`.
"""
import numpy as np


class _Site:

    def __init__(self, symbol):
        self.symbol = symbol


class DuckAtoms:

    def __init__(self, symbols, scaled_positions, cell):
        self._symbols = list(symbols)
        self._scaled = np.asarray(scaled_positions, dtype=float)
        self.cell = np.asarray(cell, dtype=float)

    def get_scaled_positions(self):
        return self._scaled

    def get_volume(self):
        return float(abs(np.linalg.det(self.cell)))

    def __iter__(self):
        return iter(_Site(s) for s in self._symbols)

    def __len__(self):
        return len(self._symbols)


def trajectory_from_arrays(symbols, frac, latt):
    """frac [N, F, 3], latt [F, 3, 3] (WITHOUT the duplicated first frame) -> list of DuckAtoms."""
    return [DuckAtoms(symbols, frac[:, f], latt[f]) for f in range(frac.shape[1])]
