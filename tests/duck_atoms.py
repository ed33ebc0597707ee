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


# ---- duck-typed stand-in for `MDAnalysis.Universe` (MDAnalysis is not installed in this image) -----------------------
# Exposes what `MDAnalysisParser` touches (kinisi/parser.py:600-672): `.trajectory[::n]` yielding timesteps with `.volume`
# and `.triclinic_dimensions` while moving the universe to that frame, `.atoms[::n]` with `.positions` (Cartesian float32,
# ONE buffer reused for every frame, as MDAnalysis readers do) and iteration yielding objects with `.type`.
class _TypedSite:

    def __init__(self, kind):
        self.type = kind


class _Timestep:

    def __init__(self, cell, dtype):
        self.triclinic_dimensions = np.asarray(cell, dtype=dtype)  # MDAnalysis holds the box in float32
        self.volume = float(abs(np.linalg.det(self.triclinic_dimensions.astype(np.float64))))


class _AtomGroup:

    def __init__(self, universe, sel):
        self._u, self._sel = universe, sel

    @property
    def positions(self):
        return self._u._buffer[self._sel]

    def __getitem__(self, sel):
        return _AtomGroup(self._u, np.arange(len(self._u._types))[self._sel][sel])

    def __iter__(self):
        return iter(_TypedSite(self._u._types[i]) for i in np.arange(len(self._u._types))[self._sel])

    def __len__(self):
        return len(np.arange(len(self._u._types))[self._sel])


class _Reader:

    def __init__(self, universe, frames):
        self._u, self._frames = universe, frames

    def __getitem__(self, sel):
        return _Reader(self._u, self._frames[sel])

    def __iter__(self):
        for f in self._frames:
            self._u._buffer[...] = self._u._positions[f]  # the reader overwrites the one position buffer
            yield _Timestep(self._u._cells[f], self._u._positions.dtype)

    def __len__(self):
        return len(self._frames)


class DuckUniverse:

    def __init__(self, types, positions, cells, dtype=np.float32):
        """positions [F, N, 3] Cartesian, cells [F, 3, 3]; dtype float32 is what MDAnalysis hands out."""
        self._types = list(types)
        self._positions = np.asarray(positions, dtype=dtype)
        self._cells = np.asarray(cells, dtype=np.float64)
        self._buffer = np.zeros(self._positions.shape[1:], dtype=dtype)
        self.trajectory = _Reader(self, np.arange(self._positions.shape[0]))
        self.atoms = _AtomGroup(self, slice(None))


def universe_from_arrays(types, frac, latt, dtype=np.float32):
    """frac [N, F, 3], latt [F, 3, 3] -> DuckUniverse holding the Cartesian positions frac . latt."""
    cart = np.einsum('nfk,fkj->fnj', frac, latt)
    return DuckUniverse(types, cart, latt, dtype)
