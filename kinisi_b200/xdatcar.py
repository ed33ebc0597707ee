"""
A reader for VASP XDATCAR files, so that `from_file` (kinisi/analyzer.py:199-234) works where pymatgen is not installed.

The reference hands the file to `pymatgen.io.vasp.Xdatcar` and walks `.structures`; this module parses the text into
frame-major arrays and exposes the same surface the pymatgen front-end reads (`.structures[i].frac_coords`,
`.lattice.matrix`, `.volume`, sites with `.specie`), as views into those arrays.  Both XDATCAR layouts are read: one header
followed by `Direct configuration=` blocks (NVT) and a header repeated before every block (NPT, variable cell).
"""
import gzip
from typing import List

import numpy as np


class _Lattice:

    def __init__(self, matrix):
        self.matrix = matrix

    @property
    def volume(self):
        return float(abs(np.linalg.det(self.matrix)))


class _Site:

    def __init__(self, specie, frac_coords):
        self.specie = specie
        self.species_string = specie
        self.frac_coords = frac_coords


class XdatcarStructure:
    """One configuration: `frac_coords [site, 3]`, `lattice.matrix [3, 3]`, `volume`, iteration over sites."""

    def __init__(self, species: List[str], frac_coords: np.ndarray, matrix: np.ndarray):
        self._species = species
        self.frac_coords = frac_coords
        self.lattice = _Lattice(matrix)

    @property
    def volume(self):
        return self.lattice.volume

    def __len__(self):
        return len(self._species)

    def __iter__(self):
        return (_Site(s, self.frac_coords[i]) for i, s in enumerate(self._species))

    def __getitem__(self, i):
        return _Site(self._species[i], self.frac_coords[i])


def _header(lines, i):
    """Parses the 7 header lines starting at lines[i]: comment, scale, three lattice vectors, species, counts."""
    scale = float(lines[i + 1].split()[0])
    matrix = np.array([[float(x) for x in lines[i + 2 + k].split()[:3]] for k in range(3)])
    if scale < 0:  # VASP: a negative scale is the cell volume
        scale = (-scale / abs(np.linalg.det(matrix)))**(1.0 / 3.0)
    names = lines[i + 5].split()
    counts = [int(x) for x in lines[i + 6].split()]
    if len(names) != len(counts):
        raise ValueError('XDATCAR: the species and count lines do not match')
    species = [n for n, c in zip(names, counts) for _ in range(c)]
    return matrix * scale, species


class Xdatcar:
    """
    The configurations of an XDATCAR file (plain text or .gz).

    :param filename: Path of the file.

    `structures` lists the configurations; `frac_coords [frame, site, 3]` and `lattices [frame, 3, 3]` hold the same
    data as arrays.
    """

    def __init__(self, filename: str):
        opener = gzip.open if str(filename).endswith('.gz') else open
        with opener(filename, 'rt') as f:
            lines = f.read().split('\n')
        matrix, species = _header(lines, 0)
        n = len(species)
        frames, lattices = [], []
        i = 7
        while i < len(lines):
            word = lines[i].strip().split(' ')[0].lower() if lines[i].strip() else ''
            if word.startswith('direct') or word.startswith('cartesian'):
                block = np.array(' '.join(lines[i + 1:i + 1 + n]).split(), dtype=np.float64)
                if block.size < 3 * n:
                    break  # a truncated last configuration
                xyz = block.reshape(n, -1)[:, :3]
                if word.startswith('cartesian'):
                    xyz = np.dot(xyz, np.linalg.inv(matrix))
                frames.append(xyz)
                lattices.append(matrix)
                i += n + 1
            elif word == '':
                i += 1
            else:  # a repeated header (variable cell)
                matrix, species_again = _header(lines, i)
                if len(species_again) != n:
                    raise ValueError('XDATCAR: the number of atoms changes between configurations')
                i += 7
        if not frames:
            raise ValueError('XDATCAR: no configurations found')
        self.site_symbols = species
        self.frac_coords = np.stack(frames, axis=0)
        self.lattices = np.stack(lattices, axis=0)
        self.structures = [XdatcarStructure(species, self.frac_coords[f], self.lattices[f]) for f in range(len(frames))]
