"""
Base class of the three analyzers (mirror of kinisi/analyzer.py:14-349 for the displacement-statistics
path): routes a trajectory through a parser and keeps `(delta_t, disp_3d, n_o, volume)`.

HDF5 save/load (kinisi/analyzer.py:33-67) is persistence, not part of this path: `save`/`load` raise.
"""
from typing import Union

import numpy as np

from .parser import MDAnalysisParser, ASEParser, LazyDisplacements, PymatgenParser


def _flatten_list(this_list: list) -> list:
    """Flatten nested lists (kinisi/analyzer.py:351-359)."""
    return [item for sublist in this_list for item in sublist]


class Analyzer:
    """
    Superclass of :py:class:`DiffusionAnalyzer`, :py:class:`JumpDiffusionAnalyzer` and
    :py:class:`ConductivityAnalyzer`.

    :param delta_t: An array of the time interval values.
    :param disp_3d: The displacements, one entry per delta_t (lazy sequence or list of arrays).
    :param n_o: Number of independent observations per time interval.
    :param volume: The volume of the simulation cell.
    """

    def __init__(self, delta_t: np.ndarray, disp_3d, n_o: np.ndarray, volume: float):
        self._delta_t = delta_t
        self._disp_3d = disp_3d
        self._n_o = n_o
        self._volume = volume

    def save(self, filename: str):
        raise NotImplementedError('HDF5 persistence (kinisi/analyzer.py:33-67) is outside the path kinisi_b200 builds')

    @classmethod
    def load(cls, filename: str):
        raise NotImplementedError('HDF5 persistence (kinisi/analyzer.py:33-67) is outside the path kinisi_b200 builds')

    def to_dict(self) -> dict:
        """:return: Dictionary description of Analyzer."""
        return {'delta_t': self._delta_t, 'disp_3d': self._disp_3d, 'n_o': self._n_o, 'volume': self._volume}

    @classmethod
    def from_dict(cls, my_dict: dict) -> 'Analyzer':
        """An analyzer from its dictionary (kinisi/analyzer.py:75-85; with the bootstrap under 'diff' as in
        kinisi/diffusion_analyzer.py:46-57 and the two sibling classes)."""
        an = cls(my_dict['delta_t'], my_dict['disp_3d'], my_dict['n_o'], my_dict['volume'])
        if my_dict.get('diff') is not None:
            from . import diffusion
            an._diff = diffusion.Bootstrap.from_dict(my_dict['diff'])
        return an

    @classmethod
    def _from_parser(cls, make_parser, trajectory, dtype, flatten, **kwargs):
        """The dtype dispatch shared by every `_from_*` (kinisi/analyzer.py:108-122 and siblings)."""
        if dtype is None:
            u = make_parser(trajectory)
            return cls(u.delta_t, u.disp_3d, u._n_o, u.volume, **kwargs)
        elif dtype == 'identical':
            u = [make_parser(f) for f in trajectory]
            n_o = np.zeros(u[0]._n_o.size)
            for i in u:
                n_o += i._n_o
            return cls(u[0].delta_t, cls._stack_trajectories(u), n_o, u[0].volume, **kwargs)
        elif dtype == 'consecutive':
            if flatten is None:
                raise ValueError('The dtype specified was not recognised, please consult the kinisi documentation.')
            u = make_parser(flatten(trajectory))
            return cls(u.delta_t, u.disp_3d, u._n_o, u.volume, **kwargs)
        else:
            raise ValueError('The dtype specified was not recognised, please consult the kinisi documentation.')

    @classmethod
    def _from_pymatgen(cls, trajectory, parser_params: dict, dtype: str = None, **kwargs):
        """From a list or nested list of pymatgen Structure-like objects (kinisi/analyzer.py:87-122)."""
        return cls._from_parser(lambda t: PymatgenParser(t, **parser_params), trajectory, dtype, _flatten_list, **kwargs)

    @classmethod
    def _from_ase(cls, trajectory, parser_params: dict, dtype: str = None, **kwargs):
        """From a list or nested list of ase Atoms-like objects (kinisi/analyzer.py:124-159)."""
        return cls._from_parser(lambda t: ASEParser(t, **parser_params), trajectory, dtype, _flatten_list, **kwargs)

    @classmethod
    def _from_Xdatcar(cls, trajectory, parser_params: dict, dtype: str = None, **kwargs):
        """From one or a list of pymatgen Xdatcar-like objects exposing `.structures` (analyzer.py:161-197)."""
        if dtype is None:
            return cls._from_parser(lambda t: PymatgenParser(t.structures, **parser_params), trajectory, None, None,
                                    **kwargs)
        return cls._from_parser(lambda t: PymatgenParser(t if isinstance(t, list) else t.structures, **parser_params),
                                trajectory, dtype, lambda tr: _flatten_list([x.structures for x in tr]), **kwargs)

    @classmethod
    def _from_file(cls, trajectory, parser_params: dict, dtype: str = None, **kwargs):
        """From one or a list of XDATCAR files (kinisi/analyzer.py:199-234).  The reference reads them with
        pymatgen.io.vasp.Xdatcar; where pymatgen is not installed the built-in reader (kinisi_b200/xdatcar.py) is used."""
        try:
            from pymatgen.io.vasp import Xdatcar
        except ModuleNotFoundError:
            from .xdatcar import Xdatcar
        if dtype is None:
            return cls._from_Xdatcar(Xdatcar(trajectory), parser_params, None, **kwargs)
        return cls._from_Xdatcar([Xdatcar(f) for f in trajectory], parser_params, dtype, **kwargs)

    @classmethod
    def _from_universe(cls, trajectory, parser_params: dict, dtype: str = None, **kwargs):
        """From an MDAnalysis Universe-like object, or a list of them with dtype='identical' (kinisi/analyzer.py:236-271;
        'consecutive' is not offered for universes, as in the reference).  MDAnalysis itself is not imported: any object
        with the attributes :py:class:`MDAnalysisParser` reads will do."""
        return cls._from_parser(lambda t: MDAnalysisParser(t, **parser_params), trajectory, dtype, None, **kwargs)

    @staticmethod
    def _stack_trajectories(u) -> LazyDisplacements:
        """
        Several trajectories stacked so that they look like additional atoms (kinisi/analyzer.py:274-290):
        here the device-resident displacement arrays are concatenated along the atom axis.
        """
        return u[0].disp_3d.stacked_with([p.disp_3d for p in u[1:]])

    def posterior_predictive(self, posterior_predictive_params: Union[dict, None] = None):
        """Sample the posterior predictive distribution (kinisi/analyzer.py:292-304)."""
        if posterior_predictive_params is None:
            posterior_predictive_params = {}
        return self._diff.posterior_predictive(**posterior_predictive_params)

    @property
    def distribution(self) -> np.ndarray:
        """:return: Samples of the linear relationship, for plotting credible intervals."""
        return self._diff.gradient.samples * self.dt[:, np.newaxis] + self._diff.intercept.samples

    @property
    def dt(self) -> np.ndarray:
        """:return: Timestep values that have been sampled."""
        return self._diff.dt

    @property
    def dr(self):
        """:return: Distributions of the Euclidean displacement at each dt (built on demand)."""
        return self._diff.euclidian_displacements

    @property
    def ngp_max(self) -> float:
        """:return: Position in dt where the non-Gaussian parameter is maximised."""
        return self.dt[self._diff.ngp.argmax()]

    @property
    def intercept(self):
        """:return: The distribution describing the intercept."""
        return self._diff.intercept

    @property
    def volume(self) -> float:
        """:return: Volume of system, in cubic angstrom."""
        return self._volume
