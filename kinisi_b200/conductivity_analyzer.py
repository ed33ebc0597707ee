"""`ConductivityAnalyzer`: mean-squared charge displacement and ionic conductivity
(mirror of kinisi/conductivity_analyzer.py)."""
from typing import Union

import numpy as np

from . import diffusion
from .analyzer import Analyzer


class ConductivityAnalyzer(Analyzer):
    """
    MSCD with per-ion charges and the Bayesian regression of sigma (kinisi/conductivity_analyzer.py:16-266).
    """

    def __init__(self, delta_t, disp_3d, n_o, volume):
        super().__init__(delta_t, disp_3d, n_o, volume)
        self._diff = None

    def to_dict(self) -> dict:
        my_dict = super().to_dict()
        my_dict['diff'] = self._diff.to_dict()
        return my_dict

    @classmethod
    def _build(cls, maker, trajectory, parser_params, dtype, uncertainty_params, ionic_charge):
        if uncertainty_params is None:
            uncertainty_params = {}
        an = maker(trajectory, parser_params, dtype=dtype)
        an._diff = diffusion.MSCDBootstrap(an._delta_t, an._disp_3d, ionic_charge, an._n_o, **uncertainty_params)
        return an

    @classmethod
    def from_pymatgen(cls, trajectory, parser_params: dict, dtype: str = None, uncertainty_params: dict = None,
                      ionic_charge: Union[np.ndarray, int] = 1):
        return cls._build(super()._from_pymatgen, trajectory, parser_params, dtype, uncertainty_params, ionic_charge)

    @classmethod
    def from_ase(cls, trajectory, parser_params: dict, dtype: str = None, uncertainty_params: dict = None,
                 ionic_charge: Union[np.ndarray, int] = 1):
        return cls._build(super()._from_ase, trajectory, parser_params, dtype, uncertainty_params, ionic_charge)

    @classmethod
    def from_Xdatcar(cls, trajectory, parser_params: dict, dtype: str = None, uncertainty_params: dict = None,
                     ionic_charge: Union[np.ndarray, int] = 1):
        return cls._build(super()._from_Xdatcar, trajectory, parser_params, dtype, uncertainty_params, ionic_charge)

    @classmethod
    def from_file(cls, trajectory, parser_params: dict, dtype: str = None, uncertainty_params: dict = None,
                  ionic_charge: Union[np.ndarray, int] = 1):
        return cls._build(super()._from_file, trajectory, parser_params, dtype, uncertainty_params, ionic_charge)

    @classmethod
    def from_universe(cls, trajectory, parser_params: dict, dtype: str = None, uncertainty_params: dict = None,
                      ionic_charge: Union[np.ndarray, int] = 1):
        return cls._build(super()._from_universe, trajectory, parser_params, dtype, uncertainty_params, ionic_charge)

    def conductivity(self, start_dt: float, temperature: float, conductivity_params: Union[dict, None] = None):
        """
        Calculate the ionic conductivity (kinisi/conductivity_analyzer.py:224-237).

        :param start_dt: The start of the diffusive regime.
        :param temperature: Simulation temperature in Kelvin.
        :param conductivity_params: Parameters for :py:meth:`kinisi_b200.diffusion.Bootstrap.bootstrap_GLS`.
        """
        if conductivity_params is None:
            conductivity_params = {}
        self._diff.conductivity(start_dt, temperature, self._volume, **conductivity_params)

    @property
    def mscd(self) -> np.ndarray:
        """:return: MSCD for the input trajectories."""
        return self._diff.n

    @property
    def mscd_std(self) -> np.ndarray:
        """:return: MSCD standard deviation values (a single standard deviation)."""
        return self._diff.s

    @property
    def sigma(self):
        """:returns: Conductivity, in mS cm^-1."""
        return self._diff.sigma

    @property
    def flatchain(self) -> np.ndarray:
        """:return: sampling flatchain"""
        return np.array([self.sigma.samples, self.intercept.samples]).T
