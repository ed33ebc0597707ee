"""`DiffusionAnalyzer`: mean-squared displacement and self-diffusion coefficient
(mirror of kinisi/diffusion_analyzer.py)."""
from typing import Union

import numpy as np

from . import diffusion
from .analyzer import Analyzer


class DiffusionAnalyzer(Analyzer):
    """
    MSD with its covariance model and the Bayesian regression of D (kinisi/diffusion_analyzer.py:16-245).

    :param delta_t: An array of the time interval values.
    :param disp_3d: The displacements, one entry per delta_t.
    :param n_o: Number of independent observations per time interval.
    :param volume: The volume of the simulation cell.
    """

    def __init__(self, delta_t, disp_3d, n_o, volume):
        super().__init__(delta_t, disp_3d, n_o, volume)
        self._diff = None

    def to_dict(self) -> dict:
        my_dict = super().to_dict()
        my_dict['diff'] = self._diff.to_dict()
        return my_dict

    @classmethod
    def _build(cls, maker, trajectory, parser_params, dtype, uncertainty_params):
        if uncertainty_params is None:
            uncertainty_params = {}
        diff = maker(trajectory, parser_params, dtype=dtype)
        diff._diff = diffusion.MSDBootstrap(diff._delta_t, diff._disp_3d, diff._n_o, **uncertainty_params)
        return diff

    @classmethod
    def from_pymatgen(cls, trajectory, parser_params: dict, dtype: str = None, uncertainty_params: dict = None):
        """From a list or nested list of pymatgen structures (kinisi/diffusion_analyzer.py:59-89)."""
        return cls._build(super()._from_pymatgen, trajectory, parser_params, dtype, uncertainty_params)

    @classmethod
    def from_ase(cls, trajectory, parser_params: dict, dtype: str = None, uncertainty_params: dict = None):
        """From a list or nested list of ase Atoms (kinisi/diffusion_analyzer.py:91-118)."""
        return cls._build(super()._from_ase, trajectory, parser_params, dtype, uncertainty_params)

    @classmethod
    def from_Xdatcar(cls, trajectory, parser_params: dict, dtype: str = None, uncertainty_params: dict = None):
        """From one or a list of pymatgen Xdatcar objects (kinisi/diffusion_analyzer.py:120-147)."""
        return cls._build(super()._from_Xdatcar, trajectory, parser_params, dtype, uncertainty_params)

    @classmethod
    def from_file(cls, trajectory, parser_params: dict, dtype: str = None, uncertainty_params: dict = None):
        """From one or a list of XDATCAR files (kinisi/diffusion_analyzer.py:149-175)."""
        return cls._build(super()._from_file, trajectory, parser_params, dtype, uncertainty_params)

    @classmethod
    def from_universe(cls, trajectory, parser_params: dict, dtype: str = None, uncertainty_params: dict = None):
        """From an MDAnalysis universe (kinisi/diffusion_analyzer.py:177-202)."""
        return cls._build(super()._from_universe, trajectory, parser_params, dtype, uncertainty_params)

    def diffusion(self, start_dt: float, diffusion_params: Union[dict, None] = None):
        """
        Calculate the diffusion coefficient (kinisi/diffusion_analyzer.py:204-216).

        :param start_dt: The start of the diffusive regime.
        :param diffusion_params: Parameters for :py:meth:`kinisi_b200.diffusion.Bootstrap.bootstrap_GLS`.
        """
        if diffusion_params is None:
            diffusion_params = {}
        self._diff.diffusion(start_dt, **diffusion_params)

    @property
    def msd(self) -> np.ndarray:
        """:return: MSD for the input trajectories."""
        return self._diff.n

    @property
    def msd_std(self) -> np.ndarray:
        """:return: MSD standard deviation values (a single standard deviation)."""
        return self._diff.s

    @property
    def D(self):
        """:return: Diffusion coefficient distribution."""
        return self._diff.D

    @property
    def flatchain(self) -> np.ndarray:
        """:return: sampling flatchain"""
        return np.array([self.D.samples, self.intercept.samples]).T
