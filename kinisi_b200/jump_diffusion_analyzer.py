"""`JumpDiffusionAnalyzer`: mean-squared total displacement and jump-diffusion coefficient
(mirror of kinisi/jump_diffusion_analyzer.py)."""
from typing import Union

import numpy as np

from . import diffusion
from .analyzer import Analyzer


class JumpDiffusionAnalyzer(Analyzer):
    """
    MSTD (collective motion) and the Bayesian regression of D_J (kinisi/jump_diffusion_analyzer.py:16-248).
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
        an = maker(trajectory, parser_params, dtype=dtype)
        an._diff = diffusion.MSTDBootstrap(an._delta_t, an._disp_3d, an._n_o, **uncertainty_params)
        return an

    @classmethod
    def from_pymatgen(cls, trajectory, parser_params: dict, dtype: str = None, uncertainty_params: dict = None):
        return cls._build(super()._from_pymatgen, trajectory, parser_params, dtype, uncertainty_params)

    @classmethod
    def from_ase(cls, trajectory, parser_params: dict, dtype: str = None, uncertainty_params: dict = None):
        return cls._build(super()._from_ase, trajectory, parser_params, dtype, uncertainty_params)

    @classmethod
    def from_Xdatcar(cls, trajectory, parser_params: dict, dtype: str = None, uncertainty_params: dict = None):
        return cls._build(super()._from_Xdatcar, trajectory, parser_params, dtype, uncertainty_params)

    @classmethod
    def from_file(cls, trajectory, parser_params: dict, dtype: str = None, uncertainty_params: dict = None):
        return cls._build(super()._from_file, trajectory, parser_params, dtype, uncertainty_params)

    @classmethod
    def from_universe(cls, trajectory, parser_params: dict, dtype: str = None, uncertainty_params: dict = None):
        return cls._build(super()._from_universe, trajectory, parser_params, dtype, uncertainty_params)

    def jump_diffusion(self, start_dt: float, jump_diffusion_params: Union[dict, None] = None):
        """
        Calculate the jump diffusion coefficient (kinisi/jump_diffusion_analyzer.py:207-219).

        :param start_dt: The start of the diffusive regime.
        :param jump_diffusion_params: Parameters for :py:meth:`kinisi_b200.diffusion.Bootstrap.bootstrap_GLS`.
        """
        if jump_diffusion_params is None:
            jump_diffusion_params = {}
        self._diff.jump_diffusion(start_dt, **jump_diffusion_params)

    @property
    def mstd(self) -> np.ndarray:
        """:return: MSTD for the input trajectories."""
        return self._diff.n

    @property
    def mstd_std(self) -> np.ndarray:
        """:return: MSTD standard deviation values (a single standard deviation)."""
        return self._diff.s

    @property
    def D_J(self):
        """:return: Jump diffusion coefficient"""
        return self._diff.D_J

    @property
    def flatchain(self) -> np.ndarray:
        """:return: sampling flatchain"""
        return np.array([self.D_J.samples, self.intercept.samples]).T
