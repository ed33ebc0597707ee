"""
Host-side mirror of kinisi/parser.py for the displacement-statistics path: same class names,
argument meaning, defaults and error behaviour, with the arithmetic executed by the sm_100a kernels
behind include/kinisi_b200.h.

What differs from the reference by design
  * `Parser.get_disp` returns a device-resident :py:class:`Trajectory` (coordinates + lattices in HBM)
    instead of a numpy array; unwrapping, the cumulative sum and the drift correction are fused into
    one pass when the `Parser` is constructed (kinisi/parser.py:101-148).  `np.asarray(traj)` still
    yields the reference's `[site, time step, axis]` array.
  * `disp_3d` is a lazy sequence (:py:class:`LazyDisplacements`): the O(N*T*n_dt) list the reference
    materialises (parser.py:208-216) is never built; one entry is materialised only if user code
    indexes it.  `memory_limit` therefore guards the device-resident displacement array instead.
"""
import warnings
from typing import List, Tuple, Union

import numpy as np
import torch

from . import dist, engine
from ._lib import MODE_DISPLACEMENT, MODE_FRACTIONAL


def _default_device(device=None):
    if device is not None:
        return torch.device(device)
    engine.require_cuda()
    return torch.device('cuda', torch.cuda.current_device())


class Trajectory:
    """
    Device-resident input of stage 1: coordinates `[site, frame, axis]` (float32 or float64) plus, in
    fractional mode, one lattice per frame.  Frame 0 is the duplicated first frame the front-ends insert
    (kinisi/parser.py:323-324), so a trajectory of F frames yields F-1 displacement frames.
    """

    def __init__(self, coords, latt=None, mode=MODE_FRACTIONAL, device=None):
        device = _default_device(device)
        self.mode = mode
        self.coords = engine.to_device(coords, device)
        if self.coords.dtype not in (torch.float32, torch.float64):
            self.coords = self.coords.to(torch.float64)
        if self.coords.dim() != 3 or self.coords.shape[2] != 3:
            raise ValueError('coordinates must have the shape [site, frame, 3]')
        self.latt = self.latt_inv = None
        self.latt_stride = 9
        if mode == MODE_FRACTIONAL:
            latt_host = latt if isinstance(latt, np.ndarray) else None
            latt = engine.to_device(latt, device, torch.float64).reshape(-1, 3, 3)
            if latt.shape[0] == 1:
                self.latt_stride = 0
            elif latt.shape[0] != self.coords.shape[1]:
                raise ValueError('one lattice per frame is required')
            elif latt_host is not None and np.all(latt_host == latt_host[0]):
                self.latt_stride = 0  # constant cell: every frame reads latt[0]
            self.latt = latt
            self.latt_inv = engine.invert_lattice(latt)

    @property
    def device(self):
        return self.coords.device

    @property
    def n_atoms(self):
        return self.coords.shape[0]

    @property
    def n_frames(self):
        """Number of displacement frames (the reference's drift_corrected.shape[1])."""
        return self.coords.shape[1] - (1 if self.mode == MODE_FRACTIONAL else 0)

    def frame_sums(self, idx, weights=None):
        """Per-frame (weighted) displacement sums over atoms idx, all-reduced over ranks."""
        pitch = engine.pitch_for(self.n_frames)
        return engine.frame_sums(self.coords, self.mode, self.latt, self.latt_inv, self.latt_stride, idx, weights, pitch)

    def displacement(self, idx, drift=None):
        """dc [len(idx), 3, pitch] float64, component-major.  `drift` is None or (frame_sums, n_framework): the
        per-frame sums over the framework atoms, whose mean is removed before the cumulative sum."""
        pitch = engine.pitch_for(self.n_frames)
        sums, scale = (None, 0.0) if drift is None else (drift[0], 1.0 / drift[1])
        return engine.unwrap_cumsum(self.coords, self.mode, self.latt, self.latt_inv, self.latt_stride, idx, sums, scale,
                                    pitch)

    def numpy(self, idx=None, drift=None):
        """The reference layout `[site, time step, axis]` on the host."""
        if idx is None:
            idx = torch.arange(self.n_atoms, dtype=torch.int32, device=self.device)
        dc = self.displacement(idx, drift)
        return dc[:, :, :self.n_frames].permute(0, 2, 1).contiguous().cpu().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype)

    @property
    def shape(self):
        return (self.n_atoms, self.n_frames, 3)


class LazyDisplacements:
    """
    Stand-in for the reference's `disp_3d` list (kinisi/parser.py:184, 208-216): a sequence with one
    entry per kept time interval, entry i being the `[atom, observation, axis]` array for lag k_i.
    Entries are materialised on demand from the device-resident cumulative displacement; the moment
    kernels never do that.
    """

    def __init__(self, dc, n_frames, lags, n_atoms_global=None, stride=1):
        self.dc = dc  # [n_local_atoms, 3, pitch] float64 CUDA
        self.n_frames = int(n_frames)
        self.lags = np.asarray(lags, dtype=np.int64)
        self.n_atoms = dc.shape[0]
        self.n_atoms_global = int(n_atoms_global if n_atoms_global is not None else dc.shape[0])

    def __len__(self):
        return len(self.lags)

    def shape_of(self, i):
        return (self.n_atoms, self.n_frames - int(self.lags[i]) + 1, 3)

    @property
    def shapes(self):
        return [self.shape_of(i) for i in range(len(self))]

    def __getitem__(self, i):
        if isinstance(i, slice):
            return LazyDisplacements(self.dc, self.n_frames, self.lags[i], self.n_atoms_global)
        k = int(self.lags[i])
        x = self.dc[:, :, :self.n_frames]
        first = x[:, :, k - 1:k]
        rest = x[:, :, k:] - x[:, :, :self.n_frames - k]
        return torch.cat([first, rest], dim=2).permute(0, 2, 1).contiguous().cpu().numpy()

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def stacked_with(self, others):
        """`Analyzer._stack_trajectories` (kinisi/analyzer.py:274-290): more atoms, same lags."""
        dc = torch.cat([self.dc] + [o.dc for o in others], dim=0)
        return LazyDisplacements(dc, self.n_frames, self.lags, self.n_atoms_global + sum(o.n_atoms_global for o in others))


class Parser:
    """
    The base class for parsing (kinisi/parser.py:23-222).

    :param disp: Displacements of atoms with the shape [site, time step, axis] (numpy / torch), or the
        :py:class:`Trajectory` returned by :py:meth:`get_disp`.
    :param indices: Indices for the atoms in the trajectory used in the diffusion calculation.
    :param drift_indices: Indices for the atoms in the trajectory that should not be used in the diffusion
        calculation, instead being used to correct for framework drift.
    :param time_step: Time step, in picoseconds, between steps in trajectory.
    :param step_skip: Sampling freqency of the trajectory.
    :param min_dt: Minimum time interval to be evaluated. Optional, defaults to time_step * step_skip.
    :param max_dt: Maximum time interval to be evaluated. Optional, defaults to the length of the simulation.
    :param n_steps: Number of steps in the time interval grid. Optional, defaults to 100.
    :param spacing: 'linear' or 'logarithmic'. Optional, defaults to 'linear'.
    :param sampling: 'multi-origin' or 'single-origin'. Optional, defaults to 'multi-origin'.
    :param memory_limit: Upper limit, in GB, on the memory the displacements may occupy (here: the
        device-resident cumulative displacement array). Optional, defaults to 8.
    :param progress: Accepted for compatibility.
    :param device: CUDA device. Optional, defaults to the current device.
    :param distributed: If True and torch.distributed is initialised, `disp` holds this rank's shard of
        the atoms; drift sums, atom counts and (later) moment sums are all-reduced.
    """

    def __init__(self,
                 disp,
                 indices,
                 drift_indices,
                 time_step: float,
                 step_skip: int,
                 min_dt: float = None,
                 max_dt: float = None,
                 n_steps: int = 100,
                 spacing: str = 'linear',
                 sampling: str = 'multi-origin',
                 memory_limit: float = 8.,
                 progress: bool = True,
                 device=None,
                 distributed: bool = False):
        self.time_step = time_step
        self.step_skip = step_skip
        self.indices = indices
        self.drift_indices = drift_indices
        self.memory_limit = memory_limit
        self.min_dt = min_dt
        self.max_dt = max_dt
        self.sampling = sampling
        self._volume = None
        self._distributed = bool(distributed) and dist.active()

        if isinstance(disp, Trajectory):
            traj = disp
        else:
            traj = Trajectory(disp, mode=MODE_DISPLACEMENT, device=device)
        self._traj = traj
        dev = traj.device
        n_frames = traj.n_frames

        idx = engine.small_to_device(indices, dev, np.int32)
        n_sel_local = idx.numel()
        n_sel = int(round(dist.allreduce_sum_scalar(n_sel_local, dev))) if self._distributed else n_sel_local
        self._n_atoms_global = n_sel

        # memory guard (kinisi/parser.py:190-200), applied to what this build keeps resident
        need_gb = n_sel_local * 3 * engine.pitch_for(n_frames) * 8 * 1e-9
        if need_gb > self.memory_limit:
            raise MemoryError(f"The memory limit of this job is {self.memory_limit:.1e} GB but the "
                              f"displacement values will use {need_gb:.1e} GB. Please either increase "
                              "the memory_limit parameter or descrease the sampling rate (see "
                              "https://kinisi.readthedocs.io/en/latest/memory_limit.html).")

        # drift correction (kinisi/parser.py:131-148) fused with the displacement construction
        self._drift = self._framework_mean(traj, drift_indices)
        if n_sel_local > 0:
            self._dc_sel = traj.displacement(idx, self._drift)
        else:
            self._dc_sel = torch.zeros((0, 3, engine.pitch_for(n_frames)), dtype=torch.float64, device=dev)

        if self.max_dt is None:
            self.max_dt = n_frames * (self.step_skip * self.time_step)
        if self.min_dt is None:
            self.min_dt = self.step_skip * self.time_step
        min_dt_int = int(self.min_dt / (self.step_skip * self.time_step))
        if min_dt_int >= n_frames:
            raise ValueError('The passed min_dt is greater than or equal to the maximum simulation length.')
        if n_steps > (n_frames - min_dt_int) + 1:
            n_steps = int(n_frames - min_dt_int) + 1

        self.time_intervals = self.get_time_intervals(n_steps, spacing)
        self.delta_t, self.disp_3d, self._n_o = self.get_disps(self.time_intervals, self._dc_sel, progress)

    def _framework_mean(self, traj, drift_indices):
        """(per-frame sums over the framework atoms [3, pitch] on the device, number of framework atoms), or None.
        The displacement kernel subtracts sums / n before its cumulative sum (kinisi/parser.py:143-145)."""
        n_fw_local = len(drift_indices)
        n_fw = int(round(dist.allreduce_sum_scalar(n_fw_local, traj.device))) if self._distributed else n_fw_local
        if n_fw == 0:
            return None
        if n_fw_local > 0:
            fw = engine.small_to_device(drift_indices, traj.device, np.int32)
            sums = traj.frame_sums(fw)
        else:
            sums = torch.zeros((3, engine.pitch_for(traj.n_frames)), dtype=torch.float64, device=traj.device)
        if self._distributed:
            dist.allreduce_sum_(sums)
        return sums, float(n_fw)

    @property
    def volume(self) -> float:
        """:return: Volume of system, in cubic angstrom."""
        return self._volume

    @property
    def dc(self) -> np.ndarray:
        """Drift-corrected displacements of ALL atoms, `[site, time step, axis]` (host copy, built on demand)."""
        return self._traj.numpy(drift=self._drift)

    @staticmethod
    def get_disp(coords, latt, progress: bool = True, device=None) -> Trajectory:
        """
        Stage the fractional coordinates and lattices on the device (kinisi/parser.py:101-129).  The
        unwrapping itself (TOR scheme, doi: 10.1021/acs.jctc.3c00308) runs fused with the drift correction
        when the :py:class:`Parser` is built.  Issues the reference's warning about non-orthorhombic cells.

        :param coords: Fractional coordinates: the reference's list of `[site, 1, axis]` arrays (one per
            frame, first frame duplicated) or one `[site, frame, axis]` array / tensor.
        :param latt: Lattice descriptions, `[frame, 3, 3]` (or a list of 3x3 arrays).
        """
        if isinstance(coords, (list, tuple)):
            coords = np.concatenate(coords, axis=1)
        if isinstance(latt, (list, tuple)):
            latt = np.array(latt)
        latt_np = latt.detach().cpu().numpy() if isinstance(latt, torch.Tensor) else np.asarray(latt)
        off_diags = np.array([[False, True, True], [True, False, True], [True, True, False]])
        if np.any(latt_np.reshape(-1, 3, 3)[:, off_diags] != 0):
            warnings.warn(
                'Converting triclinic cell to orthorhombic this may have unexpected results. '
                'Triclinic a, b, c are not equilivalent to orthorhombic x, y, z.', UserWarning)
        return Trajectory(coords, latt_np.reshape(-1, 3, 3), MODE_FRACTIONAL, device)

    @staticmethod
    def correct_drift(drift_indices, disp) -> np.ndarray:
        """
        Drift correction on its own (kinisi/parser.py:131-148): `disp - mean(disp[drift_indices])`.

        :return: Displacements corrected to account for drift of a framework, `[site, time step, axis]`.
        """
        traj = disp if isinstance(disp, Trajectory) else Trajectory(disp, mode=MODE_DISPLACEMENT)
        if len(drift_indices) == 0:
            return traj.numpy()
        fw = engine.small_to_device(drift_indices, traj.device, np.int32)
        return traj.numpy(drift=(traj.frame_sums(fw), float(len(drift_indices))))

    def get_time_intervals(self, n_steps: int, spacing: str) -> np.ndarray:
        """The integer time-interval grid (kinisi/parser.py:150-168).  Integer work: stays on the host."""
        unit = self.step_skip * self.time_step
        min_dt = int(self.min_dt / unit)
        max_dt = int(self.max_dt / unit)
        if min_dt == 0:
            min_dt = 1
        if spacing == 'linear':
            return np.linspace(min_dt, max_dt, n_steps, dtype=int)
        elif spacing == 'logarithmic':
            return np.unique(np.geomspace(min_dt, max_dt, n_steps, dtype=int))
        else:
            raise ValueError("Only linear or logarithmic spacing is allowed.")

    def get_disps(self, time_intervals: np.ndarray, drift_corrected, progress: bool = True):
        """
        The bookkeeping of kinisi/parser.py:170-222 without the displacement list: `delta_t` for every
        interval, `n_o = N * k_last / k` (float) and the lazy `disp_3d` for the intervals the reference keeps.
        """
        delta_t = time_intervals * self.time_step * self.step_skip
        n_frames = self._traj.n_frames
        n_atoms = self._n_atoms_global
        if self.sampling not in ('multi-origin', 'single-origin'):
            raise ValueError(f"The sampling option of {self.sampling} is unrecognized, "
                             "please use 'multi-origin' or 'single-origin'.")
        if np.any(time_intervals > n_frames):
            raise IndexError('a time interval exceeds the length of the trajectory')
        kept, n_samples = [], []
        for i, k in enumerate(time_intervals):
            n_obs = 1 if self.sampling == 'single-origin' else n_frames - int(k) + 1
            if n_atoms * len(range(0, n_obs, int(k))) <= 1:  # parser.py:204-205, 214-215
                continue
            kept.append(i)
            n_samples.append(n_atoms * time_intervals[-1] / k)  # parser.py:221
        kept = np.array(kept, dtype=int)
        if self.sampling == 'single-origin':
            # parser.py:203: entry i is the single frame dc[:, i] (indexed by grid position, not by lag)
            disp_3d = SingleOriginDisplacements(drift_corrected, n_frames, kept, n_atoms)
        else:
            disp_3d = LazyDisplacements(drift_corrected, n_frames, time_intervals[kept], n_atoms)
        return delta_t, disp_3d, np.array(n_samples, dtype=float)


class SingleOriginDisplacements(LazyDisplacements):
    """`sampling='single-origin'` (kinisi/parser.py:202-206): entry i is `dc[:, i:i+1]`."""

    def __init__(self, dc, n_frames, frames, n_atoms_global=None):
        super().__init__(dc, n_frames, np.ones(len(frames), dtype=np.int64), n_atoms_global)
        self.frames = np.asarray(frames, dtype=np.int64)

    def shape_of(self, i):
        return (self.n_atoms, 1, 3)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return SingleOriginDisplacements(self.dc, self.n_frames, self.frames[i], self.n_atoms_global)
        f = int(self.frames[i])
        return self.dc[:, :, f:f + 1].permute(0, 2, 1).contiguous().cpu().numpy()

    def device_entry(self, i):
        f = int(self.frames[i])
        return self.dc[:, :, f:f + 1].permute(0, 2, 1).contiguous()


def _frames_to_arrays(frames, get_frac, get_cell, sub_sample_traj, progress):
    """Walks a trajectory object frame by frame (kinisi/parser.py:309-325, 458-473) into one
    `[site, frame, 3]` array and one `[frame, 3, 3]` array, first frame duplicated."""
    structure = None
    coords, latt = [], []
    for struct in frames[::sub_sample_traj]:
        if structure is None:
            structure = struct
        coords.append(np.asarray(get_frac(struct), dtype=np.float64))
        latt.append(np.asarray(get_cell(struct), dtype=np.float64))
    coords.insert(0, coords[0])
    latt.insert(0, latt[0])
    return structure, np.stack(coords, axis=1), np.stack(latt, axis=0)


def _get_framework(structure, indices, framework_indices):
    """kinisi/parser.py:741-764 (1-based `framework_indices`, as the reference subtracts one)."""
    if isinstance(framework_indices, (list, tuple)):
        return indices, list(np.array(framework_indices, dtype=int) - 1)
    chosen = set(int(i) for i in indices)
    return indices, [i for i, _ in enumerate(structure) if i not in chosen]


class ASEParser(Parser):
    """
    A parser for ASE Atoms objects (kinisi/parser.py:225-356).  Any object exposing
    `get_scaled_positions()`, `.cell[:]`, `get_volume()` and iteration over sites with `.symbol` works.

    :param atoms: Atoms ordered in sequence of run.
    :param specie: symbol to calculate diffusivity for as a String, e.g. 'Li'.
    :param time_step: Time step, in picoseconds, between steps in trajectory.
    :param step_skip: Sampling frequency of the trajectory.
    :param sub_sample_traj: Multiple of the time_step to sub sample at. Optional, defaults to 1.
    :param specie_indices: Optional, list of indices to calculate diffusivity for. Specie must be None.
    :param masses: Optional, masses for molecule centres (not on the displacement-statistics path).
    :param framework_indices: Optional, indices used to correct framework drift; [] disables the correction.
    (min_dt, max_dt, n_steps, spacing, sampling, memory_limit, progress as for :py:class:`Parser`.)
    """

    def __init__(self,
                 atoms,
                 specie: str,
                 time_step: float,
                 step_skip: int,
                 sub_sample_traj: int = 1,
                 min_dt: float = None,
                 max_dt: float = None,
                 n_steps: int = 100,
                 spacing: str = 'linear',
                 sampling: str = 'multi-origin',
                 memory_limit: float = 8.,
                 progress: bool = True,
                 specie_indices: List[int] = None,
                 masses: List[float] = None,
                 framework_indices: List[int] = None,
                 device=None):
        structure, coords, latt = self.get_structure_coords_latt(atoms, sub_sample_traj, progress)

        if specie is not None:
            indices = self.get_indices(structure, specie, framework_indices)
        elif isinstance(specie_indices, (list, tuple)):
            if isinstance(specie_indices[0], (list, tuple)):
                raise NotImplementedError('molecular centres of mass (kinisi/parser.py:674-738) are outside the '
                                          'displacement-statistics path of kinisi_b200')
            indices = _get_framework(structure, specie_indices, framework_indices)
        else:
            raise TypeError('Unrecognized type for specie or specie_indices')

        self.coords_check = coords[:, 0:1]

        # note: like the reference (parser.py:291) sub_sample_traj is NOT folded into step_skip here
        super().__init__(self.get_disp(coords, latt, progress=progress, device=device), indices[0], indices[1], time_step,
                         step_skip, min_dt, max_dt, n_steps, spacing, sampling, memory_limit, progress, device=device)
        self._volume = structure.get_volume()

    @staticmethod
    def get_structure_coords_latt(atoms, sub_sample_traj: int = 1, progress: bool = True):
        """:return: initial structure, fractional coordinates `[site, frame, 3]`, lattices `[frame, 3, 3]`."""
        return _frames_to_arrays(atoms, lambda s: s.get_scaled_positions(), lambda s: s.cell[:], sub_sample_traj,
                                 progress)

    @staticmethod
    def get_indices(structure, specie, framework_indices) -> Tuple[list, list]:
        """Framework and non-framework indices (kinisi/parser.py:327-356)."""
        indices, drift_indices = [], []
        if not isinstance(specie, list):
            specie = [specie]
        for i, site in enumerate(structure):
            if site.symbol in specie:
                indices.append(i)
            else:
                drift_indices.append(i)
        if len(indices) == 0:
            raise ValueError("There are no species selected to calculate the mean-squared displacement of.")
        if isinstance(framework_indices, (list, tuple)):
            drift_indices = framework_indices
        return indices, drift_indices


class PymatgenParser(Parser):
    """
    A parser for pymatgen structures (kinisi/parser.py:359-508).  Any object exposing `.frac_coords`,
    `.lattice.matrix`, `.volume` and iteration over sites with `.specie` works.  `time_step` is in
    femtoseconds: `delta_t` is scaled by 1e-3 as in the reference (parser.py:440), and
    `sub_sample_traj` is folded into `step_skip` (parser.py:429).
    """

    def __init__(self,
                 structures,
                 specie,
                 time_step: float,
                 step_skip: int,
                 sub_sample_traj: int = 1,
                 min_dt: float = None,
                 max_dt: float = None,
                 n_steps: int = 100,
                 spacing: str = 'linear',
                 sampling: str = 'multi-origin',
                 memory_limit: float = 8.,
                 progress: bool = True,
                 specie_indices: List[int] = None,
                 masses: List[float] = None,
                 framework_indices: List[int] = None,
                 device=None):
        structure, coords, latt = self.get_structure_coords_latt(structures, sub_sample_traj, progress)
        if specie is not None:
            indices = self.get_indices(structure, specie, framework_indices)
        elif isinstance(specie_indices, (list, tuple)):
            if isinstance(specie_indices[0], (list, tuple)):
                raise NotImplementedError('molecular centres of mass (kinisi/parser.py:674-738) are outside the '
                                          'displacement-statistics path of kinisi_b200')
            indices = _get_framework(structure, specie_indices, framework_indices)
        else:
            raise TypeError('Unrecognized type for specie or specie_indices')
        self.coords_check = coords[:, 0:1]
        super().__init__(self.get_disp(coords, latt, progress=progress, device=device), indices[0], indices[1], time_step,
                         step_skip * sub_sample_traj, min_dt, max_dt, n_steps, spacing, sampling, memory_limit, progress,
                         device=device)
        self._volume = structure.volume
        warnings.warn("UserWarning: Be aware that the expected input unit for 'time_step' is femtoseconds, not picoseconds.")
        self.delta_t = self.delta_t * 1e-3

    @staticmethod
    def get_structure_coords_latt(structures, sub_sample_traj: int = 1, progress: bool = True):
        return _frames_to_arrays(structures, lambda s: s.frac_coords, lambda s: s.lattice.matrix, sub_sample_traj,
                                 progress)

    @staticmethod
    def get_indices(structure, specie, framework_indices) -> Tuple[list, list]:
        indices, drift_indices = [], []
        if not isinstance(specie, list):
            specie = [specie]
        for i, site in enumerate(structure):
            if site.specie.__str__() in specie:
                indices.append(i)
            else:
                drift_indices.append(i)
        if len(indices) == 0:
            raise ValueError("There are no species selected to calculate the mean-squared displacement of.")
        if isinstance(framework_indices, (list, tuple)):
            drift_indices = framework_indices
        return indices, drift_indices
