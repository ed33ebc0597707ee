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
from typing import List, Tuple

import contextlib

import numpy as np
import torch

from . import dist, engine
from ._lib import MODE_DISPLACEMENT, MODE_FRACTIONAL


_SIDE_STREAMS = {}
_GRIDS = {}  # (min_dt, max_dt, n_steps, spacing) -> integer time-interval grid
_LATTICES = {}  # (lattice bytes, device) -> (lattices, inverses, upload event) on the device
STAGE_BYTES = 32 << 20  # host->device copy granularity of the staged ingest (about 0.6 ms of PCIe time)
# A cumulative-displacement array above STREAM_BYTES is not kept: stage 1 is re-run block by block (STREAM_BLOCK_BYTES of
# float64 displacements at a time) in front of every consumer that reduces over atoms (BASELINE configs[4]: 100 000 atoms x
# 50 000 frames would be 120 GB of float64 next to 60 GB of input).  Only for device-resident input.
STREAM_BYTES = 8 << 30
STREAM_BLOCK_BYTES = 2 << 30


def _side_streams(device):
    """(copy stream, ingest stream) of a device: transfers and the stage-1 kernels that consume them run beside the
    caller's stream."""
    key = (device.type, device.index)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = (torch.cuda.Stream(device), torch.cuda.Stream(device))
    return _SIDE_STREAMS[key]


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
        self._host = None      # host copy still to be staged (None: the coordinates are resident)
        self._staged = None    # per atom: its row has been queued for the device
        if not isinstance(coords, torch.Tensor):
            coords = torch.from_numpy(np.ascontiguousarray(coords))
        if coords.dtype not in (torch.float32, torch.float64):
            coords = coords.to(torch.float64)
        if coords.dim() != 3 or coords.shape[2] != 3:
            raise ValueError('coordinates must have the shape [site, frame, 3]')
        if coords.device.type == 'cpu':
            # Host input: rows are copied atom block by atom block on a side stream, in the order the parser needs
            # them, so that the stage-1 kernels overlap the transfer (see Parser._build_displacement_staged).
            # The staged ingest lives on its own two streams and never waits for the caller's stream, so the transfer
            # of one trajectory overlaps whatever the caller's stream is still doing (the regression of the previous one).
            self._host = coords.contiguous()
            self._staged = np.zeros(self._host.shape[0], dtype=bool)
            self._side = _side_streams(device)
            with torch.cuda.stream(self._side[0]):  # from the copy stream's pool: no cross-stream reuse hazard
                self.coords = torch.empty(self._host.shape, dtype=self._host.dtype, device=device)
            self.coords.record_stream(self._side[1])
            self.coords.record_stream(engine.current_stream(device))
        else:
            self.coords = coords.to(device).contiguous()
        self.latt = self.latt_inv = None
        self.latt_stride = 9
        self._latt_event = None
        if mode == MODE_FRACTIONAL and latt is not None:
            latt_host = latt if isinstance(latt, np.ndarray) else None
            # with a staged ingest the lattices are prepared on the ingest stream, like everything else of stage 1
            work = self._side[1] if self._host is not None else engine.current_stream(device)
            # (entering a stream context costs ~15 us of host time: not when the work stream is the current one)
            with (torch.cuda.stream(work) if self._host is not None else contextlib.nullcontext()):
                cached = None
                if latt_host is not None and latt_host.nbytes <= (64 << 10):
                    # small lattices (the constant cell above all) and their inverses are kept on the device by content
                    key = (np.ascontiguousarray(latt_host, dtype=np.float64).tobytes(), device.index)
                    cached = _LATTICES.get(key)
                if cached is not None:
                    latt, inv, uploaded = cached
                    if uploaded is not None:
                        if uploaded.query():
                            _LATTICES[key] = (latt, inv, None)  # landed: later hits need not ask again
                        else:
                            work.wait_event(uploaded)
                    latt.record_stream(work)  # (an evicted entry must outlive its readers on other streams)
                    inv.record_stream(work)
                else:
                    latt = engine.to_device(latt, device, torch.float64).reshape(-1, 3, 3)
                    inv = None
                if latt.shape[0] == 1:
                    self.latt_stride = 0
                elif latt.shape[0] != self.coords.shape[1]:
                    raise ValueError('one lattice per frame is required')
                elif latt_host is not None and np.all(latt_host == latt_host[0]):
                    self.latt_stride = 0  # constant cell: every frame reads latt[0]
                self.latt = latt
                self.latt_inv = inv if inv is not None else engine.invert_lattice(latt)
                if cached is None and latt_host is not None and latt_host.nbytes <= (64 << 10):
                    uploaded = torch.cuda.Event()
                    uploaded.record(work)
                    if len(_LATTICES) >= 64:
                        _LATTICES.pop(next(iter(_LATTICES)))
                    _LATTICES[key] = (self.latt, self.latt_inv, uploaded)
                if self._host is not None:
                    self._latt_event = torch.cuda.Event()
                    self._latt_event.record(work)
            if self._host is not None:
                self.latt.record_stream(engine.current_stream(device))
                self.latt_inv.record_stream(engine.current_stream(device))

    @property
    def device(self):
        return self.coords.device

    @property
    def resident(self):
        return self._host is None

    def stage(self, rows):
        """Queue the host->device copy of the atom rows `rows` (host integers) on the copy stream, contiguous runs as
        single copies; returns an event that fires when they have landed (None: nothing left to copy)."""
        if self._host is None:
            return None
        rows = np.unique(np.asarray(rows, dtype=np.int64))
        rows = rows[~self._staged[rows]]
        if rows.size == 0:
            return None
        copy = self._side[0]
        breaks = np.flatnonzero(np.diff(rows) != 1) + 1
        with torch.cuda.stream(copy):
            for run in np.split(rows, breaks):
                a, b = int(run[0]), int(run[-1]) + 1
                self.coords[a:b].copy_(self._host[a:b], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy)
        self._staged[rows] = True
        if self._staged.all():
            self._host_keepalive = self._host  # the copies may still be in flight
            self._host = None
        return ev

    def make_resident(self, rows=None):
        """Stage `rows` (default: every atom) and make the current stream wait for the copy."""
        if self._latt_event is not None:
            engine.current_stream(self.device).wait_event(self._latt_event)
            self._latt_event = None
        if self._host is None:
            return
        ev = self.stage(np.arange(self.coords.shape[0]) if rows is None else rows)
        if ev is not None:
            engine.current_stream(self.device).wait_event(ev)

    @property
    def n_atoms(self):
        return self.coords.shape[0]

    @property
    def n_frames(self):
        """Number of displacement frames (the reference's drift_corrected.shape[1])."""
        return self.coords.shape[1] - (1 if self.mode == MODE_FRACTIONAL else 0)

    @classmethod
    def from_frames(cls, frames, latt=None, mode=MODE_FRACTIONAL, device=None, duplicate_first=True, cartesian_cells=None):
        """
        A trajectory from FRAME-major coordinates `[frame, site, 3]` (how MD codes write them and how the front-ends
        collect them, kinisi/parser.py:309-325, 458-473).  The transposition to the atom-major layout of the kernels
        runs on the device (kb2_frames_to_atoms); host input is transferred in blocks of frames on the copy stream
        and transposed on the ingest stream behind the transfer.  `duplicate_first` inserts the copy of the first
        frame the reference's front-ends prepend (parser.py:323-324).  `cartesian_cells` ([frame, 3, 3], or [1, 3, 3] for a
        constant cell; without the duplicated frame): the frames hold Cartesian positions and are multiplied by the
        inverse cell of their frame on the way (the MDAnalysis front-end, parser.py:630-635); the result is float64.
        """
        device = _default_device(device)
        if not isinstance(frames, torch.Tensor):
            frames = torch.from_numpy(np.ascontiguousarray(frames))
        if frames.dtype not in (torch.float32, torch.float64):
            frames = frames.to(torch.float64)
        if frames.dim() != 3 or frames.shape[2] != 3:
            raise ValueError('frames must have the shape [frame, site, 3]')
        frames = frames.contiguous()
        n_f, n_a = int(frames.shape[0]), int(frames.shape[1])
        off = 1 if duplicate_first else 0
        copy, ingest = _side_streams(device)
        main = engine.current_stream(device)
        out_dtype = frames.dtype if cartesian_cells is None else torch.float64
        inv = None
        if cartesian_cells is not None:
            work = ingest if frames.device.type == 'cpu' else main
            with torch.cuda.stream(work):
                cells = engine.to_device(np.ascontiguousarray(cartesian_cells, dtype=np.float64).reshape(-1, 3, 3), device,
                                         torch.float64)
                if cells.shape[0] not in (1, n_f):
                    raise ValueError('one cell per frame is required')
                inv = engine.invert_lattice(cells)

        def inv_of(f0, f1):
            return None if inv is None else (inv if inv.shape[0] == 1 else inv[f0:f1])

        if frames.device.type == 'cpu':
            with torch.cuda.stream(ingest):
                coords = torch.empty((n_a, n_f + off, 3), dtype=out_dtype, device=device)
            per = max(1, int(STAGE_BYTES // max(1, n_a * 3 * frames.element_size())))
            with torch.cuda.stream(copy):
                stage = [torch.empty((min(per, n_f), n_a, 3), dtype=frames.dtype, device=device) for _ in range(2)]
            free = [None, None]  # event: the transpose that read staging buffer i has been queued and finished
            for i, f0 in enumerate(range(0, n_f, per)):
                f1 = min(n_f, f0 + per)
                buf = stage[i & 1][:f1 - f0]
                with torch.cuda.stream(copy):
                    if free[i & 1] is not None:
                        copy.wait_event(free[i & 1])
                    buf.copy_(frames[f0:f1], non_blocking=True)
                    landed = torch.cuda.Event()
                    landed.record(copy)
                with torch.cuda.stream(ingest):
                    ingest.wait_event(landed)
                    engine.frames_to_atoms(buf, coords, off + f0, duplicate_first and f0 == 0, inv_of(f0, f1))
                    free[i & 1] = torch.cuda.Event()
                    free[i & 1].record(ingest)
            for t in stage:
                t.record_stream(ingest)
            coords.record_stream(main)
            main.wait_stream(ingest)
            keep = frames  # the transfer reads the caller's memory until the streams have drained
        else:
            frames = frames.to(device)
            coords = torch.empty((n_a, n_f + off, 3), dtype=out_dtype, device=device)
            for f0 in range(0, n_f, 65535 * 32):
                f1 = min(n_f, f0 + 65535 * 32)
                engine.frames_to_atoms(frames[f0:f1], coords, off + f0, duplicate_first and f0 == 0, inv_of(f0, f1))
            keep = None
        traj = cls(coords, latt, mode, device)
        traj._frames_keepalive = keep
        return traj

    def frame_sums(self, idx, weights=None, join=None):
        """Per-frame (weighted) displacement sums over atoms idx (an int32 device tensor); `join`: see engine.frame_sums."""
        self.make_resident()
        pitch = engine.pitch_for(self.n_frames)
        return engine.frame_sums(self.coords, self.mode, self.latt, self.latt_inv, self.latt_stride, idx, weights, pitch,
                                 join=join)

    def displacement(self, idx, drift=None):
        """dc [len(idx), 3, pitch] float64, component-major.  `drift` is None or (frame_sums, n_framework): the
        per-frame sums over the framework atoms, whose mean is removed before the cumulative sum."""
        self.make_resident()
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

    def __init__(self, dc, n_frames, lags, n_atoms_global=None, stride=1, ready=None, distributed=False, source=None):
        self._dc = dc  # [n_local_atoms, 3, pitch] float64 CUDA; None while the sequence is STREAMED from `source`
        # streamed: (trajectory, int32 device indices of the selected atoms, drift or None): the displacements of a block
        # of atoms are rebuilt from the coordinates whenever a consumer asks (`blocks()`), the whole array only for `dc`
        self._source = source
        # True: dc holds this rank's shard of the atoms; sums over atoms are all-reduced (set by a distributed Parser)
        self.distributed = bool(distributed) and dist.active()
        # staged ingest: [(row0, row1, event)] -- rows [row0, row1) of dc are valid once the event has fired
        self._ready = list(ready) if ready else []
        self.n_frames = int(n_frames)
        self.lags = np.asarray(lags, dtype=np.int64)
        self.n_atoms = dc.shape[0] if dc is not None else int(source[1].numel())
        self.n_atoms_global = int(n_atoms_global if n_atoms_global is not None else self.n_atoms)

    @property
    def streamed(self):
        return self._dc is None

    def blocks(self):
        """Yields `(row0, row1, dc_rows)`: the displacement array in blocks of atoms, each valid on the current stream
        when it is yielded.  Resident: the blocks of a staged ingest in the order they land (one block otherwise).
        Streamed: stage 1 runs for one block at a time into ONE buffer -- a block must be consumed (in stream order)
        before the next one is asked for."""
        if self._dc is not None:
            stream = engine.current_stream(self._dc.device)
            for r0, r1, ev in self.row_blocks():
                if ev is not None:
                    stream.wait_event(ev)
                yield r0, r1, self._dc[r0:r1]
            return
        traj, idx, drift = self._source
        pitch = engine.pitch_for(self.n_frames)
        per = max(1, min(self.n_atoms, int(STREAM_BLOCK_BYTES // (3 * pitch * 8))))
        buf = torch.empty((per, 3, pitch), dtype=torch.float64, device=idx.device)
        sums, scale = (None, 0.0) if drift is None else (drift[0], 1.0 / drift[1])
        traj.make_resident()
        for r0 in range(0, self.n_atoms, per):
            r1 = min(self.n_atoms, r0 + per)
            engine.unwrap_cumsum(traj.coords, traj.mode, traj.latt, traj.latt_inv, traj.latt_stride, idx[r0:r1], sums, scale,
                                 pitch, out=buf[:r1 - r0])
            yield r0, r1, buf[:r1 - r0]

    @property
    def dc(self):
        """The whole array, valid on the current stream (waits for a staged ingest to finish; a streamed sequence
        builds and keeps it: indexing, resampling and blocking need the array)."""
        if self._dc is None:
            traj, idx, drift = self._source
            self._dc = traj.displacement(idx, drift)
            self._source = None
        if self._ready:
            stream = engine.current_stream(self._dc.device)
            for _, _, ev in self._ready:
                stream.wait_event(ev)
            self._ready = []
        return self._dc

    def row_blocks(self):
        """[(row0, row1, event or None)]: blocks of atoms in the order they become valid.  A consumer that reduces
        over atoms (the moment kernels) can work block by block behind the transfer."""
        if not self._ready:
            return [(0, self.n_atoms, None)]
        return list(self._ready)

    @property
    def device(self):
        return self._dc.device if self._dc is not None else self._source[1].device

    @property
    def dc_unsynchronised(self):
        return self._dc

    def __len__(self):
        return len(self.lags)

    def shape_of(self, i):
        return (self.n_atoms, self.n_frames - int(self.lags[i]) + 1, 3)

    @property
    def shapes(self):
        return [self.shape_of(i) for i in range(len(self))]

    def n_obs_all(self):
        """Number of observations of every entry (shape_of(i)[1]) as one array."""
        return self.n_frames - self.lags + 1

    def __getitem__(self, i):
        if isinstance(i, slice):
            return LazyDisplacements(self._dc, self.n_frames, self.lags[i], self.n_atoms_global, ready=self._ready,
                                     distributed=self.distributed, source=self._source)
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
        return LazyDisplacements(dc, self.n_frames, self.lags, self.n_atoms_global + sum(o.n_atoms_global for o in others),
                                 distributed=self.distributed)


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
    :param global_counts: Optional `(selected atoms, framework atoms)` summed over all ranks.  When given, the
        collective (and the host synchronisation) that would count them is skipped.
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
                 distributed: bool = False,
                 global_counts: Tuple[int, int] = None):
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

        n_sel_local = len(indices)
        # global atom counts (selected, framework): one small collective
        if self._distributed and global_counts is not None:
            n_sel, self._n_fw_global = int(global_counts[0]), int(global_counts[1])
        elif self._distributed:
            n_sel, self._n_fw_global = (int(round(v)) for v in dist.allreduce_sum_scalars(
                [n_sel_local, len(drift_indices)], dev))
        else:
            n_sel, self._n_fw_global = n_sel_local, len(drift_indices)
        self._n_atoms_global = n_sel

        # memory guard (kinisi/parser.py:190-200), applied to what this build keeps resident: the cumulative displacements
        # of the selected atoms, or one block of them when they are streamed
        dc_bytes = n_sel_local * 3 * engine.pitch_for(n_frames) * 8
        self._streamed = bool(traj.resident and sampling == 'multi-origin' and dc_bytes > STREAM_BYTES)
        need_gb = (min(dc_bytes, STREAM_BLOCK_BYTES) if self._streamed else dc_bytes) * 1e-9
        if need_gb > self.memory_limit:
            raise MemoryError(f"The memory limit of this job is {self.memory_limit:.1e} GB but the "
                              f"displacement values will use {need_gb:.1e} GB. Please either increase "
                              "the memory_limit parameter or descrease the sampling rate (see "
                              "https://kinisi.readthedocs.io/en/latest/memory_limit.html).")

        # drift correction (kinisi/parser.py:131-148) fused with the displacement construction
        self._dc_ready = None
        if not traj.resident and n_sel_local > 0:
            self._build_displacement_staged(traj, indices, drift_indices)
        else:
            # one upload for both index lists
            both = engine.const_to_device(np.concatenate([np.asarray(indices, dtype=np.int64).reshape(-1),
                                                          np.asarray(drift_indices, dtype=np.int64).reshape(-1)]), dev, np.int32)
            idx, fw_idx = both[:n_sel_local], both[n_sel_local:]
            self._drift = self._framework_mean(traj, drift_indices, fw_idx)
            self._sel_idx_dev = idx
            if self._streamed:
                self._dc_sel = None
            elif n_sel_local > 0:
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
        self.delta_t, self.disp_3d, self._n_o = self.get_disps(self.time_intervals, self._dc_sel_staged, progress)

    def _build_displacement_staged(self, traj, indices, drift_indices):
        """Stage 1 behind the host->device transfer.  The rows are copied in the order they are needed -- framework
        atoms first (their per-frame sums are the drift), then the selected atoms, block by block -- on a copy stream;
        an ingest stream runs the frame-sum / unwrap kernels of a block as soon as it has landed and records one event
        per block of the displacement array, so a consumer on the caller's stream (the moment kernels) can follow
        block by block.  Atoms that are neither selected nor framework are never transferred."""
        dev = traj.device
        main = engine.current_stream(dev)
        _, ingest = _side_streams(dev)
        n_frames, pitch = traj.n_frames, engine.pitch_for(traj.n_frames)
        row_bytes = traj.coords.shape[1] * 3 * traj.coords.element_size()
        per_block = max(1, int(STAGE_BYTES // row_bytes))
        sel = np.asarray(indices, dtype=np.int64)
        fw = np.asarray(drift_indices, dtype=np.int64)
        n_fw_local = len(fw)
        n_fw = self._n_fw_global
        fw_sorted = np.sort(fw)
        with torch.cuda.stream(ingest):
            idx_dev = engine.const_to_device(np.concatenate([fw_sorted, sel]), dev, np.int32)  # one upload for every block
            fw_dev, sel_dev = idx_dev[:n_fw_local], idx_dev[n_fw_local:]
            sums = None
            if n_fw > 0:
                for b0 in range(0, n_fw_local, per_block):
                    block = fw_sorted[b0:b0 + per_block]
                    ev = traj.stage(block)
                    if ev is not None:
                        ingest.wait_event(ev)
                    part = engine.frame_sums(traj.coords, traj.mode, traj.latt, traj.latt_inv, traj.latt_stride,
                                             fw_dev[b0:b0 + len(block)], None, pitch)
                    sums = part if sums is None else sums.add_(part)
                if sums is None:
                    sums = torch.zeros((3, pitch), dtype=torch.float64, device=dev)
                if self._distributed:
                    dist.allreduce_sum_(sums, 'frames')
            self._drift = None if n_fw == 0 else (sums, float(n_fw))
            step_sums, scale = (None, 0.0) if self._drift is None else (sums, 1.0 / n_fw)
            dc = torch.empty((len(sel), 3, pitch), dtype=torch.float64, device=dev)
            dc.record_stream(main)  # allocated for the ingest stream, consumed on the caller's
            if sums is not None:
                sums.record_stream(main)
            ready = []
            for b0 in range(0, len(sel), per_block):
                block = sel[b0:b0 + per_block]
                ev = traj.stage(block)
                if ev is not None:
                    ingest.wait_event(ev)
                engine.unwrap_cumsum(traj.coords, traj.mode, traj.latt, traj.latt_inv, traj.latt_stride,
                                     sel_dev[b0:b0 + len(block)], step_sums, scale, pitch, out=dc[b0:b0 + len(block)])
                done = torch.cuda.Event()
                done.record(ingest)
                ready.append((b0, b0 + len(block), done))
        self._dc_sel_staged = dc
        self._dc_ready = ready

    @property
    def _dc_sel(self):
        """The drift-corrected displacement of the selected atoms, valid on the current stream."""
        if self._dc_sel_staged is None and getattr(self, '_streamed', False):
            return self.disp_3d.dc  # a streamed parser builds (and keeps) the array only on request
        if self._dc_ready:
            stream = engine.current_stream(self._dc_sel_staged.device)
            for _, _, ev in self._dc_ready:
                stream.wait_event(ev)
            self._dc_ready = None
        return self._dc_sel_staged

    @_dc_sel.setter
    def _dc_sel(self, value):
        self._dc_sel_staged = value
        self._dc_ready = None

    def _framework_mean(self, traj, drift_indices, fw=None):
        """(per-frame sums over the framework atoms [3, pitch] on the device, number of framework atoms), or None.
        The displacement kernel subtracts sums / n before its cumulative sum (kinisi/parser.py:143-145)."""
        n_fw_local = len(drift_indices)
        n_fw = self._n_fw_global
        if n_fw == 0:
            return None
        # atoms sharded over ranks: the kernel that makes the local sums ends with the sum over ranks (one fused kernel over
        # peer memory, csrc/k1_unwrap.cu frame_join_epilogue); a rank without local framework atoms joins with zeros
        join = dist.frame_sums_channel(3, engine.pitch_for(traj.n_frames), traj.device) if self._distributed else None
        if n_fw_local > 0:
            if fw is None:
                fw = engine.small_to_device(drift_indices, traj.device, np.int32)
            sums = traj.frame_sums(fw, join=join)
            if join is not None:
                dist.check_peer_status()
                return sums, float(n_fw)
        else:
            sums = torch.zeros((3, engine.pitch_for(traj.n_frames)), dtype=torch.float64, device=traj.device)
        if self._distributed:
            dist.allreduce_sum_(sums, 'frames')
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
    def get_disp_from_frames(frames, latt, progress: bool = True, device=None, cartesian: bool = False) -> Trajectory:
        """As :py:meth:`get_disp` for FRAME-major fractional coordinates `[frame, site, 3]` WITHOUT the duplicated first
        frame (it is inserted on the device): see :py:meth:`Trajectory.from_frames`.  `cartesian`: the frames hold
        Cartesian positions, converted with the inverse of their frame's cell on the device."""
        if isinstance(latt, (list, tuple)):
            latt = np.array(latt)
        latt_np = latt.detach().cpu().numpy() if isinstance(latt, torch.Tensor) else np.asarray(latt)
        latt_np = latt_np.reshape(-1, 3, 3)
        off_diags = np.array([[False, True, True], [True, False, True], [True, True, False]])
        if np.any(latt_np[:, off_diags] != 0):
            warnings.warn(
                'Converting triclinic cell to orthorhombic this may have unexpected results. '
                'Triclinic a, b, c are not equilivalent to orthorhombic x, y, z.', UserWarning)
        cells = latt_np if cartesian else None
        if cells is not None and cells.shape[0] > 1 and np.all(cells == cells[0]):
            cells = cells[:1]
        if latt_np.shape[0] > 1:
            latt_np = np.concatenate([latt_np[:1], latt_np], axis=0)  # the duplicated first frame has its lattice too
        return Trajectory.from_frames(frames, latt_np, MODE_FRACTIONAL, device, duplicate_first=True, cartesian_cells=cells)

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
        """The integer time-interval grid (kinisi/parser.py:150-168).  Integer work: stays on the host (and is kept per
        set of arguments: a loop over trajectories of one shape asks for the same grid every time)."""
        unit = self.step_skip * self.time_step
        min_dt = int(self.min_dt / unit)
        max_dt = int(self.max_dt / unit)
        if min_dt == 0:
            min_dt = 1
        key = (min_dt, max_dt, int(n_steps), spacing)
        hit = _GRIDS.get(key)
        if hit is not None:
            return hit.copy()
        if spacing == 'linear':
            grid = np.linspace(min_dt, max_dt, n_steps, dtype=int)
        elif spacing == 'logarithmic':
            grid = np.unique(np.geomspace(min_dt, max_dt, n_steps, dtype=int))
        else:
            raise ValueError("Only linear or logarithmic spacing is allowed.")
        if len(_GRIDS) >= 64:
            _GRIDS.pop(next(iter(_GRIDS)))
        _GRIDS[key] = grid
        return grid.copy()

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
        ks = np.asarray(time_intervals, dtype=np.int64)
        n_obs = np.ones_like(ks) if self.sampling == 'single-origin' else n_frames - ks + 1
        # parser.py:204-205, 214-215: an interval is kept if atoms x len(range(0, n_obs, k)) > 1
        kept = np.flatnonzero(n_atoms * np.maximum(0, -(-n_obs // ks)) > 1)
        n_samples = n_atoms * time_intervals[-1] / time_intervals[kept]  # parser.py:221
        if self.sampling == 'single-origin':
            # parser.py:203: entry i is the single frame dc[:, i] (indexed by grid position, not by lag)
            disp_3d = SingleOriginDisplacements(drift_corrected, n_frames, kept, n_atoms, ready=getattr(self, '_dc_ready', None),
                                                distributed=getattr(self, '_distributed', False))
        else:
            source = (self._traj, self._sel_idx_dev, self._drift) if getattr(self, '_streamed', False) else None
            disp_3d = LazyDisplacements(drift_corrected, n_frames, time_intervals[kept], n_atoms, ready=getattr(self, '_dc_ready', None),
                                        distributed=getattr(self, '_distributed', False), source=source)
        return delta_t, disp_3d, np.array(n_samples, dtype=float)


class SingleOriginDisplacements(LazyDisplacements):
    """`sampling='single-origin'` (kinisi/parser.py:202-206): entry i is `dc[:, i:i+1]`."""

    def __init__(self, dc, n_frames, frames, n_atoms_global=None, ready=None, distributed=False):
        super().__init__(dc, n_frames, np.ones(len(frames), dtype=np.int64), n_atoms_global, ready=ready,
                         distributed=distributed)
        self.frames = np.asarray(frames, dtype=np.int64)

    def shape_of(self, i):
        return (self.n_atoms, 1, 3)

    def n_obs_all(self):
        return np.ones(len(self.lags), dtype=np.int64)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return SingleOriginDisplacements(self._dc, self.n_frames, self.frames[i], self.n_atoms_global, ready=self._ready,
                                             distributed=self.distributed)
        f = int(self.frames[i])
        return self.dc[:, :, f:f + 1].permute(0, 2, 1).contiguous().cpu().numpy()

    def device_entry(self, i):
        f = int(self.frames[i])
        return self.dc[:, :, f:f + 1].permute(0, 2, 1).contiguous()

    def stacked_with(self, others):
        """`Analyzer._stack_trajectories` (kinisi/analyzer.py:274-290) stacks entry i = `dc[:, i:i+1]` of every
        trajectory along the atom axis: the stacked sequence is again single-origin, over the same frames."""
        dc = torch.cat([self.dc] + [o.dc for o in others], dim=0)
        return SingleOriginDisplacements(dc, self.n_frames, self.frames,
                                         self.n_atoms_global + sum(o.n_atoms_global for o in others),
                                         distributed=self.distributed)


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


def _walk_frames(frames, get_frac, get_cell, sub_sample_traj):
    """The same walk collecting FRAME-major arrays `[frame, site, 3]`, `[frame, 3, 3]` (no duplicated first frame): the
    frames are appended as they come and transposed on the device."""
    structure = None
    coords, latt = [], []
    for struct in frames[::sub_sample_traj]:
        if structure is None:
            structure = struct
        coords.append(np.asarray(get_frac(struct), dtype=np.float64))
        latt.append(np.asarray(get_cell(struct), dtype=np.float64))
    return structure, np.stack(coords, axis=0), np.stack(latt, axis=0)


def _get_molecules(structure, traj, indices, masses, framework_indices):
    """
    Centres of groups of atoms by the pseudo-centre-of-mass approach (kinisi/parser.py:674-738, arXiv:2501.14578), on
    the device (kb2_molecule_centres).  `indices` are 1-based lists of equal length, one per molecule; `masses` has one
    entry per member (None: centre of geometry); `framework_indices` 1-based or None (every atom outside the molecules).

    :return: a Trajectory whose first rows are the centres followed by the framework atoms, and the index lists
        `(centres, framework)` into it.
    """
    try:
        members = np.array(indices) - 1
        if members.ndim != 2:
            raise ValueError
    except Exception:
        raise ValueError('Molecules must be of same length')
    n_molecules = members.shape[0]
    if isinstance(framework_indices, (list, tuple)):
        drift_indices = np.array(framework_indices, dtype=np.int64) - 1
    else:
        inside = set(int(i) for i in members.reshape(-1))
        drift_indices = np.array([i for i, _ in enumerate(structure) if i not in inside], dtype=np.int64)
    weights = None
    if masses is not None:
        if len(masses) != members.shape[-1]:
            raise ValueError('Masses must be the same length as a molecule')
        weights = engine.small_to_device(np.asarray(masses, dtype=np.float64), traj.device, np.float64)
    traj.make_resident()
    centres = engine.molecule_centres(traj.coords, engine.small_to_device(members, traj.device, np.int32), weights)
    rows = [centres]
    if len(drift_indices):
        sel = engine.small_to_device(drift_indices, traj.device, np.int64)
        rows.append(traj.coords.index_select(0, sel).to(torch.float64))
    new = Trajectory(torch.cat(rows, dim=0), None, traj.mode, traj.device)
    new.latt, new.latt_inv, new.latt_stride, new._latt_event = traj.latt, traj.latt_inv, traj.latt_stride, traj._latt_event
    return new, (list(range(n_molecules)), list(range(n_molecules, n_molecules + len(drift_indices))))


def _get_framework(structure, indices, framework_indices):
    """kinisi/parser.py:741-764 (1-based `framework_indices`, as the reference subtracts one)."""
    if isinstance(framework_indices, (list, tuple)):
        return indices, list(np.array(framework_indices, dtype=int) - 1)
    chosen = set(int(i) for i in indices)
    return indices, [i for i, _ in enumerate(structure) if i not in chosen]


def _front_end_trajectory(parser, structure, frames, latt, specie, specie_indices, masses, framework_indices, progress,
                          device, cartesian=False):
    """Shared body of the front-end constructors (kinisi/parser.py:266-293, 406-435): index selection, molecule
    centres, `coords_check`, and the device trajectory built from the frame-major walk."""
    traj = Parser.get_disp_from_frames(frames, latt, progress=progress, device=device, cartesian=cartesian)
    first = (np.dot(frames[0], np.linalg.inv(latt[0])) if cartesian else frames[0])[:, None, :]
    if specie is not None:
        indices = parser.get_indices(structure, specie, framework_indices)
        parser.coords_check = first
    elif isinstance(specie_indices, (list, tuple)):
        if isinstance(specie_indices[0], (list, tuple)):
            traj, indices = _get_molecules(structure, traj, specie_indices, masses, framework_indices)
            parser.coords_check = traj.coords[:, 0:1].cpu().numpy()
        else:
            indices = _get_framework(structure, specie_indices, framework_indices)
            parser.coords_check = first
    else:
        raise TypeError('Unrecognized type for specie or specie_indices')
    return traj, indices


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
    :param masses: Optional, masses of the members of a molecule when `specie_indices` is a list of 1-based lists
        (centres by the pseudo-centre-of-mass approach, kinisi/parser.py:674-738).
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
        structure, frames, latt = _walk_frames(atoms, lambda s: s.get_scaled_positions(), lambda s: s.cell[:],
                                               sub_sample_traj)
        traj, indices = _front_end_trajectory(self, structure, frames, latt, specie, specie_indices, masses,
                                              framework_indices, progress, device)
        # note: like the reference (parser.py:291) sub_sample_traj is NOT folded into step_skip here
        super().__init__(traj, indices[0], indices[1], time_step, step_skip, min_dt, max_dt, n_steps, spacing, sampling,
                         memory_limit, progress, device=device)
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
        structure, frames, latt = _walk_frames(structures, lambda s: s.frac_coords, lambda s: s.lattice.matrix,
                                               sub_sample_traj)
        traj, indices = _front_end_trajectory(self, structure, frames, latt, specie, specie_indices, masses,
                                              framework_indices, progress, device)
        super().__init__(traj, indices[0], indices[1], time_step, step_skip * sub_sample_traj, min_dt, max_dt, n_steps,
                         spacing, sampling, memory_limit, progress, device=device)
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


class MDAnalysisParser(Parser):
    """
    A parser that consumes an MDAnalysis.Universe object (kinisi/parser.py:511-672).  Any object exposing
    `.trajectory[::n]` yielding timesteps with `.volume` and `.triclinic_dimensions`, and `.atoms[::n]` with `.positions`
    (Cartesian, of the current timestep) and iteration over sites with `.type` works.  The positions are collected as
    they are read and the per-frame product with the inverse cell (parser.py:630-634) runs on the device.

    :param universe: The MDAnalysis object of interest.
    :param specie: Specie to calculate diffusivity for as a String, e.g. 'Li'.
    :param time_step: Time step, in picoseconds, between steps in trajectory.
    :param step_skip: Sampling frequency of the trajectory.
    :param sub_sample_atoms: The sampling rate to sample the atoms in the system. Optional, defaults to 1.
    :param sub_sample_traj: Multiple of the time_step to sub sample at (folded into step_skip, parser.py:593).
    :param specie_indices, masses, framework_indices: as for :py:class:`ASEParser`.
    (min_dt, max_dt, n_steps, spacing, sampling, memory_limit, progress as for :py:class:`Parser`.)
    """

    def __init__(self,
                 universe,
                 specie: str,
                 time_step: float,
                 step_skip: int,
                 sub_sample_atoms: int = 1,
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
        if sub_sample_atoms != 1 and specie_indices is not None:
            raise ValueError(
                'sub_sample_atom cannot be used with specie_indices. Please specify only inidices you wish to sample.')
        structure, frames, cells, volume = self._walk_universe(universe, sub_sample_atoms, sub_sample_traj)
        traj, indices = _front_end_trajectory(self, structure, frames, cells, specie, specie_indices, masses,
                                              framework_indices, progress, device, cartesian=True)
        super().__init__(traj, indices[0], indices[1], time_step, step_skip * sub_sample_traj, min_dt, max_dt, n_steps,
                         spacing, sampling, memory_limit, progress, device=device)
        self._volume = volume

    @staticmethod
    def _walk_universe(universe, sub_sample_atoms, sub_sample_traj):
        """The walk of kinisi/parser.py:620-636 collecting FRAME-major Cartesian positions `[frame, site, 3]` (in the
        universe's own element type, float32 for MDAnalysis) and the cells `[frame, 3, 3]`."""
        structure, volume = None, 0
        positions, cells = [], []
        for timestep in universe.trajectory[::sub_sample_traj]:
            if structure is None:
                structure = universe.atoms[::sub_sample_atoms]
                volume = timestep.volume
            cells.append(np.array(timestep.triclinic_dimensions, dtype=np.float64))
            positions.append(np.array(universe.atoms[::sub_sample_atoms].positions))  # a copy: readers reuse the buffer
        return structure, np.stack(positions, axis=0), np.stack(cells, axis=0), volume

    @staticmethod
    def get_structure_coords_latt(universe, sub_sample_atoms: int = 1, sub_sample_traj: int = 1, progress: bool = True):
        """:return: initial structure, fractional coordinates as the reference's list of `[site, 1, 3]` arrays (first
        frame duplicated), the list of cells, and the cell volume (kinisi/parser.py:600-638; host arithmetic)."""
        structure, positions, cells, volume = MDAnalysisParser._walk_universe(universe, sub_sample_atoms, sub_sample_traj)
        coords = [np.dot(p, np.linalg.inv(m))[:, None] for p, m in zip(positions, cells)]
        latt = list(cells)
        coords.insert(0, coords[0])
        latt.insert(0, latt[0])
        return structure, coords, latt, volume

    @staticmethod
    def get_indices(structure, specie, framework_indices) -> Tuple[list, list]:
        """Framework and non-framework indices by atom type (kinisi/parser.py:640-672)."""
        indices, drift_indices = [], []
        if not isinstance(specie, list):
            specie = [specie]
        for i, site in enumerate(structure):
            if site.type in specie:
                indices.append(i)
            else:
                drift_indices.append(i)
        if len(indices) == 0:
            raise ValueError("There are no species selected to calculate the mean-squared displacement of.")
        if isinstance(framework_indices, (list, tuple)):
            drift_indices = framework_indices
        return indices, drift_indices
