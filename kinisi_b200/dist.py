"""Multi-GPU plumbing: one process per GPU, atoms sharded across ranks, per-dt moment sums and the
per-frame framework / collective sums joined by a sum over ranks -- one kernel of our own per join over
peer memory (NVLink / NVSwitch, csrc/k7_peer.cu) on the GPU box, NCCL where the windows cannot be mapped,
gloo in the CPU tests.  Everything is a no-op when torch.distributed is not initialised."""
import ctypes
import os

import torch
import torch.distributed as td


_CHANNELS = {}
_STATE = {'group': None, 'active': False}


def active():
    """True inside an initialised process group of more than one rank (cached per default group: this is asked several
    times per analysis)."""
    if not (td.is_available() and td.is_initialized()):
        _STATE['group'] = None
        return False
    g = td.group.WORLD
    if _STATE['group'] is not g:
        _STATE['group'], _STATE['active'] = g, td.get_world_size() > 1
        _CHANNELS.clear()
        for comm in _DIRECT.values():
            if comm is not None:
                comm.close()
        _DIRECT.clear()
        for ch in _PEER.values():
            if ch:
                ch.close(collective=False)  # another process group: the old one cannot be asked to wait
        _PEER.clear()
    return _STATE['active']


def world_size():
    return td.get_world_size() if active() else 1


def rank():
    return td.get_rank() if active() else 0



def _channel(name):
    """A process group of its own per kind of collective.  A communicator runs its collectives in issue order on one
    internal stream: with several analyses in flight, the frame-sum all-reduce of the next trajectory would otherwise
    queue behind the moment-sum all-reduce of the current one, which waits for its moment kernel -- stage 1 of the next
    trajectory could not start before stage 2 of this one had finished.  Created on first use, in program order (the
    same on every rank)."""
    g = _CHANNELS.get(name)
    if g is None:
        g = _CHANNELS[name] = td.group.WORLD if name is None else td.new_group()
    return g


class _NcclUniqueId(ctypes.Structure):
    _fields_ = [('internal', ctypes.c_byte * 128)]


class _DirectComm:
    """One NCCL communicator driven through the library's C API, for the small float64 all-reduces of the path.

    `ProcessGroupNCCL` runs every collective on a stream of its own (two event hand-overs per call) behind a good deal of
    bookkeeping: 115-130 us of host time per call in the analysis loop, which made the host thread the bottleneck of a
    sharded run.  Here `ncclAllReduce` is enqueued on the CALLER's stream, about 10 us.  Calls on
    one communicator must run in the same order on every rank; they come from different streams (analyses in flight), so
    each call waits for the event of the previous one on this communicator.  Collectives on DIFFERENT communicators can run
    at the same time on different streams; that is safe only because every rank issues them in the same host program order
    (one analysis after the other, the same stages in each) -- a requirement of this module, as it is of NCCL itself.
    Used only where the peer-memory windows of `_PeerChannel` cannot be mapped."""
    _lib = None

    @classmethod
    def library(cls):
        if cls._lib is None:
            path = None
            with open('/proc/self/maps') as f:  # the copy of libnccl torch has already loaded
                for line in f:
                    if 'libnccl.so' in line:
                        path = line.split()[-1]
                        break
            if path is None:
                raise OSError('libnccl is not loaded in this process')
            lib = ctypes.CDLL(path)
            lib.ncclGetUniqueId.argtypes = [ctypes.POINTER(_NcclUniqueId)]
            lib.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, _NcclUniqueId, ctypes.c_int]
            lib.ncclAllReduce.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_void_p, ctypes.c_void_p]
            lib.ncclGetErrorString.restype = ctypes.c_char_p
            lib.ncclGetErrorString.argtypes = [ctypes.c_int]
            lib.ncclCommDestroy.argtypes = [ctypes.c_void_p]
            cls._lib = lib
        return cls._lib

    def __init__(self, device):
        lib = self.library()
        uid = _NcclUniqueId()
        if td.get_rank() == 0:
            self._check(lib.ncclGetUniqueId(ctypes.byref(uid)))
        raw = torch.frombuffer(bytearray(bytes(uid)), dtype=torch.uint8).to(device)
        td.broadcast(raw, 0)
        ctypes.memmove(ctypes.byref(uid), bytes(raw.cpu().numpy().tobytes()), 128)
        comm = ctypes.c_void_p()
        with torch.cuda.device(device):
            self._check(lib.ncclCommInitRank(ctypes.byref(comm), td.get_world_size(), uid, td.get_rank()))
        self._comm, self._last, self.device = comm, None, device

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError('NCCL: ' + self.library().ncclGetErrorString(rc).decode())

    def close(self):
        """Destroys the communicator (after the work enqueued on it has finished)."""
        if self._comm is not None and self._comm.value:
            torch.cuda.synchronize(self.device)
            self.library().ncclCommDestroy(self._comm)
        self._comm = None

    def allreduce_sum_f64(self, t):
        stream = torch.cuda.current_stream(t.device)
        if self._last is not None:
            stream.wait_event(self._last)
        self._check(self.library().ncclAllReduce(t.data_ptr(), t.data_ptr(), t.numel(), 8, 0, self._comm,
                                                 stream.cuda_stream))  # 8 = ncclFloat64, 0 = ncclSum
        self._last = torch.cuda.Event()
        self._last.record(stream)


class _PeerChannel:
    """The joins of one channel through peer memory (csrc/k7_peer.cu): every rank owns a window of device memory that
    its peers map through CUDA IPC, and one launch per rank stores the array into every window, raises a flag and adds
    the copies in rank order.  No NCCL kernel, no communicator stream, no ordering events between the streams of the
    analyses in flight (calls use the slots of the window round-robin and wait on the device for a slot to be released),
    and the sums are bit-identical on all ranks.  Creation is collective: every rank reaches the first join of a
    channel (and any join of a larger array than before) at the same point of the program."""

    def __init__(self, device, capacity):
        from . import _lib
        self._libmod, self.device = _lib, device
        lib = _lib.load()
        self.world, self.rank = td.get_world_size(), td.get_rank()
        self.capacity = int(capacity)
        self.seq = 0
        self._own, self._opened, self._counters = None, [], None
        handle = (ctypes.c_ubyte * 64)()
        own = ctypes.c_void_p()
        ok, err = 1, ''
        with torch.cuda.device(device):
            if lib.kb2_peer_window_create(self.world, self.capacity, ctypes.byref(own), handle) != 0:
                ok, err = 0, lib.kb2_last_error().decode('utf-8', 'replace')
        handles = [None] * self.world
        td.all_gather_object(handles, (ok, bytes(handle), os.getpid()))
        ptrs = (ctypes.c_void_p * self.world)()
        if all(h[0] for h in handles):
            self._own = own
            with torch.cuda.device(device):
                for q, (_, raw, _pid) in enumerate(handles):
                    if q == self.rank:
                        ptrs[q] = own.value
                        continue
                    p = ctypes.c_void_p()
                    buf = (ctypes.c_ubyte * 64).from_buffer_copy(raw)
                    if lib.kb2_peer_window_open(buf, ctypes.byref(p)) != 0:
                        ok, err = 0, lib.kb2_last_error().decode('utf-8', 'replace')
                        break
                    self._opened.append(p)
                    ptrs[q] = p.value
        else:
            ok = 0
            err = err or 'a peer could not create its window'
            if own.value:
                self._own = own
        # every rank must take the same road: one more exchange of the outcome
        outcome = [None] * self.world
        td.all_gather_object(outcome, ok)
        self.ok = all(outcome)
        self.error = err
        self._ptrs = ptrs
        if not self.ok:
            self.close()

    def close(self, collective=True):
        """Unmaps the peers' windows, then frees this rank's own -- after every rank has unmapped it (`collective`)."""
        lib = self._libmod.load()
        with torch.cuda.device(self.device):
            for p in self._opened:
                lib.kb2_peer_window_close(p)
            self._opened = []
            if collective and td.is_initialized():
                td.barrier()
            if self._own is not None and self._own.value:
                lib.kb2_peer_window_destroy(self._own)
            self._own = None

    def allreduce_sum_f64(self, t):
        """In-place sum over ranks.  A [rows][even row length] array of up to 128 x 1024 columns is cut by columns, one
        slice per 1024 columns of every row (the frame sums: the cut of the fused kernel, see `frame_sums_join`), anything
        else as a flat array."""
        lib = self._libmod.load()
        stream = torch._C._cuda_getCurrentRawStream(t.device.index)
        if t.dim() == 2 and t.shape[0] > 1 and lib.kb2_peer_rows_slices(t.shape[1]) > 0:
            rc = lib.kb2_peer_allreduce_rows_f64(t.data_ptr(), t.shape[0], t.shape[1], self._ptrs, self.world, self.rank,
                                                 self.capacity, self.seq, stream)
            self._libmod.check(rc, 'kb2_peer_allreduce_rows_f64')
        else:
            rc = lib.kb2_peer_allreduce_f64(t.data_ptr(), t.numel(), self._ptrs, self.world, self.rank, self.capacity,
                                            self.seq, stream)
            self._libmod.check(rc, 'kb2_peer_allreduce_f64')
        self.seq += 1
        from . import engine
        engine.LAUNCHES['n'] += 1
        check_peer_status()

    def next_join(self):
        """(windows, world, rank, capacity, seq) of the next join of this channel, for a kernel that does the join itself
        (kb2_moment_stats_joined); the call is counted."""
        seq = self.seq
        self.seq += 1
        return self._ptrs, self.world, self.rank, self.capacity, seq

    def frame_sums_join(self):
        """The arguments after `pitch` of kb2_frame_sums_joined for the next join of this channel (the call is counted):
        (windows, world, rank, capacity, seq, tile counters)."""
        if self._counters is None:  # 32 sets of 128 arrival counters, used round-robin: zero between calls
            self._counters = torch.zeros(32 * 128, dtype=torch.int32, device=self.device)
        seq = self.seq
        self.seq += 1
        return self._ptrs, self.world, self.rank, self.capacity, seq, self._counters.data_ptr() + 4 * 128 * (seq % 32)


def check_peer_status():
    """Raises if a join through peer memory gave up waiting for a peer (the sums of that analysis are then meaningless)."""
    from . import _lib
    code = _lib.load().kb2_peer_status()
    if code:
        raise RuntimeError(f'kinisi_b200.dist: a join through peer memory timed out waiting for a peer (code {code}): '
                           'the ranks did not issue the same analyses in the same order, or a rank died')


_PEER = {}
_PEER_MIN_CAPACITY = 1 << 15


def _peer(name, t):
    """The peer-memory channel `name` for tensors like `t`, or None (CPU tensors, not float64, another backend,
    KB2_PEER_AR=0, or the windows could not be mapped: the NCCL road is taken instead).  The process group is only the
    rendezvous for the IPC handles: a gloo group with CUDA tensors works as well (ranks may even share one GPU, which is
    how tests/test_gpu_peer.py runs two ranks on the driver's single-GPU box)."""
    if (name is None or not t.is_cuda or t.dtype != torch.float64 or not t.is_contiguous() or t.data_ptr() % 16
            or os.environ.get('KB2_PEER_AR', '1') == '0' or td.get_backend() not in ('nccl', 'gloo')
            or td.get_world_size() > 16):
        return None
    key = (name, t.device.index)
    ch = _PEER.get(key)
    if ch is False:
        return None
    n = t.numel()
    if ch is None or n > ch.capacity:
        cap = _PEER_MIN_CAPACITY
        while cap < n:
            cap *= 2
        if ch is not None:  # a longer trajectory than any before: every rank is here, nothing of this channel may be in flight
            torch.cuda.synchronize(t.device)
            td.barrier()
            ch.close()
        ch = _PeerChannel(t.device, cap)
        if not ch.ok:
            import warnings
            warnings.warn(f'kinisi_b200.dist: peer-memory windows unavailable ({ch.error}); using NCCL all-reduces')
            _PEER[key] = False
            return None
        _PEER[key] = ch
    return ch


_DIRECT = {}


def _direct(name, t):
    """The direct communicator for channel `name`, or None (CPU tensors, another backend, KB2_DIRECT_NCCL=0, or the
    library could not be driven: the process group is used instead)."""
    if not t.is_cuda or t.dtype != torch.float64 or not t.is_contiguous() or os.environ.get('KB2_DIRECT_NCCL', '1') == '0':
        return None
    key = (name, t.device.index)
    if key not in _DIRECT:
        comm = None
        if td.get_backend() == 'nccl':
            try:
                comm = _DirectComm(t.device)
            except Exception as exc:  # stay on the process group
                import warnings
                warnings.warn(f'kinisi_b200.dist: direct NCCL communicator unavailable ({exc}); using torch.distributed')
        _DIRECT[key] = comm
    return _DIRECT[key]


def moments_channel(n, device):
    """The peer-memory channel through which n per-lag sums (a flat array of one slice: n <= 1024) can be joined inside the
    kernel that consumes them (kb2_moment_stats_joined), or None.  Depends on n alone: every rank takes the same road."""
    if not active() or n > 1024 or n < 2 or os.environ.get('KB2_FUSED_JOIN', '1') == '0':
        return None
    device = torch.device(device)
    if device.type != 'cuda':
        return None
    ch = _PEER.get(('moments', device.index))
    if ch:
        return ch
    return _peer('moments', torch.empty(int(n), dtype=torch.float64, device=device))


def frame_sums_channel(n_rows, pitch, device):
    """The peer-memory channel through which a [n_rows][pitch] array of frame sums on `device` can be joined INSIDE the
    kernel that produces it (kb2_frame_sums_joined), or None: the decision depends on the shape alone, so every rank takes
    the same road."""
    if not active() or os.environ.get('KB2_FUSED_JOIN', '1') == '0':
        return None
    from . import _lib
    if _lib.load().kb2_peer_rows_slices(int(pitch)) <= 0 or n_rows != 3:
        return None
    device = torch.device(device)
    ch = _PEER.get(('frames', device.index))
    if ch and n_rows * int(pitch) <= ch.capacity:
        return ch
    probe = torch.empty((n_rows, int(pitch)), dtype=torch.float64, device=device)  # first use / a longer trajectory: the collective road
    return _peer('frames', probe)


_SUM_OPTS = None


def allreduce_sum_(t, channel=None):
    """In-place sum all-reduce of a tensor (float64 moment sums / frame sums); `channel` names the communicator
    (None: the default group).  Goes to the process group's own entry point: the checks and logging of the
    torch.distributed wrapper cost more host time than the collective takes on the device."""
    global _SUM_OPTS
    if active() and t.numel() > 0:
        peer = _peer(channel, t)
        if peer is not None:
            peer.allreduce_sum_f64(t)
            return t
        comm = _direct(channel, t)
        if comm is not None:
            comm.allreduce_sum_f64(t)
            return t
        if _SUM_OPTS is None:
            _SUM_OPTS = td.AllreduceOptions()
            _SUM_OPTS.reduceOp = td.ReduceOp.SUM
        _channel(channel).allreduce([t], _SUM_OPTS).wait()
    return t


_SCALAR_STREAMS = {}


def allreduce_sum_scalars(values, device=None):
    """Sums of host scalars over ranks (atom counts, charge sums), all in one collective.  On a CUDA device the
    collective and the read-back run on a stream of their own, so the host only waits for this tiny all-reduce and
    not for whatever is queued on the caller's stream (the previous analysis)."""
    if not active():
        return list(values)
    if device is not None and torch.device(device).type == 'cuda':
        device = torch.device(device)
        key = device.index
        if key not in _SCALAR_STREAMS:
            _SCALAR_STREAMS[key] = torch.cuda.Stream(device)
        with torch.cuda.stream(_SCALAR_STREAMS[key]):
            t = torch.tensor([float(v) for v in values], dtype=torch.float64).to(device)
            td.all_reduce(t, op=td.ReduceOp.SUM)
            return [float(v) for v in t.cpu().tolist()]
    t = torch.tensor([float(v) for v in values], dtype=torch.float64)
    td.all_reduce(t, op=td.ReduceOp.SUM)
    return [float(v) for v in t.tolist()]


def allreduce_sum_scalar(x, device=None):
    """Sum of one host scalar over ranks."""
    return allreduce_sum_scalars([x], device)[0]


def shard_bounds(n, r=None, w=None):
    """Contiguous block of `n` atoms owned by rank r of w."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    base, rem = divmod(n, w)
    lo = r * base + min(r, rem)
    return lo, lo + base + (1 if r < rem else 0)
