"""Multi-GPU plumbing: one process per GPU, atoms sharded across ranks, per-dt moment sums and the
per-frame framework / collective sums joined by a sum all-reduce (NCCL over NVLink on the GPU box,
gloo in the CPU tests).  Everything is a no-op when torch.distributed is not initialised."""
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


_SUM_OPTS = None


def allreduce_sum_(t, channel=None):
    """In-place sum all-reduce of a tensor (float64 moment sums / frame sums); `channel` names the communicator
    (None: the default group).  Goes to the process group's own entry point: the checks and logging of the
    torch.distributed wrapper cost more host time than the collective takes on the device."""
    global _SUM_OPTS
    if active():
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
