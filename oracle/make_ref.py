"""oracle/make_ref.py -- TEST INFRASTRUCTURE ONLY.

Copies the UNMODIFIED reference package (bjmorgan/kinisi, /root/reference/kinisi/*.py, tests left out) into
oracle/_ref/kinisi so that it can travel to the GPU box (oracle/_ref/ is git-ignored and never part of the history, but it
is not gpurun-ignored).  Run by `__graft_entry__.build()` when /root/reference is present; `bench.py --impl reference` and
the `cpu_baseline` leg import it from there with the three stand-in modules of oracle/shims on the path (uravu, emcee,
statsmodels are not installed in this image, see oracle/shims/README.md).  Nothing under kinisi_b200/ imports it.

    python oracle/make_ref.py            # -> oracle/_ref/kinisi/, oracle/_ref/SOURCE.txt
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = '/root/reference/kinisi'
DST = os.path.join(HERE, '_ref', 'kinisi')


def make(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print(f'{SRC} is absent: keeping whatever is under {DST}')
        return os.path.isdir(DST)
    os.makedirs(DST, exist_ok=True)
    lines = []
    for name in sorted(os.listdir(SRC)):
        if not name.endswith('.py'):
            continue
        shutil.copyfile(os.path.join(SRC, name), os.path.join(DST, name))
        lines.append(f'{hashlib.sha256(open(os.path.join(SRC, name), "rb").read()).hexdigest()}  kinisi/{name}')
    with open(os.path.join(HERE, '_ref', 'SOURCE.txt'), 'w') as f:
        f.write('byte-for-byte copies of /root/reference/kinisi/*.py (bjmorgan/kinisi 1.1.1), sha256:\n' + '\n'.join(lines) + '\n')
    if verbose:
        print(f'copied {len(lines)} files to {DST}')
    return True


def import_reference():
    """Puts oracle/shims and oracle/_ref in front of sys.path and returns (kinisi.parser, kinisi.diffusion) of the unmodified
    reference, or None when oracle/_ref is missing."""
    if not os.path.isfile(os.path.join(DST, 'parser.py')):
        return None
    for p in (os.path.join(HERE, '_ref'), os.path.join(HERE, 'shims')):
        if p not in sys.path:
            sys.path.insert(0, p)
    from kinisi import diffusion, parser  # noqa: the unmodified reference
    return parser, diffusion


if __name__ == '__main__':
    make()
