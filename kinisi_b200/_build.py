"""Compiles the CUDA sources in csrc/ into libkinisi_b200.so next to this file (in-tree, sm_100a only)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libkinisi_b200.so')
SOURCES = ['api.cu', 'k0_ingest.cu', 'k1_unwrap.cu', 'k2_moments.cu', 'k2_moments_tma.cu', 'k2_moments_runs.cu', 'k2_moments_win.cu', 'k2_moments_tile.cu', 'k3_regress.cu', 'k4_eigen.cu', 'k6_resample.cu', 'k7_peer.cu']
NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC',
    '-Wno-deprecated-gpu-targets'
]


def _nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found: the kinisi_b200 CUDA extension cannot be built')
    return exe


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, '..', 'include', 'kinisi_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> kinisi_b200/libkinisi_b200.so"""
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, '..', 'build', 'obj')
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace('.cu', '.o'))
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj]
        if verbose:
            print(' '.join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{out}')
    cmd = [nvcc, '-shared', '-Wno-deprecated-gpu-targets', '-o', LIB] + objs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}')
    return LIB


if __name__ == '__main__':
    print(build(force=True, verbose=True))
