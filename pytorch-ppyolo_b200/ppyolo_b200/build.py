"""Builds libppyolo_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import glob
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(PKG_DIR))
CSRC = os.path.join(os.path.dirname(PKG_DIR), 'csrc')
LIB_PATH = os.path.join(PKG_DIR, 'libppyolo_b200.so')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '--shared',
              '-Xcompiler', '-fPIC', '-I', os.path.join(ROOT, 'include')]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(ROOT, 'include', '*.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    objs = []
    procs = []
    build_dir = os.path.join(os.path.dirname(PKG_DIR), 'build')
    os.makedirs(build_dir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != '--shared']
    for src in sources():
        obj = os.path.join(build_dir, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        cmd = [nvcc] + flags + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, pr in procs:
        out, _ = pr.communicate()
        if verbose or pr.returncode != 0:
            sys.stderr.write(out.decode())
        if pr.returncode != 0:
            raise RuntimeError('nvcc failed: ' + ' '.join(cmd))
    link = [nvcc, '--shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB_PATH] + objs
    subprocess.check_call(link)
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
