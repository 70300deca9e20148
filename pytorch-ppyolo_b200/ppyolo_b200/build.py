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


def have_nvcc():
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    return os.path.exists(nvcc)


def source_hash():
    """Content hash of everything the library is built from (mtimes do not survive a snapshot copy)."""
    import hashlib
    h = hashlib.sha256()
    deps = sources() + sorted(glob.glob(os.path.join(CSRC, '*.cuh'))) + sorted(glob.glob(os.path.join(ROOT, 'include', '*.h')))
    for d in deps:
        h.update(os.path.basename(d).encode())
        with open(d, 'rb') as f:
            h.update(f.read())
    h.update(' '.join(NVCC_FLAGS[:8]).encode())
    return h.hexdigest()


HASH_PATH = LIB_PATH + '.srchash'


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    if not os.path.exists(HASH_PATH):
        return True
    with open(HASH_PATH) as f:
        return f.read().strip() != source_hash()


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
    with open(HASH_PATH, 'w') as f:
        f.write(source_hash())
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
