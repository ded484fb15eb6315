"""Build the native library in-tree: ``terran_b200/lib/libterran_b200.so``.

nvcc cross-compiles for sm_100a without a GPU; the built ``.so`` is git-ignored
but travels with the working tree.  ``python -m terran_b200.build`` or
``__graft_entry__.build()``.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
LIB = os.path.join(LIBDIR, 'libterran_b200.so')

NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
COMMON = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC']

#: source -> extra flags.  The post-processing kernels reproduce the oracle's
#: fp32 operation order bit for bit, so fused multiply-add contraction is off.
SOURCES = {
    'conv_tc.cu': [],
    'conv_patch.cu': [],
    'conv_direct.cu': [],
    'conv_mma.cu': [],
    'detect_post.cu': ['-fmad=false'],
    'pose_parse.cu': ['-fmad=false'],
    'net.cu': [],
    'program.cu': [],
}


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    headers.append(os.path.join(os.path.dirname(HERE), 'include', 'terran_b200.h'))
    jobs = []
    objs = []
    for src, extra in SOURCES.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(LIBDIR, src.replace('.cu', '.o'))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [NVCC] + ARCH + COMMON + extra + ['-c', s, '-o', o]
            if verbose:
                cmd += ['-Xptxas', '-v']
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or verbose:
            sys.stderr.write(' '.join(cmd) + '\n' + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed: ' + ' '.join(cmd))

    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        list(ex.map(run, jobs))
    if jobs or force or _stale(LIB, objs):
        run([NVCC] + ARCH + ['-shared', '-o', LIB] + objs)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
