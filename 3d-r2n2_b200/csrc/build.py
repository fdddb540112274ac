"""Build libr2n2_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, 'libr2n2_b200.so')
SOURCES = ['elementwise.cu', 'conv_simt.cu', 'conv_umma.cu', 'conv_umma_fwd.cu', 'conv_umma_wgrad2.cu', 'conv_umma_wgrad3.cu', 'conv_umma_fwd3.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Wno-deprecated-gpu-targets']
if os.environ.get('R2N2_PTXAS_V'):
    NVCC_FLAGS += ['-Xptxas', '-v']


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(('.cu', '.cuh'))]
    deps.append(os.path.join(os.path.dirname(PKG), 'include', 'r2n2_b200.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(HERE, s.replace('.cu', '.o'))
        objs.append(o)
        cmd = [nvcc] + NVCC_FLAGS + ['-c', os.path.join(HERE, s), '-o', o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out.decode())
        if p.returncode != 0:
            raise RuntimeError('nvcc failed: ' + ' '.join(cmd))
    cmd = [nvcc, '-shared', '-o', OUT] + objs + ['-cudart', 'shared']
    subprocess.check_call(cmd)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
