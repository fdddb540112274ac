"""Build libr2n2_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import hashlib
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
# diagnostic build (tools/waitstat.py): per-warp mbarrier wait statistics; a SEPARATE library, never the product's
WAITSTAT = bool(os.environ.get('R2N2_WAITSTAT'))
if WAITSTAT:
    NVCC_FLAGS += ['-DR2N2_WAITSTAT']
    OUT = os.path.join(PKG, 'libr2n2_b200_ws.so')


HASH_FILE = OUT + '.srchash'
last_action = None       # 'compiled' / 'reused' (what the last build() call did)


def source_hash():
    """sha256 over every .cu / .cuh source, the public header and the compile flags: the library embeds it
    (r2n2_version()) and the sidecar file records it, so a stale prebuilt .so is never reused silently."""
    deps = sorted(os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(('.cu', '.cuh')))
    deps.append(os.path.join(os.path.dirname(PKG), 'include', 'r2n2_b200.h'))
    h = hashlib.sha256()
    for d in deps:
        h.update(os.path.basename(d).encode())
        with open(d, 'rb') as f:
            h.update(f.read())
    h.update(' '.join(f for f in NVCC_FLAGS if f != '-v' and f != '-Xptxas').encode())
    return h.hexdigest()[:16]


def needs_build():
    if not os.path.exists(OUT) or not os.path.exists(HASH_FILE):
        return True
    with open(HASH_FILE) as f:
        return f.read().strip() != source_hash()


def build(force=False, verbose=False):
    global last_action
    if not force and not needs_build():
        last_action = 'reused'
        return OUT
    last_action = 'compiled'
    src_hash = source_hash()
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(HERE, s.replace('.cu', '.ws.o' if WAITSTAT else '.o'))
        objs.append(o)
        cmd = [nvcc] + NVCC_FLAGS + ['-DR2N2_SRC_HASH="%s"' % src_hash, '-c', os.path.join(HERE, s), '-o', o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out.decode())
        if p.returncode != 0:
            raise RuntimeError('nvcc failed: ' + ' '.join(cmd))
    cmd = [nvcc, '-shared', '-o', OUT] + objs + ['-cudart', 'shared']
    subprocess.check_call(cmd)
    with open(HASH_FILE, 'w') as f:
        f.write(src_hash + '\n')
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
