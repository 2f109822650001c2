"""Build the C-ABI CUDA library in-tree: telogator2_b200/libtg_b200.so (sm_100a only).

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
SO = os.path.join(HERE, 'libtg_b200.so')
SOURCES = ['tg_api.cu', 'tg_dp.cu', 'tg_scan.cu', 'tg_cov.cu', 'tg_cv.cu', 'tg_prefilter.cu', 'tg_bound.cu', 'tg_linkage.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC,-fvisibility=hidden', '-I' + os.path.join(ROOT, 'include'), '-I' + CSRC]


def _stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, 'include', 'telogator2_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def host_ext_path():
    import sysconfig
    return os.path.join(HERE, '_tg_hostlists' + (sysconfig.get_config_var('EXT_SUFFIX') or '.so'))


def build_host_ext(force=False):
    """telogator2_b200/_tg_hostlists*.so: CPython extension that formats span arrays as the reference's nested lists and packs
    read strings into the upload buffer (chost/tg_hostlists.c; gcc, no third-party headers)."""
    import sysconfig
    src = os.path.join(HERE, 'chost', 'tg_hostlists.c')
    out = host_ext_path()
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    cc = os.environ.get('CC', 'gcc')
    subprocess.check_call([cc, '-O2', '-shared', '-fPIC', '-fvisibility=hidden', '-I' + sysconfig.get_paths()['include'], src, '-o', out])
    return out


def build(force=False, verbose=False):
    build_host_ext(force)
    if not force and not _stale():
        return SO
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    objs = []
    bdir = os.path.join(HERE, 'build')
    os.makedirs(bdir, exist_ok=True)
    for s in SOURCES:
        o = os.path.join(bdir, s.replace('.cu', '.o'))
        cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', os.path.join(CSRC, s), '-o', o]
        subprocess.check_call(cmd)
        objs.append(o)
    subprocess.check_call([nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', SO] + objs + ['-lcudart'])
    return SO


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
