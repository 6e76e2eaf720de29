"""In-tree build of libfreddi_b200.so (nvcc, sm_100a only).

Every translation unit is compiled to an object of its own (in parallel: the evolve kernel is compiled four times — the main,
small-grid and large-grid builds — at about a minute each) and the objects are linked into the one shared library.
"""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'csrc', '_obj')
SOURCES = ['capi.cu', 'construct.cu', 'sharded.cu', 'evolve.cu', 'fields.cu', 'evolve_big.cu', 'fields_big.cu', 'evolve_s128.cu',
           'fields_s128.cu', 'evolve_s256.cu', 'fields_s256.cu', 'evolve_s512.cu', 'fields_s512.cu', 'resolve.cpp', 'star.cpp', 'lx_table.cpp']
HEADERS = ['disc_engine.h', 'lx_table.h', 'fb200_math.h', 'fb200_model.h', 'evolve.cuh', 'gpu_exec.cuh', 'resolve.hpp', 'star.hpp', 'ensemble_impl.hpp',
           os.path.join('..', '..', 'include', 'freddi_b200.h')]
OUT = os.path.join(HERE, 'libfreddi_b200.so')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC,-ffp-contract=off']
# what a translation unit includes besides itself (a kernel build includes evolve.cu / fields.cu textually)
KERNEL_HEADERS = ['disc_engine.h', 'lx_table.h', 'fb200_math.h', 'fb200_model.h', 'evolve.cuh', 'gpu_exec.cuh', os.path.join('..', '..', 'include', 'freddi_b200.h')]


def _deps(src):
    deps = [src] + HEADERS
    if src.startswith('evolve_'):
        deps.append('evolve.cu')
    if src.startswith('fields_'):
        deps.append('fields.cu')
    return [os.path.join(CSRC, d) for d in deps]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(os.path.join(CSRC, f)) <= t for f in SOURCES + HEADERS)


def build(force=False, verbose=False, out=None, kernel_flags=(), name=None):
    """The product library, or — with `out` — a kernel-variant library for A/B measurements and profiling
    (FB200_LIB=<out> python bench.py ...): `kernel_flags` (e.g. -DFB_PROFILE, -DFB_NT=512) go to the kernel translation units
    (evolve*.cu, fields*.cu); objects live in csrc/_obj/<name>."""
    variant = out is not None
    if not variant and not force and up_to_date():
        return OUT
    nvcc = os.environ.get('NVCC', 'nvcc')
    obj_dir = os.path.join(OBJ, name or os.path.splitext(os.path.basename(out))[0]) if variant else OBJ
    os.makedirs(obj_dir, exist_ok=True)
    flags = NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else [])

    def compile_one(src):
        obj = os.path.join(obj_dir, os.path.splitext(src)[0] + '.o')
        kernel = src.startswith(('evolve', 'fields'))
        if force or verbose or (variant and kernel) or _stale(obj, _deps(src)):
            subprocess.run([nvcc] + flags + (list(kernel_flags) if kernel else []) + ['-c', src, '-o', obj], cwd=CSRC, check=True)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    target = os.path.abspath(out) if variant else OUT
    os.makedirs(os.path.dirname(target), exist_ok=True)
    subprocess.run([nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', target] + objs, cwd=CSRC, check=True)
    return target


if __name__ == '__main__':
    import argparse
    ap = argparse.ArgumentParser(description=build.__doc__)
    ap.add_argument('--out', default=None, help='variant library to write (default: the product library)')
    ap.add_argument('--force', action='store_true')
    ap.add_argument('--verbose', action='store_true', help='-Xptxas -v')
    ap.add_argument('kernel_flags', nargs='*', help='extra nvcc flags for the kernel translation units, after --')
    o = ap.parse_args()
    print(build(force=o.force, verbose=o.verbose, out=o.out, kernel_flags=o.kernel_flags))
