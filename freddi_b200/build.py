"""In-tree build of libfreddi_b200.so (nvcc, sm_100a only)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
SOURCES = ['capi.cu', 'construct.cu', 'sharded.cu', 'evolve.cu', 'fields.cu', 'evolve_big.cu', 'fields_big.cu', 'resolve.cpp', 'star.cpp']
HEADERS = ['disc_engine.h', 'fb200_math.h', 'fb200_model.h', 'evolve.cuh', 'gpu_exec.cuh', 'resolve.hpp', 'star.hpp', 'ensemble_impl.hpp',
           os.path.join('..', '..', 'include', 'freddi_b200.h')]
OUT = os.path.join(HERE, 'libfreddi_b200.so')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC,-ffp-contract=off', '-shared']


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(os.path.join(CSRC, f)) <= t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    nvcc = os.environ.get('NVCC', 'nvcc')
    cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', OUT] + SOURCES
    subprocess.run(cmd, cwd=CSRC, check=True)
    return OUT
