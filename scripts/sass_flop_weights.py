#!/usr/bin/env python
"""FP64 flop per call of CUDA's double-precision math functions on sm_100a — the weights bench.py's roofline uses
(pow, exp, log, division, sqrt; DFMA = 2 flop, DMUL = DADD = 1; SURVEY.md 8d asks for SASS-counted weights).

    python scripts/sass_flop_weights.py                 # here (no GPU): static counts from `cuobjdump -sass`, per function,
                                                        # split into the inline body and the out-of-line subroutines
    python scripts/sass_flop_weights.py --run           # on the GPU box, under ncu: EXECUTED flop per call on typical
                                                        # arguments (smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on)

Both outputs are committed as profiles/sass_flop_weights.txt.  The static count of the whole function is an upper bound
(it includes the slow paths for denormal / huge / special operands that a call with ordinary operands branches around);
the executed count is what one call really costs, and is the number the roofline arithmetic uses."""
import argparse
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, 'build', 'sass_weights')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']

SRC = r'''
#include <math.h>
#include <cstdio>
#include <cuda_runtime.h>
// one kernel per function; arguments are read from memory so that nothing is folded at compile time
extern "C" __global__ void k_pow(double* o, const double* a, const double* b, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) o[i] = pow(a[i], b[i]); }
extern "C" __global__ void k_exp(double* o, const double* a, const double* b, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) o[i] = exp(a[i]); }
extern "C" __global__ void k_log(double* o, const double* a, const double* b, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) o[i] = log(a[i]); }
extern "C" __global__ void k_div(double* o, const double* a, const double* b, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) o[i] = a[i] / b[i]; }
extern "C" __global__ void k_sqrt(double* o, const double* a, const double* b, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) o[i] = sqrt(a[i]); }
extern "C" __global__ void k_none(double* o, const double* a, const double* b, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) o[i] = a[i] + b[i]; }
#ifdef WITH_MAIN
int main() {
    const int n = 1 << 20;
    double *o, *a, *b;
    cudaMalloc(&o, n * 8); cudaMalloc(&a, n * 8); cudaMalloc(&b, n * 8);
    double* ha = new double[n]; double* hb = new double[n];
    // typical operands of the path: viscous torques 1e30..1e40 raised to 1 - m = 0.7, x = h nu / kT in 0.1..60,
    // ratios and sums of ordinary magnitude
    for (int i = 0; i < n; ++i) { ha[i] = 1e30 * pow(1e10, (i % 1000) / 1000.); hb[i] = 0.7; }
    cudaMemcpy(a, ha, n * 8, cudaMemcpyHostToDevice); cudaMemcpy(b, hb, n * 8, cudaMemcpyHostToDevice);
    k_pow<<<n / 256, 256>>>(o, a, b, n);
    for (int i = 0; i < n; ++i) { ha[i] = 0.1 + 60. * (i % 1000) / 1000.; hb[i] = 3. + (i % 7); }
    cudaMemcpy(a, ha, n * 8, cudaMemcpyHostToDevice); cudaMemcpy(b, hb, n * 8, cudaMemcpyHostToDevice);
    k_exp<<<n / 256, 256>>>(o, a, b, n);
    k_log<<<n / 256, 256>>>(o, a, b, n);
    k_div<<<n / 256, 256>>>(o, a, b, n);
    k_sqrt<<<n / 256, 256>>>(o, a, b, n);
    k_none<<<n / 256, 256>>>(o, a, b, n);
    cudaDeviceSynchronize();
    printf("launched %d threads per kernel: %s\n", n, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
#endif
'''

FP64_OPS = ('DFMA', 'DMUL', 'DADD', 'DSETP', 'MUFU.RCP64H', 'MUFU.RSQ64H', 'MUFU.EX2', 'MUFU.LG2')


def static_counts():
    os.makedirs(BUILD, exist_ok=True)
    cu = os.path.join(BUILD, 'w.cu')
    with open(cu, 'w') as f:
        f.write(SRC)
    cubin = os.path.join(BUILD, 'w.cubin')
    subprocess.run(['nvcc'] + ARCH + ['-O3', '-cubin', '-o', cubin, cu], check=True)
    sass = subprocess.run(['cuobjdump', '-sass', cubin], check=True, capture_output=True, text=True).stdout
    out = []
    fn, section = None, None
    counts = {}
    for line in sass.splitlines():
        m = re.search(r'Function : (\w+)', line)
        if m:
            fn, section = m.group(1), 'inline'
            counts[fn] = {'inline': collections.Counter(), 'subroutines': collections.Counter()}
            continue
        if fn is None:
            continue
        m = re.search(r'/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if not m:
            continue
        predicated, op = m.group(1) is not None, m.group(2)
        for k in FP64_OPS:
            if op == k or op.startswith(k + '.') or (k.startswith('MUFU') and op == k):
                counts[fn][section][k] += 1
        if op == 'EXIT' and not predicated:
            section = 'subroutines'  # what follows the kernel's EXIT are the out-of-line routines it CALLs
    out.append('static FP64 instruction counts per function (cuobjdump -sass, nvcc {} -O3):'.format(' '.join(ARCH)))
    out.append('%-8s %-12s %5s %5s %5s %6s %5s  %s' % ('function', 'section', 'DFMA', 'DMUL', 'DADD', 'DSETP', 'MUFU', 'flop = 2 DFMA + DMUL + DADD'))
    base = counts.get('k_none')
    for fn in ('k_pow', 'k_exp', 'k_log', 'k_div', 'k_sqrt', 'k_none'):
        tot = 0
        for sec in ('inline', 'subroutines'):
            c = counts[fn][sec]
            flop = 2 * c['DFMA'] + c['DMUL'] + c['DADD']
            tot += flop
            mufu = sum(v for k, v in c.items() if k.startswith('MUFU'))
            out.append('%-8s %-12s %5d %5d %5d %6d %5d  %d' % (fn[2:], sec, c['DFMA'], c['DMUL'], c['DADD'], c['DSETP'], mufu, flop))
        out.append('%-8s %-12s %s  %d' % (fn[2:], 'all paths', ' ' * 31, tot))
    out.append('(k_none = a + b: the 1 DADD of the harness itself; pow CALLs its core routine on every call — it is not a slow path —')
    out.append(' div and sqrt CALL a subroutine only for operands outside the fast range)')
    return '\n'.join(out)


def executed_counts():
    os.makedirs(BUILD, exist_ok=True)
    cu = os.path.join(BUILD, 'w_main.cu')
    with open(cu, 'w') as f:
        f.write('#define WITH_MAIN 1\n' + SRC)
    exe = os.path.join(BUILD, 'w_run')
    subprocess.run(['nvcc'] + ARCH + ['-O3', '-o', exe, cu], check=True)
    metrics = ','.join('smsp__sass_thread_inst_executed_op_%s_pred_on.sum' % op for op in ('dfma', 'dmul', 'dadd'))
    log = os.path.join(BUILD, 'w_ncu.csv')
    subprocess.run(['ncu', '--metrics', metrics, '--clock-control', 'none', '--csv', '--log-file', log, exe], check=True, capture_output=True)
    per = collections.defaultdict(dict)
    with open(log) as f:
        rows = [r for r in csv.reader(f) if len(r) > 10]
    hdr = rows[0]
    ik, im, iv = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value')
    for r in rows[1:]:
        per[r[ik].split('(')[0]][r[im]] = float(r[iv].replace(',', ''))
    n = 1 << 20
    out = ['executed FP64 thread-instructions per call (ncu, 2^20 threads, typical operands: pow(1e30..1e40, 0.7), exp/log/sqrt(0.1..60), a / b):',
           '%-8s %8s %8s %8s  %s' % ('function', 'DFMA', 'DMUL', 'DADD', 'flop = 2 DFMA + DMUL + DADD')]
    for fn in ('k_pow', 'k_exp', 'k_log', 'k_div', 'k_sqrt', 'k_none'):
        d = per[fn]
        a, b, c = (d.get('smsp__sass_thread_inst_executed_op_%s_pred_on.sum' % op, 0.) / n for op in ('dfma', 'dmul', 'dadd'))
        out.append('%-8s %8.2f %8.2f %8.2f  %.1f' % (fn[2:], a, b, c, 2 * a + b + c))
    return '\n'.join(out)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--run', action='store_true', help='also measure the executed counts with ncu (needs a GPU)')
    opts = ap.parse_args()
    print(static_counts())
    if opts.run:
        print()
        print(executed_counts())
