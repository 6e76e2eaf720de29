"""Shared helpers of the test-suite: model configurations, the host-emulation binding, comparison utilities."""
import ctypes as C
import glob
import os

import numpy as np

from freddi_b200.args import FB200Args, FB200ModelInfo, make_args

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden', 'reference_dat')
EMU_SO = os.path.join(ROOT, 'tests', 'host_emu', 'libhost_emu.so')
RTOL = 1e-10  # BASELINE.json north_star: 1e-10 relative on Sigma, F, Mdot and band fluxes at every output time

DAY = 86400.0
MSUN = 1.98892e33
KPC = 1000. * 3.08567758135e18
from freddi_b200.workloads import LAM3  # noqa: E402


def golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, '*.dat')))


def golden_passbands(paths):
    """The `passband=` entries of a golden header (relative to the golden directory) as Passband objects,
    read the way the reference's CLI does (photon counter)."""
    from freddi_b200.ensemble import Passband
    return [Passband.from_file(os.path.join(GOLDEN_DIR, p)) for p in paths]


from freddi_b200.workloads import (INI, NS_CI, WOODS, ini_args, config2_args, config3_args, config4_args,  # noqa: F401,E402
                                   config5_args)


class HostEmu:
    """tests/host_emu/libhost_emu.so: DiscEngine under the sequential host policy (test infrastructure)."""

    def __init__(self, big=False):
        self.lib = C.CDLL(EMU_SO.replace('libhost_emu.so', 'libhost_emu_big.so') if big else EMU_SO)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int64)
        self.lib.emu_run.restype = C.c_int64
        self.lib.emu_run.argtypes = [C.POINTER(FB200Args), dp, C.c_int, C.c_int, dp, C.c_int64, dp, ip, ip, ip, C.c_char_p,
                                     C.c_size_t]
        self.lib.emu_resolve.argtypes = [C.POINTER(FB200Args), C.POINTER(FB200ModelInfo), dp, dp, C.c_char_p, C.c_size_t]
        self.lib.emu_planck_nu1_nu2.restype = C.c_double
        self.lib.emu_planck_nu1_nu2.argtypes = [C.c_double] * 4 + [C.POINTER(C.c_int)]

    def set_passbands(self, passbands=()):
        passbands = list(passbands)
        dp = C.POINTER(C.c_double)
        self.lib.emu_set_passbands.argtypes = [C.c_int, C.POINTER(C.c_int), dp, dp, dp]
        lens = (C.c_int * max(1, len(passbands)))(*[pb.lambdas.size for pb in passbands])
        lam = np.concatenate([pb.lambdas for pb in passbands]) if passbands else np.zeros(1)
        tr = np.concatenate([pb.transmissions for pb in passbands]) if passbands else np.zeros(1)
        tdnu = np.zeros(max(1, len(passbands)))
        rc = self.lib.emu_set_passbands(len(passbands), lens, lam.ctypes.data_as(dp), tr.ctypes.data_as(dp), tdnu.ctypes.data_as(dp))
        if rc != 0:
            raise ValueError('emu_set_passbands: {}'.format(rc))
        return tdnu[:len(passbands)]

    def run(self, args, Nt, Nx, lambdas=(), cold_disk=False, passbands=(), star_flux=False):
        if passbands or star_flux:
            self.set_passbands(passbands)
            self.lib.emu_set_star_flux(int(bool(star_flux)))
            try:
                return self.run(args, Nt, Nx, lambdas, cold_disk)
            finally:
                self.set_passbands(())
                self.lib.emu_set_star_flux(0)
        lam = np.ascontiguousarray(lambdas, dtype=np.float64)
        ncol = (15 + (len(lam) + self.n_passbands()) * ((2 if cold_disk else 1) + (3 if self.lib.emu_star_flux() else 0))
                + (10 if args.is_ns else 0))
        rows = np.full((Nt + 1, ncol), np.nan)
        F = np.full((Nt + 1, Nx), np.nan)
        first = np.full(Nt + 1, -1, dtype=np.int64)
        last = np.full(Nt + 1, -1, dtype=np.int64)
        pic = np.zeros(Nt + 1, dtype=np.int64)
        err = C.create_string_buffer(512)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int64)
        n = self.lib.emu_run(C.byref(args), lam.ctypes.data_as(dp), len(lam), int(cold_disk), rows.ctypes.data_as(dp), Nt + 1,
                             F.ctypes.data_as(dp), first.ctypes.data_as(ip), last.ctypes.data_as(ip), pic.ctypes.data_as(ip),
                             err, 512)
        if n < 0:
            raise RuntimeError('[{}] {}'.format(n, err.value.decode()))
        return dict(rows=rows, rows_valid=int(n), F=F, first=first, last=last, picard=pic)

    def star_triangles(self, args, cap=100000):
        out = np.empty((cap, 7))
        self.lib.emu_star_triangles.restype = C.c_long
        self.lib.emu_star_triangles.argtypes = [C.POINTER(FB200Args), C.POINTER(C.c_double), C.c_long]
        n = self.lib.emu_star_triangles(C.byref(args), out.ctypes.data_as(C.POINTER(C.c_double)), cap)
        if n < 0:
            raise RuntimeError('emu_star_triangles: {}'.format(n))
        return out[:n].copy()

    def n_passbands(self):
        return int(self.lib.emu_n_passbands())

    def planck(self, T, nu1, nu2, tol=1e-4):
        n = C.c_int()
        v = self.lib.emu_planck_nu1_nu2(T, nu1, nu2, tol, C.byref(n))
        return v, n.value


def rel_err(v, g):
    """Element-wise relative error with the parity rules of SURVEY.md 8d: exact zeros compare with ==,
    NaN must match NaN, inf must match inf."""
    v = np.asarray(v, dtype=float)
    g = np.asarray(g, dtype=float)
    with np.errstate(all='ignore'):
        rel = np.abs(v - g) / np.abs(g)
    rel = np.where(g == 0, np.where(v == 0, 0., np.inf), rel)
    rel = np.where(np.isnan(g) | np.isnan(v), np.where(np.isnan(g) & np.isnan(v), 0., np.inf), rel)
    rel = np.where(np.isinf(g) | np.isinf(v), np.where(v == g, 0., np.inf), rel)
    return rel


def worst_columns(rows, ref_rows, names):
    out = {}
    for k, c in enumerate(names):
        out[c] = float(np.max(rel_err(rows[:, k], ref_rows[:, k]))) if rows.shape[0] else 0.
    return out


def assert_rows_close(rows, ref_rows, names, rtol=RTOL, what=''):
    worst = worst_columns(rows, ref_rows, names)
    bad = {k: v for k, v in worst.items() if not v <= rtol}
    assert not bad, '{}: columns beyond rtol={}: {}'.format(what, rtol, bad)
    return worst
