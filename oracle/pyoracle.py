"""TEST INFRASTRUCTURE — ctypes access to the checkers.

``Ref`` : ``oracle/_ref/libfreddi_ref.so`` — the unmodified reference sources compiled in place against
oracle/boost_shim, driven through the C functions of oracle/ref_driver.cpp (prefix ``ref_``) over ``fb200_args_t``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this module; nothing
under ``freddi_b200/`` does.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np

from freddi_b200.args import FB200Args, FB200ModelInfo, args_array

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, '_ref', 'libfreddi_ref.so')

N_BH_COLS = 15
N_NS_COLS = 10
BH_COLS = ['t', 'Mdot', 'Mdisk', 'Rhot', 'Sigmaout', 'Kirrout', 'H2R', 'Teffout', 'Tirrout', 'Qirr2Qvisout',
           'TphXmax', 'Lx', 'Lbol', 'Fx', 'Fbol']
NS_COLS = ['Rin', 'etans', 'Lxns', 'Lbolns', 'Fxns', 'Fbolns', 'Thotspot', 'fpin', 'Fmagnin', 'Fin']
FIELDS = {'h': 0, 'R': 1, 'F': 2, 'W': 3, 'Sigma': 4, 'Tph': 5, 'Tph_vis': 6, 'Tirr': 7, 'Kirr': 8, 'Height': 9,
          'Qx': 10, 'Tph_X': 11, 'windA': 12, 'windB': 13, 'windC': 14, 'Fmagn': 15, 'dFmagn_dh': 16,
          'd2Fmagn_dh2': 17}


def build(target='ref'):
    subprocess.run(['make', '-s', '-C', HERE, target], check=True)


def column_names(n_lambdas=0, cold_disk=False, is_ns=False, passbands=(), star_flux=False):
    cols = list(BH_COLS)
    for name in [str(k) for k in range(n_lambdas)] + list(passbands):
        cols.append('Fnu{}'.format(name))
        if cold_disk:
            cols.append('Fnu{}_cold'.format(name))
        if star_flux:
            cols += ['Fnu{}_star'.format(name), 'Fnu{}_star_min'.format(name), 'Fnu{}_star_max'.format(name)]
    if is_ns:
        cols += NS_COLS
    return cols


class CheckerError(Exception):
    def __init__(self, code, msg):
        super().__init__('[{}] {}'.format(code, msg))
        self.code = code
        self.msg = msg


class _Checker:
    prefix = None
    path = None

    def __init__(self):
        if not os.path.exists(self.path):
            raise FileNotFoundError(self.path)
        self.lib = C.CDLL(self.path)
        p = self.prefix
        L = self.lib
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int64)
        f = getattr(L, p + 'create')
        f.restype = C.c_void_p
        f.argtypes = [C.POINTER(FB200Args), dp, C.c_int, C.POINTER(C.c_int), C.c_char_p, C.c_size_t]
        getattr(L, p + 'destroy').argtypes = [C.c_void_p]
        getattr(L, p + 'n_cols').argtypes = [C.c_void_p, C.c_int]
        getattr(L, p + 'step').argtypes = [C.c_void_p]
        f = getattr(L, p + 'last_picard'); f.restype = C.c_long; f.argtypes = [C.c_void_p]
        getattr(L, p + 'row').argtypes = [C.c_void_p, C.c_int, dp]
        getattr(L, p + 'state').argtypes = [C.c_void_p, dp, ip, ip, ip]
        getattr(L, p + 'info').argtypes = [C.c_void_p, C.POINTER(FB200ModelInfo)]
        getattr(L, p + 'field').argtypes = [C.c_void_p, C.c_int, dp]
        f = getattr(L, p + 'flux'); f.restype = C.c_double; f.argtypes = [C.c_void_p, C.c_int, C.c_double]
        f = getattr(L, p + 'mdot_wind'); f.restype = C.c_double; f.argtypes = [C.c_void_p]
        f = getattr(L, p + 'run'); f.restype = C.c_int64
        f.argtypes = [C.POINTER(FB200Args), dp, C.c_int, C.c_int, dp, C.c_int64, dp, ip, ip, ip, C.c_char_p, C.c_size_t]
        f = getattr(L, p + 'run_ensemble'); f.restype = C.c_int64
        f.argtypes = [C.POINTER(FB200Args), C.c_int64, dp, C.c_int, C.c_int, C.c_int, C.c_int64, dp, C.c_int64,
                      C.POINTER(C.c_int32), dp, dp, ip]

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    # ---- whole light curve ----
    def set_passbands(self, passbands=()):
        """Passbands (objects with .lambdas [cm] and .transmissions) of the models built afterwards; () clears.
        Returns EnergyPassband::t_dnu of each."""
        passbands = list(passbands)
        f = self._fn('set_passbands')
        f.restype = None
        dp = C.POINTER(C.c_double)
        f.argtypes = [C.c_int, C.POINTER(C.c_int), dp, dp, dp]
        lens = (C.c_int * max(1, len(passbands)))(*[pb.lambdas.size for pb in passbands])
        lam = np.concatenate([pb.lambdas for pb in passbands]) if passbands else np.zeros(1)
        tr = np.concatenate([pb.transmissions for pb in passbands]) if passbands else np.zeros(1)
        tdnu = np.zeros(max(1, len(passbands)))
        f(len(passbands), lens, lam.ctypes.data_as(dp), tr.ctypes.data_as(dp), tdnu.ctypes.data_as(dp))
        return tdnu[:len(passbands)]

    def trapz(self, x, y, first=0, last=None):
        """trapz(x, y, first, last) of cpp/src/util.cpp:9-19."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        f = self._fn('trapz')
        f.restype = C.c_double
        dp = C.POINTER(C.c_double)
        f.argtypes = [dp, dp, C.c_size_t, C.c_size_t, C.c_size_t]
        return f(x.ctypes.data_as(dp), y.ctypes.data_as(dp), x.size, int(first), x.size - 1 if last is None else int(last))

    def define_mdot0(self, value=float('nan')):
        """DiskStructureArguments::Mdot0 of the BH models built next: the reference leaves this member uninitialised
        (cpp/src/arguments.cpp:42-55) although Cambier2013Wind reads it (cpp/src/freddi_state.cpp:408,411); NaN: untouched."""
        f = self._fn('define_mdot0')
        f.restype = None
        f.argtypes = [C.c_double]
        f(float(value))

    def set_star_flux(self, on):
        self._fn('set_star_flux')(int(bool(on)))

    def run(self, args, lambdas=(), cold_disk=False, snapshots=True, passbands=(), star_flux=False):
        """Returns dict(rows [(Nt+1) x ncol, NaN after collapse], rows_valid, F, first, last, picard)."""
        if passbands or star_flux:
            self.set_passbands(passbands)
            self.set_star_flux(star_flux)
            try:
                return self.run(args, lambdas, cold_disk, snapshots)
            finally:
                self.set_passbands(())
                self.set_star_flux(False)
        lam = np.ascontiguousarray(lambdas, dtype=np.float64)
        m = self.model(args, lambdas)
        info = m.info()
        Nt, Nx = int(info.Nt), int(info.Nx)
        ncol = m.n_cols(cold_disk)
        rows = np.full((Nt + 1, ncol), np.nan)
        F = np.full((Nt + 1, Nx), np.nan) if snapshots else None
        first = np.full(Nt + 1, -1, dtype=np.int64)
        last = np.full(Nt + 1, -1, dtype=np.int64)
        picard = np.zeros(Nt + 1, dtype=np.int64)
        err = C.create_string_buffer(512)
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int64)
        n = self._fn('run')(C.byref(args), lam.ctypes.data_as(dp), len(lam), int(cold_disk), rows.ctypes.data_as(dp),
                            Nt + 1, F.ctypes.data_as(dp) if snapshots else None, first.ctypes.data_as(ip),
                            last.ctypes.data_as(ip), picard.ctypes.data_as(ip), err, len(err))
        if n < 0:
            raise CheckerError(int(n), err.value.decode())
        return dict(rows=rows, rows_valid=int(n), F=F, first=first, last=last, picard=picard, Nt=Nt, Nx=Nx)

    def model(self, args, lambdas=()):
        return _Model(self, args, lambdas)

    def run_ensemble(self, args_list, lambdas=(), cold_disk=False, n_threads=0, max_steps=-1, want_rows=False,
                     rows_per_model=0):
        lam = np.ascontiguousarray(lambdas, dtype=np.float64)
        arr = args_array(args_list) if not isinstance(args_list, C.Array) else args_list
        n = len(arr)
        dp = C.POINTER(C.c_double)
        ncol = N_BH_COLS + len(lam) * (2 if cold_disk else 1) + (N_NS_COLS if arr[0].is_ns else 0)
        rows = np.full((n, rows_per_model, ncol), np.nan) if want_rows else None
        valid = np.zeros(n, dtype=np.int32)
        picard = np.zeros(n, dtype=np.int64)
        tc = C.c_double()
        te = C.c_double()
        steps = self._fn('run_ensemble')(arr, n, lam.ctypes.data_as(dp), len(lam), int(cold_disk), int(n_threads),
                                         int(max_steps), rows.ctypes.data_as(dp) if want_rows else None,
                                         int(rows_per_model), valid.ctypes.data_as(C.POINTER(C.c_int32)),
                                         C.byref(tc), C.byref(te), picard.ctypes.data_as(C.POINTER(C.c_int64)))
        return dict(steps=int(steps), construct_s=tc.value, evolve_s=te.value, rows=rows, rows_valid=valid, picard=picard)


class _Model:
    """One live model: step() / row() / field() — the reference's FreddiEvolution seen through C."""

    def __init__(self, checker, args, lambdas=()):
        self.c = checker
        self.lam = np.ascontiguousarray(lambdas, dtype=np.float64)
        self.is_ns = bool(args.is_ns)
        code = C.c_int()
        err = C.create_string_buffer(512)
        self.h = checker._fn('create')(C.byref(args), self.lam.ctypes.data_as(C.POINTER(C.c_double)), len(self.lam),
                                       C.byref(code), err, len(err))
        if not self.h:
            raise CheckerError(code.value, err.value.decode())
        self._info = None

    def __del__(self):
        if getattr(self, 'h', None):
            self.c._fn('destroy')(self.h)
            self.h = None

    def info(self):
        if self._info is None:
            self._info = FB200ModelInfo()
            self.c._fn('info')(self.h, C.byref(self._info))
        return self._info

    def n_cols(self, cold_disk=False):
        return int(self.c._fn('n_cols')(self.h, int(cold_disk)))

    def star_triangles(self):
        """(n, 7): centre, unit normal, area of every triangle of the companion star."""
        f = self.c._fn('star_triangles')
        f.restype = C.c_long
        f.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        n = f(self.h, None)
        out = np.empty((n, 7))
        f(self.h, out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def step(self):
        return self.c._fn('step')(self.h)

    def last_picard(self):
        return int(self.c._fn('last_picard')(self.h))

    def row(self, cold_disk=False):
        n = self.c._fn('n_cols')(self.h, int(cold_disk))
        out = np.empty(n)
        self.c._fn('row')(self.h, int(cold_disk), out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def state(self):
        t = C.c_double()
        i_t, first, last = C.c_int64(), C.c_int64(), C.c_int64()
        self.c._fn('state')(self.h, C.byref(t), C.byref(i_t), C.byref(first), C.byref(last))
        return dict(t=t.value, i_t=i_t.value, first=first.value, last=last.value)

    def field(self, name):
        out = np.empty(int(self.info().Nx))
        rc = self.c._fn('field')(self.h, FIELDS[name], out.ctypes.data_as(C.POINTER(C.c_double)))
        if rc != 0:
            raise KeyError(name)
        return out

    def flux(self, lam, region=0):
        return self.c._fn('flux')(self.h, region, float(lam))

    def mdot_wind(self):
        return self.c._fn('mdot_wind')(self.h)


class Ref(_Checker):
    prefix = 'ref_'
    path = REF_SO


# ---- the reference's golden light curves (python/test/data/*.dat) ----
def load_golden(path):
    """Parse a freddi.dat golden: (kwargs in CLI units with __cgs=False, lambdas [cm], passband files, columns dict).

    Same header handling as python/test/test_regression.py:43-60,75-82: the '### Parameters' block is an INI
    echo; dir/prefix/fulldata/precision/config/tempsparsity/stdout/windtype/passband are dropped.
    """
    with open(path) as f:
        content = f.read()
    block = re.search(r'### Parameters.+?###', content, flags=re.MULTILINE | re.DOTALL).group()
    kwargs, lambdas, passbands = {}, [], []
    for line in block.splitlines():
        m = re.match(r'^# (\w+)=(.*)$', line)
        if not m:
            continue
        key, value = m.group(1), m.group(2).split('#')[0].strip()
        if key == 'lambda':
            lambdas.append(float(value) * 1e-8)
            continue
        if key == 'passband':
            passbands.append(value)
            continue
        if key in ('dir', 'prefix', 'fulldata', 'precision', 'config', 'tempsparsity', 'stdout', 'windtype'):
            continue
        try:
            kwargs[key] = int(value)
        except ValueError:
            try:
                kwargs[key] = float(value)
            except ValueError:
                kwargs[key] = value
    kwargs['__cgs'] = False
    data = np.genfromtxt(path, names=True)
    return kwargs, lambdas, passbands, data
