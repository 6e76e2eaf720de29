"""Ensemble of independent Freddi / FreddiNeutronStar models evolved on one B200 (thin layer over the C ABI)."""
import ctypes as C

import numpy as np

from . import _lib
from .args import FB200Args, FB200Config, FB200ModelInfo, FB200_MAX_LAMBDAS, FB200_MAX_PASSBANDS, args_array

N_BH_COLS = 15
N_NS_COLS = 10
# cpp/src/output.cpp:197-214 and cpp/src/ns/ns_output.cpp:6-19
BH_COLS = ['t', 'Mdot', 'Mdisk', 'Rhot', 'Sigmaout', 'Kirrout', 'H2R', 'Teffout', 'Tirrout', 'Qirr2Qvisout',
           'TphXmax', 'Lx', 'Lbol', 'Fx', 'Fbol']
NS_COLS = ['Rin', 'etans', 'Lxns', 'Lbolns', 'Fxns', 'Fbolns', 'Thotspot', 'fpin', 'Fmagnin', 'Fin']
FIELDS = {'h': 0, 'R': 1, 'F': 2, 'W': 3, 'Sigma': 4, 'Tph': 5, 'Tph_vis': 6, 'Tirr': 7, 'Kirr': 8, 'Height': 9,
          'Qx': 10, 'Tph_X': 11, 'windA': 12, 'windB': 13, 'windC': 14, 'Fmagn': 15, 'dFmagn_dh': 16,
          'd2Fmagn_dh2': 17}


class Passband:
    """One tabulated passband as the reference's ``EnergyPassband`` holds it (cpp/include/passband.hpp:26-47):
    ``lambdas`` [cm] and ``transmissions`` (for a photon counter already multiplied by the wavelength)."""

    def __init__(self, name, lambdas, transmissions):
        self.name = str(name)
        self.lambdas = np.ascontiguousarray(lambdas, dtype=np.float64).ravel()
        self.transmissions = np.ascontiguousarray(transmissions, dtype=np.float64).ravel()
        if self.lambdas.shape != self.transmissions.shape:
            raise ValueError('passband {}: lambdas and transmissions differ in length'.format(name))

    @classmethod
    def from_file(cls, path, detector='photon-counter'):
        """Two whitespace-separated columns, wavelength [Angstrom] and transmission, read until the first token that
        is not a number (EnergyPassband::dataFromFile, cpp/src/passband.cpp:16-43).  The CLI always builds
        photon-counter passbands (cpp/src/options.cpp:316): transmission *= lambda[cm]."""
        import os
        vals = []
        with open(path) as f:
            for tok in f.read().split():
                try:
                    vals.append(float(tok))
                except ValueError:
                    break
        vals = vals[:len(vals) // 2 * 2]
        lam = np.array(vals[0::2]) * 1e-8  # angstromToCm, cpp/include/unit_transformation.hpp:30-32
        tr = np.array(vals[1::2])
        if detector == 'photon-counter':
            tr = tr * lam
        elif detector != 'energy':
            raise ValueError('detector must be "photon-counter" or "energy"')
        name = str(path)  # EnergyPassband::nameFromPath: `for (; !path.extension().empty(); path = path.stem());`
        while os.path.splitext(os.path.basename(name))[1]:
            name = os.path.splitext(os.path.basename(name))[0]
        return cls(name, lam, tr)


def column_names(n_lambdas=0, cold_disk=False, is_ns=False, passbands=(), star_flux=False):
    cols = list(BH_COLS)
    for name in [str(k) for k in range(n_lambdas)] + list(passbands):  # cpp/src/output.cpp:219-291
        cols.append('Fnu{}'.format(name))
        if cold_disk:
            cols.append('Fnu{}_cold'.format(name))
        if star_flux:
            cols += ['Fnu{}_star'.format(name), 'Fnu{}_star_min'.format(name), 'Fnu{}_star_max'.format(name)]
    if is_ns:
        cols += NS_COLS
    return cols


def resolve(args):
    """Host-only: (info, h, F0) of one model — what the reference's constructors compute."""
    info = FB200ModelInfo()
    h = np.empty(int(args.Nx))
    F = np.empty(int(args.Nx))
    dp = C.POINTER(C.c_double)
    _lib.check(_lib.lib.fb200_args_resolve(C.byref(args), C.byref(info), h.ctypes.data_as(dp), F.ctypes.data_as(dp)))
    return info, h, F


def make_config(lambdas=(), cold_disk=False, record_F=False, compute_lx=True, host_threads=0, exact_lx_walk=False, passbands=(),
                star_flux=False):
    """fb200_config_t of these options -> (cfg, lambdas, passbands); the passband arrays stay referenced by the result."""
    cfg = FB200Config()
    _lib.lib.fb200_config_default(C.byref(cfg))
    lambdas = [float(x) for x in lambdas]
    if len(lambdas) > FB200_MAX_LAMBDAS:
        raise ValueError('at most {} lambdas per ensemble'.format(FB200_MAX_LAMBDAS))
    cfg.n_lambdas = len(lambdas)
    for k, x in enumerate(lambdas):
        cfg.lambdas[k] = x
    cfg.cold_disk = int(bool(cold_disk))
    cfg.record_F = int(bool(record_F))
    cfg.compute_lx = int(bool(compute_lx))
    cfg.host_threads = int(host_threads)
    cfg.exact_lx_walk = int(bool(exact_lx_walk))
    passbands = list(passbands)
    if len(passbands) > FB200_MAX_PASSBANDS:
        raise ValueError('at most {} passbands per ensemble'.format(FB200_MAX_PASSBANDS))
    cfg.star_flux = 2 if star_flux == 'readout' else int(bool(star_flux))
    cfg.n_passbands = len(passbands)
    dp = C.POINTER(C.c_double)
    for k, pb in enumerate(passbands):
        cfg.passband_len[k] = pb.lambdas.size
        cfg.passband_lambdas[k] = pb.lambdas.ctypes.data_as(dp)
        cfg.passband_transmissions[k] = pb.transmissions.ctypes.data_as(dp)
    return cfg, lambdas, passbands


class _PinnedBlock:
    def __init__(self, nbytes):
        p = C.c_void_p()
        _lib.check(_lib.lib.fb200_host_alloc(int(nbytes), C.byref(p)))
        self.ptr = p.value
        self.nbytes = int(nbytes)
        self.__array_interface__ = {'shape': (self.nbytes,), 'typestr': '|u1', 'data': (self.ptr, False), 'version': 3}

    def __del__(self):
        try:
            if self.ptr:
                _lib.lib.fb200_host_free(C.c_void_p(self.ptr))
                self.ptr = None
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float64):
    """numpy array in pinned (page-locked) host memory from fb200_host_alloc: device <-> host copies into it run at full
    PCIe rate and asynchronously.  Freed with the last reference."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    blk = _PinnedBlock(max(n, 1))
    return np.asarray(blk)[:n].view(dtype).reshape(shape)  # the array's base chain keeps the block alive


class SweepResult:
    """What freddi_b200.sweep returns: rows (n_models, n_rows, n_cols), status / rows_valid / picard_iters (n_models,), the column
    names and the timing the library measured (stats: dict of fb200_sweep_stats_t)."""

    def __init__(self, rows, status, rows_valid, picard_iters, columns, stats, t_end=None):
        self.rows, self.status, self.rows_valid, self.picard_iters, self.columns, self.stats = rows, status, rows_valid, picard_iters, columns, stats
        self.t_end = t_end  # FreddiState::t() where each model stopped [s]


def sweep_plan(args, lambdas=(), cold_disk=False, passbands=(), star_flux=False):
    """(n_rows, n_cols) of the rows a sweep over these models returns (host only)."""
    arr = args if isinstance(args, C.Array) else args_array(list(args))
    cfg, _, _ = make_config(lambdas, cold_disk, False, True, 0, False, passbands, star_flux)
    nr, nc = C.c_size_t(), C.c_size_t()
    _lib.check(_lib.lib.fb200_sweep_plan(arr, len(arr), C.byref(cfg), C.byref(nr), C.byref(nc)))
    return int(nr.value), int(nc.value)


def sweep(args, lambdas=(), cold_disk=False, compute_lx=True, devices=None, host_threads=0, exact_lx_walk=False, passbands=(),
          star_flux=False, out=None):
    """Argument structs in, light curves out, in one call over one or several GPUs of this process (fb200_sweep_run): the
    reference's driver loop (cpp/include/application.hpp:17-43) for every model.  Construction is streamed behind the evolve
    kernels.  devices: None = all visible, or a sequence of ordinals; out: a (n_models, n_rows, n_cols) float64 array to fill
    (pinned_empty(...) for full-rate copies), allocated pinned when None."""
    arr = args if isinstance(args, C.Array) else args_array(list(args))
    cfg, lambdas, passbands = make_config(lambdas, cold_disk, False, compute_lx, host_threads, exact_lx_walk, passbands, star_flux)
    nr, nc = C.c_size_t(), C.c_size_t()
    _lib.check(_lib.lib.fb200_sweep_plan(arr, len(arr), C.byref(cfg), C.byref(nr), C.byref(nc)))
    n = len(arr)
    shape = (n, int(nr.value), int(nc.value))
    if out is None:
        out = pinned_empty(shape)
    if out.shape != shape or out.dtype != np.float64 or not out.flags.c_contiguous:
        raise ValueError('out must be a C-contiguous float64 array of shape {}'.format(shape))
    status = np.empty(n, dtype=np.int32)
    valid = np.empty(n, dtype=np.int32)
    picard = np.empty(n, dtype=np.int64)
    t_end = np.empty(n)
    devs = None if devices is None else np.ascontiguousarray(list(devices), dtype=np.int32)
    stats = _lib.FB200SweepStats()
    bad = C.c_size_t()
    i32p = C.POINTER(C.c_int32)
    rc = _lib.lib.fb200_sweep_run(arr, n, C.byref(cfg), None if devs is None else devs.ctypes.data_as(i32p), 0 if devs is None else devs.size,
                                  out.ctypes.data_as(C.POINTER(C.c_double)), out.size, status.ctypes.data_as(i32p), valid.ctypes.data_as(i32p),
                                  picard.ctypes.data_as(C.POINTER(C.c_int64)), t_end.ctypes.data_as(C.POINTER(C.c_double)), C.byref(stats),
                                  C.byref(bad))
    if rc != _lib.FB200_OK:
        try:
            _lib.check(rc)
        except Exception as e:
            e.bad_model = int(bad.value)
            raise
    columns = column_names(len(lambdas), cold_disk, bool(arr[0].is_ns), [pb.name for pb in passbands], bool(star_flux) and star_flux != 'readout')
    return SweepResult(out, status, valid, picard, columns, {k: getattr(stats, k) for k, _ in _lib.FB200SweepStats._fields_}, t_end)


class Ensemble:
    """``n`` models, one CTA each, evolved by a persistent sm_100a kernel.

    args      : sequence of FB200Args (freddi_b200.args.make_args) or a ctypes array of them
    lambdas   : wavelengths [cm] of the Fnu columns every row carries (FluxArguments::lambdas)
    passbands : sequence of Passband: one Fnu<name> column each (FluxArguments::passbands)
    cold_disk : add the Fnu_cold columns (--colddiskflux)
    star_flux : add the Fnu_star, _star_min, _star_max columns of the irradiated companion star (--starflux);
                'readout': only triangulate the star so that flux_rows(region='star') works
    record_F  : keep F(h) of every row in HBM (needed for field()/flux() at past rows)
    device    : a CUDA ordinal (one GPU), or 'all' / a sequence of ordinals: ONE handle over several GPUs of this process —
                model k on shard k mod G, all devices evolve at once, rows() gathers into one array (fb200_sharded_*);
                the per-device read-outs (field_rows, flux_rows, mdot_wind, ...) are reached through shards()
    exact_lx_walk : integrate the Lx band integral of every ring like the reference (default: skip the rings that are
                    provably negligible against the hottest one in the X-ray band; changes Lx by < 1e-16 relative)
    """

    def __init__(self, args, lambdas=(), cold_disk=False, record_F=False, compute_lx=True, device=0, host_threads=0,
                 exact_lx_walk=False, passbands=(), star_flux=False, _borrowed=None):
        self._h = None
        self._owner = True
        arr = args if isinstance(args, C.Array) else args_array(list(args))
        cfg, lambdas, passbands = make_config(lambdas, cold_disk, record_F, compute_lx, host_threads, exact_lx_walk, passbands, star_flux)
        handle = C.c_void_p()
        bad = C.c_size_t()
        # device: an ordinal -> one GPU; 'all' or a sequence of ordinals -> one handle over several GPUs of this process
        # (model k on shard k mod G; a device may be listed more than once)
        self._sharded = not isinstance(device, int)
        self._p = 'fb200_sharded_' if self._sharded else 'fb200_ensemble_'
        if _borrowed is not None:  # a shard of a sharded ensemble: the parent owns the handle
            handle, self._owner = _borrowed, False
        elif self._sharded:
            devs = None if (isinstance(device, str) and device == 'all') else np.ascontiguousarray(list(device), dtype=np.int32)
            rc = _lib.lib.fb200_sharded_create(arr, len(arr), C.byref(cfg), None if devs is None else devs.ctypes.data_as(C.POINTER(C.c_int32)),
                                               0 if devs is None else devs.size, C.byref(handle), C.byref(bad))
            self.bad_model = int(bad.value)
            _lib.check(rc)
        else:
            cfg.device = int(device)
            rc = _lib.lib.fb200_ensemble_create(arr, len(arr), C.byref(cfg), C.byref(handle), C.byref(bad))
            self.bad_model = int(bad.value)
            _lib.check(rc)
        self._h = handle
        self._shard_views = None
        self._args = arr
        self.lambdas = lambdas
        self.cold_disk = bool(cold_disk)
        self.record_F = bool(record_F)
        self.is_ns = bool(arr[0].is_ns)
        self.n_models = int(self._f('n_models')(handle))
        self.n_rows = int(self._f('n_rows')(handle))
        self.n_cols = int(self._f('n_cols')(handle))
        self.Nx = int(self._f('nx')(handle))
        self.passbands = passbands
        self.star_flux = bool(star_flux) and star_flux != 'readout'
        self.columns = column_names(len(lambdas), cold_disk, self.is_ns, [pb.name for pb in passbands], self.star_flux)
        self._info = None
        self._kw = dict(lambdas=lambdas, cold_disk=cold_disk, record_F=record_F, compute_lx=compute_lx, exact_lx_walk=exact_lx_walk,
                        passbands=passbands, star_flux=star_flux)

    def _f(self, name):
        """The C entry point `name` of this handle kind (fb200_ensemble_* / fb200_sharded_*)."""
        f = getattr(_lib.lib, self._p + name, None)
        if f is None:
            raise NotImplementedError('{} is per device: call it on the shards (Ensemble.shards())'.format(name))
        return f

    def _gather(self, method, *a, **kw):
        """A per-device read-out on a multi-GPU handle: call it on every shard and interleave the results into the caller's
        model order (model k = shard k mod G, local index k // G).  Arrays, tuples and dicts of arrays are handled."""
        if self._shard_views is None:
            self._shard_views = self.shards()
        parts = [getattr(sh, method)(*a, **kw) for sh in self._shard_views]
        G = len(parts)

        def weave(xs):
            if isinstance(xs[0], tuple):
                return tuple(weave([x[j] for x in xs]) for j in range(len(xs[0])))
            if isinstance(xs[0], dict):
                return {k: weave([x[k] for x in xs]) for k in xs[0]}
            if xs[0] is None:
                return None
            out = np.empty((self.n_models,) + xs[0].shape[1:], dtype=xs[0].dtype)
            for d, x in enumerate(xs):
                out[d::G] = x
            return out
        return weave(parts)

    @property
    def n_shards(self):
        return int(_lib.lib.fb200_sharded_n_shards(self._h)) if self._sharded else 1

    def shards(self):
        """The single-device ensembles behind a multi-GPU handle (views: the parent keeps ownership).  Shard d holds the
        models d, d + G, d + 2G, ... of the caller."""
        if not self._sharded:
            return [self]
        G = self.n_shards
        out = []
        for d in range(G):
            h = C.c_void_p(_lib.lib.fb200_sharded_shard(self._h, d))
            sub = (FB200Args * len(range(d, self.n_models, G)))(*[self._args[k] for k in range(d, self.n_models, G)])
            out.append(Ensemble(sub, _borrowed=h, **self._kw))
        return out

    # -- life cycle --
    def close(self):
        if self._h is not None:
            if self._owner:
                self._f('destroy')(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- evolution --
    def evolve(self, n_steps=-1):
        _lib.check(self._f('evolve')(self._h, int(n_steps)))
        return self

    def reset(self):
        """Back to the initial state: the resolved initial conditions are copied again host (pinned) -> device."""
        _lib.check(self._f('reset')(self._h))
        return self

    def evolve_async(self, n_steps=-1):
        _lib.check(self._f('evolve_async')(self._h, int(n_steps)))
        return self

    def sync(self):
        _lib.check(self._f('sync')(self._h))

    def last_evolve_ms(self):
        ms = C.c_float()
        _lib.check(self._f('last_evolve_ms')(self._h, C.byref(ms)))
        return float(ms.value)

    def last_launches(self):
        if self._sharded:
            return sum(sh.last_launches() for sh in self.shards())
        return int(self._f('last_launches')(self._h))

    # -- read-out --
    def rows(self, out=None):
        """(n_models, n_rows, n_cols) float64; rows a model never reached are NaN."""
        if out is None:
            out = np.empty((self.n_models, self.n_rows, self.n_cols))
        _lib.check(self._f('rows')(self._h, out.ctypes.data_as(C.POINTER(C.c_double)), out.size))
        return out

    def rows_device_ptr(self):
        p = C.c_void_p()
        _lib.check(self._f('rows_device')(self._h, C.byref(p)))
        return p.value

    # -- device-resident results (zero copy) --
    _DEVPTR = {'rows': 0, 'F': 1, 'h': 2, 'Fsnap': 3, 'bounds': 4}

    def device_array(self, which):
        """A view of the ensemble's HBM as an object with ``__cuda_array_interface__`` (version 3): ``torch.as_tensor(x,
        device='cuda')``, ``cupy.asarray(x)`` and numba take it without a copy.  which: 'rows' (n_models, n_rows, n_cols),
        'F' / 'h' (n_models, Nx), with record_F 'Fsnap' (n_models, n_rows, Nx; NaN where never written) and 'bounds'
        (n_models, n_rows, 2) int32.  The view keeps the ensemble alive; call sync() after evolve_async() before reading."""
        if self._sharded:
            raise NotImplementedError('device arrays are per device: use Ensemble.shards()')
        p, dev = C.c_void_p(), C.c_int32()
        _lib.check(_lib.lib.fb200_ensemble_device_ptr(self._h, self._DEVPTR[which], C.byref(p), C.byref(dev)))
        shape = {'rows': (self.n_models, self.n_rows, self.n_cols), 'F': (self.n_models, self.Nx), 'h': (self.n_models, self.Nx),
                 'Fsnap': (self.n_models, self.n_rows, self.Nx), 'bounds': (self.n_models, self.n_rows, 2)}[which]
        return DeviceArray(self, p.value, shape, '<i4' if which == 'bounds' else '<f8', int(dev.value))

    def rows_device(self):
        """(n_models, n_rows, n_cols) float64 in HBM, zero copy (see device_array)."""
        return self.device_array('rows')

    # -- state in / out --
    def checkpoint(self, with_rows=True):
        """Opaque blob (numpy uint8) of every model's current state — CurrentState, F and the carried closure, optionally the
        rows so far; restore() into an ensemble created from the same arguments continues bit-identically."""
        if self._sharded:
            return [sh.checkpoint(with_rows) for sh in self.shards()]
        n = int(_lib.lib.fb200_ensemble_checkpoint_bytes(self._h, int(bool(with_rows))))
        buf = np.empty(n, dtype=np.uint8)
        _lib.check(_lib.lib.fb200_ensemble_checkpoint_save(self._h, buf.ctypes.data_as(C.c_void_p), n, int(bool(with_rows))))
        return buf

    def restore(self, blob):
        if self._sharded:
            for sh, b in zip(self.shards(), blob):
                sh.restore(b)
            return self
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        _lib.check(_lib.lib.fb200_ensemble_checkpoint_load(self._h, blob.ctypes.data_as(C.c_void_p), blob.size))
        return self

    def get_state(self, with_F=True):
        """dict(F (n_models, Nx) | None, t, i_t, first, last, F_in, Mdot_in_prev): FreddiState::CurrentState per model."""
        if self._sharded:
            return self._gather('get_state', with_F)
        n = self.n_models
        F = np.empty((n, self.Nx)) if with_F else None
        t, F_in, Mp = (np.empty(n) for _ in range(3))
        i_t, first, last = (np.empty(n, dtype=np.int64) for _ in range(3))
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int64)
        _lib.check(self._f('get_state')(self._h, F.ctypes.data_as(dp) if with_F else None, t.ctypes.data_as(dp), i_t.ctypes.data_as(ip),
                                        first.ctypes.data_as(ip), last.ctypes.data_as(ip), F_in.ctypes.data_as(dp), Mp.ctypes.data_as(dp)))
        return dict(F=F, t=t, i_t=i_t, first=first, last=last, F_in=F_in, Mdot_in_prev=Mp)

    def set_state(self, F=None, t=None, i_t=None, first=None, last=None, F_in=None, Mdot_in_prev=None):
        """Replace (parts of) the models' current state: F (n_models, Nx) and per-model scalars; None keeps a component."""
        n = self.n_models
        if self._sharded:
            G = self.n_shards
            full = dict(F=F, t=t, i_t=i_t, first=first, last=last, F_in=F_in, Mdot_in_prev=Mdot_in_prev)
            for d, sh in enumerate(self.shards()):
                part = {}
                for k, v in full.items():
                    if v is None:
                        continue
                    v = np.asarray(v)
                    part[k] = v if v.ndim == 0 or v.shape[0] != n else v[d::G]
                sh.set_state(**part)
            return self
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int64)

        def dbl(x, shape):
            if x is None:
                return None, None
            a = np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=np.float64), shape))
            return a, a.ctypes.data_as(dp)

        def i64(x):
            if x is None:
                return None, None
            a = np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=np.int64), (n,)))
            return a, a.ctypes.data_as(ip)
        keep = [dbl(F, (n, self.Nx)), dbl(t, (n,)), i64(i_t), i64(first), i64(last), dbl(F_in, (n,)), dbl(Mdot_in_prev, (n,))]
        _lib.check(self._f('set_state')(self._h, *[k[1] for k in keep]))
        return self

    def result(self):
        """EnsembleResult over all recorded rows (needs record_F for the radial arrays)."""
        return EnsembleResult(self)

    def status(self):
        st = np.empty(self.n_models, dtype=np.int32)
        rv = np.empty(self.n_models, dtype=np.int32)
        pi = np.empty(self.n_models, dtype=np.int64)
        _lib.check(self._f('status')(self._h, st.ctypes.data_as(C.POINTER(C.c_int32)),
                                                  rv.ctypes.data_as(C.POINTER(C.c_int32)),
                                                  pi.ctypes.data_as(C.POINTER(C.c_int64))))
        return st, rv, pi

    def counters(self):
        """(picard_iters, lx_trial_steps) per model, summed over the steps done so far."""
        if self._sharded:
            return self._gather('counters')
        pi, lx = (np.empty(self.n_models, dtype=np.int64) for _ in range(2))
        ip = C.POINTER(C.c_int64)
        _lib.check(self._f('counters')(self._h, pi.ctypes.data_as(ip), lx.ctypes.data_as(ip)))
        return pi, lx

    WORK_COUNTERS = ('picard_iters', 'lx_trials', 'ring_steps', 'ring_iters', 'pow_evals', 'row_rings')

    def work_counters(self):
        """dict of per-model int64 arrays: what the kernels really did so far (roofline bookkeeping) — Picard iterations,
        try_step calls of the Lx quadrature, sum of (last - first) over the solves, the same weighted by the Picard
        iterations, closure pow() calls."""
        out = np.empty((self.n_models, len(self.WORK_COUNTERS)), dtype=np.int64)
        _lib.check(self._f('work_counters')(self._h, out.ctypes.data_as(C.POINTER(C.c_int64)), out.size))
        return {k: out[:, j].copy() for j, k in enumerate(self.WORK_COUNTERS)}

    def state(self):
        t = np.empty(self.n_models)
        i_t, first, last = (np.empty(self.n_models, dtype=np.int64) for _ in range(3))
        ip = C.POINTER(C.c_int64)
        _lib.check(self._f('state')(self._h, t.ctypes.data_as(C.POINTER(C.c_double)), i_t.ctypes.data_as(ip),
                                                 first.ctypes.data_as(ip), last.ctypes.data_as(ip)))
        return dict(t=t, i_t=i_t, first=first, last=last)

    def row_bounds(self, row):
        if self._sharded:
            return self._gather('row_bounds', row)
        first, last = (np.empty(self.n_models, dtype=np.int64) for _ in range(2))
        ip = C.POINTER(C.c_int64)
        _lib.check(self._f('row_bounds')(self._h, int(row), first.ctypes.data_as(ip), last.ctypes.data_as(ip)))
        return first, last

    def info(self):
        if self._info is None:
            arr = (FB200ModelInfo * self.n_models)()
            _lib.check(self._f('info')(self._h, arr))
            self._info = arr
        return self._info

    def field(self, name, row=-1):
        """(n_models, Nx): the reference getter's array (zeros below first, cold-zone values beyond last)."""
        out = np.empty((self.n_models, self.Nx))
        _lib.check(self._f('field')(self._h, FIELDS[name], int(row), out.ctypes.data_as(C.POINTER(C.c_double)),
                                                 out.size))
        return out

    def flux(self, lambdas, region='hot', row=-1):
        """(n_models, len(lambdas)) spectral flux density of the hot or cold disc region."""
        lam = np.ascontiguousarray(np.atleast_1d(lambdas), dtype=np.float64).ravel()
        out = np.empty((self.n_models, lam.size))
        reg = {'hot': 0, 'cold': 1}[region]
        dp = C.POINTER(C.c_double)
        _lib.check(self._f('flux')(self._h, reg, int(row), lam.ctypes.data_as(dp), lam.size,
                                                out.ctypes.data_as(dp), out.size))
        return out

    def field_rows(self, name, row_begin=0, n_rows=None, pad_nan=True):
        """(n_models, n_rows, Nx) in one launch; pad_nan: NaN outside [first, last] (the EvolutionResult layout)."""
        if self._sharded:
            return self._gather('field_rows', name, row_begin, n_rows, pad_nan)
        n_rows = self.n_rows - row_begin if n_rows is None else int(n_rows)
        out = np.empty((self.n_models, n_rows, self.Nx))
        _lib.check(self._f('field_rows')(self._h, FIELDS[name], int(row_begin), n_rows, int(bool(pad_nan)),
                                                      out.ctypes.data_as(C.POINTER(C.c_double)), out.size))
        return out

    def flux_rows(self, lambdas, region='hot', row_begin=0, n_rows=None, phases=None):
        """(n_models, n_rows, len(lambdas)) in one launch; region 'star' needs one orbital phase [rad] per row."""
        if self._sharded:
            return self._gather('flux_rows', lambdas, region, row_begin, n_rows, phases)
        n_rows = self.n_rows - row_begin if n_rows is None else int(n_rows)
        lam = np.ascontiguousarray(np.atleast_1d(lambdas), dtype=np.float64).ravel()
        out = np.empty((self.n_models, n_rows, lam.size))
        reg = {'hot': 0, 'cold': 1, 'star': 2}[region]
        dp = C.POINTER(C.c_double)
        ph = None
        if phases is not None:
            ph = np.ascontiguousarray(np.broadcast_to(np.asarray(phases, dtype=np.float64), (n_rows,)))
        _lib.check(self._f('flux_rows')(self._h, reg, int(row_begin), n_rows, lam.ctypes.data_as(dp), lam.size,
                                                     ph.ctypes.data_as(dp) if ph is not None else None,
                                                     out.ctypes.data_as(dp), out.size))
        return out

    def mdot_wind(self, row=-1):
        if self._sharded:
            return self._gather('mdot_wind', row)
        out = np.empty(self.n_models)
        _lib.check(self._f('mdot_wind')(self._h, int(row), out.ctypes.data_as(C.POINTER(C.c_double)), out.size))
        return out


class DeviceArray:
    """A typed view of device memory owned by an Ensemble (``__cuda_array_interface__`` version 3)."""

    def __init__(self, owner, ptr, shape, typestr, device):
        self._owner = owner  # keeps the HBM alive
        self.shape, self.device = tuple(int(x) for x in shape), device
        # (torch only imports writable views; the memory belongs to the engine: treat it as read-only)
        self.__cuda_array_interface__ = {'shape': self.shape, 'typestr': typestr, 'data': (int(ptr), False), 'version': 3, 'strides': None}

    def to_torch(self):
        import torch
        return torch.as_tensor(self, device=torch.device('cuda', self.device))


class EnsembleResult:
    """The reference's EvolutionResult (python/freddi/evolution_result.py:24-71) for a whole ensemble: attribute access gives
    the light-curve columns as (n_models, n_rows) and the radial distributions as (n_models, n_rows, Nx), NaN outside
    [first, last] and for rows a model never reached — one launch per attribute (fb200_ensemble_field_rows)."""

    def __init__(self, ens):
        self._ens = ens
        self._rows = None

    def _col(self, name):
        if self._rows is None:
            self._rows = self._ens.rows()
        return self._rows[:, :, self._ens.columns.index(name)]

    def flux(self, lmbd, region='hot', phase=None):
        """(n_models, n_rows) + lmbd.shape, like EvolutionResult.flux."""
        lam = np.asarray(lmbd, dtype=float)
        e = self._ens
        parts = {'hot': ['hot'], 'cold': ['cold'], 'disk': ['hot', 'cold'], 'star': ['star'], 'all': ['hot', 'cold', 'star']}[region]
        total = 0.
        for part in parts:
            ph = None
            if part == 'star':
                if phase is None:
                    raise ValueError('Phase must be specified if star flux is required')
                ph = phase
            total = total + e.flux_rows(lam.ravel(), part, 0, e.n_rows, ph)
        return total.reshape((e.n_models, e.n_rows) + lam.shape)

    def __getattr__(self, attr):
        e = self.__dict__['_ens']
        if attr in e.columns:
            return self._col(attr)
        alias = {'Mdot_in': 'Mdot', 'Fx': 'Fx', 'Lx': 'Lx'}
        if attr in alias and alias[attr] in e.columns:
            return self._col(alias[attr])
        if attr in FIELDS:
            return e.field_rows(attr, 0, e.n_rows, pad_nan=True)
        raise AttributeError(attr)


def measure_fp64_peak(device=0):
    """Measured FP64 (DFMA) peak of the device in TFLOP/s."""
    v = C.c_double()
    _lib.check(_lib.lib.fb200_measure_fp64_peak(int(device), C.byref(v)))
    return float(v.value)


from .sharding import shard_indices  # noqa: E402,F401
