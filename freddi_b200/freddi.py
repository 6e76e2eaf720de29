"""``Freddi`` / ``FreddiNeutronStar`` / ``EvolutionResult`` with the reference's Python surface
(python/freddi/__init__.py:9-150, python/freddi/evolution_result.py:6-71, cpp/pywrap/pywrap_freddi_state.cpp:18-52,
cpp/pywrap/pywrap_freddi_evolution.cpp:303-325), implemented over the C ABI: one model = an ensemble of one,
evolved by a single persistent-kernel launch with F(h) of every row recorded in HBM.

``EvolutionResult`` serves its (Nt+1, Nx) arrays and ``flux()`` from ONE batched read-out launch over all recorded rows
(``fb200_ensemble_field_rows`` / ``fb200_ensemble_flux_rows``) instead of one call per state.
"""
import numpy as np

from . import _lib
from .args import merge_kwargs, kwargs_to_args, REQUIRED_ARGS, NS_REQUIRED_ARGS
from .ensemble import Ensemble, BH_COLS, NS_COLS

_ARRAY_PROPS = ('h', 'R', 'F', 'W', 'Tph', 'Tph_vis', 'Tirr', 'Kirr', 'Sigma', 'Height', 'windA', 'windB', 'windC')
_NS_ARRAY_PROPS = ('Fmagn', 'dFmagn_dh', 'd2Fmagn_dh2')
_DAY = 86400.0
_KEV_PER_K = 1.3806504e-16 / (1000. * 1.602176487e-12)


class _Run:
    """The evolved model shared by all states of one Freddi object."""

    def __init__(self, args, is_ns, device):
        self.args = args
        self.is_ns = is_ns
        self.ens = Ensemble([args], record_F=True, device=device, star_flux='readout')
        self.ens.evolve(0)  # row 0 = the initial state
        self.info = self.ens.info()[0]
        self.evolved = False
        self._rows = None
        self._cache = {}

    def ensure(self, row):
        if row > 0 and not self.evolved:
            self.ens.evolve(-1)
            self.evolved = True
            self._rows = None
            self._cache.clear()

    @property
    def rows(self):
        if self._rows is None:
            self._rows = self.ens.rows()[0]
        return self._rows

    def rows_valid(self):
        return int(self.ens.status()[1][0])

    def collapsed(self):
        return int(self.ens.status()[0][0]) == _lib.MODEL_COLLAPSED

    def col(self, row, name):
        return float(self.rows[row, self.ens.columns.index(name)])

    def cached(self, key, fn):
        if key not in self._cache:
            self._cache[key] = fn()
        return self._cache[key]


class _BaseFreddi:
    _is_ns = False

    def __init__(self, *args, **kwargs):
        if args:
            # cpp/pywrap/pywrap_freddi_evolution.cpp:86-91
            raise TypeError('{} constructor accepts keyword arguments only'.format(type(self).__name__))
        device = kwargs.pop('__device', 0)
        kw = merge_kwargs(kwargs, self._is_ns, type(self).__name__)
        self._kwargs = dict(kw)
        self._run = _Run(kwargs_to_args(kw, self._is_ns), self._is_ns, device)
        self._row = 0

    @classmethod
    def _required_args(cls):
        return dict(NS_REQUIRED_ARGS if cls._is_ns else REQUIRED_ARGS)

    @classmethod
    def _at(cls, run, row, kwargs):
        obj = cls.__new__(cls)
        obj._run = run
        obj._row = row
        obj._kwargs = kwargs
        return obj

    @classmethod
    def from_astropy(cls, **kwargs):
        """Alternative constructor accepting astropy Quantity values (python/freddi/__init__.py:51-86)."""
        from collections.abc import Mapping
        from functools import singledispatch
        from astropy.units import Quantity

        @singledispatch
        def convert(value):
            return value

        @convert.register(Quantity)
        def _(value):
            return value.cgs.value

        @convert.register(Mapping)
        def _(value):
            return {k: convert(v) for k, v in value.items()}

        return cls(**{key: convert(value) for key, value in kwargs.items()})

    alt = from_astropy

    # ---- iteration / evolution (cpp/include/freddi_evolution.hpp:34-39,56-57) ----
    def __iter__(self):
        run = self._run
        run.ensure(1)
        n = run.rows_valid()
        for r in range(self._row, n):
            self._row = r
            yield type(self)._at(run, r, self._kwargs)
        if run.collapsed():
            raise RuntimeError('Rout <= Rin')  # RadiusCollapseException::what, cpp/include/exceptions.hpp:6-11

    def evolve(self):
        return EvolutionResult(self)

    # ---- scalars ----
    def _col(self, name):
        self._run.ensure(self._row)
        return self._run.col(self._row, name)

    GM = property(lambda self: self._run.info.GM)
    eta = property(lambda self: self._run.info.eta)
    distance = property(lambda self: self._run.info.distance)
    cosi = property(lambda self: self._run.info.cosi)
    Nt = property(lambda self: int(self._run.info.Nt))
    Nx = property(lambda self: int(self._run.info.Nx))
    Mdot_out = property(lambda self: float(self._run.args.Mdotout))
    Mdot_in = property(lambda self: self._col('Mdot'))
    Mdot = Mdot_in
    Lx = property(lambda self: self._col('Lx'))
    Fx = property(lambda self: self._col('Fx'))
    Lbol = property(lambda self: self._col('Lbol'))
    Mdisk = property(lambda self: self._col('Mdisk'))
    i_t = property(lambda self: self._row)
    lambdas = property(lambda self: np.empty(0))  # hard-wired empty from Python (pywrap_arguments.cpp:64-77)

    @property
    def t(self):
        t = float(self._run.args.inittime)
        tau = float(self._run.info.tau)
        for _ in range(self._row):  # accumulated like FreddiState::step (cpp/src/freddi_state.cpp:151)
            t += tau
        return t

    def _bounds(self):
        self._run.ensure(self._row)
        first, last = self._run.cached(('b', self._row), lambda: self._run.ens.row_bounds(self._row))
        return int(first[0]), int(last[0])

    first = property(lambda self: self._bounds()[0])
    last = property(lambda self: self._bounds()[1])

    @property
    def Mdot_wind(self):
        self._run.ensure(self._row)
        return float(self._run.ens.mdot_wind(self._row)[0])

    # ---- arrays ----
    def _field(self, name):
        self._run.ensure(self._row)
        return self._run.cached((name, self._row), lambda: self._run.ens.field(name, self._row)[0]).copy()

    # ---- fluxes (python/freddi/__init__.py:103-137) ----
    def _flux_hot(self, lmbd, phase=None):
        self._run.ensure(self._row)
        return self._run.ens.flux(np.atleast_1d(lmbd), 'hot', self._row)[0]

    def _flux_cold(self, lmbd, phase=None):
        self._run.ensure(self._row)
        return self._run.ens.flux(np.atleast_1d(lmbd), 'cold', self._row)[0]

    def _flux_star(self, lmbd, phase=None):
        self._run.ensure(self._row)
        return self._run.ens.flux_rows(np.atleast_1d(lmbd), 'star', self._row, 1, [float(phase)])[0, 0]

    def flux(self, lmbd, region='hot', phase=None):
        region = region.lower()
        lmbd = np.asarray(lmbd, dtype=float)
        flat = lmbd.ravel()
        if 'hot'.startswith(region):
            out = self._flux_hot(flat)
        elif 'cold'.startswith(region):
            out = self._flux_cold(flat)
        elif 'disk'.startswith(region):
            out = self._flux_hot(flat) + self._flux_cold(flat)
        elif 'star'.startswith(region):
            if phase is None:
                raise ValueError('Phase must be specified if star flux is required')
            out = self._flux_star(flat, phase)
        elif 'all'.startswith(region):
            if phase is None:
                raise ValueError('Phase must be specified if star flux is required')
            out = self._flux_hot(flat) + self._flux_cold(flat) + self._flux_star(flat, phase)
        else:
            raise ValueError(f'Zone {region} is not supported')
        return out.reshape(lmbd.shape)


def _array_prop(name):
    return property(lambda self: self._field(name))


def _edge_prop(edge, name):
    return property(lambda self: self._field(name)[getattr(self, edge)])


for _name in _ARRAY_PROPS:
    setattr(_BaseFreddi, _name, _array_prop(_name))
    setattr(_BaseFreddi, 'first_' + _name, _edge_prop('first', _name))
    setattr(_BaseFreddi, 'last_' + _name, _edge_prop('last', _name))


class Freddi(_BaseFreddi):
    """Black-hole disc (reference: freddi.Freddi)."""
    _is_ns = False


class FreddiNeutronStar(_BaseFreddi):
    """Neutron-star disc with a magnetosphere (reference: freddi.FreddiNeutronStar,
    cpp/pywrap/pywrap_freddi_evolution.cpp:313-325)."""
    _is_ns = True

    mu_magn = property(lambda self: self._run.info.mu_magn)
    R_cor = property(lambda self: self._run.info.R_cor)
    R_dead = property(lambda self: self._run.info.R_dead)
    inverse_beta = property(lambda self: self._run.info.inverse_beta)
    eta_ns = property(lambda self: self._col('etans'))
    fp = property(lambda self: self._col('fpin'))
    T_hot_spot = property(lambda self: self._col('Thotspot') / _KEV_PER_K)
    Lbol_ns = property(lambda self: self._col('Lbolns'))
    Lx_ns = property(lambda self: self._col('Lxns'))
    Fx_ns = property(lambda self: self._col('Fxns'))


for _name in _NS_ARRAY_PROPS:
    setattr(FreddiNeutronStar, _name, _array_prop(_name))
    setattr(FreddiNeutronStar, 'first_' + _name, _edge_prop('first', _name))
    setattr(FreddiNeutronStar, 'last_' + _name, _edge_prop('last', _name))


class EvolutionResult:
    """Temporal distribution of the disc parameters (python/freddi/evolution_result.py:6-71):
    attribute access stacks the per-state values into (Nt+1,) or (Nt+1, Nx) arrays, the latter NaN
    outside [first, last]."""

    def __init__(self, states):
        self.__states = tuple(states)
        self.__len_states = len(self.__states)

    def __batched(self):
        """(run, first row) when the states are consecutive recorded rows of one run, else None."""
        st = self.__states
        if not st or not all(isinstance(s, _BaseFreddi) for s in st):
            return None
        run, r0 = st[0]._run, st[0]._row
        if any(s._run is not run or s._row != r0 + i for i, s in enumerate(st)):
            return None
        run.ensure(r0 + len(st) - 1)
        return run, r0

    def flux(self, lmbd, region='hot', phase=None):
        from collections.abc import Iterable
        if isinstance(phase, Iterable):
            phase = np.asarray(phase)
            if phase.shape != (self.__len_states,):
                raise ValueError('phase length should be {}'.format(self.__len_states))
        else:
            phase = [phase] * self.__len_states
        lmbd = np.asarray(lmbd)
        b = self.__batched()
        reg = region.lower()
        if b is not None and reg and any(name.startswith(reg) for name in ('hot', 'cold', 'disk', 'star', 'all')):
            run, r0 = b
            n = self.__len_states
            flat = np.asarray(lmbd, dtype=float).ravel()
            parts = {'hot': ['hot'], 'cold': ['cold'], 'disk': ['hot', 'cold'], 'star': ['star'], 'all': ['hot', 'cold', 'star']}
            key = next(name for name in ('hot', 'cold', 'disk', 'star', 'all') if name.startswith(reg))
            if 'star' in parts[key] and any(p is None for p in phase):
                raise ValueError('Phase must be specified if star flux is required')
            total = 0.
            for part in parts[key]:  # one launch per region for all rows
                ph = np.asarray(phase, dtype=float) if part == 'star' else None
                total = total + run.ens.flux_rows(flat, part, r0, n, ph)[0]
            return total.reshape((n,) + lmbd.shape)
        arr = np.empty((self.__len_states,) + lmbd.shape, dtype=float)
        for i in range(self.__len_states):
            arr[i] = self.__states[i].flux(lmbd, region, phase[i])
        return arr

    def __getattr__(self, attr):
        if attr in _ARRAY_PROPS + _NS_ARRAY_PROPS:
            b = self.__batched()
            if b is not None:  # one launch for all rows, NaN padding done on the device
                run, r0 = b
                return run.ens.field_rows(attr, r0, self.__len_states, pad_nan=True)[0]
        first_val = np.asarray(getattr(self.__states[0], attr))
        if first_val.ndim == 0:
            arr = np.empty(self.__len_states, dtype=first_val.dtype)
            arr[0] = first_val
            for i in range(1, self.__len_states):
                arr[i] = getattr(self.__states[i], attr)
            return arr
        elif first_val.ndim == 1:
            arr = np.full((self.__len_states, self.__states[0].Nx), np.nan, dtype=first_val.dtype)
            for i, state in enumerate(self.__states):
                value = np.asarray(getattr(state, attr))
                idx = slice(state.first, state.last + 1)
                arr[i, idx] = value[idx]
            return arr
        else:
            raise ValueError("{}.ndim > 1 ({})".format(attr, first_val.ndim))
