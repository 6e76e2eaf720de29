"""Keyword arguments of ``freddi.Freddi`` / ``freddi.FreddiNeutronStar`` -> ``fb200_args_t``.

Mirrors the reference's Python constructor: names and defaults of
``cpp/pywrap/pywrap_freddi_evolution.cpp:15-83,194-224``, the ``__cgs=False`` unit conversion of
``:107-130`` (constants from ``cpp/include/unit_transformation.hpp:8-55`` and
``cpp/include/gsl_const_cgsm.h``), keyword-only / unknown-keyword errors of ``:86-104``.
String options become the enums of ``include/freddi_b200.h``; an unknown string is passed through as
an out-of-range enum so that the C++ resolver raises the same error class the reference does.
"""
import ctypes as C
import math

# ---- constants (cpp/include/gsl_const_cgsm.h, cpp/include/constants.hpp) ----
SPEED_OF_LIGHT = 2.99792458e10
GRAVITATIONAL_CONSTANT = 6.673e-8
PLANCKS_CONSTANT_H = 6.62606896e-27
PARSEC = 3.08567758135e18
ELECTRON_VOLT = 1.602176487e-12
SOLAR_MASS = 1.98892e33
SOLAR_RADIUS = 6.955e10
DAY = 86400.0


def kev_to_hertz(e_kev):
    return e_kev * 1000.0 * ELECTRON_VOLT / PLANCKS_CONSTANT_H


def rg_to_cm(r_rg, mass_gram):
    return r_rg * GRAVITATIONAL_CONSTANT * mass_gram / (SPEED_OF_LIGHT * SPEED_OF_LIGHT)


class FB200Args(C.Structure):
    """ctypes image of ``fb200_args_t`` (include/freddi_b200.h)."""
    _fields_ = [
        ('is_ns', C.c_int32), ('_pad0', C.c_int32),
        ('alpha', C.c_double), ('alphacold', C.c_double), ('Mx', C.c_double), ('kerr', C.c_double),
        ('period', C.c_double), ('Mopt', C.c_double), ('rochelobefill', C.c_double), ('Topt', C.c_double),
        ('rin', C.c_double), ('rout', C.c_double), ('risco', C.c_double),
        ('opacity', C.c_int32), ('boundcond', C.c_int32), ('initialcond', C.c_int32), ('windtype', C.c_int32),
        ('Mdotout', C.c_double), ('Thot', C.c_double), ('Qirr2Qvishot', C.c_double),
        ('F0', C.c_double), ('Mdisk0', C.c_double), ('Mdot0', C.c_double), ('powerorder', C.c_double),
        ('gaussmu', C.c_double), ('gausssigma', C.c_double),
        ('windparams', C.c_double * 4),
        ('Cirr', C.c_double), ('irrindex', C.c_double), ('Cirrcold', C.c_double), ('irrindexcold', C.c_double),
        ('h2rcold', C.c_double),
        ('angulardistdisk', C.c_int32), ('angulardistns', C.c_int32),
        ('colourfactor', C.c_double), ('emin', C.c_double), ('emax', C.c_double), ('staralbedo', C.c_double),
        ('inclination', C.c_double), ('ephemerist0', C.c_double), ('distance', C.c_double),
        ('inittime', C.c_double), ('time', C.c_double), ('tau', C.c_double), ('eps', C.c_double),
        ('Nx', C.c_uint32), ('gridscale', C.c_int32), ('starlod', C.c_int32), ('nsprop', C.c_int32),
        ('freqx', C.c_double), ('Rx', C.c_double), ('Bx', C.c_double), ('hotspotarea', C.c_double),
        ('epsilonAlfven', C.c_double), ('inversebeta', C.c_double), ('Rdead', C.c_double),
        ('fptype', C.c_int32), ('kappattype', C.c_int32), ('nsgravredshift', C.c_int32), ('_pad1', C.c_int32),
        ('fpparams', C.c_double * 2),
        ('kappatparams', C.c_double * 4),
    ]


FB200_MAX_LAMBDAS = 16
FB200_MAX_PASSBANDS = 4


class FB200Config(C.Structure):
    """ctypes image of ``fb200_config_t``."""
    _fields_ = [
        ('device', C.c_int32), ('n_lambdas', C.c_int32),
        ('lambdas', C.c_double * FB200_MAX_LAMBDAS),
        ('cold_disk', C.c_int32), ('record_F', C.c_int32), ('compute_lx', C.c_int32), ('host_threads', C.c_int32),
        ('threads_per_model', C.c_int32), ('exact_lx_walk', C.c_int32),
        ('n_passbands', C.c_int32), ('passband_len', C.c_int32 * FB200_MAX_PASSBANDS),
        ('passband_lambdas', C.POINTER(C.c_double) * FB200_MAX_PASSBANDS),
        ('passband_transmissions', C.POINTER(C.c_double) * FB200_MAX_PASSBANDS),
        ('star_flux', C.c_int32), ('wait_timeout_ms', C.c_int32),
    ]


class FB200ModelInfo(C.Structure):
    """ctypes image of ``fb200_model_info_t``."""
    _fields_ = [(n, C.c_double) for n in (
        'GM', 'eta', 'R_g', 'cosi', 'distance', 'cosiOverD2', 'tau', 'h_in', 'h_out', 'rin', 'rout', 'risco',
        'F0', 'Mdisk0', 'Mdot0', 'mu_magn', 'R_cor', 'R_dead', 'inverse_beta', 'R_x', 'redshift')] + [
        ('Nx', C.c_int64), ('Nt', C.c_int64), ('first0', C.c_int64)]


# ---- string option -> enum ----
_BAD = 1000  # out-of-range enum: the resolver raises the reference's std::invalid_argument


def _enum(table):
    return lambda s: table.get(s.decode() if isinstance(s, bytes) else s, _BAD)


OPACITY = {'Kramers': 0, 'OPAL': 1}
BOUNDCOND = {'Teff': 0, 'Tirr': 1}
GRIDSCALE = {'log': 0, 'linear': 1}
ANGDIST = {'plane': 0, 'isotropic': 1}
# aliases accepted by cpp/src/arguments.cpp:70,100,165-166 and cpp/src/ns/ns_arguments.cpp:148
INITIALCOND = {'powerF': 0, 'power': 0, 'linearF': 1, 'linear': 1, 'powerSigma': 2,
               'sineF': 3, 'sine': 3, 'sinusF': 3, 'sinus': 3, 'gaussF': 4, 'quasistat': 5,
               'quasistat-ns': 6, 'quasistat_ns': 6}
WINDTYPE = {'no': 0, 'SS73C': 1, 'Cambier2013': 2, '__testA__': 3, '__testB__': 4, '__testC__': 5,
            'ShieldsOscil1986': 6, 'Janiuk2015': 7, 'Shields1986': 8, 'Woods1996AGN': 9, 'Woods1996': 10, 'toy': 11}
# order of windparams[] per wind type (cpp/src/freddi_state.cpp:403-406,422-424,431-433,440-442,452-455,469-472,498-502,539-542,590-594,639-641)
WINDPARAM_NAMES = {0: (), 1: (), 2: ('kC', 'RIC'), 3: ('kA',), 4: ('kB',), 5: ('kC',), 6: ('C_w', 'R_w'),
                   7: ('A_0', 'B_1'), 8: ('Xi_max', 'T_ic', 'Pow'), 9: ('C_0', 'T_ic'),
                   10: ('Xi_max', 'T_ic', 'Pow'), 11: ('C_w',)}
NSPROP = {'dummy': 0, 'sibgatullinsunyaev2000': 1, 'sibsun2000': 1, 'newt': 2}
FPTYPE = {'no-outflow': 0, 'propeller': 1, 'corotation-block': 2, 'eksi-kutlu2010': 3, 'romanova2018': 4,
          'geometrical': 5}
FPPARAM_NAMES = {4: ('par1', 'par2'), 5: ('chi',)}
KAPPATTYPE = {'const': 0, 'corstep': 1, 'romanova2018': 2}
KAPPATPARAM_NAMES = {0: ('value',), 1: ('in', 'out'), 2: ('in', 'out', 'par1', 'par2')}
NSGRAVREDSHIFT = {'off': 0, 'on': 1}

_NONE = object()

# cpp/pywrap/pywrap_freddi_evolution.cpp:15-31
REQUIRED_ARGS = dict(alpha=0.0, Mx=0.0, Mopt=0.0, period=0.0, initialcond='quasistat', F0=0.0, distance=0.0, time=0.0)
# cpp/pywrap/pywrap_freddi_evolution.cpp:33-83
KWDEFAULTS = dict(
    __cgs=True,
    alphacold=None, kerr=0.0, rochelobefill=1.0, Topt=0.0, rin=None, rout=None, risco=None,
    opacity='Kramers', Mdotout=0.0, boundcond='Teff', initialcond='powerF', Thot=0.0, Qirr2Qvishot=0.0,
    F0=None, Mdisk0=None, Mdot0=None, powerorder=None, gaussmu=None, gausssigma=None,
    windtype='no', windparams=(),
    Cirr=0.0, irrindex=0.0, Cirrcold=0.0, irrindexcold=0.0, h2rcold=0.0, angulardistdisk='plane',
    colourfactor=1.7, emin=kev_to_hertz(1.0), emax=kev_to_hertz(12.0), staralbedo=0.0, inclination=0.0,
    ephemerist0=0.0,
    inittime=0.0, tau=None, Nx=1000, gridscale='log', starlod=3, eps=None,
)
# cpp/pywrap/pywrap_freddi_evolution.cpp:194-224
NS_REQUIRED_ARGS = dict(REQUIRED_ARGS, Bx=0.0)
NS_KWDEFAULTS = dict(
    KWDEFAULTS,
    angulardistns='isotropic', nsprop='dummy', freqx=None, Rx=None, hotspotarea=1.0, epsilonAlfven=1.0,
    inversebeta=0.0, Rdead=0.0, fptype='no-outflow', fpparams=(), kappattype='const',
    kappatparams={'value': 1.0 / 3.0, 'in': 1.0 / 3.0, 'out': 1.0 / 3.0}, nsgravredshift='off',
)


def merge_kwargs(kwargs, is_ns, cls_name='Freddi'):
    """Defaults + user kwargs, rejecting unknown names (pywrap_freddi_evolution.cpp:94-104,183-185)."""
    required = NS_REQUIRED_ARGS if is_ns else REQUIRED_ARGS
    defaults = NS_KWDEFAULTS if is_ns else KWDEFAULTS
    for key in kwargs:
        if key not in defaults and key not in required:
            raise TypeError('{} got an unexpected keyword argument {}'.format(cls_name, key))
    kw = dict(defaults)
    kw.update(kwargs)
    for key in required:
        if key not in kw:
            raise KeyError(key)  # boost::python raises KeyError on kw["alpha"] for a missing required key
    return kw


def _to_cgs(kw):
    """cpp/pywrap/pywrap_freddi_evolution.cpp:107-130 (applied only when __cgs is False)."""
    if kw['__cgs']:
        return kw
    kw = dict(kw)
    kw['Mx'] = float(kw['Mx']) * SOLAR_MASS
    kw['Mopt'] = float(kw['Mopt']) * SOLAR_MASS
    kw['period'] = float(kw['period']) * DAY
    if kw['rin'] is not None:
        kw['rin'] = rg_to_cm(float(kw['rin']), kw['Mx'])
    if kw['rout'] is not None:
        kw['rout'] = float(kw['rout']) * SOLAR_RADIUS
    kw['emin'] = kev_to_hertz(float(kw['emin']))
    kw['emax'] = kev_to_hertz(float(kw['emax']))
    kw['distance'] = float(kw['distance']) * 1000.0 * PARSEC
    kw['inittime'] = float(kw['inittime']) * DAY
    kw['time'] = float(kw['time']) * DAY
    if kw['tau'] is not None:
        kw['tau'] = float(kw['tau']) * DAY
    return kw


def _optd(v):
    return math.nan if v is None else float(v)


def _param_array(values, names, n, what):
    out = [0.0] * n
    values = dict(values) if not isinstance(values, dict) else values
    for i, name in enumerate(names):
        if name not in values:
            # std::map::at -> std::out_of_range in the reference (e.g. cpp/src/freddi_state.cpp:405)
            raise IndexError('{} requires parameter {!r}'.format(what, name))
        out[i] = float(values[name])
    return out


def kwargs_to_args(kw, is_ns):
    """Merged kwargs (see merge_kwargs) -> FB200Args, all CGS."""
    kw = _to_cgs(kw)
    a = FB200Args()
    a.is_ns = 1 if is_ns else 0
    a.alpha = float(kw['alpha']); a.alphacold = _optd(kw['alphacold'])
    a.Mx = float(kw['Mx']); a.kerr = float(kw['kerr']); a.period = float(kw['period']); a.Mopt = float(kw['Mopt'])
    a.rochelobefill = float(kw['rochelobefill']); a.Topt = float(kw['Topt'])
    a.rin = _optd(kw['rin']); a.rout = _optd(kw['rout']); a.risco = _optd(kw['risco'])
    a.opacity = _enum(OPACITY)(kw['opacity']); a.boundcond = _enum(BOUNDCOND)(kw['boundcond'])
    a.initialcond = _enum(INITIALCOND)(kw['initialcond']); a.windtype = _enum(WINDTYPE)(kw['windtype'])
    a.Mdotout = float(kw['Mdotout']); a.Thot = float(kw['Thot']); a.Qirr2Qvishot = float(kw['Qirr2Qvishot'])
    a.F0 = _optd(kw['F0']); a.Mdisk0 = _optd(kw['Mdisk0']); a.Mdot0 = _optd(kw['Mdot0'])
    a.powerorder = _optd(kw['powerorder']); a.gaussmu = _optd(kw['gaussmu']); a.gausssigma = _optd(kw['gausssigma'])
    wp = _param_array(kw['windparams'], WINDPARAM_NAMES.get(a.windtype, ()), 4, 'windtype')
    for i in range(4):
        a.windparams[i] = wp[i]
    a.Cirr = float(kw['Cirr']); a.irrindex = float(kw['irrindex']); a.Cirrcold = float(kw['Cirrcold'])
    a.irrindexcold = float(kw['irrindexcold']); a.h2rcold = float(kw['h2rcold'])
    a.angulardistdisk = _enum(ANGDIST)(kw['angulardistdisk'])
    a.colourfactor = float(kw['colourfactor']); a.emin = float(kw['emin']); a.emax = float(kw['emax'])
    a.staralbedo = float(kw['staralbedo']); a.inclination = float(kw['inclination'])
    a.ephemerist0 = float(kw['ephemerist0']); a.distance = float(kw['distance'])
    a.inittime = float(kw['inittime']); a.time = float(kw['time']); a.tau = _optd(kw['tau'])
    a.eps = 1e-6 if kw['eps'] is None else float(kw['eps'])  # cpp/include/arguments.hpp:327
    a.Nx = int(kw['Nx']); a.gridscale = _enum(GRIDSCALE)(kw['gridscale']); a.starlod = int(kw['starlod'])
    a.angulardistns = ANGDIST['isotropic']
    a.freqx = math.nan; a.Rx = math.nan
    a.hotspotarea = 1.0; a.epsilonAlfven = 1.0
    if is_ns:
        a.angulardistns = _enum(ANGDIST)(kw['angulardistns'])
        a.nsprop = _enum(NSPROP)(kw['nsprop'])
        a.freqx = _optd(kw['freqx']); a.Rx = _optd(kw['Rx']); a.Bx = float(kw['Bx'])
        a.hotspotarea = float(kw['hotspotarea']); a.epsilonAlfven = float(kw['epsilonAlfven'])
        a.inversebeta = float(kw['inversebeta']); a.Rdead = float(kw['Rdead'])
        a.fptype = _enum(FPTYPE)(kw['fptype']); a.kappattype = _enum(KAPPATTYPE)(kw['kappattype'])
        a.nsgravredshift = _enum(NSGRAVREDSHIFT)(kw['nsgravredshift'])
        fp = _param_array(kw['fpparams'], FPPARAM_NAMES.get(a.fptype, ()), 2, 'fptype')
        for i in range(2):
            a.fpparams[i] = fp[i]
        kp = _param_array(kw['kappatparams'], KAPPATPARAM_NAMES.get(a.kappattype, ()), 4, 'kappattype')
        for i in range(4):
            a.kappatparams[i] = kp[i]
    return a


def make_args(is_ns=False, **kwargs):
    """Convenience: user kwargs -> FB200Args."""
    return kwargs_to_args(merge_kwargs(kwargs, is_ns, 'FreddiNeutronStar' if is_ns else 'Freddi'), is_ns)


def args_array(args_list):
    arr = (FB200Args * len(args_list))()
    for i, a in enumerate(args_list):
        C.memmove(C.byref(arr, i * C.sizeof(FB200Args)), C.byref(a), C.sizeof(FB200Args))
    return arr
