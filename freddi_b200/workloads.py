"""The BASELINE.json workloads as concrete argument lists, shared by bench.py and the tests."""
import numpy as np

from .args import make_args

LAM3 = [3e-5, 5e-5, 8e-5]  # 3000, 5000, 8000 Angstrom [cm]

# freddi.ini (config #1 of BASELINE.json), CLI units.  emin/emax are spelled out [keV] as the golden files' headers do:
# with __cgs=False the reference converts them keV -> Hz even when they come from its (already Hz) defaults
# (cpp/pywrap/pywrap_freddi_evolution.cpp:117-118), which would put the X-ray band at ~1e35 Hz and make Lx == 0.
INI = dict(alpha=0.25, Mx=5, Mopt=0.5, period=0.25, F0=2e38, initialcond='powerF', powerorder=6, gaussmu=1, gausssigma=0.25,
           distance=10, time=50, tau=0.25, emin=1, emax=12, __cgs=False)


def ini_args(is_ns=False, **kw):
    d = dict(INI)
    d.update(kw)
    return make_args(is_ns, **d)


def config2_args(alphas, cirrs, **kw):
    """BASELINE config #2: quasistat, Thot=1e4, boundcond=Tirr, isotropic; sweep alpha x Cirr."""
    out = []
    for a in alphas:
        for c in cirrs:
            out.append(ini_args(initialcond='quasistat', Thot=1e4, boundcond='Tirr', angulardistdisk='isotropic',
                                alpha=float(a), Cirr=float(c), **kw))
    return out


NS_CI = dict(Mx=1.4, alpha=0.25, Mopt=0.5, Mdot0=1e17, F0=None, initialcond='quasistat', Bx=1e8, period=0.25, distance=10,
             time=50, tau=0.25, powerorder=None)


def config3_args(bxs, alphas, **kw):
    """BASELINE config #3: freddi-ns CI command, sweep Bx x alpha."""
    out = []
    for b in bxs:
        for a in alphas:
            d = dict(NS_CI)
            d.update(Bx=float(b), alpha=float(a))
            d.update(kw)
            out.append(ini_args(True, **d))
    return out


WOODS = dict(alpha=0.5, Mx=9, rout=1, period=0.5, Mopt=0.5, F0=2e37, kerr=0.4, Cirr=1e-3, opacity='OPAL',
             initialcond='quasistat', windtype='Woods1996', windparams={'Xi_max': 10, 'T_ic': 1e8, 'Pow': 1}, distance=5)


def config4_args(alphas, cirrs, **kw):
    """BASELINE config #4: README wind example (Readme.md:997-1003), sweep alpha x Cirr, 3 lambdas every row."""
    out = []
    for a in alphas:
        for c in cirrs:
            d = dict(WOODS)
            d.update(alpha=float(a), Cirr=float(c))
            d.update(kw)
            out.append(ini_args(False, **d))
    return out


def config5_args(mxs, alphas, mdisks, cirrs, **kw):
    """BASELINE config #5: Mx x alpha x Mdisk0 x Cirr, quasistat, Thot=1e4, boundcond=Tirr, isotropic."""
    out = []
    for m in mxs:
        for a in alphas:
            for md in mdisks:
                for c in cirrs:
                    out.append(ini_args(initialcond='quasistat', Thot=1e4, boundcond='Tirr', angulardistdisk='isotropic',
                                        Mx=float(m), alpha=float(a), Mdisk0=float(md), F0=None, Cirr=float(c), **kw))
    return out


# ---- the same sweeps as packed argument arrays, built without a Python call per model (10^5 models in milliseconds) ----
def sweep_array(base, n, **columns):
    """ctypes array of `n` copies of `base` (an FB200Args) with the given double fields replaced column-wise (CGS values)."""
    import ctypes as C
    from .args import FB200Args
    size = C.sizeof(FB200Args)
    arr = (FB200Args * n)()
    raw = np.frombuffer(arr, dtype=np.uint8).reshape(n, size)
    raw[:] = np.frombuffer(bytes(base), dtype=np.uint8)
    dbl = np.frombuffer(arr, dtype=np.float64).reshape(n, size // 8)
    for name, values in columns.items():
        off = getattr(FB200Args, name).offset
        assert off % 8 == 0 and getattr(FB200Args, name).size == 8, name
        dbl[:, off // 8] = np.asarray(values, dtype=np.float64)
    return arr


SOLAR_MASS_G = 1.98892e33  # cpp/include/constants.hpp (solar_mass), as freddi_b200.args applies it


def config2_sweep(alphas, cirrs, indices=None):
    """config2_args(alphas, cirrs) as a packed array; `indices` selects models of the flattened (alpha-major) sweep."""
    alphas, cirrs = np.asarray(alphas, dtype=float), np.asarray(cirrs, dtype=float)
    k = np.arange(alphas.size * cirrs.size) if indices is None else np.asarray(indices)
    base = config2_args(alphas[:1], cirrs[:1])[0]
    return sweep_array(base, k.size, alpha=alphas[k // cirrs.size], Cirr=cirrs[k % cirrs.size])


def config3_sweep(bxs, alphas, indices=None):
    """config3_args(bxs, alphas) as a packed array (Bx-major)."""
    bxs, alphas = np.asarray(bxs, dtype=float), np.asarray(alphas, dtype=float)
    k = np.arange(bxs.size * alphas.size) if indices is None else np.asarray(indices)
    base = config3_args(bxs[:1], alphas[:1])[0]
    return sweep_array(base, k.size, Bx=bxs[k // alphas.size], alpha=alphas[k % alphas.size])


def config4_sweep(alphas, cirrs, indices=None):
    alphas, cirrs = np.asarray(alphas, dtype=float), np.asarray(cirrs, dtype=float)
    k = np.arange(alphas.size * cirrs.size) if indices is None else np.asarray(indices)
    base = config4_args(alphas[:1], cirrs[:1])[0]
    return sweep_array(base, k.size, alpha=alphas[k // cirrs.size], Cirr=cirrs[k % cirrs.size])


def config5_sweep(mxs, alphas, mdisks, cirrs, indices=None):
    """config5_args(...) as a packed array: model k = ((i_Mx * n_alpha + i_alpha) * n_Mdisk + i_Mdisk) * n_Cirr + i_Cirr."""
    mxs, alphas, mdisks, cirrs = (np.asarray(x, dtype=float) for x in (mxs, alphas, mdisks, cirrs))
    k = np.arange(mxs.size * alphas.size * mdisks.size * cirrs.size) if indices is None else np.asarray(indices)
    base = config5_args(mxs[:1], alphas[:1], mdisks[:1], cirrs[:1])[0]
    ic = k % cirrs.size
    imd = (k // cirrs.size) % mdisks.size
    ia = (k // (cirrs.size * mdisks.size)) % alphas.size
    im = k // (cirrs.size * mdisks.size * alphas.size)
    return sweep_array(base, k.size, Mx=mxs[im] * SOLAR_MASS_G, alpha=alphas[ia], Mdisk0=mdisks[imd], Cirr=cirrs[ic])
