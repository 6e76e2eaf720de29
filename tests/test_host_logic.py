"""CPU tests of everything around the kernels: the C ABI loads and exports what include/freddi_b200.h declares,
the host-side argument resolver reproduces the reference's constructors, and DiscEngine — the single source the
CUDA kernel is compiled from — run under the sequential host policy (tests/host_emu, test infrastructure)
matches the oracle to 1e-10 with identical Picard counts, row counts and first/last."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import pyoracle
from tests import common as cm
from freddi_b200.args import make_args, FB200Args, merge_kwargs


def test_c_abi_exports_every_declared_symbol(built):
    from freddi_b200 import _lib
    header = open(os.path.join(cm.ROOT, 'include', 'freddi_b200.h')).read()
    declared = set(re.findall(r'\b(fb200_[a-z0-9_]+)\s*\(', header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(_lib.lib, name), name
    assert _lib.lib.fb200_abi_version() == 4


def test_args_struct_layout_matches_the_header(built):
    """sizeof(fb200_args_t) seen by C equals the ctypes image: defaults written by C read back correctly."""
    from freddi_b200 import _lib
    a = FB200Args()
    _lib.lib.fb200_args_default(C.byref(a), 1)
    assert a.is_ns == 1 and a.Nx == 1000 and a.eps == 1e-6 and a.colourfactor == 1.7 and a.hotspotarea == 1.0
    assert abs(a.kappatparams[0] - 1 / 3) < 1e-16 and a.starlod == 3 and np.isnan(a.Bx)
    assert abs(a.emax / a.emin - 12) < 1e-12


def test_no_device_fails_loudly(built):
    """There is no CPU fallback: without a CUDA device creation raises."""
    from freddi_b200 import _lib, Ensemble
    if _lib.lib.fb200_device_count() > 0:
        pytest.skip('a CUDA device is present')
    with pytest.raises(_lib.NoDeviceError):
        Ensemble([cm.ini_args()])


def test_kwargs_surface_matches_reference_constructor():
    """cpp/pywrap/pywrap_freddi_evolution.cpp:86-104: unknown keyword -> TypeError, missing required -> KeyError."""
    with pytest.raises(TypeError):
        merge_kwargs(dict(cm.INI, nosuch=1), False)
    bad = dict(cm.INI)
    del bad['alpha']
    with pytest.raises(KeyError):
        merge_kwargs(bad, False)
    with pytest.raises(TypeError):
        merge_kwargs(dict(cm.INI, Bx=1e8), False)  # NS-only keyword on a BH model
    a = make_args(True, **dict(cm.INI, Bx=1e8))
    assert a.is_ns == 1 and a.Bx == 1e8


RESOLVE_CASES = {
    'default': (False, {}),
    'quasistat_Mdisk0': (False, dict(initialcond='quasistat', Mdisk0=1e26, F0=None)),
    'quasistat_Mdot0_opal': (False, dict(initialcond='quasistat', Mdot0=1e18, F0=None, opacity='OPAL')),
    'sine_Mdot0': (False, dict(initialcond='sineF', Mdot0=1e18, F0=None)),
    'gauss': (False, dict(initialcond='gaussF', gaussmu=0.8, gausssigma=0.1)),
    'gauss_underflow_ramp': (False, dict(initialcond='gaussF', gaussmu=1.0, gausssigma=0.01)),
    'powerSigma': (False, dict(initialcond='powerSigma', powerorder=1, Mdisk0=1e25, F0=None)),
    'linear_grid_kerr': (False, dict(gridscale='linear', kerr=0.7, Nx=257, initialcond='linearF')),
    'rin_rout_given': (False, dict(rin=10, rout=2.0, alphacold=0.02)),
    'ns_dummy': (True, dict(cm.NS_CI)),
    'ns_sibsun_invbeta': (True, dict(cm.NS_CI, inversebeta=0.5, freqx=300, nsprop='sibgatullinsunyaev2000', nsgravredshift='on')),
    'ns_quasistat_ns': (True, dict(cm.NS_CI, initialcond='quasistat-ns', freqx=400, Rx=1e6, Bx=3e8, nsprop='newt')),
}


@pytest.mark.parametrize('name', sorted(RESOLVE_CASES))
def test_resolver_reproduces_reference_constructors(ref, built, name):
    """h, F(t=0), first and the derived constants are bit-identical to what the reference's constructors produce
    (same libm, same operation order): cpp/src/arguments.cpp:57-351, cpp/src/freddi_state.cpp:13-71."""
    from freddi_b200.ensemble import resolve
    is_ns, kw = RESOLVE_CASES[name]
    a = cm.ini_args(is_ns, **kw)
    info, h, F = resolve(a)
    m = ref.model(a)
    ri = m.info()
    np.testing.assert_array_equal(h, m.field('h'))
    np.testing.assert_array_equal(F, m.field('F'))
    for f in ('GM', 'eta', 'R_g', 'cosi', 'distance', 'cosiOverD2', 'tau', 'h_in', 'h_out', 'rin', 'rout', 'risco'):
        assert getattr(info, f) == getattr(ri, f), f
    assert (info.Nx, info.Nt, info.first0) == (ri.Nx, ri.Nt, ri.first0)
    if is_ns:
        for f in ('mu_magn', 'R_cor', 'R_dead', 'inverse_beta', 'R_x', 'redshift'):
            assert getattr(info, f) == getattr(ri, f), f


def test_resolver_memos_are_transparent(ref, built):
    """The per-thread memos of the resolver (grid, quasistat normalisation integral, f_F on the grid: resolve.cpp) return the
    same bits whether they hit or miss, in any order of the models of a sweep — and those bits are the reference's."""
    from freddi_b200.ensemble import resolve
    mk = lambda **kw: cm.ini_args(False, initialcond='quasistat', F0=None, Thot=1e4, boundcond='Tirr', Cirr=1e-4, **kw)  # noqa: E731
    sweep = [mk(Mx=5, alpha=0.3, Mdisk0=1e26), mk(Mx=5, alpha=0.7, Mdisk0=3e25),       # same grid: hits
             mk(Mx=9, alpha=0.3, Mdisk0=1e26), mk(Mx=9, alpha=0.3, Mdisk0=1e26, opacity='OPAL'),
             mk(Mx=9, alpha=0.3, Mdisk0=1e26, gridscale='linear'), mk(Mx=9, alpha=0.3, Mdisk0=1e26, Nx=300)]
    first = [resolve(a) for a in sweep]
    for order in (range(len(sweep)), reversed(range(len(sweep))), (0, 0, 2, 1, 3, 3, 5, 4, 0)):
        for k in order:
            info, h, F = resolve(sweep[k])
            np.testing.assert_array_equal(h, first[k][1])
            np.testing.assert_array_equal(F, first[k][2])
            assert info.first0 == first[k][0].first0
    for a, (info, h, F) in zip(sweep, first):
        m = ref.model(a)
        np.testing.assert_array_equal(h, m.field('h'))
        np.testing.assert_array_equal(F, m.field('F'))


@pytest.mark.parametrize('kw,exc', [
    (dict(F0=None), RuntimeError),                       # One of F0, Mdisk0 or Mdot0 must be specified
    (dict(Mdot0=1e18), RuntimeError),                    # Set Mdisk0 or F0 instead of Mdot0 if initialcond=powerF
    (dict(powerorder=None), RuntimeError),               # powerorder must be specified
    (dict(initialcond='gaussF', gaussmu=1.5), RuntimeError),
    (dict(opacity='nope'), ValueError),                  # std::invalid_argument
    (dict(gridscale='nope'), ValueError),
    (dict(windtype='nope'), ValueError),
    (dict(angulardistdisk='nope'), ValueError),
    (dict(initialcond='nope'), RuntimeError),            # Invalid value of initialcond
])
def test_resolver_error_classes(ref, built, kw, exc):
    """Bad arguments come back as the exception class the reference throws; the oracle agrees on the class."""
    from freddi_b200.ensemble import resolve
    a = cm.ini_args(**kw)
    with pytest.raises(exc):
        resolve(a)
    with pytest.raises(pyoracle.CheckerError) as ei:
        ref.model(a)
    assert (ei.value.code == -1) == (exc is ValueError)


def emu_vs_ref(ref, emu, a, lam=(), cold=False, passbands=(), star_flux=False):
    r = ref.run(a, lam, cold_disk=cold, passbands=passbands, star_flux=star_flux)
    e = emu.run(a, r['Nt'], r['Nx'], lam, cold, passbands=passbands, star_flux=star_flux)
    names = pyoracle.column_names(len(lam), cold, bool(a.is_ns), [p.name for p in passbands], star_flux)
    n = r['rows_valid']
    assert e['rows_valid'] == n
    np.testing.assert_array_equal(e['picard'][:n], r['picard'][:n])
    np.testing.assert_array_equal(e['first'][:n], r['first'][:n])
    np.testing.assert_array_equal(e['last'][:n], r['last'][:n])
    cm.assert_rows_close(e['rows'][:n], r['rows'][:n], names)
    assert cm.rel_err(e['F'][:n], r['F'][:n]).max() <= cm.RTOL
    assert (e['rows'][:n, 0] == r['rows'][:n, 0]).all()  # t is accumulated, bit-exact
    return r, e


@pytest.mark.parametrize('path', cm.golden_files(), ids=os.path.basename)
def test_engine_host_policy_matches_oracle_on_golden_configurations(ref, emu, path):
    kw, lam, pb, _data = pyoracle.load_golden(path)
    emu_vs_ref(ref, emu, make_args(False, **kw), lam, passbands=cm.golden_passbands(pb))


def test_passbands_hot_and_cold_regions(ref, emu):
    """Fnu<passband> and Fnu<passband>_cold (cpp/src/output.cpp:258-271) next to the wavelength columns, NS rows after
    them; t_dnu of the prepared table equals EnergyPassband::t_dnu bit for bit."""
    pbs = cm.golden_passbands(['passbands/Swift_V.dat', 'passbands/Swift_B.dat'])
    assert [p.name for p in pbs] == ['Swift_V', 'Swift_B']
    np.testing.assert_array_equal(emu.set_passbands(pbs), ref.set_passbands(pbs))
    emu.set_passbands(())
    ref.set_passbands(())
    kw = dict(initialcond='quasistat', Thot=1e4, Cirr=3e-4, Cirrcold=1e-4, h2rcold=0.03, irrindex=0.5, irrindexcold=0.2,
              boundcond='Tirr', angulardistdisk='plane', inclination=30, time=25)
    r, e = emu_vs_ref(ref, emu, cm.ini_args(False, **kw), cm.LAM3[:1], True, pbs)
    assert r['rows'].shape[1] == 15 + 2 + 4
    assert r['last'][r['rows_valid'] - 1] < r['Nx'] - 2 and r['rows'][r['rows_valid'] - 1, -1] > 0  # a cold zone exists and shines
    emu_vs_ref(ref, emu, cm.ini_args(True, **dict(cm.NS_CI, time=5)), (), False, pbs[:1])


# python/test/test_ns.py:37-55 (moving inner radius), shortened to 300 steps for the CPU suite (the GPU suite runs the 2000)
NS_MOVING = dict(Mx=1.4, alpha=0.25, Mopt=0.5, period=0.25, F0=None, Mdot0=1e16, initialcond='quasistat', Bx=1e8, freqx=500, Rx=1e6,
                 rin=1.3e6 * 2.99792458e10 ** 2 / (6.673e-8 * 1.4 * 1.98892e33), rout=1e11 / 6.955e10, Rdead=5e6, time=30, tau=0.1,
                 powerorder=None)

ENGINE_CASES = {
    'woods1996_irr_opal': (False, dict(cm.WOODS), cm.LAM3, False),
    'shields1986': (False, dict(cm.WOODS, windtype='Shields1986', time=10), cm.LAM3, False),
    'janiuk2015': (False, dict(windtype='Janiuk2015', windparams={'A_0': 10, 'B_1': 0.5}, initialcond='quasistat', time=10), (), False),
    'testA': (False, dict(windtype='__testA__', windparams={'kA': 2.0}, time=10), (), False),
    'testC_mdotout': (False, dict(windtype='__testC__', windparams={'kC': 3.0}, Mdotout=-1e18, time=10), (), False),
    'cold_disk_plane_teff': (False, dict(initialcond='quasistat', Thot=1e4, Cirr=3e-4, Cirrcold=1e-4, h2rcold=0.03, irrindex=0.5,
                                         irrindexcold=0.2, boundcond='Teff', angulardistdisk='plane', inclination=30, time=25), cm.LAM3, True),
    'ns_ci': (True, dict(cm.NS_CI, time=25), cm.LAM3, False),
    'ns_collapse_invbeta': (True, dict(cm.NS_CI, inversebeta=0.5, freqx=300, Rx=1.2e6, fptype='romanova2018',
                                       fpparams={'par1': 0.15, 'par2': 0.92}, kappattype='romanova2018',
                                       kappatparams={'in': 1 / 3, 'out': 1 / 3, 'par1': 0.15, 'par2': 0.92},
                                       nsprop='sibgatullinsunyaev2000', nsgravredshift='on', Cirr=1e-4, Thot=1e4,
                                       boundcond='Tirr', angulardistns='plane', time=30), cm.LAM3, False),
    'ns_quasistat_ns_geometrical': (True, dict(cm.NS_CI, initialcond='quasistat-ns', freqx=400, Rx=1e6, Bx=3e8, nsprop='newt',
                                               fptype='geometrical', fpparams={'chi': 30.0}, kappattype='corstep',
                                               kappatparams={'in': 0.3, 'out': 0.4}, time=20), cm.LAM3, False),
    'small_grid_nx64': (False, dict(Nx=64, initialcond='quasistat', time=10), cm.LAM3, False),
    # branches without a test in round 1 (cpp/src/freddi_state.cpp:546-586, cpp/src/ns/ns_evolution.cpp:144-225)
    'woods1996agn_strong': (False, dict(cm.WOODS, windtype='Woods1996AGN', windparams={'C_0': 100., 'T_ic': 1e8}, time=10), cm.LAM3, False),
    'woods1996agn_shrinking': (False, dict(cm.WOODS, windtype='Woods1996AGN', windparams={'C_0': 10., 'T_ic': 1e7}, Thot=1e4,
                                           boundcond='Tirr', time=10), cm.LAM3, False),
    'woods1996agn_super_eddington': (False, dict(cm.WOODS, windtype='Woods1996AGN', windparams={'C_0': 1., 'T_ic': 1e9}, F0=6e38, time=5),
                                     cm.LAM3, False),
    'shields_oscil1986': (False, dict(windtype='ShieldsOscil1986', windparams={'C_w': 1.0, 'R_w': 0.5}, initialcond='quasistat', time=10), (), False),
    'ns_propeller': (True, dict(NS_MOVING, fptype='propeller'), cm.LAM3, False),
    'ns_corotation_block': (True, dict(NS_MOVING, fptype='corotation-block'), cm.LAM3, False),
    'ns_no_outflow': (True, dict(NS_MOVING, fptype='no-outflow'), cm.LAM3, False),
    'ns_geometrical_chi0': (True, dict(NS_MOVING, fptype='geometrical', fpparams={'chi': 0.}), cm.LAM3, False),
}


@pytest.mark.parametrize('name', sorted(ENGINE_CASES))
def test_engine_host_policy_matches_oracle(ref, emu, name):
    is_ns, kw, lam, cold = ENGINE_CASES[name]
    emu_vs_ref(ref, emu, cm.ini_args(is_ns, **kw), lam, cold)


def test_ns_moving_inner_radius(ref, emu):
    """python/test/test_ns.py:37-55 set-up: the magnetosphere pushes `first` outwards during the decay."""
    kw = dict(Mx=1.4, alpha=0.25, Mopt=0.5, period=0.25, F0=None, Mdot0=1e16, initialcond='quasistat', Bx=1e8, freqx=500, Rx=1e6,
              rin=1.3e6 * 2.99792458e10 ** 2 / (6.673e-8 * 1.4 * 1.98892e33), rout=1e11 / 6.955e10, Rdead=5e6, time=40, tau=0.1,
              powerorder=None, fptype='eksi-kutlu2010')
    r, e = emu_vs_ref(ref, emu, cm.ini_args(True, **kw), cm.LAM3)
    assert r['first'][-1] > r['first'][0]


def test_corotation_block_switches_off_beyond_the_corotation_radius(ref, emu):
    """python/test/test_ns.py:74-80 on the engine (host policy): 2000 steps of 0.1 d, the inner radius crosses R_cor and
    fp drops from 1 to 0 exactly there; rows against the oracle."""
    from freddi_b200.ensemble import resolve
    a = cm.ini_args(True, **dict(NS_MOVING, fptype='corotation-block', time=200))
    r, e = emu_vs_ref(ref, emu, a, ())
    cols = pyoracle.column_names(0, False, True)
    info, _h, _F = resolve(a)
    ratio = e['rows'][:, cols.index('Rin')] / info.R_cor
    assert (ratio < 1).any() and (ratio > 1).any()
    np.testing.assert_array_equal(ratio < 1, e['rows'][:, cols.index('fpin')] == 1)


def test_cambier2013_with_the_uninitialised_member_defined(ref, emu):
    """See tests/test_gpu_branches.py: the reference's Cambier2013 reads a member it never initialises; with the member
    defined to the normalised Mdot0 (what this library uses) its code and the engine agree."""
    from freddi_b200.ensemble import resolve
    a = cm.ini_args(False, windtype='Cambier2013', windparams={'kC': 1.0, 'RIC': 1.0}, initialcond='quasistat', time=10)
    info, _h, _F = resolve(a)
    ref.define_mdot0(info.Mdot0)
    try:
        r, e = emu_vs_ref(ref, emu, a, cm.LAM3)
    finally:
        ref.define_mdot0()
    assert r['picard'][1:].min() >= 2  # a static C term: T1 of SURVEY.md 9


def test_planck_band_integral_controller(emu):
    """The adaptive Cash-Karp quadrature: exact zeros for cold rings, agreement with a fine trapezoid within the
    controller's tolerance, monotone in T; trial-step counts are those of odeint's controller (>= 4: the step can
    grow by at most x4.5 from 1% of the interval)."""
    nu1, nu2 = 2.417989e17, 2.9015868e18
    assert emu.planck(1e4, nu1, nu2) == (0.0, 0)     # exp overflows at nu1: identically zero integrand
    assert emu.planck(0.0, nu1, nu2) == (0.0, 0)
    assert emu.planck(float('nan'), nu1, nu2) == (0.0, 0)
    prev = 0.0
    for T in (3e5, 1e6, 3e6, 1e7, 3e7, 1e8):
        v, n = emu.planck(T, nu1, nu2)
        nu = np.linspace(nu1, nu2, 400001)
        with np.errstate(over='ignore'):
            B = 2 * 6.62606896e-27 / 2.99792458e10 ** 2 * nu ** 3 / np.expm1(6.62606896e-27 / 1.3806504e-16 * nu / T)
        exact = np.trapezoid(B, nu)
        assert abs(v / exact - 1) < 2e-3, (T, v, exact)
        assert v > prev and n >= 4
        prev = v


def test_large_grid_engine_build_matches_oracle(ref, emu_big):
    """DiscEngine compiled for the large-grid build (FB200_NX_MAX = 16384: more blocks than threads in the solver's block
    elimination, run-time slot counts) against the oracle at Nx = 2500: irradiated shrinking disc and a static wind on a linear grid."""
    emu_vs_ref(ref, emu_big, cm.ini_args(False, initialcond='quasistat', Thot=1e4, boundcond='Tirr', angulardistdisk='isotropic',
                                         Cirr=3e-4, Nx=2500, time=6), cm.LAM3)
    emu_vs_ref(ref, emu_big, cm.ini_args(False, windtype='__testA__', windparams={'kA': 1.0}, Mdotout=1e18, initialcond='sineF',
                                         Mdot0=1e18, F0=None, Nx=2500, gridscale='linear', time=20, tau=1), ())


STAR = dict(initialcond='quasistat', Thot=1e4, boundcond='Tirr', Cirr=3e-4, Topt=4000, inclination=40, staralbedo=0.3, h2rcold=0.03,
            ephemerist0=0.1, time=4)


@pytest.mark.parametrize('kw', [dict(), dict(Mx=9, Mopt=12, period=1.5, rochelobefill=0.8, starlod=2), dict(starlod=0, Mopt=5)],
                         ids=['default', 'heavy_donor_underfilled_lod2', 'equal_mass_lod0'])
def test_star_geometry_matches_reference(ref, emu, kw):
    """The triangulated Roche-lobe surface (star.cpp) against the reference's Star::triangles(): centres, normals, areas."""
    a = cm.ini_args(False, **kw)
    want = ref.model(a).star_triangles()
    got = emu.star_triangles(a)
    assert got.shape == want.shape == (20 * 4 ** int(a.starlod), 7)
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
    np.testing.assert_allclose(got[:, 6], want[:, 6], rtol=1e-11)


@pytest.mark.parametrize('name', ['bh_isotropic', 'bh_plane_cold_passband', 'ns_two_sources'])
def test_star_flux_columns(ref, emu, name):
    """--starflux: Fnu*_star, _star_min, _star_max (cpp/src/output.cpp:234-255,272-291).  The state at t0 has no irradiation
    sources yet (they are set at the end of step(), cpp/src/freddi_evolution.cpp:28), later rows do."""
    if name == 'bh_isotropic':
        r, e = emu_vs_ref(ref, emu, cm.ini_args(False, angulardistdisk='isotropic', **STAR), cm.LAM3, star_flux=True)
        k = 15 + 1
        np.testing.assert_allclose(r['rows'][0, k + 1:k + 3], r['rows'][0, k], rtol=1e-6)  # unirradiated: Topt only, (almost) the same from every side
        assert r['rows'][3, k + 2] > 3 * r['rows'][3, k + 1]                            # superior conjunction shows the heated side
    elif name == 'bh_plane_cold_passband':
        pbs = cm.golden_passbands(['passbands/Swift_V.dat'])
        emu_vs_ref(ref, emu, cm.ini_args(False, angulardistdisk='plane', Cirrcold=1e-4, irrindex=0.5, **dict(STAR, time=12)), cm.LAM3[:1],
                   True, pbs, star_flux=True)
    else:
        emu_vs_ref(ref, emu, cm.ini_args(True, **dict(cm.NS_CI, Topt=3500, inclination=60, staralbedo=0.5, angulardistns='plane',
                                                      Cirr=1e-3, time=3)), cm.LAM3[:2], star_flux=True)
