"""GPU parity: the sm_100a kernels, called through the C ABI, against the oracle (oracle/_ref = the reference's
own sources) on identical inputs.  Tolerance 1e-10 relative (BASELINE.json north_star) on every column of
every row and on F; integers (row count, first/last per row, Picard iterations) exact."""
import numpy as np
import pytest

from oracle import pyoracle
from tests import common as cm

pytestmark = pytest.mark.gpu


def run_gpu(args_list, lambdas=(), cold_disk=False, record_F=True, passbands=(), star_flux=False):
    from freddi_b200 import Ensemble
    with Ensemble(args_list, lambdas=lambdas, cold_disk=cold_disk, record_F=record_F, passbands=passbands, star_flux=star_flux) as ens:
        ens.evolve(-1)
        out = dict(rows=ens.rows(), columns=ens.columns)
        out['status'], out['valid'], out['picard'] = ens.status()
        if record_F:
            out['F'] = [ens.field('F', r) for r in range(ens.n_rows)]
            out['bounds'] = [ens.row_bounds(r) for r in range(ens.n_rows)]
    return out


def check_model(g, k, r, what, snapshot_rows=None):
    n = r['rows_valid']
    assert g['valid'][k] == n, '{}: rows_valid {} != {}'.format(what, g['valid'][k], n)
    assert g['picard'][k] == r['picard'].sum(), '{}: Picard iterations {} != {}'.format(what, g['picard'][k], r['picard'].sum())
    worst = cm.assert_rows_close(g['rows'][k, :n], r['rows'][:n], g['columns'], what=what)
    assert np.isnan(g['rows'][k, n:]).all(), what + ': rows after the last valid one must be NaN'
    if 'F' in g:
        rows = range(n) if snapshot_rows is None else [x for x in snapshot_rows if x < n]
        for row in rows:
            assert g['bounds'][row][0][k] == r['first'][row] and g['bounds'][row][1][k] == r['last'][row], \
                '{}: first/last at row {}'.format(what, row)
            e = cm.rel_err(g['F'][row][k], r['F'][row])
            assert e.max() <= cm.RTOL, '{}: F at row {} off by {}'.format(what, row, e.max())
    return worst


@pytest.mark.parametrize('path', cm.golden_files(), ids=lambda p: p.split('/')[-1])
def test_golden_configurations(ref, path):
    """The reference's 8 regression configurations (python/test/data/*.dat): every row, F at every row."""
    kw, lam, pb, data = pyoracle.load_golden(path)
    from freddi_b200.args import make_args
    a = make_args(False, **kw)
    passbands = cm.golden_passbands(pb)
    g = run_gpu([a], lam, passbands=passbands)
    assert g['columns'] == list(data.dtype.names)
    r = ref.run(a, lam, passbands=passbands)
    check_model(g, 0, r, path.split('/')[-1])
    for pbn in passbands:  # the golden's own passband columns (6 printed digits)
        np.testing.assert_allclose(g['rows'][0, :len(data), g['columns'].index('Fnu' + pbn.name)], data['Fnu' + pbn.name], rtol=6e-6)
    # and the reference's own criterion against the golden file (python/test/test_regression.py:64-98)
    n = min(len(data), r['rows_valid'])
    for col in ('Mdot', 'Mdisk', 'Sigmaout', 'Teffout', 'Tirrout', 'Lx', 'Fx'):
        np.testing.assert_allclose(g['rows'][0, :n, g['columns'].index(col)], data[col][:n], rtol=1e-4)
    np.testing.assert_allclose(g['rows'][0, :n, 0], data['t'][:n], rtol=1e-13, atol=0)
    assert (g['rows'][0, :n, 0] == r['rows'][:n, 0]).all()  # t is accumulated (t += tau), bit-exact


def test_config2_alpha_cirr_sweep(ref):
    """BASELINE config #2 shape (alpha x Cirr, irradiation, shrinking hot zone), reduced to 4 x 3 models."""
    args = cm.config2_args(np.linspace(0.1, 1.0, 4), np.geomspace(1e-5, 5e-3, 3))
    g = run_gpu(args, cm.LAM3)
    for k, a in enumerate(args):
        check_model(g, k, ref.run(a, cm.LAM3), 'config2[{}]'.format(k), snapshot_rows=(0, 100, 200))


def test_config3_neutron_star_sweep(ref):
    args = cm.config3_args(np.geomspace(10 ** 7.5, 10 ** 8.5, 3), np.linspace(0.1, 1.0, 2))
    g = run_gpu(args, cm.LAM3)
    for k, a in enumerate(args):
        check_model(g, k, ref.run(a, cm.LAM3), 'config3[{}]'.format(k), snapshot_rows=(0, 100, 200))


def test_config4_wind_irradiation_multiband(ref):
    args = cm.config4_args(np.linspace(0.3, 0.9, 3), np.geomspace(1e-4, 2e-3, 2))
    g = run_gpu(args, cm.LAM3)
    for k, a in enumerate(args):
        check_model(g, k, ref.run(a, cm.LAM3), 'config4[{}]'.format(k), snapshot_rows=(0, 100, 200))


def test_config5_four_parameter_sweep(ref):
    args = cm.config5_args([3, 15], [0.1, 1.0], [1e25, 1e27], [1e-5, 5e-3])
    g = run_gpu(args, cm.LAM3)
    for k, a in enumerate(args):
        check_model(g, k, ref.run(a, cm.LAM3), 'config5[{}]'.format(k), snapshot_rows=(0, 100, 200))


NS_MOVING = dict(Mx=1.4, alpha=0.25, Mopt=0.5, period=0.25, F0=None, Mdot0=1e16, initialcond='quasistat', Bx=1e8, freqx=500,
                 Rx=1e6, rin=1.3e6 * 2.99792458e10 ** 2 / (6.673e-8 * 1.4 * 1.98892e33), rout=1e11 / 6.955e10, Rdead=5e6,
                 time=40, tau=0.1, powerorder=None)

EDGE_CASES = {
    # name: (is_ns, kwargs, lambdas, cold_disk)
    'ns_moving_inner_radius_eksi_kutlu': (True, dict(NS_MOVING, fptype='eksi-kutlu2010'), cm.LAM3, False),
    'ns_inverse_beta_collapse': (True, dict(cm.NS_CI, inversebeta=0.5, freqx=300, Rx=1.2e6, fptype='romanova2018',
                                            fpparams={'par1': 0.15, 'par2': 0.92}, kappattype='romanova2018',
                                            kappatparams={'in': 1 / 3, 'out': 1 / 3, 'par1': 0.15, 'par2': 0.92},
                                            nsprop='sibgatullinsunyaev2000', nsgravredshift='on', Cirr=1e-4, Thot=1e4,
                                            boundcond='Tirr', angulardistns='plane', time=30), cm.LAM3, False),
    'ns_quasistat_ns_newt_geometrical': (True, dict(cm.NS_CI, initialcond='quasistat-ns', freqx=400, Rx=1e6, Bx=3e8, nsprop='newt',
                                                    fptype='geometrical', fpparams={'chi': 30.0}, kappattype='corstep',
                                                    kappatparams={'in': 0.3, 'out': 0.4}, time=30), cm.LAM3, False),
    'cold_disk_plane_teff': (False, dict(initialcond='quasistat', Thot=1e4, Cirr=3e-4, Cirrcold=1e-4, h2rcold=0.03, irrindex=0.5,
                                         irrindexcold=0.2, boundcond='Teff', angulardistdisk='plane', inclination=30), cm.LAM3, True),
    'linearF_linear_grid_nx500': (False, dict(initialcond='linearF', gridscale='linear', Nx=500, Mdot0=1e18, F0=None, time=20), (), False),
    'powerSigma_nx300': (False, dict(initialcond='powerSigma', powerorder=1, Mdisk0=1e25, F0=None, Nx=300, time=20), (), False),
    'wind_testA': (False, dict(windtype='__testA__', windparams={'kA': 2.0}, time=20), (), False),
    'wind_testB': (False, dict(windtype='__testB__', windparams={'kB': 16.0}, time=20), (), False),
    'wind_testC': (False, dict(windtype='__testC__', windparams={'kC': 3.0}, Mdotout=-1e18, time=20), (), False),
    'wind_janiuk2015': (False, dict(windtype='Janiuk2015', windparams={'A_0': 10, 'B_1': 0.5}, initialcond='quasistat', time=20), (), False),
    'wind_toy': (False, dict(windtype='toy', windparams={'C_w': 1.0}, initialcond='quasistat', time=20), (), False),
    'wind_shields1986': (False, dict(cm.WOODS, windtype='Shields1986', time=20), cm.LAM3, False),
}


@pytest.mark.parametrize('name', sorted(EDGE_CASES))
def test_edge_cases(ref, name):
    is_ns, kw, lam, cold = EDGE_CASES[name]
    a = cm.ini_args(is_ns, **kw)
    g = run_gpu([a], lam, cold_disk=cold)
    check_model(g, 0, ref.run(a, lam, cold_disk=cold), name)


def test_passbands_with_cold_disk_and_neutron_star(ref):
    """Fnu<passband>[_cold] columns (cpp/src/output.cpp:258-271) in a sweep with a shrinking hot zone, and ahead of the NS columns."""
    pbs = cm.golden_passbands(['passbands/Swift_V.dat', 'passbands/Swift_B.dat'])
    base = dict(initialcond='quasistat', Thot=1e4, Cirrcold=1e-4, h2rcold=0.03, irrindex=0.5, irrindexcold=0.2, boundcond='Tirr',
                angulardistdisk='plane', inclination=30, time=25)
    args = [cm.ini_args(False, **dict(base, Cirr=c, alpha=al)) for c in (1e-4, 3e-4) for al in (0.25, 0.7)]
    g = run_gpu(args, cm.LAM3[:2], cold_disk=True, passbands=pbs)
    assert g['columns'][15:] == ['Fnu0', 'Fnu0_cold', 'Fnu1', 'Fnu1_cold', 'FnuSwift_V', 'FnuSwift_V_cold', 'FnuSwift_B', 'FnuSwift_B_cold']
    for k, a in enumerate(args):
        r = ref.run(a, cm.LAM3[:2], cold_disk=True, passbands=pbs)
        check_model(g, k, r, 'passbands_cold[{}]'.format(k), snapshot_rows=(0, 50, 100))
    assert all(g['rows'][k, g['valid'][k] - 1, -1] > 0 for k in range(len(args)))  # the cold zone exists and shines in Swift_B
    a = cm.ini_args(True, **dict(cm.NS_CI, time=10))
    g = run_gpu([a], (), passbands=pbs[:1])
    assert g['columns'][15:17] == ['FnuSwift_V', 'Rin']
    check_model(g, 0, ref.run(a, (), passbands=pbs[:1]), 'passbands_ns')


def test_star_flux_columns(ref):
    """--starflux: the irradiated companion star at the current phase and the two conjunctions, for
    wavelengths and a passband, BH (two mass ratios = two triangle tables in one ensemble, isotropic and plane disc emission,
    cold disc) and NS (second source, plane NS emission)."""
    pbs = cm.golden_passbands(['passbands/Swift_V.dat'])
    star = dict(initialcond='quasistat', Thot=1e4, boundcond='Tirr', Cirr=3e-4, Topt=4000, inclination=40, staralbedo=0.3, h2rcold=0.03,
                Cirrcold=1e-4, ephemerist0=0.1, time=12)
    args = [cm.ini_args(False, angulardistdisk='isotropic', **star), cm.ini_args(False, angulardistdisk='plane', **dict(star, Mopt=2.0, starlod=2)),
            cm.ini_args(False, angulardistdisk='plane', **dict(star, rochelobefill=0.7, Topt=0))]
    g = run_gpu(args, cm.LAM3[:2], cold_disk=True, passbands=pbs, star_flux=True)
    assert g['columns'][15:20] == ['Fnu0', 'Fnu0_cold', 'Fnu0_star', 'Fnu0_star_min', 'Fnu0_star_max']
    assert g['columns'][25:30] == ['FnuSwift_V', 'FnuSwift_V_cold', 'FnuSwift_V_star', 'FnuSwift_V_star_min', 'FnuSwift_V_star_max']
    for k, a in enumerate(args):
        check_model(g, k, ref.run(a, cm.LAM3[:2], cold_disk=True, passbands=pbs, star_flux=True), 'star[{}]'.format(k), snapshot_rows=(0, 20))
    assert g['rows'][2, 0, 17] == 0 and g['rows'][2, 5, 17] > 0   # Topt = 0: dark until the disc's irradiation sources are set
    a = cm.ini_args(True, **dict(cm.NS_CI, Topt=3500, inclination=60, staralbedo=0.5, angulardistns='plane', Cirr=1e-3, time=5))
    g = run_gpu([a], cm.LAM3[:1], star_flux=True)
    check_model(g, 0, ref.run(a, cm.LAM3[:1], star_flux=True), 'star_ns')


def test_mixed_lengths_and_incremental_evolve(ref):
    """Models with different Nt in one ensemble; evolve() in pieces equals one call; rows past Nt stay NaN."""
    from freddi_b200 import Ensemble
    args = [cm.ini_args(time=10), cm.ini_args(time=5, initialcond='quasistat'), cm.ini_args(time=7.5, alpha=0.6)]
    with Ensemble(args, lambdas=cm.LAM3) as ens:
        assert ens.n_rows == 41
        ens.evolve(0)
        st, valid, _ = ens.status()
        assert valid.tolist() == [1, 1, 1]
        ens.evolve(7)
        assert ens.state()['i_t'].tolist() == [7, 7, 7]
        ens.evolve(20)
        assert ens.state()['i_t'].tolist() == [27, 20, 27]
        ens.evolve(-1)
        rows = ens.rows()
        st, valid, picard = ens.status()
        assert valid.tolist() == [41, 21, 31] and (st == 0).all()
        for k, a in enumerate(args):
            r = ref.run(a, cm.LAM3)
            cm.assert_rows_close(rows[k, :valid[k]], r['rows'][:valid[k]], ens.columns, what='piecewise[{}]'.format(k))
            assert picard[k] == r['picard'].sum()
            assert np.isnan(rows[k, valid[k]:]).all()


def test_fields_and_flux_readout(ref):
    """fb200_ensemble_field / _flux / _mdot_wind against the reference getters on a mid-run state."""
    from freddi_b200 import Ensemble
    a = cm.ini_args(**dict(cm.WOODS, Thot=1e4, boundcond='Tirr', time=10))
    m = ref.model(a, cm.LAM3)
    with Ensemble([a], lambdas=cm.LAM3, record_F=True) as ens:
        ens.evolve(15)
        for _ in range(15):
            assert m.step() == 0
        st = ens.state()
        rs = m.state()
        assert (st['i_t'][0], st['first'][0], st['last'][0]) == (rs['i_t'], rs['first'], rs['last'])
        assert st['t'][0] == rs['t']
        for name in ('h', 'R', 'F', 'W', 'Sigma', 'Tph', 'Tph_vis', 'Tirr', 'Kirr', 'Height', 'Tph_X'):
            e = cm.rel_err(ens.field(name)[0], m.field(name))
            assert e.max() <= cm.RTOL, (name, e.max())
        lam = np.geomspace(1e-5, 1e-4, 7)
        fl = ens.flux(lam, 'hot')[0]
        assert cm.rel_err(fl, [m.flux(x, 0) for x in lam]).max() <= cm.RTOL
        assert cm.rel_err(ens.flux(lam, 'cold')[0], [m.flux(x, 1) for x in lam]).max() <= cm.RTOL
        # recorded row 15 equals the current state
        np.testing.assert_array_equal(ens.field('F', 15), ens.field('F'))


def test_freddi_python_surface(ref):
    """freddi.Freddi-shaped API: iteration, evolve(), EvolutionResult NaN padding (python/test/test_freddi.py:22-44)."""
    from freddi_b200 import Freddi
    kw = dict(cm.INI, initialcond='quasistat', Thot=1e4, boundcond='Tirr', Cirr=2e-4, angulardistdisk='isotropic', time=20)
    fr = Freddi(**kw)
    a = cm.ini_args(**{k: v for k, v in kw.items() if k != '__cgs'})
    r = ref.run(a, ())
    assert fr.Nt == 80 and fr.Nx == 1000
    result = fr.evolve()
    np.testing.assert_array_equal(result.i_t, np.arange(fr.Nt + 1))
    assert cm.rel_err(result.Mdot, r['rows'][:, 1]).max() <= cm.RTOL
    assert cm.rel_err(result.Mdisk, r['rows'][:, 2]).max() <= cm.RTOL
    np.testing.assert_array_equal(result.first, r['first'])
    np.testing.assert_array_equal(result.last, r['last'])
    F = result.F
    assert F.shape == (fr.Nt + 1, fr.Nx)
    for i in range(fr.Nt + 1):
        assert np.isnan(F[i, r['last'][i] + 1:]).all() and not np.isnan(F[i, r['first'][i]:r['last'][i] + 1]).any()
    assert cm.rel_err(result.last_Sigma, r['rows'][:, 4]).max() <= cm.RTOL
    fl = result.flux([5e-5, 8e-5])
    assert fl.shape == (fr.Nt + 1, 2)
    # the batched read-out (one launch for all rows) equals the per-state getters, NaN padding included
    states = list(Freddi(**kw))  # the iterator is one-shot like the reference's; a second object for the per-state path
    assert len(states) == fr.Nt + 1
    for name in ('F', 'Sigma', 'Tph', 'Tirr', 'Height'):
        batched = getattr(result, name)
        for i in (0, 7, fr.Nt):
            per_state = np.full(fr.Nx, np.nan)
            sl = slice(states[i].first, states[i].last + 1)
            per_state[sl] = getattr(states[i], name)[sl]
            np.testing.assert_array_equal(batched[i], per_state)
    np.testing.assert_array_equal(fl[5], states[5].flux([5e-5, 8e-5]))
    m = ref.model(a)
    for i in range(6):
        m.step()
    assert cm.rel_err(result.Tph[6, states[6].first:states[6].last + 1], m.field('Tph')[states[6].first:states[6].last + 1]).max() <= cm.RTOL
    assert cm.rel_err(fl[6], [m.flux(5e-5), m.flux(8e-5)]).max() <= cm.RTOL


def test_python_star_flux_and_regions(ref):
    """Freddi.flux(lambda, region, phase) for star / disk / all (python/freddi/__init__.py:103-137) against the oracle's
    flux_star(lambda, phase), per state and batched over the whole EvolutionResult."""
    import ctypes as C
    from freddi_b200 import Freddi
    kw = dict(cm.INI, initialcond='quasistat', Thot=1e4, boundcond='Tirr', Cirr=3e-4, angulardistdisk='plane', Topt=4500, inclination=35,
              staralbedo=0.2, h2rcold=0.03, Cirrcold=1e-4, time=5)
    fr = Freddi(**kw)
    res = fr.evolve()
    a = cm.ini_args(**{k: v for k, v in kw.items() if k != '__cgs'})
    m = ref.model(a)
    f = ref.lib.ref_flux_star
    f.restype = C.c_double
    f.argtypes = [C.c_void_p, C.c_double, C.c_double]
    lam = np.array([4e-5, 6e-5])
    phases = np.linspace(0.3, 5.0, fr.Nt + 1)
    want_star = np.empty((fr.Nt + 1, 2))
    want_all = np.empty((fr.Nt + 1, 2))
    for i in range(fr.Nt + 1):
        for k, l in enumerate(lam):
            want_star[i, k] = f(m.h, l, phases[i])
            want_all[i, k] = want_star[i, k] + m.flux(l, 0) + m.flux(l, 1)
        if i < fr.Nt:
            m.step()
    got = res.flux(lam, 'star', phases)
    assert got.shape == (fr.Nt + 1, 2) and (got > 0).all()
    assert cm.rel_err(got, want_star).max() <= cm.RTOL
    assert cm.rel_err(res.flux(lam, 'all', phases), want_all).max() <= cm.RTOL
    states = list(Freddi(**kw))
    np.testing.assert_array_equal(states[3].flux(lam, 'star', phases[3]), got[3])
    with pytest.raises(ValueError):
        res.flux(lam, 'star')


def test_invalid_arguments_raise_reference_error_classes():
    from freddi_b200 import Ensemble
    with pytest.raises(RuntimeError):
        Ensemble([cm.ini_args(F0=None)])  # "One of F0, Mdisk0 or Mdot0 must be specified"
    with pytest.raises(ValueError):
        Ensemble([cm.ini_args(opacity='nope')])
    with pytest.raises(NotImplementedError):
        Ensemble([cm.ini_args(Nx=50000)])  # beyond the large-grid build (16384)


def test_piecewise_evolve_is_bit_identical():
    """evolve() cut at arbitrary steps gives the same bits as one call: the Picard state carried between work items (F and
    the closure K of the last iteration) is complete."""
    from freddi_b200 import Ensemble
    args = cm.config2_args([0.2, 0.7], [1e-4, 1e-3], time=20) + [cm.ini_args(False, windtype='Woods1996', windparams={'Xi_max': 10, 'T_ic': 1e8, 'Pow': 1},
                                                                        initialcond='quasistat', Cirr=1e-3, time=20)]
    with Ensemble(args, lambdas=cm.LAM3) as ens:
        ens.evolve(-1)
        whole = ens.rows()
    with Ensemble(args, lambdas=cm.LAM3) as ens:
        for n in (1, 6, 13, 30):
            ens.evolve(n)
        ens.evolve(-1)
        np.testing.assert_array_equal(ens.rows(), whole)


def test_reset_replays_the_ensemble_bit_for_bit():
    """fb200_ensemble_reset: initial conditions copied again from pinned host memory; the replay is bit-identical."""
    from freddi_b200 import Ensemble
    args = cm.config2_args([0.2, 0.9], [1e-4, 2e-3], time=10)
    with Ensemble(args, lambdas=cm.LAM3, record_F=True) as ens:
        ens.evolve(-1)
        rows1, F1 = ens.rows(), ens.field('F', 20)
        st1 = ens.status()
        ens.reset()
        assert ens.state()['i_t'].tolist() == [0, 0, 0, 0]
        ens.evolve(13)
        ens.evolve(-1)
        rows2, F2 = ens.rows(), ens.field('F', 20)
        st2 = ens.status()
    np.testing.assert_array_equal(rows1, rows2)
    np.testing.assert_array_equal(F1, F2)
    for a, b in zip(st1, st2):
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize('Nx', [8, 33, 128, 129, 255, 257, 512, 513, 1024])  # every kernel build and its edges (s128, s256, s512, main)
def test_grid_sizes_from_tiny_to_maximum(ref, Nx):
    """Nx = 1024 is the largest grid the kernels take (one CTA holds the model); tiny grids exercise partial
    blocks of the partitioned solve (block of 1, 2, ... rows) and partial warps."""
    args = [cm.ini_args(Nx=Nx, initialcond='quasistat', time=5), cm.ini_args(Nx=Nx, time=5),
            cm.ini_args(Nx=Nx, initialcond='quasistat', Thot=1e4, Cirr=3e-4, boundcond='Tirr', time=5)]
    g = run_gpu(args, cm.LAM3)
    for k, a in enumerate(args):
        check_model(g, k, ref.run(a, cm.LAM3), 'Nx={}[{}]'.format(Nx, k))


def test_zero_steps_and_single_step_runs(ref):
    """time = 0 gives Nt = 0: one row (the initial state) and no step; time = tau gives exactly one step."""
    args = [cm.ini_args(time=0.0, tau=0.25), cm.ini_args(time=0.25, tau=0.25, initialcond='quasistat')]
    g = run_gpu(args, cm.LAM3)
    assert g['valid'].tolist() == [1, 2]
    for k, a in enumerate(args):
        r = ref.run(a, cm.LAM3)
        cm.assert_rows_close(g['rows'][k, :g['valid'][k]], r['rows'][:r['rows_valid']], g['columns'], what='short[{}]'.format(k))
