"""GPU parity for the reference branches round 1 left without a test: Woods1996AGN wind (cpp/src/freddi_state.cpp:546-586),
the `propeller` and `corotation-block` accretion fractions (cpp/src/ns/ns_evolution.cpp:144-225), the reference's own 2000-step
moving-inner-radius behaviour tests for every fp type (python/test/test_ns.py:37-97), Cambier2013 (cpp/src/freddi_state.cpp:403-420),
the wind getters of dynamic winds at a mid-run state, the work-item give-up path of the scheduler, and the radial trapezoid's
known answer (cpp/test/util.cpp:31-41).  All through the C ABI; rows + F + Picard counts against oracle/_ref at 1e-10."""
import ctypes as C
import os

import numpy as np
import pytest

from freddi_b200.args import FB200ModelInfo
from tests import common as cm
from tests.test_gpu_parity import check_model, run_gpu

pytestmark = pytest.mark.gpu

AGN = {
    'agn_strong_wind': dict(cm.WOODS, windtype='Woods1996AGN', windparams={'C_0': 100., 'T_ic': 1e8}, time=20),
    'agn_shrinking_hot_zone': dict(cm.WOODS, windtype='Woods1996AGN', windparams={'C_0': 10., 'T_ic': 1e7}, Thot=1e4, boundcond='Tirr', time=20),
    # le > 0.1 (f_L < 1) and le > 0.01 (R_tr from the logarithms), R > R_tr (g_R > 1): a bright, large disc
    'agn_super_eddington_branches': dict(cm.WOODS, windtype='Woods1996AGN', windparams={'C_0': 1., 'T_ic': 1e9}, F0=6e38, time=10),
}


@pytest.mark.parametrize('name', sorted(AGN))
def test_woods1996agn_wind(ref, name):
    a = cm.ini_args(False, **AGN[name])
    g = run_gpu([a], cm.LAM3)
    r = ref.run(a, cm.LAM3)
    check_model(g, 0, r, name)
    assert r['picard'][1:r['rows_valid']].min() >= 2  # C != 0: the first convergence test always fails (T1 of SURVEY.md 9)


# python/test/test_ns.py:37-55 in CLI units (Mx [Msun], rin [Rg], rout [Rsun], time/tau [d])
RG_CM = 6.673e-8 * 1.4 * 1.98892e33 / 2.99792458e10 ** 2
NS_FP = dict(cm.INI, Mx=1.4, F0=None, Mdot0=1e16, initialcond='quasistat', Rx=1e6, freqx=500, rin=1.3e6 / RG_CM, rout=1e11 / 6.955e10,
             Bx=1e8, Rdead=5e6, time=200, tau=0.1, powerorder=None)
NS_FP.pop('__cgs')
FP_CASES = {
    'no-outflow': dict(fptype='no-outflow'),
    'propeller': dict(fptype='propeller'),
    'propeller-unity-fp': dict(fptype='propeller', rin=1e6 / RG_CM),
    'corotation-block': dict(fptype='corotation-block'),
    'geometrical-chi0': dict(fptype='geometrical', fpparams={'chi': 0.}),
    'geometrical-chi30': dict(fptype='geometrical', fpparams={'chi': 30.}),
    'eksi-kutlu2010': dict(fptype='eksi-kutlu2010'),
    'romanova2018': dict(fptype='romanova2018', fpparams={'par1': 0.15, 'par2': 0.92}),
}


@pytest.fixture(scope='module')
def fp_runs(ref):
    """All fp types in ONE ensemble (2000 steps each, the inner radius moves out through the corotation radius)."""
    from freddi_b200 import Ensemble
    names = sorted(FP_CASES)
    args = [cm.ini_args(True, **dict(NS_FP, **FP_CASES[n])) for n in names]
    with Ensemble(args, lambdas=cm.LAM3, record_F=True) as ens:
        ens.evolve(-1)
        g = dict(rows=ens.rows(), columns=ens.columns)
        g['status'], g['valid'], g['picard'] = ens.status()
        g['F'] = {row: ens.field('F', row) for row in (0, 500, 1000, 2000)}
        g['bounds'] = {row: ens.row_bounds(row) for row in (0, 500, 1000, 2000)}
        info = ens.info()
        g['R_cor'] = [info[k].R_cor for k in range(len(args))]
    return names, args, g


@pytest.mark.parametrize('name', sorted(FP_CASES))
def test_ns_fp_types_2000_steps_against_oracle(ref, fp_runs, name):
    names, args, g = fp_runs
    k = names.index(name)
    r = ref.run(args[k], cm.LAM3)
    assert r['rows_valid'] == 2001
    check_model(g, k, r, name, snapshot_rows=(0, 500, 1000, 2000))
    assert r['first'][2000] > r['first'][0]  # the magnetosphere pushed the inner edge outwards


def test_ns_fp_behaviour_like_the_reference_suite(fp_runs):
    """The assertions of python/test/test_ns.py:57-97 on the GPU rows."""
    names, args, g = fp_runs
    cols = g['columns']
    fp = {n: g['rows'][names.index(n), :, cols.index('fpin')] for n in names}
    Rin = {n: g['rows'][names.index(n), :, cols.index('Rin')] for n in names}
    R_cor = g['R_cor'][0]
    assert (g['valid'] == 2001).all()
    np.testing.assert_array_equal(fp['no-outflow'], 1)
    np.testing.assert_array_equal(fp['propeller'], 0)
    unity = Rin['propeller-unity-fp'] <= 1e6 * (1 + 1e-12)   # fp must be unity if R_in <= Rx or R_in <= Risco (R = h^2 / GM rounds)
    assert unity.any()
    np.testing.assert_array_equal(fp['propeller-unity-fp'][unity], 1)
    r = Rin['corotation-block'] / R_cor
    assert (r < 1).any() and (r > 1).any()
    np.testing.assert_array_equal(r < 1, fp['corotation-block'] == 1)
    np.testing.assert_array_equal(fp['corotation-block'], fp['geometrical-chi0'])
    f = fp['geometrical-chi30']
    r = Rin['geometrical-chi30'] / R_cor
    assert (f >= 0).all() and (f <= 1).all()
    assert r[0] < 1 and f[0] == 1
    assert (f[r > 1] < 1).all()


def test_cambier2013_with_the_uninitialised_member_defined(ref):
    """Cambier2013Wind reads DiskStructureArguments::Mdot0 (cpp/src/freddi_state.cpp:408,411), a member no constructor of
    the reference sets (cpp/src/arguments.cpp:42-55): the reference's result depends on what the heap held.  This library
    uses the Mdot0 the initial condition normalises.  With the member DEFINED to that value (oracle hook, the
    sources stay unmodified) the reference's own Cambier2013 code and the kernels agree to 1e-10."""
    from freddi_b200.ensemble import resolve
    a = cm.ini_args(False, windtype='Cambier2013', windparams={'kC': 1.0, 'RIC': 1.0}, initialcond='quasistat', time=20)
    info, _h, _F = resolve(a)
    assert info.Mdot0 > 0
    g = run_gpu([a], cm.LAM3)
    ref.define_mdot0(info.Mdot0)
    try:
        r = ref.run(a, cm.LAM3)
        m = ref.model(a, cm.LAM3)
        windC_ref = m.field('windC')
    finally:
        ref.define_mdot0()
    check_model(g, 0, r, 'cambier2013')
    assert (windC_ref <= 0).all() and (windC_ref < 0).any()  # the wind really acts (exp underflows to -0 at the innermost rings)
    from freddi_b200 import Ensemble
    with Ensemble([a]) as ens:
        assert cm.rel_err(ens.field('windC')[0], windC_ref).max() <= cm.RTOL


DYNAMIC_WINDS = {
    'Woods1996': dict(cm.WOODS, Thot=1e4, boundcond='Tirr', time=10),
    'Woods1996AGN': dict(cm.WOODS, windtype='Woods1996AGN', windparams={'C_0': 10., 'T_ic': 1e7}, Thot=1e4, boundcond='Tirr', time=10),
    'Shields1986': dict(cm.WOODS, windtype='Shields1986', time=10),
    'Janiuk2015': dict(windtype='Janiuk2015', windparams={'A_0': 10, 'B_1': 0.5}, initialcond='quasistat', time=10),
    'ShieldsOscil1986': dict(windtype='ShieldsOscil1986', windparams={'C_w': 1.0, 'R_w': 0.5}, initialcond='quasistat', time=10),
    'toy': dict(windtype='toy', windparams={'C_w': 1.0}, initialcond='quasistat', Thot=1e4, Cirr=3e-4, boundcond='Tirr', time=10),
    # NS: windC() adds d2Fmagn_dh2 (cpp/src/ns/ns_evolution.cpp:487-493); the wind's constructor saw the BH Mdot_in formula on the
    # initial F before the NS constructor shifted it (T12 of SURVEY.md 9), every later update the NS formula
    'ns_Woods1996': dict(cm.NS_CI, is_ns=True, inversebeta=0.5, freqx=300, Rx=1.2e6, windtype='Woods1996',
                         windparams={'Xi_max': 10, 'T_ic': 1e8, 'Pow': 1}, Cirr=1e-3, time=10),
}


@pytest.mark.parametrize('wind', sorted(DYNAMIC_WINDS))
def test_dynamic_wind_getters_hold_the_last_update(ref, wind):
    """windA/windB/windC and Mdot_wind of a dynamic wind are what wind_->update() stored inside FreddiState::step BEFORE the
    solve — values of the previous state's Mdot_in, first, last (cpp/src/freddi_state.cpp:147-153,364-383,457-654), rings that
    left [first, last] keep their last value.  Checked at t0 (the constructor's update), mid-run on the current state, and on
    recorded rows."""
    from freddi_b200 import Ensemble
    kw = dict(DYNAMIC_WINDS[wind])
    a = cm.ini_args(kw.pop('is_ns', False), **kw)
    m = ref.model(a, cm.LAM3)
    with Ensemble([a], lambdas=cm.LAM3, record_F=True) as ens:
        want = {}
        for row in range(0, 26):
            if row in (0, 1, 12, 25):
                want[row] = {f: m.field(f) for f in ('windA', 'windB', 'windC')}
                want[row]['mdot_wind'] = m.mdot_wind()
                ens.evolve(row - int(ens.state()['i_t'][0]))
                for f in ('windA', 'windB', 'windC'):
                    e = cm.rel_err(ens.field(f)[0], want[row][f])
                    assert e.max() <= cm.RTOL, (wind, row, f, e.max())
                assert cm.rel_err(ens.mdot_wind()[0], want[row]['mdot_wind']) <= cm.RTOL, (wind, row)
            assert m.step() == 0
        for row, w in want.items():  # the same from the recorded rows, after the run has moved on
            for f in ('windB', 'windC'):
                assert cm.rel_err(ens.field(f, row)[0], w[f]).max() <= cm.RTOL, (wind, row, f)
            assert cm.rel_err(ens.mdot_wind(row)[0], w['mdot_wind']) <= cm.RTOL, (wind, row)
        moved = want[25]['windC'] != want[1]['windC']
        if wind != 'Janiuk2015':
            assert moved.any(), 'the wind term must change with Mdot_in for this test to mean anything'


def test_trapz_known_answer_on_the_device():
    """cpp/test/util.cpp:31-41: trapz of sin(x) on 11 log-spaced points of [2 pi, 3 pi] is 1.9829331321624648 (1e-12); here
    through the CTA-wide weighted sum every radial integral of the kernels uses (fb200_selftest_trapz)."""
    from freddi_b200 import _lib
    N = 11
    x = 2 * np.pi * (3 * np.pi / (2 * np.pi)) ** (np.arange(N) / (N - 1.))
    y = np.sin(x)
    out = C.c_double()
    dp = C.POINTER(C.c_double)
    _lib.check(_lib.lib.fb200_selftest_trapz(0, x.ctypes.data_as(dp), y.ctypes.data_as(dp), N, 0, N - 1, C.byref(out)))
    assert abs(out.value / 1.9829331321624648 - 1) < 1e-12
    assert abs(out.value / 2 - 1) < 1e-2
    # a sub-range and the degenerate range (first >= last gives 0, util.cpp:10-12)
    _lib.check(_lib.lib.fb200_selftest_trapz(0, x.ctypes.data_as(dp), y.ctypes.data_as(dp), N, 3, 8, C.byref(out)))
    assert abs(out.value - np.trapezoid(y[3:9], x[3:9])) < 1e-15
    _lib.check(_lib.lib.fb200_selftest_trapz(0, x.ctypes.data_as(dp), y.ctypes.data_as(dp), N, 5, 5, C.byref(out)))
    assert out.value == 0.0


def test_work_item_gives_up_when_its_producer_never_publishes():
    """evolve.cu: a work item waits for the previous round of its model; past wait_timeout it flags the model and the launch
    ends (FB200_MODEL_FAILED, the rows finished so far stay valid, nothing of the model is overwritten).  Forced here with
    a 1 us timeout: model 1 collapses at once, so its CTA reaches round 1 of model 0 while round 0 is still running."""
    import subprocess
    import sys
    code = r'''
import numpy as np
from freddi_b200 import Ensemble
from tests import common as cm
slow = cm.ini_args(False, initialcond='quasistat', Thot=1e4, boundcond='Tirr', angulardistdisk='isotropic', Cirr=1e-3, alpha=0.5)
fast = cm.ini_args(False, initialcond='quasistat', Thot=1e7, boundcond='Tirr', angulardistdisk='isotropic', Cirr=1e-5, alpha=0.5)
with Ensemble([slow, fast], lambdas=cm.LAM3) as ens:
    ens.evolve(-1)
    st, valid, _ = ens.status()
    rows = ens.rows()
import json
print('RESULT ' + json.dumps([st.tolist(), valid.tolist(), int(np.isfinite(rows[0, :valid[0]]).all()), int(np.isnan(rows[0, valid[0]:]).all())]))
np.save(sys_argv_out, rows[0])
'''
    import json
    outs = {}
    for tag, env in (('normal', {}), ('forced', {'FB200_WAIT_TIMEOUT_US': '1'})):
        path = os.path.join(cm.ROOT, 'gpurun_out', 'abort_{}.npy'.format(tag))
        os.makedirs(os.path.dirname(path), exist_ok=True)
        p = subprocess.run([sys.executable, '-c', 'sys_argv_out = {!r}\n'.format(path) + code], cwd=cm.ROOT, env=dict(os.environ, **env),
                           capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        line = [l for l in p.stdout.splitlines() if l.startswith('RESULT ')][0]
        outs[tag] = (json.loads(line[len('RESULT '):]), np.load(path))
    (st_n, valid_n, fin_n, nan_n), rows_n = outs['normal']
    (st_f, valid_f, fin_f, nan_f), rows_f = outs['forced']
    assert st_n[0] == 0 and valid_n[0] == 201 and st_n[1] == 1 and valid_n[1] < 10
    assert st_f[0] == 2, 'model 0 must be reported FAILED when a work item gave up on it'
    assert 1 <= valid_f[0] < 201 and fin_f == 1 and nan_f == 1
    np.testing.assert_array_equal(rows_f[:valid_f[0]], rows_n[:valid_f[0]])  # what was finished is exactly the normal run's
    assert st_f[1] == st_n[1] and valid_f[1] == valid_n[1]
