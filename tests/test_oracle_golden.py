"""The oracle (oracle/_ref: the reference's unmodified translation units + oracle/boost_shim) pinned against the
reference's own golden light curves (python/test/data/*.dat, printed with 6 significant digits) with the
reference's own criterion rtol=1e-4 (python/test/test_regression.py:64-98) — in fact it reproduces them to
the print precision — plus the known-answer values of its unit tests and the analytical steady states."""
import os

import numpy as np
import pytest

from oracle import pyoracle
from tests import common as cm
from freddi_b200.args import make_args


@pytest.mark.parametrize('path', cm.golden_files(), ids=os.path.basename)
def test_reference_build_reproduces_golden_light_curves(ref, path):
    kw, lam, pb, data = pyoracle.load_golden(path)
    a = make_args(False, **kw)
    passbands = cm.golden_passbands(pb)
    r = ref.run(a, lam, snapshots=False, passbands=passbands)
    assert r['rows_valid'] == len(data)  # "Number of lines"
    names = pyoracle.column_names(len(lam), passbands=[p.name for p in passbands])
    assert names == list(data.dtype.names)
    np.testing.assert_allclose(r['rows'][:, 0], data['t'], rtol=1e-12)
    for col in ['Mdot', 'Mdisk', 'Rhot', 'Sigmaout', 'Kirrout', 'H2R', 'Teffout', 'Tirrout', 'Qirr2Qvisout', 'TphXmax', 'Lx', 'Lbol',
                'Fx', 'Fbol'] + ['Fnu{}'.format(k) for k in range(len(lam))] + ['Fnu' + p.name for p in passbands]:
        v = r['rows'][:, names.index(col)]
        g = data[col]
        # 6 significant digits in the file: half a unit of the last printed digit
        np.testing.assert_allclose(v, g, rtol=6e-6, atol=0, err_msg='{} {}'.format(os.path.basename(path), col))


def test_mdisk0_normalisation_pins_the_odeint_shim(ref):
    """Mdisk(t=0) = 9.99997e25 for Mdisk0=1e26 (quasistat_Mdisk0.dat, first row): exercises the fixed-step
    Cash-Karp quadrature of the Boost shim."""
    kw, lam, _pb, data = pyoracle.load_golden(os.path.join(cm.GOLDEN_DIR, 'quasistat_Mdisk0.dat'))
    m = ref.model(make_args(False, **kw))
    assert abs(m.row()[2] / 1e26 - 1) < 1e-5
    assert '{:.6g}'.format(m.row()[2]) == '{:.6g}'.format(data['Mdisk'][0])


def test_trapz_known_answer(ref, emu):
    """cpp/test/util.cpp:31-41: trapz of sin(x) on 11 log-spaced points of [2 pi, 3 pi] is 1.9829331321624648 (scipy's value,
    1e-12) and 2 analytically (1e-2).  Asserted on the reference's own trapz (oracle) and on the weighted point sum the engine
    uses for every radial integral (policy_trapz under the host policy; the GPU test drives the same through the C ABI)."""
    import ctypes as C
    N = 11
    x = 2 * np.pi * (3 * np.pi / (2 * np.pi)) ** (np.arange(N) / (N - 1.))
    y = np.sin(x)
    want = 1.9829331321624648
    got_ref = ref.trapz(x, y, 0, N - 1)
    assert abs(got_ref / want - 1) < 1e-12 and abs(got_ref / 2 - 1) < 1e-2
    dp = C.POINTER(C.c_double)
    emu.lib.emu_trapz.restype = C.c_double
    emu.lib.emu_trapz.argtypes = [dp, dp, C.c_int, C.c_int, C.c_int]
    got = emu.lib.emu_trapz(x.ctypes.data_as(dp), y.ctypes.data_as(dp), N, 0, N - 1)
    assert abs(got / want - 1) < 1e-12
    assert emu.lib.emu_trapz(x.ctypes.data_as(dp), y.ctypes.data_as(dp), N, 4, 4) == 0.0 == ref.trapz(x, y, 4, 4)  # first >= last
    assert abs(emu.lib.emu_trapz(x.ctypes.data_as(dp), y.ctypes.data_as(dp), N, 2, 9) / ref.trapz(x, y, 2, 9) - 1) < 1e-15


def test_ss73_subcritical_steady_state(ref):
    """python/test/test_analytical.py:11-26: F -> Mdot (h - h_in), rtol 1e-5 (shortened: 300 steps of 1 d)."""
    Mdot = 1e18
    a = cm.ini_args(initialcond='sineF', Mdot0=Mdot / 2, F0=None, Mdotout=Mdot, time=1000, tau=1)
    m = ref.model(a)
    for _ in range(1000):
        assert m.step() == 0
    h, F = m.field('h'), m.field('F')
    np.testing.assert_allclose(F[1:], Mdot * (h - h[0])[1:], rtol=1e-5)


def test_ns_magnetic_torque_derivatives(ref):
    """python/test/test_ns.py:12-34: the integral of dFmagn_dh is Fmagn, of d2Fmagn_dh2 is dFmagn_dh (rtol 1e-5)."""
    import scipy.interpolate
    a = cm.ini_args(True, **dict(cm.NS_CI, inversebeta=0.5, Bx=1e8, freqx=100, Rx=1e6))
    m = ref.model(a)
    h, Fm, dFm, d2Fm = m.field('h'), m.field('Fmagn'), m.field('dFmagn_dh'), m.field('d2Fmagn_dh2')
    np.testing.assert_allclose(Fm - Fm[0], scipy.interpolate.CubicSpline(h, dFm).antiderivative(1)(h), rtol=1e-5, atol=0)
    np.testing.assert_allclose(dFm - dFm[0], scipy.interpolate.CubicSpline(h, d2Fm).antiderivative(1)(h), rtol=1e-5, atol=0)
