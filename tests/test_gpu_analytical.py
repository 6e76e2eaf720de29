"""The reference's analytical integration tests (python/test/test_analytical.py) run through the `Freddi` class of this
package on the GPU — same keyword arguments, same closed-form solutions, same tolerances — and, beside each, parity of
the final state with the oracle.  Three of them use Nx = 10^4 (linear grid): the large-grid build of the kernels."""
import numpy as np
import pytest
import scipy.special
from numpy.testing import assert_allclose

from tests import common as cm

pytestmark = pytest.mark.gpu

DAY = 86400
# python/test/test_util.py:29-52: freddi.ini in CGS
DEFAULTS = dict(Mx=5 * 1.98892e33, alpha=0.25, Mopt=0.5 * 1.98892e33, period=0.25 * DAY, F0=2e38, powerorder=6.0, gaussmu=1.0,
                gausssigma=0.25, distance=10 * 1000 * 3.08567758135e18, time=50 * DAY, tau=0.25 * DAY)


def freddi_w_default(**kwargs):
    from freddi_b200 import Freddi
    return Freddi(**{**DEFAULTS, **kwargs})


def last_state(fr, ref=None):
    """Iterates like the reference tests do; with `ref` also checks F and Mdot_wind of the final state against the oracle."""
    state = None
    for state in fr:
        pass
    if ref is not None:
        from freddi_b200.args import kwargs_to_args, merge_kwargs
        a = kwargs_to_args(merge_kwargs(dict(fr._kwargs), False, 'Freddi'), False)
        m = ref.model(a)
        for _ in range(int(m.info().Nt)):
            assert m.step() == 0
        assert cm.rel_err(state.F, m.field('F')).max() <= cm.RTOL
        assert state.first == m.state()['first'] and state.last == m.state()['last']
        w = m.mdot_wind()
        assert abs(state.Mdot_wind - w) <= 1e-9 * max(abs(w), 1e-300) or w == state.Mdot_wind
    return state


def test_shakura_sunyaev_subcritical(ref):
    """test_analytical.py:11-26"""
    Mdot = 1e18
    state = last_state(freddi_w_default(initialcond='sineF', Mdot0=Mdot / 2, Mdotout=Mdot, time=1000 * DAY, tau=1 * DAY), ref)
    h = state.h
    assert_allclose(state.F, Mdot * (h - h[0]), rtol=1e-5)


def test_shakura_sunyaev_supercritical(ref):
    """test_analytical.py:29-62"""
    Mx = 2e34
    GM = 6.673e-8 * Mx
    c = 2.99792458e10
    Rin = 6 * 6.673e-8 * Mx / c ** 2
    Rout = 100 * Rin
    Ledd = 4. * np.pi * 1.67262158e-24 * c / 6.65245893699e-25 * GM
    eta = 1 - np.sqrt(8 / 9)
    Mcrit = Ledd / (c ** 2 * eta)
    state = last_state(freddi_w_default(windtype='SS73C', windparams={}, Mx=Mx, Mdot0=Mcrit, Mdotout=Mcrit * Rout / Rin,
                                        initialcond='sineF', time=100 * DAY, tau=1 * DAY, rout=Rout), ref)
    h = state.h
    F = Mcrit / Rin / (3 * GM) * (h ** 3 - h[0] ** 3)
    assert_allclose(state.F, F, rtol=1e-5)
    assert_allclose(state.Mdot_wind, Mcrit * (Rout / Rin - 1))


@pytest.mark.parametrize('k_A0', [0.1, 1., 10.])
def test_stationary_wind_A(ref, k_A0):
    """test_analytical.py:65-93 — Nx = 10000, linear grid"""
    Mdot = 1e18
    state = last_state(freddi_w_default(windtype='__testA__', windparams={'kA': k_A0}, Mdotout=Mdot, initialcond='sineF', Mdot0=Mdot,
                                        time=1000 * DAY, tau=1 * DAY, Nx=10000, gridscale='linear'), ref if k_A0 == 1. else None)
    h = state.h
    assert h.shape == (10000,)
    F = (Mdot * (h[-1] - h[0]) * np.sqrt(np.pi / 2 / k_A0) * np.exp(-k_A0 / 2)
         * scipy.special.erfi(np.sqrt(k_A0 / 2) * (h - h[0]) / (h[-1] - h[0])))
    assert_allclose(state.F, F, rtol=1e-6)
    assert_allclose(state.Mdot_wind, Mdot * (1 - np.exp(-k_A0 / 2)))


@pytest.mark.parametrize('k_B0', [16., 25., 36.])
def test_stationary_wind_B(ref, k_B0):
    """test_analytical.py:96-125 — Nx = 10000, linear grid"""
    Mdot = 1e18
    state = last_state(freddi_w_default(windtype='__testB__', windparams={'kB': k_B0}, Mdotout=Mdot, initialcond='sineF', Mdot0=Mdot,
                                        time=1000 * DAY, tau=1 * DAY, Nx=10000, gridscale='linear'), ref if k_B0 == 25. else None)
    h = state.h
    F = (Mdot / np.sqrt(k_B0) * (h[-1] - h[0]) / np.sinh(np.sqrt(k_B0)) * np.sinh((h - h[0]) / (h[-1] - h[0]) * np.sqrt(k_B0)))
    assert_allclose(state.F, F, rtol=1e-3)
    assert_allclose(state.Mdot_wind, Mdot * (1 - 1 / np.sinh(np.sqrt(k_B0))), rtol=1e-4)


@pytest.mark.parametrize('k_C0', [1.0, 2.0, 3.0])
def test_stationary_wind_C(ref, k_C0, k_hw=0.5):
    """test_analytical.py:128-156"""
    Mdot = 1e18
    state = last_state(freddi_w_default(windtype='__testC__', windparams={'kC': k_C0}, Mdotout=Mdot, initialcond='sineF', Mdot0=Mdot,
                                        time=10000 * DAY, tau=10 * DAY, gridscale='linear'), ref if k_C0 == 2.0 else None)
    h = state.h
    h_w = k_hw * h[-1]
    C0 = k_C0 * Mdot / (h[-1] - h[0])
    F = (Mdot - C0 / 2 * (h[-1] - h_w)) * (h - h[0])
    i = h >= h_w
    F[i] += C0 / 2 * (-((h[-1] - h_w) / (2 * np.pi)) ** 2 * (1 - np.cos(2 * np.pi * (h[i] - h_w) / (h[-1] - h_w))) + (h[i] - h_w) ** 2 / 2)
    assert_allclose(state.F, F, rtol=1e-5)


@pytest.mark.parametrize('opacity,a,kl', [('Kramers', (1.430, -0.460, 0.030), (3.5, 6.0)),
                                          ('OPAL', (1.400, -0.425, 0.025), (11. / 3., 19. / 3.))])
def test_lipunova_shakura_quasi_stationary(ref, opacity, a, kl):
    """test_analytical.py:159-196"""
    state = last_state(freddi_w_default(opacity=opacity, time=1000 * DAY, tau=1 * DAY), ref)
    h, F = state.h, state.F
    xi = h / h[-1]
    f = a[0] * xi + a[1] * xi ** kl[0] + a[2] * xi ** kl[1]
    assert_allclose(F / F[-1], f * (h - h[0]) / (h[-1] - h[0]) / xi, rtol=1e-3)


def test_shields_power_law_wind(ref):
    """test_analytical.py:199-225 — Nx = 10000, linear grid, dynamic wind"""
    k_C, r_wind = 1, 0.9
    Mx = 2e34
    GM = 6.673e-8 * Mx
    Mdotout = 1e18
    rout = 1e11
    Mdotin = Mdotout / (1 + k_C)
    state = last_state(freddi_w_default(windtype='ShieldsOscil1986', windparams={'C_w': k_C, 'R_w': r_wind}, F0=Mdotin * np.sqrt(GM * rout),
                                        Mdotout=Mdotout, rout=rout, Mx=Mx, initialcond='powerF', powerorder=1, time=1000 * DAY,
                                        tau=1 * DAY, Nx=10000, gridscale='linear'), ref)
    h = state.h
    h_w = h[-1] * np.sqrt(r_wind)
    F = Mdotin * (h - h[0])
    i = h > h_w
    F[i] += k_C * Mdotin * (h[i] * np.log(h[i] / h_w) - (h[i] - h_w)) / np.log(h[-1] / h_w)
    assert_allclose(state.F, F, rtol=1e-3)
    assert_allclose(state.Mdot_wind, Mdotout * k_C / (1 + k_C), rtol=1e-3)


def test_large_grid_rows_against_oracle(ref):
    """Every row and F snapshots of a Nx = 3000 / 16384 run (irradiation, shrinking hot zone, band fluxes) on the large-grid build."""
    from tests.test_gpu_parity import run_gpu, check_model
    for Nx, time in ((3000, 15), (16384, 3)):
        a = cm.ini_args(False, initialcond='quasistat', Thot=1e4, boundcond='Tirr', angulardistdisk='isotropic', Cirr=3e-4, Nx=Nx, time=time)
        g = run_gpu([a, cm.ini_args(False, Nx=Nx, time=time)], cm.LAM3)
        check_model(g, 0, ref.run(a, cm.LAM3), 'large_grid_{}'.format(Nx), snapshot_rows=(0, 5, 12))
    from freddi_b200 import Ensemble
    with pytest.raises(NotImplementedError):
        Ensemble([cm.ini_args(False, Nx=16385)])
