"""The band-integral table (freddi_b200/csrc/lx_table.h): the adaptive Cash-Karp walk of Spectrum::Planck_nu1_nu2
(cpp/src/spectrum.cpp:26-37) is, for a fixed band ratio, a piecewise-analytic function of x1 = h nu1 / kT; the kernels look it up.
CPU tests: the table against the direct walk (host arithmetic through the C ABI, no device), and the engine under the host
policy with and without the table.  GPU tests: the ensemble with the table against the ring-by-ring walk (exact_lx_walk)."""
import ctypes as C

import numpy as np
import pytest

from tests import common as cm
from freddi_b200 import _lib
from freddi_b200.args import kev_to_hertz


def _selftest(e1, e2, n, x_lo=2.0 ** -5, x_hi=256.0, seed=11):
    err, mism, wx, nsub = C.c_double(), C.c_long(), C.c_double(), C.c_int()
    rc = _lib.lib.fb200_lx_table_selftest(kev_to_hertz(e1), kev_to_hertz(e2), x_lo, x_hi, n, seed, C.byref(err), C.byref(mism),
                                          C.byref(wx), C.byref(nsub))
    return rc, err.value, mism.value, wx.value, nsub.value


def test_table_reproduces_the_walk_for_the_default_band():
    """1-12 keV (the reference's default emin / emax): 4e5 temperatures over the table's whole range.  Every look-up lands on the
    piece the walk takes (equal try_step counts: a wrong piece would differ by the quadrature error, ~1e-6) and agrees with the
    walk to the fit tolerance — six orders below the quadrature's own error."""
    rc, err, mism, wx, nsub = _selftest(1.0, 12.0, 400000)
    assert rc == 0
    assert mism == 0
    assert err <= 5e-12, (err, wx)
    assert 100 < nsub <= 2047


@pytest.mark.parametrize('band', [(0.5, 10.0), (2.0, 10.0), (0.3, 10.0), (1.0, 2.5)])
def test_other_band_ratios(band):
    rc, err, mism, wx, nsub = _selftest(band[0], band[1], 150000, seed=5)
    assert rc == 0
    assert mism == 0
    assert err <= 1e-11, (err, wx)


def test_the_hot_range_densely():
    """x1 in [0.5, 3]: where three quarters of the table's pieces sit (the first steps of the walk are rejected or not)."""
    rc, err, mism, wx, nsub = _selftest(1.0, 12.0, 400000, 0.5, 3.0, seed=3)
    assert rc == 0 and mism == 0 and err <= 5e-12, (err, mism, wx)


def test_a_band_without_a_table_is_refused_not_approximated():
    """A 1000:1 band has no tidy piece structure (tens of thousands of pieces): the builder gives up within its budget, the entry
    reports FB200_ERR_UNSUPPORTED and the kernels walk."""
    rc, *_ = _selftest(0.1, 100.0, 10)
    assert rc in (0, _lib.ERR_UNSUPPORTED)


def test_engine_with_and_without_the_table_under_the_host_policy():
    """DiscEngine under the sequential host policy (tests/host_emu): Lx, Fx through the table against the ring-by-ring walk."""
    emu = cm.HostEmu()
    Nt, Nx = 40, 1000
    a = cm.config2_args([0.55], [3e-4], time=Nt * 0.25)[0]
    try:
        emu.lib.emu_set_lx_table(1)
        tab = emu.run(a, Nt, Nx, cm.LAM3)
        emu.lib.emu_set_lx_table(0)
        walk = emu.run(a, Nt, Nx, cm.LAM3)
    finally:
        emu.lib.emu_set_lx_table(1)
    assert tab['rows_valid'] == walk['rows_valid'] == Nt + 1
    lx, fx = 11, 13  # FB200_COL_LX, FB200_COL_FX
    others = [c for c in range(tab['rows'].shape[1]) if c not in (lx, fx)]
    np.testing.assert_array_equal(tab['rows'][:, others], walk['rows'][:, others])
    for c in (lx, fx):
        e = cm.rel_err(tab['rows'][1:, c], walk['rows'][1:, c]).max()
        assert 0 < e <= 5e-12, e  # > 0: the table was really used


@pytest.mark.gpu
def test_ensemble_with_the_table_against_the_exact_walk():
    """64 models of BASELINE configs[1] (alpha x Cirr): default (pruned rings + table) against exact_lx_walk (every ring walked)."""
    from freddi_b200 import Ensemble
    args = cm.config2_args(np.linspace(0.1, 1.0, 8), np.geomspace(1e-5, 5e-3, 8))
    with Ensemble(args, lambdas=cm.LAM3) as ens:
        ens.evolve(-1)
        tab, valid, cols = ens.rows(), ens.status()[1], ens.columns
        wc_tab = ens.work_counters()
    with Ensemble(args, lambdas=cm.LAM3, exact_lx_walk=True) as ens:
        ens.evolve(-1)
        walk = ens.rows()
        np.testing.assert_array_equal(valid, ens.status()[1])
    worst = 0.
    for k in range(len(args)):
        v = valid[k]
        for c in ('Lx', 'Fx'):
            i = cols.index(c)
            worst = max(worst, cm.rel_err(tab[k, :v, i], walk[k, :v, i]).max())
        others = [i for i, c in enumerate(cols) if c not in ('Lx', 'Fx')]
        np.testing.assert_array_equal(tab[k, :v][:, others], walk[k, :v][:, others])
    assert 0 < worst <= 5e-12, worst
    assert wc_tab['lx_trials'].sum() > 0  # the table hands back the walk's try_step counts (roofline bookkeeping)


@pytest.mark.gpu
def test_neutron_star_hot_spot_through_the_table():
    """Lxns / Fxns (one band integral per row, ns_evolution.cpp:398-422) and the disc's Lx of NS models."""
    from freddi_b200 import Ensemble
    args = cm.config3_args((10 ** 7.5, 1e8, 10 ** 8.5), (0.1, 0.55, 1.0))
    with Ensemble(args) as ens:
        ens.evolve(-1)
        tab, valid, cols = ens.rows(), ens.status()[1], ens.columns
    with Ensemble(args, exact_lx_walk=True) as ens:
        ens.evolve(-1)
        walk = ens.rows()
    for k in range(len(args)):
        v = valid[k]
        for c in ('Lx', 'Fx', 'Lxns', 'Fxns'):
            i = cols.index(c)
            assert cm.rel_err(tab[k, :v, i], walk[k, :v, i]).max() <= 5e-12, (k, c)


@pytest.mark.gpu
def test_a_model_with_another_band_walks():
    """One table per ensemble (the first model's band ratio): a model with another emax / emin in the same ensemble is walked ring by
    ring (pruned rings aside), so its Lx equals the exact walk's far below the fit tolerance; the first model goes through the table."""
    from freddi_b200 import Ensemble
    args = cm.config2_args([0.4], [1e-3]) + cm.config2_args([0.4], [1e-3], emax=10)
    with Ensemble(args, lambdas=cm.LAM3) as ens:
        ens.evolve(-1)
        mixed, valid, cols = ens.rows(), ens.status()[1], ens.columns
    with Ensemble(args, lambdas=cm.LAM3, exact_lx_walk=True) as ens:
        ens.evolve(-1)
        walk = ens.rows()
    i = cols.index('Lx')
    e_tab = cm.rel_err(mixed[0, :valid[0], i], walk[0, :valid[0], i]).max()
    e_walk = cm.rel_err(mixed[1, :valid[1], i], walk[1, :valid[1], i]).max()
    assert 1e-14 < e_tab <= 5e-12, e_tab   # table
    assert e_walk <= 1e-14, e_walk         # walked: only the pruning differs
    assert not np.array_equal(mixed[0, :, i], mixed[1, :, i])


def test_random_band_ratios_never_return_a_wrong_piece():
    """Whatever the band ratio, a table that builds agrees with the walk to the fit tolerance at every sampled temperature (a look-up
    on a wrong piece would be off by the quadrature error, ~1e-6); a ratio without a tidy piece structure gets no table at all.
    (The try_step counts may differ by the walk's closing step: t + (nu2 - t) can miss nu2 by an ulp, a step that adds nothing.)"""
    rng = np.random.default_rng(7)
    built = 0
    for r in np.exp(rng.uniform(np.log(1.05), np.log(300.), 12)):
        err, mism, wx, nsub = C.c_double(), C.c_long(), C.c_double(), C.c_int()
        nu1 = 2.4e17
        rc = _lib.lib.fb200_lx_table_selftest(nu1, nu1 * r, 2.0 ** -5, 256.0, 20000, 3, C.byref(err), C.byref(mism), C.byref(wx), C.byref(nsub))
        assert rc in (0, _lib.ERR_UNSUPPORTED), (r, rc)
        if rc == 0:
            built += 1
            assert err.value <= 2e-11, (r, err.value, wx.value)
    assert built >= 8


def test_table_against_the_reference_s_own_band_integral():
    """Spectrum::Planck_nu1_nu2 of the reference itself (oracle/_ref: cpp/src/spectrum.cpp compiled in place, odeint restated by the
    shim) against the table as the kernels evaluate it, at 10^5 temperatures over the table's range and beyond (the ends walk):
    the 1e-10 parity tolerance with two orders to spare at every one of them."""
    from oracle.pyoracle import REF_SO
    ref, emu = C.CDLL(REF_SO), C.CDLL(cm.EMU_SO)
    dp = C.POINTER(C.c_double)
    ref.ref_planck_nu1_nu2.argtypes = [dp, C.c_long, C.c_double, C.c_double, C.c_double, dp]
    emu.emu_planck_tab.argtypes = [dp, C.c_long, C.c_double, C.c_double, C.c_double, dp, C.POINTER(C.c_long)]
    nu1, nu2 = kev_to_hertz(1.0), kev_to_hertz(12.0)
    h_over_k = 6.62606896e-27 / 1.3806504e-16
    n = 100000
    x1 = np.exp(np.random.default_rng(0).uniform(np.log(0.02), np.log(400.), n))
    T = h_over_k * nu1 / x1
    a, b, walked = np.empty(n), np.empty(n), C.c_long()
    ref.ref_planck_nu1_nu2(T.ctypes.data_as(dp), n, nu1, nu2, 1e-4, a.ctypes.data_as(dp))
    emu.emu_planck_tab(T.ctypes.data_as(dp), n, nu1, nu2, 1e-4, b.ctypes.data_as(dp), C.byref(walked))
    inside = (x1 >= 2.0 ** -5) & (x1 < 256.)
    assert walked.value <= (~inside).sum() + 2   # slivers: a few 1e-9 of the range
    rel = np.abs(a - b) / np.abs(a)
    assert np.isfinite(rel).all()
    assert rel.max() <= 1e-11, (rel.max(), x1[rel.argmax()])
    assert np.median(rel) <= 1e-13


def test_switching_the_table_off_and_bad_bands(monkeypatch):
    """FB200_LX_TABLE=0 disables the table for the process (every ring walks); a band with nu2 <= nu1 has none."""
    assert _selftest(12.0, 1.0, 10)[0] == _lib.ERR_UNSUPPORTED
    assert _selftest(1.0, 1.0, 10)[0] == _lib.ERR_UNSUPPORTED
    monkeypatch.setenv('FB200_LX_TABLE', '0')
    assert _selftest(1.0, 12.0, 10)[0] == _lib.ERR_UNSUPPORTED
    monkeypatch.delenv('FB200_LX_TABLE')
    assert _selftest(1.0, 12.0, 10)[0] == 0
