"""BASELINE-size runs on the GPU: the full 10^3-model alpha x Cirr sweep of config #2 (and a 10^3 slice of the NS
and wind sweeps), checked through size-independent properties plus a stride-sampled comparison with the oracle."""
import numpy as np
import pytest

from tests import common as cm

pytestmark = pytest.mark.gpu


def _full_config2():
    return cm.config2_args(np.linspace(0.1, 1.0, 40), np.geomspace(1e-5, 5e-3, 25))


@pytest.fixture(scope='module')
def full_run():
    from freddi_b200 import Ensemble
    args = _full_config2()
    with Ensemble(args, lambdas=cm.LAM3) as ens:
        ens.evolve(-1)
        rows = ens.rows()
        status, valid, picard = ens.status()
        cols = ens.columns
    return dict(args=args, rows=rows, status=status, valid=valid, picard=picard, cols=cols)


def test_full_sweep_invariants(full_run):
    r = full_run
    rows, valid, status = r['rows'], r['valid'], r['status']
    n, nrows, ncol = rows.shape
    assert (n, nrows) == (1000, 201)
    assert set(np.unique(status)) <= {0, 1}
    # a model either ran to the end or collapsed (RadiusCollapseException) at rows_valid
    assert ((status == 0) == (valid == nrows)).all()
    tcol = r['cols'].index('t')
    t_expected = np.zeros(nrows)
    t = 0.0
    for k in range(1, nrows):
        t += 0.25 * 86400.
        t_expected[k] = t / 86400.  # accumulated t += tau, then sToDay (cpp/src/freddi_state.cpp:151)
    for k in range(n):
        v = valid[k]
        assert np.isfinite(rows[k, :v]).all(), k
        assert np.isnan(rows[k, v:]).all(), k
        assert (rows[k, :v, tcol] == t_expected[:v]).all(), k
        md = rows[k, :v, r['cols'].index('Mdisk')]
        assert (np.diff(md) < 0).all(), 'the hot-disc mass only decreases in an outburst decay (model {})'.format(k)
        assert (np.diff(rows[k, :v, r['cols'].index('Rhot')]) <= 0).all(), 'the hot zone never grows (model {})'.format(k)
    # Picard iterations: every completed step needs at least 2 (T1 of SURVEY.md 9 does not apply here: C == 0) ...
    steps = valid - 1
    assert (r['picard'] >= steps).all() and (r['picard'] <= 40 * steps).all()


def test_full_sweep_is_deterministic_and_order_independent(full_run):
    """Bit-for-bit: the same model gives the same rows alone, in a small ensemble, or in the full sweep."""
    from freddi_b200 import Ensemble
    r = full_run
    pick = [0, 137, 512, 999]
    with Ensemble([r['args'][k] for k in pick], lambdas=cm.LAM3) as ens:
        ens.evolve(-1)
        small = ens.rows()
    for j, k in enumerate(pick):
        np.testing.assert_array_equal(small[j], r['rows'][k])
    with Ensemble(r['args'], lambdas=cm.LAM3) as ens:
        ens.evolve(50)
        ens.evolve(-1)  # piecewise
        again = ens.rows()
    np.testing.assert_array_equal(again, r['rows'])


def test_full_sweep_sample_against_oracle(full_run, ref):
    """Every 67th model of the sweep (15 models, corners included) against the reference's own code, all rows."""
    r = full_run
    for k in list(range(0, 1000, 67)) + [999]:
        o = ref.run(r['args'][k], cm.LAM3, snapshots=False)
        n = o['rows_valid']
        assert r['valid'][k] == n, k
        assert r['picard'][k] == o['picard'].sum(), k
        cm.assert_rows_close(r['rows'][k, :n], o['rows'][:n], r['cols'], what='config2 full[{}]'.format(k))


def test_exact_lx_walk_agrees_with_pruned_lx(full_run):
    """The default Lx — rings skipped whose total share is bounded by 1e-16 of Lx (disc_engine.h structure_and_row), the others read
    from the table of the walk's result (lx_table.h, <= 2e-12 per ring) — equals the ring-by-ring walk to far better than the
    parity tolerance."""
    from freddi_b200 import Ensemble
    r = full_run
    pick = list(range(0, 1000, 125))
    with Ensemble([r['args'][k] for k in pick], lambdas=cm.LAM3, exact_lx_walk=True) as ens:
        ens.evolve(-1)
        exact = ens.rows()
    for j, k in enumerate(pick):
        v = r['valid'][k]
        for c in ('Lx', 'Fx'):
            i = r['cols'].index(c)
            assert cm.rel_err(r['rows'][k, :v, i], exact[j, :v, i]).max() <= 5e-12, (k, c)
        others = [i for i, c in enumerate(r['cols']) if c not in ('Lx', 'Fx')]
        np.testing.assert_array_equal(r['rows'][k, :v][:, others], exact[j, :v][:, others])


SWEEPS_1000 = {
    'config3_ns': lambda: cm.config3_args(np.geomspace(10 ** 7.5, 10 ** 8.5, 25), np.linspace(0.1, 1.0, 40)),
    'config4_woods': lambda: cm.config4_args(np.linspace(0.1, 1.0, 40), np.geomspace(1e-5, 5e-3, 25)),
}


@pytest.mark.parametrize('name', sorted(SWEEPS_1000))
def test_ns_and_wind_sweeps_1000_models(ref, name):
    """10^3-model slices of configs #3 (NS, Bx x alpha) and #4 (Woods1996 wind + irradiation, 3 bands every row): invariants on
    every model, and every 62nd model (17 models, corners included) against the reference's own code — all rows, all columns,
    row counts and Picard iterations."""
    from freddi_b200 import Ensemble
    args = SWEEPS_1000[name]()
    with Ensemble(args, lambdas=cm.LAM3) as ens:
        ens.evolve(-1)
        rows = ens.rows()
        status, valid, picard = ens.status()
        cols = ens.columns
    assert len(args) == 1000 and set(np.unique(status)) <= {0, 1}
    for k in range(len(args)):
        v = valid[k]
        assert v >= 1 and np.isfinite(rows[k, :v, :3]).all() and np.isnan(rows[k, v:]).all(), k
    assert (picard >= valid - 1).all()
    for k in list(range(0, 1000, 62)) + [999]:
        o = ref.run(args[k], cm.LAM3, snapshots=False)
        n = o['rows_valid']
        assert valid[k] == n, (name, k)
        assert picard[k] == o['picard'].sum(), (name, k)
        cm.assert_rows_close(rows[k, :n], o['rows'][:n], cols, what='{}[{}]'.format(name, k))
