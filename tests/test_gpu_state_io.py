"""State in / out of an ensemble and its device-resident results: checkpoint -> resume is bit-identical to an uninterrupted
run (the reference's copyable FreddiState, cpp/src/freddi_state.cpp:84-91), explicit get_state / set_state (the
optional h/F upload), zero-copy views of rows / F / Fsnap in HBM, and the ensemble-level EvolutionResult
(python/freddi/evolution_result.py:24-71)."""
import numpy as np
import pytest

from tests import common as cm

pytestmark = pytest.mark.gpu


def _same(a, b):
    return a.shape == b.shape and (a.view(np.uint64) == b.view(np.uint64)).all()


def _cases():
    return {
        'bh_irr': cm.config2_args(np.linspace(0.1, 1.0, 7), np.geomspace(1e-5, 5e-3, 5), time=10.0),
        'woods': cm.config4_args(np.linspace(0.1, 1.0, 4), np.geomspace(1e-5, 5e-3, 3), time=7.5),
        'ns': cm.config3_args(np.geomspace(10 ** 7.5, 10 ** 8.5, 4), np.linspace(0.1, 1.0, 3), time=7.5),
    }


@pytest.mark.parametrize('case', ['bh_irr', 'woods', 'ns'])
@pytest.mark.parametrize('cut', [0, 1, 13])
def test_checkpoint_resume_is_bit_identical(case, cut):
    from freddi_b200 import Ensemble
    args = _cases()[case]
    with Ensemble(args, lambdas=cm.LAM3, record_F=True) as ens:
        ens.evolve(-1)
        rows1, status1, F1 = ens.rows(), ens.status(), ens.field('F')
        wc1 = ens.work_counters()
        mw1 = ens.mdot_wind(row=cut + 1)
    with Ensemble(args, lambdas=cm.LAM3, record_F=True) as a:
        a.evolve(cut)
        blob = a.checkpoint(with_rows=True)
    with Ensemble(args, lambdas=cm.LAM3, record_F=True) as b:
        b.restore(blob)
        st = b.state()
        assert (st['i_t'] == np.minimum(cut, status1[1] - 1)).all()
        b.evolve(-1)
        assert _same(b.rows(), rows1)
        s = b.status()
        assert all((x == y).all() for x, y in zip(s, status1))
        assert _same(b.field('F'), F1)
        wc = b.work_counters()
        assert all((wc[k] == wc1[k]).all() for k in wc)
        if case == 'woods':  # the wind record of every row travelled with the rows
            assert _same(b.mdot_wind(row=cut + 1), mw1)
    # a checkpoint of another shape is refused
    with Ensemble(args[:-1], lambdas=cm.LAM3, record_F=True) as c:
        with pytest.raises(ValueError):
            c.restore(blob)
        with pytest.raises(ValueError):
            c.restore(blob[:40])


def test_explicit_state_round_trip():
    """get_state after 9 steps -> set_state into a fresh ensemble -> the continuation agrees to rounding (the carried closure
    is not part of the explicit state), rows before the cut stay unwritten."""
    from freddi_b200 import Ensemble
    args = _cases()['bh_irr']
    with Ensemble(args, lambdas=cm.LAM3) as ens:
        ens.evolve(-1)
        rows1, (st1, rv1, _) = ens.rows(), ens.status()
    with Ensemble(args, lambdas=cm.LAM3) as a:
        a.evolve(9)
        state = a.get_state()
        assert state['F'].shape == (len(args), 1000) and (state['i_t'] == 9).all()
    with Ensemble(args, lambdas=cm.LAM3) as b:
        b.set_state(**state)
        b.evolve(-1)
        rows = b.rows()
        st, rv, _ = b.status()
        assert (rv == rv1).all() and (st == st1).all()
        assert np.isnan(rows[:, :10]).all()
        for k in range(len(args)):
            n = rv1[k]
            rel = np.abs(rows[k, 10:n] - rows1[k, 10:n]) / np.abs(rows1[k, 10:n])
            assert np.nanmax(np.where(rows1[k, 10:n] == rows[k, 10:n], 0., rel)) < cm.RTOL


def test_initial_state_upload_is_bit_identical():
    """The resolved t = 0 state uploaded explicitly (optional F[Nx] upload) reproduces the default run bit for
    bit; a different F changes it."""
    from freddi_b200 import Ensemble
    from freddi_b200.ensemble import resolve
    args = _cases()['bh_irr'][:6]
    F0 = np.stack([resolve(a)[2] for a in args])
    with Ensemble(args, lambdas=cm.LAM3) as ens:
        ens.evolve(-1)
        rows1 = ens.rows()
    with Ensemble(args, lambdas=cm.LAM3) as b:
        s0 = b.get_state()
        assert _same(s0['F'], F0) and (s0['i_t'] == 0).all() and np.isneginf(s0['Mdot_in_prev']).all()
        b.set_state(F=F0, i_t=0)
        b.evolve(-1)
        assert _same(b.rows(), rows1)
        b.reset()
        b.set_state(F=0.5 * F0)
        b.evolve(-1)
        col = b.columns.index('Mdot')
        r = b.rows()
        assert np.allclose(r[:, 0, col], 0.5 * rows1[:, 0, col], rtol=1e-12)
        with pytest.raises(ValueError):
            b.set_state(first=999, last=999)


def test_device_resident_views_are_zero_copy():
    import torch
    from freddi_b200 import Ensemble
    args = _cases()['bh_irr'][:10]
    with Ensemble(args, lambdas=cm.LAM3, record_F=True) as ens:
        ens.evolve_async(-1)
        ens.sync()
        dev = ens.rows_device()
        cai = dev.__cuda_array_interface__
        assert cai['shape'] == (10, ens.n_rows, ens.n_cols) and cai['typestr'] == '<f8' and cai['data'][0] == ens.rows_device_ptr()
        t = torch.as_tensor(dev, device='cuda')
        assert t.data_ptr() == ens.rows_device_ptr() and t.shape == (10, ens.n_rows, ens.n_cols) and t.dtype == torch.float64
        assert _same(t.cpu().numpy(), ens.rows())
        # a reduction on the device without the rows ever leaving HBM: peak Mdot of every light curve
        col = ens.columns.index('Mdot')
        peak = torch.nan_to_num(t[:, :, col], nan=-1.).max(dim=1).values.cpu().numpy()
        assert (peak == np.nanmax(ens.rows()[:, :, col], axis=1)).all()
        F = ens.device_array('F').to_torch()
        assert _same(F.cpu().numpy(), ens.field('F'))
        Fs = ens.device_array('Fsnap').to_torch().cpu().numpy()
        bounds = ens.device_array('bounds').to_torch().cpu().numpy()
        _, rv, _ = ens.status()
        want = ens.field_rows('F', pad_nan=False)
        for k in range(10):
            n = rv[k]
            assert _same(Fs[k, :n], want[k, :n]) and np.isnan(Fs[k, n:]).all()
            first, last = ens.row_bounds(n - 1)
            assert bounds[k, n - 1, 0] == first[k] and bounds[k, n - 1, 1] == last[k]
        assert _same(ens.device_array('h').to_torch().cpu().numpy(), ens.field('h'))
    with Ensemble(args, lambdas=cm.LAM3) as plain:
        with pytest.raises(RuntimeError):
            plain.device_array('Fsnap')


def test_ensemble_result_mirrors_evolution_result():
    from freddi_b200 import Ensemble
    args = _cases()['bh_irr'][:8]
    with Ensemble(args, lambdas=cm.LAM3, record_F=True) as ens:
        ens.evolve(-1)
        res = ens.result()
        rows = ens.rows()
        _, rv, _ = ens.status()
        assert _same(res.Mdot, rows[:, :, ens.columns.index('Mdot')]) and res.t.shape == (8, ens.n_rows)
        Sigma = res.Sigma
        assert Sigma.shape == (8, ens.n_rows, 1000)
        for k in range(8):
            n = rv[k]
            assert np.isnan(Sigma[k, n:]).all()
            first, last = ens.row_bounds(n - 1)
            assert np.isfinite(Sigma[k, n - 1, first[k]:last[k] + 1]).all()
            assert np.isnan(Sigma[k, n - 1, last[k] + 1:]).all() and np.isnan(Sigma[k, n - 1, :first[k]]).all()
        fl = res.flux(np.array([[3e-5, 5e-5], [8e-5, 1e-4]]))
        assert fl.shape == (8, ens.n_rows, 2, 2)
        col = ens.columns.index('Fnu0')
        ok = np.isfinite(rows[:, :, col])
        assert np.allclose(fl[:, :, 0, 0][ok], rows[:, :, col][ok], rtol=1e-12)
        with pytest.raises(AttributeError):
            res.no_such_thing
