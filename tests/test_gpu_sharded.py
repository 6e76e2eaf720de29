"""In-process multi-GPU handle (fb200_sharded_*) and the streamed one-shot sweep (fb200_sweep_run) against the
single-device ensemble: bit for bit, whatever the sharding, the scheduling groups or the upload order.  A device may be
listed more than once, so the sharding logic is exercised on a one-GPU box too; with >= 2 devices the real thing runs."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import common as cm

pytestmark = pytest.mark.gpu


def _args(n_alpha=12, n_cirr=10, **kw):
    return cm.config2_args(np.linspace(0.1, 1.0, n_alpha), np.geomspace(1e-5, 5e-3, n_cirr), **kw)


def _single(args, **kw):
    from freddi_b200 import Ensemble
    with Ensemble(args, lambdas=cm.LAM3, **kw) as ens:
        ens.evolve(-1)
        return ens.rows(), ens.status(), ens.state(), ens.field('F')


def _same(a, b):
    return a.shape == b.shape and (a.view(np.uint64) == b.view(np.uint64)).all()


def _device_lists():
    from freddi_b200 import _lib
    n = _lib.lib.fb200_device_count()
    lists = [[0, 0], [0, 0, 0]]
    if n >= 2:
        lists += [[0, 1], 'all']
    return lists


def test_sharded_handle_equals_single_device_bit_for_bit():
    from freddi_b200 import Ensemble
    args = _args(time=10.0)
    rows1, (st1, rv1, pi1), state1, F1 = _single(args)
    for devs in _device_lists():
        with Ensemble(args, lambdas=cm.LAM3, device=devs) as ens:
            assert ens.n_models == len(args) and ens.n_shards == (len(devs) if devs != 'all' else ens.n_shards)
            ens.evolve(7)          # in two calls: the state carried between launches is per shard
            ens.evolve(-1)
            rows = ens.rows()
            st, rv, pi = ens.status()
            state = ens.state()
            F = ens.field('F')
            assert _same(rows, rows1), devs
            assert (st == st1).all() and (rv == rv1).all() and (pi == pi1).all()
            assert (state['i_t'] == state1['i_t']).all() and (state['last'] == state1['last']).all() and _same(state['t'], state1['t'])
            assert _same(F, F1)
            # shard d holds the models d, d + G, ...
            G = ens.n_shards
            for d, sh in enumerate(ens.shards()):
                assert _same(sh.rows(), rows1[d::G])
            # reset + evolve again reproduces the run
            ens.reset()
            ens.evolve(-1)
            assert _same(ens.rows(), rows1)
            wc = ens.work_counters()
            assert (wc['picard_iters'] == pi1).all()


def test_sharded_rows_into_pinned_buffer():
    from freddi_b200 import Ensemble, pinned_empty
    args = _args(6, 5, time=5.0)
    rows1 = _single(args)[0]
    with Ensemble(args, lambdas=cm.LAM3, device=[0, 0]) as ens:
        ens.evolve(-1)
        out = pinned_empty((ens.n_models, ens.n_rows, ens.n_cols))
        out[:] = 7.
        ens.rows(out)
        assert _same(out, rows1)


@pytest.mark.parametrize('n_alpha,n_cirr', [(3, 2), (20, 16), (40, 25)])
def test_streamed_sweep_equals_ensemble(n_alpha, n_cirr):
    """fb200_sweep_run (kernel launched before the models are uploaded, groups picked up by watermark, rows read back per
    group) == create + evolve + rows.  40 x 25 is BASELINE configs[1] at full size (two scheduling groups per device)."""
    from freddi_b200 import sweep, sweep_plan
    args = _args(n_alpha, n_cirr, time=50.0 if n_alpha == 40 else 12.5)
    rows1, (st1, rv1, pi1), _, _ = _single(args)
    assert sweep_plan(args, lambdas=cm.LAM3) == rows1.shape[1:]
    for devs in ([0], [0, 0]) + (([0, 1], None) if len(_device_lists()) > 2 else ()):
        res = sweep(args, lambdas=cm.LAM3, devices=devs)
        assert _same(res.rows, rows1), devs
        assert (res.status == st1).all() and (res.rows_valid == rv1).all() and (res.picard_iters == pi1).all()
        assert res.stats['model_steps'] == int((rv1 - 1).sum())
        assert res.stats['launches'] == (len(devs) if devs else res.stats['n_devices'])
    # pageable output buffer, a second time (caches warm)
    out = np.empty(rows1.shape)
    res = sweep(args, lambdas=cm.LAM3, devices=[0], out=out)
    assert res.rows is out and _same(out, rows1)


def test_streamed_sweep_small_groups_many_chunks():
    """Scheduling groups of 16 / 32 models: every group boundary, the staging-slot ring and the per-group read-back are used."""
    code = r'''
import numpy as np
from tests import common as cm
from freddi_b200 import sweep, Ensemble
args = cm.config2_args(np.linspace(0.1, 1.0, 15), np.geomspace(1e-5, 5e-3, 11), time=7.5)
with Ensemble(args, lambdas=cm.LAM3) as ens:
    ens.evolve(-1); rows1 = ens.rows(); st1 = ens.status()
for devs in ([0], [0, 0, 0]):
    res = sweep(args, lambdas=cm.LAM3, devices=devs)
    assert (res.rows.view(np.uint64) == rows1.view(np.uint64)).all(), devs
    assert (res.rows_valid == st1[1]).all()
print('ok')
'''
    env = dict(os.environ, FB200_GROUP0='16', FB200_GROUP='32', PYTHONPATH=cm.ROOT)
    r = subprocess.run([sys.executable, '-c', code], env=env, cwd=cm.ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and 'ok' in r.stdout, r.stdout + r.stderr


def test_streamed_sweep_ns_and_winds():
    """Optional arrays travel through the staging slots too: NS (Fmagn, dFmagn_dh, d2Fmagn_dh2 + static C) and a dynamic wind."""
    from freddi_b200 import sweep
    ns = cm.config3_args(np.geomspace(10 ** 7.5, 10 ** 8.5, 5), np.linspace(0.1, 1.0, 4), time=5.0)
    rows1, (st1, rv1, _), _, _ = _single(ns)
    res = sweep(ns, lambdas=cm.LAM3, devices=[0, 0])
    assert _same(res.rows, rows1) and (res.rows_valid == rv1).all()
    woods = cm.config4_args(np.linspace(0.1, 1.0, 5), np.geomspace(1e-5, 5e-3, 4), time=5.0)
    rows1, (st1, rv1, _), _, _ = _single(woods)
    res = sweep(woods, lambdas=cm.LAM3, devices=[0])
    assert _same(res.rows, rows1) and (res.rows_valid == rv1).all()


def test_streamed_sweep_reports_a_bad_model_and_stays_usable():
    """An argument error in a late model surfaces as the reference's exception class with its index; the launch that was
    already waiting for uploads is withdrawn, and the library keeps working."""
    from freddi_b200 import sweep
    import time
    args = list(_args(20, 20, time=50.0))  # 200 steps: work items of several rounds are in flight when the launch is withdrawn
    bad = 377
    args[bad].opacity = 99  # std::invalid_argument("Wrong opacity") in the reference's OpacityRelated constructor
    t0 = time.perf_counter()
    with pytest.raises(ValueError) as ei:
        sweep(args, lambdas=cm.LAM3, devices=[0, 0])
    assert ei.value.bad_model == bad and 'opacity' in str(ei.value)
    assert time.perf_counter() - t0 < 5., 'a withdrawn launch must not sit out the work-item time-outs'
    args[bad].opacity = 0
    args[3].opacity = 99  # ... and in the very first chunk, of one device only
    with pytest.raises(ValueError) as ei:
        sweep(args, lambdas=cm.LAM3, devices=[0])
    assert ei.value.bad_model == 3
    good = _args(4, 3, time=2.5)
    rows1 = _single(good)[0]
    assert _same(sweep(good, lambdas=cm.LAM3, devices=[0]).rows, rows1)


def test_trim_and_cache_reuse():
    from freddi_b200 import _lib, sweep
    args = _args(4, 4, time=2.5)
    a = sweep(args, lambdas=cm.LAM3, devices=[0]).rows.copy()
    _lib.lib.fb200_trim()
    b = sweep(args, lambdas=cm.LAM3, devices=[0]).rows
    assert _same(a, b)


def test_many_sweeps_of_changing_shape_and_two_threads():
    """The caches (device arenas, pinned staging, host workers) across calls of changing size, model class and device list,
    and two host threads sweeping at once: every result equals the plain ensemble's bit for bit."""
    import threading
    from freddi_b200 import sweep
    rng = np.random.default_rng(0)
    bh = _args(30, 30, time=3.0)     # 900 models to draw from
    ns = cm.config3_args(np.geomspace(10 ** 7.5, 10 ** 8.5, 10), np.linspace(0.1, 1.0, 10), time=3.0)
    want = {}

    def reference(kind, idx):
        key = (kind, tuple(idx))
        if key not in want:
            pool = bh if kind == 'bh' else ns
            want[key] = _single([pool[i] for i in idx])[0]
        return want[key]

    cases = []
    for n in (1, 7, 16, 17, 300, 640, 900, 33, 2, 450):
        kind = 'ns' if n in (7, 33) else 'bh'
        pool = bh if kind == 'bh' else ns
        idx = sorted(rng.choice(len(pool), size=min(n, len(pool)), replace=False).tolist())
        cases.append((kind, idx, [0] if n % 2 else [0, 0]))
    for kind, idx, devs in cases:
        pool = bh if kind == 'bh' else ns
        res = sweep([pool[i] for i in idx], lambdas=cm.LAM3, devices=devs)
        assert _same(res.rows, reference(kind, idx)), (kind, len(idx), devs)

    errors = []

    def worker(which):
        try:
            for kind, idx, devs in cases[which::2]:
                pool = bh if kind == 'bh' else ns
                res = sweep([pool[i] for i in idx], lambdas=cm.LAM3, devices=devs)
                if not _same(res.rows, reference(kind, idx)):
                    errors.append((which, kind, len(idx)))
        except Exception as exc:  # noqa: BLE001
            errors.append((which, repr(exc)))

    for kind, idx, _ in cases:
        reference(kind, idx)  # (filled before the threads start)
    threads = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_per_device_readouts_through_the_multi_gpu_handle():
    """field_rows / flux_rows / mdot_wind / row_bounds / counters / get_state / set_state / result() on a multi-GPU handle gather
    from the shards into the caller's model order: equal to the single-device ensemble."""
    from freddi_b200 import Ensemble
    args = cm.config4_args(np.linspace(0.1, 1.0, 5), np.geomspace(1e-5, 5e-3, 3), time=4.0)  # Woods1996: a dynamic wind
    with Ensemble(args, lambdas=cm.LAM3, record_F=True) as one, Ensemble(args, lambdas=cm.LAM3, record_F=True, device=[0, 0, 0]) as many:
        for e in (one, many):
            e.evolve(-1)
        assert _same(many.field_rows('Sigma'), one.field_rows('Sigma'))
        assert _same(many.flux_rows([4e-5, 6e-5], 'hot'), one.flux_rows([4e-5, 6e-5], 'hot'))
        assert _same(many.mdot_wind(row=7), one.mdot_wind(row=7))
        for a, b in zip(many.row_bounds(5), one.row_bounds(5)):
            assert (a == b).all()
        for a, b in zip(many.counters(), one.counters()):
            assert (a == b).all()
        sa, sb = many.get_state(), one.get_state()
        assert all(_same(np.asarray(sa[k], dtype=np.float64), np.asarray(sb[k], dtype=np.float64)) for k in sa)
        assert _same(many.result().Tph, one.result().Tph) and _same(many.result().Mdot, one.result().Mdot)
        # put both back to the state after 3 steps and continue
        with Ensemble(args, lambdas=cm.LAM3) as src:
            src.evolve(3)
            st = src.get_state()
        for e in (one, many):
            e.reset()
            e.set_state(**st)
            e.evolve(-1)
        assert _same(many.rows(), one.rows()) and np.isnan(one.rows()[:, :4]).all() and np.isfinite(one.rows()[:, 4:, 0]).all()


def test_streamed_sweep_with_passbands_star_and_cold_disk():
    """The second instantiation of the kernel (passband / companion-star columns) and the star-geometry tables through the
    streamed construction, over two shards: rows equal the plain ensemble's bit for bit."""
    from freddi_b200 import Ensemble, sweep
    from freddi_b200.ensemble import Passband
    pb = Passband('box', np.linspace(4e-5, 6e-5, 40), np.linspace(4e-5, 6e-5, 40) * 1.0)
    args = cm.config2_args(np.linspace(0.2, 0.9, 6), np.geomspace(1e-4, 1e-3, 4), time=3.0, Topt=4000, inclination=40, h2rcold=0.03)
    for k, a in enumerate(args):  # two star geometries in one sweep
        if k % 2:
            a.Mopt = 0.8 * a.Mopt
    kw = dict(lambdas=cm.LAM3[:2], cold_disk=True, passbands=[pb], star_flux=True)
    with Ensemble(args, **kw) as ens:
        ens.evolve(-1)
        rows1, cols = ens.rows(), ens.columns
    assert 'Fnubox_star_max' in cols and 'Fnu1_cold' in cols
    for devs in ([0], [0, 0]):
        res = sweep(args, devices=devs, **kw)
        assert res.columns == cols and _same(res.rows, rows1), devs
