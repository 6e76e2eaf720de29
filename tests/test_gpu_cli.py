"""The `freddi-b200` / `freddi-ns-b200` executables (freddi_b200/cpp/freddi_main.cpp) on the GPU: the reference's golden
files are regenerated from their own parameter echo and compared with the committed files — what the reference's CI does
with `generate_test_data.sh` + `git diff --exit-code` (.github/workflows/main.yml)."""
import os
import subprocess

import numpy as np
import pytest

from tests import common as cm
from tests.test_cli import Cli, golden_argv

pytestmark = pytest.mark.gpu

BIN = os.path.join(cm.ROOT, 'build', 'freddi-b200')
BIN_NS = os.path.join(cm.ROOT, 'build', 'freddi-ns-b200')


def run_cli(binary, argv, cwd):
    env = dict(os.environ, XDG_CONFIG_HOME=os.path.join(str(cwd), 'xdg'))
    return subprocess.run([binary] + argv, cwd=str(cwd), env=env, capture_output=True, text=True, timeout=300)


def assert_same_dat(text, golden, last_digit_flips=0):
    """Identical text; `last_digit_flips` numeric tokens may differ by one unit of the last printed digit (a value
    within 1e-10 of a rounding boundary of the 6-digit print)."""
    if text == golden:
        return
    a, b = text.splitlines(), golden.splitlines()
    assert len(a) == len(b)
    flips = 0
    for x, y in zip(a, b):
        if x == y:
            continue
        assert not x.startswith('#'), (x, y)
        for u, v in zip(x.split('\t'), y.split('\t')):
            if u != v:
                assert abs(float(u) / float(v) - 1) < 1.1e-5, (u, v)
                flips += 1
    assert flips <= last_digit_flips, flips


@pytest.mark.parametrize('path', cm.golden_files(), ids=os.path.basename)
def test_executable_regenerates_the_golden_file(built, tmp_path, path):
    argv, golden = golden_argv(path)
    os.symlink(os.path.join(cm.GOLDEN_DIR, 'passbands'), tmp_path / 'passbands')
    p = run_cli(BIN, argv, tmp_path)
    assert p.returncode == 0, p.stderr
    prefix = [a.split('=', 1)[1] for a in argv if a.startswith('--prefix=')][0]
    assert_same_dat((tmp_path / (prefix + '.dat')).read_text(), golden, last_digit_flips=2)


def test_ns_executable_fulldata_and_sweep(built, ref, tmp_path):
    """freddi-ns (the CI command of .github/workflows/main.yml:27) with --fulldata, and a two-axis --sweep: every file equals
    the text rendered from the oracle's rows / fields for the same model."""
    cli = Cli()
    base = ['--Mx=1.4', '--alpha=0.25', '--Mopt=0.5', '--period=0.25', '--Mdot0=1e17', '--initialcond=quasistat', '--Bx=1e8',
            '--distance=10', '--time=5', '--tau=0.25', '--precision=10', '--lambda=5000', '--fulldata', '--tempsparsity=4']
    p = run_cli(BIN_NS, base, tmp_path)
    assert p.returncode == 0, p.stderr
    q = cli.parse(True, base)
    m = ref.model(q['args'], q['lambdas'])
    info = m.info()
    rows = []
    for i_t in range(21):
        rows.append(m.row())
        if i_t % 4 == 0:
            fields = np.stack([m.field(f) for f in ('h', 'R', 'F', 'Sigma', 'Tph', 'Tph_vis', 'Tirr', 'Height', 'Fmagn')])
            st = m.state()
            out = cli.lib.cli_render_fulldata
            import ctypes as C
            buf = C.create_string_buffer(1 << 21)
            n = out(1, np.ascontiguousarray(fields).ctypes.data_as(C.POINTER(C.c_double)), 9, C.c_long(fields.shape[1]),
                    C.c_double(rows[-1][0]), C.c_long(st['first']), C.c_long(st['last']), C.c_uint(10), buf, C.c_size_t(1 << 21))
            want = buf.raw[:n].decode()
            got = (tmp_path / 'freddi_{}.dat'.format(i_t)).read_text()
            assert_same_dat(got, want, last_digit_flips=3)
        m.step()
    want = cli.render(True, base, np.array(rows), info)
    assert_same_dat((tmp_path / 'freddi.dat').read_text(), want, last_digit_flips=1)
    assert not (tmp_path / 'freddi_1.dat').exists()

    sweep = ['--alpha=0.25', '--Mx=5', '--Mopt=0.5', '--period=0.25', '--distance=10', '--time=10', '--tau=0.25', '--F0=2e38',
             '--initialcond=quasistat', '--Thot=1e4', '--boundcond=Tirr', '--angulardistdisk=isotropic', '--Cirr=1e-4', '--precision=8',
             '--prefix=sw', '--sweep', 'alpha=0.2:0.8:3', 'Cirr=1e-4:1e-3:2:log']
    p = run_cli(BIN, sweep, tmp_path)
    assert p.returncode == 0, p.stderr
    k = 0
    for alpha in np.linspace(0.2, 0.8, 3):
        for cirr in (1e-4, 1e-3):
            argv = [a for a in sweep[:-3] if not a.startswith(('--alpha', '--Cirr', '--prefix'))]
            argv += ['--alpha={!r}'.format(float(alpha)), '--Cirr={!r}'.format(float(cirr)), '--prefix=sw-{}'.format(k)]
            q = cli.parse(False, argv)
            r = ref.run(q['args'], (), snapshots=False)
            want = cli.render(False, argv, r['rows'][:r['rows_valid']], ref.model(q['args']).info())
            assert_same_dat((tmp_path / 'sw-{}.dat'.format(k)).read_text(), want, last_digit_flips=1)
            k += 1
    assert k == 6 and not (tmp_path / 'sw-6.dat').exists()


def test_sweep_over_devices_writes_the_same_files(built, tmp_path):
    """--devices: the sweep from ONE process over several GPUs (here the same GPU listed twice, plus 'all') writes, byte for
    byte, the files of the single-device run — light curves through fb200_sweep_run, --fulldata through the sharded handle."""
    sweep = ['--alpha=0.25', '--Mx=5', '--Mopt=0.5', '--period=0.25', '--distance=10', '--time=5', '--tau=0.25', '--F0=2e38',
             '--initialcond=quasistat', '--Thot=1e4', '--boundcond=Tirr', '--angulardistdisk=isotropic', '--Cirr=1e-4', '--precision=12',
             '--lambda=5000', '--sweep', 'alpha=0.2:0.8:5', 'Cirr=1e-4:1e-3:3:log']
    for extra in ([], ['--fulldata', '--tempsparsity=10']):
        dirs = {}
        for name, dev in (('one', []), ('two', ['--devices=0,0']), ('all', ['--devices=all'])):
            d = tmp_path / (name + str(len(extra)))
            d.mkdir()
            p = run_cli(BIN, sweep + extra + dev, d)
            assert p.returncode == 0, p.stderr
            dirs[name] = d
        files = sorted(f.name for f in dirs['one'].iterdir())
        assert len(files) == (15 if not extra else 15 * 4)  # freddi-k.dat (+ _0, _10, _20 with --fulldata)
        for name in ('two', 'all'):
            assert sorted(f.name for f in dirs[name].iterdir()) == files
            for f in files:
                assert (dirs[name] / f).read_bytes() == (dirs['one'] / f).read_bytes(), (name, f)


def test_cpp_host_example_runs_and_sweep_matches(built, tmp_path):
    """examples/ensemble_sweep.cpp (the C++ host over the C ABI, built by __graft_entry__.build()): the ensemble handle on one GPU
    and fb200::sweep over all GPUs give identical rows."""
    exe = os.path.join(cm.ROOT, 'build', 'ensemble_sweep')
    p = subprocess.run([exe, '5', '4'], cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert '# 20 models, 201 rows x 16 columns each' in p.stdout and 'rows identical' in p.stdout
    assert sum(1 for line in p.stdout.splitlines() if line.startswith('model ')) == 20


def test_executable_reports_collapse_and_errors(built, tmp_path):
    base = ['--alpha=0.7', '--Mx=5', '--Mopt=0.5', '--period=0.25', '--distance=10', '--time=50', '--tau=0.25', '--F0=2e38',
            '--initialcond=quasistat', '--Thot=1e4', '--boundcond=Tirr', '--Cirr=1e-4', '--angulardistdisk=isotropic']
    p = run_cli(BIN, base, tmp_path)
    assert p.returncode == 0 and 'Freddi terminated prematurely' in p.stderr and 'Rout <= Rin' in p.stderr
    n_rows = sum(1 for line in (tmp_path / 'freddi.dat').read_text().splitlines() if not line.startswith('#'))
    assert 1 < n_rows < 201
    # the reference's line (cpp/include/application.hpp:31-40): i_t of the step that threw, and freddi->t() — which FreddiState::step
    # had already advanced by tau when truncateOuterRadius threw — at std::cerr's default precision
    import re
    m = re.search(r'i_t = (\d+), t = (\S+) \(days\), reason: Rout <= Rin', p.stderr)
    assert m and int(m.group(1)) == n_rows - 1 and m.group(2) == '{:g}'.format(n_rows * 0.25)
    p = run_cli(BIN, base[1:], tmp_path)
    assert p.returncode == 1 and "the option '--alpha' is required but missing" in p.stderr
