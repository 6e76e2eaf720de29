"""The command-line front-end (freddi_b200/cpp/freddi_options.hpp, freddi_output.hpp, freddi_main.cpp): option
parsing and unit conversion against the Python constructor path, and the freddi.dat writer byte for byte against the
reference's golden files (python/test/data/*.dat, written by the reference's `freddi` with --precision=6), with the
rows supplied by the oracle.  The `-m gpu` counterpart runs the real executable (tests/test_gpu_cli.py)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import pyoracle
from tests import common as cm
from freddi_b200.args import FB200Args, FB200ModelInfo, make_args

CLI_SO = os.path.join(cm.ROOT, 'tests', 'host_emu', 'libcli_emu.so')


class Cli:
    def __init__(self):
        self.lib = C.CDLL(CLI_SO)
        self.lib.cli_render_dat.restype = C.c_long
        self.lib.cli_render_fulldata.restype = C.c_long

    @staticmethod
    def _argv(argv):
        arr = (C.c_char_p * len(argv))(*[a.encode() for a in argv])
        return len(argv), arr

    def parse(self, ns, argv):
        n, arr = self._argv(['freddi'] + list(argv))
        a = FB200Args()
        lam = (C.c_double * 16)()
        nl, cold, prec, sparsity, npb = C.c_int(), C.c_int(), C.c_uint(), C.c_uint(), C.c_int()
        err = C.create_string_buffer(512)
        rc = self.lib.cli_parse(int(ns), n, arr, C.byref(a), lam, C.byref(nl), C.byref(cold), C.byref(prec), C.byref(sparsity),
                                C.byref(npb), err, 512)
        if rc != 0:
            raise ValueError(err.value.decode())
        return dict(args=a, lambdas=list(lam[:nl.value]), cold_disk=bool(cold.value), precision=prec.value,
                    tempsparsity=sparsity.value, n_passbands=npb.value)

    def render(self, ns, argv, rows, info, passband_names=()):
        n, arr = self._argv(['freddi'] + list(argv))
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        names = (C.c_char_p * max(1, len(passband_names)))(*[p.encode() for p in passband_names])
        err = C.create_string_buffer(512)
        cap = 1 << 22
        out = C.create_string_buffer(cap)
        m = self.lib.cli_render_dat(int(ns), n, arr, names, len(passband_names), rows.ctypes.data_as(C.POINTER(C.c_double)),
                                    C.c_long(rows.shape[0]), C.c_double(info.rout), C.c_double(info.risco), C.c_double(info.tau),
                                    out, C.c_size_t(cap), err, C.c_size_t(512))
        if m < 0:
            raise ValueError(err.value.decode())
        assert m <= cap
        return out.raw[:m].decode()


@pytest.fixture(scope='module')
def cli(built):
    return Cli()


def golden_argv(path):
    """The golden's own "### Parameters" echo as a command line (every option, defaults included)."""
    with open(path) as f:
        content = f.read()
    block = re.search(r'### Parameters\n(.+?)### Derived', content, flags=re.DOTALL).group(1)
    argv = []
    for line in block.splitlines():
        m = re.match(r'^# ([\w-]+)=(.*?)(?:  # \d+)?$', line)
        argv.append('--{}={}'.format(m.group(1), m.group(2)))
    return argv, content


NS_ONLY = ('nsprop', 'freqx', 'Rx', 'Bx', 'hotspotarea', 'epsilonAlfven', 'inversebeta', 'Rdead', 'fptype', 'kappattype',
           'nsgravredshift', 'fpparams', 'kappatparams', 'angulardistns')


def args_equal(a, b):
    for name, _t in FB200Args._fields_:
        if not a.is_ns and name in NS_ONLY:
            continue
        x, y = getattr(a, name), getattr(b, name)
        if hasattr(x, '__len__'):
            x, y = list(x), list(y)
        else:
            x, y = [x], [y]
        for u, v in zip(x, y):
            if isinstance(u, float) and np.isnan(u) and np.isnan(v):
                continue
            assert u == v, (name, u, v)


@pytest.mark.parametrize('path', cm.golden_files(), ids=os.path.basename)
def test_cli_reproduces_the_golden_file_byte_for_byte(ref, cli, path):
    argv, content = golden_argv(path)
    p = cli.parse(False, argv)
    kw, lam, pb, _data = pyoracle.load_golden(path)
    a = make_args(False, **kw)
    args_equal(p['args'], a)   # the CLI's unit conversion == the Python constructor's (__cgs=False)
    assert p['lambdas'] == lam and p['precision'] == 6 and p['n_passbands'] == len(pb)
    passbands = cm.golden_passbands(pb)
    r = ref.run(p['args'], p['lambdas'], snapshots=False, passbands=passbands)
    info = ref.model(p['args']).info()
    text = cli.render(False, argv, r['rows'][:r['rows_valid']], info, [q.name for q in passbands])
    assert text == content


def test_ini_file_command_line_priority_and_composition(cli, tmp_path, monkeypatch):
    """parseOptions (cpp/include/options.hpp:91-135): command line first, then --config, then ./freddi.ini; the first
    stored value wins; lambda composes across sources; comments and positional tokens are ignored."""
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv('XDG_CONFIG_HOME', str(tmp_path / 'xdg'))
    (tmp_path / 'freddi.ini').write_text(open(os.path.join(cm.GOLDEN_DIR, 'freddi.ini')).read() + '\nlambda=8000 # 0\n')
    (tmp_path / 'extra.ini').write_text('alpha = 0.5  # from --config\nMx=7\nlambda=5000\n')
    p = cli.parse(False, ['--config', 'extra.ini', '-M', '9', 'period=1', '--lambda', '3000', '4000', '--Thot=1e4'])
    a = p['args']
    assert a.alpha == 0.5                      # extra.ini beats freddi.ini's 0.25
    assert a.Mx == 9 * 1.98892e33              # the command line beats both files
    assert a.period == 0.25 * 86400            # the positional token `period=1` is dropped
    assert a.F0 == 2e38 and a.powerorder == 6 and a.time == 50 * 86400 and a.tau == 0.25 * 86400 and a.Thot == 1e4
    assert p['lambdas'] == [3000e-8, 4000e-8, 5000e-8, 8000e-8]
    with pytest.raises(ValueError, match="option '--alpha' cannot be specified more than once"):
        cli.parse(False, ['--alpha=1', '--alpha=2'])
    with pytest.raises(ValueError, match='unrecognised option'):
        cli.parse(False, ['--alfa=1'])
    assert cli.parse(False, ['--alphac=0.01'])['args'].alphacold == 0.01   # unambiguous prefix (allow_guessing)
    with pytest.raises(ValueError, match='ambiguous'):
        cli.parse(False, ['--wind=no'])


def test_cli_errors_of_the_reference(cli, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv('XDG_CONFIG_HOME', str(tmp_path / 'xdg'))
    base = ['--alpha=0.25', '--Mx=5', '--Mopt=0.5', '--period=0.25', '--distance=10', '--time=50', '--F0=2e38', '--powerorder=6']
    cli.parse(False, base)
    with pytest.raises(ValueError, match="the option '--Mopt' is required but missing"):
        cli.parse(False, [x for x in base if 'Mopt' not in x])
    with pytest.raises(ValueError, match='Set positive --Cirr when --boundcond=Tirr'):   # cpp/src/options.cpp:260-262
        cli.parse(False, base + ['--boundcond=Tirr'])
    with pytest.raises(ValueError, match='--windXi_max is required if --windtype=Woods1996'):
        cli.parse(False, base + ['--windtype=Woods1996'])
    with pytest.raises(ValueError, match='Unknown --windtype=Cambier2013'):               # not reachable from the CLI
        cli.parse(False, base + ['--windtype=Cambier2013'])
    with pytest.raises(ValueError, match='Invalid --gridscale value'):
        cli.parse(False, base + ['--gridscale=sqrt'])
    with pytest.raises(ValueError, match="the argument \\('abc'\\) for option '--alpha' is invalid"):
        cli.parse(False, ['--alpha=abc'] + base[1:])
    with pytest.raises(ValueError, match="required but missing"):
        cli.parse(True, base)                                                              # freddi-ns needs --Bx
    ns = cli.parse(True, base + ['--Bx=1e8', '--fptype=geometrical', '--fp-geometrical-chi=30', '--kappattype=corstep',
                                 '--kappat-corstep-in=0.3', '--initialcond=quasistat-ns', '--angulardistns=plane'])['args']
    assert ns.is_ns == 1 and ns.Bx == 1e8 and ns.fptype == 5 and ns.fpparams[0] == 30 and ns.kappattype == 1
    assert ns.kappatparams[0] == 0.3 and abs(ns.kappatparams[1] - 1 / 3) < 1e-16 and ns.initialcond == 6 and ns.angulardistns == 0


def test_sweep_extension(cli, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv('XDG_CONFIG_HOME', str(tmp_path / 'xdg'))
    base = ['--alpha=0.25', '--Mx=5', '--Mopt=0.5', '--period=0.25', '--distance=10', '--time=50', '--F0=2e38', '--powerorder=6']
    n, arr = cli._argv(['freddi'] + base + ['--sweep', 'alpha=0.1:1.0:40', 'Cirr=1e-5:5e-3:25:log'])
    vals = (C.c_double * 64)()
    err = C.create_string_buffer(512)
    assert cli.lib.cli_n_sweep_models(0, n, arr, vals, 64, err, 512) == 1000
    np.testing.assert_allclose(vals[:40], np.linspace(0.1, 1.0, 40), rtol=1e-15)
    n, arr = cli._argv(['freddi'] + base + ['--sweep', 'opacity=1:2:3'])
    assert cli.lib.cli_n_sweep_models(0, n, arr, vals, 64, err, 512) == -1 and b'not a scalar' in err.value
