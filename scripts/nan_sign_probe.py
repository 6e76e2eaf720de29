import numpy as np, sys
sys.path.insert(0,'/root/repo')
from freddi_b200 import Ensemble
from oracle import pyoracle
from tests import common as cm
from tests.test_cli import Cli
cli=Cli()
base = ['--Mx=1.4', '--alpha=0.25', '--Mopt=0.5', '--period=0.25', '--Mdot0=1e17', '--initialcond=quasistat', '--Bx=1e8',
            '--distance=10', '--time=5', '--tau=0.25', '--precision=10', '--lambda=5000']
q=cli.parse(True, base)
ref=pyoracle.Ref()
r=ref.run(q['args'], q['lambdas'])
with Ensemble([q['args']], lambdas=q['lambdas'], record_F=True) as e:
    e.evolve(-1)
    rows=e.rows()[0]
    print(e.columns)
    for name in ('Sigma','Tph','Tph_vis','Tirr','Height','Tph_X','Qx','W','Kirr'):
        f=e.field(name,0)[0]
        print(name, 'nan count', np.isnan(f).sum(), 'neg-signed', np.signbit(f[np.isnan(f)]).sum())
m=ref.model(q['args'], q['lambdas'])
for name in ('Sigma','Tph','Tph_vis','Tirr','Height','Tph_X','W','Kirr'):
    f=m.field(name)
    print('ref', name, 'nan count', np.isnan(f).sum(), 'neg-signed', np.signbit(f[np.isnan(f)]).sum())
for k,c in enumerate(e.columns):
    g=rows[:,k]; o=r['rows'][:,k]
    if np.isnan(g).any() or np.isnan(o).any():
        print(c, 'gpu nan', np.isnan(g).sum(), 'neg', np.signbit(g[np.isnan(g)]).sum(), '| ref nan', np.isnan(o).sum(), 'neg', np.signbit(o[np.isnan(o)]).sum())
