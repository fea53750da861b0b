"""Developer script: compare the CUDA library with the compiled reference on
small instances and print micro-benchmarks (run under gpurun)."""
import json, os, re, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gadfit_b200._capi import Api, product_api
from gadfit_b200.lm_solver import LMsolver
from gadfit_b200 import workloads as W
import ctypes as C

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ref = Api(os.path.join(ROOT, 'oracle/_ref/libgadfit_ref.so'), 'gadfit_ref_', product=False)
api = product_api()

for which, name in [(0, 'DFMA'), (1, 'DMMA884'), (2, 'DMMA+DFMA mixed')]:
    out = C.c_double()
    st = api.microbench(which, C.byref(out))
    print(f"microbench {name}: status {st} {out.value/1e12:.2f} TFLOP/s", flush=True)

fx = json.load(open(os.path.join(ROOT, 'tests/golden/c1_fixture.json')))
fix_d = fx['fix_d']; x1, y1, x2, y2 = (np.array(fx[k]) for k in ('x_data_1', 'y_data_1', 'x_data_2', 'y_data_2'))

def run_c1(api_, accel=-1.0, loss=0, limit=4):
    s = LMsolver('exponential', _api=api_)
    s.addDataset(x1, y1); s.addDataset(x2, y2)
    s.settings.iteration_limit = limit
    s.settings.acceleration_threshold = accel
    s.settings.loss = loss
    s.setPar(0, fix_d[0], True, 0); s.setPar(2, fix_d[1], True, 0)
    s.setPar(0, fix_d[4], True, 1); s.setPar(2, fix_d[5], True, 1)
    s.setPar(1, fix_d[3], True)
    s.fit(1.0)
    return s

def compare(tag, a, b, n_sets, n_pars):
    pa = np.array([[a.getParValue(k, s) for k in range(n_pars)] for s in range(n_sets)])
    pb = np.array([[b.getParValue(k, s) for k in range(n_pars)] for s in range(n_sets)])
    rel = np.max(np.abs(pa - pb) / np.maximum(np.abs(pb), 1e-300))
    ca, cb = a.chi2(), b.chi2()
    ta, tb = a.trace(), b.trace()
    same = [x['decision'] for x in ta] == [x['decision'] for x in tb]
    print(f"{tag}: max rel par diff {rel:.3e}  chi2 {ca!r} vs {cb!r} rel {abs(ca-cb)/abs(cb):.3e}  "
          f"iters {a.iterations()} vs {b.iterations()}  decisions equal: {same} ({len(ta)} vs {len(tb)} trials)", flush=True)
    if not same:
        for x, y in zip(ta, tb):
            print("    ", x['lambda_'], x['new_chi2'], x['decision'], '|', y['lambda_'], y['new_chi2'], y['decision'])

for kw in [dict(), dict(accel=0.9, limit=5), dict(loss=1, limit=5), dict(loss=2, limit=2), dict(limit=100)]:
    a = run_c1(api, **kw); b = run_c1(ref, **kw)
    compare(f"C1 {kw}", a, b, 2, 3)

for name, kw in [("C2", dict(n_sets=4, n_points=2000)), ("C3", dict(n_sets=8, n_points=512)),
                 ("C4", dict(n_sets=4, n_points=256)), ("C5", dict(n_sets=6, n_points=64)),
                 ("C3", dict(n_sets=40, n_points=1000)), ("C2", dict(n_sets=3, n_points=100000))]:
    w = W.WORKLOADS[name](**kw)
    sols = []
    for api_ in (api, ref):
        s = LMsolver(w.model, _api=api_)
        w.setup(s); s.settings.n_threads = 8
        t = time.time(); s.fit(w.lambda0); dt = time.time() - t
        sols.append(s)
        print(f"  {name} {kw} {'cuda' if api_ is api else 'ref '} fit {dt*1e3:.1f} ms timings {s.timings()}")
    compare(f"{name} {kw}", sols[0], sols[1], w.n_sets, w.n_pars)
