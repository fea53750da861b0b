"""Developer script (run under gpurun): direct 2-D integral sections, device
vs oracle, trial by trial."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import backends, scenarios

d = {k: np.array(v) for k, v in json.load(open(os.path.join(ROOT, "tests/golden/c1_fixture.json"))).items()
     if k != "_source"}
for sc in scenarios.DIRECT_SCENARIOS:
    g = scenarios.run_direct(sc, backends.product_solver(), d)
    o = scenarios.run_direct(sc, backends.oracle_solver(1), d)
    pg, po = np.array(g["pars"][0]), np.array(o["pars"][0])
    print(sc["id"], "chi2 gpu %r oracle %r kat %r" % (g["chi2"], o["chi2"], sc["kat"][0]))
    print("   par rel diff gpu-oracle %.2e  gpu-kat %.2e" % (np.max(np.abs(pg - po) / np.abs(po)),
          np.max(np.abs(pg - np.array(sc["kat"][1])) / np.abs(sc["kat"][1]))))
    for tg, to in zip(g["trace"], o["trace"]):
        print("   trial: old %.3e new %.3e | rel old %.2e new %.2e | dec %d %d" % (
            to["old_chi2"], to["new_chi2"], abs(tg["old_chi2"] - to["old_chi2"]) / to["old_chi2"],
            abs(tg["new_chi2"] - to["new_chi2"]) / to["new_chi2"], tg["decision"], to["decision"]))
