"""Volume walks on random rays, GPU (b200pt_volume_walks) vs oracle (orc_volume_walks): where do they part?  Saves the mismatching rays for offline tracing."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import util
import test_gpu_parity as tp
from oracle import orc

out = {}
for case in (0, 1, 2):
    name, depth, pf, kw, vols = tp._het_cases()[case]
    cfg = util.oracle_config(name, Volumes=vols)
    T = util.product_tracer(name, 32, 32, Volumes=vols)
    rs = np.random.RandomState(5 + case)
    n = 60000
    gv = [v for v in vols if v.get("Grid") is not None][0]
    lo = np.array(gv["Position"]) + np.array(gv["Grid"]["corner_min"]) * np.array(gv["Scale"]); hi = np.array(gv["Position"]) + np.array(gv["Grid"]["corner_max"]) * np.array(gv["Scale"])
    ctr, rad = (lo + hi) / 2, np.linalg.norm(hi - lo) / 2
    u = rs.randn(n, 3); u /= np.linalg.norm(u, axis=1, keepdims=True)
    org = (ctr + u * rad * rs.uniform(0.2, 1.6, (n, 1))).astype(np.float32)
    tgt = lo + rs.rand(n, 3) * (hi - lo)
    d = tgt - org; d /= np.linalg.norm(d, axis=1, keepdims=True); d = d.astype(np.float32)
    seeds = rs.randint(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    for rd in (0.0, 2.0):
        a = orc.volume_walks(cfg, org, d, seeds, rd); b = T.volume_walks(org, d, seeds, rd)
        okT = (a[0] == b[0]) | (np.abs(a[0] - b[0]) < 1e-5); okR0 = a[3][:, 0] == b[3][:, 0]
        okS = (np.abs(a[1] - b[1]) <= 1e-4 * np.maximum(np.abs(a[1]), 1.0)) & (a[2] == b[2]); okR1 = a[3][:, 1] == b[3][:, 1]
        print(f"case {case} depth {rd}: T equal {okT.mean():.5f} (rng state after: {okR0.mean():.5f})   scatter equal {okS.mean():.5f} (rng after: {okR1.mean():.5f})   "
              f"mean T {a[0].mean():.4f}/{b[0].mean():.4f}  scatter frac {(a[1] >= 0).mean():.4f}/{(b[1] >= 0).mean():.4f}", flush=True)
        if rd == 0.0:
            bad = np.where(~okR0 | ~okR1)[0][:200]
            out[f"c{case}_org"] = org[bad]; out[f"c{case}_dir"] = d[bad]; out[f"c{case}_seed"] = seeds[bad]
            out[f"c{case}_oT"] = a[0][bad]; out[f"c{case}_gT"] = b[0][bad]; out[f"c{case}_oS"] = a[1][bad]; out[f"c{case}_gS"] = b[1][bad]
            out[f"c{case}_orng"] = a[3][bad]; out[f"c{case}_grng"] = b[3][bad]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez(os.path.join(ROOT, "gpurun_out", "het_walk_mismatch.npz"), **out)
