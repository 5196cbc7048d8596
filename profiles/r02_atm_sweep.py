"""Atmosphere parity: accumulated rel L2 (GPU vs oracle, same seeds) against the sample count -- the residual is the noise of the few paths on which
the two fp32 implementations branch differently, each a firefly under the 2e5-bright sun disk, so it must fall like 1 / sqrt(spp).
usage (GPU box): python profiles/r02_atm_sweep.py > gpurun_out/r02_atm_sweep.txt"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import util

CASES = [("viking_room", 6, dict(EnableAtmosphere=1, SkyRotationAltitude=-35.0, SkyRotationAzimuth=200.0, SunColor=(1.0, 0.8, 0.6))),
         ("cornell_box", 8, dict(EnableAtmosphere=1, SkyRotationAltitude=-30.0))]
for name, depth, kw in CASES:
    S = util.oracle_scene(name)
    for frames in (16, 64, 256, 1024, 4096):
        ref, cnt = S.render(util.oracle_config(name, MaxDepth=depth, **kw), 48, 36, frames, util.BASE_SEED)
        T = util.product_tracer(name, 48, 36, MaxDepth=depth, **kw)
        T.path_trace(frames, util.BASE_SEED)
        got = T.get_hdr()
        print(f"{name:12s} spp {frames:5d}  rel L2 {util.rel_l2(got[..., :3], ref[..., :3]):.3e}  rel L2 * sqrt(spp) {util.rel_l2(got[..., :3], ref[..., :3]) * frames ** 0.5:.3e}", flush=True)
