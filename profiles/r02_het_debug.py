"""Heterogeneous-volume parity, narrowed down: matched-seed agreement GPU vs oracle for variants of one cloud scene (which estimator term diverges?)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import util
import test_gpu_parity as tp

name, depth, pf, kw, vols = tp._het_cases()[0]
S = util.oracle_scene(name)
def run(label, depth=depth, vols=vols, **extra):
    cfgkw = dict(MaxDepth=depth, PhaseFunction=pf, Volumes=vols, **extra)
    ref, cnt = S.render(util.oracle_config(name, **cfgkw), 128, 96, 1, util.BASE_SEED)
    T = util.product_tracer(name, 128, 96, **cfgkw); T.path_trace(1, util.BASE_SEED); got = T.get_hdr(); c = T.counters()
    a, b = ref[..., :3].astype(np.float64), got[..., :3].astype(np.float64)
    close = np.all(np.abs(a - b) <= 1e-4 * np.maximum(np.abs(a), 1e-2), axis=-1)
    print(f"{label:34s} agree {close.mean():.5f}  events {cnt['medium_events']:6d}/{c['medium_events']:6d}  segments {cnt['segments']:6d}/{c['extend_rays']:6d}  "
          f"shadow {cnt['shadow_rays']:6d}/{c['shadow_rays']:6d}  mean {a.mean():.5f}/{b.mean():.5f}", flush=True)
run("full")
run("no NEE", EnableSkyMIS=0, EnableMeshMIS=0)
run("sky NEE only", EnableMeshMIS=0)
run("light NEE only", EnableSkyMIS=0)
run("depth 1", depth=1)
run("depth 2", depth=2)
run("depth 2, no NEE", depth=2, EnableSkyMIS=0, EnableMeshMIS=0)
absorb = [dict(vols[0], Color=(0.0, 0.0, 0.0))]
run("absorbing", vols=absorb)
run("absorbing, no NEE", vols=absorb, EnableSkyMIS=0, EnableMeshMIS=0)
hom = [dict({k: v for k, v in vols[0].items() if k != "Grid"}, CornerMin=vols[0]["Grid"]["corner_min"], CornerMax=vols[0]["Grid"]["corner_max"])]
run("homogeneous twin (Position/Scale)", vols=hom)
