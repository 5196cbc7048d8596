"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck / initcheck): every kernel family once, tiny sizes.
usage: compute-sanitizer --tool memcheck python profiles/sanitize_small.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import util

FOG = dict(CornerMin=(-4.0, -4.0, -10.0), CornerMax=(4.5, 3.0, -1.0), Density=0.35, Color=(0.8, 0.7, 0.6), Anisotropy=0.4, Alpha=1.5, DropletSize=14.0)
def run(name, W, H, frames, depth, **kw):
    T = util.product_tracer(name, W, H, MaxDepth=depth, **kw)
    T.path_trace(frames, util.BASE_SEED)
    img = T.get_hdr(); assert np.isfinite(img).all()
    T.post_process(); T.get_ldr()
    T.post_process_rows(8, min(H, 24)); T.get_ldr_rows(8, min(H, 24))
    print(name, kw.get("Volumes") is not None, os.environ.get("B200PT_FUSE"), os.environ.get("B200PT_TRAV"), "ok", float(img[..., :3].mean()))

run("cornell_box", 64, 36, 2, 6)                                   # BVH in shared memory (TMA), flat traversal, uniform class
os.environ["B200PT_FLAT_MAX"] = "0"; run("cornell_box", 48, 32, 1, 4); del os.environ["B200PT_FLAT_MAX"]      # BVH2 stack traversal in shared memory
os.environ["B200PT_FUSE"] = "2"; run("cornell_box", 48, 32, 2, 5); del os.environ["B200PT_FUSE"]              # fused bounce kernel
run("cornell_box", 48, 32, 1, 5, Volumes=[FOG])                     # k_volume_decide / k_shade_volume
run("cornell_box", 48, 32, 2, 5, EnableAtmosphere=1, SkyRotationAltitude=-30.0)                       # atmosphere: sun NEE, delta tracking, transmittance walks
run("cornell_box", 40, 30, 1, 5, EnableAtmosphere=1, SkyRotationAltitude=-20.0, Volumes=[FOG])        # atmosphere + homogeneous volume
def _het():
    from oracle import orc
    z, y, x = np.mgrid[0:20, 0:34, 0:36].astype(np.float32)
    d = np.exp(-(((x - 18) / 9) ** 2 + ((y - 17) / 9) ** 2 + ((z - 10) / 6) ** 2)).astype(np.float32) * 2.0
    return dict(Position=(0.2, -0.2, -5.5), Scale=(3.0, 3.0, 3.0), Density=4.0, Grid=orc.prepare_density_grid(d, index_min=(-18, -17, -10), temperature=d * 100.0))
run("cornell_box", 40, 30, 1, 5, Volumes=[FOG, _het()])                                              # grid volume: delta tracking, ratio-tracked NEE transmittance in k_connect<.., true>
run("cornell_box", 40, 30, 1, 5, EnableAtmosphere=1, SkyRotationAltitude=-20.0, Volumes=[_het()])    # ... under the atmosphere
run("cornell_box_glass", 48, 48, 2, 8)                              # dynamic-fetch BVH2 kernels, class queues (glass / diffuse)
os.environ["B200PT_WIDE"] = "1"; run("viking_room", 48, 48, 1, 4); del os.environ["B200PT_WIDE"]              # BVH4 + textures
os.environ["B200PT_SORT"] = "1"; os.environ["B200PT_TOP_KB"] = "16"; os.environ["B200PT_WIDE"] = "1"
run("viking_room", 32, 32, 1, 3)                                    # ray sort + treelet staging (opt-in paths)
print("sanitize_small done")
