#!/usr/bin/env python
"""Prints the GPU-vs-oracle parity numbers quoted in DESIGN.md (run on the GPU box)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util

def both(name, W, H, frames, seed, **kw):
    S = util.oracle_scene(name); ref, _ = S.render(util.oracle_config(name, **kw), W, H, frames, seed)
    T = util.product_tracer(name, W, H, **kw); T.path_trace(frames, seed); return ref, T.get_hdr()

for name, depth, (W, H) in (("cornell_box", 8, (384, 216)), ("cornell_box_glass", 16, (256, 256)), ("viking_room", 8, (256, 256)), ("breakfast_room", 8, (256, 144))):
    ref, got = both(name, W, H, 1, util.BASE_SEED, MaxDepth=depth)
    a, b = ref[..., :3].astype(np.float64), got[..., :3].astype(np.float64)
    close = np.all(np.abs(a - b) <= 1e-4 * np.maximum(np.abs(a), 1e-2), axis=-1).mean()
    exact = np.all(a == b, axis=-1).mean()
    ref64, got64 = both(name, W // 2, H // 2, 64, util.BASE_SEED, MaxDepth=depth)
    print(f"{name:18s} 1spp: {close*100:.3f}% px within 1e-4, {exact*100:.2f}% bit-identical | 64 frames: rel L2 = {util.rel_l2(got64[..., :3], ref64[..., :3]):.2e}")
