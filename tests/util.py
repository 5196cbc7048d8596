"""Shared helpers of the test-suite: build the SAME scene for the CPU oracle and for the product (through the C-ABI)."""
import os
import sys
import functools
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import orc, gltf_ref  # noqa: E402  (test infrastructure)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_ASSETS = "/root/reference/Assets"
HAVE_REF = os.path.isdir(REF_ASSETS)
BASE_SEED = 0x1234ABCD


@functools.lru_cache(maxsize=None)
def luts():
    return gltf_ref.load_luts_npz(os.path.join(GOLDEN, "luts.npz"))


@functools.lru_cache(maxsize=None)
def scene_dict(name):
    return gltf_ref.load_scene_npz(os.path.join(GOLDEN, name + ".npz"))


@functools.lru_cache(maxsize=None)
def env_small(w=512, h=256, seed=3):
    """synthetic HDR env (raw rgba, alpha 1) + oracle-built pdf/alias"""
    raw = gltf_ref.synthetic_env(w, h, seed)
    env_pdf, alias, total = orc.build_env_alias(raw)
    return raw, env_pdf, alias


def oracle_scene(name, env=None):
    raw, env_pdf, alias = env if env is not None else env_small()
    return orc.Scene(scene_dict(name), env_pdf, alias, luts())


def camera(name):
    sc = scene_dict(name)
    return orc.camera_from_view(sc["camera_view"], sc["aspect"])


def oracle_config(name, **kw):
    vi, pi = camera(name)
    return orc.default_config(ViewInverse=vi, ProjectionInverse=pi, **kw)


_ORC2PT = {"SampleCount": "SamplesPerFrame", "EnvironmentIntensity": "SkyIntensity", "ScreenSplitCount": "ScreenChunkCount"}


def product_tracer(name, W, H, env=None, device=0, **cfg_kw):
    """PathTracer handle (C-ABI) with the same scene / env / LUTs / camera / config as oracle_scene + oracle_config."""
    import vpt_b200 as pt
    raw, env_pdf, alias = env if env is not None else env_small()
    t = pt.PathTracer(device)
    t.set_scene(scene_dict(name))
    t.set_env_map(raw)
    t.set_luts(*luts())
    cfg = pt.default_config()
    for k, v in cfg_kw.items():
        if k == "Volumes":                       # homogeneous AABB volumes: list of dicts (oracle.orc.VOLUME_DEFAULTS keys)
            for vol in v: t.add_volume(**vol)
        elif k == "PhaseFunction":
            t.set_phase_function(v)
        else:
            setattr(cfg, _ORC2PT.get(k, k), v)
    t.set_config(cfg)
    t.resize(W, H)
    return t


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.sqrt(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-30)))
