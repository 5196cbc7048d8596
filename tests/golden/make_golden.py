"""Generates the committed fixtures under tests/golden/ from the reference's shipped assets
(run here, where /root/reference exists; the GPU box only sees the generated files).

  cornell_box.npz, cornell_box_glass.npz, viking_room.npz, breakfast_room.npz
        flat scene arrays produced by the ORACLE-side loader oracle/gltf_ref.py from
        /root/reference/Assets/*.gltf  (what AssetImporter::ImportScene + PathTracer::SetScene keep)
  luts.npz
        the three energy-compensation tables /root/reference/Assets/LookupTables/*.bin (lossless)
  env_alias_kat.npz
        a small environment map + the alias table / pdf the oracle computes for it (known-answer vector)
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import gltf_ref, orc  # noqa: E402

A = "/root/reference/Assets/"
OUT = os.path.dirname(os.path.abspath(__file__))

for name, out in [("CornellBox", "cornell_box"), ("CornellBoxGlass", "cornell_box_glass"), ("VikingRoom", "viking_room"), ("BreakfastRoom", "breakfast_room")]:
    sc = gltf_ref.load_gltf(A + name + ".gltf")
    gltf_ref.save_scene_npz(os.path.join(OUT, out + ".npz"), sc)
    print(out, "meshes", len(sc["meshes"]), "tris", sum(len(i) // 3 for _, i in sc["meshes"]), "instances", len(sc["instances"]),
          "textures", [t.shape for t in sc["textures"]], os.path.getsize(os.path.join(OUT, out + ".npz")))

gltf_ref.save_luts_npz(os.path.join(OUT, "luts.npz"), *gltf_ref.load_luts_dir(A + "LookupTables"))
print("luts", os.path.getsize(os.path.join(OUT, "luts.npz")))

env = gltf_ref.synthetic_env(64, 32, seed=11, sun=800.0)
env2, alias, s = orc.build_env_alias(env)
np.savez_compressed(os.path.join(OUT, "env_alias_kat.npz"), env=env, env_pdf=env2, alias=alias.view(np.uint32).reshape(-1, 2), total=np.float32(s))
print("env kat", s)
