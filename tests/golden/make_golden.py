"""Regenerates tests/golden/*: run from the repo root with `python tests/golden/make_golden.py`.
The reference (mos9527/Foundation) has no ray tracer and no fixtures for this path, so these vectors
come from this repo's CPU oracle; they pin the arithmetic contract against accidental change and are what the GPU
tests compare with on the box (where /root/reference does not exist anyway)."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from foundation_b200 import scenes  # noqa: E402
from oracle.pt_oracle import OracleScene  # noqa: E402
from tests.util import SMALL_SCENES, ray_mix  # noqa: E402

out = {"scenes": {}}
for name, make in SMALL_SCENES.items():
    sc = make(); o = OracleScene(sc)
    nodes, tris, order = o.blas(0)
    rays = ray_mix(sc, 512)
    h, i = o.trace_closest(rays)
    img = o.render(sc.width, sc.height, 7, 0, 1, 3, background=sc.background)
    out["scenes"][name] = {"blas0_nodes_sha256": hashlib.sha256(nodes.tobytes()).hexdigest(), "hits_sha256": hashlib.sha256(h.tobytes() + i.tobytes()).hexdigest(),
                           "image_sha256": hashlib.sha256(img.tobytes()).hexdigest(), "num_nodes": int(len(nodes)), "triangles": int(sc.num_triangles)}
here = os.path.dirname(os.path.abspath(__file__))
json.dump(out, open(os.path.join(here, "golden.json"), "w"), indent=1)
sc = SMALL_SCENES["cornell"](); o = OracleScene(sc)
lo, hi = scenes.scene_bounds(sc)
rays = np.concatenate([scenes.incoherent_rays(lo, hi, 192, 21), scenes.stress_rays(sc, 64, 22)])
h, _ = o.trace_closest(rays)
np.savez_compressed(os.path.join(here, "cornell_hits.npz"), rays=rays.view(np.float32).reshape(-1, 8), prim=h["prim"], t_bits=h["t"].view(np.uint32))
print(json.dumps(out, indent=1))
