"""Config 4 with moving instances: time of a TLAS-only commit (BLAS reuse) vs the first full commit."""
import sys
import numpy as np
sys.path.insert(0, ".")
from foundation_b200 import pt, scenes  # noqa: E402

sc = scenes.instanced_patches()
tr = pt.PathTracer(sc.width, sc.height, background=sc.background)
bs = tr.load(sc)
print(f"full commit: {bs.build_ms:.2f} ms, launches {tr.stats().kernel_launches}, instances {bs.num_instances}, effective tris {bs.effective_triangles}")
inst = sc.instances.copy()
for it in range(4):
    inst["transform"][:-1, 11] += 0.01
    tr.instances_set(inst)
    bs = tr.scene_commit()
    print(f"TLAS-only commit {it}: {bs.build_ms:.2f} ms, launches {tr.stats().kernel_launches}")
tr.render(0, 1, 4)
print("render after move ok:", tr.stats().last_ms, "ms")
