"""Repeated scene builds in one process: separates first-use costs (module load, pool growth) from the steady-state build."""
import sys
import time

sys.path.insert(0, ".")
from foundation_b200 import pt, scenes  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "terrain"
sc = scenes.by_name(name)
for i in range(4):
    tr = pt.PathTracer(sc.width, sc.height, background=sc.background)
    t0 = time.perf_counter()
    tr.materials_set(sc.materials)
    for m in sc.meshes:
        tr.mesh_create(m.positions, m.indices, m.material_ids)
    if sc.instances is not None:
        tr.instances_set(sc.instances)
    t1 = time.perf_counter()
    bs = tr.scene_commit()
    t2 = time.perf_counter()
    print(f"run {i}: upload {1e3 * (t1 - t0):.1f} ms  commit wall {1e3 * (t2 - t1):.1f} ms  build_ms(events) {bs.build_ms:.2f}  sort_ms {bs.sort_ms:.2f}  "
          f"{bs.num_triangles / bs.build_ms / 1e3:.1f} Mtris/s  launches {tr.stats().kernel_launches}", flush=True)
    tr.close()
