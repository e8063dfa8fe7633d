"""Dev tool: e2e rate of foundation_pt_trace_closest (pinned host rays in, pinned host hits out) for several chunk schedules of the host pipeline."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from foundation_b200 import pt, scenes  # noqa: E402

sc = scenes.fractal_terrain()
n = 1 << 26
tr = pt.PathTracer(sc.width, sc.height, background=sc.background)
bs = tr.load(sc)
rays = scenes.incoherent_rays(np.asarray(bs.scene_lo[:]), np.asarray(bs.scene_hi[:]), n, 4)
pin_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1)).pin_memory()
pin_hits = torch.empty(n * 16, dtype=torch.uint8).pin_memory()
tr.close()
for rep in range(2):
    for big, tail in ((22, 22), (22, 20), (22, 18), (21, 21), (21, 19), (23, 20), (20, 20), (24, 21)):
        os.environ["FOUNDATION_PT_E2E_CHUNK_LOG2"] = str(big); os.environ["FOUNDATION_PT_E2E_TAIL_LOG2"] = str(tail)
        tr = pt.PathTracer(sc.width, sc.height, background=sc.background)
        tr.load(sc)
        tr.trace_closest_raw(pin_rays.data_ptr(), n, pin_hits.data_ptr())
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3):
            tr.trace_closest_raw(pin_rays.data_ptr(), n, pin_hits.data_ptr())
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
        print(f"chunk 2^{big} tail 2^{tail}: {dt * 1e3:.2f} ms/step  {n / dt / 1e6:.1f} Mrays/s  (H2D {n * 32 / dt / 1e9:.1f} GB/s)", flush=True)
        tr.close()
