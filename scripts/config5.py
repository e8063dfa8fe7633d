"""BASELINE.json config 5: 3840x2160 progressive render of the 10 M-triangle terrain, 1024 spp in 64-spp batches, tile/sample sharding over all
ranks with the C ABI's gather of the owned tiles into rank 0 (`foundation_pt_gather`) after every batch and, separately, only at the end.  Run under torchrun, one rank per GPU."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch  # noqa: E402
from foundation_b200 import distributed as fdist  # noqa: E402
from foundation_b200 import pt, scenes  # noqa: E402

rank, world, local = fdist.init_process_group()
torch.cuda.set_device(local)
dist = torch.distributed if world > 1 else None
W, H, SPP, BATCH, BOUNCES = 3840, 2160, int(os.environ.get("SPP", 1024)), 64, 8
sc = scenes.fractal_terrain(width=W, height=H)
tr = pt.PathTracer(W, H, device=local, seed=1, background=sc.background)
tr.load(sc)
dr = fdist.DistributedRenderer(tr, rank, world, direct=os.environ.get("DIRECT", "1") == "1")


def barrier():
    if dist:
        dist.barrier()
    torch.cuda.synchronize()


res = {}
for mode in ("gather_every_batch", "gather_at_end"):
    dr.render(0, BATCH, BOUNCES, gather=False)     # warm-up + clears the accumulation
    barrier(); t0 = time.perf_counter()
    for s0 in range(0, SPP, BATCH):
        last = s0 + BATCH >= SPP
        dr.render(s0, BATCH, BOUNCES, gather=(mode == "gather_every_batch") or last)
    barrier(); dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local}")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[mode] = {"seconds": float(t.item()), "spp_per_s": SPP / float(t.item())}
    if rank == 0:
        f = tr.read_accum()
        res[mode]["frame_mean_radiance"] = float(f[..., :3].mean() / SPP); res[mode]["min_samples_per_pixel"] = float(f[..., 3].min())
if rank == 0:
    print(json.dumps({"config": "5: 3840x2160, 1024 spp, 8 bounces, terrain 9,999,394 tris", "n_gpus": world, "spp": SPP, "batch": BATCH, **res}))
barrier()
if dist:
    dist.destroy_process_group()
