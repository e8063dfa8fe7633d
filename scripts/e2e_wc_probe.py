"""Dev tool: does write-combined pinned host memory (cudaHostAllocWriteCombined) for the RAY buffer change the e2e rate of foundation_pt_trace_closest?"""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from foundation_b200 import pt, scenes  # noqa: E402

rt = C.CDLL("/usr/local/cuda/lib64/libcudart.so.12")
sc = scenes.fractal_terrain()
n = 1 << 26
tr = pt.PathTracer(sc.width, sc.height, background=sc.background)
bs = tr.load(sc)
rays = scenes.incoherent_rays(np.asarray(bs.scene_lo[:]), np.asarray(bs.scene_hi[:]), n, 4)
raw = rays.view(np.uint8).reshape(-1)
pin_hits = torch.empty(n * 16, dtype=torch.uint8).pin_memory()
pin_rays = torch.from_numpy(raw).pin_memory()


def host_alloc(nbytes, flags):
    p = C.c_void_p()
    rc = rt.cudaHostAlloc(C.byref(p), C.c_size_t(nbytes), C.c_uint(flags))
    assert rc == 0, rc
    return p.value


for label, flags in (("torch pin_memory", None), ("cudaHostAlloc default", 0), ("cudaHostAlloc write-combined", 4), ("torch pin_memory", None)):
    if flags is None:
        ptr = pin_rays.data_ptr()
    else:
        ptr = host_alloc(raw.nbytes, flags)
        C.memmove(ptr, raw.ctypes.data, raw.nbytes)
    tr.trace_closest_raw(ptr, n, pin_hits.data_ptr())
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        tr.trace_closest_raw(ptr, n, pin_hits.data_ptr())
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    h = pin_hits.numpy().view(scenes.HIT_DTYPE)
    print(f"{label}: {dt * 1e3:.2f} ms/step  {n / dt / 1e6:.1f} Mrays/s  (H2D {n * 32 / dt / 1e9:.1f} GB/s)  hits {int((h['prim'] != 0xFFFFFFFF).sum())}", flush=True)
    if flags is not None:
        rt.cudaFreeHost(C.c_void_p(ptr))
