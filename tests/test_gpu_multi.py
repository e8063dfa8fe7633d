"""Stage C1 inside the C ABI: the tile-partitioned frame gathered into one member's accumulation buffer, checked against the
unpartitioned single-context frame BIT FOR BIT.  Three host shapes: several partitions on one device (`gather_local`: the pack / scatter
kernels of the NCCL form with a device copy in place of send / recv — runs on the driver's one-GPU box), one process driving several GPUs
(`foundation_pt_group_*`, ncclCommInitAll), and one process per GPU (`comm_unique_id` / `comm_init` / `gather`).  The last two need at
least two GPUs and are skipped on a one-GPU box (run with `gpurun --gpus 2`).
The reference is single-device (mos9527/Foundation src/Editor/Editor.cpp:18); the slot is Renderer::Draw (src/Renderer/Renderer.cpp:367-401)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from foundation_b200 import pt, scenes
from tests.util import SMALL_SCENES

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def gpu_count() -> int:
    import torch
    return torch.cuda.device_count()


def single_frame(sc, seed, spp, bounces, device=0):
    with pt.PathTracer(sc.width, sc.height, device=device, seed=seed, background=sc.background) as t:
        t.load(sc)
        t.render(0, 1, bounces); t.render(1, spp - 1, bounces)
        return t.read_accum().copy()


@pytest.mark.parametrize("shape", [(128, 128, 16, 3), (100, 70, 16, 4), (160, 90, 32, 8), (37, 29, 8, 5)])
def test_gather_local_partitions_equal_the_whole_frame(gpu, shape):
    """Ragged frames (width / height not multiples of the tile), more ranks than tile columns, every rank as the root."""
    W, H, tile, count = shape
    sc = scenes.cornell_box(W, H)
    ref = single_frame(sc, 11, 3, 3)
    parts = []
    for r in range(count):
        t = pt.PathTracer(W, H, seed=11, background=sc.background); t.load(sc)
        t.partition_set(r, count, tile)
        t.render(0, 1, 3); t.render(1, 2, 3)
        parts.append(t)
    for root in (0, count - 1):
        own = parts[root].read_accum().copy()
        for r in range(count):
            if r != root:
                parts[root].gather_local(parts[r])
        assert np.array_equal(parts[root].read_accum(), ref), f"{shape}: gathered frame differs (root {root})"
        parts[root].write_accum(own)                      # restore: the next root gathers from unmodified partitions
    for t in parts:
        t.close()


def test_group_of_one_equals_a_plain_context(gpu):
    sc = SMALL_SCENES["spheres"]()
    ref = single_frame(sc, 5, 4, 4)
    with pt.Group([0], sc.width, sc.height, seed=5, background=sc.background) as g:
        g.load(sc)
        g.render(0, 1, 4); g.render(1, 3, 4)
        assert np.array_equal(g.read_accum(), ref)


@pytest.mark.parametrize("mode", ["nccl", "direct"])
def test_group_over_all_gpus_is_bit_identical_to_one_gpu(gpu, mode):
    n = gpu_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    for name in ("terrain", "instanced"):
        sc = SMALL_SCENES[name]()
        ref = single_frame(sc, 9, 4, 4)
        with pt.Group(list(range(n)), sc.width, sc.height, seed=9, background=sc.background, tile=16, comm_flags=pt.COMM_DIRECT if mode == "direct" else 0) as g:
            g.load(sc)
            g.render(0, 1, 4); g.render(1, 3, 4)
            got = g.read_accum()
            assert np.array_equal(got, ref), f"{name} {mode}: {int((got != ref).any(-1).sum())} pixels differ over {n} GPUs"
            assert g.members[0].stats().gather_ms >= 0.0


@pytest.mark.parametrize("mode", ["nccl", "direct"])
def test_one_process_per_gpu_comm_init_and_gather(gpu, mode, tmp_path):
    n = gpu_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    n = min(n, 4)
    out = str(tmp_path / "frame.npy"); idf = str(tmp_path / "comm.id")
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "multi_worker.py"), str(r), str(n), mode, idf, out], cwd=ROOT) for r in range(n)]
    for p in procs:
        assert p.wait(timeout=600) == 0
    sc = SMALL_SCENES["terrain"]()
    ref = single_frame(sc, 13, 3, 4)
    assert np.array_equal(np.load(out), ref), f"{mode}: gathered frame of {n} processes differs from the single-GPU frame"


def test_cpp_editor_on_all_gpus_matches_one_gpu(gpu, tmp_path):
    """The C++ host (`Renderer(devices, ...)` / `Draw()`, headless Editor with --gpus N): its gathered frame equals its single-device frame."""
    n = gpu_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    rdir = os.path.join(ROOT, "foundation_b200", "renderer")
    subprocess.run(["make", "-C", rdir], check=True, capture_output=True)
    sc = SMALL_SCENES["spheres"]()
    path = str(tmp_path / "s.fpts"); scenes.save_scene(sc, path)
    frames = {}
    for label, extra in (("one", []), ("nccl", ["--gpus", str(n)]), ("direct", ["--gpus", str(n), "--direct"])):
        raw = str(tmp_path / f"{label}.raw")
        out = subprocess.run([os.path.join(rdir, "foundation_editor")] + extra + [path, "2", "2", "4", raw], check=True, capture_output=True, text=True).stdout
        assert f"gpus={n if extra else 1} " in out and "Memory Used: 0 bytes" in out, out
        frames[label] = np.fromfile(raw, np.float32)
    assert np.array_equal(frames["nccl"], frames["one"]) and np.array_equal(frames["direct"], frames["one"])
