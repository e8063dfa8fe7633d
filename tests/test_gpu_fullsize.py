"""BASELINE.json's five configs at their FULL size, under the driver's eyes (`-m gpu`): round 1 only ran miniatures here and kept the
full-size soak in a script.  Per config: device hits vs the CPU oracle's BVH on 2^20 rays (bit-exact ids, t, u, v), device BVH vs the
device's exhaustive search (no BVH), and a frame at the config's resolution vs the oracle's scalar path loop (bit-identical, RMSE gate
1e-3 written out).  Config 5 (3840x2160, sharded over 8 GPUs) is checked through its size-independent property: the 8 interleaved-tile
partitions, rendered one after the other, sum to the unpartitioned frame exactly.
Resolution conventions: the reference's target is 1920x1080 (mos9527/Foundation src/Renderer/Renderer.cpp:40-41)."""
import numpy as np
import pytest

from foundation_b200 import pt, scenes
from oracle.pt_oracle import OracleScene
from tests.util import assert_hits_equal, rmse

pytestmark = pytest.mark.gpu

FULL = {
    # name: (scene factory, rays vs oracle, rays vs exhaustive search, spp, bounces)
    "config1_cornell_512": (lambda: scenes.cornell_box(512, 512), 1 << 20, 1 << 20, 1, 4),
    "config2_sphere_field_1M": (scenes.sphere_field, 1 << 20, 1 << 18, 2, 8),
    "config3_terrain_10M": (scenes.fractal_terrain, 1 << 20, 1 << 16, 2, 8),
    "config4_instanced_100M": (scenes.instanced_patches, 1 << 20, 1 << 12, 2, 8),
}


@pytest.fixture(scope="module", params=list(FULL))
def full(request, gpu):
    make, n_orc, n_brute, spp, bounces = FULL[request.param]
    sc = make()
    tr = pt.PathTracer(sc.width, sc.height, seed=5, background=sc.background)
    bs = tr.load(sc)
    yield request.param, sc, tr, OracleScene(sc), bs, n_orc, n_brute, spp, bounces
    tr.close()


def test_fullsize_triangle_counts(full):
    name, sc, tr, orc, bs, *_ = full
    want = {"config1_cornell_512": 32, "config2_sphere_field_1M": 1024004, "config3_terrain_10M": 9999394, "config4_instanced_100M": 100820002}[name]
    assert int(bs.effective_triangles) == want, (name, int(bs.effective_triangles))


def test_fullsize_hits_match_oracle_and_exhaustive_search(full):
    name, sc, tr, orc, bs, n_orc, n_brute, *_ = full
    lo, hi = np.asarray(bs.scene_lo[:]), np.asarray(bs.scene_hi[:])
    rays = np.concatenate([scenes.incoherent_rays(lo, hi, n_orc // 2, 4), scenes.camera_rays(sc, n_orc // 2, 6)])
    gh, gi = tr.trace_closest(rays)
    oh, oi = orc.trace_closest(rays)
    assert_hits_equal(gh, gi, oh, oi, f"{name}: device vs oracle BVH")
    assert gh.tobytes() == oh.tobytes(), f"{name}: t/u/v not bit-identical"
    assert np.array_equal(tr.trace_any(rays[: 1 << 18]), orc.trace_any(rays[: 1 << 18])), f"{name}: occlusion differs"
    tr.rays_upload(rays[:n_brute]); tr.rays_trace_brute()
    bh, bi = tr.rays_download_hits()
    assert_hits_equal(gh[:n_brute], gi[:n_brute], bh, bi, f"{name}: device BVH vs device exhaustive search")
    print(f"{name}: hit fraction {float((gh['prim'] != 0xFFFFFFFF).mean()):.3f}")


def test_fullsize_frame_matches_oracle(full):
    name, sc, tr, orc, bs, _, _, spp, bounces = full
    tr.render(0, spp, bounces)
    g = tr.read_accum()
    o = orc.render(sc.width, sc.height, 5, 0, spp, bounces, background=sc.background)
    e = rmse(g[..., :3] / spp, o[..., :3] / spp)
    differing = int((g != o).any(axis=-1).sum())
    print(f"{name}: {sc.width}x{sc.height} {spp} spp {bounces} bounces rmse={e:.3e} differing_pixels={differing}")
    assert e <= 1e-3, f"{name}: image RMSE {e} > 1e-3 (north_star gate on linear radiance)"
    assert differing == 0, f"{name}: {differing} pixels not bit-identical"


def test_config5_4k_eight_tile_partitions_sum_to_the_whole_frame(gpu):
    """Config 5: 3840x2160 progressive render of the 10 M-triangle terrain sharded over 8 ranks by interleaved 32x32 tiles.  On one GPU
    the 8 partitions are rendered one after the other (two progressive batches each); their sum must equal the unpartitioned frame bit
    for bit, every pixel must be owned exactly once, and the sample counter must be uniform."""
    W, H, bounces = 3840, 2160, 8
    sc = scenes.fractal_terrain(width=W, height=H)
    tr = pt.PathTracer(W, H, seed=1, background=sc.background)
    tr.load(sc)
    tr.render(0, 1, bounces); tr.render(1, 1, bounces)
    whole = tr.read_accum().copy()
    total = np.zeros_like(whole)
    for rank in range(8):
        tr.partition_set(rank, 8, 32)
        tr.render(0, 1, bounces); tr.render(1, 1, bounces)
        part = tr.read_accum()
        assert np.all((part[..., 3] == 0) | (total[..., 3] == 0)), "tiles overlap"
        total += part
    assert np.all(total[..., 3] == 2.0)
    assert np.array_equal(total, whole), f"4K: {int((total != whole).any(-1).sum())} pixels differ between the 8-way partition and the whole frame"
    tr.close()


def test_fullsize_two_waves_in_flight_are_bit_identical_to_one_wave_at_a_time(gpu, monkeypatch):
    """The wave logic at sizes the miniatures cannot reach: (a) a 1080p call that fits ONE wave but is large enough to be cut into two halves on two streams (5 samples =
    10.4 M slots), (b) a call of several full waves plus a short one (37 samples at 16 per wave), (c) one rank's share of an 8-GPU frame (64 samples of 1/8 of the pixels —
    one wave by size, two by the split rule).  Each must equal, bit for bit, the same samples rendered one wave at a time (FOUNDATION_PT_DUAL_WAVE=0), ray counters included."""
    sc = scenes.sphere_field()
    frames, rays = {}, {}
    for dual in ("1", "0"):
        monkeypatch.setenv("FOUNDATION_PT_DUAL_WAVE", dual)
        tr = pt.PathTracer(sc.width, sc.height, seed=9, background=sc.background)
        tr.load(sc)
        out, cnt = [], []
        for s0, ns in ((0, 5), (0, 37)):
            tr.render(s0, ns, 8)
            st = tr.stats()
            out.append(tr.read_accum().copy()); cnt.append((st.rays_extend, st.rays_shadow))
        tr.partition_set(0, 8, 32)
        tr.render(0, 64, 8)
        st = tr.stats()
        out.append(tr.read_accum().copy()); cnt.append((st.rays_extend, st.rays_shadow))
        tr.close()
        frames[dual], rays[dual] = out, cnt
    for k, what in enumerate(("5 samples (one wave cut in two)", "37 samples (three waves)", "rank 0 of 8, 64 samples")):
        assert frames["1"][k].tobytes() == frames["0"][k].tobytes(), what
        assert rays["1"][k] == rays["0"][k] and rays["1"][k][0] > 0, what
