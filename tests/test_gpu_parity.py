"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded
inputs.  Gates (north_star): hit IDs bit-exact with ties by primitive index, t within 2 ulp (we expect 0),
image RMSE <= 1e-3 on linear radiance (we expect bit-identical).  Marked gpu: run on the B200 box."""
import numpy as np
import pytest

from foundation_b200 import pt, scenes
from oracle.pt_oracle import OracleScene
from tests.util import SMALL_SCENES, assert_hits_equal, multi_mesh_scene, ray_mix, rmse

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=list(SMALL_SCENES))
def pair(request, gpu):
    sc = SMALL_SCENES[request.param]()
    tr = pt.PathTracer(sc.width, sc.height, seed=1, background=sc.background)
    tr.load(sc)
    yield request.param, sc, tr, OracleScene(sc)
    tr.close()


def test_build_is_byte_identical_to_oracle(pair):
    """A1-A6: device radix sort + Karras emit + refit + collapse produce the oracle's BVH8 byte for byte."""
    name, sc, tr, orc = pair
    for mid in range(len(sc.meshes)):
        gn, gt, go = tr.blas_download(mid)
        on, ot, oo = orc.blas(mid)
        assert np.array_equal(go, oo), f"{name}: Morton order differs (radix sort / keys)"
        assert len(gn) == len(on) and gn.tobytes() == on.tobytes(), f"{name}: BVH8 nodes differ"
        assert gt.tobytes() == ot.tobytes(), f"{name}: leaf-ordered triangles differ"
    if sc.instances is not None:
        gn, go = tr.tlas_download()
        on, oo, _ = orc.tlas()
        assert np.array_equal(go, oo) and gn.tobytes() == on.tobytes(), f"{name}: TLAS differs"


def test_closest_hit_matches_oracle_bvh_and_brute_force(pair):
    name, sc, tr, orc = pair
    rays = ray_mix(sc)
    gh, gi = tr.trace_closest(rays)
    oh, oi = orc.trace_closest(rays)
    assert_hits_equal(gh, gi, oh, oi, f"{name} vs oracle BVH")
    assert gh.tobytes() == oh.tobytes(), f"{name}: t/u/v not bit-identical"
    n = 512 if sc.effective_triangles > 200000 else 4096
    bh, bi = orc.trace_closest(rays[:n], brute=True)
    assert_hits_equal(gh[:n], gi[:n], bh, bi, f"{name} vs CPU brute force")
    # device-side exhaustive kernel on the whole set
    tr.rays_upload(rays)
    tr.rays_trace_brute()
    dh, di = tr.rays_download_hits()
    assert_hits_equal(gh, gi, dh, di, f"{name} vs device brute force")


def test_any_hit_matches_oracle(pair):
    name, sc, tr, orc = pair
    rays = ray_mix(sc)
    assert np.array_equal(tr.trace_any(rays), orc.trace_any(rays)), f"{name}: occlusion differs"


def test_device_resident_path_equals_host_path(pair):
    name, sc, tr, orc = pair
    rays = ray_mix(sc, 2048)
    gh, gi = tr.trace_closest(rays)
    tr.rays_upload(rays)
    tr.rays_trace_closest()
    dh, di = tr.rays_download_hits()
    assert gh.tobytes() == dh.tobytes() and np.array_equal(gi, di)
    st = tr.stats()
    assert st.kernel_launches >= 1 and st.last_ms > 0


@pytest.mark.parametrize("flags", [0, pt.FLAG_MATERIAL_SORT, pt.FLAG_SOBOL_JITTER, pt.FLAG_SOBOL_PATH, pt.FLAG_SOBOL_JITTER | pt.FLAG_SOBOL_PATH])
def test_image_matches_oracle(pair, flags):
    """B1-B7: the wavefront loop (queues, compaction, material sort) against one scalar loop per path."""
    name, sc, _, orc = pair
    spp, bounces = 2, 4
    tr = pt.PathTracer(sc.width, sc.height, seed=7, flags=flags, background=sc.background)
    tr.load(sc)
    tr.render(0, 1, bounces)
    tr.render(1, spp - 1, bounces)          # progressive: second call continues the accumulation
    g = tr.read_accum()
    st = tr.stats()
    rc = np.zeros(2, np.uint64)
    o = orc.render(sc.width, sc.height, 7, 0, spp, bounces, flags=flags, background=sc.background, ray_counts=rc)
    e = rmse(g[..., :3] / spp, o[..., :3] / spp)
    differing = int((g != o).any(axis=-1).sum())
    print(f"{name} flags={flags}: rmse={e:.3e} differing_pixels={differing} mean={o[..., :3].mean() / spp:.4f}")
    assert e <= 1e-3, f"{name}: image RMSE {e}"
    assert differing == 0, f"{name}: {differing} pixels not bit-identical"
    assert np.all(g[..., 3] == spp)
    tr.close()


@pytest.mark.parametrize("wave_samples", [1, 2])
def test_two_waves_in_flight_keep_the_frame_bit_identical(pair, monkeypatch, wave_samples):
    """A render of several waves alternates them between two streams / two copies of the wavefront state (the tails of one wave's traversal kernels overlap the other
    wave's kernels); the accumulations stay chained in sample order, so the frame equals the one-wave-at-a-time frame and the oracle's bit for bit, and the ray counters add up."""
    name, sc, _, orc = pair
    spp, bounces = 5, 4
    frames, rays = [], []
    for dual in ("1", "0"):
        monkeypatch.setenv("FOUNDATION_PT_DUAL_WAVE", dual)
        monkeypatch.setenv("FOUNDATION_PT_WAVE_SAMPLES", str(wave_samples))      # 5 or 3 waves (the last one short) instead of one
        tr = pt.PathTracer(sc.width, sc.height, seed=11, background=sc.background)
        tr.load(sc)
        tr.render(0, spp, bounces)
        st = tr.stats()
        frames.append(tr.read_accum()); rays.append((st.rays_extend, st.rays_shadow))
        tr.render(0, 2, bounces); tr.render(2, spp - 2, bounces)                  # progressive continuation across calls, again several waves per call
        frames.append(tr.read_accum())
        tr.close()
    rc = np.zeros(2, np.uint64)
    o = orc.render(sc.width, sc.height, 11, 0, spp, bounces, background=sc.background, ray_counts=rc)
    for f in frames:
        assert f.tobytes() == o.tobytes(), f"{name}: frame differs from the oracle's"
    assert rays[0] == rays[1] and rays[0][0] > 0


def test_tile_partition_is_bit_identical(pair):
    """§8e: interleaved tiles -> the union of the ranks' images equals the single-GPU image exactly."""
    name, sc, _, orc = pair
    full = pt.PathTracer(sc.width, sc.height, seed=3, background=sc.background); full.load(sc)
    full.render(0, 1, 3)
    ref = full.read_accum(); full.close()
    total = np.zeros_like(ref)
    for rank in range(3):
        t = pt.PathTracer(sc.width, sc.height, seed=3, background=sc.background); t.load(sc)
        t.partition_set(rank, 3, 16)
        t.render(0, 1, 3)
        part = t.read_accum(); t.close()
        assert np.all((part[..., 3] == 0) | (total[..., 3] == 0)), "tiles overlap"
        total += part
        o = orc.render(sc.width, sc.height, 3, 0, 1, 3, background=sc.background, rank=rank, count=3, tile=16)
        assert np.array_equal(part, o), f"{name}: rank {rank} differs from oracle partition"
    assert np.array_equal(total, ref)


def test_resolve_rgba8(pair):
    name, sc, tr, _ = pair
    tr.render(0, 2, 2)
    acc = tr.read_accum()
    img = tr.resolve_rgba8()
    want = (np.clip(acc[..., :3] / acc[..., 3:4], 0, 1) * np.float32(255.0) + np.float32(0.5)).astype(np.uint8)
    assert np.abs(img[..., :3].astype(int) - want.astype(int)).max() <= 1
    assert np.all(img[..., 3] == 255)


def test_vertex_formats(gpu):
    """R16_UINT / R32_UINT / unindexed, and an over-aligned vertex stride (reference: vertex_input, Renderer.cpp:23-27,110-115)."""
    sc = scenes.cornell_box(64, 64)
    m = sc.meshes[0]
    rays = ray_mix(sc, 1024)
    outs = []
    for variant in ("u32", "u16", "soup", "stride32"):
        t = pt.PathTracer(64, 64)
        t.materials_set(sc.materials)
        if variant == "u32":
            t.mesh_create(m.positions, m.indices, m.material_ids)
        elif variant == "u16":
            t.mesh_create(m.positions, m.indices.astype(np.uint16), m.material_ids)
        elif variant == "soup":
            t.mesh_create(m.positions[m.indices.reshape(-1)], None, m.material_ids)
        else:
            padded = np.zeros((m.positions.shape[0], 8), np.float32); padded[:, :3] = m.positions; padded[:, 3:] = 7.0
            t.mesh_create(padded, m.indices, m.material_ids, stride=32)
        t.scene_commit()
        outs.append(t.trace_closest(rays)[0])
        t.close()
    for o in outs[1:]:
        assert o.tobytes() == outs[0].tobytes()


def test_error_behaviour(gpu):
    """C ABI never throws: bad arguments / call order give negative status + message (SURVEY.md §8b error convention)."""
    t = pt.PathTracer(32, 32)
    with pytest.raises(pt.FoundationPtError) as e:
        t.render(0, 1, 1)
    assert e.value.status == pt.ERR_STATE
    with pytest.raises(pt.FoundationPtError) as e:
        t.mesh_create(np.zeros((3, 3), np.float32), np.asarray([[0, 1, 5]], np.uint32))
    assert e.value.status == pt.ERR_ARGUMENT
    with pytest.raises(pt.FoundationPtError) as e:
        t.scene_commit()
    assert e.value.status == pt.ERR_STATE
    t.mesh_create(np.asarray([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), np.asarray([[0, 1, 2]], np.uint32))
    t.scene_commit()
    with pytest.raises(pt.FoundationPtError):
        t.render(0, 1, 1)            # camera not set
    bad = np.zeros(1, scenes.INSTANCE_DTYPE)                 # singular instance transform: reported by the commit that follows
    t.instances_set(bad)
    with pytest.raises(pt.FoundationPtError) as e:
        t.scene_commit()
    assert e.value.status == pt.ERR_ARGUMENT and "singular" in str(e.value)
    ident = np.zeros(1, scenes.INSTANCE_DTYPE); ident["transform"][0] = (1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0)
    t.instances_set(ident); t.scene_commit()
    # single-triangle scene, empty ray set, degenerate rays
    hits, _ = t.trace_closest(np.zeros(0, scenes.RAY_DTYPE))
    assert len(hits) == 0
    r = np.zeros(3, scenes.RAY_DTYPE)
    r["origin"] = [(0.2, 0.2, 1), (0.2, 0.2, 1), (5, 5, 1)]; r["direction"] = [(0, 0, -1), (0, 0, 0), (0, 0, -1)]; r["tmax"] = np.inf
    hits, _ = t.trace_closest(r)
    assert hits["prim"][0] == 0 and hits["t"][0] == 1.0 and hits["prim"][1] == 0xFFFFFFFF and hits["prim"][2] == 0xFFFFFFFF
    t.close()


def test_cpp_renderer_host_draw_loop_matches_oracle(gpu, tmp_path):
    """The C++ host shaped like the reference's Renderer (ctor uploads + builds, Draw() adds a sample batch and resolves to
    RGBA8) driven by the headless Editor: its accumulation buffer equals the oracle's render of the same scene file."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rdir = os.path.join(root, "foundation_b200", "renderer")
    subprocess.run(["make", "-C", rdir], check=True, capture_output=True)
    sc = SMALL_SCENES["cornell"]()
    path = str(tmp_path / "cornell.fpts"); raw = str(tmp_path / "acc.raw"); ppm = str(tmp_path / "out.ppm")
    scenes.save_scene(sc, path)
    pfm = str(tmp_path / "out.pfm")
    out = subprocess.run([os.path.join(rdir, "foundation_editor"), path, "3", "2", "4", raw, ppm, "0", pfm], check=True, capture_output=True, text=True).stdout
    assert "spp=6" in out and "Memory Used: 0 bytes" in out, out          # three Draw() calls of two samples; nothing leaked through Core::Allocator
    acc = np.fromfile(raw, np.float32).reshape(sc.height, sc.width, 4)
    o = OracleScene(sc).render(sc.width, sc.height, 7, 0, 6, 4, background=sc.background)
    assert np.array_equal(acc, o)
    assert os.path.getsize(ppm) > sc.width * sc.height * 3
    from foundation_b200 import imageio
    assert np.array_equal(imageio.read_pfm(pfm), imageio.radiance_from_accum(o))   # the C++ writer and the Python resolve agree bit for bit


def test_image_output_pfm_png(gpu, tmp_path):
    """save_pfm / save_png write what read_accum / resolve_rgba8 return."""
    from foundation_b200 import imageio
    sc = SMALL_SCENES["cornell"]()
    with pt.PathTracer(sc.width, sc.height, seed=3) as tr:
        tr.load(sc)
        tr.render(0, 2, 3)
        tr.save_pfm(str(tmp_path / "a.pfm")); tr.save_png(str(tmp_path / "a.png"))
        assert np.array_equal(imageio.read_pfm(str(tmp_path / "a.pfm")), imageio.radiance_from_accum(tr.read_accum()))
        assert np.array_equal(imageio.read_png(str(tmp_path / "a.png")), tr.resolve_rgba8())


@pytest.mark.parametrize("leaf", [1, 2, 3])
def test_leaf_sizes_and_far_origins(gpu, leaf):
    """max_leaf_tris 1..3 build byte-identically to the oracle and give identical hits, including for rays that start ~100 scene
    radii away (exercises the ray-dependent slab slack in instanced BLAS)."""
    sc = scenes.instanced_patches(num_instances=60, patch=8, width=64, height=64)
    lo, hi = scenes.scene_bounds(sc)
    rng = np.random.default_rng(5)
    n = 200000
    c = (lo + hi) / 2; half = (hi - lo) / 2
    dirs = rng.normal(size=(n, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    org = c + dirs * np.linalg.norm(half) * 100.0
    tgt = c + (rng.random((n, 3)) * 2 - 1) * half
    rays = np.zeros(n, scenes.RAY_DTYPE)
    rays["origin"] = org.astype(np.float32); d = tgt - org; d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays["direction"] = d.astype(np.float32); rays["tmax"] = np.inf
    tr = pt.PathTracer(64, 64, max_leaf_tris=leaf); tr.load(sc)
    orc = OracleScene(sc, max_leaf=leaf)
    gn, gt, go = tr.blas_download(0); on, ot, oo = orc.blas(0)
    assert gn.tobytes() == on.tobytes() and gt.tobytes() == ot.tobytes()
    tr.rays_upload(rays)
    tr.rays_trace_closest(); gh, gi = tr.rays_download_hits()
    tr.rays_trace_brute(); bh, bi = tr.rays_download_hits()
    assert_hits_equal(gh, gi, bh, bi, f"leaf {leaf}: device BVH vs device exhaustive, far origins")
    oh, oi = orc.trace_closest(rays[:20000])
    assert gh[:20000].tobytes() == oh.tobytes() and np.array_equal(gi[:20000], oi)
    tr.close()


def test_checkpoint_resume_is_bit_identical(gpu, tmp_path):
    """Progressive render interrupted and resumed in a NEW context from a file equals the uninterrupted render (§8f rank 4)."""
    sc = SMALL_SCENES["spheres"]()
    a = pt.PathTracer(sc.width, sc.height, seed=21, background=sc.background); a.load(sc)
    a.render(0, 5, 4)
    ck = str(tmp_path / "frame.fpck")
    a.save_checkpoint(ck, 5, 21)
    a.render(5, 3, 4)
    full = a.read_accum(); a.close()
    b = pt.PathTracer(sc.width, sc.height, seed=21, background=sc.background); b.load(sc)
    done, seed = b.load_checkpoint(ck)
    assert (done, seed) == (5, 21)
    b.render(done, 3, 4)
    assert np.array_equal(b.read_accum(), full)
    b.close()


def test_moving_instances_rebuild_only_the_tlas(gpu):
    """SURVEY.md §8f rank 3 (the reference animates its model matrix every Draw, Renderer.cpp:373): instances_set + scene_commit on an
    already committed scene keeps every BLAS and the flat BLAS arrays, rebuilds the TLAS, and traces exactly like a fresh build."""
    sc = SMALL_SCENES["instanced"]()
    tr = pt.PathTracer(sc.width, sc.height, seed=2, background=sc.background)
    tr.load(sc)
    full_launches = tr.stats().kernel_launches
    blas_before = [x.tobytes() for x in tr.blas_download(0)]
    moved = sc.instances.copy()
    rng = np.random.default_rng(3)
    n = len(moved) - 1                                        # last instance is the light: leave it
    ang = rng.uniform(0, 2 * np.pi, n); s = rng.uniform(0.7, 1.4, n)
    T = moved["transform"][:n].copy()
    T[:, 0] = s * np.cos(ang); T[:, 1] = -s * np.sin(ang); T[:, 4] = s * np.sin(ang); T[:, 5] = s * np.cos(ang); T[:, 10] = s
    T[:, 3] += rng.uniform(-0.3, 0.3, n); T[:, 7] += rng.uniform(-0.3, 0.3, n); T[:, 11] += rng.uniform(0.0, 0.5, n)
    moved["transform"][:n] = T
    tr.instances_set(moved)
    tr.scene_commit()
    assert tr.stats().kernel_launches < full_launches, "second commit should skip the BLAS builds"
    assert [x.tobytes() for x in tr.blas_download(0)] == blas_before
    sc2 = scenes.Scene(sc.name, sc.meshes, sc.materials, moved, sc.view, sc.proj, sc.width, sc.height, sc.background)
    orc = OracleScene(sc2)
    gn, go = tr.tlas_download(); on, oo, _ = orc.tlas()
    assert np.array_equal(go, oo) and gn.tobytes() == on.tobytes()
    rays = ray_mix(sc2)
    gh, gi = tr.trace_closest(rays); oh, oi = orc.trace_closest(rays)
    assert gh.tobytes() == oh.tobytes() and np.array_equal(gi, oi)
    tr.render(0, 2, 3)
    assert np.array_equal(tr.read_accum(), orc.render(sc.width, sc.height, 2, 0, 2, 3, background=sc.background))
    tr.close()


def test_several_meshes_without_instances(gpu):
    """scene_commit's implicit identity instances (one per mesh) against the oracle's same rule: hits, instance ids, image."""
    sc = multi_mesh_scene()
    tr = pt.PathTracer(sc.width, sc.height, seed=4); tr.load(sc)
    orc = OracleScene(sc)
    rays = ray_mix(sc, 4096)
    gh, gi = tr.trace_closest(rays); oh, oi = orc.trace_closest(rays)
    assert gh.tobytes() == oh.tobytes() and np.array_equal(gi, oi) and set(np.unique(gi[gh["prim"] != 0xFFFFFFFF])) == {0, 1, 2}
    tr.render(0, 2, 4)
    assert np.array_equal(tr.read_accum(), orc.render(sc.width, sc.height, 4, 0, 2, 4))
    tr.close()


def test_degenerate_and_duplicate_triangles_on_device(gpu):
    from tests.test_oracle import degenerate_scene
    sc = degenerate_scene()
    tr = pt.PathTracer(32, 32); tr.load(sc)
    orc = OracleScene(sc)
    gn, gt, go = tr.blas_download(0); on, ot, oo = orc.blas(0)
    assert np.array_equal(go, oo) and gn.tobytes() == on.tobytes() and gt.tobytes() == ot.tobytes()
    lo, hi = scenes.scene_bounds(sc)
    rays = np.concatenate([scenes.incoherent_rays(lo, hi, 20000, 3), scenes.incoherent_rays(np.zeros(3), np.ones(3), 20000, 4)])
    gh, gi = tr.trace_closest(rays); oh, oi = orc.trace_closest(rays)
    assert gh.tobytes() == oh.tobytes()
    tr.rays_upload(rays); tr.rays_trace_brute(); bh, bi = tr.rays_download_hits()
    assert_hits_equal(gh, gi, bh, bi, "degenerate mesh: device BVH vs device exhaustive")
    tr.close()


def test_cpp_renderer_host_spinning_instances(gpu, tmp_path):
    """The C++ Draw() loop with a per-frame model rotation (reference: Renderer.cpp:373): every frame replaces the instance list, the
    backend rebuilds only the TLAS, and the last frame equals the oracle's render of the rotated scene."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rdir = os.path.join(root, "foundation_b200", "renderer")
    subprocess.run(["make", "-C", rdir], check=True, capture_output=True)
    sc = SMALL_SCENES["instanced"]()
    path = str(tmp_path / "inst.fpts"); raw = str(tmp_path / "acc.raw")
    scenes.save_scene(sc, path)
    frames, spin = 3, 20.0
    out = subprocess.run([os.path.join(rdir, "foundation_editor"), path, str(frames), "2", "3", raw, str(tmp_path / "o.ppm"), str(spin)], check=True, capture_output=True,
                         text=True).stdout
    assert "spp=2" in out and "Memory Used: 0 bytes" in out, out
    a = frames * spin * np.pi / 180.0
    c, s = np.float32(np.cos(a)), np.float32(np.sin(a))
    moved = sc.instances.copy()
    T = sc.instances["transform"].reshape(-1, 3, 4); M = moved["transform"].reshape(-1, 3, 4)
    M[:, 0, :] = c * T[:, 0, :] - s * T[:, 1, :]; M[:, 1, :] = s * T[:, 0, :] + c * T[:, 1, :]
    sc2 = scenes.Scene(sc.name, sc.meshes, sc.materials, moved, sc.view, sc.proj, sc.width, sc.height, sc.background)
    acc = np.fromfile(raw, np.float32).reshape(sc.height, sc.width, 4)
    o = OracleScene(sc2).render(sc.width, sc.height, 7, 0, 2, 3, background=sc.background)
    assert np.array_equal(acc, o)


def test_pathological_rays_terminate_and_match_oracle(gpu):
    """NaN / inf / denormal / zero-length / inverted-interval rays: the traversal terminates and reports what the oracle reports."""
    sc = SMALL_SCENES["terrain"]()
    tr = pt.PathTracer(sc.width, sc.height); tr.load(sc)
    orc = OracleScene(sc)
    lo, hi = scenes.scene_bounds(sc)
    rays = scenes.incoherent_rays(lo, hi, 4096, 8)
    rays["origin"][0::16, 0] = np.nan
    rays["direction"][1::16] = (np.inf, 0, -1)
    rays["direction"][2::16] = (0, 0, 0)
    rays["direction"][3::16] *= np.float32(1e-42)                 # denormal direction
    rays["tmax"][4::16] = -1.0                                    # empty interval
    rays["tmin"][5::16] = np.inf
    rays["direction"][6::16] *= np.float32(1e30)                  # huge direction: hits at tiny t
    rays["origin"][7::16] = (1e30, 1e30, 1e30)
    rays["direction"][8::16, 2] = 0.0                             # axis-parallel rays (zero component -> clamped reciprocal)
    rays["direction"][9::16, :2] = 0.0
    gh, gi = tr.trace_closest(rays); oh, oi = orc.trace_closest(rays)
    assert np.array_equal(gh["prim"], oh["prim"]) and gh.tobytes() == oh.tobytes()
    assert np.array_equal(tr.trace_any(rays), orc.trace_any(rays))
    tr.close()


@pytest.mark.parametrize("ntris", [1, 2, 3, 255, 256, 257, 511, 512, 513, 1025, 70001])
def test_build_sizes_around_the_refit_tiles(gpu, ntris):
    """A4: the tiled refit (256 sorted leaves per block, rounds in shared memory, upper levels in a second kernel) must give the oracle's
    nodes byte for byte for triangle counts below, at and just above tile multiples, and for a mesh whose Morton keys are all equal
    except for a few (deep, chain-like radix tree: many rounds inside one tile)."""
    n = int(np.ceil(np.sqrt(ntris / 2.0))) + 1
    base = scenes.fractal_terrain(n=n, width=32, height=32, with_light=False)
    m = base.meshes[0]
    idx = np.ascontiguousarray(m.indices[:ntris]); mat = np.ascontiguousarray(m.material_ids[:ntris])
    assert idx.shape[0] == ntris
    sc = scenes.Scene(f"terrain_first_{ntris}", [scenes.Mesh(m.positions, idx, mat)], base.materials, None, base.view, base.proj, 32, 32)
    tr = pt.PathTracer(32, 32); tr.load(sc)
    orc = OracleScene(sc)
    gn, gt, go = tr.blas_download(0); on, ot, oo = orc.blas(0)
    assert np.array_equal(go, oo) and gn.tobytes() == on.tobytes() and gt.tobytes() == ot.tobytes(), ntris
    lo, hi = scenes.scene_bounds(sc)
    rays = scenes.incoherent_rays(lo, hi, 4096, 9)
    gh, gi = tr.trace_closest(rays); oh, oi = orc.trace_closest(rays)
    assert gh.tobytes() == oh.tobytes()
    tr.close()


def test_build_of_a_chain_like_radix_tree(gpu):
    """600 copies of one triangle (identical Morton keys: the radix tree is decided by the index tie-break) plus a few far-away ones."""
    rng = np.random.default_rng(5)
    tri = np.asarray([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    pos = np.concatenate([np.tile(tri, (600, 1)), (rng.uniform(-50, 50, (7, 1, 3)) + tri[None]).reshape(-1, 3).astype(np.float32)])
    idx = np.arange(pos.shape[0], dtype=np.uint32).reshape(-1, 3)
    base = scenes.cornell_box(32, 32)
    sc = scenes.Scene("chain", [scenes.Mesh(np.ascontiguousarray(pos), idx, np.zeros(idx.shape[0], np.uint32))], base.materials, None, base.view, base.proj, 32, 32)
    tr = pt.PathTracer(32, 32); tr.load(sc)
    orc = OracleScene(sc)
    gn, gt, go = tr.blas_download(0); on, ot, oo = orc.blas(0)
    assert np.array_equal(go, oo) and gn.tobytes() == on.tobytes() and gt.tobytes() == ot.tobytes()
    rays = scenes.incoherent_rays(np.full(3, -2.0), np.full(3, 2.0), 4096, 10)
    gh, gi = tr.trace_closest(rays); oh, oi = orc.trace_closest(rays)
    assert gh.tobytes() == oh.tobytes()
    tr.close()


def test_render_async_plus_wait_equals_render(gpu):
    """render_async enqueues, wait completes; any other entry point in between waits implicitly; a second render_async before the wait is a
    state error.  The frame equals the blocking render's bit for bit."""
    sc = SMALL_SCENES["cornell"]()
    a = pt.PathTracer(sc.width, sc.height, seed=3, background=sc.background); a.load(sc)
    b = pt.PathTracer(sc.width, sc.height, seed=3, background=sc.background); b.load(sc)
    a.render(0, 4, 4)
    b.render_async(0, 2, 4)
    with pytest.raises(pt.FoundationPtError) as e:
        b.render_async(2, 2, 4)
    assert e.value.status == pt.ERR_STATE
    b.wait()
    st = b.stats()
    assert st.rays_extend > 0 and st.last_ms > 0
    b.render_async(2, 2, 4)
    img = b.read_accum()                     # no explicit wait: read_accum settles the pending render first
    assert np.array_equal(img, a.read_accum())
    b.wait()                                 # nothing in flight: a no-op
    a.close(); b.close()


def test_stage_timing_flag_reports_where_a_render_spends_its_time(gpu):
    """FOUNDATION_PT_FLAG_STAGE_TIMING (CUDA events around each wavefront stage): the per-stage times add up to the
    call's device time, the frame is unchanged, and without the flag the array stays zero."""
    sc = SMALL_SCENES["terrain"]()
    a = pt.PathTracer(sc.width, sc.height, seed=2, background=sc.background); a.load(sc)
    b = pt.PathTracer(sc.width, sc.height, seed=2, background=sc.background, flags=pt.FLAG_STAGE_TIMING); b.load(sc)
    a.render(0, 8, 4); b.render(0, 8, 4)
    assert np.array_equal(a.read_accum(), b.read_accum())
    sa, sb = a.stats(), b.stats()
    assert all(v == 0.0 for v in sa.stage_ms)
    st = dict(zip(pt.STAGE_NAMES, sb.stage_ms))
    assert st["extend"] > 0 and st["shade"] > 0 and st["connect"] > 0 and st["raygen"] > 0 and st["accumulate"] > 0 and st["material_sort"] == 0.0
    assert 0.7 * sb.last_ms <= sum(sb.stage_ms) <= 1.02 * sb.last_ms, (sb.last_ms, st)
    a.close(); b.close()


def test_deforming_mesh_rebuilds_its_blas_and_matches_a_fresh_build(gpu):
    """SURVEY.md section 8f rank 3, deforming geometry: mesh_update_positions + scene_commit rebuilds the BLAS of that mesh only and gives exactly
    what a fresh context builds from the deformed vertices (nodes byte for byte, hits, image) — flat scene and instanced scene."""
    for name in ("terrain", "instanced"):
        sc = SMALL_SCENES[name]()
        tr = pt.PathTracer(sc.width, sc.height, seed=6, background=sc.background); tr.load(sc)
        other_before = [x.tobytes() for x in tr.blas_download(1)] if len(sc.meshes) > 1 else None
        m = sc.meshes[0]
        rng = np.random.default_rng(12)
        pos2 = (m.positions + rng.normal(0, 0.02 * float(np.ptp(m.positions[:, 2]) + 1e-3), m.positions.shape)).astype(np.float32)
        tr.mesh_update_positions(0, pos2)
        tr.scene_commit()
        sc2 = scenes.Scene(sc.name, [scenes.Mesh(pos2, m.indices, m.material_ids)] + sc.meshes[1:], sc.materials, sc.instances, sc.view, sc.proj, sc.width, sc.height, sc.background)
        orc = OracleScene(sc2)
        gn, gt, go = tr.blas_download(0); on, ot, oo = orc.blas(0)
        assert np.array_equal(go, oo) and gn.tobytes() == on.tobytes() and gt.tobytes() == ot.tobytes(), name
        if other_before is not None:
            assert [x.tobytes() for x in tr.blas_download(1)] == other_before            # untouched meshes keep their BLAS
        rays = ray_mix(sc2)
        gh, gi = tr.trace_closest(rays); oh, oi = orc.trace_closest(rays)
        assert gh.tobytes() == oh.tobytes() and np.array_equal(gi, oi), name
        tr.render(0, 2, 3)
        assert np.array_equal(tr.read_accum(), orc.render(sc.width, sc.height, 6, 0, 2, 3, background=sc.background)), name
        with pytest.raises(pt.FoundationPtError):
            tr.mesh_update_positions(0, pos2[:-1])                                        # topology is fixed
        tr.close()


def test_changing_the_emissive_materials_refreshes_the_cached_light_lists(gpu):
    """The emissive triangles of a mesh are extracted once and reused by later commits (moving instances, deformed meshes); the cache is keyed by the emissive
    set of the materials, so a commit after materials_set with a different emitter must light the scene from the new emitter — image vs the oracle on the edited scene."""
    import copy
    sc = scenes.cornell_box(96, 96)
    tr = pt.PathTracer(sc.width, sc.height, seed=13, background=sc.background)
    tr.load(sc)
    tr.render(0, 2, 3)
    first = tr.read_accum().copy()
    sc2 = copy.copy(sc)
    mats = sc.materials.copy()
    em = mats[:, 4:7].copy()
    light = int(np.argmax(em.sum(axis=1)))
    mats[light, 4:7] = 0.0                                  # the ceiling panel goes dark ...
    mats[1, 4:7] = (6.0, 2.0, 1.0)                          # ... and the red wall glows instead
    sc2.materials = mats
    tr.materials_set(mats)
    tr.scene_commit()
    tr.render(0, 2, 3)
    second = tr.read_accum().copy()
    o = OracleScene(sc2).render(sc.width, sc.height, 13, 0, 2, 3, background=sc.background)
    assert second.tobytes() == o.tobytes(), "frame after the material edit differs from the oracle's"
    assert first.tobytes() != second.tobytes()
    tr.materials_set(sc.materials)                          # and back: the first list is extracted again, not the stale one reused
    tr.scene_commit()
    tr.render(0, 2, 3)
    assert tr.read_accum().tobytes() == first.tobytes()
    tr.close()
