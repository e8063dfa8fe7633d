"""ctypes binding of the C ABI in include/foundation_pt.h, plus `PathTracer`, a thin object wrapper.

This is the host side a Python caller uses; the C++ host side shaped like the reference's
`Foundation::Renderer::Renderer` lives in foundation_b200/renderer/.  Both call only `foundation_pt_*`.
There is no CPU fallback: constructing a PathTracer without a CUDA device raises FoundationPtError
(status FOUNDATION_PT_ERR_NO_DEVICE), and a missing libfoundation_pt.so raises at load time.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import build as _build
from .scenes import HIT_DTYPE, RAY_DTYPE, Scene  # noqa: F401

OK, ERR_ARGUMENT, ERR_STATE, ERR_CUDA, ERR_OOM, ERR_NO_DEVICE, ERR_UNSUPPORTED, ERR_COMM = 0, -1, -2, -3, -4, -5, -6, -7
COMM_ID_BYTES, COMM_DIRECT = 128, 1
FLAG_NO_MATERIAL_SORT, FLAG_NO_NEE, FLAG_NO_BSDF_EMISSION, FLAG_MATERIAL_SORT, FLAG_SOBOL_JITTER, FLAG_SOBOL_PATH, FLAG_STAGE_TIMING = 1, 2, 4, 8, 16, 32, 64
STAGE_NAMES = ("raygen", "extend", "shade", "connect", "material_sort", "accumulate")

NODE_DTYPE = np.dtype([("p", "<f4", (3,)), ("e", "u1", (3,)), ("imask", "u1"), ("child_base", "<u4"), ("tri_base", "<u4"), ("meta", "u1", (8,)),
                       ("qlo", "u1", (3, 8)), ("qhi", "u1", (3, 8))])
TRI_DTYPE = np.dtype([("v0", "<f4", (3,)), ("prim", "<u4"), ("e1", "<f4", (3,)), ("mat", "<u4"), ("e2", "<f4", (3,)), ("pad", "<u4")])


class Config(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("width", C.c_uint32), ("height", C.c_uint32), ("seed", C.c_uint64),
                ("max_leaf_tris", C.c_uint32), ("flags", C.c_uint32), ("background", C.c_float * 3), ("reserved", C.c_uint32)]


class BuildStats(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("num_meshes", C.c_uint32), ("num_instances", C.c_uint32), ("num_triangles", C.c_uint64),
                ("effective_triangles", C.c_uint64), ("num_nodes8", C.c_uint64), ("device_bytes", C.c_uint64), ("build_ms", C.c_float),
                ("sort_ms", C.c_float), ("scene_lo", C.c_float * 3), ("scene_hi", C.c_float * 3)]


class Stats(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("kernel_launches", C.c_uint32), ("rays_extend", C.c_uint64), ("rays_shadow", C.c_uint64),
                ("last_ms", C.c_float), ("trace_ms", C.c_float), ("gather_ms", C.c_float), ("reserved", C.c_uint32), ("total_launches", C.c_uint64),
                ("stage_ms", C.c_float * 6)]


# every symbol include/foundation_pt.h declares (tests check the library exports exactly these)
SYMBOLS = {
    "foundation_pt_version": (C.c_uint32, []),
    "foundation_pt_last_error": (C.c_char_p, [C.c_void_p]),
    "foundation_pt_create": (C.c_int32, [C.POINTER(Config), C.c_void_p, C.POINTER(C.c_void_p)]),
    "foundation_pt_destroy": (C.c_int32, [C.c_void_p]),
    "foundation_pt_materials_set": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint32]),
    "foundation_pt_mesh_create": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32)]),
    "foundation_pt_instances_set": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint32]),
    "foundation_pt_mesh_update_positions": (C.c_int32, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t, C.c_uint32]),
    "foundation_pt_mesh_attributes_set": (C.c_int32, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "foundation_pt_texture_create": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_size_t, C.POINTER(C.c_uint32)]),
    "foundation_pt_material_textures_set": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint32]),
    "foundation_pt_scene_commit": (C.c_int32, [C.c_void_p, C.POINTER(BuildStats)]),
    "foundation_pt_camera_set": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "foundation_pt_partition_set": (C.c_int32, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]),
    "foundation_pt_render": (C.c_int32, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]),
    "foundation_pt_render_async": (C.c_int32, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]),
    "foundation_pt_wait": (C.c_int32, [C.c_void_p]),
    "foundation_pt_read_accum": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "foundation_pt_write_accum": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "foundation_pt_resolve_rgba8": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "foundation_pt_accum_device_ptr": (C.c_int32, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "foundation_pt_trace_closest": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "foundation_pt_trace_any": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "foundation_pt_rays_upload": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "foundation_pt_rays_trace_closest": (C.c_int32, [C.c_void_p, C.c_uint64, C.c_uint64]),
    "foundation_pt_rays_trace_any": (C.c_int32, [C.c_void_p, C.c_uint64, C.c_uint64]),
    "foundation_pt_rays_trace_brute": (C.c_int32, [C.c_void_p, C.c_uint64, C.c_uint64]),
    "foundation_pt_rays_download_hits": (C.c_int32, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]),
    "foundation_pt_stats_get": (C.c_int32, [C.c_void_p, C.POINTER(Stats)]),
    "foundation_pt_blas_download": (C.c_int32, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                                C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "foundation_pt_comm_unique_id": (C.c_int32, [C.c_void_p, C.c_size_t]),
    "foundation_pt_comm_init": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]),
    "foundation_pt_gather": (C.c_int32, [C.c_void_p, C.c_uint32]),
    "foundation_pt_gather_local": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "foundation_pt_group_create": (C.c_int32, [C.POINTER(Config), C.POINTER(C.c_int32), C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_void_p)]),
    "foundation_pt_group_destroy": (C.c_int32, [C.c_void_p]),
    "foundation_pt_group_size": (C.c_uint32, [C.c_void_p]),
    "foundation_pt_group_context": (C.c_void_p, [C.c_void_p, C.c_uint32]),
    "foundation_pt_group_render": (C.c_int32, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]),
    "foundation_pt_group_last_error": (C.c_char_p, [C.c_void_p]),
    "foundation_pt_tlas_download": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
}

_LIB = None


class FoundationPtError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"foundation_pt status {status}: {message}")
        self.status = status


def load_library(path: Optional[str] = None) -> C.CDLL:
    """Loads libfoundation_pt.so (building it in-tree first if nvcc is present and it is stale).  Raises if it cannot."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    p = path or _build.LIB_PATH
    if path is None:
        try:
            _build.build()
        except Exception:
            if not os.path.exists(p):
                raise
    lib = C.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        if os.environ.get("FOUNDATION_PT_LIB") and not hasattr(lib, name):
            continue                    # A/B experiments against an older build only
        fn = getattr(lib, name)          # AttributeError if the symbol is missing: fail loudly
        fn.restype = res; fn.argtypes = args
    if path is None:
        _LIB = lib
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class PathTracer:
    """One context = one GPU.  Mirrors the call sequence of the C ABI one to one."""

    def __init__(self, width: int = 1920, height: int = 1080, device: int = 0, seed: int = 1, max_leaf_tris: int = 0, flags: int = 0,
                 background=(0.0, 0.0, 0.0)):
        self._lib = load_library()
        cfg = Config(C.sizeof(Config), device, width, height, seed, max_leaf_tris, flags, (C.c_float * 3)(*background), 0)
        self._ctx = C.c_void_p()
        st = self._lib.foundation_pt_create(C.byref(cfg), None, C.byref(self._ctx))
        if st != OK:
            raise FoundationPtError(st, self._lib.foundation_pt_last_error(None).decode())
        self.width, self.height, self.device = width, height, device
        self.build_stats: Optional[BuildStats] = None
        self._nrays = 0

    # -- plumbing
    def _check(self, st: int):
        if st != OK:
            raise FoundationPtError(st, self._lib.foundation_pt_last_error(self._ctx).decode())

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._lib.foundation_pt_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- scene
    def materials_set(self, materials):
        m = np.ascontiguousarray(materials, np.float32).reshape(-1, 8)
        self._check(self._lib.foundation_pt_materials_set(self._ctx, _p(m), m.shape[0]))

    def mesh_create(self, positions, indices=None, material_ids=None, stride: Optional[int] = None) -> int:
        pos = np.ascontiguousarray(positions)
        if stride is None:
            pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3); stride = 12; nverts = pos.shape[0]
        else:
            nverts = (pos.nbytes - 12) // stride + 1
        if indices is None:
            idx, fmt, ntris = None, 0, nverts // 3
        else:
            idx = np.ascontiguousarray(indices)
            if idx.dtype == np.uint16:
                fmt = 16
            else:
                idx = np.ascontiguousarray(idx, np.uint32); fmt = 32
            ntris = idx.size // 3
        mat = None if material_ids is None else np.ascontiguousarray(material_ids, np.uint32)
        out = C.c_uint32()
        self._check(self._lib.foundation_pt_mesh_create(self._ctx, _p(pos), stride, nverts, _p(idx), fmt, ntris, _p(mat), C.byref(out)))
        return out.value

    def mesh_update_positions(self, mesh_id: int, positions, stride: int = 12):
        """New vertex positions for an existing mesh (same count / stride / indices); its BLAS is rebuilt by the next scene_commit."""
        pos = np.ascontiguousarray(positions)
        if stride == 12 and pos.dtype != np.uint8:
            pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3); nverts = pos.shape[0]
        else:
            nverts = (pos.nbytes - 12) // stride + 1
        self._check(self._lib.foundation_pt_mesh_update_positions(self._ctx, mesh_id, _p(pos), stride, nverts))

    def mesh_attributes_set(self, mesh_id: int, uv=None, colors=None, uv_stride: int = 8, color_stride: int = 12):
        """Per-vertex uv (float2) / colour (float3) streams; pass strides for interleaved buffers."""
        uv = None if uv is None else np.ascontiguousarray(uv); colors = None if colors is None else np.ascontiguousarray(colors)
        if uv is not None and uv.dtype != np.uint8:
            uv = np.ascontiguousarray(uv, np.float32)
        if colors is not None and colors.dtype != np.uint8:
            colors = np.ascontiguousarray(colors, np.float32)
        self._check(self._lib.foundation_pt_mesh_attributes_set(self._ctx, mesh_id, _p(uv), uv_stride, _p(colors), color_stride))

    def texture_create(self, rgba8, row_pitch: int = 0) -> int:
        t = np.ascontiguousarray(rgba8, np.uint8); assert t.ndim == 3 and t.shape[2] == 4
        out = C.c_uint32()
        self._check(self._lib.foundation_pt_texture_create(self._ctx, _p(t), t.shape[1], t.shape[0], row_pitch, C.byref(out)))
        return out.value

    def material_textures_set(self, ids):
        a = np.ascontiguousarray(ids, np.uint32)
        self._check(self._lib.foundation_pt_material_textures_set(self._ctx, _p(a), a.shape[0]))

    def instances_set(self, instances):
        inst = np.ascontiguousarray(instances)
        assert inst.dtype.itemsize == 64
        self._check(self._lib.foundation_pt_instances_set(self._ctx, _p(inst), inst.shape[0]))

    def scene_commit(self) -> BuildStats:
        bs = BuildStats(); bs.struct_size = C.sizeof(BuildStats)
        self._check(self._lib.foundation_pt_scene_commit(self._ctx, C.byref(bs)))
        self.build_stats = bs
        return bs

    def load(self, scene: Scene) -> BuildStats:
        self.materials_set(scene.materials)
        for m in scene.meshes:
            mid = self.mesh_create(m.positions, m.indices, m.material_ids)
            if getattr(m, "uv", None) is not None or getattr(m, "colors", None) is not None:
                self.mesh_attributes_set(mid, m.uv, m.colors)
        for tex in getattr(scene, "textures", None) or []:
            self.texture_create(tex)
        if getattr(scene, "material_textures", None) is not None:
            self.material_textures_set(scene.material_textures)
        if scene.instances is not None:
            self.instances_set(scene.instances)
        bs = self.scene_commit()
        if scene.view is not None:
            self.camera_set(scene.view, scene.proj)
        return bs

    def camera_set(self, view, proj):
        v = np.ascontiguousarray(view, np.float32).reshape(16); p = np.ascontiguousarray(proj, np.float32).reshape(16)
        self._check(self._lib.foundation_pt_camera_set(self._ctx, _p(v), _p(p)))

    def partition_set(self, rank: int, count: int, tile: int = 32):
        self._check(self._lib.foundation_pt_partition_set(self._ctx, rank, count, tile))

    # -- multi-GPU frame (stage C1): one process per GPU
    @staticmethod
    def comm_unique_id() -> bytes:
        """128-byte communicator id (rank 0 creates it, the host distributes it to the other ranks by any channel)."""
        lib = load_library()
        buf = (C.c_uint8 * COMM_ID_BYTES)()
        st = lib.foundation_pt_comm_unique_id(buf, COMM_ID_BYTES)
        if st != OK:
            raise FoundationPtError(st, lib.foundation_pt_last_error(None).decode())
        return bytes(buf)

    def comm_init(self, comm_id: bytes, rank: int, count: int, tile: int = 32, flags: int = 0):
        """Collective over all ranks: joins the communicator and fixes this context's tile partition."""
        buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(comm_id)
        self._check(self._lib.foundation_pt_comm_init(self._ctx, buf, COMM_ID_BYTES, rank, count, tile, flags))

    def gather(self, root: int = 0):
        """Collective: afterwards rank `root`'s accumulation buffer holds the whole frame."""
        self._check(self._lib.foundation_pt_gather(self._ctx, root))

    def gather_local(self, src: "PathTracer"):
        """Same-device gather: scatters `src`'s owned tiles into this context's frame (no NCCL)."""
        self._check(self._lib.foundation_pt_gather_local(self._ctx, src._ctx))

    # -- render
    def render(self, sample_begin: int, sample_count: int, max_bounces: int):
        self._check(self._lib.foundation_pt_render(self._ctx, sample_begin, sample_count, max_bounces))

    def render_async(self, sample_begin: int, sample_count: int, max_bounces: int):
        """Enqueue a render and return at once; `wait()` (or any other call) completes it."""
        self._check(self._lib.foundation_pt_render_async(self._ctx, sample_begin, sample_count, max_bounces))

    def wait(self):
        self._check(self._lib.foundation_pt_wait(self._ctx))

    def read_accum(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.height, self.width, 4), np.float32)
        self._check(self._lib.foundation_pt_read_accum(self._ctx, _p(out), out.nbytes))
        return out

    def write_accum(self, accum: np.ndarray):
        a = np.ascontiguousarray(accum, np.float32)
        self._check(self._lib.foundation_pt_write_accum(self._ctx, _p(a), a.nbytes))

    # -- checkpoint / resume of a progressive render: the float4 sum + the sample counter + the stream seed
    def save_checkpoint(self, path: str, samples_done: int, seed: int):
        acc = self.read_accum()
        with open(path, "wb") as f:
            f.write(np.asarray([0x4b435046, 1, self.width, self.height, samples_done, 0], np.uint32).tobytes())   # 'FPCK'
            f.write(np.asarray([seed], np.uint64).tobytes())
            f.write(acc.tobytes())

    def load_checkpoint(self, path: str):
        """Returns (samples_done, seed); continue with render(samples_done, more, bounces) on a context created with that seed."""
        with open(path, "rb") as f:
            hdr = np.frombuffer(f.read(24), np.uint32)
            if hdr[0] != 0x4b435046 or hdr[1] != 1 or hdr[2] != self.width or hdr[3] != self.height:
                raise ValueError("not a checkpoint of this frame size")
            seed = int(np.frombuffer(f.read(8), np.uint64)[0])
            acc = np.frombuffer(f.read(self.width * self.height * 16), np.float32).reshape(self.height, self.width, 4)
        self.write_accum(acc)
        return int(hdr[4]), seed

    def resolve_rgba8(self) -> np.ndarray:
        out = np.empty((self.height, self.width, 4), np.uint8)
        self._check(self._lib.foundation_pt_resolve_rgba8(self._ctx, _p(out), out.nbytes))
        return out

    # -- image output: linear mean radiance as PFM / OpenEXR, the presented RGBA8 frame as PNG
    def save_pfm(self, path: str):
        from . import imageio
        imageio.write_pfm(path, imageio.radiance_from_accum(self.read_accum()))

    def save_png(self, path: str):
        from . import imageio
        imageio.write_png(path, self.resolve_rgba8())

    def save_exr(self, path: str):
        from . import imageio
        imageio.write_exr(path, imageio.radiance_from_accum(self.read_accum()))

    def accum_device_ptr(self):
        ptr = C.c_void_p(); size = C.c_size_t()
        self._check(self._lib.foundation_pt_accum_device_ptr(self._ctx, C.byref(ptr), C.byref(size)))
        return ptr.value, size.value

    # -- explicit ray sets
    def trace_closest(self, rays, hits=None, inst=None, want_inst: bool = True):
        rays = np.ascontiguousarray(rays); assert rays.dtype.itemsize == 32
        n = rays.shape[0]
        if hits is None:
            hits = np.empty(n, HIT_DTYPE)
        if inst is None and want_inst:
            inst = np.empty(n, np.uint32)
        self._check(self._lib.foundation_pt_trace_closest(self._ctx, _p(rays), n, _p(hits), _p(inst)))
        return hits, inst

    def trace_closest_raw(self, rays_ptr: int, n: int, hits_ptr: int, inst_ptr: int = 0):
        """Raw host pointers (e.g. pinned torch tensors): the e2e path of bench.py."""
        self._check(self._lib.foundation_pt_trace_closest(self._ctx, C.c_void_p(rays_ptr), n, C.c_void_p(hits_ptr), C.c_void_p(inst_ptr) if inst_ptr else None))

    def trace_any(self, rays):
        rays = np.ascontiguousarray(rays); n = rays.shape[0]
        occ = np.empty(n, np.uint8)
        self._check(self._lib.foundation_pt_trace_any(self._ctx, _p(rays), n, _p(occ)))
        return occ

    def rays_upload(self, rays):
        rays = np.ascontiguousarray(rays); assert rays.dtype.itemsize == 32
        self._check(self._lib.foundation_pt_rays_upload(self._ctx, _p(rays), rays.shape[0]))
        self._nrays = rays.shape[0]

    def rays_trace_closest(self, first: int = 0, count: Optional[int] = None):
        self._check(self._lib.foundation_pt_rays_trace_closest(self._ctx, first, self._nrays - first if count is None else count))

    def rays_trace_any(self, first: int = 0, count: Optional[int] = None):
        self._check(self._lib.foundation_pt_rays_trace_any(self._ctx, first, self._nrays - first if count is None else count))

    def rays_trace_brute(self, first: int = 0, count: Optional[int] = None):
        self._check(self._lib.foundation_pt_rays_trace_brute(self._ctx, first, self._nrays - first if count is None else count))

    def rays_download_hits(self, first: int = 0, count: Optional[int] = None):
        n = self._nrays - first if count is None else count
        hits = np.empty(n, HIT_DTYPE); inst = np.empty(n, np.uint32)
        self._check(self._lib.foundation_pt_rays_download_hits(self._ctx, first, n, _p(hits), _p(inst)))
        return hits, inst

    # -- introspection
    def stats(self) -> Stats:
        s = Stats(); s.struct_size = C.sizeof(Stats)
        self._check(self._lib.foundation_pt_stats_get(self._ctx, C.byref(s)))
        return s

    def blas_download(self, mesh_id: int = 0):
        nn = C.c_uint64(); nt = C.c_uint64()
        self._check(self._lib.foundation_pt_blas_download(self._ctx, mesh_id, None, 0, None, 0, None, 0, C.byref(nn), C.byref(nt)))
        nodes = np.zeros(nn.value, NODE_DTYPE); tris = np.zeros(nt.value, TRI_DTYPE); order = np.zeros(nt.value, np.uint32)
        self._check(self._lib.foundation_pt_blas_download(self._ctx, mesh_id, _p(nodes), nodes.nbytes, _p(tris), tris.nbytes, _p(order), order.nbytes,
                                                          C.byref(nn), C.byref(nt)))
        return nodes, tris, order

    def tlas_download(self):
        nn = C.c_uint64(); ni = C.c_uint64()
        self._check(self._lib.foundation_pt_tlas_download(self._ctx, None, 0, None, 0, C.byref(nn), C.byref(ni)))
        nodes = np.zeros(nn.value, NODE_DTYPE); order = np.zeros(ni.value, np.uint32)
        if nn.value:
            self._check(self._lib.foundation_pt_tlas_download(self._ctx, _p(nodes), nodes.nbytes, _p(order), order.nbytes, C.byref(nn), C.byref(ni)))
        return nodes, order


class _Member(PathTracer):
    """A context owned by a Group (borrowed handle: never destroyed on its own)."""

    def __init__(self, lib, ctx, width, height, device):
        self._lib, self._ctx, self.width, self.height, self.device = lib, C.c_void_p(ctx), width, height, device
        self.build_stats, self._nrays = None, 0

    def close(self):
        self._ctx = C.c_void_p()


class Group:
    """One process driving several GPUs (foundation_pt_group_*): the scene is uploaded to every member, `render` renders all tile
    partitions concurrently and gathers the frame into member 0."""

    def __init__(self, devices, width: int = 1920, height: int = 1080, seed: int = 1, flags: int = 0, background=(0.0, 0.0, 0.0), tile: int = 32,
                 comm_flags: int = 0, max_leaf_tris: int = 0):
        self._lib = load_library()
        cfg = Config(C.sizeof(Config), 0, width, height, seed, max_leaf_tris, flags, (C.c_float * 3)(*background), 0)
        devs = (C.c_int32 * len(devices))(*devices)
        self._g = C.c_void_p()
        st = self._lib.foundation_pt_group_create(C.byref(cfg), devs, len(devices), tile, comm_flags, None, C.byref(self._g))
        if st != OK:
            raise FoundationPtError(st, self._lib.foundation_pt_last_error(None).decode())
        self.members = [_Member(self._lib, self._lib.foundation_pt_group_context(self._g, i), width, height, d) for i, d in enumerate(devices)]

    def load(self, scene: Scene):
        return [m.load(scene) for m in self.members]

    def render(self, sample_begin: int, sample_count: int, max_bounces: int):
        st = self._lib.foundation_pt_group_render(self._g, sample_begin, sample_count, max_bounces)
        if st != OK:
            raise FoundationPtError(st, self._lib.foundation_pt_group_last_error(self._g).decode())

    def read_accum(self):
        return self.members[0].read_accum()

    def close(self):
        if getattr(self, "_g", None) and self._g.value:
            for m in self.members:
                m.close()
            self._lib.foundation_pt_group_destroy(self._g)
            self._g = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
