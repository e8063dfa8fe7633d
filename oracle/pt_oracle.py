"""ctypes binding of the CPU oracle (oracle/pt_oracle.cpp).  TEST INFRASTRUCTURE ONLY — parity unpinned (the
reference, mos9527/Foundation, ships no ray tracer).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module; the product never does."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

RAY_DTYPE = np.dtype([("origin", "<f4", (3,)), ("tmin", "<f4"), ("direction", "<f4", (3,)), ("tmax", "<f4")])
HIT_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("prim", "<u4")])
NODE_DTYPE = np.dtype([("p", "<f4", (3,)), ("e", "u1", (3,)), ("imask", "u1"), ("child_base", "<u4"), ("tri_base", "<u4"), ("meta", "u1", (8,)),
                       ("qlo", "u1", (3, 8)), ("qhi", "u1", (3, 8))])
TRI_DTYPE = np.dtype([("v0", "<f4", (3,)), ("prim", "<u4"), ("e1", "<f4", (3,)), ("mat", "<u4"), ("e2", "<f4", (3,)), ("pad", "<u4")])
assert NODE_DTYPE.itemsize == 80 and TRI_DTYPE.itemsize == 48


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libpt_oracle.so")
    src = os.path.join(_HERE, "pt_oracle.cpp")
    hdr_dir = os.path.join(_HERE, "..", "foundation_b200", "csrc")
    newest = max([os.path.getmtime(src)] + [os.path.getmtime(os.path.join(hdr_dir, h)) for h in
                                            ("pt_math.h", "pt_layout.h", "pt_shading.h", "pt_host_shared.h") if os.path.exists(os.path.join(hdr_dir, h))])
    if force or not os.path.exists(so) or os.path.getmtime(so) < newest:
        import fcntl
        with open(os.path.join(_HERE, ".build.lock"), "w") as lock:      # concurrent test workers / ranks: one builder at a time
            fcntl.flock(lock, fcntl.LOCK_EX)
            try:
                if force or not os.path.exists(so) or os.path.getmtime(so) < newest:
                    subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
            finally:
                fcntl.flock(lock, fcntl.LOCK_UN)
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libpt_oracle.so")
        alt = os.environ.get("PT_ORACLE_LIB")          # e.g. the -fsanitize=address,undefined build of `make -C oracle asan`
        if alt:
            so = alt
        elif not os.path.exists(so) or os.environ.get("PT_ORACLE_REBUILD"):
            build()
        else:
            try:
                build()
            except Exception:  # no compiler on the box: use the shipped binary
                pass
        L = C.CDLL(so)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_uint32]
        for name in ("orc_destroy", "orc_commit"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.orc_materials_set.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_mesh_add.restype = C.c_uint32
        L.orc_mesh_add.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
        L.orc_instances_set.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_mesh_attributes_set.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_texture_add.restype = C.c_uint32; L.orc_texture_add.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.orc_material_textures_set.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_texture_sample.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_void_p]
        L.orc_blas_counts.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.orc_blas_get.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_tlas_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_tlas_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_scene_info.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.orc_trace_closest.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_trace_any.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_camera_set.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_camera_get.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_render.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p,
                                 C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_bsdf_eval.argtypes = [C.c_void_p] * 5
        L.orc_bsdf_sample.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p]
        L.orc_bsdf_sample.restype = C.c_int
        L.orc_sincos2pi.argtypes = [C.c_float, C.c_void_p, C.c_void_p]
        L.orc_pcg.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        L.orc_pcg_raw.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]
        L.orc_morton.restype = C.c_uint64
        L.orc_morton.argtypes = [C.c_void_p] * 3
        L.orc_sobol02.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.orc_sobol02_padded.argtypes = [C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p]
        L.orc_path_dim_key.restype = C.c_uint64; L.orc_path_dim_key.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
        L.orc_hw_threads.restype = C.c_int
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class OracleScene:
    """Same call sequence as foundation_b200.pt.PathTracer so parity tests read alike on both sides."""

    def __init__(self, scene=None, max_leaf: int = 0):
        self._L = lib()
        self._h = C.c_void_p(self._L.orc_create(max_leaf))
        self.flat = True
        if scene is not None:
            self.load(scene)

    def __del__(self):
        try:
            if self._h:
                self._L.orc_destroy(self._h); self._h = None
        except Exception:
            pass

    def load(self, scene):
        m = np.ascontiguousarray(scene.materials, np.float32)
        self._L.orc_materials_set(self._h, _p(m), m.shape[0])
        for mesh in scene.meshes:
            pos = np.ascontiguousarray(mesh.positions, np.float32); idx = np.ascontiguousarray(mesh.indices, np.uint32)
            mat = np.ascontiguousarray(mesh.material_ids, np.uint32)
            mid = self._L.orc_mesh_add(self._h, _p(pos), pos.shape[0], _p(idx), idx.shape[0], _p(mat))
            uv, col = getattr(mesh, "uv", None), getattr(mesh, "colors", None)
            if uv is not None or col is not None:
                uv = None if uv is None else np.ascontiguousarray(uv, np.float32); col = None if col is None else np.ascontiguousarray(col, np.float32)
                self._L.orc_mesh_attributes_set(self._h, mid, _p(uv), _p(col), _p(idx))
        for tex in getattr(scene, "textures", None) or []:
            t = np.ascontiguousarray(tex, np.uint8)
            self._L.orc_texture_add(self._h, _p(t), t.shape[1], t.shape[0])
        mt = getattr(scene, "material_textures", None)
        if mt is not None:
            mt = np.ascontiguousarray(mt, np.uint32)
            self._L.orc_material_textures_set(self._h, _p(mt), mt.shape[0])
        if scene.instances is not None:
            ids = np.ascontiguousarray(scene.instances["mesh_id"], np.uint32); xf = np.ascontiguousarray(scene.instances["transform"], np.float32)
            self._L.orc_instances_set(self._h, _p(ids), _p(xf), ids.shape[0])
            self.flat = False
        elif len(scene.meshes) > 1:
            # same rule as foundation_pt_scene_commit: several meshes and no instance list = every mesh once with the identity transform
            ids = np.arange(len(scene.meshes), dtype=np.uint32)
            xf = np.ascontiguousarray(np.tile(np.asarray([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32), (len(scene.meshes), 1)))
            self._L.orc_instances_set(self._h, _p(ids), _p(xf), ids.shape[0])
            self.flat = False
        self._L.orc_commit(self._h)
        if scene.view is not None:
            self.camera_set(scene.view, scene.proj)
        self.scene = scene

    def camera_set(self, view, proj):
        v = np.ascontiguousarray(view, np.float32); p = np.ascontiguousarray(proj, np.float32)
        if self._L.orc_camera_set(self._h, _p(v), _p(p)) != 0:
            raise ValueError("singular camera")

    def camera_get(self):
        out = np.zeros(12, np.float32); self._L.orc_camera_get(self._h, _p(out)); return out

    def info(self):
        lo = np.zeros(3, np.float32); hi = np.zeros(3, np.float32)
        eps = C.c_float(); nl = C.c_uint32(); la = C.c_float()
        self._L.orc_scene_info(self._h, _p(lo), _p(hi), C.byref(eps), C.byref(nl), C.byref(la))
        return dict(lo=lo, hi=hi, ray_eps=eps.value, num_lights=nl.value, light_area=la.value)

    def blas(self, mesh: int = 0):
        nn = C.c_uint64(); nt = C.c_uint64()
        self._L.orc_blas_counts(self._h, mesh, C.byref(nn), C.byref(nt))
        nodes = np.zeros(nn.value, NODE_DTYPE); tris = np.zeros(nt.value, TRI_DTYPE); order = np.zeros(nt.value, np.uint32)
        self._L.orc_blas_get(self._h, mesh, _p(nodes), _p(tris), _p(order))
        return nodes, tris, order

    def tlas(self):
        nn = C.c_uint64(); ni = C.c_uint64()
        self._L.orc_tlas_counts(self._h, C.byref(nn), C.byref(ni))
        nodes = np.zeros(nn.value, NODE_DTYPE); order = np.zeros(ni.value, np.uint32); rec = np.zeros((ni.value, 28), np.uint32)
        self._L.orc_tlas_get(self._h, _p(nodes), _p(order), _p(rec))
        return nodes, order, rec

    def trace_closest(self, rays, brute: bool = False, threads: int = 0, counters: bool = False):
        rays = np.ascontiguousarray(rays); assert rays.dtype.itemsize == 32
        n = rays.shape[0]
        hits = np.zeros(n, HIT_DTYPE); inst = np.zeros(n, np.uint32); cnt = np.zeros(3, np.uint64)
        self._L.orc_trace_closest(self._h, _p(rays), n, _p(hits), _p(inst), 1 if brute else 0, threads, _p(cnt))
        return (hits, inst, cnt) if counters else (hits, inst)

    def trace_any(self, rays, brute: bool = False, threads: int = 0, counters: bool = False):
        rays = np.ascontiguousarray(rays); n = rays.shape[0]
        occ = np.zeros(n, np.uint8); cnt = np.zeros(3, np.uint64)
        self._L.orc_trace_any(self._h, _p(rays), n, _p(occ), 1 if brute else 0, threads, _p(cnt))
        return (occ, cnt) if counters else occ

    def render(self, width, height, seed, sample_begin, sample_count, max_bounces, flags=0, background=(0, 0, 0), rank=0, count=1, tile=32,
               accum=None, threads=0, brute=False, ray_counts=None):
        if accum is None:
            accum = np.zeros((height, width, 4), np.float32)
        bg = np.asarray(background, np.float32)
        rc = np.zeros(2, np.uint64)
        self._L.orc_render(self._h, width, height, seed, sample_begin, sample_count, max_bounces, flags, _p(bg), rank, count, tile, _p(accum),
                           threads, 1 if brute else 0, _p(rc))
        if ray_counts is not None:
            ray_counts[:] = rc
        return accum


def bsdf_eval(mat8, wo, wi):
    L = lib(); m = np.asarray(mat8, np.float32); a = np.asarray(wo, np.float32); b = np.asarray(wi, np.float32)
    f = np.zeros(3, np.float32); pdf = C.c_float()
    L.orc_bsdf_eval(_p(m), _p(a), _p(b), _p(f), C.byref(pdf))
    return f, pdf.value


def bsdf_sample(mat8, wo, ul, u1, u2):
    L = lib(); m = np.asarray(mat8, np.float32); a = np.asarray(wo, np.float32); wi = np.zeros(3, np.float32)
    ok = L.orc_bsdf_sample(_p(m), _p(a), ul, u1, u2, _p(wi))
    return bool(ok), wi


def sincos2pi(u):
    L = lib(); s = C.c_float(); c = C.c_float(); L.orc_sincos2pi(u, C.byref(s), C.byref(c)); return s.value, c.value


def pcg(seed, pixel, sample, n):
    out = np.zeros(n, np.uint32); lib().orc_pcg(seed, pixel, sample, n, _p(out)); return out


def pcg_raw(initstate, initseq, n):
    out = np.zeros(n, np.uint32); lib().orc_pcg_raw(initstate, initseq, n, _p(out)); return out


def sobol02(sample, k0=0, k1=0):
    x = C.c_float(); y = C.c_float(); lib().orc_sobol02(sample, k0, k1, C.byref(x), C.byref(y)); return x.value, y.value


def texture_sample(rgba8, u, v):
    t = np.ascontiguousarray(rgba8, np.uint8); out = np.zeros(3, np.float32)
    lib().orc_texture_sample(_p(t), t.shape[1], t.shape[0], u, v, _p(out))
    return out


def hw_threads():
    return lib().orc_hw_threads()


def sobol02_padded(sample, key):
    x = C.c_float(); y = C.c_float(); lib().orc_sobol02_padded(sample, key, C.byref(x), C.byref(y)); return x.value, y.value


def path_dim_key(seed, pixel, bounce, which):
    return int(lib().orc_path_dim_key(seed, pixel, bounce, which))
