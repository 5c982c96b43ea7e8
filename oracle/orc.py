"""oracle/orc.py - TEST INFRASTRUCTURE ONLY: ctypes front end of oracle/libcrt_oracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The product (cudart_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libcrt_oracle.so")

RAY_DTYPE = np.dtype([("o", "<f4", 3), ("d", "<f4", 3), ("tmin", "<f4"), ("tmax", "<f4")])
HIT_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("tri", "<u4")])

RNG_RAND, RNG_COUNTER = 0, 1


class _Scene(C.Structure):
    _fields_ = [
        ("tris", C.c_void_p), ("n_tris", C.c_uint32),
        ("nodes", C.c_void_p), ("n_nodes", C.c_uint32),
        ("mat_floats", C.c_void_p), ("mat_indices", C.c_void_p), ("n_mats", C.c_uint32),
        ("light_indices", C.c_void_p), ("n_lights", C.c_uint32), ("lights_area", C.c_float),
        ("camera", C.c_void_p),
        ("tex_data", C.c_void_p), ("tex_dims", C.c_void_p), ("tex_offsets", C.c_void_p),
        ("n_tex", C.c_uint32), ("tex_texels", C.c_uint32),
    ]


class Counters(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("extend_rays", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("nodes_visited", C.c_uint64), ("tris_tested", C.c_uint64), ("max_stack", C.c_uint32)]

    def as_dict(self):
        d = {k: int(getattr(self, k)) for k, _ in self._fields_}
        rays = d["extend_rays"] + d["shadow_rays"]
        if d["samples"]:
            d["rays_per_sample"] = rays / d["samples"]
        if rays:
            d["nodes_per_ray"] = d["nodes_visited"] / rays
            d["tris_per_ray"] = d["tris_tested"] / rays
        return d


def build(force: bool = False) -> str:
    """Compile the C restatement (and, where /root/reference exists, oracle/_ref)."""
    srcs = [os.path.join(_HERE, "crt_oracle.c"), os.path.join(_HERE, "crt_oracle.h"),
            os.path.join(_HERE, "..", "include", "crt_rng.h"), os.path.join(_HERE, "..", "include", "crt_math.h")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_trace_closest.argtypes = [C.POINTER(_Scene), C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(Counters)]
        L.orc_trace_closest_unordered.argtypes = [C.POINTER(_Scene), C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_render_accum.argtypes = [C.POINTER(_Scene), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int,
                                       C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(Counters),
                                       C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.orc_resolve_rgba8.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.orc_counter_uniform.argtypes = [C.c_uint32] * 4
        L.orc_counter_uniform.restype = C.c_float
        L.orc_counter_bits.argtypes = [C.c_uint32] * 4
        L.orc_counter_bits.restype = C.c_uint32
        L.orc_sincos.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.orc_sincos.restype = None
        L.srand.argtypes = [C.c_uint]
        _lib = L
    return _lib


class OracleScene:
    """Keeps the numpy arrays of a FlatScene alive and exposes them as an orc_scene."""

    def __init__(self, fs):
        self.fs = fs
        keep = self._keep = {}
        for name in ("tris", "nodes", "mat_floats", "mat_indices", "light_indices", "camera", "tex_dims", "tex_offsets",
                     "tex_data"):
            keep[name] = np.ascontiguousarray(getattr(fs, name))
        p = lambda a: a.ctypes.data if a.size else None
        self.c = _Scene(p(keep["tris"]), fs.n_tris, p(keep["nodes"]), fs.n_nodes, p(keep["mat_floats"]),
                        p(keep["mat_indices"]), fs.n_mats, p(keep["light_indices"]), fs.n_lights,
                        float(fs.lights_area), p(keep["camera"]), p(keep["tex_data"]), p(keep["tex_dims"]),
                        p(keep["tex_offsets"]), fs.n_tex, int(keep["tex_data"].shape[0]))

    @property
    def width(self):
        return int(self.fs.camera["xpixels"][0])

    @property
    def height(self):
        return int(self.fs.camera["ypixels"][0])

    def trace_closest(self, rays, counters: Counters | None = None, unordered: bool = False):
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        hits = np.empty(rays.shape[0], HIT_DTYPE)
        if unordered:
            lib().orc_trace_closest_unordered(C.byref(self.c), rays.ctypes.data, rays.shape[0], hits.ctypes.data)
        else:
            lib().orc_trace_closest(C.byref(self.c), rays.ctypes.data, rays.shape[0], hits.ctypes.data,
                                    C.byref(counters) if counters is not None else None)
        return hits

    def render_accum(self, spp, rng=RNG_COUNTER, seed=1234, first_sample=0, threads=0, rows=None, accum=None,
                     counters: Counters | None = None, ray_log_cap=0):
        """Returns (accum[H,W,3] float32 radiance sums, ray_log or None)."""
        W, H = self.width, self.height
        if accum is None:
            accum = np.zeros((H, W, 3), np.float32)
        row0, row1 = rows if rows is not None else (0, H)
        if threads <= 0:
            threads = os.cpu_count() or 1
        log = np.empty(ray_log_cap, RAY_DTYPE) if ray_log_cap else None
        n_log = C.c_size_t(0)
        if rng == RNG_RAND:
            lib().srand(seed)
        lib().orc_render_accum(C.byref(self.c), W, H, first_sample, spp, rng, seed, threads, row0, row1,
                               accum.ctypes.data, C.byref(counters) if counters is not None else None,
                               log.ctypes.data if log is not None else None, ray_log_cap, C.byref(n_log))
        if log is not None:
            log = log[: min(n_log.value, ray_log_cap)]
        return accum, log


def resolve_rgba8(accum, spp):
    accum = np.ascontiguousarray(accum, np.float32)
    n = accum.size // 3
    out = np.empty((n, 4), np.uint8)
    lib().orc_resolve_rgba8(accum.ctypes.data, n, spp, out.ctypes.data)
    return out.reshape(accum.shape[:-1] + (4,))


def counter_uniform(seed, pixel, sample, dim):
    return float(lib().orc_counter_uniform(seed, pixel, sample, dim))


def counter_bits(seed, pixel, sample, dim):
    return int(lib().orc_counter_bits(seed, pixel, sample, dim))


def sincos(a):
    """(sin, cos) of the product's own function (include/crt_math.h)."""
    s, c = C.c_float(0), C.c_float(0)
    lib().orc_sincos(float(a), C.byref(s), C.byref(c))
    return float(s.value), float(c.value)
