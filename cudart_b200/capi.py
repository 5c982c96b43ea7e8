"""ctypes binding of libcrt_b200.so (include/crt_abi.h).  No fallback: importing the symbols fails loudly
when the library has not been built (``python -c "import __graft_entry__ as g; g.build()"``)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CRT_LIB") or os.path.join(_HERE, "libcrt_b200.so")   # CRT_LIB: tuning builds only

CRT_FLAG_EXACT_VISIT = 1
CRT_FLAG_PROFILE = 2
CRT_FLAG_STATS = 4
CRT_FLAG_REFERENCE_TOPOLOGY = 8
CRT_FLAG_NO_SMEM_NODES = 16
CRT_FLAG_SMEM_TOP_LEVELS = 32
CRT_FLAG_TIGHT_BVH = 64
CRT_FLAG_NO_TIGHT_BVH = 128
CRT_FLAG_DYNAMIC_PAD = 256
CRT_FLAG_STATIC_PAD = 512
CRT_MISS = 0xFFFFFFFF
STAGES = ("raygen", "extend", "shade", "shadow", "accumulate")


class SceneDesc(C.Structure):
    _fields_ = [
        ("tris", C.c_void_p), ("n_tris", C.c_uint32),
        ("nodes", C.c_void_p), ("n_nodes", C.c_uint32),
        ("mat_floats", C.c_void_p), ("mat_indices", C.c_void_p), ("n_mats", C.c_uint32),
        ("light_indices", C.c_void_p), ("n_lights", C.c_uint32), ("lights_area", C.c_float),
        ("camera", C.c_void_p),
        ("tex_data", C.c_void_p), ("tex_dims", C.c_void_p), ("tex_offsets", C.c_void_p),
        ("n_tex", C.c_uint32), ("tex_texels", C.c_uint32),
    ]


class Options(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("device", C.c_int32), ("rank", C.c_uint32), ("world", C.c_uint32),
        ("tile_rows", C.c_uint32), ("seed", C.c_uint32), ("spp_per_wave", C.c_uint32), ("flags", C.c_uint32),
        ("stream", C.c_void_p), ("waves_in_flight", C.c_uint32),
    ]


class Counters(C.Structure):
    _fields_ = [
        ("samples", C.c_uint64), ("extend_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("node_visits", C.c_uint64),
        ("tri_tests", C.c_uint64), ("kernel_launches", C.c_uint64), ("render_ms", C.c_double),
        ("stage_ms", C.c_double * 5),
    ]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "stage_ms"}
        d["stage_ms"] = dict(zip(STAGES, list(self.stage_ms)))
        return d


# every symbol include/crt_abi.h declares
SYMBOLS = {
    "crt_abi_version": (C.c_int, []),
    "crt_last_error": (C.c_char_p, []),
    "crt_create": (C.c_int, [C.POINTER(SceneDesc), C.c_uint32, C.c_uint32, C.POINTER(Options), C.POINTER(C.c_void_p)]),
    "crt_destroy": (None, [C.c_void_p]),
    "crt_render": (C.c_int, [C.c_void_p, C.c_uint32]),
    "crt_flush": (C.c_int, [C.c_void_p]),
    "crt_sync": (C.c_int, [C.c_void_p]),
    "crt_samples_rendered": (C.c_uint32, [C.c_void_p]),
    "crt_spp_per_wave": (C.c_uint32, [C.c_void_p]),
    "crt_resolve_rgba8": (C.c_int, [C.c_void_p, C.c_void_p]),
    "crt_resolve_f32": (C.c_int, [C.c_void_p, C.c_void_p]),
    "crt_read_accum": (C.c_int, [C.c_void_p, C.c_void_p]),
    "crt_reset": (C.c_int, [C.c_void_p]),
    "crt_update_scene": (C.c_int, [C.c_void_p, C.POINTER(SceneDesc)]),
    "crt_get_counters": (C.c_int, [C.c_void_p, C.POINTER(Counters)]),
    "crt_trace_closest": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32]),
    "crt_trace_closest_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32, C.POINTER(C.c_float)]),
    "crt_accum_device_ptr": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "crt_resolve_gathered": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "crt_resolve_local_rgba8_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "crt_assemble_rgba8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "crt_stream_wait_ctx": (C.c_int, [C.c_void_p, C.c_void_p]),
    "crt_ctx_wait_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "crt_build_bvh": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32)]),
    "crt_build_lights": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                   C.POINTER(C.c_uint32), C.POINTER(C.c_float)]),
    "crt_wide_bvh": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]),
    "crt_tight_bvh": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_void_p]),
    "crt_tight_pad_mode": (C.c_int, [C.c_void_p]),
    "crt_tight_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "crt_derived_fingerprint": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64 * 2)]),
    "crt_fingerprint": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64 * 2)]),
    "crt_lane_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64 * 8)]),
    "crt_scene_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
}


class CrtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("crt error %d: %s" % (code, msg))
        self.code = code


def build(verbose: bool = False) -> str:
    """Compile libcrt_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")], check=True,
                   stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing - build it with `make -C cudart_b200/csrc` (there is no fallback path)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.crt_abi_version() != 1:
            raise ImportError("libcrt_b200.so ABI version mismatch")
        _lib = L
    return _lib


def check(code: int) -> None:
    if code != 0:
        raise CrtError(code, lib().crt_last_error().decode("utf-8", "replace"))
