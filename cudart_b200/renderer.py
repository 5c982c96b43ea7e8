"""Host-side mirror of the reference's ``Renderer`` seam (``src/include/renderer.h:78-100``) on top of the
C ABI: same method names, argument meaning and call order as ``TerminalLauncher::render`` uses them
(``terminal_launcher.cpp:10-16``): ``getSamples``, ``getImgBytesPointer``, ``renderOneSamplePerPixel`` x spp,
``copyImageBytes``; ``saveToImage`` then reads ``h_imgPtr`` (``launcher.cpp:13-21``)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .scenefile import HIT_DTYPE, RAY_DTYPE, FlatScene


def scene_desc(fs: FlatScene):
    """(SceneDesc, keepalive) for a FlatScene."""
    keep = {k: np.ascontiguousarray(getattr(fs, k)) for k in
            ("tris", "nodes", "mat_floats", "mat_indices", "light_indices", "camera", "tex_dims", "tex_offsets", "tex_data")}
    p = lambda a: a.ctypes.data if a.size else None
    d = capi.SceneDesc(p(keep["tris"]), fs.n_tris, p(keep["nodes"]), fs.n_nodes, p(keep["mat_floats"]),
                       p(keep["mat_indices"]), fs.n_mats, p(keep["light_indices"]), fs.n_lights, float(fs.lights_area),
                       p(keep["camera"]), p(keep["tex_data"]), p(keep["tex_dims"]), p(keep["tex_offsets"]), fs.n_tex,
                       int(keep["tex_data"].shape[0]))
    return d, keep


class B200Renderer:
    """Drop-in third renderer next to SequentialRenderer / ParallelRenderer.

    ``B200Renderer(scene, width, height, samples)`` matches ``ParallelRenderer(Scene*, w, h, spp)``
    (``main.cpp:163``); extra keyword options map to ``crt_options``.
    """

    useCuda = True

    def __init__(self, scene: FlatScene, width: int, height: int, samples: int = 10, *, device: int = -1, rank: int = 0,
                 world: int = 1, tile_rows: int = 8, seed: int = 1234, spp_per_wave: int = 0, exact_visit: bool = False,
                 profile: bool = False, stats: bool = False, reference_topology: bool = False, smem_nodes: bool | None = None,
                 tight_bvh: bool | None = None, dynamic_pad: bool | None = None, stream: int | None = None,
                 waves_in_flight: int = 0):
        self._lib = capi.lib()
        self.width, self.height, self.samples = int(width), int(height), int(samples)
        self.rank, self.world = rank, world
        desc, self._keep = scene_desc(scene)
        flags = (capi.CRT_FLAG_EXACT_VISIT if exact_visit else 0) | (capi.CRT_FLAG_PROFILE if profile else 0) \
            | (capi.CRT_FLAG_STATS if stats else 0) | (capi.CRT_FLAG_REFERENCE_TOPOLOGY if reference_topology else 0) \
            | (0 if smem_nodes is None else (capi.CRT_FLAG_SMEM_TOP_LEVELS if smem_nodes else capi.CRT_FLAG_NO_SMEM_NODES)) \
            | (0 if tight_bvh is None else (capi.CRT_FLAG_TIGHT_BVH if tight_bvh else capi.CRT_FLAG_NO_TIGHT_BVH)) \
            | (0 if dynamic_pad is None else (capi.CRT_FLAG_DYNAMIC_PAD if dynamic_pad else capi.CRT_FLAG_STATIC_PAD))
        opt = capi.Options(C.sizeof(capi.Options), device, rank, world, tile_rows, seed, spp_per_wave, flags, stream,
                           waves_in_flight)
        h = C.c_void_p()
        capi.check(self._lib.crt_create(C.byref(desc), self.width, self.height, C.byref(opt), C.byref(h)))
        self._h = h
        self.h_imgPtr = np.zeros((self.height, self.width, 4), np.uint8)     # Renderer::h_imgPtr (renderer.cu:80-85)

    # ---- the reference's five-call contract -------------------------------------------------------
    def renderOneSamplePerPixel(self, p_img=None) -> None:
        """One more sample per pixel (``renderer.h:84``).  The library queues single samples and launches them in whole
        wavefront batches; ``getSamplesRendered`` advances by one per call all the same."""
        capi.check(self._lib.crt_render(self._h, 1))

    def copyImageBytes(self, p_img=None) -> None:
        capi.check(self._lib.crt_resolve_rgba8(self._h, self.h_imgPtr.ctypes.data))

    def getImgBytesPointer(self):
        return self.h_imgPtr

    def createMaterialsData(self) -> None:
        """Material tables are packed and uploaded by crt_create (parallel_renderer.cu:252-279 equivalent)."""

    def getWidth(self) -> int:
        return self.width

    def getHeight(self) -> int:
        return self.height

    def getSamples(self) -> int:
        return self.samples

    def getSamplesRendered(self) -> int:
        return int(self._lib.crt_samples_rendered(self._h))

    def spp_per_wave(self) -> int:
        """Samples per pixel the library traces in one wavefront batch."""
        return int(self._lib.crt_spp_per_wave(self._h))

    # ---- batch API ------------------------------------------------------------------------------------
    def render(self, spp: int) -> None:
        """`spp` samples per pixel in as few wavefront batches as possible (asynchronous)."""
        capi.check(self._lib.crt_render(self._h, int(spp)))

    def flush(self) -> None:
        """Launch samples still queued by ``renderOneSamplePerPixel`` (asynchronous)."""
        capi.check(self._lib.crt_flush(self._h))

    def sync(self) -> None:
        capi.check(self._lib.crt_sync(self._h))

    def reset(self) -> None:
        capi.check(self._lib.crt_reset(self._h))

    def update_scene(self, scene: FlatScene) -> None:
        """Upload another scene (same image size) into the live renderer and start a fresh accumulation."""
        desc, self._keep = scene_desc(scene)
        capi.check(self._lib.crt_update_scene(self._h, C.byref(desc)))

    def image_f32(self) -> np.ndarray:
        out = np.empty((self.height, self.width, 3), np.float32)
        capi.check(self._lib.crt_resolve_f32(self._h, out.ctypes.data))
        return out

    def accum(self) -> np.ndarray:
        out = np.empty((self.height, self.width, 3), np.float32)
        capi.check(self._lib.crt_read_accum(self._h, out.ctypes.data))
        return out

    def counters(self) -> dict:
        c = capi.Counters()
        capi.check(self._lib.crt_get_counters(self._h, C.byref(c)))
        return c.as_dict()

    def trace_closest(self, rays: np.ndarray, exact_visit: bool = False, stats: bool = False) -> np.ndarray:
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        hits = np.empty(rays.shape[0], HIT_DTYPE)
        capi.check(self._lib.crt_trace_closest(self._h, rays.ctypes.data, rays.shape[0], hits.ctypes.data,
                                               (capi.CRT_FLAG_EXACT_VISIT if exact_visit else 0) |
                                               (capi.CRT_FLAG_STATS if stats else 0)))
        return hits

    def trace_closest_device(self, d_rays: int, n: int, d_hits: int, exact_visit: bool = False) -> float:
        """Closest hit over device arrays; returns the kernel's CUDA-event time in ms."""
        ms = C.c_float(0)
        capi.check(self._lib.crt_trace_closest_device(self._h, d_rays, n, d_hits,
                                                      capi.CRT_FLAG_EXACT_VISIT if exact_visit else 0, C.byref(ms)))
        return float(ms.value)

    def tight_info(self):
        """(wide nodes of the tight per-triangle BVH the render path walks, or 0 when off; individually boxed triangles)."""
        a, b = C.c_uint32(0), C.c_uint32(0)
        capi.check(self._lib.crt_tight_info(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def lane_stats(self):
        """stats=True only: dict of warp-level loop iterations and lane occupancy of the traversal kernels."""
        out = (C.c_uint64 * 8)()
        capi.check(self._lib.crt_lane_stats(self._h, C.byref(out)))
        keys = ("node_iters", "node_lanes", "node_idle", "leaf_iters", "leaf_lanes", "leaf_idle", "refills", "refill_rays")
        return dict(zip(keys, (int(v) for v in out)))

    def scene_info(self):
        """(host rebuilds of the derived scene arrays, scene uploads, host->device bytes of the last upload)."""
        a, b, n = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        capi.check(self._lib.crt_scene_info(self._h, C.byref(a), C.byref(b), C.byref(n)))
        return int(a.value), int(b.value), int(n.value)

    def tight_pad_mode(self) -> int:
        """0 = the render path does not walk the tight BVH, 1 = pad baked into its boxes, 2 = padded per node visit."""
        return int(self._lib.crt_tight_pad_mode(self._h))

    def accum_device_ptr(self):
        p, n = C.c_void_p(), C.c_size_t()
        capi.check(self._lib.crt_accum_device_ptr(self._h, C.byref(p), C.byref(n)))
        return int(p.value), int(n.value)

    def resolve_gathered(self, d_gathered: int, want_rgba8: bool = True, want_f32: bool = False):
        rgba = self.h_imgPtr if want_rgba8 else None          # in place: the caller may have made it page-locked
        f32 = np.empty((self.height, self.width, 3), np.float32) if want_f32 else None
        capi.check(self._lib.crt_resolve_gathered(self._h, d_gathered, rgba.ctypes.data if want_rgba8 else None,
                                                  f32.ctypes.data if want_f32 else None))
        return rgba, f32

    def resolve_local_rgba8_device(self):
        """(device pointer, bytes) of this rank's owned pixels as RGBA8, compact tile-major (asynchronous)."""
        p, n = C.c_void_p(), C.c_size_t()
        capi.check(self._lib.crt_resolve_local_rgba8_device(self._h, C.byref(p), C.byref(n)))
        return int(p.value), int(n.value)

    def assemble_rgba8(self, d_gathered: int):
        """De-interleave the rank-major RGBA8 buffers of all ranks into ``h_imgPtr`` (blocking)."""
        capi.check(self._lib.crt_assemble_rgba8(self._h, d_gathered, self.h_imgPtr.ctypes.data))
        return self.h_imgPtr

    def stream_wait_ctx(self, stream: int) -> None:
        """`stream` waits (on the device) for everything this renderer has been asked to do so far."""
        capi.check(self._lib.crt_stream_wait_ctx(self._h, stream))

    def ctx_wait_stream(self, stream: int) -> None:
        """The renderer's stream waits (on the device) for everything enqueued on `stream` so far."""
        capi.check(self._lib.crt_ctx_wait_stream(self._h, stream))

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.crt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
