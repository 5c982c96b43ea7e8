"""Flattened scene arrays in the reference's own POD layouts, and the CRTS container that holds them.

The arrays are exactly what the reference's ``Scene`` getters hand to its renderers
(``/root/reference/src/include/scene.h:26-49``): ``Triangle[]`` in BVH order (128 B each,
``geometry.h:30-51``), ``LinearBVHNode[]`` (32 B, ``bvh.h:18-27``), the material tables as packed by
``SequentialRenderer::createMaterialsData`` (``sequential_renderer.cu:48-64``), the light index list,
``Camera`` (72 B, ``camera.h:9-22``) and the flattened, padded texture store
(``texture_store.cpp:37-71``).  Feeding both the CPU oracle and the B200 path the *same bytes* makes
loader quirks cancel out of every parity test.

CRTS file, little endian::

    u32[16] header: "CRTS", version=1, n_tris, n_nodes, n_mats, n_lights, n_tex, tex_texels, W, H,
                    f32 lights_area, 0...
    camera[72] tris[n_tris*128] nodes[n_nodes*32] mat_floats[n_mats*12 f32] mat_indices[n_mats*2 i32]
    light_indices[n_lights u32] tex_dims[2*n_tex u32] tex_offsets[(n_tex+1) u32] tex_data[tex_texels*3 f32]
"""
from __future__ import annotations

import dataclasses
import gzip
import io
import os

import numpy as np

TRIANGLE_DTYPE = np.dtype(
    {
        "names": ["triId", "materialId", "v1", "e1", "e2", "n1", "n2", "n3", "uv1", "uv2", "uv3", "area"],
        "formats": ["<u4", "<u4", ("<f4", 3), ("<f4", 3), ("<f4", 3), ("<f4", 3), ("<f4", 3), ("<f4", 3),
                    ("<f4", 2), ("<f4", 2), ("<f4", 2), "<f4"],
        "offsets": [0, 4, 8, 20, 32, 44, 56, 68, 80, 88, 96, 104],
        "itemsize": 128,
    }
)
NODE_DTYPE = np.dtype(
    {
        "names": ["min", "max", "offset", "numTriangles", "axis"],
        "formats": [("<f4", 3), ("<f4", 3), "<i4", "<i2", "<i2"],
        "offsets": [0, 12, 24, 28, 30],
        "itemsize": 32,
    }
)
CAMERA_DTYPE = np.dtype(
    {
        "names": ["eye", "dir", "up", "right", "fov", "aspect", "apertureWidth", "focusDistance", "xpixels", "ypixels"],
        "formats": [("<f4", 3), ("<f4", 3), ("<f4", 3), ("<f4", 3), "<f4", "<f4", "<f4", "<f4", "<i4", "<i4"],
        "offsets": [0, 12, 24, 36, 48, 52, 56, 60, 64, 68],
        "itemsize": 72,
    }
)
RAY_DTYPE = np.dtype([("o", "<f4", 3), ("d", "<f4", 3), ("tmin", "<f4"), ("tmax", "<f4")])
HIT_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("tri", "<u4")])
MISS = 0xFFFFFFFF

# refl_t, material.h:8-15
DIFFUSE, SPECULAR, DIFFSPEC, MICROFACET, REFRACTIVE, EMISSIVE = range(6)


@dataclasses.dataclass
class FlatScene:
    """One scene as flat numpy arrays (reference layouts)."""

    tris: np.ndarray            # TRIANGLE_DTYPE[n_tris]
    nodes: np.ndarray           # NODE_DTYPE[n_nodes]
    mat_floats: np.ndarray      # f32[n_mats, 4, 3]  kd, ka, ks, (ns, ni, diffuseCoefficient)
    mat_indices: np.ndarray     # i32[n_mats, 2]     (bsdf, texKdIdx)
    light_indices: np.ndarray   # u32[n_lights]
    lights_area: float
    camera: np.ndarray          # CAMERA_DTYPE scalar array, shape (1,)
    tex_dims: np.ndarray        # u32[2*n_tex]
    tex_offsets: np.ndarray     # u32[n_tex+1]
    tex_data: np.ndarray        # f32[tex_texels, 3]
    width: int = 0
    height: int = 0

    @property
    def n_tris(self) -> int:
        return int(self.tris.shape[0])

    @property
    def n_nodes(self) -> int:
        return int(self.nodes.shape[0])

    @property
    def n_mats(self) -> int:
        return int(self.mat_floats.shape[0])

    @property
    def n_lights(self) -> int:
        return int(self.light_indices.shape[0])

    @property
    def n_tex(self) -> int:
        return int(self.tex_dims.shape[0] // 2)

    def with_resolution(self, width: int, height: int) -> "FlatScene":
        """Copy with the camera re-targeted to another image size (json_loader.cpp:33-34,44)."""
        cam = self.camera.copy()
        cam["xpixels"] = width
        cam["ypixels"] = height
        cam["aspect"] = np.float32(width) / np.float32(height)
        return dataclasses.replace(self, camera=cam, width=width, height=height)


def _open(path, mode):
    return gzip.open(path, mode) if str(path).endswith(".gz") else open(path, mode)


def load_crts(path) -> FlatScene:
    with _open(path, "rb") as f:
        buf = f.read()
    hdr = np.frombuffer(buf, "<u4", 16)
    if buf[:4] != b"CRTS" or hdr[1] != 1:
        raise ValueError("%s: not a CRTS v1 file" % path)
    n_tris, n_nodes, n_mats, n_lights, n_tex, tex_texels, W, H = (int(v) for v in hdr[2:10])
    lights_area = float(np.frombuffer(buf, "<f4", 1, 40)[0])
    off = 64

    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(buf, dtype, count, off).copy()
        off += a.nbytes
        return a

    camera = take(CAMERA_DTYPE, 1)
    tris = take(TRIANGLE_DTYPE, n_tris)
    nodes = take(NODE_DTYPE, n_nodes)
    mat_floats = take("<f4", n_mats * 12).reshape(n_mats, 4, 3)
    mat_indices = take("<i4", n_mats * 2).reshape(n_mats, 2)
    light_indices = take("<u4", n_lights)
    tex_dims = take("<u4", 2 * n_tex)
    tex_offsets = take("<u4", n_tex + 1)
    tex_data = take("<f4", tex_texels * 3).reshape(tex_texels, 3)
    if off != len(buf):
        raise ValueError("%s: %d trailing bytes" % (path, len(buf) - off))
    return FlatScene(tris, nodes, mat_floats, mat_indices, light_indices, lights_area, camera, tex_dims, tex_offsets,
                     tex_data, W, H)


def save_crts(path, s: FlatScene) -> None:
    hdr = np.zeros(16, "<u4")
    hdr[0] = np.frombuffer(b"CRTS", "<u4")[0]
    hdr[1] = 1
    hdr[2:10] = [s.n_tris, s.n_nodes, s.n_mats, s.n_lights, s.n_tex, s.tex_data.shape[0], s.width, s.height]
    hdr[10] = np.frombuffer(np.float32(s.lights_area).tobytes(), "<u4")[0]
    out = io.BytesIO()
    out.write(hdr.tobytes())
    for a, dt in ((s.camera, CAMERA_DTYPE), (s.tris, TRIANGLE_DTYPE), (s.nodes, NODE_DTYPE), (s.mat_floats, "<f4"),
                  (s.mat_indices, "<i4"), (s.light_indices, "<u4"), (s.tex_dims, "<u4"), (s.tex_offsets, "<u4"),
                  (s.tex_data, "<f4")):
        out.write(np.ascontiguousarray(a, dtype=dt).tobytes())
    os.makedirs(os.path.dirname(os.path.abspath(path)) or ".", exist_ok=True)
    with _open(path, "wb") as f:
        f.write(out.getvalue())
