"""Image-space sharding across the GPUs of one box: interleaved row tiles (tile k -> rank k % world), one
process per GPU, ONE exchange at read-out: every rank resolves its own tiles to RGBA8 on its GPU and rank 0
collects them over NCCL/NVLink and de-interleaves (``gather_rgba8``, 4 bytes per pixel); the all-gather of the
float4 accumulators (``gather_framebuffer``) stays for float parity checks.  The reference has no multi-GPU path
at all (``main.cpp:155-162`` picks one device); the row-tile math here mirrors ``DTiling`` in ``csrc/kernels.cuh``.

The path shards with no data-path exchange: pixels are independent and the RNG is keyed by the GLOBAL
pixel index, so the image does not depend on ``world``.
"""
from __future__ import annotations

import numpy as np


def owned_rows(height: int, rank: int, world: int, tile_rows: int = 8) -> np.ndarray:
    """Global row indices rank `rank` renders, in its local (compact) order."""
    rows = []
    n_tiles = -(-height // tile_rows)
    for t in range(rank, n_tiles, world):
        rows.extend(range(t * tile_rows, min((t + 1) * tile_rows, height)))
    return np.asarray(rows, np.int64)


def padded_rows(height: int, world: int, tile_rows: int = 8) -> int:
    """Rows per rank in the gather buffer (identical on every rank)."""
    n_tiles = -(-height // tile_rows)
    return -(-n_tiles // world) * tile_rows


def deinterleave(gathered: np.ndarray, width: int, height: int, world: int, tile_rows: int = 8) -> np.ndarray:
    """gathered: [world, padded_rows*width, C] rank-major compact buffers -> [height, width, C] image (host restatement
    of the indexing in k_resolve, used by the CPU tests of the multi-process logic)."""
    pr = padded_rows(height, world, tile_rows)
    g = np.asarray(gathered).reshape(world, pr, width, -1)
    img = np.zeros((height, width, g.shape[-1]), g.dtype)
    for r in range(world):
        rows = owned_rows(height, r, world, tile_rows)
        img[rows] = g[r, : len(rows)]
    return img


def as_torch(ptr: int, shape, dtype="<f4"):
    """Zero-copy torch view of device memory owned by libcrt_b200 (plumbing for torch.distributed)."""
    import torch

    class _Dev:
        pass

    d = _Dev()
    d.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": dtype, "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(d, device="cuda")


def gather_framebuffer(renderer, group=None):
    """All-gather every rank's compact accumulator (float4 per owned pixel) on torch's current stream. Returns the
    rank-major device tensor [world, padded_pixels, 4] that ``B200Renderer.resolve_gathered`` consumes (float parity
    checks; the display path is ``gather_rgba8``).  No host synchronisation: torch's stream waits for the render on the
    device, and the renderer's stream waits for the collective before ``resolve_gathered`` reads its result."""
    import torch
    import torch.distributed as dist

    ptr, padded = renderer.accum_device_ptr()
    cur = torch.cuda.current_stream().cuda_stream
    renderer.stream_wait_ctx(cur)
    local = as_torch(ptr, (padded, 4))
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    out = torch.empty((world, padded, 4), dtype=torch.float32, device=local.device)
    if world == 1:
        out[0].copy_(local)
    else:
        dist.all_gather_into_tensor(out.view(-1), local.view(-1), group=group)
    renderer.ctx_wait_stream(cur)
    return out


def gather_rgba8(renderer, dst: int = 0, group=None):
    """The display read-out: every rank resolves ITS pixels to RGBA8 on its own GPU (4 bytes per pixel instead of the
    accumulator's 16), the buffers are collected on rank `dst` only (NCCL send/recv under ``torch.distributed.gather``)
    and de-interleaved there into ``renderer.h_imgPtr``.  Returns that image on `dst`, None elsewhere.  The only host
    synchronisation is the final device -> host copy on `dst`."""
    import torch
    import torch.distributed as dist

    ptr, nbytes = renderer.resolve_local_rgba8_device()
    cur = torch.cuda.current_stream().cuda_stream
    renderer.stream_wait_ctx(cur)
    local = as_torch(ptr, (nbytes,), "|u1")
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        gathered = local
    elif rank == dst:
        gathered = torch.empty((world, nbytes), dtype=torch.uint8, device=local.device)
        dist.gather(local, list(gathered.unbind(0)), dst=dst, group=group)
    else:
        dist.gather(local, None, dst=dst, group=group)
        return None
    renderer.ctx_wait_stream(cur)
    return renderer.assemble_rgba8(gathered.data_ptr())
