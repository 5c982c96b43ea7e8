"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol include/crt_abi.h
declares, fails loudly (no CPU fallback) without a GPU, and the multi-process row-tile logic gathers a
framebuffer correctly across world_size-2 gloo ranks."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from cudart_b200 import capi
    hdr = open(os.path.join(ROOT, "include", "crt_abi.h")).read()
    declared = set(re.findall(r"\b(crt_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no prototypes found"
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    L = ctypes.CDLL(capi.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert capi.lib().crt_abi_version() == 1


def test_no_cpu_fallback_without_a_gpu(golden):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from cudart_b200 import capi
    from cudart_b200.renderer import B200Renderer
    fs, g = golden("cornell")
    with pytest.raises(capi.CrtError) as e:
        B200Renderer(fs, 16, 16)
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    """The checker is never on the product path: nothing under cudart_b200/ imports, links, executes or even names
    anything under oracle/ (the CLI's --cpu hands over to a reference binary installed OUTSIDE oracle/)."""
    pat = re.compile(r"(from\s+oracle|import\s+oracle|crt_oracle|liborc|orc\.py|oracle/)")
    for dirpath, _, files in os.walk(os.path.join(ROOT, "cudart_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                hits = [m.group(0) for m in pat.finditer(text) if "make_assets" not in text[max(0, m.start() - 10):m.end() + 20]]
                assert not hits, (f, hits)


def test_row_tile_math_partitions_the_image():
    from cudart_b200.sharding import deinterleave, owned_rows, padded_rows
    for H, world, tr in [(1080, 8, 8), (50, 3, 8), (7, 4, 8), (2160, 2, 16), (33, 1, 8)]:
        rows = [owned_rows(H, r, world, tr) for r in range(world)]
        allrows = np.sort(np.concatenate(rows))
        assert np.array_equal(allrows, np.arange(H))
        pr = padded_rows(H, world, tr)
        assert all(len(r) <= pr for r in rows)
        W = 5
        img = np.random.default_rng(0).random((H, W, 4)).astype(np.float32)
        gathered = np.zeros((world, pr * W, 4), np.float32)
        for r in range(world):
            gathered[r, : len(rows[r]) * W] = img[rows[r]].reshape(-1, 4)
        assert np.array_equal(deinterleave(gathered, W, H, world, tr), img)


WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from cudart_b200.sharding import owned_rows, padded_rows, deinterleave
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank, world, W, H, tr = dist.get_rank(), 2, 12, 37, 8
img = np.random.default_rng(3).random((H, W, 4)).astype(np.float32)        # what a 1-GPU render would hold
rows = owned_rows(H, rank, world, tr)
local = torch.zeros(padded_rows(H, world, tr) * W, 4)
local[: len(rows) * W] = torch.from_numpy(img[rows].reshape(-1, 4))       # this rank's compact accumulator
out = torch.empty(world * local.shape[0], 4)
dist.all_gather_into_tensor(out.view(-1), local.view(-1))
ok = np.array_equal(deinterleave(out.numpy().reshape(world, -1, 4), W, H, world, tr), img)
# the display read-out (sharding.gather_rgba8): RGBA8 resolved per rank, collected on rank 0 only, de-interleaved there
img8 = (img * 255).astype(np.uint8)
local8 = torch.zeros(padded_rows(H, world, tr) * W, 4, dtype=torch.uint8)
local8[: len(rows) * W] = torch.from_numpy(img8[rows].reshape(-1, 4))
if rank == 0:
    got8 = torch.empty(world, local8.numel(), dtype=torch.uint8)
    dist.gather(local8.view(-1), list(got8.unbind(0)), dst=0)
    ok = ok and np.array_equal(deinterleave(got8.numpy().reshape(world, -1, 4), W, H, world, tr), img8)
else:
    dist.gather(local8.view(-1), None, dst=0)
t = torch.tensor([1.0 if ok else 0.0]); dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if t.item() == 1.0 else 1)
"""


def test_two_rank_gloo_gather_reassembles_the_framebuffer(tmp_path):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "port": port})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)]) for r in range(2)]
    assert [p.wait(timeout=120) for p in procs] == [0, 0]


def test_scene_fingerprint_sees_every_byte():
    """crt_update_scene rebuilds its derived arrays iff this fingerprint of the caller's arrays changes (crt_api.cu):
    deterministic, and a single flipped bit anywhere - first byte, a 4 MB chunk seam of the threaded path, the ragged tail -
    changes it."""
    import ctypes as C
    from cudart_b200 import capi
    lib = capi.lib()

    def fp(a):
        out = (C.c_uint64 * 2)()
        capi.check(lib.crt_fingerprint(a.ctypes.data, a.nbytes, C.byref(out)))
        return (int(out[0]), int(out[1]))

    rng = np.random.default_rng(7)
    for nbytes in (1, 31, 32, 33, 4096 + 5, (9 << 20) + 13):          # the last one takes the chunked, threaded path
        a = rng.integers(0, 256, nbytes, dtype=np.uint8)
        base = fp(a)
        assert base == fp(a.copy())
        seen = {base}
        for pos in sorted({0, nbytes // 2, nbytes - 1, min(nbytes - 1, (4 << 20) - 1), min(nbytes - 1, 4 << 20), min(nbytes - 1, 8 << 20)}):
            b = a.copy()
            b[pos] ^= 1
            seen.add(fp(b))
            assert fp(b) != base, (nbytes, pos)
        assert len(seen) >= 2
        assert fp(a[:-1]) != base if nbytes > 1 else True           # the length is part of it
    assert fp(np.zeros(64, np.uint8)) != fp(np.zeros(65, np.uint8))


def test_header_is_plain_c_and_the_library_links_from_c(tmp_path):
    """include/crt_abi.h is the boundary a C (not C++) program binds: compile a C99 translation unit against it with
    gcc -std=c99 -pedantic, link libcrt_b200.so, call it.  Without a GPU every computing entry point must come back
    with an error code and a message - never a crash, never a CPU fallback."""
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "crt_abi.h"
int main(void) {
    crt_scene_desc s;
    crt_options o;
    crt_ctx* ctx = NULL;
    uint64_t fp[2];
    memset(&s, 0, sizeof s);
    memset(&o, 0, sizeof o);
    o.struct_size = sizeof o;
    if (crt_abi_version() != CRT_ABI_VERSION) return 1;
    if (crt_create(NULL, 16, 16, &o, &ctx) != CRT_ERR_INVALID) return 2;          /* null scene */
    if (crt_create(&s, 16, 16, &o, &ctx) != CRT_ERR_INVALID || ctx != NULL) return 3; /* empty scene */
    if (strlen(crt_last_error()) == 0) return 4;
    if (crt_render(NULL, 1) != CRT_ERR_INVALID || crt_flush(NULL) != CRT_ERR_INVALID || crt_sync(NULL) != CRT_ERR_INVALID) return 5;
    if (crt_samples_rendered(NULL) != 0 || crt_spp_per_wave(NULL) != 0) return 6;
    if (crt_fingerprint("abc", 3, fp) != CRT_OK || (fp[0] == 0 && fp[1] == 0)) return 7;
    crt_destroy(NULL);
    printf("abi %d ok: %s\n", crt_abi_version(), crt_last_error());
    return 0;
}
''')
    exe = tmp_path / "abi"
    libdir = os.path.join(ROOT, "cudart_b200")
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-lcrt_b200", "-Wl,-rpath," + libdir], check=True)
    p = subprocess.run([str(exe)], capture_output=True, text=True)
    assert p.returncode == 0, (p.returncode, p.stdout, p.stderr)
    assert "abi 1 ok" in p.stdout
