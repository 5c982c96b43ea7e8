"""The drop-in surface, end to end on the GPU box:

  * integration/_build/CudaRT_dropin = the reference's UNMODIFIED main.cpp, scene.cpp, bvh.cpp, launcher.cpp and
    terminal_launcher.cpp (objects built by oracle/Makefile) linked with integration/b200_renderer.cpp - the binding
    INTEGRATION.md describes, compiled against the reference's real renderer.h / scene.h - and libcrt_b200.so.  The
    reference's own launcher loop (one renderOneSamplePerPixel per sample) drives the library; its stb PNG must hold
    exactly the pixels the repo's own CLI writes.
  * the repo CLI's --progress (stand-in for the reference's --x window, window.cpp:17-68) and --cpu hand-over.
"""
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "cudart_b200", "CudaRT")
DROPIN = os.path.join(ROOT, "integration", "_build", "CudaRT_dropin")


def read_png_rgba(path):
    """Minimal PNG reader (8-bit RGBA, non-interlaced, all five filter types): stb's writer picks filters per row."""
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    off, idat, w, h = 8, b"", 0, 0
    while off < len(data):
        n, tag = struct.unpack(">I4s", data[off:off + 8])
        body = data[off + 8:off + 8 + n]
        if tag == b"IHDR":
            w, h, depth, ctype, _, _, interlace = struct.unpack(">IIBBBBB", body)
            assert (depth, ctype, interlace) == (8, 6, 0)
        elif tag == b"IDAT":
            idat += body
        off += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + 4 * w)
    out = np.zeros((h, 4 * w), np.uint8)
    prev = np.zeros(4 * w, np.int32)
    for y in range(h):
        f, line = int(raw[y, 0]), raw[y, 1:].astype(np.int32)
        cur = np.zeros(4 * w, np.int32)
        if f == 0:
            cur = line
        elif f == 2:
            cur = (line + prev) & 255
        else:
            for x in range(4 * w):
                a = cur[x - 4] if x >= 4 else 0
                b, c = prev[x], (prev[x - 4] if x >= 4 else 0)
                if f == 1:
                    p = a
                elif f == 3:
                    p = (a + b) >> 1
                else:
                    pa, pb, pc = abs(b - c), abs(a - c), abs(a + b - 2 * c)
                    p = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                cur[x] = (line[x] + p) & 255
        out[y] = cur
        prev = cur
    return out.reshape(h, w, 4)


def scene_dir(tmp_path, name="tiny"):
    src = os.path.join(ROOT, "tests", "golden", name)
    (tmp_path / ".env").write_text("MESHES_PATH=%s/meshes\nCAMERA_PATH=%s/settings\nTEXTURES_PATH=%s/textures\n" % (src, src, src))
    return tmp_path


def run_in(wd, exe, *args, env=None):
    wd.mkdir(exist_ok=True)
    return subprocess.run([exe] + [str(a) for a in args], cwd=wd, capture_output=True, text=True, env=env)


def test_png_reader_handles_every_filter(tmp_path):
    """(CPU) the reader above against zlib streams built with each filter type."""
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (7, 9, 4), dtype=np.uint8)
    rows = []
    prev = np.zeros(36, np.int32)
    for y in range(7):
        cur = img[y].reshape(-1).astype(np.int32)
        f = y % 5
        line = np.zeros(36, np.int32)
        for x in range(36):
            a = cur[x - 4] if x >= 4 else 0
            b, c = prev[x], (prev[x - 4] if x >= 4 else 0)
            pa, pb, pc = abs(b - c), abs(a - c), abs(a + b - 2 * c)
            paeth = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
            p = (0, a, b, (a + b) >> 1, paeth)[f]
            line[x] = (cur[x] - p) & 255
        rows.append(bytes([f]) + line.astype(np.uint8).tobytes())
        prev = cur

    def chunk(tag, body):
        return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body))
    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 9, 7, 8, 6, 0, 0, 0)) + \
        chunk(b"IDAT", zlib.compress(b"".join(rows))) + chunk(b"IEND", b"")
    (tmp_path / "t.png").write_bytes(png)
    assert np.array_equal(read_png_rgba(tmp_path / "t.png"), img)


@pytest.mark.gpu
@pytest.mark.parametrize("name,spp,w,h", [("tiny", 8, 40, 30), ("soup", 5, 48, 32)])
def test_reference_main_drives_the_library_and_writes_the_same_pixels(tmp_path, name, spp, w, h):
    """Unmodified reference main.cpp + TerminalLauncher::render + stb PNG on top of libcrt_b200.so == the repo's CLI.
    The reference loader / BVH builder feed the drop-in, the repo's own loader / builder feed the CLI: equal pixels also
    re-prove, on the device, that the two scene builds agree."""
    if not os.path.exists(DROPIN):
        pytest.skip("integration/_build/CudaRT_dropin is not built (needs /root/reference at build time)")
    base = scene_dir(tmp_path, name)
    a = run_in(base / "build_dropin", DROPIN, "-f", name, "-s", spp, "-w", w, "-h", h)
    assert a.returncode == 0, a.stderr + a.stdout
    assert "Rendered to %s.png for %d samples per pixel." % (name, spp) in a.stdout       # main.cpp:183, Renderer::samplesRendered
    b = run_in(base / "build_cli", CLI, "-f", name, "-s", spp, "-w", w, "-h", h)
    assert b.returncode == 0, b.stderr
    img_a, img_b = read_png_rgba(base / "build_dropin" / (name + ".png")), read_png_rgba(base / "build_cli" / (name + ".png"))
    assert img_a.shape == (h, w, 4) and img_a[..., :3].any()
    assert np.array_equal(img_a, img_b)


@pytest.mark.gpu
def test_progressive_output_rewrites_the_image_and_ends_on_the_same_pixels(tmp_path):
    """--progress N: the running image is written every N samples, the progress lines appear, and the
    final image equals the one-shot render (sample indices do not depend on how the samples were batched)."""
    base = scene_dir(tmp_path)
    one = run_in(base / "a", CLI, "-f", "tiny", "-s", 7, "-w", 40, "-h", 30)
    prog = run_in(base / "b", CLI, "-f", "tiny", "-s", 7, "-w", 40, "-h", 30, "--progress", 3)
    assert one.returncode == 0 and prog.returncode == 0, one.stderr + prog.stderr
    assert "3 / 7 samples per pixel" in prog.stdout and "6 / 7 samples per pixel" in prog.stdout
    assert "Rendered to tiny.png for 7 samples per pixel." in prog.stdout
    assert np.array_equal(read_png_rgba(base / "a" / "tiny.png"), read_png_rgba(base / "b" / "tiny.png"))
    # an interrupted progressive run leaves the last running image behind: stop after the first write
    part = run_in(base / "c", CLI, "-f", "tiny", "-s", 3, "-w", 40, "-h", 30)
    first = run_in(base / "d", CLI, "-f", "tiny", "-s", 7, "-w", 40, "-h", 30, "--progress", 3, "--stop-after-progress", 1)
    assert first.returncode == 0, first.stderr
    assert np.array_equal(read_png_rgba(base / "c" / "tiny.png"), read_png_rgba(base / "d" / "tiny.png"))


def test_cpu_flag_hands_over_to_a_reference_binary_outside_the_oracle(tmp_path):
    """(CPU) --cpu never runs anything of this repository: it execs $CUDART_REFERENCE (or baseline/_ref/CudaRT_ref) with
    the same arguments, and fails loudly when there is none."""
    base = scene_dir(tmp_path)
    fake = tmp_path / "fake_ref.sh"
    fake.write_text("#!/bin/sh\necho REFERENCE \"$@\"\n")
    fake.chmod(0o755)
    ok = run_in(base / "a", CLI, "-f", "tiny", "-s", 2, "--cpu", env=dict(os.environ, CUDART_REFERENCE=str(fake)))
    assert ok.returncode == 0 and "REFERENCE -f tiny -s 2 --cpu" in ok.stdout
    bad = run_in(base / "b", CLI, "-f", "tiny", "--cpu", env=dict(os.environ, CUDART_REFERENCE=str(tmp_path / "missing")))
    assert bad.returncode != 0 and "no CPU renderer and no fallback" in bad.stderr
