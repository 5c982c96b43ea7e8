"""The drop-in command line's scene loading (cudart_b200/csrc/cli: .env, camera JSON, OBJ/MTL, PNG textures, material
table, BVH, light list) against dumps of the reference's own loader: every array bit for bit.  CPU only."""
import os
import subprocess

import numpy as np
import pytest

from cudart_b200 import load_crts

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "cudart_b200", "CudaRT")
ASSETS = os.path.join(ROOT, "oracle", "_ref", "assets")


def run_cli(tmp_path, assets, scene, w, h, *extra):
    """The reference reads ../.env relative to its working directory (main.cpp:35)."""
    (tmp_path / "build").mkdir(exist_ok=True)
    (tmp_path / ".env").write_text("# test\nMESHES_PATH=%s/meshes\nCAMERA_PATH=%s/settings\nTEXTURES_PATH=%s/textures\n" % (assets, assets, assets))
    out = tmp_path / (scene + ".crts")
    p = subprocess.run([CLI, "-f", scene, "-s", "1", "-w", str(w), "-h", str(h), "--dump-scene", str(out), *extra],
                       cwd=tmp_path / "build", capture_output=True, text=True)
    return p, out


def assert_same_scene(mine, ref):
    for k in mine.tris.dtype.names:
        assert np.array_equal(mine.tris[k].view(np.uint32), ref.tris[k].view(np.uint32)), "tris." + k
    assert mine.n_nodes == ref.n_nodes
    for k in mine.nodes.dtype.names:
        assert np.array_equal(mine.nodes[k], ref.nodes[k]), "nodes." + k
    assert np.array_equal(mine.mat_floats.view(np.uint32), ref.mat_floats.view(np.uint32))
    assert np.array_equal(mine.mat_indices, ref.mat_indices)
    assert np.array_equal(mine.light_indices, ref.light_indices)
    assert np.float32(mine.lights_area) == np.float32(ref.lights_area)
    assert mine.camera.tobytes() == ref.camera.tobytes()
    assert np.array_equal(mine.tex_dims, ref.tex_dims) and np.array_equal(mine.tex_offsets, ref.tex_offsets)
    assert np.array_equal(mine.tex_data.view(np.uint32), ref.tex_data.view(np.uint32))


def test_tiny_scene_equals_reference_loader(tmp_path):
    """tests/golden/tiny: v/vt/vn, v//vn, bare v and negative indices, `o` and `g`, a material switch inside a group,
    a textured material, and two materials the reference's Material::operator== cannot tell apart."""
    p, out = run_cli(tmp_path, os.path.join(ROOT, "tests", "golden", "tiny"), "tiny", 40, 30, "--no-render")
    assert p.returncode == 0, p.stderr
    ref = load_crts(os.path.join(ROOT, "tests", "golden", "tiny.crts.gz"))
    assert_same_scene(load_crts(out), ref)
    assert ref.tris["materialId"].tolist() == [0, 0, 2, 2, 3, 3, 4, 4, 4, 4, 1, 1]       # green wall got red's id, D10
    assert "Total number of triangles:\t12" in p.stdout and "Total BVH Nodes: 11" in p.stdout


@pytest.mark.parametrize("scene", ["cornell", "cornellSphere", "monkeys", "terrain64"])
def test_fixture_scenes_equal_reference_loader(tmp_path, golden, scene):
    if not os.path.exists(os.path.join(ASSETS, "meshes", scene + ".obj")):
        pytest.skip("oracle/_ref/assets not generated here (python oracle/make_assets.py)")
    fs, g = golden(scene)
    p, out = run_cli(tmp_path, ASSETS, scene, g["W"], g["H"], "--no-render")
    assert p.returncode == 0, p.stderr
    assert_same_scene(load_crts(out), fs)


def test_cli_surface(tmp_path):
    p = subprocess.run([CLI], capture_output=True, text=True)
    assert p.returncode == 1 and "Usage: CudaRT <options>" in p.stderr and "--cpu" in p.stderr      # main.cpp:59-71
    p, _ = run_cli(tmp_path, os.path.join(ROOT, "tests", "golden", "tiny"), "nosuchscene", 8, 8, "--no-render")
    assert p.returncode == 1 and "Failed to load mesh" in p.stderr
    import torch
    if not torch.cuda.is_available():
        p, _ = run_cli(tmp_path, os.path.join(ROOT, "tests", "golden", "tiny"), "tiny", 8, 8)
        assert p.returncode == 1 and "no CPU fallback" in p.stderr


QUIRKY_HEAD = """mtllib quirky.mtl
o first
v 0 0 0
v 1 0 0
v 0 1 0
v 1.5 1.25 -0.75
v -2.500000 3.250000 1e-3
v +4 5E+1 -6.5e-2
v  7 8 9
v\t1 2 3
v 1 2 3 
v 0.000000 -0.000000 -0
v 1e-42 0.1 0.2
v 123456789012345678901234567890 1 1
v 3.4028235e38 1.17549435e-38 16777217
vn 0 0 1
vn 0.0 1.0 0.0
vt 0.25 0.75
vt 1 0 0
usemtl red
f 1 2 3
f 1/1/1 2/2/2 3/1/1
f 1//1 2//2 3//1
f 1/1 2/2 3/1
f -1 -2 -3
f 4 5 6
f  1 2 3
f 1 2 3 
f 1/ 2/ 3/
g second
usemtl green
f 7 8 9
f 10 9 8
f 0000000001 2 3
f 6 8 10
"""
QUIRKY_MTL = """newmtl red
Kd 0.8 0.1 0.1
Ka 1 1 1
type EMISSIVE
newmtl green
Kd 0.1 0.8 0.1
"""
QUIRKY_CAMERA = ('{"fStop": 1000.0, "eye": [0.0, 0.0, 5.0], "focalLength": 0.032, "fieldOfView": 60.0, '
                 '"upDirection": [0.0, 1.0, 0.0], "viewDirection": [0.0, 0.0, -5.0]}')


def _quirky_assets(tmp_path, obj_text, crlf=False):
    a = tmp_path / "assets"
    for d in ("meshes", "settings", "textures"):
        (a / d).mkdir(parents=True, exist_ok=True)
    (a / "meshes" / "quirky.obj").write_bytes(obj_text.replace("\n", "\r\n").encode() if crlf else obj_text.encode())
    (a / "meshes" / "quirky.mtl").write_text(QUIRKY_MTL)
    (a / "settings" / "quirky-camera.json").write_text(QUIRKY_CAMERA)
    return str(a)


def _run_both_parsers(tmp_path, assets):
    res = []
    for slow in (False, True):
        (tmp_path / "build").mkdir(exist_ok=True)
        (tmp_path / ".env").write_text("MESHES_PATH=%s/meshes\nCAMERA_PATH=%s/settings\nTEXTURES_PATH=%s/textures\n" % (assets, assets, assets))
        out = tmp_path / ("quirky_%d.crts" % slow)
        env = dict(os.environ)
        env.pop("CRT_OBJ_SLOW", None)
        if slow:
            env["CRT_OBJ_SLOW"] = "1"
        p = subprocess.run([CLI, "-f", "quirky", "-s", "1", "-w", "16", "-h", "12", "--dump-scene", str(out), "--no-render"],
                           cwd=tmp_path / "build", capture_output=True, text=True, env=env)
        res.append((p.returncode, p.stdout.replace(str(out), "OUT"), p.stderr, out.read_bytes() if out.exists() else None))
    return res


def test_fast_obj_lines_follow_the_reference_parsers_rules(tmp_path):
    """load_obj parses the strict common form of v / vn / vt / f lines straight from the file buffer and sends everything else
    through the reference parser's token rules.  On text full of near misses - doubled blanks, tabs, trailing blanks, '+' signs,
    exponents, signed zeros, subnormal and overflowing numbers, `a/b`, `a//c`, `a/`, negative and zero-padded indices - the
    scene must be byte-identical to the one obtained with EVERY line on the slow path (CRT_OBJ_SLOW), and so must the outcome
    (exit code, messages) when a line is an error."""
    import random
    lines = QUIRKY_HEAD.splitlines()
    # 1. every prefix that ends in a face (some stop at lines the reference parser rejects: both parsers must fail alike)
    cases = ["\n".join(lines[:k]) + "\n" for k in range(20, len(lines) + 1)]
    # 2. the error lines dropped: a file that loads
    good = [l for l in lines if not (l.startswith("v  7") or l.startswith("v 1e-42") or l.startswith("v 1234567890123") or l.startswith("f  1")
                                     or l == "f 1 2 3 " or l.startswith("f 0000000001"))]
    cases += ["\n".join(good) + "\n", "\n".join(good)]                       # with and without the final newline
    # 3. shuffled faces over the same vertices, and an out-of-range index
    rng = random.Random(5)
    faces = [l for l in good if l.startswith("f ")]
    head = [l for l in good if not l.startswith("f ") and not l.startswith("g ") and not l.startswith("usemtl green")]
    for _ in range(4):
        rng.shuffle(faces)
        cases.append("\n".join(head + faces) + "\n")
    cases.append("\n".join(head + faces + ["f 1 2 99"]) + "\n")
    cases.append("\n".join(head + faces + ["f 1 2 -99"]) + "\n")
    loaded = 0
    for k, text in enumerate(cases):
        for crlf in ((False, True) if k % 7 == 0 else (False,)):
            d = tmp_path / ("case%d_%d" % (k, crlf))
            d.mkdir()
            fast, slow = _run_both_parsers(d, _quirky_assets(d, text, crlf))
            assert fast[:3] == slow[:3], (k, crlf, text, fast[:3], slow[:3])
            assert fast[3] == slow[3], (k, crlf)
            loaded += fast[0] == 0 and fast[3] is not None
    assert loaded >= 6


def test_png_writer_bands_decode_to_the_same_pixels(tmp_path):
    """write_png_rgba8 deflates the image in bands of rows on all host threads and splices the bands into ONE zlib stream
    (cli/scene_io.cpp).  The file must decode - chunk CRCs, Adler-32 and all - to exactly the pixels handed in, for sizes with
    one band, many bands, ragged last bands and rows wider than a band, and its bytes must not depend on how many CPUs the
    process may use."""
    import struct
    import zlib
    src = tmp_path / "png_main.cpp"
    src.write_text(r'''
#include "cli/scene_io.h"
#include <cstdio>
#include <cstdlib>
int main(int argc, char** argv) {
    const unsigned w = (unsigned)atoi(argv[1]), h = (unsigned)atoi(argv[2]);
    std::vector<uint8_t> img((size_t)w * h * 4);
    uint32_t s = 12345u + w * 31u + h;
    for (size_t i = 0; i < img.size(); i++) { s = s * 1664525u + 1013904223u; img[i] = (i % 4 == 3) ? 255 : (uint8_t)(((i / 4) % w) / 3 + ((s >> 24) & 15)); }
    FILE* f = fopen(argv[4], "wb"); fwrite(img.data(), 1, img.size(), f); fclose(f);
    crtcli::write_png_rgba8(argv[3], img.data(), w, h);
    return 0;
}
''')
    exe = tmp_path / "png_main"
    csrc = os.path.join(ROOT, "cudart_b200", "csrc")
    libdir = os.path.join(ROOT, "cudart_b200")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", csrc, str(src), os.path.join(csrc, "cli", "scene_io.cpp"), "-o", str(exe),
                    "-L", libdir, "-lcrt_b200", "-lz", "-lpthread", "-Wl,-rpath," + libdir], check=True)
    for w, h in ((1, 1), (40, 30), (640, 480), (1920, 1080), (7, 5000), (70000, 3)):
        png, raw = tmp_path / "o.png", tmp_path / "o.raw"
        subprocess.run([str(exe), str(w), str(h), str(png), str(raw)], check=True)
        data = png.read_bytes()
        assert data[:8] == b"\x89PNG\r\n\x1a\n"
        off, idat, tags = 8, b"", []
        while off < len(data):
            n, tag = struct.unpack(">I4s", data[off:off + 8])
            body = data[off + 8:off + 8 + n]
            assert zlib.crc32(tag + body) == struct.unpack(">I", data[off + 8 + n:off + 12 + n])[0]
            tags.append(tag)
            if tag == b"IHDR":
                assert struct.unpack(">IIBBBBB", body) == (w, h, 8, 6, 0, 0, 0)
            if tag == b"IDAT":
                idat += body
            off += 12 + n
        assert tags == [b"IHDR", b"IDAT", b"IEND"]
        rows = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + 4 * w)        # checks the combined Adler-32
        assert (rows[:, 0] == 0).all()
        assert np.array_equal(rows[:, 1:], np.frombuffer(raw.read_bytes(), np.uint8).reshape(h, 4 * w))
        for threads in ("1", "5"):
            subprocess.run([str(exe), str(w), str(h), str(tmp_path / "one.png"), str(raw)], check=True, env=dict(os.environ, CRT_HOST_THREADS=threads))
            assert (tmp_path / "one.png").read_bytes() == data


def test_fast_obj_lines_random_near_misses(tmp_path):
    """Randomised version of the test above: vertex / normal / texture / face lines assembled from well-formed and
    slightly malformed pieces (random blanks, tabs, signs, exponents, long digit strings, empty index fields, indices around
    the valid range).  Whatever the reference parser's rules make of each file - a scene or an error - the fast path must
    make the same of it."""
    import random
    rng = random.Random(20261018)

    def number():
        k = rng.random()
        if k < 0.55:
            return "%.6f" % rng.uniform(-50, 50)
        if k < 0.65:
            return str(rng.randint(-9, 9))
        if k < 0.72:
            return "%.3e" % rng.uniform(-1e3, 1e3)
        if k < 0.76:
            return rng.choice(["-0", "0.000000", "-0.000000", "+1.5", ".5", "5.", "1e-45", "1e39", "1e-39", "0x10", "nan", "inf", "1.5f", "1,5", ""])
        return "%.9g" % rng.uniform(-1e-3, 1e-3)

    def sep():
        return " " if rng.random() < 0.93 else rng.choice(["  ", "\t", " \t"])

    def index(n):
        k = rng.random()
        if k < 0.8:
            return str(rng.randint(1, max(n, 1)))
        if k < 0.88:
            return str(-rng.randint(1, max(n, 1)))
        return rng.choice(["0", str(n + 1), str(-n - 1), "0000000002", "12345678901", "", "1x", "+1"])

    loaded = failed = 0
    for case in range(40):
        nv, nn, nt = rng.randint(3, 12), rng.randint(1, 4), rng.randint(1, 4)
        # mostly clean files (so that a good share loads), a few with several near misses
        dirty = 0.0 if case % 3 == 0 else (0.02 if case % 3 == 1 else 0.15)

        def maybe(clean, messy):
            return messy() if rng.random() < dirty else clean

        out = ["mtllib quirky.mtl", "o first", "usemtl red"]
        for _ in range(nv):
            out.append("v" + maybe(" ", sep) + maybe(" ", sep).join(maybe("%.6f" % rng.uniform(-5, 5), number) for _ in range(3)) + maybe("", lambda: rng.choice([" ", "\t", " 1.0"])))
        for _ in range(nn):
            out.append("vn " + " ".join(maybe("%.4f" % rng.uniform(-1, 1), number) for _ in range(3)))
        for _ in range(nt):
            out.append("vt " + " ".join(maybe("%.4f" % rng.uniform(0, 1), number) for _ in range(rng.choice([2, 2, 3]))))
        for f in range(rng.randint(2, 10)):
            form = rng.choice(["v", "v/t", "v//n", "v/t/n"])
            corners = []
            for _ in range(3 if rng.random() > dirty / 3 else rng.choice([2, 4])):
                v, t, n = maybe(str(rng.randint(1, nv)), lambda: index(nv)), maybe(str(rng.randint(1, nt)), lambda: index(nt)), maybe(str(rng.randint(1, nn)), lambda: index(nn))
                corners.append({"v": v, "v/t": v + "/" + t, "v//n": v + "//" + n, "v/t/n": v + "/" + t + "/" + n}[form] + maybe("", lambda: rng.choice(["/", "//"])))
            out.append("f" + maybe(" ", sep) + maybe(" ", sep).join(corners) + maybe("", lambda: " "))
            if f == 3:
                out += ["g second", "usemtl green"]
        text = "\n".join(out) + ("\n" if rng.random() < 0.8 else "")
        d = tmp_path / ("case%d" % case)
        d.mkdir()
        fast, slow = _run_both_parsers(d, _quirky_assets(d, text, crlf=(case % 13 == 5)))
        assert fast[:3] == slow[:3], (case, text, fast[:3], slow[:3])
        assert fast[3] == slow[3], (case, text)
        loaded += fast[0] == 0
        failed += fast[0] != 0
    assert loaded >= 8 and failed >= 8, (loaded, failed)
