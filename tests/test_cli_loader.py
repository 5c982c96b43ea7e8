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
