"""Host side of the tight per-triangle BVH (cudart_b200/csrc/host_scene.cpp build_tight_bvh).

The structure is only allowed to SKIP a triangle for a ray when the reference's fp32 Moeller-Trumbore test
(geometry.cu:38-65) rejects it.  These tests pin that on the CPU: whenever the restated test accepts a triangle, the ray
must pass the triangle's inflated box under the kernel's slab arithmetic - for the oracle's own path rays and for
adversarial grazing rays whose determinant sits at the reference's 1e-5 acceptance threshold."""
import numpy as np
import pytest

from conftest import SCENES
from grazing import f32, grazing_rays, moeller_trumbore, region_of
from oracle import orc


def box_pass(boxes, o, d, pad=None):
    """The traversal kernel's min/max slab test (kernels.cu trace_queue, `fast` path) in numpy float32.  pad = (a1, b2)
    per box: the dynamic-pad form, every side grown by a1 * |T|max + b2 with |T|max the distance from the origin to the
    farthest corner of the box (the kernel multiplies a1 by 1.0001 to cover its approximate square root; not here)."""
    with np.errstate(all="ignore"):
        inv = (f32(1) / d).astype(f32)
        s0, s1 = (boxes[:, :3] - o).astype(f32), (boxes[:, 3:] - o).astype(f32)
        a0, a1 = (s0 * inv).astype(f32), (s1 * inv).astype(f32)
        near, far = np.minimum(a0, a1), np.maximum(a0, a1)
        if pad is not None:
            m = np.maximum(np.abs(s0), np.abs(s1)).astype(np.float64)
            tm = np.sqrt((m * m).sum(1)).astype(f32)                                        # fma chain + sqrt.approx (2 ulp)
            P = (pad[:, 0].astype(np.float64) * tm + pad[:, 1]).astype(f32)                 # fmaf; the kernel's a1 is 1.0001 x this one
            p = (P[:, None] * np.abs(inv)).astype(f32)
            near, far = (near - p).astype(f32), (far + p).astype(f32)
        return near.max(1) <= far.min(1), np.all(np.abs(inv) < 1e28, 1)


MODES = {"static-pad": 1, "dynamic-pad": 2}


def boxes_and_pads(fs, lo, hi, mode):
    """(boxes, pad terms or None, info) of the tight tree built in the given pad mode."""
    from cudart_b200.scenebuild import tight_bvh
    boxes, info, pad, dyn = tight_bvh(fs.tris, fs.nodes, lo, hi, pad_mode=MODES[mode], pads=True)
    assert dyn == (mode == "dynamic-pad")
    return boxes, (pad if dyn else None), info


def leaf_boxes(fs):
    """Per triangle: is it in a reference leaf, and that leaf's box in min/max order."""
    in_leaf = np.zeros(len(fs.tris), bool)
    box = np.zeros((len(fs.tris), 6), np.float32)
    for nd in fs.nodes[fs.nodes["numTriangles"] > 0]:
        sl = slice(int(nd["offset"]), int(nd["offset"]) + int(nd["numTriangles"]))
        in_leaf[sl] = True
        box[sl] = np.concatenate([np.minimum(nd["min"], nd["max"]), np.maximum(nd["min"], nd["max"])])
    return in_leaf, box


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("name", SCENES)
def test_accepted_grazing_hits_pass_the_inflated_box(golden, name, mode):
    fs, _ = golden(name)
    lo, hi = region_of(fs)
    boxes, pad, (n_big, n_nodes, n_wide) = boxes_and_pads(fs, lo, hi, mode)
    assert n_nodes > 0 and n_wide > 0
    o, d, k = grazing_rays(fs, 200_000)
    t = fs.tris
    acc, _, _, _, det = moeller_trumbore(o, d, t["v1"][k], t["e1"][k], t["e2"][k])
    passed, fast = box_pass(boxes[k], o, d, None if pad is None else pad[k])
    acc &= fast
    assert acc.sum() > 20_000 and (acc & (np.abs(det) < 3e-5)).sum() > 2_000      # the generator reaches the threshold regime
    # a triangle with its own inflated box: every accepted hit passes it.  A triangle kept behind its reference leaf's box
    # is only ever hit by the reference when that box passes - which then IS the box.
    in_leaf, lbox = leaf_boxes(fs)
    own = (~np.all(boxes == lbox, 1) & in_leaf)[k]
    gate, _ = box_pass(lbox[k], o, d)
    assert not (acc & own & ~passed).any()
    assert not (acc & gate & ~passed).any()


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("name", SCENES)
def test_path_ray_winners_pass_their_inflated_box(golden, name, mode):
    fs, _ = golden(name)
    small = fs.with_resolution(64, 48)
    lo, hi = region_of(small)
    boxes, pad, _ = boxes_and_pads(small, lo, hi, mode)
    osc = orc.OracleScene(small)
    _, rays = osc.render_accum(4, rng=orc.RNG_COUNTER, seed=11, threads=1, ray_log_cap=64 * 48 * 4 * 12)
    want = osc.trace_closest(rays)
    hit = want["tri"] != 0xFFFFFFFF
    o, d = rays["o"][hit], rays["d"][hit]
    assert np.all((o >= lo) & (o <= hi)), "path ray origins must stay inside the region the bound is valid for"
    passed, fast = box_pass(boxes[want["tri"][hit]], o, d, None if pad is None else pad[want["tri"][hit]])
    assert not (fast & ~passed).any()


@pytest.mark.parametrize("mode", MODES)
def test_boxes_contain_their_triangle_and_unsplit_leaves_keep_the_reference_box(golden, mode):
    for name in SCENES:
        fs, _ = golden(name)
        lo, hi = region_of(fs)
        boxes, _, (n_split, n_nodes, _) = boxes_and_pads(fs, lo, hi, mode)
        t = fs.tris
        v1 = t["v1"].astype(np.float64)
        in_leaf, leaf_box = leaf_boxes(fs)
        own = ~np.all(boxes == leaf_box, 1) & in_leaf
        assert own.sum() <= n_split <= in_leaf.sum()
        for p in (v1, v1 + t["e1"], v1 + t["e2"]):          # an inflated box contains the real triangle
            assert np.all(boxes[own, :3] <= p[own]) and np.all(p[own] <= boxes[own, 3:])


def test_degenerate_and_huge_triangles_keep_their_reference_leaf_box(golden):
    """Zero-area, huge and tiny triangles: no bound that is not deserved, no non-finite box."""
    from cudart_b200.scenebuild import tight_bvh
    from cudart_b200.scenefile import NODE_DTYPE, TRIANGLE_DTYPE
    t = np.zeros(4, TRIANGLE_DTYPE)
    t["v1"] = [[0, 0, 0], [1, 1, 1], [0, 0, 0], [2, 2, 2]]
    t["e1"] = [[1, 0, 0], [100, 0, 0], [1, 0, 0], [1e-3, 0, 0]]
    t["e2"] = [[0, 1, 0], [0, 100, 0], [2, 0, 0], [0, 1e-3, 0]]      # unit, huge, zero-area (collinear), tiny
    nodes = np.zeros(1, NODE_DTYPE)
    nodes["min"] = [[0, 0, 0]]
    nodes["max"] = [[101, 101, 3]]
    nodes["numTriangles"] = 4
    boxes, (n_split, n_nodes, n_wide) = tight_bvh(t, nodes, [-10, -10, -10], [110, 110, 10])
    assert np.all(boxes[1] == [0, 0, 0, 101, 101, 3])               # s >= 0.25: behind the leaf's own box
    assert np.all(np.isfinite(boxes)) and n_nodes >= 1 and n_wide >= 1
    ext = boxes[:, 3:] - boxes[:, :3]
    assert np.all(ext[3] < 0.1) and np.all(ext[3] > 1e-3)            # a tiny triangle keeps a tiny (but padded) box
    assert n_split >= 1


def test_pad_mode_is_chosen_by_the_size_of_the_baked_in_pad(golden):
    """build_tight_bvh's auto rule (host_scene.cpp): dynamic pad when the baked-in pad averages more than 0.3 of the
    triangle's longest edge - the coarse terrain (pad = 8 edges) and the hand-written scenes - static where the pad is a few
    per cent of the triangles (the three bundled scenes; measured: profiles/r01z_pad_ab.txt)."""
    from cudart_b200.scenebuild import PAD_AUTO, tight_bvh
    want = {"cornell": False, "cornellSphere": False, "monkeys": True, "terrain64": True, "tiny": True, "soup": True}
    for name, dynamic in want.items():
        fs, _ = golden(name)
        lo, hi = region_of(fs)
        _, _, _, dyn = tight_bvh(fs.tris, fs.nodes, lo, hi, pad_mode=PAD_AUTO, pads=True)
        assert dyn == dynamic, name
