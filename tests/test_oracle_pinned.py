"""Pins the CPU oracle (oracle/crt_oracle.c) to golden data produced by the UNMODIFIED reference
(tests/golden/make_golden.py -> oracle/_ref/ref_dump): bit-exact closest hits vs the reference's
intersectBVH, bit-exact radiance sums vs its samplePixel under the same rand() stream, and identical
RGBA8 bytes vs its Renderer API.  CPU only."""
import numpy as np
import pytest

from conftest import SCENES
from oracle import orc


def bits(a):
    return np.ascontiguousarray(a).view(np.uint8)


@pytest.mark.parametrize("name", SCENES)
def test_closest_hit_matches_reference_intersectBVH(golden, name):
    fs, g = golden(name)
    hits = orc.OracleScene(fs).trace_closest(g["rays"])
    assert np.array_equal(bits(hits), bits(g["hits"]))


@pytest.mark.parametrize("name", SCENES)
def test_order_independent_rule_matches_reference(golden, name):
    """own-leaf-box test + explicit tie key == reference traversal, bit for bit."""
    fs, g = golden(name)
    rays = g["rays"][::3] if name == "terrain64" else g["rays"]
    ref = g["hits"][::3] if name == "terrain64" else g["hits"]
    hits = orc.OracleScene(fs).trace_closest(rays, unordered=True)
    assert np.array_equal(hits["tri"], ref["tri"])
    hit = ref["tri"] != 0xFFFFFFFF
    for k in ("t", "u", "v"):
        assert np.array_equal(hits[k][hit].view(np.uint32), ref[k][hit].view(np.uint32))
    assert np.array_equal(hits["t"][~hit].view(np.uint32), ref["t"][~hit].view(np.uint32))


@pytest.mark.parametrize("name", SCENES)
def test_radiance_and_rgba_match_reference_samplePixel(golden, name):
    fs, g = golden(name)
    osc = orc.OracleScene(fs)
    assert (osc.width, osc.height) == (g["W"], g["H"])
    accum, _ = osc.render_accum(g["spp"], rng=orc.RNG_RAND, seed=g["seed"])
    assert np.array_equal(accum.view(np.uint32), g["accum"].view(np.uint32))
    assert np.array_equal(orc.resolve_rgba8(accum, g["spp"]), g["rgba"])


def test_counter_rng_is_uniform_and_order_free(golden):
    fs, _ = golden("cornell")
    osc = orc.OracleScene(fs)
    a, _ = osc.render_accum(2, rng=orc.RNG_COUNTER, seed=5, threads=1)
    b, _ = osc.render_accum(2, rng=orc.RNG_COUNTER, seed=5, threads=4)
    assert np.array_equal(a, b)                         # schedule independent
    top, _ = osc.render_accum(2, rng=orc.RNG_COUNTER, seed=5, threads=2, rows=(0, 10))
    assert np.array_equal(top[:10], a[:10]) and not top[10:].any()
    u = np.array([orc.counter_uniform(1, p, s, d) for p in range(40) for s in range(10) for d in range(10)])
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.02
