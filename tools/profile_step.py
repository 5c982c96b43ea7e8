#!/usr/bin/env python3
"""One short render for profilers: `ncu ... python tools/profile_step.py [workload] [spp] [waves]`.
Numbers printed under a profiler are never bench values."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS, load_scene  # noqa: E402
from cudart_b200.renderer import B200Renderer  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "cornellSphere_1080p"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 2
waves = int(sys.argv[3]) if len(sys.argv) > 3 else 2
scene, W, H, _ = WORKLOADS[wl]
r = B200Renderer(load_scene(scene, W, H), W, H, spp, spp_per_wave=spp, stats=bool(os.environ.get("CRT_STATS")),
                 reference_topology=bool(os.environ.get("CRT_REFTOPO")), exact_visit=bool(os.environ.get("CRT_EXACT")), smem_nodes=False if os.environ.get("CRT_NOSMEM") else (True if os.environ.get("CRT_SMEMTOP") else None))
if os.environ.get("CRT_ONECALL"):          # all waves in ONE call: consecutive waves alternate between the two wave slots / streams
    r.render(spp * waves)
else:
    for _ in range(waves):
        r.render(spp)
r.sync()
c = r.counters()
print(wl, "reftopo" if os.environ.get("CRT_REFTOPO") else "sah", {k: c[k] for k in ("samples", "extend_rays", "shadow_rays", "node_visits", "tri_tests", "kernel_launches", "render_ms")})
r.close()
