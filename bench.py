#!/usr/bin/env python3
"""bench.py - the hot path's headline metric (BASELINE.json): Msamples/s (and Mrays/s) of the per-pixel
path-tracing loop on N B200s, next to the reference's own --cpu path timed on the same box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

A "step" is one pass of the hot path over one batch: `spp_per_step` samples for every pixel of the frame.
Default workload = BASELINE.json configs[1]: cornellSphere (refractive sphere, mirror + glossy boxes),
1920x1080, 512 spp = 8 steps x 64 spp.  For N > 1 (torchrun, one rank per GPU) the frame is sharded as
interleaved row tiles, every rank renders N x 64 spp of its tiles per step (per-GPU work fixed: weak
scaling) and each step ends with ONE all-gather of the compact accumulators over NCCL + a resolve on rank 0.

One JSON line on stdout (rank 0).  `value` = whole-job Msamples/s with the scene resident in HBM, timed with
CUDA events on the launching stream; `e2e` = the same through the Renderer/C-ABI call with HOST buffers
(scene upload + render + RGBA8 readback every step); `roofline` = the dominant kernel (extend+shadow
traversal) against the measured HBM peak; `cpu_baseline` = the reference's --cpu path on a bounded sample.
"""
import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    #  name                 scene fixture     W     H    total spp
    "cornellSphere_1080p": ("cornellSphere", 1920, 1080, 512),
    "cornell_1080p": ("cornell", 1920, 1080, 512),
    "monkeys_1080p": ("monkeys", 1920, 1080, 512),
    "cornell_4k": ("cornell", 3840, 2160, 1024),
    "cornell_256": ("cornell", 256, 256, 16),
    "terrain1m_1080p": ("terrain707", 1920, 1080, 256),      # BASELINE config 4: 999 700 procedural triangles
    "terrain131k_1080p": ("terrain256", 1920, 1080, 256),
    "terrain8k_256": ("terrain64", 256, 256, 16),             # small dynamic-pad scene for compute-sanitizer runs
}
SPP_PER_STEP = 64
L2_BYTES = 126 * 1024 * 1024


_SCENE_CACHE = {}


def load_scene(name, W, H):
    """Bundled scenes come from the committed reference dumps; terrainN is generated and built by the product."""
    from cudart_b200 import load_crts
    if name.startswith("terrain") and name != "terrain64":
        from cudart_b200.procedural import terrain_scene
        if name not in _SCENE_CACHE:
            _SCENE_CACHE[name] = terrain_scene(int(name[len("terrain"):]), W, H)
        return _SCENE_CACHE[name].with_resolution(W, H)
    return load_crts(os.path.join(ROOT, "tests", "golden", name + ".crts.gz")).with_resolution(W, H)


def scene_bytes(fs):
    return int(sum(np.asarray(getattr(fs, k)).nbytes for k in
                   ("tris", "nodes", "mat_floats", "mat_indices", "light_indices", "camera", "tex_dims", "tex_offsets", "tex_data")))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md recipe): NVML polled in-process every
    20 ms (nvidia_ml_py), or an `nvidia-smi -lms 100` child when NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, gpu_index, uuid=None):
        self.rows, self.proc, self.gpu, self.uuid = [], None, str(gpu_index), uuid
        self.sm, self.mx, self.reasons, self.stop, self.t, self.source = [], [], set(), threading.Event(), None, None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        if self.uuid:
            try:
                return pynvml, pynvml.nvmlDeviceGetHandleByUUID(self.uuid)
            except Exception:
                pass
        return pynvml, pynvml.nvmlDeviceGetHandleByIndex(int(self.gpu))

    def _poll_nvml(self, nv, h):
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while True:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
                mask = int(reasons_fn(h))
                self.reasons.update(k for k, bit in self.BITS.items() if mask & bit)
            except Exception:
                pass
            if self.stop.wait(0.02):
                break

    def __enter__(self):
        try:
            nv, h = self._nvml_handle()
            self.source = "nvml"
            self.t = threading.Thread(target=self._poll_nvml, args=(nv, h), daemon=True)
            self.t.start()
            return self
        except Exception:
            pass
        try:
            self.source = "nvidia-smi"
            self.proc = subprocess.Popen(["nvidia-smi", "-i", self.gpu, "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        self.stop.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=3)
            except Exception:
                self.proc.kill()
        elif self.t:
            self.t.join(timeout=1)

    def summary(self):
        sm, mx, reasons = list(self.sm), list(self.mx), set(self.reasons)
        sm += [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx += [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": self.source}


# ---------------------------------------------------------------------------------------------------------
def reference_cpu_run(scene, W, H, spp, threads):
    """Time the UNMODIFIED reference's --cpu path (oracle/_ref/CudaRT_ref) or, when it is not there, the oracle port.
    Returns (Msamples/s, kind)."""
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "CudaRT_ref")
    assets = os.path.join(ROOT, "oracle", "_ref", "assets")
    if os.path.exists(ref_bin) and os.path.exists(os.path.join(assets, "meshes", scene + ".obj")):
        with open(os.path.join(assets, ".env"), "w") as f:
            f.write("MESHES_PATH=%s/meshes\nCAMERA_PATH=%s/settings\nTEXTURES_PATH=%s/textures\n" % (assets, assets, assets))
        cwd = os.path.join(assets, "build")
        os.makedirs(cwd, exist_ok=True)
        env = dict(os.environ, OMP_NUM_THREADS=str(threads), CUDA_VISIBLE_DEVICES="")
        out = subprocess.run([ref_bin, "-f", scene, "-s", str(spp), "-w", str(W), "-h", str(H), "--cpu"], cwd=cwd, env=env,
                             capture_output=True, text=True, timeout=1800).stdout
        m = re.search(r"Elapsed time in seconds = ([0-9.eE+-]+)", out)
        if m:
            return W * H * spp / float(m.group(1)) / 1e6, "reference"
    from oracle import orc
    osc = orc.OracleScene(load_scene(scene, W, H))
    t0 = time.perf_counter()
    osc.render_accum(spp, rng=orc.RNG_COUNTER, seed=1, threads=threads)
    return W * H * spp / (time.perf_counter() - t0) / 1e6, "port"


def cpu_baseline(scene, W, H, budget_s=12.0):
    """Bounded sample of the same workload on the host cores: quarter-resolution frame, a few spp."""
    cores = os.cpu_count() or 1
    w, h = max(W // 4, 16), max(H // 4, 16)
    v1, kind = reference_cpu_run(scene, w, h, 1, 1)              # calibration run, 1 thread
    spp = int(max(1, min(16, budget_s / 2 * v1 * 1e6 / (w * h))))
    runs = {1: reference_cpu_run(scene, w, h, spp, 1)[0]}
    if cores > 1:
        runs[cores] = reference_cpu_run(scene, w, h, spp, cores)[0]
    best = max(runs, key=runs.get)
    return {"value": round(runs[best], 5), "unit": "Msamples/s", "cores": best, "kind": kind,
            "sample": "%s %dx%d %d spp via %s; threads->Msamples/s %s (host has %d cores; the reference's rand() lock makes "
                      "more threads slower)" % (scene, w, h, spp, "oracle/_ref/CudaRT_ref --cpu" if kind == "reference" else "oracle port",
                                                {k: round(v, 4) for k, v in runs.items()}, cores)}


def algorithmic_bytes(scene, W, H):
    """A_ray = 32*N_node + 48*N_tri + 48 with N_* from the REFERENCE traversal (oracle counters)."""
    from oracle import orc
    c = orc.Counters()
    orc.OracleScene(load_scene(scene, max(W // 8, 16), max(H // 8, 16))).render_accum(1, rng=orc.RNG_COUNTER, seed=3, counters=c)
    d = c.as_dict()
    a_ray = 32.0 * d["nodes_per_ray"] + 48.0 * d["tris_per_ray"] + 48.0
    return a_ray, d


# ---------------------------------------------------------------------------------------------------------
def run_reference(args, scene, W, H, total_spp):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    w, h = max(W // 4, 16), max(H // 4, 16)
    v1, kind = reference_cpu_run(scene, w, h, 1, 1)
    spp = int(max(1, min(8, 4.0 * v1 * 1e6 / (w * h))))          # ~4 s per step
    vN, _ = reference_cpu_run(scene, w, h, 1, cores) if cores > 1 else (0.0, kind)
    threads = cores if vN > v1 else 1
    for _ in range(args.warmup):
        reference_cpu_run(scene, w, h, 1, threads)
    t, vals = 0.0, []
    for _ in range(args.steps):
        v, kind = reference_cpu_run(scene, w, h, spp, threads)
        vals.append(v)
        t += w * h * spp / (v * 1e6)
    value = args.steps * w * h * spp / t / 1e6
    line = {"impl": "reference", "metric": "path-tracing throughput", "value": round(value, 5), "unit": "Msamples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * t / args.steps, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "scene": scene, "width": W, "height": H, "spp": total_spp,
                       "timed_sample": "%dx%d x %d spp per step" % (w, h, spp)},
            "cpu_baseline": {"value": round(value, 5), "unit": "Msamples/s", "cores": threads, "kind": kind,
                             "sample": "%s %dx%d %d spp per step, render time as printed by the reference binary" % (scene, w, h, spp)},
            "e2e": {"value": round(value, 5), "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cornellSphere_1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--spp-per-step", type=int, default=SPP_PER_STEP)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-smem-nodes", action="store_true", help="A/B: never stage BVH nodes in shared memory")
    ap.add_argument("--tight-bvh", default="auto", choices=["auto", "on", "off"], help="A/B: per-triangle tight BVH (DESIGN.md 2.4)")
    ap.add_argument("--pad", default="auto", choices=["auto", "static", "dynamic"], help="A/B: where the tight BVH's boxes are inflated (DESIGN.md 2.4)")
    ap.add_argument("--stats", action="store_true", help="count node visits / triangle tests (slower; not a bench value)")
    ap.add_argument("--spp-per-wave", type=int, default=0, help="samples per pixel per wavefront batch (0 = auto, ~4M paths)")
    args = ap.parse_args()
    scene, W, H, total_spp = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, scene, W, H, total_spp)
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: `python bench.py --gpus N` re-launches itself the way the driver does (one rank per GPU)
        import socket
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
                                   "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:])

    import torch
    import torch.distributed as dist
    from cudart_b200.renderer import B200Renderer
    from cudart_b200.sharding import gather_framebuffer

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device - the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    fs = load_scene(scene, W, H)
    spp_step = args.spp_per_step * world                      # weak scaling: per-GPU work fixed
    stream = torch.cuda.current_stream().cuda_stream
    r = B200Renderer(fs, W, H, total_spp, device=local, rank=rank, world=world, profile=True, stream=stream,
                     spp_per_wave=args.spp_per_wave, smem_nodes=False if args.no_smem_nodes else None,
                     tight_bvh={"auto": None, "on": True, "off": False}[args.tight_bvh], stats=args.stats,
                     dynamic_pad={"auto": None, "static": False, "dynamic": True}[args.pad])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        r.render(spp_step)
        if world > 1:
            g = gather_framebuffer(r)
            if rank == 0:
                r.resolve_gathered(g.data_ptr(), want_rgba8=False, want_f32=False)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    c0 = r.counters()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    try:
        uuid = "GPU-" + str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        uuid = None
    with ClockSampler(local, uuid) as clk:
        barrier()
        ev0.record()
        for _ in range(args.steps):
            step()
        ev1.record()
        barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
    c1 = r.counters()
    rays = torch.tensor([float(c1["extend_rays"] - c0["extend_rays"] + c1["shadow_rays"] - c0["shadow_rays"])], device="cuda")
    trav_ms = torch.tensor([c1["stage_ms"]["extend"] - c0["stage_ms"]["extend"] + c1["stage_ms"]["shadow"] - c0["stage_ms"]["shadow"]], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(rays, op=dist.ReduceOp.SUM)
        dist.all_reduce(trav_ms, op=dist.ReduceOp.MAX)
    ms, rays, trav_ms = float(ms.item()), float(rays.item()), float(trav_ms.item())
    samples = float(W) * H * spp_step * args.steps
    value = samples / (ms * 1e-3) / 1e6
    launches = c1["kernel_launches"] - c0["kernel_launches"]
    stage = {k: round(c1["stage_ms"][k] - c0["stage_ms"][k], 3) for k in c1["stage_ms"]}
    tight_nodes, tight_always = r.tight_info()
    pad_mode = ("off", "static", "dynamic")[r.tight_pad_mode()]
    per_ray = None
    if args.stats and rays > 0 and world == 1:
        per_ray = {"wide_nodes_per_ray": round((c1["node_visits"] - c0["node_visits"]) / rays, 2),
                   "tri_tests_per_ray": round((c1["tri_tests"] - c0["tri_tests"]) / rays, 2)}
    lane_stats = r.lane_stats() if args.stats else None
    r.close()

    # ---- e2e through the public API with HOST buffers: every step uploads the scene (host -> device), renders
    # spp_step samples per pixel and reads the RGBA8 frame back (device -> host), like one frame of an animation ----
    rr = B200Renderer(fs, W, H, spp_step, device=local, rank=rank, world=world, stream=stream, spp_per_wave=args.spp_per_wave,
                      tight_bvh={"auto": None, "on": True, "off": False}[args.tight_bvh],
                      dynamic_pad={"auto": None, "static": False, "dynamic": True}[args.pad])

    # the frame lands in page-locked host memory (cudaMemcpyAsync at full PCIe rate, no staging copy)
    rr.h_imgPtr = torch.empty((H, W, 4), dtype=torch.uint8, pin_memory=True).numpy()

    def e2e_step():
        rr.update_scene(fs)
        rr.render(spp_step)
        if world > 1:
            g = gather_framebuffer(rr)
            if rank == 0:
                rr.resolve_gathered(g.data_ptr(), want_rgba8=True)
        else:
            rr.copyImageBytes()

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = samples / float(e2e_s.item()) / 1e6
    scene_rebuilds, scene_uploads, scene_upload_bytes = rr.scene_info()
    rr.close()

    if rank == 0:
        peak, peak_src = measured_peaks()
        a_ray, od = algorithmic_bytes(scene, W, H)
        n_trav_launches = 7 * args.steps * max(1, -(-spp_step // (args.spp_per_wave or max(1, min(64, (16 << 20) // max(1, W * H // world))))))
        achieved = (rays / world) * a_ray / (trav_ms * 1e-3) / 1e9 if trav_ms > 0 else None
        traffic, physical = None, None
        try:            # ncu DRAM bytes per traversal launch for this workload (profiles/, captured with the same wave size)
            t = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json"))).get(args.workload)
            if t and (args.spp_per_wave in (0, t["spp_per_wave"])) and world == 1:
                traffic, physical = t["dram_bytes_per_launch"], t["ncu"]
        except (OSError, ValueError):
            pass
        line = {
            "metric": "path-tracing throughput", "value": round(value, 3), "unit": "Msamples/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "mrays_per_s": round(rays / (ms * 1e-3) / 1e6, 3), "rays_per_sample": round(rays / samples, 4),
            "config": {"workload": args.workload, "scene": scene, "width": W, "height": H, "spp_total_config": total_spp,
                       "spp_per_step": spp_step, "parallelism": "row-tiles x%d" % world, "rng": "counter (seed 1234)",
                       "l2": "wavefront state per wave (~%d MB) exceeds L2 (%d MB); no explicit flush" %
                             (176 * min((16 << 20), W * H * spp_step // world) >> 20, L2_BYTES >> 20)},
            "clocks": clk.summary(),
            "e2e": {"value": round(e2e_value, 3), "unit": "Msamples/s", "h2d_bytes_per_step": scene_upload_bytes,
                    "d2h_bytes_per_step": W * H * 4,
                    "caller_scene_bytes": scene_bytes(fs), "host_rebuilds": scene_rebuilds, "scene_uploads": scene_uploads,
                    "note": "per step: crt_update_scene (the caller's host arrays are fingerprinted; the derived 4-ary trees and "
                            "triangle records are rebuilt on the host only when the geometry changed - once here - and copied "
                            "host -> device from page-locked memory EVERY step), crt_render, crt_resolve_rgba8 into a page-locked "
                            "host frame"},
            "gpu_launches": int(launches),
            "stage_ms": stage,
            "tight_bvh": {"wide_nodes": tight_nodes, "split_triangles": tight_always, "pad": pad_mode, "per_ray": per_ray},
            "lane_stats": lane_stats,
            "roofline": {"bound": "hbm", "kernel": "k_traverse (BVH traversal of the extend and shadow rays)", "achieved": round(achieved, 2) if achieved else None,
                         "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": round(achieved / peak, 4) if achieved else None,
                         "traffic": traffic, "algorithmic_bytes_per_launch": round((rays / world) * a_ray / max(n_trav_launches, 1), 0),
                         "bytes_per_ray": round(a_ray, 1), "launches": n_trav_launches, "physical_limiter_ncu": physical,
                         "physical": None if not traffic else {
                             "dram_gbs": round(traffic * n_trav_launches / (trav_ms * 1e-3) / 1e9, 1),
                             "frac_of_peak": round(traffic * n_trav_launches / (trav_ms * 1e-3) / 1e9 / peak, 4),
                             "what": "ncu DRAM bytes per launch x launches / live CUDA-event traversal time"},
                         "note": "algorithmic bytes follow the REFERENCE traversal (no distance culling, SURVEY.md 8(d)); the scene is "
                                 "cache resident and this path culls, so frac can exceed 1 - ncu shows the kernel is issue-bound, not HBM-bound",
                         "reference_traversal": {k: round(od[k], 3) for k in ("rays_per_sample", "nodes_per_ray", "tris_per_ray")}},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(scene, W, H)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
