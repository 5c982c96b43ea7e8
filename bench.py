#!/usr/bin/env python3
"""bench.py - the hot path's headline metric (BASELINE.json): Msamples/s (and Mrays/s) of the per-pixel
path-tracing loop on N B200s, next to the reference's own --cpu path and the reference's own CUDA renderer
timed on the same box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME] [--scaling weak|strong]

A "step" is one pass of the hot path over one batch: `spp_per_step` samples for every pixel of the frame.
Default workload = BASELINE.json configs[1]: cornellSphere (refractive sphere, mirror + glossy boxes),
1920x1080, 512 spp = 8 steps x 64 spp.  For N > 1 (torchrun, one rank per GPU) the frame is sharded as
interleaved row tiles; `--scaling weak` (default) gives every rank N x 64 spp of its tiles per step (per-GPU work
fixed), `--scaling strong` keeps 64 spp of the whole frame per step.  Each step ends with the ONE collective of the
path: every rank resolves its tiles to RGBA8 and rank 0 collects them (NCCL send/recv, 4 bytes per pixel).

One JSON line on stdout (rank 0):
  value        whole-job Msamples/s, scene resident in HBM, CUDA events on the launching stream, no profiling flag;
  e2e          the same through the Renderer / C-ABI call with HOST buffers (scene upload + render + RGBA8 read-back
               into page-locked memory every step);
  seam_literal the reference launcher's literal loop: renderOneSamplePerPixel() once per sample, copyImageBytes() at the end;
  roofline     the dominant kernel (k_traverse) against the ceiling ncu says binds it: SM issue slots;
  cpu_baseline the reference's --cpu path on a bounded sample; cuda_reference the reference's own CUDA renderer on this GPU;
  configs      short runs of the other BASELINE configs (N = 1 only) and the config-5 frame under strong scaling.
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
    "cornellSphere_1080p": ("cornellSphere", 1920, 1080, 512),   # BASELINE config 2 (the headline)
    "cornell_1080p": ("cornell", 1920, 1080, 512),
    "monkeys_1080p": ("monkeys", 1920, 1080, 512),               # BASELINE config 3
    "cornell_4k": ("cornell", 3840, 2160, 1024),                 # BASELINE config 5 (1/2/4/8 GPU sweep)
    "cornell_256": ("cornell", 256, 256, 16),                    # BASELINE config 1 (the reference's CPU-runnable case)
    "terrain1m_1080p": ("terrain707", 1920, 1080, 256),          # BASELINE config 4: 999 700 procedural triangles
    "terrain131k_1080p": ("terrain256", 1920, 1080, 256),
    "terrain8k_256": ("terrain64", 256, 256, 16),                # small dynamic-pad scene for compute-sanitizer runs
}
SPP_PER_STEP = 64
L2_BYTES = 126 * 1024 * 1024
SM_COUNT = 148
ISSUE_SLOTS_PER_SM_CLK = 4          # one warp instruction per scheduler per clock, four schedulers per SM

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
        j = json.load(open(p))
        return float(j["hbm_gbs"]), float(j.get("sm_max_mhz", 1965.0)), "measured"
    return 6650.0, 1965.0, "fallback"


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
def _reference_binary(scene):
    """(binary, cwd) of the UNMODIFIED reference compiled by oracle/Makefile, with its .env pointing at the fixtures."""
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "CudaRT_ref")
    assets = os.path.join(ROOT, "oracle", "_ref", "assets")
    if not (os.path.exists(ref_bin) and os.path.exists(os.path.join(assets, "meshes", scene + ".obj"))):
        return None, None
    with open(os.path.join(assets, ".env"), "w") as f:
        f.write("MESHES_PATH=%s/meshes\nCAMERA_PATH=%s/settings\nTEXTURES_PATH=%s/textures\n" % (assets, assets, assets))
    cwd = os.path.join(assets, "build")
    os.makedirs(cwd, exist_ok=True)
    return ref_bin, cwd


def reference_cpu_run(scene, W, H, spp, threads):
    """Time the UNMODIFIED reference's --cpu path (oracle/_ref/CudaRT_ref) or, when it is not there, the oracle port.
    Returns (Msamples/s, kind)."""
    ref_bin, cwd = _reference_binary(scene)
    if ref_bin:
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


def reference_cuda_run(scene, W, H, spp, device_uuid_or_index):
    """The reference's OWN CUDA renderer (ParallelRenderer, parallel_renderer.cu:288-342: a megakernel, one launch per
    sample), unmodified, compiled for sm_100 by oracle/Makefile, on this GPU: `CudaRT_ref -f <scene> -s <spp> -w W -h H`
    without --cpu.  CUDA_VISIBLE_DEVICES pins the device (main.cpp:160-162 picks device 1 when it sees two).  It renders
    H - H % 16 rows (its 16x16 blocks truncate); the rate is quoted on the rows it rendered."""
    ref_bin, cwd = _reference_binary(scene)
    if not ref_bin:
        return None
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=str(device_uuid_or_index))
    try:
        out = subprocess.run([ref_bin, "-f", scene, "-s", str(spp), "-w", str(W), "-h", str(H)], cwd=cwd, env=env,
                             capture_output=True, text=True, timeout=900).stdout
    except (subprocess.TimeoutExpired, OSError):
        return None
    m = re.search(r"Elapsed time in seconds = ([0-9.eE+-]+)", out)
    if not m:
        return None
    t = float(m.group(1))
    rows = H - H % 16
    return {"value": round(W * rows * spp / t / 1e6, 3), "unit": "Msamples/s", "seconds": round(t, 3),
            "sample": "%s %dx%d x %d spp (renders %d of %d rows), its own 'Elapsed time'" % (scene, W, H, spp, rows, H),
            "what": "reference ParallelRenderer (parallel_renderer.cu:288-342), unmodified, -arch=sm_100, same GPU"}


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


def ncu_per_ray(workload):
    """Per-ray instruction and DRAM figures of k_traverse for this workload, from the committed ncu launch list of the
    final kernels (profiles/r02_traverse_per_ray.json, written by tools/ncu_per_ray.py): the thread- and warp-instruction
    counts are a property of the code and the ray set, so live rays/s times these gives live instructions/s."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_traverse_per_ray.json"))).get(workload)
    except (OSError, ValueError):
        return None


def workload_config(workload):
    """`config` of the JSON line: names the workload and nothing else, so that the two arms (`--impl b200` / `reference`)
    of one workload carry the IDENTICAL dict; how a run was carried out (samples per step, batch size, sharding, the CPU
    arm's bounded sample) is in `run`."""
    scene, W, H, total_spp = WORKLOADS[workload]
    return {"workload": workload, "scene": scene, "width": W, "height": H, "spp_total_config": total_spp,
            "l2": "the wavefront state of one batch (~3.4 GB at 16.6 M paths) exceeds the %d MB L2 many times over: no explicit "
                  "flush between timed iterations" % (L2_BYTES >> 20)}


def cli_multi_gpu_check(gpus):
    """`./CudaRT --gpus N` (one process, one context and one host thread per GPU, ncclCommInitAll + ncclSend/ncclRecv of the
    RGBA8 tiles to GPU 0) against `--gpus 1` on the hand-written `tiny` scene: the two PNG files must be byte-identical.
    Runs inside the multi-GPU bench so that this in-process NCCL path is exercised on the driver's >= 2-GPU boxes."""
    import hashlib
    import tempfile
    cli = os.path.join(ROOT, "cudart_b200", "CudaRT")
    tiny = os.path.join(ROOT, "tests", "golden", "tiny")
    if not os.path.exists(cli):
        return {"skipped": "cudart_b200/CudaRT is not built"}
    out = {"gpus": gpus, "scene": "tiny 70x45, 16 spp"}
    with tempfile.TemporaryDirectory() as tmp:
        with open(os.path.join(tmp, ".env"), "w") as f:
            f.write("MESHES_PATH=%s/meshes\nCAMERA_PATH=%s/settings\nTEXTURES_PATH=%s/textures\n" % (tiny, tiny, tiny))
        digests = []
        for n in (1, gpus):
            wd = os.path.join(tmp, "build%d" % n)
            os.makedirs(wd)
            env = {k: v for k, v in os.environ.items() if k not in ("CUDA_VISIBLE_DEVICES",)}
            t0 = time.perf_counter()
            try:
                p = subprocess.run([cli, "-f", "tiny", "-s", "16", "-w", "70", "-h", "45", "--gpus", str(n)], cwd=wd, env=env,
                                   capture_output=True, text=True, timeout=300)
            except subprocess.TimeoutExpired:
                return dict(out, error="timeout with --gpus %d" % n)
            if p.returncode != 0:
                return dict(out, error=(p.stderr or p.stdout)[-300:])
            out["seconds_gpus_%d" % n] = round(time.perf_counter() - t0, 2)
            digests.append(hashlib.sha256(open(os.path.join(wd, "tiny.png"), "rb").read()).hexdigest())
    out["identical_png"] = digests[0] == digests[1]
    out["sha256"] = digests[0][:16]
    return out


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
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload),
            "run": {"timed_sample": "%dx%d x %d spp per step, scaled linearly to the workload's frame" % (w, h, spp), "threads": threads},
            "cpu_baseline": {"value": round(value, 5), "unit": "Msamples/s", "cores": threads, "kind": kind,
                             "sample": "%s %dx%d %d spp per step, render time as printed by the reference binary" % (scene, w, h, spp)},
            "e2e": {"value": round(value, 5), "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
class Harness:
    """One process = one rank = one GPU: torch for the device stream, events and NCCL; everything else is the C ABI."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank, self.world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            sys.exit("bench.py: no CUDA device - the B200 path has no CPU fallback")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        # a real (non-default) stream: the context enqueues on it, torch's events and collectives are recorded on it
        self.stream = torch.cuda.Stream()
        torch.cuda.set_stream(self.stream)
        try:
            self.uuid = "GPU-" + str(torch.cuda.get_device_properties(self.local).uuid)
        except Exception:
            self.uuid = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def renderer(self, fs, W, H, spp, **kw):
        from cudart_b200.renderer import B200Renderer
        a = self.args
        return B200Renderer(fs, W, H, spp, device=self.local, rank=self.rank, world=self.world, stream=self.stream.cuda_stream,
                            spp_per_wave=a.spp_per_wave, smem_nodes=False if a.no_smem_nodes else None,
                            tight_bvh={"auto": None, "on": True, "off": False}[a.tight_bvh],
                            dynamic_pad={"auto": None, "static": False, "dynamic": True}[a.pad],
                            waves_in_flight=a.waves_in_flight, **kw)

    def device_step(self, r, spp_step):
        """Render + the path's one collective, everything on the device (no host synchronisation, no host copies)."""
        r.render(spp_step)
        if self.world > 1:
            ptr, nbytes = r.resolve_local_rgba8_device()
            from cudart_b200.sharding import as_torch
            local = as_torch(ptr, (nbytes,), "|u1")
            if self.rank == 0:
                if getattr(self, "_gathered", None) is None or self._gathered.shape != (self.world, nbytes):
                    self._gathered = self.torch.empty((self.world, nbytes), dtype=self.torch.uint8, device="cuda")
                self.dist.gather(local, list(self._gathered.unbind(0)), dst=0)
            else:
                self.dist.gather(local, None, dst=0)

    def timed(self, fn, steps, clocks=False):
        """EXACTLY `steps` calls of fn between two events on the launching stream, barrier + synchronize on both sides;
        returns (max over ranks of the elapsed ms, clock summary or None)."""
        torch = self.torch
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        clk = ClockSampler(self.local, self.uuid) if clocks else None
        if clk:
            clk.__enter__()
        self.barrier()
        ev0.record(self.stream)
        for _ in range(steps):
            fn()
        ev1.record(self.stream)
        self.barrier()
        if clk:
            clk.__exit__()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item()), (clk.summary() if clk else None)

    def reduce(self, x, op="sum"):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([float(x)], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM if op == "sum" else self.dist.ReduceOp.MAX)
        return float(t.item())


def measure_workload(h, workload, steps, warmup, spp_step, full=True):
    """value (+ e2e, seam, stage shares when `full`) of one workload on this harness; dict on every rank."""
    scene, W, H, total_spp = WORKLOADS[workload]
    fs = load_scene(scene, W, H)
    out = {"workload": workload, "scene": scene, "width": W, "height": H, "spp_per_step": spp_step}
    samples = float(W) * H * spp_step * steps

    # ---- value: scene resident, no profiling flag, device-only step ----
    t_create = time.perf_counter()
    r = h.renderer(fs, W, H, total_spp)
    # crt_create as the caller sees it: validation, fingerprint, the host-side derivation of the trees and records on the
    # host's threads, device allocations and the enqueued uploads
    out["create_s"] = round(time.perf_counter() - t_create, 4)
    # warm-up: at least `warmup` steps AND at least 0.3 s of work (a 256x256 step lasts 2 ms: three of them would leave
    # the clocks of a cold GPU still ramping)
    t_w, n_w = time.perf_counter(), 0
    while n_w < warmup or (h.world == 1 and time.perf_counter() - t_w < 0.3 and n_w < 400):   # N > 1: same count on every rank
        h.device_step(r, spp_step)
        n_w += 1
        if n_w % 8 == 0:
            h.torch.cuda.synchronize()
    h.barrier()
    c0 = r.counters()
    ms, clocks = h.timed(lambda: h.device_step(r, spp_step), steps, clocks=True)
    c1 = r.counters()
    rays = h.reduce(c1["extend_rays"] - c0["extend_rays"] + c1["shadow_rays"] - c0["shadow_rays"])
    out.update(value=samples / (ms * 1e-3) / 1e6, ms_per_step=ms / steps, mrays_per_s=rays / (ms * 1e-3) / 1e6,
               rays_per_sample=rays / samples, clocks=clocks, gpu_launches=int(c1["kernel_launches"] - c0["kernel_launches"]),
               spp_per_wave=r.spp_per_wave(), tight_bvh={"wide_nodes": r.tight_info()[0], "split_triangles": r.tight_info()[1],
                                                         "pad": ("off", "static", "dynamic")[r.tight_pad_mode()]})
    r.close()
    if not full:
        return out

    # ---- stage shares and the traversal kernel's own time: a second pass with an event after every launch ----
    rp = h.renderer(fs, W, H, total_spp, profile=True, stats=h.args.stats)
    for _ in range(2):
        h.device_step(rp, spp_step)
    h.barrier()
    p0 = rp.counters()
    pms, _ = h.timed(lambda: h.device_step(rp, spp_step), steps)
    p1 = rp.counters()
    stage = {k: p1["stage_ms"][k] - p0["stage_ms"][k] for k in p1["stage_ms"]}
    out.update(stage_ms={k: round(v, 3) for k, v in stage.items()}, profiled_value=samples / (pms * 1e-3) / 1e6,
               trav_ms=h.reduce(stage["extend"] + stage["shadow"], "max"),
               trav_rays=h.reduce(p1["extend_rays"] - p0["extend_rays"] + p1["shadow_rays"] - p0["shadow_rays"]),
               trav_launches=7 * (int(p1["kernel_launches"] - p0["kernel_launches"]) // 15))
    if h.args.stats and h.world == 1:
        out["per_ray"] = {"wide_nodes_per_ray": round((p1["node_visits"] - p0["node_visits"]) / max(out["trav_rays"], 1), 2),
                          "tri_tests_per_ray": round((p1["tri_tests"] - p0["tri_tests"]) / max(out["trav_rays"], 1), 2)}
        out["lane_stats"] = rp.lane_stats()
    rp.close()

    # ---- e2e through the public API with HOST buffers: every step uploads the scene (host -> device), renders
    # spp_step samples per pixel and reads the RGBA8 frame back (device -> host), like one frame of an animation ----
    torch = h.torch
    rr = h.renderer(fs, W, H, spp_step)
    rr.h_imgPtr = torch.empty((H, W, 4), dtype=torch.uint8, pin_memory=True).numpy()   # page-locked: no staging copy

    def e2e_step():
        from cudart_b200.sharding import gather_rgba8
        rr.update_scene(fs)
        rr.render(spp_step)
        if h.world > 1:
            gather_rgba8(rr)
        else:
            rr.copyImageBytes()

    e2e_step()
    h.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    h.barrier()
    e2e_s = h.reduce(time.perf_counter() - t0, "max")
    rebuilds, uploads, upload_bytes = rr.scene_info()
    out["e2e"] = {"value": round(samples / e2e_s / 1e6, 3), "unit": "Msamples/s", "h2d_bytes_per_step": upload_bytes,
                  "d2h_bytes_per_step": W * H * 4 if h.rank == 0 else 0, "caller_scene_bytes": scene_bytes(fs),
                  "host_rebuilds": rebuilds, "scene_uploads": uploads,
                  "note": "per step: crt_update_scene (the caller's host arrays are fingerprinted; the derived 4-ary trees and "
                          "triangle records are rebuilt on the host only when the geometry changed - once here - and copied "
                          "host -> device from page-locked memory EVERY step), crt_render, and the RGBA8 frame into a page-locked "
                          "host buffer (N > 1: per-rank resolve, NCCL gather to rank 0, de-interleave there)"}

    # ---- the literal seam: TerminalLauncher::render's loop (terminal_launcher.cpp:10-16), one call per sample ----
    def seam_step():
        for _ in range(spp_step):
            rr.renderOneSamplePerPixel(rr.h_imgPtr)

    rr.reset()
    seam_step()
    seam_ms, _ = h.timed(seam_step, steps)
    out["seam_literal"] = {"value": round(samples / (seam_ms * 1e-3) / 1e6, 3), "unit": "Msamples/s",
                           "what": "renderOneSamplePerPixel() x %d per step through the Renderer mirror (crt_render(ctx, 1) each); "
                                   "the library queues single samples and launches whole %d-spp waves" % (spp_step, rr.spp_per_wave())}

    # ---- time to the workload's full sample count, the way a user would get the frame: upload, render, read back ----
    if h.world == 1 or h.args.scaling == "strong":
        rr.reset()
        h.barrier()
        t0 = time.perf_counter()
        rr.update_scene(fs)
        rr.render(total_spp)
        if h.world > 1:
            from cudart_b200.sharding import gather_rgba8
            gather_rgba8(rr)
        else:
            rr.copyImageBytes()
        h.barrier()
        out["time_to_full_spp_s"] = {"spp": total_spp, "seconds": round(h.reduce(time.perf_counter() - t0, "max"), 4)}
    rr.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cornellSphere_1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = N x spp-per-step per step (per-GPU work fixed), strong = spp-per-step of the whole frame per step")
    ap.add_argument("--spp-per-step", type=int, default=SPP_PER_STEP)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the short runs of the other BASELINE configs and the reference CUDA renderer")
    ap.add_argument("--no-smem-nodes", action="store_true", help="A/B: never stage BVH nodes in shared memory")
    ap.add_argument("--tight-bvh", default="auto", choices=["auto", "on", "off"], help="A/B: per-triangle tight BVH (DESIGN.md 2.4)")
    ap.add_argument("--pad", default="auto", choices=["auto", "static", "dynamic"], help="A/B: where the tight BVH's boxes are inflated (DESIGN.md 2.4)")
    ap.add_argument("--stats", action="store_true", help="count node visits / triangle tests in the profiling pass")
    ap.add_argument("--spp-per-wave", type=int, default=0, help="samples per pixel per wavefront batch (0 = auto, ~16M paths)")
    ap.add_argument("--waves-in-flight", type=int, default=0, help="wavefront batches in flight (0 = library default)")
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

    # stdout carries exactly ONE line, the JSON: everything else that writes to file descriptor 1 (NCCL's version banner,
    # library chatter) is sent to stderr from here on
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    h = Harness(args)
    world, rank = h.world, h.rank
    warmup = max(args.warmup, 3)
    spp_step = args.spp_per_step * (world if args.scaling == "weak" else 1)
    m = measure_workload(h, args.workload, args.steps, warmup, spp_step)

    # ---- the other BASELINE configs, short (same code path, 3 steps each) ----
    configs = {}
    if not args.no_configs and args.workload == "cornellSphere_1080p":
        if world == 1:
            for wl in ("monkeys_1080p", "terrain1m_1080p", "cornell_4k", "cornell_256"):
                c = measure_workload(h, wl, 3, 3, args.spp_per_step, full=False)
                configs[wl] = {"value": round(c["value"], 3), "unit": "Msamples/s", "mrays_per_s": round(c["mrays_per_s"], 3),
                               "ms_per_step": round(c["ms_per_step"], 3), "spp_per_step": c["spp_per_step"], "steps": 3,
                               "tight_bvh": c["tight_bvh"], "create_s": c["create_s"], "scaling": "n/a (1 GPU)"}
        else:
            # BASELINE config 5 under STRONG scaling: the fixed 3840x2160 frame, 64 spp of the whole frame per step, shared
            # by the N ranks, RGBA8 gather to rank 0 every step
            c = measure_workload(h, "cornell_4k", 3, 3, args.spp_per_step, full=False)
            configs["cornell_4k"] = {"value": round(c["value"], 3), "unit": "Msamples/s", "mrays_per_s": round(c["mrays_per_s"], 3),
                                     "ms_per_step": round(c["ms_per_step"], 3), "spp_per_step": c["spp_per_step"], "steps": 3,
                                     "scaling": "strong", "n_gpus": world,
                                     "time_to_1024spp_s_extrapolated": round(c["ms_per_step"] * 1e-3 * 1024 / c["spp_per_step"], 3)}

    if rank == 0:
        hbm_peak, sm_mhz, peak_src = measured_peaks()
        a_ray, od = algorithmic_bytes(scene, W, H)
        steps = args.steps
        samples = float(W) * H * spp_step * steps
        trav_s = m["trav_ms"] * 1e-3
        rays_rank = m["trav_rays"] / world
        launches = max(m["trav_launches"], 1)
        issue_peak = SM_COUNT * ISSUE_SLOTS_PER_SM_CLK * sm_mhz * 1e6 / 1e9           # G warp instructions / s
        per_ray = ncu_per_ray(args.workload)
        roof = {"bound": "issue", "kernel": "k_traverse (BVH traversal of the extend and shadow rays)", "unit": "Gwarp-inst/s",
                "peak": round(issue_peak, 1),
                "peak_source": "%d SMs x %d schedulers x 1 warp instruction per clock x %.0f MHz (%s SM clock)" %
                               (SM_COUNT, ISSUE_SLOTS_PER_SM_CLK, sm_mhz, peak_src),
                "launches": launches, "kernel_ms_per_launch": round(m["trav_ms"] / launches, 4),
                "kernel_time_from": "CUDA events around every launch on the launching stream, in a second pass of the same steps "
                                    "(CRT_FLAG_PROFILE; that pass ran at %.1f Msamples/s against value %.1f)" % (m["profiled_value"], m["value"]),
                "rays_per_launch": round(rays_rank / launches, 1)}
        if per_ray and trav_s > 0:
            winst = rays_rank * per_ray["warp_inst_per_ray"]
            roof.update(achieved=round(winst / trav_s / 1e9, 1), frac=round(winst / trav_s / 1e9 / issue_peak, 4),
                        lanes_per_inst=per_ray["thread_inst_per_ray"] / per_ray["warp_inst_per_ray"],
                        thread_inst_frac_of_simt_peak=round(rays_rank * per_ray["thread_inst_per_ray"] / trav_s / 1e9 / (issue_peak * 32), 4),
                        algorithmic_inst_per_launch=round(winst / launches, 0),
                        per_ray_source=per_ray.get("source"), traffic=round(per_ray["dram_bytes_per_ray"] * rays_rank / launches, 0),
                        traffic_what="dram__bytes_read.sum + dram__bytes_write.sum of the k_traverse launches of one wave (ncu), per ray, x live rays per launch")
            roof["lanes_per_inst"] = round(roof["lanes_per_inst"], 2)
        else:
            roof.update(achieved=None, frac=None, traffic=None)
        # the contract's algorithm-defined HBM figure stays as a side field (it exceeds what is physically moved: this path culls)
        if trav_s > 0:
            alg = rays_rank * a_ray / trav_s / 1e9
            roof["hbm_algorithmic"] = {"bytes_per_ray": round(a_ray, 1), "achieved_gbs": round(alg, 1), "peak_gbs": hbm_peak, "peak_source": peak_src,
                                       "frac": round(alg / hbm_peak, 4),
                                       "note": "SURVEY.md 8(d) A_ray of the REFERENCE traversal (no distance culling); the scene is cache "
                                               "resident and this path culls, so this is not a physical bandwidth",
                                       "reference_traversal": {k: round(od[k], 3) for k in ("rays_per_sample", "nodes_per_ray", "tris_per_ray")}}
            if per_ray:
                phys = per_ray["dram_bytes_per_ray"] * rays_rank / trav_s / 1e9
                roof["hbm_physical"] = {"dram_gbs": round(phys, 1), "frac_of_peak": round(phys / hbm_peak, 4),
                                        "queue_bytes_per_ray": 48, "what": "ncu DRAM bytes per ray x live rays / live kernel time"}
        line = {
            "metric": "path-tracing throughput", "value": round(m["value"], 3), "unit": "Msamples/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": round(m["ms_per_step"], 3),
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "mrays_per_s": round(m["mrays_per_s"], 3), "rays_per_sample": round(m["rays_per_sample"], 4),
            "config": workload_config(args.workload),
            "run": {"spp_per_step": spp_step, "spp_per_wave": m["spp_per_wave"], "parallelism": "row-tiles x%d" % world,
                    "rng": "counter (seed 1234)"},
            "clocks": m["clocks"],
            "e2e": m["e2e"],
            "seam_literal": m["seam_literal"],
            "time_to_full_spp_s": m.get("time_to_full_spp_s"),
            "time_to_512spp_s": (m.get("time_to_full_spp_s") or {}).get("seconds") if total_spp == 512 else None,
            "gpu_launches": m["gpu_launches"],
            "create_s": m["create_s"],
            "stage_ms": m["stage_ms"],
            "tight_bvh": dict(m["tight_bvh"], per_ray=m.get("per_ray")),
            "lane_stats": m.get("lane_stats"),
            "roofline": roof,
            "configs": configs or None,
        }
        if world == 1 and not args.no_configs and args.workload == "cornellSphere_1080p":
            line["cuda_reference"] = reference_cuda_run(scene, W, H, 64, h.uuid or h.local)
        if world > 1 and not args.no_configs:
            line["cli_multi_gpu"] = cli_multi_gpu_check(world)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(scene, W, H)
            if m.get("time_to_full_spp_s"):
                line["time_to_full_spp_s"]["cpu_reference_seconds_extrapolated"] = round(W * H * total_spp / (line["cpu_baseline"]["value"] * 1e6), 1)
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        h.dist.barrier()
        h.dist.destroy_process_group()


if __name__ == "__main__":
    main()
