#!/usr/bin/env python
"""Benchmark of the sightpy hot path on B200 (BASELINE.json: Mrays/s and s/frame).

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # headline: Cornell box 1920x1080, 256 spp
    python bench.py --config {example1,example2,example3,example4,cornell,stress} [...]
    python bench.py --impl reference [--config ...]                     # CPU arm: the real reference on the host cores
    torchrun --nproc-per-node N ... bench.py --gpus N ...               # one rank per GPU

One *step* renders the whole frame of the configuration once (primary-ray generation, intersection, shading,
importance-sampled bounces, accumulation, tonemap).  With N ranks the frame is cut into N shards
(sightpy/parallel.py: contiguous sample ranges, or interleaved 64x64 tiles for the stress scene / --shard tiles),
the float accumulation buffers are summed on rank 0 with one NCCL reduce and rank 0 tonemaps: total work is fixed,
i.e. strong scaling.

Metric: Mrays/s = rays traced against the collider list (primary + secondary, the reference's sum of len(ray) over
get_raycolor calls; shadow rays not counted) / second.
  value  scene already resident on the GPU, frame resolved on the device (no copy-out)
  e2e    Scene.render() through the public API: scene flattened + uploaded from host memory and the uint8 frame
         copied back to the host inside the timed region, every step.
The default run also times the other five BASELINE.json configurations briefly (`configs` block of the JSON line) and
every line carries a checksum of the frame rank 0 ends up with (`frame`: sha256 of the bytes, mean radiance and 4x4
block means), so that runs at 1/2/4/8 GPUs can be seen to produce the same image.
"""
import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
for p in (REPO, REPO / "python-raytracer_b200", REPO / "tests"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

# BASELINE.json configs, SURVEY.md §8(d) concrete inputs.  cpu: (width, height, spp) of the bounded CPU sample.
CONFIGS = {
    "example1": dict(index=0, builder="example1", width=400, height=300, spp=1, kw={}, cpu=(400, 300),
                     what="example1.py scene (2 glossy spheres, checker floor, sky box), 400x300, 1 spp"),
    "example2": dict(index=1, builder="example2", width=1920, height=1080, spp=7, kw={}, cpu=(240, 135),
                     what="example2.py scene (3 absorbing glass spheres, textured glossy floor), 1920x1080, 7 spp"),
    "example3": dict(index=1, builder="example3", width=1920, height=1080, spp=4, kw={}, cpu=(240, 135),
                     what="example3.py scene (rotated glass cuboid, textured floor), 1920x1080, 4 spp"),
    "example4": dict(index=2, builder="example4", width=3840, height=2160, spp=16, kw={}, cpu=(240, 135),
                     what="example4.py scene (thin-film bubble, blurred light-emitting sky box), 3840x2160, 16 spp"),
    "cornell": dict(index=3, builder="cornell", width=1920, height=1080, spp=256, kw={}, cpu=(160, 90),
                    what="example_cornellbox.py scene, 1920x1080, 256 spp"),
    # the reference's cost grows with (colliders x distinct colliders hit per level): at 6145 colliders it does not finish
    # a thumbnail in minutes, so its CPU sample uses a 385-collider version of the scene (and says so)
    "stress": dict(index=4, builder="stress", width=3840, height=2160, spp=64, kw={}, cpu=(24, 14), shard="tiles",
                   cpu_kw=dict(n_spheres=256, n_triangles=64),
                   what="4096 random spheres + 2 x 1024 triangles over a checker ground, 3840x2160, 64 spp"),
}
# algorithmic flops of one ray-collider test (1 FMA = 2 flop; compares / min / max not counted)
FLOP_PER_TEST = {0: 21, 1: 35, 2: 55, 3: 45}                # sphere, bounded plane, oriented cuboid, triangle
METRIC = "Mrays/sec (primary+secondary), Cornell box 1920x1080 256spp"


def metric_name(cfg_name):
    return METRIC if cfg_name == "cornell" else f"Mrays/sec (primary+secondary), {CONFIGS[cfg_name]['what']}"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="cornell", choices=sorted(CONFIGS))
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--spp", type=int, default=0)
    ap.add_argument("--shard", default="", choices=["", "samples", "tiles"], help="multi-GPU sharding (default: per config)")
    ap.add_argument("--chunk", type=int, default=0, help="primaries per wavefront chunk (0 = library default)")
    ap.add_argument("--queue-cap", type=int, default=0, help="records per wavefront queue (0 = library default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the brief runs of the other configurations")
    ap.add_argument("--in-process", action="store_true",
                    help="drive --gpus N devices from this one process through sp_render_group (no torchrun)")
    return ap.parse_args()


# ---- clocks --------------------------------------------------------------------------------------
class ClockSampler:
    """Samples nvidia-smi during the timed region (B200_PROFILING.md 'clocks line')."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    PERIOD_MS = 50

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None
        self.t_begin = self.t_end = None

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def mark_end(self):
        self.t_end = time.perf_counter()

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", str(self.PERIOD_MS)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
            self.thread.join(timeout=2)

    def summary(self):
        """Median SM clock and throttle reasons over the samples taken inside the timed region (the sampler runs
        from before the warm-up, so nvidia-smi is already up when a short timed region starts).  A region
        shorter than the sampling period falls back to the sample nearest to it and says so."""
        lo = self.t_begin if self.t_begin is not None else float("-inf")
        hi = (self.t_end if self.t_end is not None else float("inf")) + 1e-3 * self.PERIOD_MS
        parsed = []
        for t, r in self.rows:
            try:
                parsed.append((t, float(r[0]), float(r[1]), r[3:7], float(r[2]) if r[2].replace('.', '', 1).isdigit() else None))
            except (ValueError, IndexError):
                continue
        inside = [p for p in parsed if lo <= p[0] <= hi]
        note = None
        if not inside and parsed:
            mid = 0.5 * (lo + hi) if self.t_begin is not None and self.t_end is not None else parsed[-1][0]
            inside = [min(parsed, key=lambda p: abs(p[0] - mid))]
            note = "timed region shorter than the sampling period: nearest sample (%.0f ms away)" % (abs(inside[0][0] - mid) * 1e3)
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        reasons = set()
        for _, _, _, flags, _ in inside:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), flags):
                if v.lower().startswith("active"):
                    reasons.add(name)
        out = {"sm_mhz": statistics.median(p[1] for p in inside), "sm_max_mhz": max(p[2] for p in inside),
               "reasons": sorted(reasons), "samples": len(inside)}
        watts = [p[4] for p in inside if p[4] is not None]
        if watts:
            out["power_w_median"], out["power_w_max"] = statistics.median(watts), max(watts)
        if note:
            out["note"] = note
        return out


# ---- CPU arms --------------------------------------------------------------------------------------
# The real reference (oracle/reference_arm.py: the unmodified package from the git-ignored baseline/_ref copy) when it
# travelled with the repository, else the float64 numpy oracle port (pinned to the reference at 1e-16).
def _oracle_worker(job):
    builder, kw, sample, seed = job
    import scenes
    import sightpy
    from oracle.sightpy_oracle import Oracle
    from sightpy.flatten import flatten_scene
    flat = flatten_scene(scenes.BUILDERS[builder](sightpy, **kw))
    orc = Oracle(flat, rng="philox", seed=seed)
    t0 = time.perf_counter()
    orc.render_linear(1, sample_begin=sample)
    return orc.rays_total, time.perf_counter() - t0


def port_rate(builder, kw, n_samples, processes):
    """Oracle port: n_samples samples of the scene's frame over `processes` workers.  Returns (rays, seconds)."""
    jobs = [(builder, kw, s, 0) for s in range(n_samples)]
    t0 = time.perf_counter()
    if processes == 1:
        res = [_oracle_worker(j) for j in jobs]
    else:
        import multiprocessing as mp
        small = dict(kw, width=max(kw["width"] // 4, 8), height=max(kw["height"] // 4, 8))
        with mp.get_context("spawn").Pool(processes) as pool:
            pool.map(_oracle_worker, [(builder, small, 0, 0)] * processes)   # spin-up (imports) outside the timing
            t0 = time.perf_counter()
            res = pool.map(_oracle_worker, jobs)
    return sum(r for r, _ in res), time.perf_counter() - t0


def cpu_rate(builder, kw, n_samples, processes):
    """-> (rays, seconds, kind)"""
    from oracle import reference_arm
    if reference_arm.available():
        r, s = reference_arm.trace_rate(builder, kw, n_samples, processes)
        return r, s, "reference"
    r, s = port_rate(builder, kw, n_samples, processes)
    return r, s, "port"


def run_reference(args, emit=print):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import reference_arm
    cfg = CONFIGS[args.config]
    cores = os.cpu_count() or 1
    w, h = cfg["cpu"]
    if args.config == "cornell":
        w, h = 240, 135                   # 1/8-scale frame: ~1.9 M rays per sample (~6 s per core)
    kw = dict(cfg["kw"], width=w, height=h, **cfg.get("cpu_kw", {}))
    small = dict(kw, width=max(w // 4, 8), height=max(h // 4, 8))
    for _ in range(args.warmup):
        cpu_rate(cfg["builder"], small, cores, cores)
    rays = secs = 0.0
    kind = "port"
    for _ in range(args.steps):
        r, s, kind = cpu_rate(cfg["builder"], kw, cores, cores)
        rays += r; secs += s
    value = rays / secs / 1e6
    what = ("get_raycolor of the unmodified reference (baseline/_ref copy of lmondada/Python-Raytracer, float64 numpy)"
            if kind == "reference" else "float64 numpy oracle port of the reference")
    sample = f"{cfg['builder']} scene{' ' + str(cfg['cpu_kw']) if cfg.get('cpu_kw') else ''} at {w}x{h}, {cores} spp per step (one sample per worker process), {what}"
    # Scene.render exactly as the example scripts call it: process pool, deep copies and pickling included
    shipped = None
    if kind == "reference" and args.config == "cornell":
        try:
            rpp = rays / (args.steps * cores * w * h)                     # rays per primary, measured above
            secs_shipped = reference_arm.render_as_shipped("cornell", dict(width=100, height=100), cores)
            shipped = {"s_per_frame": secs_shipped, "workload": f"example_cornellbox.py as shipped: 100x100, {cores} spp, Scene.render",
                       "Mrays_per_s": rpp * 100 * 100 * cores / secs_shipped / 1e6}
        except Exception as e:  # noqa: BLE001
            shipped = {"error": f"{type(e).__name__}: {e}"}
    emit(json.dumps({
        "impl": "reference", "metric": metric_name(args.config), "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{cfg['what']} (BASELINE.json configs[{cfg['index']}]; timed on a bounded sample)",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample},
        "render_as_shipped": shipped,
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ---- native arm --------------------------------------------------------------------------------------
def frame_fingerprint(srgb, lin):
    """sha256 of the uint8 frame plus numbers that survive a different order of float additions."""
    c, h, w = lin.shape
    hb, wb = max(h // 4, 1), max(w // 4, 1)
    blocks = [[round(float(lin[:, y * hb:(y + 1) * hb if y < 3 else h, x * wb:(x + 1) * wb if x < 3 else w].mean()), 6)
               for x in range(4)] for y in range(4)]
    return {"sha256": hashlib.sha256(np.ascontiguousarray(srgb).tobytes()).hexdigest()[:16],
            "mean_radiance": round(float(lin.mean()), 6), "mean_srgb8": round(float(srgb.mean()), 4),
            "block_means_4x4": blocks,
            "note": "float atomics add in a run-dependent order: bytes may differ in a few pixels, the means agree to ~1e-6"}


class Runner:
    """One configuration resident on this rank's GPU."""

    def __init__(self, torch, dist, name, width, height, spp, shard, world, rank, chunk=0, queue_cap=0):
        import scenes
        import sightpy
        from sightpy import parallel
        from sightpy.backend import NativeScene
        from sightpy.flatten import flatten_scene
        self.torch, self.dist, self.world, self.rank = torch, dist, world, rank
        cfg = CONFIGS[name]
        self.name, self.cfg, self.width, self.height, self.spp = name, cfg, width, height, spp
        t0 = time.perf_counter()
        self.scene = scenes.BUILDERS[cfg["builder"]](sightpy, width=width, height=height, **cfg["kw"])
        self.flat = flatten_scene(self.scene)
        self.host_build_s = time.perf_counter() - t0
        self.native = NativeScene(self.flat)
        if chunk:
            self.native.set_option("chunk_primaries", chunk)
        if queue_cap:
            self.native.set_option("ray_queue_capacity", queue_cap)
            self.native.set_option("fan_queue_capacity", queue_cap)
        self.stream = torch.cuda.current_stream()
        self.native.set_stream(self.stream.cuda_stream)
        self.shard = shard or cfg.get("shard", "samples")
        if self.shard == "samples" and spp < world:
            self.shard = "tiles"
        self.range = parallel.sample_range(spp, rank, world)
        self.tiles = parallel.tile_ids(width, height, rank, world)
        self.acc = parallel.accum_as_tensor(self.native) if world > 1 else None

    def step(self):
        """Device-resident frame: this rank's shard, reduce, resolve on the device."""
        if self.shard == "samples":
            st = self.native.render_samples(self.range[0], self.range[1], seed=0, clear=True)
        else:
            st = self.native.render_tiles(self.tiles, 64, 0, self.spp, seed=0, clear=True)
        if self.world > 1:
            self.dist.reduce(self.acc, dst=0, op=self.dist.ReduceOp.SUM)
        if self.rank == 0:
            self.native.resolve_on_device(self.spp)
        return st

    def flop_per_ray(self):
        types = self.flat.colliders["type"]
        return int(sum(FLOP_PER_TEST[int(t)] for t in types))

    def fingerprint(self):
        srgb, lin = self.native.resolve(self.spp, want_linear=True)
        return frame_fingerprint(srgb, lin)

    def close(self):
        self.native.close()


def timed(torch, dist, world, runner, warmup, steps, clocks=None):
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    tot = {"rays": 0, "launches": 0, "level_ms": 0.0, "level_launches": 0, "queue_bytes": 0, "shadow": 0, "chunks": 0,
           "retries": 0, "level_ms_by_depth": None, "per_depth": None}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(warmup):
        runner.step()
    barrier()
    if clocks:
        clocks.mark_begin()
    t_wall = time.perf_counter()
    ev0.record(runner.stream)
    for _ in range(steps):
        st = runner.step()
        tot["rays"] += st["rays_total"]; tot["launches"] += st["kernel_launches"] + (1 if runner.rank == 0 else 0)
        tot["level_ms"] += st["level_kernel_ms"]; tot["level_launches"] += st["level_kernel_launches"]
        tot["queue_bytes"] += st["queue_bytes"]; tot["shadow"] += st["shadow_rays"]; tot["chunks"] += st["chunks"]
        tot["retries"] += st["chunk_retries"]
        tot["per_depth"] = st["rays_per_depth"]; tot["level_ms_by_depth"] = st["level_ms"]
    ev1.record(runner.stream)
    barrier()
    tot["wall_s"] = time.perf_counter() - t_wall
    if clocks:
        clocks.mark_end()
    tot["dev_ms"] = ev0.elapsed_time(ev1)
    return tot


def run_group(args, emit=print):
    """N GPUs driven by the library itself from one process (sp_init_devices + sp_render_group: host thread per device,
    NVLink peer-access gather) — what a plain `python example.py` gets on a multi-GPU node."""
    import scenes
    import sightpy
    from sightpy.backend import NativeGroup
    from sightpy.flatten import flatten_scene
    cfg = CONFIGS[args.config]
    width, height, spp = args.width or cfg["width"], args.height or cfg["height"], args.spp or cfg["spp"]
    shard = args.shard or cfg.get("shard", "samples")
    flat = flatten_scene(scenes.BUILDERS[cfg["builder"]](sightpy, width=width, height=height, **cfg["kw"]))
    group = NativeGroup(flat, list(range(args.gpus)))
    with ClockSampler(0) as clocks:
        for _ in range(args.warmup):
            group.render_on_device(spp, 0, shard)
        clocks.mark_begin()
        t0 = time.perf_counter()
        rays = launches = 0
        dev_ms = 0.0
        for _ in range(args.steps):
            st = group.render_on_device(spp, 0, shard)
            rays += st["rays_total"]; launches += st["kernel_launches"]; dev_ms += st["device_ms"]
        wall = time.perf_counter() - t0
        clocks.mark_end()
    srgb, lin, st = group.render(spp, 0, want_linear=True, shard=shard)
    # end to end: the frame copied out every step (the scene stays resident: sp_render_group has no re-upload to time)
    t0 = time.perf_counter()
    n_e2e = max(1, min(args.steps, 2))
    e2e_rays = 0
    for _ in range(n_e2e):
        _, _, se = group.render(spp, 0, want_linear=False, shard=shard)
        e2e_rays += se["rays_total"]
    e2e_s = time.perf_counter() - t0
    group.close()
    emit(json.dumps({
        "metric": metric_name(args.config), "value": rays / wall / 1e6, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "s_per_frame": wall / args.steps,
        "device_ms_per_step_slowest_gpu": dev_ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{cfg['builder']} scene, {width}x{height}, {spp} spp (BASELINE.json configs[{cfg['index']}])",
                   "sharding": f"in-process: sp_render_group over {args.gpus} devices, {shard}, peer-access gather (no torchrun, no NCCL)",
                   "timing": "host wall clock around the blocking sp_render_group calls (slowest device + gather + resolve)",
                   "rays_per_frame": rays / args.steps},
        "frame": frame_fingerprint(srgb, lin),
        "e2e": {"value": e2e_rays / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": int(width * height * 3), "steps": n_e2e, "s_per_frame": e2e_s / n_e2e},
        "gpu_launches": int(launches), "clocks": clocks.summary(),
    }))


def run_native(args, emit=print):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ["SIGHTPY_DEVICE"] = str(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from sightpy.backend import measure_peaks

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        return t.item()

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cfg = CONFIGS[args.config]
    width, height, spp = args.width or cfg["width"], args.height or cfg["height"], args.spp or cfg["spp"]
    runner = Runner(torch, dist, args.config, width, height, spp, args.shard, world, rank, args.chunk, args.queue_cap)
    with ClockSampler(local) as clocks:                  # started before the warm-up, samples filtered to the timed region
        tot = timed(torch, dist, world, runner, args.warmup, args.steps, clocks)
    clock_summary = clocks.summary()
    job_ms = allmax(tot["dev_ms"])
    job_rays = allsum(tot["rays"])
    job_launches = int(allsum(tot["launches"]))
    value = job_rays / (job_ms * 1e-3) / 1e6
    fingerprint = runner.fingerprint() if rank == 0 else None

    # ---- end to end through the public API ---------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        scene, flat = runner.scene, runner.flat
        scene.seed = 0
        e2e_rays = 0
        if world > 1:
            os.environ["SIGHTPY_SHARD"] = runner.shard

        def e2e_step():
            scene.invalidate(full=True)              # forget the device copy: flatten + upload everything again
            img = scene.render(samples_per_pixel=spp)
            return img, scene.last_stats["rays_total"]

        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            e2e_step()
            barrier()
            t0 = time.perf_counter()
            n_e2e = max(1, min(args.steps, 2))
            for _ in range(n_e2e):
                img, r = e2e_step()
                e2e_rays += r
            barrier()
            e2e_s = time.perf_counter() - t0
        e2e_s = allmax(e2e_s)
        h2d = sum(getattr(flat, k).nbytes for k in ("materials", "primitives", "colliders", "lights", "importance",
                                                      "shadow_colliders", "media", "ambient")) + flat.camera.nbytes \
            + sum(t.source_u8.nbytes for t in flat.textures)
        e2e = {"value": allsum(e2e_rays) / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(width * height * 3), "steps": n_e2e, "s_per_frame": e2e_s / n_e2e,
               "note": "textures named by a stable key stay resident on the device after their first upload"}

    # ---- the other configurations, briefly (every rank takes part: same sharding and reduce) ------------------
    others = None
    if not args.no_configs and args.config == "cornell" and not (args.width or args.height or args.spp):
        others = {}
        for name, c in CONFIGS.items():
            if name == "cornell":
                continue
            o_spp = 4 if name == "stress" else c["spp"]            # the stress frame at 64 spp takes ~9 s on one GPU
            r = Runner(torch, dist, name, c["width"], c["height"], o_spp, "", world, rank)
            t = timed(torch, dist, world, r, 1, 1)
            ms, rays = allmax(t["dev_ms"]), allsum(t["rays"])
            entry = {"workload": c["what"] if o_spp == c["spp"] else c["what"].replace(f"{c['spp']} spp", f"{o_spp} spp (of {c['spp']})"),
                     "baseline_config": c["index"], "Mrays_per_s": rays / (ms * 1e-3) / 1e6, "s_per_frame": ms / 1e3,
                     "rays_per_frame": rays, "shadow_rays": allsum(t["shadow"]), "sharding": r.shard,
                     "colliders": int(len(r.flat.colliders)), "chunks_rank0": t["chunks"], "host_scene_build_s": r.host_build_s}
            if rank == 0:
                entry["frame"] = r.fingerprint()
                entry["frame"].pop("note"); entry["frame"].pop("block_means_4x4")
                others[name] = entry
            r.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rooflines of the dominant kernel (the fused level kernels) -------------------------------------
    peaks_file = REPO / "MEASURED_PEAKS.json"
    hbm_peak, hbm_src = 6650.0, "fallback"
    if peaks_file.exists():
        hbm_peak, hbm_src = float(json.loads(peaks_file.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    live = measure_peaks()
    level_s = max(tot["level_ms"] * 1e-3, 1e-9)
    n_launch = max(tot["level_launches"], 1)
    fpr = runner.flop_per_ray()
    bvh = len(runner.flat.colliders) >= 64
    flops = fpr * tot["rays"]
    per_ray_file = REPO / "profiles" / "r2b_per_ray.json"    # from the committed ncu launch list of this build (tools/ncu_per_ray.py)
    if not per_ray_file.exists():
        per_ray_file = REPO / "profiles" / "r2_per_ray.json"
    per_ray = json.loads(per_ray_file.read_text()).get(args.config) if per_ray_file.exists() else None
    traffic = per_ray["dram_bytes_per_ray"] * tot["rays"] / n_launch if per_ray else None
    kernel = ("sp_warp_kernel (levels >= 1: 98 % of the rays) + sp_level_kernel (level 0)" if args.config == "cornell"
              else "sp_trace_kernel + sp_level_kernel + sp_shadow_kernel (every level)" if bvh
              else "sp_hit_kernel + sp_shade_kernel (level 0) + sp_level_kernel (levels >= 1)" if width * height * spp >= (1 << 18)
              else "sp_level_kernel (all levels)")
    roofline = {
        "bound": "fp32", "kernel": kernel,
        "achieved": None if bvh else flops / level_s / 1e12, "peak": live["fp32_tflops"], "unit": "TFLOP/s",
        "frac": None if bvh else flops / level_s / 1e12 / live["fp32_tflops"],
        "peak_source": "FFMA chain micro-benchmark run by this process (sp_measure_peaks); MEASURED_PEAKS.json has no FP32 entry",
        "flop_per_ray": fpr, "rays_per_launch": tot["rays"] / n_launch,
        "avg_launch_ms": tot["level_ms"] / n_launch, "traffic": None,
        "note": ("scene of %d colliders behind a BVH: a ray tests a few of them, so algorithmic flops per ray are not defined; "
                 "see roofline_issue" % len(runner.flat.colliders)) if bvh else
                "fused generate + intersect + shade kernel: achieved counts only the algorithmic intersection flops of SURVEY 8(d); "
                "the binding resource is instruction issue, see roofline_issue",
    }
    clock_mhz = clock_summary.get("sm_mhz") or 1965.0
    issue_peak = 148 * 4 * 32 * clock_mhz * 1e6 / 1e12               # thread instructions / s: SMs x schedulers x lanes x clock
    roofline_issue = {
        "bound": "issue", "kernel": kernel, "unit": "T thread-instructions/s", "peak": issue_peak,
        "peak_source": "148 SMs x 4 schedulers x 32 lanes x SM clock under load (%.0f MHz)" % clock_mhz,
        "achieved": per_ray["thread_inst_per_ray"] * tot["rays"] / level_s / 1e12 if per_ray else None,
        "frac": per_ray["thread_inst_per_ray"] * tot["rays"] / level_s / 1e12 / issue_peak if per_ray else None,
        "thread_inst_per_ray": per_ray["thread_inst_per_ray"] if per_ray else None,
        "warp_inst_per_ray": per_ray["warp_inst_per_ray"] if per_ray else None,
        "source": f"smsp__thread_inst_executed.sum / rays of the committed ncu launch list (profiles/{per_ray_file.name})" if per_ray else
                  "no ncu launch list committed for this configuration",
    }
    roofline_hbm = {
        "bound": "hbm", "kernel": kernel + " (queue records only)",
        "achieved": tot["queue_bytes"] / level_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
        "frac": tot["queue_bytes"] / level_s / 1e9 / hbm_peak, "peak_source": hbm_src,
        "bytes_per_record": 96, "traffic": traffic,
        "algorithmic_bytes_per_launch": tot["queue_bytes"] / n_launch,
    }

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        w, h = cfg["cpu"]
        n = 2 if args.config == "cornell" else 1
        r, s, kind = cpu_rate(cfg["builder"], dict(cfg["kw"], width=w, height=h, **cfg.get("cpu_kw", {})), n, 1)
        what = "unmodified reference (baseline/_ref), get_raycolor" if kind == "reference" else "float64 numpy oracle port"
        cpu = {"value": r / s / 1e6, "unit": "Mrays/s", "cores": 1, "kind": kind,
               "sample": f"{cfg['builder']} scene{' ' + str(cfg['cpu_kw']) if cfg.get('cpu_kw') else ''} at {w}x{h}, {n} spp, {what}, single process"}

    out = {
        "metric": metric_name(args.config), "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": job_ms / args.steps, "s_per_frame": job_ms / args.steps / 1e3,
        "wall_ms_per_step": tot["wall_s"] / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{cfg['what'] if (width, height, spp) == (cfg['width'], cfg['height'], cfg['spp']) else cfg['builder'] + f' scene, {width}x{height}, {spp} spp'} "
                               f"(BASELINE.json configs[{cfg['index']}])",
                   "sharding": (f"{world} contiguous sample ranges" if runner.shard == "samples" else f"interleaved 64x64 tiles over {world} ranks")
                               + ", NCCL reduce of the float4 accumulation buffer",
                   "rays_per_frame": job_rays / args.steps, "rays_per_depth_rank0": tot["per_depth"],
                   "level_ms_rank0": tot["level_ms_by_depth"], "chunks_per_frame_rank0": tot["chunks"] / args.steps,
                   "chunk_retries": tot["retries"],
                   "l2": "wavefront queues are several GB per chunk, far larger than the 126 MB L2"},
        "frame": fingerprint,
        "e2e": e2e, "gpu_launches": job_launches, "clocks": clock_summary,
        "roofline": roofline, "roofline_issue": roofline_issue, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu,
        "shadow_rays": tot["shadow"], "configs": others,
    }
    emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    # The contract is ONE JSON line on stdout.  Libraries underneath print there too (NCCL announces its
    # version on stdout when the first communicator is created), so file descriptor 1 points at stderr for
    # the duration of the run and is restored for the result line only.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    lines = []
    emit = lines.append
    try:
        if args.impl == "reference":
            run_reference(args, emit)
        elif args.in_process and args.gpus > 1 and int(os.environ.get("WORLD_SIZE", "1")) <= 1:
            run_group(args, emit)
        else:
            run_native(args, emit)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    for line in lines:
        print(line, flush=True)


if __name__ == "__main__":
    main()
