#!/usr/bin/env python
"""bench.py -- Mpaths/sec of the path-tracing hot path (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--frames-per-step F] [--workload NAME]
  N>1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A STEP is one pass of the hot path over one batch: F dispatches (frames) of PathTracer::PathTrace at 1 spp/frame over the
whole image, i.e. W*H*F camera paths.  Workload (BASELINE.json configs[1]): Cornell box 1920x1080, full principled BSDF,
NEE+MIS (env + emissive mesh), depth 8; K*F = 1024 spp with the default K=8, F=128.  Data: the committed Cornell fixture
(12 triangles, identical to the shipped glTF) + shipped energy-compensation LUTs + a SYNTHETIC 4096x2048 HDR environment
map of the same size/dynamic range as the reference's default meadow_2_4k.hdr (25 MB, not committed).

value   : whole-job Mpaths/s, inputs resident in HBM, timed with CUDA events on the launching stream, max over ranks.
e2e     : same metric through the C-ABI with HOST buffers: per step the parameter block is re-uploaded, and the HDR
          accumulation image + the post-processed RGBA8 image are read back to pinned host memory.
N>1     : image-tile partition (16-row bands, rank r owns bands b with b % N == r), scene replicated, ONE NCCL gather of
          the framebuffer bands to rank 0 per step, no other collective (strong scaling: total work fixed).
--impl reference : the CPU oracle (port of the reference's estimator; the Vulkan reference cannot run here) on all host
          cores, each step = 1 frame of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (scene fixture, W, H, MaxDepth)
    "cornell_1080p_d8": ("cornell_box", 1920, 1080, 8),
    "glass_1080sq_d16": ("cornell_box_glass", 1080, 1080, 16),
    "breakfast_1080p_d8": ("breakfast_room", 1920, 1080, 8),
    "viking_1080sq_d8": ("viking_room", 1080, 1080, 8),
}
BASE_SEED = 0x1234ABCD
BAND_ROWS = 16
PEAKS_FALLBACK_GBS = 6650.0


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    except Exception:
        return PEAKS_FALLBACK_GBS, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  The sampler runs from the first warm-up
    step to the end of the timed steps; mark_begin() / mark_end() bracket the timed region.  Samples inside the bracket are reported; if
    the region is shorter than the sampling period (multi-GPU runs finish a step in ~10 ms) the samples of the whole loaded window
    (warm-up + timed steps, same kernels, same clocks) are used instead and "window" says so."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows = []; self.proc = None; self.gpu = gpu_index; self.i0 = None; self.i1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True); self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        try:
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def mark_begin(self): self.i0 = len(self.rows)

    def mark_end(self): self.i1 = len(self.rows)

    def stop(self):
        if not self.proc: return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        try:
            self.proc.terminate()
            self.proc.wait(timeout=2)
        except Exception:
            pass
        rows = list(self.rows)
        i0 = self.i0 if self.i0 is not None else 0
        i1 = self.i1 if self.i1 is not None else len(rows)
        timed = [r for r in rows[i0:i1] if len(r) >= 9]
        window = "timed region"
        if not timed:
            timed = [r for r in rows if len(r) >= 9]; window = "warm-up + timed steps (timed region shorter than the 100 ms sampling period)"
        def num(x):
            try: return float(x)
            except Exception: return None
        sm = [v for v in (num(r[1]) for r in timed) if v is not None]
        mx = [v for v in (num(r[2]) for r in timed) if v is not None]
        reasons = set()
        for r in timed:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"): reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm), "window": window}


ASSETS = os.path.join(ROOT, "oracle", "_ref", "assets")    # copy of the reference's shipped assets (oracle/Makefile `assets`; git-ignored, travels to the GPU box)
ASSET_FILES = {"cornell_box": "CornellBox.gltf", "cornell_box_glass": "CornellBoxGlass.gltf", "breakfast_room": "BreakfastRoom.gltf", "viking_room": "VikingRoom.gltf"}
ENV_FILE = "meadow_2_4k.hdr"                               # PathTracer.h:208, the reference's default environment map


def have_assets(scene):
    return all(os.path.isfile(os.path.join(ASSETS, f)) for f in (ASSET_FILES[scene], ENV_FILE, "LookupTables/ReflectionLookup.bin"))


def synthetic_env_4k():
    from oracle import gltf_ref
    return gltf_ref.synthetic_env(4096, 2048, seed=3, sun=150000.0)


def data_description(scene):
    return "reference assets (shipped glTF scene, meadow_2_4k.hdr, lookup tables); random-number streams seeded by bench.py" if have_assets(scene) else "synthetic"


def env_description(real):
    return ("meadow_2_4k.hdr 4096x2048 (the reference's default map, RGBA32F 128 MiB + 64 MiB alias table)" if real
            else "synthetic 4096x2048 RGBA32F (128 MiB) + 64 MiB alias table")


def bench_config(workload, frames_per_step, steps, world, extra=None):
    """`config` of the JSON line: the same keys in both arms (the driver compares them)."""
    scene, W, H, depth = WORKLOADS[workload]
    real = have_assets(scene)
    c = {"workload": workload, "scene": scene, "image": [W, H], "max_depth": depth, "spp_per_frame": 1, "frames_per_step": frames_per_step,
         "spp_total": frames_per_step * steps, "partition": f"{BAND_ROWS}-row bands x {world} ranks",
         "scene_source": (ASSET_FILES[scene] + " through the product's own glTF loader (set_scene_file)") if real else "committed fixture tests/golden/" + scene + ".npz (set_scene_arrays)",
         "env_map": env_description(real),
         "l2_policy": "inputs larger than L2 (env map + alias table 192 MiB, wavefront state of a 16.6 M-path wave > 1 GB vs 126 MB L2): no explicit flush"}
    if extra: c.update(extra)
    return c


def algorithmic_bytes(c):
    """SURVEY.md 8(d) streaming model, per kernel, from the device counters of one step."""
    return {
        "raygen": 68 * c["paths"],
        "extend": 44 * c["extend_rays"],
        "shade": 156 * c["shade_invocations"] + 52 * c["shadow_rays"],
        "connect": 76 * c["shadow_rays"] + 8 * c["shade_invocations"],
        "resolve": 48 * c["paths"],
        "total": 116 * c["paths"] + 224 * c["extend_rays"] + 128 * c["shadow_rays"],
    }


def diff_counters(a, b):
    return {k: (b[k] - a[k]) for k in ("paths", "extend_rays", "shade_invocations", "surface_hits", "misses", "shadow_rays", "medium_events", "kernel_launches")}


def cpu_arm_threads(info):
    import math
    return max(1, min(info["logical_cpus"], int(math.ceil(info["effective_cpus"]))))


def cpu_arm_describe(info, nthreads, sample):
    return {"cores": nthreads, "kind": "port", "sample": sample, "build": "oracle/liboracle_fast.so: gcc -O3 -march=native, compiled on this host",
            "host": {"logical_cpus": info["logical_cpus"], "model": info["model"], "cgroup_cpu_max": info["cgroup_cpu_max"],
                     "effective_cpus": info["effective_cpus"], "affinity_cpus": info.get("affinity_cpus")}}


def oracle_scene_for(scene, depth):
    """CPU arm: the oracle's scene + config for a workload -- real assets through the oracle-side loaders when the copy is there, else fixtures + synthetic map."""
    import util
    from oracle import orc, gltf_ref
    if have_assets(scene):
        sd = gltf_ref.load_gltf(os.path.join(ASSETS, ASSET_FILES[scene]))
        raw = orc.load_hdr(os.path.join(ASSETS, ENV_FILE))
        luts = gltf_ref.load_luts_dir(os.path.join(ASSETS, "LookupTables"))
    else:
        sd = util.scene_dict(scene); raw = synthetic_env_4k(); luts = util.luts()
    env_pdf, alias, _ = orc.build_env_alias(raw)
    vi, pi = orc.camera_from_view(sd["camera_view"], sd["aspect"])
    return orc.Scene(sd, env_pdf, alias, luts), orc.default_config(ViewInverse=vi, ProjectionInverse=pi, MaxDepth=depth)


def run_reference(args):
    """CPU arm: the oracle (port of the reference's Slang estimator; the Vulkan reference cannot run here -- profiles/r02_vulkan_host_probe.txt),
    -O3 -march=native build, one thread per CPU the cgroup quota grants; step = 1 frame; median of >= 3 timed repeats of the K steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import util
    from oracle import orc
    orc.use_fast_build()
    info = orc.host_cpu_info()
    cores = cpu_arm_threads(info)
    scene, W, H, depth = WORKLOADS[args.workload]
    S, cfg = oracle_scene_for(scene, depth)
    img = np.zeros((H, W, 4), np.float32)
    f = 0
    for _ in range(args.warmup):
        S.render(cfg, W, H, 1, BASE_SEED, frame0=f, image=img, nthreads=cores); f += 1
    reps = []
    for _ in range(3):
        t0 = time.perf_counter()
        for _ in range(args.steps):
            S.render(cfg, W, H, 1, BASE_SEED, frame0=f, image=img, nthreads=cores); f += 1
        reps.append(time.perf_counter() - t0)
    dt = sorted(reps)[1]
    v = W * H * args.steps / dt / 1e6
    cb = cpu_arm_describe(info, cores, f"bounded sample: each step = 1 frame (1 spp of the full {W}x{H} image) of the workload's {args.frames_per_step}; "
                                       f"{args.steps} steps, median of 3 repeats (pthreads over pixel rows)")
    cb.update({"value": v, "unit": "Mpaths/s", "repeats_mpaths": [W * H * args.steps / r / 1e6 for r in reps]})
    out = {"impl": "reference", "metric": "Mpaths/sec", "value": v, "unit": "Mpaths/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": data_description(scene),
           "config": bench_config(args.workload, args.frames_per_step, args.steps, max(args.gpus, 1)),   # the b200 arm's config; the bounded sample is in cpu_baseline.sample
           "cpu_baseline": cb,
           "e2e": {"value": v, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out))


def post_row_block(H, rank, world):
    """contiguous block of output rows of rank `rank` (boundaries on multiples of 8 rows)"""
    cut = lambda r: (H * r // world + 7) // 8 * 8 if r < world else H
    return min(cut(rank), H), min(cut(rank + 1), H)


def run_post(args):
    """BASELINE configs[4]: 3840x2160 HDR accumulate + bloom (10 mips) + tonemap, post only, on 1 / 2 / 4 / 8 GPUs.  A step = one frame: every rank folds a new
    radiance frame into the accumulation image on the rows its block depends on (b200pt_accumulate_rows), runs the bloom chain + tonemap for its block of
    output rows (b200pt_post_process_rows: halo rows recomputed, no exchange) and, for N > 1, the RGBA8 blocks are gathered on rank 0 (one NCCL gather).
    value = GB/s by the reference's pass-per-pass byte model (174.7 B/pixel) over the whole image / time, max over ranks."""
    import torch
    import torch.distributed as dist
    import util
    import vpt_b200 as pt
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, H = 3840, 2160
    rng = np.random.default_rng(7)
    hdr = np.ones((H, W, 4), np.float32); hdr[..., :3] = (np.exp(rng.normal(0, 1.5, (H, W, 3))) * 0.5).astype(np.float32)
    for _ in range(64):
        y, x = rng.integers(0, H - 5), rng.integers(0, W - 5); hdr[y:y + 5, x:x + 5, :3] = 500.0
    T = pt.PathTracer(local)
    T.set_scene(util.scene_dict("cornell_box")); T.resize(W, H)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); T.set_stream(stream.cuda_stream)
    T.set_hdr(hdr)
    frame = torch.from_numpy(hdr).cuda()                         # the "new frame" folded in every step (resident in HBM)
    y0, y1 = post_row_block(H, rank, world)
    in0, in1 = T.post_input_rows(y0, y1)
    blk_rows = max(post_row_block(H, r, world)[1] - post_row_block(H, r, world)[0] for r in range(world))
    blk = torch.zeros((blk_rows, W, 4), dtype=torch.uint8, device="cuda")
    gathered = [torch.zeros_like(blk) for _ in range(world)] if (world > 1 and rank == 0) else None
    fidx = [1]

    def step():
        T.accumulate_rows(frame.data_ptr(), fidx[0], in0, in1); fidx[0] += 1
        T.post_process_rows(y0, y1)
        if world > 1:
            T.get_ldr_rows_into_device(y0, y1, blk.data_ptr())
            dist.gather(blk, gathered, dst=0)

    def barrier():
        if world > 1: dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local); sampler.start()
    iters = max(args.steps, 8) * 25
    for _ in range(max(args.warmup, 3) * 5): step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.mark_begin()
    ev0.record(stream)
    for _ in range(iters): step()
    ev1.record(stream); barrier()
    sampler.mark_end()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1) / iters
    tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1: dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    # e2e: the frame comes from pinned HOST memory every step and the RGBA8 result goes back to the host
    host_frame = torch.from_numpy(hdr).pin_memory(); host_ldr = torch.empty((y1 - y0, W, 4), dtype=torch.uint8).pin_memory()
    def e2e_step():
        frame[in0:in1].copy_(host_frame[in0:in1], non_blocking=True)
        step()
        T.get_ldr_rows(y0, y1, host_ldr.numpy())
    e2e_step(); barrier()
    n_e2e = 20
    t0 = time.perf_counter()
    for _ in range(n_e2e): e2e_step()
    barrier()
    te = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device="cuda")
    if world > 1: dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    mips = pt.bloom_mip_sizes(W, H)
    px = [w * h for w, h in mips]
    # reference pass structure: accumulate RMW 32 B/px; threshold R16+W16; down i: read mip i-1 + write mip i; up i: read mip i + RMW mip i-1; tonemap: hdr + bloom -> rgba8
    bytes_model = 32 * px[0] + 32 * px[0] + sum(16 * px[i - 1] + 16 * px[i] for i in range(1, len(px))) + sum(16 * px[i] + 32 * px[i - 1] for i in range(1, len(px))) + (16 + 16 + 4) * px[0]
    # what the fused chain really has to move for the WHOLE image: accumulate (frame read + image RMW = 48 B/px), first down pass reads the image, the final kernel reads image + mip 1 and writes RGBA8
    bytes_fused = 48 * px[0] + (16 * px[0] + 16 * px[1]) + sum(16 * px[i - 1] + 16 * px[i] for i in range(2, len(px))) + sum(16 * px[i] + 32 * px[i - 1] for i in range(2, len(px))) + (16 * px[0] + 16 * px[1] + 4 * px[0])
    peak, kind = measured_peak()
    gbs = bytes_model / (ms * 1e-3) / 1e9
    moved = bytes_fused / (ms * 1e-3) / 1e9
    if rank == 0:
        print(json.dumps({"metric": "post chain GB/s (3840x2160 HDR accumulate + bloom 10 mips + tonemap)", "value": gbs, "unit": "GB/s", "n_gpus": world, "steps": iters, "warmup": max(args.warmup, 3) * 5,
                          "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": "post_4k", "image": [W, H], "mips": mips, "bytes_per_pixel_model": bytes_model / px[0], "chain": "fused, row blocks + recomputed halo",
                                     "bytes_per_pixel_moved": bytes_fused / px[0], "row_block_rank0": [y0, y1], "input_rows_rank0": [in0, in1],
                                     "l2_policy": "inputs larger than L2 (HDR image 133 MB + frame 133 MB vs 126 MB L2), no explicit flush",
                                     "note": "value = the reference's pass-per-pass byte model (SURVEY 8d, 174.7 B/px) / time: the fused chain moves fewer bytes, so value may exceed N x the HBM peak; roofline.achieved counts the bytes the chain really moves, per GPU"},
                          "e2e": {"value": bytes_model / e2e_s / 1e9, "unit": "GB/s", "h2d_bytes_per_step": int((in1 - in0) * W * 16), "d2h_bytes_per_step": int((y1 - y0) * W * 4), "steps": n_e2e},
                          "gpu_launches": int(iters * (1 + 2 * (len(px) - 1))), "clocks": clocks,
                          "roofline": {"bound": "hbm", "kernel": "post chain (k_accumulate, k_bloom_down_first .. k_bloom_final)", "achieved": moved / world, "peak": peak, "peak_kind": kind, "unit": "GB/s",
                                       "frac": moved / world / peak, "traffic": None, "pass_per_pass_model_gbs": gbs}}))
    if world > 1: dist.destroy_process_group()


def run_lut_bake(args):
    """the energy-compensation LUT bake (LookupTableCalculator::CalculateTable, Application.cpp:35-72).  Pure ALU/MUFU
    work (4 B written per texel): reported as BSDF samples/s; the oracle restatement is timed beside it on a few texels."""
    import util
    import vpt_b200 as pt
    from oracle import orc
    T = pt.PathTracer(0)
    n_samp = 20000
    out = {}
    tot_s, tot_ms = 0, 0.0
    for kind, name in ((0, "reflect_64x64x32"), (1, "refract_outside_128x128x32"), (2, "refract_inside_128x128x32")):
        T.bake_lut(kind, 400, seed=1)                                   # warm-up
        tab, ms = T.bake_lut(kind, n_samp, seed=7)
        out[name] = {"ms": ms, "Gsamples_per_s": tab.size * n_samp / (ms * 1e-3) / 1e9}
        tot_s += tab.size * n_samp; tot_ms += ms
    L = orc.lib(); t0 = time.time(); cpu_s = 0
    for i in range(16):
        L.orc_bake_lut_texel(1, 128, 128, 32, 8 * i + 3, 5 * i + 20, 2 * i, 200000, 1); cpu_s += 200000
    cpu = cpu_s / (time.time() - t0) / 1e9
    print(json.dumps({"metric": "LUT bake Gsamples/s", "value": tot_s / (tot_ms * 1e-3) / 1e9, "unit": "Gsamples/s", "n_gpus": 1, "steps": 1, "warmup": 1, "ms_per_step": tot_ms,
                      "higher_is_better": True, "dtype": "f32", "data": "synthetic", "config": {"workload": "lut_bake", "samples_per_texel": n_samp, "tables": out,
                      "full_bake_estimate_s": 1e7 / n_samp * tot_ms * 1e-3},
                      "cpu_baseline": {"value": cpu, "unit": "Gsamples/s", "cores": 1, "kind": "port", "sample": "16 texels x 200,000 samples of the refraction table, one core"}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=128)
    ap.add_argument("--frames-in-flight", type=int, default=0)
    ap.add_argument("--workload", default="cornell_1080p_d8", choices=sorted(WORKLOADS) + ["post_4k", "lut_bake"])
    ap.add_argument("--cpu-baseline-frames", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3: args.warmup = 3
    if args.workload == "lut_bake":
        return run_lut_bake(args)
    if args.workload == "post_4k":
        return run_post(args)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import util
    import vpt_b200 as pt

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1: args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    scene, W, H, depth = WORKLOADS[args.workload]
    F = args.frames_per_step
    T = pt.PathTracer(local)
    setup_s = {}                                 # scene load + acceleration-structure build, env-map preprocessing, tables: outside the metric, reported beside it
    t_setup = time.perf_counter()
    if have_assets(scene):                       # the product's own SetScene path: glTF + textures + HDR + lookup tables from the reference's shipped files
        T.set_scene_file(os.path.join(ASSETS, ASSET_FILES[scene])); T.synchronize()
        setup_s["set_scene_s"] = time.perf_counter() - t_setup; t_setup = time.perf_counter()
        T.set_env_map_file(os.path.join(ASSETS, ENV_FILE)); T.synchronize()
        setup_s["set_env_map_s"] = time.perf_counter() - t_setup; t_setup = time.perf_counter()
        T.set_luts_dir(os.path.join(ASSETS, "LookupTables")); T.synchronize()
        setup_s["set_luts_s"] = time.perf_counter() - t_setup
    else:
        T.set_scene(util.scene_dict(scene)); T.synchronize()
        setup_s["set_scene_s"] = time.perf_counter() - t_setup; t_setup = time.perf_counter()
        T.set_env_map(synthetic_env_4k()); T.synchronize()
        setup_s["set_env_map_s"] = time.perf_counter() - t_setup; t_setup = time.perf_counter()
        T.set_luts(*util.luts()); T.synchronize()
        setup_s["set_luts_s"] = time.perf_counter() - t_setup
    cfg = pt.default_config(MaxDepth=depth, MaxSamplesAccumulated=0x7FFFFFFF, FramesInFlight=args.frames_in_flight)
    T.set_config(cfg)
    T.resize(W, H)
    T.set_partition(rank, world, BAND_ROWS)
    stream = torch.cuda.Stream()                 # a real (non-NULL) stream shared by the renderer, the events and NCCL
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    T.set_stream(stream.cuda_stream)
    rows = T.local_rows()
    max_rows = max(pt.lib().b200pt_partition_local_row_count(H, r, world, BAND_ROWS) for r in range(world))
    band = torch.zeros((max_rows, W, 4), dtype=torch.float32, device="cuda")
    gathered = [torch.zeros_like(band) for _ in range(world)] if (world > 1 and rank == 0) else None

    def gather():
        if world == 1: return
        T.get_hdr_into_device(band.data_ptr())
        dist.gather(band, gathered, dst=0)

    def step():
        T.path_trace(F, BASE_SEED)               # asynchronous: waves of consecutive steps overlap on the handle's internal streams
        gather()

    def barrier():
        if world > 1: dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local); sampler.start()
    for _ in range(args.warmup): step()
    barrier()
    c0 = T.counters()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    t0 = time.perf_counter()
    ev0.record(stream)
    for _ in range(args.steps): step()
    T.flush()                                    # the end event must sit behind every launched wave (device-side wait, no host sync)
    ev1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    sampler.mark_end()
    ms = ev0.elapsed_time(ev1)
    if abs(ms * 1e-3 - wall) > 0.05 * wall + 2e-3:   # device-event time must agree with the wall-clock bracket
        ms = max(ms, wall * 1e3)
    clocks = sampler.stop()
    c1 = T.counters()
    dc = diff_counters(c0, c1)
    tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1: dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    paths_total = W * H * F * args.steps
    value = paths_total / (ms_max * 1e-3) / 1e6

    # ---- e2e through the C-ABI with host buffers (rank-local images; rank 0 assembles for N>1)
    hdr_host = torch.empty((rows, W, 4), dtype=torch.float32, pin_memory=True).numpy()
    ldr_host = torch.empty((H, W, 4), dtype=torch.uint8, pin_memory=True).numpy() if world == 1 else None
    def e2e_step():
        T.set_config(cfg)                       # parameter block re-upload (the reference's setters, PathTracer.cpp:940-986)
        T.set_partition(rank, world, BAND_ROWS)
        T.path_trace(F, BASE_SEED)
        gather()
        T.get_hdr(hdr_host)                     # D2H of the accumulation image
        if world == 1:
            T.post_process(); T.get_ldr(ldr_host)   # bloom + tonemap + RGBA8 read-back (image-output path)
    e2e_steps = max(2, min(args.steps, 4))
    if os.environ.get("B200PT_BENCH_E2E_BREAKDOWN"):   # where does an end-to-end step spend its host time?  (stderr, not part of the JSON line)
        def tm(label, fn):
            t = time.perf_counter(); fn(); torch.cuda.synchronize(); sys.stderr.write(f"[e2e] {label}: {(time.perf_counter() - t) * 1e3:.2f} ms\n")
        for _ in range(2):
            tm("set_config", lambda: T.set_config(cfg)); tm("set_partition", lambda: T.set_partition(rank, world, BAND_ROWS))
            tm("path_trace", lambda: T.path_trace(F, BASE_SEED)); tm("gather", gather); tm("get_hdr", lambda: T.get_hdr(hdr_host))
            if world == 1: tm("post_process", T.post_process); tm("get_ldr", lambda: T.get_ldr(ldr_host))
    e2e_step(); barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps): e2e_step()
    barrier()
    e2e_dt = time.perf_counter() - t0
    te = torch.tensor([e2e_dt], dtype=torch.float64, device="cuda")
    if world > 1: dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = W * H * F * e2e_steps / float(te.item()) / 1e6
    h2d = F * 16 + 512                           # dispatch table (16 B/dispatch) + kernel parameter blocks
    d2h = rows * W * 16 + (W * H * 4 if world == 1 else 0)

    # ---- per-kernel split of one profiled step -> roofline of the dominant kernel
    T.set_config(cfg); T.set_partition(rank, world, BAND_ROWS)
    T.set_profiling(True)
    ca = T.counters(); T.path_trace(F, BASE_SEED); cb = T.counters()
    T.set_profiling(False)
    dprof = diff_counters(ca, cb)
    kms = {k: cb["ms_" + k] for k in ("raygen", "extend", "shade", "connect", "resolve")}
    ab = algorithmic_bytes(dprof)
    fused = kms["connect"] == 0.0 and dprof["shadow_rays"] > 0      # fused bounce kernel (k_shade_hit<CLASS, ., 2>): shade + NEE queries + epilogue + next TraceRay
    if fused:                                                        # its algorithmic bytes = the three per-bounce stages it replaces (SURVEY 8d model unchanged)
        ab["shade"] = ab["shade"] + ab["connect"] + 44 * max(dprof["extend_rays"] - dprof["paths"], 0)
        ab["extend"] = 44 * dprof["paths"]; ab["connect"] = 0
    dom = max(kms, key=lambda k: kms[k])
    n_launch = {"raygen": cb["waves"], "resolve": cb["waves"]}
    launches_dom = n_launch.get(dom, cb["bounces"])
    peak, peak_kind = measured_peak()
    achieved = ab[dom] / (kms[dom] * 1e-3) / 1e9 if kms[dom] > 0 else 0.0
    seg_per_path = dprof["extend_rays"] / max(dprof["paths"], 1)
    pipeline_bytes = ab["total"]
    step_ms_prof = sum(kms.values())
    kname = {"shade": "k_shade_hit<CLASS,.,2> fused bounce kernel: shade + NEE queries + roulette + next TraceRay (+k_shade_miss)" if fused else "k_shade_hit<CLASS> (+k_shade_miss)", "extend": "k_extend", "connect": "k_connect", "raygen": "k_raygen", "resolve": "k_resolve"}[dom]
    traffic = None                                # DRAM bytes per launch of that kernel from the committed ncu --set full capture, if any
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        ent = tj.get(args.workload, {}).get(dom)
        if ent: traffic = {"dram_bytes_per_launch": ent["dram_bytes_per_launch"], "algorithmic_bytes_of_that_launch": ent.get("algorithmic_bytes"), "source": ent["source"]}
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "launches_per_step": int(launches_dom), "avg_launch_ms": kms[dom] / max(launches_dom, 1),
                "algorithmic_bytes_per_launch": ab[dom] / max(launches_dom, 1),
                "kernel_ms_per_step": kms, "kernel_share": {k: (v / step_ms_prof if step_ms_prof else 0.0) for k, v in kms.items()},
                "pipeline": {"algorithmic_bytes_per_step": pipeline_bytes, "segments_per_path": seg_per_path,
                             "shadow_rays_per_path": dprof["shadow_rays"] / max(dprof["paths"], 1),
                             "roofline_mpaths": peak * 1e9 / (pipeline_bytes / max(dprof["paths"], 1)) / 1e6,
                             "frac": value / world / (peak * 1e9 / (pipeline_bytes / max(dprof["paths"], 1)) / 1e6)}}

    out = {"metric": "Mpaths/sec", "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": data_description(scene),
           "config": bench_config(args.workload, F, args.steps, world),
           "notes": {"l2_policy": "working set > L2: env map 128 MiB + alias 64 MiB + wavefront state of a 16.6 M-path wave (126 MB L2), no explicit flush",
                     "timing": "CUDA events on the launching stream, max over ranks", "wall_s": wall,
                     "setup_s": {k: round(v, 4) for k, v in setup_s.items()}},      # host wall clock of rank 0: file decode + upload + BVH build / alias table / tables; not in the metric
           "e2e": {"value": e2e_value, "unit": "Mpaths/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps},
           "gpu_launches": int(dc["kernel_launches"]), "clocks": clocks, "roofline": roofline,
           "counters_per_step": {k: dc[k] / args.steps for k in dc}}

    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            from oracle import orc
            orc.use_fast_build()
            info = orc.host_cpu_info(); cores = cpu_arm_threads(info)
            S, ocfg = oracle_scene_for(scene, depth)
            n = args.cpu_baseline_frames
            S.render(ocfg, W, H, 1, BASE_SEED, nthreads=cores)
            reps = []
            for r in range(3):
                t0 = time.perf_counter(); S.render(ocfg, W, H, n, BASE_SEED, frame0=1 + r * n, nthreads=cores); reps.append(time.perf_counter() - t0)
            dt = sorted(reps)[1]
            cb = cpu_arm_describe(info, cores, f"{n} frames x 1 spp of the full {W}x{H} image, median of 3 repeats (CPU oracle)")
            cb.update({"value": W * H * n / dt / 1e6, "unit": "Mpaths/s", "repeats_mpaths": [W * H * n / r / 1e6 for r in reps]})
            out["cpu_baseline"] = cb
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
