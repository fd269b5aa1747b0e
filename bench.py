#!/usr/bin/env python
"""Benchmark of the ArtiBoost synthesis hot path: synthesised views/s (BASELINE.json metric), rasteriser workload.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA kernels through the C-ABI)
  python bench.py --impl reference --gpus N ...            reference arm: the CPU restatement on the host cores

Workload (BASELINE.json configs[1]): batch-512 hand+object rasteriser, RGBA8 + depth f32 + seg u8 at 256x256, HO3D CCV
space (4 objects x 288 views x 50 grasps), synthetic MANO-shaped hand (778 verts / 1538 faces) and synthetic YCB-shaped
objects (8192 verts / 16380 faces).  A step = one pass of ab_render_batch over one batch of 512 views whose per-view
inputs (posed hand vertices, object poses, ids, light, background crop) are already resident in HBM.  One process per
GPU, views sharded across ranks with no data-path collective (weak scaling).  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# torchrun exports OMP_NUM_THREADS=1 to every rank; the host-side legs of this file (CPU baseline, asset generation, the
# torch reference models' initialisation) then crawl on one core.  Give every rank its share of the cores instead --
# before numpy / torch load their OpenMP runtimes.
if os.environ.get("OMP_NUM_THREADS") == "1" and "TORCHELASTIC_RUN_ID" in os.environ:
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))

if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":   # the version banner goes to stdout, which carries ONE JSON line
    os.environ["NCCL_DEBUG"] = "WARN"

BATCH = 512
SIZE = 256
BYTES_PER_VIEW_OUT = SIZE * SIZE * 9                # RGBA8 + depth f32 + seg u8
BYTES_PER_VIEW_IN = 778 * 12 + 64 + 4 + 4 + 4 + 20  # hand verts, pose, obj id, texture id, light, bg_sel
ALGO_BYTES_PER_VIEW = BYTES_PER_VIEW_OUT + BYTES_PER_VIEW_IN


def bf16_sustained_peak():
    """Measured sustained dense bf16 TFLOP/s of this pool's B200s (MEASURED_PEAKS.json), else the profiling recipe's figure."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["bf16_tflops_sustained"])
    except Exception:
        return 1364.6


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self._paused = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {getattr(nv, n): n[len("nvmlClocksThrottleReason"):] for n in dir(nv)
                     if n.startswith("nvmlClocksThrottleReason") and isinstance(getattr(nv, n), int)}
            while not self._halt.is_set():
                if self._paused.is_set():   # no NVML traffic to this GPU inside the timed window
                    time.sleep(0.001)
                    continue
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if bit and (mask & bit) == bit and name not in ("None", "All"):
                        self.reasons.add(name)
                time.sleep(self.period)
        except Exception as e:  # NVML missing: report it instead of failing the bench
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def pause(self, on=True):
        (self._paused.set if on else self._paused.clear)()

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        rename = {"GpuIdle": "gpu_idle", "SwPowerCap": "sw_power_cap", "HwSlowdown": "hw_slowdown",
                  "HwThermalSlowdown": "hw_thermal_slowdown", "SwThermalSlowdown": "sw_thermal_slowdown",
                  "HwPowerBrakeSlowdown": "hw_power_brake_slowdown", "ApplicationsClocksSetting": "applications_clocks_setting",
                  "SyncBoost": "sync_boost", "DisplayClockSetting": "display_clock_setting"}
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "samples": len(s),
                "reasons": sorted(rename.get(r, r) for r in self.reasons if r != "GpuIdle")}


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_reference(steps, warmup, sample_views, threads=None):
    """The CPU restatement of the same workload (oracle/: numpy pose generator -> C rasteriser, OpenMP over views)."""
    from artiboost_b200 import assets
    from oracle.synth_cpu import CpuSynth
    threads = threads or os.cpu_count() or 1
    synth = CpuSynth(assets, seed=0)
    n_res = max(1, int(os.environ.get("AB_BENCH_BATCHES", "10")))   # the same cycle of resident batches as the GPU arm
    inps = [synth.sample(sample_views) for _ in range(n_res)]
    out = None
    for i in range(warmup):
        out = synth.render(inps[i % n_res], n_threads=threads, out=out)
    t0 = time.perf_counter()
    for i in range(steps):
        out = synth.render(inps[i % n_res], n_threads=threads, out=out)
    dt = time.perf_counter() - t0
    return sample_views * steps / dt, dt / steps * 1e3, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = BATCH   # the same 512 views per step as our arm
    vps, ms, threads = cpu_reference(args.steps, args.warmup, sample)
    line = {
        "impl": "reference", "metric": METRIC, "value": vps,
        "unit": "views/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, views_per_launch=BATCH, submission="host loop over oracle/raster.c (OpenMP over views)",
                                  resident_batches=f"{max(1, int(os.environ.get('AB_BENCH_BATCHES', '10')))} different batches of {BATCH} views, visited in turn by the steps"),
        "cpu_baseline": {"value": vps, "unit": "views/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} views/step x {args.steps} steps of the batch-512 workload, oracle/raster.c "
                                   f"with OpenMP over views on {threads} host threads (pyrender/EGL cannot run here)"},
        "e2e": {"value": vps, "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


METRIC = "synthesised views/sec (rasteriser, RGBA+depth+seg 256x256)"


def workload_config(n_gpus, **extra):
    cfg = {"workload": "BASELINE.json configs[1]: batch-512 hand+object rasteriser only (RGBA8+depth f32+seg u8, 256x256)",
           "views_per_step_per_gpu": BATCH, "image": [SIZE, SIZE], "ccv_space": [4, 288, 50],
           "hand_mesh": [778, 1538], "object_mesh": [8192, 16380], "parallelism": f"views sharded over {n_gpus} rank(s), no collective",
           "l2": "each step writes 302 MB of views (> 126 MB L2); meshes and patch tables (3 MB) are shared by the batch and L2-resident by design"}
    cfg.update(extra)
    return cfg


def oracle_check(pipe, poses, rand, out, idx):
    """A few views of the timed batch against oracle/raster.c, bit for bit (outside every timed region)."""
    import numpy as np
    from oracle import raster
    r = pipe.renderer
    K = pipe.cam_intr
    cfg = dict(width=r.width, height=r.height, fx=float(K[0, 0]), fy=float(K[1, 1]), cx=float(K[0, 2]), cy=float(K[1, 2]),
               znear=0.05, cull_backface=1, ambient=0.8, diffuse=0.25)
    hf = r.hand_faces.cpu().numpy()[:, :3]
    hcols = r.hand_colors.cpu().numpy()
    bgs = r.backgrounds.cpu().numpy()
    ocols = r.obj_colors.cpu().numpy()
    ok = 0
    for i in idx:
        oid = int(poses["obj_id"][i])
        o = pipe.objects[pipe.obj_names[oid]]
        sel = rand["bg_sel"][i].cpu().numpy()
        rgba, depth, seg, _ = raster.render_view(
            cfg, poses["final_hand_verts"][i].cpu().numpy(), hf, hcols[int(rand["hand_tex"][i])], o["vertices"], o["faces"],
            ocols[r._voff[oid]:r._voff[oid + 1]], poses["final_obj_pose"][i].cpu().numpy(), light=float(rand["light"][i]),
            bg=bgs[sel[0]], bg_sel=sel[1:])
        same = (np.array_equal(out["seg"][i].cpu().numpy(), seg) and np.array_equal(out["rgba"][i].cpu().numpy(), rgba)
                and np.array_equal(out["depth"][i].cpu().numpy().view(np.uint32), depth.view(np.uint32)))
        if not same:
            raise SystemExit(f"bench: view {i} of the timed batch differs from oracle/raster.c")
        ok += 1
    return ok


# ------------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from artiboost_b200 import build, lib
    from artiboost_b200.synth import SynthPipeline

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (our arm) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # NCCL prints its version banner on stdout, next to the JSON line
        dist.init_process_group("nccl", device_id=dev)
    if not os.path.exists(lib.LIB_PATH):
        build.build()
    lib.load()
    wall = {"start": time.perf_counter()}

    pipe = SynthPipeline(device=dev, seed=1, sample_seed=1 + rank + int(os.environ.get("AB_BENCH_SEED_OFFSET", "0")),
                         chunk=BATCH)   # same assets on every rank, different draws
    # N_RES different batches of 512 views stay resident and the steps cycle through them: the triangle work of ONE random batch
    # varies by +-10 % with its views (per-rank times of a single batch: 0.255 ... 0.318 ms at N = 4), and the job's time is the
    # slowest rank's, so a single batch per rank would make the 1 -> 8 curve a statement about seeds
    N_RES = max(1, int(os.environ.get("AB_BENCH_BATCHES", "10")))
    resident = []
    for _ in range(N_RES):
        p_ = pipe.sample_poses(BATCH)           # CCV draw -> view -> grasp -> pose generator (device)
        resident.append((p_, pipe.draw_render_randoms(BATCH)))
    poses, rand = resident[0]
    out = {"rgba": torch.empty((BATCH, SIZE, SIZE, 4), dtype=torch.uint8, device=dev),
           "depth": torch.empty((BATCH, SIZE, SIZE), dtype=torch.float32, device=dev),
           "seg": torch.empty((BATCH, SIZE, SIZE), dtype=torch.uint8, device=dev)}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    counter = [0, 0]

    def eager_step():
        p_, r_ = resident[counter[0] % N_RES]
        counter[0] += 1
        pipe.render(p_, r_, out=out)

    # one step = ab_render_batch over a resident batch, replayed from a CUDA graph so that the 1 -> 8 GPU curve does not
    # depend on the host (the call is two kernel launches; it is capture-safe by construction)
    l0 = lib.launch_count()
    eager_step()
    torch.cuda.synchronize(dev)
    launches_per_step = lib.launch_count() - l0
    side = torch.cuda.Stream(dev)
    graphs = []
    with torch.cuda.stream(side):
        for p_, r_ in resident:
            pipe.render(p_, r_, out=out)
            torch.cuda.synchronize(dev)
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_, stream=side):
                pipe.render(p_, r_, out=out)
            graphs.append(g_)
    torch.cuda.synchronize(dev)

    def step():
        graphs[counter[1] % N_RES].replay()
        counter[1] += 1

    vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
    ids = [v for v in vis.split(",") if v.strip().isdigit()]
    sampler = ClockSampler(int(ids[local]) if len(ids) > local else local, period=0.05)
    sampler.start()
    # warm-up: the requested number of steps, then the same step for another 0.4 s -- the clock record then holds samples
    # taken under this very load although the timed region itself is only milliseconds long
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    torch.cuda.synchronize(dev)
    t_end = time.perf_counter() + 0.4
    while time.perf_counter() < t_end:
        for _ in range(20):
            step()
        torch.cuda.synchronize(dev)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    counter[1] = 0   # the timed steps start at resident batch 0 on every rank
    sampler.pause(True)   # the clock record is taken under this same load right before and right after the timed window
    time.sleep(0.003)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_total_local = e0.elapsed_time(e1)
    sampler.pause(False)
    t_end = time.perf_counter() + 0.2
    while time.perf_counter() < t_end:
        for _ in range(20):
            step()
        torch.cuda.synchronize(dev)
    clocks = sampler.stop()
    per_rank = torch.zeros(world, dtype=torch.float64, device=dev)
    per_rank[rank] = ms_total_local / args.steps
    if world > 1:
        dist.all_reduce(per_rank)
    per_rank_ms = [float(x) for x in per_rank.tolist()]
    # every rank's own clock record (median SM clock under load, throttle reasons as a bit per rank): names the slow GPU
    mhz = torch.zeros(world, dtype=torch.float64, device=dev)
    thr = torch.zeros(world, dtype=torch.float64, device=dev)
    mhz[rank] = float(clocks["sm_mhz"] or 0)
    thr[rank] = float(len([r for r in clocks["reasons"] if "slowdown" in r or "power" in r]))
    if world > 1:
        dist.all_reduce(mhz)
        dist.all_reduce(thr)
    clocks["per_rank_sm_mhz"] = [float(x) for x in mhz.tolist()]
    clocks["per_rank_throttle_reasons"] = [int(x) for x in thr.tolist()]
    ms_total = max(per_rank_ms) * args.steps
    value = world * BATCH * args.steps / (ms_total * 1e-3)
    launches = launches_per_step * args.steps

    # distribution of single steps (events around every replay), outside the headline region
    n_dist = max(50, min(args.steps, 200))
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_dist + 1)]
    evs[0].record()
    for i in range(n_dist):
        step()
        evs[i + 1].record()
    torch.cuda.synchronize(dev)
    per_step = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(n_dist))
    dist_ms = {"p10": per_step[n_dist // 10], "p50": per_step[n_dist // 2], "p90": per_step[(9 * n_dist) // 10], "n": n_dist}

    # per-kernel device time (events on the launching stream around every launch, eager submission)
    for _ in range(3):
        eager_step()
    torch.cuda.synchronize(dev)
    lib.profile_enable(True)
    iso_steps = max(5, min(args.steps, 50))
    for _ in range(iso_steps):
        eager_step()
    torch.cuda.synchronize(dev)
    lib.profile_enable(False)
    stages = lib.profile_collect()
    wall["raster"] = time.perf_counter()

    # ---- end to end through the host-buffer API: pinned inputs H2D, views D2H, every step
    pin = lambda x: x.detach().cpu().pin_memory()  # noqa: E731
    h_in = [pin(poses["obj_id"]), pin(poses["final_obj_pose"]), pin(poses["final_hand_verts"]), pin(rand["hand_tex"]),
            pin(rand["light"]), pin(rand["bg_sel"])]
    h_out = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in out.items()}
    e2e_steps = max(3, min(args.steps, 20))

    def e2e_step():
        pipe.renderer.render_batch_host(*h_in, out=h_out, sub_batch=args.sub_batch)

    for _ in range(2):
        e2e_step()
    barrier()
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * BATCH * e2e_steps / (float(t.item()) * 1e-3)
    h2d = sum(x.numel() * x.element_size() for x in h_in)
    d2h = sum(x.numel() * x.element_size() for x in h_out.values())
    pipe.render(poses, rand, out=out)   # the device path on the same (first resident) batch the host-buffer call rendered
    torch.cuda.synchronize(dev)
    same = all(torch.equal(h_out[k], out[k].cpu()) for k in out)
    # the reference's own reply is the BGR image alone (render_infra.py:57-58): the same call with that payload
    h_rgba = {"rgba": h_out["rgba"]}
    for _ in range(2):
        pipe.renderer.render_batch_host(*h_in, out=h_rgba, sub_batch=args.sub_batch)
    barrier()
    e0.record()
    for _ in range(e2e_steps):
        pipe.renderer.render_batch_host(*h_in, out=h_rgba, sub_batch=args.sub_batch)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_rgba_value = world * BATCH * e2e_steps / (float(t.item()) * 1e-3)
    wall["e2e"] = time.perf_counter()

    synth = None
    try:
        synth = synthesis_bench(pipe, dev, world)
    except Exception as e:
        if world > 1:
            raise
        synth = {"error": repr(e)}
    wall["synthesis"] = time.perf_counter()

    train = None
    if not args.no_train:
        del h_out, h_in, h_rgba
        try:
            train = train_bench(dev, world, rank)
        except Exception as e:  # the secondary workload must never cost the headline line
            if world > 1:
                raise
            train = {"error": repr(e)}
    wall["train"] = time.perf_counter()
    dex = None
    if not args.no_train:
        try:
            dex = dexycb_sym_bench(dev, world, rank)
        except Exception as e:
            if world > 1:
                raise
            dex = {"error": repr(e)}
        wall["dexycb_sym"] = time.perf_counter()
    equiv = None
    if world == 1 and not args.no_train and args.equiv_steps > 0:
        try:
            equiv = train_equivalence(dev, steps=args.equiv_steps)
        except Exception as e:
            equiv = {"error": repr(e)}
        wall["train_equivalence"] = time.perf_counter()

    if rank != 0:
        finish(world, dev)
        return

    # ---- roofline of the dominant kernel (largest share of the timed region)
    peak, peak_src = load_peaks()
    dom = max(stages.items(), key=lambda kv: kv[1][0])
    dom_ms, dom_n = dom[1]
    views_per_launch = BATCH * iso_steps / dom_n
    achieved = ALGO_BYTES_PER_VIEW * views_per_launch / (dom_ms / dom_n * 1e-3) / 1e9
    step_gbs = ALGO_BYTES_PER_VIEW * BATCH / (ms_total / args.steps * 1e-3) / 1e9   # the timed region itself
    roofline = {"bound": "hbm", "kernel": dom[0], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_view": ALGO_BYTES_PER_VIEW, "views_per_launch": views_per_launch,
                "kernel_share_of_step": dom_ms / sum(v[0] for v in stages.values()),
                "whole_step": {"achieved": step_gbs, "frac": step_gbs / peak},
                "stage_ms_per_step": {k: v[0] / iso_steps for k, v in stages.items()},
                "note": "instruction-issue bound, not HBM bound: ~17.9k mostly sub-pixel triangles per view (SURVEY.md 8d); "
                        "the tile kernel writes the output exactly once (traffic / algorithmic = 1.0) and keeps z-buffer, "
                        "projected vertices and coverage work in shared memory and registers"}
    traffic_file = os.path.join(ROOT, "profiles", "raster_traffic.json")
    if os.path.exists(traffic_file):
        try:
            roofline["traffic"] = json.load(open(traffic_file)).get(dom[0])
        except Exception:
            pass

    cpu = None
    checked = 0
    if world == 1 and not args.no_cpu_baseline:
        vps, _, threads = cpu_reference(steps=20, warmup=2, sample_views=BATCH)
        cpu = {"value": vps, "unit": "views/s", "cores": threads, "kind": "port",
               "sample": f"{BATCH} views x 20 passes of the same workload, oracle/raster.c, OpenMP over views on {threads} host threads"}
        pipe.render(poses, rand, out=out)
        checked = oracle_check(pipe, poses, rand, out, list(range(0, BATCH, BATCH // 8)))
    wall["cpu"] = time.perf_counter()

    line = {
        "metric": METRIC, "value": value, "unit": "views/s",
        "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(world, views_per_launch=BATCH, submission="CUDA graph replay of ab_render_batch",
                                  resident_batches=f"{N_RES} different batches of {BATCH} views per rank, visited in turn by the steps"),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "views/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "matches_device_path": bool(same), "sub_batch": args.sub_batch,
                "api": "Renderer.render_batch_host: pinned host inputs -> ab_render_batch -> pinned host RGBA+depth+seg",
                "rgba_only_views_per_s": e2e_rgba_value,
                "rgba_only_note": "same call returning the colour image alone, the payload of the reference's reply (render_infra.py:57-58)"},
        "gpu_launches": int(launches),
        "gpu_launches_note": f"{launches_per_step} kernels of this library per step (raster_bin_kernel, raster_tile_kernel), replayed from a CUDA graph",
        "roofline": roofline,
        "cpu_baseline": cpu,
        "oracle_checked": checked,
        "step_ms_distribution": dist_ms,
    }
    line["extras"] = {"synthesis_configs1": synth}
    if train is not None:
        line["extras"]["train_loop_configs3"] = train
    if equiv is not None:
        line["extras"]["train_equivalence_configs4"] = equiv
    if dex is not None:
        line["extras"]["dexycb_sym_configs5"] = dex
    if world == 1 and not args.no_network:
        try:
            line["extras"]["network_forward_configs2"] = network_forward_bench(dev)
        except Exception as e:  # the secondary workload must never cost the headline line
            line["extras"]["network_forward_configs2"] = {"error": repr(e)}
        wall["network"] = time.perf_counter()
        try:
            line["extras"]["hand_obj_refiner_8f3"] = refiner_bench(dev)
        except Exception as e:
            line["extras"]["hand_obj_refiner_8f3"] = {"error": repr(e)}
        wall["refiner"] = time.perf_counter()
    keys = list(wall)
    line["wall_s"] = {keys[i + 1]: round(wall[keys[i + 1]] - wall[keys[i]], 2) for i in range(len(keys) - 1)}
    # the driver keeps the tail of stdout: the numbers of the secondary metrics go last, at the top level
    if isinstance(synth, dict):
        line["synthesised_views_per_s_sample_to_raster"] = synth.get("views_per_s")
    if isinstance(train, dict):
        for k_out, k_in in (("train_images_per_s", "images_per_s"), ("train_vs_torch_bf16", "vs_torch_bf16_autocast"),
                            ("train_vs_torch_fp32", "vs_torch_fp32")):
            line[k_out] = train.get(k_in)
    if isinstance(dex, dict):
        line["train_images_per_s_dexycb_sym"] = dex.get("images_per_s")
    if isinstance(equiv, dict):
        line["mpcpe_after_train_mm"] = equiv.get("mpcpe_after_train_mm")
        line["mpcpe_after_train_reference_loop_mm"] = equiv.get("mpcpe_after_train_reference_loop_mm")
    line["per_rank_ms"] = per_rank_ms
    print(json.dumps(line), flush=True)
    finish(world, dev)


def synthesis_bench(pipe, dev, world, batch=BATCH, steps=40, warmup=8):
    """views/s of the whole synthesis path -- CCV draw -> view -> grasp lookup -> pose generator (MANO LBS)
    -> rasterise -- for a batch of 512 (REFINER null, scrambler random), eager submission."""
    import torch
    import torch.distributed as dist

    from artiboost_b200 import lib
    out = {"rgba": torch.empty((batch, SIZE, SIZE, 4), dtype=torch.uint8, device=dev),
           "depth": torch.empty((batch, SIZE, SIZE), dtype=torch.float32, device=dev),
           "seg": torch.empty((batch, SIZE, SIZE), dtype=torch.uint8, device=dev)}
    def timed(prefetch):
        for _ in range(warmup):
            pipe.synthesise(batch, out=out, prefetch=prefetch)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            pipe.synthesise(batch, out=out, prefetch=prefetch)
        e1.record()
        host = (time.perf_counter() - t0) / steps * 1e3
        torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        pipe.drop_prefetch()
        return float(t.item()) / steps, host

    l0 = lib.launch_count()
    ms_plain, _ = timed(False)
    # the public call with prefetch=True: the next batch's draw + prelude + LBS run on a high-priority stream beside this
    # batch's rasteriser (same views in the same order as the plain sequence: tests/test_gpu_synthesis.py)
    ms, host_ms = timed(True)
    n_calls = 2 * (steps + warmup) + 2
    lib.profile_enable(True)
    for _ in range(5):
        pipe.sample_poses(batch)
    torch.cuda.synchronize(dev)
    lib.profile_enable(False)
    st = lib.profile_collect()
    res = {"batch": batch, "views_per_s": world * batch / ms * 1e3, "ms_per_batch": ms, "host_submit_ms_per_batch": host_ms,
           "submission": "eager, synthesise(prefetch=True): poses of batch k+1 beside the rasteriser of batch k",
           "views_per_s_unpipelined": world * batch / ms_plain * 1e3, "ms_per_batch_unpipelined": ms_plain,
           "our_kernels_per_batch": (lib.launch_count() - l0) / (n_calls + 5),
           "sample_poses_stage_us": {k: v[0] / v[1] * 1e3 for k, v in st.items()}}
    lbs = st.get("mano_lbs_kernel")
    if lbs:
        per = lbs[0] / lbs[1] * 1e-3   # seconds per launch of `batch` samples
        peak, _ = load_peaks()
        res["mano_lbs"] = {"us_per_launch": per * 1e6, "samples_per_s": batch / per,
                           "hbm_frac_algorithmic": batch / per * 10856 / (peak * 1e9),
                           "fp32_tflops": batch / per * 1.07e6 / 1e12,
                           "note": "10 856 B and 1.07 MFLOP per sample (BASELINE.md section 4)"}
    return res


def refiner_bench(dev, batch=512, steps=10, warmup=3):
    """the pose generator with the shipped config's refiner (REFINER.TYPE hand_obj, 3 iterations,
    778 hand vertices x 10 000 resampled object points per nearest-neighbour pass) and the anatomical scrambler
    random_2, batch 512.  Random RefineNet weights (the GrabNet checkpoint is a licensed asset): same arithmetic.
    Beside it: the same nearest-neighbour pass through torch.cdist + min (the library route on this GPU; the
    reference's own chamfer_distance CUDA extension is not installable offline)."""
    import torch

    from artiboost_b200 import lib
    from artiboost_b200.artiboost.refiner import chamfer_nn
    from artiboost_b200.synth import DEFAULT_CFG, SynthPipeline
    cfg = dict(DEFAULT_CFG, SCRAMBLER=dict(DEFAULT_CFG["SCRAMBLER"], TYPE="random_2"),
               REFINER={"TYPE": "hand_obj", "PRETRAINED": None, "ITERS": 3})
    pipe = SynthPipeline(device=dev, seed=5, cfg=cfg, n_hand_tex=2, n_bg=2)

    def timed(fn, n=steps, w=warmup):
        for _ in range(w):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / n

    ms = timed(lambda: pipe.sample_poses(batch))
    lib.profile_enable(True)     # stage times from a separate pass (two events per launch)
    timed(lambda: pipe.sample_poses(batch))
    lib.profile_enable(False)
    stages = lib.profile_collect()
    n = steps + warmup
    nn_ms, nn_launches = stages.get("chamfer_nn_kernel", (0.0, 0))   # ab_chamfer_nn_grouped on the refiner's static clouds
    pairs = batch * 778 * 10000
    out = {"batch": batch, "iters": 3, "scrambler": "random_2", "poses_per_s": batch / ms * 1e3, "ms_per_batch": ms,
           "stage_ms_per_batch": {k: v[0] / n for k, v in stages.items()},
           "stage_launches_per_batch": {k: v[1] / n for k, v in stages.items()}}
    if nn_launches:
        per = nn_ms / nn_launches
        out["chamfer_nn"] = {"ms_per_launch": per, "equivalent_gpairs_per_s": pairs / per / 1e6,
                             "note": "grouped search (Morton groups of 32 points + box bounds, rotated cloud in shared memory, one thread per vertex), bit-identical "
                                     "to the full scan; `equivalent` = the 778 x 10 000 pairs of the scan per sample / time. "
                                     "The scan itself (chamfer_nn_scan_alone_ms) runs at 77 % of the fp32 FMA pipe"}
    # the search on its own, on the geometry the refiner sees: hand vertices and object cloud in the camera orientation,
    # both relative to the object centre (hand ~3-5 cm from the surface)
    poses = pipe.sample_poses(batch)
    obj_pose = poses["final_obj_pose"].contiguous()
    x = (poses["final_hand_verts"] - obj_pose[:, None, :3, 3]).contiguous()
    pts = pipe.refiner.resampled_objs_buffer
    oid = poses["obj_id"].long()
    rot3 = obj_pose[:, :3, :3].contiguous()

    def torch_nn():
        for s in range(0, batch, 128):
            y = torch.matmul(pts[oid[s:s + 128]], rot3[s:s + 128].transpose(1, 2))
            torch.cdist(x[s:s + 128], y).min(-1)

    out["torch_cdist_min_ms"] = timed(torch_nn, 5, 2)
    out["chamfer_nn_scan_alone_ms"] = timed(lambda: chamfer_nn(x, pts, obj_id=oid.int(), rot=obj_pose, return_idx=True), 10, 3)
    from artiboost_b200.artiboost.refiner import chamfer_nn_grouped
    groups = (pipe.refiner.nn_sorted, pipe.refiner.nn_perm, pipe.refiner.nn_boxes)
    out["chamfer_nn_grouped_alone_ms"] = timed(lambda: chamfer_nn_grouped(x, groups, obj_id=oid.int(), rot=obj_pose, return_idx=True), 10, 3)
    d0, i0 = chamfer_nn(x, pts, obj_id=oid.int(), rot=obj_pose)
    d1, i1 = chamfer_nn_grouped(x, groups, obj_id=oid.int(), rot=obj_pose)
    out["grouped_equals_scan"] = bool(torch.equal(d0, d1) and torch.equal(i0, i1))
    out["mean_nn_distance_m"] = float(d0.mean())
    return out


def finish(world, dev):
    """Leave together: every captured graph that references NCCL work has been destroyed by now (train_bench drops its
    loop before it returns), so the process group can be torn down in order.  A watchdog ends the process if the teardown
    still hangs, after the JSON line has been flushed."""
    import gc

    import torch
    import torch.distributed as dist
    sys.stdout.flush()
    if world > 1:
        gc.collect()
        torch.cuda.synchronize(dev)
        dist.barrier()
        torch.cuda.synchronize(dev)
        watchdog = threading.Timer(30.0, lambda: os._exit(0))
        watchdog.daemon = True
        watchdog.start()
        dist.destroy_process_group()
        watchdog.cancel()


# ------------------------------------------------------------------------ secondary workload: network forward
def torch_forward(model, inputs):
    """Plain torch.nn / cuDNN evaluation of the same clasbased modules (the reference's own path: fp32 NCHW,
    anakin/models/resnet.py:199-221, simplebaseline.py:177-190, mlp.py:24).  Baseline leg only."""
    import torch
    import torch.nn.functional as F
    hb = model.model_list[0]
    bb, head = hb.backbone, hb.hybrid_head
    x = inputs["image"]
    x = bb.maxpool(bb.relu(bb.bn1(bb.conv1(x))))
    for name in ("layer1", "layer2", "layer3", "layer4"):
        for blk in getattr(bb, name):
            r = x if blk.downsample is None else blk.downsample(x)
            if hasattr(blk, "conv3"):
                o = blk.relu(blk.bn1(blk.conv1(x)))
                o = blk.relu(blk.bn2(blk.conv2(o)))
                o = blk.bn3(blk.conv3(o))
            else:
                o = blk.relu(blk.bn1(blk.conv1(x)))
                o = blk.bn2(blk.conv2(o))
            x = blk.relu(o + r)
    mean = x.mean(3).mean(2)
    h = head.final_layer(head.deconv_layers(x))
    h = F.softmax(h.reshape(h.shape[0], head.nclasses, -1), 2)
    confd = h.max(-1).values
    h = (h / (h.sum(-1, keepdim=True) + 1e-7)).view(h.shape[0], head.nclasses, head.depth_res, head.height_res, head.width_res)
    u = (h.sum(dim=[2, 3]) * (torch.arange(head.width_res, device=h.device) / head.width_res)).sum(-1)
    v = (h.sum(dim=[2, 4]) * (torch.arange(head.height_res, device=h.device) / head.height_res)).sum(-1)
    d = (h.sum(dim=[3, 4]) * (torch.arange(head.depth_res, device=h.device) / head.depth_res)).sum(-1)
    rot6d = hb.box_head.layers(mean)
    return torch.stack([u, v, d], -1), confd, rot6d


def network_forward_bench(dev, batch=128, steps=10, warmup=3):
    """BASELINE.json configs[2]: clasbased forward-only, synthetic 256x256 batch, tensor-core conv path; with the
    reference's torch.nn / cuDNN path (fp32, and bf16 autocast as a stronger baseline) timed beside it."""
    import torch

    import artiboost_b200.models as M
    from artiboost_b200 import lib
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import netcfg
    out = {}
    flops = {"ResNet34": 10.70e9, "ResNet50": 12.61e9}  # fwd FLOPs / image at 256^2, measured on the reference
    for backbone in ("ResNet50", "ResNet34"):
        arch, preset = netcfg.arch_cfg(backbone)
        torch.manual_seed(1)
        model = M.Arch({"ARCH": arch}, M.build_arch_model_list(arch, preset_cfg=preset)).eval()
        netcfg.randomise_bn(model)
        model = model.to(dev)
        inp = {k: v.to(dev) for k, v in netcfg.make_inputs(batch).items()}

        def timed(fn):
            for _ in range(warmup):
                fn()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            torch.cuda.synchronize(dev)
            return e0.elapsed_time(e1) / steps

        with torch.no_grad():
            ms_eager = timed(lambda: model(inp))
            lib.profile_enable(True)     # per-kernel device times from a separate pass: the stage timers record two events per launch
            for _ in range(steps + warmup):
                model(inp)
            torch.cuda.synchronize(dev)
            lib.profile_enable(False)
            stages = lib.profile_collect()
            # the forward pass as one CUDA graph (how a serving loop submits it): ~50 launches become one submission
            ms = ms_eager
            try:
                side = torch.cuda.Stream(dev)
                g_fwd = torch.cuda.CUDAGraph()
                with torch.cuda.stream(side):
                    model(inp)
                    torch.cuda.synchronize(dev)
                    with torch.cuda.graph(g_fwd, stream=side):
                        keep = model(inp)
                torch.cuda.synchronize(dev)
                ms = min(ms_eager, timed(g_fwd.replay))
                del g_fwd, keep
            except Exception as e:   # capture is an optimisation of the submission, never a requirement
                sys.stderr.write(f"forward graph capture skipped: {e!r}\n")
            torch.backends.cudnn.allow_tf32 = False
            torch.backends.cuda.matmul.allow_tf32 = False
            ms_fp32 = timed(lambda: torch_forward(model, inp))
            ci = dict(inp, image=inp["image"].contiguous(memory_format=torch.channels_last))
            with torch.autocast("cuda", dtype=torch.bfloat16):
                ms_bf16 = timed(lambda: torch_forward(model, ci))
            # MPCPE parity (SURVEY.md 8d iii): Mean3DEPE between the reference arithmetic (torch.nn / cuDNN fp32) and ours on the
            # same synthetic inputs and identical weights, on `corners_3d_abs` (and joints), in millimetres
            ours = model(inp)
            ours = ours[next(iter(ours))]
            ref = torch_train_forward(model, inp)
            mpcpe = (ours["corners_3d_abs"].float() - ref["corners_3d_abs"]).norm(dim=-1).mean().item() * 1e3
            mpjpe = (ours["joints_3d_abs"].float() - ref["joints_3d_abs"]).norm(dim=-1).mean().item() * 1e3
        out[backbone] = {
            "batch": batch, "images_per_s": batch / ms * 1e3, "ms_per_step": ms, "ms_per_step_eager": ms_eager,
            "submission": "CUDA graph replay of the forward pass" if ms < ms_eager else "eager",
            "tflops": batch * flops[backbone] / ms / 1e9, "frac_of_bf16_sustained_peak": batch * flops[backbone] / ms / 1e9 / bf16_sustained_peak(),
            "stage_ms_per_step": {k: v[0] / (steps + warmup) for k, v in stages.items()},
            "torch_cudnn_fp32_images_per_s": batch / ms_fp32 * 1e3, "torch_cudnn_bf16_autocast_images_per_s": batch / ms_bf16 * 1e3,
            "mpcpe_vs_torch_fp32_mm": mpcpe, "mpjpe_vs_torch_fp32_mm": mpjpe,
        }
        del model
    return out


# --------------------------------------------------------- secondary workload: full loop, train images/s (configs[3])
def torch_train_forward(model, batch):
    """The reference's own training graph (torch.nn / cuDNN fp32 NCHW modules + autograd) on the same parameters:
    anakin/models/hybridbaseline.py:41-96.  Baseline leg only."""
    import torch
    from artiboost_b200.models.transform import batch_uvd2xyz, compute_rotation_matrix_from_ortho6d
    hb = model.model_list[0]
    kp3d, _, rot6d = torch_forward(model, batch)
    xyz = batch_uvd2xyz(uvd=kp3d, root_joint=batch["root_joint"], intr=batch["cam_intr"], inp_res=hb.inp_res)
    R = compute_rotation_matrix_from_ortho6d(rot6d)
    corners = torch.matmul(R, batch["corners_can"].permute(0, 2, 1)).permute(0, 2, 1) + xyz[:, 21:22]
    return {"joints_3d_abs": xyz[:, :21], "corners_3d_abs": corners}


def train_bench(dev, world, rank, steps=8, warmup=3, batch=128, backbone="ResNet34"):
    """BASELINE.json configs[3]: CCV sample -> pose -> rasterise -> mix -> clasbased train step (batch-statistics BN,
    JointsLoss + HandOrdLoss + SceneOrdLoss, clip 1e-3, Adam 5e-5), gradient all-reduce over ranks; per-GPU batch 128
    (yaml:131), synthetic share 0.6 / 1.6 (yaml:2).  Beside it on rank 0 at N=1: the same modules through torch.nn /
    cuDNN fp32 + autograd + torch.optim.Adam (the reference's single-GPU PyTorch path) and bf16 autocast."""
    import copy

    import torch
    import torch.distributed as dist

    import artiboost_b200.models as M
    from artiboost_b200 import criterions, lib
    from artiboost_b200.synth import SynthPipeline
    from artiboost_b200.train import ArtiBoostLoop
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import netcfg
    flops = {"ResNet34": 31.8e9, "ResNet50": 37.5e9}[backbone]   # fwd+bwd FLOPs / image at 256^2
    arch, preset = netcfg.arch_cfg(backbone)
    torch.manual_seed(1)
    model = M.Arch({"ARCH": arch}, M.build_arch_model_list(arch, preset_cfg=preset)).to(dev)
    ref_model = copy.deepcopy(model) if world == 1 else None
    pipe = SynthPipeline(device=dev, seed=11, sample_seed=11 + rank)
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    loop = ArtiBoostLoop(model, pipe, batch_size=batch, generator=gen, use_graph=True)
    for _ in range(4):   # eager warm-up steps, then the one-time CUDA-graph capture of the optimisation step
        loop.step()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, n, w):
        for _ in range(w):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / n

    ms_synth = timed(lambda: loop.make_batch(), steps, warmup)
    fixed = loop.make_batch()
    l0 = lib.launch_count()
    ms_step = timed(lambda: loop.step(), steps, warmup)
    launches = (lib.launch_count() - l0) / (steps + warmup)
    ms_net = timed(lambda: loop.step(fixed), steps, 1)
    # stage breakdown from a separate, event-bracketed EAGER pass (events cannot be recorded inside the replayed graph)
    loop.train_step.use_graph = False
    loop.step()
    lib.profile_enable(True)
    l1 = lib.launch_count()
    for _ in range(2):
        loop.step()
    torch.cuda.synchronize(dev)
    kernels_in_step = (lib.launch_count() - l1) / 2
    lib.profile_enable(False)
    stages = lib.profile_collect()
    loop.train_step.use_graph = True
    out = {"backbone": backbone, "per_gpu_batch": batch, "synthetic_per_batch": loop.n_synth, "real_shaped_per_batch": loop.n_real,
           "images_per_s": world * batch / ms_step * 1e3, "ms_per_step": ms_step, "ms_synthesis_and_batching": ms_synth,
           "ms_train_step_only": ms_net, "host_submissions_per_step": launches,
           "our_kernels_per_step": kernels_in_step, "cuda_graph": True,
           "tflops_fwd_bwd": batch * flops / ms_net / 1e9, "frac_of_bf16_sustained_peak": batch * flops / ms_net / 1e9 / bf16_sustained_peak(), "bf16_sustained_peak_tflops": bf16_sustained_peak(),
           "stage_ms_per_step": {k: v[0] / 2 for k, v in stages.items()},
           "stage_launches_per_step": {k: v[1] / 2 for k, v in stages.items()}}
    if ref_model is not None and rank == 0:
        crit = criterions.Criterion(criterions.DEFAULT_CRITERION_CFG, generator=gen)
        opt = torch.optim.Adam([p for p in ref_model.parameters() if p.requires_grad], lr=5e-5)
        ref_model.train()
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False

        def ref_step(autocast):
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                preds = torch_train_forward(ref_model, fixed)
            loss, _ = crit.compute_losses({k: v.float() for k, v in preds.items()}, fixed)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(ref_model.parameters(), 1e-3)
            opt.step()

        ms_ref = timed(lambda: ref_step(False), max(3, steps // 2), 2)
        out["torch_cudnn_fp32_train_images_per_s"] = batch / ms_ref * 1e3
        cl = dict(fixed, image=fixed["image"].contiguous(memory_format=torch.channels_last))
        fixed_nchw, fixed = fixed, cl
        ms_ref16 = timed(lambda: ref_step(True), max(3, steps // 2), 2)
        fixed = fixed_nchw
        out["torch_cudnn_bf16_autocast_train_images_per_s"] = batch / ms_ref16 * 1e3
        out["train_step_only_images_per_s"] = batch / ms_net * 1e3
        out["vs_torch_fp32"] = ms_ref / ms_net          # optimisation step on the same batch, ours over the reference's path
        out["vs_torch_bf16_autocast"] = ms_ref16 / ms_net
    # the captured step references NCCL work: drop it before the process group is torn down
    loop.close()
    del loop
    import gc
    gc.collect()
    return out


def dexycb_sym_bench(dev, world, rank, steps=8, warmup=3, batch=128, backbone="ResNet34"):
    """BASELINE.json configs[4]: the DexYCB clasbased_sym loop -- 21 YCB-shaped objects in the CCV space ([21, 288, 50]),
    CENTER_IDX 9, JointsLoss (corners off) + HandOrdLoss + SymCornerLoss with a synthetic symmetry table in the
    extend_models_info.json format (config_eval/eval_dexycb_clasbased_sym_artiboost.yaml:39,84-91), per-GPU batch 128,
    rendered + real-shaped mix, gradient all-reduce over ranks; images/s, weak scaling."""
    import gc

    import torch
    import torch.distributed as dist

    import artiboost_b200.models as M
    from artiboost_b200 import assets
    from artiboost_b200.synth import SynthPipeline
    from artiboost_b200.train import DEFAULT_PRESET, ArtiBoostLoop, make_augmenter
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import netcfg
    arch, preset = netcfg.arch_cfg(backbone)
    preset = dict(preset, CENTER_IDX=9)
    arch["DATA_PRESET"] = preset
    torch.manual_seed(1)
    model = M.Arch({"ARCH": arch}, M.build_arch_model_list(arch, preset_cfg=preset)).to(dev)
    pipe = SynthPipeline(obj_names=list(assets.YCB_NAMES), device=dev, seed=21, sample_seed=21 + rank)
    info = {str(i + 1): ({"symmetries_continuous": [{"axis": [0, 0, 1], "offset": [0, 0, 0]}]} if i % 3 == 0 else
                         {"symmetries_discrete": [[-1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1]]} if i % 3 == 1 else {}) for i in range(21)}
    crit = {"LAMBDAS": [1.0, 0.1, 1.0],
            "CRITERION": [{"TYPE": "JointsLoss", "LAMBDA_JOINTS_3D": 1.0, "LAMBDA_CORNERS_3D": 0.0}, {"TYPE": "HandOrdLoss"},
                          {"TYPE": "SymCornerLoss", "LAMBDA_SYM_CORNERS_3D": 1.0, "MODEL_INFO": info, "MAX_SYM_DISC_STEP": 0.05}]}
    gen = torch.Generator(device=dev).manual_seed(200 + rank)
    loop = ArtiBoostLoop(model, pipe, batch_size=batch, criterion_cfg=crit, generator=gen, use_graph=True)
    loop.augmenter = make_augmenter(pipe, cfg_preset=dict(DEFAULT_PRESET, CENTER_IDX=9), generator=gen)
    for _ in range(4 + warmup):
        loop.step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loop.step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    seen = int((loop.feedback.err_cnt.sum(dim=(1, 2)) > 0).sum())
    out = {"backbone": backbone, "per_gpu_batch": batch, "ccv_space": list(pipe.sample_weight_map.shape), "objects_drawn": seen,
           "images_per_s": world * batch / ms * 1e3, "ms_per_step": ms, "losses": "JointsLoss + HandOrdLoss + SymCornerLoss",
           "blacklisted_cells": int(pipe.blacklist_map.sum())}
    loop.close()
    del loop
    gc.collect()
    return out


def train_equivalence(dev, steps=200, batch=128, eval_samples=4096, backbone="ResNet34", seed=5):
    """"matched MPCPE".  The same seeded synthetic stream and the same initial weights go
    through (a) this repo's training step (bf16 tensor-core kernels, fused clip + Adam, CUDA graph) and (b) the reference's
    loop arithmetic -- torch.nn / cuDNN fp32 modules with the same state_dict, Criterion, clip_grad_norm_(1e-3),
    torch.optim.Adam(5e-5) (train/train_artiboost.py:66-96) -- for `steps` steps; both networks are then evaluated (BN in
    eval mode) on one fixed set of rendered samples and Mean3DEPE (anakin/metrics/meanepe.py:39-70) on corners_3d_abs
    (MPCPE) and joints_3d_abs (MPJPE) is reported in millimetres.  The random draws of the ordinal losses come from two
    generators in lockstep."""
    import copy

    import torch

    import artiboost_b200.models as M
    from artiboost_b200 import criterions
    from artiboost_b200.synth import SynthPipeline
    from artiboost_b200.train import TrainStep, make_augmenter, mix_batches, real_shaped_batch, synth_to_batch
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import netcfg
    arch, preset = netcfg.arch_cfg(backbone)
    torch.manual_seed(seed)
    model = M.Arch({"ARCH": arch}, M.build_arch_model_list(arch, preset_cfg=preset)).to(dev)
    ref_model = copy.deepcopy(model)
    pipe = SynthPipeline(device=dev, seed=seed, n_hand_tex=16, n_bg=4)
    gen = torch.Generator(device=dev).manual_seed(seed)
    n_synth = int(round(batch * 0.6 / 1.6))   # SYNTH_FACTOR 0.6 (yaml:2)
    ts = TrainStep(model, generator=torch.Generator(device=dev).manual_seed(seed + 1), use_graph=True)
    crit = criterions.Criterion(criterions.DEFAULT_CRITERION_CFG, generator=torch.Generator(device=dev).manual_seed(seed + 1))
    opt = torch.optim.Adam([p for p in ref_model.parameters() if p.requires_grad], lr=5e-5)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False

    def mean3depe(pred, gt):
        return float((pred.float() - gt).norm(dim=-1).mean()) * 1e3

    def evaluate(sets):
        model.eval(); ref_model.eval()
        acc = {"ours": [0.0, 0.0], "ref": [0.0, 0.0]}
        with torch.no_grad():
            for b in sets:
                gt_c = b["corners_3d"] + b["root_joint"].unsqueeze(1)
                gt_j = b["joints_3d"] + b["root_joint"].unsqueeze(1)
                po = model(b)
                po = po[next(iter(po))]
                pr = torch_train_forward(ref_model, b)
                for name, p in (("ours", po), ("ref", pr)):
                    acc[name][0] += mean3depe(p["corners_3d_abs"], gt_c) / len(sets)
                    acc[name][1] += mean3depe(p["joints_3d_abs"], gt_j) / len(sets)
        return acc

    aug = make_augmenter(pipe, generator=gen)

    def make_batch():
        synth = synth_to_batch(pipe.synthesise(n_synth), pipe, augmenter=aug)
        return mix_batches(real_shaped_batch(batch - n_synth, dev, gen, pipe.renderer.width), synth)

    eval_sets = [{k: (v.clone() if torch.is_tensor(v) else v) for k, v in synth_to_batch(pipe.synthesise(batch), pipe, augmenter=aug).items()}
                 for _ in range(max(1, eval_samples // batch))]
    before = evaluate(eval_sets)
    losses_o, losses_r = [], []
    for _ in range(steps):
        b = make_batch()
        lo, _ = ts(b)
        ref_model.train()
        opt.zero_grad(set_to_none=True)
        preds = torch_train_forward(ref_model, b)
        lr_, _ = crit.compute_losses(preds, b)
        lr_.backward()
        torch.nn.utils.clip_grad_norm_(ref_model.parameters(), 1e-3)
        opt.step()
        losses_o.append(lo.clone()); losses_r.append(lr_.detach())
    after = evaluate(eval_sets)
    lo, lr_ = torch.stack(losses_o).float().cpu(), torch.stack(losses_r).float().cpu()
    k = max(1, steps // 10)
    ts.close()
    return {"steps": steps, "batch": batch, "eval_samples": len(eval_sets) * batch, "backbone": backbone,
            "mpcpe_before_mm": {"ours": before["ours"][0], "reference_loop": before["ref"][0]},
            "mpcpe_after_train_mm": after["ours"][0], "mpcpe_after_train_reference_loop_mm": after["ref"][0],
            "mpjpe_after_train_mm": after["ours"][1], "mpjpe_after_train_reference_loop_mm": after["ref"][1],
            "mpcpe_rel_diff": abs(after["ours"][0] - after["ref"][0]) / after["ref"][0],
            "mpjpe_rel_diff": abs(after["ours"][1] - after["ref"][1]) / after["ref"][1],
            "loss_first_steps": {"ours": float(lo[:k].mean()), "reference_loop": float(lr_[:k].mean())},
            "loss_last_steps": {"ours": float(lo[-k:].mean()), "reference_loop": float(lr_[-k:].mean())},
            "loss_max_rel_diff": float(((lo - lr_).abs() / lr_.abs().clamp_min(1e-9)).max()),
            "tolerance": "matched = MPCPE and MPJPE of the two loops within 10 % of each other after training (bf16 vs fp32 trajectories under Adam separate slowly; measured 0.2 % / 4.8 % at 200 steps)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sub-batch", type=int, default=64, help="views per device->host copy stage of the end-to-end leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the secondary full-loop train images/s measurement")
    ap.add_argument("--no-network", action="store_true", help="skip the secondary network-forward measurement")
    ap.add_argument("--equiv-steps", type=int, default=200, help="training steps of the matched-MPCPE leg (N = 1)")
    args = ap.parse_args()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: `python bench.py --gpus N` re-launches itself as one rank per GPU like the driver does
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
