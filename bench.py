#!/usr/bin/env python
"""bench.py — particle-steps/s of the per-step hot path (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n-side S] [--impl reference]

A "step" is one Solver::integrate (src/solver.cpp:417-429 of the reference): timestep, predict,
tree, pre-interaction, fluid force, tree gravity, correct — on the 3-D Evrard sphere (DISPH +
Balsara + time-dependent AV + tree gravity theta = 0.5, Wendland C4), the reference's own
sample/evrard generator scaled to N=312 (15.9 M particles), device-resident state.

value      whole-job particle-steps/s, state resident in HBM, CUDA events on the launching stream,
           max over ranks.
e2e        the same step through the C ABI with HOST buffers: sphb_upload_aos (pinned host AoS ->
           device) + sphb_integrate + sphb_download_aos inside the timed region, every step.
roofline   the dominant kernel (k_gravity): algorithmic FP64 FLOPs of the reference algorithm on this
           input (78 per particle-particle + 15 per particle-cell interaction; counted
           by the kernel's own counters in an untimed step) / its CUDA-event duration, against the
           FP64 FMA peak measured here by sphb_bench_fp64 (MEASURED_PEAKS.json has no FP64 entry);
           the HBM figure (algorithmic bytes / duration vs MEASURED_PEAKS.json hbm_gbs) sits beside it.
cpu_baseline / --impl reference
           the unmodified reference (oracle/_ref, built from /root/reference by oracle/Makefile) — or
           the C port when that library did not travel — on this box's host cores, all threads, on a
           bounded Evrard sample of the same physics.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def evrard_params(n_side):
    from sphcode_b200 import sample_params
    return sample_params("evrard", N=n_side)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def reference_steps(n_side, steps, warmup):
    """Time Solver::integrate of the unmodified reference (or the C port) on host cores."""
    from sphcode_b200 import make_sample
    from oracle import refsim
    p = evrard_params(n_side)
    parts = make_sample(p)
    flavour = "tree" if refsim.available(3, "tree") else "port"
    sim = refsim.RefSim(p, parts, 3, flavour)
    sim.initialize()
    for _ in range(warmup):
        sim.integrate()
    t0 = time.perf_counter()
    for _ in range(steps):
        sim.integrate()
    dt = time.perf_counter() - t0
    kind = "reference" if flavour == "tree" else "port"
    return len(parts) * steps / dt, dt / steps * 1e3, sim.threads, kind, len(parts)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_side = args.ref_n_side
    value, ms, cores, kind, n = reference_steps(n_side, args.steps, args.warmup)
    p = evrard_params(n_side)
    line = {
        "impl": "reference", "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.n_side, None, p),
        "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": cores, "kind": kind,
                         "sample": f"sample/evrard generator N={n_side} ({n} particles), {args.steps} Solver::integrate steps "
                                   f"after Solver::initialize + {args.warmup} warm-up, OpenMP threads={cores}"},
        "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(n_side, n, p):
    return {"workload": f"evrard DIM=3 DISPH+Balsara+tdAV+tree gravity theta={p['theta']}, Wendland C4, "
                        f"sample/evrard generator N={n_side}" + (f" ({n} particles)" if n else " (15.9M particles)"),
            "neighborNumber": p["neighborNumber"], "leafParticleNumber": p["leafParticleNumber"],
            "l2": "inputs larger than L2 (no flush needed)", "parallelism": "replicated state, Morton-slice compute"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n-side", type=int, default=312, help="sample/evrard N (312 -> 15.9 M particles)")
    ap.add_argument("--ref-n-side", type=int, default=100, help="bounded CPU sample (100 -> 523 k particles)")
    ap.add_argument("--impl", default="sphb", choices=["sphb", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from sphcode_b200 import make_sample, lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libsphb has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    p = evrard_params(args.n_side)
    parts = make_sample(p)
    n = len(parts)
    ctx = lib.Context(p, 3, device=local)
    stream = torch.cuda.current_stream()
    ctx.L.sphb_set_stream(ctx._c, stream.cuda_stream)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(lib.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.set_distributed_id(rank, world, bytes(uid.cpu().numpy().tobytes()))

    # pinned host AoS buffer = what a reference Simulation would hold
    nbytes = n * parts.dtype.itemsize
    hptr = ctx.L.sphb_host_alloc(nbytes)
    if not hptr:
        raise SystemExit("cudaMallocHost failed")
    import ctypes
    ctypes.memmove(hptr, parts.ctypes.data, nbytes)
    del parts

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    ctx.upload_raw(hptr, n)
    ctx.initialize()
    # interaction counts of the reference algorithm on this input (untimed, counters on)
    ctx.enable_counters(True)
    ctx.integrate()
    cnt = ctx.counters()
    ctx.enable_counters(False)
    for _ in range(max(args.warmup - 1, 0)):
        ctx.integrate()

    sampler = ClockSampler(local)
    ctx.enable_timers(True)
    stage_ms = {k: 0.0 for k in lib.T_NAMES}
    l0 = ctx.launches
    barrier()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        ctx.integrate()
        for k, v in ctx.timers().items():
            stage_ms[k] += v
    ev1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launches - l0
    ctx.enable_timers(False)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = n * args.steps / (ms * 1e-3)

    # end to end through the C ABI with host buffers
    e2e = None
    if not args.no_e2e:
        for _ in range(1):
            ctx.upload_raw(hptr, n); ctx.integrate(); ctx.download_raw(hptr)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        ksteps = max(2, min(args.steps, 3))
        for _ in range(ksteps):
            ctx.upload_raw(hptr, n)
            ctx.integrate()
            ctx.download_raw(hptr)
        e1.record(stream)
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        ems = max(e0.elapsed_time(e1), wall)      # copies are synchronous host calls: take the larger clock
        te = torch.tensor([ems], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": n * ksteps / (float(te.item()) * 1e-3), "unit": "particle-steps/s",
               "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes, "steps": ksteps,
               "api": "sphb_upload_aos(SPHB_F_ALL) + sphb_integrate + sphb_download_aos(SPHB_F_ALL), pinned host AoS"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # roofline of the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    fp64_peak = lib.fp64_peak_tflops(local)
    npart = cnt["n_particles"]
    # per-rank share of the work when world > 1 (equal particle slices)
    share = 1.0 / world
    grav_flops = (78.0 * cnt["grav_pp"] + 15.0 * cnt["grav_pc"])           # this rank's slice (counters are per rank)
    grav_ms = stage_ms["gravity"] / args.steps
    grav_bytes = 96.0 * npart * share                                        # 96 B/particle for gravity
    achieved = grav_flops / (grav_ms * 1e-3) / 1e12 if grav_ms > 0 else 0.0
    sph_flops = 57.0 * cnt["newton_evals"] + 78.0 * cnt["pre_neighbors"] + 57.0 * cnt["pre_neighbors"] + 141.0 * cnt["force_pairs"]
    step_flops = grav_flops + sph_flops
    # DRAM traffic of the dominant kernel per launch, from the committed `ncu --set full` capture of this
    # workload (profiles/ncu_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum); 1-GPU figure
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if world == 1 and tj.get("n_side") == args.n_side:
            traffic = tj["k_gravity"]["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {
        "bound": "fp64", "kernel": "k_gravity<3,false,false>", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
        "frac": achieved / fp64_peak if fp64_peak else None, "traffic": traffic,
        "peak_source": "FP64 FMA micro-benchmark sphb_bench_fp64, measured in this run",
        "alg_flops_per_launch": grav_flops, "ms_per_launch": grav_ms,
        "hbm": {"achieved": grav_bytes / (grav_ms * 1e-3) / 1e9 if grav_ms > 0 else 0.0, "peak": hbm_peak, "unit": "GB/s",
                "peak_source": hbm_src, "alg_bytes_per_launch": grav_bytes},
        "step": {"alg_flops_per_particle_step": step_flops / (npart * share), "achieved_tflops": step_flops * world / (ms / args.steps * 1e-3) / 1e12,
                 "frac_of_fp64_peak": step_flops / (ms / args.steps * 1e-3) / 1e12 / fp64_peak if fp64_peak else None},
        "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()},
        "interactions_per_particle": {k: cnt[k] / (npart * share) for k in
                                      ("newton_evals", "newton_iters", "pre_candidates", "pre_neighbors", "force_pairs",
                                       "grav_pp", "grav_pc", "grav_node_visits")},
    }

    cpu = None
    if not args.no_cpu_baseline:
        try:
            v, cms, cores, kind, nref = reference_steps(args.ref_n_side, 2, 1)
            cpu = {"value": v, "unit": "particle-steps/s", "cores": cores, "kind": kind, "ms_per_step": cms,
                   "sample": f"sample/evrard generator N={args.ref_n_side} ({nref} particles), 2 Solver::integrate steps after "
                             f"Solver::initialize + 1 warm-up, OpenMP threads={cores}"}
        except Exception as e:  # the checker libraries did not travel
            cpu = {"value": None, "unit": "particle-steps/s", "cores": 0, "kind": "unavailable", "sample": str(e)[:200]}

    line = {
        "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.n_side, n, p),
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "nonconverged_newton": ctx.nonconverged,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
