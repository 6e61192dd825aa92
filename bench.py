#!/usr/bin/env python
"""bench.py — particle-steps/s of the per-step hot path (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c5] [--impl reference]

A "step" is one Solver::integrate (src/solver.cpp:417-429 of the reference): timestep, predict, tree,
pre-interaction, fluid force, tree gravity, correct.  Workloads (--config; BASELINE.json `configs`):
    c1      shock_tube DIM=1 SSPH cubic spline as shipped (500 particles)
    c2      khi DIM=2 DISPH + artificial conductivity, periodic, N=1152 (995 328 particles)
    c3      gresho_chan_vortex DIM=2 GSPH 2nd-order MUSCL, Wendland C4, N=2048 (4 194 304)
    c4      evrard DIM=3 DISPH + Balsara + time-dependent AV + tree gravity theta=0.5, N=124 (998 592)
    c5      the same physics, N=312 (15 902 832)      <- default, the headline metric
    c5_64m  N=496 (63.9 M)
--gpus N > 1 (launched by torchrun, one rank per GPU): ONE particle set of the same size, split over the ranks by
Morton domain decomposition ("scaling": "strong"); every rank uploads its share and owns 1/N of the state.

value      whole-job particle-steps/s, state resident in HBM, CUDA events on the launching stream, max over ranks.
e2e        the same step through the C ABI with HOST buffers, every step: sphb_upload_aos (pinned host AoS of the
           rank's particles -> device) + sphb_integrate + sphb_download_aos inside the timed region.
e2e.module the reference's module sequence with only the members Solver::predict / correct write going up
           (src/solver.cpp:442-455) and only what each module writes coming down (field masks).
roofline   the dominant kernel of the step (the stage with the largest device time): algorithmic FP64 FLOPs of the
           REFERENCE algorithm on this input (hand-counted constants of SURVEY.md 8d x interaction counts taken by
           the kernels' own counters in an untimed step) / its CUDA-event duration, against the FP64 FMA peak of this
           box (measured once by sphb_bench_fp64 and recorded in MEASURED_FP64.json next to MEASURED_PEAKS.json);
           `traffic` = dram bytes of that kernel from an ncu pass run by this script on the same workload.
cpu_baseline / --impl reference
           the unmodified reference (oracle/_ref, built from /root/reference by oracle/Makefile; the C port when that
           library did not travel) on this box's host cores, thread count set explicitly to all of them, on a bounded
           sample that `config.workload` names.
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

# name -> (sample, overrides, bounded reference sample overrides (None = same size), description)
CONFIGS = {
    "c1": ("shock_tube", dict(N=50), None, "shock_tube DIM=1 SSPH cubic spline, as shipped"),
    "c2": ("khi", dict(N=1152, SPHType="disph", useArtificialConductivity=True), None,
           "khi DIM=2 DISPH + artificial conductivity, periodic box, Wendland C4"),
    "c3": ("gresho_chan_vortex", dict(N=2048, SPHType="gsph", use2ndOrderGSPH=True), dict(N=1024),
           "gresho_chan_vortex DIM=2 GSPH 2nd-order MUSCL, Wendland C4"),
    "c4": ("evrard", dict(N=124), None, "evrard DIM=3 DISPH+Balsara+tdAV+tree gravity theta=0.5, Wendland C4"),
    "c5": ("evrard", dict(N=312), dict(N=124), "evrard DIM=3 DISPH+Balsara+tdAV+tree gravity theta=0.5, Wendland C4"),
    "c5_64m": ("evrard", dict(N=496), dict(N=124), "evrard DIM=3 DISPH+Balsara+tdAV+tree gravity theta=0.5, Wendland C4"),
}


def host_threads():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def flop_constants(dim, sph, gsph2):
    """Algorithmic FP64 operations of the REFERENCE source per interaction (each + - * / sqrt compare = 1), SURVEY.md 8d,
    hand-counted for DIM=3 DISPH there and re-counted here per DIM / formulation by the same rule:
      distance |r_ij| = 3 DIM; kernel w 20, dhw 25, dw 19 + DIM (include/kernel/*.hpp)
      newton     src/pre_interaction.cpp:241-262: distance + w + dhw + 3
      dens       density-loop neighbour (src/pre_interaction.cpp:83-103, d_pre_interaction.cpp:68-93, g_pre_interaction.cpp)
      bal        Balsara-loop neighbour (116-125); GSPH: the MUSCL gradient loop (g_pre_interaction.cpp:100-135)
      pair       force pair incl. viscosity / conductivity (fluid_force.cpp:58-116, d_fluid_force.cpp:58-81,
                 g_fluid_force.cpp:60-166 incl. van Leer limiter and HLL)
      pp, pc     gravity particle-particle / particle-cell (src/bhtree.cpp:312-316, 327-329)"""
    d = dim
    cross = {1: 0, 2: 5, 3: 15}[d]
    k = {"newton": 48 + 3 * d, "pp": 60 + 6 * d, "pc": 6 + 3 * d}
    if sph == "ssph":
        k.update(dens=54 + 6 * d, bal=(3 * d + 19 + d + d + 2 * d + 2 + cross) if d > 1 else 0, pair=71 + 18 * d)
    elif sph == "disph":
        k.update(dens=60 + 6 * d, bal=(3 * d + 19 + d + d + 2 * d + 2 + cross) if d > 1 else 0, pair=87 + 18 * d)
    else:
        k.update(dens=28 + 6 * d, bal=(21 + 11 * d + 2 * d * d) if gsph2 else 0,
                 pair=(103 + 32 * d + 10 * d * d) if gsph2 else (110 + 20 * d))
    return k


def alg_bytes(dim, gravity):
    """Compulsory HBM bytes per particle of each stage kernel with device-resident SoA state, every member touched once
    (SURVEY.md 8d): pre reads pos, vel, mass, dens, ene, sound, alpha, writes sml, dens, pres, gradh, balsara, alpha,
    neighbor; force reads pos, vel, sml, mass, dens, pres, sound, ene, gradh, alpha, balsara, writes acc, dene;
    gravity reads pos, acc, sml (+ the packed x, y, z, m record), writes acc, phi."""
    d = dim
    return {"pre": (2 * d + 11) * 8 + 4, "fluid": (3 * d + 10) * 8, "gravity": (3 * d + 6) * 8 if gravity else 0}


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


def config_params(name, over=None):
    from sphcode_b200 import sample_params
    sample, base, _, _ = CONFIGS[name]
    o = dict(base)
    if over:
        o.update(over)
    return sample_params(sample, **o)


def workload_string(name, n, p):
    sample, base, _, desc = CONFIGS[name]
    return f"{name}: {desc}, sample/{sample} generator N={base['N']} ({n} particles)"


def workload_config(name, n, p, world, ref_note=None):
    cfg = {"workload": workload_string(name, n, p) + (f" [{ref_note}]" if ref_note else ""),
           "neighborNumber": p["neighborNumber"], "leafParticleNumber": p["leafParticleNumber"],
           "l2": "inputs larger than L2 (no flush needed)" if n * 200 > 126e6 else "state smaller than L2: device-resident small case",
           "parallelism": "single GPU" if world == 1 else f"Morton domain decomposition over {world} GPUs (1/{world} of the state per rank, "
                          "halo pulls over NVLink peer memory, NCCL all-reduce of node sums / kernel sizes / dt)"}
    return cfg


def reference_steps(name, over, steps, warmup):
    """Time Solver::integrate of the unmodified reference (or the C port) on ALL host cores of this box."""
    from sphcode_b200 import make_sample
    from oracle import refsim
    p = config_params(name, over)
    parts = make_sample(p)
    dim = p["DIM"]
    flavour = "tree" if refsim.available(dim, "tree") else "port"
    cores = host_threads()
    # explicit thread count: torchrun exports OMP_NUM_THREADS=1 (mirrors omp_set_num_threads, src/solver.cpp:46-58)
    sim = refsim.RefSim(p, parts, dim, flavour, threads=cores)
    sim.initialize()
    for _ in range(warmup):
        sim.integrate()
    t0 = time.perf_counter()
    for _ in range(steps):
        sim.integrate()
    dt = time.perf_counter() - t0
    kind = "reference" if flavour == "tree" else "port"
    return len(parts) * steps / dt, dt / steps * 1e3, sim.threads, kind, len(parts), p


def stock_binary_steps(name, over):
    """The CPU baseline BASELINE.md names: the stock `./sph <sample> <threads>` binary of the unmodified reference
    (oracle/_ref/sph_d<DIM>, built by `make -C oracle stock` against the Boost stand-in), run in a scratch directory on the
    shipped parameter file with N / SPHType edits, all host threads; throughput from its own "calclation time" line
    (src/solver.cpp:313-350: the step loop only).  Two runs: one step to learn dt, then endTime = 2.5 dt (three steps)."""
    import re
    import shutil
    import tempfile
    from oracle import refsim
    from sphcode_b200 import params as P
    sample, base, _, _ = CONFIGS[name]
    o = dict(base)
    o.update(over or {})
    dim = P.SAMPLES[sample][0]
    exe = refsim.stock_binary_path(dim)
    if not os.path.exists(exe):
        return None
    cores = host_threads()
    d = tempfile.mkdtemp(prefix="sphb_stock_")
    try:
        os.makedirs(os.path.join(d, "sample", sample))
        def run(t_end):
            j = dict(P.SHIPPED[sample])
            j.update(o)
            j.update(outputDirectory=os.path.join(d, "results"), endTime=t_end, outputTime=1e9, energyTime=1e9)
            json.dump(j, open(os.path.join(d, "sample", sample, sample + ".json"), "w"))
            shutil.rmtree(os.path.join(d, "results"), ignore_errors=True)
            r = subprocess.run([exe, sample, str(cores)], cwd=d, capture_output=True, text=True, timeout=1200)
            if r.returncode:
                raise RuntimeError(r.stdout[-300:] + r.stderr[-300:])
            ms = float(re.search(r"calclation time: ([0-9.eE+-]+) ms", r.stdout).group(1))
            logs = [f for f in os.listdir(os.path.join(d, "results")) if f.endswith(".log")]
            text = open(os.path.join(d, "results", logs[0])).read()
            loops = re.findall(r"loop: (\d+), time: ([0-9.eE+-]+), dt: ([0-9.eE+-]+), num: (\d+)", text)
            return ms, int(loops[-1][0]), float(loops[-1][2]), int(loops[-1][3])
        _, _, dt, _ = run(1e-12)
        ms, loops, _, n = run(2.5 * dt)
        return {"value": n * loops / (ms * 1e-3), "unit": "particle-steps/s", "cores": cores, "steps": loops, "particles": n,
                "calclation_time_ms": ms, "command": f"oracle/_ref/sph_d{dim} {sample} {cores}  (N={o['N']}, endTime=2.5 dt, no snapshots after t=0)"}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample, base, ref_over, desc = CONFIGS[args.config]
    # bounded: the reference's step costs seconds per million particles; cap warm-up + steps so that the arm ends in minutes
    steps, warmup = args.steps, args.warmup
    value, ms, cores, kind, n, p = reference_steps(args.config, ref_over, steps, warmup)
    p_full = config_params(args.config)
    from sphcode_b200 import make_sample
    n_full = n if not ref_over else None
    note = None if not ref_over else (f"this arm ran the bounded sample N={ref_over['N']} ({n} particles) of the same generator and physics; "
                                      f"the full size needs ~{'17' if args.config == 'c5' else '67'} GB and minutes per step on the host")
    if n_full is None:
        n_full = {"c5": 15902832, "c5_64m": 63902768, "c3": 4194304}.get(args.config, n)
    line = {
        "impl": "reference", "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s",
        "value_per_core": value / max(cores, 1),
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.config, n_full, p_full, max(args.gpus, 1), note),
        "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": cores, "kind": kind,
                         "sample": f"sample/{sample} generator N={(ref_over or base)['N']} ({n} particles), {steps} Solver::integrate steps "
                                   f"after Solver::initialize + {warmup} warm-up, OpenMP threads={cores} (set explicitly)"},
        "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    try:
        line["cpu_baseline"]["stock_sph_binary"] = stock_binary_steps(args.config, ref_over)
    except Exception as e:
        line["cpu_baseline"]["stock_sph_binary"] = {"error": str(e)[:200]}
    print(json.dumps(line))


def fp64_peak(local):
    """FP64 FMA peak of this box: measured once on the idle GPU and recorded next to MEASURED_PEAKS.json."""
    from sphcode_b200 import lib
    path = os.path.join(ROOT, "MEASURED_FP64.json")
    try:
        import torch
        name = torch.cuda.get_device_name(local)
    except Exception:
        name = "?"
    try:
        j = json.load(open(path))
        if j.get("gpu_name") == name and j.get("fp64_tflops"):
            return j["fp64_tflops"], "MEASURED_FP64.json (sphb_bench_fp64 on this box, idle GPU, before the run)"
    except Exception:
        pass
    v = lib.fp64_peak_tflops(local)
    try:
        json.dump({"fp64_tflops": v, "gpu_name": name, "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
                   "how": "sphb_bench_fp64: 8 independent DFMA chains per thread, 148 x 8 blocks x 256 threads, best of 5, CUDA events"},
                  open(path, "w"), indent=1)
    except Exception:
        pass
    return v, "sphb_bench_fp64 measured at the start of this run on the idle GPU (written to MEASURED_FP64.json)"


def ncu_traffic(config, kernel_regex, timeout=600):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel on this workload: a child
    process of this script (--traffic-probe) run under ncu.  None when ncu is unavailable or fails."""
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None, "ncu not found"
    log = os.path.join(ROOT, "gpurun_out", "bench_traffic.csv")
    os.makedirs(os.path.dirname(log), exist_ok=True)
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k", f"regex:{kernel_regex}",
           "-s", "2", "-c", "1", "--csv", "--log-file", log, sys.executable, os.path.abspath(__file__), "--traffic-probe", "--config", config]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
        if r.returncode:
            return None, f"ncu rc={r.returncode}: {(r.stderr or r.stdout)[-200:]}"
        tot = 0.0
        import csv
        rows = [row for row in csv.reader(open(log)) if len(row) > 5]
        hdr = next(row for row in rows if "Metric Name" in row)
        iname, iunit, ival = hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
        for row in rows:
            if row[iname] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                v = float(row[ival].replace(",", ""))
                tot += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(row[iunit], 1)
        return (tot or None), "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum on one launch, this run"
    except Exception as e:
        return None, f"ncu failed: {str(e)[:200]}"


def traffic_probe(args):
    """child of ncu_traffic: a few steps of the workload, nothing printed"""
    from sphcode_b200 import make_sample, lib
    p = config_params(args.config)
    parts = make_sample(p)
    c = lib.Context(p, p["DIM"], device=0)
    c.upload(parts)
    c.initialize()
    for _ in range(3):
        c.integrate()
    c.synchronize()


def state_digest(ctx, dist, world):
    """A digest of the physical state that does not depend on the decomposition (to rounding): energy sums of
    src/output.cpp:72-83 (all-reduced inside libsphb)."""
    e = ctx.energy()
    return {"kinetic": float(e[0]), "thermal": float(e[1]), "potential": float(e[2]), "total": float(e.sum())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="c5", choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="sphb", choices=["sphb", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-traffic", action="store_true")
    ap.add_argument("--traffic-probe", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()

    if args.traffic_probe:
        return traffic_probe(args)
    if args.impl == "reference":
        return run_reference(args)

    import ctypes
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
    peak64, peak64_src = (fp64_peak(local) if rank == 0 else (None, None))
    if world > 1:
        dist.barrier()

    p = config_params(args.config)
    dim = p["DIM"]
    if world > 1 and p["SPHType"] == "gsph":
        raise SystemExit("GSPH is single-GPU only")
    parts = make_sample(p)
    n_glob = len(parts)
    # this rank's share of the generator's output (contiguous chunk; the first tree build ships every particle to its owner)
    lo, hi = n_glob * rank // world, n_glob * (rank + 1) // world
    parts = parts[lo:hi].copy() if world > 1 else parts
    rec = parts.dtype.itemsize
    ctx = lib.Context(p, dim, device=local)
    stream = torch.cuda.current_stream()
    ctx.L.sphb_set_stream(ctx._c, stream.cuda_stream)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(lib.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.set_distributed_id(rank, world, bytes(uid.cpu().numpy().tobytes()))

    # pinned host AoS buffer = what a reference Simulation would hold (room for the rank's count to drift by migration)
    cap = int(len(parts) * 1.35) + 8192 if world > 1 else len(parts)
    hptr = ctx.L.sphb_host_alloc(cap * rec)
    if not hptr:
        raise SystemExit("cudaMallocHost failed")
    ctypes.memmove(hptr, parts.ctypes.data, len(parts) * rec)
    n_up = len(parts)
    del parts

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    ctx.upload_raw(hptr, n_up)
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
    value = n_glob * args.steps / (ms * 1e-3)
    digest = state_digest(ctx, dist, world)
    n_loc = ctx.local_n
    loc = torch.tensor([n_loc, -n_loc], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(loc, op=dist.ReduceOp.MAX)
    n_loc_max, n_loc_min = int(loc[0].item()), int(-loc[1].item())
    halo_records, migrated = ctx.halo_records, ctx.migrated
    # every rank's stage times (the collectives make a rank wait for the slowest one: the table shows where)
    names = list(lib.T_NAMES)
    st = torch.tensor([stage_ms[k] / args.steps for k in names], dtype=torch.float64, device="cuda")
    allst = [torch.zeros_like(st) for _ in range(world)]
    if world > 1:
        dist.all_gather(allst, st)
    else:
        allst = [st]
    stage_by_rank = {k: [round(float(a[i].item()), 3) for a in allst] for i, k in enumerate(names)}

    # ---- end to end through the C ABI with host buffers
    e2e = None
    if not args.no_e2e:
        F = lib
        def full_step():
            n = ctx.local_n
            ctx.upload_raw(hptr, n)
            ctx.integrate()
            ctx.download_raw(hptr)
            return n, ctx.local_n
        ctx.download_raw(hptr)                      # the host copy of the rank's current particles
        full_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        ksteps = max(2, min(args.steps, 3))
        up_b = dn_b = 0
        for _ in range(ksteps):
            a, b = full_step()
            up_b += a * rec
            dn_b += b * rec
        e1.record(stream)
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        ems = max(e0.elapsed_time(e1), wall)      # copies are synchronous host calls: take the larger clock
        te = torch.tensor([ems, up_b / ksteps, dn_b / ksteps], dtype=torch.float64, device="cuda")
        if world > 1:
            tm = te[:1].clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dist.all_reduce(te[1:])
            te[0] = tm[0]
        e2e = {"value": n_glob * ksteps / (float(te[0].item()) * 1e-3), "unit": "particle-steps/s",
               "h2d_bytes_per_step": int(te[1].item()), "d2h_bytes_per_step": int(te[2].item()), "steps": ksteps,
               "api": "per rank: sphb_upload_aos(SPHB_F_ALL) of its particles + sphb_integrate + sphb_download_aos(SPHB_F_ALL), pinned host AoS"}
        # module mode (single GPU): what the reference Solver would move around its host-side predict / correct with the
        # sph::gpu Modules in its slots — masked transfers, in place over PCIe
        if world == 1:
            up_mask = F.F_POS | F.F_VEL | F.F_ENE | F.F_SOUND                       # Solver::predict writes (vel_p, ene_p stay on the host)
            pre_mask = F.F_SML | F.F_DENS | F.F_PRES | F.F_GRADH | F.F_BALSARA | F.F_ALPHA | F.F_NEIGHBOR
            def module_step():
                ctx.timestep()
                ctx.upload_raw(hptr, n_glob, up_mask)
                ctx.make_tree(); ctx.pre()
                ctx.download_raw(hptr, pre_mask)
                ctx.fluid()
                ctx.download_raw(hptr, F.F_ACC | F.F_DENE)
                ctx.gravity()
                ctx.download_raw(hptr, F.F_ACC | F.F_PHI)
            module_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(ksteps):
                module_step()
            barrier()
            mms = (time.perf_counter() - t0) * 1e3
            d = dim
            e2e["module"] = {"value": n_glob * ksteps / (mms * 1e-3), "unit": "particle-steps/s",
                             "h2d_bytes_per_step": n_glob * (2 * d + 2) * 8, "d2h_bytes_per_step": n_glob * ((6 * 8 + 4) + (d + 1) * 8 + (d + 1) * 8),
                             "api": "sphb_timestep; sphb_upload_aos(POS|VEL|ENE|SOUND); sphb_make_tree; sphb_pre_interaction; sphb_download_aos(pre outputs); "
                                    "sphb_fluid_force; sphb_download_aos(ACC|DENE); sphb_gravity_force; sphb_download_aos(ACC|PHI) — the calls of "
                                    "sph::gpu::{TimeStep,PreInteraction,FluidForce,GravityForce}::calculation; host-side predict / correct not included"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    K = flop_constants(dim, p["SPHType"], p["use2ndOrderGSPH"])
    npart = cnt["n_particles"]                      # this rank's particles (counters are per rank)
    flops = {"pre": K["newton"] * cnt["newton_evals"] + (K["dens"] + K["bal"]) * cnt["pre_neighbors"],
             "fluid": K["pair"] * cnt["force_pairs"],
             "gravity": K["pp"] * cnt["grav_pp"] + K["pc"] * cnt["grav_pc"]}
    per_step = {k: v / args.steps for k, v in stage_ms.items()}
    dom = max(("pre", "fluid", "gravity"), key=lambda k: per_step[k])
    kern = {"pre": "k_pre_interaction", "fluid": "k_fluid_force", "gravity": "k_gravity"}[dom]
    B = alg_bytes(dim, p["useGravity"])
    dom_ms = per_step[dom]
    achieved = flops[dom] / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
    step_flops = sum(flops.values())
    traffic, traffic_src = None, "not measured"
    if world == 1 and not args.no_traffic:
        traffic, traffic_src = ncu_traffic(args.config, kern)
    roofline = {
        "bound": "fp64", "kernel": kern, "achieved": achieved, "peak": peak64, "unit": "TFLOP/s",
        "frac": achieved / peak64 if peak64 else None, "traffic": traffic, "traffic_source": traffic_src,
        "peak_source": peak64_src,
        "alg_flops_per_launch": flops[dom], "ms_per_launch": dom_ms, "flop_constants": K,
        "hbm": {"achieved": B[dom] * npart / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0, "peak": hbm_peak, "unit": "GB/s",
                "peak_source": hbm_src, "alg_bytes_per_launch": B[dom] * npart},
        "kernels": {k: {"ms": per_step[k], "alg_tflops": flops[k] / (per_step[k] * 1e-3) / 1e12 if per_step[k] > 0 else 0.0,
                        "frac_of_fp64_peak": flops[k] / (per_step[k] * 1e-3) / 1e12 / peak64 if per_step[k] > 0 and peak64 else None}
                    for k in ("pre", "fluid", "gravity")},
        "step": {"alg_flops_per_particle_step": step_flops / max(npart, 1),
                 "achieved_tflops_this_rank": step_flops / (ms / args.steps * 1e-3) / 1e12,
                 "frac_of_fp64_peak": step_flops / (ms / args.steps * 1e-3) / 1e12 / peak64 if peak64 else None},
        "stage_ms_per_step": per_step,
        "interactions_per_particle": {k: cnt[k] / max(npart, 1) for k in
                                      ("newton_evals", "newton_iters", "pre_candidates", "pre_neighbors", "force_pairs",
                                       "grav_pp", "grav_pc", "grav_node_visits")},
    }

    cpu = None
    if not args.no_cpu_baseline:
        try:
            _, _, ref_over, _ = CONFIGS[args.config]
            v, cms, cores, kind, nref, _ = reference_steps(args.config, ref_over, 2, 1)
            cpu = {"value": v, "unit": "particle-steps/s", "cores": cores, "kind": kind, "ms_per_step": cms, "value_per_core": v / max(cores, 1),
                   "sample": f"sample/{CONFIGS[args.config][0]} generator N={(ref_over or CONFIGS[args.config][1])['N']} ({nref} particles), 2 Solver::integrate steps "
                             f"after Solver::initialize + 1 warm-up, OpenMP threads={cores} (set explicitly)"}
            try:
                cpu["stock_sph_binary"] = stock_binary_steps(args.config, ref_over)
            except Exception as e:
                cpu["stock_sph_binary"] = {"error": str(e)[:200]}
        except Exception as e:  # the checker libraries did not travel
            cpu = {"value": None, "unit": "particle-steps/s", "cores": 0, "kind": "unavailable", "sample": str(e)[:200]}

    line = {
        "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.config, n_glob, p, world),
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "nonconverged_newton": ctx.nonconverged,
        "state_digest": digest,
        "decomposition": {"particles_per_rank_min": n_loc_min, "particles_per_rank_max": n_loc_max,
                          "halo_records_rank0": halo_records, "migrated_rank0": migrated,
                          "stage_ms_per_step_by_rank": stage_by_rank if world > 1 else None},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
