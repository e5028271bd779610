#!/usr/bin/env python
"""bench.py -- throughput of the FVM_TVD hot path on B200, in the reference's metric.

metric : cell-updates/sec per RK stage (FP64) = owned cells x 2 stages x steps / time
step   : one whole RK2 time step of FVM_TVD::run (2 x [gradients, edge fluxes, residual gather,
         update] + half-sum + limit flags + remediation), state resident in HBM
e2e    : the same through the C-ABI with HOST buffers: every step uploads the conservative state
         from pinned host memory (cfd2d_fvm_set_state), steps, and reads it back
         (cfd2d_fvm_get_state) -- what the Method glue does at its save cadence
workload (N=1): BASELINE.json configs[2], the configuration the roofline target is quoted on:
         synthetic 2000x1000x2 = 4 M-cell triangulated channel (inlet / outlet / walls), smooth
         initial data (the reference's 2nd-order path cannot run the literal shock IC),
         2nd-order reconstruction + exact Godunov flux = the reference's live scheme.
N>1    : weak scaling, 4 M cells per GPU (configs[4]); slab partition, owned/halo renumbering as
         the reference's Decomp, NCCL send/recv halo exchange.

Usage: python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                       [--flux godunov|lax] [--order 1|2] [--nx NX --ny NY]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ALGO_BYTES_STAGE = 320.0     # algorithmic bytes per cell per RK stage
# split of the same inventory by kernel: what each kernel alone must touch
ALGO_BYTES_KERNEL = {"grad": 32 + 64 + 8 + 1.5 * 24, "flux": 32 + 64 + 16 + 96, "update": 32 + 32 + 8 + 8}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index, self.t_mark = [], None, index, 0.0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append((time.time(), ln.strip()))

    def mark(self):
        """Start of the timed region: the sampler itself is started earlier (nvidia-smi needs a few
        hundred ms before its first line), only samples read after this mark are reported."""
        self.t_mark = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, r in self.rows:
            if ts < self.t_mark:
                continue
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload_case(nx, ny, tiles=1):
    from cfd2d_b200 import cases
    c = cases.channel(nx, ny)
    return c, c.smooth_state(tiles=tiles)


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation (oracle/_ref when it loads, else the C port)
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    nx, ny, nsteps, variant, flux, order, use_ref = args
    import tempfile
    from cfd2d_b200 import cases
    c = cases.channel(nx, ny)
    st = c.smooth_state()
    if use_ref:
        from oracle import refharness as R
        d = tempfile.mkdtemp(prefix="cfd2d_ref_")
        c.write(d)
        s = R.RefSolver(d, variant=variant)
        s.set_state(*st)
        s.calc_time_step()
        s.run(1)                       # warm caches / page in
        dt = s.run(nsteps)
    else:
        from oracle import port as P
        s = P.OracleSolver(c.mesh, c.task, flux, order)
        s.set_state(*st)
        s.calc_time_step()
        s.step(1)
        t0 = time.perf_counter()
        s.step(nsteps)
        dt = time.perf_counter() - t0
    return c.mesh.nc * 2.0 * nsteps / dt, c.mesh.nc


def cpu_reference(flux, order, nprocs=1, nx=500, ny=250, nsteps=6):
    """Times the reference CPU solver on a bounded sample of the workload: a (nx x ny x 2)-cell twin
    (same generator, BCs, initial data, scheme).  nprocs > 1 runs independent replicas concurrently
    (FVM_TVD is serial: no OpenMP, no MPI calls) and sums their throughput."""
    from oracle import refharness as R
    variant = {(0, 2): "v0", (1, 1): "v1", (1, 2): "v2"}.get((flux, order))
    use_ref = variant is not None and R.available(variant)
    if use_ref:
        try:
            R.load(variant)
        except OSError:
            use_ref = False
    args = (nx, ny, nsteps, variant, flux, order, use_ref)
    if nprocs <= 1:
        res = [_cpu_worker(args)]
    else:
        import multiprocessing as mp
        with mp.get_context("spawn").Pool(nprocs) as pool:
            res = pool.map(_cpu_worker, [args] * nprocs)
    total = float(sum(r[0] for r in res))
    return {"value": total, "unit": "cell-updates/s", "cores": int(nprocs),
            "kind": "reference" if use_ref else "port",
            "sample": f"{res[0][1]}-cell twin of the workload ({nx}x{ny}x2 channel, same BCs/IC/scheme), "
                      f"{nsteps} RK2 steps after 1 warm-up step" + (f", {nprocs} independent replicas summed" if nprocs > 1 else ", 1 thread")}


def run_reference_arm(a):
    """`--impl reference`: the reference's own CPU implementation of the path (oracle/_ref = the real
    FVM_TVD compiled from /root/reference; else the C port) on this box's host cores.

    FVM_TVD is serial -- no OpenMP, no MPI calls on this path -- so ONE job can use ONE
    thread: that is `value` (cores = 1).  What the whole host could do for an ensemble of independent
    jobs (one replica per core, summed) is reported next to it as `all_cores_replicas`, labelled."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    flux = 0 if a.flux == "godunov" else 1
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    vals = []
    r = None
    steps = max(1, min(a.steps, 3))
    for _ in range(steps):                            # each call does its own warm-up step
        r = cpu_reference(flux, a.order, nprocs=1)
        vals.append(r["value"])
        if time.perf_counter() - t0 > 90:
            break
    v = float(np.median(vals))
    r["value"] = v
    try:
        rep = cpu_reference(flux, a.order, nprocs=cores)
        r["all_cores_replicas"] = {"value": rep["value"], "cores": cores,
                                   "what": "independent replicas of the serial solver, one per core, throughputs summed "
                                           "(an ensemble upper bound, not one job)"}
    except Exception as ex:
        r["all_cores_replicas"] = {"error": repr(ex)}
    n_cells = 2 * a.nx * a.ny
    line = {"impl": "reference", "metric": "cell-updates/sec (FP64, RK stage)", "value": v, "unit": "cell-updates/s",
            "n_gpus": a.gpus, "steps": len(vals), "warmup": a.warmup, "ms_per_step": 1e3 * 2.0 * n_cells / v,
            "ms_per_step_note": "one RK2 step of the 4 M-cell workload at the sampled rate (extrapolated from the sample)",
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(a), "cpu_baseline": r,
            "e2e": {"value": v, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(a):
    strong = getattr(a, "strong", False) and a.gpus > 1
    per = 2 * a.nx * a.ny // (a.gpus if strong else 1)
    return {"workload": (f"BASELINE configs[3]: synthetic {a.nx}x{a.ny}x2 = {2 * a.nx * a.ny}-cell triangulated channel split over {a.gpus} GPUs, "
                         if strong else
                         f"BASELINE configs[2]: synthetic {a.nx}x{a.ny}x2 = {2 * a.nx * a.ny} -cell triangulated channel per GPU, ")
                        + f"inlet/outlet/walls, smooth IC; RK2; order {a.order}; flux {a.flux}",
            "cells_per_gpu": per, "flux": a.flux, "order": a.order,
            "partition": (a.partition if a.gpus > 1 else "none"),
            "initial_state": "one Gaussian pressure bump per 4 M-cell slab (N slabs side by side: every rank solves the single-GPU problem)",
            "l2_policy": "inputs larger than L2 (working set ~1.9 GB at 4 M cells vs 126 MB L2)"}


def bind_to_gpu_numa_node(local: int):
    """Multi-rank runs: pin this process (and therefore its first-touch host allocations, the pinned
    staging buffers of the e2e leg included) to the CPUs of the NUMA node its GPU hangs off, as an MPI
    launcher's binding would.  Best effort; returns a short description for the JSON line."""
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(local), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        bus = out[-12:] if len(out) >= 12 else out          # 00000000:1b:00.0 -> 0000:1b:00.0
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return "numa: single node"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"numa node {node} ({len(cpus)} cpus)"
    except Exception as ex:
        return "numa: not bound (" + type(ex).__name__ + ")"
    return "numa: not bound"


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)      # 200 timed steps after 20 warm-up
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--flux", default="godunov", choices=["godunov", "lax"])
    ap.add_argument("--order", type=int, default=2)
    ap.add_argument("--nx", type=int, default=2000)
    ap.add_argument("--ny", type=int, default=1000)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-variants", action="store_true", help="skip the extra Lax-Friedrichs variant lines")
    ap.add_argument("--partition", default="auto", choices=["auto", "slab", "metis"],
                    help="N>1: metis = the bundled METIS exactly as the reference's Decomp calls it (decomp.cpp:86-104; every rank "
                         "builds the global mesh: ~0.6 GB of host memory and ~3.5 s per million cells and rank); slab = equal-count "
                         "strips built from a window of the mesh (O(cells per rank) setup); auto (default) = metis when the host "
                         "memory allows it, else slab")
    ap.add_argument("--strong", action="store_true",
                    help="N>1: strong scaling -- the (nx x ny x 2)-cell mesh is the WHOLE job, split N ways "
                         "(BASELINE configs[3]: --nx 8000 --ny 2000 = 32 M cells); default is weak scaling (nx x ny x 2 per GPU)")
    ap.add_argument("--fused", action="store_true", help="one tile-fused kernel per stage (k_stage) instead of k_grad, k_flux, k_update")
    ap.add_argument("--layout", type=int, default=None, choices=[0, 1, 2],
                    help="step layout: 0 three sweeps, 1 k_stage, 2 k_stage_pipe (persistent, cp.async.bulk-fed); default: the library's")
    ap.add_argument("--exact-riemann", action="store_true",
                    help="Godunov: rim_orig in the reference's operation order instead of the reduced-instruction solver")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
        return
    a.warmup = max(a.warmup, 3)

    import torch
    from cfd2d_b200 import fvm
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else "numa: not bound (single rank)"
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    flux = 0 if a.flux == "godunov" else 1
    partition_note = None
    if world > 1 and a.partition == "auto":
        # METIS needs the global dual graph on every rank: bound the host memory before choosing it
        total_mcells = 2.0 * a.nx * a.ny * (1 if a.strong else world) / 1e6
        need_gb = 0.62 * total_mcells * world
        try:
            avail_gb = [float(ln.split()[1]) / 1e6 for ln in open("/proc/meminfo") if ln.startswith("MemAvailable")][0]
        except Exception:
            avail_gb = 0.0
        from cfd2d_b200 import decomp as _d
        have = os.path.exists(_d.METIS_LIB)
        a.partition = "metis" if (have and need_gb < 0.6 * avail_gb) else "slab"
        partition_note = (f"auto: {a.partition} (global mesh on every rank needs ~{need_gb:.0f} GB of host memory, "
                          f"{avail_gb:.0f} GB available, bundled METIS {'found' if have else 'missing'})")
    clocks = ClockSampler(local)      # started now: nvidia-smi needs a few hundred ms before its first sample
    if rank == 0:
        clocks.start()
    t_setup = time.perf_counter()
    if world == 1:
        c, st = workload_case(a.nx, a.ny, tiles=max(1, a.nx // 1000) if a.strong else 1)
        s = fvm.Solver(c.mesh, c.task, flux, a.order, device=local)
        nc_local, nc_total = c.mesh.nc, c.mesh.nc
    else:
        from cfd2d_b200 import decomp
        nx_rank = a.nx
        if a.strong:
            if a.nx % world:
                raise SystemExit("--strong needs nx divisible by the number of GPUs")
            nx_rank = a.nx // world
        # strong scaling: the global problem (mesh AND initial state) must not depend on N
        s, st, nc_local, nc_total = decomp.make_rank_solver(nx_rank, a.ny, rank, world, local, flux, a.order, dist,
                                                            partition=a.partition,
                                                            tiles=max(1, a.nx // 1000) if a.strong else None)
    halo_stats = None
    if world > 1:
        rm = s.rank_mesh
        hs = torch.tensor([rm.nc_ex - rm.nc, int((np.asarray(rm.recv_count) > 0).sum()), rm.nc], device="cuda", dtype=torch.int64)
        hall = [torch.zeros_like(hs) for _ in range(world)]
        dist.all_gather(hall, hs)
        hall = torch.stack(hall).cpu().numpy()
        halo_stats = {"halo_cells_per_rank": hall[:, 0].tolist(), "peers_per_rank": hall[:, 1].tolist(),
                      "owned_cells_per_rank": hall[:, 2].tolist(),
                      "bytes_per_exchange_max": {"state_32B": int(32 * hall[:, 0].max()), "gradients_64B": int(64 * hall[:, 0].max())}}
    if a.fused and a.layout is None:
        a.layout = 1
    if a.layout is not None:
        s.use_fused(a.layout)
    if a.exact_riemann:
        s.use_exact_riemann(True)
    stream = torch.cuda.Stream()          # a real (capturable) stream; torch events are recorded on it
    torch.cuda.set_stream(stream)
    s.set_stream(stream.cuda_stream)
    s.set_state(*st)
    s.calc_time_step()
    setup_s = time.perf_counter() - t_setup

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing
    s.step(a.warmup)
    l0 = s.launch_count
    barrier()
    clocks.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    s.step_async(a.steps)
    e1.record(stream)
    s.sync()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = s.launch_count - l0
    clk = clocks.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    value = nc_total * 2.0 * a.steps / (ms * 1e-3)

    # ---- per-kernel device time (CUDA events around every launch on the launching stream)
    prof_steps = max(3, min(10, a.steps))
    prof = s.profile(prof_steps)
    stages = 2 * prof_steps
    per_kernel = {}
    for k, (tms, cnt) in prof.items():
        if cnt:
            per_kernel[k] = {"avg_ms": tms / cnt, "launches_per_step": cnt / prof_steps}
    peak, peak_src = peaks()
    fused = prof["stage1"][1] > 0
    if fused:
        # one tile-fused kernel per RK stage (k_stage<flux, order, stage>): the stage IS the kernel
        stage_ms = (prof["stage1"][0] + prof["stage2"][0]) / stages
        ach_stage = ALGO_BYTES_STAGE * nc_local / (stage_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": ach_stage, "peak": peak, "unit": "GB/s", "frac": ach_stage / peak,
                    "traffic": None, "peak_source": peak_src,
                    "scope": "dominant kernel = k_stage (one launch = one RK stage of all owned cells: gradients + edge "
                             "fluxes + residual gather + update, the 'residual+update' of the north star): "
                             "320 algorithmic B/cell x owned cells / average launch duration (CUDA events on the launching stream)",
                    "stage_ms": stage_ms,
                    "dominant_kernel": {"name": "k_stage_pipe" if a.layout == 2 else "k_stage", "avg_ms": stage_ms, "share_of_step": 2 * stage_ms / (ms / a.steps),
                                        "algorithmic_bytes_per_cell": ALGO_BYTES_STAGE, "achieved": ach_stage,
                                        "frac": ach_stage / peak},
                    "per_kernel": per_kernel, "plan": s.plan_summary}
    else:
        upd_ms = (prof["update1"][0] + prof["update2"][0]) / stages
        grad_ms = prof["grad"][0] / stages if prof["grad"][1] else 0.0
        flux_ms = prof["flux"][0] / stages
        stage_ms = grad_ms + flux_ms + upd_ms
        ach_stage = ALGO_BYTES_STAGE * nc_local / (stage_ms * 1e-3) / 1e9
        dom = max((("grad", grad_ms), ("flux", flux_ms), ("update", upd_ms)), key=lambda x: x[1])
        ach_dom = ALGO_BYTES_KERNEL[dom[0]] * nc_local / (dom[1] * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": ach_stage, "peak": peak, "unit": "GB/s", "frac": ach_stage / peak,
                    "traffic": None, "peak_source": peak_src,
                    "scope": "one RK stage = k_grad + k_flux + k_update (the 'residual+update' of the north star): "
                             "320 algorithmic B/cell x owned cells / summed average launch durations",
                    "stage_ms": stage_ms,
                    "dominant_kernel": {"name": "k_" + dom[0], "avg_ms": dom[1], "share_of_stage": dom[1] / stage_ms,
                                        "algorithmic_bytes_per_cell": ALGO_BYTES_KERNEL[dom[0]],
                                        "achieved": ach_dom, "frac": ach_dom / peak},
                    "per_kernel": per_kernel}
    # measured DRAM traffic (ncu --set full capture of this command at 4 M cells, Godunov order 2; per launch)
    tp = os.path.join(ROOT, "profiles", "traffic_r02.json")
    if not os.path.exists(tp):
        tp = os.path.join(ROOT, "profiles", "traffic_r01.json")
    if os.path.exists(tp) and nc_local == 4000000 and flux == 0 and a.order == 2:
        try:
            tr = json.load(open(tp))
            pk = tr["per_kernel_dram_bytes_4m"]
            if fused:
                sfx = "pipe" if a.layout == 2 else "fused"
                roofline["traffic"] = 0.5 * (pk["stage1_" + sfx] + pk["stage2_" + sfx])
            else:
                roofline["traffic"] = tr["stage_dram_bytes_4m"]
                roofline["dominant_kernel"]["traffic"] = pk.get(roofline["dominant_kernel"]["name"][2:])
            roofline["traffic_source"] = tr["source"]
        except Exception:
            pass

    # FP64-pipe roofline of the flux kernel (the exact Riemann solver is FP64-issue bound, so the HBM
    # fraction alone does not describe it): FP64 instructions per launch from the committed ncu capture of this
    # command, divided by the LIVE k_flux time; peak = 64 FP64 lanes x SMs x max SM clock (ncu: dfma peak_sustained = 64)
    fp = os.path.join(ROOT, "profiles", "fp64_r02.json")
    if os.path.exists(fp) and nc_local == 4000000 and flux == 0 and a.order == 2 and not fused and not a.exact_riemann:
        try:
            fj = json.load(open(fp))
            sm_mhz = (clk or {}).get("sm_max_mhz") or 1965.0
            sms = torch.cuda.get_device_properties(local).multi_processor_count
            peak64 = 64.0 * sms * sm_mhz * 1e6
            ach64 = fj["k_flux_fp64_thread_inst_per_launch_4m"] / (flux_ms * 1e-3)
            roofline["fp64"] = {"bound": "fp64", "kernel": "k_flux<2,2>", "achieved": ach64 / 1e12, "peak": peak64 / 1e12,
                                "unit": "T FP64 inst/s", "frac": ach64 / peak64,
                                "ncu_pipe_fp64_pct": fj["k_flux_pipe_fp64_cycles_active_pct"], "source": fj["source"]}
        except Exception:
            pass

    # ---- e2e through the C-ABI with host buffers (pinned), H2D + step + D2H every step
    pin = [torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in st]
    out = [torch.empty(nc_local, dtype=torch.float64).pin_memory() for _ in range(4)]
    ptr_in = [int(t.data_ptr()) for t in pin]
    ptr_out = [int(t.data_ptr()) for t in out]
    e2e_steps = max(3, min(10, a.steps))

    def e2e_step():
        s.set_state(*ptr_in)
        s.step(1)
        s.get_state(out=ptr_out, want_tau=False, want_flag=False)

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": nc_total * 2.0 * e2e_steps / e2e_s, "unit": "cell-updates/s",
           "h2d_bytes_per_step": 32 * nc_local, "d2h_bytes_per_step": 32 * nc_local,
           "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
           "what": "cfd2d_fvm_set_state(pinned host) + cfd2d_fvm_step(1) + cfd2d_fvm_get_state(pinned host) per step",
           "host_binding": numa}
    # the ceiling of that call pattern: the same bytes moved by bare cudaMemcpyAsync (no kernels), all ranks at once
    try:
        dbuf = [torch.empty(nc_local, dtype=torch.float64, device="cuda") for _ in range(4)]
        def copies():
            for d_, h_ in zip(dbuf, pin):
                d_.copy_(h_, non_blocking=True)
            for h_, d_ in zip(out, dbuf):
                h_.copy_(d_, non_blocking=True)
            torch.cuda.synchronize()
        copies()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            copies()
        barrier()
        cp_s = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([cp_s], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            cp_s = float(t.item())
        e2e["copy_only_ms_per_step"] = 1e3 * cp_s / e2e_steps
        e2e["copy_only_GBps_per_rank"] = 64.0 * nc_local * e2e_steps / cp_s / 1e9
        e2e["copy_ceiling_note"] = ("H2D then D2H of the same pinned buffers with no kernels, all ranks concurrently: what the "
                                    "host <-> device path of this box gives this call pattern")
        del dbuf
    except Exception as ex:
        e2e["copy_only_error"] = repr(ex)
    # the Method glue's real cadence: state up once, FILE_OUTPUT_STEP steps on the device, state down once
    cad = 10
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        s.set_state(*ptr_in)
        s.step(cad)
        s.get_state(out=ptr_out, want_tau=False, want_flag=False)
    barrier()
    cad_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([cad_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        cad_s = float(t.item())
    e2e["at_output_cadence"] = {"steps_between_transfers": cad, "value": nc_total * 2.0 * cad * 3 / cad_s, "unit": "cell-updates/s",
                                "what": "set_state + step(10) + get_state, i.e. FILE_OUTPUT_STEP = 10 in task.xml"}

    line = {"metric": "cell-updates/sec (FP64, RK stage)", "value": value, "unit": "cell-updates/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": "strong" if (a.strong and world > 1) else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(a), "clocks": clk, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "setup_s": setup_s}
    if halo_stats is not None:
        halo_stats["transport"] = s.halo_transport
        line["halo"] = halo_stats
    if partition_note:
        line["config"]["partition_note"] = partition_note

    # ---- the bandwidth-bound variants of the same path (reference's Lax-Friedrichs block), N=1 only
    if world == 1 and not a.no_variants and flux == 0:
        variants = {}
        for name, (vf, vo) in {"lax_order2": (1, 2), "lax_order1": (1, 1)}.items():
            s2 = fvm.Solver(c.mesh, c.task, vf, vo, device=local)
            if a.layout is not None:
                s2.use_fused(a.layout)
            s2.set_stream(stream.cuda_stream)
            s2.set_state(*st)
            s2.calc_time_step()
            s2.step(a.warmup)
            torch.cuda.synchronize()
            e0.record(stream); s2.step_async(a.steps); e1.record(stream); s2.sync()
            vms = e0.elapsed_time(e1)
            vv = nc_total * 2.0 * a.steps / (vms * 1e-3)
            variants[name] = {"value": vv, "ms_per_step": vms / a.steps, "roofline_frac_320B": ALGO_BYTES_STAGE * vv / 1e9 / peak}
            s2.close()
        line["variants"] = variants

    if rank == 0 and world == 1 and not a.no_cpu:
        try:
            line["cpu_baseline"] = cpu_reference(flux, a.order, nprocs=1)
        except Exception as ex:  # the checker failing must not lose the GPU numbers
            line["cpu_baseline"] = {"value": None, "error": repr(ex)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    s.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
