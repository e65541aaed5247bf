#!/usr/bin/env python
"""bench.py — sigma builds per second of the effective Hamiltonian (Heff::makeHeff) on B200.

One STEP = one sigma build sigma = H_eff * S of the two-site problem at the middle site pair of the named workload
(all diagram families 1-5; operators resident in HBM; S changes every step like in the Davidson loop).

  value     whole-job sigma builds / s with S and sigma resident in HBM (b2_heff_apply_device), CUDA events, max over ranks
  e2e       the same through the host-buffer C-ABI call a CheMPS2 shim makes (b2_heff_apply: pinned H2D of S, kernels, D2H of sigma)
  roofline  algorithmic FP64 FLOPs of one sigma build (2mnk per reference dgemm_) / measured kernel time,
            against the FP64 tensor (DMMA) peak measured on this GPU in this run (MEASURED_PEAKS.json has no FP64 entry)
  N > 1     operator-ownership sharding (MPIchemps2.h owner maps with mpi_size -> N GPUs): every rank applies the terms it
            owns to the replicated S and the partial sigma vectors are summed with an NCCL all-reduce  => strong scaling

--impl reference times the UNMODIFIED reference (oracle/_ref, built from /root/reference by oracle/build_ref.sh) on the host
cores on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "heff_sigma_builds_per_s"
UNIT = "sigma-builds/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="synth40")
    ap.add_argument("--D", type=int, default=None)
    ap.add_argument("--dist", default="gauss", choices=["gauss", "flat"])
    ap.add_argument("--ref-budget-s", type=float, default=420.0, help="--impl reference: stop starting new reference builds after this many seconds")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the secondary metric (DMRG sweep time at D on N2/cc-pVDZ, config 2)")
    ap.add_argument("--no-update", action="store_true", help="skip the operator-update metric")
    ap.add_argument("--sweep-D", default=None, help="comma-separated bond dimensions of the sweep metric, one full sweep each (default 250,500,1000)")
    ap.add_argument("--sweep-ref", action="store_true", help="also time the unmodified reference's DMRG::Solve on the same schedule (minutes)")
    ap.add_argument("--work-budget", type=float, default=0)
    ap.add_argument("--chunk-k", type=float, default=0)
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks / throttle reasons while the timed region runs"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.2)

    def result(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def workload_and_dims(args, device):
    from chemps2_b200 import workloads
    w = workloads.get(args.workload, D=args.D)
    ctx = w.context(device)
    if args.work_budget:
        ctx.set_option("work_budget", args.work_budget)
    if args.chunk_k:
        ctx.set_option("chunk_k", args.chunk_k)
    dims = w.apply_distribution(ctx, args.dist)
    return w, ctx, dims


def host_mem_available_gb():
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                return int(ln.split()[1]) / 1048576.0
    except OSError:
        pass
    return 0.0


def cpu_baseline(args, w, dims, flops, arena_doubles, threads=None):
    """ONE complete Heff::makeHeff of the UNMODIFIED reference (oracle/_ref) on the host cores at EXACTLY the workload the GPU arm
    runs (same D, same sector table, same hash-filled operators, same input vector): measured, not extrapolated.  Its output vector
    is returned for the full-size parity check.  The reference holds both operator tables in host memory (8 bytes x arena_doubles)."""
    from oracle import refrun
    need_gb = 8.0 * arena_doubles / 2 ** 30 * 1.08 + 6.0
    have_gb = host_mem_available_gb()
    if have_gb and have_gb < need_gb:
        raise MemoryError(f"the reference needs ~{need_gb:.0f} GB of host memory for its operator tables at D={w.D}, {have_gb:.0f} GB available")
    t0 = time.time()
    ref = refrun.run_reference_synth(w, 7, reps=1, dims=dims, threads=threads)
    gflops = flops / ref["best_s"] / 1e9
    sample = (f"ONE complete reference Heff::makeHeff (oracle/_ref: unmodified CheMPS2 + OpenBLAS, OpenMP over target blocks, {ref['threads']} threads) "
              f"on this very workload at D={w.D} (veclength {ref['veclength']}): {flops / 1e9:.1f} GFLOP in {ref['best_s']:.2f} s = {gflops:.1f} GFLOP/s; "
              f"measured, not extrapolated (process wall {time.time() - t0:.0f} s incl. operator fill)")
    return {"value": 1.0 / ref["best_s"], "unit": UNIT, "cores": ref["threads"], "kind": "reference", "sample": sample, "gflops": gflops,
            "sample_seconds": ref["best_s"]}, ref


SWEEP_WORKLOAD = "n2_ccpvdz"                       # BASELINE config 2: N2/cc-pVDZ 14e/28o, D2h, orbitals reordered (the reference's own FCIDUMP)
SWEEP_SCHEDULE = [(250, 1), (500, 1), (1000, 1)]   # (D, full sweeps); rtol 1e-5, no noise, seeded random MPS (srand(12345))
SWEEP_SEED = 12345


def sweep_metric(device, with_reference, schedule=None, world=1, rank=0, allreduce=None):
    """Secondary metric of BASELINE.json ("DMRG sweep time at D"): a complete DMRG calculation — the reference's seeded random MPS, PreSolve,
    scheduled two-site sweeps (Join, device Davidson, device-SVD Split, operator updates) — on N2/cc-pVDZ (config 2), timed by wall clock
    around b2_dmrg_sweep.  world > 1: sigma terms and operator updates sharded over the GPUs (b2_dmrg_set_world), the rest replicated.
    with_reference: the same start (same rand() stream) and schedule through DMRG::Solve of the unmodified reference (oracle/_ref) on the host
    cores of this box; its per-sweep wall times are the reference's own printed timers."""
    from chemps2_b200 import api, workloads
    schedule = schedule or SWEEP_SCHEDULE
    w = workloads.get(SWEEP_WORKLOAD)
    ctx = w.context(device)
    ctx.bk_init(schedule[0][0])
    d = api.DMRG(ctx)
    if world > 1:
        d.set_world(world, rank, allreduce)
    d.random_mps(SWEEP_SEED)
    t_begin = time.time()
    d.presolve()
    change, rows = False, []
    for D, nsweeps in schedule:
        for _ in range(nsweeps):
            for to_right in (False, True):
                d.timers(reset=True)
                t0 = time.time()
                e, dw = d.sweep(to_right, 1e-5, 0.0, D, change)
                change = True
                rows.append(dict(D=D, to_right=to_right, seconds=time.time() - t0, energy=e, max_discarded_weight=dw, **d.timers()))
    last = rows[-2:]
    hits, misses = d.plan_cache_stats()
    plans = {"reused_from_cache": hits, "built": misses, "built_on_helper_thread_during_update": d.plan_prefetched()}
    out = {"workload": f"{SWEEP_WORKLOAD}: N2/cc-pVDZ 14e/28o D2h (BASELINE config 2), seeded random MPS, schedule D=" + ",".join(str(D) for D, _ in schedule) +
                       " (one left+right sweep each), rtol 1e-5, no noise", "n_gpus": world,
           "D": schedule[-1][0], "seconds_per_sweep_at_D": last[0]["seconds"] + last[1]["seconds"], "total_seconds": time.time() - t_begin,
           "energy": min(r["energy"] for r in last), "max_discarded_weight": max(r["max_discarded_weight"] for r in last), "plans": plans,
           "half_sweeps": [{"D": r["D"], "dir": "->" if r["to_right"] else "<-", "s": round(r["seconds"], 3), "E": r["energy"], "plan_s": round(r["plan_s"], 3),
                            "solve_s": round(r["solve_s"], 3), "split_s": round(r["split_s"], 3), "update_s": round(r["update_s"], 3), "sigma_builds": r["n_matvec"]} for r in rows]}
    from oracle import refrun
    fcidump = os.path.join(ROOT, "oracle", "_ref", "N2.CCPVDZ.FCIDUMP")
    if with_reference and rank == 0 and refrun.available() and os.path.exists(fcidump):
        sched = ",".join(f"{D}:1e-14:{n}:0.0:1e-5" for D, n in schedule)
        env = dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count()), OPENBLAS_NUM_THREADS="1")
        t0 = time.time()
        res = subprocess.run([refrun.REF_DRIVER, "energies", "--fcidump", fcidump, "--group", "7", "--twoS", "0", "--N", "14", "--irrep", "0", "--reorder",
                              "--schedule", sched, "--seed", str(SWEEP_SEED)], capture_output=True, text=True, env=env, cwd="/tmp")
        walls = [float(ln.split("=")[1].split()[0]) for ln in res.stdout.splitlines() if "Elapsed wall time" in ln]
        mins = [float(ln.split("=")[1]) for ln in res.stdout.splitlines() if "Minimum energy           =" in ln]
        if len(walls) >= 2:
            out["reference"] = {"kind": "reference", "cores": os.cpu_count(), "seconds_per_sweep_at_D": walls[-2] + walls[-1], "total_seconds": time.time() - t0,
                                "half_sweep_seconds": [round(x, 2) for x in walls], "half_sweep_min_energies": mins,
                                "speedup_at_D": (walls[-2] + walls[-1]) / out["seconds_per_sweep_at_D"]}
    return out


def update_metric(torch, ctx, w, dims, old_set, stream, steps=3, with_reference=True):
    """Second half of the north-star hot path: the renormalized-operator update DMRG::updateMovingRight (DMRGoperators.cpp:243-574) at the
    bench shape — every operator of boundary site+1 (L, S0/S1, F0/F1, A/B/C/D, Q, X) from the hash-filled operators of boundary `site` and a
    synthetic site tensor, through b2_update_run_device; CUDA events on the context stream.  FLOPs = 2mnk per reference dgemm_ of
    TensorOperator::update & co (SURVEY 8(d) F_upd), computed analytically by the plan.  with_reference: the unmodified reference's
    updateMovingRight on the same inputs on the host cores (oracle/_ref `synthupdate`), timed, and a sample of the new operators compared."""
    from chemps2_b200 import api
    from chemps2_b200._lib import lib
    new_set = api.OpSet(ctx, w.site + 1, True)
    t0 = time.time()
    upd = api.Update(ctx, w.site, True, old_set, new_set)
    plan_s = time.time() - t0
    st = upd.stats()
    nt = lib.b2_tensor_t_size(ctx.h, w.site)
    from oracle import refrun
    t_dev = torch.from_numpy(api.hash_fill(nt, 7, key=refrun.op_key(4, 0, -1, -1), amp=0.1)).cuda()
    upd.run_device(t_dev.data_ptr())              # warm-up
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        upd.run_device(t_dev.data_ptr())
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out = {"what": "DMRG::updateMovingRight at the bench shape (all operators of boundary site+1)", "ms_per_update": ms,
           "gflop_per_update": st["flops_ref"] / 1e9, "tflops_fp64": st["flops_ref"] / (ms * 1e-3) / 1e12,
           "executed_tflops_fp64": st["flops_exec"] / (ms * 1e-3) / 1e12, "terms": st["terms"] + st["mix_terms"],
           "launches": st["launches"], "plan_build_s": plan_s}
    if with_reference and refrun.available():
        try:
            ref = refrun.run_reference_update(w, 7, moving_right=True, site=w.site, dims=dims)
            worst, checked = 0.0, 0
            for kind, i, j, size, rsum, rsq, rdot in ref["ops"][::max(1, len(ref["ops"]) // 48)]:
                if size == 0:
                    continue
                got = new_set.download(new_set.find(kind, i, j))
                h = api.hash_fill(size, 7 + 17, key=refrun.op_key(5, kind, i, j))
                worst = max(worst, abs(float(np.dot(got, h)) - rdot) / max(np.sqrt(rsq * size), 1e-300))
                checked += 1
            out["cpu_reference"] = {"kind": "reference", "cores": ref["threads"], "seconds_per_update": ref["update_s"],
                                    "tflops_fp64": st["flops_ref"] / ref["update_s"] / 1e12, "speedup": ref["update_s"] / (ms * 1e-3),
                                    "parity": {"operators_checked": checked, "max_rel_err_of_hash_projection": worst}}
        except Exception as e:
            out["cpu_reference"] = {"failed": str(e)}
    return out


def arena_doubles_of(*sets):
    from chemps2_b200._lib import lib
    return int(sum(lib.b2_opset_arena_size(x.h) for x in sets if x is not None))


def run_reference(args):
    """--impl reference: the unmodified reference's Heff::makeHeff on the host cores at the SAME config as the GPU arm (same D, sector
    table, operators, input).  One step = one complete sigma build; the builds actually run are reported in `steps` (a build at the
    default config takes minutes on the host, so the requested step count is capped — `steps_requested` keeps the request)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from chemps2_b200 import api
    from oracle import refrun
    if not refrun.available():
        emit({"impl": "reference", "unavailable": "oracle/_ref/ref_driver not built (needs /root/reference at build time)"})
        return
    w, ctx, dims = workload_and_dims(args, -1)
    left = api.OpSet(ctx, w.site, True)
    right = api.OpSet(ctx, w.site + 2, False)
    flops = api.Heff(ctx, w.site, left, right).stats()["flops_ref"]
    arena = arena_doubles_of(left, right)
    del left, right
    runs = []
    t0 = time.time()
    try:
        while len(runs) < max(1, args.steps) and (not runs or time.time() - t0 + 1.5 * (time.time() - t0) / len(runs) < args.ref_budget_s):
            base, _ = cpu_baseline(args, w, dims, flops, arena)
            runs.append(base)
    except Exception as e:   # e.g. not enough host memory for the reference's operator tables at this size: say so, never print a number for another size
        if not runs:
            emit({"impl": "reference", "unavailable": f"reference run at the bench config failed: {e}"})
            return
    best = max(runs, key=lambda r: r["value"])
    mean_s = float(np.mean([r["sample_seconds"] for r in runs]))
    line = {"metric": METRIC, "value": 1.0 / mean_s, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": len(runs),
            "steps_requested": args.steps, "warmup": 0, "warmup_requested": args.warmup, "ms_per_step": 1e3 * mean_s, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_of(args, w, flops),
            "cpu_baseline": {k: best[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": 1.0 / mean_s, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line["cpu_baseline"]["value"] = 1.0 / mean_s
    emit(line)


def config_of(args, w, flops):
    return {"workload": f"{w.name}: {w.N}e/{w.L}o point group {w.group} synthetic integrals, D={w.D}, sector model '{args.dist}', "
                        f"site pair ({w.site},{w.site + 1}), hash-filled renormalized operators",
            "gflop_per_sigma_build": flops / 1e9, "l2_policy": "inputs_exceed_l2 (operator arenas >> 126 MB)"}


def emit(line):
    """the ONE JSON line of the contract, on the process's real stdout (fd 1 is pointed at stderr while the bench runs, so that
    banners of libraries — NCCL prints its version to stdout — cannot end up in front of it)"""
    os.write(REAL_STDOUT, (json.dumps(line) + "\n").encode())


REAL_STDOUT = 1


def main():
    global REAL_STDOUT
    args = parse()
    sys.stdout.flush()
    REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)
    import torch
    import torch.distributed as dist

    from chemps2_b200 import api
    from chemps2_b200._lib import check, lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.current_stream()

    w, ctx, dims = workload_and_dims(args, local)
    ctx.set_stream(stream.cuda_stream)
    left = api.OpSet(ctx, w.site, True)
    right = api.OpSet(ctx, w.site + 2, False)
    left.fill_hash(7, 1.0)
    right.fill_hash(7, 1.0)
    t0 = time.time()
    heff = api.Heff(ctx, w.site, left, right, world, rank)
    plan_s = time.time() - t0
    st = heff.stats()
    n = heff.n
    # a different S every step (as in the Davidson loop), generated up front, pinned on the host for the e2e leg
    nvec = args.steps + args.warmup
    host_in = [torch.from_numpy(api.hash_fill(n, 100 + i)).pin_memory() for i in range(min(nvec, 4))]
    host_out = torch.empty(n, dtype=torch.float64).pin_memory()
    dev_in = [h.cuda(non_blocking=True) for h in host_in]
    dev_out = torch.empty(n, dtype=torch.float64, device="cuda")

    def step_device(i):
        heff.apply_device(dev_in[i % len(dev_in)].data_ptr(), dev_out.data_ptr())
        if world > 1:
            dist.all_reduce(dev_out)

    def step_e2e(i):
        if world == 1:
            hin = host_in[i % len(host_in)]
            check(lib.b2_heff_apply(heff.h, C.cast(hin.data_ptr(), C.POINTER(C.c_double)), C.cast(host_out.data_ptr(), C.POINTER(C.c_double))))
        else:
            d = dev_in[i % len(dev_in)]
            d.copy_(host_in[i % len(host_in)], non_blocking=True)
            heff.apply_device(d.data_ptr(), dev_out.data_ptr())
            dist.all_reduce(dev_out)
            host_out.copy_(dev_out, non_blocking=True)
            stream.synchronize()

    def timed(fn):
        for i in range(args.warmup):
            fn(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(args.steps):
            fn(args.warmup + i)
        e1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local)
    sampler.start()
    ms_dev = timed(step_device)
    # kernel-only time of the sigma kernels (CUDA events around the launches on the same stream), averaged live
    kern = []
    for i in range(args.steps):
        heff.apply_device(dev_in[i % len(dev_in)].data_ptr(), dev_out.data_ptr())
        kern.append(heff.kernel_seconds())
    ms_e2e = timed(step_e2e)
    sampler.stop_flag = True
    sampler.join()
    # a checksum that is comparable across N: the same input vector (index 0) through the device path (+ all-reduce)
    step_device(0)
    torch.cuda.synchronize()
    sigma_norm = float(torch.linalg.vector_norm(dev_out).item())
    sigma_probe = [float(x) for x in dev_out[:: max(1, n // 7)][:8].cpu()]

    sweep_multi = None
    if world > 1 and not args.no_sweep:   # the whole-sweep metric at N GPUs: every rank takes part (sharded sigma + updates), rank 0 reports
        try:
            del heff, left, right
            torch.cuda.empty_cache()
            ar = api.AllReduce()
            sched = [(int(x), 1) for x in args.sweep_D.split(",")] if args.sweep_D else None
            sweep_multi = sweep_metric(local, False, schedule=sched, world=world, rank=rank, allreduce=ar)
        except Exception as e:
            sweep_multi = {"failed": str(e)}
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    value = args.steps / (ms_dev * 1e-3)
    e2e = args.steps / (ms_e2e * 1e-3)
    # roofline: FP64 tensor (DMMA) pipe
    peak = C.c_double()
    check(lib.b2_probe_fp64(ctx.h, 1, C.byref(peak)))
    kavg = float(np.mean(kern))
    # algorithmic FLOPs (2mnk per reference dgemm_; at N > 1 the rank's share is taken as 1/N of the plan) and the
    # FLOPs the kernels really execute (fewer: shared stage-1 products, cheaper association order) — both against the same peak
    achieved = st["flops_ref"] / world / kavg / 1e12
    executed = st["flops_exec"] / kavg / 1e12 if world == 1 else None
    # DRAM bytes of the k_tiles launches of ONE sigma build of this workload from this round's `ncu --set full` capture (profiles/)
    traffic, traffic_src = None, None
    try:
        if world == 1 and args.workload == "synth40" and args.D is None and args.dist == "gauss" and not args.work_budget and not args.chunk_k:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
            traffic, traffic_src = tj["k_tiles_dram_bytes_per_sigma_build"], tj.get("source")
    except Exception:
        traffic = None
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak.value, "unit": "TFLOP/s", "frac": achieved / peak.value,
                "traffic": traffic, "traffic_source": traffic_src,
                "achieved_is": "ALGORITHMIC FLOPs of the reference's dgemm_ calls (SURVEY 8(d)) / measured kernel time",
                "executed_tflops": executed, "executed_frac": (executed / peak.value) if executed else None,
                "executed_is": "FLOPs the kernels really execute (useful part of the issued DMMAs) / kernel time: the hardware-utilisation figure",
                "kernel": "k_tiles (grouped FP64 DMMA contraction: stage-1 + stage-2 launches of one sigma build)",
                "kernel_ms_per_sigma_build": kavg * 1e3,
                "peak_source": "measured in this run: register-resident mma.sync.m8n8k4.f64 loop (b2_probe_fp64); MEASURED_PEAKS.json holds no FP64 figure"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_of(args, w, st["flops_ref"]),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(8 * n), "d2h_bytes_per_step": int(8 * n)},
            "gpu_launches": int(st["launches"] * args.steps), "roofline": roofline, "clocks": sampler.result(),
            "tflops_fp64": st["flops_ref"] / (ms_dev / args.steps * 1e-3) / 1e12,
            "plan": {"terms": st["terms"], "waves": st["waves"], "ctas": st["tiles"], "plan_build_s": plan_s, "veclength": int(n),
                     "exec_over_ref_flops": st["flops_exec"] / st["flops_ref"], "sigma_norm": sigma_norm, "sigma_probe": sigma_probe}}
    if sweep_multi is not None:
        line["sweep"] = sweep_multi
    if world == 1 and not args.no_cpu_baseline:
        try:
            # the GPU result for the reference's input vector (seed 7), computed BEFORE the host is loaded with the reference run
            out = heff.apply(api.hash_fill(n, 7))
            base, ref = cpu_baseline(args, w, dims, st["flops_ref"], arena_doubles_of(left, right))
            line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
            # full-size parity: this very plan on these very operators vs the unmodified reference's makeHeff output
            scale = float(np.abs(ref["vec_out"]).max())
            line["parity_vs_reference"] = {"max_rel_err": float(np.abs(out - ref["vec_out"]).max() / scale), "veclength": int(n), "D": w.D,
                                           "what": "max|sigma_gpu - sigma_ref| / max|sigma_ref| at the bench size, same operators and input"}
            del out, ref
        except Exception as e:   # the baseline must not take the bench line down
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"failed: {e}"}
    if world == 1 and not args.no_update:
        try:
            del heff                                     # frees the sigma plan (workspace + work lists) before the update plan is built
            um = update_metric(torch, ctx, w, dims, left, stream, with_reference=not args.no_cpu_baseline)
            um["frac_of_fp64_peak"] = um["tflops_fp64"] / peak.value
            um["executed_frac_of_fp64_peak"] = um["executed_tflops_fp64"] / peak.value
            line["operator_update"] = um
        except Exception as e:
            line["operator_update"] = {"failed": str(e)}
    if world == 1 and not args.no_sweep:
        try:
            left = right = None
            torch.cuda.empty_cache()
            sched = [(int(x), 1) for x in args.sweep_D.split(",")] if args.sweep_D else None
            line["sweep"] = sweep_metric(local, args.sweep_ref, schedule=sched)
        except Exception as e:   # the secondary metric must not take the bench line down
            line["sweep"] = {"failed": str(e)}
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
