#!/usr/bin/env python
"""Benchmark of the adFVM residual + discrete-adjoint hot path (BASELINE.json metric:
primal + adjoint Mcell-updates/s per RK stage; HBM GB/s as % of peak).

A "step" = one primal time step (3 SSPRK stages) + one adjoint time step (forward recompute + reverse sweep,
3 stages) of the synthetic periodic hex box, i.e. 2*3*C cell-updates. Default workload: 368^3 = 49.8 M
cells per GPU (BASELINE.json config 5; 256^3 if the host cannot hold the input arrays of all ranks, said in config).
  value        device-resident stepping (inputs already in HBM), CUDA events on the library's stream
  e2e          the same work through PrimalFunction/AdjointFunction.__call__ with pinned HOST buffers:
               state + adjoint uploaded and results downloaded every call
  roofline     dominant kernel, algorithmic bytes / CUDA-event kernel time per pass / measured peak;
               stage_model = whole-stage byte model of BASELINE.md over the step time
  cpu_baseline the oracle port timed on this host on a bounded sample (112^3)

`--impl reference` times the CPU oracle port only (rank 0), same metric and config.
Multi-GPU: one rank per GPU (torchrun), weak scaling: every rank owns an n^3 block of a (n*px, n*py, n*pz)
periodic box, halo over NCCL inside the library, overlapped with the tiles that do not depend on it.
"""
import argparse
import json
import os
import resource
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "primal+adjoint Mcell-updates/s per RK stage"
UNIT = "Mcell-updates/s"


def algorithmic_bytes(s):
    """BASELINE.md §3 unique-touch model, periodic hex box (F = Fi = 3C, G << C), s = sizeof(scalar)."""
    per_stage_primal = 149.0 + 1.0 / 3        # scalars per cell per stage
    per_stage_reverse = 196.0 + 2.0 / 3
    B_p = per_stage_primal * s + 96
    B_a = (per_stage_primal + per_stage_reverse) * s + 192
    # dominant kernels, rows of BASELINE.md §3 they cover (per cell per launch):
    #  flux_tile   = flux (20 + 51 + 3 + 6 scalars, 2F ints) + RK update (23 1/3) + primitive of next stage (10)
    #  flux_grad_tile = flux_grad row: 5 + 40 + (51 + 3) + 40 scalars, 2F ints
    #  grad_adj_update = gradCell VJP + primitive VJP + RK combination fused: Qb 5, Gb 15, cell-face metrics 24, W 5,
    #                    later-stage adjoints 5/5/15 (+ source-gradient RMW 10 on the last reverse stage) = 11 2/3 avg, out 5; 6 ints
    k = {"flux_tile": (80 + 23 + 1.0 / 3 + 10) * s + 24, "flux_grad_tile": (5 + 40 + 54 + 40) * s + 24,
         "grad_cell": (5 + 15 + 1 + 15) * s + 72, "grad_adj_update": (5 + 15 + 24 + 5 + 11 + 2.0 / 3 + 5) * s + 24}
    return B_p, B_a, k


# full passes over the mesh per (primal + adjoint) step: a kernel may be launched in two parts (tiles / cells that
# do not depend on halo data first), so rates are computed from its total time per step, not per launch
PASSES = {"flux_tile": 5, "flux_grad_tile": 3, "grad_cell": 6, "grad_adj_update": 3}


class ClockSampler:
    def __init__(self, device):
        self.rows, self.stop = [], False
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def finish(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_rate(n, steps=1, dtype=np.float64):
    """primal + adjoint step of the oracle port on an n^3 periodic box; returns (Mcell-updates/s/stage combined,
    primal, adjoint, cores)."""
    import torch
    from adfvm_b200 import cases
    from oracle import adfvm_oracle as O
    case = cases.periodic_box(n, dtype)
    C = case.mesh.nInternalCells
    adj = [np.ones_like(s) * w for s, w in zip(case.state, (1.0, 1e-2, 1e-5))]
    tp = ta = 0.0
    state = case.state
    for _ in range(steps):
        t0 = time.perf_counter(); out = O.primal(case.spec, case.inputs(state)); t1 = time.perf_counter()
        O.primal_grad(case.spec, case.adjoint_inputs(state, adj)); t2 = time.perf_counter()
        tp += t1 - t0; ta += t2 - t1
        state = [np.ascontiguousarray(o, dtype) for o in out[:3]]
    return (2 * 3 * C * steps / (tp + ta) / 1e6, 3 * C * steps / tp / 1e6, 3 * C * steps / ta / 1e6,
            torch.get_num_threads(), tp + ta)


HOST_BYTES_PER_CELL = 1100          # numpy inputs (mesh metrics, connectivity, state) + pinned buffers + tile plan, measured at 368^3


def host_memory_available():
    """bytes of host memory this job may still use (cgroup limit aware)"""
    avail = None
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        pass
    try:
        mx = open("/sys/fs/cgroup/memory.max").read().strip()
        cur = int(open("/sys/fs/cgroup/memory.current").read().strip())
        if mx != "max":
            lim = int(mx) - cur
            avail = lim if avail is None else min(avail, lim)
    except Exception:
        pass
    return avail


def default_size(world):
    """368^3 per GPU (BASELINE.json config 5) unless the host cannot hold the inputs of all ranks; then 256^3, said in config"""
    n, note = 368, None
    avail = host_memory_available()
    need = HOST_BYTES_PER_CELL * 368 ** 3 * world * 1.25
    if avail is not None and avail < need:
        n, note = 256, "host memory %.0f GB < %.0f GB needed for %d ranks at 368^3: 256^3 per GPU instead" % (avail / 1e9, need / 1e9, world)
    return n, note


def run_reference(args, rank):
    if rank != 0:
        return
    n = args.cpu_n
    # bounded sample: warm-up + K steps of the 'n^3' box
    for _ in range(min(args.warmup, 1)):
        cpu_oracle_rate(n, 1)
    v, vp, va, cores, secs = cpu_oracle_rate(n, max(1, min(args.steps, 3)))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": secs * 1e3 / max(1, min(args.steps, 3)), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "periodic_hex_box_n%d" % args.n, "cells_per_gpu": args.n ** 3, "rk_stages": 3,
                       "step": "1 primal step + 1 adjoint step (incl. forward recompute)",
                       "sample": "each step = one primal + adjoint step of a %d^3 box of the same workload on the host cores" % n},
            "primal": vp, "adjoint": va,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "oracle port (torch CPU fp64), %d^3 periodic box, %d primal+adjoint steps" % (n, max(1, min(args.steps, 3)))},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", dest="n", type=int, default=int(os.environ.get("ADFVM_BENCH_N", "0")),
                    help="cells per side per GPU (default: 368 = the ~50M cells/GPU of BASELINE.json config 5, or 256 when the "
                         "host memory cannot hold the input arrays of all ranks)")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--cpu-n", type=int, default=112, help="box size of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    size_note = None
    if args.n <= 0 and (args.impl == "reference" or world == 1):
        args.n, size_note = default_size(max(world, args.gpus if args.impl == "reference" else 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from adfvm_b200 import cases, function, _lib
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    lib = _lib.default_lib()
    assert lib.is_cuda
    dtype = np.float64 if args.dtype == "f64" else np.float32
    s = np.dtype(dtype).itemsize
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        if args.n <= 0:                       # rank 0 decides, everybody follows
            box = [default_size(world) if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            args.n, size_note = box[0]
    W = max(args.warmup, 3)
    K = args.steps

    # ---- workload
    if world > 1:
        from adfvm_b200 import decompose
        case = decompose.periodic_box_rank(args.n, rank, world, dtype)
    else:
        case = cases.periodic_box(args.n, dtype)
    C = case.mesh.nInternalCells
    torch.cuda.set_stream(torch.cuda.Stream())            # an explicit stream (events below are recorded on it); the legacy
    stream = torch.cuda.current_stream().cuda_stream      # default stream would rule out the CUDA graphs of whole steps
    f = function.PrimalFunction(case.spec, dtype, device=local, stream=stream)
    fa = f.grad()
    if world > 1:
        from adfvm_b200 import decompose
        decompose.attach_comm(f, rank, world)
    rng = np.random.RandomState(3 + rank)
    adj0 = [np.ascontiguousarray(rng.randn(*a.shape) * w, dtype) for a, w in zip(case.state, (1.0, 1e-2, 1e-5))]
    # first call uploads the static data (mesh, BCs, source) and the state; primes the adjoint buffers
    f(*case.inputs(), replace_reusable=True, return_reusable=False)
    fa(*case.adjoint_inputs(case.state, adj0), return_static=False)
    f(*case.inputs(), replace_reusable=True, return_reusable=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def resident_step():
        f.step_resident(case.dt)
        fa.step_resident(case.dt, 1.0, chain=True)

    # ---- device-resident timing (value). Working set per step >> L2 (126 MB): no explicit flush needed.
    for _ in range(W):
        resident_step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = f.launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * K + 1)]
    ev[0].record()
    for k in range(K):
        f.step_resident(case.dt); ev[2 * k + 1].record()
        fa.step_resident(case.dt, 1.0, chain=True); ev[2 * k + 2].record()
    barrier()
    launches = f.launches - l0
    total_ms = ev[0].elapsed_time(ev[-1])
    tp_ms = sum(ev[2 * k].elapsed_time(ev[2 * k + 1]) for k in range(K))
    ta_ms = sum(ev[2 * k + 1].elapsed_time(ev[2 * k + 2]) for k in range(K))
    clocks = sampler.finish() if sampler else None
    t = torch.tensor([total_ms, tp_ms, ta_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, tp_ms, ta_ms = t.tolist()

    # ---- per-kernel times (separate pass, CUDA events around every launch on the launching stream)
    f.kernel_timing(True)
    for _ in range(max(2, min(K, 5))):
        resident_step()
    rep = f.kernel_report()
    f.kernel_timing(False)
    nrep = max(2, min(K, 5))

    # ---- end to end through the public call with pinned host buffers
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    hstate = [pin(a) for a in case.state]
    hadj = [pin(a) for a in adj0]
    static = case.inputs(hstate)
    e2e_K = max(2, min(K, 5))
    for _ in range(2):
        out = f(*case.inputs(hstate), replace_reusable=True, return_reusable=True)
        g = fa(*case.adjoint_inputs(hstate, hadj), return_static=True, zero_static=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_K):
        out = f(*case.inputs(hstate), replace_reusable=True, return_reusable=True)
        g = fa(*case.adjoint_inputs(hstate, hadj), return_static=True, zero_static=True)
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = te.item()
    h2d = (5 * C * s + s) + (10 * C * s + 4 * s)                  # primal: state + dt; adjoint: state + adjoint + 4 scalars
    d2h = (5 * C * s + 2 * s) + (10 * C * s)                        # primal: state+dtc+obj; adjoint: adjoint fields + source gradients

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    B_p, B_a, kb = algorithmic_bytes(s)
    cells_total = C * world
    value = 2 * 3 * cells_total * K / (total_ms * 1e-3) / 1e6
    vp = 3 * cells_total * K / (tp_ms * 1e-3) / 1e6
    va = 3 * cells_total * K / (ta_ms * 1e-3) / 1e6
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    kernels = {}
    for name, (cnt, ms) in rep.items():
        per = ms / cnt
        e = {"launches_per_step": cnt / nrep, "ms_per_launch": per, "share": None}
        if name in kb:
            e["ms_per_pass"] = ms / nrep / PASSES[name]
            e["algorithmic_GBs"] = kb[name] * C / (e["ms_per_pass"] * 1e-3) / 1e9
            e["frac"] = e["algorithmic_GBs"] / peak
        kernels[name] = e
    tot = sum(ms for _, ms in rep.values())
    for name, (cnt, ms) in rep.items():
        kernels[name]["share"] = ms / tot
    dom = max((k for k in kernels if k in kb), key=lambda k: rep[k][1])
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = "%s_%s_n128" % (dom, args.dtype)      # ncu --set full capture at 128^3; DRAM bytes scale with the cell count
        traffic = tr.get(key)
        if traffic is not None:
            traffic = int(traffic * (C / 128.0 ** 3))
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["algorithmic_GBs"], "peak": peak, "unit": "GB/s",
            "frac": kernels[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
            "algorithmic_bytes_per_cell": kb[dom],
            "stage_model": {"primal_bytes_per_cell_stage": B_p, "adjoint_bytes_per_cell_stage": B_a,
                            "primal_GBs": B_p * 3 * C / (tp_ms / K * 1e-3) / 1e9, "adjoint_GBs": B_a * 3 * C / (ta_ms / K * 1e-3) / 1e9,
                            "primal_frac": B_p * 3 * C / (tp_ms / K * 1e-3) / 1e9 / peak,
                            "adjoint_frac": B_a * 3 * C / (ta_ms / K * 1e-3) / 1e9 / peak}}
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        v, cvp, cva, cores, secs = cpu_oracle_rate(args.cpu_n, 1)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "primal": cvp, "adjoint": cva,
               "sample": "oracle port (torch CPU fp64), %d^3 periodic box, 1 primal + 1 adjoint step, %.1f s" % (args.cpu_n, secs)}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": "periodic_hex_box_n%d" % args.n, "cells_per_gpu": C, "rk_stages": 3,
                       "step": "1 primal step + 1 adjoint step (incl. forward recompute)",
                       "l2": "working set per step (%.0f MB) exceeds the 126 MB L2; no explicit flush" % (f.device_bytes / 1e6),
                       "device_bytes": f.device_bytes, "host_rss_gb": round(resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e6, 1),
                       "size_note": size_note},
            "primal": vp, "adjoint": va, "primal_ms": tp_ms / K, "adjoint_ms": ta_ms / K,
            "e2e": {"value": 2 * 3 * cells_total * e2e_K / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3 / e2e_K,
                    "note": "PrimalFunction/AdjointFunction.__call__, pinned host buffers, full state+adjoint up and down every call"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "kernels": kernels, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
