#!/usr/bin/env python
"""Benchmark of the adFVM residual + discrete-adjoint hot path (BASELINE.json metric:
primal + adjoint Mcell-updates/s per RK stage; HBM GB/s as % of peak).

A "step" = one primal time step (3 SSPRK stages) + one adjoint time step (forward recompute + reverse sweep,
3 stages) of the synthetic periodic hex box, i.e. 2*3*C cell-updates. Default workload: 368^3 = 49.8 M
cells per GPU (BASELINE.json config 5; 256^3 if the host cannot hold the input arrays of all ranks, said in config).
  parity       BEFORE anything is timed: the CUDA path against the reference's own compiled step functions
               (oracle/_ref, kind "reference"; the oracle port if that module did not travel) at 48^3 on one GPU, and on
               N > 1 GPUs every rank's 48^3 block of the decomposed box against the single-rank run of the undecomposed
               mesh (the reference's criterion, tests/test_parallel.py:63-81) with early_tiles > 0 (the overlapped path)
  value        device-resident stepping (inputs already in HBM), CUDA events on the library's stream
  e2e          the same work through PrimalFunction/AdjointFunction.__call__ with pinned HOST buffers, driven the way the
               reference's Adjoint.run drives a checkpoint block (apps/adjoint.py:214-291): forward-mode primal calls that
               return every state, then the block backwards through primal_grad with state + adjoint uploaded and the
               adjoint downloaded every call, the source-term gradient at the end of the block (sampleInterval = block)
  roofline     dominant kernel, algorithmic bytes / CUDA-event kernel time per pass / measured peak;
               stage_model = whole-stage byte model of BASELINE.md over the step time
  cpu_baseline the reference's compiled CPU path on this host's cores on a bounded sample (see run_reference)
  fp32         the same device-resident measurement in fp32 (sub-record; skipped with --no-fp32)
  strong       N > 1: a fixed 368^3 mesh split over the N ranks (sub-record)

`--impl reference` times the reference's own CPU implementation (rank 0 only): one single-rank process per host core,
each stepping its own periodic box through the reference's generated `primal` / `primal_grad` (no MPI in this image: the
processes do not exchange halos, which favours the reference). Same metric, unit and config.
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
T0 = time.time()


def log(*a):
    print("[bench %6.1fs]" % (time.time() - T0), *a, file=sys.stderr, flush=True)


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
    # what the FUSED kernels strictly have to move per cell (no residual / LHS round trip, next-stage primitives written
    # once): flux_tile reads Q,G 20, metrics 3 faces x 16 + V, W0/S/Wx 13 1/3, writes W 5 + Q 5 (+ 24 B of packed words)
    strict = {"flux_tile": (20 + 48 + 1 + 13 + 1.0 / 3 + 10) * s + 12, "flux_grad_tile": (20 + 5 + 48 + 1 + 20) * s + 12}
    return B_p, B_a, k, strict


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


# ------------------------------------------------------------------------------------------ CPU reference arm
def cpu_port_rate(n, steps=1, dtype=np.float64):
    """fallback when oracle/_ref did not travel: primal + adjoint step of the oracle port on an n^3 periodic box"""
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
    return {"value": 2 * 3 * C * steps / (tp + ta) / 1e6, "primal": 3 * C * steps / tp / 1e6, "adjoint": 3 * C * steps / ta / 1e6,
            "cores": torch.get_num_threads(), "kind": "port", "seconds": tp + ta, "unit": UNIT,
            "sample": "oracle port (torch CPU fp64), %d^3 periodic box, %d primal + adjoint step(s)" % (n, steps)}


def _ref_worker(idx, core, n, warmup, steps, barrier, q):
    """one single-rank process of the reference's compiled step functions on its own n^3 periodic box, pinned to `core`"""
    try:
        try:
            os.sched_setaffinity(0, {core})
        except Exception:
            pass
        from adfvm_b200 import cases
        from oracle import refgraph
        g = refgraph.RefGraph("box_cyclic")
        case = cases.periodic_box(n)
        devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(devnull, 1)      # "Initializing C++ interface"
        g.initialize(case.mesh)
        os.dup2(saved, 1)
        adj = [np.ones_like(s) * w for s, w in zip(case.state, (1.0, 1e-2, 1e-5))]
        state = case.state
        tp = ta = 0.0
        barrier.wait()
        for it in range(warmup + steps):
            if it == warmup:
                barrier.wait()
                tp = ta = 0.0
                t_start = time.perf_counter()
            t0 = time.perf_counter()
            out = g.primal(*case.inputs(state), replace_reusable=True, return_reusable=True)
            t1 = time.perf_counter()
            g.primal_grad(*case.adjoint_inputs(state, adj), return_static=True, zero_static=True)
            t2 = time.perf_counter()
            tp += t1 - t0; ta += t2 - t1
            state = [np.ascontiguousarray(o) for o in out[:3]]
        q.put((idx, tp, ta, time.perf_counter() - t_start, g.variant))
    except Exception as e:           # noqa: BLE001
        try:
            barrier.abort()
        except Exception:
            pass
        q.put((idx, None, None, None, repr(e)))


def cpu_reference_rate(n, warmup, steps, max_procs=None):
    """The reference's own compiled primal / primal_grad (oracle/_ref) on every host core this process may use: P independent
    single-rank processes, each on an n^3 periodic box, host numpy arrays in and out (the reference's calling convention).
    Aggregate rate = sum over processes of their own rates during the common timed region."""
    import multiprocessing as mp
    cores = sorted(os.sched_getaffinity(0))
    per_proc = 5200 * n ** 3 + 400e6                     # bytes: the reference keeps ~5 kB per cell alive (measured at 48^3-64^3)
    avail = host_memory_available()
    P = len(cores)
    if avail:
        P = max(1, min(P, int(0.6 * avail / per_proc)))
    if max_procs:
        P = min(P, max_procs)
    ctx = mp.get_context("spawn")
    barrier, q = ctx.Barrier(P), ctx.Queue()
    procs = [ctx.Process(target=_ref_worker, args=(i, cores[i], n, warmup, steps, barrier, q)) for i in range(P)]
    for p in procs:
        p.start()
    res = [q.get() for _ in procs]
    for p in procs:
        p.join()
    bad = [r for r in res if r[1] is None]
    if bad:
        raise RuntimeError("reference worker failed: %s" % bad[0][4])
    C = n ** 3
    vp = sum(3 * C * steps / r[1] for r in res) / 1e6
    va = sum(3 * C * steps / r[2] for r in res) / 1e6
    v = sum(6 * C * steps / (r[1] + r[2]) for r in res) / 1e6
    wall = max(r[3] for r in res)
    return {"value": v, "primal": vp, "adjoint": va, "cores": P, "kind": "reference", "seconds": wall, "unit": UNIT,
            "per_core": {"primal": vp / P, "adjoint": va / P},
            "sample": "the reference's own generated primal/primal_grad (oracle/_ref, %s), %d independent single-rank processes "
                      "pinned to %d of %d usable cores, each %d step(s) of a %d^3 periodic box, fp64, %.1f s" %
                      (res[0][4], P, P, len(cores), steps, n, wall)}


def cpu_baseline(n, warmup, steps):
    from oracle import refgraph
    if refgraph.available("box_cyclic"):
        return cpu_reference_rate(n, warmup, steps)
    return cpu_port_rate(min(n, 64), max(1, min(steps, 2)))


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
    K, W = max(1, args.steps), max(0, args.warmup)
    # bounded sample: each step = one primal + adjoint step of an n^3 box per process (~2 s at 64^3); cap the run at ~3 min
    K_run, W_run = min(K, 20), min(W, 3)
    r = cpu_baseline(n, W_run, K_run)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": K,
            "warmup": W, "ms_per_step": r["seconds"] * 1e3 / K_run, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "periodic_hex_box_n%d" % args.n, "cells_per_gpu": args.n ** 3, "rk_stages": 3,
                       "step": "1 primal step + 1 adjoint step (incl. forward recompute)",
                       "sample": "each step = one primal + adjoint step of a %d^3 box of the same workload per host core; %d of the %d "
                                 "requested steps were run (bounded sample)" % (n, K_run, K)},
            "primal": r["primal"], "adjoint": r["adjoint"], "cpu_baseline": r,
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ parity before timing
def _relerr(x, y):
    return float(np.abs(x - y).max() / max(np.abs(y).max(), 1e-300))


def _group_err(a, b, sc):
    num = max(float(np.abs(x - y).max()) * s for x, y, s in zip(a, b, sc))
    den = max(float(np.abs(y).max()) * s for y, s in zip(b, sc))
    return num / max(den, 1e-300)


def parity_check(rank, world, local, stream, n=48):
    """fp64. One GPU: CUDA vs the reference's compiled functions (or the oracle port) at n^3: 2 primal steps + 1 adjoint step.
    N GPUs: every rank's n^3 block vs the single-rank run of the undecomposed (n*px, n*py, n*pz) box on its own GPU."""
    import torch
    import torch.distributed as dist
    from adfvm_b200 import cases, function, decompose
    rng = np.random.RandomState(7)
    out = {"n": n}
    if world == 1:
        case = cases.periodic_box(n)
        f = function.PrimalFunction(case.spec, np.float64, device=local, stream=stream)
        o1 = f(*case.inputs(), replace_reusable=True)
        o2 = f(*case.inputs(list(o1[:3])), replace_reusable=True)
        adj = [np.ascontiguousarray(rng.randn(*s.shape) * w) for s, w in zip(case.state, (1.0, 1e-2, 1e-5))]
        g = f.grad()(*case.adjoint_inputs(case.state, adj))
        from oracle import refgraph
        if refgraph.available("box_cyclic"):
            R = refgraph.RefGraph("box_cyclic")
            devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(devnull, 1)
            R.initialize(case.mesh)
            os.dup2(saved, 1)
            r1 = R.primal(*case.inputs(), replace_reusable=True)
            r2 = R.primal(*case.inputs([np.ascontiguousarray(x) for x in r1[:3]]), replace_reusable=True)
            gr = R.primal_grad(*case.adjoint_inputs(case.state, adj))
            out["against"] = "reference (oracle/_ref compiled primal/primal_grad, %s)" % R.variant
        else:
            from oracle import adfvm_oracle as O
            r1 = O.primal(case.spec, case.inputs())
            r2 = O.primal(case.spec, case.inputs([np.ascontiguousarray(x) for x in r1[:3]]))
            gr = O.primal_grad(case.spec, case.adjoint_inputs(case.state, adj))
            out["against"] = "oracle port (oracle/_ref did not travel)"
        sc = [float(np.abs(s).max()) for s in case.state]
        errs = [_relerr(a, b) for a, b in zip(o1, r1)] + [_relerr(a, b) for a, b in zip(o2, r2)]
        errs += [_group_err(g[:3], gr[:3], sc), _group_err(g[3:6], gr[3:6], sc)]
        out.update(maxerr=max(errs), early_tiles=f.tile_rounds()[2], tiles=f.tile_stats()[2])
        f.c.close()
        return out
    N = (n, n, n)
    g = decompose.global_box(N, world)
    fs = function.PrimalFunction(g.spec, np.float64, device=local, stream=stream)
    o1 = fs(*g.inputs(), replace_reusable=True)
    o2 = fs(*g.inputs(list(o1[:3])), replace_reusable=True)
    adj = [np.ascontiguousarray(rng.randn(*s.shape) * w) for s, w in zip(g.state, (1.0, 1e-2, 1e-5))]
    grad = fs.grad()(*g.adjoint_inputs(g.state, adj))
    fs.c.close()
    case = decompose.periodic_box_rank(N, rank, world)
    f = function.PrimalFunction(case.spec, np.float64, device=local, stream=stream)
    decompose.attach_comm(f, rank, world)
    ids = decompose.global_cell_ids(N, rank, world)
    p1 = f(*case.inputs(), replace_reusable=True)
    p2 = f(*case.inputs(list(p1[:3])), replace_reusable=True)
    gr = f.grad()(*case.adjoint_inputs(case.state, [np.ascontiguousarray(x[ids]) for x in adj]))
    sc = [float(np.abs(s).max()) for s in g.state]
    errs = [_relerr(x, y[ids]) for x, y in zip(p1[:3], o1[:3])] + [_relerr(x, y[ids]) for x, y in zip(p2[:3], o2[:3])]
    errs.append(_relerr(p1[4], o1[4]))
    errs += [_group_err(gr[:3], [y[ids] for y in grad[:3]], sc), _group_err(gr[3:6], [y[ids] for y in grad[3:6]], sc)]
    early, tiles = f.tile_rounds()[2], f.tile_stats()[2]
    f.c.close()
    t = torch.tensor([max(errs), -float(early), float(tiles)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out.update(maxerr=float(t[0]), early_tiles=int(-t[1]), tiles=int(t[2]),
               against="single-rank run of the undecomposed %dx%dx%d-block box on each rank's own GPU" % tuple(decompose.factor3(world)))
    return out


# ------------------------------------------------------------------------------------------ GPU measurement
def measure(case, dtype, rank, world, local, stream, K, W, want_kernels=True, want_e2e=True, seed=3):
    """device-resident value (+ per-kernel times, + e2e through the public call) of one workload; times are max over ranks"""
    import torch
    import torch.distributed as dist
    from adfvm_b200 import function, decompose
    s = np.dtype(dtype).itemsize
    C = case.mesh.nInternalCells
    t_setup = time.time()
    f = function.PrimalFunction(case.spec, dtype, device=local, stream=stream)
    fa = f.grad()
    if world > 1:
        decompose.attach_comm(f, rank, world)
    rng = np.random.RandomState(seed + rank)
    adj0 = [np.ascontiguousarray(rng.randn(*a.shape) * w, dtype) for a, w in zip(case.state, (1.0, 1e-2, 1e-5))]
    # first call uploads the static data (mesh, BCs, source) and the state; primes the adjoint buffers
    f(*case.inputs(), replace_reusable=True, return_reusable=False)
    t_static = time.time() - t_setup
    fa(*case.adjoint_inputs(case.state, adj0), return_static=False)
    f(*case.inputs(), replace_reusable=True, return_reusable=False)
    log("rank %d: context ready (%d cells, %s): static upload + tile plan %.1f s" % (rank, C, np.dtype(dtype).name, t_static))

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
    res = {"C": C, "total_ms": total_ms, "tp_ms": tp_ms, "ta_ms": ta_ms, "launches": launches, "clocks": clocks,
           "device_bytes": f.device_bytes, "static_s": t_static, "early_tiles": f.tile_rounds()[2], "tiles": f.tile_stats()[2]}

    # ---- per-kernel times (separate pass, CUDA events around every launch on the launching stream)
    if want_kernels:
        f.kernel_timing(True)
        nrep = max(2, min(K, 5))
        for _ in range(nrep):
            resident_step()
        res["kernel_report"] = f.kernel_report()
        res["kernel_reps"] = nrep
        f.kernel_timing(False)

    # ---- end to end through the public call with pinned host buffers: a checkpoint block as Adjoint.run drives it
    if want_e2e:
        B = max(2, min(K, 3))
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
        hstate = [pin(a) for a in case.state]
        hadj0 = [pin(a) for a in adj0]

        def block():
            states, state = [hstate], hstate
            for k in range(B):                                   # Solver.run(mode='forward'), adFVM/solver.py:296-382
                o = f(*case.inputs(state), replace_reusable=(k == 0), return_reusable=True)
                state = list(o[:3]); states.append(state)
            adj = hadj0
            for k in range(B):                                   # apps/adjoint.py:250-291
                last = k == B - 1
                g = fa(*case.adjoint_inputs(states[B - 1 - k], adj), return_static=last, zero_static=last)
                adj = list(g[:3])
            return g
        block()
        barrier()
        t0 = time.perf_counter()
        block()
        barrier()
        e2e_s = time.perf_counter() - t0
        te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        res["e2e_s_per_step"] = te.item() / B
        res["e2e_block"] = B
        # per step: primal: state up on the first step of the block only, state + dtc + obj down; adjoint: state + adjoint + 4 scalars
        # up, adjoint down, source gradient down once per block
        res["h2d"] = (5 * C * s) / B + s + (10 * C * s + 4 * s)
        res["d2h"] = (5 * C * s + 2 * s) + (5 * C * s) + (5 * C * s) / B
        # ---- the same block through adfvm_b200.blocks: the block's states, the adjoint fields and the gradient
        # accumulator stay in HBM; per block the host sends the start state + the adjoint fields and reads back the end state, the
        # adjoint fields and the accumulated source gradient
        from adfvm_b200 import blocks
        C0 = case.mesh.nInternalCells

        def block2():                  # one checkpoint of Adjoint.run: the block forward from its stored start state, then backwards
            return blocks.adjoint_blocks(f, fa, case.inputs, [hstate], B, B, case.dt, hadj0, None)
        block2()
        barrier()
        t0 = time.perf_counter()
        block2()
        barrier()
        tb = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        res["e2e_blocks_s_per_step"] = tb.item() / B
        # ---- the per-step protocol once more with the opt-in state cache (PrimalFunction(state_cache=n)): the states the forward
        # calls returned are read-only and have device copies, the adjoint calls that get the very same arrays back skip their upload
        f.c.state_cache = B + 1
        block()
        barrier()
        t0 = time.perf_counter()
        block()
        barrier()
        tc = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        res["e2e_cache_s_per_step"] = tc.item() / B
        res["e2e_cache_hits"] = f.state_cache_hits
        res["e2e_cache_h2d"] = 5 * C0 * s * (1 + 2.0 / B) + 4 * s
        f.c.state_cache = 0
        res["e2e_blocks_h2d"] = (10 * C0 * s) / B                      # the block's start state + the adjoint fields
        res["e2e_blocks_d2h"] = (10 * C0 * s) / B                      # the adjoint fields + the accumulated source gradient
    f.c.close()
    return res


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
    ap.add_argument("--cpu-n", type=int, default=64, help="box size per process of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fp32", action="store_true", help="skip the fp32 sub-record")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling sub-record (N > 1)")
    ap.add_argument("--no-parity", action="store_true")
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
    from adfvm_b200 import cases, _lib
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
    torch.cuda.set_stream(torch.cuda.Stream())            # an explicit stream (events below are recorded on it); the legacy
    stream = torch.cuda.current_stream().cuda_stream      # default stream would rule out the CUDA graphs of whole steps

    # ---- parity first: a fast wrong kernel is not a result
    parity = None
    if not args.no_parity:
        parity = parity_check(rank, world, local, stream)
        log("rank %d: parity %s" % (rank, parity))
        if not (parity["maxerr"] < 1e-10):
            raise SystemExit("parity check failed: %r" % (parity,))
        if world > 1 and not parity["early_tiles"] > 0:
            raise SystemExit("parity check did not exercise the overlapped path (early_tiles = 0): %r" % (parity,))

    # ---- workload
    t_case = time.time()
    if world > 1:
        from adfvm_b200 import decompose
        case = decompose.periodic_box_rank(args.n, rank, world, dtype)
    else:
        case = cases.periodic_box(args.n, dtype)
    log("rank %d: case arrays built in %.1f s" % (rank, time.time() - t_case))
    m = measure(case, dtype, rank, world, local, stream, K, W)
    C = m["C"]
    total_ms, tp_ms, ta_ms = m["total_ms"], m["tp_ms"], m["ta_ms"]

    # ---- sub-records
    fp32 = None
    if not args.no_fp32 and args.dtype == "f64":
        case32 = cases.Case(case.mesh.astype(np.float32), case.spec, case.state, case.source, {}, case.dt, np.float32)
        m32 = measure(case32, np.float32, rank, world, local, stream, K, W, want_kernels=True, want_e2e=False)
        if rank == 0:
            Bp32, Ba32, kb32, _ = algorithmic_bytes(4)
            fp32 = {"dtype": "f32", "value": 6 * C * world * K / (m32["total_ms"] * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": m32["total_ms"] / K,
                    "primal": 3 * C * world * K / (m32["tp_ms"] * 1e-3) / 1e6, "adjoint": 3 * C * world * K / (m32["ta_ms"] * 1e-3) / 1e6,
                    "device_bytes": m32["device_bytes"], "tolerance": "1e-5 relative against the fp64 oracle (tests/test_gpu_parity.py)"}
            peak_ = _peak()[0]
            fp32["primal_frac"] = Bp32 * 3 * C / (m32["tp_ms"] / K * 1e-3) / 1e9 / peak_
            fp32["adjoint_frac"] = Ba32 * 3 * C / (m32["ta_ms"] / K * 1e-3) / 1e9 / peak_
            fp32["kernels"] = {k: {"ms_per_pass": ms / m32["kernel_reps"] / PASSES[k],
                                   "frac": kb32[k] * C / (ms / m32["kernel_reps"] / PASSES[k] * 1e-3) / 1e9 / peak_}
                               for k, (cnt, ms) in m32["kernel_report"].items() if k in kb32}
        del case32
    strong = None
    if world > 1 and not args.no_strong:
        from adfvm_b200 import decompose
        p = decompose.factor3(world)
        G = 368 if args.n >= 368 else args.n
        blk = tuple(G // p[d] for d in range(3))
        del case
        cs = decompose.periodic_box_rank(blk, rank, world, dtype)
        ms_ = measure(cs, dtype, rank, world, local, stream, K, W, want_kernels=False, want_e2e=False)
        if rank == 0:
            cells = blk[0] * blk[1] * blk[2] * world
            strong = {"scaling": "strong", "global_cells": cells, "block": list(blk), "value": 6 * cells * K / (ms_["total_ms"] * 1e-3) / 1e6,
                      "unit": UNIT, "ms_per_step": ms_["total_ms"] / K, "early_tiles": ms_["early_tiles"], "tiles": ms_["tiles"],
                      "note": "the %d^3 mesh of the 1-GPU run split over %d ranks; efficiency = value / (N x the N=1 value)" % (G, world)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    B_p, B_a, kb, strict = algorithmic_bytes(s)
    cells_total = C * world
    value = 2 * 3 * cells_total * K / (total_ms * 1e-3) / 1e6
    vp = 3 * cells_total * K / (tp_ms * 1e-3) / 1e6
    va = 3 * cells_total * K / (ta_ms * 1e-3) / 1e6
    peak, peak_src = _peak()
    kernels = {}
    rep, nrep = m["kernel_report"], m["kernel_reps"]
    for name, (cnt, ms) in rep.items():
        per = ms / cnt
        e = {"launches_per_step": cnt / nrep, "ms_per_launch": per, "share": None}
        if name in kb:
            e["ms_per_pass"] = ms / nrep / PASSES[name]
            e["algorithmic_GBs"] = kb[name] * C / (e["ms_per_pass"] * 1e-3) / 1e9
            e["frac"] = e["algorithmic_GBs"] / peak
            if name in strict:
                e["strict_bytes_per_cell"] = strict[name]
                e["strict_frac"] = strict[name] * C / (e["ms_per_pass"] * 1e-3) / 1e9 / peak
        kernels[name] = e
    tot = sum(ms for _, ms in rep.values())
    for name, (cnt, ms) in rep.items():
        kernels[name]["share"] = ms / tot
    dom = max((k for k in kernels if k in kb), key=lambda k: rep[k][1])
    traffic, traffic_src = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = "%s_%s_n%d" % (dom, args.dtype, args.n)
        if key in tr:
            traffic, traffic_src = int(tr[key]), "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum per pass at this size (%s)" % tr.get(key + "_source", "profiles/")
        else:
            key = "%s_%s_n128" % (dom, args.dtype)
            if key in tr:
                traffic = int(tr[key] * (C / 128.0 ** 3))
                traffic_src = "n128-scaled: ncu --set full capture at 128^3 scaled by the cell count (no capture at this size)"
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["algorithmic_GBs"], "peak": peak, "unit": "GB/s",
            "frac": kernels[dom]["frac"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
            "algorithmic_bytes_per_cell": kb[dom], "strict_bytes_per_cell": strict.get(dom), "strict_frac": kernels[dom].get("strict_frac"),
            "stage_model": {"primal_bytes_per_cell_stage": B_p, "adjoint_bytes_per_cell_stage": B_a,
                            "primal_GBs": B_p * 3 * C / (tp_ms / K * 1e-3) / 1e9, "adjoint_GBs": B_a * 3 * C / (ta_ms / K * 1e-3) / 1e9,
                            "primal_frac": B_p * 3 * C / (tp_ms / K * 1e-3) / 1e9 / peak,
                            "adjoint_frac": B_a * 3 * C / (ta_ms / K * 1e-3) / 1e9 / peak}}
    # the dominant kernel's other ceiling: the fp64 pipe (SASS count of fp64 instructions per face evaluation, DESIGN.md section 5;
    # an fp64 warp instruction holds one of the 4 x 148 sub-partition pipes for two cycles)
    # 904 = ncu of the specialised (Roe + Sutherland) kernel at 128^3: fp64 pipe active 56.0 % of 727 us x 1.965 GHz x 592 sub-partitions
    # / 2 cycles per instruction / 262 144 warp-evaluations (profiles/r2f_tiles_f64.summary.txt); the generic instantiation: 953
    FP64_PER_EVAL = {"flux_grad_tile": 904}
    if dom in FP64_PER_EVAL and args.dtype == "f64":
        mhz = (m["clocks"] or {}).get("sm_mhz") or 1965.0
        floor_ms = FP64_PER_EVAL[dom] * 4.0 * C / 32 * 2 / (4 * 148 * mhz * 1e6) * 1e3
        roof["fp64_pipe"] = {"fp64_instructions_per_face_evaluation": FP64_PER_EVAL[dom], "evaluations_per_cell": 4.0,
                             "floor_ms_per_pass": floor_ms, "pipe_busy_frac": floor_ms / kernels[dom]["ms_per_pass"],
                             "hbm_floor_ms_per_pass": kb[dom] * C / peak / 1e6,
                             "note": "this kernel is bound by fp64 issue (two warps per scheduler at 254 registers), not by HBM: with the "
                                     "pipe saturated it would reach hbm_floor/floor of the HBM peak"}
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        log("cpu baseline ...")
        try:
            cpu = cpu_baseline(args.cpu_n, 1, 3)
        except Exception as e:          # noqa: BLE001
            cpu = {"error": repr(e)}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": "periodic_hex_box_n%d" % args.n, "cells_per_gpu": C, "rk_stages": 3,
                       "step": "1 primal step + 1 adjoint step (incl. forward recompute)",
                       "l2": "working set per step (%.0f MB) exceeds the 126 MB L2; no explicit flush" % (m["device_bytes"] / 1e6),
                       "device_bytes": m["device_bytes"], "host_rss_gb": round(resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e6, 1),
                       "size_note": size_note, "static_upload_and_tile_plan_s": round(m["static_s"], 1),
                       "early_tiles": m["early_tiles"], "tiles": m["tiles"]},
            "primal": vp, "adjoint": va, "primal_ms": tp_ms / K, "adjoint_ms": ta_ms / K,
            "e2e": {"value": 2 * 3 * cells_total / m["e2e_s_per_step"] / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(m["h2d"]),
                    "d2h_bytes_per_step": int(m["d2h"]), "ms_per_step": m["e2e_s_per_step"] * 1e3,
                    "note": "PrimalFunction/AdjointFunction.__call__ with pinned host buffers, one checkpoint block of %d steps driven like "
                            "the reference's Adjoint.run (forward-mode primal calls returning every state, then primal_grad backwards: state + "
                            "adjoint up, adjoint down every call, source gradient once per block)" % m["e2e_block"]},
            "e2e_blocks": {"value": 2 * 3 * cells_total / m["e2e_blocks_s_per_step"] / 1e6, "unit": UNIT,
                           "h2d_bytes_per_step": int(m["e2e_blocks_h2d"]), "d2h_bytes_per_step": int(m["e2e_blocks_d2h"]),
                           "ms_per_step": m["e2e_blocks_s_per_step"] * 1e3,
                           "note": "not the headline: the same block of %d steps through adfvm_b200.blocks (SURVEY 8(f)-1: states, adjoint fields "
                                   "and gradient accumulator resident; host traffic once per block) - what the drivers' loops cost when they call "
                                   "run_block instead of one Function call per step" % m["e2e_block"]} if "e2e_blocks_s_per_step" in m else None,
            "e2e_state_cache": {"value": 2 * 3 * cells_total / m["e2e_cache_s_per_step"] / 1e6, "unit": UNIT,
                                "h2d_bytes_per_step": int(m["e2e_cache_h2d"]), "d2h_bytes_per_step": int(m["d2h"]),
                                "ms_per_step": m["e2e_cache_s_per_step"] * 1e3, "cache_hits": m["e2e_cache_hits"],
                                "note": "not the headline (opt-in, differs from the reference's semantics: returned state arrays are read-only): "
                                        "the same per-step protocol as e2e with PrimalFunction(state_cache=n) - states passed back as the very "
                                        "array objects that were returned are taken from their device copies instead of being uploaded again"}
            if "e2e_cache_s_per_step" in m else None,
            "gpu_launches": int(m["launches"]), "clocks": m["clocks"], "roofline": roof, "kernels": kernels, "cpu_baseline": cpu,
            "parity": parity, "parity_maxerr": parity["maxerr"] if parity else None, "early_tiles": parity["early_tiles"] if parity else None,
            "fp32": fp32, "strong": strong, "wall_s": round(time.time() - T0, 1)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _peak():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    return peak, ("measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)")


if __name__ == "__main__":
    main()
