#!/usr/bin/env python
"""Pretty-print bench.py JSON lines: python tools/bench_summary.py gpurun_out/x.log ..."""
import json
import sys
for path in sys.argv[1:]:
    for l in open(path):
        if not l.startswith("{"):
            continue
        d = json.loads(l)
        if "unavailable" in d:
            print(path, d); continue
        print("%s: %s n_gpus=%d value=%.1f primal=%.1f adjoint=%.1f (ms %.3f / %.3f) e2e=%.1f launches=%s" % (
            path, d["dtype"], d["n_gpus"], d["value"], d.get("primal", 0), d.get("adjoint", 0), d.get("primal_ms", 0), d.get("adjoint_ms", 0),
            d["e2e"]["value"], d.get("gpu_launches")))
        r = d.get("roofline")
        if r:
            sm = r.get("stage_model", {})
            print("   roofline %s frac=%.3f achieved=%.0f GB/s; stage model primal %.3f adjoint %.3f; clocks %s" % (
                r.get("kernel"), r["frac"], r["achieved"], sm.get("primal_frac", 0), sm.get("adjoint_frac", 0), d.get("clocks")))
        for k, v in sorted(d.get("kernels", {}).items(), key=lambda kv: -kv[1]["share"]):
            print("   %-16s %5.1f%%  %.4f ms x %.1f/step %s" % (k, 100 * v["share"], v["ms_per_launch"], v["launches_per_step"],
                                                            ("frac %.3f" % v["frac"]) if "frac" in v else ""))
