"""Condense tools/bench_grid.py JSON lines to one short line per config."""
import json, sys
for l in sys.stdin:
    try:
        d = json.loads(l)
    except Exception:
        print(l.rstrip()[:300]); continue
    print("%s %7.2f Mcorr/s  %5.0f GB/s  frac %.3f  fwd %.3f cells %.3f ms  T=%s ctas=%s det=%d" % (
        d["config"], d["correlations_per_s"] / 1e6, d["contract_gbs"], d["frac_of_hbm_peak"], d["stage_ms"]["fwd_ms"],
        d["stage_ms"]["cells_ms"], d.get("cell_threads"), d.get("cell_ctas"), d["detected"]))
