#!/usr/bin/env python
"""dev: print the interesting parts of bench.py JSON lines (python tools/show_bench.py file.json ...)"""
import json
import sys

for f in sys.argv[1:]:
    d = json.load(open(f))
    print("=====", f)
    print("value %.0f  ms/step %.3f  e2e %.0f (h2d %.1f MB, d2h %.2f MB)" % (d["value"], d["ms_per_step"], d["e2e"]["value"],
          d["e2e"]["h2d_bytes_per_step"] / 1e6, d["e2e"]["d2h_bytes_per_step"] / 1e6))
    c = d["config"]
    print({k: c[k] for k in c if k not in ("workload", "layout", "l2", "variants", "parallelism")})
    r = d["roofline"]
    print("K2 roofline: %.0f GB/s frac %.3f ms %.3f  traffic %s" % (r["achieved"], r["frac"], r["ms_per_launch"], r["traffic"]))
    k1 = d["kernels"]["K1_affinity"]
    print("K1: %.3f ms (prep %.3f + aff %.3f) frac %.3f   K3 %.3f ms" % (k1["ms_per_launch"], k1["K1a_prep_ms"], k1["K1_affinity_ms"], k1["frac"], d["kernels"]["K3_arun_ms"]))
    print("parity", d["parity_checked"], d["parity_mismatches"], d["parity_oracle"], " clocks", d["clocks"])
    if "fast_path" in d:
        fp = d["fast_path"]
        print("fast:", {k: fp[k] for k in fp if k not in ("what", "kernels", "e2e")}, "e2e %.0f" % fp["e2e"]["value"])
        print("   K1 %.3f (prep %.3f)  K2 %.3f" % (fp["kernels"]["K1_affinity"]["ms_per_launch"], fp["kernels"]["K1_affinity"]["K1a_prep_ms"], fp["kernels"]["K2_solver"]["ms_per_launch"]))
    if "variants" in c:
        print("cone:", json.dumps(c["variants"]["cone"]))
    if "cpu_baseline" in d:
        cb = d["cpu_baseline"]
        print("cpu: %.0f /s on %d cores (%s); latency mode %s" % (cb["value"], cb["cores"], cb["kind"], cb.get("latency_mode", {}).get("value")))
