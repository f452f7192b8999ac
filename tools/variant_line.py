"""one summary line of a bench.py JSON line (used by tools/variants.sh)"""
import json
import sys

try:
    d = json.load(open(sys.argv[2]))
    r = d["roofline"]
    print("%-64s %7.1f Mrays/s e2e %7.1f  N_int %.2f N_prim %.2f it %.2f" % (
        sys.argv[1], d["value"], d["e2e"]["value"], r["n_int_per_ray"], r["n_prim_per_ray"], r["phantom_iterations_per_ray"]),
        {k: (v["steps"], round(v["lanes_per_step"], 1)) for k, v in r["warp_scheduler_rank0"].items()})
except Exception as e:  # noqa: BLE001
    print(sys.argv[1], "FAILED", e)
