"""One line per bench JSON left in gpurun_out/ by scripts/checkpoint.sh (value, ms/step, phases, e2e, CPU arm, clocks)."""
import json
import sys

for f in sys.argv[1:] or ("tensor", "fp64", "c2", "c1", "reference"):
    try:
        d = json.loads(open("gpurun_out/bench_r2_%s.json" % f).read().strip().splitlines()[-1])
        e2e, cpu = d.get("e2e"), d.get("cpu_baseline")
        print(f, round(d["value"], 4), round(d["ms_per_step"], 3),
              {k: round(v, 3) for k, v in d.get("phases_ms", {}).items()},
              "e2e", e2e and (round(e2e["value"], 3), e2e.get("set_voxels_ms") and round(e2e["set_voxels_ms"], 2)),
              "cpu", cpu and (cpu["value"], cpu.get("ref_headers_check")), d.get("clocks"), d.get("sweeps"))
    except Exception as e:
        print(f, "ERR", e)
