# speculative single sweep + stats hand-over: GPU tests, then the bench with and without them
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/b_new.json 2> gpurun_out/b_new.err
BALM_NO_SPEC=1 BALM_NO_STATS_CACHE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/b_old.json 2> gpurun_out/b_old.err
BALM_NO_SPEC=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/b_nospec.json 2>> gpurun_out/b_old.err
python - <<'PY'
import json
for f in ("new","old","nospec"):
    try:
        d=json.loads(open("gpurun_out/b_%s.json"%f).read().strip().splitlines()[-1]); print(f, round(d["value"],2), round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["phases_ms"].items()}, d["e2e"] and round(d["e2e"]["value"],2), d.get("sweeps"), d["roofline"].get("digit_planes"))
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/b_new.err
