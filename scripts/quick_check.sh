mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/b_new.json 2> gpurun_out/b_new.err
BALM_SYNC_PHASES=1 BALM_NO_BUFFER_REUSE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/b_old.json 2> gpurun_out/b_old.err
python - <<'PY'
import json
for f in ("new","old"):
    d=json.loads(open("gpurun_out/b_%s.json"%f).read().strip().splitlines()[-1]); print(f, round(d["value"],2), round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["phases_ms"].items()}, d.get("sweeps"), d["e2e"]["value"], d["e2e"]["set_voxels_ms"])
PY
tail -2 gpurun_out/b_new.err
