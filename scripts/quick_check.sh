mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/b_new.json 2> gpurun_out/b_new.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/b_new.json").read().strip().splitlines()[-1]); print(round(d["value"],2), round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["phases_ms"].items()}, d.get("sweeps"))
PY
tail -2 gpurun_out/b_new.err
