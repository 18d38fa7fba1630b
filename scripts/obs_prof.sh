mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/b_new.json 2> gpurun_out/b_new.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/b_new.json").read().strip().splitlines()[-1]); print(round(d["value"],2), round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["phases_ms"].items()}, d.get("sweeps"))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"obs_pass" -s 3 -c 2 -o gpurun_out/prof_obs_fused python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/b2.log 2>&1
tail -2 gpurun_out/b2.log
