mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "--- solve only: fused / unfused / unfused no-lookahead"
timeout 120 python scripts/solve_only.py
BALM_NO_FUSED_PANEL=1 timeout 120 python scripts/solve_only.py
BALM_NO_FUSED_PANEL=1 BALM_NO_LOOKAHEAD=1 timeout 120 python scripts/solve_only.py
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/b_new.json 2> gpurun_out/b_new.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/b_new.json").read().strip().splitlines()[-1]); print(round(d["value"],2), round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["phases_ms"].items()}, d["e2e"] and round(d["e2e"]["value"],2), d.get("sweeps"))
PY
tail -3 gpurun_out/b_new.err
