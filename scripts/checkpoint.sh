# Round checkpoint on one B200: GPU tests, the bench lines kept under profiles/, ncu launch list + full-set captures.
mkdir -p gpurun_out; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1_tensor.json 2> gpurun_out/bench_err.log; tail -2 gpurun_out/bench_err.log
timeout 300 python bench.py --steps 10 --warmup 3 --precision fp64 --no-cpu > gpurun_out/bench_r1_fp64.json 2>> gpurun_out/bench_err.log
timeout 300 python bench.py --steps 10 --warmup 3 --poses 200 --voxels 20000 --cpu-sample-voxels 512 > gpurun_out/bench_r1_c2.json 2>> gpurun_out/bench_err.log
timeout 300 python bench.py --steps 10 --warmup 3 --poses 50 --voxels 2000 --cpu-sample-voxels 2000 > gpurun_out/bench_r1_c1.json 2>> gpurun_out/bench_err.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_reference.json 2>> gpurun_out/bench_err.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r1_tensor.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"syrk_tc|voxel_stats|obs_pass" -s 9 -c 8 -o gpurun_out/prof_r1_tensor -f python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/b2.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"ldl_" -s 220 -c 8 -o gpurun_out/prof_r1_solve -f python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/b3.log 2>&1
timeout 300 python scripts/assoc_perf.py 2>&1 | tail -3
python - <<'PY'
import json
for f in ("tensor","fp64","c2","c1","reference"):
    try:
        d=json.loads(open("gpurun_out/bench_r1_%s.json"%f).read().strip().splitlines()[-1]); print(f, round(d["value"],4), round(d["ms_per_step"],3), {k:round(v,3) for k,v in d.get("phases_ms",{}).items()}, d["e2e"] and round(d["e2e"]["value"],4), d["cpu_baseline"] and d["cpu_baseline"]["value"], d.get("clocks"), d.get("sweeps"))
    except Exception as e: print(f, "ERR", e)
PY
