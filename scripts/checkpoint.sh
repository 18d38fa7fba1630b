# Round-2 checkpoint on one B200: GPU tests, the bench lines kept under profiles/, solve A/B against cuSOLVER, the
# factorisation timeline, ncu launch list + full-set captures of the dominant kernels. Everything lands in gpurun_out/;
# scripts/summarize_profiles.py turns the ncu outputs into the tracked summaries under profiles/.
mkdir -p gpurun_out; (time timeout 1200 python -m pytest tests -m gpu -q) 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_tensor.json 2> gpurun_out/bench_err.log; tail -2 gpurun_out/bench_err.log
timeout 300 python bench.py --steps 10 --warmup 3 --precision fp64 --no-cpu > gpurun_out/bench_r2_fp64.json 2>> gpurun_out/bench_err.log
timeout 300 python bench.py --steps 10 --warmup 3 --poses 200 --voxels 20000 --cpu-sample-small 512 --cpu-sample-large 2048 > gpurun_out/bench_r2_c2.json 2>> gpurun_out/bench_err.log
timeout 300 python bench.py --steps 10 --warmup 3 --poses 50 --voxels 2000 --cpu-sample-small 1000 --cpu-sample-large 2000 > gpurun_out/bench_r2_c1.json 2>> gpurun_out/bench_err.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_r2_reference.json 2>> gpurun_out/bench_err.log
timeout 200 python scripts/solve_ab.py > gpurun_out/solve_ab_r2.log 2>/dev/null; cat gpurun_out/solve_ab_r2.log | cut -c1-200
timeout 100 python scripts/dag_trace.py 500 > gpurun_out/dag_trace_r2.txt 2>&1; grep "ms_solve\|chain:\|fine\|workers\|far CTA" gpurun_out/dag_trace_r2.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_tensor.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"syrk_tc|voxel_sums|obs_pass" -s 9 -c 8 -o gpurun_out/prof_r2_tensor -f python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/b2.log 2>&1
# the persistent factorisation (cooperative, spin-waits on counters): its own capture, bounded by a short timeout
timeout 240 ncu --set full --clock-control none -k regex:"ldl_dag|ldl_back|solve_residual" -s 6 -c 4 -o gpurun_out/prof_r2_solve -f python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/b3.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 300 python scripts/assoc_perf.py 2>&1 | tail -3
python scripts/print_bench.py
