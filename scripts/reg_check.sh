mkdir -p gpurun_out
for i in 1 2; do
BALM_TRACE_REG=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu 2> gpurun_out/reg_new$i.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('new', d['e2e']['value'], d['e2e']['set_voxels_ms'])"
grep balm_set_voxels gpurun_out/reg_new$i.err
BALM_NO_BUFFER_REUSE=1 BALM_TRACE_REG=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu 2> gpurun_out/reg_old$i.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('old', d['e2e']['value'], d['e2e']['set_voxels_ms'])"
grep balm_set_voxels gpurun_out/reg_old$i.err
done
