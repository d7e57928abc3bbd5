# r02 session 2, call 22 (1 GPU): pattern-pass temporaries in persistent scratch: full suite, PC comparison, bench (twice: e2e stability)
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/h_pytest_full.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/h_pytest_full.log | head -2
timeout 120 python tools/pc_compare.py 40 100 > gpurun_out/h_pc_compare.log 2>&1; cat gpurun_out/h_pc_compare.log | cut -c1-200
timeout 400 python bench.py > gpurun_out/h_bench_c5_1gpu.log 2>&1; echo bench rc=$?; tail -1 gpurun_out/h_bench_c5_1gpu.log | cut -c1-200
timeout 300 python bench.py --no-cpu-baseline --no-parity --e2e-steps 4 --steps 3 > gpurun_out/h_bench_c5_1gpu_b.log 2>&1; echo bench2 rc=$?
python - <<'PY'
import json
for f in ('h_bench_c5_1gpu', 'h_bench_c5_1gpu_b'):
    l=[x for x in open(f'gpurun_out/{f}.log') if x.startswith('{"metric')]
    if l:
        d=json.loads(l[-1]); print(f, 'value %.4g' % d['value'], 'e2e ms %.1f' % d['e2e']['ms_per_step'], {k: round(v, 1) for k, v in d['e2e']['stage_ms'].items()})
PY
