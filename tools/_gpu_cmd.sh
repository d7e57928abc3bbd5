CTILE_VARIANTS=0 timeout 100 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --error-exitcode 3 python tools/ctile_check.py tet12 tria40 > gpurun_out/fast_memcheck.log 2>&1; echo memcheck rc=$?; tail -2 gpurun_out/fast_memcheck.log | cut -c1-200
CTILE_VARIANTS=0 timeout 400 python tools/ctile_check.py tet40 tet100 tria300 tet200 > gpurun_out/fast_check.log 2>&1; echo rc=$?
python -c "
import sys,json
for l in open('gpurun_out/fast_check.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print({k:d.get(k) for k in ('case','variant','rows_ms','ctile_ms','values_within_1e12','rhs_within_1e12','max_abs_diff','run_to_run_bit_identical','mode','error')})
"
