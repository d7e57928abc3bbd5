# r02 session 2, call 7 (1 GPU): replicated partial records; full parity file + driver tests; sync floor
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_driver_cpp.py -x -q > gpurun_out/c7_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/c7_pytest.log
timeout 300 python tools/sync_floor.py 16 32 > gpurun_out/c7_sync_floor.log 2>&1; cat gpurun_out/c7_sync_floor.log
for sy in last lean; do
  PFEM_PCG_SYNC=$sy timeout 300 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/c7_c2_$sy.log 2>&1; echo c2 $sy rc=$?; tail -1 gpurun_out/c7_c2_$sy.log | cut -c1-200
done
