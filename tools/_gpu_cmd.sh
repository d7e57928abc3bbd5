# r02 session 2, call 4 (2 GPUs): lean barrier fix, tag-validated halo; multi-GPU parity + C5 at 2 GPUs, both flavours
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "persistent_and_multi" > gpurun_out/c4_pytest1.log 2>&1; echo pytest1 rc=$?; tail -2 gpurun_out/c4_pytest1.log
timeout 120 python tools/sync_floor.py 16 > gpurun_out/c4_sync_floor.log 2>&1; cat gpurun_out/c4_sync_floor.log
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -k "(2-29621 or 2-29622 or 2-29624) and p2p-rows or (variants and gen_tet24-2)" > gpurun_out/c4_pytest_multi.log 2>&1; echo pytest multi rc=$?; tail -3 gpurun_out/c4_pytest_multi.log
run2() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1; }
PFEM_PCG_SYNC=last PFEM_PCG_HALO=flag timeout 600 bash -c "$(declare -f run2); run2 29701" > gpurun_out/c4_c5_2gpu_last_flag.log 2>&1; echo rc=$?; tail -1 gpurun_out/c4_c5_2gpu_last_flag.log | cut -c1-250
PFEM_PCG_SYNC=lean PFEM_PCG_HALO=flag timeout 600 bash -c "$(declare -f run2); run2 29702" > gpurun_out/c4_c5_2gpu_lean_flag.log 2>&1; echo rc=$?; tail -1 gpurun_out/c4_c5_2gpu_lean_flag.log | cut -c1-250
timeout 600 bash -c "$(declare -f run2); run2 29703" > gpurun_out/c4_c5_2gpu_default.log 2>&1; echo rc=$?; tail -1 gpurun_out/c4_c5_2gpu_default.log | cut -c1-250
