# r02 session 2, call 16 (2 GPUs): stash tests; BASELINE configs[2] (C3) and configs[3] (C4) at 2 GPUs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -k "stash and 2-2967" > gpurun_out/c16_pytest_stash.log 2>&1; echo pytest stash rc=$?; tail -3 gpurun_out/c16_pytest_stash.log
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 3 --warmup 3 --e2e-steps 1 --workload $2 --no-cpu-baseline; }
timeout 600 bash -c "$(declare -f run); run 29721 c3" > gpurun_out/c16_c3_2gpu.log 2>&1; echo c3 rc=$?; tail -1 gpurun_out/c16_c3_2gpu.log | cut -c1-300
timeout 900 bash -c "$(declare -f run); run 29722 c4" > gpurun_out/c16_c4_2gpu.log 2>&1; echo c4 rc=$?; tail -1 gpurun_out/c16_c4_2gpu.log | cut -c1-300
timeout 300 python bench.py --workload c3 --steps 3 --warmup 3 --e2e-steps 1 > gpurun_out/c16_c3_1gpu.log 2>&1; echo c3 1gpu rc=$?; tail -1 gpurun_out/c16_c3_1gpu.log | cut -c1-300
