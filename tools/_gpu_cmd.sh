# r02 session 2, call 19 (8 GPUs): C5 and C4 at 8 GPUs (the driver's SCALE run repeats C5; this is the builder's own check)
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 3 --warmup 3 --e2e-steps 1 --workload $2 --no-cpu-baseline; }
timeout 600 bash -c "$(declare -f run); run 29741 c5" > gpurun_out/c19_c5_8gpu.log 2>&1; echo c5 rc=$?; tail -1 gpurun_out/c19_c5_8gpu.log | cut -c1-300
timeout 400 bash -c "$(declare -f run); run 29742 c4" > gpurun_out/c19_c4_8gpu.log 2>&1; echo c4 rc=$?; tail -1 gpurun_out/c19_c4_8gpu.log | cut -c1-300
