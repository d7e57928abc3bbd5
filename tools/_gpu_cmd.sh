# r02 session 2, call 24 (2 GPUs): dedicated pushing warp: multi-GPU parity subset + tet100 at 2 GPUs
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py -x -q -k "(p2p-rows and (tet10-2 or gen_tet24-2)) or (variants and gen_tet24-2-last-tag) or (ilu0 and p2p-beam)" > gpurun_out/i_pytest_multi.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/i_pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29761 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 --cells 100 > gpurun_out/i_tet100_2gpu.log 2>&1; echo rc=$?
python - <<'PY'
import json
l=[x for x in open('gpurun_out/i_tet100_2gpu.log') if x.startswith('{"metric')]
if l:
    d=json.loads(l[-1]); print('tet100 2 GPUs: value %.4g  its %s  us/iter %.2f' % (d['value'], d['iterations_per_step'], 1e6*d['config']['dof']/d['value']))
PY
