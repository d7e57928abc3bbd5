# r02 session 2, call 9 (2 GPUs): halo flavours (flag | tag | tagf) at 2 GPUs, tet100 and C5
mkdir -p gpurun_out
run2() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 $2; }
port=29710
for cells in 100 200; do
for halo in flag tag tagf; do
  port=$((port+1))
  PFEM_PCG_HALO=$halo timeout 600 bash -c "$(declare -f run2); run2 $port '--cells $cells'" > gpurun_out/c9_${cells}_$halo.log 2>&1; echo $cells $halo rc=$?
  python - <<PY
import json
l=[x for x in open('gpurun_out/c9_${cells}_$halo.log') if x.startswith('{"metric')]
if l:
    d=json.loads(l[-1]); its=d.get('iterations_per_step'); print('  value %.4g  ms/step %.2f  its %s  us/iter %.2f' % (d['value'], d['ms_per_step'], its, 1e6*d['config']['dof']*1.0/d['value']))
else:
    print(open('gpurun_out/c9_${cells}_$halo.log').read()[-1500:])
PY
done
done
