#!/bin/bash
# compute-sanitizer over the small fixtures (race detection / sanitizers).  Run under gpurun, ONE GPU:
#   gpurun --timeout 900 -- 'bash tools/sanitize.sh'
# memcheck: out-of-bounds / misaligned global+shared accesses; racecheck: shared-memory hazards; synccheck: barrier misuse;
# initcheck: reads of uninitialised device memory.  The persistent CG kernel spins on global flags by design: racecheck only
# looks at shared memory, so it stays meaningful.  Logs land in gpurun_out/sanitize_*.log; copy the summaries to profiles/.
set -u
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
RUN="python tools/profile_step.py --cells 12 --max-it 40"
for TOOL in memcheck racecheck synccheck initcheck; do
  for ASM in default tiled tiled2; do
    if [ "$ASM" = default ]; then unset PFEM_ASM; else export PFEM_ASM=$ASM; fi
    timeout 600 $CS --tool $TOOL --error-exitcode 3 $RUN > gpurun_out/sanitize_${TOOL}_${ASM}.log 2>&1
    echo "$TOOL $ASM exit=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_${TOOL}_${ASM}.log | tail -1)"
  done
done
unset PFEM_ASM
timeout 600 $CS --tool memcheck --error-exitcode 3 python tools/profile_step.py --kind elasticity --cells 4 --max-it 60 > gpurun_out/sanitize_memcheck_elasticity.log 2>&1
echo "memcheck elasticity exit=$? $(grep 'ERROR SUMMARY' gpurun_out/sanitize_memcheck_elasticity.log | tail -1)"
