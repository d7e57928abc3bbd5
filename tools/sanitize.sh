#!/bin/bash
# compute-sanitizer over the small fixtures (race detection / sanitizers).  Run under gpurun, ONE GPU:
#   gpurun --timeout 900 -- 'bash tools/sanitize.sh [full]'
# memcheck: out-of-bounds / misaligned global+shared accesses; racecheck: shared-memory hazards; synccheck: barrier misuse;
# initcheck: reads of uninitialised device memory.  The persistent CG kernel spins on global flags by design: racecheck only
# looks at shared memory, so it stays meaningful.  Logs land in gpurun_out/sanitize_*.log; summaries are copied to profiles/.
# Default = the quick set (every tool on the default value pass + memcheck on the opt-in passes and on elasticity);
# "full" = every tool on every value-pass variant.
set -u
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
RUN="python tools/profile_step.py --cells 10 --max-it 24"
TMO=${SANITIZE_TIMEOUT:-150}
one() {   # tool, asm mode, command...
  local TOOL=$1 ASM=$2; shift 2
  if [ "$ASM" = default ]; then unset PFEM_ASM; else export PFEM_ASM=$ASM; fi
  local LOG=gpurun_out/sanitize_${TOOL}_${ASM}.log
  timeout $TMO $CS --tool $TOOL --error-exitcode 3 "$@" > $LOG 2>&1
  echo "$TOOL $ASM exit=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $LOG | tail -1)"
  unset PFEM_ASM
}
if [ "${1:-quick}" = full ]; then
  for TOOL in memcheck racecheck synccheck initcheck; do
    for ASM in default rows tiled tiled2; do one $TOOL $ASM $RUN; done
  done
else
  for TOOL in memcheck racecheck synccheck initcheck; do one $TOOL default $RUN; done
  one memcheck rows $RUN
  one memcheck tiled2 $RUN
fi
TOOLTAG=elasticity
unset PFEM_ASM
timeout $TMO $CS --tool memcheck --error-exitcode 3 python tools/profile_step.py --kind elasticity --cells 4 --max-it 60 > gpurun_out/sanitize_memcheck_elasticity.log 2>&1
echo "memcheck elasticity exit=$? $(grep 'ERROR SUMMARY' gpurun_out/sanitize_memcheck_elasticity.log | tail -1)"
