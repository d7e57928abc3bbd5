#!/bin/bash
# Round-2 starting point for the tiled value pass (run under gpurun, ONE GPU):
#   1. default vs tiled (gather) vs tiled2 (scatter) on C5: bit-identity + device time per pass
#   2. one `ncu --set full` capture of each tiled kernel (compile has -lineinfo: the source page maps to the .cuh)
#   3. summaries for profiles/
# Usage: gpurun --timeout 900 -- 'bash tools/profile_tiled.sh'
set -u
mkdir -p gpurun_out
N=${1:-200}
python tools/tiled_check.py "$N" 2>&1 | tee gpurun_out/tiled_check_r02.log
for ASM in tiled tiled2; do
  PFEM_ASM=$ASM ncu --set full --clock-control none --import-source on -k regex:assemble_tiled -c 1 \
      -o gpurun_out/asm_${ASM}_r02 -f python tools/profile_step.py --cells "$N" --max-it 1 > gpurun_out/prof_${ASM}.log 2>&1
  python tools/ncu_summary.py gpurun_out/asm_${ASM}_r02.ncu-rep > gpurun_out/asm_${ASM}_r02_summary.txt 2>/dev/null || true
done
ls -la gpurun_out | tail -12
