#!/bin/bash
# time the value pass on C5 for each rows-per-CTA shape (t_assemble is CUDA-event timed inside the library)
for R in 256 128 64 32; do
  echo "R=$R: $(PFEM_ASM_ROWS=$R python tools/profile_step.py --cells ${1:-200} --max-it 1 2>&1 | tail -1)"
done
