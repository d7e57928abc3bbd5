# r02 session 2, call 14 (1 GPU): the complete GPU suite, timed
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests/ -x -q -m gpu --durations=12 ) > gpurun_out/c14_pytest_full.log 2>&1; echo pytest rc=$?; tail -25 gpurun_out/c14_pytest_full.log
