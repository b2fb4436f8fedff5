timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_tests.log 2>&1; echo tests rc=$?; tail -4 gpurun_out/r2z_tests.log
timeout 300 python profiles/scripts/timing.py 2>&1 | tail -8
