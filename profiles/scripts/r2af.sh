timeout 900 python -m pytest tests -m gpu -x -q -k "not train and not config5 and not multi_gpu" > gpurun_out/r2af_tests.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/r2af_tests.log
timeout 300 python profiles/scripts/timing.py 2>&1 | tail -8
