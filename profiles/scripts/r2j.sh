timeout 900 python -m pytest tests -m gpu -x -q -k "not train" > gpurun_out/r2j_tests.log 2>&1; echo tests rc=$?; tail -15 gpurun_out/r2j_tests.log
timeout 300 python profiles/small_calls.py > gpurun_out/r2j_small.log 2>&1; tail -3 gpurun_out/r2j_small.log
