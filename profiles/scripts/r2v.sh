timeout 1500 python -m pytest tests -m gpu -x -q -k "config5 or small_call" --durations=5 > gpurun_out/r2v_tests.log 2>&1; echo tests rc=$?; tail -12 gpurun_out/r2v_tests.log
