timeout 900 python -m pytest tests -m gpu -x -q -k "not train" > gpurun_out/r2m_tests.log 2>&1; echo tests rc=$?; tail -5 gpurun_out/r2m_tests.log
timeout 300 python profiles/small_call_split.py 2>&1 | tail -1
timeout 300 python profiles/small_calls.py 2>&1 | tail -1
