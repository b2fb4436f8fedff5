timeout 600 python -m pytest tests -m gpu -x -q -k "not train" > gpurun_out/r2e_tests.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/r2e_tests.log
timeout 300 python - > gpurun_out/r2e_timing.log 2>&1 <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
sys.argv = ["x", "1000000000", "2"]
from subword_tokenizers_b200 import device
device.tune("timing", 1)
exec(open("profiles/prof_encode.py").read())
PY
cat gpurun_out/r2e_timing.log
timeout 300 python profiles/many_types.py > gpurun_out/r2e_many.log 2>&1; tail -12 gpurun_out/r2e_many.log
