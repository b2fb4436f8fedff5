set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "not train" > gpurun_out/r2c_tests.log 2>&1; echo tests rc=$?; tail -5 gpurun_out/r2c_tests.log
timeout 300 python - > gpurun_out/r2c_timing.log 2>&1 <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
sys.argv = ["x", "1000000000", "2"]
from subword_tokenizers_b200 import device
device.tune("timing", 1)
exec(open("profiles/prof_encode.py").read())
PY
cat gpurun_out/r2c_timing.log
