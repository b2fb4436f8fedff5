set -x
python profiles/prof_encode.py 1000000000 3 > gpurun_out/r2b_prof.log 2>&1
SWT_TIMING=1 python - > gpurun_out/r2b_timing.log 2>&1 <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
sys.argv = ["x", "1000000000", "2"]
from subword_tokenizers_b200 import device
device.tune("timing", 1)
exec(open("profiles/prof_encode.py").read())
PY
ncu --set full --import-source on --clock-control none -k regex:"encode_(count|emit)_kernel" --launch-skip 6 --launch-count 2 -f -o gpurun_out/r2b_wp python profiles/prof_encode.py 1000000000 3 wp > gpurun_out/r2b_ncu_wp.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"encode_(count|emit)_kernel" --launch-skip 6 --launch-count 2 -f -o gpurun_out/r2b_bpe python profiles/prof_encode.py 1000000000 3 bpe > gpurun_out/r2b_ncu_bpe.log 2>&1
ls -la gpurun_out
