timeout 900 python -m pytest tests -m gpu -x -q -k "not train" > gpurun_out/r2i_tests.log 2>&1; echo tests rc=$?; tail -15 gpurun_out/r2i_tests.log
timeout 300 python profiles/small_calls.py > gpurun_out/r2i_small.log 2>&1; tail -3 gpurun_out/r2i_small.log
echo "== default (4/4)"; timeout 300 python profiles/scripts/timing.py 2>&1 | tail -6
for v in "-DSWT_COUNT_CTAS=3 -DSWT_EMIT_CTAS=3" "-DSWT_COUNT_CTAS=5 -DSWT_EMIT_CTAS=5"; do
  echo "== $v"; SWT_NVCC_EXTRA="$v" python -m subword_tokenizers_b200.build --force > /dev/null 2>&1
  timeout 300 python profiles/scripts/timing.py 2>&1 | tail -6
done
