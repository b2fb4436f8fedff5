timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2aj_tests.log 2>&1; echo tests rc=$?; tail -4 gpurun_out/r2aj_tests.log
(time python bench.py --steps 10 --warmup 3) > gpurun_out/r2aj_bench.json 2> gpurun_out/r2aj_bench.err; echo bench rc=$?; tail -c 300 gpurun_out/r2aj_bench.err
