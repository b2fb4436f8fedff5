timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2al_tests.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/r2al_tests.log
(time python bench.py --steps 10 --warmup 3) > gpurun_out/r2al_bench.json 2> gpurun_out/r2al_bench.err; echo bench rc=$?; tail -c 300 gpurun_out/r2al_bench.err
timeout 300 python profiles/small_calls.py 2>&1 | tail -1
ncu --set full --import-source on --clock-control none -k regex:"encode_|memo_clear" --launch-skip 36 --launch-count 12 -f -o gpurun_out/r02_final_bpe python profiles/prof_encode.py 1000000000 3 bpe > gpurun_out/r02_final_ncu_bpe.log 2>&1
