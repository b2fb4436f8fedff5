set -x
ncu --set full --import-source on --clock-control none -k regex:"encode_|memo_clear" --launch-skip 21 --launch-count 7 -f -o gpurun_out/r02_final_wp python profiles/prof_encode.py 1000000000 3 wp > gpurun_out/r02_final_ncu_wp.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"encode_|memo_clear" --launch-skip 21 --launch-count 7 -f -o gpurun_out/r02_final_bpe python profiles/prof_encode.py 1000000000 3 bpe > gpurun_out/r02_final_ncu_bpe.log 2>&1
python bench.py --steps 2 --warmup 1 --no-also > gpurun_out/r02_b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-also > gpurun_out/r02_ncu_bench.log 2>&1
ncu --set full --clock-control none -k regex:"tokenize_small" --launch-skip 20 --launch-count 1 -f -o gpurun_out/r02_small python profiles/small_calls.py > gpurun_out/r02_small.log 2>&1
ls -la gpurun_out
