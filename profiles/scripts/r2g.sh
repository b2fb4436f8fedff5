ncu --set full --import-source on --clock-control none -k regex:"encode_(count|emit)_kernel" --launch-skip 6 --launch-count 2 -f -o gpurun_out/r2g_bpe python profiles/prof_encode.py 1000000000 3 bpe > gpurun_out/r2g_ncu_bpe.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"encode_(count|emit)_kernel" --launch-skip 6 --launch-count 2 -f -o gpurun_out/r2g_wp python profiles/prof_encode.py 1000000000 3 wp > gpurun_out/r2g_ncu_wp.log 2>&1
timeout 400 python profiles/many_types.py > gpurun_out/r2g_many.log 2>&1; tail -5 gpurun_out/r2g_many.log
