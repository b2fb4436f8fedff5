(time python bench.py --steps 10 --warmup 3) > gpurun_out/r2an_bench.json 2> gpurun_out/r2an_bench.err; echo bench rc=$?; tail -c 300 gpurun_out/r2an_bench.err
