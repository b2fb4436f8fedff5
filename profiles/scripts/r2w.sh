(time python bench.py --steps 10 --warmup 3) > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo bench rc=$?; tail -c 1500 gpurun_out/r2w_bench.err
for mb in 16 64; do python bench.py --steps 5 --warmup 3 --no-also --e2e-batch-mb $mb 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('batch', d['e2e']['batch_mib'], 'e2e', d['e2e']['value'], 'ceiling', d['e2e']['copy_ceiling']['value'], 'frac', d['e2e']['frac_of_copy_ceiling'])"; done
