CUDA_VISIBLE_DEVICES=0 timeout 600 python profiles/train_scale.py --types 1000000 > gpurun_out/r2ac_a.log 2>&1 &
CUDA_VISIBLE_DEVICES=1 timeout 600 python profiles/train_scale.py --types 1000000 > gpurun_out/r2ac_b.log 2>&1 &
wait
tail -1 gpurun_out/r2ac_a.log | cut -c1-260; tail -1 gpurun_out/r2ac_b.log | cut -c1-260
