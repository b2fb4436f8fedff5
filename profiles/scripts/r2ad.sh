timeout 900 python -m pytest tests -m gpu -x -q -k "multi_gpu" > gpurun_out/r2ad_tests.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/r2ad_tests.log
for n in 8 4; do timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n profiles/train_scale.py --types 1000000 2>&1 | grep -v OMP | tail -1 | cut -c1-1400; done
for g in 0 3 7; do CUDA_VISIBLE_DEVICES=$g timeout 300 python profiles/train_scale.py --types 1000000 2>&1 | tail -1 | cut -c1-200; done
