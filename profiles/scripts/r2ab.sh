timeout 900 python -m pytest tests -m gpu -x -q -k "multi_gpu or bpe_train or smoke" > gpurun_out/r2ab_tests.log 2>&1; echo tests rc=$?; tail -4 gpurun_out/r2ab_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 profiles/train_scale.py --types 1000000 2>&1 | grep -v OMP | tail -1 | cut -c1-1400
