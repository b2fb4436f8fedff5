timeout 900 python -m pytest tests -m gpu -x -q -k "multi_gpu" > gpurun_out/r2p_tests.log 2>&1; echo tests rc=$?; tail -5 gpurun_out/r2p_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 profiles/train_scale.py --types 1000000 --timing 2>&1 | grep -v "^\[swt train timing\].*rank1" | tail -34
