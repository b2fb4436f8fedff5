timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_tests.log 2>&1; echo tests rc=$?; tail -15 gpurun_out/r2o_tests.log
for n in 1 2; do
  if [ $n = 1 ]; then timeout 600 python profiles/train_scale.py --types 1000000 --check-oracle-steps 2 2>&1 | tail -1
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 profiles/train_scale.py --types 1000000 2>&1 | tail -2
       SWT_NO_PEER_EXCHANGE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 profiles/train_scale.py --types 1000000 --no-peer 2>&1 | tail -2
  fi
done
timeout 600 python profiles/train_scale.py --types 1000000 --timing 2>&1 | tail -30
