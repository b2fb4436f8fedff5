for mode in sleep spin; do
  if [ $mode = spin ]; then export SWT_TRAIN_SPIN_WAIT=1; fi
  CUDA_VISIBLE_DEVICES=0 timeout 600 python profiles/train_scale.py --types 1000000 > gpurun_out/r2ae_a.log 2>&1 &
  CUDA_VISIBLE_DEVICES=1 timeout 600 python profiles/train_scale.py --types 1000000 > gpurun_out/r2ae_b.log 2>&1 &
  wait
  echo "== two concurrent single-GPU runs, wait=$mode"; tail -1 gpurun_out/r2ae_a.log | cut -c150-260; tail -1 gpurun_out/r2ae_b.log | cut -c150-260
  echo "== 2-GPU sharded, wait=$mode"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 profiles/train_scale.py --types 1000000 2>&1 | grep -v OMP | tail -1 | cut -c150-260
done
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null
