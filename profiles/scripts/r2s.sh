SWT_TRAIN_NO_GRAPH=1 timeout 600 python profiles/train_scale.py --types 1000000 2>&1 | tail -2 | cut -c1-330
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 profiles/train_scale.py --types 1000000 2>&1 | grep -v OMP | tail -3 | cut -c1-330
SWT_TRAIN_NO_GRAPH=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 profiles/train_scale.py --types 1000000 2>&1 | grep -v OMP | tail -3 | cut -c1-330
