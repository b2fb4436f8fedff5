timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 profiles/peer_barrier.py 2>&1 | tail -3
nvidia-smi topo -m 2>&1 | head -8
