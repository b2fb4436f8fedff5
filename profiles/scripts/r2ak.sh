timeout 900 python -m pytest tests -m gpu -x -q -k "split_count or large_stream or more_than or fastbpe or smoke" > gpurun_out/r2ak_tests.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/r2ak_tests.log
timeout 300 python profiles/scripts/timing.py 2>&1 | tail -8
timeout 400 python - <<'PY'
import sys, os, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch, bench, bench_data as BD
from subword_tokenizers_b200 import device, packing as P
# many-type stream (bench shape) FastBPE with the pretrained merges: slow words + rate
dev = torch.device("cuda", 0)
mat, lens = BD.synth_type_table(2_000_000, 1)
t_arena, t_off = BD.table_to_utf8(mat, lens)
zs = BD.ZipfStream(t_arena, t_off, 200_000, 100)
d_arena, d_off, n_words, _ = zs.device_stream(500_000_000, dev)
benc = device.BpeEncoder(P.BpeTables([tuple(p) for p in bench.load_golden("pretrained_bpe_merges.json.gz")]))
r = bench.time_encode(benc, d_arena, d_off, n_words, 3, 2, None, 1)
print("many-type FastBPE ms", np.mean(r["step_ms"]), "slow", r["slow_words"], "warm gate word", int(r["status"].cpu()[6]))
PY
