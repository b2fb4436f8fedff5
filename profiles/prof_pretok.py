"""Times the device pre-tokenizer (swt_pretok_count / swt_pretok_write) + FastWP encode on raw text built from the
bench Zipf stream (words joined by single spaces)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from subword_tokenizers_b200 import device, packing as P, _lib
from subword_tokenizers_b200.utils import naive_wp_encode_ids

nbytes = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
dev = torch.device("cuda", 0)
stream = bench.ZipfStream.train5k(0)
d_arena, d_off, n_words, off32 = stream.device_stream(nbytes, dev)
d_text, n_text = bench.device_text(d_arena, d_off, n_words)
tab = P.WpTables(bench.load_golden("ref_wp_train5k_v8000_vocab.json.gz"))
wenc = device.WpEncoder(tab, naive_wp_encode_ids("##", tab))
pt = device.Pretokenizer.get()
lib = _lib.load()
ws = torch.empty(lib.swt_pretok_workspace_bytes(n_text), dtype=torch.uint8, device=dev)
st = torch.empty(8, dtype=torch.int32, device=dev)
out_arena = torch.empty(n_text + 16, dtype=torch.uint8, device=dev)
out_off = torch.empty(n_words + 2, dtype=torch.int32, device=dev)
sp = torch.cuda.current_stream().cuda_stream
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for it in range(4):
    ev[0].record()
    device.check(lib.swt_pretok_count(pt._handle, d_text.data_ptr(), n_text, ws.data_ptr(), ws.numel(), st.data_ptr(), sp))
    ev[1].record()
    device.check(lib.swt_pretok_write(pt._handle, d_text.data_ptr(), n_text, ws.data_ptr(), ws.numel(), out_arena.data_ptr(), n_text,
                                      out_off.data_ptr(), None, n_words + 2, n_words, int(d_arena.numel()), st.data_ptr(), sp))
    ev[2].record(); torch.cuda.synchronize()
    print("pretok count %.3f ms  write %.3f ms  (%d text bytes, %d words)" % (ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), n_text, n_words))
s = st.cpu().numpy()
assert int(s[0]) == 0 and int(s[1]) == n_words, s
assert torch.equal(out_arena[: d_arena.numel()], d_arena) and torch.equal(out_off[: n_words + 1], d_off[: n_words + 1])
print("pre-tokenizer output equals the packed stream")
