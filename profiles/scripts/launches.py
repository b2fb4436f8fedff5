import sys, os
sys.path.insert(0, os.getcwd())
from subword_tokenizers_b200 import device
for kv in sys.argv[4:]:
    k, v = kv.split("="); device.tune(k, int(v))
sys.argv = sys.argv[:4]
exec(open("profiles/prof_encode.py").read())
