"""Per-kernel timing of the FastWP / FastBPE encode call over the 1 GB bench stream (swt_tune("timing", 1))."""
import sys, os
sys.path.insert(0, os.getcwd())
sys.argv = ["x", "1000000000", "2"] + sys.argv[1:]
from subword_tokenizers_b200 import device
device.tune("timing", 1)
exec(open("profiles/prof_encode.py").read())
