"""B200-native implementation of the three hot paths of phtryll/subword-tokenizers.

The four classes keep the reference's surface (train / tokenize / encode_word / save_resources /
load_resources, merges.json / vocab.json layout); FastBPE tokenization, FastWP tokenization and BPE
training run as hand-written sm_100a CUDA kernels behind the C ABI of include/swt.h (libswt.so).
"""
from .bpe import FastBPE, NaiveBPE
from .wordpiece import FastWP, NaiveWP
from .utils import SubwordTokenizer, WPTrie_E2E
from .hf_shim import make_hf_tokenizer

__all__ = ["NaiveBPE", "FastBPE", "NaiveWP", "FastWP", "SubwordTokenizer", "WPTrie_E2E",
           "make_hf_tokenizer"]
