"""The C-ABI library loads on a CPU-only box and exports every symbol include/swt.h declares (no compute)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _header_functions():
    src = open(os.path.join(ROOT, "include", "swt.h"), encoding="utf-8").read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(swt_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_a_plain_c_abi():
    src = open(os.path.join(ROOT, "include", "swt.h"), encoding="utf-8").read()
    assert 'extern "C"' in src
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    assert "torch" not in code.lower() and "at::" not in code and "std::" not in code     # no torch / C++ types
    fns = _header_functions()
    assert len(fns) >= 25 and "swt_bpe_encode" in fns and "swt_wp_encode" in fns and "swt_bpe_train_steps" in fns


def test_library_exports_every_declared_symbol():
    from subword_tokenizers_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _header_functions():
        assert hasattr(lib, name), "libswt.so does not export %s" % name
    # the Python binding declares a prototype for every function of the header and nothing else
    assert sorted(_lib.SIGNATURES) == _header_functions()
    assert _lib.load().swt_abi_version() == 1


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from subword_tokenizers_b200 import FastBPE, FastWP, make_hf_tokenizer
    from subword_tokenizers_b200._lib import SwtError
    hf = make_hf_tokenizer()
    fb = FastBPE(hf)
    fb.merges_list = [("a", "b")]
    fb._bpe_ranks = {("a", "b"): 0}
    with pytest.raises(SwtError):
        fb.tokenize("ab ab")
    with pytest.raises(SwtError):
        fb.train(["ab ab"], 10)
    fw = FastWP(hf)
    with pytest.raises(SwtError):
        fw.train(["ab ab"], 4)
    from subword_tokenizers_b200 import NaiveWP
    with pytest.raises(SwtError):
        NaiveWP(hf).train(["ab ab"], 4)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "subword_tokenizers_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text and "swt_oracle" not in text, f
