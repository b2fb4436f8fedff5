"""Offline loader for the upstream reference (test infrastructure only).

Imports the unmodified reference classes from /root/reference and
builds the offline stand-in for ``AutoTokenizer.from_pretrained("bert-base-uncased")``
(reference cli.py:163): the reference only ever calls
``tokenizer.backend_tokenizer.pre_tokenizer.pre_tokenize_str`` (source/utils.py:27) and
BertPreTokenizer has no parameters, so a locally constructed one is behaviourally identical.

Only the golden-vector generator (tests/golden/make_golden.py) and CPU-side differential
tests (skipped when /root/reference is absent) may import this module.
"""
import os
import sys

REFERENCE_ROOT = os.environ.get("SWT_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "source", "bpe.py"))


def make_hf_tokenizer():
    """The offline shim of SURVEY.md §8(c)."""
    import tokenizers
    from tokenizers import models, normalizers, pre_tokenizers
    import transformers

    tk = tokenizers.Tokenizer(models.WordPiece({"[UNK]": 0}, unk_token="[UNK]"))
    tk.normalizer = normalizers.BertNormalizer(lowercase=True)
    tk.pre_tokenizer = pre_tokenizers.BertPreTokenizer()
    return transformers.PreTrainedTokenizerFast(tokenizer_object=tk, unk_token="[UNK]")


def load_reference():
    """Returns (NaiveBPE, FastBPE, NaiveWP, FastWP) classes of the reference."""
    if not reference_available():
        raise RuntimeError("reference checkout not present at %s" % REFERENCE_ROOT)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from source.bpe import NaiveBPE, FastBPE  # type: ignore
    from source.wordpiece import NaiveWP, FastWP  # type: ignore
    return NaiveBPE, FastBPE, NaiveWP, FastWP
