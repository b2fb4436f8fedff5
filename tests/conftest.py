import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    path = os.path.join(GOLDEN, name)
    if name.endswith(".gz"):
        with gzip.open(path, "rt", encoding="utf-8") as f:
            return json.load(f)
    with open(path, encoding="utf-8") as f:
        return json.load(f)


@pytest.fixture(scope="session")
def hf_tokenizer():
    """Offline stand-in for AutoTokenizer.from_pretrained('bert-base-uncased')."""
    from subword_tokenizers_b200.hf_shim import make_hf_tokenizer
    return make_hf_tokenizer()


@pytest.fixture(scope="session")
def pre_tokenize(hf_tokenizer):
    pre = hf_tokenizer.backend_tokenizer.pre_tokenizer

    def f(text):
        return [w for w, _ in pre.pre_tokenize_str(text.lower())]
    return f


@pytest.fixture(scope="session")
def random_cases():
    return load_golden("ref_random_cases.json.gz")
