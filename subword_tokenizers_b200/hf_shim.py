"""Locally constructed BERT normalizer/pre-tokenizer (no network).

The reference loads ``AutoTokenizer.from_pretrained("bert-base-uncased")`` (cli.py:163) and then only
ever calls ``tokenizer.backend_tokenizer.pre_tokenizer.pre_tokenize_str`` (source/utils.py:27).
BertPreTokenizer has no parameters, so building it locally is behaviourally identical for that
call; BASELINE.json's north star asks for exactly this because there is no network.
"""


def make_hf_tokenizer():
    import tokenizers
    from tokenizers import models, normalizers, pre_tokenizers
    import transformers

    tk = tokenizers.Tokenizer(models.WordPiece({"[UNK]": 0}, unk_token="[UNK]"))
    tk.normalizer = normalizers.BertNormalizer(lowercase=True)
    tk.pre_tokenizer = pre_tokenizers.BertPreTokenizer()
    return transformers.PreTrainedTokenizerFast(tokenizer_object=tk, unk_token="[UNK]")
