"""Generates bert_pretok_classes.json: the character classes of the Rust `tokenizers` BertPreTokenizer, probed through the
library itself (pre_tokenize_str on "a" + ch + "b" for every code point): `space` = characters it drops as whitespace
(Rust char::is_whitespace), `punct` = characters it isolates (is_bert_punc: ASCII punctuation or Unicode P*).
The reference reaches this code at source/utils.py:27.  Run in an environment that has `tokenizers`; the output is
shipped with the package and spot-checked against the live library at run time (device.BertClasses.verify)."""
import json, os, sys
import tokenizers
from tokenizers import pre_tokenizers

pre = pre_tokenizers.BertPreTokenizer()
space, punct = [], []
for cp in range(0x110000):
    if 0xD800 <= cp < 0xE000:
        continue
    n = len(pre.pre_tokenize_str("a" + chr(cp) + "b"))
    if n == 2:
        space.append(cp)
    elif n == 3:
        punct.append(cp)
    elif n != 1:
        raise SystemExit("unexpected split for U+%04X" % cp)


def ranges(cps):
    out = []
    for c in cps:
        if out and out[-1][1] + 1 == c:
            out[-1][1] = c
        else:
            out.append([c, c])
    return out


dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bert_pretok_classes.json")
json.dump({"tokenizers_version": tokenizers.__version__, "space": ranges(space), "punct": ranges(punct)}, open(dst, "w"))
print(len(space), "space,", len(punct), "punct ->", dst)
