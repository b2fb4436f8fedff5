"""list / dict / set subclasses that count their mutations.

The reference's classes read their live Python containers on every call (``merges_list``, ``_bpe_ranks``, ``vocab``;
source/bpe.py:126,212-217, source/wordpiece.py:138); here those containers are mirrored by device tables, so the classes
must notice ANY change -- including in-place edits that keep the length -- without hashing 20,000 entries per
``tokenize`` call.  Every mutating method bumps ``version``; the encoders compare ``(id, version)`` in O(1).
"""
from __future__ import annotations


def _make(base, name):
    fn = getattr(base, name)

    def method(self, *args, **kwargs):
        self.version += 1
        return fn(self, *args, **kwargs)
    method.__name__ = name
    return method


def _tracked(base, mutators, cls_name):
    ns = {"version": 0}
    for m in mutators:
        ns[m] = _make(base, m)
    return type(cls_name, (base,), ns)


TrackedList = _tracked(list, ["append", "extend", "insert", "remove", "pop", "clear", "sort", "reverse", "__setitem__",
                              "__delitem__", "__iadd__", "__imul__"], "TrackedList")
TrackedDict = _tracked(dict, ["__setitem__", "__delitem__", "pop", "popitem", "clear", "update", "setdefault", "__ior__"],
                       "TrackedDict")
TrackedSet = _tracked(set, ["add", "discard", "remove", "pop", "clear", "update", "difference_update", "intersection_update",
                            "symmetric_difference_update", "__ior__", "__iand__", "__isub__", "__ixor__"], "TrackedSet")


def tracked_attribute(name: str, kind):
    """Property that wraps whatever is assigned to it into the tracked container type (a copy), so that replacing the
    attribute and mutating it in place are both visible as a new ``(id, version)`` stamp."""
    slot = "_tracked_" + name

    def get(self):
        return getattr(self, slot)

    def set_(self, value):
        setattr(self, slot, value if type(value) is kind else kind(value))
    return property(get, set_)


def stamp(container):
    return (id(container), container.version)
