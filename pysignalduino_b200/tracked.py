"""Mutation-tracking containers for the protocol dict.

The reference reads ``self._protocols`` afresh on every call, so a caller may edit a protocol row between two
``demodulate`` calls (the reference's own tests do: tests/test_manchester_protocols.py:54,84).  The device works on a
compiled copy of the table, so the drop-in class has to notice edits.  Hashing the whole dict per call cost 0.9 ms —
twice the reference's entire MS decode — so the dict, its rows and their lists are wrapped in subclasses that bump a
shared version counter on every mutating method instead; ``SDProtocols.engine()`` compares one integer.

``TrackedDict`` is a ``dict`` and ``TrackedList`` a ``list`` (``isinstance`` checks such as
message_unsynced.py:71 keep working; ``json.dumps`` and ``copy.deepcopy`` too).
"""
from __future__ import annotations

from typing import Any


class Version:
    """The shared counter: ``n`` changes whenever any container that holds it was mutated."""

    __slots__ = ("n",)

    def __init__(self) -> None:
        self.n = 0


def wrap(value: Any, ver: Version) -> Any:
    """Recursively replace plain dicts / lists by tracked ones bound to ``ver``."""
    if isinstance(value, TrackedDict) or isinstance(value, TrackedList):
        if value._ver is ver:
            return value
        value = dict(value) if isinstance(value, dict) else list(value)
    if isinstance(value, dict):
        return TrackedDict(value, ver)
    if isinstance(value, list):
        return TrackedList(value, ver)
    return value


class TrackedDict(dict):
    __slots__ = ("_ver",)

    def __init__(self, src=(), ver: Version | None = None):
        super().__init__()
        self._ver = ver if ver is not None else Version()
        for k, v in dict(src).items():
            dict.__setitem__(self, k, wrap(v, self._ver))

    def _touch(self) -> None:
        self._ver.n += 1

    def __setitem__(self, key, value):
        dict.__setitem__(self, key, wrap(value, self._ver))
        self._touch()

    def __delitem__(self, key):
        dict.__delitem__(self, key)
        self._touch()

    def __ior__(self, other):
        self.update(other)
        return self

    def update(self, *args, **kwargs):
        for k, v in dict(*args, **kwargs).items():
            dict.__setitem__(self, k, wrap(v, self._ver))
        self._touch()

    def setdefault(self, key, default=None):
        if key not in self:
            dict.__setitem__(self, key, wrap(default, self._ver))
            self._touch()
        return dict.__getitem__(self, key)

    def pop(self, *args):
        r = dict.pop(self, *args)
        self._touch()
        return r

    def popitem(self):
        r = dict.popitem(self)
        self._touch()
        return r

    def clear(self):
        dict.clear(self)
        self._touch()

    def __reduce_ex__(self, protocol):
        # copy / pickle as plain data: the copy gets its own counter when it is wrapped again
        return (TrackedDict, (dict(self),))


class TrackedList(list):
    __slots__ = ("_ver",)

    def __init__(self, src=(), ver: Version | None = None):
        self._ver = ver if ver is not None else Version()
        super().__init__(wrap(v, self._ver) for v in src)

    def _touch(self) -> None:
        self._ver.n += 1

    def __setitem__(self, idx, value):
        if isinstance(idx, slice):
            value = [wrap(v, self._ver) for v in value]
        else:
            value = wrap(value, self._ver)
        list.__setitem__(self, idx, value)
        self._touch()

    def __delitem__(self, idx):
        list.__delitem__(self, idx)
        self._touch()

    def __iadd__(self, other):
        self.extend(other)
        return self

    def __imul__(self, k):
        list.__imul__(self, k)
        self._touch()
        return self

    def append(self, v):
        list.append(self, wrap(v, self._ver))
        self._touch()

    def extend(self, it):
        list.extend(self, [wrap(v, self._ver) for v in it])
        self._touch()

    def insert(self, i, v):
        list.insert(self, i, wrap(v, self._ver))
        self._touch()

    def pop(self, *args):
        r = list.pop(self, *args)
        self._touch()
        return r

    def remove(self, v):
        list.remove(self, v)
        self._touch()

    def clear(self):
        list.clear(self)
        self._touch()

    def sort(self, **kw):
        list.sort(self, **kw)
        self._touch()

    def reverse(self):
        list.reverse(self)
        self._touch()

    def __reduce_ex__(self, protocol):
        return (TrackedList, (list(self),))
