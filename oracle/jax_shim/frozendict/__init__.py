"""Minimal `frozendict` stand-in (hashable immutable dict)."""


class frozendict(dict):
  def __hash__(self):
    return hash(frozenset(self.items()))

  def _ro(self, *a, **k):
    raise TypeError('frozendict is immutable')

  __setitem__ = __delitem__ = clear = pop = popitem = setdefault = update = _ro
