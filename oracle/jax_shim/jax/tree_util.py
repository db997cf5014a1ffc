"""Minimal pytrees: None, tuple, list, dict, namedtuple and registered classes."""
_REGISTRY = {}


def register_pytree_node(cls, flatten, unflatten):
  _REGISTRY[cls] = (flatten, unflatten)


def _is_namedtuple(x):
  return isinstance(x, tuple) and hasattr(x, '_fields')


def _children(x):
  """Returns (children, rebuild) or None for leaves."""
  t = type(x)
  if t in _REGISTRY:
    fl, unfl = _REGISTRY[t]
    data, meta = fl(x)
    return list(data), (lambda ch, meta=meta, unfl=unfl: unfl(meta, tuple(ch)))
  if x is None:
    return [], (lambda ch: None)
  if _is_namedtuple(x):
    return list(x), (lambda ch, t=t: t(*ch))
  if isinstance(x, tuple):
    return list(x), (lambda ch: tuple(ch))
  if isinstance(x, list):
    return list(x), (lambda ch: list(ch))
  if isinstance(x, dict):
    keys = sorted(x.keys())
    return [x[k] for k in keys], (lambda ch, keys=keys, t=t: t(zip(keys, ch)))
  return None


def tree_map(f, tree, *rest, is_leaf=None):
  if is_leaf is not None and is_leaf(tree):
    return f(tree, *rest)
  c = _children(tree)
  if c is None:
    return f(tree, *rest)
  children, rebuild = c
  rest_children = []
  for r in rest:
    rc = _children(r)
    if rc is None or len(rc[0]) != len(children):
      raise ValueError('tree structure mismatch')
    rest_children.append(rc[0])
  return rebuild([tree_map(f, ch, *[rc[i] for rc in rest_children], is_leaf=is_leaf)
                  for i, ch in enumerate(children)])


def tree_leaves(tree, is_leaf=None):
  if is_leaf is not None and is_leaf(tree):
    return [tree]
  c = _children(tree)
  if c is None:
    return [tree]
  out = []
  for ch in c[0]:
    out.extend(tree_leaves(ch, is_leaf))
  return out


def tree_all(tree):
  return all(tree_leaves(tree))


def tree_reduce(f, tree, initializer=None):
  import functools
  leaves = tree_leaves(tree)
  if initializer is None:
    return functools.reduce(f, leaves)
  return functools.reduce(f, leaves, initializer)


class _TreeDef:
  def __init__(self, tree):
    self.tree = tree


def tree_flatten(tree, is_leaf=None):
  return tree_leaves(tree, is_leaf), _TreeDef(tree)


def tree_unflatten(treedef, leaves):
  it = iter(leaves)
  return tree_map(lambda _: next(it), treedef.tree)


def tree_structure(tree):
  return _TreeDef(tree)
