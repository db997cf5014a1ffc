def _ni(*a, **k):
  raise NotImplementedError


gelu = relu = softmax = _ni
