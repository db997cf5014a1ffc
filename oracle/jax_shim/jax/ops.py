def segment_sum(*a, **k):
  raise NotImplementedError('segment_sum')
