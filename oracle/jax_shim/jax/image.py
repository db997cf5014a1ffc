class ResizeMethod:
  NEAREST = 'nearest'
  LINEAR = 'linear'

  @staticmethod
  def from_string(s):
    return s


def resize(*a, **k):
  raise NotImplementedError('resize')
