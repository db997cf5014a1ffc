"""Global numeric configuration (the analogue of `jax.config`).

The reference casts inputs with `jax.dtypes.canonicalize_dtype(float64)`
(`_src/stax/requirements.py:794`): float32 unless `jax_enable_x64` is set.  The
same switch exists here: `config.update('enable_x64', True)` or the environment
variable `NT_B200_ENABLE_X64=1`.
"""
import os

import numpy as np


class _Config:
  def __init__(self):
    self.enable_x64 = os.environ.get('NT_B200_ENABLE_X64', '0') not in ('0', '', 'false', 'False')
    self.device = int(os.environ.get('NT_B200_DEVICE', os.environ.get('LOCAL_RANK', '0')))
    # bytes of device workspace per context; 0 = half of the free memory (capped)
    self.workspace_bytes = int(os.environ.get('NT_B200_WORKSPACE_BYTES', '0'))
    self.disable_fusion = os.environ.get('NT_B200_NO_FUSION', '0') not in ('0', '')
    # stencil kernels with one Conv+ABRelu layer per launch (one HBM round trip per layer)
    self.per_layer = os.environ.get('NT_B200_PER_LAYER', '0') not in ('0', '')
    # x2=None: compute the full n x n square like the reference instead of triangle + mirror
    self.full_square = os.environ.get('NT_B200_FULL_SQUARE', '0') not in ('0', '')

  def update(self, name, value):
    if name in ('enable_x64', 'jax_enable_x64'):
      self.enable_x64 = bool(value)
    elif name in ('device', 'workspace_bytes', 'disable_fusion', 'per_layer', 'full_square'):
      setattr(self, name, type(getattr(self, name))(value))
    else:
      raise AttributeError(f'unknown config option {name!r}')

  @property
  def dtype(self):
    return np.float64 if self.enable_x64 else np.float32


config = _Config()
