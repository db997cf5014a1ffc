"""The closed forms the CUDA kernels evaluate, restated in NumPy and checked on the CPU against the reference
formulas (`_src/stax/elementwise.py:444-455` ABRelu, `:67-112` Erf): the committed polynomial coefficients of
G(c) = acos(c)/sqrt(1-c^2) are parsed out of the .cuh sources, so a typo in a constant fails here, without a GPU."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, 'neural-tangents_b200', 'csrc')


def _body(path, signature):
  src = open(os.path.join(CSRC, path)).read()
  start = src.index(signature)
  return src[start:src.index('\n}\n', start)]


def _numbers(text):
  return [float(x) for x in re.findall(r'(?<![\w.])-?\d+\.\d+(?:e-?\d+)?', text)]


def G_exact(c):
  c = np.asarray(c, dtype=np.float64)
  return np.where(c < 1, np.arccos(np.minimum(c, 1 - 1e-300)) / np.sqrt((1 - c) * (1 + c) + 1e-300), 1.0)


def test_fp32_fit_of_G():
  coef = _numbers(_body('fused_kernels.cuh', '__device__ __forceinline__ float acos_over_sin(float c) {'))
  assert len(coef) == 9                                      # degree 8, Horner order (highest first)
  c = np.linspace(0, 1, 20001)[:-1]
  p = np.zeros_like(c)
  for a in coef:
    p = p * c + a
  assert np.abs(p - G_exact(c)).max() < 1.5e-7
  # the packed kernel evaluates -G with negated coefficients
  neg = _numbers(_body('stage_packed.cuh', '__device__ __forceinline__ float2 neg_acos_over_sin2(float2 c) {'))
  np.testing.assert_allclose(neg, [-a for a in coef], rtol=0, atol=0)


def test_fp64_fit_of_G_even_odd_horner():
  body = _body('fused_kernels.cuh', '__device__ __forceinline__ double acos_over_sin(double c) {')
  lines = [l for l in body.splitlines() if '__fma_rn' in l or l.strip().startswith('double e =') or l.strip().startswith('double o =')]
  e_coef, o_coef, cur = [], [], None
  for l in lines:
    if l.strip().startswith('double e ='):
      cur = e_coef
      cur.append(_numbers(l)[0])
    elif l.strip().startswith('double o ='):
      cur = o_coef
      cur.append(_numbers(l)[0])
    elif 'return' not in l:
      cur.append(_numbers(l)[-1])
  assert len(e_coef) == 9 and len(o_coef) == 8               # degree 16 = even part (9) + odd part (8)
  c = np.linspace(0, 1, 20001)[:-1]
  y = c * c
  e = np.zeros_like(c)
  for a in e_coef:
    e = e * y + a
  o = np.zeros_like(c)
  for a in o_coef:
    o = o * y + a
  assert np.abs(c * o + e - G_exact(c)).max() < 1e-13


def _abrelu_reference(K, T, q1, q2, a, b):
  """elementwise.py:444-455 in float64."""
  prod = q1 * q2
  s = np.sqrt(np.maximum(prod - K * K, 0))
  theta = np.where((s == 0) & (K == 0), np.pi / 2, np.arctan2(s, K))
  dot = (a * a + b * b) / 2 - (a - b) ** 2 / (2 * np.pi) * theta
  return (a - b) ** 2 / (2 * np.pi) * s + dot * K, dot * T


def test_kernel_formulation_of_abrelu_matches_reference():
  """kd = hab2 + coef * copysign(pi/2 - sn G(|c|), c) with sn = s / sqrt(q1 q2), c = K / sqrt(q1 q2)
  (act_point / act_pair) is the reference's half_ab - coef * atan2(s, K)."""
  rng = np.random.default_rng(0)
  n = 200000
  q1, q2 = rng.uniform(0.05, 4, n), rng.uniform(0.05, 4, n)
  rho = np.concatenate([rng.uniform(-1, 1, n - 6), [1, -1, 0, 1 - 1e-9, -1 + 1e-9, 1e-12]])
  K = rho * np.sqrt(q1 * q2)
  T = rng.standard_normal(n)
  for a, b in ((0., 1.), (0.2, 1.), (-1., 1.)):
    coef, half_ab = (a - b) ** 2 / (2 * np.pi), (a * a + b * b) / 2
    hab2 = half_ab - coef * np.pi / 2
    p = q1 * q2
    s = np.sqrt(np.abs(p - K * K))
    rb = 1 / np.sqrt(p)
    sn, c = s * rb, K * rb
    u = np.pi / 2 - sn * G_exact(np.minimum(np.abs(c), 1))
    kd = hab2 + coef * np.copysign(u, c)
    Ko, To = kd * K + coef * s, kd * T
    Kr, Tr = _abrelu_reference(K, T, q1, q2, a, b)
    np.testing.assert_allclose(Ko, Kr, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(To, Tr, rtol=1e-9, atol=1e-9)
    # the carried form of the packed kernel: U' = kd * (T_conv) + K'  ==  T' + K'
    np.testing.assert_allclose(kd * T + Ko, Tr + Kr, rtol=1e-9, atol=1e-9)


def test_kernel_formulation_of_erf_matches_reference():
  """erf_act_point / act_pair_erf: with Kh = 2 b^2 K and D = 1 + 2 b^2 q the reference's
  (2/pi) atan2(2 b^2 K, sqrt(D1 D2 - 4 b^4 K^2)) is (2/pi) asin(Kh / sqrt(D1 D2)) via the same G."""
  rng = np.random.default_rng(1)
  n = 200000
  q1, q2 = rng.uniform(0.01, 5, n), rng.uniform(0.01, 5, n)
  K = rng.uniform(-1, 1, n) * np.sqrt(q1 * q2)
  T = rng.standard_normal(n)
  for a, b, c_ in ((1., 1., 0.), (0.8, 1.3, 0.2)):
    kb, tb = b * b * K, b * b * T                                   # k *= b  (Kernel.__mul__: b^2)
    prod = (1 + 2 * b * b * q1) * (1 + 2 * b * b * q2)
    sr = np.sqrt(np.maximum(prod - 4 * kb * kb, 0))
    Kr = a * a * (2 / np.pi) * np.arctan2(2 * kb, sr) + c_ * c_
    Tr = a * a * tb * (4 / np.pi) / sr
    e_in, eA, eT, eC = 2 * b * b, a * a * 2 / np.pi, a * a * b * b * 4 / np.pi, c_ * c_
    D1, D2 = 1 + e_in * q1, 1 + e_in * q2
    Kh = e_in * K
    d = D1 * D2 - Kh * Kh
    assert d.min() >= 1 - 1e-12                                     # no singularity: s >= 1
    rs, rb = 1 / np.sqrt(d), 1 / np.sqrt(D1 * D2)
    sn, cc = d * rs * rb, Kh * rb
    u = np.pi / 2 - sn * G_exact(np.minimum(np.abs(cc), 1))
    np.testing.assert_allclose(eA * np.copysign(u, cc) + eC, Kr, rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(eT * rs * T, Tr, rtol=1e-9, atol=1e-10)
