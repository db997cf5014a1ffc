"""Register-file source words per element-layer of a SASS loop body (reuse-cache hits excluded).

  cuobjdump -sass -fun <kernel> lib.o | python profiles/rf_words.py <element-layers per loop trip>

Model (profiles/microbench/rf_bandwidth.cu): every SMSP reads 2 32-bit register words per lane and
clock, shared by all pipes; immediates, uniform registers and constants are free; an operand whose
slot was flagged .reuse by the previous instruction comes from the operand-reuse cache.
"""
import collections
import re
import sys

lines = [l for l in sys.stdin if re.search(r'/\*[0-9a-f]{4,5}\*/', l)]
text = [re.sub(r'/\* 0x[0-9a-f]* \*/', '', l) for l in lines]
# innermost big loop: last backward uniform branch
back = [i for i, l in enumerate(text) if re.search(r'BRA(\.U)?\s+!?U?P\d, 0x', l) or ' BRA 0x' in l]
best = None
for i in back:
  m = re.search(r'0x([0-9a-f]+)', text[i].split('BRA')[1])
  tgt = m.group(1).rjust(4, '0')
  for j in range(i):
    if re.search(r'/\*0*%s\*/' % tgt.lstrip('0'), text[j]):
      if best is None or i - j > best[1] - best[0]:
        best = (j, i)
body = text[best[0]:best[1] + 1]
per = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
words = collections.Counter()
cnt = collections.Counter()
prev_reuse = {}
tot = hits = 0
for line in body:
  m = re.search(r'\*/\s+(@!?U?P\d\s+)?(\S+)\s+(.*);', line)
  if not m:
    continue
  opfull = m.group(2)
  op = opfull.split('.')[0]
  ops = [a.strip() for a in m.group(3).split(',')]
  srcs = ops if op in ('STS', 'STG', 'REDG', 'RED', 'BAR', 'BRA') else ops[1:]
  cur_reuse = {}
  seen = set()
  for slot, a in enumerate(srcs):
    mm = re.search(r'(?<![U\w])R(\d+)', a)
    if not mm or a.startswith('RZ'):
      continue
    r = mm.group(1)
    wide = 2 if 'F32x2' in a else 1
    if op in ('STS', 'STG') and not a.startswith('['):
      wide = 4 if '.128' in opfull else 2 if '.64' in opfull else 1
    if '.reuse' in a:
      cur_reuse[slot] = r
    if prev_reuse.get(slot) == r:
      hits += wide
      continue
    if r in seen:
      continue
    seen.add(r)
    words[op] += wide
    tot += wide
  cnt[op] += 1
  prev_reuse = cur_reuse
n = sum(cnt.values())
print(f'loop body {n} instructions = {n / per:.2f} per element-layer')
print(f'register source words per element-layer {tot / per:.2f} (reuse-cache hits {hits / per:.2f}) '
      f'-> RF floor {tot / per / 2:.1f} clk')
for op, w in words.most_common(12):
  print(f'  {op:8s} x{cnt[op]:4d}  {w / per:6.2f} words')
