import csv,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
pairs=float(sys.argv[2]); layers_elems=float(sys.argv[3])
# find header rows (may repeat per kernel); use first kernel only
hdr=None; ops=collections.Counter(); samp=collections.Counter(); tot=0; nk=0
for r in rows:
    if r and r[0]=='Kernel Name':
        nk+=1
        if nk>1: break
        continue
    if r and r[0]=='Address': hdr=r; ia=hdr.index('Source'); ie=hdr.index('Instructions Executed'); isamp=hdr.index('# Samples'); continue
    if hdr is None or len(r)<=ie: continue
    src=r[ia].strip(); n=int(r[ie]); toks=src.split(); op=toks[0]
    if op.startswith('@'): op=toks[1]
    op=op.split('.')[0]
    ops[op]+=n; tot+=n; samp[op]+=int(r[isamp])
print('total warp instr',tot)
elems=pairs*layers_elems/32
for op,n in ops.most_common(28):
    print(f'{op:10s} {n:14d} {100*n/tot:6.2f}%  per elem-layer {n/elems:6.2f}   samples {samp[op]}')
print('instr per element-layer', tot/elems)
