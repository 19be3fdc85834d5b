"""Phase-level stall table of k_pair4095_tmem from the per-SASS-instruction page of an ncu report
(ncu -i x.ncu-rep --page source --csv --print-source sass > x.csv).  The phase boundaries below are the SASS line numbers
of the build that produced profiles/r2i_pair_sass_stalls.csv.gz (found from the STTM / LDTM runs and the loop head).

    python tools/ncu_stalls.py [profiles/r2i_pair_sass_stalls.csv.gz]
"""
import collections
import csv
import gzip
import io
import sys

path = sys.argv[1] if len(sys.argv) > 1 else 'profiles/r2i_pair_sass_stalls.csv.gz'
raw = gzip.open(path, 'rt').read() if path.endswith('.gz') else open(path).read()
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; data = [r for r in rows[2:] if len(r)>=len(hdr)]; ix = {h:i for i,h in enumerate(hdr)}
sc = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
# loop body starts where? find first index with Instructions Executed ~2031616
first = next(i for i,r in enumerate(data) if int(r[ix['Instructions Executed']] or 0) > 2000000)
print('loop starts at', first)
phases = [('prologue', 0, first), ('A1: PHAT + DFT-5 -> TMEM (spectrum loads)', first, 964), ('A2: DFT-13 -> exchange tile', 964, 1640),
          ('B1: Hermitian merge + DFT-7 -> TMEM', 1640, 2244), ('B2: DFT-9 -> natural-order row scatter', 2244, 2957), ('odd column (7 lanes) DFT-9', 2957, 3055),
          ('fetch next pair (indices, bound, 30 loads)', 3055, 3165), ('peak pick + stores', 3165, len(data))]
tot = sum(int(r[ix['# Samples']] or 0) for r in data)
print('| phase | SASS instr | warp-inst/pair | samples % | top stall reasons (% of phase samples) |')
print('|---|---:|---:|---:|---|')
for name,a,b in phases:
    s = 0; ie = 0; st = collections.Counter()
    for r in data[a:b]:
        s += int(r[ix['# Samples']] or 0); ie += int(r[ix['Instructions Executed']] or 0)
        for c in sc: st[c[6:]] += int(r[ix[c]] or 0)
    top = ', '.join(f'{k} {100*v/max(s,1):.0f}' for k,v in st.most_common(5))
    print(f'| {name} | {b-a} | {ie/2031616:.0f} | {100*s/tot:.1f} | {top} |')
