#!/usr/bin/env python
"""Per-region stall-sample attribution of a kernel from an .ncu-rep (regions split at BAR.SYNC)."""
import csv, subprocess, sys
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
isrc = hdr.index('Source'); iss = hdr.index('Warp Stall Sampling (All Samples)'); iex = hdr.index('Instructions Executed')
data = [(int(r[iss] or 0), r[isrc].strip(), int(r[iex] or 0), k) for k, r in enumerate(rows[2:])]
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
for d in sorted(data, reverse=True)[:topn]:
    print(f"{d[0]:7d} {100*d[0]/tot:5.1f}%  idx={d[3]:5d} exec={d[2]:9d}  {d[1][:90]}")
bars = [d[3] for d in data if 'BAR.SYNC' in d[1]]
prev = 0
for b in bars + [len(data)]:
    seg = data[prev:b + 1]
    ssum = sum(x[0] for x in seg); esum = sum(x[2] for x in seg)
    f64 = sum(x[2] for x in seg if x[1].split()[0] in ('DFMA', 'DMUL', 'DADD') or (len(x[1].split()) > 1 and x[1].split()[1] in ('DFMA', 'DMUL', 'DADD')))
    lds = sum(x[2] for x in seg if 'LDS' in x[1]); sts = sum(x[2] for x in seg if 'STS' in x[1]); ldg = sum(x[2] for x in seg if 'LDG' in x[1]); red = sum(x[2] for x in seg if 'RED' in x[1] or 'ATOM' in x[1])
    print(f"region {prev:5d}-{b:5d}: samples {ssum:7d} ({100*ssum/tot:5.1f}%) warp-instr {esum:11d} fp64 {f64:11d} lds {lds:10d} sts {sts:10d} ldg {ldg:10d} red {red:10d}")
    prev = b + 1
