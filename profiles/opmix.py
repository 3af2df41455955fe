"""Opcode mix (weighted by executions) of one kernel from an .ncu-rep with source: python profiles/opmix.py rep kernel"""
import csv, collections, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--kernel-name", sys.argv[2], "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
h = rows[hi]; ie = h.index('Instructions Executed'); isrc = h.index('Source')
agg = collections.Counter(); tot = 0; static = 0
for r in rows[hi + 1:]:
    try: n = int(r[ie])
    except Exception: continue
    s = r[isrc].strip()
    if s.startswith('@'): s = s.split(None, 1)[1]
    agg[s.split()[0].split('.')[0]] += n; tot += n; static += 1
print(f"# {sys.argv[2]}: {tot} warp-instructions executed, {static} SASS instructions")
for k, v in agg.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 30): print(f"{k:12s} {v:11d} {100*v/tot:5.1f}%")
