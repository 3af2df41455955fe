"""Per-source-line instruction / stall-sample totals of one kernel from an .ncu-rep captured with --import-source on.
   python profiles/srcprof.py <rep> <kernel-name> [top]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern, "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = [r for r in rows if r and r[0] == 'Line No'][0]
n = len(hdr); ie = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples')
cur = None; res = []
srcs = {}
for r in rows:
    if r and r[0] == 'File Path': cur = r[1].split('/')[-1]
    if r and r[0].isdigit():
        off = len(r) - n
        try: inst = int(r[ie + off]); s = int(r[isamp + off])
        except Exception: continue
        res.append((cur, int(r[0]), inst, s, ",".join(r[1:2 + off])[:90]))
tot = sum(o[2] for o in res); ts = sum(o[3] for o in res)
print(f"# {kern}: {tot} warp-instructions attributed, {ts} stall samples")
for o in sorted(res, key=lambda o: -o[2])[:top]:
    print(f"{o[0]:20s} {o[1]:4d} inst {o[2]:10d} {100*o[2]/tot:5.1f}%  samples {o[3]:6d} {100*o[3]/max(ts,1):5.1f}%  | {o[4].strip()}")
