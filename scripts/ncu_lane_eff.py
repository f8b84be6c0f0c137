"""Share of samples and warp instructions per lanes-active bucket of an .ncu-rep (source page)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
b = {}
ts=ti=0
for r in csv.reader(io.StringIO(src)):
    if len(r) < 12 or r[0] in ("Line No", ""): continue
    try: ln = int(r[0]); s = int(r[4]); i = int(r[7]); t = int(r[8])
    except ValueError: continue
    if i == 0: continue
    e = t / i / 32
    k = round(e * 10) / 10
    x = b.setdefault(k, [0, 0, 0]); x[0] += s; x[1] += i; x[2] += t
    ts += s; ti += i
for k in sorted(b): print(f"eff~{k:.1f}: samples {100*b[k][0]/ts:5.1f}%  warp-inst {100*b[k][1]/ti:5.1f}%  thread-inst {b[k][2]/1e9:.2f}G")
