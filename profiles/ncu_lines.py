"""Per-source-line hot spots of one kernel in an .ncu-rep (needs -lineinfo + --import-source on).
usage: python profiles/ncu_lines.py <report.ncu-rep> <kernel regex> [top N]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern,
                      "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, lines = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit():
        if len(r) > len(hdr):                      # source text with unescaped quotes / commas
            extra = len(r) - len(hdr)
            r = [r[0], ",".join(r[1:2 + extra])] + r[2 + extra:]
        d = dict(zip(hdr, [x if x not in ("-", "") else "0" for x in r]))
        # two "Source" columns: first is the CUDA line
        lines.append((fname, int(r[0]), r[1].strip(), int(d.get("# Samples", 0) or 0), int(d.get("Instructions Executed", 0) or 0),
                      int(d.get("Thread Instructions Executed", 0) or 0) / max(int(d.get("Instructions Executed", 0) or 0), 1)))
ti = sum(l[4] for l in lines) or 1
ts = sum(l[3] for l in lines) or 1
print(f"total instructions {ti:.3e}, samples {ts}")
print("by instructions executed:")
for l in sorted(lines, key=lambda x: -x[4])[:top]:
    print(f"{100*l[4]/ti:5.1f}% inst {100*l[3]/ts:5.1f}% smp  thr {l[5]:4.1f}  {l[0]}:{l[1]:<5d} {l[2][:110]}")
if len(sys.argv) > 4:
    print("by line range (file sphb_stages.cuh):")
    for rng in sys.argv[4].split(","):
        a, b = map(int, rng.split("-"))
        sel = [l for l in lines if l[0] == "sphb_stages.cuh" and a <= l[1] <= b]
        print(f"  {rng}: {100*sum(l[4] for l in sel)/ti:5.1f}% inst {100*sum(l[3] for l in sel)/ts:5.1f}% smp")
    oth = [l for l in lines if l[0] != "sphb_stages.cuh"]
    for f in sorted(set(l[0] for l in oth)):
        sel = [l for l in oth if l[0] == f]
        print(f"  {f}: {100*sum(l[4] for l in sel)/ti:5.1f}% inst {100*sum(l[3] for l in sel)/ts:5.1f}% smp")
