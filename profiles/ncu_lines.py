"""Per-source-line instruction / stall-sample shares of one kernel from an .ncu-rep captured with --import-source on.
usage: python profiles/ncu_lines.py report.ncu-rep kernel_regex [top_n]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 60
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{rx}",
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = None; out = []; cur = None
num = lambda x: int(x) if x.strip().lstrip('-').isdigit() else 0
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) > 8 and r[0].isdigit(): out.append((cur, r))
ie = hdr.index('Instructions Executed'); ss = hdr.index('# Samples'); te = hdr.index('Thread Instructions Executed')
tot = sum(num(r[ie]) for f, r in out); tots = sum(num(r[ss]) for f, r in out)
print('total warp instructions', tot, 'samples', tots, 'source lines', len(out))
top = sorted(out, key=lambda x: -max(num(x[1][ie]) / max(tot, 1), num(x[1][ss]) / max(tots, 1)))[:top_n]
for f, r in sorted(top, key=lambda x: (x[0], int(x[1][0]))):
    i = num(r[ie]); t = num(r[te])
    print(f"{f[-18:]:18s} L{r[0]:>5s} inst {i / tot * 100:5.1f}% samp {num(r[ss]) / tots * 100:5.1f}% thr/inst {t / max(i, 1):4.1f}  {r[1].strip()[:100]}")
