#!/usr/bin/env python
"""Summarise an ncu report per CUDA source line: python scripts/ncu_lines.py report.ncu-rep [top] [samp|instr]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
key = 4 if (len(sys.argv) > 3 and sys.argv[3] == "instr") else 3
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None; fname = ""; out = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0] not in ("", "Function Name") and len(r) >= 12:
        try:
            out.append((fname, int(r[0]), r[1].strip(), float(r[6] or 0), float(r[7] or 0), float(r[8] or 0)))
        except ValueError:
            pass
ts = sum(o[3] for o in out); ti = sum(o[4] for o in out); tt = sum(o[5] for o in out)
print(f"samples {ts:.0f}  warp-instr {ti:.3g}  thread-instr {tt:.3g}  avg threads/instr {tt/ti:.2f}")
for o in sorted(out, key=lambda o: -o[key])[:top]:
    print(f"{o[0][:16]:16s}{o[1]:5d} samp {100*o[3]/ts:5.1f}% instr {100*o[4]/ti:5.1f}% thr/instr {o[5]/max(o[4],1):5.1f} | {o[2][:100]}")
