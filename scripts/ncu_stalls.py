#!/usr/bin/env python
"""Stall reasons of an ncu report, whole kernel and per CUDA source line: python scripts/ncu_stalls.py report.ncu-rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None; fname = ""; out = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r
        st = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
        continue
    if hdr and r[0] not in ("", "Function Name") and len(r) >= 12:
        try:
            out.append((fname, int(r[0]), r[1].strip(), float(r[6] or 0), float(r[7] or 0), float(r[8] or 0), [float(r[i] or 0) for i in st]))
        except ValueError:
            pass
names = [hdr[i][6:] for i in st]
ts = sum(o[3] for o in out); ti = sum(o[4] for o in out); tt = sum(o[5] for o in out)
tot = [sum(o[6][k] for o in out) for k in range(len(st))]
print(f"samples {ts:.0f}  warp-instr {ti:.3g}  thread-instr {tt:.3g}  avg threads/instr {tt/ti:.2f}")
print("stall reasons: " + "  ".join(f"{n} {100*v/ts:.1f}%" for n, v in sorted(zip(names, tot), key=lambda x: -x[1]) if v > 0.005 * ts))
for o in sorted(out, key=lambda o: -o[3])[:top]:
    big = sorted(zip(names, o[6]), key=lambda x: -x[1])[:3]
    print(f"{o[0][:14]:14s}{o[1]:5d} samp {100*o[3]/ts:5.1f}% instr {100*o[4]/ti:5.1f}% thr {o[5]/max(o[4],1):5.1f} {' '.join(f'{n}:{100*v/max(o[3],1):.0f}' for n, v in big):40s}| {o[2][:80]}")
