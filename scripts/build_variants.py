#!/usr/bin/env python
"""A/B builds of the CUDA library: python scripts/build_variants.py NAME=-DFOO=1,-DBAR=2 ...  -> pyafv_b200/ab/libafv_NAME.so
(select one at run time with AFV_B200_LIB=pyafv_b200/ab/libafv_NAME.so)."""
import subprocess, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pyafv_b200 import _build
out = _build.PKG / "ab"
out.mkdir(exist_ok=True)
procs = []
for spec in sys.argv[1:]:
    name, _, flags = spec.partition("=")
    cmd = [_build.nvcc_path(), *_build.NVCC_FLAGS, *[f for f in flags.split(",") if f], "-o", str(out / f"libafv_{name}.so"), str(_build.SRC)]
    procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for name, p in procs:
    o, _ = p.communicate()
    print(name, "ok" if p.returncode == 0 else "FAILED\n" + o)
