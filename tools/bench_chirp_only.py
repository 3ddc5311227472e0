#!/usr/bin/env python
"""Chirp rows of tools/bench_modes.py only (kernel experiments)."""
import os, sys, re, io, contextlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
src = open(os.path.join(ROOT, "tools", "bench_modes.py")).read()
# run the module but keep only chirp lines
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    exec(compile(src, "bench_modes.py", "exec"), {"__name__": "__main__", "__file__": os.path.join(ROOT, "tools", "bench_modes.py")})
for line in buf.getvalue().splitlines():
    if "CHIRP" in line:
        print(line[:330])
