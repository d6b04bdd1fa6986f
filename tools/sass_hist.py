"""Instruction histogram of an emitted kernel (NVRTC -> cubin -> cuobjdump), no GPU needed.
usage: python tools/sass_hist.py build/emit_nvrtc_extra_ordinary_wave_efit_rk4.cu [options]"""
import collections, ctypes, re, subprocess, sys, tempfile
from graph_framework_b200._lib import lib

src = open(sys.argv[1]).read().encode()
opts = sys.argv[2].encode() if len(sys.argv) > 2 else None
cubin, size, log = ctypes.c_void_p(), ctypes.c_size_t(), ctypes.c_void_p()
rc = lib.gfb_compile_to_cubin(src, opts, ctypes.byref(cubin), ctypes.byref(size), ctypes.byref(log))
assert rc == 0, ctypes.string_at(log).decode()
with tempfile.NamedTemporaryFile(suffix=".cubin") as f:
    f.write(ctypes.string_at(cubin, size.value)); f.flush()
    sass = subprocess.run(["cuobjdump", "-sass", f.name], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", f.name], capture_output=True, text=True).stdout
print(res.strip().splitlines()[-1])
hist = collections.Counter()
for line in sass.splitlines():
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m:
        hist[m.group(2).split(".")[0] + ("." + m.group(2).split(".")[1] if m.group(2).startswith(("MUFU", "F2I", "I2F", "F2F", "LD", "ST")) and "." in m.group(2) else "")] += 1
total = sum(hist.values())
print("total", total)
for k, v in hist.most_common(60):
    print("%-14s %6d %5.1f%%" % (k, v, 100.0*v/total))
