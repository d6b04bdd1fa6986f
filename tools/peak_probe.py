#!/usr/bin/env python3
"""Runs the FP64 peak probe (gfb_measure_fp64_peak) alone, for `ncu --kernel-name regex:fp64_peak`."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_framework_b200._lib import lib, check

ctx = lib.gfb_ctx_create(0)
t, ms = ctypes.c_double(0), ctypes.c_float(0)
check(lib.gfb_measure_fp64_peak(ctx, ctypes.byref(t), ctypes.byref(ms)), "peak")
print("fp64 peak, uniform multiplier/addend: %.2f TFLOP/s (%.3f ms)" % (t.value, ms.value))
check(lib.gfb_measure_fp64_peak_registers(ctx, ctypes.byref(t), ctypes.byref(ms)), "peak regs")
print("fp64 peak, three register operands:    %.2f TFLOP/s (%.3f ms)" % (t.value, ms.value))
lib.gfb_ctx_destroy(ctx)
