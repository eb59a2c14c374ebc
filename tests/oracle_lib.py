"""TEST INFRASTRUCTURE: ctypes binding of oracle/_ref/liboracle_gpsl1.so (the
plain-C restatement of the reference sample loop). Imported only by tests,
__graft_entry__.smoke() and bench.py's cpu_baseline leg -- never by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_ref", "liboracle_gpsl1.so")


class OracleChan(C.Structure):
    _fields_ = [("prn", C.c_int32), ("iword", C.c_int32), ("ibit", C.c_int32), ("icode", C.c_int32),
                ("f_carr", C.c_double), ("f_code", C.c_double), ("carr_phase", C.c_double),
                ("code_phase", C.c_double), ("gain", C.c_double), ("dwrd", C.c_uint32 * 60)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "_ref/liboracle_gpsl1.so"])
        _lib = C.CDLL(SO)
        _lib.oracle_codegen.argtypes = [C.c_int, C.c_void_p]
        _lib.oracle_codegen.restype = C.c_int
        _lib.oracle_tables.argtypes = [C.c_void_p, C.c_void_p]
        _lib.oracle_synth_block.argtypes = [C.POINTER(OracleChan), C.c_int, C.c_int, C.c_void_p]
        _lib.oracle_quantize8.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    return _lib


def codegen(prn):
    ca = np.zeros(1023, np.uint8)
    assert lib().oracle_codegen(prn, ca.ctypes.data) == 0
    return ca


def tables():
    s, c = np.zeros(512, np.int32), np.zeros(512, np.int32)
    lib().oracle_tables(s.ctypes.data, c.ctypes.data)
    return s, c


def make_chans(rec_row, nav_row, carr_phase=None):
    """rec_row: refdump.CHAN_DT[C] (one block); nav_row: uint32[C, 60]."""
    n = len(rec_row)
    arr = (OracleChan * n)()
    for i in range(n):
        r = rec_row[i]
        a = arr[i]
        a.prn, a.iword, a.ibit, a.icode = int(r["prn"]), int(r["iword"]), int(r["ibit"]), int(r["icode"])
        a.f_carr, a.f_code = float(r["f_carr"]), float(r["f_code"])
        a.carr_phase = float(r["carr_phase"] if carr_phase is None else carr_phase[i])
        a.code_phase, a.gain = float(r["code_phase"]), float(r["gain"])
        for k in range(60):
            a.dwrd[k] = int(nav_row[i][k])
    return arr


def synth_block(chans, nsamp=300000):
    iq = np.zeros(2 * nsamp, np.int16)
    lib().oracle_synth_block(chans, len(chans), nsamp, iq.ctypes.data)
    return iq


def quantize8(iq16):
    out = np.zeros(iq16.size, np.int8)
    lib().oracle_quantize8(iq16.ctypes.data, iq16.size, out.ctypes.data)
    return out
