"""Runs one device-vs-device engine cross-check (sfno_b200_selftest_gemm) in its own process so that a trap in a
tensor-core kernel cannot poison the CUDA context of the test session.  Prints one JSON line."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "spherical-dyffusion_b200", "libsfno_b200.so")


def main():
    op = int(sys.argv[1])
    dims = [int(v) for v in sys.argv[2:]]
    dims += [0] * (6 - len(dims))
    lib = ctypes.CDLL(LIB)
    lib.sfno_b200_selftest_gemm.restype = ctypes.c_int
    lib.sfno_b200_selftest_gemm.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    lib.sfno_b200_last_error.restype = ctypes.c_char_p
    if os.environ.get("SFNO_TC_DEBUG"):
        lib.sfno_b200_set_option.argtypes = [ctypes.c_char_p, ctypes.c_int64]
        lib.sfno_b200_set_option(b"tc_debug", int(os.environ["SFNO_TC_DEBUG"]))
    res = (ctypes.c_double * 5)()
    arr = (ctypes.c_int * len(dims))(*dims)
    st = lib.sfno_b200_selftest_gemm(op, arr, len(dims), res)
    counters = None
    if int(os.environ.get("SFNO_TC_DEBUG", "0")) & 128:
        buf = (ctypes.c_ulonglong * 12)()
        lib.sfno_b200_tc_counters(buf)
        c = list(buf)
        ctas = max(c[7], 1)
        life = max(c[5], 1)
        counters = {"producer_wait_stage": c[0] / life, "mma_wait_operands": c[1] / life, "mma_wait_accumulator": c[2] / life,
                    "epi_wait_accumulator": c[3] / life, "epi_wait_residual": c[4] / life, "epi_busy": c[6] / life,
                    "epi_setup": c[8] / life, "epi_drain": c[9] / life, "epi_store": c[10] / life,
                    "cta_cycles": life / ctas, "ctas": c[7]}
    out = {"status": st, "error": lib.sfno_b200_last_error().decode() if st else "", "max_err": res[0], "max_ref": res[1],
           "tc_used": res[2], "ms": res[3], "nonfinite": res[4], "op": op, "dims": dims, "counters": counters}
    print(json.dumps(out))
    return 0 if st == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
