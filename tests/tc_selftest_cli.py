"""Runs device-vs-device engine cross-checks (sfno_b200_selftest_gemm) in their own process so that a trap in a
tensor-core kernel cannot poison the CUDA context of the test session.  One case from the command line (op, dims...;
SFNO_TC_DEBUG from the environment) or, with ``--batch``, a JSON list of ``[op, dims, tc_debug]`` cases from stdin (one
CUDA context for all of them).  Prints one JSON line per case."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "spherical-dyffusion_b200", "libsfno_b200.so")


def load():
    lib = ctypes.CDLL(os.environ.get("SFNO_B200_LIB") or LIB)
    lib.sfno_b200_selftest_gemm.restype = ctypes.c_int
    lib.sfno_b200_selftest_gemm.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    lib.sfno_b200_last_error.restype = ctypes.c_char_p
    lib.sfno_b200_set_option.argtypes = [ctypes.c_char_p, ctypes.c_int64]
    return lib


def run_case(lib, op, dims, tc_debug):
    dims = list(dims) + [0] * (6 - len(dims))
    lib.sfno_b200_set_option(b"tc_debug", int(tc_debug))
    res = (ctypes.c_double * 5)()
    arr = (ctypes.c_int * len(dims))(*dims)
    st = lib.sfno_b200_selftest_gemm(op, arr, len(dims), res)
    counters = None
    if int(tc_debug) & 128:
        buf = (ctypes.c_ulonglong * 12)()
        lib.sfno_b200_tc_counters(buf)
        c = list(buf)
        ctas = max(c[7], 1)
        life = max(c[5], 1)
        counters = {"producer_wait_stage": c[0] / life, "mma_wait_operands": c[1] / life, "mma_wait_accumulator": c[2] / life,
                    "epi_wait_accumulator": c[3] / life, "epi_wait_residual": c[4] / life, "epi_busy": c[6] / life,
                    "epi_setup": c[8] / life, "epi_drain": c[9] / life, "epi_store": c[10] / life,
                    "cta_cycles": life / ctas, "ctas": c[7]}
    return {"status": st, "error": lib.sfno_b200_last_error().decode() if st else "", "max_err": res[0], "max_ref": res[1],
            "tc_used": res[2], "ms": res[3], "nonfinite": res[4], "op": op, "dims": dims, "tc_debug": int(tc_debug), "counters": counters}


def main():
    lib = load()
    if len(sys.argv) > 1 and sys.argv[1] == "--batch":
        for op, dims, tc_debug in json.loads(sys.stdin.read()):
            print(json.dumps(run_case(lib, int(op), [int(v) for v in dims], int(tc_debug))), flush=True)
        return 0
    out = run_case(lib, int(sys.argv[1]), [int(v) for v in sys.argv[2:]], int(os.environ.get("SFNO_TC_DEBUG", "0") or 0))
    print(json.dumps(out))
    return 0 if out["status"] == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
