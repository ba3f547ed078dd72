"""CPU soak of the device front end under emulation (no GPU needed): tools/fuzz_frontend.py mail through the
frontend.cuh / frontend_warp.cuh sources (tests/emu) against the host front end, with and without foreign-signature skipping.

    python tools/fe_soak.py [seed] [n_emails]

A negative code from emu_fe_compare is a mismatch between what the device accepted and what the host front end
produces for the same message; 0 = the device declined (host path), 1 = accepted and byte-identical, 2 = both report a
mail parse error.  Round 1: 180 000 comparisons over three seeds, no mismatch."""
import collections
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 30000
    spec = importlib.util.spec_from_file_location("fuzz_frontend", os.path.join(HERE, "fuzz_frontend.py"))
    fz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fz)
    from tests import emu
    lib = emu.lib()
    hist, bad = collections.Counter(), 0
    import ctypes as C
    whist = collections.Counter()
    for e in fz.build(n, seed):
        dom = e.from_domain.encode()
        for skip in (0, 1):
            sc = C.c_int(-99)
            # the warp-cooperative kernel source (32 emulated lanes) and, through it, the scalar twin
            rw = lib.emu_fe_compare_warp(e.raw_email, len(e.raw_email), dom, len(dom), 128, 32, skip, C.byref(sc))
            r = sc.value
            hist[(skip, r)] += 1
            whist[(skip, rw)] += 1
            if r < 0 or rw < 0 or (rw == 1 and r == 2) or (rw == 2 and r == 1):
                bad += 1
                if bad < 6:
                    print("MISMATCH scalar", r, "warp", rw, skip, e.from_domain, e.raw_email[:500], file=sys.stderr)
    print(f"fe_soak seed {seed}: scalar {dict(sorted(hist.items()))} warp {dict(sorted(whist.items()))} mismatches {bad}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
