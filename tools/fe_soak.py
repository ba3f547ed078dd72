"""CPU soak of the device front end under emulation (no GPU needed): tools/fuzz_frontend.py mail through the
frontend.cuh source (tests/emu) against the host front end, with and without foreign-signature skipping.

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
    for e in fz.build(n, seed):
        dom = e.from_domain.encode()
        for skip in (0, 1):
            r = lib.emu_fe_compare(e.raw_email, len(e.raw_email), dom, len(dom), 128, 32, skip)
            hist[(skip, r)] += 1
            if r < 0:
                bad += 1
                if bad < 6:
                    print("MISMATCH", r, skip, e.from_domain, e.raw_email[:500], file=sys.stderr)
    print(f"fe_soak seed {seed}: {dict(sorted(hist.items()))} mismatches {bad}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
