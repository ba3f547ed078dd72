"""Differential fuzz of the regex compiler (csrc/regexc.hpp through zkb_regex_compile) and the DFA search semantics against
Python `re` in bytes mode: random patterns from a small grammar (literals, classes, alternation, greedy / lazy / bounded
repetition, groups, the `(\\r\\n|^)` line prefix zk-email patterns use), random haystacks over a matching alphabet.
Patterns that can match the empty string are skipped (Rust's and Python's iteration differ there).  CPU only.

    python tools/fuzz_regex.py [seed] [n_patterns] [--emu | --gpu]

--emu also runs the DFA scan KERNEL SOURCE (csrc/dfa.cuh under host emulation, tests/emu) over every haystack and holds its
match count and first span to the oracle's: random tables with anchors, look-behind start states and word boundaries.
--gpu does the same with the real kernel through zkb_dfa_scan_batch (needs a B200), with and without soft-break removal.
"""
import os
import random
import re
import signal
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402  (test infrastructure: the checker)
import zkemail_rs_b200 as z  # noqa: E402

ALPHA = "abcxyz019 .:@\r\n"
CLASSES = ["[a-c]", "[^a-c\r\n]", "[0-9]", "[a-z0-9]", r"\d", r"[^\r\n]", "[xyz@.]", "."]


def gen_atom(rng, depth):
    """-> (pattern for the compiler under test, the same pattern in Python's spelling)"""
    r = rng.random()
    if r < 0.42:
        c = rng.choice(ALPHA)
        t = {"\r": "\\r", "\n": "\\n", ".": "\\.", " ": " "}.get(c, c)
        return t, t
    if r < 0.70:
        t = rng.choice(CLASSES)
        return t, t
    if r < 0.76:
        # zero-width assertions: Rust's `$` is the end of the haystack only (Python: \Z); ASCII word boundaries need
        # the (?-u:..) spelling in Rust; (?m) anchors mean the same in both
        return rng.choice([("$", "\\Z"), ("(?-u:\\b)", "\\b"), ("(?-u:\\B)", "\\B"), ("(?m:^)", "(?m:^)"), ("(?m:$)", "(?m:$)"), ("^", "^")])
    if depth < 2:
        a, b = gen_alt(rng, depth + 1)
        return "(" + a + ")", "(" + b + ")"
    t = rng.choice(CLASSES)
    return t, t


def gen_piece(rng, depth):
    a, b = gen_atom(rng, depth)
    if rng.random() < 0.55 or a in ("$", "^") or a.startswith("(?"):
        return a, b
    q = rng.choice(["+", "*", "?", "{2}", "{1,3}", "{2,}", "+?", "*?", "??", "{1,2}?"])
    return a + q, b + q


def gen_concat(rng, depth):
    ps = [gen_piece(rng, depth) for _ in range(rng.randint(1, 3))]
    return "".join(p[0] for p in ps), "".join(p[1] for p in ps)


def gen_alt(rng, depth):
    cs = [gen_concat(rng, depth) for _ in range(1 if rng.random() < 0.7 else rng.randint(2, 3))]
    return "|".join(c[0] for c in cs), "|".join(c[1] for c in cs)


def gen_pattern(rng):
    p, q = gen_alt(rng, 0)
    if rng.random() < 0.15:
        p, q = r"(\r\n|^)" + p, r"(\r\n|^)" + q
    if rng.random() < 0.1:
        p, q = "(?i)" + p, "(?i)" + q
    if rng.random() < 0.1:
        p, q = "(?s)" + p, "(?s)" + q
    return p, q


def _alarm(signum, frame):
    raise TimeoutError


def has_nullable_loop(pat):
    """True if some repetition that may run more than once has a body that can match the empty string.  On those a
    backtracking engine (an empty iteration ends the loop) and the Thompson construction (closure order) can disagree
    after a non-empty iteration; DESIGN.md section 5 and tests/golden/regex_pin_extra.json."""
    try:
        import re._parser as sp        # Python >= 3.11
    except ImportError:                 # pragma: no cover
        import sre_parse as sp

    def walk(items):
        for op, av in items:
            name = str(op)
            if name in ("MAX_REPEAT", "MIN_REPEAT", "POSSESSIVE_REPEAT"):
                lo, hi, body = av
                if hi > 1 and body.getwidth()[0] == 0:
                    return True
                if walk(body):
                    return True
            elif name == "SUBPATTERN":
                if walk(av[3]):
                    return True
            elif name == "BRANCH":
                if any(walk(b) for b in av[1]):
                    return True
        return False
    return walk(sp.parse(pat))


def main():
    signal.signal(signal.SIGALRM, _alarm)
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    with_gpu = "--gpu" in sys.argv
    with_emu = "--emu" in sys.argv or with_gpu
    if with_gpu:
        eng = z.Engine()

        class emu:    # same call shape as tests.emu
            @staticmethod
            def dfa_scan(fwd, bwd, hays, qp=False):
                return eng.dfa_scan_batch(z.DFA(fwd, bwd), hays, qp=qp)
    elif with_emu:
        from tests import emu
    seed = int(args[0]) if len(args) > 0 else 1
    n = int(args[1]) if len(args) > 1 else 2000
    kernel_bad = 0
    rng = random.Random(seed)
    tested = skipped = bad = nullable = 0
    for _ in range(n):
        pat, pypat = gen_pattern(rng)
        try:
            py = re.compile(pypat.encode())
        except re.error:
            skipped += 1
            continue
        nullable_pattern = py.search(b"") is not None     # can match the empty string: no comparison with Python
        if nullable_pattern and not with_emu:
            skipped += 1
            continue
        try:
            d = z.compile_regex(pat)
        except Exception as ex:   # noqa: BLE001
            if "too large" in str(ex):
                skipped += 1        # the compiler's state limit, a documented refusal
                continue
            print("COMPILE FAILED", repr(pat), ex, file=sys.stderr)
            bad += 1
            continue
        tested += 1
        if with_emu:
            hays = ["".join(rng.choice(ALPHA) for _ in range(rng.randint(0, 80))).encode() for _ in range(16)]
            if with_gpu and rng.random() < 0.3:       # quoted-printable soft breaks removed on the fly (core/src/email.rs:61-86)
                hays = [h.replace(b"y", b"=\r\n") for h in hays]
                rows = emu.dfa_scan(d.fwd, d.bwd, hays, qp=True)
                hays = [oracle.qp_clean(h)[0] for h in hays]
            else:
                rows = emu.dfa_scan(d.fwd, d.bwd, hays)
            for hay, row in zip(hays, rows):
                cnt, spans = oracle.dfa_find_iter(d.fwd, d.bwd, hay)
                first = tuple(spans[0]) if cnt else (0, 0)
                if int(row[0]) != cnt or (cnt and (int(row[1]), int(row[2])) != first):
                    kernel_bad += 1
                    if kernel_bad < 6:
                        print("KERNEL MISMATCH", repr(pat), hay, "oracle", cnt, first, "kernel", [int(x) for x in row], file=sys.stderr)
        if nullable_pattern:
            skipped += 1            # (the kernel source was still held to the oracle above)
            continue
        for _h in range(12):
            hay = "".join(rng.choice(ALPHA) for _ in range(rng.randint(0, 60))).encode()
            try:                     # Python's backtracking matcher can take exponential time on nested repetitions
                signal.alarm(2)
                ms = list(py.finditer(hay))
                signal.alarm(0)
            except TimeoutError:
                break
            if any(m.start() == m.end() for m in ms):
                continue            # an empty match somewhere: iteration rules differ
            want = [(m.start(), m.end()) for m in ms]
            cnt, spans = oracle.dfa_find_iter(d.fwd, d.bwd, hay)
            got = [tuple(s) for s in spans[:cnt]]
            if any(a == b for a, b in got):
                continue            # an empty match on the automaton's side (e.g. \B on an empty haystack, which Python refuses)
            if cnt != len(want) or got != want[:len(got)]:
                if has_nullable_loop(pypat):
                    nullable += 1   # the one class where the two families of engines differ by construction
                    break
                bad += 1
                if bad < 8:
                    print("MISMATCH", repr(pat), hay, "python", want, "dfa", cnt, got, file=sys.stderr)
                break
    print(f"fuzz_regex seed {seed}: {tested} patterns tested, {skipped} skipped, mismatches {bad}, "
          f"disagreements inside loops with a nullable body {nullable}" + (f", kernel-source mismatches {kernel_bad}" if with_emu else ""))
    return 1 if bad or kernel_bad else 0


if __name__ == "__main__":
    sys.exit(main())
