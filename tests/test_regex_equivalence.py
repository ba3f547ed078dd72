"""Second opinion on the regex compiler (besides Python `re` and the real regex-automata blobs): exhaustive language
equivalence, over the ASCII alphabet, between the tables csrc/regexc.hpp produces and the finite-state machine the
independent `interegular` package builds for the same pattern.

  * reverse table (`DFA.bwd`: anchored, match kind "all" — helpers/src/regex.rs:7-14 via dfa::regex::Builder) must accept
    exactly reverse(L(pattern));
  * forward table (anchored start, leftmost-first) must accept a subset of L(pattern) that contains, for every string
    of the language, the prefix leftmost-first semantics stops at (checked on sampled members).
The comparison walks the PRODUCT automaton breadth first over the 128 ASCII bytes, so it is a proof over that alphabet,
not a sample.  Table semantics as everywhere: match states are entered one byte late, the last class is end-of-input."""
import collections

import pytest

import zkemail_rs_b200 as z
from oracle import ra_wire as W

interegular = pytest.importorskip("interegular")
from interegular.fsm import anything_else  # noqa: E402

PATTERNS = [
    r"Transaction ID: [A-Z0-9]+", r"subject:[^\r\n]+", r"from:[^\r\n]*@example\.com", r"ab+c|[^x]d", r"(a|ab)(c|bcd)?",
    r"[a-f0-9]{4}-[a-f0-9]{2}", r"x*y?z{2,3}", r"(foo|bar|ba)+z", r"\r\nto:[^\r\n]+\r\n", r"[^a]b|a", r"a{0,2}b{1,}",
    r"(0|1(01*0)*1)+", r"[A-Za-z_][A-Za-z0-9_]*=", r"(ab|a)(bc|c)", r"email was meant for @[a-z]+\.",
]


class _Table:
    def __init__(self, blob):
        p = W.parse_zdf(blob)
        self.ns, self.nc, self.mn, self.mx = p["ns"], p["nc"], p["mn"], p["mx"]
        self.trans, self.cls = p["trans"], p["classes"]
        self.start = p["start"][6 + 2]                   # anchored, look-behind = start of text

    def step(self, s, b):
        return self.trans[s * self.nc + self.cls[b]]

    def accepts_here(self, s):                           # end of input right after the bytes read so far
        t = self.trans[s * self.nc + self.nc - 1]
        return self.mn <= t <= self.mx


def _their_step(fsm, s, ch):
    if s is None:
        return None
    sym = fsm.alphabet[ch] if ch in fsm.alphabet else fsm.alphabet[anything_else]
    return fsm.map.get(s, {}).get(sym)


def _product_equal(table, fsm, want="equal"):
    """BFS over (table state, fsm state).  want = "equal": same acceptance everywhere; "subset": table accepts => fsm accepts."""
    seen = {(table.start, fsm.initial)}
    todo = collections.deque([(table.start, fsm.initial, b"")])
    while todo:
        s, t, w = todo.popleft()
        ours, theirs = table.accepts_here(s), t is not None and t in fsm.finals
        if want == "equal" and ours != theirs:
            return w
        if want == "subset" and ours and not theirs:
            return w
        for b in range(128):
            s2 = table.step(s, b)
            t2 = _their_step(fsm, t, chr(b))
            if s2 == 0 and (t2 is None or want == "subset"):
                continue                                  # dead on our side (and nothing to disprove on theirs)
            if (s2, t2) not in seen:
                seen.add((s2, t2))
                todo.append((s2, t2, w + bytes([b])))
    return None


@pytest.mark.parametrize("pattern", PATTERNS)
def test_reverse_table_is_the_reversed_language(pattern):
    dfa = z.compile_regex(pattern)
    fsm = interegular.parse_pattern(pattern).to_fsm().reversed().reduce()
    # a table state that is dead on our side while theirs lives is a disagreement only if theirs can still accept:
    # reduce() leaves no useless states, so `None` (no transition) is the only dead state on their side
    bad = _product_equal(_Table(dfa.bwd), fsm, "equal")
    assert bad is None, (pattern, bad)


@pytest.mark.parametrize("pattern", PATTERNS)
def test_forward_table_accepts_only_members_and_finds_the_leftmost_first_prefix(pattern):
    import re
    dfa = z.compile_regex(pattern)
    fsm = interegular.parse_pattern(pattern).to_fsm().reduce()
    t = _Table(dfa.fwd)
    assert _product_equal(t, fsm, "subset") is None, pattern
    # members of the language: the anchored leftmost-first match of Python `re` (same preference order for these
    # patterns) is where the forward table reports its last match
    rx = re.compile(pattern.encode())
    members = _members(fsm, 80)
    assert len(members) >= 3, pattern
    for w in members:
        m = rx.match(w)
        st, last = t.start, None
        for i, b in enumerate(w):
            st = t.step(st, b)
            if t.mn <= st <= t.mx:
                last = i
            if st == 0:
                break
        else:
            if t.accepts_here(st):
                last = len(w)
        assert m is not None and last == m.end(), (pattern, w, last, m and m.end())


def _members(fsm, limit, max_len=40):
    """Strings of the language by seeded random walks that are steered towards a final state (distance to the nearest
    final state computed first), one or two representative ASCII bytes per alphabet symbol."""
    import random
    rng = random.Random(7)
    reps = collections.defaultdict(list)
    for b in range(128):
        ch = chr(b)
        sym = fsm.alphabet[ch] if ch in fsm.alphabet else fsm.alphabet[anything_else]
        if len(reps[sym]) < 2:
            reps[sym].append(b)
    dist = {f: 0 for f in fsm.finals}
    changed = True
    while changed:                                       # Bellman-Ford over a few dozen states
        changed = False
        for st, row in fsm.map.items():
            for sym, nxt in row.items():
                if nxt in dist and reps.get(sym) and dist.get(st, 1 << 30) > dist[nxt] + 1:
                    dist[st] = dist[nxt] + 1
                    changed = True
    out = set()
    for _ in range(limit * 20):
        st, w = fsm.initial, b""
        while len(w) < max_len:
            if st in fsm.finals and rng.random() < 0.4:
                break
            moves = [(sym, nxt) for sym, nxt in fsm.map.get(st, {}).items() if nxt in dist and reps.get(sym)]
            if not moves:
                break
            closer = [mv for mv in moves if dist[mv[1]] < dist.get(st, 1 << 30)]
            sym, nxt = rng.choice(closer if closer and rng.random() < 0.7 else moves)
            w += bytes([rng.choice(reps[sym])])
            st = nxt
        if st in fsm.finals:
            out.add(w)
        if len(out) >= limit:
            break
    return sorted(out)
