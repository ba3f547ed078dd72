"""Patterns of the kind zk-email circuits use (header lines with a `(\\r\\n|^)` prefix, address and amount extractors,
lazy and bounded repetitions, POSIX classes): the library's compiler + the oracle's find_iter + the kernel source
under emulation must all give Python `re`'s spans (none of these patterns has empty matches, where Rust and Python
differ)."""
import re

import numpy as np

import oracle
import zkemail_rs_b200 as z
from tests import emu

PATTERNS = [
    r"(\r\n|^)to:[^\r\n]+\r\n", r"(\r\n|^)subject:[^\r\n]+\r\n", r"dkim-signature:([a-z]+=[^;]+; )+t=[0-9]+;",
    r"email was meant for @[a-zA-Z0-9_]+", r"[A-Za-z0-9!#$%&'*+=?\-\^_`{|}~./@]+@[A-Za-z0-9.\-]+",
    r"(\r\n|^)from:([^\r\n]+<)?[A-Za-z0-9.]+@[a-z.]+>?\r\n", r"\$[0-9]+(\.[0-9]{2})?", r"(?i)x-mailer: [a-z ]+",
    r"[^\x00-\x1f]+?@", r"\d{4}-\d{2}-\d{2}", r"(?s)begin.*?end", r"a{2,}b|c+?d", r"[[:alpha:]]+[[:digit:]]",
]
HAYS = [
    b"to:alice@example.com\r\nsubject:hello $12.50 world\r\nfrom:Bob <bob.x@mail.example.com>\r\ndkim-signature:v=1; a=rsa; t=1700;\r\n"
    b"email was meant for @alice_1 x-mailer: Foo Bar 2024-01-31 begin a\nb end end aab ccd abc1\r\n",
    b"subject:only\r\n", b"", b"from:x@y.z\r\nto:a@b.c\r\nto:d@e.f\r\n", b"$5 $6.7 $8.90 1999-12-31x2000-01-01", b"X-Mailer: abc\r\nx-mailer: DEF ghi\r\n",
]


def _py(pattern):
    return re.compile(pattern.replace("[[:alpha:]]", "[A-Za-z]").replace("[[:digit:]]", "[0-9]").encode())


def test_compiler_oracle_and_kernel_source_agree_with_python_re():
    for pat in PATTERNS:
        d = z.compile_regex(pat)
        rows = np.asarray(emu.dfa_scan(d.fwd, d.bwd, HAYS))
        for hay, row in zip(HAYS, rows):
            want = [(m.start(), m.end()) for m in _py(pat).finditer(hay)]
            cnt, spans = oracle.dfa_find_iter(d.fwd, d.bwd, hay)
            assert cnt == len(want) and [tuple(s) for s in spans[:cnt]] == want, (pat, hay)
            assert int(row[0]) == cnt, (pat, hay)
            if cnt:
                assert (int(row[1]), int(row[2])) == want[0], (pat, hay)


def test_repetitions_of_sub_expressions_that_can_match_the_empty_string():
    """Leftmost-first preference in loops whose body can match the empty string depends on the SHAPE of the Thompson NFA,
    not only on the language.  regex-automata compiles `e+` as one copy of e with a union behind it, and `e*` with such an
    e as `(e+)?` (rust-lang/regex issue 779: with a plain loop `(|a)*` takes "aaa" where the crate - like Perl - takes
    the empty string at every position).  csrc/regexc.hpp mirrors those shapes; found by tools/fuzz_regex.py against
    Python `re`, which agrees with the crate on these patterns."""
    def spans(pat, hay):
        d = z.compile_regex(pat)
        cnt, sp = oracle.dfa_find_iter(d.fwd, d.bwd, hay)
        return [tuple(x) for x in sp[:cnt]]
    assert spans(r"(?:|a)*", b"aaa") == [(0, 0), (1, 1), (2, 2), (3, 3)]
    assert spans(r"(?:|a)+", b"aaa") == [(0, 0), (1, 1), (2, 2), (3, 3)]
    assert spans(r"(?:a|)+", b"aaa") == [(0, 3)]
    assert spans(r"1(b?|[xyz]+\d)*", b"01xy9z") == [(1, 2)]            # the empty first alternative ends the loop
    assert spans(r"1([xyz]+\d|b?)*", b"01xy9z") == [(1, 5)]
    assert spans(r"b(y*|0??)*", b"xb0bby") == [(1, 2), (3, 4), (4, 6)]
    assert spans(r"b{1,3}(y*|0??)*", b"xb0bby") == [(1, 2), (3, 6)]
    assert spans(r"x(?:a??|b)+?y", b"xaby xy") == [(0, 4), (5, 7)]
    for pat, hay in [(r"1(b?|[xyz]+\d)*", b"01xy9z 1b1 1x0y1"), (r"(\r\n|^)k:( ?[a-z]*)+;", b"k: ab cd;\r\nk:;\r\nk:x y z ;")]:
        want = [(m.start(), m.end()) for m in re.finditer(pat.encode(), hay) if m.start() != m.end()]
        assert spans(pat, hay) == want, pat


def test_ascii_word_boundaries():
    r"""(?-u:\b) / (?-u:\B) compile (one "previous byte was a word byte" bit per DFA state) and agree with Python's
    bytes-mode \b; Unicode \b is rejected like DFARegex::new rejects it (helpers/src/regex.rs:20 would return Err)."""
    cases = [(r"(?-u:\b)foo(?-u:\b)", rb"\bfoo\b"), (r"(?-u)\bfoo\b", rb"\bfoo\b"), (r"(?-u:\B)oo", rb"\Boo"), (r"(?-u)\b[a-z]+\b", rb"\b[a-z]+\b"),
             (r"(?-u)\b\d+\b", rb"\b\d+\b"), (r"(?-u)x\b", rb"x\b"), (r"(?-u)\bx", rb"\bx"), (r"(?-u)a\Bb", rb"a\Bb"),
             (r"(?-u)\b(cat|dog)s?\b", rb"\b(cat|dog)s?\b"), (r"(?i-u)\bID: [a-z0-9]+\b", rb"(?i)\bID: [a-z0-9]+\b"), (r"(?m-u)^\w+\b", rb"(?m)^\w+\b")]
    hays = [b"foo", b" foo ", b"foofoo foo_foo foo-foo", b"a foo\xc3\xa9 foo", b"", b"x", b"xx x.x ax xa", b"ab a b", b"cats dog dogs cat. catsup",
            b"id: a1 ID: zz9_ ID: Q ", b"12 a12 12a 3", b"the cat sat on the mat", b"\nfoo\n", b"foo\xff", b"one\ntwo three\nfour"]
    for pat, py in cases:
        d = z.compile_regex(pat)
        rows = np.asarray(emu.dfa_scan(d.fwd, d.bwd, hays))
        for hay, row in zip(hays, rows):
            want = [(m.start(), m.end()) for m in re.finditer(py, hay)]
            cnt, spans = oracle.dfa_find_iter(d.fwd, d.bwd, hay)
            assert [tuple(s) for s in spans[:cnt]] == want, (pat, hay)
            assert int(row[0]) == cnt and (not cnt or (int(row[1]), int(row[2])) == want[0]), (pat, hay)
    # patterns that match the empty string: Rust's iteration rule, kernel source against the oracle
    for pat in (r"(?-u)\b", r"(?-u)\B", r"(?-u)\b|x"):
        d = z.compile_regex(pat)
        for hay, row in zip(hays, np.asarray(emu.dfa_scan(d.fwd, d.bwd, hays))):
            cnt, spans = oracle.dfa_find_iter(d.fwd, d.bwd, hay)
            assert int(row[0]) == cnt and (not cnt or (int(row[1]), int(row[2])) == tuple(spans[0])), (pat, hay)
    d = z.compile_regex(r"(?-u)\b")
    assert [tuple(s) for s in oracle.dfa_find_iter(d.fwd, d.bwd, b"ab cd")[1][:4]] == [(0, 0), (2, 2), (3, 3), (5, 5)]
    import pytest
    for pat in (r"\bfoo", r"(?u)\bx", r"[\b]"):
        with pytest.raises(z.RegexError):
            z.compile_regex(pat)


def test_capture_resolution_accepts_rust_flag_groups():
    """compile_regex_parts (helpers/src/regex.rs:16-51) resolves capture strings with Python `re`; Rust's `u` flag
    spellings must not trip it."""
    from zkemail_rs_b200.structs import RegexPattern
    hay = b"x ID: abc9 y\r\nsubject:Hi there\r\n"
    for pat, want in ((r"(?-u:\b)ID: ([a-z0-9]+)", ["abc9"]), (r"(?-u)\bID: ([a-z0-9]+)\b", ["abc9"]), (r"(?i-u)\bid: ([a-z0-9]+)", ["abc9"]),
                      (r"(?u)subject:(\w+) (\w+)", ["Hi", "there"]), (r"(?<n>ID): (\w+)", ["ID", "abc9"])):
        idx = list(range(1, len(want) + 1))
        assert z.compile_regex_parts([RegexPattern(pat, idx)], hay)[0].captures == want, pat


def test_unicode_general_category_classes():
    r"""\p{..} / \P{..} with general categories against Python's unicodedata (the tables are generated from it)."""
    import unicodedata
    import pytest
    cat = unicodedata.category
    text = "Hello Wörld ÀÉî ßtraße ΑΒγδ Жд 123 ٤٥٦ x_y — “quoted” €5 + ∑    end\t!"

    def runs(pred):
        out, i = [], 0
        while i < len(text):
            if pred(text[i]):
                j = i
                while j < len(text) and pred(text[j]):
                    j += 1
                out.append((len(text[:i].encode()), len(text[:j].encode())))
                i = j
            else:
                i += 1
        return out

    cases = [(r"\p{Lu}+", lambda c: cat(c) == "Lu"), (r"\p{L}+", lambda c: cat(c)[0] == "L"), (r"\pL+", lambda c: cat(c)[0] == "L"),
             (r"\p{Letter}+", lambda c: cat(c)[0] == "L"), (r"\p{Nd}+", lambda c: cat(c) == "Nd"), (r"\p{gc=Decimal_Number}+", lambda c: cat(c) == "Nd"),
             (r"\P{L}+", lambda c: cat(c)[0] != "L"), (r"\p{^L}+", lambda c: cat(c)[0] != "L"), (r"\p{P}+", lambda c: cat(c)[0] == "P"),
             (r"\p{Sc}+", lambda c: cat(c) == "Sc"), (r"\p{Sm}+", lambda c: cat(c) == "Sm"), (r"\p{Z}+", lambda c: cat(c)[0] == "Z"),
             (r"[\p{Lu}\p{Nd}]+", lambda c: cat(c) in ("Lu", "Nd")), (r"[^\p{L}\p{Z}]+", lambda c: cat(c)[0] not in "LZ"),
             (r"\p{Ll}+", lambda c: cat(c) == "Ll"), (r"\p{ASCII}+", lambda c: ord(c) < 128)]
    hay = text.encode()
    for pat, pred in cases:
        d = z.compile_regex(pat)
        cnt, spans = oracle.dfa_find_iter(d.fwd, d.bwd, hay)
        want = runs(pred)
        assert [tuple(s) for s in spans[:cnt]] == want, pat
        row = np.asarray(emu.dfa_scan(d.fwd, d.bwd, [hay]))[0]
        assert int(row[0]) == cnt and (int(row[1]), int(row[2])) == want[0], pat
    for pat in (r"\p{Greek}", r"(?-u)\p{L}", r"(?i)\p{Lu}", r"\p{", r"\p{Xx}"):
        with pytest.raises(z.RegexError):
            z.compile_regex(pat)


def test_unicode_simple_case_folding():
    """(?i) uses Unicode simple case folding (CaseFolding C + S), as the Rust regex crate does: the folding classes of
    non-ASCII letters are honoured, the Turkic dotted / dotless i are NOT folded onto i (Python's re does fold them)."""
    text = "É é STRASSE Straße ẞ Я я K k K s S ſ ǅ Ǆ ǆ σ ς Σ İ i I ı µ Μ μ Ꭰ ꭰ ÿ Ÿ"
    hay = text.encode()

    def found(pat):
        d = z.compile_regex("(?i)" + pat)
        cnt, spans = oracle.dfa_find_iter(d.fwd, d.bwd, hay)
        row = np.asarray(emu.dfa_scan(d.fwd, d.bwd, [hay]))[0]
        assert int(row[0]) == cnt
        return [hay[s:e].decode() for s, e in spans[:cnt]]

    for pat in ("é", "я", "k", "s", "ǆ", "σ", "ß", "ẞ", "µ", "ꭰ", "ÿ", "[а-я]", "[α-ω]"):
        assert found(pat) == [m.group() for m in re.finditer(pat, text, re.IGNORECASE)], pat
    assert found("i") == ["i", "I"] and found("ı") == ["ı"] and found("İ") == ["İ"]
    # byte mode folds ASCII letters only
    d = z.compile_regex(r"(?i-u)[a-z]+")
    assert oracle.dfa_find_iter(d.fwd, d.bwd, b"abC \xc3\x89")[1][:2] == [(0, 3)]
