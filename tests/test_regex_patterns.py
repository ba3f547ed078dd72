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
