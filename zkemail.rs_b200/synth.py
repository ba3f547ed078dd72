"""Synthetic DKIM-signed mail generator (offline stand-in for helpers/src/generator.rs:11-53, which
needs DNS/HTTP for the key) and an INDEPENDENT RFC 6376 canonicaliser written from the RFC text
(§3.4.2, §3.4.4), used to sign.  Because a signature either verifies or not, messages signed with
this canonicaliser pin the C oracle and the CUDA path from a third, independent implementation.

Pure Python + `cryptography`; used by tests/ and by bench.py for small pools.  (bench.py's large
pools are produced by the C generator in workload/zk_gen.c with the same recipe.)
"""
from __future__ import annotations

import base64
import hashlib
import re
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
from cryptography.hazmat.primitives import hashes, serialization
from cryptography.hazmat.primitives.asymmetric import padding, rsa

from .structs import Email, PublicKey

# ------------------------------------------------------------------ RFC 6376 canonicalisation


def relaxed_body(body: bytes) -> bytes:
    """RFC 6376 §3.4.4, for bodies whose line terminators are CRLF (the RFC-clean envelope)."""
    lines = body.split(b"\r\n")
    out = []
    for ln in lines:
        ln = re.sub(rb"[ \t]+", b" ", ln)
        ln = re.sub(rb" +$", b"", ln)
        out.append(ln)
    res = b"\r\n".join(out)
    # ignore all empty lines at the end; ensure a trailing CRLF unless empty
    while res.endswith(b"\r\n\r\n"):
        res = res[:-2]
    if res == b"\r\n":
        return res  # cfdkim quirk kept out of generated mail; see tests
    if res and not res.endswith(b"\r\n"):
        res += b"\r\n"
    return res


def simple_body(body: bytes) -> bytes:
    """RFC 6376 §3.4.3."""
    if body == b"":
        return b"\r\n"
    while body.endswith(b"\r\n\r\n"):
        body = body[:-2]
    return body


def relaxed_header(name: bytes, value: bytes) -> bytes:
    """RFC 6376 §3.4.2.  `value` is the raw field body after the colon (folds included)."""
    v = value.replace(b"\r\n", b"")
    v = re.sub(rb"[ \t]+", b" ", v).strip(b" ")
    return name.lower().rstrip() + b":" + v + b"\r\n"


def simple_header_cfdkim(name: bytes, value: bytes) -> bytes:
    """cfdkim's reconstruction ``key: value`` (NOT the verbatim line) — SURVEY.md A.2."""
    return name + b": " + value.lstrip(b" ") + b"\r\n"


# ------------------------------------------------------------------ keys


@dataclass
class KeyPair:
    private: rsa.RSAPrivateKey
    der: bytes  # PKCS#1 RSAPublicKey DER — the PublicKey.key contract (helpers/src/dkim.rs:50)
    bits: int

    @property
    def n(self) -> int:
        return self.private.public_key().public_numbers().n

    @property
    def e(self) -> int:
        return self.private.public_key().public_numbers().e

    def private_der(self) -> bytes:
        return self.private.private_bytes(
            serialization.Encoding.DER,
            serialization.PrivateFormat.TraditionalOpenSSL,
            serialization.NoEncryption(),
        )

    @staticmethod
    def from_private_der(der: bytes) -> "KeyPair":
        k = serialization.load_der_private_key(der, password=None)
        return KeyPair._wrap(k)

    @staticmethod
    def _wrap(k) -> "KeyPair":
        pub = k.public_key().public_bytes(serialization.Encoding.DER, serialization.PublicFormat.PKCS1)
        return KeyPair(k, pub, k.key_size)

    @staticmethod
    def generate(bits: int = 2048, e: int = 65537) -> "KeyPair":
        return KeyPair._wrap(rsa.generate_private_key(public_exponent=e, key_size=bits))


# ------------------------------------------------------------------ message construction

DEFAULT_H = ("from", "to", "subject", "date", "message-id")


def fold_b64(s: str, first: int = 60, width: int = 72) -> str:
    """Fold a base64 string over several header lines (CRLF + one TAB continuation)."""
    parts = [s[:first]]
    s = s[first:]
    while s:
        parts.append(s[:width])
        s = s[width:]
    return "\r\n\t".join(parts)


def synth_body(rng: np.random.Generator, canon_len: int, qp_soft_breaks: bool = False,
               token: Optional[bytes] = None) -> bytes:
    """Printable-ASCII body of lines <= 76 chars + CRLF whose relaxed-canonical length is exactly
    `canon_len` (SURVEY.md §8d C1/C2 recipe).  No trailing WSP, so canonical == raw for
    canon_len >= 3 (shorter bodies are a bare run of letters without a line terminator)."""
    if canon_len < 3:
        return b"x" * canon_len
    alphabet = np.frombuffer(
        b"abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789 ,.;:-_()", dtype=np.uint8
    )

    def text(n: int) -> bytes:
        line = bytearray(alphabet[rng.integers(0, len(alphabet), size=n)].tobytes())
        if line[0:1] == b" ":
            line[0:1] = b"x"
        if line[-1:] == b" ":
            line[-1:] = b"x"
        return re.sub(rb"  +", lambda m: b" " + b"y" * (len(m.group(0)) - 1), bytes(line))

    lines: List[bytes] = []
    total = 0
    token_at = None
    if token is not None:
        tok_line = bytes(token)
        if qp_soft_breaks and len(tok_line) > 6:
            cut = int(rng.integers(1, len(tok_line) - 1))
            tok_line = tok_line[:cut] + b"=\r\n" + tok_line[cut:]
        tok_line += b"\r\n"
        assert canon_len >= len(tok_line) + 3, "body too short for the token line"
        token_at = int(rng.integers(0, max(1, canon_len - len(tok_line) - 80)))
    while total < canon_len:
        remaining = canon_len - total
        if token_at is not None and total >= token_at and remaining >= len(tok_line):
            lines.append(tok_line)
            total += len(tok_line)
            token_at = None
            continue
        budget = remaining - (len(tok_line) if token_at is not None else 0)
        if budget < 3:
            budget = remaining
        ll = int(min(rng.integers(40, 77), budget - 2))
        left = budget - (ll + 2)
        if left in (1, 2):
            ll = ll + left if ll + left <= 76 else ll - (3 - left)
        ll = max(1, ll)
        line = text(ll)
        if qp_soft_breaks and rng.random() < 0.3 and ll > 12:
            line = line[: ll - 3] + b"=\r\n" + line[ll - 3:]
            # the soft break adds 3 raw bytes that are part of the canonical body as well
            if total + len(line) + 2 > canon_len - (len(tok_line) if token_at is not None else 0):
                line = line[: ll - 3] + line[ll:]
        lines.append(line + b"\r\n")
        total += len(lines[-1])
    body = b"".join(lines)
    if len(body) != canon_len:  # final adjustment: trim or pad the last text line
        diff = len(body) - canon_len
        for i in range(len(lines) - 1, -1, -1):
            ln = lines[i]
            if token is not None and ln == tok_line:
                continue
            core = ln[:-2]
            if diff > 0 and len(core) - diff >= 1 and b"=\r\n" not in core[-(diff + 3):]:
                core = core[: len(core) - diff]
                if core.endswith(b" "):
                    core = core[:-1] + b"x"
                lines[i] = core + b"\r\n"
                break
            if diff < 0:
                lines[i] = core + b"z" * (-diff) + b"\r\n"
                break
        body = b"".join(lines)
    assert len(body) == canon_len, (len(body), canon_len)
    assert relaxed_body(body) == body, "synthetic body must be canonical-stable"
    return body


def sign_email(
    headers: Sequence[Tuple[str, str]],
    body: bytes,
    key: KeyPair,
    domain: str,
    selector: str = "sel1",
    canon: str = "relaxed/relaxed",
    h: Sequence[str] = DEFAULT_H,
    extra_tags: str = "",
    sig_position: str = "top",
    algo: str = "rsa-sha256",
    omit_c: bool = False,
    meta: Optional[dict] = None,
) -> bytes:
    """Build a raw RFC 5322 message with one DKIM-Signature (RFC 6376 §3.5/§3.7)."""
    hc, bc = (canon.split("/") + ["simple"])[:2] if "/" in canon else (canon, "simple")
    cbody = relaxed_body(body) if bc == "relaxed" else simple_body(body)
    bh = base64.b64encode(hashlib.sha256(cbody).digest()).decode()
    ctag = "" if omit_c else f" c={canon};"
    value_nob = (
        f" v=1; a={algo};{ctag} d={domain}; s={selector};\r\n"
        f"\th={':'.join(h)};{extra_tags}\r\n"
        f"\tbh={bh};\r\n"
        f"\tb="
    )
    hdr_bytes = [(k.encode(), (" " + v).encode()) for k, v in headers]
    # select headers bottom-up per RFC 6376 §5.4.2
    used: dict = {}
    pre = b""
    for name in h:
        lname = name.strip().lower().encode()
        start = used.get(lname, len(hdr_bytes))
        hit = None
        for j in range(start - 1, -1, -1):
            if hdr_bytes[j][0].lower() == lname:
                hit = j
                break
        used[lname] = hit if hit is not None else 0
        if hit is not None:
            k, v = hdr_bytes[hit]
            pre += relaxed_header(k, v) if hc == "relaxed" else simple_header_cfdkim(k, v)
    sigv = value_nob.encode()
    if hc == "relaxed":
        pre += relaxed_header(b"DKIM-Signature", sigv)[:-2]
    else:
        pre += simple_header_cfdkim(b"DKIM-Signature", sigv)[:-2]
    sig = key.private.sign(pre, padding.PKCS1v15(), hashes.SHA256())
    if meta is not None:  # what an RFC 6376 verifier must hash (used to build golden fixtures)
        meta.update(header_preimage=pre, canonical_body=cbody, signature=sig)
    b = fold_b64(base64.b64encode(sig).decode())
    sig_header = b"DKIM-Signature:" + value_nob.encode() + b.encode() + b"\r\n"
    lines = [k + b":" + v + b"\r\n" for k, v in hdr_bytes]
    block = (sig_header + b"".join(lines)) if sig_position == "top" else (b"".join(lines) + sig_header)
    return block + b"\r\n" + body


def default_headers(rng: np.random.Generator, domain: str, idx: int) -> List[Tuple[str, str]]:
    user = "".join(chr(97 + int(c)) for c in rng.integers(0, 26, size=8))
    subj = "".join(chr(97 + int(c)) for c in rng.integers(0, 26, size=24))
    return [
        ("Received", f"from mx.{domain} by relay.example.net; Mon, 1 Jan 2024 00:00:00 +0000"),
        ("From", f"{user.capitalize()} <{user}@{domain}>"),
        ("To", f"recipient{idx}@example.org"),
        ("Subject", f"Order {subj} update {idx}"),
        ("Date", "Mon, 01 Jan 2024 00:00:00 +0000"),
        ("Message-ID", f"<{idx:08d}.{user}@{domain}>"),
        ("MIME-Version", "1.0"),
        ("Content-Type", "text/plain; charset=us-ascii"),
    ]


def make_email(rng: np.random.Generator, key: KeyPair, domain: str, idx: int = 0,
               body_len: int = 4096, **kw) -> Email:
    body_kw = {k: kw.pop(k) for k in ("qp_soft_breaks", "token") if k in kw}
    body = synth_body(rng, body_len, **body_kw) if body_len > 0 else b""
    raw = sign_email(default_headers(rng, domain, idx), body, key, domain, **kw)
    return Email(from_domain=domain, raw_email=raw, public_key=PublicKey(key.der, "rsa"))


# ------------------------------------------------------------------ negative-case mutators


def mutate(email: Email, kind: str, rng: np.random.Generator, other_key: Optional[KeyPair] = None) -> Email:
    """Fault injection: each kind must flip the verdict identically in the
    oracle and on the GPU."""
    raw = bytearray(email.raw_email)
    sep = raw.find(b"\r\n\r\n")
    if kind == "body_flip":
        i = sep + 4 + int(rng.integers(0, max(1, len(raw) - sep - 4)))
        raw[i] = raw[i] ^ 0x01 if raw[i] ^ 0x01 not in (0x0D, 0x0A, 0x20, 0x09) else raw[i] ^ 0x40
    elif kind == "sig_flip":
        i = raw.find(b"\tb=") + 3
        c = raw[i + 5]
        raw[i + 5] = ord("A") if c != ord("A") else ord("B")
    elif kind == "wrong_key":
        assert other_key is not None
        return Email(email.from_domain, bytes(raw), PublicKey(other_key.der, "rsa"), email.external_inputs)
    elif kind == "bh_flip":
        i = raw.find(b"\tbh=") + 4
        raw[i] = ord("A") if raw[i] != ord("A") else ord("B")
    elif kind == "header_flip":
        i = raw.find(b"Subject: ") + 9
        raw[i] = raw[i] ^ 0x01
    elif kind == "domain_mismatch":
        return Email("other-" + email.from_domain, bytes(raw), email.public_key, email.external_inputs)
    elif kind == "missing_tag":
        i = raw.find(b" s=")
        j = raw.find(b";", i)
        del raw[i : j + 1]
    else:
        raise ValueError(kind)
    return Email(email.from_domain, bytes(raw), email.public_key, email.external_inputs)


NEGATIVE_KINDS = ("body_flip", "sig_flip", "wrong_key", "bh_flip", "header_flip", "domain_mismatch", "missing_tag")
