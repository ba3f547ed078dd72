"""Offline input generator — mirror of the reference's helpers/src/generator.rs:11-87
(generate_email_inputs, generate_email_with_regex_inputs), helpers/src/dkim.rs:72-111 (DNS TXT record ->
(key bytes, key type)) and core/src/email.rs:61-86 (remove_quoted_printable_soft_breaks).

The reference fetches the selector's key over DNS / HTTP (helpers/src/dkim.rs:32-70); there is no network
here, so the caller supplies a resolver: any callable (domain, selector) -> (key_bytes, key_type) | TXT record
string | None.  Everything else follows the reference: DKIM-Signature headers are walked top to bottom, a header
is a candidate when validate_header accepts it and its d= equals from_domain case-insensitively, the key of its
s= selector is loaded, and the first candidate for which verify_email_with_key passes yields the Email.  The
verification itself is the batched device path (Engine.verify_batch); the batch form advances every pending
message one candidate per round."""
from __future__ import annotations

import base64
import binascii
import ctypes as C
import struct
from typing import Callable, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from .engine import Engine, RegexError, _check, _py_pattern, canonicalize_signed_email, compile_regex, default_engine, load_library
from .structs import CompiledRegex, Email, EmailWithRegex, ExternalInput, PublicKey, RegexConfig, RegexInfo, RegexPattern

KeyAnswer = Union[None, str, Tuple[bytes, str]]
KeyResolver = Callable[[str, str], KeyAnswer]


class GeneratorError(Exception):
    """The anyhow::Error of helpers/src/generator.rs (message text as in the reference)."""


# ------------------------------------------------------------------ DER helpers (PKCS#1 <-> SPKI)
def _der_read(b: bytes, at: int) -> Tuple[int, int, int]:
    """-> (tag, content start, content end); raises ValueError on malformed / non-minimal lengths."""
    if at + 2 > len(b):
        raise ValueError("truncated DER")
    tag, l0 = b[at], b[at + 1]
    at += 2
    if l0 < 0x80:
        ln = l0
    else:
        k = l0 & 0x7F
        if k == 0 or k > 4 or at + k > len(b) or b[at] == 0:
            raise ValueError("bad DER length")
        ln = int.from_bytes(b[at:at + k], "big")
        if ln < 0x80:
            raise ValueError("non-minimal DER length")
        at += k
    if at + ln > len(b):
        raise ValueError("truncated DER")
    return tag, at, at + ln


def _der_uint(b: bytes, at: int) -> Tuple[int, int]:
    tag, s, e = _der_read(b, at)
    if tag != 0x02 or e == s or b[s] & 0x80 or (e - s > 1 and b[s] == 0 and not b[s + 1] & 0x80):
        raise ValueError("bad DER INTEGER")
    return int.from_bytes(b[s:e], "big"), e


def _der_len(n: int) -> bytes:
    if n < 0x80:
        return bytes([n])
    raw = n.to_bytes((n.bit_length() + 7) // 8, "big")
    return bytes([0x80 | len(raw)]) + raw


def _der_int(v: int) -> bytes:
    raw = v.to_bytes(v.bit_length() // 8 + 1, "big")
    return b"\x02" + _der_len(len(raw)) + raw


def pkcs1_from_numbers(n: int, e: int) -> bytes:
    body = _der_int(n) + _der_int(e)
    return b"\x30" + _der_len(len(body)) + body


def _parse_pkcs1(der: bytes) -> Tuple[int, int]:
    tag, s, e = _der_read(der, 0)
    if tag != 0x30 or e != len(der):
        raise ValueError("not an RSAPublicKey")
    n, at = _der_uint(der, s)
    ex, at = _der_uint(der, at)
    if at != e:
        raise ValueError("trailing data in RSAPublicKey")
    return n, ex


_RSA_ALGID = bytes.fromhex("300d06092a864886f70d0101010500")


def _parse_spki(der: bytes) -> Tuple[int, int]:
    tag, s, e = _der_read(der, 0)
    if tag != 0x30 or e != len(der):
        raise ValueError("not a SubjectPublicKeyInfo")
    t2, s2, e2 = _der_read(der, s)
    if t2 != 0x30 or der[s:e2] != _RSA_ALGID:
        raise ValueError("not an rsaEncryption key")
    t3, s3, e3 = _der_read(der, e2)
    if t3 != 0x03 or e3 != e or s3 == e3 or der[s3] != 0:
        raise ValueError("bad BIT STRING")
    return _parse_pkcs1(der[s3 + 1:e3])


def rsa_key_to_pkcs1(decoded: bytes) -> bytes:
    """helpers/src/dkim.rs:96-102: from_public_key_der(..).or_else(from_pkcs1_der(..))?.to_pkcs1_der()."""
    try:
        n, e = _parse_spki(decoded)
    except ValueError:
        n, e = _parse_pkcs1(decoded)
    return pkcs1_from_numbers(n, e)


def parse_dkim_key_record(value: str) -> Tuple[bytes, str]:
    """helpers/src/dkim.rs:72-111: `k=` / `p=` of a DKIM TXT record -> (key bytes, key type)."""
    key_type, public_key = "", ""
    for part in (p.strip() for p in value.split(";")):
        if part.startswith("k="):
            key_type = part[2:]
        if part.startswith("p="):
            public_key = part[2:]
    if not key_type:
        key_type = "rsa"
    if not public_key:
        raise GeneratorError("No public key found")
    if key_type not in ("rsa", "ed25519"):
        raise GeneratorError(f"Unsupported key type: {key_type}")
    try:
        decoded = base64.b64decode(public_key, validate=True)
    except (binascii.Error, ValueError) as e:
        raise GeneratorError(f"base64: {e}") from e
    if key_type == "rsa":
        try:
            return rsa_key_to_pkcs1(decoded), "rsa"
        except ValueError as e:
            raise GeneratorError(f"rsa key: {e}") from e
    if len(decoded) != 32:
        raise GeneratorError("Invalid Ed25519 key length")
    return decoded, "ed25519"


class StaticKeys:
    """Resolver over a dict {(domain, selector): TXT record | (key, key_type)} (domain compared lower-case)."""

    def __init__(self, table: Dict[Tuple[str, str], KeyAnswer]):
        self.table = {(d.lower(), s): v for (d, s), v in table.items()}

    def __call__(self, domain: str, selector: str) -> KeyAnswer:
        return self.table.get((domain.lower(), selector))


def _resolve(keys: KeyResolver, domain: str, selector: str) -> Optional[Tuple[bytes, str]]:
    """fetch_dkim_key(..) as an Option: None where the reference's `if let Ok(..)` falls through."""
    try:
        ans = keys(domain, selector)
        if ans is None:
            return None
        if isinstance(ans, str):
            if "p=" not in ans or ans.endswith("p="):  # helpers/src/dkim.rs:66-68
                return None
            ans = parse_dkim_key_record(ans)
        key, kt = ans
        return bytes(key), str(kt)
    except GeneratorError:
        return None


# ------------------------------------------------------------------ message inspection (host front end)
def dkim_signatures(raw_email: bytes, now_unix: int = 0) -> List[Optional[Tuple[str, str]]]:
    """One entry per DKIM-Signature header, top to bottom: (d, s) when validate_header accepts it, else None.
    Raises GeneratorError when the message does not parse (mailparse::parse_mail(..)?)."""
    L = load_library()
    if not getattr(L, "_zkb_sigs_bound", False):
        L.zkb_host_dkim_signatures.argtypes = [C.c_char_p, C.c_size_t, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L._zkb_sigs_bound = True
    out, ol, ns = C.c_void_p(), C.c_size_t(), C.c_size_t()
    if L.zkb_host_dkim_signatures(raw_email, len(raw_email), now_unix, C.byref(out), C.byref(ol), C.byref(ns)) != 0:
        raise GeneratorError("mail parse error")
    try:
        buf = C.string_at(out.value, ol.value)
    finally:
        L.zkb_free(out)
    res, at = [], 0
    for _ in range(ns.value):
        valid, dl, sl = struct.unpack_from("<III", buf, at)
        at += 12
        if valid:
            res.append((buf[at:at + dl].decode("utf-8", "replace"), buf[at + dl:at + dl + sl].decode("utf-8", "replace")))
        else:
            res.append(None)
        at += dl + sl
    return res


def remove_quoted_printable_soft_breaks(body: bytes) -> Tuple[bytes, np.ndarray]:
    """core/src/email.rs:61-86: drops every `=\\r\\n`, zero-pads to the input length; the index map holds the
    original position of each kept byte and usize::MAX for the padding."""
    n = len(body)
    a = np.frombuffer(body, dtype=np.uint8)
    keep = np.ones(n, dtype=bool)
    i = body.find(b"=\r\n")
    while i >= 0:
        keep[i:i + 3] = False
        i = body.find(b"=\r\n", i + 3)
    idx = np.nonzero(keep)[0].astype(np.uint64)
    cleaned = np.zeros(n, dtype=np.uint8)
    cleaned[:idx.size] = a[idx.astype(np.int64)]
    imap = np.full(n, np.iinfo(np.uint64).max, dtype=np.uint64)
    imap[:idx.size] = idx
    return cleaned.tobytes(), imap


# ------------------------------------------------------------------ generate_email_inputs
def generate_email_inputs_batch(items: Sequence[Tuple[str, bytes]], keys: KeyResolver,
                                external_inputs: Optional[Sequence[Optional[List[ExternalInput]]]] = None,
                                engine: Optional[Engine] = None) -> List[Union[Email, GeneratorError]]:
    """generate_email_inputs (helpers/src/generator.rs:11-53) for many (from_domain, raw_email) pairs.  Each round
    verifies, in one device batch, the next candidate signature of every message that has not passed yet."""
    eng = engine or default_engine()
    n = len(items)
    out: List[Union[None, Email, GeneratorError]] = [None] * n
    cands: List[List[str]] = []
    for i, (dom, raw) in enumerate(items):
        try:
            sigs = dkim_signatures(raw, eng.now_unix)
        except GeneratorError as e:
            out[i] = e
            cands.append([])
            continue
        if not sigs:
            out[i] = GeneratorError("No DKIM signatures found")  # generator.rs:20-22
        cands.append([s for ds in sigs if ds is not None and ds[0].lower() == dom.lower() for s in [ds[1]]])
    pos = [0] * n
    while True:
        batch, owners = [], []
        for i, (dom, raw) in enumerate(items):
            if out[i] is not None:
                continue
            email = None
            while pos[i] < len(cands[i]) and email is None:
                got = _resolve(keys, dom, cands[i][pos[i]])
                pos[i] += 1
                if got is not None:
                    ext = list(external_inputs[i] or []) if external_inputs is not None else []
                    email = Email(dom, bytes(raw), PublicKey(got[0], got[1]), ext)
            if email is None:
                out[i] = GeneratorError("No valid DKIM key found for any signature")  # generator.rs:52
            else:
                batch.append(email)
                owners.append(i)
        if not batch:
            break
        recs = eng.verify_batch(batch)
        for e, i, r in zip(batch, owners, recs):
            if int(r["status"]) == 0:   # try_from_bytes ok, verify ok, detail starts with "pass"
                out[i] = e
    return out  # type: ignore[return-value]


def generate_email_inputs(from_domain: str, raw_email: bytes, keys: KeyResolver,
                          external_inputs: Optional[List[ExternalInput]] = None, engine: Optional[Engine] = None) -> Email:
    r = generate_email_inputs_batch([(from_domain, raw_email)], keys, [external_inputs], engine)[0]
    if isinstance(r, GeneratorError):
        raise r
    return r


# ------------------------------------------------------------------ generate_email_with_regex_inputs
def compile_regex_parts_exact(parts: Sequence[RegexPattern], haystack: bytes, engine: Optional[Engine] = None) -> List[CompiledRegex]:
    """helpers/src/regex.rs:16-51 with the exactly-one-match test done by the device DFA scan (regex-automata
    find_iter semantics, not Python's); capture strings from Python `re` anchored on the device's span."""
    import re
    eng = engine or default_engine()
    out = []
    for part in parts:
        dfa = compile_regex(part.pattern)
        cnt, start, end, _ = (int(x) for x in eng.dfa_scan_batch(dfa, [haystack])[0])
        if cnt != 1:
            raise RegexError(f"Input doesn't match regex pattern: {part!r}")
        caps: List[str] = []
        if part.capture_indices:
            # MetaRegex::captures searches the WHOLE input (helpers/src/regex.rs:25-27); the only match is the
            # device's span, so group 0 is that span.  Sub-groups come from Python `re` over a translated pattern,
            # searched from the span start without an end limit (`$`, `\b` see the same context as in the reference);
            # a translation whose overall match is not exactly the automaton's span is rejected, never trusted.
            m = None
            if any(idx != 0 for idx in part.capture_indices):
                rx = re.compile(_py_pattern(part.pattern))
                m = rx.search(haystack, start)
                if m is None or m.span() != (start, end):
                    raise RegexError(f"capture groups of {part.pattern!r} cannot be resolved exactly "
                                     f"(automaton span {(start, end)}, translated pattern {m.span() if m else None})")
            for idx in part.capture_indices:
                if idx == 0:
                    caps.append(haystack[start:end].decode("utf-8", errors="replace"))
                    continue
                if idx > m.re.groups or m.group(idx) is None:
                    raise RegexError("Capture group not found")
                caps.append(m.group(idx).decode("utf-8", errors="replace"))
        out.append(CompiledRegex(verify_re=dfa, captures=caps))
    return out


def generate_email_with_regex_inputs(from_domain: str, raw_email: bytes, regex_config: RegexConfig, keys: KeyResolver,
                                     external_inputs: Optional[List[ExternalInput]] = None,
                                     engine: Optional[Engine] = None) -> EmailWithRegex:
    """helpers/src/generator.rs:55-87."""
    eng = engine or default_engine()
    email = generate_email_inputs(from_domain, raw_email, keys, external_inputs, eng)
    try:
        header, body = canonicalize_signed_email(raw_email, eng.now_unix)
    except Exception as e:
        raise GeneratorError(str(e)) from e
    cleaned, _ = remove_quoted_printable_soft_breaks(body)
    bp = compile_regex_parts_exact(regex_config.body_parts, cleaned, eng) if regex_config.body_parts else None
    hp = compile_regex_parts_exact(regex_config.header_parts, header, eng) if regex_config.header_parts else None
    return EmailWithRegex(email, RegexInfo(header_parts=hp, body_parts=bp))
