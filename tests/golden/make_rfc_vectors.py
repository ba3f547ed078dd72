"""Writes tests/golden/rfc_vectors.json: the signed sample messages of RFC 8463 Appendix A (one message
carrying an ed25519-sha256 and an rsa-sha256 signature, relaxed/relaxed, RSA-1024) and of RFC 6376
Appendix A.2 (rsa-sha256, simple/simple, RSA-1024), with the public keys the RFCs publish.

These are signatures the builder of this repository did NOT make.  They are self-validating: a
mis-transcribed byte cannot verify.  This script checks each vector with an implementation that shares
nothing with the oracle or the engine (`cryptography`/OpenSSL for RSA and Ed25519, `hashlib`, a
20-line RFC 6376 canonicaliser below) and refuses to write the file if any check fails.  The expected
hashes stored in the file come from that independent implementation.

Whitespace of RFC 6376 A.2 (the RFC text is indented; simple canonicalisation signs the indentation):
the printed signature verifies with continuation lines indented by six spaces, "com  [" with two
spaces in Received, and ONE space in "game. Are" (RFC 8463 prints two; relaxed collapses them)."""
import base64
import hashlib
import json
import os
import re

from cryptography.hazmat.primitives import hashes, serialization
from cryptography.hazmat.primitives.asymmetric import ed25519, padding

HERE = os.path.dirname(os.path.abspath(__file__))

RFC8463_RSA_P = ("MIGfMA0GCSqGSIb3DQEBAQUAA4GNADCBiQKBgQDkHlOQoBTzWRiGs5V6NpP3idY6Wk08a5qhdR6wy5bdOKb2jLQiY/J16JYi0Qvx/"
                 "byYzCNb3W91y3FutACDfzwQ/BC/e/8uBsCR+yz1Lxj+PL6lHvqMKrM3rG4hstT5QjvHO9PzoxZyVYLzBfO2EeC3Ip3G+2kryOTIKT+l/"
                 "K4w3QIDAQAB")
RFC8463_ED_P = "11qYAYKxCrfVS/7TyWQHOg7hcvPapiMlrwIaaPcHURo="
RFC8463_MSG = (
    "DKIM-Signature: v=1; a=ed25519-sha256; c=relaxed/relaxed;\r\n"
    " d=football.example.com; i=@football.example.com;\r\n"
    " q=dns/txt; s=brisbane; t=1528637909; h=from : to :\r\n"
    " subject : date : message-id : from : subject : date;\r\n"
    " bh=2jUSOH9NhtVGCQWNr9BrIAPreKQjO6Sn7XIkfJVOzv8=;\r\n"
    " b=/gCrinpcQOoIfuHNQIbq4pgh9kyIK3AQUdt9OdqQehSwhEIug4D11Bus\r\n"
    " Fa3bT3FY5OsU7ZbnKELq+eXdp1Q1Dw==\r\n"
    "DKIM-Signature: v=1; a=rsa-sha256; c=relaxed/relaxed;\r\n"
    " d=football.example.com; i=@football.example.com;\r\n"
    " q=dns/txt; s=test; t=1528637909; h=from : to : subject :\r\n"
    " date : message-id : from : subject : date;\r\n"
    " bh=2jUSOH9NhtVGCQWNr9BrIAPreKQjO6Sn7XIkfJVOzv8=;\r\n"
    " b=F45dVWDfMbQDGHJFlXUNB2HKfbCeLRyhDXgFpEL8GwpsRe0IeIixNTe3\r\n"
    " DhCVlUrSjV4BwcVcOF6+FF3Zo9Rpo1tFOeS9mPYQTnGdaSGsgeefOsk2Jz\r\n"
    " dA+L10TeYt9BgDfQNZtKdN1WO//KgIqXP7OdEFE4LjFYNcUxZQ4FADY+8=\r\n"
    "From: Joe SixPack <joe@football.example.com>\r\n"
    "To: Suzie Q <suzie@shopping.example.net>\r\n"
    "Subject: Is dinner ready?\r\n"
    "Date: Fri, 11 Jul 2003 21:00:37 -0700 (PDT)\r\n"
    "Message-ID: <20030712040037.46341.5F8J@football.example.com>\r\n"
    "\r\n"
    "Hi.\r\n"
    "\r\n"
    "We lost the game.  Are you hungry yet?\r\n"
    "\r\n"
    "Joe.\r\n")

RFC6376_RSA_P = ("MIGfMA0GCSqGSIb3DQEBAQUAA4GNADCBiQKBgQDwIRP/UC3SBsEmGqZ9ZJW3/DkMoGeLnQg1fWn7/zYtIxN2SnFCjxOCKG9v3b4jYfcTNh5ij"
                 "Ssq631uBItLa7od+v/RtdC2UzJ1lWT947qR+Rcac2gbto/NMqJ0fzfVjH4OuKhitdY9tf6mcwGjaNBcWToIMmPSPDdQPNUYckcQ2QIDAQAB")
IND = "      "
RFC6376_MSG = (
    "DKIM-Signature: v=1; a=rsa-sha256; s=brisbane; d=example.com;\r\n"
    + IND + "c=simple/simple; q=dns/txt; i=joe@football.example.com;\r\n"
    + IND + "h=Received : From : To : Subject : Date : Message-ID;\r\n"
    + IND + "bh=2jUSOH9NhtVGCQWNr9BrIAPreKQjO6Sn7XIkfJVOzv8=;\r\n"
    + IND + "b=AuUoFEfDxTDkHlLXSZEpZj79LICEps6eda7W3deTVFOk4yAUoqOB\r\n"
    + IND + "  4nujc7YopdG5dWLSdNg6xNAZpOPr+kHxt1IrE+NahM6L/LbvaHut\r\n"
    + IND + "  KVdkLLkpVaVVQPzeRDI009SO2Il5Lu7rDNH6mZckBdrIx0orEtZV\r\n"
    + IND + "  4bmp/YzhwvcubU4=;\r\n"
    "Received: from client1.football.example.com  [192.0.2.1]\r\n"
    + IND + "by submitserver.example.com with SUBMISSION;\r\n"
    + IND + "Fri, 11 Jul 2003 21:01:54 -0700 (PDT)\r\n"
    "From: Joe SixPack <joe@football.example.com>\r\n"
    "To: Suzie Q <suzie@shopping.example.net>\r\n"
    "Subject: Is dinner ready?\r\n"
    "Date: Fri, 11 Jul 2003 21:00:37 -0700 (PDT)\r\n"
    "Message-ID: <20030712040037.46341.5F8J@football.example.com>\r\n"
    "\r\n"
    "Hi.\r\n"
    "\r\n"
    "We lost the game. Are you hungry yet?\r\n"
    "\r\n"
    "Joe.\r\n")


# ---------------------------------------------------------------- independent RFC 6376 canonicaliser
def _headers(msg):
    hdr, body = msg.split("\r\n\r\n", 1)
    return re.split(r"\r\n(?=[^ \t])", hdr), body


def _relaxed_header(h):
    k, v = h.split(":", 1)
    v = re.sub(r"[ \t]+", " ", v.replace("\r\n", "")).strip(" ")
    return k.lower().rstrip(" \t") + ":" + v


def _relaxed_body(b):
    s = "\r\n".join(re.sub(r"[ \t]+", " ", l).rstrip(" ") for l in b.split("\r\n"))
    while s.endswith("\r\n\r\n"):
        s = s[:-2]
    return s


def _simple_body(b):
    while b.endswith("\r\n\r\n"):
        b = b[:-2]
    return b or "\r\n"


def preimages(msg, sig_index):
    hs, body = _headers(msg)
    sig = hs[sig_index]
    tags = sig.split(":", 1)[1].replace("\r\n", "")
    canon = re.search(r"\bc=\s*([a-z/]+)", tags).group(1)
    hc, bc = (canon.split("/") + ["simple"])[:2]
    names = [n.strip().lower() for n in re.search(r"\bh=([^;]*);", tags).group(1).split(":")]
    used, out = {}, []
    ch = _relaxed_header if hc == "relaxed" else (lambda h: h)
    for n in names:
        cands = [h for h in hs if h.split(":", 1)[0].strip().lower() == n]
        i = used.get(n, 0)
        if i < len(cands):
            out.append(ch(cands[len(cands) - 1 - i]) + "\r\n")
            used[n] = i + 1
    bval = re.search(r"[;\s]b=([^;]*)", sig, re.S).group(1)
    out.append(ch(sig.replace(bval, "")))
    cb = _relaxed_body(body) if bc == "relaxed" else _simple_body(body)
    return "".join(out).encode(), cb.encode(), base64.b64decode(re.sub(r"\s+", "", bval))


def pkcs1_der(spki_b64):
    pk = serialization.load_der_public_key(base64.b64decode(spki_b64))
    return pk, pk.public_bytes(serialization.Encoding.DER, serialization.PublicFormat.PKCS1)


def main():
    out = []
    b64 = lambda b: base64.b64encode(b).decode()
    sha = lambda b: hashlib.sha256(b).digest()

    # RFC 8463: signature 1 (second header) is the RSA one, signature 0 the Ed25519 one
    pk, der = pkcs1_der(RFC8463_RSA_P)
    hdr, body, sig = preimages(RFC8463_MSG, 1)
    pk.verify(sig, hdr, padding.PKCS1v15(), hashes.SHA256())            # raises when mis-transcribed
    assert b64(sha(body)) == "2jUSOH9NhtVGCQWNr9BrIAPreKQjO6Sn7XIkfJVOzv8="
    ehdr, ebody, esig = preimages(RFC8463_MSG, 0)
    ed = ed25519.Ed25519PublicKey.from_public_bytes(base64.b64decode(RFC8463_ED_P))
    ed.verify(esig, sha(ehdr))                                          # RFC 8463 §3: PureEdDSA over the SHA-256 digest
    common = dict(from_domain="football.example.com", raw_email=b64(RFC8463_MSG.encode()))
    out.append(dict(name="rfc8463_appendix_a_rsa_key", source="RFC 8463 Appendix A.2/A.3, selector test", **common,
                    key=b64(der), key_type="rsa", signature_index=1,
                    header_preimage=b64(hdr), canonical_body=b64(body),
                    expect=dict(verifies=True, body_hash=sha(body).hex(), header_hash=sha(hdr).hex(),
                                from_domain_hash=sha(b"football.example.com").hex(), public_key_hash=sha(der).hex())))
    raw_ed = base64.b64decode(RFC8463_ED_P)
    out.append(dict(name="rfc8463_appendix_a_ed25519_key", source="RFC 8463 Appendix A.2/A.3, selector brisbane", **common,
                    key=b64(raw_ed), key_type="ed25519", signature_index=0,
                    header_preimage=b64(ehdr), canonical_body=b64(ebody),
                    expect=dict(verifies=True, body_hash=sha(ebody).hex(), header_hash=sha(ehdr).hex(),
                                from_domain_hash=sha(b"football.example.com").hex(), public_key_hash=sha(raw_ed).hex())))

    pk, der = pkcs1_der(RFC6376_RSA_P)
    hdr, body, sig = preimages(RFC6376_MSG, 0)
    pk.verify(sig, hdr, padding.PKCS1v15(), hashes.SHA256())
    assert b64(sha(body)) == "2jUSOH9NhtVGCQWNr9BrIAPreKQjO6Sn7XIkfJVOzv8="
    out.append(dict(name="rfc6376_appendix_a2", source="RFC 6376 Appendix A.2, key of Appendix C", from_domain="example.com",
                    raw_email=b64(RFC6376_MSG.encode()), key=b64(der), key_type="rsa", signature_index=0,
                    header_preimage=b64(hdr), canonical_body=b64(body),
                    expect=dict(verifies=True, body_hash=sha(body).hex(), header_hash=sha(hdr).hex(),
                                from_domain_hash=sha(b"example.com").hex(), public_key_hash=sha(der).hex())))
    # negatives derived from the RFC messages: one flipped body byte, the other RFC's key
    bad = RFC8463_MSG.replace("Joe.\r\n", "Jof.\r\n")
    out.append(dict(name="rfc8463_body_flip", source="derived", from_domain="football.example.com", raw_email=b64(bad.encode()),
                    key=out[0]["key"], key_type="rsa", signature_index=1, expect=dict(verifies=False)))
    out.append(dict(name="rfc8463_wrong_key", source="derived", from_domain="football.example.com",
                    raw_email=b64(RFC8463_MSG.encode()), key=b64(der), key_type="rsa", signature_index=1,
                    expect=dict(verifies=False)))
    with open(os.path.join(HERE, "rfc_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out), "vectors; all RFC signatures verified with cryptography/OpenSSL")


if __name__ == "__main__":
    main()
