"""Synthetic workload generator — ctypes front end of workload/zk_gen.c (BENCH / TEST INFRASTRUCTURE, neither product
code nor the oracle): large seeded pools of
DKIM-signed synthetic mail as flat numpy arrays, plus zero-copy views for the engine's C ABI
(zkb_email_view records) and for the oracle's batch driver (zo_email records)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libzk_gen.so")
DER_STRIDE, DOM_STRIDE = 512, 64
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "zk_gen.c")):
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        L = C.CDLL(_LIB)
        L.zg_keys_create.restype = C.c_void_p
        L.zg_keys_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.zg_keys_destroy.argtypes = [C.c_void_p]
        L.zg_raw_bound.restype = C.c_size_t
        L.zg_raw_bound.argtypes = [C.c_size_t]
        L.zg_compact.restype = C.c_size_t
        L.zg_compact.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
        L.zg_generate.argtypes = [C.c_void_p, C.c_uint64, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        _lib = L
    return _lib


class KeyPool:
    def __init__(self, n2048: int, n1024: int = 0, threads: int = 0):
        self.n = n2048 + n1024
        self.n2048, self.n1024 = n2048, n1024
        self.der = np.zeros((self.n, DER_STRIDE), dtype=np.uint8)
        self.der_len = np.zeros(self.n, dtype=np.uint32)
        self.handle = lib().zg_keys_create(n2048, n1024, threads or (os.cpu_count() or 1),
                                           self.der.ctypes.data, self.der_len.ctypes.data)
        assert self.handle, "key generation failed"
        self.dom = np.zeros((self.n, DOM_STRIDE), dtype=np.uint8)
        self.dom_len = np.zeros(self.n, dtype=np.uint32)
        for i in range(self.n):
            d = f"d{i}.example.com".encode()
            self.dom[i, : len(d)] = np.frombuffer(d, dtype=np.uint8)
            self.dom_len[i] = len(d)

    def key_der(self, i: int) -> bytes:
        return self.der[i, : self.der_len[i]].tobytes()

    def close(self):
        if self.handle:
            lib().zg_keys_destroy(self.handle)
            self.handle = None


class MailPool:
    """n synthetic emails in one flat buffer."""

    _RSA = np.frombuffer(b"rsa\0", dtype=np.uint8).copy()

    def __init__(self, keys: KeyPool, n: int, body_len, seed: int = 0xD1C1, neg_fraction: float = 0.0,
                 token: bool = False, qp_percent: int = 0, threads: int = 0, key_idx=None):
        rng = np.random.default_rng(seed)
        self.keys, self.n = keys, n
        self.body_len = (np.full(n, body_len, dtype=np.uint32) if np.isscalar(body_len)
                         else np.asarray(body_len, dtype=np.uint32))
        self.key_idx = (rng.integers(0, keys.n, size=n).astype(np.uint32) if key_idx is None
                        else np.asarray(key_idx, dtype=np.uint32))
        # negatives: 1 body flip, 2 signature flip, 3 wrong key (verifier is handed another key)
        self.neg_kind = np.zeros(n, dtype=np.uint8)
        n_neg = int(round(n * neg_fraction))
        if n_neg:
            idx = rng.choice(n, size=n_neg, replace=False)
            self.neg_kind[idx] = (np.arange(n_neg) % 3 + 1).astype(np.uint8)
        bound = self.body_len.astype(np.uint64) + 1400
        self.raw_off = np.zeros(n, dtype=np.uint64)
        if n > 1:
            self.raw_off[1:] = np.cumsum((bound[:-1] + 63) // 64 * 64, dtype=np.uint64)
        total = int(self.raw_off[-1] + bound[-1]) if n else 0
        self.raw = np.zeros(total + 64, dtype=np.uint8)
        self.raw_len = np.zeros(n, dtype=np.uint32)
        rc = lib().zg_generate(keys.handle, seed, n, self.body_len.ctypes.data, self.key_idx.ctypes.data,
                               self.neg_kind.ctypes.data, self.raw_off.ctypes.data, self.raw.ctypes.data,
                               self.raw_len.ctypes.data, 1 if token else 0, qp_percent,
                               threads or (os.cpu_count() or 1))
        assert rc == 0, "generator failed"
        # pack the arena the way a mail spool would be: messages back to back, 64-byte aligned starts
        used = lib().zg_compact(self.raw.ctypes.data, self.raw_off.ctypes.data, self.raw_len.ctypes.data, n, 64)
        self.raw = self.raw[: used + 64]
        # the key each email is verified with ("wrong key" negatives get a same-size neighbour)
        self.verify_key = self.key_idx.copy()
        wrong = self.neg_kind == 3
        if wrong.any():
            k = self.key_idx[wrong]
            lo = np.where(k < keys.n2048, 0, keys.n2048)
            cnt = np.where(k < keys.n2048, max(keys.n2048, 1), max(keys.n1024, 1))
            self.verify_key[wrong] = (lo + (k - lo + 1) % cnt).astype(np.uint32)
            if (self.verify_key[wrong] == k).any():  # pool of one key: cannot make a wrong-key case
                self.neg_kind[wrong & (self.verify_key == self.key_idx)] = 0

    def expected_ok(self) -> np.ndarray:
        return self.neg_kind == 0

    def engine_views(self, order=None) -> np.ndarray:
        """(n, 8) uint64 zkb_email_view records (include/zkemail_b200.h)."""
        sel = np.arange(self.n) if order is None else np.asarray(order)
        v = np.empty((len(sel), 8), dtype=np.uint64)
        v[:, 0] = self.keys.dom.ctypes.data + self.key_idx[sel].astype(np.uint64) * DOM_STRIDE
        v[:, 1] = self.keys.dom_len[self.key_idx[sel]]
        v[:, 2] = self.raw.ctypes.data + self.raw_off[sel]
        v[:, 3] = self.raw_len[sel]
        v[:, 4] = self.keys.der.ctypes.data + self.verify_key[sel].astype(np.uint64) * DER_STRIDE
        v[:, 5] = self.keys.der_len[self.verify_key[sel]]
        v[:, 6] = self._RSA.ctypes.data
        v[:, 7] = 3
        return v

    def oracle_views(self, order=None) -> np.ndarray:
        """(n, 7) uint64 zo_email records (oracle/zk_oracle.h)."""
        ev = self.engine_views(order)
        return np.ascontiguousarray(ev[:, :7])

    def email(self, i: int):
        """One email as the API struct (for small cross-checks)."""
        import zkemail_rs_b200 as z
        raw = self.raw[int(self.raw_off[i]): int(self.raw_off[i]) + int(self.raw_len[i])].tobytes()
        k = int(self.verify_key[i])
        dom = self.keys.dom[int(self.key_idx[i]), : self.keys.dom_len[int(self.key_idx[i])]].tobytes().decode()
        return z.Email(dom, raw, z.PublicKey(self.keys.key_der(k), "rsa"))
