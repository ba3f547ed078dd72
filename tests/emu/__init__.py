"""ctypes front end of the host emulation of the kernel sources (tests/emu/emu_kernels.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libzkb_emu.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        alt = os.environ.get("ZKB_EMU_LIB")     # e.g. a -fsanitize=thread build (tools/emu_tsan.sh): lanes are real threads
        if not alt:
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        L = C.CDLL(alt or _LIB)
        L.emu_regex_compile.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                        C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_char_p, C.c_size_t]
        L.emu_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def sha256_batch(msgs, prefetch=False, rot=0):
    """The SHA-256 kernel source over a packed arena (garbage in the padding), sorted order.
    prefetch=True runs the double-buffered instantiation used for launches with few lanes; rot = 1 .. 3 the
    instantiations that issue that many rotate families on the FMA pipe."""
    n = len(msgs)
    off, cur = [], 0
    for m in msgs:
        off.append(cur)
        cur += (len(m) // 64 + 1) * 64
    arena = np.full(cur + 64, 0xAA, dtype=np.uint8)
    for o, m in zip(off, msgs):
        arena[o:o + len(m)] = np.frombuffer(m, dtype=np.uint8)
    offa = np.array(off, dtype=np.uint64)
    lena = np.array([len(m) for m in msgs], dtype=np.uint32)
    order = np.argsort(-lena.astype(np.int64), kind="stable").astype(np.uint32)
    dig = np.zeros((n, 8), dtype=np.uint32)
    if rot:
        lib().emu_sha256_batch_rot(_p(arena), _p(offa), _p(lena), _p(order), n, _p(dig), rot)
    else:
        (lib().emu_sha256_batch_prefetch if prefetch else lib().emu_sha256_batch)(_p(arena), _p(offa), _p(lena), _p(order), n, _p(dig))
    return [dig[i].astype(">u4").tobytes() for i in range(n)]


def rsa_verify(keys_der, digests, sigs, limbs, lanes, generic=False):
    """The RSA kernel source: one group of `lanes` threads per signature.  Returns flags (1 = ok)."""
    n = len(sigs)
    uniq = sorted(set(keys_der))
    keytab = np.zeros((len(uniq), 264), dtype=np.uint32)
    for i, d in enumerate(uniq):
        assert lib().emu_build_key_entry(d, len(d), _p(keytab[i])) == 0
    dig = np.zeros((n, 8), dtype=np.uint32)
    sigar = np.zeros(n * limbs, dtype=np.uint32)
    items = np.zeros((n, 4), dtype=np.uint32)
    for i in range(n):
        dig[i] = np.frombuffer(digests[i], dtype=">u4")
        s = int.from_bytes(sigs[i], "big")
        for j in range(limbs):
            sigar[i * limbs + j] = (s >> (32 * j)) & 0xFFFFFFFF
        items[i] = [i * limbs, uniq.index(keys_der[i]), i, i]
    flags = np.zeros(n, dtype=np.uint32)
    rc = lib().emu_rsa_verify(limbs, lanes, 1 if generic else 0, _p(sigar), _p(items), n, _p(keytab), _p(dig), _p(flags))
    assert rc == 0, "instantiation not compiled in the emulator"
    return [int(f & 1) for f in flags]


def regex_compile(pattern: bytes):
    f, b = C.c_void_p(), C.c_void_p()
    fl, bl = C.c_size_t(), C.c_size_t()
    err = C.create_string_buffer(256)
    rc = lib().emu_regex_compile(pattern, len(pattern), C.byref(f), C.byref(fl), C.byref(b), C.byref(bl), err, 256)
    if rc:
        raise ValueError(err.value.decode())
    r = (C.string_at(f, fl.value), C.string_at(b, bl.value))
    lib().emu_free(f)
    lib().emu_free(b)
    return r


def dfa_scan(fwd, bwd, hays, qp=False, use_smem=True, table_form=0):
    n = len(hays)
    off, cur = [], 0
    for h in hays:
        off.append(cur)
        cur += (len(h) // 64 + 1) * 64
    arena = np.full(cur + 64, 0x3D, dtype=np.uint8)  # '=' garbage in the padding: must never be read as data
    for o, h in zip(off, hays):
        arena[o:o + len(h)] = np.frombuffer(h, dtype=np.uint8)
    offa = np.array(off, dtype=np.uint64)
    lena = np.array([len(h) for h in hays], dtype=np.uint32)
    out = np.zeros((n, 4), dtype=np.uint32)
    rc = lib().emu_dfa_scan(fwd, len(fwd), bwd, len(bwd), _p(arena), _p(offa), _p(lena), n, 1 if qp else 0,
                            1 if use_smem else 0, _p(out), table_form)
    assert rc == 0
    return out


def canon_bodies(bodies, relaxed=True, l=None):
    """The device canonicalisation kernel source over bodies placed at odd offsets of a span buffer."""
    n = len(bodies)
    off, cur = [], 5
    for b in bodies:
        off.append(cur)
        cur += len(b) + 3   # deliberately unaligned, bodies adjacent to foreign bytes
    span = np.full(cur + 64, 0x20, dtype=np.uint8)   # SP garbage around the bodies
    for o, b in zip(off, bodies):
        span[o:o + len(b)] = np.frombuffer(b, dtype=np.uint8)
    slot, c2 = [], 0
    for b in bodies:
        slot.append(c2)
        c2 += ((len(b) + 2) // 64 + 1) * 64
    arena = np.full(c2 + 64, 0xEE, dtype=np.uint8)
    offa = np.array(off, dtype=np.uint64)
    lena = np.array([len(b) for b in bodies], dtype=np.uint32)
    flags = np.full(n, (1 if relaxed else 0) | (2 if l is not None else 0), dtype=np.uint32)
    lval = np.full(n, l if l is not None else 0, dtype=np.uint32)
    slota = np.array(slot, dtype=np.uint64)
    out_len = np.zeros(n, dtype=np.uint32)
    lib().emu_canon_body(_p(span), _p(offa), _p(lena), _p(flags), _p(lval), n, _p(arena), _p(slota), _p(out_len))
    return [arena[s:s + int(k)].tobytes() for s, k in zip(slot, out_len)]
