"""Instruction-level Python model of the distributed (T lanes x L limbs) Montgomery multiply used
by zkemail.rs_b200/csrc/rsa.cuh.  Mirrors the even/odd IMAD.WIDE accumulator dataflow, the carry
flag, and the warp shuffles, so the algorithm can be checked against Python big ints before any
GPU time is spent.  Development aid only (not imported by product or tests)."""
import random, sys
M32 = 0xFFFFFFFF
maxtop = 0

class Lane:
    def __init__(s, L):
        s.L = L; s.X = [0]*(L+2); s.Y = [0]*(L+2); s.cf = 0
    # PTX-like ops on this lane's carry flag
    def mad_wide(s, arr, i, a, b, src, j, cin):
        """(arr[i],arr[i+1]) = a*b + (src[j],src[j+1]) + (cf if cin) ; sets cf"""
        add = src[j] | (src[j+1] << 32)
        v = a*b + add + (s.cf if cin else 0)
        s.cf = v >> 64
        assert s.cf <= 1
        arr[i] = v & M32; arr[i+1] = (v >> 32) & M32
    def addc(s, arr, i, val, cin, cout=True):
        v = arr[i] + val + (s.cf if cin else 0)
        arr[i] = v & M32
        if cout: s.cf = v >> 32
        else:
            global maxtop
            assert (v >> 32) == 0, "overflow out of top limb"

def mont_mul(a, b, n, n0inv, T, L):
    """a,b,n: integers < R. returns canonical integer result (< R) after conditional subtract."""
    global maxtop
    LIMBS = T*L; R = 1 << (32*LIMBS)
    limbs = lambda v: [(v >> (32*i)) & M32 for i in range(LIMBS)]
    al, bl, nl = limbs(a), limbs(b), limbs(n)
    lanes = [Lane(L) for _ in range(T)]
    A = [al[p*L:(p+1)*L] for p in range(T)]
    N = [nl[p*L:(p+1)*L] for p in range(T)]
    # arrays: lane.X is "Xp" (pre-shift even array of previous iteration), lane.Y is "Yp"
    for i in range(LIMBS):
        bi = bl[i]                       # shfl broadcast of b[i%L] from lane i//L
        for p, ln in enumerate(lanes):
            Xp, Yp = ln.X, ln.Y
            # A: Yp[0] += Xp[1]
            ln.addc(Yp, 0, Xp[1], cin=False)
            # B: Ynew (in place in Xp): Xp[2k..] = a[2k+1]*bi + Xp[2k+2..] + carry
            for k in range(L//2):
                ln.mad_wide(Xp, 2*k, A[p][2*k+1], bi, Xp, 2*k+2, cin=True)
            Xp[L] = ln.cf; Xp[L+1] = 0; ln.cf = 0           # addc Xp[L] = 0+0+cf
            # C: Yp += a_even*bi
            for k in range(L//2):
                ln.mad_wide(Yp, 2*k, A[p][2*k], bi, Yp, 2*k, cin=(k > 0))
            ln.addc(Yp, L, 0, cin=True); ln.addc(Yp, L+1, 0, cin=True, cout=False)
            ln.X, ln.Y = Yp, Xp              # role swap
        m = (lanes[0].X[0] * n0inv) & M32    # lane 0 computes, shfl broadcast
        for p, ln in enumerate(lanes):
            X, Y = ln.X, ln.Y
            for k in range(L//2):
                ln.mad_wide(X, 2*k, N[p][2*k], m, X, 2*k, cin=(k > 0))
            ln.addc(X, L, 0, cin=True); ln.addc(X, L+1, 0, cin=True, cout=False)
            for k in range(L//2):
                ln.mad_wide(Y, 2*k, N[p][2*k+1], m, Y, 2*k, cin=(k > 0))
            ln.addc(Y, L, 0, cin=True); ln.addc(Y, L+1, 0, cin=True, cout=False)
        assert lanes[0].X[0] == 0
        lows = [ln.X[0] for ln in lanes]
        for p, ln in enumerate(lanes):
            recv = lows[p+1] if p+1 < T else 0   # shfl_down
            ln.addc(ln.X, L, recv, cin=False); ln.addc(ln.X, L+1, 0, cin=True, cout=False)
            maxtop = max(maxtop, ln.X[L+1], ln.Y[L+1], ln.Y[L])
    # merge: res = Y + X[1] + (X>>64)<<32  per lane  -> res[0..L+1]
    total = 0
    res_l = []
    for p, ln in enumerate(lanes):
        X, Y = ln.X, ln.Y
        res = [0]*(L+3)
        v = Y[0] + X[1]; res[0] = v & M32; c = v >> 32
        for j in range(1, L+2):
            xv = X[j+1] if j+1 < L+2 else 0
            v = Y[j] + xv + c; res[j] = v & M32; c = v >> 32
        assert c == 0
        res_l.append(res)
    # inter-lane resolve phase 1: add hi (res[L], res[L+1]) of lane p-1 into lane p limbs 0,1 ...
    his = [(r[L], r[L+1]) for r in res_l]
    g = []; Pm = []
    for p in range(T):
        r = res_l[p]
        h0, h1 = his[p-1] if p > 0 else (0, 0)
        v = r[0] + h0; r[0] = v & M32; c = v >> 32
        v = r[1] + h1 + c; r[1] = v & M32; c = v >> 32
        for j in range(2, L):
            v = r[j] + c; r[j] = v & M32; c = v >> 32
        g.append(c); Pm.append(all(x == M32 for x in r[:L]))
    # phase 2: single-bit carries via generate/propagate (ballot trick)
    G = sum(gb << p for p, gb in enumerate(g)); Pb = sum((1 if pb else 0) << p for p, pb in enumerate(Pm))
    # carry INTO lane p+1 = carry out of bit p of (G + (G|P)) ; cin vector = ((G+(G|P)) ^ (G|P) ^ G) >> ... derive directly:
    s = G + (G | Pb)
    # carry out of bit p in this sum:
    cins = [0]*(T+1)
    for p in range(T):
        gp = (G >> p) & 1; pp = (Pb >> p) & 1
        cins[p+1] = gp | (pp & cins[p])
    # check the add trick reproduces cins:
    x = G ^ (G | Pb)            # a^b per bit = p (when g,p exclusive)
    carr = s ^ x                # carry-in per bit
    for p in range(T+1):
        assert ((carr >> p) & 1) == cins[p], (bin(G), bin(Pb), p)
    for p in range(T):
        r = res_l[p]; c = cins[p]
        for j in range(L):
            v = r[j] + c; r[j] = v & M32; c = v >> 32
    ov = his[T-1][0] + (his[T-1][1] << 32) + cins[T]
    assert ov <= 1, ov
    t = sum(sum(r[j] << (32*j) for j in range(L)) << (32*L*p) for p, r in enumerate(res_l))
    full = t + (ov << (32*LIMBS))
    if ov: t = (full - n)
    assert 0 <= t < R
    return t

def test(T, L, trials=20, adversarial=True):
    LIMBS = T*L; R = 1 << (32*LIMBS)
    rnd = random.Random(1234 + T*100 + L)
    for it in range(trials):
        if adversarial and it < 4:
            n = R - 1 - 2*rnd.randrange(4) if it % 2 == 0 else (1 << (32*LIMBS-1)) + 1 + 2*rnd.randrange(1000)
            a = R - 1 - rnd.randrange(3); b = R - 1 - rnd.randrange(3)
        else:
            n = rnd.getrandbits(32*LIMBS) | 1 | (1 << (32*LIMBS-1)) if it % 3 else (rnd.getrandbits(32*LIMBS - rnd.randrange(1, 70)) | 1)
            a = rnd.randrange(R); b = rnd.randrange(R)
        n0inv = (-pow(n, -1, 1 << 32)) & M32
        t = mont_mul(a, b, n, n0inv, T, L)
        assert (t * R - a*b) % n == 0, "wrong residue"
    # modexp e=65537
    n = rnd.getrandbits(32*LIMBS) | 1 | (1 << (32*LIMBS-1)); s = rnd.randrange(n)
    n0inv = (-pow(n, -1, 1 << 32)) & M32; RR = (R*R) % n
    x = mont_mul(s, RR, n, n0inv, T, L)
    for _ in range(16): x = mont_mul(x, x, n, n0inv, T, L)
    x = mont_mul(x, s, n, n0inv, T, L)
    if x >= n: x -= n
    assert x == pow(s, 65537, n)
    print(f"T={T} L={L} ok, max top limb seen {maxtop}")

if __name__ == "__main__":
    for T, L in [(1, 8), (2, 4), (4, 2), (4, 4), (2, 16), (4, 16), (8, 8), (2, 32)]:
        test(T, L, trials=12)
