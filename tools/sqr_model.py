"""Model of the dedicated Montgomery squaring of rsa.cuh (T = 4 lanes x L = 16 limbs): which chunk products each lane
computes, where their limbs go in shared memory, which 8-limb pieces every owner lane sums (the table SQ_TAB that
rsa.cuh carries), and the reduction that follows.  Run it to regenerate the table and to check the scheme on random
operands with Python integers:   python tools/sqr_model.py [--emit]"""
import random
import sys

L, T = 16, 4
B = 1 << 32
AREA_D, AREA_F, AREA_H = 0, 32, 72      # limb offsets of the three frames inside a producer lane's area
AREA = 104                              # limbs per producer lane: D 32 | F 40 (33 used) | H 32 (25 used)


def limbs(v, n):
    return [(v >> (32 * i)) & (B - 1) for i in range(n)]


def frames(a):
    """a: 64 limbs.  Returns per producer lane p the three frames as integers and their global limb offsets."""
    A = [sum(a[16 * p + i] << (32 * i) for i in range(16)) for p in range(4)]
    out = []
    for p in range(4):
        q, h = (p + 1) % 4, (p + 2) % 4
        D = A[p] * A[p]
        F = 2 * A[p] * A[q]
        # the pair (p, p+2) is split between its two lanes: the lower lane multiplies its chunk by the LOW half of the
        # other chunk, the upper lane multiplies the other chunk (fetched into registers) by its own HIGH half
        if p < 2:
            H, hoff = 2 * A[p] * (A[h] & ((1 << 256) - 1)), 16 * (p + h)
        else:
            H, hoff = 2 * A[h] * (A[p] >> 256), 16 * (p + h) + 8
        out.append(((D, 32 * p), (F, 16 * (p + q)), (H, hoff)))
    return out


def table():
    """SQ_TAB[lane][22]: for the 6 + 5 + 6 + 5 piece slots (lo0, lo1, hi0, hi1) the limb offset of the 8-limb piece in
    the signature's shared-memory region, or 0xFFFF."""
    cover = {o: [] for o in range(16)}
    for p in range(4):
        q, h = (p + 1) % 4, (p + 2) % 4
        for j in range(4):
            cover[4 * p + j].append(p * AREA + AREA_D + 8 * j)
        for j in range(5):
            cover[2 * (p + q) + j].append(p * AREA + AREA_F + 8 * j)
        hs = 2 * (p + h) + (0 if p < 2 else 1)
        for j in range(4):
            cover[hs + j].append(p * AREA + AREA_H + 8 * j)
    slots = (6, 5, 6, 5)
    tab = []
    for r in range(4):
        row = []
        for d, o in enumerate((2 * r, 2 * r + 1, 8 + 2 * r, 9 + 2 * r)):
            assert len(cover[o]) <= slots[d], (r, d, len(cover[o]))
            row += cover[o] + [0xFFFF] * (slots[d] - len(cover[o]))
        tab.append(row)
    return tab


def check(n_trials=200):
    tab = table()
    rng = random.Random(1)
    for _ in range(n_trials):
        x = rng.getrandbits(2048) if rng.random() < 0.8 else (1 << 2048) - 1 - rng.getrandbits(40)
        a = limbs(x, 64)
        smem = [0] * (4 * AREA)
        for p, fr in enumerate(frames(a)):
            (D, _), (F, _), (H, _) = fr
            smem[p * AREA + AREA_D:p * AREA + AREA_D + 32] = limbs(D, 32)
            smem[p * AREA + AREA_F:p * AREA + AREA_F + 40] = limbs(F, 40)
            smem[p * AREA + AREA_H:p * AREA + AREA_H + 32] = limbs(H, 32)
            assert F < 1 << (32 * 33) and H < 1 << (32 * 25)
        total = 0
        for r in range(4):
            k = 0
            for d, cnt in enumerate((6, 5, 6, 5)):
                o = (2 * r, 2 * r + 1, 8 + 2 * r, 9 + 2 * r)[d]
                acc = 0
                for _i in range(cnt):
                    off = tab[r][k]; k += 1
                    if off != 0xFFFF:
                        acc += sum(smem[off + t] << (32 * t) for t in range(8))
                total += acc << (256 * o)
        assert total == x * x, "combine"
        # reduction: window = low 64 limbs, the high limbs enter one per step at the top
        n = rng.getrandbits(2048) | (1 << 2047) | 1
        n0inv = (-pow(n, -1, B)) % B
        t = total
        w = t & ((1 << 2048) - 1)
        hi = limbs(t >> 2048, 64)
        for i in range(64):
            m = ((w & (B - 1)) * n0inv) & (B - 1)
            w = (w + m * n) >> 32
            w += hi[i] << (32 * 63)
        assert w % n == (x * x * pow(1 << 2048, -1, n)) % n and w < (1 << 2048) + n, "reduce"
    return True


if __name__ == "__main__":
    assert check()
    print("model ok: frames + 22-slot piece table reproduce x^2; fed reduction gives x^2 R^-1 mod n")
    if "--emit" in sys.argv:
        for row in table():
            print("  {" + ", ".join("0x%04x" % v for v in row) + "},")
