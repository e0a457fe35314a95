"""Numpy emulation of the K1 DMMA tiling (fragment tables + sliding operand windows), used to check the
index arithmetic of csrc/banded_kernel.cu on the CPU.  Lane (gid, q) = (lane>>2, lane&3) owns chain gid and
times 8J+q, 8J+q+4 of tile J; DMMA m8n8k4: D[m][n] += sum_k A[m][k] B[k][n] with A-frag lane -> A[gid][q],
B-frag lane -> B[q][gid], C-frag lane -> C[gid][2q], C[gid][2q+1]."""
import numpy as np


def geom(n, b):
    HB = (b + 3) // 4
    NCH = 2 * HB + 2
    LAGT = (HB + 1) // 2
    WN = 2 * LAGT + 2 + HB
    NT = (n + 7) // 8
    return HB, NCH, LAGT, WN, NT


def build_fragtab(band, b, n, transpose):
    """band: (2b+1, n) diagonal-major, band[b + (j-i), i] = A[i, j].  Returns frag[NT][NCH][32]."""
    HB, NCH, LAGT, WN, NT = geom(n, b)
    ft = np.zeros((NT, NCH, 32))
    for J in range(NT):
        for hh in range(NCH):
            for lane in range(32):
                gid, q = lane >> 2, lane & 3
                o = 8 * J + (gid >> 1) + 4 * (gid & 1)          # output time of C slot n = gid
                i = 4 * (2 * J - HB + hh) + q                   # input time of k slot q
                if o < n and 0 <= i < n and abs(i - o) <= b:
                    ft[J, hh, lane] = band[b + (o - i), i] if transpose else band[b + (i - o), o]
    return ft


def dmma(c0, c1, a, bfrag):
    """a, bfrag: (32,) per-lane operands; c0, c1: (32,) accumulators."""
    A = np.zeros((8, 4)); B = np.zeros((4, 8))
    for lane in range(32):
        A[lane >> 2, lane & 3] = a[lane]
        B[lane & 3, lane >> 2] = bfrag[lane]
    C = A @ B
    for lane in range(32):
        c0[lane] += C[lane >> 2, 2 * (lane & 3)]
        c1[lane] += C[lane >> 2, 2 * (lane & 3) + 1]


def sweep_matvec(ft, X, n, b):
    """X: (8 chains, n).  Returns Y (8, n) = X @ A^T via the windowed sweep."""
    HB, NCH, LAGT, WN, NT = geom(n, b)
    xw = np.zeros((WN, 32))
    Y = np.zeros((8, NT * 8))
    for s in range(NT + LAGT):
        xw[:-2] = xw[2:]
        v0 = np.zeros(32); v1 = np.zeros(32)
        if s < NT:
            for lane in range(32):
                gid, q = lane >> 2, lane & 3
                t0, t1 = 8 * s + q, 8 * s + q + 4
                v0[lane] = X[gid, t0] if t0 < n else 0.0
                v1[lane] = X[gid, t1] if t1 < n else 0.0
        xw[WN - 2], xw[WN - 1] = v0, v1
        Ja = s - LAGT
        if 0 <= Ja < NT:
            c0 = np.zeros(32); c1 = np.zeros(32)
            for hh in range(NCH):
                dmma(c0, c1, xw[hh], ft[Ja, hh])
            for lane in range(32):
                gid, q = lane >> 2, lane & 3
                Y[gid, 8 * Ja + q] = c0[lane]; Y[gid, 8 * Ja + q + 4] = c1[lane]
            # own tile must sit at window slots HB, HB+1
            for lane in range(32):
                gid, q = lane >> 2, lane & 3
                if 8 * Ja + q < n: assert xw[HB, lane] == X[gid, 8 * Ja + q]
                if 8 * Ja + q + 4 < n: assert xw[HB + 1, lane] == X[gid, 8 * Ja + q + 4]
    return Y[:, :n]


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for n, b in [(201, 20), (37, 5), (16, 4), (9, 1), (3, 1), (5, 0), (50, 16), (41, 32), (6, 2), (1, 0)]:
        A = rng.normal(size=(n, n))
        i, j = np.indices((n, n)); A[np.abs(i - j) > b] = 0
        band = np.zeros((2 * b + 1, n))
        for off in range(-b, b + 1):
            for r in range(max(0, -off), min(n, n - off)):
                band[b + off, r] = A[r, r + off]
        X = rng.normal(size=(8, n))
        Y = sweep_matvec(build_fragtab(band, b, n, False), X, n, b)
        Yt = sweep_matvec(build_fragtab(band, b, n, True), X, n, b)
        print(n, b, geom(n, b), np.max(np.abs(Y - X @ A.T)), np.max(np.abs(Yt - X @ A)))
