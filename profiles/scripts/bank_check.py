"""Brute-force bank-conflict check of the shared-memory exchange layouts of the one-warp FFT-512
(c-ofdm_b200/csrc/fft512w.cuh).  A 128-bit warp access is served in 4 phases of 8 lanes; a phase is
conflict-free when its 8 lanes touch 8 different 16-byte bank groups (address/16 mod 8) or the same
address.  64-bit accesses: 2 phases of 16 lanes, 16 different 8-byte groups.  Run: python bank_check.py"""
import itertools


def phases128(addrs):            # addrs: float2 index per lane (16-byte aligned pairs)
    worst = 1
    for p in range(4):
        units = {}
        for l in range(8 * p, 8 * p + 8):
            u = addrs[l] // 2
            units.setdefault(u % 8, set()).add(u)
        worst = max(worst, max(len(v) for v in units.values()))
    return worst


def phases64(addrs):
    worst = 1
    for p in range(2):
        units = {}
        for l in range(16 * p, 16 * p + 16):
            units.setdefault(addrs[l] % 16, set()).add(addrs[l])
        worst = max(worst, max(len(v) for v in units.values()))
    return worst


# ---- E1: writer lane l holds t = 2l, 2l+1 (t = 8 n2 + n3); element (k1, t) ----
def e1(k1, t):
    return 64 * k1 + (t ^ ((k1 & 1) << 3))


# ---- E2: element (k1, k2, n3) ----
def e2(k1, k2, n3):
    return 64 * k1 + 8 * (k2 ^ (k1 & 1)) + ((((n3 >> 1) ^ ((k2 >> 1) & 3)) << 1) | (n3 & 1))


def main():
    w = 1
    # E1 write: fixed k1, lane l writes pair at t = 2l
    for k1 in range(8):
        w = max(w, phases128([e1(k1, 2 * l) for l in range(32)]))
    # E1 read: lane l' = 4 k1 + j reads for n2: pair at t = 8 n2 + 2 j
    for n2 in range(8):
        w = max(w, phases128([e1(l >> 2, 8 * n2 + 2 * (l & 3)) for l in range(32)]))
    print("E1 worst ways:", w)
    # injectivity
    assert len({e1(k1, t) for k1 in range(8) for t in range(64)}) == 512
    assert all(e1(k1, 2 * u) % 2 == 0 and e1(k1, 2 * u + 1) == e1(k1, 2 * u) + 1 for k1 in range(8) for u in range(32))
    w = 1
    # E2 write: lane l' = 4 k1 + j writes for k2: pair n3 = 2j
    for k2 in range(8):
        w = max(w, phases128([e2(l >> 2, k2, 2 * (l & 3)) for l in range(32)]))
    # E2 read: lane l'' = k2 + 8 a reads for b in 0,1 and jj: pair (k1 = 2a + b, k2, n3 = 2 jj)
    for b in range(2):
        for jj in range(4):
            w = max(w, phases128([e2(2 * (l >> 3) + b, l & 7, 2 * jj) for l in range(32)]))
    print("E2 worst ways:", w)
    assert len({e2(k1, k2, n3) for k1 in range(8) for k2 in range(8) for n3 in range(8)}) == 512
    assert all(e2(k1, k2, 2 * u) % 2 == 0 and e2(k1, k2, 2 * u + 1) == e2(k1, k2, 2 * u) + 1
               for k1 in range(8) for k2 in range(8) for u in range(4))


if __name__ == "__main__":
    main()
