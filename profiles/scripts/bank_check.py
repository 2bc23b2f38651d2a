"""Brute-force bank-conflict check of the shared-memory exchange layouts of the one-warp FFT-512
(c-ofdm_b200/csrc/fft512w.cuh).  A 128-bit warp access is served in 4 phases of 8 lanes; a phase is
conflict-free when its 8 lanes touch 8 different 16-byte bank groups (address/16 mod 8).  A 64-bit
access is served in 2 phases of 16 lanes that must touch 16 different 8-byte groups (address/8 mod 16).
Addresses below are float2 (8-byte) slot indices.  Run: python bank_check.py"""


def phases128(addrs):
    worst = 1
    for p in range(4):
        units = {}
        for l in range(8 * p, 8 * p + 8):
            assert addrs[l] % 2 == 0
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


def e1(k1, t):                       # plane B (t odd) starts at slot 324
    return 324 * (t & 1) + 40 * k1 + (t >> 1)


def e2(k1, k2, n3):                  # plane A: n3 in {0,1,4,5}; plane B (slot 272 on): n3 in {2,3,6,7}
    return 272 * ((n3 >> 1) & 1) + 34 * k2 + 4 * k1 + (n3 & 1) + 2 * (n3 >> 2)


def main():
    w = 1
    for k1 in range(8):              # E1 write: lane l stores t = 2l (plane A) and t = 2l + 1 (plane B)
        w = max(w, phases64([e1(k1, 2 * l) for l in range(32)]), phases64([e1(k1, 2 * l + 1) for l in range(32)]))
    n3a = lambda j: (j & 1) + 4 * (j >> 1)
    for n2 in range(8):              # E1 read: lane l' = 4 k1 + j reads (t, t + 2), t = 8 n2 + n3a(j)
        addrs = [e1(l >> 2, 8 * n2 + n3a(l & 3)) for l in range(32)]
        assert all(e1(l >> 2, 8 * n2 + n3a(l & 3) + 2) == addrs[l] + 1 for l in range(32))
        w = max(w, phases128(addrs))
    print("E1 worst ways:", w)
    assert len({e1(k1, t) for k1 in range(8) for t in range(64)}) == 512
    assert max(e1(k1, t) for k1 in range(8) for t in range(64)) < 644
    w = 1
    for k2 in range(8):              # E2 write: lane l' = 4 k1 + j stores n3a(j) (plane A) and n3a(j) + 2 (plane B)
        w = max(w, phases64([e2(l >> 2, k2, n3a(l & 3)) for l in range(32)]), phases64([e2(l >> 2, k2, n3a(l & 3) + 2) for l in range(32)]))
    for b in range(2):               # E2 read: lane l'' = k2 + 8 a reads pairs (n3, n3 + 1), n3 in {0, 4, 2, 6}, of k1 = 2a + b
        for n3 in (0, 4, 2, 6):
            addrs = [e2(2 * (l >> 3) + b, l & 7, n3) for l in range(32)]
            assert all(e2(2 * (l >> 3) + b, l & 7, n3 + 1) == addrs[l] + 1 for l in range(32))
            w = max(w, phases128(addrs))
    print("E2 worst ways:", w)
    assert len({e2(k1, k2, n3) for k1 in range(8) for k2 in range(8) for n3 in range(8)}) == 512
    assert max(e2(k1, k2, n3) for k1 in range(8) for k2 in range(8) for n3 in range(8)) < 644


if __name__ == "__main__":
    main()
