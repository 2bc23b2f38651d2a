"""Brute-force bank-conflict check of the shared-memory exchange of the one-warp FFT-512
(c-ofdm_b200/csrc/fft512w.cuh, radix 16 x 16 x 2).  A 128-bit warp access is served in 4 phases of 8
lanes; a phase is conflict-free when its 8 lanes touch 8 different 16-byte bank groups (address/16
mod 8).  A 64-bit access is served in 2 phases of 16 lanes that must touch 16 different 8-byte groups
(address/8 mod 16).  Addresses below are float2 (8-byte) slot indices.  Run: python bank_check.py"""


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


def ex(g, k1, m):                    # element A'[k1; 2m + g]
    return 296 * g + 18 * k1 + m


def main():
    w = 1
    for k1 in range(16):             # write: lane l = 2m + g stores its k1-th output
        w = max(w, phases64([ex(l & 1, k1, l >> 1) for l in range(32)]))
    print("exchange write worst ways:", w)
    w = 1
    for mm in range(8):              # read: lane l' = 2 k1 + g loads (m, m + 1) = (2 mm, 2 mm + 1)
        addrs = [ex(l & 1, l >> 1, 2 * mm) for l in range(32)]
        assert all(ex(l & 1, l >> 1, 2 * mm + 1) == addrs[l] + 1 for l in range(32))
        w = max(w, phases128(addrs))
    print("exchange read worst ways:", w)
    assert len({ex(g, k1, m) for g in range(2) for k1 in range(16) for m in range(16)}) == 512
    assert max(ex(g, k1, m) for g in range(2) for k1 in range(16) for m in range(16)) < 584
    # coarse 640-point transform of the acquire kernel: pass-1 output 10 j + q at slot 11 j + q (lanes j, j + 32)
    w = 1
    for q in range(10):
        w = max(w, phases64([11 * l + q for l in range(32)]), phases64([11 * (l + 32) + q for l in range(32)]))
    print("coarse pass-1 write worst ways:", w)


if __name__ == "__main__":
    main()
