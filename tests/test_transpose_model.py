"""CPU model of k_transpose_reg's two index tricks (panacus_b200/csrc/pgx_gm.cu), checked exhaustively:
  * the in-register 32 x 32 bit transpose: two PRMT stages (selectors 0x5410 / 0x7632 and 0x6240 / 0x7351) and three
    shift + select stages;
  * the XOR swizzle of the staged tile: a bijection for every tile width, and the 32 lanes of every warp read 32
    different banks at every step (lanes that are 32 rows apart would otherwise all hit one bank).
The kernel itself is compared with numpy through its consumers in tests/test_gpu_parity.py::test_transpose_kernels."""
import random


def prmt(x, y, sel):
    b = [(x >> (8 * i)) & 255 for i in range(4)] + [(y >> (8 * i)) & 255 for i in range(4)]
    r = 0
    for i in range(4):
        r |= b[(sel >> (4 * i)) & 7] << (8 * i)
    return r


def transpose32(a):
    a = list(a)
    M = 0xFFFFFFFF
    for k in range(16):
        x, y = a[k], a[k + 16]
        a[k], a[k + 16] = prmt(x, y, 0x5410), prmt(x, y, 0x7632)
    for k in range(32):
        if k & 8:
            continue
        x, y = a[k], a[k + 8]
        a[k], a[k + 8] = prmt(x, y, 0x6240), prmt(x, y, 0x7351)
    for sh, m in ((4, 0x0F0F0F0F), (2, 0x33333333), (1, 0x55555555)):
        for k in range(32):
            if k & sh:
                continue
            x, y = a[k], a[k + sh]
            a[k] = (x & m) | (((y << sh) & M) & ~m & M)
            a[k + sh] = ((x >> sh) & m) | (y & ~m & M)
    return a


def test_register_transpose_network():
    rng = random.Random(7)
    for _ in range(50):
        a = [rng.getrandbits(32) for _ in range(32)]
        t = transpose32(a)
        assert all(((t[b] >> r) & 1) == ((a[r] >> b) & 1) for b in range(32) for r in range(32))
    ident = [1 << r for r in range(32)]
    assert transpose32(ident) == ident


def test_tile_swizzle_is_a_conflict_free_bijection():
    for wc_log in range(1, 6):
        WC, IB = 1 << wc_log, 512 >> wc_log

        def swz(ib):
            return (ib & 31) if IB >= 32 else ((ib << 1) & 31)
        pos = {(R * WC + c) ^ swz(R >> 5) for R in range(IB * 32) for c in range(WC)}
        assert len(pos) == 16384 and max(pos) < 16384  # every word of the 64 KB tile exactly once
        for R in range(IB * 32):  # a 16-byte staging chunk stays one aligned 16-byte chunk (words permuted inside)
            for q in range(max(WC // 4, 1)):
                w = min(WC, 4)
                dst = {((R * WC + w * q + j) ^ swz(R >> 5)) // w for j in range(w)}
                assert len(dst) == 1
        for warp in range(16):
            for r in range(32):
                banks = set()
                for lane in range(32):
                    tid = warp * 32 + lane
                    ib, c = tid % IB, tid // IB
                    banks.add((((32 * ib + r) * WC + c) ^ swz(ib)) % 32)
                assert len(banks) == 32, (wc_log, warp, r)
