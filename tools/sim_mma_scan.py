"""CPU simulation of one warp of scan_mma_kernel: fragment layouts of mma.sync.m16n8k32.u8 (PTX ISA), the A-fragment
construction from packed code words, the record byte order (rec_pos) and the ballot -> bitmap rotation.  Catches index bugs
without a GPU."""
import numpy as np

rng = np.random.default_rng(1)
D = 192
W32 = D // 32
JJ = W32 // 2
NT = 2


def rec_pos(d):
    j, r = d >> 5, d & 31
    return 64 * (j >> 1) + 16 * (r & 3) + 8 * (j & 1) + 4 * ((r >> 2) & 1) + (r >> 3)


bits = rng.integers(0, 2, size=(32, D), dtype=np.uint32)       # 32 vectors of the warp
q = rng.integers(0, 16, size=(8 * NT, D), dtype=np.uint32)      # records
codes = np.zeros((32, W32), np.uint32)
for d in range(D):
    codes[:, d // 32] |= bits[:, d] << np.uint32(d % 32)
rec = np.zeros((8 * NT, D), np.uint8)
for d in range(D):
    rec[:, rec_pos(d)] = (q[:, d] & 15) << (3 - (d & 3))
assert len({rec_pos(d) for d in range(D)}) == D and max(rec_pos(d) for d in range(D)) == D - 1


def mma(acc, A, B):
    """acc[lane][4], A[lane][4] u32, B[lane][2] u32 -> per PTX fragment layout of m16n8k32 (u8)."""
    Am = np.zeros((16, 32), np.int64)
    Bm = np.zeros((32, 8), np.int64)
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for r in range(4):
            row = g + 8 * (r & 1)
            k0 = 4 * t + 16 * (r >> 1)
            for i in range(4):
                Am[row, k0 + i] = (int(A[lane][r]) >> (8 * i)) & 255
        for r in range(2):
            k0 = 4 * t + 16 * r
            for i in range(4):
                Bm[k0 + i, g] = (int(B[lane][r]) >> (8 * i)) & 255
    C = Am @ Bm
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        acc[lane][0] += C[g, 2 * t]
        acc[lane][1] += C[g, 2 * t + 1]
        acc[lane][2] += C[g + 8, 2 * t]
        acc[lane][3] += C[g + 8, 2 * t + 1]


acc = np.zeros((2, NT, 32, 4), np.int64)
for jj in range(JJ):
    af = np.zeros((2, 2, 32, 4), np.uint32)
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        amask = (0x01010101 << t) & 0xFFFFFFFF
        for s in range(4):
            wx, wy = int(codes[4 * g + s, 2 * jj]), int(codes[4 * g + s, 2 * jj + 1])
            af[0][s >> 1][lane][s & 1] = wx & amask
            af[0][s >> 1][lane][(s & 1) + 2] = (wx >> 4) & amask
            af[1][s >> 1][lane][s & 1] = wy & amask
            af[1][s >> 1][lane][(s & 1) + 2] = (wy >> 4) & amask
    for nt in range(NT):
        b = np.zeros((32, 4), np.uint32)
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            raw = rec[nt * 8 + g, 64 * jj + 16 * t: 64 * jj + 16 * t + 16]
            b[lane] = raw.view(np.uint32)
        mma(acc[0][nt], af[0][0], b[:, 0:2])
        mma(acc[1][nt], af[0][1], b[:, 0:2])
        mma(acc[0][nt], af[1][0], b[:, 2:4])
        mma(acc[1][nt], af[1][1], b[:, 2:4])

want = bits.astype(np.int64) @ q.astype(np.int64).T  # [vector, record]
pred = np.zeros((32, 8 * NT), bool)
thr = np.median(want)
for nt in range(NT):
    bal = np.zeros((2, 4), np.uint32)
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for o in range(2):
            col = nt * 8 + 2 * t + o
            for s in range(4):
                iacc = acc[s >> 1][nt][lane][(s & 1) * 2 + o]
                assert iacc == 8 * want[4 * g + s, col], (lane, s, o)
                if want[4 * g + s, col] < thr:
                    bal[o][s] |= np.uint32(1 << lane)
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for o in range(2):
            bm = 0
            for s in range(4):
                x = int(bal[o][s])
                n = (t - s) & 31
                rot = ((x >> n) | (x << (32 - n))) & 0xFFFFFFFF if n else x
                bm |= rot & ((0x11111111 << s) & 0xFFFFFFFF)
            col = nt * 8 + 2 * t + o
            ref = 0
            for v in range(32):
                if want[v, col] < thr:
                    ref |= 1 << v
            assert bm == ref, (lane, o, hex(bm), hex(ref))
print("sim ok: acc == 8*abdp for every fragment element, bitmaps in vector order")
