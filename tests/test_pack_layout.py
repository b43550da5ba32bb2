"""CPU model of the packed-operand 3M kernel's shared-memory layout (mcmc-multispde_b200/csrc/zgemm_tn.cu: M3P,
pack_operand_kernel, zgemm3m_pre_tn_kernel).  No device: the pack kernel's index arithmetic and the consumer lanes' fragment
offsets are restated in numpy and checked for the two things the design claims --

  * addressing: reading the packed stage image at the offsets the lanes use, and combining the fragments the way
    mma.sync.m8n8k4.f64 does (A: row = lane>>2, k = lane&3; B: k = lane&3, col = lane>>2; C: row = lane>>2, cols 2(lane&3)+{0,1}),
    reproduces conj(A)^T B / A^T B for the tile, edge rows/columns and the k tail included (zero fill);
  * bank conflicts: every 128-bit value load of a quarter-warp hits 8 distinct 16-byte slots of a 128-byte line and every
    64-bit sum load of a half-warp hits 16 distinct 8-byte slots (the XOR swizzles c ^ 2(r&3) and c ^ 4(r&3)).

The constants are read from the source so the model cannot drift from the kernel silently."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "mcmc-multispde_b200", "csrc", "zgemm_tn.cu")).read()


def _const(name, block):
    m = re.search(r"\b%s\s*=\s*(\d+)" % name, block)
    assert m, name
    return int(m.group(1))


M3P_BLOCK = SRC[SRC.index("struct M3P {"):SRC.index("static_assert(M3P::SMEM_BYTES")]
BM, BN, BK, STAGES, MF = (_const(k, M3P_BLOCK) for k in ("BM", "BN", "BK", "STAGES", "MF"))
WI = BM // (8 * MF)
WARPS = WI * (BN // 16)


def test_source_still_has_the_modelled_swizzles_and_offsets():
    assert (BM, BN, BK, STAGES, MF) == (64, 64, 16, 4, 4) and WARPS == 8
    for snippet in ("data[r * W + (c ^ (2 * t))] = v;", "sums[r * W + (c ^ (4 * t))] = fma(sign, v.y, v.x);",
                    "const int a_off = (t * BM + wi * 8 * MF + (g ^ (2 * t))) * 2;",
                    "const int b_off = B_DATA + (t * BN + wj * 16 + (g ^ (2 * t))) * 2;",
                    "sa_off[a] = A_SUM + t * BM + wi * 8 * MF + ((a * 8 + g) ^ (4 * t));",
                    "sb_off[b] = B_SUM + t * BN + wj * 16 + ((b * 8 + g) ^ (4 * t));",
                    "S + a_off + kk * 4 * BM * 2 + a * 16", "S[sa_off[a] + kk * 4 * BM]",
                    "S + b_off + kk * 4 * BN * 2 + b * 16", "S[sb_off[b] + kk * 4 * BN]"):
        assert snippet in SRC, snippet
    assert STAGES * (BK * BM * 3 + BK * BN * 3) * 8 + 2 * STAGES * 8 == 196672      # one CTA per SM: > 227 KB / 2


def pack_tile(X, tk, tc, W, sign):
    """pack_operand_kernel<W> for tile (tk, tc): [16 x W complex | 16 x W sums] as a flat array of doubles."""
    rows, cols = X.shape
    data = np.zeros((16 * W,), dtype=np.complex128)
    sums = np.zeros((16 * W,), dtype=np.float64)
    for r in range(16):
        for c in range(W):
            kr, gc = tk * 16 + r, tc * W + c
            v = X[kr, gc] if (kr < rows and gc < cols) else 0.0
            t = r & 3
            data[r * W + (c ^ (2 * t))] = v
            sums[r * W + (c ^ (4 * t))] = np.real(v) + sign * np.imag(v)
    return np.concatenate((data.view(np.float64), sums))


def lane_offsets(warp, lane):
    """Offsets (in doubles, inside a stage) of zgemm3m_pre_tn_kernel's consumer lane, for kk = 0."""
    g, t = lane >> 2, lane & 3
    wi, wj = warp % WI, warp // WI
    A_SUM, B_DATA = BK * BM * 2, BK * BM * 3
    B_SUM = B_DATA + BK * BN * 2
    a_val = [(t * BM + wi * 8 * MF + (g ^ (2 * t))) * 2 + a * 16 for a in range(MF)]
    b_val = [B_DATA + (t * BN + wj * 16 + (g ^ (2 * t))) * 2 + b * 16 for b in range(2)]
    a_sum = [A_SUM + t * BM + wi * 8 * MF + ((a * 8 + g) ^ (4 * t)) for a in range(MF)]
    b_sum = [B_SUM + t * BN + wj * 16 + ((b * 8 + g) ^ (4 * t)) for b in range(2)]
    return a_val, b_val, a_sum, b_sum


@pytest.mark.parametrize("conj", [True, False])
@pytest.mark.parametrize("m,n,k", [(64, 64, 32), (50, 37, 21), (64, 20, 16), (7, 64, 5)])
def test_fragment_addressing_reproduces_the_product(m, n, k, conj):
    rng = np.random.default_rng(11)
    A = rng.standard_normal((k, m)) + 1j * rng.standard_normal((k, m))
    B = rng.standard_normal((k, n)) + 1j * rng.standard_normal((k, n))
    ktiles = -(-k // BK)
    t1, t2, t3 = (np.zeros((BM, BN)) for _ in range(3))
    for kt in range(ktiles):
        S = np.concatenate((pack_tile(A, kt, 0, BM, -1.0 if conj else 1.0), pack_tile(B, kt, 0, BN, 1.0)))
        assert S.size == BK * BM * 3 + BK * BN * 3
        for warp in range(WARPS):
            wi, wj = warp % WI, warp // WI
            off = [lane_offsets(warp, lane) for lane in range(32)]
            for kk in range(BK // 4):
                for a in range(MF):
                    Ar, Ai, As = (np.zeros((8, 4)) for _ in range(3))
                    for lane in range(32):
                        g, t = lane >> 2, lane & 3
                        o = off[lane][0][a] + kk * 4 * BM * 2
                        Ar[g, t], Ai[g, t] = S[o], S[o + 1]
                        As[g, t] = S[off[lane][2][a] + kk * 4 * BM]
                    for b in range(2):
                        Br, Bi, Bs = (np.zeros((4, 8)) for _ in range(3))
                        for lane in range(32):
                            g, t = lane >> 2, lane & 3
                            o = off[lane][1][b] + kk * 4 * BN * 2
                            Br[t, g], Bi[t, g] = S[o], S[o + 1]
                            Bs[t, g] = S[off[lane][3][b] + kk * 4 * BN]
                        rows = slice(wi * 8 * MF + a * 8, wi * 8 * MF + a * 8 + 8)
                        cols = slice(wj * 16 + b * 8, wj * 16 + b * 8 + 8)
                        t1[rows, cols] += Ar @ Br
                        t2[rows, cols] += Ai @ Bi
                        t3[rows, cols] += As @ Bs
    re = t1 + t2 if conj else t1 - t2                       # the kernel's epilogue
    im = (t3 - t1) + t2 if conj else (t3 - t1) - t2
    C = (re + 1j * im)[:m, :n]
    want = (A.conj() if conj else A).T @ B
    assert np.abs(C - want).max() <= 1e-12 * np.abs(want).max()
    assert np.all((re + 1j * im)[m:, :] == 0) and np.all((re + 1j * im)[:, n:] == 0)      # zero fill beyond the matrix


def test_fragment_loads_are_bank_conflict_free():
    """Shared memory has 32 four-byte banks = one 128-byte line.  A 128-bit load is served a quarter-warp at a time, a 64-bit
    load a half-warp at a time; within such a group all addresses must fall into distinct slots of the line."""
    for warp in range(WARPS):
        off = [lane_offsets(warp, lane) for lane in range(32)]
        for kk in range(BK // 4):
            for a in range(MF):
                for q in range(4):                                                   # quarter-warps: lanes 8q .. 8q+7
                    slots = {((off[l][0][a] + kk * 4 * BM * 2) * 8 // 16) % 8 for l in range(8 * q, 8 * q + 8)}
                    assert len(slots) == 8
                for h in range(2):                                                   # half-warps
                    slots = {((off[l][2][a] + kk * 4 * BM) * 8 // 8) % 16 for l in range(16 * h, 16 * h + 16)}
                    assert len(slots) == 16
            for b in range(2):
                for q in range(4):
                    slots = {((off[l][1][b] + kk * 4 * BN * 2) * 8 // 16) % 8 for l in range(8 * q, 8 * q + 8)}
                    assert len(slots) == 8
                for h in range(2):
                    slots = {((off[l][3][b] + kk * 4 * BN) * 8 // 8) % 16 for l in range(16 * h, 16 * h + 16)}
                    assert len(slots) == 16


def test_value_loads_are_16_byte_aligned_and_inside_the_stage():
    stage = BK * BM * 3 + BK * BN * 3
    for warp in range(WARPS):
        for lane in range(32):
            a_val, b_val, a_sum, b_sum = lane_offsets(warp, lane)
            for kk in range(BK // 4):
                for o in [x + kk * 4 * BM * 2 for x in a_val] + [x + kk * 4 * BN * 2 for x in b_val]:
                    assert o % 2 == 0 and 0 <= o < stage - 1
                for o in [x + kk * 4 * BM for x in a_sum] + [x + kk * 4 * BN for x in b_sum]:
                    assert 0 <= o < stage


@pytest.mark.parametrize("MFi", [2, 4])
def test_in_loop_3m_kernel_value_loads_are_conflict_free(MFi):
    """zgemm3m_tn_kernel (struct M3): operand rows are stored with a pitch of BM + 2 / BN + 2 complex elements (== 2 mod 8), so
    the 8 lanes of a quarter-warp (g in {2q', 2q'+1}, t = 0..3) hit 8 distinct 16-byte slots."""
    assert "static constexpr int PA = BM + 2;" in SRC and "static constexpr int PB = BN + 2;" in SRC
    assert "const int a_off = (t * PA + wi * 8 * MF + g) * 2;" in SRC
    assert "const int b_off = (BK * PA + t * PB + wj * 16 + g) * 2;" in SRC
    bm, bn, bk = 64, 32, 16
    pa, pb = bm + 2, bn + 2
    wi_n = bm // (8 * MFi)
    for warp in range(wi_n * (bn // 16)):
        wi, wj = warp % wi_n, warp // wi_n
        for kk in range(bk // 4):
            for q in range(4):
                lanes = range(8 * q, 8 * q + 8)
                for a in range(MFi):
                    slots = {(((l & 3) * pa + wi * 8 * MFi + (l >> 2)) + kk * 4 * pa + a * 8) % 8 for l in lanes}
                    assert len(slots) == 8
                for b in range(2):
                    slots = {((bk * pa + (l & 3) * pb + wj * 16 + (l >> 2)) + kk * 4 * pb + b * 8) % 8 for l in lanes}
                    assert len(slots) == 8
