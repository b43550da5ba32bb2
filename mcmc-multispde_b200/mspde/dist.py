"""Block-distributed Phi(L) for one chain at the largest basis sizes (BASELINE config 5).

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).  The N x N and M x M Hermitian systems of
pCN._log_ratio (/root/reference/mcmc/image_cupy.py:634-643) are held in a 1-D block-cyclic layout over row blocks
(row block i of width `nb` lives on rank i mod P, stored row-major, upper part used):

  r = L^H L + HtH + delta I   each rank forms its own row blocks with local DMMA GEMMs (L is replicated: the fused
                              assembly writes it from the (2n-1)^2 tables in a few ms, no communication)
  U = chol(r)                 right-looking: the owner factorises the diagonal block and solves its row panel locally,
                              broadcasts the nb x (n - p nb) panel (ncclBroadcast), every rank updates its own row
                              blocks below with U12_i^H U12 (zgemm_tn, K = nb)
  Ht = U^-H H^H               the M right-hand sides are split across ranks (U is complete on every rank after the
                              panel broadcasts), then all_gather
  Q_inv = I - Ht^H Ht, V = chol(Q_inv), Phi = 0.5 ||V y||^2 - sum log diag V      same layout, final all_reduce of 2 scalars

Every numerical operation is a libmspde C-ABI call on sub-blocks; torch is used for buffers, copies and NCCL only.
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from . import _ffi

CDTYPE = torch.complex128
c_double = ctypes.c_double


def _round_up(x, a):
    return (x + a - 1) // a * a


class BlockRows:
    """Owned row blocks of an n x n matrix, row-major with a padded leading dimension."""

    def __init__(self, n, nb, world, rank, device):
        self.n, self.nb, self.world, self.rank = n, nb, world, rank
        self.nblk = (n + nb - 1) // nb
        self.owned = [i for i in range(self.nblk) if i % world == rank]
        self.ld = _round_up(n, 8)
        self.row0 = {}
        r = 0
        for i in self.owned:
            self.row0[i] = r
            r += self.rows(i)
        self.data = torch.zeros((max(r, 1), self.ld), dtype=CDTYPE, device=device)

    def rows(self, i):
        return min(self.nb, self.n - i * self.nb)

    def block(self, i):
        """[rows(i), n] view of row block i."""
        r = self.row0[i]
        return self.data[r:r + self.rows(i), :self.n]


def drive(tasks):
    """Round-robin over (generator, torch.cuda.Stream) pairs: each `next` enqueues one panel's worth of kernels and NCCL calls
    on that task's stream, so independent block-distributed factorizations overlap on the GPUs (one fills the other's
    panel-phase bubbles).  Returns the generators' return values."""
    results = [None] * len(tasks)
    live = list(range(len(tasks)))
    while live:
        for i in list(live):
            gen, stream = tasks[i]
            with torch.cuda.stream(stream):
                try:
                    next(gen)
                except StopIteration as e:
                    results[i] = e.value
                    live.remove(i)
    return results


def run_to_completion(gen):
    try:
        while True:
            next(gen)
    except StopIteration as e:
        return e.value


class DistPhi:
    def __init__(self, N, M, nb=512, device=None, group=None):
        self.group = group
        self.lib = _ffi.lib()
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.N, self.M, self.nb = N, M, nb
        self.info = torch.zeros(4, dtype=torch.int32, device=self.device)
        self.timings = {}
        self._events = None

    # ---- thin C-ABI wrappers on strided views ------------------------------------------------------------------
    def _gemm(self, A, B, C, alpha, beta, upper):
        k, m = A.shape
        n = B.shape[1]
        _ffi.check(self.lib.mspde_zgemm_tn(_ffi.current_stream(), 1, m, n, k, c_double(alpha), _ffi.ptr(A), A.stride(0),
                                           _ffi.ptr(B), B.stride(0), c_double(beta), _ffi.ptr(C), C.stride(0),
                                           1 if upper else 0, None, 1), "zgemm_tn")

    def _potrf(self, A, winv):
        _ffi.check(self.lib.mspde_zpotrf_upper(_ffi.current_stream(), _ffi.ptr(A), A.stride(0), A.shape[0], _ffi.ptr(winv),
                                               _ffi.ptr(self.info)), "zpotrf")

    def _trsm(self, U, winv, B):
        _ffi.check(self.lib.mspde_ztrsm_uh(_ffi.current_stream(), _ffi.ptr(U), U.stride(0), _ffi.ptr(winv), U.shape[0],
                                           _ffi.ptr(B), B.stride(0), B.shape[1]), "ztrsm")

    # ---- building blocks ---------------------------------------------------------------------------------------
    def gram_rows(self, A, out: BlockRows, alpha, beta):
        """out[i-rows, i*nb:] = alpha * A[:, i-rows]^H A[:, i*nb:] + beta * out (upper part of every owned row block)."""
        nb = self.nb
        for i in out.owned:
            c0 = i * nb
            blk = out.block(i)
            self._gemm(A[:, c0:c0 + out.rows(i)], A[:, c0:], blk[:, c0:], alpha, beta, True)

    def cholesky(self, mat: BlockRows, keep_full=None, winv_all=None, y=None):
        return run_to_completion(self.cholesky_gen(mat, keep_full, winv_all, y))

    def cholesky_gen(self, mat: BlockRows, keep_full=None, winv_all=None, y=None):
        """In-place distributed upper Cholesky (generator: yields after every panel), look-ahead of one panel: the owner of
        row block p+1 applies panel p to that block first, factorises it and starts its broadcast while every rank is still
        busy with the rest of panel p's trailing update, so a panel factorisation stalls one rank instead of all of them.
        keep_full: optional n x n buffer receiving the complete factor (panels are seen by every rank anyway).  winv_all:
        inverses of the 64-blocks (needed with keep_full for the later solve).  y: if given, returns the local partial sums
        (sum |V y|^2 rows, sum log diag) of the owned panels."""
        n, nb, P = mat.n, mat.nb, self.world
        nblk = mat.nblk
        ldp = _round_up(n, 8)
        panels = [torch.empty(nb * ldp, dtype=CDTYPE, device=self.device) for _ in range(2)]
        wblk = (nb // 64) * 64 * 64
        wbufs = [torch.empty(wblk, dtype=CDTYPE, device=self.device) for _ in range(2)]
        acc = torch.zeros(2, dtype=torch.float64, device=self.device)
        rowsq = torch.empty(nb, dtype=torch.float64, device=self.device)
        rowlog = torch.empty(nb, dtype=torch.float64, device=self.device)

        def view(p):
            bp, w = mat.rows(p), n - p * nb
            return panels[p % 2][: bp * w].view(bp, w)

        def factor(p):          # owner only: diagonal block + row panel, copy into the broadcast buffer
            bp, c0 = mat.rows(p), p * nb
            w = n - c0
            blk = mat.block(p)
            self._potrf(blk[:, c0:c0 + bp], wbufs[p % 2])
            if w > bp:
                self._trsm(blk[:, c0:c0 + bp], wbufs[p % 2], blk[:, c0 + bp:])
            pv = view(p)
            pv.copy_(blk[:, c0:])
            if y is not None:
                _ffi.check(self.lib.mspde_upper_rows_times_y(_ffi.current_stream(), _ffi.ptr(pv), pv.stride(0), bp, w,
                                                             _ffi.ptr(y[c0:]), _ffi.ptr(rowsq), _ffi.ptr(rowlog)), "vy")
                acc[0] += rowsq[:bp].sum()
                acc[1] += rowlog[:bp].sum()

        def publish(p):         # everyone: receive panel p, keep the pieces of the factor that later steps need
            owner = p % P
            pv = view(p)
            if P > 1:
                dist.broadcast(torch.view_as_real(pv), src=owner, group=self.group)
                if winv_all is not None:
                    dist.broadcast(torch.view_as_real(wbufs[p % 2]), src=owner, group=self.group)
            c0 = p * nb
            if keep_full is not None:
                keep_full[c0:c0 + mat.rows(p), c0:] = pv
            if winv_all is not None:
                winv_all[p * wblk:(p + 1) * wblk] = wbufs[p % 2]

        def update(p, blocks):  # trailing update of the given owned row blocks with panel p
            pv = view(p)
            for i in blocks:
                o = (i - p) * nb
                bi = mat.rows(i)
                self._gemm(pv[:, o:o + bi], pv[:, o:], mat.block(i)[:, i * nb:], -1.0, 1.0, True)

        if self.rank == 0:
            factor(0)
        publish(0)
        for p in range(nblk):
            mine = [i for i in mat.owned if i > p]
            if p + 1 < nblk:
                if self.rank == (p + 1) % P:
                    update(p, [p + 1])
                    factor(p + 1)
                    publish(p + 1)
                    update(p, [i for i in mine if i != p + 1])
                else:
                    update(p, mine)
                    publish(p + 1)
            yield
        return acc

    # ---- Phi ---------------------------------------------------------------------------------------------------
    def setup(self, H, y, delta, HtH_rows: BlockRows | None = None):
        """H: full M x N (replicated), y: M.  HtH row blocks are formed with the same distributed Gram kernel unless given."""
        self.H, self.y, self.delta = H, y, float(delta)
        N, M, nb = self.N, self.M, self.nb
        if HtH_rows is None:
            HtH_rows = BlockRows(N, nb, self.world, self.rank, self.device)
            self.gram_rows(H, HtH_rows, 1.0, 0.0)
            for i in HtH_rows.owned:               # Hermitian by construction; clear the rounding-level imaginary diagonal
                d = HtH_rows.block(i)[:, i * nb:].diagonal()
                d.imag.zero_()
        self.HtH_rows = HtH_rows
        self.r = BlockRows(N, nb, self.world, self.rank, self.device)
        self.Q = BlockRows(M, nb, self.world, self.rank, self.device)
        self.U = torch.zeros((N, _round_up(N, 8)), dtype=CDTYPE, device=self.device)[:, :N]
        self.winv = torch.zeros(self.r.nblk * (nb // 64) * 64 * 64, dtype=CDTYPE, device=self.device)
        self.mc = _round_up((M + self.world - 1) // self.world, 8)          # right-hand sides per rank
        self.Ht_chunk = torch.zeros((N, self.mc), dtype=CDTYPE, device=self.device)
        self.Ht_all = torch.zeros((self.world, N, self.mc), dtype=CDTYPE, device=self.device)
        self.Ht = torch.zeros((N, _round_up(self.mc * self.world, 8)), dtype=CDTYPE, device=self.device)
        lo = self.rank * self.mc
        hi = min(M, lo + self.mc)
        self.Ht_chunk.zero_()
        if hi > lo:
            self._Hct_chunk = H[lo:hi, :].conj().transpose(0, 1).contiguous()      # H^H columns of this rank (setup copy)
        else:
            self._Hct_chunk = None
        self._lohi = (lo, hi)

    def phi(self, L):
        """Phi(L) for the replicated N x N matrix L; collective call, returns a Python float on every rank."""
        ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
        ev0.record()
        acc = run_to_completion(self.phi_gen(L, timed=True))
        ev1.record()
        return self.finalize(acc)

    def finalize(self, acc):
        """Synchronise and turn the device-side sums of phi_gen into the Python float (raises LinAlgError on a failed Cholesky)."""
        torch.cuda.synchronize()
        if self._events:
            names = ["gram_N", "chol_N", "trsm", "gram_M", "chol_M"]
            self.timings = {nm: self._events[k].elapsed_time(self._events[k + 1]) for k, nm in enumerate(names)}
        info = int(self.info[0].item())
        if info:
            self.info.zero_()
        _ffi.check(info, "distributed Phi (Cholesky)")
        a = acc.cpu()
        return 0.5 * float(a[0]) - float(a[1])

    def phi_gen(self, L, timed=False):
        """Generator form of phi(): enqueues the work panel by panel on the current stream and returns the device tensor
        {sum |V y|^2 rows, sum log diag V} (all-reduced); finalize() turns it into Phi."""
        N, M, nb = self.N, self.M, self.nb
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)] if timed else None
        self._events = ev
        mark = (lambda k: ev[k].record()) if timed else (lambda k: None)
        mark(0)
        # r = HtH + delta I + L^H L on the owned row blocks
        self.r.data.copy_(self.HtH_rows.data)
        for i in self.r.owned:
            blk = self.r.block(i)[:, i * nb:]
            _ffi.check(self.lib.mspde_add_diag(_ffi.current_stream(), _ffi.ptr(blk), blk.stride(0), blk.shape[0],
                                               c_double(self.delta)), "add_diag")
        self.gram_rows(L, self.r, 1.0, 1.0)
        mark(1)
        yield
        yield from self.cholesky_gen(self.r, keep_full=self.U, winv_all=self.winv)
        mark(2)
        # Ht = U^-H H^H for this rank's right-hand sides, then gather
        lo, hi = self._lohi
        if hi > lo:
            self.Ht_chunk[:, :hi - lo].copy_(self._Hct_chunk)
            self._trsm(self.U, self.winv, self.Ht_chunk)
        yield
        if self.world > 1:
            dist.all_gather_into_tensor(torch.view_as_real(self.Ht_all), torch.view_as_real(self.Ht_chunk), group=self.group)
            for q in range(self.world):
                self.Ht[:, q * self.mc:(q + 1) * self.mc] = self.Ht_all[q]
        else:
            self.Ht[:, :self.mc] = self.Ht_chunk
        mark(3)
        yield
        # Q_inv = I - Ht^H Ht on the owned row blocks
        Ht = self.Ht[:, :M]
        self.gram_rows(Ht, self.Q, -1.0, 0.0)
        for i in self.Q.owned:
            blk = self.Q.block(i)[:, i * nb:]
            _ffi.check(self.lib.mspde_add_diag(_ffi.current_stream(), _ffi.ptr(blk), blk.stride(0), blk.shape[0], c_double(1.0)),
                       "add_diag")
        mark(4)
        yield
        acc = yield from self.cholesky_gen(self.Q, y=self.y)
        if self.world > 1:
            dist.all_reduce(acc, group=self.group)
            dist.all_reduce(self.info, op=dist.ReduceOp.MAX, group=self.group)
        mark(5)
        return acc


# ======================================================================================================================
# Block-distributed LU solve and QR least squares (the remaining dense pieces of one n=96 iteration)
# ======================================================================================================================
class BlockCols:
    """Owned 64-column blocks of a rows x ntot matrix in a 1-D block-cyclic layout (block b on rank b mod P), stored as
    one local row-major matrix whose columns are the owned blocks in ascending order."""

    NB = 64

    def __init__(self, rows, ntot, world, rank, device, nb=None):
        if nb is not None:
            self.NB = int(nb)            # instance override: the rank-128 QR distributes 128-column blocks
        nb = self.NB
        self.rows, self.ntot, self.world, self.rank = rows, ntot, world, rank
        self.nblk = (ntot + nb - 1) // nb
        self.owned = [b for b in range(self.nblk) if b % world == rank]
        self.col0 = {}
        c = 0
        for b in self.owned:
            self.col0[b] = c
            c += self.width(b)
        self.ncols = c
        self.ld = _round_up(max(c, 1), 8)
        self.data = torch.zeros((rows, self.ld), dtype=CDTYPE, device=device)

    def width(self, b):
        return min(self.NB, self.ntot - b * self.NB)

    def right_of(self, p, pw):
        """(local column offset, number of local columns) strictly right of panel p's pw pivot columns."""
        if p in self.col0:
            start = self.col0[p] + pw
        else:
            nxt = [b for b in self.owned if b > p]
            start = self.col0[nxt[0]] if nxt else self.ncols
        return start, self.ncols - start

    def ptr(self, row, col):
        return ctypes.c_void_p(self.data.data_ptr() + (row * self.ld + col) * 16)

    def fill_from(self, full, aug=None):
        """Columns < full.shape[1] come from the replicated matrix `full`; the remaining (augmented) columns from `aug`."""
        nb, nreal = self.NB, full.shape[1]
        for b in self.owned:
            g0, w, l0 = b * nb, self.width(b), self.col0[b]
            wr = max(0, min(w, nreal - g0))
            if wr > 0:
                self.data[:, l0:l0 + wr] = full[:, g0:g0 + wr]
            if wr < w:
                self.data[:, l0 + wr:l0 + w] = aug[:, g0 + wr - nreal:g0 + w - nreal]


class _DistFactor:
    def __init__(self, device=None, group=None):
        self.group = group
        self.lib = _ffi.lib()
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        for name in ("mspde_qr_work_bytes", "mspde_lu_work_bytes", "mspde_qr_pair_work_bytes"):
            getattr(self.lib, name).restype = ctypes.c_longlong

    def _bcast_bytes(self, work, off, size, src):
        if self.world > 1:
            dist.broadcast(work[off:off + size], src=src, group=self.group)

    def _gather_upper(self, mat: BlockCols, n, ntot):
        return run_to_completion(self._gather_upper_gen(mat, n, ntot))

    def _gather_upper_gen(self, mat: BlockCols, n, ntot):
        """Every rank receives the upper-trapezoidal factor (rows < n) of the distributed matrix as one n x ntot matrix."""
        nb = mat.NB
        full = torch.zeros((n, _round_up(ntot, 8)), dtype=CDTYPE, device=self.device)[:, :ntot]
        for b in range(mat.nblk):
            w = mat.width(b)
            nr = min(n, (b + 1) * nb)
            buf = torch.empty((nr, w), dtype=CDTYPE, device=self.device)
            if b % self.world == self.rank:
                buf.copy_(mat.data[:nr, mat.col0[b]:mat.col0[b] + w])
            if self.world > 1:
                dist.broadcast(torch.view_as_real(buf), src=b % self.world, group=self.group)
            full[:nr, b * nb:b * nb + w] = buf
            if b % 8 == 7:
                yield
        return full

    def _back_substitute(self, full, n, dcol):
        x = torch.empty(n, dtype=CDTYPE, device=self.device)
        _ffi.check(self.lib.mspde_trsv_upper(_ffi.current_stream(), _ffi.ptr(full), full.stride(0), n, dcol, _ffi.ptr(x)), "trsv")
        return x


class DistLU(_DistFactor):
    """x = L^-1 rhs and log|det L| with the pivoted LU distributed over 64-column blocks (cp.linalg.solve, image_cupy.py:561)."""

    def solve(self, L, rhs):
        x, logabs, info = run_to_completion(self.solve_gen(L, rhs))
        _ffi.check(int(info[0].item()), "distributed LU (singular matrix)")
        return x, float(logabs.item())

    def solve_gen(self, L, rhs):
        """Generator: yields after every panel; returns (x, logabs tensor, info tensor) without synchronising."""
        n = L.shape[0]
        ntot = n + 1
        mat = BlockCols(n, ntot, self.world, self.rank, self.device)
        mat.fill_from(L, rhs.reshape(n, 1))
        nbytes = int(self.lib.mspde_lu_work_bytes(n, ntot))
        works = [torch.empty(nbytes, dtype=torch.uint8, device=self.device) for _ in range(2)]
        lay = (ctypes.c_longlong * 6)()
        self.lib.mspde_lu_work_layout(n, ntot, _ffi.ptr(works[0]), lay)
        logabs = torch.zeros(1, dtype=torch.float64, device=self.device)
        info = torch.zeros(4, dtype=torch.int32, device=self.device)
        st = _ffi.current_stream
        nb = mat.NB
        npan = (n + nb - 1) // nb
        dims = lambda p: (p * nb, min(nb, n - p * nb), n - p * nb)       # c0, pw, mp

        def factor(p):
            c0, pw, mp = dims(p)
            _ffi.check(self.lib.mspde_lu_panel(st(), mat.ptr(c0, mat.col0[p]), mat.ld, mp, pw, c0, n, ntot, _ffi.ptr(works[p % 2]),
                                               _ffi.ptr(info), _ffi.ptr(logabs)), "lu_panel")

        def publish(p):
            for k in (0, 2, 4):
                self._bcast_bytes(works[p % 2], lay[k], lay[k + 1], p % self.world)

        def apply(p, start, n2):
            c0, pw, mp = dims(p)
            if n2 > 0:
                _ffi.check(self.lib.mspde_lu_apply(st(), mp, pw, n, ntot, _ffi.ptr(works[p % 2]), mat.ptr(c0, start), mat.ld, n2),
                           "lu_apply")

        # look-ahead of one panel: the owner of panel p+1 brings that block up to date and factors it before it joins the rest
        # of panel p's trailing update; the other ranks update first and only then wait for the broadcast
        if self.rank == 0:
            factor(0)
        publish(0)
        for p in range(npan):
            c0, pw, mp = dims(p)
            start, n2 = mat.right_of(p, pw)
            if p + 1 < npan:
                if self.rank == (p + 1) % self.world:
                    wn = mat.width(p + 1)
                    s1 = mat.col0[p + 1]
                    if p in mat.col0:          # single-rank group: panel p's own block may carry extra columns before block p+1
                        apply(p, start, s1 - start)
                    apply(p, s1, wn)
                    factor(p + 1)
                    publish(p + 1)
                    apply(p, s1 + wn, mat.ncols - (s1 + wn))
                else:
                    apply(p, start, n2)
                    publish(p + 1)
            else:
                apply(p, start, n2)
            yield
        if self.world > 1:
            dist.all_reduce(logabs, group=self.group)
            dist.all_reduce(info, op=dist.ReduceOp.MAX, group=self.group)
        full = yield from self._gather_upper_gen(mat, n, ntot)
        x = self._back_substitute(full, n, n)
        return x, logabs, info


class DistQR(_DistFactor):
    """v = argmin || [H; L] v - rhs || with the Householder QR distributed over 64-column blocks (util.lstsq, util_cupy.py:590-609)."""

    def lstsq(self, blocks, rhs):
        """blocks: list of replicated row blocks (e.g. [H, L]) stacked vertically; rhs: vector of the total row count, or a
        list of such vectors (they ride along as extra columns; one solution per right-hand side is returned)."""
        return run_to_completion(self.lstsq_gen(blocks, rhs))

    def lstsq_gen(self, blocks, rhs):
        """Rank-128 block reflectors: the matrix is distributed over 128-column blocks; the owner of a block factors both of
        its 64-column panels (mspde_qr_pair_prepare), broadcasts V (live rows x 128), V^T, T1, T2 and V1^H V2, and every rank
        applies the pair to its own columns with one rank-128 update (mspde_qr_pair_apply).  Look-ahead of one block."""
        rows = sum(b.shape[0] for b in blocks)
        n = blocks[0].shape[1]
        many = isinstance(rhs, (list, tuple))
        rhs_list = list(rhs) if many else [rhs]
        naug = len(rhs_list)
        ntot = n + naug
        nbk = 128
        mat = BlockCols(rows, ntot, self.world, self.rank, self.device, nb=nbk)
        r0 = 0
        for blk in blocks:                     # fill the owned columns block row by block row (no stacked copy is formed)
            for b in mat.owned:
                g0, w, l0 = b * nbk, mat.width(b), mat.col0[b]
                wr = max(0, min(w, n - g0))
                if wr > 0:
                    mat.data[r0:r0 + blk.shape[0], l0:l0 + wr] = blk[:, g0:g0 + wr]
            r0 += blk.shape[0]
        for j, rj in enumerate(rhs_list):       # right-hand side j is global column n + j
            b = (n + j) // nbk
            if b in mat.col0:
                mat.data[:, mat.col0[b] + (n + j - b * nbk)] = rj
        nbytes = int(self.lib.mspde_qr_pair_work_bytes(rows, n, naug))
        works = [torch.empty(nbytes, dtype=torch.uint8, device=self.device) for _ in range(2)]
        lay = (ctypes.c_longlong * 10)()
        self.lib.mspde_qr_pair_work_layout(rows, n, naug, _ffi.ptr(works[0]), lay)
        st = _ffi.current_stream
        npan = (n + nbk - 1) // nbk
        dims = lambda p: (p * nbk, min(nbk, n - p * nbk), rows - p * nbk)     # c0, span, mp
        kks = {}

        def factor(p):
            c0, span, mp = dims(p)
            kk = ctypes.c_int(0)
            _ffi.check(self.lib.mspde_qr_pair_prepare(st(), _ffi.ptr(works[p % 2]), rows, n, naug, mat.ptr(c0, mat.col0[p]), mat.ld, mp,
                                                      span, None, ctypes.byref(kk)), "qr_pair_prepare")

        def publish(p):
            c0, span, mp = dims(p)
            kks[p] = 128 if (span > 64 and mp > 64) else 64            # what mspde_qr_pair_prepare decided, known to every rank
            w, owner = works[p % 2], p % self.world
            self._bcast_bytes(w, lay[0], mp * nbk * 16, owner)             # V: only the mp live rows
            for k in (2, 4, 6, 8):                                         # V^T, T1, T2, V1^H V2
                self._bcast_bytes(w, lay[k], lay[k + 1], owner)

        def apply(p, start, n2):
            c0, span, mp = dims(p)
            if n2 > 0:
                _ffi.check(self.lib.mspde_qr_pair_apply(st(), _ffi.ptr(works[p % 2]), rows, n, naug, kks[p], mp, mat.ptr(c0, start),
                                                        mat.ld, n2), "qr_pair_apply")

        # look-ahead of one block (see DistLU.solve_gen)
        if self.rank == 0:
            factor(0)
        publish(0)
        for p in range(npan):
            c0, span, mp = dims(p)
            start, n2 = mat.right_of(p, span)
            if p + 1 < npan:
                if self.rank == (p + 1) % self.world:
                    wn = mat.width(p + 1)
                    s1 = mat.col0[p + 1]
                    if p in mat.col0:
                        apply(p, start, s1 - start)
                    apply(p, s1, wn)
                    factor(p + 1)
                    publish(p + 1)
                    apply(p, s1 + wn, mat.ncols - (s1 + wn))
                else:
                    apply(p, start, n2)
                    publish(p + 1)
            else:
                apply(p, start, n2)
            yield
        full = yield from self._gather_upper_gen(mat, n, ntot)
        xs = [self._back_substitute(full, n, n + j) for j in range(len(rhs_list))]
        return xs if many else xs[0]


class DistPCN:
    """``pCN.one_step`` (image_cupy.py:553-617) for ONE chain whose dense work is block-distributed over the ranks of the
    default process group (BASELINE config 5).  A sampler, not a timing harness: it carries the chain state of
    SURVEY Appendix A.5, takes injected noise, decides accept/reject and promotes current <- new like the single-GPU
    ``mspde_pcn_step`` -- and is checked against it (tests/test_gpu_dist.py, bench.py's config5 self-check).

    What is replicated: the vectors (noise / samples, N complex each), H, y, and the last layer's L_current / L_latest (the
    fused assembly writes an N x N matrix from the (2n-1)^2 tables in milliseconds, so nothing is communicated for it).
    What is distributed: every factorization -- DistLU (middle-layer solves), DistPhi (both Phi evaluations), DistQR (Gibbs
    least squares, carrying the right-hand sides of BOTH accept outcomes as two extra columns, A.4/A.5).  Independent tasks
    are paired on two streams / two NCCL communicators: (first middle-layer solve || Phi(L_current)), then
    (Phi(L_latest) || Gibbs QR).  The accept test runs on the host from the all-reduced Phi values, identical on every rank."""

    def __init__(self, n, n_ext, n_layers, H, y, D_diag, stdev_sym, sqrt_betas, device, delta=None, chol_epsilon=1e-6, beta=1.0,
                 nb=512, packed_gemm=False):
        from .core import MspdeContext
        self.device = device
        if not packed_gemm:
            # PROCESS-WIDE: large 3M products go through the in-loop kernel.  The packed-operand kernel keeps a workspace per
            # stream (up to 6 GB each, three streams here) that the replicated N x N matrices of n=96 leave no room for, and the
            # per-rank panel updates are short-k products it does not speed up; 5.61 s/iteration was measured this way.
            _ffi.set_gemm_pre_min(0)
        il = 2 * n - 1
        self.n, self.n_ext, self.K = n, n_ext, n_layers
        self.N, self.nr = il * il, 2 * n * n - 2 * n + 1
        self.M = H.shape[0]
        self.H, self.y = H, y
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.ctx = MspdeContext(n, n_ext, self.M, n_layers, device=device.index, assembly_only=True)
        self._D = torch.as_tensor(D_diag).to(device)
        _ffi.check(self.ctx.lib.mspde_ctx_set_inputs(self.ctx._h, None, 0, None, 0, None, 0, None, _ffi.ptr(self._D), None,
                                                     c_double(0.0)))
        self.stdev_sym = torch.as_tensor(stdev_sym, dtype=torch.float64).to(device)
        self.sqrt_betas = [float(x) for x in sqrt_betas]
        self.set_step(beta)
        if self.world > 1:                                # two communicators: two factorizations in flight at once
            self.gA, self.gB = dist.new_group(list(range(self.world))), dist.new_group(list(range(self.world)))
        else:
            self.gA = self.gB = None
        self.sA, self.sB = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)
        self.phi = DistPhi(self.N, self.M, nb=nb, device=device, group=self.gB)
        self.phi.setup(H, y, 0.0)
        if delta is None:                                 # chol_epsilon * ||HtH||_F from the distributed row blocks (473-478)
            nrm2 = torch.zeros(1, dtype=torch.float64, device=device)
            for i in self.phi.HtH_rows.owned:
                blk = self.phi.HtH_rows.block(i)[:, i * nb:]
                nrm2 += 2 * (torch.triu(blk, 1).abs() ** 2).sum() + (blk.diagonal().abs() ** 2).sum()
            if self.world > 1:
                dist.all_reduce(nrm2)
            delta = chol_epsilon * float(nrm2.sqrt())
        self.delta = self.phi.delta = float(delta)
        self.lu, self.qr = DistLU(device, group=self.gA), DistQR(device, group=self.gA)
        z = lambda: torch.zeros(self.N, dtype=CDTYPE, device=device)
        K = n_layers
        self.noise_cur, self.noise_new = [z() for _ in range(K)], [z() for _ in range(K)]
        self.u_cur, self.u_new = [z() for _ in range(K)], [z() for _ in range(K)]
        self.L_cur = None                                 # last layer's current_L (replicated)
        self.accepted_total, self.iteration = 0, 0
        self.last = None

    def set_step(self, beta):
        self.beta = float(beta)
        self.betaZ = float(max(0.0, 1.0 - self.beta ** 2) ** 0.5)

    def _sym(self, w_half):
        return torch.cat((w_half[1:].flip(0).conj(), w_half))          # util.symmetrize (util_cupy.py:201-204)

    def init_state(self, noise_cur, u_cur, u_new=None, noise_new=None):
        """Chain state as lists of N-vectors (numpy or torch): Simulation.__init__'s result (image_cupy.py:902-946).  The last
        layer's current_L is assembled from u_cur[K-2] (753, 577-579)."""
        to = lambda a: torch.as_tensor(a).to(self.device).to(CDTYPE)
        for k in range(self.K):
            self.noise_cur[k].copy_(to(noise_cur[k]))
            self.noise_new[k].copy_(to(noise_new[k] if noise_new is not None else noise_cur[k]))
            self.u_cur[k].copy_(to(u_cur[k]))
            self.u_new[k].copy_(to(u_new[k] if u_new is not None else u_cur[k]))
        self.L_cur = self.ctx.construct_L(self.u_cur[self.K - 2], self.sqrt_betas[self.K - 1])

    def init_from_prior(self, draw_w_sym):
        """Layer.__init__ / Simulation.__init__ (717-766, 902-946): layer 0 u = stdev_sym * w; layer k >= 1: L = assemble(u_{k-1}),
        u = solve(L, w') with the block-distributed LU.  draw_w_sym() returns one symmetric noise vector (identical on all ranks)."""
        K = self.K
        for k in range(K):
            if k == 0:
                u = self.stdev_sym * draw_w_sym()
                self.noise_cur[0].copy_(draw_w_sym())
            else:
                self.noise_cur[k].copy_(draw_w_sym())
                Lk = self.ctx.construct_L(self.u_cur[k - 1], self.sqrt_betas[k])
                u, _ = self.lu.solve(Lk, draw_w_sym())
                if k == K - 1:
                    self.L_cur = Lk
                del Lk
            self.u_cur[k].copy_(u); self.u_new[k].copy_(u)
            self.noise_new[k].copy_(self.noise_cur[k])

    def one_step(self, w_half, log_u, w_last, e):
        """One iteration with injected noise (device tensors: K-1 half vectors, last layer's half vector, M normals).
        Returns dict(accepted, log_ratio, phi_cur, phi_new); raises numpy.linalg.LinAlgError like the single-GPU path."""
        import time
        K, M, N = self.K, self.M, self.N
        main = torch.cuda.current_stream()
        beta, betaZ = self.beta, self.betaZ
        acc_cur = None
        t0 = time.perf_counter()
        for k in range(K - 1):                                                                   # 557-564
            self.noise_new[k].copy_(betaZ * self.noise_cur[k] + beta * self._sym(w_half[k]))
            if k == 0:
                self.u_new[0].copy_(self.stdev_sym * self.noise_new[0])
            else:
                Lk = self.ctx.construct_L(self.u_new[k - 1], self.sqrt_betas[k])                 # 560
                if acc_cur is None:        # pair the first middle-layer solve with Phi(L_current) (567-568)
                    self.sA.wait_stream(main); self.sB.wait_stream(main)
                    (x, _, info), acc_cur = drive([(self.lu.solve_gen(Lk, self.noise_new[k]), self.sA),
                                                   (self.phi.phi_gen(self.L_cur), self.sB)])
                    main.wait_stream(self.sA); main.wait_stream(self.sB)
                    phi_cur = self.phi.finalize(acc_cur)
                    _ffi.check(int(info[0].item()), "distributed LU (singular matrix)")
                else:
                    x, _ = self.lu.solve(Lk, self.noise_new[k])                                  # 561
                self.u_new[k].copy_(x)
                del Lk
        if acc_cur is None:                # two layers: no middle-layer solve to pair with
            phi_cur = self.phi.phi(self.L_cur)
        t1 = time.perf_counter()                 # finalize() synchronised: phase 1 = middle-layer solves || Phi(L_current)
        L_latest = self.ctx.construct_L(self.u_new[K - 2], self.sqrt_betas[K - 1])               # 570
        # Gibbs right-hand sides of both accept outcomes (583-586): noise_cur[K-1] is read AFTER the accept copy
        sw = self._sym(w_last)
        cand_acc = betaZ * self.noise_new[K - 1] + beta * sw
        cand_rej = betaZ * self.noise_cur[K - 1] + beta * sw
        top = (self.y - e).to(CDTYPE)
        rhs = [torch.cat((top, -cand_acc)), torch.cat((top, -cand_rej))]
        self.sA.wait_stream(main); self.sB.wait_stream(main)
        (v_acc, v_rej), acc_new = drive([(self.qr.lstsq_gen([self.H, L_latest], rhs), self.sA),
                                         (self.phi.phi_gen(L_latest), self.sB)])                # 571, 596-599
        main.wait_stream(self.sA); main.wait_stream(self.sB)
        phi_new = self.phi.finalize(acc_new)
        log_ratio = phi_cur - phi_new
        accepted = bool(log_ratio > log_u)                                                       # 573 (strict)
        if accepted:                                                                             # 575-579
            for k in range(K):
                self.u_cur[k].copy_(self.u_new[k])
                self.noise_cur[k].copy_(self.noise_new[k])
            self.L_cur = L_latest
        self.noise_new[K - 1].copy_(cand_acc if accepted else cand_rej)                          # 583
        self.u_new[K - 1].copy_(v_acc if accepted else v_rej)                                    # 598
        self.accepted_total += int(accepted)
        self.iteration += 1
        self.last = dict(accepted=accepted, log_ratio=log_ratio, phi_cur=phi_cur, phi_new=phi_new, log_u=float(log_u),
                         phase_s={"solves||phi_current": round(t1 - t0, 4), "gibbs_qr||phi_latest": round(time.perf_counter() - t1, 4)})
        return self.last

    def current_halves(self):
        """Layer.record_sample payload (792-795)."""
        return torch.stack([u[self.nr - 1:] for u in self.u_cur])


def build_dist_chain(n, n_ext, n_theta, n_layers, device, seed=7, beta=0.3, meas_std=0.2, chol_epsilon=1e-6, kappa=5e9, sigma_0=4e7,
                     sigma_v=3.2e4, sigma_scaling=0.4):
    """``Simulation(..., distributed=True)``'s construction (image_cupy.py:839-946) for the block-distributed chain: tables and
    prior exactly as the single-GPU mirror builds them, H generated on the device (f1), synthetic sinogram y = Re(H v_true) +
    noise, initial state from the prior.  Every rank builds identical inputs (same seeds); returns (chain, noise_fn) where
    noise_fn() draws one iteration's injected noise in the reference's RNG order, identical on every rank."""
    import math
    import numpy as np
    from . import image as im, util
    f = im.FourierAnalysis_2D(n, n_ext)
    m = 2 * n_ext - 1
    N, nr = f.basis_number_2D_sym, f.basis_number_2D_ravel
    sino = im.Sinogram("shepp.png", target_size=m, n_theta=n_theta, stdev=meas_std, relative_location="phantom_images",
                       rng=np.random.Generator(np.random.PCG64(seed)))
    H0 = sino.get_measurement_matrix_device(f.ix.ravel("F"), f.iy.ravel("F"), device)
    sino.synthesize(H0, f)
    H = H0 / meas_std
    del H0
    y = torch.as_tensor(sino.y, dtype=torch.float64).to(device)
    pi = np.float64(util.PI)
    beta_u = (sigma_0 ** 2) * (2 ** 2 * pi * math.gamma(2.0)) / math.gamma(1.0)
    sb0 = np.float32(np.sqrt(beta_u))
    sbv = np.float32(np.sqrt(beta_u * (sigma_v / sigma_0) ** 2))
    Lu = ((f.D_diag / kappa - kappa * np.ones(N, np.float32)) / np.float64(sb0)).astype(np.float32)
    sd = (-1.0 / Lu.astype(np.float64)).astype(np.float32).astype(np.float64)
    sd[nr - 1] /= 2
    sbs = [float(sb0)] + [float(np.float64(sbv) * np.sqrt(sigma_scaling))] * (n_layers - 2) + [float(sbv)]
    chain = DistPCN(n, n_ext, n_layers, H, y, f.D_diag, sd, sbs, device, chol_epsilon=chol_epsilon, beta=beta)
    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    M = H.shape[0]

    def w_half():
        return torch.as_tensor(util.w_half_from_normals(rng.standard_normal(nr), rng.standard_normal(nr))).to(device)

    def noise_fn():
        w = [w_half() for _ in range(n_layers - 1)]
        log_u = float(np.log(rng.random()))
        return w, log_u, w_half(), torch.as_tensor(rng.standard_normal(M)).to(device)
    chain.init_from_prior(lambda: chain._sym(w_half()))
    return chain, noise_fn


class _HistoryLayer:
    """What Simulation.run / the result file need from a Layer (image_cupy.py:717-807) when the state lives in a DistPCN."""

    def __init__(self, order_number, sqrt_beta, history_length, nr):
        import numpy as np
        self.order_number = order_number
        self.is_stationary = order_number == 0
        self.sqrt_beta = float(sqrt_beta)
        self.samples_history = np.empty((history_length, nr), dtype=np.complex128)
        self.sqrt_beta_history = np.empty(history_length, dtype=np.float64)
        self.sqrt_beta_history[0] = self.sqrt_beta
        self.i_record = 0

    def record_sample(self, half_host):                                                          # 792-795
        if self.i_record < self.samples_history.shape[0]:
            self.samples_history[self.i_record, :] = half_host
            self.i_record += 1

    def record_sqrt_beta(self):                                                                  # 797-799
        if self.i_record < self.sqrt_beta_history.shape[0]:
            self.sqrt_beta_history[self.i_record] = self.sqrt_beta


class _DistStepper:
    """The part of the ``pCN`` surface that ``Simulation.run`` drives (one_step, adapt_step, adapt_step_sqrtBeta, record_skip),
    on top of a DistPCN.  ``beta`` / ``betaZ`` ARE the chain's step sizes, so the reference's adaptation rules -- borrowed
    unchanged from mspde.image.pCN -- act on the distributed sampler directly."""

    def __init__(self, chain, noise_fn, target_acceptance_rate=0.234):
        self.chain, self.noise_fn = chain, noise_fn
        self.aggresiveness = 0.2
        self.target_acceptance_rate = target_acceptance_rate
        self.beta_feedback_gain = 2.1
        self.record_skip = 1
        self.record_count = 0
        self.max_record_history = 1000000
        self.use_beta_adaptation = False            # the sqrt-beta hyper-move is not distributed (single-GPU path only)
        self.pcn_step_sqrtBetas = 5e-1
        self.pcn_step_sqrtBetas_Z = 0.5 ** 0.5
        self.last_step = None

    beta = property(lambda self: self.chain.beta, lambda self, v: setattr(self.chain, "beta", float(v)))
    betaZ = property(lambda self: self.chain.betaZ, lambda self, v: setattr(self.chain, "betaZ", float(v)))

    def one_step(self, Layers):
        """pCN.one_step (553-617) through DistPCN.one_step; records the current half-samples of every layer (609-613).
        Returns (accepted, 0); numpy.linalg.LinAlgError propagates to Simulation.run like on the single-GPU path."""
        self.last_step = self.chain.one_step(*self.noise_fn())
        if self.record_count % self.record_skip == 0:
            halves = self.chain.current_halves().cpu().numpy()
            for k, l in enumerate(Layers):
                l.record_sample(halves[k])
                l.record_sqrt_beta()
        self.record_count += 1
        return int(self.last_step["accepted"]), 0


def _borrow_step_rules():
    from . import image as im
    for name in ("adapt_step", "more_aggresive", "less_aggresive", "set_step", "set_step_sqrtBeta", "adapt_step_sqrtBeta"):
        setattr(_DistStepper, name, getattr(im.pCN, name))


class DistSimulation:
    """``Simulation(..., distributed=True)``: the reference's ``Simulation`` surface (image_cupy.py:810-1014: constructor
    arguments, ``run``, ``Layers[k].samples_history``, ``acceptancePercentage``, ``total_time``) for ONE chain whose dense work
    is block-distributed over the ranks of the default process group (BASELINE config 5).  Launch one process per GPU under
    torchrun; every rank builds the same inputs and draws the same noise, steps the chain collectively, and holds the same
    histories; rank 0 writes the result file.

    ``run`` IS ``mspde.image.Simulation.run`` (pinned to the reference's loop by tests/test_run_loop.py); the stepper underneath
    is DistPCN (tests/test_gpu_dist.py).  Not carried over: the sqrt-beta hyper-move and the naive Phi branch."""

    def __init__(self, n_layers, n_samples, n, n_extended, beta, kappa, sigma_0, sigma_v, sigma_scaling, meas_std,
                 evaluation_interval, printProgress, seed, burn_percentage, enable_step_adaptation, pcn_variant="dunlop",
                 phantom_name="shepp.png", meas_type="tomo", n_theta=50, verbose=False, hybrid_GPU_CPU=False, device=0,
                 chol_epsilon=1e-6):
        if meas_type != "tomo" or hybrid_GPU_CPU or phantom_name != "shepp.png":
            raise NotImplementedError("the distributed chain covers the Shepp-Logan tomography measurement on GPUs only")
        dev = torch.device("cuda", device)
        chain, noise_fn = build_dist_chain(n, n_extended, n_theta, n_layers, dev, seed=seed, beta=beta, meas_std=meas_std,
                                           chol_epsilon=chol_epsilon, kappa=kappa, sigma_0=sigma_0, sigma_v=sigma_v,
                                           sigma_scaling=sigma_scaling)
        self._finish(chain, noise_fn, n_layers, n_samples, evaluation_interval, printProgress, burn_percentage,
                     enable_step_adaptation, seed=seed, pcn_variant=pcn_variant, verbose=verbose)

    @classmethod
    def from_chain(cls, chain, noise_fn, n_samples, evaluation_interval=10, printProgress=False, burn_percentage=25.0,
                   enable_step_adaptation=True, **kw):
        """Wrap an existing DistPCN (or any object with its beta / betaZ / one_step / current_halves / sqrt_betas / nr / K)."""
        self = object.__new__(cls)
        self._finish(chain, noise_fn, chain.K, n_samples, evaluation_interval, printProgress, burn_percentage,
                     enable_step_adaptation, **kw)
        return self

    def _finish(self, chain, noise_fn, n_layers, n_samples, evaluation_interval, printProgress, burn_percentage,
                enable_step_adaptation, seed=None, pcn_variant="dunlop", verbose=False):
        if not hasattr(_DistStepper, "adapt_step"):
            _borrow_step_rules()
        self.chain = chain
        self.n_layers, self.n_samples = n_layers, n_samples
        self.evaluation_interval, self.printProgress = evaluation_interval, printProgress
        self.burn_percentage, self.enable_step_adaptation = burn_percentage, enable_step_adaptation
        self.random_seed, self.pcn_variant, self.verbose = seed, pcn_variant, verbose
        self.acceptancePercentage = 0.
        self.accepted_count = 0
        self.accepted_count_SqrtBeta = 0
        self.total_time = 0.0
        self.pcn = _DistStepper(chain, noise_fn)
        self.pcn.record_skip = max(1, n_samples // self.pcn.max_record_history)                 # 899
        history_length = min(n_samples, self.pcn.max_record_history)                            # 900
        self.Layers = [_HistoryLayer(k, chain.sqrt_betas[k], history_length, chain.nr) for k in range(n_layers)]

    def run(self):
        from . import image as im
        return im.Simulation.run(self)

    def save(self, file_name):
        """Rank 0 writes the histories and the run's scalars (key names of util._save_object where they exist, .npz)."""
        import numpy as np
        if dist.is_initialized() and dist.get_rank() != 0:
            return
        out = {"n_layers": self.n_layers, "n_samples": self.n_samples, "evaluation_interval": self.evaluation_interval,
               "burn_percentage": self.burn_percentage, "acceptancePercentage": self.acceptancePercentage,
               "accepted_count": self.accepted_count, "total_time": self.total_time, "pcn/beta": self.pcn.beta,
               "pcn/betaZ": self.pcn.betaZ, "pcn/record_skip": self.pcn.record_skip}
        for k, l in enumerate(self.Layers):
            out[f"Layers {k}/samples_history"] = l.samples_history[:l.i_record]
            out[f"Layers {k}/sqrt_beta"] = l.sqrt_beta
            out[f"Layers {k}/sqrt_beta_history"] = l.sqrt_beta_history[:l.i_record]
        np.savez(file_name, **out)
