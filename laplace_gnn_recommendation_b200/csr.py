"""Device-resident CSR graph + the fused SpMM call (host side of lgb_csr_build / lgb_spmm).

HBM layout per graph (int32 indices, fp32 values):
    rowptr[n_rows+1] | colidx[nnz] | val[nnz] (optional) | split plan (long_rows, long_ptr, task_row,
    task_start, task_end) | optional degree-bucketed row_order[n_rows] | per-d partial-sum scratch [n_tasks, d]
The transposed graph (CSC arrays, used by every backward) is a second DeviceCSR built lazily and cached.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import torch

from . import _lib
from ._lib import LgbCsr, check, ptr, stream

# kernel variant used when neither the call nor autotune() names one (0 = the default; see LGB_SPMM_VARIANT_SHIFT in the header)
SPMM_VARIANT = int(os.environ.get("LGB_SPMM_VARIANT", "0"))
# rows with more non-zeros are split into chunk-sized slices (csrc/spmm.cu): bounds the length of any sequential fp32
# accumulation chain (accuracy) and the work of one warp (load balance)
DEFAULT_CHUNK = int(os.environ.get("LGB_SPMM_CHUNK", "1024"))
# execute long-row slices in column-sweep order by default (DeviceCSR.use_sweep_order; autotune decides per graph when it runs)
DEFAULT_SWEEP = os.environ.get("LGB_SPMM_SWEEP", "0") == "1"
STAGE2_SEG = 32                                                   # partial rows per warp of the stage-2 tree
# stage 2 (sum of the partial rows of a sliced row): flat = one CTA per long row (round 1) | tree = one warp per 32 partial rows,
# a launch of its own | fused = the tree inside the main launch (default: -2 % epoch time at N = 1, -8 % per SpMM pair on a 1/8
# shard, profiles/README.md r2r; kernel families without the fused path fall back to the tree)
STAGE2_MODE = os.environ.get("LGB_SPMM_STAGE2", "fused")
STAGE2_TREE = STAGE2_MODE != "flat"
STAGE2_FUSED = STAGE2_MODE == "fused"                              # the tree inside the main launch (LGB_SPMM_FUSED_STAGE2)


# candidates of DeviceCSR.autotune for d <= 64: the default sub-warp kernel, its CTA-wide-slice and chain-shortening forms,
# the warp-per-row kernel it replaced, the sub-warp kernel at gather-unroll 4, and the four-rows-per-warp forms (20, 22: d in 33..64;
# 23, 25: with 256-bit loads, d = 64; other widths run the default under those numbers)
AUTOTUNE_CANDIDATES = (0, 16, 18, 19, 12, 13, 20, 22, 23, 25)
# the hot-column-cache kernels (variants 30 / 31, d in {32, 64}) with the number of operand rows their plan keeps in shared
# memory (64 KB x two CTAs per SM, 128 / 192 KB x one CTA per SM at d = 64).  A/B'd on the B200 in round 2 and 1.3-1.6x SLOWER
# than the sub-warp kernel on both halves of the H&M graph (the L1 already keeps the hottest rows; persistent 1024-thread CTAs
# lose the block scheduler's fine-grained balancing): kept as variants, tried by autotune only with LGB_SPMM_TRY_HOT=1 / hot=
HOT_CANDIDATES = ((30, 256), (31, 512), (31, 768))
# variants served by the sub-warp kernel family at d <= 64 (the family that carries the fused stage 2); launch accounting only
FUSED_STAGE2_VARIANTS = frozenset((0, 13, 14, 15, 16, 18, 19, 20, 21, 22, 23, 24, 25))


def _time_ms(fn, reps: int, device) -> float:
    """Mean duration of back-to-back calls, CUDA events on the current stream.  One warm-up call, one timed call to size
    the sample (at least ``reps`` launches, more for short kernels so that ~10 ms are measured)."""
    def timed(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) / n
    fn()
    first = timed(1)
    n = int(min(max(reps, 10.0 / max(first, 1e-3)), 200))
    return timed(n)


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def rows_bitmap(n_rows: int, index_lists, device, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Bitmap of row ids for ``DeviceCSR.spmm(x_rows=)``: bit r set for every r = idx + offset of the ``(idx int64, offset)``
    pairs in ``index_lists`` (ids outside [0, n_rows) are ignored).  One memset + one tiny launch per list (lgb_rows_bitmap)."""
    lib = _lib.load()
    words = (int(n_rows) + 31) // 32
    if out is None:
        out = torch.empty(max(words, 1), dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        check(lib.lgb_zero(ptr(out), out.numel() * 4, stream()), "zero")
        for idx, offset in index_lists:
            idx = _lib.i64c(idx)
            _lib.require_cuda(idx)
            check(lib.lgb_rows_bitmap(ptr(idx), idx.numel(), int(offset), int(n_rows), ptr(out), stream()), "rows_bitmap")
            _lib.count_launch()
    return out


class DeviceCSR:
    """An [n_rows, n_cols] sparse matrix in CSR form living in HBM."""

    def __init__(self, n_rows: int, n_cols: int, rowptr: torch.Tensor, colidx: torch.Tensor,
                 val: Optional[torch.Tensor] = None, chunk: int = DEFAULT_CHUNK):
        _lib.require_cuda(rowptr, colidx, val)
        assert rowptr.dtype == torch.int32 and colidx.dtype == torch.int32
        self.n_rows, self.n_cols, self.nnz = int(n_rows), int(n_cols), int(colidx.numel())
        self.rowptr, self.colidx, self.val = rowptr, colidx, val
        self.device = rowptr.device
        self.chunk = int(chunk)
        self.row_order: Optional[torch.Tensor] = None
        self.n_long = 0
        self.n_tasks = 0
        self.long_rows = self.long_ptr = self.task_row = self.task_start = self.task_end = None
        self.n_seg = 0
        self.seg_row = self.seg_t0 = self.seg_t1 = self.row_seg0 = self.task_seg = None
        self.sweep = DEFAULT_SWEEP                         # execute the long-row slices in column-sweep order: use_sweep_order()
        self.task_exec: Optional[torch.Tensor] = None
        self.perm: Optional[torch.Tensor] = None      # COO -> CSR permutation (int64) when built from COO
        self.csr2csc: Optional[torch.Tensor] = None   # set on the TRANSPOSED graph: its entry i is CSR entry csr2csc[i]
        self._t: Optional["DeviceCSR"] = None
        self._partials: Dict[tuple, torch.Tensor] = {}
        self._struct: Optional[LgbCsr] = None
        self.n_hot = 0                                  # hot-column plan (variants 30 / 31): see set_hot()
        self.hot_cols: Optional[torch.Tensor] = None
        self.colidx_hot: Optional[torch.Tensor] = None
        self.variant: Optional[int] = None             # kernel variant chosen by autotune() for THIS graph (None: the default)
        self.autotune_report: Optional[dict] = None
        if self.chunk > 0:
            self._build_plan()

    # ---- construction -------------------------------------------------------------------
    @classmethod
    def from_coo(cls, row: torch.Tensor, col: torch.Tensor, n_rows: int, n_cols: int,
                 chunk: int = DEFAULT_CHUNK, want_perm: bool = False, validate: bool = True) -> "DeviceCSR":
        """SparseTensor(row, col, sparse_sizes) semantics (reference data/lightgcn_loader.py:65-79), including its
        range assertion: an index outside ``sparse_sizes`` raises (``validate`` costs one stream synchronisation; a caller
        that builds several graphs per batch may pass False and call ``check_built()`` once on the last one -- out-of-range
        entries are clamped on the device either way, so nothing is ever written out of bounds)."""
        _lib.require_cuda(row, col)
        lib = _lib.load()
        row, col = _lib.i64c(row), _lib.i64c(col)
        nnz = row.numel()
        dev = row.device
        need = C.c_size_t(0)
        check(lib.lgb_csr_build_ws_bytes(nnz, n_rows, C.byref(need)), "csr_build_ws_bytes")
        ws = _ws(need.value, dev)
        rowptr = torch.empty(n_rows + 1, dtype=torch.int32, device=dev)
        colidx = torch.empty(nnz, dtype=torch.int32, device=dev)
        perm = torch.empty(nnz, dtype=torch.int64, device=dev) if want_perm else None
        with torch.cuda.device(dev):
            check(lib.lgb_csr_build(ptr(row), ptr(col), nnz, n_rows, n_cols, ptr(rowptr), ptr(colidx), ptr(perm),
                                    ptr(ws), ws.numel(), stream()), "csr_build")
            if validate:
                check(lib.lgb_csr_build_check(ptr(ws), stream()), "csr_build (index range)")
        _lib.count_launch(3)
        g = cls(n_rows, n_cols, rowptr, colidx, None, chunk)
        g.perm = perm
        g._build_ws = None if validate else ws          # keeps the status word alive until check_built()
        return g

    def check_built(self) -> None:
        """Deferred index-range check of a graph built with ``validate=False`` (synchronises the current stream)."""
        ws = getattr(self, "_build_ws", None)
        if ws is not None:
            with torch.cuda.device(self.device):
                check(_lib.load().lgb_csr_build_check(ptr(ws), stream()), "csr_build (index range)")
            self._build_ws = None

    def _build_plan(self) -> None:
        lib = _lib.load()
        need = C.c_size_t(0)
        check(lib.lgb_spmm_plan_ws_bytes(self.n_rows, C.byref(need)), "spmm_plan_ws_bytes")
        ws = _ws(need.value, self.device)
        counts = (C.c_int64 * 2)()
        with torch.cuda.device(self.device):
            check(lib.lgb_spmm_plan_count(ptr(self.rowptr), self.n_rows, self.chunk, counts, ptr(ws), ws.numel(),
                                          stream()), "spmm_plan_count")
            self.n_long, self.n_tasks = int(counts[0]), int(counts[1])
            if self.n_long:
                i32 = dict(dtype=torch.int32, device=self.device)
                self.long_rows = torch.empty(self.n_long, **i32)
                self.long_ptr = torch.empty(self.n_long + 1, **i32)
                self.task_row = torch.empty(self.n_tasks, **i32)
                self.task_start = torch.empty(self.n_tasks, **i32)
                self.task_end = torch.empty(self.n_tasks, **i32)
                check(lib.lgb_spmm_plan_fill(ptr(self.rowptr), self.n_rows, self.chunk, self.n_long, self.n_tasks,
                                             ptr(self.long_rows), ptr(self.long_ptr), ptr(self.task_row),
                                             ptr(self.task_start), ptr(self.task_end), ptr(ws), ws.numel(), stream()),
                      "spmm_plan_fill")
                self._build_segments()
        self._build_sweep()
        self._struct = None

    def use_sweep_order(self, on: bool = True) -> "DeviceCSR":
        """Execute the long-row slices sorted by their first column instead of row by row.  Rows are stored with ascending
        columns, so a slice covers a narrow column range; in sweep order the slices of ALL long rows that gather the same band
        of operand rows run next to each other, and the band is fetched from DRAM once instead of once per long row (the
        351 MB user table of the H&M-shaped item rows does not fit the 126 MB L2).  Results are bit-identical: partial sums
        stay indexed by slice id."""
        self.sweep = bool(on)
        self._build_sweep()
        self._struct = None
        return self

    def _build_sweep(self) -> None:
        self.task_exec = None
        if self.sweep and self.n_tasks > 1:
            first = self.colidx[self.task_start.long()]
            self.task_exec = torch.argsort(first, stable=True).to(torch.int32)

    def _build_segments(self) -> None:
        """Stage-2 tree plan: the partial rows of every long row cut into segments of 32 (one warp each, csrc/spmm.cu
        spmm_long_reduce_tree_kernel).  Plan-time index arithmetic on the device (torch ops)."""
        lp = self.long_ptr.long()
        nseg = (lp[1:] - lp[:-1] + STAGE2_SEG - 1) // STAGE2_SEG
        row_seg0 = torch.zeros(self.n_long + 1, dtype=torch.int64, device=self.device)
        row_seg0[1:] = torch.cumsum(nseg, 0)
        self.n_seg = int(row_seg0[-1])
        seg_row = torch.repeat_interleave(torch.arange(self.n_long, device=self.device), nseg)
        first = lp[seg_row] + STAGE2_SEG * (torch.arange(self.n_seg, device=self.device) - row_seg0[seg_row])
        i32 = torch.int32
        self.seg_row, self.seg_t0 = seg_row.to(i32), first.to(i32)
        self.seg_t1 = torch.minimum(first + STAGE2_SEG, lp[seg_row + 1]).to(i32)
        self.row_seg0 = row_seg0.to(i32)
        # segment of every slice (fused stage 2): the segments tile the slices in order
        self.task_seg = torch.repeat_interleave(torch.arange(self.n_seg, device=self.device), (self.seg_t1 - self.seg_t0).long()).to(i32)

    def use_degree_order(self, on: bool = True) -> "DeviceCSR":
        """Process ordinary rows in descending degree-bucket order (better intra-CTA balance on skewed graphs)."""
        if not on:
            self.row_order = None
        elif self.row_order is None and self.n_rows:
            lib = _lib.load()
            need = C.c_size_t(0)
            check(lib.lgb_degree_order_ws_bytes(self.n_rows, C.byref(need)), "degree_order_ws_bytes")
            ws = _ws(need.value, self.device)
            self.row_order = torch.empty(self.n_rows, dtype=torch.int32, device=self.device)
            with torch.cuda.device(self.device):
                check(lib.lgb_degree_order(ptr(self.rowptr), self.n_rows, ptr(self.row_order), ptr(ws), ws.numel(),
                                           stream()), "degree_order")
        self._struct = None
        return self

    # ---- plan-time kernel selection --------------------------------------------------------
    def autotune(self, d: int, candidates=None, reps: int = 5, fused_epilogue: bool = True, chunks=None,
                 degree_orders=(False,), hot=None, sweeps=(False,)) -> int:
        """Pick the fastest lgb_spmm configuration for THIS graph and width on THIS device (one-off, like an FFT plan):
        kernel variant x slice size of the long-row plan (``chunks``, default: the current one) x row processing order
        (``degree_orders``) x slice execution order (``sweeps``, see use_sweep_order).  Every configuration is first checked against the first one's result on random data (one that
        disagrees beyond fp32 summation-order noise is dropped and reported), then timed with CUDA events over the call
        shape the LightGCN layers use (Y + fused accumulate).  The winner is installed on the graph (plan, row order,
        ``self.variant``) and used by every later ``spmm`` that does not name a variant.  Returns the chosen variant."""
        if candidates is None:
            candidates = AUTOTUNE_CANDIDATES if d % 4 == 0 and d <= 64 else ((0, 26, 27) if d == 128 else (0,))
            if hot is None and d in (32, 64) and os.environ.get("LGB_SPMM_TRY_HOT", "0") == "1":
                hot = tuple((v, h * 64 // d) for v, h in HOT_CANDIDATES)       # measured slower on B200 (profiles/README.md r2d): opt-in
        # every candidate is a (variant, rows of the hot-column plan) pair; 0 rows = no plan
        candidates = [(v, 0) for v in candidates] + [(int(v), int(h)) for v, h in (hot or ())]
        chunks = tuple(chunks) if chunks else (self.chunk,)
        if self.chunk <= 0:
            chunks = (self.chunk,)
        dev = self.device
        report = {"d": d, "ms": {}, "rejected": {}}
        sweeps = tuple(bool(x) for x in sweeps) or (False,)
        n_configs = len(candidates) * len(chunks) * len(degree_orders) * len(sweeps)
        if self.n_rows == 0 or self.nnz == 0 or n_configs <= 1:
            self.variant = candidates[0][0] if candidates else None
            self.set_hot(candidates[0][1] if candidates else 0)
            self.use_sweep_order(sweeps[0])
            report["chosen"] = {"variant": self.variant, "chunk": self.chunk, "degree_order": self.row_order is not None,
                                "sweep": self.sweep}
            self.autotune_report = report
            return self.variant
        gen = torch.Generator(device=dev).manual_seed(1234)
        X = torch.randn(self.n_cols, d, device=dev, generator=gen)
        acc = torch.randn(self.n_rows, d, device=dev, generator=gen) if fused_epilogue else None
        Y = torch.empty(self.n_rows, d, device=dev)
        out = torch.empty(self.n_rows, d, device=dev) if fused_epilogue else None
        ref = ref_mag = None
        best, best_ms = (chunks[0], bool(degree_orders[0]), candidates[0], sweeps[0]), float("inf")
        for chunk in chunks:
            self._set_chunk(chunk)
            for order in degree_orders:
                self.use_degree_order(bool(order))
                for sweep in sweeps:
                    if sweep and self.n_tasks <= 1 and len(sweeps) > 1:
                        continue                                          # nothing to reorder: same launch as sweep off
                    self.use_sweep_order(sweep)
                    for cand in candidates:
                        v, n_hot = cand
                        key = (f"v{v}" + (f"h{n_hot}" if n_hot else "") + (f"/chunk{chunk}" if len(chunks) > 1 else "")
                               + ("/degree-order" if order else "") + ("/sweep" if sweep else ""))
                        try:
                            self.set_hot(n_hot)
                            run = (lambda: self.spmm(X, Y=Y, acc_in=acc, acc_out=out, variant=v)) if fused_epilogue else (lambda: self.spmm(X, Y=Y, variant=v))   # noqa: E731
                            run()
                            if ref is None:
                                ref = Y.clone()
                                ref_mag = float(ref.abs().max()) + 1e-30
                            else:
                                err = float((Y - ref).abs().max())
                                if not err <= 1e-4 * ref_mag:            # also catches NaN
                                    report["rejected"][key] = f"max |diff| {err:.3e} vs magnitude {ref_mag:.3e}"
                                    continue
                            ms = _time_ms(run, reps, dev)
                        except RuntimeError as exc:                       # a variant that is not available for this shape
                            report["rejected"][key] = str(exc)[:200]
                            continue
                        report["ms"][key] = ms
                        if ms < best_ms:
                            best, best_ms = (chunk, bool(order), cand, sweep), ms
        self._set_chunk(best[0])
        self.use_degree_order(best[1])
        self.use_sweep_order(best[3])
        self.set_hot(best[2][1])
        self.variant, self.autotune_report = best[2][0], report
        report["chosen"] = {"variant": best[2][0], "chunk": best[0], "degree_order": best[1], "sweep": best[3]}
        if best[2][1]:
            report["chosen"].update(hot_rows=best[2][1], hot_share=round(getattr(self, "hot_share", 0.0), 4))
        return self.variant

    def _set_chunk(self, chunk: int) -> None:
        """Rebuild the long-row plan for another slice size."""
        if int(chunk) != self.chunk:
            self.chunk = int(chunk)
            self.n_long = self.n_tasks = self.n_seg = 0
            self.long_rows = self.long_ptr = self.task_row = self.task_start = self.task_end = self.task_exec = None
            self.seg_row = self.seg_t0 = self.seg_t1 = self.row_seg0 = self.task_seg = None
            if self.chunk > 0:
                self._build_plan()
            self._struct = None

    def set_hot(self, n_hot: int) -> "DeviceCSR":
        """Build (or drop, n_hot = 0) the hot-column plan of variants 30 / 31: the n_hot most-referenced columns of THIS matrix
        (by column degree over its own entries) and a copy of colidx in which such a column c appears as ~slot.  Plan-time
        work (torch ops: bincount, topk, one gather); the hot path only reads the two arrays."""
        n_hot = int(min(max(n_hot, 0), self.n_cols))
        if n_hot == self.n_hot and (n_hot == 0 or self.colidx_hot is not None):
            return self
        self.n_hot, self.hot_cols, self.colidx_hot, self._struct = 0, None, None, None
        if n_hot == 0 or self.nnz == 0:
            return self
        lo, hi = int(self.rowptr[0]), int(self.rowptr[-1])          # a row view shares colidx with its parent: its own range only
        cols = self.colidx[lo:hi].long()
        deg = torch.bincount(cols, minlength=self.n_cols)
        hot = torch.topk(deg, n_hot).indices
        slot = torch.full((self.n_cols,), -1, dtype=torch.int64, device=self.device)
        slot[hot] = torch.arange(n_hot, device=self.device)
        sl = slot[cols]
        colh = self.colidx.clone()
        colh[lo:hi] = torch.where(sl >= 0, -sl - 1, cols).to(torch.int32)
        self.n_hot, self.hot_cols, self.colidx_hot = n_hot, hot.to(torch.int32).contiguous(), colh
        self.hot_share = float(deg[hot].sum()) / max(hi - lo, 1)     # fraction of the entries served from shared memory
        return self

    def with_values(self, val: Optional[torch.Tensor]) -> "DeviceCSR":
        """Same structure (arrays and plan shared), different values."""
        g = DeviceCSR.__new__(DeviceCSR)
        g.__dict__.update(self.__dict__)
        g.val = val
        g._t = None
        g._struct = None
        g._partials = self._partials
        return g

    def transpose(self) -> "DeviceCSR":
        """CSC arrays as the CSR of A^T (cached); values follow through csr2csc."""
        if self._t is not None:
            return self._t
        lib = _lib.load()
        need = C.c_size_t(0)
        check(lib.lgb_csr_transpose_ws_bytes(self.nnz, self.n_cols, C.byref(need)), "csr_transpose_ws_bytes")
        ws = _ws(need.value, self.device)
        i32 = dict(dtype=torch.int32, device=self.device)
        colptr = torch.empty(self.n_cols + 1, **i32)
        rowidx = torch.empty(self.nnz, **i32)
        csr2csc = torch.empty(self.nnz, **i32)
        with torch.cuda.device(self.device):
            check(lib.lgb_csr_transpose(ptr(self.rowptr), ptr(self.colidx), self.n_rows, self.n_cols, self.nnz,
                                        ptr(colptr), ptr(rowidx), ptr(csr2csc), ptr(ws), ws.numel(), stream()),
                  "csr_transpose")
            val_t = None
            if self.val is not None:
                val_t = torch.empty_like(self.val)
                check(lib.lgb_gather_f32(ptr(self.val), ptr(csr2csc), self.nnz, ptr(val_t), stream()), "gather_f32")
        _lib.count_launch(4)
        t = DeviceCSR(self.n_cols, self.n_rows, colptr, rowidx, val_t, self.chunk)
        t.csr2csc = csr2csc
        if self.row_order is not None:
            t.use_degree_order()
        if self.sweep:
            t.use_sweep_order()
        t._t = self
        self._t = t
        return t

    def gcn_norm(self):
        """(dinv, val) of gcn_norm(add_self_loops=False) -- reference model/lightgcn.py:56."""
        if self.n_rows != self.n_cols:
            raise RuntimeError("gcn_norm needs a square adjacency")
        lib = _lib.load()
        dinv = torch.empty(self.n_rows, dtype=torch.float32, device=self.device)
        val = torch.empty(self.nnz, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib.lgb_gcn_norm(ptr(self.rowptr), ptr(self.colidx), self.n_rows, self.nnz, ptr(dinv), ptr(val),
                                   stream()), "gcn_norm")
        _lib.count_launch(2)
        return dinv, val

    # ---- C struct -----------------------------------------------------------------------
    @property
    def struct(self) -> LgbCsr:
        if self._struct is None:
            s = LgbCsr()
            s.n_rows, s.n_cols, s.nnz = self.n_rows, self.n_cols, self.nnz
            s.rowptr, s.colidx, s.val, s.row_order = ptr(self.rowptr), ptr(self.colidx), ptr(self.val), ptr(self.row_order)
            s.chunk = self.chunk
            s.n_long, s.n_tasks = self.n_long, self.n_tasks
            s.long_rows, s.long_ptr = ptr(self.long_rows), ptr(self.long_ptr)
            s.task_row, s.task_start, s.task_end = ptr(self.task_row), ptr(self.task_start), ptr(self.task_end)
            s.colidx_hot, s.hot_cols, s.n_hot = ptr(self.colidx_hot), ptr(self.hot_cols), int(self.n_hot)
            s.seg_row, s.seg_t0, s.seg_t1, s.row_seg0 = ptr(self.seg_row), ptr(self.seg_t0), ptr(self.seg_t1), ptr(self.row_seg0)
            s.n_seg = int(self.n_seg)
            s.task_exec = ptr(self.task_exec)
            s.task_seg = ptr(self.task_seg)
            self._struct = s
        return self._struct

    def _partial_ws(self, d: int) -> Optional[torch.Tensor]:
        """Partial-sum scratch of the long-row slices: one per (width, STREAM) -- the sharded engine runs the same graph on two
        streams at once (layer k of one chain next to layer k+1 of the other), and launches that overlap must not share it."""
        if self.n_tasks == 0 and self.n_hot == 0:
            return None
        key = (d, self.n_tasks, torch.cuda.current_stream(self.device).cuda_stream if self.device.type == "cuda" else 0)
        buf = self._partials.get(key)
        if buf is None:
            # behind the partial sums: 64 floats for the work counters of the hot-column kernels (variants 30 / 31), the level-2
            # rows and the tickets of the stage-2 tree; all of it starts at zero and every launch leaves counters / tickets zero
            # ... and the segment tickets of the fused stage 2
            buf = torch.zeros(self.n_tasks * d + 64 + self.n_seg * d + self.n_long + self.n_seg + 64, dtype=torch.float32, device=self.device)
            self._partials[key] = buf
        return buf

    # ---- the hot call -------------------------------------------------------------------
    def spmm(self, X: torch.Tensor, Y: Optional[torch.Tensor] = None, resid: Optional[torch.Tensor] = None,
             acc_in: Optional[torch.Tensor] = None, acc_out: Optional[torch.Tensor] = None, acc_div: float = 1.0,
             mean: bool = False, want_y: bool = True, variant: Optional[int] = None, split_row: Optional[int] = None,
             y_tail: Optional[torch.Tensor] = None, x_rows: Optional[torch.Tensor] = None,
             resid_rows: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
        """Y = A @ X with the fused epilogue of lgb_spmm.  Allocates Y when want_y and Y is None.
        ``x_rows``: bitmap (int32 words, ``rows_bitmap``) of the rows of X that may be non-zero -- lgb_spmm_rowsparse: same
        result, entries that multiply a zero row are never gathered.  ``resid_rows``: the same kind of bitmap for ``resid`` (used
        with ``x_rows`` only): rows that are not flagged are not read."""
        _lib.require_cuda(X)
        if X.dim() != 2 or X.shape[0] != self.n_cols:
            raise RuntimeError(f"spmm: X has shape {tuple(X.shape)}, expected [{self.n_cols}, d]")
        X = _lib.f32c(X)
        d = X.shape[1]
        if Y is None and want_y:
            Y = torch.empty(self.n_rows, d, dtype=torch.float32, device=self.device)
        for name, t in (("Y", Y), ("resid", resid), ("acc_in", acc_in), ("acc_out", acc_out)):
            if t is not None and (tuple(t.shape) != (self.n_rows, d) or not t.is_contiguous() or t.dtype != torch.float32):
                raise RuntimeError(f"spmm: {name} must be contiguous float32 [{self.n_rows}, {d}]")
        lib = _lib.load()
        if variant is None:
            variant = SPMM_VARIANT if self.variant is None else self.variant
        if variant in (30, 31) and self.n_hot == 0 and os.environ.get("LGB_SPMM_HOT"):
            self.set_hot(int(os.environ["LGB_SPMM_HOT"]))      # a variant pinned from the environment brings its plan along
        flags = (1 if mean else 0) | (2 if (STAGE2_TREE and self.n_seg) else 0) | (4 if (STAGE2_FUSED and self.n_seg) else 0) | (variant << 4)
        with torch.cuda.device(self.device):
            if y_tail is not None:
                # split epilogue: rows >= split_row write raw sums to y_tail (see lgb_spmm_split)
                if tuple(y_tail.shape) != (self.n_rows - int(split_row), d) or not y_tail.is_contiguous():
                    raise RuntimeError(f"spmm: y_tail must be contiguous float32 [{self.n_rows - int(split_row)}, {d}]")
                check(lib.lgb_spmm_split(C.byref(self.struct), ptr(X), d, ptr(Y), ptr(resid), ptr(acc_in), ptr(acc_out),
                                         float(acc_div), flags, ptr(self._partial_ws(d)), int(split_row), ptr(y_tail),
                                         stream()), "spmm_split")
            elif x_rows is not None:
                if x_rows.dtype != torch.int32 or x_rows.numel() * 32 < self.n_cols or not x_rows.is_contiguous():
                    raise RuntimeError(f"spmm: x_rows must be a contiguous int32 bitmap of >= {self.n_cols} bits")
                _lib.require_cuda(x_rows)
                if resid_rows is not None and (resid_rows.dtype != torch.int32 or resid_rows.numel() * 32 < self.n_rows
                                               or not resid_rows.is_contiguous()):
                    raise RuntimeError(f"spmm: resid_rows must be a contiguous int32 bitmap of >= {self.n_rows} bits")
                check(lib.lgb_spmm_rowsparse(C.byref(self.struct), ptr(X), ptr(x_rows), d, ptr(Y), ptr(resid), ptr(resid_rows), ptr(acc_in),
                                             ptr(acc_out), float(acc_div), flags, ptr(self._partial_ws(d)), stream()),
                      "spmm_rowsparse")
            else:
                check(lib.lgb_spmm(C.byref(self.struct), ptr(X), d, ptr(Y), ptr(resid), ptr(acc_in), ptr(acc_out),
                                   float(acc_div), flags, ptr(self._partial_ws(d)), stream()), "spmm")
        fused = STAGE2_FUSED and self.n_seg and variant in FUSED_STAGE2_VARIANTS and d % 4 == 0 and d <= 64
        _lib.count_launch(2 if (self.n_long and not fused) else 1)
        return Y
