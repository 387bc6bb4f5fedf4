"""Host-side mirror of the reference's likelihood functions, backed by the CUDA kernels.

`MapleEngine` keeps the same names, argument meaning and "impossible" conventions as the
reference functions it stands in for (MAPLEv0.7.5.4.py): ``appendProbNode`` (:6505) returns a float
or ``-inf``; ``mergeVectors`` (:4446) returns a list, ``None`` or ``(list, lk)``;
``estimateBranchLengthWithDerivative`` (:5040) returns a float or ``False``;
``areVectorsDifferent`` (:5419) returns a bool.  The batch forms are what the search drivers use;
the single-call forms pack their arguments, run a batch of one and unpack, so that parity tests
can be written exactly like calls into the reference.

PyTorch is used for device memory and streams only.  Nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import capi
from .genome_list import PackedLists, pack_lists, decode_stream
from .model import MapleModel


def _dp(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class DeviceLists:
    """A PackedLists arena resident in HBM (torch.cuda tensors)."""

    def __init__(self, packed: PackedLists, device: torch.device):
        self.lRef, self.U = packed.lRef, packed.U
        self.key = torch.from_numpy(packed.key.view(np.int32)).to(device)
        self.pay = torch.from_numpy(packed.pay).to(device)
        self.key_start = torch.from_numpy(packed.key_start).to(device)
        self.pay_start = torch.from_numpy(packed.pay_start).to(device)
        self.nkeys = torch.from_numpy(packed.nkeys).to(device)
        self.n = len(packed)

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.key, self.pay, self.key_start, self.pay_start, self.nkeys))


class MergeResult:
    """Device-side output of merge_batch: one slot per pair in a scratch arena."""

    def __init__(self, key, pay, key_start, pay_start, nkeys, npay, lk, status, lRef, U):
        self.key, self.pay, self.key_start, self.pay_start = key, pay, key_start, pay_start
        self.nkeys, self.npay, self.lk, self.status = nkeys, npay, lk, status
        self.lRef, self.U = lRef, U

    def to_lists(self):
        key = self.key.cpu().numpy().view(np.uint32)
        pay = self.pay.cpu().numpy()
        ks, ps = self.key_start.cpu().numpy(), self.pay_start.cpu().numpy()
        st, nk = self.status.cpu().numpy(), self.nkeys.cpu().numpy()
        return [None if st[i] != 0 else decode_stream(key, pay, ks[i], ps[i], self.lRef, self.U, int(nk[i])) for i in range(len(st))]


class MapleEngine:
    def __init__(self, model: MapleModel, device: int = 0):
        self.lib = capi.load()
        if not torch.cuda.is_available():
            raise capi.MapleError("no CUDA device visible; maple_b200 has no CPU fallback")
        self.model = model
        self.device = torch.device("cuda", device)
        flags = ((capi.MAPLE_F_USING_ERROR_RATE if model.usingErrorRate else 0)
                 | (capi.MAPLE_F_ERROR_SITE_SPECIFIC if model.errorRateSiteSpecific else 0)
                 | (capi.MAPLE_F_RATE_VARIATION if model.useRateVariation else 0))
        ctx = C.c_void_p()
        pi = (C.c_double * 4)(*[float(x) for x in model.rootFreqs])
        rc = self.lib.maple_ctx_create(C.byref(ctx), device, model.lRef, pi, flags)
        capi.check(None, rc, "maple_ctx_create")
        self.ctx = ctx
        try:
            self.num_sms = int(torch.cuda.get_device_properties(self.device).multi_processor_count)
        except Exception:
            self.num_sms = 148
        self.lists: Optional[DeviceLists] = None
        self.update_model()

    def __del__(self):
        try:
            if getattr(self, "ctx", None):
                self.lib.maple_ctx_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass

    # ---------------------------------------------------------------- model / lists
    def update_model(self):
        m = self.model
        Q = (C.c_double * 16)(*[float(x) for x in m.Q.reshape(-1)])

        def hp(a):
            return None if a is None else a.ctypes.data_as(C.c_void_p)

        rc = self.lib.maple_ctx_set_model(self.ctx, Q, hp(m.siteRates if m.useRateVariation else None), float(m.errorRate),
                                          hp(m.errorRates if (m.usingErrorRate and m.errorRateSiteSpecific) else None),
                                          hp(m.cumulativeRate), hp(m.cumulativeErrorRate), float(m.totError))
        capi.check(self.ctx, rc, "maple_ctx_set_model")
        rc = self.lib.maple_ctx_set_thresholds(self.ctx, m.thresholdProb, m.thresholdDiffForUpdate, m.thresholdFoldChangeUpdate,
                                               m.minBLenSensitivity)
        capi.check(self.ctx, rc, "maple_ctx_set_thresholds")
        self._root_tables = False

    def _ensure_root_tables(self):
        """cumulativeBases / rootFreqsLogErrorCumulative: only findProbRoot reads them, so they are uploaded on first use."""
        if self._root_tables:
            return
        m = self.model
        cb = np.ascontiguousarray(m.cumulative_bases(), np.int32)
        pl = np.ascontiguousarray(m.root_freqs_log_error_cumulative()) if m.usingErrorRate else None
        rc = self.lib.maple_ctx_set_root_tables(self.ctx, cb.ctypes.data_as(C.c_void_p), None if pl is None else pl.ctypes.data_as(C.c_void_p))
        capi.check(self.ctx, rc, "maple_ctx_set_root_tables")
        self._root_tables = True

    def bind(self, lists) -> DeviceLists:
        if isinstance(lists, PackedLists):
            lists = DeviceLists(lists, self.device)
        self.lists = lists
        rc = self.lib.maple_lists_bind(self.ctx, _dp(lists.key), _dp(lists.pay), _dp(lists.key_start), _dp(lists.pay_start), lists.n)
        capi.check(self.ctx, rc, "maple_lists_bind")
        return lists

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _t(self, a, dtype):
        if isinstance(a, torch.Tensor):
            return a.to(device=self.device, dtype=dtype).contiguous()
        return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(self.device)

    def set_search_variant(self, variant: int):
        """0 = warp-converged state-machine search kernel (default), 1 = straight-line kernel (A/B measurements)."""
        capi.check(self.ctx, self.lib.maple_ctx_set_search_variant(self.ctx, int(variant)), "maple_ctx_set_search_variant")

    def set_place_variant(self, variant: int):
        """0 = one new sample per thread (default), 1 = one per warp with windowed scans (place_scan.cuh), 2 = the same with MAT
        trees covered, 3 = 2 with the parallel window replay; same results."""
        capi.check(self.ctx, self.lib.maple_ctx_set_place_variant(self.ctx, int(variant)), "maple_ctx_set_place_variant")
        self.place_variant = int(variant)

    def set_scan_min_size(self, n: int):
        capi.check(self.ctx, self.lib.maple_ctx_set_scan_min_size(self.ctx, int(n)), "maple_ctx_set_scan_min_size")

    def set_lanes_per_warp(self, lanes: int):
        """Searches a warp of the search kernel runs at a time (0 = chosen per launch)."""
        capi.check(self.ctx, self.lib.maple_ctx_set_lanes_per_warp(self.ctx, int(lanes)), "maple_ctx_set_lanes_per_warp")

    def set_critical_searches(self, count: int):
        """The first `count` entries of every following search batch run on an SM of their own each (0 = off)."""
        capi.check(self.ctx, self.lib.maple_ctx_set_critical_searches(self.ctx, int(count)), "maple_ctx_set_critical_searches")

    def set_head_searches(self, count: int):
        """The first `count` entries of every following search batch are handed out one per warp before all lanes pull (0 = off)."""
        capi.check(self.ctx, self.lib.maple_ctx_set_head_searches(self.ctx, int(count)), "maple_ctx_set_head_searches")

    def set_scan_service(self, fsm_sms: int):
        """SMs whose CTAs own the searches while all others only serve subtree scans (-1 = chosen per launch, 0 = off)."""
        capi.check(self.ctx, self.lib.maple_ctx_set_scan_service(self.ctx, int(fsm_sms)), "maple_ctx_set_scan_service")

    def set_dense_scoring(self, mode: int, max_bytes: int = 0):
        """The dense scoring pass of the search (-1 = in deep rounds when it applies, 0 = never, 1 = whenever it applies) and the
        HBM its score matrix may take (0 = keep the current limit)."""
        capi.check(self.ctx, self.lib.maple_ctx_set_dense_scoring(self.ctx, int(mode), int(max_bytes)), "maple_ctx_set_dense_scoring")

    def search_stats(self, enable: bool = True, read: bool = True):
        """Profiling counters of the search kernel (see scripts/time_search.py for their meaning)."""
        out = (C.c_uint64 * 40)() if read else None
        capi.check(self.ctx, self.lib.maple_search_stats(self.ctx, 1 if enable else 0, out), "maple_search_stats")
        return None if out is None else list(out)

    @property
    def launches(self) -> int:
        return int(self.lib.maple_launch_count(self.ctx))

    # ---------------------------------------------------------------- batch forms (device tensors)
    def append_prob_batch(self, pIdx, cIdx, isTipC, bLen, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        pIdx, cIdx = self._t(pIdx, torch.int32), self._t(cIdx, torch.int32)
        isTipC, bLen = self._t(isTipC, torch.uint8), self._t(bLen, torch.float64)
        n = pIdx.numel()
        if out is None:
            out = torch.empty(n, dtype=torch.float64, device=self.device)
        rc = self.lib.maple_append_prob_batch(self.ctx, n, _dp(pIdx), _dp(cIdx), _dp(isTipC), _dp(bLen), _dp(out), self._stream())
        capi.check(self.ctx, rc, "maple_append_prob_batch")
        return out

    def append_prob_batch_host(self, pIdx: np.ndarray, cIdx: np.ndarray, isTipC: np.ndarray, bLen: np.ndarray,
                               out: Optional[np.ndarray] = None) -> np.ndarray:
        """HOST buffers in, host scores out (copies inside the call)."""
        pIdx, cIdx = np.ascontiguousarray(pIdx, np.int32), np.ascontiguousarray(cIdx, np.int32)
        isTipC, bLen = np.ascontiguousarray(isTipC, np.uint8), np.ascontiguousarray(bLen, np.float64)
        n = len(pIdx)
        if out is None:
            out = np.empty(n, dtype=np.float64)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        rc = self.lib.maple_append_prob_batch_host(self.ctx, n, vp(pIdx), vp(cIdx), vp(isTipC), vp(bLen), vp(out))
        capi.check(self.ctx, rc, "maple_append_prob_batch_host")
        return out

    def merge_batch(self, idx1, bLen1, fromTip1, idx2, bLen2, fromTip2, flags, numMinor1=None, numMinor2=None,
                    shorten: bool = False) -> MergeResult:
        L = self.lists
        idx1, idx2 = self._t(idx1, torch.int32), self._t(idx2, torch.int32)
        bLen1, bLen2 = self._t(bLen1, torch.float64), self._t(bLen2, torch.float64)
        fromTip1, fromTip2 = self._t(fromTip1, torch.uint8), self._t(fromTip2, torch.uint8)
        flags = self._t(flags, torch.uint8)
        nm1 = None if numMinor1 is None else self._t(numMinor1, torch.int32)
        nm2 = None if numMinor2 is None else self._t(numMinor2, torch.int32)
        n = idx1.numel()
        # slot sizing: nkeys1+nkeys2 keys (16-byte aligned), 6 doubles of payload per key
        cap = (L.nkeys[idx1.long()].long() + L.nkeys[idx2.long()].long() + 3) // 4 * 4
        ks = torch.cumsum(cap, 0) - cap
        ps = ks * 6
        total = int(cap.sum().item())
        out_key = torch.empty(total + 4, dtype=torch.int32, device=self.device)
        out_pay = torch.empty(total * 6 + 4, dtype=torch.float64, device=self.device)
        nk = torch.empty(n, dtype=torch.int32, device=self.device)
        npay = torch.empty(n, dtype=torch.int32, device=self.device)
        lk = torch.zeros(n, dtype=torch.float64, device=self.device)
        st = torch.empty(n, dtype=torch.int32, device=self.device)
        rc = self.lib.maple_merge_batch(self.ctx, n, _dp(idx1), _dp(bLen1), _dp(fromTip1), _dp(idx2), _dp(bLen2), _dp(fromTip2),
                                        _dp(flags), _dp(nm1), _dp(nm2), _dp(out_key), _dp(out_pay), _dp(ks), _dp(ps), _dp(nk),
                                        _dp(npay), _dp(lk), _dp(st), 1 if shorten else 0, self._stream())
        capi.check(self.ctx, rc, "maple_merge_batch")
        return MergeResult(out_key, out_pay, ks, ps, nk, npay, lk, st, L.lRef, L.U)

    def blen_batch(self, pIdx, cIdx, fromTipC):
        L = self.lists
        pIdx, cIdx = self._t(pIdx, torch.int32), self._t(cIdx, torch.int32)
        fromTipC = self._t(fromTipC, torch.uint8)
        n = pIdx.numel()
        cap = L.nkeys[pIdx.long()].long() + L.nkeys[cIdx.long()].long()
        ss = torch.cumsum(cap, 0) - cap
        scratch = torch.empty(int(cap.sum().item()) + 1, dtype=torch.float64, device=self.device)
        out = torch.empty(n, dtype=torch.float64, device=self.device)
        st = torch.empty(n, dtype=torch.int32, device=self.device)
        rc = self.lib.maple_blen_batch(self.ctx, n, _dp(pIdx), _dp(cIdx), _dp(fromTipC), _dp(scratch), _dp(ss), _dp(out), _dp(st),
                                       self._stream())
        capi.check(self.ctx, rc, "maple_blen_batch")
        return out, st

    def vectors_differ_batch(self, idx1, idx2) -> torch.Tensor:
        idx1, idx2 = self._t(idx1, torch.int32), self._t(idx2, torch.int32)
        n = idx1.numel()
        out = torch.empty(n, dtype=torch.uint8, device=self.device)
        rc = self.lib.maple_vectors_differ_batch(self.ctx, n, _dp(idx1), _dp(idx2), _dp(out), self._stream())
        capi.check(self.ctx, rc, "maple_vectors_differ_batch")
        return out

    def prob_root_batch(self, idx) -> torch.Tensor:
        self._ensure_root_tables()
        idx = self._t(idx, torch.int32)
        out = torch.empty(idx.numel(), dtype=torch.float64, device=self.device)
        rc = self.lib.maple_prob_root_batch(self.ctx, idx.numel(), _dp(idx), _dp(out), self._stream())
        capi.check(self.ctx, rc, "maple_prob_root_batch")
        return out

    def root_vector_batch(self, idx, bLen, isFromTip, shorten: bool = True) -> MergeResult:
        """rootVector(probVect, bLen, isFromTip) (:4916) for n lists of the bound arena; shorten=True returns what the reference
        returns (it shortens before returning, :4994)."""
        L = self.lists
        idx, bLen, isFromTip = self._t(idx, torch.int32), self._t(bLen, torch.float64), self._t(isFromTip, torch.uint8)
        n = idx.numel()
        cap = (L.nkeys[idx.long()].long() + 3) // 4 * 4 + 4
        ks = torch.cumsum(cap, 0) - cap
        ps = ks * 6
        total = int(cap.sum().item())
        out_key = torch.empty(total + 4, dtype=torch.int32, device=self.device)
        out_pay = torch.empty(total * 6 + 4, dtype=torch.float64, device=self.device)
        nk = torch.empty(n, dtype=torch.int32, device=self.device)
        npay = torch.empty(n, dtype=torch.int32, device=self.device)
        rc = self.lib.maple_root_vector_batch(self.ctx, n, _dp(idx), _dp(bLen), _dp(isFromTip), _dp(out_key), _dp(out_pay), _dp(ks), _dp(ps),
                                              _dp(nk), _dp(npay), 1 if shorten else 0, self._stream())
        capi.check(self.ctx, rc, "maple_root_vector_batch")
        st = torch.zeros(n, dtype=torch.int32, device=self.device)
        return MergeResult(out_key, out_pay, ks, ps, nk, npay, None, st, L.lRef, L.U)

    def pass_branch_batch(self, idx, mutNode, dirIsUp, mutStart: torch.Tensor, mut: torch.Tensor) -> MergeResult:
        """passGenomeListThroughBranch for n lists; mutStart/mut: CSR mutation lists on the device (int32)."""
        L = self.lists
        idx, mutNode, dirIsUp = self._t(idx, torch.int32), self._t(mutNode, torch.int32), self._t(dirIsUp, torch.uint8)
        n = idx.numel()
        nm = (mutStart[mutNode.long() + 1] - mutStart[mutNode.long()]).long()
        cap = (L.nkeys[idx.long()].long() + 2 * nm + 2 + 3) // 4 * 4
        ks = torch.cumsum(cap, 0) - cap
        ps = ks * 6
        total = int(cap.sum().item())
        out_key = torch.empty(total + 4, dtype=torch.int32, device=self.device)
        out_pay = torch.empty(total * 6 + 4, dtype=torch.float64, device=self.device)
        nk = torch.empty(n, dtype=torch.int32, device=self.device)
        npay = torch.empty(n, dtype=torch.int32, device=self.device)
        rc = self.lib.maple_pass_branch_batch(self.ctx, n, _dp(idx), _dp(mutNode), _dp(dirIsUp), _dp(mutStart), _dp(mut), _dp(out_key),
                                              _dp(out_pay), _dp(ks), _dp(ps), _dp(nk), _dp(npay), self._stream())
        capi.check(self.ctx, rc, "maple_pass_branch_batch")
        st = torch.zeros(n, dtype=torch.int32, device=self.device)
        return MergeResult(out_key, out_pay, ks, ps, nk, npay, None, st, L.lRef, L.U)

    # ---------------------------------------------------------------- reference-named single calls
    def _bind_pair(self, a, b):
        keep = self.lists
        self.bind(pack_lists([a, b], self.model.lRef, self.model.usingErrorRate))
        return keep

    def _restore(self, keep):
        if keep is not None:
            self.bind(keep)

    def appendProbNode(self, probVectP, probVectC, isTipC, bLen) -> float:
        keep = self._bind_pair(probVectP, probVectC)
        try:
            return float(self.append_prob_batch([0], [1], [1 if isTipC else 0], [float(bLen)]).cpu()[0])
        finally:
            self._restore(keep)

    def mergeVectors(self, probVect1, bLen1, fromTip1, probVect2, bLen2, fromTip2, returnLK=False, isUpDown=False, numMinor1=0,
                     numMinor2=0):
        keep = self._bind_pair(probVect1, probVect2)
        try:
            fl = (capi.MAPLE_MERGE_UPDOWN if isUpDown else 0) | (capi.MAPLE_MERGE_RETURN_LK if returnLK else 0)
            r = self.merge_batch([0], [float(bLen1)], [1 if fromTip1 else 0], [1], [float(bLen2)], [1 if fromTip2 else 0], [fl],
                                 [int(numMinor1)], [int(numMinor2)])
            st = int(r.status.cpu()[0])
            if st == 2:
                raise capi.MapleError("mergeVectors: likelihood underflow (the reference raises Exception('exit'))")
            out = r.to_lists()[0]
            if returnLK:
                return (out, float(r.lk.cpu()[0])) if out is not None else (None, None)
            return out
        finally:
            self._restore(keep)

    def estimateBranchLengthWithDerivative(self, probVectP, probVectC, fromTipC=False):
        keep = self._bind_pair(probVectP, probVectC)
        try:
            out, st = self.blen_batch([0], [1], [1 if fromTipC else 0])
            return False if int(st.cpu()[0]) == 1 else float(out.cpu()[0])
        finally:
            self._restore(keep)

    def findProbRoot(self, probVect) -> float:
        keep = self._bind_pair(probVect, probVect)
        try:
            return float(self.prob_root_batch([0]).cpu()[0])
        finally:
            self._restore(keep)

    def rootVector(self, probVect, bLen, isFromTip):
        keep = self._bind_pair(probVect, probVect)
        try:
            return self.root_vector_batch([0], [float(bLen) if bLen else 0.0], [1 if isFromTip else 0]).to_lists()[0]
        finally:
            self._restore(keep)

    def passGenomeListThroughBranch(self, probVect, mutations, dirIsUp=False):
        keep = self._bind_pair(probVect, probVect)
        try:
            ms = torch.tensor([0, len(mutations)], dtype=torch.int32, device=self.device)
            mu = torch.tensor([list(m) for m in mutations] or [[0, 0, 0]], dtype=torch.int32, device=self.device).reshape(-1)
            return self.pass_branch_batch([0], [0], [1 if dirIsUp else 0], ms, mu).to_lists()[0]
        finally:
            self._restore(keep)

    def areVectorsDifferent(self, probVect1, probVect2) -> bool:
        if probVect2 is None:
            return True
        keep = self._bind_pair(probVect1, probVect2)
        try:
            return bool(self.vectors_differ_batch([0], [1]).cpu()[0])
        finally:
            self._restore(keep)
