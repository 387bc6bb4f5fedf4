"""Substitution / error model state the likelihood kernels consume.

Mirrors the module globals of the reference that its hot-path functions read
(MAPLEv0.7.5.4.py:3606-3693 reference-derived constants, :4055-4069 and :6350-6390 rate tables)
and that ``startTopologyUpdatesParallel`` ships to its workers (:9581).  Everything derived
here (cumulative tables, totError) is recomputed with the reference's own operation order so
that the values are bit-identical to what the reference holds.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

_ALLELES = {"a": 0, "c": 1, "g": 2, "t": 3, "A": 0, "C": 1, "G": 2, "T": 3}


def ref_indices(ref: str) -> np.ndarray:
    """refIndeces (:3680-3685): non-ACGT reference characters count as A."""
    return np.fromiter((_ALLELES.get(ch, 0) for ch in ref), dtype=np.int8, count=len(ref))


def ref_valid(ref: str) -> np.ndarray:
    """True where the reference holds A/C/G/T (the positions cumulativeBases counts, :3640-3647)."""
    return np.fromiter((ch in _ALLELES for ch in ref), dtype=bool, count=len(ref))


def root_freqs_from_ref(ref: str):
    """rootFreqs (:3670-3678): base composition over ACGT characters, divided by lRef."""
    counts = [0, 0, 0, 0]
    for ch in ref:
        i = _ALLELES.get(ch)
        if i is not None:
            counts[i] += 1
    return [c / float(len(ref)) for c in counts]


@dataclass
class MapleModel:
    lRef: int
    refIdx: np.ndarray  # int8 [lRef]
    rootFreqs: np.ndarray  # float64 [4]
    Q: np.ndarray  # float64 [4,4] mutMatrixGlobal
    usingErrorRate: bool = False
    errorRateSiteSpecific: bool = False
    useRateVariation: bool = False
    errorRate: float = 0.0  # errorRateGlobal
    siteRates: Optional[np.ndarray] = None  # [lRef] when useRateVariation
    errorRates: Optional[np.ndarray] = None  # [lRef] when errorRateSiteSpecific
    # thresholds (argparse defaults of the reference, :51, :61-62, :65)
    thresholdProb: float = 1e-8
    thresholdDiffForUpdate: float = 1e-5
    thresholdFoldChangeUpdate: float = 1.01
    minBLenSensitivityMutations: float = 0.001
    # derived
    cumulativeRate: np.ndarray = field(default=None, repr=False)
    cumulativeErrorRate: Optional[np.ndarray] = field(default=None, repr=False)
    totError: float = 0.0
    minBLenSensitivity: float = 0.0
    # reference positions holding A/C/G/T; ambiguous or N reference characters are indexed as A by refIndeces (:3680-3685)
    # but are NOT counted in cumulativeBases (:3640-3647).  None = every position is valid.
    refValid: Optional[np.ndarray] = field(default=None, repr=False)

    def __post_init__(self):
        self.refIdx = np.ascontiguousarray(self.refIdx, dtype=np.int8)
        self.rootFreqs = np.ascontiguousarray(self.rootFreqs, dtype=np.float64)
        self.Q = np.ascontiguousarray(self.Q, dtype=np.float64).reshape(4, 4)
        if self.siteRates is not None:
            self.siteRates = np.ascontiguousarray(self.siteRates, dtype=np.float64)
        if self.errorRates is not None:
            self.errorRates = np.ascontiguousarray(self.errorRates, dtype=np.float64)
        if self.useRateVariation and self.siteRates is None:
            raise ValueError("useRateVariation needs siteRates")
        if self.usingErrorRate and self.errorRateSiteSpecific and self.errorRates is None:
            raise ValueError("site-specific error model needs errorRates")
        self.minBLenSensitivity = self.minBLenSensitivityMutations * (1.0 / self.lRef)  # :3617-3618
        self.refresh()

    def refresh(self):
        """Recompute the cumulative tables after Q / siteRates / errorRates changed (:6350-6390)."""
        non_mut = np.array([self.Q[i, i] for i in range(4)])
        terms = non_mut[self.refIdx.astype(np.int64)]
        if self.useRateVariation:
            terms = terms * self.siteRates  # nonMutRates[refIndeces[i]]*siteRates[i]
        cr = np.empty(self.lRef + 1, dtype=np.float64)
        cr[0] = 0.0
        np.cumsum(terms, out=cr[1:])  # sequential accumulate == the reference's running sum
        self.cumulativeRate = cr
        if self.usingErrorRate and self.errorRateSiteSpecific:
            ce = np.empty(self.lRef + 1, dtype=np.float64)
            ce[0] = 0.0
            np.cumsum(self.errorRates, out=ce[1:])
            self.cumulativeErrorRate = ce
            self.totError = -float(ce[-1])
        else:
            self.cumulativeErrorRate = None
            self.totError = -self.errorRate * self.lRef if self.usingErrorRate else 0.0

    # tables only findProbRoot needs (tree log-likelihood, :4865-4912)
    def cumulative_bases(self) -> np.ndarray:
        cb = np.zeros((self.lRef + 1, 4), dtype=np.int32)
        onehot = np.zeros((self.lRef, 4), dtype=np.int32)
        onehot[np.arange(self.lRef), self.refIdx.astype(np.int64)] = 1
        if self.refValid is not None:
            onehot[~np.asarray(self.refValid, bool)] = 0
        np.cumsum(onehot, axis=0, out=cb[1:])
        return cb

    def root_freqs_log_error_cumulative(self) -> np.ndarray:
        out = np.zeros(self.lRef + 1, dtype=np.float64)
        pi = [float(x) for x in self.rootFreqs]
        acc = 0.0
        ridx = self.refIdx
        for i in range(self.lRef):
            e = float(self.errorRates[i]) if (self.errorRateSiteSpecific and self.errorRates is not None) else self.errorRate
            acc = acc + math.log(pi[ridx[i]] * (1.0 - 1.33333 * e) + 0.333333 * e)  # :6384 / :6389
            out[i + 1] = acc
        return out

    @classmethod
    def from_ref(cls, ref: str, Q, **kw) -> "MapleModel":
        """Model for a reference genome string: refIndeces, rootFreqs and the validity mask are derived from it
        (:3640-3685), everything else is passed through."""
        return cls(lRef=len(ref), refIdx=ref_indices(ref), rootFreqs=np.array(root_freqs_from_ref(ref), dtype=np.float64),
                   Q=np.array(Q, dtype=np.float64), refValid=ref_valid(ref), **kw)

    @classmethod
    def from_reference_snapshot(cls, env: dict, model: dict) -> "MapleModel":
        """Build from a golden fixture (tests/golden/make_golden.py: 'env' + 'model')."""
        ref = env["ref"]
        m = cls(
            lRef=env["lRef"],
            refIdx=ref_indices(ref),
            rootFreqs=np.array(env["rootFreqs"], dtype=np.float64),
            Q=np.array(model["mutMatrixGlobal"], dtype=np.float64),
            usingErrorRate=bool(env["usingErrorRate"]),
            errorRateSiteSpecific=bool(env["errorRateSiteSpecific"]),
            useRateVariation=bool(env["useRateVariation"]),
            errorRate=float(model["errorRateGlobal"]),
            siteRates=None if model["siteRates"] is None else np.array(model["siteRates"], dtype=np.float64),
            errorRates=None if model["errorRates"] is None else np.array(model["errorRates"], dtype=np.float64),
            thresholdProb=env["thresholdProb"],
            thresholdDiffForUpdate=env["thresholdDiffForUpdate"],
            thresholdFoldChangeUpdate=env["thresholdFoldChangeUpdate"],
        )
        m.refValid = ref_valid(ref)
        # the fixture records minBLenSensitivity already scaled by 1/lRef
        m.minBLenSensitivity = float(env["minBLenSensitivity"])
        return m
