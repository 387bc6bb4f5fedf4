"""Packed MAPLE genome lists: the HBM layout the kernels (and the CPU oracle) read.

A MAPLE genome list (reference spec: MAPLEv0.7.5.4.py:378-390) is a python list of tuples
covering genome positions 1..lRef.  The packed form splits every list into two streams:

  key  : uint32 per entry
           bits 0-2   type      0-3 = A,C,G,T   4 = R   5 = N   6 = O
           bits 3-4   nLens     how many branch-length fields the tuple carries (0, 1 or 2)
           bit  5     flag      "observation comes from a tip without minor sequences"
                                (only meaningful under the error model, reference :383-384)
           bits 6-7   nuc       local-reference nucleotide (entry[1] for types 0-3 and 6)
           bits 8-31  end       1-based inclusive last position covered by the entry.  The
                                reference keeps this implicit for single-site entries; it is
                                explicit here so that entries are self-describing.
  pay  : float64 payload, in entry order: the nLens lengths, then (type 6 only) the 4-vector.

Tuple length <-> (nLens, flag) under usingErrorRate U (reference tests such as
``len(entry)==3+usingErrorRate``, :6589-6599, :4687-4696):
  types 0-4:  U=0: len = 2 + nLens            U=1: len = 2 (nLens=0) | 4 (nLens=1) | 5 (nLens=2)
  type  5  :  len = 2
  type  6  :  len = 3 (nLens=0) | 4 (nLens=1)
Length-3 tuples of type<5 under U=1 only arise in reference branches that index out of
range (:4516, :4536-4538, :4631-4633) and are rejected here.

Lists of a collection sit back to back in one arena; list i starts at key[key_start[i]] and
pay[pay_start[i]], both aligned to 16 bytes.  key_start = -1 encodes python ``None``.
Genomes up to 2**24-1 positions are representable.
"""
from __future__ import annotations

import numpy as np

TYPE_R, TYPE_N, TYPE_O = 4, 5, 6
MAX_LREF = (1 << 24) - 1
KEY_ALIGN = 4  # entries (16 B)
PAY_ALIGN = 2  # doubles (16 B)


def make_key(typ: int, nlens: int, flag: int, nuc: int, end: int) -> int:
    return (typ & 7) | ((nlens & 3) << 3) | ((1 if flag else 0) << 5) | ((nuc & 3) << 6) | (end << 8)


def split_key(k: int):
    k = int(k)
    return k & 7, (k >> 3) & 3, (k >> 5) & 1, (k >> 6) & 3, k >> 8


class PackedLists:
    """A collection of genome lists in arena form (host numpy arrays)."""

    __slots__ = ("key", "pay", "key_start", "pay_start", "nkeys", "npay", "lRef", "U")

    def __init__(self, key, pay, key_start, pay_start, nkeys, npay, lRef, U):
        self.key, self.pay = key, pay
        self.key_start, self.pay_start = key_start, pay_start
        self.nkeys, self.npay = nkeys, npay
        self.lRef, self.U = int(lRef), int(bool(U))

    def __len__(self):
        return len(self.key_start)

    def get(self, i: int):
        return unpack_list(self, i)


def _encode_entry(e, pos: int, U: int, keys: list, pay: list) -> int:
    """Append one tuple to the streams; returns the new position (sites consumed)."""
    t = e[0]
    L = len(e)
    if t == TYPE_N:
        if L != 2:
            raise ValueError("N entry with %d fields" % L)
        keys.append(make_key(5, 0, 0, 0, e[1]))
        return e[1]
    if t == TYPE_O:
        if L == 3:
            nl = 0
        elif L == 4:
            nl = 1
            pay.append(float(e[2]))
        else:
            raise ValueError("O entry with %d fields" % L)
        v = e[-1]
        pay.extend((float(v[0]), float(v[1]), float(v[2]), float(v[3])))
        keys.append(make_key(6, nl, 0, e[1], pos + 1))
        return pos + 1
    if not 0 <= t <= 4:
        raise ValueError("bad entry type %r" % (t,))
    flag = 0
    if U:
        if L == 2:
            nl = 0
        elif L == 4:
            nl = 1
            pay.append(float(e[2]))
            flag = 1 if e[3] else 0
        elif L == 5:
            nl = 2
            pay.append(float(e[2]))
            pay.append(float(e[3]))
            flag = 1 if e[4] else 0
        else:
            raise ValueError("entry of type %d with %d fields under the error model" % (t, L))
    else:
        nl = L - 2
        if nl < 0 or nl > 2:
            raise ValueError("entry of type %d with %d fields" % (t, L))
        for x in e[2:]:
            pay.append(float(x))
    if t == TYPE_R:
        keys.append(make_key(4, nl, flag, 0, e[1]))
        return e[1]
    keys.append(make_key(t, nl, flag, e[1], pos + 1))
    return pos + 1


def pack_lists(lists, lRef: int, U) -> PackedLists:
    """Pack python genome lists (``None`` allowed) into one arena."""
    if lRef > MAX_LREF:
        raise ValueError("lRef %d exceeds the 24-bit position field" % lRef)
    U = int(bool(U))
    n = len(lists)
    key_start = np.full(n, -1, dtype=np.int64)
    pay_start = np.full(n, -1, dtype=np.int64)
    nkeys = np.zeros(n, dtype=np.int32)
    npay = np.zeros(n, dtype=np.int32)
    keys: list = []
    pay: list = []
    for i, gl in enumerate(lists):
        if gl is None:
            continue
        while len(keys) % KEY_ALIGN:
            keys.append(0)
        while len(pay) % PAY_ALIGN:
            pay.append(0.0)
        k0, p0 = len(keys), len(pay)
        pos = 0
        for e in gl:
            pos = _encode_entry(e, pos, U, keys, pay)
        if pos != lRef:
            raise ValueError("genome list %d covers %d of %d positions" % (i, pos, lRef))
        key_start[i], pay_start[i] = k0, p0
        nkeys[i], npay[i] = len(keys) - k0, len(pay) - p0
    # tail padding so that 16-byte vector loads never run off the arena
    while len(keys) % KEY_ALIGN or not keys:
        keys.append(0)
    while len(pay) % PAY_ALIGN or not pay:
        pay.append(0.0)
    return PackedLists(np.asarray(keys, dtype=np.uint32), np.asarray(pay, dtype=np.float64),
                       key_start, pay_start, nkeys, npay, lRef, U)


def decode_stream(key, pay, k0: int, p0: int, lRef: int, U: int, nkeys=None):
    """Decode one list from raw streams back into reference-style tuples."""
    if k0 < 0:
        return None
    out = []
    k, p, pos = int(k0), int(p0), 0
    while pos < lRef:
        if nkeys is not None and k - k0 >= nkeys:
            raise ValueError("list ends at position %d before lRef" % pos)
        t, nl, flag, nuc, end = split_key(key[k])
        k += 1
        lens = [float(pay[p + j]) for j in range(nl)]
        p += nl
        if t == TYPE_N:
            out.append((5, end))
        elif t == TYPE_O:
            vec = [float(pay[p]), float(pay[p + 1]), float(pay[p + 2]), float(pay[p + 3])]
            p += 4
            out.append((6, nuc) + tuple(lens) + (vec,))
        else:
            second = end if t == TYPE_R else nuc
            if U and nl:
                out.append((t, second) + tuple(lens) + (bool(flag),))
            else:
                out.append((t, second) + tuple(lens))
        if end <= pos:
            raise ValueError("non-increasing end position in packed list")
        pos = end
    return out


def unpack_list(pl: PackedLists, i: int):
    return decode_stream(pl.key, pl.pay, pl.key_start[i], pl.pay_start[i], pl.lRef, pl.U, int(pl.nkeys[i]))


def lists_equal(a, b) -> bool:
    """Exact structural + bitwise-value equality of two python genome lists."""
    if a is None or b is None:
        return a is None and b is None
    if len(a) != len(b):
        return False
    for x, y in zip(a, b):
        if len(x) != len(y):
            return False
        for u, v in zip(x, y):
            if isinstance(u, (list, tuple)):
                if list(u) != list(v):
                    return False
            elif u != v:
                return False
    return True
