"""A reproducible random chain of list operations, run by any backend that offers the reference's list functions (the reference
itself when fixtures are generated, the C oracle and the host-compiled CUDA source when they are tested).  The random stream
decides operations, operands and branch lengths; results feed back into the operand pools, so the chain wanders into lists the
recorded runs never produced (zero and huge branch lengths, repeated merges of merged lists, tips flags on internal lists ...).
Every step yields a compact result: a digest of the output list (bit-exact content) or the scalar itself."""
import hashlib
import json
import random


def canon(gl):
    if gl is None:
        return None
    return [[[(float(v) + 0.0).hex() for v in x] if isinstance(x, (list, tuple)) else (float(x) + 0.0).hex() for x in e] for e in gl]


def digest(gl):
    return None if gl is None else hashlib.md5(json.dumps(canon(gl)).encode()).hexdigest()[:16]


def initial_pools(tree, lists_of):
    """tree: the frozen tree of a fixture (or the reference's Tree); lists_of(family, node) -> list or None.  Only nodes without MAT
    mutations: their lists are in the reference genome's coordinates."""
    n = len(tree["up"])
    clean = [i for i in range(n) if not tree["mutations"][i]]
    lower = [lists_of("probVect", i) for i in clean]
    lower = [v for v in lower if v is not None][:120]
    upper = [lists_of(f, i) for f in ("probVectUpRight", "probVectUpLeft", "probVectTotUp") for i in clean]
    upper = [v for v in upper if v is not None][:200]
    return lower, upper


def run_chain(be, lower, upper, lRef, seed, steps):
    """be: backend with merge(a,b1,t1,b,b2,t2,returnLK,isUpDown,numMinor1,numMinor2) -> list | None | (list, lk); append; blen
    (None for False); differ; root_vector (shortened, as the reference returns it); prob_root; shorten (returns a new list)."""
    rng = random.Random(seed)
    lower, upper = list(lower), list(upper)
    out = []

    def blen():
        return rng.choice([0.0, 0.0, 1e-9, rng.random() / lRef, 3 * rng.random() / lRef, 0.1 * rng.random(), 0.1])

    for _ in range(steps):
        op = rng.randrange(8)
        if op == 0:
            a, b = rng.choice(lower), rng.choice(lower)
            b1, t1, b2, t2 = blen(), rng.random() < 0.3, blen(), rng.random() < 0.3
            lk, nm1, nm2 = rng.random() < 0.3, rng.choice([0, 0, 2]), rng.choice([0, 0, 1])
            r = be.merge(a, b1, t1, b, b2, t2, lk, False, nm1, nm2)
            v = r[0] if (lk and r is not None) else r
            out.append([0, digest(v), r[1] if (lk and r is not None) else None])
            if v is not None and len(lower) < 400:
                lower.append(be.shorten(v))
        elif op == 1:
            a, b = rng.choice(upper), rng.choice(lower)
            b1, b2, t2 = blen(), blen(), rng.random() < 0.3
            v = be.merge(a, b1, False, b, b2, t2, False, True, 0, 0)
            out.append([1, digest(v)])
            if v is not None and len(upper) < 500:
                upper.append(be.shorten(v))
        elif op == 2:
            a, b, tip, bl = rng.choice(upper), rng.choice(lower), rng.random() < 0.5, blen()
            r = be.append(a, b, tip, bl)
            out.append([2, "-inf" if r == float("-inf") else r])
        elif op == 3:
            a, b, tip = rng.choice(upper), rng.choice(lower), rng.random() < 0.5
            out.append([3, be.blen(a, b, tip)])
        elif op == 4:
            pool = rng.choice([lower, upper])
            a, b = rng.choice(pool), rng.choice(pool)
            out.append([4, bool(be.differ(a, b))])
        elif op == 5:
            a, bl, tip = rng.choice(lower), blen(), rng.random() < 0.5
            out.append([5, digest(be.root_vector(a, bl, tip))])
        elif op == 6:
            out.append([6, be.prob_root(rng.choice(lower))])
        else:
            out.append([7, digest(be.shorten(rng.choice(rng.choice([lower, upper]))))])
    return out


class OracleBackend:
    """The C oracle or the host-compiled CUDA source (same python interface)."""

    def __init__(self, orc):
        self.o = orc

    def merge(self, a, b1, t1, b, b2, t2, lk, updown, nm1, nm2):
        return self.o.merge(a, b1, t1, b, b2, t2, returnLK=lk, isUpDown=updown, numMinor1=nm1, numMinor2=nm2)

    def append(self, a, b, tip, bl):
        return self.o.append(a, b, tip, bl)

    def blen(self, a, b, tip):
        return self.o.blen(a, b, tip)

    def differ(self, a, b):
        return self.o.differ(a, b)

    def root_vector(self, a, bl, tip):
        return self.o.shorten(self.o.root_vector(a, bl, tip))

    def prob_root(self, a):
        return self.o.prob_root(a)

    def shorten(self, a):
        return self.o.shorten(a)
