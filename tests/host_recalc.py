"""Test helper: reCalculateAllGenomeLists (MAPLEv0.7.5.4.py:6013-6347) restated over the CPU oracle's
primitives, including the MAT re-referencing at nodes that carry mutation lists."""


def recalc_lists(orc, up, children, dist, mutations, isTip, root, lower):
    """lower: {tip node: genome list}.  Returns dicts lower, upRight, upLeft, totUp (node -> list / None)."""
    lower = dict(lower)
    upR, upL, tot = {}, {}, {}
    order, stack = [], [root]
    while stack:
        n = stack.pop()
        order.append(n)
        stack.extend(children[n])

    def pass_up(c, v):
        return orc.pass_branch(v, mutations[c], True) if mutations[c] else v

    def pass_down(c, v):
        return orc.pass_branch(v, mutations[c], False) if mutations[c] else v

    for n in reversed(order):
        ch = children[n]
        if ch:
            v = orc.merge(pass_up(ch[0], lower[ch[0]]), dist[ch[0]], isTip[ch[0]], pass_up(ch[1], lower[ch[1]]), dist[ch[1]], isTip[ch[1]])
            assert v is not None
            lower[n] = orc.shorten(v)

    def root_vector(v, bLen, tip):
        # rootVector re-expresses relative to the reference genome and back (:4928-4940, :4990-4993)
        v = pass_up(root, v)
        r = orc.root_vector(v, bLen, tip)
        r = pass_down(root, r)
        return orc.shorten(r)

    ch = children[root]
    if ch:
        upR[root] = root_vector(pass_up(ch[1], lower[ch[1]]), dist[ch[1]], isTip[ch[1]])
        upL[root] = root_vector(pass_up(ch[0], lower[ch[0]]), dist[ch[0]], isTip[ch[0]])
    for n in order[1:]:
        p = up[n]
        vectUp = upR[p] if children[p][0] == n else upL[p]
        vectUp = pass_down(n, vectUp)
        if dist[n]:
            tot[n] = orc.shorten(orc.merge(vectUp, dist[n] / 2, False, lower[n], dist[n] / 2, isTip[n], isUpDown=True))
        else:
            tot[n] = None
        ch = children[n]
        if ch:
            v1 = pass_up(ch[1], lower[ch[1]])
            v0 = pass_up(ch[0], lower[ch[0]])
            r = orc.merge(vectUp, dist[n], False, v1, dist[ch[1]], isTip[ch[1]], isUpDown=True)
            upR[n] = None if r is None else orc.shorten(r)
            r = orc.merge(vectUp, dist[n], False, v0, dist[ch[0]], isTip[ch[0]], isUpDown=True)
            upL[n] = None if r is None else orc.shorten(r)
    return lower, upR, upL, tot
