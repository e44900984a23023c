"""Executable specification of the *parallel* formulation of stage 1 used by the CUDA kernels
(DESIGN.md "id model").  Pure Python, small images only.  TEST INFRASTRUCTURE:
it is fuzzed against the oracle on CPU (tests/test_model_vs_oracle.py) so that the decomposition the kernels
implement (static edges -> provisional clusters -> events -> reconstruction tree -> attributed counts ->
winners -> death flags -> prefix-scan numbering -> listing order) is known to reproduce the sequential
reference before any GPU time is spent.

Each step is written as a whole-array pass (what one kernel does), never as a raster simulation.
"""
from __future__ import annotations

import numpy as np

NEW, UP, LEFT, UPLEFT, NONE = 0, 1, 2, 3, 4


class Unsupported(Exception):
    pass


def _same(a, b, shift, thres):
    return all(abs((int(a[k]) >> shift) - (int(b[k]) >> shift)) <= thres for k in range(3))


def static_edges_colour(px, diagonal, shift, thres, key, has_key):
    """step 1: per-pixel join code J and merge flag M, pure functions of pixel values."""
    h, w = px.shape[:2]
    n = w * h
    J = np.zeros(n, dtype=np.int8)
    M = np.zeros(n, dtype=bool)
    keyed = np.zeros(n, dtype=bool)
    if has_key:
        keyed = (px.reshape(n, 4) == np.array(key, dtype=np.uint8)).all(axis=1)
        for i in range(n):
            if not keyed[i] and _same(px.reshape(n, 4)[i], key, shift, thres):
                raise Unsupported("unclean key colour")
    for y in range(h):
        for x in range(w):
            i = y * w + x
            c = px[y, x]
            up = px[y - 1, x] if y > 0 else None
            left = px[y, x - 1] if x > 0 else None
            ul = px[y - 1, x - 1] if (x > 0 and y > 0) else None
            s = lambda a, b: a is not None and b is not None and _same(a, b, shift, thres)
            if x > 0 and y > 0 and not keyed[i - 1] and not keyed[i - w]:
                M[i] = s(left, up) and (diagonal or (s(c, left) and s(c, up)))
            if keyed[i]:
                J[i] = NONE
            elif s(c, up) and s(c, ul):
                J[i] = UP
            elif s(c, left) and s(c, ul):
                J[i] = LEFT
            elif diagonal and s(c, ul):
                J[i] = UPLEFT
            else:
                J[i] = NEW
    return J, M, keyed


def static_edges_binary(mask, diagonal):
    h, w = mask.shape
    n = w * h
    J = np.full(n, NONE, dtype=np.int8)
    M = np.zeros(n, dtype=bool)
    g = lambda x, y: bool(mask[y, x]) if (0 <= x < w and 0 <= y < h) else False
    for y in range(h):
        for x in range(w):
            i = y * w + x
            v, vu, vl, vul = g(x, y), g(x, y - 1), g(x - 1, y), g(x - 1, y - 1)
            M[i] = (v or diagonal) and vu and vl
            if v:
                J[i] = UP if vu else LEFT if vl else UPLEFT if (vul and diagonal) else NEW
    return J, M


def run_model(J, M, w, h, base, check_stale_upleft):
    """steps 2-9.  Returns dict(labels, num_slots, listing per slot, areas per slot)."""
    n = w * h
    off = {UP: -w, LEFT: -1, UPLEFT: -w - 1}
    # step 2: provisional clusters = trees of the J forest (parent index < own index -> one forward pass == pointer jumping)
    pc = np.full(n, -1, dtype=np.int64)
    for i in range(n):
        if J[i] == NEW:
            pc[i] = i
        elif J[i] != NONE:
            pc[i] = pc[i + off[int(J[i])]]
    new_pix = [i for i in range(n) if J[i] == NEW]
    K = len(new_pix)
    leaf_of = {p: k for k, p in enumerate(new_pix)}           # compaction of NEW flags (prefix scan)
    # step 4: candidate events, time-sorted by construction
    events = [(i, leaf_of[pc[i - 1]], leaf_of[pc[i - w]]) for i in range(n)
              if M[i] and pc[i - 1] != pc[i - w]]
    # step 5: reconstruction tree (structure only; union-find over leaves)
    uf = list(range(K))

    def find(a):
        while uf[a] != a:
            uf[a] = uf[uf[a]]
            a = uf[a]
        return a

    cur = list(range(K))                 # current tree node of each set (indexed by uf root)
    parent, birth, child_l, child_u = [-1] * K, [-1] * K, [-1] * K, [-1] * K
    stale = []                           # (node, side) candidates for the stale-upleft quirk
    for (i, a, b) in events:
        ra, rb = find(a), find(b)
        if ra == rb:
            continue
        v = len(parent)
        parent.append(-1); birth.append(i); child_l.append(cur[ra]); child_u.append(cur[rb])
        parent[cur[ra]] = v
        parent[cur[rb]] = v
        if check_stale_upleft and J[i] == UPLEFT:
            ru = find(leaf_of[pc[i - w - 1]])
            if ru == ra:
                stale.append((v, 0))
            elif ru == rb:
                stale.append((v, 1))
        uf[ra] = rb
        cur[rb] = v
    nn = len(parent)
    # step 6: attribute every pixel to the highest ancestor of its leaf born no later than the pixel
    cnt = [0] * nn
    attr = np.full(n, -1, dtype=np.int64)
    for p in range(n):
        if J[p] == NONE:
            continue
        v = leaf_of[pc[p]]
        while parent[v] != -1 and birth[parent[v]] <= p:
            v = parent[v]
        attr[p] = v
        cnt[v] += 1
    # W = subtree sums (node ids increase with time, children before parents)
    W = list(cnt)
    for v in range(K, nn):
        W[v] += W[child_l[v]] + W[child_u[v]]
    # step 7: winners / carriers / deaths
    carrier = list(range(K)) + [-1] * (nn - K)
    left_wins = [False] * nn
    death_time = [None] * K
    death_left = [False] * K
    for v in range(K, nn):
        l, u = child_l[v], child_u[v]
        if W[l] <= W[u]:                       # left loses, up's label survives
            carrier[v] = carrier[u]
            death_time[carrier[l]] = birth[v]; death_left[carrier[l]] = True
        else:
            left_wins[v] = True
            carrier[v] = carrier[l]
            death_time[carrier[u]] = birth[v]; death_left[carrier[u]] = False
    for (v, side) in stale:
        loser_side = 1 if left_wins[v] else 0
        if side == loser_side:
            raise Unsupported("stale upleft label resurrects a dead slot")
    # step 8: D flags and prefix-scan numbering
    D = [False] * K
    for k in range(K):
        nxt = new_pix[k + 1] if k + 1 < K else None
        if death_time[k] is not None and death_left[k] and (nxt is None or death_time[k] <= nxt):
            D[k] = True
    label = [0] * K
    run = base
    for k in range(K):
        label[k] = run
        if not D[k]:
            run += 1
    num_slots = run + (1 if (K > 0 and D[K - 1]) else 0)
    # root / final labels
    root = list(range(nn))
    for v in range(nn - 1, -1, -1):
        root[v] = root[parent[v]] if parent[v] != -1 else v
    labels = np.zeros(n, dtype=np.int64)
    for p in range(n):
        if J[p] != NONE:
            labels[p] = label[carrier[root[leaf_of[pc[p]]]]]
        else:
            labels[p] = -1
    # step 9: listing order: listing(v) = listing(winner) ++ listing(loser) ++ own(v)
    start = [0] * nn                     # offset of the node's block inside its cluster listing
    own_start = [0] * nn
    for v in range(nn - 1, -1, -1):
        if v >= K:
            l, u = child_l[v], child_u[v]
            win, lose = (l, u) if left_wins[v] else (u, l)
            start[win] = start[v]
            start[lose] = start[v] + W[win]
            own_start[v] = start[v] + W[win] + W[lose]
        else:
            own_start[v] = start[v]
    listing = {}
    order = sorted((p for p in range(n) if J[p] != NONE), key=lambda p: (labels[p], own_start[attr[p]], p))
    for p in order:
        listing.setdefault(int(labels[p]), []).append(p)
    return dict(labels=labels, num_slots=max(num_slots, base), listing=listing, K=K)
