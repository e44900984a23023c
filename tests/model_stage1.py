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


def run_model(J, M, w, h, base, check_stale_upleft, res=frozenset()):
    """steps 2-9.  Returns dict(labels, num_slots, listing per slot, areas per slot).

    `res` = pixels at which the reference's stale `cluster_upleft` read (builder.rs:301-305 vs :341-343) resurrects the
    slot that has just lost the merge at the same pixel: such a pixel starts a provisional cluster of its own (like a NEW
    pixel) that carries the dead label instead of a fresh one.  Whether a pixel is in that set depends on merge outcomes
    before it, so the set is the fixpoint of `stale -> res` (run_model_fixpoint); the model reports the set it observes in
    `stale_set`.  With check_stale_upleft and an inconsistent `res` the result is only meaningful as the next iterate."""
    n = w * h
    off = {UP: -w, LEFT: -1, UPLEFT: -w - 1}
    res = frozenset(int(p) for p in res if J[int(p)] == UPLEFT)
    # step 2: provisional clusters = trees of the J forest (parent index < own index -> one forward pass == pointer jumping)
    pc = np.full(n, -1, dtype=np.int64)
    for i in range(n):
        if J[i] == NEW or i in res:
            pc[i] = i
        elif J[i] != NONE:
            pc[i] = pc[i + off[int(J[i])]]
    new_pix = [i for i in range(n) if J[i] == NEW or i in res]   # all leaves, ascending; resurrections are not NEW events
    K = len(new_pix)
    is_new = [J[p] == NEW for p in new_pix]
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
    stale = []                           # (node, side, pixel) candidates for the stale-upleft quirk
    node_at = {}                         # event pixel -> tree node
    for (i, a, b) in events:
        ra, rb = find(a), find(b)
        if ra == rb:
            continue
        v = len(parent)
        node_at[i] = v
        parent.append(-1); birth.append(i); child_l.append(cur[ra]); child_u.append(cur[rb])
        parent[cur[ra]] = v
        parent[cur[rb]] = v
        if check_stale_upleft and J[i] == UPLEFT:
            ru = find(leaf_of[pc[i - w - 1]])
            if ru == ra:
                stale.append((v, 0, i))
            elif ru == rb:
                stale.append((v, 1, i))
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
    loser_car = [-1] * nn
    for v in range(K, nn):
        l, u = child_l[v], child_u[v]
        if W[l] <= W[u]:                       # left loses, up's label survives
            carrier[v] = carrier[u]
            loser_car[v] = carrier[l]
            death_time[carrier[l]] = birth[v]; death_left[carrier[l]] = True
        else:
            left_wins[v] = True
            carrier[v] = carrier[l]
            loser_car[v] = carrier[u]
            death_time[carrier[u]] = birth[v]; death_left[carrier[u]] = False
    stale_set = set()
    for (v, side, i) in stale:
        loser_side = 1 if left_wins[v] else 0
        if side == loser_side:
            stale_set.add(i)
    # a resurrected leaf carries the label of the leaf that lost at its own pixel; chains resolve to a NEW leaf (origin)
    src = [-1] * K
    for k in range(K):
        if not is_new[k]:
            v = node_at.get(new_pix[k])
            src[k] = loser_car[v] if v is not None else -1
    origin = list(range(K))
    for k in range(K):                         # sources are earlier leaves: one forward pass
        if not is_new[k]:
            origin[k] = origin[src[k]] if src[k] >= 0 else -1
    # step 8: D flags and prefix-scan numbering (NEW events only)
    news = [k for k in range(K) if is_new[k]]
    nxt_new = {}
    for j, k in enumerate(news):
        nxt_new[k] = new_pix[news[j + 1]] if j + 1 < len(news) else None
    D = [False] * K
    fence = False
    for k in range(K):
        o = origin[k]
        if o < 0 or death_time[k] is None or not death_left[k]:
            continue
        nx = nxt_new[o]
        if nx is None or death_time[k] <= nx:
            D[o] = True                            # builder.rs:315-317: the slot is handed to the next new cluster
            if death_time[k] in stale_set:
                fence = True                       # ... and resurrected at the same pixel: the next NEW overwrites it
    label = [0] * K
    run = base
    for k in news:
        label[k] = run
        if not D[k]:
            run += 1
    for k in range(K):
        if not is_new[k]:
            label[k] = label[origin[k]] if origin[k] >= 0 else 0
    num_slots = run + (1 if (news and D[news[-1]]) else 0)
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
    return dict(labels=labels, num_slots=max(num_slots, base), listing=listing, K=K, stale_set=frozenset(stale_set), fence=fence)


def run_model_fixpoint(J, M, w, h, base, max_iter=64):
    """diagonal=true colour path: iterate `res <- stale_set(res)` to the (unique, by causality in time) fixpoint."""
    res = frozenset()
    for it in range(max_iter):
        out = run_model(J, M, w, h, base, True, res)
        if out["stale_set"] == res:
            if out["fence"]:
                raise Unsupported("a slot freed by builder.rs:315-317 is resurrected at the same pixel")
            out["iterations"] = it + 1
            return out
        res = out["stale_set"]
    raise Unsupported("no fixpoint")
