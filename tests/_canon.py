"""Slot-permutation-invariant canonical form of a DCGrid block pool.

Blocks are ordered by (level, x, y, z); slot ids inside parent / child / apron entries are
rewritten to ranks in that order, and field bricks are permuted accordingly.  Two pools
that differ only in which pool slot each block landed in have equal canonical forms."""
import numpy as np

NONE = np.uint64(0xFFFFFFFFFFFFFFFF)


def canonical(topo, fields=None):
    pos = np.asarray(topo["pos"]).reshape(-1, 3).astype(np.int64)
    lvl = np.asarray(topo["level"]).astype(np.int64)
    parent = np.asarray(topo["parent"]).astype(np.uint64)
    child = np.asarray(topo["child"]).reshape(-1, 8).astype(np.uint64)
    M = lvl.shape[0]
    active = np.nonzero(lvl != 0xFF)[0]
    order = active[np.lexsort((pos[active, 2], pos[active, 1], pos[active, 0], lvl[active]))]
    rank = np.full(M + 1, -1, dtype=np.int64)  # rank[M] = sentinel for "none"
    rank[order] = np.arange(order.size)
    out = {}
    out["blocks"] = np.concatenate([lvl[order, None], pos[order]], axis=1)

    def remap_sub(idx):  # 8*slot+sub -> 8*rank+sub, NONE -> -1
        idx = np.asarray(idx, dtype=np.uint64)
        none = idx == NONE
        slot = np.where(none, M, (idx // np.uint64(8)).astype(np.int64))
        slot = np.minimum(slot, M)
        r = rank[slot]
        return np.where(none | (r < 0), -1, r * 8 + (idx % np.uint64(8)).astype(np.int64))

    def remap_slot(idx):
        idx = np.asarray(idx, dtype=np.uint64)
        none = idx == NONE
        slot = np.minimum(np.where(none, M, idx.astype(np.int64)), M)
        r = rank[slot]
        return np.where(none, -1, np.where(r < 0, -2, r))

    out["parent"] = remap_sub(parent[order])
    out["child"] = remap_slot(child[order])
    if topo.get("apron") is not None:
        apron = np.asarray(topo["apron"]).reshape(-1, 216).astype(np.uint64)[order]
        slot = np.minimum((apron // np.uint64(64)).astype(np.int64), M)
        r = rank[slot]
        out["apron"] = np.where(r < 0, -2, r * 64 + (apron % np.uint64(64)).astype(np.int64))
    if fields:
        for name, arr in fields.items():
            a = np.asarray(arr)
            comps = a.size // (M * 64)
            out[name] = a.reshape(M, 64, comps)[order]
    return out


def diff_report(a, b):
    """Returns a dict name -> number of mismatching entries (0 everywhere == canonical equality)."""
    rep = {}
    for k in a:
        if k not in b:
            continue
        x, y = np.asarray(a[k]), np.asarray(b[k])
        if x.shape != y.shape:
            rep[k] = f"shape {x.shape} vs {y.shape}"
            continue
        if x.dtype.kind == "f":
            rep[k] = int(np.count_nonzero(x.view(np.uint32) != y.view(np.uint32)))
        else:
            rep[k] = int(np.count_nonzero(x != y))
    return rep
