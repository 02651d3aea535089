import sys, time, numpy as np
sys.path.insert(0, "/root/repo/oracle")
import hv_oracle, qhull_oracle
d, n = int(sys.argv[1]), int(sys.argv[2])
xs = np.random.default_rng(0).random((n, d))
base, normal = qhull_oracle.cuboid(d)
o = hv_oracle.run(xs, base, normal, nthreads=8)
sig = o["sig"]; V = len(sig)
print("V", V, flush=True)
# edges
D1 = d + 1
keys = []
for k in range(D1):
    e = np.delete(sig, k, axis=1)
    keys.append(e)
E = np.concatenate(keys)                      # (V*D1, d)
vid = np.tile(np.arange(V), D1); kid = np.repeat(np.arange(D1), V)
real = (E <= n).any(axis=1)
order = np.lexsort(E.T[::-1])
Es = E[order]
same = np.r_[False, (Es[1:] == Es[:-1]).all(axis=1)]
eid_sorted = np.cumsum(~same) - 1
eid = np.empty(len(E), dtype=np.int64); eid[order] = eid_sorted
NE = eid_sorted[-1] + 1
edge_of = np.full((V, D1), -1, dtype=np.int64); edge_of[vid, kid] = np.where(real, eid, -1)
# neighbour vertex across each edge
cnt = np.bincount(eid, minlength=NE)
first = np.full(NE, -1, dtype=np.int64); second = np.full(NE, -1, dtype=np.int64)
o_v = vid[order]
starts = np.r_[0, np.nonzero(~same)[0][1:]] if False else np.nonzero(~same)[0]
first[eid_sorted[starts]] = o_v[starts]
has2 = cnt >= 2
second[eid_sorted[starts[has2[eid_sorted[starts]]] ]] = o_v[starts[has2[eid_sorted[starts]]] + 1]
nbr = np.full((V, D1), -1, dtype=np.int64)
for k in range(D1):
    e = edge_of[:, k]; ok = e >= 0
    a, b = first[e[ok]], second[e[ok]]
    me = np.arange(V)[ok]
    nbr[ok, k] = np.where(a == me, b, a)
print("edges", NE, "max cnt", cnt.max(), flush=True)

def simulate(W, mode, nseeds, rng):
    found = np.zeros(V, bool); state = np.zeros(NE, np.int8)   # 0 unregistered, 1 open, 2 closed
    q_v = []; q_k = []
    def commit(new_vs):
        # register edges of new vertices; returns open entries
        ov, ok_ = [], []
        for k in range(D1):
            e = edge_of[new_vs, k]; m = e >= 0
            ee = e[m]; vv = new_vs[m]
            # sequential semantics within batch: handle duplicates of the same edge in this commit
            uniq, idx, c = np.unique(ee, return_index=True, return_counts=True)
            st = state[uniq]
            # edges seen once now
            one = c == 1
            # previously unregistered & once -> open; previously open -> closed; twice in this commit -> closed
            opn = one & (st == 0)
            state[uniq[opn]] = 1
            ov.append(vv[idx[opn]]); ok_.append(np.full(opn.sum(), k))
            state[uniq[~opn]] = 2
        return np.concatenate(ov), np.concatenate(ok_)
    seeds = rng.choice(V, nseeds, replace=False)
    found[seeds] = True
    av, ak = commit(seeds)
    # edges between two seeds got closed; open ones enqueued -- but entries of edges closed later are skipped at pop
    qv, qk = list(av), list(ak)
    qv = np.array(qv); qk = np.array(qk)
    rays = dups = skips = 0
    head = 0
    pend_v, pend_k = qv, qk
    while len(pend_v):
        if mode == "fifo":
            bv, bk = pend_v[:W], pend_k[:W]; pend_v, pend_k = pend_v[W:], pend_k[W:]
        elif mode == "lifo":
            bv, bk = pend_v[-W:], pend_k[-W:]; pend_v, pend_k = pend_v[:-W], pend_k[:-W]
        else:
            idx = rng.permutation(len(pend_v)); sel = idx[:W]; rest = np.sort(idx[W:])
            bv, bk = pend_v[sel], pend_k[sel]; pend_v, pend_k = pend_v[rest], pend_k[rest]
        e = edge_of[bv, bk]
        live = state[e] == 1
        skips += (~live).sum()
        bv, bk = bv[live], bk[live]
        rays += len(bv)
        tgt = nbr[bv, bk]
        tgt = tgt[tgt >= 0]
        u, c = np.unique(tgt, return_counts=True)
        isnew = ~found[u]
        dups += (c - 1).sum() + (~isnew).sum()
        newv = u[isnew]
        found[newv] = True
        # the walked edges become closed by the registration of the target
        if len(newv):
            av, ak = commit(newv)
            pend_v = np.concatenate([pend_v, av]); pend_k = np.concatenate([pend_k, ak])
    return rays, dups, skips, found.sum()

rng = np.random.default_rng(1)
for W in (2000, 20000, 40000, 75000):
    for mode in ("fifo", "lifo", "rand"):
        t = time.time()
        r, du, sk, f = simulate(W, mode, max(n // 8, 1), rng)
        print("W=%6d %-4s rays %8d dups %7d (%.1f%%) skips %8d found %d  rays/vertex %.3f  [%.0fs]" % (W, mode, r, du, 100.0 * du / r, sk, f, r / V, time.time() - t), flush=True)
