"""Per-block timeline of the rank chain kernel (DMO_RANK_TRACE): where does a link of the chain spend its time?

Slots (thread 0 of each block; [0..7] globaltimer ns, [8..15] SM clock):
  0 block start   1 tables done (bulk loop starts)   2 starts waiting for tile b-2   3 tile b-2 in shared memory
  4 fold done (starts waiting for predecessor b-1)   5 predecessor ranks seen   6 published   7 smid
"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
    M = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    path = os.path.join(tempfile.gettempdir(), "rank_trace.bin")
    os.environ["DMO_RANK_TRACE"] = path
    from dmosopt_b200 import _lib as L

    L.context()
    rng = np.random.default_rng(0)
    Y = rng.random((n, M))
    L.rank_nd(Y)
    L.rank_nd(Y)
    t = np.fromfile(path, dtype=np.int64).reshape(-1, 32)
    nb = t.shape[0]
    g = t[:, :16].astype(np.float64)
    c = t[:, 16:32].astype(np.float64)
    g0 = g[:, 0].min()
    g -= g0
    pub = g[:, 6]
    link = np.diff(pub)
    print(f"n={n} M={M}: {nb} blocks, kernel span {pub.max() / 1e3:.1f} us; globaltimer granularity ~{np.min(link[link > 0]) if (link > 0).any() else 0:.0f} ns")
    print(f"publish-to-publish: mean {link.mean():.0f} ns, median {np.median(link):.0f}, p10 {np.percentile(link, 10):.0f}, p90 {np.percentile(link, 90):.0f}")
    sel = slice(max(2, nb // 8), nb)  # steady state
    clk = lambda a, b: (c[sel, b] - c[sel, a])
    names = [("tables (0->1)", 0, 1), ("bulk up to tile b-3 (1->2)", 1, 2), ("wait+load tile b-2 (2->3)", 2, 3), ("tile b-2 pairs + fold (3->4)", 3, 4),
             ("wait predecessor (4->5)", 4, 5), ("resolve + publish (5->6)", 5, 6)]
    for nm, a, b in names:
        ok = (t[sel, a] > 0) & (t[sel, b] > 0)  # a skipped tile leaves its stamps empty
        d = clk(a, b)[ok]
        if d.size == 0:
            continue
        print(f"  {nm:32s} SM cycles: mean {d.mean():9.0f}  median {np.median(d):9.0f}  p90 {np.percentile(d, 90):9.0f}")
    # how long after the predecessor's publish does this block see it, and publish itself (global ns)
    see = g[1:, 5] - pub[:-1]
    own = pub[1:] - g[1:, 5]
    ready_before = (g[1:, 4] <= pub[:-1]).mean()
    print(f"  predecessor publish -> seen here: mean {see[sel].mean():.0f} ns (median {np.median(see[sel]):.0f});  seen -> own publish: mean {own[sel].mean():.0f} ns")
    print(f"  fraction of blocks already waiting when the predecessor published: {ready_before:.2f}")
    # tile b-2: publish of b-2 -> this block has it in shared memory
    ok3 = t[2:, 3] > 0
    lag2 = (g[2:, 3] - pub[:-2])[ok3]
    if lag2.size:
        print(f"  publish(b-2) -> tile b-2 loaded here: mean {lag2.mean():.0f} ns (median {np.median(lag2):.0f})")
    print("  blocks per SM:", np.bincount(t[:, 7].astype(int)).max())
    dump_gaps(g, pub)
    if t[:, 8].any():  # grid phase stamps (M <= 3): 8 query start, 9 query end, 10 tree update fenced, 11 done flag set
        gq = t[:, 8] > 0
        lag = int(os.environ.get("DMO_RANK_LAG", "32"))
        idx = np.flatnonzero(gq)
        print(f"  grid query duration (8->9): mean {(g[gq, 9] - g[gq, 8]).mean():.0f} ns, p90 {np.percentile(g[gq, 9] - g[gq, 8], 90):.0f}")
        print(f"  publish -> tree update fenced (6->10): mean {(g[:, 10] - g[:, 6]).mean():.0f} ns, p90 {np.percentile(g[:, 10] - g[:, 6], 90):.0f}")
        print(f"  fenced -> done flag (10->11): mean {(g[:, 11] - g[:, 10]).mean():.0f} ns, p90 {np.percentile(g[:, 11] - g[:, 10], 90):.0f}")
        dn = np.diff(g[:, 11])
        print(f"  done-to-done: mean {dn.mean():.0f} ns, median {np.median(dn):.0f}, p90 {np.percentile(dn, 90):.0f}")
        print(f"  done[b] - publish[b]: mean {(g[:, 11] - g[:, 6]).mean():.0f} ns")
        slack = g[idx, 8] - g[idx - lag - 1, 11]
        print(f"  query start - done[b-lag-1]: mean {slack.mean():.0f} ns, median {np.median(slack):.0f}")
        turn = pub[idx - 1] - g[idx, 9]
        print(f"  predecessor publish - query end (slack before own turn; negative = late): mean {turn.mean():.0f} ns, p10 {np.percentile(turn, 10):.0f}")


def dump_gaps(g, pub, k=14):
    link = np.diff(pub)
    order = np.argsort(-link)[:k]
    print("  largest publish gaps: block, gap us, [fold done - pred publish] us, [tile b-2 loaded - publish(b-2)] us, [start(0) - pred publish] us, [tables done(1) - pred publish] us")
    for j in sorted(order):
        b = j + 1
        print(f"    b={b:5d} gap={link[j] / 1e3:7.1f}  ready-after-pred={(g[b, 4] - pub[b - 1]) / 1e3:8.1f}  tile2-lag={(g[b, 3] - pub[b - 2]) / 1e3 if b >= 2 else 0:8.1f}"
              f"  start={(g[b, 0] - pub[b - 1]) / 1e3:9.1f}  tables={(g[b, 1] - pub[b - 1]) / 1e3:9.1f}  query-end={(g[b, 9] - pub[b - 1]) / 1e3 if g[b, 9] > 0 else 0:9.1f}")


if __name__ == "__main__":
    main()
