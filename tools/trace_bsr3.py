#!/usr/bin/env python
"""Timeline of k_bsr3 on one SM (CTA 0) from clock64 stamps (-DSBP_B3_TRACE=1 build, tools/mkvar.sh trace):
consumer warps 0-7: [0 item top | 1 fragments ready | 2 boxes ready | 3 chain start | 4 chain issued | 5 stages released |
6 staging buffer free | 7 results staged], box / fragment producers (warps 8, 9): [0 stage free | 1 requests issued],
store helpers (warps 10-13): [0/2 buffer full | 1/3 stored and handed back] for their two consumer warps.

    SBP_B200_LIB=.../libsbp_b200_trace.so python tools/trace_bsr3.py [--mt 7] [--opts bsr_cw=12,bsr_help=0,bsr_sem=2]
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sbp_b200 as S  # noqa: E402

NI = 96


def build_semi():
    n1 = 36
    rng = np.random.default_rng(0)
    D2 = S.tensor_product_operator_2D(S.mattsson_nordstrom_2004(2, -1.0, 1.0, n1))
    nodes = D2.grid.copy()
    interior = np.ones(D2.N, bool)
    interior[np.unique(D2.boundary_indices)] = False
    nodes[interior] += (2.0 / (n1 - 1)) / 6 * (1 - 2 * rng.uniform(size=(int(interior.sum()), 2)))
    Ds = []
    for dim, lengths in ((0, (0.14, 0.08)), (1, (0.08, 0.14))):
        pat = S.neighborhood_sparsity_pattern(nodes, lengths)
        pat = pat | pat.T | np.eye(D2.N, dtype=bool)
        Dd = D2.Ds[dim].copy()
        extra = pat & (Dd == 0)
        Dd[extra] = 0.05 * rng.normal(size=int(extra.sum()))
        Ds.append(Dd)
    Dm = S.MultidimensionalMatrixDerivativeOperator(nodes, D2.boundary_indices, D2.normals, D2.weights(), D2.weights_boundary,
                                                    Ds, 2, corners=D2.corners)
    return S.MultidimensionalLinearAdvectionNonperiodicSemidiscretization(
        Dm, (1.0, 1.0), lambda x, t: float(np.exp(-30 * ((x[0] - t) ** 2 + (x[1] - t) ** 2))), storage="sparse")


def stats_free_running(T, cw):
    """12 consumer warps: a warp has a block in two of three items, so its stamps are indexed by ITS block count"""
    cons = T[:cw]
    lo, hi = 14, 56
    valid = (cons[:, lo:hi, 3] > 0) & (cons[:, lo:hi, 4] > 0)
    per_block = (cons[:, hi, 0] - cons[:, lo, 0]).mean() / (hi - lo)
    print(f"cycles per block of a warp {per_block:.0f} -> per item of 8 blocks on {cw} warps: {per_block * 8 / cw:.0f}")
    f = lambda a, b: (cons[:, lo:hi, b] - cons[:, lo:hi, a])[valid].mean()
    print("mean phase lengths per block: top->frag %.0f | frag->box %.0f | box->chain start %.0f | chain %.0f | chain end->released %.0f"
          " | released->next top %.0f" % (f(0, 1), f(1, 2), f(2, 3), f(3, 4), f(4, 5), (cons[:, lo + 1:hi + 1, 0] - cons[:, lo:hi, 5]).mean()))
    t_begin, t_end = cons[:, lo, 0].min(), cons[:, hi, 0].max()
    for s_ in range(4):
        ev = []
        for w in range(s_, cw, 4):
            for it in range(lo, hi):
                if cons[w, it, 3] > 0 and cons[w, it, 4] > 0:
                    ev += [(cons[w, it, 3], 1), (cons[w, it, 4], -1)]
        ev.sort()
        lvl, last, hist = 0, t_begin, {k: 0 for k in range(5)}
        for t, dlt in ev:
            if t > last:
                hist[lvl] += t - last
                last = t
            lvl += dlt
        tot = sum(hist.values())
        print(f"scheduler {s_}: chains in flight " + "  ".join(f"{k}: {100 * hist[k] / tot:.0f}%" for k in range(4)))
    prod = T[cw]
    print("box producer: stage free -> requests issued mean %.0f cycles; then waits %.0f for the next free stage" %
          ((prod[20:80, 1] - prod[20:80, 0]).mean(), (prod[21:81, 0] - prod[20:80, 1]).mean()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--opts", default="")
    ap.add_argument("--first", type=int, default=40)
    ap.add_argument("--count", type=int, default=2)
    ap.add_argument("--cw", type=int, default=8, help="consumer warps of the traced configuration (8 or 12)")
    ap.add_argument("--lockstep", action="store_true", help="12 consumer warps with chunks of 12 (-DSBP_BSR_CH=12): a block per warp and item")
    ap.add_argument("--nch", type=int, default=21, help="chunks per tile (21 for config 3 with chunks of 8, 14 with chunks of 12)")
    args = ap.parse_args()
    ctx = S.Context(0)
    for kv in filter(None, args.opts.split(",")):
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    semi = build_semi()
    N, B = semi.derivative.N, 65536
    u = ctx.upload(np.random.default_rng(1).uniform(-1, 1, (B, N)))
    du = u.similar()
    buf = S.RawBuffer(ctx, 32 * NI * 8 * 8)
    S._lib.check(ctx._lib.sbp_memset(ctx.handle, C.c_void_p(buf.ptr), 0, buf.nbytes))
    ctx._lib.sbp_debug_set_trace.argtypes = [C.c_void_p]
    assert ctx._lib.sbp_debug_set_trace(C.c_void_p(buf.ptr)) == 0
    for _ in range(3):
        semi(du, u, None, 0.1)
    ctx.synchronize()
    T = buf.to_host(np.int64, 32 * NI * 8).reshape(32, NI, 8)
    if args.cw != 8 and not args.lockstep:
        return stats_free_running(T, args.cw)
    NC = args.cw
    cons = T[:NC]
    t0 = cons[:, args.first, 0].min()
    names = ["top", "frag", "box", "c0", "c1", "rel", "free", "stg"]
    print("consumer warps, items", args.first, "..", args.first + args.count - 1, "(cycles relative to the first stamp)")
    for it in range(args.first, args.first + args.count):
        for w in range(NC):
            print(f"  item {it} warp {w}: " + " ".join(f"{n}={cons[w, it, k] - t0:6d}" for k, n in enumerate(names)))
    # statistics over items 21..83 (three whole tiles of 21 chunks for config 3)
    lo, hi = args.nch, 4 * args.nch
    sl = slice(lo, hi)
    valid = (cons[:, sl, 3] > 0) & (cons[:, sl, 4] > 0)  # blocks that multiplied (padding records of a tile's last chunk do not)

    def d(a, b, mask=None):
        x = (cons[:, sl, b] - cons[:, sl, a]).astype(np.float64)
        m = np.ones_like(x, bool) if mask is None else mask
        return x[m].mean()

    period = (cons[:, hi, 0] - cons[:, lo, 0]).mean() / (hi - lo)
    print(f"period per item {period:.0f} cycles; blocks that multiply: {valid.mean() * 100:.0f}%")
    print("mean phase lengths: top->frag %.0f | frag->box %.0f | box->chain start %.0f | chain %.0f | chain end->released %.0f | "
          "released->buffer free %.0f | staging %.0f | staged->next top %.0f" %
          (d(0, 1), d(1, 2), (cons[:, sl, 3] - cons[:, sl, 2])[valid].mean(), (cons[:, sl, 4] - cons[:, sl, 3])[valid].mean(),
           (cons[:, sl, 5] - cons[:, sl, 4])[valid].mean(), d(5, 6), d(6, 7), (cons[:, lo + 1:hi + 1, 0] - cons[:, sl, 7]).mean()))
    # how many warps of a scheduler are inside their chain at the same time
    t_begin, t_end = cons[:, lo, 0].min(), cons[:, hi, 0].max()
    for s_ in range(4):
        ws = [w for w in range(NC) if w % 4 == s_]
        ev = []
        for w in ws:
            for it in range(lo, hi):
                if cons[w, it, 3] > 0 and cons[w, it, 4] > 0:
                    ev.append((cons[w, it, 3], 1))
                    ev.append((cons[w, it, 4], -1))
        ev.sort()
        lvl, last, hist = 0, t_begin, {0: 0, 1: 0, 2: 0, 3: 0}
        for t, dlt in ev:
            hist[lvl] += t - last
            lvl += dlt
            last = t
        hist[0] += t_end - last
        tot = sum(hist.values())
        print(f"scheduler {s_}: chains in flight " + "  ".join(f"{k}: {100 * hist[k] / tot:.0f}%" for k in range(4)))
    # per chunk of the tile: when do the boxes arrive relative to the moment the LAST warp is ready for them?
    print("chunk: mean over 3 tiles of (box-ready - fragments-ready) of the last warp, chain length, release spread")
    for c in range(args.nch):
        its = [lo + c + args.nch * k for k in range(3)]
        wait = np.mean([(cons[:, it, 2] - cons[:, it, 1]).min() for it in its])
        chain = np.mean([np.mean([cons[w, it, 4] - cons[w, it, 3] for w in range(NC) if cons[w, it, 3] > 0]) for it in its])
        rel = np.mean([cons[:, it, 5].max() - cons[:, it, 5].min() for it in its])
        per = np.mean([cons[:, it + 1, 0].mean() - cons[:, it, 0].mean() for it in its])
        print(f"  chunk {c:2d}: min box wait {wait:6.0f}  chain {chain:6.0f}  release spread {rel:6.0f}  item period {per:6.0f}")
    prod = T[NC]
    print("box producer: stage free -> requests issued mean %.0f cycles; then waits %.0f for the next free stage" %
          ((prod[sl, 1] - prod[sl, 0]).mean(), (prod[lo + 1:hi + 1, 0] - prod[sl, 1]).mean()))
    for h in range(4):
        hp = T[NC + 2 + h]
        print(f"helper {h}: full->stored warp {h}: %.0f, warp {h + 4}: %.0f" % ((hp[sl, 1] - hp[sl, 0]).mean(), (hp[sl, 3] - hp[sl, 2]).mean()))


if __name__ == "__main__":
    main()
