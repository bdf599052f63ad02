"""The warp-balanced N-D tally of the CUDA path (mcb_device.cuh: tally_nd_balanced / ndb_item) computes every cell of
Field::accumulate's N-D walk (field.cpp:156-218) independently: "the cell entered at the m-th crossing of axis d".  This test
restates BOTH formulations in plain Python -- the reference's walk over the sorted, tie-merged crossing parameters with its (1.0)
sentinel, and the per-item formulation with the same par() expression, count estimate + fix-up and tie ownership -- and checks
that they give the same (cell, weight) deposits on random and on degenerate segments (corner crossings, integer coordinates,
ends a rounding error outside the grid, axes that are not crossed).  The GPU parity tests check the device code itself."""
import math
import random

INF = float("inf")
DBL_MIN = 2.2250738585072014e-308


def c2i(c, mx):
    f = math.floor(c)
    return 0 if f < 0 else (mx if f > mx else int(f))


def setup(b3, e3, mx, strd):
    rec, col = [], 0
    for d in range(3):
        b, e = c2i(b3[d], mx[d]), c2i(e3[d], mx[d])
        col += b * strd[d]
        dc = e3[d] - b3[d]
        on = (not abs(dc) < DBL_MIN) and b != e
        fwd = b < e
        rec.append(dict(bc=b3[d], dc=dc, idc=(1.0 / dc if on else 0.0), nxt=(b + 1 if fwd else b), n=(abs(e - b) if on else 0),
                        dst=(strd[d] if fwd else -strd[d]), pm=(1 if fwd else -1)))
    return rec, col


def par(r, c):
    return (float(r["nxt"] + c * r["pm"]) - r["bc"]) * r["idc"]


def reference_walk(b3, e3, mx, strd):
    """field.cpp:156-218 as a 3-way merge of the monotone crossing sequences + the (1.0) sentinel (oracle/mc_oracle.cpp)."""
    rec, col = setup(b3, e3, mx, strd)
    cnt, prev, sentinel, out = [0, 0, 0], 0.0, True, {}
    pr = [par(rec[d], 0) if rec[d]["n"] > 0 else INF for d in range(3)]
    more = True
    while more:
        best = min(pr)
        key = 1.0 if (sentinel and 1.0 <= best) else best
        out[col] = out.get(col, 0.0) + (key - prev)
        prev = key
        if key == 1.0:
            sentinel = False
        rest = sentinel
        for d in range(3):
            if pr[d] == key:
                col += rec[d]["dst"]; cnt[d] += 1
                pr[d] = par(rec[d], cnt[d]) if cnt[d] < rec[d]["n"] else INF
            rest = rest or pr[d] < INF
        more = rest
    return out


def item_walk(b3, e3, mx, strd):
    """tally_nd_balanced: the first cell by the owner lane, then one independent item per face crossing (ndb_item)."""
    rec, col0 = setup(b3, e3, mx, strd)
    out = {}
    w0 = 1.0
    for r in rec:
        if r["n"] > 0:
            w0 = min(w0, par(r, 0))
    out[col0] = out.get(col0, 0.0) + w0
    for k in range(sum(r["n"] for r in rec)):
        m, d = k, 0
        if m >= rec[0]["n"]:
            m -= rec[0]["n"]; d = 1
            if m >= rec[1]["n"]:
                m -= rec[1]["n"]; d = 2
        r = rec[d]
        t, tn, col, skip = par(r, m), INF, col0 + (m + 1) * r["dst"], False
        if m + 1 < r["n"]:
            tn = par(r, m + 1)
        ea, eb = (0 if d == 2 else d + 1), (2 if d == 0 else d - 1)
        if rec[ea]["n"] == 0:
            ea, eb = eb, ea
        for e in (ea, eb):
            r = rec[e]; n = r["n"]
            if n > 0:
                x = t * r["dc"] + r["bc"]
                c = (math.floor(x) - r["nxt"] + 1) if r["dst"] > 0 else (r["nxt"] - math.ceil(x) + 1)
                c = min(max(c, 0), n)
                pc = par(r, c) if c < n else INF
                while pc <= t:
                    c += 1; pc = par(r, c) if c < n else INF
                tie = False
                while c > 0:
                    pp = par(r, c - 1)
                    if pp <= t:
                        tie = pp == t
                        break
                    c -= 1; pc = pp
                if tie and e < d:
                    skip = True
                tn = min(tn, pc); col += c * r["dst"]
        w = min(tn, 1.0) - t
        if not skip and w > 0:
            out[col] = out.get(col, 0.0) + w
    return out


def test_item_walk_equals_the_reference_walk():
    rnd = random.Random(1)
    mx, strd = [7, 15, 31], [1, 8, 128]

    def coord(mode, d):
        top = mx[d] + 1
        if mode == 0:
            return rnd.uniform(0, top)
        if mode == 1:                      # integer and half-integer coordinates: exact ties (corner crossings)
            return rnd.choice([0.0, float(top), rnd.uniform(0, top), float(rnd.randint(0, top)), rnd.randint(0, top) + 0.5])
        return rnd.choice([-1e-17, top + 1e-14, top * (1 + 1e-16), rnd.uniform(0, top)])      # a rounding error outside the grid

    worst = 0.0
    for it in range(30000):
        mode = it % 3
        b3 = [coord(mode, d) for d in range(3)]; e3 = [coord(mode, d) for d in range(3)]
        if it % 7 == 0:
            e3[0] = b3[0]                  # axis not crossed (2-D grid)
        if it % 11 == 0:
            e3[1] = b3[1]
        a, b = reference_walk(b3, e3, mx, strd), item_walk(b3, e3, mx, strd)
        err = max(abs(a.get(k, 0.0) - b.get(k, 0.0)) for k in set(a) | set(b))
        worst = max(worst, err)
        assert err <= 4e-16, (b3, e3, err)
    assert worst <= 4e-16
