"""CPU: Python model of decode_tile() in csrc/gemm_nt.cu (grouped, L2-friendly tile walk).  Every tile of the
rectangle / lower trapezoid must be produced exactly once for any grid shape -- the GPU tests only see a
handful of shapes."""
import pytest


def decode_full(lin, tm, tn, GM):
    per = GM * tn
    g = lin // per
    first = g * GM
    gs = min(GM, tm - first)
    ing = lin - g * per
    return first + ing % gs, ing // gs


def decode_lower(lin, tm, tn, GM):
    rem, first = lin, 0
    while True:
        gs = min(GM, tm - first)
        cnt = sum(min(r, tn - 1) + 1 for r in range(first, first + gs))
        if rem < cnt or first + gs >= tm:
            break
        rem -= cnt
        first += gs
    full = min(first, tn)
    if rem < full * gs:
        return first + rem % gs, rem // gs
    rem -= full * gs
    bn = full
    while True:
        nrows = first + gs - max(bn, first)
        if rem < nrows or bn + 1 >= tn:
            break
        rem -= nrows
        bn += 1
    return max(bn, first) + rem, bn


@pytest.mark.parametrize("tm,tn", [(1, 1), (5, 7), (37, 3), (64, 64), (17, 40), (33, 1), (16, 16), (31, 2), (256, 256)])
@pytest.mark.parametrize("GM", [4, 16])
def test_full_grid_walk_is_a_permutation(tm, tn, GM):
    got = sorted(decode_full(l, tm, tn, GM) for l in range(tm * tn))
    assert got == sorted((a, b) for a in range(tm) for b in range(tn))


@pytest.mark.parametrize("tm,tn", [(1, 1), (5, 5), (37, 37), (100, 30), (17, 1), (33, 32), (256, 256), (252, 4)])
@pytest.mark.parametrize("GM", [4, 16])
def test_lower_trapezoid_walk_is_a_permutation(tm, tn, GM):
    n = sum(min(r, tn - 1) + 1 for r in range(tm))          # grid size computed by gemm_nt_launch
    got = sorted(decode_lower(l, tm, tn, GM) for l in range(n))
    assert got == sorted((bm, bn) for bm in range(tm) for bn in range(min(bm, tn - 1) + 1))
    # heaviest tiles (small bm: longest k range in the W'W launch) come first
    assert decode_lower(0, tm, tn, GM)[0] == 0


def _own_before(t, rem, div, mod):
    """tiles t' < t owned by rank `rem` under block-cyclic ownership (t' // div) % mod == rem  (shard_own_before, gemm_nt.cu)"""
    cyc = div * mod
    c, r = divmod(t, cyc)
    extra = min(max(r - rem * div, 0), div)
    return c * div + extra


def test_compact_ownership_enumeration_matches_brute_force():
    """gemm_nt.cu, own_compact launches: ordinal -> global tile row in closed form, grid = ceil(n_own / 8) * 8 * tn.  Every owned
    tile row of the launch must be visited exactly once per column, none that is foreign."""
    import itertools
    for div, mod, off, tm, tn in itertools.product((1, 2, 4, 8), (2, 3, 4, 8), (0, 1, 5, 12, 40), (1, 7, 16, 33), (1, 4)):
        for rem in range(mod):
            owned = [bm for bm in range(tm) if ((bm + off) // div) % mod == rem]
            first = _own_before(off, rem, div, mod)
            n_own = _own_before(off + tm, rem, div, mod) - first
            assert n_own == len(owned), (div, mod, off, tm, rem)
            seen = []
            GO = 8
            ntiles = (n_own + GO - 1) // GO * GO * tn
            for lin in range(ntiles):
                group, in_group = divmod(lin, GO * tn)
                o = group * GO + in_group % GO
                bn = in_group // GO
                if o >= n_own:
                    continue
                og = o + first
                tg = ((og // div) * mod + rem) * div + og % div
                seen.append((tg - off, bn))
            assert sorted(seen) == sorted((bm, bn) for bm in owned for bn in range(tn)), (div, mod, off, tm, tn, rem)
