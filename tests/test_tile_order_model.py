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
