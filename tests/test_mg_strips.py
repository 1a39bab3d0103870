"""Strip-distributed multigrid ("cg.mg" on N GPUs): the row ranges of hdd_mg_strip_plan, checked on the CPU by running a
V(1,1)-cycle in numpy the way csrc/multigrid.cu's vcycle() does - every rank sweeps only the rows of its plan on
NaN-poisoned full-size arrays, ghost rows of the level-0 right-hand side come from the neighbours once, the first
replicated level is summed over the ranks - and comparing the rows each rank owns with the replicated V-cycle.
A plan that is one row short anywhere shows up as NaN, a wrong range as a different number."""
import ctypes as C

import numpy as np
import pytest

from dune_hdd_b200 import capi


def plan(ny, c0, c1, n_dist):
    out = (C.c_int * (8 * n_dist + 3))()
    capi.check(capi.lib().hdd_mg_strip_plan(ny, c0, c1, n_dist, out))
    lv = [dict(zip(("pre_lo", "pre_hi", "b_lo", "b_hi", "up_lo", "up_hi", "pro_lo", "pro_hi"), out[8 * l:8 * l + 8]))
          for l in range(n_dist)]
    return lv, out[8 * n_dist], out[8 * n_dist + 1], out[8 * n_dist + 2]


# ---- a 9-point operator with variable coefficients and the V-cycle pieces on (ny+1) x (nx+1) vertex arrays ------------
def stencil(nx, ny, seed):
    rng = np.random.default_rng(seed)
    S = -0.1 - 0.05 * rng.random((9, ny + 1, nx + 1))
    S[4] = 1.5 + rng.random((ny + 1, nx + 1))
    return S


def apply_rows(S, x, rows):
    """(A x)[rows], reading x on rows +- 1 (clipped at the grid boundary)"""
    ny1, nx1 = x.shape
    lo, hi = rows
    out = np.zeros((hi - lo + 1, nx1))
    for ey in (-1, 0, 1):
        for ex in (-1, 0, 1):
            e = (ey + 1) * 3 + ex + 1
            for iy in range(lo, hi + 1):
                jy = iy + ey
                if jy < 0 or jy >= ny1:
                    continue
                xs = np.zeros(nx1)
                if ex == 0:
                    xs[:] = x[jy]
                elif ex == 1:
                    xs[:-1] = x[jy, 1:]
                else:
                    xs[1:] = x[jy, :-1]
                valid = np.ones(nx1, bool)
                if ex == 1:
                    valid[-1] = False
                if ex == -1:
                    valid[0] = False
                out[iy - lo] += np.where(valid, S[e, iy] * xs, 0.0)
    return out


def restrict_rows(r, rows):
    """full weighting onto coarse rows `rows`, reading fine rows 2 I +- 1"""
    nyf1, nxf1 = r.shape
    nxc1 = (nxf1 - 1) // 2 + 1
    lo, hi = rows
    out = np.zeros((hi - lo + 1, nxc1))
    for IY in range(lo, hi + 1):
        for dy in (-1, 0, 1):
            fy = 2 * IY + dy
            if fy < 0 or fy >= nyf1:
                continue
            for dx in (-1, 0, 1):
                w = (1.0 if dx == 0 else 0.5) * (1.0 if dy == 0 else 0.5)
                fx = 2 * np.arange(nxc1) + dx
                ok = (fx >= 0) & (fx < nxf1)
                out[IY - lo, ok] += w * r[fy, fx[ok]]
    return out


def interp_rows(xc, nxf1, rows):
    lo, hi = rows
    out = np.zeros((hi - lo + 1, nxf1))
    fx = np.arange(nxf1)
    cx, ox = fx >> 1, fx & 1
    for fy in range(lo, hi + 1):
        cy, oy = fy >> 1, fy & 1
        v = xc[cy, cx].copy()
        v[ox == 1] += xc[cy, cx[ox == 1] + 1]
        if oy:
            v += xc[cy + 1, cx]
            v[ox == 1] += xc[cy + 1, cx[ox == 1] + 1]
        out[fy - lo] = v * np.where(ox == 1, 0.5, 1.0) * (0.5 if oy else 1.0)
    return out


OMEGA = 0.8


def vcycle(Ss, b0, n_levels, ranges=None, reduce_coarse=None):
    """V(1,1) on levels 0 .. n_levels-1 (the last one: many Jacobi sweeps stand in for the dense solve).
    ranges: per distributed level the plan dict, None = whole levels; reduce_coarse(l, b) sums the right-hand side of the
    first replicated level over the ranks.  Arrays are NaN outside what this rank computed."""
    nd = 0 if ranges is None else len(ranges[0])
    lv, own_lo, own_hi = (ranges if ranges is not None else ([], 0, 0))
    b = [b0] + [None] * (n_levels - 1)
    x, r, y = [None] * n_levels, [None] * n_levels, [None] * n_levels
    full = lambda l: (0, Ss[l].shape[1] - 1)
    for l in range(n_levels - 1):
        S = Ss[l]
        dinv = OMEGA / S[4]
        pre = (lv[l]["pre_lo"], lv[l]["pre_hi"]) if l < nd else full(l)
        x[l] = np.full(S.shape[1:], np.nan)
        r[l] = np.full(S.shape[1:], np.nan)
        xb = dinv * b[l]  # x on every row where b is valid
        x[l][pre[0]:pre[1] + 1] = xb[pre[0]:pre[1] + 1]
        r[l][pre[0]:pre[1] + 1] = b[l][pre[0]:pre[1] + 1] - apply_rows(S, xb, pre)
        nxt = Ss[l + 1].shape[1:]
        b[l + 1] = np.full(nxt, np.nan)
        if l + 1 < nd:
            rc = (lv[l + 1]["b_lo"], lv[l + 1]["b_hi"])
        elif l + 1 == nd and nd > 0:
            b[l + 1][:] = 0.0
            rc = (own_lo, own_hi)
        else:
            rc = full(l + 1)
        b[l + 1][rc[0]:rc[1] + 1] = restrict_rows(r[l], rc)
        if l + 1 == nd and nd > 0:
            b[l + 1] = reduce_coarse(b[l + 1])
    L = n_levels - 1
    S = Ss[L]
    res = np.zeros(S.shape[1:])
    for _ in range(30):
        res = res + (OMEGA / S[4]) * (b[L] - apply_rows(S, res, full(L)))
    result = [None] * n_levels
    result[L] = res
    for l in range(n_levels - 2, -1, -1):
        S = Ss[l]
        dinv = OMEGA / S[4]
        up = (lv[l]["up_lo"], lv[l]["up_hi"]) if l < nd else full(l)
        pro = (lv[l]["pro_lo"], lv[l]["pro_hi"]) if l < nd else full(l)
        xp = x[l].copy()
        xp[pro[0]:pro[1] + 1] += interp_rows(result[l + 1], S.shape[2], pro)
        xp_masked = np.full_like(xp, np.nan)
        xp_masked[pro[0]:pro[1] + 1] = xp[pro[0]:pro[1] + 1]
        y[l] = np.full(S.shape[1:], np.nan)
        y[l][up[0]:up[1] + 1] = xp_masked[up[0]:up[1] + 1] + dinv[up[0]:up[1] + 1] * (
            b[l][up[0]:up[1] + 1] - apply_rows(S, xp_masked, up))
        result[l] = y[l]
    return result[0]


@pytest.mark.parametrize("world,n_dist", [(2, 1), (2, 2), (3, 2), (2, 3), (4, 3), (8, 3), (2, 4), (3, 4)])
def test_strip_plan_reproduces_the_replicated_vcycle(world, n_dist):
    g_probe = plan(1 << 12, 1 << 10, 1 << 11, n_dist)[3]
    rows_per_rank = ((2 * g_probe + 2 + (1 << n_dist) - 1) >> n_dist) << n_dist
    ny = rows_per_rank * world
    while (ny >> n_dist) % 2 == 1 and (ny >> n_dist) > 1:  # one more level below the distributed ones
        rows_per_rank += 1 << n_dist
        ny = rows_per_rank * world
    nx = max(16, 1 << (n_dist + 2))  # at least two cells per row on the coarsest level
    n_levels = n_dist + 2
    assert ny % (1 << (n_levels - 1)) == 0
    Ss = [stencil(nx >> l, ny >> l, 100 + l) for l in range(n_levels)]
    rng = np.random.default_rng(7)
    b_full = rng.standard_normal((ny + 1, nx + 1))
    reference = vcycle(Ss, b_full, n_levels)
    assert np.isfinite(reference).all()
    # every rank: its plan, the level-0 right-hand side on its rows plus the ghost rows, NaN elsewhere
    plans = [plan(ny, r * rows_per_rank, (r + 1) * rows_per_rank, n_dist) for r in range(world)]
    coarse_parts = {}

    def run(rank, reduce_coarse):
        lv, own_lo, own_hi, ghost = plans[rank]
        c0, c1 = rank * rows_per_rank, (rank + 1) * rows_per_rank
        assert ghost == g_probe or rank in (0, world - 1)
        b = np.full_like(b_full, np.nan)
        lo, hi = lv[0]["b_lo"], lv[0]["b_hi"]
        assert lo >= max(0, c0 - g_probe) and hi <= min(ny, c1 + g_probe)  # only the adjacent strips are needed
        b[lo:hi + 1] = b_full[lo:hi + 1]
        return vcycle(Ss, b, n_levels, (lv, own_lo, own_hi), reduce_coarse)

    # pass 1 collects every rank's contribution to the first replicated level, pass 2 uses the sum
    def collect(rank):
        def f(bpart):
            coarse_parts[rank] = bpart.copy()
            return bpart
        return f

    for r in range(world):
        run(r, collect(r))
    total = sum(coarse_parts.values())
    # own rows partition the first replicated level
    cover = sum((p != 0).any(axis=1).astype(int) for p in coarse_parts.values())
    assert (cover <= 1).all()
    for r in range(world):
        out = run(r, lambda bpart: total)
        c0, c1 = r * rows_per_rank, (r + 1) * rows_per_rank
        mine = out[c0:c1 + 1]
        assert np.isfinite(mine).all(), "rank %d: the plan is short of rows" % r
        assert np.abs(mine - reference[c0:c1 + 1]).max() <= 1e-13 * np.abs(reference).max()


def test_ghost_widths():
    """2 rows for one distributed level, 8 for two, 18 for three, 38 for four, symmetric for an interior strip"""
    for n_dist, g in ((1, 2), (2, 8), (3, 18), (4, 38)):
        lv, own_lo, own_hi, ghost = plan(4096, 1024, 1536, n_dist)
        assert ghost == g
        assert lv[0]["b_lo"] == 1024 - g and lv[0]["b_hi"] == 1536 + g
        assert (own_lo, own_hi) == (1024 >> n_dist, (1536 >> n_dist) - 1)
    lv, own_lo, own_hi, ghost = plan(4096, 3584, 4096, 3)
    assert lv[0]["b_hi"] == 4096 and own_hi == 512 and ghost == 18
