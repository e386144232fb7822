"""Independent cross-checks of the oracle's third-party-shaped parts (the Cabana list criterion and the ghost layer),
written out in numpy on random systems.  They pin what the reference's goldens pin only at one system: the inclusive
d^2 <= r^2 criterion, the coordinate tie-break of the half list, and which periodic images become ghosts.  CPU only."""
import ctypes as C
import itertools

import numpy as np
import pytest


def random_system(seed, n=400, box=(7.0, 8.0, 9.0)):
    rng = np.random.default_rng(seed)
    box = np.asarray(box)
    pos = rng.random((n, 3)) * box
    # a few exact ties along x and exactly-at-cutoff pairs exercise the tie-break and the inclusive criterion
    pos[1] = pos[0] + (0.0, 0.5, 0.0)
    pos[3] = pos[2] + (0.0, 0.0, 1.25)
    pos[5] = pos[4] + (1.5, 0.0, 0.0)  # |d| == r exactly (1.5 is exactly representable and so is its square)
    return np.mod(pos, box), box


def with_ghosts(oracle, pos, box, thickness):
    L = oracle.lib()
    n = len(pos)
    sub = oracle.subdomain([0, 0, 0], box, thickness)
    oa = np.zeros(12 * n, dtype=oracle.ATOM)
    oa["pos"][:n], oa["mass"][:n] = pos, 1.0
    corr = np.full(len(oa), -1, dtype=np.int64)
    ng = L.or_ghost_create_xyz(oa.ctypes.data, n, len(oa), C.byref(sub), corr.ctypes.data)
    assert ng >= 0
    return oa, corr, ng, sub


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_ghost_layer_is_the_set_of_qualifying_images(oracle, seed):
    """communication/GhostExchange.cpp:59-169: after the x, y, z passes the ghosts are exactly the images x + s L
    (s in {-1, 0, 1}^3 without 0) with s_d = +1 only for x_d < minInner_d and s_d = -1 only for x_d >= maxInner_d"""
    pos, box = random_system(seed)
    thickness = 1.5
    oa, corr, ng, sub = with_ghosts(oracle, pos, box, thickness)
    n = len(pos)
    lo, hi = np.array(sub.minInnerCorner), np.array(sub.maxInnerCorner)
    want = []
    for s in itertools.product((-1, 0, 1), repeat=3):
        if s == (0, 0, 0):
            continue
        ok = np.ones(n, dtype=bool)
        for d in range(3):
            if s[d] == 1:
                ok &= pos[:, d] < lo[d]
            elif s[d] == -1:
                ok &= pos[:, d] >= hi[d]
        for i in np.nonzero(ok)[0]:
            want.append((int(i), *s))
    got = []
    for g in range(n, n + ng):
        real = int(corr[g])
        while real >= n:  # ghosts of ghosts point at ghosts: follow to the real atom
            real = int(corr[real])
        shift = np.rint((oa["pos"][g] - pos[real]) / box).astype(int)
        assert np.array_equal(oa["pos"][g], pos[real] + shift * box) or np.allclose(oa["pos"][g], pos[real] + shift * box, atol=1e-12)
        got.append((real, *shift))
    assert sorted(got) == sorted(want) and len(set(got)) == len(got)


@pytest.mark.parametrize("seed,half", [(1, False), (2, False), (1, True), (3, True)])
def test_verlet_list_is_the_brute_force_pair_set(oracle, seed, half):
    """Cabana VerletList over local + ghost atoms against an O(N^2) enumeration: neighbours of local atom i are all
    j != i with d^2 <= r^2 (inclusive); the half list keeps j only if its position is lexicographically greater"""
    pos, box = random_system(seed)
    radius = 1.5
    oa, corr, ng, sub = with_ghosts(oracle, pos, box, radius)
    n = len(pos)
    counts, neigh = oracle.verlet_build(oa, 13, n + ng, 0, n, radius, 1.0, np.array(sub.minGhostCorner),
                                        np.array(sub.maxGhostCorner), half=half, width=96)
    allpos = oa["pos"][:n + ng]
    for i in range(n):
        d = allpos[i] - allpos
        d2 = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]
        sel = d2 <= radius * radius
        sel[i] = False
        if half:
            q, p = allpos, allpos[i]
            greater = (q[:, 0] > p[0]) | ((q[:, 0] == p[0]) & ((q[:, 1] > p[1]) | ((q[:, 1] == p[1]) & (q[:, 2] > p[2]))))
            sel &= greater
        assert sorted(neigh[i, :counts[i]].tolist()) == np.nonzero(sel)[0].tolist(), i
    if not half:
        assert 5 in neigh[4, :counts[4]] and 4 in neigh[5, :counts[5]]  # the pair at exactly |d| = r is a neighbour pair


def test_force_fold_conserves_the_total_force(oracle):
    """communication/AccumulateForce.cpp:25-47: ghost forces are added to their real atoms (through ghosts of ghosts) and
    zeroed: the sum over all atoms before equals the sum over the real atoms afterwards"""
    pos, box = random_system(7)
    oa, corr, ng, sub = with_ghosts(oracle, pos, box, 1.5)
    n = len(pos)
    rng = np.random.default_rng(0)
    oa["force"][:n + ng] = rng.normal(size=(n + ng, 3))
    total = oa["force"][:n + ng].sum(axis=0)
    oracle.lib().or_ghost_fold_force(oa.ctypes.data, n, ng, corr.ctypes.data)
    assert np.allclose(oa["force"][:n].sum(axis=0), total, rtol=0, atol=1e-10)
    assert np.all(oa["force"][n:n + ng] == 0.0)


def test_periodic_mapping_is_idempotent_and_lands_in_the_box(oracle):
    """communication/PeriodicMapping.cpp:30-58 for atoms at most one image away"""
    rng = np.random.default_rng(5)
    box = np.array([7.0, 8.0, 9.0])
    n = 2000
    a = np.zeros(n, dtype=oracle.ATOM)
    a["pos"] = (rng.random((n, 3)) * 3.0 - 1.0) * box  # [-L, 2L)
    a["pos"][0] = box                                   # exactly on the upper faces
    a["pos"][1] = 0.0
    sub = oracle.subdomain([0, 0, 0], box, 1.0)
    before = a["pos"].copy()
    oracle.lib().or_periodic_map(a.ctypes.data, n, C.byref(sub))
    once = a["pos"].copy()
    assert np.all(once >= 0.0) and np.all(once < box)
    assert np.allclose(np.mod(before, box), once, atol=1e-12) or np.all(np.abs(np.rint((before - once) / box) * box - (before - once)) < 1e-12)
    oracle.lib().or_periodic_map(a.ctypes.data, n, C.byref(sub))
    assert np.array_equal(a["pos"], once)
