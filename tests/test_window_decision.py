"""CPU check of the claim behind the thread-per-digest kernel's boundary decision
(crick_b200/csrc/tdigest_lanes.cu): outside the per-centroid weight window
[Wlo, Whi] = total * (qlim(q1) -+ margin) the reference's own decision fl(kend - k1) > 1
(tdigest_stubs.c:178-216, glibc asin) is what the window says, so only cumulative weights inside the
window need the exact k-scale pair.  The window is recomputed here with the kernel's formulas in
float64 and attacked with cumulative weights placed a few ulp to 1e-9 either side of both window
edges, over compressions, totals of very different magnitude and boundaries from the far tails to
the top (where no centroid can open any more)."""
import math

import numpy as np
import pytest

MARGIN = 6e-13
PI, PI_2 = math.pi, math.pi / 2


def kscale(c, q):
    # tdigest_stubs.c:178-189 with the explicit roundings the kernels use (no FMA)
    return (c * (math.asin(2.0 * q - 1.0) + PI_2)) / PI


def window(c, total, q1):
    kC, kS = math.cos(PI / c), math.sin(PI / c)
    kH = math.sin(PI_2 / c) ** 2
    kQtop = 0.5 * (1.0 + kC)
    q1 = min(q1, 1.0)
    r = math.sqrt(q1 * (1.0 - q1))
    qlim = q1 * kC + kH + r * kS
    m = MARGIN
    if 0.0 < r < 3.2e-4:          # the quantisation of y1 = fl(2 q1 - 1) next to -+1, seen through asin
        m += 5e-17 / r
    wlo = total * (qlim - m)
    whi = math.inf if q1 > kQtop - MARGIN else total * (qlim + m)
    if q1 > kQtop - MARGIN:
        wlo = total * (1.0 - 16.0 * MARGIN - 1e-11)
    return wlo, whi, qlim


@pytest.mark.parametrize("comp", [20.0, 100.0, 200.0, 500.0, 1000.0])
def test_window_never_contradicts_the_exact_decision(comp):
    rng = np.random.default_rng(int(comp))
    checked = inside = 0
    for total in (1.0, 1000.0, 1e9 + 0.5, 3.7e15, 2.5e-3):
        rtotal = 1.0 / total
        # q1 of the open centroid's start: tails, middle, near the top, exactly 0 (first centroid)
        q1s = np.concatenate([[0.0], 10.0 ** rng.uniform(-30, -1, 150), rng.uniform(0, 1, 120),
                              1.0 - 10.0 ** rng.uniform(-15, -1, 90)])
        for q1 in q1s:
            W1 = q1 * total
            q1k = min(W1 * rtotal, 1.0)                    # what the kernel derives from W1
            wlo, whi, qlim = window(comp, total, q1k)
            k1 = 0.0 if q1 == 0.0 else kscale(comp, W1 / total)
            cands = []
            for edge in (wlo, whi):
                if not math.isfinite(edge):
                    continue
                for rel in (0.0, 1e-16, 4e-16, 1e-15, 1e-14, 1e-12, 1e-9):
                    cands += [edge * (1.0 + rel), edge * (1.0 - rel)]
                cands += [np.nextafter(edge, math.inf), np.nextafter(edge, -math.inf)]
            cands += list(total * rng.uniform(q1k, 1.0, 6)) + [total, W1]
            for Wc in cands:
                if not (W1 <= Wc <= total * (1.0 + 1e-12)):
                    continue                               # cumulative weights only grow, up to ~total
                exact_open = not (kscale(comp, Wc / total) - k1 <= 1.0)
                if Wc > whi:
                    assert exact_open, (comp, total, q1, Wc, wlo, whi)
                elif Wc < wlo:
                    assert not exact_open, (comp, total, q1, Wc, wlo, whi)
                else:
                    inside += 1
                checked += 1
    assert checked > 5000 and inside < checked          # the attack did reach both sides of the window


def test_window_is_narrow():
    """The fallback is rare because the window is 1.2e-12 of the total weight wide (plus the top
    region, which only the last centroid of a flush can reach); it only widens for a boundary behind
    less than ~1e-7 of the total weight."""
    for comp in (20.0, 100.0, 1000.0):
        for q1 in (0.0, 1e-6, 0.3, 0.9):
            wlo, whi, _ = window(comp, 1.0, q1)
            assert 0.0 < whi - wlo < 1.6e-12
        wlo, whi, _ = window(comp, 1.0, 1e-12)
        assert 1e-10 < whi - wlo < 1e-9
