#!/usr/bin/env python3
"""Derive the fully symmetric triangle quadrature rules that petgem_b200/quadrature2d.py tabulates.

A rule is a set of orbits under the symmetry group of the triangle (barycentric coordinates):
S3 = the centroid (1 point, unknown: weight), S21(a) = permutations of (a, a, 1-2a) (3 points,
unknowns: weight, a), S111(a, b) = permutations of (a, b, 1-a-b) (6 points, unknowns: weight, a, b).
For each degree d the orbit structure below (the classical minimal-point structures of
D. A. Dunavant, Int. J. Numer. Meth. Eng. 21 (1985) 1129-1148) is solved for exactness on every
monomial x^i y^j, i + j <= d, over the unit triangle (integral i! j! / (i+j+2)!) by
Gauss-Newton from random starts, keeping the solution with positive weights and interior points.
Prints the orbit parameters with 17 significant digits.
"""
import math
import sys

import numpy as np
from scipy.optimize import least_squares

STRUCTURE = {  # degree: (number of S3, S21, S111 orbits)
    2: (0, 1, 0), 4: (0, 2, 0), 6: (0, 2, 1), 8: (1, 3, 1), 10: (1, 2, 3), 12: (0, 5, 3),
}


def expand(params, struct):
    n3, n21, n111 = struct
    pts, wts = [], []
    k = 0
    for _ in range(n3):
        pts.append((1 / 3, 1 / 3, 1 / 3)); wts.append(params[k]); k += 1
    for _ in range(n21):
        w, a = params[k], params[k + 1]; k += 2
        for q in ((a, a, 1 - 2 * a), (a, 1 - 2 * a, a), (1 - 2 * a, a, a)):
            pts.append(q); wts.append(w)
    for _ in range(n111):
        w, a, b = params[k], params[k + 1], params[k + 2]; k += 3
        c = 1 - a - b
        for q in ((a, b, c), (a, c, b), (b, a, c), (b, c, a), (c, a, b), (c, b, a)):
            pts.append(q); wts.append(w)
    return np.array(pts), np.array(wts)


def residual(params, struct, degree):
    pts, wts = expand(params, struct)
    x, y = pts[:, 0], pts[:, 1]
    res = []
    for i in range(degree + 1):
        for j in range(degree + 1 - i):
            exact = math.factorial(i) * math.factorial(j) / math.factorial(i + j + 2)
            res.append((wts * x**i * y**j).sum() - exact)
    return np.array(res)


def solve(degree, seed=0, tries=400):
    struct = STRUCTURE[degree]
    n3, n21, n111 = struct
    rng = np.random.default_rng(seed)
    best = None
    for _ in range(tries):
        p0 = []
        for _ in range(n3):
            p0.append(rng.uniform(0.01, 0.1))
        for _ in range(n21):
            p0 += [rng.uniform(0.005, 0.08), rng.uniform(0.02, 0.49)]
        for _ in range(n111):
            a = rng.uniform(0.01, 0.4); b = rng.uniform(0.05, 0.9 - a)
            p0 += [rng.uniform(0.005, 0.05), a, b]
        sol = least_squares(residual, p0, args=(struct, degree), xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=2000)
        pts, wts = expand(sol.x, struct)
        if np.abs(sol.fun).max() < 5e-16 and (wts > 0).all() and (pts > 0).all():
            key = tuple(np.round(np.sort(wts), 10))
            if best is None:
                best = {}
            best.setdefault(key, sol.x)
    return struct, best


if __name__ == "__main__":
    degrees = [int(a) for a in sys.argv[1:]] or sorted(STRUCTURE)
    for d in degrees:
        struct, sols = solve(d)
        print("degree", d, "structure", struct, "distinct solutions:", 0 if sols is None else len(sols))
        for key, x in (sols or {}).items():
            print("   ", ", ".join("%.17g" % v for v in x))
