#!/usr/bin/env python3
"""Derive the fully symmetric triangle quadrature rules that petgem_b200/quadrature2d.py tabulates.

A rule is a set of orbits under the symmetry group of the triangle (barycentric coordinates):
S3 = the centroid (1 point, unknown: weight), S21(a) = permutations of (a, a, 1-2a) (3 points,
unknowns: weight, a), S111(a, b) = permutations of (a, b, 1-a-b) (6 points, unknowns: weight, a, b).
For each degree d the orbit structure below (the classical minimal-point structures of
D. A. Dunavant, Int. J. Numer. Meth. Eng. 21 (1985) 1129-1148) is solved for exactness on every
monomial x^i y^j, i + j <= d, over the unit triangle (integral i! j! / (i+j+2)!) by
Gauss-Newton from random starts, keeping the solution with positive weights and interior points.
Prints the orbit parameters with 17 significant digits.  Degrees 2..8 converge in seconds and the
25-point rule (degree 10) in ~10 minutes of random starts; for the 33-point rule (degree 12) random
starts did not converge; re-solving it from the published 15-digit values (`START`) takes the moment
residual from 2e-15 to 3e-17 but moves the points by 1e-10 (ill conditioned), so quadrature2d.py keeps the
published digits for that rule.
"""
import math
import sys

import numpy as np
from scipy.optimize import least_squares

STRUCTURE = {  # degree: (number of S3, S21, S111 orbits)
    2: (0, 1, 0), 4: (0, 2, 0), 6: (0, 2, 1), 8: (1, 3, 1), 10: (1, 2, 3), 12: (0, 5, 3),
}


# published 33-point rule (weights for the unit-area normalisation halved), starting guess only
START = {
    12: [0.025731066440455 / 2, 0.488217389773805, 0.043692544538038 / 2, 0.439724392294460,
         0.062858224217885 / 2, 0.271210385012116, 0.034796112930709 / 2, 0.127576145541586,
         0.006166261051559 / 2, 0.021317350453210,
         0.040371557766381 / 2, 0.115343494534698, 0.275713269685514,
         0.022356773202303 / 2, 0.022838332222257, 0.281325580989940,
         0.017316231108659 / 2, 0.025734050548330, 0.116251915907597],
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


def _moments(degree):
    ij = [(i, j) for i in range(degree + 1) for j in range(degree + 1 - i)]
    ex = np.array([math.factorial(i) * math.factorial(j) / math.factorial(i + j + 2) for i, j in ij])
    return np.array([i for i, _ in ij]), np.array([j for _, j in ij]), ex


def residual(params, struct, degree, mom=None):
    I, J, ex = mom or _moments(degree)
    pts, wts = expand(params, struct)
    x, y = pts[:, 0], pts[:, 1]
    return (wts[None, :] * x[None, :] ** I[:, None] * y[None, :] ** J[:, None]).sum(axis=1) - ex


def jacobian(params, struct, degree, mom=None):
    """Analytic Jacobian of `residual` with respect to the orbit parameters."""
    I, J, _ = mom or _moments(degree)
    n3, n21, n111 = struct

    def mono(x, y):
        return x**I * y**J

    def dmono(x, y):
        dx = np.where(I > 0, I * x ** np.maximum(I - 1, 0) * y**J, 0.0)
        dy = np.where(J > 0, J * x**I * y ** np.maximum(J - 1, 0), 0.0)
        return dx, dy

    cols, k = [], 0
    for _ in range(n3):
        cols.append(mono(1 / 3, 1 / 3))
        k += 1
    for _ in range(n21):
        w, a = params[k], params[k + 1]
        k += 2
        c = 1 - 2 * a
        orbit = [(a, a, (1, 1)), (a, c, (1, -2)), (c, a, (-2, 1))]  # (x, y, d(x, y)/da)
        cols.append(sum(mono(x, y) for x, y, _ in orbit))
        d = 0
        for x, y, (sx, sy) in orbit:
            dx, dy = dmono(x, y)
            d = d + w * (dx * sx + dy * sy)
        cols.append(d)
    for _ in range(n111):
        w, a, b = params[k], params[k + 1], params[k + 2]
        k += 3
        c = 1 - a - b
        orbit = [(a, b, (1, 0), (0, 1)), (a, c, (1, -1), (0, -1)), (b, a, (0, 1), (1, 0)),
                 (b, c, (0, -1), (1, -1)), (c, a, (-1, 1), (-1, 0)), (c, b, (-1, 0), (-1, 1))]
        cols.append(sum(mono(x, y) for x, y, _, _ in orbit))
        da = db = 0
        for x, y, (xa, ya), (xb, yb) in orbit:
            dx, dy = dmono(x, y)
            da = da + w * (dx * xa + dy * ya)
            db = db + w * (dx * xb + dy * yb)
        cols.append(da)
        cols.append(db)
    return np.stack(cols, axis=1)


def solve(degree, seed=0, tries=4000, want=4):
    """Levenberg-Marquardt from random starts until `want` admissible solutions (positive weights,
    interior points; usually copies of one or two distinct rules) have been found or the tries are
    used up."""
    struct = STRUCTURE[degree]
    n3, n21, n111 = struct
    mom = _moments(degree)
    rng = np.random.default_rng(seed)
    best = {}
    found = 0
    if degree in START:
        sol = least_squares(residual, START[degree], jac=jacobian, args=(struct, degree, mom), method="lm",
                            xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=2000)
        return struct, {"from the published starting guess": sol.x}
    for _ in range(tries):
        p0 = []
        for _ in range(n3):
            p0.append(rng.uniform(0.01, 0.1))
        for _ in range(n21):
            p0 += [rng.uniform(0.005, 0.08), rng.uniform(0.02, 0.49)]
        for _ in range(n111):
            a = rng.uniform(0.01, 0.4)
            b = rng.uniform(0.05, 0.9 - a)
            p0 += [rng.uniform(0.005, 0.05), a, b]
        sol = least_squares(residual, p0, jac=jacobian, args=(struct, degree, mom), method="lm", xtol=1e-15,
                            ftol=1e-15, gtol=1e-15, max_nfev=2000)
        pts, wts = expand(sol.x, struct)
        if np.abs(sol.fun).max() < 5e-16 and (wts > 0).all() and (pts > 0).all():
            key = tuple(np.round(np.sort(wts), 9))
            best.setdefault(key, sol.x)
            found += 1
            if found >= want:
                break
    return struct, best


if __name__ == "__main__":
    degrees = [int(a) for a in sys.argv[1:]] or sorted(STRUCTURE)
    for d in degrees:
        struct, sols = solve(d)
        print("degree", d, "structure", struct, "distinct solutions:", 0 if sols is None else len(sols))
        for key, x in (sols or {}).items():
            print("   ", ", ".join("%.17g" % v for v in x))
