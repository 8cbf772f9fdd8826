"""Fully symmetric quadrature rules on the unit triangle, degrees 2..12 (even).

The reference tabulates "order 2p" rules for the MT boundary integrals
(``hvfem.compute2DGaussPoints``, hvfem.py:1613-2300, called at solver.py:323-324): the classical
minimal-point symmetric rules of D. A. Dunavant (Int. J. Numer. Meth. Eng. 21, 1985) with 3, 6, 12,
16, 25 and 33 points.  The integrand of that boundary term is not polynomial (the 1-D MT field
varies along z), so unlike the volume rule (``basis.tet_quadrature``) the points themselves matter
for agreement with the reference.  The orbit parameters below were re-derived, not copied:
``tools/make_triangle_rules.py`` solves the moment equations for the published orbit structures
(S3 = centroid, S21(a) = permutations of (a, a, 1-2a), S111(a, b) = permutations of (a, b, 1-a-b));
``tests/test_mt.py`` checks exactness on every monomial up to the degree and the agreement with the
reference's points (as a set) to 1e-12 (the reference prints ~15 digits).  The 33-point rule of degree
12 is the exception: random starts did not converge for it and its moment system is ill conditioned, so
the published 15-digit values are used as they are (moment residual 2e-15).

Weights sum to 1/2 (the area of the unit triangle), as in the reference.
"""
from __future__ import annotations

import numpy as np

# degree -> list of orbits: ("s3", w) | ("s21", w, a) | ("s111", w, a, b)
_ORBITS = {
    2: [("s21", 1.0 / 6.0, 1.0 / 6.0)],
    4: [("s21", 0.054975871827660873, 0.09157621350977066),
        ("s21", 0.11169079483900581, 0.44594849091596495)],
    6: [("s21", 0.025422453185101706, 0.063089014491499909),
        ("s21", 0.058393137863180428, 0.24928674517092173),
        ("s111", 0.041425537809192274, 0.0531450498448248, 0.31035245103377535)],
    8: [("s3", 0.072157803838942242),
        ("s21", 0.04754581713360953, 0.45929258829279074),
        ("s21", 0.016229248811595157, 0.050547228317029569),
        ("s21", 0.051608685267356312, 0.17056930775183615),
        ("s111", 0.013615157087229119, 0.2631128296344189, 0.0083947774100495039)],
    10: [("s3", 0.045408995191662985),
         ("s21", 0.018362978878201666, 0.48557763338372001),
         ("s21", 0.022660529717727262, 0.10948157548484638),
         ("s111", 0.014163621265468333, 0.24667256063955986, 0.72832390459791163),
         ("s111", 0.03637895842277928, 0.55035294182147942, 0.14170721941450812),
         ("s111", 0.0047108334818440926, 0.0095408154002988143, 0.92365593358769249)],
    # 33 points: the published 15-digit values (weights for the unit-area normalisation halved); the moment
    # system of this rule is so ill conditioned that re-solving it moves the points by 1e-10 while the
    # residual only drops from 2e-15 to 3e-17, so the published digits are kept
    12: [("s21", 0.025731066440455 / 2, 0.488217389773805),
         ("s21", 0.043692544538038 / 2, 0.439724392294460),
         ("s21", 0.062858224217885 / 2, 0.271210385012116),
         ("s21", 0.034796112930709 / 2, 0.127576145541586),
         ("s21", 0.006166261051559 / 2, 0.021317350453210),
         ("s111", 0.040371557766381 / 2, 0.115343494534698, 0.275713269685514),
         ("s111", 0.022356773202303 / 2, 0.022838332222257, 0.281325580989940),
         ("s111", 0.017316231108659 / 2, 0.025734050548330, 0.116251915907597)],
}


def triangle_quadrature(degree: int):
    """Points (xi, eta) [ng, 2] and weights [ng] of the symmetric rule exact for polynomials of the
    given (even) degree, 2..12, on the triangle (0,0), (1,0), (0,1)."""
    if degree not in _ORBITS:
        raise ValueError("triangle_quadrature: degree %r not tabulated (%s)" % (degree, sorted(_ORBITS)))
    pts, wts = [], []
    for orbit in _ORBITS[degree]:
        kind, w = orbit[0], orbit[1]
        if kind == "s3":
            bary = [(1.0 / 3.0, 1.0 / 3.0, 1.0 / 3.0)]
        elif kind == "s21":
            a = orbit[2]
            c = 1.0 - 2.0 * a
            bary = [(a, a, c), (a, c, a), (c, a, a)]
        else:
            a, b = orbit[2], orbit[3]
            c = 1.0 - a - b
            bary = [(a, b, c), (a, c, b), (b, a, c), (b, c, a), (c, a, b), (c, b, a)]
        for q in bary:
            pts.append((q[0], q[1]))
            wts.append(w)
    return np.array(pts, dtype=np.float64), np.array(wts, dtype=np.float64)
