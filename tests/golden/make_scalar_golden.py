"""Generates tests/golden/scalar_cases.json: the known-answer vectors of the reference's Scalar tests.

Every case names the reference test it transcribes (file:line in /root/reference/tests).  Expected values are
the closed forms written in those tests (math.* here instead of std::*); where the reference compares against
Maple-generated derivative expressions of a formula, the same formula is differentiated here with sympy.
`null` = entry not asserted by the reference.  Run:  python tests/golden/make_scalar_golden.py
"""
import json
import math
import os

import sympy as sp

cases = []


def case(name, params, k, expected, tol, ref, n_out=1):
    """expected: list (one per returned scalar) of dicts val / grad / hess (lists, None = not asserted)."""
    cases.append({"name": name, "params": [float(p) for p in params], "k": k, "n_out": n_out, "expected": expected, "tol": tol, "ref": ref})


def e1(val, g, h):
    return [{"val": val, "grad": [g], "hess": [[h]]}]


U = "ScalarTestUnaryOperators.cc"
B = "ScalarTestBinaryOperators.cc"
a = (4.0, 3.0, 2.0)      # a(x) = x^2 + x + 2 at x = 1
s4, c4 = math.sin(4.0), math.cos(4.0)
case("neg", a, 1, e1(-4.0, -3.0, -2.0), 0.0, f"{U}:11-26")
case("sqrt", a, 1, e1(2.0, 3.0 / 4.0, 7.0 / 32.0), 1e-12, f"{U}:39-53")
case("fabs", (1.0, 3.0, 6.0), 1, e1(1.0, 3.0, 6.0), 1e-12, f"{U}:117-127")
case("fabs", (-1.0, 3.0, -6.0), 1, e1(1.0, -3.0, 6.0), 1e-12, f"{U}:129-139")
case("abs", (1.0, 3.0, 6.0), 1, e1(1.0, 3.0, 6.0), 1e-12, f"{U}:158-168")
case("abs", (-1.0, 3.0, -6.0), 1, e1(1.0, -3.0, 6.0), 1e-12, f"{U}:170-180")
case("exp", a, 1, e1(math.exp(4.0), 3.0 * math.exp(4.0), 11.0 * math.exp(4.0)), 1e-12 * math.exp(4.0) * 11, f"{U}:194-208")
case("log", a, 1, e1(2.0 * math.log(2.0), 3.0 / 4.0, -1.0 / 16.0), 1e-12, f"{U}:221-235")
case("log2", a, 1, e1(2.0, 3.0 / 4.0 / math.log(2.0), -1.0 / 16.0 / math.log(2.0)), 1e-12, f"{U}:248-262")
case("log10", a, 1, e1(math.log10(4.0), 3.0 / 4.0 / math.log(10.0), -1.0 / 16.0 / math.log(10.0)), 1e-12, f"{U}:275-289")
case("sin", a, 1, e1(s4, 3.0 * c4, 2.0 * c4 - 9.0 * s4), 1e-12, f"{U}:302-316")
case("cos", a, 1, e1(c4, -3.0 * s4, -2.0 * s4 - 9.0 * c4), 1e-12, f"{U}:329-343")
case("tan", a, 1, e1(math.tan(4.0), 3.0 / c4 ** 2, 4.0 * (1.0 + 9.0 * math.tan(4.0)) / (1.0 + math.cos(8.0))), 1e-12, f"{U}:356-370")
h = (0.5, 3.0, 2.0)
case("asin", h, 1, e1(math.asin(0.5), 3.4641, 9.2376), 1e-4, f"{U}:383-397")
case("acos", h, 1, e1(math.acos(0.5), -3.4641, -9.2376), 1e-4, f"{U}:410-424")
case("atan", h, 1, e1(math.atan(0.5), 2.4, -4.16), 1e-12, f"{U}:437-451")
case("sinh", a, 1, e1(math.sinh(4.0), 3.0 * math.cosh(4.0), 9.0 * math.sinh(4.0) + 2.0 * math.cosh(4.0)), 1e-12 * 300, f"{U}:464-478")
case("cosh", a, 1, e1(math.cosh(4.0), 3.0 * math.sinh(4.0), 2.0 * math.sinh(4.0) + 9.0 * math.cosh(4.0)), 1e-12 * 300, f"{U}:491-505")
case("tanh", a, 1, e1(math.tanh(4.0), 3.0 / math.cosh(4.0) ** 2,
                      2.0 * (1.0 - 9.0 * math.sinh(4.0) / math.cosh(4.0)) / math.cosh(4.0) ** 2), 1e-12, f"{U}:518-532")
case("asinh", h, 1, e1(math.asinh(0.5), 2.68328, -1.43108), 1e-5, f"{U}:545-559")
case("acosh", a, 1, e1(math.acosh(4.0), math.sqrt(3.0 / 5.0), -2.0 / 5.0 / math.sqrt(15.0)), 1e-12, f"{U}:572-586")
case("atanh", h, 1, e1(math.atanh(0.5), 4.0, 18.6667), 1e-4, f"{U}:599-613")
# pow
case("pow_int", a + (0,), 1, e1(1.0, 0.0, 0.0), 1e-12, f"{B}:19-28")
case("pow_int", a + (1,), 1, e1(4.0, 3.0, 2.0), 1e-12, f"{B}:29-38")
case("pow_int", a + (3,), 1, e1(64.0, 144.0, 312.0), 1e-12, f"{B}:39-48")
case("pow_real", a + (1.5,), 1, e1(8.0, 9.0, 75.0 / 8.0), 1e-12, f"{B}:69-78")
case("pow_real", a + (0.0,), 1, e1(1.0, 0.0, 0.0), 1e-12, f"{B}:79-88")
case("pow_real", a + (1.0,), 1, e1(4.0, 3.0, 2.0), 1e-12, f"{B}:89-98")
case("pow_real", a + (3.0,), 1, e1(64.0, 144.0, 312.0), 1e-12, f"{B}:99-108")
# binary: a = (4,3,2), b = (0,1,4)  (x^3 - x^2 at x = 1)
ab = a + (0.0, 1.0, 4.0)
case("add", ab, 1, e1(4.0, 4.0, 6.0), 1e-12, f"{B}:131-140")
case("add_s", ab + (1.0,), 1, e1(5.0, 3.0, 2.0), 1e-12, f"{B}:142-151")
case("s_add", ab + (1.0,), 1, e1(5.0, 3.0, 2.0), 1e-12, f"{B}:153-162")
case("iadd", ab, 1, e1(4.0, 4.0, 6.0), 1e-12, f"{B}:164-173")
case("iadd_s", (4.0, 4.0, 6.0, 0, 0, 0, 1.0), 1, e1(5.0, 4.0, 6.0), 1e-12, f"{B}:175-184")
case("sub", ab, 1, e1(4.0, 2.0, -2.0), 1e-12, f"{B}:207-216")
case("sub_s", ab + (1.0,), 1, e1(3.0, 3.0, 2.0), 1e-12, f"{B}:218-227")
case("s_sub", ab + (1.0,), 1, e1(-3.0, -3.0, -2.0), 1e-12, f"{B}:229-238")
case("isub", ab, 1, e1(4.0, 2.0, -2.0), 1e-12, f"{B}:240-249")
case("isub_s", (4.0, 2.0, -2.0, 0, 0, 0, 1.0), 1, e1(3.0, 2.0, -2.0), 1e-12, f"{B}:251-260")
case("mul", ab, 1, e1(0.0, 4.0, 22.0), 1e-12, f"{B}:283-292")
case("mul_s", ab + (2.0,), 1, e1(8.0, 6.0, 4.0), 1e-12, f"{B}:294-303")
case("s_mul", ab + (2.0,), 1, e1(8.0, 6.0, 4.0), 1e-12, f"{B}:305-314")
case("imul", ab, 1, e1(0.0, 4.0, 22.0), 1e-12, f"{B}:316-325")
case("imul_s", (0.0, 4.0, 22.0, 0, 0, 0, 2.0), 1, e1(0.0, 8.0, 44.0), 1e-12, f"{B}:327-336")
# division: a = (1,1,4), b = (4,3,2)
dv = (1.0, 1.0, 4.0, 4.0, 3.0, 2.0)
case("div", dv, 1, e1(1.0 / 4.0, 1.0 / 16.0, 25.0 / 32.0), 1e-12, f"{B}:359-368")
case("div_s", dv + (2.0,), 1, e1(0.5, 0.5, 2.0), 1e-12, f"{B}:370-379")
case("s_div", dv + (2.0,), 1, e1(2.0, -2.0, -4.0), 1e-12, f"{B}:381-390")
case("idiv", dv, 1, e1(1.0 / 4.0, 1.0 / 16.0, 25.0 / 32.0), 1e-12, f"{B}:392-401")
case("idiv_s", (1.0 / 4.0, 1.0 / 16.0, 25.0 / 32.0, 0, 0, 0, 2.0), 1, e1(1.0 / 8.0, 1.0 / 32.0, 25.0 / 64.0), 1e-12, f"{B}:403-412")
case("quadratic", (1.0,), 1, e1(4.0, 3.0, 2.0), 0.0, "ScalarTestMisc.cc:11-26")

# ---- compound expressions: differentiate the reference's formula with sympy ----
x, y = sp.symbols("x y", real=True)


def sym_case(name, expr, pts, k, tol, ref, check_val=True):
    vars_ = [x, y][:k]
    grad = [sp.diff(expr, v) for v in vars_]
    hess = [[sp.diff(expr, v, w) for w in vars_] for v in vars_]
    for pt in pts:
        sub = dict(zip(vars_, pt))
        ev = lambda e: float(sp.N(e.subs(sub), 30))
        exp = [{"val": ev(expr) if check_val else None, "grad": [ev(g) for g in grad], "hess": [[ev(hh) for hh in row] for row in hess]}]
        case(name, pt, k, exp, tol, ref)


sym_case("atan2_1", sp.atan2(x ** 2 - x - 1, x), [(-2.0,), (-1.0,), (-0.5,), (-0.25,), (0.25,), (0.5,), (1.0,), (2.0,)], 1, 1e-12, f"{B}:468-510")
# atan2(y, x) with independent variables: closed forms of ScalarTestBinaryOperators.cc:426-456 (sympy gives the same)
sym_case("atan2_const", sp.atan2(y, x), [(1.0, 2.0), (2.0, 2.0), (-1.0, 2.0), (-2.0, 3.0), (1.0, 0.0), (0.0, 1.0), (-1.0, 0.0), (0.0, -1.0)], 2, 1e-12, f"{B}:426-456")
sym_case("hypot", sp.sqrt(x ** 2 + y ** 2), [(3.0, 4.0)], 2, 1e-12, f"{B}:590-611")
sym_case("div2d", x ** 2 / y, [(-1.0, -0.5)], 2, 1e-12, f"{B}:623-642")
aa = sp.Rational(1, 2) * x ** 2 - y ** 2 + 2 * x - y
bb = -(x - 2) ** 2 - (y - 3) ** 2 + 1
sym_case("div2d_2", aa / bb, [(5.0, 0.0), (1.0, 1.0), (0.0, 5.0), (-1.0, 1.0), (-5.0, 0.0), (-1.0, -1.0), (0.0, -5.0), (1.0, -1.0)], 2, 1e-12, f"{B}:655-700 (Maple)")
sym_case("plus_minus_mult_div_2d", (x ** 2 + x) * (y ** 2 - y) / (y - 1), [(1.0, 1.5)], 2, 1e-12, f"{B}:712-727")
# atan2_2: two outputs (atan2(b, a), atan(b / a)) with identical derivatives
a2 = sp.Rational(1, 2) * x ** 2 - y ** 2 - y
b2 = -(x - 2) ** 2 - (y - 3) ** 2 + 1
expr = sp.atan(b2 / a2)
for pt in [(1.0, 0.0), (0.5, 0.5), (0.0, 1.0), (-0.5, 0.5), (-1.0, 0.0), (-0.5, -0.5), (0.5, -0.5)]:
    sub = {x: pt[0], y: pt[1]}
    ev = lambda e: float(sp.N(e.subs(sub), 30))
    g = [ev(sp.diff(expr, v)) for v in (x, y)]
    hs = [[ev(sp.diff(expr, v, w)) for w in (x, y)] for v in (x, y)]
    one = {"val": None, "grad": g, "hess": hs}
    case("atan2_2", pt, 2, [one, one], 1e-12, f"{B}:523-570 (Maple)", n_out=2)
# sqr(a) == pow(a, 2) == a * a  (ScalarTestUnaryOperators.cc:66-101)
q = x * x + 7 * y * y - 9 * x + x + 2 * y
sq = q ** 2
sub = {x: 4.0, y: 6.0}
ev = lambda e: float(sp.N(e.subs(sub), 30))
one = {"val": ev(sq), "grad": [ev(sp.diff(sq, v)) for v in (x, y)], "hess": [[ev(sp.diff(sq, v, w)) for w in (x, y)] for v in (x, y)]}
case("sqr_pow_mul", (4.0, 6.0), 2, [one, one, one], 1e-12 * abs(one["hess"][1][1]), f"{U}:66-101", n_out=3)
# sphere parametrisation (ScalarTestMisc.cc:38-86)
al = be = math.pi / 8.0
sa, ca, sb, cb = math.sin(al), math.cos(al), math.sin(be), math.cos(be)
case("sphere", (al, be), 2, [
    {"val": sa * cb, "grad": [ca * cb, -sa * sb], "hess": [[-sa * cb, -ca * sb], [-ca * sb, -sa * cb]]},
    {"val": sa * sb, "grad": [ca * sb, cb * sa], "hess": [[-sa * sb, ca * cb], [ca * cb, -sa * sb]]},
    {"val": ca, "grad": [-sa, 0.0], "hess": [[-ca, 0.0], [0.0, 0.0]]}], 1e-12, "ScalarTestMisc.cc:38-86", n_out=3)


# ---- tests/ScalarTestComparison.cc ----
C = "ScalarTestComparison.cc"
NA = [{"val": None, "grad": [], "hess": []}]


def flag(v):
    return [{"val": float(v), "grad": [0.0], "hess": [[0.0]]}]


# isnan / isinf / isfinite of passive scalars (:12-32): bit 0 isnan, bit 1 isinf, bit 2 isfinite
for v, bits in ((0.0, 4), (math.inf, 2), (-math.inf, 2), (math.nan, 1)):
    case("isnan_isinf", (v,), 1, flag(bits), 0.0, f"{C}:12-32")


def cmp_mask(a, b, sc):
    """The 18 comparisons of test_comparison (:41-108) as bits, in the order ==, !=, <, <=, >, >= for (a, b), (a, double), (double, a).
    Comparisons look at val only (Scalar.hh:933-1095): a = (1,1,4) == b = (1,2,8)."""
    ops = [lambda p, q: p == q, lambda p, q: p != q, lambda p, q: p < q, lambda p, q: p <= q, lambda p, q: p > q, lambda p, q: p >= q]
    bits = [op(a, b) for op in ops] + [op(a, sc) for op in ops] + [op(sc, a) for op in ops]
    return sum(1 << i for i, v in enumerate(bits) if v)


ca, cb, cc = (1.0, 1.0, 4.0), (1.0, 2.0, 8.0), (2.0, 2.0, 8.0)
for p_, q_ in ((ca, cb), (cb, ca), (ca, cc), (cc, ca), (cb, cc), (cc, cb)):
    for sc in (1.0, 2.0):
        case("cmp", p_ + q_ + (sc,), 1, flag(cmp_mask(p_[0], q_[0], sc)), 0.0, f"{C}:41-108")
# spot values the reference spells out: a == b, !(a < b), a <= b, a >= b; a == 1.0, a < 2.0, !(a > 2.0)
assert cmp_mask(1.0, 1.0, 1.0) & 0b111111 == 0b101001 and cmp_mask(1.0, 2.0, 2.0) & 0b111111 == 0b001110
# min / fmin / max / fmax select the whole scalar (:110-139): a = (1,2,3), b = (2,3,4)
ma = (1.0, 2.0, 3.0, 2.0, 3.0, 4.0)
for nm in ("min", "fmin"):
    case(nm, ma, 1, e1(1.0, 2.0, 3.0), 0.0, f"{C}:110-139")
for nm in ("max", "fmax"):
    case(nm, ma, 1, e1(2.0, 3.0, 4.0), 0.0, f"{C}:110-139")
# clamp(x, lo, hi) with double bounds (:141-160): x = (4,3,2)
case("clamp_d", (4.0, 3.0, 2.0, 0.0, 5.0), 1, e1(4.0, 3.0, 2.0), 0.0, f"{C}:141-160")
case("clamp_d", (4.0, 3.0, 2.0, -5.0, 0.0), 1, e1(0.0, 0.0, 0.0), 0.0, f"{C}:141-160")
case("clamp_d", (4.0, 3.0, 2.0, 5.0, 10.0), 1, e1(5.0, 0.0, 0.0), 0.0, f"{C}:141-160")
# clamp with scalar bounds returns the bound with ITS derivatives (Scalar.hh:1133-1145)
case("clamp", (4.0, 3.0, 2.0, 5.0, 1.0, 7.0, 10.0, 0.5, 0.25), 1, e1(5.0, 1.0, 7.0), 0.0, "Scalar.hh:1133-1145")

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "scalar_cases.json")
json.dump(cases, open(out, "w"), indent=0)
print(len(cases), "cases ->", out)
