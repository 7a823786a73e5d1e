"""Generates the correctly-rounded-in-practice double atan / atan2 / acos used by the
vanishing-point stage on BOTH sides (oracle/orc_atan.h in C, vplines-slam_b200/csrc/vpl_atan.cuh
in CUDA), next to the sincos of tools/gen_sincos.py whose double-double helpers it reuses.

Why it exists: the vanishing-point vote adds sqrt(l_i l_j) (sin(2 |o_i - o_j|) + 0.2) into sphere
cells in a fixed order, o = atan2(dy, dx) per line, and the hypotheses are built from
atan / sin / cos (feature_tracker/src/vanishing_point_detection.cpp:67-177, :180-276).  Equivalent
hypotheses (lambda and lambda + 90 degrees) read the same three cells in a different order, so which
of them wins can hang on the last bit of a cell.  libm and the CUDA math library round these
functions differently in ~0.1 % of the calls; with one shared definition the device and the oracle
produce the same bits.

Method: octant reduction to q = min/max in [0, 1] (double-double division), q = c + ..., c = k/64,
t = (q - c) / (1 + q c), atan q = atan c (table, double-double) + t (1 - t^2/3 + ... - t^18/19)
in double-double (the terms from t^8 on in plain double), then the octant/quadrant reflections with a double-double pi.  acos x =
atan2(sqrt((1 - x)(1 + x)), x) with the square root in double-double.  Only IEEE add/mul/fma/div/sqrt.

    python tools/gen_atan.py      # rewrites the two headers
    python tools/gen_atan.py --check N   # compares N random arguments with mpmath through the oracle
"""
import os
import sys

import mpmath as mp

mp.mp.prec = 400
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def split(x):
    hi = float(x)
    lo = float(x - mp.mpf(hi))
    return hi, lo


def main():
    tab = [split(mp.atan(mp.mpf(k) / 64)) for k in range(65)]
    coef = [split(mp.mpf((-1) ** n) / (2 * n + 1)) for n in range(10)]  # 1, -1/3, ... -1/19
    pi_h, pi_l = split(mp.pi)
    pio2_h, pio2_l = split(mp.pi / 2)

    def arr(name, c):
        return ("static VPL_SC_CONST double %s[%d][2] = {\n" % (name, len(c)) +
                ",\n".join("  {%s, %s}" % (float.hex(h), float.hex(l)) for h, l in c) + "};\n")

    body = f"""
#define VPL_AT_PI_H {float.hex(pi_h)}
#define VPL_AT_PI_L {float.hex(pi_l)}
#define VPL_AT_PIO2_H {float.hex(pio2_h)}
#define VPL_AT_PIO2_L {float.hex(pio2_l)}
{arr('vpl_at_tab', tab)}{arr('vpl_at_coef', coef)}
VPL_SC_FN vpl_dd vpl_dd_make(double h, double l) {{ vpl_dd r; r.hi = h; r.lo = l; return r; }}
VPL_SC_FN vpl_dd vpl_dd_neg(vpl_dd a) {{ a.hi = -a.hi; a.lo = -a.lo; return a; }}
/* a / b, ~2^-100 relative: three quotient digits from ONE reciprocal (each residual is formed exactly) */
VPL_SC_FN vpl_dd vpl_dd_div(vpl_dd a, vpl_dd b) {{
  double rb = 1.0 / b.hi;
  double q1 = a.hi * rb;
  vpl_dd r = vpl_dd_add(a, vpl_dd_neg(vpl_dd_mul(b, vpl_dd_make(q1, 0.0))));
  double q2 = r.hi * rb;
  r = vpl_dd_add(r, vpl_dd_neg(vpl_dd_mul(b, vpl_dd_make(q2, 0.0))));
  double q3 = r.hi * rb;
  vpl_dd q = vpl_dd_quick(q1, q2);
  return vpl_dd_add(q, vpl_dd_make(q3, 0.0));
}}
/* sqrt(a), a >= 0 */
VPL_SC_FN vpl_dd vpl_dd_sqrt(vpl_dd a) {{
  if (a.hi <= 0.0) return vpl_dd_make(0.0, 0.0);
  double x = sqrt(a.hi);
  double p = x * x;
  double e = fma(x, x, -p);
  vpl_dd r = vpl_dd_add(a, vpl_dd_make(-p, -e));
  return vpl_dd_quick(x, r.hi / (2.0 * x));
}}
/* atan(q), q in [0, 1] as a double-double */
VPL_AT_CORE vpl_dd vpl_atan_dd01(vpl_dd q) {{
  double kd = rint(q.hi * 64.0);
  int k = (int)kd;
  double c = kd * 0.015625;
  vpl_dd num = vpl_dd_add(q, vpl_dd_make(-c, 0.0));
  vpl_dd den = vpl_dd_add(vpl_dd_make(1.0, 0.0), vpl_dd_mul(q, vpl_dd_make(c, 0.0)));
  vpl_dd t = vpl_dd_div(num, den);
  vpl_dd u = vpl_dd_mul(t, t);
  /* u <= 2^-14: the terms from u^4 on are below 2^-56 of the sum, a plain double carries them */
  double tail = vpl_at_coef[9][0];
  for (int i = 8; i >= 4; --i) tail = vpl_at_coef[i][0] + u.hi * tail;
  vpl_dd p = vpl_dd_make(tail, 0.0);
  VPL_SC_ROLLED
  for (int i = 3; i >= 0; --i)
    p = vpl_dd_add_ord(vpl_dd_make(vpl_at_coef[i][0], vpl_at_coef[i][1]), vpl_dd_mul(u, p));
  return vpl_dd_add(vpl_dd_make(vpl_at_tab[k][0], vpl_at_tab[k][1]), vpl_dd_mul(t, p));
}}
/* atan2(y, x) of double-double arguments, as a double-double in (-pi, pi] */
VPL_SC_FN vpl_dd vpl_atan2_dd(vpl_dd y, vpl_dd x) {{
  int yneg = y.hi < 0.0, xneg = x.hi < 0.0;
  vpl_dd a = yneg ? vpl_dd_neg(y) : y;
  vpl_dd b = xneg ? vpl_dd_neg(x) : x;
  int swap = a.hi > b.hi || (a.hi == b.hi && a.lo > b.lo);
  vpl_dd lo = swap ? b : a, hi = swap ? a : b;
  vpl_dd q;
  if (hi.hi == 0.0 || (hi.hi - hi.hi != 0.0 && lo.hi - lo.hi == 0.0)) q = vpl_dd_make(0.0, 0.0); /* 0/0, finite/inf */
  else if (hi.hi - hi.hi != 0.0) q = vpl_dd_make(1.0, 0.0);                                       /* inf/inf */
  else q = vpl_dd_div(lo, hi);
  vpl_dd r = vpl_atan_dd01(q);
  if (swap) r = vpl_dd_add(vpl_dd_make(VPL_AT_PIO2_H, VPL_AT_PIO2_L), vpl_dd_neg(r));
  if (xneg) r = vpl_dd_add(vpl_dd_make(VPL_AT_PI_H, VPL_AT_PI_L), vpl_dd_neg(r));
  return yneg ? vpl_dd_neg(r) : r;
}}
/* std::atan2(double, double), correctly rounded in practice (signed zeros as IEEE 754 / C99 F.9.1.4) */
VPL_SC_FN double vpl_atan2_cr(double y, double x) {{
  if (y != y || x != x) return y + x;
  if (y == 0.0) {{
    double z = signbit(x) ? VPL_AT_PI_H : 0.0;
    return signbit(y) ? -z : z;
  }}
  vpl_dd r = vpl_atan2_dd(vpl_dd_make(y, 0.0), vpl_dd_make(x, 0.0));
  return r.hi + r.lo;
}}
/* std::atan(double): atan2(t, 1) without the division by one */
VPL_SC_FN double vpl_atan_cr(double t) {{
  if (t != t) return t;
  if (t == 0.0) return t;
  double a = t < 0.0 ? -t : t;
  vpl_dd r;
  if (a <= 1.0) {{
    r = vpl_atan_dd01(vpl_dd_make(a, 0.0));
  }} else {{
    vpl_dd q = a - a != 0.0 ? vpl_dd_make(0.0, 0.0) : vpl_dd_div(vpl_dd_make(1.0, 0.0), vpl_dd_make(a, 0.0));
    r = vpl_dd_add(vpl_dd_make(VPL_AT_PIO2_H, VPL_AT_PIO2_L), vpl_dd_neg(vpl_atan_dd01(q)));
  }}
  double v = r.hi + r.lo;
  return t < 0.0 ? -v : v;
}}
/* std::acos(double): NaN outside [-1, 1] */
VPL_SC_FN double vpl_acos_cr(double x) {{
  if (!(x >= -1.0 && x <= 1.0)) return x - x != 0.0 ? x - x : (x - x) / (x - x);
  if (x == 1.0) return 0.0;
  vpl_dd s = vpl_dd_sqrt(vpl_dd_mul(vpl_dd_two_sum(1.0, -x), vpl_dd_two_sum(1.0, x)));
  vpl_dd r = vpl_atan2_dd(s, vpl_dd_make(x, 0.0));
  return r.hi + r.lo;
}}
"""
    c_head = ("/* GENERATED by tools/gen_atan.py -- do not edit.  Deterministic double atan / atan2 / acos\n"
              " * (table + double-double Taylor; IEEE add/mul/fma/div/sqrt only). */\n")
    with open(os.path.join(ROOT, "oracle", "orc_atan.h"), "w") as f:
        f.write(c_head + "#ifndef ORC_ATAN_H\n#define ORC_ATAN_H\n#include \"orc_sincos.h\"\n#define VPL_AT_CORE static inline\n" + body + "#endif\n")
    with open(os.path.join(ROOT, "vplines-slam_b200", "csrc", "vpl_atan.cuh"), "w") as f:
        f.write(c_head + "#pragma once\n#include \"vpl_sincos.cuh\"\n"
                "/* one out-of-line copy of the table + series per kernel image: atan, atan2 and acos all end in it */\n"
                "#define VPL_AT_CORE static __device__ __noinline__\nnamespace vpl {\n" + body + "}  // namespace vpl\n")
    print("wrote oracle/orc_atan.h and vplines-slam_b200/csrc/vpl_atan.cuh")


def check(n):
    """Compares the oracle's build of these functions with mpmath (round to nearest)."""
    import ctypes
    import random
    sys.path.insert(0, ROOT)
    from oracle import oracle as orc
    L = orc.lib()
    for name in ("orc_atan2_cr", "orc_atan_cr", "orc_acos_cr"):
        getattr(L, name).restype = ctypes.c_double
    L.orc_atan2_cr.argtypes = [ctypes.c_double, ctypes.c_double]
    L.orc_atan_cr.argtypes = [ctypes.c_double]
    L.orc_acos_cr.argtypes = [ctypes.c_double]
    rnd = random.Random(7)
    bad = [0, 0, 0]
    for i in range(n):
        sc = 10.0 ** rnd.uniform(-6, 6)
        y, x = rnd.uniform(-1, 1) * sc, rnd.uniform(-1, 1) * 10.0 ** rnd.uniform(-6, 6)
        if float(mp.atan2(mp.mpf(y), mp.mpf(x))) != L.orc_atan2_cr(y, x):
            bad[0] += 1
        t = rnd.uniform(-1, 1) * sc
        if float(mp.atan(mp.mpf(t))) != L.orc_atan_cr(t):
            bad[1] += 1
        u = rnd.uniform(-1, 1) if i % 4 else 1.0 - 10.0 ** rnd.uniform(-16, -1)
        if float(mp.acos(mp.mpf(u))) != L.orc_acos_cr(u):
            bad[2] += 1
    print("misrounded of %d: atan2 %d, atan %d, acos %d" % (n, bad[0], bad[1], bad[2]))
    return bad


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--check":
        check(int(sys.argv[2]))
    else:
        main()
