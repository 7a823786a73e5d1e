"""Generates the correctly-rounded-in-practice double sincos used by LSD's region2rect
on BOTH sides (oracle/orc_sincos.h in C, vplines-slam_b200/csrc/vpl_sincos.cuh in CUDA).

Why it exists: a rectangle's edges pass through the centres of its extreme pixels, so
rect_nfa's ceil()/trunc() row limits sit within an ulp of integers and a 1-ulp change in
dx = cos(theta), dy = sin(theta) moves pixels in or out of the rectangle.  libm and the
CUDA math library both differ from the correctly rounded value by up to an ulp, in
different places.  Both sides therefore use this one definition: a 3-term Cody-Waite reduction, a second
reduction by a 27-entry double-double table of sin/cos(m/32) and short double-double Taylor series, built only from IEEE add/mul/fma, which
yields the correctly rounded result unless the true value lies within ~1e-22 (relative)
of a rounding midpoint.  It agrees with glibc wherever glibc is itself correctly rounded.

    python tools/gen_sincos.py      # rewrites the two headers
"""
import os
import mpmath as mp

mp.mp.prec = 400
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def split(x):
    hi = float(x)
    lo = float(x - mp.mpf(hi))
    return hi, lo


def trunc_bits(x, bits):
    """x truncated to `bits` significant bits (so that small-integer multiples are exact)."""
    m, e = mp.frexp(x)
    m = mp.floor(m * 2 ** bits) / 2 ** bits
    return mp.ldexp(m, e)


def main():
    pio2 = mp.pi / 2
    p1 = trunc_bits(pio2, 40)
    p2 = trunc_bits(pio2 - p1, 40)
    p3 = pio2 - p1 - p2
    P1, P2 = float(p1), float(p2)
    assert mp.mpf(P1) == p1 and mp.mpf(P2) == p2
    P3h, P3l = split(p3)
    two_over_pi = float(2 / mp.pi)
    sin_c = [split(mp.mpf((-1) ** k) / mp.factorial(2 * k + 1)) for k in range(1, 13)]  # r^3 .. r^25
    cos_c = [split(mp.mpf((-1) ** k) / mp.factorial(2 * k)) for k in range(1, 13)]      # r^2 .. r^24

    tab = []
    for m in range(27):
        sh, sl = split(mp.sin(mp.mpf(m) / 32))
        ch, cl = split(mp.cos(mp.mpf(m) / 32))
        tab.append((sh, sl, ch, cl))
    tab_txt = ("static VPL_SC_CONST double vpl_sc_tab[27][4] = {\n" +
               ",\n".join("  {%s, %s, %s, %s}" % tuple(float.hex(v) for v in t) for t in tab) + "};\n")

    def arr(name, c):
        return ("static VPL_SC_CONST double %s[%d][2] = {\n" % (name, len(c)) +
                ",\n".join("  {%s, %s}" % (float.hex(h), float.hex(l)) for h, l in c) + "};\n")

    body = f"""
/* pi/2 = P1 + P2 + P3 (P1, P2: 40 significant bits each, so k*P1 and k*P2 are exact) */
#define VPL_SC_P1 {float.hex(P1)}
#define VPL_SC_P2 {float.hex(P2)}
#define VPL_SC_P3H {float.hex(P3h)}
#define VPL_SC_P3L {float.hex(P3l)}
#define VPL_SC_2OPI {float.hex(two_over_pi)}
{arr('vpl_sc_sin', sin_c)}{arr('vpl_sc_cos', cos_c)}/* sin(m/32), cos(m/32), m = 0..26, as double-doubles */
{tab_txt}
typedef struct {{ double hi, lo; }} vpl_dd;

VPL_SC_FN vpl_dd vpl_dd_two_sum(double a, double b) {{
  vpl_dd r; r.hi = a + b; double bb = r.hi - a; r.lo = (a - (r.hi - bb)) + (b - bb); return r; }}
VPL_SC_FN vpl_dd vpl_dd_quick(double a, double b) {{
  vpl_dd r; r.hi = a + b; r.lo = b - (r.hi - a); return r; }}
VPL_SC_FN vpl_dd vpl_dd_add(vpl_dd a, vpl_dd b) {{
  vpl_dd s = vpl_dd_two_sum(a.hi, b.hi); double e = s.lo + (a.lo + b.lo); return vpl_dd_quick(s.hi, e); }}
VPL_SC_FN vpl_dd vpl_dd_mul(vpl_dd a, vpl_dd b) {{
  double p = a.hi * b.hi; double e = fma(a.hi, b.hi, -p); e = e + (a.hi * b.lo + a.lo * b.hi); return vpl_dd_quick(p, e); }}

/* a + b for |a.hi| >= |b.hi| (a Horner coefficient and the smaller product): the exact sum of the high parts needs
 * no branch-free two_sum; same result, three operations fewer */
VPL_SC_FN vpl_dd vpl_dd_add_ord(vpl_dd a, vpl_dd b) {{
  vpl_dd s = vpl_dd_quick(a.hi, b.hi); double e = s.lo + (a.lo + b.lo); return vpl_dd_quick(s.hi, e); }}

/* sin r and cos r for a double-double |r| <= pi/4 (+ rounding): r = c + d, c = m/32 from a table of double-double
 * sin c / cos c, |d| <= 1/64: sin d and cos d from three double-double Taylor steps (the terms from d^9 / d^8 on
 * are below 2^-63 of the sums, a plain double carries them), then the angle-addition formulas. */
VPL_SC_FN void vpl_sincos_dd(vpl_dd r, vpl_dd* s_out, vpl_dd* c_out, int want) {{
  double md = rint(r.hi * 32.0);
  int m = (int)md;
  int a = m < 0 ? -m : m;
  vpl_dd mc; mc.hi = -(md * 0.03125); mc.lo = 0.0;
  vpl_dd d = vpl_dd_add(r, mc);
  vpl_dd d2 = vpl_dd_mul(d, d);
  double ts = vpl_sc_sin[6][0], tc = vpl_sc_cos[6][0];
  VPL_SC_ROLLED
  for (int i = 5; i >= 3; --i) {{
    ts = vpl_sc_sin[i][0] + d2.hi * ts;
    tc = vpl_sc_cos[i][0] + d2.hi * tc;
  }}
  vpl_dd ps, pc;
  ps.hi = ts; ps.lo = 0.0;
  pc.hi = tc; pc.lo = 0.0;
  VPL_SC_ROLLED
  for (int i = 2; i >= 0; --i) {{
    vpl_dd cs; cs.hi = vpl_sc_sin[i][0]; cs.lo = vpl_sc_sin[i][1];
    vpl_dd cc; cc.hi = vpl_sc_cos[i][0]; cc.lo = vpl_sc_cos[i][1];
    ps = vpl_dd_add_ord(cs, vpl_dd_mul(d2, ps));
    pc = vpl_dd_add_ord(cc, vpl_dd_mul(d2, pc));
  }}
  /* sin d = d + d*d2*ps ; cos d = 1 + d2*pc */
  vpl_dd sd = vpl_dd_add(d, vpl_dd_mul(d, vpl_dd_mul(d2, ps)));
  vpl_dd one; one.hi = 1.0; one.lo = 0.0;
  vpl_dd cd = vpl_dd_add_ord(one, vpl_dd_mul(d2, pc));
  vpl_dd S, C;
  S.hi = vpl_sc_tab[a][0]; S.lo = vpl_sc_tab[a][1];
  C.hi = vpl_sc_tab[a][2]; C.lo = vpl_sc_tab[a][3];
  if (m < 0) {{ S.hi = -S.hi; S.lo = -S.lo; }}
  if (want & 1) *s_out = vpl_dd_add(vpl_dd_mul(S, cd), vpl_dd_mul(C, sd));
  if (want & 2) {{
    vpl_dd t = vpl_dd_mul(S, sd); t.hi = -t.hi; t.lo = -t.lo;
    *c_out = vpl_dd_add(vpl_dd_mul(C, cd), t);
  }}
}}

/* theta - k pi/2 as a double-double, k = rint(theta 2/pi) */
VPL_SC_FN vpl_dd vpl_sc_reduce(double theta, int* k_out) {{
  double kd = rint(theta * VPL_SC_2OPI);
  *k_out = (int)kd;
  double t1 = theta - kd * VPL_SC_P1;              /* exact */
  vpl_dd r = vpl_dd_two_sum(t1, -(kd * VPL_SC_P2)); /* kd*P2 exact */
  vpl_dd p3; p3.hi = kd * VPL_SC_P3H; p3.lo = fma(kd, VPL_SC_P3H, -p3.hi) + kd * VPL_SC_P3L;
  p3.hi = -p3.hi; p3.lo = -p3.lo;
  return vpl_dd_add(r, p3);
}}

/* sin and cos of theta (|theta| < 1e4), correctly rounded in practice. */
VPL_SC_FN void vpl_sincos_cr(double theta, double* s_out, double* c_out) {{
  int k;
  vpl_dd r = vpl_sc_reduce(theta, &k);
  vpl_dd sr, cr;
  vpl_sincos_dd(r, &sr, &cr, 3);
  double sv = sr.hi + sr.lo, cv = cr.hi + cr.lo;
  switch (k & 3) {{
    case 0: *s_out = sv; *c_out = cv; break;
    case 1: *s_out = cv; *c_out = -sv; break;
    case 2: *s_out = -sv; *c_out = -cv; break;
    default: *s_out = -cv; *c_out = sv; break;
  }}
}}

/* sin(theta) alone (bit-identical to the sine of vpl_sincos_cr): only the combination the quadrant needs */
VPL_SC_FN double vpl_sin_cr(double theta) {{
  int k;
  vpl_dd r = vpl_sc_reduce(theta, &k);
  vpl_dd sr, cr;
  double v;
  if (k & 1) {{ vpl_sincos_dd(r, &sr, &cr, 2); v = cr.hi + cr.lo; }}
  else {{ vpl_sincos_dd(r, &sr, &cr, 1); v = sr.hi + sr.lo; }}
  return (k & 2) ? -v : v;
}}
"""
    head = ("/* GENERATED by tools/gen_sincos.py -- do not edit.  Deterministic double sincos\n"
            " * (double-double Taylor after Cody-Waite reduction; IEEE add/mul/fma only). */\n")
    with open(os.path.join(ROOT, "oracle", "orc_sincos.h"), "w") as f:
        f.write(head + "#ifndef ORC_SINCOS_H\n#define ORC_SINCOS_H\n#include <math.h>\n"
                "#define VPL_SC_CONST const\n#define VPL_SC_FN static inline\n#define VPL_SC_ROLLED\n" + body + "#endif\n")
    with open(os.path.join(ROOT, "vplines-slam_b200", "csrc", "vpl_sincos.cuh"), "w") as f:
        f.write(head + "#pragma once\n#define VPL_SC_CONST __device__ const\n"
                "#define VPL_SC_FN __device__ __forceinline__\n"
                "#define VPL_SC_ROLLED _Pragma(\"unroll 1\") /* keeps the kernels that call this small enough for the instruction cache */\n"
                "namespace vpl {\n" + body + "}  // namespace vpl\n")
    print("written")


def check(n):
    """Compares the oracle's build of these functions with mpmath (round to nearest)."""
    import ctypes
    import random
    import sys
    sys.path.insert(0, ROOT)
    from oracle import oracle as orc
    L = orc.lib()
    L.orc_sin_cr.restype = ctypes.c_double; L.orc_sin_cr.argtypes = [ctypes.c_double]
    L.orc_sincos_cr.argtypes = [ctypes.c_double, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    rnd = random.Random(11)
    bad = [0, 0, 0]
    s, c = ctypes.c_double(), ctypes.c_double()
    for i in range(n):
        t = rnd.uniform(-7, 7) if i % 3 else rnd.uniform(-1, 1) * 10.0 ** rnd.uniform(-8, 3)
        L.orc_sincos_cr(t, ctypes.byref(s), ctypes.byref(c))
        es, ec = float(mp.sin(mp.mpf(t))), float(mp.cos(mp.mpf(t)))
        bad[0] += s.value != es
        bad[1] += c.value != ec
        bad[2] += L.orc_sin_cr(t) != s.value
    print("misrounded of %d: sin %d, cos %d; vpl_sin_cr != sine of vpl_sincos_cr: %d" % (n, bad[0], bad[1], bad[2]))
    return bad


if __name__ == "__main__":
    import sys
    if len(sys.argv) > 2 and sys.argv[1] == "--check":
        check(int(sys.argv[2]))
    else:
        main()
