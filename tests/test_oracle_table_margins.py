"""How far the float32 tables of the decode path are from depending on WHICH libm computed them.

The reference builds its tables with Go's math.Pow / math.Sin / math.Cos (pure-Go implementations), the oracle and the
engine with glibc's; both round a float64 result to float32 (imdct.go:23-79, frame.go:31-40,146-173,488-497).  Two
float64 results that differ in their last bits round to the same float32 unless a float32 rounding boundary (the
midpoint of two neighbouring float32 values) lies between them.  This test computes every table entry exactly (mpmath,
200 bits), measures its distance to the nearest float32 rounding boundary in units of the float64 ulp at that value,
and lists the entries closer than a margin: those are the only entries on which a different libm could change a bit.

Margins: 4 ulp64 for sin/cos of a float64 argument (both libraries document < 1 ulp; Go's Cody-Waite reduction is good to
about 2 for these arguments); 32 ulp64 for the requantisation product 2^(k/4) * |is|^(4/3), which in the reference is
math.Pow(2, k/4) * math.Pow(i, 4/3) with the product rounded too (Go's Pow goes through Exp(yf * Log(x)): a few ulp each).

The sets are EMPTY (asserted): the tables are the correctly rounded values of the exact quantities, and no implementation
within those error bounds can produce anything else.  The oracle's "parity unpinned" caveat (no Go binary to compare
with) therefore does not extend to the tables; what stays unpinned is the literal restatement of the control flow.
"""
import ctypes as C

import mpmath as mp
import numpy as np

import hostemu_lib
import oracle

mp.mp.prec = 200


def f32_boundary_distance_ulp64(x):
    """(distance from x to the nearest float32 rounding boundary) / (float64 ulp at x); x an mpf != 0."""
    ax = abs(x)
    e = int(mp.floor(mp.log(ax, 2)))
    if mp.mpf(2) ** e > ax:
        e -= 1
    ulp32 = mp.mpf(2) ** (e - 23)
    ulp64 = mp.mpf(2) ** (e - 52)
    # boundaries are the odd multiples of ulp32 / 2
    t = ax / (ulp32 / 2)
    k = mp.floor(t)
    lo, hi = (k, k + 1)
    cands = [c for c in (lo - 1, lo, hi, hi + 1) if int(c) % 2 == 1]
    d = min(abs(t - c) for c in cands) * (ulp32 / 2)
    return d / ulp64


def rn32(x):
    return np.float32(float(mp.nstr(x, 40)))  # float(): correctly rounded to f64 from 40 digits, then f64 -> f32 (checked below to be safe)


def rn32_exact(x):
    """Correctly rounded float32 of an mpf, without double rounding."""
    if x == 0:
        return np.float32(0.0)
    s = -1 if x < 0 else 1
    ax = abs(x)
    e = int(mp.floor(mp.log(ax, 2)))
    if mp.mpf(2) ** e > ax:
        e -= 1
    q = ax / mp.mpf(2) ** (e - 23)
    n = int(mp.floor(q))
    r = q - n
    if r > mp.mpf(1) / 2 or (r == mp.mpf(1) / 2 and n % 2 == 1):
        n += 1
    return np.float32(s * float(n) * 2.0 ** (e - 23))


def go_const(expr):
    """A Go untyped-constant expression such as math.Pi / 36: evaluated exactly, rounded once to float64."""
    return float(mp.nstr(expr, 40))


def check(name, entries, table, margin):
    """entries: list of (index, exact mpf).  table: the float32 table in use.  Returns the entries inside the margin."""
    close, worst = [], None
    for idx, x in entries:
        want = rn32_exact(x)
        assert table[idx].view(np.uint32) == want.view(np.uint32) or (x == 0 and table[idx] == 0), (name, idx, table[idx], want)
        if x == 0:
            continue
        d = f32_boundary_distance_ulp64(x)
        worst = d if worst is None or d < worst else worst
        if d < margin:
            close.append((idx, float(d)))
    print(f"{name}: {len(entries)} entries, closest to a float32 rounding boundary: {float(worst):.1f} ulp64 (margin {margin}); inside the margin: {close}")
    return close


def test_trig_tables_are_libm_independent():
    pi = mp.pi
    c36, c12, c72, c24, c64 = (go_const(pi / 36), go_const(pi / 12), go_const(pi / 72), go_const(pi / 24), go_const(pi / 64))
    cos36 = hostemu_lib.table(0)
    cos12 = hostemu_lib.table(1)
    win = hostemu_lib.table(2)
    synth = hostemu_lib.table(3)
    close = []
    # imdct.go:61-79: float32(math.Cos(math.Pi / (2 N) * (2 j + 1 + N / 2) * (2 i + 1))), float64 arithmetic left to right
    close += check("cosN36", [(i * 36 + j, mp.cos(mp.mpf(c72 * (2.0 * j + 1.0 + 18.0) * (2.0 * i + 1.0)))) for i in range(18) for j in range(36)], cos36, 4)
    close += check("cosN12", [(i * 12 + j, mp.cos(mp.mpf(c24 * (2.0 * j + 1.0 + 6.0) * (2.0 * i + 1.0)))) for i in range(6) for j in range(12)], cos12, 4)
    # frame.go:490-497: float32(math.Cos(float64((16 + i) * (2 j + 1)) * (math.Pi / 64.0)))
    close += check("synthNWin", [(i * 32 + j, mp.cos(mp.mpf(float((16 + i) * (2 * j + 1)) * c64))) for i in range(64) for j in range(32)], synth, 4)
    # imdct.go:23-57
    w = []
    w += [(0 * 36 + i, mp.sin(mp.mpf(c36 * (i + 0.5)))) for i in range(36)]
    w += [(1 * 36 + i, mp.sin(mp.mpf(c36 * (i + 0.5)))) for i in range(18)]
    w += [(1 * 36 + i, mp.mpf(1)) for i in range(18, 24)]
    w += [(1 * 36 + i, mp.sin(mp.mpf(c12 * (i + 0.5 - 18.0)))) for i in range(24, 30)]
    w += [(2 * 36 + i, mp.sin(mp.mpf(c12 * (i + 0.5)))) for i in range(12)]
    w += [(3 * 36 + i, mp.sin(mp.mpf(c12 * (i + 0.5 - 6.0)))) for i in range(6, 12)]
    w += [(3 * 36 + i, mp.mpf(1)) for i in range(12, 18)]
    w += [(3 * 36 + i, mp.sin(mp.mpf(c36 * (i + 0.5)))) for i in range(18, 36)]
    close += check("imdctWin", [(k, x) for k, x in w if x != 1], win, 4)
    # synthNWin row 16 is cos(odd * pi / 2): tiny values (~1e-17 .. 1e-15) whose float32 rounding is far from any boundary too
    assert close == [], close


def test_requantisation_rows_are_libm_independent():
    """powq4[q][i] = float32(2^(q/4) * i^(4/3)), the only data the requantiser multiplies by an exact power of two."""
    L = hostemu_lib.lib()
    L.emu_powq4.restype = C.POINTER(C.c_float)
    L.emu_powq4.argtypes = [C.POINTER(C.c_int)]
    n = C.c_int()
    p = L.emu_powq4(C.byref(n))
    tab = np.ctypeslib.as_array(p, shape=(n.value,)).copy()
    row = n.value // 4
    close = []
    y = mp.mpf(4.0 / 3.0)  # the reference's exponent is the float64 constant 4.0/3.0 (frame.go:38), not 4/3
    for q in range(4):
        scale = mp.mpf(2) ** (mp.mpf(q) / 4)
        entries = [(q * row + i, scale * mp.mpf(i) ** y) for i in range(1, 8207)]
        close += check(f"powq4[q={q}]", entries, tab, 32)
    assert close == [], close
    # and the oracle's float64 powtab34 (glibc pow) is within 1 ulp64 of the exact i^(4.0/3.0) everywhere
    ref = np.ctypeslib.as_array(oracle.lib().orc_table_powtab34(), shape=(8207,))
    worst = 0.0
    for i in range(1, 8207):
        x = mp.mpf(i) ** y
        e = int(mp.floor(mp.log(x, 2)))
        worst = max(worst, float(abs(mp.mpf(float(ref[i])) - x) / mp.mpf(2) ** (e - 52)))
    print(f"powtab34 (float64, glibc pow): max error {worst:.3f} ulp64")
    assert worst <= 1.0
