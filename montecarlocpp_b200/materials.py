"""Synthetic material files in the reference's on-disk format (material.cpp:86-114).

The reference ships no material data (its .gitignore excludes everything outside
montecarlo/), only the parser, so the build fixes two self-consistent synthetic
materials.  The SAME files feed the CPU oracle and the CUDA path, so
their physical accuracy is irrelevant to parity.

disp file : line 1 `nw np`; then nw rows of `omega domega [vel_p dos_p]*np`.
relax file: np rows of 8 numbers = 2 mechanisms x (A, a, b, c), tau^-1 = sum A w^a T^b e^(-c/T).
"""
import math
import os

# constants.h:17-19 (the reference's literal values)
PI = 3.141592653589793
HBAR = 1.054560652927e-34
KB = 1.380648e-23


def _dedT(omega, temp):
    x = HBAR / (KB * temp) * omega
    return KB * (x / (2.0 * math.sinh(x / 2.0))) ** 2


def _write(path, rows):
    with open(path, "w") as f:
        for row in rows:
            f.write(" ".join(repr(v) if isinstance(v, float) else str(v) for v in row) + "\n")


def write_grey(dirname, temp=300.0, omega=5e13, vel=6000.0, energy_sum=1.66e6, inv_tau=1.5e11):
    """One frequency, one branch: C = 1.66e6 J/m^3K, v = 6 km/s, tau = 6.67 ps (MFP 40 nm)."""
    os.makedirs(dirname, exist_ok=True)
    dos = energy_sum / _dedT(omega, temp)        # domega = 1
    disp = os.path.join(dirname, "grey_disp.txt")
    relax = os.path.join(dirname, "grey_relax2.txt")
    _write(disp, [[1, 1], [float(omega), 1.0, float(vel), dos]])
    _write(relax, [[float(inv_tau), 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0]])
    return disp, relax


SI_BRANCHES = [  # (v_s, c, A_U, b_U, c_U)   omega = v_s k + c k^2
    (9.01e3, -2.0e-7, 2.0e-19, 1.49, 180.0),     # LA
    (5.23e3, -2.26e-7, 1.2e-19, 1.65, 180.0),    # TA
    (5.23e3, -2.26e-7, 1.2e-19, 1.65, 180.0),    # TA
]
SI_KMAX = 2.0 * PI / 5.43e-10
SI_IMPURITY = 3.0e-45


def write_silicon(dirname, nw=1000):
    """Si-like full dispersion: nw uniform bins on a common grid, LA + 2 TA quadratic branches.

    Bins above a branch's zone-edge frequency carry zero density of states for that branch.
    """
    os.makedirs(dirname, exist_ok=True)
    wmax = [vs * SI_KMAX + c * SI_KMAX ** 2 for vs, c, *_ in SI_BRANCHES]
    dw = max(wmax) / nw
    rows = [[nw, len(SI_BRANCHES)]]
    for i in range(nw):
        w = (i + 0.5) * dw
        row = [w, dw]
        for b, (vs, c, *_rest) in enumerate(SI_BRANCHES):
            if w < wmax[b]:
                k = (-vs + math.sqrt(vs * vs + 4.0 * c * w)) / (2.0 * c)
                v = vs + 2.0 * c * k
                dos = k * k / (2.0 * PI * PI * v)
            else:
                v, dos = 1.0, 0.0
            row += [v, dos]
        rows.append(row)
    disp = os.path.join(dirname, "Si_disp.txt")
    relax = os.path.join(dirname, "Si_relax2.txt")
    _write(disp, rows)
    _write(relax, [[a, 2.0, b, c, SI_IMPURITY, 4.0, 0.0, 0.0] for (_vs, _c, a, b, c) in SI_BRANCHES])
    return disp, relax


def write_all(dirname, nw=1000):
    g = write_grey(dirname)
    s = write_silicon(dirname, nw)
    return {"grey": g, "silicon": s}
