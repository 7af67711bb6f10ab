"""Dormand-Prince 5(4) tableau with the 4th-order dense-output polynomial, as rational numbers (J. R. Dormand & P. J.
Prince 1980; L. F. Shampine 1986 for the interpolant) — the published constants SciPy's ``RK45`` uses, which is what the
reference's ``solver='scipy'`` runs (pyrates/backend/base/base_backend.py:802-812).  Emitted into the adaptive instance
kernel as ``constexpr`` tables (csrc/templates/inst_kernel.cuh)."""
from __future__ import annotations

C = [0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1]
A = [[0, 0, 0, 0, 0],
     [1 / 5, 0, 0, 0, 0],
     [3 / 40, 9 / 40, 0, 0, 0],
     [44 / 45, -56 / 15, 32 / 9, 0, 0],
     [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729, 0],
     [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656]]
B = [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84]
E = [-71 / 57600, 0, 71 / 16695, -71 / 1920, 17253 / 339200, -22 / 525, 1 / 40]
P = [[1, -8048581381 / 2820520608, 8663915743 / 2820520608, -12715105075 / 11282082432],
     [0, 0, 0, 0],
     [0, 131558114200 / 32700410799, -68118460800 / 10900136933, 87487479700 / 32700410799],
     [0, -1754552775 / 470086768, 14199869525 / 1410260304, -10690763975 / 1880347072],
     [0, 127303824393 / 49829197408, -318862633887 / 49829197408, 701980252875 / 199316789632],
     [0, -282668133 / 205662961, 2019193451 / 616988883, -1453857185 / 822651844],
     [0, 40617522 / 29380423, -110615467 / 29380423, 69997945 / 29380423]]


# Bogacki-Shampine 3(2) pair with its cubic Hermite interpolant — SciPy's ``RK23`` (the reference's own solver test uses
# ``method='RK23'``, tests/test_backend_simulations.py:370-373)
RK23 = dict(
    C=[0, 1 / 2, 3 / 4],
    A=[[0, 0], [1 / 2, 0], [0, 3 / 4]],
    B=[2 / 9, 1 / 3, 4 / 9],
    E=[5 / 72, -1 / 12, -1 / 9, 1 / 8],
    P=[[1, -4 / 3, 5 / 9], [0, 1, -2 / 3], [0, 4 / 3, -8 / 9], [0, -1, 1]],
    error_exponent=-1 / 3)
RK45 = dict(C=C, A=A, B=B, E=E, P=P, error_exponent=-1 / 5)
METHODS = {"RK45": RK45, "RK23": RK23}


def _lit(v) -> str:
    r = repr(float(v))
    return r if ("." in r or "e" in r) else r + ".0"


def _rows(m) -> str:
    return "{" + ", ".join("{" + ", ".join(_lit(v) for v in row) + "}" for row in m) + "}"


def cuda_method_tables(method: str = "RK45") -> str:
    """Tableau of one explicit Runge-Kutta pair for the instance kernel: stage count, interpolant degree, controller
    exponent and the coefficient initialisers (exact double literals, repr round-trips)."""
    m = METHODS[method]
    nst, npoly = len(m["B"]), len(m["P"][0])
    a_rows = [list(row) + [0] * (nst - 1 - len(row)) for row in m["A"]]
    return "\n".join([
        f"#define PRB_RK_NST {nst}", f"#define PRB_RK_NP {npoly}", f"#define PRB_RK_EXPONENT {_lit(m['error_exponent'])}",
        "#define PRB_RK_C_INIT {" + ", ".join(_lit(v) for v in m["C"]) + "}",
        "#define PRB_RK_A_INIT " + _rows(a_rows),
        "#define PRB_RK_B_INIT {" + ", ".join(_lit(v) for v in m["B"]) + "}",
        "#define PRB_RK_E_INIT {" + ", ".join(_lit(v) for v in m["E"]) + "}",
        "#define PRB_RK_P_INIT " + _rows(m["P"])])


def cuda_tables() -> str:
    """``#define`` initialisers consumed by the kernel template (exact double literals, repr round-trips)."""
    return "\n".join([
        "#define PRB_RK45_C_INIT {" + ", ".join(_lit(v) for v in C) + "}",
        "#define PRB_RK45_A_INIT " + _rows(A),
        "#define PRB_RK45_B_INIT {" + ", ".join(_lit(v) for v in B) + "}",
        "#define PRB_RK45_E_INIT {" + ", ".join(_lit(v) for v in E) + "}",
        "#define PRB_RK45_P_INIT " + _rows(P)])
