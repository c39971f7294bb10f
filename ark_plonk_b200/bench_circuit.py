"""The reference's benchmark circuit as plain arrays (host side, front end).

Builds what `BenchCircuit::gadget` leaves in the `StandardComposer` (benches/plonk.rs:45-68;
constraint_system/composer.rs:202-238 zero row + 3 blinding rows, :493-548 dummy rows, :553-574
dummy lookup table): selector columns, wire columns, the copy-constraint permutation and the
lookup table, for a padded size of 2^degree.  The 8 blinding witnesses the reference draws from
OsRng (composer.rs:580-648) are explicit inputs.  Circuit construction is CPU front-end work in
the reference too (out of scope for the GPU).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .encoding import FR_MODULUS

SELECTORS = ("q_m", "q_l", "q_r", "q_o", "q_c", "q_4", "q_arith", "q_range", "q_logic",
             "q_fixed_group_add", "q_variable_group_add", "q_lookup")


@dataclass
class CircuitArrays:
    curve: int
    n: int                      # padded size (power of two)
    rows: int                   # real rows
    selectors: dict             # name -> (n,) int64 index into `values`
    wires: np.ndarray           # (4, n) int64 index into `values`
    values: list                # distinct field values (python ints)
    sigma: np.ndarray           # (4, n, 2) int64: sigma[col, row] = (col', row')
    table: list                 # lookup table rows (python ints)
    public_inputs: dict = field(default_factory=dict)   # row -> value (non-zero entries only)


def build(curve: int, degree: int, blinders) -> CircuitArrays:
    p = FR_MODULUS[curve]
    assert len(blinders) == 8
    size = 1 << degree
    # distinct values: 0, 1, 2, 3, 4, 6, 7, 127, -20, then the 8 blinders
    values = [0, 1, 2, 3, 4, 6, 7, 127, (-20) % p] + [b % p for b in blinders]
    vid = {v: i for i, v in enumerate(values[:9])}
    Z, ONE, SIX, SEVEN, M20 = vid[0], vid[1], vid[6], vid[7], vid[(-20) % p]
    # rows: 0 = zero constraint, 1..3 = blinding, then pairs of dummy rows until the bound is met
    n_pairs = 0
    rows = 4
    while (1 << max(max(rows, 3) - 1, 0).bit_length()) < size - 1:
        rows += 2
        n_pairs += 1
    n = 1 << max(max(rows, 3) - 1, 0).bit_length()
    sel = {s: np.zeros(n, dtype=np.int64) for s in SELECTORS}
    wires = np.zeros((4, n), dtype=np.int64)
    # variable ids: 0 = zero; 1..8 = blinders (r1,r2,r3,r4 of row 1, then of row 2); then 4 per pair
    var_val = [Z] + [9 + k for k in range(8)]
    wvar = np.zeros((4, rows), dtype=np.int64)             # variable id per (col, row)
    sel["q_l"][0] = ONE
    sel["q_arith"][0] = ONE
    wvar[:, 1] = [1, 2, 3, 4]
    wvar[:, 2] = [5, 6, 7, 8]
    wvar[:, 3] = [5, 6, 0, 0]
    base = 9
    pair = np.arange(n_pairs, dtype=np.int64)
    six, one, seven, m20 = base + 4 * pair, base + 4 * pair + 1, base + 4 * pair + 2, base + 4 * pair + 3
    r0, r1 = 4 + 2 * pair, 5 + 2 * pair
    wvar[0, r0], wvar[1, r0], wvar[2, r0], wvar[3, r0] = six, seven, m20, one
    wvar[0, r1], wvar[1, r1], wvar[2, r1], wvar[3, r1] = m20, six, seven, 0
    for name, v0, v1 in (("q_m", 1, 1), ("q_l", 2, 1), ("q_r", 3, 1), ("q_o", 4, 1), ("q_c", 4, 127), ("q_4", 1, 0),
                         ("q_arith", 1, 1), ("q_lookup", 1, 1)):
        sel[name][r0] = vid[v0]
        sel[name][r1] = vid[v1]
    nvars = base + 4 * n_pairs
    val_of_var = np.zeros(nvars, dtype=np.int64)
    val_of_var[:9] = var_val
    val_of_var[six], val_of_var[one], val_of_var[seven], val_of_var[m20] = SIX, ONE, SEVEN, M20
    wires[:, :rows] = val_of_var[wvar]                       # padding rows hold the zero value
    # copy constraints: each variable's wire list in insertion order (row-major, cols a,b,c,d) is a cycle
    pos_var = wvar.T.reshape(-1)                             # index = row*4 + col
    order = np.argsort(pos_var, kind="stable")
    sorted_var = pos_var[order]
    nxt = np.empty_like(order)
    start = np.flatnonzero(np.r_[True, sorted_var[1:] != sorted_var[:-1]])
    end = np.r_[start[1:], len(order)]
    nxt[:-1] = order[1:]
    nxt[end - 1] = order[start]
    sigma = np.empty((4, n, 2), dtype=np.int64)
    sigma[:, :, 0] = np.arange(4)[:, None]
    sigma[:, :, 1] = np.arange(n)[None, :]
    src_row, src_col = order // 4, order % 4
    sigma[src_col, src_row, 0] = nxt % 4
    sigma[src_col, src_row, 1] = nxt // 4
    table = [[6, 7, (-20) % p, 1], [(-20) % p, 6, 7, 0], [3, 1, 4, 9]]
    return CircuitArrays(curve=curve, n=n, rows=rows, selectors=sel, wires=wires, values=values, sigma=sigma, table=table)
