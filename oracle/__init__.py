"""CPU oracle for the ark-plonk proving hot path (TEST INFRASTRUCTURE ONLY).

This package restates, with Python big integers, the arithmetic that
heliaxdev/ark-plonk reaches through the arkworks 0.3 crates (ark-ff, ark-ec,
ark-poly, ark-poly-commit, ark-serialize) and merlin 3.0.  None of those crates
is vendored in /root/reference and no Rust toolchain exists in this image, so
the oracle follows their published algorithms and is anchored on the reference's
own call sites (cited per function as `plonk-core/src/...:line`).

PARITY UNPINNED BY REFERENCE BYTES: the reference holds no byte-level golden
vectors for this path; the oracle is pinned instead by
 (1) the curve/field constants re-derived numerically (tests/test_oracle_*.py),
 (2) the Merlin conformance vector,
 (3) mathematical uniqueness: commit(p) == [p(tau)]G, fft == naive DFT,
 (4) the reference's own algebraic tests restated (permutation identities,
     `test_combine_split` expected vectors, prove -> verify acceptance).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference arm
may import this package.  The product (ark_plonk_b200) never does.
"""
