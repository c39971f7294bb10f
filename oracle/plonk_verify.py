"""CPU restatement of ark-plonk's verifier (TEST INFRASTRUCTURE): the acceptance oracle for proofs.

Oracle = test infrastructure (see oracle/__init__.py).  Follows proof_system/proof.rs:111-426
(`Proof::verify`): transcript replay (:128-223,250-295), `compute_r0` (:428-486),
`compute_linearisation_commitment` (:489-603) with the verifier-key terms of
widget/arithmetic.rs:128-158, widget/lookup.rs:236-291, proof_system/permutation.rs:328-386, the
compressed table commitment (:316-325) and the two batched KZG checks (:345-425).

The pairing equation of `PC::check`, e(C - v G, H) = e(W, tau H - z H), is evaluated in its
discrete-log form in G1, C - v G == (tau - z) W, because the harness knows tau.
Range / logic / fixed-base / curve-addition selector commitments enter through
`extend_linearisation_commitment` (widget/mod.rs:152-170, proof.rs:530-559); public inputs through
the barycentric pi_eval of compute_r0.
"""
from __future__ import annotations

from . import gates
from .curves import Curve
from .ntt import Domain
from .plonk import K1, K2, K3, PlonkTranscript, lc
from .serialize import deser_g1, ser_field, ser_u64


def parse_proof(curve: Curve, blob: bytes):
    """Proof wire format (proof.rs:51-103)"""
    pos = 0

    def take(n):
        nonlocal pos
        b = blob[pos:pos + n]
        assert len(b) == n, "truncated proof"
        pos += n
        return b

    comms = [deser_g1(curve, take(48)) for _ in range(13)]
    openings = []
    for _ in range(2):
        w = deser_g1(curve, take(48))
        assert take(1) == b"\x00", "random_v must be None"
        openings.append(w)
    evals = [int.from_bytes(take(32), "little") for _ in range(16)]
    count = int.from_bytes(take(8), "little")
    custom = []
    for _ in range(count):
        ln = int.from_bytes(take(8), "little")
        label = take(ln).decode()
        custom.append((label, int.from_bytes(take(32), "little")))
    assert pos == len(blob), "trailing bytes"
    names = ("a", "b", "c", "d", "z", "f", "h_1", "h_2", "z_2", "t_1", "t_2", "t_3", "t_4")
    return dict(zip(names, comms)), openings, evals, custom


def verify(curve: Curve, vk: dict, n: int, blob: bytes, tau: int, label: bytes = b"ark", public_inputs=None) -> bool:
    """vk: name -> affine point for q_m q_l q_r q_o q_4 q_c q_arith q_range q_logic q_fixed_group_add
    q_variable_group_add q_lookup left_sigma right_sigma out_sigma fourth_sigma table_1..table_4
    (missing selectors = identity).  public_inputs: row -> value."""
    p = curve.fr.p
    C, (aw_open, saw_open), ev, custom = parse_proof(curve, blob)
    (a_e, b_e, c_e, d_e, s1, s2, s3, zhat, q_lookup_e, z2_next, h1_e, h1_next, h2_e, f_e, table_e, table_next) = ev
    cust = dict(custom)
    tr = PlonkTranscript(curve, label)
    public_inputs = public_inputs or {}
    pi_ser = ser_u64(len(public_inputs))
    for pos in sorted(public_inputs):
        pi_ser += ser_u64(pos) + ser_field(curve.fr, public_inputs[pos])
    tr.append_bytes(b"pi", pi_ser)
    for lab, key in ((b"w_l", "a"), (b"w_r", "b"), (b"w_o", "c"), (b"w_4", "d")):
        tr.append_g1(lab, C[key])
    zeta = tr.challenge(b"zeta"); tr.append_fr(b"zeta", zeta)
    tr.append_g1(b"f", C["f"]); tr.append_g1(b"h1", C["h_1"]); tr.append_g1(b"h2", C["h_2"])
    beta = tr.challenge(b"beta"); tr.append_fr(b"beta", beta)
    gamma = tr.challenge(b"gamma"); tr.append_fr(b"gamma", gamma)
    delta = tr.challenge(b"delta"); tr.append_fr(b"delta", delta)
    epsilon = tr.challenge(b"epsilon"); tr.append_fr(b"epsilon", epsilon)
    tr.append_g1(b"z", C["z"])
    alpha = tr.challenge(b"alpha"); tr.append_fr(b"alpha", alpha)
    seps = []
    for ch, ap in ((b"range separation challenge", b"range seperation challenge"),
                   (b"logic separation challenge", b"logic seperation challenge"),
                   (b"fixed base separation challenge", b"fixed base separation challenge"),
                   (b"variable base separation challenge", b"variable base separation challenge")):
        v = tr.challenge(ch); tr.append_fr(ap, v)
        seps.append(v)
    ls = tr.challenge(b"lookup separation challenge"); tr.append_fr(b"lookup separation challenge", ls)
    for lab, key in ((b"t_1", "t_1"), (b"t_2", "t_2"), (b"t_3", "t_3"), (b"t_4", "t_4")):
        tr.append_g1(lab, C[key])
    zc = tr.challenge(b"z"); tr.append_fr(b"z", zc)

    zh = (pow(zc, n, p) - 1) % p
    l1 = zh * pow(n * (zc - 1) % p, -1, p) % p
    alpha_sq = alpha * alpha % p
    lsq, lcu = ls * ls % p, pow(ls, 3, p)
    opd = (1 + delta) % p
    eopd = epsilon * opd % p
    # compute_r0 (proof.rs:426-486) with pi_eval by barycentric evaluation (proof.rs:635-677)
    omega_inv = pow(Domain.for_size(curve.fr, n).group_gen, -1, p)
    pi_eval = zh * pow(n, -1, p) % p * sum(v * pow((pow(omega_inv, pos, p) * zc - 1) % p, -1, p)
                                           for pos, v in public_inputs.items()) % p
    r0 = (pi_eval - (a_e + beta * s1 + gamma) * (b_e + beta * s2 + gamma) % p * (c_e + beta * s3 + gamma) % p
          * ((d_e + gamma) * zhat % p * alpha % p)
          - l1 * alpha_sq
          - lsq * z2_next % p * (eopd + delta * h2_e) % p * (eopd + h2_e + delta * h1_next)
          - lcu * l1) % p
    for lab, v in ((b"a_eval", a_e), (b"b_eval", b_e), (b"c_eval", c_e), (b"d_eval", d_e), (b"left_sig_eval", s1),
                   (b"right_sig_eval", s2), (b"out_sig_eval", s3), (b"perm_eval", zhat), (b"f_eval", f_e),
                   (b"q_lookup_eval", q_lookup_e), (b"lookup_perm_eval", z2_next), (b"h_1_eval", h1_e),
                   (b"h_1_next_eval", h1_next), (b"h_2_eval", h2_e)):
        tr.append_fr(lab, v)
    for lab, v in custom:
        tr.append_fr(lab.encode(), v)

    # linearisation commitment: 6 arithmetic + 3 lookup + 2 permutation + 4 quotient terms
    q_arith = cust["q_arith_eval"]
    terms = [(a_e * b_e % p * q_arith, vk.get("q_m")), (a_e * q_arith, vk.get("q_l")), (b_e * q_arith, vk.get("q_r")),
             (c_e * q_arith, vk.get("q_o")), (d_e * q_arith, vk.get("q_4")), (q_arith, vk.get("q_c"))]
    # custom gates: selector commitment * constraints(evaluations) (proof.rs:530-559)
    EA, ED = gates.embedded_params(curve)
    cg = gates.custom_gate_sum([1, 1, 1, 1], seps, (a_e, b_e, c_e, d_e),
                               (cust["a_next_eval"], cust["b_next_eval"], cust["d_next_eval"]),
                               cust["q_l_eval"], cust["q_r_eval"], cust["q_c_eval"], EA, ED, p)
    terms += list(zip(cg, (vk.get("q_range"), vk.get("q_logic"), vk.get("q_fixed_group_add"),
                           vk.get("q_variable_group_add"))))
    terms.append(((lc([a_e, b_e, c_e, d_e], zeta, p) - f_e) * ls, vk.get("q_lookup")))
    terms.append((opd * (epsilon + f_e) % p * (eopd + table_e + delta * table_next) % p * lsq + l1 * lcu, C["z_2"]))
    terms.append(((-z2_next) * lsq % p * (eopd + h2_e + delta * h1_next), C["h_1"]))
    bz = beta * zc % p
    x = (a_e + bz + gamma) * (b_e + K1 * bz + gamma) % p * (c_e + K2 * bz + gamma) % p * ((d_e + K3 * bz + gamma) * alpha % p)
    terms.append((x + l1 * alpha_sq, C["z"]))
    y = -((a_e + beta * s1 + gamma) * (b_e + beta * s2 + gamma) % p * (c_e + beta * s3 + gamma) % p * (beta * zhat % p * alpha % p))
    terms.append((y, vk.get("fourth_sigma")))
    z_n = (zh + 1) % p
    t_scal = [(-zh) % p]
    for _ in range(3):
        t_scal.append(t_scal[-1] * z_n % p)
    terms += list(zip(t_scal, (C["t_1"], C["t_2"], C["t_3"], C["t_4"])))

    def msm(pairs):
        acc = (0, 1, 0)
        for s, P in pairs:
            if P is None or s % p == 0:
                continue
            acc = curve.jadd(acc, curve.to_jac(curve.mul(P, s % p)))
        return curve.to_affine(acc)

    lin_comm = msm(terms)
    table_comm = msm([(1, vk.get("table_1")), (zeta, vk.get("table_2")), (zeta * zeta % p, vk.get("table_3")),
                      (pow(zeta, 3, p), vk.get("table_4"))])
    aw_ch = tr.challenge(b"aggregate_witness")
    saw_ch = tr.challenge(b"aggregate_witness")
    omega = Domain.for_size(curve.fr, n).group_gen

    def kzg_check(comms, point, values, W, ch):
        # sonic_pc::check: combine with powers of the opening challenge starting at 1
        pairs, v, cur = [], 0, 1
        for Cc, val in zip(comms, values):
            pairs.append((cur, Cc))
            v = (v + cur * val) % p
            cur = cur * ch % p
        lhs = curve.add(msm(pairs), curve.neg(curve.mul(curve.G, v)))        # C - v G
        rhs = curve.mul(W, (tau - point) % p) if W is not None else None    # (tau - z) W
        return lhs == rhs

    ok1 = kzg_check([lin_comm, vk.get("left_sigma"), vk.get("right_sigma"), vk.get("out_sigma"), C["f"], C["h_2"], table_comm,
                     C["a"], C["b"], C["c"], C["d"]], zc,
                    [(-r0) % p, s1, s2, s3, f_e, h2_e, table_e, a_e, b_e, c_e, d_e], aw_open, aw_ch)
    ok2 = kzg_check([C["z"], C["a"], C["b"], C["d"], C["h_1"], C["z_2"], table_comm], zc * omega % p,
                    [zhat, cust["a_next_eval"], cust["b_next_eval"], cust["d_next_eval"], h1_next, z2_next, table_next],
                    saw_open, saw_ch)
    return ok1 and ok2
