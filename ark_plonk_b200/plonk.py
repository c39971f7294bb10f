"""Device-resident PLONK prover over the apb C ABI (host orchestration).

Mirrors the reference's call schedule: `preprocess` = StandardComposer::preprocess_prover
(plonk-core/src/proof_system/preprocess.rs:126-243,267-423), `Prover.prove` =
Prover::prove_with_preprocessed (proof_system/prover.rs:163-638) with quotient_poly.rs:34-178 and
linearisation_poly.rs:164-349.  Every FFT / commitment the reference issues through
`EvaluationDomain` / `PC::commit` / `PC::open` is issued here through `apb_ntt*` / `apb_msm*`;
the pointwise loops between them run as CUDA kernels (`apb_plonk_*`, `apb_fr_lincomb`,
`apb_poly_*`), so polynomials never leave HBM.  The host handles the Fiat-Shamir transcript
(`apb_transcript_*`), ~50 scalar field operations per proof (Python ints) and proof assembly.

Scope: all of the reference's gate terms -- arithmetic, permutation, lookup, and the range /
logic / fixed-base / curve-addition custom gates (widget/*.rs) -- and public inputs.  A custom
gate selector that is identically zero (all four are in `BenchCircuit`, composer.rs:506-509,
531-534) costs nothing: no resident vectors, no term in the kernel.
There is no CPU fallback: all vector work goes through the CUDA library.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import encoding as enc
from . import gates
from ._lib import NTT_COSET_FFT, NTT_COSET_IFFT, NTT_FFT, NTT_IFFT, ApbError, Lib, get_lib
from .bench_circuit import SELECTORS, CircuitArrays
from .domain import Radix2EvaluationDomain
from .kzg import CommitterKey

K1, K2, K3 = 7, 13, 17
CUSTOM_SELECTORS = ("q_range", "q_logic", "q_fixed_group_add", "q_variable_group_add")
FR_GENERATOR = (7, 22)


class Transcript:
    """transcript.rs:16-50 over the library's merlin implementation"""

    def __init__(self, lib: Lib, curve: int, label: bytes):
        self.lib, self.curve = lib, curve
        self.p = enc.FR_MODULUS[curve]
        h = C.c_void_p()
        lib.check(lib.c.apb_transcript_new(label, len(label), C.byref(h)))
        self._h = h

    def append_bytes(self, label: bytes, msg: bytes):
        self.lib.check(self.lib.c.apb_transcript_append(self._h, label, len(label), msg, len(msg)))

    def append_fr(self, label: bytes, v: int):
        self.append_bytes(label, int(v).to_bytes(32, "little"))

    def challenge(self, label: bytes) -> int:
        nbytes = self.p.bit_length() // 8                 # size_in_bits() / 8 = 31
        buf = (C.c_uint8 * nbytes)()
        self.lib.check(self.lib.c.apb_transcript_challenge(self._h, label, len(label), buf, nbytes))
        return int.from_bytes(bytes(buf), "little") % self.p

    def __del__(self):
        try:
            self.lib.c.apb_transcript_free(self._h)
        except Exception:
            pass


class Arena:
    """One HBM allocation, bump-allocated in Fr elements (MSM batches address scalars by offset)."""

    def __init__(self, lib: Lib, elems: int, torch_device: str | None = None):
        """`torch_device` ("cuda" / "cpu"): back the arena by a torch tensor so that slices can be used
        in torch.distributed collectives (multi-GPU commit split); otherwise the library allocates."""
        self.lib = lib
        self.elems = elems
        self.tensor = None
        if torch_device is not None:
            import torch
            self.tensor = torch.zeros(elems * 4, dtype=torch.int64, device=torch_device)
            self.base = self.tensor.data_ptr()
        else:
            p = C.c_void_p()
            lib.check(lib.c.apb_dev_alloc(elems * 32, C.byref(p)))
            self.base = p.value
        self.used = 0

    def alloc(self, elems: int) -> int:
        """returns the element offset"""
        off = self.used
        if off + elems > self.elems:
            raise ApbError(6, "arena exhausted")
        self.used += elems
        return off

    def ptr(self, off: int) -> int:
        return self.base + off * 32

    def mark(self) -> int:
        return self.used

    def release(self, mark: int):
        self.used = mark

    def upload(self, off: int, arr: np.ndarray):
        arr = np.ascontiguousarray(arr, dtype=np.uint64)
        self.lib.check(self.lib.c.apb_dev_upload(self.ptr(off), arr.ctypes.data, arr.nbytes))

    def download(self, off: int, elems: int) -> np.ndarray:
        out = np.empty((elems, 4), dtype=np.uint64)
        self.lib.check(self.lib.c.apb_dev_download(out.ctypes.data, self.ptr(off), elems * 32))
        return out

    def view(self, off: int, elems: int):
        """torch view (int64 words) of `elems` elements at element offset `off` (tensor-backed arenas)"""
        return self.tensor[off * 4:(off + elems) * 4]

    def close(self):
        if self.tensor is not None:
            self.lib.c.apb_dev_sync()
            self.tensor = None
            self.base = 0
        elif self.base:
            self.lib.c.apb_dev_free(self.base)
            self.base = 0


def _mont(curve: int, v: int) -> np.ndarray:
    return enc.fr_to_mont(curve, [v])[0]


def _mont_list(curve: int, vals) -> np.ndarray:
    return np.ascontiguousarray(enc.fr_to_mont(curve, list(vals)))


@dataclass
class ProverKey:
    curve: int
    n: int
    arena: Arena
    dom: Radix2EvaluationDomain
    dom4: Radix2EvaluationDomain
    poly: dict = field(default_factory=dict)        # name -> element offset of n coefficients
    ev4: dict = field(default_factory=dict)         # name -> element offset of 4n coset evaluations
    tables: list = None                             # 4 offsets of padded table columns (evaluation form)
    vh_inv: np.ndarray = None                       # (4, 4) Montgomery
    commitments: dict = field(default_factory=dict) # verifier-key commitments, compressed bytes
    q_lookup_evals: int = 0                         # offset: q_lookup on the n-domain (prover.rs:252-254)
    custom: tuple = ()                              # the custom gate selectors that are not identically zero
    public_inputs: dict = field(default_factory=dict)


class Prover:
    def __init__(self, curve: int, ck: CommitterKey, lib: Lib | None = None, committer=None,
                 arena_device: str | None = None):
        """`committer`: optional parallel.DistributedCommitter that spreads the polynomials of each
        PC::commit over the ranks of a process group (needs a tensor-backed arena: `arena_device`)."""
        self.lib = lib or get_lib()
        self.curve = curve
        self.ck = ck
        self.committer = committer
        self.arena_device = arena_device
        self.p = enc.FR_MODULUS[curve]
        self.msm_calls = 0
        self.ntt_calls = 0
        self.phase_log = None          # set to a list to record (k, per-phase ms) of every commit call

    # ---- thin wrappers -----------------------------------------------------------------------
    def _ntt(self, dom, kind, arena, src, in_len, dst, batch=1, in_stride=None, out_stride=None):
        self.ntt_calls += batch
        size = dom.size
        if batch == 1:
            dom.ntt_dev(kind, arena.ptr(src), in_len, arena.ptr(dst))
        else:
            dom.ntt_batch_dev(kind, arena.ptr(src), in_len, in_stride or size, arena.ptr(dst), out_stride or size, batch)

    def _commit(self, arena: Arena, offs, lens):
        """PC::commit of device-resident polynomials -> list of (xyz, compressed bytes)"""
        k = len(offs)
        self.msm_calls += k
        if self.committer is not None and self.committer.world > 1:
            out = self.committer.commit(arena, offs, lens)
            if self.phase_log is not None:       # this rank's share of the batch
                self.phase_log.append((k, self.lib.last_device_ms(), self.lib.msm_phase_ms()))
            return [(out[i], self.lib.g1_compress(self.curve, out[i])) for i in range(k)]
        so = (C.c_size_t * k)(*offs)
        bo = (C.c_size_t * k)(*([0] * k))
        ln = (C.c_size_t * k)(*lens)
        out = np.zeros((k, 18), dtype=np.uint64)
        self.lib.check(self.lib.c.apb_msm_batch_dev(self.ck._h, k, arena.base, so, bo, ln, 1, out.ctypes.data))
        if self.phase_log is not None:
            self.phase_log.append((k, self.lib.last_device_ms(), self.lib.msm_phase_ms()))
        return [(out[i], self.lib.g1_compress(self.curve, out[i])) for i in range(k)]

    def _lincomb(self, arena, offs, lens, scalars, dst, out_len):
        k = len(offs)
        ptrs = (C.c_void_p * k)(*[arena.ptr(o) for o in offs])
        ln = (C.c_size_t * k)(*lens)
        sc = _mont_list(self.curve, scalars)
        self.lib.check(self.lib.c.apb_fr_lincomb(self.curve, k, ptrs, ln, sc.ctypes.data, arena.ptr(dst), out_len))

    def _eval(self, arena, offs, lens, points):
        k = len(offs)
        ptrs = (C.c_void_p * k)(*[arena.ptr(o) for o in offs])
        ln = (C.c_size_t * k)(*lens)
        pts = _mont_list(self.curve, points)
        out = np.zeros((k, 4), dtype=np.uint64)
        self.lib.check(self.lib.c.apb_poly_eval(self.curve, k, ptrs, ln, pts.ctypes.data, out.ctypes.data))
        return enc.fr_from_mont(self.curve, out)

    # ---- preprocessing ---------------------------------------------------------------------
    def preprocess(self, circ: CircuitArrays, commit_verifier_key: bool = True) -> ProverKey:
        curve, n, p = self.curve, circ.n, self.p
        lib = self.lib
        # a custom gate selector that is identically zero contributes nothing to the quotient or the
        # linearisation polynomial and commits to the identity: it gets no resident vectors
        vals_obj = np.array(circ.values, dtype=object)
        custom = tuple(s for s in CUSTOM_SELECTORS if np.any(vals_obj[circ.selectors[s]] != 0))
        dom = Radix2EvaluationDomain(curve, n, lib=lib)
        dom4 = Radix2EvaluationDomain(curve, 4 * n, lib=lib)
        names = [s for s in SELECTORS if s not in CUSTOM_SELECTORS or s in custom]
        sig_names = ["left_sigma", "right_sigma", "out_sigma", "fourth_sigma"]
        # residents: (8 selectors + 4 sigmas) polys + 4 tables + q_lookup evals + scratch ; 14 vectors of 4n
        # + the per-proof working set of `prove` (24 n-vectors, 11 4n-vectors, openings)
        arena = Arena(lib, (len(names) + 4 + 4 + 1 + 6 + 36) * n + (len(names) + 4 + 2 + 12) * 4 * n,
                      torch_device=self.arena_device)
        pk = ProverKey(curve=curve, n=n, arena=arena, dom=dom, dom4=dom4, custom=custom,
                       public_inputs={int(k): int(v) % p for k, v in circ.public_inputs.items() if int(v) % p})
        vals_mont = _mont_list(curve, circ.values)
        tmp = arena.alloc(n)
        # selector polynomials: ifft of the evaluation columns (preprocess.rs:304-340)
        for s in names:
            arena.upload(tmp, vals_mont[circ.selectors[s]])
            pk.poly[s] = arena.alloc(n)
            self._ntt(dom, NTT_IFFT, arena, tmp, n, pk.poly[s])
            if s == "q_lookup":
                pk.q_lookup_evals = arena.alloc(n)
                arena.upload(pk.q_lookup_evals, vals_mont[circ.selectors[s]])
        # sigma polynomials (permutation/mod.rs:140-213): roots w^i = fft(X), scaled by 1, K1, K2, K3
        x_poly = arena.alloc(n)
        arena.upload(x_poly, _mont_list(curve, [0, 1] if n > 1 else [0]))
        roots = arena.alloc(n)
        self._ntt(dom, NTT_FFT, arena, x_poly, min(2, n), roots)
        scaled = []
        sc_off = arena.alloc(n)
        for kk in (1, K1, K2, K3):
            self._lincomb(arena, [roots], [n], [kk], sc_off, n)
            scaled.append(arena.download(sc_off, n))
        scaled = np.stack(scaled)                                   # (4, n, 4)
        for c, s in enumerate(sig_names):
            lag = scaled[circ.sigma[c, :, 0], circ.sigma[c, :, 1]]
            arena.upload(tmp, lag)
            pk.poly[s] = arena.alloc(n)
            self._ntt(dom, NTT_IFFT, arena, tmp, n, pk.poly[s])
        # lookup table columns, padded with element 0 (lookup/multiset.rs:70-79)
        pk.tables = []
        tcols = [[row[i] for row in circ.table] for i in range(4)]
        for col in tcols:
            col = list(col) if col else [0]
            col = col + [col[0]] * (n - len(col))
            off = arena.alloc(n)
            arena.upload(off, _mont_list(curve, col))
            pk.tables.append(off)
        # 4n coset evaluations (preprocess.rs:145-210) incl. linear = coset_fft([0, 1]) and L1
        for s in names + sig_names:
            pk.ev4[s] = arena.alloc(4 * n)
            self._ntt(dom4, NTT_COSET_FFT, arena, pk.poly[s], n, pk.ev4[s])
        pk.ev4["linear"] = arena.alloc(4 * n)
        self._ntt(dom4, NTT_COSET_FFT, arena, x_poly, min(2, n), pk.ev4["linear"])
        l1 = np.zeros((n, 4), dtype=np.uint64)
        l1[0] = _mont(curve, 1)
        arena.upload(tmp, l1)
        l1_poly = arena.alloc(n)
        self._ntt(dom, NTT_IFFT, arena, tmp, n, l1_poly)
        pk.ev4["l1"] = arena.alloc(4 * n)
        self._ntt(dom4, NTT_COSET_FFT, arena, l1_poly, n, pk.ev4["l1"])
        # vanishing polynomial on the coset has period 4 (preprocess.rs:429-452)
        g_n = pow(FR_GENERATOR[curve], n, p)
        w4n = pow(self._root_of_unity(4 * n), n, p)
        vh = [(g_n * pow(w4n, i, p) - 1) % p for i in range(4)]
        pk.vh_inv = _mont_list(curve, [pow(v, -1, p) for v in vh])
        if commit_verifier_key:
            all_names = names + sig_names
            res = self._commit(arena, [pk.poly[s] for s in all_names], [n] * len(all_names))
            for s, (_, comp) in zip(all_names, res):
                pk.commitments[s] = comp
            for i, off in enumerate(pk.tables):
                self._ntt(dom, NTT_IFFT, arena, off, n, tmp)
                pk.commitments["table_%d" % (i + 1)] = self._commit(arena, [tmp], [n])[0][1]
        lib.check(lib.c.apb_dev_sync())
        return pk

    def load_prover_key(self, blob: bytes) -> ProverKey:
        """`ProverKey::deserialize_unchecked` (proof_system/widget/mod.rs:285-328; circuit.rs:264-268 takes such a
        key by value): builds the device-resident key from arkworks' serialized form instead of compiling the
        circuit.  Residents the reference recomputes per proof (L1 evaluations, q_lookup on the n-domain, 1/Z_H)
        are derived here once.  Public inputs are not part of the key: pass them to `prove`."""
        from . import serialize as ser
        curve, p, lib = self.curve, self.p, self.lib
        arr = ser.read_prover_key_arrays(curve, blob)
        n = arr["n"]
        field_id = 0 if curve == 0 else 2

        def mont(a):
            a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
            return lib.field_op(field_id, 3, a, None) if a.shape[0] else a

        custom = tuple(s for s in CUSTOM_SELECTORS if arr["poly"][s].shape[0])
        dom = Radix2EvaluationDomain(curve, n, lib=lib)
        dom4 = Radix2EvaluationDomain(curve, 4 * n, lib=lib)
        names = [s for s in SELECTORS if s not in CUSTOM_SELECTORS or s in custom]
        sig_names = ["left_sigma", "right_sigma", "out_sigma", "fourth_sigma"]
        arena = Arena(lib, (len(names) + 4 + 4 + 1 + 6 + 36) * n + (len(names) + 4 + 2 + 12) * 4 * n,
                      torch_device=self.arena_device)
        pk = ProverKey(curve=curve, n=n, arena=arena, dom=dom, dom4=dom4, custom=custom, public_inputs={})
        for s in names + sig_names:
            c = np.zeros((n, 4), dtype=np.uint64)
            m = mont(arr["poly"][s])
            c[: m.shape[0]] = m
            pk.poly[s] = arena.alloc(n)
            arena.upload(pk.poly[s], c)
            pk.ev4[s] = arena.alloc(4 * n)
            arena.upload(pk.ev4[s], mont(arr["ev4"][s]))
        pk.q_lookup_evals = arena.alloc(n)
        self._ntt(dom, NTT_FFT, arena, pk.poly["q_lookup"], n, pk.q_lookup_evals)
        pk.tables = []
        for t in arr["tables"]:
            off = arena.alloc(n)
            arena.upload(off, mont(t))
            pk.tables.append(off)
        pk.ev4["linear"] = arena.alloc(4 * n)
        arena.upload(pk.ev4["linear"], mont(arr["linear"]))
        tmp = arena.alloc(n)
        l1 = np.zeros((n, 4), dtype=np.uint64)
        l1[0] = _mont(curve, 1)
        arena.upload(tmp, l1)
        l1_poly = arena.alloc(n)
        self._ntt(dom, NTT_IFFT, arena, tmp, n, l1_poly)
        pk.ev4["l1"] = arena.alloc(4 * n)
        self._ntt(dom4, NTT_COSET_FFT, arena, l1_poly, n, pk.ev4["l1"])
        vh = enc.limbs_to_ints(arr["v_h"][:4])
        pk.vh_inv = _mont_list(curve, [pow(v, -1, p) for v in vh])
        lib.check(lib.c.apb_dev_sync())
        return pk

    def _root_of_unity(self, size: int) -> int:
        p = self.p
        adicity = 32 if self.curve == 0 else 47
        root = pow(FR_GENERATOR[self.curve], (p - 1) >> adicity, p)
        return pow(root, 1 << (adicity - (size.bit_length() - 1)), p)

    # ---- prove -----------------------------------------------------------------------------
    def upload_wires(self, pk: ProverKey, wires_mont) -> int:
        """copies the (4, n, 4) witness columns into the key's arena (stays until the key is closed);
        `wires_mont` may be a numpy array or an int host pointer (e.g. pinned memory)"""
        off = pk.arena.alloc(4 * pk.n)
        self._upload_wires(pk.arena, off, wires_mont, pk.n)
        return off

    def _upload_wires(self, arena, off, wires_mont, n):
        if isinstance(wires_mont, int):
            self.lib.check(self.lib.c.apb_dev_upload(arena.ptr(off), wires_mont, 4 * n * 32))
        else:
            arena.upload(off, np.ascontiguousarray(wires_mont).reshape(4 * n, 4))

    def prove(self, pk: ProverKey, wires_mont, transcript_label: bytes = b"ark", faithful: bool = True,
              trace: dict | None = None, wires_resident: int | None = None, public_inputs: dict | None = None):
        """wires_mont: (4, n, 4) uint64 Montgomery wire values (w_l, w_r, w_o, w_4), padded to n.
        `faithful`: also issue the 14 commitments of prover.rs:579,606 whose results the reference
        discards (SonicKZG10::open ignores them) so that the MSM count matches the reference's 29.
        `public_inputs`: {row: value} of THIS proof (the reference reads them from the composer per
        proof, prover.rs:182,392; the key holds structure only); default = the values of the circuit the
        key was compiled from."""
        curve, n, p, lib = self.curve, pk.n, self.p, self.lib
        N4 = 4 * n
        dom, dom4 = pk.dom, pk.dom4
        A = pk.arena
        mark = A.mark()
        T = trace if trace is not None else {}
        if public_inputs is None:
            pi = pk.public_inputs
        else:               # proof_system/pi.rs: zero values are not stored
            pi = {int(k): int(v) % p for k, v in public_inputs.items() if int(v) % p}
            if any(r < 0 or r >= n for r in pi):
                raise ApbError(1, "public input row outside the circuit")
        try:
            return self._prove(pk, wires_mont, transcript_label, faithful, T, A, dom, dom4, n, N4, p, curve, lib,
                               wires_resident, pi)
        finally:
            lib.c.apb_dev_sync()
            A.release(mark)

    def _prove(self, pk, wires_mont, label, faithful, T, A, dom, dom4, n, N4, p, curve, lib, wires_resident=None,
               public_inputs=None):
        public_inputs = pk.public_inputs if public_inputs is None else public_inputs
        tr = Transcript(lib, curve, label)
        # PublicInputs (proof_system/pi.rs) = BTreeMap<usize, F>: u64 length, (u64 row, element) in row order
        pi_ser = len(public_inputs).to_bytes(8, "little")
        for pos in sorted(public_inputs):
            pi_ser += pos.to_bytes(8, "little") + public_inputs[pos].to_bytes(32, "little")
        tr.append_bytes(b"pi", pi_ser)
        omega = self._root_of_unity(n)

        # -- round 1: wire polynomials (prover.rs:188-220)
        if wires_resident is not None:
            w_ev = wires_resident
        else:
            w_ev = A.alloc(4 * n)
            self._upload_wires(A, w_ev, wires_mont, n)
        w_poly = A.alloc(4 * n)
        self._ntt(dom, NTT_IFFT, A, w_ev, n, w_poly, batch=4)
        wp = [w_poly + c * n for c in range(4)]
        wv = [w_ev + c * n for c in range(4)]
        w_comm = self._commit(A, wp, [n] * 4)
        for lab, (_, comp) in zip((b"w_l", b"w_r", b"w_o", b"w_4"), w_comm):
            tr.append_bytes(lab, comp)

        # -- round 2: lookup (prover.rs:225-321)
        zeta = tr.challenge(b"zeta")
        tr.append_fr(b"zeta", zeta)
        t_comp = A.alloc(n)
        self._lincomb(A, pk.tables, [n] * 4, [1, zeta, zeta * zeta % p, pow(zeta, 3, p)], t_comp, n)
        table_poly = A.alloc(n)
        self._ntt(dom, NTT_IFFT, A, t_comp, n, table_poly)
        f_comp = A.alloc(n)
        zeta_m = _mont(curve, zeta)
        lib.check(lib.c.apb_plonk_lookup_f(curve, A.ptr(pk.q_lookup_evals), A.ptr(wv[0]), A.ptr(wv[1]), A.ptr(wv[2]),
                                           A.ptr(wv[3]), A.ptr(t_comp), zeta_m.ctypes.data, A.ptr(f_comp), n))
        f_poly = A.alloc(n)
        self._ntt(dom, NTT_IFFT, A, f_comp, n, f_poly)
        # combine_split needs no further challenge, so f, h1, h2 are committed in one batched pass
        # (the reference issues three PC::commit calls, prover.rs:290,313,316); transcript order kept
        h_ev = A.alloc(2 * n)
        lib.check(lib.c.apb_plonk_combine_split(curve, A.ptr(t_comp), A.ptr(f_comp), n, A.ptr(h_ev), A.ptr(h_ev + n)))
        h_poly = A.alloc(2 * n)
        self._ntt(dom, NTT_IFFT, A, h_ev, n, h_poly, batch=2)
        h1_poly, h2_poly = h_poly, h_poly + n
        f_comm, h1_comm, h2_comm = self._commit(A, [f_poly, h1_poly, h2_poly], [n] * 3)
        tr.append_bytes(b"f", f_comm[1])
        tr.append_bytes(b"h1", h1_comm[1])
        tr.append_bytes(b"h2", h2_comm[1])

        # -- round 3: permutation polynomials (prover.rs:326-392)
        beta = tr.challenge(b"beta"); tr.append_fr(b"beta", beta)
        gamma = tr.challenge(b"gamma"); tr.append_fr(b"gamma", gamma)
        delta = tr.challenge(b"delta"); tr.append_fr(b"delta", delta)
        epsilon = tr.challenge(b"epsilon"); tr.append_fr(b"epsilon", epsilon)
        if len({beta, gamma, delta, epsilon}) != 4:
            raise ApbError(1, "challenges must be different")
        sig_names = ("left_sigma", "right_sigma", "out_sigma", "fourth_sigma")
        sig_ev = A.alloc(4 * n)
        for c, s in enumerate(sig_names):                                    # 4 x domain.fft (mod.rs:671-674)
            self._ntt(dom, NTT_FFT, A, pk.poly[s], n, sig_ev + c * n)
        z_ev = A.alloc(n)
        wires_p = (C.c_void_p * 4)(*[A.ptr(o) for o in wv])
        sig_p = (C.c_void_p * 4)(*[A.ptr(sig_ev + c * n) for c in range(4)])
        beta_m, gamma_m, delta_m, eps_m = (_mont(curve, v) for v in (beta, gamma, delta, epsilon))
        lib.check(lib.c.apb_plonk_perm_z(dom._h, wires_p, sig_p, beta_m.ctypes.data, gamma_m.ctypes.data, A.ptr(z_ev)))
        z_poly = A.alloc(n)
        self._ntt(dom, NTT_IFFT, A, z_ev, n, z_poly)
        z2_ev = A.alloc(n)
        lib.check(lib.c.apb_plonk_lookup_z2(dom._h, A.ptr(f_comp), A.ptr(t_comp), A.ptr(h_ev), A.ptr(h_ev + n),
                                            delta_m.ctypes.data, eps_m.ctypes.data, A.ptr(z2_ev)))
        z2_poly = A.alloc(n)
        self._ntt(dom, NTT_IFFT, A, z2_ev, n, z2_poly)
        z_comm, z2_comm = self._commit(A, [z_poly, z2_poly], [n] * 2)       # prover.rs:362,388 in one pass
        tr.append_bytes(b"z", z_comm[1])                                    # z2 is not absorbed (prover.rs:387-389)

        # -- round 4: quotient (prover.rs:398-475, quotient_poly.rs:34-178)
        alpha = tr.challenge(b"alpha"); tr.append_fr(b"alpha", alpha)
        range_sep = tr.challenge(b"range separation challenge"); tr.append_fr(b"range seperation challenge", range_sep)
        logic_sep = tr.challenge(b"logic separation challenge"); tr.append_fr(b"logic seperation challenge", logic_sep)
        fixed_sep = tr.challenge(b"fixed base separation challenge"); tr.append_fr(b"fixed base separation challenge", fixed_sep)
        var_sep = tr.challenge(b"variable base separation challenge"); tr.append_fr(b"variable base separation challenge", var_sep)
        lookup_sep = tr.challenge(b"lookup separation challenge"); tr.append_fr(b"lookup separation challenge", lookup_sep)
        pi_poly = None
        if public_inputs:                                                    # prover.rs:392 (pi_poly = ifft of the PI column)
            pi_col = np.zeros((n, 4), dtype=np.uint64)
            rows = sorted(public_inputs)
            pi_col[rows] = _mont_list(curve, [public_inputs[r] for r in rows])
            pi_poly = A.alloc(n)
            A.upload(pi_poly, pi_col)
            self._ntt(dom, NTT_IFFT, A, pi_poly, n, pi_poly)
        mark4 = A.mark()
        srcs = [z_poly, wp[0], wp[1], wp[2], wp[3], z2_poly, f_poly, table_poly, h1_poly, h2_poly]
        ev_names = ["z", "wl", "wr", "wo", "w4", "z2", "f", "table", "h1", "h2"]
        if pi_poly is not None:
            srcs.append(pi_poly)
            ev_names.append("pi")
        ev = A.alloc(len(srcs) * N4)
        # one proof over several GPUs: the coset FFTs are independent, so rank r transforms polynomials
        # [r * per, (r + 1) * per) and the evaluation vectors are all-gathered in place over NVLink; the
        # K mod world left-over polynomials are transformed by every rank
        com = self.committer
        split = com is not None and com.world > 1 and A.tensor is not None and N4 % com.world == 0
        per = len(srcs) // com.world if split else 0
        for k, s in enumerate(srcs):                                         # 10 coset FFTs on 4n (quotient_poly.rs:74-120)
            if k >= per * (com.world if split else 0) or k // per == com.rank:
                self._ntt(dom4, NTT_COSET_FFT, A, s, n, ev + k * N4)
        if per:
            com.all_gather_inplace(A, ev, per * N4)
        E = {nm: ev + k * N4 for k, nm in enumerate(ev_names)}
        q_ev = A.alloc(N4)
        order = [E["wl"], E["wr"], E["wo"], E["w4"], E["z"], E["z2"], E["f"], E["table"], E["h1"], E["h2"], E.get("pi"),
                 pk.ev4["q_m"], pk.ev4["q_l"], pk.ev4["q_r"], pk.ev4["q_o"], pk.ev4["q_4"], pk.ev4["q_c"], pk.ev4["q_arith"],
                 pk.ev4["q_lookup"], pk.ev4["left_sigma"], pk.ev4["right_sigma"], pk.ev4["out_sigma"], pk.ev4["fourth_sigma"],
                 pk.ev4["linear"], pk.ev4["l1"]] + [pk.ev4.get(s) for s in CUSTOM_SELECTORS]
        ptrs = (C.c_void_p * 29)(*[A.ptr(o) if o is not None else None for o in order])
        scal = _mont_list(curve, [alpha, beta, gamma, delta, epsilon, zeta, lookup_sep, K1, K2, K3,
                                  range_sep, logic_sep, fixed_sep, var_sep, gates.EMBEDDED_A[curve], gates.EMBEDDED_D[curve]])
        if split:                                                            # every rank evaluates its slice of the coset
            chunk = N4 // com.world
            lib.check(lib.c.apb_plonk_quotient_range(curve, ptrs, scal.ctypes.data, pk.vh_inv.ctypes.data, A.ptr(q_ev), N4,
                                                     com.rank * chunk, chunk))
            com.all_gather_inplace(A, q_ev, chunk)
        else:
            lib.check(lib.c.apb_plonk_quotient_full(curve, ptrs, scal.ctypes.data, pk.vh_inv.ctypes.data, A.ptr(q_ev), N4))
        # the evaluation vectors are dead: t_poly reuses their space (the stream is in order, so
        # the coset_ifft reads q_ev before anything enqueued later can overwrite it)
        A.release(mark4)
        t_poly = A.alloc(N4)
        self._ntt(dom4, NTT_COSET_IFFT, A, q_ev, N4, t_poly)
        t_off = [t_poly + k * n for k in range(4)]
        t_comm = self._commit(A, t_off, [n] * 4)
        for lab, (_, comp) in zip((b"t_1", b"t_2", b"t_3", b"t_4"), t_comm):
            tr.append_bytes(lab, comp)

        # -- round 5: linearisation (prover.rs:480-554, linearisation_poly.rs:164-349)
        zc = tr.challenge(b"z"); tr.append_fr(b"z", zc)
        zw = zc * omega % p
        P = pk.poly
        ev_offs = [wp[0], wp[1], wp[2], wp[3], P["left_sigma"], P["right_sigma"], P["out_sigma"], P["q_arith"], P["q_lookup"],
                   P["q_c"], P["q_l"], P["q_r"], h1_poly, h2_poly, f_poly, table_poly,
                   z_poly, wp[0], wp[1], wp[3], z2_poly, h1_poly, table_poly]
        pts = [zc] * 16 + [zw] * 7
        (a_e, b_e, c_e, d_e, left_e, right_e, out_e, q_arith_e, q_lookup_e, q_c_e, q_l_e, q_r_e, h1_e, h2_e, f_e, table_e,
         perm_e, a_next, b_next, d_next, z2_next, h1_next, table_next) = self._eval(A, ev_offs, [n] * len(ev_offs), pts)
        vanishing = (pow(zc, n, p) - 1) % p
        z_n = (vanishing + 1) % p
        l1_z = vanishing * pow(n * (zc - 1) % p, -1, p) % p
        opd = (1 + delta) % p
        eopd = epsilon * opd % p
        lsq = lookup_sep * lookup_sep % p
        lcu = lsq * lookup_sep % p
        comp_tuple = (((d_e * zeta + c_e) * zeta + b_e) * zeta + a_e) % p
        bz = beta * zc % p
        pa = (a_e + bz + gamma) * (b_e + K1 * bz + gamma) % p * (c_e + K2 * bz + gamma) % p * (d_e + K3 * bz + gamma) % p * alpha % p
        pb = (a_e + beta * left_e + gamma) * (b_e + beta * right_e + gamma) % p * (c_e + beta * out_e + gamma) % p \
            * (beta * perm_e % p) % p * alpha % p
        terms = [
            (P["q_m"], a_e * b_e % p * q_arith_e), (P["q_l"], a_e * q_arith_e), (P["q_r"], b_e * q_arith_e),
            (P["q_o"], c_e * q_arith_e), (P["q_4"], d_e * q_arith_e), (P["q_c"], q_arith_e),
            (z_poly, pa + l1_z * alpha % p * alpha), (P["fourth_sigma"], -pb),
            (P["q_lookup"], (comp_tuple - f_e) * lookup_sep),
            (z2_poly, opd * (epsilon + f_e) % p * (eopd + table_e + delta * table_next) % p * lsq + l1_z * lcu),
            (h1_poly, (-z2_next) * lsq % p * (eopd + h2_e + delta * h1_next)),
            (t_off[0], -vanishing), (t_off[1], -vanishing * z_n), (t_off[2], -vanishing * z_n % p * z_n),
            (t_off[3], -vanishing * pow(z_n, 3, p)),
        ]
        # custom gates: selector polynomial * constraints(evaluations) (linearisation_poly.rs:382-410)
        w_e, nxt_e = (a_e, b_e, c_e, d_e), (a_next, b_next, d_next)
        for s in pk.custom:
            if s == "q_range":
                sc = gates.range_scalar(range_sep, w_e, nxt_e, p)
            elif s == "q_logic":
                sc = gates.logic_scalar(logic_sep, w_e, nxt_e, q_c_e, p)
            elif s == "q_fixed_group_add":
                sc = gates.fixed_base_scalar(fixed_sep, w_e, nxt_e, q_l_e, q_r_e, q_c_e, curve, p)
            else:
                sc = gates.curve_add_scalar(var_sep, w_e, nxt_e, curve, p)
            terms.append((P[s], sc))
        lin_poly = A.alloc(n)
        self._lincomb(A, [o for o, _ in terms], [n] * len(terms), [s % p for _, s in terms], lin_poly, n)
        for lab, v in ((b"a_eval", a_e), (b"b_eval", b_e), (b"c_eval", c_e), (b"d_eval", d_e), (b"left_sig_eval", left_e),
                       (b"right_sig_eval", right_e), (b"out_sig_eval", out_e), (b"perm_eval", perm_e), (b"f_eval", f_e),
                       (b"q_lookup_eval", q_lookup_e), (b"lookup_perm_eval", z2_next), (b"h_1_eval", h1_e),
                       (b"h_1_next_eval", h1_next), (b"h_2_eval", h2_e)):
            tr.append_fr(lab, v)
        custom = [("q_arith_eval", q_arith_e), ("q_c_eval", q_c_e), ("q_l_eval", q_l_e), ("q_r_eval", q_r_e),
                  ("a_next_eval", a_next), ("b_next_eval", b_next), ("d_next_eval", d_next)]
        for lab, v in custom:
            tr.append_fr(lab.encode(), v)

        # -- openings (prover.rs:563-618; sonic_pc::open: p = sum challenge^i p_i, witness = p / (X - z))
        aw_challenge = tr.challenge(b"aggregate_witness")
        saw_challenge = tr.challenge(b"aggregate_witness")          # prover.rs:593: no absorb between the two challenges
        aw_polys = [lin_poly, P["left_sigma"], P["right_sigma"], P["out_sigma"], f_poly, h2_poly, table_poly] + wp
        saw_polys = [z_poly, wp[0], wp[1], wp[3], h1_poly, z2_poly, table_poly]
        comb = A.alloc(2 * n)
        wit = A.alloc(2 * n)
        zc_m, zw_m = _mont(curve, zc), _mont(curve, zw)
        self._lincomb(A, aw_polys, [n] * len(aw_polys), [pow(aw_challenge, i, p) for i in range(len(aw_polys))], comb, n)
        lib.check(lib.c.apb_poly_divide_linear(curve, A.ptr(comb), n, zc_m.ctypes.data, A.ptr(wit)))
        self._lincomb(A, saw_polys, [n] * 7, [pow(saw_challenge, i, p) for i in range(7)], comb + n, n)
        lib.check(lib.c.apb_poly_divide_linear(curve, A.ptr(comb + n), n, zw_m.ctypes.data, A.ptr(wit + n)))
        # The saw opening does not depend on the aw opening, so everything round 5 commits goes through ONE
        # batched pass: PC::commit(aw_polys) and PC::commit(saw_polys) (prover.rs:579,606 - 7 + 7 commitments
        # that SonicKZG10::open then ignores; issued when `faithful`) and the two opening MSMs (prover.rs:582,609).
        if faithful:
            res = self._commit(A, aw_polys[:7] + [wit] + saw_polys + [wit + n], [n] * 7 + [n - 1] + [n] * 7 + [n - 1])
            aw_open, saw_open = res[7], res[15]
        else:
            aw_open, saw_open = self._commit(A, [wit, wit + n], [n - 1, n - 1])

        if T is not None and T.get("want"):
            T.update(zeta=zeta, beta=beta, gamma=gamma, delta=delta, epsilon=epsilon, alpha=alpha, lookup_sep=lookup_sep,
                     z_challenge=zc, aw_challenge=aw_challenge, saw_challenge=saw_challenge,
                     w_polys=[enc.fr_from_mont(curve, A.download(o, n)) for o in wp],
                     t_poly=enc.fr_from_mont(curve, A.download(t_poly, N4)),
                     lin_poly=enc.fr_from_mont(curve, A.download(lin_poly, n)),
                     z_poly=enc.fr_from_mont(curve, A.download(z_poly, n)),
                     z2_poly=enc.fr_from_mont(curve, A.download(z2_poly, n)),
                     h1=enc.fr_from_mont(curve, A.download(h_ev, n)), h2=enc.fr_from_mont(curve, A.download(h_ev + n, n)))

        comms = [w_comm[0], w_comm[1], w_comm[2], w_comm[3], z_comm, f_comm, h1_comm, h2_comm, z2_comm,
                 t_comm[0], t_comm[1], t_comm[2], t_comm[3]]
        out = b"".join(c[1] for c in comms)
        out += aw_open[1] + b"\x00" + saw_open[1] + b"\x00"
        for v in (a_e, b_e, c_e, d_e, left_e, right_e, out_e, perm_e,
                  q_lookup_e, z2_next, h1_e, h1_next, h2_e, f_e, table_e, table_next):
            out += int(v).to_bytes(32, "little")
        out += len(custom).to_bytes(8, "little")
        for lab, v in custom:
            b = lab.encode()
            out += len(b).to_bytes(8, "little") + b + int(v).to_bytes(32, "little")
        return out


def wires_to_mont(circ: CircuitArrays) -> np.ndarray:
    """(4, n, 4) Montgomery wire values of a CircuitArrays instance"""
    vals = _mont_list(circ.curve, circ.values)
    return vals[circ.wires]
