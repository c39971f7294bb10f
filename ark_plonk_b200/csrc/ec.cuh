// G1 group law for short-Weierstrass curves with a = 0 (BLS12-381, BLS12-377) in extended
// Jacobian "XYZZ" coordinates:  x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2, identity <=> ZZ == 0.
//
// Stands in for ark-ec 0.3 GroupProjective::add_assign_mixed / add_assign / double_in_place as
// used by VariableBaseMSM (plonk-core/src/commitment.rs:45).  arkworks uses Jacobian
// coordinates; the projective representative is not observable (every result leaves the
// library as a normalised affine point), so XYZZ (8M+2S mixed add instead of 7M+4S) is free
// to choose.
#pragma once
#include "arith.cuh"

namespace apb {

template <class FQ>
struct Affine {        // 2*N u32, Montgomery; (0, 0) encodes the point at infinity
    Fp<FQ> x, y;
    APB_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
};

template <class FQ>
struct XYZZ {
    typedef Fp<FQ> F;
    F x, y, zz, zzz;

    APB_HD static XYZZ identity() {
        XYZZ r;
        r.x = F::zero(); r.y = F::zero(); r.zz = F::zero(); r.zzz = F::zero();
        return r;
    }
    APB_HD bool is_identity() const { return zz.is_zero(); }

    APB_HD static XYZZ from_affine(const Affine<FQ>& p) {
        XYZZ r;
        if (p.is_inf()) return identity();
        r.x = p.x; r.y = p.y; r.zz = F::one(); r.zzz = F::one();
        return r;
    }

    // 2*(x, y) for an affine point (mdbl-2008-s-1), y != 0 on these prime-order groups
    APB_HD static XYZZ dbl_affine(const F& px, const F& py) {
        XYZZ r;
        F U = py.dbl();
        F V = U.sqr();
        F W = U * V;
        F S = px * V;
        F X2 = px.sqr();
        F M = X2.dbl() + X2;
        r.x = M.sqr() - S.dbl();
        r.y = M * (S - r.x) - W * py;
        r.zz = V;
        r.zzz = W;
        return r;
    }

    // dbl-2008-s-1
    APB_HD XYZZ dbl() const {
        if (is_identity()) return *this;
        XYZZ r;
        F U = y.dbl();
        F V = U.sqr();
        F W = U * V;
        F S = x * V;
        F X2 = x.sqr();
        F M = X2.dbl() + X2;
        r.x = M.sqr() - S.dbl();
        r.y = M * (S - r.x) - W * y;
        r.zz = V * zz;
        r.zzz = W * zzz;
        return r;
    }

    // this += (px, +-py)   (madd-2008-s); (px,py) must not be infinity
    APB_HD void add_affine(const F& px, const F& py) {
        if (is_identity()) {
            x = px; y = py; zz = F::one(); zzz = F::one();
            return;
        }
        F U2 = px * zz;
        F S2 = py * zzz;
        F Pp = U2 - x;
        F R = S2 - y;
        if (Pp.is_zero()) {
            if (R.is_zero()) *this = dbl_affine(px, py);
            else *this = identity();
            return;
        }
        F PP = Pp.sqr();
        F PPP = Pp * PP;
        F Q = x * PP;
        F X3 = R.sqr() - PPP - Q.dbl();
        y = R * (Q - X3) - y * PPP;
        x = X3;
        zz = zz * PP;
        zzz = zzz * PPP;
    }

    // this += o   (add-2008-s)
    APB_HD void add(const XYZZ& o) {
        if (o.is_identity()) return;
        if (is_identity()) { *this = o; return; }
        F U1 = x * o.zz;
        F U2 = o.x * zz;
        F S1 = y * o.zzz;
        F S2 = o.y * zzz;
        F Pp = U2 - U1;
        F R = S2 - S1;
        if (Pp.is_zero()) {
            if (R.is_zero()) *this = dbl();
            else *this = identity();
            return;
        }
        F PP = Pp.sqr();
        F PPP = Pp * PP;
        F Q = U1 * PP;
        F X3 = R.sqr() - PPP - Q.dbl();
        y = R * (Q - X3) - S1 * PPP;
        x = X3;
        zz = zz * o.zz * PP;
        zzz = zzz * o.zzz * PPP;
    }

    APB_HD XYZZ neg() const {
        XYZZ r = *this;
        r.y = y.neg();
        return r;
    }
};

}  // namespace apb
