// GROUNDWORK (lever 2) -- batched affine addition.  Not used by a product kernel yet.
//
// out[i] = P[i] + Q[i] for a chunk of m independent pairs of AFFINE points with ONE field inversion (Montgomery's
// trick): 5M + 1S per addition + one inversion per chunk, against the 8M + 2S of the XYZZ mixed addition the bucket
// accumulation uses today.  The inversion is fp_inv's binary Euclid: shifts and additions on the ALU pipe, none
// of the multiplier that bounds the accumulation -- which is what makes per-thread chunks
// of a few dozen pairs worthwhile without any cross-thread cooperation.
//
//   pass 1 (forward):   d_i = x2 - x1   (2 y1 for a doubling; 1 when the pair needs no division)
//                       pre[i] = d_0 d_1 ... d_i
//   one inversion:      inv = 1 / pre[m-1]
//   pass 2 (backward):  1/d_i = inv * pre[i-1];  inv *= d_i
//                       lambda = (y2 - y1) / d_i          (3 x1^2 / (2 y1) for a doubling)
//                       x3 = lambda^2 - x1 - x2,  y3 = lambda (x1 - x3) - y1
// Special cases are resolved per pair without breaking the chain: inf + Q, P + inf, P + (-P) = inf, P + P.
// The bucket-sum plan built on it (rounds of pairwise additions inside each bucket, tested on the host as
// hc_bucket_sums_batch_affine in tests/host/hostcheck.cpp) is described in DESIGN.md.
#pragma once
#include "ec.cuh"

namespace b200zk {

enum : uint8_t { BA_ADD = 0, BA_DBL = 1, BA_COPY_P = 2, BA_COPY_Q = 3, BA_INF = 4 };

// classification of one pair and the factor it contributes to the product chain
template <class F>
HD uint8_t ba_classify(const Affine<F>& p, const Affine<F>& q, F& d) {
    d = F::one();
    if (p.is_inf()) return BA_COPY_Q;
    if (q.is_inf()) return BA_COPY_P;
    if (p.x == q.x) {
        if (p.y == q.y && !p.y.is_zero()) {
            d = fp_dbl(p.y);
            return BA_DBL;
        }
        return BA_INF;  // opposite points (or a 2-torsion point doubled)
    }
    d = fp_sub(q.x, p.x);
    return BA_ADD;
}

// pre: m field elements of scratch.  out may alias p or q.
template <class F>
HD void ec_batch_add_affine(const Affine<F>* p, const Affine<F>* q, Affine<F>* out, int m, F* pre) {
    if (m <= 0) return;
    F acc = F::one();
    for (int i = 0; i < m; i++) {
        F d;
        ba_classify(p[i], q[i], d);
        acc = fp_mul(acc, d);
        pre[i] = acc;
    }
    F inv = fp_inv(acc);  // every d_i != 0, so acc != 0
    for (int i = m - 1; i >= 0; i--) {
        F d;
        const uint8_t kind = ba_classify(p[i], q[i], d);
        const F dinv = i ? fp_mul(inv, pre[i - 1]) : inv;
        inv = fp_mul(inv, d);
        const Affine<F> a = p[i], b = q[i];
        Affine<F> r;
        switch (kind) {
            case BA_COPY_P: r = a; break;
            case BA_COPY_Q: r = b; break;
            case BA_INF: r = Affine<F>::inf(); break;
            default: {
                F num;
                if (kind == BA_DBL) {
                    const F xx = fp_sqr(a.x);
                    num = fp_add(fp_dbl(xx), xx);
                } else {
                    num = fp_sub(b.y, a.y);
                }
                const F lam = fp_mul(num, dinv);
                r.x = fp_sub(fp_sub(fp_sqr(lam), a.x), b.x);
                r.y = fp_sub(fp_mul(lam, fp_sub(a.x, r.x)), a.y);
            }
        }
        out[i] = r;
    }
}

}  // namespace b200zk
