// Test-only shared library: runs the PORTABLE (host) path of zk-apps_b200/csrc/{field,ec}.cuh so
// the limb algorithms and EC formulas the CUDA kernels share can be diffed against the Python
// oracle without a GPU.  Not part of the product; never loaded by it.
#include <cstring>
#include "../../zk-apps_b200/csrc/ec.cuh"
#include "../../zk-apps_b200/csrc/pairing.cuh"
#include "../../zk-apps_b200/csrc/verify.cuh"
#include "../../zk-apps_b200/csrc/glv.cuh"
#include "../../zk-apps_b200/csrc/field_dfma.cuh"
#include "../../zk-apps_b200/csrc/ec_batch_affine.cuh"
#include <vector>
#include <algorithm>
using namespace b200zk;

template <class F> static void field_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        F x, y, r;
        memcpy(&x, a + i * sizeof(F), sizeof(F));
        if (b) memcpy(&y, b + i * sizeof(F), sizeof(F)); else y = x;
        switch (op) {
            case 0: r = fp_add(x, y); break;
            case 1: r = fp_sub(x, y); break;
            case 2: r = fp_mul(x, y); break;
            case 3: r = fp_sqr(x); break;
            case 4: r = fp_inv(x); break;
            case 8: if constexpr (sizeof(F) == sizeof(Fq2)) r = fp_inv(x); else r = fp_inv_uniform(x); break;  // groundwork: branch-uniform inversion
            case 9: if constexpr (sizeof(F) == sizeof(Fq2)) r = fp_inv(x); else r = fp_inv_binary(x); break;  // the binary Euclid (fp_inv itself is the safegcd inversion)
            default: r = x;
        }
        memcpy(out + i * sizeof(F), &r, sizeof(F));
    }
}
template <class C> static void conv(int op, const uint8_t* a, uint8_t* out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        Fp<C> x; memcpy(&x, a + i * sizeof(x), sizeof(x));
        x = op == 5 ? fp_to_mont(x) : fp_from_mont(x);
        memcpy(out + i * sizeof(x), &x, sizeof(x));
    }
}
template <class F> static void msm_naive(const uint8_t* bases, const uint8_t* scalars, size_t n, uint8_t* out) {
    XYZZ<F> acc = XYZZ<F>::inf();
    for (size_t i = 0; i < n; i++) {
        Affine<F> p; memcpy(&p, bases + i * sizeof(p), sizeof(p));
        uint32_t k[8]; memcpy(k, scalars + 32 * i, 32);
        XYZZ<F> t = ec_mul_scalar(XYZZ<F>::from_affine(p), k);
        ec_add(acc, t);
    }
    Affine<F> r = ec_to_affine(acc);
    memcpy(out, &r, sizeof(r));
}
// sum of +-points with mixed additions, in order (exercises every ec_madd special case)
template <class F> static void madd_chain(const uint8_t* pts, const uint8_t* neg, size_t n, uint8_t* out) {
    XYZZ<F> acc = XYZZ<F>::inf();
    for (size_t i = 0; i < n; i++) {
        Affine<F> p; memcpy(&p, pts + i * sizeof(p), sizeof(p));
        ec_madd(acc, p, neg && neg[i]);
    }
    Affine<F> r = ec_to_affine(acc);
    memcpy(out, &r, sizeof(r));
}
// groundwork for the batched-affine bucket accumulation (ec_batch_affine.cuh): per-bucket sums by rounds of pairwise
// additions inside each bucket, every round one flat array of pairs cut into chunks of `chunk` pairs with one inversion
// each -- the schedule a GPU kernel would run (thread = chunk)
template <class F>
static void bucket_sums_batch_affine(const uint8_t* points, size_t n, const uint32_t* ids, uint32_t n_buckets, int chunk,
                                     uint8_t* out) {
    std::vector<Affine<F>> cur(n);
    if (n) memcpy(cur.data(), points, n * sizeof(Affine<F>));
    std::vector<size_t> off(n_buckets + 1, 0);
    for (size_t i = 0; i < n; i++) off[ids[i] + 1]++;
    for (uint32_t b = 0; b < n_buckets; b++) off[b + 1] += off[b];
    for (;;) {
        bool more = false;
        std::vector<size_t> noff(n_buckets + 1, 0);
        for (uint32_t b = 0; b < n_buckets; b++) {
            const size_t sz = off[b + 1] - off[b];
            noff[b + 1] = noff[b] + (sz + 1) / 2;
            if (sz > 1) more = true;
        }
        if (!more) break;
        const size_t total = noff[n_buckets];
        std::vector<Affine<F>> P(total), Q(total), R(total);
        for (uint32_t b = 0; b < n_buckets; b++) {
            const size_t sz = off[b + 1] - off[b];
            for (size_t i = 0; i < (sz + 1) / 2; i++) {
                P[noff[b] + i] = cur[off[b] + 2 * i];
                Q[noff[b] + i] = 2 * i + 1 < sz ? cur[off[b] + 2 * i + 1] : Affine<F>::inf();  // odd one out: P + inf
            }
        }
        std::vector<F> pre((size_t)chunk);
        for (size_t c = 0; c < total; c += (size_t)chunk) {
            const int m = (int)std::min((size_t)chunk, total - c);
            ec_batch_add_affine(P.data() + c, Q.data() + c, R.data() + c, m, pre.data());
        }
        cur.swap(R);
        off.swap(noff);
    }
    for (uint32_t b = 0; b < n_buckets; b++) {
        const Affine<F> r = off[b + 1] > off[b] ? cur[off[b]] : Affine<F>::inf();
        memcpy(out + (size_t)b * sizeof(Affine<F>), &r, sizeof(Affine<F>));
    }
}

template <class F>
static void batch_add_affine(const uint8_t* p, const uint8_t* q, size_t n, int chunk, uint8_t* out) {
    std::vector<Affine<F>> P(n), Q(n), R(n);
    memcpy(P.data(), p, n * sizeof(Affine<F>));
    memcpy(Q.data(), q, n * sizeof(Affine<F>));
    std::vector<F> pre((size_t)chunk);
    for (size_t c = 0; c < n; c += (size_t)chunk) {
        const int m = (int)std::min((size_t)chunk, n - c);
        ec_batch_add_affine(P.data() + c, Q.data() + c, R.data() + c, m, pre.data());
    }
    memcpy(out, R.data(), n * sizeof(Affine<F>));
}

extern "C" {
int hc_field_op(int field, int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
    if (op == 5 || op == 6) { if (field == 0) conv<FrCfg>(op, a, out, n); else conv<FqCfg>(op, a, out, n); return 0; }
    if (field == 0) field_op<Fr>(op, a, b, out, n);
    else if (field == 1) field_op<Fq>(op, a, b, out, n);
    else field_op<Fq2>(op, a, b, out, n);
    return 0;
}
int hc_msm_naive(int group, const uint8_t* bases, const uint8_t* scalars, size_t n, uint8_t* out) {
    if (group == 1) msm_naive<Fq>(bases, scalars, n, out); else msm_naive<Fq2>(bases, scalars, n, out);
    return 0;
}
int hc_madd_chain(int group, const uint8_t* pts, const uint8_t* neg, size_t n, uint8_t* out) {
    if (group == 1) madd_chain<Fq>(pts, neg, n, out); else madd_chain<Fq2>(pts, neg, n, out);
    return 0;
}
// ---- pairing.cuh: tower arithmetic, Miller loop, final exponentiation, wire format
// Fq12 = 576 B: c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2 (each Fq2 = c0 || c1, Montgomery)
int hc_fq12_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out) {
    Fq12 x, y, r;
    memcpy(&x, a, sizeof(x));
    if (b) memcpy(&y, b, sizeof(y)); else y = x;
    switch (op) {
        case 0: r = fq12_mul(x, y); break;
        case 1: r = fq12_sqr(x); break;
        case 2: r = fq12_inv(x); break;
        case 3: r = fq12_frob(x); break;
        case 4: r = fq12_conj(x); break;
        case 5: r = final_exponentiation(x); break;
        case 6: r = fq12_mul_by_014(x, y.c0.c0, y.c0.c1, y.c1.c1); break;  // sparse operand taken from y's slots 0, 1, 4
        default: r = x;
    }
    memcpy(out, &r, sizeof(r));
    return 0;
}
int hc_miller_loop(const uint8_t* g1, const uint8_t* g2, uint8_t* out) {
    Affine<Fq> p; Affine<Fq2> q;
    memcpy(&p, g1, sizeof(p)); memcpy(&q, g2, sizeof(q));
    Fq12 f = miller_loop(p, q);
    memcpy(out, &f, sizeof(f));
    return 0;
}
int hc_pairing(const uint8_t* g1, const uint8_t* g2, uint8_t* out) {
    Affine<Fq> p; Affine<Fq2> q;
    memcpy(&p, g1, sizeof(p)); memcpy(&q, g2, sizeof(q));
    Fq12 f = final_exponentiation(miller_loop(p, q));
    memcpy(out, &f, sizeof(f));
    return 0;
}
// returns the POINT_* status
int hc_decompress(int group, const uint8_t* in, uint8_t* out) {
    if (group == 1) { Affine<Fq> p = Affine<Fq>::inf(); int rc = g1_decompress(in, p); memcpy(out, &p, sizeof(p)); return rc; }
    Affine<Fq2> p = Affine<Fq2>::inf(); int rc = g2_decompress(in, p); memcpy(out, &p, sizeof(p)); return rc;
}
int hc_compress(int group, const uint8_t* in, uint8_t* out) {
    if (group == 1) { Affine<Fq> p; memcpy(&p, in, sizeof(p)); g1_compress(p, out); return 0; }
    Affine<Fq2> p; memcpy(&p, in, sizeof(p)); g2_compress(p, out); return 0;
}
int hc_in_subgroup_plain(int group, const uint8_t* in) {   // [r]P == O
    if (group == 1) { Affine<Fq> p; memcpy(&p, in, sizeof(p)); return ec_on_curve(p, g1_b()) && ec_in_subgroup_plain(p); }
    Affine<Fq2> p; memcpy(&p, in, sizeof(p)); return ec_on_curve(p, g2_b()) && ec_in_subgroup_plain(p);
}
int hc_in_subgroup(int group, const uint8_t* in) {
    if (group == 1) { Affine<Fq> p; memcpy(&p, in, sizeof(p)); return ec_on_curve(p, g1_b()) && ec_in_subgroup(p); }
    Affine<Fq2> p; memcpy(&p, in, sizeof(p)); return ec_on_curve(p, g2_b()) && ec_in_subgroup(p);
}
// the verifier's per-proof logic (verify.cuh) end to end: vk in b200zk_groth16_setup's vk_out layout; returns PROOF_*
int hc_verify(const uint8_t* vk_raw, uint32_t num_inputs, const uint8_t* proof, const uint8_t* public_inputs, int check_subgroup) {
    Affine<Fq> alpha; Affine<Fq2> g2s[3];
    memcpy(&alpha, vk_raw, 96); memcpy(g2s, vk_raw + 96, 576);
    std::vector<Affine<Fq>> abc(num_inputs);
    memcpy(abc.data(), vk_raw + 672, (size_t)num_inputs * 96);
    std::vector<Fr> x(num_inputs ? num_inputs - 1 : 0);
    if (!x.empty()) memcpy(x.data(), public_inputs, x.size() * 32);
    PreparedVk pv;
    prepare_vk(alpha, g2s[0], g2s[1], g2s[2], pv);
    return verify_one(pv, abc.data(), num_inputs - 1, proof, x.data(), check_subgroup != 0);
}
// GLV (glv.cuh): k (n x 32 B canonical LE) -> k1, k2 (n x 20 B LE each)
int hc_glv_split(const uint8_t* k, size_t n, uint8_t* k1, uint8_t* k2) {
    for (size_t i = 0; i < n; i++) {
        uint32_t kk[8], a[GLV_LIMBS], b[GLV_LIMBS];
        memcpy(kk, k + i * 32, 32);
        glv_split(kk, a, b);
        memcpy(k1 + i * 4 * GLV_LIMBS, a, 4 * GLV_LIMBS);
        memcpy(k2 + i * 4 * GLV_LIMBS, b, 4 * GLV_LIMBS);
    }
    return 0;
}
// phi(P) = (beta x, y) for n affine points (G1: 96 B, G2: 192 B, Montgomery)
int hc_glv_phi(int group, const uint8_t* in, size_t n, uint8_t* out) {
    for (size_t i = 0; i < n; i++) {
        if (group == 1) {
            Affine<Fq> p;
            memcpy(&p, in + i * 96, 96);
            p.x = glv_phi_x(p.x);
            memcpy(out + i * 96, &p, 96);
        } else {
            Affine<Fq2> p;
            memcpy(&p, in + i * 192, 192);
            p.x = glv_phi_x(p.x);
            memcpy(out + i * 192, &p, 192);
        }
    }
    return 0;
}
// experiment: the Fq Montgomery product on double-precision FMAs (field_dfma.cuh); n pairs of 48 B Montgomery words
int hc_fq_mul_dfma(const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        Fq x, y;
        memcpy(&x, a + i * 48, 48);
        memcpy(&y, b + i * 48, 48);
        const Fq r = dfma::fq_mul_dfma(x, y);
        memcpy(out + i * 48, &r, 48);
    }
    return 0;
}
int hc_batch_add_affine(int group, const uint8_t* p, const uint8_t* q, size_t n, int chunk, uint8_t* out) {
    if (group == 1) batch_add_affine<Fq>(p, q, n, chunk, out); else batch_add_affine<Fq2>(p, q, n, chunk, out);
    return 0;
}
// ids: bucket of every point, ascending
int hc_bucket_sums_batch_affine(int group, const uint8_t* points, size_t n, const uint32_t* ids, uint32_t n_buckets,
                                int chunk, uint8_t* out) {
    if (group == 1) bucket_sums_batch_affine<Fq>(points, n, ids, n_buckets, chunk, out);
    else bucket_sums_batch_affine<Fq2>(points, n, ids, n_buckets, chunk, out);
    return 0;
}
}
