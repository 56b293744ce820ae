// Test-only shared library: runs the PORTABLE (host) path of zk-apps_b200/csrc/{field,ec}.cuh so
// the limb algorithms and EC formulas the CUDA kernels share can be diffed against the Python
// oracle without a GPU.  Not part of the product; never loaded by it.
#include <cstring>
#include "../../zk-apps_b200/csrc/ec.cuh"
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
}
