// ORACLE (test infrastructure + reported CPU baseline; never on the product path).
//
// C++ restatement, 64-bit limbs + unsigned __int128, of the CPU algorithms arkworks 0.4 uses for
// the Groth16 path BASELINE.json names ([recall], SURVEY.md Appendix B -- ark-ff/ark-ec/ark-poly
// 0.4.2 and ark-bls12-381 0.4.0 are lockfile-only in the reference, shielder/contract/Cargo.lock:
// 195-281; ark-groth16 has no pin at all).  PARITY UNPINNED against the reference (no vectors
// exist there); pinned instead by tests/test_oracle_c.py against the Python big-int oracle
// (oracle/pyref) and the SURVEY Appendix A known answers.
//   * Montgomery CIOS field arithmetic (Fp<MontBackend<N>>)
//   * VariableBaseMSM::msm_bigint: c = 3 if n < 32 else ln_without_floats(n) + 2, unsigned digits,
//     2^c - 1 buckets per window, running-sum reduction, windows in parallel (rayon -> std::thread),
//     combined high to low with c doublings
//   * Radix2EvaluationDomain fft/ifft and the coset forms
//   * LibsnarkReduction witness map and create_proof_with_reduction(r, s)
// It is the "port" CPU baseline of bench.py: multi-threaded over the host cores (std::thread; the image has no OpenMP runtime).
#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

// rayon stand-in: dynamic-chunk parallel for over std::thread (the image has no OpenMP runtime)
static int g_threads = (int)std::thread::hardware_concurrency() > 0 ? (int)std::thread::hardware_concurrency() : 1;
template <class Fn>
static void parallel_for(size_t n, size_t chunk, Fn fn) {
    int nt = g_threads;
    if ((size_t)nt > (n + chunk - 1) / chunk) nt = (int)((n + chunk - 1) / chunk);
    if (nt <= 1) { for (size_t i = 0; i < n; i++) fn(i); return; }
    std::atomic<size_t> next(0);
    auto work = [&]() {
        for (;;) {
            size_t lo = next.fetch_add(chunk);
            if (lo >= n) break;
            size_t hi = lo + chunk < n ? lo + chunk : n;
            for (size_t i = lo; i < hi; i++) fn(i);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
}

typedef unsigned __int128 u128;

template <int N>
struct Params {
    uint64_t p[N];
    uint64_t inv;  // -p^-1 mod 2^64
    uint64_t one[N], r2[N];
};

template <int N, const Params<N>* P>
struct Fp {
    uint64_t v[N];
    static Fp zero() { Fp r; for (int i = 0; i < N; i++) r.v[i] = 0; return r; }
    static Fp one() { Fp r; for (int i = 0; i < N; i++) r.v[i] = P->one[i]; return r; }
    bool is_zero() const { uint64_t o = 0; for (int i = 0; i < N; i++) o |= v[i]; return o == 0; }
    bool operator==(const Fp& b) const { uint64_t o = 0; for (int i = 0; i < N; i++) o |= v[i] ^ b.v[i]; return o == 0; }
    static bool geq_p(const uint64_t* a) {
        for (int i = N - 1; i >= 0; i--) { if (a[i] > P->p[i]) return true; if (a[i] < P->p[i]) return false; }
        return true;
    }
    static void sub_p(uint64_t* a) {
        u128 br = 0;
        for (int i = 0; i < N; i++) { u128 d = (u128)a[i] - P->p[i] - br; a[i] = (uint64_t)d; br = (d >> 64) & 1; }
    }
    Fp operator+(const Fp& b) const {
        Fp r; u128 c = 0;
        for (int i = 0; i < N; i++) { c += (u128)v[i] + b.v[i]; r.v[i] = (uint64_t)c; c >>= 64; }
        if (c || geq_p(r.v)) sub_p(r.v);
        return r;
    }
    Fp operator-(const Fp& b) const {
        Fp r; u128 br = 0;
        for (int i = 0; i < N; i++) { u128 d = (u128)v[i] - b.v[i] - br; r.v[i] = (uint64_t)d; br = (d >> 64) & 1; }
        if (br) { u128 c = 0; for (int i = 0; i < N; i++) { c += (u128)r.v[i] + P->p[i]; r.v[i] = (uint64_t)c; c >>= 64; } }
        return r;
    }
    Fp neg() const { return is_zero() ? *this : zero() - *this; }
    Fp dbl() const { return *this + *this; }
    Fp operator*(const Fp& b) const {  // CIOS
        uint64_t t[N + 2];
        for (int i = 0; i < N + 2; i++) t[i] = 0;
        for (int i = 0; i < N; i++) {
            u128 c = 0;
            for (int j = 0; j < N; j++) { c += (u128)v[j] * b.v[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
            c += t[N]; t[N] = (uint64_t)c; t[N + 1] = (uint64_t)(c >> 64);
            uint64_t m = t[0] * P->inv;
            c = (u128)m * P->p[0] + t[0]; c >>= 64;
            for (int j = 1; j < N; j++) { c += (u128)m * P->p[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
            c += t[N]; t[N - 1] = (uint64_t)c; t[N] = t[N + 1] + (uint64_t)(c >> 64);
        }
        Fp r;
        for (int i = 0; i < N; i++) r.v[i] = t[i];
        if (t[N] || geq_p(r.v)) sub_p(r.v);
        return r;
    }
    Fp sqr() const { return *this * *this; }
    Fp pow(const uint64_t* e, int n) const {
        Fp r = one();
        for (int i = n - 1; i >= 0; i--)
            for (int b = 63; b >= 0; b--) { r = r.sqr(); if ((e[i] >> b) & 1) r = r * *this; }
        return r;
    }
    Fp inv() const {
        uint64_t e[N]; uint64_t br = 2;
        for (int i = 0; i < N; i++) { e[i] = P->p[i] - br; br = P->p[i] < br ? 1 : 0; }
        return pow(e, N);
    }
    Fp to_mont() const { Fp r2; for (int i = 0; i < N; i++) r2.v[i] = P->r2[i]; return *this * r2; }
    Fp from_mont() const { Fp o = zero(); o.v[0] = 1; return *this * o; }
    static Fp from_u64(uint64_t x) { Fp r = zero(); r.v[0] = x; return r.to_mont(); }
};

template <int N>
static Params<N> make_params(const uint64_t* p) {
    Params<N> P;
    for (int i = 0; i < N; i++) P.p[i] = p[i];
    uint64_t inv = 1;
    for (int i = 0; i < 63; i++) { inv *= inv; inv *= p[0]; }
    P.inv = (uint64_t)0 - inv;
    // R mod p and R^2 mod p by repeated doubling of 1
    uint64_t x[N];
    for (int i = 0; i < N; i++) x[i] = 0;
    x[0] = 1;
    auto dbl_mod = [&](uint64_t* a) {
        uint64_t c = 0;
        for (int i = 0; i < N; i++) { uint64_t nc = a[i] >> 63; a[i] = (a[i] << 1) | c; c = nc; }
        bool ge = c != 0;
        if (!ge) { ge = true; for (int i = N - 1; i >= 0; i--) { if (a[i] > p[i]) break; if (a[i] < p[i]) { ge = false; break; } } }
        if (ge) { u128 br = 0; for (int i = 0; i < N; i++) { u128 d = (u128)a[i] - p[i] - br; a[i] = (uint64_t)d; br = (d >> 64) & 1; } }
    };
    for (int i = 0; i < 64 * N; i++) dbl_mod(x);
    for (int i = 0; i < N; i++) P.one[i] = x[i];
    for (int i = 0; i < 64 * N; i++) dbl_mod(x);
    for (int i = 0; i < N; i++) P.r2[i] = x[i];
    return P;
}

static const uint64_t FQ_MOD[6] = {0xb9feffffffffaaabull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull,
                                   0x64774b84f38512bfull, 0x4b1ba7b6434bacd7ull, 0x1a0111ea397fe69aull};
static const uint64_t FR_MOD[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull};
static const Params<6> FQ_P = make_params<6>(FQ_MOD);
static const Params<4> FR_P = make_params<4>(FR_MOD);
typedef Fp<6, &FQ_P> Fq;
typedef Fp<4, &FR_P> Fr;

struct Fq2 {
    Fq c0, c1;
    static Fq2 zero() { return {Fq::zero(), Fq::zero()}; }
    static Fq2 one() { return {Fq::one(), Fq::zero()}; }
    bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    bool operator==(const Fq2& b) const { return c0 == b.c0 && c1 == b.c1; }
    Fq2 operator+(const Fq2& b) const { return {c0 + b.c0, c1 + b.c1}; }
    Fq2 operator-(const Fq2& b) const { return {c0 - b.c0, c1 - b.c1}; }
    Fq2 neg() const { return {c0.neg(), c1.neg()}; }
    Fq2 dbl() const { return {c0.dbl(), c1.dbl()}; }
    Fq2 operator*(const Fq2& b) const {
        Fq t0 = c0 * b.c0, t1 = c1 * b.c1, t2 = (c0 + c1) * (b.c0 + b.c1);
        return {t0 - t1, t2 - t0 - t1};
    }
    Fq2 sqr() const { return {(c0 + c1) * (c0 - c1), (c0 * c1).dbl()}; }
    Fq2 inv() const { Fq d = (c0.sqr() + c1.sqr()).inv(); return {c0 * d, (c1 * d).neg()}; }
};

// ------------------------------------------------------------------ curve (a = 0), Jacobian
template <class F> struct Aff { F x, y; bool is_inf() const { return x.is_zero() && y.is_zero(); } };
template <class F> struct Jac {
    F x, y, z;
    static Jac inf() { return {F::one(), F::one(), F::zero()}; }
    bool is_inf() const { return z.is_zero(); }
};
template <class F> static Jac<F> jdbl(const Jac<F>& p) {
    if (p.is_inf()) return p;
    F A = p.x.sqr(), B = p.y.sqr(), C = B.sqr();
    F D = ((p.x + B).sqr() - A - C).dbl();
    F E = A.dbl() + A, Fv = E.sqr();
    Jac<F> r;
    r.x = Fv - D.dbl();
    r.y = E * (D - r.x) - C.dbl().dbl().dbl();
    r.z = (p.y * p.z).dbl();
    return r;
}
template <class F> static Jac<F> jadd(const Jac<F>& a, const Jac<F>& b) {
    if (a.is_inf()) return b;
    if (b.is_inf()) return a;
    F z1z1 = a.z.sqr(), z2z2 = b.z.sqr();
    F u1 = a.x * z2z2, u2 = b.x * z1z1, s1 = a.y * b.z * z2z2, s2 = b.y * a.z * z1z1;
    F h = u2 - u1, rr = s2 - s1;
    if (h.is_zero()) return rr.is_zero() ? jdbl(a) : Jac<F>::inf();
    F hh = h.sqr(), hhh = h * hh, v = u1 * hh;
    Jac<F> r;
    r.x = rr.sqr() - hhh - v.dbl();
    r.y = rr * (v - r.x) - s1 * hhh;
    r.z = a.z * b.z * h;
    return r;
}
template <class F> static Jac<F> jmadd(const Jac<F>& a, const Aff<F>& b) {  // mixed addition
    if (b.is_inf()) return a;
    if (a.is_inf()) return {b.x, b.y, F::one()};
    F z1z1 = a.z.sqr();
    F u2 = b.x * z1z1, s2 = b.y * a.z * z1z1;
    F h = u2 - a.x, rr = s2 - a.y;
    if (h.is_zero()) return rr.is_zero() ? jdbl(a) : Jac<F>::inf();
    F hh = h.sqr(), hhh = h * hh, v = a.x * hh;
    Jac<F> r;
    r.x = rr.sqr() - hhh - v.dbl();
    r.y = rr * (v - r.x) - a.y * hhh;
    r.z = a.z * h;
    return r;
}
template <class F> static Aff<F> to_aff(const Jac<F>& p) {
    if (p.is_inf()) return {F::zero(), F::zero()};
    F zi = p.z.inv(), zi2 = zi.sqr();
    return {p.x * zi2, p.y * zi2 * zi};
}
template <class F> static Jac<F> jmul(const Jac<F>& p, const uint64_t* k, int limbs = 4) {
    Jac<F> r = Jac<F>::inf();
    for (int i = limbs - 1; i >= 0; i--)
        for (int b = 63; b >= 0; b--) { r = jdbl(r); if ((k[i] >> b) & 1) r = jadd(r, p); }
    return r;
}

// ------------------------------------------------------------------ VariableBaseMSM::msm_bigint
static int ln_without_floats(size_t a) { int lg = 0; while ((a >> (lg + 1)) != 0) lg++; return lg * 69 / 100; }

template <class F>
static Jac<F> msm_bigint(const Aff<F>* bases, const uint64_t* scalars /* n x 4 canonical */, size_t n) {
    const int c = n < 32 ? 3 : ln_without_floats(n) + 2;
    const int num_bits = 255;
    std::vector<int> starts;
    for (int s = 0; s < num_bits; s += c) starts.push_back(s);
    std::vector<Jac<F>> wsum(starts.size());
    parallel_for(starts.size(), 1, [&](size_t wi) {
        const int w0 = starts[wi];
        std::vector<Jac<F>> buckets(((size_t)1 << c) - 1, Jac<F>::inf());
        for (size_t i = 0; i < n; i++) {
            const uint64_t* k = scalars + 4 * i;
            if ((k[0] | k[1] | k[2] | k[3]) == 0) continue;  // (lambda body: `continue` stays inside this for)
            int limb = w0 >> 6, sh = w0 & 63;
            uint64_t d = k[limb] >> sh;
            if (sh + c > 64 && limb + 1 < 4) d |= k[limb + 1] << (64 - sh);
            d &= ((uint64_t)1 << c) - 1;
            if (d) buckets[d - 1] = jmadd(buckets[d - 1], bases[i]);
        }
        Jac<F> run = Jac<F>::inf(), res = Jac<F>::inf();
        for (size_t b = buckets.size(); b-- > 0;) { run = jadd(run, buckets[b]); res = jadd(res, run); }
        wsum[wi] = res;
    });
    Jac<F> total = wsum.back();
    for (size_t wi = wsum.size() - 1; wi-- > 0;) {
        for (int d = 0; d < c; d++) total = jdbl(total);
        total = jadd(total, wsum[wi]);
    }
    return total;
}

// ------------------------------------------------------------------ Radix2EvaluationDomain
static Fr fr_from_limbs(const uint64_t* p) { Fr r; memcpy(r.v, p, 32); return r; }
static const uint64_t ROOT_2_32_CANON[4] = {0x3829971f439f0d2bull, 0xb63683508c2280b9ull, 0xd09b681922c813b4ull, 0x16a2a19edfe81f20ull};

static Fr root_of_unity(uint32_t log_n) {
    Fr w = fr_from_limbs(ROOT_2_32_CANON).to_mont();
    for (uint32_t i = log_n; i < 32; i++) w = w.sqr();
    return w;
}

static void ntt_core(Fr* a, uint32_t log_n, const Fr& w) {
    const size_t n = (size_t)1 << log_n;
    for (size_t i = 0; i < n; i++) {
        size_t j = 0;
        for (uint32_t b = 0; b < log_n; b++) j |= ((i >> b) & 1) << (log_n - 1 - b);
        if (i < j) { Fr t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    std::vector<Fr> tw(n / 2 ? n / 2 : 1);
    tw[0] = Fr::one();
    for (size_t i = 1; i < n / 2; i++) tw[i] = tw[i - 1] * w;
    for (size_t m = 1; m < n; m <<= 1) {
        const size_t step = n / (2 * m);
        auto bfly = [&](size_t idx) {
            const size_t k = (idx / m) * 2 * m, j = idx % m;
            Fr u = a[k + j], v = a[k + j + m] * tw[j * step];
            a[k + j] = u + v;
            a[k + j + m] = u - v;
        };
        if (n >= (1u << 15)) parallel_for(n / 2, 4096, bfly);
        else for (size_t idx = 0; idx < n / 2; idx++) bfly(idx);
    }
}

static void domain_fft(Fr* a, uint32_t log_n, bool inverse, const Fr* offset) {
    const size_t n = (size_t)1 << log_n;
    const Fr w = root_of_unity(log_n);
    if (!inverse) {
        if (offset) { Fr g = Fr::one(); for (size_t i = 0; i < n; i++) { a[i] = a[i] * g; g = g * *offset; } }
        ntt_core(a, log_n, w);
    } else {
        ntt_core(a, log_n, w.inv());
        Fr g = Fr::from_u64(n).inv();
        const Fr oi = offset ? offset->inv() : Fr::one();
        for (size_t i = 0; i < n; i++) { a[i] = a[i] * g; if (offset) g = g * oi; }
    }
}

// ------------------------------------------------------------------ Poseidon (constants injected by the Python oracle)
static Fr P_RC[64][5], P_MDS[5][5];
static void poseidon_permute(Fr* s) {
    for (int rnd = 0; rnd < 64; rnd++) {
        for (int i = 0; i < 5; i++) s[i] = s[i] + P_RC[rnd][i];
        const bool full = rnd < 4 || rnd >= 60;
        for (int i = 0; i < (full ? 5 : 1); i++) { Fr x2 = s[i].sqr(); s[i] = x2.sqr() * s[i]; }
        Fr ns[5];
        for (int i = 0; i < 5; i++) { Fr acc = Fr::zero(); for (int j = 0; j < 5; j++) acc = acc + P_MDS[i][j] * s[j]; ns[i] = acc; }
        for (int i = 0; i < 5; i++) s[i] = ns[i];
    }
}
static Fr poseidon_hash(const Fr* in, int n) {
    Fr s[5] = {Fr::from_u64(1ull << 32) * Fr::from_u64(1ull << 32), Fr::zero(), Fr::zero(), Fr::zero(), Fr::zero()};
    const int chunks = n / 4 + 1;
    for (int c = 0; c < chunks; c++) {
        const int lo = c * 4, len = lo < n ? (n - lo < 4 ? n - lo : 4) : 0;
        for (int i = 0; i < len; i++) s[1 + i] = s[1 + i] + in[lo + i];
        if (len + 1 < 5) s[len + 1] = s[len + 1] + Fr::one();
        poseidon_permute(s);
    }
    return s[1];
}

typedef Aff<Fq> G1A;
typedef Aff<Fq2> G2A;

extern "C" {

int orc_threads() { return g_threads; }
void orc_set_threads(int t) { if (t > 0) g_threads = t; }

void orc_field_mul(int field, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        if (field == 0) { Fr x, y; memcpy(&x, a + 32 * i, 32); memcpy(&y, b + 32 * i, 32); Fr r = x * y; memcpy(out + 32 * i, &r, 32); }
        else { Fq x, y; memcpy(&x, a + 48 * i, 48); memcpy(&y, b + 48 * i, 48); Fq r = x * y; memcpy(out + 48 * i, &r, 48); }
    }
}

void orc_ntt(uint8_t* data, uint32_t log_n, int inverse, const uint8_t* coset_offset, size_t batch) {
    Fr off;
    if (coset_offset) memcpy(&off, coset_offset, 32);
    for (size_t b = 0; b < batch; b++) domain_fft((Fr*)data + (b << log_n), log_n, inverse != 0, coset_offset ? &off : nullptr);
}

void orc_msm_g1(const uint8_t* bases, const uint8_t* scalars, size_t n, uint8_t* out) {
    G1A r = to_aff(msm_bigint<Fq>((const G1A*)bases, (const uint64_t*)scalars, n));
    memcpy(out, &r, 96);
}
void orc_msm_g2(const uint8_t* bases, const uint8_t* scalars, size_t n, uint8_t* out) {
    G2A r = to_aff(msm_bigint<Fq2>((const G2A*)bases, (const uint64_t*)scalars, n));
    memcpy(out, &r, 192);
}
// k_i * G (parallel), used to make keys / bases for CPU-only tests
void orc_fixed_base_mul(int group, const uint8_t* gen, const uint8_t* scalars, size_t n, uint8_t* out) {
    parallel_for(n, 16, [&](size_t i) {
        const uint64_t* k = (const uint64_t*)(scalars + 32 * i);
        if (group == 1) {
            G1A g; memcpy(&g, gen, 96);
            G1A r = to_aff(jmul(Jac<Fq>{g.x, g.y, Fq::one()}, k));
            memcpy(out + 96 * i, &r, 96);
        } else {
            G2A g; memcpy(&g, gen, 192);
            G2A r = to_aff(jmul(Jac<Fq2>{g.x, g.y, Fq2::one()}, k));
            memcpy(out + 192 * i, &r, 192);
        }
    });
}

void orc_set_poseidon_constants(const uint8_t* rc, const uint8_t* mds) {
    memcpy(P_RC, rc, sizeof(P_RC));
    memcpy(P_MDS, mds, sizeof(P_MDS));
}
void orc_poseidon_hash_batch(const uint8_t* in, size_t n_hashes, uint32_t arity, uint8_t* out) {
    parallel_for(n_hashes, 4, [&](size_t i) {
        Fr h = poseidon_hash((const Fr*)in + i * arity, (int)arity);
        memcpy(out + 32 * i, &h, 32);
    });
}

// create_proof_with_reduction(r, s) for one full assignment z; matrices as CSR (row_ptr u64, cols u32,
// vals Montgomery Fr).  out_points: affine A (96) | B (192) | C (96).
void orc_groth16_prove(uint64_t nc, uint64_t num_inputs, uint64_t num_vars, uint32_t log_n,
                       const uint64_t* rpA, const uint32_t* cA, const uint8_t* vA,
                       const uint64_t* rpB, const uint32_t* cB, const uint8_t* vB,
                       const uint64_t* rpC, const uint32_t* cC, const uint8_t* vC,
                       const uint8_t* alpha_g1, const uint8_t* beta_g1, const uint8_t* beta_g2, const uint8_t* delta_g1,
                       const uint8_t* delta_g2, const uint8_t* a_query, const uint8_t* b_g1_query,
                       const uint8_t* b_g2_query, const uint8_t* l_query, const uint8_t* h_query, const uint8_t* z_mont,
                       const uint8_t* r_canon, const uint8_t* s_canon, uint8_t* out_points) {
    const size_t n = (size_t)1 << log_n;
    const Fr* z = (const Fr*)z_mont;
    std::vector<Fr> a(n, Fr::zero()), b(n, Fr::zero()), c(n, Fr::zero());
    auto rows = [&](const uint64_t* rp, const uint32_t* cols, const uint8_t* vals, std::vector<Fr>& dst) {
        parallel_for(nc, 256, [&](size_t i) {
            Fr acc = Fr::zero();
            for (uint64_t k = rp[i]; k < rp[i + 1]; k++) acc = acc + ((const Fr*)vals)[k] * z[cols[k]];
            dst[i] = acc;
        });
    };
    rows(rpA, cA, vA, a);
    rows(rpB, cB, vB, b);
    rows(rpC, cC, vC, c);
    for (uint64_t j = 0; j < num_inputs; j++) a[nc + j] = z[j];
    const Fr g = Fr::from_u64(7);
    for (std::vector<Fr>* v : {&a, &b, &c}) { domain_fft(v->data(), log_n, true, nullptr); domain_fft(v->data(), log_n, false, &g); }
    Fr gn = g;
    for (uint32_t i = 0; i < log_n; i++) gn = gn.sqr();
    const Fr zinv = (gn - Fr::one()).inv();
    for (size_t i = 0; i < n; i++) a[i] = (a[i] * b[i] - c[i]) * zinv;
    domain_fft(a.data(), log_n, true, &g);
    // scalars to canonical form (into_bigint)
    std::vector<uint64_t> zc(4 * num_vars), hc(4 * n);
    for (uint64_t i = 0; i < num_vars; i++) { Fr t = z[i].from_mont(); memcpy(&zc[4 * i], t.v, 32); }
    for (size_t i = 0; i < n; i++) { Fr t = a[i].from_mont(); memcpy(&hc[4 * i], t.v, 32); }
    const uint64_t* rr = (const uint64_t*)r_canon;
    const uint64_t* ss = (const uint64_t*)s_canon;
    G1A al, be1, de1; G2A be2, de2;
    memcpy(&al, alpha_g1, 96); memcpy(&be1, beta_g1, 96); memcpy(&de1, delta_g1, 96);
    memcpy(&be2, beta_g2, 192); memcpy(&de2, delta_g2, 192);
    Jac<Fq> A = msm_bigint<Fq>((const G1A*)a_query, zc.data(), num_vars);
    A = jmadd(A, al);
    A = jadd(A, jmul(Jac<Fq>{de1.x, de1.y, Fq::one()}, rr));
    Jac<Fq> B1 = msm_bigint<Fq>((const G1A*)b_g1_query, zc.data(), num_vars);
    B1 = jmadd(B1, be1);
    B1 = jadd(B1, jmul(Jac<Fq>{de1.x, de1.y, Fq::one()}, ss));
    Jac<Fq2> B2 = msm_bigint<Fq2>((const G2A*)b_g2_query, zc.data(), num_vars);
    B2 = jmadd(B2, be2);
    B2 = jadd(B2, jmul(Jac<Fq2>{de2.x, de2.y, Fq2::one()}, ss));
    Jac<Fq> C = msm_bigint<Fq>((const G1A*)l_query, zc.data() + 4 * num_inputs, num_vars - num_inputs);
    C = jadd(C, msm_bigint<Fq>((const G1A*)h_query, hc.data(), n - 1));
    C = jadd(C, jmul(A, ss));
    C = jadd(C, jmul(B1, rr));
    Fr fr_r = fr_from_limbs(rr).to_mont(), fr_s = fr_from_limbs(ss).to_mont();
    Fr rs = (fr_r * fr_s).from_mont();
    Jac<Fq> rsd = jmul(Jac<Fq>{de1.x, de1.y, Fq::one()}, rs.v);
    rsd.y = rsd.y.neg();
    C = jadd(C, rsd);
    G1A oa = to_aff(A), oc = to_aff(C);
    G2A ob = to_aff(B2);
    memcpy(out_points, &oa, 96);
    memcpy(out_points + 96, &ob, 192);
    memcpy(out_points + 288, &oc, 96);
}

}  // extern "C"
