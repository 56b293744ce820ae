// Host side of the shielder relation: Poseidon parameter generation and the R1CS of the
// "update note" statement -- the ConstraintSynthesizer::generate_constraints half of the drop-in
// (the witness half runs on the GPU, relation.cu).
//
// Functional spec (reference file:line; nothing in the reference executes these builders):
//   update_note_circuit          shielder/relations/src/relations/update_note.rs:106-149
//   UpdateNoteInput::new         update_note.rs:47-88      (allocation order of the witnesses)
//   verify_note_circuit          update_note.rs:91-103
//   CircuitMerkleProof::verify   shielder/relations/src/merkle_proof.rs:38-61
//   update_account_circuit       shielder/relations/src/relations/update_account.rs:68-95
//   Poseidon parameters          shielder/relations/src/lib.rs:17-26 (T=5, RATE=4, R_F=8, R_P=56)
// Account / Operation are abstract in the reference (account.rs:8-21, operation.rs:3-23); the
// concrete shapes follow the mock: 2 (token, balance) pairs, OpPub{amount,token,user},
// OpPriv{user} (shielder/mocked_zk/src/account.rs:12-14,36-79; ops.rs:6-32,47-62).
// Variable numbering is arkworks': z = [1, instance..., witness...].
#pragma once
#include <cstdint>
#include <utility>
#include <vector>

#include "field.cuh"

namespace b200zk {
namespace host {

constexpr int POSEIDON_T = 5, POSEIDON_RATE = 4, POSEIDON_RF = 8, POSEIDON_RP = 56;
constexpr int POSEIDON_ROUNDS = POSEIDON_RF + POSEIDON_RP;
constexpr int POSEIDON_SBOXES = POSEIDON_RF * POSEIDON_T + POSEIDON_RP;  // 96
constexpr int POSEIDON_TRACE = POSEIDON_SBOXES * 3;                      // witnesses per permutation
constexpr int BALANCE_BITS = 128;
constexpr int KIND_DEPOSIT = 0, KIND_WITHDRAW = 1;
constexpr int RELATION_UPDATE_NOTE = 0, RELATION_UPDATE_ACCOUNT = 1;

struct PoseidonConsts {
    Fr rc[POSEIDON_ROUNDS][POSEIDON_T];
    Fr mds[POSEIDON_T][POSEIDON_T];
    Fr two64;  // initial state[0] = 2^64
};

inline Fr fr_from_u64(uint64_t v) {
    Fr r = Fr::zero();
    r.v[0] = (uint32_t)v;
    r.v[1] = (uint32_t)(v >> 32);
    return fp_to_mont(r);
}

// Poseidon reference Grain LFSR (80 bit, self-shrinking); see oracle/pyref/poseidon.py for the
// provenance of every convention.
class Grain {
   public:
    Grain(int field_bits, int t, int r_f, int r_p) {
        int pos = 0;
        auto push = [&](uint32_t v, int n) {
            for (int i = n - 1; i >= 0; i--) st_[pos++] = (v >> i) & 1;
        };
        push(1, 2);
        push(0, 4);
        push(field_bits, 12);
        push(t, 12);
        push(r_f, 10);
        push(r_p, 10);
        push(0x3fffffff, 30);
        head_ = 0;
        for (int i = 0; i < 160; i++) clock();
    }
    int next_bit() {
        for (;;) {
            if (clock()) return clock();
            clock();
        }
    }
    // n-bit big-endian integer as canonical little-endian limbs
    Fr random_bits(int n) {
        Fr r = Fr::zero();
        for (int i = n - 1; i >= 0; i--)
            if (next_bit()) r.v[i / 32] |= 1u << (i % 32);
        return r;
    }
    Fr field_element() {  // rejection sampling, canonical -> Montgomery
        for (;;) {
            Fr c = random_bits(255);
            if (less_than_modulus(c)) return fp_to_mont(c);
        }
    }
    Fr field_element_no_reject() {  // value mod r (one subtraction suffices: 2^255 < 2r)
        Fr c = random_bits(255);
        if (!less_than_modulus(c)) {
            uint64_t br = 0;
            for (int i = 0; i < 8; i++) {
                uint64_t d = (uint64_t)c.v[i] - FrCfg::mod(i) - br;
                c.v[i] = (uint32_t)d;
                br = (d >> 63) & 1;
            }
        }
        return fp_to_mont(c);
    }

   private:
    static bool less_than_modulus(const Fr& c) {
        for (int i = 7; i >= 0; i--) {
            if (c.v[i] < FrCfg::mod(i)) return true;
            if (c.v[i] > FrCfg::mod(i)) return false;
        }
        return false;
    }
    int bit(int i) const { return st_[(head_ + i) % 80]; }
    int clock() {
        int nb = bit(62) ^ bit(51) ^ bit(38) ^ bit(23) ^ bit(13) ^ bit(0);
        st_[head_] = nb;  // slot of the bit shifted out becomes the new last bit
        head_ = (head_ + 1) % 80;
        return nb;
    }
    int st_[80];
    int head_;
};

inline const PoseidonConsts& poseidon_consts() {
    static PoseidonConsts pc;
    static bool ready = false;
    if (!ready) {
        Grain g(255, POSEIDON_T, POSEIDON_RF, POSEIDON_RP);
        for (int r = 0; r < POSEIDON_ROUNDS; r++)
            for (int i = 0; i < POSEIDON_T; i++) pc.rc[r][i] = g.field_element();
        Fr xs[POSEIDON_T], ys[POSEIDON_T];
        for (;;) {
            Fr v[2 * POSEIDON_T];
            for (auto& e : v) e = g.field_element_no_reject();
            bool ok = true;
            for (int i = 0; i < 2 * POSEIDON_T && ok; i++)
                for (int j = 0; j < i; j++)
                    if (v[i] == v[j]) ok = false;
            for (int i = 0; i < POSEIDON_T; i++) {
                xs[i] = v[i];
                ys[i] = v[POSEIDON_T + i];
            }
            for (int i = 0; i < POSEIDON_T && ok; i++)
                for (int j = 0; j < POSEIDON_T; j++)
                    if (fp_add(xs[i], ys[j]).is_zero()) ok = false;
            if (ok) break;
        }
        for (int i = 0; i < POSEIDON_T; i++)
            for (int j = 0; j < POSEIDON_T; j++) pc.mds[i][j] = fp_inv(fp_add(xs[i], ys[j]));
        Fr two32 = fr_from_u64(1ull << 32);
        pc.two64 = fp_mul(two32, two32);
        ready = true;
    }
    return pc;
}

// ---------------------------------------------------------------------------- linear combinations
struct LC {
    std::vector<std::pair<uint32_t, Fr>> t;  // sorted by variable, no zero coefficients

    static LC var(uint32_t v) {
        LC r;
        r.t.emplace_back(v, Fr::one());
        return r;
    }
    static LC constant(const Fr& c) {
        LC r;
        if (!c.is_zero()) r.t.emplace_back(0u, c);
        return r;
    }
    LC scaled(const Fr& c) const {
        LC r;
        if (c.is_zero()) return r;
        r.t.reserve(t.size());
        for (auto& e : t) r.t.emplace_back(e.first, fp_mul(e.second, c));
        return r;
    }
    LC neg() const {
        LC r;
        r.t.reserve(t.size());
        for (auto& e : t) r.t.emplace_back(e.first, fp_neg(e.second));
        return r;
    }
    // this + c * o
    LC add_scaled(const LC& o, const Fr* c) const {
        LC r;
        r.t.reserve(t.size() + o.t.size());
        size_t i = 0, j = 0;
        while (i < t.size() || j < o.t.size()) {
            if (j == o.t.size() || (i < t.size() && t[i].first < o.t[j].first)) {
                r.t.push_back(t[i++]);
            } else {
                Fr oc = c ? fp_mul(o.t[j].second, *c) : o.t[j].second;
                if (i < t.size() && t[i].first == o.t[j].first) {
                    Fr s = fp_add(t[i].second, oc);
                    if (!s.is_zero()) r.t.emplace_back(t[i].first, s);
                    i++;
                } else if (!oc.is_zero()) {
                    r.t.emplace_back(o.t[j].first, oc);
                }
                j++;
            }
        }
        return r;
    }
    LC operator+(const LC& o) const { return add_scaled(o, nullptr); }
    LC operator-(const LC& o) const { return add_scaled(o.neg(), nullptr); }
};

struct R1CS {
    uint32_t num_inputs = 1;  // counts the constant ONE
    uint32_t num_aux = 0;
    std::vector<LC> A, B, C;
    int kind = KIND_WITHDRAW;
    int relation = 0;         // RELATION_UPDATE_NOTE / RELATION_UPDATE_ACCOUNT
    uint32_t tree_height = 0;
    uint32_t inputs_per_instance = 0;  // Fr elements of one input row of the witness generator

    uint32_t num_variables() const { return num_inputs + num_aux; }
    LC alloc_input() { return LC::var(num_inputs++); }
    LC alloc_witness() { return LC::var(num_inputs + num_aux++); }
    void enforce(const LC& a, const LC& b, const LC& c) {
        A.push_back(a);
        B.push_back(b);
        C.push_back(c);
    }
};

// ---------------------------------------------------------------------------- gadgets
inline LC g_mul(R1CS& cs, const LC& a, const LC& b) {
    LC out = cs.alloc_witness();
    cs.enforce(a, b, out);
    return out;
}
inline void g_assert_equal(R1CS& cs, const LC& a, const LC& b) { cs.enforce(a - b, LC::constant(Fr::one()), LC()); }
inline LC g_is_zero(R1CS& cs, const LC& x) {  // GateChip::is_zero: x*inv = 1 - out ; x*out = 0
    LC inv = cs.alloc_witness();
    LC out = cs.alloc_witness();
    cs.enforce(x, inv, LC::constant(Fr::one()) - out);
    cs.enforce(x, out, LC());
    return out;
}
inline LC g_select(R1CS& cs, const LC& a, const LC& b, const LC& sel) {  // GateChip::select(a, b, sel)
    return g_mul(cs, sel, a - b) + b;
}
inline void g_range_bits(R1CS& cs, const LC& x, int nbits) {
    LC acc;
    Fr pow2 = Fr::one();
    const LC one = LC::constant(Fr::one());
    for (int i = 0; i < nbits; i++) {
        LC b = cs.alloc_witness();
        cs.enforce(b, b - one, LC());
        acc = acc.add_scaled(b, &pow2);
        pow2 = fp_dbl(pow2);
    }
    g_assert_equal(cs, acc, x);
}
inline void g_poseidon_permute(R1CS& cs, LC* s) {
    const PoseidonConsts& pc = poseidon_consts();
    const int half = POSEIDON_RF / 2;
    for (int rnd = 0; rnd < POSEIDON_ROUNDS; rnd++) {
        for (int i = 0; i < POSEIDON_T; i++) s[i] = s[i] + LC::constant(pc.rc[rnd][i]);
        const bool full = rnd < half || rnd >= half + POSEIDON_RP;
        for (int i = 0; i < (full ? POSEIDON_T : 1); i++) {
            LC x2 = g_mul(cs, s[i], s[i]);
            LC x4 = g_mul(cs, x2, x2);
            s[i] = g_mul(cs, x4, s[i]);
        }
        LC ns[POSEIDON_T];
        for (int i = 0; i < POSEIDON_T; i++)
            for (int j = 0; j < POSEIDON_T; j++) ns[i] = ns[i].add_scaled(s[j], &pc.mds[i][j]);
        for (int i = 0; i < POSEIDON_T; i++) s[i] = ns[i];
    }
}
inline LC g_poseidon_hash(R1CS& cs, const std::vector<LC>& in) {  // PoseidonHasher::hash_fix_len_array
    const PoseidonConsts& pc = poseidon_consts();
    LC s[POSEIDON_T];
    s[0] = LC::constant(pc.two64);
    const LC one = LC::constant(Fr::one());
    size_t n = in.size();
    size_t chunks = n / POSEIDON_RATE + 1;  // the extra empty chunk when n % RATE == 0 is the "+1"
    for (size_t c = 0; c < chunks; c++) {
        size_t lo = c * POSEIDON_RATE, len = lo < n ? std::min<size_t>(POSEIDON_RATE, n - lo) : 0;
        for (size_t i = 0; i < len; i++) s[1 + i] = s[1 + i] + in[lo + i];
        if (len + 1 < (size_t)POSEIDON_T) s[len + 1] = s[len + 1] + one;
        g_poseidon_permute(cs, s);
    }
    return s[1];
}

// update_account_circuit (update_account.rs:68-95) with the mock's concrete Account / Operation: verify_account_circuit
// on the old account (:79-85), `update` (account.rs:36-79 of the mock: the balance whose token matches moves by
// `amount`, checked_add / checked_sub as 128-bit range checks, exactly one token must match), verify_account_circuit
// on the new one (:88-94).  Shared by the standalone relation and by update_note_circuit, which calls it as its last
// step (update_note.rs:141-148) -- the witness kernels allocate in exactly this order.
inline void g_update_account(R1CS& cs, int kind, const LC& amount, const LC& token, const LC acc_token[2],
                             const LC acc_balance[2], const LC& old_account_hash, const LC& new_account_hash) {
    std::vector<LC> old_vec = {acc_token[0], acc_balance[0], acc_token[1], acc_balance[1]};
    g_assert_equal(cs, g_poseidon_hash(cs, old_vec), old_account_hash);
    g_range_bits(cs, amount, BALANCE_BITS);
    LC matches;
    std::vector<LC> new_vec;
    for (int i = 0; i < 2; i++) {
        LC eq = g_is_zero(cs, acc_token[i] - token);
        LC delta = g_mul(cs, eq, amount);
        LC nb = kind == KIND_DEPOSIT ? acc_balance[i] + delta : acc_balance[i] - delta;
        g_range_bits(cs, nb, BALANCE_BITS);
        matches = matches + eq;
        new_vec.push_back(acc_token[i]);
        new_vec.push_back(nb);
    }
    g_assert_equal(cs, matches, LC::constant(Fr::one()));
    g_assert_equal(cs, g_poseidon_hash(cs, new_vec), new_account_hash);
}

// update_note_circuit as R1CS.  Must stay in lock-step with update_note_witness_kernel (relation.cu):
// both allocate witnesses in exactly this order.
inline R1CS synthesize_update_note(int kind, uint32_t tree_height) {
    R1CS cs;
    cs.kind = kind;
    cs.relation = RELATION_UPDATE_NOTE;
    cs.tree_height = tree_height;
    cs.inputs_per_instance = 18 + 2 * tree_height;
    // instance variables (make_public order, update_note.rs:121,127)
    LC amount = cs.alloc_input(), token = cs.alloc_input(), user = cs.alloc_input();
    LC new_note_hash = cs.alloc_input(), merkle_root = cs.alloc_input(), old_nullifier = cs.alloc_input();
    // witnesses in UpdateNoteInput::new order (update_note.rs:58-76)
    std::vector<LC> new_note(4);
    for (auto& v : new_note) v = cs.alloc_witness();
    LC old_zk_id = cs.alloc_witness(), old_trapdoor = cs.alloc_witness(), old_account_hash = cs.alloc_witness();
    std::vector<LC> old_note = {old_zk_id, old_trapdoor, old_nullifier, old_account_hash};
    std::vector<LC> path_shape(tree_height), path(tree_height);
    for (auto& v : path_shape) v = cs.alloc_witness();
    for (auto& v : path) v = cs.alloc_witness();
    LC op_priv_user = cs.alloc_witness();
    LC acc_token[2], acc_balance[2];
    for (int i = 0; i < 2; i++) {
        acc_token[i] = cs.alloc_witness();
        acc_balance[i] = cs.alloc_witness();
    }
    // verify_note_circuit(new_note, new_note_hash)                 update_note.rs:129
    g_assert_equal(cs, g_poseidon_hash(cs, new_note), new_note_hash);
    // old_note_hash                                                 update_note.rs:131
    LC current = g_poseidon_hash(cs, old_note);
    // merkle_proof.verify                                           merkle_proof.rs:49-60
    for (uint32_t i = 0; i < tree_height; i++) {
        LC selector = g_is_zero(cs, path_shape[i]);
        LC left = g_select(cs, path[i], current, selector);
        LC right = g_select(cs, current, path[i], selector);
        current = g_poseidon_hash(cs, {left, right});
    }
    g_assert_equal(cs, current, merkle_root);
    // CircuitOperation::combine(op_priv, op_pub).unwrap()           update_note.rs:139
    g_assert_equal(cs, user, op_priv_user);
    // update_account_circuit                                        update_note.rs:141-148 -> update_account.rs:68-95
    g_update_account(cs, kind, amount, token, acc_token, acc_balance, old_account_hash, new_note[3]);
    return cs;
}

// update_account_circuit as a relation of its own (update_account.rs:68-95; UpdateAccountInput :18-30).
// Instance variables in the struct's field order ("public inputs", :23-26): old_account_hash, new_account_hash,
// operation = (amount, token, user) -- the operation arrives already combined (A::Op), so `user` is carried as an
// instance variable without a constraint of its own, like every other field `update` does not look at.
// Witness: old_account (:29) = (token0, balance0, token1, balance1).  Lock-step with update_account_witness_kernel.
inline R1CS synthesize_update_account(int kind) {
    R1CS cs;
    cs.kind = kind;
    cs.relation = RELATION_UPDATE_ACCOUNT;
    cs.tree_height = 0;
    cs.inputs_per_instance = 9;
    LC old_account_hash = cs.alloc_input(), new_account_hash = cs.alloc_input();
    LC amount = cs.alloc_input(), token = cs.alloc_input(), user = cs.alloc_input();
    (void)user;
    LC acc_token[2], acc_balance[2];
    for (int i = 0; i < 2; i++) {
        acc_token[i] = cs.alloc_witness();
        acc_balance[i] = cs.alloc_witness();
    }
    g_update_account(cs, kind, amount, token, acc_token, acc_balance, old_account_hash, new_account_hash);
    return cs;
}

}  // namespace host
}  // namespace b200zk
