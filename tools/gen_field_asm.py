#!/usr/bin/env python3
"""Generates zk-apps_b200/csrc/field_asm.cuh: one inline-PTX statement per carry chain.

Why generated: a Montgomery product is ~4*N^2 carry-chained mad instructions.  Emitting each as
its own `asm volatile` makes the NVVM front end crawl (minutes per translation unit once the EC
formulas inline ~100 products); one statement per chain keeps the condition-code dependency
inside a single asm block (no reliance on statement adjacency) and cuts compile time ~10x.

Chains (N = limbs, 8 for Fr, 12 for Fq), all on aligned register pairs so that every
mad.lo.cc/madc.hi.cc pair becomes one IMAD.WIDE.U32(.X):
  row_even_N(X, a, b):        X[0..N]  += sum_{j even} a[j]*b << 32j        (carry into X[N])
  row_odd_N(Y, a, b):         Y[0..N-1] += sum_{j odd} a[j]*b << 32(j-1)    (no carry out, see field.cuh)
  row_odd_cin_N(nx0,y0,x1,Y,a,b): nx0 = y0 + x1, its carry enters the row_odd chain
  add_N / sub_N:              plain limb chains
"""
import sys

def ops(names):
    return ", ".join(names)

def gen(N):
    out = []
    H = N // 2
    # ---------------- row_even
    lines = []
    # operands: %0..%N = X[0..N] (+r), then a_even (H), then b
    a0 = N + 1
    b = a0 + H
    for k in range(H):
        j = 2 * k
        lo = "mad.lo.cc.u32" if k == 0 else "madc.lo.cc.u32"
        lines.append(f"{lo} %{j}, %{a0 + k}, %{b}, %{j};")
        lines.append(f"madc.hi.cc.u32 %{j + 1}, %{a0 + k}, %{b}, %{j + 1};")
    lines.append(f"addc.u32 %{N}, %{N}, 0;")
    body = "\\n\\t".join(lines)
    outs = ", ".join(f'"+r"(X[{i}])' for i in range(N + 1))
    ins = ", ".join([f'"r"(a[{2 * k}])' for k in range(H)] + ['"r"(b)'])
    out.append(f"""B200ZK_ASM_DEV void row_even_{N}(uint32_t* X, const uint32_t* a, uint32_t b) {{
    asm volatile("{body}"
                 : {outs}
                 : {ins});
}}""")
    # ---------------- row_odd (no carry in)
    def row_odd(cin):
        lines = []
        # operands: %0..%(N-1) = Y (+r); [cin: %N = nx0 (=r)]; a_odd (H); b; [cin: y0, x1]
        base = N + (1 if cin else 0)
        a0 = base
        b = a0 + H
        if cin:
            lines.append(f"add.cc.u32 %{N}, %{b + 1}, %{b + 2};")
        for k in range(H):
            j = 2 * k  # Y index of the pair (Y[j], Y[j+1]) <- a[2k+1]
            lo = "madc.lo.cc.u32" if (cin or k > 0) else "mad.lo.cc.u32"
            lines.append(f"{lo} %{j}, %{a0 + k}, %{b}, %{j};")
            hi = "madc.hi.cc.u32" if k < H - 1 else "madc.hi.u32"
            lines.append(f"{hi} %{j + 1}, %{a0 + k}, %{b}, %{j + 1};")
        body = "\\n\\t".join(lines)
        outs = [f'"+r"(Y[{i}])' for i in range(N)]
        if cin:
            outs.append('"=&r"(nx0)')
        ins = [f'"r"(a[{2 * k + 1}])' for k in range(H)] + ['"r"(b)']
        if cin:
            ins += ['"r"(y0)', '"r"(x1)']
        sig = (f"row_odd_cin_{N}(uint32_t& nx0, uint32_t y0, uint32_t x1, uint32_t* Y, const uint32_t* a, uint32_t b)"
               if cin else f"row_odd_{N}(uint32_t* Y, const uint32_t* a, uint32_t b)")
        return f"""B200ZK_ASM_DEV void {sig} {{
    asm volatile("{body}"
                 : {", ".join(outs)}
                 : {", ".join(ins)});
}}"""
    out.append(row_odd(False))
    out.append(row_odd(True))
    # ---------------- add (no carry out) : r = a + b
    lines = []
    for i in range(N):
        op = "add.cc.u32" if i == 0 else ("addc.cc.u32" if i < N - 1 else "addc.u32")
        lines.append(f"{op} %{i}, %{N + i}, %{2 * N + i};")
    body = "\\n\\t".join(lines)
    outs = ", ".join(f'"=&r"(r[{i}])' for i in range(N))
    ins = ", ".join([f'"r"(a[{i}])' for i in range(N)] + [f'"r"(b[{i}])' for i in range(N)])
    out.append(f"""B200ZK_ASM_DEV void add_{N}(uint32_t* r, const uint32_t* a, const uint32_t* b) {{
    asm volatile("{body}"
                 : {outs}
                 : {ins});
}}""")
    # ---------------- sub with borrow mask : r = a - b, returns 0xffffffff if a < b
    lines = []
    for i in range(N):
        op = "sub.cc.u32" if i == 0 else "subc.cc.u32"
        lines.append(f"{op} %{i}, %{N + 1 + i}, %{2 * N + 1 + i};")
    lines.append(f"subc.u32 %{N}, 0, 0;")
    body = "\\n\\t".join(lines)
    outs = ", ".join([f'"=&r"(r[{i}])' for i in range(N)] + ['"=&r"(borrow)'])
    ins = ", ".join([f'"r"(a[{i}])' for i in range(N)] + [f'"r"(b[{i}])' for i in range(N)])
    out.append(f"""B200ZK_ASM_DEV uint32_t sub_{N}(uint32_t* r, const uint32_t* a, const uint32_t* b) {{
    uint32_t borrow;
    asm volatile("{body}"
                 : {outs}
                 : {ins});
    return borrow;
}}""")
    return "\n\n".join(out)


# ----------------------------------------------------------------------------------------------
# Wide (unreduced) products for lazy reduction: T[0..2N) = a*b and T = a^2, and the pieces of a
# word-by-word Montgomery reduction of a 2N-limb value.  Same alignment trick as the CIOS rows: two
# accumulators, E for even columns and O for odd columns (O[k] is column k+1), so that every
# mad.lo/madc.hi pair lands on an aligned register pair.  Rows are processed in increasing order, so the
# limb that receives a chain's final carry has only ever collected carries (a handful) and cannot overflow.
def chain_stmt(pairs, carry):
    """pairs: list of (acc_lo_expr, acc_hi_expr, x_expr, y_expr); carry: acc expr or None."""
    outs, ins, lines = [], [], []
    def o(expr):
        if expr not in outs:
            outs.append(expr)
        return outs.index(expr)
    for lo, hi, x, y in pairs:
        o(lo); o(hi)
    if carry:
        o(carry)
    nout = len(outs)
    def i(expr):
        if expr not in ins:
            ins.append(expr)
        return nout + ins.index(expr)
    first = True
    for lo, hi, x, y in pairs:
        xi, yi = i(x), i(y)
        lines.append(f"{'mad.lo.cc.u32' if first else 'madc.lo.cc.u32'} %{o(lo)}, %{xi}, %{yi}, %{o(lo)};")
        last = (lo, hi, x, y) == pairs[-1] and not carry
        lines.append(f"{'madc.hi.u32' if last else 'madc.hi.cc.u32'} %{o(hi)}, %{xi}, %{yi}, %{o(hi)};")
        first = False
    if carry:
        lines.append(f"addc.u32 %{o(carry)}, %{o(carry)}, 0;")
    body = "\\n\\t".join(lines)
    return ('    asm volatile("' + body + '"\n                 : ' + ", ".join(f'"+r"({e})' for e in outs) +
            "\n                 : " + ", ".join(f'"r"({e})' for e in ins) + ");")


def add_chain_stmt(dst, a, b, final_carry_out=None):
    """dst[k] = a[k] + b[k] with carries, lists of expressions."""
    n = len(dst)
    outs = list(dst) + ([final_carry_out] if final_carry_out else [])
    ins = list(a) + list(b)
    lines = []
    for k in range(n):
        op = "add.cc.u32" if k == 0 else ("addc.cc.u32" if (k < n - 1 or final_carry_out) else "addc.u32")
        lines.append(f"{op} %{k}, %{len(outs) + k}, %{len(outs) + n + k};")
    if final_carry_out:
        lines.append(f"addc.u32 %{n}, 0, 0;")
    body = "\\n\\t".join(lines)
    return ('    asm volatile("' + body + '"\n                 : ' + ", ".join(f'"=&r"({e})' for e in outs) +
            "\n                 : " + ", ".join(f'"r"({e})' for e in ins) + ");")


def gen_wide(N):
    H = N // 2
    out = []
    # ---------------- mul_wide
    body = [f"    uint32_t E[{2 * N}], O[{2 * N - 1}];",
            f"#pragma unroll\n    for (int k = 0; k < {2 * N}; k++) E[k] = 0;",
            f"#pragma unroll\n    for (int k = 0; k < {2 * N - 1}; k++) O[k] = 0;"]
    for i in range(N):
        if i % 2 == 0:
            body.append(chain_stmt([(f"E[{i + 2 * k}]", f"E[{i + 2 * k + 1}]", f"a[{2 * k}]", f"b[{i}]") for k in range(H)],
                                   f"E[{i + N}]"))
            body.append(chain_stmt([(f"O[{i + 2 * k}]", f"O[{i + 2 * k + 1}]", f"a[{2 * k + 1}]", f"b[{i}]") for k in range(H)],
                                   f"O[{i + N}]" if i + N <= 2 * N - 2 else None))
        else:
            body.append(chain_stmt([(f"O[{i - 1 + 2 * k}]", f"O[{i + 2 * k}]", f"a[{2 * k}]", f"b[{i}]") for k in range(H)],
                                   f"O[{i - 1 + N}]"))
            body.append(chain_stmt([(f"E[{i + 1 + 2 * k}]", f"E[{i + 2 + 2 * k}]", f"a[{2 * k + 1}]", f"b[{i}]") for k in range(H)],
                                   f"E[{i + 1 + N}]" if i + 1 + N <= 2 * N - 1 else None))
    body.append("    T[0] = E[0];")
    body.append(add_chain_stmt([f"T[{k}]" for k in range(1, 2 * N)], [f"E[{k}]" for k in range(1, 2 * N)],
                               [f"O[{k}]" for k in range(0, 2 * N - 1)]))
    out.append(f"B200ZK_ASM_DEV void mul_wide_{N}(uint32_t* T, const uint32_t* a, const uint32_t* b) {{\n" + "\n".join(body) + "\n}")
    # ---------------- sqr_wide: cross products (i < j) into E / O, doubled, plus the diagonal
    body = [f"    uint32_t E[{2 * N}], O[{2 * N - 1}], X[{2 * N}];",
            f"#pragma unroll\n    for (int k = 0; k < {2 * N}; k++) E[k] = 0;",
            f"#pragma unroll\n    for (int k = 0; k < {2 * N - 1}; k++) O[k] = 0;"]
    for i in range(N - 1):
        js_o = list(range(i + 1, N, 2))     # i + j odd  -> O index i + j - 1 (even)
        js_e = list(range(i + 2, N, 2))     # i + j even -> E index i + j (even)
        if js_o:
            last = i + js_o[-1] - 1 + 2
            body.append(chain_stmt([(f"O[{i + j - 1}]", f"O[{i + j}]", f"a[{i}]", f"a[{j}]") for j in js_o],
                                   f"O[{last}]" if last <= 2 * N - 2 else None))
        if js_e:
            last = i + js_e[-1] + 2
            body.append(chain_stmt([(f"E[{i + j}]", f"E[{i + j + 1}]", f"a[{i}]", f"a[{j}]") for j in js_e],
                                   f"E[{last}]" if last <= 2 * N - 1 else None))
    body.append("    X[0] = 0;")   # column 0 has no cross term
    body.append(add_chain_stmt([f"X[{k}]" for k in range(1, 2 * N)], [f"E[{k}]" for k in range(1, 2 * N)],
                               [f"O[{k}]" for k in range(0, 2 * N - 1)]))
    body.append(add_chain_stmt([f"T[{k}]" for k in range(0, 2 * N)], [f"X[{k}]" for k in range(0, 2 * N)],
                               [f"X[{k}]" for k in range(0, 2 * N)]))
    diag = [(f"T[{2 * k}]", f"T[{2 * k + 1}]", f"a[{k}]", f"a[{k}]") for k in range(N)]
    body.append(chain_stmt(diag, None))
    out.append(f"B200ZK_ASM_DEV void sqr_wide_{N}(uint32_t* T, const uint32_t* a) {{\n" + "\n".join(body) + "\n}")
    # ---------------- redc step: nx0 = y0 + x1 (carry into the odd row), m = nx0 * inv, Y += m * p_odd
    outs = [f'"+r"(Y[{k}])' for k in range(N)] + ['"=&r"(nx0)', '"=&r"(m)']
    ins = [f'"r"(p[{2 * k + 1}])' for k in range(H)] + ['"r"(y0)', '"r"(x1)', '"r"(inv)']
    p0 = N + 2
    lines = [f"add.cc.u32 %{N}, %{p0 + H}, %{p0 + H + 1};", f"mul.lo.u32 %{N + 1}, %{N}, %{p0 + H + 2};"]
    for k in range(H):
        lines.append(f"madc.lo.cc.u32 %{2 * k}, %{p0 + k}, %{N + 1}, %{2 * k};")
        lines.append(f"{'madc.hi.cc.u32' if k < H - 1 else 'madc.hi.u32'} %{2 * k + 1}, %{p0 + k}, %{N + 1}, %{2 * k + 1};")
    body = "\\n\\t".join(lines)
    out.append(f"""B200ZK_ASM_DEV void redc_step_{N}(uint32_t& nx0, uint32_t& m, uint32_t y0, uint32_t x1, uint32_t* Y, const uint32_t* p, uint32_t inv) {{
    asm volatile("{body}"
                 : {", ".join(outs)}
                 : {", ".join(ins)});
}}""")
    # ---------------- wide add / sub (2N limbs); sub returns the borrow mask
    out.append(f"B200ZK_ASM_DEV void add_wide_{N}(uint32_t* r, const uint32_t* a, const uint32_t* b) {{\n" +
               add_chain_stmt([f"r[{k}]" for k in range(2 * N)], [f"a[{k}]" for k in range(2 * N)], [f"b[{k}]" for k in range(2 * N)]) + "\n}")
    n = 2 * N
    lines = []
    for k in range(n):
        lines.append(f"{'sub.cc.u32' if k == 0 else 'subc.cc.u32'} %{k}, %{n + 1 + k}, %{2 * n + 1 + k};")
    lines.append(f"subc.u32 %{n}, 0, 0;")
    body = "\\n\\t".join(lines)
    outs = ", ".join([f'"=&r"(r[{k}])' for k in range(n)] + ['"=&r"(borrow)'])
    ins = ", ".join([f'"r"(a[{k}])' for k in range(n)] + [f'"r"(b[{k}])' for k in range(n)])
    out.append(f"""B200ZK_ASM_DEV uint32_t sub_wide_{N}(uint32_t* r, const uint32_t* a, const uint32_t* b) {{
    uint32_t borrow;
    asm volatile("{body}"
                 : {outs}
                 : {ins});
    return borrow;
}}""")
    return "\n\n".join(out)

HEADER = """// GENERATED by tools/gen_field_asm.py -- do not edit.  One inline-PTX statement per carry chain.
#pragma once
#include <cstdint>
#if defined(__CUDA_ARCH__)
namespace b200zk {
namespace ptx {
#define B200ZK_ASM_DEV __device__ __forceinline__

"""
FOOTER = """
// dispatch on the limb count
template <int N> B200ZK_ASM_DEV void row_even(uint32_t* X, const uint32_t* a, uint32_t b);
template <> B200ZK_ASM_DEV void row_even<8>(uint32_t* X, const uint32_t* a, uint32_t b) { row_even_8(X, a, b); }
template <> B200ZK_ASM_DEV void row_even<12>(uint32_t* X, const uint32_t* a, uint32_t b) { row_even_12(X, a, b); }
template <int N> B200ZK_ASM_DEV void row_odd(uint32_t* Y, const uint32_t* a, uint32_t b);
template <> B200ZK_ASM_DEV void row_odd<8>(uint32_t* Y, const uint32_t* a, uint32_t b) { row_odd_8(Y, a, b); }
template <> B200ZK_ASM_DEV void row_odd<12>(uint32_t* Y, const uint32_t* a, uint32_t b) { row_odd_12(Y, a, b); }
template <int N> B200ZK_ASM_DEV void row_odd_cin(uint32_t& nx0, uint32_t y0, uint32_t x1, uint32_t* Y, const uint32_t* a, uint32_t b);
template <> B200ZK_ASM_DEV void row_odd_cin<8>(uint32_t& nx0, uint32_t y0, uint32_t x1, uint32_t* Y, const uint32_t* a, uint32_t b) { row_odd_cin_8(nx0, y0, x1, Y, a, b); }
template <> B200ZK_ASM_DEV void row_odd_cin<12>(uint32_t& nx0, uint32_t y0, uint32_t x1, uint32_t* Y, const uint32_t* a, uint32_t b) { row_odd_cin_12(nx0, y0, x1, Y, a, b); }
template <int N> B200ZK_ASM_DEV void add_n(uint32_t* r, const uint32_t* a, const uint32_t* b);
template <> B200ZK_ASM_DEV void add_n<8>(uint32_t* r, const uint32_t* a, const uint32_t* b) { add_8(r, a, b); }
template <> B200ZK_ASM_DEV void add_n<12>(uint32_t* r, const uint32_t* a, const uint32_t* b) { add_12(r, a, b); }
template <int N> B200ZK_ASM_DEV uint32_t sub_n(uint32_t* r, const uint32_t* a, const uint32_t* b);
template <> B200ZK_ASM_DEV uint32_t sub_n<8>(uint32_t* r, const uint32_t* a, const uint32_t* b) { return sub_8(r, a, b); }
template <> B200ZK_ASM_DEV uint32_t sub_n<12>(uint32_t* r, const uint32_t* a, const uint32_t* b) { return sub_12(r, a, b); }
template <int N> B200ZK_ASM_DEV void mul_wide(uint32_t* T, const uint32_t* a, const uint32_t* b);
template <> B200ZK_ASM_DEV void mul_wide<8>(uint32_t* T, const uint32_t* a, const uint32_t* b) { mul_wide_8(T, a, b); }
template <> B200ZK_ASM_DEV void mul_wide<12>(uint32_t* T, const uint32_t* a, const uint32_t* b) { mul_wide_12(T, a, b); }
template <int N> B200ZK_ASM_DEV void sqr_wide(uint32_t* T, const uint32_t* a);
template <> B200ZK_ASM_DEV void sqr_wide<8>(uint32_t* T, const uint32_t* a) { sqr_wide_8(T, a); }
template <> B200ZK_ASM_DEV void sqr_wide<12>(uint32_t* T, const uint32_t* a) { sqr_wide_12(T, a); }
template <int N> B200ZK_ASM_DEV void redc_step(uint32_t& nx0, uint32_t& m, uint32_t y0, uint32_t x1, uint32_t* Y, const uint32_t* p, uint32_t inv);
template <> B200ZK_ASM_DEV void redc_step<8>(uint32_t& nx0, uint32_t& m, uint32_t y0, uint32_t x1, uint32_t* Y, const uint32_t* p, uint32_t inv) { redc_step_8(nx0, m, y0, x1, Y, p, inv); }
template <> B200ZK_ASM_DEV void redc_step<12>(uint32_t& nx0, uint32_t& m, uint32_t y0, uint32_t x1, uint32_t* Y, const uint32_t* p, uint32_t inv) { redc_step_12(nx0, m, y0, x1, Y, p, inv); }
template <int N> B200ZK_ASM_DEV void add_wide(uint32_t* r, const uint32_t* a, const uint32_t* b);
template <> B200ZK_ASM_DEV void add_wide<8>(uint32_t* r, const uint32_t* a, const uint32_t* b) { add_wide_8(r, a, b); }
template <> B200ZK_ASM_DEV void add_wide<12>(uint32_t* r, const uint32_t* a, const uint32_t* b) { add_wide_12(r, a, b); }
template <int N> B200ZK_ASM_DEV uint32_t sub_wide(uint32_t* r, const uint32_t* a, const uint32_t* b);
template <> B200ZK_ASM_DEV uint32_t sub_wide<8>(uint32_t* r, const uint32_t* a, const uint32_t* b) { return sub_wide_8(r, a, b); }
template <> B200ZK_ASM_DEV uint32_t sub_wide<12>(uint32_t* r, const uint32_t* a, const uint32_t* b) { return sub_wide_12(r, a, b); }

#undef B200ZK_ASM_DEV
}  // namespace ptx
}  // namespace b200zk
#endif
"""

if __name__ == "__main__":
    path = sys.argv[1] if len(sys.argv) > 1 else "zk-apps_b200/csrc/field_asm.cuh"
    with open(path, "w") as f:
        f.write(HEADER + gen(8) + "\n\n" + gen(12) + "\n\n" + gen_wide(8) + "\n\n" + gen_wide(12) + "\n" + FOOTER)
    print("wrote", path)
