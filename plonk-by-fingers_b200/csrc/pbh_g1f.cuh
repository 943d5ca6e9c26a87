// pbh_g1f.cuh — G1, GT and the pairing over F_101 as exact small-integer FP32 arithmetic (PBH_ALGO_ARITH).
//
// Why: the integer version of this arithmetic (pbh_arith.cuh) is bound by the ALU pipe — a complete affine addition spends
// more instructions on compares, selects and the shifts of its reductions than on multiplies (profiles/r01i: 574 SEL, 451 SHF
// and ~430 ISETP of 3964 instructions per verification).  Here
//   * residues are floats, a reduction mod 101 is three FMA-pipe instructions (FFMA with the 1.5 * 2^23 rounding constant,
//     FADD, FFMA) and differences need no correction term, because centred residues may be negative;
//   * ONE slope formula serves addition and doubling:  lambda = (x1^2 + x1 x2 + x2^2) / (y1 + y2).  It equals the chord slope
//     (y2 - y1) / (x2 - x1) whenever x1 != x2 (multiply both by (y2 + y1)(x2 - x1) and use y^2 = x^3 + 3) and the tangent slope
//     3 x^2 / 2 y when the points are equal.  On THIS curve it never fails for a non-trivial sum: 101 = 2 (mod 3), so cubing
//     is a bijection of F_101 and y1 = -y2 forces x1^3 = x2^3, i.e. x1 = x2, i.e. P = -Q (or P = Q of order two).  Hence
//     y1 + y2 = 0  <=>  the sum is the identity, and the reference's case analysis (src/pbh/g1.rs:119-144: self.inf, rhs.inf,
//     self == -rhs, self == rhs, chord) collapses to one table lookup whose zero entry is the identity flag.
// Both operands must be ON THE CURVE (or flagged identities): every caller checks that first — the verifier's Step 1
// (src/plonk.rs:523-534) runs before any arithmetic and decides the result alone when it fails; the sweep kernels flag
// off-curve inputs.  Results are the group sum, the same function the reference computes, with the canonical identity
// (0, 0, infinite) of src/pbh/g1.rs:83-89.
//
// Exactness: every value is an integer below 2^21 in magnitude (red101 is exact there: checked for every integer of the
// range by the CPU test suite, and the bounds of every expression are machine-checked by instantiating these templates
// with the magnitude-propagating scalar type of the CPU test suite).
#pragma once
#include "pbh_arith.cuh"
#include "pbh_f32.cuh"

namespace pbh {

// ---- FP32 primitives for F_101 (plain float policy; the CPU test suite adds a bound-propagating policy) ----------------
// centred residue in [-50, 50] of an exact integer |x| <= 2^21
PBH_HD F32 f_red101(F32 x) {
  float t = fmaf(x.v, 0.009900990099009901f, 12582912.0f);
  float q = t - 12582912.0f;
  return F32(fmaf(q, -101.0f, x.v));
}
// 1 / s (mod 101) as a centred float, 0 when s = 0 (mod 101), for an integer |s| <= 202: one lookup, no range reduction
// of the index
PBH_HD F32 f_inv101(F32 s, const float* inv101c) {
#if defined(__CUDA_ARCH__)
  return F32(inv101c[(uint32_t)__float_as_int(s.v + (12582912.0f + 202.0f)) & 0x1FFu]);
#else
  return F32(inv101c[(int)s.v + 202]);
#endif
}
PBH_HD F32 f_sel(bool c, F32 a, F32 b) { return c ? a : b; }
PBH_HD F32 f_neg(F32 a) { return F32(-a.v); }
// canonical residue 0..100 of a centred one
PBH_HD uint32_t f_canon101(F32 x) {
  float c = x.v < 0.0f ? x.v + 101.0f : x.v;
#if defined(__CUDA_ARCH__)
  return (uint32_t)__float_as_int(c + 12582912.0f) & 0xFFu;
#else
  return (uint32_t)(int)c;
#endif
}
// byte k of a packed word as an exact float (one PRMT that also plants the exponent, one FADD)
PBH_HD F32 f_from_byte(uint32_t w, int k, F32*) {
#if defined(__CUDA_ARCH__)
  return F32(__int_as_float((int)__byte_perm(w, 0x4B000000u, 0x7440u + (uint32_t)k)) - 8388608.0f);
#else
  return F32((float)((w >> (8 * k)) & 0xFFu));
#endif
}
PBH_HD bool f_eq(F32 a, F32 b) { return a.v == b.v; }
// an exact non-negative integer below 2^13 as a table index
PBH_HD uint32_t f_to_index(F32 x) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__float_as_int(x.v + 12582912.0f) & 0x1FFFu;
#else
  return (uint32_t)(int)x.v;
#endif
}

// ---- G1 ------------------------------------------------------------------------------------------------------------
template <class T>
struct G1F {
  T x, y;      // any representative with |.| <= 100; (0, 0) for the identity
  bool inf;
};
template <class T> PBH_HD G1F<T> g1f_identity() { G1F<T> r; T* tag = nullptr; r.x = f_const(0.f, tag); r.y = f_const(0.f, tag); r.inf = true; return r; }
// packed x | y << 8 | inf << 16 (Tables::pt17, PairTables) -> point
template <class T> PBH_HD G1F<T> g1f_unpack(uint32_t w) {
  G1F<T> r; T* tag = nullptr;
  r.x = f_from_byte(w, 0, tag); r.y = f_from_byte(w, 1, tag); r.inf = (w & 0x10000u) != 0u;
  return r;
}
template <class T> PBH_HD uint32_t g1f_pack(const G1F<T>& p) { return f_canon101(p.x) | (f_canon101(p.y) << 8) | (p.inf ? 0x10000u : 0u); }
template <class T> PBH_HD G1F<T> g1f_neg(const G1F<T>& p) { G1F<T> r = p; r.y = f_neg(p.y); return r; }
// src/pbh/g1.rs:63-65 (ignores the flag, Q9): y^2 == x^3 + 3 for raw coordinates 0..100
template <class T> PBH_HD bool g1f_in_curve(T x, T y) {
  T* tag = nullptr;
  T x2 = f_mul(x, x);
  T d = f_sub(f_sub(f_mul(y, y), f_mul(x2, x)), f_const(3.f, tag));     // |d| < 2^21
  return f_is_zero(f_red101(d));
}

// P + Q for curve points (header comment): 20 FP32-pipe instructions, one lookup, the selects of the identity cases
template <class T>
PBH_HD G1F<T> g1f_add(const G1F<T>& p, const G1F<T>& q, const float* inv101c) {
  T* tag = nullptr;
  const T sinv = f_inv101(f_add(p.y, q.y), inv101c);                    // 0  <=>  P = -Q
  const T t = f_add(p.x, q.x);
  const T n = f_fma(p.x, p.x, f_mul(q.x, t));                           // x1^2 + x1 x2 + x2^2, <= 3 * 10^4
  const T lam = f_red101(f_mul(n, sinv));                               // |n sinv| <= 1.5 * 10^6
  const T x3 = f_red101(f_sub(f_mul(lam, lam), t));
  const T y3 = f_red101(f_sub(f_mul(lam, f_sub(p.x, x3)), p.y));
  const bool zero = f_is_zero(sinv);
  const T o = f_const(0.f, tag);
  G1F<T> r;
  // q flagged -> p as it is (so q may be a flagged point that still carries coordinates: the "no addition" of
  // g1f_add_if and of Straus); p must be canonical, i.e. (0, 0) when flagged
  r.x = f_sel(q.inf, p.x, f_sel(p.inf, q.x, f_sel(zero, o, x3)));
  r.y = f_sel(q.inf, p.y, f_sel(p.inf, q.y, f_sel(zero, o, y3)));
  r.inf = q.inf ? p.inf : (p.inf ? false : zero);
  return r;
}
// P + (take ? Q : identity): the conditional additions of double-and-add and of Straus cost one flag, not three selects
template <class T>
PBH_HD G1F<T> g1f_add_if(const G1F<T>& p, const G1F<T>& q, bool take, const float* inv101c) {
  G1F<T> qq = q;
  qq.inf = q.inf || !take;
  return g1f_add(p, qq, inv101c);
}

// [k]P for k < 2^BITS, MSB first (src/pbh/g1.rs:146-168 computes the same group multiple LSB first)
template <int BITS, class T>
PBH_HD G1F<T> g1f_smul(const G1F<T>& p_in, uint32_t k, const float* inv101c) {
  G1F<T> p = p_in;
  if (p.inf) p = g1f_identity<T>();
  G1F<T> r = p;
  r.inf = p.inf || !((k >> (BITS - 1)) & 1u);
  if (r.inf) r = g1f_identity<T>();
#pragma unroll
  for (int j = BITS - 2; j >= 0; j--) {
    r = g1f_add(r, r, inv101c);
    r = g1f_add_if(r, p, ((k >> j) & 1u) != 0u, inv101c);
  }
  return r;
}

// ---- GT = F_101[u]/(u^2 + 2) -----------------------------------------------------------------------------------------
template <class T> struct GTF { T a, b; };
// src/pbh/gt.rs:61-69
template <class T> PBH_HD GTF<T> gtf_mul(const GTF<T>& p, const GTF<T>& q) {
  T* tag = nullptr;
  GTF<T> r;
  r.a = f_red101(f_fma(f_const(-2.f, tag), f_mul(p.b, q.b), f_mul(p.a, q.a)));
  r.b = f_red101(f_fma(p.a, q.b, f_mul(p.b, q.a)));
  return r;
}
template <class T> PBH_HD GTF<T> gtf_sqr(const GTF<T>& p) {
  T* tag = nullptr;
  GTF<T> r;
  r.a = f_red101(f_fma(f_const(-2.f, tag), f_mul(p.b, p.b), f_mul(p.a, p.a)));
  r.b = f_red101(f_mul(f_add(p.a, p.a), p.b));
  return r;
}
// x^600 = (conj(x)^2 / norm(x))^6, 0 -> 0: GTP::pow(600) of src/pbh/gt.rs:33-59 on all of F_101^2 (pbh_arith.cuh gt_final_exp)
template <class T> PBH_HD GTF<T> gtf_final_exp(const GTF<T>& f, const float* inv101c) {
  T* tag = nullptr;
  const T norm = f_red101(f_fma(f_const(2.f, tag), f_mul(f.b, f.b), f_mul(f.a, f.a)));
  const T ninv = f_inv101(norm, inv101c);
  GTF<T> c; c.a = f.a; c.b = f_neg(f.b);
  const GTF<T> c2 = gtf_sqr(c);
  GTF<T> x; x.a = f_red101(f_mul(c2.a, ninv)); x.b = f_red101(f_mul(c2.b, ninv));    // f^100
  const GTF<T> x2 = gtf_sqr(x);
  const GTF<T> x3 = gtf_mul(x2, x);
  return gtf_sqr(x3);
}
// The line through a and b at Q = (qa, qb u): src/pbh/pairing.rs:25-34, 41, 45 on the raw coordinates (identity = (0, 0), Q10)
template <class T> PBH_HD GTF<T> millerf_line(const G1F<T>& a, const G1F<T>& b, T qa, T nqb) {
  const T m = f_sub(b.x, a.x), n = f_sub(b.y, a.y);
  const T c = f_sub(f_mul(m, a.y), f_mul(n, a.x));
  GTF<T> r;
  r.a = f_red101(f_fma(qa, n, c));
  r.b = f_red101(f_mul(m, nqb));          // nqb = -qb
  return r;
}
// pairing_f(17, P, Q), the recursion of src/pbh/pairing.rs:23-47 unrolled (pbh_arith.cuh miller_f17)
template <class T> PBH_HD GTF<T> millerf_f17(const G1F<T>& p_in, T qa, T qb, const float* inv101c) {
  G1F<T> p = p_in;
  if (p.inf) p = g1f_identity<T>();        // G1P * k returns the canonical identity for a flagged point (g1.rs:148-150)
  const T nqb = f_neg(qb);
  const G1F<T> p2 = g1f_add(p, p, inv101c), p4 = g1f_add(p2, p2, inv101c), p8 = g1f_add(p4, p4, inv101c), p16 = g1f_add(p8, p8, inv101c);
  GTF<T> f = millerf_line(p, g1f_neg(p2), qa, nqb);
  f = gtf_mul(gtf_sqr(f), millerf_line(p2, g1f_neg(p4), qa, nqb));
  f = gtf_mul(gtf_sqr(f), millerf_line(p4, g1f_neg(p8), qa, nqb));
  f = gtf_mul(gtf_sqr(f), millerf_line(p8, g1f_neg(p16), qa, nqb));
  f = gtf_mul(f, millerf_line(p16, p_in, qa, nqb));   // the reference passes the caller's P itself here (pairing.rs:40)
  return f;
}
// src/pbh/pairing.rs:12-20; qa, qb: coordinates of Q (0..100)
template <class T> PBH_HD GTF<T> pairingf(const G1F<T>& p, T qa, T qb, const float* inv101c) {
  return gtf_final_exp(millerf_f17(p, qa, qb, inv101c), inv101c);
}

}  // namespace pbh
