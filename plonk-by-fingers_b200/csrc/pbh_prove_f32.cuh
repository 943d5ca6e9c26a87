// pbh_prove_f32.cuh — the prover's F_17 arithmetic on the FP32 FMA pipes.
//
// Why floating point for integer residues: on sm_100a an IMAD issues on one FMA pipe only (measured 64
// thread-ops/clk/SM) while FFMA issues on both (measured 124/clk/SM with a constant operand, 93 with three register
// operands) — see profiles/r01_pipe_rates.txt.  Every value here is a small integer, so FP32 arithmetic is EXACT as
// long as magnitudes stay below 2^24; the proof is bit-identical to the integer path.
//
// Representation: centred residues in [-8, 8].  red17(x) = x - 17*rint(x/17) is three full-rate ops (FFMA with the
// 1.5*2^23 rounding constant, FADD, FFMA) and is valid for |x| <= 2^23 (the error of the fp32 reciprocal times |x|
// stays below the 1/34 margin to the nearest rounding boundary).  Reductions are lazy: sums of products run
// unreduced through whole polynomial products.  Every bound is machine-checked: the core below is templated on the
// scalar type, and the CPU test suite instantiates it with a type that propagates worst-case magnitudes and fails if
// any operation could leave the exact range (test_f32_bounds).
//
// The reference's SubAssign quirk Q1 (src/poly.rs:192-203) is reproduced inside the core: it can only act when the top
// coefficient alpha b1 b3 b5 b7 of t1 + t2 vanishes, which a warp vote detects, so the common case pays one reduction and
// one vote for it.  (Round 1 sent such items to the exact-length integer routine of pbh_prove.cuh, which made every warp
// of a uniformly drawn batch run both routines: 146 us per 2^20 proofs against 74 us on a full-path batch.)
#pragma once
#include <math.h>

#include "pbh_prove.cuh"
#include "pbh_f32.cuh"
#include "pbh_g1f.cuh"

namespace pbh {

// schoolbook products, unreduced, in outer-product order: consecutive FFMAs share the multiplicand a[i] (operand
// reuse cache, fewer register-bank conflicts) and write different accumulators (independent chains)
#ifndef PBH_MAC_ORDER
#define PBH_MAC_ORDER 0
#endif
template <class T, int LA, int LB>
PBH_HD void fpoly_mul(const T (&a)[LA], const T (&b)[LB], T (&out)[LA + LB - 1]) {
#if PBH_MAC_ORDER == 0
#pragma unroll
  for (int i = 0; i < LA; i++) {
#pragma unroll
    for (int j = 0; j < LB; j++) {
      // out[i + j] is touched for the first time when i == 0 or j == LB - 1
      out[i + j] = (i == 0 || j == LB - 1) ? f_mul(a[i], b[j]) : f_fma(a[i], b[j], out[i + j]);
    }
  }
#elif PBH_MAC_ORDER == 1
#pragma unroll
  for (int k = 0; k < LA + LB - 1; k++) {
    bool first = true;
#pragma unroll
    for (int i = 0; i < LA; i++) {
      if (k - i >= 0 && k - i < LB) { out[k] = first ? f_mul(a[i], b[k - i]) : f_fma(a[i], b[k - i], out[k]); first = false; }
    }
  }
#else
  // column order: fixed b[j], varying a[i]
#pragma unroll
  for (int j = 0; j < LB; j++) {
#pragma unroll
    for (int i = 0; i < LA; i++) {
      out[i + j] = (j == 0 || i == LA - 1) ? f_mul(a[i], b[j]) : f_fma(a[i], b[j], out[i + j]);
    }
  }
#endif
}
template <class T, int LA, int LB, int LO>
PBH_HD void fpoly_mac(const T (&a)[LA], const T (&b)[LB], T (&acc)[LO]) {
#pragma unroll
  for (int i = 0; i < LA; i++) {
#pragma unroll
    for (int j = 0; j < LB; j++) acc[i + j] = f_fma(a[i], b[j], acc[i + j]);
  }
}

// interpolate_at_h with centred matrix entries: 13 -> -4, 16 -> -1  (h_pows_inv of SURVEY.md §3.1)
template <class T>
PBH_HD void fintt4(T v0, T v1, T v2, T v3, T (&f)[4]) {
  T* tag = nullptr;
  const T m4 = f_const(-4.f, tag), m1 = f_const(-1.f, tag), p4 = f_const(4.f, tag);
  f[0] = f_red(f_mul(m4, f_add(f_add(v0, v1), f_add(v2, v3))));
  f[1] = f_red(f_fma(m4, v0, f_fma(m1, v1, f_fma(p4, v2, v3))));
  f[2] = f_red(f_fma(m4, f_add(v0, v2), f_mul(p4, f_add(v1, v3))));
  f[3] = f_red(f_fma(m4, v0, f_fma(p4, v2, f_fma(m1, v3, v1))));
}

// Circuit / SRS constants as centred floats (filled by the host from Consts)
struct ConstsF {
  float q_l[4], q_o[4], q_m[4], q_c[4];        // selector vectors (satisfies uses q_l twice: Q8)
  float QL[4], QR[4], QO[4], QM[4], QC[4];     // interpolated selectors
  float sig[3][4], S[3][4], L1[4];
  float srs_dlog[10];
  float vdlog[8];                              // verifier: dlogs of the 8 constant commitments (pbh_verify.cuh)
};

// Where the core reads circuit / SRS constants from.  RuntimeCK: the context's ConstsF (any 4-gate circuit, any SRS).
// PbhCK: the reference's own circuit (src/pbh/mod.rs:56-67) with SRS::create(2, 6) as compile-time constants, so zero
// coefficients vanish, +-1 become additions and the SRS-bounds checks fold away; used only when the context's constants
// are exactly these (checked at context creation).
struct RuntimeCK {
  const ConstsF& F;
  uint32_t npts;
  PBH_HD float QL(int i) const { return F.QL[i]; }
  PBH_HD float QR(int i) const { return F.QR[i]; }
  PBH_HD float QO(int i) const { return F.QO[i]; }
  PBH_HD float QM(int i) const { return F.QM[i]; }
  PBH_HD float QC(int i) const { return F.QC[i]; }
  PBH_HD float sig(int w, int i) const { return F.sig[w][i]; }
  PBH_HD float S(int w, int i) const { return F.S[w][i]; }
  PBH_HD float L1(int i) const { return F.L1[i]; }
  PBH_HD float srs_dlog(int i) const { return F.srs_dlog[i]; }
  PBH_HD uint32_t n_pts() const { return npts; }
};
struct PbhCK {
  PBH_HD constexpr float QL(int i) const { return i == 0 ? -4.f : (i == 1 ? 1.f : (i == 2 ? 4.f : -1.f)); }
  PBH_HD constexpr float QR(int i) const { return QL(i); }
  PBH_HD constexpr float QO(int i) const { return i == 0 ? -1.f : 0.f; }
  PBH_HD constexpr float QM(int i) const { return i == 0 ? 5.f : (i == 1 ? -1.f : (i == 2 ? -4.f : 1.f)); }
  PBH_HD constexpr float QC(int) const { return 0.f; }
  PBH_HD constexpr float sig(int w, int i) const {
    return w == 0 ? (i == 0 ? 2.f : (i == 1 ? 8.f : (i == 2 ? -2.f : 3.f)))
         : w == 1 ? (i == 0 ? 1.f : (i == 1 ? 4.f : (i == 2 ? -1.f : -5.f)))
                  : (i == 0 ? -4.f : (i == 1 ? -8.f : (i == 2 ? 5.f : -3.f)));
  }
  PBH_HD constexpr float S(int w, int i) const {
    return w == 0 ? (i == 0 ? 7.f : (i == 1 ? -4.f : (i == 2 ? -7.f : 6.f)))
         : w == 1 ? (i == 0 ? 4.f : (i == 1 ? 0.f : (i == 2 ? -4.f : 1.f)))
                  : (i == 0 ? 6.f : (i == 1 ? 7.f : (i == 2 ? 3.f : -3.f)));
  }
  PBH_HD constexpr float L1(int) const { return -4.f; }
  PBH_HD constexpr float srs_dlog(int i) const {
    return i == 0 ? 1.f : (i == 1 ? 2.f : (i == 2 ? 4.f : (i == 3 ? 8.f : (i == 4 ? -1.f : (i == 5 ? -2.f : (i == 6 ? -4.f : 0.f))))));
  }
  PBH_HD constexpr uint32_t n_pts() const { return 7u; }
};
// true when a context's float constants are exactly PbhCK's
inline bool consts_match_pbh(const ConstsF& F, uint32_t n_pts) {
  PbhCK c;
  bool ok = n_pts == 7;
  for (int i = 0; i < 4; i++) {
    ok = ok && F.QL[i] == c.QL(i) && F.QR[i] == c.QR(i) && F.QO[i] == c.QO(i) && F.QM[i] == c.QM(i) && F.QC[i] == c.QC(i) && F.L1[i] == c.L1(i);
    for (int w = 0; w < 3; w++) ok = ok && F.sig[w][i] == c.sig(w, i) && F.S[w][i] == c.S(w, i);
  }
  for (int i = 0; i < 10; i++) ok = ok && F.srs_dlog[i] == c.srs_dlog(i);
  return ok;
}

struct ProofF {
  uint32_t e[9];     // a_s b_s c_s z_s t_lo_s t_mid_s t_hi_s w_z_s w_z_omega_s: discrete logs base G (TABLE) or packed points (ARITH)
  uint32_t ev[7];    // canonical evaluations
};

// sum_i [c_i] g1s[i] from canonical coefficients through the three-point tables: ceil(L/3) lookups, one addition fewer
template <int L>
PBH_HD uint32_t commit_pairs_f32(const uint32_t (&ci)[L], const Tables& Tb) {
  G1F<F32> acc = g1f_unpack<F32>(pair_lookup(Tb.fixed->srs_tri[0], tri_index<L, 0>(ci)));
#pragma unroll
  for (int j = 1; j < (L + 2) / 3; j++) {
    const uint32_t idx = j == 1 ? tri_index<L, 1>(ci) : (j == 2 ? tri_index<L, 2>(ci) : tri_index<L, 3>(ci));
    acc = g1f_add(acc, g1f_unpack<F32>(pair_lookup(Tb.fixed->srs_tri[j], idx)), Tb.inv101c);
  }
  return g1f_pack(acc);
}

// SRS::eval_at_s of a coefficient array (src/plonk.rs:51-58); `oob` is set when a coefficient at or beyond n_pts is
// non-zero (the reference then indexes g1s out of bounds, src/plonk.rs:56).
//   ALGO_TABLE: dot product with the SRS discrete logs -> exponent of G (see commit<> of pbh_prove.cuh); returns 0..16.
//   ALGO_ARITH: per-term fixed-base multiples [c_i]g1s[i] added with the affine group law; returns the packed point.
template <int ALGO, class T, class CK, int L>
PBH_HD uint32_t fcommit(const T (&c)[L], const CK& ck, const Tables& Tb, bool reduced, bool& oob) {
  T* tag = nullptr;
  const uint32_t n_pts = ck.n_pts();
  if (n_pts < (uint32_t)L) {                       // uniform, false for the usual 7-point SRS except for w_z
#pragma unroll
    for (int j = 0; j < L; j++)
      if ((uint32_t)j >= n_pts) oob = oob | !f_is_zero(reduced ? c[j] : f_red(c[j]));
  }
  if (ALGO == ALGO_TABLE) {
    T e = f_mul(c[0], f_const(ck.srs_dlog(0), tag));
#pragma unroll
    for (int i = 1; i < L; i++) e = f_fma(c[i], f_const(ck.srs_dlog(i), tag), e);
    return f_canon(f_red(e));
  } else {
    // three coefficients per lookup (FixedBaseTables), the first lookup needs no addition; exact FP32 curve arithmetic
    // (pbh_g1f.cuh: the table entries are multiples of SRS points, all on the curve)
    static_assert(L >= 2 && L <= 10, "the SRS tables cover 10 points");
    T cr[L];
#pragma unroll
    for (int i = 0; i < L; i++) cr[i] = reduced ? c[i] : f_red(c[i]);
    // table index of a triple straight from the centred residues: (c0 + 8) + 17 (c1 + 8) + 289 (c2 + 8)
    auto index = [&](int j) -> uint32_t {
      T x = f_add(cr[3 * j], f_const(2456.f, tag));
      if (3 * j + 1 < L) x = f_fma(cr[(3 * j + 1 < L) ? 3 * j + 1 : 0], f_const(17.f, tag), x);
      if (3 * j + 2 < L) x = f_fma(cr[(3 * j + 2 < L) ? 3 * j + 2 : 0], f_const(289.f, tag), x);
      return f_to_index(x);
    };
    G1F<F32> acc = g1f_unpack<F32>(pair_lookup(Tb.fixed->srs_tri[0], index(0)));
#pragma unroll
    for (int j = 1; j < (L + 2) / 3; j++) acc = g1f_add(acc, g1f_unpack<F32>(pair_lookup(Tb.fixed->srs_tri[j], index(j))), Tb.inv101c);
    return g1f_pack(acc);
  }
}

// w[12], rnd[9], ch[5]: inputs as exact small integers (0..16).  inv17c: centred inverses as floats, indexed by the
// canonical residue.  Returns the status byte among {0, 2, 3, 4, 5} (satisfiability, status 1, is the caller's).
// The challenges come from `cs` (pbh_fs.cuh) as values 0..16 at the point where the reference first uses each.
template <int ALGO, class T, class CK, class CS>
PBH_HD uint32_t prove_core_f32_cs(const T (&w)[12], const T (&rnd_in)[9], CS& cs, const CK& ck, const Tables& Tb,
                                  const float* inv17c, ProofF& P) {
  T* tag = nullptr;
  // blinders and challenges are used as they come (0..16); the bound check shows that centring them is not needed
  // (worst-case magnitude 4.8 M, below red17's 2^23 range)
  const T (&rnd)[9] = rnd_in;
  // packed point of a commitment, for the challenge sources that hash it
  auto point_of = [&](uint32_t e) -> uint32_t { return CS::kNeedsPoints ? ((ALGO == ALGO_TABLE) ? Tb.pt17[e] : e) : 0u; };

  // ---- wire polynomials                                                     src/plonk.rs:233-235, 248-252
  T fa[4], fb[4], fc[4];
  fintt4(w[0], w[1], w[2], w[3], fa);
  fintt4(w[4], w[5], w[6], w[7], fb);
  fintt4(w[8], w[9], w[10], w[11], fc);
  T a[6], b[6], c[6];
  a[0] = f_sub(fa[0], rnd[1]); a[1] = f_sub(fa[1], rnd[0]); a[2] = fa[2]; a[3] = fa[3]; a[4] = rnd[1]; a[5] = rnd[0];
  b[0] = f_sub(fb[0], rnd[3]); b[1] = f_sub(fb[1], rnd[2]); b[2] = fb[2]; b[3] = fb[3]; b[4] = rnd[3]; b[5] = rnd[2];
  c[0] = f_sub(fc[0], rnd[5]); c[1] = f_sub(fc[1], rnd[4]); c[2] = fc[2]; c[3] = fc[3]; c[4] = rnd[5]; c[5] = rnd[4];
  bool oob_abc = false, oob_z = false, oob_t = false, oob_w = false;
  P.e[0] = fcommit<ALGO, T>(a, ck, Tb, false, oob_abc);                             // src/plonk.rs:255-257
  P.e[1] = fcommit<ALGO, T>(b, ck, Tb, false, oob_abc);
  P.e[2] = fcommit<ALGO, T>(c, ck, Tb, false, oob_abc);
  T beta, gamma;
  cs.beta_gamma(point_of(P.e[0]), point_of(P.e[1]), point_of(P.e[2]), beta, gamma);

  // ---- accumulator                                                          src/plonk.rs:278-299
  T acc[4];
  acc[0] = f_const(1.f, tag);
  bool div0 = false;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    // omega^i, k1 omega^i, k2 omega^i centred: i=0: 1,2,3; i=1: 4,8,-5; i=2: -1,-2,-3
    const float o1 = (i == 0) ? 1.f : (i == 1 ? 4.f : -1.f), o2 = (i == 0) ? 2.f : (i == 1 ? 8.f : -2.f),
                o3 = (i == 0) ? 3.f : (i == 1 ? -5.f : -3.f);
    T wa = f_add(w[i], gamma), wb = f_add(w[4 + i], gamma), wc = f_add(w[8 + i], gamma);
    // factors stay unreduced (|.| <= 16 + 8 + 8*8 = 88): a product of three is below 2^20
    T n1 = f_fma(beta, f_const(o1, tag), wa), n2 = f_fma(beta, f_const(o2, tag), wb), n3 = f_fma(beta, f_const(o3, tag), wc);
    T d1 = f_fma(beta, f_const(ck.sig(0, i), tag), wa), d2 = f_fma(beta, f_const(ck.sig(1, i), tag), wb),
      d3 = f_fma(beta, f_const(ck.sig(2, i), tag), wc);
    T dsor = f_red(f_mul(f_mul(d1, d2), d3));
    div0 = div0 | f_is_zero(dsor);                                           // src/plonk.rs:297 unwrap
    T dinv = f_const(inv17c[f_canon(dsor)], tag);
    T dend = f_red(f_mul(f_mul(n1, n2), n3));
    acc[i + 1] = f_red(f_mul(f_mul(acc[i], dinv), dend));                    // |.| <= 8^3
  }
  T accx[4];
  fintt4(acc[0], acc[1], acc[2], acc[3], accx);
  T z[7];
  z[0] = f_sub(accx[0], rnd[8]); z[1] = f_sub(accx[1], rnd[7]); z[2] = f_sub(accx[2], rnd[6]); z[3] = accx[3];
  z[4] = rnd[8]; z[5] = rnd[7]; z[6] = rnd[6];
  P.e[3] = fcommit<ALGO, T>(z, ck, Tb, false, oob_z);                               // src/plonk.rs:313
  const T alpha = cs.alpha(point_of(P.e[3]));

  // ---- quotient numerator: t1 + alpha (A'B'C' z - A''B''C'' z_omega) + alpha^2 (z - 1) L1     src/plonk.rs:339-369
  T num[22];
  const T a2 = f_red(f_mul(alpha, alpha));
  {
    // t1 = a b q_m + a q_l + b q_r + c q_o + q_c
    T ab[11];
    fpoly_mul(a, b, ab);
    T qm[4], ql[4], qr[4], qo[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { qm[i] = f_const(ck.QM(i), tag); ql[i] = f_const(ck.QL(i), tag); qr[i] = f_const(ck.QR(i), tag); qo[i] = f_const(ck.QO(i), tag); }
    T t1[14];
    fpoly_mul(ab, qm, t1);
    fpoly_mac(a, ql, t1);
    fpoly_mac(b, qr, t1);
    fpoly_mac(c, qo, t1);
#pragma unroll
    for (int i = 0; i < 4; i++) t1[i] = f_add(t1[i], f_const(ck.QC(i), tag));
#pragma unroll
    for (int i = 0; i < 22; i++) num[i] = (i < 14) ? t1[i] : f_const(0.f, tag);
  }
  T zw[7];   // z(omega x): coefficient i times omega^i = 1, 4, -1, -4, 1, 4, -1
  zw[0] = z[0]; zw[1] = f_mul(z[1], f_const(4.f, tag)); zw[2] = f_sub(f_const(0.f, tag), z[2]); zw[3] = f_mul(z[3], f_const(-4.f, tag));
  zw[4] = z[4]; zw[5] = f_mul(z[5], f_const(4.f, tag)); zw[6] = f_sub(f_const(0.f, tag), z[6]);
  {
    // t2' = (a + gamma + beta x)(b + gamma + 2 beta x)(c + gamma + 3 beta x) z
    T A[6], B[6], C[6];
#pragma unroll
    for (int i = 0; i < 6; i++) { A[i] = a[i]; B[i] = b[i]; C[i] = c[i]; }
    A[0] = f_add(A[0], gamma); A[1] = f_add(A[1], beta);
    B[0] = f_add(B[0], gamma); B[1] = f_fma(beta, f_const(2.f, tag), B[1]);
    C[0] = f_add(C[0], gamma); C[1] = f_fma(beta, f_const(3.f, tag), C[1]);
    T AB[11], ABC[16], t2[22];
    fpoly_mul(A, B, AB);
#pragma unroll
    for (int i = 0; i < 11; i++) AB[i] = f_red(AB[i]);
    fpoly_mul(AB, C, ABC);
    fpoly_mul(ABC, z, t2);
#pragma unroll
    for (int i = 0; i < 22; i++) num[i] = f_fma(alpha, t2[i], num[i]);
  }
  // Q1 (src/poly.rs:192-203): `t1 + t2 -= t3` pushes the coefficients of t3 at or beyond len(t1 + t2) UN-NEGATED.  num holds
  // t1 + t2 here.  Its top coefficient is alpha b1 b3 b5 b7, so the quirk needs a zero among those five (27 % of uniformly
  // drawn inputs, none of a full-path batch); the scan below finds len(t1 + t2) and runs only in warps that hold such an item.
  uint32_t unnegated = 0u;        // bit n: coefficient n of t3 is added, not subtracted
  const bool q1_warp = f_any(f_is_zero(f_red(num[21])), tag);
  if (q1_warp) {
    bool tail_zero = true;        // (t1 + t2)[n..21] all zero; the zero polynomial keeps one coefficient
#pragma unroll
    for (int n = 21; n >= 1; n--) {
      tail_zero = tail_zero && f_is_zero(f_red(num[n]));
      unnegated |= tail_zero ? (1u << n) : 0u;
    }
  }
  {
    // t3' = (a + beta S1 + gamma)(b + beta S2 + gamma)(c + beta S3 + gamma) z(omega x)
    T A[6], B[6], C[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
      A[i] = (i < 4) ? f_fma(beta, f_const(ck.S(0, i), tag), a[i]) : a[i];
      B[i] = (i < 4) ? f_fma(beta, f_const(ck.S(1, i), tag), b[i]) : b[i];
      C[i] = (i < 4) ? f_fma(beta, f_const(ck.S(2, i), tag), c[i]) : c[i];
    }
    A[0] = f_add(A[0], gamma); B[0] = f_add(B[0], gamma); C[0] = f_add(C[0], gamma);
#pragma unroll
    for (int i = 0; i < 4; i++) { A[i] = f_red(A[i]); B[i] = f_red(B[i]); C[i] = f_red(C[i]); }
    T AB[11], ABC[16], t3[22];
    fpoly_mul(A, B, AB);
#pragma unroll
    for (int i = 0; i < 11; i++) AB[i] = f_red(AB[i]);
    fpoly_mul(AB, C, ABC);
    fpoly_mul(ABC, zw, t3);
    T nalpha = f_sub(f_const(0.f, tag), alpha);
    if (q1_warp) {
#pragma unroll
      for (int i = 0; i < 22; i++) num[i] = f_fma(((unnegated >> i) & 1u) ? alpha : nalpha, t3[i], num[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 22; i++) num[i] = f_fma(nalpha, t3[i], num[i]);
    }
  }
  {
    // + alpha^2 (z - 1) L1                                                     src/plonk.rs:356-369
    T zm[7], l1[4], t4[10];
#pragma unroll
    for (int i = 0; i < 7; i++) zm[i] = (i == 0) ? f_sub(z[0], f_const(1.f, tag)) : z[i];
#pragma unroll
    for (int i = 0; i < 4; i++) l1[i] = f_const(ck.L1(i), tag);
    fpoly_mul(zm, l1, t4);
#pragma unroll
    for (int i = 0; i < 22; i++) num[i] = f_red((i < 10) ? f_fma(a2, t4[i], num[i]) : num[i]);
  }

  // ---- divide by Z_H = x^4 - 1                                              src/plonk.rs:369-378
  T t[18];
#pragma unroll
  for (int j = 17; j >= 0; j--) t[j] = (j + 4 < 18) ? f_add(num[j + 4], t[j + 4]) : num[j + 4];   // unreduced, |.| <= 40
  bool rem_nz = false;
#pragma unroll
  for (int j = 0; j < 4; j++) rem_nz = rem_nz | !f_is_zero(f_red(f_add(num[j], t[j])));           // src/plonk.rs:370
  bool t_short = f_is_zero(t[17]);                                                                   // src/plonk.rs:376 (Q5)
  T tlo[6], tmid[6], thi[6];
#pragma unroll
  for (int i = 0; i < 6; i++) { tlo[i] = t[i]; tmid[i] = t[6 + i]; thi[i] = t[12 + i]; }
  P.e[6] = fcommit<ALGO, T>(thi, ck, Tb, false, oob_t);                             // src/plonk.rs:383-385
  P.e[5] = fcommit<ALGO, T>(tmid, ck, Tb, false, oob_t);
  P.e[4] = fcommit<ALGO, T>(tlo, ck, Tb, false, oob_t);
  const T zc = cs.zeta(point_of(P.e[4]), point_of(P.e[5]), point_of(P.e[6]));

  // ---- evaluations at z                                                     src/plonk.rs:393-399
  T zp[10];
  zp[0] = f_const(1.f, tag); zp[1] = zc;
#pragma unroll
  for (int i = 2; i < 10; i++) zp[i] = f_red(f_mul(zp[i - 1], zc));
  T a_z = a[0], b_z = b[0], c_z = c[0], tlo_z = tlo[0], tmid_z = tmid[0], thi_z = thi[0], zw_z = zw[0];
#pragma unroll
  for (int i = 1; i < 6; i++) {
    a_z = f_fma(a[i], zp[i], a_z); b_z = f_fma(b[i], zp[i], b_z); c_z = f_fma(c[i], zp[i], c_z);
    tlo_z = f_fma(tlo[i], zp[i], tlo_z); tmid_z = f_fma(tmid[i], zp[i], tmid_z); thi_z = f_fma(thi[i], zp[i], thi_z);
  }
#pragma unroll
  for (int i = 1; i < 7; i++) zw_z = f_fma(zw[i], zp[i], zw_z);
  T s1_z = f_const(ck.S(0, 0), tag), s2_z = f_const(ck.S(1, 0), tag), l1_z = f_const(ck.L1(0), tag);
#pragma unroll
  for (int i = 1; i < 4; i++) {
    s1_z = f_fma(f_const(ck.S(0, i), tag), zp[i], s1_z); s2_z = f_fma(f_const(ck.S(1, i), tag), zp[i], s2_z);
    l1_z = f_fma(f_const(ck.L1(i), tag), zp[i], l1_z);
  }
  a_z = f_red(a_z); b_z = f_red(b_z); c_z = f_red(c_z); s1_z = f_red(s1_z); s2_z = f_red(s2_z); zw_z = f_red(zw_z); l1_z = f_red(l1_z);
  const T z6 = zp[6];
  const T z12 = f_red(f_mul(z6, z6));
  T t_z = f_red(f_fma(z12, thi_z, f_fma(z6, tmid_z, tlo_z)));

  // ---- linearisation polynomial r (unreduced, 10 coefficients)             src/plonk.rs:401-422 (Q2)
  T r[10];
  {
    T bz = f_mul(beta, zc);
    T f1 = f_red(f_add(f_add(a_z, bz), gamma)), f2 = f_red(f_add(f_fma(bz, f_const(2.f, tag), b_z), gamma)),
      f3 = f_red(f_add(f_fma(bz, f_const(3.f, tag), c_z), gamma));
    T k2s = f_mul(f_red(f_mul(f_mul(f1, f2), f3)), alpha);
    T kz = f_red(f_fma(l1_z, a2, k2s));
    T g1v = f_red(f_add(f_fma(beta, s1_z, a_z), gamma)), g2v = f_red(f_add(f_fma(beta, s2_z, b_z), gamma));
    T k3s = f_red(f_mul(f_red(f_mul(f_mul(g1v, g2v), alpha)), f_red(f_mul(beta, zw_z))));
    T s3[4], zs3[10];
#pragma unroll
    for (int i = 0; i < 4; i++) s3[i] = f_const(ck.S(2, i), tag);
    fpoly_mul(z, s3, zs3);
    T ab_z = f_red(f_mul(a_z, b_z));
#pragma unroll
    for (int i = 0; i < 10; i++) {
      T s = f_mul(k3s, zs3[i]);
      if (i < 7) s = f_fma(kz, z[i], s);
      if (i < 4) {
        s = f_fma(ab_z, f_const(ck.QM(i), tag), s); s = f_fma(a_z, f_const(ck.QL(i), tag), s);
        s = f_fma(b_z, f_const(ck.QR(i), tag), s); s = f_fma(c_z, f_const(ck.QO(i), tag), s);
        s = f_add(s, f_const(ck.QC(i), tag));
      }
      r[i] = s;
    }
  }
  T r_z = r[0];
#pragma unroll
  for (int i = 1; i < 10; i++) r_z = f_fma(r[i], zp[i], r_z);
  r_z = f_red(r_z);
  P.ev[0] = f_canon(a_z); P.ev[1] = f_canon(b_z); P.ev[2] = f_canon(c_z); P.ev[3] = f_canon(s1_z); P.ev[4] = f_canon(s2_z);
  P.ev[5] = f_canon(r_z); P.ev[6] = f_canon(zw_z);
  const T v = cs.v(P.ev);

  // ---- opening polynomials                                                  src/plonk.rs:430-446
  T v2 = f_red(f_mul(v, v)), v3 = f_red(f_mul(v2, v)), v4 = f_red(f_mul(v3, v)), v5 = f_red(f_mul(v4, v)), v6 = f_red(f_mul(v5, v));
  T wn[10];
#pragma unroll
  for (int i = 0; i < 10; i++) {
    T s = f_mul(v, r[i]);
    if (i < 6) {
      s = f_add(s, tlo[i]); s = f_fma(z6, tmid[i], s); s = f_fma(z12, thi[i], s);
      s = f_fma(v2, a[i], s); s = f_fma(v3, b[i], s); s = f_fma(v4, c[i], s);
    }
    if (i < 4) { s = f_fma(v5, f_const(ck.S(0, i), tag), s); s = f_fma(v6, f_const(ck.S(1, i), tag), s); }
    wn[i] = s;
  }
  {
    T c0 = f_fma(v6, s2_z, f_fma(v5, s1_z, f_fma(v4, c_z, f_fma(v3, b_z, f_fma(v2, a_z, f_fma(v, r_z, t_z))))));
    wn[0] = f_sub(wn[0], c0);
  }
  // synthetic division by (x - z): remainder identically zero                 src/plonk.rs:437-438
  T wz[9];
  wz[8] = f_red(wn[9]);
#pragma unroll
  for (int k = 7; k >= 0; k--) wz[k] = f_red(f_fma(zc, wz[k + 1], wn[k + 1]));
  // (z_x - z_omega_z) / (x - z omega)                                         src/plonk.rs:441-442
  T wzw[6];
  {
    T zo = f_mul(zc, f_const(4.f, tag));
    wzw[5] = (ALGO == ALGO_ARITH) ? f_red(z[6]) : z[6];   // the ARITH tables are indexed by centred residues, z[6] = b7 arrives as 0..16
#pragma unroll
    for (int k = 4; k >= 0; k--) wzw[k] = f_red(f_fma(zo, wzw[k + 1], z[k + 1]));
  }
  P.e[7] = fcommit<ALGO, T>(wz, ck, Tb, true, oob_w);                               // src/plonk.rs:445-446 -> :56 (Q2)
  P.e[8] = fcommit<ALGO, T>(wzw, ck, Tb, true, oob_w);
  cs.u(point_of(P.e[7]), point_of(P.e[8]));

  uint32_t status = 0;
  if (oob_w) status = 5;
  if (oob_t) status = 5;
  if (t_short) status = 4;
  if (rem_nz) status = 3;
  if (oob_z) status = 5;
  if (div0) status = 2;
  if (oob_abc) status = 5;
  return status;
}

// the reference's interface: the caller supplies the Challange
template <int ALGO, class T, class CK>
PBH_HD uint32_t prove_core_f32(const T (&w)[12], const T (&rnd_in)[9], const T (&ch_in)[5], const CK& ck, const Tables& Tb,
                               const float* inv17c, ProofF& P) {
  FixedChal<T> cs(ch_in);
  return prove_core_f32_cs<ALGO, T>(w, rnd_in, cs, ck, Tb, inv17c, P);
}

// One proof through the FP32 core (every input class, Q1 included).  Inputs are canonical bytes (< 17).  Output as packed
// points + canonical evaluations, like prove_one<ALGO_TABLE>.
// `unsat_known`: -1 = evaluate constraints.satisfies here; 0 / 1 = already evaluated by the caller.
// PBH_CIRCUIT = true: the context's constants equal PbhCK's (checked by the host), use the compile-time instantiation.
template <int ALGO, bool PBH_CIRCUIT = false>
PBH_HD uint32_t prove_item_f32(const uint32_t (&w)[12], const uint32_t (&rnd)[9], const uint32_t (&ch)[5], const Consts& K,
                               const ConstsF& KF, const Tables& T, ProofRegs& P, int unsat_known = -1) {
  const bool unsat = unsat_known < 0 ? unsatisfied(w, K) : (unsat_known != 0);
  F32* tag = nullptr;
  F32 wf[12], rf[9], cf[5];
#pragma unroll
  for (int i = 0; i < 12; i++) wf[i] = f_from_u32(w[i], tag);
#pragma unroll
  for (int i = 0; i < 9; i++) rf[i] = f_from_u32(rnd[i], tag);
#pragma unroll
  for (int i = 0; i < 5; i++) cf[i] = f_from_u32(ch[i], tag);
  ProofF pf;
  uint32_t status;
  if (PBH_CIRCUIT) {
    status = prove_core_f32<ALGO, F32>(wf, rf, cf, PbhCK(), T, T.inv17c, pf);
  } else {
    status = prove_core_f32<ALGO, F32>(wf, rf, cf, RuntimeCK{KF, K.n_pts}, T, T.inv17c, pf);
  }
#pragma unroll
  for (int k = 0; k < 9; k++) P.pt[k] = (ALGO == ALGO_TABLE) ? T.pt17[pf.e[k]] : pf.e[k];
#pragma unroll
  for (int k = 0; k < 7; k++) P.ev[k] = pf.ev[k];
  // program order of the reference: the satisfiability assert comes first; an SRS too short for a, b, c (only
  // with fewer than 6 SRS points) precedes the accumulator
  if (unsat) status = 1;
  return status;
}

// ---- Fiat-Shamir prover item: the challenges come from the SHA-256 transcript seeded with the
// context's `seed`.  `derived` receives alpha beta gamma z v u (those derived before a failing site; the caller zeroes
// them with the proof when status != 0).  FP32 = false runs the int32 routine for every item.
struct ConvF32 { PBH_HD F32 operator()(uint32_t x) const { return f_from_u32(x, (F32*)nullptr); } };

template <int ALGO, bool FP32, bool PBH_CIRCUIT>
PBH_HD uint32_t prove_item_fs(const uint32_t (&w)[12], const uint32_t (&rnd)[9], const uint32_t (&seed)[8], const Consts& K,
                              const ConstsF& KF, const Tables& T, ProofRegs& P, uint32_t (&derived)[6], bool want_u = true) {
  if (!FP32) {
    FsChal<uint32_t, ConvU32> cs(seed, want_u);
    const uint32_t status = prove_one_cs<ALGO, false>(w, rnd, cs, K, T, P, -1);
#pragma unroll
    for (int k = 0; k < 6; k++) derived[k] = cs.derived[k];
    return status;
  }
  const bool unsat = unsatisfied(w, K);
  F32* tag = nullptr;
  F32 wf[12], rf[9];
#pragma unroll
  for (int i = 0; i < 12; i++) wf[i] = f_from_u32(w[i], tag);
#pragma unroll
  for (int i = 0; i < 9; i++) rf[i] = f_from_u32(rnd[i], tag);
  // the last two steps (v, u), where few polynomials are still live, call the shared copy of the compression: measured
  // 490 -> 441 us per 2^20 complete transcripts; outlining the first three as well loses it again to register shuffling
  FsChal<F32, ConvF32, 24> cs(seed, want_u);
  ProofF pf;
  uint32_t status;
  if (PBH_CIRCUIT) status = prove_core_f32_cs<ALGO, F32>(wf, rf, cs, PbhCK(), T, T.inv17c, pf);
  else status = prove_core_f32_cs<ALGO, F32>(wf, rf, cs, RuntimeCK{KF, K.n_pts}, T, T.inv17c, pf);
#pragma unroll
  for (int k = 0; k < 9; k++) P.pt[k] = (ALGO == ALGO_TABLE) ? T.pt17[pf.e[k]] : pf.e[k];
#pragma unroll
  for (int k = 0; k < 7; k++) P.ev[k] = pf.ev[k];
#pragma unroll
  for (int k = 0; k < 6; k++) derived[k] = cs.derived[k];
  if (unsat) status = 1;
  return status;
}

}  // namespace pbh
