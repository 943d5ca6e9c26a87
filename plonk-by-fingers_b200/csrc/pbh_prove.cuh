// pbh_prove.cuh — one Plonk-by-hand proof per thread, all state in registers.
//
// Reproduces Plonk::prove (src/plonk.rs:191-466) and Constrains::satisfies (src/constraints.rs:198-230)
// bit for bit, including the quirks of SURVEY.md §2.3.  The reference works on heap-allocated,
// normalised, variable-length polynomials; here every polynomial is a fixed-width zero-padded register
// array and the reference's length-dependent behaviour is recovered exactly where it is observable:
//   - Q1  (src/poly.rs:192-203): `t1+t2-t3` adds the tail of t3 instead of subtracting it when t3 is longer
//         than t1+t2.  Only possible when coefficient 21 of t1+t2 is zero, handled in a rare branch.
//   - Q5  (src/plonk.rs:376): the slice [12..18] panics unless t has exactly 18 coefficients.
//   - Q2  (src/plonk.rs:56 via :445): eval_at_s panics when w_z has more coefficients than SRS points.
//   - Q15 (src/poly.rs:220-228) and normalisation only change lengths, never values.
// The reference's panics become the status byte of include/pbh_b200.h, first failing site in program order.
#pragma once
#include "pbh_arith.cuh"
#include "pbh_fs.cuh"

namespace pbh {

// Per-context constants, passed BY VALUE as a kernel parameter so that uniform accesses are constant-bank
// operands.  Everything here is circuit- or SRS-constant work that the reference redoes on every call
// (src/plonk.rs:222-243, 328-333, 506-517, 557-562).
struct Consts {
  // selector vectors as given (for satisfies) and the copy-constraint permutation: witness value k (a0..3,
  // b0..3, c0..3) must equal witness value perm[k]                         src/constraints.rs:198-230
  uint32_t q_l[4], q_r[4], q_o[4], q_m[4], q_c[4];
  uint32_t perm[12];
  // interpolate_at_h of the selector vectors, zero padded                 src/plonk.rs:236-240
  uint32_t QL[4], QR[4], QO[4], QM[4], QC[4];
  // copy_constraints_to_roots (values) and their interpolations            src/plonk.rs:222-224, 241-243
  uint32_t sig[3][4];
  uint32_t S[3][4];
  uint32_t L1[4];          // interpolate_at_h([1,0,0,0])                     src/plonk.rs:328-333
  // SRS: discrete log (base G, mod 17) of g1s[i]; n_pts = g1s.len()         src/plonk.rs:35-48 (Q11)
  uint32_t srs_dlog[10];
  uint32_t n_pts;
  // verifier preprocessing: dlog (mod 17) of q_m_s q_l_s q_r_s q_o_s q_c_s sigma_1_s sigma_2_s sigma_3_s
  uint32_t vdlog[8];       //                                                 src/plonk.rs:510-517
  uint32_t g2_1[2], g2_s[2];
};

enum { ALGO_ARITH = 0, ALGO_TABLE = 1 };

// schoolbook product with one reduction per output coefficient              src/poly.rs:205-218
template <int LA, int LB>
PBH_HD void poly_mul17(const uint32_t (&a)[LA], const uint32_t (&b)[LB], uint32_t (&out)[LA + LB - 1]) {
#pragma unroll
  for (int k = 0; k < LA + LB - 1; k++) {
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < LA; i++) {
      if (k - i >= 0 && k - i < LB) acc += a[i] * b[k - i];
    }
    out[k] = mod17(acc);
  }
}
// unreduced accumulate: acc[k] += sum a[i] b[k-i]
template <int LA, int LB, int LO>
PBH_HD void poly_mac17(const uint32_t (&a)[LA], const uint32_t (&b)[LB], uint32_t (&acc)[LO]) {
#pragma unroll
  for (int k = 0; k < LA + LB - 1; k++) {
#pragma unroll
    for (int i = 0; i < LA; i++) {
      if (k - i >= 0 && k - i < LB) acc[k] += a[i] * b[k - i];
    }
  }
}

// Plonk::interpolate_at_h for H = {1,4,16,13}: h_pows_inv = 13 * [w^(-ij)], w^-1 = 13
// (src/plonk.rs:153-160, 177-179; matrix of SURVEY.md §3.1).  Output zero padded to 4 coefficients.
PBH_HD void intt4(uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3, uint32_t (&f)[4]) {
  f[0] = mod17(13u * (v0 + v1 + v2 + v3));
  f[1] = mod17(13u * v0 + 16u * v1 + 4u * v2 + v3);
  f[2] = mod17(13u * v0 + 4u * v1 + 13u * v2 + 4u * v3);
  f[3] = mod17(13u * v0 + v1 + 4u * v2 + 16u * v3);
}
// forward size-4 NTT, omega = 4: evals[i] = sum_j c[j] 4^(ij)                src/fft.rs:90-106
PBH_HD void ntt4(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t (&e)[4]) {
  e[0] = mod17(c0 + c1 + c2 + c3);
  e[1] = mod17(c0 + 4u * c1 + 16u * c2 + 13u * c3);
  e[2] = mod17(c0 + 16u * c1 + c2 + 16u * c3);
  e[3] = mod17(c0 + 13u * c1 + 16u * c2 + 4u * c3);
}

// index of the canonical coefficients 3J, 3J+1, 3J+2 (those below L) in a three-point SRS table, whose digits are the
// centred residues plus 8 (FixedBaseTables)
PBH_HD uint32_t tri_digit(uint32_t c) { return c <= 8u ? c + 8u : c - 9u; }
template <int L, int J>
PBH_HD uint32_t tri_index(const uint32_t (&c)[L]) {
  uint32_t idx = (3 * J < L) ? tri_digit(c[(3 * J < L) ? 3 * J : 0]) : 8u;
  idx += 17u * ((3 * J + 1 < L) ? tri_digit(c[(3 * J + 1 < L) ? 3 * J + 1 : 0]) : 8u);
  idx += 289u * ((3 * J + 2 < L) ? tri_digit(c[(3 * J + 2 < L) ? 3 * J + 2 : 0]) : 8u);
  return idx;
}

// SRS::eval_at_s (src/plonk.rs:51-58) of a zero-padded coefficient array, as a packed point x | y<<8 | inf<<16.
//   TABLE: every g1s[i] is [d_i]G, so the commitment is [sum c_i d_i mod 17]G — one conflict-free lookup.
//   ARITH: per-term fixed-base multiples [c_i]g1s[i] added with the affine group law.
template <int ALGO, int L>
PBH_HD uint32_t commit(const uint32_t (&c)[L], const Consts& K, const Tables& T) {
  if (ALGO == ALGO_TABLE) {
    uint32_t e = 0;
#pragma unroll
    for (int i = 0; i < L; i++) e += c[i] * K.srs_dlog[i];
    return T.pt17[mod17(e)];
  } else {
    // three coefficients per lookup (FixedBaseTables), the first lookup needs no addition
    static_assert(L >= 2 && L <= 10, "the SRS tables cover 10 points");
    G1 acc = g1_unpack(pair_lookup(T.fixed->srs_tri[0], tri_index<L, 0>(c)));
#pragma unroll
    for (int j = 1; j < (L + 2) / 3; j++) {
      const uint32_t idx = j == 1 ? tri_index<L, 1>(c) : (j == 2 ? tri_index<L, 2>(c) : tri_index<L, 3>(c));
      acc = g1_add(acc, g1_unpack(pair_lookup(T.fixed->srs_tri[j], idx)), T.inv101);
    }
    return g1_pack(acc);
  }
}

// true when the normalised polynomial has more than n_pts coefficients (the reference then indexes g1s out of
// bounds: src/plonk.rs:56)
template <int L>
PBH_HD bool longer_than(const uint32_t (&c)[L], uint32_t n_pts) {
  bool r = false;
#pragma unroll
  for (int j = 0; j < L; j++) r = r || ((uint32_t)j >= n_pts && c[j] != 0u);
  return r;
}

struct ProofRegs {
  uint32_t pt[9];   // packed points a_s b_s c_s z_s t_lo_s t_mid_s t_hi_s w_z_s w_z_omega_s
  uint32_t ev[7];   // a_z b_z c_z s_sigma_1_z s_sigma_2_z r_z z_omega_z
};

// gate part of constraints.satisfies, src/constraints.rs:200-210 (Q8: q_l multiplies b as well)
PBH_HD bool gates_unsatisfied(const uint32_t (&w)[12], const Consts& K) {
  bool unsat = false;
#pragma unroll
  for (int n = 0; n < 4; n++) {
    uint32_t r = K.q_l[n] * (w[n] + w[4 + n]) + K.q_o[n] * w[8 + n] + K.q_m[n] * mod17(w[n] * w[4 + n]) + K.q_c[n];
    unsat = unsat | (mod17(r) != 0u);
  }
  return unsat;
}

// constraints.satisfies(assigments), src/constraints.rs:198-230
PBH_HD bool unsatisfied(const uint32_t (&w)[12], const Consts& K) {
  bool unsat = gates_unsatisfied(w, K);
  // witness value k must equal witness value perm[k]; values are 5-bit fields of a 64-bit word so that the
  // (uniform, runtime) permutation needs no local-memory indexing
  unsigned long long packed = 0;
#pragma unroll
  for (int k = 0; k < 12; k++) packed |= (unsigned long long)w[k] << (5 * k);
#pragma unroll
  for (int k = 0; k < 12; k++) {
    uint32_t other = (uint32_t)(packed >> (5u * K.perm[k])) & 31u;
    unsat = unsat || (other != w[k]);
  }
  return unsat;
}

// w[12] = a[0..4) b[0..4) c[0..4); r[9] = b1..b9; ch[5] = alpha beta gamma z v.  All inputs < 17.
// Returns the status byte.  `P` is fully defined only for status 0.  With UPTO_T the routine stops after the
// quotient (sites :199 .. :376) — enough for items whose t(x) is known to be short (status 1-4 only).
// `unsat_known`: -1 = evaluate constraints.satisfies here; 0 / 1 = the caller already did (e.g. from shared memory).
// The challenges come from `cs` (pbh_fs.cuh) at the point where the reference first uses each.
template <int ALGO, bool UPTO_T, class CS>
PBH_HD uint32_t prove_one_cs(const uint32_t (&w)[12], const uint32_t (&rnd)[9], CS& cs, const Consts& K, const Tables& T,
                             ProofRegs& P, int unsat_known = -1) {
  const uint32_t n_pts = K.n_pts;

  const bool unsat = unsat_known < 0 ? unsatisfied(w, K) : (unsat_known != 0);

  // ---- wire polynomials                                                     src/plonk.rs:233-235, 248-252
  uint32_t fa[4], fb[4], fc[4];
  intt4(w[0], w[1], w[2], w[3], fa);
  intt4(w[4], w[5], w[6], w[7], fb);
  intt4(w[8], w[9], w[10], w[11], fc);
  // (b1 x + b2)(x^4 - 1) + f_a = [f0 - b2, f1 - b1, f2, f3, b2, b1]
  uint32_t a[6], b[6], c[6];
  a[0] = sub17(fa[0], rnd[1]); a[1] = sub17(fa[1], rnd[0]); a[2] = fa[2]; a[3] = fa[3]; a[4] = rnd[1]; a[5] = rnd[0];
  b[0] = sub17(fb[0], rnd[3]); b[1] = sub17(fb[1], rnd[2]); b[2] = fb[2]; b[3] = fb[3]; b[4] = rnd[3]; b[5] = rnd[2];
  c[0] = sub17(fc[0], rnd[5]); c[1] = sub17(fc[1], rnd[4]); c[2] = fc[2]; c[3] = fc[3]; c[4] = rnd[5]; c[5] = rnd[4];

  bool oob_abc = longer_than(a, n_pts) || longer_than(b, n_pts) || longer_than(c, n_pts);
  P.pt[0] = commit<ALGO>(a, K, T);                                           // src/plonk.rs:255-257
  P.pt[1] = commit<ALGO>(b, K, T);
  P.pt[2] = commit<ALGO>(c, K, T);
  uint32_t beta, gamma;
  cs.beta_gamma(P.pt[0], P.pt[1], P.pt[2], beta, gamma);

  // ---- accumulator                                                          src/plonk.rs:278-299
  const uint32_t bg = gamma;
  uint32_t acc[4];
  acc[0] = 1u;
  bool div0 = false;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const uint32_t om = (i == 0) ? 1u : (i == 1 ? 4u : 16u);   // omega^i, omega = 4 (src/pbh/mod.rs:29)
    uint32_t wa = w[i] + bg, wb = w[4 + i] + bg, wc = w[8 + i] + bg;
    uint32_t dend = mul17(mul17(mod17(wa + beta * om), mod17(wb + beta * (2u * om))), mod17(wc + beta * (3u * om)));
    uint32_t dsor = mul17(mul17(mod17(wa + beta * K.sig[0][i]), mod17(wb + beta * K.sig[1][i])),
                          mod17(wc + beta * K.sig[2][i]));
    div0 = div0 || (dsor == 0u);                                             // src/plonk.rs:297 unwrap
    acc[i + 1] = mul17(acc[i], mul17(dend, T.inv17[dsor]));
  }
  uint32_t accx[4];
  intt4(acc[0], acc[1], acc[2], acc[3], accx);                              // src/plonk.rs:303 (Q7 never fires)
  // (b7 x^2 + b8 x + b9)(x^4 - 1) + acc_x                                    src/plonk.rs:309
  uint32_t z[7];
  z[0] = sub17(accx[0], rnd[8]); z[1] = sub17(accx[1], rnd[7]); z[2] = sub17(accx[2], rnd[6]); z[3] = accx[3];
  z[4] = rnd[8]; z[5] = rnd[7]; z[6] = rnd[6];
  bool oob_z = longer_than(z, n_pts);
  P.pt[3] = commit<ALGO>(z, K, T);                                           // src/plonk.rs:313
  const uint32_t alpha = cs.alpha(P.pt[3]);

  // ---- quotient numerator                                                   src/plonk.rs:339-369
  uint32_t t12[22];   // t1 + t2
  uint32_t num[22];
  {
    // t2 = alpha (a + gamma + beta x)(b + gamma + beta k1 x)(c + gamma + beta k2 x) z
    uint32_t A[6], B[6], C[6];
#pragma unroll
    for (int i = 0; i < 6; i++) { A[i] = a[i]; B[i] = b[i]; C[i] = c[i]; }
    A[0] = add17(A[0], gamma); A[1] = mod17(A[1] + beta);
    B[0] = add17(B[0], gamma); B[1] = mod17(B[1] + 2u * beta);               // K1 = 2  src/pbh/mod.rs:27
    C[0] = add17(C[0], gamma); C[1] = mod17(C[1] + 3u * beta);               // K2 = 3  src/pbh/mod.rs:28
#pragma unroll
    for (int i = 0; i < 6; i++) A[i] = mul17(A[i], alpha);
    uint32_t AB[11], ABC[16], t2[22];
    poly_mul17(A, B, AB);
    poly_mul17(AB, C, ABC);
    poly_mul17(ABC, z, t2);
    // t1 = a b q_m + a q_l + b q_r + c q_o + q_c   (PI = 0)
    uint32_t ab[11];
    poly_mul17(a, b, ab);
    uint32_t t1[14];
#pragma unroll
    for (int i = 0; i < 14; i++) t1[i] = 0;
    poly_mac17(ab, K.QM, t1);
    poly_mac17(a, K.QL, t1);
    poly_mac17(b, K.QR, t1);
    poly_mac17(c, K.QO, t1);
#pragma unroll
    for (int i = 0; i < 4; i++) t1[i] += K.QC[i];
#pragma unroll
    for (int i = 0; i < 22; i++) t12[i] = (i < 14) ? mod17(t1[i] + t2[i]) : t2[i];
  }
  {
    // t3 = alpha (a + beta S1 + gamma)(b + beta S2 + gamma)(c + beta S3 + gamma) z(omega x)
    uint32_t A[6], B[6], C[6], zw[7];
#pragma unroll
    for (int i = 0; i < 6; i++) {
      A[i] = (i < 4) ? mod17(a[i] + beta * K.S[0][i]) : a[i];
      B[i] = (i < 4) ? mod17(b[i] + beta * K.S[1][i]) : b[i];
      C[i] = (i < 4) ? mod17(c[i] + beta * K.S[2][i]) : c[i];
    }
    A[0] = add17(A[0], gamma); B[0] = add17(B[0], gamma); C[0] = add17(C[0], gamma);
#pragma unroll
    for (int i = 0; i < 6; i++) A[i] = mul17(A[i], alpha);
    // z_omega_x: coefficient i times omega^i, omega^i = 1,4,16,13,1,4,16     src/plonk.rs:346-352
    zw[0] = z[0]; zw[1] = mod17(4u * z[1]); zw[2] = mod17(16u * z[2]); zw[3] = mod17(13u * z[3]);
    zw[4] = z[4]; zw[5] = mod17(4u * z[5]); zw[6] = mod17(16u * z[6]);
    uint32_t AB[11], ABC[16], t3[22];
    poly_mul17(A, B, AB);
    poly_mul17(AB, C, ABC);
    poly_mul17(ABC, zw, t3);
    // (t1 + t2) - t3 with the reference's SubAssign (Q1)
#pragma unroll
    for (int i = 0; i < 22; i++) num[i] = sub17(t12[i], t3[i]);
    if (t12[21] == 0u) {
      // rare: t1+t2 is shorter than 22 coefficients, so t3 may be longer than it; coefficients of t3 at or
      // beyond len(t1+t2) are pushed un-negated.  len >= 1 (the zero polynomial is [0]).
      bool tail_zero = true;   // t12[n..21] all zero
#pragma unroll
      for (int n = 21; n >= 1; n--) {
        tail_zero = tail_zero && (t12[n] == 0u);
        if (tail_zero) num[n] = t3[n];
      }
    }
  }
  {
    // t4 = alpha^2 (z - 1) L1                                                src/plonk.rs:356
    uint32_t zm[7];
    uint32_t a2 = mul17(alpha, alpha);
#pragma unroll
    for (int i = 0; i < 7; i++) zm[i] = mul17((i == 0) ? sub17(z[0], 1u) : z[i], a2);
    uint32_t t4[10];
    poly_mul17(zm, K.L1, t4);
#pragma unroll
    for (int i = 0; i < 10; i++) num[i] = add17(num[i], t4[i]);
  }

  // ---- divide by Z_H = x^4 - 1                                              src/plonk.rs:369-378, src/poly.rs:230-247
  uint32_t t[18];
#pragma unroll
  for (int j = 17; j >= 0; j--) t[j] = (j + 4 < 18) ? add17(num[j + 4], t[j + 4]) : num[j + 4];
  bool rem_nz = false;
#pragma unroll
  for (int j = 0; j < 4; j++) rem_nz = rem_nz || (add17(num[j], t[j]) != 0u);  // src/plonk.rs:370
  bool t_short = (t[17] == 0u);                                                 // src/plonk.rs:376 (Q5)

  if (UPTO_T) {
    uint32_t st = t_short ? 4u : 0u;
    if (rem_nz) st = 3;
    if (oob_z) st = 5;
    if (div0) st = 2;
    if (oob_abc) st = 5;
    if (unsat) st = 1;
    return st;
  }
  uint32_t tlo[6], tmid[6], thi[6];
#pragma unroll
  for (int i = 0; i < 6; i++) { tlo[i] = t[i]; tmid[i] = t[6 + i]; thi[i] = t[12 + i]; }
  bool oob_t = longer_than(thi, n_pts) || longer_than(tmid, n_pts) || longer_than(tlo, n_pts);
  P.pt[4] = commit<ALGO>(tlo, K, T);                                         // src/plonk.rs:383-385
  P.pt[5] = commit<ALGO>(tmid, K, T);
  P.pt[6] = commit<ALGO>(thi, K, T);
  const uint32_t zc = cs.zeta(P.pt[4], P.pt[5], P.pt[6]);

  // ---- evaluations at z                                                     src/plonk.rs:393-399
  uint32_t zp[10];   // powers of the challenge z
  zp[0] = 1u;
#pragma unroll
  for (int i = 1; i < 10; i++) zp[i] = mul17(zp[i - 1], zc);
  uint32_t a_z = 0, b_z = 0, c_z = 0, s1_z = 0, s2_z = 0, tlo_z = 0, tmid_z = 0, thi_z = 0, zw_z = 0;
#pragma unroll
  for (int i = 0; i < 6; i++) {
    a_z += a[i] * zp[i]; b_z += b[i] * zp[i]; c_z += c[i] * zp[i];
    tlo_z += tlo[i] * zp[i]; tmid_z += tmid[i] * zp[i]; thi_z += thi[i] * zp[i];
  }
#pragma unroll
  for (int i = 0; i < 4; i++) { s1_z += K.S[0][i] * zp[i]; s2_z += K.S[1][i] * zp[i]; }
  {
    const uint32_t omp[7] = {1u, 4u, 16u, 13u, 1u, 4u, 16u};
#pragma unroll
    for (int i = 0; i < 7; i++) zw_z += mod17(z[i] * omp[i]) * zp[i];
  }
  a_z = mod17(a_z); b_z = mod17(b_z); c_z = mod17(c_z); s1_z = mod17(s1_z); s2_z = mod17(s2_z); zw_z = mod17(zw_z);
  const uint32_t z6 = zp[6];                    // z^(n+2)
  const uint32_t z12 = mul17(z6, z6);           // z^(2n+4)
  uint32_t t_z = mod17(mod17(tlo_z) + z6 * mod17(tmid_z) + z12 * mod17(thi_z));
  uint32_t l1_z = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) l1_z += K.L1[i] * zp[i];
  l1_z = mod17(l1_z);

  // ---- linearisation polynomial r                                           src/plonk.rs:401-422 (Q2)
  uint32_t r[10];
  {
    uint32_t ab_z = mul17(a_z, b_z);
    uint32_t bz = mul17(beta, zc);
    uint32_t k2s = mul17(mul17(mul17(mod17(a_z + bz + gamma), mod17(b_z + 2u * bz + gamma)), mod17(c_z + 3u * bz + gamma)),
                         alpha);
    uint32_t k4s = mul17(l1_z, mul17(alpha, alpha));
    uint32_t kz = add17(k2s, k4s);
    uint32_t k3s = mul17(mul17(mul17(mod17(a_z + beta * s1_z + gamma), mod17(b_z + beta * s2_z + gamma)), alpha),
                         mul17(beta, zw_z));
    uint32_t zs3[10];
    poly_mul17(z, K.S[2], zs3);
#pragma unroll
    for (int i = 0; i < 10; i++) {
      uint32_t acc_r = k3s * zs3[i];
      if (i < 7) acc_r += kz * z[i];
      if (i < 4) acc_r += ab_z * K.QM[i] + a_z * K.QL[i] + b_z * K.QR[i] + c_z * K.QO[i] + K.QC[i];
      r[i] = mod17(acc_r);
    }
  }
  uint32_t r_z = 0;
#pragma unroll
  for (int i = 0; i < 10; i++) r_z += r[i] * zp[i];
  r_z = mod17(r_z);
  P.ev[0] = a_z; P.ev[1] = b_z; P.ev[2] = c_z; P.ev[3] = s1_z; P.ev[4] = s2_z; P.ev[5] = r_z; P.ev[6] = zw_z;
  const uint32_t v = cs.v(P.ev);

  // ---- opening polynomials                                                  src/plonk.rs:430-446
  uint32_t v2 = mul17(v, v), v3 = mul17(v2, v), v4 = mul17(v3, v), v5 = mul17(v4, v), v6 = mul17(v5, v);
  uint32_t wn[10];
#pragma unroll
  for (int i = 0; i < 10; i++) {
    uint32_t s = v * r[i];
    if (i < 6) s += tlo[i] + z6 * tmid[i] + z12 * thi[i] + v2 * a[i] + v3 * b[i] + v4 * c[i];
    if (i < 4) s += v5 * K.S[0][i] + v6 * K.S[1][i];
    wn[i] = s;
  }
  {
    uint32_t c0 = t_z + v * r_z + v2 * a_z + v3 * b_z + v4 * c_z + v5 * s1_z + v6 * s2_z;
    wn[0] += 17u * 128u - c0;   // c0 < 7 * 256
  }
#pragma unroll
  for (int i = 0; i < 10; i++) wn[i] = mod17(wn[i]);
  // divide by (x - z): synthetic division, remainder is identically zero    src/plonk.rs:437-438
  uint32_t wz[9];
  wz[8] = wn[9];
#pragma unroll
  for (int k = 7; k >= 0; k--) wz[k] = mod17(wn[k + 1] + zc * wz[k + 1]);
  // (z_x - z_omega_z) / (x - z omega)                                        src/plonk.rs:441-442
  uint32_t wzw[6];
  {
    uint32_t zo = mod17(4u * zc);
    wzw[5] = z[6];
#pragma unroll
    for (int k = 4; k >= 0; k--) wzw[k] = mod17(z[k + 1] + zo * wzw[k + 1]);
  }
  bool oob_w = longer_than(wz, n_pts) || longer_than(wzw, n_pts);            // src/plonk.rs:445-446 -> :56 (Q2)
  // table slots at or beyond n_pts are only reachable together with oob_w
  P.pt[7] = commit<ALGO>(wz, K, T);
  P.pt[8] = commit<ALGO>(wzw, K, T);
  cs.u(P.pt[7], P.pt[8]);

  // first failing site in program order
  uint32_t status = 0;
  if (oob_w) status = 5;
  if (oob_t) status = 5;
  if (t_short) status = 4;
  if (rem_nz) status = 3;
  if (oob_z) status = 5;
  if (div0) status = 2;
  if (oob_abc) status = 5;
  if (unsat) status = 1;
  return status;
}

// the reference's interface: the caller supplies the Challange
template <int ALGO, bool UPTO_T = false>
PBH_HD uint32_t prove_one(const uint32_t (&w)[12], const uint32_t (&rnd)[9], const uint32_t (&ch)[5], const Consts& K,
                          const Tables& T, ProofRegs& P, int unsat_known = -1) {
  FixedChal<uint32_t> cs(ch);
  return prove_one_cs<ALGO, UPTO_T>(w, rnd, cs, K, T, P, unsat_known);
}

}  // namespace pbh
