// pbh_verify.cuh — one Plonk-by-hand verification per thread.
//
// Reproduces Plonk::verify (src/plonk.rs:468-650) bit for bit: the on-curve test that ignores the infinity
// flag (Q9), the missing alpha in t_z (Q3), the panic on Z_H(z) = 0 (Q4), the identity pairing value 0+0u (Q10)
// and acceptance of any of the 101 affine curve points, not only the order-17 subgroup (Q17).  The eight
// selector / permutation commitments the reference recomputes on every call (src/plonk.rs:506-517) are context
// constants.
//
// Group work, two bit-exact algorithms:
//   ARITH  real affine arithmetic per item: one Straus (shared-doubling) multi-scalar multiplication over the
//          nine proof points with 5-bit scalars, fixed-base multiples for the constant commitments, two Miller
//          loops over r = 17 and two final exponentiations in F_101^2.
//   TABLE  E(F_101) is cyclic of order 102, so a point is its discrete log to a fixed generator; G1 arithmetic
//          becomes arithmetic mod 102 and the two pairings against the fixed G2 points become 102-entry
//          tables.  All tables fit in <= 32 shared-memory banks, so divergent lookups are conflict free.
#pragma once
#include "pbh_prove_f32.cuh"

namespace pbh {

enum : uint32_t {
  VR_ACCEPT = 0x01, VR_REJECT_PAIRING = 0x00, VR_NOT_ON_CURVE = 0x02, VR_NOT_IN_FIELD = 0x04, VR_PANIC_ZH0 = 0x10,
  VR_BAD_ENCODING = 0x20
};

// px, py: coordinates of a_s b_s c_s z_s t_lo_s t_mid_s t_hi_s w_z_s w_z_omega_s; infbits: bit k = point k's
// `infinite` flag; ev: a_z b_z c_z s_sigma_1_z s_sigma_2_z r_z z_omega_z; ch: alpha beta gamma z v; u = rand[0].
// All raw bytes.  Returns the PBH_VR_* result byte; e1/e2 are defined when the pairing check was reached.
// KF: non-null selects the FP32-pipe scalar path for PBH_ALGO_TABLE (verify_scalars_f32 below); null = int32 arithmetic.
template <class TT> PBH_HD void verify_scalars_f32(const TT (&idx)[9], const TT (&ev)[7], const TT (&ch)[5], TT u, const ConstsF& VF,
                                                   const float* inv17c, uint32_t& i1, uint32_t& i2, bool& zh0);
template <int ALGO>
PBH_HD uint32_t verify_one(const uint32_t (&px_in)[9], const uint32_t (&py_in)[9], uint32_t infbits, const uint32_t (&ev)[7],
                           const uint32_t (&ch_in)[5], uint32_t u, const Consts& K, const Tables& T, GT& e1, GT& e2,
                           const ConstsF* KF = nullptr) {
  e1.a = e1.b = e2.a = e2.b = 0;
  // ---- encoding (see include/pbh_b200.h).  Out-of-range bytes are replaced by 0 so that no table is indexed out
  // of bounds; the verdict of such an item is PBH_VR_BAD_ENCODING whatever is computed below.
  // The range checks are two maxima (three-input integer max on the device) instead of 24 compares.  PBH_ALGO_TABLE only
  // indexes its 128-entry tables with min(x, 127) and merely compares y, so it reads the coordinates as they are; the
  // curve arithmetic of PBH_ALGO_ARITH needs them in range.
  uint32_t max_xy = 0u, max_s = u;
#pragma unroll
  for (int k = 0; k < 9; k++) { max_xy = px_in[k] > max_xy ? px_in[k] : max_xy; max_xy = py_in[k] > max_xy ? py_in[k] : max_xy; }
#pragma unroll
  for (int k = 0; k < 5; k++) max_s = ch_in[k] > max_s ? ch_in[k] : max_s;
  const bool bad = (infbits >> 9) != 0u || max_xy >= 101u || max_s >= 17u;
  uint32_t px[9], py[9], ch[5];
#pragma unroll
  for (int k = 0; k < 9; k++) {
    px[k] = (ALGO == ALGO_TABLE) ? px_in[k] : (bad ? 0u : px_in[k]);
    py[k] = (ALGO == ALGO_TABLE) ? py_in[k] : (bad ? 0u : py_in[k]);
  }
#pragma unroll
  for (int k = 0; k < 5; k++) ch[k] = bad ? 0u : ch_in[k];
  if (bad) u = 0u;

  // ---- Step 1: in_curve on the raw coordinates (flag ignored)              src/plonk.rs:523-534
  bool off_curve = false;
  uint32_t idx[9];   // TABLE: discrete logs
  G1F<F32> pt[9];    // ARITH: points (pbh_g1f.cuh)
  F32* ftag = nullptr;
#pragma unroll
  for (int k = 0; k < 9; k++) {
    bool inf = (infbits >> k) & 1u;
    if (ALGO == ALGO_TABLE) {
      // the table holds the root y <= 50 of x^3 + 3 (0xFF when there is none) and its point's index; the other root is
      // 101 - y with index 102 - i.  Fold y into 0..50 first: one compare then decides "on the curve".
      uint32_t xi = px[k] < 127u ? px[k] : 127u;
      uint32_t ys = T.y_of_x[xi], i0 = T.idx_of_x[xi];
      const bool upper = py[k] > 50u;
      const uint32_t yf = upper ? 101u - py[k] : py[k];          // out-of-range y gives a value that matches no root
      off_curve = off_curve || yf != ys;
      uint32_t i = upper ? 102u - i0 : i0;
      idx[k] = inf ? 0u : i;                 // a flagged point acts as the identity (g1.rs:121, 148-150)
    } else {
      const F32 fx = f_from_u32(px[k], ftag), fy = f_from_u32(py[k], ftag);
      off_curve = off_curve || !g1f_in_curve(fx, fy);
      pt[k].x = inf ? f_const(0.f, ftag) : fx; pt[k].y = inf ? f_const(0.f, ftag) : fy; pt[k].inf = inf;
    }
  }
  // ---- Step 2: in_field                                                     src/plonk.rs:538-547
  bool off_field = false;
#pragma unroll
  for (int k = 0; k < 7; k++) off_field = off_field || ev[k] >= 17u;

  if (ALGO == ALGO_TABLE && KF != nullptr) {
    F32* tag = nullptr;
    F32 fidx[9], fev[7], fch[5];
#pragma unroll
    for (int k = 0; k < 9; k++) fidx[k] = f_from_u32(idx[k], tag);
#pragma unroll
    for (int k = 0; k < 7; k++) fev[k] = f_red(f_from_u32(ev[k], tag));      // also folds bytes >= 17 into range
#pragma unroll
    for (int k = 0; k < 5; k++) fch[k] = f_from_u32(ch[k], tag);
    uint32_t i1, i2;
    bool zh0f;
    verify_scalars_f32<F32>(fidx, fev, fch, f_from_u32(u, tag), *KF, T.inv17c, i1, i2, zh0f);
    e1.a = T.pair_s_a[i1]; e1.b = T.pair_s_b[i1];
    e2.a = T.pair_1_a[i2]; e2.b = T.pair_1_b[i2];
    uint32_t resf = (e1.a == e2.a && e1.b == e2.b) ? VR_ACCEPT : VR_REJECT_PAIRING;
    if (zh0f) resf = VR_PANIC_ZH0;
    if (off_field) resf = VR_NOT_IN_FIELD;
    if (off_curve) resf = VR_NOT_ON_CURVE;
    if (bad) resf = VR_BAD_ENCODING;
    if (resf != VR_ACCEPT && resf != VR_REJECT_PAIRING) e1.a = e1.b = e2.a = e2.b = 0;
    return resf;
  }
  const uint32_t alpha = ch[0], beta = ch[1], gamma = ch[2], z = ch[3], v = ch[4];
  // evaluations >= 17 are representable inputs (verdict false at Step 2); reduce them so that everything computed
  // below stays inside the table ranges
  const uint32_t a_z = mod17(ev[0]), b_z = mod17(ev[1]), c_z = mod17(ev[2]), s1_z = mod17(ev[3]), s2_z = mod17(ev[4]),
                 r_z = mod17(ev[5]), zw_z = mod17(ev[6]);

  // ---- Steps 4-7                                                            src/plonk.rs:553-579
  // Lazy reduction: a product of up to four residues (<= 16^4 = 65536) is inside mod17's exact range, so chains of
  // multiplications are reduced once (the host test build records the largest argument mod17 ever receives).
  uint32_t z2 = mul17(z, z), z3 = mul17(z2, z), z4 = mul17(z2, z2);
  uint32_t zh_z = sub17(z4, 1u);
  uint32_t l1_z = mod17(K.L1[0] + K.L1[1] * z + K.L1[2] * z2 + K.L1[3] * z3);
  uint32_t l1a2 = mod17(l1_z * alpha * alpha);
  uint32_t perm_a = mod17(beta * s1_z + gamma + a_z), perm_b = mod17(beta * s2_z + gamma + b_z);
  uint32_t perm_ab = perm_a * perm_b;                                        // <= 256, kept unreduced
  uint32_t perm = mod17(perm_ab * mod17(c_z + gamma) * zw_z);                // Q3: no alpha here
  bool zh0 = zh_z == 0u;                                                     // Q4
  uint32_t t_z = mul17(mod17(r_z + 34u - perm - l1a2), T.inv17[zh_z]);

  // ---- scalars of Steps 8-11                                                src/plonk.rs:583-644
  uint32_t v2 = mul17(v, v), v3 = mul17(v2, v), v4 = mul17(v3, v), v5 = mul17(v4, v), v6 = mul17(v5, v);
  uint32_t z6 = mul17(z4, z2), z12 = mul17(z6, z6);
  uint32_t bz = mul17(beta, z);
  uint32_t s_qm = mod17(a_z * b_z * v), s_ql = mul17(a_z, v), s_qr = mul17(b_z, v), s_qo = mul17(c_z, v), s_qc = v;
  uint32_t f123a = mod17(mod17(a_z + bz + gamma) * mod17(b_z + 2u * bz + gamma) * mod17(c_z + 3u * bz + gamma) * alpha);
  uint32_t s_zs = mod17((f123a + l1a2) * v + u);
  uint32_t s_s3 = mod17(mod17(perm_ab * alpha * v) * beta * zw_z);
  uint32_t s_e = mod17(t_z + v * r_z + v2 * a_z + v3 * b_z + v4 * c_z + v5 * s1_z + v6 * s2_z + u * zw_z);
  uint32_t s_wzw = mod17(u * z * 4u);   // u z omega

  if (ALGO == ALGO_TABLE) {
    // fixed-base part in the exponent of G: d_1 - d_3 + v^5 sigma_1 + v^6 sigma_2 - e
    uint32_t efix = mod17(s_qm * K.vdlog[0] + s_ql * K.vdlog[1] + s_qr * K.vdlog[2] + s_qo * K.vdlog[3] + s_qc * K.vdlog[4] +
                          v5 * K.vdlog[5] + v6 * K.vdlog[6] + neg17(s_s3) * K.vdlog[7] + neg17(s_e));
    uint32_t i1 = mod102(idx[7] + u * idx[8]);
    uint32_t i2 = mod102(z * idx[7] + s_wzw * idx[8] + idx[4] + z6 * idx[5] + z12 * idx[6] + s_zs * idx[3] + v2 * idx[0] +
                         v3 * idx[1] + v4 * idx[2] + 6u * efix);
    e1.a = T.pair_s_a[i1]; e1.b = T.pair_s_b[i1];
    e2.a = T.pair_1_a[i2]; e2.b = T.pair_1_b[i2];
  } else {
    // Exact FP32 curve arithmetic (pbh_g1f.cuh).  The points are on the curve or the verdict is decided by Step 1 alone (the
    // arithmetic below then runs on garbage that stays in range: every table index is bounded by construction).
    const float* inv = T.inv101c;
    // fixed-base multiples (the constant commitments lie in <G>, of order 17, so -[s]P = [17-s]P)
    // three scalars per lookup for the eight constant commitments and G (FixedBaseTables): 3 lookups, 2 additions
    G1F<F32> acc = g1f_unpack<F32>(pair_lookup(T.fixed->vfix_tri[0], s_qm + 17u * s_ql + 289u * s_qr));
    acc = g1f_add(acc, g1f_unpack<F32>(pair_lookup(T.fixed->vfix_tri[1], s_qo + 17u * s_qc + 289u * v5)), inv);
    acc = g1f_add(acc, g1f_unpack<F32>(pair_lookup(T.fixed->vfix_tri[2], v6 + 17u * neg17(s_s3) + 289u * neg17(s_e))), inv);
    // Straus with joint two-point windows: sum_k [sc_k] pt_k with shared doublings, scalars < 17 (5 bits).  The points
    // are taken in pairs; per bit level one addition of 0, P, Q or P + Q (precomputed) serves both scalars of a pair:
    // 4 + 21 additions and 4 doublings instead of 44 + 4.  t_lo_s has the scalar 1 (one addition at the last level).
    // "No addition" is an addend flagged as the identity, which the addition handles with the flag alone.
    const uint32_t sc[9] = {v2, v3, v4, s_zs, 1u, z6, z12, z, s_wzw};
    const int pa[4] = {0, 2, 5, 7}, pb[4] = {1, 3, 6, 8};
    G1F<F32> both[4];
#pragma unroll
    for (int p = 0; p < 4; p++) both[p] = g1f_add(pt[pa[p]], pt[pb[p]], inv);
    G1F<F32> r = g1f_identity<F32>();
#pragma unroll
    for (int j = 4; j >= 0; j--) {
      if (j != 4) r = g1f_add(r, r, inv);
#pragma unroll
      for (int p = 0; p < 4; p++) {
        const bool ba = (sc[pa[p]] >> j) & 1u, bb = (sc[pb[p]] >> j) & 1u;
        G1F<F32> addend;
        addend.x = f_sel(ba, f_sel(bb, both[p].x, pt[pa[p]].x), pt[pb[p]].x);
        addend.y = f_sel(ba, f_sel(bb, both[p].y, pt[pa[p]].y), pt[pb[p]].y);
        addend.inf = ba ? (bb ? both[p].inf : pt[pa[p]].inf) : (bb ? pt[pb[p]].inf : true);
        r = g1f_add(r, addend, inv);
      }
      if (j == 0) r = g1f_add(r, pt[4], inv);
    }
    const G1F<F32> q2 = g1f_add(r, acc, inv);                                          // e_2_q1
    const G1F<F32> q1 = g1f_add(pt[7], g1f_smul<5>(pt[8], u, inv), inv);               // e_1_q1
    const GTF<F32> f1 = pairingf(q1, f_from_u32(K.g2_s[0], ftag), f_from_u32(K.g2_s[1], ftag), inv);   // src/plonk.rs:646
    const GTF<F32> f2 = pairingf(q2, f_from_u32(K.g2_1[0], ftag), f_from_u32(K.g2_1[1], ftag), inv);   // src/plonk.rs:647
    e1.a = f_canon101(f1.a); e1.b = f_canon101(f1.b); e2.a = f_canon101(f2.a); e2.b = f_canon101(f2.b);
  }

  uint32_t res = (e1.a == e2.a && e1.b == e2.b) ? VR_ACCEPT : VR_REJECT_PAIRING;
  if (zh0) res = VR_PANIC_ZH0;
  if (off_field) res = VR_NOT_IN_FIELD;
  if (off_curve) res = VR_NOT_ON_CURVE;
  if (bad) res = VR_BAD_ENCODING;
  if (res != VR_ACCEPT && res != VR_REJECT_PAIRING) e1.a = e1.b = e2.a = e2.b = 0;
  return res;
}

// ---- PBH_ALGO_TABLE verifier with the F_17 scalar work on the FP32 pipes (see pbh_prove_f32.cuh for the rationale and
// the exactness argument; bounds are machine-checked by the CPU test suite through the same template) ----------------
// x mod 102 as an integer index 0..101 for an exact non-negative integer x < 2^20
template <class T>
PBH_HD T f_red102c(T x) {   // centred residue in [-51, 51]
  T* tag = nullptr;
  return f_fma(f_rint_div(x, 102.0f, 0.00980392156862745f, tag), f_const(-102.f, tag), x);
}

// idx[9]: discrete logs (0..101) of the nine proof points, flagged points already mapped to 0.  ev/ch/u: exact small
// integers (evaluations already reduced mod 17 by the caller when >= 17).  Outputs the table indices of e_1_q1 and
// e_2_q1 and whether Z_H(z) = 0.
template <class T>
PBH_HD void verify_scalars_f32(const T (&idx)[9], const T (&ev)[7], const T (&ch)[5], T u, const ConstsF& VF, const float* inv17c,
                               uint32_t& i1, uint32_t& i2, bool& zh0) {
  T* tag = nullptr;
  const T alpha = ch[0], beta = ch[1], gamma = ch[2], z = ch[3], v = ch[4];
  const T a_z = ev[0], b_z = ev[1], c_z = ev[2], s1_z = ev[3], s2_z = ev[4], r_z = ev[5], zw_z = ev[6];
  // Steps 4-7                                                                 src/plonk.rs:553-579
  T z2 = f_red(f_mul(z, z)), z3 = f_red(f_mul(z2, z)), z4 = f_red(f_mul(z2, z2));
  T zh_z = f_red(f_sub(z4, f_const(1.f, tag)));
  T l1_z = f_red(f_fma(f_const(VF.L1[3], tag), z3, f_fma(f_const(VF.L1[2], tag), z2, f_fma(f_const(VF.L1[1], tag), z, f_const(VF.L1[0], tag)))));
  T a2 = f_red(f_mul(alpha, alpha));
  T l1a2 = f_red(f_mul(l1_z, a2));
  T perm_a = f_red(f_add(f_fma(beta, s1_z, gamma), a_z)), perm_b = f_red(f_add(f_fma(beta, s2_z, gamma), b_z));
  T perm_ab = f_red(f_mul(perm_a, perm_b));
  T perm = f_red(f_mul(f_mul(perm_ab, f_add(c_z, gamma)), zw_z));            // Q3: no alpha here
  zh0 = f_is_zero(zh_z);                                                       // Q4
  T t_z = f_red(f_mul(f_sub(f_sub(r_z, perm), l1a2), f_const(inv17c[f_canon(zh_z)], tag)));
  // scalars of Steps 8-11                                                     src/plonk.rs:583-644
  T v2 = f_red(f_mul(v, v)), v3 = f_red(f_mul(v2, v)), v4 = f_red(f_mul(v3, v)), v5 = f_red(f_mul(v4, v)), v6 = f_red(f_mul(v5, v));
  T z6 = f_red(f_mul(z4, z2)), z12 = f_red(f_mul(z6, z6));
  T bz = f_mul(beta, z);
  T av = f_red(f_mul(a_z, v));
  T s_qm = f_red(f_mul(av, b_z)), s_ql = av, s_qr = f_red(f_mul(b_z, v)), s_qo = f_red(f_mul(c_z, v)), s_qc = v;
  T f1 = f_red(f_add(f_add(a_z, bz), gamma)), f2 = f_red(f_add(f_fma(bz, f_const(2.f, tag), b_z), gamma)),
    f3 = f_red(f_add(f_fma(bz, f_const(3.f, tag), c_z), gamma));
  T av_ = f_red(f_mul(alpha, v));
  T s_zs = f_red(f_add(f_fma(f_red(f_mul(f_mul(f1, f2), f3)), av_, f_mul(l1a2, v)), u));
  T s_s3 = f_red(f_mul(f_red(f_mul(perm_ab, av_)), f_red(f_mul(beta, zw_z))));
  T s_e = f_red(f_fma(u, zw_z, f_fma(v6, s2_z, f_fma(v5, s1_z, f_fma(v4, c_z, f_fma(v3, b_z, f_fma(v2, a_z, f_fma(v, r_z, t_z))))))));
  T s_wzw = f_red(f_mul(f_mul(u, z), f_const(4.f, tag)));
  // fixed-base part in the exponent of G (mod 17): d_1 - d_3 + v^5 sigma_1 + v^6 sigma_2 - e
  T efix = f_sub(f_fma(v6, f_const(VF.vdlog[6], tag), f_fma(v5, f_const(VF.vdlog[5], tag), f_fma(s_qc, f_const(VF.vdlog[4], tag),
                 f_fma(s_qo, f_const(VF.vdlog[3], tag), f_fma(s_qr, f_const(VF.vdlog[2], tag), f_fma(s_ql, f_const(VF.vdlog[1], tag),
                 f_mul(s_qm, f_const(VF.vdlog[0], tag)))))))), f_fma(s_s3, f_const(VF.vdlog[7], tag), s_e));
  T efix_c = f_canon_f(f_red(efix), tag);
  // group arithmetic in the exponent of g102 (mod 102); scalars must be the canonical integers 0..16 (gf of src/pbh/mod.rs:30-32)
  T cu = u, cz = z, cw = f_canon_f(s_wzw, tag),   // u and z arrive canonical (0..16)
    c6 = f_canon_f(z6, tag), c12 = f_canon_f(z12, tag),
    czs = f_canon_f(s_zs, tag), c2 = f_canon_f(v2, tag), c3 = f_canon_f(v3, tag), c4 = f_canon_f(v4, tag);
  T x1 = f_fma(cu, idx[8], idx[7]);
  T x2 = f_fma(efix_c, f_const(6.f, tag), f_fma(c4, idx[2], f_fma(c3, idx[1], f_fma(c2, idx[0], f_fma(czs, idx[3], f_fma(c12, idx[6],
         f_fma(c6, idx[5], f_add(f_fma(cw, idx[8], f_mul(cz, idx[7])), idx[4]))))))));
  i1 = f_index102(f_red102c(x1));
  i2 = f_index102(f_red102c(x2));
}

// ---- Fiat-Shamir verifier: replay the transcript over the proof bytes as they are (each point
// as x, y, infinite, 0; the evaluation bytes unreduced), then Plonk::verify with the derived Challange and rand[0] = u.
// `derived` = alpha beta gamma z v u.
template <int ALGO>
PBH_HD uint32_t verify_one_fs(const uint32_t (&px)[9], const uint32_t (&py)[9], uint32_t infbits, const uint32_t (&ev)[7],
                              const uint32_t (&seed)[8], const Consts& K, const Tables& T, GT& e1, GT& e2, const ConstsF* KF,
                              uint32_t (&derived)[6]) {
  uint32_t pk[9];
#pragma unroll
  for (int k = 0; k < 9; k++) pk[k] = (px[k] & 0xFFu) | ((py[k] & 0xFFu) << 8) | (((infbits >> k) & 1u) << 16);
  FsChal<uint32_t, ConvU32, ALGO == ALGO_TABLE ? 31 : 0> cs(seed);   // out-of-line compression where it measured faster (pbh_sha256.cuh)
  uint32_t beta, gamma;
  cs.beta_gamma(pk[0], pk[1], pk[2], beta, gamma);
  cs.alpha(pk[3]);
  cs.zeta(pk[4], pk[5], pk[6]);
  cs.v(ev);
  cs.u(pk[7], pk[8]);
  const uint32_t ch[5] = {cs.derived[0], cs.derived[1], cs.derived[2], cs.derived[3], cs.derived[4]};
  const uint32_t res = verify_one<ALGO>(px, py, infbits, ev, ch, cs.derived[5], K, T, e1, e2, KF);
#pragma unroll
  for (int k = 0; k < 6; k++) derived[k] = (res == VR_BAD_ENCODING) ? 0u : cs.derived[k];
  return res;
}

}  // namespace pbh
