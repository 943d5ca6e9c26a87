// pbh_setup.hpp — context creation on the host: SRS::create (src/plonk.rs:35-48), Plonk::new
// (src/plonk.rs:120-175) and every circuit-constant value the reference recomputes per call
// (src/plonk.rs:222-243, 328-333, 506-517, 557-562), using the same arithmetic routines as the kernels
// (pbh_arith.cuh is __host__ __device__).  Runs once per context; the results are uploaded as `Consts`
// (kernel parameter) and `Tables` (global memory, staged into shared memory by each block).
#pragma once
#include <string>
#include <vector>
#include <cstring>

#include "../../include/pbh_b200.h"
#include "pbh_verify.cuh"
#include "pbh_prove_f32.cuh"

namespace pbh {

struct HostSetup {
  Consts K;
  ConstsF KF;   // the same constants as centred floats, for the FP32 prover
  Tables T;
  std::vector<G1> g1s;        // SRS.g1s
  uint32_t g2_1[2], g2_s[2];  // SRS.g2_1, SRS.g2_s
  G1 vconst[8];               // q_m_s q_l_s q_r_s q_o_s q_c_s sigma_1_s sigma_2_s sigma_3_s
  uint32_t g102_idx_of_G;     // == 6
  uint32_t fs_seed[8];        // initial state of the Fiat-Shamir transcript (include/pbh_b200.h), SHA-256 words
  FixedBaseTables P;          // three-point fixed-base tables; T.fixed points here until the context uploads them
};

// "G2" of src/pbh/g2.rs:58-101: (a, b*u), u^2 = -2, no identity; returns false where the reference panics (Q12)
struct G2h { uint32_t a, b; };
inline bool g2_add(const G2h& p, const G2h& q, const uint8_t* inv101, G2h& out) {
  if (p.a == q.a && p.b == q.b) {
    uint32_t den = mod101(2u * p.b);
    if (den == 0) return false;                                   // (three*a^2 / (two*b)).unwrap()
    uint32_t m_u = mul101(mod101(3u * p.a * p.a), inv101[den]);
    uint32_t u2inv = inv101[99];                                  // 1/(u^2) = (-2)^-1
    uint32_t m2 = mul101(mul101(m_u, m_u), u2inv);
    uint32_t na = mod101(m2 + 202u - 2u * p.a);
    uint32_t nb = mod101(mul101(mul101(u2inv, m_u), mod101(3u * p.a + 101u - m2)) + 101u - p.b);
    out.a = na; out.b = nb;   // `out` may alias p or q
    return true;
  }
  uint32_t den = sub101(q.a, p.a);
  if (den == 0) return false;                                     // ((rhs.b - self.b) / (rhs.a - self.a)).unwrap()
  uint32_t l_u = mul101(sub101(q.b, p.b), inv101[den]);
  uint32_t l2 = mul101(mul101(l_u, l_u), 99u);                    // lambda_u^2 * (-2)
  uint32_t a = mod101(l2 + 202u - p.a - q.a);
  uint32_t b = mod101(mul101(l_u, sub101(p.a, a)) + 101u - p.b);
  out.a = a; out.b = b;
  return true;
}
inline bool g2_mul(const G2h& p, uint32_t k, const uint8_t* inv101, G2h& out) {
  bool have = false;
  G2h result{0, 0}, base = p;
  while (k > 0) {
    if (k & 1u) {
      if (have) { if (!g2_add(result, base, inv101, result)) return false; }
      else { result = base; have = true; }
    }
    k >>= 1;
    if (!g2_add(base, base, inv101, base)) return false;   // the reference doubles once more after the last bit
  }
  if (!have) return false;                                  // result.unwrap() on None (scalar 0)
  out = result;
  return true;
}

inline uint32_t pow_mod(uint32_t b, uint32_t e, uint32_t m) {
  uint32_t r = 1 % m;
  b %= m;
  while (e) { if (e & 1) r = (r * b) % m; b = (b * b) % m; e >>= 1; }
  return r;
}

// Returns PBH_OK or an error; `err` receives a description.
inline int host_setup(const pbh_circuit& c, uint32_t srs_secret, uint32_t srs_n, uint32_t omega_pows, HostSetup& hs,
                      std::string& err) {
  std::memset(&hs.K, 0, sizeof(hs.K));
  std::memset(&hs.T, 0, sizeof(hs.T));
  Consts& K = hs.K;
  Tables& T = hs.T;
  if (omega_pows != 4) { err = "omega_pows must be 4 (H = <4> in F_17; prove() is hard-wired to 4 gates)"; return PBH_ERR_UNSUPPORTED; }
  if (srs_n < 3 || srs_n > 64) { err = "srs_n must be in [3, 64]"; return PBH_ERR_UNSUPPORTED; }
  if (srs_secret >= 101) { err = "srs_secret must be a canonical F_101 value"; return PBH_ERR_BAD_ARGUMENT; }

  // inverse tables (the inverse is unique, so any method is bit-exact with the reference's extended GCD)
  for (uint32_t a = 1; a < 17; a++) T.inv17[a] = (uint8_t)pow_mod(a, 15, 17);
  for (uint32_t a = 1; a < 101; a++) T.inv101[a] = (uint8_t)pow_mod(a, 99, 101);

  // ---- SRS::create                                                          src/plonk.rs:35-48
  const G1 G = g1_make(1, 2);
  hs.g1s.clear();
  hs.g1s.push_back(G);
  uint32_t s_pow = srs_secret;
  for (uint32_t k = 0; k < srs_n; k++) {
    hs.g1s.push_back(g1_smul<7>(G, s_pow, T.inv101));
    s_pow = mul101(s_pow, srs_secret);                    // Q11: reduced mod 101, not mod 17
  }
  G2h g2gen{36, 31}, g2s;
  if (!g2_mul(g2gen, srs_secret, T.inv101, g2s)) {
    err = "SRS::create panics in the reference: G2 * s is not defined for this s (src/pbh/g2.rs:58-101)";
    return PBH_ERR_SETUP_PANIC;
  }
  hs.g2_1[0] = 36; hs.g2_1[1] = 31; hs.g2_s[0] = g2s.a; hs.g2_s[1] = g2s.b;
  K.g2_1[0] = 36; K.g2_1[1] = 31; K.g2_s[0] = g2s.a; K.g2_s[1] = g2s.b;

  // ---- [e]G and discrete logs in <G>
  G1 eg[17];
  eg[0] = g1_identity();
  for (int e = 1; e < 17; e++) eg[e] = g1_add(eg[e - 1], G, T.inv101);
  for (int e = 0; e < 17; e++) T.pt17[e] = g1_pack(eg[e]);
  auto dlog17 = [&](const G1& p) -> int {
    for (int e = 0; e < 17; e++) if (g1_pack(eg[e]) == g1_pack(p)) return e;
    return -1;
  };
  K.n_pts = (uint32_t)hs.g1s.size();
  for (size_t i = 0; i < hs.g1s.size() && i < 10; i++) {
    int d = dlog17(hs.g1s[i]);
    if (d < 0) { err = "internal: SRS point outside <G>"; return PBH_ERR_BAD_ARGUMENT; }
    K.srs_dlog[i] = (uint32_t)d;
  }

  // ---- Plonk::new                                                           src/plonk.rs:120-175
  const uint32_t h[4] = {1, 4, 16, 13};      // OMEGA^n, OMEGA = 4; K1 = 2, K2 = 3 are not in H, K2 not in K1*H
  // ---- circuit                                                              src/constraints.rs:109-153
  auto sel = [&](const uint8_t* v, uint32_t (&raw)[4], uint32_t (&poly)[4]) -> bool {
    for (int i = 0; i < 4; i++) { if (v[i] >= 17) return false; raw[i] = v[i]; }
    intt4(raw[0], raw[1], raw[2], raw[3], poly);
    return true;
  };
  if (!sel(c.q_l, K.q_l, K.QL) || !sel(c.q_r, K.q_r, K.QR) || !sel(c.q_o, K.q_o, K.QO) || !sel(c.q_m, K.q_m, K.QM) ||
      !sel(c.q_c, K.q_c, K.QC)) {
    err = "selector values must be canonical F_17 values";
    return PBH_ERR_BAD_ARGUMENT;
  }
  const uint8_t* wires[3] = {c.c_a_wire, c.c_b_wire, c.c_c_wire};
  const uint8_t* idxs[3] = {c.c_a_index, c.c_b_index, c.c_c_index};
  for (int wv = 0; wv < 3; wv++) {
    for (int n = 0; n < 4; n++) {
      uint32_t wire = wires[wv][n], index = idxs[wv][n];
      if (wire > 2) { err = "copy constraint wire must be PBH_COPY_A/B/C"; return PBH_ERR_BAD_ARGUMENT; }
      if (index < 1 || index > 4) {
        err = "copy constraint index outside 1..4: the reference indexes h[n-1] out of bounds (src/plonk.rs:181-189)";
        return PBH_ERR_SETUP_PANIC;
      }
      // copy_constraints_to_roots: A(n) -> h[n-1], B(n) -> k1*h[n-1], C(n) -> k2*h[n-1]
      K.sig[wv][n] = mod17((wire + 1u) * h[index - 1]);
      K.perm[wv * 4 + n] = wire * 4 + (index - 1);
    }
    intt4(K.sig[wv][0], K.sig[wv][1], K.sig[wv][2], K.sig[wv][3], K.S[wv]);
  }
  intt4(1, 0, 0, 0, K.L1);

  // ---- verifier preprocessing: eval_at_s of the 8 constant polynomials      src/plonk.rs:510-517
  const uint32_t* polys[8] = {K.QM, K.QL, K.QR, K.QO, K.QC, K.S[0], K.S[1], K.S[2]};
  for (int j = 0; j < 8; j++) {
    uint32_t e = 0;
    for (int i = 0; i < 4; i++) {
      if (polys[j][i] != 0 && (uint32_t)i >= K.n_pts) { err = "verifier preprocessing panics: SRS too short"; return PBH_ERR_SETUP_PANIC; }
      e += polys[j][i] * K.srs_dlog[i];
    }
    K.vdlog[j] = e % 17;
    hs.vconst[j] = eg[K.vdlog[j]];
  }

  // ---- fixed-base multiples for PBH_ALGO_ARITH
  for (int i = 0; i < 10; i++) {
    G1 base = (size_t)i < hs.g1s.size() ? hs.g1s[i] : g1_identity();
    G1 m = g1_identity();
    for (int k = 0; k < 17; k++) { T.srs_mult[i][k] = g1_pack(m); m = g1_add(m, base, T.inv101); }
  }
  for (int j = 0; j < 9; j++) {
    G1 base = j < 8 ? hs.vconst[j] : G;
    G1 m = g1_identity();
    for (int k = 0; k < 17; k++) { T.vfix_mult[j][k] = g1_pack(m); m = g1_add(m, base, T.inv101); }
  }

  // three-point tables: entry a + 17 b + 289 c = [a]P_3j + [b]P_3j+1 + [c]P_3j+2 (identity for points beyond the table)
  auto mult = [&](const uint32_t (*tab)[17], int rows, int i, int k) { return i < rows ? g1_unpack(tab[i][k]) : g1_identity(); };
  for (int j = 0; j < 4; j++)
    for (int cc = 0; cc < 17; cc++)
      for (int b = 0; b < 17; b++)
        for (int a = 0; a < 17; a++)
          // centred digits: entry (a, b, cc) holds the multiples (a - 8, b - 8, cc - 8) mod 17
          hs.P.srs_tri[j][a + 17 * b + 289 * cc] = g1_pack(g1_add(g1_add(mult(T.srs_mult, 10, 3 * j, (a + 9) % 17), mult(T.srs_mult, 10, 3 * j + 1, (b + 9) % 17), T.inv101),
                                                                mult(T.srs_mult, 10, 3 * j + 2, (cc + 9) % 17), T.inv101));
  for (int j = 0; j < 3; j++)
    for (int cc = 0; cc < 17; cc++)
      for (int b = 0; b < 17; b++)
        for (int a = 0; a < 17; a++)
          hs.P.vfix_tri[j][a + 17 * b + 289 * cc] = g1_pack(g1_add(g1_add(mult(T.vfix_mult, 9, 3 * j, a), mult(T.vfix_mult, 9, 3 * j + 1, b), T.inv101),
                                                                 mult(T.vfix_mult, 9, 3 * j + 2, cc), T.inv101));
  T.fixed = &hs.P;

  // ---- group structure of E(F_101) for PBH_ALGO_TABLE
  std::vector<G1> pts;
  for (uint32_t x = 0; x < 101; x++)
    for (uint32_t y = 0; y < 101; y++)
      if (g1_in_curve(x, y)) pts.push_back(g1_make(x, y));
  if (pts.size() != 101) { err = "internal: curve order"; return PBH_ERR_BAD_ARGUMENT; }
  bool found = false;
  G1 gen = g1_identity();
  for (const G1& p : pts) {
    // order of p, and whether [6]p == G
    G1 m = p;
    int order = 1;
    G1 six = g1_identity();
    while (!m.inf) { m = g1_add(m, p, T.inv101); order++; if (order == 6) six = m; }
    if (order == 102 && g1_pack(six) == g1_pack(G)) { gen = p; found = true; break; }
  }
  if (!found) { err = "internal: no generator with [6]g = G"; return PBH_ERR_BAD_ARGUMENT; }
  hs.g102_idx_of_G = 6;
  std::memset(T.y_of_x, 0xFF, sizeof(T.y_of_x));
  G1 m = g1_identity();
  for (uint32_t i = 0; i < 102; i++) {
    T.x_of_idx[i] = (uint8_t)m.x; T.y_of_idx[i] = (uint8_t)m.y;
    if (!m.inf && m.y <= 50) { T.y_of_x[m.x] = (uint8_t)m.y; T.idx_of_x[m.x] = (uint8_t)i; }
    GT es = pairing(m, K.g2_s[0], K.g2_s[1], T.inv101);
    GT e1 = pairing(m, K.g2_1[0], K.g2_1[1], T.inv101);
    T.pair_s_a[i] = (uint8_t)es.a; T.pair_s_b[i] = (uint8_t)es.b;
    T.pair_1_a[i] = (uint8_t)e1.a; T.pair_1_b[i] = (uint8_t)e1.b;
    m = g1_add(m, gen, T.inv101);
  }
  // ---- centred float copies for the FP32 prover (pbh_prove_f32.cuh)
  auto cen = [](uint32_t x) { return x > 8u ? (float)x - 17.0f : (float)x; };
  ConstsF& F = hs.KF;
  for (int i = 0; i < 4; i++) {
    F.q_l[i] = cen(K.q_l[i]); F.q_o[i] = cen(K.q_o[i]); F.q_m[i] = cen(K.q_m[i]); F.q_c[i] = cen(K.q_c[i]);
    F.QL[i] = cen(K.QL[i]); F.QR[i] = cen(K.QR[i]); F.QO[i] = cen(K.QO[i]); F.QM[i] = cen(K.QM[i]); F.QC[i] = cen(K.QC[i]);
    F.L1[i] = cen(K.L1[i]);
    for (int wv = 0; wv < 3; wv++) { F.sig[wv][i] = cen(K.sig[wv][i]); F.S[wv][i] = cen(K.S[wv][i]); }
  }
  for (int i = 0; i < 10; i++) F.srs_dlog[i] = cen(K.srs_dlog[i]);
  for (int i = 0; i < 8; i++) F.vdlog[i] = cen(K.vdlog[i]);
  for (uint32_t a = 0; a < 17; a++) T.inv17c[a] = cen(T.inv17[a]);
  for (int i = 0; i < 408; i++) {                       // f_inv101 of pbh_g1f.cuh: index = s + 202
    const uint32_t r = (uint32_t)(((i - 202) % 101 + 101) % 101), v = T.inv101[r];
    T.inv101c[i] = v > 50u ? (float)v - 101.0f : (float)v;
  }

  // ---- Fiat-Shamir seed: SHA-256(tag || omega_pows || circuit || SRS), see include/pbh_b200.h
  {
    std::vector<uint8_t> m;
    for (const char* t = "plonk-by-fingers/fiat-shamir/v1"; *t; t++) m.push_back((uint8_t)*t);
    m.push_back((uint8_t)omega_pows);
    const uint8_t* fields[11] = {c.q_l, c.q_r, c.q_o, c.q_m, c.q_c, c.c_a_wire, c.c_a_index, c.c_b_wire, c.c_b_index, c.c_c_wire, c.c_c_index};
    for (int f = 0; f < 11; f++) for (int i = 0; i < 4; i++) m.push_back(fields[f][i]);
    m.push_back((uint8_t)hs.g1s.size());
    for (const G1& pnt : hs.g1s) { m.push_back((uint8_t)pnt.x); m.push_back((uint8_t)pnt.y); m.push_back((uint8_t)pnt.inf); m.push_back(0); }
    m.push_back((uint8_t)hs.g2_1[0]); m.push_back((uint8_t)hs.g2_1[1]); m.push_back((uint8_t)hs.g2_s[0]); m.push_back((uint8_t)hs.g2_s[1]);
    sha256_host(m.data(), m.size(), hs.fs_seed);
  }
  return PBH_OK;
}

}  // namespace pbh
