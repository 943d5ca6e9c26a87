// hostemul.cpp — TEST INFRASTRUCTURE.  Compiles the product's __host__ __device__ routines
// (plonk-by-fingers_b200/csrc/pbh_{arith,prove,verify}.cuh, pbh_setup.hpp) with g++ so that the exact
// per-item logic the kernels run can be differential-tested against the oracle on a machine without a GPU.
// It is NOT part of the product and is never loaded by it: the shipped library has no CPU path.
#include <cstdint>
#include <cstring>
#include <string>

#include "../../plonk-by-fingers_b200/csrc/pbh_setup.hpp"

using namespace pbh;

static std::string g_err;

// ---- worst-case magnitude propagation through the FP32 prover core (pbh_prove_f32.cuh) ----------------------------
// Bnd carries an upper bound of |value| for an integer-valued quantity.  Every arithmetic result must stay below 2^24
// (exactly representable) and every reduction input at or below 2^23 (validity range of red17).
namespace pbh {
struct Bnd { double m; Bnd() : m(0) {} explicit Bnd(double x) : m(x) {} };
static double g_max_exact = 0, g_max_red = 0;
static bool g_violation = false;
inline Bnd chk(double m) { if (m > g_max_exact) g_max_exact = m; if (m >= 16777216.0) g_violation = true; return Bnd(m); }
inline Bnd f_const(float c, Bnd*) { return Bnd(c < 0 ? -c : c); }
inline Bnd f_fma(Bnd a, Bnd b, Bnd c) { return chk(a.m * b.m + c.m); }
inline Bnd f_mul(Bnd a, Bnd b) { return chk(a.m * b.m); }
inline Bnd f_add(Bnd a, Bnd b) { return chk(a.m + b.m); }
inline Bnd f_sub(Bnd a, Bnd b) { return chk(a.m + b.m); }
inline Bnd f_red(Bnd x) { if (x.m > g_max_red) g_max_red = x.m; if (x.m > 8388608.0) g_violation = true; return Bnd(8); }
inline bool f_is_zero(Bnd) { return false; }
inline bool f_any(bool, Bnd*) { return true; }     // the bound analysis walks the Q1 path
inline uint32_t f_canon(Bnd) { return 0; }
inline Bnd f_rint_div(Bnd x, float d, float, Bnd*) { if (x.m > g_max_red) g_max_red = x.m; if (x.m >= 2097152.0) g_violation = true; return chk(x.m / d + 1.0); }
inline Bnd f_canon_f(Bnd, Bnd*) { return Bnd(16); }
inline uint32_t f_index102(Bnd) { return 0; }
// F_101 policy (pbh_g1f.cuh): red101 is exact for |x| <= 2^21, the inverse lookup takes an integer |s| <= 202
inline Bnd f_red101(Bnd x) { if (x.m > g_max_red) g_max_red = x.m; if (x.m > 2097152.0) g_violation = true; return Bnd(50); }
inline Bnd f_inv101(Bnd s, const float*) { if (s.m > 202.0) g_violation = true; return Bnd(50); }
inline Bnd f_sel(bool, Bnd a, Bnd b) { return Bnd(a.m > b.m ? a.m : b.m); }
inline Bnd f_neg(Bnd a) { return a; }
inline uint32_t f_canon101(Bnd) { return 0; }
inline Bnd f_from_byte(uint32_t, int, Bnd*) { return Bnd(100); }
inline bool f_eq(Bnd, Bnd) { return false; }
inline uint32_t f_to_index(Bnd x) { if (x.m >= 8192.0) g_violation = true; return 0; }
}  // namespace pbh

extern "C" {

// Runs prove_core_f32 on magnitude bounds: inputs <= 16, every circuit / SRS constant at its centred maximum 8, SRS
// discrete logs at 8, every n_pts from 1 to 10 (all out-of-bounds checks compiled in).  Returns 0 when no operation can
// leave the exact range; max_exact / max_red receive the largest magnitudes seen.
int emul_f32_bounds(double* max_exact, double* max_red) {
  using namespace pbh;
  g_max_exact = g_max_red = 0; g_violation = false;
  ConstsF KF;
  float* kf = reinterpret_cast<float*>(&KF);
  for (size_t i = 0; i < sizeof(KF) / sizeof(float); i++) kf[i] = 8.0f;
  float inv[32];
  for (int i = 0; i < 32; i++) inv[i] = 8.0f;
  for (uint32_t n_pts = 1; n_pts <= 10; n_pts++) {
    Bnd w[12], r[9], c[5];
    for (auto& x : w) x = Bnd(16);
    for (auto& x : r) x = Bnd(16);
    for (auto& x : c) x = Bnd(16);
    ProofF pf;
    Tables tb; std::memset(&tb, 0, sizeof tb);
    static FixedBaseTables fixed_zero{};
    tb.fixed = &fixed_zero;
    prove_core_f32<ALGO_TABLE, Bnd>(w, r, c, RuntimeCK{KF, n_pts}, tb, inv, pf);
    prove_core_f32<ALGO_ARITH, Bnd>(w, r, c, RuntimeCK{KF, n_pts}, tb, inv, pf);
    prove_core_f32<ALGO_TABLE, Bnd>(w, r, c, PbhCK(), tb, inv, pf);
  }
  {
    // the verifier's FP32 scalar path: discrete logs <= 101, evaluations reduced (<= 8), challenges and u <= 16
    Bnd idx[9], ev[7], ch[5];
    for (auto& x : idx) x = Bnd(101);
    for (auto& x : ev) x = Bnd(8);
    for (auto& x : ch) x = Bnd(16);
    uint32_t i1, i2; bool zh0;
    verify_scalars_f32<Bnd>(idx, ev, ch, Bnd(16), KF, inv, i1, i2, zh0);
  }
  {
    // the FP32 curve arithmetic: coordinates as raw bytes (<= 100) or centred residues, G2 coordinates <= 100
    G1F<Bnd> p, q;
    p.x = p.y = q.x = q.y = Bnd(100); p.inf = q.inf = false;
    float inv[408] = {0};
    G1F<Bnd> r = g1f_add(p, q, inv);
    r = g1f_add(r, p, inv);
    r = g1f_smul<7>(p, 127u, inv);
    (void)g1f_in_curve(Bnd(100), Bnd(100));
    GTF<Bnd> e = pairingf(p, Bnd(100), Bnd(100), inv);
    GTF<Bnd> raw; raw.a = raw.b = Bnd(100);
    e = gtf_final_exp(raw, inv);
    (void)e; (void)r;
  }
  *max_exact = g_max_exact; *max_red = g_max_red;
  return g_violation ? 1 : 0;
}

// red101 of pbh_g1f.cuh against x mod 101 for every integer |x| <= 2^21; returns the number of mismatches (the result must
// be THE centred residue in [-50, 50], not merely congruent: zero tests and table indices rely on it)
uint64_t emul_check_red101_f32() {
  using namespace pbh;
  uint64_t bad = 0;
  for (int64_t x = -2097152; x <= 2097152; x++) {
    float r = f_red101(F32((float)x)).v;
    int64_t m = ((x % 101) + 101) % 101;
    if (m > 50) m -= 101;
    if ((int64_t)r != m || r != (float)(int64_t)r) bad++;
  }
  return bad;
}

// red17 of pbh_prove_f32.cuh against x mod 17 for every integer |x| <= 2^23; returns the number of mismatches
uint64_t emul_check_red17_f32() {
  using namespace pbh;
  uint64_t bad = 0;
  for (int64_t x = -8388608; x <= 8388608; x++) {
    float r = f_red(F32((float)x)).v;
    int64_t m = ((x % 17) + 17) % 17;
    int64_t c = m > 8 ? m - 17 : m;
    bad += (r != (float)c);
    bad += f_canon(F32(r)) != (uint32_t)m;
  }
  return bad;
}

const char* emul_last_error() { return g_err.c_str(); }

// exhaustive check of the reciprocal constants over their documented ranges; returns the number of mismatches
// largest arguments mod17 / mod101 / mod102 have received so far in this process
uint32_t emul_mod_max_argument(int which) { return g_mod_max[which]; }

uint64_t emul_check_reductions() {
  uint64_t bad = 0;
  for (uint32_t x = 0; x < 69631u; x++) bad += (x - 17u * ((x * 61681u) >> 20)) != x % 17u;
  for (uint32_t x = 0; x < 103000u; x++) bad += (x - 101u * ((x * 41528u) >> 22)) != x % 101u;
  for (uint32_t x = 0; x < 104000u; x++) bad += (x - 102u * ((x * 41121u) >> 22)) != x % 102u;
  uint32_t keep[3] = {g_mod_max[0], g_mod_max[1], g_mod_max[2]};   // the sweep below is not a use of the arithmetic
  for (uint32_t x = 0; x < 60000u; x++) bad += (mod17(x) != x % 17u) + (mod101(x) != x % 101u) + (mod102(x) != x % 102u);
  for (int i = 0; i < 3; i++) g_mod_max[i] = keep[i];
  return bad;
}

int emul_setup(const pbh_circuit* c, uint8_t s, uint32_t srs_n, uint8_t omega_pows, uint8_t* g1s, uint8_t g2[4], uint8_t consts[24],
               uint8_t* tables_out, size_t tables_cap) {
  HostSetup hs;
  int rc = host_setup(*c, s, srs_n, omega_pows, hs, g_err);
  if (rc) return rc;
  for (size_t i = 0; i < hs.g1s.size(); i++) { g1s[3 * i] = hs.g1s[i].x; g1s[3 * i + 1] = hs.g1s[i].y; g1s[3 * i + 2] = hs.g1s[i].inf; }
  g2[0] = hs.g2_1[0]; g2[1] = hs.g2_1[1]; g2[2] = hs.g2_s[0]; g2[3] = hs.g2_s[1];
  for (int j = 0; j < 8; j++) { consts[3 * j] = hs.vconst[j].x; consts[3 * j + 1] = hs.vconst[j].y; consts[3 * j + 2] = hs.vconst[j].inf; }
  if (tables_out && tables_cap >= sizeof(Tables)) std::memcpy(tables_out, &hs.T, sizeof(Tables));
  return 0;
}
size_t emul_tables_size() { return sizeof(Tables); }

int emul_prove_batch(const pbh_circuit* c, uint8_t s, uint32_t srs_n, uint8_t omega_pows, int algo, size_t n, const uint8_t* wit,
                     const uint8_t* rnd, const uint8_t* chal, uint8_t* proof, uint8_t* status) {
  HostSetup hs;
  int rc = host_setup(*c, s, srs_n, omega_pows, hs, g_err);
  if (rc) return rc;
  for (size_t i = 0; i < n; i++) {
    uint32_t w[12], r[9], ch[5];
    bool bad = false;
    for (int k = 0; k < 12; k++) { w[k] = wit[k * n + i]; bad |= w[k] >= 17; }
    for (int k = 0; k < 9; k++) { r[k] = rnd[k * n + i]; bad |= r[k] >= 17; }
    for (int k = 0; k < 5; k++) { ch[k] = chal[k * n + i]; bad |= ch[k] >= 17; }
    if (bad) { std::memset(w, 0, sizeof w); std::memset(r, 0, sizeof r); std::memset(ch, 0, sizeof ch); }
    ProofRegs P;
    const bool special = consts_match_pbh(hs.KF, hs.K.n_pts);
    uint32_t st = algo == 4 ? (special ? prove_item_f32<ALGO_TABLE, true>(w, r, ch, hs.K, hs.KF, hs.T, P) : 0xFFu)
                : algo == 3 ? prove_item_f32<ALGO_ARITH>(w, r, ch, hs.K, hs.KF, hs.T, P)
                : algo == 2 ? prove_item_f32<ALGO_TABLE>(w, r, ch, hs.K, hs.KF, hs.T, P)
                            : (algo == 1 ? prove_one<ALGO_TABLE>(w, r, ch, hs.K, hs.T, P) : prove_one<ALGO_ARITH>(w, r, ch, hs.K, hs.T, P));
    if (bad) st = PBH_ST_BAD_ENCODING;
    for (int k = 0; k < 27; k++) proof[k * n + i] = 0;
    status[i] = (uint8_t)st;
    if (st == 0) {
      for (int k = 0; k < 9; k++) {
        proof[(2 * k) * n + i] = P.pt[k] & 0xFF; proof[(2 * k + 1) * n + i] = (P.pt[k] >> 8) & 0xFF;
        if ((P.pt[k] >> 16) & 1) { if (k < 8) proof[18 * n + i] |= 1u << k; else proof[19 * n + i] |= 1; }
      }
      for (int k = 0; k < 7; k++) proof[(20 + k) * n + i] = (uint8_t)P.ev[k];
    }
  }
  return 0;
}

int emul_verify_batch(const pbh_circuit* c, uint8_t s, uint32_t srs_n, uint8_t omega_pows, int algo, size_t n, const uint8_t* proof,
                      const uint8_t* chal, const uint8_t* u, uint8_t* result, uint8_t* gt) {
  HostSetup hs;
  int rc = host_setup(*c, s, srs_n, omega_pows, hs, g_err);
  if (rc) return rc;
  for (size_t i = 0; i < n; i++) {
    uint32_t px[9], py[9], ev[7], ch[5];
    for (int k = 0; k < 9; k++) { px[k] = proof[(2 * k) * n + i]; py[k] = proof[(2 * k + 1) * n + i]; }
    uint32_t infbits = proof[18 * n + i] | ((uint32_t)proof[19 * n + i] << 8);
    for (int k = 0; k < 7; k++) ev[k] = proof[(20 + k) * n + i];
    for (int k = 0; k < 5; k++) ch[k] = chal[k * n + i];
    GT e1, e2;
    uint32_t res = algo == 2 ? verify_one<ALGO_TABLE>(px, py, infbits, ev, ch, u[i], hs.K, hs.T, e1, e2, &hs.KF)
                   : (algo == 1 ? verify_one<ALGO_TABLE>(px, py, infbits, ev, ch, u[i], hs.K, hs.T, e1, e2)
                                : verify_one<ALGO_ARITH>(px, py, infbits, ev, ch, u[i], hs.K, hs.T, e1, e2));
    result[i] = (uint8_t)res;
    if (gt) { gt[i] = e1.a; gt[n + i] = e1.b; gt[2 * n + i] = e2.a; gt[3 * n + i] = e2.b; }
  }
  return 0;
}

// ---- Fiat-Shamir item routines (pbh_fs.cuh, pbh_sha256.cuh) ----
int emul_sha256(const uint8_t* data, size_t len, uint8_t out[32]) {
  uint32_t h[8];
  sha256_host(data, len, h);
  for (int i = 0; i < 8; i++) for (int b = 0; b < 4; b++) out[4 * i + b] = (uint8_t)(h[i] >> (24 - 8 * b));
  return 0;
}
// state <- SHA-256(state || msg) through the single-compression device routine; len <= 15
int emul_sha256_absorb(uint8_t state[32], const uint8_t* msg, int len) {
  uint32_t st[8], m[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 8; i++) st[i] = ((uint32_t)state[4 * i] << 24) | ((uint32_t)state[4 * i + 1] << 16) | ((uint32_t)state[4 * i + 2] << 8) | state[4 * i + 3];
  for (int i = 0; i < len; i++) m[i / 4] |= (uint32_t)msg[i] << (24 - 8 * (i % 4));
  sha256_absorb(st, m, len);
  for (int i = 0; i < 8; i++) for (int b = 0; b < 4; b++) state[4 * i + b] = (uint8_t)(st[i] >> (24 - 8 * b));
  return 0;
}
int emul_fs_seed(const pbh_circuit* c, uint8_t s, uint32_t srs_n, uint8_t omega_pows, uint8_t out[32]) {
  HostSetup hs;
  int rc = host_setup(*c, s, srs_n, omega_pows, hs, g_err);
  if (rc) return rc;
  for (int i = 0; i < 8; i++) for (int b = 0; b < 4; b++) out[4 * i + b] = (uint8_t)(hs.fs_seed[i] >> (24 - 8 * b));
  return 0;
}
// algo: 0 ARITH int, 1 TABLE int, 2 TABLE f32, 3 ARITH f32, 4 TABLE f32 specialised
int emul_prove_fs_batch(const pbh_circuit* c, uint8_t s, uint32_t srs_n, uint8_t omega_pows, int algo, size_t n, const uint8_t* wit,
                        const uint8_t* rnd, uint8_t* proof, uint8_t* status, uint8_t* chal_out) {
  HostSetup hs;
  int rc = host_setup(*c, s, srs_n, omega_pows, hs, g_err);
  if (rc) return rc;
  for (size_t i = 0; i < n; i++) {
    uint32_t w[12], r[9], d[6];
    bool bad = false;
    for (int k = 0; k < 12; k++) { w[k] = wit[k * n + i]; bad |= w[k] >= 17; }
    for (int k = 0; k < 9; k++) { r[k] = rnd[k * n + i]; bad |= r[k] >= 17; }
    if (bad) { std::memset(w, 0, sizeof w); std::memset(r, 0, sizeof r); }
    ProofRegs P;
    const bool special = consts_match_pbh(hs.KF, hs.K.n_pts);
    uint32_t st = algo == 4 ? (special ? prove_item_fs<ALGO_TABLE, true, true>(w, r, hs.fs_seed, hs.K, hs.KF, hs.T, P, d) : 0xFFu)
                : algo == 3 ? prove_item_fs<ALGO_ARITH, true, false>(w, r, hs.fs_seed, hs.K, hs.KF, hs.T, P, d)
                : algo == 2 ? prove_item_fs<ALGO_TABLE, true, false>(w, r, hs.fs_seed, hs.K, hs.KF, hs.T, P, d)
                : algo == 1 ? prove_item_fs<ALGO_TABLE, false, false>(w, r, hs.fs_seed, hs.K, hs.KF, hs.T, P, d)
                            : prove_item_fs<ALGO_ARITH, false, false>(w, r, hs.fs_seed, hs.K, hs.KF, hs.T, P, d);
    if (bad) st = PBH_ST_BAD_ENCODING;
    for (int k = 0; k < 27; k++) proof[k * n + i] = 0;
    status[i] = (uint8_t)st;
    for (int k = 0; k < 6; k++) chal_out[k * n + i] = st == 0 ? (uint8_t)d[k] : 0;
    if (st == 0) {
      for (int k = 0; k < 9; k++) {
        proof[(2 * k) * n + i] = P.pt[k] & 0xFF; proof[(2 * k + 1) * n + i] = (P.pt[k] >> 8) & 0xFF;
        if ((P.pt[k] >> 16) & 1) { if (k < 8) proof[18 * n + i] |= 1u << k; else proof[19 * n + i] |= 1; }
      }
      for (int k = 0; k < 7; k++) proof[(20 + k) * n + i] = (uint8_t)P.ev[k];
    }
  }
  return 0;
}
int emul_verify_fs_batch(const pbh_circuit* c, uint8_t s, uint32_t srs_n, uint8_t omega_pows, int algo, size_t n, const uint8_t* proof,
                         uint8_t* result, uint8_t* chal_out, uint8_t* gt) {
  HostSetup hs;
  int rc = host_setup(*c, s, srs_n, omega_pows, hs, g_err);
  if (rc) return rc;
  for (size_t i = 0; i < n; i++) {
    uint32_t px[9], py[9], ev[7], d[6];
    for (int k = 0; k < 9; k++) { px[k] = proof[(2 * k) * n + i]; py[k] = proof[(2 * k + 1) * n + i]; }
    uint32_t infbits = proof[18 * n + i] | ((uint32_t)proof[19 * n + i] << 8);
    for (int k = 0; k < 7; k++) ev[k] = proof[(20 + k) * n + i];
    GT e1, e2;
    uint32_t res = algo == 2 ? verify_one_fs<ALGO_TABLE>(px, py, infbits, ev, hs.fs_seed, hs.K, hs.T, e1, e2, &hs.KF, d)
                   : (algo == 1 ? verify_one_fs<ALGO_TABLE>(px, py, infbits, ev, hs.fs_seed, hs.K, hs.T, e1, e2, nullptr, d)
                                : verify_one_fs<ALGO_ARITH>(px, py, infbits, ev, hs.fs_seed, hs.K, hs.T, e1, e2, nullptr, d));
    result[i] = (uint8_t)res;
    for (int k = 0; k < 6; k++) chal_out[k * n + i] = (uint8_t)d[k];
    if (gt) { gt[i] = e1.a; gt[n + i] = e1.b; gt[2 * n + i] = e2.a; gt[3 * n + i] = e2.b; }
  }
  return 0;
}

// G1 / pairing primitives of pbh_arith.cuh
int emul_g1_add(const uint8_t p[3], const uint8_t q[3], uint8_t out[3]) {
  HostSetup hs; pbh_circuit c; std::memset(&c, 0, sizeof c);
  for (int i = 0; i < 4; i++) c.c_a_index[i] = c.c_b_index[i] = c.c_c_index[i] = 1;
  if (host_setup(c, 2, 6, 4, hs, g_err)) return -1;
  G1 a{p[0], p[1], (uint32_t)(p[2] != 0)}, b{q[0], q[1], (uint32_t)(q[2] != 0)};
  bool bad;
  G1 r = g1_add(a, b, hs.T.inv101, &bad);
  out[0] = r.x; out[1] = r.y; out[2] = r.inf;
  return bad ? 1 : 0;
}
// all 102 x 102 additions, all 102 x 101 scalar multiples, and pairings of all 102 points (plus flagged points
// with coordinates) against q; outputs as flat arrays for the caller to compare with the oracle
int emul_g1_smul(const uint8_t p[3], uint8_t k, uint8_t out[3]) {
  static HostSetup hs; static bool init = false;
  if (!init) { pbh_circuit c; std::memset(&c, 0, sizeof c); for (int i = 0; i < 4; i++) c.c_a_index[i] = c.c_b_index[i] = c.c_c_index[i] = 1;
    if (host_setup(c, 2, 6, 4, hs, g_err)) return -1; init = true; }
  G1 a{p[0], p[1], (uint32_t)(p[2] != 0)};
  G1 r = g1_smul<7>(a, k, hs.T.inv101);
  out[0] = r.x; out[1] = r.y; out[2] = r.inf;
  return 0;
}
int emul_pairing(const uint8_t p[3], const uint8_t q[2], uint8_t out[2], uint8_t miller_out[2]) {
  static HostSetup hs; static bool init = false;
  if (!init) { pbh_circuit c; std::memset(&c, 0, sizeof c); for (int i = 0; i < 4; i++) c.c_a_index[i] = c.c_b_index[i] = c.c_c_index[i] = 1;
    if (host_setup(c, 2, 6, 4, hs, g_err)) return -1; init = true; }
  G1 a{p[0], p[1], (uint32_t)(p[2] != 0)};
  GT f = miller_f17(a, q[0], q[1], hs.T.inv101);
  GT e = gt_final_exp(f, hs.T.inv101);
  out[0] = e.a; out[1] = e.b; miller_out[0] = f.a; miller_out[1] = f.b;
  return 0;
}
// the same three through the exact FP32 curve arithmetic of pbh_g1f.cuh (what PBH_ALGO_ARITH runs on the device)
static const HostSetup& shared_setup() {
  static HostSetup hs; static bool init = false;
  if (!init) { pbh_circuit c; std::memset(&c, 0, sizeof c); for (int i = 0; i < 4; i++) c.c_a_index[i] = c.c_b_index[i] = c.c_c_index[i] = 1;
    host_setup(c, 2, 6, 4, hs, g_err); init = true; }
  return hs;
}
static G1F<F32> g1f_of(const uint8_t p[3]) { G1F<F32> a; a.x = F32((float)p[0]); a.y = F32((float)p[1]); a.inf = p[2] != 0; return a; }
int emul_g1f_add(const uint8_t p[3], const uint8_t q[3], uint8_t out[3]) {
  const HostSetup& hs = shared_setup();
  G1F<F32> a = g1f_of(p), b = g1f_of(q);
  if (a.inf) a = g1f_identity<F32>();
  if (b.inf) b = g1f_identity<F32>();
  const G1F<F32> r = g1f_add(a, b, hs.T.inv101c);
  out[0] = (uint8_t)f_canon101(r.x); out[1] = (uint8_t)f_canon101(r.y); out[2] = r.inf ? 1 : 0;
  return 0;
}
int emul_g1f_smul(const uint8_t p[3], uint8_t k, uint8_t out[3]) {
  const HostSetup& hs = shared_setup();
  const G1F<F32> r = g1f_smul<7>(g1f_of(p), k, hs.T.inv101c);
  out[0] = (uint8_t)f_canon101(r.x); out[1] = (uint8_t)f_canon101(r.y); out[2] = r.inf ? 1 : 0;
  return 0;
}
int emul_pairingf(const uint8_t p[3], const uint8_t q[2], uint8_t out[2], uint8_t miller_out[2]) {
  const HostSetup& hs = shared_setup();
  const GTF<F32> f = millerf_f17(g1f_of(p), F32((float)q[0]), F32((float)q[1]), hs.T.inv101c);
  const GTF<F32> e = gtf_final_exp(f, hs.T.inv101c);
  out[0] = (uint8_t)f_canon101(e.a); out[1] = (uint8_t)f_canon101(e.b);
  miller_out[0] = (uint8_t)f_canon101(f.a); miller_out[1] = (uint8_t)f_canon101(f.b);
  return 0;
}
int emul_gtf_final_exp(const uint8_t f[2], uint8_t out[2]) {
  const HostSetup& hs = shared_setup();
  GTF<F32> x; x.a = F32((float)f[0]); x.b = F32((float)f[1]);
  const GTF<F32> e = gtf_final_exp(x, hs.T.inv101c);
  out[0] = (uint8_t)f_canon101(e.a); out[1] = (uint8_t)f_canon101(e.b);
  return 0;
}
int emul_gt_final_exp(const uint8_t f[2], uint8_t out[2]) {
  static HostSetup hs; static bool init = false;
  if (!init) { pbh_circuit c; std::memset(&c, 0, sizeof c); for (int i = 0; i < 4; i++) c.c_a_index[i] = c.c_b_index[i] = c.c_c_index[i] = 1;
    if (host_setup(c, 2, 6, 4, hs, g_err)) return -1; init = true; }
  GT x{f[0], f[1]};
  GT e = gt_final_exp(x, hs.T.inv101);
  out[0] = e.a; out[1] = e.b;
  return 0;
}

}  // extern "C"
