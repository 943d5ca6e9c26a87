// pbh_oracle.hpp — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
//
// A C++17 restatement of the prove/verify path of adria0/plonk-by-fingers, written so that every
// data-dependent decision of the Rust reference (normalised variable-length polynomials, the
// order of panics, the quirks Q1..Q17 of SURVEY.md §2.3) happens here exactly as it does there.
// It deliberately keeps the reference's cost profile (heap-allocated coefficient vectors, one
// extended-GCD inversion per affine G1 addition, recursive Miller function, pow(600)) because it
// doubles as the CPU baseline ("C++ restatement of the reference", never "the Rust crate").
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// use anything in oracle/.  The product (plonk-by-fingers_b200/) never includes this file.
//
// PARITY PINNING: the reference cannot be built here (no cargo/rustc).  This restatement is pinned
// by every known-answer vector in the reference's own unit tests (tests/golden/reference_vectors.json,
// transcribed from the cited file:line) and differential-tested against a second, independently
// written Python restatement (oracle/pyref.py).  Inputs other than the reference's golden tuple
// are "parity unpinned" by the reference itself.
//
// Citations are relative to /root/reference/.
#pragma once
#include <cstdint>
#include <cstddef>
#include <optional>
#include <utility>
#include <vector>
#include <algorithm>

namespace pbh_oracle {

// A Rust panic site, carried as a C++ exception.  `site` is the status code of include/pbh_b200.h.
struct Panic {
  int site;
  const char* what;
};

enum PanicSite : int {
  SITE_UNSATISFIED = 1,   // src/plonk.rs:199   assert!(constraints.satisfies(assigments))
  SITE_ACC_DIV0 = 2,      // src/plonk.rs:297   (dend / dsor).unwrap()
  SITE_T_REMAINDER = 3,   // src/plonk.rs:370   assert_eq!(rem, Poly::zero())
  SITE_T_SLICE = 4,       // src/plonk.rs:376   t_x.coeffs()[12..18]
  SITE_SRS_OOB = 5,       // src/plonk.rs:56    self.g1s[n]
  SITE_WZ_REMAINDER = 6,  // src/plonk.rs:438
  SITE_WZW_REMAINDER = 7, // src/plonk.rs:442
  SITE_ACC_ASSERT = 8,    // src/plonk.rs:307
  SITE_VERIFY_ZH0 = 16,   // src/plonk.rs:579   (… / z_h_z).unwrap()
  SITE_OTHER = 64,        // any other unwrap/assert (G2 without identity, "cannot add", …)
};

// ---------------------------------------------------------------------------------------------
// src/utils/u64field.rs
// ---------------------------------------------------------------------------------------------

// src/utils/u64field.rs:10-25
inline void extended_gcd(int64_t a, int64_t b, int64_t& g, int64_t& x, int64_t& y) {
  int64_t s = 0, old_s = 1, t = 1, old_t = 0, r = b, old_r = a;
  while (r != 0) {
    int64_t quotient = old_r / r;
    old_r -= quotient * r; std::swap(old_r, r);
    old_s -= quotient * s; std::swap(old_s, s);
    old_t -= quotient * t; std::swap(old_t, t);
  }
  g = old_r; x = old_s; y = old_t;
}

// src/utils/u64field.rs:28-228.  Every operator ends in one `% M`, like the reference.
template <uint64_t M>
struct Fp {
  uint64_t v;
  Fp() : v(0) {}
  static Fp from_u64(uint64_t n) { Fp r; r.v = n % M; return r; }          // :95-99
  static Fp from_i64(int64_t n) {                                           // :85-93
    return n < 0 ? -from_u64((uint64_t)(-n)) : from_u64((uint64_t)n);
  }
  static Fp zero() { return from_u64(0); }
  static Fp one() { return from_u64(1); }
  static uint64_t order() { return M; }
  bool is_zero() const { return v == 0; }
  uint64_t as_u64() const { return v; }
  bool in_field() const { return v < M; }                                   // :49-51
  std::optional<Fp> inv() const {                                           // :52-63
    int64_t g, c, unused;
    extended_gcd((int64_t)v, (int64_t)M, g, c, unused);
    if (g != 1) return std::nullopt;
    Fp r; r.v = c < 0 ? (uint64_t)((int64_t)M + c) : (uint64_t)c;
    return r;
  }
  Fp pow(uint64_t e) const {                                                // :64-75
    Fp result = one(), base = *this;
    while (e > 0) {
      if (e % 2 == 1) result = result * base;
      e >>= 1;
      base = base * base;
    }
    return result;
  }
  friend Fp operator+(Fp a, Fp b) { Fp r; r.v = (a.v + b.v) % M; return r; }  // :107-147
  friend Fp operator-(Fp a) { Fp r; r.v = (M - a.v) % M; return r; }          // :162-174
  friend Fp operator-(Fp a, Fp b) { return a + (-b); }                        // :149-160
  friend Fp operator*(Fp a, Fp b) { Fp r; r.v = (a.v * b.v) % M; return r; }  // :176-200
  friend std::optional<Fp> operator/(Fp a, Fp b) {                            // :222-228
    auto i = b.inv();
    if (!i) return std::nullopt;
    return *i * a;
  }
  Fp& operator+=(Fp b) { v = (v + b.v) % M; return *this; }
  Fp& operator-=(Fp b) { *this += -b; return *this; }
  Fp& operator*=(Fp b) { v = (v * b.v) % M; return *this; }
  friend bool operator==(Fp a, Fp b) { return a.v == b.v; }
  friend bool operator!=(Fp a, Fp b) { return a.v != b.v; }
};

template <class T>
inline T unwrap(const std::optional<T>& o, int site, const char* what) {
  if (!o) throw Panic{site, what};
  return *o;
}

// ---------------------------------------------------------------------------------------------
// src/poly.rs
// ---------------------------------------------------------------------------------------------
template <class F>
struct Poly {
  std::vector<F> c;  // c[i] is the coefficient of x^i; always normalised; zero poly is {0}

  Poly() : c{F::zero()} {}
  explicit Poly(std::vector<F> coeffs) : c(std::move(coeffs)) { normalize(); }   // :17-21
  static Poly raw(std::vector<F> coeffs) { Poly p; p.c = std::move(coeffs); return p; }
  static Poly from_i64(std::initializer_list<int64_t> l) {                        // :24-26
    std::vector<F> v; for (auto n : l) v.push_back(F::from_i64(n));
    return Poly(std::move(v));
  }
  static Poly zero() { return raw({F::zero()}); }                                 // :34-36
  static Poly one() { return raw({F::one()}); }                                   // :39-41
  const std::vector<F>& coeffs() const { return c; }
  size_t len() const { return c.size(); }
  size_t degree() const { return c.size() - 1; }                                  // :91-93
  bool is_zero() const { return c.size() == 1 && c[0].is_zero(); }                // :108-110

  void normalize() {                                                              // :96-105
    if (c.size() > 1 && c.back().is_zero()) {
      size_t n = c.size();
      while (n > 0 && c[n - 1].is_zero()) --n;
      c.resize(n == 0 ? 1 : n, F::zero());
    }
  }
  void set(size_t i, F p) {                                                       // :113-119
    if (c.size() < i + 1) c.resize(i + 1, F::zero());
    c[i] = p;
    normalize();
  }
  // (x-p_1)(x-p_2)...                                                            // :64-68
  static Poly z(const std::vector<F>& points) {
    Poly acc = one();
    for (const F& x : points) acc = acc * Poly(std::vector<F>{-x, F::one()});
    return acc;
  }
  // Lagrange interpolation through (x,y) pairs                                   // :45-61
  static Poly lagrange(const std::vector<std::pair<F, F>>& p) {
    size_t k = p.size();
    Poly l = zero();
    for (size_t j = 0; j < k; j++) {
      Poly l_j = one();
      for (size_t i = 0; i < k; i++) {
        if (i != j) {
          F cc = unwrap((p[j].first - p[i].first).inv(), SITE_OTHER, "lagrange x points must be unique");
          l_j = l_j * Poly(std::vector<F>{-(cc * p[i].first), cc});
        }
      }
      l += l_j * p[j].second;
    }
    return l;
  }
  // power-accumulating evaluation (not Horner)                                   // :71-79
  F eval(F x) const {
    F x_pow = F::one();
    F y = c[0];
    for (size_t i = 1; i < c.size(); i++) {
      x_pow *= x;
      y += x_pow * c[i];
    }
    return y;
  }
  // buggy in the reference (double-counts c0); kept for completeness           // :82-88
  F eval_with_pows(const std::vector<F>& x_pow) const {
    F y = c[0];
    for (size_t i = 0; i < c.size(); i++) y += x_pow[i] * c[i];
    return y;
  }

  Poly& operator+=(const Poly& rhs) {                                             // :165-176
    size_t n_max = std::max(c.size(), rhs.c.size());
    for (size_t n = 0; n < n_max; n++) {
      if (n >= c.size()) c.push_back(rhs.c[n]);
      else if (n < rhs.c.size()) c[n] += rhs.c[n];
    }
    normalize();
    return *this;
  }
  Poly& operator+=(F rhs) { c[0] += rhs; normalize(); return *this; }             // :178-183
  Poly& operator-=(F rhs) { c[0] -= rhs; normalize(); return *this; }             // :185-190
  // Q1: the tail of a longer rhs is pushed UN-NEGATED                            // :192-203
  Poly& operator-=(const Poly& rhs) {
    size_t n_max = std::max(c.size(), rhs.c.size());
    for (size_t n = 0; n < n_max; n++) {
      if (n >= c.size()) c.push_back(rhs.c[n]);
      else if (n < rhs.c.size()) c[n] -= rhs.c[n];
    }
    normalize();
    return *this;
  }
  // schoolbook into a len_a+len_b buffer, then trim                              // :205-218
  friend Poly operator*(const Poly& a, const Poly& b) {
    std::vector<F> mul(a.c.size() + b.c.size(), F::zero());
    for (size_t n = 0; n < a.c.size(); n++)
      for (size_t m = 0; m < b.c.size(); m++) mul[n + m] += a.c[n] * b.c[m];
    Poly r = raw(std::move(mul));
    r.normalize();
    return r;
  }
  // Q15: *= 0 gives Poly::zero(); otherwise no normalisation                     // :220-228
  Poly& operator*=(F rhs) {
    if (rhs.is_zero()) *this = zero();
    else for (auto& v : c) v = v * rhs;
    return *this;
  }
  friend Poly operator*(Poly a, F b) { a *= b; return a; }                        // :346-376
  friend Poly operator+(Poly a, const Poly& b) { a += b; return a; }              // :250-282
  friend Poly operator+(Poly a, F b) { a += b; return a; }                        // :284-315
  friend Poly operator-(Poly a, F b) { a -= b; return a; }                        // :300-306
  friend Poly operator-(Poly a, const Poly& b) { a -= b; return a; }              // :317-323
  friend bool operator==(const Poly& a, const Poly& b) { return a.c == b.c; }
  friend bool operator!=(const Poly& a, const Poly& b) { return !(a == b); }
};

// long division, allocating a monomial and a product per step                    // src/poly.rs:230-247
template <class F>
inline std::pair<Poly<F>, Poly<F>> poly_div(Poly<F> self, const Poly<F>& rhs) {
  Poly<F> q = Poly<F>::zero();
  Poly<F> r = std::move(self);
  while (!r.is_zero() && r.degree() >= rhs.degree()) {
    F lead_r = r.c.back();
    F lead_d = rhs.c.back();
    Poly<F> t = Poly<F>::zero();
    t.set(r.c.size() - rhs.c.size(), lead_r * unwrap(lead_d.inv(), SITE_OTHER, "div: lead inverse"));
    q += t;
    r -= rhs * t;
  }
  q.normalize();
  r.normalize();
  return {q, r};
}

// ---------------------------------------------------------------------------------------------
// src/matrix.rs
// ---------------------------------------------------------------------------------------------
template <class F>
struct Matrix {
  size_t m, n;  // rows, cols
  std::vector<F> v;

  static Matrix zero(size_t m, size_t n) { return Matrix{m, n, std::vector<F>(m * n, F::zero())}; }  // :16-20
  static Matrix from_u64(const std::vector<uint64_t>& vals, size_t m, size_t n) {                    // :26-33
    if (vals.size() != m * n) throw Panic{SITE_OTHER, "matrix: size"};
    Matrix r = zero(m, n);
    for (size_t i = 0; i < vals.size(); i++) r.v[i] = F::from_u64(vals[i]);
    return r;
  }
  size_t cols() const { return n; }
  size_t rows() const { return m; }
  F& at(size_t r, size_t c) {                                                                         // :107-121
    if (!(r < m && c < n)) throw Panic{SITE_OTHER, "matrix: index"};
    return v[c + r * n];
  }
  const F& at(size_t r, size_t c) const {
    if (!(r < m && c < n)) throw Panic{SITE_OTHER, "matrix: index"};
    return v[c + r * n];
  }
  Matrix inv() const {                                                                                // :40-59
    size_t len = n;
    Matrix aug = zero(len, len * 2);
    for (size_t i = 0; i < len; i++) {
      for (size_t j = 0; j < len; j++) aug.at(i, j) = at(i, j);
      aug.at(i, i + len) = F::one();
    }
    aug.gauss_jordan_general();
    Matrix un = zero(len, len);
    for (size_t i = 0; i < len; i++)
      for (size_t j = 0; j < len; j++) un.at(i, j) = aug.at(i, j + len);
    return un;
  }
  void gauss_jordan_general() {                                                                       // :61-104
    size_t lead = 0, row_count = m, col_count = n;
    for (size_t r = 0; r < row_count; r++) {
      if (col_count <= lead) break;
      size_t i = r;
      while (at(i, lead) == F::zero()) {
        i += 1;
        if (row_count == i) {
          i = r;
          lead += 1;
          if (col_count == lead) break;
        }
      }
      for (size_t col = 0; col < n; col++) std::swap(v[n * i + col], v[n * r + col]);
      if (at(r, lead) != F::zero()) {
        F div = at(r, lead);
        for (size_t j = 0; j < col_count; j++) at(r, j) = unwrap(at(r, j) / div, SITE_OTHER, "gj: div");
      }
      for (size_t k = 0; k < row_count; k++) {
        if (k != r) {
          F mult = at(k, lead);
          for (size_t j = 0; j < col_count; j++) at(k, j) = at(k, j) - at(r, j) * mult;
        }
      }
      lead += 1;
    }
  }
  friend Matrix operator*(const Matrix& a, const Matrix& b) {                                         // :123-145
    if (a.n != b.m) throw Panic{SITE_OTHER, "matrix: mul shape"};
    Matrix cc = zero(a.m, b.n);
    for (size_t i = 0; i < a.m; i++)
      for (size_t j = 0; j < b.n; j++)
        for (size_t k = 0; k < a.n; k++) cc.at(i, j) += a.at(i, k) * b.at(k, j);
    return cc;
  }
  friend Matrix operator+(const Matrix& a, const Matrix& b) {                                         // :157-168
    if (a.m != b.m || a.n != b.n) throw Panic{SITE_OTHER, "matrix: add shape"};
    Matrix cc = zero(a.m, a.n);
    for (size_t i = 0; i < a.v.size(); i++) cc.v[i] = a.v[i] + b.v[i];
    return cc;
  }
  // &Matrix * Poly -> Poly: pad coeffs to m rows, multiply, re-normalise (Q14)                       // :147-155
  Poly<F> mul_poly(const Poly<F>& p) const {
    std::vector<F> coeffs = p.c;
    if (coeffs.size() > m) throw Panic{SITE_OTHER, "matrix*poly: too many coeffs"};
    coeffs.resize(m, F::zero());
    Matrix col{m, 1, std::move(coeffs)};
    Matrix prod = (*this) * col;
    return Poly<F>(std::move(prod.v));
  }
  friend bool operator==(const Matrix& a, const Matrix& b) { return a.m == b.m && a.n == b.n && a.v == b.v; }
};

// ---------------------------------------------------------------------------------------------
// src/fft.rs
// ---------------------------------------------------------------------------------------------
template <class F>
struct EvaluationDomainGenerator { F omega; size_t size; };                       // :6-15

// O(n^2) Vandermonde "FFT"                                                      // :23-49
template <class F>
struct VandermondeFFT {
  Matrix<F> mat;
  explicit VandermondeFFT(EvaluationDomainGenerator<F> d) : mat(Matrix<F>::zero(d.size, d.size)) {
    for (size_t n = 0; n < d.size; n++)
      for (size_t mm = 0; mm < d.size; mm++) mat.v[n * d.size + mm] = d.omega.pow((uint64_t)(n * mm));
  }
  // Q14: the output is a normalised Poly's coefficient vector, so trailing zeros are dropped
  std::vector<F> fft(const std::vector<F>& values) const { return mat.mul_poly(Poly<F>(values)).c; }
  std::vector<F> fft_inv(const std::vector<F>& freq) const {
    std::vector<F> vals = fft(freq);
    F len_inv = unwrap(F::from_u64(freq.size()).inv(), SITE_OTHER, "fft_inv: len inverse");
    std::vector<F> out;
    out.push_back(len_inv * vals[0]);
    for (size_t k = 0; k + 1 < vals.size(); k++) out.push_back(len_inv * vals[vals.size() - 1 - k]);
    return out;
  }
};

// recursive radix-2                                                             // :81-106
template <class F>
inline std::vector<F> cooley_tukey_fft(const std::vector<F>& vals, const std::vector<F>& domain) {
  if (vals.size() == 1) return {vals[0]};
  auto split = [](const std::vector<F>& v, bool even) {
    std::vector<F> o;
    for (size_t n = 0; n < v.size(); n++) if ((n % 2 == 0) == even) o.push_back(v[n]);
    return o;
  };
  std::vector<F> half_domain = split(domain, true);
  std::vector<F> l = cooley_tukey_fft(split(vals, true), half_domain);
  std::vector<F> r = cooley_tukey_fft(split(vals, false), half_domain);
  std::vector<F> o(vals.size(), F::zero());
  size_t cnt = std::min(l.size(), r.size());
  for (size_t i = 0; i < cnt; i++) {
    if (i >= domain.size()) throw Panic{SITE_OTHER, "fft: domain index"};
    F y_times_root = r[i] * domain[i];
    o[i] = l[i] + y_times_root;
    o[i + vals.size() / 2] = l[i] - y_times_root;
  }
  return o;
}

template <class F>
struct CooleyTukeyFFT {                                                           // :51-79
  std::vector<F> pows;
  explicit CooleyTukeyFFT(EvaluationDomainGenerator<F> d) {
    F mm = F::one();
    pows.push_back(mm);
    for (size_t k = 1; k < d.size; k++) { mm = mm * d.omega; pows.push_back(mm); }
  }
  std::vector<F> fft(const std::vector<F>& values) const { return cooley_tukey_fft(values, pows); }
  std::vector<F> fft_inv(const std::vector<F>& freq) const {
    std::vector<F> vals = fft(freq);
    F len_inv = unwrap(F::from_u64(freq.size()).inv(), SITE_OTHER, "fft_inv: len inverse");
    std::vector<F> out;
    out.push_back(len_inv * vals[0]);
    for (size_t k = 0; k + 1 < vals.size(); k++) out.push_back(len_inv * vals[vals.size() - 1 - k]);
    return out;
  }
};

// NTT-based polynomial product                                                   // :109-132
template <class F, class FFTI>
inline std::vector<F> mul_ntt(const FFTI& fft, std::vector<F> a_vals, std::vector<F> b_vals) {
  size_t sum = a_vals.size() + b_vals.size();
  a_vals.resize(sum, F::zero());
  b_vals.resize(sum, F::zero());
  std::vector<F> a_freq = fft.fft(a_vals), b_freq = fft.fft(b_vals);
  std::vector<F> c_freq;
  for (size_t n = 0; n < a_freq.size(); n++) {
    F l = a_freq[n];
    F r = n < b_freq.size() ? b_freq[n] : F::zero();
    c_freq.push_back(l * r);
  }
  return fft.fft_inv(c_freq);
}

// ---------------------------------------------------------------------------------------------
// src/pbh/mod.rs, src/pbh/g1.rs, g2.rs, gt.rs, pairing.rs
// ---------------------------------------------------------------------------------------------
using F101 = Fp<101>;  // src/pbh/mod.rs:8-11
using F17 = Fp<17>;    // src/pbh/mod.rs:13-16
inline F101 f101(uint64_t x) { return F101::from_u64(x); }
inline F17 f17(uint64_t x) { return F17::from_u64(x); }

// y^2 = x^3 + 3 over F_101, affine with an explicit flag                         // src/pbh/g1.rs:18-26
struct G1P {
  F101 x, y;
  bool infinite = false;
  static G1P make(F101 x, F101 y) { G1P p; p.x = x; p.y = y; p.infinite = false; return p; }   // :55-61
  static G1P generator() { return make(f101(1), f101(2)); }                                      // :71-77
  static F101 generator_subgroup_size() { return f101(17); }                                     // :79-81
  static G1P identity() { G1P p; p.x = F101::zero(); p.y = F101::zero(); p.infinite = true; return p; }  // :83-89
  bool in_curve() const { return y.pow(2) == x.pow(3) + f101(3); }    // ignores the flag (Q9)   // :63-65
  bool is_identity() const { return infinite; }
  friend bool operator==(const G1P& a, const G1P& b) { return a.x == b.x && a.y == b.y && a.infinite == b.infinite; }
  friend bool operator!=(const G1P& a, const G1P& b) { return !(a == b); }
  friend G1P operator-(const G1P& p) { return p.infinite ? p : make(p.x, -p.y); }                // :108-117
  friend G1P operator+(const G1P& self, const G1P& rhs) {                                        // :119-144 (Q16)
    if (self.infinite) return rhs;
    if (rhs.infinite) return self;
    if (self == -rhs) return identity();
    if (self == rhs) {
      F101 two = f101(2), three = f101(3);
      F101 mm = unwrap((three * self.x.pow(2)) / (two * self.y), SITE_OTHER, "g1 double");
      return make(mm * mm - two * self.x, mm * (three * self.x - mm.pow(2)) - self.y);
    }
    F101 lambda = unwrap((rhs.y - self.y) / (rhs.x - self.x), SITE_OTHER, "cannot add");
    F101 xx = lambda.pow(2) - self.x - rhs.x;
    return make(xx, lambda * (self.x - xx) - self.y);
  }
  // LSB-first double-and-add; always doubles once more after the last bit                      // :146-168
  friend G1P operator*(const G1P& self, F101 k) {
    uint64_t rhs = k.as_u64();
    if (rhs == 0 || self.is_identity()) return identity();
    std::optional<G1P> result;
    G1P base = self;
    while (rhs > 0) {
      if (rhs % 2 == 1) result = result ? (*result + base) : base;
      rhs >>= 1;
      base = base + base;
    }
    return unwrap(result, SITE_OTHER, "g1 mul");
  }
};
inline G1P g1f(uint64_t x, uint64_t y) { return G1P::make(f101(x), f101(y)); }   // src/pbh/g1.rs:11-13

// "G2": (a, b·u), u^2 = -2, no identity handling (Q12)                          // src/pbh/g2.rs:14-18
struct G2P {
  F101 a, b;
  static G2P make(F101 a, F101 b) { G2P p; p.a = a; p.b = b; return p; }
  static G2P generator() { return make(f101(36), f101(31)); }                     // :28-33
  static uint64_t embeeding_degree() { return 2; }                                // :34-36
  friend bool operator==(const G2P& p, const G2P& q) { return p.a == q.a && p.b == q.b; }
  friend G2P operator-(const G2P& p) { return make(p.a, -p.b); }                  // :51-56
  friend G2P operator+(const G2P& self, const G2P& rhs) {                         // :58-80
    if (self == rhs) {
      F101 two = f101(2), three = f101(3);
      F101 m_u = unwrap((three * self.a.pow(2)) / (two * self.b), SITE_OTHER, "g2 double");
      F101 u_pow_2_inv = unwrap((-f101(2)).inv(), SITE_OTHER, "g2 u^-2");
      F101 m_pow_2 = m_u.pow(2) * u_pow_2_inv;
      return make(m_pow_2 - two * self.a, u_pow_2_inv * m_u * (three * self.a - m_pow_2) - self.b);
    }
    F101 lambda_u = unwrap((rhs.b - self.b) / (rhs.a - self.a), SITE_OTHER, "g2 add");
    F101 lambda_pow_2 = lambda_u.pow(2) * -f101(2);
    F101 aa = lambda_pow_2 - self.a - rhs.a;
    F101 bb = lambda_u * (self.a - aa) - self.b;
    return make(aa, bb);
  }
  friend G2P operator*(const G2P& self, F101 k) {                                 // :82-101
    uint64_t rhs = k.as_u64();
    std::optional<G2P> result;
    G2P base = self;
    while (rhs > 0) {
      if (rhs % 2 == 1) result = result ? (*result + base) : base;
      rhs >>= 1;
      base = base + base;
    }
    return unwrap(result, SITE_OTHER, "g2 mul by zero");
  }
};
inline G2P g2f(uint64_t a, uint64_t b) { return G2P::make(f101(a), f101(b)); }

// F_101[u]/(u^2+2)                                                               // src/pbh/gt.rs:9-75
struct GTP {
  F101 a, b;
  static GTP make(F101 a, F101 b) { GTP g; g.a = a; g.b = b; return g; }
  friend bool operator==(const GTP& p, const GTP& q) { return p.a == q.a && p.b == q.b; }
  friend bool operator!=(const GTP& p, const GTP& q) { return !(p == q); }
  friend GTP operator-(const GTP& p) { return make(p.a, -p.b); }   // Q13: Neg is conjugation   // :21-29
  friend GTP operator*(const GTP& p, const GTP& q) {                                           // :61-69
    return make(p.a * q.a - f101(2) * p.b * q.b, p.a * q.b + p.b * q.a);
  }
  GTP pow(uint64_t n) const {                                                                   // :33-59
    GTP p, base;
    if (n >= 101) {
      p = -(this->pow(n / 101));
      n %= 101;
      base = *this;
    } else {
      p = make(F101::one(), F101::zero());
      base = *this;
    }
    while (n > 0) {
      if (n % 2 == 1) p = p * base;
      n >>= 1;
      base = base * base;
    }
    return p;
  }
};

// recursive Miller function over r                                               // src/pbh/pairing.rs:23-47
inline GTP pairing_f(uint64_t r, const G1P& p, const G2P& q) {
  auto line = [](const G1P& a, const G1P& b, F101& x, F101& y, F101& c) {
    F101 mm = b.x - a.x;
    F101 nn = b.y - a.y;
    x = nn;
    y = -mm;
    c = mm * a.y - nn * a.x;
  };
  if (r == 1) return GTP::make(f101(1), f101(0));
  F101 x, y, c;
  if (r % 2 == 1) {
    uint64_t r1 = r - 1;
    line(p * f101(r1), p, x, y, c);
    return pairing_f(r1, p, q) * GTP::make(q.a * x + c, q.b * y);
  }
  uint64_t r2 = r / 2;
  line(p * f101(r2), (-p) * f101(r2) * f101(2), x, y, c);
  return pairing_f(r2, p, q).pow(2) * GTP::make(q.a * x + c, q.b * y);
}

inline GTP pairing(const G1P& g1, const G2P& g2) {                                // src/pbh/pairing.rs:12-20
  uint64_t p = F101::order();
  uint64_t r = G1P::generator_subgroup_size().as_u64();
  uint64_t k = G2P::embeeding_degree();
  uint64_t pk = 1;
  for (uint64_t i = 0; i < k; i++) pk *= p;
  uint64_t exp = (pk - 1) / r;
  return pairing_f(r, g1, g2).pow(exp);
}

// PlonkByHandTypes                                                               // src/pbh/mod.rs:18-33
inline F17 K1() { return f17(2); }
inline F17 K2() { return f17(3); }
inline F17 OMEGA() { return f17(4); }
inline F101 gf(F17 s) { return F101::from_u64(s.as_u64()); }

// ---------------------------------------------------------------------------------------------
// src/constraints.rs (Gate, CopyOf, Constrains, Assigments, satisfies)
// ---------------------------------------------------------------------------------------------
struct Gate {                                                                     // :10-64
  F17 q_l, q_r, q_o, q_m, q_c;
  static Gate sum_a_b() { return {F17::one(), F17::one(), -F17::one(), F17::zero(), F17::zero()}; }
  static Gate sub_a_b() { return {F17::one(), F17::one(), F17::one(), F17::zero(), F17::zero()}; }
  static Gate mul_a_b() { return {F17::zero(), F17::zero(), -F17::one(), F17::one(), F17::zero()}; }
  static Gate bind_a(F17 value) { return {F17::one(), F17::zero(), F17::zero(), F17::one(), value}; }
};
struct CopyOf { char wire; size_t n; };                                           // :67-71  ('A'|'B'|'C', 1-based)

struct Assigments { std::vector<F17> a, b, c; };                                  // :132-136

struct Constrains {                                                               // :109-118
  std::vector<F17> q_l, q_r, q_o, q_m, q_c;
  std::vector<CopyOf> c_a, c_b, c_c;
  static Constrains make(const std::vector<Gate>& gates, std::vector<CopyOf> ca, std::vector<CopyOf> cb,
                         std::vector<CopyOf> cc) {                                // :139-153
    Constrains k;
    for (auto& g : gates) {
      k.q_l.push_back(g.q_l); k.q_r.push_back(g.q_r); k.q_o.push_back(g.q_o);
      k.q_m.push_back(g.q_m); k.q_c.push_back(g.q_c);
    }
    k.c_a = std::move(ca); k.c_b = std::move(cb); k.c_c = std::move(cc);
    return k;
  }
  bool satisfies(const Assigments& v) const {                                     // :198-230
    if (v.a.size() != q_l.size()) throw Panic{SITE_OTHER, "satisfies: len"};
    for (size_t n = 0; n < v.a.size(); n++) {
      // Q8: q_l is used where q_r is meant
      F17 r = q_l[n] * v.a[n] + q_l[n] * v.b[n] + q_o[n] * v.c[n] + q_m[n] * v.a[n] * v.b[n] + q_c[n];
      if (r != F17::zero()) return false;
    }
    if (v.a.size() != c_a.size() || v.a.size() != c_b.size() || v.a.size() != c_c.size())
      throw Panic{SITE_OTHER, "satisfies: copy len"};
    auto value = [&](const CopyOf& c) -> const F17& {
      const std::vector<F17>& w = c.wire == 'A' ? v.a : (c.wire == 'B' ? v.b : v.c);
      if (c.n < 1 || c.n > w.size()) throw Panic{SITE_OTHER, "satisfies: copy index"};
      return w[c.n - 1];
    };
    for (size_t n = 0; n < c_a.size(); n++) {
      if (v.a[n] != value(c_a[n]) || v.b[n] != value(c_b[n]) || v.c[n] != value(c_c[n])) return false;
    }
    return true;
  }
};

// ---------------------------------------------------------------------------------------------
// src/plonk.rs
// ---------------------------------------------------------------------------------------------
struct SRS {                                                                      // :28-32
  std::vector<G1P> g1s;
  G2P g2_1, g2_s;
  static SRS create(F101 s, size_t n) {                                           // :35-48 (Q11)
    SRS srs;
    F101 s_pow = s;
    srs.g1s.push_back(G1P::generator());
    for (size_t k = 0; k < n; k++) {
      srs.g1s.push_back(G1P::generator() * s_pow);
      s_pow = s_pow * s;
    }
    srs.g2_1 = G2P::generator();
    srs.g2_s = G2P::generator() * s;
    return srs;
  }
  G1P eval_at_s(const Poly<F17>& vs) const {                                      // :51-58
    G1P acc = G1P::identity();
    for (size_t n = 0; n < vs.c.size(); n++) {
      if (n >= g1s.size()) throw Panic{SITE_SRS_OOB, "eval_at_s: more coefficients than SRS points"};
      acc = acc + g1s[n] * gf(vs.c[n]);
    }
    return acc;
  }
};

struct Proof {                                                                    // :61-95
  G1P a_s, b_s, c_s, z_s, t_lo_s, t_mid_s, t_hi_s, w_z_s, w_z_omega_s;
  F17 a_z, b_z, c_z, s_sigma_1_z, s_sigma_2_z, r_z, z_omega_z;
  friend bool operator==(const Proof& p, const Proof& q) {
    return p.a_s == q.a_s && p.b_s == q.b_s && p.c_s == q.c_s && p.z_s == q.z_s && p.t_lo_s == q.t_lo_s &&
           p.t_mid_s == q.t_mid_s && p.t_hi_s == q.t_hi_s && p.w_z_s == q.w_z_s &&
           p.w_z_omega_s == q.w_z_omega_s && p.a_z == q.a_z && p.b_z == q.b_z && p.c_z == q.c_z &&
           p.s_sigma_1_z == q.s_sigma_1_z && p.s_sigma_2_z == q.s_sigma_2_z && p.r_z == q.r_z &&
           p.z_omega_z == q.z_omega_z;
  }
};

struct Challange { F17 alpha, beta, gamma, z, v; };                               // :97-108

// challenge source of Plonk::prove_cs that returns the caller's Challange: the reference's behaviour
struct FixedChallenges {
  Challange ch;
  void beta_gamma(const G1P&, const G1P&, const G1P&, F17& beta, F17& gamma) { beta = ch.beta; gamma = ch.gamma; }
  F17 alpha(const G1P&) { return ch.alpha; }
  F17 zeta(const G1P&, const G1P&, const G1P&) { return ch.z; }
  F17 v(const F17 (&)[7]) { return ch.v; }
  void u(const G1P&, const G1P&) {}
};

// ---- Fiat-Shamir transcript (no counterpart in the reference, which leaves the challenges to
// the caller: src/plonk.rs:201-206).  Specified in include/pbh_b200.h ("Fiat-Shamir transcript"); restated here
// independently of the product's implementation and pinned against hashlib in tests/test_fiat_shamir.py. ----
struct Sha256 {                                                                   // FIPS 180-4
  uint32_t h[8];
  uint8_t buf[64];
  uint64_t len = 0;
  Sha256() {
    static const uint32_t iv[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
    for (int i = 0; i < 8; i++) h[i] = iv[i];
  }
  static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
  static const uint32_t* round_constants() {
    // first 32 bits of the fractional parts of the cube roots of the first 64 primes, computed rather than tabulated
    static uint32_t k[64];
    static bool done = false;
    if (!done) {
      int count = 0;
      for (uint32_t p = 2; count < 64; p++) {
        bool prime = true;
        for (uint32_t d = 2; d * d <= p; d++) if (p % d == 0) { prime = false; break; }
        if (!prime) continue;
        // floor(frac(cbrt(p)) * 2^32) by integer bisection on y^3 <= p * 2^96, y = cbrt(p) * 2^32
        unsigned __int128 target = (unsigned __int128)p << 96;
        uint64_t lo = 0, hi = (uint64_t)8 << 32;
        while (hi - lo > 1) {
          uint64_t mid = lo + (hi - lo) / 2;
          // mid^3 may overflow 128 bits only if mid >= 2^42.67; mid < 2^35 here
          unsigned __int128 cube = (unsigned __int128)mid * mid * mid;
          if (cube <= target) lo = mid; else hi = mid;
        }
        k[count++] = (uint32_t)lo;
      }
      done = true;
    }
    return k;
  }
  void block(const uint8_t* p) {
    const uint32_t* K = round_constants();
    uint32_t w[64];
    for (int i = 0; i < 16; i++) w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
    for (int i = 16; i < 64; i++) {
      uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
      uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
      w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; i++) {
      uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25), ch = (e & f) ^ (~e & g);
      uint32_t t1 = hh + S1 + ch + K[i] + w[i];
      uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22), maj = (a & b) ^ (a & c) ^ (b & c);
      uint32_t t2 = S0 + maj;
      hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
  }
  void update(const uint8_t* p, size_t n) {
    for (size_t i = 0; i < n; i++) {
      buf[len % 64] = p[i];
      len++;
      if (len % 64 == 0) block(buf);
    }
  }
  void finish(uint8_t out[32]) {
    uint64_t bits = len * 8;
    uint8_t pad = 0x80;
    update(&pad, 1);
    pad = 0;
    while (len % 64 != 56) update(&pad, 1);
    uint8_t lb[8];
    for (int i = 0; i < 8; i++) lb[i] = (uint8_t)(bits >> (56 - 8 * i));
    update(lb, 8);
    for (int i = 0; i < 8; i++) { out[4 * i] = (uint8_t)(h[i] >> 24); out[4 * i + 1] = (uint8_t)(h[i] >> 16); out[4 * i + 2] = (uint8_t)(h[i] >> 8); out[4 * i + 3] = (uint8_t)h[i]; }
  }
  static void hash(const std::vector<uint8_t>& m, uint8_t out[32]) { Sha256 s; s.update(m.data(), m.size()); s.finish(out); }
};

// state_{k+1} = SHA-256(state_k || message_k); a challenge is a big-endian 64-bit slice of the new state, mod 17
struct FsTranscript {
  uint8_t state[32];
  uint8_t derived[6] = {0, 0, 0, 0, 0, 0};   // alpha beta gamma z v u, in the Challange order followed by u
  int n_derived = 0;                          // how many absorb steps ran (for tests: challenges after a panic stay 0)
  explicit FsTranscript(const uint8_t seed[32]) { for (int i = 0; i < 32; i++) state[i] = seed[i]; }
  static void put_point(std::vector<uint8_t>& m, const G1P& p) {
    m.push_back((uint8_t)p.x.v); m.push_back((uint8_t)p.y.v); m.push_back(p.infinite ? 1 : 0); m.push_back(0);
  }
  void absorb(const std::vector<uint8_t>& msg) {
    std::vector<uint8_t> m(state, state + 32);
    m.insert(m.end(), msg.begin(), msg.end());
    Sha256::hash(m, state);
    n_derived++;
  }
  uint8_t squeeze(int k) const {               // k-th 8-byte slice of the state as a big-endian integer, mod 17
    uint64_t x = 0;
    for (int i = 0; i < 8; i++) x = (x << 8) | state[8 * k + i];
    return (uint8_t)(x % 17);
  }
  void beta_gamma(const G1P& a, const G1P& b, const G1P& c, F17& beta, F17& gamma) {
    std::vector<uint8_t> m; put_point(m, a); put_point(m, b); put_point(m, c);
    absorb(m);
    derived[1] = squeeze(0); derived[2] = squeeze(1);
    beta = f17(derived[1]); gamma = f17(derived[2]);
  }
  F17 alpha(const G1P& z) {
    std::vector<uint8_t> m; put_point(m, z);
    absorb(m);
    derived[0] = squeeze(0);
    return f17(derived[0]);
  }
  F17 zeta(const G1P& lo, const G1P& mid, const G1P& hi) {
    std::vector<uint8_t> m; put_point(m, lo); put_point(m, mid); put_point(m, hi);
    absorb(m);
    derived[3] = squeeze(0);
    return f17(derived[3]);
  }
  F17 v(const F17 (&ev)[7]) {
    std::vector<uint8_t> m;
    for (int i = 0; i < 7; i++) m.push_back((uint8_t)ev[i].v);
    absorb(m);
    derived[4] = squeeze(0);
    return f17(derived[4]);
  }
  void u(const G1P& wz, const G1P& wzw) {
    std::vector<uint8_t> m; put_point(m, wz); put_point(m, wzw);
    absorb(m);
    derived[5] = squeeze(0);
  }
};

// optional trace of prover intermediates, for cross-checking SURVEY.md §9
struct ProveTrace {
  std::vector<F17> sigma_1, sigma_2, sigma_3, acc;
  Poly<F17> f_a, f_b, f_c, q_m, q_l, q_r, q_o, q_c, s1, s2, s3, l1, a, b, c, acc_x, z, z_omega, numerator, t, r,
      w_z, w_z_omega;
  F17 t_z;
};
struct VerifyTrace {
  G1P q_m_s, q_l_s, q_r_s, q_o_s, q_c_s, sigma_1_s, sigma_2_s, sigma_3_s, d_s, f_s, e_s, e_1_q1, e_2_q1;
  F17 z_h_z, l_1_z, t_z;
  GTP e_1, e_2;
  int reason = 0;  // 0 = reached the pairing check, 1 = not on curve, 2 = not in field
};

struct Plonk {                                                                    // :110-117
  SRS srs;
  std::vector<F17> h, k1_h, k2_h;
  Matrix<F17> h_pows_inv;
  Poly<F17> z_h_x;

  static bool contains(const std::vector<F17>& v, F17 x) { return std::find(v.begin(), v.end(), x) != v.end(); }

  Plonk(SRS srs_, F17 omega_pows) : srs(std::move(srs_)), h_pows_inv(Matrix<F17>::zero(1, 1)) {   // :120-175
    for (uint64_t n = 0; n < omega_pows.as_u64(); n++) h.push_back(OMEGA().pow(n));
    if (contains(h, K1()) || contains(h, K2())) throw Panic{SITE_OTHER, "K1/K2 in H"};
    for (auto r : h) k1_h.push_back(r * K1());
    if (contains(k1_h, K2())) throw Panic{SITE_OTHER, "K2 in k1H"};
    for (auto r : h) k2_h.push_back(r * K2());
    Matrix<F17> h_pows = Matrix<F17>::zero(h.size(), h.size());
    for (size_t c = 0; c < h_pows.cols(); c++)
      for (size_t r = 0; r < h_pows.rows(); r++) h_pows.at(r, c) = h[r].pow((uint64_t)c);
    h_pows_inv = h_pows.inv();
    z_h_x = Poly<F17>::z(h);
  }

  Poly<F17> interpolate_at_h(const std::vector<F17>& vv) const {                  // :177-179
    return h_pows_inv.mul_poly(Poly<F17>(vv));
  }
  std::vector<F17> copy_constraints_to_roots(const std::vector<CopyOf>& c) const {   // :181-189
    std::vector<F17> out;
    for (auto& k : c) {
      const std::vector<F17>& src = k.wire == 'A' ? h : (k.wire == 'B' ? k1_h : k2_h);
      if (k.n < 1 || k.n > src.size()) throw Panic{SITE_OTHER, "copy constraint index"};
      out.push_back(src[k.n - 1]);
    }
    return out;
  }

  // src/plonk.rs:191-466
  Proof prove(const Constrains& constraints, const Assigments& assigments, const Challange& ch,
              const F17 (&rand)[9], ProveTrace* tr = nullptr) const {
    FixedChallenges cs{ch};
    return prove_cs(constraints, assigments, cs, rand, tr);
  }

  // The same function with the five challenges asked from `cs` at the point where the reference first needs each of
  // them (the reference receives them as an argument, src/plonk.rs:195, 201-206): beta and gamma after the round-1
  // commitments, alpha after [z], z after the quotient commitments, v after the evaluations.  FixedChallenges hands
  // back the caller's Challange, which makes this the reference's prove; FsTranscript (below) derives them.
  template <class CS>
  Proof prove_cs(const Constrains& constraints, const Assigments& assigments, CS& cs, const F17 (&rand)[9],
                 ProveTrace* tr = nullptr) const {
    using P = Poly<F17>;
    if (!constraints.satisfies(assigments)) throw Panic{SITE_UNSATISFIED, "constraints not satisfied"};  // :199
    const F17 omega = OMEGA(), k1 = K1(), k2 = K2();
    const uint64_t n = constraints.c_a.size();

    std::vector<F17> sigma_1 = copy_constraints_to_roots(constraints.c_a);        // :222-224
    std::vector<F17> sigma_2 = copy_constraints_to_roots(constraints.c_b);
    std::vector<F17> sigma_3 = copy_constraints_to_roots(constraints.c_c);

    P f_a_x = interpolate_at_h(assigments.a);                                     // :233-243
    P f_b_x = interpolate_at_h(assigments.b);
    P f_c_x = interpolate_at_h(assigments.c);
    P q_o_x = interpolate_at_h(constraints.q_o);
    P q_m_x = interpolate_at_h(constraints.q_m);
    P q_l_x = interpolate_at_h(constraints.q_l);
    P q_r_x = interpolate_at_h(constraints.q_r);
    P q_c_x = interpolate_at_h(constraints.q_c);
    P s_sigma_1 = interpolate_at_h(sigma_1);
    P s_sigma_2 = interpolate_at_h(sigma_2);
    P s_sigma_3 = interpolate_at_h(sigma_3);

    // round 1                                                                    // :248-257
    F17 b1 = rand[0], b2 = rand[1], b3 = rand[2], b4 = rand[3], b5 = rand[4], b6 = rand[5];
    P a_x = P({b2, b1}) * z_h_x + f_a_x;
    P b_x = P({b4, b3}) * z_h_x + f_b_x;
    P c_x = P({b6, b5}) * z_h_x + f_c_x;
    G1P a_s = srs.eval_at_s(a_x);
    G1P b_s = srs.eval_at_s(b_x);
    G1P c_s = srs.eval_at_s(c_x);
    F17 beta, gamma;
    cs.beta_gamma(a_s, b_s, c_s, beta, gamma);

    // round 2                                                                    // :267-313
    F17 b7 = rand[6], b8 = rand[7], b9 = rand[8];
    std::vector<F17> acc{F17::one()};
    for (size_t i = 1; i < (size_t)n; i++) {
      F17 a = assigments.a[i - 1], b = assigments.b[i - 1], c = assigments.c[i - 1];
      F17 omega_pow = omega.pow((uint64_t)i - 1);
      F17 dend = (a + beta * omega_pow + gamma) * (b + beta * k1 * omega_pow + gamma) *
                 (c + beta * k2 * omega_pow + gamma);
      F17 dsor = (a + beta * s_sigma_1.eval(omega_pow) + gamma) * (b + beta * s_sigma_2.eval(omega_pow) + gamma) *
                 (c + beta * s_sigma_3.eval(omega_pow) + gamma);
      F17 vv = acc[i - 1] * unwrap(dend / dsor, SITE_ACC_DIV0, "accumulator denominator is zero");   // :297
      acc.push_back(vv);
    }
    P acc_x = interpolate_at_h(acc);
    if (!(acc_x.eval(omega.pow(n)) == F17::one())) throw Panic{SITE_ACC_ASSERT, "acc(w^n) != 1"};   // :307 (Q7)
    P z_x = P({b9, b8, b7}) * z_h_x + acc_x;
    G1P z_s = srs.eval_at_s(z_x);
    const F17 alpha = cs.alpha(z_s);

    // round 3                                                                    // :328-385
    std::vector<F17> lagrange_vector(h.size(), F17::zero());
    if (!lagrange_vector.empty()) lagrange_vector[0] = F17::one();
    P l_1_x = interpolate_at_h(lagrange_vector);
    P p_i_x = P::zero();

    P a_x_b_x_q_m_x = (a_x * b_x) * q_m_x;
    P a_x_q_l_x = a_x * q_l_x;
    P b_x_q_r_x = b_x * q_r_x;
    P c_x_q_o_x = c_x * q_o_x;
    P alpha_a_x_beta_x_gamma = (a_x + P({gamma, beta})) * alpha;
    P b_x_beta_k1_x_gamma = b_x + P({gamma, beta * k1});
    P c_x_beta_k2_x_gamma = c_x + P({gamma, beta * k2});
    std::vector<F17> zw;
    for (size_t k = 0; k < z_x.c.size(); k++) zw.push_back(z_x.c[k] * omega.pow((uint64_t)k));
    P z_omega_x = P(zw);
    P alpha_a_x_beta_s_sigma1_x_gamma = (a_x + s_sigma_1 * beta + gamma) * alpha;
    P b_x_beta_s_sigma2_x_gamma = b_x + s_sigma_2 * beta + gamma;
    P c_x_beta_s_sigma3_x_gamma = c_x + s_sigma_3 * beta + gamma;
    P alpha_2_z_x_1_l_1_x = ((z_x + P({-F17::one()})) * alpha.pow(2)) * l_1_x;

    P t_1_z_h = a_x_b_x_q_m_x + a_x_q_l_x + b_x_q_r_x + c_x_q_o_x + p_i_x + q_c_x;
    P t_2_z_h = alpha_a_x_beta_x_gamma * b_x_beta_k1_x_gamma * c_x_beta_k2_x_gamma * z_x;
    P t_3_z_h = alpha_a_x_beta_s_sigma1_x_gamma * b_x_beta_s_sigma2_x_gamma * c_x_beta_s_sigma3_x_gamma * z_omega_x;
    P t_4_z_h = alpha_2_z_x_1_l_1_x;

    P numerator = t_1_z_h + t_2_z_h - t_3_z_h + t_4_z_h;     // the `-` is Q1
    auto [t_x, rem] = poly_div(numerator, z_h_x);
    if (tr) { tr->numerator = numerator; tr->t = t_x; }
    if (rem != P::zero()) throw Panic{SITE_T_REMAINDER, "t(x) remainder is not zero"};              // :370

    if (t_x.c.size() < 18) throw Panic{SITE_T_SLICE, "t(x) has fewer than 18 coefficients"};        // :376 (Q5)
    P t_hi_x = P(std::vector<F17>(t_x.c.begin() + 12, t_x.c.begin() + 18));
    P t_mid_x = P(std::vector<F17>(t_x.c.begin() + 6, t_x.c.begin() + 12));
    P t_lo_x = P(std::vector<F17>(t_x.c.begin() + 0, t_x.c.begin() + 6));
    G1P t_hi_s = srs.eval_at_s(t_hi_x);
    G1P t_mid_s = srs.eval_at_s(t_mid_x);
    G1P t_lo_s = srs.eval_at_s(t_lo_x);
    const F17 z = cs.zeta(t_lo_s, t_mid_s, t_hi_s);

    // round 4                                                                    // :393-422
    F17 a_z = a_x.eval(z), b_z = b_x.eval(z), c_z = c_x.eval(z);
    F17 s_sigma_1_z = s_sigma_1.eval(z), s_sigma_2_z = s_sigma_2.eval(z);
    F17 t_z = t_x.eval(z);
    F17 z_omega_z = z_omega_x.eval(z);

    P a_z_b_z_q_m_x = q_m_x * a_z * b_z;
    P a_z_q_l_x = q_l_x * a_z;
    P b_z_q_r_x = q_r_x * b_z;
    P c_z_q_o_x = q_o_x * c_z;
    P r_1_x = a_z_b_z_q_m_x + a_z_q_l_x + b_z_q_r_x + c_z_q_o_x + q_c_x;
    P r_2_x = z_x * ((a_z + beta * z + gamma) * (b_z + beta * k1 * z + gamma) * (c_z + beta * k2 * z + gamma) * alpha);
    // Q2: a polynomial product with z_x, and added
    P r_3_x = (z_x * (s_sigma_3 * beta * z_omega_z)) *
              ((a_z + beta * s_sigma_1_z + gamma) * (b_z + beta * s_sigma_2_z + gamma) * alpha);
    P r_4_x = z_x * l_1_x.eval(z) * alpha.pow(2);
    P r_x = r_1_x + r_2_x + r_3_x + r_4_x;
    F17 r_z = r_x.eval(z);
    const F17 evals[7] = {a_z, b_z, c_z, s_sigma_1_z, s_sigma_2_z, r_z, z_omega_z};
    const F17 v = cs.v(evals);

    // round 5                                                                    // :430-446
    P w_num = (t_lo_x + t_mid_x * z.pow(n + 2) + t_hi_x * z.pow(2 * n + 4) - t_z) + (r_x - r_z) * v +
              (a_x - a_z) * v.pow(2) + (b_x - b_z) * v.pow(3) + (c_x - c_z) * v.pow(4) +
              (s_sigma_1 - s_sigma_1_z) * v.pow(5) + (s_sigma_2 - s_sigma_2_z) * v.pow(6);
    auto [w_z_x, rem2] = poly_div(w_num, P({-z, F17::one()}));
    if (rem2 != P::zero()) throw Panic{SITE_WZ_REMAINDER, "w_z remainder"};                          // :438
    auto [w_z_omega_x, rem3] = poly_div(z_x - z_omega_z, P({(-z) * omega, F17::one()}));
    if (rem3 != P::zero()) throw Panic{SITE_WZW_REMAINDER, "w_zw remainder"};                        // :442

    if (tr) {
      tr->sigma_1 = sigma_1; tr->sigma_2 = sigma_2; tr->sigma_3 = sigma_3; tr->acc = acc;
      tr->f_a = f_a_x; tr->f_b = f_b_x; tr->f_c = f_c_x; tr->q_m = q_m_x; tr->q_l = q_l_x; tr->q_r = q_r_x;
      tr->q_o = q_o_x; tr->q_c = q_c_x; tr->s1 = s_sigma_1; tr->s2 = s_sigma_2; tr->s3 = s_sigma_3;
      tr->l1 = l_1_x; tr->a = a_x; tr->b = b_x; tr->c = c_x; tr->acc_x = acc_x; tr->z = z_x;
      tr->z_omega = z_omega_x; tr->r = r_x; tr->w_z = w_z_x; tr->w_z_omega = w_z_omega_x; tr->t_z = t_z;
    }

    G1P w_z_s = srs.eval_at_s(w_z_x);                                             // :445 (Q2 -> SITE_SRS_OOB)
    G1P w_z_omega_s = srs.eval_at_s(w_z_omega_x);

    Proof pr;
    pr.a_s = a_s; pr.b_s = b_s; pr.c_s = c_s; pr.z_s = z_s; pr.t_lo_s = t_lo_s; pr.t_mid_s = t_mid_s;
    pr.t_hi_s = t_hi_s; pr.w_z_s = w_z_s; pr.w_z_omega_s = w_z_omega_s;
    pr.a_z = a_z; pr.b_z = b_z; pr.c_z = c_z; pr.s_sigma_1_z = s_sigma_1_z; pr.s_sigma_2_z = s_sigma_2_z;
    pr.r_z = r_z; pr.z_omega_z = z_omega_z;
    cs.u(w_z_s, w_z_omega_s);
    return pr;
  }

  // src/plonk.rs:468-650
  bool verify(const Constrains& constraints, const Proof& proof, const Challange& ch, const F17 (&rand)[1],
              VerifyTrace* tr = nullptr) const {
    const G1P &a_s = proof.a_s, &b_s = proof.b_s, &c_s = proof.c_s, &z_s = proof.z_s, &t_lo_s = proof.t_lo_s,
              &t_mid_s = proof.t_mid_s, &t_hi_s = proof.t_hi_s, &w_z_s = proof.w_z_s,
              &w_z_omega_s = proof.w_z_omega_s;
    const F17 a_z = proof.a_z, b_z = proof.b_z, c_z = proof.c_z, s_sigma_1_z = proof.s_sigma_1_z,
              s_sigma_2_z = proof.s_sigma_2_z, r_z = proof.r_z, z_omega_z = proof.z_omega_z;
    const F17 alpha = ch.alpha, beta = ch.beta, gamma = ch.gamma, z = ch.z, v = ch.v;
    const F17 omega = OMEGA(), k1 = K1(), k2 = K2();

    // verifier preprocessing (circuit-constant, recomputed per call in the reference)      // :506-517
    std::vector<F17> sigma_1 = copy_constraints_to_roots(constraints.c_a);
    std::vector<F17> sigma_2 = copy_constraints_to_roots(constraints.c_b);
    std::vector<F17> sigma_3 = copy_constraints_to_roots(constraints.c_c);
    G1P q_m_s = srs.eval_at_s(interpolate_at_h(constraints.q_m));
    G1P q_l_s = srs.eval_at_s(interpolate_at_h(constraints.q_l));
    G1P q_r_s = srs.eval_at_s(interpolate_at_h(constraints.q_r));
    G1P q_o_s = srs.eval_at_s(interpolate_at_h(constraints.q_o));
    G1P q_c_s = srs.eval_at_s(interpolate_at_h(constraints.q_c));
    G1P sigma_1_s = srs.eval_at_s(interpolate_at_h(sigma_1));
    G1P sigma_2_s = srs.eval_at_s(interpolate_at_h(sigma_2));
    G1P sigma_3_s = srs.eval_at_s(interpolate_at_h(sigma_3));
    if (tr) {
      tr->q_m_s = q_m_s; tr->q_l_s = q_l_s; tr->q_r_s = q_r_s; tr->q_o_s = q_o_s; tr->q_c_s = q_c_s;
      tr->sigma_1_s = sigma_1_s; tr->sigma_2_s = sigma_2_s; tr->sigma_3_s = sigma_3_s;
    }
    F17 u = rand[0];

    // Step 1                                                                     // :523-534 (Q9)
    if (!a_s.in_curve() || !b_s.in_curve() || !c_s.in_curve() || !z_s.in_curve() || !t_lo_s.in_curve() ||
        !t_mid_s.in_curve() || !t_hi_s.in_curve() || !w_z_s.in_curve() || !w_z_omega_s.in_curve()) {
      if (tr) tr->reason = 1;
      return false;
    }
    // Step 2                                                                     // :538-547
    if (!a_z.in_field() || !b_z.in_field() || !c_z.in_field() || !s_sigma_1_z.in_field() ||
        !s_sigma_2_z.in_field() || !r_z.in_field() || !z_omega_z.in_field()) {
      if (tr) tr->reason = 2;
      return false;
    }
    // Step 4, 5                                                                  // :553-562
    F17 z_h_z = z_h_x.eval(z);
    std::vector<F17> lagrange_vector(h.size(), F17::zero());
    if (!lagrange_vector.empty()) lagrange_vector[0] = F17::one();
    F17 l_1_z = interpolate_at_h(lagrange_vector).eval(z);
    F17 p_i_z = F17::zero();

    // Step 7 (Q3: no alpha on the permutation term; Q4: unwrap)                  // :570-579
    F17 a_z_beta_s_sigma_1_z_gamma = beta * s_sigma_1_z + gamma + a_z;
    F17 b_z_beta_s_sigma_2_z_gamma = beta * s_sigma_2_z + gamma + b_z;
    F17 c_z_gamma = c_z + gamma;
    F17 l_1_z_alpha_2 = l_1_z * alpha.pow(2);
    F17 t_z = unwrap((r_z + p_i_z - (a_z_beta_s_sigma_1_z_gamma * b_z_beta_s_sigma_2_z_gamma * c_z_gamma * z_omega_z) -
                      l_1_z_alpha_2) / z_h_z,
                     SITE_VERIFY_ZH0, "Z_H(z) is zero");

    // Step 8                                                                     // :583-610
    G1P d_1_s = q_m_s * gf(a_z * b_z * v) + q_l_s * gf(a_z * v) + q_r_s * gf(b_z * v) + q_o_s * gf(c_z * v) +
                q_c_s * gf(v);
    G1P d_2_s = z_s * gf((a_z + beta * z + gamma) * (b_z + beta * k1 * z + gamma) * (c_z + beta * k2 * z + gamma) *
                             alpha * v +
                         l_1_z * alpha.pow(2) * v + u);
    G1P d_3_s = sigma_3_s * gf((a_z + beta * s_sigma_1_z + gamma) * (b_z + beta * s_sigma_2_z + gamma) * alpha * v *
                               beta * z_omega_z);
    G1P d_s = d_1_s + d_2_s + -d_3_s;

    // Step 9                                                                     // :614-624
    uint64_t n = constraints.c_a.size();
    G1P f_s = t_lo_s + t_mid_s * gf(z.pow(n + 2)) + t_hi_s * gf(z.pow(2 * n + 4)) + d_s + a_s * gf(v.pow(2)) +
              b_s * gf(v.pow(3)) + c_s * gf(v.pow(4)) + sigma_1_s * gf(v.pow(5)) + sigma_2_s * gf(v.pow(6));

    // Step 10                                                                    // :628-637
    G1P e_s = srs.eval_at_s(Poly<F17>::from_i64({1})) *
              gf(t_z + v * r_z + v.pow(2) * a_z + v.pow(3) * b_z + v.pow(4) * c_z + v.pow(5) * s_sigma_1_z +
                 v.pow(6) * s_sigma_2_z + u * z_omega_z);

    // Step 11                                                                    // :641-649
    G1P e_1_q1 = w_z_s + w_z_omega_s * gf(u);
    G2P e_1_q2 = srs.g2_s;
    G1P e_2_q1 = w_z_s * gf(z) + w_z_omega_s * gf(u * z * omega) + f_s + -e_s;
    G2P e_2_q2 = srs.g2_1;
    GTP e_1 = pairing(e_1_q1, e_1_q2);
    GTP e_2 = pairing(e_2_q1, e_2_q2);
    if (tr) {
      tr->z_h_z = z_h_z; tr->l_1_z = l_1_z; tr->t_z = t_z; tr->d_s = d_s; tr->f_s = f_s; tr->e_s = e_s;
      tr->e_1_q1 = e_1_q1; tr->e_2_q1 = e_2_q1; tr->e_1 = e_1; tr->e_2 = e_2; tr->reason = 0;
    }
    return e_1 == e_2;
  }
};

}  // namespace pbh_oracle
