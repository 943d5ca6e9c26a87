// pbh_sha256.cuh — SHA-256 (FIPS 180-4) for the Fiat-Shamir transcript of include/pbh_b200.h, usable from host and device.
// Every transcript step hashes one block: the 32-byte state followed by at most 12 message bytes, so a step is one
// compression from the initial hash value.  The general multi-block routine is host-only (context seed).
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifndef PBH_HD
#ifdef __CUDACC__
#define PBH_HD __host__ __device__ __forceinline__
#else
#define PBH_HD inline
#endif
#endif

namespace pbh {

PBH_HD uint32_t sha_rotr(uint32_t x, int n) {
#ifdef __CUDA_ARCH__
  return __funnelshift_r(x, x, n);
#else
  return (x >> n) | (x << (32 - n));
#endif
}
// round constants: first 32 bits of the fractional parts of the cube roots of the first 64 primes
PBH_HD constexpr uint32_t sha256_k_table(int i) {
  constexpr uint32_t K[64] = {
    0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u,
    0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u,
    0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau,
    0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u,
    0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
    0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u,
    0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u,
    0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};
  return K[i];
}

// one compression of the 16-word block `w` (modified in place by the message schedule) from the state `h`
PBH_HD void sha256_compress(uint32_t (&h)[8], uint32_t (&w)[16]) {
  uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
  for (int i = 0; i < 64; i++) {
    if (i >= 16) {
      const uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
      const uint32_t s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
      const uint32_t s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
      w[i & 15] = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
    }
    const uint32_t S1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25);
    const uint32_t ch = (e & f) ^ (~e & g);
    const uint32_t t1 = hh + S1 + ch + sha256_k_table(i) + w[i & 15];
    const uint32_t S0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22);
    const uint32_t maj = (a & b) ^ (a & c) ^ (b & c);
    const uint32_t t2 = S0 + maj;
    hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

PBH_HD void sha256_init(uint32_t (&h)[8]) {
  h[0] = 0x6a09e667u; h[1] = 0xbb67ae85u; h[2] = 0x3c6ef372u; h[3] = 0xa54ff53au;
  h[4] = 0x510e527fu; h[5] = 0x9b05688cu; h[6] = 0x1f83d9abu; h[7] = 0x5be0cd19u;
}

// One transcript step: SHA-256(state || message) for a message of at most 15 bytes whose padded words w8..w11 and
// length word w15 the caller has prepared.  Two copies: inlined, and an out-of-line device function.  Five inlined
// compressions make a kernel of about 190 KB of SASS that stalls on instruction fetch (ncu: 0.9-1.8 "no instruction"
// stalls per issue); one shared copy avoids that but pays the call's register shuffling.  Measured per 2^20 items:
// the verifier gains (411 -> 393 us); the prover, with many more live registers around the early calls, loses when
// every step is outlined (379 -> 414 us without the u step) and gains when only the last two are (490 -> 441 us).
struct ShaState { uint32_t v[8]; };
PBH_HD ShaState sha256_step_inline(ShaState st, uint32_t w8, uint32_t w9, uint32_t w10, uint32_t w11, uint32_t w15) {
  uint32_t w[16];
#pragma unroll
  for (int i = 0; i < 8; i++) w[i] = st.v[i];
  w[8] = w8; w[9] = w9; w[10] = w10; w[11] = w11; w[12] = 0u; w[13] = 0u; w[14] = 0u; w[15] = w15;
  uint32_t h[8];
  sha256_init(h);
  sha256_compress(h, w);
  ShaState out;
#pragma unroll
  for (int i = 0; i < 8; i++) out.v[i] = h[i];
  return out;
}
#ifdef __CUDACC__
__host__ __device__ __noinline__
#else
inline
#endif
ShaState sha256_step_outline(ShaState st, uint32_t w8, uint32_t w9, uint32_t w10, uint32_t w11, uint32_t w15) {
  return sha256_step_inline(st, w8, w9, w10, w11, w15);
}

// state <- SHA-256(state || message), the message being `nbytes` (<= 15, a constant at every call site) bytes packed
// big-endian into m[0..3] with zero bits after them (the padding bit and the length are added here)
template <bool OUTLINE = false>
PBH_HD void sha256_absorb(uint32_t (&state)[8], const uint32_t (&m)[6], int nbytes) {
  uint32_t w[4] = {m[0], m[1], m[2], m[3]};
  w[nbytes / 4] |= 0x80000000u >> (8 * (nbytes % 4));
  ShaState st;
#pragma unroll
  for (int i = 0; i < 8; i++) st.v[i] = state[i];
  st = OUTLINE ? sha256_step_outline(st, w[0], w[1], w[2], w[3], (uint32_t)(8 * (32 + nbytes)))
               : sha256_step_inline(st, w[0], w[1], w[2], w[3], (uint32_t)(8 * (32 + nbytes)));
#pragma unroll
  for (int i = 0; i < 8; i++) state[i] = st.v[i];
}

// k-th challenge of a state: its k-th big-endian 64-bit slice reduced mod 17 (2^32 = 1 mod 17)
PBH_HD uint32_t sha256_squeeze17(const uint32_t (&state)[8], int k) {
  return (uint32_t)(((uint64_t)state[2 * k] + (uint64_t)state[2 * k + 1]) % 17u);
}

// general message, host only
inline void sha256_host(const uint8_t* data, size_t len, uint32_t (&out)[8]) {
  sha256_init(out);
  uint8_t block[64];
  size_t off = 0;
  auto run = [&](const uint8_t* p) {
    uint32_t w[16];
    for (int i = 0; i < 16; i++) w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
    sha256_compress(out, w);
  };
  for (; off + 64 <= len; off += 64) run(data + off);
  size_t rem = len - off;
  for (size_t i = 0; i < 64; i++) block[i] = i < rem ? data[off + i] : 0;
  block[rem] = 0x80;
  if (rem >= 56) {
    run(block);
    for (size_t i = 0; i < 64; i++) block[i] = 0;
  }
  const uint64_t bits = (uint64_t)len * 8;
  for (int i = 0; i < 8; i++) block[56 + i] = (uint8_t)(bits >> (56 - 8 * i));
  run(block);
}

}  // namespace pbh
