// fr_host.h -- host-side BN254 Fr helpers used by the planner to build the Montgomery constant tables
// and the Barrett constants the kernels use.  Everything is derived from p at start-up; nothing is a
// trusted magic number except p itself (halo2curves bn256::Fr modulus).
#pragma once
#include <stdint.h>
#include <string.h>

namespace h2sha {

struct U256 {
  uint64_t l[4];
  bool operator==(const U256& o) const { return memcmp(l, o.l, 32) == 0; }
};

namespace fr {
typedef unsigned __int128 u128;
static const uint64_t P[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};

inline bool geq_p(const uint64_t a[4]) {
  for (int i = 3; i >= 0; i--) {
    if (a[i] != P[i]) return a[i] > P[i];
  }
  return true;
}
inline void sub_p(uint64_t a[4]) {
  uint64_t borrow = 0;
  for (int i = 0; i < 4; i++) {
    u128 t = (u128)a[i] - P[i] - borrow;
    a[i] = (uint64_t)t;
    borrow = (uint64_t)(t >> 64) & 1;
  }
}
// (a + b) mod p for a, b < p
inline U256 add(const U256& a, const U256& b) {
  U256 r;
  u128 c = 0;
  for (int i = 0; i < 4; i++) {
    c += (u128)a.l[i] + b.l[i];
    r.l[i] = (uint64_t)c;
    c >>= 64;
  }
  if (geq_p(r.l)) sub_p(r.l);
  return r;
}
inline bool is_zero(const U256& a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0; }
inline U256 neg(const U256& a) {
  if (is_zero(a)) return a;
  U256 r;
  uint64_t borrow = 0;
  for (int i = 0; i < 4; i++) {
    u128 t = (u128)P[i] - a.l[i] - borrow;
    r.l[i] = (uint64_t)t;
    borrow = (uint64_t)(t >> 64) & 1;
  }
  return r;
}
inline U256 dbl(const U256& a) { return add(a, a); }
// (a * b) mod p by double-and-add (slow but obviously right; only used at plan time for < 2000 values)
inline U256 mul(const U256& a, const U256& b) {
  U256 r = {{0, 0, 0, 0}};
  for (int i = 255; i >= 0; i--) {
    r = dbl(r);
    if ((b.l[i / 64] >> (i % 64)) & 1) r = add(r, a);
  }
  return r;
}
inline U256 from_u64(uint64_t v) { return U256{{v, 0, 0, 0}}; }
inline U256 pow(const U256& a, const uint64_t e[4]) {
  U256 r = from_u64(1);
  for (int i = 255; i >= 0; i--) {
    r = mul(r, r);
    if ((e[i / 64] >> (i % 64)) & 1) r = mul(r, a);
  }
  return r;
}
inline U256 inv(const U256& a) {
  uint64_t e[4] = {P[0] - 2, P[1], P[2], P[3]};
  return pow(a, e);
}
// R = 2^256 mod p
inline U256 mont_r() {
  U256 r = from_u64(1);
  for (int i = 0; i < 256; i++) r = dbl(r);
  return r;
}
// canonical -> Montgomery
inline U256 to_mont(const U256& a) {
  static const U256 R = mont_r();
  return mul(a, R);
}
// Montgomery product a * b * 2^-256 mod p (CIOS, 64-bit limbs) for a, b < p; -p^-1 mod 2^64 by Newton iteration
inline uint64_t neg_p_inv64() {
  uint64_t inv = 1;
  for (int i = 0; i < 6; i++) inv *= 2 - P[0] * inv;
  return 0 - inv;
}
inline U256 mont_mul(const U256& a, const U256& b) {
  static const uint64_t ninv = neg_p_inv64();
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) {
      c += (u128)a.l[j] * b.l[i] + t[j];
      t[j] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[4] = (uint64_t)c;
    t[5] = (uint64_t)(c >> 64);
    const uint64_t m = t[0] * ninv;
    c = (u128)m * P[0] + t[0];
    c >>= 64;
    for (int j = 1; j < 4; j++) {
      c += (u128)m * P[j] + t[j];
      t[j - 1] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[3] = (uint64_t)c;
    t[4] = t[5] + (uint64_t)(c >> 64);
  }
  U256 r = {{t[0], t[1], t[2], t[3]}};
  if (t[4] != 0 || geq_p(r.l)) sub_p(r.l);
  return r;
}
// canonical value of a signed small integer
inline U256 from_i64(int64_t v) { return v >= 0 ? from_u64((uint64_t)v) : neg(from_u64((uint64_t)(-v))); }

// floor(2^k / p) for k <= 318, as a u64 (caller guarantees it fits)
inline uint64_t floor_pow2_div_p(int k) {
  // long division of 2^k by p, bit by bit
  uint64_t rem[5] = {0, 0, 0, 0, 0};
  uint64_t q = 0;
  for (int i = k; i >= 0; i--) {
    // rem = rem * 2 + bit(i of 2^k)
    uint64_t carry = (i == k) ? 1 : 0;
    for (int j = 0; j < 5; j++) {
      uint64_t n = (rem[j] << 1) | carry;
      carry = rem[j] >> 63;
      rem[j] = n;
    }
    bool ge = rem[4] != 0 || geq_p(rem);
    q <<= 1;
    if (ge) {
      uint64_t borrow = 0;
      for (int j = 0; j < 4; j++) {
        u128 t = (u128)rem[j] - P[j] - borrow;
        rem[j] = (uint64_t)t;
        borrow = (uint64_t)(t >> 64) & 1;
      }
      rem[4] -= borrow;
      q |= 1;
    }
  }
  return q;
}
}  // namespace fr
}  // namespace h2sha
