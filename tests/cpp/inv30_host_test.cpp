// Host build of mira_b200/csrc/inv30.cuh (the branch-free safegcd inversion the device code uses where every lane
// inverts at once).  Reads "field hex64" lines on stdin (field 0 = BN254 Fq, 1 = Fr; the value as 64 hex digits, < M)
// and prints the inverse (0 for 0).  tests/test_inv30.py checks the output against Python's pow(x, -1, M).
#include <cstdio>
#include <cstdlib>
#include "../../mira_b200/csrc/inv30.cuh"

static constexpr uint32_t FQ[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
static constexpr uint32_t FR[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};

template <const uint32_t (&W)[8]>
struct Mod {
  static constexpr uint32_t word(int i) { return W[i]; }
  static constexpr int32_t limb(int i) {
    const int bit = 30 * i, k = bit >> 5, sh = bit & 31;
    const uint64_t lo = W[k], hi = k + 1 < 8 ? W[k + 1] : 0u;
    return (int32_t)((((hi << 32) | lo) >> sh) & 0x3fffffffu);
  }
  static constexpr uint32_t minv30() {
    uint32_t x = W[0];
    for (int i = 0; i < 5; i++) x *= 2u - W[0] * x;
    return x & 0x3fffffffu;
  }
};

int main() {
  char buf[256];
  int fld;
  while (scanf("%d %255s", &fld, buf) == 2) {
    uint32_t a[8], r[8];
    for (int i = 0; i < 8; i++) {
      char t[9];
      for (int k = 0; k < 8; k++) t[k] = buf[(7 - i) * 8 + k];
      t[8] = 0;
      a[i] = (uint32_t)strtoul(t, nullptr, 16);
    }
    if (fld == 0) mira::inv30::modinv<Mod<FQ>>(r, a);
    else mira::inv30::modinv<Mod<FR>>(r, a);
    for (int i = 7; i >= 0; i--) printf("%08x", r[i]);
    printf("\n");
  }
  return 0;
}
