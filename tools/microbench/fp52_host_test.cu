// Host-side driver of mira_b200/csrc/fp52.cuh (the FP64-pipe field arithmetic): reads operations on stdin, prints the
// results, so that tests/test_fp52_host.py can compare them bit for bit with Python big integers on a box without a GPU.
// Line format:  <field: q|r> <op> <hex operands...>   (operands: 64 hex digits, canonical integers < modulus, taken as
// std-Montgomery residues);  output: one line of 64 hex digits (std-Montgomery result) plus the maximal limb magnitude
// seen on the fp52 side, as log2.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include "fp52.cuh"

using namespace mira;
using namespace mira::fp52;

template <class F> static Fe<F> parse(const char* s) {
  Fe<F> r;
  for (int i = 0; i < 8; i++) {
    char buf[9];
    memcpy(buf, s + (7 - i) * 8, 8);
    buf[8] = 0;
    r.v[i] = (uint32_t)strtoul(buf, nullptr, 16);
  }
  return r;
}
template <class F> static void print(const Fe<F>& a, double maxlimb) {
  for (int i = 7; i >= 0; i--) printf("%08x", a.v[i]);
  printf(" %.3f\n", maxlimb > 0 ? std::log2(maxlimb) : 0.0);
}
template <class F> static double maxl(const Fd<F>& a) {
  double m = 0;
  for (int i = 0; i < 4; i++) m = std::fmax(m, std::fabs(a.v[i]));
  return m;
}

template <class F> static int run(const char* op, char** t, int nt) {
  Fd<F> a[4];
  for (int i = 0; i < nt && i < 4; i++) a[i] = fd_from_std(parse<F>(t[i]));
  Fd<F> r;
  if (!strcmp(op, "mul")) r = fd_mul(a[0], a[1]);
  else if (!strcmp(op, "sqr")) r = fd_sqr(a[0]);
  else if (!strcmp(op, "dual")) r = fd_mul_add_mul(a[0], a[1], a[2], a[3]);
  else if (!strcmp(op, "conv")) r = a[0];
  else if (!strcmp(op, "submul")) r = fd_mul(fd_norm(fd_sub(a[0], a[1])), a[2]);                // (a - b) * c
  else if (!strcmp(op, "lazy")) r = fd_mul(fd_sub(a[0], a[1]), a[2]);                           // one un-normalised operand
  else if (!strcmp(op, "lazy3")) r = fd_mul(fd_norm(fd_sub(fd_sub(a[0], a[1]), fd_add(a[2], a[2]))), a[3]);   // (a - b - 2c) * d
  else if (!strcmp(op, "chain")) {                                                              // ((a*b)^2 - c) * d, squared
    Fd<F> x = fd_mul(a[0], a[1]);
    x = fd_sqr(x);
    x = fd_norm(fd_sub(x, a[2]));
    x = fd_mul(x, a[3]);
    r = fd_sqr(x);
  } else return 1;
  print<F>(fd_to_std(r), maxl(r));
  return 0;
}

int main() {
  char line[2048];
  while (fgets(line, sizeof line, stdin)) {
    char* tok[8];
    int n = 0;
    for (char* p = strtok(line, " \n"); p && n < 8; p = strtok(nullptr, " \n")) tok[n++] = p;
    if (n < 3) continue;
    int rc = tok[0][0] == 'q' ? run<FqTag>(tok[1], tok + 2, n - 2) : run<FrTag>(tok[1], tok + 2, n - 2);
    if (rc) { fprintf(stderr, "bad op %s\n", tok[1]); return 1; }
  }
  return 0;
}
