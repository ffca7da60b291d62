/*
 * oracle_field.h — 4x64-limb Montgomery Fq/Fr shared by the oracle translation units.
 * TEST INFRASTRUCTURE ONLY (see mira_oracle.h).  Everything is static: each TU initialises its own copy.
 */
#ifndef ORACLE_FIELD_H
#define ORACLE_FIELD_H
#include <stdint.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe;
typedef struct { fe x, y; } aff;          /* (0,0) = identity */
typedef struct { fe x, y, z; } jac;       /* z = 0 identity   */

typedef struct {
    uint64_t m[4];   /* modulus */
    uint64_t inv;    /* -m^{-1} mod 2^64 */
    fe one;          /* R mod m */
    fe r2;           /* R^2 mod m */
} field_t;

/* BN254 base field Fq: p = 21888242871839275222246405745257275088696311157297823662689037894645226208583 */
/* BN254 scalar field Fr: r = 21888242871839275222246405745257275088548364400416034343698204186575808495617 */
static field_t FIELDS[2] = {
    { { 0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL }, 0, {{0}}, {{0}} },
    { { 0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL }, 0, {{0}}, {{0}} },
};

/* ------------------------------------------------------------------ field */
static inline int ge_mod(const uint64_t a[4], const uint64_t m[4]) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] > m[i]) return 1;
        if (a[i] < m[i]) return 0;
    }
    return 1;
}
static inline uint64_t sub4(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) {
    u128 br = 0;
    for (int i = 0; i < 4; i++) {
        u128 t = (u128)a[i] - b[i] - (uint64_t)br;
        r[i] = (uint64_t)t;
        br = (t >> 64) & 1;
    }
    return (uint64_t)br;
}
static inline uint64_t add4(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (u128)a[i] + b[i];
        r[i] = (uint64_t)c;
        c >>= 64;
    }
    return (uint64_t)c;
}
static inline int fe_is_zero(const fe *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int fe_eq(const fe *a, const fe *b) {
    return ((a->l[0] ^ b->l[0]) | (a->l[1] ^ b->l[1]) | (a->l[2] ^ b->l[2]) | (a->l[3] ^ b->l[3])) == 0;
}
static inline void fe_add(const field_t *f, fe *r, const fe *a, const fe *b) {
    uint64_t t[4];
    uint64_t c = add4(t, a->l, b->l);
    if (c || ge_mod(t, f->m)) sub4(t, t, f->m);
    memcpy(r->l, t, 32);
}
static inline void fe_sub(const field_t *f, fe *r, const fe *a, const fe *b) {
    uint64_t t[4];
    if (sub4(t, a->l, b->l)) add4(t, t, f->m);
    memcpy(r->l, t, 32);
}
static inline void fe_neg(const field_t *f, fe *r, const fe *a) {
    if (fe_is_zero(a)) { *r = *a; return; }
    sub4(r->l, f->m, a->l);
}
static inline void fe_dbl(const field_t *f, fe *r, const fe *a) { fe_add(f, r, a, a); }

/* CIOS Montgomery product, 4 x 64-bit limbs: r = a*b*R^{-1} mod m */
static inline void fe_mul(const field_t *f, fe *r, const fe *a, const fe *b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a->l[j] * b->l[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t q = t[0] * f->inv;
        c = (u128)q * f->m[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)q * f->m[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    if (t[4] || ge_mod(t, f->m)) sub4(t, t, f->m);
    memcpy(r->l, t, 32);
}
static inline void fe_sqr(const field_t *f, fe *r, const fe *a) { fe_mul(f, r, a, a); }

static void fe_pow(const field_t *f, fe *r, const fe *a, const uint64_t e[4]) {
    fe acc = f->one;
    for (int i = 255; i >= 0; i--) {
        fe_sqr(f, &acc, &acc);
        if ((e[i / 64] >> (i % 64)) & 1) fe_mul(f, &acc, &acc, a);
    }
    *r = acc;
}
static void fe_inv(const field_t *f, fe *r, const fe *a) {  /* Fermat; 0 -> 0 */
    uint64_t e[4], two[4] = {2, 0, 0, 0};
    sub4(e, f->m, two);
    fe_pow(f, r, a, e);
}
static void fe_from_canonical(const field_t *f, fe *r, const fe *c) { fe_mul(f, r, c, &f->r2); }
static void fe_to_canonical(const field_t *f, fe *r, const fe *a) {
    fe one_raw = {{1, 0, 0, 0}};
    fe_mul(f, r, a, &one_raw);
}
static void fe_from_u64(const field_t *f, fe *r, uint64_t v) {
    fe c = {{v, 0, 0, 0}};
    fe_from_canonical(f, r, &c);
}

static void field_init(field_t *f) {
    /* inv = -m^{-1} mod 2^64 by Newton iteration */
    uint64_t x = 1;
    for (int i = 0; i < 6; i++) x *= 2 - f->m[0] * x;
    f->inv = (uint64_t)0 - x;
    /* R mod m and R^2 mod m by repeated doubling of 1 */
    fe v = {{1, 0, 0, 0}};
    for (int i = 0; i < 512; i++) {
        uint64_t t[4];
        uint64_t c = add4(t, v.l, v.l);
        if (c || ge_mod(t, f->m)) sub4(t, t, f->m);
        memcpy(v.l, t, 32);
        if (i == 255) f->one = v;
    }
    f->r2 = v;
}

static inline void fields_init(void) { field_init(&FIELDS[0]); field_init(&FIELDS[1]); }
#endif
