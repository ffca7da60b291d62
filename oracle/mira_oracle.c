/*
 * mira_oracle.c — CPU ORACLE (test infrastructure, never shipped; see mira_oracle.h).
 *
 * Restates, in plain C with 4x64-bit-limb Montgomery arithmetic, the algorithm the reference
 * calls at src/commitment.rs:80:
 *     best_multiexp(v, &self.ck[..v.len()]).to_affine()
 * best_multiexp / multiexp_serial are in halo2_proofs::arithmetic (un-vendored git dependency,
 * Cargo.toml:60-62); their published algorithm is restated below and anchored on the reference's
 * call site and error behaviour (src/commitment.rs:78-87).  Field and curve formulas are the
 * published ones halo2curves uses for a = 0 short-Weierstrass curves in Jacobian coordinates
 * (EFD dbl-2009-l, madd-2007-bl, add-2007-bl).
 *
 * PARITY UNPINNED at the commit() boundary (no golden vector exists upstream, reference cannot be
 * built here); pinned for the field/curve layer by the reference KATs listed in mira_oracle.h.
 */
#include "mira_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "oracle_field.h"

typedef struct {
    const field_t *base, *scalar;
    fe b;        /* curve constant, Montgomery */
    aff gen;
} curve_t;
static curve_t CURVES[2];
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

/* ------------------------------------------------------------------ curve (a = 0) */
static inline int aff_is_id(const aff *p) { return fe_is_zero(&p->x) && fe_is_zero(&p->y); }
static inline int jac_is_id(const jac *p) { return fe_is_zero(&p->z); }
static inline void jac_set_id(const field_t *f, jac *p) {
    memset(p, 0, sizeof *p);
    p->y = f->one;
}
static inline void jac_from_aff(const field_t *f, jac *r, const aff *p) {
    if (aff_is_id(p)) { jac_set_id(f, r); return; }
    r->x = p->x; r->y = p->y; r->z = f->one;
}
/* dbl-2009-l */
static void jac_dbl(const field_t *f, jac *r, const jac *p) {
    if (jac_is_id(p)) { *r = *p; return; }
    fe a, b, c, d, e, ff, t, x3, y3, z3;
    fe_sqr(f, &a, &p->x);
    fe_sqr(f, &b, &p->y);
    fe_sqr(f, &c, &b);
    fe_add(f, &d, &p->x, &b); fe_sqr(f, &d, &d); fe_sub(f, &d, &d, &a); fe_sub(f, &d, &d, &c); fe_dbl(f, &d, &d);
    fe_dbl(f, &e, &a); fe_add(f, &e, &e, &a);
    fe_sqr(f, &ff, &e);
    fe_mul(f, &z3, &p->z, &p->y); fe_dbl(f, &z3, &z3);
    fe_dbl(f, &t, &d); fe_sub(f, &x3, &ff, &t);
    fe_dbl(f, &c, &c); fe_dbl(f, &c, &c); fe_dbl(f, &c, &c);
    fe_sub(f, &t, &d, &x3); fe_mul(f, &y3, &e, &t); fe_sub(f, &y3, &y3, &c);
    r->x = x3; r->y = y3; r->z = z3;
}
/* madd-2007-bl with the exceptional cases handled */
static void jac_add_mixed(const field_t *f, jac *r, const jac *p, const aff *q) {
    if (aff_is_id(q)) { *r = *p; return; }
    if (jac_is_id(p)) { jac_from_aff(f, r, q); return; }
    fe z1z1, u2, s2, h, hh, i, j, rr, v, t, x3, y3, z3;
    fe_sqr(f, &z1z1, &p->z);
    fe_mul(f, &u2, &q->x, &z1z1);
    fe_mul(f, &s2, &q->y, &p->z); fe_mul(f, &s2, &s2, &z1z1);
    if (fe_eq(&u2, &p->x)) {
        if (fe_eq(&s2, &p->y)) { jac_dbl(f, r, p); return; }
        jac_set_id(f, r); return;
    }
    fe_sub(f, &h, &u2, &p->x);
    fe_sqr(f, &hh, &h);
    fe_dbl(f, &i, &hh); fe_dbl(f, &i, &i);
    fe_mul(f, &j, &h, &i);
    fe_sub(f, &rr, &s2, &p->y); fe_dbl(f, &rr, &rr);
    fe_mul(f, &v, &p->x, &i);
    fe_sqr(f, &x3, &rr); fe_sub(f, &x3, &x3, &j); fe_sub(f, &x3, &x3, &v); fe_sub(f, &x3, &x3, &v);
    fe_mul(f, &t, &p->y, &j); fe_dbl(f, &t, &t);
    fe_sub(f, &y3, &v, &x3); fe_mul(f, &y3, &y3, &rr); fe_sub(f, &y3, &y3, &t);
    fe_add(f, &z3, &p->z, &h); fe_sqr(f, &z3, &z3); fe_sub(f, &z3, &z3, &z1z1); fe_sub(f, &z3, &z3, &hh);
    r->x = x3; r->y = y3; r->z = z3;
}
/* add-2007-bl with the exceptional cases handled */
static void jac_add(const field_t *f, jac *r, const jac *p, const jac *q) {
    if (jac_is_id(q)) { *r = *p; return; }
    if (jac_is_id(p)) { *r = *q; return; }
    fe z1z1, z2z2, u1, u2, s1, s2, h, i, j, rr, v, t, x3, y3, z3;
    fe_sqr(f, &z1z1, &p->z);
    fe_sqr(f, &z2z2, &q->z);
    fe_mul(f, &u1, &p->x, &z2z2);
    fe_mul(f, &u2, &q->x, &z1z1);
    fe_mul(f, &s1, &p->y, &q->z); fe_mul(f, &s1, &s1, &z2z2);
    fe_mul(f, &s2, &q->y, &p->z); fe_mul(f, &s2, &s2, &z1z1);
    if (fe_eq(&u1, &u2)) {
        if (fe_eq(&s1, &s2)) { jac_dbl(f, r, p); return; }
        jac_set_id(f, r); return;
    }
    fe_sub(f, &h, &u2, &u1);
    fe_dbl(f, &i, &h); fe_sqr(f, &i, &i);
    fe_mul(f, &j, &h, &i);
    fe_sub(f, &rr, &s2, &s1); fe_dbl(f, &rr, &rr);
    fe_mul(f, &v, &u1, &i);
    fe_sqr(f, &x3, &rr); fe_sub(f, &x3, &x3, &j); fe_sub(f, &x3, &x3, &v); fe_sub(f, &x3, &x3, &v);
    fe_mul(f, &t, &s1, &j); fe_dbl(f, &t, &t);
    fe_sub(f, &y3, &v, &x3); fe_mul(f, &y3, &y3, &rr); fe_sub(f, &y3, &y3, &t);
    fe_add(f, &z3, &p->z, &q->z); fe_sqr(f, &z3, &z3); fe_sub(f, &z3, &z3, &z1z1); fe_sub(f, &z3, &z3, &z2z2);
    fe_mul(f, &z3, &z3, &h);
    r->x = x3; r->y = y3; r->z = z3;
}
/* Curve::to_affine: identity -> (0,0) */
static void jac_to_aff(const field_t *f, aff *r, const jac *p) {
    if (jac_is_id(p)) { memset(r, 0, sizeof *r); return; }
    fe zi, zi2, zi3;
    fe_inv(f, &zi, &p->z);
    fe_sqr(f, &zi2, &zi);
    fe_mul(f, &zi3, &zi2, &zi);
    fe_mul(f, &r->x, &p->x, &zi2);
    fe_mul(f, &r->y, &p->y, &zi3);
}
/* Curve::batch_normalize (Montgomery's trick) */
static void jac_batch_to_aff(const field_t *f, aff *out, const jac *in, size_t n) {
    fe *pre = (fe *)malloc(sizeof(fe) * (n ? n : 1));
    fe acc = f->one;
    for (size_t i = 0; i < n; i++) {
        pre[i] = acc;
        if (!jac_is_id(&in[i])) fe_mul(f, &acc, &acc, &in[i].z);
    }
    fe inv;
    fe_inv(f, &inv, &acc);
    for (size_t i = n; i-- > 0;) {
        if (jac_is_id(&in[i])) { memset(&out[i], 0, sizeof(aff)); continue; }
        fe zi, zi2, zi3;
        fe_mul(f, &zi, &inv, &pre[i]);
        fe_mul(f, &inv, &inv, &in[i].z);
        fe_sqr(f, &zi2, &zi);
        fe_mul(f, &zi3, &zi2, &zi);
        fe_mul(f, &out[i].x, &in[i].x, &zi2);
        fe_mul(f, &out[i].y, &in[i].y, &zi3);
    }
    free(pre);
}

static void curves_init(void) {
    field_init(&FIELDS[0]);
    field_init(&FIELDS[1]);
    /* BN254 G1: y^2 = x^3 + 3 over Fq, generator (1, 2) */
    CURVES[0].base = &FIELDS[ORACLE_FQ];
    CURVES[0].scalar = &FIELDS[ORACLE_FR];
    fe_from_u64(&FIELDS[0], &CURVES[0].b, 3);
    fe_from_u64(&FIELDS[0], &CURVES[0].gen.x, 1);
    fe_from_u64(&FIELDS[0], &CURVES[0].gen.y, 2);
    /* Grumpkin: y^2 = x^3 - 17 over Fr, generator (1, sqrt(-16)) */
    CURVES[1].base = &FIELDS[ORACLE_FR];
    CURVES[1].scalar = &FIELDS[ORACLE_FQ];
    fe t;
    fe_from_u64(&FIELDS[1], &t, 17);
    fe_neg(&FIELDS[1], &CURVES[1].b, &t);
    fe_from_u64(&FIELDS[1], &CURVES[1].gen.x, 1);
    /* canonical y = sqrt(-16) = 0x2cf135e7506a45d632d270d45f1181294833fc48d823f272c (checked on-curve in tests) */
    fe gyc = {{ 0x833fc48d823f272cULL, 0x2d270d45f1181294ULL, 0xcf135e7506a45d63ULL, 0x0000000000000002ULL }};
    fe_from_canonical(&FIELDS[1], &CURVES[1].gen.y, &gyc);
}
static inline void init(void) { pthread_once(&g_once, curves_init); }

static int on_curve(const curve_t *c, const aff *p) {
    if (aff_is_id(p)) return 1;
    fe l, r;
    fe_sqr(c->base, &l, &p->y);
    fe_sqr(c->base, &r, &p->x); fe_mul(c->base, &r, &r, &p->x); fe_add(c->base, &r, &r, &c->b);
    return fe_eq(&l, &r);
}

/* ------------------------------------------------------------------ multiexp (halo2_proofs::arithmetic) */
/* get_at(segment, c, bytes): c-bit unsigned window `segment` of the 32-byte LE canonical scalar */
static inline size_t get_at(size_t segment, size_t c, const uint8_t bytes[32]) {
    size_t skip_bits = segment * c;
    size_t skip_bytes = skip_bits / 8;
    if (skip_bytes >= 32) return 0;
    uint8_t v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    size_t avail = 32 - skip_bytes;
    memcpy(v, bytes + skip_bytes, avail < 8 ? avail : 8);
    uint64_t tmp;
    memcpy(&tmp, v, 8);
    tmp >>= skip_bits - skip_bytes * 8;
    tmp %= ((uint64_t)1 << c);
    return (size_t)tmp;
}

enum { B_NONE = 0, B_AFFINE = 1, B_PROJ = 2 };
typedef struct { int tag; aff a; jac p; } bucket_t;

static void multiexp_serial(const curve_t *cv, const fe *coeffs, const aff *bases, size_t n, jac *acc) {
    const field_t *f = cv->base;
    uint8_t (*repr)[32] = (uint8_t (*)[32])malloc(32 * (n ? n : 1));
    for (size_t i = 0; i < n; i++) {
        fe c;
        fe_to_canonical(cv->scalar, &c, &coeffs[i]);   /* a.to_repr() */
        memcpy(repr[i], c.l, 32);
    }
    size_t c;
    if (n < 4) c = 1;
    else if (n < 32) c = 3;
    else c = (size_t)ceil(log((double)(uint32_t)n));
    size_t segments = 256 / c + 1;
    size_t nb = ((size_t)1 << c) - 1;
    bucket_t *buckets = (bucket_t *)malloc(sizeof(bucket_t) * nb);
    for (size_t seg = segments; seg-- > 0;) {
        for (size_t k = 0; k < c; k++) jac_dbl(f, acc, acc);
        for (size_t b = 0; b < nb; b++) buckets[b].tag = B_NONE;
        for (size_t i = 0; i < n; i++) {
            size_t d = get_at(seg, c, repr[i]);
            if (d == 0) continue;
            bucket_t *bk = &buckets[d - 1];
            if (bk->tag == B_NONE) { bk->tag = B_AFFINE; bk->a = bases[i]; }
            else if (bk->tag == B_AFFINE) {
                jac t; jac_from_aff(f, &t, &bk->a);
                jac_add_mixed(f, &bk->p, &t, &bases[i]);
                bk->tag = B_PROJ;
            } else jac_add_mixed(f, &bk->p, &bk->p, &bases[i]);
        }
        /* summation by parts */
        jac running; jac_set_id(f, &running);
        for (size_t b = nb; b-- > 0;) {
            if (buckets[b].tag == B_AFFINE) jac_add_mixed(f, &running, &running, &buckets[b].a);
            else if (buckets[b].tag == B_PROJ) jac_add(f, &running, &running, &buckets[b].p);
            jac_add(f, acc, acc, &running);
        }
    }
    free(buckets);
    free(repr);
}

typedef struct { const curve_t *cv; const fe *coeffs; const aff *bases; size_t n; jac acc; } mx_task;
static void *mx_worker(void *arg) {
    mx_task *t = (mx_task *)arg;
    multiexp_serial(t->cv, t->coeffs, t->bases, t->n, &t->acc);
    return NULL;
}

int oracle_num_cores(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

static void best_multiexp(const curve_t *cv, const fe *coeffs, const aff *bases, size_t n, int threads, jac *out) {
    const field_t *f = cv->base;
    size_t num_threads = threads > 0 ? (size_t)threads : (size_t)oracle_num_cores();
    if (n > num_threads) {
        size_t chunk = n / num_threads;
        size_t num_chunks = (n + chunk - 1) / chunk;
        mx_task *tasks = (mx_task *)calloc(num_chunks, sizeof(mx_task));
        pthread_t *tids = (pthread_t *)calloc(num_chunks, sizeof(pthread_t));
        for (size_t k = 0; k < num_chunks; k++) {
            size_t lo = k * chunk, hi = lo + chunk < n ? lo + chunk : n;
            tasks[k].cv = cv; tasks[k].coeffs = coeffs + lo; tasks[k].bases = bases + lo; tasks[k].n = hi - lo;
            jac_set_id(f, &tasks[k].acc);
            pthread_create(&tids[k], NULL, mx_worker, &tasks[k]);
        }
        jac acc; jac_set_id(f, &acc);
        for (size_t k = 0; k < num_chunks; k++) {
            pthread_join(tids[k], NULL);
            jac_add(f, &acc, &acc, &tasks[k].acc);
        }
        *out = acc;
        free(tasks); free(tids);
    } else {
        jac acc; jac_set_id(f, &acc);
        multiexp_serial(cv, coeffs, bases, n, &acc);
        *out = acc;
    }
}

int oracle_commit(int curve, const void *bases, size_t n_bases, const void *scalars, size_t n, int threads,
                  void *out_affine64) {
    init();
    if (n > n_bases) return -1;   /* Error::TooLongInput, src/commitment.rs:82-85 */
    const curve_t *cv = &CURVES[curve];
    jac acc;
    best_multiexp(cv, (const fe *)scalars, (const aff *)bases, n, threads, &acc);
    aff r;
    jac_to_aff(cv->base, &r, &acc);
    memcpy(out_affine64, &r, 64);
    return 0;
}

static void scalar_mul_canon(const curve_t *cv, jac *r, const aff *base, const fe *canon) {
    jac acc; jac_set_id(cv->base, &acc);
    for (int i = 255; i >= 0; i--) {
        jac_dbl(cv->base, &acc, &acc);
        if ((canon->l[i / 64] >> (i % 64)) & 1) jac_add_mixed(cv->base, &acc, &acc, base);
    }
    *r = acc;
}

void oracle_commit_naive(int curve, const void *bases, const void *scalars, size_t n, void *out_affine64) {
    init();
    const curve_t *cv = &CURVES[curve];
    jac acc; jac_set_id(cv->base, &acc);
    for (size_t i = 0; i < n; i++) {
        fe c; jac t;
        fe_to_canonical(cv->scalar, &c, &((const fe *)scalars)[i]);
        scalar_mul_canon(cv, &t, &((const aff *)bases)[i], &c);
        jac_add(cv->base, &acc, &acc, &t);
    }
    aff r;
    jac_to_aff(cv->base, &r, &acc);
    memcpy(out_affine64, &r, 64);
}

/* ------------------------------------------------------------------ exported field / curve helpers */
void oracle_fe_from_u64(int field, uint64_t v, void *out) { init(); fe r; fe_from_u64(&FIELDS[field], &r, v); memcpy(out, &r, 32); }
void oracle_fe_from_canonical(int field, const void *c, void *out) { init(); fe a, r; memcpy(&a, c, 32); fe_from_canonical(&FIELDS[field], &r, &a); memcpy(out, &r, 32); }
void oracle_fe_to_canonical(int field, const void *a_, void *out) { init(); fe a, r; memcpy(&a, a_, 32); fe_to_canonical(&FIELDS[field], &r, &a); memcpy(out, &r, 32); }
void oracle_fe_add(int field, const void *a_, const void *b_, void *out) { init(); fe a, b, r; memcpy(&a, a_, 32); memcpy(&b, b_, 32); fe_add(&FIELDS[field], &r, &a, &b); memcpy(out, &r, 32); }
void oracle_fe_sub(int field, const void *a_, const void *b_, void *out) { init(); fe a, b, r; memcpy(&a, a_, 32); memcpy(&b, b_, 32); fe_sub(&FIELDS[field], &r, &a, &b); memcpy(out, &r, 32); }
void oracle_fe_mul(int field, const void *a_, const void *b_, void *out) { init(); fe a, b, r; memcpy(&a, a_, 32); memcpy(&b, b_, 32); fe_mul(&FIELDS[field], &r, &a, &b); memcpy(out, &r, 32); }
void oracle_fe_inv(int field, const void *a_, void *out) { init(); fe a, r; memcpy(&a, a_, 32); fe_inv(&FIELDS[field], &r, &a); memcpy(out, &r, 32); }
void oracle_fe_mul_many(int field, const void *a_, const void *b_, size_t n, void *out) {
    init();
    const fe *a = (const fe *)a_, *b = (const fe *)b_;
    fe *o = (fe *)out;
    for (size_t i = 0; i < n; i++) fe_mul(&FIELDS[field], &o[i], &a[i], &b[i]);
}

void oracle_generator(int curve, void *out) { init(); memcpy(out, &CURVES[curve].gen, 64); }
int oracle_is_on_curve(int curve, const void *p) { init(); aff a; memcpy(&a, p, 64); return on_curve(&CURVES[curve], &a); }
void oracle_point_add_affine(int curve, const void *p_, const void *q_, void *out) {
    init();
    const curve_t *cv = &CURVES[curve];
    aff p, q, r; jac t;
    memcpy(&p, p_, 64); memcpy(&q, q_, 64);
    jac_from_aff(cv->base, &t, &p);
    jac_add_mixed(cv->base, &t, &t, &q);
    jac_to_aff(cv->base, &r, &t);
    memcpy(out, &r, 64);
}
void oracle_point_neg_affine(int curve, const void *p_, void *out) {
    init();
    aff p; memcpy(&p, p_, 64);
    fe_neg(CURVES[curve].base, &p.y, &p.y);
    memcpy(out, &p, 64);
}
void oracle_scalar_mul(int curve, const void *base, const void *scalar, void *out) {
    init();
    const curve_t *cv = &CURVES[curve];
    aff b, r; fe s, c; jac t;
    memcpy(&b, base, 64); memcpy(&s, scalar, 32);
    fe_to_canonical(cv->scalar, &c, &s);
    scalar_mul_canon(cv, &t, &b, &c);
    jac_to_aff(cv->base, &r, &t);
    memcpy(out, &r, 64);
}

/* ------------------------------------------------------------------ synthetic inputs */
static inline uint64_t sm64_word(uint64_t seed, uint64_t k) {
    uint64_t z = seed + (k + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
/* canonical uniform-ish scalar i of stream `seed` in field f */
static void gen_canon(const field_t *f, uint64_t seed, uint64_t i, fe *out) {
    fe c;
    for (int k = 0; k < 4; k++) c.l[k] = sm64_word(seed, 4 * i + k);
    c.l[3] &= 0x3FFFFFFFFFFFFFFFULL;
    if (ge_mod(c.l, f->m)) sub4(c.l, c.l, f->m);
    *out = c;
}
void oracle_gen_scalars(int curve, uint64_t seed, size_t first, size_t n, int dist, void *out) {
    init();
    const field_t *f = CURVES[curve].scalar;
    fe *o = (fe *)out;
    for (size_t k = 0; k < n; k++) {
        uint64_t i = first + k;
        fe c;
        gen_canon(f, seed, i, &c);
        if (dist == 1) {
            uint64_t sel = sm64_word(seed ^ 0x5EEDULL, i) % 100;
            if (sel < 60) memset(&c, 0, sizeof c);
            else if (sel < 85) { c.l[0] &= 1; c.l[1] = c.l[2] = c.l[3] = 0; }
            else if (sel < 95) { c.l[0] &= 0xFFFFFFFFULL; c.l[1] = c.l[2] = c.l[3] = 0; }
        }
        fe_from_canonical(f, &o[k], &c);
    }
}

/* fixed-base table for the generator: T[w][d] = d * 2^(8w) * G, w < 32, d < 256 */
typedef struct { const curve_t *cv; uint64_t seed; size_t first, n; aff *out; const aff *table; } gb_task;
static void *gb_worker(void *arg) {
    gb_task *t = (gb_task *)arg;
    const curve_t *cv = t->cv;
    const size_t CH = 1024;
    jac *tmp = (jac *)malloc(sizeof(jac) * CH);
    for (size_t s = 0; s < t->n; s += CH) {
        size_t m = t->n - s < CH ? t->n - s : CH;
        for (size_t k = 0; k < m; k++) {
            fe c;
            gen_canon(cv->scalar, t->seed, t->first + s + k, &c);
            jac acc; jac_set_id(cv->base, &acc);
            for (int w = 0; w < 32; w++) {
                unsigned d = (unsigned)((c.l[w / 8] >> (8 * (w % 8))) & 0xFF);
                if (d) jac_add_mixed(cv->base, &acc, &acc, &t->table[w * 256 + d]);
            }
            tmp[k] = acc;
        }
        jac_batch_to_aff(cv->base, t->out + s, tmp, m);
    }
    free(tmp);
    return NULL;
}
void oracle_gen_bases(int curve, uint64_t seed, size_t first, size_t n, int threads, void *out) {
    init();
    const curve_t *cv = &CURVES[curve];
    aff *table = (aff *)calloc(32 * 256, sizeof(aff));
    jac *row = (jac *)malloc(sizeof(jac) * 256);
    jac base; jac_from_aff(cv->base, &base, &cv->gen);
    for (int w = 0; w < 32; w++) {
        aff ba; jac_to_aff(cv->base, &ba, &base);
        jac_set_id(cv->base, &row[0]);
        for (int d = 1; d < 256; d++) jac_add_mixed(cv->base, &row[d], &row[d - 1], &ba);
        jac_batch_to_aff(cv->base, table + w * 256, row, 256);
        for (int k = 0; k < 8; k++) jac_dbl(cv->base, &base, &base);
    }
    free(row);
    size_t nt = threads > 0 ? (size_t)threads : (size_t)oracle_num_cores();
    if (nt > n) nt = n ? n : 1;
    gb_task *tasks = (gb_task *)calloc(nt, sizeof(gb_task));
    pthread_t *tids = (pthread_t *)calloc(nt, sizeof(pthread_t));
    size_t per = (n + nt - 1) / nt;
    for (size_t k = 0; k < nt; k++) {
        size_t lo = k * per, hi = lo + per < n ? lo + per : n;
        if (lo > n) lo = n;
        tasks[k].cv = cv; tasks[k].seed = seed; tasks[k].first = first + lo; tasks[k].n = hi > lo ? hi - lo : 0;
        tasks[k].out = (aff *)out + lo; tasks[k].table = table;
        pthread_create(&tids[k], NULL, gb_worker, &tasks[k]);
    }
    for (size_t k = 0; k < nt; k++) pthread_join(tids[k], NULL);
    free(tasks); free(tids); free(table);
}
