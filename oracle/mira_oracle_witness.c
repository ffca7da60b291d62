/*
 * mira_oracle_witness.c — CPU ORACLE for the witness side of the hot path.
 * TEST INFRASTRUCTURE ONLY (see mira_oracle.h): never linked, imported or called by the product.
 *
 * Restates, with the same 4x64-limb Montgomery field layer as mira_oracle.c:
 *   oracle_eval_rows   GraphEvaluator::evaluate + Calculation::evaluate
 *                      (/root/reference/src/polynomial/graph_evaluator.rs:93-149, 361-388) over a
 *                      PlonkEvalDomain (src/plonk/eval.rs:93-106, 153-228) with GetDataForEval's
 *                      eval_column_var dispatch (src/plonk/eval.rs:57-70) and get_rotation_idx
 *                      (graph_evaluator.rs:51-53);
 *   oracle_fold_w/_e   RelaxedPlonkWitness::fold (src/plonk/mod.rs:1097-1134);
 *   oracle_concat_pad  util::concatenate_with_padding (src/util.rs:189-193);
 *   oracle_fft         fft::best_fft (src/fft.rs:51-115: bit-reversal + radix-2 DIT butterflies) and
 *                      get_omega_or_inv / get_ifft_divisor / ifft (src/fft.rs:12-27, 165-175).
 *
 * PINNED by: the 8-point FFT known-answer test (src/fft.rs:239-258, tests/golden/fft_kat_fr.json), the
 * fft->ifft round trip (src/fft.rs:266-279), and tests that restate the reference's own GraphEvaluator tests
 * (graph_evaluator.rs:447-634) against direct big-integer evaluation.  The program encoding is documented in
 * include/mira_b200.h (mira_eval_program_create).
 */
#include <pthread.h>
#include <stdlib.h>

#include "mira_oracle.h"
#include "oracle_field.h"

static pthread_once_t w_once = PTHREAD_ONCE_INIT;
static inline void winit(void) { pthread_once(&w_once, fields_init); }

/* ------------------------------------------------------------------ fold (src/plonk/mod.rs:1097-1134) */
void oracle_fold_w(int field, const void *w1_, const void *w2_, size_t n, const void *r_, void *out_) {
    winit();
    const field_t *f = &FIELDS[field];
    const fe *w1 = w1_, *w2 = w2_;
    fe *out = out_;
    fe r;
    memcpy(&r, r_, 32);
    for (size_t i = 0; i < n; i++) {      /* *w1 + *r * *w2 */
        fe t;
        fe_mul(f, &t, &r, &w2[i]);
        fe_add(f, &out[i], &w1[i], &t);
    }
}

void oracle_fold_e(int field, const void *e_, const void *const *terms, size_t n_terms, size_t n, const void *r_,
                   void *out_) {
    winit();
    const field_t *f = &FIELDS[field];
    const fe *e = e_;
    fe *out = out_;
    fe r;
    memcpy(&r, r_, 32);
    /* powers_or_r = r^1, r^2, ... (iter::successors) */
    fe *pw = malloc(sizeof(fe) * (n_terms ? n_terms : 1));
    for (size_t k = 0; k < n_terms; k++) {
        if (k == 0) pw[0] = r;
        else fe_mul(f, &pw[k], &pw[k - 1], &r);
    }
    for (size_t i = 0; i < n; i++) {      /* fold(*ei, |acc, (tk, p)| acc + p * tk[i]) */
        fe acc = e[i];
        for (size_t k = 0; k < n_terms; k++) {
            fe t;
            fe_mul(f, &t, &pw[k], &((const fe *)terms[k])[i]);
            fe_add(f, &acc, &acc, &t);
        }
        out[i] = acc;
    }
    free(pw);
}

/* src/util.rs:189-193: each column followed by zero padding up to pad_size (a longer column is kept whole).
 * returns the number of elements written; out may be NULL to query the length. */
size_t oracle_concat_pad(const void *const *cols, const size_t *lens, size_t n_cols, size_t pad_size, void *out_) {
    fe *out = out_;
    size_t w = 0;
    for (size_t c = 0; c < n_cols; c++) {
        size_t l = lens[c], tot = l > pad_size ? l : pad_size;
        if (out) {
            memcpy(&out[w], cols[c], l * 32);
            memset(&out[w + l], 0, (tot - l) * 32);
        }
        w += tot;
    }
    return w;
}

/* ------------------------------------------------------------------ row evaluator */
enum { OP_ADD = 0, OP_SUB, OP_MUL, OP_SQUARE, OP_DOUBLE, OP_NEGATE, OP_HORNER, OP_STORE };
enum { VS_CONSTANT = 0, VS_INTERMEDIATE, VS_FIXED, VS_POLY, VS_CHALLENGE };

typedef struct {
    const oracle_eval_domain *d;
    const field_t *f;
    const fe *constants;
    size_t n_constants;
    const fe *inter;
    uint32_t n_inter;
    const size_t *rot; /* resolved row per rotation index */
    uint32_t n_rot;
} ectx;

/* PlonkEvalDomain::eval_advice_var (src/plonk/eval.rs:153-228) */
static int advice_var(const ectx *c, size_t row, size_t index, fe *out) {
    const oracle_eval_domain *d = c->d;
    if (d->flags & ORACLE_EVAL_LOOKUP_DOMAIN) {   /* LookupEvalDomain::eval_advice_var (src/plonk/eval.rs:125-135) */
        if (index >= d->num_w1) return ORACLE_EVAL_COLUMN_OUT_OF_BOUNDARY;
        if (row >= d->w1_len[index]) return ORACLE_EVAL_ROW_OUT_OF_BOUNDARY;
        *out = ((const fe *)d->w1[index])[row];
        return 0;
    }
    size_t row_size = d->row_size, num_advice = d->num_advice, num_lookup = d->num_lookup;
    size_t max_width = num_advice + num_lookup * 5;
    int first = index < max_width;
    if (!first) index -= max_width;
    size_t num_witness = first ? d->num_w1 : d->num_w2;
    size_t i, j;
    if (index < num_advice) {
        i = 0; j = index;
    } else {
        size_t li = (index - num_advice) / 5, ls = (index - num_advice) % 5;
        int first_round = ls < 3;
        if (!first_round) ls -= 3;
        if (num_witness == 2) {
            if (first_round) { i = 0; j = num_advice + li * 3 + ls; }
            else { i = 1; j = li * 2 + ls; }
        } else if (num_witness == 3) {
            if (first_round) { i = 1; j = li * 3 + ls; }
            else { i = 2; j = li * 2 + ls; }
        } else {
            return ORACLE_EVAL_INVALID_WITNESS_INDEX;
        }
    }
    const void *const *W = first ? d->w1 : d->w2;
    const uint64_t *L = first ? d->w1_len : d->w2_len;
    if (num_witness <= i || L[i] <= j * row_size + row) return ORACLE_EVAL_INVALID_WITNESS_INDEX;
    *out = ((const fe *)W[i])[j * row_size + row];
    return 0;
}

/* GetDataForEval::eval_column_var (src/plonk/eval.rs:57-70): selectors, then fixed, then advice */
static int column_var(const ectx *c, size_t row, size_t index, fe *out) {
    const oracle_eval_domain *d = c->d;
    if (index < d->num_selectors) {
        int s = ((const uint8_t *)d->selectors[index])[row] != 0;
        if (s) *out = c->f->one; else memset(out, 0, 32);
        return 0;
    }
    if (index - d->num_selectors < d->num_fixed) {
        *out = ((const fe *)d->fixed[index - d->num_selectors])[row];
        return 0;
    }
    return advice_var(c, row, index - d->num_selectors - d->num_fixed, out);
}

/* the get_value closure of Calculation::evaluate (graph_evaluator.rs:101-131) */
static int get_value(const ectx *c, const uint32_t *opnd, fe *out) {
    uint32_t kind = opnd[0] & 0xff, rotation = opnd[0] >> 8, index = opnd[1];
    switch (kind) {
        case VS_CONSTANT:
            if (index >= c->n_constants) return ORACLE_EVAL_BAD_PROGRAM;
            *out = c->constants[index];
            return 0;
        case VS_INTERMEDIATE:
            if (index >= c->n_inter) return ORACLE_EVAL_BAD_PROGRAM;
            *out = c->inter[index];
            return 0;
        case VS_FIXED:
            if (index >= c->d->num_fixed) return ORACLE_EVAL_COLUMN_OUT_OF_BOUNDARY;
            if (rotation >= c->n_rot) return ORACLE_EVAL_BAD_PROGRAM;
            if (c->rot[rotation] >= c->d->row_size) return ORACLE_EVAL_ROW_OUT_OF_BOUNDARY;
            *out = ((const fe *)c->d->fixed[index])[c->rot[rotation]];
            return 0;
        case VS_POLY:
            if (rotation >= c->n_rot) return ORACLE_EVAL_BAD_PROGRAM;
            return column_var(c, c->rot[rotation], index, out);
        case VS_CHALLENGE:
            if (index >= c->d->num_challenges) return ORACLE_EVAL_CHALLENGE_OUT_OF_BOUNDARY;
            *out = ((const fe *)c->d->challenges)[index];
            return 0;
    }
    return ORACLE_EVAL_BAD_PROGRAM;
}

int oracle_eval_rows(int field, const uint32_t *code, size_t code_words, const void *constants, size_t n_constants,
                     const int32_t *rotations, size_t n_rotations, uint32_t num_intermediates,
                     const oracle_eval_domain *dom, size_t row_begin, size_t row_end, void *out_) {
    winit();
    const field_t *f = &FIELDS[field];
    fe *out = out_;
    fe *inter = calloc(num_intermediates ? num_intermediates : 1, sizeof(fe));
    size_t *rot = calloc(n_rotations ? n_rotations : 1, sizeof(size_t));
    ectx c = {dom, f, constants, n_constants, inter, num_intermediates, rot, (uint32_t)n_rotations};
    int rc = 0;
    for (size_t row = row_begin; row < row_end && !rc; row++) {
        /* graph_evaluator.rs:366-371: fresh intermediates, rotations resolved with rem_euclid */
        memset(inter, 0, sizeof(fe) * num_intermediates);
        for (size_t k = 0; k < n_rotations; k++) {
            int64_t n = (int64_t)dom->row_size, v = ((int64_t)row + rotations[k]) % n;
            rot[k] = (size_t)(v < 0 ? v + n : v);
        }
        uint32_t last_target = 0;
        int any = 0;
        size_t pc = 0;
        while (pc < code_words && !rc) {
            if (pc + 2 > code_words) { rc = ORACLE_EVAL_BAD_PROGRAM; break; }
            uint32_t op = code[pc] & 0xff, nops = code[pc] >> 8, target = code[pc + 1];
            const uint32_t *o = &code[pc + 2];
            if (pc + 2 + 2 * (size_t)nops > code_words || target >= num_intermediates) { rc = ORACLE_EVAL_BAD_PROGRAM; break; }
            fe a, b, r;
            memset(&r, 0, 32);
            switch (op) {
                case OP_ADD: case OP_SUB: case OP_MUL:
                    if (nops != 2) { rc = ORACLE_EVAL_BAD_PROGRAM; break; }
                    if ((rc = get_value(&c, o, &a)) || (rc = get_value(&c, o + 2, &b))) break;
                    if (op == OP_ADD) fe_add(f, &r, &a, &b);
                    else if (op == OP_SUB) fe_sub(f, &r, &a, &b);
                    else fe_mul(f, &r, &a, &b);
                    break;
                case OP_SQUARE: case OP_DOUBLE: case OP_NEGATE: case OP_STORE:
                    if (nops != 1) { rc = ORACLE_EVAL_BAD_PROGRAM; break; }
                    if ((rc = get_value(&c, o, &a))) break;
                    if (op == OP_SQUARE) fe_sqr(f, &r, &a);
                    else if (op == OP_DOUBLE) fe_dbl(f, &r, &a);
                    else if (op == OP_NEGATE) fe_neg(f, &r, &a);
                    else r = a;
                    break;
                case OP_HORNER: { /* operands: start, factor, parts... ; value = value * factor + part */
                    if (nops < 2) { rc = ORACLE_EVAL_BAD_PROGRAM; break; }
                    fe fac;
                    if ((rc = get_value(&c, o + 2, &fac)) || (rc = get_value(&c, o, &r))) break;
                    for (uint32_t k = 2; k < nops && !rc; k++) {
                        fe part;
                        if ((rc = get_value(&c, o + 2 * k, &part))) break;
                        fe_mul(f, &r, &r, &fac);
                        fe_add(f, &r, &r, &part);
                    }
                    break;
                }
                default: rc = ORACLE_EVAL_BAD_PROGRAM;
            }
            if (rc) break;
            inter[target] = r;
            last_target = target;
            any = 1;
            pc += 2 + 2 * (size_t)nops;
        }
        if (rc) break;
        if (any) out[row - row_begin] = inter[last_target];   /* result of the last calculation */
        else memset(&out[row - row_begin], 0, 32);             /* or ZERO if there is none      */
    }
    free(inter);
    free(rot);
    return rc;
}

/* ------------------------------------------------------------------ FFT (src/fft.rs) */
/* PrimeField::ROOT_OF_UNITY of halo2curves bn256::Fr (S = 28), canonical; = 7^((r-1)/2^28).
 * Validated by the reference's own KAT (src/fft.rs:239-258) in tests/test_oracle.py. */
static const uint64_t FR_ROOT_OF_UNITY[4] = {0xd34f1ed960c37c9cULL, 0x3215cf6dd39329c8ULL, 0x98865ea93dd31f74ULL, 0x03ddb9f5166d18b7ULL};
#define FR_S 28

/* get_omega_or_inv (src/fft.rs:12-24).  Only Fr has a 2-adic subgroup worth transforming (Fq: S = 1). */
int oracle_fft_omega(int field, uint32_t k, int is_inverse, void *out) {
    winit();
    if (field != ORACLE_FR || k > FR_S) return -1;
    const field_t *f = &FIELDS[field];
    fe w, c;
    memcpy(&c, FR_ROOT_OF_UNITY, 32);
    fe_from_canonical(f, &w, &c);
    if (is_inverse) fe_inv(f, &w, &w);
    for (uint32_t i = k; i < FR_S; i++) fe_sqr(f, &w, &w);
    memcpy(out, &w, 32);
    return 0;
}
/* get_ifft_divisor (src/fft.rs:25-27): TWO_INV^k */
void oracle_fft_divisor(int field, uint32_t k, void *out) {
    winit();
    const field_t *f = &FIELDS[field];
    fe two, inv2, acc = f->one;
    fe_from_u64(f, &two, 2);
    fe_inv(f, &inv2, &two);
    for (uint32_t i = 0; i < k; i++) fe_mul(f, &acc, &acc, &inv2);
    memcpy(out, &acc, 32);
}

/* best_fft (src/fft.rs:51-115), serial branch: in-place, a has 2^log_n elements */
void oracle_fft(int field, void *a_, uint32_t log_n, const void *omega_) {
    winit();
    const field_t *f = &FIELDS[field];
    fe *a = a_;
    size_t n = (size_t)1 << log_n;
    fe omega;
    memcpy(&omega, omega_, 32);
    for (size_t k = 0; k < n; k++) {      /* bitreverse swap */
        size_t rk = 0;
        for (uint32_t b = 0; b < log_n; b++) rk |= ((k >> b) & 1) << (log_n - 1 - b);
        if (k < rk) { fe t = a[rk]; a[rk] = a[k]; a[k] = t; }
    }
    size_t half = n / 2;
    fe *tw = malloc(sizeof(fe) * (half ? half : 1));
    fe w = f->one;
    for (size_t i = 0; i < half; i++) { tw[i] = w; fe_mul(f, &w, &w, &omega); }
    size_t chunk = 2, twiddle_chunk = n / 2;
    for (uint32_t s = 0; s < log_n; s++) {
        for (size_t base = 0; base < n; base += chunk) {
            fe *left = &a[base], *right = &a[base + chunk / 2];
            for (size_t i = 0; i < chunk / 2; i++) {
                fe t = right[i];
                if (i) fe_mul(f, &t, &t, &tw[i * twiddle_chunk]);
                right[i] = left[i];
                fe_add(f, &left[i], &left[i], &t);
                fe_sub(f, &right[i], &right[i], &t);
            }
        }
        chunk *= 2;
        twiddle_chunk /= 2;
    }
    free(tw);
}

/* fft / ifft wrappers (src/fft.rs:160-175) */
int oracle_fft_forward(int field, void *a, uint32_t log_n) {
    fe w;
    if (oracle_fft_omega(field, log_n, 0, &w)) return -1;
    oracle_fft(field, a, log_n, &w);
    return 0;
}
int oracle_fft_inverse(int field, void *a_, uint32_t log_n) {
    fe w, d;
    if (oracle_fft_omega(field, log_n, 1, &w)) return -1;
    oracle_fft_divisor(field, log_n, &d);
    oracle_fft(field, a_, log_n, &w);
    const field_t *f = &FIELDS[field];
    fe *a = a_;
    for (size_t i = 0; i < ((size_t)1 << log_n); i++) fe_mul(f, &a[i], &a[i], &d);
    return 0;
}

/* ------------------------------------------------------------------ multi-threaded wrappers
 * The reference runs these loops under rayon (`into_par_iter` over rows, src/nifs/vanilla/mod.rs:109-116;
 * `par_iter` in fold, src/plonk/mod.rs:1104-1131).  Used by bench.py's CPU baseline so the host cores are all busy. */
typedef struct {
    int kind, field, rc;
    const uint32_t *code; size_t code_words; const void *constants; size_t n_constants;
    const int32_t *rotations; size_t n_rotations; uint32_t num_intermediates; const oracle_eval_domain *dom;
    const void *a, *b, *r; const void *const *terms; size_t n_terms;
    size_t begin, end; void *out;
} wjob;
static void *wjob_run(void *p) {
    wjob *j = p;
    if (j->kind == 0)
        j->rc = oracle_eval_rows(j->field, j->code, j->code_words, j->constants, j->n_constants, j->rotations, j->n_rotations,
                                 j->num_intermediates, j->dom, j->begin, j->end, (char *)j->out + j->begin * 32);
    else if (j->kind == 1)
        oracle_fold_w(j->field, (const char *)j->a + j->begin * 32, (const char *)j->b + j->begin * 32, j->end - j->begin, j->r,
                      (char *)j->out + j->begin * 32);
    else {
        const void *shifted[64];
        for (size_t k = 0; k < j->n_terms; k++) shifted[k] = (const char *)j->terms[k] + j->begin * 32;
        oracle_fold_e(j->field, (const char *)j->a + j->begin * 32, shifted, j->n_terms, j->end - j->begin, j->r,
                      (char *)j->out + j->begin * 32);
    }
    return NULL;
}
static int run_jobs(wjob proto, size_t n, int threads) {
    if (threads <= 0) threads = oracle_num_cores();
    if ((size_t)threads > n) threads = n ? (int)n : 1;
    pthread_t *th = malloc(sizeof(pthread_t) * threads);
    wjob *jobs = malloc(sizeof(wjob) * threads);
    size_t per = (n + threads - 1) / threads;
    for (int t = 0; t < threads; t++) {
        jobs[t] = proto;
        jobs[t].begin = (size_t)t * per < n ? (size_t)t * per : n;
        jobs[t].end = jobs[t].begin + per < n ? jobs[t].begin + per : n;
        pthread_create(&th[t], NULL, wjob_run, &jobs[t]);
    }
    int rc = 0;
    for (int t = 0; t < threads; t++) {
        pthread_join(th[t], NULL);
        if (jobs[t].rc) rc = jobs[t].rc;
    }
    free(th);
    free(jobs);
    return rc;
}
int oracle_eval_rows_mt(int field, const uint32_t *code, size_t code_words, const void *constants, size_t n_constants,
                        const int32_t *rotations, size_t n_rotations, uint32_t num_intermediates,
                        const oracle_eval_domain *dom, int threads, void *out) {
    wjob j = {0};
    j.kind = 0; j.field = field; j.code = code; j.code_words = code_words; j.constants = constants; j.n_constants = n_constants;
    j.rotations = rotations; j.n_rotations = n_rotations; j.num_intermediates = num_intermediates; j.dom = dom; j.out = out;
    return run_jobs(j, dom->row_size, threads);
}
void oracle_fold_w_mt(int field, const void *w1, const void *w2, size_t n, const void *r, int threads, void *out) {
    wjob j = {0};
    j.kind = 1; j.field = field; j.a = w1; j.b = w2; j.r = r; j.out = out;
    run_jobs(j, n, threads);
}
void oracle_fold_e_mt(int field, const void *e, const void *const *terms, size_t n_terms, size_t n, const void *r, int threads,
                      void *out) {
    wjob j = {0};
    j.kind = 2; j.field = field; j.a = e; j.terms = terms; j.n_terms = n_terms; j.r = r; j.out = out;
    run_jobs(j, n, threads);
}

/* ------------------------------------------------------------------ lookup argument
 * Arguments::evaluate_m (src/plonk/lookup.rs:278-305): m_i = number of j with l_j == t_i, reported at the FIRST
 * occurrence of each distinct t value and ZERO at later duplicates (`processed_t`); equality is equality of
 * `to_repr()`, i.e. of field elements.  out: n_t elements, F::from_u128(count) in Montgomery form. */
typedef struct { fe v; size_t idx; } keyed;
static int keyed_cmp(const void *a_, const void *b_) {
    const keyed *a = a_, *b = b_;
    for (int i = 3; i >= 0; i--) {
        if (a->v.l[i] != b->v.l[i]) return a->v.l[i] < b->v.l[i] ? -1 : 1;
    }
    return a->idx < b->idx ? -1 : (a->idx > b->idx ? 1 : 0);
}
static int fe_cmp(const fe *a, const fe *b) {
    for (int i = 3; i >= 0; i--) {
        if (a->l[i] != b->l[i]) return a->l[i] < b->l[i] ? -1 : 1;
    }
    return 0;
}
void oracle_lookup_m(int field, const void *l_, size_t n_l, const void *t_, size_t n_t, void *out_) {
    winit();
    const field_t *f = &FIELDS[field];
    const fe *l = l_, *t = t_;
    fe *out = out_;
    keyed *ls = malloc(sizeof(keyed) * (n_l ? n_l : 1)), *ts = malloc(sizeof(keyed) * (n_t ? n_t : 1));
    for (size_t i = 0; i < n_l; i++) { ls[i].v = l[i]; ls[i].idx = i; }
    for (size_t i = 0; i < n_t; i++) { ts[i].v = t[i]; ts[i].idx = i; }
    qsort(ls, n_l, sizeof(keyed), keyed_cmp);
    qsort(ts, n_t, sizeof(keyed), keyed_cmp);
    size_t lp = 0;
    for (size_t i = 0; i < n_t;) {
        size_t j = i;
        while (j < n_t && fe_cmp(&ts[j].v, &ts[i].v) == 0) j++;          /* group of equal t values, index-sorted */
        while (lp < n_l && fe_cmp(&ls[lp].v, &ts[i].v) < 0) lp++;
        size_t q = lp;
        while (q < n_l && fe_cmp(&ls[q].v, &ts[i].v) == 0) q++;
        fe_from_u64(f, &out[ts[i].idx], (uint64_t)(q - lp));              /* first occurrence gets the count ... */
        for (size_t k = i + 1; k < j; k++) memset(&out[ts[k].idx], 0, 32); /* ... later duplicates ZERO          */
        lp = q;
        i = j;
    }
    free(ls);
    free(ts);
}

/* Arguments::evaluate_h_g (src/plonk/lookup.rs:307-319): h_i = 1/(l_i + r), g_i = m_i/(t_i + r); a zero
 * denominator yields ZERO (`Option::from(x.invert()).unwrap_or(F::ZERO)`). */
void oracle_lookup_h_g(int field, const void *l_, const void *t_, const void *m_, size_t n, const void *r_, void *h_, void *g_) {
    winit();
    const field_t *f = &FIELDS[field];
    const fe *l = l_, *t = t_, *m = m_;
    fe *h = h_, *g = g_;
    fe r;
    memcpy(&r, r_, 32);
    for (size_t i = 0; i < n; i++) {
        fe s, inv;
        fe_add(f, &s, &l[i], &r);
        fe_inv(f, &h[i], &s);
        fe_add(f, &s, &t[i], &r);
        fe_inv(f, &inv, &s);
        fe_mul(f, &g[i], &m[i], &inv);
    }
}
