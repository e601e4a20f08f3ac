/*
 * sd_corpus.c — deterministic synthetic pulse-train corpora (bench / test infrastructure).
 *
 * Generates the BASELINE.json workloads directly in the packed batch format of
 * include/sdb200.h.  Message i is a pure function of
 * (seed, i), so any rank can generate exactly its own shard [lo, hi).
 *
 * This is neither the product nor the oracle: it only makes inputs.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#include "../include/sdb200.h"

#define GEN_MAXLIST 16
#define CORPUS_MAX_DIGITS 1024   /* BASELINE config 3: D <= 1 024 digits (the packed domain itself allows SDB_MAX_DIGITS) */

typedef struct GenProto {
    int32_t is_ms;                 /* has a numeric list `sync` */
    int32_t has_clockabs;
    double  clockabs;
    int32_t nsync;   double sync[GEN_MAXLIST];
    int32_t nstart;  double start[GEN_MAXLIST];
    int32_t none;    double one[GEN_MAXLIST];
    int32_t nzero;   double zero[GEN_MAXLIST];
    int32_t nfloat;  double flt[GEN_MAXLIST];
    int32_t npause;  double pause[GEN_MAXLIST];
    int32_t nend;    double end[GEN_MAXLIST];
    int32_t length_min;            /* -1 absent */
    int32_t length_max;            /* -1 absent */
    int32_t reconstruct;
    /* MC / MN */
    int32_t method;                /* same numbering as oracle METHOD ids; 0 none */
    int32_t clock_min, clock_max;  /* clockrange, 0/0 if absent */
    int32_t polarity_invert;
    int32_t table_index;           /* index in protocol-table order */
    int32_t is_119;
} GenProto;

/* ---- RNG: splitmix64 keyed by (seed, message index) ---- */
typedef struct { uint64_t s; } Rng;
static inline uint64_t rng_next(Rng *r)
{
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline void rng_seed(Rng *r, uint64_t seed, uint64_t idx)
{
    r->s = seed * 0xD1342543DE82EF95ull + idx * 0x2545F4914F6CDD1Dull + 0x1234567ull;
    rng_next(r); rng_next(r);
}
static inline uint32_t rng_below(Rng *r, uint32_t n) { return n ? (uint32_t)(rng_next(r) % n) : 0; }
static inline int rng_range(Rng *r, int lo, int hi) { return hi <= lo ? lo : lo + (int)rng_below(r, (uint32_t)(hi - lo + 1)); }
static inline double rng_unit(Rng *r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline double rng_uniform(Rng *r, double lo, double hi) { return lo + (hi - lo) * rng_unit(r); }

/* ---- growable digit pool ---- */
typedef struct { uint8_t *nib; size_t n, cap; } Pool;
static void pool_reserve(Pool *p, size_t extra)
{
    if (p->n + extra > p->cap) {
        while (p->n + extra > p->cap) p->cap = p->cap ? p->cap * 2 : (1u << 20);
        p->nib = realloc(p->nib, p->cap);
    }
}

/* value table of one message: distinct template values -> slot */
typedef struct { double v[SDB_MAX_SLOTS]; int n; } Vals;
static int vals_add(Vals *vs, double v)
{
    for (int i = 0; i < vs->n; i++) if (vs->v[i] == v) return i;
    if (vs->n >= SDB_MAX_SLOTS) return -1;
    vs->v[vs->n] = v;
    return vs->n++;
}
static int vals_add_list(Vals *vs, const double *l, int n)
{
    for (int i = 0; i < n; i++) if (vals_add(vs, l[i]) < 0) return -1;
    return 0;
}
static int vals_find(const Vals *vs, double v)
{
    for (int i = 0; i < vs->n; i++) if (vs->v[i] == v) return i;
    return -1;
}

typedef struct {
    uint8_t d[CORPUS_MAX_DIGITS + 64];
    int n;
} Digits;
static void dig_list(Digits *D, const Vals *vs, const int *slot_id, const double *l, int n)
{
    for (int i = 0; i < n && D->n < CORPUS_MAX_DIGITS + 32; i++) {
        int k = vals_find(vs, l[i]);
        D->d[D->n++] = (uint8_t)(k >= 0 ? slot_id[k] : 9);
    }
}

/* Fill pat[]/pat_ids for nv template values + nnoise noise pulses in a shuffled slot order. */
static void assign_slots(Rng *r, const Vals *vs, int nnoise, double clock, SdbPulseMsg *m, int *slot_id)
{
    int total = vs->n + nnoise;
    if (total > SDB_MAX_SLOTS) total = SDB_MAX_SLOTS;
    /* random distinct ids out of 0..7 */
    int ids[8] = {0, 1, 2, 3, 4, 5, 6, 7};
    for (int i = 7; i > 0; i--) { int j = (int)rng_below(r, (uint32_t)i + 1); int t = ids[i]; ids[i] = ids[j]; ids[j] = t; }
    /* random slot order */
    int order[8];
    for (int i = 0; i < total; i++) order[i] = i;
    for (int i = total - 1; i > 0; i--) { int j = (int)rng_below(r, (uint32_t)i + 1); int t = order[i]; order[i] = order[j]; order[j] = t; }
    uint32_t pid = 0;
    for (int s = 0; s < total; s++) {
        int k = order[s];                       /* which value sits in slot s */
        double val;
        if (k < vs->n) {
            val = rint(vs->v[k] * clock * rng_uniform(r, 0.95, 1.05));
            slot_id[k] = ids[s];
        } else {
            val = (double)rng_range(r, 100, 9000) * (rng_below(r, 2) ? 1.0 : -1.0);
        }
        if (val > 2000000000.0) val = 2000000000.0;
        if (val < -2000000000.0) val = -2000000000.0;
        m->pat[s] = (int32_t)val;
        pid |= (uint32_t)ids[s] << (4 * s);
    }
    m->pat_ids = pid;
    m->npat = (uint8_t)total;
}

static int slot_of_id(const SdbPulseMsg *m, int id)
{
    for (int s = 0; s < m->npat; s++) if ((int)((m->pat_ids >> (4 * s)) & 0xF) == id) return s;
    return 0xFF;
}

static int pick_bits(Rng *r, const GenProto *p, int lo_default, int hi_default)
{
    int nb = rng_range(r, lo_default, hi_default);
    if (rng_below(r, 100) < 80) {
        int lo = p->length_min > 0 ? p->length_min : 1;
        int hi = p->length_max > 0 ? p->length_max : lo + 40;
        if (hi < lo) hi = lo;
        if (nb < lo || nb > hi) nb = rng_range(r, lo, hi);
    }
    return nb;
}

/* ---- MS message ---- */
static void gen_ms(Rng *r, const GenProto *tab, const int *ms_ids, int nms, SdbPulseMsg *m, Digits *D, int16_t *rssi)
{
    const GenProto *p = &tab[ms_ids[rng_below(r, (uint32_t)nms)]];
    double clock = p->has_clockabs && p->clockabs > 0 ? p->clockabs : (double)rng_range(r, 250, 600);
    clock *= rng_uniform(r, 0.9, 1.1);
    Vals vs = {.n = 0};
    vals_add_list(&vs, p->sync, p->nsync);
    vals_add_list(&vs, p->one, p->none);
    vals_add_list(&vs, p->zero, p->nzero);
    int nnoise = rng_range(r, 0, 2);
    if (vs.n + nnoise > SDB_MAX_SLOTS) nnoise = SDB_MAX_SLOTS - vs.n;
    int slot_id[SDB_MAX_SLOTS] = {0};
    assign_slots(r, &vs, nnoise, clock, m, slot_id);
    int nb = pick_bits(r, p, 24, 64);
    D->n = 0;
    dig_list(D, &vs, slot_id, p->sync, p->nsync);
    for (int b = 0; b < nb && D->n + p->none <= CORPUS_MAX_DIGITS; b++) {
        if (rng_below(r, 2)) dig_list(D, &vs, slot_id, p->one, p->none);
        else dig_list(D, &vs, slot_id, p->zero, p->nzero > 0 ? p->nzero : p->none);
    }
    /* CP: the slot whose template value is closest to +1 */
    int best = 0; double bd = 1e300;
    for (int k = 0; k < vs.n; k++) { double d = fabs(vs.v[k] - 1.0); if (d < bd) { bd = d; best = k; } }
    m->cp = (uint8_t)slot_of_id(m, slot_id[best]);
    m->flags = SDB_MSG_VALID;
    *rssi = rng_below(r, 2) ? (int16_t)rng_range(r, 0, 255) : (int16_t)-1;
    if (rng_below(r, 100) < 5) {                      /* 5 % corrupted */
        switch (rng_below(r, 3)) {
        case 0: if (D->n) D->d[rng_below(r, (uint32_t)D->n)] = (uint8_t)rng_below(r, 8); break;
        case 1: if (D->n > 1) D->n = rng_range(r, 1, D->n - 1); break;
        default: *rssi = -2; m->flags = 0; break;  /* non-digit R -> message_synced.py:42-47 rejects */
        }
    }
}

/* ---- MU message ---- */
static void gen_mu(Rng *r, const GenProto *tab, const int *mu_ids, int nmu, SdbPulseMsg *m, Digits *D, int16_t *rssi)
{
    const GenProto *p = &tab[mu_ids[rng_below(r, (uint32_t)nmu)]];   /* mu_ids lists MU-only ids 3x */
    double clock = p->clockabs > 0 ? p->clockabs : (double)rng_range(r, 250, 600);
    clock *= rng_uniform(r, 0.9, 1.1);
    Vals vs = {.n = 0};
    vals_add_list(&vs, p->start, p->nstart);
    vals_add_list(&vs, p->one, p->none);
    vals_add_list(&vs, p->zero, p->nzero);
    int use_pause = p->npause > 0 && vals_add_list(&vs, p->pause, p->npause) == 0;
    int use_end = p->nend > 0 && vals_add_list(&vs, p->end, p->nend) == 0;
    int use_sync = p->is_ms && vals_add_list(&vs, p->sync, p->nsync) == 0;
    int gap_k = -1;
    if (!use_pause && !use_end && !use_sync && p->nstart == 0 && vs.n < SDB_MAX_SLOTS)
        gap_k = vals_add(&vs, -(double)rng_range(r, 20, 40));            /* inter-frame gap */
    int nnoise = rng_range(r, 0, 2);
    if (vs.n + nnoise > SDB_MAX_SLOTS) nnoise = SDB_MAX_SLOTS - vs.n;
    int slot_id[SDB_MAX_SLOTS] = {0};
    assign_slots(r, &vs, nnoise, clock, m, slot_id);

    int nb;
    if (rng_below(r, 100) < 80) {
        int lo = p->length_min > 0 ? p->length_min : 8;
        int hi = p->length_max > 0 ? p->length_max : lo + 40;
        if (hi > 160) hi = lo + 40 < 160 ? 160 : lo + 40;
        if (hi < lo) hi = lo;
        nb = rng_range(r, lo, hi);
    } else nb = rng_range(r, 8, 80);
    int frames = rng_range(r, 2, 4);
    int w = p->none > 0 ? p->none : 2;
    int per_frame = p->nstart + nb * w + p->npause + p->nend + p->nsync + 1;
    while (frames > 1 && frames * per_frame > CORPUS_MAX_DIGITS) frames--;
    while (frames * per_frame > CORPUS_MAX_DIGITS && nb > 1) { nb--; per_frame -= w; }
    uint64_t bits[4] = {rng_next(r), rng_next(r), rng_next(r), rng_next(r)};
    int truncate_last = p->reconstruct && rng_below(r, 100) < 30;
    D->n = 0;
    for (int f = 0; f < frames; f++) {
        dig_list(D, &vs, slot_id, p->start, p->nstart);
        for (int b = 0; b < nb; b++) {
            int bit = (int)((bits[(b >> 6) & 3] >> (b & 63)) & 1);
            int before = D->n;
            if (bit || p->nzero == 0) dig_list(D, &vs, slot_id, p->one, p->none);
            else dig_list(D, &vs, slot_id, p->zero, p->nzero);
            if (truncate_last && f == frames - 1 && b == nb - 1 && D->n > before) D->n--;
        }
        if (truncate_last && f == frames - 1) break;
        if (use_pause) dig_list(D, &vs, slot_id, p->pause, p->npause);
        if (use_end) dig_list(D, &vs, slot_id, p->end, p->nend);
        if (use_sync) dig_list(D, &vs, slot_id, p->sync, p->nsync);
        if (gap_k >= 0 && D->n < CORPUS_MAX_DIGITS + 32) D->d[D->n++] = (uint8_t)slot_id[gap_k];
    }
    if (D->n > CORPUS_MAX_DIGITS) D->n = CORPUS_MAX_DIGITS;
    m->cp = 0xFF;
    m->flags = SDB_MSG_VALID;
    *rssi = rng_below(r, 2) ? (int16_t)rng_range(r, 0, 255) : (int16_t)-1;
    if (rng_below(r, 100) < 5) {
        if (rng_below(r, 2)) { if (D->n) D->d[rng_below(r, (uint32_t)D->n)] = (uint8_t)rng_below(r, 8); }
        else if (D->n > 1) D->n = rng_range(r, 1, D->n - 1);
    }
    if (D->n == 0) m->flags = 0;
}

/*
 * Generate messages [lo, hi) of the corpus (kind = SDB_KIND_MS / SDB_KIND_MU).
 * ids: candidate protocol indices (MU: MU-only ids repeated 3x by the caller).
 * msgs: hi-lo records.  Digit pool is malloc'ed here: *pool_out (bytes, nibble-packed),
 * free with sdc_free().  rssi[i]: -1 absent, -2 corrupt ("1q"), else 0..255.
 */
int sdc_gen_pulse(const GenProto *tab, int ntab, const int32_t *ids, int nids, int kind,
                  uint64_t seed, int64_t lo, int64_t hi,
                  SdbPulseMsg *msgs, int16_t *rssi, uint8_t **pool_out, int64_t *pool_bytes)
{
    (void)ntab;
    Pool P = {0};
    Digits D;
    for (int64_t i = lo; i < hi; i++) {
        Rng r;
        rng_seed(&r, seed, (uint64_t)i);
        SdbPulseMsg *m = &msgs[i - lo];
        memset(m, 0, sizeof *m);
        if (kind == SDB_KIND_MS) gen_ms(&r, tab, ids, nids, m, &D, &rssi[i - lo]);
        else gen_mu(&r, tab, ids, nids, m, &D, &rssi[i - lo]);
        if (!(m->flags & SDB_MSG_VALID)) {            /* same record the host packer writes for a rejected message */
            int16_t keep = rssi[i - lo];
            memset(m, 0, sizeof *m);
            m->cp = 0xFF;
            rssi[i - lo] = keep;
            D.n = 0;
        }
        size_t padded = ((size_t)D.n + 31) / 32 * 32;
        pool_reserve(&P, padded + 64);
        m->doff = (uint32_t)(P.n / 32);
        m->dlen = (uint16_t)D.n;
        memcpy(P.nib + P.n, D.d, (size_t)D.n);
        memset(P.nib + P.n + D.n, SDB_DIGIT_PAD, padded - (size_t)D.n);
        P.n += padded;
    }
    pool_reserve(&P, 64);
    memset(P.nib + P.n, SDB_DIGIT_PAD, 64);
    size_t nbytes = (P.n + 64) / 2;
    uint8_t *out = malloc(nbytes ? nbytes : 1);
    for (size_t b = 0; b < nbytes; b++) out[b] = (uint8_t)(P.nib[2 * b] | (P.nib[2 * b + 1] << 4));
    free(P.nib);
    *pool_out = out;
    *pool_bytes = (int64_t)nbytes;
    return 0;
}

void sdc_free(void *p) { free(p); }

/* ------------------------------------------------------------------------------------------
 * Packed MS / MU batch -> firmware payload lines ("MS;P0=-3886;...;D=1310...;CP=1;SP=3;R=33;"), one per message,
 * '\n'-separated: the input of the line-parser row (signalduino/parser/ms.py, mu.py).  A record without
 * SDB_MSG_VALID is rendered with an empty D (the parser drops it just the same).  Returns the bytes needed
 * (render again with a larger buffer when > cap).
 * ------------------------------------------------------------------------------------------ */
int64_t sdc_render_lines(int kind, const SdbPulseMsg *msgs, const uint8_t *digits, const int16_t *rssi, int64_t n,
                         char *text, int64_t cap, uint32_t *off, uint32_t *len, int framed)
{
    int64_t used = 0;
    char buf[CORPUS_MAX_DIGITS + 256];
    for (int64_t i = 0; i < n; i++) {
        const SdbPulseMsg *m = &msgs[i];
        int k = 0;
        if (framed) buf[k++] = 0x02;                        /* STX ... ETX as the firmware sends it (base.py:174-193) */
        k += sprintf(buf + k, kind == SDB_KIND_MS ? "MS;" : "MU;");
        const int valid = (m->flags & SDB_MSG_VALID) != 0;
        if (valid)
            for (int s = 0; s < m->npat; s++) k += sprintf(buf + k, "P%u=%d;", (m->pat_ids >> (4 * s)) & 0xF, m->pat[s]);
        else
            k += sprintf(buf + k, "P0=1;P1=-1;");
        buf[k++] = 'D'; buf[k++] = '=';
        if (valid) {
            const uint8_t *d = digits + (size_t)m->doff * 16;
            for (int j = 0; j < m->dlen; j++) {
                int nib = (d[j >> 1] >> ((j & 1) * 4)) & 0xF;
                buf[k++] = nib <= 9 ? (char)('0' + nib) : 'x';
            }
        }
        buf[k++] = ';';
        if (kind == SDB_KIND_MS) {
            int cpid = m->cp != 0xFF ? (int)((m->pat_ids >> (4 * m->cp)) & 0xF) : 9;
            k += sprintf(buf + k, "CP=%d;SP=0;", cpid);
        }
        if (rssi && rssi[i] >= 0) k += sprintf(buf + k, "R=%d;", (int)rssi[i]);
        if (framed) buf[k++] = 0x03;
        if (off) off[i] = (uint32_t)used;
        if (len) len[i] = (uint32_t)k;
        if (used + k + 1 <= cap) { memcpy(text + used, buf, (size_t)k); text[used + k] = '\n'; }
        used += k + 1;
    }
    return used;
}
