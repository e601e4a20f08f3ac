/*
 * sd_corpus_hex.c — deterministic synthetic MC (Manchester) and MN corpora.
 * Bench / test infrastructure: makes inputs only.  Message i is a pure function of (seed, i).
 *
 * MC: protocol ~ U(12 manchester ids) with Oregon v2/v3 ('10') weighted 40 %; C inside the
 * clockrange 90 %; L inside [length_min, length_max] 85 %; hex = random nibbles with the
 * protocol's sync structure planted (AS '1100', TFA duplicate frames, Funkbus frames with valid
 * parity + checksum, Sainlogic '010100', Grothe 8 nibbles, Somfy 14 nibbles) so that both length
 * domains of SURVEY App. A.5 are respected.
 * MN: protocol ~ U(ids with a Conv* method); 50 % of the frames carry a valid check
 * (CRC16 / LFSR16 digest / CRC8 / XOR / bit-count), the rest are random hex of the same length.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/sdb200.h"

/* must match GenProto in sd_corpus.c */
#define GEN_MAXLIST 16
typedef struct GenProto {
    int32_t is_ms, has_clockabs;
    double  clockabs;
    int32_t nsync;   double sync[GEN_MAXLIST];
    int32_t nstart;  double start[GEN_MAXLIST];
    int32_t none;    double one[GEN_MAXLIST];
    int32_t nzero;   double zero[GEN_MAXLIST];
    int32_t nfloat;  double flt[GEN_MAXLIST];
    int32_t npause;  double pause[GEN_MAXLIST];
    int32_t nend;    double end[GEN_MAXLIST];
    int32_t length_min, length_max, reconstruct;
    int32_t method, clock_min, clock_max, polarity_invert, table_index, is_119;
} GenProto;

enum { M_FUNKBUS = 1, M_SAINLOGIC, M_AS, M_HIDEKI, M_MAVERICK, M_OSV1, M_OSV2O3, M_OSPIR, M_MCRAW, M_MCRAW_H, M_TFA,
       M_GROTHE, M_SOMFY, M_LIGHTNING, M_5IN1, M_6IN1, M_7IN1, M_PCA301, M_KOPP, M_LACROSSE };

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
    r->s = seed * 0xD1342543DE82EF95ull + idx * 0x2545F4914F6CDD1Dull + 0x7654321ull;
    rng_next(r); rng_next(r);
}
static inline uint32_t rng_below(Rng *r, uint32_t n) { return n ? (uint32_t)(rng_next(r) % n) : 0; }
static inline int rng_range(Rng *r, int lo, int hi) { return hi <= lo ? lo : lo + (int)rng_below(r, (uint32_t)(hi - lo + 1)); }

typedef struct { uint8_t nib[SDB_MAX_HEX + 8]; int n; } Hex;

static void hex_random(Rng *r, Hex *h, int n)
{
    if (n > SDB_MAX_HEX) n = SDB_MAX_HEX;
    if (n < 1) n = 1;
    for (int i = 0; i < n; i++) h->nib[i] = (uint8_t)rng_below(r, 16);
    h->n = n;
}
/* bits (0/1 bytes) -> nibbles, optionally inverted (so that the decoder's inversion restores them) */
static void hex_from_bits(Hex *h, const uint8_t *bits, int nbits, int invert)
{
    int nn = (nbits + 3) / 4;
    if (nn > SDB_MAX_HEX) nn = SDB_MAX_HEX;
    for (int i = 0; i < nn; i++) {
        int v = 0;
        for (int k = 0; k < 4; k++) { int bi = 4 * i + k; v = (v << 1) | (bi < nbits ? bits[bi] : 1); }
        h->nib[i] = (uint8_t)(invert ? 15 - v : v);
    }
    h->n = nn;
}
static int byte_of(const Hex *h, int i) { return (h->nib[2 * i] << 4) | h->nib[2 * i + 1]; }
static void set_byte(Hex *h, int i, int v) { h->nib[2 * i] = (uint8_t)((v >> 4) & 15); h->nib[2 * i + 1] = (uint8_t)(v & 15); }

/* ---- MC ---- */
static void gen_mc(Rng *r, const GenProto *tab, const int32_t *ids, int nids, int osv_slot, SdbHexMsg *m, Hex *h)
{
    int pick = (osv_slot >= 0 && rng_below(r, 100) < 40) ? osv_slot : (int)rng_below(r, (uint32_t)nids);
    const GenProto *p = &tab[ids[pick]];
    int inv = p->polarity_invert;
    int lmin = p->length_min > 0 ? p->length_min : 8, lmax = p->length_max > 0 ? p->length_max : lmin + 64;
    if (lmax < lmin) lmax = lmin;
    int clock;
    if (p->clock_max > p->clock_min + 1 && rng_below(r, 100) < 90) clock = rng_range(r, p->clock_min + 1, p->clock_max - 1);
    else clock = rng_below(r, 2) ? p->clock_min - rng_range(r, 0, 200) : p->clock_max + rng_range(r, 0, 200);
    int L = rng_below(r, 100) < 85 ? rng_range(r, lmin, lmax) : rng_range(r, 8, 260);
    hex_random(r, h, (L + 3) / 4);
    uint8_t bits[SDB_MAX_HEX * 4];
    switch (p->method) {
    case M_GROTHE:
        if (rng_below(r, 100) < 70) { hex_random(r, h, 8); if (h->nib[0] == (inv ? 15 : 0)) h->nib[0] = 5; }
        break;
    case M_SOMFY:
        hex_random(r, h, 14);
        if (h->nib[0] == (inv ? 15 : 0)) h->nib[0] = 10;
        break;
    case M_AS:
        if (h->n > 6 && rng_below(r, 2)) h->nib[rng_range(r, 4, h->n - 1)] = (uint8_t)(inv ? 3 : 12);   /* '1100' */
        break;
    case M_SAINLOGIC:
        if (h->n >= 3 && rng_below(r, 2)) {           /* decoded bits start x0101 0100: sync at bit 1 */
            int a = 5, b = 4;
            if (rng_below(r, 2)) hex_random(r, h, rng_range(r, 28, 31));   /* < 128 bits: the re-sync branch (:331-346) */
            h->nib[0] = (uint8_t)(inv ? 15 - a : a);
            h->nib[1] = (uint8_t)(inv ? 15 - b : b);
        }
        break;
    case M_FUNKBUS:
        if (rng_below(r, 100) < 30) {
            /* differential frame t: 0x2C, 4 random bytes, last byte = free nibble | checksum nibble */
            int t[6];
            t[0] = 0x2C;
            for (int i = 1; i < 5; i++) t[i] = (int)rng_below(r, 256);
            int hi = (int)rng_below(r, 16) << 4;
            int xorv = t[0] ^ t[1] ^ t[2] ^ t[3] ^ t[4] ^ (hi & 0xE0);
            int xn = ((xorv & 0xF0) >> 4) ^ (xorv & 0x0F), chk = 0;
            if (xn & 8) chk ^= 0xC;
            if (xn & 4) chk ^= 0x2;
            if (xn & 2) chk ^= 0x8;
            if (xn & 1) chk ^= 0x3;
            int par = 0;
            for (int i = 0; i < 5; i++) par ^= __builtin_parity((unsigned)t[i]);
            par ^= __builtin_parity((unsigned)(hi & 0xF0));
            if (par) hi ^= 0x10;                      /* bit 4 is outside the checksum: use it to even the parity */
            t[5] = hi | chk;
            /* s = t[3:48] (45 bits) + 2 random; manchester bits b: b0 = 1, b[k+1] = s[k] ? b[k] : !b[k] */
            uint8_t s[64];
            int ns = 0;
            for (int k = 3; k < 48; k++) s[ns++] = (uint8_t)((t[k >> 3] >> (7 - (k & 7))) & 1);
            s[ns++] = (uint8_t)rng_below(r, 2); s[ns++] = (uint8_t)rng_below(r, 2);
            bits[0] = 1;
            for (int k = 0; k < ns; k++) bits[k + 1] = s[k] ? bits[k] : (uint8_t)!bits[k];
            hex_from_bits(h, bits, ns + 1, inv);
            if (L < lmin || L > lmax) L = rng_range(r, lmin, lmax);
        }
        break;
    case M_TFA:
        if (rng_below(r, 100) < 60) {
            uint8_t msg[64];
            int ml = lmin > 0 && lmin <= 60 ? lmin : 52;
            for (int i = 0; i < ml; i++) msg[i] = (uint8_t)rng_below(r, 2);
            int nb = 0, copies = rng_range(r, 2, 3);
            for (int i = 0; i < 10; i++) bits[nb++] = 1;
            bits[nb++] = 0; bits[nb++] = 1;                                  /* '111111111101' */
            for (int c = 0; c < copies; c++) {
                for (int i = 0; i < ml; i++) bits[nb++] = msg[i];
                if (c == 1 && copies == 3 && rng_below(r, 2)) bits[nb - 1] ^= 1;   /* sometimes a corrupted repeat */
                if (c + 1 < copies) { for (int i = 0; i < 11; i++) bits[nb++] = 1; bits[nb++] = 0; bits[nb++] = 1; }
            }
            hex_from_bits(h, bits, nb, inv);
            L = rng_range(r, lmin, lmax);
        }
        break;
    default:
        if (rng_below(r, 100) < 90 && h->nib[0] == (inv ? 15 : 0)) h->nib[0] = (uint8_t)(inv ? 6 : 9);
        break;
    }
    m->proto = (uint16_t)p->table_index;
    m->clock = clock;
    m->bitlen = (int16_t)L;
    m->flags = SDB_MSG_VALID;
}

/* ---- MN ---- */
static int lfsr16(const Hex *x, int first_nib, int bytes, int gen, int key)
{
    int lfsr = 0;
    for (int k = 0; k < bytes; k++) {
        int data = (x->nib[first_nib + 2 * k] << 4) | x->nib[first_nib + 2 * k + 1];
        for (int i = 7; i >= 0; i--) {
            if ((data >> i) & 1) lfsr ^= key;
            key = (key & 1) ? ((key >> 1) ^ gen) : (key >> 1);
        }
    }
    return lfsr;
}
static int crc16(const Hex *h, int first_byte, int nbytes, int poly)
{
    int crc = 0;
    for (int k = 0; k < nbytes; k++) {
        crc ^= byte_of(h, first_byte + k) << 8;
        for (int i = 0; i < 8; i++) crc = (crc & 0x8000) ? (((crc << 1) ^ poly) & 0xFFFF) : ((crc << 1) & 0xFFFF);
    }
    return crc;
}

static void gen_mn(Rng *r, const GenProto *tab, const int32_t *ids, int nids, SdbHexMsg *m, Hex *h)
{
    const GenProto *p = &tab[ids[rng_below(r, (uint32_t)nids)]];
    int valid = rng_below(r, 2);
    switch (p->method) {
    case M_6IN1: {
        hex_random(r, h, 36 + 2 * rng_range(r, 0, 2));
        if (valid) {
            int sum = 0;
            for (int i = 2; i < 17; i++) sum += byte_of(h, i);
            set_byte(h, 17, (0xFF - (sum & 0xFF)) & 0xFF);
            int c = crc16(h, 2, 15, 0x1021);
            set_byte(h, 0, c >> 8); set_byte(h, 1, c & 0xFF);
        }
        break;
    }
    case M_5IN1: {
        hex_random(r, h, 52 + 2 * rng_range(r, 0, 2));
        if (valid) {
            int pop = 0;
            for (int i = 1; i < 13; i++) { int inv = (~byte_of(h, i)) & 0xFF; set_byte(h, i + 13, inv); pop += __builtin_popcount((unsigned)inv); }
            set_byte(h, 13, pop & 0xFF);
            set_byte(h, 0, (~pop) & 0xFF);
        }
        break;
    }
    case M_7IN1: case M_LIGHTNING: {
        int seven = p->method == M_7IN1;
        int bytes = seven ? 21 : 8;
        hex_random(r, h, seven ? 46 + 2 * rng_range(r, 0, 2) : 20 + 2 * rng_range(r, 0, 2));
        if (valid) {
            Hex x = *h;                                   /* work in the XOR 0xA domain */
            for (int i = 0; i < x.n; i++) x.nib[i] ^= 0xA;
            if (seven && h->nib[42] == 0 && h->nib[43] == 0) { h->nib[43] = 7; x.nib[43] = 7 ^ 0xA; }
            int d = lfsr16(&x, 4, bytes, 0x8810, seven ? 0xBA95 : 0xABF9) ^ (seven ? 0x6DF1 : 0x899E);
            for (int k = 0; k < 4; k++) h->nib[k] = (uint8_t)(((d >> (12 - 4 * k)) & 15) ^ 0xA);
        }
        break;
    }
    case M_PCA301: {
        hex_random(r, h, 24);
        if (valid) { int c = crc16(h, 0, 10, 0x8005); set_byte(h, 10, c >> 8); set_byte(h, 11, c & 0xFF); }
        break;
    }
    case M_KOPP: {
        int anz = rng_range(r, 5, 12);
        hex_random(r, h, anz * 2 + 2 + 2 * rng_range(r, 0, 2));
        set_byte(h, 0, anz - 1);
        if (valid) { int b = 0xAA; for (int i = 0; i < anz; i++) b ^= byte_of(h, i); set_byte(h, anz, b); }
        break;
    }
    case M_LACROSSE: {
        hex_random(r, h, 10 + 2 * rng_range(r, 0, 1));
        if (rng_below(r, 4)) {                            /* mostly decimal temperature nibbles */
            h->nib[3] = (uint8_t)rng_range(r, 0, 9); h->nib[4] = (uint8_t)rng_range(r, 0, 9); h->nib[5] = (uint8_t)rng_range(r, 0, 9);
        }
        if (valid) {
            int crc = 0;
            for (int k = 0; k < 4; k++) {
                crc ^= byte_of(h, k);
                for (int i = 0; i < 8; i++) crc = (crc & 0x80) ? (((crc << 1) ^ 0x31) & 0xFF) : ((crc << 1) & 0xFF);
            }
            set_byte(h, 4, crc);
        }
        break;
    }
    default:
        hex_random(r, h, p->length_min > 0 ? p->length_min : 20);
        break;
    }
    m->proto = (uint16_t)p->table_index;
    m->clock = 0;
    m->bitlen = 0;
    m->flags = SDB_MSG_VALID;
}

/*
 * Messages [lo, hi) of an MC (kind 2) or MN (kind 3) corpus.  ids = candidate protocol indices;
 * osv_slot = position of protocol '10' in ids (MC weighting), -1 if absent.
 */
int sdc_gen_hex(const GenProto *tab, int ntab, const int32_t *ids, int nids, int osv_slot, int kind,
                uint64_t seed, int64_t lo, int64_t hi, SdbHexMsg *msgs, uint8_t **pool_out, int64_t *pool_bytes)
{
    (void)ntab;
    size_t cap = 1u << 20, n = 0;
    uint8_t *nib = malloc(cap);
    Hex h;
    for (int64_t i = lo; i < hi; i++) {
        Rng r;
        rng_seed(&r, seed, (uint64_t)i);
        SdbHexMsg *m = &msgs[i - lo];
        memset(m, 0, sizeof *m);
        h.n = 0;
        if (kind == SDB_KIND_MC) gen_mc(&r, tab, ids, nids, osv_slot, m, &h);
        else gen_mn(&r, tab, ids, nids, m, &h);
        size_t padded = ((size_t)h.n + 31) / 32 * 32;
        if (n + padded + 64 > cap) { while (n + padded + 64 > cap) cap *= 2; nib = realloc(nib, cap); }
        m->doff = (uint32_t)(n / 32);
        m->hlen = (uint16_t)h.n;
        memcpy(nib + n, h.nib, (size_t)h.n);
        memset(nib + n + h.n, SDB_DIGIT_PAD, padded - (size_t)h.n);
        n += padded;
    }
    memset(nib + n, SDB_DIGIT_PAD, 64);
    size_t nbytes = (n + 64) / 2;
    uint8_t *out = malloc(nbytes ? nbytes : 1);
    for (size_t b = 0; b < nbytes; b++) out[b] = (uint8_t)(nib[2 * b] | (nib[2 * b + 1] << 4));
    free(nib);
    *pool_out = out;
    *pool_bytes = (int64_t)nbytes;
    return 0;
}
