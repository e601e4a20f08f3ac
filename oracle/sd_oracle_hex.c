/*
 * sd_oracle_hex.c — CPU ORACLE (test infrastructure, NOT product code): MC and MN paths.
 *
 * Literal C restatement of
 *   SDProtocols.demodulate_mc        sd_protocols/sd_protocols.py:76-111
 *   _demodulate_mc_data              sd_protocols/manchester.py:49-144
 *   the 12 mcBit2* / mcRaw decoders  sd_protocols/manchester.py:207-795, helpers.py:90-122
 *   SDProtocols.demodulate_mn        sd_protocols/sd_protocols.py:113-155
 *   the 7 Conv* converters           sd_protocols/helpers.py:190-716
 *
 * "strict" mode (mc_repaired = 0) reproduces the reference AS SHIPPED (TypeError at
 * manchester.py:84 / :120); "repaired" applies the two one-line edits documented in SURVEY §8c:
 *   manchester.py:83   clock_min, clock_max = clockrange[0], clockrange[1]
 *   manchester.py:120  method_func(name, bit_data, protocol_id, len(bit_data))
 */
#define _GNU_SOURCE
#include "sd_oracle.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define HEXMAX (SDB_MAX_HEX + 8)
#define BITMAX (SDB_MAX_HEX * 4 + 64)

typedef struct {
    OraHit *hits; int64_t nh, caph;
    char *pool;   int64_t np, capp;
} HOut;

static void hout_hit(HOut *o, int msg, int proto, const char *payload)
{
    int len = (int)strlen(payload);
    if (o->nh == o->caph) { o->caph = o->caph ? o->caph * 2 : 1024; o->hits = realloc(o->hits, o->caph * sizeof(OraHit)); }
    if (o->np + len > o->capp) { while (o->np + len > o->capp) o->capp = o->capp ? o->capp * 2 : 65536; o->pool = realloc(o->pool, o->capp); }
    OraHit *h = &o->hits[o->nh++];
    h->msg = msg; h->proto = proto; h->bit_length = -1;
    h->payload_off = (int32_t)o->np; h->payload_len = len;
    memcpy(o->pool + o->np, payload, len);
    o->np += len;
}

static int hexval(char c)
{
    if (c >= '0' && c <= '9') return c - '0';
    if (c >= 'A' && c <= 'F') return c - 'A' + 10;
    if (c >= 'a' && c <= 'f') return c - 'a' + 10;
    return -1;
}
static int byte_at(const char *hex, int i) { return hexval(hex[2 * i]) * 16 + hexval(hex[2 * i + 1]); }

static const char *sfind(const char *s, const char *pat, int from)
{
    int n = (int)strlen(s);
    if (from > n) return NULL;
    return strstr(s + from, pat);
}

/* helpers.length_in_range — helpers.py:124-166 */
static int in_range(const OraProto *p, int n)
{
    int min_len = p->has_length_min ? p->length_min : -1;
    if (min_len != -1 && n < min_len) return 0;
    if (p->has_length_max && n > p->length_max) return 0;
    return 1;
}

/* common "min/max then hex" decoders: Hideki :418, Maverick :452, OSV1 :486, OSV2o3 :520, OSPIR :554 */
static int dec_minmax_hex(const OraProto *p, const char *bits, int n, char *res)
{
    int lmin = p->has_length_min ? p->length_min : -1;
    if (n < lmin) return -1;
    int lmax = p->has_length_max ? p->length_max : 9999;
    if (n > lmax) return -1;
    return ora_bin2hex(bits, res) < 0 ? -1 : 1;
}

/* mcBit2Funkbus — manchester.py:207-300.  rc 1 / -1, or -3 = ValueError (int('') on a short frame) */
static int dec_funkbus(const OraProto *p, const char *bits, int n, char *res)
{
    int lmin = p->has_length_min ? p->length_min : -1;
    if (n < lmin) return -1;
    if (p->has_length_max && n > p->length_max) return -1;
    /* :238-239 1->lh 0->hl then mc2dmc (helpers.py:6-26): out[k] = conv[2k+1]==conv[2k+2] ? 0 : 1 */
    char conv[2 * BITMAX + 4], s[BITMAX + 8], t[BITMAX + 16];
    for (int i = 0; i < n; i++) {
        conv[2 * i] = bits[i] == '1' ? 'l' : 'h';
        conv[2 * i + 1] = bits[i] == '1' ? 'h' : 'l';
    }
    int ns = 0;
    for (int i = 1; i < 2 * n - 1; i += 2) s[ns++] = conv[i] == conv[i + 1] ? '0' : '1';
    s[ns] = 0;
    int pid_is_119 = strcmp(p->id, "119") == 0;
    if (pid_is_119) {                                        /* :244-252 */
        const char *f = strstr(s, "01100");
        int pos = f ? (int)(f - s) : -1;
        if (pos >= 0 && pos < 5) {
            snprintf(t, sizeof t, "001%s", s + pos);
            if ((int)strlen(t) < 48) return -1;
        } else return -1;
    } else {
        snprintf(t, sizeof t, "0%s", s);
    }
    int tl = (int)strlen(t);
    int xorv = 0, chk = 0, parity = 0;
    char hex[16];
    for (int i = 0; i < 6; i++) {                            /* :262-278 */
        int from = i * 8, to = from + 8;
        if (to > tl) to = tl;
        if (from >= to) return -3;                           /* int('', 2) */
        int data = 0;
        for (int k = from; k < to; k++) data = data * 2 + (t[k] - '0');
        sprintf(hex + 2 * i, "%02X", data);
        if (i < 5) xorv ^= data;
        else { chk = data & 0x0F; xorv ^= data & 0xE0; data &= 0xF0; }
        for (int tmp = data; tmp; tmp >>= 1) parity ^= tmp & 1;
    }
    if (parity == 1) return -1;
    int xn = ((xorv & 0xF0) >> 4) ^ (xorv & 0x0F), r = 0;    /* :284-293 */
    if (xn & 8) r ^= 0xC;
    if (xn & 4) r ^= 0x2;
    if (xn & 2) r ^= 0x8;
    if (xn & 1) r ^= 0x3;
    if (r != chk) return -1;
    strcpy(res, hex);
    return 1;
}

/* mcBit2Sainlogic — manchester.py:302-354 */
static int dec_sainlogic(const OraProto *p, const char *bits, int n, char *res)
{
    int lmax = p->has_length_max ? p->length_max : 0;
    if (n > lmax) return -1;
    char b[BITMAX + 32];
    strcpy(b, bits);
    if (n < 128) {
        const char *f = strstr(b, "010100");
        int start = f ? (int)(f - b) : -1;
        if (start < 0 || start > 10) return -1;
        while (start < 10) {
            memmove(b + 1, b, strlen(b) + 1);
            b[0] = '1';
            start = (int)(strstr(b, "010100") - b);
        }
        b[128 < (int)strlen(b) ? 128 : strlen(b)] = 0;
        n = (int)strlen(b);
    }
    int lmin = p->has_length_min ? p->length_min : 0;
    if (n < lmin) return -1;
    return ora_bin2hex(b, res) < 0 ? -1 : 1;
}

/* mcBit2AS — manchester.py:356-416 */
static int dec_as(const OraProto *p, const char *bits, int n, char *res)
{
    int lmin = p->has_length_min ? p->length_min : -1;
    int lmax = p->has_length_max ? p->length_max : 9999;
    const char *f = sfind(bits, "1100", 16);
    if (f) {
        int start_pos = (int)(f - bits);
        const char *e = sfind(bits, "1100", start_pos + 16);
        int end_pos = e ? (int)(e - bits) : n;
        int ml = end_pos - start_pos;
        if (ml < lmin) return -1;
        if (ml > lmax) return -1;
        return ora_bin2hex(bits + start_pos, res) < 0 ? -1 : 1;
    }
    if (n < lmin) return -1;
    if (n > lmax) return -1;
    return ora_bin2hex(bits, res) < 0 ? -1 : 1;
}

/* mcBit2TFA — manchester.py:615-719.  On success res holds the Python repr of the list. */
static int dec_tfa(const OraProto *p, const char *bits, int n, char *res)
{
    const char *f = strstr(bits, "111111111101");
    if (!f) return -1;
    int preamble_pos = (int)(f - bits) + 12;
    int message_end = -1, i = 1, nm = 0;
    static __thread char msgs[64][BITMAX / 4 + 4];
    while (message_end < n) {
        const char *e = preamble_pos >= 0 ? sfind(bits, "1111111111101", preamble_pos) : NULL;
        message_end = e ? (int)(e - bits) : -1;
        if (message_end < preamble_pos) message_end = n;
        int ml = message_end - preamble_pos;
        if (in_range(p, ml)) {
            char part[BITMAX + 4];
            memcpy(part, bits + preamble_pos, ml); part[ml] = 0;
            if (nm < 64) ora_bin2hex(part, msgs[nm++]);
        }
        const char *q = sfind(bits, "1101", message_end);
        if (q) preamble_pos = (int)(q - bits) + 4;
        else { preamble_pos = -1; message_end = n; }
        i++;
    }
    if (i == 10) return -1;
    /* :706-711 duplicates: every 2nd occurrence */
    int seen[64] = {0}, nd = 0;
    char *w = res;
    *w++ = '[';
    for (int a = 0; a < nm; a++) {
        int first = a;
        for (int b = 0; b < a; b++) if (strcmp(msgs[b], msgs[a]) == 0) { first = b; break; }
        if (seen[first] == 1) {
            if (nd++) { *w++ = ','; *w++ = ' '; }
            w += sprintf(w, "'%s'", msgs[a]);
        }
        seen[first]++;
    }
    *w++ = ']'; *w = 0;
    return nd > 0 ? 1 : -1;
}

/* hex -> bit string as _convert_mc_hex_to_bits + hex_to_bin_str do (manchester.py:18-47, helpers.py:168-188) */
static int mc_bits(const char *hex, int hlen, int invert, char *bits)
{
    if (hlen == 0) return -1;                                /* int('', 16) -> None */
    int lead = 0, v[HEXMAX];
    for (int i = 0; i < hlen; i++) {
        int x = hexval(hex[i]);
        if (invert && !(hex[i] >= 'a' && hex[i] <= 'f')) x = 15 - x;   /* tr table is upper-case only */
        v[i] = x;
    }
    while (lead < hlen - 1 && v[lead] == 0) lead++;          /* bin(int()) drops leading zero nibbles */
    int n = 0;
    for (int i = lead; i < hlen; i++)
        for (int k = 3; k >= 0; k--) bits[n++] = (char)('0' + ((v[i] >> k) & 1));
    bits[n] = 0;
    return n;
}

/* demodulate_mc — returns SDB_ST_* */
static int demod_mc_one(const OraProto *tab, int np, int repaired, const SdbHexMsg *m, const char *hex,
                        int mi, HOut *o)
{
    if (!(m->flags & SDB_MSG_VALID) || m->proto >= np) return SDB_ST_OK;   /* sd_protocols.py:81-83 */
    const OraProto *p = &tab[m->proto];
    int mcbitnum = m->bitlen;
    int lmin = p->has_length_min ? p->length_min : -1;       /* manchester.py:70-79 */
    if (mcbitnum < lmin) return SDB_ST_OK;
    int lmax = p->has_length_max ? p->length_max : 9999;
    if (mcbitnum > lmax) return SDB_ST_OK;
    if (p->has_clockrange) {                                 /* :81-86 */
        if (!repaired) return SDB_ST_TYPEERROR;              /* int > list */
        if (!(m->clock > p->clock_min && m->clock < p->clock_max)) return SDB_ST_OK;
    }
    int invert = p->polarity_invert;                         /* :91-96 */
    if (m->flags & SDB_HEX_TOGGLE_POLARITY) invert ^= 1;
    char bits[BITMAX];
    int n = mc_bits(hex, m->hlen, invert, bits);             /* :99-102 */
    if (p->method == ORA_M_NONE) return SDB_ST_VALUEERROR;   /* :109 returns a 1-list -> unpack fails */
    if (p->method == ORA_M_UNKNOWN) return SDB_ST_OK;        /* :121-123 */
    if (n < 0) return SDB_ST_TYPEERROR;                      /* :120 len(None) */
    if (!repaired) return SDB_ST_TYPEERROR;                  /* :120 self passed twice */
    if (p->method >= ORA_M_BRESSER_LIGHTNING) return SDB_ST_TYPEERROR;   /* Conv*(msg_data, msg_type) given 4 args */

    char res[BITMAX];
    int rc = -1;
    switch (p->method) {
    case ORA_M_FUNKBUS:   rc = dec_funkbus(p, bits, n, res); if (rc == -3) return SDB_ST_VALUEERROR; break;
    case ORA_M_SAINLOGIC: rc = dec_sainlogic(p, bits, n, res); break;
    case ORA_M_AS:        rc = dec_as(p, bits, n, res); break;
    case ORA_M_HIDEKI: case ORA_M_MAVERICK: case ORA_M_OSV1: case ORA_M_OSV2O3: case ORA_M_OSPIR:
        rc = dec_minmax_hex(p, bits, n, res); break;
    case ORA_M_MCRAW_MANCHESTER: {                           /* manchester.py:588-613 */
        int mx = p->has_length_max ? p->length_max : 0;
        rc = n > mx ? -1 : (ora_bin2hex(bits, res) < 0 ? -1 : 1);
        break;
    }
    case ORA_M_MCRAW_HELPERS:                                /* helpers.py:90-122 */
        if (p->has_length_max) {
            if (p->length_max_is_str) return SDB_ST_TYPEERROR;   /* int > str */
            if (n > p->length_max) { rc = -1; break; }
        }
        rc = ora_bin2hex(bits, res) < 0 ? -1 : 1;
        break;
    case ORA_M_TFA:       rc = dec_tfa(p, bits, n, res); break;
    case ORA_M_GROTHE:    rc = n != 32 ? -1 : (ora_bin2hex(bits, res) < 0 ? -1 : 1); break;   /* :721-754 */
    case ORA_M_SOMFY: {                                      /* :756-795 */
        const char *b = bits;
        int len = n;
        if (n == 57) { b = bits + 1; len = 56; }
        if (len != 56) { rc = -1; break; }
        char tmp[64];
        memcpy(tmp, b, 56); tmp[56] = 0;
        rc = ora_bin2hex(tmp, res) < 0 ? -1 : 1;
        break;
    }
    }
    if (rc != 1) return SDB_ST_OK;
    char payload[BITMAX + 64];
    snprintf(payload, sizeof payload, "%s%s", p->preamble, res);   /* :131-132 */
    hout_hit(o, mi, m->proto, payload);
    return SDB_ST_OK;
}

/* ---------------------------------------------------------------------------------------- MN */

/* helpers.lfsr_digest16 — helpers.py:190-221 */
static int lfsr16(int bytes, int gen, int key, const char *hex)
{
    int lfsr = 0;
    for (int k = 0; k < bytes; k++) {
        int data = byte_at(hex, k);
        for (int i = 7; i >= 0; i--) {
            if ((data >> i) & 1) lfsr ^= key;
            if (key & 1) key = (key >> 1) ^ gen; else key >>= 1;
        }
    }
    return lfsr;
}
/* helpers._calc_crc16 (refin = refout = False, xorout 0) — helpers.py:281-309 */
static int crc16(const char *hex, int nbytes, int poly)
{
    int crc = 0;
    for (int k = 0; k < nbytes; k++) {
        crc ^= byte_at(hex, k) << 8;
        for (int i = 0; i < 8; i++) {
            if (crc & 0x8000) crc = (crc << 1) ^ poly; else crc <<= 1;
            crc &= 0xFFFF;
        }
    }
    return crc;
}

static int demod_mn_one(const OraProto *tab, int np, const SdbHexMsg *m, const char *hex, int mi, HOut *o)
{
    if (!(m->flags & SDB_MSG_VALID) || m->proto >= np) return SDB_ST_OK;   /* sd_protocols.py:115-123 */
    const OraProto *p = &tab[m->proto];
    int n = m->hlen;
    char out[HEXMAX + 96];
    if (p->method < ORA_M_BRESSER_LIGHTNING || p->method == ORA_M_UNKNOWN) return SDB_ST_OK;   /* :125-149 */
    if (n == 0) return SDB_ST_OK;                            /* `if not hex_data` */
    switch (p->method) {
    case ORA_M_BRESSER_LIGHTNING: case ORA_M_BRESSER_7IN1: { /* helpers.py:223-280, :473-523 */
        int seven = p->method == ORA_M_BRESSER_7IN1;
        if (n < (seven ? 46 : 20)) return SDB_ST_OK;
        if (seven && hex[42] == '0' && hex[43] == '0') return SDB_ST_OK;
        char x[HEXMAX];
        for (int i = 0; i < n; i++) x[i] = "0123456789ABCDEF"[hexval(hex[i]) ^ 0xA];
        x[n] = 0;
        int cs = seven ? lfsr16(21, 0x8810, 0xBA95, x + 4) : lfsr16(8, 0x8810, 0xABF9, x + 4);
        int first = (hexval(x[0]) << 12) | (hexval(x[1]) << 8) | (hexval(x[2]) << 4) | hexval(x[3]);
        if ((cs ^ first) != (seven ? 0x6DF1 : 0x899E)) return SDB_ST_OK;
        if (!seven) x[20] = 0;
        strcpy(out, x);
        break;
    }
    case ORA_M_BRESSER_5IN1: {                               /* :382-425 */
        if (n < 52) return SDB_ST_OK;
        int bit_add = 0, ref = 0;
        for (int i = 0; i < 13; i++) {
            int b = byte_at(hex, i), inv = byte_at(hex, i + 13);
            if ((b ^ inv) != 0xFF) return SDB_ST_OK;
            if (i == 0) ref = inv;
            else for (int d = inv; d; d >>= 1) bit_add += d & 1;
        }
        if (bit_add != ref) return SDB_ST_OK;
        memcpy(out, hex + 28, 24); out[24] = 0;
        break;
    }
    case ORA_M_BRESSER_6IN1: {                               /* :427-471 */
        if (n < 36) return SDB_ST_OK;
        int want = (hexval(hex[0]) << 12) | (hexval(hex[1]) << 8) | (hexval(hex[2]) << 4) | hexval(hex[3]);
        if (crc16(hex + 4, 15, 0x1021) != want) return SDB_ST_OK;
        int sum = 0;
        for (int i = 2; i < 18; i++) sum += byte_at(hex, i);
        if ((sum & 0xFF) != 0xFF) return SDB_ST_OK;
        memcpy(out, hex, n); out[n] = 0;
        break;
    }
    case ORA_M_PCA301: {                                     /* :525-579 */
        if (n < 24) return SDB_ST_OK;
        int want = (hexval(hex[20]) << 12) | (hexval(hex[21]) << 8) | (hexval(hex[22]) << 4) | hexval(hex[23]);
        if (crc16(hex, 10, 0x8005) != want) return SDB_ST_OK;
        snprintf(out, sizeof out, "OK 24 %d %d %d %d %d %d %d %d %d %d %04X",
                 byte_at(hex, 0), byte_at(hex, 1), byte_at(hex, 2), byte_at(hex, 3), byte_at(hex, 4),
                 byte_at(hex, 5) & 0x0F, byte_at(hex, 6), byte_at(hex, 7), byte_at(hex, 8), byte_at(hex, 9), want);
        break;
    }
    case ORA_M_KOPP: {                                       /* :581-628 */
        if (n < 4) return SDB_ST_OK;
        int anz = byte_at(hex, 0) + 1;
        if (n < anz * 2 + 2) return SDB_ST_OK;
        int blk = 0xAA;
        for (int i = 0; i < anz; i++) blk ^= byte_at(hex, i);
        if (blk != byte_at(hex, anz)) return SDB_ST_OK;
        out[0] = 'k'; out[1] = 'r';
        memcpy(out + 2, hex, anz * 2); out[2 + anz * 2] = 0;
        break;
    }
    case ORA_M_LACROSSE: {                                   /* :630-716 */
        if (n < 10) return SDB_ST_OK;
        int crc = 0;
        for (int k = 0; k < 4; k++) {
            crc ^= byte_at(hex, k);
            for (int i = 0; i < 8; i++) { if (crc & 0x80) crc = (crc << 1) ^ 0x31; else crc <<= 1; crc &= 0xFF; }
        }
        if (crc != byte_at(hex, 4)) return SDB_ST_OK;
        int b0 = byte_at(hex, 0), b1 = byte_at(hex, 1), b2 = byte_at(hex, 2), b3 = byte_at(hex, 3);
        int addr = ((b0 & 0x0F) << 2) | ((b1 & 0xC0) >> 6);
        int traw = (b1 & 0x0F) * 100 + ((b2 & 0xF0) >> 4) * 10 + (b2 & 0x0F);
        volatile double temperature = ((double)traw / 10.0) - 40.0;      /* float64, as Python */
        if (temperature >= 60 || temperature <= -40) return SDB_ST_OK;
        int typ = ((b3 & 0x7F) == 125) ? 2 : 1;
        volatile double scaled = temperature * 10.0 + 1000.0;
        int ts = ((int)scaled) & 0xFFFF;
        snprintf(out, sizeof out, "OK 9 %d %d %d %d %d", addr, typ | ((b1 & 0x20) << 2), (ts >> 8) & 0xFF, ts & 0xFF, b3);
        break;
    }
    default:
        return SDB_ST_OK;
    }
    hout_hit(o, mi, m->proto, out);
    return SDB_ST_OK;
}

/* ---------------------------------------------------------------------------------------- driver */
typedef struct {
    const OraProto *tab; int nproto; int kind; int repaired;
    const SdbHexMsg *msgs; const uint8_t *digits; int64_t lo, hi;
    uint8_t *status; HOut *buf;
} HexJob;

static void *hex_worker(void *arg)
{
    HexJob *j = arg;
    char hex[HEXMAX];
    for (int64_t i = j->lo; i < j->hi; i++) {
        const SdbHexMsg *m = &j->msgs[i];
        const uint8_t *d = j->digits + (size_t)m->doff * 16;
        for (int k = 0; k < m->hlen; k++) hex[k] = "0123456789ABCDEF"[(d[k >> 1] >> ((k & 1) * 4)) & 0xF];
        hex[m->hlen] = 0;
        int64_t nh0 = j->buf->nh, np0 = j->buf->np;
        int st = j->kind == SDB_KIND_MC ? demod_mc_one(j->tab, j->nproto, j->repaired, m, hex, (int)i, j->buf)
                                        : demod_mn_one(j->tab, j->nproto, m, hex, (int)i, j->buf);
        if (st != SDB_ST_OK) { j->buf->nh = nh0; j->buf->np = np0; }
        j->status[i] = (uint8_t)st;
    }
    return NULL;
}

int ora_demod_hex(const OraProto *tab, int nproto, int kind, int mc_repaired,
                  const SdbHexMsg *msgs, const uint8_t *digits, int64_t n,
                  uint8_t *status, OraHit *hits, int64_t hits_cap,
                  char *pool, int64_t pool_cap, int64_t *nhits, int64_t *pool_used,
                  int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    HOut *bufs = calloc(nthreads, sizeof(HOut));
    HexJob *jobs = calloc(nthreads, sizeof(HexJob));
    pthread_t *th = calloc(nthreads, sizeof(pthread_t));
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (HexJob){tab, nproto, kind, mc_repaired, msgs, digits, n * t / nthreads, n * (t + 1) / nthreads,
                           status, &bufs[t]};
        if (nthreads == 1) hex_worker(&jobs[t]);
        else pthread_create(&th[t], NULL, hex_worker, &jobs[t]);
    }
    if (nthreads > 1) for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    int64_t tot_h = 0, tot_p = 0;
    for (int t = 0; t < nthreads; t++) { tot_h += bufs[t].nh; tot_p += bufs[t].np; }
    *nhits = tot_h; *pool_used = tot_p;
    int rc = 0;
    if (tot_h > hits_cap || tot_p > pool_cap) rc = -3;
    else {
        int64_t ho = 0, po = 0;
        for (int t = 0; t < nthreads; t++) {
            for (int64_t i = 0; i < bufs[t].nh; i++) { hits[ho] = bufs[t].hits[i]; hits[ho].payload_off += (int32_t)po; ho++; }
            if (bufs[t].np) memcpy(pool + po, bufs[t].pool, bufs[t].np);
            po += bufs[t].np;
        }
    }
    for (int t = 0; t < nthreads; t++) { free(bufs[t].hits); free(bufs[t].pool); }
    free(bufs); free(jobs); free(th);
    return rc;
}
