/*
 * sd_oracle.c — CPU ORACLE (test infrastructure, NOT product code).  See sd_oracle.h.
 *
 * Literal, string-based C restatement of the PySignalduino demodulation hot path.
 * Every function cites the reference file:line it follows (paths relative to the
 * PySignalduino tree).  Build: `make -C oracle` (gcc -O2 -ffp-contract=off -pthread).
 */
#define _GNU_SOURCE
#include "sd_oracle.h"

#include <math.h>
#include <regex.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

/* ------------------------------------------------------------------------------------------
 * float helpers
 * ---------------------------------------------------------------------------------------- */

/* CPython round(x, 1) == correctly rounded one-decimal conversion of the exact binary double,
 * read back with strtod (Objects/floatobject.c double_round).  Independent slow form. */
double ora_round1_printf(double x)
{
    char buf[400];
    snprintf(buf, sizeof buf, "%.1f", x);
    return strtod(buf, NULL);
}

/* Fast form: 10*x = y + e exactly (FMA residual); round-half-even on the exact value. */
double ora_round1(double x)
{
    double y = x * 10.0;
    double e = fma(x, 10.0, -y);
    double r = nearbyint(y);
    double d = y - r;
    if (d == 0.5 || d == -0.5) {
        double fl = floor(y);
        if (e > 0) r = fl + 1.0;
        else if (e < 0) r = fl;
    }
    return r / 10.0;
}

/* pattern_utils.py:15-26 */
double ora_tolerance(double v)
{
    double a = fabs(v);
    if (a > 3) {
        if (a > 16) return a * 0.18;
        return a * 0.3;
    }
    return 1.0;
}

/* ------------------------------------------------------------------------------------------
 * pattern_exists — pattern_utils.py:34-136
 * ids[i] is the one-character pattern id of dict entry i (dict order), vals[i] its normalised value.
 * Returns the length of the matching id string written to out, or -1.
 * ---------------------------------------------------------------------------------------- */
int ora_pattern_exists(const double *search, int ns, const char *ids, const double *vals, int npat,
                       const char *raw, char *out)
{
    double uniq[ORA_MAXLIST];
    int nu = 0;
    /* :54-57 unique values, first-appearance order */
    for (int i = 0; i < ns; i++) {
        int seen = 0;
        for (int j = 0; j < nu; j++) if (uniq[j] == search[i]) { seen = 1; break; }
        if (!seen) uniq[nu++] = search[i];
    }
    int cand[ORA_MAXLIST][SDB_MAX_SLOTS];
    int ncand[ORA_MAXLIST];
    for (int u = 0; u < nu; u++) {
        double tol = ora_tolerance(uniq[u]);                       /* :63 */
        double gaps[SDB_MAX_SLOTS];
        int n = 0;
        for (int p = 0; p < npat; p++) {                           /* :73-76 */
            double gap = fabs(vals[p] - uniq[u]);
            if (gap <= 0.001 || gap <= tol) {
                /* stable insertion by gap (:83 list.sort is stable) */
                int k = n;
                while (k > 0 && gaps[k - 1] > gap) { gaps[k] = gaps[k - 1]; cand[u][k] = cand[u][k - 1]; k--; }
                gaps[k] = gap; cand[u][k] = p; n++;
            }
        }
        if (n == 0) return -1;                                      /* :78-80 */
        ncand[u] = n;
    }
    long total = 1;                                                 /* :93-101 */
    for (int u = 0; u < nu; u++) total *= ncand[u];
    if (total > 10000) return -1;

    int odo[ORA_MAXLIST];
    for (int u = 0; u < nu; u++) odo[u] = 0;
    for (long c = 0; c < total; c++) {                              /* :111 itertools.product order */
        int dup = 0;                                                /* :114 */
        for (int a = 0; a < nu && !dup; a++)
            for (int b = a + 1; b < nu; b++)
                if (cand[a][odo[a]] == cand[b][odo[b]]) { dup = 1; break; }
        if (!dup) {
            for (int i = 0; i < ns; i++) {                          /* :118-127 */
                int u = 0;
                while (uniq[u] != search[i]) u++;
                out[i] = ids[cand[u][odo[u]]];
            }
            out[ns] = 0;
            if (strstr(raw, out)) return ns;                        /* :133 */
        }
        for (int u = nu - 1; u >= 0; u--) {                         /* last list varies fastest */
            if (++odo[u] < ncand[u]) break;
            odo[u] = 0;
        }
    }
    return -1;
}

/* ------------------------------------------------------------------------------------------
 * helpers.bin_str_2_hex_str — helpers.py:28-64.  Returns length, or -1 for None.
 * ---------------------------------------------------------------------------------------- */
int ora_bin2hex(const char *bits, char *out)
{
    int n = (int)strlen(bits);
    if (n == 0) { out[0] = 0; return 0; }
    for (int i = 0; i < n; i++) if (bits[i] != '0' && bits[i] != '1') return -1;
    int nd = (n + 3) / 4;
    int index = n - 4, pos = nd;
    for (;;) {
        int cw = 4;
        if (index < 0) { cw += index; index = 0; }
        int v = 0;
        for (int k = 0; k < cw; k++) v = v * 2 + (bits[index + k] - '0');
        out[--pos] = "0123456789ABCDEF"[v];
        index -= 4;
        if (index <= -4) break;
    }
    out[nd] = 0;
    return nd;
}

/* ------------------------------------------------------------------------------------------
 * postDemo_* — postdemodulation.py.  in/out: one byte per bit.  rc 1 ok, 0 reject, -2 ValueError.
 * ---------------------------------------------------------------------------------------- */
static int bits_val(const uint8_t *b, int from, int n)
{
    int v = 0;
    for (int i = 0; i < n; i++) v = v * 2 + b[from + i];
    return v;
}
static int bits_find(const uint8_t *b, int n, const char *pat)
{
    int m = (int)strlen(pat);
    for (int i = 0; i + m <= n; i++) {
        int ok = 1;
        for (int k = 0; k < m; k++) if (b[i + k] != (uint8_t)(pat[k] - '0')) { ok = 0; break; }
        if (ok) return i;
    }
    return -1;
}

/* postdemodulation.py:27-88 */
static int pd_em(const uint8_t *in, int n, uint8_t *out, int *no)
{
    int st = bits_find(in, n, "0000000001");
    if (st < 0) return 0;
    const uint8_t *s = in + st + 10;
    int len = n - (st + 10);
    if (len != 89) return 0;
    int crc = 0, k = 0;
    for (int count = 0; count < len; count += 9) {
        if (count + 8 < len) {
            int byte = bits_val(s, count, 8);
            if (count < len - 10) {
                for (int j = 7; j >= 0; j--) out[k++] = s[count + j];
                crc ^= byte;
            }
        }
    }
    if (crc != bits_val(s, len - 8, 8)) return 0;
    *no = k;
    return 1;
}

/* postdemodulation.py:90-137 */
static int pd_revolt(const uint8_t *in, int n, uint8_t *out, int *no)
{
    if (n < 96) return 0;
    int chk = bits_val(in, 88, 8), sum = 0;
    for (int b = 0; b < 88; b += 8) sum += bits_val(in, b, 8);
    if ((sum & 0xFF) != chk) return 0;
    memcpy(out, in, 88);
    *no = 88;
    return 1;
}

static int first_one(const uint8_t *in, int n)
{
    for (int i = 0; i < n; i++) if (in[i] == 1) return i;
    return -1;
}
static int parity9_ok(const uint8_t *b, int len)
{
    for (int s = 0; s < len; s += 9) {
        int p = 0;
        for (int i = s; i < s + 9 && i < len; i++) p += b[i];
        if (p % 2) return 0;
    }
    return 1;
}
/* remove index k*9+8 bits: "for b in range(len-1, 0, -9): pop(b)" */
static int strip9(const uint8_t *b, int len, uint8_t *out)
{
    int k = 0;
    for (int i = 0; i < len; i++) if (i % 9 != 8) out[k++] = b[i];
    return k;
}

/* postdemodulation.py:139-243 */
static int pd_fs20(const uint8_t *in, int n, uint8_t *out, int *no)
{
    int ds = first_one(in, n);
    if (ds < 0) return 0;
    const uint8_t *b = in + ds + 1;
    int len = n - ds - 1;
    if (len == 46 || len == 55) len--;
    if (len != 45 && len != 54) return 0;
    int sum = 6;
    for (int i = 0; i < len - 9; i += 9) sum += bits_val(b, i, 8);
    int chk = bits_val(b, len - 9, 8);
    if (((sum + 6) & 0xFF) == chk) return 0;
    if ((sum & 0xFF) != chk) return 0;
    if (!parity9_ok(b, len)) return 0;
    uint8_t t[64];
    int k = strip9(b, len, t);
    int o = 0;
    if (len == 45) {           /* k == 40: del [32:40]; insert 8 zeros at 24 */
        for (int i = 0; i < 24; i++) out[o++] = t[i];
        for (int i = 0; i < 8; i++) out[o++] = 0;
        for (int i = 24; i < 32; i++) out[o++] = t[i];
    } else {                   /* k == 48: del [40:48] */
        for (int i = 0; i < 40; i++) out[o++] = t[i];
    }
    (void)k;
    *no = o;
    return 1;
}

/* postdemodulation.py:245-337 */
static int pd_fht80(const uint8_t *in, int n, uint8_t *out, int *no)
{
    int ds = first_one(in, n);
    if (ds < 0) return 0;
    const uint8_t *b = in + ds + 1;
    int len = n - ds - 1;
    if (len == 55) len--;
    if (len != 54) return 0;
    int sum = 12;
    for (int i = 0; i < 45; i += 9) sum += bits_val(b, i, 8);
    int chk = bits_val(b, 45, 8);
    if (((sum - 6) & 0xFF) == chk) return 0;
    if ((sum & 0xFF) != chk) return 0;
    if (!parity9_ok(b, 54)) return 0;
    *no = strip9(b, 54, out);
    return 1;
}

/* postdemodulation.py:339-423 */
static int pd_fht80tf(const uint8_t *in, int n, uint8_t *out, int *no)
{
    if (n < 46) return 0;
    int ds = first_one(in, n);
    if (ds < 0) return 0;
    const uint8_t *b = in + ds + 1;
    int len = n - ds - 1;
    if (len != 45) return 0;
    int sum = 12;
    for (int i = 0; i < 36; i += 9) sum += bits_val(b, i, 8);
    int chk = bits_val(b, 36, 8);
    if ((sum & 0xFF) != chk) return 0;
    if (!parity9_ok(b, 45)) return 0;
    uint8_t t[64];
    strip9(b, 45, t);          /* 40 bits */
    if (t[26] != 0) return 0;
    memcpy(out, t, 32);        /* del [32:40] */
    *no = 32;
    return 1;
}

static int rev4(const uint8_t *b, int from, int avail)
{
    /* int("".join(reversed(bits[from:from+4])), 2) for a slice clipped to `avail` bits */
    int v = 0;
    for (int i = avail - 1; i >= 0; i--) v = v * 2 + b[from + i];
    return v;
}

/* postdemodulation.py:425-578 */
static int pd_ws2000(const uint8_t *in, int n, uint8_t *out, int *no)
{
    static const int dlw[8] = {35, 50, 35, 50, 70, 40, 40, 85};
    int ds = first_one(in, n);
    if (ds < 0) return 0;
    int dl = n - ds;
    int dl1 = dl - (dl % 5);
    int avail = n - (ds + 1);
    if (avail > 4) avail = 4;
    if (avail <= 0) return -2;                     /* int('', 2) -> ValueError (:471) */
    int typ = rev4(in, ds + 1, avail);
    if (typ > 7) return 0;
    if (typ == 1 && (dl == 45 || dl == 46)) dl1 += 5;
    if (dlw[typ] != dl1) return 0;
    if (ds > 10) return 0;
    int index = 0, dataindex = 0, check = 0, sum = 5;
    while (index < dl - 1) {
        if (in[index + ds] != 1) return 0;
        dataindex = index + ds + 1;
        int rest = n - dataindex;
        if (rest < 4) return 0;
        int data = rev4(in, dataindex, 4);
        if (dl == 45 || dl == 46) {
            if (index <= dl - 5) check ^= data;
        } else {
            if (index <= dl - 10) { check ^= data; sum += data; }
        }
        index += 5;
    }
    if (check != 0) return 0;
    if (dl < 45 || dl > 46) {
        int data = rev4(in, dataindex, 4);
        if (data != (sum & 0x0F)) return 0;
    }
    ds += 1;
    int o = 0;
#define REV4_OUT(off) do { for (int q = 3; q >= 0; q--) out[o++] = in[ds + (off) + q]; } while (0)
    REV4_OUT(5); REV4_OUT(0); REV4_OUT(15); REV4_OUT(10);
    if (typ == 0 || typ == 2) {
        REV4_OUT(20);
    } else if (typ == 1 || typ == 3 || typ == 4 || typ == 7) {
        REV4_OUT(25); REV4_OUT(20); REV4_OUT(35); REV4_OUT(30);
        if (typ == 4) { REV4_OUT(55); REV4_OUT(50); REV4_OUT(45); REV4_OUT(40); }
    }
#undef REV4_OUT
    *no = o;
    return 1;
}

/* postdemodulation.py:580-640 */
static int pd_ws7035(const uint8_t *in, int n, uint8_t *out, int *no)
{
    if (n < 8 || bits_find(in, 8, "10100000") != 0) return 0;
    if (n != 44) return 0;
    int par = 0;
    for (int i = 15; i < 28; i++) par += in[i];
    if (par % 2) return 0;
    int s = 0;
    for (int i = 0; i < 40; i += 4) s += bits_val(in, i, 4);
    if (s % 16 != bits_val(in, 40, 4)) return 0;
    int o = 0;
    for (int i = 0; i < 44; i++) if (!(27 <= i && i < 31)) out[o++] = in[i];
    *no = o;
    return 1;
}

/* postdemodulation.py:642-706 */
static int pd_ws7053(const uint8_t *in, int n, uint8_t *out, int *no)
{
    uint8_t s[SDB_MAX_DIGITS + 64];
    int sp = bits_find(in, n, "10100000");
    int len = n;
    if (sp > 0) {
        len = n - sp;
        memcpy(s, in + sp, len);
        s[len++] = 0;
    } else {
        memcpy(s, in, n);
    }
    if (sp < 0) return 0;
    if (len < 32) return 0;
    int par = 0;
    for (int i = 15; i < 28; i++) par += s[i];
    if (par % 2) return 0;
    int o = 0;
    for (int i = 0; i < 28; i++) out[o++] = s[i];
    for (int i = 16; i < 24; i++) out[o++] = s[i];
    for (int i = 28; i < 32; i++) out[o++] = s[i];
    *no = o;
    return 1;
}

/* postdemodulation.py:708-730 */
static int pd_lenprefix(const uint8_t *in, int n, uint8_t *out, int *no)
{
    int nb = 8;
    while ((n >> nb) != 0) nb++;                   /* format(len,'08b') grows past 255 */
    int o = 0;
    for (int i = nb - 1; i >= 0; i--) out[o++] = (n >> i) & 1;
    memcpy(out + o, in, n);
    *no = o + n;
    return 1;
}

int ora_postdemod(int method, const uint8_t *in, int n, uint8_t *out, int out_cap, int *n_out)
{
    (void)out_cap;
    *n_out = 0;
    switch (method) {
    case ORA_PD_EM:           return pd_em(in, n, out, n_out);
    case ORA_PD_REVOLT:       return pd_revolt(in, n, out, n_out);
    case ORA_PD_FS20:         return pd_fs20(in, n, out, n_out);
    case ORA_PD_FHT80:        return pd_fht80(in, n, out, n_out);
    case ORA_PD_FHT80TF:      return pd_fht80tf(in, n, out, n_out);
    case ORA_PD_WS2000:       return pd_ws2000(in, n, out, n_out);
    case ORA_PD_WS7035:       return pd_ws7035(in, n, out, n_out);
    case ORA_PD_WS7053:       return pd_ws7053(in, n, out, n_out);
    case ORA_PD_LENGTHPREFIX: return pd_lenprefix(in, n, out, n_out);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * growable per-thread output
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    OraHit *hits; int64_t nh, caph;
    char *pool;   int64_t np, capp;
} OutBuf;

static void ob_hit(OutBuf *o, int msg, int proto, int bit_length, const char *payload)
{
    int len = (int)strlen(payload);
    if (o->nh == o->caph) { o->caph = o->caph ? o->caph * 2 : 1024; o->hits = realloc(o->hits, o->caph * sizeof(OraHit)); }
    if (o->np + len > o->capp) { while (o->np + len > o->capp) o->capp = o->capp ? o->capp * 2 : 65536; o->pool = realloc(o->pool, o->capp); }
    OraHit *h = &o->hits[o->nh++];
    h->msg = msg; h->proto = proto; h->bit_length = bit_length;
    h->payload_off = (int32_t)o->np; h->payload_len = len;
    memcpy(o->pool + o->np, payload, len);
    o->np += len;
}

/* ------------------------------------------------------------------------------------------
 * message view: packed record -> the strings / floats the reference works on
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    char   D[SDB_MAX_DIGITS + 8];
    int    dlen;
    int    npat;
    char   ids[SDB_MAX_SLOTS];
    double raw[SDB_MAX_SLOTS];
    int    cp;        /* slot index or -1 */
    int    valid;
} MsgView;

static void view_msg(const SdbPulseMsg *m, const uint8_t *digits, MsgView *v)
{
    const uint8_t *d = digits + (size_t)m->doff * 16;
    v->dlen = m->dlen;
    for (int i = 0; i < m->dlen; i++) {
        int nib = (d[i >> 1] >> ((i & 1) * 4)) & 0xF;
        v->D[i] = nib <= 9 ? (char)('0' + nib) : '?';
    }
    v->D[m->dlen] = 0;
    v->npat = m->npat;
    for (int s = 0; s < m->npat; s++) {
        v->ids[s] = (char)('0' + ((m->pat_ids >> (4 * s)) & 0xF));
        v->raw[s] = (double)m->pat[s];
    }
    v->cp = m->cp == 0xFF ? -1 : m->cp;
    v->valid = (m->flags & SDB_MSG_VALID) != 0;
}

/* tiny ordered dict: str -> repr char ('\0' = '') */
typedef struct { char key[8][ORA_MAXLIST + 1]; char val[8]; int n; } Lookup;
static int lk_find(const Lookup *l, const char *k)
{
    for (int i = 0; i < l->n; i++) if (strcmp(l->key[i], k) == 0) return i;
    return -1;
}
static void lk_set(Lookup *l, const char *k, char v)      /* d[k] = v */
{
    int i = lk_find(l, k);
    if (i < 0) { i = l->n++; strcpy(l->key[i], k); }
    l->val[i] = v;
}
static void lk_setdefault(Lookup *l, const char *k, char v)  /* if k not in d: d[k] = v */
{
    if (lk_find(l, k) < 0) lk_set(l, k, v);
}

/* helpers.length_in_range — helpers.py:124-166 */
static int length_in_range(const OraProto *p, int n)
{
    int min_len = p->has_length_min ? p->length_min : -1;
    if (min_len != -1 && n < min_len) return 0;
    if (p->has_length_max && n > p->length_max) return 0;
    return 1;
}

#define MAXBITS (SDB_MAX_DIGITS + 64)

/* ------------------------------------------------------------------------------------------
 * demodulate_ms — message_synced.py:10-243.  Returns SDB_ST_*.
 * ---------------------------------------------------------------------------------------- */
static int demod_ms_one(const OraProto *tab, int np, const MsgView *m, int mi, OutBuf *o)
{
    if (!m->valid) return SDB_ST_OK;                         /* :21-47 gates (host) */
    if (m->cp < 0) return SDB_ST_OK;                         /* :60-62 */
    double clock_abs = fabs(m->raw[m->cp]);                  /* :64 */
    if (clock_abs == 0) return SDB_ST_OK;
    double norm[SDB_MAX_SLOTS];
    for (int s = 0; s < m->npat; s++) norm[s] = ora_round1(m->raw[s] / clock_abs);   /* :70-72 */

    for (int pi = 0; pi < np; pi++) {                        /* :79-81 get_keys('sync') */
        const OraProto *p = &tab[pi];
        if (p->sync_kind == 0) continue;
        double proto_clock = p->has_clockabs ? p->clockabs : 0.0;   /* :83 */
        if (proto_clock > 0 && fabs(proto_clock - clock_abs) > clock_abs * 0.3) continue;   /* :84-88 */

        Lookup lk = {.n = 0}, el = {.n = 0};
        int message_start = 0, failed = 0;
        int w = p->none > 0 ? p->none : 0;                    /* :106-107 */
        const double *lists[4] = {p->sync, p->one, p->zero, p->flt};
        const int lens[4] = {p->sync_kind == 3 ? 0 : p->nsync, p->none, p->nzero, p->nfloat};
        const char reprs[4] = {0, '1', '0', 'F'};
        for (int key = 0; key < 4; key++) {                   /* :109 */
            if (lens[key] <= 0 && !(key == 0 && p->sync_kind == 2)) continue;   /* :111 falsy */
            if (key == 0 && p->sync_kind == 2) { failed = 1; break; }            /* :114-118 float('D') */
            char pstr[ORA_MAXLIST + 1];
            int r = ora_pattern_exists(lists[key], lens[key], m->ids, norm, m->npat, m->D, pstr);   /* :128 */
            if (r >= 0) {
                lk_set(&lk, pstr, reprs[key]);                /* :133 */
                if (r > 0) {                                  /* :135-138 */
                    char sh[ORA_MAXLIST + 1];
                    memcpy(sh, pstr, r - 1); sh[r - 1] = 0;
                    lk_setdefault(&el, sh, reprs[key]);
                }
                if (key == 0) {                               /* :140-158 */
                    const char *f = strstr(m->D, pstr);
                    if (!f) { failed = 1; break; }
                    message_start = (int)(f - m->D) + r;
                    double bit_length = w > 0 ? (double)(m->dlen - message_start) / (double)w : 0.0;
                    int length_min = p->has_length_min ? p->length_min : -1;
                    if ((double)length_min > bit_length) { failed = 1; break; }
                    el.n = 0;
                }
            } else if (key != 3) { failed = 1; break; }       /* :160-163 */
        }
        if (failed) continue;
        if (lk.n == 0) continue;                              /* :168 */
        if (w == 0) return SDB_ST_VALUEERROR;                 /* :174 range(start, stop, 0) */

        char bits[MAXBITS];
        int nb = 0;
        for (int i = message_start; i < m->dlen; i += w) {    /* :174-189 */
            char chunk[ORA_MAXLIST + 1];
            int cl = m->dlen - i < w ? m->dlen - i : w;
            memcpy(chunk, m->D + i, cl); chunk[cl] = 0;
            int k = lk_find(&lk, chunk);
            if (k >= 0) {
                if (lk.val[k]) bits[nb++] = lk.val[k];
            } else if (p->reconstruct) {
                if (cl == w) chunk[cl - 1] = 0;               /* :182 */
                int e = lk_find(&el, chunk);
                if (e >= 0) bits[nb++] = el.val[e]; else break;
            } else break;
        }
        if (nb == 0) continue;                                /* :191 */
        if (!length_in_range(p, nb)) continue;                /* :194 */
        while (nb % p->paddingbits > 0) bits[nb++] = '0';     /* :198-200 (pad BEFORE postDemod) */
        if (p->postdemod) {                                   /* :203-219 */
            uint8_t in[MAXBITS], out[MAXBITS + 16];
            for (int i = 0; i < nb; i++) {
                if (bits[i] != '0' && bits[i] != '1') return SDB_ST_VALUEERROR;   /* :209 int('F') */
                in[i] = (uint8_t)(bits[i] - '0');
            }
            int no = 0;
            int rc = ora_postdemod(p->postdemod, in, nb, out, sizeof out, &no);
            if (rc == -2) return SDB_ST_VALUEERROR;
            if (rc < 1) continue;
            if (no > 0) { for (int i = 0; i < no; i++) bits[i] = (char)('0' + out[i]); nb = no; }   /* :218 */
        }
        bits[nb] = 0;
        char hex[MAXBITS / 4 + 4];
        if (ora_bin2hex(bits, hex) < 0) continue;             /* :224-226 */
        char payload[MAXBITS / 4 + 64];
        snprintf(payload, sizeof payload, "%s%s%s", p->preamble, hex, p->postamble);
        ob_hit(o, mi, pi, nb, payload);                       /* :233-241 */
    }
    return SDB_ST_OK;
}

/* ------------------------------------------------------------------------------------------
 * demodulate_mu — message_unsynced.py:11-296.  Returns SDB_ST_*.
 * ---------------------------------------------------------------------------------------- */
static int starts_with(const char *s, const char *pre)
{
    while (*pre) { if (*s++ != *pre++) return 0; }
    return 1;
}

static int demod_mu_one(const OraProto *tab, int np, const regex_t *mm, const MsgView *m, int mi, OutBuf *o)
{
    if (m->dlen == 0) return SDB_ST_OK;                      /* :22-25 */
    for (int pi = 0; pi < np; pi++) {                        /* :45-47 get_keys('clockabs') */
        const OraProto *p = &tab[pi];
        if (!p->has_clockabs) continue;
        if (!p->active) continue;                            /* :48 */
        double clock_abs = p->clockabs;                      /* :59 */
        double norm[SDB_MAX_SLOTS];
        for (int s = 0; s < m->npat; s++) norm[s] = ora_round1(m->raw[s] / clock_abs);   /* :62-64 */

        const char *D = m->D;
        char start_str[ORA_MAXLIST + 1] = "";
        if (p->start_is_list) {                              /* :71-88 */
            int r = ora_pattern_exists(p->start, p->nstart, m->ids, norm, m->npat, D, start_str);
            if (r < 0) continue;
            const char *f = strstr(D, start_str);
            if (!f) continue;
            D = f;                                           /* current_raw_data[message_start:] */
        }
        int dlen = (int)strlen(D);

        Lookup lk = {.n = 0}, el = {.n = 0};
        char parts[3][ORA_MAXLIST + 1];
        int nparts = 0, failed = 0;
        const double *lists[3] = {p->one, p->zero, p->flt};
        const int lens[3] = {p->none, p->nzero, p->nfloat};
        const char reprs[3] = {'1', '0', 'F'};
        for (int key = 0; key < 3; key++) {                  /* :99-141 */
            if (lens[key] <= 0) continue;
            char pstr[ORA_MAXLIST + 1];
            int r = ora_pattern_exists(lists[key], lens[key], m->ids, norm, m->npat, D, pstr);
            if (r >= 0) {
                lk_set(&lk, pstr, reprs[key]);
                if (r > 0) {
                    char sh[ORA_MAXLIST + 1];
                    memcpy(sh, pstr, r - 1); sh[r - 1] = 0;
                    lk_setdefault(&el, sh, reprs[key]);
                }
                strcpy(parts[nparts++], pstr);
            } else if (key != 2) { failed = 1; break; }
        }
        if (failed || nparts == 0) continue;                 /* :143 */

        int use_tail = p->reconstruct && el.n > 0;           /* :175 */
        int length_min = p->has_length_min ? p->length_min : 0;   /* :178 */
        int w = p->none > 0 ? p->none : 0;                   /* :201-203 */
        int ls = (int)strlen(start_str);

        /* re.finditer("(?:START)((?:S1|S2|S3){MIN,}(?:E1|E2|..)?)", D)  :181-192 */
        int pos = 0;
        while (pos <= dlen) {
            int ms = -1, cs = 0, ce = 0;
            for (int s = pos; s <= dlen; s++) {
                if (!starts_with(D + s, start_str)) continue;
                int q = s + ls, n = 0;
                for (;;) {
                    int hit = 0;
                    for (int a = 0; a < nparts; a++)
                        if (starts_with(D + q, parts[a])) { q += (int)strlen(parts[a]); hit = 1; break; }
                    if (!hit) break;
                    n++;
                }
                if (n < length_min) continue;
                if (use_tail)
                    for (int a = 0; a < el.n; a++)
                        if (starts_with(D + q, el.key[a])) { q += (int)strlen(el.key[a]); break; }
                ms = s; cs = s + ls; ce = q;
                break;
            }
            if (ms < 0) break;
            pos = (ce == ms) ? ce + 1 : ce;                  /* empty match: scanner must advance */

            /* ---- one match: :193-290 ---- */
            if (w == 0) continue;                            /* :205 */
            int caplen = ce - cs;
            int nchunks = (caplen + w - 1) / w;
            if (nchunks == 0) return SDB_ST_INDEXERROR;      /* :212 chunks[-1] */
            if (p->length_max_truthy && nchunks > p->length_max) continue;   /* :217 */
            char bits[MAXBITS + 16];
            int nb = 0;
            for (int c = 0; c < nchunks; c++) {              /* :221-228 */
                char chunk[ORA_MAXLIST + 1];
                int cl = caplen - c * w < w ? caplen - c * w : w;
                memcpy(chunk, D + cs + c * w, cl); chunk[cl] = 0;
                int k = lk_find(&lk, chunk);
                if (k >= 0) bits[nb++] = lk.val[k];
                else if (p->reconstruct && (k = lk_find(&el, chunk)) >= 0) bits[nb++] = el.val[k];
            }
            if (p->postdemod) {                              /* :231-250 */
                uint8_t in[MAXBITS], out[MAXBITS + 16];
                int ok = 1;
                for (int i = 0; i < nb; i++) {
                    if (bits[i] != '0' && bits[i] != '1') { ok = 0; break; }   /* ValueError -> pass */
                    in[i] = (uint8_t)(bits[i] - '0');
                }
                if (ok) {
                    int no = 0;
                    int rc = ora_postdemod(p->postdemod, in, nb, out, sizeof out, &no);
                    if (rc != -2) {                          /* -2: ValueError swallowed at :249 */
                        if (rc < 1) continue;
                        for (int i = 0; i < no; i++) bits[i] = (char)('0' + out[i]);
                        nb = no;
                    }
                }
            }
            while (nb % p->paddingbits > 0) bits[nb++] = '0';   /* :257-259 (pad AFTER postDemod) */
            bits[nb] = 0;
            char dmsg[MAXBITS + 16];
            if (p->dispatch_bin) {                           /* :264-265 */
                strcpy(dmsg, bits);
            } else {
                if (ora_bin2hex(bits, dmsg) < 0) {
                    /* None: remove_zero would raise AttributeError (:269) — not reachable with the shipped table */
                    strcpy(dmsg, "None");
                } else if (p->remove_zero) {
                    int z = 0;
                    while (dmsg[z] == '0') z++;
                    memmove(dmsg, dmsg + z, strlen(dmsg + z) + 1);
                }
            }
            char payload[MAXBITS + 96];
            snprintf(payload, sizeof payload, "%s%s%s", p->preamble, dmsg, p->postamble);
            if (p->has_modulematch && regexec(&mm[pi], payload, 0, NULL, 0) != 0) continue;   /* :277-280 */
            ob_hit(o, mi, pi, nb, payload);                  /* :282-290 */
        }
    }
    return SDB_ST_OK;
}

/* ------------------------------------------------------------------------------------------
 * batch drivers
 * ---------------------------------------------------------------------------------------- */
static int merge_out(OutBuf *bufs, int nt, OraHit *hits, int64_t hits_cap, char *pool, int64_t pool_cap,
                     int64_t *nhits, int64_t *pool_used)
{
    int64_t th = 0, tp = 0;
    for (int t = 0; t < nt; t++) { th += bufs[t].nh; tp += bufs[t].np; }
    *nhits = th; *pool_used = tp;
    int rc = 0;
    if (th > hits_cap || tp > pool_cap) rc = -3;
    else {
        int64_t ho = 0, po = 0;
        for (int t = 0; t < nt; t++) {
            for (int64_t i = 0; i < bufs[t].nh; i++) {
                hits[ho] = bufs[t].hits[i];
                hits[ho].payload_off += (int32_t)po;
                ho++;
            }
            if (bufs[t].np) memcpy(pool + po, bufs[t].pool, bufs[t].np);
            po += bufs[t].np;
        }
    }
    for (int t = 0; t < nt; t++) { free(bufs[t].hits); free(bufs[t].pool); }
    return rc;
}

typedef struct {
    const OraProto *tab; int nproto; int kind; const regex_t *mm;
    const SdbPulseMsg *msgs; const uint8_t *digits; int64_t lo, hi;
    uint8_t *status; OutBuf *buf;
} PulseJob;

static void *pulse_worker(void *arg)
{
    PulseJob *j = arg;
    MsgView v;
    for (int64_t i = j->lo; i < j->hi; i++) {
        view_msg(&j->msgs[i], j->digits, &v);
        int64_t nh0 = j->buf->nh, np0 = j->buf->np;
        int st = j->kind == SDB_KIND_MS ? demod_ms_one(j->tab, j->nproto, &v, (int)i, j->buf)
                                        : demod_mu_one(j->tab, j->nproto, j->mm, &v, (int)i, j->buf);
        if (st != SDB_ST_OK) { j->buf->nh = nh0; j->buf->np = np0; }   /* exception: results lost */
        j->status[i] = (uint8_t)st;
    }
    return NULL;
}

int ora_demod_pulse(const OraProto *tab, int nproto, int kind,
                    const SdbPulseMsg *msgs, const uint8_t *digits, int64_t n,
                    uint8_t *status, OraHit *hits, int64_t hits_cap,
                    char *pool, int64_t pool_cap, int64_t *nhits, int64_t *pool_used,
                    int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    regex_t *mm = calloc(nproto, sizeof(regex_t));
    for (int i = 0; i < nproto; i++)
        if (tab[i].has_modulematch && regcomp(&mm[i], tab[i].modulematch, REG_EXTENDED | REG_NOSUB) != 0) {
            fprintf(stderr, "oracle: cannot compile modulematch '%s'\n", tab[i].modulematch);
            free(mm);
            return -1;
        }
    OutBuf *bufs = calloc(nthreads, sizeof(OutBuf));
    PulseJob *jobs = calloc(nthreads, sizeof(PulseJob));
    pthread_t *th = calloc(nthreads, sizeof(pthread_t));
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (PulseJob){tab, nproto, kind, mm, msgs, digits, n * t / nthreads, n * (t + 1) / nthreads,
                             status, &bufs[t]};
        if (nthreads == 1) pulse_worker(&jobs[t]);
        else pthread_create(&th[t], NULL, pulse_worker, &jobs[t]);
    }
    if (nthreads > 1) for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    int rc = merge_out(bufs, nthreads, hits, hits_cap, pool, pool_cap, nhits, pool_used);
    for (int i = 0; i < nproto; i++) if (tab[i].has_modulematch) regfree(&mm[i]);
    free(mm); free(bufs); free(jobs); free(th);
    return rc;
}
