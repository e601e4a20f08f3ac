/*
 * sdb_fmt.h — the payload string of one hit.  MS / MU: preamble + hex / bits + postamble
 * (sd_protocols/message_synced.py:224-231, message_unsynced.py:254-274, helpers.py:28-64); MC: preamble + hex, or preamble +
 * repr(list) for TFA (manchester.py:131-132, :713-717); MN: the converter's own string (helpers.py:223-716).
 * ONE definition for the host formatter (sdb_capi.cu, sdb_format_hits) and the device formatter (sdb_format.cu,
 * format_kernel), so the two cannot drift apart.
 */
#pragma once
#include <stdint.h>
#include "../../include/sdb200.h"
#include "sdb_table.h"

#ifdef __CUDACC__
#define SDB_HD __host__ __device__ __forceinline__
#else
#define SDB_HD static inline
#endif

SDB_HD int sdb_fmt_bit(const uint32_t *w, uint32_t i) { return (int)((w[i >> 5] >> (i & 31)) & 1u); }

/* hex digit j of nb bits (LSB-first words), right-aligned nibbles as helpers.py:28-64 builds them */
SDB_HD int sdb_fmt_hex_digit(const uint32_t *w, int nb, int nd, int j)
{
    const int b0 = nb - 4 * (nd - j);
    if (b0 >= 0) {                                    /* message bits b0 .. b0+3, the first one is the digit's MSB */
        const int sh = b0 & 31;
        uint32_t x = w[b0 >> 5] >> sh;
        if (sh > 28) x |= w[(b0 >> 5) + 1] << (32 - sh);      /* b0 + 3 < nb: that word exists */
        x &= 0xFu;
        return (int)(((x & 1u) << 3) | ((x & 2u) << 1) | ((x & 4u) >> 1) | ((x & 8u) >> 3));
    }
    int v = 0;
    for (int k = 0; k < 4; k++) { const int bi = b0 + k; v = (v << 1) | (bi >= 0 ? sdb_fmt_bit(w, (uint32_t)bi) : 0); }
    return v;
}

/* Characters of a pulse hit written to dst (dst == nullptr: count only); returns the count. */
SDB_HD uint32_t sdb_fmt_pulse(const SdbPulseProto *pp, const SdbHit &ht, const uint32_t *w, char *dst)
{
    uint32_t n = 0;
    const uint32_t nb = ht.nbits, nwv = (nb + 31) >> 5;
    const bool has_f = (ht.flags & SDB_HIT_HAS_F) != 0;
    for (int k = 0; k < pp->pre_len; k++) { if (dst) dst[n] = pp->preamble[k]; n++; }
    if (pp->flags & SDB_PF_DISPATCH_BIN) {
        if (dst) for (uint32_t b = 0; b < nb; b++) dst[n + b] = has_f && sdb_fmt_bit(w + nwv, b) ? 'F' : (char)('0' + sdb_fmt_bit(w, b));
        n += nb;
    } else if (has_f) {                               /* f"{None}" (message_unsynced.py:267,274) */
        if (dst) { dst[n] = 'N'; dst[n + 1] = 'o'; dst[n + 2] = 'n'; dst[n + 3] = 'e'; }
        n += 4;
    } else {
        const int nd = (int)((nb + 3) >> 2);
        int j = 0;
        if (pp->flags & SDB_PF_REMOVE_ZERO) while (j < nd && sdb_fmt_hex_digit(w, (int)nb, nd, j) == 0) j++;    /* lstrip('0') :268-269 */
        if (dst) for (int q = j; q < nd; q++) { const int d = sdb_fmt_hex_digit(w, (int)nb, nd, q); dst[n + (uint32_t)(q - j)] = (char)(d < 10 ? '0' + d : 'A' + d - 10); }
        n += (uint32_t)(nd - j);
    }
    for (int k = 0; k < pp->post_len; k++) { if (dst) dst[n] = pp->postamble[k]; n++; }
    return n;
}

/* plain (unstripped) hex digits of nb bits */
SDB_HD uint32_t sdb_fmt_hexbits(const uint32_t *w, uint32_t nb, char *dst)
{
    const int nd = (int)((nb + 3) >> 2);
    if (dst) for (int q = 0; q < nd; q++) { const int d = sdb_fmt_hex_digit(w, (int)nb, nd, q); dst[q] = (char)(d < 10 ? '0' + d : 'A' + d - 10); }
    return (uint32_t)nd;
}

/* "%u" */
SDB_HD uint32_t sdb_fmt_u32(uint32_t v, char *dst)
{
    char tmp[10];
    int k = 0;
    do { tmp[k++] = (char)('0' + v % 10); v /= 10; } while (v);
    if (dst) for (int i = 0; i < k; i++) dst[i] = tmp[k - 1 - i];
    return (uint32_t)k;
}

/* Characters of MC / MN hit i (hits of one TFA list are consecutive; the first element carries the whole string, the others
 * are empty); dst == nullptr: count only. */
SDB_HD uint32_t sdb_fmt_hexkind(int kind, const SdbHexProto *hx, const SdbHit *hits, uint32_t i, uint32_t nhits, const uint32_t *bits, char *dst)
{
    const SdbHit &ht = hits[i];
    const SdbHexProto &p = hx[ht.proto];
    const uint32_t *w = bits + ht.bits_off;
    uint32_t n = 0;
#define SDB_PUT(c) do { if (dst) dst[n] = (c); n++; } while (0)
    if (kind == SDB_KIND_MC) {
        if ((ht.flags & SDB_HIT_LIST) && ht.aux != 0) return 0;
        for (int k = 0; k < p.pre_len; k++) SDB_PUT(p.preamble[k]);
        if (ht.flags & SDB_HIT_LIST) {
            SDB_PUT('[');
            for (uint32_t e = i; e < nhits && hits[e].msg == ht.msg && (hits[e].flags & SDB_HIT_LIST) && hits[e].aux == e - i; e++) {
                if (e > i) { SDB_PUT(','); SDB_PUT(' '); }
                SDB_PUT('\'');
                n += sdb_fmt_hexbits(bits + hits[e].bits_off, hits[e].nbits, dst ? dst + n : nullptr);
                SDB_PUT('\'');
            }
            SDB_PUT(']');
        } else if (ht.flags & SDB_HIT_HAS_F) {            /* f"{preamble}{None}" (as-shipped mcRaw) */
            SDB_PUT('N'); SDB_PUT('o'); SDB_PUT('n'); SDB_PUT('e');
        } else n += sdb_fmt_hexbits(w, ht.nbits, dst ? dst + n : nullptr);
    } else if (ht.flags & SDB_HIT_FIELDS) {               /* MN hits carry the converter that produced them in aux */
        const bool pca = ht.aux == SDB_M_PCA301;          /* "OK 24 %u x10 %04X" (helpers.py:525-579) | "OK 9 %u x5" (:630-716) */
        SDB_PUT('O'); SDB_PUT('K'); SDB_PUT(' ');
        if (pca) { SDB_PUT('2'); SDB_PUT('4'); } else SDB_PUT('9');
        const int nf = pca ? 10 : 5;
        for (int f = 0; f < nf; f++) { SDB_PUT(' '); n += sdb_fmt_u32(w[f], dst ? dst + n : nullptr); }
        if (pca) {
            SDB_PUT(' ');
            for (int q = 3; q >= 0; q--) { const uint32_t d = (w[10] >> (4 * q)) & 0xFu; SDB_PUT((char)(d < 10 ? '0' + d : 'A' + d - 10)); }
        }
    } else {
        if (ht.aux == SDB_M_KOPP) { SDB_PUT('k'); SDB_PUT('r'); }
        n += sdb_fmt_hexbits(w, ht.nbits, dst ? dst + n : nullptr);
    }
#undef SDB_PUT
    return n;
}
