/*
 * sdb_fmt.h — the payload string of one MS / MU hit: preamble + hex / bits + postamble
 * (sd_protocols/message_synced.py:224-231, message_unsynced.py:254-274, helpers.py:28-64).
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
