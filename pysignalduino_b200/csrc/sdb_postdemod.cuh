/*
 * sdb_postdemod.cuh — the nine postDemo_* validators / rewriters as device functions.
 *
 * Reference: sd_protocols/postdemodulation.py:27-730.  Input and output are bit strings packed
 * LSB-first into 32-bit words (bit i of the message = word i>>5, bit i&31).  Each function is
 * executed by ONE lane (frames are <= ~150 bits and only ~10 of 129 protocols have a
 * post-demodulation step); return codes: 1 accept (out/no filled), 0 reject,
 * -2 the reference raises ValueError (postdemodulation.py:471 int('', 2)).
 */
#pragma once
#include <stdint.h>
#include "sdb_table.h"

namespace sdb {

__device__ __forceinline__ int gbit(const uint32_t *w, int i) { return (w[i >> 5] >> (i & 31)) & 1; }
__device__ __forceinline__ void sbit(uint32_t *w, int i, int v)
{
    uint32_t m = 1u << (i & 31);
    if (v) w[i >> 5] |= m; else w[i >> 5] &= ~m;
}

/* Word-level access (the validators used to walk bit by bit; unrolled, that was 48 KB of code inside kernels whose hot code
 * has to fit the instruction cache, and 19 % of the emit kernel's instructions on a single lane):
 * n <= 32 message bits starting at `from`, message bit from + i at bit i.  Reads w[from >> 5] and the word after it
 * (the bit planes have spare words). */
__device__ __forceinline__ uint32_t bits32(const uint32_t *w, int from, int n)
{
    const uint32_t x = __funnelshift_r(w[from >> 5], w[(from >> 5) + 1], from & 31);
    return n >= 32 ? x : x & ((1u << n) - 1u);
}
/* int(bits[from:from+n], 2) — MSB first, 0 <= n <= 32 */
__device__ __forceinline__ int bval(const uint32_t *w, int from, int n)
{
    return n <= 0 ? 0 : (int)(__brev(bits32(w, from, n)) >> (32 - n));
}
/* int("".join(reversed(bits[from:from+n])), 2) */
__device__ __forceinline__ int bval_rev(const uint32_t *w, int from, int n) { return n <= 0 ? 0 : (int)bits32(w, from, n); }
/* n <= 32 bits (bit i of v -> message bit o + i) ORed into a zero-initialised plane */
__device__ __forceinline__ void put_bits(uint32_t *w, int o, uint32_t v, int n)
{
    const int sh = o & 31;
    w[o >> 5] |= v << sh;
    if (sh + n > 32) w[(o >> 5) + 1] |= v >> (32 - sh);
}
/* out[o : o + n] = in[p : p + n] (out zero-initialised) */
__device__ __forceinline__ void copy_bits(uint32_t *out, int o, const uint32_t *in, int p, int n)
{
#pragma unroll 1
    for (int i = 0; i < n; i += 32) { const int c = n - i < 32 ? n - i : 32; put_bits(out, o + i, bits32(in, p + i, c), c); }
}
/* str.find of an MSB-first pattern of `m` bits */
__device__ __forceinline__ int bfind(const uint32_t *w, int n, uint32_t pat, int m)
{
#pragma unroll 1
    for (int i = 0; i + m <= n; i++)
        if ((uint32_t)bval(w, i, m) == pat) return i;
    return -1;
}
__device__ __forceinline__ int first_one(const uint32_t *w, int n)
{
#pragma unroll 1
    for (int i = 0; i < n; i += 32) {
        const uint32_t x = bits32(w, i, n - i < 32 ? n - i : 32);
        if (x) return i + __ffs(x) - 1;
    }
    return -1;
}
/* every 9-bit group (last may be short) has even parity */
__device__ __forceinline__ bool parity9_ok(const uint32_t *w, int base, int len)
{
#pragma unroll 1
    for (int s = 0; s < len; s += 9)
        if (__popc(bits32(w, base + s, len - s < 9 ? len - s : 9)) & 1) return false;
    return true;
}

/* postDemo_EM — postdemodulation.py:27-88 */
__device__ inline int pd_em(const uint32_t *in, int n, uint32_t *out, int *no)
{
    int st = bfind(in, n, 1u, 10);                 /* "0000000001" */
    if (st < 0) return 0;
    int base = st + 10, len = n - base;
    if (len != 89) return 0;
    int crc = 0, k = 0;
#pragma unroll 1
    for (int c = 0; c < len; c += 9) {
        if (c + 8 < len) {
            int byte = bval(in, base + c, 8);
            if (c < len - 10) {
                put_bits(out, k, (uint32_t)byte, 8);          /* the byte's bits in reverse order: as a number that is `byte` again */
                k += 8;
                crc ^= byte;
            }
        }
    }
    if (crc != bval(in, base + len - 8, 8)) return 0;
    *no = k;
    return 1;
}

/* postDemo_Revolt — :90-137 */
__device__ inline int pd_revolt(const uint32_t *in, int n, uint32_t *out, int *no)
{
    if (n < 96) return 0;
    int chk = bval(in, 88, 8), sum = 0;
#pragma unroll 1
    for (int b = 0; b < 88; b += 8) sum += bval(in, b, 8);
    if ((sum & 0xFF) != chk) return 0;
    copy_bits(out, 0, in, 0, 88);
    *no = 88;
    return 1;
}

/* groups of 9 bits -> their first 8 bits, `ngroups` of them, appended at out[o ...] */
__device__ __forceinline__ int strip_ninth(uint32_t *out, int o, const uint32_t *in, int base, int g0, int ngroups)
{
#pragma unroll 1
    for (int g = g0; g < g0 + ngroups; g++) { put_bits(out, o, bits32(in, base + 9 * g, 8), 8); o += 8; }
    return o;
}

/* postDemo_FS20 — :139-243 */
__device__ inline int pd_fs20(const uint32_t *in, int n, uint32_t *out, int *no)
{
    int ds = first_one(in, n);
    if (ds < 0) return 0;
    int base = ds + 1, len = n - base;
    if (len == 46 || len == 55) len--;
    if (len != 45 && len != 54) return 0;
    int sum = 6;
#pragma unroll 1
    for (int i = 0; i < len - 9; i += 9) sum += bval(in, base + i, 8);
    int chk = bval(in, base + len - 9, 8);
    if (((sum + 6) & 0xFF) == chk) return 0;
    if ((sum & 0xFF) != chk) return 0;
    if (!parity9_ok(in, base, len)) return 0;
    /* strip every 9th bit -> t[]; 45: t[0:24] + 8 zeros + t[24:32]; 54: t[0:40] */
    int o;
    if (len == 45) {
        o = strip_ninth(out, 0, in, base, 0, 3);
        o = strip_ninth(out, o + 8, in, base, 3, 1);
    } else o = strip_ninth(out, 0, in, base, 0, 5);
    *no = o;
    return 1;
}

/* postDemo_FHT80 — :245-337 */
__device__ inline int pd_fht80(const uint32_t *in, int n, uint32_t *out, int *no)
{
    int ds = first_one(in, n);
    if (ds < 0) return 0;
    int base = ds + 1, len = n - base;
    if (len == 55) len--;
    if (len != 54) return 0;
    int sum = 12;
#pragma unroll 1
    for (int i = 0; i < 45; i += 9) sum += bval(in, base + i, 8);
    int chk = bval(in, base + 45, 8);
    if (((sum - 6) & 0xFF) == chk) return 0;
    if ((sum & 0xFF) != chk) return 0;
    if (!parity9_ok(in, base, 54)) return 0;
    *no = strip_ninth(out, 0, in, base, 0, 6);
    return 1;
}

/* postDemo_FHT80TF — :339-423 */
__device__ inline int pd_fht80tf(const uint32_t *in, int n, uint32_t *out, int *no)
{
    if (n < 46) return 0;
    int ds = first_one(in, n);
    if (ds < 0) return 0;
    int base = ds + 1, len = n - base;
    if (len != 45) return 0;
    int sum = 12;
#pragma unroll 1
    for (int i = 0; i < 36; i += 9) sum += bval(in, base + i, 8);
    if ((sum & 0xFF) != bval(in, base + 36, 8)) return 0;
    if (!parity9_ok(in, base, 45)) return 0;
    if (gbit(in, base + 27 + 2) != 0) return 0;    /* bit 26 of the stripped string = bit 2 of the fourth group */
    *no = strip_ninth(out, 0, in, base, 0, 4);      /* del [32:40] */
    return 1;
}

/* postDemo_WS2000 — :425-578 */
__device__ inline int pd_ws2000(const uint32_t *in, int n, uint32_t *out, int *no)
{
    int ds = first_one(in, n);
    if (ds < 0) return 0;
    int dl = n - ds;
    int dl1 = dl - (dl % 5);
    int avail = n - (ds + 1);
    if (avail > 4) avail = 4;
    if (avail <= 0) return -2;                    /* int('', 2) -> ValueError */
    int typ = bval_rev(in, ds + 1, avail);
    if (typ > 7) return 0;
    if (typ == 1 && (dl == 45 || dl == 46)) dl1 += 5;
    /* datalength per sensor type: 35, 50, 35, 50, 70, 40, 40, 85 (in units of 5 bits, one byte each) */
    if ((int)((0x1108080E0A070A07ull >> (8 * typ)) & 0xFFull) * 5 != dl1) return 0;
    if (ds > 10) return 0;
    int index = 0, dataindex = 0, check = 0, sum = 5;
#pragma unroll 1
    while (index < dl - 1) {
        if (gbit(in, index + ds) != 1) return 0;
        dataindex = index + ds + 1;
        if (n - dataindex < 4) return 0;
        int data = bval_rev(in, dataindex, 4);
        if (dl == 45 || dl == 46) {
            if (index <= dl - 5) check ^= data;
        } else if (index <= dl - 10) { check ^= data; sum += data; }
        index += 5;
    }
    if (check != 0) return 0;
    if (dl < 45 || dl > 46) {
        if (bval_rev(in, dataindex, 4) != (sum & 0x0F)) return 0;
    }
    ds += 1;
    int o = 0;
    /* nibbles at the offsets 5, 0, 15, 10 (then 20 | 25, 20, 35, 30 (then 55, 50, 45, 40)), each with its 4 bits reversed;
     * ord = offset / 5 of the k-th nibble, one byte each, low byte first */
    uint64_t ord = 0x02030001ull;
    int cntn = 4;
    if (typ == 0 || typ == 2) { ord |= 4ull << 32; cntn = 5; }
    else if (typ == 1 || typ == 3 || typ == 4 || typ == 7) { ord |= 0x06070405ull << 32; cntn = 8; }
#pragma unroll 1
    for (int k = 0; k < cntn; k++) {
        const int off = 5 * (int)((ord >> (8 * k)) & 0xFFull);
        put_bits(out, o, __brev(bits32(in, ds + off, 4)) >> 28, 4);
        o += 4;
    }
    if (typ == 4) {
#pragma unroll 1
        for (int k = 11; k >= 8; k--) { put_bits(out, o, __brev(bits32(in, ds + 5 * k, 4)) >> 28, 4); o += 4; }
    }
    *no = o;
    return 1;
}

/* postDemo_WS7035 — :580-640 */
__device__ inline int pd_ws7035(const uint32_t *in, int n, uint32_t *out, int *no)
{
    if (n < 8 || bval(in, 0, 8) != 0xA0) return 0;   /* startswith('10100000') */
    if (n != 44) return 0;
    if (__popc(bits32(in, 15, 13)) & 1) return 0;
    int s = 0;
#pragma unroll 1
    for (int i = 0; i < 40; i += 4) s += bval(in, i, 4);
    if ((s & 15) != bval(in, 40, 4)) return 0;
    copy_bits(out, 0, in, 0, 27);                    /* without bits 27..30 */
    copy_bits(out, 27, in, 31, 13);
    *no = 40;
    return 1;
}

/* postDemo_WS7053 — :642-706 */
__device__ inline int pd_ws7053(const uint32_t *in, int n, uint32_t *out, int *no)
{
    int sp = bfind(in, n, 0xA0u, 8);
    if (sp < 0) return 0;
    int len = n, base = 0;
    if (sp > 0) { base = sp; len = n - sp + 1; }     /* cut, then append one '0' */
    if (len < 32) return 0;
    /* bits [i, i + c) of the working string: i < n-base ? in[base+i] : 0 */
    const int have = n - base;
#define SDB_WS(i, c) ((i) >= have ? 0u : bits32(in, base + (i), have - (i) < (c) ? have - (i) : (c)))
    if (__popc(SDB_WS(15, 13)) & 1) return 0;
    put_bits(out, 0, SDB_WS(0, 28), 28);
    put_bits(out, 28, SDB_WS(16, 8), 8);
    put_bits(out, 36, SDB_WS(28, 4), 4);
#undef SDB_WS
    *no = 40;
    return 1;
}

/* postDemo_lengtnPrefix — :708-730 */
__device__ inline int pd_lenprefix(const uint32_t *in, int n, uint32_t *out, int *no)
{
    int nb = 8;
    while ((n >> nb) != 0) nb++;
    put_bits(out, 0, __brev((uint32_t)n) >> (32 - nb), nb);      /* n as nb binary digits, most significant first */
    copy_bits(out, nb, in, 0, n);
    *no = nb + n;
    return 1;
}

__device__ inline int postdemod(int method, const uint32_t *in, int n, uint32_t *out, int *no)
{
    *no = 0;
    switch (method) {
    case SDB_PD_EM:           return pd_em(in, n, out, no);
    case SDB_PD_REVOLT:       return pd_revolt(in, n, out, no);
    case SDB_PD_FS20:         return pd_fs20(in, n, out, no);
    case SDB_PD_FHT80:        return pd_fht80(in, n, out, no);
    case SDB_PD_FHT80TF:      return pd_fht80tf(in, n, out, no);
    case SDB_PD_WS2000:       return pd_ws2000(in, n, out, no);
    case SDB_PD_WS7035:       return pd_ws7035(in, n, out, no);
    case SDB_PD_WS7053:       return pd_ws7053(in, n, out, no);
    case SDB_PD_LENGTHPREFIX: return pd_lenprefix(in, n, out, no);
    }
    return 0;
}

}  // namespace sdb
