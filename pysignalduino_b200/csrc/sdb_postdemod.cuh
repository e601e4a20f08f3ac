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
/* int(bits[from:from+n], 2) — MSB first */
__device__ __forceinline__ int bval(const uint32_t *w, int from, int n)
{
    int v = 0;
    for (int i = 0; i < n; i++) v = (v << 1) | gbit(w, from + i);
    return v;
}
/* int("".join(reversed(bits[from:from+n])), 2) */
__device__ __forceinline__ int bval_rev(const uint32_t *w, int from, int n)
{
    int v = 0;
    for (int i = n - 1; i >= 0; i--) v = (v << 1) | gbit(w, from + i);
    return v;
}
/* str.find of an MSB-first pattern of `m` bits */
__device__ __forceinline__ int bfind(const uint32_t *w, int n, uint32_t pat, int m)
{
    for (int i = 0; i + m <= n; i++)
        if ((uint32_t)bval(w, i, m) == pat) return i;
    return -1;
}
__device__ __forceinline__ int first_one(const uint32_t *w, int n)
{
    for (int i = 0; i < n; i++) if (gbit(w, i)) return i;
    return -1;
}
/* every 9-bit group (last may be short) has even parity */
__device__ __forceinline__ bool parity9_ok(const uint32_t *w, int base, int len)
{
    for (int s = 0; s < len; s += 9) {
        int p = 0;
        for (int i = s; i < s + 9 && i < len; i++) p ^= gbit(w, base + i);
        if (p) return false;
    }
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
    for (int c = 0; c < len; c += 9) {
        if (c + 8 < len) {
            int byte = bval(in, base + c, 8);
            if (c < len - 10) {
                for (int j = 7; j >= 0; j--) sbit(out, k++, gbit(in, base + c + j));
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
    for (int b = 0; b < 88; b += 8) sum += bval(in, b, 8);
    if ((sum & 0xFF) != chk) return 0;
    for (int i = 0; i < 88; i++) sbit(out, i, gbit(in, i));
    *no = 88;
    return 1;
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
    for (int i = 0; i < len - 9; i += 9) sum += bval(in, base + i, 8);
    int chk = bval(in, base + len - 9, 8);
    if (((sum + 6) & 0xFF) == chk) return 0;
    if ((sum & 0xFF) != chk) return 0;
    if (!parity9_ok(in, base, len)) return 0;
    /* strip every 9th bit -> t[]; 45: t[0:24] + 8 zeros + t[24:32]; 54: t[0:40] */
    int o = 0, k = 0;
    for (int i = 0; i < len; i++) {
        if (i % 9 == 8) continue;
        int b = gbit(in, base + i);
        if (len == 45) {
            if (k == 24) for (int z = 0; z < 8; z++) sbit(out, o++, 0);
            if (k < 32) sbit(out, o++, b);
        } else {
            if (k < 40) sbit(out, o++, b);
        }
        k++;
    }
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
    for (int i = 0; i < 45; i += 9) sum += bval(in, base + i, 8);
    int chk = bval(in, base + 45, 8);
    if (((sum - 6) & 0xFF) == chk) return 0;
    if ((sum & 0xFF) != chk) return 0;
    if (!parity9_ok(in, base, 54)) return 0;
    int o = 0;
    for (int i = 0; i < 54; i++) if (i % 9 != 8) sbit(out, o++, gbit(in, base + i));
    *no = o;
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
    for (int i = 0; i < 36; i += 9) sum += bval(in, base + i, 8);
    if ((sum & 0xFF) != bval(in, base + 36, 8)) return 0;
    if (!parity9_ok(in, base, 45)) return 0;
    int o = 0, k = 0, bit26 = 0;
    for (int i = 0; i < 45; i++) {
        if (i % 9 == 8) continue;
        int b = gbit(in, base + i);
        if (k == 26) bit26 = b;
        if (k < 32) sbit(out, o++, b);            /* del [32:40] */
        k++;
    }
    if (bit26 != 0) return 0;
    *no = o;
    return 1;
}

/* postDemo_WS2000 — :425-578 */
__device__ inline int pd_ws2000(const uint32_t *in, int n, uint32_t *out, int *no)
{
    const int dlw[8] = {35, 50, 35, 50, 70, 40, 40, 85};
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
    if (dlw[typ] != dl1) return 0;
    if (ds > 10) return 0;
    int index = 0, dataindex = 0, check = 0, sum = 5;
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
#define SDB_REV4(off) do { for (int q = 3; q >= 0; q--) sbit(out, o++, gbit(in, ds + (off) + q)); } while (0)
    SDB_REV4(5); SDB_REV4(0); SDB_REV4(15); SDB_REV4(10);
    if (typ == 0 || typ == 2) {
        SDB_REV4(20);
    } else if (typ == 1 || typ == 3 || typ == 4 || typ == 7) {
        SDB_REV4(25); SDB_REV4(20); SDB_REV4(35); SDB_REV4(30);
        if (typ == 4) { SDB_REV4(55); SDB_REV4(50); SDB_REV4(45); SDB_REV4(40); }
    }
#undef SDB_REV4
    *no = o;
    return 1;
}

/* postDemo_WS7035 — :580-640 */
__device__ inline int pd_ws7035(const uint32_t *in, int n, uint32_t *out, int *no)
{
    if (n < 8 || bval(in, 0, 8) != 0xA0) return 0;   /* startswith('10100000') */
    if (n != 44) return 0;
    int par = 0;
    for (int i = 15; i < 28; i++) par ^= gbit(in, i);
    if (par) return 0;
    int s = 0;
    for (int i = 0; i < 40; i += 4) s += bval(in, i, 4);
    if ((s & 15) != bval(in, 40, 4)) return 0;
    int o = 0;
    for (int i = 0; i < 44; i++) if (!(27 <= i && i < 31)) sbit(out, o++, gbit(in, i));
    *no = o;
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
    /* bit i of the working string: i < n-base ? in[base+i] : 0 */
#define SDB_WS(i) (((i) < n - base) ? gbit(in, base + (i)) : 0)
    int par = 0;
    for (int i = 15; i < 28; i++) par ^= SDB_WS(i);
    if (par) return 0;
    int o = 0;
    for (int i = 0; i < 28; i++) sbit(out, o++, SDB_WS(i));
    for (int i = 16; i < 24; i++) sbit(out, o++, SDB_WS(i));
    for (int i = 28; i < 32; i++) sbit(out, o++, SDB_WS(i));
#undef SDB_WS
    *no = o;
    return 1;
}

/* postDemo_lengtnPrefix — :708-730 */
__device__ inline int pd_lenprefix(const uint32_t *in, int n, uint32_t *out, int *no)
{
    int nb = 8;
    while ((n >> nb) != 0) nb++;
    int o = 0;
    for (int i = nb - 1; i >= 0; i--) sbit(out, o++, (n >> i) & 1);
    for (int i = 0; i < n; i++) sbit(out, o++, gbit(in, i));
    *no = o;
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
