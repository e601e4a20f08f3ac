/*
 * sdb_hex.cu — MC (Manchester) and MN batch kernels for sm_100a.
 *
 * Replaces, for a whole batch of packed messages,
 *   demodulate_mc / _demodulate_mc_data   sd_protocols/sd_protocols.py:76-111, manchester.py:49-144
 *   _convert_mc_hex_to_bits, hex_to_bin_str manchester.py:18-47, helpers.py:168-188
 *   mcBit2* / mcRaw / helpers.mcraw        manchester.py:207-795, helpers.py:90-122
 *   demodulate_mn + Conv*                  sd_protocols.py:113-155, helpers.py:190-716
 *
 * These messages are tiny (<= 512 hex characters, one named protocol each), so the mapping is
 * one THREAD per message: adjacent threads read adjacent 16-byte digit units of the pool, every
 * decoder works on a nibble accessor (no bit string is ever materialised), and an accepted message
 * publishes its hit with one atomicAdd.  mc_repaired selects the reference "as shipped"
 * (TypeError at manchester.py:84/:120) or with the two documented one-line repairs.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sdb200.h"
#include "sdb_table.h"
#include "sdb_pulse.h"

namespace sdb {

#define RES_WORDS (SDB_MAX_HEX * 4 / 32 + 2)

struct HArgs {
    SdbDevTable tab;
    const SdbHexMsg *msgs;
    const uint8_t *digits;
    uint32_t n;
    int kind, repaired;
    SdbMsgOut *out;
    SdbHit *hits;   uint32_t hits_cap;
    uint32_t *bits; uint32_t bits_cap;
    SdbCounters *ctr;
};

/* hex nibble i of a message */
struct Nibs {
    const uint8_t *d;
    int n;
    __device__ __forceinline__ int at(int i) const { return (__ldg(&d[i >> 1]) >> ((i & 1) * 4)) & 0xF; }
    __device__ __forceinline__ int byte(int i) const { return (at(2 * i) << 4) | at(2 * i + 1); }
};

/* the bit string bin(int(hex,16)).zfill(4k) of the (optionally inverted) hex: leading zero nibbles dropped */
struct McBits {
    Nibs h;
    int lead;      /* dropped nibbles */
    int n;         /* bits */
    bool inv;
    __device__ __forceinline__ int nib(int i) const { int x = h.at(lead + i); return inv ? 15 - x : x; }
    __device__ __forceinline__ int bit(int i) const { return (nib(i >> 2) >> (3 - (i & 3))) & 1; }
};

struct Res {
    uint32_t w[RES_WORDS];
    int n;
    __device__ __forceinline__ void clear() { for (int i = 0; i < RES_WORDS; i++) w[i] = 0; n = 0; }
    __device__ __forceinline__ void push(int b) { if (b) w[n >> 5] |= 1u << (n & 31); n++; }
};

__device__ void publish(const HArgs &A, SdbMsgOut &mo, uint32_t mi, uint16_t proto, const uint32_t *w, int nbits,
                        int nwords, uint8_t flags, uint16_t aux, uint32_t hb, uint32_t k)
{
    uint32_t wb = atomicAdd(&A.ctr->words, (uint32_t)nwords);
    if (hb + k < A.hits_cap && wb + nwords <= A.bits_cap) {
        SdbHit h;
        h.msg = mi; h.bits_off = wb; h.proto = proto; h.nbits = (uint16_t)nbits; h.aux = aux; h.flags = flags; h.rsv = 0;
        A.hits[hb + k] = h;
        for (int i = 0; i < nwords; i++) A.bits[wb + i] = w[i];
    }
    (void)mo;
}

__device__ void emit_one(const HArgs &A, SdbMsgOut &mo, uint32_t mi, uint16_t proto, const uint32_t *w, int nbits,
                         int nwords, uint8_t flags, uint16_t aux)
{
    uint32_t hb = atomicAdd(&A.ctr->hits, 1u);
    mo.hit_off = hb; mo.nhits = 1;
    publish(A, mo, mi, proto, w, nbits, nwords, flags, aux, hb, 0);
}

/* helpers.length_in_range — helpers.py:124-166 */
__device__ __forceinline__ bool in_range(const SdbHexProto &p, int n)
{
    if (!(p.flags & SDB_HF_EXISTS)) return false;                        /* 'protocol does not exists' */
    int mn = (p.flags & SDB_HF_HAS_MIN) ? p.length_min : -1;
    if (mn != -1 && n < mn) return false;
    if ((p.flags & SDB_HF_HAS_MAX) && n > p.length_max) return false;
    return true;
}

/* ---- MC decoders, generic over the bit source (hex nibbles in the batch path, bytes in the unit-op path) ----
 * Reason codes of a rejected frame; the host turns them into the reference's message strings. */
/* (SDB_MCR_* reason codes: include/sdb200.h) */

/* bit string given as one byte per bit (unit-op path: mcBit2*(name, bit_data, protocol_id, mcbitnum)) */
struct ByteBits {
    const uint8_t *b;
    int n;
    __device__ __forceinline__ int bit(int i) const { return b[i] & 1; }
};

/* str.find(pattern of m bits MSB-first, from) */
template <class Bits>
__device__ int find_bits(const Bits &B, uint32_t pat, int m, int from)
{
    if (from < 0) from = 0;
    uint32_t acc = 0, mask = m >= 32 ? 0xffffffffu : ((1u << m) - 1);
    int have = 0;
    for (int i = from; i < B.n; i++) {
        acc = ((acc << 1) | (uint32_t)B.bit(i)) & mask;
        if (++have >= m && acc == pat) return i - m + 1;
    }
    return -1;
}

#define SDB_TFA_MAX 48
struct TfaOut { int start[SDB_TFA_MAX], len[SDB_TFA_MAX], n; };

/*
 * One mcBit2* / mcRaw call (manchester.py:207-795, helpers.py:90-122) on bit string B with the caller's mcbitnum.
 * Returns 1 (R = result bits, or T = list of duplicate parts for TFA), -1 (reason set), or -(100 + SDB_ST_*)
 * when the reference raises.
 */
template <class Bits>
__device__ int mc_decode(const SdbHexProto &p, const Bits &B, int mcbitnum, Res &R, TfaOut &T, int &reason)
{
    const int n = B.n;                                                   /* len(bit_data) */
    const int lmin = (p.flags & SDB_HF_HAS_MIN) ? p.length_min : -1;
    const int lmax = (p.flags & SDB_HF_HAS_MAX) ? p.length_max : 9999;
    reason = SDB_MCR_NONE;
    T.n = 0;
    R.clear();
    switch (p.method) {
    case SDB_M_HIDEKI: case SDB_M_MAVERICK: case SDB_M_OSV1: case SDB_M_OSV2O3: case SDB_M_OSPIR:   /* :418-586 */
        if (mcbitnum < lmin) { reason = SDB_MCR_TOO_SHORT; return -1; }
        if (mcbitnum > lmax) { reason = SDB_MCR_TOO_LONG; return -1; }
        for (int i = 0; i < n; i++) R.push(B.bit(i));
        return 1;
    case SDB_M_MCRAW_MANCHESTER: {                                       /* :588-613 */
        int mx = (p.flags & SDB_HF_HAS_MAX) ? p.length_max : 0;
        if (mcbitnum > mx) { reason = SDB_MCR_TOO_LONG; return -1; }
        for (int i = 0; i < n; i++) R.push(B.bit(i));
        return 1;
    }
    case SDB_M_MCRAW_HELPERS:                                            /* helpers.py:90-122 */
        if (p.flags & SDB_HF_HAS_MAX) {
            if (p.flags & SDB_HF_MAX_IS_STR) return -(100 + SDB_ST_TYPEERROR);   /* int > str */
            if (mcbitnum > p.length_max) { reason = SDB_MCR_TOO_LONG; return -1; }
        }
        for (int i = 0; i < n; i++) R.push(B.bit(i));
        return 1;
    case SDB_M_GROTHE:                                                   /* :721-754 */
        if (mcbitnum != 32) { reason = SDB_MCR_NOT_32; return -1; }
        for (int i = 0; i < n; i++) R.push(B.bit(i));
        return 1;
    case SDB_M_SOMFY: {                                                  /* :756-795 */
        int from = 0, len = n;
        if (mcbitnum == 57) { from = n > 0 ? 1 : 0; len = n > 57 ? 56 : (n > 0 ? n - 1 : 0); }   /* bit_data[1:57] */
        if (len != 56) { reason = SDB_MCR_NOT_56; return -1; }
        for (int i = 0; i < 56; i++) R.push(B.bit(from + i));
        return 1;
    }
    case SDB_M_AS: {                                                     /* :356-416 */
        int sp = n >= 16 ? find_bits(B, 0xC, 4, 16) : -1;
        if (sp >= 0) {
            int ep = find_bits(B, 0xC, 4, sp + 16);
            if (ep < 0) ep = n;
            int ml = ep - sp;
            if (ml < lmin) { reason = SDB_MCR_TOO_SHORT; return -1; }
            if (ml > lmax) { reason = SDB_MCR_TOO_LONG; return -1; }
            for (int i = sp; i < n; i++) R.push(B.bit(i));
        } else {
            if (mcbitnum < lmin) { reason = SDB_MCR_TOO_SHORT; return -1; }
            if (mcbitnum > lmax) { reason = SDB_MCR_TOO_LONG; return -1; }
            for (int i = 0; i < n; i++) R.push(B.bit(i));
        }
        return 1;
    }
    case SDB_M_SAINLOGIC: {                                              /* :302-354 */
        int mx = (p.flags & SDB_HF_HAS_MAX) ? p.length_max : 0;
        if (mcbitnum > mx) { reason = SDB_MCR_TOO_LONG; return -1; }
        int ones = 0, total = n, mb = mcbitnum;
        if (mcbitnum < 128) {
            int start = find_bits(B, 0x14, 6, 0);                        /* '010100' */
            if (start < 0 || start > 10) { reason = SDB_MCR_NO_START; return -1; }
            ones = 10 - start;                                           /* prepend '1' until the sync sits at 10 */
            total = min(128, n + ones);
            mb = total;                                                  /* mcbitnum = len(bit_data) */
        }
        int mn = (p.flags & SDB_HF_HAS_MIN) ? p.length_min : 0;
        if (mb < mn) { reason = SDB_MCR_TOO_SHORT; return -1; }
        for (int i = 0; i < total; i++) R.push(i < ones ? 1 : B.bit(i - ones));
        return 1;
    }
    case SDB_M_FUNKBUS: {                                                /* :207-300 */
        if (mcbitnum < lmin) { reason = SDB_MCR_TOO_SHORT; return -1; }
        if ((p.flags & SDB_HF_HAS_MAX) && mcbitnum > p.length_max) { reason = SDB_MCR_TOO_LONG; return -1; }
        /* differential manchester (helpers.py:6-26): s[k] = (b[k] == b[k+1]), k < n-1 */
        const int ns = n > 0 ? n - 1 : 0;
        int tl, off, plen;     /* t = prefix + s[off:], prefix '001' (119) or '0' */
        if (p.flags & SDB_HF_IS_119) {
            int pos = -1;
            uint32_t acc = 0;
            for (int k = 0; k < ns && k < 9; k++) {                      /* '01100' must start at 0..4 */
                acc = ((acc << 1) | (uint32_t)(B.bit(k) == B.bit(k + 1))) & 0x1F;
                if (k >= 4 && acc == 0x0C) { pos = k - 4; break; }
            }
            if (pos < 0) { reason = SDB_MCR_WRONG_BEGIN; return -1; }
            off = pos; plen = 3; tl = 3 + ns - pos;
            if (tl < 48) { reason = SDB_MCR_WRONG_BEGIN; return -1; }
        } else { off = 0; plen = 1; tl = 1 + ns; }
        int xorv = 0, chk = 0, parity = 0;
        for (int i = 0; i < 6; i++) {
            int from = i * 8, to = min(from + 8, tl);
            if (from >= to) return -(100 + SDB_ST_VALUEERROR);           /* int('', 2) */
            int data = 0;
            for (int k = from; k < to; k++) {
                int b;
                if (k < plen) b = (plen == 3 && k == 2) ? 1 : 0;
                else { int si = off + (k - plen); b = (B.bit(si) == B.bit(si + 1)); }
                data = (data << 1) | b;
            }
            for (int k = 7; k >= 0; k--) R.push((data >> k) & 1);        /* f"{data:02X}" */
            if (i < 5) xorv ^= data;
            else { chk = data & 0x0F; xorv ^= data & 0xE0; data &= 0xF0; }
            parity ^= __popc(data) & 1;
        }
        if (parity) { reason = SDB_MCR_PARITY; return -1; }
        int xn = ((xorv & 0xF0) >> 4) ^ (xorv & 0x0F), r = 0;
        if (xn & 8) r ^= 0xC;
        if (xn & 4) r ^= 0x2;
        if (xn & 2) r ^= 0x8;
        if (xn & 1) r ^= 0x3;
        if (r != chk) { reason = SDB_MCR_CHECKSUM; return -1; }
        return 1;
    }
    case SDB_M_TFA: {                                                    /* :615-719 */
        int f = find_bits(B, 0xFFD, 12, 0);                              /* '111111111101' */
        if (f < 0) { reason = SDB_MCR_NO_SYNC; return -1; }
        int pre = f + 12, mend = -1, it = 1, nm = 0, last_fail = 0;
        int ps[SDB_TFA_MAX], pl[SDB_TFA_MAX];
        while (mend < mcbitnum) {
            mend = (pre >= 0 && pre <= n) ? find_bits(B, 0x1FFD, 13, pre) : -1;   /* '1111111111101' */
            if (mend < pre) mend = mcbitnum;
            int ml = mend - pre;
            if (in_range(p, ml)) {
                /* bit_data[pre:mend] is clipped to the data that exists */
                int e = mend < n ? mend : n, s = pre < n ? pre : n;
                if (nm < SDB_TFA_MAX) { ps[nm] = s; pl[nm] = e > s ? e - s : 0; nm++; }
            } else {
                int mn = (p.flags & SDB_HF_HAS_MIN) ? p.length_min : -1;
                last_fail = !(p.flags & SDB_HF_EXISTS) ? SDB_MCR_DUP_NOPROTO
                            : (mn != -1 && ml < mn) ? SDB_MCR_DUP_SHORT : SDB_MCR_DUP_LONG;
            }
            int q = mend <= n ? find_bits(B, 0xD, 4, mend) : -1;         /* '1101' */
            if (q >= 0) pre = q + 4; else { pre = -1; mend = mcbitnum; }
            it++;
        }
        if (it == 10) { reason = SDB_MCR_LOOP; return -1; }
        /* :706-711: every element whose hex string was seen exactly once before */
        int seen[SDB_TFA_MAX];
        for (int a = 0; a < nm; a++) {
            int first = a;
            for (int b = 0; b < a; b++) {
                /* hex strings equal <=> same digit count and same right-aligned value */
                if (((pl[a] + 3) >> 2) != ((pl[b] + 3) >> 2)) continue;
                int la = pl[a], lb = pl[b], L = max(la, lb);
                bool eq = true;
                for (int k = 0; k < L && eq; k++) {
                    int ia = la - 1 - k, ib = lb - 1 - k;
                    int ba = ia >= 0 ? B.bit(ps[a] + ia) : 0, bb = ib >= 0 ? B.bit(ps[b] + ib) : 0;
                    eq = ba == bb;
                }
                if (eq) { first = b; break; }
            }
            seen[a] = 0;
            if (seen[first] == 1) { T.start[T.n] = ps[a]; T.len[T.n] = pl[a]; T.n++; }
            seen[first]++;
        }
        if (T.n == 0) { reason = SDB_MCR_NO_DUP | last_fail; return -1; }
        return 1;
    }
    default:
        return -1;
    }
}

/* ---- MC batch path: returns SDB_ST_* ------------------------------------------------------- */
__device__ int mc_one(const HArgs &A, SdbMsgOut &mo, uint32_t mi, const SdbHexMsg &m)
{
    /* mo.reason = why the message was rejected (SDB_MCR_*; _demodulate_mc_data returns that text, manchester.py:70-128) */
    const SdbHexProto p = A.tab.hex[m.proto];
    const int mcbitnum = m.bitlen;
    const int lmin = (p.flags & SDB_HF_HAS_MIN) ? p.length_min : -1;     /* manchester.py:70-79 */
    if (mcbitnum < lmin) { mo.reason = SDB_MCR_TOO_SHORT; return SDB_ST_OK; }
    const int lmax = (p.flags & SDB_HF_HAS_MAX) ? p.length_max : 9999;
    if (mcbitnum > lmax) { mo.reason = SDB_MCR_TOO_LONG; return SDB_ST_OK; }
    if (p.flags & SDB_HF_CLOCKRANGE) {                                   /* :81-86 */
        if (!A.repaired) return SDB_ST_TYPEERROR;                        /* int > list */
        if (!(m.clock > p.clock_min && m.clock < p.clock_max)) { mo.reason = SDB_MCR_CLOCK; return SDB_ST_OK; }
    }
    bool inv = (p.flags & SDB_HF_INVERT) != 0;                           /* :91-96 */
    if (m.flags & SDB_HEX_TOGGLE_POLARITY) inv = !inv;
    if (p.method == SDB_M_NONE) { mo.reason = SDB_MCR_NO_METHOD; return SDB_ST_VALUEERROR; }   /* :109 1-list cannot unpack into 3 */
    if (p.method == SDB_M_UNKNOWN) { mo.reason = SDB_MCR_UNKNOWN_METHOD; return SDB_ST_OK; }  /* :121-123 */
    if (m.hlen == 0) return SDB_ST_TYPEERROR;                            /* hex_to_bin_str('') is None -> len(None) */
    if (!A.repaired && p.method == SDB_M_MCRAW_MANCHESTER) {
        /* :120 passes self twice; mcRaw has a spare parameter, so the call works with shifted arguments (:588-613):
         * bit_data = the name, protocol_id = the bit string (no such id: length_max 0), mcbitnum = int(protocol id) */
        if (p.pid_int == 0) return SDB_ST_VALUEERROR;                    /* int('13.2') */
        if (p.pid_int == 1) { mo.reason = SDB_MCR_TOO_LONG; return SDB_ST_OK; }
        Res R0;
        R0.clear();
        emit_one(A, mo, mi, m.proto, R0.w, 0, 0, SDB_HIT_HAS_F, 0);     /* bin_str_2_hex_str(name) is None: payload = preamble + "None" */
        return SDB_ST_OK;
    }
    if (!A.repaired) return SDB_ST_TYPEERROR;                            /* :120 self passed twice */
    if (p.method >= SDB_M_BRESSER_LIGHTNING) return SDB_ST_TYPEERROR;    /* Conv*(msg_data, msg_type) called with 4 args */

    McBits B;
    B.h.d = A.digits + (size_t)m.doff * 16; B.h.n = m.hlen; B.inv = inv; B.lead = 0;
    while (B.lead < m.hlen - 1 && B.nib(0) == 0) B.lead++;               /* helpers.py:183-186; nib(0) is relative to lead */
    B.n = 4 * (m.hlen - B.lead);

    Res R;
    TfaOut T;
    int reason;
    const int rc = mc_decode(p, B, B.n, R, T, reason);                   /* :120 mcbitnum = len(bit_data) */
    if (rc <= -100) return -rc - 100;
    if (rc != 1) {
        int r8 = reason & 0xFF;                                          /* the TFA suffix flags fold into codes 12..14 */
        if (reason & SDB_MCR_DUP_SHORT) r8 = 12; else if (reason & SDB_MCR_DUP_LONG) r8 = 13; else if (reason & SDB_MCR_DUP_NOPROTO) r8 = 14;
        mo.reason = (uint8_t)r8;
        return SDB_ST_OK;
    }
    if (p.method == SDB_M_TFA) {
        uint32_t hb = atomicAdd(&A.ctr->hits, (uint32_t)T.n);
        mo.hit_off = hb; mo.nhits = (uint16_t)T.n;
        for (int e = 0; e < T.n; e++) {
            R.clear();
            for (int k = 0; k < T.len[e]; k++) R.push(B.bit(T.start[e] + k));
            publish(A, mo, mi, m.proto, R.w, R.n, (R.n + 31) >> 5, SDB_HIT_LIST, (uint16_t)e, hb, (uint32_t)e);
        }
        return SDB_ST_OK;
    }
    emit_one(A, mo, mi, m.proto, R.w, R.n, (R.n + 31) >> 5, 0, 0);       /* sd_protocols.py:102-109 */
    return SDB_ST_OK;
}

/* unit op: one mcBit2* call on a byte-per-bit string.  out = result bits (bytes); for TFA the parts are
 * concatenated and seg[] holds their lengths.  res = {rc, reason, n_out, nseg}. */
__global__ void unit_mc_kernel(SdbDevTable tab, uint32_t proto, int method_override, const uint8_t *bits, int n, int mcbitnum,
                               uint8_t *out, int out_cap, int32_t *seg, int32_t *res)
{
    if (threadIdx.x != 0) return;
    SdbHexProto p;
    if (proto < tab.nproto) p = tab.hex[proto];
    else { p = SdbHexProto{}; if (method_override & 0x100) p.flags = SDB_HF_IS_119; }   /* id not in the table: defaults only */
    if (method_override & 0xFF) p.method = (uint8_t)(method_override & 0xFF);
    ByteBits B;
    B.b = bits; B.n = n;
    Res R;
    TfaOut T;
    int reason = 0, nout = 0, nseg = 0;
    int rc = mc_decode(p, B, mcbitnum, R, T, reason);
    if (rc == 1) {
        if (p.method == SDB_M_TFA) {
            for (int e = 0; e < T.n; e++) {
                seg[nseg++] = T.len[e];
                for (int k = 0; k < T.len[e] && nout < out_cap; k++) out[nout++] = (uint8_t)B.bit(T.start[e] + k);
            }
        } else {
            for (int i = 0; i < R.n && nout < out_cap; i++) out[nout++] = (uint8_t)((R.w[i >> 5] >> (i & 31)) & 1);
        }
    }
    res[0] = rc; res[1] = reason; res[2] = nout; res[3] = nseg;
}

int launch_unit_mc(const SdbDevTable &tab, uint32_t proto, int method_override, const uint8_t *d_bits, int n, int mcbitnum,
                   uint8_t *d_out, int out_cap, int32_t *d_seg, int32_t *d_res, cudaStream_t stream)
{
    if (n > SDB_MAX_HEX * 4) return -1;
    unit_mc_kernel<<<1, 32, 0, stream>>>(tab, proto, method_override, d_bits, n, mcbitnum, d_out, out_cap, d_seg, d_res);
    return (int)cudaGetLastError();
}

/* ---- MN ------------------------------------------------------------------------------------ */
/* helpers.lfsr_digest16 (helpers.py:190-221) over nibbles already XORed with 0xA */
__device__ int lfsr16(const Nibs &h, int first_nib, int bytes, int gen, int key)
{
    int lfsr = 0;
    for (int k = 0; k < bytes; k++) {
        int data = ((h.at(first_nib + 2 * k) ^ 0xA) << 4) | (h.at(first_nib + 2 * k + 1) ^ 0xA);
        for (int i = 7; i >= 0; i--) {
            if ((data >> i) & 1) lfsr ^= key;
            key = (key & 1) ? ((key >> 1) ^ gen) : (key >> 1);
        }
    }
    return lfsr;
}
/* helpers._calc_crc16, refin = refout = False, init 0, xorout 0 (helpers.py:281-309) */
__device__ int crc16(const Nibs &h, int first_byte, int nbytes, int poly)
{
    int crc = 0;
    for (int k = 0; k < nbytes; k++) {
        crc ^= h.byte(first_byte + k) << 8;
        for (int i = 0; i < 8; i++) crc = (crc & 0x8000) ? (((crc << 1) ^ poly) & 0xFFFF) : ((crc << 1) & 0xFFFF);
    }
    return crc;
}

__device__ int mn_one(const HArgs &A, SdbMsgOut &mo, uint32_t mi, const SdbHexMsg &m)
{
    SdbHexProto p = A.tab.hex[m.proto];
    if (m.rsv) p.method = m.rsv;                                         /* a direct Conv*(msg_data) call names the converter */
    if (p.method < SDB_M_BRESSER_LIGHTNING || p.method == SDB_M_UNKNOWN) return SDB_ST_OK;   /* sd_protocols.py:125-149 */
    Nibs h;
    h.d = A.digits + (size_t)m.doff * 16; h.n = m.hlen;
    const int n = m.hlen;
    if (n == 0) return SDB_ST_OK;                                        /* `if not hex_data` */
    Res R;
    R.clear();
    uint8_t flags = 0;
    int nwords = 0;
    switch (p.method) {
    case SDB_M_BRESSER_LIGHTNING: case SDB_M_BRESSER_7IN1: {             /* helpers.py:223-280, :473-523 */
        const bool seven = p.method == SDB_M_BRESSER_7IN1;
        if (n < (seven ? 46 : 20)) return SDB_ST_OK;
        if (seven && h.at(42) == 0 && h.at(43) == 0) return SDB_ST_OK;
        int cs = seven ? lfsr16(h, 4, 21, 0x8810, 0xBA95) : lfsr16(h, 4, 8, 0x8810, 0xABF9);
        int first = ((h.at(0) ^ 0xA) << 12) | ((h.at(1) ^ 0xA) << 8) | ((h.at(2) ^ 0xA) << 4) | (h.at(3) ^ 0xA);
        if ((cs ^ first) != (seven ? 0x6DF1 : 0x899E)) return SDB_ST_OK;
        int outn = seven ? n : 20;
        for (int i = 0; i < outn; i++) { int x = h.at(i) ^ 0xA; for (int k = 3; k >= 0; k--) R.push((x >> k) & 1); }
        break;
    }
    case SDB_M_BRESSER_5IN1: {                                           /* :382-425 */
        if (n < 52) return SDB_ST_OK;
        int bit_add = 0, ref = 0;
        for (int i = 0; i < 13; i++) {
            int b = h.byte(i), iv = h.byte(i + 13);
            if ((b ^ iv) != 0xFF) return SDB_ST_OK;
            if (i == 0) ref = iv; else bit_add += __popc(iv);
        }
        if (bit_add != ref) return SDB_ST_OK;
        for (int i = 28; i < 52; i++) { int x = h.at(i); for (int k = 3; k >= 0; k--) R.push((x >> k) & 1); }
        break;
    }
    case SDB_M_BRESSER_6IN1: {                                           /* :427-471 */
        if (n < 36) return SDB_ST_OK;
        int want = (h.byte(0) << 8) | h.byte(1);
        if (crc16(h, 2, 15, 0x1021) != want) return SDB_ST_OK;
        int sum = 0;
        for (int i = 2; i < 18; i++) sum += h.byte(i);
        if ((sum & 0xFF) != 0xFF) return SDB_ST_OK;
        for (int i = 0; i < n; i++) { int x = h.at(i); for (int k = 3; k >= 0; k--) R.push((x >> k) & 1); }
        break;
    }
    case SDB_M_PCA301: {                                                 /* :525-579 */
        if (n < 24) return SDB_ST_OK;
        int want = (h.byte(10) << 8) | h.byte(11);
        if (crc16(h, 0, 10, 0x8005) != want) return SDB_ST_OK;
        for (int i = 0; i < 10; i++) R.w[i] = (uint32_t)(i == 5 ? (h.byte(5) & 0x0F) : h.byte(i));
        R.w[10] = (uint32_t)want;
        R.n = 11; nwords = 11; flags = SDB_HIT_FIELDS;
        break;
    }
    case SDB_M_KOPP: {                                                   /* :581-628 */
        if (n < 4) return SDB_ST_OK;
        int anz = h.byte(0) + 1;
        if (n < anz * 2 + 2) return SDB_ST_OK;
        int blk = 0xAA;
        for (int i = 0; i < anz; i++) blk ^= h.byte(i);
        if (blk != h.byte(anz)) return SDB_ST_OK;
        for (int i = 0; i < anz * 2; i++) { int x = h.at(i); for (int k = 3; k >= 0; k--) R.push((x >> k) & 1); }
        break;
    }
    case SDB_M_LACROSSE: {                                               /* :630-716 */
        if (n < 10) return SDB_ST_OK;
        int crc = 0;
        for (int k = 0; k < 4; k++) {
            crc ^= h.byte(k);
            for (int i = 0; i < 8; i++) crc = (crc & 0x80) ? (((crc << 1) ^ 0x31) & 0xFF) : ((crc << 1) & 0xFF);
        }
        if (crc != h.byte(4)) return SDB_ST_OK;
        int b0 = h.byte(0), b1 = h.byte(1), b2 = h.byte(2), b3 = h.byte(3);
        int traw = (b1 & 0x0F) * 100 + ((b2 & 0xF0) >> 4) * 10 + (b2 & 0x0F);
        /* float64 exactly as Python: (raw / 10) - 40, then int(t * 10 + 1000) */
        double temperature = __dsub_rn(__ddiv_rn((double)traw, 10.0), 40.0);
        if (temperature >= 60.0 || temperature <= -40.0) return SDB_ST_OK;
        int ts = ((int)__dadd_rn(__dmul_rn(temperature, 10.0), 1000.0)) & 0xFFFF;
        int typ = ((b3 & 0x7F) == 125) ? 2 : 1;
        R.w[0] = (uint32_t)(((b0 & 0x0F) << 2) | ((b1 & 0xC0) >> 6));
        R.w[1] = (uint32_t)(typ | ((b1 & 0x20) << 2));
        R.w[2] = (uint32_t)((ts >> 8) & 0xFF);
        R.w[3] = (uint32_t)(ts & 0xFF);
        R.w[4] = (uint32_t)b3;
        R.n = 5; nwords = 5; flags = SDB_HIT_FIELDS;
        break;
    }
    default:
        return SDB_ST_OK;
    }
    if (!(flags & SDB_HIT_FIELDS)) nwords = (R.n + 31) >> 5;
    emit_one(A, mo, mi, m.proto, R.w, R.n, nwords, flags, p.method);
    return SDB_ST_OK;
}

__global__ void __launch_bounds__(SDB_HEX_THREADS) hex_kernel(HArgs A)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t mi = blockIdx.x * blockDim.x + threadIdx.x; mi < A.n; mi += stride) {
        const SdbHexMsg m = A.msgs[mi];
        SdbMsgOut mo;
        mo.hit_off = 0; mo.nhits = 0; mo.status = SDB_ST_OK; mo.reason = 0;
        if ((m.flags & SDB_MSG_DOMAIN) || ((m.flags & SDB_MSG_VALID) && m.hlen > SDB_MAX_HEX)) {
            mo.status = SDB_ST_DOMAIN;                      /* not representable: reported, never decoded differently */
            atomicAdd(&A.ctr->domain, 1u);
        } else if ((m.flags & SDB_MSG_VALID) && m.proto < A.tab.nproto) {
            int st = A.kind == SDB_KIND_MC ? mc_one(A, mo, mi, m) : mn_one(A, mo, mi, m);
            if (st != SDB_ST_OK) {
                mo.status = (uint8_t)st; mo.nhits = 0;        /* (mo.reason keeps SDB_MCR_NO_METHOD: the direct call returns a list there) */
                atomicAdd(&A.ctr->raised, 1u);
            }
        }
        A.out[mi] = mo;
    }
}

int launch_hex(int kind, int mc_repaired, const SdbDevTable &tab, const SdbHexMsg *d_msgs, const uint8_t *d_digits,
               uint32_t n, SdbMsgOut *d_out, SdbHit *d_hits, uint32_t hits_cap, uint32_t *d_bits, uint32_t bits_cap,
               SdbCounters *d_ctr, int grid, cudaStream_t stream)
{
    HArgs A;
    A.tab = tab; A.msgs = d_msgs; A.digits = d_digits; A.n = n; A.kind = kind; A.repaired = mc_repaired;
    A.out = d_out; A.hits = d_hits; A.hits_cap = hits_cap; A.bits = d_bits; A.bits_cap = bits_cap; A.ctr = d_ctr;
    if (n == 0) return 0;
    hex_kernel<<<grid, SDB_HEX_THREADS, 0, stream>>>(A);
    return (int)cudaGetLastError();
}

}  // namespace sdb
