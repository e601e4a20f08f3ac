/*
 * sdb_lines.cu — firmware text lines -> packed MS / MU records, on the device.
 *
 * Replaces, for a whole batch of payload lines of ONE message type,
 *   MSParser._parse_to_dict / MUParser._parse_to_dict     signalduino/parser/ms.py:71-84, mu.py:84-95
 *   the MU validity regex                                  signalduino/parser/mu.py:48-52
 *   "D" present -> msg_data["data"] = msg_data["D"]        ms.py:41-46, mu.py:57-61
 *   the input gates of demodulate_ms / demodulate_mu       sd_protocols/message_synced.py:21-66,
 *                                                          sd_protocols/message_unsynced.py:22-35
 * i.e. everything between extract_payload() (base.py:174-193, host) and the demodulation kernels.
 *
 * One THREAD per line: the work is a byte-serial field scan (100-900 bytes per line, ~10 instructions per
 * byte) and 32 lines per warp keep the issue slots busy; the per-line cost is ~1 % of the MU demodulation
 * of the same line, so there is nothing to gain from a warp-cooperative scan.  Each line's digit stream goes to
 * the digit pool at unit (line_off >> 5) + i (16-byte units): D is shorter than the line, so the regions never
 * overlap and no prefix sum is needed.
 *
 * Exactness: the device accepts the canonical grammar the firmware emits (keys D, CP, SP, R, P<d> with
 * values -?[0-9]{1,10}); every line outside it (non-ASCII bytes, duplicate or multi-digit pattern ids,
 * values float() parses differently, D longer than 1024 digits) is flagged SDB_LINE_HOSTPATH and goes
 * through the host packer (pack.py), never decoded differently.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sdb200.h"
#include "sdb_pulse.h"

namespace sdb {

struct LArgs {
    const uint8_t *text;
    const uint32_t *line_off, *line_len;
    uint32_t n;
    SdbPulseMsg *msgs;
    uint8_t *pool;            /* digit pool, ((text_len >> 5) + n + 2) * 16 bytes */
    SdbLineInfo *info;
};

struct Span { int a, b; };                     /* [a, b) inside the line; a < 0 = key absent */

__device__ __forceinline__ bool is_dig(uint8_t c) { return c >= '0' && c <= '9'; }
__device__ __forceinline__ bool all_digits(const uint8_t *s, Span v)
{
    if (v.a < 0 || v.b <= v.a) return false;   /* "".isdigit() is False */
    for (int i = v.a; i < v.b; i++) if (!is_dig(s[i])) return false;
    return true;
}
/* -?[0-9]{1,10} inside int32 (without INT_MIN): the only value syntax the device packs itself */
__device__ __forceinline__ bool canon_int(const uint8_t *s, int a, int b, int32_t &out)
{
    bool neg = false;
    if (a < b && s[a] == '-') { neg = true; a++; }
    if (b - a < 1 || b - a > 10) return false;
    int64_t v = 0;
    for (int i = a; i < b; i++) { if (!is_dig(s[i])) return false; v = v * 10 + (s[i] - '0'); }
    if (v > 2147483647LL) return false;
    out = (int32_t)(neg ? -v : v);
    return true;
}

struct Pats {
    int32_t val[SDB_MAX_SLOTS];
    uint32_t ids;
    int n;
    __device__ __forceinline__ int find(int id) const
    {
        for (int s = 0; s < n; s++) if ((int)((ids >> (4 * s)) & 0xF) == id) return s;
        return -1;
    }
};

/* nibble-pack D (digits only by construction) into the pool region of this line, 0xF padding to the 16-byte unit */
__device__ __forceinline__ void pack_digits(const uint8_t *s, Span d, uint8_t *pool, uint32_t unit)
{
    const int dlen = d.b - d.a;
    uint32_t *w = reinterpret_cast<uint32_t *>(pool + (size_t)unit * 16);
    const int nwords = ((dlen + 31) >> 5) * 4;
    for (int k = 0; k < nwords; k++) {
        uint32_t x = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int i = 8 * k + j;
            const uint32_t nib = i < dlen ? (is_dig(s[d.a + i]) ? (uint32_t)(s[d.a + i] - '0') : 0xEu) : 0xFu;
            x |= nib << (4 * j);
        }
        w[k] = x;
    }
}

template <bool MU>
__device__ void tokenize_one(const uint8_t *s, int len, uint32_t unit, uint8_t *pool, SdbPulseMsg &rec, SdbLineInfo &info)
{
    info.status = SDB_LINE_INVALID; info.has_r = 0; info.r_len = 0; info.r_off = 0; info.clock = 0;
    for (int i = 0; i < SDB_MAX_SLOTS; i++) rec.pat[i] = 0;
    rec.doff = unit; rec.dlen = 0; rec.npat = 0; rec.cp = 0xFF; rec.pat_ids = 0; rec.flags = 0;
    rec.rsv[0] = rec.rsv[1] = rec.rsv[2] = 0;
    if (len > 8191) { info.status = SDB_LINE_HOSTPATH; return; }
    for (int i = 0; i < len; i++)
        if (s[i] >= 0x80) { info.status = SDB_LINE_HOSTPATH; return; }      /* str.isdigit() / \d know more digits than ASCII */

    Span D = {-1, -1}, CP = {-1, -1}, SP = {-1, -1}, R = {-1, -1};
    Pats P;
    P.ids = 0; P.n = 0;
    bool host = false;

    if (MU) {
        /* ^(?=.*D=\d+)(?:MU;(?:P[0-7]=-?[0-9]{1,5};){2,8}((?:D=\d{2,};)|(?:CP=\d;)|(?:R=\d+;)|(?:O;)|(?:e;)|(?:p;)|(?:w=\d;))*)$
         * (mu.py:48): no alternative of the second group matches a pattern field and vice versa, so the split is unique */
        if (len < 3 || s[0] != 'M' || s[1] != 'U' || s[2] != ';') return;
        int p = 3, npf = 0;
        bool tail = false;
        while (p < len) {
            int q = p;
            while (q < len && s[q] != ';') q++;
            if (q >= len) return;                                           /* every field ends with ';' */
            const int fl = q - p;
            bool is_pat = false;
            if (fl >= 4 && s[p] == 'P' && s[p + 1] >= '0' && s[p + 1] <= '7' && s[p + 2] == '=') {
                int a = p + 3;
                if (s[a] == '-') a++;
                const int nd = q - a;
                is_pat = nd >= 1 && nd <= 5;
                for (int i = a; i < q && is_pat; i++) is_pat = is_dig(s[i]);
            }
            if (is_pat) {
                if (tail) return;
                if (++npf > 8) return;
                int32_t v = 0;
                canon_int(s, p + 3, q, v);
                const int id = s[p + 1] - '0';
                const int slot = P.find(id);                                /* same dict key: the value is overwritten in place */
                if (slot >= 0) P.val[slot] = v;
                else { P.val[P.n] = v; P.ids |= (uint32_t)id << (4 * P.n); P.n++; }
            } else {
                if (!tail) { if (npf < 2) return; tail = true; }
                bool ok = false;
                if (fl >= 4 && s[p] == 'D' && s[p + 1] == '=') {
                    ok = true;
                    for (int i = p + 2; i < q && ok; i++) ok = is_dig(s[i]);
                    if (ok) { D.a = p + 2; D.b = q; }
                } else if (fl == 4 && s[p] == 'C' && s[p + 1] == 'P' && s[p + 2] == '=' && is_dig(s[p + 3])) ok = true;
                else if (fl >= 3 && s[p] == 'R' && s[p + 1] == '=') {
                    ok = true;
                    for (int i = p + 2; i < q && ok; i++) ok = is_dig(s[i]);
                    if (ok) { R.a = p + 2; R.b = q; }
                } else if (fl == 1 && (s[p] == 'O' || s[p] == 'e' || s[p] == 'p')) ok = true;
                else if (fl == 3 && s[p] == 'w' && s[p + 1] == '=' && is_dig(s[p + 2])) ok = true;
                if (!ok) return;
            }
            p = q + 1;
        }
        if (npf < 2 || D.a < 0) return;                                     /* {2,8} and the D=\d+ look-ahead */
    } else {
        /* _parse_to_dict (ms.py:71-84): split on ';', key = text before the first '=', later keys overwrite */
        int p = 0;
        while (p < len) {
            int q = p, eq = -1;
            while (q < len && s[q] != ';') { if (s[q] == '=' && eq < 0) eq = q; q++; }
            if (q > p) {
                const int ke = eq >= 0 ? eq : q;                            /* key = [p, ke) */
                const Span v = {eq >= 0 ? eq + 1 : q, q};
                const int kl = ke - p;
                if (kl == 1 && s[p] == 'D') D = v;
                else if (kl == 2 && s[p] == 'C' && s[p + 1] == 'P') CP = v;
                else if (kl == 2 && s[p] == 'S' && s[p + 1] == 'P') SP = v;
                else if (kl == 1 && s[p] == 'R') R = v;
                else if (kl >= 2 && s[p] == 'P') {
                    bool kd = true;
                    int id = 0;
                    for (int i = p + 1; i < ke; i++) { kd = kd && is_dig(s[i]); id = id * 10 + (s[i] - '0'); if (id > 1000) id = 1000; }
                    if (kd) {                                                /* a pattern key (message_synced.py:50-57) */
                        int32_t val = 0;
                        if (id > 9 || !canon_int(s, v.a, v.b, val) || P.find(id) >= 0 || P.n >= SDB_MAX_SLOTS) host = true;
                        else { P.val[P.n] = val; P.ids |= (uint32_t)id << (4 * P.n); P.n++; }
                    }
                }
            }
            p = q + 1;
        }
        if (D.a < 0) return;                                                /* ms.py:41-43: no D, nothing to demodulate */
        /* message_synced.py:21-47 */
        if (!all_digits(s, D) || !all_digits(s, CP) || !all_digits(s, SP)) return;
        if (R.a >= 0 && !all_digits(s, R)) return;
    }
    if (host || D.b - D.a > SDB_MAX_DIGITS) { info.status = SDB_LINE_HOSTPATH; return; }

    for (int i = 0; i < P.n; i++) rec.pat[i] = P.val[i];
    rec.npat = (uint8_t)P.n; rec.pat_ids = P.ids;
    rec.dlen = (uint16_t)(D.b - D.a);
    rec.flags = SDB_MSG_VALID;
    if (!MU) {
        int cpv = 0;                                                         /* str(int(CP)) (message_synced.py:33,59) */
        for (int i = CP.a; i < CP.b; i++) { cpv = cpv * 10 + (s[i] - '0'); if (cpv > 1000) cpv = 1000; }
        const int slot = cpv <= 9 ? P.find(cpv) : -1;
        if (slot >= 0) { rec.cp = (uint8_t)slot; info.clock = P.val[slot] < 0 ? -P.val[slot] : P.val[slot]; }
    }
    if (R.a >= 0) { info.r_off = (uint32_t)R.a; info.r_len = (uint16_t)(R.b - R.a); info.has_r = 1; }
    pack_digits(s, D, pool, unit);
    info.status = SDB_LINE_OK;
}

template <bool MU>
__global__ void __launch_bounds__(128) tokenize_kernel(LArgs A)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < A.n; i += gridDim.x * blockDim.x) {
        const uint32_t off = A.line_off[i], len = A.line_len[i];
        SdbPulseMsg rec;
        SdbLineInfo info;
        tokenize_one<MU>(A.text + off, (int)len, (off >> 5) + i, A.pool, rec, info);
        A.msgs[i] = rec;
        A.info[i] = info;
    }
}

size_t lines_pool_bytes(size_t text_len, uint32_t n) { return ((text_len >> 5) + (size_t)n + 4) * 16; }

int launch_tokenize(int kind, const uint8_t *d_text, const uint32_t *d_off, const uint32_t *d_len, uint32_t n,
                    SdbPulseMsg *d_msgs, uint8_t *d_pool, SdbLineInfo *d_info, int sm_count, cudaStream_t stream)
{
    if (n == 0) return 0;
    LArgs A;
    A.text = d_text; A.line_off = d_off; A.line_len = d_len; A.n = n; A.msgs = d_msgs; A.pool = d_pool; A.info = d_info;
    uint32_t need = (n + 127) / 128;
    uint32_t grid = (uint32_t)sm_count * 8;
    if (need < grid) grid = need;
    if (kind == SDB_KIND_MU) tokenize_kernel<true><<<grid, 128, 0, stream>>>(A);
    else tokenize_kernel<false><<<grid, 128, 0, stream>>>(A);
    return (int)cudaGetLastError();
}

}  // namespace sdb
