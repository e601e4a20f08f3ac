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
 * One WARP per line (a first version with one thread per line diverged on every byte loop and cost 6-11 k warp
 * instructions per line — as much as demodulating it):
 *   1. the line is staged in shared memory with coalesced 16-byte loads;
 *   2. 32 bytes per step: __ballot_sync finds the ';' separators (and any non-ASCII byte), the lanes holding one
 *      record the field boundaries;
 *   3. one LANE per field classifies and parses its (short) field: key, '=' position, canonical integer value;
 *   4. ballots / shuffles apply the dict semantics across fields: last D / CP / SP / R wins, pattern slots in order
 *      of first appearance, the MU regex as conditions on the field-class masks;
 *   5. the warp checks and nibble-packs the D digits, 8 per lane, straight into the digit pool (coalesced), at unit
 *      (line_off >> 5) + i (16-byte units): D is shorter than the line, so regions never overlap and no prefix
 *      sum is needed.
 *
 * Exactness: the device accepts the canonical grammar the firmware emits (keys D, CP, SP, R, P<d> with
 * values -?[0-9]{1,10}); every line outside it (non-ASCII bytes, duplicate or multi-digit pattern ids in MS,
 * values float() parses differently, D longer than SDB_MAX_DIGITS digits, more than 32 fields, longer than LN_LONG_MAX
 * bytes) is flagged SDB_LINE_HOSTPATH and goes through the host packer (pack.py), never decoded differently.
 * Lines longer than LN_MAX bytes (D of more than ~1000 digits) are listed by the first pass and tokenized by a second
 * instantiation of the same kernel with a 4.6 KB staging buffer per warp (2 warps per CTA; exits at once when the list is empty).
 */
#include <cuda_runtime.h>
#include <cstdlib>
#include <stdint.h>

#include "../../include/sdb200.h"
#include "sdb_pulse.h"

namespace sdb {

#define FULL 0xffffffffu
#define LN_MAX 1280                              /* staged bytes per line in the first pass */
#define LN_WARPS 8
#define LN_LONG_MAX (SDB_MAX_DIGITS + 512)       /* second pass (listed long lines); longer lines take the host path */
#define LN_LONG_WARPS 2

struct LArgs {
    const uint8_t *text;
    const uint32_t *line_off, *line_len;
    uint32_t n;
    uint32_t base;            /* index of line 0 of this launch in the caller's batch (digit-pool unit = (off >> 5) + base + i) */
    SdbPulseMsg *msgs;
    uint8_t *pool;            /* digit pool, ((text_len >> 5) + n + 4) * 16 bytes */
    SdbLineInfo *info;
    uint32_t *long_list;      /* lines of this launch with LN_MAX < len <= LN_LONG_MAX (first pass -> second pass) */
    uint32_t *long_cnt;       /* [0] list length, [1] the second pass's work counter */
};

template <int LNMAX>
struct __align__(16) LineSm {
    uint8_t  buf[LNMAX + 48];                    /* the line, at the same 16-byte phase as in global memory */
    uint16_t fend[33];                           /* position of the ';' that ends field k */
    uint32_t rec[12];                            /* SdbPulseMsg being assembled */
};

/* field classes (step 3) */
enum { F_EMPTY = 0, F_IGNORE, F_PAT, F_D, F_CP, F_SP, F_R, F_OTHER_OK, F_HOST, F_BAD };

__device__ __forceinline__ bool is_dig(uint8_t c) { return c >= '0' && c <= '9'; }

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
__device__ __forceinline__ bool short_digits(const uint8_t *s, int a, int b)      /* str.isdigit() on a short value */
{
    if (b <= a) return false;
    for (int i = a; i < b; i++) if (!is_dig(s[i])) return false;
    return true;
}

/* lane: classify field [a, b).  MS follows _parse_to_dict (key = text before the first '='), MU the alternatives of the
 * validity regex.  va / vb = value span, id / val = pattern id and value. */
template <bool MU>
__device__ __forceinline__ int classify(const uint8_t *s, int a, int b, int &va, int &vb, int &id, int32_t &val)
{
    const int fl = b - a;
    va = vb = b; id = 0; val = 0;
    if (fl == 0) return MU ? F_BAD : F_EMPTY;
    const uint8_t c0 = s[a];
    if (MU) {
        if (fl >= 4 && c0 == 'P' && s[a + 1] >= '0' && s[a + 1] <= '7' && s[a + 2] == '=') {
            int p = a + 3;
            if (s[p] == '-') p++;
            const int nd = b - p;
            if (nd >= 1 && nd <= 5 && short_digits(s, p, b)) { canon_int(s, a + 3, b, val); id = s[a + 1] - '0'; return F_PAT; }
            return F_BAD;
        }
        if (fl >= 4 && c0 == 'D' && s[a + 1] == '=') { va = a + 2; vb = b; return F_D; }     /* \d{2,}: checked by the warp */
        if (fl == 4 && c0 == 'C' && s[a + 1] == 'P' && s[a + 2] == '=' && is_dig(s[a + 3])) return F_OTHER_OK;
        if (fl >= 3 && c0 == 'R' && s[a + 1] == '=') { va = a + 2; vb = b; return short_digits(s, va, vb) ? F_R : F_BAD; }
        if (fl == 1 && (c0 == 'O' || c0 == 'e' || c0 == 'p')) return F_OTHER_OK;
        if (fl == 3 && c0 == 'w' && s[a + 1] == '=' && is_dig(s[a + 2])) return F_OTHER_OK;
        return F_BAD;
    }
    if (c0 == 'D' && (fl == 1 || s[a + 1] == '=')) { va = fl == 1 ? b : a + 2; return F_D; }
    if (c0 == 'R' && (fl == 1 || s[a + 1] == '=')) { va = fl == 1 ? b : a + 2; return F_R; }
    if (fl >= 2 && (c0 == 'C' || c0 == 'S') && s[a + 1] == 'P' && (fl == 2 || s[a + 2] == '=')) {
        va = fl == 2 ? b : a + 3;
        return c0 == 'C' ? F_CP : F_SP;
    }
    if (c0 == 'P' && fl >= 2 && is_dig(s[a + 1])) {
        int e = a + 1, pid = 0;
        while (e < b && is_dig(s[e])) { pid = pid * 10 + (s[e] - '0'); if (pid > 1000) pid = 1000; e++; }
        if (e < b && s[e] != '=') return F_IGNORE;                      /* key with other characters: not a pattern key */
        /* a pattern key (message_synced.py:50-57): canonical single-digit id and canonical value, or the host decides */
        if (e == b || pid > 9 || !canon_int(s, e + 1, b, val)) return F_HOST;
        id = pid;
        return F_PAT;
    }
    return F_IGNORE;
}

template <bool MU, int LNMAX, int NW>
__global__ void __launch_bounds__(NW * 32) tokenize_kernel(LArgs A)
{
    constexpr bool LONG = LNMAX > LN_MAX;
    __shared__ LineSm<LNMAX> g_ls[NW];
    LineSm<LNMAX> &sm = g_ls[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;

    for (uint32_t it = wid;; it += warps) {
        uint32_t li = it;
        if (LONG) {                               /* second pass: draw from the list the first pass wrote */
            uint32_t t = 0;
            if (lane == 0) { t = atomicAdd(&A.long_cnt[1], 1u); t = t < A.long_cnt[0] ? A.long_list[t] : 0xFFFFFFFFu; }
            li = __shfl_sync(FULL, t, 0);
            if (li == 0xFFFFFFFFu) break;
        } else if (li >= A.n) break;
        const uint32_t off = A.line_off[li];
        const int len = (int)A.line_len[li];
        const uint32_t unit = (off >> 5) + A.base + li;
        int status = SDB_LINE_INVALID;
        /* record defaults: everything zero, cp = 0xFF */
        if (lane < 12) sm.rec[lane] = lane == 8 ? unit : (lane == 9 ? 0xFF000000u : 0u);
        int32_t clock = 0;
        int r_off = 0, r_len = 0, has_r = 0;
        __syncwarp();

        if (len > LNMAX) {
            status = SDB_LINE_HOSTPATH;
            if (!LONG && len <= LN_LONG_MAX) {    /* the second pass tokenizes it (and overwrites what is written below) */
                if (lane == 0) A.long_list[atomicAdd(&A.long_cnt[0], 1u)] = li;
            }
        } else if (len > 0) {
            /* 1. stage (16-byte loads from the aligned address below the line start) */
            const uint32_t ph = off & 15u;
            const uint4 *src = reinterpret_cast<const uint4 *>(A.text + (off - ph));
            const int nq = (int)((ph + (uint32_t)len + 15u) >> 4);
            for (int q = lane; q < nq; q += 32) reinterpret_cast<uint4 *>(sm.buf)[q] = __ldg(&src[q]);
            __syncwarp();
            const uint8_t *s = sm.buf + ph;

            /* 2. separators; a trailing part without ';' is a field too (str.split) */
            int nf = 0;
            bool nonascii = false;
            for (int w0 = 0; w0 < len; w0 += 32) {
                const int p = w0 + lane;
                const uint8_t c = p < len ? s[p] : 0;
                const uint32_t semi = __ballot_sync(FULL, c == ';');
                nonascii |= __any_sync(FULL, c >= 0x80);
                if (c == ';') {
                    const int k = nf + __popc(semi & ((1u << lane) - 1));
                    if (k < 33) sm.fend[k] = (uint16_t)p;
                }
                nf += __popc(semi);
            }
            const bool open_tail = s[len - 1] != ';';
            if (open_tail) { if (lane == 0 && nf < 33) sm.fend[nf] = (uint16_t)len; nf++; }
            __syncwarp();

            if (nonascii || nf > 32) status = SDB_LINE_HOSTPATH;      /* str.isdigit() / \d know more digits than ASCII */
            else {
                /* 3. one lane per field */
                const bool have = lane < nf;
                const int a = have ? (lane ? sm.fend[lane - 1] + 1 : 0) : 0, b = have ? sm.fend[lane] : 0;
                int va = 0, vb = 0, id = 0;
                int32_t val = 0;
                int cls = have ? classify<MU>(s, a, b, va, vb, id, val) : F_EMPTY;
                if (MU && have && lane == 0) cls = (b == 2 && s[0] == 'M' && s[1] == 'U') ? F_OTHER_OK : F_BAD;

                /* 4. dict semantics across the fields */
                const uint32_t m_pat = __ballot_sync(FULL, cls == F_PAT), m_d = __ballot_sync(FULL, cls == F_D);
                const uint32_t m_r = __ballot_sync(FULL, cls == F_R);
                bool host = __any_sync(FULL, cls == F_HOST);
                bool ok = true;
                if (MU) {
                    /* ^(?=.*D=\d+)(?:MU;(?:P[0-7]=-?[0-9]{1,5};){2,8}((?:D=\d{2,};)|(?:CP=\d;)|(?:R=\d+;)|(?:O;)|(?:e;)|(?:p;)|(?:w=\d;))*)$
                     * no alternative of the tail matches a pattern field and vice versa, so the split is unique */
                    const int k = __ffs(~(m_pat >> 1)) - 1;                /* leading pattern fields after "MU" */
                    ok = !open_tail && !__any_sync(FULL, have && cls == F_BAD) && k >= 2 && k <= 8 &&
                         (k + 1 >= 32 || (m_pat >> (k + 1)) == 0) && m_d != 0;
                } else {
                    ok = m_d != 0;                                         /* ms.py:41-43: no D, nothing to demodulate */
                }
                /* the D fields: digits only (MU: every D field, \d{2,}; MS: the last one, non-empty); pack the last */
                const int d_lane = m_d ? 31 - __clz(m_d) : 0;
                int dlen = 0;
                if (ok) {
                    uint32_t todo = MU ? m_d : (1u << d_lane);
                    while (todo && ok) {
                        const int dl = __ffs(todo) - 1;
                        todo &= todo - 1;
                        const int da = __shfl_sync(FULL, va, dl), db = __shfl_sync(FULL, vb, dl);
                        const int n = db - da;
                        const bool last = dl == d_lane;
                        if (n < (MU ? 2 : 1)) { ok = false; break; }
                        if (last && n > SDB_MAX_DIGITS) { host = true; }
                        bool good = true;
                        uint32_t *dst = reinterpret_cast<uint32_t *>(A.pool + (size_t)unit * 16);
                        const int nwords = ((n + 31) >> 5) * 4;
                        for (int w = lane; w * 8 < n || (last && !host && w < nwords); w += 32) {
                            uint32_t x = 0;
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                const int i = 8 * w + j;
                                uint32_t nib = 0xFu;
                                if (i < n) { const uint8_t c = s[da + i]; good = good && is_dig(c); nib = (uint32_t)(c - '0') & 0xFu; }
                                x |= nib << (4 * j);
                            }
                            if (last && !host && w < nwords) dst[w] = x;
                        }
                        if (!__all_sync(FULL, good)) ok = false;
                        if (last) dlen = n;
                    }
                }
                int cp_slot = -1;
                if (ok && !MU) {
                    /* message_synced.py:21-47: CP and SP non-empty digit strings, R too when present (the last of each wins) */
                    const uint32_t m_cp = __ballot_sync(FULL, cls == F_CP), m_sp = __ballot_sync(FULL, cls == F_SP);
                    const bool vdig = (cls == F_CP || cls == F_SP || cls == F_R) && short_digits(s, va, vb);
                    const uint32_t m_vd = __ballot_sync(FULL, vdig);
                    if (!m_cp || !m_sp) ok = false;
                    else {
                        const int lc = 31 - __clz(m_cp), lsp = 31 - __clz(m_sp);
                        if (!((m_vd >> lc) & 1) || !((m_vd >> lsp) & 1)) ok = false;
                        if (m_r && !((m_vd >> (31 - __clz(m_r))) & 1)) ok = false;
                        if (ok) {
                            int cpv = 0;                                    /* str(int(CP)) (message_synced.py:33,59) */
                            if (lane == lc) for (int i = va; i < vb; i++) { cpv = cpv * 10 + (s[i] - '0'); if (cpv > 1000) cpv = 1000; }
                            cpv = __shfl_sync(FULL, cpv, lc);
                            cp_slot = cpv <= 9 ? cpv + 100 : -1;             /* resolved to a slot below */
                        }
                    }
                }
                if (ok) {
                    /* pattern slots in order of first appearance; the same key again overwrites the value in place
                     * (MU: keys are P0..P7; MS: any duplicate id goes to the host, "P1" and "P01" are different keys) */
                    const uint32_t grp = __match_any_sync(FULL, cls == F_PAT ? id : 64 + lane);
                    const bool leader = cls == F_PAT && lane == __ffs(grp) - 1;
                    const uint32_t m_lead = __ballot_sync(FULL, leader);
                    if (!MU && m_lead != m_pat) host = true;
                    const int npat = __popc(m_lead);
                    if (npat > SDB_MAX_SLOTS) host = true;
                    if (!host) {
                        const int32_t v_last = __shfl_sync(FULL, val, 31 - __clz(grp));     /* value of the last field of the group */
                        uint32_t ids_part = 0;
                        if (leader) {
                            const int slot = __popc(m_lead & ((1u << lane) - 1));
                            sm.rec[slot] = (uint32_t)v_last;
                            ids_part = (uint32_t)id << (4 * slot);
                            if (cp_slot == id + 100) cp_slot = slot;
                        }
                        const uint32_t ids = __reduce_or_sync(FULL, ids_part);
                        const int cps = __reduce_max_sync(FULL, (leader && cp_slot < 100) ? cp_slot : -1);
                        __syncwarp();
                        if (lane == 0) {
                            sm.rec[9] = (uint32_t)dlen | ((uint32_t)npat << 16) | ((uint32_t)(cps >= 0 ? cps : 0xFF) << 24);
                            sm.rec[10] = ids;
                            sm.rec[11] = SDB_MSG_VALID;
                        }
                        __syncwarp();
                        if (cps >= 0) { const int32_t pv = (int32_t)sm.rec[cps]; clock = pv < 0 ? -pv : pv; }
                        if (m_r) {
                            const int lr = 31 - __clz(m_r);
                            r_off = __shfl_sync(FULL, va, lr); r_len = __shfl_sync(FULL, vb, lr) - r_off; has_r = 1;
                        }
                        status = SDB_LINE_OK;
                    }
                }
                if (host && (MU ? ok : true)) status = status == SDB_LINE_OK ? SDB_LINE_OK : SDB_LINE_HOSTPATH;
                if (!ok && !MU) status = SDB_LINE_INVALID;               /* the gates fail before any pattern is looked at */
            }
        }
        __syncwarp();
        if (status != SDB_LINE_OK && lane < 12) sm.rec[lane] = lane == 8 ? unit : (lane == 9 ? 0xFF000000u : 0u);
        __syncwarp();
        if (lane < 12) reinterpret_cast<uint32_t *>(&A.msgs[li])[lane] = sm.rec[lane];
        if (lane == 0) {
            SdbLineInfo inf;
            inf.status = (uint8_t)status; inf.has_r = (uint8_t)(status == SDB_LINE_OK ? has_r : 0);
            inf.r_len = (uint16_t)(status == SDB_LINE_OK ? r_len : 0); inf.r_off = (uint32_t)(status == SDB_LINE_OK ? r_off : 0);
            inf.clock = status == SDB_LINE_OK ? clock : 0;
            A.info[li] = inf;
        }
        __syncwarp();
    }
}

size_t lines_pool_bytes(size_t text_len, uint32_t n) { return ((text_len >> 5) + (size_t)n + 4) * 16; }

int launch_tokenize(int kind, const uint8_t *d_text, const uint32_t *d_off, const uint32_t *d_len, uint32_t n, uint32_t base,
                    SdbPulseMsg *d_msgs, uint8_t *d_pool, SdbLineInfo *d_info, uint32_t *d_long, int sm_count, cudaStream_t stream)
{
    if (n == 0) return 0;
    LArgs A;
    A.text = d_text; A.line_off = d_off; A.line_len = d_len; A.n = n; A.base = base; A.msgs = d_msgs; A.pool = d_pool; A.info = d_info;
    A.long_list = d_long + 2; A.long_cnt = d_long;
    cudaError_t e = cudaMemsetAsync(d_long, 0, 2 * sizeof(uint32_t), stream);
    if (e != cudaSuccess) return (int)e;
    uint32_t need = (n + LN_WARPS - 1) / LN_WARPS;
    uint32_t grid = (uint32_t)sm_count * 24;       /* (8 CTAs per SM: 101.9 M lines/s through the whole text path, 24: 103.2 M) */
    if (need < grid) grid = need;
    const uint32_t grid_long = (uint32_t)sm_count * 4;
    if (kind == SDB_KIND_MU) {
        tokenize_kernel<true, LN_MAX, LN_WARPS><<<grid, LN_WARPS * 32, 0, stream>>>(A);
        tokenize_kernel<true, LN_LONG_MAX, LN_LONG_WARPS><<<grid_long, LN_LONG_WARPS * 32, 0, stream>>>(A);
    } else {
        tokenize_kernel<false, LN_MAX, LN_WARPS><<<grid, LN_WARPS * 32, 0, stream>>>(A);
        tokenize_kernel<false, LN_LONG_MAX, LN_LONG_WARPS><<<grid_long, LN_LONG_WARPS * 32, 0, stream>>>(A);
    }
    return (int)cudaGetLastError();
}

}  // namespace sdb
