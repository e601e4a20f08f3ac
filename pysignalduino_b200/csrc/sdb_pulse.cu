/*
 * sdb_pulse.cu — MS / MU batch demodulation kernel for sm_100a.
 *
 * Replaces, for a whole batch of packed messages,
 *   demodulate_ms      sd_protocols/message_synced.py:10-243
 *   demodulate_mu      sd_protocols/message_unsynced.py:11-296
 *   pattern_exists     sd_protocols/pattern_utils.py:34-136
 *   length_in_range / bin_str_2_hex_str   sd_protocols/helpers.py:28-64, :124-166
 *   postDemo_*         sd_protocols/postdemodulation.py (sdb_postdemod.cuh)
 *
 * Work decomposition: ONE WARP PER MESSAGE, persistent warps striding over the batch, several SMALL kernels per
 * message class (each kernel's hot code has to fit the 32 KB L1.5 instruction cache — see DESIGN.md §4.1):
 *   resolve_kernel<MS|MU>   stage the message (nibble-packed digits by coalesced 128-bit loads; digit / digram first-
 *                           and last-occurrence tables so that "target in D" is a table lookup; tenths table
 *                           T[clock][slot] = 10*round(P/clock, 1)), candidate-slot masks per (clock, interval)
 *                           pair + kill masks = prefilter, then one LANE per surviving protocol: exact
 *                           pattern_exists at thread level (warp level for long starts / 1- and 4-digit symbols)
 *                           -> 16-byte survivor records in protocol-table order.
 *   scan_kernel<MS>         survivors -> chunk classification by ballots -> bits -> finish_match.
 *   mu_match_kernel         survivors 32 at a time, one per lane: distinct symbol / start bitmaps built once per
 *                           message (SWAR), each lane walks re.finditer for its own survivor over the shared
 *                           bitmaps -> 4-byte match records in reference order.
 *   mu_emit_kernel          match records -> bits (ballots) -> finish_match.
 *   scan_kernel<MU>         fused match + emit, only for messages with more than MU_MCAP matches.
 *   finish_match            length rules, padding, post-demodulation, modulematch, hit staging.
 * Hits of one message are staged in shared memory and published with one atomicAdd, so
 * they are contiguous and already in reference order (protocol order, then match order).
 *
 * Float parity: the only float64 work is x = P/clock and CPython's round(x, 1), done with
 * correctly-rounded division and an FMA residual; every tolerance /
 * gap comparison is an integer compare against tables the host derived by running the
 * reference's own float expressions.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sdb200.h"
#include "sdb_table.h"
#include "sdb_postdemod.cuh"
#include "sdb_pulse.h"

/*
 * This file is compiled TWICE:
 *   sdb_pulse.cu       namespace sdb       the fast kernels: D <= SDB_FAST_DIGITS (1024) digits staged per warp, 8 warps / CTA
 *   sdb_pulse_long.cu  namespace sdb_long  (#define SDB_PULSE_LONG + #include of this file) the same code sized for
 *                                          D <= SDB_MAX_DIGITS (4096): 2 warps / CTA, resolve + fused scan only
 * The fast resolve kernel appends every message longer than SDB_FAST_DIGITS to a per-launch list; the long kernels
 * (tiny persistent grids that exit at once when the list is empty) decode exactly those, so the reference's "no length
 * limit on D" (message_unsynced.py:22-25) holds up to the firmware-scale cap without taxing the common case.
 */
#ifdef SDB_PULSE_LONG
#define KNS sdb_long
#define KMAXD SDB_MAX_DIGITS
#define KTHREADS SDB_LONG_THREADS
#define KMIN_CTAS 1
#define KRESOLVE_CTAS 1
#define POS_BITS 13                   /* p <= 4095, n <= 4096 */
#else
#define KNS sdb
#define KMAXD SDB_FAST_DIGITS
#define KTHREADS SDB_PULSE_THREADS
#define KMIN_CTAS SDB_PULSE_MIN_CTAS
#ifndef SDB_MATCH_CTAS
#define SDB_MATCH_CTAS SDB_PULSE_MIN_CTAS
#endif
#ifndef SDB_EMIT_CTAS
#define SDB_EMIT_CTAS SDB_PULSE_MIN_CTAS
#endif
#ifndef SDB_RESOLVE_CTAS
#define SDB_RESOLVE_CTAS SDB_PULSE_MIN_CTAS
#endif
#define KRESOLVE_CTAS SDB_RESOLVE_CTAS
#define KMATCH_CTAS SDB_MATCH_CTAS    /* per-kernel register caps (65536 / (256 * CTAs)) for the two smaller MU kernels */
#define KEMIT_CTAS SDB_EMIT_CTAS
#define POS_BITS 11                   /* p <= 1023, n <= 1024 */
#endif
#define POS_MASK ((1u << POS_BITS) - 1u)
/* SdbSurv.meta: bits 0..12 where the scanned text begins (s0 / message_start), bit 15: `float` resolved */
#define SURV_POS_MASK 0x1FFFu
#define SURV_HASF 0x8000u

namespace KNS {
#ifdef SDB_PULSE_LONG
using sdb::gbit; using sdb::sbit; using sdb::postdemod;      /* sdb_postdemod.cuh lives in namespace sdb */
#endif

/* Message digits, survivor and match records are touched once per kernel: loading them without allocating in L1 keeps them
 * from evicting the protocol table (83 KB, read by every lane all the time); L2 still keeps them for the next kernel of the
 * launch group. */
#ifdef SDB_STREAM_HINTS
__device__ __forceinline__ uint4 ld_stream(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ld_stream(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
#else
__device__ __forceinline__ uint4 ld_stream(const uint4 *p) { return __ldg(p); }
__device__ __forceinline__ uint32_t ld_stream(const uint32_t *p) { return *p; }
#endif
#define LD_STREAM(p) ld_stream(p)

#define FULL 0xffffffffu
/* Functions with exactly ONE call site per kernel are inlined: no code growth, and a call is expensive here (at the 64-register
 * cap the ABI saves and restores the live registers around it) — emit kernel −24 %, MS scan −9 %, match −3 % (profiles/README.md).
 * Functions with several call sites stay out of line (instruction-cache footprint); the per-message drivers of the match and
 * emit kernels have a second, rare call site (re-run writing in place), which goes through an out-of-line wrapper.
 * The resolve kernels are a local optimum in BOTH directions: inlining tpre / resolve_general / resolve_*_warp costs 3-9 %,
 * moving prepare_tables / stage_message / thread_resolve_mu out of line costs 9-18 % (measured, profiles/README.md). */
#define FN_ONE_SITE __device__ __forceinline__
#define DIG_WORDS (KMAXD / 8 + 4)
#define BIT_WORDS (KMAXD / 32 + 4)
#define ST_HITS 24
#define ST_WORDS 112
#define NONE32 0xffffffffu
#define WARPS (KTHREADS / 32)

/* compute-sanitizer is not available on the GPU pool, so the library can be built with its own bounds checks
 * (-DSDB_BOUNDS_CHECK, python -m pysignalduino_b200.build_ext --check): every data-dependent shared-memory index
 * goes through IDX(), violations are counted in g_sdb_oob and read back with sdb_debug_violations(). */
#ifdef SDB_BOUNDS_CHECK
__device__ unsigned int g_sdb_oob = 0;
__device__ __forceinline__ int sdb_chk_idx(int i, int n) { if ((unsigned)i >= (unsigned)n) { atomicAdd(&g_sdb_oob, 1u); return 0; } return i; }
#define IDX(i, n) sdb_chk_idx((int)(i), (int)(n))
#define SDB_CHK(c) do { if (!(c)) atomicAdd(&g_sdb_oob, 1u); } while (0)
#else
#define IDX(i, n) (i)
#define SDB_CHK(c) do { } while (0)
#endif

typedef SdbPulseArgs KArgs;

/* Messages differ a lot in cost (dlen 20..1024, 0..129 survivors), so a static message -> warp map leaves a tail at the
 * end of every launch (measured: chunks of 1 M instead of 262 144 messages were 7 % faster).  Warps therefore draw
 * their next messages from a per-launch counter. */
#ifdef SDB_PULSE_LONG
#define KLIST_ALWAYS true
#else
#define KLIST_ALWAYS false
#endif
/* LIST: the launch draws from a list an earlier kernel of the group wrote (complete: stream order) — the long kernels (messages
 * of more than SDB_FAST_DIGITS digits) and the overflow pass of the fast resolve kernel; lpos = position in that list. */
template <bool LIST>
__device__ __forceinline__ bool next_message(const KArgs &A, uint32_t &base, uint32_t &left, uint32_t &mi, uint32_t &lpos)
{
    if (LIST || KLIST_ALWAYS) {
        (void)left;
        uint32_t b = 0, v = 0xFFFFFFFFu;
        if ((threadIdx.x & 31) == 0) {
            b = atomicAdd(A.ticket, 1u);
            if (b < min(*A.list_cnt, A.list_max)) v = A.list[b];
        }
        base = __shfl_sync(0xffffffffu, v, 0);
        if (base == 0xFFFFFFFFu) return false;
        lpos = __shfl_sync(0xffffffffu, b, 0);
        mi = base;
        return true;
    }
    if (!left) {
        uint32_t b = 0;
        if ((threadIdx.x & 31) == 0) b = atomicAdd(A.ticket, A.ticket_batch);
        base = __shfl_sync(0xffffffffu, b, 0);
        if (base >= A.n) return false;
        left = min(A.ticket_batch, A.n - base);
    }
    mi = base++;
    left--;
    lpos = 0;
    return true;
}

/* A record the packed domain cannot represent (flagged by the packer, or malformed: > 8 slots, an id > 9, D longer than
 * SDB_MAX_DIGITS): no hits and status SDB_ST_DOMAIN — reported per message, never decoded differently from the reference. */
__device__ __forceinline__ bool msg_domain(const SdbPulseMsg *m)
{
    const uint32_t fl = m->flags;
    if (fl & SDB_MSG_DOMAIN) return true;
    if (!(fl & SDB_MSG_VALID)) return false;
    if (m->dlen > SDB_MAX_DIGITS || m->npat > SDB_MAX_SLOTS) return true;
    const uint32_t ids = m->pat_ids;
    bool bad = false;
    for (int s = 0; s < (int)m->npat; s++) bad |= ((ids >> (4 * s)) & 0xF) > 9;
    return bad;
}

/* MU scan kernel: symbol / start bitmaps of the distinct id-string sets of up to 32 survivors */
#define MU_NB 6                       /* symbol bitmap triples resident at a time             */
#define MU_NS 6                       /* start bitmaps resident at a time                     */
#define MU_BW (KMAXD / 32 + 2)        /* words per bitmap: KMAXD positions + 2 zero words     */
#define MU_K 4                        /* matches a lane records before the warp emits them    */
#define MU_MCAP 64                    /* match records per message handed to the emit kernel  */
#define MU_MARK 0xFFFFFFFFu           /* more than that: the fused fallback kernel takes the message */

struct __align__(16) WarpSm {
    uint32_t dig[DIG_WORDS];          /* nibble-packed digits, 0xF beyond dlen               */
    union {
        struct {                      /* resolve kernels */
            uint32_t first2[100], last2[100]; /* digram ab: first position / last position + 1       */
            uint32_t first1[12], last1[12];   /* digit a                                             */
            uint32_t cnt2[100];               /* digram ab: occurrences at even | odd << 16 positions */
            uint16_t T[SDB_MAX_CLK][8];       /* tenths per (clock, slot), biased: T_BIAS + clamp(t, +-T_CLAMP) in 0 .. 0x7FFF so that two
                                               * of them are range-checked per 32-bit operation; T_EMPTY = empty / unused slot       */
            uint8_t  M[SDB_MAX_VALS];         /* candidate-slot mask per (clock, interval) pair      */
            uint8_t  plist[256];              /* table rows that passed the prefilter, in table order */
            uint32_t dead[SDB_KILL_WORDS];    /* bit r: protocol row r lacks a candidate slot for a mandatory value */
        };
        struct {                      /* scan_kernel<MU> */
            uint32_t Bm[MU_NB][3][MU_BW];     /* [0] bit p: a symbol of the set starts at position p; [1] / [2]: 8 / 16
                                               * consecutive symbols (p, p + w, ...) start at p      */
            uint32_t Sm[MU_NS][MU_BW];        /* bit p: the start string occurs at position p        */
        };
    };
    uint32_t val[BIT_WORDS];          /* bit plane of the current match (LSB-first)          */
    uint32_t fpl[BIT_WORDS];          /* 'F' plane                                           */
    uint32_t tmp[BIT_WORDS];          /* post-demodulation output                            */
    SdbHit   st_hits[ST_HITS];        /* staged hits of the current message                  */
    uint32_t st_bits[ST_WORDS];
    int32_t  pat[8];
    int32_t  pd_rc, pd_no;
    /* per-message scalars (so that the out-of-line helpers need few arguments) */
    const uint16_t *rank;
    int32_t  dlen, npat;
    uint32_t pat_ids, msg;
    uint32_t blk_off, blk_left;       /* resolve / match kernels: the warp's current block of the compact arena (kept here, not
                                       * in registers: the message loops are at the 64-register cap already)                */
    /* hit sink */
    uint32_t nh, nw;                  /* hits / words produced so far                        */
    uint32_t hbase, wbase;            /* second pass: global bases                           */
    int32_t  direct, overflow;
};

__shared__ WarpSm g_sm[WARPS];
#define SM() (g_sm[threadIdx.x >> 5])

/* CTA-shared copies of what every lane of the resolve kernels reads all the time (dynamic shared memory, filled once per CTA):
 * the (clock, interval) pairs and the hot fields of the protocol rows.  The full 248-byte rows stay in global memory for the
 * rare paths (several candidates, long starts); with each lane on a different row those reads were the main source of
 * long-scoreboard stalls (profiles/README.md, round 2). */
struct HotKey {                       /* 12 bytes */
    uint8_t  len, nuniq;
    uint16_t vidx[2];                 /* candidate-mask rows of the first two distinct values */
    uint16_t rsv;
    uint32_t uidx;
};
struct HotRow {                       /* 52 bytes */
    HotKey   key[4];
    uint8_t  width, clk_idx;
    int16_t  regex_min;
};
extern __shared__ __align__(16) uint8_t g_dyn[];
#define HOT_VALS() (reinterpret_cast<const SdbValRow *>(g_dyn))
#define HOT_ROWS(nvals) (reinterpret_cast<const HotRow *>(g_dyn + (((nvals) * sizeof(SdbValRow) + 15) & ~(size_t)15)))
/* then, 16-byte aligned, doubles: MU the clocks (n_clk values and n_clk times 10 / clock), MS the clock of every protocol row */
#define HOT_CLK(nvals, nrows) (reinterpret_cast<const double *>(g_dyn + (((nvals) * sizeof(SdbValRow) + 15) & ~(size_t)15) + (((nrows) * sizeof(HotRow) + 15) & ~(size_t)15)))
__host__ __device__ __forceinline__ size_t hot_bytes(uint32_t nvals, uint32_t nrows)
{
    const size_t ndbl = nrows > 2 * SDB_MAX_CLK ? nrows : 2 * SDB_MAX_CLK;       /* MU: clocks and 10 / clock; MS: one clock per protocol row */
    return ((nvals * sizeof(SdbValRow) + 15) & ~(size_t)15) + (((size_t)nrows * sizeof(HotRow) + 15) & ~(size_t)15) + ndbl * sizeof(double) + 16;
}

/* biased tenths (see WarpSm::T): every accept interval of a compiled table lies well inside +-T_CLAMP (table.py checks) */
#define T_BIAS 0x4000
#define T_CLAMP 16000
#define T_EMPTY 0x7FFFu
#define T_GET(x) ((int)(x) - T_BIAS)
__device__ __forceinline__ uint32_t t_biased(int t) { return (uint32_t)(min(max(t, -T_CLAMP), T_CLAMP) + T_BIAS); }

/* ---- small helpers ------------------------------------------------------------------- */
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

/* 8 digits starting at position p (nibble-packed, low nibble first) */
__device__ __forceinline__ uint32_t win32(const uint32_t *dig, int p)
{
    int wi = p >> 3, sh = (p & 7) * 4;
    SDB_CHK(p >= 0 && wi + 1 < DIG_WORDS);
    return __funnelshift_r(dig[wi], dig[wi + 1], sh);
}
/* 16 digits starting at position p */
__device__ __forceinline__ uint64_t win64(const uint32_t *dig, int p)
{
    int wi = p >> 3, sh = (p & 7) * 4;
    SDB_CHK(p >= 0 && wi + 2 < DIG_WORDS);
    uint32_t a = dig[wi], b = dig[wi + 1], c = dig[wi + 2];
    return ((uint64_t)__funnelshift_r(b, c, sh) << 32) | __funnelshift_r(a, b, sh);
}
__device__ __forceinline__ uint64_t nibmask64(int n) { return n >= 16 ? ~0ull : ((1ull << (4 * n)) - 1); }
__device__ __forceinline__ uint32_t nibmask32(int n) { return n >= 8 ? ~0u : ((1u << (4 * n)) - 1); }

/* 10 * round(p / c, 1) as an integer: CPython float.__round__(x, 1) on x = p / c (float64). */
__device__ __forceinline__ int tenths(int p, double c)
{
    double x = __ddiv_rn((double)p, c);
    double y = __dmul_rn(x, 10.0);
    double e = __fma_rn(x, 10.0, -y);          /* 10x = y + e exactly */
    double r = rint(y);                        /* ties-to-even on y */
    double d = __dsub_rn(y, r);
    if (d == 0.5 || d == -0.5) {               /* y sits on a tie: the exact value decides */
        double fl = floor(y);
        if (e > 0.0) r = fl + 1.0;
        else if (e < 0.0) r = fl;
    }
    if (r > 32000.0) r = 32000.0;              /* far outside every accept interval */
    if (r < -32000.0) r = -32000.0;
    return (int)r;
}

__device__ __noinline__ int tenths_cold(int p, double c) { return tenths(p, c); }   /* the rare exact path of tenths_fast, out of the hot loop */

/* Same value, usually without the division: y' = p * (10/c) differs from the exact 10*RN(p/c) by < 1e-10,
 * so whenever y' is further than 1e-6 from a tie both round to the same integer; only near-ties (and
 * they do occur: p/c = 0.25, 0.35, ...) take the exact path. */
__device__ __forceinline__ int tenths_fast(int p, double c, double inv10c)
{
    double y = __dmul_rn((double)p, inv10c);
    if (y >= 40000.0) return 32000;
    if (y <= -40000.0) return -32000;
    double r = rint(y);
    double d = fabs(__dsub_rn(y, r));
    if (fabs(d - 0.5) > 1e-6) {
        if (r > 32000.0) r = 32000.0;
        if (r < -32000.0) r = -32000.0;
        return (int)r;
    }
    return tenths_cold(p, c);
}

/* ---- warp-cooperative substring search: first p >= from with D[p:p+L] == tgt, else -1 ------- */
__device__ __noinline__ int warp_find(uint64_t tgt, int L, int from)
{
    WarpSm &sm = SM();
    const int dlen = sm.dlen;
    uint64_t m = nibmask64(L);
    for (int base = from; base + L <= dlen; base += 32) {
        int p = base + lane_id();
        bool hit = false;
        if (p + L <= dlen) hit = (win64(sm.dig, p) & m) == tgt;      /* never read past the staged digits */
        uint32_t bal = __ballot_sync(FULL, hit);
        if (bal) return base + __ffs(bal) - 1;
    }
    return -1;
}

/* Is the id string `tg` (L digits, nibble-packed) a substring of D[from:] ?  (pattern_utils.py:133) */
__device__ __forceinline__ bool target_present(WarpSm &sm, uint64_t tg, int L, int from, bool want_pos, int &pos)
{
    const int d0 = (int)(tg & 0xF), d1 = (int)((tg >> 4) & 0xF);
    if (L == 1) {
        if (sm.last1[IDX(d0, 12)] <= (uint32_t)from) return false;
        if (want_pos) pos = (int)sm.first1[IDX(d0, 12)];           /* want_pos callers search from 0 */
        return true;
    }
    if (L == 2) {
        if (sm.last2[IDX(d0 * 10 + d1, 100)] <= (uint32_t)from) return false;
        if (want_pos) pos = (int)sm.first2[IDX(d0 * 10 + d1, 100)];
        return true;
    }
    int prev = d0;
    for (int i = 1; i < L; i++) {                         /* necessary: every digram occurs late enough */
        int d = (int)((tg >> (4 * i)) & 0xF);
        if (sm.last2[IDX(prev * 10 + d, 100)] <= (uint32_t)(from + i - 1)) return false;
        prev = d;
    }
    int p = warp_find(tg, L, from);
    if (p < 0) return false;
    pos = p;
    return true;
}

__constant__ uint32_t k_inv16[9] = {0u, 65536u, 32768u, 21846u, 16384u, 13108u, 10923u, 9363u, 8192u};   /* ceil(2^16 / d) */

/*
 * General pattern_exists (pattern_utils.py:34-136): some distinct value has several candidate
 * slots, so the candidates are ordered by gap rank (then slot) and the cartesian product is
 * walked in itertools.product order, 32 combinations per step.
 */
__device__ __noinline__ bool resolve_general(const SdbKeyTpl *__restrict__ k, int t_slot, int from, bool want_pos,
                                             uint64_t &tgt, int &pos, bool in, int lo, uint32_t bal)
{
    WarpSm &sm = SM();
    const int lane = lane_id();
    const int L = k->len, K = k->nuniq;
    const int v = lane >> 3, j = lane & 7;
    const uint32_t pat_ids = sm.pat_ids;
    int key = 0x7fffffff;
    if (in) key = ((int)__ldg(&sm.rank[k->rank_off[v] + (t_slot - lo)]) << 3) | j;   /* gap rank, then slot order (:83 stable sort) */
    int cnt[SDB_MAX_UNIQ];
#pragma unroll
    for (int u = 0; u < SDB_MAX_UNIQ; u++) cnt[u] = __popc((bal >> (8 * u)) & 0xff);
    int ord = 0;
#pragma unroll 1
    for (int o = 1; o < 8; o++) {
        int other = __shfl_sync(FULL, key, (lane & ~7) | ((j + o) & 7));
        ord += other < key;
    }
    uint32_t list = in ? ((uint32_t)j << (4 * ord)) : 0u;
    list |= __shfl_xor_sync(FULL, list, 1);
    list |= __shfl_xor_sync(FULL, list, 2);
    list |= __shfl_xor_sync(FULL, list, 4);
    uint32_t lists[SDB_MAX_UNIQ];
#pragma unroll
    for (int u = 0; u < SDB_MAX_UNIQ; u++) lists[u] = __shfl_sync(FULL, list, 8 * u);
    int total = 1;
#pragma unroll
    for (int u = 0; u < SDB_MAX_UNIQ; u++) if (u < K) total *= cnt[u];
    /* total <= 8^4 = 4096 < 10000: the explosion guard (:97-101) can never fire with <= 8 slots */
    const uint32_t uidx = k->uidx;
    for (int base = 0; base < total; base += 32) {                  /* :111 product order, last list fastest */
        int c = base + lane;
        bool ok = c < total;
        int slot[SDB_MAX_UNIQ] = {0, 0, 0, 0};
        int rem = ok ? c : 0;
#pragma unroll
        for (int u = SDB_MAX_UNIQ - 1; u >= 0; u--) {
            if (u < K) {
                /* rem / cnt[u] without an integer division: cnt <= 8, rem < 8^4, so (rem * ceil(2^16 / cnt)) >> 16 is exact */
                const int q = (int)(((uint32_t)rem * k_inv16[IDX(cnt[u], 9)]) >> 16);
                const int ch = rem - q * cnt[u];
                rem = q;
                slot[u] = (lists[u] >> (4 * ch)) & 0xF;
            }
        }
#pragma unroll
        for (int a = 0; a < SDB_MAX_UNIQ; a++)
#pragma unroll
            for (int b = a + 1; b < SDB_MAX_UNIQ; b++)
                if (b < K && slot[a] == slot[b]) ok = false;        /* :114 one id for two values */
        uint64_t tg = 0;
        int prev = 0;
        for (int i = 0; i < L; i++) {                               /* :118-127 */
            int u = (uidx >> (2 * i)) & 3;
            int s = u == 0 ? slot[0] : (u == 1 ? slot[1] : (u == 2 ? slot[2] : slot[3]));
            int d = (pat_ids >> (4 * s)) & 0xF;
            tg |= (uint64_t)d << (4 * i);
            if (ok) {
                if (L == 1) ok = sm.last1[IDX(d, 12)] > (uint32_t)from;
                else if (i > 0) ok = sm.last2[IDX(prev * 10 + d, 100)] > (uint32_t)(from + i - 1);   /* digram filter (exact for L == 2) */
            }
            prev = d;
        }
        uint32_t good = __ballot_sync(FULL, ok);
        while (good) {                                              /* :133 first combination present in D wins */
            int w = __ffs(good) - 1;
            good &= good - 1;
            uint32_t tlo = __shfl_sync(FULL, (uint32_t)tg, w);
            uint32_t thi = __shfl_sync(FULL, (uint32_t)(tg >> 32), w);
            uint64_t cand = ((uint64_t)thi << 32) | tlo;
            if (target_present(sm, cand, L, from, want_pos, pos)) { tgt = cand; return true; }
        }
    }
    return false;
}

/*
 * pattern_exists for one template, warp-cooperative.  lane = (distinct value v, slot j).
 *   t_slot : this lane's tenths value for slot (lane & 7);  from : the searched text is D[from:]
 * Fast path: every distinct value has exactly one candidate slot -> a single combination.
 */
__device__ __noinline__ bool resolve_key(const SdbKeyTpl *__restrict__ k, int t_slot, int from, bool want_pos,
                                         uint64_t &tgt, int &pos)
{
    WarpSm &sm = SM();
    const int lane = lane_id();
    const int L = k->len, K = k->nuniq;
    const int v = lane >> 3, j = lane & 7;
    bool in = false;
    int lo = 0;
    if (v < K && j < sm.npat) {                                     /* :73-76 candidates = tenths inside [lo, hi] */
        lo = k->lo[v];
        in = t_slot >= lo && t_slot <= k->hi[v];
    }
    const uint32_t bal = __ballot_sync(FULL, in);
    const uint32_t b0 = bal & 0xff, b1 = (bal >> 8) & 0xff, b2 = (bal >> 16) & 0xff, b3 = bal >> 24;
    if (!b0 || (K > 1 && !b1) || (K > 2 && !b2) || (K > 3 && !b3)) return false;       /* :78-80 */
    const bool single = !(b0 & (b0 - 1)) && !(b1 & (b1 - 1)) && !(b2 & (b2 - 1)) && !(b3 & (b3 - 1));
    if (!single) return resolve_general(k, t_slot, from, want_pos, tgt, pos, in, lo, bal);
    /* one combination: distinct sentinels keep absent values out of the duplicate test (:114) */
    const int s0 = __ffs(b0) - 1, s1 = K > 1 ? __ffs(b1) - 1 : 8, s2 = K > 2 ? __ffs(b2) - 1 : 9, s3 = K > 3 ? __ffs(b3) - 1 : 10;
    if (s0 == s1 || s0 == s2 || s0 == s3 || s1 == s2 || s1 == s3 || s2 == s3) return false;
    const uint32_t ids = sm.pat_ids;
    const int d0 = (ids >> (4 * s0)) & 0xF, d1 = (ids >> (4 * (s1 & 7))) & 0xF, d2 = (ids >> (4 * (s2 & 7))) & 0xF,
              d3 = (ids >> (4 * (s3 & 7))) & 0xF;
    const uint32_t uidx = k->uidx;
    uint64_t tg = 0;
    for (int i = 0; i < L; i++) {                                   /* :118-127 */
        int u = (uidx >> (2 * i)) & 3;
        int d = u == 0 ? d0 : (u == 1 ? d1 : (u == 2 ? d2 : d3));
        tg |= (uint64_t)d << (4 * i);
    }
    if (!target_present(sm, tg, L, from, want_pos, pos)) return false;
    tgt = tg;
    return true;
}

/* ---- payload characters (for modulematch, message_unsynced.py:254-280) ------------------- */
struct Payload {
    const SdbPulseProto *pp;
    const uint32_t *val, *fpl;
    int nb;          /* bits after padding */
    int ndig;        /* hex digits ceil(nb/4) */
    int lz;          /* leading '0' digits removed by remove_zero */
    int body;        /* length of the middle part */
    bool has_f, bin;
};
__device__ __forceinline__ int hex_digit(const uint32_t *val, int nb, int ndig, int j)
{
    /* right-aligned nibbles (helpers.py:28-64): digit j covers bits nb-4(ndig-j) .. +3 */
    int b0 = nb - 4 * (ndig - j), v = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { int bi = b0 + k; v = (v << 1) | (bi >= 0 ? gbit(val, bi) : 0); }
    return v;
}
__device__ int payload_char(const Payload &P, int i)
{
    const SdbPulseProto *pp = P.pp;
    if (i < pp->pre_len) return (unsigned char)pp->preamble[i];
    i -= pp->pre_len;
    if (i < P.body) {
        if (P.bin) return gbit(P.fpl, i) ? 'F' : ('0' + gbit(P.val, i));
        if (P.has_f) return "None"[i];
        int d = hex_digit(P.val, P.nb, P.ndig, i + P.lz);
        return d < 10 ? '0' + d : 'A' + d - 10;
    }
    i -= P.body;
    if (i < pp->post_len) return (unsigned char)pp->postamble[i];
    return -1;
}
/* modulematch (message_unsynced.py:277-280) as a fixed-offset character-class program */
FN_ONE_SITE bool modulematch(const SdbPulseProto *pp, const SdbMmItem *__restrict__ items, int nb, bool has_f)
{
    WarpSm &sm = SM();
    const int flags = pp->flags;
    Payload P;
    P.pp = pp; P.val = sm.val; P.fpl = sm.fpl; P.nb = nb; P.ndig = (nb + 3) >> 2;
    P.has_f = has_f; P.bin = (flags & SDB_PF_DISPATCH_BIN) != 0; P.lz = 0;
    if (P.bin) P.body = nb;
    else if (has_f) P.body = 4;
    else {
        if (flags & SDB_PF_REMOVE_ZERO) {                 /* lstrip('0') :268-269 */
            int z = 0;
            while (z < P.ndig && hex_digit(sm.val, nb, P.ndig, z) == 0) z++;
            P.lz = z;
        }
        P.body = P.ndig - P.lz;
    }
    const int total = pp->pre_len + P.body + pp->post_len;
    const bool end = (flags & SDB_PF_MM_END) != 0;
    int pos = pp->pre_len;
    const int n = pp->mm_nitems;
    for (int it = 0; it < n; it++) {
        const SdbMmItem *m = &items[pp->mm_off + it];
        int mn = m->min, mx = m->max;
        int cnt = 0;
        int lim = (end && it == n - 1) ? mx : mn;
        if (m->mask[0] == 0xFFFFFBFFu && (m->mask[1] & m->mask[2] & m->mask[3]) == 0xFFFFFFFFu) {
            cnt = min(lim, total - pos);          /* '.': the payload never contains a newline, only the length matters */
            pos += cnt;
            if (cnt < mn) return false;
            continue;
        }
        /* one lane per character: the atom consumes the leading characters that are in its class, up to lim */
        const int avail = min(lim, total - pos);
        cnt = max(avail, 0);
        for (int b0 = 0; b0 < avail; b0 += 32) {
            const int j = b0 + lane_id();
            bool bad = false;
            if (j < avail) {
                const int c = payload_char(P, pos + j);
                bad = c < 0 || c >= 128 || !((m->mask[c >> 5] >> (c & 31)) & 1);
            }
            const uint32_t bm = __ballot_sync(FULL, bad);
            if (bm) { cnt = b0 + __ffs(bm) - 1; break; }
        }
        pos += cnt;
        if (cnt < mn) return false;
    }
    if (end && pos != total) return false;
    return true;
}

/* emit_hit <- finish_match <- mu_emit_match / scan_ms: each has exactly one call site per kernel (inside the out-of-line
 * mu_emit_records / scan_survivors_mu / scan_survivors), so inlining the chain costs no code and saves three calls per match */
#define CHAIN_FN FN_ONE_SITE

/* ---- hit sink: shared-memory staging, or direct global writes on the rare second pass ------ */
CHAIN_FN void emit_hit(const KArgs &A, const SdbPulseProto *pp, int nb, bool has_f, int ordinal, bool mm_host)
{
    WarpSm &sm = SM();
    const int lane = lane_id();
    const int nwv = (nb + 31) >> 5;
    const int nw = has_f ? 2 * nwv : nwv;
    SdbHit h;
    h.msg = sm.msg; h.proto = pp->proto; h.nbits = (uint16_t)nb; h.aux = (uint16_t)ordinal;
    h.flags = (has_f ? SDB_HIT_HAS_F : 0) | (mm_host ? SDB_HIT_MM_HOST : 0); h.rsv = 0;
    const uint32_t nh0 = sm.nh, nw0 = sm.nw;
    __syncwarp();
    if (!sm.direct) {
        if (nh0 < ST_HITS && nw0 + nw <= ST_WORDS && !sm.overflow) {
            h.bits_off = nw0;
            if (lane == 0) sm.st_hits[IDX(nh0, ST_HITS)] = h;
            for (int i = lane; i < nw; i += 32) sm.st_bits[IDX(nw0 + i, ST_WORDS)] = i < nwv ? sm.val[IDX(i, BIT_WORDS)] : sm.fpl[IDX(i - nwv, BIT_WORDS)];
        } else if (lane == 0) sm.overflow = 1;
    } else {
        h.bits_off = sm.wbase + nw0;
        if (sm.hbase + nh0 < A.hits_cap && sm.wbase + nw0 + nw <= A.bits_cap) {
            if (lane == 0) A.hits[sm.hbase + nh0] = h;
            for (int i = lane; i < nw; i += 32) A.bits[sm.wbase + nw0 + i] = i < nwv ? sm.val[IDX(i, BIT_WORDS)] : sm.fpl[IDX(i - nwv, BIT_WORDS)];
        }
    }
    __syncwarp();
    if (lane == 0) { sm.nh = nh0 + 1; sm.nw = nw0 + nw; }
    __syncwarp();
}

/* nb rounded up to a multiple of paddingbits (1, 4 or 8 in the shipped table: no integer division) */
__device__ __forceinline__ int pad_up(int nb, int pad)
{
    return (pad & (pad - 1)) == 0 ? (nb + pad - 1) & ~(pad - 1) : (nb + pad - 1) / pad * pad;
}

/* post-demodulation on lane 0 (rare: ~10 of 129 protocols, frames <= ~150 bits) */
FN_ONE_SITE void run_postdemod(int method, int nb)
{
    WarpSm &sm = SM();
    for (int w = lane_id(); w < BIT_WORDS; w += 32) sm.tmp[IDX(w, BIT_WORDS)] = 0;
    __syncwarp();
    if (lane_id() == 0) {
        int no = 0;
        sm.pd_rc = postdemod(method, sm.val, nb, sm.tmp, &no);
        sm.pd_no = no;
    }
    __syncwarp();
}

/* Shared tail of an MS / MU match: the bits are in sm.val / sm.fpl (nb of them, zero beyond).
 * Returns SDB_ST_* (non-OK aborts the message). */
template <bool MS>
CHAIN_FN int finish_match(const KArgs &A, const SdbPulseProto *pp, int nb, int ordinal, bool maybe_f)
{
    WarpSm &sm = SM();
    const int lane = lane_id();
    const int flags = pp->flags;
    __syncwarp();
    bool has_f = false;
    if (maybe_f) {
        uint32_t fany = 0;
        for (int i = lane; i < ((nb + 31) >> 5); i += 32) fany |= sm.fpl[IDX(i, BIT_WORDS)];
        has_f = __any_sync(FULL, fany != 0);
    }
    if (MS) {
        /* length_in_range (helpers.py:124-166), message_synced.py:194 */
        if (pp->lir_min != -1 && nb < pp->lir_min) return SDB_ST_OK;
        if ((flags & SDB_PF_HAS_LIR_MAX) && nb > pp->lir_max) return SDB_ST_OK;
    }
    const int pad = pp->padbits;
    if (MS) nb = pad_up(nb, pad);                 /* message_synced.py:198-200: pad BEFORE postDemod (appended bits are 0 already) */

    if (pp->postdemod) {
        if (has_f) {
            if (MS) return SDB_ST_VALUEERROR;     /* message_synced.py:209 int('F') escapes */
            /* MU: ValueError swallowed (message_unsynced.py:249), bits kept */
        } else {
            run_postdemod(pp->postdemod, nb);
            int rc = sm.pd_rc, no = sm.pd_no;
            if (rc == -2) {
                if (MS) return SDB_ST_VALUEERROR;
            } else {
                if (rc < 1) return SDB_ST_OK;
                if (!MS || no > 0) {              /* MS keeps the old bits when ret_bits is empty (:218) */
                    for (int w = lane; w < BIT_WORDS; w += 32) sm.val[IDX(w, BIT_WORDS)] = sm.tmp[IDX(w, BIT_WORDS)];
                    nb = no;
                }
            }
            __syncwarp();
        }
    }
    if (!MS) nb = pad_up(nb, pad);                /* message_unsynced.py:257-259: pad AFTER postDemod */

    bool mm_host = false;
    if (MS) {
        if (has_f) return SDB_ST_OK;              /* bin_str_2_hex_str -> None (:224-226) */
    } else {
        if (flags & SDB_PF_MM_NEVER) return SDB_ST_OK;
        if (flags & SDB_PF_MM_HOST) mm_host = true;      /* a regex shape the program cannot express: the host formatter applies it */
        else if (pp->mm_off != 0xFFFF && !modulematch(pp, A.tab.mm, nb, has_f)) return SDB_ST_OK;   /* :277-280 */
    }
    emit_hit(A, pp, nb, has_f, ordinal, mm_host);
    return SDB_ST_OK;
}

/* exact per-byte equality flags (0x80 in every byte of x equal to the byte c) */
__device__ __forceinline__ uint32_t eq_bytes(uint32_t x, uint32_t c4)
{
    uint32_t z = x ^ c4;
    return ~(((z & 0x7f7f7f7fu) + 0x7f7f7f7fu) | z) & 0x80808080u;
}
/* 0x80 flags of bytes 0..3 -> bits 0,2,4,6 */
__device__ __forceinline__ uint32_t spread_even(uint32_t t)
{
    uint32_t u = t >> 7;
    return (u | (u >> 6) | (u >> 12) | (u >> 18)) & 0x55u;
}

/* exact per-nibble equality of the 8 digits of x with digit c -> bits 0..7 */
__device__ __forceinline__ uint32_t eq_nibbles8(uint32_t x, uint32_t c8)
{
    uint32_t z = x ^ c8;
    uint32_t u = (~(((z & 0x77777777u) + 0x77777777u) | z) & 0x88888888u) >> 3;   /* bit 4k = digit k equal */
    u = (u | (u >> 3)) & 0x03030303u;
    u = (u | (u >> 6)) & 0x000F000Fu;
    return (u | (u >> 12)) & 0xFFu;
}

/*
 * The regex scan of the MU survivors of one message: message_unsynced.py:146-290.
 *
 * re.finditer("START((?:S1|S2|S3){MIN,}(?:E1|..)?)", D') is backtracking-free because all symbol alternatives have
 * the same width w: a match at i needs START at i and a run of n >= MIN symbols at p = i + len(START),
 * p, p + w, ... (greedy: up to the first non-symbol of that residue class), then an optional tail.  Two bitmaps
 * describe everything: B (a symbol of the set starts here) and S (the start string occurs here).  Survivors of one
 * message share them heavily (most resolve to the same two id strings), so the warp builds each DISTINCT bitmap once
 * (SWAR compares, 8 windows per lane) and then every LANE walks the matches of its own survivor with word scans over
 * the shared bitmaps; the warp only cooperates again to turn a match into bits (ballots), post-demodulate and emit.
 */
/* warp: bitmap of the positions where a w-digit symbol of {c1, c0, cf} starts -> dst[0 .. nw + 1] */
FN_ONE_SITE void mu_build_B(uint32_t *dst, int w, uint32_t c1, uint32_t c0, uint32_t cf, int nw)
{
    WarpSm &sm = SM();
    const int lane = lane_id();
    const int dlen = sm.dlen;
    if (w == 2) {
        /* each lane compares the 8 two-digit windows of one digit word against the symbol bytes; windows reaching
         * past dlen contain 0xF padding and never equal a symbol (digits <= 9) */
        const uint32_t k1 = c1 * 0x01010101u, k0 = c0 * 0x01010101u, kf = cf * 0x01010101u;
        const int nwords = (dlen + 7) >> 3;
        uint8_t *d8 = reinterpret_cast<uint8_t *>(dst);
        const int nbytes = 4 * (nw + 2);
#pragma unroll 1
        for (int wi = lane; wi < nbytes; wi += 32) {
            uint32_t b8 = 0;
            if (wi < nwords) {
                uint32_t x = sm.dig[IDX(wi, DIG_WORDS)], nx = sm.dig[IDX(wi + 1, DIG_WORDS)];
                uint32_t y = __funnelshift_r(x, nx, 4);                /* windows at odd positions */
                uint32_t be = eq_bytes(x, k1) | eq_bytes(x, k0) | eq_bytes(x, kf);
                uint32_t bo = eq_bytes(y, k1) | eq_bytes(y, k0) | eq_bytes(y, kf);
                b8 = spread_even(be) | (spread_even(bo) << 1);
            }
            d8[IDX(wi, 4 * MU_BW)] = (uint8_t)b8;
        }
    } else if (w == 1) {
        /* one-digit symbols: 8 digits per lane, plain nibble compares (0xF padding never equals a digit) */
        const uint32_t k1 = (c1 & 0xF) * 0x11111111u, k0 = (c0 & 0xF) * 0x11111111u, kf = (cf & 0xF) * 0x11111111u;
        const int nwords = (dlen + 7) >> 3;
        uint8_t *d8 = reinterpret_cast<uint8_t *>(dst);
        const int nbytes = 4 * (nw + 2);
#pragma unroll 1
        for (int wi = lane; wi < nbytes; wi += 32) {
            uint32_t b8 = 0;
            if (wi < nwords) {
                const uint32_t x = sm.dig[IDX(wi, DIG_WORDS)];
                b8 = eq_nibbles8(x, k1) | eq_nibbles8(x, k0) | eq_nibbles8(x, kf);
            }
            d8[IDX(wi, 4 * MU_BW)] = (uint8_t)b8;
        }
    } else {
        const uint32_t wm = nibmask32(w);
#pragma unroll 1
        for (int r = 0; r <= nw + 1; r++) {
            int p = r * 32 + lane;
            uint32_t x = p < dlen ? win32(sm.dig, p) & wm : 0xFFFFFFFFu;
            bool sym = (p + w <= dlen) && (x == c1 || x == c0 || x == cf);
            uint32_t bw = __ballot_sync(FULL, sym);
            if (lane == 0) dst[IDX(r, MU_BW)] = bw;
        }
    }
    __syncwarp();
    /* "8 / 16 symbols in a row start here" by shifted-AND doubling on the words: the per-lane scans take their candidates
     * from these, so the many short accidental runs cost nothing */
#ifdef SDB_PULSE_LONG
    /* more words than lanes: the doubling runs in place on plane [2], block after block in ascending order (a word is
     * combined with its still unmodified successor), plane [1] is the copy taken after the third level */
    uint32_t *P1 = dst + MU_BW, *P2 = dst + 2 * MU_BW;
    for (int j = lane; j < MU_BW; j += 32) P2[IDX(j, MU_BW)] = j <= nw + 1 ? dst[IDX(j, MU_BW)] : 0u;
    __syncwarp();
#pragma unroll 1
    for (int lv = 0; lv < 4; lv++) {
        const int sh = w << lv;                                        /* w, 2w, 4w, 8w <= 32 */
#pragma unroll 1
        for (int b0 = 0; b0 < MU_BW; b0 += 32) {
            const int j = b0 + lane;
            uint32_t x = 0, nx = 0;
            if (j < MU_BW) { x = P2[IDX(j, MU_BW)]; nx = j + 1 < MU_BW ? P2[IDX(j + 1, MU_BW)] : 0u; }
            __syncwarp();
            if (j < MU_BW) P2[IDX(j, MU_BW)] = x & (sh < 32 ? __funnelshift_r(x, nx, sh) : nx);
            __syncwarp();
        }
        if (lv == 2) { for (int j = lane; j < MU_BW; j += 32) P1[IDX(j, MU_BW)] = P2[IDX(j, MU_BW)]; __syncwarp(); }
    }
#else
    /* (lane j = word j; word 32 is all zero) */
    uint32_t x = lane <= nw + 1 ? dst[IDX(lane, MU_BW)] : 0u;
    uint32_t x8 = 0;
#pragma unroll
    for (int lv = 0; lv < 4; lv++) {
        const int sh = w << lv;                                        /* w, 2w, 4w, 8w <= 32 */
        uint32_t nx = __shfl_down_sync(FULL, x, 1);
        if (lane == 31) nx = 0;
        x &= sh < 32 ? __funnelshift_r(x, nx, sh) : nx;
        if (lv == 2) x8 = x;
    }
    dst[IDX(MU_BW + lane, 3 * MU_BW)] = x8;
    dst[IDX(2 * MU_BW + lane, 3 * MU_BW)] = x;
    if (lane < 2) { dst[IDX(MU_BW + 32 + lane, 3 * MU_BW)] = 0; dst[IDX(2 * MU_BW + 32 + lane, 3 * MU_BW)] = 0; }
    __syncwarp();
#endif
}

/* warp: bitmap of the occurrences of the Ls-digit start string -> dst[0 .. nw + 1] */
FN_ONE_SITE void mu_build_S(uint32_t *dst, int Ls, uint64_t start_t, int nw)
{
    WarpSm &sm = SM();
    const int lane = lane_id();
    const int dlen = sm.dlen;
    if (Ls <= 2) {
        const uint32_t ks = Ls == 2 ? (uint32_t)start_t * 0x01010101u : ((uint32_t)start_t & 0xF) * 0x11111111u;
        const int nwords = (dlen + 7) >> 3;
        uint8_t *d8 = reinterpret_cast<uint8_t *>(dst);
        const int nbytes = 4 * (nw + 2);
#pragma unroll 1
        for (int wi = lane; wi < nbytes; wi += 32) {
            uint32_t s8 = 0;
            if (wi < nwords) {
                uint32_t x = sm.dig[IDX(wi, DIG_WORDS)];
                if (Ls == 2) {
                    uint32_t y = __funnelshift_r(x, sm.dig[IDX(wi + 1, DIG_WORDS)], 4);
                    s8 = spread_even(eq_bytes(x, ks)) | (spread_even(eq_bytes(y, ks)) << 1);
                } else s8 = eq_nibbles8(x, ks);                        /* one-pulse start: plain digit compare */
            }
            d8[IDX(wi, 4 * MU_BW)] = (uint8_t)s8;
        }
    } else {
        const uint64_t lm = nibmask64(Ls);
#pragma unroll 1
        for (int r = 0; r <= nw + 1; r++) {
            int p = r * 32 + lane;
            bool st = (p + Ls <= dlen) && ((win64(sm.dig, p) & lm) == start_t);
            uint32_t sw = __ballot_sync(FULL, st);
            if (lane == 0) dst[IDX(r, MU_BW)] = sw;
        }
    }
    __syncwarp();
}

/* per-lane description and progress of one survivor, packed so that the whole scan state stays in registers */
#define MU_NONE 0xFFFE
struct MuLane {
    const uint32_t *B, *Bx, *S;  /* Bx: B, or its 8- / 16-in-a-row level when MIN allows; S == nullptr: no start string */
    uint32_t sym;                /* id strings of one | zero << 8 ... kept as c1, c0, cf below */
    uint32_t c1, c0, cf;         /* one / zero / float (zero, float = one when absent) */
    int lw, Ls, MIN, lenmax;
    uint32_t flags;              /* 1 use_tail, 2 has0, 4 hasf, 8 done */
    int pos, status, nm;
    uint32_t m0, m1, m2, m3;     /* recorded matches: p | n << POS_BITS | (tail + 1) << 2 POS_BITS */
    uint64_t dead;               /* 16 bits per residue class of p: no run of >= MIN symbols starts before this position */
    uint64_t cand;               /* 16 bits per residue class of p: cached next candidate i + 1 (0 unknown, 0xFFFF none) */
};
#define MU_F_TAIL 1u
#define MU_F_HAS0 2u
#define MU_F_HASF 4u
#define MU_F_DONE 8u

__device__ __forceinline__ uint32_t mu_cls_mask(int lw, int p)
{
    return lw == 0 ? FULL : (lw == 1 ? 0x55555555u << (p & 1) : 0x11111111u << (p & 3));
}

/* lane: one step of re.finditer for this survivor = the leftmost remaining candidate, accepted (recorded) or rejected.
 * A candidate is a position i with S[i] and (MIN == 0 or Bx[i + Ls]). */
__device__ __forceinline__ void mu_step(const uint32_t *dig, MuLane &L, int nwB)
{
    const int w = 1 << L.lw, Ls = L.Ls;
    const int pos = L.pos;
    /* leftmost candidate over the residue classes; a class is dead up to the end of a run that was too short.
     * p = i + Ls in class c  <=>  i in class (c - Ls) mod w */
    int i = MU_NONE;
#pragma unroll 1
    for (int c = 0; c < w; c++) {
        const int from = max(pos, (int)((uint32_t)(L.dead >> (16 * c)) & 0xFFFFu) - Ls);
        int cd = (int)((uint32_t)(L.cand >> (16 * c)) & 0xFFFFu) - 1;
        if (cd < from) {
            const uint32_t cmask = mu_cls_mask(L.lw, (c - Ls) & 3);
            int j = from >> 5;
            uint32_t first = FULL << (from & 31);
            cd = MU_NONE;
#pragma unroll 2
            for (; j <= nwB; j++) {
                uint32_t cw = cmask & first;
                first = FULL;
                if (L.MIN > 0) cw &= Ls ? __funnelshift_r(L.Bx[IDX(j, MU_BW)], L.Bx[IDX(j + 1, MU_BW)], Ls) : L.Bx[IDX(j, MU_BW)];
                if (L.S) cw &= L.S[IDX(j, MU_BW)];
                if (cw) { cd = j * 32 + __ffs(cw) - 1; break; }
            }
            L.cand = (L.cand & ~(0xFFFFull << (16 * c))) | ((uint64_t)(cd + 1) << (16 * c));
        }
        i = min(i, cd);
    }
    if (i == MU_NONE) { L.flags |= MU_F_DONE; return; }
    const int p = i + Ls;
    /* greedy run: first position q >= p, q = p (mod w), that is not a symbol (B is zero from dlen on) */
    const uint32_t cls = mu_cls_mask(L.lw, p);
    int ns;
    {
        int j = p >> 5;
        uint32_t z = ~L.B[IDX(j, MU_BW)] & cls & (FULL << (p & 31));
        while (!z) { j++; z = ~L.B[IDX(j, MU_BW)] & cls; }           /* ends at word (dlen >> 5) <= nwB + 1 at the latest */
        ns = j * 32 + __ffs(z) - 1;
    }
    const int n = (ns - p) >> L.lw;
    if (n < L.MIN) {                                                    /* {MIN,} not met: every start inside this run fails too */
        const int sh = 16 * (p & (w - 1));
        L.dead = (L.dead & ~(0xFFFFull << sh)) | ((uint64_t)ns << sh);
        L.pos = i + 1;
        return;
    }
    int end = p + n * w;
    int tail = -1;                                                      /* 0/1/2 = bit '1'/'0'/'F' of the reconstructed chunk */
    if (L.flags & MU_F_TAIL) {
        const uint32_t em = nibmask32(w - 1);
        uint32_t x = win32(dig, end) & em;                              /* beyond dlen the 0xF padding never matches */
        if (x == (L.c1 & em)) tail = 0;
        else if ((L.flags & MU_F_HAS0) && x == (L.c0 & em)) tail = 1;
        else if ((L.flags & MU_F_HASF) && x == (L.cf & em)) tail = 2;
        if (tail >= 0) end += w - 1;
    }
    L.pos = end > i ? end : i + 1;
    const int nch = n + (tail >= 0 ? 1 : 0);
    if (nch == 0) { L.status = SDB_ST_INDEXERROR; L.flags |= MU_F_DONE; return; }   /* :212 chunks[-1] on an empty capture */
    if (L.lenmax >= 0 && nch > L.lenmax) return;                        /* :217 */
    const uint32_t rec = (uint32_t)p | ((uint32_t)n << POS_BITS) | ((uint32_t)(tail + 1) << (2 * POS_BITS));
    if (L.nm == 0) L.m0 = rec; else if (L.nm == 1) L.m1 = rec; else if (L.nm == 2) L.m2 = rec; else L.m3 = rec;
    L.nm++;
}

/* warp: one match -> bits (:220-228: later keys overwrite earlier ones on identical strings) -> finish_match */
CHAIN_FN int mu_emit_match(const KArgs &A, const SdbPulseProto *pp, uint32_t rec, uint32_t c1, uint32_t c0,
                                          uint32_t cf, bool hasf, int ordinal)
{
    WarpSm &sm = SM();
    const int lane = lane_id();
    const int w = pp->width;
    const bool has0 = pp->key[2].len != 0;
    const uint32_t wm = nibmask32(w);
    const int p = rec & POS_MASK, n = (rec >> POS_BITS) & POS_MASK, tail = (int)(rec >> (2 * POS_BITS)) - 1;
    __syncwarp();
#pragma unroll 1
    for (int b0 = 0; b0 < n; b0 += 32) {
        int c = b0 + lane;
        bool isf = false, one = false;
        if (c < n) {
            uint32_t x = win32(sm.dig, p + c * w) & wm;
            if (hasf && x == cf) isf = true;
            else if (has0 && x == c0) one = false;
            else one = true;
        }
        uint32_t vw = __ballot_sync(FULL, one), fw = __ballot_sync(FULL, isf);
        if (lane == 0) { sm.val[IDX(b0 >> 5, BIT_WORDS)] = vw; sm.fpl[IDX(b0 >> 5, BIT_WORDS)] = fw; }
    }
    __syncwarp();
    {
        /* only the words padding can reach need clearing (pad <= 64 bits): the ballots above wrote every word below them, the
         * consumers read (nb_padded + 31) / 32 words */
        const int wq = ((n + 31) >> 5) + lane;
        if (lane < 4 && wq < BIT_WORDS) { sm.val[IDX(wq, BIT_WORDS)] = 0; sm.fpl[IDX(wq, BIT_WORDS)] = 0; }
    }
    __syncwarp();
    if (lane == 0) {
        if (tail == 0) sm.val[IDX(n >> 5, BIT_WORDS)] |= 1u << (n & 31);
        else if (tail == 2) sm.fpl[IDX(n >> 5, BIT_WORDS)] |= 1u << (n & 31);
    }
    return finish_match<false>(A, pp, n + (tail >= 0 ? 1 : 0), ordinal, hasf);
}

__device__ __forceinline__ uint64_t mu_sort3_key(uint32_t a, uint32_t b, uint32_t c, int w)
{
    uint32_t lo = min(a, min(b, c)), hi = max(a, max(b, c));
    uint32_t mid = a + b + c - lo - hi;
    if (mid == hi) mid = lo;                    /* {x, y} is (x, x, y) whichever of the two the absent keys duplicate */
    return ((uint64_t)w << 48) | ((uint64_t)lo << 32) | ((uint64_t)mid << 16) | hi;
}

/* REC = true: matches are appended to `recs` (survivor index << 24 | p | n << 11 | (tail + 1) << 22) for the emit kernel and
 * nrec counts them (SDB_ST_MU_OVERFLOW when they do not fit); REC = false: every match is emitted on the spot. */
#define SDB_ST_MU_OVERFLOW 0x7F
template <bool REC>
FN_ONE_SITE int scan_survivors_mu_impl(const KArgs &A, const SdbSurv *slots, uint32_t nsurv, uint32_t *recs, uint32_t &nrec)
{
    WarpSm &sm = SM();
    const int lane = lane_id();
    const SdbPulseProto *rows = A.tab.mu;
    const int nwB = (sm.dlen - 1) >> 5;                                /* candidates live in words 0 .. nwB; words 0 .. nwB + 1 are valid */
#pragma unroll 1
    for (uint32_t s0i = 0; s0i < nsurv; s0i += 32) {
        const bool have = s0i + lane < nsurv;
        SdbSurv mine;
        mine.start = 0; mine.c1 = mine.c0 = mine.cf = 0; mine.meta = 0;
        if (have) *reinterpret_cast<uint4 *>(&mine) = LD_STREAM(reinterpret_cast<const uint4 *>(&slots[s0i + lane]));   /* coalesced 16-byte loads */
        const int cnt = min(32u, nsurv - s0i);
        const uint32_t row = have ? (uint32_t)(mine.start >> 56) : 0u;
        const uint64_t start_t = mine.start & 0x00FFFFFFFFFFFFFFull;
        MuLane L;
        int w, lvl;
        {
            const SdbPulseProto *pp = &rows[row];
            w = pp->width;
            L.lw = w == 1 ? 0 : (w == 2 ? 1 : 2);
            L.Ls = pp->key[0].len;
            const bool has0 = pp->key[2].len != 0, hasf = (mine.meta & SURV_HASF) != 0;
            L.c1 = mine.c1; L.c0 = has0 ? mine.c0 : mine.c1; L.cf = hasf ? mine.cf : mine.c1;
            L.MIN = pp->regex_min; L.lenmax = pp->mu_len_max;
            /* w == 1: the tail key is '' and matches nothing extra */
            L.flags = (((pp->flags & SDB_PF_RECONSTRUCT) && w > 1) ? MU_F_TAIL : 0u) | (has0 ? MU_F_HAS0 : 0u) | (hasf ? MU_F_HASF : 0u) |
                      (have ? 0u : MU_F_DONE);
            lvl = L.MIN >= 16 ? 2 : (L.MIN >= 8 ? 1 : 0);
        }
        L.sym = 0;
        L.B = nullptr; L.Bx = nullptr; L.S = nullptr;
        L.pos = mine.meta & SURV_POS_MASK; L.status = SDB_ST_OK; L.nm = 0; L.m0 = L.m1 = L.m2 = L.m3 = 0;
        L.dead = 0; L.cand = 0;
        int ordbase = 0;
        const uint64_t keyB = mu_sort3_key(L.c1, L.c0, L.cf, w);
        const uint64_t keyS = ((uint64_t)L.Ls << 56) | start_t;

        /* Rounds: build the distinct bitmaps of the pending lanes (at most MU_NB / MU_NS per round) and let the lanes whose
         * bitmaps are resident scan until they are done or hold MU_K unread matches; once nothing is pending, emit in survivor
         * order.  A survivor with more than MU_K matches becomes the only pending lane of a further round. */
        uint32_t pending = __ballot_sync(FULL, have);
        int k = 0, rounds = 0;
        bool cont = false;
#pragma unroll 1
        for (;;) {
            const bool pend = (pending >> lane) & 1;
            bool now = pend;
            if (!(cont && rounds == 1)) {                               /* (a continuation after a single round finds every bitmap still resident) */
            const uint32_t gB = __match_any_sync(FULL, pend ? keyB : (0xF000000000000000ull | lane));
            const bool needS = pend && L.Ls != 0;
            const uint32_t gS = __match_any_sync(FULL, needS ? keyS : (0xF000000000000000ull | lane));
            const uint32_t leadB = __ballot_sync(FULL, pend && lane == __ffs(gB) - 1);
            const uint32_t leadS = __ballot_sync(FULL, needS && lane == __ffs(gS) - 1);
            const int slotB = __popc(leadB & ((1u << (__ffs(gB) - 1)) - 1));
            const int slotS = needS ? __popc(leadS & ((1u << (__ffs(gS) - 1)) - 1)) : 0;
            now = pend && slotB < MU_NB && slotS < MU_NS;
            __syncwarp();
            uint32_t lb = leadB;
#pragma unroll 1
            for (int sl = 0; lb && sl < MU_NB; sl++) {
                const int src = __ffs(lb) - 1;
                lb &= lb - 1;
                mu_build_B(sm.Bm[sl][0], __shfl_sync(FULL, w, src), __shfl_sync(FULL, L.c1, src), __shfl_sync(FULL, L.c0, src),
                           __shfl_sync(FULL, L.cf, src), nwB);
            }
            uint32_t ls = leadS;
#pragma unroll 1
            for (int sl = 0; ls && sl < MU_NS; sl++) {
                const int src = __ffs(ls) - 1;
                ls &= ls - 1;
                const uint32_t lo = __shfl_sync(FULL, (uint32_t)start_t, src), hi = __shfl_sync(FULL, (uint32_t)(start_t >> 32), src);
                mu_build_S(sm.Sm[sl], __shfl_sync(FULL, L.Ls, src), ((uint64_t)hi << 32) | lo, nwB);
            }
            if (now) {
                L.B = sm.Bm[IDX(slotB, MU_NB)][0];
                L.Bx = sm.Bm[IDX(slotB, MU_NB)][lvl];
                L.S = L.Ls ? sm.Sm[IDX(slotS, MU_NS)] : nullptr;
            }
            rounds += rounds < 2;
            }
            /* warp-synchronous on purpose: one step per lane per trip, reconverged at every trip, so the cost is the LONGEST
             * lane's step count and not the sum over lanes */
            bool run = now && !(L.flags & MU_F_DONE) && L.nm < MU_K;
#pragma unroll 1
            while (__any_sync(FULL, run)) {
                if (run) {
                    mu_step(sm.dig, L, nwB);
                    run = !(L.flags & MU_F_DONE) && L.nm < MU_K;
                }
                __syncwarp();
            }
            pending &= ~__ballot_sync(FULL, now);
            if (pending) continue;
            /* an exception anywhere loses every hit of the message (it escapes demodulate_mu): no need to emit first */
            if (__any_sync(FULL, L.status != SDB_ST_OK)) return SDB_ST_INDEXERROR;
            if (REC) {
                /* lanes k .. q flush their recorded matches, q = the first lane that is not done yet (it continues alone) */
                const uint32_t notdone = __ballot_sync(FULL, !(L.flags & MU_F_DONE)) & (FULL << k);
                const int q = notdone ? __ffs(notdone) - 1 : cnt;
                const int mine_n = (lane >= k && lane <= q && lane < cnt) ? L.nm : 0;
                int incl = mine_n;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int t = __shfl_up_sync(FULL, incl, d);
                    if (lane >= d) incl += t;
                }
                const uint32_t total = (uint32_t)__shfl_sync(FULL, incl, 31);
                if (nrec + total > MU_MCAP) return SDB_ST_MU_OVERFLOW;
                if (mine_n) {
                    uint32_t *dst = recs + nrec + (incl - mine_n);
                    const uint32_t tag = (s0i + lane) << 24;
                    dst[0] = L.m0 | tag;
                    if (mine_n > 1) dst[1] = L.m1 | tag;
                    if (mine_n > 2) dst[2] = L.m2 | tag;
                    if (mine_n > 3) dst[3] = L.m3 | tag;
                    L.nm = 0;
                }
                nrec += total;
                k = q;
            } else
#pragma unroll 1
            for (; k < cnt; k++) {
                const int nmk = __shfl_sync(FULL, L.nm, k);
                if (nmk) {
                    const SdbPulseProto *pk = &rows[__shfl_sync(FULL, row, k)];
                    const uint32_t c1k = __shfl_sync(FULL, L.c1, k), c0k = __shfl_sync(FULL, L.c0, k), cfk = __shfl_sync(FULL, L.cf, k);
                    const bool hasfk = (__shfl_sync(FULL, L.flags, k) & MU_F_HASF) != 0;
                    const int ob = __shfl_sync(FULL, ordbase, k);
#pragma unroll 1
                    for (int j = 0; j < nmk; j++) {
                        const uint32_t rec = __shfl_sync(FULL, j == 0 ? L.m0 : (j == 1 ? L.m1 : (j == 2 ? L.m2 : L.m3)), k);
                        const int st = mu_emit_match(A, pk, rec, c1k, c0k, cfk, hasfk, ob + j);
                        if (st != SDB_ST_OK) return st;
                    }
                    if (lane == k) { ordbase += L.nm; L.nm = 0; }
                }
                if (!(__shfl_sync(FULL, L.flags, k) & MU_F_DONE)) break;
            }
            if (k >= cnt) break;
            cont = true;
            pending = 1u << k;                                         /* more than MU_K matches: this survivor continues alone */
        }
    }
    return SDB_ST_OK;
}

/* ---- thread-level pattern_exists for templates of <= 2 pulses (<= 2 distinct values) -----------
 * One LANE resolves one protocol: candidates are ordered by (gap rank, slot) by repeated
 * min-extraction from 8 registers, the product is walked in itertools.product order, and
 * "target in D[from:]" is one shared-memory table lookup.  (pattern_utils.py:34-136) */
#define TKEY_NONE 0x7fffffff
__device__ __forceinline__ int tkey_min8(const int k[8])
{
    int m = min(min(min(k[0], k[1]), min(k[2], k[3])), min(min(k[4], k[5]), min(k[6], k[7])));
    return m;
}
__device__ __forceinline__ bool tres(const HotKey &hk, const SdbKeyTpl *__restrict__ k, int clk_idx, const WarpSm &sm, int from,
                                     uint32_t &code, int &pos)
{
    const int L = hk.len, K = hk.nuniq;
    const uint32_t uidx = hk.uidx, ids = sm.pat_ids;
    /* candidate slots of the (<= 2) distinct values (:73-76): looked up in the per-message mask table */
    const uint32_t ca = sm.M[IDX(hk.vidx[0], SDB_MAX_VALS)];
    const uint32_t cb = K > 1 ? sm.M[IDX(hk.vidx[1], SDB_MAX_VALS)] : 0u;
    if (!ca || (K > 1 && !cb)) return false;                  /* :78-80 */
    if (!(ca & (ca - 1)) && !(cb & (cb - 1))) {
        /* the common case: one candidate per value -> a single combination, no ordering needed */
        const int sa = __ffs(ca) - 1;
        const int da = (ids >> (4 * sa)) & 0xF;
        int d0 = da, d1 = da;
        if (K > 1) {
            const int sb = __ffs(cb) - 1;
            if (sb == sa) return false;                       /* :114 */
            const int db = (ids >> (4 * sb)) & 0xF;
            d0 = (uidx & 3) ? db : da; d1 = ((uidx >> 2) & 3) ? db : da;
        }
        if (L == 1) {
            if (sm.last1[IDX(d0, 12)] <= (uint32_t)from) return false;
            code = (uint32_t)d0; pos = (int)sm.first1[IDX(d0, 12)];
        } else {
            if (sm.last2[IDX(d0 * 10 + d1, 100)] <= (uint32_t)from) return false;
            code = (uint32_t)(d0 | (d1 << 4)); pos = (int)sm.first2[IDX(d0 * 10 + d1, 100)];
        }
        return true;
    }
    /* several candidates for some value: order them by (gap rank, slot).  (A sort-free variant — "smallest key above the
     * previous one" scans over the set bits of the mask — was measured: its short divergent loops cost 1 % more than these
     * unrolled 8-register selections.) */
    int ka[8], kb[8];
    const int na = __popc(ca), nb = __popc(cb);
    {
        const uint16_t *__restrict__ rank = sm.rank;
        const int lo0 = k->lo[0], lo1 = K > 1 ? k->lo[1] : 0;
        const uint32_t ro0 = k->rank_off[0], ro1 = K > 1 ? k->rank_off[1] : 0;
        const uint16_t *trow = sm.T[IDX(clk_idx, SDB_MAX_CLK)];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int tj = T_GET(trow[j]);
            ka[j] = ((ca >> j) & 1) ? (((int)__ldg(&rank[ro0 + (tj - lo0)]) << 3) | j) : TKEY_NONE;
            kb[j] = ((cb >> j) & 1) ? (((int)__ldg(&rank[ro1 + (tj - lo1)]) << 3) | j) : TKEY_NONE;
        }
    }
    for (int ia = 0; ia < na; ia++) {                         /* :111 product order: first list slowest */
        const int ma = tkey_min8(ka);
        const int sa = ma & 7;
#pragma unroll
        for (int j = 0; j < 8; j++) if (ka[j] == ma) ka[j] = TKEY_NONE;
        const int da = (ids >> (4 * sa)) & 0xF;
        if (K == 1) {
            bool ok;
            if (L == 1) { ok = sm.last1[IDX(da, 12)] > (uint32_t)from; if (ok) { code = da; pos = (int)sm.first1[IDX(da, 12)]; } }
            else { ok = sm.last2[IDX(da * 11, 100)] > (uint32_t)from; if (ok) { code = da * 0x11; pos = (int)sm.first2[IDX(da * 11, 100)]; } }
            if (ok) return true;
            continue;
        }
        int kc[8];
#pragma unroll
        for (int j = 0; j < 8; j++) kc[j] = kb[j];
        for (int ib = 0; ib < nb; ib++) {
            const int mb = tkey_min8(kc);
            const int sb = mb & 7;
#pragma unroll
            for (int j = 0; j < 8; j++) if (kc[j] == mb) kc[j] = TKEY_NONE;
            if (sb == sa) continue;                           /* :114 one id for two values */
            const int db = (ids >> (4 * sb)) & 0xF;
            const int d0 = (uidx & 3) ? db : da, d1 = ((uidx >> 2) & 3) ? db : da;   /* :118-127 */
            if (sm.last2[IDX(d0 * 10 + d1, 100)] > (uint32_t)from) {    /* :133 */
                code = (uint32_t)(d0 | (d1 << 4));
                pos = (int)sm.first2[IDX(d0 * 10 + d1, 100)];
                return true;
            }
        }
    }
    return false;
}

/* Thread-level necessary condition for a long start (only the warp-level path resolves it exactly): when every distinct
 * value has exactly ONE candidate slot the id string is determined, and each of its digrams must occur in D.
 * false = pattern_exists certainly returns -1. */
__device__ __noinline__ bool tpre(const SdbKeyTpl *__restrict__ k, const WarpSm &sm)
{
    const int L = k->len, K = k->nuniq;
    const uint32_t ids = sm.pat_ids;
    uint32_t dg = 0, seen = 0;                                 /* 4 digits, one nibble each */
    for (int u = 0; u < K; u++) {
        const uint32_t c = sm.M[IDX(k->vidx[u], SDB_MAX_VALS)];
        if (!c) return false;                                  /* pattern_utils.py:78-80 */
        if (c & (c - 1)) return true;                          /* several candidates: undecided here */
        if (seen & c) return false;                            /* :114 one slot cannot stand for two values */
        seen |= c;
        dg |= ((ids >> (4 * (__ffs(c) - 1))) & 0xF) << (4 * u);
    }
    const uint32_t uidx = k->uidx;
    int prev = dg & 0xF;
    for (int i = 1; i < L; i++) {
        const int d = (dg >> (4 * ((uidx >> (2 * i)) & 3))) & 0xF;
        if (sm.last2[IDX(prev * 10 + d, 100)] <= (uint32_t)(i - 1)) return false;
        prev = d;
    }
    return true;
}

/* Thread-level resolution of one MU protocol (2-digit symbols, start of <= 2 pulses).
 * Returns 0 dead, 1 resolved (codes = start | one<<8 | zero<<16 | float<<24, s0f = s0 | hasf<<16),
 * 2 = needs the warp-level path (for a long start: after one / zero passed a pre-screen on the whole D).
 * after_start: the warp has resolved a long start meanwhile (D' begins at s0_in); one / zero / float follow here. */
__device__ __forceinline__ int thread_resolve_mu(const HotRow &hr, const SdbPulseProto *__restrict__ pp, const WarpSm &sm,
                                                 uint32_t &codes, uint32_t &s0f, bool after_start, int s0_in)
{
    const int width = hr.width;
    if (width > 2 || (width == 1 && hr.key[0].len > 2)) return 2;   /* 4-digit symbols: warp-level path for every key */
    const int clk_idx = hr.clk_idx;
    const bool long_start = !after_start && hr.key[0].len > 2;   /* needs a warp-wide search: only pre-screen one / zero here */
    if (long_start && !tpre(&pp->key[0], sm)) return 0;
    uint32_t acc = 0, hasf = 0;
    int s0 = after_start ? s0_in : 0;
#pragma unroll 1
    for (int kk = (long_start || after_start) ? 1 : 0; kk < 4; kk++) {   /* start (:67-88), then one / zero / float (:99-141) */
        const HotKey &hk = hr.key[kk];
        if (!hk.len) continue;
        uint32_t code = 0;
        int p = 0;
        if (!tres(hk, &pp->key[kk], clk_idx, sm, kk == 0 ? 0 : s0, code, p)) {
            if (kk == 3) break;                               /* float is optional (:138) */
            return 0;
        }
        if (kk == 0) s0 = p;                                  /* D' = D[find(start):] */
        if (kk == 3) hasf = 1;
        acc |= code << (8 * kk);
    }
    if (long_start) return 2;                                 /* one / zero exist somewhere in D: worth the warp-level path */
    /* a match needs regex_min consecutive symbols, all at positions of one parity: count them (necessary condition) */
    if (width == 2) {
        const uint32_t c1 = (acc >> 8) & 0xFF, c0 = (acc >> 16) & 0xFF, cf = acc >> 24;
        uint32_t cnt = sm.cnt2[IDX((c1 & 15) * 10 + (c1 >> 4), 100)];
        if (hr.key[2].len && c0 != c1) cnt += sm.cnt2[IDX((c0 & 15) * 10 + (c0 >> 4), 100)];
        if (hasf && cf != c1 && cf != c0) cnt += sm.cnt2[IDX((cf & 15) * 10 + (cf >> 4), 100)];
        const int best = max((int)(cnt & 0xFFFF), (int)(cnt >> 16));
        if (best < (int)hr.regex_min) return 0;
    }
    codes = acc;
    s0f = (uint32_t)s0 | (hasf << 16);
    return 1;
}

/* ---- one (message x MS protocol) task: message_synced.py:90-241 ----------------------------- */
/* Thread-level resolution of one MS protocol (2-digit symbols, sync of <= 2 pulses): message_synced.py:109-163.
 * Returns 0 dead, 1 resolved (codes = sync | one<<8 | zero<<16 | float<<24, msf = message_start | hasf<<16),
 * 2 = needs the warp-level path.  Every pattern_exists call of MS searches the whole D (from = 0). */
__device__ __forceinline__ int thread_resolve_ms(const HotRow &hr, const SdbPulseProto *__restrict__ pp, const WarpSm &sm,
                                                 uint32_t &codes, uint32_t &msf)
{
    if (hr.width != 2 || hr.key[0].len > 2) return 2;
    uint32_t acc = 0, hasf = 0;
    int spos = 0;
#pragma unroll 1
    for (int kk = 0; kk < 4; kk++) {                          /* sync, one, zero, float (:109) */
        const HotKey &hk = hr.key[kk];
        if (!hk.len) continue;
        uint32_t code = 0;
        int p = 0;
        if (!tres(hk, &pp->key[kk], 0, sm, 0, code, p)) {
            if (kk == 3) break;                               /* :160-163 float may be missing */
            return 0;
        }
        if (kk == 0) {                                        /* :140-156 */
            spos = p + hk.len;
            if ((int)hr.regex_min * 2 > sm.dlen - spos) return 0;
        }
        if (kk == 3) hasf = 1;
        acc |= code << (8 * kk);
    }
    codes = acc;
    msf = (uint32_t)spos | (hasf << 16);
    return 1;
}

FN_ONE_SITE int scan_ms(const KArgs &A, const SdbPulseProto *pp, int ms, uint32_t cs, uint32_t c1, uint32_t c0,
                                    uint32_t cf, bool hasf);

/* the chunk loop of one resolved (message x MS protocol) task: message_synced.py:171-241 */
FN_ONE_SITE int scan_ms(const KArgs &A, const SdbPulseProto *pp, int ms, uint32_t cs, uint32_t c1, uint32_t c0,
                                    uint32_t cf, bool hasf)
{
    WarpSm &sm = SM();
    const int lane = lane_id();
    const int dlen = sm.dlen;
    const int w = pp->width;
    const int flags = pp->flags;
    const int Lsy = pp->key[0].len;
    const bool has0 = pp->key[2].len != 0;
    const uint32_t wm = nibmask32(w), em = nibmask32(w - 1), sm_ = nibmask32(Lsy);
    const bool recon = (flags & SDB_PF_RECONSTRUCT) != 0;

    /* chunk loop (:174-189): class per chunk = '1' / '0' / 'F' / skip (sync string) / stop */
    __syncwarp();
    for (int i = lane; i < BIT_WORDS; i += 32) { sm.val[IDX(i, BIT_WORDS)] = 0; sm.fpl[IDX(i, BIT_WORDS)] = 0; }
    __syncwarp();
    int nb = 0;
    const int nchunks = (dlen - ms + w - 1) >> (w == 1 ? 0 : (w == 2 ? 1 : 2));      /* w is 1, 2 or 4 (table.py) */
    for (int b0 = 0; b0 < nchunks; b0 += 32) {
        int c = b0 + lane;
        int cls = 4;                               /* 0 '1', 1 '0', 2 'F', 3 skip, 4 stop, 5 none */
        if (c < nchunks) {
            int q = ms + c * w;
            int cl = min(w, dlen - q);
            uint32_t x = win32(sm.dig, q);
            if (cl == w) {
                uint32_t xs = x & wm;
                if (hasf && xs == cf) cls = 2;
                else if (has0 && xs == c0) cls = 1;
                else if (xs == c1) cls = 0;
                else if (Lsy == w && xs == cs) cls = 3;
                else if (recon) {
                    uint32_t xe = x & em;          /* chunk[:-1] (:182) against the first-wins end table */
                    if (xe == (c1 & em)) cls = 0;
                    else if (has0 && xe == (c0 & em)) cls = 1;
                    else if (hasf && xe == (cf & em)) cls = 2;
                }
            } else {                               /* short last chunk */
                if (cl == Lsy && (x & sm_) == cs) cls = 3;
                else if (recon && cl == w - 1) {
                    uint32_t xe = x & em;
                    if (xe == (c1 & em)) cls = 0;
                    else if (has0 && xe == (c0 & em)) cls = 1;
                    else if (hasf && xe == (cf & em)) cls = 2;
                }
            }
        } else cls = 5;
        uint32_t stop = __ballot_sync(FULL, cls == 4);
        uint32_t live = stop ? ((1u << (__ffs(stop) - 1)) - 1) : FULL;      /* chunks before the first stop */
        uint32_t emitm = __ballot_sync(FULL, cls <= 2) & live;
        uint32_t onem = __ballot_sync(FULL, cls == 0) & live;
        uint32_t fm = __ballot_sync(FULL, cls == 2) & live;
        if (emitm) {
            bool mine = (emitm >> lane) & 1;
            if (mine) {
                int idx = nb + __popc(emitm & ((1u << lane) - 1));
                if ((onem >> lane) & 1) atomicOr(&sm.val[IDX(idx >> 5, BIT_WORDS)], 1u << (idx & 31));
                if ((fm >> lane) & 1) atomicOr(&sm.fpl[IDX(idx >> 5, BIT_WORDS)], 1u << (idx & 31));
            }
            nb += __popc(emitm);
        }
        if (stop) break;
    }
    if (nb == 0) return SDB_ST_OK;                 /* :191 */
    return finish_match<true>(A, pp, nb, 0, hasf);
}

/* ---- phase 0: stage one message ---------------------------------------------------------- */
/* digits + per-message scalars only (all the scan kernel needs) */
__device__ __forceinline__ void stage_digits(const KArgs &A, WarpSm &sm, const SdbPulseMsg *m, int dlen, uint32_t mi)
{
    const int lane = lane_id();
    const uint4 *src = reinterpret_cast<const uint4 *>(A.digits + (size_t)m->doff * 16);
    const int nq = (dlen + 31) >> 5;               /* 16-byte units */
#ifdef SDB_PULSE_LONG
    for (int q = lane; q < nq; q += 32) reinterpret_cast<uint4 *>(sm.dig)[q] = LD_STREAM(&src[q]);
#else
    if (lane < nq) {
        uint4 v = LD_STREAM(&src[lane]);
        reinterpret_cast<uint4 *>(sm.dig)[lane] = v;
    }
#endif
    if (lane < 4) sm.dig[IDX(4 * nq + lane, DIG_WORDS)] = FULL;    /* windows may read 3 words past the last digit */
    if (lane < 8) sm.pat[lane] = m->pat[lane];
    if (lane == 0) {
        sm.rank = A.tab.rank; sm.dlen = dlen; sm.npat = m->npat; sm.pat_ids = m->pat_ids; sm.msg = A.msg_base + mi;
        sm.nh = 0; sm.nw = 0; sm.hbase = 0; sm.wbase = 0; sm.direct = 0; sm.overflow = 0;
    }
    __syncwarp();
}

/* one round of the occurrence tables: 32 positions, one writer per distinct key (match_any) */
__device__ __forceinline__ void occurrence_round(WarpSm &sm, int base, int dlen, int lane)
{
    int p = base + lane;
    uint32_t x = win32(sm.dig, p);
    int a = x & 0xF, b = (x >> 4) & 0xF;
    bool va = p < dlen && a <= 9;
    bool vb = va && (p + 1 < dlen) && b <= 9;
    uint32_t ga = __match_any_sync(FULL, va ? a : 16 + lane);
    uint32_t gb = __match_any_sync(FULL, vb ? a * 10 + b : 128 + lane);
    if (va) {
        if (lane == __ffs(ga) - 1 && sm.first1[IDX(a, 12)] == NONE32) sm.first1[IDX(a, 12)] = p;
        if (lane == 31 - __clz(ga)) sm.last1[IDX(a, 12)] = p + 1;
    }
    if (vb) {
        int code = a * 10 + b;
        if (lane == __ffs(gb) - 1) {
            if (sm.first2[IDX(code, 100)] == NONE32) sm.first2[IDX(code, 100)] = p;
            /* base is a multiple of 32, so lane parity == position parity */
            sm.cnt2[IDX(code, 100)] += (uint32_t)__popc(gb & 0x55555555u) | ((uint32_t)__popc(gb & 0xAAAAAAAAu) << 16);
        }
        if (lane == 31 - __clz(gb)) sm.last2[IDX(code, 100)] = p + 1;
    }
    __syncwarp();
}

/* COMPACT: the MU resolve kernel's hot code sits at the edge of the 32 KB instruction cache, so its loops
 * are not unrolled; the MS kernel (smaller, short messages) keeps the compiler's unrolling. */
template <bool COMPACT>
__device__ __forceinline__ void stage_message(const KArgs &A, WarpSm &sm, const SdbPulseMsg *m, int dlen, uint32_t mi)
{
    const int lane = lane_id();
    for (int i = lane; i < 100; i += 32) { sm.first2[IDX(i, 100)] = NONE32; sm.last2[IDX(i, 100)] = 0; sm.cnt2[IDX(i, 100)] = 0; }
    if (lane < 12) { sm.first1[IDX(lane, 12)] = NONE32; sm.last1[IDX(lane, 12)] = 0; }
    stage_digits(A, sm, m, dlen, mi);
    /* occurrence tables, ascending rounds */
    if (COMPACT) {
#pragma unroll 1
        for (int base = 0; base < dlen; base += 32) occurrence_round(sm, base, dlen, lane);
    } else {
        for (int base = 0; base < dlen; base += 32) occurrence_round(sm, base, dlen, lane);
    }
}

/* =========================================================================================
 * MS and MU each run as TWO kernels so that each has a small instruction footprint (ncu: a fused
 * kernel spends > 50 % of its stall samples waiting for instruction fetch, because the warps of an
 * SM sit in different phases of a large body of code):
 *   resolve_kernel<MS|MU>  template resolution for every protocol -> SdbSurv records
 *                          (message_synced.py:83-163 / message_unsynced.py:59-141)
 *   scan_kernel<MS|MU>     chunk loop / regex scan for every survivor, in protocol-table order
 *                          (message_synced.py:171-241 / message_unsynced.py:146-290)
 * Both are one warp per message; the survivor slots of message i are surv[i*stride .. +stride).
 * ========================================================================================= */

/* warp-level template resolution of one "complex" MU protocol (long start / 1- or 4-digit symbols) */
__device__ __noinline__ bool resolve_mu_warp(const SdbPulseProto *pp, SdbSurv &rec)
{
    WarpSm &sm = SM();
    const int t_slot = T_GET(sm.T[IDX(pp->clk_idx, SDB_MAX_CLK)][lane_id() & 7]);
    int s0 = 0, dummy;
    uint64_t start_t = 0, t1 = 0, t0 = 0, tf = 0;
    if (pp->key[0].len && !resolve_key(&pp->key[0], t_slot, 0, true, start_t, s0)) return false;      /* :67-88 */
    if (!resolve_key(&pp->key[1], t_slot, s0, false, t1, dummy)) return false;                        /* :99-141 */
    if (pp->key[2].len && !resolve_key(&pp->key[2], t_slot, s0, false, t0, dummy)) return false;
    bool hasf = false;
    if (pp->key[3].len) hasf = resolve_key(&pp->key[3], t_slot, s0, false, tf, dummy);
    rec.start = start_t;
    rec.c1 = (uint16_t)t1; rec.c0 = (uint16_t)t0; rec.cf = (uint16_t)tf;
    rec.meta = (uint16_t)(s0 | (hasf ? SURV_HASF : 0));
    return true;
}

/* warp-level resolution of one "complex" MS protocol (4-digit symbols or a 4-pulse sync): message_synced.py:109-163 */
__device__ __noinline__ bool resolve_ms_warp(const SdbPulseProto *pp, SdbSurv &rec)
{
    WarpSm &sm = SM();
    const int t_slot = T_GET(sm.T[IDX(0, SDB_MAX_CLK)][lane_id() & 7]);
    const int w = pp->width;
    uint64_t ts = 0, t1 = 0, t0 = 0, tf = 0;
    int spos = 0, dummy;
    if (!resolve_key(&pp->key[0], t_slot, 0, true, ts, spos)) return false;
    const int ms = spos + pp->key[0].len;
    if ((int)pp->regex_min * w > sm.dlen - ms) return false;          /* :150-156 length_min > (len - start) / width */
    if (!resolve_key(&pp->key[1], t_slot, 0, false, t1, dummy)) return false;
    if (pp->key[2].len && !resolve_key(&pp->key[2], t_slot, 0, false, t0, dummy)) return false;
    bool hasf = false;
    if (pp->key[3].len) hasf = resolve_key(&pp->key[3], t_slot, 0, false, tf, dummy);
    rec.start = ts;
    rec.c1 = (uint16_t)t1; rec.c0 = (uint16_t)t0; rec.cf = (uint16_t)tf;
    rec.meta = (uint16_t)(ms | (hasf ? SURV_HASF : 0));
    return true;
}

/* per-message preparation of the tenths rows and candidate-slot masks; false = the message yields [] */
template <bool MS>
__device__ __forceinline__ bool prepare_tables(const KArgs &A, WarpSm &sm, const SdbPulseMsg *m, double &clock_abs)
{
    const int lane = lane_id();
    const int npat = sm.npat;
    int v0, v1;
    /* A slot whose id digit never occurs in D cannot be part of a target string that is found in D (pattern_utils.py:133), so
     * it is left out of every candidate list: same result (combinations are tried in the same order, the ones dropped could only
     * fail), but more protocols die in the prefilter and fewer values have several candidates.  Such a slot gets the tenths
     * value of an empty slot. */
#ifndef SDB_NO_SLOT_FILTER
#define SLOT_USED(j) (sm.first1[IDX((sm.pat_ids >> (4 * (j))) & 0xF, 12)] != NONE32)
#else
#define SLOT_USED(j) true
#endif
    /* the slots that take part, as a bit mask and as a list of slot numbers (one nibble each, ascending) */
    const bool slot_on = lane < npat && SLOT_USED(lane & 7);
    const uint32_t usedm = __ballot_sync(FULL, slot_on) & 0xFFu;
    if (MS) {
        const int cp = m->cp;
        if (cp == 0xFF) return false;                                /* message_synced.py:60-62 */
        const int pc = sm.pat[cp];
        if (pc == 0) return false;                                   /* :65-66 */
        clock_abs = fabs((double)pc);
        /* tenths of the (<= 8) slots, normalised by the message's own clock (:70-72), into row 0 of T */
        if (lane < 8) sm.T[IDX(0, SDB_MAX_CLK)][lane] = (uint16_t)(slot_on ? t_biased(tenths(sm.pat[lane], clock_abs)) : T_EMPTY);
        v0 = (int)A.tab.n_mu_vals; v1 = (int)A.tab.n_vals;          /* the MS intervals follow the MU pairs */
    } else {
        /* tenths table for every distinct protocol clock (message_unsynced.py:59-64): one lane per (clock, slot that takes
         * part) — 4.9 of 8 slots on average — after the rows have been filled with "empty" */
        const int ncl = A.tab.n_clk;
        const double *hclk = HOT_CLK(A.tab.n_vals, A.tab.n_mu);
        const int nused = __popc(usedm);
        const uint32_t ulist = __reduce_or_sync(FULL, slot_on ? (uint32_t)lane << (4 * __popc(usedm & ((1u << lane) - 1))) : 0u);
#pragma unroll 1
        for (int c = lane; c < ncl; c += 32)
            *reinterpret_cast<uint4 *>(&sm.T[IDX(c, SDB_MAX_CLK)][0]) = make_uint4(T_EMPTY * 0x10001u, T_EMPTY * 0x10001u, T_EMPTY * 0x10001u, T_EMPTY * 0x10001u);
        __syncwarp();
        const uint32_t inv = k_inv16[IDX(nused, 9)];
#pragma unroll 1
        for (int idx = lane; idx < ncl * nused; idx += 32) {
            const int c = (int)(((uint32_t)idx * inv) >> 16);        /* idx / nused (exact: idx < 8^4, nused <= 8) */
            const int j = (int)((ulist >> (4 * (idx - c * nused))) & 7u);
            sm.T[IDX(c, SDB_MAX_CLK)][j] = (uint16_t)t_biased(tenths_fast(sm.pat[j], hclk[c], hclk[ncl + c]));
        }
        v0 = 0; v1 = (int)A.tab.n_mu_vals;
    }
    __syncwarp();
    /* candidate-slot mask of every distinct (clock, accept interval) pair: one lane per pair, two slots per 32-bit operation
     * (biased 15-bit values: bit 15 of (x | 0x8000) - y says x >= y, and no borrow crosses the halfword boundary); a pair
     * without any candidate kills every protocol that needs it (pattern_utils.py:78-80) */
    uint4 ka = make_uint4(0, 0, 0, 0), kb = make_uint4(0, 0, 0, 0);
    const uint32_t H = 0x80008000u;
#pragma unroll 1
    for (int v = v0 + lane; v < v1; v += 32) {
        const SdbValRow vr = HOT_VALS()[v];                          /* CTA-shared copy: lo = biased lower bound, hi = biased upper bound | 0x8000 */
        const uint4 row = *reinterpret_cast<const uint4 *>(&sm.T[IDX(vr.clk_idx, SDB_MAX_CLK)][0]);
        const uint32_t lo2 = (uint32_t)(uint16_t)vr.lo * 0x10001u, hi2 = (uint32_t)(uint16_t)vr.hi * 0x10001u;
        const uint32_t m0 = ((row.x | H) - lo2) & (hi2 - row.x) & H, m1 = ((row.y | H) - lo2) & (hi2 - row.y) & H;
        const uint32_t m2 = ((row.z | H) - lo2) & (hi2 - row.z) & H, m3 = ((row.w | H) - lo2) & (hi2 - row.w) & H;
        const uint32_t r = (m0 >> 15) | (m1 >> 13) | (m2 >> 11) | (m3 >> 9);      /* slot 2i at bit 2i, slot 2i + 1 at bit 16 + 2i */
        const uint32_t mk = (r | (r >> 15)) & 0xFFu;
        sm.M[IDX(v, SDB_MAX_VALS)] = (uint8_t)mk;                                       /* empty slots hold T_EMPTY and never qualify */
        if (!mk) {
            const uint4 *kr = reinterpret_cast<const uint4 *>(A.tab.kill + (size_t)v * SDB_KILL_WORDS);
            const uint4 k0 = __ldg(&kr[0]);
            ka.x |= k0.x; ka.y |= k0.y; ka.z |= k0.z; ka.w |= k0.w;
            if (!MS) { const uint4 k1 = __ldg(&kr[1]); kb.x |= k1.x; kb.y |= k1.y; kb.z |= k1.z; kb.w |= k1.w; }
        }
    }
    {
        const uint32_t d0 = __reduce_or_sync(FULL, ka.x), d1 = __reduce_or_sync(FULL, ka.y), d2 = __reduce_or_sync(FULL, ka.z),
                       d3 = __reduce_or_sync(FULL, ka.w);
        uint32_t d4 = 0, d5 = 0, d6 = 0, d7 = 0;
        if (!MS) { d4 = __reduce_or_sync(FULL, kb.x); d5 = __reduce_or_sync(FULL, kb.y); d6 = __reduce_or_sync(FULL, kb.z); d7 = __reduce_or_sync(FULL, kb.w); }
        if (lane == 0) {
            sm.dead[0] = d0; sm.dead[1] = d1; sm.dead[2] = d2; sm.dead[3] = d3;
            sm.dead[4] = d4; sm.dead[5] = d5; sm.dead[6] = d6; sm.dead[7] = d7;
        }
    }
    __syncwarp();
    return true;
}

/* A warp's next block of a compact arena (survivor or match records): `blk` records, or what is left of the arena; false when
 * that is less than `need`.  Out of line on purpose: it runs once per ~15 messages, and the resolve kernel's hot code sits at
 * the edge of the instruction cache (a staging + copy epilogue of 250 instructions cost the kernel 18 %). */
__device__ __noinline__ bool claim_block(uint32_t *ctr, uint32_t cap, uint32_t blk, uint32_t need)
{
    WarpSm &sm = SM();
    const uint32_t want = blk > need ? blk : need;
    uint32_t o = 0;
    if (lane_id() == 0) o = atomicAdd(ctr, want);
    o = __shfl_sync(FULL, o, 0);
    const uint32_t left = (o <= cap && want <= cap - o) ? want : (o < cap ? cap - o : 0u);     /* the arena's tail: what is left of it, if anything */
    __syncwarp();
    if (lane_id() == 0) { sm.blk_off = o; sm.blk_left = left; }
    __syncwarp();
    return left >= need;
}
#ifndef SDB_PULSE_LONG
/* no room in the compact arena: the message goes to the overflow pass (worst-case slots); SDB_SURV_SHORT when that is full too */
__device__ __noinline__ uint32_t overflow_push(uint32_t *ovf_cnt, uint32_t *ovf_list, uint32_t ovf_max, uint32_t mi)
{
    uint32_t b = 0;
    if (lane_id() == 0) {
        b = atomicAdd(ovf_cnt, 1u);
        if (b < ovf_max) ovf_list[b] = mi;                            /* the overflow pass writes this message's surv_meta */
    }
    b = __shfl_sync(FULL, b, 0);
    return b < ovf_max ? 0u : SDB_SURV_SHORT;
}
#endif

/* OVF = false: the launch group's messages by ticket; survivors go straight into the warp's current block of the compact arena
 * (claimed SDB_SURV_BLOCK records at a time; a block is abandoned when fewer records are left than protocols passed the
 * prefilter).  When the arena is exhausted the message is listed for the overflow pass.
 * OVF = true (fast build only): the overflow pass — list-driven, survivors written into worst-case slots. */
template <bool MS, bool OVF>
__global__ void __launch_bounds__(KTHREADS, KRESOLVE_CTAS) resolve_kernel(KArgs A)
{
    WarpSm &sm = SM();
    const int lane = lane_id();
    const uint32_t nrows = MS ? A.tab.n_ms : A.tab.n_mu;
    const SdbPulseProto *rows = MS ? A.tab.ms : A.tab.mu;

    /* list-driven launches (overflow pass, long-message kernels) usually find an empty list: leave before the table copies */
    if ((OVF || KLIST_ALWAYS) && *A.list_cnt == 0) return;
    /* once per CTA: the pairs and the hot protocol fields into shared memory */
    {
        SdbValRow *sv = reinterpret_cast<SdbValRow *>(g_dyn);
        for (uint32_t v = threadIdx.x; v < A.tab.n_vals; v += blockDim.x) {
            SdbValRow r = A.tab.vals[v];                             /* bounds biased like T (prepare_tables) */
            r.lo = (int16_t)(r.lo + T_BIAS); r.hi = (int16_t)((r.hi + T_BIAS) | 0x8000);
            sv[v] = r;
        }
        HotRow *hrw = const_cast<HotRow *>(HOT_ROWS(A.tab.n_vals));
        for (uint32_t r = threadIdx.x; r < nrows; r += blockDim.x) {
            const SdbPulseProto *pr = &rows[r];
            HotRow h;
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
                h.key[kk].len = pr->key[kk].len; h.key[kk].nuniq = pr->key[kk].nuniq; h.key[kk].uidx = pr->key[kk].uidx;
                h.key[kk].vidx[0] = pr->key[kk].vidx[0]; h.key[kk].vidx[1] = pr->key[kk].vidx[1]; h.key[kk].rsv = 0;
            }
            h.width = pr->width; h.clk_idx = (uint8_t)pr->clk_idx; h.regex_min = pr->regex_min;
            hrw[r] = h;
        }
        double *hc = const_cast<double *>(HOT_CLK(A.tab.n_vals, nrows));
        if (!MS) {
            for (uint32_t c = threadIdx.x; c < 2 * A.tab.n_clk; c += blockDim.x) hc[c] = A.tab.clk[c];
        } else {
            for (uint32_t r = threadIdx.x; r < nrows; r += blockDim.x) hc[r] = rows[r].clock;     /* for the 30 % clock gate of pass 1 */
        }
        __syncthreads();
    }
    const HotRow *hot = HOT_ROWS(A.tab.n_vals);

    uint32_t tk_base = 0, tk_left = 0, mi = 0, lpos = 0;
    if (lane == 0) { sm.blk_off = 0; sm.blk_left = 0; }              /* this warp's block of the compact survivor arena */
    __syncwarp();
    while (next_message<OVF>(A, tk_base, tk_left, mi, lpos)) {
        const SdbPulseMsg *m = &A.msgs[mi];
        const int dlen = m->dlen;
        uint32_t nsurv = 0;
        uint32_t off = OVF ? A.ovf_base + lpos * A.surv_stride : 0u, cnt = 0;
        /* records outside the packed domain (ids > 9, > 8 slots, D too long) yield no hits instead of undefined lookups */
        const bool ids_ok = m->npat <= SDB_MAX_SLOTS &&
                            !__any_sync(FULL, lane < m->npat && ((m->pat_ids >> (4 * (lane & 7))) & 0xF) > 9);
        const bool decodable = (m->flags & SDB_MSG_VALID) && !(m->flags & SDB_MSG_DOMAIN) && dlen > 0 && dlen <= SDB_MAX_DIGITS && ids_ok;
#ifndef SDB_PULSE_LONG
        if (decodable && dlen > KMAXD) {               /* too long for this kernel's staging: the long kernels take it */
            if (lane == 0) A.long_list[atomicAdd(SDB_CTL(A, SDB_CTL_LONG_CNT), 1u)] = mi;
        } else
#endif
        if (decodable) {
            stage_message<!MS>(A, sm, m, dlen, mi);
            double clock_abs = 0.0;
            if (prepare_tables<MS>(A, sm, m, clock_abs)) {
                /* pass 1 (lane = protocol): keep the protocols whose mandatory values all have a candidate slot
                 * (pattern_utils.py:78-80; MS also the 30 % clock gate, message_synced.py:83-88); compact their
                 * row numbers, in table order, into plist */
                uint32_t nalive = 0;
#pragma unroll 1
                for (uint32_t q0 = 0; q0 < nrows; q0 += 32) {
                    const uint32_t q = q0 + lane;
                    bool ok = false;
                    if (q < nrows) {
                        ok = !((sm.dead[IDX(q >> 5, SDB_KILL_WORDS)] >> (q & 31)) & 1);
                        if (MS && ok) {
                            const double pclk = HOT_CLK(A.tab.n_vals, nrows)[q];
                            ok = !(pclk > 0.0 && fabs(__dsub_rn(pclk, clock_abs)) > __dmul_rn(clock_abs, 0.3));
                        }
                    }
                    const uint32_t bal = __ballot_sync(FULL, ok);
                    if (ok) sm.plist[IDX(nalive + __popc(bal & ((1u << lane) - 1)), 256)] = (uint8_t)q;
                    nalive += __popc(bal);
                }
                __syncwarp();
                if (!OVF) {
                    if (sm.blk_left < nalive && !claim_block(SDB_CTL(A, SDB_CTL_SURV), A.surv_cap, SDB_SURV_BLOCK, nalive)) {
#ifndef SDB_PULSE_LONG
                        cnt = overflow_push(SDB_CTL(A, SDB_CTL_OVF_CNT), A.ovf_list, A.ovf_max, mi);
#else
                        cnt = SDB_SURV_SHORT;                          /* the host grows the arena and runs the call again */
#endif
                        nalive = 0;
                    }
                    off = sm.blk_off;
                }
                SdbSurv *const slots = A.surv + off;
                /* pass 2 (lane = surviving protocol): exact template resolution */
#pragma unroll 1
                for (uint32_t q0 = 0; q0 < nalive; q0 += 32) {
                    const bool have = q0 + lane < nalive;
                    const uint32_t q = have ? sm.plist[IDX(q0 + lane, 256)] : 0;
                    int state = have ? 4 : 0;                         /* 4 = to be resolved by this lane */
                    SdbSurv rec;
                    rec.start = 0; rec.c1 = rec.c0 = rec.cf = 0; rec.meta = 0;
                    uint64_t long_start = 0;
                    int s0w = 0;
#pragma unroll 1
                    for (int pass = 0; pass < 2; pass++) {
                        if (state == 4) {
                            const SdbPulseProto *pq = &rows[q];
                            uint32_t codes = 0, sf = 0;
                            state = MS ? thread_resolve_ms(hot[q], pq, sm, codes, sf) : thread_resolve_mu(hot[q], pq, sm, codes, sf, pass == 1, s0w);
                            rec.start = pass == 1 ? long_start : (uint64_t)(codes & 0xFF);
                            rec.c1 = (codes >> 8) & 0xFF; rec.c0 = (codes >> 16) & 0xFF; rec.cf = codes >> 24;
                            rec.meta = (uint16_t)((sf & SURV_POS_MASK) | ((sf >> 16) ? SURV_HASF : 0));
                        }
                        if (pass == 1) break;
                        /* the few protocols that need warp-wide searches are resolved one after the other: an MU protocol with
                         * 2-digit symbols only needs its long start found by the warp, its lane does the rest in pass 1 */
                        uint32_t cx = __ballot_sync(FULL, state == 2);
                        while (cx) {
                            const int b = __ffs(cx) - 1;
                            cx &= cx - 1;
                            const uint32_t qb = __shfl_sync(FULL, q, b);
                            const SdbPulseProto *pb = &rows[qb];
                            if (!MS && pb->width == 2) {
                                uint64_t st = 0;
                                int sp = 0;
                                const int t_slot = T_GET(sm.T[IDX(pb->clk_idx, SDB_MAX_CLK)][lane & 7]);
                                const bool ok = resolve_key(&pb->key[0], t_slot, 0, true, st, sp);     /* :67-88 */
                                if (lane == b) { state = ok ? 4 : 0; long_start = st; s0w = sp; }
                            } else {
                                SdbSurv r2;
                                const bool ok = MS ? resolve_ms_warp(&rows[qb], r2) : resolve_mu_warp(&rows[qb], r2);
                                if (lane == b) { state = ok ? 1 : 0; rec = r2; }
                            }
                        }
                        if (!__any_sync(FULL, state == 4)) break;
                    }
                    const uint32_t alive = __ballot_sync(FULL, state == 1);
                    if (state == 1) {                                 /* protocol-table order is the slot order */
                        rec.start |= (uint64_t)q << 56;
                        slots[nsurv + __popc(alive & ((1u << lane) - 1))] = rec;
                    }
                    nsurv += __popc(alive);
                }
                if (!OVF && nsurv) {
                    __syncwarp();
                    if (lane == 0) { sm.blk_off += nsurv; sm.blk_left -= nsurv; }
                }
                if (nalive) cnt = nsurv;
            }
        }
        if (lane == 0) A.surv_meta[mi] = make_uint2(off, cnt);
        __syncwarp();
    }
}

/* the fused fallback kernel has two call sites (second pass: writing in place): always out of line */
__device__ __noinline__ int scan_survivors_mu_fused(const KArgs &A, const SdbSurv *slots, uint32_t nsurv)
{
    uint32_t unused = 0;
    return scan_survivors_mu_impl<false>(A, slots, nsurv, nullptr, unused);
}

template <bool MS>
__device__ __noinline__ int scan_survivors(const KArgs &A, const SdbSurv *slots, uint32_t nsurv)
{
    const int lane = lane_id();
    const SdbPulseProto *rows = MS ? A.tab.ms : A.tab.mu;
#pragma unroll 1
    for (uint32_t s0i = 0; s0i < nsurv; s0i += 32) {
        SdbSurv mine;
        mine.start = 0; mine.c1 = mine.c0 = mine.cf = 0; mine.meta = 0;
        if (s0i + lane < nsurv) *reinterpret_cast<uint4 *>(&mine) = LD_STREAM(reinterpret_cast<const uint4 *>(&slots[s0i + lane]));   /* coalesced 16-byte loads */
        const int cnt = min(32u, nsurv - s0i);
#pragma unroll 1
        for (int k = 0; k < cnt; k++) {
            const uint32_t slo = __shfl_sync(FULL, (uint32_t)mine.start, k);
            const uint32_t shi = __shfl_sync(FULL, (uint32_t)(mine.start >> 32), k);
            const uint32_t c10 = __shfl_sync(FULL, (uint32_t)mine.c1 | ((uint32_t)mine.c0 << 16), k);
            const uint32_t cfm = __shfl_sync(FULL, (uint32_t)mine.cf | ((uint32_t)mine.meta << 16), k);
            const SdbPulseProto *pp = &rows[shi >> 24];
            const uint64_t start_t = (((uint64_t)(shi & 0x00FFFFFFu)) << 32) | slo;
            const int pos0 = (int)((cfm >> 16) & SURV_POS_MASK);
            const bool hasf = ((cfm >> 16) & SURV_HASF) != 0;
            const int st = scan_ms(A, pp, pos0, (uint32_t)start_t, c10 & 0xFFFF, c10 >> 16, cfm & 0xFFFF, hasf);
            if (st != SDB_ST_OK) return st;
        }
    }
    return SDB_ST_OK;
}

template <bool MS>
__global__ void __launch_bounds__(KTHREADS, KMIN_CTAS) scan_kernel(KArgs A)
{
    WarpSm &sm = SM();
    const int lane = lane_id();

    uint32_t tk_base = 0, tk_left = 0, mi = 0, lpos = 0;
    while (next_message<false>(A, tk_base, tk_left, mi, lpos)) {
        const SdbPulseMsg *m = &A.msgs[mi];
        SdbMsgOut mo;
        mo.hit_off = 0; mo.nhits = 0; mo.status = SDB_ST_OK; mo.reason = 0;
#ifndef SDB_PULSE_LONG
        if (!MS && A.match_meta[mi].y != MU_MARK) continue;      /* MU: only what the match kernel could not hand over */
#endif
        const uint2 sv = A.surv_meta[mi];
        uint32_t nsurv = sv.y;
        if (nsurv == SDB_SURV_SHORT) {                           /* not resolved: the scratch was too small for this launch group */
            nsurv = 0;
            mo.status = SDB_ST_SCRATCH;
            if (lane == 0) atomicAdd(SDB_CTL(A, SDB_CTL_SHORT), 1u);
        } else if (MS && nsurv == 0 && msg_domain(m)) {   /* (a record outside the domain never has survivors) */
            mo.status = SDB_ST_DOMAIN;
            if (lane == 0) atomicAdd(&A.ctr->domain, 1u);
        }
        if (nsurv) {
            const SdbSurv *slots = A.surv + sv.x;
            stage_digits(A, sm, m, m->dlen, mi);
            int status = MS ? scan_survivors<MS>(A, slots, nsurv) : scan_survivors_mu_fused(A, slots, nsurv);
            __syncwarp();
            const uint32_t nh = sm.nh, nw = sm.nw;
            if (status != SDB_ST_OK) {
                mo.status = (uint8_t)status;                     /* exception: earlier hits are lost */
                if (lane == 0) atomicAdd(&A.ctr->raised, 1u);
            } else if (nh) {
                uint32_t hb = 0, wb = 0;
                if (lane == 0) {
                    hb = atomicAdd(&A.ctr->hits, nh);
                    wb = atomicAdd(&A.ctr->words, nw);
                }
                hb = __shfl_sync(FULL, hb, 0);
                wb = __shfl_sync(FULL, wb, 0);
                mo.hit_off = hb; mo.nhits = (uint16_t)nh;
                if (!sm.overflow) {
                    if (hb + nh <= A.hits_cap && wb + nw <= A.bits_cap) {
                        for (uint32_t i = lane; i < nh; i += 32) {
                            SdbHit h = sm.st_hits[IDX(i, ST_HITS)];
                            h.bits_off += wb;
                            A.hits[hb + i] = h;
                        }
                        for (uint32_t i = lane; i < nw; i += 32) A.bits[wb + i] = sm.st_bits[IDX(i, ST_WORDS)];
                    }
                } else {
                    /* rare: more output than the staging area holds -> scan again, writing in place */
                    __syncwarp();
                    if (lane == 0) { sm.nh = 0; sm.nw = 0; sm.direct = 1; sm.overflow = 0; sm.hbase = hb; sm.wbase = wb; }
                    __syncwarp();
                    if (MS) scan_survivors<MS>(A, slots, nsurv); else scan_survivors_mu_fused(A, slots, nsurv);
                }
            }
        }
        if (lane == 0) A.out[mi] = mo;
        __syncwarp();
    }
}

#ifndef SDB_PULSE_LONG
/* MU, kernel 2 of 3: every survivor's regex matches -> match records (message_unsynced.py:146-217), written into the warp's
 * current block of the compact match arena; when that arena is exhausted the message is left to the fused fallback kernel like
 * one with more than MU_MCAP matches. */
__global__ void __launch_bounds__(KTHREADS, KMATCH_CTAS) mu_match_kernel(KArgs A)
{
    WarpSm &sm = SM();
    const int lane = lane_id();

    uint32_t tk_base = 0, tk_left = 0, mi = 0, lpos = 0;
    if (lane == 0) { sm.blk_off = 0; sm.blk_left = 0; }              /* this warp's block of the compact match arena */
    __syncwarp();
    while (next_message<false>(A, tk_base, tk_left, mi, lpos)) {
        const SdbPulseMsg *m = &A.msgs[mi];
        const uint2 sv = A.surv_meta[mi];
        const bool scratch_short = sv.y == SDB_SURV_SHORT;
        const uint32_t nsurv = scratch_short ? 0u : sv.y;
        uint32_t nrec = 0, moff = 0;
        int status = SDB_ST_OK;
        if (nsurv) {
            stage_digits(A, sm, m, m->dlen, mi);
            if (sm.blk_left < MU_MCAP && !claim_block(SDB_CTL(A, SDB_CTL_MATCH), A.match_cap, SDB_MATCH_BLOCK, MU_MCAP))
                status = SDB_ST_MU_OVERFLOW;                          /* match arena exhausted: the fused fallback kernel takes the message */
            else {
                moff = sm.blk_off;
                status = scan_survivors_mu_impl<true>(A, A.surv + sv.x, nsurv, A.match + moff, nrec);
                __syncwarp();
                if (status == SDB_ST_OK && lane == 0) { sm.blk_off = moff + nrec; sm.blk_left -= nrec; }
            }
            __syncwarp();
        }
        if (lane == 0) {
            if (status == SDB_ST_MU_OVERFLOW) A.match_meta[mi] = make_uint2(0u, MU_MARK);
            else {
                const bool raised = status != SDB_ST_OK;
                A.match_meta[mi] = make_uint2(moff, raised ? 0u : nrec);
                if (raised || nrec == 0) {                       /* nothing left to do for the emit kernel */
                    SdbMsgOut mo;
                    mo.hit_off = 0; mo.nhits = 0; mo.status = (uint8_t)status; mo.reason = 0;
                    if (scratch_short) { mo.status = SDB_ST_SCRATCH; atomicAdd(SDB_CTL(A, SDB_CTL_SHORT), 1u); }
                    else if (nsurv == 0 && msg_domain(m)) { mo.status = SDB_ST_DOMAIN; atomicAdd(&A.ctr->domain, 1u); }
                    A.out[mi] = mo;
                    if (raised) atomicAdd(&A.ctr->raised, 1u);   /* exception: earlier hits are lost */
                }
            }
        }
        __syncwarp();
    }
}

/* the records of one message -> bits, post-demodulation, modulematch, staged hits */
FN_ONE_SITE void mu_emit_records_impl(const KArgs &A, const SdbSurv *slots, const uint32_t *recs, uint32_t nrec)
{
    const int lane = lane_id();
    int prev = -1, ordinal = 0;
#pragma unroll 1
    for (uint32_t r0 = 0; r0 < nrec; r0 += 32) {
        const bool have = r0 + lane < nrec;
        const uint32_t rec = have ? LD_STREAM(&recs[r0 + lane]) : 0u; /* coalesced */
        SdbSurv sv;
        sv.start = 0; sv.c1 = sv.c0 = sv.cf = 0; sv.meta = 0;
        if (have) *reinterpret_cast<uint4 *>(&sv) = LD_STREAM(reinterpret_cast<const uint4 *>(&slots[rec >> 24]));
        const int cnt = min(32u, nrec - r0);
#pragma unroll 1
        for (int k = 0; k < cnt; k++) {
            const uint32_t rk = __shfl_sync(FULL, rec, k);
            const uint32_t row = __shfl_sync(FULL, (uint32_t)(sv.start >> 56), k);
            const uint32_t c10 = __shfl_sync(FULL, (uint32_t)sv.c1 | ((uint32_t)sv.c0 << 16), k);
            const uint32_t cfm = __shfl_sync(FULL, (uint32_t)sv.cf | ((uint32_t)sv.meta << 16), k);
            const SdbPulseProto *pp = &A.tab.mu[row];
            const bool has0 = pp->key[2].len != 0, hasf = ((cfm >> 16) & SURV_HASF) != 0;
            const uint32_t c1 = c10 & 0xFFFF, c0 = has0 ? c10 >> 16 : c1, cf = hasf ? cfm & 0xFFFF : c1;
            const int si = (int)(rk >> 24);
            ordinal = si == prev ? ordinal + 1 : 0;                   /* records of one survivor are consecutive, in match order */
            prev = si;
            mu_emit_match(A, pp, rk & 0x00FFFFFFu, c1, c0, cf, hasf, ordinal);
        }
    }
}

__device__ __noinline__ void mu_emit_records_again(const KArgs &A, const SdbSurv *slots, const uint32_t *recs, uint32_t nrec)
{
    mu_emit_records_impl(A, slots, recs, nrec);              /* the rare second pass, writing in place */
}

/* MU, kernel 3 of 3: match records -> hits (message_unsynced.py:220-290) */
__global__ void __launch_bounds__(KTHREADS, KEMIT_CTAS) mu_emit_kernel(KArgs A)
{
    WarpSm &sm = SM();
    const int lane = lane_id();

    uint32_t tk_base = 0, tk_left = 0, mi = 0, lpos = 0;
    while (next_message<false>(A, tk_base, tk_left, mi, lpos)) {
        const uint2 mm = A.match_meta[mi];
        const uint32_t nrec = mm.y;
        if (nrec == 0 || nrec == MU_MARK) continue;
        const SdbPulseMsg *m = &A.msgs[mi];
        const SdbSurv *slots = A.surv + A.surv_meta[mi].x;
        const uint32_t *recs = A.match + mm.x;
        SdbMsgOut mo;
        mo.hit_off = 0; mo.nhits = 0; mo.status = SDB_ST_OK; mo.reason = 0;
        stage_digits(A, sm, m, m->dlen, mi);
        mu_emit_records_impl(A, slots, recs, nrec);
        __syncwarp();
        const uint32_t nh = sm.nh, nw = sm.nw;
        if (nh) {
            uint32_t hb = 0, wb = 0;
            if (lane == 0) {
                hb = atomicAdd(&A.ctr->hits, nh);
                wb = atomicAdd(&A.ctr->words, nw);
            }
            hb = __shfl_sync(FULL, hb, 0);
            wb = __shfl_sync(FULL, wb, 0);
            mo.hit_off = hb; mo.nhits = (uint16_t)nh;
            if (!sm.overflow) {
                if (hb + nh <= A.hits_cap && wb + nw <= A.bits_cap) {
                    for (uint32_t i = lane; i < nh; i += 32) {
                        SdbHit h = sm.st_hits[IDX(i, ST_HITS)];
                        h.bits_off += wb;
                        A.hits[hb + i] = h;
                    }
                    for (uint32_t i = lane; i < nw; i += 32) A.bits[wb + i] = sm.st_bits[IDX(i, ST_WORDS)];
                }
            } else {
                /* rare: more output than the staging area holds -> emit again, writing in place */
                __syncwarp();
                if (lane == 0) { sm.nh = 0; sm.nw = 0; sm.direct = 1; sm.overflow = 0; sm.hbase = hb; sm.wbase = wb; }
                __syncwarp();
                mu_emit_records_again(A, slots, recs, nrec);
            }
        }
        if (lane == 0) A.out[mi] = mo;
        __syncwarp();
    }
}

/* dynamic shared memory of the resolve kernels (the CTA-shared table copies); opting in is needed beyond 48 KB in total.
 * The attribute is per FUNCTION (not per handle): it is set to the largest size any table can need, so that handles with
 * different tables can coexist (a handle created later with a smaller table must not lower the cap of an earlier one). */
template <bool MS>
static size_t resolve_dyn_smem(const SdbDevTable &tab)
{
    cudaFuncSetAttribute(resolve_kernel<MS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hot_bytes(SDB_MAX_VALS, 255));
    cudaFuncSetAttribute(resolve_kernel<MS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hot_bytes(SDB_MAX_VALS, 255));
    return hot_bytes(tab.n_vals, MS ? tab.n_ms : tab.n_mu);
}

/* resident CTAs per SM of the resolve, match / scan and emit kernels (each launch is a persistent grid of its own size) */
void pulse_blocks_per_sm(int kind, const SdbDevTable &tab, int per_sm[3])
{
    int a = 0, b = 0, c = 0;
    if (kind == SDB_KIND_MS) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, resolve_kernel<true, false>, SDB_PULSE_THREADS, resolve_dyn_smem<true>(tab));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, scan_kernel<true>, SDB_PULSE_THREADS, 0);
        c = b;
    } else {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, resolve_kernel<false, false>, SDB_PULSE_THREADS, resolve_dyn_smem<false>(tab));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, mu_match_kernel, SDB_PULSE_THREADS, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c, mu_emit_kernel, SDB_PULSE_THREADS, 0);
    }
    per_sm[0] = a > 0 ? a : 1; per_sm[1] = b > 0 ? b : 1; per_sm[2] = c > 0 ? c : 1;
}

/* ---- scratch layout (one block per handle; SdbScratchCfg in sdb_pulse.h) ---- */
struct ScratchLayout {
    size_t surv, surv_meta, match, match_meta, ctl, stats, long_list, ovf_list, total;
    uint32_t surv_cap, match_cap;
};
static ScratchLayout scratch_layout(uint32_t stride, const SdbScratchCfg &c)
{
    ScratchLayout L;
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const uint64_t scap = (uint64_t)c.chunk * c.surv_avg + (uint64_t)c.warps * SDB_SURV_BLOCK,
                   mcap = (uint64_t)c.chunk * c.match_avg + (uint64_t)c.warps * SDB_MATCH_BLOCK;
    L.surv_cap = scap > 0xF0000000ull ? 0xF0000000u : (uint32_t)scap;
    L.match_cap = mcap > 0xF0000000ull ? 0xF0000000u : (uint32_t)mcap;
    size_t o = 0;
    L.surv = o;       o = up(o + ((size_t)L.surv_cap + (size_t)c.ovf_max * stride) * sizeof(SdbSurv));
    L.surv_meta = o;  o = up(o + (size_t)c.chunk * sizeof(uint2));
    L.match = o;      o = up(o + (size_t)L.match_cap * sizeof(uint32_t));
    L.match_meta = o; o = up(o + (size_t)c.chunk * sizeof(uint2));
    L.ctl = o;        o = up(o + SDB_CTL_WORDS * sizeof(uint32_t));
    L.stats = o;      o = up(o + SDB_STAT_WORDS * sizeof(uint32_t));
    L.long_list = o;  o = up(o + (size_t)c.chunk * sizeof(uint32_t));
    L.ovf_list = o;   o = up(o + (size_t)c.ovf_max * sizeof(uint32_t));
    L.total = o;
    return L;
}
size_t pulse_scratch_bytes(uint32_t stride, const SdbScratchCfg &cfg) { return scratch_layout(stride, cfg).total; }
size_t pulse_scratch_stats_offset(uint32_t stride, const SdbScratchCfg &cfg) { return scratch_layout(stride, cfg).stats; }
size_t pulse_scratch_ctl_offset(uint32_t stride, const SdbScratchCfg &cfg) { return scratch_layout(stride, cfg).ctl; }
void pulse_scratch_caps(uint32_t stride, const SdbScratchCfg &cfg, uint32_t caps[2])
{
    const ScratchLayout L = scratch_layout(stride, cfg);
    caps[0] = L.surv_cap; caps[1] = L.match_cap;
}

/* between launch groups: fold the group's allocation counters into the persistent statistics (what the host sizes the
 * scratch from) and zero the work / allocation counters for the next group */
__global__ void fold_ctl_kernel(uint32_t *ctl, uint32_t *stats)
{
    const int t = threadIdx.x;
    if (t == 0) {
        stats[SDB_STAT_SURV] = max(stats[SDB_STAT_SURV], ctl[SDB_CTL_SURV * SDB_CTL_STRIDE]);
        stats[SDB_STAT_MATCH] = max(stats[SDB_STAT_MATCH], ctl[SDB_CTL_MATCH * SDB_CTL_STRIDE]);
        stats[SDB_STAT_OVF] = max(stats[SDB_STAT_OVF], ctl[SDB_CTL_OVF_CNT * SDB_CTL_STRIDE]);
        stats[SDB_STAT_SHORT] += ctl[SDB_CTL_SHORT * SDB_CTL_STRIDE];
    }
    __syncwarp();
    if (t < SDB_CTL_COUNTERS) ctl[t * SDB_CTL_STRIDE] = 0;
}

/* unit op: one postDemo_* call on one bit list (bytes 0/1), executed by the device function above */
#define UNIT_WORDS (SDB_MAX_DIGITS / 32 + 8)
__global__ void unit_postdemod_kernel(int method, const uint8_t *in, uint32_t n_in, uint8_t *out, uint32_t out_cap, int32_t *res)
{
    __shared__ uint32_t a[UNIT_WORDS], b[UNIT_WORDS];
    if (threadIdx.x == 0) {
        for (int i = 0; i < UNIT_WORDS; i++) { a[i] = 0; b[i] = 0; }
        for (uint32_t i = 0; i < n_in; i++) sbit(a, (int)i, in[i] & 1);
        int no = 0;
        int rc = postdemod(method, a, (int)n_in, b, &no);
        res[0] = rc; res[1] = no;
        for (int i = 0; i < no && (uint32_t)i < out_cap; i++) out[i] = (uint8_t)gbit(b, i);
    }
}
int launch_unit_postdemod(int method, const uint8_t *d_in, uint32_t n_in, uint8_t *d_out, uint32_t out_cap,
                          int32_t *d_res, cudaStream_t stream)
{
    if (n_in > (uint32_t)SDB_MAX_DIGITS) return -1;
    unit_postdemod_kernel<<<1, 32, 0, stream>>>(method, d_in, n_in, d_out, out_cap, d_res);
    return (int)cudaGetLastError();
}

unsigned int debug_violations(bool reset)
{
#ifdef SDB_BOUNDS_CHECK
    unsigned int v = 0, z = 0;
    cudaMemcpyFromSymbol(&v, g_sdb_oob, sizeof v);
    if (reset) cudaMemcpyToSymbol(g_sdb_oob, &z, sizeof z);
    return v + sdb_long::debug_violations_long(reset);
#else
    (void)reset;
    return 0xFFFFFFFFu;      /* not a checked build */
#endif
}

int launch_pulse(int kind, const SdbDevTable &tab, const SdbPulseMsg *d_msgs, const uint8_t *d_digits, uint32_t n,
                 SdbMsgOut *d_out, SdbHit *d_hits, uint32_t hits_cap, uint32_t *d_bits, uint32_t bits_cap,
                 SdbCounters *d_ctr, const int grid[3], int grid_long, void *scratch, const SdbScratchCfg &cfg, uint32_t msg_base0, cudaStream_t stream)
{
    if (n == 0) return 0;
    if (!scratch || !cfg.chunk) return (int)cudaErrorInvalidValue;
    const bool ms = kind == SDB_KIND_MS;
    const uint32_t stride = tab.n_ms > tab.n_mu ? tab.n_ms : tab.n_mu;       /* scratch is sized for the larger class */
    const ScratchLayout L = scratch_layout(stride, cfg);
    uint8_t *sb = static_cast<uint8_t *>(scratch);
    KArgs A;
    A.tab = tab; A.digits = d_digits; A.hits = d_hits; A.hits_cap = hits_cap; A.bits = d_bits; A.bits_cap = bits_cap; A.ctr = d_ctr;
    A.surv = reinterpret_cast<SdbSurv *>(sb + L.surv); A.surv_cap = L.surv_cap; A.ovf_base = L.surv_cap; A.ovf_max = cfg.ovf_max;
    A.surv_stride = ms ? tab.n_ms : tab.n_mu;
    A.surv_meta = reinterpret_cast<uint2 *>(sb + L.surv_meta);
    A.match = reinterpret_cast<uint32_t *>(sb + L.match); A.match_cap = L.match_cap;
    A.match_meta = reinterpret_cast<uint2 *>(sb + L.match_meta);
    A.ctl = reinterpret_cast<uint32_t *>(sb + L.ctl);
    uint32_t *stats = reinterpret_cast<uint32_t *>(sb + L.stats);
    A.long_list = reinterpret_cast<uint32_t *>(sb + L.long_list);
    A.ovf_list = reinterpret_cast<uint32_t *>(sb + L.ovf_list);
    A.list = nullptr; A.list_cnt = nullptr; A.list_max = 0;
    const uint32_t wpc = SDB_PULSE_THREADS / 32;
    const size_t dyn = hot_bytes(tab.n_vals, ms ? tab.n_ms : tab.n_mu);
    for (uint32_t off = 0; off < n; off += cfg.chunk) {
        A.msgs = d_msgs + off; A.out = d_out + off; A.msg_base = msg_base0 + off;
        A.n = n - off < cfg.chunk ? n - off : cfg.chunk;
        const uint32_t need = (A.n + wpc - 1) / wpc;
        int g[3];
        for (int i = 0; i < 3; i++) g[i] = need < (uint32_t)grid[i] ? (int)need : grid[i];
        A.ticket_batch = ms ? 4 * SDB_TICKET_BATCH : SDB_TICKET_BATCH;   /* MS messages are ~10x cheaper; the fallback kernel skips nearly everything: 256 */
        fold_ctl_kernel<<<1, 32, 0, stream>>>(A.ctl, stats);
        A.list = nullptr; A.list_cnt = nullptr; A.list_max = 0;
        A.ticket = A.ctl + SDB_CTL_STRIDE * 0;
        if (ms) resolve_kernel<true, false><<<g[0], SDB_PULSE_THREADS, dyn, stream>>>(A);
        else resolve_kernel<false, false><<<g[0], SDB_PULSE_THREADS, dyn, stream>>>(A);
        /* overflow pass: the messages whose survivors did not fit the compact arena (usually none: the grid exits at once) */
        A.list = A.ovf_list; A.list_cnt = SDB_CTL(A, SDB_CTL_OVF_CNT); A.list_max = cfg.ovf_max;
        A.ticket = A.ctl + SDB_CTL_STRIDE * 6;
        {
            const uint32_t need_o = (cfg.ovf_max + wpc - 1) / wpc;
            const int go = need_o < (uint32_t)g[0] ? (int)need_o : g[0];
            if (ms) resolve_kernel<true, true><<<go, SDB_PULSE_THREADS, dyn, stream>>>(A);
            else resolve_kernel<false, true><<<go, SDB_PULSE_THREADS, dyn, stream>>>(A);
        }
        A.list = nullptr; A.list_cnt = nullptr; A.list_max = 0;
        if (ms) {
            A.ticket = A.ctl + SDB_CTL_STRIDE * 1; scan_kernel<true><<<g[1], SDB_PULSE_THREADS, 0, stream>>>(A);
        } else {
            A.ticket = A.ctl + SDB_CTL_STRIDE * 1; mu_match_kernel<<<g[1], SDB_PULSE_THREADS, 0, stream>>>(A);
            A.ticket = A.ctl + SDB_CTL_STRIDE * 2; mu_emit_kernel<<<g[2], SDB_PULSE_THREADS, 0, stream>>>(A);
            A.ticket = A.ctl + SDB_CTL_STRIDE * 3; A.ticket_batch = 256;
            scan_kernel<false><<<g[0] < g[1] ? g[0] : g[1], SDB_PULSE_THREADS, 0, stream>>>(A);     /* fused fallback: messages with > MU_MCAP matches only */
        }
        /* messages with more than SDB_FAST_DIGITS digits (listed by the resolve kernel above; usually none) */
        A.ticket_batch = 1;
        A.list = A.long_list; A.list_cnt = SDB_CTL(A, SDB_CTL_LONG_CNT); A.list_max = A.n;
        cudaError_t e = (cudaError_t)sdb_long::launch_long(kind, A, grid_long, stream);
        if (e != cudaSuccess) return (int)e;
    }
    fold_ctl_kernel<<<1, 32, 0, stream>>>(A.ctl, stats);
    return (int)cudaGetLastError();
}

#else  /* SDB_PULSE_LONG */

unsigned int debug_violations_long(bool reset)
{
#ifdef SDB_BOUNDS_CHECK
    unsigned int v = 0, z = 0;
    cudaMemcpyFromSymbol(&v, g_sdb_oob, sizeof v);
    if (reset) cudaMemcpyToSymbol(g_sdb_oob, &z, sizeof z);
    return v;
#else
    (void)reset;
    return 0u;
#endif
}

int long_blocks_per_sm(const SdbDevTable &tab)
{
    int a = 0, b = 0, c = 0, d = 0;
    const size_t dms = hot_bytes(tab.n_vals, tab.n_ms), dmu = hot_bytes(tab.n_vals, tab.n_mu);
    cudaFuncSetAttribute(resolve_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hot_bytes(SDB_MAX_VALS, 255));
    cudaFuncSetAttribute(resolve_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hot_bytes(SDB_MAX_VALS, 255));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, resolve_kernel<true, false>, KTHREADS, dms);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, scan_kernel<true>, KTHREADS, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c, resolve_kernel<false, false>, KTHREADS, dmu);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d, scan_kernel<false>, KTHREADS, 0);
    int nb = a;
    if (b < nb) nb = b;
    if (c < nb) nb = c;
    if (d < nb) nb = d;
    return nb > 0 ? nb : 1;
}

/* unit op: pattern_exists (pattern_utils.py:34-136) for ONE template on ONE digit string — the warp-level resolver the
 * kernels use (resolve_key / resolve_general / warp_find), behind the module-level pattern_utils.pattern_exists of the
 * drop-in package.  res = {found, target digits 0..7, target digits 8..15, position of the first occurrence}. */
__global__ void __launch_bounds__(KTHREADS, KMIN_CTAS) unit_pattern_kernel(SdbKeyTpl tpl, const uint16_t *rank, const int16_t *tenths,
                                                                            uint32_t pat_ids, int npat, const uint8_t *digits, int dlen,
                                                                            int32_t *res)
{
    if (threadIdx.x >= 32) return;
    WarpSm &sm = SM();
    const int lane = lane_id();
    KArgs A;
    A.digits = digits; A.tab.rank = rank; A.msg_base = 0;
    __shared__ SdbPulseMsg m;
    if (lane == 0) { m.doff = 0; m.dlen = (uint16_t)dlen; m.npat = (uint8_t)npat; m.cp = 0xFF; m.pat_ids = pat_ids; m.flags = SDB_MSG_VALID; }
    if (lane < 8) m.pat[lane] = 0;
    __syncwarp();
    stage_message<true>(A, sm, &m, dlen, 0);
    const int t_slot = tenths[lane & 7];
    uint64_t tgt = 0;
    int pos = 0;
    const SdbKeyTpl k = tpl;
    const bool ok = resolve_key(&k, t_slot, 0, true, tgt, pos);
    if (lane == 0) { res[0] = ok ? 1 : 0; res[1] = (int32_t)(uint32_t)tgt; res[2] = (int32_t)(uint32_t)(tgt >> 32); res[3] = pos; }
}

int launch_unit_pattern(const SdbKeyTpl &tpl, const uint16_t *d_rank, const int16_t *d_tenths, uint32_t pat_ids, int npat,
                        const uint8_t *d_digits, int dlen, int32_t *d_res, cudaStream_t stream)
{
    if (dlen < 0 || dlen > SDB_MAX_DIGITS || npat < 0 || npat > SDB_MAX_SLOTS || tpl.len > SDB_MAX_TPL || tpl.nuniq > SDB_MAX_UNIQ) return -1;
    unit_pattern_kernel<<<1, 32, 0, stream>>>(tpl, d_rank, d_tenths, pat_ids, npat, d_digits, dlen, d_res);
    return (int)cudaGetLastError();
}

/* The messages the fast resolve kernel listed (SDB_FAST_DIGITS < dlen <= SDB_MAX_DIGITS): resolve + fused scan, sized for
 * the long staging buffers.  A0 = the fast launch group's own arguments with `list` set to the long list; survivors are
 * claimed from the same compact arena, after the fast kernels' (which are done with theirs: stream order). */
int launch_long(int kind, const SdbPulseArgs &A0, int grid, cudaStream_t stream)
{
    KArgs A = A0;
    A.ticket_batch = 1;
    if (kind == SDB_KIND_MS) {
        A.ticket = A.ctl + SDB_CTL_STRIDE * 4; resolve_kernel<true, false><<<grid, KTHREADS, hot_bytes(A.tab.n_vals, A.tab.n_ms), stream>>>(A);
        A.ticket = A.ctl + SDB_CTL_STRIDE * 5; scan_kernel<true><<<grid, KTHREADS, 0, stream>>>(A);
    } else {
        A.ticket = A.ctl + SDB_CTL_STRIDE * 4; resolve_kernel<false, false><<<grid, KTHREADS, hot_bytes(A.tab.n_vals, A.tab.n_mu), stream>>>(A);
        A.ticket = A.ctl + SDB_CTL_STRIDE * 5; scan_kernel<false><<<grid, KTHREADS, 0, stream>>>(A);
    }
    return (int)cudaGetLastError();
}
#endif /* SDB_PULSE_LONG */

}  // namespace KNS
