/*
 * sdb_table.h — binary layout of the compiled protocol table (host + device).
 * Mirrors the numpy dtypes in pysignalduino_b200/table.py field for field.
 */
#ifndef SDB_TABLE_H
#define SDB_TABLE_H

#include <stdint.h>

#define SDB_TBL_MAGIC   0x31424453u   /* "SDB1" */
#define SDB_TBL_VERSION 10u

#define SDB_MAX_UNIQ 4     /* distinct values per template (shipped table: <= 4)   */
#define SDB_MAX_TPL  14    /* template length (longest `start` has 14 pulses)      */
#define SDB_MAX_REQ  12    /* prefilter intervals per protocol                     */
#define SDB_MAX_CLK  64    /* distinct MU clockabs values (shipped table: 55)      */

/* SdbPulseProto.flags */
#define SDB_PF_RECONSTRUCT  0x01
#define SDB_PF_DISPATCH_BIN 0x02
#define SDB_PF_REMOVE_ZERO  0x04
#define SDB_PF_MM_END       0x08
#define SDB_PF_MM_NEVER     0x10
#define SDB_PF_HAS_LIR_MAX  0x20
#define SDB_PF_MM_HOST      0x40   /* modulematch outside the device program's shapes: hits carry SDB_HIT_MM_HOST */

/* postDemodulation ids (sd_protocols/postdemodulation.py) */
#define SDB_PD_EM 1
#define SDB_PD_REVOLT 2
#define SDB_PD_FS20 3
#define SDB_PD_FHT80 4
#define SDB_PD_FHT80TF 5
#define SDB_PD_WS2000 6
#define SDB_PD_WS7035 7
#define SDB_PD_WS7053 8
#define SDB_PD_LENGTHPREFIX 9

/* MC / MN method ids (sd_protocols/manchester.py, helpers.py) */
#define SDB_M_NONE 0
#define SDB_M_FUNKBUS 1
#define SDB_M_SAINLOGIC 2
#define SDB_M_AS 3
#define SDB_M_HIDEKI 4
#define SDB_M_MAVERICK 5
#define SDB_M_OSV1 6
#define SDB_M_OSV2O3 7
#define SDB_M_OSPIR 8
#define SDB_M_MCRAW_MANCHESTER 9
#define SDB_M_MCRAW_HELPERS 10
#define SDB_M_TFA 11
#define SDB_M_GROTHE 12
#define SDB_M_SOMFY 13
#define SDB_M_BRESSER_LIGHTNING 14
#define SDB_M_BRESSER_5IN1 15
#define SDB_M_BRESSER_6IN1 16
#define SDB_M_BRESSER_7IN1 17
#define SDB_M_PCA301 18
#define SDB_M_KOPP 19
#define SDB_M_LACROSSE 20
#define SDB_M_UNKNOWN 21

/* SdbHexProto.flags */
#define SDB_HF_EXISTS     0x01
#define SDB_HF_HAS_MIN    0x02
#define SDB_HF_HAS_MAX    0x04
#define SDB_HF_CLOCKRANGE 0x08
#define SDB_HF_INVERT     0x10
#define SDB_HF_MAX_IS_STR 0x20
#define SDB_HF_IS_119     0x40

/* One pulse template (sync / start / one / zero / float), 48 bytes. */
typedef struct SdbKeyTpl {
    uint8_t  len;                     /* pulses in the template, 0 = key absent                     */
    uint8_t  nuniq;                   /* distinct values, first-appearance order                    */
    uint16_t rsv;
    uint32_t uidx;                    /* 2 bits per position: which distinct value                  */
    int16_t  lo[SDB_MAX_UNIQ];        /* accepted tenths interval of each distinct value            */
    int16_t  hi[SDB_MAX_UNIQ];
    uint32_t rank_off[SDB_MAX_UNIQ];  /* gap-rank table slice: rank[rank_off + (t - lo)]            */
    uint16_t vidx[SDB_MAX_UNIQ];      /* MU: row of (clock, interval) in SdbValRow[] = slot of the per-message candidate mask */
} SdbKeyTpl;

/* One distinct (clock, accept interval) pair of the MU table: the resolve kernel computes, once per message,
 * the 8-bit mask of pattern slots whose tenths value lies inside it. */
typedef struct SdbValRow {
    uint16_t clk_idx;
    int16_t  lo, hi;
    uint16_t rsv;
} SdbValRow;
#define SDB_MAX_VALS 512
#define SDB_KILL_WORDS 8

/* One MS or MU protocol, 248 bytes. key[0] = sync (MS) / start (MU), [1] one, [2] zero, [3] float. */
typedef struct SdbPulseProto {
    SdbKeyTpl key[4];
    double   clock;                   /* MS: clockabs for the 30 % gate (0 = no gate); MU: clockabs */
    uint16_t proto;                   /* index in protocol-table order                              */
    uint16_t clk_idx;                 /* MU: index into the distinct-clock list                     */
    int16_t  regex_min;               /* MU: {MIN,}; MS: int(length_min) or -1                      */
    int16_t  lir_min;                 /* length_in_range minimum, -1 = none                         */
    int16_t  lir_max;                 /* length_in_range maximum (valid with SDB_PF_HAS_LIR_MAX)    */
    int16_t  mu_len_max;              /* MU: len(chunks) limit, -1 = none                           */
    uint8_t  width;                   /* digits per symbol                                          */
    uint8_t  padbits;
    uint8_t  postdemod;               /* SDB_PD_*                                                   */
    uint8_t  flags;                   /* SDB_PF_*                                                   */
    uint8_t  pre_len, post_len;
    uint16_t mm_off;                  /* first modulematch item, 0xFFFF = none                      */
    char     preamble[16];
    char     postamble[4];
    uint8_t  mm_nitems;
    uint8_t  rsv[7];
} SdbPulseProto;

/* Prefilter row, 28 bytes: rows of the candidate-mask table that must all be non-zero.  The kernels use the transposed
 * form (kill masks per pair, off_kill); these rows stay in the blob as its source and for inspection. */
typedef struct SdbPrefilter {
    uint16_t clk_idx;
    uint16_t nreq;
    uint16_t vreq[SDB_MAX_REQ];
} SdbPrefilter;

/* One modulematch atom: `min` (..`max` for the last, '$'-anchored atom) characters inside `mask`. */
typedef struct SdbMmItem {
    uint32_t mask[4];                 /* 128-bit ASCII membership                                   */
    uint16_t min, max;
} SdbMmItem;

/* One protocol as seen by the MC / MN paths, 36 bytes, one row per protocol id (table order). */
typedef struct SdbHexProto {
    int32_t length_min, length_max;
    int32_t clock_min, clock_max;
    uint8_t method;                   /* SDB_M_*  */
    uint8_t flags;                    /* SDB_HF_* */
    uint8_t pre_len;
    uint8_t pid_int;                  /* int(protocol id): 1 = positive, 2 = zero or negative, 0 = int() raises ValueError
                                         (only read by the as-shipped MC path for method manchester.mcRaw, whose shifted
                                         arguments turn the id into mcbitnum, manchester.py:120 / :606-611) */
    char    preamble[16];             /* MC payload prefix (manchester.py:131-132), host formatting only */
} SdbHexProto;

typedef struct SdbTblHeader {
    uint32_t magic, version, nproto;
    uint32_t n_ms, n_mu, n_clk, n_rank, n_mm;
    uint32_t off_ms, off_mu, off_ms_pf, off_mu_pf, off_clk, off_rank, off_mm, off_hex;
    uint32_t total;
    uint32_t n_vals, off_vals;   /* SdbValRow[]: MU pairs first, then the MS intervals (clock slot 0) */
    uint32_t n_mu_vals;
    uint32_t off_kill;           /* uint32[n_vals][SDB_KILL_WORDS]: protocol rows (of the pair's class) that need pair v */
    uint32_t rsv[3];
} SdbTblHeader;

#ifdef __cplusplus
static_assert(sizeof(SdbKeyTpl) == 48, "SdbKeyTpl layout");
static_assert(sizeof(SdbValRow) == 8, "SdbValRow layout");
static_assert(sizeof(SdbPulseProto) == 248, "SdbPulseProto layout");
static_assert(sizeof(SdbPrefilter) == 28, "SdbPrefilter layout");
static_assert(sizeof(SdbMmItem) == 20, "SdbMmItem layout");
static_assert(sizeof(SdbHexProto) == 36, "SdbHexProto layout");
static_assert(sizeof(SdbTblHeader) == 96, "SdbTblHeader layout");
#endif

/* Device-side view of the table (pointers into the device copy of the blob). */
typedef struct SdbDevTable {
    const SdbPulseProto *ms;    const SdbPrefilter *ms_pf;   uint32_t n_ms;
    const SdbPulseProto *mu;    const SdbPrefilter *mu_pf;   uint32_t n_mu;
    const double        *clk;   uint32_t n_clk;     /* clk[0..n_clk) clocks, clk[n_clk..2n_clk) = 10/clock */
    const uint16_t      *rank;
    const SdbValRow     *vals;  uint32_t n_vals, n_mu_vals;
    const uint32_t      *kill;
    const SdbMmItem     *mm;
    const SdbHexProto   *hex;   uint32_t nproto;
} SdbDevTable;

#endif /* SDB_TABLE_H */
