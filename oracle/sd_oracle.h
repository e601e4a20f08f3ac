/*
 * sd_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C, string-based restatement of PySignalduino's SDProtocols.demodulate()
 * hot path.  It is deliberately literal: D= stays a C string, pattern_exists walks the
 * cartesian product, every tolerance is evaluated in float64 at run time — i.e. it shares
 * no tables, no precomputed intervals and no code with the CUDA path it checks.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (pysignalduino_b200/) never does.
 *
 * Parity pin: oracle/validate_vs_reference.py runs this oracle against the imported
 * Python reference on fuzz corpora; tests/golden/ holds reference-generated fixtures.
 */
#ifndef SD_ORACLE_H
#define SD_ORACLE_H

#include <stdint.h>
#include "../include/sdb200.h"   /* only the packed-batch structs of the boundary */

#ifdef __cplusplus
extern "C" {
#endif

#define ORA_MAXLIST 16

/* postDemodulation method ids (sd_protocols/postdemodulation.py) */
enum {
    ORA_PD_NONE = 0, ORA_PD_EM, ORA_PD_REVOLT, ORA_PD_FS20, ORA_PD_FHT80, ORA_PD_FHT80TF,
    ORA_PD_WS2000, ORA_PD_WS7035, ORA_PD_WS7053, ORA_PD_LENGTHPREFIX
};

/* MC / MN method ids (sd_protocols/manchester.py, helpers.py) */
enum {
    ORA_M_NONE = 0,
    ORA_M_FUNKBUS, ORA_M_SAINLOGIC, ORA_M_AS, ORA_M_HIDEKI, ORA_M_MAVERICK, ORA_M_OSV1,
    ORA_M_OSV2O3, ORA_M_OSPIR, ORA_M_MCRAW_MANCHESTER, ORA_M_MCRAW_HELPERS, ORA_M_TFA,
    ORA_M_GROTHE, ORA_M_SOMFY,
    ORA_M_BRESSER_LIGHTNING, ORA_M_BRESSER_5IN1, ORA_M_BRESSER_6IN1, ORA_M_BRESSER_7IN1,
    ORA_M_PCA301, ORA_M_KOPP, ORA_M_LACROSSE,
    ORA_M_UNKNOWN            /* a method name that does not exist on the class */
};

/* One protocol, filled literally from the protocol dict by oracle/oracle.py. */
typedef struct OraProto {
    char    id[16];
    int32_t has_clockabs;      double clockabs;          /* float(clockabs) */
    int32_t sync_kind;         /* 0 absent/falsy, 1 every element float()-able, 2 float() raises */
    int32_t nsync;             double sync[ORA_MAXLIST];
    int32_t start_is_list;     /* truthy and isinstance(list) (message_unsynced.py:71) */
    int32_t nstart;            double start[ORA_MAXLIST];
    int32_t none;              double one[ORA_MAXLIST];  /* n = -1: float() raises, 0: absent/falsy */
    int32_t nzero;             double zero[ORA_MAXLIST];
    int32_t nfloat;            double flt[ORA_MAXLIST];
    int32_t has_length_min;    int32_t length_min;
    int32_t has_length_max;    int32_t length_max;
    int32_t length_max_truthy; /* MU: `if length_max and ...` (message_unsynced.py:217) */
    int32_t paddingbits;       /* int(check_property(pid,'paddingbits',4)) */
    int32_t postdemod;         /* ORA_PD_* of an EXISTING method, else 0 (hasattr gate) */
    int32_t reconstruct;       /* truthiness of reconstructBit */
    int32_t dispatch_bin;      /* int(dispatchBin) == 1 */
    int32_t remove_zero;       /* truthiness */
    int32_t active;            /* check_property(pid,'active',True) truthiness */
    int32_t has_modulematch;   /* truthiness */
    char    modulematch[96];
    char    preamble[32];
    char    postamble[16];
    /* MC / MN */
    int32_t method;            /* ORA_M_* */
    int32_t has_clockrange;    int32_t clock_min, clock_max;
    int32_t polarity_invert;
    int32_t length_max_is_str; /* helpers.mcraw compares int > str -> TypeError (helpers.py:113-114) */
} OraProto;

typedef struct OraHit {
    int32_t msg;
    int32_t proto;             /* index into the OraProto array */
    int32_t bit_length;
    int32_t payload_off;       /* into the char pool */
    int32_t payload_len;
} OraHit;

/*
 * Demodulate n MS or MU messages (kind = SDB_KIND_MS / SDB_KIND_MU) on `nthreads` host threads.
 * status[i] = SDB_ST_*; hits are ordered by (msg, protocol order, match order).
 * Returns 0, or -3 if hits_cap / pool_cap are too small (needed sizes in *nhits, *pool_used).
 */
int ora_demod_pulse(const OraProto *tab, int nproto, int kind,
                    const SdbPulseMsg *msgs, const uint8_t *digits, int64_t n,
                    uint8_t *status, OraHit *hits, int64_t hits_cap,
                    char *pool, int64_t pool_cap, int64_t *nhits, int64_t *pool_used,
                    int nthreads);

/* MC / MN (kind = SDB_KIND_MC / SDB_KIND_MN); mc_repaired as in sdb200.h. */
int ora_demod_hex(const OraProto *tab, int nproto, int kind, int mc_repaired,
                  const SdbHexMsg *msgs, const uint8_t *digits, int64_t n,
                  uint8_t *status, OraHit *hits, int64_t hits_cap,
                  char *pool, int64_t pool_cap, int64_t *nhits, int64_t *pool_used,
                  int nthreads);

/* Scalar pieces, exported so tests can pin them on the reference's own unit vectors. */
double ora_round1(double x);                 /* CPython round(x, 1), fast path (FMA residual) */
double ora_round1_printf(double x);          /* same via correctly-rounded "%.1f" (slow, independent) */
double ora_tolerance(double v);              /* pattern_utils.calculate_tolerance */
int    ora_pattern_exists(const double *search, int ns, const char *ids, const double *vals, int npat,
                          const char *raw, char *out /* >= ns+1 */);
int    ora_postdemod(int method, const uint8_t *in, int n, uint8_t *out, int out_cap, int *n_out); /* rc 1 / 0; -2 = ValueError */
int    ora_bin2hex(const char *bits, char *out);   /* helpers.bin_str_2_hex_str; returns len or -1 for None */

#ifdef __cplusplus
}
#endif
#endif
