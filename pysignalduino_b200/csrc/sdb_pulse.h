/* sdb_pulse.h — host-callable launchers of the device kernels (internal to libsdb200.so). */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/sdb200.h"
#include "sdb_table.h"

#define SDB_PULSE_THREADS 256   /* 8 warps = 8 messages in flight per CTA */
#ifndef SDB_PULSE_MIN_CTAS
#define SDB_PULSE_MIN_CTAS 4  /* register cap = 65536 / (256 * MIN_CTAS) */
#endif
#define SDB_HEX_THREADS   128
#define SDB_FAST_DIGITS   1024  /* digits per message the fast MS / MU kernels stage; longer ones (<= SDB_MAX_DIGITS) go to the long kernels */
#define SDB_LONG_THREADS  64    /* long kernels: 2 warps per CTA (17 KB of staging per warp) */

/* Scratch of one handle: what the kernels of a launch group hand each other.  Survivor and match records
 * are COMPACT: a message claims exactly as many records as it produced from a per-group arena (one atomicAdd per message),
 * so the arenas are sized by the AVERAGE number of records per message (`surv_avg`, `match_avg`) and not by the worst case
 * (every protocol of every message: 129 x 16 B per message).  A message whose survivors do not fit — arena full, or more of
 * them than a warp can stage — is listed and resolved again by a second launch that writes into worst-case slots
 * (`ovf_max` messages); beyond that it is flagged SDB_ST_SCRATCH and the host grows the budgets from the recorded need. */
struct SdbScratchCfg {
    uint32_t chunk;        /* messages per launch group                              */
    uint32_t surv_avg;     /* survivor records per message the compact arena holds   */
    uint32_t match_avg;    /* MU match records per message the compact arena holds   */
    uint32_t ovf_max;      /* messages the worst-case overflow region holds          */
    uint32_t warps;        /* warps of the largest persistent grid (one block of each arena per warp on top of the budgets) */
};
#define SDB_SURV_AVG_DEFAULT 19u
#define SDB_MATCH_AVG_DEFAULT 8u
#define SDB_OVF_MAX_DEFAULT 4096u
/* persistent statistics of a scratch block (device words, read back by the host paths) */
#define SDB_STAT_SURV 0    /* largest number of survivor records any launch group asked for        */
#define SDB_STAT_MATCH 1   /* ... of match records                                                 */
#define SDB_STAT_OVF 2     /* ... of messages sent to the overflow pass                            */
#define SDB_STAT_SHORT 3   /* messages flagged SDB_ST_SCRATCH since the statistics were last reset */
#define SDB_STAT_WORDS 8

/* everything a pulse kernel launch needs (the KArgs of both builds of sdb_pulse.cu) */
/* One resolved (message x protocol) task handed from the resolve kernel to the scan / match / emit kernels, 16 bytes. */
struct __align__(16) SdbSurv {
    uint64_t start;        /* bits 0..55: id string of `start` / `sync` (nibble-packed), bits 56..63: table row */
    uint16_t c1, c0, cf;   /* id strings of one / zero / float (<= 4 digits) */
    uint16_t meta;         /* SURV_POS_MASK: s0 (where D' begins), SURV_HASF: float resolved */
};
struct SdbPulseArgs {
    SdbDevTable tab;
    const SdbPulseMsg *msgs;   /* already offset to the first message of this launch */
    const uint8_t *digits;
    uint32_t n;                /* messages of this launch */
    uint32_t msg_base;         /* batch index of msgs[0] (hit.msg is a batch index) */
    SdbMsgOut *out;            /* already offset */
    SdbHit *hits;  uint32_t hits_cap;
    uint32_t *bits; uint32_t bits_cap;
    SdbCounters *ctr;
    SdbSurv *surv;             /* survivor arena: surv_cap compact records, then ovf_max x surv_stride worst-case slots */
    uint32_t surv_cap, ovf_base, ovf_max;
    uint32_t surv_stride;      /* protocols of this class (47 MS / 129 MU) */
    uint2 *surv_meta;          /* per message: {first record, count}; count SDB_SURV_SHORT = not resolved (scratch too small) */
    uint32_t *match;           /* MU match arena (match kernel -> emit kernel), match_cap records */
    uint32_t match_cap;
    uint2 *match_meta;         /* per message: {first record, count}; count MU_MARK = left to the fused fallback kernel */
    uint32_t *ctl;             /* this launch group's work and allocation counters (SDB_CTL_*), zeroed per group */
    uint32_t *ticket;          /* = &ctl[...]: warps draw messages ticket_batch at a time */
    uint32_t ticket_batch;
    uint32_t *long_list;       /* messages with SDB_FAST_DIGITS < dlen <= SDB_MAX_DIGITS (fast resolve -> long kernels) */
    uint32_t *ovf_list;        /* messages whose survivors did not fit the compact arena (fast resolve -> overflow pass) */
    const uint32_t *list;      /* list-driven launches: the list to draw from, its length and the most entries to take */
    const uint32_t *list_cnt;
    uint32_t list_max;
};
/* every counter sits on a 128-byte line of its own (the tickets are drawn once per message by every warp of the grid) */
#define SDB_CTL_STRIDE 32
#define SDB_CTL_TICKET0 0      /* [0..3] fast kernels, [4..5] long kernels, [6] overflow pass */
#define SDB_CTL_LONG_CNT 7
#define SDB_CTL_OVF_CNT 8
#define SDB_CTL_SURV 9         /* survivor records claimed (may exceed surv_cap: the need) */
#define SDB_CTL_MATCH 10
#define SDB_CTL_SHORT 11
#define SDB_CTL_COUNTERS 12
#define SDB_CTL_WORDS (SDB_CTL_COUNTERS * SDB_CTL_STRIDE)
#define SDB_CTL(A, i) ((A).ctl + (i) * SDB_CTL_STRIDE)
/* records a warp claims per atomicAdd: it writes message after message into its block and abandons the block when fewer
 * records are left than the next message could need (protocols that passed the prefilter / MU_MCAP).  The arenas hold one
 * block per resident warp on top of the per-message budgets (the blocks in use when a kernel ends). */
#define SDB_SURV_BLOCK 256u    /* >= protocols per class (<= 255) */
#define SDB_MATCH_BLOCK 512u   /* >= MU_MCAP */
#define SDB_SURV_SHORT 0xFFFFFFFFu

namespace sdb {

int launch_pulse(int kind, const SdbDevTable &tab, const SdbPulseMsg *d_msgs, const uint8_t *d_digits, uint32_t n,
                 SdbMsgOut *d_out, SdbHit *d_hits, uint32_t hits_cap, uint32_t *d_bits, uint32_t bits_cap,
                 SdbCounters *d_ctr, const int grid[3], int grid_long, void *scratch, const SdbScratchCfg &cfg, uint32_t msg_base0, cudaStream_t stream);
void pulse_blocks_per_sm(int kind, const SdbDevTable &tab, int per_sm[3]);   /* resolve, match / scan, emit */
unsigned int debug_violations(bool reset);   /* bounds-check build only; 0xFFFFFFFF otherwise */
size_t pulse_scratch_bytes(uint32_t stride, const SdbScratchCfg &cfg);
size_t pulse_scratch_stats_offset(uint32_t stride, const SdbScratchCfg &cfg);   /* SDB_STAT_WORDS words */
size_t pulse_scratch_ctl_offset(uint32_t stride, const SdbScratchCfg &cfg);     /* SDB_CTL_WORDS words */
void pulse_scratch_caps(uint32_t stride, const SdbScratchCfg &cfg, uint32_t caps[2]);   /* records of the survivor / match arena */

#ifndef SDB_TICKET_BATCH
#define SDB_TICKET_BATCH 1     /* messages a warp draws per atomicAdd on the launch's work counter */
#endif
#ifndef SDB_MU_CHUNK
#define SDB_MU_CHUNK 1048576u
#endif
/* SDB_MU_CHUNK: messages per launch group of the device-resident calls (bounds the survivor scratch: chunk * n_mu * 16 B) */
#ifndef SDB_PIPE_CHUNK
#define SDB_PIPE_CHUNK 262144u
#endif
/* SDB_PIPE_CHUNK: messages per pipeline stage of the host-buffer calls (H2D / kernels / D2H overlap) */

int launch_hex(int kind, int mc_repaired, const SdbDevTable &tab, const SdbHexMsg *d_msgs, const uint8_t *d_digits,
               uint32_t n, SdbMsgOut *d_out, SdbHit *d_hits, uint32_t hits_cap, uint32_t *d_bits, uint32_t bits_cap,
               SdbCounters *d_ctr, int grid, cudaStream_t stream);

size_t lines_pool_bytes(size_t text_len, uint32_t n);
int launch_tokenize(int kind, const uint8_t *d_text, const uint32_t *d_off, const uint32_t *d_len, uint32_t n, uint32_t base,
                    SdbPulseMsg *d_msgs, uint8_t *d_pool, SdbLineInfo *d_info, uint32_t *d_long /* 2 + n words */, int sm_count, cudaStream_t stream);

/* sdb_format.cu: payload strings of the MS / MU hits [range[0], ctr->hits) into a device pool (NUL-terminated, one SdbPayloadHit per hit) */
int launch_format(int kind, const SdbHit *d_hits, const uint32_t *d_bits, const SdbPulseProto *rows, const uint16_t *row_of_proto,
                  const SdbHexProto *hx, uint32_t nproto,
                  uint32_t *d_range, const SdbCounters *d_ctr, uint32_t hits_cap, uint32_t bits_cap, char *d_pool, uint32_t pool_cap,
                  SdbPayloadHit *d_phits, uint32_t *d_used, int grid, cudaStream_t stream);

int launch_unit_mc(const SdbDevTable &tab, uint32_t proto, int method_override, const uint8_t *d_bits, int n, int mcbitnum,
                   uint8_t *d_out, int out_cap, int32_t *d_seg, int32_t *d_res, cudaStream_t stream);

int launch_unit_postdemod(int method, const uint8_t *d_in, uint32_t n_in, uint8_t *d_out, uint32_t out_cap,
                          int32_t *d_res /* [rc, n_out] */, cudaStream_t stream);

}  // namespace sdb

/* sdb_pulse_long.cu: the same kernels sized for SDB_FAST_DIGITS < D <= SDB_MAX_DIGITS */
namespace sdb_long {
int long_blocks_per_sm(const SdbDevTable &tab);
int launch_unit_pattern(const SdbKeyTpl &tpl, const uint16_t *d_rank, const int16_t *d_tenths, uint32_t pat_ids, int npat,
                        const uint8_t *d_digits, int dlen, int32_t *d_res, cudaStream_t stream);
unsigned int debug_violations_long(bool reset);
int launch_long(int kind, const SdbPulseArgs &A0, int grid, cudaStream_t stream);   /* A0: the fast launch group's own arguments */
}
