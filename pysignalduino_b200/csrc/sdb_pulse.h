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

namespace sdb {

int launch_pulse(int kind, const SdbDevTable &tab, const SdbPulseMsg *d_msgs, const uint8_t *d_digits, uint32_t n,
                 SdbMsgOut *d_out, SdbHit *d_hits, uint32_t hits_cap, uint32_t *d_bits, uint32_t bits_cap,
                 SdbCounters *d_ctr, int grid, int grid_long, void *mu_scratch, uint32_t mu_chunk, uint32_t msg_base0, cudaStream_t stream);
int pulse_blocks_per_sm(int kind, const SdbDevTable &tab);
unsigned int debug_violations(bool reset);   /* bounds-check build only; 0xFFFFFFFF otherwise */
size_t mu_scratch_bytes(uint32_t n_mu, uint32_t chunk);   /* survivor slots handed from mu_resolve_kernel to mu_scan_kernel */

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
int launch_long(int kind, const SdbDevTable &tab, const SdbPulseMsg *d_msgs, const uint8_t *d_digits, uint32_t n,
                uint32_t msg_base, SdbMsgOut *d_out, SdbHit *d_hits, uint32_t hits_cap, uint32_t *d_bits, uint32_t bits_cap,
                SdbCounters *d_ctr, void *surv, uint32_t *surv_cnt, uint32_t surv_stride, uint32_t *long_list, uint32_t *long_cnt,
                uint32_t *tickets, int grid, cudaStream_t stream);
}
