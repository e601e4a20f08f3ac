/*
 * sdb200.h — C ABI of libsdb200.so, the B200-native batch demodulator that sits
 * behind PySignalduino's SDProtocols.demodulate() hot path.
 *
 * Every entry point takes plain pointers and sizes (no torch / C++ types), so it
 * can be bound from ctypes (what pysignalduino_b200/capi.py does), cffi, or any
 * other FFI.  Each function cites the reference interface it replaces
 * (paths relative to the PySignalduino source tree).
 *
 * Threading: one handle may be used by one host thread at a time
 * (reference: engine called through asyncio.to_thread, signalduino/controller.py:252).
 */
#ifndef SDB200_H
#define SDB200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDB_ABI_VERSION 3      /* 2: SdbMsgOut.reason, SdbCounters.domain, SDB_ST_DOMAIN, payload / reserve / pattern entry points;
                                  3: SDB_ST_SCRATCH + sdb_scratch_short (compact scratch arenas) */

/* ---- limits of the packed domain ------------------------------------------------ */
#define SDB_MAX_SLOTS   8      /* P0..P7: pulse-pattern slots per message (firmware emits <= 8) */
#define SDB_MAX_DIGITS  4096   /* digits in one D= stream (fast kernels stage 1024, longer messages take the long kernels) */
#define SDB_DIGIT_OTHER 0xE    /* any character of D that is not '0'..'9'                      */
#define SDB_DIGIT_PAD   0xF    /* padding nibble after the last digit                          */
#define SDB_MAX_HEX     512    /* hex characters in one MC / MN D= field                       */

/* ---- return codes --------------------------------------------------------------- */
#define SDB_OK            0
#define SDB_E_ARG        -1    /* bad argument / malformed table blob          */
#define SDB_E_CUDA       -2    /* CUDA runtime error, see sdb_last_error()     */
#define SDB_E_OVERFLOW   -3    /* hit / bit arena too small: counters say how much is needed */
#define SDB_E_NOGPU      -4    /* no CUDA device: there is NO CPU fallback     */
#define SDB_E_SCRATCH    -5    /* messages still flagged SDB_ST_SCRATCH after the scratch was grown 6 times (a user table
                                  with more than 151 protocols in one class and messages longer than 1024 digits) */

/* ---- message kinds (SDProtocols.demodulate(msg, msg_type), sd_protocols/sd_protocols.py:60-74) */
#define SDB_KIND_MS 0          /* demodulate_ms, sd_protocols/message_synced.py:10-243   */
#define SDB_KIND_MU 1          /* demodulate_mu, sd_protocols/message_unsynced.py:11-296 */
#define SDB_KIND_MC 2          /* demodulate_mc, sd_protocols/sd_protocols.py:76-111 + manchester.py:49-144 */
#define SDB_KIND_MN 3          /* demodulate_mn, sd_protocols/sd_protocols.py:113-155    */

/* ---- per-message status (replaces Python exceptions that escape demodulate()) ---- */
#define SDB_ST_OK          0   /* returned a (possibly empty) list                                  */
#define SDB_ST_INDEXERROR  1   /* message_unsynced.py:212 chunks[-1] on an empty capture            */
#define SDB_ST_TYPEERROR   2   /* manchester.py:84 / :120, helpers.py:114 (MC as shipped; id 57)    */
#define SDB_ST_VALUEERROR  3   /* message_synced.py:174 range(..., 0) with signal_width == 0        */
#define SDB_ST_DOMAIN      4   /* NOT a reference outcome: the message is outside the packed domain (SDB_MSG_DOMAIN, or a
                                  malformed record); it was not decoded — no hits, reported per message, never silently wrong */
#define SDB_ST_SCRATCH     5   /* NOT a reference outcome, device-pointer calls only: the message was not decoded because the
                                  handle's scratch arenas (sized by AVERAGE survivor counts) were too small for this batch;
                                  sdb_scratch_short() grows them, then submit the flagged messages again.  The host-buffer
                                  calls do that themselves and never return this status. */

/*
 * One MS / MU message after host-side packing of the parser dict
 * (input contract: message_synced.py:21-66, message_unsynced.py:22-35).
 * 48 bytes, 16-byte aligned; algorithmic input bytes/message = 48 + ceil(dlen/2).
 */
typedef struct SdbPulseMsg {
    int32_t  pat[SDB_MAX_SLOTS]; /* pulse values, slot order = insertion order of the P<d> keys (pattern_utils.py:73,83 tie-break) */
    uint32_t doff;               /* start of this message's digit stream in the digit pool, in 16-byte units */
    uint16_t dlen;               /* number of digits                                           */
    uint8_t  npat;               /* valid slots 0..8                                           */
    uint8_t  cp;                 /* MS: slot index whose id == int(CP) (message_synced.py:59-66), 0xFF if none */
    uint32_t pat_ids;            /* nibble s = decimal id (0..9) that slot s is addressed by in D */
    uint8_t  flags;              /* SDB_MSG_* */
    uint8_t  rsv[3];
} SdbPulseMsg;

#define SDB_MSG_VALID 0x01       /* passed the host-side input gates; a message without it yields [] */
#define SDB_MSG_DOMAIN 0x04      /* the packer could not represent the message (non-integer pulse value, pattern id > 9,
                                    more than SDB_MAX_SLOTS slots, D longer than SDB_MAX_DIGITS / SDB_MAX_HEX, lower-case hex):
                                    status SDB_ST_DOMAIN, the rest of the batch is decoded normally */

/*
 * One MC / MN message (sd_protocols.py:79-88: protocol_id, data, clock, bit_length).
 * 16 bytes.
 */
typedef struct SdbHexMsg {
    uint32_t doff;               /* start of the hex nibbles in the digit pool, in 16-byte units */
    uint16_t hlen;               /* number of hex characters                                     */
    uint16_t proto;              /* index of protocol_id in protocol-table order, 0xFFFF = unknown/missing */
    int32_t  clock;              /* MC: C=                                                       */
    int16_t  bitlen;             /* MC: L= (clamped to int16)                                    */
    uint8_t  flags;              /* SDB_MSG_VALID | SDB_HEX_* */
    uint8_t  rsv;                /* MN: != 0 runs that converter instead of the protocol's own (direct Conv* call) */
} SdbHexMsg;

#define SDB_HEX_TOGGLE_POLARITY 0x02  /* messagetype 'Mc' or firmware "V 3.2." (manchester.py:94-96)          */

/* Per-message result slot, 8 bytes. hits of message m = hits[hit_off .. hit_off+nhits) in reference order. */
typedef struct SdbMsgOut {
    uint32_t hit_off;
    uint16_t nhits;
    uint8_t  status;             /* SDB_ST_* */
    uint8_t  reason;             /* MC: why _demodulate_mc_data rejected the message (SDB_MCR_*, 0 = none); 0 otherwise */
} SdbMsgOut;

/* MC reject reasons (manchester.py:70-128 and the mcBit2* decoders): the host renders the reference's message text */
#define SDB_MCR_NONE 0
#define SDB_MCR_TOO_SHORT 1        /* 'message is too short' */
#define SDB_MCR_TOO_LONG 2         /* 'message is too long' */
#define SDB_MCR_WRONG_BEGIN 3      /* 'wrong bits at begin' */
#define SDB_MCR_PARITY 4           /* 'parity error' */
#define SDB_MCR_CHECKSUM 5         /* 'checksum error' */
#define SDB_MCR_NO_START 6         /* '<name>: lib/mcBit2Sainlogic, start 010100 not found' */
#define SDB_MCR_NOT_32 7           /* 'message must be 32 bits, got <n>' */
#define SDB_MCR_NOT_56 8           /* 'message must be 56 bits, got <n>' */
#define SDB_MCR_NO_SYNC 9          /* 'sync not found' */
#define SDB_MCR_NO_DUP 10          /* ' no duplicate found'; sdb_unit_mc ORs SDB_MCR_DUP_*; SdbMsgOut.reason uses 12 / 13 / 14 */
#define SDB_MCR_LOOP 11            /* 'loop error, please report this data <bits>' */
#define SDB_MCR_CLOCK 20           /* 'clock out of range' (manchester.py:86) */
#define SDB_MCR_NO_METHOD 21       /* [(-1, 'Protocol method not defined', {})] (:108-109) */
#define SDB_MCR_UNKNOWN_METHOD 22  /* 'Unknown protocol method <name>' (:121-123) */
#define SDB_MCR_DUP_SHORT 0x100    /* ', message is too short' */
#define SDB_MCR_DUP_LONG 0x200     /* ', message is too long' */
#define SDB_MCR_DUP_NOPROTO 0x400  /* ', protocol does not exists' */

/* One decoded message ("hit"), 16 bytes.  payload bits live in the bit arena. */
typedef struct SdbHit {
    uint32_t msg;                /* index of the message in the batch                                   */
    uint32_t bits_off;           /* first 32-bit word of this hit in the bit arena                      */
    uint16_t proto;              /* index in protocol-table order (JSON order, sd_protocols.py:49-52)   */
    uint16_t nbits;              /* meta.bit_length (message_synced.py:237 / message_unsynced.py:286)   */
    uint16_t aux;                /* MS/MU: match ordinal within (msg, proto); MC/MN: decoder-specific   */
    uint8_t  flags;              /* SDB_HIT_* */
    uint8_t  rsv;
} SdbHit;

/* One hit as sdb_demod_host_payloads returns it, 12 bytes: what a caller needs next to the payload string (the message is
 * implied by SdbMsgOut.hit_off / nhits, the bits stay on the device). */
typedef struct SdbPayloadHit {
    uint32_t str_off;            /* the NUL-terminated payload string starts at pool[str_off]                 */
    uint16_t proto;              /* as SdbHit                                                                */
    uint16_t nbits;
    uint16_t aux;
    uint8_t  flags;
    uint8_t  rsv;
} SdbPayloadHit;

#define SDB_HIT_HAS_F   0x01     /* a second plane of nbits follows: 1 = symbol is 'F' (float)           */
#define SDB_HIT_LIST    0x02     /* MC TFA: element of the duplicate list (manchester.py:705-717)        */
#define SDB_HIT_FIELDS  0x04     /* MN: bits hold decoder fields, host renders "OK 9 ..." / "OK 24 ..."  */
#define SDB_HIT_MM_HOST 0x08     /* MU, user-edited tables only: the protocol's modulematch is a regex the device program cannot
                                    express; the caller must still apply re.search(modulematch, payload) (message_unsynced.py:277-280)
                                    — pysignalduino_b200's formatters do.  Never set with the shipped protocol table. */

/* Device-side counters written by every demod call (16 bytes). */
typedef struct SdbCounters {
    uint32_t hits;               /* hit records needed (may exceed capacity -> SDB_E_OVERFLOW)           */
    uint32_t words;              /* bit-arena words needed                                               */
    uint32_t raised;             /* messages for which the reference raises (status 1..3)                */
    uint32_t domain;             /* messages with status SDB_ST_DOMAIN                                   */
} SdbCounters;

typedef struct SdbHandle SdbHandle;

/* Library / device ------------------------------------------------------------- */
int         sdb_abi_version(void);
const char *sdb_last_error(const SdbHandle *h);   /* h may be NULL for creation errors */

/*
 * Create an engine on CUDA device `device` from a compiled protocol table blob
 * (pysignalduino_b200/table.py; replaces SDProtocols.__init__/_load_protocols,
 * sd_protocols/sd_protocols.py:25-41).  Fails with SDB_E_NOGPU when no device exists.
 */
int  sdb_create(const void *blob, size_t blob_len, int device, SdbHandle **out);
void sdb_destroy(SdbHandle *h);

/*
 * Device-resident batch demodulation of MS or MU messages: all pointers are DEVICE
 * pointers, the call only enqueues work on `stream` (a cudaStream_t passed as void*).
 * Replaces SDProtocols.demodulate(msg, "MS"|"MU") for n messages
 * (sd_protocols.py:64-71 -> message_synced.py:10 / message_unsynced.py:11).
 */
int sdb_demod_pulse_device(SdbHandle *h, int kind,
                           const SdbPulseMsg *d_msgs, const uint8_t *d_digits, uint32_t n,
                           SdbMsgOut *d_out, SdbHit *d_hits, uint32_t hits_cap,
                           uint32_t *d_bits, uint32_t bits_cap,
                           SdbCounters *d_counters, void *stream);

/*
 * Pre-size the handle's scratch (survivor / match records the kernels of one launch group hand each other) for launch
 * groups of up to n_messages.  Without it the first sdb_demod_pulse_device() call that needs more scratch than any call
 * before synchronises `stream`, frees and reallocates (it blocks the host and cannot be captured into a CUDA graph); after
 * it, device calls with n <= n_messages (or any n once 1 048 576 is reserved: larger batches run as several groups) only
 * enqueue.  All work on ONE handle must be stream-ordered: the scratch and the work counters belong to the handle, so two
 * calls in flight on different streams would share them — use one handle per stream.
 */
int sdb_reserve(SdbHandle *h, uint32_t n_messages);

/*
 * The scratch holds COMPACT survivor / match records: every warp claims blocks of 256 survivor records (16 B each) / 512 match
 * records (4 B each) from per-launch-group arenas and fills them message after message.  The arenas are sized for 19 survivor
 * and 8 match records per message on average plus one block per resident warp (the benchmark corpus writes 14.6 and 6.7 per
 * message and claims 18.4 and 8.6 with the block remainders), and worst-case slots for 4096 messages per launch group take
 * what did not fit — 0.42 GB per 1 048 576-message group instead of the 2.2 GB that room for every protocol of every message
 * took.  A launch group that needs more (more than 19 claimed records per message AND more than 4096 messages that did not
 * fit) leaves the excess messages undecoded with
 * status SDB_ST_SCRATCH.  The host-buffer calls notice, grow the scratch from the recorded need and repeat the call.  After
 * device-pointer calls, sdb_scratch_short() synchronises the device, stores in *n_short how many messages were flagged since
 * the last check and grows the budgets (the next call reallocates), so that submitting the flagged messages again succeeds.
 */
int sdb_scratch_short(SdbHandle *h, uint32_t *n_short);
/* Set the budgets (0 = keep the current one): survivor records / MU match records per message on average, messages of the
 * worst-case region; slack_warps = 0 (automatic) unless a test wants the arenas smaller than one block per resident warp on top
 * of the budgets.  Synchronises and releases the current block; the next call allocates with the new budgets. */
int sdb_scratch_budget(SdbHandle *h, uint32_t surv_avg, uint32_t match_avg, uint32_t ovf_max, uint32_t slack_warps);
/* Bytes of the scratch block as allocated (0 = none yet); cfg = {messages per launch group, surv_avg, match_avg, ovf_max, arena
 * blocks added on top (warps), and the largest values any sdb_scratch_short() / host-buffer call has read back: the
 * survivor-arena claim of a launch group (records), the match-arena claim, the overflow list}. */
size_t sdb_scratch_info(const SdbHandle *h, uint32_t cfg[8]);

/* Same for MC / MN (sd_protocols.py:76-155, manchester.py, helpers.py:223-716).
 * mc_repaired: 0 = as shipped (TypeError, SURVEY §8c "strict"), 1 = the two documented one-line repairs. */
int sdb_demod_hex_device(SdbHandle *h, int kind, int mc_repaired,
                         const SdbHexMsg *d_msgs, const uint8_t *d_digits, uint32_t n,
                         SdbMsgOut *d_out, SdbHit *d_hits, uint32_t hits_cap,
                         uint32_t *d_bits, uint32_t bits_cap,
                         SdbCounters *d_counters, void *stream);

/*
 * Host-buffer convenience call (what SDProtocols.demodulate()/demodulate_batch() use):
 * copies the packed batch to the device, runs the kernels, copies results back and
 * synchronises.  All pointers are HOST pointers.  `msgs` is SdbPulseMsg[] for MS/MU and
 * SdbHexMsg[] for MC/MN.  On SDB_E_OVERFLOW `counters` holds the required capacities.
 */
int sdb_demod_host(SdbHandle *h, int kind, int mc_repaired,
                   const void *msgs, const uint8_t *digits, size_t digits_len, uint32_t n,
                   SdbMsgOut *out, SdbHit *hits, uint32_t hits_cap,
                   uint32_t *bits, uint32_t bits_cap, SdbCounters *counters);

/*
 * sdb_demod_host plus the payload string of every hit in one call: what SDProtocols.demodulate() returns per hit is a
 * string (preamble + hex / bits + postamble, message_synced.py:224-231, message_unsynced.py:254-274, manchester.py:131-132).
 * A format kernel runs after the decode kernels of every pipeline stage and the strings travel back with the hits under the
 * next stage's kernels, as 12-byte SdbPayloadHit records (string offset, protocol, bit length, flags) + the string pool:
 * 27.5 bytes per hit on the wire instead of 16 (SdbHit) + 4 (offset) + 15.5 (string) — with N ranks on one host the host's
 * copy bandwidth is what bounds the end-to-end path.  `hits` and `bits` may be NULL (the 16-byte records / the bit arena are
 * then not copied back; hits_cap / bits_cap still size the device arenas).  phits has hits_cap entries; hit i's string
 * starts at pool[phits[i].str_off] and is NUL-terminated (the pool is NOT in hit order).  SDB_E_OVERFLOW when an arena or
 * the pool is too small (counters / *pool_used say how much is needed).
 */
int sdb_demod_host_payloads(SdbHandle *h, int kind, int mc_repaired,
                            const void *msgs, const uint8_t *digits, size_t digits_len, uint32_t n,
                            SdbMsgOut *out, SdbHit *hits, uint32_t hits_cap,
                            uint32_t *bits, uint32_t bits_cap, SdbCounters *counters,
                            char *pool, size_t pool_cap, SdbPayloadHit *phits, size_t *pool_used);

/*
 * Host-side result formatting (the payload string of every hit:
 * preamble + hex/bits + postamble, message_synced.py:224-231, message_unsynced.py:254-274,
 * manchester.py:131-132).  Writes NUL-free strings into `pool`; str_off has nhits+1 entries.
 * Returns SDB_E_OVERFLOW (and the needed size in *pool_used) when pool_cap is too small.
 */
int sdb_format_hits(const SdbHandle *h, int kind,
                    const SdbHit *hits, uint32_t nhits, const uint32_t *bits,
                    char *pool, size_t pool_cap, uint64_t *str_off, size_t *pool_used);

/*
 * Firmware text lines -> hits (signalduino/parser/ms.py:26-69, mu.py:26-82 without the
 * per-line Python work).  `text` holds payload lines of ONE message type (MS or MU), already stripped of the
 * STX/ETX framing (base.py:174-193); line i is text[line_off[i] .. +line_len[i]), offsets ascending, lines disjoint.
 * A tokenizer kernel replaces _parse_to_dict (ms.py:71-84), the MU validity regex (mu.py:48-52) and the input gates
 * of demodulate_ms / _mu (message_synced.py:21-66, message_unsynced.py:22-35); the demodulation kernels follow.
 * Results are indexed by line.  info[i].status: SDB_LINE_INVALID = the reference yields [] for the line,
 * SDB_LINE_OK = decoded here, SDB_LINE_HOSTPATH = outside the canonical grammar (non-ASCII, duplicate / multi-digit
 * pattern ids, values float() reads differently, D > SDB_MAX_DIGITS digits): the caller packs that line itself.
 * All pointers are HOST pointers; pass PINNED buffers (text, offsets, out, info, hits, bits): the call pipelines H2D /
 * kernels / D2H per stage, and a pageable buffer makes each asynchronous copy synchronous with the host.
 */
typedef struct SdbLineInfo {
    uint8_t  status;             /* SDB_LINE_*                                                   */
    uint8_t  has_r;              /* the line has an R= field (meta.rssi = its text)               */
    uint16_t r_len;
    uint32_t r_off;              /* offset of the R value inside the line                         */
    int32_t  clock;              /* MS: abs(P[CP]) (meta.clock, message_synced.py:239)            */
} SdbLineInfo;
#define SDB_LINE_INVALID  0
#define SDB_LINE_OK       1
#define SDB_LINE_HOSTPATH 2

int sdb_demod_lines_host(SdbHandle *h, int kind,
                         const uint8_t *text, size_t text_len,
                         const uint32_t *line_off, const uint32_t *line_len, uint32_t n,
                         SdbMsgOut *out, SdbHit *hits, uint32_t hits_cap,
                         uint32_t *bits, uint32_t bits_cap, SdbCounters *counters, SdbLineInfo *info);

/*
 * Raw receive buffer -> payload lines (host code, no handle, no device): signalduino/parser/base.py:13-193
 * (strip, ^\x02(M[sSuUcCNOo];.*;)\x03$, decompression of the reduced "Mred=1" format) and the message-type routing
 * of signalduino/parser/__init__.py:43-77.  raw is split on '\n'; raw line i gets line_type[i] = SDB_KIND_MS / _MU /
 * _MC / _MN, SDB_FRAME_OTHER (framed, but no parser for the type) or SDB_FRAME_NONE (not a framed message), and its
 * payload at text[line_off[i] .. +line_len[i]) ('\n' after each, offsets ascending: ready for sdb_demod_lines_host).
 * SDB_FRAME_PYPATH set: the payload needs str.upper() on a non-ASCII character — frame that line in Python.
 * Returns SDB_E_OVERFLOW (with *n_lines / *text_used = what is needed) when max_lines / text_cap are too small.
 */
#define SDB_FRAME_OTHER  4
#define SDB_FRAME_SIDE   0x40      /* sdb_frame_lines_inplace: the (decompressed) payload is in the side buffer */
#define SDB_FRAME_PYPATH 0x80
#define SDB_FRAME_NONE   0xFF
int sdb_frame_lines(const uint8_t *raw, size_t raw_len,
                    uint8_t *text, size_t text_cap,
                    uint32_t *line_off, uint32_t *line_len, uint8_t *line_type, uint32_t max_lines,
                    uint32_t *n_lines, size_t *text_used);

/* The same without copying: the payload of a plain line is raw[line_off[i] .. +line_len[i]) (inside the caller's buffer,
 * STX / ETX excluded); a reduced payload is decompressed into `side` and flagged SDB_FRAME_SIDE (offsets into side).
 * Both offset sequences are ascending, so each subset can go to sdb_demod_lines_host with its own text buffer. */
int sdb_frame_lines_inplace(const uint8_t *raw, size_t raw_len,
                            uint32_t *line_off, uint32_t *line_len, uint8_t *line_type, uint32_t max_lines,
                            uint8_t *side, size_t side_cap, uint32_t *n_lines, size_t *side_used);

/*
 * Host-side JSON of the MS / MU hits of an sdb_demod_lines_host call: one string per hit, exactly
 * MqttPublisher._message_to_json(DecodedMessage) (signalduino/mqtt.py:228-245) = json.dumps({"protocol_id", "payload",
 * "metadata": {"bit_length", "rssi", "clock"}}, indent=4).  id_pool / id_off[nproto + 1]: the protocol id strings in table
 * order; text / line_off / info: the arguments / results of the lines call (rssi is the text of the R field).
 */
int sdb_format_json(const SdbHandle *h, int kind,
                    const SdbHit *hits, uint32_t nhits, const uint32_t *bits,
                    const char *id_pool, const uint32_t *id_off,
                    const uint8_t *text, const uint32_t *line_off, const SdbLineInfo *info,
                    char *pool, size_t pool_cap, uint64_t *str_off, size_t *pool_used);

/*
 * Unit-op entry points: run ONE device function on ONE input (a 1-warp launch).  They back
 * the scalar helper methods of the drop-in class so that the reference's own unit tests
 * (tests/test_postdemodulation.py, tests/test_manchester_protocols.py, tests/test_helpers.py)
 * exercise the device code.  bits are one byte per bit (0/1).
 */
int sdb_unit_postdemod(SdbHandle *h, int method, const uint8_t *bits_in, uint32_t n_in,
                       uint8_t *bits_out, uint32_t out_cap, uint32_t *n_out, int *rcode);

/*
 * One mcBit2* / mcRaw call (sd_protocols/manchester.py:207-795, helpers.py:90-122) on a bit string given as one
 * byte per bit, with the length rules of protocol `proto` (table-order index) and the caller's mcbitnum.
 * proto == 0xFFFFFFFF: an id that is not in the table (every property takes the reference's default; bit 8 of
 * method_override says the id is 119).  method_override & 0xFF != 0 runs that SDB_M_* decoder instead of the
 * protocol's own.  rcode: 1 ok, -1 rejected
 * (`reason` = SDB_MCR_* code), <= -100: the reference raises (-(100 + SDB_ST_*)).  For mcBit2TFA the duplicate
 * parts are concatenated in bits_out and seg[] holds their lengths.
 */
int sdb_unit_mc(SdbHandle *h, uint32_t proto, int method_override, const uint8_t *bits, uint32_t n, int mcbitnum,
                uint8_t *bits_out, uint32_t out_cap, uint32_t *n_out, int32_t *seg, uint32_t seg_cap,
                uint32_t *n_seg, int *rcode, int *reason);

/*
 * One pattern_exists call (sd_protocols/pattern_utils.py:34-136) through the warp-level resolver of the MS / MU kernels:
 * does the pulse template occur in the digit string, and with which pattern ids?
 *   tpl / tpl_len   one compiled template (48 bytes: pysignalduino_b200/table.py KEYTPL_DTYPE = csrc/sdb_table.h SdbKeyTpl:
 *                   length, distinct values, their accepted tenths intervals, offsets into `rank`)
 *   rank / n_rank   dense gap ranks of the template's distinct values over their intervals (candidate order, :83)
 *   tenths[8]       10 * pattern value per slot (slot order = dict order); pat_ids = nibble s: the id digit of slot s
 *   digits / dlen   the data string, nibble-packed as in SdbPulseMsg (16-byte padded with 0xF)
 * *found = 1: target_digits receives the len matched id digits (one per byte), *pos the first occurrence; 0: returns -1 there.
 */
int sdb_unit_pattern_exists(SdbHandle *h, const void *tpl, size_t tpl_len, const uint16_t *rank, uint32_t n_rank,
                            const int16_t *tenths, uint32_t pat_ids, uint32_t npat,
                            const uint8_t *digits, size_t digits_len, uint32_t dlen,
                            int *found, uint8_t *target_digits, uint32_t target_cap, int *pos);

/*
 * Bounds-check build only (libsdb200_chk.so, -DSDB_BOUNDS_CHECK): number of out-of-range shared-memory
 * indices the kernels have seen so far (compute-sanitizer is not available on the GPU pool).
 * Returns 0xFFFFFFFF from the normal build.
 */
unsigned int sdb_debug_violations(SdbHandle *h, int reset);

#ifdef __cplusplus
}
#endif
#endif /* SDB200_H */
