/*
 * sdb_capi.cu — the extern "C" boundary of libsdb200.so (declared in include/sdb200.h).
 *
 * Host logic only: table upload, buffer management, kernel launches, host<->device copies and
 * result formatting.  There is NO CPU decode path in this library: without a CUDA device
 * sdb_create() fails with SDB_E_NOGPU.
 */
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sdb200.h"
#include "sdb_fmt.h"
#include "sdb_pulse.h"
#include "sdb_table.h"

struct SdbHandle {
    int device = 0;
    int sm_count = 0;
    int grid_ms[3] = {0, 0, 0}, grid_mu[3] = {0, 0, 0}, grid_hex = 0, grid_long = 0;   /* MS / MU: resolve, match or scan, emit */
    std::vector<uint8_t> blob;          /* host copy (formatting needs preamble / flags) */
    uint8_t *d_blob = nullptr;
    SdbDevTable tab{};
    std::string err;
    /* buffers of the host-pointer convenience path (grown on demand) */
    void *d_msgs = nullptr;   size_t cap_msgs = 0;
    uint8_t *d_digits = nullptr; size_t cap_digits = 0;
    SdbMsgOut *d_out = nullptr;  size_t cap_out = 0;
    SdbHit *d_hits = nullptr;    size_t cap_hits = 0;
    uint32_t *d_bits = nullptr;  size_t cap_bits = 0;
    SdbCounters *d_ctr = nullptr;
    uint8_t *d_unit = nullptr;          /* unit-op scratch */
    uint8_t *d_text = nullptr;  size_t cap_text = 0;     /* sdb_demod_lines_host: line text */
    uint8_t *d_lines = nullptr; size_t cap_lines = 0;    /* line offsets, lengths, SdbLineInfo */
    void *d_scratch = nullptr;          /* what the pulse kernels of a launch group hand each other (sdb_pulse.h), allocated on the first MS / MU call */
    SdbScratchCfg scfg{0, 0, 0, 0, 0};     /* ... as allocated */
    uint32_t want_surv = SDB_SURV_AVG_DEFAULT, want_match = SDB_MATCH_AVG_DEFAULT, want_ovf = SDB_OVF_MAX_DEFAULT;   /* budgets (grown from the recorded need) */
    uint32_t slack_warps = 0;           /* 0 = one arena block per resident warp on top of the budgets; else that many (tests of the overflow paths) */
    uint32_t *h_stats = nullptr;        /* pinned copy of the scratch statistics */
    uint32_t seen[3] = {0, 0, 0};       /* largest survivor-arena claim / match-arena claim / overflow list any check has read back */
    cudaStream_t stream = nullptr;      /* compute + final copies of the host-buffer path */
    cudaStream_t copy_stream = nullptr; /* pipelined H2D */
    cudaStream_t d2h_stream = nullptr;  /* pipelined D2H of the per-message result slots */
    std::vector<cudaEvent_t> ev_h2d, ev_done, ev_d2h;
    SdbCounters *h_snap = nullptr;      /* pinned: counters after each chunk (which arena ranges are final) */
    uint32_t *h_used = nullptr;         /* pinned: payload-pool bytes in use after each chunk */
    uint32_t snap_cap = 0;
    uint16_t *d_rowmap = nullptr;       /* [nproto] MS rows, then [nproto] MU rows: table-order protocol index -> row (format kernel) */
    char *d_chars = nullptr;    size_t cap_chars = 0;     /* device payload pool of sdb_demod_host_payloads */
    SdbPayloadHit *d_phits = nullptr; size_t cap_phits = 0;
    uint32_t *d_fmt = nullptr;          /* [0] first hit not formatted yet, [1] scratch, [2] pool bytes handed out */
};

static int enqueue_pulse(SdbHandle *h, int kind, const SdbPulseMsg *d_msgs, const uint8_t *d_digits, uint32_t n,
                         uint32_t msg_base, SdbMsgOut *d_out, SdbHit *d_hits, uint32_t hits_cap,
                         uint32_t *d_bits, uint32_t bits_cap, SdbCounters *d_counters, cudaStream_t st);

static thread_local std::string g_create_err;

static int set_err(SdbHandle *h, int code, const char *what, cudaError_t ce = cudaSuccess)
{
    std::string s = what;
    if (ce != cudaSuccess) { s += ": "; s += cudaGetErrorString(ce); }
    if (h) h->err = s; else g_create_err = s;
    return code;
}

#define CK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) return set_err(h, SDB_E_CUDA, #call, _e); } while (0)

extern "C" int sdb_abi_version(void) { return SDB_ABI_VERSION; }

extern "C" unsigned int sdb_debug_violations(SdbHandle *h, int reset)
{
    if (!h) return 0xFFFFFFFFu;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    return sdb::debug_violations(reset != 0);
}

extern "C" const char *sdb_last_error(const SdbHandle *h) { return h ? h->err.c_str() : g_create_err.c_str(); }

extern "C" int sdb_create(const void *blob, size_t blob_len, int device, SdbHandle **out)
{
    SdbHandle *h = nullptr;
    if (!blob || !out || blob_len < sizeof(SdbTblHeader)) return set_err(nullptr, SDB_E_ARG, "sdb_create: bad arguments");
    const SdbTblHeader *hd = static_cast<const SdbTblHeader *>(blob);
    if (hd->magic != SDB_TBL_MAGIC || hd->version != SDB_TBL_VERSION || hd->total != blob_len)
        return set_err(nullptr, SDB_E_ARG, "sdb_create: not a protocol table blob of this version");
    if (hd->n_clk > SDB_MAX_CLK) return set_err(nullptr, SDB_E_ARG, "sdb_create: too many distinct clocks");
    if (hd->n_vals > SDB_MAX_VALS) return set_err(nullptr, SDB_E_ARG, "sdb_create: too many distinct (clock, interval) pairs");
    if (hd->n_ms > 255 || hd->n_mu > 255) return set_err(nullptr, SDB_E_ARG, "sdb_create: more than 255 protocols in one class");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) return set_err(nullptr, SDB_E_NOGPU, "sdb_create: no CUDA device (there is no CPU fallback)", ce);
    if (device < 0 || device >= ndev) return set_err(nullptr, SDB_E_ARG, "sdb_create: bad device index");
    h = new SdbHandle();
    h->device = device;
    auto fail = [&](int code) { std::string e = h->err; sdb_destroy(h); g_create_err = e; return code; };
#define CKC(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { set_err(h, SDB_E_CUDA, #call, _e); return fail(SDB_E_CUDA); } } while (0)
    CKC(cudaSetDevice(device));
    cudaDeviceProp prop;
    CKC(cudaGetDeviceProperties(&prop, device));
    h->sm_count = prop.multiProcessorCount;
    h->blob.assign(static_cast<const uint8_t *>(blob), static_cast<const uint8_t *>(blob) + blob_len);
    CKC(cudaMalloc(&h->d_blob, blob_len));
    CKC(cudaMemcpy(h->d_blob, blob, blob_len, cudaMemcpyHostToDevice));
    CKC(cudaMalloc(&h->d_ctr, sizeof(SdbCounters)));
    CKC(cudaMalloc(&h->d_fmt, 4 * sizeof(uint32_t)));
    {
        /* protocol index -> pulse row, for the device formatter */
        std::vector<uint16_t> map(2 * (size_t)hd->nproto, 0);
        const SdbPulseProto *ms = reinterpret_cast<const SdbPulseProto *>(static_cast<const uint8_t *>(blob) + hd->off_ms);
        const SdbPulseProto *mu = reinterpret_cast<const SdbPulseProto *>(static_cast<const uint8_t *>(blob) + hd->off_mu);
        for (uint32_t i = 0; i < hd->n_ms; i++) if (ms[i].proto < hd->nproto) map[ms[i].proto] = (uint16_t)i;
        for (uint32_t i = 0; i < hd->n_mu; i++) if (mu[i].proto < hd->nproto) map[hd->nproto + mu[i].proto] = (uint16_t)i;
        CKC(cudaMalloc(&h->d_rowmap, map.size() * sizeof(uint16_t) + 16));
        CKC(cudaMemcpy(h->d_rowmap, map.data(), map.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    }
    CKC(cudaMalloc(&h->d_unit, 32768));
    CKC(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CKC(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    CKC(cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking));
    const uint8_t *b = h->d_blob;
    h->tab.ms = reinterpret_cast<const SdbPulseProto *>(b + hd->off_ms);
    h->tab.mu = reinterpret_cast<const SdbPulseProto *>(b + hd->off_mu);
    h->tab.ms_pf = reinterpret_cast<const SdbPrefilter *>(b + hd->off_ms_pf);
    h->tab.mu_pf = reinterpret_cast<const SdbPrefilter *>(b + hd->off_mu_pf);
    h->tab.clk = reinterpret_cast<const double *>(b + hd->off_clk);
    h->tab.rank = reinterpret_cast<const uint16_t *>(b + hd->off_rank);
    h->tab.vals = reinterpret_cast<const SdbValRow *>(b + hd->off_vals);
    h->tab.kill = reinterpret_cast<const uint32_t *>(b + hd->off_kill);
    h->tab.n_vals = hd->n_vals; h->tab.n_mu_vals = hd->n_mu_vals;
    h->tab.mm = reinterpret_cast<const SdbMmItem *>(b + hd->off_mm);
    h->tab.hex = reinterpret_cast<const SdbHexProto *>(b + hd->off_hex);
    h->tab.n_ms = hd->n_ms; h->tab.n_mu = hd->n_mu; h->tab.n_clk = hd->n_clk; h->tab.nproto = hd->nproto;
    /* persistent grids: every SM filled with as many CTAs as fit */
    sdb::pulse_blocks_per_sm(SDB_KIND_MS, h->tab, h->grid_ms);
    sdb::pulse_blocks_per_sm(SDB_KIND_MU, h->tab, h->grid_mu);
    for (int i = 0; i < 3; i++) { h->grid_ms[i] *= h->sm_count; h->grid_mu[i] *= h->sm_count; }
    CKC(cudaMallocHost(reinterpret_cast<void **>(&h->h_stats), SDB_STAT_WORDS * sizeof(uint32_t)));
    /* hex kernel: one thread per message, grid-stride; 16 CTAs of 128 threads are resident per SM, and three times that many
     * CTAs even out the very different message costs (sweep 8 / 16 / 24 / 32 / 48 / 64 CTAs per SM: MC 1.33 / 1.49 / 1.54 / 1.61 /
     * 1.68 / 1.69 G msg/s, MN 1.79 ... 1.96; tools/hex_sweep.py with SDB_HEX_CTAS) */
    h->grid_hex = h->sm_count * (getenv("SDB_HEX_CTAS") ? atoi(getenv("SDB_HEX_CTAS")) : 48);
    h->grid_long = h->sm_count * sdb_long::long_blocks_per_sm(h->tab);
#undef CKC
    *out = h;
    return SDB_OK;
}

extern "C" void sdb_destroy(SdbHandle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaFree(h->d_blob); cudaFree(h->d_ctr); cudaFree(h->d_unit); cudaFree(h->d_scratch); cudaFree(h->d_text); cudaFree(h->d_lines);
    cudaFree(h->d_msgs); cudaFree(h->d_digits); cudaFree(h->d_out); cudaFree(h->d_hits); cudaFree(h->d_bits);
    cudaFree(h->d_rowmap); cudaFree(h->d_chars); cudaFree(h->d_phits); cudaFree(h->d_fmt);
    if (h->h_used) cudaFreeHost(h->h_used);
    if (h->h_stats) cudaFreeHost(h->h_stats);
    for (cudaEvent_t e : h->ev_h2d) cudaEventDestroy(e);
    for (cudaEvent_t e : h->ev_done) cudaEventDestroy(e);
    for (cudaEvent_t e : h->ev_d2h) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->h_snap) cudaFreeHost(h->h_snap);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->d2h_stream) cudaStreamDestroy(h->d2h_stream);
    delete h;
}

extern "C" int sdb_demod_pulse_device(SdbHandle *h, int kind,
                                      const SdbPulseMsg *d_msgs, const uint8_t *d_digits, uint32_t n,
                                      SdbMsgOut *d_out, SdbHit *d_hits, uint32_t hits_cap,
                                      uint32_t *d_bits, uint32_t bits_cap,
                                      SdbCounters *d_counters, void *stream)
{
    if (!h) return SDB_E_ARG;
    if (kind != SDB_KIND_MS && kind != SDB_KIND_MU) return set_err(h, SDB_E_ARG, "sdb_demod_pulse_device: kind must be MS or MU");
    if (n && (!d_msgs || !d_digits || !d_out || !d_counters)) return set_err(h, SDB_E_ARG, "sdb_demod_pulse_device: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(h->device));
    CK(cudaMemsetAsync(d_counters, 0, sizeof(SdbCounters), st));
    return enqueue_pulse(h, kind, d_msgs, d_digits, n, 0, d_out, d_hits, hits_cap, d_bits, bits_cap, d_counters, st);
}

/* ---- scratch of the pulse kernels (sdb_pulse.h: compact survivor / match arenas sized by average record counts) ---- */
static uint32_t pulse_stride(const SdbHandle *h) { return h->tab.n_ms > h->tab.n_mu ? h->tab.n_ms : h->tab.n_mu; }

/* (re)allocate when the launch group or one of the budgets outgrew the block; st is synchronised first (the kernels in
 * flight use the old block) */
static int ensure_scratch(SdbHandle *h, uint32_t n, cudaStream_t st)
{
    const uint32_t chunk0 = n < SDB_MU_CHUNK ? ((n + 1023u) & ~1023u) : SDB_MU_CHUNK;
    const uint32_t chunk = chunk0 > h->scfg.chunk ? chunk0 : h->scfg.chunk;
    const uint32_t stride = pulse_stride(h);
    SdbScratchCfg c;
    c.chunk = chunk;
    c.surv_avg = h->want_surv < stride ? h->want_surv : stride;
    c.match_avg = h->want_match < 64u ? h->want_match : 64u;
    c.ovf_max = h->want_ovf < chunk ? h->want_ovf : chunk;
    {
        /* warps of the largest persistent grid; a launch group of `chunk` messages never starts more than chunk + 8 */
        int g = h->grid_ms[0];
        for (int i = 0; i < 3; i++) { if (h->grid_ms[i] > g) g = h->grid_ms[i]; if (h->grid_mu[i] > g) g = h->grid_mu[i]; }
        const uint32_t w = (uint32_t)g * (SDB_PULSE_THREADS / 32) + (uint32_t)h->grid_long * (SDB_LONG_THREADS / 32);
        c.warps = w < chunk + 8 ? w : chunk + 8;
        if (h->slack_warps && h->slack_warps < c.warps) c.warps = h->slack_warps;
    }
    if (h->d_scratch && c.chunk <= h->scfg.chunk && c.surv_avg <= h->scfg.surv_avg && c.match_avg <= h->scfg.match_avg &&
        c.ovf_max <= h->scfg.ovf_max)
        return SDB_OK;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(st));
    CK(cudaDeviceSynchronize());
    if (h->d_scratch) CK(cudaFree(h->d_scratch));
    h->d_scratch = nullptr; h->scfg = SdbScratchCfg{0, 0, 0, 0, 0};
    const size_t bytes = sdb::pulse_scratch_bytes(stride, c);
    CK(cudaMalloc(&h->d_scratch, bytes));
    /* counters and statistics start at zero (the lists and arenas need no initialisation) */
    CK(cudaMemset(static_cast<uint8_t *>(h->d_scratch) + sdb::pulse_scratch_ctl_offset(stride, c), 0, SDB_CTL_WORDS * sizeof(uint32_t)));
    CK(cudaMemset(static_cast<uint8_t *>(h->d_scratch) + sdb::pulse_scratch_stats_offset(stride, c), 0, SDB_STAT_WORDS * sizeof(uint32_t)));
    h->scfg = c;
    return SDB_OK;
}

/* Read (and reset) the scratch statistics after the work on `st` has finished; grow the budgets from the recorded need so that
 * the next call — or the repetition of this one, when messages were flagged SDB_ST_SCRATCH — finds room.  *n_short = messages
 * that were flagged since the last check. */
static int check_scratch(SdbHandle *h, cudaStream_t st, uint32_t *n_short)
{
    *n_short = 0;
    if (!h->d_scratch) return SDB_OK;
    const uint32_t stride = pulse_stride(h);
    uint8_t *stats = static_cast<uint8_t *>(h->d_scratch) + sdb::pulse_scratch_stats_offset(stride, h->scfg);
    CK(cudaMemcpyAsync(h->h_stats, stats, SDB_STAT_WORDS * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaMemsetAsync(stats, 0, SDB_STAT_WORDS * sizeof(uint32_t), st));
    CK(cudaStreamSynchronize(st));
    const uint64_t chunk = h->scfg.chunk ? h->scfg.chunk : 1;
    const uint64_t need_surv = h->h_stats[SDB_STAT_SURV], need_match = h->h_stats[SDB_STAT_MATCH], need_ovf = h->h_stats[SDB_STAT_OVF];
    if (need_surv > h->seen[0]) h->seen[0] = (uint32_t)need_surv;
    if (need_match > h->seen[1]) h->seen[1] = (uint32_t)need_match;
    if (need_ovf > h->seen[2]) h->seen[2] = (uint32_t)need_ovf;
    uint32_t caps[2];
    sdb::pulse_scratch_caps(stride, h->scfg, caps);
    if (need_surv > caps[0]) {                                  /* the overflow pass had to help: 25 % above the need from now on */
        const uint64_t per = need_surv - (uint64_t)h->scfg.warps * SDB_SURV_BLOCK;
        const uint32_t w = (uint32_t)((per + per / 4 + chunk - 1) / chunk);
        if (w > h->want_surv) h->want_surv = w;
    }
    if (need_match > caps[1]) {
        const uint64_t per = need_match - (uint64_t)h->scfg.warps * SDB_MATCH_BLOCK;
        const uint32_t w = (uint32_t)((per + per / 4 + chunk - 1) / chunk);
        if (w > h->want_match) h->want_match = w;
    }
    if (need_ovf > h->scfg.ovf_max) {
        const uint32_t w = (uint32_t)(2 * need_ovf);
        if (w > h->want_ovf) h->want_ovf = w;
    }
    *n_short = h->h_stats[SDB_STAT_SHORT];
    if (*n_short) h->slack_warps = 0;
    if (*n_short && h->want_surv <= h->scfg.surv_avg && h->want_ovf <= h->scfg.ovf_max) {
        /* flagged although the recorded need fits (long messages share the arena): double */
        h->want_surv = 2 * h->scfg.surv_avg; h->want_ovf = 2 * h->scfg.ovf_max;
    }
    return SDB_OK;
}

/* Scratch for launch groups of up to n_messages (the survivor / match records the kernels of one group hand each other).
 * After this call sdb_demod_pulse_device() with n <= n_messages per launch group only enqueues work. */
extern "C" int sdb_reserve(SdbHandle *h, uint32_t n_messages)
{
    if (!h) return SDB_E_ARG;
    CK(cudaSetDevice(h->device));
    return ensure_scratch(h, n_messages, h->stream);
}

/* Device-pointer calls cannot repeat themselves: after synchronising, this says how many messages of the calls since the last
 * check were flagged SDB_ST_SCRATCH (not decoded: the compact scratch arenas were too small for that batch) and grows the
 * budgets, so that submitting those messages again succeeds.  0 with the shipped table unless a batch averages more than
 * 18 surviving protocols per message. */
extern "C" int sdb_scratch_short(SdbHandle *h, uint32_t *n_short)
{
    if (!h || !n_short) return SDB_E_ARG;
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    return check_scratch(h, h->stream, n_short);
}

/* Set the budgets the scratch is sized by (0 = keep): survivor records and MU match records per message on average, messages
 * of the worst-case overflow region; slack_warps (0 = automatic: every resident warp) caps the number of arena blocks that are
 * added on top of the per-message budgets — only the tests of the overflow paths want less than automatic.  The current block
 * is released; the next call allocates with the new budgets. */
extern "C" int sdb_scratch_budget(SdbHandle *h, uint32_t surv_avg, uint32_t match_avg, uint32_t ovf_max, uint32_t slack_warps)
{
    if (!h) return SDB_E_ARG;
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    if (surv_avg) h->want_surv = surv_avg;
    if (match_avg) h->want_match = match_avg;
    if (ovf_max) h->want_ovf = ovf_max;
    h->slack_warps = slack_warps;
    if (h->d_scratch) CK(cudaFree(h->d_scratch));
    h->d_scratch = nullptr; h->scfg = SdbScratchCfg{0, 0, 0, 0, 0};
    return SDB_OK;
}

/* bytes of the scratch block as allocated; cfg = {messages per launch group, surv_avg, match_avg, ovf_max, arena blocks on top
 * (warps), then the largest values any sdb_scratch_short() / host-buffer call has read back: survivor-arena claim of a launch
 * group (records), match-arena claim, overflow list} */
extern "C" size_t sdb_scratch_info(const SdbHandle *h, uint32_t cfg[8])
{
    if (cfg) for (int i = 0; i < 8; i++) cfg[i] = 0;
    if (!h || !h->d_scratch) return 0;
    if (cfg) {
        cfg[0] = h->scfg.chunk; cfg[1] = h->scfg.surv_avg; cfg[2] = h->scfg.match_avg; cfg[3] = h->scfg.ovf_max; cfg[4] = h->scfg.warps;
        cfg[5] = h->seen[0]; cfg[6] = h->seen[1]; cfg[7] = h->seen[2];
    }
    return sdb::pulse_scratch_bytes(pulse_stride(h), h->scfg);
}

/* Enqueue the kernels for messages [0, n) at d_msgs whose batch indices start at msg_base; counters are NOT reset. */
static int enqueue_pulse(SdbHandle *h, int kind, const SdbPulseMsg *d_msgs, const uint8_t *d_digits, uint32_t n,
                         uint32_t msg_base, SdbMsgOut *d_out, SdbHit *d_hits, uint32_t hits_cap,
                         uint32_t *d_bits, uint32_t bits_cap, SdbCounters *d_counters, cudaStream_t st)
{
    const int *grid = kind == SDB_KIND_MS ? h->grid_ms : h->grid_mu;
    if (n) {
        int rc = ensure_scratch(h, n, st);
        if (rc != SDB_OK) return rc;
    }
    int rc = sdb::launch_pulse(kind, h->tab, d_msgs, d_digits, n, d_out, d_hits, hits_cap, d_bits, bits_cap, d_counters, grid,
                               h->grid_long, h->d_scratch, h->scfg, msg_base, st);
    if (rc != 0) return set_err(h, SDB_E_CUDA, "pulse kernel launch", static_cast<cudaError_t>(rc));
    return SDB_OK;
}

extern "C" int sdb_demod_hex_device(SdbHandle *h, int kind, int mc_repaired,
                                    const SdbHexMsg *d_msgs, const uint8_t *d_digits, uint32_t n,
                                    SdbMsgOut *d_out, SdbHit *d_hits, uint32_t hits_cap,
                                    uint32_t *d_bits, uint32_t bits_cap,
                                    SdbCounters *d_counters, void *stream)
{
    if (!h) return SDB_E_ARG;
    if (kind != SDB_KIND_MC && kind != SDB_KIND_MN) return set_err(h, SDB_E_ARG, "sdb_demod_hex_device: kind must be MC or MN");
    if (n && (!d_msgs || !d_digits || !d_out || !d_counters)) return set_err(h, SDB_E_ARG, "sdb_demod_hex_device: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(h->device));
    CK(cudaMemsetAsync(d_counters, 0, sizeof(SdbCounters), st));
    int grid = h->grid_hex;
    uint32_t need = (n + SDB_HEX_THREADS - 1) / SDB_HEX_THREADS;
    if (need < (uint32_t)grid) grid = (int)(need ? need : 1);
    int rc = sdb::launch_hex(kind, mc_repaired, h->tab, d_msgs, d_digits, n, d_out, d_hits, hits_cap, d_bits, bits_cap, d_counters, grid, st);
    if (rc != 0) return set_err(h, SDB_E_CUDA, "hex kernel launch", static_cast<cudaError_t>(rc));
    return SDB_OK;
}

/* The pipelined host calls queue asynchronous copies into the caller's buffers on three streams: before an error code goes
 * back to the caller nothing of that may still be in flight. */
static int settle(SdbHandle *h, int rc)
{
    if (rc != SDB_OK && h) {
        if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
        if (h->stream) cudaStreamSynchronize(h->stream);
        if (h->d2h_stream) cudaStreamSynchronize(h->d2h_stream);
    }
    return rc;
}

template <typename T>
static int grow(SdbHandle *h, T *&p, size_t &cap, size_t need_bytes)
{
    if (need_bytes <= cap) return SDB_OK;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = need_bytes + need_bytes / 4 + 256;
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&p), want);
    if (e != cudaSuccess) return set_err(h, SDB_E_CUDA, "cudaMalloc", e);
    cap = want;
    return SDB_OK;
}

/* events + pinned counter snapshots for an nchunks-deep pipeline */
static int pipeline_prepare(SdbHandle *h, uint32_t nchunks)
{
    while (h->ev_h2d.size() < nchunks) {
        cudaEvent_t a, b, c;
        CK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c, cudaEventDisableTiming));
        h->ev_h2d.push_back(a); h->ev_done.push_back(b); h->ev_d2h.push_back(c);
    }
    if (h->snap_cap < nchunks) {
        if (h->h_snap) cudaFreeHost(h->h_snap);
        h->h_snap = nullptr; h->snap_cap = 0;
        CK(cudaMallocHost(reinterpret_cast<void **>(&h->h_snap), sizeof(SdbCounters) * (size_t)(nchunks + 8)));
        if (h->h_used) cudaFreeHost(h->h_used);
        h->h_used = nullptr;
        CK(cudaMallocHost(reinterpret_cast<void **>(&h->h_used), sizeof(uint32_t) * (size_t)(nchunks + 8)));
        h->snap_cap = nchunks + 8;
    }
    return SDB_OK;
}

/* Hits are appended through one atomic counter and the chunks run in stream order, so once chunk k is done the arena
 * ranges [done, snapshot k) are final: copy them to the host while the next chunk's kernels run, instead of one big
 * D2H after the last kernel (0.7 GB per 10 M mixed messages). */
struct ArenaDrain { uint32_t hits = 0, words = 0, chars = 0; };
/* sdb_demod_host_payloads: the strings of every stage are formatted on the device (sdb_format.cu) right after its decode
 * kernels and copied back with the hits, under the next stage's kernels */
struct PayloadSink {
    char *pool = nullptr; uint32_t pool_cap = 0;
    SdbPayloadHit *phits = nullptr;
    uint32_t used = 0;
};
static int drain_chunk(SdbHandle *h, uint32_t k, ArenaDrain &d, SdbHit *hits, uint32_t hits_cap, uint32_t *bits, uint32_t bits_cap,
                       PayloadSink *sink = nullptr)
{
    CK(cudaEventSynchronize(h->ev_done[k]));            /* chunk k + 1 is already queued: the GPU stays busy */
    const SdbCounters c = h->h_snap[k];
    if (c.hits > hits_cap || c.words > bits_cap) return SDB_OK;      /* overflow: reported after the last chunk */
    if (c.hits > d.hits && hits)
        CK(cudaMemcpyAsync(hits + d.hits, h->d_hits + d.hits, sizeof(SdbHit) * (size_t)(c.hits - d.hits), cudaMemcpyDeviceToHost, h->d2h_stream));
    if (c.words > d.words && bits)
        CK(cudaMemcpyAsync(bits + d.words, h->d_bits + d.words, sizeof(uint32_t) * (size_t)(c.words - d.words), cudaMemcpyDeviceToHost, h->d2h_stream));
    if (sink) {
        const uint32_t used = h->h_used[k];
        sink->used = used;
        if (used <= sink->pool_cap) {
            if (c.hits > d.hits)
                CK(cudaMemcpyAsync(sink->phits + d.hits, h->d_phits + d.hits, sizeof(SdbPayloadHit) * (size_t)(c.hits - d.hits), cudaMemcpyDeviceToHost, h->d2h_stream));
            if (used > d.chars)
                CK(cudaMemcpyAsync(sink->pool + d.chars, h->d_chars + d.chars, (size_t)(used - d.chars), cudaMemcpyDeviceToHost, h->d2h_stream));
            d.chars = used;
        }
    }
    d.hits = c.hits; d.words = c.words;
    return SDB_OK;
}

/* format kernel of the stage just decoded + snapshot of the pool counter (payload mode, MS / MU) */
static int enqueue_format(SdbHandle *h, int kind, uint32_t hits_cap, uint32_t bits_cap, PayloadSink *sink, uint32_t k, cudaStream_t st)
{
    const SdbPulseProto *rows = kind == SDB_KIND_MS ? h->tab.ms : h->tab.mu;
    const uint16_t *map = h->d_rowmap + (kind == SDB_KIND_MS ? 0 : h->tab.nproto);
    int rc = sdb::launch_format(kind, h->d_hits, h->d_bits, rows, map, h->tab.hex, h->tab.nproto, h->d_fmt, h->d_ctr, hits_cap, bits_cap,
                                h->d_chars, sink->pool_cap, h->d_phits, h->d_fmt + 2, h->sm_count * 8, st);
    if (rc != 0) return set_err(h, SDB_E_CUDA, "format kernel launch", static_cast<cudaError_t>(rc));
    CK(cudaMemcpyAsync(&h->h_used[k], h->d_fmt + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    return SDB_OK;
}

static int demod_host_once(SdbHandle *h, int kind, int mc_repaired,
                           const void *msgs, const uint8_t *digits, size_t digits_len, uint32_t n,
                           SdbMsgOut *out, SdbHit *hits, uint32_t hits_cap,
                           uint32_t *bits, uint32_t bits_cap, SdbCounters *counters, PayloadSink *sink);

/* The compact scratch arenas are sized by average record counts: a batch that needs more has its excess messages flagged
 * SDB_ST_SCRATCH by the kernels; the host paths then grow the scratch from the recorded need and run the call again, so that
 * callers of the host-buffer entry points never see that status. */
#define SDB_SCRATCH_RETRIES 6
static int demod_host_impl(SdbHandle *h, int kind, int mc_repaired,
                           const void *msgs, const uint8_t *digits, size_t digits_len, uint32_t n,
                           SdbMsgOut *out, SdbHit *hits, uint32_t hits_cap,
                           uint32_t *bits, uint32_t bits_cap, SdbCounters *counters, PayloadSink *sink)
{
    for (int attempt = 0;; attempt++) {
        int rc = demod_host_once(h, kind, mc_repaired, msgs, digits, digits_len, n, out, hits, hits_cap, bits, bits_cap, counters, sink);
        if (rc != SDB_OK || (kind != SDB_KIND_MS && kind != SDB_KIND_MU)) return rc;
        uint32_t n_short = 0;
        if ((rc = check_scratch(h, h->stream, &n_short)) != SDB_OK || !n_short) return rc;
        if (attempt == SDB_SCRATCH_RETRIES) return set_err(h, SDB_E_SCRATCH, "pulse scratch: messages still flagged SDB_ST_SCRATCH after growing it");
    }
}

static int demod_host_once(SdbHandle *h, int kind, int mc_repaired,
                           const void *msgs, const uint8_t *digits, size_t digits_len, uint32_t n,
                           SdbMsgOut *out, SdbHit *hits, uint32_t hits_cap,
                           uint32_t *bits, uint32_t bits_cap, SdbCounters *counters, PayloadSink *sink)
{
    if (!h) return SDB_E_ARG;
    if (!counters || (n && (!msgs || !digits || !out))) return set_err(h, SDB_E_ARG, "sdb_demod_host: null pointer");
    const bool pulse = kind == SDB_KIND_MS || kind == SDB_KIND_MU;
    if (!pulse && kind != SDB_KIND_MC && kind != SDB_KIND_MN) return set_err(h, SDB_E_ARG, "sdb_demod_host: bad kind");
    CK(cudaSetDevice(h->device));
    memset(counters, 0, sizeof *counters);
    if (n == 0) return SDB_OK;
    const size_t rec = pulse ? sizeof(SdbPulseMsg) : sizeof(SdbHexMsg);
    int rc;
    if ((rc = grow(h, reinterpret_cast<uint8_t *&>(h->d_msgs), h->cap_msgs, rec * n))) return rc;
    if ((rc = grow(h, h->d_digits, h->cap_digits, digits_len + 64))) return rc;
    if ((rc = grow(h, h->d_out, h->cap_out, sizeof(SdbMsgOut) * (size_t)n))) return rc;
    if ((rc = grow(h, h->d_hits, h->cap_hits, sizeof(SdbHit) * (size_t)(hits_cap ? hits_cap : 1)))) return rc;
    if ((rc = grow(h, h->d_bits, h->cap_bits, sizeof(uint32_t) * (size_t)(bits_cap ? bits_cap : 1)))) return rc;
    cudaStream_t st = h->stream;
    if (sink) {                                               /* payload mode (MS / MU): device pool + one offset per hit */
        if ((rc = grow(h, h->d_chars, h->cap_chars, (size_t)sink->pool_cap + 16))) return rc;
        if ((rc = grow(h, h->d_phits, h->cap_phits, sizeof(SdbPayloadHit) * (size_t)(hits_cap ? hits_cap : 1)))) return rc;
        if ((rc = pipeline_prepare(h, 1))) return rc;
        CK(cudaMemsetAsync(h->d_fmt, 0, 4 * sizeof(uint32_t), st));
    }
    const SdbPulseMsg *pm = static_cast<const SdbPulseMsg *>(msgs);
    /* Pipelined path (MS / MU, several chunks): the H2D copy of chunk k+1 and the D2H copy of the result slots of
     * chunk k-1 overlap the kernels of chunk k.  Needs the digit streams stored in message order (doff non-decreasing),
     * which is what pack.py and the corpus generator produce; checked at the chunk boundaries. */
    const uint32_t C = SDB_PIPE_CHUNK;
    bool pipelined = pulse && n > C;
    if (pipelined)
        for (uint32_t b = C; b < n; b += C)
            if (pm[b].doff < pm[b - 1].doff || (size_t)pm[b].doff * 16 > digits_len) { pipelined = false; break; }
    if (pipelined) {
        const uint32_t nchunks = (n + C - 1) / C;
        if ((rc = pipeline_prepare(h, nchunks))) return rc;
        SdbPulseMsg *dm = static_cast<SdbPulseMsg *>(h->d_msgs);
        CK(cudaMemsetAsync(h->d_ctr, 0, sizeof(SdbCounters), st));
        auto h2d = [&](uint32_t k) -> int {
            const uint32_t lo = k * C, cnt = n - lo < C ? n - lo : C;
            const size_t dlo = (size_t)pm[lo].doff * 16;
            const size_t dhi = k + 1 < nchunks ? (size_t)pm[lo + cnt].doff * 16 : digits_len;
            CK(cudaMemcpyAsync(dm + lo, pm + lo, sizeof(SdbPulseMsg) * (size_t)cnt, cudaMemcpyHostToDevice, h->copy_stream));
            if (dhi > dlo) CK(cudaMemcpyAsync(h->d_digits + dlo, digits + dlo, dhi - dlo, cudaMemcpyHostToDevice, h->copy_stream));
            CK(cudaEventRecord(h->ev_h2d[k], h->copy_stream));
            return SDB_OK;
        };
        ArenaDrain drain;
        if ((rc = h2d(0))) return rc;
        for (uint32_t k = 0; k < nchunks; k++) {
            const uint32_t lo = k * C, cnt = n - lo < C ? n - lo : C;
            if (k + 1 < nchunks && (rc = h2d(k + 1))) return rc;
            CK(cudaStreamWaitEvent(st, h->ev_h2d[k], 0));
            rc = enqueue_pulse(h, kind, dm + lo, h->d_digits, cnt, lo, h->d_out + lo, h->d_hits, hits_cap, h->d_bits, bits_cap, h->d_ctr, st);
            if (rc != SDB_OK) return rc;
            if (sink && (rc = enqueue_format(h, kind, hits_cap, bits_cap, sink, k, st))) return rc;
            CK(cudaMemcpyAsync(&h->h_snap[k], h->d_ctr, sizeof(SdbCounters), cudaMemcpyDeviceToHost, st));
            CK(cudaEventRecord(h->ev_done[k], st));
            CK(cudaStreamWaitEvent(h->d2h_stream, h->ev_done[k], 0));
            CK(cudaMemcpyAsync(out + lo, h->d_out + lo, sizeof(SdbMsgOut) * (size_t)cnt, cudaMemcpyDeviceToHost, h->d2h_stream));
            if (k && (rc = drain_chunk(h, k - 1, drain, hits, hits_cap, bits, bits_cap, sink))) return rc;
        }
        if ((rc = drain_chunk(h, nchunks - 1, drain, hits, hits_cap, bits, bits_cap, sink))) return rc;
        *counters = h->h_snap[nchunks - 1];
        CK(cudaStreamSynchronize(h->d2h_stream));
        if (counters->hits > hits_cap || counters->words > bits_cap) return set_err(h, SDB_E_OVERFLOW, "hit / bit arena too small");
        return SDB_OK;
    } else {
        CK(cudaMemcpyAsync(h->d_msgs, msgs, rec * n, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(h->d_digits, digits, digits_len, cudaMemcpyHostToDevice, st));
        if (pulse)
            rc = sdb_demod_pulse_device(h, kind, static_cast<const SdbPulseMsg *>(h->d_msgs), h->d_digits, n, h->d_out,
                                        h->d_hits, hits_cap, h->d_bits, bits_cap, h->d_ctr, st);
        else
            rc = sdb_demod_hex_device(h, kind, mc_repaired, static_cast<const SdbHexMsg *>(h->d_msgs), h->d_digits, n, h->d_out,
                                      h->d_hits, hits_cap, h->d_bits, bits_cap, h->d_ctr, st);
        if (rc != SDB_OK) return rc;
        if (sink && (rc = enqueue_format(h, kind, hits_cap, bits_cap, sink, 0, st))) return rc;
        CK(cudaMemcpyAsync(counters, h->d_ctr, sizeof(SdbCounters), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(out, h->d_out, sizeof(SdbMsgOut) * (size_t)n, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    if (counters->hits > hits_cap || counters->words > bits_cap) return set_err(h, SDB_E_OVERFLOW, "hit / bit arena too small");
    if (counters->hits && hits) CK(cudaMemcpyAsync(hits, h->d_hits, sizeof(SdbHit) * (size_t)counters->hits, cudaMemcpyDeviceToHost, st));
    if (counters->words && bits) CK(cudaMemcpyAsync(bits, h->d_bits, sizeof(uint32_t) * (size_t)counters->words, cudaMemcpyDeviceToHost, st));
    if (sink) {
        sink->used = h->h_used[0];
        if (sink->used <= sink->pool_cap) {
            if (counters->hits) CK(cudaMemcpyAsync(sink->phits, h->d_phits, sizeof(SdbPayloadHit) * (size_t)counters->hits, cudaMemcpyDeviceToHost, st));
            if (sink->used) CK(cudaMemcpyAsync(sink->pool, h->d_chars, sink->used, cudaMemcpyDeviceToHost, st));
        }
    }
    CK(cudaStreamSynchronize(st));
    return SDB_OK;
}

extern "C" int sdb_demod_host(SdbHandle *h, int kind, int mc_repaired,
                              const void *msgs, const uint8_t *digits, size_t digits_len, uint32_t n,
                              SdbMsgOut *out, SdbHit *hits, uint32_t hits_cap,
                              uint32_t *bits, uint32_t bits_cap, SdbCounters *counters)
{
    if (n && (!hits || !bits)) return h ? set_err(h, SDB_E_ARG, "sdb_demod_host: null pointer") : SDB_E_ARG;
    return settle(h, demod_host_impl(h, kind, mc_repaired, msgs, digits, digits_len, n, out, hits, hits_cap, bits, bits_cap, counters, nullptr));
}

static int demod_lines_impl(SdbHandle *h, int kind,
                            const uint8_t *text, size_t text_len,
                            const uint32_t *line_off, const uint32_t *line_len, uint32_t n,
                            SdbMsgOut *out, SdbHit *hits, uint32_t hits_cap,
                            uint32_t *bits, uint32_t bits_cap, SdbCounters *counters, SdbLineInfo *info);

extern "C" int sdb_demod_lines_host(SdbHandle *h, int kind,
                                    const uint8_t *text, size_t text_len,
                                    const uint32_t *line_off, const uint32_t *line_len, uint32_t n,
                                    SdbMsgOut *out, SdbHit *hits, uint32_t hits_cap,
                                    uint32_t *bits, uint32_t bits_cap, SdbCounters *counters, SdbLineInfo *info)
{
    for (int attempt = 0;; attempt++) {
        int rc = settle(h, demod_lines_impl(h, kind, text, text_len, line_off, line_len, n, out, hits, hits_cap, bits, bits_cap, counters, info));
        if (rc != SDB_OK || !h) return rc;
        uint32_t n_short = 0;
        if ((rc = check_scratch(h, h->stream, &n_short)) != SDB_OK || !n_short) return rc;       /* (see demod_host_impl) */
        if (attempt == SDB_SCRATCH_RETRIES) return set_err(h, SDB_E_SCRATCH, "pulse scratch: messages still flagged SDB_ST_SCRATCH after growing it");
    }
}

static int demod_lines_impl(SdbHandle *h, int kind,
                            const uint8_t *text, size_t text_len,
                            const uint32_t *line_off, const uint32_t *line_len, uint32_t n,
                            SdbMsgOut *out, SdbHit *hits, uint32_t hits_cap,
                            uint32_t *bits, uint32_t bits_cap, SdbCounters *counters, SdbLineInfo *info)
{
    if (!h) return SDB_E_ARG;
    if (kind != SDB_KIND_MS && kind != SDB_KIND_MU) return set_err(h, SDB_E_ARG, "sdb_demod_lines_host: kind must be MS or MU");
    if (!counters || (n && (!text || !line_off || !line_len || !out || !info))) return set_err(h, SDB_E_ARG, "sdb_demod_lines_host: null pointer");
    CK(cudaSetDevice(h->device));
    memset(counters, 0, sizeof *counters);
    if (n == 0) return SDB_OK;
    if (text_len >= (1ull << 32)) return set_err(h, SDB_E_ARG, "sdb_demod_lines_host: text larger than 4 GiB");
    int rc;
    const size_t pool_bytes = sdb::lines_pool_bytes(text_len, n);
    if ((rc = grow(h, reinterpret_cast<uint8_t *&>(h->d_msgs), h->cap_msgs, sizeof(SdbPulseMsg) * (size_t)n))) return rc;
    if ((rc = grow(h, h->d_digits, h->cap_digits, pool_bytes + 64))) return rc;
    if ((rc = grow(h, h->d_out, h->cap_out, sizeof(SdbMsgOut) * (size_t)n))) return rc;
    if ((rc = grow(h, h->d_hits, h->cap_hits, sizeof(SdbHit) * (size_t)(hits_cap ? hits_cap : 1)))) return rc;
    if ((rc = grow(h, h->d_bits, h->cap_bits, sizeof(uint32_t) * (size_t)(bits_cap ? bits_cap : 1)))) return rc;
    if ((rc = grow(h, h->d_text, h->cap_text, text_len + 16))) return rc;
    if ((rc = grow(h, h->d_lines, h->cap_lines, (3 * sizeof(uint32_t) + sizeof(SdbLineInfo)) * (size_t)n + 64))) return rc;
    cudaStream_t st = h->stream;
    uint32_t *d_off = reinterpret_cast<uint32_t *>(h->d_lines), *d_len = d_off + n;
    SdbLineInfo *d_info = reinterpret_cast<SdbLineInfo *>(d_len + n);
    uint32_t *d_long = reinterpret_cast<uint32_t *>(d_info + n);      /* [0] list length, [1] work counter, then the list of long lines */
    SdbPulseMsg *dm = static_cast<SdbPulseMsg *>(h->d_msgs);
    /* per stage of SDB_PIPE_CHUNK lines: H2D of the chunk's text / offsets (copy stream) -> tokenizer + demodulation kernels
     * (compute stream) -> D2H of the result slots and line infos (d2h stream); chunk k+1's copy overlaps chunk k's kernels.
     * Device offsets are the caller's own (global) offsets, so digit-pool units and hit.msg need no rebasing. */
    const uint32_t C = SDB_PIPE_CHUNK;
    const uint32_t nchunks = (n + C - 1) / C;
    if ((rc = pipeline_prepare(h, nchunks))) return rc;
    CK(cudaMemsetAsync(h->d_ctr, 0, sizeof(SdbCounters), st));
    auto h2d = [&](uint32_t k) -> int {
        const uint32_t lo = k * C, cnt = n - lo < C ? n - lo : C;
        /* argument check of this stage's lines, here so that it overlaps the kernels of the stages already queued */
        for (uint32_t i = lo; i < lo + cnt; i++)
            if ((size_t)line_off[i] + line_len[i] > text_len || (i && line_off[i] < line_off[i - 1] + line_len[i - 1]))
                return set_err(h, SDB_E_ARG, "sdb_demod_lines_host: lines must be ascending, disjoint and inside the text");
        const size_t t0 = line_off[lo], t1 = (size_t)line_off[lo + cnt - 1] + line_len[lo + cnt - 1];
        CK(cudaMemcpyAsync(h->d_text + t0, text + t0, t1 - t0, cudaMemcpyHostToDevice, h->copy_stream));
        CK(cudaMemcpyAsync(d_off + lo, line_off + lo, sizeof(uint32_t) * (size_t)cnt, cudaMemcpyHostToDevice, h->copy_stream));
        CK(cudaMemcpyAsync(d_len + lo, line_len + lo, sizeof(uint32_t) * (size_t)cnt, cudaMemcpyHostToDevice, h->copy_stream));
        CK(cudaEventRecord(h->ev_h2d[k], h->copy_stream));
        return SDB_OK;
    };
    ArenaDrain drain;
    if ((rc = h2d(0))) return rc;
    for (uint32_t k = 0; k < nchunks; k++) {
        const uint32_t lo = k * C, cnt = n - lo < C ? n - lo : C;
        if (k + 1 < nchunks && (rc = h2d(k + 1))) return rc;
        CK(cudaStreamWaitEvent(st, h->ev_h2d[k], 0));
        rc = sdb::launch_tokenize(kind, h->d_text, d_off + lo, d_len + lo, cnt, lo, dm + lo, h->d_digits, d_info + lo, d_long, h->sm_count, st);
        if (rc != 0) return set_err(h, SDB_E_CUDA, "tokenize kernel launch", static_cast<cudaError_t>(rc));
        rc = enqueue_pulse(h, kind, dm + lo, h->d_digits, cnt, lo, h->d_out + lo, h->d_hits, hits_cap, h->d_bits, bits_cap, h->d_ctr, st);
        if (rc != SDB_OK) return rc;
        CK(cudaMemcpyAsync(&h->h_snap[k], h->d_ctr, sizeof(SdbCounters), cudaMemcpyDeviceToHost, st));
        CK(cudaEventRecord(h->ev_done[k], st));
        CK(cudaStreamWaitEvent(h->d2h_stream, h->ev_done[k], 0));
        CK(cudaMemcpyAsync(out + lo, h->d_out + lo, sizeof(SdbMsgOut) * (size_t)cnt, cudaMemcpyDeviceToHost, h->d2h_stream));
        CK(cudaMemcpyAsync(info + lo, d_info + lo, sizeof(SdbLineInfo) * (size_t)cnt, cudaMemcpyDeviceToHost, h->d2h_stream));
        if (k && (rc = drain_chunk(h, k - 1, drain, hits, hits_cap, bits, bits_cap))) return rc;
    }
    if ((rc = drain_chunk(h, nchunks - 1, drain, hits, hits_cap, bits, bits_cap))) return rc;
    *counters = h->h_snap[nchunks - 1];
    CK(cudaStreamSynchronize(h->d2h_stream));
    if (counters->hits > hits_cap || counters->words > bits_cap) return set_err(h, SDB_E_OVERFLOW, "hit / bit arena too small");
    return SDB_OK;
}

/* ---- host-side formatting --------------------------------------------------------------- */
/* Payload strings of the hits: preamble + hex / bits + postamble.  Two passes over the hits (string lengths -> offsets ->
 * characters), each split over the host threads for large batches: 27 M hits per 10 M-message mixed corpus would take
 * seconds on one core, and the reference side of the comparison (its CPU path) produces these strings too. */
namespace {
struct Fmt {
    const SdbTblHeader *hd;
    int kind;
    bool pulse;
    std::vector<const SdbPulseProto *> row;       /* table-order index -> pulse row (preamble / flags) */
    const SdbHexProto *hx;
    const SdbHit *hits;
    uint32_t nhits;
    const uint32_t *bits;

    /* characters of hit i appended at dst (dst == nullptr: count only); returns the count, or (size_t)-1 for a bad record.
     * sdb_fmt.h holds the formatting itself: the same code the device formatter runs. */
    size_t one(uint32_t i, char *dst) const
    {
        const SdbHit &ht = hits[i];
        if (ht.proto >= hd->nproto) return (size_t)-1;
        if (pulse) {
            const SdbPulseProto *pp = row[ht.proto];
            if (!pp) return (size_t)-1;
            return sdb_fmt_pulse(pp, ht, bits + ht.bits_off, dst);
        }
        return sdb_fmt_hexkind(kind, hx, hits, i, nhits, bits, dst);
    }
};

template <typename F>
void parallel_ranges(uint32_t n, unsigned threads, F fn)
{
    if (threads <= 1 || n < 65536) { fn(0u, 0u, n); return; }
    std::vector<std::thread> th;
    const uint32_t per = (n + threads - 1) / threads;
    for (unsigned t = 0; t < threads; t++) {
        const uint32_t lo = t * per < n ? t * per : n, hi = lo + per < n ? lo + per : n;
        th.emplace_back([=] { fn(t, lo, hi); });
    }
    for (auto &x : th) x.join();
}
}  // namespace

static int fmt_init(Fmt &F, const SdbHandle *h, int kind)
{
    F.hd = reinterpret_cast<const SdbTblHeader *>(h->blob.data());
    F.kind = kind;
    F.pulse = kind == SDB_KIND_MS || kind == SDB_KIND_MU;
    F.hx = reinterpret_cast<const SdbHexProto *>(h->blob.data() + F.hd->off_hex);
    F.hits = nullptr; F.nhits = 0; F.bits = nullptr;
    if (!F.pulse && kind != SDB_KIND_MC && kind != SDB_KIND_MN) return SDB_E_ARG;
    if (F.pulse) {
        F.row.assign(F.hd->nproto, nullptr);
        const SdbPulseProto *tab = reinterpret_cast<const SdbPulseProto *>(h->blob.data() + (kind == SDB_KIND_MS ? F.hd->off_ms : F.hd->off_mu));
        const uint32_t cnt = kind == SDB_KIND_MS ? F.hd->n_ms : F.hd->n_mu;
        for (uint32_t i = 0; i < cnt; i++) F.row[tab[i].proto] = &tab[i];
    }
    return SDB_OK;
}

static unsigned fmt_threads()
{
    unsigned t = std::thread::hardware_concurrency();
    if (t == 0) t = 1;
    return t > 64 ? 64 : t;
}

/* Strings of hits [lo, hi) appended at pool + base (only when they fit); str_off[lo + 1 .. hi] filled (str_off[lo] must
 * hold base).  Returns the new end offset, or UINT64_MAX for a malformed hit record. */
static uint64_t format_range(const Fmt &F, uint32_t lo, uint32_t hi, uint64_t base, char *pool, size_t pool_cap, uint64_t *str_off)
{
    const unsigned threads = fmt_threads();
    const uint32_t n = hi - lo;
    std::vector<uint64_t> part(threads + 1, 0);
    std::vector<int> bad(threads, 0);
    parallel_ranges(n, threads, [&](unsigned t, uint32_t a, uint32_t b) {       /* pass 1: lengths, per-range totals */
        uint64_t sum = 0;
        for (uint32_t i = lo + a; i < lo + b; i++) {
            const size_t len = F.one(i, nullptr);
            if (len == (size_t)-1) { bad[t] = 1; str_off[i + 1] = 0; continue; }
            str_off[i + 1] = len;
            sum += len;
        }
        part[t + 1] = sum;
    });
    for (unsigned t = 0; t < threads; t++) if (bad[t]) return UINT64_MAX;
    part[0] = base;
    for (unsigned t = 0; t < threads; t++) part[t + 1] += part[t];
    const uint64_t end = part[threads];
    const bool fits = end <= pool_cap && (pool || end == 0);
    parallel_ranges(n, threads, [&](unsigned t, uint32_t a, uint32_t b) {       /* pass 2: offsets, then the characters */
        uint64_t at = part[t];                        /* (a serial run is range 0) */
        for (uint32_t i = lo + a; i < lo + b; i++) {
            const uint64_t len = str_off[i + 1];
            if (fits && len) F.one(i, pool + at);
            at += len;
            str_off[i + 1] = at;
        }
    });
    return end;
}

extern "C" int sdb_format_hits(const SdbHandle *h, int kind,
                               const SdbHit *hits, uint32_t nhits, const uint32_t *bits,
                               char *pool, size_t pool_cap, uint64_t *str_off, size_t *pool_used)
{
    if (!h || !str_off || !pool_used || (nhits && !hits)) return SDB_E_ARG;
    Fmt F;
    if (fmt_init(F, h, kind) != SDB_OK) return SDB_E_ARG;
    F.hits = hits; F.nhits = nhits; F.bits = bits;
    str_off[0] = 0;
    const uint64_t end = format_range(F, 0, nhits, 0, pool, pool_cap, str_off);
    if (end == UINT64_MAX) return SDB_E_ARG;
    *pool_used = (size_t)end;
    return end <= pool_cap && (pool || end == 0) ? SDB_OK : SDB_E_OVERFLOW;
}

/* ---- decode + payload strings in one pipelined call --------------------------------------- */
extern "C" int sdb_demod_host_payloads(SdbHandle *h, int kind, int mc_repaired,
                                       const void *msgs, const uint8_t *digits, size_t digits_len, uint32_t n,
                                       SdbMsgOut *out, SdbHit *hits, uint32_t hits_cap,
                                       uint32_t *bits, uint32_t bits_cap, SdbCounters *counters,
                                       char *pool, size_t pool_cap, SdbPayloadHit *phits, size_t *pool_used)
{
    if (!h) return SDB_E_ARG;
    if (!phits || !pool_used || (pool_cap && !pool)) return set_err(h, SDB_E_ARG, "sdb_demod_host_payloads: null pointer");
    *pool_used = 0;
    if (kind < SDB_KIND_MS || kind > SDB_KIND_MN) return set_err(h, SDB_E_ARG, "sdb_demod_host_payloads: bad kind");
    /* a format kernel per pipeline stage (sdb_format.cu): the strings travel back instead of the bit arena (bits may be NULL) */
    PayloadSink sink;
    sink.pool = pool; sink.pool_cap = pool_cap > 0xFFFFFFF0ull ? 0xFFFFFFF0u : (uint32_t)pool_cap; sink.phits = phits;
    const int rc = settle(h, demod_host_impl(h, kind, mc_repaired, msgs, digits, digits_len, n, out, hits, hits_cap, bits, bits_cap, counters, &sink));
    *pool_used = sink.used;
    if (rc != SDB_OK) return rc;
    if (sink.used > sink.pool_cap) return set_err(h, SDB_E_OVERFLOW, "payload pool too small");
    return SDB_OK;
}

/* ---- unit ops ----------------------------------------------------------------------------- */
/* Python's repr(float): the shortest digit string that round-trips, fixed notation for 1e-4 <= |v| < 1e16 */
static void py_float_repr(double v, char *out, size_t cap)
{
    if (v != v) { snprintf(out, cap, "nan"); return; }
    if (v == 0.0) { snprintf(out, cap, "%s", std::signbit(v) ? "-0.0" : "0.0"); return; }
    if (v - v != 0.0) { snprintf(out, cap, "%s", v < 0 ? "-inf" : "inf"); return; }
    char e[40];
    int prec = 1;
    for (; prec <= 17; prec++) {
        snprintf(e, sizeof e, "%.*e", prec - 1, v);
        if (strtod(e, nullptr) == v) break;
    }
    /* e = [-]d[.ddd]e[+-]xx */
    std::string digits;
    const char *p = e;
    const bool neg = *p == '-';
    if (neg) p++;
    for (; *p && *p != 'e'; p++) if (*p != '.') digits.push_back(*p);
    const int ex = atoi(p + 1);
    std::string r = neg ? "-" : "";
    if (ex >= -4 && ex < 16) {
        if (ex >= 0) {
            for (int i = 0; i <= ex; i++) r.push_back(i < (int)digits.size() ? digits[i] : '0');
            r.push_back('.');
            if ((int)digits.size() > ex + 1) r.append(digits, ex + 1, std::string::npos); else r.push_back('0');
        } else {
            r += "0.";
            r.append((size_t)(-ex - 1), '0');
            r += digits;
        }
    } else {
        r.push_back(digits[0]);
        if (digits.size() > 1) { r.push_back('.'); r.append(digits, 1, std::string::npos); }
        char x[8];
        snprintf(x, sizeof x, "e%c%02d", ex < 0 ? '-' : '+', ex < 0 ? -ex : ex);
        r += x;
    }
    snprintf(out, cap, "%s", r.c_str());
}

extern "C" int sdb_format_json(const SdbHandle *h, int kind,
                               const SdbHit *hits, uint32_t nhits, const uint32_t *bits,
                               const char *id_pool, const uint32_t *id_off,
                               const uint8_t *text, const uint32_t *line_off, const SdbLineInfo *info,
                               char *pool, size_t pool_cap, uint64_t *str_off, size_t *pool_used)
{
    if (!h || !str_off || !pool_used || !id_pool || !id_off || (nhits && (!hits || !info || !line_off || !text))) return SDB_E_ARG;
    if (kind != SDB_KIND_MS && kind != SDB_KIND_MU) return SDB_E_ARG;
    const SdbTblHeader *hd = reinterpret_cast<const SdbTblHeader *>(h->blob.data());
    /* payload strings first (same formatter as sdb_format_hits) */
    std::vector<uint64_t> poff(nhits + 1, 0);
    std::vector<char> ppool(64 + 48 * (size_t)nhits);
    size_t pused = 0;
    int rc = sdb_format_hits(h, kind, hits, nhits, bits, ppool.data(), ppool.size(), poff.data(), &pused);
    if (rc == SDB_E_OVERFLOW) { ppool.resize(pused + 16); rc = sdb_format_hits(h, kind, hits, nhits, bits, ppool.data(), ppool.size(), poff.data(), &pused); }
    if (rc != SDB_OK) return rc;
    /* MU: meta.clock is the protocol's clockabs (message_unsynced.py:288) */
    std::vector<std::string> mu_clock(hd->nproto);
    if (kind == SDB_KIND_MU) {
        const SdbPulseProto *tab = reinterpret_cast<const SdbPulseProto *>(h->blob.data() + hd->off_mu);
        char buf[40];
        for (uint32_t i = 0; i < hd->n_mu; i++) { py_float_repr(tab[i].clock, buf, sizeof buf); mu_clock[tab[i].proto] = buf; }
    }
    size_t used = 0;
    auto put = [&](char c) { if (used < pool_cap) pool[used] = c; used++; };
    auto put_str = [&](const char *s) { while (*s) put(*s++); };
    auto put_json_string = [&](const char *s, size_t n) {          /* json.dumps(str), ensure_ascii=True */
        put('"');
        for (size_t i = 0; i < n; i++) {
            const unsigned char c = (unsigned char)s[i];
            char esc[8];
            switch (c) {
            case '"': put_str("\\\""); break;
            case '\\': put_str("\\\\"); break;
            case '\n': put_str("\\n"); break;
            case '\r': put_str("\\r"); break;
            case '\t': put_str("\\t"); break;
            case '\b': put_str("\\b"); break;
            case '\f': put_str("\\f"); break;
            default:
                if (c < 0x20 || c >= 0x7f) { snprintf(esc, sizeof esc, "\\u%04x", c); put_str(esc); }
                else put((char)c);
            }
        }
        put('"');
    };
    char num[48];
    for (uint32_t i = 0; i < nhits; i++) {
        str_off[i] = used;
        const SdbHit &ht = hits[i];
        if (ht.proto >= hd->nproto) return SDB_E_ARG;
        const SdbLineInfo &li = info[ht.msg];
        /* MqttPublisher._message_to_json (signalduino/mqtt.py:228-245): asdict(message) without "raw", json.dumps(indent=4) */
        put_str("{\n    \"protocol_id\": ");
        put_json_string(id_pool + id_off[ht.proto], id_off[ht.proto + 1] - id_off[ht.proto]);
        put_str(",\n    \"payload\": ");
        put_json_string(ppool.data() + poff[i], (size_t)(poff[i + 1] - poff[i]));
        snprintf(num, sizeof num, "%u", (unsigned)ht.nbits);
        put_str(",\n    \"metadata\": {\n        \"bit_length\": "); put_str(num);
        put_str(",\n        \"rssi\": ");
        if (li.has_r) put_json_string(reinterpret_cast<const char *>(text) + line_off[ht.msg] + li.r_off, li.r_len);
        else put_str("null");
        put_str(",\n        \"clock\": ");
        if (kind == SDB_KIND_MS) { snprintf(num, sizeof num, "%d.0", li.clock); put_str(num); }
        else put_str(mu_clock[ht.proto].c_str());
        put_str("\n    }\n}");
    }
    str_off[nhits] = used;
    *pool_used = used;
    return used > pool_cap ? SDB_E_OVERFLOW : SDB_OK;
}

extern "C" int sdb_unit_postdemod(SdbHandle *h, int method, const uint8_t *bits_in, uint32_t n_in,
                                  uint8_t *bits_out, uint32_t out_cap, uint32_t *n_out, int *rcode)
{
    if (!h || !n_out || !rcode || (n_in && !bits_in)) return SDB_E_ARG;
    if (n_in > SDB_MAX_DIGITS || out_cap > 8192) return set_err(h, SDB_E_ARG, "sdb_unit_postdemod: input too long");
    CK(cudaSetDevice(h->device));
    uint8_t *d_in = h->d_unit, *d_out = h->d_unit + 8192;
    int32_t *d_res = reinterpret_cast<int32_t *>(h->d_unit + 24576);
    if (n_in) CK(cudaMemcpyAsync(d_in, bits_in, n_in, cudaMemcpyHostToDevice, h->stream));
    int rc = sdb::launch_unit_postdemod(method, d_in, n_in, d_out, out_cap, d_res, h->stream);
    if (rc != 0) return set_err(h, rc < 0 ? SDB_E_ARG : SDB_E_CUDA, "unit postdemod launch", rc > 0 ? static_cast<cudaError_t>(rc) : cudaSuccess);
    int32_t res[2];
    CK(cudaMemcpyAsync(res, d_res, sizeof res, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    *rcode = res[0];
    uint32_t no = res[1] > 0 ? (uint32_t)res[1] : 0;
    *n_out = no;
    if (no > out_cap) no = out_cap;
    if (no && bits_out) { CK(cudaMemcpy(bits_out, d_out, no, cudaMemcpyDeviceToHost)); }
    return SDB_OK;
}

extern "C" int sdb_unit_mc(SdbHandle *h, uint32_t proto, int method_override, const uint8_t *bits, uint32_t n, int mcbitnum,
                           uint8_t *bits_out, uint32_t out_cap, uint32_t *n_out, int32_t *seg, uint32_t seg_cap,
                           uint32_t *n_seg, int *rcode, int *reason)
{
    if (!h || !n_out || !n_seg || !rcode || !reason || (n && !bits)) return SDB_E_ARG;
    if (n > SDB_MAX_HEX * 4 || out_cap > 4096 || (proto >= h->tab.nproto && proto != 0xFFFFFFFFu)) return set_err(h, SDB_E_ARG, "sdb_unit_mc: bad argument");
    CK(cudaSetDevice(h->device));
    uint8_t *d_in = h->d_unit, *d_out = h->d_unit + 8192;
    int32_t *d_seg = reinterpret_cast<int32_t *>(h->d_unit + 16384);     /* <= 48 segments */
    int32_t *d_res = reinterpret_cast<int32_t *>(h->d_unit + 24576);
    if (n) CK(cudaMemcpyAsync(d_in, bits, n, cudaMemcpyHostToDevice, h->stream));
    int rc = sdb::launch_unit_mc(h->tab, proto, method_override, d_in, (int)n, mcbitnum, d_out, (int)out_cap, d_seg, d_res, h->stream);
    if (rc != 0) return set_err(h, rc < 0 ? SDB_E_ARG : SDB_E_CUDA, "unit mc launch", rc > 0 ? static_cast<cudaError_t>(rc) : cudaSuccess);
    int32_t res[4];
    CK(cudaMemcpyAsync(res, d_res, sizeof res, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    *rcode = res[0]; *reason = res[1];
    uint32_t no = res[2] > 0 ? (uint32_t)res[2] : 0, ns = res[3] > 0 ? (uint32_t)res[3] : 0;
    *n_out = no; *n_seg = ns;
    if (no && bits_out) CK(cudaMemcpy(bits_out, d_out, no < out_cap ? no : out_cap, cudaMemcpyDeviceToHost));
    if (ns && seg) CK(cudaMemcpy(seg, d_seg, sizeof(int32_t) * (ns < seg_cap ? ns : seg_cap), cudaMemcpyDeviceToHost));
    return SDB_OK;
}

extern "C" int sdb_unit_pattern_exists(SdbHandle *h, const void *tpl, size_t tpl_len, const uint16_t *rank, uint32_t n_rank,
                                       const int16_t *tenths, uint32_t pat_ids, uint32_t npat,
                                       const uint8_t *digits, size_t digits_len, uint32_t dlen,
                                       int *found, uint8_t *target_digits, uint32_t target_cap, int *pos)
{
    if (!h || !tpl || tpl_len != sizeof(SdbKeyTpl) || !tenths || !found || !pos || (n_rank && !rank) || (dlen && !digits))
        return set_err(h, SDB_E_ARG, "sdb_unit_pattern_exists: bad argument");
    if (dlen > SDB_MAX_DIGITS || npat > SDB_MAX_SLOTS || digits_len < (((size_t)dlen + 31) / 32) * 16 || n_rank > 65536)
        return set_err(h, SDB_E_ARG, "sdb_unit_pattern_exists: outside the packed domain");
    SdbKeyTpl k;
    memcpy(&k, tpl, sizeof k);
    if (k.len == 0 || k.len > SDB_MAX_TPL || k.nuniq == 0 || k.nuniq > SDB_MAX_UNIQ) return set_err(h, SDB_E_ARG, "sdb_unit_pattern_exists: bad template");
    for (int u = 0; u < k.nuniq; u++)
        if (k.hi[u] < k.lo[u] || (size_t)k.rank_off[u] + (size_t)(k.hi[u] - k.lo[u]) >= n_rank) return set_err(h, SDB_E_ARG, "sdb_unit_pattern_exists: rank table too short");
    CK(cudaSetDevice(h->device));
    /* scratch: [digits + 64 B | rank | tenths | res] */
    const size_t dbytes = (((size_t)dlen + 31) / 32) * 16 + 64, rbytes = ((size_t)n_rank * 2 + 15) & ~(size_t)15;
    uint8_t *buf = nullptr;
    CK(cudaMalloc(&buf, dbytes + rbytes + 64));
    auto done = [&](int code) { cudaFree(buf); return code; };
    cudaStream_t st = h->stream;
    cudaError_t e = cudaMemsetAsync(buf, 0xFF, dbytes, st);
    if (e == cudaSuccess && dlen) e = cudaMemcpyAsync(buf, digits, (((size_t)dlen + 31) / 32) * 16, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && n_rank) e = cudaMemcpyAsync(buf + dbytes, rank, (size_t)n_rank * 2, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(buf + dbytes + rbytes, tenths, 16, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return done(set_err(h, SDB_E_CUDA, "sdb_unit_pattern_exists: copy", e));
    int32_t *d_res = reinterpret_cast<int32_t *>(buf + dbytes + rbytes + 16);
    int rc = sdb_long::launch_unit_pattern(k, reinterpret_cast<const uint16_t *>(buf + dbytes), reinterpret_cast<const int16_t *>(buf + dbytes + rbytes),
                                           pat_ids, (int)npat, buf, (int)dlen, d_res, st);
    if (rc != 0) return done(set_err(h, rc < 0 ? SDB_E_ARG : SDB_E_CUDA, "unit pattern launch", rc > 0 ? static_cast<cudaError_t>(rc) : cudaSuccess));
    int32_t res[4] = {0, 0, 0, 0};
    e = cudaMemcpyAsync(res, d_res, sizeof res, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return done(set_err(h, SDB_E_CUDA, "sdb_unit_pattern_exists: sync", e));
    *found = res[0];
    *pos = res[3];
    const uint64_t tg = ((uint64_t)(uint32_t)res[2] << 32) | (uint32_t)res[1];
    if (target_digits) for (uint32_t i = 0; i < k.len && i < target_cap; i++) target_digits[i] = (uint8_t)((tg >> (4 * i)) & 0xF);
    return done(SDB_OK);
}
