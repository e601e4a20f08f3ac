/*
 * sdb_format.cu — payload strings of the hits (all four message kinds), on the device.
 *
 * What SDProtocols.demodulate() hands back per hit is a string: preamble + hex (or bits) + postamble
 * (message_synced.py:224-231, message_unsynced.py:254-274).  Formatting 27 M hits per 10 M-message corpus on the host cost
 * more than decoding them on the GPU, and every rank of a multi-GPU job would compete for the same host cores, so the
 * strings are produced here and copied back instead of the bit arena.
 *
 * One thread per hit; a CTA of 256 hits lays its strings out contiguously (block-wide exclusive scan of the lengths + one
 * atomicAdd on the pool counter per CTA), stages them in shared memory when they fit and writes them out coalesced.
 * Every string is followed by a NUL; phits[i] (12 bytes: string offset, protocol, bit length, flags) is what travels back
 * per hit.  The pool order follows the CTA order of the atomics, not the hit order — callers index through phits[i].str_off.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sdb200.h"
#include "sdb_fmt.h"
#include "sdb_pulse.h"

namespace sdb {

#define FMT_THREADS 256
#define FMT_STAGE 12288                       /* bytes of shared staging per CTA (typical CTA total: 256 x ~17 = 4.4 KB) */

struct FArgs {
    int kind;
    const SdbHexProto *hx;                    /* MC / MN: one row per protocol (preamble) */
    const SdbHit *hits;
    const uint32_t *bits;
    const SdbPulseProto *rows;
    const uint16_t *row_of_proto;             /* table-order protocol index -> row of this class */
    uint32_t nproto;
    const uint32_t *range;                    /* [0] first hit to format, [1] = SdbCounters.hits: one past the last */
    const SdbCounters *ctr;                   /* an overflowed arena holds unwritten records: nothing is formatted then */
    uint32_t hits_cap, bits_cap;
    char *pool;
    uint32_t pool_cap;
    SdbPayloadHit *phits;
    uint32_t *used;                           /* pool bytes handed out so far (may exceed pool_cap: overflow) */
};

__global__ void __launch_bounds__(FMT_THREADS) format_kernel(FArgs A)
{
    __shared__ char stage[FMT_STAGE];
    __shared__ uint32_t wsum[FMT_THREADS / 32];
    __shared__ uint32_t base_sh;
    const uint32_t h0 = A.range[0], h1 = A.range[1];
    if (A.ctr->hits > A.hits_cap || A.ctr->words > A.bits_cap) return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint32_t b0 = h0 + blockIdx.x * FMT_THREADS; b0 < h1; b0 += gridDim.x * FMT_THREADS) {
        const uint32_t i = b0 + threadIdx.x;
        SdbHit ht;
        const SdbPulseProto *pp = nullptr;
        const uint32_t *w = nullptr;
        uint32_t len = 0;
        if (i < h1) {
            ht = A.hits[i];
            if (ht.proto < A.nproto) {
                if (A.kind <= SDB_KIND_MU) {
                    pp = &A.rows[A.row_of_proto[ht.proto]];
                    w = A.bits + ht.bits_off;
                    len = sdb_fmt_pulse(pp, ht, w, nullptr) + 1;        /* + NUL */
                } else len = sdb_fmt_hexkind(A.kind, A.hx, A.hits, i, h1, A.bits, nullptr) + 1;
            }
        }
        /* exclusive scan of the lengths over the CTA */
        uint32_t incl = len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
        if (lane == 31) wsum[wid] = incl;
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int k = 0; k < FMT_THREADS / 32; k++) { const uint32_t s = wsum[k]; if (k < wid) before += s; total += s; }
        const uint32_t excl = before + incl - len;
        if (threadIdx.x == 0) base_sh = atomicAdd(A.used, total);
        __syncthreads();
        const uint32_t base = base_sh;
        const bool fits = base + total <= A.pool_cap && base + total >= base;
        if (i < h1) {
            SdbPayloadHit ph;
            ph.str_off = base + excl; ph.proto = ht.proto; ph.nbits = ht.nbits; ph.aux = ht.aux; ph.flags = ht.flags; ph.rsv = 0;
            A.phits[i] = ph;
        }
        if (fits && total) {
            const bool staged = total <= FMT_STAGE;
            if (len) {
                char *dst = staged ? stage + excl : A.pool + base + excl;
                const uint32_t n = A.kind <= SDB_KIND_MU ? sdb_fmt_pulse(pp, ht, w, dst) : sdb_fmt_hexkind(A.kind, A.hx, A.hits, i, h1, A.bits, dst);
                dst[n] = 0;
            }
            if (staged) {
                __syncthreads();
                char *out = A.pool + base;
                /* coalesced copy: bytes up to the first 16-byte boundary of the destination, then 16-byte words */
                const uint32_t head = min(total, (uint32_t)((16 - ((uintptr_t)out & 15)) & 15));
                if (threadIdx.x < head) out[threadIdx.x] = stage[threadIdx.x];
                const uint32_t nvec = (total - head) >> 4;
                for (uint32_t v = threadIdx.x; v < nvec; v += FMT_THREADS) {
                    uint4 x;
                    const char *s = stage + head + 16 * v;           /* the staging side is unaligned relative to the destination */
                    uint32_t q[4];
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        q[k] = (uint32_t)(uint8_t)s[4 * k] | ((uint32_t)(uint8_t)s[4 * k + 1] << 8) | ((uint32_t)(uint8_t)s[4 * k + 2] << 16) | ((uint32_t)(uint8_t)s[4 * k + 3] << 24);
                    x.x = q[0]; x.y = q[1]; x.z = q[2]; x.w = q[3];
                    *reinterpret_cast<uint4 *>(out + head + 16 * v) = x;
                }
                const uint32_t tail0 = head + 16 * nvec;
                if (tail0 + threadIdx.x < total) out[tail0 + threadIdx.x] = stage[tail0 + threadIdx.x];
            }
        }
        __syncthreads();
    }
}

/* after the format kernel of a stage: the next stage starts where this one ended */
__global__ void format_advance_kernel(uint32_t *range) { range[0] = range[1]; }

int launch_format(int kind, const SdbHit *d_hits, const uint32_t *d_bits, const SdbPulseProto *rows, const uint16_t *row_of_proto,
                  const SdbHexProto *hx, uint32_t nproto,
                  uint32_t *d_range, const SdbCounters *d_ctr, uint32_t hits_cap, uint32_t bits_cap, char *d_pool, uint32_t pool_cap,
                  SdbPayloadHit *d_phits, uint32_t *d_used, int grid, cudaStream_t stream)
{
    /* range[1] = the hit counter now (device-side copy, stream-ordered after the decode kernels of this stage) */
    cudaError_t e = cudaMemcpyAsync(d_range + 1, &d_ctr->hits, sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream);
    if (e != cudaSuccess) return (int)e;
    FArgs A;
    A.kind = kind; A.hx = hx;
    A.hits = d_hits; A.bits = d_bits; A.rows = rows; A.row_of_proto = row_of_proto; A.nproto = nproto; A.range = d_range;
    A.pool = d_pool; A.pool_cap = pool_cap; A.phits = d_phits; A.used = d_used;
    A.ctr = d_ctr; A.hits_cap = hits_cap; A.bits_cap = bits_cap;
    format_kernel<<<grid, FMT_THREADS, 0, stream>>>(A);
    format_advance_kernel<<<1, 1, 0, stream>>>(d_range);
    return (int)cudaGetLastError();
}

}  // namespace sdb
