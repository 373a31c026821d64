// K1, streaming form -- the LRGB fit as a persistent, warp-specialised sm_100a kernel.
//
//   producer warp : feeds a shared-memory ring with 1-D bulk copies (cp.async.bulk, SASS
//                   UBLKCP): each copy is one light's 3072-byte run of 1024 consecutive pixels;
//                   L lights per stage (one producer lane each), S stages, completion counted
//                   in bytes on an mbarrier per stage.  Default L = 6, S = 4 (72 KB ring).
//   consumer warps: 8 warps x 32 lanes x 4 pixels = the 1024-pixel tile.  Per light a lane
//                   reads its 12 bytes as 3 words (lane stride 3 words: conflict free), turns
//                   the four luma bytes into floats with PRMT + FADD, and issues 12 FFMA2 (packed
//                   pairs of coefficients) against the pseudo-inverse column broadcast from shared memory.  All 12
//                   bytes are also summed for ptm_cbcr_avg in two 32-bit accumulators per word:
//                   `ev += w` (no unpacking at all) and `od += bytes 1,3 as 16-bit lanes`;
//                   the even-byte sums are recovered at the end as ev - (S1 << 8) - (S3 << 24).
//   epilogue      : integer mean, JFIF colour conversion, 256 / Y normalisation, running
//                   extrema, and a warp-private staging area laid out like the output so the
//                   96 + 12 bytes per lane leave as fully coalesced 16-byte stores.
//
// HBM latency is covered by the ring, not by occupancy (two CTAs per SM), so the consumers keep
// their 24 accumulators in registers without spilling.  Measured on B200 (24 MP x 60 lights):
// a read-only ring of this shape streams the stack at 7.2 TB/s; with K1's own 27 B/px of output
// mixed in the floor is 0.77 ms (tools/ubench_read.cu), the kernel runs in 0.83 ms.
//
// Reference arithmetic fused here: rti-builder/ptmlib.c:692-725 (fit), :762-802 (average),
// rti-builder/ptm-encoder.c:327-353 (colour + normalise), rti-builder/ptmlib.c:826-847 (extrema).
#include "ptm_device.cuh"
#include "ptm_kernels.h"
#include "ptm_tma.cuh"

#include <cstdlib>

namespace ptm {

constexpr int kConsumerWarps = 8;
constexpr int kConsumers = kConsumerWarps * 32;          // 256
constexpr int kTmaThreads = kConsumers + 32;             // + producer warp
constexpr int kTilePx = kConsumers * 4;                  // 1024
constexpr int kTileBytes = kTilePx * 3;                  // 3072 per light
constexpr int kWarpStageBytes = 128 * 24 + 128 * 3;      // a warp's 128 pixels: float coefficients + colour bytes

// dynamic shared memory: [ring: S stages x L lights x 3072 B][2 S mbarriers][matrix: n x 32 B]
// W = consumer warps: 8 (1024-pixel tiles, two CTAs per SM) or 4 (512-pixel tiles, four CTAs per SM: the same
// bytes in flight per SM, half the finish spread -- for slabs with few tiles per CTA, see launch_fit_lrgb_tma)
template <int L, int S, int W = kConsumerWarps>
struct TmaSmem {
    static constexpr int tile_px = W * 128;
    static constexpr int tile_bytes = tile_px * 3;
    static constexpr size_t stage_bytes = (size_t) L * tile_bytes;
    static constexpr size_t bars = (size_t) S * stage_bytes;
    static constexpr size_t staging = bars + 2 * S * sizeof(uint64_t);        // W warps x (3072 + 384) B
    static constexpr size_t matrix = staging + (size_t) W * kWarpStageBytes;
    static constexpr size_t total(unsigned n_lights) { return matrix + (size_t) n_lights * 32; }
};

// One light's contribution to a lane's four pixels.
//   word0 = Y0 Cb0 Cr0 Y1 | word1 = Cb1 Cr1 Y2 Cb2 | word2 = Cr2 Y3 Cb3 Cr3
template <bool kSums = true, bool kRgbIn = false>
__device__ __forceinline__ void consume_light(const uint32_t *s32, const float4 *m, float (&acc)[4][PTM_NCOEF],
                                              unsigned (&ev)[3], unsigned (&od)[3]) {
    unsigned w0 = s32[0], w1 = s32[1], w2 = s32[2];
    if (kRgbIn)
        rgb_words_to_ycc(w0, w1, w2);   // RGB-input stacks: per-sample integer JFIF conversion first
    const float4 ma = m[0], mb = m[1];
    const float y[4] = {byte_to_float(w0, PTM_SEL_B0), byte_to_float(w0, PTM_SEL_B3),
                        byte_to_float(w1, PTM_SEL_B2), byte_to_float(w2, PTM_SEL_B1)};
    // Packed form (fma.rn.f32x2, SASS FFMA2): coefficient pairs (0,1) (2,3) (4,5) of a pixel share one
    // instruction with the sample broadcast to both halves.  Same FLOPs, same results bit for bit, half
    // the FMA instructions: FFMA2 issues at half rate, so it buys no issue cycles, but it cuts SM power
    // -- under the 1 kW cap the sustained fit went from 1.018 to 0.928 ms per 24 MP.
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        unsigned long long c01, c23, c45, m01, m23, m45, yy;
        asm("mov.b64 %0, {%1, %2};" : "=l"(c01) : "f"(acc[p][0]), "f"(acc[p][1]));
        asm("mov.b64 %0, {%1, %2};" : "=l"(c23) : "f"(acc[p][2]), "f"(acc[p][3]));
        asm("mov.b64 %0, {%1, %2};" : "=l"(c45) : "f"(acc[p][4]), "f"(acc[p][5]));
        asm("mov.b64 %0, {%1, %2};" : "=l"(m01) : "f"(ma.x), "f"(ma.y));
        asm("mov.b64 %0, {%1, %2};" : "=l"(m23) : "f"(ma.z), "f"(ma.w));
        asm("mov.b64 %0, {%1, %2};" : "=l"(m45) : "f"(mb.x), "f"(mb.y));
        asm("mov.b64 %0, {%1, %1};" : "=l"(yy) : "f"(y[p]));
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c01) : "l"(yy), "l"(m01));
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c23) : "l"(yy), "l"(m23));
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c45) : "l"(yy), "l"(m45));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[p][0]), "=f"(acc[p][1]) : "l"(c01));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[p][2]), "=f"(acc[p][3]) : "l"(c23));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[p][4]), "=f"(acc[p][5]) : "l"(c45));
    }
    if (kSums) {
        ev[0] += w0;
        ev[1] += w1;
        ev[2] += w2;
        od[0] += __byte_perm(w0, 0, 0x4341);
        od[1] += __byte_perm(w1, 0, 0x4341);
        od[2] += __byte_perm(w2, 0, 0x4341);
    }
}

// Byte sums of TWO lights at once: three-operand adds (IADD3) halve the integer add count.
__device__ __forceinline__ void sum_pair(const uint32_t *a, const uint32_t *b, unsigned (&ev)[3], unsigned (&od)[3]) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const unsigned wa = a[j], wb = b[j];
        ev[j] = ev[j] + wa + wb;
        od[j] = od[j] + __byte_perm(wa, 0, 0x4341) + __byte_perm(wb, 0, 0x4341);
    }
}

// kDebug (only instantiated in -DPTM_TUNING builds): 1 = consumers skip the arithmetic
// (ring delivery ceiling), 2 = consumers skip the waits (arithmetic ceiling).
// kFull: the light count is a multiple of L, so every stage is complete (no per-stage count).
__device__ __forceinline__ unsigned tile_px_of(unsigned tile, unsigned n_tiles, unsigned tail_px, unsigned full_px) {
    return (tail_px && tile + 1 == n_tiles) ? tail_px : full_px;
}

__device__ __forceinline__ void st_hint(float4 *p, float4 v, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w), "l"(policy)
                 : "memory");
}

#ifdef PTM_TUNING
// per-CTA timeline of the last launch: entry, first stage consumed, last tile stored, exit (globaltimer ns)
__device__ unsigned long long g_cta_times[4 * 1024];
__device__ __forceinline__ void cta_stamp(int slot) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (blockIdx.x < 1024)
        g_cta_times[4 * blockIdx.x + slot] = t;
}
#define PTM_STAMP(slot) cta_stamp(slot)
#else
#define PTM_STAMP(slot)
#endif

template <int L, int S, bool kWriteCoeffs, bool kFull, int kDebug = 0, bool kRgbIn = false, int W = kConsumerWarps>
__global__ void __launch_bounds__(W * 32 + 32, 16 / W)
fit_lrgb_u8_tma_kernel(FitArgs a, unsigned n_tiles, unsigned tail_px) {
    using Smem = TmaSmem<L, S, W>;
    constexpr int kConsumerWarps = W, kConsumers = W * 32, kTilePx = Smem::tile_px, kTileBytes = Smem::tile_bytes;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ int keys_s[2 * PTM_NCOEF];
    // Tile schedule.  The first tile of a block is blockIdx.x; with a.tile_counter the following ones are
    // drawn from a device-wide counter by the producer (blocks on faster SMs take more tiles: the spread
    // between the first and the last block to finish was 10 % of the kernel with a static stride) and
    // announced to the consumers through tile_pub[tile parity], published by the mbarrier arrive of the
    // tile's first stage.  A tile index >= n_tiles tells the consumers to stop.
    __shared__ unsigned tile_pub[2];
    uint8_t *ring = smem;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + Smem::bars);
    uint64_t *empty = full + S;
    float4 *m_s = reinterpret_cast<float4 *>(smem + Smem::matrix);

    const unsigned n = a.n_lights;
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kConsumerWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();
    // In the fused encode this kernel is itself launched programmatically dependent on the previous
    // K2 (its blocks become resident while K2's last blocks drain): nothing is read or written before
    // that grid has completed.  A no-op for an ordinary launch.
    // Exception: when the caller guarantees that the stack is not written by earlier work on the stream
    // (a.stack_stable, ptmb_set_stack_stable), the PRODUCER warp starts its bulk copies right away -- the ring
    // fills while the previous grid drains, which hides the 2-4 us ramp of the first stages.  The producer only
    // reads the stack and draws tile numbers; it joins the dependency before it touches anything else (below).
    const bool early_producer = a.stack_stable && warp == kConsumerWarps;
    if (!early_producer)
        asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x == 0)
        PTM_STAMP(0);
    // a programmatically dependent K2 (the fused encode) may be scheduled as soon as SM resources free up
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const unsigned n_chunks = (n + L - 1) / L;
    Extrema ext;
    ext.init();

    if (warp != kConsumerWarps) {
        // consumers fetch the pinv table while the producer already streams the first stages
        for (unsigned i = threadIdx.x; i < n; i += kConsumers) {
            const float *M = a.M;
            m_s[2 * i] = make_float4(M[i], M[n + i], M[2 * n + i], M[3 * n + i]);
            m_s[2 * i + 1] = make_float4(M[4 * n + i], M[5 * n + i], 0.f, 0.f);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory");
    }

    if (warp == kConsumerWarps) {
        // ------------------------------------------------------------ producer
        // lane 0 owns the barrier protocol; lanes 0..L-1 each issue one light's copy of the stage
        const uint64_t policy = l2_policy_evict_first();
        unsigned s = 0, ph = 0, slot = 0;
        unsigned tile = blockIdx.x;
        for (;;) {
            if (lane == 0) {
                mbar_wait(&empty[s], ph ^ 1);          // also orders this write after the readers of the slot
                tile_pub[slot] = tile;
            }
            if (tile >= n_tiles) {
                if (lane == 0)
                    mbar_arrive(&full[s]);             // wake the consumers: nothing more to do
                break;
            }
            // draw the next tile now, its latency hides behind this tile's copies
            unsigned next = tile + gridDim.x;
            if (a.tile_counter) {
                if (lane == 0)
                    next = gridDim.x + atomicAdd(a.tile_counter, 1u);
                next = __shfl_sync(0xFFFFFFFFu, next, 0);
            }
            const uint8_t *src = a.stack + (size_t) tile * kTileBytes + (size_t) lane * a.image_stride;
            // the last tile may be partial (tail_px pixels, a multiple of 16 => bytes a multiple of 16)
            const unsigned bytes = (tail_px && tile + 1 == n_tiles) ? tail_px * 3 : (unsigned) kTileBytes;
            for (unsigned c = 0; c < n_chunks; ++c) {
                const unsigned l0 = c * L;
                const unsigned cnt = kFull ? (unsigned) L : min((unsigned) L, n - l0);
                if (lane == 0) {
                    if (c != 0)
                        mbar_wait(&empty[s], ph ^ 1);
                    mbar_arrive_expect_tx(&full[s], cnt * bytes);
                }
                __syncwarp();
                if (lane < cnt)
                    bulk_g2s(ring + (size_t) s * Smem::stage_bytes + (size_t) lane * kTileBytes,
                             src + (size_t) l0 * a.image_stride, bytes, &full[s], policy);
                if (++s == S) {
                    s = 0;
                    ph ^= 1;
                }
            }
            tile = next;
            slot ^= 1;
        }
        if (early_producer)
            asm volatile("griddepcontrol.wait;" ::: "memory");   // before the key hand-off below
    } else {
        // ----------------------------------------------------------- consumers
        const unsigned t = threadIdx.x;  // 0..255, owns pixels 4t..4t+3 of the tile
        unsigned s = 0, ph = 0, slot = 0;
        uint64_t st_policy = 0;
        if (a.store_hint == 1)
            st_policy = l2_policy_evict_first();
        else if (a.store_hint == 2)
            asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(st_policy));
        for (bool first = true;; first = false, slot ^= 1) {
            unsigned tile = 0;
            float acc[4][PTM_NCOEF];
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int k = 0; k < PTM_NCOEF; ++k)
                    acc[p][k] = 0.f;
            unsigned ev[3] = {0, 0, 0}, od[3] = {0, 0, 0};

            bool stop = false;
            for (unsigned c = 0; c < n_chunks; ++c) {
                mbar_wait(&full[s], ph);
                if (c == 0) {
                    tile = *reinterpret_cast<volatile unsigned *>(&tile_pub[slot]);
                    if (tile >= n_tiles) {
                        stop = true;
                        break;
                    }
                }
                if (threadIdx.x == 0 && first && c == 0)
                    PTM_STAMP(1);
                const uint32_t *buf = reinterpret_cast<const uint32_t *>(ring + (size_t) s * Smem::stage_bytes) + 3 * t;
                const unsigned l0 = c * L;
                const unsigned cnt = kFull ? (unsigned) L : min((unsigned) L, n - l0);
                if (kDebug == 1) {
                    ev[0] += buf[0];
                } else if (kRgbIn) {
                    for (unsigned l = 0; l < cnt; ++l)
                        consume_light<true, true>(buf + l * (kTileBytes / 4), m_s + 2 * (l0 + l), acc, ev, od);
                } else if (cnt == L) {
                    // straight-line body: all loads of the stage can be hoisted; byte sums go pairwise
#pragma unroll
                    for (unsigned l = 0; l + 1 < L; l += 2) {
                        consume_light<false>(buf + l * (kTileBytes / 4), m_s + 2 * (l0 + l), acc, ev, od);
                        consume_light<false>(buf + (l + 1) * (kTileBytes / 4), m_s + 2 * (l0 + l + 1), acc, ev, od);
                        sum_pair(buf + l * (kTileBytes / 4), buf + (l + 1) * (kTileBytes / 4), ev, od);
                    }
                    if (L & 1)
                        consume_light(buf + (L - 1) * (kTileBytes / 4), m_s + 2 * (l0 + L - 1), acc, ev, od);
                } else {
                    for (unsigned l = 0; l < cnt; ++l)
                        consume_light(buf + l * (kTileBytes / 4), m_s + 2 * (l0 + l), acc, ev, od);
                }
                __syncwarp();
                if (lane == 0)
                    mbar_arrive(&empty[s]);
                if (++s == S) {
                    s = 0;
                    ph ^= 1;
                }
            }

            if (stop)
                break;
            // per-byte sums: odd bytes from the 16-bit lanes, even bytes from the plain word sum
            unsigned sum[12];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const unsigned s1 = od[j] & 0xFFFFu, s3 = od[j] >> 16;
                const unsigned e = ev[j] - (s1 << 8) - (s3 << 24);
                sum[4 * j + 0] = e & 0xFFFFu;
                sum[4 * j + 1] = s1;
                sum[4 * j + 2] = e >> 16;
                sum[4 * j + 3] = s3;
            }
            unsigned out_b[12];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const unsigned ya = div_magic(sum[3 * p], a.magic);
                const unsigned cba = div_magic(sum[3 * p + 1], a.magic);
                const unsigned cra = div_magic(sum[3 * p + 2], a.magic);
                lrgb_or_lum_finish(a.mode & kFitLum, ya, cba, cra, acc[p], out_b[3 * p], out_b[3 * p + 1], out_b[3 * p + 2]);
            }
            if (4 * t < tile_px_of(tile, n_tiles, tail_px, kTilePx))   // lanes past a partial tile hold stale bytes
                ext.add4(acc);

            // The lane's 96 bytes of coefficients and 12 of colour go through a warp-private
            // staging area laid out exactly like the output, so that they leave as fully
            // coalesced 16-byte stores (512 contiguous bytes per warp instruction).
            uint8_t *stg = smem + Smem::staging + (size_t) warp * kWarpStageBytes;
            const size_t wpx = (size_t) tile * kTilePx + (size_t) warp * 128;  // the warp's first pixel
            // pixels of this warp that exist (all 128 except in a partial last tile; multiple of 16)
            const unsigned tile_px = (tail_px && tile + 1 == n_tiles) ? tail_px : (unsigned) kTilePx;
            const unsigned valid = min(128u, tile_px > warp * 128 ? tile_px - warp * 128 : 0u);
            if (kWriteCoeffs) {
                float4 *row = reinterpret_cast<float4 *>(stg) + lane * 6;
                const float *f = &acc[0][0];
#pragma unroll
                for (int j = 0; j < 6; ++j)
                    row[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            }
            uint32_t *rgbw = reinterpret_cast<uint32_t *>(stg + 128 * 24);
            rgbw[3 * lane + 0] = out_b[0] | (out_b[1] << 8) | (out_b[2] << 16) | (out_b[3] << 24);
            rgbw[3 * lane + 1] = out_b[4] | (out_b[5] << 8) | (out_b[6] << 16) | (out_b[7] << 24);
            rgbw[3 * lane + 2] = out_b[8] | (out_b[9] << 8) | (out_b[10] << 16) | (out_b[11] << 24);
            __syncwarp();
            if (kWriteCoeffs) {
                float4 *dst = reinterpret_cast<float4 *>(a.coeffs + wpx * PTM_NCOEF);
                const float4 *src = reinterpret_cast<const float4 *>(stg);
#pragma unroll
                for (int r = 0; r < 6; ++r)
                    if (lane + 32 * r < valid * 6 / 4) {
                        if (a.store_hint)
                            st_hint(dst + lane + 32 * r, src[lane + 32 * r], st_policy);
                        else
                            __stcg(dst + lane + 32 * r, src[lane + 32 * r]);
                    }
            }
            if (lane < valid * 3 / 16)
                __stcg(reinterpret_cast<uint4 *>(a.rgb + wpx * 3) + lane,
                       reinterpret_cast<const uint4 *>(rgbw)[lane]);
            __syncwarp();
        }
    }
    if (threadIdx.x == 0)
        PTM_STAMP(2);
    ext.flush(a.keys, keys_s);
    if (a.handoff.ticket)
        keys_handoff<KeyHandoff, XchgLayout>(a.handoff, a.keys, kMaxRanks, kSlotInts);
    if (threadIdx.x == 0)
        PTM_STAMP(3);
}

#ifdef PTM_TUNING
extern "C" int ptmb_tuning_cta_times(unsigned long long *out, int n) {
    return (int) cudaMemcpyFromSymbol(out, g_cta_times, sizeof(unsigned long long) * (size_t) n);
}
#endif

template <int L, int S, int kDebug = 0, int W = kConsumerWarps>
static cudaError_t launch_variant(const FitArgs &a, size_t n_tiles, unsigned tail_px, int sm_count, cudaStream_t st) {
    using Smem = TmaSmem<L, S, W>;
    constexpr int kTmaThreads = W * 32 + 32;
    static_assert(L <= 32, "one producer lane per light of a stage");
    // opt in to > 48 KB of dynamic shared memory and look up residency, once per device
    static int per_sm_cache[64][5];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess)
        return e;
    if (dev < 0 || dev >= 64)
        return cudaErrorInvalidDevice;
    const int wc = a.coeffs != nullptr;
    const bool full = a.n_lights % L == 0;
    auto kern = wc ? (full ? fit_lrgb_u8_tma_kernel<L, S, true, true, kDebug, false, W>
                                : fit_lrgb_u8_tma_kernel<L, S, true, false, kDebug, false, W>)
                   : (full ? fit_lrgb_u8_tma_kernel<L, S, false, true, kDebug, false, W>
                           : fit_lrgb_u8_tma_kernel<L, S, false, false, kDebug, false, W>);
    int slot = wc * 2 + (full ? 1 : 0);
    if (a.mode & kFitRgbIn) {   // RGB-input stacks: the instantiation with the per-sample colour conversion
        kern = fit_lrgb_u8_tma_kernel<L, S, true, false, 0, true, W>;
        slot = 4;
        if (!wc)
            return cudaErrorNotSupported;
    }
    if (per_sm_cache[dev][slot] == 0) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) Smem::total(257));
        if (e != cudaSuccess)
            return e;
        int nb = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kTmaThreads, Smem::total(257));
        if (e != cudaSuccess)
            return e;
        if (nb < 1)
            return cudaErrorLaunchOutOfResources;
        per_sm_cache[dev][slot] = nb;
    }
    size_t grid = (size_t) sm_count * per_sm_cache[dev][slot];
    if (grid > n_tiles)
        grid = n_tiles;
    // Invariant: the tile counter is 0 between launches.  The hand-off block of a fused launch resets it on
    // the device; any other launch is followed by a stream-ordered clear (below).
    if (a.handoff.ticket) {
        // fused encode: K1 -> K2 -> K1 -> ... each launched programmatically dependent on its predecessor
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned) grid);
        cfg.blockDim = dim3(kTmaThreads);
        cfg.dynamicSmemBytes = Smem::total(a.n_lights);
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, kern, a, (unsigned) n_tiles, tail_px);
    }
    kern<<<(unsigned) grid, kTmaThreads, Smem::total(a.n_lights), st>>>(a, (unsigned) n_tiles, tail_px);
    e = cudaGetLastError();
    if (e == cudaSuccess && a.tile_counter)
        e = cudaMemsetAsync(a.tile_counter, 0, sizeof(unsigned), st);
    return e;
}

// ---------------------------------------------------------------------------
// RGB formats: the same ring, but every byte of the stack is a sample of its own fit
// (rti-builder/ptm-encoder.c:272-284): a lane keeps 4 pixels x 3 channels x 6 = 72 accumulators,
// as 36 packed pairs updated with FFMA2 (fma.rn.f32x2, the sample broadcast to both halves and a
// coefficient pair of the pinv column as the other operand).  Output: three planes [pixels][6],
// only plane 0 feeds the extrema (rti-builder/ptmlib.c:831).
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ffma2_bcast(float s, unsigned long long m, unsigned long long c) {
    unsigned long long ss, d;
    asm("mov.b64 %0, {%1, %1};" : "=l"(ss) : "f"(s));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(ss), "l"(m), "l"(c));
    return d;
}

__device__ __forceinline__ void consume_light_rgb(const uint32_t *s32, const float4 *m,
                                                  unsigned long long (&acc)[12][3]) {
    const unsigned w[3] = {s32[0], s32[1], s32[2]};
    const float4 ma = m[0], mb = m[1];
    unsigned long long m01, m23, m45;
    asm("mov.b64 %0, {%1, %2};" : "=l"(m01) : "f"(ma.x), "f"(ma.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(m23) : "f"(ma.z), "f"(ma.w));
    asm("mov.b64 %0, {%1, %2};" : "=l"(m45) : "f"(mb.x), "f"(mb.y));
#pragma unroll
    for (int j = 0; j < 12; ++j) {   // byte j of the 12 = pixel j / 3, channel j % 3
        const float v = byte_to_float(w[j >> 2], 0x7650u + (j & 3));
        acc[j][0] = ffma2_bcast(v, m01, acc[j][0]);
        acc[j][1] = ffma2_bcast(v, m23, acc[j][1]);
        acc[j][2] = ffma2_bcast(v, m45, acc[j][2]);
    }
}

constexpr int kRgbWarps = 7;                             // 7 consumer warps + 1 producer = 256 threads:
constexpr int kRgbThreads = kRgbWarps * 32 + 32;         //   two CTAs per SM get 128 registers per thread
constexpr int kRgbTilePx = kRgbWarps * 128;              // 896 pixels
constexpr int kRgbTileBytes = kRgbTilePx * 3;            // 2688 bytes per light (a multiple of 16)

template <int L, int S>
struct RgbSmem {
    static constexpr size_t stage_bytes = (size_t) L * kRgbTileBytes;
    static constexpr size_t bars = (size_t) S * stage_bytes;
    static constexpr size_t staging = bars + 2 * S * sizeof(uint64_t);
    static constexpr size_t matrix = staging + (size_t) kRgbWarps * kWarpStageBytes;
    static constexpr size_t total(unsigned n_lights) { return matrix + (size_t) n_lights * 32; }
};

template <int L, int S, bool kFull>
__global__ void __launch_bounds__(kRgbThreads, 2)
fit_rgb_u8_tma_kernel(FitArgs a, unsigned n_tiles, unsigned tail_px) {
    using Smem = RgbSmem<L, S>;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ int keys_s[2 * PTM_NCOEF];
    uint8_t *ring = smem;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + Smem::bars);
    uint64_t *empty = full + S;
    float4 *m_s = reinterpret_cast<float4 *>(smem + Smem::matrix);

    const unsigned n = a.n_lights;
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
        const float *M = a.M;
        m_s[2 * i] = make_float4(M[i], M[n + i], M[2 * n + i], M[3 * n + i]);
        m_s[2 * i + 1] = make_float4(M[4 * n + i], M[5 * n + i], 0.f, 0.f);
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kRgbWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned n_chunks = (n + L - 1) / L;
    Extrema ext;
    ext.init();

    if (warp == kRgbWarps) {
        const uint64_t policy = l2_policy_evict_first();
        unsigned s = 0, ph = 0;
        for (unsigned tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const uint8_t *src = a.stack + (size_t) tile * kRgbTileBytes + (size_t) lane * a.image_stride;
            const unsigned bytes = (tail_px && tile + 1 == n_tiles) ? tail_px * 3 : (unsigned) kRgbTileBytes;
            for (unsigned c = 0; c < n_chunks; ++c) {
                const unsigned l0 = c * L;
                const unsigned cnt = kFull ? (unsigned) L : min((unsigned) L, n - l0);
                if (lane == 0) {
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_arrive_expect_tx(&full[s], cnt * bytes);
                }
                __syncwarp();
                if (lane < cnt)
                    bulk_g2s(ring + (size_t) s * Smem::stage_bytes + (size_t) lane * kRgbTileBytes,
                             src + (size_t) l0 * a.image_stride, bytes, &full[s], policy);
                if (++s == S) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
    } else {
        const unsigned t = threadIdx.x;
        unsigned s = 0, ph = 0;
        for (unsigned tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            unsigned long long acc[12][3];
#pragma unroll
            for (int j = 0; j < 12; ++j)
                acc[j][0] = acc[j][1] = acc[j][2] = 0ull;   // two +0.0f

            for (unsigned c = 0; c < n_chunks; ++c) {
                mbar_wait(&full[s], ph);
                const uint32_t *buf = reinterpret_cast<const uint32_t *>(ring + (size_t) s * Smem::stage_bytes) + 3 * t;
                const unsigned l0 = c * L;
                const unsigned cnt = kFull ? (unsigned) L : min((unsigned) L, n - l0);
                if (cnt == L) {
#pragma unroll
                    for (unsigned l = 0; l < L; ++l)
                        consume_light_rgb(buf + l * (kRgbTileBytes / 4), m_s + 2 * (l0 + l), acc);
                } else {
                    for (unsigned l = 0; l < cnt; ++l)
                        consume_light_rgb(buf + l * (kRgbTileBytes / 4), m_s + 2 * (l0 + l), acc);
                }
                __syncwarp();
                if (lane == 0)
                    mbar_arrive(&empty[s]);
                if (++s == S) {
                    s = 0;
                    ph ^= 1;
                }
            }

            const unsigned tile_px = (tail_px && tile + 1 == n_tiles) ? tail_px : (unsigned) kRgbTilePx;
            const unsigned valid = min(128u, tile_px > warp * 128 ? tile_px - warp * 128 : 0u);
            const size_t wpx = (size_t) tile * kRgbTilePx + (size_t) warp * 128;
            uint8_t *stg = smem + Smem::staging + (size_t) warp * kWarpStageBytes;
            if (4 * t < tile_px) {
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    float c0[PTM_NCOEF];
#pragma unroll
                    for (int h = 0; h < 3; ++h) {
                        c0[2 * h] = __uint_as_float((unsigned) (acc[p * 3][h] & 0xFFFFFFFFull));
                        c0[2 * h + 1] = __uint_as_float((unsigned) (acc[p * 3][h] >> 32));
                    }
                    ext.add(c0);
                }
            }
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                // lane's 4 pixels of this channel: 24 floats = 12 packed pairs, laid out like the plane
                unsigned long long *row = reinterpret_cast<unsigned long long *>(stg) + lane * 12;
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int h = 0; h < 3; ++h)
                        row[p * 3 + h] = acc[p * 3 + ch][h];
                __syncwarp();
                float4 *dst = reinterpret_cast<float4 *>(a.coeffs + (size_t) ch * a.plane_stride + wpx * PTM_NCOEF);
                const float4 *src = reinterpret_cast<const float4 *>(stg);
#pragma unroll
                for (int r = 0; r < 6; ++r)
                    if (lane + 32 * r < valid * 6 / 4)
                        __stcg(dst + lane + 32 * r, src[lane + 32 * r]);
                __syncwarp();
            }
        }
    }
    ext.flush(a.keys, keys_s);
}

template <int L, int S>
static cudaError_t launch_rgb_variant(const FitArgs &a, size_t n_tiles, unsigned tail_px, int sm_count,
                                      cudaStream_t st) {
    using Smem = RgbSmem<L, S>;
    static int per_sm_cache[64][2];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess)
        return e;
    if (dev < 0 || dev >= 64)
        return cudaErrorInvalidDevice;
    const bool full = a.n_lights % L == 0;
    auto kern = full ? fit_rgb_u8_tma_kernel<L, S, true> : fit_rgb_u8_tma_kernel<L, S, false>;
    if (per_sm_cache[dev][full] == 0) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) Smem::total(257));
        if (e != cudaSuccess)
            return e;
        int nb = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kRgbThreads, Smem::total(257));
        if (e != cudaSuccess)
            return e;
        if (nb < 1)
            return cudaErrorLaunchOutOfResources;
        per_sm_cache[dev][full] = nb;
    }
    size_t grid = (size_t) sm_count * per_sm_cache[dev][full];
    if (grid > n_tiles)
        grid = n_tiles;
    kern<<<(unsigned) grid, kRgbThreads, Smem::total(a.n_lights), st>>>(a, (unsigned) n_tiles, tail_px);
    return cudaGetLastError();
}

cudaError_t launch_fit_rgb_tma(const FitArgs &a, int sm_count, cudaStream_t st, size_t *done) {
    *done = 0;
    const bool ok = a.n_lights <= 2048 && ((uintptr_t) a.stack % 16 == 0) && (a.image_stride % 16 == 0) &&
                    ((uintptr_t) a.coeffs % 16 == 0) && (a.plane_stride % 4 == 0);
    const size_t full_tiles = a.pixels / kRgbTilePx;
    const unsigned rem = (unsigned) (a.pixels % kRgbTilePx);
    const unsigned tail_px = (rem % 16 == 0) ? rem : 0;
    const size_t n_tiles = full_tiles + (tail_px ? 1 : 0);
    if (!ok || n_tiles == 0 || n_tiles > 0xFFFFFFFFull || a.n_lights > 257)
        return cudaSuccess;
    cudaError_t e = launch_rgb_variant<6, 4>(a, n_tiles, tail_px, sm_count, st);
    if (e == cudaSuccess)
        *done = full_tiles * kRgbTilePx + tail_px;
    return e;
}

// Launches the streaming kernel over the whole tiles of the slab (plus one partial tile) and returns how many
// pixels it covered through *done (0 if the layout does not allow bulk copies).
//
// Tile size.  Tiles are handed out dynamically, so the blocks finish within about one tile of each other; with
// 1024-pixel tiles that is ~10 us at 60 lights -- nothing in a 24 MP fit (79 tiles per CTA), 8 % of a 3 MP slab
// (the 8-GPU shape, 9.9 tiles per CTA).  A 4-consumer-warp instantiation (512-pixel tiles, four CTAs per SM, the
// same bytes in flight per SM) halves that spread but delivers worse: measured on B200 (profiles/r02_tile_sizes.txt)
// 3 MP 0.1102 -> 0.1163 ms, 6 MP 0.2208 -> 0.2349, 24 MP 0.862 -> 0.898 -- 1536-byte bulk copies and twice the
// barrier traffic cost more than the shorter tail saves.  It stays selectable (PTM_TMA_TILE=512, result-equivalent,
// tests/test_gpu_parity.py runs every case with both) but no slab size chooses it.
constexpr size_t kSmallSlabTilesPerCta = 0;

template <int W>
static cudaError_t launch_lrgb_tiles(const FitArgs &a, int sm_count, cudaStream_t st, size_t *done, int variant) {
    constexpr size_t tile_px = (size_t) W * 128;
    const size_t full_tiles = a.pixels / tile_px;
    const unsigned rem = (unsigned) (a.pixels % tile_px);
    // the remainder rides along as one partial tile when its byte count keeps the 16-byte granularity of bulk copies
    const unsigned tail_px = (rem % 16 == 0) ? rem : 0;
    const size_t n_tiles = full_tiles + (tail_px ? 1 : 0);
    if (n_tiles == 0 || n_tiles > 0xFFFFFFFFull)
        return cudaSuccess;
    cudaError_t e;
    if (W == 8) {
        switch (variant) {
        case 1: e = launch_variant<5, 5>(a, n_tiles, tail_px, sm_count, st); break;
        case 2: e = launch_variant<3, 8>(a, n_tiles, tail_px, sm_count, st); break;
        case 3: e = launch_variant<5, 4>(a, n_tiles, tail_px, sm_count, st); break;
        case 4: e = launch_variant<4, 5>(a, n_tiles, tail_px, sm_count, st); break;
        case 5: e = launch_variant<8, 3>(a, n_tiles, tail_px, sm_count, st); break;
#ifdef PTM_TUNING   // delivery ceiling only: skips the arithmetic, does NOT produce results
        case 11: e = launch_variant<6, 4, 1>(a, n_tiles, tail_px, sm_count, st); break;
#endif
        default: e = launch_variant<6, 4>(a, n_tiles, tail_px, sm_count, st); break;
        }
    } else {
        switch (variant) {
        case 2: e = launch_variant<3, 8, 0, W>(a, n_tiles, tail_px, sm_count, st); break;
        case 4: e = launch_variant<4, 5, 0, W>(a, n_tiles, tail_px, sm_count, st); break;
        default: e = launch_variant<6, 4, 0, W>(a, n_tiles, tail_px, sm_count, st); break;
        }
    }
    if (e == cudaSuccess)
        *done = full_tiles * tile_px + tail_px;
    return e;
}

cudaError_t launch_fit_lrgb_tma(const FitArgs &a, int sm_count, cudaStream_t st, size_t *done) {
    *done = 0;
    const bool ok = a.n_lights <= 257 && ((uintptr_t) a.stack % 16 == 0) && (a.image_stride % 16 == 0) &&
                    ((uintptr_t) a.rgb % 16 == 0) && (a.coeffs == nullptr || (uintptr_t) a.coeffs % 16 == 0);
    if (!ok || ((a.mode & kFitRgbIn) && !a.coeffs))
        return cudaSuccess;
    static int variant = -1;
    if (variant < 0) {
        const char *v = getenv("PTM_TMA_VARIANT");  // tuning aids; the defaults are what ships
        variant = v ? atoi(v) : 0;
    }
    const char *t = getenv("PTM_TMA_TILE");         // read per call: the tests run every case with both tile sizes
    const int forced_tile = t ? atoi(t) : 0;
    const size_t big_tiles = a.pixels / kTilePx;
    bool small = big_tiles < kSmallSlabTilesPerCta * (size_t) sm_count * 2;
    if (forced_tile == 512)
        small = true;
    else if (forced_tile == 1024)
        small = false;
    if (a.mode & kFitRgbIn)
        small = false;   // the RGB-input instantiation exists for the big tile only
    return small ? launch_lrgb_tiles<4>(a, sm_count, st, done, variant) : launch_lrgb_tiles<8>(a, sm_count, st, done, variant);
}

}  // namespace ptm
