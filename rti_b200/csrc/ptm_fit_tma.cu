// K1, streaming form -- the LRGB fit as a persistent, warp-specialised sm_100a kernel.
//
//   producer warp : one elected lane feeds a shared-memory ring with 1-D bulk copies
//                   (cp.async.bulk, SASS UBLKCP): each copy is one light's 3072-byte run of
//                   1024 consecutive pixels; four lights per stage, four stages in flight,
//                   completion counted in bytes on an mbarrier per stage.
//   consumer warps: 8 warps x 32 lanes x 4 pixels = the 1024-pixel tile.  Per light a lane
//                   reads its 12 bytes as 3 words (lane stride 3 words: conflict free), turns
//                   the four luma bytes into floats with PRMT + FADD, and issues 24 FFMA
//                   against the pseudo-inverse column broadcast from shared memory.  All 12
//                   bytes are also summed for ptm_cbcr_avg in two 32-bit accumulators per word:
//                   `ev += w` (no unpacking at all) and `od += bytes 1,3 as 16-bit lanes`;
//                   the even-byte sums are recovered at the end as ev - (S1 << 8) - (S3 << 24).
//   epilogue      : integer mean, JFIF colour conversion, 256 / Y normalisation, running
//                   extrema, and a per-warp shared-memory transpose so the 96 bytes of float
//                   coefficients per lane leave as fully coalesced 16-byte stores.
//
// HBM latency is covered by the ring (up to 48 KB in flight per CTA, two CTAs per SM), not by
// occupancy, so the consumers keep their accumulators in registers without spilling.
//
// Reference arithmetic fused here: rti-builder/ptmlib.c:692-725 (fit), :762-802 (average),
// rti-builder/ptm-encoder.c:327-353 (colour + normalise), rti-builder/ptmlib.c:826-847 (extrema).
#include "ptm_device.cuh"
#include "ptm_kernels.h"
#include "ptm_tma.cuh"

namespace ptm {

constexpr int kConsumerWarps = 8;
constexpr int kConsumers = kConsumerWarps * 32;          // 256
constexpr int kTmaThreads = kConsumers + 32;             // + producer warp
constexpr int kTilePx = kConsumers * 4;                  // 1024
constexpr int kTileBytes = kTilePx * 3;                  // 3072 per light
constexpr int kLightsPerStage = 4;
constexpr int kStages = 4;
constexpr int kStageBytes = kLightsPerStage * kTileBytes;
constexpr int kStageWordsPerWarp = 32 * 28 + 32 * 3;     // padded coefficient rows + rgb words

struct TmaSmem {
    // dynamic shared memory layout (all offsets multiples of 16 bytes)
    static __host__ __device__ size_t ring() { return 0; }
    static __host__ __device__ size_t staging() { return (size_t) kStages * kStageBytes; }
    static __host__ __device__ size_t bars() { return staging() + (size_t) kConsumerWarps * kStageWordsPerWarp * 4; }
    static __host__ __device__ size_t matrix() { return bars() + 2 * kStages * sizeof(uint64_t); }
    static __host__ __device__ size_t total(unsigned n_lights) { return matrix() + (size_t) n_lights * 32; }
};

// One light's contribution to a lane's four pixels.
//   word0 = Y0 Cb0 Cr0 Y1 | word1 = Cb1 Cr1 Y2 Cb2 | word2 = Cr2 Y3 Cb3 Cr3
__device__ __forceinline__ void consume_light(const uint32_t *s32, const float4 *m, float (&acc)[4][PTM_NCOEF],
                                              unsigned (&ev)[3], unsigned (&od)[3]) {
    const unsigned w0 = s32[0], w1 = s32[1], w2 = s32[2];
    const float4 ma = m[0], mb = m[1];
    const float y[4] = {byte_to_float(w0, PTM_SEL_B0), byte_to_float(w0, PTM_SEL_B3),
                        byte_to_float(w1, PTM_SEL_B2), byte_to_float(w2, PTM_SEL_B1)};
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        acc[p][0] = fmaf(ma.x, y[p], acc[p][0]);
        acc[p][1] = fmaf(ma.y, y[p], acc[p][1]);
        acc[p][2] = fmaf(ma.z, y[p], acc[p][2]);
        acc[p][3] = fmaf(ma.w, y[p], acc[p][3]);
        acc[p][4] = fmaf(mb.x, y[p], acc[p][4]);
        acc[p][5] = fmaf(mb.y, y[p], acc[p][5]);
    }
    ev[0] += w0;
    ev[1] += w1;
    ev[2] += w2;
    od[0] += __byte_perm(w0, 0, 0x4341);
    od[1] += __byte_perm(w1, 0, 0x4341);
    od[2] += __byte_perm(w2, 0, 0x4341);
}

template <bool kWriteCoeffs>
__global__ void __launch_bounds__(kTmaThreads, 2)
fit_lrgb_u8_tma_kernel(FitArgs a, unsigned n_tiles) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ int keys_s[2 * PTM_NCOEF];
    uint8_t *ring = smem + TmaSmem::ring();
    float *staging = reinterpret_cast<float *>(smem + TmaSmem::staging());
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + TmaSmem::bars());
    uint64_t *empty = full + kStages;
    float4 *m_s = reinterpret_cast<float4 *>(smem + TmaSmem::matrix());

    const unsigned n = a.n_lights;
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
        const float *M = a.M;
        m_s[2 * i] = make_float4(M[i], M[n + i], M[2 * n + i], M[3 * n + i]);
        m_s[2 * i + 1] = make_float4(M[4 * n + i], M[5 * n + i], 0.f, 0.f);
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kConsumerWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned n_chunks = (n + kLightsPerStage - 1) / kLightsPerStage;
    Extrema ext;
    ext.init();

    if (warp == kConsumerWarps) {
        // ------------------------------------------------------------ producer
        if (lane == 0) {
            const uint64_t policy = l2_policy_evict_first();
            unsigned it = 0;
            for (unsigned tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const uint8_t *src = a.stack + (size_t) tile * kTileBytes;
                for (unsigned c = 0; c < n_chunks; ++c, ++it) {
                    const unsigned s = it % kStages, ph = (it / kStages) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    const unsigned l0 = c * kLightsPerStage;
                    const unsigned cnt = min((unsigned) kLightsPerStage, n - l0);
                    mbar_arrive_expect_tx(&full[s], cnt * kTileBytes);
                    for (unsigned l = 0; l < cnt; ++l)
                        bulk_g2s(ring + (size_t) s * kStageBytes + (size_t) l * kTileBytes,
                                 src + (size_t) (l0 + l) * a.image_stride, kTileBytes, &full[s], policy);
                }
            }
        }
    } else {
        // ----------------------------------------------------------- consumers
        const unsigned t = threadIdx.x;  // 0..255, owns pixels 4t..4t+3 of the tile
        float *stg = staging + (size_t) warp * kStageWordsPerWarp;
        unsigned it = 0;
        for (unsigned tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            float acc[4][PTM_NCOEF];
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int k = 0; k < PTM_NCOEF; ++k)
                    acc[p][k] = 0.f;
            unsigned ev[3] = {0, 0, 0}, od[3] = {0, 0, 0};

            for (unsigned c = 0; c < n_chunks; ++c, ++it) {
                const unsigned s = it % kStages, ph = (it / kStages) & 1;
                mbar_wait(&full[s], ph);
                const uint32_t *buf = reinterpret_cast<const uint32_t *>(ring + (size_t) s * kStageBytes) + 3 * t;
                const unsigned l0 = c * kLightsPerStage;
                const unsigned cnt = min((unsigned) kLightsPerStage, n - l0);
                if (cnt == kLightsPerStage) {
                    // straight-line body: all loads of the stage can be hoisted, adds pair into IADD3
#pragma unroll
                    for (unsigned l = 0; l < kLightsPerStage; ++l)
                        consume_light(buf + l * (kTileBytes / 4), m_s + 2 * (l0 + l), acc, ev, od);
                } else {
                    for (unsigned l = 0; l < cnt; ++l)
                        consume_light(buf + l * (kTileBytes / 4), m_s + 2 * (l0 + l), acc, ev, od);
                }
                __syncwarp();
                if (lane == 0)
                    mbar_arrive(&empty[s]);
            }

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
                lrgb_finish(ya, cba, cra, acc[p], out_b[3 * p], out_b[3 * p + 1], out_b[3 * p + 2]);
                ext.add(acc[p]);
            }

            // warp-local transpose through shared memory -> coalesced 16-byte stores
            const size_t px0 = (size_t) tile * kTilePx + (size_t) warp * 128;  // first pixel of this warp
            if (kWriteCoeffs) {
                float4 *row = reinterpret_cast<float4 *>(stg + lane * 28);
                const float *f = &acc[0][0];
#pragma unroll
                for (int j = 0; j < 6; ++j)
                    row[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            }
            uint32_t *rgbw = reinterpret_cast<uint32_t *>(stg + 32 * 28);
            rgbw[3 * lane + 0] = out_b[0] | (out_b[1] << 8) | (out_b[2] << 16) | (out_b[3] << 24);
            rgbw[3 * lane + 1] = out_b[4] | (out_b[5] << 8) | (out_b[6] << 16) | (out_b[7] << 24);
            rgbw[3 * lane + 2] = out_b[8] | (out_b[9] << 8) | (out_b[10] << 16) | (out_b[11] << 24);
            __syncwarp();
            if (kWriteCoeffs) {
                float4 *dst = reinterpret_cast<float4 *>(a.coeffs + px0 * PTM_NCOEF);
#pragma unroll
                for (int r = 0; r < 6; ++r) {
                    const unsigned q = lane + 32 * r;  // float4 index within the warp's 192
                    const unsigned tp = q / 6, j = q - tp * 6;
                    __stcg(dst + q, *reinterpret_cast<const float4 *>(stg + tp * 28 + j * 4));
                }
            }
            if (lane < 24)
                __stcg(reinterpret_cast<uint4 *>(a.rgb + px0 * 3) + lane,
                       *reinterpret_cast<const uint4 *>(rgbw + 4 * lane));
            __syncwarp();
        }
    }
    ext.flush(a.keys, keys_s);
}

// Number of resident CTAs per SM for the streaming kernel (computed once per launch config).
static int tma_blocks_per_sm(size_t smem, bool write_coeffs) {
    int nb = 0;
    cudaError_t e = write_coeffs
        ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fit_lrgb_u8_tma_kernel<true>, kTmaThreads, smem)
        : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fit_lrgb_u8_tma_kernel<false>, kTmaThreads, smem);
    return e == cudaSuccess ? nb : 0;
}

// Launches the streaming kernel over the whole 1024-pixel tiles of the slab and returns how many
// pixels it covered through *done (0 if the layout does not allow bulk copies).
cudaError_t launch_fit_lrgb_tma(const FitArgs &a, int sm_count, cudaStream_t st, size_t *done) {
    *done = 0;
    const bool ok = a.n_lights <= 257 && ((uintptr_t) a.stack % 16 == 0) && (a.image_stride % 16 == 0) &&
                    ((uintptr_t) a.rgb % 16 == 0) && (a.coeffs == nullptr || (uintptr_t) a.coeffs % 16 == 0);
    const size_t n_tiles = a.pixels / kTilePx;
    if (!ok || n_tiles == 0 || n_tiles > 0xFFFFFFFFull)
        return cudaSuccess;
    const size_t smem = TmaSmem::total(a.n_lights);
    // opt in to > 48 KB of dynamic shared memory and look up residency, once per device
    static int per_sm_cache[64][2];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess)
        return e;
    const int wc = a.coeffs != nullptr;
    if (dev < 0 || dev >= 64)
        return cudaErrorInvalidDevice;
    if (per_sm_cache[dev][wc] == 0) {
        e = wc ? cudaFuncSetAttribute(fit_lrgb_u8_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int) TmaSmem::total(257))
               : cudaFuncSetAttribute(fit_lrgb_u8_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int) TmaSmem::total(257));
        if (e != cudaSuccess)
            return e;
        per_sm_cache[dev][wc] = tma_blocks_per_sm(TmaSmem::total(257), wc);
        if (per_sm_cache[dev][wc] < 1)
            return cudaErrorLaunchOutOfResources;
    }
    const int per_sm = per_sm_cache[dev][wc];
    size_t grid = (size_t) sm_count * per_sm;
    if (grid > n_tiles)
        grid = n_tiles;
    if (a.coeffs)
        fit_lrgb_u8_tma_kernel<true><<<(unsigned) grid, kTmaThreads, smem, st>>>(a, (unsigned) n_tiles);
    else
        fit_lrgb_u8_tma_kernel<false><<<(unsigned) grid, kTmaThreads, smem, st>>>(a, (unsigned) n_tiles);
    *done = n_tiles * kTilePx;
    return cudaGetLastError();
}

}  // namespace ptm
