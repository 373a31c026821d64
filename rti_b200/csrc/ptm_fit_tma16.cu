// K1 for 16-bit stacks (the RAW-derived configuration: 100 lights, uint16 linear samples).
// Same structure as ptm_fit_tma.cu -- producer warp + bulk-copy ring + register accumulators --
// with 7 consumer warps (256 threads, 128 registers per thread at two CTAs per SM) and 6 bytes
// per pixel and light: a tile is 896 pixels = 5376 bytes per light.
//
//   kRgb = false  LRGB.  No counterpart in the reference (its ptm_cbcr_avg / colour loop are 8 bit):
//                 fit on the 16-bit luma like ptm_fit_poly_uint (rti-builder/ptmlib.c:727-760), 16-bit
//                 truncated means, colour conversion of rti-builder/ptm-encoder.c:331-339 on their top
//                 8 bits, norm = 256 / mean luma (oracle: orc_lrgb16_finish).
//   kRgb = true   RGB formats: three fits per pixel (rti-builder/ptm-encoder.c:272-284 with
//                 ptm_fit_poly_uint semantics), extrema from plane 0 only (rti-builder/ptmlib.c:831).
//
// A lane owns 4 pixels = 24 bytes = six words per light, read as three 8-byte shared loads
// (lane stride 6 words: conflict free per half warp):
//   w0 = Y0|Cb0  w1 = Cr0|Y1  w2 = Cb1|Cr1  w3 = Y2|Cb2  w4 = Cr2|Y3  w5 = Cb3|Cr3   (low | high half)
// Sums for the means: `all += w` and `hi += w >> 16` per word; the low-half sum is all - (hi << 16).
#include "ptm_device.cuh"
#include "ptm_kernels.h"
#include "ptm_tma.cuh"

namespace ptm {

constexpr int k16Warps = 7;
constexpr int k16Threads = k16Warps * 32 + 32;
constexpr int k16TilePx = k16Warps * 128;          // 896
constexpr int k16TileBytes = k16TilePx * 6;        // 5376 per light
constexpr int k16WarpStage = 128 * 24 + 128 * 3;   // a warp's 128 pixels: floats + colour bytes

template <int L, int S>
struct Smem16 {
    static constexpr size_t stage_bytes = (size_t) L * k16TileBytes;
    static constexpr size_t bars = (size_t) S * stage_bytes;
    static constexpr size_t staging = bars + 2 * S * sizeof(uint64_t);
    static constexpr size_t matrix = staging + (size_t) k16Warps * k16WarpStage;
    static constexpr size_t total(unsigned n_lights) { return matrix + (size_t) n_lights * 32; }
};

__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}

__device__ __forceinline__ void fma3(unsigned long long (&c)[3], float v, unsigned long long m01,
                                     unsigned long long m23, unsigned long long m45) {
    const unsigned long long vv = pack2(v, v);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c[0]) : "l"(vv), "l"(m01));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c[1]) : "l"(vv), "l"(m23));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c[2]) : "l"(vv), "l"(m45));
}

// coefficient store with an L2 policy (FitArgs::store_hint), see fit_lrgb_u8_tma_kernel
__device__ __forceinline__ void st_coeff16(float4 *p, float4 v, int hint, uint64_t policy) {
    if (hint)
        asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y),
                     "f"(v.z), "f"(v.w), "l"(policy)
                     : "memory");
    else
        __stcg(p, v);
}

__device__ __forceinline__ float pair_lo(unsigned long long p) { return __uint_as_float((unsigned) (p & 0xFFFFFFFFull)); }
__device__ __forceinline__ float pair_hi(unsigned long long p) { return __uint_as_float((unsigned) (p >> 32)); }

template <int L, int S, bool kRgb, bool kFull>
__global__ void __launch_bounds__(k16Threads, 2)
fit_u16_tma_kernel(FitArgs a, unsigned n_tiles, unsigned tail_px) {
    using Smem = Smem16<L, S>;
    constexpr int kAcc = kRgb ? 12 : 4;   // fits per lane: 4 pixels x (3 channels | luma)
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ int keys_s[2 * PTM_NCOEF];
    __shared__ unsigned tile_pub[2];   // dynamic tile schedule, see fit_lrgb_u8_tma_kernel
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
            mbar_init(&empty[s], k16Warps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned n_chunks = (n + L - 1) / L;
    Extrema ext;
    ext.init();

    if (warp == k16Warps) {
        // ------------------------------------------------------------ producer
        const uint64_t policy = l2_policy_evict_first();
        unsigned s = 0, ph = 0, slot = 0;
        unsigned tile = blockIdx.x;
        for (;;) {
            if (lane == 0) {
                mbar_wait(&empty[s], ph ^ 1);
                tile_pub[slot] = tile;
            }
            if (tile >= n_tiles) {
                if (lane == 0)
                    mbar_arrive(&full[s]);             // wake the consumers: nothing more to do
                break;
            }
            unsigned next = tile + gridDim.x;
            if (a.tile_counter) {                      // the next tile's draw hides behind this tile's copies
                if (lane == 0)
                    next = gridDim.x + atomicAdd(a.tile_counter, 1u);
                next = __shfl_sync(0xFFFFFFFFu, next, 0);
            }
            const uint8_t *src = a.stack + (size_t) tile * k16TileBytes + (size_t) lane * a.image_stride;
            const unsigned bytes = (tail_px && tile + 1 == n_tiles) ? tail_px * 6 : (unsigned) k16TileBytes;
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
                    bulk_g2s(ring + (size_t) s * Smem::stage_bytes + (size_t) lane * k16TileBytes,
                             src + (size_t) l0 * a.image_stride, bytes, &full[s], policy);
                if (++s == S) {
                    s = 0;
                    ph ^= 1;
                }
            }
            tile = next;
            slot ^= 1;
        }
    } else {
        // ----------------------------------------------------------- consumers
        const unsigned t = threadIdx.x;  // owns pixels 4t..4t+3 of the tile
        unsigned s = 0, ph = 0, slot = 0;
        uint64_t st_policy = 0;
        if (a.store_hint == 1)
            st_policy = l2_policy_evict_first();
        else if (a.store_hint == 2)
            asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(st_policy));
        for (;; slot ^= 1) {
            unsigned tile = 0;
            bool stop = false;
            unsigned long long acc[kAcc][3];
#pragma unroll
            for (int j = 0; j < kAcc; ++j)
                acc[j][0] = acc[j][1] = acc[j][2] = 0ull;
            unsigned all[6] = {0, 0, 0, 0, 0, 0}, hi[6] = {0, 0, 0, 0, 0, 0};

            for (unsigned c = 0; c < n_chunks; ++c) {
                mbar_wait(&full[s], ph);
                if (c == 0) {
                    tile = *reinterpret_cast<volatile unsigned *>(&tile_pub[slot]);
                    if (tile >= n_tiles) {
                        stop = true;
                        break;
                    }
                }
                const uint8_t *buf = ring + (size_t) s * Smem::stage_bytes + (size_t) t * 24;
                const unsigned l0 = c * L;
                const unsigned cnt = kFull ? (unsigned) L : min((unsigned) L, n - l0);
#pragma unroll
                for (unsigned l = 0; l < L; ++l) {
                    if (kFull || l < cnt) {
                        const uint2 *p2 = reinterpret_cast<const uint2 *>(buf + (size_t) l * k16TileBytes);
                        const uint2 d0 = p2[0], d1 = p2[1], d2 = p2[2];
                        const unsigned w[6] = {d0.x, d0.y, d1.x, d1.y, d2.x, d2.y};
                        const float4 ma = m_s[2 * (l0 + l)], mb = m_s[2 * (l0 + l) + 1];
                        const unsigned long long m01 = pack2(ma.x, ma.y), m23 = pack2(ma.z, ma.w),
                                                 m45 = pack2(mb.x, mb.y);
                        if (kRgb) {
#pragma unroll
                            for (int j = 0; j < 12; ++j)   // half-word j = pixel j / 3, channel j % 3
                                fma3(acc[kRgb ? j : 0], (j & 1) ? half_to_float_hi(w[j >> 1]) : half_to_float_lo(w[j >> 1]),
                                     m01, m23, m45);
                        } else {
                            fma3(acc[0], half_to_float_lo(w[0]), m01, m23, m45);
                            fma3(acc[1], half_to_float_hi(w[1]), m01, m23, m45);
                            fma3(acc[2], half_to_float_lo(w[3]), m01, m23, m45);
                            fma3(acc[3], half_to_float_hi(w[4]), m01, m23, m45);
#pragma unroll
                            for (int j = 0; j < 6; ++j) {
                                all[j] += w[j];
                                hi[j] += w[j] >> 16;
                            }
                        }
                    }
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

            const unsigned tile_px = (tail_px && tile + 1 == n_tiles) ? tail_px : (unsigned) k16TilePx;
            const unsigned valid = min(128u, tile_px > warp * 128 ? tile_px - warp * 128 : 0u);
            const size_t wpx = (size_t) tile * k16TilePx + (size_t) warp * 128;
            uint8_t *stg = smem + Smem::staging + (size_t) warp * k16WarpStage;
            const bool live = 4 * t < tile_px;

            if (!kRgb) {
                // half-word sums in pixel order: Y0 Cb0 Cr0 Y1 Cb1 Cr1 ...
                unsigned sum[12];
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    sum[2 * j] = all[j] - (hi[j] << 16);
                    sum[2 * j + 1] = hi[j];
                }
                float c[4][PTM_NCOEF];
                unsigned out_b[12];
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const unsigned ya = sum[3 * p] / n, cba = sum[3 * p + 1] / n, cra = sum[3 * p + 2] / n;
                    float dummy[PTM_NCOEF] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    lrgb_finish(ya >> 8, cba >> 8, cra >> 8, dummy, out_b[3 * p], out_b[3 * p + 1], out_b[3 * p + 2]);
                    const float norm = __fdiv_rn(256.0f, (float) ya);
#pragma unroll
                    for (int h = 0; h < 3; ++h) {
                        c[p][2 * h] = __fmul_rn(pair_lo(acc[p][h]), norm);
                        c[p][2 * h + 1] = __fmul_rn(pair_hi(acc[p][h]), norm);
                    }
                }
                if (live)
                    ext.add4(c);
                if (a.coeffs) {
                    float4 *row = reinterpret_cast<float4 *>(stg) + lane * 6;
                    const float *f = &c[0][0];
#pragma unroll
                    for (int j = 0; j < 6; ++j)
                        row[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                }
                uint32_t *rgbw = reinterpret_cast<uint32_t *>(stg + 128 * 24);
                rgbw[3 * lane + 0] = out_b[0] | (out_b[1] << 8) | (out_b[2] << 16) | (out_b[3] << 24);
                rgbw[3 * lane + 1] = out_b[4] | (out_b[5] << 8) | (out_b[6] << 16) | (out_b[7] << 24);
                rgbw[3 * lane + 2] = out_b[8] | (out_b[9] << 8) | (out_b[10] << 16) | (out_b[11] << 24);
                __syncwarp();
                if (a.coeffs) {
                    float4 *dst = reinterpret_cast<float4 *>(a.coeffs + wpx * PTM_NCOEF);
                    const float4 *src = reinterpret_cast<const float4 *>(stg);
#pragma unroll
                    for (int r = 0; r < 6; ++r)
                        if (lane + 32 * r < valid * 6 / 4)
                            st_coeff16(dst + lane + 32 * r, src[lane + 32 * r], a.store_hint, st_policy);
                }
                if (lane < valid * 3 / 16)
                    __stcg(reinterpret_cast<uint4 *>(a.rgb + wpx * 3) + lane, reinterpret_cast<const uint4 *>(rgbw)[lane]);
                __syncwarp();
            } else {
                if (live) {
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        float c0[PTM_NCOEF];
#pragma unroll
                        for (int h = 0; h < 3; ++h) {
                            c0[2 * h] = pair_lo(acc[kRgb ? p * 3 : 0][h]);
                            c0[2 * h + 1] = pair_hi(acc[kRgb ? p * 3 : 0][h]);
                        }
                        ext.add(c0);
                    }
                }
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    unsigned long long *row = reinterpret_cast<unsigned long long *>(stg) + lane * 12;
#pragma unroll
                    for (int p = 0; p < 4; ++p)
#pragma unroll
                        for (int h = 0; h < 3; ++h)
                            row[p * 3 + h] = acc[kRgb ? p * 3 + ch : 0][h];
                    __syncwarp();
                    float4 *dst = reinterpret_cast<float4 *>(a.coeffs + (size_t) ch * a.plane_stride + wpx * PTM_NCOEF);
                    const float4 *src = reinterpret_cast<const float4 *>(stg);
#pragma unroll
                    for (int r = 0; r < 6; ++r)
                        if (lane + 32 * r < valid * 6 / 4)
                            st_coeff16(dst + lane + 32 * r, src[lane + 32 * r], a.store_hint, st_policy);
                    __syncwarp();
                }
            }
        }
    }
    ext.flush(a.keys, keys_s);
}

template <int L, int S, bool kRgb>
static cudaError_t launch16(const FitArgs &a, size_t n_tiles, unsigned tail_px, int sm_count, cudaStream_t st) {
    using Smem = Smem16<L, S>;
    static int per_sm_cache[64][2];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess)
        return e;
    if (dev < 0 || dev >= 64)
        return cudaErrorInvalidDevice;
    const bool full = a.n_lights % L == 0;
    auto kern = full ? fit_u16_tma_kernel<L, S, kRgb, true> : fit_u16_tma_kernel<L, S, kRgb, false>;
    if (per_sm_cache[dev][full] == 0) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) Smem::total(512));
        if (e != cudaSuccess)
            return e;
        int nb = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, k16Threads, Smem::total(512));
        if (e != cudaSuccess)
            return e;
        if (nb < 1)
            return cudaErrorLaunchOutOfResources;
        per_sm_cache[dev][full] = nb;
    }
    size_t grid = (size_t) sm_count * per_sm_cache[dev][full];
    if (grid > n_tiles)
        grid = n_tiles;
    kern<<<(unsigned) grid, k16Threads, Smem::total(a.n_lights), st>>>(a, (unsigned) n_tiles, tail_px);
    e = cudaGetLastError();
    if (e == cudaSuccess && a.tile_counter)   // the tile counter is 0 between launches
        e = cudaMemsetAsync(a.tile_counter, 0, sizeof(unsigned), st);
    return e;
}

// Streaming fit of a uint16 stack over whole 896-pixel tiles (+ a partial tile whose pixel count is a
// multiple of 16).  *done = pixels covered (0 when the layout does not allow bulk copies).
cudaError_t launch_fit_u16_tma(const FitArgs &a, bool rgb, int sm_count, cudaStream_t st, size_t *done) {
    *done = 0;
    bool ok = a.n_lights <= 512 && ((uintptr_t) a.stack % 16 == 0) && (a.image_stride % 16 == 0) &&
              (a.coeffs == nullptr || (uintptr_t) a.coeffs % 16 == 0);
    if (rgb)
        ok = ok && a.coeffs != nullptr && a.plane_stride % 4 == 0;
    else
        ok = ok && (uintptr_t) a.rgb % 16 == 0;
    const size_t full_tiles = a.pixels / k16TilePx;
    const unsigned rem = (unsigned) (a.pixels % k16TilePx);
    const unsigned tail_px = (rem % 16 == 0) ? rem : 0;
    const size_t n_tiles = full_tiles + (tail_px ? 1 : 0);
    if (!ok || n_tiles == 0 || n_tiles > 0xFFFFFFFFull)
        return cudaSuccess;
    cudaError_t e = rgb ? launch16<4, 3, true>(a, n_tiles, tail_px, sm_count, st)
                        : launch16<4, 3, false>(a, n_tiles, tail_px, sm_count, st);
    if (e == cudaSuccess)
        *done = full_tiles * k16TilePx + tail_px;
    return e;
}

}  // namespace ptm
