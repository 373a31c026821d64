// Micro-benchmark: achievable HBM read bandwidth on B200 for the PTM stack access pattern.
//   R1  linear LDG.128 stream over the whole buffer
//   R2  LDG.128, but in K1's order: a CTA walks tile t through all 60 planes, 3072 B per plane
//   R3  bulk-copy (TMA) ring in K1's order, consumers only wait + release
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I rti_b200/csrc -o tools/ubench_read tools/ubench_read.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptm_tma.cuh"

using namespace ptm;

__global__ void r1(const uint4 *p, size_t n16, unsigned *sink) {
    unsigned acc = 0;
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x, step = (size_t) gridDim.x * blockDim.x;
    for (; i + 7 * step < n16; i += 8 * step) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcs(p + i + u * step);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    for (; i < n16; i += step) { uint4 v = __ldcs(p + i); acc += v.x ^ v.w; }
    if (acc == 0x12345678u) *sink = acc;
}

// tile = 3072 B per plane = 192 uint4; block of 192 threads reads one uint4 per plane, unroll over planes
__global__ void r2(const uint8_t *base, size_t plane_stride, unsigned n_planes, unsigned n_tiles, unsigned *sink) {
    unsigned acc = 0;
    for (unsigned t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const uint8_t *src = base + (size_t) t * 3072 + threadIdx.x * 16;
        for (unsigned n = 0; n + 11 < n_planes; n += 12) {
            uint4 v[12];
#pragma unroll
            for (int u = 0; u < 12; ++u) v[u] = __ldcs((const uint4 *) (src + (size_t) (n + u) * plane_stride));
#pragma unroll
            for (int u = 0; u < 12; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

template <int L, int S, int TILE>
__global__ void __launch_bounds__(288, 2) r3(const uint8_t *base, size_t plane_stride, unsigned n_planes, unsigned n_tiles, unsigned *sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *full = (uint64_t *) (smem + (size_t) S * L * TILE);
    uint64_t *empty = full + S;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
        mbar_fence_init();
    }
    __syncthreads();
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned n_chunks = n_planes / L;
    unsigned acc = 0;
    if (warp == 8) {
        const uint64_t pol = l2_policy_evict_first();
        unsigned s = 0, ph = 0;
        for (unsigned t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const uint8_t *src = base + (size_t) t * TILE + (size_t) lane * plane_stride;
            for (unsigned c = 0; c < n_chunks; ++c) {
                if (lane == 0) { mbar_wait(&empty[s], ph ^ 1); mbar_arrive_expect_tx(&full[s], L * TILE); }
                __syncwarp();
                if (lane < L) bulk_g2s(smem + ((size_t) s * L + lane) * TILE, src + (size_t) c * L * plane_stride, TILE, &full[s], pol);
                if (++s == S) { s = 0; ph ^= 1; }
            }
        }
    } else {
        unsigned s = 0, ph = 0;
        for (unsigned t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            for (unsigned c = 0; c < n_chunks; ++c) {
                mbar_wait(&full[s], ph);
                acc += ((const unsigned *) (smem + (size_t) s * L * TILE))[threadIdx.x];
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
                if (++s == S) { s = 0; ph ^= 1; }
            }
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}


// R4: R3 plus K1's output traffic (24 B/px floats + 3 B/px bytes) with different store shapes.
//   W=0 none, W=1 16 B per lane at 96 B stride (register direct), W=2 warp-coalesced 16 B stores,
//   W=3 TMA bulk store of the warp's 3072 + 384 bytes from shared memory
template <int L, int S, int W>
__global__ void __launch_bounds__(288, 2) r4(const uint8_t *base, size_t plane_stride, unsigned n_planes, unsigned n_tiles,
                                              float *coef, uint8_t *rgb, unsigned *sink) {
    constexpr int TILE = 3072;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *full = (uint64_t *) (smem + (size_t) S * L * TILE);
    uint64_t *empty = full + S;
    uint8_t *stg = smem + (size_t) S * L * TILE + 2 * S * 8 + 64;   // 8 warps x 3456 B (W=3 only)
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
        mbar_fence_init();
    }
    __syncthreads();
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned n_chunks = n_planes / L;
    unsigned acc = 0;
    if (warp == 8) {
        const uint64_t pol = l2_policy_evict_first();
        unsigned s = 0, ph = 0;
        for (unsigned t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const uint8_t *src = base + (size_t) t * TILE + (size_t) lane * plane_stride;
            for (unsigned c = 0; c < n_chunks; ++c) {
                if (lane == 0) { mbar_wait(&empty[s], ph ^ 1); mbar_arrive_expect_tx(&full[s], L * TILE); }
                __syncwarp();
                if (lane < L) bulk_g2s(smem + ((size_t) s * L + lane) * TILE, src + (size_t) c * L * plane_stride, TILE, &full[s], pol);
                if (++s == S) { s = 0; ph ^= 1; }
            }
        }
    } else {
        unsigned s = 0, ph = 0;
        for (unsigned t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            for (unsigned c = 0; c < n_chunks; ++c) {
                mbar_wait(&full[s], ph);
                acc += ((const unsigned *) (smem + (size_t) s * L * TILE))[threadIdx.x];
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
                if (++s == S) { s = 0; ph ^= 1; }
            }
            const size_t px = (size_t) t * 1024 + threadIdx.x * 4;
            const float f = __uint_as_float(acc);
            if (W == 1) {
                float4 *d = (float4 *) (coef + px * 6);
#pragma unroll
                for (int j = 0; j < 6; ++j) __stcg(d + j, make_float4(f, f, f, f));
                unsigned *r = (unsigned *) (rgb + px * 3);
                __stcg(r, acc); __stcg(r + 1, acc); __stcg(r + 2, acc);
            } else if (W == 2) {
                const size_t wpx = (size_t) t * 1024 + warp * 128;
                float4 *d = (float4 *) (coef + wpx * 6);
#pragma unroll
                for (int j = 0; j < 6; ++j) __stcg(d + lane + 32 * j, make_float4(f, f, f, f));
                if (lane < 24) __stcg((uint4 *) (rgb + wpx * 3) + lane, make_uint4(acc, acc, acc, acc));
            } else if (W == 3) {
                const size_t wpx = (size_t) t * 1024 + warp * 128;
                uint8_t *my = stg + warp * 3456;
                // wait until the previous bulk store has finished READING the staging buffer
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
                float4 *sf = (float4 *) my;
#pragma unroll
                for (int j = 0; j < 6; ++j) sf[lane * 6 + j] = make_float4(f, f, f, f);
                ((unsigned *) (my + 3072))[lane * 3] = acc; ((unsigned *) (my + 3072))[lane * 3 + 1] = acc; ((unsigned *) (my + 3072))[lane * 3 + 2] = acc;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(coef + wpx * 6), "r"(smem_addr(my)), "r"(3072) : "memory");
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(rgb + wpx * 3), "r"(smem_addr(my + 3072)), "r"(384) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        }
        if (W == 3 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    if (acc == 0x12345678u) *sink = acc;
}

template <typename F>
void timeit(const char *name, double bytes, F launch) {
    launch();
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    printf("%-44s %7.3f ms  %7.1f GB/s  (%s)\n", name, best, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const size_t plane = 72000000, n_planes = 60, total = plane * n_planes;
    uint8_t *buf; unsigned *sink;
    cudaMalloc(&buf, total); cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, total);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    const unsigned n_tiles = plane / 3072;
    timeit("R1 linear LDG.128 x8, 148x8 CTAs x 256", total, [&] { r1<<<sms * 8, 256>>>((const uint4 *) buf, total / 16, sink); });
    timeit("R1 linear LDG.128 x8, 148x4 CTAs x 512", total, [&] { r1<<<sms * 4, 512>>>((const uint4 *) buf, total / 16, sink); });
    timeit("R2 plane-walk LDG.128 x12, 148x8 CTAs x 192", total, [&] { r2<<<sms * 8, 192>>>(buf, plane, n_planes, n_tiles, sink); });
    timeit("R2 plane-walk LDG.128 x12, 148x10 CTAs x 192", total, [&] { r2<<<sms * 10, 192>>>(buf, plane, n_planes, n_tiles, sink); });
#define R3(L, S, TILE) { size_t sm = (size_t) S * L * TILE + 2 * S * 8; cudaFuncSetAttribute(r3<L, S, TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm); \
      timeit("R3 TMA ring L=" #L " S=" #S " tile=" #TILE, total, [&] { r3<L, S, TILE><<<sms * 2, 288, sm>>>(buf, plane, n_planes, plane / TILE, sink); }); }
    R3(4, 8, 3072)
    R3(10, 3, 3072)
    R3(6, 5, 3072)
    R3(5, 3, 6144)
    R3(10, 3, 1536)
    R3(2, 8, 6144)
    R3(1, 8, 12288)
    float *coef; uint8_t *rgb;
    cudaMalloc(&coef, 24000000ull * 24); cudaMalloc(&rgb, 24000000ull * 3);
#define R4(L, S, W) { size_t sm = (size_t) S * L * 3072 + 2 * S * 8 + 64 + 8 * 3456; cudaFuncSetAttribute(r4<L, S, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm); \
      timeit("R4 ring L=" #L " S=" #S " + writes, shape W=" #W, total + 27.0 * 24e6 * (W != 0), [&] { r4<L, S, W><<<sms * 2, 288, sm>>>(buf, plane, n_planes, plane / 3072, coef, rgb, sink); }); }
    R4(6, 4, 0)
    R4(6, 4, 1)
    R4(6, 4, 2)
    R4(6, 4, 3)
    R4(6, 4, 1)
    R4(6, 4, 2)
    R4(6, 4, 3)
    return 0;
}
