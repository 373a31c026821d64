// K2 -- scale / bias derivation and uint8 quantisation
// (phases B and C of ptm_scale_coefficients, rti-builder/ptmlib.c:858-880).
//
// The coefficient plane is a flat float array whose element e belongs to
// coefficient e % 6.  Each thread handles float4 q (elements 4q..4q+3), so its
// coefficient pattern depends only on q % 3; the grid stride is a multiple of 3
// float4s, which makes that pattern a per-thread constant held in registers.
// Loads are 16 B per lane, stores 4 B per lane, both fully coalesced.
#include <algorithm>
#include <cstdlib>

#include "ptm_device.cuh"
#include "ptm_kernels.h"

namespace ptm {

constexpr int kQuantThreads = 384;  // multiple of 3 and of 32
constexpr int kQuantUnroll = 4;

// kConsume: the float plane is scratch that nobody reads afterwards (ptmb_set_coeffs_scratch).  The last `tail_q`
// float4 of the plane -- the part K1 wrote last and that is still dirty in L2 -- are then quantised FIRST, and every
// 128-byte line is discarded from L2 once read, so it is never written back to HBM; the rest of the plane follows as
// an ordinary forward stream.  Measured with ncu range replay over whole encodes (profiles/r02_k2_tail.txt): walking
// the WHOLE plane backwards saves 114 MB of HBM traffic per 24 MP encode but streams the HBM-resident part slower
// (0.860 -> 0.869 ms); the tail-first form keeps the saving without that cost.
template <bool kConsume>
__global__ void __launch_bounds__(kQuantThreads)
quantise_kernel(const float *__restrict__ coeffs, size_t n_float4, size_t tail_q, size_t n_floats,
                KeySource ksrc, uint8_t *__restrict__ bytes, void *sb_out) {
    __shared__ float inv_s[PTM_NCOEF];
    __shared__ float bias_s[PTM_NCOEF];
    __shared__ int keys[2 * PTM_NCOEF];
    keys_fetch<KeySource, XchgLayout>(ksrc, keys, kMaxRanks, kSlotInts);
    // the next fit of a fused encode loop may start occupying SMs as this grid drains (it waits for
    // this grid's completion before touching memory, see fit_lrgb_u8_tma_kernel)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x < PTM_NCOEF) {
        float scale, inv;
        int bias;
        scale_bias_from_keys(keys, threadIdx.x, scale, bias, inv);
        inv_s[threadIdx.x] = inv;
        bias_s[threadIdx.x] = (float) bias;
        if (blockIdx.x == 0 && sb_out) {
            reinterpret_cast<float *>(sb_out)[threadIdx.x] = scale;
            reinterpret_cast<int *>(sb_out)[PTM_NCOEF + threadIdx.x] = bias;
        }
    }
    __syncthreads();

    const float4 *src = reinterpret_cast<const float4 *>(coeffs);
    uint32_t *dst = reinterpret_cast<uint32_t *>(bytes);
    const size_t step = (size_t) gridDim.x * kQuantThreads;  // multiple of 3 (and of 8)
    const size_t q0 = (size_t) blockIdx.x * kQuantThreads + threadIdx.x;
    // coefficient of the float4's first lane: (4 q) % 6 depends on q % 3 only, and q % 3 is the same for every index
    // this thread visits (step % 3 == 0; with kConsume the tail starts at a multiple of 24 float4 as well)
    const int k0 = (int) ((4 * (q0 % 3)) % PTM_NCOEF);
    float inv[4], bias[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        inv[j] = inv_s[(k0 + j) % PTM_NCOEF];
        bias[j] = bias_s[(k0 + j) % PTM_NCOEF];
    }
    auto quantise4 = [&](const float4 v) {
        const unsigned b0 = clip_u8(__fadd_rn(__fmul_rn(v.x, inv[0]), bias[0]));
        const unsigned b1 = clip_u8(__fadd_rn(__fmul_rn(v.y, inv[1]), bias[1]));
        const unsigned b2 = clip_u8(__fadd_rn(__fmul_rn(v.z, inv[2]), bias[2]));
        const unsigned b3 = clip_u8(__fadd_rn(__fmul_rn(v.w, inv[3]), bias[3]));
        return b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
    };

    const size_t head = kConsume ? n_float4 - tail_q : n_float4;   // float4 [0, head) stream forward, [head, n) first
    if (kConsume) {
        size_t q = q0;
        for (; q + (kQuantUnroll - 1) * step < tail_q; q += kQuantUnroll * step) {
            float4 v[kQuantUnroll];
#pragma unroll
            for (int u = 0; u < kQuantUnroll; ++u)
                v[u] = __ldcg(src + head + q + u * step);
#pragma unroll
            for (int u = 0; u < kQuantUnroll; ++u) {
                const size_t i = head + q + u * step;
                __stcs(dst + i, quantise4(v[u]));
                if ((i & 7) == 0)   // first float4 of a whole 128-byte line, read by 8 lanes of this instruction
                    asm volatile("discard.global.L2 [%0], 128;" ::"l"(src + i) : "memory");
            }
        }
        for (; q < tail_q; q += step) {
            const size_t i = head + q;
            __stcs(dst + i, quantise4(__ldcg(src + i)));
            if ((i & 7) == 0)
                asm volatile("discard.global.L2 [%0], 128;" ::"l"(src + i) : "memory");
        }
    }
    size_t q = q0;
    for (; q + (kQuantUnroll - 1) * step < head; q += kQuantUnroll * step) {
        float4 v[kQuantUnroll];
#pragma unroll
        for (int u = 0; u < kQuantUnroll; ++u)
            v[u] = __ldcg(src + q + u * step);
#pragma unroll
        for (int u = 0; u < kQuantUnroll; ++u)
            __stcs(dst + q + u * step, quantise4(v[u]));
    }
    for (; q < head; q += step)
        __stcs(dst + q, quantise4(__ldcg(src + q)));
    // scalar tail (fewer than 4 floats)
    if (blockIdx.x == 0) {
        for (size_t e = n_float4 * 4 + threadIdx.x; e < n_floats; e += kQuantThreads) {
            const int k = (int) (e % PTM_NCOEF);
            bytes[e] = (uint8_t) clip_u8(__fadd_rn(__fmul_rn(coeffs[e], inv_s[k]), bias_s[k]));
        }
    }
}

// Everything through the scalar path: used when pointers are not 16 / 4 byte aligned.
__global__ void __launch_bounds__(kQuantThreads)
quantise_scalar_kernel(const float *__restrict__ coeffs, size_t n_floats, KeySource ksrc,
                       uint8_t *__restrict__ bytes, void *sb_out) {
    __shared__ float inv_s[PTM_NCOEF];
    __shared__ float bias_s[PTM_NCOEF];
    __shared__ int keys[2 * PTM_NCOEF];
    keys_fetch<KeySource, XchgLayout>(ksrc, keys, kMaxRanks, kSlotInts);
    if (threadIdx.x < PTM_NCOEF) {
        float scale, inv;
        int bias;
        scale_bias_from_keys(keys, threadIdx.x, scale, bias, inv);
        inv_s[threadIdx.x] = inv;
        bias_s[threadIdx.x] = (float) bias;
        if (blockIdx.x == 0 && sb_out) {
            reinterpret_cast<float *>(sb_out)[threadIdx.x] = scale;
            reinterpret_cast<int *>(sb_out)[PTM_NCOEF + threadIdx.x] = bias;
        }
    }
    __syncthreads();
    for (size_t e = (size_t) blockIdx.x * kQuantThreads + threadIdx.x; e < n_floats;
         e += (size_t) gridDim.x * kQuantThreads) {
        const int k = (int) (e % PTM_NCOEF);
        bytes[e] = (uint8_t) clip_u8(__fadd_rn(__fmul_rn(coeffs[e], inv_s[k]), bias_s[k]));
    }
}

cudaError_t launch_quantise(const float *coeffs, size_t pixels, const KeySource &keys, uint8_t *bytes, void *sb,
                            bool consume, int sm_count, cudaStream_t st) {
    const size_t n_floats = pixels * PTM_NCOEF;
    const bool aligned = ((uintptr_t) coeffs % 16 == 0) && ((uintptr_t) bytes % 4 == 0);
    if (!aligned) {
        size_t want = (n_floats + kQuantThreads - 1) / kQuantThreads;
        size_t cap = (size_t) sm_count * 16;
        unsigned grid = (unsigned) (want < 1 ? 1 : (want < cap ? want : cap));
        quantise_scalar_kernel<<<grid, kQuantThreads, 0, st>>>(coeffs, n_floats, keys, bytes, sb);
        return cudaGetLastError();
    }
    const size_t n_float4 = n_floats / 4;
    size_t want = (n_float4 + (size_t) kQuantThreads * kQuantUnroll - 1) / ((size_t) kQuantThreads * kQuantUnroll);
    size_t cap = (size_t) sm_count * 5;
    unsigned grid = (unsigned) (want < 1 ? 1 : (want < cap ? want : cap));
    // Tail-first consumption needs the tail to start on a multiple of 24 float4 (coefficient pattern: multiple of 3;
    // 128-byte lines read by 8 lanes of one warp instruction: multiple of 8) in a line-aligned plane: pixels % 16 == 0.
    // Anything else takes the plain forward kernel.  The tail is what K1 can have left in L2 (PTM_K2_TAIL_MB, default
    // 96 MB: 24 MP ran in 0.876 / 0.869 / 0.865 ms with 32 / 64 / 96 MB against 0.875 without; a plane that fits is all tail).
    const char *tail_env = getenv("PTM_K2_TAIL_MB");   // read per call (the tests split small planes with it)
    const long tail_mb = tail_env && atol(tail_env) > 0 ? atol(tail_env) : 96;
    size_t tail_q = std::min<size_t>(n_float4, ((size_t) tail_mb << 20) / 16);
    tail_q -= tail_q % 24;
    const bool can_consume = consume && n_float4 % 24 == 0 && tail_q > 0 && (uintptr_t) coeffs % 128 == 0;
    if (keys.mailbox != nullptr) {
        // Fused hand-off: K2 waits for the keys in its prologue (mailbox flags), so it may be launched
        // programmatically dependent on K1 -- its blocks become resident while K1 drains and the launch
        // latency disappears from the critical path.  K1 issues griddepcontrol.launch_dependents at its
        // start; K2 never reads K1's output before the flags are up (acquire loads + L1-bypassing reads).
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(kQuantThreads);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (can_consume)
            return cudaLaunchKernelEx(&cfg, quantise_kernel<true>, coeffs, n_float4, tail_q, n_floats, keys, bytes, sb);
        return cudaLaunchKernelEx(&cfg, quantise_kernel<false>, coeffs, n_float4, (size_t) 0, n_floats, keys, bytes, sb);
    } else if (can_consume)
        quantise_kernel<true><<<grid, kQuantThreads, 0, st>>>(coeffs, n_float4, tail_q, n_floats, keys, bytes, sb);
    else
        quantise_kernel<false><<<grid, kQuantThreads, 0, st>>>(coeffs, n_float4, (size_t) 0, n_floats, keys, bytes, sb);
    return cudaGetLastError();
}

}  // namespace ptm
