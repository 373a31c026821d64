// Device-side helpers shared by the PTM kernels (sm_100a).
//
// Parity rules: the reference is built without FMA
// contraction (rti-builder/Makefile:5-7), truncates float->byte after clipping
// (rti-builder/ptmlib.h:156-158) and does its bias arithmetic in int.  This
// translation unit is compiled with -fmad=false; the only fused multiply-adds
// are the explicit fmaf() calls of the per-pixel dot product, whose summation
// order the reference itself leaves to its BLAS.
#pragma once

#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

#define PTM_NCOEF 6

namespace ptm {

// Order-preserving int32 image of a float: a < b  <=>  key(a) < key(b).
__host__ __device__ __forceinline__ int float_to_key(float f) {
#ifdef __CUDA_ARCH__
    int b = __float_as_int(f);
#else
    int b;
    memcpy(&b, &f, 4);
#endif
    return b ^ ((b >> 31) & 0x7FFFFFFF);
}

__host__ __device__ __forceinline__ float key_to_float(int k) {
    int b = k ^ ((k >> 31) & 0x7FFFFFFF);
#ifdef __CUDA_ARCH__
    return __int_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}

// rti-builder/ptmlib.h:156-158: f > 255 -> 255, f < 0 -> 0, else truncate.
// cvt.rzi.u32.f32 saturates (negative -> 0) and truncates; min() does the rest.
__device__ __forceinline__ unsigned clip_u8(float f) {
    return min(__float2uint_rz(f), 255u);
}

// Exact uint8 -> float without the conversion unit: place the byte in the
// mantissa of 2^23 and subtract 2^23.  `sel` is a PRMT selector that moves the
// wanted byte of `w` to byte 0 and takes bytes 1..3 from 0x4B000000.
__device__ __forceinline__ float byte_to_float(unsigned w, unsigned sel) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, sel)) - 8388608.0f;
}
#define PTM_SEL_B0 0x7650u
#define PTM_SEL_B1 0x7651u
#define PTM_SEL_B2 0x7652u
#define PTM_SEL_B3 0x7653u

// Exact uint16 -> float the same way (selector takes two bytes).
__device__ __forceinline__ float half_to_float_lo(unsigned w) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7610u)) - 8388608.0f;
}
__device__ __forceinline__ float half_to_float_hi(unsigned w) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632u)) - 8388608.0f;
}

// x / d for x <= 255 * d, d < 4104, with magic = floor(2^32 / d) + 1 (0 means d == 1).
__device__ __forceinline__ unsigned div_magic(unsigned x, unsigned magic) {
    return magic ? __umulhi(x, magic) : x;
}

struct Extrema {
    float mn[PTM_NCOEF];
    float mx[PTM_NCOEF];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int k = 0; k < PTM_NCOEF; ++k) {
            mn[k] = FLT_MAX;
            mx[k] = -FLT_MAX;
        }
    }
    // fminf/fmaxf ignore NaNs like their libm namesakes (rti-builder/ptmlib.c:90-104)
    __device__ __forceinline__ void add(const float *c) {
#pragma unroll
        for (int k = 0; k < PTM_NCOEF; ++k) {
            mn[k] = fminf(mn[k], c[k]);
            mx[k] = fmaxf(mx[k], c[k]);
        }
    }
    // four pixels at once with the 3-input min / max of sm_100 (FMNMX3): 2 instructions per
    // coefficient and bound instead of 4.  Like fminf / fmaxf, NaN operands are ignored.
    __device__ __forceinline__ void add4(const float (&c)[4][PTM_NCOEF]) {
#pragma unroll
        for (int k = 0; k < PTM_NCOEF; ++k) {
            float a, b;
            asm("min.f32 %0, %1, %2, %3;" : "=f"(a) : "f"(mn[k]), "f"(c[0][k]), "f"(c[1][k]));
            asm("min.f32 %0, %1, %2, %3;" : "=f"(mn[k]) : "f"(a), "f"(c[2][k]), "f"(c[3][k]));
            asm("max.f32 %0, %1, %2, %3;" : "=f"(b) : "f"(mx[k]), "f"(c[0][k]), "f"(c[1][k]));
            asm("max.f32 %0, %1, %2, %3;" : "=f"(mx[k]) : "f"(b), "f"(c[2][k]), "f"(c[3][k]));
        }
    }
    // warp shuffle tree, then one atomicMax per warp and coefficient on the keys
    __device__ __forceinline__ void flush(int *keys, int *smem_keys /* [12] in shared memory */) {
        // block-level merge in shared memory first, then 12 global atomics per block
        if (threadIdx.x < 2 * PTM_NCOEF)
            smem_keys[threadIdx.x] = float_to_key(-FLT_MAX);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PTM_NCOEF; ++k) {
            // integer max of the order-preserving keys == float max; one REDUX per value
            const int a = __reduce_max_sync(0xFFFFFFFFu, float_to_key(-mn[k]));
            const int b = __reduce_max_sync(0xFFFFFFFFu, float_to_key(mx[k]));
            if ((threadIdx.x & 31) == 0) {
                atomicMax(&smem_keys[k], a);
                atomicMax(&smem_keys[PTM_NCOEF + k], b);
            }
        }
        __syncthreads();
        if (threadIdx.x < 2 * PTM_NCOEF)
            atomicMax(&keys[threadIdx.x], smem_keys[threadIdx.x]);
    }
};

// Called by every thread of every block right after Extrema::flush when a hand-off is set up.
// The block that takes the last ticket owns the merged keys: it swaps them out (resetting them
// for the next launch), stores them into slot [parity][rank] of every rank's mailbox, fences
// system-wide and then raises the flags.  Peer mailboxes are plain P2P pointers over NVLink.
template <typename Handoff, typename Layout>
__device__ __forceinline__ void keys_handoff(const Handoff &h, int *keys, int max_ranks, int slot_ints) {
    __shared__ int last_s;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0)
        last_s = (atomicAdd(h.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last_s)
        return;
    __threadfence();
    const int parity = h.epoch & 1;
    // the merged keys leave the accumulator (which is reset for the next launch) through shared memory;
    // lane r then fills slot [parity][rank] of rank r's mailbox, fences ONCE and raises that rank's flag
    __shared__ int merged_s[2 * PTM_NCOEF];
    if (threadIdx.x < 2 * PTM_NCOEF)
        merged_s[threadIdx.x] = atomicExch(&keys[threadIdx.x], float_to_key(-FLT_MAX));
    if (threadIdx.x == 0) {
        *h.ticket = 0;   // before the flags: once they are up the next fit may be on its way
        if (h.tile_counter)
            *h.tile_counter = 0;   // every block has stopped drawing tiles before it took its ticket
    }
    __syncthreads();
    if (threadIdx.x < h.world) {
        int *box = h.mailbox[threadIdx.x];
        int *slot = box + Layout::slots_off / sizeof(int) + (parity * max_ranks + h.rank) * slot_ints;
#pragma unroll
        for (int k = 0; k < 2 * PTM_NCOEF; ++k)
            slot[k] = merged_s[k];
        // one mailbox on this GPU: device scope is enough; peers over NVLink need system scope.  Thread 0's
        // resets above are ordered before the flag by the block barrier and this fence (cumulativity).
        if (h.world == 1)
            __threadfence();
        else
            __threadfence_system();
        volatile int *flags = box + Layout::flags_off / sizeof(int);
        flags[parity * max_ranks + h.rank] = (int) h.epoch;
    }
}

// K2 prologue: the 12 merged keys into shared memory, from plain keys or from the local mailbox
// (one lane per rank polls its flag, bounded by a timeout that is reported in the mailbox status).
template <typename Source, typename Layout>
__device__ __forceinline__ void keys_fetch(const Source &src, int *keys_sh, int max_ranks, int slot_ints) {
    __shared__ int ok_s;
    if (src.mailbox == nullptr) {
        if (threadIdx.x < 2 * PTM_NCOEF)
            keys_sh[threadIdx.x] = src.keys[threadIdx.x];
        __syncthreads();
        return;
    }
    const int parity = src.epoch & 1;
    const int *flags = src.mailbox + Layout::flags_off / sizeof(int);
    if (threadIdx.x == 0)
        ok_s = 1;
    __syncthreads();
    if (threadIdx.x < src.world) {
        // acquire load at system scope pairs with the publisher's fence + flag store; the block
        // barrier below extends the ordering to the lanes that read the slots (no fence per CTA)
        const int *f = flags + parity * max_ranks + threadIdx.x;
        unsigned long long t0 = 0, t;
        int v;
        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
        if (v != (int) src.epoch) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            do {
                __nanosleep(64);
                asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                if (t - t0 > 5000000000ull) {
                    ok_s = 0;
                    break;
                }
            } while (v != (int) src.epoch);
        }
    }
    __syncthreads();
    if (threadIdx.x < 2 * PTM_NCOEF) {
        const volatile int *slots = src.mailbox + Layout::slots_off / sizeof(int);
        int m = slots[(parity * max_ranks + 0) * slot_ints + threadIdx.x];
        for (int r = 1; r < src.world; ++r)
            m = max(m, slots[(parity * max_ranks + r) * slot_ints + threadIdx.x]);
        keys_sh[threadIdx.x] = m;
    }
    if (threadIdx.x == 0 && !ok_s)
        src.mailbox[Layout::status_off / sizeof(int)] = 1;
    __syncthreads();
}

// The LRGB colour + normalise step for one pixel (rti-builder/ptm-encoder.c:331-348).
// (ya, cba, cra) are the truncated integer averages of rti-builder/ptmlib.c:794-796.
__device__ __forceinline__ void lrgb_finish(unsigned ya, unsigned cba, unsigned cra, float *c,
                                            unsigned &r, unsigned &g, unsigned &b) {
    const float y = (float) ya;
    const float cb = (float) ((int) cba - 128);
    const float cr = (float) ((int) cra - 128);
    r = clip_u8(__fadd_rn(y, __fmul_rn(1.40200f, cr)));
    g = clip_u8(__fsub_rn(__fsub_rn(y, __fmul_rn(0.34414f, cb)), __fmul_rn(0.71414f, cr)));
    b = clip_u8(__fadd_rn(y, __fmul_rn(1.77200f, cb)));
    const float norm = __fdiv_rn(256.0f, y);
#pragma unroll
    for (int k = 0; k < PTM_NCOEF; ++k)
        c[k] = __fmul_rn(c[k], norm);
}

// The LUM variant of the same step (rti-builder/ptm-encoder.c:286-305): the averaged YCbCr triplet is
// stored as it is and the coefficients stay unnormalised.
__device__ __forceinline__ void lrgb_or_lum_finish(bool lum, unsigned ya, unsigned cba, unsigned cra, float *c,
                                                   unsigned &r, unsigned &g, unsigned &b) {
    if (lum) {
        r = ya;
        g = cba;
        b = cra;
    } else {
        lrgb_finish(ya, cba, cra, c, r, g, b);
    }
}

// RGB -> YCbCr per sample, libjpeg's integer JFIF conversion (jccolor.c rgb_ycc_convert, third party; the
// reference itself never converts -- its decoder delivers YCbCr, rti-builder/ptm-encoder.c:188 -- so this is the
// optional front end for RGB-input stacks; oracle/ptm_oracle.c:orc_rgb_to_ycbcr is the
// same arithmetic on the host):
//   Y  = ( 19595 R + 38470 G +  7471 B + 32768) >> 16
//   Cb = (-11059 R - 21709 G + 32768 B + 8388608 + 32767) >> 16
//   Cr = ( 32768 R - 27439 G -  5329 B + 8388608 + 32767) >> 16         (all sums are positive)
__device__ __forceinline__ void rgb_to_ycc(unsigned r, unsigned g, unsigned b, unsigned &y, unsigned &cb, unsigned &cr) {
    y = (19595u * r + 38470u * g + 7471u * b + 32768u) >> 16;
    cb = (8388608u + 32767u + 32768u * b - 11059u * r - 21709u * g) >> 16;
    cr = (8388608u + 32767u + 32768u * r - 27439u * g - 5329u * b) >> 16;
}

// the same on a lane's 12 bytes = 4 interleaved pixels held in three words
__device__ __forceinline__ void rgb_words_to_ycc(unsigned &w0, unsigned &w1, unsigned &w2) {
    const unsigned in[3] = {w0, w1, w2};
    unsigned out[12];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int j = 3 * p;
        const unsigned r = (in[j >> 2] >> (8 * (j & 3))) & 0xFFu;
        const unsigned g = (in[(j + 1) >> 2] >> (8 * ((j + 1) & 3))) & 0xFFu;
        const unsigned b = (in[(j + 2) >> 2] >> (8 * ((j + 2) & 3))) & 0xFFu;
        rgb_to_ycc(r, g, b, out[j], out[j + 1], out[j + 2]);
    }
    w0 = out[0] | (out[1] << 8) | (out[2] << 16) | (out[3] << 24);
    w1 = out[4] | (out[5] << 8) | (out[6] << 16) | (out[7] << 24);
    w2 = out[8] | (out[9] << 8) | (out[10] << 16) | (out[11] << 24);
}

// scale / bias / inverse scale from the extrema (rti-builder/ptmlib.c:858-863).
__device__ __forceinline__ void scale_bias_from_keys(const int *keys, int k, float &scale, int &bias, float &inv) {
    const float mn = -key_to_float(keys[k]);
    const float mx = key_to_float(keys[PTM_NCOEF + k]);
    const float range = __fsub_rn(mx, mn);
    scale = __fdiv_rn(range, 256.0f);
    // (int) of a NaN or out-of-range float is undefined in C; the reference runs on x86, where the
    // conversion (cvttss2si) then yields INT_MIN -- e.g. for a constant image (range == 0).  CUDA's
    // conversion would saturate / return 0 instead, so the x86 result is spelled out.
    const float fb = __fmul_rn(__fdiv_rn(-256.0f, range), mn);
    bias = (fb >= -2147483648.0f && fb < 2147483648.0f) ? (int) fb : (int) 0x80000000;
    inv = __fdiv_rn(1.0f, scale);
}

}  // namespace ptm
