// K3 for the RGB formats on the tensor cores -- relighting sweeps, tolerance form (+-1 per byte).
//
// Per output byte (one channel of one pixel) and direction d the reference evaluates
//     CLIP( sum_k  (byte_k - bias_k) * scale_k * l_k(d) )          rti-builder/ptmlib.c:553-562,610-619
// -- a GEMM with K = 6: [raster bytes x 6] . [6 x directions].  The FFMA form of it
// (relight_rgb_fast_kernel) is bound by the FMA pipe: 18 multiply-adds per pixel-direction for 3 bytes of
// output (64 % pipe utilisation at 0.62 of the HBM roofline, profiles/r02_ncu_relight_rgb_fast_summary.txt).
// Here the multiply-adds go to the tensor cores and the CUDA cores only convert and pack:
//
//   * B operand = (byte_k - bias_k): small integers, EXACT in f16 (|value| < 2048).  Built once per 16 output
//     bytes and kept in registers for the whole sweep (one 32-bit register per 8 output bytes and lane).
//   * A operand = scale_k * l_k(d) as a TWO-term f16 split (hi + lo; K = 2 x 6 of 16), one table row per
//     direction in shared memory.  Products are exact, accumulation is fp32 in the tensor core.  The split
//     leaves |x| 2^-22 of each term: with |scale_k l_k| <= 64 (checked on the host, else the FFMA kernel
//     serves the call) the sum is within 0.05 of the fused fp32 value -- only bytes whose value lies that
//     close to an integer can differ from the reference, by one.
//   * mma.sync.m16n8k8 (one step for the hi, one for the lo terms) with M = 16 DIRECTIONS and N = 8 OUTPUT BYTES: the accumulator fragment then holds,
//     per lane, two ADJACENT bytes of one raster (columns 2t, 2t+1).  The columns of two n-tiles are
//     interleaved (tile 0: bytes 4j, 4j+1; tile 1: bytes 4j+2, 4j+3) so a lane ends up with four consecutive
//     bytes of a direction: 2 FMUL2.RZ (denormal trick, see ptm_relight_fast.cu) + 2 I2IP make the word.
//     This is why the legacy register-fragment MMA is used and not tcgen05: a TMEM accumulator puts adjacent
//     raster bytes in different LANES (lane = row), which would cost a cross-lane byte transpose per word,
//     and the GEMM itself is tiny (K = 16, 3 bytes of HBM per 32 flops) -- the kernel is bound by the
//     raster writes, not by tensor throughput.
//   * a warp owns 512 consecutive output bytes (32 sixteen-byte chunks, 512-byte aligned); the words of 16
//     directions are staged in shared memory (row pitch 528 B: conflict free) and leave as 16-byte stores, one
//     store instruction = the 512 contiguous bytes of one direction.
#include <cstdlib>

#include <cuda_fp16.h>

#include "ptm_device.cuh"
#include "ptm_kernels.h"

namespace ptm {

namespace {

// One m16n8k16 step instead: half the tensor instructions (the legacy HMMA path issues one instruction per 8 cycles
// and sub-partition whatever its K), but B must sit in a register PAIR (b, b2 hold the same value).
__device__ __forceinline__ void hmma16(float (&d)[4], const unsigned (&a)[4], unsigned b, unsigned b2) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
                 "{%10, %10, %10, %10};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b), "r"(b2), "f"(0.0f));
}

constexpr int kGroup = 4;
// staging row pitch in bytes: unit + 16 (NC = 24: 100 words, lanes (g, t) hit banks 4g + t; same for 16 and 32)
__host__ __device__ constexpr int row_pitch(int nc) { return nc * 16 + 16; }
__host__ __device__ constexpr int stage_bytes(int nc) { return 16 * row_pitch(nc); }   // 16 directions per m-tile
constexpr float kTwoM149 = 1.40129846e-45f;

typedef unsigned long long f32x2;

__device__ __forceinline__ f32x2 pk(float lo, float hi) {
    f32x2 d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ void mul2_rz_bits(f32x2 a, f32x2 b, int &lo, int &hi) {
    f32x2 d;
    asm("mul.rz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(d));
}
__device__ __forceinline__ unsigned pack_sat_u8(int hi, int lo, unsigned upper) {
    unsigned d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(hi), "r"(lo), "r"(upper));
    return d;
}
// D(16 x 8, f32) = Ahi(16 x 8) * B(8 x 8) + Alo(16 x 8) * B(8 x 8): two m16n8k8 steps on the SAME B register.  (One
// m16n8k16 with B repeated in both K halves needs the B value in a register PAIR: twice the registers for B, and
// ptxas spilled them into the inner loop.)
__device__ __forceinline__ void hmma(float (&d)[4], const unsigned (&a)[4], unsigned b, unsigned) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%7, %7, %7, %7};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a[2]), "r"(a[3]), "r"(b), "f"(0.0f));
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(b));
}

// direction table: per direction 16 halves = 8 words: words 0..2 hi(x_0,x_1) hi(x_2,x_3) hi(x_4,x_5), word 3 zero,
// words 4..6 the lo parts, word 7 zero; x_k = scale_k * l_k(d)
__device__ __forceinline__ void build_table(uint32_t *tab, const RelightArgs &a, unsigned d_pad) {
    for (unsigned i = threadIdx.x; i < d_pad * 4; i += blockDim.x) {
        const unsigned d = i >> 2, w = i & 3;
        uint32_t hi = 0, lo = 0;
        if (w < 3 && d < a.n_dirs) {
            const float s0 = w == 0 ? a.scale[0] : w == 1 ? a.scale[2] : a.scale[4];
            const float s1 = w == 0 ? a.scale[1] : w == 1 ? a.scale[3] : a.scale[5];
            const float x0 = s0 * a.light[d * PTM_NCOEF + 2 * w];
            const float x1 = s1 * a.light[d * PTM_NCOEF + 2 * w + 1];
            const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
            const __half l0 = __float2half_rn(x0 - __half2float(h0)), l1 = __float2half_rn(x1 - __half2float(h1));
            hi = (uint32_t) __half_as_ushort(h0) | ((uint32_t) __half_as_ushort(h1) << 16);
            lo = (uint32_t) __half_as_ushort(l0) | ((uint32_t) __half_as_ushort(l1) << 16);
        }
        tab[d * 8 + w] = hi;
        tab[d * 8 + 4 + w] = lo;
    }
}

// kChunks 16-byte chunks per warp unit, kWarps warps per CTA, kMinCtas CTAs per SM (the register budget)
template <int kChunks, int kWarps, int kMinCtas, bool kK16 = false>
__global__ void __launch_bounds__(kWarps * 32, kMinCtas)
relight_rgb_mma_kernel(RelightArgs a, unsigned d_pad, unsigned long long n_chunks_total) {
    constexpr int kRowPitch = row_pitch(kChunks), kStageBytes = stage_bytes(kChunks);
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t *tab = reinterpret_cast<uint32_t *>(smem + kWarps * kStageBytes);
    build_table(tab, a, d_pad);
    __syncthreads();

    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned g = lane >> 2, t = lane & 3;
    uint8_t *stage = smem + warp * kStageBytes;
    const unsigned row_bytes = (unsigned) (a.width * 3);   // a multiple of 16; the raster is < 4 GiB (launch conditions)
    const size_t raster = (size_t) row_bytes * a.height;
    const f32x2 kDen = pk(kTwoM149, kTwoM149);
    // 1024 + bias of this lane's two coefficients, as f16x2 (exact: |bias| <= 1024)
    const int bias_lo = t == 0 ? a.bias[0] : t == 1 ? a.bias[2] : a.bias[4];
    const int bias_hi = t == 0 ? a.bias[1] : t == 1 ? a.bias[3] : a.bias[5];
    const __half2 kb2 = __floats2half2_rn(1024.0f + (float) bias_lo, 1024.0f + (float) bias_hi);
    // the lane supplies column g of both n-tiles of a chunk: output bytes 4 (g / 2) + g % 2 and + 2
    const unsigned col0 = 4 * (g >> 1) + (g & 1);

    const unsigned long long n_units = (n_chunks_total + kChunks - 1) / kChunks;
    for (unsigned long long unit = (unsigned long long) blockIdx.x * kWarps + warp; unit < n_units;
         unit += (unsigned long long) gridDim.x * kWarps) {
        const unsigned long long c0 = unit * kChunks;
        const unsigned nch = (unsigned) min((unsigned long long) kChunks, n_chunks_total - c0);
        // ---- B operand: (byte - bias) of coefficients (2t, 2t+1) for the lane's two output bytes of every chunk
        unsigned B[kChunks][2], B2[kK16 ? kChunks : 1][2];
        // flat output byte of the unit's first chunk -> output row y, byte xb within the row; chunks never straddle rows
        unsigned y = (unsigned) ((c0 * 16) / row_bytes);
        unsigned xb = (unsigned) (c0 * 16 - (unsigned long long) y * row_bytes);
        const unsigned tt = t < 3 ? t : 2;            // lanes t = 3 hold the K padding: they load (harmlessly) and keep 0
        const unsigned keep = t < 3 ? 0xFFFFFFFFu : 0u;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
            // branch free: chunks past the end of the image re-read the unit's last valid chunk (never stored)
            const unsigned s0 = ((unsigned) a.height - 1u - y) * row_bytes + xb;   // stored row height-1-y (:587-589)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const unsigned s = s0 + col0 + 2 * h;
                const unsigned p = s / 3u;
                const unsigned ch = s - 3u * p;
                const uint8_t *plane = ch == 0 ? a.plane[0] : ch == 1 ? a.plane[1] : a.plane[2];
                const unsigned v = __ldg(reinterpret_cast<const unsigned short *>(plane + (size_t) p * PTM_NCOEF) + tt);
                const unsigned w = __byte_perm(v, 0x64646464u, 0x4140);   // halves 1024 + byte
                const __half2 r = __hsub2(*reinterpret_cast<const __half2 *>(&w), kb2);
                B[c][h] = *reinterpret_cast<const unsigned *>(&r) & keep;
                if (kK16)   // a second physical register with the same value (opaque to the optimiser)
                    asm volatile("mov.b32 %0, %1;" : "=r"(B2[c][h]) : "r"(B[c][h]));
            }
            if ((unsigned) (c + 1) < nch) {
                xb += 16;
                if (xb >= row_bytes) {
                    xb -= row_bytes;
                    ++y;
                }
            }
        }
        uint8_t *out_unit = a.out + c0 * 16;
        for (unsigned d0 = 0; d0 < d_pad; d0 += 16) {
            unsigned A[4];
            A[0] = tab[(d0 + g) * 8 + t];
            A[1] = tab[(d0 + g + 8) * 8 + t];
            A[2] = tab[(d0 + g) * 8 + 4 + t];
            A[3] = tab[(d0 + g + 8) * 8 + 4 + t];
            uint32_t *st0 = reinterpret_cast<uint32_t *>(stage + g * kRowPitch) + t;
            uint32_t *st1 = reinterpret_cast<uint32_t *>(stage + (g + 8) * kRowPitch) + t;
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
                {   // chunks past the end of the image have B = 0: computed, staged, never stored
                    float e[4], f[4];
                    if (kK16) {
                        hmma16(e, A, B[c][0], B2[c][0]);
                        hmma16(f, A, B[c][1], B2[c][1]);
                    } else {
                        hmma(e, A, B[c][0], 0u);   // rows g, g+8; bytes 4t, 4t+1
                        hmma(f, A, B[c][1], 0u);   //             bytes 4t+2, 4t+3
                    }
                    int z0, z1, z2, z3;
                    mul2_rz_bits(pk(e[0], e[1]), kDen, z0, z1);
                    mul2_rz_bits(pk(f[0], f[1]), kDen, z2, z3);
                    st0[4 * c] = pack_sat_u8(z1, z0, pack_sat_u8(z3, z2, 0u));
                    mul2_rz_bits(pk(e[2], e[3]), kDen, z0, z1);
                    mul2_rz_bits(pk(f[2], f[3]), kDen, z2, z3);
                    st1[4 * c] = pack_sat_u8(z1, z0, pack_sat_u8(z3, z2, 0u));
                }
                // keep at most kGroup chunks (2 kGroup HMMAs, 8 kGroup accumulators) in flight: left alone the
                // scheduler interleaves all of them and spills the B registers into the inner loop
                if (c % kGroup == kGroup - 1)
                    asm volatile("" ::: "memory");
            }
            __syncwarp();
            // one store instruction per direction: lanes 0..nch-1 write its 16-byte chunks (384 contiguous bytes)
            if (lane < nch) {
                uint8_t *dst = out_unit + (size_t) d0 * raster + lane * 16;
                const unsigned rows = min(16u, a.n_dirs - d0);
#pragma unroll 4
                for (unsigned r = 0; r < rows; ++r, dst += raster)
                    __stcs(reinterpret_cast<uint4 *>(dst), *reinterpret_cast<const uint4 *>(stage + r * kRowPitch + lane * 16));
            }
            __syncwarp();
        }
    }
}

}  // namespace

namespace {
template <int kChunks, int kWarps, int kMinCtas, bool kK16 = false>
cudaError_t launch_chunks(const RelightArgs &a, int sm_count, cudaStream_t st) {
    constexpr int kThreads = kWarps * 32;
    const unsigned d_pad = (a.n_dirs + 15u) & ~15u;
    const size_t smem = (size_t) kWarps * stage_bytes(kChunks) + (size_t) d_pad * 32;
    if (smem > 110 * 1024)
        return cudaErrorNotSupported;
    static bool attr_done[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess)
        return e;
    if (dev >= 0 && dev < 64 && !attr_done[dev]) {
        e = cudaFuncSetAttribute(relight_rgb_mma_kernel<kChunks, kWarps, kMinCtas, kK16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
        if (e != cudaSuccess)
            return e;
        attr_done[dev] = true;
    }
    const unsigned long long n_chunks = (unsigned long long) a.width * a.height * 3 / 16;
    const unsigned long long n_units = (n_chunks + kChunks - 1) / kChunks;
    const unsigned long long want = (n_units + kWarps - 1) / kWarps;
    const unsigned long long cap = (unsigned long long) sm_count * kMinCtas;
    const unsigned grid = (unsigned) (want < cap ? want : cap);
    relight_rgb_mma_kernel<kChunks, kWarps, kMinCtas, kK16><<<grid, kThreads, smem, st>>>(a, d_pad, n_chunks);
    return cudaGetLastError();
}
}  // namespace

// cudaErrorNotSupported: layout or value range outside what the f16 operands take (the caller falls back to the
// FFMA tolerance kernel).  max_term = max over directions and coefficients of |scale_k * l_k(d)| (host-computed).
cudaError_t launch_relight_rgb_mma(const RelightArgs &a, float max_term, int sm_count, cudaStream_t st) {
    const size_t pixels = a.width * a.height;
    if (pixels == 0 || a.n_dirs == 0)
        return cudaSuccess;
    bool ok = a.width % 16 == 0 && ((uintptr_t) a.out % 16 == 0) && max_term <= 64.0f && max_term == max_term &&
              pixels * 3 < 0xFFFFFFFFull;
    for (int k = 0; k < PTM_NCOEF; ++k)
        ok = ok && a.bias[k] >= -1024 && a.bias[k] <= 1024;
    for (int p = 0; p < 3; ++p)
        ok = ok && ((uintptr_t) a.plane[p] % 2 == 0);
    if (!ok)
        return cudaErrorNotSupported;
    static int shape = -1;
    if (shape < 0) {
        const char *v = getenv("PTM_RELIGHT_MMA_SHAPE");   // tuning aid (result-equivalent kernel shapes)
        shape = v ? atoi(v) : 0;
    }
    // Measured on B200, 24 MP x 360 directions (profiles/r02_relight_rgb_mma.txt): units of 32 chunks (512 bytes,
    // aligned) 4.89-4.99 ms in every CTA shape; 16 chunks 5.04-5.28 ms; 24 chunks (384-byte units that straddle the
    // 256-byte granules of the memory system) 6.0-6.3 ms; one m16n8k16 step with B in register pairs 5.04-5.16 ms.
    switch (shape) {
    case 1: return launch_chunks<32, 4, 3>(a, sm_count, st);
    case 2: return launch_chunks<16, 8, 2>(a, sm_count, st);
    case 3: return launch_chunks<24, 8, 2>(a, sm_count, st);
    case 4: return launch_chunks<16, 6, 2, true>(a, sm_count, st);
    default: return launch_chunks<32, 2, 6>(a, sm_count, st);
    }
}

}  // namespace ptm
