// K1, tensor-core form -- the fit as a byte-wise integer GEMM on the sm_100a tensor cores.
//
// Every byte of the stack is one row of a GEMM whose K dimension is the light index:
//
//     D[byte g][column j] = sum_n stack[n][g] * B[n][j]            (u8 x s8 -> s32, tcgen05.mma kind::i8)
//
// B holds the pseudo-inverse as 32-bit fixed point, split into four balanced base-256 digits
// per coefficient (columns 4k .. 4k+3), plus one column of ones (column 24) that yields the
// per-byte sums ptm_cbcr_avg needs.  The integer dot products are exact, so
//
//     sum_n sample_n * q_kn = D0 * 2^24 + D1 * 2^16 + D2 * 2^8 + D3        (q = round(M * 2^F_k))
//
// is the exact dot product with a matrix that differs from M by < 2^-31 of its largest entry,
// and one int64 -> float conversion rounds it once.  (The reference leaves the summation order
// to its BLAS, rti-builder/ptmlib.c:717-720; this result is order independent.)  16-bit samples
// are two byte rows, combined as lo + 256 * hi in integers.
//
// Why tensor cores here: the RGB formats need 18 multiply-adds per byte of the stack, the FFMA
// kernels are issue bound on them (profiles/r01_formats.txt, 39 % of the HBM roofline), and
// north_star allows tensor cores where ncu shows the FFMA pipe as the limiter.  The raw bytes go
// from HBM to shared memory by TMA (SWIZZLE_128B boxes of 128 bytes x all lights) and are
// consumed by the tensor core AS THEY LIE (MN-major unsigned 8-bit A operand); the CUDA cores
// only run the epilogue.
//
//   warp 0      : TMA producer, ring of S stages, one stage = 3 column blocks of 128 bytes x n lights
//   warp 1      : allocates TMEM, issues tcgen05.mma (3 blocks x kpad/32 instructions per unit) into
//                 one of 4 accumulator buffers (3 x 32 columns each), commits to the ring/TMEM barriers
//   warps 2..   : G epilogue groups of 4 warps; a lane owns TMEM lane r of its quarter = byte r of
//                 each column block.  128 = 2 (mod 3), so of a lane's three bytes (one per block)
//                 exactly one is a luma byte: every lane finishes exactly one LRGB pixel per unit.
//                 Outputs are staged in shared memory and leave as bulk stores (cp.async.bulk s2g).
//
// Reference arithmetic fused here: rti-builder/ptmlib.c:692-760 (fit), :762-802 (average),
// rti-builder/ptm-encoder.c:272-284 (RGB planes), :327-353 (colour + normalise),
// rti-builder/ptmlib.c:826-847 (extrema).
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cmath>
#include <cstdlib>
#include <cstring>

#include "ptm_device.cuh"
#include "ptm_kernels.h"
#include "ptm_tma.cuh"

namespace ptm {

namespace {

constexpr int kBlk = 128;                 // bytes per column block = MMA M
constexpr int kUnitBlocks = 3;            // 384 bytes: whole pixels for 3- and 6-byte pixels
constexpr int kUnitBytes = kBlk * kUnitBlocks;
constexpr int kNB = 32;                   // MMA N: 24 digit columns + ones + padding
constexpr int kSumCol = 24;
constexpr int kAccBufs = 4;               // TMEM accumulator buffers
constexpr int kAccCols = 128;             // columns reserved per buffer (96 used)
constexpr int kTmemCols = kAccBufs * kAccCols;
constexpr int kMaxStages = 12;
constexpr int kMaxDynSmem = 226 * 1024;   // 227 KB opt-in limit minus the kernel's static shared memory

// instruction descriptor (cute/arch/mma_sm100_desc.hpp field layout): D = s32, A = u8 MN-major,
// B = s8 K-major, N = 32, M = 128
constexpr uint32_t kIdesc = (2u << 4) | (0u << 7) | (1u << 10) | (1u << 15) | (0u << 16) |
                            ((uint32_t) (kNB >> 3) << 17) | ((uint32_t) (kBlk >> 4) << 24);

// staging per epilogue group: RGB = 3 planes x 128 px x 24 B; LRGB = 128 x 24 + 384 (rgb) + 384 (means)
constexpr int kStagePlane = 128 * 24;
constexpr int kStageBytes = 3 * kStagePlane;

struct MmaArgs {
    const int8_t *digits;     // device, shared-memory image of B: [kpad/16][4][8][16]
    float scale[PTM_NCOEF];   // 2^-F_k
    unsigned kpad;            // lights rounded up to 32
    unsigned stages;
    unsigned n_units;
    unsigned tail_px;         // pixels of the last unit when it is partial, else 0
    unsigned row_bytes;       // bytes of one image that the units cover
    unsigned grouped;         // 1: whole units are loaded light-group-major (see the producer)
    unsigned store_hint;      // L2 policy of the output bulk stores: 0 = none, 1 = evict_first, 2 = evict_last
    unsigned run_units;       // units per run: a CTA draws runs of consecutive units (DRAM row locality inside a run)
    unsigned *counter;        // device-wide run counter, 0 at launch (nullptr: runs strided by the grid)
};

__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t) ((saddr & 0x3FFFFu) >> 4) | ((uint64_t) (lbo >> 4) << 16) | ((uint64_t) (sbo >> 4) << 32) |
           (1ull << 46) | ((uint64_t) layout << 61);
}

__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate)
        : "memory");
}

// one lane of a converged warp (the issue code stays warp-uniform, so descriptors live in uniform registers)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xFFFFFFFF;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar))
                 : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 consecutive columns of this lane's TMEM row
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, int (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar,
                                            uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_addr(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_addr(bar)), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}

__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar,
                                            uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(smem_addr(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_addr(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
        : "memory");
}

__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, int c3,
                                            uint64_t *bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;" ::"r"(smem_addr(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_addr(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(policy)
        : "memory");
}

__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(smem_addr(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g_hint(void *dst_gmem, const void *src_smem, uint32_t bytes, uint64_t policy) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst_gmem),
                 "r"(smem_addr(src_smem)), "r"(bytes), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void store_out(unsigned hint, uint64_t policy, void *dst_gmem, const void *src_smem,
                                          uint32_t bytes) {
    if (hint)
        bulk_s2g_hint(dst_gmem, src_smem, bytes, policy);
    else
        bulk_s2g(dst_gmem, src_smem, bytes);
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void group_sync(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// exact sum_n sample_n * q_n of one byte row from its four digit sums
__device__ __forceinline__ long long digits_to_int(const int *d) {
    const int hi = d[0] * 256 + d[1];
    const int lo = d[2] * 256 + d[3];
    return (long long) hi * 65536 + (long long) lo;
}

}  // namespace

// kMode: 0 = RGB formats, uint8; 1 = LRGB, uint8; 2 = RGB formats, uint16; 3 = LRGB, uint16
// kDebug (only instantiated in -DPTM_TUNING builds): 1 = epilogue only drains TMEM (load + MMA ceiling),
// 2 = additionally no MMAs are issued (TMA delivery ceiling), 3 = MMAs and drain but no loads (MMA ceiling)
template <int kMode, int G, int kDebug = 0>
__global__ void __launch_bounds__(64 + 128 * G, 1)
fit_mma_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_grp,
               const __grid_constant__ CUtensorMap tmap_rem, FitArgs a, MmaArgs m) {
    constexpr bool kLrgb = (kMode & 1) != 0;
    constexpr bool k16 = kMode >= 2;
    constexpr unsigned kUnitPx = k16 ? 64 : 128;

    extern __shared__ uint8_t smem_raw[];
    __shared__ int keys_s[2 * PTM_NCOEF];
    uint8_t *smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
    const unsigned stage_bytes = kUnitBytes * m.kpad;
    const unsigned blk_bytes = kBlk * m.kpad;
    uint8_t *ring = smem;
    uint8_t *bdig = ring + (size_t) m.stages * stage_bytes;
    uint8_t *stg_all = bdig + (size_t) m.kpad * 32;
    uint64_t *full = reinterpret_cast<uint64_t *>(stg_all + (size_t) G * kStageBytes);
    uint64_t *empty = full + m.stages;
    uint64_t *tfull = empty + m.stages;
    uint64_t *tempty = tfull + kAccBufs;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + kAccBufs);

    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // a programmatically dependent K2 (the fused encode) may become resident as SMs drain
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x == 0) {
        for (unsigned s = 0; s < m.stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < kAccBufs; ++b) {
            mbar_init(&tfull[b], 1);
            mbar_init(&tempty[b], 4);
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(tmem_slot)),
                     "r"((uint32_t) kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {   // B operand image into shared memory (generic proxy), then visible to the tensor core
        const uint4 *src = reinterpret_cast<const uint4 *>(m.digits);
        uint4 *dst = reinterpret_cast<uint4 *>(bdig);
        for (unsigned i = threadIdx.x; i < m.kpad * 2; i += blockDim.x)
            dst[i] = src[i];
        fence_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    Extrema ext;
    ext.init();

    // Schedule: a CTA works on RUNS of consecutive units (the ring then holds several adjacent 384-byte pieces
    // of every light: more DRAM row hits).  Its first run is blockIdx.x, the following ones are drawn from a
    // device-wide counter by the producer (CTAs on faster SMs take more; with a fixed split the slowest CTA set
    // the kernel time).  The producer announces each unit through unit_pub[local index % 32] before it arms
    // the stage's barrier; kNoUnit tells the MMA warp and, through it, every epilogue group to stop.
    __shared__ unsigned unit_pub[32];
    constexpr unsigned kNoUnit = 0xFFFFFFFFu;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        // (the whole warp runs the loop converged; one elected lane issues)
        const uint64_t policy = l2_policy_evict_first();
        const unsigned groups = a.n_lights / 8, rem = a.n_lights % 8;
        unsigned s = 0, ph = 0, i = 0;
        unsigned run = blockIdx.x;
        while ((size_t) run * m.run_units < m.n_units) {
            unsigned next = run + gridDim.x;
            if (m.counter) {   // draw the next run now: the atomic's latency hides behind this run's copies
                if (lane == 0)
                    next = gridDim.x + atomicAdd(m.counter, 1u);
                next = __shfl_sync(0xFFFFFFFFu, next, 0);
            }
            const unsigned u0 = run * m.run_units, u1 = min(m.n_units, u0 + m.run_units);
            for (unsigned u = u0; u < u1; ++u, ++i) {
                const unsigned byte0 = u * kUnitBytes;
                const unsigned left = m.row_bytes - byte0;
                const unsigned nblk = left >= (unsigned) kUnitBytes ? kUnitBlocks : (left + kBlk - 1) / kBlk;
                mbar_wait(&empty[s], ph ^ 1);
                uint8_t *dst = ring + (size_t) s * stage_bytes;
                if (elect_one()) {
                    unit_pub[i & 31] = u;
                    if (kDebug == 3) {
                        mbar_arrive(&full[s]);
                    } else if (m.grouped && left >= (unsigned) kUnitBytes) {
                        // Whole unit, light-group-major: one box = [8-light group][column block][8 lights][128 B],
                        // so the three 128-byte pieces of a light are requested 8 pieces apart (DRAM row hits)
                        // instead of a whole light set apart.  Stage layout [group][block][8 x 128 B swizzle atom].
                        mbar_arrive_expect_tx(&full[s], (groups + (rem ? 1 : 0)) * (kUnitBytes * 8));
                        if (groups)
                            tma_load_4d(dst, &tmap_grp, 0, 0, (int) (u * kUnitBlocks), 0, &full[s], policy);
                        if (rem)
                            tma_load_3d(dst + (size_t) groups * (kUnitBytes * 8), &tmap_rem, 0, 0,
                                        (int) (u * kUnitBlocks), &full[s], policy);
                    } else {
                        mbar_arrive_expect_tx(&full[s], nblk * kBlk * a.n_lights);
                        for (unsigned c = 0; c < nblk; ++c)
                            tma_load_2d(dst + (size_t) c * blk_bytes, &tmap, (int) (byte0 + c * kBlk), 0, &full[s],
                                        policy);
                    }
                }
                __syncwarp();
                if (++s == m.stages) {
                    s = 0;
                    ph ^= 1;
                }
            }
            run = next;
        }
        // one stop mark per epilogue group
        for (int e = 0; e < G; ++e, ++i) {
            mbar_wait(&empty[s], ph ^ 1);
            if (elect_one()) {
                unit_pub[i & 31] = kNoUnit;
                mbar_arrive(&full[s]);
            }
            __syncwarp();
            if (++s == m.stages) {
                s = 0;
                ph ^= 1;
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        // (whole warp converged, one elected lane issues; descriptors advance by adding to their low word)
        const uint32_t ring_a = smem_addr(ring);
        const unsigned ksteps = m.kpad / 32;
        const uint64_t bdesc0 = smem_desc(smem_addr(bdig), 512, 128, 0);
        unsigned s = 0, ph = 0, stops = 0;
        for (unsigned i = 0; stops < (unsigned) G; ++i) {
            const unsigned b = i % kAccBufs, use = i / kAccBufs;
            mbar_wait(&tempty[b], (use & 1) ^ 1);
            mbar_wait(&full[s], ph);
            tc_fence_after();
            const unsigned u = *reinterpret_cast<volatile unsigned *>(&unit_pub[i & 31]);
            if (u == kNoUnit) {
                // pass the stop mark on: the commit arrives once every earlier MMA has completed
                ++stops;
                if (elect_one()) {
                    mma_commit(&empty[s]);   // the producer may need this stage for a later stop mark
                    mma_commit(&tfull[b]);
                }
                __syncwarp();
                if (++s == m.stages) {
                    s = 0;
                    ph ^= 1;
                }
                continue;
            }
            const unsigned left = m.row_bytes - u * kUnitBytes;
            const unsigned nblk = left >= (unsigned) kUnitBytes ? kUnitBlocks : (left + kBlk - 1) / kBlk;
            if (elect_one()) {
                // A: 32 lights = 4 swizzle atoms of 8 rows x 128 B (SBO = atom stride); B: two 16-byte K chunks
                // 512 B apart (LBO), 8-column groups 128 B apart (SBO), 1024 B per 32 lights
                const bool grp = m.grouped && left >= (unsigned) kUnitBytes;   // whole unit
                const uint32_t atom_stride = grp ? kUnitBytes * 8 : 1024;   // between 8-light swizzle atoms
                const uint32_t blk_stride = grp ? 1024u : blk_bytes;        // between column blocks of a stage
                const uint64_t adesc0 = smem_desc(ring_a + s * stage_bytes, atom_stride, atom_stride, 2);
                const uint32_t a_step = (4 * atom_stride) >> 4;
                const uint32_t d0 = tmem_base + b * kAccCols;
                if (kDebug != 2) {
                    if (nblk == kUnitBlocks) {
#pragma unroll
                        for (int c = 0; c < kUnitBlocks; ++c) {
                            uint64_t ad = adesc0 + ((c * blk_stride) >> 4), bd = bdesc0;
                            mma_i8(d0 + c * kNB, ad, bd, 0);
                            for (unsigned kk = 1; kk < ksteps; ++kk) {
                                ad += a_step;
                                bd += 1024 >> 4;
                                mma_i8(d0 + c * kNB, ad, bd, 1);
                            }
                        }
                    } else {
                        for (unsigned c = 0; c < nblk; ++c) {
                            uint64_t ad = adesc0 + ((c * blk_stride) >> 4), bd = bdesc0;
                            for (unsigned kk = 0; kk < ksteps; ++kk) {
                                mma_i8(d0 + c * kNB, ad, bd, kk != 0);
                                ad += a_step;
                                bd += 1024 >> 4;
                            }
                        }
                    }
                }
                mma_commit(&empty[s]);   // the stage may be refilled once these MMAs have read it
                mma_commit(&tfull[b]);   // ... and the accumulators are complete
            }
            __syncwarp();
            if (++s == m.stages) {
                s = 0;
                ph ^= 1;
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue groups
        const unsigned g = (warp - 2) / 4;
        const unsigned quarter = warp & 3;               // TMEM lanes this warp may read
        const unsigned r = quarter * 32 + lane;          // byte row inside a column block
        const bool issuer = (warp - 2) % 4 == 0 && lane == 0;
        uint8_t *stg = stg_all + (size_t) g * kStageBytes;
        const uint32_t lane_base = tmem_base + ((quarter * 32u) << 16);
        float sc[PTM_NCOEF];
#pragma unroll
        for (int k = 0; k < PTM_NCOEF; ++k)
            sc[k] = m.scale[k];
        uint64_t st_policy = 0;
        if (m.store_hint == 1)
            st_policy = l2_policy_evict_first();
        else if (m.store_hint == 2)
            asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(st_policy));

        for (unsigned i = g;; i += G) {
            const unsigned b = i % kAccBufs, use = i / kAccBufs;
            mbar_wait(&tfull[b], use & 1);
            tc_fence_after();
            const unsigned u = *reinterpret_cast<volatile unsigned *>(&unit_pub[i & 31]);
            if (u == kNoUnit)
                break;
            const unsigned valid = (m.tail_px && u + 1 == m.n_units) ? m.tail_px : kUnitPx;
            const size_t px0 = (size_t) u * kUnitPx;

            if (kDebug != 0) {
                int v[32];
                tmem_ld32(lane_base + b * kAccCols, v);
                if (v[0] == 0x7FFFFFFF)
                    ext.mn[0] = 0.f;
                tc_fence_before();
                __syncwarp();
                if (lane == 0)
                    mbar_arrive(&tempty[b]);
            } else if (kMode == 0) {
                // every byte row is a sample row of its own fit
                float f[kUnitBlocks][PTM_NCOEF];
#pragma unroll
                for (int c = 0; c < kUnitBlocks; ++c) {
                    int v[32];
                    tmem_ld32(lane_base + b * kAccCols + c * kNB, v);
#pragma unroll
                    for (int k = 0; k < PTM_NCOEF; ++k)
                        f[c][k] = __fmul_rn(__ll2float_rn(digits_to_int(&v[4 * k])), sc[k]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0)
                    mbar_arrive(&tempty[b]);
                if (issuer)
                    bulk_wait_read();
                group_sync(1 + g);
#pragma unroll
                for (int c = 0; c < kUnitBlocks; ++c) {
                    const unsigned gb = c * kBlk + r, px = gb / 3, ch = gb - 3 * px;
                    if (ch == 0 && px < valid)
                        ext.add(f[c]);
                    float2 *row = reinterpret_cast<float2 *>(stg + ch * kStagePlane + px * 24);
                    row[0] = make_float2(f[c][0], f[c][1]);
                    row[1] = make_float2(f[c][2], f[c][3]);
                    row[2] = make_float2(f[c][4], f[c][5]);
                }
                fence_async_smem();
                group_sync(1 + g);
                if (issuer) {
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch)
                        store_out(m.store_hint, st_policy, a.coeffs + (size_t) ch * a.plane_stride + px0 * PTM_NCOEF, stg + ch * kStagePlane,
                                 valid * 24);
                    bulk_commit();
                }
            } else if (kMode == 1) {
                // of this lane's three byte rows exactly one (block r % 3) is a luma row
                const unsigned mine = r % 3;
                int sums[kUnitBlocks];
                float cf[PTM_NCOEF];
#pragma unroll
                for (int c = 0; c < kUnitBlocks; ++c) {
                    int v[32];
                    tmem_ld32(lane_base + b * kAccCols + c * kNB, v);
                    sums[c] = v[kSumCol];
                    if (mine == (unsigned) c) {
#pragma unroll
                        for (int k = 0; k < PTM_NCOEF; ++k)
                            cf[k] = __fmul_rn(__ll2float_rn(digits_to_int(&v[4 * k])), sc[k]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0)
                    mbar_arrive(&tempty[b]);
                // truncated integer means of all 384 bytes (rti-builder/ptmlib.c:794-796)
                uint8_t *means = stg + kStagePlane + 384;
                unsigned avg[kUnitBlocks];
#pragma unroll
                for (int c = 0; c < kUnitBlocks; ++c) {
                    avg[c] = div_magic((unsigned) sums[c], a.magic);
                    means[c * kBlk + r] = (uint8_t) avg[c];
                }
                const unsigned ya = mine == 0 ? avg[0] : mine == 1 ? avg[1] : avg[2];
                if (issuer)
                    bulk_wait_read();
                group_sync(1 + g);
                const unsigned q = (mine * kBlk + r) / 3;     // this lane's pixel inside the unit
                const unsigned cba = means[3 * q + 1], cra = means[3 * q + 2];
                unsigned R, Gc, B;
                lrgb_finish(ya, cba, cra, cf, R, Gc, B);
                if (q < valid)
                    ext.add(cf);
                if (a.coeffs) {
                    float2 *row = reinterpret_cast<float2 *>(stg + q * 24);
                    row[0] = make_float2(cf[0], cf[1]);
                    row[1] = make_float2(cf[2], cf[3]);
                    row[2] = make_float2(cf[4], cf[5]);
                }
                uint8_t *rgb = stg + kStagePlane + 3 * q;
                rgb[0] = (uint8_t) R;
                rgb[1] = (uint8_t) Gc;
                rgb[2] = (uint8_t) B;
                fence_async_smem();
                group_sync(1 + g);
                if (issuer) {
                    if (a.coeffs)
                        store_out(m.store_hint, st_policy, a.coeffs + px0 * PTM_NCOEF, stg, valid * 24);
                    store_out(m.store_hint, st_policy, a.rgb + px0 * 3, stg + kStagePlane, valid * 3);
                    bulk_commit();
                }
            } else if (kMode == 2) {
                // 16-bit samples (ptm_fit_poly_uint semantics, rti-builder/ptmlib.c:727-760): byte rows come
                // in (low, high) pairs on neighbouring lanes, sample = lo + 256 * hi, so the exact integer
                // dot products combine the same way.  The even lane finishes coefficients 0..2 of the
                // pair's sample, the odd lane 3..5.
                const bool odd = (r & 1) != 0;
                float f[kUnitBlocks][3];
#pragma unroll
                for (int c = 0; c < kUnitBlocks; ++c) {
                    int v[32];
                    tmem_ld32(lane_base + b * kAccCols + c * kNB, v);
                    long long t[PTM_NCOEF];
#pragma unroll
                    for (int k = 0; k < PTM_NCOEF; ++k)
                        t[k] = digits_to_int(&v[4 * k]);
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const long long send = odd ? t[j] : t[3 + j];
                        const long long recv = __shfl_xor_sync(0xFFFFFFFFu, send, 1);
                        const long long whole = odd ? recv + 256 * t[3 + j] : t[j] + 256 * recv;
                        f[c][j] = __fmul_rn(__ll2float_rn(whole), odd ? sc[3 + j] : sc[j]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0)
                    mbar_arrive(&tempty[b]);
                if (issuer)
                    bulk_wait_read();
                group_sync(1 + g);
#pragma unroll
                for (int c = 0; c < kUnitBlocks; ++c) {
                    const unsigned sg = (c * kBlk + r) >> 1, px = sg / 3, ch = sg - 3 * px;
                    if (ch == 0 && px < valid) {
#pragma unroll
                        for (int k = 0; k < PTM_NCOEF; ++k) {
                            if ((k >= 3) == odd) {
                                ext.mn[k] = fminf(ext.mn[k], f[c][k % 3]);
                                ext.mx[k] = fmaxf(ext.mx[k], f[c][k % 3]);
                            }
                        }
                    }
                    float *row = reinterpret_cast<float *>(stg + ch * kStagePlane + px * 24 + (odd ? 12 : 0));
                    row[0] = f[c][0];
                    row[1] = f[c][1];
                    row[2] = f[c][2];
                }
                fence_async_smem();
                group_sync(1 + g);
                if (issuer) {
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch)
                        store_out(m.store_hint, st_policy, a.coeffs + (size_t) ch * a.plane_stride + px0 * PTM_NCOEF, stg + ch * kStagePlane,
                                 valid * 24);
                    bulk_commit();
                }
            } else if (kMode == 3) {
                // LRGB on 16-bit samples (no counterpart in the reference; semantics of oracle/ptm_oracle.c:
                // orc_lrgb16_finish and of fit_u16_tma_kernel).  Lanes come in (low byte, high byte) pairs; pair
                // j = r / 2 holds sample 64 c + j of column block c, whose channel is (c + j) % 3: each pair
                // owns exactly one luma sample, i.e. finishes one of the unit's 64 pixels.
                const bool odd = (r & 1) != 0;
                const unsigned j = r >> 1;
                const unsigned cy = (3u - j % 3u) % 3u;       // the block in which this pair's sample is luma
                int sums[kUnitBlocks];
                long long t[PTM_NCOEF] = {0, 0, 0, 0, 0, 0};
#pragma unroll
                for (int c = 0; c < kUnitBlocks; ++c) {
                    int v[32];
                    tmem_ld32(lane_base + b * kAccCols + c * kNB, v);
                    sums[c] = v[kSumCol];
                    if (cy == (unsigned) c) {
#pragma unroll
                        for (int k = 0; k < PTM_NCOEF; ++k)
                            t[k] = digits_to_int(&v[4 * k]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0)
                    mbar_arrive(&tempty[b]);
                // 16-bit sums and truncated means of the pair's three samples
                uint16_t *means = reinterpret_cast<uint16_t *>(stg + kStagePlane + 384);
                unsigned mean[kUnitBlocks];
#pragma unroll
                for (int c = 0; c < kUnitBlocks; ++c) {
                    const int other = __shfl_xor_sync(0xFFFFFFFFu, sums[c], 1);
                    const unsigned s16 = odd ? (unsigned) other + 256u * (unsigned) sums[c]
                                             : (unsigned) sums[c] + 256u * (unsigned) other;
                    mean[c] = s16 / a.n_lights;
                    if (!odd)
                        means[c * 64 + j] = (uint16_t) mean[c];
                }
                const unsigned ya = cy == 0 ? mean[0] : cy == 1 ? mean[1] : mean[2];
                // the even lane finishes coefficients 0..2 of the luma fit, the odd lane 3..5
                float f[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const long long send = odd ? t[k] : t[3 + k];
                    const long long recv = __shfl_xor_sync(0xFFFFFFFFu, send, 1);
                    const long long whole = odd ? recv + 256 * t[3 + k] : t[k] + 256 * recv;
                    f[k] = __fmul_rn(__ll2float_rn(whole), odd ? sc[3 + k] : sc[k]);
                }
                if (issuer)
                    bulk_wait_read();
                group_sync(1 + g);
                const unsigned q = (cy * 64 + j) / 3;         // this pair's pixel inside the unit
                const unsigned cba = means[3 * q + 1], cra = means[3 * q + 2];
                float dummy[PTM_NCOEF] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                unsigned R, Gc, B;
                lrgb_finish(ya >> 8, cba >> 8, cra >> 8, dummy, R, Gc, B);
                const float norm = __fdiv_rn(256.0f, (float) ya);
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    f[k] = __fmul_rn(f[k], norm);
                if (q < valid) {
#pragma unroll
                    for (int k = 0; k < PTM_NCOEF; ++k) {
                        if ((k >= 3) == odd) {
                            ext.mn[k] = fminf(ext.mn[k], f[k % 3]);
                            ext.mx[k] = fmaxf(ext.mx[k], f[k % 3]);
                        }
                    }
                }
                if (a.coeffs) {
                    float *row = reinterpret_cast<float *>(stg + q * 24 + (odd ? 12 : 0));
                    row[0] = f[0];
                    row[1] = f[1];
                    row[2] = f[2];
                }
                if (!odd) {
                    uint8_t *rgb = stg + kStagePlane + 3 * q;
                    rgb[0] = (uint8_t) R;
                    rgb[1] = (uint8_t) Gc;
                    rgb[2] = (uint8_t) B;
                }
                fence_async_smem();
                group_sync(1 + g);
                if (issuer) {
                    if (a.coeffs)
                        store_out(m.store_hint, st_policy, a.coeffs + px0 * PTM_NCOEF, stg, valid * 24);
                    store_out(m.store_hint, st_policy, a.rgb + px0 * 3, stg + kStagePlane, valid * 3);
                    bulk_commit();
                }
            }
        }
        if (issuer) {
            bulk_wait_all();
            asm volatile("fence.proxy.async.global;" ::: "memory");
        }
    }
    ext.flush(a.keys, keys_s);
    if (a.handoff.ticket)
        keys_handoff<KeyHandoff, XchgLayout>(a.handoff, a.keys, kMaxRanks, kSlotInts);
    tc_fence_before();
    __syncthreads();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t) kTmemCols)
                     : "memory");
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------

// Fixed-point digits of the pseudo-inverse, laid out as the shared-memory image of the B operand
// (K-major, no swizzle: [16-light chunk][8-column group][column][light]).  Returns false when
// the matrix cannot be represented (non-finite or all-zero rows are left to the FFMA kernels).
bool pack_mma_digits(const float *M, unsigned n_lights, std::vector<int8_t> &image, float *scale, unsigned *kpad_out) {
    if (n_lights == 0 || n_lights > 224)
        return false;
    const unsigned kpad = (n_lights + 31) / 32 * 32;
    image.assign((size_t) kpad * 32, 0);
    for (int k = 0; k < PTM_NCOEF; ++k) {
        double mx = 0.0;
        for (unsigned n = 0; n < n_lights; ++n) {
            const double v = M[(size_t) k * n_lights + n];
            if (!std::isfinite(v))
                return false;
            mx = std::fmax(mx, std::fabs(v));
        }
        int F = 0;
        if (mx > 0.0) {
            int e;
            std::frexp(mx, &e);      // mx = f * 2^e, f in [0.5, 1)  =>  mx * 2^(30 - e) < 2^30
            F = 30 - e;
        }
        if (F > 120 || F < -90)
            return false;
        scale[k] = (float) std::ldexp(1.0, -F);
        for (unsigned n = 0; n < n_lights; ++n) {
            long long q = std::llrint(std::ldexp((double) M[(size_t) k * n_lights + n], F));
            // balanced base-256 digits, least significant first
            int d[4];
            for (int j = 3; j >= 1; --j) {
                int low = (int) (((q % 256) + 256) % 256);
                if (low >= 128)
                    low -= 256;
                d[j] = low;
                q = (q - low) / 256;
            }
            d[0] = (int) q;          // |q| <= 2^30 => |d0| <= 65
            for (int j = 0; j < 4; ++j) {
                const unsigned col = 4 * k + j;
                image[(size_t) (n / 16) * 512 + (col / 8) * 128 + (col % 8) * 16 + (n % 16)] = (int8_t) d[j];
            }
        }
    }
    for (unsigned n = 0; n < n_lights; ++n)
        image[(size_t) (n / 16) * 512 + (kSumCol / 8) * 128 + (kSumCol % 8) * 16 + (n % 16)] = 1;
    *kpad_out = kpad;
    return true;
}

static unsigned long long g_mma_launches = 0;   // launches of the tensor-core kernel by this process

static PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

template <int kMode, int G, int kDebug = 0>
static cudaError_t launch_mode(const CUtensorMap *maps, const FitArgs &a, const MmaArgs &m, size_t smem, unsigned grid,
                               cudaStream_t st) {
    static bool configured[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess)
        return e;
    if (dev < 0 || dev >= 64)
        return cudaErrorInvalidDevice;
    auto kern = fit_mma_kernel<kMode, G, kDebug>;
    if (!configured[dev]) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
        if (e != cudaSuccess)
            return e;
        configured[dev] = true;
    }
    kern<<<grid, 64 + 128 * G, smem, st>>>(maps[0], maps[1], maps[2], a, m);
    return cudaGetLastError();
}

// mode: 0 = RGB u8, 1 = LRGB u8, 2 = RGB u16, 3 = LRGB u16.  Covers whole 384-byte units (+ a 16-pixel-granular partial
// last unit) and reports the pixels it covered through *done; 0 = layout not eligible.
cudaError_t launch_fit_mma(const FitArgs &a, int mode, int sm_count, cudaStream_t st, size_t *done) {
    *done = 0;
    if (mode < 0 || mode > 3)
        return cudaSuccess;
    if (!a.digits || a.kpad == 0 || a.kpad > 224 || a.n_lights > a.kpad)
        return cudaSuccess;
    const unsigned unit_px = mode >= 2 ? 64 : 128, px_bytes = mode >= 2 ? 6 : 3;
    const size_t row_bytes_all = a.pixels * px_bytes;
    bool ok = ((uintptr_t) a.stack % 16 == 0) && (a.image_stride % 16 == 0) && row_bytes_all < 0x7FFFFFFFull &&
              a.image_stride >= row_bytes_all;
    if (mode == 0 || mode == 2)
        ok = ok && a.coeffs && ((uintptr_t) a.coeffs % 16 == 0) && (a.plane_stride % 4 == 0);
    else
        ok = ok && ((uintptr_t) a.rgb % 16 == 0) && (a.coeffs == nullptr || (uintptr_t) a.coeffs % 16 == 0);
    const size_t full_units = a.pixels / unit_px;
    const unsigned rem = (unsigned) (a.pixels % unit_px);
    const unsigned tail_px = (rem % 16 == 0) ? rem : 0;
    const size_t n_units = full_units + (tail_px ? 1 : 0);
    if (!ok || n_units == 0)
        return cudaSuccess;
    PFN_cuTensorMapEncodeTiled_v12000 encode = tensor_map_encoder();
    if (!encode)
        return cudaSuccess;

    static int variant = -1;
    if (variant < 0) {
        const char *v = getenv("PTM_MMA_VARIANT");   // tuning aid; the default is what ships
        variant = v ? atoi(v) : 0;
    }
#ifdef PTM_TUNING
    const int G = variant == 1 ? 4 : 2;
#else
    const int G = 2;
#endif
    MmaArgs m;
    m.digits = a.digits;
    memcpy(m.scale, a.digit_scale, sizeof(m.scale));
    m.kpad = a.kpad;
    m.n_units = (unsigned) n_units;
    m.tail_px = tail_px;
    m.row_bytes = (unsigned) ((full_units * unit_px + tail_px) * px_bytes);
    const size_t stage_bytes = (size_t) kUnitBytes * a.kpad;
    const size_t fixed = 1024 + (size_t) a.kpad * 32 + (size_t) G * kStageBytes + (2 * kMaxStages + 2 * kAccBufs) * 8 + 16;
    size_t stages = (kMaxDynSmem - fixed) / stage_bytes;
    if (stages > (size_t) kMaxStages)
        stages = kMaxStages;
    if (stages < 2)
        return cudaSuccess;
    m.stages = (unsigned) stages;
    const size_t smem = fixed - (2 * kMaxStages) * 8 + 2 * stages * 8 + stages * stage_bytes;

    CUtensorMap maps[3];
    CUtensorMap &map = maps[0];
    memset(maps, 0, sizeof(maps));
    const cuuint64_t gdim[2] = {(cuuint64_t) m.row_bytes, (cuuint64_t) a.n_lights};
    const cuuint64_t gstride[1] = {(cuuint64_t) a.image_stride};
    const cuuint32_t box[2] = {(cuuint32_t) kBlk, (cuuint32_t) a.n_lights};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t *>(a.stack), gdim, gstride, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return cudaSuccess;   // not representable: the FFMA kernels take the slab

    // light-group-major maps for whole units (see the producer): the 128-byte window and the column block
    // index are separate dimensions over the same bytes; only whole column blocks are addressable, the
    // partial last unit goes through the plain 2-D map above
    static int load_mode = -1, l2_promo = -1;
    if (load_mode < 0) {
        const char *v = getenv("PTM_MMA_LOAD");
        load_mode = v ? atoi(v) : 1;
        const char *w = getenv("PTM_MMA_L2");
        l2_promo = w ? atoi(w) : 128;
    }
    const CUtensorMapL2promotion promo = l2_promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                         : l2_promo == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                                         : l2_promo == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                                                          : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    m.grouped = 0;
    const unsigned groups = a.n_lights / 8, rem_l = a.n_lights % 8;
    const cuuint64_t whole_blocks = m.row_bytes / kBlk;
    if (load_mode == 1 && whole_blocks >= (cuuint64_t) kUnitBlocks) {
        bool ok2 = true;
        if (groups) {
            const cuuint64_t d4[4] = {(cuuint64_t) kBlk, 8, whole_blocks, groups};
            const cuuint64_t s4[3] = {(cuuint64_t) a.image_stride, (cuuint64_t) kBlk, (cuuint64_t) a.image_stride * 8};
            const cuuint32_t b4[4] = {(cuuint32_t) kBlk, 8, (cuuint32_t) kUnitBlocks, groups};
            const cuuint32_t e4[4] = {1, 1, 1, 1};
            ok2 = encode(&maps[1], CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<uint8_t *>(a.stack), d4, s4, b4, e4,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        }
        if (ok2 && rem_l) {
            const cuuint64_t d3[3] = {(cuuint64_t) kBlk, rem_l, whole_blocks};
            const cuuint64_t s3[2] = {(cuuint64_t) a.image_stride, (cuuint64_t) kBlk};
            const cuuint32_t b3[3] = {(cuuint32_t) kBlk, 8, (cuuint32_t) kUnitBlocks};
            const cuuint32_t e3[3] = {1, 1, 1};
            ok2 = encode(&maps[2], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3,
                         const_cast<uint8_t *>(a.stack) + (size_t) groups * 8 * a.image_stride, d3, s3, b3, e3,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        }
        m.grouped = ok2 ? 1 : 0;
    }
    if (!m.grouped) {
        maps[1] = maps[0];
        maps[2] = maps[0];
    }

    unsigned grid = (unsigned) sm_count;
    if ((size_t) grid > n_units)
        grid = (unsigned) n_units;
    static int run_units = -1;
    if (run_units < 0) {
        const char *v = getenv("PTM_MMA_RUN");   // tuning aid: units per run
        run_units = v ? atoi(v) : 8;
        if (run_units < 1)
            run_units = 1;
    }
    static int store_hint = -1;
    if (store_hint < 0) {
        const char *v = getenv("PTM_MMA_STORE_HINT");
        store_hint = v ? atoi(v) : 2;
    }
    m.store_hint = (unsigned) store_hint;
    m.run_units = (unsigned) run_units;
    m.counter = a.tile_counter;
    const size_t n_runs = (n_units + m.run_units - 1) / m.run_units;
    if ((size_t) grid > n_runs)
        grid = (unsigned) n_runs;
    // Invariant: the run counter is 0 between launches.  The hand-off block of a fused launch resets it on the
    // device; any other launch is followed by a stream-ordered clear (below).
    cudaError_t e;
    switch (mode == 3 ? 0 : variant) {   // the tuning variants exist for modes 0-2

#ifdef PTM_TUNING   // make EXTRA_NVFLAGS=-DPTM_TUNING: 1 = four epilogue groups; 2-4 = ceilings only, they skip work and do NOT produce results
    case 1:
        e = mode == 0 ? launch_mode<0, 4>(maps, a, m, smem, grid, st)
            : mode == 1 ? launch_mode<1, 4>(maps, a, m, smem, grid, st) : launch_mode<2, 4>(maps, a, m, smem, grid, st);
        break;
    case 2: e = mode == 2 ? launch_mode<2, 2, 1>(maps, a, m, smem, grid, st) : launch_mode<1, 2, 1>(maps, a, m, smem, grid, st); break;
    case 3: e = mode == 2 ? launch_mode<2, 2, 2>(maps, a, m, smem, grid, st) : launch_mode<1, 2, 2>(maps, a, m, smem, grid, st); break;
    case 4: e = mode == 2 ? launch_mode<2, 2, 3>(maps, a, m, smem, grid, st) : launch_mode<1, 2, 3>(maps, a, m, smem, grid, st); break;
#endif
    default:
        e = mode == 0   ? launch_mode<0, 2>(maps, a, m, smem, grid, st)
            : mode == 1 ? launch_mode<1, 2>(maps, a, m, smem, grid, st)
            : mode == 2 ? launch_mode<2, 2>(maps, a, m, smem, grid, st)
                        : launch_mode<3, 2>(maps, a, m, smem, grid, st);
        break;
    }
    if (e == cudaSuccess) {
        *done = full_units * unit_px + tail_px;
        ++g_mma_launches;
        if (m.counter && !a.handoff.ticket)
            e = cudaMemsetAsync(m.counter, 0, sizeof(unsigned), st);
    }
    return e;
}

unsigned long long mma_launch_count() { return g_mma_launches; }

}  // namespace ptm
