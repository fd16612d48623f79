// reverb.cu — FIR convolutional reverb as a bf16 tcgen05 GEMM with TMEM accumulators.
//
//   y[v][n] = sum_{k<L} bf16(h_c[k]) * bf16(x[v][n-k])          (per IR channel c; products exact, fp32 accumulate)
//
// For one IR channel all voices share h, so a tile of outputs is a plain GEMM with a long reduction:
//   D[128 voices][BN frames] = A[128][K] * Bt[BN][K]^T,   K = L + BN - 1 (padded to 64), BN in {256, 224, 192, 128}
//   A[v][j]  = xh[row v][H + n0 - Lr + j]         a TMA window of the bf16 sample history, K-major as stored;
//                                                 Lr = roundup(L-1, 8) keeps every box start 16-byte aligned (TMA rule)
//   Bt[i][j] = h[Lr - (j - i)]  (0 <= Lr-(j-i) < L)  Toeplitz expansion of the reversed IR, built ONCE per IR (12 MB per
//                                                 channel for L = 48000) and then L2-resident: it is the same for
//                                                 every time tile and every voice tile.
// The band is (L / K) = 99.5 % dense, so the GEMM does no meaningful wasted work.
//
// Kernel anatomy (persistent grid, one CTA per SM, tiles strided by the grid; 256 threads):
//   warp 0   TMA producer   cp.async.bulk.tensor.2d -> 128B-swizzled smem stages, mbarrier expect_tx
//   warp 1   MMA issuer     one elected thread: 4 x tcgen05.mma.kind::f16 (M128 N256 K16) per 64-wide k-block,
//                           tcgen05.commit frees the stage / signals the epilogue
//   warp 2   TMEM allocator (512 columns = two accumulators: a segment's epilogue overlaps the next segment's MMAs)
//   warps 4-7 epilogue      tcgen05.ld 32x32b.x32 -> registers -> 16-byte stores of y
// 4 stages x (16 KB A + 32 KB B) = 192 KB of shared memory.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <utility>

#include "kernels.cuh"
#include "plan.hpp"

namespace fw {

constexpr uint32_t RV_BM = 128, RV_BN = 256, RV_BK = 64, RV_STAGES = 4;
constexpr uint32_t RV_A_BYTES = RV_BM * RV_BK * 2, RV_B_BYTES = RV_BN * RV_BK * 2;
constexpr uint32_t RV_SMEM_BYTES = RV_STAGES * (RV_A_BYTES + RV_B_BYTES) + 1024 /*align*/ + 256 /*barriers*/;

// ---------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int32_t x, int32_t y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tcgen05_alloc(uint32_t* smem_result, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_dealloc(uint32_t taddr, uint32_t cols) { asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tcgen05_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// UMMA shared-memory descriptor, K-major operand, 128-byte swizzle: rows at 128 B pitch, 8-row groups at SBO = 1024 B,
// LBO = 1 (unused for swizzled K-major), descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024u >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// Instruction descriptor, kind::f16: D = F32, A = B = BF16, both K-major, N >> 3 at bit 17, M >> 4 at bit 24.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(uint32_t M, uint32_t N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24); }

// ---------------------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------------------
// Bt[c][i][j] = bf16(h_c[Lr-(j-i)]) where 0 <= Lr-(j-i) < L, else 0.    [ir_ch][256][Kpad]
__global__ void reverb_build_toeplitz(const float* __restrict__ ir, __nv_bfloat16* __restrict__ bt, uint32_t L, uint32_t Lr, uint32_t ir_ch, uint32_t kpad) {
    const size_t n = (size_t)ir_ch * RV_BN * kpad;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        const uint32_t j = idx % kpad; const size_t ci = idx / kpad; const uint32_t i = ci % RV_BN, c = (uint32_t)(ci / RV_BN);
        float v = 0.0f;
        const int64_t k = (int64_t)Lr - ((int64_t)j - (int64_t)i);
        if (k >= 0 && k < (int64_t)L) v = ir[(size_t)c * L + (size_t)k];
        bt[idx] = __float2bfloat16_rn(v);
    }
}

// xh[c*V + v][cursor + t] = bf16(in[v][c][t]) for t < T: appends the call's block behind the history (8 samples per thread)
__global__ void reverb_prepare(const float* __restrict__ in, __nv_bfloat16* __restrict__ xh, uint32_t V, uint32_t C, uint32_t T, uint32_t in_pitch, uint32_t cursor,
                               uint32_t pitch, uint32_t zero_first, uint32_t chan_base) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // the GEMM's set-up (barriers, TMEM, tensor maps) overlaps this kernel
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t row = blockIdx.x;  // c * V + v (rows on grid.x: no 65535 cap)
    const uint32_t c = row / V, v = row % V;
    __nv_bfloat16* dst = xh + ((size_t)chan_base * V + row) * pitch + cursor;
    const float* x = in + ((size_t)v * C + c) * in_pitch;
    const bool vec = (T % 8u) == 0 && (in_pitch % 4u) == 0 && (cursor % 8u) == 0 && (pitch % 8u) == 0 && (reinterpret_cast<uintptr_t>(in) % 16u) == 0;
    if (vec) {
        for (uint32_t i = (blockIdx.y * blockDim.x + threadIdx.x) * 8u; i < T; i += gridDim.y * blockDim.x * 8u) {
            float4 a = __ldcs(reinterpret_cast<const float4*>(x + i)), b = __ldcs(reinterpret_cast<const float4*>(x + i + 4));
            if (i < zero_first) { a = make_float4(0.f, 0.f, 0.f, 0.f); b = a; }  // zero_first is a multiple of the block size here
            __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w), p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
            uint4 o;
            o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1); o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
            *reinterpret_cast<uint4*>(dst + i) = o;
        }
    } else {
        for (uint32_t i = blockIdx.y * blockDim.x + threadIdx.x; i < T; i += gridDim.y * blockDim.x) dst[i] = __float2bfloat16_rn(i < zero_first ? 0.0f : x[i]);
    }
}

struct ReverbGemmArgs {
    float* out;                   // row (v * C + c) at out + row * out_pitch
    uint32_t out_pitch, V, C, T, Lr, cursor, ir_ch, num_kb, chan_base;
    uint32_t tiles_n, tiles_m, total_tiles;  // output tiles along frames / voices (per channel); tiles_n * tiles_m * C
    // CTA-pair kernel, tail wave: the last `tail_tiles` tiles (fewer than half a wave of pairs) are each split along K between
    // `tail_split` pairs; the pair with the first slice adds the others' partial sums (fix-up workspace, one flag per CTA).
    uint32_t full_tiles, tail_tiles, tail_split;
    float* ws; uint32_t* flags; uint32_t epoch;
};
struct RvSeg { uint32_t tile, k0, k1; };
// segment i of pair P: whole tiles P, P + NP, ... of the full waves, then (at most) one slice of a tail tile
__device__ __forceinline__ uint32_t rv_seg_count(const ReverbGemmArgs& a, uint32_t P, uint32_t NP) {
    const uint32_t full = a.full_tiles > P ? (a.full_tiles - P + NP - 1) / NP : 0u;
    return full + (P < a.tail_tiles * a.tail_split ? 1u : 0u);
}
__device__ __forceinline__ RvSeg rv_seg(const ReverbGemmArgs& a, uint32_t P, uint32_t NP, uint32_t i) {
    const uint32_t t = P + i * NP;
    if (t < a.full_tiles) return RvSeg{t, 0u, a.num_kb};
    const uint32_t j = P / a.tail_split, sl = P % a.tail_split;
    return RvSeg{a.full_tiles + j, (uint32_t)((uint64_t)sl * a.num_kb / a.tail_split), (uint32_t)((uint64_t)(sl + 1) * a.num_kb / a.tail_split)};
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) { uint32_t v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }

// Persistent grid, one CTA per SM: CTA g computes tiles g, g + G, g + 2G, ... each over the whole reduction. All CTAs walk
// the k-blocks of their tiles IN STEP, and the Toeplitz operand B depends only on (channel, k-block): at any moment the whole
// chip reads the same few B blocks, so B streams through L2 once instead of living there. (A stream-K split that gives every
// SM an equal k-range was measured on config 4 and lost badly for exactly that reason — 0.39 ms against 0.28 ms — the CTAs
// then sit at 148 different k-offsets and B + the sample history no longer fit L2.) SM coverage comes from the tile width
// instead: BN is chosen per call among 256 / 224 / 192 / 128 frames so that the tile count fills whole waves of SMs (config 4:
// 37 x 2 x 2 = 148 tiles of 224 frames on 148 SMs, instead of 128 tiles of 256). Two TMEM accumulators: the epilogue of a tile
// overlaps the next tile's MMAs.
template <uint32_t BN>
__global__ void __launch_bounds__(256, 1) reverb_gemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const ReverbGemmArgs a) {
    constexpr uint32_t B_BYTES = BN * RV_BK * 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));  // SWIZZLE_128B wants 1024-byte tiles
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + RV_STAGES * RV_A_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + RV_STAGES * (RV_A_BYTES + RV_B_BYTES));
    uint64_t* empty_bar = full_bar + RV_STAGES;
    uint64_t* tmem_full_bar = empty_bar + RV_STAGES;   // [2]: accumulator complete (MMA -> epilogue)
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2]: accumulator drained  (epilogue -> MMA)
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // the next kernel's launch latency hides behind this one; it waits for our results itself
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t g = blockIdx.x, G = gridDim.x, num_kb = a.num_kb;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (uint32_t s = 0; s < RV_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (uint32_t b = 0; b < 2; ++b) { mbar_init(&tmem_full_bar[b], 1); mbar_init(&tmem_empty_bar[b], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tcgen05_alloc(tmem_base_slot, 512);  // two accumulators of up to 256 columns
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;
    const uint32_t tiles_per_ch = a.tiles_n * a.tiles_m;

    if (warp == 0) {
        // ===== TMA producer =====
        asm volatile("griddepcontrol.wait;" ::: "memory");  // the history buffer is written by reverb_prepare just before us
        if (lane == 0) {
            uint32_t it = 0;
            for (uint32_t t = g; t < a.total_tiles; t += G) {
                const uint32_t c = t / tiles_per_ch, rem = t % tiles_per_ch, mt = rem / a.tiles_n, nt = rem % a.tiles_n;
                const int32_t col_a0 = (int32_t)(a.cursor + nt * BN) - (int32_t)a.Lr;  // multiple of 8 elements = 16 bytes
                const int32_t row_a = (int32_t)((a.chan_base + c) * a.V + mt * RV_BM), row_b = (int32_t)(((a.chan_base + c) % a.ir_ch) * RV_BN);
                for (uint32_t kb = 0; kb < num_kb; ++kb, ++it) {
                    const uint32_t s = it % RV_STAGES, ph = (it / RV_STAGES) & 1u;
                    mbar_wait(&empty_bar[s], ph ^ 1u);
                    mbar_expect_tx(&full_bar[s], RV_A_BYTES + B_BYTES);
                    tma_load_2d(smem_a + s * RV_A_BYTES, &tm_a, &full_bar[s], col_a0 + (int32_t)(kb * RV_BK), row_a);
                    tma_load_2d(smem_b + s * RV_B_BYTES, &tm_b, &full_bar[s], (int32_t)(kb * RV_BK), row_b);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one elected lane) =====
        constexpr uint32_t idesc = umma_idesc_bf16(RV_BM, BN);
        uint32_t it = 0, seg = 0;
        for (uint32_t t = g; t < a.total_tiles; t += G, ++seg) {
            const uint32_t buf = seg & 1u;
            mbar_wait(&tmem_empty_bar[buf], ((seg >> 1) & 1u) ^ 1u);  // the epilogue has drained this accumulator (free on first use)
            tcgen05_fence_after();
            const uint32_t tmem_d = tmem_base + buf * RV_BN;
            for (uint32_t kb = 0; kb < num_kb; ++kb, ++it) {
                const uint32_t s = it % RV_STAGES, ph = (it / RV_STAGES) & 1u;
                mbar_wait(&full_bar[s], ph);
                tcgen05_fence_after();
                if (elect_one()) {
                    const uint64_t adesc = umma_desc_sw128(smem_u32(smem_a + s * RV_A_BYTES));
                    const uint64_t bdesc = umma_desc_sw128(smem_u32(smem_b + s * RV_B_BYTES));
#pragma unroll
                    for (uint32_t k = 0; k < RV_BK / 16; ++k)  // +32 bytes per K=16 step inside the 128-byte swizzle atom
                        tcgen05_mma_f16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
                    tcgen05_commit(&empty_bar[s]);                                // smem stage free once these MMAs retire
                    if (kb + 1 == num_kb) tcgen05_commit(&tmem_full_bar[buf]);    // accumulator complete
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> global =====
        const uint32_t q = warp & 3u;  // a warp may only touch TMEM lanes [32q, 32q+32)
        uint32_t seg = 0;
        for (uint32_t t = g; t < a.total_tiles; t += G, ++seg) {
            const uint32_t buf = seg & 1u;
            const uint32_t c = t / tiles_per_ch, rem = t % tiles_per_ch, mt = rem / a.tiles_n, nt = rem % a.tiles_n;
            const uint32_t n0 = nt * BN, v = mt * RV_BM + q * 32u + lane;
            mbar_wait(&tmem_full_bar[buf], (seg >> 1) & 1u);
            tcgen05_fence_after();
            float* dst_row = a.out + ((size_t)v * a.C + c) * a.out_pitch + n0;
#pragma unroll 1
            for (uint32_t col = 0; col < BN; col += 32) {
                uint32_t r[32];
                tcgen05_ld_32x32b_x32(tmem_base + ((q * 32u) << 16) + buf * RV_BN + col, r);
                if (v < a.V) {
                    if (n0 + col + 32 <= a.T && ((a.T | a.out_pitch) & 3u) == 0) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4)
                            __stcs(reinterpret_cast<float4*>(dst_row + col + i), make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]), __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])));
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) if (n0 + col + i < a.T) dst_row[col + i] = __uint_as_float(r[i]);
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) tcgen05_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): two SMs of a cluster share one 256-voice x BN-frame tile.
// Why: the single-CTA kernel is bound by SHARED-MEMORY bandwidth, not by the tensor pipe. Per 64-wide k-block an SM writes
// (TMA) and reads (UMMA operand fetch) A 16 KB + B BN*128 B each: 96 KB at BN = 256, i.e. 768 cycles at 128 B/clk against
// 512 cycles of MMA (measured on config 5: 730 cycles per k-block at BN = 256, 555 at BN = 128 — the fixed A term does not
// shrink with the tile). In a pair each SM stages its own 128 voices of A and only HALF of B (the tensor cores of both SMs
// consume both halves), so the traffic is 64 KB per k-block = 512 cycles: level with the MMA.
//   * cluster (2,1,1); rank 0 (leader) issues every tcgen05.mma.cta_group::2 (M = 256), both CTAs run TMA + epilogue;
//   * full barriers live in the leader: both CTAs' TMA loads complete_tx on them (cp.async.bulk.tensor .cta_group::2);
//   * tcgen05.commit ... multicast::cluster releases the smem stage / publishes the accumulator in BOTH CTAs;
//   * the leader's MMA warp waits for the epilogues of both CTAs (8 warps arrive on its tmem_empty barrier, the peer's
//     through mapa + mbarrier.arrive.shared::cluster);
//   * cluster barrier after set-up and before tear-down (the leader's MMAs read the peer's smem and TMEM).
// ---------------------------------------------------------------------------------------------------------------
constexpr uint32_t RV2_STAGES = 6;
constexpr uint32_t RV2_B_BYTES_MAX = (RV_BN / 2) * RV_BK * 2;  // 16 KB: half of B
constexpr uint32_t RV2_SMEM_BYTES = RV2_STAGES * (RV_A_BYTES + RV2_B_BYTES_MAX) + 1024 /*align*/ + 256 /*barriers*/;

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t smem_addr, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank)); return r; }
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) { asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory"); }
// TMA load into THIS CTA's smem whose completion bytes are counted on a barrier given by its shared::cluster address (the leader's)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* tm, uint32_t bar_cluster_addr, int32_t x, int32_t y) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(bar_cluster_addr), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tcgen05_alloc2(uint32_t* smem_result, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_dealloc2(uint32_t taddr, uint32_t cols) { asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory"); }
__device__ __forceinline__ void tcgen05_commit2(uint64_t* bar) {  // arrives on the barrier at this offset in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tcgen05_mma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

template <uint32_t BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
reverb_gemm2_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const ReverbGemmArgs a) {
    constexpr uint32_t B_BYTES = (BN / 2) * RV_BK * 2;  // this CTA's half of B
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + RV2_STAGES * RV_A_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + RV2_STAGES * (RV_A_BYTES + RV2_B_BYTES_MAX));  // used in the leader
    uint64_t* empty_bar = full_bar + RV2_STAGES;       // armed in both CTAs by the leader's multicast commit
    uint64_t* tmem_full_bar = empty_bar + RV2_STAGES;  // [2], both CTAs
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2], used in the leader: 8 epilogue warps of the pair
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const uint32_t P = blockIdx.x >> 1, NP = gridDim.x >> 1, num_kb = a.num_kb;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (uint32_t s = 0; s < RV2_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (uint32_t b = 0; b < 2; ++b) { mbar_init(&tmem_full_bar[b], 1); mbar_init(&tmem_empty_bar[b], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tcgen05_alloc2(tmem_base_slot, 512);
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();  // the peer's barriers are initialised and its TMEM is allocated before anything targets them
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;
    const uint32_t tiles_per_ch = a.tiles_n * a.tiles_m;

    if (warp == 0) {
        // ===== TMA producer (both CTAs): own 128 voices of A, own half of B =====
        asm volatile("griddepcontrol.wait;" ::: "memory");
        if (lane == 0) {
            uint32_t it = 0;
            const uint32_t nseg = rv_seg_count(a, P, NP);
            for (uint32_t si = 0; si < nseg; ++si) {
                const RvSeg sg = rv_seg(a, P, NP, si);
                const uint32_t t = sg.tile;
                const uint32_t c = t / tiles_per_ch, rem = t % tiles_per_ch, mt = rem / a.tiles_n, nt = rem % a.tiles_n;
                const int32_t col_a0 = (int32_t)(a.cursor + nt * BN) - (int32_t)a.Lr;
                const int32_t row_a = (int32_t)((a.chan_base + c) * a.V + mt * 2u * RV_BM + rank * RV_BM);
                const int32_t row_b = (int32_t)(((a.chan_base + c) % a.ir_ch) * RV_BN + rank * (BN / 2));
                for (uint32_t kb = sg.k0; kb < sg.k1; ++kb, ++it) {
                    const uint32_t s = it % RV2_STAGES, ph = (it / RV2_STAGES) & 1u;
                    mbar_wait(&empty_bar[s], ph ^ 1u);
                    const uint32_t full_leader = mapa_rank(smem_u32(&full_bar[s]), 0);
                    if (leader) mbar_expect_tx(&full_bar[s], 2u * (RV_A_BYTES + B_BYTES));  // the bytes of both CTAs
                    tma_load_2d_pair(smem_a + s * RV_A_BYTES, &tm_a, full_leader, col_a0 + (int32_t)(kb * RV_BK), row_a);
                    tma_load_2d_pair(smem_b + s * RV2_B_BYTES_MAX, &tm_b, full_leader, (int32_t)(kb * RV_BK), row_b);
                }
            }
        }
    } else if (warp == 1 && leader) {
        // ===== MMA issuer: the leader's elected lane drives both tensor cores =====
        constexpr uint32_t idesc = umma_idesc_bf16(2 * RV_BM, BN);
        uint32_t it = 0;
        const uint32_t nseg = rv_seg_count(a, P, NP);
        for (uint32_t seg = 0; seg < nseg; ++seg) {
            const RvSeg sg = rv_seg(a, P, NP, seg);
            const uint32_t buf = seg & 1u;
            mbar_wait(&tmem_empty_bar[buf], ((seg >> 1) & 1u) ^ 1u);
            tcgen05_fence_after();
            const uint32_t tmem_d = tmem_base + buf * RV_BN;
            for (uint32_t kb = sg.k0; kb < sg.k1; ++kb, ++it) {
                const uint32_t s = it % RV2_STAGES, ph = (it / RV2_STAGES) & 1u;
                mbar_wait(&full_bar[s], ph);
                tcgen05_fence_after();
                if (elect_one()) {
                    const uint64_t adesc = umma_desc_sw128(smem_u32(smem_a + s * RV_A_BYTES));
                    const uint64_t bdesc = umma_desc_sw128(smem_u32(smem_b + s * RV2_B_BYTES_MAX));
#pragma unroll
                    for (uint32_t k = 0; k < RV_BK / 16; ++k)
                        tcgen05_mma2_f16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb != sg.k0 || k != 0) ? 1u : 0u);
                    tcgen05_commit2(&empty_bar[s]);
                    if (kb + 1 == sg.k1) tcgen05_commit2(&tmem_full_bar[buf]);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue (both CTAs): own 128 rows of the 256-row tile =====
        const uint32_t q = warp & 3u;
        const uint32_t rrow = q * 32u + lane;
        const uint32_t nseg = rv_seg_count(a, P, NP);
        constexpr uint32_t WS_F4 = RV_BM * RV_BN / 4;  // float4 per CTA partial: [col / 4][row]
        for (uint32_t seg = 0; seg < nseg; ++seg) {
            const RvSeg sg = rv_seg(a, P, NP, seg);
            const uint32_t buf = seg & 1u, t = sg.tile;
            const uint32_t c = t / tiles_per_ch, rem = t % tiles_per_ch, mt = rem / a.tiles_n, nt = rem % a.tiles_n;
            const uint32_t n0 = nt * BN, v = mt * 2u * RV_BM + rank * RV_BM + rrow;
            const bool partial = sg.k0 != 0;                                  // a later K-slice of a tail tile: park the partial sums
            const uint32_t n_follow = (sg.k0 == 0 && sg.k1 < num_kb) ? a.tail_split - 1u : 0u;  // first slice: add the others'
            if (n_follow) {  // the other slices run in this same wave on pairs P + 1 ...: they finish when we do
                for (uint32_t f = lane; f < n_follow; f += 32u) while (ld_acquire_u32(a.flags + (P + 1u + f) * 2u + rank) != a.epoch) { }
                __syncwarp();
            }
            mbar_wait(&tmem_full_bar[buf], (seg >> 1) & 1u);
            tcgen05_fence_after();
            float* dst_row = a.out + ((size_t)v * a.C + c) * a.out_pitch + n0;
            float4* ws_me = reinterpret_cast<float4*>(a.ws) + (size_t)(P * 2u + rank) * WS_F4 + rrow;
#pragma unroll 1
            for (uint32_t col = 0; col < BN; col += 32) {
                uint32_t r[32];
                tcgen05_ld_32x32b_x32(tmem_base + ((q * 32u) << 16) + buf * RV_BN + col, r);
                if (partial) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        __stcg(ws_me + (size_t)((col + i) >> 2) * RV_BM, make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]), __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])));
                    continue;
                }
                for (uint32_t f = 0; f < n_follow; ++f) {  // ascending K: ((first + slice 1) + slice 2) ...
                    const float4* wf = reinterpret_cast<const float4*>(a.ws) + (size_t)((P + 1u + f) * 2u + rank) * WS_F4 + rrow;
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 pp = __ldcg(wf + (size_t)((col + i) >> 2) * RV_BM);
                        r[i] = __float_as_uint(__uint_as_float(r[i]) + pp.x); r[i + 1] = __float_as_uint(__uint_as_float(r[i + 1]) + pp.y);
                        r[i + 2] = __float_as_uint(__uint_as_float(r[i + 2]) + pp.z); r[i + 3] = __float_as_uint(__uint_as_float(r[i + 3]) + pp.w);
                    }
                }
                if (v < a.V) {
                    if (n0 + col + 32 <= a.T && ((a.T | a.out_pitch) & 3u) == 0) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4)
                            __stcs(reinterpret_cast<float4*>(dst_row + col + i), make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]), __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])));
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) if (n0 + col + i < a.T) dst_row[col + i] = __uint_as_float(r[i]);
                    }
                }
            }
            tcgen05_fence_before();
            if (partial) {  // publish: all 128 epilogue threads of this CTA have stored, then one release store
                __threadfence();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (threadIdx.x == 128) st_release_u32(a.flags + P * 2u + rank, a.epoch);
            } else {
                __syncwarp();
            }
            if (lane == 0) mbar_arrive_cluster(mapa_rank(smem_u32(&tmem_empty_bar[buf]), 0));
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();  // nobody leaves while the pair still reads this CTA's smem / TMEM or arrives on its barriers
    if (warp == 2) tcgen05_dealloc2(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr; cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
static bool make_map_bf16_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_elems, uint32_t box_inner, uint32_t box_outer) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn) return false;
    const cuuint64_t dims[2] = {inner, outer};
    const cuuint64_t strides[1] = {pitch_elems * 2};
    const cuuint32_t box[2] = {box_inner, box_outer};
    const cuuint32_t estr[2] = {1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static uint32_t reverb_lr(uint32_t L) { return ((L - 1 + 7) / 8) * 8; }
uint32_t reverb_kpad(uint32_t L) { return ((reverb_lr(L) + RV_BN + RV_BK - 1) / RV_BK) * RV_BK; }
uint32_t reverb_hist(uint32_t L) { return ((L - 1 + 63) / 64) * 64; }  // history samples kept in front of each call's block

cudaError_t launch_reverb_build(const float* d_ir, void* d_bt, uint32_t L, uint32_t ir_ch, cudaStream_t st) {
    const size_t n = (size_t)ir_ch * RV_BN * reverb_kpad(L);
    reverb_build_toeplitz<<<(unsigned)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096), 256, 0, st>>>(d_ir, static_cast<__nv_bfloat16*>(d_bt), L, reverb_lr(L), ir_ch, reverb_kpad(L));
    return cudaGetLastError();
}

// One call: append the block (bf16) behind the history at `cursor`, run the GEMM over windows ending in it.
// The caller owns the cursor policy (compaction when the buffer is full); cursor is a multiple of 8 and >= Lr.
size_t reverb_ws_bytes() { return (size_t)reverb_grid_max() * RV_BM * RV_BN * sizeof(float); }  // one [128][256] f32 partial per CTA
uint32_t reverb_grid_max() {
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
    return (uint32_t)sms;
}

// Programmatic dependent launch: every kernel here executes griddepcontrol.wait before it touches its predecessor's results,
// so the stream-order chain of completions stays intact while launch latency and prologues overlap.
template <class... KArgs, class... Args>
static cudaError_t launch_pdl_r(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

template <uint32_t BN>
static cudaError_t launch_gemm(const CUtensorMap& tm_a, const CUtensorMap& tm_b, const ReverbGemmArgs& ga, uint32_t grid, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(reverb_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RV_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    return launch_pdl_r(reverb_gemm_kernel<BN>, dim3(grid), dim3(256), RV_SMEM_BYTES, st, tm_a, tm_b, ga);
}

template <uint32_t BN>
static cudaError_t launch_gemm2(const CUtensorMap& tm_a, const CUtensorMap& tm_b, const ReverbGemmArgs& ga, uint32_t pairs_max, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(reverb_gemm2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RV2_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const uint32_t pairs = ga.total_tiles < pairs_max ? ga.total_tiles : pairs_max;
    return launch_pdl_r(reverb_gemm2_kernel<BN>, dim3(2 * pairs), dim3(256), RV2_SMEM_BYTES, st, tm_a, tm_b, ga);  // cluster dims are compiled in
}
// CTA pairs that can be co-resident (a GPC with an odd number of usable SMs strands one)
static uint32_t reverb_pairs_max() {
    static int pairs = 0;
    if (!pairs) {
        cudaFuncSetAttribute(reverb_gemm2_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RV2_SMEM_BYTES);
        cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(2 * reverb_grid_max()); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = RV2_SMEM_BYTES;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, reverb_gemm2_kernel<256>, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = (int)reverb_grid_max() / 2; }
        pairs = n;
    }
    return (uint32_t)pairs;
}

// Tile width for a call: least (waves) x (cycles per k-block). Cycles per k-block = max(MMA, shared-memory traffic at
// 128 B/clk): one CTA per tile 4 x BN/2 against (2 x (16 KB + BN x 128 B)) / 128; a CTA pair 4 x BN/2 against (2 x (16 KB + BN x 64 B)) / 128.
static uint32_t reverb_pick_bn(uint32_t T, uint32_t tiles_mc, uint32_t units, bool pair) {
    static const uint32_t cand[4] = {256, 224, 192, 128};
    uint32_t best = 256; uint64_t best_cost = ~0ull;
    for (uint32_t bn : cand) {
        const uint64_t tiles = (uint64_t)((T + bn - 1) / bn) * tiles_mc, waves = (tiles + units - 1) / units;
        const uint64_t mma = 2ull * bn, smem = pair ? 256ull + bn : 256ull + 2ull * bn;
        const uint64_t cost = waves * ((mma > smem ? mma : smem) + 8u);
        if (cost < best_cost) { best_cost = cost; best = bn; }
    }
    return best;
}

cudaError_t launch_reverb(const ReverbCall& rc, cudaStream_t st, std::string* err) {
    const uint32_t in_pitch = rc.in_pitch ? rc.in_pitch : rc.T, out_pitch = rc.out_pitch ? rc.out_pitch : rc.T;
    {
        const uint32_t per_block = 256 * 8;
        dim3 grid(rc.C * rc.V, (rc.T + per_block - 1) / per_block < 32 ? (rc.T + per_block - 1) / per_block : 32);
        cudaError_t e = launch_pdl_r(reverb_prepare, grid, dim3(256), 0, st, rc.in, static_cast<__nv_bfloat16*>(rc.xh), rc.V, rc.C, rc.T, in_pitch, rc.cursor, rc.pitch, rc.zero_first, rc.chan_base);
        if (e != cudaSuccess) return e;
    }
    const bool pair = rc.V > RV_BM;  // a pair covers 256 voices: with 128 or fewer the second SM would idle
    const uint32_t sms = reverb_grid_max(), tiles_m = pair ? (rc.V + 2 * RV_BM - 1) / (2 * RV_BM) : (rc.V + RV_BM - 1) / RV_BM;
    const uint32_t units = pair ? reverb_pairs_max() : sms;
    const uint32_t bn = reverb_pick_bn(rc.T, tiles_m * rc.C, units, pair);
    const uint32_t kpad = reverb_kpad(rc.L);  // pitch of the Toeplitz rows (built for the widest tile)
    CUtensorMap tm_a, tm_b;
    if (!make_map_bf16_2d(&tm_a, rc.xh, (uint64_t)rc.cursor + rc.T, (uint64_t)(rc.chan_base + rc.C) * rc.V, rc.pitch, RV_BK, RV_BM) ||
        !make_map_bf16_2d(&tm_b, rc.bt, kpad, (uint64_t)rc.ir_ch * RV_BN, kpad, RV_BK, pair ? bn / 2 : bn)) {
        if (err) *err = "cuTensorMapEncodeTiled failed";
        return cudaErrorInvalidValue;
    }
    ReverbGemmArgs ga{};
    ga.out = rc.out; ga.out_pitch = out_pitch; ga.V = rc.V; ga.C = rc.C; ga.T = rc.T; ga.Lr = reverb_lr(rc.L); ga.cursor = rc.cursor; ga.ir_ch = rc.ir_ch;
    ga.num_kb = (reverb_lr(rc.L) + bn + RV_BK - 1) / RV_BK;  // Bt[i][j] is zero for j > Lr + i: a narrower tile has a shorter reduction
    ga.chan_base = rc.chan_base;
    ga.tiles_n = (rc.T + bn - 1) / bn; ga.tiles_m = tiles_m; ga.total_tiles = ga.tiles_n * ga.tiles_m * rc.C;
    if (ga.total_tiles == 0) return cudaSuccess;
    ga.full_tiles = ga.total_tiles; ga.tail_tiles = 0; ga.tail_split = 1; ga.ws = rc.ws; ga.flags = rc.flags; ga.epoch = rc.epoch;
    if (pair && rc.ws && rc.flags && ga.total_tiles > units) {  // a tail wave at most half full is split along K (see ReverbGemmArgs)
        const uint32_t rem = ga.total_tiles % units;
        if (rem && rem * 2u <= units && ga.num_kb >= 64u) { ga.tail_tiles = rem; ga.tail_split = units / rem; ga.full_tiles = ga.total_tiles - rem; }
    }
    if (pair) {
        switch (bn) {
            case 224: return launch_gemm2<224>(tm_a, tm_b, ga, units, st);
            case 192: return launch_gemm2<192>(tm_a, tm_b, ga, units, st);
            case 128: return launch_gemm2<128>(tm_a, tm_b, ga, units, st);
            default: return launch_gemm2<256>(tm_a, tm_b, ga, units, st);
        }
    }
    const uint32_t G = ga.total_tiles < sms ? ga.total_tiles : sms;
    switch (bn) {
        case 224: return launch_gemm<224>(tm_a, tm_b, ga, G, st);
        case 192: return launch_gemm<192>(tm_a, tm_b, ga, G, st);
        case 128: return launch_gemm<128>(tm_a, tm_b, ga, G, st);
        default: return launch_gemm<256>(tm_a, tm_b, ga, G, st);
    }
}

}  // namespace fw
