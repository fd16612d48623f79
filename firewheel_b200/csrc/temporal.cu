// temporal.cu — sm_100a kernels for the nodes that carry state along time: biquad cascade and
// integer delay line (a12), fused into one pass per voice-channel row.
//
// These nodes are serial recurrences whose results must match the f32 oracle bit for bit, so the op order is the
// oracle's (oracle/fw_oracle.hpp BiquadProcessor / DelayProcessor), spelled with __fmul_rn/__fadd_rn/__fsub_rn:
//     y = (b0*x) + s1;   s1 = ((b1*x) - (a1*y)) + s2;   s2 = (b2*x) - (a2*y)
// Rows (voice-channels) are independent but few (8192 in config 3), and each row is a serial recurrence, so the
// fast kernel spreads the STAGES of a row over adjacent lanes (see biquad_delay_lanes).
// Algorithmic bytes per mono-sample: in 4 + out 4 (+ ring read 4 + ring write 4 with a delay) = 8 / 16.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <type_traits>
#include <utility>

#include "kernels.cuh"
#include "plan.hpp"

namespace fw {

__device__ __forceinline__ void t_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void t_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void cp_async16(float4* smem_dst, const float* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kStateStages = 8;  // state layout [row][8][2] regardless of the cascade length

// Row r of a pass -> its segment (TemporalArgs::seg_rows), the row inside the segment, and its state / ring row.
struct RowMap { uint32_t seg, rr; size_t srow; };
__device__ __forceinline__ RowMap row_map(const TemporalArgs& a, uint32_t r) {
    RowMap m; m.seg = (a.seg_rows != 0u && r >= a.seg_rows) ? 1u : 0u; m.rr = r - m.seg * a.seg_rows;
    m.srow = (size_t)m.rr * a.srow_mul + a.srow_add + m.seg;
    return m;
}
__device__ __forceinline__ const float* row_in(const TemporalArgs& a, const RowMap& m) { return (m.seg ? a.in2 : a.in) + (size_t)m.rr * a.in_pitch; }
__device__ __forceinline__ float* row_out(const TemporalArgs& a, const RowMap& m) { return (m.seg ? a.out2 : a.out) + (size_t)m.rr * a.out_pitch; }

// Fast path: stage-parallel lanes, one self-contained warp per 32/L rows.
//   * A row (voice-channel) is owned by L consecutive lanes; lane s runs biquad stage s. Lane s hands its y to lane
//     s+1 with shfl_up and is skewed by TWO iterations per stage (lane s works on sample n-2s), so the shuffle
//     issued in iteration n is first consumed in iteration n+2: its latency never sits on the recurrence.
//     Same arithmetic per stage as the oracle, only interleaved differently => bit-identical.
//   * Each warp stages its own rows: tiles are [rows][8 x float4] per 32-frame chunk with an XOR swizzle on the
//     16-byte granule (granule g of row r lives at g ^ (r & 7)), so cooperative 16-byte cp.async / STG.128 move
//     full 128-byte row segments and the per-row LDS.128 / STS.128 of the stage lanes are bank-conflict free.
//     Warps never meet at a CTA barrier (only __syncwarp), so they drift freely and cover each other's stalls.
//   * x tiles: 4-deep cp.async pipeline; old-ring tiles: 3-deep; y tiles: double-buffered by chunk parity and
//     flushed (coalesced) two chunks later to the ring (DELAY) or to `out`. With a delay, `out` is the old ring chunk.
// Requires T % 32 == 0, zero_first % 32 == 0 and, with a delay, D % 32 == 0, pos % 32 == 0, D >= 160 (an old-ring
// chunk is read two chunks ahead and must already hold the y flushed D/32 chunks earlier).
//   * CHF = 64: 64-frame chunks (T % 64 == 0, zero_first % 64 == 0, D >= 320). The ~140 instructions of per-chunk bookkeeping
//     (addresses, pipeline slots, flushes) are then paid every 64 samples instead of every 32 — 2.2 instead of 4.5 instructions per
//     sample on a kernel that is bound by its instruction stream. A ring chunk may wrap between its two 32-frame halves.
//   * FULL = true: every row of the CTA exists (the host launches the ragged last CTA separately with FULL = false), so the
//     cooperative copies carry no per-lane predicates or branches.
//   * SVF = true: the per-stage update is the trapezoidal SVF's (include/fw_b200.h) instead of the TDF-II biquad's; the lane /
//     tile / pipeline machinery is identical. Coefficient rows are then 6 floats {a1, a2, a3, m0, m1, m2}, state {ic1, ic2}.
//   * PK = true (RPL == 2): the two rows of a lane run as ONE packed-f32x2 recurrence (Blackwell FFMA2): half the FP
//     instructions per row and two independent chains per dependent-issue slot. ptxas contracts mul.rn.f32x2 + add.rn.f32x2
//     into one FFMA2 even under --fmad=false (single rounding: NOT the oracle's arithmetic), so every packed operation is
//     spelled as an exact fma against OPAQUE constants handed in through the kernel parameters: a*b = fma(a, b, -0.0),
//     a + c = fma(a, 1.0, c), a - b = fma(b, -1.0, a). Each is bit-identical to the separately rounded scalar operation
//     (signed zeros included), and ptxas cannot simplify what it cannot see. Result: bit-exact like the scalar lanes.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pk2(float lo, float hi) { f32x2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float lo2(f32x2_t v) { return __uint_as_float((uint32_t)(v & 0xffffffffull)); }
__device__ __forceinline__ float hi2(f32x2_t v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c) { f32x2_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
template <int NS, int L, bool DELAY, int RPL, bool FULL, bool SVF = false, bool PK = false, int CHF = 32>
__global__ void __launch_bounds__(32) biquad_delay_lanes(TemporalArgs a) {
    static_assert(!PK || (RPL == 2 && NS > 0), "the packed variant pairs the two rows of a lane");
    static_assert(CHF == 32 || CHF == 64, "frames per chunk");
    constexpr uint32_t G = CHF / 4;  // 16-byte granules per tile row
    constexpr uint32_t YM = 1u;                       // y tile slots - 1
    // RPL rows per lane: each lane runs stage s of RPL independent rows.
    constexpr int RSET = 32 / L, ROWS = RPL * RSET, PER = RPL * (CHF / 4) / L, LAG = NS > 0 ? 2 * (NS - 1) : 0;
    static_assert(NS <= L && (L == 1 || L == 2 || L == 4 || L == 8), "lanes per row");
    // No early launch_dependents here: this kernel is issue-bound, and dependents parked at griddepcontrol.wait
    // cost it issue slots (measured: 0.92 vs 0.64 ms per step). The implicit trigger at exit is enough.
    t_pdl_wait();  // `in` is produced by the previous kernel of this call
    const uint32_t lane = threadIdx.x & 31u, s = lane % L;
    const uint32_t row0 = a.row_base + blockIdx.x * ROWS, R = a.R, T = a.T, D = a.D;
    const bool is_first = s == 0, is_last = NS == 0 ? s == 0 : s == (uint32_t)(NS > 0 ? NS - 1 : 0);
    __shared__ float4 xt[4][ROWS][G];
    __shared__ float4 rt[DELAY ? 3 : 1][ROWS][G];
    __shared__ float4 yt[2][ROWS][G];

    uint32_t row_l[RPL], rsw[RPL]; bool lane_ok[RPL], last_ok[RPL];
    float b0[RPL], b1[RPL], b2[RPL], a1[RPL], a2[RPL], c5[RPL], s1[RPL], s2[RPL], q0[RPL], q1[RPL], yb[RPL][4];
#pragma unroll
    for (int j = 0; j < RPL; ++j) {
        row_l[j] = j * RSET + lane / L; rsw[j] = row_l[j] & 7u;
        const uint32_t r = row0 + row_l[j];
        lane_ok[j] = (FULL || r < R) && (NS == 0 ? s == 0 : s < (uint32_t)NS);
        last_ok[j] = is_last && lane_ok[j];
        b0[j] = b1[j] = b2[j] = a1[j] = a2[j] = c5[j] = s1[j] = s2[j] = q0[j] = q1[j] = 0.0f;
        yb[j][0] = yb[j][1] = yb[j][2] = yb[j][3] = 0.0f;
        if (NS > 0 && lane_ok[j]) {
            const RowMap rm = row_map(a, r);
            const float* k = a.coeffs + ((size_t)(rm.rr / a.C) * NS + s) * (SVF ? 6 : 5);
            b0[j] = k[0]; b1[j] = k[1]; b2[j] = k[2]; a1[j] = k[3]; a2[j] = k[4];
            if (SVF) c5[j] = k[5];
            const size_t sr = rm.srow;
            s1[j] = a.state[(sr * kStateStages + s) * 2]; s2[j] = a.state[(sr * kStateStages + s) * 2 + 1];
        }
    }
    // packed copies of the coefficients / state of the lane's two rows (PK), and the opaque constants of the exact-fma spelling
    f32x2_t B0 = 0, B1 = 0, B2 = 0, A1 = 0, A2 = 0, C5 = 0, S1 = 0, S2 = 0, K1 = 0, KN0 = 0, KN1 = 0, K2 = 0;
    if constexpr (PK) {
        B0 = pk2(b0[0], b0[RPL - 1]); B1 = pk2(b1[0], b1[RPL - 1]); B2 = pk2(b2[0], b2[RPL - 1]); A1 = pk2(a1[0], a1[RPL - 1]); A2 = pk2(a2[0], a2[RPL - 1]);
        C5 = pk2(c5[0], c5[RPL - 1]); S1 = pk2(s1[0], s1[RPL - 1]); S2 = pk2(s2[0], s2[RPL - 1]);
        K1 = pk2(a.k_one, a.k_one); KN0 = pk2(a.k_negzero, a.k_negzero); KN1 = pk2(a.k_negone, a.k_negone); K2 = pk2(a.k_two, a.k_two);
    }
    auto mul2 = [&](f32x2_t x, f32x2_t y) { return fma2(x, y, KN0); };   // RN(x*y): adding -0.0 changes nothing, signed zeros included
    auto add2 = [&](f32x2_t x, f32x2_t y) { return fma2(x, K1, y); };    // RN(x+y): x*1.0 is exact
    auto sub2 = [&](f32x2_t x, f32x2_t y) { return fma2(y, KN1, x); };   // RN(x-y): y*-1.0 is exact

    // cooperative copies: lane handles PER (row, granule) pairs of every tile
    const float* in_p[PER]; float* out_p[PER]; float* ring_p[PER]; uint32_t sw[PER]; bool ok[PER], hi[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const uint32_t idx = lane + 32u * i, rr = idx / G, g = idx % G;
        hi[i] = g >= 8u;  // 64-frame chunks: the second 32 frames of a ring chunk may lie beyond the wrap
        ok[i] = FULL || row0 + rr < R;
        const RowMap rm = row_map(a, ok[i] ? row0 + rr : 0u);
        in_p[i] = row_in(a, rm) + g * 4u;
        out_p[i] = row_out(a, rm) + g * 4u;
        ring_p[i] = DELAY ? a.ring + rm.srow * D + (g & 7u) * 4u : nullptr;
        sw[i] = rr * G + (g ^ (rr & 7u));  // float4 index inside a tile
    }
    const uint32_t nch = T / (uint32_t)CHF;
    // Ring offsets and tile slots advance incrementally (ring chunks are issued, consumed and flushed strictly in order):
    // a runtime `% D` costs ~20 dependent instructions through MUFU.RCP, three times per chunk, on a one-warp critical path
    // (measured on config 3: 0.506 -> 0.467 ms per step).
    uint32_t ring_issue_off = DELAY ? a.pos % D : 0u, ring_flush_off = ring_issue_off;  // (pos + 32 * chunk) % D
    uint32_t ring_issue_slot = 0, ring_use_slot = 0;                                   // chunk % 3
    auto advance = [&](uint32_t& off) { off += (uint32_t)CHF; if (off >= D) off -= D; };
    // offset of a lane's granule inside the ring for a chunk that starts at `off` (D % 32 == 0: a chunk wraps only between its halves)
    auto ring_off = [&](uint32_t off, bool second_half) { if (CHF == 32 || !second_half) return off; const uint32_t o = off + 32u; return o >= D ? o - D : o; };
    auto issue = [&](uint32_t chx, uint32_t chr) {  // x tile of chunk chx and old-ring tile of chunk chr, one commit group
        if (chx < nch) {
#pragma unroll
            for (int i = 0; i < PER; ++i) if (FULL || ok[i]) cp_async16(&xt[chx & 3u][0][0] + sw[i], in_p[i] + chx * (uint32_t)CHF);
        }
        if (DELAY && chr < nch) {
#pragma unroll
            for (int i = 0; i < PER; ++i) if (FULL || ok[i]) cp_async16(&rt[ring_issue_slot][0][0] + sw[i], ring_p[i] + ring_off(ring_issue_off, hi[i]));
            advance(ring_issue_off);
            ring_issue_slot = ring_issue_slot == 2u ? 0u : ring_issue_slot + 1u;
        }
        cp_async_commit();  // always one group per call so wait_group counts stay uniform
    };
    auto flush_y = [&](uint32_t ch) {  // y tile of chunk ch -> ring (DELAY) or out, coalesced; called for ch = 0, 1, 2, ... in order
#pragma unroll
        for (int i = 0; i < PER; ++i) if (FULL || ok[i]) {
            const float4 v = (&yt[ch & YM][0][0])[sw[i]];
            if (DELAY) *reinterpret_cast<float4*>(ring_p[i] + ring_off(ring_flush_off, hi[i])) = v;
            else __stcs(reinterpret_cast<float4*>(out_p[i] + ch * (uint32_t)CHF), v);
        }
        if (DELAY) advance(ring_flush_off);
    };

    // One skewed iteration of all RPL rows; gi = global iteration index, u4 = gi & 3 (compile-time when unrolled).
    // q0/q1: inter-stage pipeline — the value shuffled in iteration n is the input of iteration n+2.
    // CHECK = true (first chunk, zeroed chunks, drain): stage s is live only while its sample n = gi - 2s is in [0, T).
    // CHECK = false (every other chunk): all stages are live; lanes that own no stage run on garbage that is never stored.
    // per-lane swizzled granule offsets of this lane's row(s) inside a tile row: granule c lives at c ^ (row & 7)
    uint32_t goff[RPL][G];
#pragma unroll
    for (int j = 0; j < RPL; ++j)
#pragma unroll
        for (int c = 0; c < (int)G; ++c) goff[j][c] = (uint32_t)c ^ rsw[j];
    const float4* xbase[RPL] = {}; float4* ybase[RPL][2] = {};  // refreshed per chunk: x tile row, y tile rows of this / the previous chunk
    auto body = [&](auto check, uint32_t gi, const float (&x)[RPL], int u4, int n4) {
        constexpr bool CHECK = decltype(check)::value;
        const bool in_range = CHECK ? (gi - 2u * s) < T : true;  // unsigned compare: also false during warm-up (gi < 2s)
        const int slot = (u4 - LAG) & 3;  // the last stage emits sample m = gi - LAG; (gi - LAG) & 3 == (u4 - LAG) & 3
        if constexpr (PK) {
            // both rows of the lane in one packed recurrence (same op order per element as the scalar branch below)
            const f32x2_t X = pk2(is_first ? x[0] : q0[0], is_first ? x[1] : q0[1]);
            f32x2_t Y, N1, N2;
            if (SVF) {  // (B0, B1, B2, A1, A2, C5) hold (a1, a2, a3, m0, m1, m2); (S1, S2) hold (ic1, ic2)
                const f32x2_t v3 = sub2(X, S2);
                const f32x2_t v1 = add2(mul2(B0, S1), mul2(B1, v3));
                const f32x2_t v2 = add2(S2, add2(mul2(B1, S1), mul2(B2, v3)));
                N1 = sub2(mul2(K2, v1), S1);
                N2 = sub2(mul2(K2, v2), S2);
                Y = add2(mul2(A1, X), add2(mul2(A2, v1), mul2(C5, v2)));
            } else {
                Y = add2(mul2(B0, X), S1);
                N1 = add2(sub2(mul2(B1, X), mul2(A1, Y)), S2);
                N2 = sub2(mul2(B2, X), mul2(A2, Y));
            }
            if (!CHECK) { S1 = N1; S2 = N2; }
            else {
                const bool act0 = lane_ok[0] && in_range, act1 = lane_ok[1] && in_range;
                S1 = pk2(act0 ? lo2(N1) : lo2(S1), act1 ? hi2(N1) : hi2(S1));
                S2 = pk2(act0 ? lo2(N2) : lo2(S2), act1 ? hi2(N2) : hi2(S2));
            }
            const float y0 = lo2(Y), y1 = hi2(Y);
            q0[0] = q1[0]; q0[1] = q1[1];
            q1[0] = __shfl_up_sync(0xffffffffu, y0, 1); q1[1] = __shfl_up_sync(0xffffffffu, y1, 1);
            yb[0][slot] = y0; yb[1][slot] = y1;
            if (slot == 3) {
                const int ml = n4 * 4 + u4 - LAG;
                if (CHECK ? (is_last && lane_ok[0] && in_range) : last_ok[0]) ybase[0][ml >= 0 ? 0 : 1][goff[0][(ml >> 2) & (int)(G - 1)]] = make_float4(yb[0][0], yb[0][1], yb[0][2], yb[0][3]);
                if (CHECK ? (is_last && lane_ok[1] && in_range) : last_ok[1]) ybase[1][ml >= 0 ? 0 : 1][goff[1][(ml >> 2) & (int)(G - 1)]] = make_float4(yb[1][0], yb[1][1], yb[1][2], yb[1][3]);
            }
        } else {
#pragma unroll
        for (int j = 0; j < RPL; ++j) {
            const bool active = CHECK ? (lane_ok[j] && in_range) : true;
            float y;
            if (NS == 0) {
                y = x[j];
            } else {
                const float xi = is_first ? x[j] : q0[j];
                float n1, n2;
                if (SVF) {  // (b0, b1, b2, a1, a2, c5) hold (a1, a2, a3, m0, m1, m2); (s1, s2) hold (ic1, ic2)
                    const float v3 = __fsub_rn(xi, s2[j]);
                    const float v1 = __fadd_rn(__fmul_rn(b0[j], s1[j]), __fmul_rn(b1[j], v3));
                    const float v2 = __fadd_rn(s2[j], __fadd_rn(__fmul_rn(b1[j], s1[j]), __fmul_rn(b2[j], v3)));
                    n1 = __fsub_rn(__fmul_rn(2.0f, v1), s1[j]);
                    n2 = __fsub_rn(__fmul_rn(2.0f, v2), s2[j]);
                    y = __fadd_rn(__fmul_rn(a1[j], xi), __fadd_rn(__fmul_rn(a2[j], v1), __fmul_rn(c5[j], v2)));
                } else {
                    y = __fadd_rn(__fmul_rn(b0[j], xi), s1[j]);
                    n1 = __fadd_rn(__fsub_rn(__fmul_rn(b1[j], xi), __fmul_rn(a1[j], y)), s2[j]);
                    n2 = __fsub_rn(__fmul_rn(b2[j], xi), __fmul_rn(a2[j], y));
                }
                if (!CHECK || active) { s1[j] = n1; s2[j] = n2; }
                q0[j] = q1[j];
                q1[j] = __shfl_up_sync(0xffffffffu, y, 1);
            }
            yb[j][slot] = y;
            if (slot == 3 && (CHECK ? (is_last && active) : last_ok[j])) {
                // m = gi - LAG lies in this chunk iff n4 * 4 + u4 >= LAG (all compile-time); granule (m >> 2) & 7
                const int ml = n4 * 4 + u4 - LAG;
                ybase[j][ml >= 0 ? 0 : 1][goff[j][(ml >> 2) & (int)(G - 1)]] = make_float4(yb[j][0], yb[j][1], yb[j][2], yb[j][3]);
            }
        }
        }
    };
    auto chunk = [&](auto check, uint32_t ch, bool zero_in) {
#pragma unroll
        for (int j = 0; j < RPL; ++j) {
            xbase[j] = &xt[ch & 3u][row_l[j]][0];
            ybase[j][0] = &yt[ch & YM][row_l[j]][0];
            ybase[j][1] = &yt[(ch + YM) & YM][row_l[j]][0];
        }
        float4 xnext[RPL];  // software-pipelined: the LDS.128 for step n4+1 is issued before step n4 is consumed
#pragma unroll
        for (int j = 0; j < RPL; ++j) xnext[j] = xbase[j][goff[j][0]];
#pragma unroll
        for (uint32_t n4 = 0; n4 < G; ++n4) {
            float4 xq[RPL];  // every lane of a row reads the same granule (broadcast); only stage 0 uses it
#pragma unroll
            for (int j = 0; j < RPL; ++j) {
                xq[j] = xnext[j];
                if (n4 + 1u < G) xnext[j] = xbase[j][goff[j][(n4 + 1u) & (G - 1u)]];
                if (decltype(check)::value && zero_in) xq[j] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
            const uint32_t gi = ch * (uint32_t)CHF + n4 * 4u;
            float x[RPL];
#pragma unroll
            for (int j = 0; j < RPL; ++j) x[j] = xq[j].x;
            body(check, gi, x, 0, (int)n4);
#pragma unroll
            for (int j = 0; j < RPL; ++j) x[j] = xq[j].y;
            body(check, gi + 1u, x, 1, (int)n4);
#pragma unroll
            for (int j = 0; j < RPL; ++j) x[j] = xq[j].z;
            body(check, gi + 2u, x, 2, (int)n4);
#pragma unroll
            for (int j = 0; j < RPL; ++j) x[j] = xq[j].w;
            body(check, gi + 3u, x, 3, (int)n4);
        }
    };

    // groups: G(-3) = {x0}, G(-2) = {x1, ring0}, G(-1) = {x2, ring1}, G(ch) = {x(ch+3), ring(ch+2)}
    issue(0, nch); issue(1, 0); issue(2, 1);
    for (uint32_t ch = 0; ch < nch; ++ch) {
        __syncwarp();  // every lane is done with x tile ch-1, ring tile ch-1 and the y tile about to be flushed
        if (ch >= 2u) flush_y(ch - 2u);
        __syncwarp();  // order the flush's ring stores before the ring loads issued next (they may alias when D is small)
        issue(ch + 3u, ch + 2u);
        cp_async_wait<2>();  // groups up to G(ch-2) have landed: x tile ch and old-ring tile ch
        __syncwarp();
        if (DELAY) {
#pragma unroll
            for (int i = 0; i < PER; ++i) if (FULL || ok[i]) __stcs(reinterpret_cast<float4*>(out_p[i] + ch * (uint32_t)CHF), (&rt[ring_use_slot][0][0])[sw[i]]);
            ring_use_slot = ring_use_slot == 2u ? 0u : ring_use_slot + 1u;
        }
        const bool zero_in = ch * (uint32_t)CHF < a.zero_first;  // Q11 (chunk-uniform)
        if (ch == 0 || zero_in) chunk(std::true_type{}, ch, zero_in);  // warm-up: stage s starts at iteration 2s
        else chunk(std::false_type{}, ch, false);
    }
    if (nch > 0) {
        const float zx[RPL] = {};
#pragma unroll
        for (int j = 0; j < RPL; ++j) { ybase[j][0] = &yt[nch & YM][row_l[j]][0]; ybase[j][1] = &yt[(nch + YM) & YM][row_l[j]][0]; }
#pragma unroll
        for (int it = 0; it < ((LAG + 3) & ~3); ++it) body(std::true_type{}, T + it, zx, it & 3, it >> 2);  // drain (T % 4 == 0)
        __syncwarp();
        if (nch >= 2u) flush_y(nch - 2u);
        flush_y(nch - 1u);
    }
    cp_async_wait<0>();
    if constexpr (PK) { s1[0] = lo2(S1); s1[1] = hi2(S1); s2[0] = lo2(S2); s2[1] = hi2(S2); }
#pragma unroll
    for (int j = 0; j < RPL; ++j) {
        if (NS > 0 && lane_ok[j]) {
            const size_t sr = row_map(a, row0 + row_l[j]).srow;
            a.state[(sr * kStateStages + s) * 2] = s1[j];
            a.state[(sr * kStateStages + s) * 2 + 1] = s2[j];
        }
    }
}

// Generic path: any T, D, pos. One thread per row, scalar, unskewed. Bit-identical results (same per-stage op order).
__global__ void __launch_bounds__(64) biquad_delay_generic(TemporalArgs a) {
    t_pdl_wait();
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.R) return;
    const uint32_t NS = a.ns, T = a.T, D = a.D;
    float b0[kStateStages], b1[kStateStages], b2[kStateStages], a1[kStateStages], a2[kStateStages], s1[kStateStages], s2[kStateStages];
    const RowMap rm = row_map(a, r);
    for (uint32_t s = 0; s < NS; ++s) {
        const float* k = a.coeffs + ((size_t)(rm.rr / a.C) * NS + s) * 5;
        b0[s] = k[0]; b1[s] = k[1]; b2[s] = k[2]; a1[s] = k[3]; a2[s] = k[4];
        s1[s] = a.state[(rm.srow * kStateStages + s) * 2]; s2[s] = a.state[(rm.srow * kStateStages + s) * 2 + 1];
    }
    const float* in = row_in(a, rm);
    float* out = row_out(a, rm);
    float* ring = D ? a.ring + rm.srow * D : nullptr;
    uint32_t p = D ? a.pos % D : 0;
    for (uint32_t n = 0; n < T; ++n) {
        float x = n < a.zero_first ? 0.0f : in[n];
        for (uint32_t s = 0; s < NS; ++s) {
            const float y = __fadd_rn(__fmul_rn(b0[s], x), s1[s]);
            s1[s] = __fadd_rn(__fsub_rn(__fmul_rn(b1[s], x), __fmul_rn(a1[s], y)), s2[s]);
            s2[s] = __fsub_rn(__fmul_rn(b2[s], x), __fmul_rn(a2[s], y));
            x = y;
        }
        if (D) { const float d = ring[p]; ring[p] = x; x = d; p = p + 1 == D ? 0 : p + 1; }
        out[n] = x;
    }
    for (uint32_t s = 0; s < NS; ++s) { const size_t sr = rm.srow; a.state[(sr * kStateStages + s) * 2] = s1[s]; a.state[(sr * kStateStages + s) * 2 + 1] = s2[s]; }
}

// SVF cascade (spec ours, include/fw_b200.h): one thread per row, scalar. State rows as the biquad's: [row][8][2] = {ic1, ic2}.
__global__ void __launch_bounds__(64) svf_generic(TemporalArgs a) {
    t_pdl_wait();
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.R) return;
    const uint32_t NS = a.ns, T = a.T;
    float a1[kStateStages], a2[kStateStages], a3[kStateStages], m0[kStateStages], m1[kStateStages], m2[kStateStages], ic1[kStateStages], ic2[kStateStages];
    const RowMap rm = row_map(a, r);
    const size_t sr = rm.srow;
    for (uint32_t s = 0; s < NS; ++s) {
        const float* k = a.coeffs + ((size_t)(rm.rr / a.C) * NS + s) * 6;
        a1[s] = k[0]; a2[s] = k[1]; a3[s] = k[2]; m0[s] = k[3]; m1[s] = k[4]; m2[s] = k[5];
        ic1[s] = a.state[(sr * kStateStages + s) * 2]; ic2[s] = a.state[(sr * kStateStages + s) * 2 + 1];
    }
    const float* in = row_in(a, rm);
    float* out = row_out(a, rm);
    for (uint32_t n = 0; n < T; ++n) {
        float x = n < a.zero_first ? 0.0f : in[n];
        for (uint32_t s = 0; s < NS; ++s) {
            const float v3 = __fsub_rn(x, ic2[s]);
            const float v1 = __fadd_rn(__fmul_rn(a1[s], ic1[s]), __fmul_rn(a2[s], v3));
            const float v2 = __fadd_rn(ic2[s], __fadd_rn(__fmul_rn(a2[s], ic1[s]), __fmul_rn(a3[s], v3)));
            ic1[s] = __fsub_rn(__fmul_rn(2.0f, v1), ic1[s]);
            ic2[s] = __fsub_rn(__fmul_rn(2.0f, v2), ic2[s]);
            x = __fadd_rn(__fmul_rn(m0[s], x), __fadd_rn(__fmul_rn(m1[s], v1), __fmul_rn(m2[s], v2)));
        }
        out[n] = x;
    }
    for (uint32_t s = 0; s < NS; ++s) { a.state[(sr * kStateStages + s) * 2] = ic1[s]; a.state[(sr * kStateStages + s) * 2 + 1] = ic2[s]; }
}

// Temporal kernels are launched in plain stream order: measured on config 3, programmatic dependent launch made
// the step 0.62 ms instead of 0.51 ms (dependents parked at griddepcontrol.wait compete with an issue-bound kernel).
template <class... KArgs, class... Args>
static cudaError_t launch_pdl_t(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cfg.attrs = nullptr; cfg.numAttrs = 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// Full CTAs of `rows_per_warp` rows + predicated one-row-per-lane CTAs for the ragged tail.
// Packed variant (two rows per lane, FFMA2: 10.9 instead of 16.9 instructions per row and sample): chosen when even at two
// rows per lane every scheduler of the chip (148 SMs x 4) still has two warps to interleave. With fewer rows the kernel is
// bound by the dependent-issue latency of the recurrence, not by issue slots, and one row per lane keeps twice the warps in
// flight (config 3, 8192 rows: packed 0.490 ms, one row per lane 0.456 ms).
template <int NS, int L, bool DELAY, bool SVF, int CHF>
static cudaError_t launch_lanes_c(const TemporalArgs& a, cudaStream_t st) {
    constexpr uint32_t rows1 = 32 / L;
    bool packed = false;
    if constexpr (L > 1 && NS > 0) packed = a.R / (2 * rows1) >= 2 * 4 * 148;  // >= 2 warps per scheduler even at two rows per lane
    uint32_t done = 0;
    if constexpr (L > 1 && NS > 0) {
        if (packed) {
            const uint32_t n_full = a.R / (2 * rows1);
            cudaError_t e = launch_pdl_t(biquad_delay_lanes<NS, L, DELAY, 2, true, SVF, true, 32>, dim3(n_full), dim3(32), st, a);
            if (e != cudaSuccess) return e;
            done = n_full * 2 * rows1;
        }
    }
    if (!packed) {
        const uint32_t n_full = a.R / rows1;
        if (n_full) {
            cudaError_t e = launch_pdl_t(biquad_delay_lanes<NS, L, DELAY, 1, true, SVF, false, CHF>, dim3(n_full), dim3(32), st, a);
            if (e != cudaSuccess) return e;
        }
        done = n_full * rows1;
    }
    if (done < a.R) {  // ragged tail
        TemporalArgs t = a; t.row_base = done;
        return launch_pdl_t(biquad_delay_lanes<NS, L, DELAY, 1, false, SVF, false, 32>, dim3((a.R - done + rows1 - 1) / rows1), dim3(32), st, t);
    }
    return cudaSuccess;
}

template <int NS, int L, bool DELAY, bool SVF = false>
static cudaError_t launch_lanes(const TemporalArgs& a, cudaStream_t st) {
    // 64-frame chunks when the shape allows (see the kernel's comment); only instantiated for the cascade lengths that matter
    if constexpr (L == 4 || L == 2) {
        if (a.T % 64u == 0 && a.zero_first % 64u == 0 && (!DELAY || a.D >= 320u)) return launch_lanes_c<NS, L, DELAY, SVF, 64>(a, st);
    }
    return launch_lanes_c<NS, L, DELAY, SVF, 32>(a, st);
}

bool temporal_fast_path(const TemporalArgs& a) {
    if (a.T == 0 || a.T % 32u || a.zero_first % 32u) return false;
    if ((reinterpret_cast<uintptr_t>(a.in) | reinterpret_cast<uintptr_t>(a.out) | reinterpret_cast<uintptr_t>(a.ring) | reinterpret_cast<uintptr_t>(a.in2) | reinterpret_cast<uintptr_t>(a.out2)) % 16u) return false;
    if ((a.in_pitch | a.out_pitch) % 4u) return false;
    if (a.D && (a.D % 32u || a.pos % 32u || a.D < 160u)) return false;
    return true;
}

cudaError_t launch_temporal(const TemporalArgs& a0, cudaStream_t st) {
    if (a0.R == 0 || a0.T == 0) return cudaSuccess;
    TemporalArgs a = a0;
    if (a.in_pitch == 0) a.in_pitch = a.T;
    if (a.out_pitch == 0) a.out_pitch = a.T;
    a.k_one = 1.0f; a.k_negzero = -0.0f; a.k_negone = -1.0f; a.k_two = 2.0f;  // see the PK note in the kernel's comment
    if (a.svf) {
        if (a.ns >= 1 && a.D == 0 && temporal_fast_path(a)) {
            switch (a.ns) {
                case 1: return launch_lanes<1, 1, false, true>(a, st);
                case 2: return launch_lanes<2, 2, false, true>(a, st);
                case 3: return launch_lanes<3, 4, false, true>(a, st);
                case 4: return launch_lanes<4, 4, false, true>(a, st);
                case 5: return launch_lanes<5, 8, false, true>(a, st);
                case 6: return launch_lanes<6, 8, false, true>(a, st);
                case 7: return launch_lanes<7, 8, false, true>(a, st);
                default: return launch_lanes<8, 8, false, true>(a, st);
            }
        }
        return launch_pdl_t(svf_generic, dim3((a.R + 63) / 64), dim3(64), st, a);
    }
    if (temporal_fast_path(a)) {
#define FW_LANES(NS_, L_) (a.D ? launch_lanes<NS_, L_, true>(a, st) : launch_lanes<NS_, L_, false>(a, st))
        switch (a.ns) {
            case 0: return FW_LANES(0, 1);
            case 1: return FW_LANES(1, 1);
            case 2: return FW_LANES(2, 2);
            case 3: return FW_LANES(3, 4);
            case 4: return FW_LANES(4, 4);
            case 5: return FW_LANES(5, 8);
            case 6: return FW_LANES(6, 8);
            case 7: return FW_LANES(7, 8);
            default: return FW_LANES(8, 8);
        }
#undef FW_LANES
    }
    return launch_pdl_t(biquad_delay_generic, dim3((a.R + 63) / 64), dim3(64), st, a);
}

}  // namespace fw
