// kernels.cu — sm_100a kernels for the per-block audio-graph DSP path.
//
// Bit-exactness rules: this TU is compiled with --fmad=false, -ftz=false,
// -prec-div=true; recurrences additionally spell out __fmul_rn/__fadd_rn. Sum order equals the
// graph's association order: no atomics, no order-agnostic shuffles on sample data.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <utility>

#include "../../include/fw_b200.h"
#include "kernels.cuh"
#include "plan.hpp"

namespace fw {

// =============================================================================================
// K-ctl: per-voice control pass. One thread per voice walks the compiled schedule block by block
// and restates the reference's control logic, emitting records for the data kernels.
//   executor flags      schedule.rs:305-341
//   ParamSmoother       smoother.rs:115-194 (Q1-Q3, Q10 stall)
//   VolumeProcessor     volume.rs:92-110      SumNodeProcessor masks  sum.rs:52-65
//   MonoToStereo        mono_to_stereo.rs:41  StereoToMono stereo_to_mono.rs:41-47
//   HardClip masks      hard_clip.rs:60-93
// Only state transitions and gain curves are computed here; no sample data is touched.
// =============================================================================================
// Programmatic dependent launch (sm_90+): a kernel launched with the PDL attribute may start while its
// predecessor in the stream is still running; it must not touch the predecessor's results before pdl_wait().
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

struct SmLocal { float input, last; uint32_t status; };

__device__ __forceinline__ uint64_t all_silent_mask(uint32_t n) { return n >= 64 ? ~0ull : ((1ull << n) - 1ull); }
__device__ __forceinline__ bool all_channels_silent(uint64_t m, uint32_t n) { uint64_t a = all_silent_mask(n); return (m & a) == a; }

__device__ __forceinline__ void sm_reset(SmLocal& s, float val, bool& changed) {  // smoother.rs:115-129
    if (s.status != SM_INACTIVE) { s.status = SM_INACTIVE; s.input = val; s.last = val; changed = true; }
    else if (s.input != val) { s.input = val; s.last = val; changed = true; }
}

// set_and_process (smoother.rs:133-140,159-194). Returns the record mode; *first = values[0].
__device__ __forceinline__ uint32_t sm_set_and_process(SmLocal& s, float val, uint32_t frames, float a, float b, float eps,
                                                       float* curve, float* const_val, bool* smoothing, bool& changed) {
    if (!(s.input == val)) { s.input = val; s.status = SM_ACTIVE; changed = true; }
    if (s.status != SM_ACTIVE || frames == 0) {  // Q1: the constant buffer (== input for any non-Active state)
        *const_val = s.input; *smoothing = s.status != SM_INACTIVE;
        return REC_CONST;
    }
    const float t = __fmul_rn(s.input, a);
    const float y0 = __fadd_rn(t, __fmul_rn(s.last, b));
    *smoothing = true;
    if (fabsf(__fsub_rn(s.input, y0)) < eps) {  // Q3: settle test on output[0]; reset() overwrites the whole curve
        s.last = s.input; s.status = SM_DEACTIVATING;  // Q2: stays Deactivating forever
        changed = true;
        *const_val = s.input;
        return REC_CONST;
    }
    if (y0 == s.last) {  // Q10: f32 fixed point outside epsilon — the curve is constant and no state changes
        *const_val = y0;
        return REC_CONST;
    }
    float y = y0;
    curve[0] = y0;
    for (uint32_t i = 1; i < frames; ++i) { y = __fadd_rn(t, __fmul_rn(y, b)); curve[i] = y; }
    s.last = y;
    changed = true;
    *const_val = y0;
    return REC_CURVE;
}

__global__ void __launch_bounds__(128) control_kernel(const __grid_constant__ ControlArgs a) {
    pdl_launch_dependents();  // the data kernel may start loading samples now; it waits for us before reading records
    pdl_wait();               // the previous call's data kernels still read the record buffers we are about to rewrite
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t V = a.num_voices;
    if (v >= V) return;
    const CtlTables& tb = a.tables;
    const uint32_t NS = tb.n_smoothers, F = a.block_frames;
    const uint32_t n_blocks = (a.frames + F - 1) / F;

    SmLocal sm[kMaxSmoothers];
    for (uint32_t s = 0; s < NS; ++s) { sm[s].input = tb.sm_input[s][v]; sm[s].last = tb.sm_last[s][v]; sm[s].status = tb.sm_status[s][v]; }
    uint64_t flags = a.flags[v];
    uint64_t gout_mask = 0;
    uint32_t steady = 0xffffffffu, k = 0, last_modes = 0;
    float last_vals[kMaxSmoothers];
    for (uint32_t s = 0; s < NS; ++s) last_vals[s] = 0.0f;
    uint64_t last_sum_mask[kMaxSumMasks];
    for (int s = 0; s < kMaxSumMasks; ++s) last_sum_mask[s] = 0;

    for (; k < n_blocks; ++k) {
        if (k >= a.rec.kt_max) { *a.rec.error = 1; break; }
        const uint32_t frames = min(F, a.frames - k * F);
        bool changed = false;  // smoother state moved during this block
        const uint64_t flags0 = flags;  // the block is a pure function of (flags, smoother state): equal at both ends => it replays
        uint32_t modes = 0;
        for (uint32_t n = 0; n < tb.n_nodes; ++n) {
            const CtlNode nd = tb.nodes[n];
            uint64_t in_mask = 0;
            for (uint32_t i = 0; i < nd.n_in; ++i) {  // schedule.rs:305-320
                const uint64_t bit = 1ull << tb.in_buf[nd.in_off + i];
                if (tb.in_clear[nd.in_off + i]) flags |= bit;
                if (flags & bit) in_mask |= 1ull << i;
            }
            if (nd.mask_slot) { a.rec.sum_masks[((size_t)k * a.rec.n_sum_masks + (nd.mask_slot - 1)) * V + v] = in_mask; last_sum_mask[nd.mask_slot - 1] = in_mask; }
            uint64_t out_mask = 0;  // processor.rs:233 NONE_SILENT
            switch (nd.kind) {
                case FW_NODE_VOLUME: {
                    SmLocal& s = sm[nd.sm0];
                    const float g = tb.sm_target[nd.sm0][v];  // volume.rs:92
                    if (all_channels_silent(in_mask, nd.n_in)) {  // volume.rs:94-100
                        sm_reset(s, g, changed);
                        modes |= REC_CLEAR << (2 * nd.sm0);
                        out_mask = all_silent_mask(nd.n_out);
                    } else {
                        float cv; bool smoothing;
                        float* curve = a.rec.curves + ((size_t)(k * NS + nd.sm0) * V + v) * F;
                        uint32_t m = sm_set_and_process(s, g, frames, a.a, a.b, a.eps, curve, &cv, &smoothing, changed);
                        if (!smoothing && cv < 0.00001f) {  // volume.rs:104-108
                            m = REC_CLEAR; out_mask = all_silent_mask(nd.n_out);
                        } else {
                            out_mask = in_mask;  // volume.rs:110
                        }
                        modes |= m << (2 * nd.sm0);
                        a.rec.vals[(size_t)(k * NS + nd.sm0) * V + v] = cv; last_vals[nd.sm0] = cv;
                    }
                    break;
                }
                case FW_NODE_PAN: {
                    const float gl = tb.sm_target[nd.sm0][v], gr = tb.sm_target[nd.sm1][v];
                    if (all_channels_silent(in_mask, nd.n_in)) {
                        sm_reset(sm[nd.sm0], gl, changed); sm_reset(sm[nd.sm1], gr, changed);
                        modes |= (REC_CLEAR << (2 * nd.sm0)) | (REC_CLEAR << (2 * nd.sm1));
                        out_mask = all_silent_mask(nd.n_out);
                    } else {
                        float cv; bool smoothing;
                        uint32_t m = sm_set_and_process(sm[nd.sm0], gl, frames, a.a, a.b, a.eps,
                                                        a.rec.curves + ((size_t)(k * NS + nd.sm0) * V + v) * F, &cv, &smoothing, changed);
                        modes |= m << (2 * nd.sm0);
                        a.rec.vals[(size_t)(k * NS + nd.sm0) * V + v] = cv; last_vals[nd.sm0] = cv;
                        m = sm_set_and_process(sm[nd.sm1], gr, frames, a.a, a.b, a.eps,
                                               a.rec.curves + ((size_t)(k * NS + nd.sm1) * V + v) * F, &cv, &smoothing, changed);
                        modes |= m << (2 * nd.sm1);
                        a.rec.vals[(size_t)(k * NS + nd.sm1) * V + v] = cv; last_vals[nd.sm1] = cv;
                        out_mask = in_mask;
                    }
                    break;
                }
                case FW_NODE_SUM:  // sum.rs:52-65; the unrolled / generic sums never write the mask (Q7)
                    if (all_channels_silent(in_mask, nd.n_in)) out_mask = all_silent_mask(nd.n_out);
                    else if (nd.n_in == nd.n_out) out_mask = in_mask;
                    break;
                case FW_NODE_MONO_TO_STEREO:  // mono_to_stereo.rs:41-44
                    if (in_mask & 1ull) out_mask = all_silent_mask(nd.n_out);
                    break;
                case FW_NODE_STEREO_TO_MONO:  // stereo_to_mono.rs:41-47
                    if (all_channels_silent(in_mask, 2) || nd.n_in < 2 || nd.n_out == 0) out_mask = all_silent_mask(nd.n_out);
                    break;
                case FW_NODE_HARD_CLIP:  // hard_clip.rs:60-80 leaves the mask untouched on the stereo fast path
                    if (!(nd.n_in == 2 && nd.n_out == 2 && (in_mask & 3ull) == 0)) out_mask = in_mask;
                    break;
                default: break;  // dummy / graph_in / graph_out / biquad / delay / reverb: NONE_SILENT
            }
            if (n + 1 == tb.n_nodes) gout_mask = in_mask;  // graph_out is scheduled last (compiler.rs:291)
            for (uint32_t i = 0; i < nd.n_out; ++i) {  // schedule.rs:338-341
                const uint64_t bit = 1ull << tb.out_buf[nd.out_off + i];
                flags = (out_mask >> i) & 1ull ? (flags | bit) : (flags & ~bit);
            }
        }
        a.rec.modes[(size_t)k * V + v] = modes;
        last_modes = modes;
        if (!changed && flags == flags0) { steady = k; break; }  // nothing moved: every later block replays this record
    }
    if (steady == 0xffffffffu) steady = (k == 0 ? 0 : min(k, n_blocks) - 1);
    a.rec.steady_k[v] = steady;
    a.rec.st_modes[v] = last_modes;  // the record of block `steady`, flattened
    for (uint32_t s = 0; s < NS; ++s) a.rec.st_vals[(size_t)s * V + v] = last_vals[s];
    for (uint32_t s = 0; s < a.rec.n_sum_masks; ++s) a.rec.st_sum_masks[(size_t)s * V + v] = last_sum_mask[s];
    a.rec.gout_mask[v] = gout_mask;
    for (uint32_t s = 0; s < NS; ++s) { tb.sm_input[s][v] = sm[s].input; tb.sm_last[s][v] = sm[s].last; tb.sm_status[s][v] = sm[s].status; }
    a.flags[v] = flags;
}

// =============================================================================================
// K-chain: fused pointwise voice chain (+ master-bus tree). Replaces, for every voice at once,
// the executor loop schedule.rs:299-342 over {VolumeProcessor volume.rs:116-143, PanProcessor,
// HardClip hard_clip.rs:70-90, MonoToStereo :46-48, StereoToMono :49-54} and the SumNode tree
// (sum.rs:69-81). Intermediate edges live in registers; HBM sees the input once and either the
// per-voice output or a 1/64-size partial bus.
//
// Mapping: a warp owns kVPW consecutive voices x one 32*VEC-frame tile; a CTA of kWarps warps owns
// kVPC = 64 consecutive voices. All 2*kVPW loads of a thread are issued before first use.
// =============================================================================================
template <int VEC> struct VecT;
template <> struct VecT<4> {
    using type = float4;
    static __device__ __forceinline__ void load(const float* p, float (&x)[4]) { float4 v = __ldcs(reinterpret_cast<const float4*>(p)); x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w; }
    static __device__ __forceinline__ void load_ca(const float* p, float (&x)[4]) { float4 v = __ldg(reinterpret_cast<const float4*>(p)); x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w; }
    static __device__ __forceinline__ void store(float* p, const float (&x)[4]) { __stcs(reinterpret_cast<float4*>(p), make_float4(x[0], x[1], x[2], x[3])); }
};
template <> struct VecT<1> {
    using type = float;
    static __device__ __forceinline__ void load(const float* p, float (&x)[1]) { x[0] = __ldcs(p); }
    static __device__ __forceinline__ void load_ca(const float* p, float (&x)[1]) { x[0] = __ldg(p); }
    static __device__ __forceinline__ void store(float* p, const float (&x)[1]) { __stcs(p, x[0]); }
};

struct RecView {  // per-voice record access, either from the CTA's smem stage or straight from global
    uint32_t kk, modes;
    const float* vals;  // vals[s * stride]
    uint32_t stride;
};

template <int VEC>
__device__ __forceinline__ void apply_chain(const ChainArgs& a, const RecView& r, uint32_t v, uint32_t t_in_block, float (&x)[2][VEC]) {
    const uint32_t NS = a.rec.n_smoothers, V = a.num_voices, F = a.block_frames;
#pragma unroll 1
    for (uint32_t o = 0; o < a.prog.n_ops; ++o) {
        const ChainOp op = a.prog.ops[o];
        switch (op.kind) {
            case OP_GAIN: {  // volume.rs:116-143: out = in * gain[i] (both channels share the curve)
                const uint32_t m = (r.modes >> (2 * op.sm0)) & 3u;
                if (m == REC_CLEAR) {
#pragma unroll
                    for (int i = 0; i < VEC; ++i) { x[0][i] = 0.0f; x[1][i] = 0.0f; }
                } else if (m == REC_CONST) {
                    const float g = r.vals[op.sm0 * r.stride];
#pragma unroll
                    for (int i = 0; i < VEC; ++i) { x[0][i] = __fmul_rn(x[0][i], g); x[1][i] = __fmul_rn(x[1][i], g); }
                } else {
                    float g[VEC];
                    VecT<VEC>::load_ca(a.rec.curves + ((size_t)(r.kk * NS + op.sm0) * V + v) * F + t_in_block, g);
#pragma unroll
                    for (int i = 0; i < VEC; ++i) { x[0][i] = __fmul_rn(x[0][i], g[i]); x[1][i] = __fmul_rn(x[1][i], g[i]); }
                }
                break;
            }
            case OP_PAN: {
                const uint32_t m0 = (r.modes >> (2 * op.sm0)) & 3u, m1 = (r.modes >> (2 * op.sm1)) & 3u;
                if (m0 == REC_CLEAR) {
#pragma unroll
                    for (int i = 0; i < VEC; ++i) { x[0][i] = 0.0f; x[1][i] = 0.0f; }
                } else {
                    float gl[VEC], gr[VEC];
                    if (m0 == REC_CURVE) VecT<VEC>::load_ca(a.rec.curves + ((size_t)(r.kk * NS + op.sm0) * V + v) * F + t_in_block, gl);
                    else { const float g = r.vals[op.sm0 * r.stride];
#pragma unroll
                        for (int i = 0; i < VEC; ++i) gl[i] = g; }
                    if (m1 == REC_CURVE) VecT<VEC>::load_ca(a.rec.curves + ((size_t)(r.kk * NS + op.sm1) * V + v) * F + t_in_block, gr);
                    else { const float g = r.vals[op.sm1 * r.stride];
#pragma unroll
                        for (int i = 0; i < VEC; ++i) gr[i] = g; }
#pragma unroll
                    for (int i = 0; i < VEC; ++i) { x[0][i] = __fmul_rn(x[0][i], gl[i]); x[1][i] = __fmul_rn(x[1][i], gr[i]); }
                }
                break;
            }
            case OP_CLIP: {  // hard_clip.rs:70-76: in.min(t).max(-t)
                const float th = op.f0;
#pragma unroll
                for (int i = 0; i < VEC; ++i) { x[0][i] = fmaxf(fminf(x[0][i], th), -th); x[1][i] = fmaxf(fminf(x[1][i], th), -th); }
                break;
            }
            case OP_M2S:  // mono_to_stereo.rs:46-48
#pragma unroll
                for (int i = 0; i < VEC; ++i) x[1][i] = x[0][i];
                break;
            case OP_S2M:  // stereo_to_mono.rs:49-54: (l + r) * 0.5
#pragma unroll
                for (int i = 0; i < VEC; ++i) x[0][i] = __fmul_rn(__fadd_rn(x[0][i], x[1][i]), 0.5f);
                break;
            default: break;
        }
    }
}

constexpr int kVPC = 64;  // voices per CTA (one partial bus per CTA in the bus variant)

template <int VEC, int CIN, bool BUS, int kVPW, int kWarps, int kMinBlocks>
__global__ void __launch_bounds__(kWarps * 32, kMinBlocks) chain_kernel(ChainArgs a) {
    static_assert(kVPW * kWarps == kVPC, "a CTA owns 64 voices");
    pdl_launch_dependents();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t T = a.frames, V = a.num_voices, F = a.block_frames, NS = a.rec.n_smoothers;
    const uint32_t c_out = a.prog.c_out;
    constexpr uint32_t kTile = 32u * VEC;
    const uint32_t t = blockIdx.x * kTile + lane * VEC;
    const bool t_ok = t < T;
    const uint32_t vcta = blockIdx.y * kVPC, v0 = vcta + warp * kVPW;
    const uint32_t k = t_ok ? t / F : 0;
    const uint32_t t_in_block = t - k * F;
    // FULL: every voice and every frame of this CTA's tile exists and nothing is zeroed: no bounds checks at all
    const bool full = (vcta + kVPC <= V) && ((blockIdx.x + 1u) * kTile <= T) && !a.zero_first_block;

    // ---- issue every sample load of this thread up front (independent of the control kernel, unless this stage
    //      consumes the output of the preceding kernel) -------------------------------------------------
    if (a.in_from_prev_kernel) pdl_wait();
    float x[kVPW][2][VEC];
    if (full) {
        const float* p[2] = {a.in_ch[0] + (size_t)v0 * a.in_vstride + t, a.in_ch[CIN - 1] + (size_t)v0 * a.in_vstride + t};
#pragma unroll
        for (int j = 0; j < kVPW; ++j) {
#pragma unroll
            for (int c = 0; c < CIN; ++c) { VecT<VEC>::load(p[c], x[j][c]); p[c] += a.in_vstride; }
            if (CIN == 1) {
#pragma unroll
                for (int i = 0; i < VEC; ++i) x[j][1][i] = 0.0f;
            }
        }
    } else {
        const bool zero_in = a.zero_first_block && k == 0;  // Q11: first block after a schedule swap reads a fresh (zero) pool
#pragma unroll
        for (int j = 0; j < kVPW; ++j) {
            const uint32_t v = v0 + j;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
#pragma unroll
                for (int i = 0; i < VEC; ++i) x[j][c][i] = 0.0f;
                if (c < CIN && t_ok && v < V && !zero_in) VecT<VEC>::load(a.in_ch[c < CIN ? c : 0] + (size_t)v * a.in_vstride + t, x[j][c]);
            }
        }
    }

    pdl_wait();  // records written by the control kernel of this call are visible from here on

    // ---- per-warp record staging: one independent load per voice, no CTA-wide barrier ---------------
    // Steady voices (block >= steady_k[v], all smoothers REC_CONST) take the fast path; a warp that owns any
    // transient / CLEAR / CURVE voice, or a tile that straddles block boundaries, takes the generic path.
    __shared__ float s_wv[kWarps][kMaxSmoothers][kVPW];
    const bool cta_uniform = (F % kTile) == 0u;
    bool warp_fast = false;
    if (cta_uniform) {
        const uint32_t kb = (blockIdx.x * kTile) / F;
        uint32_t special = 0;
        if (lane < kVPW && v0 + lane < V) special = (kb < a.rec.steady_k[v0 + lane]) || (a.rec.st_modes[v0 + lane] != 0u);
        for (uint32_t i = lane; i < NS * kVPW; i += 32u) {
            const uint32_t s = i / kVPW, j = i % kVPW;
            s_wv[warp][s][j] = (v0 + j < V) ? a.rec.st_vals[(size_t)s * V + v0 + j] : 0.0f;
        }
        warp_fast = !__any_sync(0xffffffffu, special);
        __syncwarp();
    }

    if (warp_fast) {
        // ops outermost (one uniform decode per op), the warp's voices innermost; multipliers are smem broadcasts
#pragma unroll 1
        for (uint32_t o = 0; o < a.prog.n_ops; ++o) {
            const ChainOp op = a.prog.ops[o];
            if (op.kind == OP_GAIN) {  // volume.rs:123-126
#pragma unroll
                for (int j = 0; j < kVPW; ++j) {
                    const float g = s_wv[warp][op.sm0][j];
#pragma unroll
                    for (int i = 0; i < VEC; ++i) { x[j][0][i] = __fmul_rn(x[j][0][i], g); x[j][1][i] = __fmul_rn(x[j][1][i], g); }
                }
            } else if (op.kind == OP_PAN) {
#pragma unroll
                for (int j = 0; j < kVPW; ++j) {
                    const float gl = s_wv[warp][op.sm0][j], gr = s_wv[warp][op.sm1][j];
#pragma unroll
                    for (int i = 0; i < VEC; ++i) { x[j][0][i] = __fmul_rn(x[j][0][i], gl); x[j][1][i] = __fmul_rn(x[j][1][i], gr); }
                }
            } else if (op.kind == OP_CLIP) {  // hard_clip.rs:70-76
                const float th = op.f0;
#pragma unroll
                for (int j = 0; j < kVPW; ++j)
#pragma unroll
                    for (int i = 0; i < VEC; ++i) { x[j][0][i] = fmaxf(fminf(x[j][0][i], th), -th); x[j][1][i] = fmaxf(fminf(x[j][1][i], th), -th); }
            } else if (op.kind == OP_M2S) {  // mono_to_stereo.rs:46-48
#pragma unroll
                for (int j = 0; j < kVPW; ++j)
#pragma unroll
                    for (int i = 0; i < VEC; ++i) x[j][1][i] = x[j][0][i];
            } else if (op.kind == OP_S2M) {  // stereo_to_mono.rs:49-54
#pragma unroll
                for (int j = 0; j < kVPW; ++j)
#pragma unroll
                    for (int i = 0; i < VEC; ++i) x[j][0][i] = __fmul_rn(__fadd_rn(x[j][0][i], x[j][1][i]), 0.5f);
            }
        }
    } else {
        // rare: park the tile in local memory and run the generic per-voice interpreter over it
        float xs[kVPW][2][VEC];
#pragma unroll
        for (int j = 0; j < kVPW; ++j)
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int i = 0; i < VEC; ++i) xs[j][c][i] = x[j][c][i];
#pragma unroll 1
        for (int j = 0; j < kVPW; ++j) {  // xs is indexed dynamically on purpose: it lives in local memory
            const uint32_t v = v0 + j;
            if (v < V && t_ok) {
                RecView r;
                r.kk = min(k, a.rec.steady_k[v]); r.modes = a.rec.modes[(size_t)r.kk * V + v];
                r.vals = a.rec.vals + (size_t)r.kk * NS * V + v; r.stride = V;
                apply_chain<VEC>(a, r, v, t_in_block, xs[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < kVPW; ++j)
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int i = 0; i < VEC; ++i) x[j][c][i] = xs[j][c][i];
    }

    if (!BUS) {
        if (full) {
            float* q[2] = {a.out_ch[0] + (size_t)v0 * a.out_vstride + t, a.out_ch[1] + (size_t)v0 * a.out_vstride + t};
#pragma unroll
            for (int j = 0; j < kVPW; ++j)
#pragma unroll
                for (int c = 0; c < 2; ++c) if (c < c_out) { VecT<VEC>::store(q[c], x[j][c]); q[c] += a.out_vstride; }
        } else {
#pragma unroll
            for (int j = 0; j < kVPW; ++j) {
                const uint32_t v = v0 + j;
                if (v < V && t_ok) {
#pragma unroll
                    for (int c = 0; c < 2; ++c) if (c < c_out) VecT<VEC>::store(a.out_ch[c] + (size_t)v * a.out_vstride + t, x[j][c]);
                }
            }
        }
    } else {
        // Balanced tree over voices: (2i, 2i+1) per level; a right operand that lies beyond the last voice is
        // skipped (the 1-port SumNode copy, sum.rs:58-65). log2(kVPW) levels in registers per thread.
#pragma unroll
        for (int step = 1; step < kVPW; step <<= 1)
#pragma unroll
            for (int j = 0; j + step < kVPW; j += 2 * step)
                if (full || v0 + j + step < V) {
#pragma unroll
                    for (int c = 0; c < 2; ++c)
#pragma unroll
                        for (int i = 0; i < VEC; ++i) x[j][c][i] = __fadd_rn(x[j][c][i], x[j + step][c][i]);
                }
        // Remaining log2(kWarps) levels through shared memory; each lane owns its own frames. Producer warps
        // arrive on a named barrier and leave; warp c (< c_out) waits, finishes channel c and stores it.
        __shared__ float s_red[kWarps][2][32 * VEC];
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int i = 0; i < VEC; ++i) s_red[warp][c][lane * VEC + i] = x[0][c][i];
        if (warp >= c_out) { asm volatile("bar.arrive 1, %0;" ::"n"(kWarps * 32) : "memory"); return; }
        asm volatile("bar.sync 1, %0;" ::"n"(kWarps * 32) : "memory");
        if (t_ok) {
            const uint32_t c = warp;
            float p[kWarps][VEC];
#pragma unroll
            for (int w = 0; w < kWarps; ++w)
#pragma unroll
                for (int i = 0; i < VEC; ++i) p[w][i] = s_red[w][c][lane * VEC + i];
#pragma unroll
            for (int step = 1; step < kWarps; step <<= 1)
#pragma unroll
                for (int w = 0; w + step < kWarps; w += 2 * step)
                    if (full || vcta + (w + step) * kVPW < V) {
#pragma unroll
                        for (int i = 0; i < VEC; ++i) p[w][i] = __fadd_rn(p[w][i], p[w + step][i]);
                    }
            VecT<VEC>::store(a.out + ((size_t)blockIdx.y * c_out + c) * T + t, p[0]);
        }
    }
}

// K-sum: multi-port SumNode over pool buffers (generic lowering). sum.rs:52-56: all inputs flagged silent -> outputs
// cleared to +0.0 (flagged buffers hold +-0.0, and -0.0 + -0.0 would give -0.0); sum.rs:69-110: ports 2-4 add left to
// right unconditionally; sum.rs:111-133: ports >= 5 start from port 0 and skip ports flagged silent.
template <int VEC>
__global__ void __launch_bounds__(128) sum_kernel(const __grid_constant__ SumArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) * VEC, v = blockIdx.y, T = a.frames, V = a.num_voices;
    if (t >= T) return;
    const size_t off = (size_t)v * T + t;
    uint64_t mask = 0;
    if (a.mask_slot >= 0) {
        const uint32_t k = t / a.block_frames, sk = a.rec.steady_k[v];
        mask = k >= sk ? a.rec.st_sum_masks[(size_t)a.mask_slot * V + v] : a.rec.sum_masks[((size_t)k * a.rec.n_sum_masks + a.mask_slot) * V + v];
    }
    float acc[VEC];
    if (a.mask_slot >= 0 && (mask & a.all_mask) == a.all_mask) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = 0.0f;
        VecT<VEC>::store(a.out + off, acc);
        return;
    }
    VecT<VEC>::load(a.in[0] + off, acc);
    for (uint32_t p = 1; p < a.n_ports; ++p) {
        if (a.skip_silent && ((mask >> a.mask_bit[p]) & 1ull)) continue;
        float x[VEC];
        VecT<VEC>::load(a.in[p] + off, x);
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = __fadd_rn(acc[i], x[i]);
    }
    VecT<VEC>::store(a.out + off, acc);
}

// K-silence-fix (generic lowering): the reference's non-fused node bodies write +0.0 to an output channel whose input
// channel(s) are flagged silent (volume.rs:131-135, hard_clip.rs:78-82, mono_to_stereo.rs:41-44, stereo_to_mono.rs:41-47)
// where the arithmetic on the flagged +-0.0 samples could give -0.0. Rewrites `out` with +0.0 for every (voice, block)
// whose input mask contains all bits of `test`.
template <int VEC>
__global__ void __launch_bounds__(128) silence_fix_kernel(const __grid_constant__ SilenceFixArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) * VEC, v = blockIdx.y, T = a.frames, V = a.num_voices;
    if (t >= T) return;
    const uint32_t k = t / a.block_frames, sk = a.rec.steady_k[v];
    const uint64_t mask = k >= sk ? a.rec.st_sum_masks[(size_t)a.mask_slot * V + v] : a.rec.sum_masks[((size_t)k * a.rec.n_sum_masks + a.mask_slot) * V + v];
    if ((mask & a.test) != a.test) return;
    float z[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) z[i] = 0.0f;
    VecT<VEC>::store(a.out + (size_t)v * T + t, z);
}

// K-combine: radix-16 levels of the same balanced tree over partial buses [n_in][rows][T] -> [ceil(n_in/16)][rows][T].
template <int VEC>
__global__ void __launch_bounds__(128) combine_kernel(const float* __restrict__ pin, float* __restrict__ pout, uint32_t n_in, uint32_t rows, uint32_t T) {
    pdl_launch_dependents();
    pdl_wait();  // the partial buses come from the preceding kernel
    const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (t >= T) return;
    const uint32_t row = blockIdx.y, g = blockIdx.z, p0 = g * 16u;
    float p[16][VEC];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) p[j][i] = 0.0f;
        if (p0 + j < n_in) VecT<VEC>::load(pin + ((size_t)(p0 + j) * rows + row) * T + t, p[j]);
    }
#define FW_COMB1(d, s, first)                                                                  \
    if ((first) < n_in) { _Pragma("unroll") for (int i = 0; i < VEC; ++i) p[d][i] = __fadd_rn(p[d][i], p[s][i]); }
#pragma unroll
    for (int step = 1; step < 16; step <<= 1)
#pragma unroll
        for (int j = 0; j + step < 16; j += 2 * step) FW_COMB1(j, j + step, p0 + j + step)
#undef FW_COMB1
    VecT<VEC>::store(pout + ((size_t)g * rows + row) * T + t, p[0]);
}

// =============================================================================================
// Stream boundary: (de)interleave (util.rs:44-147) for the host-facing process_interleaved.
// =============================================================================================
__global__ void deinterleave_kernel(const float* __restrict__ inter, float* __restrict__ planar, uint32_t V, uint32_t C, uint32_t T) {
    const size_t n = (size_t)V * C * T;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t f = i % T; const size_t vc = i / T; const uint32_t c = vc % C; const size_t v = vc / C;
        planar[i] = inter[(v * T + f) * C + c];
    }
}
// masks: per-voice graph_out silence masks (or one bus mask when rows_per_mask == 0 is not used: n_masks = 1)
__global__ void interleave_kernel(const float* __restrict__ planar, float* __restrict__ inter, const uint64_t* __restrict__ masks,
                                  uint32_t V, uint32_t C, uint32_t T, uint32_t block_frames) {
    const size_t n = (size_t)V * C * T;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t c = i % C; const size_t vf = i / C; const uint32_t f = vf % T; const size_t v = vf / T;
        // The mask handed to interleave is the last block's; silent-flagged buffers hold +0.0 anyway
        // (util.rs:171), so applying it to the final block only is value-identical for earlier blocks.
        const uint64_t m = masks ? masks[v] : 0ull;
        const bool last_block = f >= ((T - 1) / block_frames) * block_frames;
        bool silent;
        if (C == 2) silent = (m & 3ull) == 3ull;            // interleave_stereo util.rs:129-134
        else silent = c < 64 && ((m >> c) & 1ull);           // interleave util.rs:103-107
        inter[i] = (silent && last_block) ? 0.0f : planar[(v * C + c) * T + f];
    }
}

__global__ void fill_kernel(float* __restrict__ p, size_t n, float val) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = val;
}

// bus mask: V == 1 -> the voice's mask; else ALL(n_out) iff every voice is all-silent (2-port SumNode tree, sum.rs:52-56)
__global__ void bus_mask_kernel(const uint64_t* __restrict__ gout_mask, uint32_t V, uint32_t n_out, uint64_t* __restrict__ bus_mask) {
    __shared__ int any_audible;
    if (threadIdx.x == 0) any_audible = 0;
    __syncthreads();
    const uint64_t all = all_silent_mask(n_out);
    for (uint32_t v = threadIdx.x; v < V; v += blockDim.x) if ((gout_mask[v] & all) != all) any_audible = 1;
    __syncthreads();
    if (threadIdx.x == 0) *bus_mask = (V == 1) ? gout_mask[0] : (any_audible ? 0ull : all);
}

// =============================================================================================
// launchers
// =============================================================================================
static inline unsigned grid_for(size_t n) { size_t b = (n + 255) / 256; return (unsigned)(b < 148u * 16u ? b : 148u * 16u); }
#define FW_LAUNCH_CHECK() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return e_; } while (0)

// All three per-call kernels are launched with programmatic stream serialization: each begins with
// griddepcontrol.launch_dependents and reads its predecessor's results only after griddepcontrol.wait.
template <class... KArgs, class... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool no_pdl = getenv("FW_NO_PDL") != nullptr;  // A/B knob
    cfg.attrs = attr; cfg.numAttrs = no_pdl ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

cudaError_t launch_control(const ControlArgs& a, cudaStream_t st) {
    const uint32_t threads = 128, blocks = (a.num_voices + threads - 1) / threads;
    return launch_pdl(control_kernel, dim3(blocks), dim3(threads), st, a);
}

static int chain_variant() {  // A/B knob for tuning runs; the default is what ships
    static int v = -1;
    if (v < 0) { const char* e = getenv("FW_CHAIN_VARIANT"); v = e ? atoi(e) : 0; }
    return v;
}
template <int VEC, int CIN, int VPW, int WARPS, int MINB>
static cudaError_t launch_chain_v(const ChainArgs& a, bool bus, cudaStream_t st) {
    dim3 grid((a.frames + 32 * VEC - 1) / (32 * VEC), (a.num_voices + kVPC - 1) / kVPC);
    if (bus) return launch_pdl(chain_kernel<VEC, CIN, true, VPW, WARPS, MINB>, grid, dim3(WARPS * 32), st, a);
    return launch_pdl(chain_kernel<VEC, CIN, false, VPW, WARPS, MINB>, grid, dim3(WARPS * 32), st, a);
}
template <int VEC, int CIN>
static cudaError_t launch_chain_t(const ChainArgs& a, bool bus, cudaStream_t st) {
    if (VEC == 4) {
        switch (chain_variant()) {
            case 1: return launch_chain_v<VEC, CIN, 8, 8, 2>(a, bus, st);
            case 2: return launch_chain_v<VEC, CIN, 8, 8, 3>(a, bus, st);
            default: return launch_chain_v<VEC, CIN, 4, 16, 2>(a, bus, st);  // 64 regs, 2 x 512 threads per SM
        }
    }
    return launch_chain_v<VEC, CIN, 8, 8, 2>(a, bus, st);
}
cudaError_t launch_chain(const ChainArgs& a, bool bus, cudaStream_t st) {
    const uintptr_t al = reinterpret_cast<uintptr_t>(a.in_ch[0]) | reinterpret_cast<uintptr_t>(a.in_ch[1]) | reinterpret_cast<uintptr_t>(a.out_ch[0]) |
                         reinterpret_cast<uintptr_t>(a.out_ch[1]) | reinterpret_cast<uintptr_t>(a.out) | (uintptr_t)((a.in_vstride | a.out_vstride) * 4);
    const bool vec4 = (a.frames % 4 == 0) && (a.block_frames % 4 == 0) && (al % 16 == 0);
    if (a.prog.c_in == 2) return vec4 ? launch_chain_t<4, 2>(a, bus, st) : launch_chain_t<1, 2>(a, bus, st);
    return vec4 ? launch_chain_t<4, 1>(a, bus, st) : launch_chain_t<1, 1>(a, bus, st);
}
uint32_t chain_voice_groups(uint32_t num_voices) { return (num_voices + kVPC - 1) / kVPC; }

cudaError_t launch_combine(const float* pin, float* pout, uint32_t n_in, uint32_t rows, uint32_t T, cudaStream_t st) {
    const bool vec4 = (T % 4 == 0) && ((reinterpret_cast<uintptr_t>(pin) | reinterpret_cast<uintptr_t>(pout)) % 16 == 0);
    const uint32_t n_out = (n_in + 15) / 16;
    if (vec4) return launch_pdl(combine_kernel<4>, dim3((T / 4 + 127) / 128, rows, n_out), dim3(128), st, pin, pout, n_in, rows, T);
    return launch_pdl(combine_kernel<1>, dim3((T + 127) / 128, rows, n_out), dim3(128), st, pin, pout, n_in, rows, T);
}
cudaError_t launch_sum(const SumArgs& a, cudaStream_t st) {
    uintptr_t al = reinterpret_cast<uintptr_t>(a.out);
    for (uint32_t p = 0; p < a.n_ports; ++p) al |= reinterpret_cast<uintptr_t>(a.in[p]);
    const bool vec4 = (a.frames % 4 == 0) && (a.block_frames % 4 == 0) && (al % 16 == 0);
    if (vec4) return launch_pdl(sum_kernel<4>, dim3((a.frames / 4 + 127) / 128, a.num_voices), dim3(128), st, a);
    return launch_pdl(sum_kernel<1>, dim3((a.frames + 127) / 128, a.num_voices), dim3(128), st, a);
}
cudaError_t launch_silence_fix(const SilenceFixArgs& a, cudaStream_t st) {
    const bool vec4 = (a.frames % 4 == 0) && (a.block_frames % 4 == 0) && (reinterpret_cast<uintptr_t>(a.out) % 16 == 0);
    if (vec4) return launch_pdl(silence_fix_kernel<4>, dim3((a.frames / 4 + 127) / 128, a.num_voices), dim3(128), st, a);
    return launch_pdl(silence_fix_kernel<1>, dim3((a.frames + 127) / 128, a.num_voices), dim3(128), st, a);
}
cudaError_t launch_deinterleave(const float* inter, float* planar, uint32_t V, uint32_t C, uint32_t T, cudaStream_t st) {
    const size_t n = (size_t)V * C * T; if (n == 0) return cudaSuccess;
    deinterleave_kernel<<<grid_for(n), 256, 0, st>>>(inter, planar, V, C, T);
    return cudaGetLastError();
}
cudaError_t launch_interleave(const float* planar, float* inter, const uint64_t* masks, uint32_t V, uint32_t C, uint32_t T, uint32_t block_frames, cudaStream_t st) {
    const size_t n = (size_t)V * C * T; if (n == 0) return cudaSuccess;
    interleave_kernel<<<grid_for(n), 256, 0, st>>>(planar, inter, masks, V, C, T, block_frames);
    return cudaGetLastError();
}
cudaError_t launch_fill(float* p, size_t n, float val, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    fill_kernel<<<grid_for(n), 256, 0, st>>>(p, n, val);
    return cudaGetLastError();
}
cudaError_t launch_bus_mask(const uint64_t* gout_mask, uint32_t V, uint32_t n_out, uint64_t* bus_mask, cudaStream_t st) {
    bus_mask_kernel<<<1, 256, 0, st>>>(gout_mask, V, n_out, bus_mask);
    return cudaGetLastError();
}

}  // namespace fw
