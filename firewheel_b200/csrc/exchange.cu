// exchange.cu — hand-over between the processor's main stream and the side stream that carries the master-bus exchange
// between ranks.
//
// The exchange of call e (ncclAllGather of the per-rank buses + the top levels of the tree, runtime.cu: run_bus_stage) runs
// on a high-priority side stream so that it overlaps control + chain of call e + 1. What it needs from the main stream is
// "the rank-local bus of call e is complete". A CUDA event recorded on the main stream would say that, but an event between
// two kernels cuts their programmatic-dependent-launch overlap (measured: ~5 us per call on a 94 us step). Instead:
//   signal    the last CTA of the combine kernel that completes the rank-local bus publishes e in a device word (kernels.cu:
//             combine_kernel, done_word); K-signal below does the same as a one-warp PDL kernel when there is no combine level
//             (<= 64 voices: the chain kernel writes the bus itself);
//   K-wait    (side stream): one warp polls that word until it reaches e; the all-gather is enqueued behind it.
// The main stream therefore carries no event at all in steady state; a poll that exceeds ~2 s raises the plan's error word
// instead of hanging the GPU.
//
// History: rounds 1 and 2 also carried a hand-written peer-memory exchange (CUDA-IPC mailboxes, the last tree level fused
// with NVLink stores into every rank's mailbox, flag words, a receive kernel). It was bit-exact and won narrowly at 2 ranks,
// but lost to NCCL's all-gather at 8 (config 2: 0.120 ms against 0.107 ms per step; with the reverb's persistent CTA pairs
// resident its polling kernels even starved) — the payload is 512 KiB per rank, far below where store bandwidth matters, and
// NCCL's single low-latency kernel beats push + system fence + wait + receive. It was removed rather than kept as an opt-in.
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.cuh"
#include "plan.hpp"

namespace fw {
namespace {

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

__global__ void __launch_bounds__(32) bus_signal_kernel(uint32_t* word, uint32_t epoch) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // the next call's control kernel follows like any kernel of the chain
    asm volatile("griddepcontrol.wait;" ::: "memory");               // the rank-local bus is complete
    if (threadIdx.x == 0) { __threadfence(); st_release_gpu(word, epoch); }
}

// wait until *word >= epoch (wrapping compare)
__global__ void __launch_bounds__(32) bus_wait_kernel(const uint32_t* word, uint32_t epoch, uint32_t* error, uint32_t error_value) {
    if (threadIdx.x != 0) return;
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_gpu(word) - epoch) < 0) {
        __nanosleep(200);
        if (clock64() - t0 > 4000000000ll) { *error = error_value; return; }  // ~2 s at 2 GHz
    }
}

}  // namespace

cudaError_t launch_bus_signal(uint32_t* word, uint32_t epoch, cudaStream_t st) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(1); cfg.blockDim = dim3(32); cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, bus_signal_kernel, word, epoch);
}
cudaError_t launch_bus_wait(const uint32_t* word, uint32_t epoch, uint32_t* error, uint32_t error_value, cudaStream_t st) {
    bus_wait_kernel<<<1, 32, 0, st>>>(word, epoch, error, error_value);
    return cudaGetLastError();
}

}  // namespace fw
