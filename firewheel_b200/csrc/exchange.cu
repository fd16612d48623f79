// exchange.cu — master-bus exchange between ranks over NVLink peer memory.
//
// Every rank owns a MAILBOX (cudaMalloc'ed, opened by all peers through CUDA IPC):
//     ready[2][16]  u32   ready[q][s] = last epoch whose slot (q, s) sender s has completely written
//     ack[2][16]    u32   ack[q][c]   = last epoch of parity q that consumer c has completely read from ITS OWN mailbox,
//                                       pushed to every sender's mailbox
//     data[2][world][cap] f32         slot (q, s) = sender s's bus of an epoch with parity q, [rows][T]
//
// Per call (epoch e, parity q = e & 1), on every rank:
//   K-push  (MAIN stream, programmatic dependent launch behind the chain / combine kernels) the LAST level of the rank-local bus
//           tree fused with the transfer: each thread finishes its tile of the tree (<= 16 partial buses) and stores it straight
//           into slot (q, me) of all `world` mailboxes — NVLink stores for the peers, a local store for itself — so no local copy
//           of the bus is ever written and the transfer overlaps the tree tile by tile. The last CTA to finish publishes
//           ready[q][me] = e in every mailbox (st.release.sys). No event is recorded on the main stream: the next call's control
//           kernel follows K-push like any other kernel of the chain.
//   K-wait  (SIDE stream, high priority) one warp polls the local ready[q][*] words until all senders published e.
//   K-recv  (SIDE stream) the top log2(world) levels of the same balanced tree in rank order over the local slots, written to the
//           caller's bus buffer — every rank performs the identical additions, so all ranks hold the same bits (an NCCL
//           all-reduce gives no such guarantee). The last CTA acknowledges: ack[q][me] = e in every mailbox.
// The side stream depends on the main stream only through the mailbox words, so the exchange of call e overlaps control + chain
// of call e + 1 with no stream-level hand-over at all. K-push of epoch e + 2 reuses parity q: it first polls the local ack[q][*]
// words for e. Senders never wait on anything but acknowledgements of an epoch two calls back, and receivers only on pushes that
// precede them in every rank's stream order, so the protocol cannot deadlock; a poll that exceeds ~2 s raises the plan's error
// word instead of hanging the GPU.
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.cuh"
#include "plan.hpp"

namespace fw {
namespace {

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// wait until word >= epoch (wrapping compare); false on timeout
__device__ __forceinline__ bool poll_at_least(const uint32_t* word, uint32_t epoch) {
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(word) - epoch) < 0) {
        __nanosleep(64);
        if (clock64() - t0 > 4000000000ll) return false;  // ~2 s at 2 GHz
    }
    return true;
}

template <int VEC> struct V4;
template <> struct V4<4> {
    static __device__ __forceinline__ void load(const float* p, float (&x)[4]) { float4 v = __ldcs(reinterpret_cast<const float4*>(p)); x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w; }
    static __device__ __forceinline__ void load_cg(const float* p, float (&x)[4]) { float4 v = __ldcg(reinterpret_cast<const float4*>(p)); x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w; }
    static __device__ __forceinline__ void store(float* p, const float (&x)[4]) { *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]); }
};
template <> struct V4<1> {
    static __device__ __forceinline__ void load(const float* p, float (&x)[1]) { x[0] = __ldcs(p); }
    static __device__ __forceinline__ void load_cg(const float* p, float (&x)[1]) { x[0] = __ldcg(p); }
    static __device__ __forceinline__ void store(float* p, const float (&x)[1]) { *p = x[0]; }
};

// balanced pairwise tree over p[0..n) (n <= 16), unpaired partials carried up — the same order as combine_kernel
template <int VEC>
__device__ __forceinline__ void tree16(float (&p)[16][VEC], uint32_t n) {
#pragma unroll
    for (int step = 1; step < 16; step <<= 1)
#pragma unroll
        for (int j = 0; j + step < 16; j += 2 * step)
            if ((uint32_t)(j + step) < n) {
#pragma unroll
                for (int i = 0; i < VEC; ++i) p[j][i] = __fadd_rn(p[j][i], p[j + step][i]);
            }
}

template <int VEC>
__global__ void __launch_bounds__(128) bus_push_kernel(const __grid_constant__ BusPushArgs a) {
    __shared__ bool s_last;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // main stream: the next call's control kernel may start
    asm volatile("griddepcontrol.wait;" ::: "memory");               // the partial buses come from the preceding kernel
    const uint32_t q = a.epoch & 1u;
    if (threadIdx.x < a.world) {  // slot (q, me) of every mailbox must have been consumed (epoch - 2)
        if (a.epoch > 2 && !poll_at_least(a.ack_local + q * 16 + threadIdx.x, a.epoch - 2)) *a.error = a.error_value;
    }
    __syncthreads();
    const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) * VEC, row = blockIdx.y, T = a.T;
    if (t < T) {
        float p[16][VEC];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) p[j][i] = 0.0f;
            if ((uint32_t)j < a.n_in) V4<VEC>::load(a.pin + ((size_t)j * a.rows + row) * T + t, p[j]);
        }
        tree16<VEC>(p, a.n_in);
        const size_t off = ((size_t)q * a.world + a.me) * a.cap + (size_t)row * T + t;
        for (uint32_t r = 0; r < a.world; ++r) V4<VEC>::store(a.data[r] + off, p[0]);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(a.counter, 1u) == gridDim.x * gridDim.y - 1;
    __syncthreads();
    if (s_last) {
        __threadfence_system();
        if (threadIdx.x == 0) { *a.counter = 0; st_release_sys(a.push_done, a.epoch); }
        if (threadIdx.x < a.world) st_release_sys(a.ready[threadIdx.x] + q * 16 + a.me, a.epoch);
    }
}

// polls words[0..count) until each is >= epoch
__global__ void __launch_bounds__(32) bus_wait_kernel(const uint32_t* words, uint32_t count, uint32_t epoch, uint32_t* error, uint32_t error_value) {
    if (threadIdx.x < count && !poll_at_least(words + threadIdx.x, epoch)) *error = error_value;
}

template <int VEC>
__global__ void __launch_bounds__(128) bus_recv_kernel(const __grid_constant__ BusRecvArgs a) {
    __shared__ bool s_last;
    const uint32_t q = a.epoch & 1u;
    const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) * VEC, row = blockIdx.y, T = a.T;
    if (t < T) {
        float p[16][VEC];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) p[j][i] = 0.0f;
            // peers wrote these lines over NVLink: read them at L2, never from a stale L1 line of an earlier epoch
            if ((uint32_t)j < a.world) V4<VEC>::load_cg(a.data_local + ((size_t)q * a.world + j) * a.cap + (size_t)row * T + t, p[j]);
        }
        tree16<VEC>(p, a.world);
        V4<VEC>::store(a.out + (size_t)row * a.out_pitch + t, p[0]);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(a.counter, 1u) == gridDim.x * gridDim.y - 1;
    __syncthreads();
    if (s_last) {
        __threadfence_system();
        if (threadIdx.x == 0) *a.counter = 0;
        if (threadIdx.x < a.world) st_release_sys(a.ack[threadIdx.x] + q * 16 + a.me, a.epoch);
    }
}

inline bool vec4_ok(uint32_t T, uint32_t cap, const void* p0, const void* p1) {
    return T % 4 == 0 && cap % 4 == 0 && reinterpret_cast<uintptr_t>(p0) % 16 == 0 && reinterpret_cast<uintptr_t>(p1) % 16 == 0;
}

}  // namespace

cudaError_t launch_bus_push(const BusPushArgs& a, cudaStream_t st) {
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(128); cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (vec4_ok(a.T, a.cap, a.pin, a.data[0])) { cfg.gridDim = dim3((a.T / 4 + 127) / 128, a.rows); return cudaLaunchKernelEx(&cfg, bus_push_kernel<4>, a); }
    cfg.gridDim = dim3((a.T + 127) / 128, a.rows);
    return cudaLaunchKernelEx(&cfg, bus_push_kernel<1>, a);
}
cudaError_t launch_bus_wait(const uint32_t* words, uint32_t count, uint32_t epoch, uint32_t* error, uint32_t error_value, cudaStream_t st) {
    bus_wait_kernel<<<1, 32, 0, st>>>(words, count, epoch, error, error_value);
    return cudaGetLastError();
}
cudaError_t launch_bus_recv(const BusRecvArgs& a, cudaStream_t st) {
    if (vec4_ok(a.T, a.cap | a.out_pitch, a.out, a.data_local)) bus_recv_kernel<4><<<dim3((a.T / 4 + 127) / 128, a.rows), 128, 0, st>>>(a);
    else bus_recv_kernel<1><<<dim3((a.T + 127) / 128, a.rows), 128, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace fw
