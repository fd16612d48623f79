// fir1_plugin.cu — a third-party node written against include/fw_b200.h's fw_node_vtable (TEST FIXTURE, not product code):
//     y[n] = x[n] - k_v * x[n-1]      per (voice, channel), one float of state, k_v = k0 + 0.01 * voice
// with the SumNode-style silence rule: a block whose inputs are all flagged silent clears its outputs, flags them and
// resets the state. It implements BOTH forms of AudioNodeProcessor::process: `process` (host slices, one voice and one block
// — what the CPU oracle calls) and `process_device` (all voices and blocks of a call, device pointers, on the processor's
// stream — what the product calls). The two must agree bit for bit; tests/test_gpu_plugin.py checks that through both libraries.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/fw_b200.h"

namespace {
uint32_t g_count[5];  // activate, deactivate, drop_node, drop_processor, update

struct Node { float k0; uint32_t rule; int fail_activate; };
struct Proc {
    uint32_t V = 0, C = 0, F = 0; int device = -1; float k0 = 0;
    std::vector<float> h_state;   // [V][C] (oracle side)
    float* d_state = nullptr;     // [V][C] (product side)
};

inline uint64_t all_mask(uint32_t n) { return n >= 64 ? ~0ull : ((1ull << n) - 1ull); }

__global__ void fir1_kernel(const float* const* in_ptrs, float* const* out_ptrs, uint32_t C, uint32_t V, uint32_t T, uint32_t F, uint64_t in_stride,
                            uint64_t out_stride, const uint64_t* masks, const float* state, float k0) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y, c = blockIdx.z;
    if (n >= T) return;
    const uint32_t k = n / F;
    const uint64_t all = C >= 64 ? ~0ull : ((1ull << C) - 1ull);
    const float* x = in_ptrs[c] + (size_t)v * in_stride;
    float* y = out_ptrs[c] + (size_t)v * out_stride;
    if ((masks[(size_t)k * V + v] & all) == all) { y[n] = 0.0f; return; }
    float xp;
    if (n == 0) xp = state[(size_t)v * C + c];
    else if (n % F == 0 && (masks[(size_t)(k - 1) * V + v] & all) == all) xp = 0.0f;  // the previous block reset the state
    else xp = x[n - 1];
    const float kv = __fadd_rn(k0, __fmul_rn(0.01f, (float)v));
    y[n] = __fsub_rn(x[n], __fmul_rn(kv, xp));
}
__global__ void fir1_state_kernel(const float* const* in_ptrs, uint32_t C, uint32_t V, uint32_t T, uint32_t F, uint64_t in_stride, const uint64_t* masks, float* state) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V * C) return;
    const uint32_t v = i / C, c = i % C, kl = (T - 1) / F;
    const uint64_t all = C >= 64 ? ~0ull : ((1ull << C) - 1ull);
    state[i] = (masks[(size_t)kl * V + v] & all) == all ? 0.0f : (in_ptrs[c] + (size_t)v * in_stride)[T - 1];
}

const char* debug_name(void*) { return "fir1_plugin"; }
void info(void* node, fw_audio_node_info* out) {
    std::memset(out, 0, sizeof(*out));
    out->num_min_supported_inputs = 1; out->num_max_supported_inputs = 8; out->num_min_supported_outputs = 1; out->num_max_supported_outputs = 8;
    out->updates = 1; out->out_silence_rule = static_cast<Node*>(node)->rule;
}
int activate(void* node, uint32_t, uint32_t max_block_frames, uint32_t ni, uint32_t no, uint32_t V, int32_t device, void** out, char* err, uint32_t cap) {
    Node* nd = static_cast<Node*>(node);
    if (nd->fail_activate || ni != no) { std::snprintf(err, cap, "fir1_plugin: %s", nd->fail_activate ? "asked to fail" : "inputs must equal outputs"); return 1; }
    Proc* p = new Proc();
    p->V = V; p->C = ni; p->F = max_block_frames; p->device = device; p->k0 = nd->k0;
    if (device >= 0) {
        cudaSetDevice(device);
        if (cudaMalloc(&p->d_state, sizeof(float) * V * ni) != cudaSuccess) { delete p; std::snprintf(err, cap, "fir1_plugin: cudaMalloc failed"); return 2; }
        cudaMemset(p->d_state, 0, sizeof(float) * V * ni);
    } else {
        p->h_state.assign((size_t)V * ni, 0.0f);
    }
    g_count[0]++;
    *out = p;
    return 0;
}
void free_proc(Proc* p) { if (p->d_state) { cudaSetDevice(p->device); cudaFree(p->d_state); } delete p; }
void deactivate(void*, void* proc) { g_count[1]++; if (proc) free_proc(static_cast<Proc*>(proc)); }
void update(void*) { g_count[4]++; }
void drop_node(void* node) { g_count[2]++; delete static_cast<Node*>(node); }
void drop_processor(void* proc) { g_count[3]++; free_proc(static_cast<Proc*>(proc)); }

void process(void* proc, uint32_t voice, uint64_t frames, const float* const* in, uint32_t ni, float* const* out, uint32_t no, fw_proc_info* info) {
    Proc* p = static_cast<Proc*>(proc);
    const uint64_t all = all_mask(ni);
    if ((info->in_silence_mask & all) == all) {
        for (uint32_t c = 0; c < no; ++c) { for (uint64_t n = 0; n < frames; ++n) out[c][n] = 0.0f; p->h_state[(size_t)voice * p->C + c] = 0.0f; }
        *info->out_silence_mask = p->C ? all_mask(no) : 0;
        return;
    }
    const float kv = p->k0 + 0.01f * (float)voice;  // compiled with -ffp-contract=off: separate multiply and add, like the kernel
    for (uint32_t c = 0; c < no; ++c) {
        float xp = p->h_state[(size_t)voice * p->C + c];
        for (uint64_t n = 0; n < frames; ++n) { const float t = kv * xp; out[c][n] = in[c][n] - t; xp = in[c][n]; }
        p->h_state[(size_t)voice * p->C + c] = xp;
    }
}

struct DevPtrs { const float** in; float** out; };
int process_device(void* proc, const fw_device_block* b) {
    Proc* p = static_cast<Proc*>(proc);
    cudaStream_t st = static_cast<cudaStream_t>(b->cuda_stream);
    // the channel pointer arrays live on the host: stage them in device memory for the kernel (stream-ordered, freed after use)
    const float** d_in = nullptr; float** d_out = nullptr;
    if (cudaMallocAsync(&d_in, sizeof(float*) * b->num_inputs, st) != cudaSuccess || cudaMallocAsync(&d_out, sizeof(float*) * b->num_outputs, st) != cudaSuccess) return 1;
    cudaMemcpyAsync(d_in, b->inputs, sizeof(float*) * b->num_inputs, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_out, b->outputs, sizeof(float*) * b->num_outputs, cudaMemcpyHostToDevice, st);
    const uint32_t T = (uint32_t)b->frames;
    fir1_kernel<<<dim3((T + 127) / 128, b->num_voices, b->num_outputs), 128, 0, st>>>(d_in, d_out, b->num_inputs, b->num_voices, T, b->block_frames, b->in_voice_stride,
                                                                                         b->out_voice_stride, b->in_silence_masks, p->d_state, p->k0);
    fir1_state_kernel<<<(b->num_voices * b->num_inputs + 127) / 128, 128, 0, st>>>(d_in, b->num_inputs, b->num_voices, T, b->block_frames, b->in_voice_stride, b->in_silence_masks, p->d_state);
    cudaFreeAsync(d_in, st); cudaFreeAsync(d_out, st);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

const fw_node_vtable g_vt = {debug_name, info, activate, deactivate, update, drop_node, process, process_device, drop_processor};
}  // namespace

extern "C" {
__attribute__((visibility("default"))) const fw_node_vtable* fw_test_plugin_vtable(void) { return &g_vt; }
__attribute__((visibility("default"))) void* fw_test_plugin_new(float k0, uint32_t rule, int fail_activate) { return new Node{k0, rule, fail_activate}; }
__attribute__((visibility("default"))) void fw_test_plugin_counters(uint32_t* out5) { std::memcpy(out5, g_count, sizeof(g_count)); }
__attribute__((visibility("default"))) void fw_test_plugin_reset_counters(void) { std::memset(g_count, 0, sizeof(g_count)); }
}
