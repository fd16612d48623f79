// fw_oracle_capi.cpp — CPU ORACLE C API (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
//
// Exports include/fw_b200.h with the `fwo_` prefix on top of fw_oracle.hpp. A context
// with num_voices = V is literally V reference contexts (context.rs:29) driven in
// lock-step; the master bus is a balanced tree of the restated 2-port SumNode
// processor (sum.rs:69-81) evaluated per block, exactly as a tree of SumNodes inside
// one reference graph would be.
#define FW_API_PREFIX fwo_
#include "../include/fw_b200.h"
#include "fw_oracle.hpp"

#include <chrono>
#include <array>
#include <functional>
#include <cstdio>
#include <cstdlib>
#include <thread>

using namespace fwo;

static inline fw_node_id pack(Index i) { return (uint64_t)i.slot | ((uint64_t)i.generation << 32); }
static inline Index unpack(uint64_t v) { return Index{(uint32_t)(v & 0xffffffffu), (uint32_t)(v >> 32)}; }
static inline NodeID nid(uint64_t v) { return NodeID{unpack(v), ""}; }

struct fw_ctx {
    fw_graph_config cfg;
    std::vector<std::unique_ptr<FirewheelGraphCtx>> voices;
    std::unique_ptr<CompiledSchedule> dbg_schedule;
    std::string last_error;
    fw_processor* live_processor = nullptr;
    std::vector<std::shared_ptr<const SampleResource>> resources;  // handle - 1
    // ctx_set_event_block: stores / messages stamped with a block offset into the next process call. The reference's processor
    // polls per block (processor.rs:214), so "at block b" is simply: process b blocks, perform the store, go on — which is what
    // the process loop below does with this list.
    uint32_t event_block = 0;
    struct Deferred { uint32_t block; std::function<void()> fn; };
    std::vector<Deferred> deferred;
};
// run `fn` now (event block 0) or when the process loop reaches the stamped block
template <class F> static int defer_or_run(fw_ctx* c, int rc_ok, F&& fn) {
    if (c->event_block == 0 || !c->voices[0]->is_activated()) return fn();
    c->deferred.push_back(fw_ctx::Deferred{c->event_block, [fn]() { (void)fn(); }});
    return rc_ok;
}
static void run_deferred_at(fw_ctx* c, uint32_t block) {
    for (size_t i = 0; i < c->deferred.size(); ++i) if (c->deferred[i].block == block) c->deferred[i].fn();
}
static void rebase_deferred(fw_ctx* c, uint32_t n_blocks) {
    std::vector<fw_ctx::Deferred> keep;
    for (auto& d : c->deferred) if (d.block >= n_blocks) keep.push_back(fw_ctx::Deferred{d.block - n_blocks, d.fn});
    c->deferred.swap(keep);
}
struct fw_processor {
    fw_ctx* ctx;
    std::vector<std::unique_ptr<FirewheelProcessor>> procs;
    uint32_t max_block_frames;
};

static std::unique_ptr<AudioNode> make_node(const fw_node_desc* d) {
    switch (d->kind) {
        case FW_NODE_DUMMY: return std::make_unique<DummyAudioNode>();
        case FW_NODE_VOLUME: return std::make_unique<VolumeNode>(d->f0);
        case FW_NODE_SUM: return std::make_unique<SumNode>();
        case FW_NODE_MONO_TO_STEREO: return std::make_unique<MonoToStereoNode>();
        case FW_NODE_STEREO_TO_MONO: return std::make_unique<StereoToMonoNode>();
        case FW_NODE_HARD_CLIP: return std::make_unique<HardClipNode>(d->f0);
        case FW_NODE_PAN: return std::make_unique<PanNode>(d->f0);
        case FW_NODE_BIQUAD: return std::make_unique<BiquadNode>(d->u0);
        case FW_NODE_DELAY: return std::make_unique<DelayNode>(d->u0);
        case FW_NODE_CONV_REVERB:
            if (!d->data || d->data_len < (uint64_t)d->u0 * d->u1) return nullptr;
            return std::make_unique<ConvReverbNode>(d->data, d->u0, d->u1);
        case FW_NODE_SAMPLER: return std::make_unique<SamplerNode>(d->f0);
        case FW_NODE_SVF: return std::make_unique<SvfNode>(d->u0);
        case FW_NODE_RESAMPLER:
            if (!d->data || d->data_len < (uint64_t)d->u0 * d->u1) return nullptr;
            return std::make_unique<ResamplerNode>(d->data, d->u0, d->u1);
        default: return nullptr;
    }
}
// ---- the plugin boundary (include/fw_b200.h fw_node_vtable) on the oracle: the user node behind the restated traits -----------
// One user node object serves all voices (the batching extension): the V per-voice graphs hold adapters that share it; the
// first adapter to be activated calls the plugin's activate() once for all voices, each adapter's processor calls the
// reference-shaped `process` with its voice index. The declared out_silence_rule is enforced: the product's control plane
// relies on it, so a plugin whose `process` writes a different mask is a plugin bug and trips an assert here.
struct CustomShared {
    fw_node_vtable vt{}; void* node = nullptr; void* proc = nullptr; uint32_t num_voices = 1; fw_audio_node_info info{}; std::string name;
    bool deactivate_on_release = false;
    ~CustomShared() {
        if (proc) { if (deactivate_on_release && vt.deactivate) vt.deactivate(node, proc); else if (vt.drop_processor) vt.drop_processor(proc); }
        if (vt.drop_node) vt.drop_node(node);
    }
};
struct CustomProcessor : AudioNodeProcessor {
    std::shared_ptr<CustomShared> sh; uint32_t voice = 0;
    void process(size_t frames, const std::vector<const float*>& inputs, const std::vector<float*>& outputs, ProcInfo info) override {
        uint64_t out_mask = 0;
        fw_proc_info pi{info.in_silence_mask.bits, &out_mask, info.stream_time_secs, info.stream_status, 0, info.cx};
        sh->vt.process(sh->proc, voice, frames, inputs.data(), (uint32_t)inputs.size(), outputs.data(), (uint32_t)outputs.size(), &pi);
        uint64_t want = 0;
        const SilenceMask in = info.in_silence_mask;
        if (sh->info.out_silence_rule == FW_OUT_SILENCE_PASSTHROUGH) want = in.bits & SilenceMask::new_all_silent(outputs.size()).bits;
        else if (sh->info.out_silence_rule == FW_OUT_SILENCE_ALL_IF_ALL_INPUTS && !inputs.empty() && in.all_channels_silent(inputs.size())) want = SilenceMask::new_all_silent(outputs.size()).bits;
        if (out_mask != want) { std::fprintf(stderr, "fw_oracle: custom node '%s' wrote out_silence_mask %llx, its declared rule gives %llx\n", sh->name.c_str(), (unsigned long long)out_mask, (unsigned long long)want); std::abort(); }
        *info.out_silence_mask = SilenceMask{out_mask};
    }
};
struct CustomAudioNode : AudioNode {
    std::shared_ptr<CustomShared> sh; uint32_t voice = 0;
    const char* debug_name() const override { return sh->name.c_str(); }
    AudioNodeInfo info() const override {
        AudioNodeInfo i; i.num_min_supported_inputs = sh->info.num_min_supported_inputs; i.num_max_supported_inputs = sh->info.num_max_supported_inputs;
        i.num_min_supported_outputs = sh->info.num_min_supported_outputs; i.num_max_supported_outputs = sh->info.num_max_supported_outputs; i.updates = sh->info.updates != 0;
        return i;
    }
    std::unique_ptr<AudioNodeProcessor> activate(uint32_t sample_rate, size_t max_block_frames, size_t num_inputs, size_t num_outputs, std::string* err) override {
        if (!sh->vt.process) { if (err) *err = "custom node has no host `process`: the oracle cannot run it"; return nullptr; }
        if (!sh->proc) {
            char msg[256] = {0};
            void* pr = nullptr;
            if (sh->vt.activate(sh->node, sample_rate, (uint32_t)max_block_frames, (uint32_t)num_inputs, (uint32_t)num_outputs, sh->num_voices, -1, &pr, msg, sizeof(msg) - 1) != 0 || !pr) {
                if (err) *err = msg[0] ? msg : "custom node activation failed";
                return nullptr;
            }
            sh->proc = pr;
        }
        auto p = std::make_unique<CustomProcessor>(); p->sh = sh; p->voice = voice;
        return p;
    }
    void deactivate(std::unique_ptr<AudioNodeProcessor> p) override { if (p) sh->deactivate_on_release = true; }  // the shared processor goes with the last adapter
    void update() override { if (voice == 0 && sh->vt.update) sh->vt.update(sh->node); }
};
fw_node_id fwo_graph_add_custom_node(fw_ctx* c, uint32_t ni, uint32_t no, const fw_node_vtable* vt, void* node) {
    if (!c || !vt || ni > 64 || no > 64 || !vt->debug_name || !vt->info || !vt->activate) {
        if (vt && vt->drop_node) vt->drop_node(node);
        if (c) c->last_error = "bad custom node (vtable needs debug_name, info and activate)";
        return FW_ID_DANGLING;
    }
    auto sh = std::make_shared<CustomShared>();
    sh->vt = *vt; sh->node = node; sh->num_voices = (uint32_t)c->voices.size();
    const char* nm = vt->debug_name(node); sh->name = nm ? nm : "custom";
    vt->info(node, &sh->info);
    fw_node_id id = FW_ID_DANGLING;
    for (size_t v = 0; v < c->voices.size(); ++v) {
        auto n = std::make_unique<CustomAudioNode>(); n->sh = sh; n->voice = (uint32_t)v;
        id = pack(c->voices[v]->graph.add_node(ni, no, std::move(n)).idx);
    }
    return id;
}

static uint32_t kind_of(const AudioNode* n) {
    if (dynamic_cast<const CustomAudioNode*>(n)) return FW_NODE_CUSTOM;
    if (dynamic_cast<const VolumeNode*>(n)) return FW_NODE_VOLUME;
    if (dynamic_cast<const SumNode*>(n)) return FW_NODE_SUM;
    if (dynamic_cast<const MonoToStereoNode*>(n)) return FW_NODE_MONO_TO_STEREO;
    if (dynamic_cast<const StereoToMonoNode*>(n)) return FW_NODE_STEREO_TO_MONO;
    if (dynamic_cast<const HardClipNode*>(n)) return FW_NODE_HARD_CLIP;
    if (dynamic_cast<const PanNode*>(n)) return FW_NODE_PAN;
    if (dynamic_cast<const BiquadNode*>(n)) return FW_NODE_BIQUAD;
    if (dynamic_cast<const DelayNode*>(n)) return FW_NODE_DELAY;
    if (dynamic_cast<const ConvReverbNode*>(n)) return FW_NODE_CONV_REVERB;
    if (dynamic_cast<const SamplerNode*>(n)) return FW_NODE_SAMPLER;
    if (dynamic_cast<const SvfNode*>(n)) return FW_NODE_SVF;
    if (dynamic_cast<const ResamplerNode*>(n)) return FW_NODE_RESAMPLER;
    return FW_NODE_DUMMY;
}
template <class F> static void each_voice(fw_ctx* c, uint32_t voice, F&& f) {
    if (voice == FW_ALL_VOICES) { for (auto& v : c->voices) f(*v); }
    else if (voice < c->voices.size()) f(*c->voices[voice]);
}
static void write_edges(const std::vector<EdgeID>& rm, fw_edge_id* out, uint32_t cap, uint32_t* n) {
    if (n) *n = (uint32_t)rm.size();
    for (size_t i = 0; i < rm.size() && i < cap && out; ++i) out[i] = pack(rm[i].idx);
}

template <class F> static int each_sampler(fw_ctx* c, fw_node_id node, uint32_t voice, F&& f) {
    if (!c || (voice != FW_ALL_VOICES && voice >= c->voices.size())) return FW_SAMPLER_NOT_A_SAMPLER;
    int rc = FW_SAMPLER_NOT_A_SAMPLER; bool first = true;
    each_voice(c, voice, [&](FirewheelGraphCtx& v) {
        auto* n = dynamic_cast<SamplerNode*>(v.graph.node(nid(node)));
        int r = n ? f(*n) : FW_SAMPLER_NOT_A_SAMPLER;
        if (first || r != 0) { if (first || rc == 0) rc = r; first = false; }
    });
    return rc;
}
extern "C" {

void fwo_graph_config_default(fw_graph_config* c) { *c = fw_graph_config{0, 2, 64, 256, 1, 0, 0, 0}; }

fw_ctx* fwo_ctx_new(const fw_graph_config* cfg) {
    if (!cfg || cfg->num_voices == 0 || cfg->num_graph_inputs > 64 || cfg->num_graph_outputs > 64) return nullptr;
    auto* c = new fw_ctx();
    c->cfg = *cfg;
    AudioGraphConfig g{cfg->num_graph_inputs, cfg->num_graph_outputs, cfg->initial_node_capacity, cfg->initial_edge_capacity};
    for (uint32_t v = 0; v < cfg->num_voices; ++v) c->voices.push_back(std::make_unique<FirewheelGraphCtx>(g));
    return c;
}
void fwo_ctx_free(fw_ctx* c) {
    if (!c) return;
    if (c->live_processor) { fw_processor* p = c->live_processor; c->live_processor = nullptr; p->procs.clear(); delete p; }
    for (auto& v : c->voices) if (v->is_activated()) v->deactivate(false);
    delete c;
}
const char* fwo_ctx_last_error(fw_ctx* c) { return c ? c->last_error.c_str() : "null context"; }
void fwo_ctx_set_event_block(fw_ctx* c, uint32_t block) { if (c) c->event_block = block; }
fw_node_id fwo_graph_in_node(fw_ctx* c) { return pack(c->voices[0]->graph.graph_in_node().idx); }
fw_node_id fwo_graph_out_node(fw_ctx* c) { return pack(c->voices[0]->graph.graph_out_node().idx); }

fw_node_id fwo_graph_add_node(fw_ctx* c, uint32_t ni, uint32_t no, const fw_node_desc* d) {
    if (!c || !d || ni > 64 || no > 64) return FW_ID_DANGLING;
    fw_node_id id = FW_ID_DANGLING;
    for (auto& v : c->voices) {
        auto n = make_node(d);
        if (!n) { c->last_error = "bad node description"; return FW_ID_DANGLING; }
        id = pack(v->graph.add_node(ni, no, std::move(n)).idx);
    }
    return id;
}
int fwo_graph_remove_node(fw_ctx* c, fw_node_id node, fw_edge_id* removed, uint32_t cap, uint32_t* n_removed) {
    int rc = 0; std::vector<EdgeID> rm0;
    for (size_t v = 0; v < c->voices.size(); ++v) {
        std::vector<EdgeID> rm;
        if (!c->voices[v]->graph.remove_node(nid(node), &rm)) rc = -1;
        if (v == 0) rm0 = rm;
    }
    if (rc == 0) write_edges(rm0, removed, cap, n_removed); else if (n_removed) *n_removed = 0;
    return rc;
}
int fwo_graph_set_num_inputs(fw_ctx* c, fw_node_id node, uint32_t n, fw_edge_id* removed, uint32_t cap, uint32_t* n_removed) {
    if (n > 64) return -1;
    int rc = 0; std::vector<EdgeID> rm0;
    for (size_t v = 0; v < c->voices.size(); ++v) {
        std::vector<EdgeID> rm;
        if (!c->voices[v]->graph.set_num_inputs(nid(node), n, &rm)) rc = -1;
        if (v == 0) rm0 = rm;
    }
    if (rc == 0) write_edges(rm0, removed, cap, n_removed); else if (n_removed) *n_removed = 0;
    return rc;
}
int fwo_graph_set_num_outputs(fw_ctx* c, fw_node_id node, uint32_t n, fw_edge_id* removed, uint32_t cap, uint32_t* n_removed) {
    if (n > 64) return -1;
    int rc = 0; std::vector<EdgeID> rm0;
    for (size_t v = 0; v < c->voices.size(); ++v) {
        std::vector<EdgeID> rm;
        if (!c->voices[v]->graph.set_num_outputs(nid(node), n, &rm)) rc = -1;
        if (v == 0) rm0 = rm;
    }
    if (rc == 0) write_edges(rm0, removed, cap, n_removed); else if (n_removed) *n_removed = 0;
    return rc;
}
int fwo_graph_connect(fw_ctx* c, fw_node_id src, uint32_t sp, fw_node_id dst, uint32_t dp, int check, fw_edge_id* out_edge,
                      fw_node_id* err_node, uint32_t* err_port) {
    AddEdgeError e0 = AddEdgeError::Ok; EdgeID id0{};
    for (size_t v = 0; v < c->voices.size(); ++v) {
        EdgeID id{};
        AddEdgeError e = c->voices[v]->graph.connect(nid(src), sp, nid(dst), dp, check != 0, &id);
        if (v == 0) { e0 = e; id0 = id; }
    }
    if (e0 == AddEdgeError::Ok) { if (out_edge) *out_edge = pack(id0.idx); return FW_EDGE_OK; }
    if (err_node) {
        switch (e0) {
            case AddEdgeError::SrcNodeNotFound: case AddEdgeError::OutPortOutOfRange: *err_node = src; break;
            case AddEdgeError::DstNodeNotFound: case AddEdgeError::InPortOutOfRange: case AddEdgeError::InputPortAlreadyConnected: *err_node = dst; break;
            default: *err_node = FW_ID_DANGLING;
        }
    }
    if (err_port) *err_port = (e0 == AddEdgeError::OutPortOutOfRange) ? sp : dp;
    return (int)e0;
}
int fwo_graph_disconnect(fw_ctx* c, fw_node_id src, uint32_t sp, fw_node_id dst, uint32_t dp) {
    int r = 0;
    for (size_t v = 0; v < c->voices.size(); ++v) { bool b = c->voices[v]->graph.disconnect(nid(src), sp, nid(dst), dp); if (v == 0) r = b; }
    return r;
}
int fwo_graph_disconnect_by_edge_id(fw_ctx* c, fw_edge_id e) {
    int r = 0;
    for (size_t v = 0; v < c->voices.size(); ++v) { bool b = c->voices[v]->graph.disconnect_by_edge_id(EdgeID{unpack(e)}); if (v == 0) r = b; }
    return r;
}
int fwo_graph_edge(fw_ctx* c, fw_edge_id e, fw_edge_info* out) {
    const Edge* ed = c->voices[0]->graph.edge(EdgeID{unpack(e)});
    if (!ed) return 0;
    if (out) *out = fw_edge_info{pack(ed->id.idx), pack(ed->src_node.idx), pack(ed->dst_node.idx), ed->src_port, ed->dst_port};
    return 1;
}
int fwo_graph_node_info(fw_ctx* c, fw_node_id node, fw_node_info* out) {
    const NodeEntry* ne = c->voices[0]->graph.node_info(nid(node));
    if (!ne) return 0;
    if (out) {
        std::memset(out, 0, sizeof(*out));
        AudioNodeInfo i = ne->weight.node->info();
        out->num_inputs = ne->num_inputs; out->num_outputs = ne->num_outputs; out->kind = kind_of(ne->weight.node.get());
        out->num_min_supported_inputs = i.num_min_supported_inputs; out->num_max_supported_inputs = i.num_max_supported_inputs;
        out->num_min_supported_outputs = i.num_min_supported_outputs; out->num_max_supported_outputs = i.num_max_supported_outputs;
        out->updates = i.updates;
        std::strncpy(out->debug_name, ne->id.debug_name, sizeof(out->debug_name) - 1);
    }
    return 1;
}
uint32_t fwo_graph_num_nodes(fw_ctx* c) { return (uint32_t)c->voices[0]->graph.nodes.len(); }
uint32_t fwo_graph_num_edges(fw_ctx* c) { return (uint32_t)c->voices[0]->graph.edges.len(); }
uint32_t fwo_graph_nodes(fw_ctx* c, fw_node_id* out, uint32_t cap) {
    uint32_t n = 0;
    c->voices[0]->graph.nodes.for_each([&](Index i, NodeEntry&) { if (out && n < cap) out[n] = pack(i); ++n; });
    return n;
}
uint32_t fwo_graph_edges(fw_ctx* c, fw_edge_id* out, uint32_t cap) {
    uint32_t n = 0;
    c->voices[0]->graph.edges.for_each([&](Index i, Edge&) { if (out && n < cap) out[n] = pack(i); ++n; });
    return n;
}
int fwo_graph_cycle_detected(fw_ctx* c) { int r = 0; for (size_t v = 0; v < c->voices.size(); ++v) { bool b = c->voices[v]->graph.cycle_detected(); if (v == 0) r = b; } return r; }
void fwo_graph_reset(fw_ctx* c) { for (auto& v : c->voices) v->graph.reset(); }
int fwo_graph_needs_compile(fw_ctx* c) { return c->voices[0]->graph.needs_compile(); }

int fwo_graph_compile_internal(fw_ctx* c, uint32_t max_block_frames) {
    if (max_block_frames == 0) return FW_COMPILE_NODE_ACTIVATION_FAILED;
    c->dbg_schedule.reset();
    CompileErrorInfo e = c->voices[0]->graph.compile_internal(max_block_frames, &c->dbg_schedule);
    return (int)e.code;
}
uint32_t fwo_schedule_len(fw_ctx* c) { return c->dbg_schedule ? (uint32_t)c->dbg_schedule->schedule.size() : 0; }
uint32_t fwo_schedule_num_buffers(fw_ctx* c) { return c->dbg_schedule ? (uint32_t)c->dbg_schedule->num_buffers : 0; }
int fwo_schedule_node(fw_ctx* c, uint32_t i, fw_scheduled_node* out) {
    if (!c->dbg_schedule || i >= c->dbg_schedule->schedule.size() || !out) return 0;
    const ScheduledNode& sn = c->dbg_schedule->schedule[i];
    std::memset(out, 0, sizeof(*out));
    out->id = pack(sn.id.idx);
    out->num_inputs = (uint32_t)sn.input_buffers.size(); out->num_outputs = (uint32_t)sn.output_buffers.size();
    for (size_t k = 0; k < sn.input_buffers.size() && k < 64; ++k) { out->in_buffer[k] = (uint32_t)sn.input_buffers[k].buffer_index; out->in_should_clear[k] = sn.input_buffers[k].should_clear; }
    for (size_t k = 0; k < sn.output_buffers.size() && k < 64; ++k) out->out_buffer[k] = (uint32_t)sn.output_buffers[k].buffer_index;
    return 1;
}

// ---- parameters -------------------------------------------------------------------------------
static int volume_set_percent_volume_now(fw_ctx* c, fw_node_id node, uint32_t voice, float pct) {
    int ok = -1;
    each_voice(c, voice, [&](FirewheelGraphCtx& v) { if (auto* n = dynamic_cast<VolumeNode*>(v.graph.node(nid(node)))) { n->set_percent_volume(pct); ok = 0; } });
    return ok;
}
int fwo_volume_set_percent_volume(fw_ctx* c, fw_node_id node, uint32_t voice, float pct) { return defer_or_run(c, 0, [=]() { return volume_set_percent_volume_now(c, node, voice, pct); }); }
int fwo_volume_set_percent_volumes(fw_ctx* c, fw_node_id node, const float* pct, uint32_t n) {
    if (n != c->voices.size()) return -1;
    for (uint32_t v = 0; v < n; ++v) if (fwo_volume_set_percent_volume(c, node, v, pct[v]) != 0) return -1;
    return 0;
}
static int pan_set_pan_now(fw_ctx* c, fw_node_id node, uint32_t voice, float pan) {
    int ok = -1;
    each_voice(c, voice, [&](FirewheelGraphCtx& v) { if (auto* n = dynamic_cast<PanNode*>(v.graph.node(nid(node)))) { n->set_pan(pan); ok = 0; } });
    return ok;
}
int fwo_pan_set_pan(fw_ctx* c, fw_node_id node, uint32_t voice, float pan) { return defer_or_run(c, 0, [=]() { return pan_set_pan_now(c, node, voice, pan); }); }
int fwo_pan_set_pans(fw_ctx* c, fw_node_id node, const float* pan, uint32_t n) {
    if (n != c->voices.size()) return -1;
    for (uint32_t v = 0; v < n; ++v) if (fwo_pan_set_pan(c, node, v, pan[v]) != 0) return -1;
    return 0;
}
static int pan_set_gains_now(fw_ctx* c, fw_node_id node, uint32_t voice, float gl, float gr) {
    int ok = -1;
    each_voice(c, voice, [&](FirewheelGraphCtx& v) { if (auto* n = dynamic_cast<PanNode*>(v.graph.node(nid(node)))) { n->set_gains(gl, gr); ok = 0; } });
    return ok;
}
int fwo_pan_set_gains(fw_ctx* c, fw_node_id node, uint32_t voice, float gl, float gr) { return defer_or_run(c, 0, [=]() { return pan_set_gains_now(c, node, voice, gl, gr); }); }
static int biquad_set_coeffs_now(fw_ctx* c, fw_node_id node, uint32_t voice, uint32_t stage, const float* k) {
    int ok = -1;
    each_voice(c, voice, [&](FirewheelGraphCtx& v) {
        if (auto* n = dynamic_cast<BiquadNode*>(v.graph.node(nid(node)))) if (stage < n->params->num_stages) { n->params->st[stage] = BiquadCoeffs{k[0], k[1], k[2], k[3], k[4]}; ok = 0; }
    });
    return ok;
}
int fwo_biquad_set_coeffs(fw_ctx* c, fw_node_id node, uint32_t voice, uint32_t stage, const float* k) { if (!k) return -1; std::array<float, 5> kk; std::copy(k, k + 5, kk.begin()); return defer_or_run(c, 0, [=]() { return biquad_set_coeffs_now(c, node, voice, stage, kk.data()); }); }
int fwo_biquad_set_all_coeffs(fw_ctx* c, fw_node_id node, const float* k, uint32_t nv, uint32_t ns) {
    if (nv != c->voices.size()) return -1;
    for (uint32_t v = 0; v < nv; ++v) for (uint32_t s = 0; s < ns; ++s) if (fwo_biquad_set_coeffs(c, node, v, s, k + ((size_t)v * ns + s) * 5) != 0) return -1;
    return 0;
}
// RBJ Audio-EQ-Cookbook, evaluated in f64, rounded once to f32. (Design helper; the DSP contract is the 5 coefficients.)
void fwo_biquad_design_rbj(uint32_t type, double fc, double q, double gain_db, double sr, float* out) {
    double w0 = 2.0 * M_PI * fc / sr, cw = std::cos(w0), sw = std::sin(w0), alpha = sw / (2.0 * q);
    double A = std::pow(10.0, gain_db / 40.0), b0, b1, b2, a0, a1, a2;
    switch (type) {
        case 0: b0 = (1 - cw) / 2; b1 = 1 - cw; b2 = (1 - cw) / 2; a0 = 1 + alpha; a1 = -2 * cw; a2 = 1 - alpha; break;
        case 1: b0 = (1 + cw) / 2; b1 = -(1 + cw); b2 = (1 + cw) / 2; a0 = 1 + alpha; a1 = -2 * cw; a2 = 1 - alpha; break;
        case 2: b0 = alpha; b1 = 0; b2 = -alpha; a0 = 1 + alpha; a1 = -2 * cw; a2 = 1 - alpha; break;
        case 3: b0 = 1; b1 = -2 * cw; b2 = 1; a0 = 1 + alpha; a1 = -2 * cw; a2 = 1 - alpha; break;
        case 4: b0 = 1 + alpha * A; b1 = -2 * cw; b2 = 1 - alpha * A; a0 = 1 + alpha / A; a1 = -2 * cw; a2 = 1 - alpha / A; break;
        case 5: { double s = 2 * std::sqrt(A) * alpha;
            b0 = A * ((A + 1) - (A - 1) * cw + s); b1 = 2 * A * ((A - 1) - (A + 1) * cw); b2 = A * ((A + 1) - (A - 1) * cw - s);
            a0 = (A + 1) + (A - 1) * cw + s; a1 = -2 * ((A - 1) + (A + 1) * cw); a2 = (A + 1) + (A - 1) * cw - s; break; }
        default: { double s = 2 * std::sqrt(A) * alpha;
            b0 = A * ((A + 1) + (A - 1) * cw + s); b1 = -2 * A * ((A - 1) + (A + 1) * cw); b2 = A * ((A + 1) + (A - 1) * cw - s);
            a0 = (A + 1) - (A - 1) * cw + s; a1 = 2 * ((A - 1) - (A + 1) * cw); a2 = (A + 1) - (A - 1) * cw - s; break; }
    }
    out[0] = (float)(b0 / a0); out[1] = (float)(b1 / a0); out[2] = (float)(b2 / a0); out[3] = (float)(a1 / a0); out[4] = (float)(a2 / a0);
}

// ---- SVF + polyphase resampler (spec ours) ----------------------------------------------------
static int svf_set_coeffs_now(fw_ctx* c, fw_node_id node, uint32_t voice, uint32_t stage, const float* k) {
    int ok = -1;
    each_voice(c, voice, [&](FirewheelGraphCtx& v) {
        if (auto* n = dynamic_cast<SvfNode*>(v.graph.node(nid(node)))) if (stage < n->params->num_stages) { n->params->st[stage] = SvfCoeffs{k[0], k[1], k[2], k[3], k[4], k[5]}; ok = 0; }
    });
    return ok;
}
int fwo_svf_set_coeffs(fw_ctx* c, fw_node_id node, uint32_t voice, uint32_t stage, const float* k) { if (!k) return -1; std::array<float, 6> kk; std::copy(k, k + 6, kk.begin()); return defer_or_run(c, 0, [=]() { return svf_set_coeffs_now(c, node, voice, stage, kk.data()); }); }
int fwo_svf_set_all_coeffs(fw_ctx* c, fw_node_id node, const float* k, uint32_t nv, uint32_t ns) {
    if (nv != c->voices.size()) return -1;
    for (uint32_t v = 0; v < nv; ++v) for (uint32_t s = 0; s < ns; ++s) if (fwo_svf_set_coeffs(c, node, v, s, k + ((size_t)v * ns + s) * 6) != 0) return -1;
    return 0;
}
void fwo_svf_design(uint32_t type, double fc, double q, double sr, float* out) {
    const double g = std::tan(M_PI * fc / sr), k = 1.0 / q;
    const double a1 = 1.0 / (1.0 + g * (g + k)), a2 = g * a1, a3 = g * a2;
    double m0 = 0, m1 = 0, m2 = 1;  // lowpass
    switch (type) {
        case 1: m0 = 0; m1 = 1; m2 = 0; break;          // bandpass
        case 2: m0 = 1; m1 = -k; m2 = -1; break;        // highpass
        case 3: m0 = 1; m1 = -k; m2 = 0; break;         // notch
        case 4: m0 = 1; m1 = -k; m2 = -2; break;        // peak
        case 5: m0 = 1; m1 = -2 * k; m2 = 0; break;     // allpass
        default: break;
    }
    out[0] = (float)a1; out[1] = (float)a2; out[2] = (float)a3; out[3] = (float)m0; out[4] = (float)m1; out[5] = (float)m2;
}
static int resampler_set_now(fw_ctx* c, fw_node_id node, uint32_t voice, uint32_t res, uint64_t step, int playing, int loop) {
    if (!c || res > c->resources.size()) return -1;
    int ok = -1;
    each_voice(c, voice, [&](FirewheelGraphCtx& v) {
        if (auto* n = dynamic_cast<ResamplerNode*>(v.graph.node(nid(node)))) { n->sh->sample = res ? c->resources[res - 1] : nullptr; n->sh->step = step; n->sh->playing = playing != 0; n->sh->loop = loop != 0; ok = 0; }
    });
    return ok;
}
int fwo_resampler_set(fw_ctx* c, fw_node_id node, uint32_t voice, uint32_t res, uint64_t step, int playing, int loop) { return defer_or_run(c, 0, [=]() { return resampler_set_now(c, node, voice, res, step, playing, loop); }); }
static int resampler_seek_now(fw_ctx* c, fw_node_id node, uint32_t voice, uint64_t pos_frames) {
    int ok = -1;
    each_voice(c, voice, [&](FirewheelGraphCtx& v) { if (auto* n = dynamic_cast<ResamplerNode*>(v.graph.node(nid(node)))) { n->sh->seek_pending = true; n->sh->seek_frames = pos_frames; ok = 0; } });
    return ok;
}
int fwo_resampler_seek(fw_ctx* c, fw_node_id node, uint32_t voice, uint64_t pos_frames) { return defer_or_run(c, 0, [=]() { return resampler_seek_now(c, node, voice, pos_frames); }); }
static double bessel_i0(double x) { double s = 1.0, t = 1.0; for (int k = 1; k < 64; ++k) { t *= (x / (2.0 * k)) * (x / (2.0 * k)); s += t; if (t < 1e-18 * s) break; } return s; }
void fwo_resampler_design(uint32_t P, uint32_t T, double cutoff, double beta, float* table) {
    const double half = (double)T / 2.0, i0b = bessel_i0(beta);
    for (uint32_t p = 0; p < P; ++p) for (uint32_t t = 0; t < T; ++t) {
        const double x = (double)t - (half - 1.0) - (double)p / (double)P;   // distance of tap t from the fractional read point
        const double s = x == 0.0 ? 1.0 : std::sin(M_PI * cutoff * x) / (M_PI * cutoff * x);
        const double r = x / half, w = std::fabs(r) >= 1.0 ? 0.0 : bessel_i0(beta * std::sqrt(1.0 - r * r)) / i0b;
        table[(size_t)p * T + t] = (float)(cutoff * s * w);
    }
}

// ---- sample resources + sampler messages ---------------------------------------------------------
uint32_t fwo_sample_resource_create(fw_ctx* c, uint32_t format, uint32_t channels, uint64_t frames, const void* data) {
    if (!c || !data || format > 5 || channels == 0 || channels > 64 || frames == 0) return 0;
    auto r = std::make_shared<SampleResource>();
    r->fmt = (SampleFormat)format; r->channels = channels; r->frames = frames;
    const size_t n = (size_t)channels * frames;
    if (format <= 1) r->f32.assign((const float*)data, (const float*)data + n);
    else if (format == 2 || format == 4) r->i16.assign((const int16_t*)data, (const int16_t*)data + n);
    else r->u16.assign((const uint16_t*)data, (const uint16_t*)data + n);
    c->resources.push_back(std::move(r));
    return (uint32_t)c->resources.size();
}
static int sampler_set_sample_now(fw_ctx* c, fw_node_id node, uint32_t voice, uint32_t res, int stop_playback) {
    if (!c || res == 0 || res > c->resources.size()) return FW_SAMPLER_BAD_ARGS;
    return each_sampler(c, node, voice, [&](SamplerNode& n) { return n.set_sample(c->resources[res - 1], stop_playback != 0); });
}
int fwo_sampler_set_sample(fw_ctx* c, fw_node_id node, uint32_t voice, uint32_t res, int stop_playback) { return defer_or_run(c, 0, [=]() { return sampler_set_sample_now(c, node, voice, res, stop_playback); }); }
static int sampler_play_now(fw_ctx* c, fw_node_id node, uint32_t voice) { return each_sampler(c, node, voice, [](SamplerNode& n) { return n.play(); }); }
int fwo_sampler_play(fw_ctx* c, fw_node_id node, uint32_t voice) { return defer_or_run(c, 0, [=]() { return sampler_play_now(c, node, voice); }); }
static int sampler_pause_now(fw_ctx* c, fw_node_id node, uint32_t voice) { return each_sampler(c, node, voice, [](SamplerNode& n) { return n.pause(); }); }
int fwo_sampler_pause(fw_ctx* c, fw_node_id node, uint32_t voice) { return defer_or_run(c, 0, [=]() { return sampler_pause_now(c, node, voice); }); }
static int sampler_stop_now(fw_ctx* c, fw_node_id node, uint32_t voice) { return each_sampler(c, node, voice, [](SamplerNode& n) { return n.stop(); }); }
int fwo_sampler_stop(fw_ctx* c, fw_node_id node, uint32_t voice) { return defer_or_run(c, 0, [=]() { return sampler_stop_now(c, node, voice); }); }
static int sampler_set_playhead_now(fw_ctx* c, fw_node_id node, uint32_t voice, double secs) { return each_sampler(c, node, voice, [&](SamplerNode& n) { return n.set_playhead(secs); }); }
int fwo_sampler_set_playhead(fw_ctx* c, fw_node_id node, uint32_t voice, double secs) { return defer_or_run(c, 0, [=]() { return sampler_set_playhead_now(c, node, voice, secs); }); }
static int sampler_set_loop_range_now(fw_ctx* c, fw_node_id node, uint32_t voice, uint32_t mode, double s, double e) {
    if (mode > 2) return FW_SAMPLER_BAD_ARGS;
    if (mode == 2 && c && !c->voices.empty()) {
        const uint32_t sr = c->voices[0]->sample_rate();
        if (sr && !(secs_to_frame(s, sr) < secs_to_frame(e, sr))) return FW_SAMPLER_BAD_ARGS;
    }
    return each_sampler(c, node, voice, [&](SamplerNode& n) { return n.set_loop_range(mode, s, e); });
}
int fwo_sampler_set_loop_range(fw_ctx* c, fw_node_id node, uint32_t voice, uint32_t mode, double s, double e) { return defer_or_run(c, 0, [=]() { return sampler_set_loop_range_now(c, node, voice, mode, s, e); }); }
static int sampler_set_percent_volume_now(fw_ctx* c, fw_node_id node, uint32_t voice, float pct) { return each_sampler(c, node, voice, [&](SamplerNode& n) { n.set_percent_volume(pct); return 0; }); }
int fwo_sampler_set_percent_volume(fw_ctx* c, fw_node_id node, uint32_t voice, float pct) { return defer_or_run(c, 0, [=]() { return sampler_set_percent_volume_now(c, node, voice, pct); }); }
int fwo_sampler_is_playing(fw_ctx* c, fw_node_id node, uint32_t voice) {
    if (!c || voice >= c->voices.size()) return FW_SAMPLER_NOT_A_SAMPLER;
    auto* n = dynamic_cast<SamplerNode*>(c->voices[voice]->graph.node(nid(node)));
    return n ? (n->playing ? 1 : 0) : FW_SAMPLER_NOT_A_SAMPLER;
}

// ---- lifecycle -------------------------------------------------------------------------------
int fwo_ctx_activate(fw_ctx* c, uint32_t sr, uint32_t n_in, uint32_t n_out, uint32_t mbf, void* user_cx, fw_processor** out) {
    if (!c || !out || sr == 0 || mbf == 0 || n_in > 64 || n_out > 64) return -1;
    if (c->voices[0]->is_activated()) return 1;
    auto* p = new fw_processor();
    p->ctx = c; p->max_block_frames = mbf;
    for (auto& v : c->voices) p->procs.push_back(v->activate(sr, n_in, n_out, mbf, user_cx));
    c->live_processor = p;
    *out = p;
    return 0;
}
int fwo_ctx_is_activated(fw_ctx* c) { return c->voices[0]->is_activated(); }
int fwo_ctx_update(fw_ctx* c, fw_update_status* out) {
    UpdateStatus s0;
    for (size_t v = 0; v < c->voices.size(); ++v) { UpdateStatus s = c->voices[v]->update(); if (v == 0) s0 = s; }
    if (out) {
        std::memset(out, 0, sizeof(*out));
        out->kind = (int32_t)s0.kind; out->graph_error = (int32_t)s0.graph_error.code;
        out->error_node = pack(s0.graph_error.node.idx); out->error_port = s0.graph_error.port; out->returned_user_cx = s0.returned_user_cx;
    }
    if (s0.graph_error.code != CompileGraphError::Ok) c->last_error = s0.graph_error.message;
    return 0;
}
void* fwo_ctx_deactivate(fw_ctx* c, int stream_is_running) {
    void* cx = nullptr;
    for (size_t v = 0; v < c->voices.size(); ++v) { void* r = c->voices[v]->deactivate(false); if (v == 0) cx = r; }
    // The oracle's rings are single-threaded: a still-live processor cannot observe Stop from another
    // thread, so `stream_is_running` is honoured only as "the caller already freed the processor".
    (void)stream_is_running;
    return cx;
}

// ---- hot path --------------------------------------------------------------------------------
namespace {
// Balanced tree of 2-port SumNodes (sum.rs:69-81) over V leaves of n_out channels x bf frames, evaluated in a
// preallocated arena: nodes [0,V) are the voices' graph_out buffers, internal nodes are appended behind them.
struct BusArena {
    std::vector<float> buf; std::vector<SilenceMask> mask; size_t n_out = 0, mbf = 0;
    std::vector<const float*> in; std::vector<float*> out; std::vector<size_t> level, next;
    void reserve(size_t V, size_t n_out_, size_t mbf_) {
        if (buf.size() < 2 * V * n_out_ * mbf_ || n_out != n_out_ || mbf != mbf_) { buf.assign(2 * V * n_out_ * mbf_, 0.0f); mask.assign(2 * V, SilenceMask::none()); }
        n_out = n_out_; mbf = mbf_;
    }
    float* node(size_t i) { return buf.data() + i * n_out * mbf; }
    size_t reduce(size_t V, size_t bf) {  // returns the root node index
        SumNodeProcessor sum2(2);
        level.resize(V); for (size_t v = 0; v < V; ++v) level[v] = v;
        size_t next_free = V;
        while (level.size() > 1) {
            next.clear();
            for (size_t i = 0; i + 1 < level.size(); i += 2) {
                const size_t o = next_free++;
                in.clear(); out.clear(); SilenceMask in_mask = SilenceMask::none();
                for (size_t port = 0; port < 2; ++port) for (size_t ch = 0; ch < n_out; ++ch) {
                    const size_t s = level[i + port];
                    in.push_back(node(s) + ch * mbf);
                    if (mask[s].is_channel_silent(ch)) in_mask.set_channel(port * n_out + ch, true);
                }
                for (size_t ch = 0; ch < n_out; ++ch) out.push_back(node(o) + ch * mbf);
                SilenceMask om = SilenceMask::none();
                sum2.process(bf, in, out, ProcInfo{in_mask, &om, 0.0, 0, nullptr});
                mask[o] = om;
                next.push_back(o);
            }
            if (level.size() & 1) next.push_back(level.back());  // unpaired: carried up (1-port SumNode copy, sum.rs:58-65)
            level.swap(next);
        }
        return level[0];
    }
};
thread_local BusArena g_arena;
}  // namespace

int fwo_processor_process_planar(fw_processor* p, const float* input, float* output, uint32_t n_in, uint32_t n_out, uint64_t frames,
                                 double t, uint32_t status, uint64_t* out_mask) {
    if (!p) return FW_PROC_BAD_ARGS;
    size_t V = p->procs.size(); bool bus = p->ctx->cfg.master_bus != 0;
    size_t mbf = p->max_block_frames; int rc = FW_PROC_OK;
    if (out_mask) *out_mask = 0;
    const float* in_ptrs[64]; float* out_ptrs[64];
    BusArena& ar = g_arena;
    if (bus) ar.reserve(V, n_out, mbf);
    size_t done = 0;
    uint32_t block = 0;
    // frames == 0 still polls messages like processor.rs:76-89
    do {
        size_t bf = std::min<size_t>(frames - done, mbf);
        if (!p->ctx->deferred.empty()) run_deferred_at(p->ctx, block);  // stores stamped with this block (ctx_set_event_block)
        ++block;
        for (size_t v = 0; v < V; ++v) {
            for (uint32_t c = 0; c < n_in; ++c) in_ptrs[c] = input + ((size_t)v * n_in + c) * frames + done;
            uint64_t m = 0; ProcessorStatus st;
            if (bus) {
                for (uint32_t c = 0; c < n_out; ++c) out_ptrs[c] = ar.node(v) + (size_t)c * mbf;
                st = p->procs[v]->process_planar(in_ptrs, out_ptrs, n_in, n_out, bf, t, status, &m);
                ar.mask[v] = SilenceMask{m};
            } else {
                for (uint32_t c = 0; c < n_out; ++c) out_ptrs[c] = output + ((size_t)v * n_out + c) * frames + done;
                st = p->procs[v]->process_planar(in_ptrs, out_ptrs, n_in, n_out, bf, t, status, &m);
                if (out_mask && v == 0) *out_mask = m;
            }
            if (st == ProcessorStatus::DropProcessor) rc = FW_PROC_DROP_PROCESSOR;
        }
        if (rc != FW_PROC_OK) {  // processor.rs:71-74,150-155: zero-fill what was not produced
            size_t rows = bus ? n_out : V * n_out;
            for (size_t r = 0; r < rows; ++r) for (size_t i = done; i < frames; ++i) output[r * frames + i] = 0.0f;
            break;
        }
        if (bus && bf > 0) {
            size_t root = ar.reduce(V, bf);
            for (uint32_t c = 0; c < n_out; ++c) std::memcpy(output + (size_t)c * frames + done, ar.node(root) + (size_t)c * mbf, bf * sizeof(float));
            if (out_mask) *out_mask = ar.mask[root].bits;
        }
        done += bf;
    } while (done < frames);
    if (!p->ctx->deferred.empty()) rebase_deferred(p->ctx, frames ? (uint32_t)((frames + mbf - 1) / mbf) : 0u);
    return rc;
}

int fwo_processor_process_interleaved(fw_processor* p, const float* input, float* output, uint32_t n_in, uint32_t n_out, uint64_t frames,
                                      double t, uint32_t status) {
    if (!p) return FW_PROC_BAD_ARGS;
    size_t V = p->procs.size(); bool bus = p->ctx->cfg.master_bus != 0;
    if (!bus && p->ctx->deferred.empty()) {
        int rc = FW_PROC_OK;
        for (size_t v = 0; v < V; ++v) {
            ProcessorStatus st = p->procs[v]->process_interleaved(input + v * frames * n_in, frames * n_in, output + v * frames * n_out, frames * n_out,
                                                                  n_in, n_out, frames, t, status);
            if (st == ProcessorStatus::DropProcessor) rc = FW_PROC_DROP_PROCESSOR;
        }
        return rc;
    }
    if (!bus) {  // stores stamped with a block offset: one device callback per block, the store in between (what a host of the reference does)
        int rc = FW_PROC_OK;
        const size_t mbf = p->max_block_frames;
        uint32_t block = 0;
        for (size_t done = 0; done < frames; done += mbf, ++block) {
            const size_t bf = std::min<size_t>(frames - done, mbf);
            run_deferred_at(p->ctx, block);
            for (size_t v = 0; v < V; ++v) {
                ProcessorStatus st = p->procs[v]->process_interleaved(input + (v * frames + done) * n_in, bf * n_in, output + (v * frames + done) * n_out, bf * n_out,
                                                                      n_in, n_out, bf, t, status);
                if (st == ProcessorStatus::DropProcessor) rc = FW_PROC_DROP_PROCESSOR;
            }
        }
        rebase_deferred(p->ctx, (uint32_t)((frames + mbf - 1) / mbf));
        return rc;
    }
    // master bus: de-interleave per voice, run the planar path, interleave the bus with the root mask
    std::vector<float> pin((size_t)V * n_in * frames), pout((size_t)n_out * frames);
    for (size_t v = 0; v < V; ++v) for (size_t f = 0; f < frames; ++f) for (uint32_t c = 0; c < n_in; ++c)
        pin[((size_t)v * n_in + c) * frames + f] = input[((size_t)v * frames + f) * n_in + c];
    uint64_t m = 0;
    int rc = fwo_processor_process_planar(p, pin.data(), pout.data(), n_in, n_out, frames, t, status, &m);
    SilenceMask mask{m};
    std::vector<const float*> chans(n_out);
    for (uint32_t c = 0; c < n_out; ++c) chans[c] = pout.data() + (size_t)c * frames;
    if (n_out == 2) interleave_stereo(chans[0], chans[1], output, frames * 2, &mask);
    else interleave(chans, frames, output, frames * n_out, n_out, &mask);
    return rc;
}
int fwo_processor_process_planar_device(fw_processor*, const float*, float*, uint32_t, uint32_t, uint64_t, double, uint32_t) { return FW_PROC_DEVICE_ERROR; }
void fwo_processor_free(fw_processor* p) {
    if (!p) return;
    if (p->ctx && p->ctx->live_processor == p) p->ctx->live_processor = nullptr;
    p->procs.clear();  // Drop: each voice's processor posts Dropped{..} to its context
    delete p;
}

// ---- device plumbing: not available in the oracle ---------------------------------------------
int fwo_device_count(void) { return 0; }
const char* fwo_last_device_error(void) { return "oracle: no device"; }
void* fwo_dev_malloc(int, uint64_t) { return nullptr; }
void fwo_dev_free(int, void*) {}
void* fwo_host_alloc_pinned(uint64_t) { return nullptr; }
void fwo_host_free_pinned(void*) {}
int fwo_processor_h2d(fw_processor*, void*, const void*, uint64_t) { return -1; }
int fwo_processor_d2h(fw_processor*, void*, const void*, uint64_t) { return -1; }
int fwo_processor_sync(fw_processor*) { return 0; }
int fwo_processor_event_record(fw_processor*, int) { return -1; }
float fwo_processor_event_elapsed_ms(fw_processor*, int, int) { return -1.0f; }
uint64_t fwo_processor_kernel_launches(fw_processor*) { return 0; }
uint64_t fwo_processor_graph_replays(fw_processor*) { return 0; }
// Isomorphic-voice detection and the parameter-table getter are host logic of the product; the oracle runs a flat graph as it is
// (that is what the detection's equivalence tests compare against) and only answers "not here".
int fwo_graph_detect_voices(fw_ctx* c, fw_voice_template* out) { if (out) *out = fw_voice_template{}; if (c) c->last_error = "the oracle does not batch: it runs the flat graph"; return -1; }
uint32_t fwo_graph_voice_nodes(fw_ctx*, uint32_t, fw_node_id*, uint32_t) { return 0; }
fw_ctx* fwo_ctx_new_batched(fw_ctx* c, int32_t, uint32_t, fw_node_id*, uint32_t) { if (c) c->last_error = "the oracle does not batch: it runs the flat graph"; return nullptr; }
uint32_t fwo_node_read_params(fw_ctx*, fw_node_id, uint32_t, float*, uint32_t) { return 0; }
int fwo_processor_l2_flush(fw_processor*) { return -1; }
int fwo_processor_profile(fw_processor*, int) { return -1; }
int fwo_processor_profile_read(fw_processor*, double*, uint64_t*) { return -1; }
int fwo_comm_unique_id(uint8_t*) { return -1; }
int fwo_processor_comm_init(fw_processor*, int, int, const uint8_t*) { return -1; }
int fwo_processor_comm_allgather(fw_processor*, const void*, void*, uint64_t) { return -1; }

// ---- oracle-only extras for known-answer tests and the CPU baseline ---------------------------
// ParamSmoother driven directly (smoother.rs:93-205): returns the status after the last block.
FW_EXPORT int fwo_smoother_run(float initial, uint32_t sample_rate, uint32_t max_block_frames, const float* targets, const uint32_t* frames_per_block,
                               uint32_t n_blocks, float* out_curves /* [n_blocks][max_block_frames] */, uint32_t* out_len, uint32_t* out_status,
                               float* out_ab) {
    ParamSmoother s(initial, sample_rate, max_block_frames);
    if (out_ab) { out_ab[0] = s.a; out_ab[1] = s.b; }
    for (uint32_t k = 0; k < n_blocks; ++k) {
        SmoothedOutput o = s.set_and_process(targets[k], frames_per_block[k]);
        if (out_len) out_len[k] = (uint32_t)o.len;
        if (out_status) out_status[k] = (uint32_t)o.status;
        if (out_curves) std::memcpy(out_curves + (size_t)k * max_block_frames, o.values, std::min<size_t>(o.len, max_block_frames) * sizeof(float));
    }
    return (int)s.status;
}
FW_EXPORT float fwo_percent_volume_to_raw_gain(float p) { return percent_volume_to_raw_gain(p); }
FW_EXPORT float fwo_db_to_gain_clamped_neg_100_db(float db) { return db_to_gain_clamped_neg_100_db(db); }
FW_EXPORT float fwo_bf16_round(float x) { return bf16_round(x); }
FW_EXPORT void fwo_pan_to_gains(float pan, float* gl, float* gr) { pan_to_gains(pan, gl, gr); }
FW_EXPORT uint64_t fwo_silence_mask_new_all_silent(uint32_t n) { return SilenceMask::new_all_silent(n).bits; }
FW_EXPORT int fwo_silence_mask_query(uint64_t bits, uint32_t n, int which) {
    SilenceMask m{bits};
    return which == 0 ? m.any_channel_silent(n) : which == 1 ? m.all_channels_silent(n) : m.is_channel_silent(n);
}
FW_EXPORT uint64_t fwo_deinterleave(float* planar /* [n_ch_out][frames], pre-filled (stale) */, uint32_t n_ch_out, uint32_t frames,
                                    const float* interleaved, uint32_t n_interleaved, int calc_mask) {
    std::vector<float*> ch(n_ch_out);
    for (uint32_t c = 0; c < n_ch_out; ++c) ch[c] = planar + (size_t)c * frames;
    return deinterleave(ch, frames, interleaved, (size_t)frames * n_interleaved, n_interleaved, calc_mask != 0).bits;
}
FW_EXPORT void fwo_interleave(const float* planar, uint32_t n_ch_in, uint32_t frames, float* interleaved, uint32_t n_interleaved,
                              int use_mask, uint64_t mask_bits, int stereo_fast_path) {
    std::vector<const float*> ch(n_ch_in);
    for (uint32_t c = 0; c < n_ch_in; ++c) ch[c] = planar + (size_t)c * frames;
    SilenceMask m{mask_bits};
    if (stereo_fast_path) interleave_stereo(ch[0], ch[1], interleaved, (size_t)frames * 2, use_mask ? &m : nullptr);
    else interleave(ch, frames, interleaved, (size_t)frames * n_interleaved, n_interleaved, use_mask ? &m : nullptr);
}

// The pull-style stream backend is product plumbing (a host thread around process_interleaved): not restated here.
struct fw_stream;
fw_stream* fwo_stream_open(fw_processor*, uint32_t, uint32_t, uint32_t, uint32_t) { return nullptr; }
int64_t fwo_stream_pull(fw_stream*, float*, uint64_t, uint32_t*, double*) { return -1; }
uint64_t fwo_stream_frames_ready(fw_stream*) { return 0; }
void fwo_stream_close(fw_stream*) {}

}  // extern "C"
