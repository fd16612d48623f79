// fw_oracle.hpp — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
//
// A C++ restatement of the per-block audio-graph DSP path of BillyDM/firewheel
// @ 2dfa7ea (Rust, /root/reference). Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may link or call this.
//
// Build flags for fidelity (oracle/Makefile): -O2 -ffp-contract=off -fno-fast-math.
// Rust/LLVM never contracts a*b+c into an FMA, so neither may this file.
//
// PARITY STATUS
//   * Pinned by restating the Rust source line by line (single f32 ops in a
//     fixed order => bit-exactness is well defined): SilenceMask, ParamSmoother,
//     util::{deinterleave,interleave,interleave_stereo,clear_all_outputs},
//     percent_volume_to_raw_gain, Volume/Sum/MonoToStereo/StereoToMono/HardClip
//     processors, AudioGraph, the compiler (Kahn sort + buffer allocator),
//     CompiledSchedule::{prepare_graph_inputs,process,read_graph_outputs},
//     FirewheelProcessor::process_interleaved, FirewheelGraphCtx, SamplerNode /
//     SamplerProcessor and the SampleResource implementations.
//     The reference's only tests are 5 *structural* schedule tests
//     (schedule.rs:407,451,539,662,685); all five are re-run against this file
//     in tests/test_structural.py. The reference holds NO numeric golden
//     vectors, so numeric parity is anchored on (i) the restated formulas plus
//     hand-derived known answers (tests/test_oracle_kat.py, tests/test_sampler_oracle.py),
//     (ii) a second, independent restatement of the executor, the smoother, the
//     pinned nodes and the sampler in pure Python (tests/pyref.py) that must agree
//     with this file bit for bit on random graphs (tests/test_oracle_vs_pyref.py),
//     (iii) committed golden vectors (tests/golden/).
//   * PARITY UNPINNED (no reference code or tests exist; the spec below IS the
//     definition): PanNode, BiquadNode, SvfNode, DelayNode, ResamplerNode,
//     ConvReverbNode. They follow SURVEY.md §8 a10-a14 and are cross-checked
//     against scipy / closed forms (tests/test_oracle_kat.py,
//     tests/test_svf_resampler_oracle.py).
//   * thunderdome 0.6.1 (generational arena; absent from /root/reference, a
//     Cargo dependency: crates/firewheel-graph/Cargo.toml:21) is restated from
//     its published algorithm: LIFO free list, generation bumped on slot reuse,
//     iteration in ascending slot order.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <optional>
#include <set>
#include <string>
#include <utility>
#include <vector>

namespace fwo {

// ---------------------------------------------------------------------------
// thunderdome::Arena (0.6.1) — restated
// ---------------------------------------------------------------------------
struct Index {
    uint32_t slot = UINT32_MAX;
    uint32_t generation = UINT32_MAX;  // Index::DANGLING
    bool operator==(const Index& o) const { return slot == o.slot && generation == o.generation; }
    bool operator!=(const Index& o) const { return !(*this == o); }
    bool operator<(const Index& o) const {  // derive(Ord): generation first, then slot
        return generation != o.generation ? generation < o.generation : slot < o.slot;
    }
};

template <class T>
class Arena {
    struct Entry {
        bool occupied = false;
        uint32_t generation = 0;     // for Empty: generation of the last occupant
        uint32_t next_free = UINT32_MAX;
        std::optional<T> value;
    };
    std::vector<Entry> storage_;
    uint32_t len_ = 0;
    uint32_t first_free_ = UINT32_MAX;
    size_t cap_hint_ = 0;

  public:
    Arena() = default;
    explicit Arena(size_t cap) : cap_hint_(cap) { storage_.reserve(cap); }
    size_t capacity() const { return std::max(storage_.capacity(), cap_hint_); }
    size_t len() const { return len_; }
    size_t num_slots() const { return storage_.size(); }

    Index insert(T v) {
        len_ += 1;
        if (first_free_ != UINT32_MAX) {
            uint32_t slot = first_free_;
            Entry& e = storage_[slot];
            first_free_ = e.next_free;
            e.generation = e.generation + 1;  // Generation::next()
            e.occupied = true;
            e.value.emplace(std::move(v));
            return Index{slot, e.generation};
        }
        Entry e;
        e.occupied = true;
        e.generation = 1;  // Generation::first()
        e.value.emplace(std::move(v));
        storage_.push_back(std::move(e));
        return Index{(uint32_t)(storage_.size() - 1), 1};
    }

    // Insert at a specific index (slot + generation), returning the previous occupant.
    std::optional<T> insert_at(Index idx, T v) {
        while (storage_.size() <= idx.slot) {  // pad with empty entries, chained on the free list
            Entry e;
            e.occupied = false;
            e.generation = 0;
            e.next_free = first_free_;
            storage_.push_back(std::move(e));
            first_free_ = (uint32_t)(storage_.size() - 1);
        }
        Entry& e = storage_[idx.slot];
        std::optional<T> old;
        if (e.occupied) {
            old = std::move(e.value);
        } else {
            // unlink idx.slot from the free list
            uint32_t* link = &first_free_;
            while (*link != UINT32_MAX) {
                if (*link == idx.slot) { *link = storage_[idx.slot].next_free; break; }
                link = &storage_[*link].next_free;
            }
            len_ += 1;
        }
        e.occupied = true;
        e.generation = idx.generation;
        e.value.emplace(std::move(v));
        return old;
    }

    std::optional<T> remove(Index idx) {
        if (idx.slot >= storage_.size()) return std::nullopt;
        Entry& e = storage_[idx.slot];
        if (!e.occupied || e.generation != idx.generation) return std::nullopt;
        std::optional<T> out = std::move(e.value);
        e.value.reset();
        e.occupied = false;
        e.next_free = first_free_;
        first_free_ = idx.slot;
        len_ -= 1;
        return out;
    }

    bool contains(Index idx) const { return get(idx) != nullptr; }
    T* get(Index idx) {
        if (idx.slot >= storage_.size()) return nullptr;
        Entry& e = storage_[idx.slot];
        return (e.occupied && e.generation == idx.generation) ? &*e.value : nullptr;
    }
    const T* get(Index idx) const { return const_cast<Arena*>(this)->get(idx); }
    T& at(Index idx) { T* p = get(idx); assert(p && "arena index"); return *p; }
    // get_by_slot: (Index, &T)
    std::pair<Index, T*> get_by_slot(uint32_t slot) {
        if (slot >= storage_.size() || !storage_[slot].occupied) return {Index{}, nullptr};
        return {Index{slot, storage_[slot].generation}, &*storage_[slot].value};
    }
    // iteration in ascending slot order
    template <class F> void for_each(F&& f) {
        for (uint32_t s = 0; s < storage_.size(); ++s)
            if (storage_[s].occupied) f(Index{s, storage_[s].generation}, *storage_[s].value);
    }
    template <class F> void for_each(F&& f) const {
        for (uint32_t s = 0; s < storage_.size(); ++s)
            if (storage_[s].occupied) f(Index{s, storage_[s].generation}, *storage_[s].value);
    }
    // drain(): yields every element in slot order and empties the arena
    std::vector<std::pair<Index, T>> drain() {
        std::vector<std::pair<Index, T>> out;
        for (uint32_t s = 0; s < storage_.size(); ++s)
            if (storage_[s].occupied) out.emplace_back(Index{s, storage_[s].generation}, std::move(*storage_[s].value));
        storage_.clear();
        len_ = 0;
        first_free_ = UINT32_MAX;
        return out;
    }
};

// ---------------------------------------------------------------------------
// firewheel-core/src/silence_mask.rs:7-74
// ---------------------------------------------------------------------------
struct SilenceMask {
    uint64_t bits = 0;
    static SilenceMask none() { return SilenceMask{0}; }
    static SilenceMask new_all_silent(size_t n) {  // :23-29
        return n >= 64 ? SilenceMask{UINT64_MAX} : SilenceMask{(uint64_t(1) << n) - 1};
    }
    bool is_channel_silent(size_t i) const { return (bits & (uint64_t(1) << i)) != 0; }  // :35
    bool any_channel_silent(size_t n) const {                                              // :43
        return n >= 64 ? bits != 0 : (bits & ((uint64_t(1) << n) - 1)) != 0;
    }
    bool all_channels_silent(size_t n) const {  // :55
        if (n >= 64) return bits == UINT64_MAX;
        uint64_t m = (uint64_t(1) << n) - 1;
        return (bits & m) == m;
    }
    void set_channel(size_t i, bool silent) {  // :67
        if (silent) bits |= uint64_t(1) << i; else bits &= ~(uint64_t(1) << i);
    }
};

// ---------------------------------------------------------------------------
// firewheel-core/src/param/range.rs:32-35, util.rs:7-41
// ---------------------------------------------------------------------------
inline float percent_volume_to_raw_gain(float percent_volume) {
    float n = std::fmax(percent_volume, 0.0f) * (1.0f / 100.0f);
    return n * n;
}
inline float db_to_gain(float db) { return std::pow(10.0f, 0.05f * db); }              // util.rs:7
inline float gain_to_db(float amp) { return 20.0f * std::log10(amp); }                  // util.rs:13
inline float db_to_gain_clamped_neg_100_db(float db) { return db <= -100.0f ? 0.0f : db_to_gain(db); }
inline float gain_to_db_clamped_neg_100_db(float amp) { return amp <= 0.00001f ? -100.0f : gain_to_db(amp); }

// ---------------------------------------------------------------------------
// util.rs:44-175 — (de)interleave, clear_all_outputs
// ---------------------------------------------------------------------------
// channels: each pointer addresses `frames` samples.
inline SilenceMask deinterleave(std::vector<float*>& channels, size_t frames, const float* interleaved,
                                size_t interleaved_len, size_t num_interleaved_channels,
                                bool calculate_silence_mask) {  // util.rs:44-87
    SilenceMask mask = SilenceMask::none();
    size_t i = 0, next = 0;
    for (size_t k = 0; k < num_interleaved_channels; ++k) {
        if (next >= channels.size()) return mask;
        float* ch = channels[next++];
        if (calculate_silence_mask && i < 64) {
            // Q4: the scan reads the DESTINATION buffer's stale contents (util.rs:58-62)
            bool all_zero = true;
            for (size_t f = 0; f < frames; ++f) if (ch[f] != 0.0f) { all_zero = false; break; }
            if (all_zero) mask.set_channel(i, true);
        }
        size_t f = 0;
        for (size_t src = i; src < interleaved_len && f < frames; src += num_interleaved_channels, ++f) ch[f] = interleaved[src];
        i += 1;
    }
    while (next < channels.size()) {
        float* ch = channels[next++];
        for (size_t f = 0; f < frames; ++f) ch[f] = 0.0f;
        if (calculate_silence_mask && i < 64) mask.set_channel(i, true);
        i += 1;
    }
    return mask;
}

inline void interleave(const std::vector<const float*>& channels, size_t frames, float* interleaved,
                       size_t interleaved_len, size_t num_interleaved_channels,
                       const SilenceMask* silence_mask) {  // util.rs:90-120
    for (size_t k = 0; k < interleaved_len; ++k) interleaved[k] = 0.0f;
    for (size_t ch_i = 0; ch_i < num_interleaved_channels; ++ch_i) {
        if (ch_i >= channels.size()) return;
        const float* ch = channels[ch_i];
        if (silence_mask && ch_i < 64 && silence_mask->is_channel_silent(ch_i)) continue;
        size_t f = 0;
        for (size_t dst = ch_i; dst < interleaved_len && f < frames; dst += num_interleaved_channels, ++f) interleaved[dst] = ch[f];
    }
}

inline void interleave_stereo(const float* in_l, const float* in_r, float* interleaved, size_t interleaved_len,
                              const SilenceMask* silence_mask) {  // util.rs:123-147
    if (silence_mask && silence_mask->all_channels_silent(2)) {
        for (size_t k = 0; k < interleaved_len; ++k) interleaved[k] = 0.0f;
        return;
    }
    size_t frames = interleaved_len / 2;
    for (size_t f = 0; f < frames; ++f) { interleaved[2 * f] = in_l[f]; interleaved[2 * f + 1] = in_r[f]; }
}

inline void deinterleave_stereo(float* out_l, float* out_r, const float* interleaved, size_t interleaved_len) {  // :150
    size_t frames = interleaved_len / 2;
    for (size_t f = 0; f < frames; ++f) { out_l[f] = interleaved[2 * f]; out_r[f] = interleaved[2 * f + 1]; }
}

inline void clear_all_outputs(size_t frames, const std::vector<float*>& outputs, SilenceMask* out_mask) {  // :165-175
    for (float* o : outputs) for (size_t i = 0; i < frames; ++i) o[i] = 0.0f;
    *out_mask = SilenceMask::new_all_silent(outputs.size());
}

// ---------------------------------------------------------------------------
// firewheel-core/src/param/smoother.rs:7-239
// ---------------------------------------------------------------------------
struct SmootherConfig { float smooth_secs = 10.0f / 1000.0f; float settle_epsilon = 0.00001f; };  // :18-25
enum class SmootherStatus : uint32_t { Inactive = 0, Active = 1, Deactivating = 2 };

struct SmoothedOutput {
    const float* values; size_t len; SmootherStatus status;
    bool is_smoothing() const { return status != SmootherStatus::Inactive; }  // :54-56
    float operator[](size_t i) const { return values[i]; }
};

class ParamSmoother {
  public:
    std::vector<float> output; float input; SmootherStatus status; float a, b, last_output, settle_epsilon;
    ParamSmoother(float val, uint32_t sample_rate, size_t max_block_frames, SmootherConfig cfg = {}) {  // :93-112
        b = std::exp(-1.0f / (cfg.smooth_secs * (float)sample_rate));
        a = 1.0f - b;
        status = SmootherStatus::Inactive; input = val; output.assign(max_block_frames, val);
        last_output = val; settle_epsilon = cfg.settle_epsilon;
    }
    bool is_active() const { return status != SmootherStatus::Inactive; }
    void reset(float val) {  // :115-129
        if (is_active()) {
            status = SmootherStatus::Inactive; input = val; last_output = val;
            std::fill(output.begin(), output.end(), val);
        } else if (input != val) {
            input = val; last_output = val; std::fill(output.begin(), output.end(), val);
        }
    }
    void set(float val) { if (input == val) return; input = val; status = SmootherStatus::Active; }  // :133-140
    SmoothedOutput process(size_t frames) {  // :159-194
        frames = std::min(frames, output.size());
        if (status != SmootherStatus::Active || frames == 0 || output.empty())
            return SmoothedOutput{output.data(), output.size(), status};  // Q1: FULL-length buffer
        float in = input * a;
        output[0] = in + (last_output * b);
        for (size_t i = 1; i < frames; ++i) output[i] = in + (output[i - 1] * b);
        last_output = output[frames - 1];
        if (status == SmootherStatus::Active) {
            if (std::fabs(input - output[0]) < settle_epsilon) {  // Q3: tests output[0], resets whole curve
                reset(input);
                status = SmootherStatus::Deactivating;  // Q2: Deactivating -> Inactive is unreachable
            }
        } else if (status == SmootherStatus::Deactivating) {
            status = SmootherStatus::Inactive;
        }
        return SmoothedOutput{output.data(), frames, status};
    }
    SmoothedOutput set_and_process(float val, size_t frames) { set(val); return process(frames); }  // :202-205
};

// ---------------------------------------------------------------------------
// firewheel-core/src/node.rs:6-132 — the plugin API
// ---------------------------------------------------------------------------
struct AudioNodeInfo {
    uint32_t num_min_supported_inputs = 0, num_max_supported_inputs = 0;
    uint32_t num_min_supported_outputs = 0, num_max_supported_outputs = 0;
    bool updates = false;
};
enum StreamStatus : uint32_t { INPUT_OVERFLOW = 1, OUTPUT_UNDERFLOW = 2 };
struct ProcInfo {
    SilenceMask in_silence_mask; SilenceMask* out_silence_mask; double stream_time_secs; uint32_t stream_status; void* cx;
};
struct AudioNodeProcessor {
    virtual ~AudioNodeProcessor() = default;
    virtual void process(size_t frames, const std::vector<const float*>& inputs, const std::vector<float*>& outputs, ProcInfo info) = 0;
};
struct AudioNode {
    virtual ~AudioNode() = default;
    virtual const char* debug_name() const = 0;
    virtual AudioNodeInfo info() const = 0;
    // Ok => processor, Err => nullptr + message
    virtual std::unique_ptr<AudioNodeProcessor> activate(uint32_t sample_rate, size_t max_block_frames, size_t num_inputs,
                                                          size_t num_outputs, std::string* err) = 0;
    virtual void deactivate(std::unique_ptr<AudioNodeProcessor>) {}
    virtual void update() {}
};

// ---------------------------------------------------------------------------
// basic_nodes/dummy.rs
// ---------------------------------------------------------------------------
struct DummyProcessor : AudioNodeProcessor {
    void process(size_t, const std::vector<const float*>&, const std::vector<float*>&, ProcInfo) override {}
};
struct DummyAudioNode : AudioNode {
    const char* debug_name() const override { return "dummy"; }
    AudioNodeInfo info() const override { AudioNodeInfo i; i.num_max_supported_inputs = 64; i.num_max_supported_outputs = 64; return i; }
    std::unique_ptr<AudioNodeProcessor> activate(uint32_t, size_t, size_t, size_t, std::string*) override {
        return std::make_unique<DummyProcessor>();
    }
};

// ---------------------------------------------------------------------------
// basic_nodes/volume.rs:8-151  ("GainNode")
// ---------------------------------------------------------------------------
struct VolumeProcessor : AudioNodeProcessor {
    std::shared_ptr<float> raw_gain; ParamSmoother gain_smoother;
    VolumeProcessor(std::shared_ptr<float> g, uint32_t sr, size_t mbf) : raw_gain(std::move(g)), gain_smoother(*raw_gain, sr, mbf) {}
    void process(size_t frames, const std::vector<const float*>& inputs, const std::vector<float*>& outputs, ProcInfo pi) override {
        float g = *raw_gain;  // :92
        if (pi.in_silence_mask.all_channels_silent(inputs.size())) {  // :94-100
            gain_smoother.reset(g);
            clear_all_outputs(frames, outputs, pi.out_silence_mask);
            return;
        }
        SmoothedOutput gain = gain_smoother.set_and_process(g, frames);  // :102
        if (!gain.is_smoothing() && gain.values[0] < 0.00001f) {  // :104-108
            clear_all_outputs(frames, outputs, pi.out_silence_mask);
            return;
        }
        *pi.out_silence_mask = pi.in_silence_mask;  // :110
        assert(frames <= gain.len);
        if (inputs.size() == 2 && outputs.size() == 2) {  // :116-129 (Q7: both channels regardless of flags)
            for (size_t i = 0; i < frames; ++i) {
                outputs[0][i] = inputs[0][i] * gain[i];
                outputs[1][i] = inputs[1][i] * gain[i];
            }
            return;
        }
        size_t n = std::min(outputs.size(), inputs.size());
        for (size_t c = 0; c < n; ++c) {  // :131-143
            if (pi.in_silence_mask.is_channel_silent(c)) { for (size_t i = 0; i < frames; ++i) outputs[c][i] = 0.0f; continue; }
            for (size_t i = 0; i < frames; ++i) outputs[c][i] = inputs[c][i] * gain[i];
        }
    }
};
struct VolumeNode : AudioNode {
    std::shared_ptr<float> raw_gain; float percent_volume;
    explicit VolumeNode(float percent) {  // :16-24
        percent = std::fmax(percent, 0.0f);
        raw_gain = std::make_shared<float>(percent_volume_to_raw_gain(percent));
        percent_volume = percent;
    }
    void set_percent_volume(float p) { *raw_gain = percent_volume_to_raw_gain(p); percent_volume = std::fmax(p, 0.0f); }  // :28-34
    const char* debug_name() const override { return "volume"; }
    AudioNodeInfo info() const override { return AudioNodeInfo{1, 64, 1, 64, false}; }
    std::unique_ptr<AudioNodeProcessor> activate(uint32_t sr, size_t mbf, size_t ni, size_t no, std::string* err) override {
        if (ni != no) {  // :63-65
            if (err) *err = "The number of inputs on a VolumeNode node must equal the number of outputs. Got num_inputs: " +
                            std::to_string(ni) + ", num_outputs: " + std::to_string(no);
            return nullptr;
        }
        return std::make_unique<VolumeProcessor>(raw_gain, sr, mbf);
    }
};

// ---------------------------------------------------------------------------
// basic_nodes/sum.rs:3-142
// ---------------------------------------------------------------------------
struct SumNodeProcessor : AudioNodeProcessor {
    size_t num_in_ports;
    explicit SumNodeProcessor(size_t p) : num_in_ports(p) {}
    void process(size_t frames, const std::vector<const float*>& inputs, const std::vector<float*>& outputs, ProcInfo pi) override {
        size_t num_inputs = inputs.size(), num_outputs = outputs.size();
        if (pi.in_silence_mask.all_channels_silent(inputs.size())) {  // :52-56
            clear_all_outputs(frames, outputs, pi.out_silence_mask);
            return;
        }
        if (num_inputs == num_outputs) {  // :58-65 copy
            for (size_t c = 0; c < num_outputs; ++c) std::memcpy(outputs[c], inputs[c], frames * sizeof(float));
            *pi.out_silence_mask = pi.in_silence_mask;
            return;
        }
        switch (num_in_ports) {
            case 2:  // :70-81
                for (size_t c = 0; c < num_outputs; ++c) {
                    const float *in1 = inputs[c], *in2 = inputs[num_outputs + c]; float* out = outputs[c];
                    for (size_t i = 0; i < frames; ++i) out[i] = in1[i] + in2[i];
                }
                break;
            case 3:  // :82-94
                for (size_t c = 0; c < num_outputs; ++c) {
                    const float *in1 = inputs[c], *in2 = inputs[num_outputs + c], *in3 = inputs[num_outputs * 2 + c]; float* out = outputs[c];
                    for (size_t i = 0; i < frames; ++i) out[i] = in1[i] + in2[i] + in3[i];
                }
                break;
            case 4:  // :95-110
                for (size_t c = 0; c < num_outputs; ++c) {
                    const float *in1 = inputs[c], *in2 = inputs[num_outputs + c], *in3 = inputs[num_outputs * 2 + c],
                                *in4 = inputs[num_outputs * 3 + c]; float* out = outputs[c];
                    for (size_t i = 0; i < frames; ++i) out[i] = in1[i] + in2[i] + in3[i] + in4[i];
                }
                break;
            default: {  // :111-133
                size_t n = num_in_ports;
                for (size_t c = 0; c < num_outputs; ++c) {
                    float* out = outputs[c];
                    std::memcpy(out, inputs[c], frames * sizeof(float));
                    for (size_t p = 1; p < n; ++p) {
                        size_t in_ch = num_outputs * p + c;
                        if (pi.in_silence_mask.is_channel_silent(in_ch)) continue;
                        const float* in = inputs[in_ch];
                        for (size_t i = 0; i < frames; ++i) out[i] += in[i];
                    }
                }
            }
        }
    }
};
struct SumNode : AudioNode {
    const char* debug_name() const override { return "sum"; }
    AudioNodeInfo info() const override { return AudioNodeInfo{1, 64, 1, 64, false}; }
    std::unique_ptr<AudioNodeProcessor> activate(uint32_t, size_t, size_t ni, size_t no, std::string* err) override {
        if (no == 0 || ni % no != 0) {  // :27-29 (num_outputs == 0 would be a Rust divide-by-zero panic)
            if (err) *err = "The number of inputs on a SumNode must be a multiple of the number of outputs. Got num_inputs: " +
                            std::to_string(ni) + ", num_outputs: " + std::to_string(no);
            return nullptr;
        }
        return std::make_unique<SumNodeProcessor>(ni / no);
    }
};

// ---------------------------------------------------------------------------
// basic_nodes/mono_to_stereo.rs:34-49, stereo_to_mono.rs:34-55, hard_clip.rs:52-94
// ---------------------------------------------------------------------------
struct MonoToStereoProcessor : AudioNodeProcessor {
    void process(size_t frames, const std::vector<const float*>& inputs, const std::vector<float*>& outputs, ProcInfo pi) override {
        if (pi.in_silence_mask.is_channel_silent(0)) { clear_all_outputs(frames, outputs, pi.out_silence_mask); return; }
        std::memcpy(outputs[0], inputs[0], frames * sizeof(float));
        std::memcpy(outputs[1], inputs[0], frames * sizeof(float));
    }
};
struct MonoToStereoNode : AudioNode {
    const char* debug_name() const override { return "mono_to_stereo"; }
    AudioNodeInfo info() const override { return AudioNodeInfo{1, 1, 2, 2, false}; }
    std::unique_ptr<AudioNodeProcessor> activate(uint32_t, size_t, size_t, size_t, std::string*) override {
        return std::make_unique<MonoToStereoProcessor>();
    }
};
struct StereoToMonoProcessor : AudioNodeProcessor {
    void process(size_t frames, const std::vector<const float*>& inputs, const std::vector<float*>& outputs, ProcInfo pi) override {
        if (pi.in_silence_mask.all_channels_silent(2) || inputs.size() < 2 || outputs.empty()) {
            clear_all_outputs(frames, outputs, pi.out_silence_mask);
            return;
        }
        // zip over the slices: the schedule hands exactly `frames`-long slices (schedule.rs:347-379)
        for (size_t i = 0; i < frames; ++i) outputs[0][i] = (inputs[0][i] + inputs[1][i]) * 0.5f;
    }
};
struct StereoToMonoNode : AudioNode {
    const char* debug_name() const override { return "stereo_to_mono"; }
    AudioNodeInfo info() const override { return AudioNodeInfo{2, 2, 1, 1, false}; }
    std::unique_ptr<AudioNodeProcessor> activate(uint32_t, size_t, size_t, size_t, std::string*) override {
        return std::make_unique<StereoToMonoProcessor>();
    }
};
struct HardClipProcessor : AudioNodeProcessor {
    float threshold_gain;
    explicit HardClipProcessor(float t) : threshold_gain(t) {}
    void process(size_t frames, const std::vector<const float*>& inputs, const std::vector<float*>& outputs, ProcInfo pi) override {
        float t = threshold_gain;
        if (inputs.size() == 2 && outputs.size() == 2 && !pi.in_silence_mask.any_channel_silent(2)) {  // :60-80 (Q7: mask unwritten)
            for (size_t i = 0; i < frames; ++i) {
                outputs[0][i] = std::fmax(std::fmin(inputs[0][i], t), -t);
                outputs[1][i] = std::fmax(std::fmin(inputs[1][i], t), -t);
            }
            return;
        }
        size_t n = std::min(outputs.size(), inputs.size());
        for (size_t c = 0; c < n; ++c) {
            if (pi.in_silence_mask.is_channel_silent(c)) { for (size_t i = 0; i < frames; ++i) outputs[c][i] = 0.0f; continue; }
            for (size_t i = 0; i < frames; ++i) outputs[c][i] = std::fmax(std::fmin(inputs[c][i], t), -t);
        }
        *pi.out_silence_mask = pi.in_silence_mask;
    }
};
struct HardClipNode : AudioNode {
    float threshold_gain;
    explicit HardClipNode(float threshold_db) : threshold_gain(db_to_gain_clamped_neg_100_db(threshold_db)) {}
    const char* debug_name() const override { return "hard_clip"; }
    AudioNodeInfo info() const override { return AudioNodeInfo{1, 64, 1, 64, false}; }
    std::unique_ptr<AudioNodeProcessor> activate(uint32_t, size_t, size_t ni, size_t no, std::string* err) override {
        if (ni != no) {
            if (err) *err = "The number of inputs on a HardClip node must equal the number of outputs. Got num_inputs: " +
                            std::to_string(ni) + ", num_outputs: " + std::to_string(no);
            return nullptr;
        }
        return std::make_unique<HardClipProcessor>(threshold_gain);
    }
};

// ===========================================================================
// OUR-SPEC NODES — PARITY UNPINNED (absent from the reference: README.md:14,18,
// DESIGN_DOC.md:13-20). Spec = SURVEY.md §8 a10-a14. Written in the reference's
// node idiom so that they read like siblings of VolumeNode.
// ===========================================================================

// a10 — stereo pan, equal power. Gains evaluated on the host in f64 -> f32.
inline void pan_to_gains(float pan, float* gl, float* gr) {
    double p = std::fmin(std::fmax((double)pan, -1.0), 1.0);
    double theta = (p + 1.0) * (M_PI / 4.0);
    *gl = (float)std::cos(theta);
    *gr = (float)std::sin(theta);
}
struct PanParams { float gl, gr; };
struct PanProcessor : AudioNodeProcessor {
    std::shared_ptr<PanParams> params; ParamSmoother sm_l, sm_r;
    PanProcessor(std::shared_ptr<PanParams> p, uint32_t sr, size_t mbf) : params(std::move(p)), sm_l(params->gl, sr, mbf), sm_r(params->gr, sr, mbf) {}
    void process(size_t frames, const std::vector<const float*>& inputs, const std::vector<float*>& outputs, ProcInfo pi) override {
        float gl = params->gl, gr = params->gr;
        if (pi.in_silence_mask.all_channels_silent(inputs.size())) {
            sm_l.reset(gl); sm_r.reset(gr);
            clear_all_outputs(frames, outputs, pi.out_silence_mask);
            return;
        }
        SmoothedOutput cl = sm_l.set_and_process(gl, frames);
        SmoothedOutput cr = sm_r.set_and_process(gr, frames);
        *pi.out_silence_mask = pi.in_silence_mask;
        for (size_t i = 0; i < frames; ++i) {
            outputs[0][i] = inputs[0][i] * cl[i];
            outputs[1][i] = inputs[1][i] * cr[i];
        }
    }
};
struct PanNode : AudioNode {
    std::shared_ptr<PanParams> params; float pan;
    explicit PanNode(float p) : params(std::make_shared<PanParams>()), pan(p) { pan_to_gains(p, &params->gl, &params->gr); }
    void set_pan(float p) { pan = p; pan_to_gains(p, &params->gl, &params->gr); }
    void set_gains(float gl, float gr) { params->gl = gl; params->gr = gr; }
    const char* debug_name() const override { return "pan"; }
    AudioNodeInfo info() const override { return AudioNodeInfo{2, 2, 2, 2, false}; }
    std::unique_ptr<AudioNodeProcessor> activate(uint32_t sr, size_t mbf, size_t ni, size_t no, std::string* err) override {
        if (ni != 2 || no != 2) { if (err) *err = "A PanNode must have 2 inputs and 2 outputs. Got num_inputs: " + std::to_string(ni) + ", num_outputs: " + std::to_string(no); return nullptr; }
        return std::make_unique<PanProcessor>(params, sr, mbf);
    }
};

// a11 — biquad cascade, transposed direct form II, f32, separately rounded ops:
//   y = (b0*x) + s1;  s1 = ((b1*x) - (a1*y)) + s2;  s2 = (b2*x) - (a2*y)
// Coefficients arrive as f32 (designed on the host in f64). No silence early-out
// (an IIR has a tail); out mask = NONE_SILENT.
struct BiquadCoeffs { float b0 = 1, b1 = 0, b2 = 0, a1 = 0, a2 = 0; };
constexpr size_t MAX_BIQUAD_STAGES = 8;
struct BiquadParams { uint32_t num_stages = 0; BiquadCoeffs st[MAX_BIQUAD_STAGES]; };
struct BiquadProcessor : AudioNodeProcessor {
    std::shared_ptr<BiquadParams> params; std::vector<float> state;  // [ch][stage][2]
    BiquadProcessor(std::shared_ptr<BiquadParams> p, size_t channels) : params(std::move(p)), state(channels * MAX_BIQUAD_STAGES * 2, 0.0f) {}
    void process(size_t frames, const std::vector<const float*>& inputs, const std::vector<float*>& outputs, ProcInfo) override {
        size_t n = std::min(inputs.size(), outputs.size());
        uint32_t ns = params->num_stages;
        for (size_t c = 0; c < n; ++c) {
            float* st = &state[c * MAX_BIQUAD_STAGES * 2];
            for (size_t i = 0; i < frames; ++i) {
                float x = inputs[c][i];
                for (uint32_t s = 0; s < ns; ++s) {
                    const BiquadCoeffs& k = params->st[s];
                    float s1 = st[2 * s], s2 = st[2 * s + 1];
                    float y = (k.b0 * x) + s1;
                    st[2 * s] = ((k.b1 * x) - (k.a1 * y)) + s2;
                    st[2 * s + 1] = (k.b2 * x) - (k.a2 * y);
                    x = y;
                }
                outputs[c][i] = x;
            }
        }
    }
};
struct BiquadNode : AudioNode {
    std::shared_ptr<BiquadParams> params;
    explicit BiquadNode(uint32_t num_stages) : params(std::make_shared<BiquadParams>()) { params->num_stages = std::min<uint32_t>(num_stages, MAX_BIQUAD_STAGES); }
    const char* debug_name() const override { return "biquad"; }
    AudioNodeInfo info() const override { return AudioNodeInfo{1, 64, 1, 64, false}; }
    std::unique_ptr<AudioNodeProcessor> activate(uint32_t, size_t, size_t ni, size_t no, std::string* err) override {
        if (ni != no) { if (err) *err = "The number of inputs on a BiquadNode must equal the number of outputs. Got num_inputs: " + std::to_string(ni) + ", num_outputs: " + std::to_string(no); return nullptr; }
        return std::make_unique<BiquadProcessor>(params, ni);
    }
};

// a12 — integer delay line: y[n] = x[n-D], zero-initialised ring per channel.
struct DelayParams { uint32_t delay = 0; };
struct DelayProcessor : AudioNodeProcessor {
    std::shared_ptr<DelayParams> params; uint32_t delay; std::vector<std::vector<float>> ring; size_t pos = 0;
    DelayProcessor(std::shared_ptr<DelayParams> p, size_t channels) : params(std::move(p)), delay(params->delay), ring(channels, std::vector<float>(std::max<uint32_t>(delay, 1), 0.0f)) {}
    void process(size_t frames, const std::vector<const float*>& inputs, const std::vector<float*>& outputs, ProcInfo) override {
        size_t n = std::min(inputs.size(), outputs.size());
        if (delay == 0) { for (size_t c = 0; c < n; ++c) std::memcpy(outputs[c], inputs[c], frames * sizeof(float)); return; }
        for (size_t c = 0; c < n; ++c) {
            size_t p = pos;
            for (size_t i = 0; i < frames; ++i) {
                float x = inputs[c][i];
                outputs[c][i] = ring[c][p];
                ring[c][p] = x;
                p = (p + 1 == delay) ? 0 : p + 1;
            }
        }
        pos = (pos + frames) % delay;
    }
};
struct DelayNode : AudioNode {
    std::shared_ptr<DelayParams> params;
    explicit DelayNode(uint32_t delay) : params(std::make_shared<DelayParams>()) { params->delay = delay; }
    const char* debug_name() const override { return "delay"; }
    AudioNodeInfo info() const override { return AudioNodeInfo{1, 64, 1, 64, false}; }
    std::unique_ptr<AudioNodeProcessor> activate(uint32_t, size_t, size_t ni, size_t no, std::string* err) override {
        if (ni != no) { if (err) *err = "The number of inputs on a DelayNode must equal the number of outputs. Got num_inputs: " + std::to_string(ni) + ", num_outputs: " + std::to_string(no); return nullptr; }
        return std::make_unique<DelayProcessor>(params, ni);
    }
};

// a14 — FIR convolutional reverb: y_c[n] = sum_{k<L} bf16(h_c[k]) * bf16(x_c[n-k]),
// products exact, accumulated here in f64 (the GPU accumulates in f32 on the
// tensor pipe; compared with normalised-max tolerance 1e-5). History carried
// across blocks. Per-channel IR (L->L, R->R); channel c uses IR row c % ir_channels.
inline float bf16_round(float x) {  // round-to-nearest-even to bfloat16, returned as f32
    uint32_t u; std::memcpy(&u, &x, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) { u |= 0x00400000u; u &= 0xffff0000u; }  // NaN stays NaN
    else { uint32_t lsb = (u >> 16) & 1u; u += 0x7fffu + lsb; u &= 0xffff0000u; }
    float r; std::memcpy(&r, &u, 4); return r;
}
struct ReverbParams { uint32_t ir_len = 0, ir_channels = 0; std::vector<float> ir; /* [ch][len], already bf16-rounded */ };
struct ConvReverbProcessor : AudioNodeProcessor {
    std::shared_ptr<ReverbParams> params; std::vector<std::vector<float>> hist;  // [ch][L-1] past inputs (bf16-rounded), oldest first
    ConvReverbProcessor(std::shared_ptr<ReverbParams> p, size_t channels) : params(std::move(p)), hist(channels, std::vector<float>(params->ir_len > 0 ? params->ir_len - 1 : 0, 0.0f)) {}
    void process(size_t frames, const std::vector<const float*>& inputs, const std::vector<float*>& outputs, ProcInfo) override {
        size_t n = std::min(inputs.size(), outputs.size());
        size_t L = params->ir_len, H = L > 0 ? L - 1 : 0;
        std::vector<float> ext(H + frames);
        for (size_t c = 0; c < n; ++c) {
            const float* h = &params->ir[(c % params->ir_channels) * L];
            std::memcpy(ext.data(), hist[c].data(), H * sizeof(float));
            for (size_t i = 0; i < frames; ++i) ext[H + i] = bf16_round(inputs[c][i]);
            for (size_t i = 0; i < frames; ++i) {
                double acc = 0.0;
                const float* xe = &ext[H + i];  // x[n-k] = xe[-k]
                for (size_t k = 0; k < L; ++k) acc += (double)h[k] * (double)xe[-(ptrdiff_t)k];
                outputs[c][i] = (float)acc;
            }
            if (H) std::memcpy(hist[c].data(), ext.data() + frames, H * sizeof(float));
        }
    }
};
struct ConvReverbNode : AudioNode {
    std::shared_ptr<ReverbParams> params;
    ConvReverbNode(const float* ir, uint32_t ir_len, uint32_t ir_channels) : params(std::make_shared<ReverbParams>()) {
        params->ir_len = ir_len; params->ir_channels = ir_channels; params->ir.resize((size_t)ir_len * ir_channels);
        for (size_t i = 0; i < params->ir.size(); ++i) params->ir[i] = bf16_round(ir[i]);
    }
    const char* debug_name() const override { return "conv_reverb"; }
    AudioNodeInfo info() const override { return AudioNodeInfo{1, 64, 1, 64, false}; }
    std::unique_ptr<AudioNodeProcessor> activate(uint32_t, size_t, size_t ni, size_t no, std::string* err) override {
        if (ni != no || params->ir_len == 0 || params->ir_channels == 0) { if (err) *err = "A ConvReverbNode needs num_inputs == num_outputs and a non-empty IR. Got num_inputs: " + std::to_string(ni) + ", num_outputs: " + std::to_string(no); return nullptr; }
        return std::make_unique<ConvReverbProcessor>(params, ni);
    }
};

// ---------------------------------------------------------------------------
// firewheel-core/src/sample_resource.rs:1-456 — sample resources (the sampler's data source)
// ---------------------------------------------------------------------------
inline float pcm_i16_to_f32(int16_t s) { return (float)s * (1.0f / (float)INT16_MAX); }            // :337-340
inline float pcm_u16_to_f32(uint16_t s) { return ((float)s * (2.0f / (float)UINT16_MAX)) - 1.0f; }  // :342-345
enum class SampleFormat : uint32_t { F32Planar = 0, F32Interleaved = 1, I16Interleaved = 2, U16Interleaved = 3, I16Planar = 4, U16Planar = 5 };
// One struct stands for the twelve `impl SampleResource` blocks (:33-335): interleaved data is [frame][ch], planar is
// [ch][frame]. All of fill_buffers_interleaved (:348-401, mono / stereo / generic loops), fill_buffers_deinterleaved
// (:404-439) and fill_buffers_deinterleaved_f32 (:442-456) reduce to: buffer c < min(buffers, channels) receives
// convert(data[c][start_frame + i]). A read past the end of the data panics in the reference (slice bounds); here it
// yields 0.0 — the one place this restatement defines behaviour the reference does not have.
struct SampleResource {
    SampleFormat fmt = SampleFormat::F32Planar; size_t channels = 1; uint64_t frames = 0;
    std::vector<float> f32; std::vector<int16_t> i16; std::vector<uint16_t> u16;
    size_t num_channels() const { return channels; }
    uint64_t len_frames() const { return frames; }
    float at(size_t ch, uint64_t frame) const {
        if (frame >= frames) return 0.0f;
        switch (fmt) {
            case SampleFormat::F32Planar: return f32[ch * frames + frame];
            case SampleFormat::F32Interleaved: return f32[frame * channels + ch];
            case SampleFormat::I16Interleaved: return pcm_i16_to_f32(i16[frame * channels + ch]);
            case SampleFormat::U16Interleaved: return pcm_u16_to_f32(u16[frame * channels + ch]);
            case SampleFormat::I16Planar: return pcm_i16_to_f32(i16[ch * frames + frame]);
            default: return pcm_u16_to_f32(u16[ch * frames + frame]);
        }
    }
    void fill_buffers(const std::vector<float*>& buffers, size_t r0, size_t r1, uint64_t start_frame) const {  // :20-25
        size_t n = std::min(buffers.size(), channels);
        for (size_t c = 0; c < n; ++c) for (size_t i = r0; i < r1; ++i) buffers[c][i] = at(c, start_frame + (i - r0));
    }
};

// ---------------------------------------------------------------------------
// basic_nodes/sampler.rs:14-577
// ---------------------------------------------------------------------------
constexpr size_t SAMPLER_CHANNEL_CAPACITY = 128;  // :14
struct SamplerMsg {  // NodeToProcessorMsg :21-28
    enum Kind : uint32_t { SetSample = 0, Play = 1, Pause = 2, Stop = 3, SetPlayheadSecs = 4, SetLoopRange = 5 } kind;
    explicit SamplerMsg(Kind k) : kind(k) {}
    std::shared_ptr<const SampleResource> sample; bool stop_playback = false;
    double secs = 0.0;                                  // SetPlayheadSecs
    uint32_t loop_mode = 0; double loop_start = 0.0, loop_end = 0.0;  // SetLoopRange: 0 None, 1 Full, 2 RangeSecs
};
struct SamplerShared {  // what the node and its processor share: the atomic gain and the two rings
    float raw_gain = 0.0f;
    std::deque<SamplerMsg> to_processor;                               // rtrb ring, capacity 128
    std::deque<std::shared_ptr<const SampleResource>> from_processor;  // ReturnSample
};
// `(secs * sample_rate).round() as u64` (:250-251,:394): Rust float->int casts saturate, NaN -> 0
inline uint64_t secs_to_frame(double secs, uint32_t sample_rate) {
    double f = std::round(secs * (double)sample_rate);
    if (!(f > 0.0)) return 0;
    if (f >= 18446744073709551616.0) return UINT64_MAX;
    return (uint64_t)f;
}
struct ProcLoopRange {  // :235-281
    uint64_t start = 0, end = 0; bool full_range = false;
    static ProcLoopRange make(uint32_t mode, double s, double e, uint32_t sample_rate, const std::shared_ptr<const SampleResource>& sample) {
        ProcLoopRange r;
        if (mode == 1) { r.start = 0; r.end = sample ? sample->len_frames() : 0; r.full_range = true; }
        else { r.start = secs_to_frame(s, sample_rate); r.end = secs_to_frame(e, sample_rate); r.full_range = false; }
        return r;
    }
    void update_sample(const std::shared_ptr<const SampleResource>& sample) { if (!sample || !full_range) return; start = 0; end = sample->len_frames(); }
    bool contains(uint64_t p) const { return p >= start && p < end; }
};
struct SamplerProcessor : AudioNodeProcessor {
    std::shared_ptr<SamplerShared> sh; ParamSmoother gain_smoother; bool playing = false; uint32_t sample_rate; uint64_t playhead = 0;
    bool has_loop = false; ProcLoopRange loop_range; std::shared_ptr<const SampleResource> sample;
    SamplerProcessor(std::shared_ptr<SamplerShared> s, uint32_t sr, size_t mbf) : sh(std::move(s)), gain_smoother(sh->raw_gain, sr, mbf), sample_rate(sr) {}
    ~SamplerProcessor() override { if (sample) sh->from_processor.push_back(sample); }  // :562-570
    uint64_t loop_start_or_zero() const { return has_loop ? loop_range.start : 0; }
    void process(size_t frames, const std::vector<const float*>&, const std::vector<float*>& outputs, ProcInfo pi) override {
        while (!sh->to_processor.empty()) {  // :331-414
            SamplerMsg m = std::move(sh->to_processor.front()); sh->to_processor.pop_front();
            switch (m.kind) {
                case SamplerMsg::SetSample:
                    if (sample) sh->from_processor.push_back(sample);
                    sample = m.sample;
                    if (has_loop) loop_range.update_sample(sample);
                    if (m.stop_playback) { playhead = loop_start_or_zero(); if (playing) playing = false; }
                    break;
                case SamplerMsg::Play: if (!playing) playing = true; break;
                case SamplerMsg::Pause: if (playing) playing = false; break;
                case SamplerMsg::Stop: playhead = loop_start_or_zero(); if (playing) playing = false; break;
                case SamplerMsg::SetPlayheadSecs: { uint64_t p = secs_to_frame(m.secs, sample_rate); if (p != playhead) playhead = p; break; }
                case SamplerMsg::SetLoopRange:
                    has_loop = m.loop_mode != 0;
                    if (has_loop) { loop_range = ProcLoopRange::make(m.loop_mode, m.loop_start, m.loop_end, sample_rate, sample); if (loop_range.contains(playhead)) playhead = loop_range.start; }
                    break;
            }
        }
        if (!sample) { clear_all_outputs(frames, outputs, pi.out_silence_mask); return; }   // :416-422
        if (!playing) { clear_all_outputs(frames, outputs, pi.out_silence_mask); return; }  // :424-430
        float raw_gain = sh->raw_gain;
        SmoothedOutput gain = gain_smoother.set_and_process(raw_gain, frames);  // :432-433
        // :435 assert_eq!(gain.values.len(), frames) panics for a short block while the smoother is inactive (Q1); a
        // panic is not a result, so short blocks are processed like full ones here.
        if (!gain.is_smoothing() && gain.values[0] < 0.00001f) { clear_all_outputs(frames, outputs, pi.out_silence_mask); return; }  // :437-443
        if (has_loop) {  // :445-484
            if (playhead >= loop_range.end) playhead = loop_range.start;
            uint64_t frames_left = loop_range.end - playhead;
            size_t first_copy_frames = (size_t)std::min<uint64_t>(frames, frames_left);
            sample->fill_buffers(outputs, 0, first_copy_frames, playhead);
            if (first_copy_frames < frames) {
                playhead = loop_range.start;
                size_t second_copy_frames = frames - first_copy_frames;
                sample->fill_buffers(outputs, first_copy_frames, frames, playhead);
                playhead += second_copy_frames;
            } else {
                playhead += frames;
            }
        } else {  // :485-516
            if (playhead >= sample->len_frames()) { playing = false; clear_all_outputs(frames, outputs, pi.out_silence_mask); return; }
            size_t copy_frames = (size_t)std::min<uint64_t>(frames, sample->len_frames() - playhead);
            sample->fill_buffers(outputs, 0, copy_frames, playhead);
            if (copy_frames < frames) {
                playing = false; playhead = 0;
                for (float* o : outputs) for (size_t i = copy_frames; i < frames; ++i) o[i] = 0.0f;
            } else {
                playhead += frames;
            }
        }
        size_t sample_channels = sample->num_channels();
        if (outputs.size() >= 2 && sample_channels == 2) {  // :522-533
            for (size_t i = 0; i < frames; ++i) { outputs[0][i] *= gain[i]; outputs[1][i] *= gain[i]; }
        } else {  // :534-543
            size_t n = std::min(outputs.size(), sample_channels);
            for (size_t c = 0; c < n; ++c) for (size_t i = 0; i < frames; ++i) outputs[c][i] *= gain[i];
        }
        if (outputs.size() > sample_channels) {  // :545-559
            if (outputs.size() == 2 && sample_channels == 1) {
                std::memcpy(outputs[1], outputs[0], frames * sizeof(float));
            } else {
                for (size_t c = sample_channels; c < outputs.size(); ++c) { for (size_t i = 0; i < frames; ++i) outputs[c][i] = 0.0f; pi.out_silence_mask->set_channel(c, true); }
            }
        }
    }
};
struct SamplerNode : AudioNode {  // :46-233
    std::shared_ptr<SamplerShared> sh; bool active = false; float percent_volume; bool playing = false;
    explicit SamplerNode(float percent) : sh(std::make_shared<SamplerShared>()) {
        percent = std::fmax(percent, 0.0f); sh->raw_gain = percent_volume_to_raw_gain(percent); percent_volume = percent;
    }
    // 0 ok, -2 ring full (push Err), -3 not activated (the reference hits todo!() there)
    int push(SamplerMsg m) {
        if (!active) return -3;
        if (sh->to_processor.size() >= SAMPLER_CHANNEL_CAPACITY) return -2;
        sh->to_processor.push_back(std::move(m)); return 0;
    }
    int set_sample(std::shared_ptr<const SampleResource> s, bool stop_playback) { SamplerMsg m(SamplerMsg::SetSample); m.sample = std::move(s); m.stop_playback = stop_playback; return push(std::move(m)); }
    int play() { if (playing) return 0; int rc = push(SamplerMsg(SamplerMsg::Play)); if (rc == 0) playing = true; return rc; }     // :82-98
    int pause() { if (!playing) return 0; int rc = push(SamplerMsg(SamplerMsg::Pause)); if (rc == 0) playing = false; return rc; }  // :101-117
    int stop() { if (!playing) return 0; int rc = push(SamplerMsg(SamplerMsg::Stop)); if (rc == 0) playing = false; return rc; }    // :120-136
    int set_playhead(double secs) { SamplerMsg m(SamplerMsg::SetPlayheadSecs); m.secs = secs; return push(std::move(m)); }
    int set_loop_range(uint32_t mode, double s, double e) { SamplerMsg m(SamplerMsg::SetLoopRange); m.loop_mode = mode; m.loop_start = s; m.loop_end = e; return push(std::move(m)); }
    void set_percent_volume(float p) { sh->raw_gain = percent_volume_to_raw_gain(p); percent_volume = std::fmax(p, 0.0f); }  // :174-180
    const char* debug_name() const override { return "beep_test"; }  // Q8 :186
    AudioNodeInfo info() const override { AudioNodeInfo i; i.num_min_supported_outputs = 1; i.num_max_supported_outputs = 64; i.updates = true; return i; }  // :189-196
    std::unique_ptr<AudioNodeProcessor> activate(uint32_t sr, size_t mbf, size_t, size_t, std::string*) override {  // :198-221
        sh->to_processor.clear(); sh->from_processor.clear();
        active = true;
        return std::make_unique<SamplerProcessor>(sh, sr, mbf);
    }
    void update() override { if (active) sh->from_processor.clear(); }  // :223-232
};

// ---------------------------------------------------------------------------
// SVF cascade (spec ours; parity unpinned by the reference): Simper/Cytomic trapezoidal
// state-variable filter. coeffs per stage {a1, a2, a3, m0, m1, m2}; op order as written (include/fw_b200.h).
// ---------------------------------------------------------------------------
struct SvfCoeffs { float a1 = 1, a2 = 0, a3 = 0, m0 = 1, m1 = 0, m2 = 0; };  // identity: y = x
struct SvfParams { uint32_t num_stages = 0; SvfCoeffs st[MAX_BIQUAD_STAGES]; };
struct SvfProcessor : AudioNodeProcessor {
    std::shared_ptr<SvfParams> params; std::vector<float> ic;  // [ch][stage][2]
    SvfProcessor(std::shared_ptr<SvfParams> p, size_t channels) : params(std::move(p)), ic(channels * MAX_BIQUAD_STAGES * 2, 0.0f) {}
    void process(size_t frames, const std::vector<const float*>& inputs, const std::vector<float*>& outputs, ProcInfo) override {
        size_t n = std::min(inputs.size(), outputs.size());
        for (size_t c = 0; c < n; ++c) {
            float* st = &ic[c * MAX_BIQUAD_STAGES * 2];
            for (size_t i = 0; i < frames; ++i) {
                float x = inputs[c][i];
                for (uint32_t s = 0; s < params->num_stages; ++s) {
                    const SvfCoeffs& k = params->st[s];
                    float& ic1 = st[s * 2]; float& ic2 = st[s * 2 + 1];
                    float v3 = x - ic2;
                    float v1 = (k.a1 * ic1) + (k.a2 * v3);
                    float v2 = ic2 + ((k.a2 * ic1) + (k.a3 * v3));
                    ic1 = (2.0f * v1) - ic1;
                    ic2 = (2.0f * v2) - ic2;
                    x = (k.m0 * x) + ((k.m1 * v1) + (k.m2 * v2));
                }
                outputs[c][i] = x;
            }
        }
    }
};
struct SvfNode : AudioNode {
    std::shared_ptr<SvfParams> params;
    explicit SvfNode(uint32_t ns) : params(std::make_shared<SvfParams>()) { params->num_stages = std::min<uint32_t>(ns, MAX_BIQUAD_STAGES); }
    const char* debug_name() const override { return "svf"; }
    AudioNodeInfo info() const override { return AudioNodeInfo{1, 64, 1, 64, false}; }
    std::unique_ptr<AudioNodeProcessor> activate(uint32_t, size_t, size_t ni, size_t no, std::string* err) override {
        if (ni != no) { if (err) *err = "The number of inputs on an SvfNode must equal the number of outputs. Got num_inputs: " + std::to_string(ni) + ", num_outputs: " + std::to_string(no); return nullptr; }
        return std::make_unique<SvfProcessor>(params, ni);
    }
};

// ---------------------------------------------------------------------------
// Polyphase resampler (spec ours; parity unpinned): a sample player with a Q32.32 position and step
// reading a SampleResource through a P-phase, T-tap windowed-sinc table. Semantics: include/fw_b200.h.
// ---------------------------------------------------------------------------
struct ResamplerShared {
    std::vector<float> table; uint32_t phases = 0, taps = 0, phase_shift = 32;
    std::shared_ptr<const SampleResource> sample; uint64_t step = 1ull << 32; bool playing = false, loop = false;
    bool seek_pending = false; uint64_t seek_frames = 0;
};
struct ResamplerProcessor : AudioNodeProcessor {
    std::shared_ptr<ResamplerShared> sh; uint64_t pos = 0;
    explicit ResamplerProcessor(std::shared_ptr<ResamplerShared> s) : sh(std::move(s)) {}
    void process(size_t frames, const std::vector<const float*>&, const std::vector<float*>& outputs, ProcInfo pi) override {
        if (sh->seek_pending) { pos = sh->seek_frames << 32; sh->seek_pending = false; }
        if (!sh->sample || !sh->playing) { clear_all_outputs(frames, outputs, pi.out_silence_mask); return; }
        const SampleResource& smp = *sh->sample;
        const int64_t len = (int64_t)smp.len_frames(); const size_t sch = smp.num_channels(), T = sh->taps;
        const size_t filled = std::min(outputs.size(), sch);
        for (size_t c = 0; c < filled; ++c) {
            for (size_t n = 0; n < frames; ++n) {
                const uint64_t p = pos + (uint64_t)n * sh->step;
                const int64_t i0 = (int64_t)(p >> 32) - (int64_t)(T / 2 - 1);
                const float* h = &sh->table[(size_t)((uint32_t)(p & 0xffffffffull) >> sh->phase_shift) * T];
                float y = 0.0f;
                for (size_t t = 0; t < T; ++t) {
                    int64_t idx = i0 + (int64_t)t; float x;
                    if (sh->loop) { idx %= len; if (idx < 0) idx += len; x = smp.at(c, (uint64_t)idx); }
                    else x = (idx >= 0 && idx < len) ? smp.at(c, (uint64_t)idx) : 0.0f;
                    y = y + (h[t] * x);
                }
                outputs[c][n] = y;
            }
        }
        if (outputs.size() > sch) {
            if (outputs.size() == 2 && sch == 1) std::memcpy(outputs[1], outputs[0], frames * sizeof(float));
            else for (size_t c = sch; c < outputs.size(); ++c) { for (size_t i = 0; i < frames; ++i) outputs[c][i] = 0.0f; pi.out_silence_mask->set_channel(c, true); }
        }
        pos += (uint64_t)frames * sh->step;
    }
};
struct ResamplerNode : AudioNode {
    std::shared_ptr<ResamplerShared> sh;
    ResamplerNode(const float* table, uint32_t phases, uint32_t taps) : sh(std::make_shared<ResamplerShared>()) {
        sh->phases = phases; sh->taps = taps; sh->table.assign(table, table + (size_t)phases * taps);
        uint32_t lg = 0; while ((1u << lg) < phases) ++lg;
        sh->phase_shift = 32 - lg;
    }
    const char* debug_name() const override { return "resampler"; }
    AudioNodeInfo info() const override { AudioNodeInfo i; i.num_min_supported_outputs = 1; i.num_max_supported_outputs = 64; return i; }
    std::unique_ptr<AudioNodeProcessor> activate(uint32_t, size_t, size_t, size_t no, std::string* err) override {
        const uint32_t P = sh->phases, T = sh->taps;
        if (no == 0 || P == 0 || P > 1024 || (P & (P - 1)) || T < 2 || T > 64 || (T & 1)) { if (err) *err = "A ResamplerNode needs >= 1 output, a power-of-two phase count <= 1024 and an even tap count <= 64."; return nullptr; }
        return std::make_unique<ResamplerProcessor>(sh);
    }
};

// ===========================================================================
// firewheel-graph: graph.rs, graph/compiler.rs, graph/compiler/schedule.rs,
// graph/error.rs, processor.rs, context.rs
// ===========================================================================
struct NodeID { Index idx; const char* debug_name = "dangling";
    bool operator==(const NodeID& o) const { return idx == o.idx; }  // graph.rs:38-42
    bool operator!=(const NodeID& o) const { return !(idx == o.idx); }
    bool operator<(const NodeID& o) const { return idx < o.idx; } };
struct EdgeID { Index idx; bool operator==(const EdgeID& o) const { return idx == o.idx; } };
struct Edge { EdgeID id; NodeID src_node; uint32_t src_port; NodeID dst_node; uint32_t dst_port; };  // compiler.rs:68-78

enum class AddEdgeError : int {  // error.rs:14-37
    Ok = 0, SrcNodeNotFound = 1, DstNodeNotFound = 2, InPortOutOfRange = 3, OutPortOutOfRange = 4,
    EdgeAlreadyExists = 5, InputPortAlreadyConnected = 6, CycleDetected = 7 };
enum class CompileGraphError : int {  // error.rs:101-116
    Ok = 0, CycleDetected = 1, NodeOnEdgeNotFound = 2, NodeIDNotUnique = 3, EdgeIDNotUnique = 4, ManyToOneError = 5,
    NodeActivationFailed = 6, MessageChannelFull = 7 };
struct CompileErrorInfo { CompileGraphError code = CompileGraphError::Ok; NodeID node; uint32_t port = 0; std::string message; };

struct InBufferAssignment { size_t buffer_index; bool should_clear; size_t generation; };   // schedule.rs:105-115
struct OutBufferAssignment { size_t buffer_index; size_t generation; };                      // schedule.rs:119-126
struct ScheduledNode { NodeID id; std::vector<InBufferAssignment> input_buffers; std::vector<OutBufferAssignment> output_buffers; };

class CompiledSchedule {  // schedule.rs:166-344
  public:
    std::vector<ScheduledNode> schedule; std::vector<float> buffers; std::vector<uint8_t> buffer_silence_flags;
    size_t num_buffers = 0, max_block_frames = 0;
    // scratch for the ArrayVec<_, 64> the reference keeps on the stack (schedule.rs:223,265,296-297): no per-block allocation
    std::vector<const float*> scratch_in_; std::vector<float*> scratch_out_;
    CompiledSchedule(std::vector<ScheduledNode> s, size_t nb, size_t mbf)
        : schedule(std::move(s)), buffers(nb * mbf, 0.0f), buffer_silence_flags(nb, 0), num_buffers(nb), max_block_frames(mbf) {}
    float* buffer_slice(size_t buffer_index) { return buffers.data() + buffer_index * max_block_frames; }  // :347-379

    template <class F> void prepare_graph_inputs(size_t frames, size_t num_stream_inputs, F&& fill_inputs) {  // :213-253
        frames = std::min(frames, max_block_frames);
        ScheduledNode& gin = schedule.front();
        std::vector<float*>& inputs = scratch_out_; inputs.clear();
        size_t fill_len = std::min(num_stream_inputs, gin.output_buffers.size());
        for (size_t i = 0; i < fill_len; ++i) inputs.push_back(buffer_slice(gin.output_buffers[i].buffer_index));
        SilenceMask m = fill_inputs(inputs, frames);
        for (size_t i = 0; i < fill_len; ++i) buffer_silence_flags[gin.output_buffers[i].buffer_index] = m.is_channel_silent(i);
        for (size_t i = fill_len; i < gin.output_buffers.size(); ++i) {
            float* b = buffer_slice(gin.output_buffers[i].buffer_index);
            for (size_t f = 0; f < frames; ++f) b[f] = 0.0f;
            buffer_silence_flags[gin.output_buffers[i].buffer_index] = 1;
        }
    }
    template <class F> void read_graph_outputs(size_t frames, size_t num_stream_outputs, F&& read_outputs) {  // :255-287
        frames = std::min(frames, max_block_frames);
        ScheduledNode& gout = schedule.back();
        std::vector<const float*>& outputs = scratch_in_; outputs.clear(); SilenceMask m = SilenceMask::none();
        size_t read_len = std::min(num_stream_outputs, gout.input_buffers.size());
        for (size_t i = 0; i < read_len; ++i) {
            size_t bi = gout.input_buffers[i].buffer_index;
            if (buffer_silence_flags[bi]) m.set_channel(i, true);
            outputs.push_back(buffer_slice(bi));
        }
        read_outputs(outputs, m, frames);
    }
    template <class F> void process(size_t frames, F&& proc) {  // :289-343
        frames = std::min(frames, max_block_frames);
        std::vector<const float*>& inputs = scratch_in_; std::vector<float*>& outputs = scratch_out_;
        for (ScheduledNode& sn : schedule) {
            SilenceMask in_mask = SilenceMask::none();
            inputs.clear(); outputs.clear();
            for (size_t i = 0; i < sn.input_buffers.size(); ++i) {
                const InBufferAssignment& b = sn.input_buffers[i];
                float* buf = buffer_slice(b.buffer_index);
                if (b.should_clear) { for (size_t f = 0; f < frames; ++f) buf[f] = 0.0f; buffer_silence_flags[b.buffer_index] = 1; }
                if (buffer_silence_flags[b.buffer_index]) in_mask.set_channel(i, true);
                inputs.push_back(buf);
            }
            for (const OutBufferAssignment& b : sn.output_buffers) outputs.push_back(buffer_slice(b.buffer_index));
            SilenceMask out_mask = proc(sn.id, in_mask, inputs, outputs);
            for (size_t i = 0; i < sn.output_buffers.size(); ++i)
                buffer_silence_flags[sn.output_buffers[i].buffer_index] = out_mask.is_channel_silent(i);
        }
    }
};

struct ScheduleHeapData {  // schedule.rs:128-150
    CompiledSchedule schedule; std::vector<NodeID> nodes_to_remove;
    std::vector<std::pair<NodeID, std::unique_ptr<AudioNodeProcessor>>> removed_node_processors, new_node_processors;
    ScheduleHeapData(CompiledSchedule s, std::vector<NodeID> rm, std::vector<std::pair<NodeID, std::unique_ptr<AudioNodeProcessor>>> nw)
        : schedule(std::move(s)), nodes_to_remove(std::move(rm)), new_node_processors(std::move(nw)) {}
};

struct NodeWeight { std::unique_ptr<AudioNode> node; bool activated = false; bool updates = false; };  // graph.rs:76-80
struct NodeEntry {  // compiler.rs:12-39
    NodeID id; uint32_t num_inputs = 0, num_outputs = 0; NodeWeight weight; std::vector<Edge> incoming, outgoing;
};

// compiler.rs:92-136
struct BufferRef { size_t idx; size_t generation; };
struct BufferAllocator {
    std::vector<BufferRef> free_list; size_t count = 0;
    std::shared_ptr<BufferRef> acquire() {
        BufferRef e;
        if (!free_list.empty()) { e = free_list.back(); free_list.pop_back(); } else { e = BufferRef{count, 0}; count += 1; }
        return std::make_shared<BufferRef>(e);
    }
    void release(std::shared_ptr<BufferRef>& r) {  // Rc::strong_count == 1
        if (r.use_count() == 1) free_list.push_back(BufferRef{r->idx, r->generation + 1});
        r.reset();
    }
};

// compiler.rs:139-418
class GraphCompiler {
  public:
    static CompileErrorInfo sort_topologically(Arena<NodeEntry>& nodes, NodeID gin, NodeID gout, bool build, std::vector<ScheduledNode>* schedule) {
        size_t nslots = nodes.num_slots();
        std::vector<int32_t> in_degree(nslots, 0); std::deque<uint32_t> queue; size_t num_visited = 0;
        nodes.for_each([&](Index, NodeEntry& n) { for (const Edge& e : n.outgoing) in_degree[e.dst_node.idx.slot] += 1; });
        queue.push_back(gin.idx.slot);  // :252
        nodes.for_each([&](Index, NodeEntry& n) { if (n.incoming.empty() && n.id.idx.slot != gin.idx.slot) queue.push_back(n.id.idx.slot); });
        while (!queue.empty()) {
            uint32_t slot = queue.front(); queue.pop_front(); num_visited += 1;
            NodeEntry* n = nodes.get_by_slot(slot).second;
            for (const Edge& e : n->outgoing) {
                in_degree[e.dst_node.idx.slot] -= 1;
                if (in_degree[e.dst_node.idx.slot] == 0) queue.push_back(e.dst_node.idx.slot);
            }
            if (build && slot != gout.idx.slot) schedule->push_back(ScheduledNode{n->id, {}, {}});
        }
        if (build) schedule->push_back(ScheduledNode{gout, {}, {}});  // :291
        CompileErrorInfo err;
        if (num_visited != nodes.len()) err.code = CompileGraphError::CycleDetected;  // :295-297
        return err;
    }
    static void preprocess(Arena<NodeEntry>& nodes, Arena<Edge>& edges) {  // :191-228
        nodes.for_each([&](Index, NodeEntry& n) { assert(n.num_inputs <= 64 && n.num_outputs <= 64); n.incoming.clear(); n.outgoing.clear(); });
        edges.for_each([&](Index, Edge& e) { nodes.at(e.src_node.idx).outgoing.push_back(e); nodes.at(e.dst_node.idx).incoming.push_back(e); });
    }
    static bool cycle_detected(Arena<NodeEntry>& nodes, Arena<Edge>& edges, NodeID gin, NodeID gout) {  // :154-168
        preprocess(nodes, edges);
        return sort_topologically(nodes, gin, gout, false, nullptr).code == CompileGraphError::CycleDetected;
    }
    static CompileErrorInfo compile(Arena<NodeEntry>& nodes, Arena<Edge>& edges, NodeID gin, NodeID gout, size_t max_block_frames,
                                    std::unique_ptr<CompiledSchedule>* out) {
        preprocess(nodes, edges);
        std::vector<ScheduledNode> schedule;
        CompileErrorInfo err = sort_topologically(nodes, gin, gout, true, &schedule);
        if (err.code != CompileGraphError::Ok) return err;
        // solve_buffer_requirements :302-412
        BufferAllocator allocator;
        std::map<Index, std::shared_ptr<BufferRef>> assignment_table;  // Arena keyed by edge index
        std::vector<std::shared_ptr<BufferRef>> to_release;
        for (ScheduledNode& entry : schedule) {
            NodeEntry& ne = nodes.at(entry.id.idx);
            to_release.clear();
            for (uint32_t port = 0; port < ne.num_inputs; ++port) {
                std::vector<const Edge*> es;
                for (const Edge& e : ne.incoming) if (e.dst_port == port) es.push_back(&e);
                if (es.empty()) {
                    auto buf = allocator.acquire();
                    entry.input_buffers.push_back(InBufferAssignment{buf->idx, true, buf->generation});
                    to_release.push_back(std::move(buf));
                } else if (es.size() == 1) {
                    auto it = assignment_table.find(es[0]->id.idx);
                    assert(it != assignment_table.end() && "No buffer assigned to edge!");
                    auto buf = std::move(it->second); assignment_table.erase(it);
                    entry.input_buffers.push_back(InBufferAssignment{buf->idx, false, buf->generation});
                    to_release.push_back(std::move(buf));
                } else {
                    err.code = CompileGraphError::ManyToOneError; err.node = entry.id; err.port = port;  // :363-365
                    return err;
                }
            }
            for (uint32_t port = 0; port < ne.num_outputs; ++port) {
                std::vector<const Edge*> es;
                for (const Edge& e : ne.outgoing) if (e.src_port == port) es.push_back(&e);
                if (es.empty()) {
                    auto buf = allocator.acquire();
                    entry.output_buffers.push_back(OutBufferAssignment{buf->idx, buf->generation});
                    to_release.push_back(std::move(buf));
                } else {
                    auto buf = allocator.acquire();
                    for (const Edge* e : es) assignment_table[e->id.idx] = buf;
                    entry.output_buffers.push_back(OutBufferAssignment{buf->idx, buf->generation});
                }  // `buf` dropped here: only the assignment table holds it
            }
            // drain(..): each element is moved out, released, dropped — later elements still hold their counts
            for (size_t i = 0; i < to_release.size(); ++i) allocator.release(to_release[i]);
        }
        *out = std::make_unique<CompiledSchedule>(std::move(schedule), allocator.count, max_block_frames);
        return err;
    }
};

struct AudioGraphConfig { size_t num_graph_inputs = 0, num_graph_outputs = 2, initial_node_capacity = 64, initial_edge_capacity = 256; };  // graph.rs:91-107

class AudioGraph {  // graph.rs:109-698
  public:
    Arena<NodeEntry> nodes; Arena<Edge> edges;
    std::set<std::pair<Index, uint32_t>> connected_input_ports;
    struct EdgeHash { Index s; uint32_t sp; Index d; uint32_t dp;
        bool operator<(const EdgeHash& o) const { return std::tie(s, sp, d, dp) < std::tie(o.s, o.sp, o.d, o.dp); } };
    std::map<EdgeHash, EdgeID> existing_edges;
    NodeID graph_in_id, graph_out_id; bool needs_compile_ = true;
    std::vector<NodeID> nodes_to_remove_from_schedule, nodes_to_activate;
    std::map<Index, NodeEntry> active_nodes_to_remove;

    explicit AudioGraph(const AudioGraphConfig& c) : nodes(c.initial_node_capacity), edges(c.initial_edge_capacity) {  // :125-168
        graph_in_id = insert_node(0, c.num_graph_inputs, std::make_unique<DummyAudioNode>(), "graph_in");
        graph_out_id = insert_node(c.num_graph_outputs, 0, std::make_unique<DummyAudioNode>(), "graph_out");
        nodes_to_activate = {graph_in_id, graph_out_id};
    }
    NodeID graph_in_node() const { return graph_in_id; }
    NodeID graph_out_node() const { return graph_out_id; }
    size_t current_node_capacity() const { return nodes.capacity(); }

    NodeID add_node(size_t num_inputs, size_t num_outputs, std::unique_ptr<AudioNode> node) {  // :201-231
        const char* name = node->debug_name();
        NodeID id = insert_node(num_inputs, num_outputs, std::move(node), name);
        nodes_to_activate.push_back(id);
        needs_compile_ = true;
        return id;
    }
    AudioNode* node(NodeID id) { NodeEntry* n = nodes.get(id.idx); return n ? n->weight.node.get() : nullptr; }  // :237-247
    const NodeEntry* node_info(NodeID id) const { return nodes.get(id.idx); }                                      // :253

    bool remove_node(NodeID id, std::vector<EdgeID>* removed) {  // :268-299
        if (id == graph_in_id || id == graph_out_id) return false;
        std::optional<NodeEntry> ne = nodes.remove(id.idx);
        if (!ne) return false;
        for (uint32_t p = 0; p < ne->num_inputs; ++p) remove_edges_with_input_port(id, p, removed);
        for (uint32_t p = 0; p < ne->num_outputs; ++p) remove_edges_with_output_port(id, p, removed);
        for (uint32_t p = 0; p < ne->num_inputs; ++p) connected_input_ports.erase({id.idx, p});
        nodes_to_remove_from_schedule.push_back(id);
        if (ne->weight.activated) active_nodes_to_remove.emplace(id.idx, std::move(*ne));
        needs_compile_ = true;
        return true;
    }
    bool set_num_inputs(NodeID id, size_t num_inputs, std::vector<EdgeID>* removed) {  // :315-343
        if (id == graph_in_id) return false;
        NodeEntry* ne = nodes.get(id.idx); if (!ne) return false;
        uint32_t ni = (uint32_t)num_inputs, old = ne->num_inputs;
        if (ni < old) for (uint32_t p = ni; p < old; ++p) { remove_edges_with_input_port(id, p, removed); connected_input_ports.erase({id.idx, p}); }
        nodes.at(id.idx).num_inputs = ni; needs_compile_ = true; return true;
    }
    bool set_num_outputs(NodeID id, size_t num_outputs, std::vector<EdgeID>* removed) {  // :349-375
        if (id == graph_out_id) return false;
        NodeEntry* ne = nodes.get(id.idx); if (!ne) return false;
        uint32_t no = (uint32_t)num_outputs, old = ne->num_outputs;
        if (no < old) for (uint32_t p = no; p < old; ++p) remove_edges_with_output_port(id, p, removed);
        nodes.at(id.idx).num_outputs = no; needs_compile_ = true; return true;
    }
    AddEdgeError connect(NodeID src, uint32_t src_port, NodeID dst, uint32_t dst_port, bool check_for_cycles, EdgeID* out_id) {  // :396-477
        const NodeEntry* s = nodes.get(src.idx); if (!s) return AddEdgeError::SrcNodeNotFound;
        const NodeEntry* d = nodes.get(dst.idx); if (!d) return AddEdgeError::DstNodeNotFound;
        if (src_port >= s->num_outputs) return AddEdgeError::OutPortOutOfRange;
        if (dst_port >= d->num_inputs) return AddEdgeError::InPortOutOfRange;
        if (src.idx == dst.idx) return AddEdgeError::CycleDetected;
        EdgeHash h{src.idx, src_port, dst.idx, dst_port};
        if (existing_edges.count(h)) return AddEdgeError::EdgeAlreadyExists;
        if (!connected_input_ports.insert({dst.idx, dst_port}).second) return AddEdgeError::InputPortAlreadyConnected;
        Edge e{EdgeID{}, src, src_port, dst, dst_port};
        EdgeID id{edges.insert(e)};
        edges.at(id.idx).id = id;
        existing_edges[h] = id;
        if (check_for_cycles && cycle_detected()) {
            // Q9: graph.rs:467-471 removes the edge only from `edges`; `existing_edges` and
            // `connected_input_ports` keep their stale entries. Restated verbatim.
            edges.remove(id.idx);
            return AddEdgeError::CycleDetected;
        }
        needs_compile_ = true;
        if (out_id) *out_id = id;
        return AddEdgeError::Ok;
    }
    bool disconnect(NodeID src, uint32_t src_port, NodeID dst, uint32_t dst_port) {  // :483-501
        auto it = existing_edges.find(EdgeHash{src.idx, src_port, dst.idx, dst_port});
        if (it == existing_edges.end()) return false;
        EdgeID id = it->second; existing_edges.erase(it);
        disconnect_by_edge_id(id);
        return true;
    }
    bool disconnect_by_edge_id(EdgeID id) {  // :507-524
        std::optional<Edge> e = edges.remove(id.idx);
        if (!e) return false;
        existing_edges.erase(EdgeHash{e->src_node.idx, e->src_port, e->dst_node.idx, e->dst_port});
        connected_input_ports.erase({e->dst_node.idx, e->dst_port});
        needs_compile_ = true;
        return true;
    }
    const Edge* edge(EdgeID id) const { return edges.get(id.idx); }  // :527
    bool cycle_detected() { return GraphCompiler::cycle_detected(nodes, edges, graph_in_id, graph_out_id); }  // :573
    bool needs_compile() const { return needs_compile_; }
    void reset() {  // :171-182
        std::vector<NodeID> ids;
        nodes.for_each([&](Index, NodeEntry& n) { if (n.id != graph_in_id && n.id != graph_out_id) ids.push_back(n.id); });
        for (NodeID id : ids) remove_node(id, nullptr);
    }
    CompileErrorInfo compile_internal(size_t max_block_frames, std::unique_ptr<CompiledSchedule>* out) {  // :629-642
        assert(max_block_frames > 0);
        return GraphCompiler::compile(nodes, edges, graph_in_id, graph_out_id, max_block_frames, out);
    }
    CompileErrorInfo compile(uint32_t sample_rate, size_t max_block_frames, std::unique_ptr<ScheduleHeapData>* out) {  // :586-627
        std::unique_ptr<CompiledSchedule> sched;
        CompileErrorInfo err = compile_internal(max_block_frames, &sched);
        if (err.code != CompileGraphError::Ok) return err;
        std::vector<std::pair<NodeID, std::unique_ptr<AudioNodeProcessor>>> procs;
        for (NodeID id : nodes_to_activate) {
            NodeEntry* ne = nodes.get(id.idx);
            if (!ne) continue;
            std::string msg;
            auto p = ne->weight.node->activate(sample_rate, max_block_frames, ne->num_inputs, ne->num_outputs, &msg);
            if (p) { procs.emplace_back(id, std::move(p)); }
            else {
                for (auto& np : procs) nodes.at(np.first.idx).weight.node->deactivate(std::move(np.second));  // rollback :603-609
                err.code = CompileGraphError::NodeActivationFailed; err.node = id; err.message = msg;
                return err;
            }
        }
        *out = std::make_unique<ScheduleHeapData>(std::move(*sched), nodes_to_remove_from_schedule, std::move(procs));
        needs_compile_ = false; nodes_to_activate.clear(); nodes_to_remove_from_schedule.clear();
        return err;
    }
    void on_schedule_returned(std::unique_ptr<ScheduleHeapData> sd) {  // :644-658
        for (auto& rp : sd->removed_node_processors) {
            auto it = active_nodes_to_remove.find(rp.first.idx);
            if (it != active_nodes_to_remove.end()) {
                it->second.weight.node->deactivate(std::move(rp.second)); active_nodes_to_remove.erase(it);
            } else if (NodeEntry* ne = nodes.get(rp.first.idx)) {
                if (ne->weight.activated) {  // Q5: never true
                    ne->weight.node->deactivate(std::move(rp.second)); ne->weight.activated = false; nodes_to_activate.push_back(rp.first);
                }
            }
        }
        sd->removed_node_processors.clear();
    }
    void on_processor_dropped(Arena<std::unique_ptr<AudioNodeProcessor>>& procs) {  // :660-669
        for (auto& kv : procs.drain()) {
            if (NodeEntry* ne = nodes.get(kv.first)) if (ne->weight.activated) { ne->weight.node->deactivate(std::move(kv.second)); ne->weight.activated = false; }
        }
    }
    void deactivate() {  // :671-689
        active_nodes_to_remove.clear(); nodes_to_remove_from_schedule.clear(); needs_compile_ = true;
        nodes.for_each([&](Index idx, NodeEntry& ne) {
            if (ne.weight.activated) { ne.weight.node->deactivate(nullptr); ne.weight.activated = false; }
            nodes_to_activate.push_back(NodeID{idx, ne.weight.node->debug_name()});
        });
    }
    void update() { nodes.for_each([&](Index, NodeEntry& ne) { if (ne.weight.updates) ne.weight.node->update(); }); }  // :691-697

  private:
    NodeID insert_node(size_t ni, size_t no, std::unique_ptr<AudioNode> node, const char* name) {
        NodeEntry ne; ne.num_inputs = (uint32_t)ni; ne.num_outputs = (uint32_t)no;
        ne.weight.updates = node->info().updates; ne.weight.node = std::move(node);
        NodeID id{nodes.insert(std::move(ne)), name};
        nodes.at(id.idx).id = id;
        return id;
    }
    void remove_edges_with_input_port(NodeID id, uint32_t port, std::vector<EdgeID>* removed) {  // :531-550
        std::vector<EdgeID> rm;
        edges.for_each([&](Index ei, Edge& e) { if (e.dst_node == id && e.dst_port == port) rm.push_back(EdgeID{ei}); });
        for (EdgeID e : rm) disconnect_by_edge_id(e);
        if (removed) removed->insert(removed->end(), rm.begin(), rm.end());
    }
    void remove_edges_with_output_port(NodeID id, uint32_t port, std::vector<EdgeID>* removed) {  // :552-571
        std::vector<EdgeID> rm;
        edges.for_each([&](Index ei, Edge& e) { if (e.src_node == id && e.src_port == port) rm.push_back(EdgeID{ei}); });
        for (EdgeID e : rm) disconnect_by_edge_id(e);
        if (removed) removed->insert(removed->end(), rm.begin(), rm.end());
    }
};

// rtrb::RingBuffer (0.3.1) restated as a bounded FIFO; single-threaded here.
template <class T> struct Ring {
    std::deque<T> q; size_t cap;
    explicit Ring(size_t c) : cap(c) {}
    bool push(T&& v) { if (q.size() >= cap) return false; q.push_back(std::move(v)); return true; }
    bool pop(T* out) { if (q.empty()) return false; *out = std::move(q.front()); q.pop_front(); return true; }
};

enum class ProcessorStatus : int { Ok = 0, DropProcessor = 1 };  // processor.rs:12-16

struct ContextToProcessorMsg { bool stop = false; std::unique_ptr<ScheduleHeapData> new_schedule; };  // processor.rs:265-268
struct ProcessorToContextMsg {  // processor.rs:270-277
    bool dropped = false; std::unique_ptr<ScheduleHeapData> schedule;
    Arena<std::unique_ptr<AudioNodeProcessor>> nodes; void* user_cx = nullptr;
};
struct Channels { Ring<ContextToProcessorMsg> to_proc{16}; Ring<ProcessorToContextMsg> to_ctx{16}; };  // context.rs:14,61-64

class FirewheelProcessor {  // processor.rs:18-263
  public:
    Arena<std::unique_ptr<AudioNodeProcessor>> nodes; std::unique_ptr<ScheduleHeapData> schedule_data; void* user_cx;
    std::shared_ptr<Channels> ch; bool running = true; size_t max_block_frames;
    FirewheelProcessor(std::shared_ptr<Channels> c, size_t node_capacity, size_t n_in, size_t n_out, size_t mbf, void* cx)
        : nodes(node_capacity * 2), user_cx(cx), ch(std::move(c)), max_block_frames(mbf) { assert(n_in <= 64 && n_out <= 64); }
    ~FirewheelProcessor() {  // Drop :251-263
        ProcessorToContextMsg m; m.dropped = true; m.nodes = std::move(nodes); m.schedule = std::move(schedule_data); m.user_cx = user_cx;
        ch->to_ctx.push(std::move(m));
    }
    ProcessorStatus process_interleaved(const float* input, size_t input_len, float* output, size_t output_len, size_t n_in, size_t n_out,
                                        size_t frames, double stream_time_secs, uint32_t stream_status) {  // :61-165
        auto fill0 = [&](size_t from) { for (size_t i = from; i < output_len; ++i) output[i] = 0.0f; };
        if (!running) { fill0(0); return ProcessorStatus::DropProcessor; }
        if (!schedule_data) { poll_messages(); if (!running) { fill0(0); return ProcessorStatus::DropProcessor; } }
        if (!schedule_data || frames == 0) { fill0(0); return ProcessorStatus::Ok; }
        assert(input_len == frames * n_in); assert(output_len == frames * n_out);
        size_t done = 0;
        while (done < frames) {
            size_t bf = std::min(frames - done, max_block_frames);
            schedule_data->schedule.prepare_graph_inputs(bf, n_in, [&](std::vector<float*>& chans, size_t fr) {
                return deinterleave(chans, fr, input + done * n_in, bf * n_in, n_in, true);
            });
            process_block(bf, stream_time_secs, stream_status);
            schedule_data->schedule.read_graph_outputs(bf, n_out, [&](const std::vector<const float*>& chans, SilenceMask m, size_t fr) {
                float* dst = output + done * n_out; size_t dst_len = bf * n_out;
                if (chans.size() == 2 && n_out == 2) interleave_stereo(chans[0], chans[1], dst, dst_len, &m);
                else interleave(chans, fr, dst, dst_len, n_out, &m);
            });
            if (!running) { if (done < frames) fill0(done * n_out); break; }
            done += bf;
        }
        return running ? ProcessorStatus::Ok : ProcessorStatus::DropProcessor;
    }
    // Planar variant used by the batched drivers ("driver bypasses interleave and feeds planar").
    // Same loop as process_interleaved with memcpy in place of (de)interleave; the fill closure returns
    // NONE_SILENT, which is what survives downstream anyway (Q4).
    ProcessorStatus process_planar(const float* const* input, float* const* output, size_t n_in, size_t n_out, size_t frames,
                                   double stream_time_secs, uint32_t stream_status, uint64_t* out_mask_last) {
        auto fill0 = [&](size_t from) { for (size_t c = 0; c < n_out; ++c) for (size_t i = from; i < frames; ++i) output[c][i] = 0.0f; };
        if (!running) { fill0(0); return ProcessorStatus::DropProcessor; }
        if (!schedule_data) { poll_messages(); if (!running) { fill0(0); return ProcessorStatus::DropProcessor; } }
        if (!schedule_data || frames == 0) { fill0(0); return ProcessorStatus::Ok; }
        size_t done = 0;
        while (done < frames) {
            size_t bf = std::min(frames - done, max_block_frames);
            schedule_data->schedule.prepare_graph_inputs(bf, n_in, [&](std::vector<float*>& chans, size_t fr) {
                for (size_t c = 0; c < chans.size(); ++c) std::memcpy(chans[c], input[c] + done, fr * sizeof(float));
                return SilenceMask::none();
            });
            process_block(bf, stream_time_secs, stream_status);
            schedule_data->schedule.read_graph_outputs(bf, n_out, [&](const std::vector<const float*>& chans, SilenceMask m, size_t fr) {
                for (size_t c = 0; c < n_out; ++c) {
                    if (c < chans.size()) std::memcpy(output[c] + done, chans[c], fr * sizeof(float));
                    else for (size_t i = 0; i < fr; ++i) output[c][done + i] = 0.0f;
                }
                if (out_mask_last) *out_mask_last = m.bits;
            });
            if (!running) { fill0(done); break; }
            done += bf;
        }
        return running ? ProcessorStatus::Ok : ProcessorStatus::DropProcessor;
    }

  private:
    void poll_messages() {  // :167-206
        ContextToProcessorMsg msg;
        while (ch->to_proc.pop(&msg)) {
            if (msg.stop) { running = false; continue; }
            std::unique_ptr<ScheduleHeapData> nw = std::move(msg.new_schedule);
            assert(nw->schedule.max_block_frames == max_block_frames);
            if (schedule_data) {
                std::unique_ptr<ScheduleHeapData> old = std::move(schedule_data);
                std::swap(old->removed_node_processors, nw->removed_node_processors);
                for (NodeID id : nw->nodes_to_remove) {
                    auto p = nodes.remove(id.idx);
                    if (p) old->removed_node_processors.emplace_back(id, std::move(*p));
                }
                ProcessorToContextMsg r; r.schedule = std::move(old);
                bool ok = ch->to_ctx.push(std::move(r)); assert(ok); (void)ok;
            }
            for (auto& np : nw->new_node_processors) { auto prev = nodes.insert_at(np.first.idx, std::move(np.second)); assert(!prev); (void)prev; }
            nw->new_node_processors.clear();
            schedule_data = std::move(nw);
        }
    }
    void process_block(size_t block_frames, double t, uint32_t status) {  // :208-248
        poll_messages();
        if (!running || !schedule_data) return;
        schedule_data->schedule.process(block_frames, [&](NodeID id, SilenceMask in_mask, const std::vector<const float*>& in, const std::vector<float*>& out) {
            SilenceMask out_mask = SilenceMask::none();
            ProcInfo pi{in_mask, &out_mask, t, status, user_cx};
            nodes.at(id.idx)->process(block_frames, in, out, pi);
            return out_mask;
        });
    }
};

enum class UpdateStatusKind : int { Inactive = 0, Active = 1, Deactivated = 2 };  // context.rs:245-254
struct UpdateStatus { UpdateStatusKind kind = UpdateStatusKind::Inactive; CompileErrorInfo graph_error; void* returned_user_cx = nullptr; };

class FirewheelGraphCtx {  // context.rs:29-243
  public:
    AudioGraph graph;
    explicit FirewheelGraphCtx(const AudioGraphConfig& c) : graph(c) {}
    ~FirewheelGraphCtx() { if (is_activated()) deactivate(true); }
    std::unique_ptr<FirewheelProcessor> activate(uint32_t sample_rate, size_t n_in, size_t n_out, size_t max_block_frames, void* user_cx) {  // :46-82
        assert(sample_rate != 0 && max_block_frames > 0);
        if (active_) return nullptr;
        ch_ = std::make_shared<Channels>(); active_ = true; sample_rate_ = sample_rate; max_block_frames_ = max_block_frames;
        return std::make_unique<FirewheelProcessor>(ch_, graph.current_node_capacity(), n_in, n_out, max_block_frames, user_cx);
    }
    bool is_activated() const { return active_; }
    uint32_t sample_rate() const { return active_ ? sample_rate_ : 0; }
    UpdateStatus update() {  // :93-148
        UpdateStatus st;
        graph.update();
        if (!active_) return st;
        bool dropped = false; void* cx = nullptr;
        update_internal(&dropped, &cx);
        if (dropped) { graph.deactivate(); active_ = false; ch_.reset(); st.kind = UpdateStatusKind::Deactivated; st.returned_user_cx = cx; return st; }
        st.kind = UpdateStatusKind::Active;
        if (graph.needs_compile()) {
            std::unique_ptr<ScheduleHeapData> sd;
            CompileErrorInfo err = graph.compile(sample_rate_, max_block_frames_, &sd);
            if (err.code != CompileGraphError::Ok) { st.graph_error = err; return st; }
            ContextToProcessorMsg m; m.new_schedule = std::move(sd);
            if (ch_->to_proc.q.size() >= ch_->to_proc.cap) graph.on_schedule_returned(std::move(m.new_schedule));  // :128-136
            else ch_->to_proc.push(std::move(m));
        }
        return st;
    }
    // The stream side is driven by the caller here: `pump` stands for "the audio thread keeps
    // calling process_interleaved until it sees Stop" (context.rs:162-211 polls for that).
    template <class Pump> void* deactivate_with(bool stream_is_running, Pump&& pump) {
        if (!active_) return nullptr;
        bool dropped = false; void* cx = nullptr;
        if (stream_is_running) { ContextToProcessorMsg m; m.stop = true; if (!ch_->to_proc.push(std::move(m))) dropped = true; }
        int spins = 0;
        while (!dropped) { pump(); update_internal(&dropped, &cx); if (!dropped && ++spins > 1500) dropped = true; }  // 3 s / 2 ms timeout
        graph.deactivate(); active_ = false; ch_.reset();
        return cx;
    }
    void* deactivate(bool stream_is_running) { return deactivate_with(stream_is_running, [] {}); }

  private:
    void update_internal(bool* dropped, void** cx) {  // :213-234
        if (!active_) return;
        ProcessorToContextMsg m;
        while (ch_->to_ctx.pop(&m)) {
            if (m.dropped) { graph.on_processor_dropped(m.nodes); *dropped = true; *cx = m.user_cx; }
            else graph.on_schedule_returned(std::move(m.schedule));
        }
    }
    bool active_ = false; std::shared_ptr<Channels> ch_; uint32_t sample_rate_ = 0; size_t max_block_frames_ = 0;
};

}  // namespace fwo
