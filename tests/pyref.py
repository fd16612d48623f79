"""A second, independent restatement of the reference's per-block path — pure Python with numpy.float32 scalars, written
straight from the Rust source (file:line cited), sharing no code with oracle/fw_oracle.hpp. It exists to cross-check the C++
oracle where the reference itself cannot be built (no Rust toolchain): small cases only, clarity over speed.

Covers: the executor loop and its silence flags (schedule.rs:213-343), ParamSmoother (smoother.rs:93-205), VolumeNode
(volume.rs:85-144), SumNode (sum.rs:42-135), MonoToStereo / StereoToMono (mono_to_stereo.rs:34-49, stereo_to_mono.rs:34-55),
HardClip (hard_clip.rs:52-94), clear_all_outputs (util.rs:165-175), SamplerNode with its message drain and the sample-resource
conversions (sampler.rs:283-560, sample_resource.rs:337-456). The schedule (node order, buffer indices, should_clear) is
taken from the library under test through `compile_internal` — the compiler has its own equivalence test."""
import math

import numpy as np

f32 = np.float32
ZERO = f32(0.0)


def all_silent(n):
    return (1 << n) - 1 if n < 64 else (1 << 64) - 1


class Smoother:  # smoother.rs:93-205
    def __init__(self, val, sample_rate, max_block_frames, smooth_secs=f32(10.0 / 1000.0), eps=f32(0.00001)):
        self.b = f32(math.exp(float(f32(-1.0) / (f32(smooth_secs) * f32(sample_rate)))))  # :99  (-1.0 / (secs * sr)).exp()
        self.a = f32(1.0) - self.b                                                          # :100
        self.status = "inactive"
        self.input = f32(val)
        self.output = np.full(max_block_frames, f32(val), f32)
        self.last = f32(val)
        self.eps = f32(eps)

    def reset(self, val):  # :115-129
        val = f32(val)
        if self.status != "inactive":
            self.status = "inactive"; self.input = val; self.last = val; self.output[:] = val
        elif self.input != val:
            self.input = val; self.last = val; self.output[:] = val

    def set(self, val):  # :133-140
        val = f32(val)
        if self.input == val:
            return
        self.input = val
        self.status = "active"

    def process(self, frames):  # :159-194 — returns (values, smoothing)
        frames = min(frames, len(self.output))
        if self.status != "active" or frames == 0:
            return self.output, self.status != "inactive"  # Q1: the whole buffer
        t = self.input * self.a
        self.output[0] = t + (self.last * self.b)
        for i in range(1, frames):
            self.output[i] = t + (self.output[i - 1] * self.b)
        self.last = self.output[frames - 1]
        if abs(self.input - self.output[0]) < self.eps:  # Q3
            self.reset(self.input)
            self.status = "deactivating"  # Q2: never leaves this state again
        return self.output[:frames], self.status != "inactive"

    def set_and_process(self, val, frames):  # :202-205
        self.set(val)
        return self.process(frames)


def clear_all_outputs(frames, outputs):  # util.rs:165-175
    for o in outputs:
        o[:frames] = ZERO
    return all_silent(len(outputs))


def percent_to_raw_gain(p):  # range.rs:32-35
    n = max(f32(p), ZERO) * (f32(1.0) / f32(100.0))
    return f32(n * n)


class Volume:  # volume.rs:56-144
    def __init__(self, percent, sr, mbf):
        self.raw_gain = percent_to_raw_gain(max(f32(percent), ZERO))
        self.sm = Smoother(self.raw_gain, sr, mbf)

    def set_percent(self, p):
        self.raw_gain = percent_to_raw_gain(p)

    def process(self, frames, inputs, outputs, in_mask):
        g = self.raw_gain
        if in_mask & all_silent(len(inputs)) == all_silent(len(inputs)):  # :94-100
            self.sm.reset(g)
            return clear_all_outputs(frames, outputs)
        gain, smoothing = self.sm.set_and_process(g, frames)
        if not smoothing and gain[0] < f32(0.00001):  # :104-108
            return clear_all_outputs(frames, outputs)
        if len(inputs) == 2 and len(outputs) == 2:  # :116-129, Q7
            for i in range(frames):
                outputs[0][i] = inputs[0][i] * gain[i]
                outputs[1][i] = inputs[1][i] * gain[i]
            return in_mask
        for c, (o, x) in enumerate(zip(outputs, inputs)):  # :131-143
            if (in_mask >> c) & 1:
                o[:frames] = ZERO
                continue
            for i in range(frames):
                o[i] = x[i] * gain[i]
        return in_mask  # :110


class Sum:  # sum.rs:20-135
    def __init__(self, n_in, n_out):
        self.ports = n_in // n_out

    def process(self, frames, inputs, outputs, in_mask):
        ni, no = len(inputs), len(outputs)
        if in_mask & all_silent(ni) == all_silent(ni):  # :52-56
            return clear_all_outputs(frames, outputs)
        if ni == no:  # :58-65
            for o, x in zip(outputs, inputs):
                o[:frames] = x[:frames]
            return in_mask
        if self.ports in (2, 3, 4):  # :69-110 — left to right, mask untouched (NONE_SILENT)
            for ch, o in enumerate(outputs):
                for i in range(frames):
                    acc = inputs[ch][i]
                    for p in range(1, self.ports):
                        acc = acc + inputs[no * p + ch][i]
                    o[i] = acc
            return 0
        for ch, o in enumerate(outputs):  # :111-133
            o[:frames] = inputs[ch][:frames]
            for p in range(1, self.ports):
                idx = no * p + ch
                if (in_mask >> idx) & 1:
                    continue
                for i in range(frames):
                    o[i] = o[i] + inputs[idx][i]
        return 0


class MonoToStereo:  # mono_to_stereo.rs:34-49
    def process(self, frames, inputs, outputs, in_mask):
        if in_mask & 1:
            return clear_all_outputs(frames, outputs)
        outputs[0][:frames] = inputs[0][:frames]
        outputs[1][:frames] = inputs[0][:frames]
        return 0


class StereoToMono:  # stereo_to_mono.rs:34-55
    def process(self, frames, inputs, outputs, in_mask):
        if in_mask & 3 == 3 or len(inputs) < 2 or not outputs:
            return clear_all_outputs(frames, outputs)
        for i in range(frames):
            outputs[0][i] = (inputs[0][i] + inputs[1][i]) * f32(0.5)
        return 0


def db_to_gain_clamped(db):  # util.rs:21-27: if db <= -100 { 0 } else { 10^(0.05 * db) }, in f32
    db = f32(db)
    return ZERO if db <= f32(-100.0) else f32(np.power(f32(10.0), f32(0.05) * db))


class HardClip:  # hard_clip.rs:8-94
    def __init__(self, threshold_gain):
        self.t = f32(threshold_gain)

    def process(self, frames, inputs, outputs, in_mask):
        if len(inputs) == 2 and len(outputs) == 2 and in_mask & 3 == 0:  # :60-80, leaves the mask NONE_SILENT
            for c in range(2):
                for i in range(frames):
                    outputs[c][i] = max(min(inputs[c][i], self.t), -self.t)
            return 0
        for c, (o, x) in enumerate(zip(outputs, inputs)):  # :82-91
            if (in_mask >> c) & 1:
                o[:frames] = ZERO
                continue
            for i in range(frames):
                o[i] = max(min(x[i], self.t), -self.t)
        return in_mask  # :93


class Dummy:  # dummy.rs:34-41 — graph_in / graph_out: leaves outputs alone, mask NONE_SILENT
    def process(self, frames, inputs, outputs, in_mask):
        return 0


class Executor:
    """CompiledSchedule (schedule.rs:166-343) + the block loop of process_interleaved (processor.rs:93-150), planar I/O."""

    def __init__(self, schedule, num_buffers, max_block_frames, processors):
        self.sched, self.mbf, self.procs = schedule, max_block_frames, processors  # schedule: [(node_key, [(buf, should_clear)], [buf])]
        self.buffers = np.zeros((num_buffers, max_block_frames), f32)
        self.flags = [False] * num_buffers

    def process(self, x, n_out):  # x: [n_in][T] -> ([n_out][T], out mask of the last block)
        n_in, T = x.shape
        y = np.zeros((n_out, T), f32)
        done, mask = 0, 0
        while done < T:
            frames = min(T - done, self.mbf)
            # prepare_graph_inputs (schedule.rs:213-253); Q4: these flags are overwritten by graph_in's NONE_SILENT below
            gin_outs = self.sched[0][2]
            fill = min(n_in, len(gin_outs))
            for i in range(fill):
                self.buffers[gin_outs[i], :frames] = x[i, done:done + frames]
                self.flags[gin_outs[i]] = False
            for b in gin_outs[fill:]:
                self.buffers[b, :frames] = ZERO
                self.flags[b] = True
            # CompiledSchedule::process (schedule.rs:289-343)
            for key, ins, outs in self.sched:
                in_mask = 0
                for i, (b, clear) in enumerate(ins):
                    if clear:
                        self.buffers[b, :frames] = ZERO
                        self.flags[b] = True
                    if self.flags[b]:
                        in_mask |= 1 << i
                out_mask = self.procs[key].process(frames, [self.buffers[b] for b, _ in ins], [self.buffers[b] for b in outs], in_mask)
                for i, b in enumerate(outs):
                    self.flags[b] = bool((out_mask >> i) & 1)
            # read_graph_outputs (schedule.rs:255-287)
            gout_ins = self.sched[-1][1]
            mask = 0
            for i in range(min(n_out, len(gout_ins))):
                b = gout_ins[i][0]
                if self.flags[b]:
                    mask |= 1 << i
                y[i, done:done + frames] = self.buffers[b, :frames]
            done += frames
        return y, mask

    def process_interleaved(self, x, n_out):  # processor.rs:93-150 with util.rs:43-147; x: [T][n_in] -> [T][n_out]
        T, n_in = x.shape
        y = np.zeros((T, n_out), f32)
        done = 0
        while done < T:
            frames = min(T - done, self.mbf)
            gin_outs = self.sched[0][2]
            fill = min(n_in, len(gin_outs))
            for i in range(fill):  # deinterleave (util.rs:43-87); its stale-buffer silence mask (Q4) is overwritten by graph_in below
                self.buffers[gin_outs[i], :frames] = x[done:done + frames, i]
            for b in gin_outs[fill:]:
                self.buffers[b, :frames] = ZERO
                self.flags[b] = True
            for key, ins, outs in self.sched:
                in_mask = 0
                for i, (b, clear) in enumerate(ins):
                    if clear:
                        self.buffers[b, :frames] = ZERO
                        self.flags[b] = True
                    if self.flags[b]:
                        in_mask |= 1 << i
                out_mask = self.procs[key].process(frames, [self.buffers[b] for b, _ in ins], [self.buffers[b] for b in outs], in_mask)
                for i, b in enumerate(outs):
                    self.flags[b] = bool((out_mask >> i) & 1)
            gout_ins = self.sched[-1][1]
            chans = [gout_ins[i][0] for i in range(min(n_out, len(gout_ins)))]
            mask = sum(1 << i for i, b in enumerate(chans) if self.flags[b])
            blk = y[done:done + frames]
            if len(chans) == 2 and n_out == 2:  # interleave_stereo (util.rs:123-147): zeros only if BOTH channels are flagged
                if mask & 3 != 3:
                    blk[:, 0] = self.buffers[chans[0], :frames]; blk[:, 1] = self.buffers[chans[1], :frames]
            else:  # interleave (util.rs:90-120): zero-fill, then every channel that is not flagged
                for i, b in enumerate(chans):
                    if not (mask >> i) & 1:
                        blk[:, i] = self.buffers[b, :frames]
            done += frames
        return y


# ---- SamplerNode (sampler.rs:283-560) + sample resources (sample_resource.rs:337-456) --------------------------------------
def pcm_i16_to_f32(s):  # sample_resource.rs:337-340
    return f32(s) * (f32(1.0) / f32(32767.0))


def pcm_u16_to_f32(s):  # sample_resource.rs:342-345
    return (f32(s) * (f32(2.0) / f32(65535.0))) - f32(1.0)


class Resource:
    """data: [channels][frames] float32 / int16 / uint16 (layout does not matter to the restatement: fill_buffers reads
    channel c, frame f whichever way the resource stores it, sample_resource.rs:348-456)."""

    def __init__(self, data):
        self.data = np.asarray(data)
        self.channels, self.frames = self.data.shape
        self.conv = {np.dtype(np.float32): f32, np.dtype(np.int16): pcm_i16_to_f32, np.dtype(np.uint16): pcm_u16_to_f32}[self.data.dtype]

    def fill_buffers(self, buffers, r0, r1, start_frame):
        for c in range(min(len(buffers), self.channels)):
            for i in range(r0, r1):
                f = start_frame + (i - r0)
                buffers[c][i] = self.conv(self.data[c, f]) if f < self.frames else ZERO  # past the end: the reference panics; defined as 0.0


def secs_to_frame(secs, sr):  # `(secs * sr as f64).round() as u64`, saturating (sampler.rs:250-251, 394)
    f = math.floor(abs(secs * sr) + 0.5) * (1 if secs >= 0 else -1)  # f64::round: half away from zero
    return 0 if not f > 0 else int(f)


class Sampler:
    def __init__(self, percent, sr, mbf):
        self.raw_gain = percent_to_raw_gain(max(f32(percent), ZERO))
        self.sm = Smoother(self.raw_gain, sr, mbf)
        self.sr, self.playing, self.playhead, self.loop, self.sample, self.msgs = sr, False, 0, None, None, []

    def set_percent(self, p):
        self.raw_gain = percent_to_raw_gain(p)

    def loop_start_or_zero(self):
        return self.loop[0] if self.loop else 0

    def process(self, frames, inputs, outputs, in_mask):
        for m in self.msgs:  # sampler.rs:331-414
            kind = m[0]
            if kind == "set_sample":
                self.sample = m[1]
                if self.loop and self.loop[2]:
                    self.loop = [0, self.sample.frames, True]
                if m[2]:
                    self.playhead = self.loop_start_or_zero(); self.playing = False
            elif kind == "play":
                self.playing = True
            elif kind == "pause":
                self.playing = False
            elif kind == "stop":
                self.playhead = self.loop_start_or_zero(); self.playing = False
            elif kind == "playhead":
                self.playhead = secs_to_frame(m[1], self.sr)
            else:  # ("loop", None | "full" | (start, end))
                if m[1] is None:
                    self.loop = None
                else:
                    self.loop = [0, self.sample.frames if self.sample else 0, True] if m[1] == "full" else [secs_to_frame(m[1][0], self.sr), secs_to_frame(m[1][1], self.sr), False]
                    if self.loop[0] <= self.playhead < self.loop[1]:
                        self.playhead = self.loop[0]
        self.msgs = []
        if self.sample is None or not self.playing:  # :416-430
            return clear_all_outputs(frames, outputs)
        gain, smoothing = self.sm.set_and_process(self.raw_gain, frames)
        if not smoothing and gain[0] < f32(0.00001):  # :437-443
            return clear_all_outputs(frames, outputs)
        smp = self.sample
        if self.loop:  # :445-484
            if self.playhead >= self.loop[1]:
                self.playhead = self.loop[0]
            first = min(frames, self.loop[1] - self.playhead)
            smp.fill_buffers(outputs, 0, first, self.playhead)
            if first < frames:
                self.playhead = self.loop[0]
                smp.fill_buffers(outputs, first, frames, self.playhead)
                self.playhead += frames - first
            else:
                self.playhead += frames
        else:  # :485-516
            if self.playhead >= smp.frames:
                self.playing = False
                return clear_all_outputs(frames, outputs)
            copy = min(frames, smp.frames - self.playhead)
            smp.fill_buffers(outputs, 0, copy, self.playhead)
            if copy < frames:
                self.playing = False; self.playhead = 0
                for o in outputs:
                    o[copy:frames] = ZERO
            else:
                self.playhead += frames
        sch = smp.channels
        for c in range(min(len(outputs), sch)):  # :522-543 (the stereo loop is the same arithmetic)
            for i in range(frames):
                outputs[c][i] = outputs[c][i] * gain[i]
        mask = 0
        if len(outputs) > sch:  # :545-559
            if len(outputs) == 2 and sch == 1:
                outputs[1][:frames] = outputs[0][:frames]
            else:
                for c in range(sch, len(outputs)):
                    outputs[c][:frames] = ZERO
                    mask |= 1 << c
        return mask
