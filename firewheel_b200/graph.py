"""Host-side mirror of the reference's graph / context / processor API over the C ABI.

Names, argument meaning and error behaviour follow firewheel-graph so that tests read
like the reference's own (crates/firewheel-graph/src/graph/compiler/schedule.rs:392-711):

    cx = FirewheelGraphCtx(lib, AudioGraphConfig(num_graph_inputs=2, num_graph_outputs=2))
    vol = cx.graph.add_node(2, 2, VolumeNode(50.0))
    cx.graph.connect(cx.graph.graph_in_node(), 0, vol, 0, False)
    ...
    proc = cx.activate(48000, 2, 2, 256)
    cx.update()
    proc.process_interleaved(inp, out, 2, 2, frames, 0.0, 0)

Every class is a thin handle; all state and all arithmetic live behind the C ABI
(`lib` is a `_capi.Lib`: the CUDA product, or the CPU oracle in tests).
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _capi as K


# ---- ids -------------------------------------------------------------------------------------
class NodeID(int):
    """thunderdome index: slot | generation << 32 (graph.rs:20-23)."""

    @property
    def slot(self):
        return int(self) & 0xFFFFFFFF

    @property
    def generation(self):
        return int(self) >> 32

    def __repr__(self):
        return f"NodeID({self.slot}-{self.generation})"


class EdgeID(NodeID):
    def __repr__(self):
        return f"EdgeID({self.slot}-{self.generation})"


class AddEdgeError(Exception):
    """graph/error.rs:14-37. `kind` is the variant name; node/port carry its payload."""

    def __init__(self, code, node=None, port=None):
        self.code, self.kind, self.node, self.port = code, K.ADD_EDGE_ERRORS[code], node, port
        super().__init__(f"Could not add edge: {self.kind} node={node} port={port}")


class CompileGraphError(Exception):
    """graph/error.rs:101-116."""

    def __init__(self, code, node=None, port=None, message=""):
        self.code, self.kind, self.node, self.port = code, K.COMPILE_ERRORS.get(code, str(code)), node, port
        super().__init__(f"Failed to compile audio graph: {self.kind} {message}")


# ---- nodes (values handed to add_node, like `impl Into<Box<dyn AudioNode>>`) -----------------
@dataclass
class _Node:
    kind: int = K.NODE_DUMMY
    u: tuple = (0, 0, 0)
    f: tuple = (0.0, 0.0, 0.0, 0.0)
    data: np.ndarray = None

    def desc(self):
        d = K.NodeDesc(kind=self.kind, u0=self.u[0], u1=self.u[1], u2=self.u[2],
                       f0=self.f[0], f1=self.f[1], f2=self.f[2], f3=self.f[3])
        if self.data is not None:
            self._keep = np.ascontiguousarray(self.data, dtype=np.float32)
            d.data = self._keep.ctypes.data_as(C.POINTER(C.c_float))
            d.data_len = self._keep.size
        return d


def DummyAudioNode():
    return _Node(K.NODE_DUMMY)


def VolumeNode(percent_volume):  # volume.rs:16
    return _Node(K.NODE_VOLUME, f=(float(percent_volume), 0.0, 0.0, 0.0))


def SumNode():
    return _Node(K.NODE_SUM)


def MonoToStereoNode():
    return _Node(K.NODE_MONO_TO_STEREO)


def StereoToMonoNode():
    return _Node(K.NODE_STEREO_TO_MONO)


def HardClipNode(threshold_db):  # hard_clip.rs:8
    return _Node(K.NODE_HARD_CLIP, f=(float(threshold_db), 0.0, 0.0, 0.0))


def PanNode(pan):
    return _Node(K.NODE_PAN, f=(float(pan), 0.0, 0.0, 0.0))


def BiquadNode(num_stages):
    return _Node(K.NODE_BIQUAD, u=(int(num_stages), 0, 0))


def DelayNode(delay_frames):
    return _Node(K.NODE_DELAY, u=(int(delay_frames), 0, 0))


def ConvReverbNode(ir):
    ir = np.atleast_2d(np.asarray(ir, dtype=np.float32))
    return _Node(K.NODE_CONV_REVERB, u=(ir.shape[1], ir.shape[0], 0), data=ir)


def SvfNode(num_stages):
    return _Node(K.NODE_SVF, u=(int(num_stages), 0, 0))


def ResamplerNode(table):
    """table: [phases][taps] float32 (see design_resampler)."""
    table = np.ascontiguousarray(table, dtype=np.float32)
    return _Node(K.NODE_RESAMPLER, u=(table.shape[0], table.shape[1], 0), data=table)


def SamplerNode(percent_volume):  # sampler.rs:56
    return _Node(K.NODE_SAMPLER, f=(float(percent_volume), 0.0, 0.0, 0.0))


class SamplerError(Exception):
    """Err(()) of the SamplerNode setters (sampler.rs:67-169), with the reason the C ABI reports."""

    def __init__(self, code):
        self.code, self.kind = code, K.SAMPLER_ERRORS.get(code, str(code))
        super().__init__(self.kind)


@dataclass
class AudioGraphConfig:  # graph.rs:91-107 + batching
    num_graph_inputs: int = 0
    num_graph_outputs: int = 2
    initial_node_capacity: int = 64
    initial_edge_capacity: int = 256
    num_voices: int = 1
    master_bus: bool = False
    device: int = 0
    max_call_frames: int = 0  # product: device memory is reserved for stretches of this many frames (0 = 64 blocks); longer calls are chunked


@dataclass
class ScheduledNode:  # schedule.rs:13-30
    id: NodeID
    input_buffers: list = field(default_factory=list)   # [(buffer_index, should_clear)]
    output_buffers: list = field(default_factory=list)  # [buffer_index]


@dataclass
class VoiceTemplate:  # include/fw_b200.h fw_voice_template
    num_voices: int
    num_template_nodes: int
    voice_inputs: int
    voice_outputs: int
    num_tree_nodes: int


@dataclass
class UpdateStatus:  # context.rs:245-254
    kind: str
    graph_error: CompileGraphError = None
    returned_user_cx: int = None


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"], "f32 C-contiguous buffers only"
        return a.ctypes.data
    return int(a)  # raw (device or pinned) address


class AudioGraph:
    def __init__(self, lib, ctx):
        self._lib, self._ctx = lib, ctx

    def graph_in_node(self):
        return NodeID(self._lib.graph_in_node(self._ctx))

    def graph_out_node(self):
        return NodeID(self._lib.graph_out_node(self._ctx))

    def add_node(self, num_inputs, num_outputs, node):
        d = node.desc()
        nid = self._lib.graph_add_node(self._ctx, num_inputs, num_outputs, C.byref(d))
        if nid == K.FW_ID_DANGLING:
            raise ValueError(self._lib.ctx_last_error(self._ctx).decode())
        return NodeID(nid)

    def add_custom_node(self, num_inputs, num_outputs, vtable_ptr, node_ptr):
        """add_node with a user node behind the plugin vtable (include/fw_b200.h fw_node_vtable): `vtable_ptr` / `node_ptr` are
        raw addresses produced by the plugin library; the graph takes the node over."""
        nid = self._lib.graph_add_custom_node(self._ctx, num_inputs, num_outputs, vtable_ptr, node_ptr)
        if nid == K.FW_ID_DANGLING:
            raise ValueError("bad custom node")
        return NodeID(nid)

    def _removed(self, fn, *args):
        cap = 4096
        buf = (C.c_uint64 * cap)()
        n = C.c_uint32(0)
        rc = fn(self._ctx, *args, buf, cap, C.byref(n))
        if rc != 0:
            raise KeyError("Err(())")
        return [EdgeID(buf[i]) for i in range(min(n.value, cap))]

    def remove_node(self, node_id):
        return self._removed(self._lib.graph_remove_node, node_id)

    def set_num_inputs(self, node_id, n):
        return self._removed(self._lib.graph_set_num_inputs, node_id, n)

    def set_num_outputs(self, node_id, n):
        return self._removed(self._lib.graph_set_num_outputs, node_id, n)

    def connect(self, src_node, src_port, dst_node, dst_port, check_for_cycles):
        e, en, ep = C.c_uint64(0), C.c_uint64(0), C.c_uint32(0)
        rc = self._lib.graph_connect(self._ctx, src_node, src_port, dst_node, dst_port, int(check_for_cycles),
                                     C.byref(e), C.byref(en), C.byref(ep))
        if rc != 0:
            raise AddEdgeError(rc, NodeID(en.value), ep.value)
        return EdgeID(e.value)

    def disconnect(self, src_node, src_port, dst_node, dst_port):
        return bool(self._lib.graph_disconnect(self._ctx, src_node, src_port, dst_node, dst_port))

    def disconnect_by_edge_id(self, edge_id):
        return bool(self._lib.graph_disconnect_by_edge_id(self._ctx, edge_id))

    def edge(self, edge_id):
        info = K.EdgeInfoC()
        if not self._lib.graph_edge(self._ctx, edge_id, C.byref(info)):
            return None
        return info

    def node_info(self, node_id):
        info = K.NodeInfoC()
        if not self._lib.graph_node_info(self._ctx, node_id, C.byref(info)):
            return None
        return info

    def nodes(self):
        n = self._lib.graph_num_nodes(self._ctx)
        buf = (C.c_uint64 * max(n, 1))()
        self._lib.graph_nodes(self._ctx, buf, n)
        return [NodeID(buf[i]) for i in range(n)]

    def edges(self):
        n = self._lib.graph_num_edges(self._ctx)
        buf = (C.c_uint64 * max(n, 1))()
        self._lib.graph_edges(self._ctx, buf, n)
        return [EdgeID(buf[i]) for i in range(n)]

    def cycle_detected(self):
        return bool(self._lib.graph_cycle_detected(self._ctx))

    def reset(self):
        self._lib.graph_reset(self._ctx)

    def needs_compile(self):
        return bool(self._lib.graph_needs_compile(self._ctx))

    def compile_internal(self, max_block_frames):
        """graph.rs:629 — returns the schedule (list of ScheduledNode) + num_buffers, or raises."""
        rc = self._lib.graph_compile_internal(self._ctx, max_block_frames)
        if rc != 0:
            raise CompileGraphError(rc)
        out = []
        for i in range(self._lib.schedule_len(self._ctx)):
            sn = K.ScheduledNodeC()
            assert self._lib.schedule_node(self._ctx, i, C.byref(sn))
            out.append(ScheduledNode(NodeID(sn.id),
                                     [(sn.in_buffer[k], bool(sn.in_should_clear[k])) for k in range(sn.num_inputs)],
                                     [sn.out_buffer[k] for k in range(sn.num_outputs)]))
        return out, self._lib.schedule_num_buffers(self._ctx)

    # ---- isomorphic-voice detection (ours; include/fw_b200.h graph_detect_voices) ------------------
    def detect_voices(self):
        """Recognise V isomorphic voices under a balanced SumNode tree in this flat graph; returns VoiceTemplate or raises ValueError with
        the reason. The batched equivalent is FirewheelGraphCtx.new_batched(flat_ctx)."""
        t = K.VoiceTemplateC()
        if self._lib.graph_detect_voices(self._ctx, C.byref(t)) != 0:
            raise ValueError(self._lib.ctx_last_error(self._ctx).decode())
        return VoiceTemplate(t.num_voices, t.num_template_nodes, t.voice_inputs, t.voice_outputs, t.num_tree_nodes)

    def voice_nodes(self, template_node):
        """ids, in this flat graph, of template node `template_node` in voice 0 .. V-1 (after detect_voices)"""
        n = self._lib.graph_voice_nodes(self._ctx, template_node, None, 0)
        buf = (C.c_uint64 * max(n, 1))()
        self._lib.graph_voice_nodes(self._ctx, template_node, buf, n)
        return [NodeID(buf[i]) for i in range(n)]

    def read_params(self, node_id, which):
        """main-thread view of a node's parameter table (K.FW_PARAM_*), as float32 array; empty if the node has no such table"""
        n = self._lib.node_read_params(self._ctx, node_id, which, None, 0)
        out = np.zeros(n, dtype=np.float32)
        if n:
            self._lib.node_read_params(self._ctx, node_id, which, out.ctypes.data, n)
        return out

    # ---- parameters -------------------------------------------------------------------------
    def set_event_block(self, block):
        """Stamp the parameter stores and sampler / resampler messages that follow with a block offset into the next process call
        (0 = at its start); see include/fw_b200.h ctx_set_event_block."""
        self._lib.ctx_set_event_block(self._ctx, int(block))

    def _chk(self, rc, what):
        if rc != 0:
            raise ValueError(f"{what}: node is not of that kind / bad voice index")

    def set_percent_volume(self, node_id, percent, voice=K.FW_ALL_VOICES):  # volume.rs:28
        if np.ndim(percent) == 0:
            self._chk(self._lib.volume_set_percent_volume(self._ctx, node_id, voice, float(percent)), "volume")
        else:
            a = np.ascontiguousarray(percent, dtype=np.float32)
            self._chk(self._lib.volume_set_percent_volumes(self._ctx, node_id, a.ctypes.data, a.size), "volume")

    def set_pan(self, node_id, pan, voice=K.FW_ALL_VOICES):
        if np.ndim(pan) == 0:
            self._chk(self._lib.pan_set_pan(self._ctx, node_id, voice, float(pan)), "pan")
        else:
            a = np.ascontiguousarray(pan, dtype=np.float32)
            self._chk(self._lib.pan_set_pans(self._ctx, node_id, a.ctypes.data, a.size), "pan")

    def set_pan_gains(self, node_id, gl, gr, voice=K.FW_ALL_VOICES):
        self._chk(self._lib.pan_set_gains(self._ctx, node_id, voice, float(gl), float(gr)), "pan")

    def set_biquad_coeffs(self, node_id, coeffs, voice=K.FW_ALL_VOICES, stage=None):
        a = np.ascontiguousarray(coeffs, dtype=np.float32)
        if a.ndim == 1:
            self._chk(self._lib.biquad_set_coeffs(self._ctx, node_id, voice, stage, a.ctypes.data), "biquad")
        elif a.ndim == 2:  # [stage][5], same for the selected voices
            for s in range(a.shape[0]):
                self._chk(self._lib.biquad_set_coeffs(self._ctx, node_id, voice, s, a[s].ctypes.data), "biquad")
        else:  # [voice][stage][5]
            self._chk(self._lib.biquad_set_all_coeffs(self._ctx, node_id, a.ctypes.data, a.shape[0], a.shape[1]), "biquad")


    def set_svf_coeffs(self, node_id, coeffs, voice=K.FW_ALL_VOICES):
        """coeffs: [stage][6] for the selected voices, or [voice][stage][6]; rows {a1, a2, a3, m0, m1, m2}."""
        a = np.ascontiguousarray(coeffs, dtype=np.float32)
        if a.ndim == 2:
            for s in range(a.shape[0]):
                self._chk(self._lib.svf_set_coeffs(self._ctx, node_id, voice, s, a[s].ctypes.data), "svf")
        else:
            self._chk(self._lib.svf_set_all_coeffs(self._ctx, node_id, a.ctypes.data, a.shape[0], a.shape[1]), "svf")

    def resampler_set(self, node_id, resource, ratio=None, step_q32=None, playing=True, loop=False, voice=K.FW_ALL_VOICES):
        """ratio: source frames advanced per output frame (Q32.32 on the wire)."""
        step = int(step_q32) if step_q32 is not None else int(round(float(ratio) * 4294967296.0))
        self._chk(self._lib.resampler_set(self._ctx, node_id, voice, int(resource), step, int(bool(playing)), int(bool(loop))), "resampler")

    def resampler_seek(self, node_id, pos_frames, voice=K.FW_ALL_VOICES):
        self._chk(self._lib.resampler_seek(self._ctx, node_id, voice, int(pos_frames)), "resampler")

    # ---- sample resources + SamplerNode (sample_resource.rs, sampler.rs:46-181) ----
    _FORMATS = {("float32", False): K.SAMPLE_F32_PLANAR, ("float32", True): K.SAMPLE_F32_INTERLEAVED,
                ("int16", True): K.SAMPLE_I16_INTERLEAVED, ("uint16", True): K.SAMPLE_U16_INTERLEAVED,
                ("int16", False): K.SAMPLE_I16_PLANAR, ("uint16", False): K.SAMPLE_U16_PLANAR}

    def create_sample_resource(self, data, interleaved=False):
        """data: [frames][channels] when interleaved, else [channels][frames]; dtype float32 / int16 / uint16."""
        a = np.ascontiguousarray(data)
        if a.ndim == 1:
            a = a[:, None] if interleaved else a[None, :]
        fmt = self._FORMATS[(a.dtype.name, bool(interleaved))]
        frames, channels = (a.shape[0], a.shape[1]) if interleaved else (a.shape[1], a.shape[0])
        h = self._lib.sample_resource_create(self._ctx, fmt, channels, frames, a.ctypes.data)
        if h == 0:
            raise ValueError("bad sample resource")
        return h

    def _smp(self, rc):
        if rc != 0:
            raise SamplerError(rc)

    def sampler_set_sample(self, node_id, resource, stop_playback, voice=K.FW_ALL_VOICES):  # sampler.rs:67
        self._smp(self._lib.sampler_set_sample(self._ctx, node_id, voice, resource, int(bool(stop_playback))))

    def sampler_play(self, node_id, voice=K.FW_ALL_VOICES):  # sampler.rs:82
        self._smp(self._lib.sampler_play(self._ctx, node_id, voice))

    def sampler_pause(self, node_id, voice=K.FW_ALL_VOICES):  # sampler.rs:101
        self._smp(self._lib.sampler_pause(self._ctx, node_id, voice))

    def sampler_stop(self, node_id, voice=K.FW_ALL_VOICES):  # sampler.rs:120
        self._smp(self._lib.sampler_stop(self._ctx, node_id, voice))

    def sampler_set_playhead(self, node_id, playhead_secs, voice=K.FW_ALL_VOICES):  # sampler.rs:139
        self._smp(self._lib.sampler_set_playhead(self._ctx, node_id, voice, float(playhead_secs)))

    def sampler_set_loop_range(self, node_id, loop_range, voice=K.FW_ALL_VOICES):  # sampler.rs:153
        """loop_range: None | "full" | (start_secs, end_secs)  — Option<LoopRange> (sampler.rs:16-19)."""
        if loop_range is None:
            mode, s, e = K.LOOP_NONE, 0.0, 0.0
        elif isinstance(loop_range, str):
            mode, s, e = K.LOOP_FULL, 0.0, 0.0
        else:
            mode, (s, e) = K.LOOP_RANGE_SECS, loop_range
        self._smp(self._lib.sampler_set_loop_range(self._ctx, node_id, voice, mode, float(s), float(e)))

    def sampler_set_percent_volume(self, node_id, percent, voice=K.FW_ALL_VOICES):  # sampler.rs:174
        if np.ndim(percent) == 0:
            self._smp(self._lib.sampler_set_percent_volume(self._ctx, node_id, voice, float(percent)))
        else:
            for v, pc in enumerate(np.asarray(percent, dtype=np.float32)):
                self._smp(self._lib.sampler_set_percent_volume(self._ctx, node_id, v, float(pc)))

    def sampler_is_playing(self, node_id, voice=0):  # sampler.rs:163
        return self._lib.sampler_is_playing(self._ctx, node_id, voice) == 1


def design_svf(lib, ftype, fc, q, sample_rate):
    out = np.zeros(6, dtype=np.float32)
    lib.svf_design(int(ftype), float(fc), float(q), float(sample_rate), out.ctypes.data)
    return out


def design_resampler(lib, phases=256, taps=32, cutoff=1.0, beta=9.0):
    out = np.zeros((phases, taps), dtype=np.float32)
    lib.resampler_design(int(phases), int(taps), float(cutoff), float(beta), out.ctypes.data)
    return out


def design_rbj(lib, ftype, fc, q, gain_db, sample_rate):
    out = np.zeros(5, dtype=np.float32)
    lib.biquad_design_rbj(int(ftype), float(fc), float(q), float(gain_db), float(sample_rate), out.ctypes.data)
    return out


class Stream:
    """Pull-style backend (the role of firewheel-cpal's DataCallback, crates/firewheel-cpal/src/lib.rs:378-449)."""

    def __init__(self, lib, handle, n_out):
        self._lib, self._h, self.n_out = lib, handle, n_out

    def frames_ready(self):
        return self._lib.stream_frames_ready(self._h)

    def pull(self, frames):
        """-> (interleaved [frames][n_out] float32, frames delivered, status bits, stream_time_secs)"""
        out = np.full((frames, self.n_out), np.nan, dtype=np.float32)
        st, t = C.c_uint32(0), C.c_double(0.0)
        n = self._lib.stream_pull(self._h, out.ctypes.data, frames, C.byref(st), C.byref(t))
        return out, n, st.value, t.value

    def close(self):
        if self._h:
            self._lib.stream_close(self._h)
            self._h = None


class FirewheelProcessor:
    """processor.rs:18 — owned by the stream side; `free()` is Drop."""

    def __init__(self, lib, handle, cfg):
        self._lib, self._h, self.cfg = lib, handle, cfg

    def process_interleaved(self, input, output, num_in_channels, num_out_channels, frames, stream_time_secs=0.0, stream_status=0):
        return self._lib.processor_process_interleaved(self._h, _ptr(input), _ptr(output), num_in_channels, num_out_channels,
                                                       frames, stream_time_secs, stream_status)

    def process_planar(self, input, output, num_in_channels, num_out_channels, frames, stream_time_secs=0.0, stream_status=0):
        m = C.c_uint64(0)
        rc = self._lib.processor_process_planar(self._h, _ptr(input), _ptr(output), num_in_channels, num_out_channels,
                                                frames, stream_time_secs, stream_status, C.byref(m))
        return rc, m.value

    def process_planar_device(self, d_input, d_output, num_in_channels, num_out_channels, frames, stream_time_secs=0.0, stream_status=0):
        return self._lib.processor_process_planar_device(self._h, _ptr(d_input), _ptr(d_output), num_in_channels,
                                                         num_out_channels, frames, stream_time_secs, stream_status)

    def h2d(self, dst, src, nbytes):
        return self._lib.processor_h2d(self._h, _ptr(dst), _ptr(src), nbytes)

    def d2h(self, dst, src, nbytes):
        return self._lib.processor_d2h(self._h, _ptr(dst), _ptr(src), nbytes)

    def sync(self):
        return self._lib.processor_sync(self._h)

    def event_record(self, slot):
        return self._lib.processor_event_record(self._h, slot)

    def event_elapsed_ms(self, a, b):
        return self._lib.processor_event_elapsed_ms(self._h, a, b)

    def kernel_launches(self):
        return self._lib.processor_kernel_launches(self._h)

    def graph_replays(self):
        return int(self._lib.processor_graph_replays(self._h))

    def profile(self, enable):
        return self._lib.processor_profile(self._h, int(enable))

    def profile_read(self):
        """-> (ms[4], launches[4]) for classes control / chain / combine / temporal, then resets."""
        ms = (C.c_double * 4)()
        n = (C.c_uint64 * 4)()
        rc = self._lib.processor_profile_read(self._h, ms, n)
        if rc != 0:
            raise RuntimeError("profile_read failed")
        return list(ms), list(n)

    def l2_flush(self):
        return self._lib.processor_l2_flush(self._h)

    def open_stream(self, num_out_channels, sample_rate, period_frames, ring_periods=4):
        h = self._lib.stream_open(self._h, num_out_channels, sample_rate, period_frames, ring_periods)
        return Stream(self._lib, h, num_out_channels) if h else None

    def comm_init(self, rank, world_size, id128):
        self.world_size = world_size
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(id128))
        return self._lib.processor_comm_init(self._h, rank, world_size, buf)

    def comm_allgather(self, array):
        """All-gather a small host array over the processor's communicator: returns shape (world, *array.shape)."""
        a = np.ascontiguousarray(array)
        world = max(int(getattr(self, "world_size", 1)), 1)
        out = np.empty((world,) + a.shape, a.dtype)
        if self._lib.processor_comm_allgather(self._h, a.ctypes.data, out.ctypes.data, a.nbytes) != 0:
            raise RuntimeError("comm_allgather failed: " + (self._lib.last_device_error() or b"").decode())
        return out

    def free(self):
        if self._h:
            self._lib.processor_free(self._h)
            self._h = None


class FirewheelGraphCtx:
    """context.rs:29."""

    def __init__(self, lib, graph_config=None):
        self._lib = lib
        gc = graph_config or AudioGraphConfig()
        self.config = gc
        c = K.GraphConfig(gc.num_graph_inputs, gc.num_graph_outputs, gc.initial_node_capacity, gc.initial_edge_capacity,
                          gc.num_voices, int(gc.master_bus), gc.device, int(gc.max_call_frames))
        self._ctx = lib.ctx_new(C.byref(c))
        if not self._ctx:
            raise RuntimeError("ctx_new failed: " + (lib.last_device_error() or b"").decode())
        self.graph = AudioGraph(lib, self._ctx)

    @classmethod
    def new_batched(cls, flat, device=0, max_call_frames=0):
        """The batched context of a flat context whose graph passed detect_voices(): the voice graph once, num_voices = V, master bus in
        place of the SumNode tree. Returns (ctx, template node ids in the new graph)."""
        t = flat.graph.detect_voices()
        ids = (C.c_uint64 * max(t.num_template_nodes, 1))()
        h = flat._lib.ctx_new_batched(flat._ctx, device, max_call_frames, ids, t.num_template_nodes)
        if not h:
            raise ValueError(flat.last_error())
        self = cls.__new__(cls)
        self._lib, self._ctx = flat._lib, h
        self.config = AudioGraphConfig(num_graph_inputs=t.voice_inputs, num_graph_outputs=t.voice_outputs, num_voices=t.num_voices,
                                       master_bus=t.num_voices > 1 or flat.config.master_bus, device=device, max_call_frames=max_call_frames)
        self.graph = AudioGraph(flat._lib, h)
        return self, [NodeID(ids[i]) for i in range(t.num_template_nodes)]

    def activate(self, sample_rate, num_stream_in_channels, num_stream_out_channels, max_block_frames, user_cx=None):
        h = C.c_void_p(None)
        rc = self._lib.ctx_activate(self._ctx, sample_rate, num_stream_in_channels, num_stream_out_channels,
                                    max_block_frames, user_cx, C.byref(h))
        if rc == 1:
            return None  # already active (context.rs:57-59)
        if rc != 0:
            raise RuntimeError("activate failed: " + self.last_error())
        return FirewheelProcessor(self._lib, h.value, self.config)

    def is_activated(self):
        return bool(self._lib.ctx_is_activated(self._ctx))

    def update(self):
        st = K.UpdateStatusC()
        self._lib.ctx_update(self._ctx, C.byref(st))
        kind = ["Inactive", "Active", "Deactivated"][st.kind]
        err = None
        if st.graph_error != 0:
            err = CompileGraphError(st.graph_error, NodeID(st.error_node), st.error_port, self.last_error())
        return UpdateStatus(kind, err, st.returned_user_cx)

    def deactivate(self, stream_is_running):
        return self._lib.ctx_deactivate(self._ctx, int(stream_is_running))

    def last_error(self):
        return (self._lib.ctx_last_error(self._ctx) or b"").decode()

    def free(self):
        if self._ctx:
            self._lib.ctx_free(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
