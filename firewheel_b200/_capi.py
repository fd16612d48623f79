"""ctypes declarations for include/fw_b200.h.

`bind(path, prefix)` loads a shared library that exports the header with the given
prefix and returns a namespace whose attributes are the un-prefixed entry points.
The product library uses prefix ``fw_``; the CPU oracle (tests only) uses ``fwo_``.
"""
import ctypes as C

FW_ID_DANGLING = 0xFFFFFFFFFFFFFFFF
FW_ALL_VOICES = 0xFFFFFFFF
FW_PARAM_PERCENT_VOLUME, FW_PARAM_RAW_GAIN, FW_PARAM_PAN, FW_PARAM_GAIN_L, FW_PARAM_GAIN_R, FW_PARAM_COEFFS = range(6)
FW_MAX_PORTS = 64

# fw_node_kind
NODE_DUMMY, NODE_VOLUME, NODE_SUM, NODE_MONO_TO_STEREO, NODE_STEREO_TO_MONO, NODE_HARD_CLIP = range(6)
NODE_PAN, NODE_BIQUAD, NODE_DELAY, NODE_CONV_REVERB, NODE_SAMPLER, NODE_SVF, NODE_RESAMPLER = 6, 7, 8, 9, 10, 11, 12

# fw_sample_format / fw_loop_mode / fw_sampler_status
SAMPLE_F32_PLANAR, SAMPLE_F32_INTERLEAVED, SAMPLE_I16_INTERLEAVED, SAMPLE_U16_INTERLEAVED, SAMPLE_I16_PLANAR, SAMPLE_U16_PLANAR = range(6)
LOOP_NONE, LOOP_FULL, LOOP_RANGE_SECS = 0, 1, 2
SAMPLER_ERRORS = {-1: "NotASampler", -2: "RingFull", -3: "NotActivated", -4: "BadArgs"}

# fw_add_edge_error / fw_compile_error names, index = code
ADD_EDGE_ERRORS = ["Ok", "SrcNodeNotFound", "DstNodeNotFound", "InPortOutOfRange", "OutPortOutOfRange",
                   "EdgeAlreadyExists", "InputPortAlreadyConnected", "CycleDetected"]
COMPILE_ERRORS = {0: "Ok", 1: "CycleDetected", 2: "NodeOnEdgeNotFound", 3: "NodeIDNotUnique", 4: "EdgeIDNotUnique",
                  5: "ManyToOneError", 6: "NodeActivationFailed", 7: "MessageChannelFull", 100: "UnsupportedOnDevice"}

PROC_OK, PROC_DROP_PROCESSOR, PROC_DEVICE_ERROR, PROC_BAD_ARGS = 0, 1, -1, -2
UPDATE_INACTIVE, UPDATE_ACTIVE, UPDATE_DEACTIVATED = 0, 1, 2


class GraphConfig(C.Structure):
    _fields_ = [("num_graph_inputs", C.c_uint32), ("num_graph_outputs", C.c_uint32),
                ("initial_node_capacity", C.c_uint32), ("initial_edge_capacity", C.c_uint32),
                ("num_voices", C.c_uint32), ("master_bus", C.c_uint32), ("device", C.c_int32),
                ("max_call_frames", C.c_uint32)]


class NodeDesc(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("u0", C.c_uint32), ("u1", C.c_uint32), ("u2", C.c_uint32),
                ("f0", C.c_float), ("f1", C.c_float), ("f2", C.c_float), ("f3", C.c_float),
                ("data", C.POINTER(C.c_float)), ("data_len", C.c_uint64)]


class UpdateStatusC(C.Structure):
    _fields_ = [("kind", C.c_int32), ("graph_error", C.c_int32), ("error_node", C.c_uint64),
                ("error_port", C.c_uint32), ("reserved", C.c_uint32), ("returned_user_cx", C.c_void_p)]


class NodeInfoC(C.Structure):
    _fields_ = [("num_inputs", C.c_uint32), ("num_outputs", C.c_uint32), ("kind", C.c_uint32),
                ("num_min_supported_inputs", C.c_uint32), ("num_max_supported_inputs", C.c_uint32),
                ("num_min_supported_outputs", C.c_uint32), ("num_max_supported_outputs", C.c_uint32),
                ("updates", C.c_uint32), ("debug_name", C.c_char * 32)]


class EdgeInfoC(C.Structure):
    _fields_ = [("id", C.c_uint64), ("src_node", C.c_uint64), ("dst_node", C.c_uint64),
                ("src_port", C.c_uint32), ("dst_port", C.c_uint32)]


class ScheduledNodeC(C.Structure):
    _fields_ = [("id", C.c_uint64), ("num_inputs", C.c_uint32), ("num_outputs", C.c_uint32),
                ("in_buffer", C.c_uint32 * FW_MAX_PORTS), ("in_should_clear", C.c_uint8 * FW_MAX_PORTS),
                ("out_buffer", C.c_uint32 * FW_MAX_PORTS)]


class VoiceTemplateC(C.Structure):
    _fields_ = [("num_voices", C.c_uint32), ("num_template_nodes", C.c_uint32), ("voice_inputs", C.c_uint32), ("voice_outputs", C.c_uint32),
                ("num_tree_nodes", C.c_uint32)]


_vp, _u32, _u64, _i32, _f32, _f64 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_float, C.c_double
_pf = C.c_void_p  # float* passed as raw addresses (numpy .ctypes.data or device pointers)
_pu64 = C.POINTER(C.c_uint64)
_pu32 = C.POINTER(C.c_uint32)

# name -> (restype, argtypes): every symbol include/fw_b200.h declares
SIGNATURES = {
    "graph_config_default": (None, [C.POINTER(GraphConfig)]),
    "ctx_new": (_vp, [C.POINTER(GraphConfig)]),
    "ctx_free": (None, [_vp]),
    "ctx_last_error": (C.c_char_p, [_vp]),
    "graph_in_node": (_u64, [_vp]),
    "graph_out_node": (_u64, [_vp]),
    "graph_add_node": (_u64, [_vp, _u32, _u32, C.POINTER(NodeDesc)]),
    "graph_add_custom_node": (_u64, [_vp, _u32, _u32, _vp, _vp]),
    "graph_remove_node": (_i32, [_vp, _u64, _pu64, _u32, _pu32]),
    "graph_set_num_inputs": (_i32, [_vp, _u64, _u32, _pu64, _u32, _pu32]),
    "graph_set_num_outputs": (_i32, [_vp, _u64, _u32, _pu64, _u32, _pu32]),
    "graph_connect": (_i32, [_vp, _u64, _u32, _u64, _u32, _i32, _pu64, _pu64, _pu32]),
    "graph_disconnect": (_i32, [_vp, _u64, _u32, _u64, _u32]),
    "graph_disconnect_by_edge_id": (_i32, [_vp, _u64]),
    "graph_edge": (_i32, [_vp, _u64, C.POINTER(EdgeInfoC)]),
    "graph_node_info": (_i32, [_vp, _u64, C.POINTER(NodeInfoC)]),
    "graph_num_nodes": (_u32, [_vp]),
    "graph_num_edges": (_u32, [_vp]),
    "graph_nodes": (_u32, [_vp, _pu64, _u32]),
    "graph_edges": (_u32, [_vp, _pu64, _u32]),
    "graph_cycle_detected": (_i32, [_vp]),
    "graph_reset": (None, [_vp]),
    "graph_needs_compile": (_i32, [_vp]),
    "graph_compile_internal": (_i32, [_vp, _u32]),
    "schedule_len": (_u32, [_vp]),
    "schedule_num_buffers": (_u32, [_vp]),
    "schedule_node": (_i32, [_vp, _u32, C.POINTER(ScheduledNodeC)]),
    "graph_detect_voices": (_i32, [_vp, C.POINTER(VoiceTemplateC)]),
    "graph_voice_nodes": (_u32, [_vp, _u32, C.POINTER(C.c_uint64), _u32]),
    "ctx_new_batched": (_vp, [_vp, _i32, _u32, C.POINTER(C.c_uint64), _u32]),
    "node_read_params": (_u32, [_vp, C.c_uint64, _u32, _vp, _u32]),
    "ctx_set_event_block": (None, [_vp, _u32]),
    "volume_set_percent_volume": (_i32, [_vp, _u64, _u32, _f32]),
    "volume_set_percent_volumes": (_i32, [_vp, _u64, _pf, _u32]),
    "pan_set_pan": (_i32, [_vp, _u64, _u32, _f32]),
    "pan_set_pans": (_i32, [_vp, _u64, _pf, _u32]),
    "pan_set_gains": (_i32, [_vp, _u64, _u32, _f32, _f32]),
    "biquad_set_coeffs": (_i32, [_vp, _u64, _u32, _u32, _pf]),
    "biquad_set_all_coeffs": (_i32, [_vp, _u64, _pf, _u32, _u32]),
    "biquad_design_rbj": (None, [_u32, _f64, _f64, _f64, _f64, _pf]),
    "svf_set_coeffs": (_i32, [_vp, _u64, _u32, _u32, _pf]),
    "svf_set_all_coeffs": (_i32, [_vp, _u64, _pf, _u32, _u32]),
    "svf_design": (None, [_u32, _f64, _f64, _f64, _pf]),
    "resampler_set": (_i32, [_vp, _u64, _u32, _u32, _u64, _i32, _i32]),
    "resampler_seek": (_i32, [_vp, _u64, _u32, _u64]),
    "resampler_design": (None, [_u32, _u32, _f64, _f64, _pf]),
    "sample_resource_create": (_u32, [_vp, _u32, _u32, _u64, _vp]),
    "sampler_set_sample": (_i32, [_vp, _u64, _u32, _u32, _i32]),
    "sampler_play": (_i32, [_vp, _u64, _u32]),
    "sampler_pause": (_i32, [_vp, _u64, _u32]),
    "sampler_stop": (_i32, [_vp, _u64, _u32]),
    "sampler_set_playhead": (_i32, [_vp, _u64, _u32, _f64]),
    "sampler_set_loop_range": (_i32, [_vp, _u64, _u32, _u32, _f64, _f64]),
    "sampler_set_percent_volume": (_i32, [_vp, _u64, _u32, _f32]),
    "sampler_is_playing": (_i32, [_vp, _u64, _u32]),
    "ctx_activate": (_i32, [_vp, _u32, _u32, _u32, _u32, _vp, C.POINTER(_vp)]),
    "ctx_is_activated": (_i32, [_vp]),
    "ctx_update": (_i32, [_vp, C.POINTER(UpdateStatusC)]),
    "ctx_deactivate": (_vp, [_vp, _i32]),
    "processor_process_interleaved": (_i32, [_vp, _pf, _pf, _u32, _u32, _u64, _f64, _u32]),
    "processor_process_planar": (_i32, [_vp, _pf, _pf, _u32, _u32, _u64, _f64, _u32, _pu64]),
    "processor_process_planar_device": (_i32, [_vp, _pf, _pf, _u32, _u32, _u64, _f64, _u32]),
    "processor_free": (None, [_vp]),
    "stream_open": (_vp, [_vp, _u32, _u32, _u32, _u32]),
    "stream_pull": (C.c_int64, [_vp, _vp, _u64, _pu32, C.POINTER(C.c_double)]),
    "stream_frames_ready": (_u64, [_vp]),
    "stream_close": (None, [_vp]),
    "device_count": (_i32, []),
    "last_device_error": (C.c_char_p, []),
    "dev_malloc": (_vp, [_i32, _u64]),
    "dev_free": (None, [_i32, _vp]),
    "host_alloc_pinned": (_vp, [_u64]),
    "host_free_pinned": (None, [_vp]),
    "processor_h2d": (_i32, [_vp, _vp, _vp, _u64]),
    "processor_d2h": (_i32, [_vp, _vp, _vp, _u64]),
    "processor_sync": (_i32, [_vp]),
    "processor_event_record": (_i32, [_vp, _i32]),
    "processor_event_elapsed_ms": (_f32, [_vp, _i32, _i32]),
    "processor_kernel_launches": (_u64, [_vp]),
    "processor_graph_replays": (_u64, [_vp]),
    "processor_l2_flush": (_i32, [_vp]),
    "processor_profile": (_i32, [_vp, _i32]),
    "processor_profile_read": (_i32, [_vp, C.c_void_p, C.c_void_p]),
    "comm_unique_id": (_i32, [C.c_void_p]),
    "processor_comm_init": (_i32, [_vp, _i32, _i32, C.c_void_p]),
    "processor_comm_allgather": (_i32, [_vp, C.c_void_p, C.c_void_p, _u64]),
}


class Lib:
    """A loaded implementation of the C ABI. Attributes are the un-prefixed entry points."""

    def __init__(self, path, prefix, extra=None):
        self.path, self.prefix = str(path), prefix
        self.cdll = C.CDLL(self.path, mode=C.RTLD_GLOBAL if prefix == "fw_" else C.RTLD_LOCAL)
        sigs = dict(SIGNATURES)
        sigs.update(extra or {})
        for name, (res, args) in sigs.items():
            fn = getattr(self.cdll, prefix + name)  # AttributeError => symbol missing: fail loudly
            fn.restype, fn.argtypes = res, args
            setattr(self, name, fn)

    def __repr__(self):
        return f"<fw Lib {self.prefix}* from {self.path}>"


def bind(path, prefix, extra=None):
    return Lib(path, prefix, extra)
