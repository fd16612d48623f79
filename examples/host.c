/* examples/host.c — driving the C ABI from plain C (what a cgo / JNI / Rust -sys binding does underneath).
 *
 *   gcc -std=c99 -I include examples/host.c -o host -L firewheel_b200/lib -lfirewheel_b200 -Wl,-rpath,$PWD/firewheel_b200/lib
 *
 * Builds a voice graph (sampler -> gain -> graph_out), prints the compiled schedule, and — when a CUDA device is present —
 * activates it for 64 voices with a master bus, starts every voice on a looping sample and renders a few blocks, one call with
 * parameter stores stamped to take effect in the middle of it (fw_ctx_set_event_block).
 * Without a device it stops after the schedule: there is no CPU fallback. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "fw_b200.h"

int main(void) {
    fw_graph_config cfg;
    fw_graph_config_default(&cfg);
    cfg.num_graph_inputs = 0; cfg.num_graph_outputs = 2; cfg.num_voices = 64; cfg.master_bus = 1;
    fw_ctx* cx = fw_ctx_new(&cfg);
    if (!cx) { fprintf(stderr, "fw_ctx_new failed\n"); return 1; }

    fw_node_desc smp_d = {0}, vol_d = {0};
    smp_d.kind = FW_NODE_SAMPLER; smp_d.f0 = 100.0f;   /* percent volume */
    vol_d.kind = FW_NODE_VOLUME; vol_d.f0 = 50.0f;
    fw_node_id smp = fw_graph_add_node(cx, 0, 2, &smp_d), vol = fw_graph_add_node(cx, 2, 2, &vol_d), out = fw_graph_out_node(cx);
    for (uint32_t c = 0; c < 2; ++c) {
        if (fw_graph_connect(cx, smp, c, vol, c, 0, NULL, NULL, NULL) != FW_EDGE_OK || fw_graph_connect(cx, vol, c, out, c, 0, NULL, NULL, NULL) != FW_EDGE_OK) {
            fprintf(stderr, "connect failed\n"); return 1;
        }
    }
    if (fw_graph_compile_internal(cx, 256) != FW_COMPILE_OK) { fprintf(stderr, "compile failed: %s\n", fw_ctx_last_error(cx)); return 1; }
    printf("schedule: %u nodes, %u buffers\n", fw_schedule_len(cx), fw_schedule_num_buffers(cx));
    for (uint32_t i = 0; i < fw_schedule_len(cx); ++i) {
        fw_scheduled_node sn; fw_node_info ni;
        fw_schedule_node(cx, i, &sn); fw_graph_node_info(cx, sn.id, &ni);
        printf("  %-10s in %u out %u\n", ni.debug_name, sn.num_inputs, sn.num_outputs);
    }
    if (fw_device_count() <= 0) { printf("no CUDA device: stopping after the schedule (no CPU fallback)\n"); fw_ctx_free(cx); return 0; }

    fw_processor* proc = NULL;
    fw_update_status st;
    if (fw_ctx_activate(cx, 48000, 0, 2, 256, NULL, &proc) != 0 || fw_ctx_update(cx, &st) != 0 || st.graph_error != FW_COMPILE_OK) {
        fprintf(stderr, "activate / update failed: %s\n", fw_ctx_last_error(cx)); return 1;
    }
    enum { FRAMES = 48000 };
    float* pcm = (float*)malloc(sizeof(float) * 2 * FRAMES);   /* planar stereo: [channel][frame] */
    for (int i = 0; i < FRAMES; ++i) { pcm[i] = sinf(6.2831853f * 220.0f * (float)i / 48000.0f); pcm[FRAMES + i] = pcm[i]; }
    uint32_t res = fw_sample_resource_create(cx, FW_SAMPLE_F32_PLANAR, 2, FRAMES, pcm);
    fw_sampler_set_sample(cx, smp, FW_ALL_VOICES, res, 1);
    fw_sampler_set_loop_range(cx, smp, FW_ALL_VOICES, FW_LOOP_FULL, 0.0, 0.0);
    for (uint32_t v = 0; v < cfg.num_voices; ++v) fw_sampler_set_playhead(cx, smp, v, (double)v / 640.0);   /* spread the phases */
    fw_sampler_play(cx, smp, FW_ALL_VOICES);

    enum { T = 1024 };
    static float bus[T * 2];   /* interleaved stereo master bus */
    for (int call = 0; call < 3; ++call) {
        if (call == 1) {   /* block-stamped control: duck the gain at block 1 of this 4-block call and pause voice 0 at block 3 */
            fw_ctx_set_event_block(cx, 1); fw_volume_set_percent_volume(cx, vol, FW_ALL_VOICES, 20.0f);
            fw_ctx_set_event_block(cx, 3); fw_sampler_pause(cx, smp, 0);
            fw_ctx_set_event_block(cx, 0);
        }
        int rc = fw_processor_process_interleaved(proc, NULL, bus, 0, 2, T, call * (double)T / 48000.0, 0);
        double e = 0.0;
        for (int i = 0; i < T * 2; ++i) e += (double)bus[i] * bus[i];
        printf("call %d: rc %d, bus rms %.4f\n", call, rc, sqrt(e / (T * 2)));
    }
    fw_processor_free(proc);
    fw_ctx_update(cx, &st);
    fw_ctx_free(cx);
    free(pcm);
    return 0;
}
