// renderer_internal.h — what integrator.cu (single-device renderer) and multi.cu (one renderer per device +
// NCCL reduce) know about each other.  Not part of the ABI.
#pragma once
#include <cuda_runtime.h>

#include "../../include/ptb200.h"

struct pt_multi;  // multi.cu

// integrator.cu: fields of a single-device renderer that multi.cu needs
float4* pt_r_accum(pt_renderer* r);
cudaStream_t pt_r_stream(pt_renderer* r);
int pt_r_device(pt_renderer* r);
int pt_r_width(pt_renderer* r);
int pt_r_height(pt_renderer* r);
void pt_r_mark_started(pt_renderer* r);
void pt_r_set_ms_reduce(pt_renderer* r, float ms);
pt_multi* pt_r_multi(pt_renderer* r);
pt_renderer* pt_r_new_multi_handle(pt_scene* scene, int w, int h, pt_multi* m);  // a handle whose calls fan out
void pt_r_delete_handle(pt_renderer* r);
void pt_release_slab_cache();

// multi.cu: the fan-out behind the public entry points (called when pt_r_multi(r) != nullptr)
int pt_multi_destroy(pt_multi* m);
int pt_multi_render(pt_multi* m, int32_t first_frame, int32_t nframes, int32_t max_depth, int32_t russian, int32_t clear_accum);
int pt_multi_sync(pt_multi* m);
int pt_multi_merge(pt_multi* m);  // NCCL sum-reduce of the per-device buffers into the root's (no-op when merged)
pt_renderer* pt_multi_root(pt_multi* m);
int pt_multi_write_accum(pt_multi* m, const float* host_rgba);
int pt_multi_get_stats(pt_multi* m, pt_stats* out);
int pt_multi_set_flags(pt_multi* m, int32_t flags);
int pt_multi_set_pool(pt_multi* m, int64_t paths);
