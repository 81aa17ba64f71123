// sr_wave.h -- records that cross HBM between the stage kernels (sr_wave.cu) and their host-side launcher.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "sr_types.h"

namespace sr {

// a reflection ray waiting for its search (PathTracingMethod.cs:52 precedent: from pos + n*0.001)
struct SR_ALIGN(16) RefRay {
    double o[3], d[3];
    uint32_t sample, depth;       // sample of the chunk it belongs to; bounce number (1 = first reflection)
    uint32_t _pad[2];
};
static_assert(sizeof(RefRay) == 64, "RefRay must be 64 bytes");

// a shading point waiting for its ShadowMethod rays (ShadowMethod.cs:144-180)
struct SR_ALIGN(16) ShadowItem {
    double end[3];                // pos + normal * 0.001
    uint32_t slot;                // depth * n_samples + sample
    uint32_t inst;
};
static_assert(sizeof(ShadowItem) == 32, "ShadowItem must be 32 bytes");

// a shadow ray the FP32 filter could not decide
struct ShadowFallback {
    uint32_t item, sample;
    int32_t list[4];              // the undecided triangles (-1 unused); list[0] == -2: look at all of them
};

struct WaveCounts {
    uint32_t n_ref[2];            // reflection rays in ref[0] / ref[1]
    uint32_t n_shadow;            // shading points in `shadow`
    uint32_t shadow_head;         // next group of 32 shading points the shadow kernel hands to a warp
    uint32_t n_shadow_fallback;
    uint32_t n_walk, walk_head;   // shading points whose rays have to walk; next group k_shadow_walk hands out
    uint32_t n_fallback[8];       // per bounce: rays the candidate search could not bracket
    uint32_t _pad[1];
};

struct WaveBufs {
    WaveCounts* counts;
    int4* cand;                   // per ray of the batch: <= 4 candidate triangles (position in leaf order), -1 = none
    uint32_t* meta;               // per ray: bits 0-1 state, bits 4+7j..: instance of candidate j
    uint32_t* slot_color;         // [depth][sample]: colour of the shading point (Texture3D + ShadingMethod applied)
    uint32_t* slot_escaped;       // [depth][sample]: shadow rays that escaped
    uint8_t* sample_state;        // bit 7 valid; bits 0-2 shading points along the path; bit 3: then the background
    int32_t* sample_id;           // hit id of the camera ray
    RefRay* ref[2];
    ShadowItem* shadow;
    uint32_t* fallback;
    uint32_t* walk_list;          // positions in `shadow` of the points the cone walks could not settle
    ShadowFallback* shadow_fallback;
};

struct WaveLayout { size_t counts, cand, meta, slot_color, slot_escaped, sample_state, sample_id, ref0, ref1, shadow, fallback, walk_list, shadow_fallback; };

struct WaveArgs {
    DevFrame f;
    DevScene sc;
    const DevInstance* insts;
    const double* offsets;
    WaveBufs b;
    DevCounters* counters;
    int tile0;                    // first tile of the chunk
    uint32_t n_tiles, n_samples;  // of the chunk
    uint32_t n_rays;              // camera batch: = n_samples
    uint32_t cap_shadow_fallback;
    int ref_in, ref_out, fb_slot;
};

struct StageTimer {
    bool on = false;
    cudaStream_t st = nullptr;
    std::vector<cudaEvent_t> ev;
    std::vector<int> stage;
    void mark(int which);                       // an event now; the time since the previous mark is charged to `which`
    void collect(double* ms_stage, int n);      // after the stream has been synchronised
};

// Host surface of a frame whose chunks leave the GPU by DMA as they finish (softray_render with a page-locked surface):
// after each chunk's compose kernel its rows are copied device -> host on `copy_stream` while the next chunk traces.
struct HostCopy {
    uint32_t* h_pixels = nullptr; int32_t* h_ids = nullptr;          // page-locked host surface (full-frame indexing)
    const uint32_t* d_pixels = nullptr; const int32_t* d_ids = nullptr;
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t>* events = nullptr;                      // pool, grown on demand (owned by the context)
};

size_t wave_buffer_bytes(uint32_t cap_samples, int depth_slots, WaveLayout* lay);
void wave_bind(void* base, const WaveLayout& lay, WaveBufs* b);
cudaError_t wave_render(const DevFrame& f, const DevScene& sc, const DevInstance* d_insts, const double* d_offsets, const WaveBufs bufs[2],
                        uint32_t cap_samples, uint32_t* d_pixels, int32_t* d_ids, DevCounters* d_counters, int sm_count,
                        cudaStream_t stream, cudaStream_t side_stream, cudaEvent_t ev_fork, cudaEvent_t ev_join, int* launches,
                        StageTimer* prof, const HostCopy* host);

}  // namespace sr
