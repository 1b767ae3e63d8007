// kernels.h -- host-callable launchers of the CUDA kernels (integrate.cu, mesh.cu).
#pragma once

#include <cuda_runtime.h>

#include "../../include/chisel_b200.h"
#include "device_map.cuh"

namespace chs
{

struct FrameGraph;
FrameGraph *frame_graph_create();
void frame_graph_destroy(FrameGraph *fg);
// newHint: expected number of new-chunk candidates (sizes that kernel's grid; any value is correct).
// Enqueue one frame (prepare -> candidates -> {new chunks || bricks} -> counter snapshot into hostSlot) as one graph launch.
cudaError_t frame_graph_launch(FrameGraph *fg, const FrameParams &fp, const DeviceMap &map, long long candidates, long long newHint, HostSnapshot *hostSlot,
                               bool profiling, cudaEvent_t *evt, cudaStream_t st);
float host_truncation(int kind, float param, float depth);

// ---- fused multi-frame integration (integrate_batch_impl.cuh) ----
// K <= kMaxBatch consecutive frames in one pass: every voxel of the map is independent of every other voxel, so applying
// frames f0 < f1 < ... to a voxel while its state sits in registers gives the same bits as K separate passes.
constexpr int kMaxBatch = 16;

struct BatchCounters
{
    int unit_count, new_count, tickets, next_task;     // unit_count: units of cost bucket 0 (most frames)
    int chunks_at_start, light_count, bucket1, bucket2; // units of buckets 3 (fewest frames), 1 and 2
    int candidates[kMaxBatch], n_new[kMaxBatch], updated_chunks[kMaxBatch], pad2[kMaxBatch];
    unsigned long long n_upd[kMaxBatch], n_carve[kMaxBatch], n_col[kMaxBatch];
    unsigned long long span_start_inv, span_end;       // brick kernel: max over CTAs of ~(start time) and of the end time (%globaltimer, ns)
};

// Device timeline of a batch (chs_set_profiling(map, 2)): %globaltimer stamps taken by the kernels themselves -- no events between
// the kernels, so the overlap of the product path (PDL, the exchange beside the brick kernel) is seen as it is.
enum TimelineStamp
{
    kTlPushStart = 0,   // peer-memory exchange of this step: first CTA of the push kernel ...
    kTlPushEnd,         // ... and its last CTA (data out, flags raised)
    kTlWaitEnd,         // the images of all ranks have arrived (peer_wait_kernel)
    kTlHizStart,
    kTlHizEnd,
    kTlCandStart,
    kTlCandEnd,
    kTlBricksStart,
    kTlBricksEnd,
    kTimelineStamps
};

// Written into a pinned slot by the last CTA of a batch: head, payload, checksum of the payload, tail -- WITHOUT a system-scope
// fence (a fence makes the kernel wait for a PCIe round trip). Valid for batch b when head == tail == b and the checksum matches
// the payload the host reads (stores that have not landed yet make it mismatch: the host simply polls again).
struct HostBatchSnapshot
{
    int head, n_chunks, n_dirty, error_flags;
    int unit_count, new_count, K, pad;
    int candidates[kMaxBatch], n_new[kMaxBatch], updated_chunks[kMaxBatch];
    long long n_upd[kMaxBatch], n_carve[kMaxBatch], n_col[kMaxBatch];
    long long bricks_span_ns;       // first CTA start -> last CTA end of the brick kernel, from %globaltimer (no launch / event overhead)
    long long timeline[kTimelineStamps];   // device timeline of the batch (%globaltimer, ns; 0 = not recorded): see BatchParams::timeline
    unsigned long long checksum;    // batch_snapshot_checksum of everything between head and here
    int tail, pad3[3];
};

// 64-bit sum of the payload words, seeded with the batch id (host and device compute it the same way)
__host__ __device__ inline unsigned long long batch_snapshot_checksum(const HostBatchSnapshot &h)
{
    unsigned long long s = 0x9E3779B97F4A7C15ull * (unsigned long long)(unsigned)h.head;
    s += (unsigned)h.n_chunks + ((unsigned long long)(unsigned)h.n_dirty << 1) + ((unsigned long long)(unsigned)h.error_flags << 2) +
         ((unsigned long long)(unsigned)h.unit_count << 3) + ((unsigned long long)(unsigned)h.new_count << 4) + ((unsigned long long)(unsigned)h.K << 5);
    for (int t = 0; t < kMaxBatch; t++)
        s += (unsigned long long)(unsigned)h.candidates[t] * 3 + (unsigned long long)(unsigned)h.n_new[t] * 5 + (unsigned long long)(unsigned)h.updated_chunks[t] * 7 +
             (unsigned long long)h.n_upd[t] * 11 + (unsigned long long)h.n_carve[t] * 13 + (unsigned long long)h.n_col[t] * 17;
    for (int t = 0; t < kTimelineStamps; t++)
        s += (unsigned long long)h.timeline[t] * (unsigned long long)(23 + 2 * t);
    return s + (unsigned long long)h.bricks_span_ns * 19;
}

struct BatchParams
{
    const FrameParams *frames;      // device array [K]; each entry is complete, exactly as the single-frame path would fill it
    int K;
    int lo[3], n[3];                // union of the K candidate ID boxes
    int4 *units;                    // bricks: {id key low, id key high, slot | brick << 24, band frames | free-space frames << 16};
                                    // slot 0xFFFFFF: the chunk does not exist yet. TWO buffers of units_cap entries: cost buckets 0 / 1
                                    // grow from the front / back of the first, buckets 2 / 3 from the front / back of the second
    int units_cap;
    int cand_stride;                // multiplicative permutation of the union box enumeration (coprime to its size)
    int batch_id;                   // > 0, increases by one per batch
    int total_ctas;                 // CTAs of the brick kernel (the last one takes the snapshot)
    BatchCounters *bctr;
    HostBatchSnapshot *host_slot;   // pinned, device-mapped
    unsigned long long *slot_batch; // [capacity] (batch id << 32) | mask of the batch's frames that updated the chunk
    int reserve_sms;                // the brick kernel leaves every 16th SM idle: room for the NCCL kernels of the next batch's exchange
    int coarse_in_shared;           // the candidates kernel builds the Hi-Z levels >= 4 in shared memory (they fit: <= 48 tiles per frame)
    // peer-memory frame exchange (capi_comm.inc): the last CTA of the batch tells every rank that this rank has finished reading
    // the step's images, i.e. that the staging set may be overwritten by the pushes of the step after next
    unsigned long long *timeline;   // [kTimelineStamps] device words, zero before the batch: starts as max of ~time, ends as max of time; nullptr: off
    unsigned *const *peer_done;     // [peer_world] device table (local): entry p = this rank's "done" word in rank p's exchange arena
    int peer_world;                 // 0: not a peer-memory step
    unsigned peer_step;
};

// What the fast brick kernel needs of one frame, 128 bytes, passed BY VALUE in the kernel's parameter block (constant bank): a
// warp reads the frame it is working on with warp-uniform constant loads, nothing is staged through shared memory.
struct alignas(16) BrickFrame
{
    float R[9];                     // row-major rotation, camera -> world
    float t[3];
    float fx, fy, cx, cy;           // cx, cy with -0.0f canonicalised to +0.0f (same pixel decisions; lets the on-image test be a
                                    // compare of the float bit patterns)
    float Wf, Hf;
    int W;
    int pix_bias;                   // -(M * W + M) mod 2^32, M = 0x4B000000: turns the magic-number floors into the pixel index
    float thr_band;                 // trunc + diag          (ProjectionIntegrator.h:82 / :143)
    float thr_carve;                // trunc + carvingDist   (:88 / :166); +inf when carving is off
    float wu;                       // weight update: 1.0f (depth path, quirk Q7) or weight / (5 * trunc)
    float cutoff;                   // 50 / 100
    const float *depth;
    const unsigned *color;          // packed r | g << 8 | b << 16
    float carve_max;                // sdf < 1e-5 as a binary32 threshold
    int pad[3];
};
static_assert(sizeof(BrickFrame) == 128, "BrickFrame is read as eight 16-byte constant loads");
struct BrickFrames
{
    BrickFrame f[kMaxBatch];
};

struct BatchLaunchInfo
{
    int W, H, cW, cH;               // depth / colour image size (identical for all frames of a batch)
    long long unionCandidates;      // chunk IDs in the union box
    bool colorPath, perPixel, profiling;
    int hizFirst, hizCount;         // frames whose Hi-Z pyramid THIS launch builds (distributed batches: the frames this rank ingests; the
                                    // other ranks' pyramids arrive by all-gather, see afterHiz); hizCount == 0: all K frames
    int (*afterHiz)(void *);        // called right after the Hi-Z kernel has been enqueued on the prepare stream (or nullptr)
    void *afterHizCtx;
    bool hizTma;                    // Hi-Z by the TMA bulk-copy kernel (float depth, constant truncator, aligned rows)
    bool skipHiz;                   // the pyramids of this step arrive through the exchange arenas (launch_hiz_sharded on every rank)
    bool fastBricks;                // every frame satisfies the preconditions of batch_bricks_fast_kernel (brick_frames is filled)
    const BrickFrames *brickFrames; // host pointer; copied into the kernel's parameter block
};
// Streams of one batch. The Hi-Z kernel runs on `prep` and records `prepared`; the candidates kernel waits for it on `main`. The
// colour packing kernel (only the brick kernel reads its output) runs on `pack` BESIDE the candidates kernel and records
// `packed`, which the brick kernel waits for. pack == prep: it simply follows the Hi-Z kernel there (host frames: both follow
// the copies on the copy stream). pack != prep (device frames, prep == main): forked from `main` with `fork`.
// CHS_HOST_PROFILE: laps inside launch_batch (capi.cu); i < 0 restarts the clock
void host_launch_lap(int i);

// Sharded Hi-Z of a distributed step (peer-memory exchange): a rank builds the pyramids of the frames IT ingests and stores every
// tile into the exchange arena of every rank; its last CTA then raises the rank's "arrived" word everywhere.
constexpr int kMaxPeersDev = 16;
struct HizPeers
{
    long long delta[kMaxPeersDev];  // byte distance from this rank's arena to rank d's, as mapped here (0 for the rank itself)
    unsigned *flag[kMaxPeersDev];   // this rank's arrived word in rank d's arena
    unsigned *hdr;                  // this rank's arena header: ticket word [129], stamps at byte 2048
    int world, rank;
    unsigned step;
    int stamp_slot;                 // which pair of stamps (the step's staging set)
    // what the kernel needs of the frames, by value (it runs before the step's frame table is on the device): frame f's image is
    // depth0 + f * npx, level l of its pyramid hiz0 + f * tiles_per_frame + level_off[l]
    const float *depth0;
    float2 *hiz0;
    long long npx, tiles_per_frame;
    int level_off[4], hizW[4], hizH[4];
    int W, H, carve;
    float cutoff, trunc, diag, carve_dist;
};

// Sharded Hi-Z (see HizPeers): pyramids of the frames [first, first + count) of the table `frames`, on stream st.
namespace half
{
cudaError_t launch_hiz_sharded(int first, int count, const HizPeers &hp, cudaStream_t st);
}

struct BatchStreams
{
    cudaStream_t prep, pack, main;
    cudaEvent_t prepared, packed, fork;
};
// events (profiling): [0] start, [1] after the Hi-Z kernel (both on prep), [2] = [7] after candidates, [3] after bricks (on main)
// phases: bit 0 = prepare + candidates, bit 1 = bricks (3 = the whole batch; the host may size the pool between the two)
// integrate_batch_impl.cuh is built twice (half-brick tasks for depth-only batches, quarter-brick tasks for colour batches)
namespace half
{
cudaError_t launch_batch(const BatchParams &bp, const DeviceMap &map, const BatchLaunchInfo &info, cudaEvent_t *evt, const BatchStreams &bs, int phases);
}
namespace quarter
{
cudaError_t launch_batch(const BatchParams &bp, const DeviceMap &map, const BatchLaunchInfo &info, cudaEvent_t *evt, const BatchStreams &bs, int phases);
}
inline cudaError_t launch_batch(const BatchParams &bp, const DeviceMap &map, const BatchLaunchInfo &info, cudaEvent_t *evt, const BatchStreams &bs, int phases)
{
    return info.colorPath ? quarter::launch_batch(bp, map, info, evt, bs, phases) : half::launch_batch(bp, map, info, evt, bs, phases);
}

// dCounters[4] (zeroed): rcp mismatches, rcp tested, div mismatches, div tested
namespace half
{
cudaError_t launch_selftest_arithmetic(unsigned long long *dCounters, unsigned long long divPairs, cudaStream_t st);
}

// table maintenance (capi.cu)
void launch_fill_u64(unsigned long long *p, size_t n, unsigned long long v, cudaStream_t st);
void launch_rebuild_hash(const DeviceMap &map, int nChunks, cudaStream_t st);
void launch_rebuild_dirty(const DeviceMap &map, int nDirty, cudaStream_t st);

// meshing (mesh.cu)
struct MeshParams
{
    const unsigned long long *dirty_list;   // packed IDs to consider
    int n_dirty;
    int *mesh_slots;                         // [n_dirty] compacted: pool slot of each existing dirty chunk
    int *tri_counts;                         // [n] triangles per remeshed chunk, then exclusive offsets
    int *grid_counts;                        // [n] occupied cells per remeshed chunk, then exclusive offsets
    long long *vert_offsets;                 // [n+1]
    long long *grid_offsets;                 // [n+1]
    unsigned char *cfg_scratch;              // [n * V] cube configuration of every cell in the reference's cell order (0: no triangle)
    float *vertices, *normals, *colors, *grids;
    long long cap_vertices, cap_grids;
    float w_observed_min;                    // largest float T with double(T) <= 1e-12: weight > 1e-12 <=> weight > T
};
void launch_mesh_select(const MeshParams &mp, const DeviceMap &map, cudaStream_t st);
void launch_mesh_count(const MeshParams &mp, const DeviceMap &map, int nChunks, cudaStream_t st);
void launch_mesh_scan(const MeshParams &mp, const DeviceMap &map, int nChunks, cudaStream_t st);
void launch_mesh_emit(const MeshParams &mp, const DeviceMap &map, int nChunks, cudaStream_t st);

} // namespace chs
