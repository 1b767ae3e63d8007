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
    float *vertices, *normals, *colors, *grids;
    long long cap_vertices, cap_grids;
    float w_observed_min;                    // largest float T with double(T) <= 1e-12: weight > 1e-12 <=> weight > T
};
void launch_mesh_select(const MeshParams &mp, const DeviceMap &map, cudaStream_t st);
void launch_mesh_count(const MeshParams &mp, const DeviceMap &map, int nChunks, cudaStream_t st);
void launch_mesh_scan(const MeshParams &mp, const DeviceMap &map, int nChunks, cudaStream_t st);
void launch_mesh_emit(const MeshParams &mp, const DeviceMap &map, int nChunks, cudaStream_t st);

} // namespace chs
