/* chisel_b200.h -- C ABI of the B200-native OpenChisel hot path (TSDF integration + incremental
 * marching cubes). This is the drop-in boundary: the open_chisel C++ facade under
 * cvids_b200/include/open_chisel/ and the Python harness (cvids_b200/capi.py) both call ONLY these
 * entry points. Plain pointers and sizes; no C++ or torch types; int status codes; no exceptions.
 *
 * Reference interface each entry point replaces (paths relative to the reference's
 * OpenChisel/open_chisel/, "OC/"; callers in OpenChisel/chisel_ros/, "CR/"):
 *
 *   chs_create               chisel::Chisel::Chisel(chunkSize, res, useColor)       OC/src/Chisel.cpp:34-37, CR/src/ChiselServer.cpp:189
 *   chs_destroy              chisel::Chisel::~Chisel                                 OC/src/Chisel.cpp:39-42
 *   chs_reset                chisel::Chisel::Reset                                   OC/src/Chisel.cpp:44-48, CR/src/ChiselServer.cpp:200
 *   chs_integrate_depth      chisel::Chisel::IntegrateDepthScan<float>               OC/include/open_chisel/Chisel.h:59-112, CR/src/ChiselServer.cpp:501
 *   chs_integrate_depth_color chisel::Chisel::IntegrateDepthScanColor<float,uint8_t> OC/include/open_chisel/Chisel.h:114-213, CR/src/ChiselServer.cpp:497
 *   chs_integrate_batch      n consecutive calls of the two entries above (the caller queues frames; OC Chisel.h:59-213)
 *   chs_update_meshes        ChunkManager::RecomputeMeshes(meshesToUpdate) + clear    OC/src/ChunkManager.cpp:130-169, OC/src/Chisel.cpp:55-57
 *                            (the every-10th-call gate of Chisel::UpdateMeshes, Chisel.cpp:50-59, lives in the facade)
 *   chs_num_dirty/_dirty_ids chisel::Chisel::GetMeshesToUpdate                       OC/include/open_chisel/Chisel.h:221-224, CR/src/ChiselServer.cpp:346,554
 *   chs_num_chunks/_chunk_ids/_download_chunk  ChunkManager::GetChunks, HasChunk, GetChunk, Chunk::GetVoxels/GetColorVoxels
 *                                                                                    OC/include/open_chisel/ChunkManager.h:69-87, CR/include/chisel_ros/Serialization.h:31-84
 *   chs_mesh_counts/_download_meshes  ChunkManager::GetAllMeshes / Mesh fields       OC/include/open_chisel/ChunkManager.h:163-182, OC/include/open_chisel/mesh/Mesh.h:52-57
 *   chs_frustum              PinholeCamera::SetupFrustum, Frustum::GetLines/GetCorners OC/src/camera/PinholeCamera.cpp:55-59, OC/src/geometry/Frustum.cpp:143-219
 *   chs_candidate_ids        ChunkManager::GetChunkIDsIntersecting(Frustum)          OC/src/ChunkManager.cpp:182-212
 *   chs_truncation           Truncator::GetTruncationDistance (3 shipped subclasses) OC/include/open_chisel/truncation/ (all three headers)
 *
 * Threading: one caller thread per map (the reference's ChiselNode is a single-threaded ros::spin,
 * CR/src/ChiselNode.cpp:135). Host buffers are caller-owned; device memory is library-owned.
 * All work of a map runs on one CUDA stream; calls return after ENQUEUEING unless stated otherwise.
 * There is no CPU fallback: every entry point that computes fails with CHS_ERR_CUDA without a device.
 */
#ifndef CHISEL_B200_H
#define CHISEL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CHS_ABI_VERSION 1

typedef struct chs_map chs_map; /* opaque */

enum
{
    CHS_OK = 0,
    CHS_ERR_INVALID = 1,   /* bad argument (null, non-cubic chunk, non-finite pose, ID out of the packable range) */
    CHS_ERR_CUDA = 2,      /* a CUDA call failed; see chs_last_error_string */
    CHS_ERR_CAPACITY = 3,  /* a device table overflowed (should not happen: capacities are grown ahead of need) */
    CHS_ERR_NOT_FOUND = 4
};

enum { CHS_TRUNC_CONSTANT = 0, CHS_TRUNC_QUADRATIC = 1, CHS_TRUNC_INVERSE = 2, CHS_TRUNC_PER_PIXEL = 3 };
/* CHS_MEM_HOST_ASYNC (chs_integrate_batch only): host buffers (ideally pinned) that the caller leaves untouched until the
 * call's ticket has been waited for (chs_wait_batch / any synchronising call); the call returns right after enqueueing, so
 * the copies of successive batches run back to back on the copy engine. */
enum { CHS_MEM_HOST = 0, CHS_MEM_DEVICE = 1, CHS_MEM_HOST_ASYNC = 2, CHS_MEM_DEVICE_ASYNC = 3 };
/* CHS_MEM_DEVICE_ASYNC (chs_integrate_batch only): device buffers whose contents are COMPLETE when the call is made (not merely
 * ordered on the map's stream) and stay untouched until the call's ticket has been waited for. The library then prepares the
 * batch (Hi-Z pyramid, colour packing) on its copy stream, beside the kernels of the previous batch, instead of behind them. */

typedef struct
{
    int chunk_size;              /* voxels per chunk edge; cubic chunks only (8, 16 or 32), see DESIGN.md Q12 */
    float resolution;            /* voxel edge in metres */
    int use_color;
    int device;                  /* CUDA ordinal; -1 = current device */
    int rank, world;             /* chunk-ownership shard: this map keeps chunk IDs with chs_owner(id) % world == rank; world <= 1: all */
    int64_t initial_chunks;      /* initial pool capacity in chunks; 0 = default */
    void *stream;                /* cudaStream_t to enqueue on; NULL = the library creates a non-blocking stream */
} chs_config;

typedef struct
{
    float fx, fy, cx, cy;
    int width, height;
    float near_plane, far_plane;
} chs_camera;

/* ProjectionIntegrator state (OC/include/open_chisel/ProjectionIntegrator.h:185-229) */
typedef struct
{
    int trunc_kind;              /* CHS_TRUNC_* */
    float trunc_param;           /* constant: metres; quadratic / inverse: scale */
    const float *trunc_per_pixel;/* CHS_TRUNC_PER_PIXEL: W*H truncation distances evaluated by the caller (any Truncator subclass), same memory space as the depth image */
    float weight;                /* ConstantWeighter(weight); colour path only (the depth path integrates with 1.0f, quirk Q7) */
    int carving_enabled;
    float carving_dist;
} chs_integrator;

/* Counters of the last integrated frame (SURVEY.md section 8(d) definitions). */
typedef struct
{
    int64_t candidates;          /* chunk IDs the reference would iterate (box passing Frustum::Intersects), owned by this rank */
    int64_t new_candidates;      /* non-existing candidates that survived the conservative depth-range cull (tested exactly) */
    int64_t brick_units;         /* 8^3 bricks of existing chunks that survived the cull (each visited once) */
    int64_t n_upd;               /* voxels on which DistVoxel::Integrate ran */
    int64_t n_carve;             /* voxels reset or decremented by carving */
    int64_t n_col;               /* voxels whose colour was written */
    int64_t n_new;               /* chunks created this frame (all survive: untouched ones are never materialised) */
    int64_t updated_chunks;      /* chunks whose integrate returned true */
    int64_t total_chunks;        /* chunks in the map after the frame */
    int64_t dirty_chunks;        /* size of the dirty set after the frame */
    int64_t error_flags;         /* non-zero: a device table overflowed */
} chs_frame_stats;

typedef struct
{
    int64_t n_chunks;            /* remeshed chunks in the last chs_update_meshes (existing dirty chunks, incl. those with no triangle) */
    int64_t n_vertices;          /* total vertices (3 per triangle) */
    int64_t n_grids;             /* total occupied-cell centres (Mesh::grids) */
    int has_colors;
} chs_mesh_counts;

/* Device time of the phases of the last frame / last re-mesh, from CUDA events on the map's stream
 * (recorded only while profiling is enabled; chs_get_timings synchronises). Milliseconds. */
typedef struct
{
    float prepare_ms, candidates_ms, new_chunks_ms, integrate_ms, frame_ms;   /* integrate_ms: the brick kernel (existing chunks) */
    float mesh_count_ms, mesh_emit_ms, mesh_ms;
    float bricks_span_ms;        /* fused path: first CTA start -> last CTA end of the brick kernel of the last finished batch, from the device's
                                    %globaltimer (kernel time without launch and event-record overhead; integrate_ms is the event-timed figure) */
} chs_timings;

/* Device timeline of the most recent fused batches (chs_set_profiling(map, 2); bit 0 = the event timings above, which serialise
   the kernels): absolute %globaltimer stamps in ns taken by the kernels themselves, 0 = not recorded. Per batch
   CHS_TIMELINE_STAMPS values: push start, push end (peer-memory exchange of the step), arrival of all ranks' images, Hi-Z start /
   end, candidates start / end, bricks start / end. out: room for max_batches batches (the library keeps 256), oldest first. */
#define CHS_TIMELINE_STAMPS 9
int chs_get_device_timeline(chs_map *map, long long *out, int max_batches, int *n_batches);

const char *chs_last_error_string(void);
int chs_abi_version(void);

int chs_create(const chs_config *cfg, chs_map **out);
int chs_destroy(chs_map *map);
int chs_reset(chs_map *map);
int chs_synchronize(chs_map *map);
int chs_set_stream(chs_map *map, void *cuda_stream);
int chs_set_profiling(chs_map *map, int enabled);

/* pose: row-major 3x4 [R|t], camera -> world. depth: width*height float metres (NaN = invalid).
 * mem: CHS_MEM_HOST (copied H2D on the map's stream) or CHS_MEM_DEVICE (used in place). */
int chs_integrate_depth(chs_map *map, const chs_integrator *integ, const float *depth, int mem,
                        const float pose[12], const chs_camera *cam);
/* color: cam.width*cam.height*channels uint8, channels 1 (mono), 3 (BGR) or 4 (BGRA) (OC ColorImage.h:61-101) */
int chs_integrate_depth_color(chs_map *map, const chs_integrator *integ, const float *depth, int mem,
                              const float pose[12], const chs_camera *cam, const uint8_t *color, int channels,
                              const float color_pose[12], const chs_camera *color_cam);
/* One frame of a batch. Pointers live in the memory space named by `mem` of the call. */
typedef struct
{
    const float *depth;            /* width*height float metres (ignored when depth_mm is set) */
    const uint16_t *depth_mm;      /* or: width*height uint16 millimetres, the ROS 16UC1 depth encoding. Converted on the device exactly as
                                      chisel_ros does on the host, (1.0f / 1000.0f) * value (CR Conversions.h:141-152): half the bytes to move */
    const uint8_t *color;          /* colour path only: color_cam.width*height*channels */
    const float *trunc_per_pixel;  /* CHS_TRUNC_PER_PIXEL only */
    float pose[12];
    float color_pose[12];          /* colour path only */
} chs_frame;
/* n consecutive frames of one stream (same image size, intrinsics and integrator) in one call: n calls of
 * Chisel::IntegrateDepthScan (color_cam == NULL) or IntegrateDepthScanColor, in order, with bit-identical results. Groups of
 * up to 16 frames run through the fused multi-frame kernels (every touched voxel is read and written once per group; see
 * DESIGN.md section 5); frames whose colour camera differs from the depth camera are integrated one by one. */
int chs_integrate_batch(chs_map *map, const chs_integrator *integ, int n_frames, const chs_frame *frames, int mem,
                        const chs_camera *cam, int channels, const chs_camera *color_cam);
/* per-frame counters of the last chs_integrate_batch call; *n = its frame count. Synchronises. */
int chs_get_batch_stats(chs_map *map, chs_frame_stats *out, int cap, int *n);
/* Pipelined use: ticket of the last chs_integrate_batch call, and a wait for ONE call (later calls stay in flight; the
 * counters of the last 4 calls are kept). chs_integrate_batch itself returns as soon as its host buffers have been copied. */
int chs_last_batch_ticket(chs_map *map, int64_t *ticket);
int chs_wait_batch(chs_map *map, int64_t ticket, chs_frame_stats *out, int cap, int *n);
int chs_get_frame_stats(chs_map *map, chs_frame_stats *out);   /* synchronises */
int chs_get_timings(chs_map *map, chs_timings *out);           /* synchronises */

/* Re-mesh every chunk of the dirty set that exists, then clear the dirty set. Synchronises (the vertex
 * count decides the output allocation). Results stay on the device until downloaded. */
int chs_update_meshes(chs_map *map);
int chs_mesh_counts_last(chs_map *map, chs_mesh_counts *out);
/* ids [3*n_chunks]; vert_offsets, grid_offsets [n_chunks+1] (prefix sums, in vertices / grid points);
 * vertices, normals, colors [3*n_vertices] (colors may be NULL); grids [3*n_grids]. Any pointer may be NULL. */
int chs_download_meshes(chs_map *map, int32_t *ids, int64_t *vert_offsets, int64_t *grid_offsets,
                        float *vertices, float *normals, float *colors, float *grids);

int chs_num_chunks(chs_map *map, int64_t *n);                  /* synchronises */
int chs_chunk_ids(chs_map *map, int32_t *ids, int64_t cap);    /* pool order */
int chs_has_chunk(chs_map *map, const int32_t id[3], int *found); /* ChunkManager::HasChunk */
/* sdf, weight [V]; rgbw [4V] (r, g, b, colour weight) or NULL. CHS_ERR_NOT_FOUND if the chunk is absent. */
int chs_download_chunk(chs_map *map, const int32_t id[3], float *sdf, float *weight, uint8_t *rgbw);
/* Whole map in one transfer, pool order: ids [3n], sdf/weight [n*V], rgbw [n*4V] or NULL. */
int chs_download_all(chs_map *map, int64_t cap_chunks, int32_t *ids, float *sdf, float *weight, uint8_t *rgbw);
/* Chunk export / import and explicit dirty set: the primitives behind sharded meshing (ghost copies of neighbour chunks
 * owned by other ranks) and map checkpoint / resume (the reference's GetAllChunks wire format, CR Serialization.h:31-84, is
 * broken). Host buffers: sdf, weight [n*V]; rgbw [n*4V] or NULL. */
int chs_export_chunks(chs_map *map, int64_t n, const int32_t *ids, uint8_t *found, float *sdf, float *weight, uint8_t *rgbw);
int chs_import_chunks(chs_map *map, int64_t n, const int32_t *ids, const float *sdf, const float *weight, const uint8_t *rgbw);
int chs_set_dirty(chs_map *map, int64_t n, const int32_t *ids);
/* Map checkpoint on disk and resume (stands in for the reference's chunk wire format, CR/include/chisel_ros/
 * Serialization.h:31-84 + CR/msg/ChunkMessage.msg, whose bit packing is broken). One file: header (magic "CHSMAP01", version, chunk
 * size, resolution, colour flag, counts), then ids, sdf, weight, rgbw and the dirty set as flat little-endian arrays. A map that
 * loads it continues exactly like the map that wrote it (same voxels, same dirty set); chs_load_map replaces the map's contents. */
int chs_save_map(chs_map *map, const char *path);
int chs_load_map(chs_map *map, const char *path);
int chs_num_dirty(chs_map *map, int64_t *n);                   /* synchronises */
int chs_dirty_ids(chs_map *map, int32_t *ids, int64_t cap);

/* ---- multi-GPU: one process per GPU, chunk-ownership shards (chs_config.rank / world, chs_owner), NCCL over NVLink ----
 * Replaces, for a map spread over several GPUs: the frame hand-over of chisel_ros (CR/src/ChiselServer.cpp:297-367: every agent's
 * frame lands in ONE map in callback order), Chisel::meshesToUpdate as one set (OC/include/open_chisel/Chisel.h:221-228), the
 * neighbour-chunk reads of meshing (OC/src/ChunkManager.cpp:319-364, 449-474) and the mesh consumer's view of all meshes
 * (CR/src/ChiselServer.cpp:613-716). The library loads libnccl.so.2 on first use; nothing NCCL is needed on one GPU.
 * All of these are COLLECTIVE: every rank of the map's world calls them in the same order. */
#define CHS_NCCL_ID_BYTES 128
int chs_comm_unique_id(uint8_t id[CHS_NCCL_ID_BYTES]);      /* ncclGetUniqueId: one rank calls it and hands the bytes to the others out of band */
int chs_comm_init(chs_map *map, const uint8_t id[CHS_NCCL_ID_BYTES]);   /* ncclCommInitRank(cfg.world, id, cfg.rank) */
int chs_comm_attach(chs_map *map, void *nccl_comm);         /* or: adopt the caller's ncclComm_t (same size and rank as the map's world) */
int chs_comm_destroy(chs_map *map);
/* One step of n_total <= 16 frames (n_total % world == 0), e.g. the frames of all agents of one time step, in arrival order.
 * frames[0 .. n_total): poses of ALL frames; image pointers only for the frames this rank ingests,
 * [rank * n_total / world, (rank + 1) * n_total / world) (the other entries' pointers are ignored). Every rank PUSHES its images
 * over NVLink into exchange arenas of all ranks (peer memory mapped through CUDA IPC at the first call; flag words instead of a
 * collective; NCCL all-gather where IPC is not available or with CHS_NCCL_EXCHANGE=1), a step ahead of the kernels that read
 * them, and the step is integrated as one fused batch, every rank updating the chunks it owns: the union of the ranks' maps equals the single-GPU map of
 * chs_integrate_batch on the same frames. Colour camera and pose must equal the depth ones; constant / computed truncators only. */
int chs_integrate_batch_distributed(chs_map *map, const chs_integrator *integ, int n_total, const chs_frame *frames, int mem,
                                    const chs_camera *cam, int channels, const chs_camera *color_cam);
/* The dirty set of every rank becomes the union over ranks. */
int chs_comm_sync_dirty(chs_map *map);
/* Distributed Chisel::UpdateMeshes: dirty-set union; device-side exchange of the chunks inside the 27-neighbourhood of the dirty
 * set ("ghost" copies, all-gather-v over NVLink); every rank re-meshes the dirty chunks it owns; the meshes are gathered on
 * `root`, where chs_mesh_counts_last / chs_download_meshes then return the union (elsewhere: the rank's own part). The dirty
 * set is cleared. Index for index the meshes of a single-GPU map (one documented exception: DESIGN.md section 8, quirk Q9). */
int chs_update_meshes_distributed(chs_map *map, int root);

/* Device self-test of the two range-check-free arithmetic forms the fused kernels use (cvids_b200/csrc/integrate_device.cuh)
 * against the IEEE intrinsics: the reciprocal over EVERY binary32 value of its guarded range, the quotient over div_pairs
 * pseudo-random operand pairs. out = {rcp mismatches, rcp values tested, div mismatches, div pairs tested}. */
int chs_selftest_arithmetic(int64_t div_pairs, int64_t out[4]);

/* Page-locked host memory for frame queues (the facade's batching queue lives in it): H2D copies from it run at full PCIe
 * speed and asynchronously. NULL on failure. */
void *chs_host_alloc(size_t bytes);
void chs_host_free(void *p);

/* Frame ingestion of the collaborative server, the step right before the path (SURVEY.md 8(f) item 2;
 * server_pose_graph/src/collaborative_server_system.cpp:213-276): cv::resize (INTER_LINEAR, half-pixel centres, OpenCV's float path) of
 * the float depth map to dst_w x dst_h, then NaN for every depth outside [valid_min, valid_max] (0.1 m / 20 m there). Equal sizes:
 * masking only. CHS_MEM_HOST (copied through the device, synchronous) or CHS_MEM_DEVICE (in place on the map's stream: the result
 * can go straight into chs_integrate_*, no host hop). The 16UC1 millimetre conversion of chisel_ros (CR Conversions.h:141-152) is
 * chs_frame.depth_mm. */
int chs_ingest_depth(chs_map *map, const float *src, int src_w, int src_h, float *dst, int dst_w, int dst_h, float valid_min, float valid_max, int mem);

/* Device-side frame queue for callers that receive frames one at a time and reuse their image buffers (the facade's
 * Chisel::IntegrateDepthScan[Color] with frame batching, CR/src/ChiselServer.cpp:285-295): device memory on the map's device,
 * and an upload that returns as soon as the HOST buffer may be reused (one PCIe copy on a dedicated stream -- no staging memcpy
 * on the host). Frames uploaded this way are handed to chs_integrate_batch with CHS_MEM_DEVICE_ASYNC. */
void *chs_device_alloc(chs_map *map, size_t bytes);
void chs_device_free(chs_map *map, void *p);
int chs_upload(chs_map *map, void *dst_device, const void *src_host, size_t bytes);

/* Host-side exact restatements the facade needs (no device work). */
int chs_frustum(const float pose[12], const chs_camera *cam, float corners[24], float lines[72], float planes[24]);
int chs_candidate_ids(int chunk_size, float resolution, const float pose[12], const chs_camera *cam,
                      int32_t *ids, int64_t cap, int64_t *n);
float chs_truncation(int trunc_kind, float trunc_param, float depth);
uint32_t chs_owner(int32_t x, int32_t y, int32_t z);

#ifdef __cplusplus
}
#endif
#endif /* CHISEL_B200_H */
