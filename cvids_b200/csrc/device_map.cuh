// device_map.cuh -- device-side layout of the chunked TSDF map and the per-frame parameter block.
//
// HBM layout:
//   * chunk hash table: open addressing, linear probing; 64-bit packed chunk ID -> pool slot.
//     Replaces std::unordered_map<ChunkID, ChunkPtr, ChunkHasher> (OC ChunkManager.h:40-53).
//   * chunk pool: slabs of kSlabChunks chunks. Per chunk, V consecutive float2 {sdf, weight}
//     (x fastest, then y, then z: Chunk::GetVoxelID, OC Chunk.h:81-84) and, with colour, V consecutive
//     uchar4 {r, g, b, colour weight}. A warp walking x touches 256 contiguous bytes of distance state.
//     Replaces the reference's two 16-byte-per-voxel AoS vectors (DistVoxel / ColorVoxel carry a vptr, Q11).
//   * dirty set: hash set + append list of packed chunk IDs (Chisel::meshesToUpdate, OC Chisel.h:221-228).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace chs
{

constexpr int kSlabChunksLog2 = 10;
constexpr int kSlabChunks = 1 << kSlabChunksLog2;          // 1024 chunks per slab (32 MiB of distance state at 16^3)
constexpr int kMaxSlabs = 1 << 13;
constexpr unsigned long long kEmptyKey = ~0ull;
constexpr int kIdBias = 1 << 20;                           // chunk IDs must lie in [-2^20, 2^20)
constexpr int kHizLevels = 8;                              // tiles of 8, 16, 32, 64 pixels (per 64x64 block) and 128 .. 1024 pixels (built by the last block)
constexpr int kCounterSlots = 32;                          // per-frame voxel counters are spread over this many addresses

enum : int
{
    kErrPoolFull = 1,
    kErrHashFull = 2,
    kErrDirtyFull = 4,
    kErrWorkFull = 8,
    kErrMeshFull = 16,
    kErrPeerTimeout = 32            // peer-memory frame exchange: a rank's images (or its release of a staging set) did not arrive in time
};

__host__ __device__ inline unsigned long long pack_id(int x, int y, int z)
{
    return ((unsigned long long)(unsigned)(x + kIdBias) << 42) | ((unsigned long long)(unsigned)(y + kIdBias) << 21) |
           (unsigned long long)(unsigned)(z + kIdBias);
}
__host__ __device__ inline void unpack_id(unsigned long long k, int *x, int *y, int *z)
{
    *x = (int)((k >> 42) & 0x1FFFFF) - kIdBias;
    *y = (int)((k >> 21) & 0x1FFFFF) - kIdBias;
    *z = (int)(k & 0x1FFFFF) - kIdBias;
}
// splitmix64 finaliser: probe start and (with % world) chunk ownership
__host__ __device__ inline unsigned long long mix64(unsigned long long h)
{
    h ^= h >> 30;
    h *= 0xbf58476d1ce4e5b9ull;
    h ^= h >> 27;
    h *= 0x94d049bb133111ebull;
    h ^= h >> 31;
    return h;
}
__host__ __device__ inline unsigned owner_hash(int x, int y, int z) { return (unsigned)(mix64(pack_id(x, y, z)) >> 32); }

struct Counters
{
    // persistent
    int n_chunks;
    int n_dirty;
    int error_flags;
    // per frame (zeroed by frame_prepare)
    int unit_count;                // brick units of existing chunks
    int new_count;                 // new-chunk candidates
    int candidates;
    int n_new;
    int updated_chunks;
    int tickets;                   // CTAs of the two integrate kernels that have finished (the last one takes the snapshot)
    int pad;
    unsigned long long n_upd[kCounterSlots], n_carve[kCounterSlots], n_col[kCounterSlots];
    // per re-mesh
    unsigned long long mesh_verts, mesh_grids;
    int mesh_chunks;
    int pad2;
};

// What the host needs to know about a finished frame; written into a pinned ring slot by the last CTA of the frame as
// four 16-byte lines, EACH carrying the frame id in its first word: a slot is valid for frame f when all four ids equal f,
// so no system-scope fence or flag ordering is needed (a 16-byte aligned store is one transaction on the bus).
struct HostSnapshot
{
    int id0, n_chunks, n_dirty, error_flags;
    int id1, unit_count, new_count, candidates;
    int id2, n_new, updated_chunks, n_carve;
    int id3, n_col, n_upd_lo, n_upd_hi;
};

struct DeviceMap
{
    unsigned long long *keys;      // chunk hash keys
    int *vals;                     // chunk hash values (pool slot)
    unsigned mask;                 // table size - 1
    float2 **dist_slabs;           // [kMaxSlabs] device pointers
    uchar4 **color_slabs;
    int *slot_ids;                 // [capacity*3] slot -> chunk ID
    unsigned long long *brick_flags; // [capacity] bit b: 8^3 brick b may hold a voxel with weight > 0 && sdf < 1e-5 (carvable)
    int *slot_epoch;               // [capacity] id of the last frame that updated the chunk (dirty marking runs once per chunk and frame)
    int capacity;                  // chunks the allocated slabs can hold
    unsigned long long *dirty_keys;
    unsigned dirty_mask;
    unsigned long long *dirty_list;
    int dirty_cap;
    Counters *ctr;
    int cs, V;
    float res, half;               // half = res * 0.5f (ChunkManager.cpp:52)
    int use_color;
    int rank, world;
};

__device__ __forceinline__ float2 *dist_ptr(const DeviceMap &m, int slot)
{
    return m.dist_slabs[slot >> kSlabChunksLog2] + (size_t)(slot & (kSlabChunks - 1)) * m.V;
}
__device__ __forceinline__ uchar4 *color_ptr(const DeviceMap &m, int slot)
{
    return m.color_slabs[slot >> kSlabChunksLog2] + (size_t)(slot & (kSlabChunks - 1)) * m.V;
}

__device__ __forceinline__ int hash_lookup(const DeviceMap &m, unsigned long long key)
{
    unsigned i = (unsigned)mix64(key) & m.mask;
    while (true)
    {
        const unsigned long long k = m.keys[i];
        if (k == key)
            return m.vals[i];
        if (k == kEmptyKey)
            return -1;
        i = (i + 1) & m.mask;
    }
}

// Insert a key that is known to be absent (each candidate ID is unique within a frame).
__device__ __forceinline__ void hash_insert_new(const DeviceMap &m, unsigned long long key, int slot)
{
    unsigned i = (unsigned)mix64(key) & m.mask;
    while (true)
    {
        const unsigned long long prev = atomicCAS(&m.keys[i], kEmptyKey, key);
        if (prev == kEmptyKey)
        {
            m.vals[i] = slot;
            return;
        }
        i = (i + 1) & m.mask;
    }
}

// Dirty set: insert-if-absent; appends to the list on first insertion.
__device__ __forceinline__ void dirty_insert(const DeviceMap &m, unsigned long long key)
{
    unsigned i = (unsigned)mix64(key) & m.dirty_mask;
    for (unsigned probes = 0; probes <= m.dirty_mask; probes++)
    {
        const unsigned long long k = m.dirty_keys[i];
        if (k == key)
            return;
        if (k == kEmptyKey)
        {
            const unsigned long long prev = atomicCAS(&m.dirty_keys[i], kEmptyKey, key);
            if (prev == kEmptyKey)
            {
                const int pos = atomicAdd(&m.ctr->n_dirty, 1);
                if (pos < m.dirty_cap)
                    m.dirty_list[pos] = key;
                else
                    atomicOr(&m.ctr->error_flags, kErrDirtyFull);
                return;
            }
            if (prev == key)
                return;
        }
        i = (i + 1) & m.dirty_mask;
    }
    atomicOr(&m.ctr->error_flags, kErrDirtyFull);
}

struct CameraDev
{
    float R[9];       // row-major rotation, camera -> world
    float t[3];
    float fx, fy, cx, cy;
    int W, H;
    float Wf, Hf;
};

struct FrameParams
{
    CameraDev cam;            // depth camera + pose
    CameraDev ccam;           // colour camera + pose
    const float *depth;       // W*H metres
    const uint16_t *depth_u16;// W*H millimetres (ROS 16UC1) or nullptr: frame_prepare converts it into `depth`, (1.0f / 1000.0f) * value
                              // as ROSImgToDepthImg does (CR Conversions.h:141-152)
    const float *trunc_img;   // W*H or nullptr (constant truncator)
    const uint8_t *color;     // cW*cH*channels, as the caller handed it over
    unsigned *color_packed;   // cW*cH packed r | g << 8 | b << 16, written by frame_prepare (ColorImage::At, OC ColorImage.h:61-101)
    int channels;
    int color_path;           // 0: Integrate (ProjectionIntegrator.h:51-99), 1: IntegrateColor (:101-183)
    int trunc_kind;
    float trunc_param;
    float diag;               // float(2.0 * sqrt(3.0f) * res), evaluated on the host in double like the reference
    float carve_dist;
    int carve;
    float weight;             // ConstantWeighter weight
    float depth_cutoff;       // 50 (depth path) / 100 (colour path)
    float sdf_carve_max;      // smallest float T with double(T) >= 1e-5: (double)sdf < 1e-5  <=>  sdf < T
    int same_cam;             // colour pose + intrinsics bit-identical to the depth ones: reuse the projection
    float wu_const;           // constant truncator: weight / (5 * trunc), formed on the host in binary32
    // candidate enumeration
    int lo[3], n[3];
    float planes[6][4];       // the reference's (quirky) frustum planes, for the exact Frustum::Intersects
    float view_planes[5][4];  // the camera's real view pyramid in world space (left, right, top, bottom with a 3 pixel margin, z >= 0):
                              // n.x + d >= 0 inside. Culling only: a box entirely outside one of them projects off the image.
    float2 *hiz[kHizLevels];  // {min lo, max hi} per tile
    int hizW[kHizLevels], hizH[kHizLevels];
    int hiz_levels;           // levels in use: the coarsest one has at most 3x3 tiles (or is level kHizLevels - 1)
    int hiz_blocks;           // 64x64 pixel blocks of the image = CTAs that build levels 0..3
    int *hiz_ticket;          // counts finished blocks of this frame; the last one builds levels 4.. and resets it
    int4 *units;              // brick units of existing chunks: {x, y, z, slot | brick << 24}
    int units_cap;
    int4 *news;               // new-chunk candidates: {x, y, z, -1}
    int news_cap;
    int cand_stride;          // multiplicative permutation of the candidate enumeration (coprime to the box size)
    int frame_id;             // > 0, increases by one per integrated frame
    int total_ctas;           // CTAs of integrate_new_chunks + integrate_bricks of this frame
    HostSnapshot *host_slot;  // pinned, device-mapped
};

} // namespace chs
