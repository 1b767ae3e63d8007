// capi.cu -- map lifecycle, capacity management and the C ABI declared in include/chisel_b200.h.
//
// Host orchestration of one frame (replaces Chisel::IntegrateDepthScan[Color], OC Chisel.h:59-213):
//   host:   frustum -> candidate ID box (host_geometry.h, exact)       [Chisel.h:64-68 / :119-123]
//   device: frame_prepare -> chunk_candidates -> integrate             [Chisel.h:70-107 / :133-207]
// Chunks that the reference would create and immediately garbage-collect are never materialised, so there
// is no erase path; the pool is a bump allocator and the hash table is insert-only between resets.
// Capacities are grown on the host BEFORE a frame from conservative upper bounds (candidate count), using
// the stream-ordered allocator, so no kernel can overflow a table and no per-frame host sync is needed.
#include <cuda_runtime.h>

#include <algorithm>
#include <immintrin.h>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <array>
#include <chrono>
#include <deque>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/chisel_b200.h"
#include "device_map.cuh"
#include "host_geometry.h"
#include "kernels.h"
#include "tma.cuh"

namespace chs
{

static thread_local std::string g_last_error;

static int fail(int code, const std::string &msg)
{
    g_last_error = msg;
    return code;
}

#define CHS_CUDA(expr)                                                                                      \
    do                                                                                                      \
    {                                                                                                       \
        cudaError_t _e = (expr);                                                                            \
        if (_e != cudaSuccess)                                                                              \
            return fail(CHS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));                  \
    } while (0)

__global__ void fill_u64_kernel(unsigned long long *p, size_t n, unsigned long long v)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = v;
}
void launch_fill_u64(unsigned long long *p, size_t n, unsigned long long v, cudaStream_t st)
{
    if (n == 0)
        return;
    const int grid = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
    fill_u64_kernel<<<grid, 256, 0, st>>>(p, n, v);
}

__global__ void fill_i32_kernel(int *p, size_t n, int v)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = v;
}
static void launch_fill_i32(int *p, size_t n, int v, cudaStream_t st)
{
    if (n == 0)
        return;
    fill_i32_kernel<<<(int)std::min<size_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(p, n, v);
}

// Pool invariant: every slot at or above n_chunks holds the state Chunk::Chunk produces -- {sdf 99999, weight 0} and colour 0
// (OC Chunk.cpp:33-48, DistVoxel.cpp:29-33, ColorVoxel.cpp) -- so that the fused kernels create a chunk without writing it.
__global__ void fill_pool_kernel(DeviceMap map, int firstSlot, int nSlots)
{
    const size_t n = (size_t)nSlots * map.V;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    {
        const int slot = firstSlot + (int)(i / map.V);
        const int v = (int)(i % map.V);
        dist_ptr(map, slot)[v] = make_float2(99999.0f, 0.0f);
        if (map.use_color)
            color_ptr(map, slot)[v] = make_uchar4(0, 0, 0, 0);
    }
}
static void launch_fill_pool(const DeviceMap &map, int firstSlot, int nSlots, cudaStream_t st)
{
    if (nSlots > 0)
        fill_pool_kernel<<<148 * 8, 256, 0, st>>>(map, firstSlot, nSlots);
}

// Re-insert every pool slot into a freshly emptied (larger) table.
__global__ void rebuild_hash_kernel(DeviceMap map)
{
    const int n = map.ctr->n_chunks;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x)
        hash_insert_new(map, pack_id(map.slot_ids[3 * s], map.slot_ids[3 * s + 1], map.slot_ids[3 * s + 2]), s);
}
void launch_rebuild_hash(const DeviceMap &map, int, cudaStream_t st) { rebuild_hash_kernel<<<148, 256, 0, st>>>(map); }

__global__ void rebuild_dirty_kernel(DeviceMap map)
{
    const int n = min(map.ctr->n_dirty, map.dirty_cap);
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x)
    {
        const unsigned long long key = map.dirty_list[s];
        unsigned i = (unsigned)mix64(key) & map.dirty_mask;
        while (atomicCAS(&map.dirty_keys[i], kEmptyKey, key) != kEmptyKey)
            i = (i + 1) & map.dirty_mask;
    }
}
void launch_rebuild_dirty(const DeviceMap &map, int, cudaStream_t st) { rebuild_dirty_kernel<<<148, 256, 0, st>>>(map); }

struct InFlight
{
    int frameId;            // the snapshot kernel writes this id into the ring slot when the frame is done
    long long newBound;     // upper bound of chunks this frame may add
    long long dirtyBound;   // upper bound of dirty IDs this frame may add
    int slot;               // index into the pinned counter ring
    bool batch = false;     // a fused multi-frame batch: frameId is the batch id, slot indexes the batch snapshot ring
    int batchIndex = -1;    // index of this launch's first frame within its chs_integrate_batch call (-1: not part of one)
    int callId = 0;         // the chs_integrate_batch call it belongs to
};

} // namespace chs

using namespace chs;

struct chs_map
{
    chs_config cfg;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool ownStream = false;
    DeviceMap dm{};
    std::vector<float2 *> distSlabs;
    std::vector<uchar4 *> colorSlabs;
    size_t hashSize = 0, dirtySize = 0;
    // frame staging
    float *dDepth = nullptr, *dTrunc = nullptr;
    uint8_t *dColor = nullptr;
    unsigned *dColorPacked = nullptr;
    size_t colorPackedCap = 0;
    float2 *dHiz = nullptr;
    int4 *dUnits = nullptr, *dNews = nullptr;
    size_t depthCap = 0, truncCap = 0, colorCap = 0, hizCap = 0, unitsCap = 0, newsCap = 0;
    int frameId = 0;
    cudaEvent_t h2dDone = nullptr;
    // counters
    Counters *dCtr = nullptr;
    Counters *hCtr = nullptr;           // pinned: staging for direct reads of the device counters
    HostSnapshot *hSnap = nullptr;      // pinned, device-mapped ring written by the last CTA of every frame
    static constexpr int kRing = 16;
    int ringNext = 0;
    std::deque<InFlight> inflight;
    FrameGraph *frameGraph = nullptr;
    long long knownChunks = 0, knownDirty = 0;
    HostSnapshot lastFrame{};
    bool haveFrame = false;
    // fused multi-frame path (integrate_batch_impl.cuh)
    // Two staging sets, used alternately: while the kernels of batch i read set i & 1 on the map's stream, the H2D copies and
    // the prepare kernel of batch i + 1 fill the other set on the copy stream.
    struct BatchSet
    {
        FrameParams *dFrames = nullptr;
        BatchCounters *dBctr = nullptr;
        unsigned long long *dTimeline = nullptr;   // [kTimelineStamps]
        float *depth = nullptr, *trunc = nullptr;
        uint16_t *depthMm = nullptr;
        size_t depthMmCap = 0;
        uint8_t *color = nullptr;
        unsigned *packed = nullptr;
        float2 *hiz = nullptr;
        size_t depthCap = 0, truncCap = 0, colorCap = 0, packedCap = 0, hizCap = 0;
        cudaEvent_t copied = nullptr, prepared = nullptr, released = nullptr, packDone = nullptr, fork = nullptr;
        bool used = false;
    } bset[2];
    cudaStream_t copyStream = nullptr;
    cudaStream_t pushStream = nullptr;            // peer-memory exchange of device frames: pushes queue up here, beside everything else
    cudaStream_t uploadStream = nullptr;          // chs_upload
    cudaEvent_t callEvent = nullptr;
    int *dHizTickets = nullptr;                // [2 * kMaxBatch + 1] self-resetting block counters of frame_prepare
    HostBatchSnapshot *hBatchSnap = nullptr;   // pinned, device-mapped ring [kRing]
    FrameParams *hFrameTables = nullptr;       // pinned ring [kRing][kMaxBatch]: staging of the per-batch frame table (at most 4 batches are in flight)
    int frameTableNext = 0;
    unsigned long long *dSlotBatch = nullptr;  // [capacity]
    int batchId = 0, batchRingNext = 0;
    long long lastBricksSpanNs = 0;
    // per-frame counters of the most recent chs_integrate_batch calls (a call is identified by its ticket)
    struct CallStats
    {
        int id = 0, pending = 0;
        std::vector<chs_frame_stats> st;
    } callStats[4];
    int callId = 0;
    // profiling
    bool profiling = false;
    bool timeline = false;                     // chs_set_profiling bit 1: device timeline of every fused batch (no events)
    std::deque<std::array<long long, kTimelineStamps>> timelineHist;   // the most recent batches (at most 256), oldest first
    cudaEvent_t evt[8] = {};
    bool frameTimed = false, meshTimed = false;
    // host mirror of slot -> id (extended lazily; slots are never recycled between resets)
    std::vector<int32_t> hostIds;
    std::unordered_map<unsigned long long, int> hostIndex;
    // meshing
    int *dMeshSlots = nullptr, *dTriCounts = nullptr, *dGridCounts = nullptr;
    long long *dVertOffsets = nullptr, *dGridOffsets = nullptr;
    size_t meshChunkCap = 0;
    float *dVerts = nullptr, *dNormals = nullptr, *dColors = nullptr, *dGrids = nullptr;
    unsigned char *dCfgScratch = nullptr;
    size_t cfgScratchCap = 0;
    long long vertCap = 0, gridCap = 0;
    chs_mesh_counts lastMesh{};
    int lastMeshChunks = 0;
    // multi-GPU (capi_comm.inc)
    void *comm = nullptr;                      // ncclComm_t
    bool ownComm = false;
    chs_map *ghost = nullptr;                  // scratch one-rank map the distributed re-mesh runs on
    long long *dCommScratch = nullptr;
    size_t commScratchCap = 0;
    int distFirst = 0, distCount = 0;          // distributed batch in progress: the frames [distFirst, distFirst + distCount) are ingested here
    // Peer-memory frame exchange (capi_comm.inc): this rank's exchange arena -- a header of flag words and two staging sets for the
    // images of a whole step -- is mapped by every other rank over CUDA IPC; every rank PUSHES the frames it ingests into all
    // arenas with one kernel (NVLink stores), no collective on the hot path.
    static constexpr int kMaxPeers = 16;
    static constexpr int kArenaSets = 3;
    struct PeerArena
    {
        char *base = nullptr;                  // cudaMalloc: [header 4 KB][set 0][set 1][set 2]
        size_t setBytes = 0, depthOff = 0, mmOff = 0, colorOff = 0, hizOff = 0;    // layout of a set
        size_t depthCap = 0, mmCap = 0, colorCap = 0, hizCap = 0;      // bytes per set
        char *peer[kMaxPeers] = {};            // the arenas of all ranks as mapped here (peer[rank] == base)
        bool tried = false, ok = false;
        unsigned step = 0;                     // distributed steps pushed so far (the same number on every rank)
    } arena;
    struct Gathered                            // the root's copy of all ranks' meshes of the last distributed re-mesh
    {
        int *ids = nullptr;
        long long *vertOffsets = nullptr, *gridOffsets = nullptr;
        float *verts = nullptr, *normals = nullptr, *colors = nullptr, *grids = nullptr;
        size_t idsCap = 0, voCap = 0, goCap = 0, vCap = 0, nCap = 0, cCap = 0, gCap = 0;
        long long nChunks = 0, nVerts = 0, nGrids = 0;
    } gathered;
    bool meshGathered = false;                 // chs_download_meshes serves the gathered union (root of a distributed re-mesh)
    bool meshFromGhost = false;                // ... or this rank's part, held by the ghost map (other ranks)
};

namespace chs
{

// The library's stream-ordered allocations come from its OWN memory pool (one per device, shared by the maps on it), which keeps
// freed memory cached instead of returning it to the OS at every synchronisation. The device's default pool -- which the rest of
// the process (PyTorch, the caller) may use -- is left alone.
static cudaMemPool_t g_pools[64] = {};
static int library_pool(cudaMemPool_t *out)
{
    int dev = 0;
    CHS_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64)
        return fail(CHS_ERR_INVALID, "device ordinal out of range");
    if (!g_pools[dev])
    {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        cudaMemPool_t pool = nullptr;
        CHS_CUDA(cudaMemPoolCreate(&pool, &props));
        unsigned long long thr = ~0ull;
        CHS_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
        g_pools[dev] = pool;
    }
    *out = g_pools[dev];
    return CHS_OK;
}

static int alloc_async(void **p, size_t bytes, cudaStream_t st)
{
    cudaMemPool_t pool = nullptr;
    int rc = library_pool(&pool);
    if (rc)
        return rc;
    CHS_CUDA(cudaMallocFromPoolAsync(p, std::max<size_t>(bytes, 256), pool, st));
    return CHS_OK;
}

template <typename T>
static int grow_buffer(T **buf, size_t *cap, size_t need, cudaStream_t st)
{
    if (need <= *cap)
        return CHS_OK;
    size_t ncap = std::max(need, *cap + *cap / 2);
    if (*buf)
        CHS_CUDA(cudaFreeAsync(*buf, st));
    void *p = nullptr;
    int rc = alloc_async(&p, ncap * sizeof(T), st);
    if (rc)
        return rc;
    *buf = (T *)p;
    *cap = ncap;
    return CHS_OK;
}

static size_t pow2_at_least(size_t n)
{
    size_t p = 1024;
    while (p < n)
        p <<= 1;
    return p;
}

// Add slabs until the pool holds `chunks`; grows the slot->id array alongside.
static int ensure_pool(chs_map *m, long long chunks)
{
    if (chunks <= m->dm.capacity)
        return CHS_OK;
    const int V = m->dm.V;
    const size_t needSlabs = (size_t)((chunks + kSlabChunks - 1) / kSlabChunks);
    if (needSlabs > (size_t)kMaxSlabs)
        return fail(CHS_ERR_CAPACITY, "chunk pool would exceed kMaxSlabs");
    // failure-atomic: the slab vectors always describe exactly dm.capacity chunks; slabs allocated by a call that fails later are
    // given back, so a retry starts from a consistent state
    const size_t oldSlabs = (size_t)m->dm.capacity / kSlabChunks;
    auto roll_back = [&]()
    {
        while (m->distSlabs.size() > oldSlabs)
        {
            cudaFreeAsync(m->distSlabs.back(), m->stream);
            m->distSlabs.pop_back();
        }
        while (m->colorSlabs.size() > (m->cfg.use_color ? oldSlabs : 0))
        {
            cudaFreeAsync(m->colorSlabs.back(), m->stream);
            m->colorSlabs.pop_back();
        }
    };
    roll_back();                                                     // leftovers of an earlier failed call
    for (size_t s = oldSlabs; s < needSlabs; s++)
    {
        void *p = nullptr;
        int rc = alloc_async(&p, sizeof(float2) * (size_t)kSlabChunks * V, m->stream);
        if (rc)
        {
            roll_back();
            return rc;
        }
        m->distSlabs.push_back((float2 *)p);
        if (m->cfg.use_color)
        {
            rc = alloc_async(&p, sizeof(uchar4) * (size_t)kSlabChunks * V, m->stream);
            if (rc)
            {
                roll_back();
                return rc;
            }
            m->colorSlabs.push_back((uchar4 *)p);
        }
    }
    CHS_CUDA(cudaMemcpyAsync(m->dm.dist_slabs + oldSlabs, m->distSlabs.data() + oldSlabs, sizeof(float2 *) * (needSlabs - oldSlabs),
                             cudaMemcpyHostToDevice, m->stream));
    if (m->cfg.use_color)
        CHS_CUDA(cudaMemcpyAsync(m->dm.color_slabs + oldSlabs, m->colorSlabs.data() + oldSlabs, sizeof(uchar4 *) * (needSlabs - oldSlabs),
                                 cudaMemcpyHostToDevice, m->stream));
    // the pageable source vectors must outlive the copy: cudaMemcpyAsync from pageable memory stages before returning
    const long long newCap = (long long)needSlabs * kSlabChunks;
    void *ids = nullptr;
    int rc = alloc_async(&ids, sizeof(int) * 3 * (size_t)newCap, m->stream);
    if (rc)
        return rc;
    if (m->dm.slot_ids)
    {
        CHS_CUDA(cudaMemcpyAsync(ids, m->dm.slot_ids, sizeof(int) * 3 * (size_t)m->dm.capacity, cudaMemcpyDeviceToDevice, m->stream));
        CHS_CUDA(cudaFreeAsync(m->dm.slot_ids, m->stream));
    }
    m->dm.slot_ids = (int *)ids;
    void *flags = nullptr;
    rc = alloc_async(&flags, sizeof(unsigned long long) * (size_t)newCap, m->stream);
    if (rc)
        return rc;
    if (m->dm.brick_flags)
    {
        CHS_CUDA(cudaMemcpyAsync(flags, m->dm.brick_flags, sizeof(unsigned long long) * (size_t)m->dm.capacity, cudaMemcpyDeviceToDevice, m->stream));
        CHS_CUDA(cudaFreeAsync(m->dm.brick_flags, m->stream));
    }
    m->dm.brick_flags = (unsigned long long *)flags;
    void *epoch = nullptr;
    rc = alloc_async(&epoch, sizeof(int) * (size_t)newCap, m->stream);
    if (rc)
        return rc;
    if (m->dm.slot_epoch)
    {
        CHS_CUDA(cudaMemcpyAsync(epoch, m->dm.slot_epoch, sizeof(int) * (size_t)m->dm.capacity, cudaMemcpyDeviceToDevice, m->stream));
        CHS_CUDA(cudaFreeAsync(m->dm.slot_epoch, m->stream));
    }
    m->dm.slot_epoch = (int *)epoch;
    void *sb = nullptr;
    rc = alloc_async(&sb, sizeof(unsigned long long) * (size_t)newCap, m->stream);
    if (rc)
        return rc;
    CHS_CUDA(cudaMemsetAsync(sb, 0, sizeof(unsigned long long) * (size_t)newCap, m->stream));
    if (m->dSlotBatch)
    {
        CHS_CUDA(cudaMemcpyAsync(sb, m->dSlotBatch, sizeof(unsigned long long) * (size_t)m->dm.capacity, cudaMemcpyDeviceToDevice, m->stream));
        CHS_CUDA(cudaFreeAsync(m->dSlotBatch, m->stream));
    }
    m->dSlotBatch = (unsigned long long *)sb;
    const int oldCap = m->dm.capacity;
    m->dm.capacity = (int)newCap;
    launch_fill_pool(m->dm, oldCap, (int)newCap - oldCap, m->stream);
    CHS_CUDA(cudaGetLastError());
    return CHS_OK;
}

static int ensure_hash(chs_map *m, long long chunks)
{
    const size_t need = pow2_at_least((size_t)chunks * 2);
    if (need <= m->hashSize)
        return CHS_OK;
    if (m->dm.keys)
    {
        CHS_CUDA(cudaFreeAsync(m->dm.keys, m->stream));
        CHS_CUDA(cudaFreeAsync(m->dm.vals, m->stream));
    }
    void *k = nullptr, *v = nullptr;
    int rc = alloc_async(&k, sizeof(unsigned long long) * need, m->stream);
    if (rc)
        return rc;
    rc = alloc_async(&v, sizeof(int) * need, m->stream);
    if (rc)
        return rc;
    m->dm.keys = (unsigned long long *)k;
    m->dm.vals = (int *)v;
    m->dm.mask = (unsigned)(need - 1);
    m->hashSize = need;
    launch_fill_u64(m->dm.keys, need, kEmptyKey, m->stream);
    launch_fill_i32(m->dm.vals, need, -1, m->stream);               // empty entries hold -1: "claimed, value not published yet"
    launch_rebuild_hash(m->dm, 0, m->stream);
    CHS_CUDA(cudaGetLastError());
    return CHS_OK;
}

static int ensure_dirty(chs_map *m, long long ids)
{
    const size_t need = pow2_at_least((size_t)ids * 2);
    if (need <= m->dirtySize)
        return CHS_OK;
    void *k = nullptr, *l = nullptr;
    int rc = alloc_async(&k, sizeof(unsigned long long) * need, m->stream);
    if (rc)
        return rc;
    rc = alloc_async(&l, sizeof(unsigned long long) * (need / 2), m->stream);
    if (rc)
        return rc;
    if (m->dm.dirty_list)
    {
        CHS_CUDA(cudaMemcpyAsync(l, m->dm.dirty_list, sizeof(unsigned long long) * (size_t)m->dm.dirty_cap, cudaMemcpyDeviceToDevice, m->stream));
        CHS_CUDA(cudaFreeAsync(m->dm.dirty_list, m->stream));
        CHS_CUDA(cudaFreeAsync(m->dm.dirty_keys, m->stream));
    }
    m->dm.dirty_keys = (unsigned long long *)k;
    m->dm.dirty_list = (unsigned long long *)l;
    m->dm.dirty_mask = (unsigned)(need - 1);
    m->dm.dirty_cap = (int)(need / 2);
    m->dirtySize = need;
    launch_fill_u64(m->dm.dirty_keys, need, kEmptyKey, m->stream);
    launch_rebuild_dirty(m->dm, 0, m->stream);
    CHS_CUDA(cudaGetLastError());
    return CHS_OK;
}

static chs_frame_stats stats_of(const HostSnapshot &c)
{
    chs_frame_stats o{};
    o.candidates = c.candidates;
    o.new_candidates = c.new_count;
    o.brick_units = c.unit_count;
    o.n_upd = (int64_t)(((unsigned long long)(unsigned)c.n_upd_hi << 32) | (unsigned)c.n_upd_lo);
    o.n_carve = c.n_carve;
    o.n_col = c.n_col;
    o.n_new = c.n_new;
    o.updated_chunks = c.updated_chunks;
    o.total_chunks = c.n_chunks;
    o.dirty_chunks = c.n_dirty;
    o.error_flags = c.error_flags;
    return o;
}

// A fused batch has finished: per-frame counters of its K frames; the map totals are those after the last frame.
static chs_map::CallStats *call_stats(chs_map *m, int callId)
{
    chs_map::CallStats &c = m->callStats[callId & 3];
    return (callId > 0 && c.id == callId) ? &c : nullptr;
}

static void retire_batch(chs_map *m, const HostBatchSnapshot &b, int base, int callId)
{
    m->lastBricksSpanNs = b.bricks_span_ns;
    if (m->timeline)
    {
        std::array<long long, kTimelineStamps> tl;
        std::memcpy(tl.data(), b.timeline, sizeof(long long) * kTimelineStamps);
        m->timelineHist.push_back(tl);
        if (m->timelineHist.size() > 256)
            m->timelineHist.pop_front();
    }
    m->knownChunks = b.n_chunks;
    m->knownDirty = b.n_dirty;
    const int K = std::min(std::max(b.K, 0), kMaxBatch);
    if (base < 0)
        base = 0;
    chs_map::CallStats *cs = call_stats(m, callId);
    std::vector<chs_frame_stats> local((size_t)K);
    for (int f = 0; f < K; f++)
    {
        chs_frame_stats &o = local[f];
        o.candidates = b.candidates[f];
        o.new_candidates = b.new_count;          // per batch: the work lists are shared by the K frames
        o.brick_units = b.unit_count;
        o.n_upd = b.n_upd[f];
        o.n_carve = b.n_carve[f];
        o.n_col = b.n_col[f];
        o.n_new = b.n_new[f];
        o.updated_chunks = b.updated_chunks[f];
        o.total_chunks = b.n_chunks;
        o.dirty_chunks = b.n_dirty;
        o.error_flags = b.error_flags;
    }
    if (K > 0)
    {
        // chs_get_frame_stats after a batch reports its last frame
        const chs_frame_stats &o = local[K - 1];
        HostSnapshot &c = m->lastFrame;
        c.n_chunks = b.n_chunks; c.n_dirty = b.n_dirty; c.error_flags = b.error_flags;
        c.unit_count = b.unit_count; c.new_count = b.new_count; c.candidates = (int)o.candidates;
        c.n_new = (int)o.n_new; c.updated_chunks = (int)o.updated_chunks; c.n_carve = (int)o.n_carve;
        c.n_col = (int)o.n_col; c.n_upd_lo = (int)(o.n_upd & 0xffffffffll); c.n_upd_hi = (int)(o.n_upd >> 32);
    }
    if (cs)
    {
        for (int f = 0; f < K && base + f < (int)cs->st.size(); f++)
            cs->st[base + f] = local[f];
        cs->pending -= K;
    }
}

// Retire completed counter snapshots (the frame graph's last node writes them into the pinned ring); `block` waits
// for all of them.
static int poll_inflight(chs_map *m, bool block)
{
    if (block && !m->inflight.empty())
        CHS_CUDA(cudaStreamSynchronize(m->stream));
    while (!m->inflight.empty())
    {
        InFlight &f = m->inflight.front();
        if (f.batch)
        {
            const volatile HostBatchSnapshot *b = &m->hBatchSnap[f.slot];
            bool arrived = b->head == f.frameId && b->tail == f.frameId;
            HostBatchSnapshot snap;
            if (arrived)
            {
                // no fence on the device side: the checksum tells whether every payload word has landed
                std::atomic_thread_fence(std::memory_order_acquire);
                std::memcpy(&snap, (const void *)&m->hBatchSnap[f.slot], sizeof(snap));
                arrived = snap.head == f.frameId && snap.tail == f.frameId && snap.checksum == batch_snapshot_checksum(snap);
            }
            if (!arrived)
            {
                if (block)
                    return fail(CHS_ERR_CUDA, "batch counter snapshot missing after synchronisation");
                break;
            }
            retire_batch(m, snap, f.batchIndex, f.callId);
            m->inflight.pop_front();
            continue;
        }
        const volatile HostSnapshot *c = &m->hSnap[f.slot];
        if (c->id0 != f.frameId || c->id1 != f.frameId || c->id2 != f.frameId || c->id3 != f.frameId)
        {
            if (block)
                return fail(CHS_ERR_CUDA, "counter snapshot missing after synchronisation");
            break;
        }
        std::atomic_thread_fence(std::memory_order_acquire);
        m->lastFrame = m->hSnap[f.slot];
        m->knownChunks = m->lastFrame.n_chunks;
        m->knownDirty = m->lastFrame.n_dirty;
        if (chs_map::CallStats *cs = call_stats(m, f.callId))
            if (f.batchIndex >= 0 && f.batchIndex < (int)cs->st.size())
            {
                cs->st[f.batchIndex] = stats_of(m->lastFrame);
                cs->pending -= 1;
            }
        m->inflight.pop_front();
    }
    return CHS_OK;
}

// Keep at most `depth` launches in flight: wait for the OLDEST ones only (their counter snapshots arrive in pinned memory), never
// drain the stream -- the GPU keeps the younger launches to work on while the host waits. Without this bound a caller that
// enqueues faster than the GPU works piles up worst-case capacity reservations (one candidate box of chunks per launch in flight)
// until ensure_capacity has to synchronise with everything.
static int wait_inflight_below(chs_map *m, size_t depth)
{
    unsigned spins = 0;
    while (m->inflight.size() >= depth)
    {
        int rc = poll_inflight(m, false);
        if (rc)
            return rc;
        if (m->inflight.size() < depth)
            break;
        _mm_pause();
        if ((++spins & 0xFFFu) == 0)
        {
            const cudaError_t q = cudaStreamQuery(m->stream);
            if (q == cudaSuccess)
                return poll_inflight(m, true);                      // stream idle: everything has arrived (or is reported missing)
            if (q != cudaErrorNotReady)
                return fail(CHS_ERR_CUDA, std::string("stream error while waiting for a batch: ") + cudaGetErrorString(q));
        }
    }
    return CHS_OK;
}

// Reserve a ring slot for the frame about to be launched.
static int reserve_snapshot(chs_map *m, int frameId, long long newBound, long long dirtyBound, HostSnapshot **slotOut)
{
    if ((int)m->inflight.size() >= chs_map::kRing)
    {
        int rc = poll_inflight(m, false);
        if (rc)
            return rc;
        if ((int)m->inflight.size() >= chs_map::kRing && (rc = poll_inflight(m, true)))
            return rc;
    }
    InFlight f;
    f.frameId = frameId;
    f.slot = m->ringNext;
    m->ringNext = (m->ringNext + 1) % chs_map::kRing;
    f.newBound = newBound;
    f.dirtyBound = dirtyBound;
    m->hSnap[f.slot].id0 = m->hSnap[f.slot].id1 = m->hSnap[f.slot].id2 = m->hSnap[f.slot].id3 = -1;
    m->inflight.push_back(f);
    *slotOut = &m->hSnap[f.slot];
    return CHS_OK;
}

static bool finite12(const float *p)
{
    for (int i = 0; i < 12; i++)
        if (!std::isfinite(p[i]))
            return false;
    return true;
}

static void fill_camera(CameraDev *c, const float pose[12], const chs_camera &cam)
{
    for (int r = 0; r < 3; r++)
    {
        for (int k = 0; k < 3; k++)
            c->R[r * 3 + k] = pose[r * 4 + k];
        c->t[r] = pose[r * 4 + 3];
    }
    c->fx = cam.fx; c->fy = cam.fy; c->cx = cam.cx; c->cy = cam.cy;
    c->W = cam.width; c->H = cam.height;
    c->Wf = (float)cam.width; c->Hf = (float)cam.height;
}

struct FramePlan
{
    FrustumGeom fg;
    CandidateBox box;
    long long cand;         // IDs in the candidate box (ChunkManager.cpp:187-196)
    long long dirtyBound;   // IDs in the box grown by one chunk per side: what the frame can mark dirty
};

// Host part of Chisel.h:64-68 / :119-123: frustum and candidate ID box, exact.
static int plan_frame(chs_map *m, const float pose[12], const chs_camera *cam, FramePlan *pl)
{
    if (cam->width <= 0 || cam->height <= 0 || !finite12(pose))
        return fail(CHS_ERR_INVALID, "bad camera size or non-finite pose (quirk Q14: rejected at the boundary)");
    build_frustum(pose, *cam, &pl->fg);
    if (!candidate_box(pl->fg, m->cfg.chunk_size, m->cfg.resolution, &pl->box))
        return fail(CHS_ERR_INVALID, "frustum is not finite or lies outside the packable chunk-ID range");
    for (int k = 0; k < 3; k++)
        if (pl->box.lo[k] - 1 < -kIdBias || pl->box.hi[k] + 1 >= kIdBias)
            return fail(CHS_ERR_INVALID, "chunk IDs outside [-2^20, 2^20)");
    pl->cand = pl->box.count();
    if (pl->cand > (1ll << 26))
        return fail(CHS_ERR_CAPACITY, "candidate box larger than 2^26 chunks");
    pl->dirtyBound = 1;
    for (int k = 0; k < 3; k++)
        pl->dirtyBound *= (long long)(pl->box.hi[k] - pl->box.lo[k] + 3);
    return CHS_OK;
}

// Capacity: known counts + bounds of launches whose counters have not come back yet + this launch (`cand` new chunks,
// `dirtyBound` dirty IDs at most); work lists for `cand` chunks.
// poolLater (fused path): if the pool cannot hold the worst case (every candidate becomes a chunk -- 48 KB each, far too
// pessimistic for large candidate boxes), do not grow it here: *poolLater is set and the caller sizes the pool from the exact
// number of chunks the candidates kernel found possible, before it launches the kernel that creates them.
static int ensure_capacity(chs_map *m, long long cand, long long dirtyBound, bool *poolLater = nullptr)
{
    cudaStream_t st = m->stream;
    int rc = poll_inflight(m, false);
    if (rc)
        return rc;
    long long chunkUb = m->knownChunks + cand, dirtyUb = m->knownDirty + dirtyBound;
    for (const InFlight &f : m->inflight)
    {
        chunkUb += f.newBound;
        dirtyUb += f.dirtyBound;
    }
    if (poolLater)
        *poolLater = false;
    if (chunkUb > m->dm.capacity || (size_t)chunkUb * 2 > m->hashSize || (size_t)dirtyUb * 2 > m->dirtySize)
    {
        // tighten the bound before growing: wait for outstanding counters
        if ((rc = poll_inflight(m, true)))
            return rc;
        // head-room for several launches in flight: the bound per launch is the candidate count
        const long long wantChunks = m->knownChunks + 4 * cand + m->knownChunks / 4;
        const long long wantDirty = m->knownDirty + 4 * dirtyBound + m->knownDirty / 4;
        const bool defer = poolLater && m->knownChunks + cand > m->dm.capacity && cand > 32768;
        if (defer)
            *poolLater = true;
        else if ((rc = ensure_pool(m, wantChunks)))
            return rc;
        if ((rc = ensure_hash(m, defer ? m->knownChunks + cand : wantChunks)) || (rc = ensure_dirty(m, wantDirty)))
            return rc;
    }
    const int bpa = m->cfg.chunk_size / 8;
    // the fused path keeps its units in two buffers of cand * bricks entries (four cost buckets)
    if ((rc = grow_buffer(&m->dUnits, &m->unitsCap, 2 * (size_t)cand * bpa * bpa * bpa, st)) || (rc = grow_buffer(&m->dNews, &m->newsCap, (size_t)cand, st)))
        return rc;
    return CHS_OK;
}

static size_t hiz_tiles(const chs_camera *cam)
{
    size_t total = 0;
    for (int l = 0; l < kHizLevels; l++)
    {
        const int tile = 8 << l;
        total += (size_t)((cam->width + tile - 1) / tile) * ((cam->height + tile - 1) / tile);
    }
    return total;
}

// The Hi-Z levels >= 4 of a frame fit the candidates kernel's shared memory (then the TMA Hi-Z kernel, which stops at level 3, can be used).
static bool hiz_coarse_fits(const chs_camera *cam)
{
    int levels = kHizLevels, coarse = 0;
    for (int l = 0; l < kHizLevels; l++)
    {
        const int tile = 8 << l, w = (cam->width + tile - 1) / tile, h = (cam->height + tile - 1) / tile;
        if (l >= 3 && w <= 3 && h <= 3 && levels == kHizLevels)
            levels = l + 1;
    }
    for (int l = 4; l < levels; l++)
    {
        const int tile = 8 << l;
        coarse += ((cam->width + tile - 1) / tile) * ((cam->height + tile - 1) / tile);
    }
    return coarse <= 48;
}

// Everything of FrameParams except the image pointers, the work lists and the ids.
static void fill_frame_params(chs_map *m, const chs_integrator *integ, const float pose[12], const chs_camera *cam, const float cpose[12],
                              const chs_camera *ccam, bool colorPath, int channels, const FramePlan &pl, float2 *hizBase, FrameParams *out)
{
    FrameParams &fp = *out;
    fill_camera(&fp.cam, pose, *cam);
    if (colorPath)
        fill_camera(&fp.ccam, cpose, *ccam);
    else
        fp.ccam = fp.cam;
    fp.channels = channels;
    fp.color_path = colorPath ? 1 : 0;
    fp.trunc_kind = integ->trunc_kind;
    fp.trunc_param = integ->trunc_param;
    // ProjectionIntegrator.h:59 / :108 -- double expression, narrowed once
    fp.diag = (float)(2.0 * std::sqrt((double)3.0f) * (double)m->cfg.resolution);
    fp.carve_dist = integ->carving_dist;
    fp.carve = integ->carving_enabled ? 1 : 0;
    fp.weight = integ->weight;
    fp.depth_cutoff = colorPath ? 100.0f : 50.0f;
    fp.same_cam = (colorPath && std::memcmp(&fp.cam, &fp.ccam, sizeof(CameraDev)) == 0) ? 1 : 0;
    fp.wu_const = integ->weight / (5 * integ->trunc_param);            // ConstantWeighter.h:43-46 (binary32, used for the constant truncator)
    {
        float T = (float)1e-5;
        if ((double)T < 1e-5)
            T = std::nextafterf(T, INFINITY);
        fp.sdf_carve_max = T;
    }
    for (int k = 0; k < 3; k++)
    {
        fp.lo[k] = pl.box.lo[k];
        fp.n[k] = pl.box.hi[k] - pl.box.lo[k] + 1;
    }
    for (int p = 0; p < 6; p++)
    {
        for (int k = 0; k < 3; k++)
            fp.planes[p][k] = pl.fg.plane[p].n[k];
        fp.planes[p][3] = pl.fg.plane[p].d;
    }
    {
        // view pyramid for culling (classify_box): u >= -3  <=>  fx x + (cx + 3) z >= 0 for z > 0, etc., in camera coordinates;
        // n_world = R n_cam, d = -n_world . t
        const float m = 3.0f;
        const float nc[5][3] = {{cam->fx, 0.0f, cam->cx + m}, {-cam->fx, 0.0f, (float)cam->width + m - cam->cx},
                                {0.0f, cam->fy, cam->cy + m}, {0.0f, -cam->fy, (float)cam->height + m - cam->cy}, {0.0f, 0.0f, 1.0f}};
        for (int p = 0; p < 5; p++)
        {
            float d = 0.0f;
            for (int r = 0; r < 3; r++)
            {
                const float nw = pose[r * 4 + 0] * nc[p][0] + pose[r * 4 + 1] * nc[p][1] + pose[r * 4 + 2] * nc[p][2];
                fp.view_planes[p][r] = nw;
                d -= nw * pose[r * 4 + 3];
            }
            fp.view_planes[p][3] = d;
        }
    }
    size_t off = 0;
    fp.hiz_levels = kHizLevels;
    for (int l = 0; l < kHizLevels; l++)
    {
        const int tile = 8 << l;
        fp.hizW[l] = (cam->width + tile - 1) / tile;
        fp.hizH[l] = (cam->height + tile - 1) / tile;
        fp.hiz[l] = hizBase + off;
        off += (size_t)fp.hizW[l] * fp.hizH[l];
        if (l >= 3 && fp.hizW[l] <= 3 && fp.hizH[l] <= 3 && fp.hiz_levels == kHizLevels)
            fp.hiz_levels = l + 1;
    }
    fp.hiz_blocks = fp.hizW[3] * fp.hizH[3];
}

static int integrate_common(chs_map *m, const chs_integrator *integ, const float *depth, int mem, const float pose[12],
                            const chs_camera *cam, const uint8_t *color, int channels, const float cpose[12],
                            const chs_camera *ccam, bool colorPath, int batchIndex = -1)
{
    if (!m || !integ || !depth || !pose || !cam)
        return fail(CHS_ERR_INVALID, "null argument");
    if (colorPath && (!color || !cpose || !ccam || channels < 1 || channels > 4 || !finite12(cpose) || ccam->width <= 0 || ccam->height <= 0))
        return fail(CHS_ERR_INVALID, "bad colour arguments");
    if (integ->trunc_kind == CHS_TRUNC_PER_PIXEL && !integ->trunc_per_pixel)
        return fail(CHS_ERR_INVALID, "CHS_TRUNC_PER_PIXEL without trunc_per_pixel");
    CHS_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = m->stream;
    FramePlan pl;
    int rc = plan_frame(m, pose, cam, &pl);
    if (rc)
        return rc;
    const long long cand = pl.cand, dirtyBound = pl.dirtyBound;
    if ((rc = ensure_capacity(m, cand, dirtyBound)))
        return rc;

    const size_t npx = (size_t)cam->width * cam->height;
    if ((rc = grow_buffer(&m->dHiz, &m->hizCap, hiz_tiles(cam), st)))
        return rc;
    FrameParams fp{};
    fill_frame_params(m, integ, pose, cam, cpose, ccam, colorPath, channels, pl, m->dHiz, &fp);
    fp.hiz_ticket = m->dHizTickets + 2 * kMaxBatch;
    // inputs
    bool copied = false;
    if (mem == CHS_MEM_HOST)
    {
        if ((rc = grow_buffer(&m->dDepth, &m->depthCap, npx, st)))
            return rc;
        CHS_CUDA(cudaMemcpyAsync(m->dDepth, depth, npx * sizeof(float), cudaMemcpyHostToDevice, st));
        fp.depth = m->dDepth;
        copied = true;
    }
    else
        fp.depth = depth;
    fp.trunc_img = nullptr;
    if (integ->trunc_kind == CHS_TRUNC_PER_PIXEL)
    {
        if (mem == CHS_MEM_HOST)
        {
            if ((rc = grow_buffer(&m->dTrunc, &m->truncCap, npx, st)))
                return rc;
            CHS_CUDA(cudaMemcpyAsync(m->dTrunc, integ->trunc_per_pixel, npx * sizeof(float), cudaMemcpyHostToDevice, st));
            fp.trunc_img = m->dTrunc;
        }
        else
            fp.trunc_img = integ->trunc_per_pixel;
    }
    else if (integ->trunc_kind != CHS_TRUNC_CONSTANT)
    {
        if ((rc = grow_buffer(&m->dTrunc, &m->truncCap, npx, st)))
            return rc;
        fp.trunc_img = m->dTrunc;                                   // written by frame_prepare
    }
    if (colorPath)
    {
        const size_t nb = (size_t)ccam->width * ccam->height * channels;
        if (mem == CHS_MEM_HOST)
        {
            if ((rc = grow_buffer(&m->dColor, &m->colorCap, nb, st)))
                return rc;
            CHS_CUDA(cudaMemcpyAsync(m->dColor, color, nb, cudaMemcpyHostToDevice, st));
            fp.color = m->dColor;
        }
        else
            fp.color = color;
        if ((rc = grow_buffer(&m->dColorPacked, &m->colorPackedCap, (size_t)ccam->width * ccam->height, st)))
            return rc;
        fp.color_packed = m->dColorPacked;
    }
    if (copied)
        CHS_CUDA(cudaEventRecord(m->h2dDone, st));
    fp.units = m->dUnits;
    fp.units_cap = (int)std::min<size_t>(m->unitsCap, 0x1fffffff);
    fp.news = m->dNews;
    fp.news_cap = (int)std::min<size_t>(m->newsCap, 0x7fffffff);
    fp.frame_id = ++m->frameId;
    {
        // smallest odd stride >= 7919 that is coprime to the box size
        long long stride = 7919;
        auto gcd = [](long long a, long long b) { while (b) { long long t = a % b; a = b; b = t; } return a; };
        while (cand > 1 && gcd(stride, cand) != 1)
            stride += 2;
        fp.cand_stride = (int)(cand > 1 ? stride % cand : 1);
        if (fp.cand_stride == 0)
            fp.cand_stride = 1;
    }

    HostSnapshot *slotPtr = nullptr;
    if ((rc = reserve_snapshot(m, fp.frame_id, cand, dirtyBound, &slotPtr)))
        return rc;
    m->inflight.back().batchIndex = batchIndex;
    m->inflight.back().callId = batchIndex >= 0 ? m->callId : 0;
    // size the new-chunk kernel's grid from what recent frames needed (the kernel strides, so any size is correct)
    const long long newHint = m->haveFrame ? std::max<long long>(64, 2ll * m->lastFrame.new_count) : cand;
    CHS_CUDA(frame_graph_launch(m->frameGraph, fp, m->dm, cand, newHint, slotPtr, m->profiling, m->evt, st));
    if (m->profiling)
        m->frameTimed = true;
    m->haveFrame = true;
    // the caller may reuse its host buffers as soon as we return (chisel_ros does: CR ChiselServer.cpp:285-295)
    if (copied)
        CHS_CUDA(cudaEventSynchronize(m->h2dDone));
    return CHS_OK;
}

// K host images of `bytes` bytes each -> K consecutive device images. Callers usually keep their frames in one ring or array,
// i.e. at a constant stride: then the K copies are ONE strided copy (one DMA descriptor chain instead of K submissions).
static int copy_images_h2d(void *dst, const void *const *src, int K, size_t bytes, cudaStream_t st, cudaMemcpyKind kind = cudaMemcpyHostToDevice)
{
    bool strided = K > 1;
    const ptrdiff_t stride = K > 1 ? (const char *)src[1] - (const char *)src[0] : 0;
    for (int f = 1; f < K && strided; f++)
        strided = (const char *)src[f] - (const char *)src[f - 1] == stride;
    if (strided && stride >= (ptrdiff_t)bytes)
    {
        CHS_CUDA(cudaMemcpy2DAsync(dst, bytes, src[0], (size_t)stride, bytes, (size_t)K, kind, st));
        return CHS_OK;
    }
    for (int f = 0; f < K; f++)
        CHS_CUDA(cudaMemcpyAsync((char *)dst + bytes * f, src[f], bytes, kind, st));
    return CHS_OK;
}

// capi_comm.inc
static int exchange_frames(chs_map *m, void *depth, size_t depthBytesPerFrame, void *color, size_t colorBytesPerFrame, cudaStream_t st);
struct PushSeg
{
    const void *src;
    unsigned long long dst_off, bytes;         // destination offset inside an arena
};
static int ensure_peer_arena(chs_map *m, size_t depthBytes, size_t mmBytes, size_t colorBytes, size_t hizBytes);
static int peer_push(chs_map *m, const PushSeg *segs, int nSeg, cudaStream_t cs, bool raiseFlags);
static int peer_wait(chs_map *m, cudaStream_t st, unsigned long long *timeline, bool hizStamps);
static void release_peer_arena(chs_map *m);
static int world_any_mm(chs_map *m, bool *anyMm);

// ---------------------------------------------------------------------------------------------------------
// chs_integrate_batch: n consecutive frames of one sensor stream (same image size, intrinsics and integrator). Sub-batches of
// up to kMaxBatch frames go through the fused kernels (integrate_batch_impl.cuh); the map afterwards is bit-identical to n calls of
// chs_integrate_depth[_color] in order. Frames whose colour camera differs from the depth camera are integrated one by one.
// CHS_HOST_PROFILE=1: where the host time of a fused batch call goes (printed per process at exit)
static double g_launchLapNs[12] = {};
static std::chrono::steady_clock::time_point g_launchLapT;
struct HostProfile
{
    static constexpr int kSections = 9;
    const char *names[kSections] = {"plan", "capacity+slot", "arena+buffers", "frame params", "copies+push", "frame table upload", "brick frames", "launch", "waiting for a batch in flight"};
    double ns[kSections] = {};
    long long calls = 0;
    bool on = std::getenv("CHS_HOST_PROFILE") != nullptr;
    ~HostProfile()
    {
        if (!on || !calls)
            return;
        double tot = 0;
        for (double v : ns)
            tot += v;
        std::fprintf(stderr, "[chs host profile] %lld fused batches, %.1f us per call:", calls, tot / calls / 1e3);
        for (int i = 0; i < kSections; i++)
            std::fprintf(stderr, " %s %.1f;", names[i], ns[i] / calls / 1e3);
        std::fprintf(stderr, "\n[chs host profile] inside launch: in capi %.1f; dispatch %.1f; resident %.1f; grids %.1f; before %.1f; released event %.1f; pack+fork %.1f; hiz %.1f; prepared event+wait %.1f; candidates %.1f; pack wait %.1f; bricks %.1f\n",
                     g_launchLapNs[8] / calls / 1e3, g_launchLapNs[9] / calls / 1e3, g_launchLapNs[10] / calls / 1e3, g_launchLapNs[11] / calls / 1e3, g_launchLapNs[7] / calls / 1e3, g_launchLapNs[6] / calls / 1e3, g_launchLapNs[0] / calls / 1e3, g_launchLapNs[1] / calls / 1e3, g_launchLapNs[2] / calls / 1e3, g_launchLapNs[3] / calls / 1e3,
                     g_launchLapNs[4] / calls / 1e3, g_launchLapNs[5] / calls / 1e3);
    }
};
static HostProfile g_hostProfile;
void host_launch_lap(int i)
{
    if (!g_hostProfile.on)
        return;
    const auto t1 = std::chrono::steady_clock::now();
    if (i >= 0)
        g_launchLapNs[i] += (double)std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - g_launchLapT).count();
    g_launchLapT = t1;
}
struct HostSection
{
    std::chrono::steady_clock::time_point t0;
    HostSection() : t0(std::chrono::steady_clock::now()) {}
    void lap(int i)
    {
        if (!g_hostProfile.on)
            return;
        const auto t1 = std::chrono::steady_clock::now();
        g_hostProfile.ns[i] += (double)std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count();
        t0 = t1;
    }
};

static int integrate_batch_fused(chs_map *m, const chs_integrator *integ, int K, const chs_frame *frames, int mem, const chs_camera *cam,
                                 int channels, const chs_camera *ccam, bool colorPath, int statsBase)
{
    cudaStream_t st = m->stream;
    int rc;
    HostSection hs;
    g_hostProfile.calls++;
    FramePlan pl[kMaxBatch];
    int ulo[3], uhi[3];
    for (int f = 0; f < K; f++)
    {
        if ((rc = plan_frame(m, frames[f].pose, cam, &pl[f])))
            return rc;
        for (int k = 0; k < 3; k++)
        {
            ulo[k] = f == 0 ? pl[f].box.lo[k] : std::min(ulo[k], pl[f].box.lo[k]);
            uhi[k] = f == 0 ? pl[f].box.hi[k] : std::max(uhi[k], pl[f].box.hi[k]);
        }
    }
    long long unionCand = 1, dirtyBound = 1;
    for (int k = 0; k < 3; k++)
    {
        unionCand *= (long long)(uhi[k] - ulo[k] + 1);
        dirtyBound *= (long long)(uhi[k] - ulo[k] + 3);
    }
    if (unionCand > (1ll << 26))
        return CHS_ERR_NOT_FOUND;                                   // frames too far apart to share a box: the caller falls back to single frames
    hs.lap(0);
    if ((rc = wait_inflight_below(m, 4)))
        return rc;
    hs.lap(8);
    bool poolLater = false;
    if ((rc = ensure_capacity(m, unionCand, dirtyBound, &poolLater)))
        return rc;

    hs.lap(1);
    const size_t npx = (size_t)cam->width * cam->height;
    const size_t cpx = colorPath ? (size_t)ccam->width * ccam->height : 0;
    const size_t tiles = hiz_tiles(cam);
    const bool computeTrunc = integ->trunc_kind == CHS_TRUNC_QUADRATIC || integ->trunc_kind == CHS_TRUNC_INVERSE;
    const bool perPixel = integ->trunc_kind != CHS_TRUNC_CONSTANT;
    // distributed batch (chs_integrate_batch_distributed): only the frames [distFirst, distFirst + distCount) carry images; they
    // are staged like host frames (device frames: copied device to device) and the other ranks' images arrive by all-gather
    const bool dist = m->distCount > 0;
    const int locFirst = dist ? m->distFirst : 0, locEnd = dist ? m->distFirst + m->distCount : K;
    const bool devSrc = mem == CHS_MEM_DEVICE || mem == CHS_MEM_DEVICE_ASYNC;
    const bool hostMem = mem == CHS_MEM_HOST || mem == CHS_MEM_HOST_ASYNC || dist;
    // device frames that are already complete: Hi-Z + colour packing run on the copy stream, beside the kernels of the previous batch
    const bool devAsync = mem == CHS_MEM_DEVICE_ASYNC && !dist;
    bool anyMm = false;
    for (int f = locFirst; f < locEnd; f++)
        anyMm |= frames[f].depth_mm != nullptr;
    if (dist && world_any_mm(m, &anyMm))
        return CHS_ERR_CUDA;
    // Distributed batch on more than one rank: the images of the other ranks arrive by peer-memory pushes into this rank's exchange
    // arena when CUDA IPC is available (collective set-up, first distributed batch), else by NCCL all-gather into the staging set.
    bool push = false;
    if (dist && m->cfg.world > 1)
    {
        if ((rc = ensure_peer_arena(m, anyMm ? 0 : npx * sizeof(float) * kMaxBatch, anyMm ? npx * sizeof(uint16_t) * kMaxBatch : 0,
                                    colorPath ? cpx * channels * kMaxBatch : 0, tiles * kMaxBatch * sizeof(float2))))
            return rc;
        push = m->arena.ok;
    }
    // Device frames over the peer-memory exchange, opt-in (CHS_SHARD_HIZ_PUSH=1): the Hi-Z pyramids are built rank by rank, each
    // rank those of the frames it ingests, and stored into every arena behind the frames (conditions of the TMA Hi-Z kernel: float
    // depth, constant truncator). Measured on 2 B200s: the push stream then carries push + Hi-Z of every
    // step back to back and becomes the longest chain of the step (127 us vs 79 us) -- off by default until the push is faster.
    static const bool shardHizOn = std::getenv("CHS_SHARD_HIZ_PUSH") != nullptr;
    const bool shardHiz = shardHizOn && push && devSrc && !perPixel && !anyMm && (cam->width % 4) == 0 && hiz_coarse_fits(cam) &&
                          std::getenv("CHS_NO_TMA_HIZ") == nullptr;
    char *const pushSet = push ? m->arena.base + 4096 + (size_t)((m->arena.step + 1) % chs_map::kArenaSets) * m->arena.setBytes : nullptr;
    const int setIdx = (m->batchId + 1) & 1;
    chs_map::BatchSet &bs = m->bset[setIdx];
    // Host frames: copies and prepare run on the copy stream, beside the kernels of the previous batch. Device frames are ordered by
    // the map's stream anyway (the caller produced them there), so everything stays on it: no cross-stream hand-overs.
    cudaStream_t cs = (hostMem || devAsync) ? m->copyStream : st;
    const bool direct = push && devSrc;
    if (dist && mem == CHS_MEM_DEVICE)
    {
        // the caller produced its device frames on the map's stream: the copy / push stream picks them up from there
        CHS_CUDA(cudaEventRecord(bs.fork, st));
        CHS_CUDA(cudaStreamWaitEvent(direct ? m->pushStream : cs, bs.fork, 0));
    }
    // Peer-memory exchange: this rank's share of the step goes into the step's set of every arena. Device frames are pushed from
    // where they are, on a stream of their own: the push is neither held up by this staging set (still read by the batch before
    // last) nor by the Hi-Z kernels on the copy stream, and runs up to a whole step ahead (the arenas hold three sets). The
    // consumers -- this rank's too -- wait for the flag words, not for an event.
    auto push_frames = [&]() -> int
    {
        PushSeg segs[2 * kMaxBatch];
        int nSeg = 0;
        const size_t dBytes = npx * (anyMm ? sizeof(uint16_t) : sizeof(float)), cBytes = cpx * channels;
        const size_t dOff = (size_t)(pushSet - m->arena.base) + (anyMm ? m->arena.mmOff : m->arena.depthOff);
        const size_t cOff = (size_t)(pushSet - m->arena.base) + m->arena.colorOff;
        if (direct)
        {
            for (int f = locFirst; f < locEnd; f++)
            {
                if ((frames[f].depth_mm != nullptr) != anyMm)
                    return fail(CHS_ERR_INVALID, "distributed batches need one depth encoding for all frames");
                segs[nSeg++] = PushSeg{anyMm ? (const void *)frames[f].depth_mm : (const void *)frames[f].depth, dOff + dBytes * f, dBytes};
                if (colorPath)
                    segs[nSeg++] = PushSeg{frames[f].color, cOff + cBytes * f, cBytes};
            }
        }
        else
        {
            // staged host frames: this rank's share is one contiguous block per image kind
            const size_t n = (size_t)(locEnd - locFirst);
            segs[nSeg++] = PushSeg{anyMm ? (const void *)(bs.depthMm + npx * locFirst) : (const void *)(bs.depth + npx * locFirst), dOff + dBytes * locFirst, dBytes * n};
            if (colorPath)
                segs[nSeg++] = PushSeg{bs.color + cBytes * locFirst, cOff + cBytes * locFirst, cBytes * n};
        }
        return peer_push(m, segs, nSeg, direct ? m->pushStream : cs, !shardHiz);
    };
    if (direct && (rc = push_frames()))
        return rc;
    // the set is free once the kernels of the batch that used it last (two batches ago) are done
    if (bs.used && (hostMem || devAsync))
        CHS_CUDA(cudaStreamWaitEvent(cs, bs.released, 0));
    {
        const bool grow = tiles * kMaxBatch > bs.hizCap || ((hostMem || anyMm) && npx * kMaxBatch > bs.depthCap) || (anyMm && hostMem && npx * kMaxBatch > bs.depthMmCap) ||
                          ((computeTrunc || (perPixel && hostMem)) && npx * kMaxBatch > bs.truncCap) ||
                          (colorPath && ((hostMem && cpx * channels * kMaxBatch > bs.colorCap) || cpx * kMaxBatch > bs.packedCap));
        if (grow)
        {
            // rare (first batch, or a larger image): nothing may still be reading the old buffers
            CHS_CUDA(cudaStreamSynchronize(st));
            CHS_CUDA(cudaStreamSynchronize(cs));
        }
    }
    if ((rc = grow_buffer(&bs.hiz, &bs.hizCap, tiles * kMaxBatch, cs)))
        return rc;
    if ((hostMem || anyMm) && (rc = grow_buffer(&bs.depth, &bs.depthCap, npx * kMaxBatch, cs)))
        return rc;
    if (anyMm && hostMem && (rc = grow_buffer(&bs.depthMm, &bs.depthMmCap, npx * kMaxBatch, cs)))
        return rc;
    if ((computeTrunc || (perPixel && hostMem)) && (rc = grow_buffer(&bs.trunc, &bs.truncCap, npx * kMaxBatch, cs)))
        return rc;
    if (colorPath)
    {
        if (hostMem && (rc = grow_buffer(&bs.color, &bs.colorCap, cpx * channels * kMaxBatch, cs)))
            return rc;
        if ((rc = grow_buffer(&bs.packed, &bs.packedCap, cpx * kMaxBatch, cs)))
            return rc;
    }
    hs.lap(2);
    FrameParams fps[kMaxBatch];
    std::memset(fps, 0, sizeof(fps));
    for (int f = 0; f < K; f++)
    {
        FrameParams &fp = fps[f];
        fill_frame_params(m, integ, frames[f].pose, cam, frames[f].color_pose, ccam, colorPath, channels, pl[f],
                          shardHiz ? reinterpret_cast<float2 *>(pushSet + m->arena.hizOff) + tiles * f : bs.hiz + tiles * f, &fp);
        fp.hiz_ticket = m->dHizTickets + setIdx * kMaxBatch + f;
        if (dist ? anyMm : frames[f].depth_mm != nullptr)
        {
            // millimetres: half the bytes over PCIe; batch_prepare converts into the float image
            if (hostMem)
                fp.depth_u16 = bs.depthMm + npx * f;
            else
                fp.depth_u16 = frames[f].depth_mm;
            fp.depth = bs.depth + npx * f;
        }
        else if (hostMem)
            fp.depth = bs.depth + npx * f;
        else
            fp.depth = frames[f].depth;
        fp.trunc_img = nullptr;
        if (integ->trunc_kind == CHS_TRUNC_PER_PIXEL)
        {
            fp.trunc_img = hostMem ? bs.trunc + npx * f : frames[f].trunc_per_pixel;
        }
        else if (computeTrunc)
            fp.trunc_img = bs.trunc + npx * f;                      // written by batch_prepare
        if (colorPath)
        {
            fp.color = hostMem ? bs.color + cpx * channels * f : frames[f].color;
            fp.color_packed = bs.packed + cpx * f;
        }
        if (push)
        {
            // the step's images are read from the exchange arena (set = parity of the step), where every rank has pushed its share
            if (anyMm)
                fp.depth_u16 = reinterpret_cast<const uint16_t *>(pushSet + m->arena.mmOff) + npx * f;
            else
                fp.depth = reinterpret_cast<const float *>(pushSet + m->arena.depthOff) + npx * f;
            if (colorPath)
                fp.color = reinterpret_cast<const uint8_t *>(pushSet + m->arena.colorOff) + cpx * channels * f;
        }
        fp.frame_id = ++m->frameId;
    }
    hs.lap(3);
    if (hostMem)
    {
        // runs of frames of the same kind (float / millimetre depth) go in one strided copy each
        const void *src[kMaxBatch];
        const cudaMemcpyKind kind = (dist && devSrc) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        for (int f0 = locFirst; f0 < locEnd && !direct;)
        {
            const bool mm = frames[f0].depth_mm != nullptr;
            if (dist && mm != anyMm)
                return fail(CHS_ERR_INVALID, "distributed batches need one depth encoding for all frames");
            int f1 = f0;
            for (; f1 < locEnd && (frames[f1].depth_mm != nullptr) == mm; f1++)
                src[f1 - f0] = mm ? (const void *)frames[f1].depth_mm : (const void *)frames[f1].depth;
            if ((rc = mm ? copy_images_h2d(bs.depthMm + npx * f0, src, f1 - f0, npx * sizeof(uint16_t), cs, kind)
                         : copy_images_h2d(bs.depth + npx * f0, src, f1 - f0, npx * sizeof(float), cs, kind)))
                return rc;
            f0 = f1;
        }
        if (integ->trunc_kind == CHS_TRUNC_PER_PIXEL)
        {
            for (int f = 0; f < K; f++)
                src[f] = frames[f].trunc_per_pixel;
            if ((rc = copy_images_h2d(bs.trunc, src, K, npx * sizeof(float), cs, kind)))
                return rc;
        }
        if (colorPath && !direct)
        {
            for (int f = locFirst; f < locEnd; f++)
                src[f - locFirst] = frames[f].color;
            if ((rc = copy_images_h2d(bs.color + cpx * channels * locFirst, src, locEnd - locFirst, cpx * channels, cs, kind)))
                return rc;
        }
        CHS_CUDA(cudaEventRecord(bs.copied, cs));
        if (push)
        {
            if (!direct && (rc = push_frames()))
                return rc;
        }
        else if (dist && (rc = exchange_frames(m, anyMm ? (void *)bs.depthMm : (void *)bs.depth, npx * (anyMm ? sizeof(uint16_t) : sizeof(float)),
                                               colorPath ? bs.color : nullptr, cpx * channels, cs)))
            return rc;
    }
    hs.lap(4);
    // the frame table, through a pinned ring slot (a pageable source costs the driver a staging copy per call)
    {
        FrameParams *slot = m->hFrameTables + (size_t)m->frameTableNext * kMaxBatch;
        m->frameTableNext = (m->frameTableNext + 1) % chs_map::kRing;
        std::memcpy(slot, fps, sizeof(FrameParams) * K);
        if (shardHiz)
        {
            // this rank's Hi-Z kernel follows its push at once, on the push stream -- a step ahead of the kernels that consume the
            // pyramids -- and raises the rank's arrived word everywhere. What it needs of the frames travels in its parameters.
            HizPeers hp{};
            for (int d = 0; d < m->cfg.world; d++)
            {
                hp.delta[d] = (long long)(m->arena.peer[d] - m->arena.base);
                hp.flag[d] = reinterpret_cast<unsigned *>(m->arena.peer[d]) + m->cfg.rank;
            }
            hp.hdr = reinterpret_cast<unsigned *>(m->arena.base);
            hp.world = m->cfg.world;
            hp.rank = m->cfg.rank;
            hp.step = m->arena.step;
            hp.stamp_slot = (int)(m->arena.step % chs_map::kArenaSets);
            hp.depth0 = reinterpret_cast<const float *>(pushSet + m->arena.depthOff);
            hp.hiz0 = reinterpret_cast<float2 *>(pushSet + m->arena.hizOff);
            hp.npx = (long long)npx;
            hp.tiles_per_frame = (long long)tiles;
            for (int l = 0; l < 4; l++)
            {
                hp.level_off[l] = (int)(fps[0].hiz[l] - fps[0].hiz[0]);
                hp.hizW[l] = fps[0].hizW[l];
                hp.hizH[l] = fps[0].hizH[l];
            }
            hp.W = cam->width;
            hp.H = cam->height;
            hp.carve = fps[0].carve;
            hp.cutoff = fps[0].depth_cutoff;
            hp.trunc = fps[0].trunc_param;
            hp.diag = fps[0].diag;
            hp.carve_dist = fps[0].carve_dist;
            CHS_CUDA(half::launch_hiz_sharded(locFirst, locEnd - locFirst, hp, m->pushStream));
        }
        CHS_CUDA(cudaMemcpyAsync(bs.dFrames, slot, sizeof(FrameParams) * K, cudaMemcpyHostToDevice, cs));
    }
    hs.lap(5);

    // reserve a slot of the batch snapshot ring
    if ((int)m->inflight.size() >= chs_map::kRing && (rc = poll_inflight(m, true)))
        return rc;
    InFlight inf;
    inf.frameId = ++m->batchId;
    inf.slot = m->batchRingNext;
    m->batchRingNext = (m->batchRingNext + 1) % chs_map::kRing;
    inf.newBound = unionCand;
    inf.dirtyBound = dirtyBound;
    inf.batch = true;
    inf.batchIndex = statsBase;
    inf.callId = m->callId;
    m->hBatchSnap[inf.slot].head = m->hBatchSnap[inf.slot].tail = -1;
    m->inflight.push_back(inf);

    BatchParams bp{};
    bp.frames = bs.dFrames;
    bp.K = K;
    for (int k = 0; k < 3; k++)
    {
        bp.lo[k] = ulo[k];
        bp.n[k] = uhi[k] - ulo[k] + 1;
    }
    bp.units = m->dUnits;
    bp.units_cap = (int)std::min<size_t>(m->unitsCap / 2, 0x1fffffff);
    {
        // smallest odd stride >= 7919 that is coprime to the box size
        long long stride = 7919;
        auto gcd = [](long long a, long long b) { while (b) { long long t = a % b; a = b; b = t; } return a; };
        while (unionCand > 1 && gcd(stride, unionCand) != 1)
            stride += 2;
        bp.cand_stride = (int)(unionCand > 1 ? stride % unionCand : 1);
        if (bp.cand_stride == 0)
            bp.cand_stride = 1;
    }
    bp.batch_id = inf.frameId;
    bp.bctr = bs.dBctr;
    bp.timeline = m->timeline ? bs.dTimeline : nullptr;
    bp.host_slot = &m->hBatchSnap[inf.slot];
    bp.slot_batch = m->dSlotBatch;
    BatchLaunchInfo info{};
    info.W = cam->width;
    info.H = cam->height;
    info.cW = colorPath ? ccam->width : 0;
    info.cH = colorPath ? ccam->height : 0;
    info.unionCandidates = unionCand;
    info.colorPath = colorPath;
    info.perPixel = perPixel;
    info.profiling = m->profiling;
    {
        // coarse Hi-Z levels in the candidates kernel's shared memory when they fit; then the Hi-Z kernel can be the TMA one
        int coarse = 0;
        for (int l = 4; l < fps[0].hiz_levels; l++)
            coarse += fps[0].hizW[l] * fps[0].hizH[l];
        bp.coarse_in_shared = coarse <= 48 ? 1 : 0;
        bool tma = bp.coarse_in_shared && !perPixel && !anyMm && (cam->width % 4) == 0 && std::getenv("CHS_NO_TMA_HIZ") == nullptr;
        for (int f = 0; f < K && tma; f++)
            tma = (reinterpret_cast<size_t>(fps[f].depth) & 15) == 0;
        info.hizTma = tma;
    }
    hs.lap(1);
    // the fast brick kernel's per-frame constants and its preconditions (integrate_batch_impl.cuh: batch_bricks_fast_kernel)
    BrickFrames brickFrames;
    info.fastBricks = !perPixel && cam->width < (1 << 22) && cam->height < (1 << 22) && std::getenv("CHS_NO_FAST_BRICKS") == nullptr;
    for (int f = 0; f < K && info.fastBricks; f++)
    {
        const FrameParams &fp = fps[f];
        BrickFrame &b = brickFrames.f[f];
        std::memset(&b, 0, sizeof(b));
        std::memcpy(b.R, fp.cam.R, sizeof(b.R));
        std::memcpy(b.t, fp.cam.t, sizeof(b.t));
        b.fx = fp.cam.fx;
        b.fy = fp.cam.fy;
        b.cx = fp.cam.cx == 0.0f ? 0.0f : fp.cam.cx;           // -0.0f -> +0.0f: identical pixel decisions (see BrickFrame)
        b.cy = fp.cam.cy == 0.0f ? 0.0f : fp.cam.cy;
        b.Wf = fp.cam.Wf;
        b.Hf = fp.cam.Hf;
        b.W = fp.cam.W;
        b.pix_bias = (int)(0u - (0x4B000000u * (unsigned)fp.cam.W + 0x4B000000u));
        b.thr_band = fp.trunc_param + fp.diag;                  // binary32 sums, as the kernels form them (__fadd_rn)
        b.thr_carve = fp.carve ? fp.trunc_param + fp.carve_dist : INFINITY;
        b.wu = colorPath ? fp.wu_const : 1.0f;
        b.cutoff = fp.depth_cutoff;
        b.depth = fp.depth;
        b.color = fp.color_packed;
        b.carve_max = fp.sdf_carve_max;
        info.fastBricks = (!colorPath || fp.same_cam) && b.wu >= 0.125f && b.wu <= 1024.0f && b.thr_band <= 256.0f && std::isfinite(b.thr_band) &&
                          (b.thr_carve == b.thr_carve);
    }
    info.brickFrames = &brickFrames;
    hs.lap(6);
    host_launch_lap(-1);
    // distributed batch: every rank builds the Hi-Z pyramids of the frames it ingests; the others arrive by all-gather right after
    struct HizGather
    {
        chs_map *m;
        float2 *hiz;
        size_t bytesPerFrame;
        cudaStream_t st;
    } hizGather{m, bs.hiz, tiles * sizeof(float2), cs};
    // (measured: on 2 GPUs the second collective costs more than the Hi-Z kernel it saves; kept for N >= 8 experiments)
    if (dist && m->cfg.world > 1 && !computeTrunc && !anyMm && std::getenv("CHS_SHARD_HIZ") != nullptr)
    {
        info.hizFirst = locFirst;
        info.hizCount = locEnd - locFirst;
        info.afterHizCtx = &hizGather;
        info.afterHiz = [](void *p) -> int
        {
            HizGather *g = (HizGather *)p;
            return exchange_frames(g->m, g->hiz, g->bytesPerFrame, nullptr, 0, g->st);
        };
    }
    // colour packing runs beside the candidates kernel: on the copy stream (host frames: it follows the copies and the Hi-Z kernel
    // there; device frames: forked from the map's stream)
    // Distributed batches: copies and the NCCL exchange run on the copy stream, beside the kernels of the previous batch (the brick
    // kernel keeps a few SMs free for the NCCL kernels: BatchParams::reserve_sms); Hi-Z and colour packing then run on the map's
    // stream, on the whole GPU, once the exchange has landed.
    BatchStreams streams{cs, m->copyStream, st, bs.prepared, bs.packDone, bs.fork};
    if (dist)
    {
        if (push)
        {
            // the other ranks' pushes of this step have landed (flag words in this rank's arena, written behind the data): Hi-Z and
            // colour packing follow on the copy stream, i.e. beside the kernels of the previous step, as for one rank
            if (shardHiz)
            {
                // ... and so have the pyramids: only the frame table went over the copy stream
                CHS_CUDA(cudaEventRecord(bs.prepared, cs));
                CHS_CUDA(cudaStreamWaitEvent(st, bs.prepared, 0));
                streams.prep = st;
                streams.pack = st;
                info.skipHiz = true;
            }
            if ((rc = peer_wait(m, shardHiz ? st : cs, m->timeline ? bs.dTimeline : nullptr, shardHiz)))
                return rc;
            bp.peer_done = reinterpret_cast<unsigned *const *>(m->arena.base + 1024);
            bp.peer_world = m->cfg.world;
            bp.peer_step = m->arena.step;
        }
        else
        {
            // NCCL exchange: Hi-Z and colour packing run on the map's stream, on the whole GPU, once the all-gather has landed
            CHS_CUDA(cudaEventRecord(bs.prepared, cs));
            CHS_CUDA(cudaStreamWaitEvent(st, bs.prepared, 0));
            streams.prep = st;
            streams.pack = st;
        }
    }
    bp.reserve_sms = (dist && m->cfg.world > 1) ? 1 : 0;
    host_launch_lap(8);
    if (!poolLater)
        CHS_CUDA(launch_batch(bp, m->dm, info, m->evt, streams, 3));
    else
    {
        // large candidate box and a pool that could not take the worst case: run prepare + candidates, read how many chunks the
        // batch can create at most (its virtual candidates), size the pool for exactly that, then run the brick kernel
        CHS_CUDA(launch_batch(bp, m->dm, info, m->evt, streams, 1));
        BatchCounters hc;
        CHS_CUDA(cudaMemcpyAsync(&hc, bs.dBctr, sizeof(BatchCounters), cudaMemcpyDeviceToHost, st));
        CHS_CUDA(cudaStreamSynchronize(st));
        m->inflight.pop_back();                                     // everything before this batch has completed: retire it first
        if ((rc = poll_inflight(m, true)))
            return rc;
        Counters cc;
        CHS_CUDA(cudaMemcpyAsync(&cc, m->dCtr, sizeof(Counters), cudaMemcpyDeviceToHost, st));
        CHS_CUDA(cudaStreamSynchronize(st));
        m->knownChunks = cc.n_chunks;
        if ((rc = ensure_pool(m, (long long)cc.n_chunks + hc.new_count + 1024)))
            return rc;
        inf.newBound = hc.new_count;
        m->inflight.push_back(inf);
        bp.slot_batch = m->dSlotBatch;                              // the pool's side arrays may have moved
        CHS_CUDA(launch_batch(bp, m->dm, info, m->evt, streams, 2));
    }
    CHS_CUDA(cudaEventRecord(bs.released, st));
    host_launch_lap(6);
    hs.lap(7);
    bs.used = true;
    if (m->profiling)
        m->frameTimed = true;
    m->haveFrame = true;
    // the caller may reuse its host buffers as soon as we return (unless it promised not to: CHS_MEM_HOST_ASYNC)
    if (mem == CHS_MEM_HOST)
        CHS_CUDA(cudaEventSynchronize(bs.copied));
    return CHS_OK;
}

static int refresh_host_ids(chs_map *m, long long n)
{
    const long long have = (long long)m->hostIds.size() / 3;
    if (n > have)
    {
        m->hostIds.resize((size_t)n * 3);
        CHS_CUDA(cudaMemcpyAsync(m->hostIds.data() + have * 3, m->dm.slot_ids + have * 3, sizeof(int) * 3 * (size_t)(n - have),
                                 cudaMemcpyDeviceToHost, m->stream));
        CHS_CUDA(cudaStreamSynchronize(m->stream));
        for (long long s = have; s < n; s++)
            m->hostIndex[pack_id(m->hostIds[3 * s], m->hostIds[3 * s + 1], m->hostIds[3 * s + 2])] = (int)s;
    }
    return CHS_OK;
}

static int sync_counts(chs_map *m)
{
    CHS_CUDA(cudaSetDevice(m->device));
    CHS_CUDA(cudaStreamSynchronize(m->stream));
    int rc = poll_inflight(m, true);
    if (rc)
        return rc;
    // counters may have changed outside integrate (reset, update_meshes): read them directly
    Counters c;
    CHS_CUDA(cudaMemcpyAsync(&m->hCtr[chs_map::kRing], m->dCtr, sizeof(Counters), cudaMemcpyDeviceToHost, m->stream));
    CHS_CUDA(cudaStreamSynchronize(m->stream));
    c = m->hCtr[chs_map::kRing];
    m->knownChunks = c.n_chunks;
    m->knownDirty = c.n_dirty;
    if (c.error_flags)
        return fail(CHS_ERR_CAPACITY, "device table overflow, flags=" + std::to_string(c.error_flags));
    return CHS_OK;
}

} // namespace chs

// ---------------------------------------------------------------------------------------------------------
// C ABI

static int create_body(chs_map *m, const chs_config *cfg);

extern "C"
{

const char *chs_last_error_string(void) { return g_last_error.c_str(); }
int chs_abi_version(void) { return CHS_ABI_VERSION; }

int chs_create(const chs_config *cfg, chs_map **out)
{
    if (!cfg || !out)
        return fail(CHS_ERR_INVALID, "null argument");
    if (cfg->chunk_size != 8 && cfg->chunk_size != 16 && cfg->chunk_size != 32)
        return fail(CHS_ERR_INVALID, "chunk_size must be 8, 16 or 32 (cubic chunks only, quirk Q12)");
    if (!(cfg->resolution > 0.0f) || !std::isfinite(cfg->resolution))
        return fail(CHS_ERR_INVALID, "resolution must be positive");
    int count = 0;
    CHS_CUDA(cudaGetDeviceCount(&count));
    if (count <= 0)
        return fail(CHS_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
    chs_map *m = new chs_map();
    m->cfg = *cfg;
    if (m->cfg.world < 1)
        m->cfg.world = 1;
    if (m->cfg.rank < 0 || m->cfg.rank >= m->cfg.world)
    {
        delete m;
        return fail(CHS_ERR_INVALID, "rank outside [0, world)");
    }
    // everything that can fail runs in create_body: a failure there must not leak the half-built map
    const int rc = create_body(m, cfg);
    if (rc)
    {
        const std::string why = g_last_error;
        chs_destroy(m);
        g_last_error = why;
        return rc;
    }
    *out = m;
    return CHS_OK;
}

} // extern "C"

static int create_body(chs_map *m, const chs_config *cfg)
{
    if (cfg->device >= 0)
        m->device = cfg->device;
    else
        CHS_CUDA(cudaGetDevice(&m->device));
    CHS_CUDA(cudaSetDevice(m->device));
    if (cfg->stream)
        m->stream = (cudaStream_t)cfg->stream;
    else
    {
        CHS_CUDA(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
        m->ownStream = true;
    }
    DeviceMap &d = m->dm;
    d.cs = cfg->chunk_size;
    d.V = d.cs * d.cs * d.cs;
    d.res = cfg->resolution;
    d.half = cfg->resolution * 0.5f;
    d.use_color = cfg->use_color ? 1 : 0;
    d.rank = m->cfg.rank;
    d.world = m->cfg.world;
    CHS_CUDA(cudaMalloc((void **)&d.dist_slabs, sizeof(float2 *) * kMaxSlabs));
    CHS_CUDA(cudaMalloc((void **)&d.color_slabs, sizeof(uchar4 *) * kMaxSlabs));
    CHS_CUDA(cudaMalloc((void **)&m->dCtr, sizeof(Counters)));
    CHS_CUDA(cudaMemsetAsync(m->dCtr, 0, sizeof(Counters), m->stream));
    d.ctr = m->dCtr;
    CHS_CUDA(cudaHostAlloc((void **)&m->hCtr, sizeof(Counters) * (chs_map::kRing + 1), cudaHostAllocPortable | cudaHostAllocMapped));
    std::memset(m->hCtr, 0, sizeof(Counters) * (chs_map::kRing + 1));
    CHS_CUDA(cudaHostAlloc((void **)&m->hSnap, sizeof(HostSnapshot) * chs_map::kRing, cudaHostAllocPortable | cudaHostAllocMapped));
    std::memset(m->hSnap, 0, sizeof(HostSnapshot) * chs_map::kRing);
    CHS_CUDA(cudaMalloc((void **)&m->dHizTickets, sizeof(int) * (2 * kMaxBatch + 1)));
    CHS_CUDA(cudaMemsetAsync(m->dHizTickets, 0, sizeof(int) * (2 * kMaxBatch + 1), m->stream));
    CHS_CUDA(cudaStreamCreateWithFlags(&m->copyStream, cudaStreamNonBlocking));
    CHS_CUDA(cudaStreamCreateWithFlags(&m->pushStream, cudaStreamNonBlocking));
    CHS_CUDA(cudaEventCreateWithFlags(&m->callEvent, cudaEventDisableTiming));
    for (chs_map::BatchSet &bs : m->bset)
    {
        CHS_CUDA(cudaMalloc((void **)&bs.dFrames, sizeof(FrameParams) * kMaxBatch));
        CHS_CUDA(cudaMalloc((void **)&bs.dBctr, sizeof(BatchCounters)));
        CHS_CUDA(cudaMemsetAsync(bs.dBctr, 0, sizeof(BatchCounters), m->stream));
        CHS_CUDA(cudaMalloc((void **)&bs.dTimeline, sizeof(unsigned long long) * kTimelineStamps));
        CHS_CUDA(cudaMemsetAsync(bs.dTimeline, 0, sizeof(unsigned long long) * kTimelineStamps, m->stream));
        CHS_CUDA(cudaEventCreateWithFlags(&bs.copied, cudaEventDisableTiming));
        CHS_CUDA(cudaEventCreateWithFlags(&bs.prepared, cudaEventDisableTiming));
        CHS_CUDA(cudaEventCreateWithFlags(&bs.released, cudaEventDisableTiming));
        CHS_CUDA(cudaEventCreateWithFlags(&bs.packDone, cudaEventDisableTiming));
        CHS_CUDA(cudaEventCreateWithFlags(&bs.fork, cudaEventDisableTiming));
    }
    CHS_CUDA(cudaHostAlloc((void **)&m->hBatchSnap, sizeof(HostBatchSnapshot) * chs_map::kRing, cudaHostAllocPortable | cudaHostAllocMapped));
    std::memset(m->hBatchSnap, 0, sizeof(HostBatchSnapshot) * chs_map::kRing);
    CHS_CUDA(cudaHostAlloc((void **)&m->hFrameTables, sizeof(FrameParams) * kMaxBatch * chs_map::kRing, cudaHostAllocPortable));
    CHS_CUDA(cudaEventCreateWithFlags(&m->h2dDone, cudaEventDisableTiming));
    m->frameGraph = frame_graph_create();
    for (int i = 0; i < 8; i++)
        CHS_CUDA(cudaEventCreate(&m->evt[i]));
    const long long initial = cfg->initial_chunks > 0 ? cfg->initial_chunks : 4096;
    int rc;
    if ((rc = ensure_pool(m, initial)) || (rc = ensure_hash(m, initial)) || (rc = ensure_dirty(m, initial * 2)))
        return rc;
    CHS_CUDA(cudaStreamSynchronize(m->stream));
    return CHS_OK;
}

extern "C"
{

int chs_destroy(chs_map *m)
{
    if (!m)
        return CHS_OK;
    cudaSetDevice(m->device);
    if (m->copyStream)
        cudaStreamSynchronize(m->copyStream);
    if (m->pushStream)
        cudaStreamSynchronize(m->pushStream);
    cudaStreamSynchronize(m->stream);
    chs_comm_destroy(m);                        // first: collective when the map takes part in a peer-memory exchange; needs the streams
    for (float2 *p : m->distSlabs)
        cudaFreeAsync(p, m->stream);
    for (uchar4 *p : m->colorSlabs)
        cudaFreeAsync(p, m->stream);
    void *bufs[] = {m->dm.keys, m->dm.vals, m->dm.slot_ids, m->dm.brick_flags, m->dm.slot_epoch, m->dm.dirty_keys, m->dm.dirty_list, m->dDepth, m->dTrunc, m->dColor, m->dColorPacked, m->dHiz,
                    m->dUnits, m->dNews, m->dMeshSlots, m->dTriCounts, m->dGridCounts, m->dVertOffsets, m->dGridOffsets, m->dVerts, m->dNormals,
                    m->dColors, m->dGrids, m->dCfgScratch, m->dSlotBatch, m->bset[0].depthMm, m->bset[1].depthMm, m->bset[0].depth, m->bset[0].trunc, m->bset[0].color, m->bset[0].packed, m->bset[0].hiz,
                    m->bset[1].depth, m->bset[1].trunc, m->bset[1].color, m->bset[1].packed, m->bset[1].hiz};
    for (void *p : bufs)
        if (p)
            cudaFreeAsync(p, m->stream);
    cudaStreamSynchronize(m->stream);
    cudaFree(m->dm.dist_slabs);
    cudaFree(m->dm.color_slabs);
    cudaFree(m->dCtr);
    for (chs_map::BatchSet &bs : m->bset)
    {
        cudaFree(bs.dFrames);
        cudaFree(bs.dBctr);
        cudaFree(bs.dTimeline);
        if (bs.copied) cudaEventDestroy(bs.copied);
        if (bs.prepared) cudaEventDestroy(bs.prepared);
        if (bs.released) cudaEventDestroy(bs.released);
        if (bs.packDone) cudaEventDestroy(bs.packDone);
        if (bs.fork) cudaEventDestroy(bs.fork);
    }
    cudaFree(m->dHizTickets);
    if (m->callEvent)
        cudaEventDestroy(m->callEvent);
    if (m->copyStream)
        cudaStreamDestroy(m->copyStream);
    if (m->pushStream)
        cudaStreamDestroy(m->pushStream);
    if (m->uploadStream)
        cudaStreamDestroy(m->uploadStream);
    cudaFreeHost(m->hBatchSnap);
    cudaFreeHost(m->hFrameTables);
    cudaFreeHost(m->hCtr);
    cudaFreeHost(m->hSnap);
    if (m->dCommScratch)
        cudaFree(m->dCommScratch);
    for (void *p : {(void *)m->gathered.ids, (void *)m->gathered.vertOffsets, (void *)m->gathered.gridOffsets, (void *)m->gathered.verts,
                    (void *)m->gathered.normals, (void *)m->gathered.colors, (void *)m->gathered.grids})
        if (p)
            cudaFree(p);
    frame_graph_destroy(m->frameGraph);
    if (m->h2dDone)
        cudaEventDestroy(m->h2dDone);
    for (int i = 0; i < 8; i++)
        if (m->evt[i])
            cudaEventDestroy(m->evt[i]);
    if (m->ownStream)
        cudaStreamDestroy(m->stream);
    delete m;
    return CHS_OK;
}

// Chisel::Reset (OC Chisel.cpp:44-48): drop every chunk, mesh and dirty flag; keep the allocations.
int chs_reset(chs_map *m)
{
    if (!m)
        return fail(CHS_ERR_INVALID, "null map");
    CHS_CUDA(cudaSetDevice(m->device));
    int rc = poll_inflight(m, true);
    if (rc)
        return rc;
    if ((rc = sync_counts(m)))
        return rc;
    launch_fill_pool(m->dm, 0, (int)std::min<long long>(m->knownChunks, m->dm.capacity), m->stream);   // restore the pool invariant
    launch_fill_u64(m->dm.keys, m->hashSize, kEmptyKey, m->stream);
    launch_fill_i32(m->dm.vals, m->hashSize, -1, m->stream);
    launch_fill_u64(m->dm.dirty_keys, m->dirtySize, kEmptyKey, m->stream);
    CHS_CUDA(cudaMemsetAsync(m->dCtr, 0, sizeof(Counters), m->stream));
    CHS_CUDA(cudaStreamSynchronize(m->stream));
    m->knownChunks = m->knownDirty = 0;
    m->hostIds.clear();
    m->hostIndex.clear();
    m->lastMesh = chs_mesh_counts{};
    m->lastMeshChunks = 0;
    m->haveFrame = false;
    return CHS_OK;
}

int chs_synchronize(chs_map *m)
{
    if (!m)
        return fail(CHS_ERR_INVALID, "null map");
    CHS_CUDA(cudaSetDevice(m->device));
    CHS_CUDA(cudaStreamSynchronize(m->stream));
    return CHS_OK;
}

int chs_set_stream(chs_map *m, void *stream)
{
    if (!m)
        return fail(CHS_ERR_INVALID, "null map");
    CHS_CUDA(cudaSetDevice(m->device));
    CHS_CUDA(cudaStreamSynchronize(m->stream));
    if (m->ownStream)
    {
        cudaStreamDestroy(m->stream);
        m->ownStream = false;
    }
    if (stream)
        m->stream = (cudaStream_t)stream;
    else
    {
        CHS_CUDA(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
        m->ownStream = true;
    }
    return CHS_OK;
}

int chs_get_device_timeline(chs_map *m, long long *out, int max_batches, int *n_batches)
{
    if (!m || !out || !n_batches || max_batches < 0)
        return fail(CHS_ERR_INVALID, "null argument");
    CHS_CUDA(cudaSetDevice(m->device));
    int rc = poll_inflight(m, true);
    if (rc)
        return rc;
    const int n = (int)std::min<size_t>(m->timelineHist.size(), (size_t)max_batches);
    for (int i = 0; i < n; i++)
        std::memcpy(out + (size_t)i * kTimelineStamps, m->timelineHist[m->timelineHist.size() - (size_t)n + (size_t)i].data(), sizeof(long long) * kTimelineStamps);
    *n_batches = n;
    return CHS_OK;
}

int chs_set_profiling(chs_map *m, int enabled)
{
    if (!m)
        return fail(CHS_ERR_INVALID, "null map");
    m->profiling = (enabled & 1) != 0;
    m->timeline = (enabled & 2) != 0;
    m->timelineHist.clear();
    m->frameTimed = m->meshTimed = false;
    return CHS_OK;
}

int chs_integrate_depth(chs_map *m, const chs_integrator *integ, const float *depth, int mem, const float pose[12], const chs_camera *cam)
{
    return integrate_common(m, integ, depth, mem, pose, cam, nullptr, 0, nullptr, nullptr, false);
}

int chs_integrate_depth_color(chs_map *m, const chs_integrator *integ, const float *depth, int mem, const float pose[12],
                              const chs_camera *cam, const uint8_t *color, int channels, const float cpose[12], const chs_camera *ccam)
{
    return integrate_common(m, integ, depth, mem, pose, cam, color, channels, cpose, ccam, true);
}

int chs_integrate_batch(chs_map *m, const chs_integrator *integ, int n, const chs_frame *frames, int mem, const chs_camera *cam, int channels,
                        const chs_camera *ccam)
{
    if (!m || !integ || !frames || !cam || n < 0)
        return fail(CHS_ERR_INVALID, "null argument");
    const bool colorPath = ccam != nullptr;
    if (colorPath && (channels < 1 || channels > 4 || ccam->width <= 0 || ccam->height <= 0))
        return fail(CHS_ERR_INVALID, "bad colour arguments");
    if (mem < CHS_MEM_HOST || mem > CHS_MEM_DEVICE_ASYNC)
        return fail(CHS_ERR_INVALID, "bad memory space");
    bool fusable = true, anyMm = false;
    for (int f = 0; f < n; f++)
    {
        if ((!frames[f].depth && !frames[f].depth_mm) || (colorPath && !frames[f].color) || (integ->trunc_kind == CHS_TRUNC_PER_PIXEL && !frames[f].trunc_per_pixel))
            return fail(CHS_ERR_INVALID, "frame without depth / colour / truncation image");
        anyMm |= frames[f].depth_mm != nullptr;
        if (!finite12(frames[f].pose) || (colorPath && !finite12(frames[f].color_pose)))
            return fail(CHS_ERR_INVALID, "non-finite pose (quirk Q14: rejected at the boundary)");
        // the fused kernels reuse the depth projection for the colour lookup
        if (colorPath && (std::memcmp(frames[f].pose, frames[f].color_pose, sizeof(float) * 12) != 0 || std::memcmp(cam, ccam, 4 * sizeof(float) + 2 * sizeof(int)) != 0))
            fusable = false;
    }
    if (anyMm && !fusable)
        return fail(CHS_ERR_INVALID, "depth_mm frames need the colour camera to coincide with the depth camera (they only run through the fused kernels)");
    CHS_CUDA(cudaSetDevice(m->device));
    int rc = poll_inflight(m, false);
    if (rc)
        return rc;
    {
        chs_map::CallStats &c = m->callStats[++m->callId & 3];
        c.id = m->callId;
        c.pending = n;
        c.st.assign((size_t)n, chs_frame_stats{});
    }
    int f = 0;
    while (f < n)
    {
        const int K = std::min(n - f, (int)kMaxBatch);
        rc = CHS_ERR_NOT_FOUND;
        if (fusable && (K >= 2 || anyMm))
            rc = integrate_batch_fused(m, integ, K, frames + f, mem, cam, channels, ccam, colorPath, f);
        if (rc == CHS_ERR_NOT_FOUND && anyMm)
            return fail(CHS_ERR_INVALID, "depth_mm frames too far apart to share a candidate box");
        if (rc == CHS_ERR_NOT_FOUND)
        {
            // frame by frame (single frame, separate colour camera, or frames too far apart to share a candidate box)
            for (int j = f; j < f + K; j++)
            {
                chs_integrator one = *integ;
                one.trunc_per_pixel = frames[j].trunc_per_pixel;
                if ((rc = integrate_common(m, &one, frames[j].depth, mem == CHS_MEM_HOST_ASYNC ? CHS_MEM_HOST : (mem == CHS_MEM_DEVICE_ASYNC ? CHS_MEM_DEVICE : mem), frames[j].pose, cam, frames[j].color, channels, frames[j].color_pose, ccam, colorPath, j)))
                    return rc;
            }
        }
        else if (rc)
            return rc;
        f += K;
    }
    return CHS_OK;
}

static int copy_call_stats(chs_map *m, int callId, chs_frame_stats *out, int cap, int *n)
{
    chs_map::CallStats *cs = call_stats(m, callId);
    if (!cs)
    {
        *n = 0;
        return callId == 0 ? CHS_OK : fail(CHS_ERR_NOT_FOUND, "the counters of that chs_integrate_batch call are no longer kept (only the last 4 calls are)");
    }
    *n = (int)cs->st.size();
    int flags = 0;
    for (int i = 0; i < *n && i < cap && out; i++)
    {
        out[i] = cs->st[i];
        flags |= (int)out[i].error_flags;
    }
    if (flags)
        return fail(CHS_ERR_CAPACITY, "device table overflow, flags=" + std::to_string(flags));
    return CHS_OK;
}

int chs_get_batch_stats(chs_map *m, chs_frame_stats *out, int cap, int *n)
{
    if (!m || !n)
        return fail(CHS_ERR_INVALID, "null argument");
    CHS_CUDA(cudaSetDevice(m->device));
    int rc = poll_inflight(m, true);
    if (rc)
        return rc;
    return copy_call_stats(m, m->callId, out, cap, n);
}

int chs_last_batch_ticket(chs_map *m, int64_t *ticket)
{
    if (!m || !ticket)
        return fail(CHS_ERR_INVALID, "null argument");
    *ticket = m->callId;
    return CHS_OK;
}

// Wait for ONE chs_integrate_batch call (not for the calls issued after it) and return its per-frame counters: lets a caller
// keep the next batch's copies and kernels in flight while it reads the previous batch's result.
int chs_wait_batch(chs_map *m, int64_t ticket, chs_frame_stats *out, int cap, int *n)
{
    if (!m || !n)
        return fail(CHS_ERR_INVALID, "null argument");
    CHS_CUDA(cudaSetDevice(m->device));
    chs_map::CallStats *cs = call_stats(m, (int)ticket);
    if (!cs)
        return copy_call_stats(m, (int)ticket, out, cap, n);
    int rc;
    for (int spin = 0; cs->pending > 0; spin++)
    {
        if ((rc = poll_inflight(m, false)))
            return rc;
        if (cs->pending <= 0)
            break;
        if (spin > 2000000)
        {
            // the snapshot is written by the last CTA of the call's last kernel: fall back to a full synchronisation
            if ((rc = poll_inflight(m, true)))
                return rc;
            break;
        }
        _mm_pause();
    }
    return copy_call_stats(m, (int)ticket, out, cap, n);
}

int chs_get_frame_stats(chs_map *m, chs_frame_stats *out)
{
    if (!m || !out)
        return fail(CHS_ERR_INVALID, "null argument");
    CHS_CUDA(cudaSetDevice(m->device));
    int rc = poll_inflight(m, true);
    if (rc)
        return rc;
    const HostSnapshot &c = m->lastFrame;
    *out = stats_of(c);
    if (c.error_flags)
        return fail(CHS_ERR_CAPACITY, "device table overflow, flags=" + std::to_string(c.error_flags));
    return CHS_OK;
}

int chs_get_timings(chs_map *m, chs_timings *out)
{
    if (!m || !out)
        return fail(CHS_ERR_INVALID, "null argument");
    CHS_CUDA(cudaSetDevice(m->device));
    CHS_CUDA(cudaStreamSynchronize(m->stream));
    std::memset(out, 0, sizeof(*out));
    if (m->frameTimed)
    {
        CHS_CUDA(cudaEventElapsedTime(&out->prepare_ms, m->evt[0], m->evt[1]));
        CHS_CUDA(cudaEventElapsedTime(&out->candidates_ms, m->evt[1], m->evt[2]));
        CHS_CUDA(cudaEventElapsedTime(&out->new_chunks_ms, m->evt[2], m->evt[7]));
        CHS_CUDA(cudaEventElapsedTime(&out->integrate_ms, m->evt[7], m->evt[3]));
        CHS_CUDA(cudaEventElapsedTime(&out->frame_ms, m->evt[0], m->evt[3]));
    }
    {
        int rc = poll_inflight(m, true);
        if (rc)
            return rc;
        out->bricks_span_ms = (float)((double)m->lastBricksSpanNs * 1e-6);
    }
    if (m->meshTimed)
    {
        CHS_CUDA(cudaEventElapsedTime(&out->mesh_count_ms, m->evt[4], m->evt[5]));
        CHS_CUDA(cudaEventElapsedTime(&out->mesh_emit_ms, m->evt[5], m->evt[6]));
        CHS_CUDA(cudaEventElapsedTime(&out->mesh_ms, m->evt[4], m->evt[6]));
    }
    return CHS_OK;
}

int chs_num_chunks(chs_map *m, int64_t *n)
{
    if (!m || !n)
        return fail(CHS_ERR_INVALID, "null argument");
    int rc = sync_counts(m);
    if (rc)
        return rc;
    *n = m->knownChunks;
    return CHS_OK;
}

int chs_chunk_ids(chs_map *m, int32_t *ids, int64_t cap)
{
    if (!m || !ids)
        return fail(CHS_ERR_INVALID, "null argument");
    int rc = sync_counts(m);
    if (rc)
        return rc;
    if ((rc = refresh_host_ids(m, m->knownChunks)))
        return rc;
    const int64_t n = std::min<int64_t>(cap, m->knownChunks);
    std::memcpy(ids, m->hostIds.data(), sizeof(int32_t) * 3 * (size_t)n);
    return CHS_OK;
}

int chs_has_chunk(chs_map *m, const int32_t id[3], int *found)
{
    if (!m || !id || !found)
        return fail(CHS_ERR_INVALID, "null argument");
    int rc = sync_counts(m);
    if (rc)
        return rc;
    if ((rc = refresh_host_ids(m, m->knownChunks)))
        return rc;
    *found = 0;
    if (id[0] < -kIdBias || id[0] >= kIdBias || id[1] < -kIdBias || id[1] >= kIdBias || id[2] < -kIdBias || id[2] >= kIdBias)
        return CHS_OK;
    *found = m->hostIndex.find(pack_id(id[0], id[1], id[2])) != m->hostIndex.end() ? 1 : 0;
    return CHS_OK;
}

static int download_slot(chs_map *m, int slot, float *sdf, float *weight, uint8_t *rgbw)
{
    const int V = m->dm.V;
    std::vector<float2> tmp((size_t)V);
    const float2 *src = m->distSlabs[slot >> kSlabChunksLog2] + (size_t)(slot & (kSlabChunks - 1)) * V;
    CHS_CUDA(cudaMemcpyAsync(tmp.data(), src, sizeof(float2) * V, cudaMemcpyDeviceToHost, m->stream));
    if (rgbw && m->cfg.use_color)
    {
        const uchar4 *cs = m->colorSlabs[slot >> kSlabChunksLog2] + (size_t)(slot & (kSlabChunks - 1)) * V;
        CHS_CUDA(cudaMemcpyAsync(rgbw, cs, 4 * (size_t)V, cudaMemcpyDeviceToHost, m->stream));
    }
    CHS_CUDA(cudaStreamSynchronize(m->stream));
    for (int i = 0; i < V; i++)
    {
        if (sdf)
            sdf[i] = tmp[i].x;
        if (weight)
            weight[i] = tmp[i].y;
    }
    return CHS_OK;
}

int chs_download_chunk(chs_map *m, const int32_t id[3], float *sdf, float *weight, uint8_t *rgbw)
{
    if (!m || !id)
        return fail(CHS_ERR_INVALID, "null argument");
    int rc = sync_counts(m);
    if (rc)
        return rc;
    if ((rc = refresh_host_ids(m, m->knownChunks)))
        return rc;
    if (id[0] < -kIdBias || id[0] >= kIdBias || id[1] < -kIdBias || id[1] >= kIdBias || id[2] < -kIdBias || id[2] >= kIdBias)
        return fail(CHS_ERR_NOT_FOUND, "no such chunk");
    auto it = m->hostIndex.find(pack_id(id[0], id[1], id[2]));
    if (it == m->hostIndex.end())
        return fail(CHS_ERR_NOT_FOUND, "no such chunk");
    return download_slot(m, it->second, sdf, weight, rgbw);
}

int chs_download_all(chs_map *m, int64_t cap, int32_t *ids, float *sdf, float *weight, uint8_t *rgbw)
{
    if (!m)
        return fail(CHS_ERR_INVALID, "null map");
    int rc = sync_counts(m);
    if (rc)
        return rc;
    if ((rc = refresh_host_ids(m, m->knownChunks)))
        return rc;
    const int64_t n = std::min<int64_t>(cap, m->knownChunks);
    const int V = m->dm.V;
    if (ids)
        std::memcpy(ids, m->hostIds.data(), sizeof(int32_t) * 3 * (size_t)n);
    std::vector<float2> tmp;
    for (int64_t s0 = 0; s0 < n; s0 += kSlabChunks)
    {
        const int64_t cnt = std::min<int64_t>(kSlabChunks, n - s0);
        const size_t slab = (size_t)(s0 >> kSlabChunksLog2);
        if (sdf || weight)
        {
            tmp.resize((size_t)cnt * V);
            CHS_CUDA(cudaMemcpyAsync(tmp.data(), m->distSlabs[slab], sizeof(float2) * (size_t)cnt * V, cudaMemcpyDeviceToHost, m->stream));
        }
        if (rgbw && m->cfg.use_color)
            CHS_CUDA(cudaMemcpyAsync(rgbw + (size_t)s0 * V * 4, m->colorSlabs[slab], 4 * (size_t)cnt * V, cudaMemcpyDeviceToHost, m->stream));
        CHS_CUDA(cudaStreamSynchronize(m->stream));
        if (sdf || weight)
            for (size_t i = 0; i < (size_t)cnt * V; i++)
            {
                if (sdf)
                    sdf[(size_t)s0 * V + i] = tmp[i].x;
                if (weight)
                    weight[(size_t)s0 * V + i] = tmp[i].y;
            }
    }
    return CHS_OK;
}

// ---- chunk export / import / explicit dirty set: the primitives behind sharded meshing (ghost chunks) and map checkpoints ----

__global__ void import_insert_kernel(DeviceMap map, const int *ids, int firstSlot, int n)
{
    // append n new chunks (IDs known to be absent) at slots firstSlot .. firstSlot + n - 1
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const int s = firstSlot + i;
        map.slot_ids[3 * s] = ids[3 * i];
        map.slot_ids[3 * s + 1] = ids[3 * i + 1];
        map.slot_ids[3 * s + 2] = ids[3 * i + 2];
        map.brick_flags[s] = ~0ull;              // unknown history: assume every brick may hold a carvable voxel (conservative)
        map.slot_epoch[s] = 0;
        hash_insert_new(map, pack_id(ids[3 * i], ids[3 * i + 1], ids[3 * i + 2]), s);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        map.ctr->n_chunks = firstSlot + n;
}

__global__ void set_dirty_kernel(DeviceMap map, const int *ids, int n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        dirty_insert(map, pack_id(ids[3 * i], ids[3 * i + 1], ids[3 * i + 2]));
}

static bool id_in_range(const int32_t *id)
{
    return id[0] >= -kIdBias && id[0] < kIdBias && id[1] >= -kIdBias && id[1] < kIdBias && id[2] >= -kIdBias && id[2] < kIdBias;
}

// Voxels of the listed chunks (host buffers, n*V each; rgbw n*4V or NULL). found[i] = 0 for IDs this map does not hold
// (their output rows are left untouched).
int chs_export_chunks(chs_map *m, int64_t n, const int32_t *ids, uint8_t *found, float *sdf, float *weight, uint8_t *rgbw)
{
    if (!m || !ids || !found)
        return fail(CHS_ERR_INVALID, "null argument");
    int rc = sync_counts(m);
    if (rc)
        return rc;
    if ((rc = refresh_host_ids(m, m->knownChunks)))
        return rc;
    const int V = m->dm.V;
    std::vector<float2> tmp((size_t)V);
    for (int64_t i = 0; i < n; i++)
    {
        found[i] = 0;
        if (!id_in_range(ids + 3 * i))
            continue;
        auto it = m->hostIndex.find(pack_id(ids[3 * i], ids[3 * i + 1], ids[3 * i + 2]));
        if (it == m->hostIndex.end())
            continue;
        found[i] = 1;
        if ((rc = download_slot(m, it->second, sdf ? sdf + (size_t)i * V : nullptr, weight ? weight + (size_t)i * V : nullptr,
                                rgbw ? rgbw + (size_t)i * V * 4 : nullptr)))
            return rc;
    }
    return CHS_OK;
}

// One CTA per imported chunk: voxels from the staging buffers into the chunk's pool slot. Every imported chunk -- new or
// overwritten -- gets all brick flags set: its history is unknown, so every brick may hold a carvable voxel (conservative; an
// overwritten chunk that kept its old flags could make later free-space frames skip a brick that needs carving).
__global__ void import_scatter_kernel(DeviceMap map, int n, const int *slots, const float *sdf, const float *weight, const unsigned *rgbw)
{
    const int i = blockIdx.x;
    if (i >= n)
        return;
    const int s = slots[i];
    float2 *dst = dist_ptr(map, s);
    const float *ps = sdf + (size_t)i * map.V, *pw = weight + (size_t)i * map.V;
    for (int v = threadIdx.x; v < map.V; v += blockDim.x)
        dst[v] = make_float2(ps[v], pw[v]);
    if (map.use_color)
    {
        unsigned *cd = reinterpret_cast<unsigned *>(color_ptr(map, s));
        for (int v = threadIdx.x; v < map.V; v += blockDim.x)
            cd[v] = rgbw ? rgbw[(size_t)i * map.V + v] : 0u;
    }
    if (threadIdx.x == 0)
        map.brick_flags[s] = ~0ull;
}

// Insert or overwrite chunks from host buffers (n*V each; rgbw may be NULL). Existing chunks keep their slot. The voxels travel
// in slices of a few thousand chunks: three bulk copies and one scatter kernel per slice, one synchronisation per slice (the
// host staging of the slice is reused) -- not one per chunk.
int chs_import_chunks(chs_map *m, int64_t n, const int32_t *ids, const float *sdf, const float *weight, const uint8_t *rgbw)
{
    if (!m || !ids || !sdf || !weight)
        return fail(CHS_ERR_INVALID, "null argument");
    int rc = sync_counts(m);
    if (rc)
        return rc;
    if ((rc = refresh_host_ids(m, m->knownChunks)))
        return rc;
    const int V = m->dm.V;
    std::vector<int> slots((size_t)n);
    std::vector<int32_t> newIds;
    std::unordered_map<unsigned long long, int> fresh;      // IDs this call adds (committed to the host mirror only after the device insert)
    long long next = m->knownChunks;
    for (int64_t i = 0; i < n; i++)
    {
        if (!id_in_range(ids + 3 * i))
            return fail(CHS_ERR_INVALID, "chunk IDs outside [-2^20, 2^20)");
        const unsigned long long key = pack_id(ids[3 * i], ids[3 * i + 1], ids[3 * i + 2]);
        auto it = m->hostIndex.find(key);
        if (it != m->hostIndex.end())
            slots[(size_t)i] = it->second;
        else
        {
            auto jt = fresh.find(key);
            if (jt != fresh.end())
                slots[(size_t)i] = jt->second;               // the same new ID twice in one call: the later entry wins
            else
            {
                slots[(size_t)i] = (int)next;
                fresh[key] = (int)next++;
                for (int k = 0; k < 3; k++)
                    newIds.push_back(ids[3 * i + k]);
            }
        }
    }
    const long long nNew = next - m->knownChunks;
    if ((rc = ensure_pool(m, next + 1)) || (rc = ensure_hash(m, next + 1)))
        return rc;
    cudaStream_t st = m->stream;
    if (nNew > 0)
    {
        int *dIds = nullptr;
        CHS_CUDA(cudaMallocAsync((void **)&dIds, sizeof(int) * 3 * (size_t)nNew, st));
        CHS_CUDA(cudaMemcpyAsync(dIds, newIds.data(), sizeof(int) * 3 * (size_t)nNew, cudaMemcpyHostToDevice, st));
        import_insert_kernel<<<(int)std::min<long long>((nNew + 255) / 256, 148), 256, 0, st>>>(m->dm, dIds, (int)m->knownChunks, (int)nNew);
        CHS_CUDA(cudaGetLastError());
        CHS_CUDA(cudaFreeAsync(dIds, st));
    }
    const int64_t slice = 2048;
    int *dSlots = nullptr;
    float *dSdf = nullptr, *dW = nullptr;
    unsigned *dCol = nullptr;
    const size_t sliceVox = (size_t)std::min<int64_t>(slice, std::max<int64_t>(n, 1)) * V;
    CHS_CUDA(cudaMallocAsync((void **)&dSlots, sizeof(int) * (size_t)std::min<int64_t>(slice, std::max<int64_t>(n, 1)), st));
    CHS_CUDA(cudaMallocAsync((void **)&dSdf, sizeof(float) * sliceVox, st));
    CHS_CUDA(cudaMallocAsync((void **)&dW, sizeof(float) * sliceVox, st));
    if (m->cfg.use_color && rgbw)
        CHS_CUDA(cudaMallocAsync((void **)&dCol, sizeof(unsigned) * sliceVox, st));
    for (int64_t i0 = 0; i0 < n; i0 += slice)
    {
        const int64_t cnt = std::min<int64_t>(slice, n - i0);
        CHS_CUDA(cudaMemcpyAsync(dSlots, slots.data() + i0, sizeof(int) * (size_t)cnt, cudaMemcpyHostToDevice, st));
        CHS_CUDA(cudaMemcpyAsync(dSdf, sdf + (size_t)i0 * V, sizeof(float) * (size_t)cnt * V, cudaMemcpyHostToDevice, st));
        CHS_CUDA(cudaMemcpyAsync(dW, weight + (size_t)i0 * V, sizeof(float) * (size_t)cnt * V, cudaMemcpyHostToDevice, st));
        if (dCol)
            CHS_CUDA(cudaMemcpyAsync(dCol, rgbw + (size_t)i0 * V * 4, 4 * (size_t)cnt * V, cudaMemcpyHostToDevice, st));
        import_scatter_kernel<<<(int)cnt, 256, 0, st>>>(m->dm, (int)cnt, dSlots, dSdf, dW, dCol);
        CHS_CUDA(cudaGetLastError());
        CHS_CUDA(cudaStreamSynchronize(st));                         // the staging buffers are reused by the next slice
    }
    for (void *p : {(void *)dSlots, (void *)dSdf, (void *)dW, (void *)dCol})
        if (p)
            CHS_CUDA(cudaFreeAsync(p, st));
    CHS_CUDA(cudaStreamSynchronize(st));
    // the device insert has succeeded: now the host mirror may know the new chunks
    for (const auto &kv : fresh)
        m->hostIndex[kv.first] = kv.second;
    m->hostIds.insert(m->hostIds.end(), newIds.begin(), newIds.end());
    m->knownChunks = next;
    return CHS_OK;
}

// ---- map checkpoint on disk (replaces the reference's broken chunk wire format, CR Serialization.h:31-84) ----
// Layout (little endian): header {magic "CHSMAP01", int32 version = 1, chunk_size, use_color, float resolution, int64 n_chunks, int64
// n_dirty, int32 voxels per chunk, int32 reserved}, then SoA payload: ids int32[3 n], sdf float[n V], weight float[n V],
// rgbw uint8[4 n V] (colour maps only), dirty ids int32[3 n_dirty]. Chunks in pool order; everything a resumed run needs to
// continue bit-identically (brick flags are a conservative cache and are rebuilt as "all set").
namespace
{
struct CheckpointHeader
{
    char magic[8];
    int32_t version, chunk_size, use_color;
    float resolution;
    int64_t n_chunks, n_dirty;
    int32_t voxels, reserved;
};
} // namespace

int chs_save_map(chs_map *m, const char *path)
{
    if (!m || !path)
        return fail(CHS_ERR_INVALID, "null argument");
    int rc = sync_counts(m);
    if (rc)
        return rc;
    const int64_t n = m->knownChunks, nd = m->knownDirty;
    const size_t V = (size_t)m->dm.V;
    std::vector<int32_t> ids((size_t)n * 3), dirty((size_t)nd * 3);
    std::vector<float> sdf((size_t)n * V), w((size_t)n * V);
    std::vector<uint8_t> rgbw(m->cfg.use_color ? (size_t)n * V * 4 : 0);
    if (n && (rc = chs_download_all(m, n, ids.data(), sdf.data(), w.data(), m->cfg.use_color ? rgbw.data() : nullptr)))
        return rc;
    if (nd && (rc = chs_dirty_ids(m, dirty.data(), nd)))
        return rc;
    CheckpointHeader h{};
    std::memcpy(h.magic, "CHSMAP01", 8);
    h.version = 1;
    h.chunk_size = m->cfg.chunk_size;
    h.use_color = m->cfg.use_color ? 1 : 0;
    h.resolution = m->cfg.resolution;
    h.n_chunks = n;
    h.n_dirty = nd;
    h.voxels = (int32_t)V;
    FILE *f = std::fopen(path, "wb");
    if (!f)
        return fail(CHS_ERR_INVALID, std::string("cannot open ") + path + " for writing");
    bool ok = std::fwrite(&h, sizeof(h), 1, f) == 1;
    auto put = [&](const void *p, size_t bytes) { ok = ok && (bytes == 0 || std::fwrite(p, 1, bytes, f) == bytes); };
    put(ids.data(), ids.size() * 4);
    put(sdf.data(), sdf.size() * 4);
    put(w.data(), w.size() * 4);
    put(rgbw.data(), rgbw.size());
    put(dirty.data(), dirty.size() * 4);
    ok = (std::fclose(f) == 0) && ok;
    return ok ? CHS_OK : fail(CHS_ERR_INVALID, std::string("short write to ") + path);
}

int chs_load_map(chs_map *m, const char *path)
{
    if (!m || !path)
        return fail(CHS_ERR_INVALID, "null argument");
    FILE *f = std::fopen(path, "rb");
    if (!f)
        return fail(CHS_ERR_NOT_FOUND, std::string("cannot open ") + path);
    CheckpointHeader h{};
    if (std::fread(&h, sizeof(h), 1, f) != 1 || std::memcmp(h.magic, "CHSMAP01", 8) != 0 || h.version != 1)
    {
        std::fclose(f);
        return fail(CHS_ERR_INVALID, "not a chisel_b200 map checkpoint (version 1)");
    }
    if (h.chunk_size != m->cfg.chunk_size || h.resolution != m->cfg.resolution || (h.use_color != 0) != (m->cfg.use_color != 0) || h.voxels != m->dm.V ||
        h.n_chunks < 0 || h.n_dirty < 0)
    {
        std::fclose(f);
        return fail(CHS_ERR_INVALID, "the checkpoint was written by a map with another chunk size, resolution or colour setting");
    }
    const size_t n = (size_t)h.n_chunks, nd = (size_t)h.n_dirty, V = (size_t)h.voxels;
    std::vector<int32_t> ids(n * 3), dirty(nd * 3);
    std::vector<float> sdf(n * V), w(n * V);
    std::vector<uint8_t> rgbw(h.use_color ? n * V * 4 : 0);
    bool ok = true;
    auto get = [&](void *p, size_t bytes) { ok = ok && (bytes == 0 || std::fread(p, 1, bytes, f) == bytes); };
    get(ids.data(), ids.size() * 4);
    get(sdf.data(), sdf.size() * 4);
    get(w.data(), w.size() * 4);
    get(rgbw.data(), rgbw.size());
    get(dirty.data(), dirty.size() * 4);
    std::fclose(f);
    if (!ok)
        return fail(CHS_ERR_INVALID, "truncated checkpoint");
    int rc = chs_reset(m);
    if (rc)
        return rc;
    if (n && (rc = chs_import_chunks(m, (int64_t)n, ids.data(), sdf.data(), w.data(), h.use_color ? rgbw.data() : nullptr)))
        return rc;
    return chs_set_dirty(m, (int64_t)nd, dirty.data());
}

// Replace the dirty set by the given IDs (Chisel::meshesToUpdate assigned from outside).
int chs_set_dirty(chs_map *m, int64_t n, const int32_t *ids)
{
    if (!m || (n > 0 && !ids))
        return fail(CHS_ERR_INVALID, "null argument");
    int rc = sync_counts(m);
    if (rc)
        return rc;
    for (int64_t i = 0; i < n; i++)
        if (!id_in_range(ids + 3 * i))
            return fail(CHS_ERR_INVALID, "chunk IDs outside [-2^20, 2^20)");
    cudaStream_t st = m->stream;
    if ((rc = ensure_dirty(m, n + 16)))
        return rc;
    launch_fill_u64(m->dm.dirty_keys, m->dirtySize, kEmptyKey, st);
    CHS_CUDA(cudaMemsetAsync(&m->dCtr->n_dirty, 0, sizeof(int), st));
    if (n > 0)
    {
        int *dIds = nullptr;
        CHS_CUDA(cudaMallocAsync((void **)&dIds, sizeof(int) * 3 * (size_t)n, st));
        CHS_CUDA(cudaMemcpyAsync(dIds, ids, sizeof(int) * 3 * (size_t)n, cudaMemcpyHostToDevice, st));
        set_dirty_kernel<<<(int)std::min<long long>((n + 255) / 256, 148), 256, 0, st>>>(m->dm, dIds, (int)n);
        CHS_CUDA(cudaGetLastError());
        CHS_CUDA(cudaFreeAsync(dIds, st));
    }
    CHS_CUDA(cudaStreamSynchronize(st));
    m->knownDirty = n;
    return CHS_OK;
}

int chs_num_dirty(chs_map *m, int64_t *n)
{
    if (!m || !n)
        return fail(CHS_ERR_INVALID, "null argument");
    int rc = sync_counts(m);
    if (rc)
        return rc;
    *n = m->knownDirty;
    return CHS_OK;
}

int chs_dirty_ids(chs_map *m, int32_t *ids, int64_t cap)
{
    if (!m || !ids)
        return fail(CHS_ERR_INVALID, "null argument");
    int rc = sync_counts(m);
    if (rc)
        return rc;
    const int64_t n = std::min<int64_t>(cap, m->knownDirty);
    std::vector<unsigned long long> keys((size_t)n);
    if (n)
    {
        CHS_CUDA(cudaMemcpyAsync(keys.data(), m->dm.dirty_list, sizeof(unsigned long long) * (size_t)n, cudaMemcpyDeviceToHost, m->stream));
        CHS_CUDA(cudaStreamSynchronize(m->stream));
    }
    for (int64_t i = 0; i < n; i++)
        unpack_id(keys[(size_t)i], &ids[3 * i], &ids[3 * i + 1], &ids[3 * i + 2]);
    return CHS_OK;
}

// Chisel::UpdateMeshes without its every-10th gate (OC Chisel.cpp:50-59): RecomputeMeshes(meshesToUpdate), clear.
int chs_update_meshes(chs_map *m)
{
    if (m)
        m->meshGathered = m->meshFromGhost = false;
    if (!m)
        return fail(CHS_ERR_INVALID, "null map");
    int rc = sync_counts(m);
    if (rc)
        return rc;
    cudaStream_t st = m->stream;
    const long long nd = m->knownDirty;
    m->lastMesh = chs_mesh_counts{};
    m->lastMesh.has_colors = m->cfg.use_color ? 1 : 0;
    m->lastMeshChunks = 0;
    if (nd == 0)
        return CHS_OK;
    if ((size_t)nd > m->meshChunkCap)
    {
        const size_t want = (size_t)nd + (size_t)nd / 2;
        size_t c0 = m->meshChunkCap, c1 = m->meshChunkCap, c2 = m->meshChunkCap;
        if ((rc = grow_buffer(&m->dMeshSlots, &c0, want, st)) || (rc = grow_buffer(&m->dTriCounts, &c1, want, st)) ||
            (rc = grow_buffer(&m->dGridCounts, &c2, want, st)))
            return rc;
        size_t o0 = m->meshChunkCap ? m->meshChunkCap + 1 : 0, o1 = o0;
        if ((rc = grow_buffer(&m->dVertOffsets, &o0, want + 1, st)) || (rc = grow_buffer(&m->dGridOffsets, &o1, want + 1, st)))
            return rc;
        m->meshChunkCap = want;
    }
    if ((rc = grow_buffer(&m->dCfgScratch, &m->cfgScratchCap, (size_t)nd * m->dm.V, st)))
        return rc;
    MeshParams mp{};
    mp.cfg_scratch = m->dCfgScratch;
    mp.dirty_list = m->dm.dirty_list;
    mp.n_dirty = (int)nd;
    mp.mesh_slots = m->dMeshSlots;
    mp.tri_counts = m->dTriCounts;
    mp.grid_counts = m->dGridCounts;
    mp.vert_offsets = m->dVertOffsets;
    mp.grid_offsets = m->dGridOffsets;
    {
        float T = (float)1e-12;
        if ((double)T > 1e-12)
            T = std::nextafterf(T, -INFINITY);
        mp.w_observed_min = T;                                  // weight > 1e-12 (double)  <=>  weight > T
    }
    CHS_CUDA(cudaMemsetAsync(&m->dCtr->mesh_chunks, 0, sizeof(int), st));
    if (m->profiling)
        CHS_CUDA(cudaEventRecord(m->evt[4], st));
    launch_mesh_select(mp, m->dm, st);
    launch_mesh_count(mp, m->dm, (int)nd, st);
    launch_mesh_scan(mp, m->dm, (int)nd, st);
    if (m->profiling)
        CHS_CUDA(cudaEventRecord(m->evt[5], st));
    CHS_CUDA(cudaGetLastError());
    CHS_CUDA(cudaMemcpyAsync(&m->hCtr[chs_map::kRing], m->dCtr, sizeof(Counters), cudaMemcpyDeviceToHost, st));
    CHS_CUDA(cudaStreamSynchronize(st));
    const Counters c = m->hCtr[chs_map::kRing];
    const long long nv = (long long)c.mesh_verts, ng = (long long)c.mesh_grids;
    if (nv > m->vertCap)
    {
        size_t a = (size_t)m->vertCap * 3, b = a, cc = a;
        const size_t want = (size_t)(nv + nv / 4) * 3;
        if ((rc = grow_buffer(&m->dVerts, &a, want, st)) || (rc = grow_buffer(&m->dNormals, &b, want, st)))
            return rc;
        if (m->cfg.use_color && (rc = grow_buffer(&m->dColors, &cc, want, st)))
            return rc;
        m->vertCap = (long long)(want / 3);
    }
    if (ng > m->gridCap)
    {
        size_t a = (size_t)m->gridCap * 3;
        const size_t want = (size_t)(ng + ng / 4) * 3;
        if ((rc = grow_buffer(&m->dGrids, &a, want, st)))
            return rc;
        m->gridCap = (long long)(want / 3);
    }
    mp.vertices = m->dVerts;
    mp.normals = m->dNormals;
    mp.colors = m->cfg.use_color ? m->dColors : nullptr;
    mp.grids = m->dGrids;
    mp.cap_vertices = m->vertCap;
    mp.cap_grids = m->gridCap;
    if (c.mesh_chunks > 0)
        launch_mesh_emit(mp, m->dm, c.mesh_chunks, st);
    if (m->profiling)
    {
        CHS_CUDA(cudaEventRecord(m->evt[6], st));
        m->meshTimed = true;
    }
    // meshesToUpdate.clear() (Chisel.cpp:57)
    launch_fill_u64(m->dm.dirty_keys, m->dirtySize, kEmptyKey, st);
    CHS_CUDA(cudaMemsetAsync(&m->dCtr->n_dirty, 0, sizeof(int), st));
    CHS_CUDA(cudaGetLastError());
    m->knownDirty = 0;
    m->lastMesh.n_chunks = c.mesh_chunks;
    m->lastMesh.n_vertices = nv;
    m->lastMesh.n_grids = ng;
    m->lastMeshChunks = c.mesh_chunks;
    return CHS_OK;
}

int chs_mesh_counts_last(chs_map *m, chs_mesh_counts *out)
{
    if (!m || !out)
        return fail(CHS_ERR_INVALID, "null argument");
    *out = m->lastMesh;
    return CHS_OK;
}

int chs_download_meshes(chs_map *m, int32_t *ids, int64_t *vertOffsets, int64_t *gridOffsets, float *vertices, float *normals,
                        float *colors, float *grids)
{
    if (!m)
        return fail(CHS_ERR_INVALID, "null map");
    CHS_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = m->stream;
    if (m->meshGathered)
    {
        // root of a distributed re-mesh: the union of all ranks' meshes (capi_comm.inc)
        const chs_map::Gathered &G = m->gathered;
        if (ids && G.nChunks)
            CHS_CUDA(cudaMemcpyAsync(ids, G.ids, sizeof(int) * 3 * (size_t)G.nChunks, cudaMemcpyDeviceToHost, st));
        if (vertOffsets)
            CHS_CUDA(cudaMemcpyAsync(vertOffsets, G.vertOffsets, sizeof(int64_t) * (size_t)(G.nChunks + 1), cudaMemcpyDeviceToHost, st));
        if (gridOffsets)
            CHS_CUDA(cudaMemcpyAsync(gridOffsets, G.gridOffsets, sizeof(int64_t) * (size_t)(G.nChunks + 1), cudaMemcpyDeviceToHost, st));
        if (G.nVerts)
        {
            if (vertices)
                CHS_CUDA(cudaMemcpyAsync(vertices, G.verts, sizeof(float) * 3 * (size_t)G.nVerts, cudaMemcpyDeviceToHost, st));
            if (normals)
                CHS_CUDA(cudaMemcpyAsync(normals, G.normals, sizeof(float) * 3 * (size_t)G.nVerts, cudaMemcpyDeviceToHost, st));
            if (colors && m->cfg.use_color)
                CHS_CUDA(cudaMemcpyAsync(colors, G.colors, sizeof(float) * 3 * (size_t)G.nVerts, cudaMemcpyDeviceToHost, st));
        }
        if (G.nGrids && grids)
            CHS_CUDA(cudaMemcpyAsync(grids, G.grids, sizeof(float) * 3 * (size_t)G.nGrids, cudaMemcpyDeviceToHost, st));
        CHS_CUDA(cudaStreamSynchronize(st));
        return CHS_OK;
    }
    if (m->meshFromGhost && m->ghost)
        return chs_download_meshes(m->ghost, ids, vertOffsets, gridOffsets, vertices, normals, colors, grids);   // non-root: this rank's part
    const int n = m->lastMeshChunks;
    const long long nv = m->lastMesh.n_vertices, ng = m->lastMesh.n_grids;
    std::vector<int> slots((size_t)n);
    if (n)
    {
        CHS_CUDA(cudaMemcpyAsync(slots.data(), m->dMeshSlots, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st));
        static_assert(sizeof(long long) == sizeof(int64_t), "offset type");
        if (vertOffsets)
            CHS_CUDA(cudaMemcpyAsync(vertOffsets, m->dVertOffsets, sizeof(int64_t) * (size_t)(n + 1), cudaMemcpyDeviceToHost, st));
        if (gridOffsets)
            CHS_CUDA(cudaMemcpyAsync(gridOffsets, m->dGridOffsets, sizeof(int64_t) * (size_t)(n + 1), cudaMemcpyDeviceToHost, st));
    }
    else
    {
        if (vertOffsets)
            vertOffsets[0] = 0;
        if (gridOffsets)
            gridOffsets[0] = 0;
    }
    if (nv)
    {
        if (vertices)
            CHS_CUDA(cudaMemcpyAsync(vertices, m->dVerts, sizeof(float) * 3 * (size_t)nv, cudaMemcpyDeviceToHost, st));
        if (normals)
            CHS_CUDA(cudaMemcpyAsync(normals, m->dNormals, sizeof(float) * 3 * (size_t)nv, cudaMemcpyDeviceToHost, st));
        if (colors && m->cfg.use_color)
            CHS_CUDA(cudaMemcpyAsync(colors, m->dColors, sizeof(float) * 3 * (size_t)nv, cudaMemcpyDeviceToHost, st));
    }
    if (ng && grids)
        CHS_CUDA(cudaMemcpyAsync(grids, m->dGrids, sizeof(float) * 3 * (size_t)ng, cudaMemcpyDeviceToHost, st));
    CHS_CUDA(cudaStreamSynchronize(st));
    if (ids && n)
    {
        int rc = refresh_host_ids(m, m->knownChunks);
        if (rc)
            return rc;
        for (int i = 0; i < n; i++)
            for (int k = 0; k < 3; k++)
                ids[3 * i + k] = m->hostIds[3 * (size_t)slots[i] + k];
    }
    return CHS_OK;
}

int chs_frustum(const float pose[12], const chs_camera *cam, float corners[24], float lines[72], float planes[24])
{
    if (!pose || !cam)
        return fail(CHS_ERR_INVALID, "null argument");
    FrustumGeom g;
    build_frustum(pose, *cam, &g);
    if (corners)
        for (int i = 0; i < 8; i++)
            for (int k = 0; k < 3; k++)
                corners[3 * i + k] = g.corner[i][k];
    if (lines)
        frustum_lines(g, lines);
    if (planes)
        for (int p = 0; p < 6; p++)
        {
            for (int k = 0; k < 3; k++)
                planes[4 * p + k] = g.plane[p].n[k];
            planes[4 * p + 3] = g.plane[p].d;
        }
    return CHS_OK;
}

// ChunkManager::GetChunkIDsIntersecting(Frustum) on the host (OC ChunkManager.cpp:182-212), for the facade and tests.
int chs_candidate_ids(int chunk_size, float resolution, const float pose[12], const chs_camera *cam, int32_t *ids, int64_t cap, int64_t *n)
{
    if (!pose || !cam || !n)
        return fail(CHS_ERR_INVALID, "null argument");
    FrustumGeom g;
    build_frustum(pose, *cam, &g);
    CandidateBox box;
    if (!candidate_box(g, chunk_size, resolution, &box))
        return fail(CHS_ERR_INVALID, "frustum is not finite or out of range");
    int64_t count = 0;
    const float ext = (float)chunk_size * resolution;
    for (int x = box.lo[0]; x <= box.hi[0]; x++)
        for (int y = box.lo[1]; y <= box.hi[1]; y++)
            for (int z = box.lo[2]; z <= box.hi[2]; z++)
            {
                const F3 bmin = f3((float)(x * chunk_size) * resolution, (float)(y * chunk_size) * resolution, (float)(z * chunk_size) * resolution);
                const F3 bmax = f3(bmin[0] + ext, bmin[1] + ext, bmin[2] + ext);
                bool hit = false;
                for (int p = 0; p < 6 && !hit; p++)
                {
                    const PlaneEq &pl = g.plane[p];
                    const F3 a = f3(pl.n[0] < 0.0f ? bmin[0] : bmax[0], pl.n[1] < 0.0f ? bmin[1] : bmax[1], pl.n[2] < 0.0f ? bmin[2] : bmax[2]);
                    hit = dot3(a, pl.n) + pl.d > 0.0f;
                }
                if (hit)
                {
                    if (ids && count < cap)
                    {
                        ids[3 * count] = x;
                        ids[3 * count + 1] = y;
                        ids[3 * count + 2] = z;
                    }
                    count++;
                }
            }
    *n = count;
    return CHS_OK;
}

void *chs_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess)
    {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

void chs_host_free(void *p)
{
    if (p)
        cudaFreeHost(p);
}

// ---- frame ingestion of the collaborative server (SPG/src/collaborative_server_system.cpp:213-276): cv::resize + validity mask ----
// cv::resize, INTER_LINEAR, float images: per destination column x = (dx + 0.5) * scale - 0.5 with scale = 1 / (double(dst) / src),
// sx = floor(x), fx = x - sx, clamped at the borders (weight 0 on the missing neighbour); rows alike. Horizontal pass first
// (S[sx] * (1 - fx) + S[sx + 1] * fx), then vertical. OpenCV's own SIMD and scalar paths differ in the last bit (fused vs. separate
// multiply-add), so this agrees with cv2 to 1 ulp, not bit for bit (tests/test_parity_gpu.py::test_ingest_depth_resize_and_mask).
// Then NaN outside [vmin, vmax] (exact).
__global__ void ingest_depth_kernel(const float *src, int sw, int sh, float *dst, int dw, int dh, double scaleX, double scaleY, float vmin, float vmax)
{
    const float nanv = __int_as_float(0x7fc00000);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < dw * dh; i += gridDim.x * blockDim.x)
    {
        const int dx = i % dw, dy = i / dw;
        float v;
        if (sw == dw && sh == dh)
            v = src[i];
        else
        {
            // source position and its fraction in double (OpenCV 4 keeps the fraction's precision: rounding the position to float
            // first costs up to 1e-5 of relative weight at x ~ 100), the weights themselves in float
            const double px = (dx + 0.5) * scaleX - 0.5, py = (dy + 0.5) * scaleY - 0.5;
            int sx = (int)floor(px), sy = (int)floor(py);
            float fx = (float)(px - (double)sx), fy = (float)(py - (double)sy);
            if (sx < 0) { sx = 0; fx = 0.0f; }
            if (sx >= sw - 1) { sx = sw - 1; fx = 0.0f; }
            if (sy < 0) { sy = 0; fy = 0.0f; }
            if (sy >= sh - 1) { sy = sh - 1; fy = 0.0f; }
            const int sx1 = min(sx + 1, sw - 1), sy1 = min(sy + 1, sh - 1);
            const float a0 = 1.0f - fx, a1 = fx, b0 = 1.0f - fy, b1 = fy;
            const float r0 = __fadd_rn(__fmul_rn(src[(size_t)sy * sw + sx], a0), __fmul_rn(src[(size_t)sy * sw + sx1], a1));
            const float r1 = __fadd_rn(__fmul_rn(src[(size_t)sy1 * sw + sx], a0), __fmul_rn(src[(size_t)sy1 * sw + sx1], a1));
            v = __fmaf_rn(r0, b0, __fmul_rn(r1, b1));
        }
        dst[i] = (v < vmin || v > vmax) ? nanv : v;            // a NaN input stays NaN (both comparisons are false, the value is kept)
    }
}

int chs_ingest_depth(chs_map *m, const float *src, int src_w, int src_h, float *dst, int dst_w, int dst_h, float valid_min, float valid_max, int mem)
{
    if (!m || !src || !dst || src_w <= 0 || src_h <= 0 || dst_w <= 0 || dst_h <= 0)
        return fail(CHS_ERR_INVALID, "bad argument");
    if (mem != CHS_MEM_HOST && mem != CHS_MEM_DEVICE)
        return fail(CHS_ERR_INVALID, "chs_ingest_depth takes CHS_MEM_HOST or CHS_MEM_DEVICE");
    CHS_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = m->stream;
    const double scaleX = 1.0 / ((double)dst_w / src_w), scaleY = 1.0 / ((double)dst_h / src_h);
    const size_t ns = (size_t)src_w * src_h, nd = (size_t)dst_w * dst_h;
    const float *dSrc = src;
    float *dDst = dst;
    float *tmpS = nullptr, *tmpD = nullptr;
    if (mem == CHS_MEM_HOST)
    {
        CHS_CUDA(cudaMallocAsync((void **)&tmpS, ns * sizeof(float), st));
        CHS_CUDA(cudaMallocAsync((void **)&tmpD, nd * sizeof(float), st));
        CHS_CUDA(cudaMemcpyAsync(tmpS, src, ns * sizeof(float), cudaMemcpyHostToDevice, st));
        dSrc = tmpS;
        dDst = tmpD;
    }
    ingest_depth_kernel<<<(int)std::min<size_t>((nd + 255) / 256, 148 * 8), 256, 0, st>>>(dSrc, src_w, src_h, dDst, dst_w, dst_h, scaleX, scaleY, valid_min, valid_max);
    CHS_CUDA(cudaGetLastError());
    if (mem == CHS_MEM_HOST)
    {
        CHS_CUDA(cudaMemcpyAsync(dst, tmpD, nd * sizeof(float), cudaMemcpyDeviceToHost, st));
        CHS_CUDA(cudaFreeAsync(tmpS, st));
        CHS_CUDA(cudaFreeAsync(tmpD, st));
        CHS_CUDA(cudaStreamSynchronize(st));
    }
    return CHS_OK;
}

void *chs_device_alloc(chs_map *m, size_t bytes)
{
    if (!m || cudaSetDevice(m->device) != cudaSuccess)
        return nullptr;
    void *p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess)
    {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

void chs_device_free(chs_map *m, void *p)
{
    if (m && p && cudaSetDevice(m->device) == cudaSuccess)
    {
        cudaStreamSynchronize(m->stream);
        cudaStreamSynchronize(m->copyStream);
        cudaStreamSynchronize(m->pushStream);
        cudaFree(p);
    }
}

int chs_upload(chs_map *m, void *dst, const void *src, size_t bytes)
{
    if (!m || !dst || !src)
        return fail(CHS_ERR_INVALID, "null argument");
    CHS_CUDA(cudaSetDevice(m->device));
    if (!m->uploadStream)
        CHS_CUDA(cudaStreamCreateWithFlags(&m->uploadStream, cudaStreamNonBlocking));
    CHS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, m->uploadStream));
    CHS_CUDA(cudaStreamSynchronize(m->uploadStream));       // the caller reuses its buffer right away (CR ChiselServer.cpp:285-295)
    return CHS_OK;
}

int chs_selftest_arithmetic(int64_t div_pairs, int64_t out[4])
{
    if (!out || div_pairs < 0)
        return fail(CHS_ERR_INVALID, "null argument");
    unsigned long long *d = nullptr;
    CHS_CUDA(cudaMalloc((void **)&d, 4 * sizeof(unsigned long long)));
    CHS_CUDA(cudaMemset(d, 0, 4 * sizeof(unsigned long long)));
    CHS_CUDA(half::launch_selftest_arithmetic(d, (unsigned long long)div_pairs, nullptr));
    unsigned long long h[4];
    CHS_CUDA(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(d);
    for (int i = 0; i < 4; i++)
        out[i] = (int64_t)h[i];
    return CHS_OK;
}

float chs_truncation(int kind, float param, float depth) { return host_truncation(kind, param, depth); }
uint32_t chs_owner(int32_t x, int32_t y, int32_t z) { return owner_hash(x, y, z); }

} // extern "C"

#include "capi_comm.inc"
