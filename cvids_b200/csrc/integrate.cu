// integrate.cu -- per-frame kernels of the TSDF integration path (sm_100a).
//
//   frame_prepare_kernel     per-pixel truncation + {min near, max far} depth-range tiles ("Hi-Z") for culling
//   chunk_candidates_kernel  enumerate the reference's candidate ID box, reproduce Frustum::Intersects, drop
//                            chunks / 8^3 bricks that provably cannot change, warp-ballot compact the rest into two
//                            work lists (new-chunk candidates, brick units of existing chunks)
//   integrate_new_chunks_kernel  exact band test of the new-chunk candidates; survivors are allocated and written once
//   integrate_bricks_kernel      projective SDF / weight / colour update of existing chunks, a warp per half brick
//
// Exact-arithmetic rule: everything that decides a branch or produces stored state follows SURVEY.md
// Appendix A operation by operation with __f*_rn intrinsics (never contracted into FMA; the file is also
// built with -fmad=false). Culling code is free-form float math with explicit slack, and is conservative:
// it may keep a chunk that turns out to be untouched, never drop one that would be touched.
#include <algorithm>
#include <cstddef>
#include <cstdlib>

#include "device_map.cuh"
#include "integrate_device.cuh"
#include "kernels.h"

namespace chs
{


// ColorImage::At once per pixel (color_pack_body). Runs as a graph branch parallel to frame_prepare -> chunk_candidates
// (only the integrate kernels consume its output).
__global__ void __launch_bounds__(256) color_pack_kernel(FrameParams fp)
{
    color_pack_body(fp, (blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x, gridDim.x * gridDim.y * blockDim.x);
}

__global__ void __launch_bounds__(256) frame_prepare_kernel(FrameParams fp, DeviceMap map)
{
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < kCounterSlots)
    {
        Counters *c = map.ctr;
        if (threadIdx.x == 0)
        {
            c->unit_count = 0;
            c->new_count = 0;
            c->candidates = 0;
            c->n_new = 0;
            c->updated_chunks = 0;
            c->tickets = 0;
        }
        c->n_upd[threadIdx.x] = 0;
        c->n_carve[threadIdx.x] = 0;
        c->n_col[threadIdx.x] = 0;
    }
    frame_prepare_tile(fp, blockIdx.x, blockIdx.y);
}

// One lane per (candidate chunk, 8^3 brick): groups of GL lanes share a chunk ID of the candidate box (the reference's
// order -- x outer, y, z inner; ChunkManager.cpp:192-196 -- does not affect the result, so IDs are enumerated by linear
// index). Every brick is classified against the Hi-Z tiles directly; the hash table is consulted (by the group leader)
// only for chunks with a brick the depth test could not reject. Survivors go, by warp-ballot compaction, to
//   news   non-existing chunks with a brick that may receive a band hit   -> integrate_new_chunks_kernel (CTA per chunk)
//   units  bricks of existing chunks that may change                       -> integrate_bricks_kernel (warp per half brick)
template <int CS>
__global__ void __launch_bounds__(256) chunk_candidates_kernel(FrameParams fp, DeviceMap map)
{
    constexpr int BPA = CS / 8, NB = BPA * BPA * BPA;
    constexpr int GL = NB >= 32 ? 32 : NB;                          // lanes per chunk: 1 (8^3), 8 (16^3), 32 (32^3)
    constexpr int BPL = NB / GL;                                     // bricks per lane: 2 for 32^3 chunks, else 1
    const int total = fp.n[0] * fp.n[1] * fp.n[2];
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int i = (int)(tid / GL);
    const int gl = (int)(tid % GL);
    const unsigned lane = threadIdx.x & 31;
    const unsigned groupMask = GL == 32 ? 0xffffffffu : (((1u << GL) - 1u) << (lane & ~(unsigned)(GL - 1)));
    const bool leader = gl == 0;
    bool candidate = false;
    int x = 0, y = 0, z = 0;
    float bx = 0.0f, by = 0.0f, bz = 0.0f;
    int code[BPL];
#pragma unroll
    for (int k = 0; k < BPL; k++)
        code[k] = 0;
    if (i < total)
    {
        const int nyz = fp.n[1] * fp.n[2];
        x = fp.lo[0] + i / nyz;
        const int r = i - (i / nyz) * nyz;
        y = fp.lo[1] + r / fp.n[2];
        z = fp.lo[2] + r % fp.n[2];
        // chunk box exactly as ChunkManager.cpp:199-201
        const float ext = __fmul_rn((float)CS, map.res);
        bx = __fmul_rn((float)(x * CS), map.res);
        by = __fmul_rn((float)(y * CS), map.res);
        bz = __fmul_rn((float)(z * CS), map.res);
        candidate = frustum_intersects_exact(fp, bx, by, bz, __fadd_rn(bx, ext), __fadd_rn(by, ext), __fadd_rn(bz, ext));
        if (candidate && map.world > 1)
            candidate = (owner_hash(x, y, z) % (unsigned)map.world) == (unsigned)map.rank;
    }
    // one lane per brick: a single classification stage keeps the dependent-latency chain short (a chunk-level pre-test
    // saved instructions but added a dependent stage and measured slower)
    if (candidate)
    {
#pragma unroll
        for (int k = 0; k < BPL; k++)
        {
            const int b = gl + k * GL;
            const int qx = b % BPA, qy = (b / BPA) % BPA, qz = b / (BPA * BPA);
            code[k] = (NB == 1) ? classify_box(fp, bx + map.half, by + map.half, bz + map.half, (float)(CS - 1) * map.res)
                                : classify_box(fp, bx + (float)(qx * 8) * map.res + map.half, by + (float)(qy * 8) * map.res + map.half,
                                               bz + (float)(qz * 8) * map.res + map.half, 7.0f * map.res);
        }
    }
    bool has2 = false, has1 = false;
#pragma unroll
    for (int k = 0; k < BPL; k++)
    {
        has2 |= code[k] == 2;
        has1 |= code[k] == 1;
    }
    const unsigned m2 = __ballot_sync(0xffffffffu, has2) & groupMask;
    const unsigned m1 = __ballot_sync(0xffffffffu, has1) & groupMask;
    const bool chunkBand = m2 != 0u, chunkFree = m1 != 0u;
    // group leader: one hash lookup per undecided chunk, one flag word per existing one
    int slot = -1;
    unsigned long long flags = 0ull;
    if (leader && (chunkBand || (chunkFree && fp.carve)))
    {
        slot = hash_lookup(map, pack_id(x, y, z));
        if (slot >= 0 && chunkFree && fp.carve)
            flags = map.brick_flags[slot];
    }
    const int leaderLane = (int)(lane & ~(unsigned)(GL - 1));
    slot = __shfl_sync(0xffffffffu, slot, leaderLane);
    flags = __shfl_sync(0xffffffffu, flags, leaderLane);

    const bool keepNew = leader && slot < 0 && chunkBand;
    const unsigned candMask = __ballot_sync(0xffffffffu, candidate && leader);
    const unsigned newMask = __ballot_sync(0xffffffffu, keepNew);
    int base = 0;
    if (lane == 0)
    {
        if (candMask)
            atomicAdd(&map.ctr->candidates, __popc(candMask));
        if (newMask)
            base = atomicAdd(&map.ctr->new_count, __popc(newMask));
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keepNew)
    {
        const int pos = base + __popc(newMask & ((1u << lane) - 1));
        if (pos < fp.news_cap)
            fp.news[pos] = make_int4(x, y, z, -1);
        else
            atomicOr(&map.ctr->error_flags, kErrWorkFull);
    }
#pragma unroll
    for (int k = 0; k < BPL; k++)
    {
        const int b = gl + k * GL;
        const bool keepB = slot >= 0 && (code[k] == 2 || (code[k] == 1 && fp.carve && ((flags >> b) & 1ull)));
        const unsigned bm = __ballot_sync(0xffffffffu, keepB);
        int ubase = 0;
        if (lane == 0 && bm)
            ubase = atomicAdd(&map.ctr->unit_count, __popc(bm));
        ubase = __shfl_sync(0xffffffffu, ubase, 0);
        if (keepB)
        {
            const int pos = ubase + __popc(bm & ((1u << lane) - 1));
            if (pos < fp.units_cap)
                fp.units[pos] = make_int4(x, y, z, slot | (b << 24));
            else
                atomicOr(&map.ctr->error_flags, kErrWorkFull);
        }
    }
}


// Chunk `slot` changed this frame: the FIRST warp to say so marks all 27 neighbour IDs dirty, whether or not they exist
// (Chisel.h:89-101, 175-189), and counts the chunk. slot_epoch[slot] holds the id of the last frame that updated it.
__device__ __forceinline__ void mark_chunk_updated(const FrameParams &fp, const DeviceMap &map, int slot, int x, int y, int z, int lane)
{
    int first = 0;
    if (lane == 0)
        first = atomicExch(&map.slot_epoch[slot], fp.frame_id) != fp.frame_id;
    first = __shfl_sync(0xffffffffu, first, 0);
    if (first)
    {
        if (lane < 27)
            dirty_insert(map, pack_id(x + lane / 9 - 1, y + (lane / 3) % 3 - 1, z + lane % 3 - 1));
        if (lane == 31)
            atomicAdd(&map.ctr->updated_chunks, 1);
    }
}

// Block-level flush of the per-lane counters (one set of global atomics per CTA, spread over kCounterSlots addresses),
// then a ticket: the LAST CTA of the frame's two integrate kernels copies what the host needs into the pinned ring slot
// (no snapshot kernel, no memcpy, no event: the host polls frame_id).
__device__ __forceinline__ void flush_counters(const FrameParams &fp, const DeviceMap &map, const VoxelStats &st, int *sCnt)
{
    int a = st.nUpd, b = st.nCarve, c = st.nCol;
    for (int o = 16; o > 0; o >>= 1)
    {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if ((threadIdx.x & 31) == 0)
    {
        if (a) atomicAdd(&sCnt[0], a);
        if (b) atomicAdd(&sCnt[1], b);
        if (c) atomicAdd(&sCnt[2], c);
    }
    __syncthreads();
    if (threadIdx.x < 3 && sCnt[threadIdx.x])
    {
        unsigned long long *dst = threadIdx.x == 0 ? map.ctr->n_upd : (threadIdx.x == 1 ? map.ctr->n_carve : map.ctr->n_col);
        atomicAdd(&dst[blockIdx.x % kCounterSlots], (unsigned long long)sCnt[threadIdx.x]);
    }
    __syncthreads();
    if (threadIdx.x < 32)
    {
        int last = 0;
        if (threadIdx.x == 0)
        {
            __threadfence();
            last = atomicAdd(&map.ctr->tickets, 1) == fp.total_ctas - 1;
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last)
        {
            __threadfence();
            const volatile Counters *ctr = map.ctr;
            static_assert(kCounterSlots == 32, "one lane per counter slot");
            long long u = (long long)ctr->n_upd[threadIdx.x], v = (long long)ctr->n_carve[threadIdx.x], w = (long long)ctr->n_col[threadIdx.x];
            for (int o = 16; o > 0; o >>= 1)
            {
                u += __shfl_xor_sync(0xffffffffu, u, o);
                v += __shfl_xor_sync(0xffffffffu, v, o);
                w += __shfl_xor_sync(0xffffffffu, w, o);
            }
            if (threadIdx.x == 0)
            {
                int4 *h = reinterpret_cast<int4 *>(fp.host_slot);
                const int id = fp.frame_id;
                h[0] = make_int4(id, ctr->n_chunks, ctr->n_dirty, ctr->error_flags);
                h[1] = make_int4(id, ctr->unit_count, ctr->new_count, ctr->candidates);
                h[2] = make_int4(id, ctr->n_new, ctr->updated_chunks, (int)v);
                h[3] = make_int4(id, (int)w, (int)(u & 0xffffffffll), (int)(u >> 32));
            }
        }
    }
}

// Existing chunks: one warp per half brick (8 x 8 x 4 voxels = two batches); no block-level synchronisation in the loop.
// The grid is sized to what is resident at once and strides over the unit list, whose length only the device knows.
template <int CS, bool COLOR_PATH, bool PER_PIXEL>
__global__ void __launch_bounds__(256, 3) integrate_bricks_kernel(FrameParams fp, DeviceMap map)
{
    constexpr int BPA = CS / 8;
    __shared__ int sCnt[3];
    if (threadIdx.x < 3)
        sCnt[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int nTasks = min(map.ctr->unit_count, fp.units_cap) * 2;
    VoxelStats st;
    st.nUpd = st.nCarve = st.nCol = 0;
    for (int g = blockIdx.x * 8 + (threadIdx.x >> 5); g < nTasks; g += gridDim.x * 8)
    {
        const int4 unit = fp.units[g >> 1];
        const int slot = unit.w & 0xFFFFFF, b = unit.w >> 24;
        // origin_k = float(CS * ID_k) * res (Chunk.cpp:43)
        const float orgx = __fmul_rn((float)(CS * unit.x), map.res), orgy = __fmul_rn((float)(CS * unit.y), map.res), orgz = __fmul_rn((float)(CS * unit.z), map.res);
        const BrickLane L = brick_lane_setup(fp, map, orgx, orgy, orgz, b % BPA, (b / BPA) % BPA, b / (BPA * BPA), lane);
        float2 *dist = dist_ptr(map, slot);
        unsigned *col = map.use_color ? reinterpret_cast<unsigned *>(color_ptr(map, slot)) : nullptr;
        st.updated = st.carvable = false;
        process_batch<CS, COLOR_PATH, PER_PIXEL, 0>(fp, map, L, (g & 1) * 2, dist, col, &st);
        process_batch<CS, COLOR_PATH, PER_PIXEL, 0>(fp, map, L, (g & 1) * 2 + 1, dist, col, &st);
        if (__any_sync(0xffffffffu, st.carvable) && lane == 0)
            atomicOr(&map.brick_flags[slot], 1ull << b);
        if (__any_sync(0xffffffffu, st.updated))
            mark_chunk_updated(fp, map, slot, unit.x, unit.y, unit.z, lane);
    }
    flush_counters(fp, map, st, sCnt);
}

// New chunks: one CTA per candidate, a warp per brick (CS = 32: eight bricks per warp; CS = 8: warp 0 only). The chunk
// is materialised only if the exact test finds a band hit: "created and untouched => garbage collected"
// (Chisel.h:76-80,102-110 / :133-143,170-173,202-207) never allocates anything.
template <int CS, bool COLOR_PATH, bool PER_PIXEL>
__global__ void __launch_bounds__(256) integrate_new_chunks_kernel(FrameParams fp, DeviceMap map)
{
    constexpr int BPA = CS / 8, NB = BPA * BPA * BPA;
    __shared__ int sCnt[3];
    __shared__ int sSlotMem;
    int *sSlot = &sSlotMem;
    const int nBlocks = gridDim.x;
    if (threadIdx.x < 3)
        sCnt[threadIdx.x] = 0;
    __syncthreads();
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nWarps = blockDim.x >> 5;
    const int nNew = min(map.ctr->new_count, fp.news_cap);
    VoxelStats st;
    st.nUpd = st.nCarve = st.nCol = 0;
    for (int w = blockIdx.x; w < nNew; w += nBlocks)
    {
        const int4 item = fp.news[w];
        const float orgx = __fmul_rn((float)(CS * item.x), map.res), orgy = __fmul_rn((float)(CS * item.y), map.res), orgz = __fmul_rn((float)(CS * item.z), map.res);
        // Would ProjectionIntegrator::Integrate[Color] report an update? (carving cannot touch a fresh chunk: weight 0)
        bool any = false;
        for (int b = warp; b < NB && !any; b += nWarps)
        {
            const int bx = b % BPA, by = (b / BPA) % BPA, bz = b / (BPA * BPA);
            if (NB > 1 && classify_box(fp, orgx + (float)(bx * 8) * map.res + map.half, orgy + (float)(by * 8) * map.res + map.half,
                                       orgz + (float)(bz * 8) * map.res + map.half, 7.0f * map.res) != 2)
                continue;
            const BrickLane L = brick_lane_setup(fp, map, orgx, orgy, orgz, bx, by, bz, lane);
            // all four batches in one basic block: the 16 depth gathers of the lane are in flight together
            bool hit = false;
#pragma unroll
            for (int q = 0; q < 4; q++)
                hit |= process_batch<CS, COLOR_PATH, PER_PIXEL, 2>(fp, map, L, q, nullptr, nullptr, &st);
            any = __any_sync(0xffffffffu, hit);
        }
        if (!__syncthreads_or(any))
            continue;
        // The chunk survives, hence it is updated: mark its 27-neighbourhood dirty (Chisel.h:89-101, 175-189) on warp 1
        // while thread 0 allocates -- the two chains of global atomics overlap.
        if (warp == (nWarps > 1 ? 1 : 0))
        {
            if (lane < 27)
                dirty_insert(map, pack_id(item.x + lane / 9 - 1, item.y + (lane / 3) % 3 - 1, item.z + lane % 3 - 1));
            if (lane == 31)
                atomicAdd(&map.ctr->updated_chunks, 1);
        }
        if (t == 0)
        {
            int s = atomicAdd(&map.ctr->n_chunks, 1);
            if (s >= map.capacity)
            {
                atomicOr(&map.ctr->error_flags, kErrPoolFull);
                atomicSub(&map.ctr->n_chunks, 1);
                s = -1;
            }
            else
            {
                map.slot_ids[3 * s] = item.x;
                map.slot_ids[3 * s + 1] = item.y;
                map.slot_ids[3 * s + 2] = item.z;
                map.brick_flags[s] = 0ull;
                map.slot_epoch[s] = fp.frame_id;
                hash_insert_new(map, pack_id(item.x, item.y, item.z), s);
                atomicAdd(&map.ctr->n_new, 1);
            }
            *sSlot = s;
        }
        __syncthreads();
        const int slot = *sSlot;
        if (slot < 0)
        {
            __syncthreads();
            continue;
        }
        float2 *dist = dist_ptr(map, slot);
        unsigned *col = map.use_color ? reinterpret_cast<unsigned *>(color_ptr(map, slot)) : nullptr;
        for (int b = warp; b < NB; b += nWarps)
        {
            const BrickLane L = brick_lane_setup(fp, map, orgx, orgy, orgz, b % BPA, (b / BPA) % BPA, b / (BPA * BPA), lane);
            st.updated = st.carvable = false;
#pragma unroll 1
            for (int q = 0; q < 4; q++)
                process_batch<CS, COLOR_PATH, PER_PIXEL, 1>(fp, map, L, q, dist, col, &st);
            // brick_flags[slot] was zeroed by thread 0 before the barrier above
            if (__any_sync(0xffffffffu, st.carvable) && lane == 0)
                atomicOr(&map.brick_flags[slot], 1ull << b);
        }
        __syncthreads();                                                // sSlot is reused by the next item
    }
    flush_counters(fp, map, st, sCnt);
}

// ------------------------------------------------------------------------------------------------------
// host launchers

template <typename K>
static int resident_blocks(K kernel, int threads)
{
    int dev = 0, sms = 148, perSm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, threads, 0);
    return sms * (perSm > 0 ? perSm : 1);
}

// ------------------------------------------------------------------------------------------------------
// The per-frame work as a CUDA graph:
//     frame_prepare -> chunk_candidates -> { integrate_new_chunks || integrate_bricks }   (the last CTA snapshots the counters)
// built once per kernel variant; every frame only the kernel arguments and grid sizes are patched
// (cudaGraphExecKernelNodeSetParams) and the graph is launched once. The profiling variant serialises the two integrate
// kernels and brackets every kernel with event-record nodes.
struct FrameGraphVariant
{
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    cudaGraphNode_t nPrepare = nullptr, nCand = nullptr, nNew = nullptr, nBricks = nullptr, nPack = nullptr;
    void *fPrepare = nullptr, *fCand = nullptr, *fNew = nullptr, *fBricks = nullptr;
    int newResident = 1, brickResident = 1;
};

struct FrameGraph
{
    FrameGraphVariant v[8];   // index = color_path | per_pixel << 1 | profiling << 2
};

template <int CS, bool COLOR_PATH, bool PER_PIXEL>
static void variant_functions(FrameGraphVariant *g)
{
    g->fPrepare = (void *)frame_prepare_kernel;
    g->fCand = (void *)chunk_candidates_kernel<CS>;
    g->fNew = (void *)integrate_new_chunks_kernel<CS, COLOR_PATH, PER_PIXEL>;
    g->fBricks = (void *)integrate_bricks_kernel<CS, COLOR_PATH, PER_PIXEL>;
    g->newResident = resident_blocks(integrate_new_chunks_kernel<CS, COLOR_PATH, PER_PIXEL>, 256);
    g->brickResident = resident_blocks(integrate_bricks_kernel<CS, COLOR_PATH, PER_PIXEL>, 256);
}

template <int CS>
static void variant_functions_cs(FrameGraphVariant *g, bool color, bool pp)
{
    if (color)
    {
        if (pp)
            variant_functions<CS, true, true>(g);
        else
            variant_functions<CS, true, false>(g);
    }
    else
    {
        if (pp)
            variant_functions<CS, false, true>(g);
        else
            variant_functions<CS, false, false>(g);
    }
}

FrameGraph *frame_graph_create() { return new FrameGraph(); }

void frame_graph_destroy(FrameGraph *fg)
{
    if (!fg)
        return;
    for (FrameGraphVariant &g : fg->v)
    {
        if (g.exec)
            cudaGraphExecDestroy(g.exec);
        if (g.graph)
            cudaGraphDestroy(g.graph);
    }
    delete fg;
}

static cudaKernelNodeParams kernel_params(void *func, dim3 grid, dim3 block, void **args)
{
    cudaKernelNodeParams p{};
    p.func = func;
    p.gridDim = grid;
    p.blockDim = block;
    p.sharedMemBytes = 0;
    p.kernelParams = args;
    p.extra = nullptr;
    return p;
}

// events: [0] before prepare, [1] after prepare, [2] after candidates, [7] after new chunks, [3] after bricks (profiling only)
cudaError_t frame_graph_launch(FrameGraph *fg, const FrameParams &fpIn, const DeviceMap &mapIn, long long candidates, long long newHint, HostSnapshot *hostSlot,
                               bool profiling, cudaEvent_t *evt, cudaStream_t st)
{
    FrameParams fp = fpIn;
    DeviceMap map = mapIn;
    fp.host_slot = hostSlot;
    const bool pp = fp.trunc_img != nullptr;
    FrameGraphVariant &g = fg->v[(fp.color_path ? 1 : 0) | (pp ? 2 : 0) | (profiling ? 4 : 0)];
    const int total = fp.n[0] * fp.n[1] * fp.n[2];
    const long long nb = (long long)(map.cs / 8) * (map.cs / 8) * (map.cs / 8);
    void *argsFrame[2] = {&fp, &map};
    void *argsPack[1] = {&fp};
    cudaError_t e;
    const bool build = g.exec == nullptr;
    if (build)
    {
        switch (map.cs)
        {
        case 8: variant_functions_cs<8>(&g, fp.color_path != 0, pp); break;
        case 16: variant_functions_cs<16>(&g, fp.color_path != 0, pp); break;
        default: variant_functions_cs<32>(&g, fp.color_path != 0, pp); break;
        }
    }
    const dim3 gPrepare((fp.cam.W + 63) / 64, (fp.cam.H + 63) / 64);
    const long long candLanes = (long long)total * std::min<long long>(nb, 32);
    const dim3 gCand((unsigned)std::max(1ll, (candLanes + 255) / 256));
    const dim3 gNew((unsigned)std::max(1ll, std::min<long long>(std::min<long long>(candidates, newHint), g.newResident)));
    const dim3 gBricks((unsigned)std::max(1ll, std::min<long long>((candidates * nb * 2 + 7) / 8, g.brickResident)));
    fp.total_ctas = (int)(gNew.x + gBricks.x);
    const int packPixels = fp.color_path ? fp.ccam.W * fp.ccam.H : 0;
    cudaKernelNodeParams pPack = kernel_params((void *)color_pack_kernel, dim3((unsigned)std::max(1, std::min(148 * 4, (packPixels / 4 + 255) / 256))), dim3(256), argsPack);
    cudaKernelNodeParams pPrepare = kernel_params(g.fPrepare, gPrepare, dim3(256), argsFrame);
    cudaKernelNodeParams pCand = kernel_params(g.fCand, gCand, dim3(256), argsFrame);
    cudaKernelNodeParams pNew = kernel_params(g.fNew, gNew, dim3(256), argsFrame);
    cudaKernelNodeParams pBricks = kernel_params(g.fBricks, gBricks, dim3(256), argsFrame);
    if (build)
    {
        if ((e = cudaGraphCreate(&g.graph, 0)) != cudaSuccess)
            return e;
        cudaGraphNode_t prev = nullptr, ev;
        auto addEvent = [&](cudaEvent_t event) -> cudaError_t {
            cudaError_t r = cudaGraphAddEventRecordNode(&ev, g.graph, prev ? &prev : nullptr, prev ? 1 : 0, event);
            prev = ev;
            return r;
        };
        // production: colour packing runs beside the candidates kernel (both wait for prepare; only the integrate kernels
        // consume it). As a second ROOT node it measured 15 us slower per frame, hence not there.
        const bool packBeside = fp.color_path && !profiling;
        if (fp.color_path && !packBeside)
        {
            // profiling variant: first in the serial chain
            if ((e = cudaGraphAddKernelNode(&g.nPack, g.graph, nullptr, 0, &pPack)) != cudaSuccess)
                return e;
            prev = g.nPack;
        }
        if (profiling && (e = addEvent(evt[0])) != cudaSuccess)
            return e;
        if ((e = cudaGraphAddKernelNode(&g.nPrepare, g.graph, prev ? &prev : nullptr, prev ? 1 : 0, &pPrepare)) != cudaSuccess)
            return e;
        prev = g.nPrepare;
        if (profiling && (e = addEvent(evt[1])) != cudaSuccess)
            return e;
        if (packBeside && (e = cudaGraphAddKernelNode(&g.nPack, g.graph, &prev, 1, &pPack)) != cudaSuccess)
            return e;
        if ((e = cudaGraphAddKernelNode(&g.nCand, g.graph, &prev, 1, &pCand)) != cudaSuccess)
            return e;
        prev = g.nCand;
        if (profiling)
        {
            // serial: candidates -> e2 -> new -> e7 -> bricks -> e3 -> snapshot
            if ((e = addEvent(evt[2])) != cudaSuccess)
                return e;
            if ((e = cudaGraphAddKernelNode(&g.nNew, g.graph, &prev, 1, &pNew)) != cudaSuccess)
                return e;
            prev = g.nNew;
            if ((e = addEvent(evt[7])) != cudaSuccess)
                return e;
            if ((e = cudaGraphAddKernelNode(&g.nBricks, g.graph, &prev, 1, &pBricks)) != cudaSuccess)
                return e;
            prev = g.nBricks;
            if ((e = addEvent(evt[3])) != cudaSuccess)
                return e;
        }
        else
        {
            // the two integrate kernels touch disjoint chunks: parallel branches (both wait for the packed colour image)
            cudaGraphNode_t deps[2] = {prev, g.nPack};
            const size_t nDeps = packBeside ? 2 : 1;
            if ((e = cudaGraphAddKernelNode(&g.nNew, g.graph, deps, nDeps, &pNew)) != cudaSuccess)
                return e;
            if ((e = cudaGraphAddKernelNode(&g.nBricks, g.graph, deps, nDeps, &pBricks)) != cudaSuccess)
                return e;
        }
        if ((e = cudaGraphInstantiate(&g.exec, g.graph, 0)) != cudaSuccess)
            return e;
    }
    else
    {
        if (g.nPack && (e = cudaGraphExecKernelNodeSetParams(g.exec, g.nPack, &pPack)) != cudaSuccess)
            return e;
        if ((e = cudaGraphExecKernelNodeSetParams(g.exec, g.nPrepare, &pPrepare)) != cudaSuccess)
            return e;
        if ((e = cudaGraphExecKernelNodeSetParams(g.exec, g.nCand, &pCand)) != cudaSuccess)
            return e;
        if ((e = cudaGraphExecKernelNodeSetParams(g.exec, g.nNew, &pNew)) != cudaSuccess)
            return e;
        if ((e = cudaGraphExecKernelNodeSetParams(g.exec, g.nBricks, &pBricks)) != cudaSuccess)
            return e;
    }
    return cudaGraphLaunch(g.exec, st);
}

float host_truncation(int kind, float param, float depth) { return truncation_of(kind, param, depth); }

} // namespace chs
