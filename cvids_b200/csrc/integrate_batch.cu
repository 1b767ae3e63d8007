// integrate_batch.cu -- fused multi-frame TSDF integration (sm_100a): K <= 16 consecutive frames in ONE pass over the map.
//
// Why this is exact. ProjectionIntegrator::Integrate[Color] (OC ProjectionIntegrator.h:51-183) updates every voxel from
// that voxel's own state and the frame alone; chunk creation / garbage collection (Chisel.h:76-110, :133-207) depends only
// on whether a frame touches the chunk; the dirty set (Chisel.h:89-101) is a set. So for a batch of frames f0 < f1 < ...
//   * a voxel may receive all its updates back to back, in frame order, while its state sits in registers;
//   * a chunk that does not exist at the start of the batch comes into existence at the first frame with a band hit and is an
//     ordinary chunk for the later frames;
//   * the union of the per-frame dirty marks is the dirty set after the last frame.
// The map after the batch is bit-identical to K calls of the single-frame path (tests/test_batch_gpu.py), and the per-frame
// counters (candidates, N_upd, N_carve, N_col, N_new, updated chunks) are kept per frame.
//
// Why it is faster. A 752x480 frame at 2 cm touches ~10 MB of voxel state -- 1.6 us of HBM time, far below the latency of the
// four dependent kernels a frame needs. Fusing K frames amortises the launches and the dependent-latency chains K-fold,
// reads and writes each touched voxel once instead of K times, and gives every warp K frames of independent work.
//
//   batch_prepare_kernel      grid.z = frame: Hi-Z tiles (+ per-pixel truncation) of every frame; zeroes the batch counters
//   batch_color_pack_kernel   grid.y = frame: packed colour images
//   batch_candidates_kernel   thread per (chunk of the UNION candidate box, 8^3 brick): exact Frustum::Intersects and the
//                             conservative depth-range class per frame -> per-brick frame masks; warp-ballot compaction
//   batch_new_chunks_kernel   CTA per chunk that does not exist yet: frames in order, exact band test until the first hit,
//                             allocation, then ordinary integration of the remaining frames
//   batch_bricks_kernel       warp per half brick of an existing chunk: state of 8 voxels per lane in registers, loop over the
//                             brick's frame mask, one store per changed voxel at the end
#include <algorithm>

#include "device_map.cuh"
#include "integrate_device.cuh"
#include "kernels.h"

namespace chs
{

static_assert(sizeof(FrameParams) % 4 == 0, "FrameParams is copied word-wise into shared memory");

// All frames of the batch into shared memory: afterwards `sF[f]` is read with warp-uniform shared loads.
__device__ __forceinline__ void load_frames(FrameParams *sF, const BatchParams &bp)
{
    const int nWords = bp.K * (int)(sizeof(FrameParams) / 4);
    const int *src = reinterpret_cast<const int *>(bp.frames);
    int *dst = reinterpret_cast<int *>(sF);
    for (int i = threadIdx.x; i < nWords; i += blockDim.x)
        dst[i] = __ldg(src + i);
    __syncthreads();
}

__global__ void __launch_bounds__(256) batch_prepare_kernel(BatchParams bp)
{
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)
    {
        int *c = reinterpret_cast<int *>(bp.bctr);
        for (int i = threadIdx.x; i < (int)(sizeof(BatchCounters) / 4); i += blockDim.x)
            c[i] = 0;
    }
    frame_prepare_tile(bp.frames[blockIdx.z], blockIdx.x, blockIdx.y);
}

__global__ void __launch_bounds__(256) batch_color_pack_kernel(BatchParams bp)
{
    color_pack_body(bp.frames[blockIdx.y], blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

// ------------------------------------------------------------------------------------------------------
// One lane per (chunk of the union box, 8^3 brick). Per frame: is the chunk inside that frame's candidate ID box and does it
// pass Frustum::Intersects (ChunkManager.cpp:182-212, exact)? If so, classify the brick against the frame's Hi-Z tiles.
//   bandM  frames in which some voxel of the brick may fall inside the truncation band
//   freeM  frames in which the brick lies in free space (only carving of observed voxels can act)
// A free-space frame is kept when the brick can hold a carvable voxel: the brick's flag is set already, or an earlier band
// frame of this batch may create one (conservative: any band frame of the batch).
template <int CS>
__global__ void __launch_bounds__(256) batch_candidates_kernel(BatchParams bp, DeviceMap map)
{
    constexpr int BPA = CS / 8, NB = BPA * BPA * BPA;
    constexpr int GL = NB >= 32 ? 32 : NB;
    constexpr int BPL = NB / GL;
    __shared__ FrameParams sF[kMaxBatch];
    load_frames(sF, bp);
    const int K = bp.K;
    const int total = bp.n[0] * bp.n[1] * bp.n[2];
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int i = (int)(tid / GL);
    const int gl = (int)(tid % GL);
    const unsigned lane = threadIdx.x & 31;
    const bool leader = gl == 0;
    const bool carve = sF[0].carve != 0;
    int x = 0, y = 0, z = 0;
    unsigned candM = 0u, bandM[BPL], freeM[BPL];
#pragma unroll
    for (int k = 0; k < BPL; k++)
        bandM[k] = freeM[k] = 0u;
    if (i < total)
    {
        const int nyz = bp.n[1] * bp.n[2];
        x = bp.lo[0] + i / nyz;
        const int r = i - (i / nyz) * nyz;
        y = bp.lo[1] + r / bp.n[2];
        z = bp.lo[2] + r % bp.n[2];
        const bool mine = map.world <= 1 || (owner_hash(x, y, z) % (unsigned)map.world) == (unsigned)map.rank;
        if (mine)
        {
            // chunk box exactly as ChunkManager.cpp:199-201
            const float ext = __fmul_rn((float)CS, map.res);
            const float bx = __fmul_rn((float)(x * CS), map.res), by = __fmul_rn((float)(y * CS), map.res), bz = __fmul_rn((float)(z * CS), map.res);
            const float ex = __fadd_rn(bx, ext), ey = __fadd_rn(by, ext), ez = __fadd_rn(bz, ext);
            for (int f = 0; f < K; f++)
            {
                const FrameParams &fp = sF[f];
                if ((unsigned)(x - fp.lo[0]) >= (unsigned)fp.n[0] || (unsigned)(y - fp.lo[1]) >= (unsigned)fp.n[1] || (unsigned)(z - fp.lo[2]) >= (unsigned)fp.n[2])
                    continue;
                if (!frustum_intersects_exact(fp, bx, by, bz, ex, ey, ez))
                    continue;
                candM |= 1u << f;
#pragma unroll
                for (int k = 0; k < BPL; k++)
                {
                    const int b = gl + k * GL;
                    const int qx = b % BPA, qy = (b / BPA) % BPA, qz = b / (BPA * BPA);
                    const int code = (NB == 1) ? classify_box(fp, bx + map.half, by + map.half, bz + map.half, (float)(CS - 1) * map.res)
                                               : classify_box(fp, bx + (float)(qx * 8) * map.res + map.half, by + (float)(qy * 8) * map.res + map.half,
                                                              bz + (float)(qz * 8) * map.res + map.half, 7.0f * map.res);
                    bandM[k] |= (code == 2 ? 1u : 0u) << f;
                    freeM[k] |= (code == 1 ? 1u : 0u) << f;
                }
            }
        }
    }
    // chunk-level unions over the group's lanes (groups are aligned sub-warps of GL lanes)
    unsigned chunkBand = 0u, chunkFree = 0u;
#pragma unroll
    for (int k = 0; k < BPL; k++)
    {
        chunkBand |= bandM[k];
        chunkFree |= freeM[k];
    }
#pragma unroll
    for (int o = 1; o < GL; o <<= 1)
    {
        chunkBand |= __shfl_xor_sync(0xffffffffu, chunkBand, o);
        chunkFree |= __shfl_xor_sync(0xffffffffu, chunkFree, o);
    }
    int slot = -1;
    unsigned long long flags = 0ull;
    if (leader && (chunkBand || (chunkFree && carve)))
    {
        slot = hash_lookup(map, pack_id(x, y, z));
        if (slot >= 0 && chunkFree && carve)
            flags = map.brick_flags[slot];
    }
    const int leaderLane = (int)(lane & ~(unsigned)(GL - 1));
    slot = __shfl_sync(0xffffffffu, slot, leaderLane);
    flags = __shfl_sync(0xffffffffu, flags, leaderLane);

    // per-frame candidate counts: lane f of the warp accumulates frame f
    int myCount = 0;
    for (int f = 0; f < K; f++)
    {
        const unsigned m = __ballot_sync(0xffffffffu, leader && ((candM >> f) & 1u));
        if ((int)lane == f)
            myCount = __popc(m);
    }
    if ((int)lane < K && myCount)
        atomicAdd(&bp.bctr->candidates[lane], myCount);

    const bool keepNew = leader && slot < 0 && chunkBand;
    const unsigned newMask = __ballot_sync(0xffffffffu, keepNew);
    int base = 0;
    if (lane == 0 && newMask)
        base = atomicAdd(&bp.bctr->new_count, __popc(newMask));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keepNew)
    {
        const int pos = base + __popc(newMask & ((1u << lane) - 1));
        if (pos < bp.news_cap)
            bp.news[pos] = make_int4(x, y, z, (int)(candM | (chunkBand << 16)));
        else
            atomicOr(&map.ctr->error_flags, kErrWorkFull);
    }
    const unsigned long long key = pack_id(x, y, z);
#pragma unroll
    for (int k = 0; k < BPL; k++)
    {
        const int b = gl + k * GL;
        unsigned m = 0u;
        if (slot >= 0)
            m = bandM[k] | ((carve && (((flags >> b) & 1ull) || bandM[k])) ? freeM[k] : 0u);
        const bool keepB = m != 0u;
        const unsigned bm = __ballot_sync(0xffffffffu, keepB);
        int ubase = 0;
        if (lane == 0 && bm)
            ubase = atomicAdd(&bp.bctr->unit_count, __popc(bm));
        ubase = __shfl_sync(0xffffffffu, ubase, 0);
        if (keepB)
        {
            const int pos = ubase + __popc(bm & ((1u << lane) - 1));
            if (pos < bp.units_cap)
                bp.units[pos] = make_int4((int)(unsigned)(key & 0xffffffffull), (int)(unsigned)(key >> 32), slot | (b << 24), (int)m);
            else
                atomicOr(&map.ctr->error_flags, kErrWorkFull);
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// per-CTA, per-frame counters in shared memory; flushed once per CTA
struct BatchShared
{
    int upd[kMaxBatch], carve[kMaxBatch], col[kMaxBatch], chunks[kMaxBatch], fresh[kMaxBatch];
};

__device__ __forceinline__ void batch_shared_zero(BatchShared *s)
{
    int *p = reinterpret_cast<int *>(s);
    for (int i = threadIdx.x; i < (int)(sizeof(BatchShared) / 4); i += blockDim.x)
        p[i] = 0;
}

// warp-level: add this frame's per-lane voxel counts to the CTA's shared counters
__device__ __forceinline__ void batch_count_frame(BatchShared *s, int f, int nUpd, int nCarve, int nCol, int lane)
{
    const int a = __reduce_add_sync(0xffffffffu, nUpd), b = __reduce_add_sync(0xffffffffu, nCarve), c = __reduce_add_sync(0xffffffffu, nCol);
    if (lane == 0)
    {
        if (a) atomicAdd(&s->upd[f], a);
        if (b) atomicAdd(&s->carve[f], b);
        if (c) atomicAdd(&s->col[f], c);
    }
}

__device__ __forceinline__ void batch_flush(const BatchParams &bp, BatchShared *s)
{
    __syncthreads();
    const int t = threadIdx.x;
    if (t < bp.K)
    {
        BatchCounters *c = bp.bctr;
        if (s->upd[t]) atomicAdd(&c->n_upd[t], (unsigned long long)s->upd[t]);
        if (s->carve[t]) atomicAdd(&c->n_carve[t], (unsigned long long)s->carve[t]);
        if (s->col[t]) atomicAdd(&c->n_col[t], (unsigned long long)s->col[t]);
        if (s->chunks[t]) atomicAdd(&c->updated_chunks[t], s->chunks[t]);
        if (s->fresh[t]) atomicAdd(&c->n_new[t], s->fresh[t]);
    }
}

// The last CTA of the batch's last kernel copies the counters into the pinned host slot.
__device__ __forceinline__ void batch_snapshot(const BatchParams &bp, const DeviceMap &map)
{
    __shared__ int sLast;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        __threadfence();
        sLast = atomicAdd(&bp.bctr->tickets, 1) == bp.total_ctas - 1;
    }
    __syncthreads();
    if (!sLast)
        return;
    __threadfence();
    const volatile BatchCounters *c = bp.bctr;
    const volatile Counters *g = map.ctr;
    volatile HostBatchSnapshot *h = bp.host_slot;
    const int t = threadIdx.x;
    if (t == 0)
    {
        h->head = bp.batch_id;
        h->n_chunks = g->n_chunks;
        h->n_dirty = g->n_dirty;
        h->error_flags = g->error_flags;
        h->unit_count = c->unit_count;
        h->new_count = c->new_count;
        h->K = bp.K;
    }
    if (t < kMaxBatch)
    {
        h->candidates[t] = c->candidates[t];
        h->n_new[t] = c->n_new[t];
        h->updated_chunks[t] = c->updated_chunks[t];
        h->n_upd[t] = (long long)c->n_upd[t];
        h->n_carve[t] = (long long)c->n_carve[t];
        h->n_col[t] = (long long)c->n_col[t];
    }
    __threadfence_system();
    __syncthreads();
    if (t == 0)
        h->tail = bp.batch_id;
}

// ------------------------------------------------------------------------------------------------------
// One frame applied to the lane's eight voxels of a half brick (z slices 4*half .. 4*half+3, y rows ly and ly+4), state in
// registers. Same arithmetic as process_batch<MODE 0> (ProjectionIntegrator.h:51-183); colour and depth cameras coincide.
template <int CS, bool COLOR_PATH, bool PER_PIXEL>
__device__ __forceinline__ void frame_on_half_brick(const FrameParams &fp, const DeviceMap &map, const BrickLane &L, int half, bool hasCol,
                                                    float2 (&dv)[8], unsigned (&cv)[8], unsigned &wroteD, unsigned &wroteC,
                                                    int &nUpd, int &nCarve, int &nCol, bool &carvable)
{
    const CameraDev &c = fp.cam;
    int pix[8];
    float cz[8], depth[8], trunc[8];
    unsigned cpx[8];
#pragma unroll
    for (int s = 0; s < 4; s++)
    {
        const float pz = __fadd_rn(__fadd_rn(__fmul_rn((float)(L.vz0 + 4 * half + s), map.res), map.half), L.orgz);
        const float d2 = __fsub_rn(pz, c.t[2]);
        const float m20 = __fmul_rn(c.R[6], d2), m21 = __fmul_rn(c.R[7], d2), m22 = __fmul_rn(c.R[8], d2);
#pragma unroll
        for (int h = 0; h < 2; h++)
        {
            const int k = s * 2 + h;
            const float cx = __fadd_rn(L.m0[0], __fadd_rn(L.m1[h][0], m20));
            const float cy = __fadd_rn(L.m0[1], __fadd_rn(L.m1[h][1], m21));
            cz[k] = __fadd_rn(L.m0[2], __fadd_rn(L.m1[h][2], m22));
            const float invZ = __frcp_rn(cz[k]);                                           // == 1.0f / z, correctly rounded
            const float u = __fadd_rn(__fmul_rn(__fmul_rn(c.fx, cx), invZ), c.cx);
            const float v = __fadd_rn(__fmul_rn(__fmul_rn(c.fy, cy), invZ), c.cy);
            const bool on = u >= 0.0f && v >= 0.0f && u < c.Wf && v < c.Hf && !(cz[k] < 0.0f);
            pix[k] = on ? (int)u + (int)v * c.W : -1;
        }
    }
#pragma unroll
    for (int k = 0; k < 8; k++)
    {
        depth[k] = pix[k] >= 0 ? __ldg(fp.depth + pix[k]) : __int_as_float(0x7fc00000);
        cpx[k] = (COLOR_PATH && hasCol && pix[k] >= 0) ? __ldg(fp.color_packed + pix[k]) : 0u;
        trunc[k] = PER_PIXEL ? (pix[k] >= 0 ? __ldg(fp.trunc_img + pix[k]) : 0.0f) : fp.trunc_param;
    }
#pragma unroll
    for (int k = 0; k < 8; k++)
    {
        const float d = depth[k];
        const bool skip = (pix[k] < 0) || (COLOR_PATH ? (d != d || d > 100.0f) : (d > 50.0f));
        if (skip)
            continue;
        const float sd = __fsub_rn(d, cz[k]);
        if (fabsf(sd) < __fadd_rn(trunc[k], fp.diag))                                       // :82 / :143
        {
            float wu = 1.0f;
            if (COLOR_PATH)
            {
                if (hasCol && (cv[k] >> 24) < 8u)                                           // :153
                {
                    cv[k] = color_integrate_packed(cv[k], cpx[k]);
                    wroteC |= 1u << k;
                    nCol++;
                }
                wu = PER_PIXEL ? __fdiv_rn(fp.weight, __fmul_rn(5.0f, trunc[k])) : fp.wu_const;
            }
            dv[k] = dist_integrate(dv[k], sd, wu);
            wroteD |= 1u << k;
            nUpd++;
            carvable |= dv[k].y > 0.0f && dv[k].x < fp.sdf_carve_max;
        }
        else if (fp.carve && sd > __fadd_rn(trunc[k], fp.carve_dist))                       // :88 / :166
        {
            if (dv[k].y > 0.0f && dv[k].x < fp.sdf_carve_max)                               // :90 / :169
            {
                if (COLOR_PATH && !(dv[k].y < 5.0f))
                    dv[k].y = __fsub_rn(dv[k].y, 1.0f);                                     // :171-175
                else
                    dv[k] = make_float2(99999.0f, 0.0f);
                wroteD |= 1u << k;
                nCarve++;
                carvable |= dv[k].y > 0.0f && dv[k].x < fp.sdf_carve_max;
            }
        }
    }
}

// Existing chunks: a warp per half brick, persistent grid striding over the unit list.
template <int CS, bool COLOR_PATH, bool PER_PIXEL>
__global__ void __launch_bounds__(256, 2) batch_bricks_kernel(BatchParams bp, DeviceMap map)
{
    constexpr int BPA = CS / 8;
    __shared__ FrameParams sF[kMaxBatch];
    __shared__ BatchShared sB;
    batch_shared_zero(&sB);
    load_frames(sF, bp);
    const int lane = threadIdx.x & 31;
    const int nTasks = min(bp.bctr->unit_count, bp.units_cap) * 2;
    const bool hasCol = COLOR_PATH && map.use_color;
    for (int g = blockIdx.x * 8 + (threadIdx.x >> 5); g < nTasks; g += gridDim.x * 8)
    {
        const int4 unit = bp.units[g >> 1];
        const int half = g & 1;
        int x, y, z;
        unpack_id(((unsigned long long)(unsigned)unit.y << 32) | (unsigned long long)(unsigned)unit.x, &x, &y, &z);
        const int slot = unit.z & 0xFFFFFF, b = unit.z >> 24;
        unsigned mask = (unsigned)unit.w;
        const float orgx = __fmul_rn((float)(CS * x), map.res), orgy = __fmul_rn((float)(CS * y), map.res), orgz = __fmul_rn((float)(CS * z), map.res);
        const int bx = b % BPA, by = (b / BPA) % BPA, bz = b / (BPA * BPA);
        float2 *dist = dist_ptr(map, slot);
        unsigned *col = hasCol ? reinterpret_cast<unsigned *>(color_ptr(map, slot)) : nullptr;
        const int idx0 = ((bz * 8 + 4 * half) * CS + (by * 8 + (lane >> 3))) * CS + bx * 8 + (lane & 7);
        float2 dv[8];
        unsigned cv[8];
#pragma unroll
        for (int k = 0; k < 8; k++)
        {
            const int idx = idx0 + (k >> 1) * CS * CS + (k & 1) * 4 * CS;
            dv[k] = dist[idx];
            cv[k] = hasCol ? col[idx] : 0u;
        }
        unsigned wroteD = 0u, wroteC = 0u, updMask = 0u;
        bool carvable = false;
        while (mask)
        {
            const int f = __ffs(mask) - 1;
            mask &= mask - 1;
            const FrameParams &fp = sF[f];
            const BrickLane L = brick_lane_setup(fp, map, orgx, orgy, orgz, bx, by, bz, lane);
            int nUpd = 0, nCarve = 0, nCol = 0;
            const unsigned before = wroteD;
            unsigned wd = 0u;
            frame_on_half_brick<CS, COLOR_PATH, PER_PIXEL>(fp, map, L, half, hasCol, dv, cv, wd, wroteC, nUpd, nCarve, nCol, carvable);
            wroteD = before | wd;
            batch_count_frame(&sB, f, nUpd, nCarve, nCol, lane);
            if (__any_sync(0xffffffffu, wd != 0u))
                updMask |= 1u << f;
        }
#pragma unroll
        for (int k = 0; k < 8; k++)
        {
            const int idx = idx0 + (k >> 1) * CS * CS + (k & 1) * 4 * CS;
            if ((wroteD >> k) & 1u)
                dist[idx] = dv[k];
            if (hasCol && ((wroteC >> k) & 1u))
                col[idx] = cv[k];
        }
        if (__any_sync(0xffffffffu, carvable) && lane == 0)
            atomicOr(&map.brick_flags[slot], 1ull << b);
        if (updMask)
        {
            // (batch id << 32) | frames of this batch that updated the chunk: the first warp of the batch marks the 27 neighbour IDs
            // dirty (Chisel.h:89-101, 175-189); every newly set frame bit counts the chunk once for that frame
            unsigned newBits = 0u;
            int first = 0;
            if (lane == 0)
            {
                const unsigned long long tag = (unsigned long long)(unsigned)bp.batch_id << 32;
                unsigned long long old = bp.slot_batch[slot], assumed, cur;
                do
                {
                    assumed = old;
                    cur = ((assumed >> 32) == (unsigned long long)(unsigned)bp.batch_id) ? assumed : tag;
                    const unsigned long long nw = cur | updMask;
                    if (nw == assumed)
                        break;
                    old = atomicCAS(&bp.slot_batch[slot], assumed, nw);
                } while (old != assumed);
                first = (assumed >> 32) != (unsigned long long)(unsigned)bp.batch_id;
                newBits = updMask & ~(unsigned)(cur & 0xffffffffull);
            }
            first = __shfl_sync(0xffffffffu, first, 0);
            newBits = __shfl_sync(0xffffffffu, newBits, 0);
            if (first && lane < 27)
                dirty_insert(map, pack_id(x + lane / 9 - 1, y + (lane / 3) % 3 - 1, z + lane % 3 - 1));
            if (lane < kMaxBatch && ((newBits >> lane) & 1u))
                atomicAdd(&sB.chunks[lane], 1);
        }
    }
    batch_flush(bp, &sB);
    batch_snapshot(bp, map);
}

// ------------------------------------------------------------------------------------------------------
// Chunks that do not exist at the start of the batch: one CTA per chunk, a warp per brick. Frames in order: until the chunk
// exists only frames with a possible band hit are tested (exactly); the first hit allocates it ("created and untouched =>
// garbage collected", Chisel.h:76-80,102-110 / :133-143,170-173,202-207, never allocates anything) and writes every voxel once;
// for the remaining frames it is an ordinary chunk (state goes through global memory: the same lane owns the same voxels).
template <int CS, bool COLOR_PATH, bool PER_PIXEL>
__global__ void __launch_bounds__(256) batch_new_chunks_kernel(BatchParams bp, DeviceMap map)
{
    constexpr int BPA = CS / 8, NB = BPA * BPA * BPA;
    __shared__ FrameParams sF[kMaxBatch];
    __shared__ BatchShared sB;
    __shared__ int sSlot;
    batch_shared_zero(&sB);
    load_frames(sF, bp);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nWarps = blockDim.x >> 5;
    const int nNew = min(bp.bctr->new_count, bp.news_cap);
    const int K = bp.K;
    for (int w = blockIdx.x; w < nNew; w += gridDim.x)
    {
        const int4 item = bp.news[w];
        const unsigned candM = (unsigned)item.w & 0xFFFFu, bandM = (unsigned)item.w >> 16;
        const float orgx = __fmul_rn((float)(CS * item.x), map.res), orgy = __fmul_rn((float)(CS * item.y), map.res), orgz = __fmul_rn((float)(CS * item.z), map.res);
        int slot = -1;
        float2 *dist = nullptr;
        unsigned *col = nullptr;
        for (int f = 0; f < K; f++)
        {
            if (!((candM >> f) & 1u))
                continue;
            const FrameParams &fp = sF[f];
            VoxelStats st;
            st.nUpd = st.nCarve = st.nCol = 0;
            st.updated = st.carvable = false;
            if (slot < 0)
            {
                if (!((bandM >> f) & 1u))
                    continue;
                bool any = false;
                for (int b = warp; b < NB && !any; b += nWarps)
                {
                    const int bx = b % BPA, by = (b / BPA) % BPA, bz = b / (BPA * BPA);
                    if (NB > 1 && classify_box(fp, orgx + (float)(bx * 8) * map.res + map.half, orgy + (float)(by * 8) * map.res + map.half,
                                               orgz + (float)(bz * 8) * map.res + map.half, 7.0f * map.res) != 2)
                        continue;
                    const BrickLane L = brick_lane_setup(fp, map, orgx, orgy, orgz, bx, by, bz, lane);
                    bool hit = false;
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        hit |= process_batch<CS, COLOR_PATH, PER_PIXEL, 2>(fp, map, L, q, nullptr, nullptr, &st);
                    any = __any_sync(0xffffffffu, hit);
                }
                if (!__syncthreads_or(any))
                    continue;
                if (warp == (nWarps > 1 ? 1 : 0) && lane < 27)
                    dirty_insert(map, pack_id(item.x + lane / 9 - 1, item.y + (lane / 3) % 3 - 1, item.z + lane % 3 - 1));
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
                        map.slot_epoch[s] = 0;
                        bp.slot_batch[s] = ((unsigned long long)(unsigned)bp.batch_id << 32) | 0xFFFFull;
                        hash_insert_new(map, pack_id(item.x, item.y, item.z), s);
                        sB.fresh[f] += 1;
                        sB.chunks[f] += 1;
                    }
                    sSlot = s;
                }
                __syncthreads();
                slot = sSlot;
                __syncthreads();
                if (slot < 0)
                    break;                                              // pool full: reported through error_flags
                dist = dist_ptr(map, slot);
                col = map.use_color ? reinterpret_cast<unsigned *>(color_ptr(map, slot)) : nullptr;
                for (int b = warp; b < NB; b += nWarps)
                {
                    const BrickLane L = brick_lane_setup(fp, map, orgx, orgy, orgz, b % BPA, (b / BPA) % BPA, b / (BPA * BPA), lane);
                    st.carvable = false;
#pragma unroll 1
                    for (int q = 0; q < 4; q++)
                        process_batch<CS, COLOR_PATH, PER_PIXEL, 1>(fp, map, L, q, dist, col, &st);
                    if (__any_sync(0xffffffffu, st.carvable) && lane == 0)
                        atomicOr(&map.brick_flags[slot], 1ull << b);
                }
                batch_count_frame(&sB, f, st.nUpd, st.nCarve, st.nCol, lane);
            }
            else
            {
                bool updated = false;
                for (int b = warp; b < NB; b += nWarps)
                {
                    const int bx = b % BPA, by = (b / BPA) % BPA, bz = b / (BPA * BPA);
                    const int code = (NB == 1) ? classify_box(fp, orgx + map.half, orgy + map.half, orgz + map.half, (float)(CS - 1) * map.res)
                                               : classify_box(fp, orgx + (float)(bx * 8) * map.res + map.half, orgy + (float)(by * 8) * map.res + map.half,
                                                              orgz + (float)(bz * 8) * map.res + map.half, 7.0f * map.res);
                    if (code == 0 || (code == 1 && !fp.carve))
                        continue;
                    const BrickLane L = brick_lane_setup(fp, map, orgx, orgy, orgz, bx, by, bz, lane);
                    st.updated = st.carvable = false;
#pragma unroll 1
                    for (int q = 0; q < 4; q++)
                        process_batch<CS, COLOR_PATH, PER_PIXEL, 0>(fp, map, L, q, dist, col, &st);
                    if (__any_sync(0xffffffffu, st.carvable) && lane == 0)
                        atomicOr(&map.brick_flags[slot], 1ull << b);
                    updated |= __any_sync(0xffffffffu, st.updated);
                }
                batch_count_frame(&sB, f, st.nUpd, st.nCarve, st.nCol, lane);
                if (__syncthreads_or(updated) && t == 0)
                    sB.chunks[f] += 1;
            }
        }
        __syncthreads();
    }
    batch_flush(bp, &sB);
}

// ------------------------------------------------------------------------------------------------------
// host launcher

template <typename Kern>
static int batch_resident(Kern kernel, int threads)
{
    int dev = 0, sms = 148, perSm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, threads, 0);
    return sms * (perSm > 0 ? perSm : 1);
}

template <int CS, bool COLOR_PATH, bool PER_PIXEL>
static cudaError_t launch_batch_variant(BatchParams bp, const DeviceMap &map, const BatchLaunchInfo &info, cudaEvent_t *evt, cudaStream_t st)
{
    static int residentNew = 0, residentBricks = 0;
    if (!residentNew)
    {
        residentNew = batch_resident(batch_new_chunks_kernel<CS, COLOR_PATH, PER_PIXEL>, 256);
        residentBricks = batch_resident(batch_bricks_kernel<CS, COLOR_PATH, PER_PIXEL>, 256);
    }
    constexpr long long NB = (CS / 8) * (CS / 8) * (CS / 8);
    const long long lanes = info.unionCandidates * std::min<long long>(NB, 32);
    const unsigned gCand = (unsigned)std::max(1ll, (lanes + 255) / 256);
    const unsigned gNew = (unsigned)std::max(1ll, std::min<long long>(std::min<long long>(info.unionCandidates, info.newHint), residentNew));
    const unsigned gBricks = (unsigned)std::max(1ll, std::min<long long>((info.unionCandidates * NB * 2 + 7) / 8, residentBricks));
    bp.total_ctas = (int)gBricks;
    cudaError_t e;
    if (info.profiling && (e = cudaEventRecord(evt[0], st)) != cudaSuccess)
        return e;
    batch_prepare_kernel<<<dim3((info.W + 63) / 64, (info.H + 63) / 64, bp.K), 256, 0, st>>>(bp);
    if (info.colorPath)
    {
        const int px = info.cW * info.cH;
        batch_color_pack_kernel<<<dim3((unsigned)std::max(1, std::min(148 * 2, (px / 4 + 255) / 256)), bp.K), 256, 0, st>>>(bp);
    }
    if (info.profiling && (e = cudaEventRecord(evt[1], st)) != cudaSuccess)
        return e;
    batch_candidates_kernel<CS><<<gCand, 256, 0, st>>>(bp, map);
    if (info.profiling && (e = cudaEventRecord(evt[2], st)) != cudaSuccess)
        return e;
    batch_new_chunks_kernel<CS, COLOR_PATH, PER_PIXEL><<<gNew, 256, 0, st>>>(bp, map);
    if (info.profiling && (e = cudaEventRecord(evt[7], st)) != cudaSuccess)
        return e;
    batch_bricks_kernel<CS, COLOR_PATH, PER_PIXEL><<<gBricks, 256, 0, st>>>(bp, map);
    if (info.profiling && (e = cudaEventRecord(evt[3], st)) != cudaSuccess)
        return e;
    return cudaGetLastError();
}

template <int CS>
static cudaError_t launch_batch_cs(const BatchParams &bp, const DeviceMap &map, const BatchLaunchInfo &info, cudaEvent_t *evt, cudaStream_t st)
{
    if (info.colorPath)
        return info.perPixel ? launch_batch_variant<CS, true, true>(bp, map, info, evt, st) : launch_batch_variant<CS, true, false>(bp, map, info, evt, st);
    return info.perPixel ? launch_batch_variant<CS, false, true>(bp, map, info, evt, st) : launch_batch_variant<CS, false, false>(bp, map, info, evt, st);
}

cudaError_t launch_batch(const BatchParams &bp, const DeviceMap &map, const BatchLaunchInfo &info, cudaEvent_t *evt, cudaStream_t st)
{
    switch (map.cs)
    {
    case 8: return launch_batch_cs<8>(bp, map, info, evt, st);
    case 16: return launch_batch_cs<16>(bp, map, info, evt, st);
    default: return launch_batch_cs<32>(bp, map, info, evt, st);
    }
}

} // namespace chs
