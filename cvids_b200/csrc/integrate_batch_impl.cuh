// integrate_batch_impl.cuh -- fused multi-frame TSDF integration (sm_100a): K <= 16 consecutive frames in ONE pass over the map.
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
//   batch_prepare_kernel      grid.y = frame: Hi-Z tiles (+ per-pixel truncation) and the packed colour image of every frame
//   batch_candidates_kernel   a group of lanes per chunk of the UNION candidate box: exact Frustum::Intersects and the
//                             conservative depth-range class per frame (chunk level over frames, then brick level for the
//                             survivors) -> per-brick frame masks; warp-ballot compaction, long units first
//   batch_bricks_kernel       warp per quarter brick (8 x 8 x 2 voxels): state of 4 voxels per lane in registers, straight-line
//                             update per frame of the brick's frame mask, one store per changed voxel at the end; tasks from an
//                             atomic queue. Bricks of chunks that do not exist yet start from the initial state in registers and
//                             create the chunk (hash insert + pool bump, no voxel traffic: free pool slots are kept
//                             initialised) on the first hit
#include <algorithm>
#include <cstddef>
#include <cstdlib>

#include "device_map.cuh"
#include "integrate_device.cuh"
#include "kernels.h"
#include "tma.cuh"

#ifndef CHS_BATCH_VARIANT
#error "compiled through integrate_batch_half.cu / integrate_batch_quarter.cu, which choose the task size of the brick kernels"
#endif

namespace chs
{
// Two builds of this file live in the library, in namespaces of their own: `half` (half-brick tasks, 8 voxels per lane, 128
// threads x 4 CTAs/SM) for depth-only batches and `quarter` (quarter-brick tasks, 4 voxels per lane, 256 threads x 2 CTAs/SM) for
// colour batches. Measured on B200: depth-only 67 us (half) vs 84 us (quarter) per 16-frame step of configs[4]; colour 80 us
// (quarter) vs 100 us (half, 120 bytes of spills per thread) per 10-frame step of configs[1]. capi.cu picks per batch.
namespace CHS_BATCH_VARIANT
{

static_assert(sizeof(FrameParams) % 4 == 0, "FrameParams is copied word-wise into shared memory");
constexpr int kVirtualSlot = 0xFFFFFF;      // unit of a chunk that does not exist yet
constexpr int kCoarseTiles = 48;            // Hi-Z tiles of the levels >= 4 of one frame that batch_candidates_kernel keeps in shared memory
constexpr int kNoSlot = -2;                 // hash value of a key whose chunk could not be allocated (pool exhausted); -1 = "being created"

// z slices of a brick per task: 4 = half brick (8 voxels per lane), 2 = quarter brick (4 voxels per lane)
#ifndef CHS_BRICK_SLICES
#define CHS_BRICK_SLICES 4                   // measured on B200 (fast kernel, configs[4], 16 frames per step, 128 registers, 16 warps / SM):
#endif                                       // 67 us with 4 vs 84 us with 2; with 4 and 168 registers (12 warps) 71 us, 96 registers (20 warps, spills) 77 us
#ifndef CHS_BRICK_THREADS
#define CHS_BRICK_THREADS 256
#endif
#ifndef CHS_BRICK_MIN_CTAS
#define CHS_BRICK_MIN_CTAS 3
#endif
constexpr int kNS = CHS_BRICK_SLICES;        // z slices per task
constexpr int kVPL = 2 * kNS;                // voxels per lane: kNS slices x the lane's two y rows
constexpr int kParts = 8 / kNS;              // tasks per brick

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may START while the
// kernel before it in the stream is still running; pdl_wait() blocks until that kernel has completed and its writes are visible,
// pdl_launch_dependents() lets the next kernel of the stream start launching. Without the attribute both are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// device timeline (BatchParams::timeline): one thread per CTA
__device__ __forceinline__ void tl_start(const BatchParams &bp, int i)
{
    if (bp.timeline && threadIdx.x == 0)
        atomicMax(&bp.timeline[i], ~global_timer_ns());
}
__device__ __forceinline__ void tl_end(const BatchParams &bp, int i)
{
    if (bp.timeline)
    {
        __syncthreads();
        if (threadIdx.x == 0)
            atomicMax(&bp.timeline[i], global_timer_ns());
    }
}

// All frames of the batch into shared memory: afterwards `sF[f]` is read with warp-uniform shared loads.
__device__ __forceinline__ void load_frames(FrameParams *sF, const BatchParams &bp)
{
    static_assert(sizeof(FrameParams) % 16 == 0, "FrameParams is copied as 16-byte words");
    const int nQuads = bp.K * (int)(sizeof(FrameParams) / 16);
    const uint4 *src = reinterpret_cast<const uint4 *>(bp.frames);
    uint4 *dst = reinterpret_cast<uint4 *>(sF);
    // all loads of a thread in flight at once (the table is a few KB: one round trip instead of one per word)
    uint4 v[8];
#pragma unroll
    for (int r = 0; r < 8; r++)
    {
        const int i = (int)threadIdx.x + r * (int)blockDim.x;
        if (i < nQuads)
            v[r] = __ldg(src + i);
    }
#pragma unroll
    for (int r = 0; r < 8; r++)
    {
        const int i = (int)threadIdx.x + r * (int)blockDim.x;
        if (i < nQuads)
            dst[i] = v[r];
    }
    for (int i = (int)threadIdx.x + 8 * (int)blockDim.x; i < nQuads; i += blockDim.x)
        dst[i] = __ldg(src + i);
    __syncthreads();
}

// grid = (64x64 pixel tiles, K): Hi-Z levels (+ per-pixel truncation, millimetre conversion) of every frame of the batch
__global__ void __launch_bounds__(256) batch_hiz_kernel(BatchParams bp, DeviceMap map, int tilesX, int frameBase)
{
    tl_start(bp, kTlHizStart);
    const FrameParams &fp = bp.frames[frameBase + blockIdx.y];
    frame_prepare_tile(fp, blockIdx.x % tilesX, blockIdx.x / tilesX);
    tl_end(bp, kTlHizEnd);
}

// grid = (pack blocks, K): packed colour image of every frame (ColorImage::At once per pixel). Only the brick kernel reads it, so
// this runs beside the candidates kernel.
__global__ void __launch_bounds__(256) batch_pack_kernel(BatchParams bp)
{
    const FrameParams &fp = bp.frames[blockIdx.y];
    color_pack_body(fp, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

// ---- Hi-Z with TMA bulk copies -----------------------------------------------------------------------------------------
// The common case (float depth, constant truncator, W % 4 == 0, 16-byte aligned image): a PERSISTENT grid; every CTA walks
// (frame, 64x64 tile) items with a three-stage ring of shared-memory tiles filled by cp.async.bulk row copies (the TMA
// engine; one mbarrier per stage counts the bytes), so that each CTA keeps up to 48 KB of DRAM reads in flight without a
// register or a thread waiting on them. Levels 0..3 go to global memory; the coarser levels are built by the candidates
// kernel in shared memory (no "last block of the frame" serialisation).
constexpr int kHizStages = 3;
constexpr int kHizPitch = 68;                          // floats per shared row: 64 + 4 (rows stay 16-byte aligned)
constexpr int kHizThreads = 256;
constexpr size_t kHizSmem = (size_t)kHizStages * 64 * kHizPitch * sizeof(float);

// PEERS: sharded Hi-Z of a distributed step -- every tile is stored into the arena of every rank (HizPeers), the last CTA raises
// this rank's arrived word everywhere. The kernel then runs on the push stream, possibly a whole step ahead: it touches neither
// the batch counters nor the timeline words of a staging set.
template <bool PEERS>
__device__ __forceinline__ void hiz_store(float2 *p, float2 v, const HizPeers &hp)
{
    if (!PEERS)
        *p = v;
    else
        for (int d = 0; d < hp.world; d++)
            *reinterpret_cast<float2 *>(reinterpret_cast<char *>(p) + hp.delta[d]) = v;
}

template <bool PEERS>
__global__ void __launch_bounds__(kHizThreads) batch_hiz_tma_kernel(BatchParams bp, int tilesX, int tilesPerFrame, int nItems, int frameBase, const __grid_constant__ HizPeers hp)
{
    extern __shared__ __align__(128) unsigned char hizSmem[];
    float *stage = reinterpret_cast<float *>(hizSmem);
    __shared__ __align__(8) unsigned long long full[kHizStages];
    __shared__ float2 s0[64], s1[16], s2[4];
    const int t = threadIdx.x, lane = t & 31;
    if (!PEERS)
        tl_start(bp, kTlHizStart);
    else if (blockIdx.x == 0 && t == 0)
        reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(hp.hdr) + 2048)[8 + 2 * hp.stamp_slot] = global_timer_ns();
    if (t == 0)
    {
        for (int s = 0; s < kHizStages; s++)
            mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int nMine = ((int)blockIdx.x < nItems) ? (nItems - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    // what the arithmetic needs of a frame is the same for all frames of a batch (one integrator, one camera)
    int W, H, carve, hw0, hh0, hw1, hh1, hw2, hh2, hw3;
    float cutoff, tr, diag, carveDist;
    if (PEERS)
    {
        W = hp.W; H = hp.H; carve = hp.carve; cutoff = hp.cutoff; tr = hp.trunc; diag = hp.diag; carveDist = hp.carve_dist;
        hw0 = hp.hizW[0]; hh0 = hp.hizH[0]; hw1 = hp.hizW[1]; hh1 = hp.hizH[1]; hw2 = hp.hizW[2]; hh2 = hp.hizH[2]; hw3 = hp.hizW[3];
    }
    else
    {
        const FrameParams &f0 = bp.frames[0];
        W = f0.cam.W; H = f0.cam.H; carve = f0.carve; cutoff = f0.depth_cutoff; tr = f0.trunc_param; diag = f0.diag; carveDist = f0.carve_dist;
        hw0 = f0.hizW[0]; hh0 = f0.hizH[0]; hw1 = f0.hizW[1]; hh1 = f0.hizH[1]; hw2 = f0.hizW[2]; hh2 = f0.hizH[2]; hw3 = f0.hizW[3];
    }
    // warp 0 issues the row copies of item j into stage s: lane r rows r and r + 32
    auto issue = [&](int j, int s)
    {
        const int item = (int)blockIdx.x + j * (int)gridDim.x;
        const int fr = frameBase + item / tilesPerFrame;
        const float *depth = PEERS ? hp.depth0 + (size_t)fr * hp.npx : bp.frames[fr].depth;
        const int tile = item % tilesPerFrame, x0 = (tile % tilesX) * 64, y0 = (tile / tilesX) * 64;
        const unsigned rowBytes = (unsigned)min(64, W - x0) * 4u;
        const int nRows = min(64, H - y0);
        if (lane == 0)
            mbar_expect_tx(&full[s], rowBytes * (unsigned)nRows);
        __syncwarp();
        float *dst = stage + (size_t)s * 64 * kHizPitch;
        for (int r = lane; r < nRows; r += 32)
            bulk_g2s(dst + r * kHizPitch, depth + (size_t)(y0 + r) * W + x0, rowBytes, &full[s]);
    };
    if (t < 32)
        for (int j = 0; j < min(kHizStages, nMine); j++)
            issue(j, j);
    const int tile8 = t >> 2, sub = t & 3;                 // 64 tiles of 8x8 pixels, 4 threads per tile (2 rows each)
    for (int j = 0; j < nMine; j++)
    {
        const int s = j % kHizStages;
        const int item = (int)blockIdx.x + j * (int)gridDim.x;
        const int fr = frameBase + item / tilesPerFrame;
        float2 *h0, *h1, *h2, *h3;
        if (PEERS)
        {
            float2 *base = hp.hiz0 + (size_t)fr * hp.tiles_per_frame;
            h0 = base + hp.level_off[0]; h1 = base + hp.level_off[1]; h2 = base + hp.level_off[2]; h3 = base + hp.level_off[3];
        }
        else
        {
            const FrameParams &fp = bp.frames[fr];
            h0 = fp.hiz[0]; h1 = fp.hiz[1]; h2 = fp.hiz[2]; h3 = fp.hiz[3];
        }
        const int tile = item % tilesPerFrame, bx = tile % tilesX, by = tile / tilesX;
        mbar_wait(&full[s], (unsigned)(j / kHizStages) & 1u);
        const float *src = stage + (size_t)s * 64 * kHizPitch;
        const int lx = (tile8 & 7) * 8, ly = (tile8 >> 3) * 8 + sub * 2;
        // two independent min / max chains per thread (one per row)
        float lo0 = INFINITY, hi0 = -INFINITY, lo1 = INFINITY, hi1 = -INFINITY;
        if (bx * 64 + lx < W)
        {
            const float nanv = __int_as_float(0x7fc00000);
            const float4 nan4 = make_float4(nanv, nanv, nanv, nanv);
            const bool in0 = by * 64 + ly < H, in1 = by * 64 + ly + 1 < H, right = bx * 64 + lx + 4 < W;
            const float4 *r0 = reinterpret_cast<const float4 *>(src + ly * kHizPitch + lx);
            const float4 *r1 = reinterpret_cast<const float4 *>(src + (ly + 1) * kHizPitch + lx);
            const float4 a = in0 ? r0[0] : nan4, b = (in0 && right) ? r0[1] : nan4;
            const float4 c = in1 ? r1[0] : nan4, d = (in1 && right) ? r1[1] : nan4;
            hiz_minmax(a.x, cutoff, &lo0, &hi0); hiz_minmax(c.x, cutoff, &lo1, &hi1);
            hiz_minmax(a.y, cutoff, &lo0, &hi0); hiz_minmax(c.y, cutoff, &lo1, &hi1);
            hiz_minmax(a.z, cutoff, &lo0, &hi0); hiz_minmax(c.z, cutoff, &lo1, &hi1);
            hiz_minmax(a.w, cutoff, &lo0, &hi0); hiz_minmax(c.w, cutoff, &lo1, &hi1);
            hiz_minmax(b.x, cutoff, &lo0, &hi0); hiz_minmax(d.x, cutoff, &lo1, &hi1);
            hiz_minmax(b.y, cutoff, &lo0, &hi0); hiz_minmax(d.y, cutoff, &lo1, &hi1);
            hiz_minmax(b.z, cutoff, &lo0, &hi0); hiz_minmax(d.z, cutoff, &lo1, &hi1);
            hiz_minmax(b.w, cutoff, &lo0, &hi0); hiz_minmax(d.w, cutoff, &lo1, &hi1);
        }
        float lo = fminf(lo0, lo1), hi = fmaxf(hi0, hi1);
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, 1));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, 1));
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, 2));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, 2));
        const int tx = bx * 8 + (tile8 & 7), ty = by * 8 + (tile8 >> 3);
        if (sub == 0)
        {
            const float2 v = hiz_apply_band(lo, hi, tr, diag, carve, carveDist);
            s0[tile8] = v;
            if (tx < hw0 && ty < hh0)
                hiz_store<PEERS>(h0 + ty * hw0 + tx, v, hp);
        }
        __syncthreads();                                    // also: every thread is done reading the stage
        if (t < 32 && j + kHizStages < nMine)
            issue(j + kHizStages, s);
        if (t < 16)
        {
            const int ax = t & 3, ay = t >> 2;
            float2 a = s0[(ay * 2) * 8 + ax * 2], b = s0[(ay * 2) * 8 + ax * 2 + 1], c = s0[(ay * 2 + 1) * 8 + ax * 2], d = s0[(ay * 2 + 1) * 8 + ax * 2 + 1];
            const float2 v = make_float2(fminf(fminf(a.x, b.x), fminf(c.x, d.x)), fmaxf(fmaxf(a.y, b.y), fmaxf(c.y, d.y)));
            s1[t] = v;
            const int gx = bx * 4 + ax, gy = by * 4 + ay;
            if (gx < hw1 && gy < hh1)
                hiz_store<PEERS>(h1 + gy * hw1 + gx, v, hp);
        }
        __syncthreads();
        if (t < 4)
        {
            const int ax = t & 1, ay = t >> 1;
            float2 a = s1[(ay * 2) * 4 + ax * 2], b = s1[(ay * 2) * 4 + ax * 2 + 1], c = s1[(ay * 2 + 1) * 4 + ax * 2], d = s1[(ay * 2 + 1) * 4 + ax * 2 + 1];
            const float2 v = make_float2(fminf(fminf(a.x, b.x), fminf(c.x, d.x)), fmaxf(fmaxf(a.y, b.y), fmaxf(c.y, d.y)));
            s2[t] = v;
            const int gx = bx * 2 + ax, gy = by * 2 + ay;
            if (gx < hw2 && gy < hh2)
                hiz_store<PEERS>(h2 + gy * hw2 + gx, v, hp);
            if (t == 0)
            {
                // level 3 from the four level-2 values of this thread's neighbours: recomputed from s1 to save a barrier
                float mn = INFINITY, mx = -INFINITY;
#pragma unroll
                for (int i = 0; i < 16; i++)
                {
                    mn = fminf(mn, s1[i].x);
                    mx = fmaxf(mx, s1[i].y);
                }
                hiz_store<PEERS>(h3 + by * hw3 + bx, make_float2(mn, mx), hp);
            }
        }
    }
    if (!PEERS)
    {
        tl_end(bp, kTlHizEnd);
        return;
    }
    // every tile is out (system-scope fence per thread); the last CTA raises this rank's word in every arena
    __threadfence_system();
    __shared__ int sLast;
    __syncthreads();
    if (t == 0)
        sLast = atomicAdd(hp.hdr + 129, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!sLast)
        return;
    __threadfence_system();
    if (t == 0)
        hp.hdr[129] = 0u;
    if (t < hp.world)
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(hp.flag[t]), "r"(hp.step) : "memory");
    if (t == 0)
        reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(hp.hdr) + 2048)[9 + 2 * hp.stamp_slot] = global_timer_ns();
}

// Emit the unit of brick b of chunk (x, y, z): `active` lanes hold one brick each. Free-space frames can only carve: the brick must
// hold an observed voxel -- its flag says so, or an EARLIER band frame of this batch may create one (the brick kernels check
// the actual register state before they spend a frame on it). Units are ordered by COST (frames to apply), in four buckets:
// the brick kernels hand tasks out bucket by bucket, so the long tasks start first and the shortest ones fill the tail.
__device__ __forceinline__ int unit_bucket(int frames, int K) { return 4 * frames > 3 * K ? 0 : (2 * frames > K ? 1 : (4 * frames > K ? 2 : 3)); }

// The free-space frames of a brick that are worth a visit, and whether the brick becomes a unit at all.
__device__ __forceinline__ bool brick_unit_masks(bool active, bool exists, bool virt, bool carve, unsigned long long flags, int b, unsigned band, unsigned freeFrames,
                                                 unsigned *fmOut)
{
    const unsigned afterBand = band ? ~((band & (0u - band)) | ((band & (0u - band)) - 1u)) : 0u;
    unsigned fm = carve ? freeFrames : 0u;
    if (!(exists && ((flags >> b) & 1ull)))
        fm &= afterBand;
    *fmOut = fm;
    return active && ((exists && (band | fm) != 0u) || (virt && band != 0u));
}

__device__ __forceinline__ void emit_brick_unit(const BatchParams &bp, bool active, unsigned long long key, int slot, bool exists, bool virt, bool carve,
                                                unsigned long long flags, int b, unsigned band, unsigned freeFrames, int K, unsigned lane)
{
    unsigned fm;
    const bool keepB = brick_unit_masks(active, exists, virt, carve, flags, b, band, freeFrames, &fm);
    const int bucket = keepB ? unit_bucket(__popc(band | fm), K) : -1;
    unsigned bm[4];
#pragma unroll
    for (int q = 0; q < 4; q++)
        bm[q] = __ballot_sync(0xffffffffu, bucket == q);
    int base = 0;
    const unsigned laneMask = lane == 0 ? bm[0] : (lane == 1 ? bm[1] : (lane == 2 ? bm[2] : bm[3]));     // lanes 0..3 reserve for buckets 0..3
    if (lane < 4 && laneMask)
    {
        int *ctr = lane == 0 ? &bp.bctr->unit_count : (lane == 1 ? &bp.bctr->bucket1 : (lane == 2 ? &bp.bctr->bucket2 : &bp.bctr->light_count));
        base = atomicAdd(ctr, __popc(laneMask));
    }
    const int myBase = __shfl_sync(0xffffffffu, base, bucket & 3);
    if (keepB)
    {
        const unsigned mine = bucket == 0 ? bm[0] : (bucket == 1 ? bm[1] : (bucket == 2 ? bm[2] : bm[3]));
        const int rank = myBase + __popc(mine & ((1u << lane) - 1));
        // buckets 0 / 2 grow from the front of their buffer, 1 / 3 from the back
        const int pos = ((bucket & 1) ? bp.units_cap - 1 - rank : rank) + (bucket >> 1) * bp.units_cap;
        bp.units[pos] = make_int4((int)(unsigned)(key & 0xffffffffull), (int)(unsigned)(key >> 32), (exists ? slot : kVirtualSlot) | (b << 24),
                                  (int)(band | (fm << 16)));
    }
}

// Unit number u of the batch in cost order (bucket 0 first). nb[q]: units in bucket q.
struct UnitCounts
{
    int n0, n01, n012, total;
};
__device__ __forceinline__ UnitCounts unit_counts(const BatchParams &bp)
{
    UnitCounts c;
    c.n0 = bp.bctr->unit_count;
    c.n01 = c.n0 + bp.bctr->bucket1;
    c.n012 = c.n01 + bp.bctr->bucket2;
    c.total = c.n012 + bp.bctr->light_count;
    return c;
}
__device__ __forceinline__ int unit_position(const BatchParams &bp, const UnitCounts &c, int u)
{
    if (u < c.n0)
        return u;
    if (u < c.n01)
        return bp.units_cap - 1 - (u - c.n0);
    if (u < c.n012)
        return bp.units_cap + (u - c.n01);
    return 2 * bp.units_cap - 1 - (u - c.n012);
}

// ------------------------------------------------------------------------------------------------------
// batch_candidates_kernel: which bricks of which chunks can change in which frames of the batch.
// A warp per tile of the union candidate box:
//   0. the chunk indices of the tile that THIS RANK OWNS are compacted by ballot (8 x world indices are looked at), so that at
//      N ranks a warp still works on full tiles and the kernel's work divides by N;
//   1. chunk level. (a) four lanes per chunk, a quarter of the frames each, cheap tests only: is the chunk inside frame f's candidate ID box and
//      does it pass Frustum::Intersects (ChunkManager.cpp:182-212, exact)? Is it inside the frame's view pyramid at all? The
//      (chunk, frame) pairs that are go into a queue in shared memory (ballot compaction). (b) a lane per QUEUED PAIR: the
//      chunk against the frame's Hi-Z tiles. Most pairs end here: behind the surface or in free space with nothing to carve.
//      One hash lookup per chunk that is left (all lanes in parallel);
//   2. brick level, the whole warp per surviving chunk: lanes = (brick, one of four frames) -> per-brick frame masks
//        band   frames in which some voxel of the brick may fall inside the truncation band
//        free   frames in which the brick lies in free space (only carving of observed voxels can act)
//      A free-space frame is kept when the brick can hold a carvable voxel by then: its flag is set already, or an EARLIER band
//      frame of this batch may create one (the brick kernels check the actual register state before spending the frame);
//   3. warp-ballot compaction into the unit lists (four cost buckets).
// Chunk indices go through a multiplicative permutation so that the surviving chunks spread over all warps.
// Lanes of a pass = (chunk lane & 7, frame group lane >> 3): EIGHT chunks per warp pass, each lane walks a quarter of the frames.
// (32 chunks per pass, a lane per chunk, left 7 warps per SM on a 35 K chunk union box: the kernel ran at the latency of one
// warp's chain. Four times the warps, a quarter of the chain each.)
constexpr int kCandWarps = 4;                       // warps per CTA
constexpr int kCandChunks = 8;                      // chunks per warp pass
constexpr int kCandList = kCandChunks * 8;          // owned chunk indices of one tile (8 x min(world, 8) indices are looked at)

#ifndef CHS_CAND_MIN_CTAS
#define CHS_CAND_MIN_CTAS 8                  // 64 registers: 32 warps per SM, so that a 35 K chunk union box is ONE wave of warps
#endif
template <int CS>
__global__ void __launch_bounds__(32 * kCandWarps, CHS_CAND_MIN_CTAS) batch_candidates_kernel(BatchParams bp, DeviceMap map)
{
    constexpr int BPA = CS / 8, NB = BPA * BPA * BPA;
    __shared__ __align__(16) FrameParams sF[kMaxBatch];
    __shared__ float2 sCoarse[kMaxBatch][kCoarseTiles];
    __shared__ int sList[kCandWarps][kCandList];
    __shared__ unsigned short sPairs[kCandWarps][kCandChunks * kMaxBatch];     // (chunk of the pass << 8) | frame: pairs that need the depth test
    __shared__ unsigned sBandC[kCandWarps][kCandChunks], sFreeC[kCandWarps][kCandChunks];
    // units of a pass (8 chunks x 8 bricks at most), written to the global lists in ONE go: four atomics per pass instead of four
    // per surviving chunk (the warp waits for their round trip)
    __shared__ int4 sUnits[kCandWarps][NB == 8 ? kCandChunks * 8 : 1];
    pdl_launch_dependents();                       // the brick kernel may start launching: it waits (pdl_wait) before it reads the unit lists
    load_frames(sF, bp);                           // the frame table was uploaded before the Hi-Z kernel: complete
    pdl_wait();                                    // the Hi-Z kernel's output is read from here on
    tl_start(bp, kTlCandStart);
    if (bp.coarse_in_shared)
    {
        // Hi-Z levels >= 4 (tiles of 128 pixels and up, until at most 3x3 tiles cover the image) of every frame, from level 3:
        // a few dozen tiles per frame, rebuilt by every CTA in shared memory instead of by the last block of the Hi-Z kernel
        const int K = bp.K;
        int off = 0;
        for (int l = 4; l < sF[0].hiz_levels; l++)
        {
            const int w = sF[0].hizW[l], h = sF[0].hizH[l], pw = sF[0].hizW[l - 1], ph = sF[0].hizH[l - 1];
            for (int i = threadIdx.x; i < K * w * h; i += blockDim.x)
            {
                const int f = i / (w * h), r = i - f * (w * h), ox = r % w, oy = r / w;
                const float2 *prev = (l == 4) ? sF[f].hiz[3] : &sCoarse[f][off - pw * ph];
                float mn = INFINITY, mx = -INFINITY;
                for (int dy = 0; dy < 2; dy++)
                    for (int dx = 0; dx < 2; dx++)
                    {
                        const int qx = ox * 2 + dx, qy = oy * 2 + dy;
                        if (qx < pw && qy < ph)
                        {
                            const float2 v = prev[qy * pw + qx];
                            mn = fminf(mn, v.x);
                            mx = fmaxf(mx, v.y);
                        }
                    }
                sCoarse[f][off + r] = make_float2(mn, mx);
            }
            __syncthreads();
            off += w * h;
        }
        // point the frames' coarse levels at the shared copies (generic addresses; classify_box<true> reads them with plain loads)
        if (threadIdx.x < K)
        {
            int o = 0;
            for (int l = 4; l < sF[0].hiz_levels; l++)
            {
                sF[threadIdx.x].hiz[l] = &sCoarse[threadIdx.x][o];
                o += sF[0].hizW[l] * sF[0].hizH[l];
            }
        }
        __syncthreads();
    }
    // chunks created by this batch will occupy the pool slots from here on (this kernel runs on the map's stream, after every
    // earlier batch; the prepare kernels may run ahead on the copy stream)
    if (blockIdx.x == 0 && threadIdx.x == 0)
        bp.bctr->chunks_at_start = map.ctr->n_chunks;
    const int K = bp.K;
    const int total = bp.n[0] * bp.n[1] * bp.n[2];
    const int nyz = bp.n[1] * bp.n[2];
    const unsigned lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const unsigned below = (1u << lane) - 1u;
    const int world = map.world > 1 ? map.world : 1;
    const int tileSize = kCandChunks * (world < 8 ? world : 8);       // indices per tile: about 8 of them are owned
    const int rounds = (tileSize + 31) / 32;
    const int cl = (int)lane & (kCandChunks - 1), fg = (int)lane / kCandChunks;      // chunk of the pass, frame group
    constexpr int kGroups = 32 / kCandChunks;
    const int nTiles = (total + tileSize - 1) / tileSize;
    const bool carve = sF[0].carve != 0;
    const float ext = __fmul_rn((float)CS, map.res);
    int myCount = 0;                                                   // lane f: candidates of frame f seen by this warp
    for (int tile = blockIdx.x * kCandWarps + warp; tile < nTiles; tile += gridDim.x * kCandWarps)
    {
        // 0. owned chunk indices of the tile
        int nList = 0;
        for (int r = 0; r < rounds; r++)
        {
            const int inTile = r * 32 + (int)lane;
            const int slotIdx = tile * tileSize + inTile;
            bool own = false;
            int i = 0;
            if (inTile < tileSize && slotIdx < total)
            {
                i = (int)(((long long)slotIdx * bp.cand_stride) % total);
                own = true;
                if (world > 1)
                {
                    const int x = bp.lo[0] + i / nyz, rr = i - (i / nyz) * nyz;
                    own = (owner_hash(x, bp.lo[1] + rr / bp.n[2], bp.lo[2] + rr % bp.n[2]) % (unsigned)world) == (unsigned)map.rank;
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, own);
            if (own)
                sList[warp][nList + __popc(m & below)] = i;
            nList += __popc(m);
        }
        __syncwarp();
        for (int base = 0; base < nList; base += kCandChunks)
        {
            const bool have = base + cl < nList;
            int x = 0, y = 0, z = 0;
            float bx = 0.0f, by = 0.0f, bz = 0.0f;
            unsigned candM = 0u;
            if (have)
            {
                const int i = sList[warp][base + cl];
                x = bp.lo[0] + i / nyz;
                const int r = i - (i / nyz) * nyz;
                y = bp.lo[1] + r / bp.n[2];
                z = bp.lo[2] + r % bp.n[2];
                // chunk box exactly as ChunkManager.cpp:199-201
                bx = __fmul_rn((float)(x * CS), map.res);
                by = __fmul_rn((float)(y * CS), map.res);
                bz = __fmul_rn((float)(z * CS), map.res);
            }
            if (lane < kCandChunks)
            {
                sBandC[warp][lane] = 0u;
                sFreeC[warp][lane] = 0u;
            }
            // 1a. chunk level, cheap part, four lanes per chunk with a quarter of the frames each: candidate of the frame (exact)?
            // inside its view pyramid? The (chunk, frame) pairs that are go into the warp's queue.
            int qn = 0;
            {
                const float ex = __fadd_rn(bx, ext), ey = __fadd_rn(by, ext), ez = __fadd_rn(bz, ext);
                for (int f0 = 0; f0 < K; f0 += kGroups)
                {
                    const int f = f0 + fg;
                    const FrameParams &fp = sF[f < K ? f : 0];
                    bool vis = false;
                    if (have && f < K && (unsigned)(x - fp.lo[0]) < (unsigned)fp.n[0] && (unsigned)(y - fp.lo[1]) < (unsigned)fp.n[1] &&
                        (unsigned)(z - fp.lo[2]) < (unsigned)fp.n[2] && frustum_intersects_exact(fp, bx, by, bz, ex, ey, ez))
                    {
                        candM |= 1u << f;
                        vis = !box_outside_view(fp, bx + map.half, by + map.half, bz + map.half, (float)(CS - 1) * map.res);
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, vis);
                    if (vis)
                        sPairs[warp][qn + __popc(m & below)] = (unsigned short)(((unsigned)cl << 8) | (unsigned)f);
                    qn += __popc(m);
                }
            }
            __syncwarp();
            // 1b. the depth test of the queued pairs, a lane per pair (dense)
            for (int q0 = 0; q0 < qn; q0 += 32)
            {
                if (q0 + (int)lane < qn)
                {
                    const unsigned pr = sPairs[warp][q0 + lane];
                    const int pc = (int)(pr >> 8), f = (int)(pr & 255u);
                    const int i = sList[warp][base + pc];
                    const int px = bp.lo[0] + i / nyz, r = i - (i / nyz) * nyz, py = bp.lo[1] + r / bp.n[2], pz = bp.lo[2] + r % bp.n[2];
                    const int code = classify_box_depth<true>(sF[f], __fmul_rn((float)(px * CS), map.res) + map.half, __fmul_rn((float)(py * CS), map.res) + map.half,
                                                              __fmul_rn((float)(pz * CS), map.res) + map.half, (float)(CS - 1) * map.res);
                    if (code == 2)
                        atomicOr(&sBandC[warp][pc], 1u << f);
                    else if (code == 1)
                        atomicOr(&sFreeC[warp][pc], 1u << f);
                }
            }
            __syncwarp();
            const unsigned chunkBand = sBandC[warp][cl], chunkFree = sFreeC[warp][cl];
            __syncwarp();
            // per-frame candidate counts: lane f of the warp accumulates frame f
            for (int f = 0; f < K; f++)
            {
                const unsigned m = __ballot_sync(0xffffffffu, (candM >> f) & 1u);
                if ((int)lane == f)
                    myCount += __popc(m);
            }
            // Does the chunk exist? One lookup per chunk that stage 1 could not dismiss: most of the view pyramid is free space
            // without chunks, and a chunk that does not exist only matters if some frame may hit it.
            int slot = -1;
            unsigned long long flags = 0ull;
            if (fg == 0 && (chunkBand || (chunkFree && carve)))
            {
                slot = hash_lookup(map, pack_id(x, y, z));
                // the flags also matter when the chunk as a whole is in the band of a frame: single bricks of it may still lie in
                // free space and hold carvable voxels from before
                if (slot >= 0 && carve)
                    flags = map.brick_flags[slot];
            }
            // free-space frames can only carve observed voxels: they matter for an existing chunk with a carvable brick, or
            // after a band frame of this batch (which may create one)
            const unsigned freeTodo = (carve && (chunkBand || (slot >= 0 && flags != 0ull))) ? chunkFree : 0u;
            // 2. brick level: the warp takes the surviving chunks one at a time
            unsigned surv = __ballot_sync(0xffffffffu, fg == 0 && (chunkBand | freeTodo) != 0u);
            int nUnits = 0;
            while (surv)
            {
                const int src = __ffs(surv) - 1;
                surv &= surv - 1;
                const int sx = __shfl_sync(0xffffffffu, x, src), sy = __shfl_sync(0xffffffffu, y, src), sz = __shfl_sync(0xffffffffu, z, src);
                const float cbx = __shfl_sync(0xffffffffu, bx, src), cby = __shfl_sync(0xffffffffu, by, src), cbz = __shfl_sync(0xffffffffu, bz, src);
                const unsigned sBand = __shfl_sync(0xffffffffu, chunkBand, src), sFree = __shfl_sync(0xffffffffu, freeTodo, src);
                const int sSlot = __shfl_sync(0xffffffffu, slot, src);
                const unsigned long long sFlags = __shfl_sync(0xffffffffu, flags, src);
                const unsigned long long key = pack_id(sx, sy, sz);
                const bool exists = sSlot >= 0;
                auto brick_code = [&](int b, int f) -> int
                {
                    const int qx = b % BPA, qy = (b / BPA) % BPA, qz = b / (BPA * BPA);
                    return classify_box<true>(sF[f], cbx + (float)(qx * 8) * map.res + map.half, cby + (float)(qy * 8) * map.res + map.half,
                                              cbz + (float)(qz * 8) * map.res + map.half, 7.0f * map.res);
                };
                if (NB == 1)
                {
                    if (lane == 0 && !exists && sBand)
                        atomicAdd(&bp.bctr->new_count, 1);
                    emit_brick_unit(bp, lane == 0, key, sSlot, exists, !exists && sBand != 0u, carve, sFlags, 0, sBand, sFree, K, lane);
                }
                else if (NB == 8)
                {
                    // lanes = (brick lane & 7, frame slot lane >> 3): the four lowest frames of the to-do mask per round
                    const int b = lane & 7, j = lane >> 3;
                    unsigned bandB = 0u, freeB = 0u;
                    unsigned t = sBand | sFree;
                    while (t)
                    {
                        unsigned pick = 0u;
#pragma unroll
                        for (int q = 0; q < 4; q++)
                        {
                            const unsigned low = t & (0u - t);
                            t ^= low;
                            pick = (q == j) ? low : pick;
                        }
                        if (pick)
                        {
                            const int code = brick_code(b, __ffs(pick) - 1);
                            bandB |= code == 2 ? pick : 0u;
                            freeB |= code == 1 ? pick : 0u;
                        }
                    }
                    bandB |= __shfl_xor_sync(0xffffffffu, bandB, 8);
                    freeB |= __shfl_xor_sync(0xffffffffu, freeB, 8);
                    bandB |= __shfl_xor_sync(0xffffffffu, bandB, 16);
                    freeB |= __shfl_xor_sync(0xffffffffu, freeB, 16);
                    // the chunk-level band mask that decides creation is the union of the brick-level ones (tighter than stage 1)
                    unsigned u = bandB;
                    u |= __shfl_xor_sync(0xffffffffu, u, 1);
                    u |= __shfl_xor_sync(0xffffffffu, u, 2);
                    u |= __shfl_xor_sync(0xffffffffu, u, 4);
                    const bool virt = !exists && u != 0u;
                    if (lane == 0 && virt)
                        atomicAdd(&bp.bctr->new_count, 1);
                    unsigned fm;
                    const bool keepB = brick_unit_masks(lane < 8, exists, virt, carve, sFlags, b, bandB, freeB, &fm);
                    const unsigned km = __ballot_sync(0xffffffffu, keepB);
                    if (keepB)
                        sUnits[warp][nUnits + __popc(km & below)] = make_int4((int)(unsigned)(key & 0xffffffffull), (int)(unsigned)(key >> 32),
                                                                              (exists ? sSlot : kVirtualSlot) | (b << 24), (int)(bandB | (fm << 16)));
                    nUnits += __popc(km);
                }
                else
                {
                    // 64 bricks: two per lane, frames in a loop
                    unsigned bandB[2] = {0u, 0u}, freeB[2] = {0u, 0u};
#pragma unroll
                    for (int hb = 0; hb < 2; hb++)
                    {
                        unsigned t = sBand | sFree;
                        while (t)
                        {
                            const unsigned low = t & (0u - t);
                            t ^= low;
                            const int code = brick_code((int)lane + 32 * hb, __ffs(low) - 1);
                            bandB[hb] |= code == 2 ? low : 0u;
                            freeB[hb] |= code == 1 ? low : 0u;
                        }
                    }
                    unsigned u = bandB[0] | bandB[1];
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1)
                        u |= __shfl_xor_sync(0xffffffffu, u, o);
                    const bool virt = !exists && u != 0u;
                    if (lane == 0 && virt)
                        atomicAdd(&bp.bctr->new_count, 1);
#pragma unroll
                    for (int hb = 0; hb < 2; hb++)
                        emit_brick_unit(bp, true, key, sSlot, exists, virt, carve, sFlags, (int)lane + 32 * hb, bandB[hb], freeB[hb], K, lane);
                }
            }
            if (NB == 8 && nUnits > 0)
            {
                // one reservation per cost bucket for the whole pass, then every unit goes to its place
                __syncwarp();
                int cnt[4] = {0, 0, 0, 0};
                for (int q = 0; q < nUnits; q += 32)
                {
                    const bool in = q + (int)lane < nUnits;
                    const int w = in ? sUnits[warp][q + lane].w : 0;
                    const int bucket = in ? unit_bucket(__popc((unsigned)w & 0xFFFFu | ((unsigned)w >> 16)), K) : -1;
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        cnt[k] += __popc(__ballot_sync(0xffffffffu, bucket == k));
                }
                int base = 0;
                const int mineCnt = lane == 0 ? cnt[0] : (lane == 1 ? cnt[1] : (lane == 2 ? cnt[2] : cnt[3]));
                if (lane < 4 && mineCnt)
                {
                    int *ctr = lane == 0 ? &bp.bctr->unit_count : (lane == 1 ? &bp.bctr->bucket1 : (lane == 2 ? &bp.bctr->bucket2 : &bp.bctr->light_count));
                    base = atomicAdd(ctr, mineCnt);
                }
                int run[4];
#pragma unroll
                for (int k = 0; k < 4; k++)
                    run[k] = __shfl_sync(0xffffffffu, base, k);
                for (int q = 0; q < nUnits; q += 32)
                {
                    const bool in = q + (int)lane < nUnits;
                    const int4 un = in ? sUnits[warp][q + lane] : make_int4(0, 0, 0, 0);
                    const int bucket = in ? unit_bucket(__popc((unsigned)un.w & 0xFFFFu | ((unsigned)un.w >> 16)), K) : -1;
#pragma unroll
                    for (int k = 0; k < 4; k++)
                    {
                        const unsigned bmk = __ballot_sync(0xffffffffu, bucket == k);
                        if (bucket == k)
                        {
                            const int rank = run[k] + __popc(bmk & below);
                            // buckets 0 / 2 grow from the front of their buffer, 1 / 3 from the back
                            bp.units[((k & 1) ? bp.units_cap - 1 - rank : rank) + (k >> 1) * bp.units_cap] = un;
                        }
                        run[k] += __popc(bmk);
                    }
                }
                __syncwarp();
            }
        }
        __syncwarp();
    }
    if ((int)lane < K && myCount)
        atomicAdd(&bp.bctr->candidates[lane], myCount);
    tl_end(bp, kTlCandEnd);
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

// Distributed batches: the CTAs that land on every 16th SM take no tasks, so that those SMs stay free for the NCCL kernels of the
// NEXT batch's frame exchange (a persistent grid that fills every SM would keep them queued until it drains).
__device__ __forceinline__ bool on_reserved_sm(const BatchParams &bp)
{
    unsigned smid;
    asm("mov.u32 %0, %%smid;" : "=r"(smid));
    return bp.reserve_sms && (smid & 15u) == 0u;
}

__device__ __forceinline__ void batch_span_start(const BatchParams &bp)
{
    if (threadIdx.x == 0)
        atomicMax(&bp.bctr->span_start_inv, ~global_timer_ns());
}

__device__ __forceinline__ void batch_flush(const BatchParams &bp, BatchShared *s)
{
    __syncthreads();
    const int t = threadIdx.x;
    if (t == 0)
        atomicMax(&bp.bctr->span_end, global_timer_ns());
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
    if (threadIdx.x < bp.peer_world)
    {
        // every CTA has finished (ticket): nothing of this step's images will be read again
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(bp.peer_done[threadIdx.x]), "r"(bp.peer_step) : "memory");
    }
    const volatile BatchCounters *c = bp.bctr;
    const volatile Counters *g = map.ctr;
    volatile HostBatchSnapshot *h = bp.host_slot;
    const int t = threadIdx.x;
    // chunks created by this batch occupy the slots [chunks_at_start, n_chunks); a chunk is born in the first frame that
    // updated it (the lowest bit of its frame mask)
    {
        __shared__ int sNew[kMaxBatch];
        if (t < kMaxBatch)
            sNew[t] = 0;
        __syncthreads();
        const int n0 = c->chunks_at_start, n1 = min(g->n_chunks, map.capacity);
        for (int s = n0 + t; s < n1; s += blockDim.x)
        {
            const unsigned m = (unsigned)(*reinterpret_cast<volatile unsigned long long *>(&bp.slot_batch[s]) & 0xFFFFull);
            if (m)
                atomicAdd(&sNew[__ffs(m) - 1], 1);
        }
        __syncthreads();
        if (t < kMaxBatch)
            bp.bctr->n_new[t] = sNew[t];
        __syncthreads();
    }
    // the snapshot is assembled in shared memory (so that its checksum is computed from exactly what is sent), then copied out
    __shared__ HostBatchSnapshot sSnap;
    if (t == 0)
    {
        sSnap.head = bp.batch_id;
        sSnap.n_chunks = g->n_chunks;
        sSnap.n_dirty = g->n_dirty;
        sSnap.error_flags = g->error_flags;
        sSnap.unit_count = c->unit_count + c->bucket1 + c->bucket2 + c->light_count;
        sSnap.new_count = c->new_count;
        sSnap.K = bp.K;
        sSnap.pad = 0;
        sSnap.bricks_span_ns = (long long)(c->span_end - ~c->span_start_inv);
        for (int i = 0; i < kTimelineStamps; i++)
            sSnap.timeline[i] = 0;
        if (bp.timeline)
        {
            // starts are kept as max of ~time; the words are zeroed for the next batch that uses this staging set
            volatile unsigned long long *tl = bp.timeline;
            for (int i = 0; i < kTlBricksStart; i++)
            {
                const unsigned long long v = tl[i];
                const bool isStart = i == kTlPushStart || i == kTlHizStart || i == kTlCandStart;
                sSnap.timeline[i] = v ? (long long)(isStart ? ~v : v) : 0;
                tl[i] = 0ull;
            }
            sSnap.timeline[kTlBricksStart] = (long long)~c->span_start_inv;
            sSnap.timeline[kTlBricksEnd] = (long long)c->span_end;
        }
        sSnap.tail = bp.batch_id;
        sSnap.pad3[0] = sSnap.pad3[1] = sSnap.pad3[2] = 0;
    }
    if (t < kMaxBatch)
    {
        sSnap.candidates[t] = c->candidates[t];
        sSnap.n_new[t] = c->n_new[t];
        sSnap.updated_chunks[t] = c->updated_chunks[t];
        sSnap.n_upd[t] = (long long)c->n_upd[t];
        sSnap.n_carve[t] = (long long)c->n_carve[t];
        sSnap.n_col[t] = (long long)c->n_col[t];
    }
    __syncthreads();
    if (t == 0)
        sSnap.checksum = batch_snapshot_checksum(sSnap);
    __syncthreads();
    static_assert(sizeof(HostBatchSnapshot) % 4 == 0, "copied word-wise");
    const int *src = reinterpret_cast<const int *>(&sSnap);
    volatile int *dst = reinterpret_cast<volatile int *>(h);
    for (int i = t; i < (int)(sizeof(HostBatchSnapshot) / 4); i += blockDim.x)
        dst[i] = src[i];
    // the counters of this staging set start from zero in the batch after next (nobody else is running: this is the last CTA)
    __syncthreads();
    int *cz = reinterpret_cast<int *>(bp.bctr);
    for (int i = t; i < (int)(sizeof(BatchCounters) / 4); i += blockDim.x)
        cz[i] = 0;
}

// ------------------------------------------------------------------------------------------------------
// One frame applied to the lane's kVPL voxels of a brick part (z slices kNS*half .. kNS*half+kNS-1, y rows ly and ly+4), state in
// registers. Same arithmetic as process_batch<MODE 0> (ProjectionIntegrator.h:51-183); colour and depth cameras coincide.
//
// FAST = true: straight-line code. The reciprocal of the projection and the quotient of DistVoxel::Integrate use the
// range-check-free correctly rounded forms (integrate_device.cuh), every update is computed for all kVPL voxels and selected
// by predicate, so their dependency chains interleave. The caller's warp votes guarantee the operand ranges; if a vote
// fails, nothing has been modified yet and the frame is redone with FAST = false (__frcp_rn / __fdiv_rn, branches).
// Returns false (FAST only) when an operand is out of range.
template <int CS, bool COLOR_PATH, bool PER_PIXEL, bool FAST>
__device__ __forceinline__ bool frame_on_half_brick(const FrameParams &fp, const DeviceMap &map, const BrickLane &L, int half, bool hasCol,
                                                    float2 (&dv)[kVPL], unsigned (&cv)[kVPL], unsigned &wroteD, unsigned &wroteC,
                                                    int &nUpd, int &nCarve, int &nCol, bool &carvable)
{
    const CameraDev &c = fp.cam;
    int pix[kVPL];
    float cz[kVPL];
    bool ok = true;
#pragma unroll
    for (int s = 0; s < kNS; s++)
    {
        const float pz = __fadd_rn(__fadd_rn(__fmul_rn((float)(L.vz0 + kNS * half + s), map.res), map.half), L.orgz);
        const float d2 = __fsub_rn(pz, c.t[2]);
        const float m20 = __fmul_rn(c.R[6], d2), m21 = __fmul_rn(c.R[7], d2), m22 = __fmul_rn(c.R[8], d2);
#pragma unroll
        for (int h = 0; h < 2; h++)
        {
            const int k = s * 2 + h;
            const float cx = __fadd_rn(L.m0[0], __fadd_rn(L.m1[h][0], m20));
            const float cy = __fadd_rn(L.m0[1], __fadd_rn(L.m1[h][1], m21));
            cz[k] = __fadd_rn(L.m0[2], __fadd_rn(L.m1[h][2], m22));
            if (FAST)
                ok &= rcp_in_range(cz[k]);
            const float invZ = FAST ? rcp_rn_inrange(cz[k]) : __frcp_rn(cz[k]);                // == 1.0f / z, correctly rounded
            const float u = __fadd_rn(__fmul_rn(__fmul_rn(c.fx, cx), invZ), c.cx);
            const float v = __fadd_rn(__fmul_rn(__fmul_rn(c.fy, cy), invZ), c.cy);
            // bitwise, not short-circuit: five compares feeding one predicate instead of four branches
            const bool on = (u >= 0.0f) & (v >= 0.0f) & (u < c.Wf) & (v < c.Hf) & !(cz[k] < 0.0f);
            pix[k] = on ? (int)u + (int)v * c.W : -1;
        }
    }
    if (FAST && !__all_sync(0xffffffffu, ok))
        return false;
    float depth[kVPL], trunc[kVPL];
    unsigned cpx[kVPL];
#pragma unroll
    for (int k = 0; k < kVPL; k++)
    {
        depth[k] = pix[k] >= 0 ? __ldg(fp.depth + pix[k]) : __int_as_float(0x7fc00000);
        // the colour of a voxel is frozen once its colour weight reaches 8 (:153): no fetch for those
        cpx[k] = (COLOR_PATH && hasCol && pix[k] >= 0 && (cv[k] >> 24) < 8u) ? __ldg(fp.color_packed + pix[k]) : 0u;
        trunc[k] = PER_PIXEL ? (pix[k] >= 0 ? __ldg(fp.trunc_img + pix[k]) : 0.0f) : fp.trunc_param;
    }
    if constexpr (FAST)
    {
        // predicates and the operands of the division for all kVPL voxels
        unsigned band = 0u, crv = 0u;
        float sd[kVPL], num[kVPL], den[kVPL], wu[kVPL];
#pragma unroll
        for (int k = 0; k < kVPL; k++)
        {
            const float d = depth[k];
            // !(d <= cutoff) is "NaN or beyond the cutoff" (:134, :141); the depth path keeps NaN (its comparisons then all fail)
            const bool skip = (pix[k] < 0) | (COLOR_PATH ? !(d <= 100.0f) : (d > 50.0f));
            sd[k] = __fsub_rn(d, cz[k]);
            const bool inBand = !skip & (fabsf(sd[k]) < __fadd_rn(trunc[k], fp.diag));                      // :82 / :143
            const bool canCarve = !skip & !inBand & (fp.carve != 0) & (sd[k] > __fadd_rn(trunc[k], fp.carve_dist))  // :88 / :166
                                  & (dv[k].y > 0.0f) & (dv[k].x < fp.sdf_carve_max);                         // :90 / :169
            band |= (inBand ? 1u : 0u) << k;
            crv |= (canCarve ? 1u : 0u) << k;
            wu[k] = COLOR_PATH ? (PER_PIXEL ? fp.weight : fp.wu_const) : 1.0f;
            if (COLOR_PATH && PER_PIXEL)
            {
                const float t5 = __fmul_rn(5.0f, trunc[k]);                                                  // ConstantWeighter.h:43-46
                ok &= !inBand | div_in_range(fp.weight, t5);
                wu[k] = div_rn_inrange(fp.weight, inBand ? t5 : 1.0f);
            }
            num[k] = inBand ? __fadd_rn(__fmul_rn(dv[k].y, dv[k].x), __fmul_rn(wu[k], sd[k])) : 0.0f;      // DistVoxel.h:52-60
            den[k] = inBand ? __fadd_rn(wu[k], dv[k].y) : 1.0f;
            ok &= div_in_range(num[k], den[k]);
        }
        if (!__all_sync(0xffffffffu, ok))
            return false;
        if (!__any_sync(0xffffffffu, (band | crv) != 0u))
            return true;
        unsigned colM = 0u;
        if (COLOR_PATH && hasCol)
        {
#pragma unroll
            for (int k = 0; k < kVPL; k++)
                colM |= ((((band >> k) & 1u) && (cv[k] >> 24) < 8u) ? 1u : 0u) << k;                         // :153
        }
        const bool anyCol = COLOR_PATH && __any_sync(0xffffffffu, colM != 0u);
#pragma unroll
        for (int k = 0; k < kVPL; k++)
        {
            const bool inBand = (band >> k) & 1u, canCarve = (crv >> k) & 1u;
            const float q = div_rn_inrange(num[k], den[k]);
            const float2 upd = make_float2(q, __fadd_rn(dv[k].y, wu[k]));
            const float2 carved = (COLOR_PATH && !(dv[k].y < 5.0f)) ? make_float2(dv[k].x, __fsub_rn(dv[k].y, 1.0f))   // :171-175
                                                                      : make_float2(99999.0f, 0.0f);                    // DistVoxel::Carve -> Reset
            if (COLOR_PATH && anyCol)
            {
                const bool colOk = (colM >> k) & 1u;
                const unsigned w = min(cv[k] >> 24, 7u);
                const unsigned mrec = cRecip20[w + 1];
                const unsigned nr = ((w * (cv[k] & 0xFFu) + (cpx[k] & 0xFFu)) * mrec) >> 20;
                const unsigned ng = ((w * ((cv[k] >> 8) & 0xFFu) + ((cpx[k] >> 8) & 0xFFu)) * mrec) >> 20;
                const unsigned nb = ((w * ((cv[k] >> 16) & 0xFFu) + ((cpx[k] >> 16) & 0xFFu)) * mrec) >> 20;
                cv[k] = colOk ? (nr | (ng << 8) | (nb << 16) | ((w + 1) << 24)) : cv[k];
                wroteC |= (colOk ? 1u : 0u) << k;
                nCol += colOk;
            }
            dv[k].x = inBand ? upd.x : (canCarve ? carved.x : dv[k].x);
            dv[k].y = inBand ? upd.y : (canCarve ? carved.y : dv[k].y);
            carvable |= (inBand | canCarve) & (dv[k].y > 0.0f) & (dv[k].x < fp.sdf_carve_max);
        }
        wroteD |= band | crv;
        nUpd += __popc(band);
        nCarve += __popc(crv);
        return true;
    }
    if constexpr (!FAST)
#pragma unroll
    for (int k = 0; k < kVPL; k++)
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
    return true;
}

// Pool slot of chunk `key`, creating the chunk if it does not exist (lane 0 of a warp whose brick part of a VIRTUAL unit was
// hit). Several warps of one new chunk race: the one whose CAS claims the hash entry bumps the pool, fills the side arrays and
// publishes the slot; the others wait for the value (vals of empty entries hold -1). No voxel is written here: pool slots at
// and above n_chunks always hold the initial state {99999, 0} / colour 0 (capi.cu keeps that invariant), which is exactly
// what Chunk::Chunk produces (Chunk.cpp:33-48, DistVoxel.cpp:29-33).
__device__ __forceinline__ int get_or_create_chunk(const BatchParams &bp, const DeviceMap &map, unsigned long long key, int x, int y, int z)
{
    unsigned i = (unsigned)mix64(key) & map.mask;
    while (true)
    {
        unsigned long long k = *reinterpret_cast<volatile unsigned long long *>(&map.keys[i]);
        if (k == kEmptyKey)
        {
            k = atomicCAS(&map.keys[i], kEmptyKey, key);
            if (k == kEmptyKey)
            {
                int s = atomicAdd(&map.ctr->n_chunks, 1);
                if (s >= map.capacity)
                {
                    // pool exhausted (the host sizes it ahead of need, so this is a bug or an out-of-memory map): reported
                    // through error_flags. The key stays in the table with the value kNoSlot: no chunk, nothing is stored, and the
                    // waiters give up as well -- never alias another chunk's slot.
                    atomicOr(&map.ctr->error_flags, kErrPoolFull);
                    atomicSub(&map.ctr->n_chunks, 1);
                    s = kNoSlot;
                }
                else
                {
                    map.slot_ids[3 * s] = x;
                    map.slot_ids[3 * s + 1] = y;
                    map.slot_ids[3 * s + 2] = z;
                    map.brick_flags[s] = 0ull;
                    map.slot_epoch[s] = 0;
                    bp.slot_batch[s] = 0ull;
                }
                __threadfence();
                *reinterpret_cast<volatile int *>(&map.vals[i]) = s;
                return s;
            }
        }
        if (k == key)
        {
            int s;
            while ((s = *reinterpret_cast<volatile int *>(&map.vals[i])) == -1)
                __nanosleep(20);
            __threadfence();
            return s;
        }
        i = (i + 1) & map.mask;
    }
}

// A warp per brick part (kNS z slices); tasks are handed out by an atomic counter because their cost (1 .. K frames) varies. Units of
// chunks that do not exist yet start from the initial state; the chunk is created when (and only if) a frame hits
// ("created and untouched => garbage collected", Chisel.h:76-80,102-110 / :133-143,170-173,202-207, never allocates anything).
template <int CS, bool COLOR_PATH, bool PER_PIXEL>
__global__ void __launch_bounds__(CHS_BRICK_THREADS, CHS_BRICK_MIN_CTAS) batch_bricks_kernel(BatchParams bp, DeviceMap map)
{
    constexpr int BPA = CS / 8;
    __shared__ __align__(16) FrameParams sF[kMaxBatch];
    __shared__ BatchShared sB;
    batch_shared_zero(&sB);
    load_frames(sF, bp);
    pdl_wait();                                    // the candidates kernel's unit lists and counters are read from here on
    batch_span_start(bp);
    const int lane = threadIdx.x & 31;
    // the list holds every brick of the union box at most once, so heavy (front) and light (back) units cannot collide
    const UnitCounts uc = unit_counts(bp);
    const int nTasks = uc.total * kParts;
    const bool hasCol = COLOR_PATH && map.use_color;
    const float carveMax = sF[0].sdf_carve_max;
    while (!on_reserved_sm(bp))
    {
        int g = 0;
        if (lane == 0)
            g = atomicAdd(&bp.bctr->next_task, 1);
        g = __shfl_sync(0xffffffffu, g, 0);
        if (g >= nTasks)
            break;
        const int u = g / kParts;
        const int4 unit = bp.units[unit_position(bp, uc, u)];
        const int half = g % kParts;                     // which group of kNS z slices of the brick
        int x, y, z;
        const unsigned long long key = ((unsigned long long)(unsigned)unit.y << 32) | (unsigned long long)(unsigned)unit.x;
        unpack_id(key, &x, &y, &z);
        int slot = unit.z & 0xFFFFFF;
        const int b = unit.z >> 24;
        const bool virt = slot == kVirtualSlot;
        const unsigned bandM = (unsigned)unit.w & 0xFFFFu;
        unsigned mask = bandM | ((unsigned)unit.w >> 16);
        const float orgx = __fmul_rn((float)(CS * x), map.res), orgy = __fmul_rn((float)(CS * y), map.res), orgz = __fmul_rn((float)(CS * z), map.res);
        const int bx = b % BPA, by = (b / BPA) % BPA, bz = b / (BPA * BPA);
        const int idx0 = ((bz * 8 + kNS * half) * CS + (by * 8 + (lane >> 3))) * CS + bx * 8 + (lane & 7);
        float2 dv[kVPL];
        unsigned cv[kVPL];
        if (!virt)
        {
            const float2 *dist = dist_ptr(map, slot);
            const unsigned *col = hasCol ? reinterpret_cast<const unsigned *>(color_ptr(map, slot)) : nullptr;
#pragma unroll
            for (int k = 0; k < kVPL; k++)
            {
                const int idx = idx0 + (k >> 1) * CS * CS + (k & 1) * 4 * CS;
                dv[k] = dist[idx];
                cv[k] = hasCol ? col[idx] : 0u;
            }
        }
        else
        {
#pragma unroll
            for (int k = 0; k < kVPL; k++)
            {
                dv[k] = make_float2(99999.0f, 0.0f);                // Chunk::Chunk initial state (DistVoxel.cpp:29-33)
                cv[k] = 0u;
            }
        }
        unsigned wroteD = 0u, wroteC = 0u, updMask = 0u;
        bool carvable = false;
        while (mask)
        {
            const int f = __ffs(mask) - 1;
            mask &= mask - 1;
            const FrameParams &fp = sF[f];
            if (!((bandM >> f) & 1u))
            {
                // free-space frame: it can only carve, and only voxels with weight > 0 && sdf < 1e-5 (:90 / :169)
                bool c = false;
#pragma unroll
                for (int k = 0; k < kVPL; k++)
                    c |= (dv[k].y > 0.0f) & (dv[k].x < carveMax);
                if (!__any_sync(0xffffffffu, c))
                    continue;
            }
            const BrickLane L = brick_lane_setup(fp, map, orgx, orgy, orgz, bx, by, bz, lane);
            int nUpd = 0, nCarve = 0, nCol = 0;
            unsigned wd = 0u;
            if (!frame_on_half_brick<CS, COLOR_PATH, PER_PIXEL, true>(fp, map, L, half, hasCol, dv, cv, wd, wroteC, nUpd, nCarve, nCol, carvable))
                frame_on_half_brick<CS, COLOR_PATH, PER_PIXEL, false>(fp, map, L, half, hasCol, dv, cv, wd, wroteC, nUpd, nCarve, nCol, carvable);
            wroteD |= wd;
            if (__any_sync(0xffffffffu, wd != 0u))
            {
                batch_count_frame(&sB, f, nUpd, nCarve, nCol, lane);
                updMask |= 1u << f;
            }
        }
        if (!updMask)
            continue;                                               // nothing changed: no store, no chunk, no dirty mark
        if (virt)
        {
            if (lane == 0)
                slot = get_or_create_chunk(bp, map, key, x, y, z);
            slot = __shfl_sync(0xffffffffu, slot, 0);
            if (slot < 0)
                continue;                                           // pool exhausted: the unit is dropped, error_flags says so
        }
        {
            float2 *dist = dist_ptr(map, slot);
            unsigned *col = hasCol ? reinterpret_cast<unsigned *>(color_ptr(map, slot)) : nullptr;
#pragma unroll
            for (int k = 0; k < kVPL; k++)
            {
                const int idx = idx0 + (k >> 1) * CS * CS + (k & 1) * 4 * CS;
                if ((wroteD >> k) & 1u)
                    dist[idx] = dv[k];
                if (hasCol && ((wroteC >> k) & 1u))
                    col[idx] = cv[k];
            }
        }
        if (__any_sync(0xffffffffu, carvable) && lane == 0)
            atomicOr(&map.brick_flags[slot], 1ull << b);
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
    batch_flush(bp, &sB);
    batch_snapshot(bp, map);
}

// ------------------------------------------------------------------------------------------------------
// batch_bricks_fast_kernel -- the brick kernel for the common case (constant truncator, colour camera == depth camera, weight
// update in [2^-3, 2^10]); everything else goes through batch_bricks_kernel above. Same tasks, same results, about half the
// instructions per voxel and no dependent gather latency between the frames of a task:
//   * the frame's constants come from the kernel's parameter block (constant bank, warp-uniform loads): no per-CTA staging,
//     no shared-memory reads in the loop;
//   * SOFTWARE PIPELINE over the frames of a task: projection + depth/colour gathers of frame f+1 are issued BEFORE the
//     update of frame f (the pixel a voxel projects to does not depend on the voxel's state), so a task's chain is one gather
//     latency plus K x ALU instead of K x (gather latency + ALU);
//   * straight-line, predicate-selected update. What the reference decides with a branch is evaluated with exactly its
//     operations; the operand-range preconditions of the guard-free reciprocal / quotient are
//     established ONCE per task from the loaded state (weights in [0, 2^20], |sdf| <= 2^17) plus one chained compare per
//     voxel and frame, and anything outside them (also a numerator that is exactly zero) redoes the frame with the IEEE
//     intrinsics (exact_frame) -- nothing has been modified at that point;
//   * on-image test on the bit patterns (u >= 0 && u < W  <=>  bits(u) < bits(W) for u != -0, which cannot occur once the
//     principal point's -0 is canonicalised), pixel index from two round-toward-zero magic-number adds instead of F2I;
//   * carving, rare in practice, is detected by one chained compare per voxel and handled out of line;
//   * per-lane change flags instead of per-voxel masks: a lane that changed anything stores its kVPL voxels (same sectors).
#ifndef CHS_FAST_THREADS
#define CHS_FAST_THREADS 128
#endif
#ifndef CHS_FAST_MIN_CTAS
#define CHS_FAST_MIN_CTAS 4
#endif

struct VoxState
{
    float sdf[kVPL], w[kVPL];
    unsigned cv[kVPL];
    unsigned cnt;          // this frame's per-lane counts: n_upd | n_carve << 10 | n_col << 20 (the order of BatchShared's arrays)
    unsigned touchedC;     // != 0: a colour voxel of the lane changed
};

// What project_gather leaves for apply_frame.
struct Fetch
{
    float d[kVPL];         // depth under the voxel's projection; NaN when the voxel is not on the image (then nothing can happen)
    float cz[kVPL];        // camera-space z
    unsigned c[kVPL];      // packed colour pixel (only fetched while the voxel's colour weight is below 8)
    unsigned slow;         // != 0: some camera-space z of the lane is outside the guard-free reciprocal's range
};

// guard-free correctly rounded quotient for 2^-40 <= |a| < 2^40, 2^-40 <= b < 2^40 (a != 0: the callers send zero numerators
// through the exact path, so the sign-of-zero fix of div_rn_inrange is not needed here)
__device__ __forceinline__ float div_rn_fast(float a, float b)
{
    float r = mufu_rcp(b);
    r = __fmaf_rn(r, __fmaf_rn(-b, r, 1.0f), r);
    const float q = __fmul_rn(a, r);
    return __fmaf_rn(r, __fmaf_rn(-b, q, a), q);
}

// 32-bit read-only load under a predicate, without a branch: `dflt` when the predicate is false (the address is not touched).
// Default and load write the same register inside one asm block, so that no copy of the loaded value (= a wait for the load)
// lands behind it.
__device__ __forceinline__ unsigned ldg_if(const unsigned *p, bool pred, unsigned dflt)
{
    unsigned r;
    asm("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\tmov.b32 %0, %3;\n\t@p ld.global.nc.b32 %0, [%1];\n\t}" : "=r"(r) : "l"(p), "r"((int)pred), "r"(dflt));
    return r;
}

// Projection of the lane's kVPL voxels into frame F and the gathers under them (PinholeCamera.cpp:38-45, 61-64;
// ProjectionIntegrator.h:64-72 / :117-131). px, py[2], pz[kNS]: world coordinates of the lane's voxel centres.
// FIRST: the first frame of a task, issued while the voxel state is still on its way from memory -- the colour pixel is then
// fetched without looking at the colour weight.
template <bool COLOR_PATH, bool FIRST>
__device__ __forceinline__ void project_gather(const BrickFrame &F, float px, const float (&py)[2], const float (&pz)[kNS], bool hasCol,
                                               const unsigned (&cv)[kVPL], Fetch &o)
{
    const float d0 = __fsub_rn(px, F.t[0]);
    const float m00 = __fmul_rn(F.R[0], d0), m01 = __fmul_rn(F.R[1], d0), m02 = __fmul_rn(F.R[2], d0);
    float m1[2][3];
#pragma unroll
    for (int h = 0; h < 2; h++)
    {
        const float d1 = __fsub_rn(py[h], F.t[1]);
        m1[h][0] = __fmul_rn(F.R[3], d1);
        m1[h][1] = __fmul_rn(F.R[4], d1);
        m1[h][2] = __fmul_rn(F.R[5], d1);
    }
    const unsigned wBits = __float_as_uint(F.Wf), hBits = __float_as_uint(F.Hf);
    unsigned slow = 0u;
#pragma unroll
    for (int s = 0; s < kNS; s++)
    {
        const float d2 = __fsub_rn(pz[s], F.t[2]);
        const float m20 = __fmul_rn(F.R[6], d2), m21 = __fmul_rn(F.R[7], d2), m22 = __fmul_rn(F.R[8], d2);
#pragma unroll
        for (int h = 0; h < 2; h++)
        {
            const int k = s * 2 + h;
            // c_j = R(0,j) d0 + (R(1,j) d1 + R(2,j) d2): Eigen's order
            const float cx = __fadd_rn(m00, __fadd_rn(m1[h][0], m20));
            const float cy = __fadd_rn(m01, __fadd_rn(m1[h][1], m21));
            const float cz = __fadd_rn(m02, __fadd_rn(m1[h][2], m22));
            o.cz[k] = cz;
            // 2^-64 <= cz < 2^64: the guard-free reciprocal is exact there. Anything else (voxels behind or in the plane of the
            // camera, absurd coordinates) sends the warp's frame through the exact path; bricks that straddle the camera plane
            // are rare (they lie in free space unless a surface is closer than the band).
            const unsigned zb = __float_as_uint(cz);
            const bool inR = (zb - 0x1F800000u) < 0x40000000u;
            slow |= inR ? 0u : 1u;
            const float invZ = rcp_rn_inrange(cz);
            const float u = __fadd_rn(__fmul_rn(__fmul_rn(F.fx, cx), invZ), F.cx);
            const float v = __fadd_rn(__fmul_rn(__fmul_rn(F.fy, cy), invZ), F.cy);
            // u >= 0 && u < W && v >= 0 && v < H on the bit patterns (u, v are never -0: see BrickFrame)
            const bool on = inR & (__float_as_uint(u) < wBits) & (__float_as_uint(v) < hBits);
            // (int)u + (int)v * W: 2^23 + floor(x) has floor(x) in its low mantissa bits for 0 <= x < 2^23
            const unsigned iu = __float_as_uint(__fadd_rz(u, 8388608.0f)), iv = __float_as_uint(__fadd_rz(v, 8388608.0f));
            const unsigned pix = iv * (unsigned)F.W + (unsigned)F.pix_bias + iu;
            // predicated loads (no branch): NaN depth / colour 0 for voxels that are not on the image. The colour of a voxel is
            // frozen once its colour weight reaches 8 (:153); weights only grow, so a voxel that is below 8 now may need this
            // frame's pixel, one that is not never will
            o.d[k] = __uint_as_float(ldg_if(reinterpret_cast<const unsigned *>(F.depth) + pix, on, 0x7fc00000u));
            o.c[k] = (COLOR_PATH && hasCol) ? ldg_if(F.color + pix, FIRST ? on : (on & (cv[k] < 0x08000000u)), 0u) : 0u;
        }
    }
    o.slow = slow;
}

// One frame applied to the lane's voxels (ProjectionIntegrator.h:72-97 / :131-179, DistVoxel.h:52-60, ColorVoxel.h:65-85).
// Returns false -- before anything is modified -- when an operand is outside the guard-free forms' ranges.
template <bool COLOR_PATH>
__device__ __forceinline__ bool apply_frame(const BrickFrame &F, const Fetch &g, bool hasCol, VoxState &v)
{
    float sd[kVPL], num[kVPL], den[kVPL];
    bool band[kVPL];
    bool far = false, tiny = false, anyCol = false;
#pragma unroll
    for (int k = 0; k < kVPL; k++)
    {
        sd[k] = __fsub_rn(g.d[k], g.cz[k]);                                              // :80 / :139
        // depth path: skip depth > 50 (:74); colour path: skip NaN (:134) and depth > 100 (:141). A NaN depth (also: voxel off
        // the image) fails every comparison below in both paths.
        const bool valid = g.d[k] <= F.cutoff;
        band[k] = valid & (fabsf(sd[k]) < F.thr_band);                                   // :82 / :143
        far |= sd[k] > F.thr_carve;                                                      // :88 / :166 (superset: the rest is tested out of line)
        num[k] = __fadd_rn(__fmul_rn(v.w[k], v.sdf[k]), __fmul_rn(F.wu, sd[k]));         // DistVoxel.h:52-60
        den[k] = __fadd_rn(F.wu, v.w[k]);
        tiny |= fabsf(num[k]) < 9.094947017729282e-13f;                                  // 2^-40 (also exactly zero)
        if (COLOR_PATH)
            anyCol |= band[k] & (v.cv[k] < 0x08000000u);
    }
    if (__any_sync(0xffffffffu, tiny | (g.slow != 0u)))
        return false;
    unsigned cnt = 0u;
    if (__any_sync(0xffffffffu, far))
    {
        // carving: weight > 0 && sdf < 1e-5 (:90 / :169); colour path: weight < 5 resets, else weight -= 1 (:171-175)
#pragma unroll
        for (int k = 0; k < kVPL; k++)
        {
            const bool crv = (g.d[k] <= F.cutoff) & !band[k] & (sd[k] > F.thr_carve) & (v.w[k] > 0.0f) & (v.sdf[k] < F.carve_max);
            const bool dec = COLOR_PATH && !(v.w[k] < 5.0f);
            const float cs = dec ? v.sdf[k] : 99999.0f, cw = dec ? __fsub_rn(v.w[k], 1.0f) : 0.0f;
            v.sdf[k] = crv ? cs : v.sdf[k];
            v.w[k] = crv ? cw : v.w[k];
            cnt += crv ? (1u << 10) : 0u;
        }
    }
    if (COLOR_PATH && hasCol && __any_sync(0xffffffffu, anyCol))
    {
#pragma unroll
        for (int k = 0; k < kVPL; k++)
        {
            const bool colOk = band[k] & (v.cv[k] < 0x08000000u);                        // :153
            const unsigned cw = v.cv[k] >> 24;
            const unsigned mrec = cRecip20[min(cw, 7u) + 1];
            // r and b share one multiply-add (16-bit lanes: w * old + new <= 2040)
            const unsigned nrb = cw * (v.cv[k] & 0x00FF00FFu) + (g.c[k] & 0x00FF00FFu);
            const unsigned ng = cw * __byte_perm(v.cv[k], 0u, 0x4441) + __byte_perm(g.c[k], 0u, 0x4441);
            const unsigned r = ((nrb & 0xFFFFu) * mrec) >> 20, b = ((nrb >> 16) * mrec) >> 20, gg = (ng * mrec) >> 20;
            // r | g << 8 | b << 16 | (w + 1) << 24 (r, g, b <= 255: their upper bytes are the zero bytes of the permutes)
            const unsigned nv = __byte_perm(__byte_perm(__byte_perm(r, gg, 0x1140), b, 0x3410), v.cv[k] + 0x01000000u, 0x7210);
            v.cv[k] = colOk ? nv : v.cv[k];
            cnt += colOk ? (1u << 20) : 0u;
        }
        v.touchedC |= cnt & (0x3FFu << 20);
    }
#pragma unroll
    for (int k = 0; k < kVPL; k++)
    {
        const float q = div_rn_fast(num[k], den[k]);
        const float w2 = __fadd_rn(v.w[k], F.wu);
        v.sdf[k] = band[k] ? q : v.sdf[k];
        v.w[k] = band[k] ? w2 : v.w[k];
        cnt += band[k] ? 1u : 0u;
    }
    v.cnt = cnt;
    return true;
}

// The frame redone with the IEEE intrinsics and per-voxel branches (frame_on_half_brick<..., FAST = false>): rare, out of line.
template <int CS, bool COLOR_PATH>
__device__ __noinline__ VoxState exact_frame(const FrameParams *fpp, const DeviceMap *mapp, VoxState v, int packedBrick, float orgx, float orgy, float orgz, bool hasCol)
{
    const int lane = threadIdx.x & 31;
    const int bx = packedBrick & 0xFF, by = (packedBrick >> 8) & 0xFF, bz = (packedBrick >> 16) & 0xFF, half = packedBrick >> 24;
    const FrameParams &fp = *fpp;
    const DeviceMap &map = *mapp;
    const BrickLane L = brick_lane_setup(fp, map, orgx, orgy, orgz, bx, by, bz, lane);
    float2 dv[kVPL];
#pragma unroll
    for (int k = 0; k < kVPL; k++)
        dv[k] = make_float2(v.sdf[k], v.w[k]);
    unsigned wd = 0u, wc = 0u;
    int nUpd = 0, nCarve = 0, nCol = 0;
    bool carvable = false;
    frame_on_half_brick<CS, COLOR_PATH, false, false>(fp, map, L, half, hasCol, dv, v.cv, wd, wc, nUpd, nCarve, nCol, carvable);
#pragma unroll
    for (int k = 0; k < kVPL; k++)
    {
        v.sdf[k] = dv[k].x;
        v.w[k] = dv[k].y;
    }
    v.cnt = (unsigned)nUpd | ((unsigned)nCarve << 10) | ((unsigned)nCol << 20);
    v.touchedC |= wc;
    return v;
}

template <int CS, bool COLOR_PATH, bool HAS_COL>
__global__ void __launch_bounds__(CHS_FAST_THREADS, CHS_FAST_MIN_CTAS)
batch_bricks_fast_kernel(const __grid_constant__ BatchParams bp, const __grid_constant__ DeviceMap map, const __grid_constant__ BrickFrames bf)
{
    constexpr int BPA = CS / 8;
    constexpr int kSlabCache = 256;                     // slab base pointers kept in shared memory (256 K chunks); beyond: global
    __shared__ BatchShared sB;
    __shared__ float2 *sDist[kSlabCache];
    __shared__ uchar4 *sCol[kSlabCache];
    batch_shared_zero(&sB);
    constexpr bool hasCol = COLOR_PATH && HAS_COL;          // HAS_COL: the map stores colour voxels (compile time: no duplicated code paths)
    {
        const int nSlabs = min((map.capacity + kSlabChunks - 1) >> kSlabChunksLog2, kSlabCache);
        for (int i = threadIdx.x; i < nSlabs; i += blockDim.x)
        {
            sDist[i] = map.dist_slabs[i];
            sCol[i] = hasCol ? map.color_slabs[i] : nullptr;
        }
    }
    __syncthreads();
    pdl_wait();                                    // the candidates kernel's unit lists and counters are read from here on
    batch_span_start(bp);
    const int lane = threadIdx.x & 31;
    const UnitCounts uc = unit_counts(bp);
    const int nTasks = uc.total * kParts;
    const bool idle = on_reserved_sm(bp);          // this CTA takes no tasks (and must not draw any from the queue)
    const float carveMax = bf.f[0].carve_max;
    // Task queue, two entries deep, kept in lanes 0 and 1: entry p is requested (atomicAdd) at a task boundary, its unit record is
    // loaded at the next boundary and it is consumed at the one after, so neither round trip is ever waited for. Tasks are handed
    // out in cost order (bucket 0 first).
    int qg = 0x7fffffff;
    int4 qu = make_int4(0, 0, 0, 0);
    if (!idle)
    {
        int base = 0;
        if (lane == 0)
            base = atomicAdd(&bp.bctr->next_task, 2);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (lane < 2)
            qg = base + lane;
        if (lane == 0 && qg < nTasks)
            qu = bp.units[unit_position(bp, uc, qg / kParts)];
    }
    for (int t = 0;; t++)
    {
        const int p = t & 1;
        const int g = __shfl_sync(0xffffffffu, qg, p);
        if (g >= nTasks)
            break;
        int4 unit;
        unit.x = __shfl_sync(0xffffffffu, qu.x, p);
        unit.y = __shfl_sync(0xffffffffu, qu.y, p);
        unit.z = __shfl_sync(0xffffffffu, qu.z, p);
        unit.w = __shfl_sync(0xffffffffu, qu.w, p);
        if (lane == p)
            qg = atomicAdd(&bp.bctr->next_task, 1);                     // consumed two boundaries from now
        if (lane == 1 - p && qg < nTasks)
            qu = bp.units[unit_position(bp, uc, qg / kParts)];         // requested at the previous boundary, consumed at the next
        const int half = g % kParts;                     // which group of kNS z slices of the brick
        int x, y, z;
        const unsigned long long key = ((unsigned long long)(unsigned)unit.y << 32) | (unsigned long long)(unsigned)unit.x;
        unpack_id(key, &x, &y, &z);
        int slot = unit.z & 0xFFFFFF;
        const int b = unit.z >> 24;
        const bool virt = slot == kVirtualSlot;
        const unsigned bandM = (unsigned)unit.w & 0xFFFFu;
        unsigned mask = bandM | ((unsigned)unit.w >> 16);
        const float orgx = __fmul_rn((float)(CS * x), map.res), orgy = __fmul_rn((float)(CS * y), map.res), orgz = __fmul_rn((float)(CS * z), map.res);
        const int bx = b % BPA, by = (b / BPA) % BPA, bz = b / (BPA * BPA);
        const int idx0 = ((bz * 8 + kNS * half) * CS + (by * 8 + (lane >> 3))) * CS + bx * 8 + (lane & 7);
        VoxState v;
        v.touchedC = 0u;
        v.cnt = 0u;
        if (!virt)
        {
            const int sl = slot >> kSlabChunksLog2;
            const size_t off = (size_t)(slot & (kSlabChunks - 1)) * map.V;
            const float2 *dist = (sl < kSlabCache ? sDist[sl] : map.dist_slabs[sl]) + off;
            const unsigned *col = hasCol ? reinterpret_cast<const unsigned *>((sl < kSlabCache ? sCol[sl] : map.color_slabs[sl]) + off) : nullptr;
#pragma unroll
            for (int k = 0; k < kVPL; k++)
            {
                const int idx = idx0 + (k >> 1) * CS * CS + (k & 1) * 4 * CS;
                const float2 d = dist[idx];
                v.sdf[k] = d.x;
                v.w[k] = d.y;
                v.cv[k] = hasCol ? col[idx] : 0u;
            }
        }
        else
        {
#pragma unroll
            for (int k = 0; k < kVPL; k++)
            {
                v.sdf[k] = 99999.0f;                                // Chunk::Chunk initial state (DistVoxel.cpp:29-33)
                v.w[k] = 0.0f;
                v.cv[k] = 0u;
            }
        }
        // world coordinates of the lane's voxel centres: centre_k = float(k) * res + res/2 (ChunkManager.cpp:52,61), p = centre +
        // origin (ProjectionIntegrator.h:64). Lane = (x = lane & 7, y rows lane >> 3 and + 4), kNS z slices.
        const float px = __fadd_rn(__fadd_rn(__fmul_rn((float)(bx * 8 + (lane & 7)), map.res), map.half), orgx);
        float py[2], pz[kNS];
#pragma unroll
        for (int h = 0; h < 2; h++)
            py[h] = __fadd_rn(__fadd_rn(__fmul_rn((float)(by * 8 + (lane >> 3) + 4 * h), map.res), map.half), orgy);
#pragma unroll
        for (int s = 0; s < kNS; s++)
            pz[s] = __fadd_rn(__fadd_rn(__fmul_rn((float)(bz * 8 + kNS * half + s), map.res), map.half), orgz);
        // first frame's gathers go out before the voxel state is needed
        int f = __ffs(mask) - 1;
        mask &= mask - 1;
        Fetch cur, nxt;
        project_gather<COLOR_PATH, true>(bf.f[f], px, py, pz, hasCol, v.cv, cur);
        // preconditions of the guard-free quotient over the whole task: 0 <= weight <= 2^20 (bit pattern compare: negative and
        // NaN weights fail), |sdf| <= 2^17. Weights grow by at most 16 * 2^10 and |sdf| stays below max(|sdf|, band) inside a task.
        bool stateOk = true;
#pragma unroll
        for (int k = 0; k < kVPL; k++)
            stateOk &= (__float_as_uint(v.w[k]) <= 0x49800000u) & (fabsf(v.sdf[k]) <= 131072.0f);
        const bool exactTask = !__all_sync(0xffffffffu, stateOk);
        unsigned touchedD = 0u, updMask = 0u;
        // One step: issue projection + gathers of the next frame of the mask into `b`, then apply frame f from `a`. The two
        // buffers swap roles every step (the loop below is unrolled by two), so nothing is copied.
        auto step = [&](const Fetch &a, Fetch &b) -> bool
        {
            int fn = -1;
            if (mask)
            {
                fn = __ffs(mask) - 1;
                mask &= mask - 1;
                project_gather<COLOR_PATH, false>(bf.f[fn], px, py, pz, hasCol, v.cv, b);
            }
            bool run = true;
            if (!((bandM >> f) & 1u))
            {
                // free-space frame: it can only carve, and only voxels with weight > 0 && sdf < 1e-5 (:90 / :169)
                bool c = false;
#pragma unroll
                for (int k = 0; k < kVPL; k++)
                    c |= (v.w[k] > 0.0f) & (v.sdf[k] < carveMax);
                run = __any_sync(0xffffffffu, c);
            }
            if (run)
            {
                if (exactTask || !apply_frame<COLOR_PATH>(bf.f[f], a, hasCol, v))
                    v = exact_frame<CS, COLOR_PATH>(bp.frames + f, &map, v, bx | (by << 8) | (bz << 16) | (half << 24), orgx, orgy, orgz, hasCol);
                const unsigned tot = __reduce_add_sync(0xffffffffu, v.cnt);
                touchedD |= v.cnt;
                if (tot)
                {
                    updMask |= 1u << f;
                    if (lane < 3)
                    {
                        const unsigned n = (tot >> (10 * lane)) & 0x3FFu;
                        if (n)
                            atomicAdd(&sB.upd[lane * kMaxBatch + f], (int)n);      // upd, carve, col are consecutive arrays of BatchShared
                    }
                }
            }
            f = fn;
            return fn >= 0;
        };
        while (step(cur, nxt) && step(nxt, cur))
        {
        }
        if (!updMask)
            continue;                                               // nothing changed: no store, no chunk, no dirty mark
        if (virt)
        {
            if (lane == 0)
                slot = get_or_create_chunk(bp, map, key, x, y, z);
            slot = __shfl_sync(0xffffffffu, slot, 0);
            if (slot < 0)
                continue;                                           // pool exhausted: the unit is dropped, error_flags says so
        }
        bool carvable = false;
        {
            const int sl = slot >> kSlabChunksLog2;
            const size_t off = (size_t)(slot & (kSlabChunks - 1)) * map.V;
            float2 *dist = (sl < kSlabCache ? sDist[sl] : map.dist_slabs[sl]) + off;
            unsigned *col = hasCol ? reinterpret_cast<unsigned *>((sl < kSlabCache ? sCol[sl] : map.color_slabs[sl]) + off) : nullptr;
#pragma unroll
            for (int k = 0; k < kVPL; k++)
            {
                const int idx = idx0 + (k >> 1) * CS * CS + (k & 1) * 4 * CS;
                if (touchedD)
                    dist[idx] = make_float2(v.sdf[k], v.w[k]);
                if (hasCol && v.touchedC)
                    col[idx] = v.cv[k];
                carvable |= (v.w[k] > 0.0f) & (v.sdf[k] < carveMax);
            }
        }
        if (__any_sync(0xffffffffu, carvable) && lane == 0)
            atomicOr(&map.brick_flags[slot], 1ull << b);
        // (batch id << 32) | frames of this batch that updated the chunk: the first warp of the batch marks the 27 neighbour IDs
        // dirty (Chisel.h:89-101, 175-189); every newly set frame bit counts the chunk once for that frame
        unsigned newBits = 0u;
        int first = 0;
        if (lane == 0)
        {
            const unsigned long long tag = (unsigned long long)(unsigned)bp.batch_id << 32;
            unsigned long long old = bp.slot_batch[slot], assumed, curv;
            do
            {
                assumed = old;
                curv = ((assumed >> 32) == (unsigned long long)(unsigned)bp.batch_id) ? assumed : tag;
                const unsigned long long nw = curv | updMask;
                if (nw == assumed)
                    break;
                old = atomicCAS(&bp.slot_batch[slot], assumed, nw);
            } while (old != assumed);
            first = (assumed >> 32) != (unsigned long long)(unsigned)bp.batch_id;
            newBits = updMask & ~(unsigned)(curv & 0xffffffffull);
        }
        first = __shfl_sync(0xffffffffu, first, 0);
        newBits = __shfl_sync(0xffffffffu, newBits, 0);
        if (first && lane < 27)
            dirty_insert(map, pack_id(x + lane / 9 - 1, y + (lane / 3) % 3 - 1, z + lane % 3 - 1));
        if (lane < kMaxBatch && ((newBits >> lane) & 1u))
            atomicAdd(&sB.chunks[lane], 1);
    }
    batch_flush(bp, &sB);
    batch_snapshot(bp, map);
}

// ------------------------------------------------------------------------------------------------------
// Self-test of the range-check-free reciprocal and quotient against the IEEE intrinsics (chs_selftest_arithmetic).
__global__ void selftest_rcp_kernel(unsigned long long *mismatches, unsigned long long *tested)
{
    unsigned long long bad = 0, n = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < (1ull << 32); i += (unsigned long long)gridDim.x * blockDim.x)
    {
        const float z = __uint_as_float((unsigned)i);
        if (!rcp_in_range(z))
            continue;
        n++;
        bad += __float_as_uint(rcp_rn_inrange(z)) != __float_as_uint(__frcp_rn(z));
    }
    atomicAdd(mismatches, bad);
    atomicAdd(tested, n);
}

__global__ void selftest_div_kernel(unsigned long long *mismatches, unsigned long long *tested, unsigned long long pairs)
{
    unsigned long long bad = 0, n = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += (unsigned long long)gridDim.x * blockDim.x)
    {
        const unsigned long long h = mix64(i * 0x9E3779B97F4A7C15ull + 1);
        float a = __uint_as_float((unsigned)h), b = __uint_as_float((unsigned)(h >> 32));
        if ((i & 3) == 1)
        {
            // the shape DistVoxel::Integrate produces: (w * sdf + wu * d) / (wu + w) with small weights
            const float w = (float)((h >> 8) & 63), wu = 0.5f * (float)(1 + ((h >> 14) & 15));
            const float sdf = 0.4f * (__uint_as_float(0x3f800000u | ((unsigned)(h >> 20) & 0x7fffffu)) - 1.5f);
            const float d = 0.4f * (__uint_as_float(0x3f800000u | ((unsigned)(h >> 41) & 0x7fffffu)) - 1.5f);
            a = __fadd_rn(__fmul_rn(w, sdf), __fmul_rn(wu, d));
            b = __fadd_rn(wu, w);
        }
        else if ((i & 3) == 2)
            a = (i & 4) ? 0.0f : -0.0f;
        if (!div_in_range(a, b))
            continue;
        n++;
        bad += __float_as_uint(div_rn_inrange(a, b)) != __float_as_uint(__fdiv_rn(a, b));
    }
    atomicAdd(mismatches, bad);
    atomicAdd(tested, n);
}

cudaError_t launch_selftest_arithmetic(unsigned long long *dCounters, unsigned long long divPairs, cudaStream_t st)
{
    selftest_rcp_kernel<<<148 * 8, 256, 0, st>>>(dCounters, dCounters + 1);
    selftest_div_kernel<<<148 * 8, 256, 0, st>>>(dCounters + 2, dCounters + 3, divPairs);
    return cudaGetLastError();
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

// Launch with the programmatic-stream-serialization attribute: the kernel may begin while its predecessor in the stream drains
// (it calls pdl_wait() before touching the predecessor's output).
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args)
{
    static const bool off = std::getenv("CHS_NO_PDL") != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = off ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

template <int CS, bool COLOR_PATH, bool PER_PIXEL>
static cudaError_t launch_batch_variant(BatchParams bp, const DeviceMap &map, const BatchLaunchInfo &info, cudaEvent_t *evt, const BatchStreams &bs, int phases)
{
    host_launch_lap(9);
    // (the occupancy query costs 15 us of host time: once per kernel variant, whatever it answers)
    static int residentBricks = -1;
    if (residentBricks < 0)
        residentBricks = std::max(1, batch_resident(batch_bricks_kernel<CS, COLOR_PATH, PER_PIXEL>, CHS_BRICK_THREADS));
    host_launch_lap(10);
    constexpr long long NB = (CS / 8) * (CS / 8) * (CS / 8);
    // a warp per tile of 8 chunk indices (at N ranks: 8 x N indices, about 8 of them owned)
    const long long tileIdx = kCandChunks * std::min(std::max(map.world, 1), 8);
    const long long candTiles = (info.unionCandidates + tileIdx - 1) / tileIdx;
    const unsigned gCand = (unsigned)std::max(1ll, std::min<long long>((candTiles + kCandWarps - 1) / kCandWarps, 148 * 2 * CHS_CAND_MIN_CTAS));
    const unsigned gBricks = (unsigned)std::max(1ll, std::min<long long>((info.unionCandidates * NB * kParts + CHS_BRICK_THREADS / 32 - 1) / (CHS_BRICK_THREADS / 32), residentBricks));
    static_assert(CHS_BRICK_THREADS % 32 == 0, "whole warps");
    bp.total_ctas = (int)gBricks;
    cudaStream_t st = bs.main;
    cudaError_t e;
    host_launch_lap(11);
    if (phases & 1)
    {
    if (info.profiling && (e = cudaEventRecord(evt[0], bs.prep)) != cudaSuccess)
        return e;
    host_launch_lap(7);
    const int tilesX = (info.W + 63) / 64, tiles = tilesX * ((info.H + 63) / 64);
    const int packBlocks = info.colorPath ? std::max(1, std::min(148, (info.cW * info.cH / 4 + 255) / 256)) : 0;
    if (packBlocks && bs.pack != bs.prep)
    {
        // device frames: fork the packing off the main stream so that it runs beside Hi-Z + candidates
        if ((e = cudaEventRecord(bs.fork, st)) != cudaSuccess || (e = cudaStreamWaitEvent(bs.pack, bs.fork, 0)) != cudaSuccess)
            return e;
        batch_pack_kernel<<<dim3(packBlocks, bp.K), 256, 0, bs.pack>>>(bp);
        if ((e = cudaEventRecord(bs.packed, bs.pack)) != cudaSuccess)
            return e;
    }
    host_launch_lap(0);
    const int hizFirst = info.hizCount > 0 ? info.hizFirst : 0, hizCount = info.hizCount > 0 ? info.hizCount : bp.K;
    if (info.skipHiz)
    {
        // the pyramids of this step were built rank by rank and have arrived through the exchange arenas
    }
    else if (info.hizTma)
    {
        static bool attr = false;
        if (!attr)
        {
            if ((e = cudaFuncSetAttribute(batch_hiz_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHizSmem)) != cudaSuccess)
                return e;
            attr = true;
        }
        const int nItems = tiles * hizCount;
        batch_hiz_tma_kernel<false><<<std::min(nItems, 148 * 4), kHizThreads, kHizSmem, bs.prep>>>(bp, tilesX, tiles, nItems, hizFirst, HizPeers{});
    }
    else
        batch_hiz_kernel<<<dim3(tiles, hizCount), 256, 0, bs.prep>>>(bp, map, tilesX, hizFirst);
    host_launch_lap(1);
    if (info.afterHiz && info.afterHiz(info.afterHizCtx) != 0)
        return cudaErrorUnknown;
    if (info.profiling && (e = cudaEventRecord(evt[1], bs.prep)) != cudaSuccess)
        return e;
    if (bs.prep != st && ((e = cudaEventRecord(bs.prepared, bs.prep)) != cudaSuccess || (e = cudaStreamWaitEvent(st, bs.prepared, 0)) != cudaSuccess))
        return e;
    if (packBlocks && bs.pack == bs.prep)
    {
        batch_pack_kernel<<<dim3(packBlocks, bp.K), 256, 0, bs.pack>>>(bp);
        if (bs.pack != st && (e = cudaEventRecord(bs.packed, bs.pack)) != cudaSuccess)
            return e;
    }
    host_launch_lap(2);
    if ((e = launch_pdl(batch_candidates_kernel<CS>, dim3(gCand), dim3(32 * kCandWarps), 0, st, bp, map)) != cudaSuccess)
        return e;
    host_launch_lap(3);
    if (info.profiling && (e = cudaEventRecord(evt[2], st)) != cudaSuccess)
        return e;
    if (packBlocks && bs.pack != st && (e = cudaStreamWaitEvent(st, bs.packed, 0)) != cudaSuccess)
        return e;
    }
    host_launch_lap(4);
    if (!(phases & 2))
        return cudaGetLastError();
    if (info.profiling && (e = cudaEventRecord(evt[7], st)) != cudaSuccess)
        return e;
    if (!PER_PIXEL && info.fastBricks)
    {
        static int residentFast = -1;
        if (residentFast < 0)
            residentFast = std::max(1, batch_resident(batch_bricks_fast_kernel<CS, COLOR_PATH, COLOR_PATH>, CHS_FAST_THREADS));
        const unsigned gFast = (unsigned)std::max(1ll, std::min<long long>((info.unionCandidates * NB * kParts + CHS_FAST_THREADS / 32 - 1) / (CHS_FAST_THREADS / 32), residentFast));
        bp.total_ctas = (int)gFast;
        if (COLOR_PATH && map.use_color)
            e = launch_pdl(batch_bricks_fast_kernel<CS, COLOR_PATH, COLOR_PATH>, dim3(gFast), dim3(CHS_FAST_THREADS), 0, st, bp, map, *info.brickFrames);
        else
            e = launch_pdl(batch_bricks_fast_kernel<CS, COLOR_PATH, false>, dim3(gFast), dim3(CHS_FAST_THREADS), 0, st, bp, map, *info.brickFrames);
    }
    else
        e = launch_pdl(batch_bricks_kernel<CS, COLOR_PATH, PER_PIXEL>, dim3(gBricks), dim3(CHS_BRICK_THREADS), 0, st, bp, map);
    host_launch_lap(5);
    if (e != cudaSuccess)
        return e;
    if (info.profiling && (e = cudaEventRecord(evt[3], st)) != cudaSuccess)
        return e;
    return cudaGetLastError();
}

template <int CS>
static cudaError_t launch_batch_cs(const BatchParams &bp, const DeviceMap &map, const BatchLaunchInfo &info, cudaEvent_t *evt, const BatchStreams &bs, int phases)
{
    if (info.colorPath)
        return info.perPixel ? launch_batch_variant<CS, true, true>(bp, map, info, evt, bs, phases)
                             : launch_batch_variant<CS, true, false>(bp, map, info, evt, bs, phases);
    return info.perPixel ? launch_batch_variant<CS, false, true>(bp, map, info, evt, bs, phases)
                         : launch_batch_variant<CS, false, false>(bp, map, info, evt, bs, phases);
}

cudaError_t launch_hiz_sharded(int first, int count, const HizPeers &hp, cudaStream_t st)
{
    static bool attr = false;
    if (!attr)
    {
        const cudaError_t e = cudaFuncSetAttribute(batch_hiz_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHizSmem);
        if (e != cudaSuccess)
            return e;
        attr = true;
    }
    const int tilesX = (hp.W + 63) / 64, tiles = tilesX * ((hp.H + 63) / 64);
    const int nItems = tiles * count;
    // beside the brick kernel of an earlier step only the SMs that kernel leaves idle are free: a small grid is enough, the
    // kernel has a whole step of slack
    batch_hiz_tma_kernel<true><<<std::min(nItems, 148 * 2), kHizThreads, kHizSmem, st>>>(BatchParams{}, tilesX, tiles, nItems, first, hp);
    return cudaGetLastError();
}

cudaError_t launch_batch(const BatchParams &bp, const DeviceMap &map, const BatchLaunchInfo &info, cudaEvent_t *evt, const BatchStreams &bs, int phases)
{
    switch (map.cs)
    {
    case 8: return launch_batch_cs<8>(bp, map, info, evt, bs, phases);
    case 16: return launch_batch_cs<16>(bp, map, info, evt, bs, phases);
    default: return launch_batch_cs<32>(bp, map, info, evt, bs, phases);
    }
}

} // namespace CHS_BATCH_VARIANT
} // namespace chs
