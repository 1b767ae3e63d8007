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
#include "kernels.h"

namespace chs
{

// ------------------------------------------------------------------------------------------------------
// truncation distance, bit-identical to the three shipped Truncator subclasses
//   ConstantTruncator.h:48-51, QuadraticTruncator.h:42-45 (+ :65-67 constants), InverseTruncator.h:42-52
__host__ __device__ inline float truncation_of(int kind, float param, float reading)
{
    if (kind == CHS_TRUNC_QUADRATIC)
    {
        // float members initialised from double constant expressions; the polynomial runs in double
        // because pow() returns double; the linear term is a float product (float * float)
        const float q = (float)(0.0019 * 10), l = (float)(0.00152 * 10), c = (float)(0.001504 * 10);
        const double r = (double)reading;
        const double p = (double)q * (r * r) + (double)(l * reading) + (double)c;   // pow(x, 2) == x*x exactly for binary32 x
        return (float)(fabs(p) * (double)param);
    }
    if (kind == CHS_TRUNC_INVERSE)
    {
        const float base = (float)0.10, focal = (float)471.27;
        const float depSample = 1.0f / (base * focal);
        const float inv = (float)(1.0 / (double)reading);
        return (depSample / (inv * inv)) * param;
    }
    return param;
}

// ------------------------------------------------------------------------------------------------------
// frame_prepare: one CTA per 64x64 pixel block; writes the per-pixel truncation image (non-constant truncators) and
// the four Hi-Z levels (tiles of 8, 16, 32, 64 pixels) holding {min over pixels of depth - band, max of depth + band}.
__device__ __forceinline__ void hiz_accumulate(const FrameParams &fp, float d, float tr, float *lo, float *hi)
{
    // pixels that can never change a voxel: NaN, +-inf, beyond the cutoff (ProjectionIntegrator.h:74,134,141)
    const bool valid = (d == d) && fabsf(d) <= 3.0e38f && !(d > fp.depth_cutoff) && (tr == tr);
    if (valid)
    {
        const float band = tr + fp.diag;
        // carving reaches every z < d - (trunc + carveDist); that is inside (.., d + band) unless carveDist is very negative
        const float farExt = fp.carve ? fmaxf(band, -(tr + fp.carve_dist)) : band;
        *lo = fminf(*lo, d - band);
        *hi = fmaxf(*hi, d + farExt);
    }
}

// ColorImage::At (OC ColorImage.h:61-101) once per pixel instead of once per voxel: mono replicates, 3/4 channels are
// B,G,R(,A); stored as r | g << 8 | b << 16 so that the integrate kernels fetch a colour with one 32-bit load. Runs as a
// graph branch parallel to frame_prepare -> chunk_candidates (only the integrate kernels consume its output).
__global__ void __launch_bounds__(256) color_pack_kernel(FrameParams fp)
{
    {
        const int n = fp.ccam.W * fp.ccam.H, ch = fp.channels;
        const int nThreads = gridDim.x * gridDim.y * blockDim.x;
        const int gtid = (blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
        int done = 0;
        if (ch == 3 && (reinterpret_cast<size_t>(fp.color) & 3) == 0 && (reinterpret_cast<size_t>(fp.color_packed) & 15) == 0)
        {
            // BGR fast path: four pixels = three aligned 32-bit words in, one 128-bit word out
            const unsigned *p32 = reinterpret_cast<const unsigned *>(fp.color);
            const int n4 = n >> 2;
            for (int i = gtid; i < n4; i += nThreads)
            {
                const unsigned a = __ldg(p32 + 3 * i), b = __ldg(p32 + 3 * i + 1), c = __ldg(p32 + 3 * i + 2);
                uint4 o;
                o.x = ((a >> 16) & 0xFFu) | (a & 0xFF00u) | ((a & 0xFFu) << 16);                         // B0 G0 R0
                o.y = ((b >> 8) & 0xFFu) | ((b & 0xFFu) << 8) | ((a >> 24) << 16);                       // B1 | G1 R1
                o.z = (c & 0xFFu) | ((b >> 24) << 8) | (((b >> 16) & 0xFFu) << 16);                      // B2 G2 | R2
                o.w = (c >> 24) | (((c >> 16) & 0xFFu) << 8) | (((c >> 8) & 0xFFu) << 16);               // B3 G3 R3
                reinterpret_cast<uint4 *>(fp.color_packed)[i] = o;
            }
            done = n4 << 2;
        }
        for (int i = done + gtid; i < n; i += nThreads)
        {
            const uint8_t *p = fp.color + (size_t)i * ch;
            unsigned r, g, b;
            if (ch >= 3)
            {
                b = __ldg(p);
                g = __ldg(p + 1);
                r = __ldg(p + 2);
            }
            else if (ch == 2)
            {
                r = __ldg(p);
                g = b = __ldg(p + 1);
            }
            else
                r = g = b = __ldg(p);
            fp.color_packed[i] = r | (g << 8) | (b << 16);
        }
    }
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
    const int W = fp.cam.W, H = fp.cam.H;
    const int t = threadIdx.x;
    const int tile = t >> 2, sub = t & 3;                  // 64 tiles of 8x8, 4 threads per tile (2 rows each)
    const int tx = blockIdx.x * 8 + (tile & 7), ty = blockIdx.y * 8 + (tile >> 3);
    float lo = INFINITY, hi = -INFINITY;
    const bool perPixel = fp.trunc_img != nullptr;
    const bool computeTrunc = fp.trunc_kind == CHS_TRUNC_QUADRATIC || fp.trunc_kind == CHS_TRUNC_INVERSE;
    float *truncOut = computeTrunc ? const_cast<float *>(fp.trunc_img) : nullptr;
    const bool vec = ((W & 3) == 0) && ((reinterpret_cast<size_t>(fp.depth) & 15) == 0) && !perPixel;
    const int x0 = tx * 8;
    if (vec && x0 + 8 <= W)
    {
        // constant truncator, aligned interior: four independent 128-bit loads per thread
        float4 v[4];
#pragma unroll
        for (int r = 0; r < 2; r++)
        {
            const int y = ty * 8 + sub * 2 + r;
            const bool in = y < H;
            const float4 *row = reinterpret_cast<const float4 *>(fp.depth + (size_t)(in ? y : 0) * W + x0);
            const float nanv = __int_as_float(0x7fc00000);
            v[2 * r] = in ? __ldg(row) : make_float4(nanv, nanv, nanv, nanv);
            v[2 * r + 1] = in ? __ldg(row + 1) : make_float4(nanv, nanv, nanv, nanv);
        }
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            hiz_accumulate(fp, v[k].x, fp.trunc_param, &lo, &hi);
            hiz_accumulate(fp, v[k].y, fp.trunc_param, &lo, &hi);
            hiz_accumulate(fp, v[k].z, fp.trunc_param, &lo, &hi);
            hiz_accumulate(fp, v[k].w, fp.trunc_param, &lo, &hi);
        }
    }
    else
    {
#pragma unroll
        for (int r = 0; r < 2; r++)
        {
            const int y = ty * 8 + sub * 2 + r;
            if (y >= H)
                continue;
#pragma unroll
            for (int i = 0; i < 8; i++)
            {
                const int x = x0 + i;
                if (x >= W)
                    continue;
                const float d = __ldg(fp.depth + (size_t)y * W + x);
                float tr;
                if (truncOut)
                {
                    tr = truncation_of(fp.trunc_kind, fp.trunc_param, d);
                    truncOut[(size_t)y * W + x] = tr;
                }
                else
                    tr = perPixel ? __ldg(fp.trunc_img + (size_t)y * W + x) : fp.trunc_param;
                hiz_accumulate(fp, d, tr, &lo, &hi);
            }
        }
    }
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, 1));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, 1));
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, 2));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, 2));
    __shared__ float2 s0[64], s1[16], s2[4];
    if (sub == 0)
    {
        s0[tile] = make_float2(lo, hi);
        if (tx < fp.hizW[0] && ty < fp.hizH[0])
            fp.hiz[0][ty * fp.hizW[0] + tx] = make_float2(lo, hi);
    }
    __syncthreads();
    if (t < 16)
    {
        const int ax = t & 3, ay = t >> 2;
        float2 a = s0[(ay * 2) * 8 + ax * 2], b = s0[(ay * 2) * 8 + ax * 2 + 1], c = s0[(ay * 2 + 1) * 8 + ax * 2], d = s0[(ay * 2 + 1) * 8 + ax * 2 + 1];
        const float2 v = make_float2(fminf(fminf(a.x, b.x), fminf(c.x, d.x)), fmaxf(fmaxf(a.y, b.y), fmaxf(c.y, d.y)));
        s1[t] = v;
        const int gx = blockIdx.x * 4 + ax, gy = blockIdx.y * 4 + ay;
        if (gx < fp.hizW[1] && gy < fp.hizH[1])
            fp.hiz[1][gy * fp.hizW[1] + gx] = v;
    }
    __syncthreads();
    if (t < 4)
    {
        const int ax = t & 1, ay = t >> 1;
        float2 a = s1[(ay * 2) * 4 + ax * 2], b = s1[(ay * 2) * 4 + ax * 2 + 1], c = s1[(ay * 2 + 1) * 4 + ax * 2], d = s1[(ay * 2 + 1) * 4 + ax * 2 + 1];
        const float2 v = make_float2(fminf(fminf(a.x, b.x), fminf(c.x, d.x)), fmaxf(fmaxf(a.y, b.y), fmaxf(c.y, d.y)));
        s2[t] = v;
        const int gx = blockIdx.x * 2 + ax, gy = blockIdx.y * 2 + ay;
        if (gx < fp.hizW[2] && gy < fp.hizH[2])
            fp.hiz[2][gy * fp.hizW[2] + gx] = v;
    }
    __syncthreads();
    if (t == 0)
    {
        const float2 v = make_float2(fminf(fminf(s2[0].x, s2[1].x), fminf(s2[2].x, s2[3].x)), fmaxf(fmaxf(s2[0].y, s2[1].y), fmaxf(s2[2].y, s2[3].y)));
        fp.hiz[3][blockIdx.y * fp.hizW[3] + blockIdx.x] = v;
    }
}

// ------------------------------------------------------------------------------------------------------
// Frustum::Intersects (OC Frustum.cpp:41-79), exact: true at the first plane whose far vertex is in front.
__device__ __forceinline__ bool frustum_intersects_exact(const FrameParams &fp, float bminx, float bminy, float bminz,
                                                         float bmaxx, float bmaxy, float bmaxz)
{
#pragma unroll
    for (int p = 0; p < 6; p++)
    {
        const float nx = fp.planes[p][0], ny = fp.planes[p][1], nz = fp.planes[p][2], d = fp.planes[p][3];
        const float ax = (nx < 0.0f) ? bminx : bmaxx;
        const float ay = (ny < 0.0f) ? bminy : bmaxy;
        const float az = (nz < 0.0f) ? bminz : bmaxz;
        const float dot = __fadd_rn(__fmul_rn(ax, nx), __fadd_rn(__fmul_rn(ay, ny), __fmul_rn(az, nz)));
        if (__fadd_rn(dot, d) > 0.0f)
            return true;
    }
    return false;
}

// Conservative depth-range classification of an axis-aligned box of voxel CENTRES (first centre at world position
// (wx, wy, wz), `ext` metres along each axis) against the Hi-Z tiles:
//   0  no voxel of the box can change this frame (off-image, behind the camera, no valid pixel, or behind every surface)
//   1  the box lies entirely in free space in front of every surface: only carving of already-observed voxels can act
//   2  some voxel may fall inside the truncation band
// Free-form float math with explicit slack: may over-report, never under-report.
__device__ int classify_box(const FrameParams &fp, float wx, float wy, float wz, float ext)
{
    const CameraDev &c = fp.cam;
    const float ox = wx - c.t[0], oy = wy - c.t[1], oz = wz - c.t[2];
    float zmin = INFINITY, zmax = -INFINITY, umin = INFINITY, umax = -INFINITY, vmin = INFINITY, vmax = -INFINITY;
    bool nearCross = false;
#pragma unroll
    for (int k = 0; k < 8; k++)
    {
        const float dx = ox + ((k & 1) ? ext : 0.0f), dy = oy + ((k & 2) ? ext : 0.0f), dz = oz + ((k & 4) ? ext : 0.0f);
        const float cx = c.R[0] * dx + c.R[3] * dy + c.R[6] * dz;
        const float cy = c.R[1] * dx + c.R[4] * dy + c.R[7] * dz;
        const float cz = c.R[2] * dx + c.R[5] * dy + c.R[8] * dz;
        zmin = fminf(zmin, cz);
        zmax = fmaxf(zmax, cz);
        if (cz > 1e-2f)
        {
            const float iz = 1.0f / cz;
            const float u = c.fx * cx * iz + c.cx, v = c.fy * cy * iz + c.cy;
            umin = fminf(umin, u);
            umax = fmaxf(umax, u);
            vmin = fminf(vmin, v);
            vmax = fmaxf(vmax, v);
        }
        else
            nearCross = true;
    }
    const float slack = 1e-3f + 1e-5f * fmaxf(fabsf(zmin), fabsf(zmax));
    zmin -= slack;
    zmax += slack;
    if (zmax < 0.0f)
        return 0;                                                       // every centre behind the camera (ProjectionIntegrator.h:68)
    int x0, x1, y0, y1;
    if (nearCross)
    {
        x0 = 0; y0 = 0; x1 = c.W - 1; y1 = c.H - 1;
    }
    else
    {
        // pad by 2 pixels for rounding of the exact projection; clamp in float first (huge values)
        const float fx0 = fmaxf(umin - 2.0f, 0.0f), fx1 = fminf(umax + 2.0f, c.Wf - 1.0f);
        const float fy0 = fmaxf(vmin - 2.0f, 0.0f), fy1 = fminf(vmax + 2.0f, c.Hf - 1.0f);
        if (!(fx0 <= fx1) || !(fy0 <= fy1))
            return 0;                                                   // projects entirely off the image
        x0 = (int)fx0; x1 = (int)fx1; y0 = (int)fy0; y1 = (int)fy1;
    }
    // pick the finest level at which the rectangle spans at most 3 tiles per axis
    int level = 0, shift = 3;
    while (level < kHizLevels - 1 && (((x1 >> shift) - (x0 >> shift)) > 2 || ((y1 >> shift) - (y0 >> shift)) > 2))
    {
        level++;
        shift++;
    }
    float lo = INFINITY, hi = -INFINITY;
    const float2 *tiles = fp.hiz[level];
    const int tw = fp.hizW[level];
    const int tx0 = x0 >> shift, tx1 = x1 >> shift, ty0 = y0 >> shift, ty1 = y1 >> shift;
    if (tx1 - tx0 <= 2 && ty1 - ty0 <= 2)
    {
        // common case: up to 3x3 tiles, all loads issued before the reduction
        float2 v[9];
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
            for (int i = 0; i < 3; i++)
            {
                const int tx = min(tx0 + i, tx1), ty = min(ty0 + j, ty1);
                v[j * 3 + i] = __ldg(tiles + ty * tw + tx);
            }
#pragma unroll
        for (int k = 0; k < 9; k++)
        {
            lo = fminf(lo, v[k].x);
            hi = fmaxf(hi, v[k].y);
        }
    }
    else
        for (int ty = ty0; ty <= ty1; ty++)
            for (int tx = tx0; tx <= tx1; tx++)
            {
                const float2 v = __ldg(tiles + ty * tw + tx);
                lo = fminf(lo, v.x);
                hi = fmaxf(hi, v.y);
            }
    if (!(lo <= hi))
        return 0;                                                       // no valid depth pixel under the box
    const float s2 = 1e-3f + 1e-5f * fmaxf(fabsf(lo), fabsf(hi));
    if (zmin > hi + s2)
        return 0;                                                       // entirely behind every surface it projects onto
    if (zmax < lo - s2)
        return 1;                                                       // entirely in free space
    return 2;
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

// ------------------------------------------------------------------------------------------------------
// per-voxel arithmetic

// PinholeCamera::ProjectPoint + IsPointOnImage (OC PinholeCamera.cpp:38-45, 61-64)
__device__ __forceinline__ bool project_on_image(const CameraDev &c, float x, float y, float z, float *u, float *v)
{
    const float invZ = __fdiv_rn(1.0f, z);
    *u = __fadd_rn(__fmul_rn(__fmul_rn(c.fx, x), invZ), c.cx);
    *v = __fadd_rn(__fmul_rn(__fmul_rn(c.fy, y), invZ), c.cy);
    return *u >= 0.0f && *v >= 0.0f && *u < c.Wf && *v < c.Hf;
}

// pose.linear().transpose() * (p - t), each coefficient c0 + (c1 + c2)
__device__ __forceinline__ void to_camera(const CameraDev &c, float px, float py, float pz, float *cx, float *cy, float *cz)
{
    const float d0 = __fsub_rn(px, c.t[0]), d1 = __fsub_rn(py, c.t[1]), d2 = __fsub_rn(pz, c.t[2]);
    *cx = __fadd_rn(__fmul_rn(c.R[0], d0), __fadd_rn(__fmul_rn(c.R[3], d1), __fmul_rn(c.R[6], d2)));
    *cy = __fadd_rn(__fmul_rn(c.R[1], d0), __fadd_rn(__fmul_rn(c.R[4], d1), __fmul_rn(c.R[7], d2)));
    *cz = __fadd_rn(__fmul_rn(c.R[2], d0), __fadd_rn(__fmul_rn(c.R[5], d1), __fmul_rn(c.R[8], d2)));
}

// DistVoxel::Integrate (OC DistVoxel.h:52-60)
__device__ __forceinline__ float2 dist_integrate(float2 v, float d, float wu)
{
    const float nd = __fdiv_rn(__fadd_rn(__fmul_rn(v.y, v.x), __fmul_rn(wu, d)), __fadd_rn(wu, v.y));
    return make_float2(nd, __fadd_rn(v.y, wu));
}

// ColorVoxel::Integrate with weightUpdate = 1 (OC ColorVoxel.h:65-85) on a packed voxel (r | g << 8 | b << 16 | w << 24),
// for w < 8 (the only weights ProjectionIntegrator.h:153 lets through). The reference evaluates, per channel,
//     uint8( saturate( float(w * old + new) / float(w + 1) ) )
// in binary32. N = w * old + new <= 2040 and D = w + 1 <= 8 are exact integers, a correctly rounded quotient of integers
// that is not itself an integer stays at least 1/D - 2^-13 away from the next integer, and the cast truncates, so the
// result is exactly floor(N / D): integer arithmetic, no rounding at all. floor(N / D) = (N * ceil(2^20 / D)) >> 20 for
// N < 2048, D <= 8. tests/test_host_logic.py::test_color_integrate_integer_identity checks all 8 * 256 * 256 cases.
__constant__ unsigned cRecip20[9] = {0u, 1048576u, 524288u, 349526u, 262144u, 209716u, 174763u, 149797u, 131072u};

__device__ __forceinline__ unsigned color_integrate_packed(unsigned cv, unsigned rgb)
{
    const unsigned w = cv >> 24;
    const unsigned m = cRecip20[w + 1];
    const unsigned nr = ((w * (cv & 0xFFu) + (rgb & 0xFFu)) * m) >> 20;
    const unsigned ng = ((w * ((cv >> 8) & 0xFFu) + ((rgb >> 8) & 0xFFu)) * m) >> 20;
    const unsigned nb = ((w * ((cv >> 16) & 0xFFu) + ((rgb >> 16) & 0xFFu)) * m) >> 20;
    return nr | (ng << 8) | (nb << 16) | ((w + 1) << 24);
}

struct VoxelStats
{
    int nUpd, nCarve, nCol;
    bool updated, carvable;
};

// Per-lane terms of one 8x8x8 brick. Lane = (x = lane & 7, y sub-row = lane >> 3): the lane owns x and two y rows
// (ly, ly + 4). The camera-space coordinates are assembled from per-axis products m0 (x), m1 (y), m2 (z):
//   c_j = R(0,j)*d0 + (R(1,j)*d1 + R(2,j)*d2)   (Eigen order) -- bit-identical to the unhoisted form.
struct BrickLane
{
    float px, pyv[2];
    float m0[3], m1[2][3];
    int vx, vy0, vz0;
    float orgz;
};

__device__ __forceinline__ BrickLane brick_lane_setup(const FrameParams &fp, const DeviceMap &map, float orgx, float orgy, float orgz,
                                                      int bx, int by, int bz, int lane)
{
    const CameraDev &c = fp.cam;
    BrickLane L;
    L.vx = bx * 8 + (lane & 7);
    L.vy0 = by * 8 + (lane >> 3);
    L.vz0 = bz * 8;
    L.orgz = orgz;
    // centre_k = float(k) * res + res/2 (ChunkManager.cpp:52,61); p = centre + origin (ProjectionIntegrator.h:64)
    L.px = __fadd_rn(__fadd_rn(__fmul_rn((float)L.vx, map.res), map.half), orgx);
    const float d0 = __fsub_rn(L.px, c.t[0]);
    L.m0[0] = __fmul_rn(c.R[0], d0);
    L.m0[1] = __fmul_rn(c.R[1], d0);
    L.m0[2] = __fmul_rn(c.R[2], d0);
#pragma unroll
    for (int h = 0; h < 2; h++)
    {
        L.pyv[h] = __fadd_rn(__fadd_rn(__fmul_rn((float)(L.vy0 + 4 * h), map.res), map.half), orgy);
        const float d1 = __fsub_rn(L.pyv[h], c.t[1]);
        L.m1[h][0] = __fmul_rn(c.R[3], d1);
        L.m1[h][1] = __fmul_rn(c.R[4], d1);
        L.m1[h][2] = __fmul_rn(c.R[5], d1);
    }
    return L;
}

// One batch = the z pair (2q, 2q+1) of the brick x the lane's two y rows = four voxels per lane; the four depth gathers
// and the four 8-byte state loads are issued together. Within a batch a warp touches four 64-byte row segments per z.
//   MODE 0: existing chunk -- speculative loads of {sdf, weight} (and colour), stores only where a voxel changed
//   MODE 1: fresh chunk    -- no loads; every voxel of the batch is written (initial or integrated value)
//   MODE 2: test only      -- returns whether any of the lane's voxels falls inside the band (no traffic on the map)
template <int CS, bool COLOR_PATH, bool PER_PIXEL, int MODE>
__device__ __forceinline__ bool process_batch(const FrameParams &fp, const DeviceMap &map, const BrickLane &L, int q,
                                              float2 *dist, unsigned *col, VoxelStats *st)
{
    const CameraDev &c = fp.cam;
    float pzv[2], m2[2][3];
#pragma unroll
    for (int s = 0; s < 2; s++)
    {
        pzv[s] = __fadd_rn(__fadd_rn(__fmul_rn((float)(L.vz0 + 2 * q + s), map.res), map.half), L.orgz);
        const float d2 = __fsub_rn(pzv[s], c.t[2]);
        m2[s][0] = __fmul_rn(c.R[6], d2);
        m2[s][1] = __fmul_rn(c.R[7], d2);
        m2[s][2] = __fmul_rn(c.R[8], d2);
    }
    int pix[4], idx[4];
    float cz[4], depth[4], trunc[4];
    float2 dv[4];
    unsigned cv[4], cpx[4];
    const bool specColor = COLOR_PATH && MODE != 2 && fp.same_cam && col != nullptr;
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
        const int s = k >> 1, h = k & 1;
        const float cx = __fadd_rn(L.m0[0], __fadd_rn(L.m1[h][0], m2[s][0]));
        const float cy = __fadd_rn(L.m0[1], __fadd_rn(L.m1[h][1], m2[s][1]));
        cz[k] = __fadd_rn(L.m0[2], __fadd_rn(L.m1[h][2], m2[s][2]));
        // PinholeCamera::ProjectPoint + IsPointOnImage (PinholeCamera.cpp:38-45, 61-64); __frcp_rn is the correctly
        // rounded reciprocal, i.e. exactly 1.0f / z
        const float invZ = __frcp_rn(cz[k]);
        const float u = __fadd_rn(__fmul_rn(__fmul_rn(c.fx, cx), invZ), c.cx);
        const float v = __fadd_rn(__fmul_rn(__fmul_rn(c.fy, cy), invZ), c.cy);
        const bool on = u >= 0.0f && v >= 0.0f && u < c.Wf && v < c.Hf && !(cz[k] < 0.0f);      // ProjectionIntegrator.h:68 / :125
        pix[k] = on ? (int)u + (int)v * c.W : -1;
        idx[k] = ((L.vz0 + 2 * q + s) * CS + (L.vy0 + 4 * h)) * CS + L.vx;
    }
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
        depth[k] = pix[k] >= 0 ? __ldg(fp.depth + pix[k]) : __int_as_float(0x7fc00000);
        cpx[k] = (specColor && pix[k] >= 0) ? __ldg(fp.color_packed + pix[k]) : 0u;   // same pixel as the depth when the cameras coincide
        trunc[k] = PER_PIXEL ? (pix[k] >= 0 ? __ldg(fp.trunc_img + pix[k]) : 0.0f) : fp.trunc_param;
        if (MODE == 0)
        {
            dv[k] = dist[idx[k]];
            if (COLOR_PATH && col)
                cv[k] = col[idx[k]];
        }
        else
        {
            dv[k] = make_float2(99999.0f, 0.0f);           // Chunk::Chunk initial state (DistVoxel.cpp:29-33)
            cv[k] = 0u;
        }
    }
    bool anyHit = false;
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
        const float d = depth[k];
        int status = 0;                                     // 1 band, 2 carve candidate
        float sd = 0.0f;
        // depth path: skip depth > 50 (:74); colour path: skip NaN (:134) and depth > 100 (:141). Off-image voxels have pix < 0.
        const bool skip = (pix[k] < 0) || (COLOR_PATH ? (d != d || d > 100.0f) : (d > 50.0f));
        if (!skip)
        {
            sd = __fsub_rn(d, cz[k]);
            if (fabsf(sd) < __fadd_rn(trunc[k], fp.diag))                        // :82 / :143
                status = 1;
            else if (fp.carve && sd > __fadd_rn(trunc[k], fp.carve_dist))        // :88 / :166
                status = 2;
        }
        if (MODE == 2)
        {
            anyHit |= status == 1;
            continue;
        }
        bool wroteDist = false, wroteCol = false;
        if (status == 1)
        {
            float wu = 1.0f;                                                      // the depth path ignores the weighter (Q7)
            if (COLOR_PATH)
            {
                if (col)
                {
                    // colour first (:146-159)
                    bool onC = true;
                    int cpix = pix[k];
                    if (!fp.same_cam)
                    {
                        const int s = k >> 1, h = k & 1;
                        float ccx, ccy, ccz, cu, cvv;
                        to_camera(fp.ccam, L.px, L.pyv[h], pzv[s], &ccx, &ccy, &ccz);
                        onC = project_on_image(fp.ccam, ccx, ccy, ccz, &cu, &cvv);
                        cpix = (int)cu + (int)cvv * fp.ccam.W;
                    }
                    if (onC && (cv[k] >> 24) < 8u)                                // ProjectionIntegrator.h:153
                    {
                        const unsigned rgb = fp.same_cam ? cpx[k] : __ldg(fp.color_packed + cpix);
                        cv[k] = color_integrate_packed(cv[k], rgb);
                        wroteCol = true;
                    }
                }
                wu = PER_PIXEL ? __fdiv_rn(fp.weight, __fmul_rn(5.0f, trunc[k])) : fp.wu_const;   // ConstantWeighter.h:43-46
            }
            dv[k] = dist_integrate(dv[k], sd, wu);
            wroteDist = true;
            st->nUpd++;
            st->nCol += wroteCol;
        }
        else if (MODE == 0 && status == 2)
        {
            if (dv[k].y > 0.0f && dv[k].x < fp.sdf_carve_max)                     // weight > 0 && sdf < 1e-5 (:90 / :169)
            {
                if (COLOR_PATH && !(dv[k].y < 5.0f))
                    dv[k].y = __fsub_rn(dv[k].y, 1.0f);                           // :171-175
                else
                    dv[k] = make_float2(99999.0f, 0.0f);                          // DistVoxel::Carve -> Reset
                wroteDist = true;
                st->nCarve++;
            }
        }
        if (wroteDist)
        {
            st->updated = true;
            st->carvable |= dv[k].y > 0.0f && dv[k].x < fp.sdf_carve_max;
        }
        if (MODE == 1 || wroteDist)
            dist[idx[k]] = dv[k];
        if (col && (MODE == 1 || wroteCol))
            col[idx[k]] = cv[k];
    }
    return anyHit;
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
