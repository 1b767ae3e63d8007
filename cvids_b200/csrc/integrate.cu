// integrate.cu -- per-frame kernels of the TSDF integration path (sm_100a).
//
//   frame_prepare_kernel     per-pixel truncation + {min near, max far} depth-range tiles ("Hi-Z") for culling
//   chunk_candidates_kernel  enumerate the reference's candidate ID box, reproduce Frustum::Intersects, drop
//                            chunks that provably cannot change, warp-ballot compact the rest into a work list
//   integrate_kernel         projective SDF / weight / colour update, one CTA per work-list chunk
//
// Exact-arithmetic rule: everything that decides a branch or produces stored state follows SURVEY.md
// Appendix A operation by operation with __f*_rn intrinsics (never contracted into FMA; the file is also
// built with -fmad=false). Culling code is free-form float math with explicit slack, and is conservative:
// it may keep a chunk that turns out to be untouched, never drop one that would be touched.
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
// frame_prepare: one CTA per 64x64 pixel block.
__global__ void __launch_bounds__(256) frame_prepare_kernel(FrameParams fp, DeviceMap map)
{
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0)
    {
        Counters *c = map.ctr;
        c->work_count = 0;
        c->candidates = 0;
        c->n_new = 0;
        c->updated_chunks = 0;
        c->n_upd = 0;
        c->n_carve = 0;
        c->n_col = 0;
    }
    const int W = fp.cam.W, H = fp.cam.H;
    const int t = threadIdx.x;
    const int tile = t >> 2, sub = t & 3;                  // 64 tiles of 8x8, 4 threads per tile (2 rows each)
    const int tx = blockIdx.x * 8 + (tile & 7), ty = blockIdx.y * 8 + (tile >> 3);
    float lo = INFINITY, hi = -INFINITY;
    const bool perPixel = fp.trunc_img != nullptr;
    float *truncOut = (fp.trunc_kind == CHS_TRUNC_QUADRATIC || fp.trunc_kind == CHS_TRUNC_INVERSE) ? const_cast<float *>(fp.trunc_img) : nullptr;
#pragma unroll
    for (int r = 0; r < 2; r++)
    {
        const int y = ty * 8 + sub * 2 + r;
        if (y >= H)
            continue;
#pragma unroll
        for (int i = 0; i < 8; i++)
        {
            const int x = tx * 8 + i;
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
            // pixels that can never change a voxel: NaN, +-inf, beyond the cutoff (ProjectionIntegrator.h:74,134,141)
            const bool valid = (d == d) && fabsf(d) <= 3.0e38f && !(d > fp.depth_cutoff) && (tr == tr);
            if (valid)
            {
                const float band = tr + fp.diag;
                // carving reaches every z < d - (trunc + carveDist); that is inside (.., d + band) unless carveDist is very negative
                const float farExt = fp.carve ? fmaxf(band, -(tr + fp.carve_dist)) : band;
                lo = fminf(lo, d - band);
                hi = fmaxf(hi, d + farExt);
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

// Conservative depth-range test of one chunk against the Hi-Z tiles. Returns true if the chunk must be processed.
__device__ bool chunk_may_change(const FrameParams &fp, const DeviceMap &map, int idx, int idy, int idz, bool exists)
{
    const CameraDev &c = fp.cam;
    const float ext = (float)(map.cs - 1) * map.res;                   // span of voxel centres along one edge
    const float ox = (float)(map.cs * idx) * map.res + map.half - c.t[0];
    const float oy = (float)(map.cs * idy) * map.res + map.half - c.t[1];
    const float oz = (float)(map.cs * idz) * map.res + map.half - c.t[2];
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
        return false;                                                   // every centre behind the camera (ProjectionIntegrator.h:68)
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
            return false;                                               // projects entirely off the image
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
    for (int ty = y0 >> shift; ty <= (y1 >> shift); ty++)
        for (int tx = x0 >> shift; tx <= (x1 >> shift); tx++)
        {
            const float2 v = __ldg(tiles + ty * tw + tx);
            lo = fminf(lo, v.x);
            hi = fmaxf(hi, v.y);
        }
    if (!(lo <= hi))
        return false;                                                   // no valid depth pixel under the chunk
    const float s2 = 1e-3f + 1e-5f * fmaxf(fabsf(lo), fabsf(hi));
    if (zmin > hi + s2)
        return false;                                                   // entirely behind every surface it projects onto
    if (zmax < lo - s2 && !(exists && fp.carve))
        return false;                                                   // entirely in free space: only carving (of existing voxels) can act
    return true;
}

__global__ void __launch_bounds__(256) chunk_candidates_kernel(FrameParams fp, DeviceMap map)
{
    const int total = fp.n[0] * fp.n[1] * fp.n[2];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool candidate = false, keep = false;
    int x = 0, y = 0, z = 0, slot = -1;
    if (i < total)
    {
        // x outer, y, z inner -- the reference's order (ChunkManager.cpp:192-196); order does not affect the result
        const int nyz = fp.n[1] * fp.n[2];
        x = fp.lo[0] + i / nyz;
        const int r = i - (i / nyz) * nyz;
        y = fp.lo[1] + r / fp.n[2];
        z = fp.lo[2] + r % fp.n[2];
        // chunk box exactly as ChunkManager.cpp:199-201
        const float ext = __fmul_rn((float)map.cs, map.res);
        const float bx = __fmul_rn((float)(x * map.cs), map.res), by = __fmul_rn((float)(y * map.cs), map.res), bz = __fmul_rn((float)(z * map.cs), map.res);
        candidate = frustum_intersects_exact(fp, bx, by, bz, __fadd_rn(bx, ext), __fadd_rn(by, ext), __fadd_rn(bz, ext));
        if (candidate && map.world > 1)
            candidate = (owner_hash(x, y, z) % (unsigned)map.world) == (unsigned)map.rank;
        if (candidate)
        {
            slot = hash_lookup(map, pack_id(x, y, z));
            keep = chunk_may_change(fp, map, x, y, z, slot >= 0);
        }
    }
    const unsigned lane = threadIdx.x & 31;
    const unsigned candMask = __ballot_sync(0xffffffffu, candidate);
    const unsigned keepMask = __ballot_sync(0xffffffffu, keep);
    int base = 0;
    if (lane == 0)
    {
        if (candMask)
            atomicAdd(&map.ctr->candidates, __popc(candMask));
        if (keepMask)
            base = atomicAdd(&map.ctr->work_count, __popc(keepMask));
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep)
    {
        const int pos = base + __popc(keepMask & ((1u << lane) - 1));
        if (pos < fp.work_cap)
            fp.work[pos] = make_int4(x, y, z, slot);
        else
            atomicOr(&map.ctr->error_flags, kErrWorkFull);
    }
}

// ------------------------------------------------------------------------------------------------------
// per-voxel evaluation

struct VoxelEval
{
    int status;      // 0 nothing, 1 in band (integrate), 2 carve candidate
    float sd;        // surfaceDist
    float trunc;
    float px, py, pz; // voxel centre, world
};

// pose.linear().transpose() * (p - t), each coefficient c0 + (c1 + c2)
__device__ __forceinline__ void to_camera(const CameraDev &c, float px, float py, float pz, float *cx, float *cy, float *cz)
{
    const float d0 = __fsub_rn(px, c.t[0]), d1 = __fsub_rn(py, c.t[1]), d2 = __fsub_rn(pz, c.t[2]);
    *cx = __fadd_rn(__fmul_rn(c.R[0], d0), __fadd_rn(__fmul_rn(c.R[3], d1), __fmul_rn(c.R[6], d2)));
    *cy = __fadd_rn(__fmul_rn(c.R[1], d0), __fadd_rn(__fmul_rn(c.R[4], d1), __fmul_rn(c.R[7], d2)));
    *cz = __fadd_rn(__fmul_rn(c.R[2], d0), __fadd_rn(__fmul_rn(c.R[5], d1), __fmul_rn(c.R[8], d2)));
}

// PinholeCamera::ProjectPoint + IsPointOnImage (OC PinholeCamera.cpp:38-45, 61-64)
__device__ __forceinline__ bool project_on_image(const CameraDev &c, float x, float y, float z, float *u, float *v)
{
    const float invZ = __fdiv_rn(1.0f, z);
    *u = __fadd_rn(__fmul_rn(__fmul_rn(c.fx, x), invZ), c.cx);
    *v = __fadd_rn(__fmul_rn(__fmul_rn(c.fy, y), invZ), c.cy);
    return *u >= 0.0f && *v >= 0.0f && *u < c.Wf && *v < c.Hf;
}

template <bool COLOR_PATH, bool PER_PIXEL>
__device__ __forceinline__ VoxelEval eval_voxel(const FrameParams &fp, float px, float py, float pz)
{
    VoxelEval e;
    e.status = 0;
    e.px = px; e.py = py; e.pz = pz;
    float cx, cy, cz, u, v;
    to_camera(fp.cam, px, py, pz, &cx, &cy, &cz);
    if (!project_on_image(fp.cam, cx, cy, cz, &u, &v) || cz < 0.0f)
        return e;
    const int pix = (int)u + (int)v * fp.cam.W;
    const float depth = __ldg(fp.depth + pix);
    if (COLOR_PATH)
    {
        if (depth != depth || depth > 100.0f)                           // ProjectionIntegrator.h:134, :141
            return e;
    }
    else if (depth > 50.0f)                                             // :74
        return e;
    const float trunc = PER_PIXEL ? __ldg(fp.trunc_img + pix) : fp.trunc_param;
    const float sd = __fsub_rn(depth, cz);
    e.sd = sd;
    e.trunc = trunc;
    if (fabsf(sd) < __fadd_rn(trunc, fp.diag))                          // :82 / :143
        e.status = 1;
    else if (fp.carve && sd > __fadd_rn(trunc, fp.carve_dist))          // :88 / :166
        e.status = 2;
    return e;
}

// DistVoxel::Integrate (OC DistVoxel.h:52-60)
__device__ __forceinline__ float2 dist_integrate(float2 v, float d, float wu)
{
    const float nd = __fdiv_rn(__fadd_rn(__fmul_rn(v.y, v.x), __fmul_rn(wu, d)), __fadd_rn(wu, v.y));
    return make_float2(nd, __fadd_rn(v.y, wu));
}

// ColorVoxel::Integrate with weightUpdate = 1 (OC ColorVoxel.h:65-85); returns true if it wrote
__device__ __forceinline__ bool color_integrate(uchar4 *cv, unsigned char r, unsigned char g, unsigned char b)
{
    const int w = cv->w;
    if (w >= 255 - 1)
        return false;
    const float wf = (float)w, den = (float)(1 + w);
    float fr = __fdiv_rn(__fadd_rn(__fmul_rn(wf, (float)cv->x), (float)(int)r), den);
    float fg = __fdiv_rn(__fadd_rn(__fmul_rn(wf, (float)cv->y), (float)(int)g), den);
    float fb = __fdiv_rn(__fadd_rn(__fmul_rn(wf, (float)cv->z), (float)(int)b), den);
    fr = fminf(fmaxf(fr, 0.0f), 255.0f);
    fg = fminf(fmaxf(fg, 0.0f), 255.0f);
    fb = fminf(fmaxf(fb, 0.0f), 255.0f);
    *cv = make_uchar4((unsigned char)fr, (unsigned char)fg, (unsigned char)fb, (unsigned char)(w + 1));
    return true;
}

// ColorImage::At (OC ColorImage.h:61-101)
__device__ __forceinline__ void color_fetch(const FrameParams &fp, int row, int col, unsigned char *r, unsigned char *g, unsigned char *b)
{
    const uint8_t *p = fp.color + ((size_t)col + (size_t)row * fp.ccam.W) * fp.channels;
    if (fp.channels >= 3)
    {
        *b = __ldg(p);
        *g = __ldg(p + 1);
        *r = __ldg(p + 2);
    }
    else if (fp.channels == 2)
    {
        *r = __ldg(p);
        *g = *b = __ldg(p + 1);
    }
    else
        *r = *g = *b = __ldg(p);
}

// Band hit on one voxel: colour first (ProjectionIntegrator.h:146-159), then distance (:161-162 / :84-85).
template <bool COLOR_PATH>
__device__ __forceinline__ void apply_band(const FrameParams &fp, const VoxelEval &e, float2 *dv, uchar4 *cv, bool hasColorVoxel, int *nCol)
{
    float wu = 1.0f;
    if (COLOR_PATH)
    {
        float cx, cy, cz, u, v;
        to_camera(fp.ccam, e.px, e.py, e.pz, &cx, &cy, &cz);
        if (hasColorVoxel && project_on_image(fp.ccam, cx, cy, cz, &u, &v) && cv->w < 8)
        {
            unsigned char r, g, b;
            color_fetch(fp, (int)v, (int)u, &r, &g, &b);
            if (color_integrate(cv, r, g, b))
                (*nCol)++;
        }
        wu = __fdiv_rn(fp.weight, __fmul_rn(5.0f, e.trunc));            // ConstantWeighter.h:43-46
    }
    *dv = dist_integrate(*dv, e.sd, wu);
}

// One CTA (256 threads) per work-list chunk. Thread t owns voxels t, t+256, ...: x is fixed per thread
// (256 % CS == 0) and a warp reads/writes 32 consecutive voxels = 256 contiguous bytes of {sdf, weight}.
template <int CS, bool COLOR_PATH, bool PER_PIXEL>
__global__ void __launch_bounds__(256) integrate_kernel(FrameParams fp, DeviceMap map)
{
    constexpr int V = CS * CS * CS;
    constexpr int ITER = V / 256;
    __shared__ int sSlot;
    __shared__ int sRed[3][8];
    const int t = threadIdx.x;
    const int nWork = min(map.ctr->work_count, fp.work_cap);
    for (int w = blockIdx.x; w < nWork; w += gridDim.x)
    {
        const int4 item = fp.work[w];
        int slot = item.w;
        const bool isNew = slot < 0;
        // origin_k = float(CS * ID_k) * res (Chunk.cpp:43); centre_k = float(k) * res + res/2 (ChunkManager.cpp:52,61)
        const float orgx = __fmul_rn((float)(CS * item.x), map.res), orgy = __fmul_rn((float)(CS * item.y), map.res), orgz = __fmul_rn((float)(CS * item.z), map.res);
        const int vx = t % CS;
        const float px = __fadd_rn(__fadd_rn(__fmul_rn((float)vx, map.res), map.half), orgx);
        int nUpd = 0, nCarve = 0, nCol = 0;
        bool updated = false;

        if (isNew)
        {
            // Pass 1: would ProjectionIntegrator::Integrate report an update? (carving cannot touch a fresh chunk: weight 0)
            bool any = false;
#pragma unroll 4
            for (int it = 0; it < ITER; it++)
            {
                const int i = t + it * 256;
                const int vy = (i / CS) % CS, vz = i / (CS * CS);
                const float py = __fadd_rn(__fadd_rn(__fmul_rn((float)vy, map.res), map.half), orgy);
                const float pz = __fadd_rn(__fadd_rn(__fmul_rn((float)vz, map.res), map.half), orgz);
                any |= eval_voxel<COLOR_PATH, PER_PIXEL>(fp, px, py, pz).status == 1;
            }
            if (!__syncthreads_or(any))
                continue;                                               // created-and-untouched => garbage collected (Chisel.h:102-110,170-173,202-207)
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
                    hash_insert_new(map, pack_id(item.x, item.y, item.z), s);
                    atomicAdd(&map.ctr->n_new, 1);
                }
                sSlot = s;
            }
            __syncthreads();
            slot = sSlot;
            __syncthreads();
            if (slot < 0)
                continue;
        }

        float2 *dist = dist_ptr(map, slot);
        uchar4 *col = map.use_color ? color_ptr(map, slot) : nullptr;
#pragma unroll 4
        for (int it = 0; it < ITER; it++)
        {
            const int i = t + it * 256;
            const int vy = (i / CS) % CS, vz = i / (CS * CS);
            const float py = __fadd_rn(__fadd_rn(__fmul_rn((float)vy, map.res), map.half), orgy);
            const float pz = __fadd_rn(__fadd_rn(__fmul_rn((float)vz, map.res), map.half), orgz);
            const VoxelEval e = eval_voxel<COLOR_PATH, PER_PIXEL>(fp, px, py, pz);
            if (isNew)
            {
                // first write of the chunk: Chunk::Chunk initial state (DistVoxel.cpp:29-33, ColorVoxel.cpp) or the integrated value
                float2 dv = make_float2(99999.0f, 0.0f);
                uchar4 cv = make_uchar4(0, 0, 0, 0);
                if (e.status == 1)
                {
                    apply_band<COLOR_PATH>(fp, e, &dv, &cv, col != nullptr, &nCol);
                    nUpd++;
                    updated = true;
                }
                dist[i] = dv;
                if (col)
                    col[i] = cv;
            }
            else if (e.status == 1)
            {
                float2 dv = dist[i];
                uchar4 cv = make_uchar4(0, 0, 0, 0);
                const int colBefore = nCol;
                if (COLOR_PATH && col)
                    cv = col[i];
                apply_band<COLOR_PATH>(fp, e, &dv, &cv, col != nullptr, &nCol);
                dist[i] = dv;
                if (COLOR_PATH && nCol != colBefore)
                    col[i] = cv;
                nUpd++;
                updated = true;
            }
            else if (e.status == 2)
            {
                float2 dv = dist[i];
                if (dv.y > 0.0f && dv.x < fp.sdf_carve_max)             // weight > 0 && sdf < 1e-5 (:90 / :169)
                {
                    if (COLOR_PATH && !(dv.y < 5.0f))
                        dv.y = __fsub_rn(dv.y, 1.0f);                   // :171-175
                    else
                        dv = make_float2(99999.0f, 0.0f);               // DistVoxel::Carve -> Reset
                    dist[i] = dv;
                    nCarve++;
                    updated = true;
                }
            }
        }

        // block reduction of the counters and of the chunk's `updated` flag
        const unsigned lane = t & 31, warp = t >> 5;
        for (int o = 16; o > 0; o >>= 1)
        {
            nUpd += __shfl_xor_sync(0xffffffffu, nUpd, o);
            nCarve += __shfl_xor_sync(0xffffffffu, nCarve, o);
            nCol += __shfl_xor_sync(0xffffffffu, nCol, o);
        }
        if (lane == 0)
        {
            sRed[0][warp] = nUpd;
            sRed[1][warp] = nCarve;
            sRed[2][warp] = nCol;
        }
        const bool anyUpdated = __syncthreads_or(updated);
        if (t < 3)
        {
            int s = 0;
#pragma unroll
            for (int k = 0; k < 8; k++)
                s += sRed[t][k];
            if (s)
                atomicAdd(t == 0 ? &map.ctr->n_upd : (t == 1 ? &map.ctr->n_carve : &map.ctr->n_col), (unsigned long long)s);
        }
        if (anyUpdated)
        {
            // all 27 neighbour IDs become dirty, whether or not they exist (Chisel.h:89-101, 175-189)
            if (t >= 32 && t < 32 + 27)
            {
                const int k = t - 32;
                dirty_insert(map, pack_id(item.x + k / 9 - 1, item.y + (k / 3) % 3 - 1, item.z + k % 3 - 1));
            }
            if (t == 64)
                atomicAdd(&map.ctr->updated_chunks, 1);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------------
// host launchers

template <int CS>
static void launch_integrate_cs(const FrameParams &fp, const DeviceMap &map, int grid, cudaStream_t st)
{
    const bool pp = fp.trunc_img != nullptr;
    if (fp.color_path)
    {
        if (pp)
            integrate_kernel<CS, true, true><<<grid, 256, 0, st>>>(fp, map);
        else
            integrate_kernel<CS, true, false><<<grid, 256, 0, st>>>(fp, map);
    }
    else
    {
        if (pp)
            integrate_kernel<CS, false, true><<<grid, 256, 0, st>>>(fp, map);
        else
            integrate_kernel<CS, false, false><<<grid, 256, 0, st>>>(fp, map);
    }
}

void launch_frame_prepare(const FrameParams &fp, const DeviceMap &map, cudaStream_t st)
{
    dim3 grid((fp.cam.W + 63) / 64, (fp.cam.H + 63) / 64);
    frame_prepare_kernel<<<grid, 256, 0, st>>>(fp, map);
}

void launch_chunk_candidates(const FrameParams &fp, const DeviceMap &map, cudaStream_t st)
{
    const int total = fp.n[0] * fp.n[1] * fp.n[2];
    if (total <= 0)
        return;
    chunk_candidates_kernel<<<(total + 255) / 256, 256, 0, st>>>(fp, map);
}

void launch_integrate(const FrameParams &fp, const DeviceMap &map, int grid, cudaStream_t st)
{
    switch (map.cs)
    {
    case 8: launch_integrate_cs<8>(fp, map, grid, st); break;
    case 16: launch_integrate_cs<16>(fp, map, grid, st); break;
    case 32: launch_integrate_cs<32>(fp, map, grid, st); break;
    default: break;
    }
}

float host_truncation(int kind, float param, float depth) { return truncation_of(kind, param, depth); }

} // namespace chs
