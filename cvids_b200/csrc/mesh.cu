// mesh.cu -- incremental marching cubes over the dirty chunk set (sm_100a).
//
//   mesh_select_kernel  dirty ID list -> pool slots of the IDs that exist (warp-ballot compaction)
//   mesh_count_kernel   per chunk: one-byte halo tile in shared memory, cube configuration of every cell ONCE (memory order),
//                       written by reference rank to a per-chunk configuration row; triangle / grid counts
//   mesh_scan_kernel    exclusive prefix sums over chunks -> vertex and grid offsets
//   mesh_emit_kernel    per chunk: configuration row + SDF halo tile, block prefix sum in the REFERENCE'S cell order, one
//                       occupied cell per thread -> vertices + flat normals + grids, then gradient normals and colours per vertex
//
// Replaces ChunkManager::RecomputeMeshes / RecomputeMesh / GenerateMesh / Extract{Inside,Border}VoxelMesh
// (OC ChunkManager.cpp:91-169, 259-447), MarchingCubes::MeshCube & friends (OC MarchingCubes.h:73-146),
// ComputeNormalsFromGradients / GetSDFAndGradient / GetSDF (ChunkManager.cpp:449-499, 609-626) and
// ColorizeMesh / InterpolateColor / GetColorVoxel / Chunk::GetColorAt (ChunkManager.cpp:501-607, 628-639,
// Chunk.cpp:118-136). Emission rank = the reference's traversal order (interior z,y,x; +X face; +Y face; +Z face),
// so vertex arrays match the reference index for index. Arithmetic that produces output follows the
// reference's binary32 operation order with __f*_rn intrinsics.
#include "device_map.cuh"
#include "kernels.h"

#include "mc_table.inc"

namespace chs
{

__constant__ unsigned long long cTriPacked[256] = MC_TRI_PACKED_INIT;
__constant__ unsigned char cTriCount[256] = MC_TRI_COUNT_INIT;
__constant__ unsigned char cEdgePairs[12] = MC_EDGE_PAIRS_INIT;

__global__ void mesh_select_kernel(MeshParams mp, DeviceMap map)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int slot = -1;
    if (i < mp.n_dirty)
        slot = hash_lookup(map, mp.dirty_list[i]);          // RecomputeMesh skips IDs without a chunk (ChunkManager.cpp:94-98)
    const unsigned lane = threadIdx.x & 31;
    const unsigned mask = __ballot_sync(0xffffffffu, slot >= 0);
    int base = 0;
    if (lane == 0 && mask)
        base = atomicAdd(&map.ctr->mesh_chunks, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (slot >= 0)
        mp.mesh_slots[base + __popc(mask & ((1u << lane) - 1))] = slot;
}

// cube corner offsets, ChunkManager.cpp:67-69: x 0 1 1 0 0 1 1 0 ; y 0 0 1 1 0 0 1 1 ; z 0 0 0 0 1 1 1 1
__device__ __forceinline__ int corner_dx(int i) { return (0x66 >> i) & 1; }
__device__ __forceinline__ int corner_dy(int i) { return (0xCC >> i) & 1; }
__device__ __forceinline__ int corner_dz(int i) { return (0xF0 >> i) & 1; }

// Cell k of the reference's traversal (ChunkManager.cpp:393-441) -> voxel index of the cell's corner 0.
template <int CS>
__device__ __forceinline__ void cell_of_rank(int k, int *x, int *y, int *z)
{
    constexpr int M = CS - 1;
    constexpr int nInterior = M * M * M, nX = M * CS, nY = M * M;
    if (k < nInterior)
    {
        *z = k / (M * M);
        *y = (k / M) % M;
        *x = k % M;
    }
    else if (k < nInterior + nX)
    {
        const int j = k - nInterior;
        *x = M;
        *z = j / CS;
        *y = j % CS;
    }
    else if (k < nInterior + nX + nY)
    {
        const int j = k - nInterior - nX;
        *y = M;
        *z = j / M;
        *x = j % M;
    }
    else
    {
        const int j = k - nInterior - nX - nY;
        *z = M;
        *y = j / CS;
        *x = j % CS;
    }
}

// Halo tiles: (CS+1)^3 raw SDF values of the chunk and its +x/+y/+z neighbours plus one class byte per voxel. The weight is
// only ever compared with two thresholds, so the tile keeps the two outcomes instead of the value (5 bytes per halo voxel;
// 33^3 voxels of a 32^3 chunk fit the 227 KB of one CTA): bit 0 = meshable corner, !(weight <= 0.5) (ChunkManager.cpp:274-278,
// 311-314, 336-367); bit 1 = observed for the gradient taps, weight > 1e-12 (:476-499). 0 = the chunk does not exist.
template <int CS>
__device__ void load_halo(const DeviceMap &map, int slot, float wMin, float *tileS, unsigned char *tileW, int *nbrSlots)
{
    constexpr int H = CS + 1;
    const int t = threadIdx.x;
    const int idx = map.slot_ids[3 * slot], idy = map.slot_ids[3 * slot + 1], idz = map.slot_ids[3 * slot + 2];
    if (t < 8)
        nbrSlots[t] = (t == 0) ? slot : hash_lookup(map, pack_id(idx + (t & 1), idy + ((t >> 1) & 1), idz + (t >> 2)));
    __syncthreads();
    // linear walk over the tile, 256 voxels per round; (x, y, z) advance incrementally (no division in the loop) and four
    // rounds are batched so that four loads per thread are in flight
    constexpr int H3 = H * H * H, dX = 256 % H, dY = (256 / H) % H, dZ = (256 / H) / H;
    int x = t % H, y = (t / H) % H, z = t / (H * H);
    for (int i0 = t; i0 < H3; i0 += 1024)
    {
        float2 d[4];
        bool have[4];
#pragma unroll
        for (int r = 0; r < 4; r++)
        {
            have[r] = false;
            d[r] = make_float2(0.0f, 0.0f);
            if (i0 + 256 * r < H3)
            {
                const int s = nbrSlots[(x == CS ? 1 : 0) | (y == CS ? 2 : 0) | (z == CS ? 4 : 0)];
                if (s >= 0)
                {
                    have[r] = true;
                    d[r] = dist_ptr(map, s)[((z == CS ? 0 : z) * CS + (y == CS ? 0 : y)) * CS + (x == CS ? 0 : x)];
                }
            }
            x += dX;
            int c = x >= H;
            x -= c * H;
            y += dY + c;
            c = y >= H;
            y -= c * H;
            z += dZ + c;
        }
#pragma unroll
        for (int r = 0; r < 4; r++)
            if (i0 + 256 * r < H3)
            {
                tileS[i0 + 256 * r] = have[r] ? d[r].x : 0.0f;
                tileW[i0 + 256 * r] = have[r] ? (unsigned char)((!(d[r].y <= 0.5f) ? 1 : 0) | ((d[r].y > wMin) ? 2 : 0)) : (unsigned char)0;
            }
    }
    __syncthreads();
}


// Rank of cell (x, y, z) in the reference's traversal (ChunkManager.cpp:393-441): interior z,y,x; +X face; +Y face; +Z face.
// Inverse of cell_of_rank.
template <int CS>
__device__ __forceinline__ int rank_of_cell(int x, int y, int z)
{
    constexpr int M = CS - 1;
    constexpr int nInterior = M * M * M, nX = M * CS, nY = M * M;
    if (z == M)
        return nInterior + nX + nY + y * CS + x;
    if (x == M)
        return nInterior + z * CS + y;
    if (y == M)
        return nInterior + nX + z * M + x;
    return (z * M + y) * M + x;
}

// Pass 1 of a re-mesh: classify every cell ONCE. The halo tile holds one byte per voxel (bit 0 = meshable corner,
// !(weight <= 0.5); bit 2 = sdf < 0), cells are visited in memory order (conflict-free shared loads, shifts instead of the
// divisions cell_of_rank needs), and the cube configuration of every cell that produces triangles is written, in the
// reference's cell order, to a per-chunk scratch row that mesh_emit_kernel reads back -- it never classifies again.
template <int CS>
__global__ void __launch_bounds__(256) mesh_count_kernel(MeshParams mp, DeviceMap map)
{
    extern __shared__ unsigned char smemB[];
    constexpr int V = CS * CS * CS, H = CS + 1, H3 = H * H * H;
    unsigned char *tileB = smemB;                       // [H3]
    unsigned char *sCfg = smemB + ((H3 + 15) & ~15);    // [V] by reference rank
    __shared__ int nbr[8];
    __shared__ int red[2][8];
    const int t = threadIdx.x;
    const int n = map.ctr->mesh_chunks;
    for (int c = blockIdx.x; c < n; c += gridDim.x)
    {
        const int slot = mp.mesh_slots[c];
        const int idx = map.slot_ids[3 * slot], idy = map.slot_ids[3 * slot + 1], idz = map.slot_ids[3 * slot + 2];
        if (t < 8)
            nbr[t] = (t == 0) ? slot : hash_lookup(map, pack_id(idx + (t & 1), idy + ((t >> 1) & 1), idz + (t >> 2)));
        __syncthreads();
        {
            constexpr int dX = 256 % H, dY = (256 / H) % H, dZ = (256 / H) / H;
            int x = t % H, y = (t / H) % H, z = t / (H * H);
            for (int i0 = t; i0 < H3; i0 += 1024)
            {
                float2 d[4];
                bool have[4];
#pragma unroll
                for (int r = 0; r < 4; r++)
                {
                    have[r] = false;
                    d[r] = make_float2(0.0f, 0.0f);
                    if (i0 + 256 * r < H3)
                    {
                        const int s = nbr[(x == CS ? 1 : 0) | (y == CS ? 2 : 0) | (z == CS ? 4 : 0)];
                        if (s >= 0)
                        {
                            have[r] = true;
                            d[r] = dist_ptr(map, s)[((z == CS ? 0 : z) * CS + (y == CS ? 0 : y)) * CS + (x == CS ? 0 : x)];
                        }
                    }
                    x += dX;
                    int c = x >= H;
                    x -= c * H;
                    y += dY + c;
                    c = y >= H;
                    y -= c * H;
                    z += dZ + c;
                }
#pragma unroll
                for (int r = 0; r < 4; r++)
                    if (i0 + 256 * r < H3)
                        tileB[i0 + 256 * r] = have[r] ? (unsigned char)((!(d[r].y <= 0.5f) ? 1 : 0) | ((d[r].x < 0.0f) ? 4 : 0)) : (unsigned char)0;
            }
        }
        __syncthreads();
        int tris = 0, grids = 0;
        for (int i = t; i < V; i += 256)
        {
            const int x = i % CS, y = (i / CS) % CS, z = i / (CS * CS);
            int cfg = 0, ok = 1;
#pragma unroll
            for (int k = 0; k < 8; k++)
            {
                const int b = tileB[((z + corner_dz(k)) * H + (y + corner_dy(k))) * H + (x + corner_dx(k))];
                ok &= b;
                cfg |= ((b >> 2) & 1) << k;                      // MarchingCubes::CalculateVertexConfiguration (MarchingCubes.h:106-116)
            }
            const int nt = (ok & 1) ? cTriCount[cfg] : 0;
            sCfg[rank_of_cell<CS>(x, y, z)] = (unsigned char)(nt ? cfg : 0);     // configurations 0 and 255 have no triangle either
            tris += nt;
            grids += nt > 0;
        }
        for (int o = 16; o > 0; o >>= 1)
        {
            tris += __shfl_xor_sync(0xffffffffu, tris, o);
            grids += __shfl_xor_sync(0xffffffffu, grids, o);
        }
        if ((t & 31) == 0)
        {
            red[0][t >> 5] = tris;
            red[1][t >> 5] = grids;
        }
        __syncthreads();
        if (t < 2)
        {
            int sum = 0;
            for (int k = 0; k < 8; k++)
                sum += red[t][k];
            (t == 0 ? mp.tri_counts : mp.grid_counts)[c] = sum;
        }
        // the chunk's configuration row, coalesced
        uint4 *dst = reinterpret_cast<uint4 *>(mp.cfg_scratch + (size_t)c * V);
        const uint4 *src = reinterpret_cast<const uint4 *>(sCfg);
        for (int i = t; i < V / 16; i += 256)
            dst[i] = src[i];
        __syncthreads();
    }
}

// one block; exclusive scans of 3*tri_counts and grid_counts over the remeshed chunks
__global__ void __launch_bounds__(1024) mesh_scan_kernel(MeshParams mp, DeviceMap map)
{
    __shared__ long long warpSum[2][32];
    __shared__ long long carry[2];
    const int n = map.ctr->mesh_chunks;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0)
        carry[0] = carry[1] = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024)
    {
        const int i = base + t;
        long long v[2] = {i < n ? 3ll * mp.tri_counts[i] : 0, i < n ? (long long)mp.grid_counts[i] : 0};
        long long inc[2] = {v[0], v[1]};
        for (int o = 1; o < 32; o <<= 1)
            for (int q = 0; q < 2; q++)
            {
                const long long up = __shfl_up_sync(0xffffffffu, inc[q], o);
                if (lane >= o)
                    inc[q] += up;
            }
        if (lane == 31)
        {
            warpSum[0][warp] = inc[0];
            warpSum[1][warp] = inc[1];
        }
        __syncthreads();
        if (warp == 0)
            for (int q = 0; q < 2; q++)
            {
                long long s = warpSum[q][lane];
                for (int o = 1; o < 32; o <<= 1)
                {
                    const long long up = __shfl_up_sync(0xffffffffu, s, o);
                    if (lane >= o)
                        s += up;
                }
                warpSum[q][lane] = s;                           // inclusive over warps
            }
        __syncthreads();
        for (int q = 0; q < 2; q++)
        {
            const long long before = carry[q] + (warp ? warpSum[q][warp - 1] : 0) + inc[q] - v[q];
            if (i < n)
                (q == 0 ? mp.vert_offsets : mp.grid_offsets)[i] = before;
        }
        __syncthreads();
        if (t == 0)
        {
            carry[0] += warpSum[0][31];
            carry[1] += warpSum[1][31];
        }
        __syncthreads();
    }
    if (t == 0)
    {
        mp.vert_offsets[n] = carry[0];
        mp.grid_offsets[n] = carry[1];
        map.ctr->mesh_verts = (unsigned long long)carry[0];
        map.ctr->mesh_grids = (unsigned long long)carry[1];
    }
}

// ------------------------------------------------------------------------------------------------------
// exact helpers

struct P3
{
    float x, y, z;
};

// Lookup context of one thread: the chunk being meshed, its +neighbour slots and halo tiles (shared memory), and a
// one-entry cache of the last hash lookup. Position -> (chunk, voxel) arithmetic is the reference's, bit for bit; only the
// memory the voxel is fetched from differs (shared-memory tile when the voxel lies inside it, else the pool).
template <int CS>
struct LookupCtx
{
    int ox, oy, oz;               // ID of the chunk being meshed
    const int *nbr;               // [8] slots of chunks (ox+dx, oy+dy, oz+dz), -1 if missing
    const float *tileS;
    const unsigned char *tileW;
    float rfChunk, rfVoxel, wMin;
    unsigned long long cKey;
    int cSlot;
};

// ChunkManager::GetIDAt (OC ChunkManager.h:136-145): floor(pos * 1/(chunkSize*res)) per axis; then the chunk's slot
template <int CS>
__device__ __forceinline__ int lookup_chunk_at(const DeviceMap &map, LookupCtx<CS> &c, P3 pos, int *ix, int *iy, int *iz)
{
    const float fx = floorf(__fmul_rn(pos.x, c.rfChunk)), fy = floorf(__fmul_rn(pos.y, c.rfChunk)), fz = floorf(__fmul_rn(pos.z, c.rfChunk));
    // outside the packable range (or NaN) there is no chunk
    if (!(fabsf(fx) < (float)(kIdBias - 1)) || !(fabsf(fy) < (float)(kIdBias - 1)) || !(fabsf(fz) < (float)(kIdBias - 1)))
        return -1;
    *ix = (int)fx;
    *iy = (int)fy;
    *iz = (int)fz;
    const int dx = *ix - c.ox, dy = *iy - c.oy, dz = *iz - c.oz;
    if ((unsigned)dx < 2u && (unsigned)dy < 2u && (unsigned)dz < 2u)
        return c.nbr[dx | (dy << 1) | (dz << 2)];
    const unsigned long long key = pack_id(*ix, *iy, *iz);
    if (key != c.cKey)
    {
        c.cKey = key;
        c.cSlot = hash_lookup(map, key);
    }
    return c.cSlot;
}

// Chunk origin from its ID: float(CS * ID_k) * res (Chunk.cpp:43)
template <int CS>
__device__ __forceinline__ P3 origin_of(const DeviceMap &map, int ix, int iy, int iz)
{
    P3 o = {__fmul_rn((float)(CS * ix), map.res), __fmul_rn((float)(CS * iy), map.res), __fmul_rn((float)(CS * iz), map.res)};
    return o;
}

// ChunkManager::GetSDF (ChunkManager.cpp:476-499); Chunk::GetVoxelID(const Vec3&) (Chunk.cpp:72-86): floor(rel * (1/res))
template <int CS, bool TILED>
__device__ __forceinline__ bool get_sdf(const DeviceMap &map, LookupCtx<CS> &c, P3 pos, float *out)
{
    int ix, iy, iz;
    const int slot = lookup_chunk_at<CS>(map, c, pos, &ix, &iy, &iz);
    if (slot < 0)
        return false;
    const P3 o = origin_of<CS>(map, ix, iy, iz);
    const int vx = (int)floorf(__fmul_rn(__fsub_rn(pos.x, o.x), c.rfVoxel)), vy = (int)floorf(__fmul_rn(__fsub_rn(pos.y, o.y), c.rfVoxel)),
              vz = (int)floorf(__fmul_rn(__fsub_rn(pos.z, o.z), c.rfVoxel));
    const int id = (vz * CS + vy) * CS + vx;
    if (id < 0 || id >= CS * CS * CS)
        return false;
    const int tx = (ix - c.ox) * CS + vx, ty = (iy - c.oy) * CS + vy, tz = (iz - c.oz) * CS + vz;
    if (TILED && (unsigned)vx < (unsigned)CS && (unsigned)vy < (unsigned)CS && (unsigned)vz < (unsigned)CS && (unsigned)tx <= (unsigned)CS &&
        (unsigned)ty <= (unsigned)CS && (unsigned)tz <= (unsigned)CS)
    {
        const int ti = (tz * (CS + 1) + ty) * (CS + 1) + tx;
        if (!(c.tileW[ti] & 2))                             // weight > 1e-12
            return false;
        *out = c.tileS[ti];
        return true;
    }
    const float2 d = dist_ptr(map, slot)[id];
    if (!(d.y > c.wMin))                                    // weight > 1e-12
        return false;
    *out = d.x;
    return true;
}

// ChunkManager::GetSDFAndGradient + ComputeNormalsFromGradients (ChunkManager.cpp:449-474, 609-626)
template <int CS, bool TILED>
__device__ bool gradient_normal(const DeviceMap &map, LookupCtx<CS> &c, P3 v, P3 *n)
{
    const float r = map.res, h = __fdiv_rn(r, 2.0f);
    P3 p = {__fadd_rn(__fmul_rn(floorf(__fdiv_rn(v.x, r)), r), h), __fadd_rn(__fmul_rn(floorf(__fdiv_rn(v.y, r)), r), h),
            __fadd_rn(__fmul_rn(floorf(__fdiv_rn(v.z, r)), r), h)};
    float ctr, xp, yp, zp, xm, ym, zm;
    if (!get_sdf<CS, TILED>(map, c, p, &ctr))
        return false;
    // posf +/- Vector3f(res, 0, 0): the untouched components add 0.0f (identity for non-zero values)
    if (!get_sdf<CS, TILED>(map, c, P3{__fadd_rn(p.x, r), __fadd_rn(p.y, 0.0f), __fadd_rn(p.z, 0.0f)}, &xp)) return false;
    if (!get_sdf<CS, TILED>(map, c, P3{__fadd_rn(p.x, 0.0f), __fadd_rn(p.y, r), __fadd_rn(p.z, 0.0f)}, &yp)) return false;
    if (!get_sdf<CS, TILED>(map, c, P3{__fadd_rn(p.x, 0.0f), __fadd_rn(p.y, 0.0f), __fadd_rn(p.z, r)}, &zp)) return false;
    if (!get_sdf<CS, TILED>(map, c, P3{__fsub_rn(p.x, r), __fsub_rn(p.y, 0.0f), __fsub_rn(p.z, 0.0f)}, &xm)) return false;
    if (!get_sdf<CS, TILED>(map, c, P3{__fsub_rn(p.x, 0.0f), __fsub_rn(p.y, r), __fsub_rn(p.z, 0.0f)}, &ym)) return false;
    if (!get_sdf<CS, TILED>(map, c, P3{__fsub_rn(p.x, 0.0f), __fsub_rn(p.y, 0.0f), __fsub_rn(p.z, r)}, &zm)) return false;
    // Eigen::Vector3f(double, double, double): differences in double, narrowed once
    float gx = (float)((double)xp - (double)xm), gy = (float)((double)yp - (double)ym), gz = (float)((double)zp - (double)zm);
    float z2 = __fadd_rn(__fmul_rn(gx, gx), __fadd_rn(__fmul_rn(gy, gy), __fmul_rn(gz, gz)));
    if (z2 > 0.0f)                                          // grad->normalize()
    {
        const float s = __fsqrt_rn(z2);
        gx = __fdiv_rn(gx, s);
        gy = __fdiv_rn(gy, s);
        gz = __fdiv_rn(gz, s);
    }
    const float mag = __fsqrt_rn(__fadd_rn(__fmul_rn(gx, gx), __fadd_rn(__fmul_rn(gy, gy), __fmul_rn(gz, gz))));
    if (!((double)mag > 1e-12))
        return false;
    const float inv = __fdiv_rn(1.0f, mag);
    n->x = __fmul_rn(gx, inv);
    n->y = __fmul_rn(gy, inv);
    n->z = __fmul_rn(gz, inv);
    return true;
}

// ChunkManager::GetColorVoxel (ChunkManager.cpp:588-607)
template <int CS>
__device__ __forceinline__ bool get_color_voxel(const DeviceMap &map, LookupCtx<CS> &c, P3 pos, uchar4 *out)
{
    int ix, iy, iz;
    const int slot = lookup_chunk_at<CS>(map, c, pos, &ix, &iy, &iz);
    if (slot < 0)
        return false;
    const P3 o = origin_of<CS>(map, ix, iy, iz);
    const int vx = (int)floorf(__fmul_rn(__fsub_rn(pos.x, o.x), c.rfVoxel)), vy = (int)floorf(__fmul_rn(__fsub_rn(pos.y, o.y), c.rfVoxel)),
              vz = (int)floorf(__fmul_rn(__fsub_rn(pos.z, o.z), c.rfVoxel));
    const int id = (vz * CS + vy) * CS + vx;
    if (id < 0 || id >= CS * CS * CS)
        return false;
    *out = color_ptr(map, slot)[id];
    return true;
}

__device__ __forceinline__ float lerp_channel(float a000, float a100, float a010, float a110, float a001, float a101, float a011, float a111,
                                              float xd, float yd, float zd)
{
    const float ix = __fsub_rn(1.0f, xd), iy = __fsub_rn(1.0f, yd), iz = __fsub_rn(1.0f, zd);
    const float c00 = __fadd_rn(__fmul_rn(a000, ix), __fmul_rn(a100, xd));
    const float c10 = __fadd_rn(__fmul_rn(a010, ix), __fmul_rn(a110, xd));
    const float c01 = __fadd_rn(__fmul_rn(a001, ix), __fmul_rn(a101, xd));
    const float c11 = __fadd_rn(__fmul_rn(a011, ix), __fmul_rn(a111, xd));
    const float c0 = __fadd_rn(__fmul_rn(c00, iy), __fmul_rn(c10, yd));
    const float c1 = __fadd_rn(__fmul_rn(c01, iy), __fmul_rn(c11, yd));
    const float c = __fadd_rn(__fmul_rn(c0, iz), __fmul_rn(c1, zd));
    return __fdiv_rn(c, 255.0f);
}

// ChunkManager::InterpolateColor (ChunkManager.cpp:501-573), bug-compatible (quirk Q9: the eight lookups pass voxel
// INDICES where GetColorVoxel expects metres), with the Chunk::GetColorAt fallback (Chunk.cpp:118-136).
template <int CS>
__device__ P3 interpolate_color(const DeviceMap &map, LookupCtx<CS> &c, P3 p)
{
    const float r = map.res;
    const float fx0 = floorf(__fdiv_rn(p.x, r)), fy0 = floorf(__fdiv_rn(p.y, r)), fz0 = floorf(__fdiv_rn(p.z, r));
    // static_cast<int>(floor(.)) then back to float inside Vec3(int, int, int); exact for |value| < 2^24
    const float x0 = (float)(int)fx0, y0 = (float)(int)fy0, z0 = (float)(int)fz0;
    const float x1 = (float)((int)fx0 + 1), y1 = (float)((int)fy0 + 1), z1 = (float)((int)fz0 + 1);
    uchar4 v000, v001, v011, v111, v110, v100, v010, v101;
    bool all = get_color_voxel<CS>(map, c, P3{x0, y0, z0}, &v000);
    all = all && get_color_voxel<CS>(map, c, P3{x0, y0, z1}, &v001);
    all = all && get_color_voxel<CS>(map, c, P3{x0, y1, z1}, &v011);
    all = all && get_color_voxel<CS>(map, c, P3{x1, y1, z1}, &v111);
    all = all && get_color_voxel<CS>(map, c, P3{x1, y1, z0}, &v110);
    all = all && get_color_voxel<CS>(map, c, P3{x1, y0, z0}, &v100);
    all = all && get_color_voxel<CS>(map, c, P3{x0, y1, z0}, &v010);
    all = all && get_color_voxel<CS>(map, c, P3{x1, y0, z1}, &v101);
    if (!all)
    {
        int ix, iy, iz;
        const int slot = lookup_chunk_at<CS>(map, c, p, &ix, &iy, &iz);
        P3 zero = {0.0f, 0.0f, 0.0f};
        if (slot < 0)
            return zero;
        const P3 o = origin_of<CS>(map, ix, iy, iz);
        const float ext = __fmul_rn((float)CS, r);
        const P3 mx = {__fadd_rn(o.x, ext), __fadd_rn(o.y, ext), __fadd_rn(o.z, ext)};
        if (!(p.x >= o.x && p.y >= o.y && p.z >= o.z && p.x <= mx.x && p.y <= mx.y && p.z <= mx.z))
            return zero;
        const int cx = (int)__fdiv_rn(__fsub_rn(p.x, o.x), r), cy = (int)__fdiv_rn(__fsub_rn(p.y, o.y), r), cz = (int)__fdiv_rn(__fsub_rn(p.z, o.z), r);
        if (!(cx >= 0 && cx < CS && cy >= 0 && cy < CS && cz >= 0 && cz < CS))
            return zero;
        const uchar4 cv = color_ptr(map, slot)[(cz * CS + cy) * CS + cx];
        P3 out = {__fdiv_rn((float)cv.x, 255.0f), __fdiv_rn((float)cv.y, 255.0f), __fdiv_rn((float)cv.z, 255.0f)};
        return out;
    }
    const float xd = __fdiv_rn(__fsub_rn(p.x, x0), 1.0f), yd = __fdiv_rn(__fsub_rn(p.y, y0), 1.0f), zd = __fdiv_rn(__fsub_rn(p.z, z0), 1.0f);
    P3 out;
    out.x = lerp_channel(v000.x, v100.x, v010.x, v110.x, v001.x, v101.x, v011.x, v111.x, xd, yd, zd);
    out.y = lerp_channel(v000.y, v100.y, v010.y, v110.y, v001.y, v101.y, v011.y, v111.y, xd, yd, zd);
    out.z = lerp_channel(v000.z, v100.z, v010.z, v110.z, v001.z, v101.z, v011.z, v111.z, xd, yd, zd);
    return out;
}

// MarchingCubes::InterpolateVertex (MarchingCubes.h:134-146), incl. the v1 + 0.5*v2 branch (quirk Q8)
__device__ __forceinline__ P3 interpolate_vertex(P3 a, P3 b, float s1, float s2)
{
    const float diff = __fsub_rn(s1, s2);
    P3 o;
    if (fabsf(diff) < 1e-6f)
    {
        o.x = __fadd_rn(a.x, __fmul_rn(0.5f, b.x));
        o.y = __fadd_rn(a.y, __fmul_rn(0.5f, b.y));
        o.z = __fadd_rn(a.z, __fmul_rn(0.5f, b.z));
        return o;
    }
    const float t = __fdiv_rn(s1, diff);
    o.x = __fadd_rn(a.x, __fmul_rn(t, __fsub_rn(b.x, a.x)));
    o.y = __fadd_rn(a.y, __fmul_rn(t, __fsub_rn(b.y, a.y)));
    o.z = __fadd_rn(a.z, __fmul_rn(t, __fsub_rn(b.z, a.z)));
    return o;
}

template <int CS, bool TILED>
__global__ void __launch_bounds__(256) mesh_emit_kernel(MeshParams mp, DeviceMap map)
{
    extern __shared__ float tile[];
    __shared__ int nbr[8];
    __shared__ int warpTot[2][8];
    __shared__ int sBaseT[257], sBaseG[257];
    constexpr int V = CS * CS * CS, CPT = V / 256, H3 = (CS + 1) * (CS + 1) * (CS + 1);
    float *tileS = tile;
    unsigned char *tileW = reinterpret_cast<unsigned char *>(tile + H3);
    unsigned char *sCfg = reinterpret_cast<unsigned char *>(tile) + ((5 * H3 + 15) & ~15);   // [V] configurations by reference rank (16-byte aligned)
    const int n = map.ctr->mesh_chunks;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int c = blockIdx.x; c < n; c += gridDim.x)
    {
        const int slot = mp.mesh_slots[c];
        {
            const uint4 *src = reinterpret_cast<const uint4 *>(mp.cfg_scratch + (size_t)c * V);
            uint4 *dst = reinterpret_cast<uint4 *>(sCfg);
            for (int i = t; i < V / 16; i += 256)
                dst[i] = src[i];
        }
        load_halo<CS>(map, slot, mp.w_observed_min, tileS, tileW, nbr);              // ends with a barrier
        // exclusive prefix sums, over the reference's cell order, of triangles and occupied cells (thread t owns ranks t*CPT ..)
        int tris = 0, grids = 0;
        for (int j = 0; j < CPT; j++)
        {
            const int nt = cTriCount[sCfg[t * CPT + j]];
            tris += nt;
            grids += nt > 0;
        }
        int incT = tris, incG = grids;
        for (int o = 1; o < 32; o <<= 1)
        {
            const int a = __shfl_up_sync(0xffffffffu, incT, o), b = __shfl_up_sync(0xffffffffu, incG, o);
            if (lane >= o)
            {
                incT += a;
                incG += b;
            }
        }
        if (lane == 31)
        {
            warpTot[0][warp] = incT;
            warpTot[1][warp] = incG;
        }
        __syncthreads();
        int baseT = incT - tris, baseG = incG - grids;
        for (int w = 0; w < warp; w++)
        {
            baseT += warpTot[0][w];
            baseG += warpTot[1][w];
        }
        sBaseT[t] = baseT;
        sBaseG[t] = baseG;
        if (t == 255)
        {
            sBaseT[256] = baseT + tris;
            sBaseG[256] = baseG + grids;
        }
        __syncthreads();
        const int nActive = sBaseG[256];
        const long long vBase = mp.vert_offsets[c], vEnd = mp.vert_offsets[c + 1];
        const long long gBase = mp.grid_offsets[c];
        const int idx = map.slot_ids[3 * slot], idy = map.slot_ids[3 * slot + 1], idz = map.slot_ids[3 * slot + 2];
        const P3 org = {__fmul_rn((float)(CS * idx), map.res), __fmul_rn((float)(CS * idy), map.res), __fmul_rn((float)(CS * idz), map.res)};
        // pass 2: one occupied cell per thread and round (balanced: 1..5 triangles each), positions and flat normals written where
        // the reference's emission order puts them
        for (int g = t; g < nActive; g += 256)
        {
            // owner thread of occupied cell g: the last tau with sBaseG[tau] <= g
            int lo = 0, hi = 255;
            while (lo < hi)
            {
                const int mid = (lo + hi + 1) >> 1;
                if (sBaseG[mid] <= g)
                    lo = mid;
                else
                    hi = mid - 1;
            }
            int rank = lo * CPT, triBase = sBaseT[lo], cfg = 0;
            for (int left = g - sBaseG[lo];; rank++)
            {
                cfg = sCfg[rank];
                if (cfg)
                {
                    if (left == 0)
                        break;
                    left--;
                    triBase += cTriCount[cfg];
                }
            }
            int x, y, z;
            cell_of_rank<CS>(rank, &x, &y, &z);
            float sdf[8];
#pragma unroll
            for (int i = 0; i < 8; i++)
                sdf[i] = tileS[((z + corner_dz(i)) * (CS + 1) + (y + corner_dy(i))) * (CS + 1) + (x + corner_dx(i))];
            const int nt = cTriCount[cfg];
            long long vOut = vBase + 3ll * triBase;
            const long long gOut = gBase + g;
            // coords = centroid + origin (ChunkManager.cpp:400 etc.), centroid_k = float(k)*res + res/2 (:61)
            const P3 c0 = {__fadd_rn(__fadd_rn(__fmul_rn((float)x, map.res), map.half), org.x),
                           __fadd_rn(__fadd_rn(__fmul_rn((float)y, map.res), map.half), org.y),
                           __fadd_rn(__fadd_rn(__fmul_rn((float)z, map.res), map.half), org.z)};
            if (gOut < mp.cap_grids)
            {
                mp.grids[3 * gOut] = c0.x;
                mp.grids[3 * gOut + 1] = c0.y;
                mp.grids[3 * gOut + 2] = c0.z;
            }
            const unsigned long long row = cTriPacked[cfg];
            for (int tri = 0; tri < nt; tri++)
            {
                P3 p[3];
#pragma unroll
                for (int k = 0; k < 3; k++)
                {
                    // vertices are pushed in the order col+2, col+1, col (MarchingCubes.h:88-90)
                    const int e = (int)((row >> (4 * (3 * tri + 2 - k))) & 0xF);
                    const int a = cEdgePairs[e] >> 3, b = cEdgePairs[e] & 7;
                    // cornerCoords = coords + float(offset) * res (ChunkManager.cpp:262,278)
                    const P3 pa = {__fadd_rn(c0.x, __fmul_rn((float)corner_dx(a), map.res)), __fadd_rn(c0.y, __fmul_rn((float)corner_dy(a), map.res)),
                                   __fadd_rn(c0.z, __fmul_rn((float)corner_dz(a), map.res))};
                    const P3 pb = {__fadd_rn(c0.x, __fmul_rn((float)corner_dx(b), map.res)), __fadd_rn(c0.y, __fmul_rn((float)corner_dy(b), map.res)),
                                   __fadd_rn(c0.z, __fmul_rn((float)corner_dz(b), map.res))};
                    p[k] = interpolate_vertex(pa, pb, sdf[a], sdf[b]);
                }
                // flat normal: (p1 - p0) x (p2 - p0), normalised (MarchingCubes.h:94-99)
                const P3 u = {__fsub_rn(p[1].x, p[0].x), __fsub_rn(p[1].y, p[0].y), __fsub_rn(p[1].z, p[0].z)};
                const P3 w = {__fsub_rn(p[2].x, p[0].x), __fsub_rn(p[2].y, p[0].y), __fsub_rn(p[2].z, p[0].z)};
                P3 nf = {__fsub_rn(__fmul_rn(u.y, w.z), __fmul_rn(u.z, w.y)), __fsub_rn(__fmul_rn(u.z, w.x), __fmul_rn(u.x, w.z)),
                         __fsub_rn(__fmul_rn(u.x, w.y), __fmul_rn(u.y, w.x))};
                const float z2 = __fadd_rn(__fmul_rn(nf.x, nf.x), __fadd_rn(__fmul_rn(nf.y, nf.y), __fmul_rn(nf.z, nf.z)));
                if (z2 > 0.0f)
                {
                    const float s = __fsqrt_rn(z2);
                    nf.x = __fdiv_rn(nf.x, s);
                    nf.y = __fdiv_rn(nf.y, s);
                    nf.z = __fdiv_rn(nf.z, s);
                }
#pragma unroll
                for (int k = 0; k < 3; k++)
                {
                    if (vOut < mp.cap_vertices)
                    {
                        mp.vertices[3 * vOut] = p[k].x;
                        mp.vertices[3 * vOut + 1] = p[k].y;
                        mp.vertices[3 * vOut + 2] = p[k].z;
                        mp.normals[3 * vOut] = nf.x;
                        mp.normals[3 * vOut + 1] = nf.y;
                        mp.normals[3 * vOut + 2] = nf.z;
                    }
                    vOut++;
                }
            }
        }
        __syncthreads();
        // pass 3, vertex-parallel (balanced, coalesced): gradient normals overwrite the flat ones where all seven taps
        // exist (ChunkManager.cpp:609-626); colours (ChunkManager.cpp:628-639)
        LookupCtx<CS> ctx;
        ctx.ox = idx; ctx.oy = idy; ctx.oz = idz;
        ctx.nbr = nbr;
        ctx.tileS = tileS; ctx.tileW = tileW;
        ctx.rfChunk = __fdiv_rn(1.0f, __fmul_rn((float)CS, map.res));
        ctx.rfVoxel = __fdiv_rn(1.0f, map.res);
        ctx.wMin = mp.w_observed_min;
        ctx.cKey = kEmptyKey;
        ctx.cSlot = -1;
        const long long vLimit = vEnd < mp.cap_vertices ? vEnd : mp.cap_vertices;
        for (long long v = vBase + t; v < vLimit; v += 256)
        {
            const P3 p = {mp.vertices[3 * v], mp.vertices[3 * v + 1], mp.vertices[3 * v + 2]};
            P3 nn;
            if (gradient_normal<CS, TILED>(map, ctx, p, &nn))
            {
                mp.normals[3 * v] = nn.x;
                mp.normals[3 * v + 1] = nn.y;
                mp.normals[3 * v + 2] = nn.z;
            }
            if (mp.colors)
            {
                const P3 col = interpolate_color<CS>(map, ctx, p);
                mp.colors[3 * v] = col.x;
                mp.colors[3 * v + 1] = col.y;
                mp.colors[3 * v + 2] = col.z;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------------

void launch_mesh_select(const MeshParams &mp, const DeviceMap &map, cudaStream_t st)
{
    if (mp.n_dirty > 0)
        mesh_select_kernel<<<(mp.n_dirty + 255) / 256, 256, 0, st>>>(mp, map);
}

template <int CS>
static void mesh_launch_cs(const MeshParams &mp, const DeviceMap &map, int grid, cudaStream_t st, bool emit)
{
    constexpr size_t H3 = (size_t)(CS + 1) * (CS + 1) * (CS + 1), V = (size_t)CS * CS * CS;
    // emit: SDF tile + class-byte tile + configuration row; count: class-byte tile + configuration row
    const size_t smem = emit ? ((5 * H3 + 15) & ~(size_t)15) + V : ((H3 + 15) & ~(size_t)15) + V;
    if (emit)
    {
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(mesh_emit_kernel<CS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        mesh_emit_kernel<CS, true><<<grid, 256, smem, st>>>(mp, map);
    }
    else
    {
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(mesh_count_kernel<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        mesh_count_kernel<CS><<<grid, 256, smem, st>>>(mp, map);
    }
}

static void mesh_dispatch(const MeshParams &mp, const DeviceMap &map, int nChunks, cudaStream_t st, bool emit)
{
    const int grid = nChunks < 148 * 8 ? (nChunks > 0 ? nChunks : 1) : 148 * 8;
    switch (map.cs)
    {
    case 8: mesh_launch_cs<8>(mp, map, grid, st, emit); break;
    case 16: mesh_launch_cs<16>(mp, map, grid, st, emit); break;
    case 32: mesh_launch_cs<32>(mp, map, grid, st, emit); break;
    default: break;
    }
}

void launch_mesh_count(const MeshParams &mp, const DeviceMap &map, int nChunks, cudaStream_t st) { mesh_dispatch(mp, map, nChunks, st, false); }
void launch_mesh_scan(const MeshParams &mp, const DeviceMap &map, int, cudaStream_t st) { mesh_scan_kernel<<<1, 1024, 0, st>>>(mp, map); }
void launch_mesh_emit(const MeshParams &mp, const DeviceMap &map, int nChunks, cudaStream_t st) { mesh_dispatch(mp, map, nChunks, st, true); }

} // namespace chs
