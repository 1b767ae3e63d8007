/* chisel_oracle.c -- plain-C, single-threaded CPU restatement of OpenChisel's hot path
 * (projective TSDF depth(+colour) integration and per-chunk marching cubes).
 *
 * TEST INFRASTRUCTURE, NOT A PRODUCT COMPONENT (see chisel_oracle.h). It restates the reference's
 * ALGORITHM and its exact IEEE binary32 operation order; every function cites the reference lines it
 * follows (paths relative to /root/reference/OpenChisel/open_chisel). It shares no code with the CUDA
 * library: data structures here are deliberately naive (one malloc per chunk, linear-probe tables).
 *
 * Arithmetic rules: every float operation is one rounding, no FMA
 * (-ffp-contract=off, x86-64 SSE2 => FLT_EVAL_METHOD 0); Eigen 3-element reductions are
 * c0 + (c1 + c2); libm atan2/tan/sqrt/pow are the C double functions where the reference's
 * unqualified calls bind to them.
 *
 * Defined semantics where the reference is racy: Q2 every new-and-untouched
 * chunk is removed; Q3 meshes are rebuilt serially; Q4 the every-10th-call gate of
 * Chisel::UpdateMeshes is NOT applied here (orc_update_meshes always re-meshes the dirty set).
 */
#include "chisel_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "mc_table.inc"
static const uint64_t kTriPacked[256] = MC_TRI_PACKED_INIT;
static const uint8_t kEdgePairs[12] = MC_EDGE_PAIRS_INIT;

typedef struct { float x, y, z; } v3;

static v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static v3 vmul(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static v3 vdiv(v3 a, float s) { return V(a.x / s, a.y / s, a.z / s); }
/* Eigen redux order for 3 coefficients: c0 + (c1 + c2) */
static float vdot(v3 a, v3 b) { float p0 = a.x * b.x, p1 = a.y * b.y, p2 = a.z * b.z; return p0 + (p1 + p2); }
static v3 vcross(v3 a, v3 b) { return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static v3 vnormalized(v3 a) { float z = vdot(a, a); if (z > 0.0f) return vdiv(a, sqrtf(z)); return a; }

/* ------------------------------------------------------------------------------------------------ */
/* containers                                                                                        */

typedef struct
{
    int id[3];
    float *sdf;      /* [V] DistVoxel::sdf    (DistVoxel.h:73) */
    float *weight;   /* [V] DistVoxel::weight (DistVoxel.h:74) */
    uint8_t *rgbw;   /* [V][4] ColorVoxel red, green, blue, weight (ColorVoxel.h:94-97), NULL without colour */
    v3 origin;       /* Chunk.cpp:43 */
} Chunk;

typedef struct
{
    long nVerts, nGrids, capVerts, capGrids;
    float *verts, *normals, *colors, *grids; /* 3 floats each */
    int hasColors;
} Mesh;

/* open-addressing map (ChunkID -> index); stands in for std::unordered_map<ChunkID, ...> */
typedef struct
{
    int *keys;  /* 3 per slot */
    int *vals;  /* -1 = empty */
    long cap, n;
} IdMap;

static uint64_t id_hash(const int *k)
{
    /* ChunkHasher (ChunkManager.h:40-51): x*p1 ^ y*p2 ^ z*p3 in size_t, then mixed for probing */
    uint64_t h = (uint64_t)(int64_t)k[0] * 73856093u ^ (uint64_t)(int64_t)k[1] * 19349663u ^ (uint64_t)(int64_t)k[2] * 8349279u;
    h ^= h >> 29; h *= 0xbf58476d1ce4e5b9ULL; h ^= h >> 32;
    return h;
}
static void idmap_init(IdMap *m, long cap)
{
    m->cap = cap; m->n = 0;
    m->keys = (int *)malloc(sizeof(int) * 3 * cap);
    m->vals = (int *)malloc(sizeof(int) * cap);
    for (long i = 0; i < cap; i++) m->vals[i] = -1;
}
static void idmap_free(IdMap *m) { free(m->keys); free(m->vals); m->keys = m->vals = NULL; m->cap = m->n = 0; }
static int idmap_get(const IdMap *m, const int *k)
{
    long i = (long)(id_hash(k) & (uint64_t)(m->cap - 1));
    while (m->vals[i] != -1)
    {
        const int *q = m->keys + 3 * i;
        if (q[0] == k[0] && q[1] == k[1] && q[2] == k[2]) return m->vals[i];
        i = (i + 1) & (m->cap - 1);
    }
    return -1;
}
static void idmap_put(IdMap *m, const int *k, int v);
static void idmap_grow(IdMap *m)
{
    IdMap b; idmap_init(&b, m->cap * 2);
    for (long i = 0; i < m->cap; i++) if (m->vals[i] != -1) idmap_put(&b, m->keys + 3 * i, m->vals[i]);
    idmap_free(m); *m = b;
}
static void idmap_put(IdMap *m, const int *k, int v)
{
    if ((m->n + 1) * 2 > m->cap) idmap_grow(m);
    long i = (long)(id_hash(k) & (uint64_t)(m->cap - 1));
    while (m->vals[i] != -1)
    {
        const int *q = m->keys + 3 * i;
        if (q[0] == k[0] && q[1] == k[1] && q[2] == k[2]) { m->vals[i] = v; return; }
        i = (i + 1) & (m->cap - 1);
    }
    m->keys[3 * i] = k[0]; m->keys[3 * i + 1] = k[1]; m->keys[3 * i + 2] = k[2];
    m->vals[i] = v; m->n++;
}
static void idmap_clear(IdMap *m) { for (long i = 0; i < m->cap; i++) m->vals[i] = -1; m->n = 0; }

typedef struct
{
    float fx, fy, cx, cy; int W, H; float nearPlane, farPlane;
} Cam;
typedef struct { float R[3][3]; v3 t; } Pose; /* camera -> world */

typedef struct { v3 normal; float distance; } Plane;
typedef struct { v3 corners[8]; v3 lines[24]; Plane top, left, right, bottom, nearP, farP; } Frustum;

typedef struct
{
    int cs;          /* cubic chunks only (Q12: Chunk::GetVoxelID uses numVoxels(2) for the y stride) */
    int V;
    float res;
    int useColor;
    Chunk *chunks; long nChunks, capChunks;
    IdMap chunkMap;  /* ChunkManager::chunks    (ChunkManager.h:211) */
    IdMap dirty;     /* Chisel::meshesToUpdate  (Chisel.h:221-228); value 1 = true */
    Mesh *meshes; long nMeshes, capMeshes;
    IdMap meshMap;   /* ChunkManager::allMeshes (ChunkManager.h:216) */
    v3 *centroids;   /* ChunkManager::centroids (ChunkManager.cpp:50-65) */
    /* ProjectionIntegrator state (ProjectionIntegrator.h:224-229) */
    int truncKind; float truncParam; float weight; int carve; float carveDist;
    long counters[8]; long last3[3];
    /* scratch chunk for "create, integrate, erase if untouched" */
    float *scrSdf, *scrW; uint8_t *scrC;
} Map;

/* ------------------------------------------------------------------------------------------------ */
/* truncators and weighter                                                                           */

/* ConstantTruncator.h:48-51, QuadraticTruncator.h:42-45,65-67, InverseTruncator.h:42-52 */
static float truncation_distance(int kind, float param, float reading)
{
    if (kind == 0)
        return param;
    if (kind == 1)
    {
        /* const float members narrowed from double constant expressions */
        const float quadraticTerm = (float)(0.0019 * 10), linearTerm = (float)(0.00152 * 10), constantTerm = (float)(0.001504 * 10);
        /* pow(float, int) -> double; float*double -> double; whole polynomial in double; std::abs(double);
         * then * scalingFactor (float -> double), narrowed to float by the return */
        double p = (double)quadraticTerm * pow((double)reading, 2) + (double)(linearTerm * reading) + (double)constantTerm;
        return (float)(fabs(p) * (double)param);
    }
    {
        /* InverseTruncator: float inv_reading = 1.0 / reading (double division, narrowed) */
        const float BASE_LINE = (float)0.10, FOCAL = (float)471.27;
        const float DEP_SAMPLE = 1.0f / (BASE_LINE * FOCAL);
        float inv_reading = (float)(1.0 / (double)reading);
        return (DEP_SAMPLE / (inv_reading * inv_reading)) * param;
    }
}
float orc_truncation(int kind, float param, float depth) { return truncation_distance(kind, param, depth); }

/* ConstantWeighter.h:43-46: weight / (5 * truncationDist); 5 is int -> float */
static float constant_weight(float weight, float truncationDist) { return weight / (5.0f * truncationDist); }

/* ------------------------------------------------------------------------------------------------ */
/* camera and frustum                                                                                */

static Pose make_pose(const float *p)
{
    Pose r;
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) r.R[i][j] = p[i * 4 + j]; }
    r.t = V(p[3], p[7], p[11]);
    return r;
}
static Cam make_cam(const float *c)
{
    Cam k; k.fx = c[0]; k.fy = c[1]; k.cx = c[2]; k.cy = c[3]; k.W = (int)c[4]; k.H = (int)c[5]; k.nearPlane = c[6]; k.farPlane = c[7];
    return k;
}

/* Plane.cpp:44-52: normal = cross.normalized(); distance = -(cross . a) with the UN-normalised cross */
static Plane plane_from_points(v3 a, v3 b, v3 c)
{
    v3 ab = vsub(b, a), ac = vsub(c, a);
    v3 cross = vcross(ab, ac);
    Plane p; p.normal = vnormalized(cross); p.distance = -(vdot(cross, a));
    return p;
}

/* PinholeCamera.cpp:55-59 -> Frustum.cpp:143-153 (SetFromParams) -> :155-219 (SetFromVectors).
 * SetupFrustum passes fy for BOTH fx and fy, cx is ignored (Q6). */
static void setup_frustum(const Cam *cam, const Pose *view, Frustum *f)
{
    const float fx = cam->fy, fy = cam->fy, cy = cam->cy;
    const float imgWidth = (float)cam->W, imgHeight = (float)cam->H;
    v3 rightVec = V(view->R[0][0], view->R[1][0], view->R[2][0]);        /* r.col(0)  */
    v3 up = V(-view->R[0][1], -view->R[1][1], -view->R[2][1]);            /* -r.col(1) */
    v3 forward = V(view->R[0][2], view->R[1][2], view->R[2][2]);          /* r.col(2)  */
    v3 pos = view->t;
    float aspect = (fx * imgWidth) / (fy * imgHeight);
    /* atan2 binds to the double overload; the double sum is narrowed once */
    float fov = (float)(atan2((double)cy, (double)fy) + atan2((double)(imgHeight - cy), (double)fy));
    const float nearDist = cam->nearPlane, farDist = cam->farPlane;

    float angleTangent = (float)tan((double)(fov / 2));
    float heightFar = angleTangent * farDist;
    float widthFar = heightFar * aspect;
    float heightNear = angleTangent * nearDist;
    float widthNear = heightNear * aspect;
    v3 farCenter = vadd(pos, vmul(forward, farDist));
    v3 farTopLeft = vsub(vadd(farCenter, vmul(up, heightFar)), vmul(rightVec, widthFar));
    v3 farTopRight = vadd(vadd(farCenter, vmul(up, heightFar)), vmul(rightVec, widthFar));
    v3 farBotLeft = vsub(vsub(farCenter, vmul(up, heightFar)), vmul(rightVec, widthFar));
    v3 farBotRight = vadd(vsub(farCenter, vmul(up, heightFar)), vmul(rightVec, widthFar));
    v3 nearCenter = vadd(pos, vmul(forward, nearDist));
    v3 nearTopLeft = vsub(vadd(nearCenter, vmul(up, heightNear)), vmul(rightVec, widthNear));
    v3 nearTopRight = vadd(vadd(nearCenter, vmul(up, heightNear)), vmul(rightVec, widthNear));
    v3 nearBotLeft = vsub(vsub(nearCenter, vmul(up, heightNear)), vmul(rightVec, widthNear));
    v3 nearBotRight = vadd(vsub(nearCenter, vmul(up, heightNear)), vmul(rightVec, widthNear));

    f->nearP = plane_from_points(nearBotLeft, nearTopLeft, nearBotRight);
    f->farP = plane_from_points(farTopRight, farTopLeft, farBotRight);
    f->left = plane_from_points(farTopLeft, nearTopLeft, farBotLeft);
    f->right = plane_from_points(nearTopRight, farTopRight, nearBotRight);
    f->top = plane_from_points(nearTopLeft, farTopLeft, nearTopRight);
    f->bottom = plane_from_points(nearBotRight, farBotLeft, nearBotLeft);

    v3 *c = f->corners;
    c[0] = farTopLeft; c[1] = farTopRight; c[2] = farBotLeft; c[3] = farBotRight;
    c[4] = nearBotRight; c[5] = nearTopLeft; c[6] = nearTopRight; c[7] = nearBotLeft;
    static const int lineIdx[24] = {0, 1, 3, 2, 1, 3, 2, 0, 4, 7, 6, 5, 5, 7, 6, 4, 0, 5, 1, 6, 2, 7, 3, 4};
    for (int i = 0; i < 24; i++) f->lines[i] = c[lineIdx[i]];
}

/* Frustum.cpp:41-79: returns true at the FIRST plane whose far vertex has positive distance (Q5) */
static int frustum_intersects(const Frustum *f, v3 bmin, v3 bmax)
{
    const Plane *planes[6] = {&f->farP, &f->nearP, &f->top, &f->bottom, &f->left, &f->right};
    for (int i = 0; i < 6; i++)
    {
        v3 n = planes[i]->normal, a;
        a.x = (n.x < 0.0f) ? bmin.x : bmax.x;
        a.y = (n.y < 0.0f) ? bmin.y : bmax.y;
        a.z = (n.z < 0.0f) ? bmin.z : bmax.z;
        if (vdot(a, n) + planes[i]->distance > 0.0f) return 1;
    }
    return 0;
}

/* ChunkManager.h:136-145 (per-instance factors; identical to the reference's first instance, Q1) */
static void id_at(const Map *m, v3 pos, int *out)
{
    const float rf = 1.0f / ((float)m->cs * m->res);
    out[0] = (int)floorf(pos.x * rf); out[1] = (int)floorf(pos.y * rf); out[2] = (int)floorf(pos.z * rf);
}

/* ChunkManager.cpp:182-212 + Frustum.cpp:101-122. Returns the count; writes up to cap IDs
 * (x outer, y, z inner -- the reference's order). */
static long candidate_ids(const Map *m, const Frustum *f, int *out, long cap)
{
    const float big = 3.402823466e+38f;
    v3 lo = V(big, big, big), hi = V(-big, -big, -big);
    for (int i = 0; i < 8; i++)
    {
        v3 c = f->corners[i];
        lo.x = fminf(lo.x, c.x); lo.y = fminf(lo.y, c.y); lo.z = fminf(lo.z, c.z);
        hi.x = fmaxf(hi.x, c.x); hi.y = fmaxf(hi.y, c.y); hi.z = fmaxf(hi.z, c.z);
    }
    int minID[3], maxID[3];
    id_at(m, lo, minID); id_at(m, hi, maxID);
    for (int k = 0; k < 3; k++) maxID[k] += 1;
    long n = 0;
    const int cs = m->cs;
    for (int x = minID[0] - 1; x <= maxID[0] + 1; x++)
        for (int y = minID[1] - 1; y <= maxID[1] + 1; y++)
            for (int z = minID[2] - 1; z <= maxID[2] + 1; z++)
            {
                v3 bmin = vmul(V((float)(x * cs), (float)(y * cs), (float)(z * cs)), m->res);
                v3 bmax = vadd(bmin, vmul(V((float)cs, (float)cs, (float)cs), m->res));
                if (frustum_intersects(f, bmin, bmax))
                {
                    if (n < cap) { out[3 * n] = x; out[3 * n + 1] = y; out[3 * n + 2] = z; }
                    n++;
                }
            }
    return n;
}

/* ------------------------------------------------------------------------------------------------ */
/* voxels                                                                                            */

/* Chunk.cpp:43: origin_k = (numVoxels_k * ID_k) * res, int product first */
static v3 chunk_origin(const Map *m, const int *id)
{
    return V((float)(m->cs * id[0]) * m->res, (float)(m->cs * id[1]) * m->res, (float)(m->cs * id[2]) * m->res);
}

/* PinholeCamera.cpp:38-45 */
static v3 project_point(const Cam *c, v3 p)
{
    const float invZ = 1.0f / p.z;
    return V(c->fx * p.x * invZ + c->cx, c->fy * p.y * invZ + c->cy, p.z);
}
/* PinholeCamera.cpp:61-64 (width/height are ints, converted to float by the comparison) */
static int is_point_on_image(const Cam *c, v3 p) { return p.x >= 0 && p.y >= 0 && p.x < (float)c->W && p.y < (float)c->H; }

/* pose.linear().transpose() * (p - pose.translation()) */
static v3 world_to_camera(const Pose *pose, v3 p)
{
    v3 d = vsub(p, pose->t);
    v3 c;
    c.x = pose->R[0][0] * d.x + (pose->R[1][0] * d.y + pose->R[2][0] * d.z);
    c.y = pose->R[0][1] * d.x + (pose->R[1][1] * d.y + pose->R[2][1] * d.z);
    c.z = pose->R[0][2] * d.x + (pose->R[1][2] * d.y + pose->R[2][2] * d.z);
    return c;
}

/* DistVoxel.h:52-60 */
static void dist_integrate(float *sdf, float *weight, float distUpdate, float weightUpdate)
{
    float oldSDF = *sdf, oldWeight = *weight;
    float newDist = (oldWeight * oldSDF + weightUpdate * distUpdate) / (weightUpdate + oldWeight);
    *sdf = newDist;
    *weight = oldWeight + weightUpdate;
}
/* DistVoxel.h:62-72 */
static void dist_reset(float *sdf, float *weight) { *sdf = 99999; *weight = 0; }

static float saturate255(float v) { return fminf(fmaxf(v, 0.0f), 255.0f); }

/* ColorVoxel.h:65-85; returns 1 if it wrote */
static int color_integrate(uint8_t *c, uint8_t newRed, uint8_t newGreen, uint8_t newBlue, uint8_t weightUpdate)
{
    uint8_t weight = c[3];
    if (weight >= 255 - weightUpdate) return 0;
    const uint8_t in[3] = {newRed, newGreen, newBlue};
    for (int k = 0; k < 3; k++)
    {
        float oldC = (float)c[k];
        /* weight * old is float; weightUpdate * new is int, converted by the float addition */
        float upd = saturate255(((float)weight * oldC + (float)((int)weightUpdate * (int)in[k])) / (float)((int)weightUpdate + (int)weight));
        c[k] = (uint8_t)upd;
    }
    c[3] = (uint8_t)(weight + weightUpdate);
    return 1;
}

/* ColorImage.h:61-101 */
static void color_at(const uint8_t *data, int width, int channels, int row, int col, uint8_t *rgb)
{
    const int index = (col + row * width) * channels;
    switch (channels)
    {
    case 1: rgb[0] = rgb[1] = rgb[2] = data[index]; break;
    case 2: rgb[0] = data[index]; rgb[1] = rgb[2] = data[index + 1]; break;
    case 3: case 4: rgb[0] = data[index + 2]; rgb[1] = data[index + 1]; rgb[2] = data[index]; break;
    default: rgb[0] = rgb[1] = rgb[2] = 0; break; /* Color<> left uninitialised by the reference */
    }
}

/* ProjectionIntegrator.h:51-99 */
static int integrate_depth_chunk(Map *m, const float *depthImg, const Cam *cam, const Pose *pose,
                                 v3 origin, float *sdf, float *weight)
{
    const float resolution = m->res;
    const float diag = (float)(2.0 * sqrt((double)3.0f) * (double)resolution); /* :59, double expression narrowed */
    int updated = 0;
    for (int i = 0; i < m->V; i++)
    {
        m->counters[1]++;
        v3 voxelCenter = vadd(m->centroids[i], origin);
        v3 inCam = world_to_camera(pose, voxelCenter);
        v3 cameraPos = project_point(cam, inCam);
        if (!is_point_on_image(cam, cameraPos) || inCam.z < 0) continue;
        float voxelDist = inCam.z;
        float depth = depthImg[(int)cameraPos.x + (int)cameraPos.y * cam->W];    /* DepthAt(row=(int)v, col=(int)u) */
        if (depth > 50.) continue;
        float truncation = truncation_distance(m->truncKind, m->truncParam, depth);
        float surfaceDist = depth - voxelDist;
        if (fabs(surfaceDist) < truncation + diag)
        {
            dist_integrate(&sdf[i], &weight[i], surfaceDist, 1.0f);               /* weighter ignored (Q7) */
            updated = 1; m->counters[2]++;
        }
        else if (m->carve && surfaceDist > truncation + m->carveDist)
        {
            if (weight[i] > 0 && sdf[i] < 1e-5)
            {
                dist_reset(&sdf[i], &weight[i]);
                updated = 1; m->counters[3]++;
            }
        }
    }
    return updated;
}

/* ProjectionIntegrator.h:101-183 */
static int integrate_color_chunk(Map *m, const float *depthImg, const Cam *cam, const Pose *pose,
                                 const uint8_t *colorImg, int channels, const Cam *ccam, const Pose *cpose,
                                 v3 origin, float *sdf, float *weight, uint8_t *rgbw)
{
    const float resolution = m->res;
    const float resolutionDiagonal = (float)(2.0 * sqrt((double)3.0f) * (double)resolution);
    int updated = 0;
    for (int i = 0; i < m->V; i++)
    {
        m->counters[1]++;
        v3 voxelCenter = vadd(m->centroids[i], origin);
        v3 inCam = world_to_camera(pose, voxelCenter);
        v3 cameraPos = project_point(cam, inCam);
        if (!is_point_on_image(cam, cameraPos) || inCam.z < 0) continue;
        float voxelDist = inCam.z;
        float depth = depthImg[(int)cameraPos.x + (int)cameraPos.y * cam->W];
        if (isnan(depth)) continue;
        float truncation = truncation_distance(m->truncKind, m->truncParam, depth);
        float surfaceDist = depth - voxelDist;
        if (depth > 100.0f) continue;
        if (fabsf(surfaceDist) < truncation + resolutionDiagonal)
        {
            v3 inColorCam = world_to_camera(cpose, voxelCenter);
            v3 colorCameraPos = project_point(ccam, inColorCam);
            if (is_point_on_image(ccam, colorCameraPos))
            {
                uint8_t *cv = rgbw + 4 * i;
                if (cv[3] < 8)
                {
                    int r = (int)colorCameraPos.y, c = (int)colorCameraPos.x;
                    uint8_t rgb[3];
                    color_at(colorImg, ccam->W, channels, r, c, rgb);
                    if (color_integrate(cv, rgb[0], rgb[1], rgb[2], 1)) m->counters[4]++;
                }
            }
            dist_integrate(&sdf[i], &weight[i], surfaceDist, constant_weight(m->weight, truncation));
            updated = 1; m->counters[2]++;
        }
        else if (m->carve && surfaceDist > truncation + m->carveDist)
        {
            if (weight[i] > 0 && sdf[i] < 1e-5)
            {
                if (weight[i] < 5) dist_reset(&sdf[i], &weight[i]);
                else weight[i] = weight[i] - 1;
                updated = 1; m->counters[3]++;
            }
        }
    }
    return updated;
}

/* ------------------------------------------------------------------------------------------------ */
/* frame orchestration                                                                               */

static long add_chunk(Map *m, const int *id, const float *sdf, const float *w, const uint8_t *c)
{
    if (m->nChunks == m->capChunks)
    {
        m->capChunks = m->capChunks ? m->capChunks * 2 : 256;
        m->chunks = (Chunk *)realloc(m->chunks, sizeof(Chunk) * m->capChunks);
    }
    Chunk *ch = &m->chunks[m->nChunks];
    memcpy(ch->id, id, sizeof(int) * 3);
    ch->origin = chunk_origin(m, id);
    ch->sdf = (float *)malloc(sizeof(float) * m->V); memcpy(ch->sdf, sdf, sizeof(float) * m->V);
    ch->weight = (float *)malloc(sizeof(float) * m->V); memcpy(ch->weight, w, sizeof(float) * m->V);
    ch->rgbw = NULL;
    if (m->useColor) { ch->rgbw = (uint8_t *)malloc(4 * m->V); memcpy(ch->rgbw, c, 4 * m->V); }
    idmap_put(&m->chunkMap, id, (int)m->nChunks);
    return m->nChunks++;
}

static void mark_dirty27(Map *m, const int *id)
{
    /* Chisel.h:89-101 / :175-189: all 27 neighbours, whether or not they exist */
    for (int dx = -1; dx <= 1; dx++)
        for (int dy = -1; dy <= 1; dy++)
            for (int dz = -1; dz <= 1; dz++)
            {
                int k[3] = {id[0] + dx, id[1] + dy, id[2] + dz};
                idmap_put(&m->dirty, k, 1);
            }
}

/* Chisel.h:59-112 (depth) and Chisel.h:115-213 (colour, serial; Q2). A chunk that does not exist is
 * integrated into a scratch chunk initialised like Chunk::Chunk (Chunk.cpp:33-63: sdf 99999, w 0,
 * colour 0) and kept iff the integrator reports an update: the same result as the reference's
 * create -> integrate -> erase-if-untouched (Chisel.h:76-80,102-110 / :133-143,170-173,202-207). */
static void integrate_frame(Map *m, const float *depth, const Cam *cam, const Pose *pose, const uint8_t *color,
                            int channels, const Cam *ccam, const Pose *cpose, int colorPath)
{
    Frustum f;
    setup_frustum(cam, pose, &f);
    long n = candidate_ids(m, &f, NULL, 0);
    int *ids = (int *)malloc(sizeof(int) * 3 * (n ? n : 1));
    candidate_ids(m, &f, ids, n);
    memset(m->counters, 0, sizeof(m->counters));
    m->counters[0] = n;
    long nNew = 0, nGarbage = 0;
    for (long k = 0; k < n; k++)
    {
        const int *id = ids + 3 * k;
        int idx = idmap_get(&m->chunkMap, id);
        int isNew = idx < 0;
        float *sdf, *w; uint8_t *c; v3 origin;
        if (isNew)
        {
            nNew++;
            for (int i = 0; i < m->V; i++) { m->scrSdf[i] = 99999; m->scrW[i] = 0; }
            if (m->useColor) memset(m->scrC, 0, 4 * m->V);
            sdf = m->scrSdf; w = m->scrW; c = m->scrC; origin = chunk_origin(m, id);
        }
        else
        {
            Chunk *ch = &m->chunks[idx];
            sdf = ch->sdf; w = ch->weight; c = ch->rgbw; origin = ch->origin;
        }
        int updated = colorPath ? integrate_color_chunk(m, depth, cam, pose, color, channels, ccam, cpose, origin, sdf, w, c)
                                : integrate_depth_chunk(m, depth, cam, pose, origin, sdf, w);
        if (updated)
        {
            if (isNew) { add_chunk(m, id, sdf, w, c); m->counters[5]++; }
            mark_dirty27(m, id);
            m->counters[6]++;
        }
        else if (isNew) nGarbage++;
    }
    m->counters[7] = nGarbage;
    m->last3[0] = n; m->last3[1] = nNew; m->last3[2] = nGarbage;
    free(ids);
}

/* ------------------------------------------------------------------------------------------------ */
/* meshing                                                                                           */

static void mesh_clear(Mesh *ms) { ms->nVerts = ms->nGrids = 0; ms->hasColors = 0; }
static void mesh_push_vert(Mesh *ms, v3 p, v3 n)
{
    if (ms->nVerts == ms->capVerts)
    {
        ms->capVerts = ms->capVerts ? ms->capVerts * 2 : 192;
        ms->verts = (float *)realloc(ms->verts, sizeof(float) * 3 * ms->capVerts);
        ms->normals = (float *)realloc(ms->normals, sizeof(float) * 3 * ms->capVerts);
        ms->colors = (float *)realloc(ms->colors, sizeof(float) * 3 * ms->capVerts);
    }
    float *v = ms->verts + 3 * ms->nVerts, *q = ms->normals + 3 * ms->nVerts;
    v[0] = p.x; v[1] = p.y; v[2] = p.z; q[0] = n.x; q[1] = n.y; q[2] = n.z;
    ms->nVerts++;
}
static void mesh_push_grid(Mesh *ms, v3 p)
{
    if (ms->nGrids == ms->capGrids)
    {
        ms->capGrids = ms->capGrids ? ms->capGrids * 2 : 64;
        ms->grids = (float *)realloc(ms->grids, sizeof(float) * 3 * ms->capGrids);
    }
    float *g = ms->grids + 3 * ms->nGrids;
    g[0] = p.x; g[1] = p.y; g[2] = p.z;
    ms->nGrids++;
}

/* MarchingCubes.h:134-146 (Q8: the near-equal branch returns v1 + 0.5*v2, not the midpoint) */
static v3 interpolate_vertex(v3 vertex1, v3 vertex2, float sdf1, float sdf2)
{
    const float minDiff = 1e-6f;
    const float sdfDiff = sdf1 - sdf2;
    if (fabsf(sdfDiff) < minDiff) return vadd(vertex1, vmul(vertex2, 0.5f));
    const float t = sdf1 / sdfDiff;
    return vadd(vertex1, vmul(vsub(vertex2, vertex1), t));
}

/* MarchingCubes.h:73-104 (mesh overload), :106-132; returns 1 if the configuration has a triangle */
static int mesh_cube(const v3 *cornerCoords, const float *cornerSDF, Mesh *mesh)
{
    int index = 0;
    for (int i = 0; i < 8; i++) if (cornerSDF[i] < 0) index |= 1 << i;
    v3 edgeCoords[12];
    for (int i = 0; i < 12; i++)
    {
        const int e0 = kEdgePairs[i] >> 3, e1 = kEdgePairs[i] & 7;
        edgeCoords[i] = V(0, 0, 0);
        if ((cornerSDF[e0] < 0 && cornerSDF[e1] >= 0) || (cornerSDF[e0] >= 0 && cornerSDF[e1] < 0))
            edgeCoords[i] = interpolate_vertex(cornerCoords[e0], cornerCoords[e1], cornerSDF[e0], cornerSDF[e1]);
    }
    const uint64_t row = kTriPacked[index];
    int col = 0;
    while (((row >> (4 * col)) & 0xF) != 0xF)
    {
        v3 p0 = edgeCoords[(row >> (4 * (col + 2))) & 0xF];
        v3 p1 = edgeCoords[(row >> (4 * (col + 1))) & 0xF];
        v3 p2 = edgeCoords[(row >> (4 * col)) & 0xF];
        v3 n = vnormalized(vcross(vsub(p1, p0), vsub(p2, p0)));
        mesh_push_vert(mesh, p0, n); mesh_push_vert(mesh, p1, n); mesh_push_vert(mesh, p2, n);
        col += 3;
    }
    return (row & 0xF) != 0xF; /* MarchingCubes::IsOccupied (MarchingCubes.h:50-54) */
}

static const int kCubeOff[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}}; /* ChunkManager.cpp:67-69 */

static int voxel_id(const Map *m, int x, int y, int z) { return (z * m->cs + y) * m->cs + x; } /* Chunk.h:81-84 */

/* ChunkManager.cpp:259-294 (inside) and :296-379 (border) in one routine: the border variant reduces to
 * the inside one when every corner index is valid. */
static void extract_voxel_mesh(Map *m, const Chunk *chunk, int ix, int iy, int iz, v3 coords, Mesh *mesh)
{
    v3 cornerCoords[8]; float cornerSDF[8];
    const int cs = m->cs;
    for (int i = 0; i < 8; i++)
    {
        int c[3] = {ix + kCubeOff[i][0], iy + kCubeOff[i][1], iz + kCubeOff[i][2]};
        const Chunk *src = chunk;
        if (!(c[0] >= 0 && c[0] < cs && c[1] >= 0 && c[1] < cs && c[2] >= 0 && c[2] < cs))
        {
            int nid[3];
            for (int j = 0; j < 3; j++)
            {
                int off = 0;
                if (c[j] < 0) { off = -1; c[j] = cs - 1; }
                else if (c[j] >= cs) { off = 1; c[j] = 0; }
                nid[j] = chunk->id[j] + off;
            }
            int idx = idmap_get(&m->chunkMap, nid);
            if (idx < 0) return;                               /* neighbour chunk missing */
            src = &m->chunks[idx];
        }
        const int vid = voxel_id(m, c[0], c[1], c[2]);
        if (src->weight[vid] <= 0.5) return;                   /* unobserved corner */
        /* cubeCoordOffsets = cubeIndexOffsets.cast<float>() * res */
        cornerCoords[i] = vadd(coords, V((float)kCubeOff[i][0] * m->res, (float)kCubeOff[i][1] * m->res, (float)kCubeOff[i][2] * m->res));
        cornerSDF[i] = src->sdf[vid];
    }
    if (mesh_cube(cornerCoords, cornerSDF, mesh)) mesh_push_grid(mesh, coords);
}

/* ChunkManager.cpp:381-447: interior z,y,x; +X face; +Y face; +Z face */
static void generate_mesh(Map *m, const Chunk *chunk, Mesh *mesh)
{
    mesh_clear(mesh);
    const int maxX = m->cs, maxY = m->cs, maxZ = m->cs;
    int x, y, z;
#define CELL(x, y, z) extract_voxel_mesh(m, chunk, x, y, z, vadd(m->centroids[voxel_id(m, x, y, z)], chunk->origin), mesh)
    for (z = 0; z < maxZ - 1; z++) for (y = 0; y < maxY - 1; y++) for (x = 0; x < maxX - 1; x++) CELL(x, y, z);
    x = maxX - 1;
    for (z = 0; z < maxZ - 1; z++) for (y = 0; y < maxY; y++) CELL(x, y, z);
    y = maxY - 1;
    for (z = 0; z < maxZ - 1; z++) for (x = 0; x < maxX - 1; x++) CELL(x, y, z);
    z = maxZ - 1;
    for (y = 0; y < maxY; y++) for (x = 0; x < maxX; x++) CELL(x, y, z);
#undef CELL
}

/* ChunkManager::GetChunkAt (ChunkManager.h:123-134) */
static const Chunk *chunk_at(const Map *m, v3 pos)
{
    int id[3]; id_at(m, pos, id);
    int idx = idmap_get(&m->chunkMap, id);
    return idx < 0 ? NULL : &m->chunks[idx];
}

/* Chunk::GetVoxelID(const Vec3&) (Chunk.cpp:72-86): floor(rel * (1/res)) per axis, then the linear index */
static int voxel_id_of_rel(const Map *m, v3 rel)
{
    const float rf = 1.0f / m->res;
    int x = (int)floorf(rel.x * rf), y = (int)floorf(rel.y * rf), z = (int)floorf(rel.z * rf);
    return voxel_id(m, x, y, z);
}

/* ChunkManager.cpp:476-499 */
static int get_sdf(const Map *m, v3 posf, double *dist)
{
    const Chunk *chunk = chunk_at(m, posf);
    if (!chunk) return 0;
    v3 relativePos = vsub(posf, chunk->origin);
    int id = voxel_id_of_rel(m, relativePos);
    if (id >= 0 && id < m->V)
    {
        if (chunk->weight[id] > 1e-12) { *dist = chunk->sdf[id]; return 1; }
    }
    return 0;
}

/* ChunkManager.cpp:449-474 */
static int get_sdf_and_gradient(const Map *m, v3 pos, double *dist, v3 *grad)
{
    const float r = m->res;
    v3 posf = V(floorf(pos.x / r) * r + r / 2.0f, floorf(pos.y / r) * r + r / 2.0f, floorf(pos.z / r) * r + r / 2.0f);
    if (!get_sdf(m, posf, dist)) return 0;
    double ddxplus, ddyplus, ddzplus, ddxminus, ddyminus, ddzminus;
    if (!get_sdf(m, vadd(posf, V(r, 0, 0)), &ddxplus)) return 0;
    if (!get_sdf(m, vadd(posf, V(0, r, 0)), &ddyplus)) return 0;
    if (!get_sdf(m, vadd(posf, V(0, 0, r)), &ddzplus)) return 0;
    if (!get_sdf(m, vsub(posf, V(r, 0, 0)), &ddxminus)) return 0;
    if (!get_sdf(m, vsub(posf, V(0, r, 0)), &ddyminus)) return 0;
    if (!get_sdf(m, vsub(posf, V(0, 0, r)), &ddzminus)) return 0;
    /* Vector3f(double, double, double): double differences narrowed to float */
    v3 g = V((float)(ddxplus - ddxminus), (float)(ddyplus - ddyminus), (float)(ddzplus - ddzminus));
    float z = vdot(g, g);                                      /* grad->normalize() */
    if (z > 0.0f) g = vdiv(g, sqrtf(z));
    *grad = g;
    return 1;
}

/* ChunkManager.cpp:609-626 */
static void compute_normals_from_gradients(const Map *m, Mesh *mesh)
{
    for (long i = 0; i < mesh->nVerts; i++)
    {
        v3 vertex = V(mesh->verts[3 * i], mesh->verts[3 * i + 1], mesh->verts[3 * i + 2]);
        double dist; v3 grad;
        if (get_sdf_and_gradient(m, vertex, &dist, &grad))
        {
            float mag = sqrtf(vdot(grad, grad));
            if (mag > 1e-12)
            {
                v3 n = vmul(grad, 1.0f / mag);
                mesh->normals[3 * i] = n.x; mesh->normals[3 * i + 1] = n.y; mesh->normals[3 * i + 2] = n.z;
            }
        }
    }
}

/* ChunkManager.cpp:588-607 */
static const uint8_t *get_color_voxel(const Map *m, v3 pos)
{
    const Chunk *chunk = chunk_at(m, pos);
    if (!chunk) return NULL;
    v3 rel = vsub(pos, chunk->origin);
    int id = voxel_id_of_rel(m, rel);
    if (id >= 0 && id < m->V) return chunk->rgbw + 4 * id;
    return NULL;
}

/* Chunk.cpp:118-136 */
static v3 chunk_color_at(const Map *m, const Chunk *chunk, v3 pos)
{
    v3 bmin = chunk->origin;
    v3 bmax = vadd(bmin, vmul(V((float)m->cs, (float)m->cs, (float)m->cs), m->res)); /* Chunk.cpp:65-70 */
    if (pos.x >= bmin.x && pos.y >= bmin.y && pos.z >= bmin.z && pos.x <= bmax.x && pos.y <= bmax.y && pos.z <= bmax.z)
    {
        v3 chunkPos = vdiv(vsub(pos, chunk->origin), m->res);
        int cx = (int)chunkPos.x, cy = (int)chunkPos.y, cz = (int)chunkPos.z;
        if (cx >= 0 && cx < m->cs && cy >= 0 && cy < m->cs && cz >= 0 && cz < m->cs)
        {
            const uint8_t *c = chunk->rgbw + 4 * voxel_id(m, cx, cy, cz);
            const float maxVal = 255.0f;
            return V((float)c[0] / maxVal, (float)c[1] / maxVal, (float)c[2] / maxVal);
        }
    }
    return V(0, 0, 0);
}

static float trilerp(const uint8_t *v000, const uint8_t *v100, const uint8_t *v010, const uint8_t *v110,
                     const uint8_t *v001, const uint8_t *v101, const uint8_t *v011, const uint8_t *v111,
                     int ch, float xd, float yd, float zd)
{
    float c_00 = (float)v000[ch] * (1 - xd) + (float)v100[ch] * xd;
    float c_10 = (float)v010[ch] * (1 - xd) + (float)v110[ch] * xd;
    float c_01 = (float)v001[ch] * (1 - xd) + (float)v101[ch] * xd;
    float c_11 = (float)v011[ch] * (1 - xd) + (float)v111[ch] * xd;
    float c_0 = c_00 * (1 - yd) + c_10 * yd;
    float c_1 = c_01 * (1 - yd) + c_11 * yd;
    float c = c_0 * (1 - zd) + c_1 * zd;
    return c / 255.0f;
}

/* ChunkManager.cpp:501-573. Q9: the eight lookups pass voxel INDICES where GetColorVoxel expects
 * metres; reproduced as written. */
static v3 interpolate_color(const Map *m, v3 colorPos)
{
    const float x = colorPos.x, y = colorPos.y, z = colorPos.z;
    const int x_0 = (int)floorf(x / m->res), y_0 = (int)floorf(y / m->res), z_0 = (int)floorf(z / m->res);
    const int x_1 = x_0 + 1, y_1 = y_0 + 1, z_1 = z_0 + 1;
    const uint8_t *v_000 = get_color_voxel(m, V((float)x_0, (float)y_0, (float)z_0));
    const uint8_t *v_001 = get_color_voxel(m, V((float)x_0, (float)y_0, (float)z_1));
    const uint8_t *v_011 = get_color_voxel(m, V((float)x_0, (float)y_1, (float)z_1));
    const uint8_t *v_111 = get_color_voxel(m, V((float)x_1, (float)y_1, (float)z_1));
    const uint8_t *v_110 = get_color_voxel(m, V((float)x_1, (float)y_1, (float)z_0));
    const uint8_t *v_100 = get_color_voxel(m, V((float)x_1, (float)y_0, (float)z_0));
    const uint8_t *v_010 = get_color_voxel(m, V((float)x_0, (float)y_1, (float)z_0));
    const uint8_t *v_101 = get_color_voxel(m, V((float)x_1, (float)y_0, (float)z_1));
    if (!v_000 || !v_001 || !v_011 || !v_111 || !v_110 || !v_100 || !v_010 || !v_101)
    {
        const Chunk *chunk = chunk_at(m, colorPos);
        if (!chunk) return V(0, 0, 0);
        return chunk_color_at(m, chunk, colorPos);
    }
    float xd = (x - (float)x_0) / (float)(x_1 - x_0);
    float yd = (y - (float)y_0) / (float)(y_1 - y_0);
    float zd = (z - (float)z_0) / (float)(z_1 - z_0);
    return V(trilerp(v_000, v_100, v_010, v_110, v_001, v_101, v_011, v_111, 0, xd, yd, zd),
             trilerp(v_000, v_100, v_010, v_110, v_001, v_101, v_011, v_111, 1, xd, yd, zd),
             trilerp(v_000, v_100, v_010, v_110, v_001, v_101, v_011, v_111, 2, xd, yd, zd));
}

/* ChunkManager.cpp:628-639 */
static void colorize_mesh(const Map *m, Mesh *mesh)
{
    mesh->hasColors = 1;
    for (long i = 0; i < mesh->nVerts; i++)
    {
        v3 c = interpolate_color(m, V(mesh->verts[3 * i], mesh->verts[3 * i + 1], mesh->verts[3 * i + 2]));
        mesh->colors[3 * i] = c.x; mesh->colors[3 * i + 1] = c.y; mesh->colors[3 * i + 2] = c.z;
    }
}

/* ChunkManager.cpp:91-128 (Q10: an existing Mesh object is rebuilt in place even when it ends up empty;
 * a NEW mesh is published only if grids is non-empty) */
static void recompute_mesh(Map *m, const int *id)
{
    int cidx = idmap_get(&m->chunkMap, id);
    if (cidx < 0) return;
    int midx = idmap_get(&m->meshMap, id);
    Mesh tmp; memset(&tmp, 0, sizeof(tmp));
    Mesh *mesh = midx >= 0 ? &m->meshes[midx] : &tmp;
    generate_mesh(m, &m->chunks[cidx], mesh);
    if (m->useColor) colorize_mesh(m, mesh);
    compute_normals_from_gradients(m, mesh);
    if (midx < 0)
    {
        if (mesh->nGrids > 0)
        {
            if (m->nMeshes == m->capMeshes)
            {
                m->capMeshes = m->capMeshes ? m->capMeshes * 2 : 256;
                m->meshes = (Mesh *)realloc(m->meshes, sizeof(Mesh) * m->capMeshes);
            }
            m->meshes[m->nMeshes] = tmp;
            idmap_put(&m->meshMap, id, (int)m->nMeshes);
            m->nMeshes++;
        }
        else { free(tmp.verts); free(tmp.normals); free(tmp.colors); free(tmp.grids); }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* C ABI                                                                                             */

void *orc_create(int chunkSize, float resolution, int useColor)
{
    Map *m = (Map *)calloc(1, sizeof(Map));
    m->cs = chunkSize; m->V = chunkSize * chunkSize * chunkSize; m->res = resolution; m->useColor = useColor;
    idmap_init(&m->chunkMap, 1024); idmap_init(&m->dirty, 1024); idmap_init(&m->meshMap, 1024);
    /* ChunkManager::CacheCentroids (ChunkManager.cpp:50-65) */
    m->centroids = (v3 *)malloc(sizeof(v3) * m->V);
    const float half = resolution * 0.5f;
    int i = 0;
    for (int z = 0; z < chunkSize; z++) for (int y = 0; y < chunkSize; y++) for (int x = 0; x < chunkSize; x++)
        m->centroids[i++] = V((float)x * resolution + half, (float)y * resolution + half, (float)z * resolution + half);
    m->scrSdf = (float *)malloc(sizeof(float) * m->V); m->scrW = (float *)malloc(sizeof(float) * m->V);
    m->scrC = (uint8_t *)malloc(4 * m->V);
    m->truncKind = 0; m->truncParam = 4 * resolution; m->weight = 1; m->carve = 1; m->carveDist = 0.05f;
    return m;
}

static void free_contents(Map *m)
{
    for (long i = 0; i < m->nChunks; i++) { free(m->chunks[i].sdf); free(m->chunks[i].weight); free(m->chunks[i].rgbw); }
    for (long i = 0; i < m->nMeshes; i++) { free(m->meshes[i].verts); free(m->meshes[i].normals); free(m->meshes[i].colors); free(m->meshes[i].grids); }
    m->nChunks = m->nMeshes = 0;
}

void orc_destroy(void *h)
{
    Map *m = (Map *)h;
    free_contents(m);
    free(m->chunks); free(m->meshes); free(m->centroids); free(m->scrSdf); free(m->scrW); free(m->scrC);
    idmap_free(&m->chunkMap); idmap_free(&m->dirty); idmap_free(&m->meshMap);
    free(m);
}

/* Chisel::Reset (Chisel.cpp:44-48) + ChunkManager::Reset (ChunkManager.cpp:176-180) */
void orc_reset(void *h)
{
    Map *m = (Map *)h;
    free_contents(m);
    idmap_clear(&m->chunkMap); idmap_clear(&m->dirty); idmap_clear(&m->meshMap);
}

void orc_setup_integrator(void *h, int truncKind, float truncParam, float weight, int carve, float carveDist)
{
    Map *m = (Map *)h;
    m->truncKind = truncKind; m->truncParam = truncParam; m->weight = weight; m->carve = carve; m->carveDist = carveDist;
}

void orc_integrate_depth(void *h, const float *depth, int W, int H, const float *pose, const float *cam)
{
    (void)W; (void)H;
    Cam c = make_cam(cam); Pose p = make_pose(pose);
    integrate_frame((Map *)h, depth, &c, &p, NULL, 0, NULL, NULL, 0);
}

void orc_integrate_color(void *h, const float *depth, int W, int H, const float *pose, const float *cam,
                         const uint8_t *color, int cW, int cH, int channels, const float *cpose, const float *ccam, int unused)
{
    (void)W; (void)H; (void)cW; (void)cH; (void)unused;
    Cam c = make_cam(cam), cc = make_cam(ccam); Pose p = make_pose(pose), cp = make_pose(cpose);
    integrate_frame((Map *)h, depth, &c, &p, color, channels, &cc, &cp, 1);
}

int orc_candidate_ids(void *h, const float *pose, const float *cam, int *out, int cap)
{
    Cam c = make_cam(cam); Pose p = make_pose(pose); Frustum f;
    setup_frustum(&c, &p, &f);
    return (int)candidate_ids((Map *)h, &f, out, cap);
}

void orc_frustum(const float *pose, const float *cam, float *corners, float *lines, float *planes)
{
    Cam c = make_cam(cam); Pose p = make_pose(pose); Frustum f;
    setup_frustum(&c, &p, &f);
    for (int i = 0; i < 8; i++) { corners[3 * i] = f.corners[i].x; corners[3 * i + 1] = f.corners[i].y; corners[3 * i + 2] = f.corners[i].z; }
    for (int i = 0; i < 24; i++) { lines[3 * i] = f.lines[i].x; lines[3 * i + 1] = f.lines[i].y; lines[3 * i + 2] = f.lines[i].z; }
    const Plane *pl[6] = {&f.farP, &f.nearP, &f.top, &f.bottom, &f.left, &f.right};
    for (int i = 0; i < 6; i++) { planes[4 * i] = pl[i]->normal.x; planes[4 * i + 1] = pl[i]->normal.y; planes[4 * i + 2] = pl[i]->normal.z; planes[4 * i + 3] = pl[i]->distance; }
}

/* ChunkManager::RecomputeMeshes over Chisel::meshesToUpdate, serially, then clear (Chisel.cpp:55-57) */
void orc_update_meshes(void *h, int unused)
{
    (void)unused;
    Map *m = (Map *)h;
    for (long i = 0; i < m->dirty.cap; i++)
        if (m->dirty.vals[i] != -1) recompute_mesh(m, m->dirty.keys + 3 * i);
    idmap_clear(&m->dirty);
}

static int id_cmp(const void *a, const void *b)
{
    const int *p = (const int *)a, *q = (const int *)b;
    for (int k = 0; k < 3; k++) if (p[k] != q[k]) return p[k] < q[k] ? -1 : 1;
    return 0;
}
static void sorted_keys(const IdMap *m, int *out)
{
    long n = 0;
    for (long i = 0; i < m->cap; i++) if (m->vals[i] != -1) { memcpy(out + 3 * n, m->keys + 3 * i, sizeof(int) * 3); n++; }
    qsort(out, (size_t)n, sizeof(int) * 3, id_cmp);
}

int orc_num_chunks(void *h) { return (int)((Map *)h)->chunkMap.n; }
void orc_chunk_ids(void *h, int *out) { sorted_keys(&((Map *)h)->chunkMap, out); }
int orc_chunk_voxels(void *h, const int *id, float *sdf, float *weight, uint8_t *rgbw)
{
    Map *m = (Map *)h;
    int idx = idmap_get(&m->chunkMap, id);
    if (idx < 0) return 0;
    memcpy(sdf, m->chunks[idx].sdf, sizeof(float) * m->V);
    memcpy(weight, m->chunks[idx].weight, sizeof(float) * m->V);
    if (rgbw && m->chunks[idx].rgbw) memcpy(rgbw, m->chunks[idx].rgbw, 4 * m->V);
    return 1;
}
/* State injection for the import / checkpoint-resume parity tests (the reference has no such entry: its chunks are public
 * objects a caller can fill, ChunkManager.h:79-87 AddChunk + Chunk::GetVoxelsMutable). Creates the chunk if absent. */
void orc_set_chunk_voxels(void *h, const int *id, const float *sdf, const float *weight, const uint8_t *rgbw)
{
    Map *m = (Map *)h;
    int idx = idmap_get(&m->chunkMap, id);
    if (idx < 0)
    {
        uint8_t *zero = (uint8_t *)calloc(4, m->V);
        add_chunk(m, id, sdf, weight, rgbw ? rgbw : zero);
        free(zero);
        return;
    }
    memcpy(m->chunks[idx].sdf, sdf, sizeof(float) * m->V);
    memcpy(m->chunks[idx].weight, weight, sizeof(float) * m->V);
    if (rgbw && m->chunks[idx].rgbw) memcpy(m->chunks[idx].rgbw, rgbw, 4 * m->V);
}
int orc_num_dirty(void *h) { return (int)((Map *)h)->dirty.n; }
void orc_dirty_ids(void *h, int *out) { sorted_keys(&((Map *)h)->dirty, out); }
int orc_num_meshes(void *h) { return (int)((Map *)h)->meshMap.n; }
void orc_mesh_ids(void *h, int *out) { sorted_keys(&((Map *)h)->meshMap, out); }
int orc_mesh_sizes(void *h, const int *id, long *sizes)
{
    Map *m = (Map *)h;
    int idx = idmap_get(&m->meshMap, id);
    if (idx < 0) return 0;
    const Mesh *ms = &m->meshes[idx];
    sizes[0] = ms->nVerts; sizes[1] = ms->nVerts; sizes[2] = ms->hasColors ? ms->nVerts : 0; sizes[3] = ms->nGrids; sizes[4] = ms->nVerts;
    return 1;
}
int orc_mesh_data(void *h, const int *id, float *verts, float *normals, float *colors, float *grids, long *indices)
{
    Map *m = (Map *)h;
    int idx = idmap_get(&m->meshMap, id);
    if (idx < 0) return 0;
    const Mesh *ms = &m->meshes[idx];
    memcpy(verts, ms->verts, sizeof(float) * 3 * ms->nVerts);
    memcpy(normals, ms->normals, sizeof(float) * 3 * ms->nVerts);
    if (colors && ms->hasColors) memcpy(colors, ms->colors, sizeof(float) * 3 * ms->nVerts);
    memcpy(grids, ms->grids, sizeof(float) * 3 * ms->nGrids);
    if (indices) for (long i = 0; i < ms->nVerts; i++) indices[i] = i;   /* MarchingCubes.h:91-93 */
    return 1;
}
void orc_last_counts(void *h, long *out3) { memcpy(out3, ((Map *)h)->last3, sizeof(long) * 3); }
void orc_frame_counters(void *h, long *out8) { memcpy(out8, ((Map *)h)->counters, sizeof(long) * 8); }
