// host_geometry.h -- host-side, once-per-frame geometry of the hot path, evaluated with the reference's
// exact binary32 operation order so that the candidate box and the (ineffective but reproduced) frustum
// predicate select the same chunk IDs. Compiled by nvcc's host pass with -fmad=false (no contraction).
//
// Follows: PinholeCamera::SetupFrustum (OC/src/camera/PinholeCamera.cpp:55-59) ->
// Frustum::SetFromParams / SetFromVectors (OC/src/geometry/Frustum.cpp:143-219), Plane(a,b,c)
// (OC/src/geometry/Plane.cpp:44-52), Frustum::ComputeBoundingBox (Frustum.cpp:101-122),
// ChunkManager::GetIDAt (OC/include/open_chisel/ChunkManager.h:136-145) and the ID range of
// ChunkManager::GetChunkIDsIntersecting (OC/src/ChunkManager.cpp:182-212).
// Eigen conventions: 3-term reductions are c0 + (c1 + c2).
#pragma once

#include <cmath>
#include <limits>

#include "../../include/chisel_b200.h"

namespace chs
{

struct F3
{
    float v[3];
    float operator[](int i) const { return v[i]; }
    float &operator[](int i) { return v[i]; }
};

inline F3 f3(float a, float b, float c) { return F3{{a, b, c}}; }
inline F3 operator+(const F3 &a, const F3 &b) { return f3(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
inline F3 operator-(const F3 &a, const F3 &b) { return f3(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
inline F3 operator*(const F3 &a, float s) { return f3(a[0] * s, a[1] * s, a[2] * s); }
inline float dot3(const F3 &a, const F3 &b)
{
    const float p0 = a[0] * b[0], p1 = a[1] * b[1], p2 = a[2] * b[2];
    return p0 + (p1 + p2);
}
inline F3 cross3(const F3 &a, const F3 &b)
{
    return f3(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}

struct PlaneEq
{
    F3 n;
    float d;
};

// Plane.cpp:44-52 -- the offset uses the un-normalised cross product (quirk Q5).
inline PlaneEq plane_through(const F3 &a, const F3 &b, const F3 &c)
{
    const F3 cr = cross3(b - a, c - a);
    const float z = dot3(cr, cr);
    PlaneEq p;
    p.n = cr;
    if (z > 0.0f)
    {
        const float s = std::sqrt(z);
        p.n = f3(cr[0] / s, cr[1] / s, cr[2] / s);
    }
    p.d = -dot3(cr, a);
    return p;
}

struct FrustumGeom
{
    F3 corner[8];        // far TL, far TR, far BL, far BR, near BR, near TL, near TR, near BL (Frustum.cpp:181-188)
    PlaneEq plane[6];    // far, near, top, bottom, left, right: the order Frustum::Intersects walks (Frustum.cpp:43)
};

inline void build_frustum(const float pose[12], const chs_camera &cam, FrustumGeom *out)
{
    // SetupFrustum hands fy to both focal arguments and drops cx (quirk Q6)
    const float fxArg = cam.fy, fyArg = cam.fy, cy = cam.cy;
    const float w = static_cast<float>(cam.width), h = static_cast<float>(cam.height);
    const F3 right = f3(pose[0], pose[4], pose[8]);
    const F3 up = f3(-pose[1], -pose[5], -pose[9]);
    const F3 fwd = f3(pose[2], pose[6], pose[10]);
    const F3 eye = f3(pose[3], pose[7], pose[11]);
    const float aspect = (fxArg * w) / (fyArg * h);
    // unqualified atan2 / tan bind to the C double functions in the reference
    const float fov = static_cast<float>(std::atan2(static_cast<double>(cy), static_cast<double>(fyArg)) +
                                         std::atan2(static_cast<double>(h - cy), static_cast<double>(fyArg)));
    const float tanHalf = static_cast<float>(std::tan(static_cast<double>(fov / 2)));
    const float hFar = tanHalf * cam.far_plane, wFar = hFar * aspect;
    const float hNear = tanHalf * cam.near_plane, wNear = hNear * aspect;
    const F3 fc = eye + fwd * cam.far_plane, nc = eye + fwd * cam.near_plane;
    const F3 ftl = (fc + up * hFar) - right * wFar, ftr = (fc + up * hFar) + right * wFar;
    const F3 fbl = (fc - up * hFar) - right * wFar, fbr = (fc - up * hFar) + right * wFar;
    const F3 ntl = (nc + up * hNear) - right * wNear, ntr = (nc + up * hNear) + right * wNear;
    const F3 nbl = (nc - up * hNear) - right * wNear, nbr = (nc - up * hNear) + right * wNear;
    out->plane[0] = plane_through(ftr, ftl, fbr);   // far
    out->plane[1] = plane_through(nbl, ntl, nbr);   // near
    out->plane[2] = plane_through(ntl, ftl, ntr);   // top
    out->plane[3] = plane_through(nbr, fbl, nbl);   // bottom
    out->plane[4] = plane_through(ftl, ntl, fbl);   // left
    out->plane[5] = plane_through(ntr, ftr, nbr);   // right
    const F3 cs[8] = {ftl, ftr, fbl, fbr, nbr, ntl, ntr, nbl};
    for (int i = 0; i < 8; i++)
        out->corner[i] = cs[i];
}

// Frustum.cpp:192-218
inline void frustum_lines(const FrustumGeom &g, float lines[72])
{
    static const int order[24] = {0, 1, 3, 2, 1, 3, 2, 0, 4, 7, 6, 5, 5, 7, 6, 4, 0, 5, 1, 6, 2, 7, 3, 4};
    for (int i = 0; i < 24; i++)
        for (int k = 0; k < 3; k++)
            lines[3 * i + k] = g.corner[order[i]][k];
}

struct CandidateBox
{
    int lo[3], hi[3];    // inclusive chunk-ID range, ChunkManager.cpp:192-196
    long long count() const
    {
        long long n = 1;
        for (int k = 0; k < 3; k++)
            n *= (hi[k] >= lo[k]) ? (long long)(hi[k] - lo[k] + 1) : 0;
        return n;
    }
};

// false if the range cannot be represented (non-finite corners or IDs beyond the packable range)
inline bool candidate_box(const FrustumGeom &g, int chunkSize, float res, CandidateBox *box)
{
    const float big = std::numeric_limits<float>::max();
    float mn[3] = {big, big, big}, mx[3] = {-big, -big, -big};
    for (int i = 0; i < 8; i++)
        for (int k = 0; k < 3; k++)
        {
            const float c = g.corner[i][k];
            if (!std::isfinite(c))
                return false;
            mn[k] = std::fmin(mn[k], c);
            mx[k] = std::fmax(mx[k], c);
        }
    const float rf = 1.0f / (static_cast<float>(chunkSize) * res);     // ChunkManager.h:138-140, per instance (Q1)
    for (int k = 0; k < 3; k++)
    {
        const float a = std::floor(mn[k] * rf), b = std::floor(mx[k] * rf);
        if (!(std::fabs(a) < 1000000.0f) || !(std::fabs(b) < 1000000.0f))
            return false;
        box->lo[k] = static_cast<int>(a) - 1;          // minID - 1
        box->hi[k] = static_cast<int>(b) + 1 + 1;      // (GetIDAt(max) + 1) + 1
    }
    return true;
}

} // namespace chs
