// open_chisel/geometry/Frustum.h -- facade. The geometry comes from chs_frustum (host-side exact restatement of
// Frustum::SetFromParams / SetFromVectors, OC/src/geometry/Frustum.cpp:143-219), so corners, lines and planes are
// bit-identical to the reference's.
#ifndef CHISEL_B200_FRUSTUM_H_
#define CHISEL_B200_FRUSTUM_H_
#include <memory>
#include "AABB.h"
#include "Geometry.h"
#include "Plane.h"

namespace chisel
{
class Frustum
{
  public:
    Frustum() {}
    const Plane &GetFarPlane() const { return planes_[0]; }
    const Plane &GetNearPlane() const { return planes_[1]; }
    const Plane &GetTopPlane() const { return planes_[2]; }
    const Plane &GetBottomPlane() const { return planes_[3]; }
    const Plane &GetLeftPlane() const { return planes_[4]; }
    const Plane &GetRightPlane() const { return planes_[5]; }
    const Vec3 *GetLines() const { return lines_; }
    const Vec3 *GetCorners() const { return corners_; }

    // Frustum.cpp:41-79, reproduced as written (true at the first plane whose far vertex lies in front, quirk Q5)
    bool Intersects(const AABB &box) const
    {
        for (int p = 0; p < 6; p++)
        {
            const Vec3 &n = planes_[p].normal;
            const Vec3 v(n(0) < 0.0f ? box.min(0) : box.max(0), n(1) < 0.0f ? box.min(1) : box.max(1), n(2) < 0.0f ? box.min(2) : box.max(2));
            if (v.dot(n) + planes_[p].distance > 0.0f)
                return true;
        }
        return false;
    }
    bool Contains(const Vec3 &point) const
    {
        for (int p = 0; p < 6; p++)
            if (planes_[p].ClassifyPoint(point) == Plane::IntersectionType::Outside)
                return false;
        return true;
    }
    void ComputeBoundingBox(AABB *box) const
    {
        Vec3 lo = corners_[0], hi = corners_[0];
        for (int i = 1; i < 8; i++)
            for (int k = 0; k < 3; k++)
            {
                lo(k) = corners_[i](k) < lo(k) ? corners_[i](k) : lo(k);
                hi(k) = corners_[i](k) > hi(k) ? corners_[i](k) : hi(k);
            }
        box->min = lo;
        box->max = hi;
    }
    // filled by PinholeCamera::SetupFrustum
    void SetFromArrays(const float corners[24], const float lines[72], const float planes[24])
    {
        for (int i = 0; i < 8; i++)
            corners_[i] = Vec3(corners[3 * i], corners[3 * i + 1], corners[3 * i + 2]);
        for (int i = 0; i < 24; i++)
            lines_[i] = Vec3(lines[3 * i], lines[3 * i + 1], lines[3 * i + 2]);
        for (int p = 0; p < 6; p++)
            planes_[p] = Plane(planes[4 * p], planes[4 * p + 1], planes[4 * p + 2], planes[4 * p + 3]);
    }

  protected:
    Vec3 corners_[8];
    Vec3 lines_[24];
    Plane planes_[6]; // far, near, top, bottom, left, right
};
typedef std::shared_ptr<Frustum> FrustumPtr;
} // namespace chisel
#endif
