// open_chisel/geometry/AABB.h -- facade; cf. OC/include/open_chisel/geometry/AABB.h:33-74.
#ifndef CHISEL_B200_AABB_H_
#define CHISEL_B200_AABB_H_
#include <memory>
#include "Geometry.h"
#include "Plane.h"

namespace chisel
{
class AABB
{
  public:
    AABB() : min(Vec3::Zero()), max(Vec3::Zero()) {}
    AABB(const Vec3 &lo, const Vec3 &hi) : min(lo), max(hi) {}
    bool Contains(const Vec3 &p) const
    {
        return p(0) >= min(0) && p(1) >= min(1) && p(2) >= min(2) && p(0) <= max(0) && p(1) <= max(1) && p(2) <= max(2);
    }
    bool Intersects(const AABB &o) const
    {
        for (int k = 0; k < 3; k++)
            if (min(k) > o.max(k) || max(k) < o.min(k))
                return false;
        return true;
    }
    Vec3 GetCenter() const { return (max + min) * 0.5f; }
    Vec3 GetExtents() const { return (max - min) * 0.5f; }
    Vec3 min, max;
};
typedef std::shared_ptr<AABB> AABBPtr;
} // namespace chisel
#endif
