// open_chisel/geometry/Plane.h -- facade; cf. OC/include/open_chisel/geometry/Plane.h:32-71.
#ifndef CHISEL_B200_PLANE_H_
#define CHISEL_B200_PLANE_H_
#include <memory>
#include "Geometry.h"

namespace chisel
{
class Plane
{
  public:
    enum class IntersectionType { Inside, Outside, Intersects };
    Plane() : normal(Vec3::Zero()), distance(0) {}
    Plane(const Vec3 &n, float d) : normal(n), distance(d) {}
    Plane(float a, float b, float c, float d) : normal(a, b, c), distance(d) {}
    float GetSignedDistance(const Vec3 &p) const { return p.dot(normal) + distance; }
    IntersectionType ClassifyPoint(const Vec3 &p) const
    {
        const float d = GetSignedDistance(p);
        return d < 0 ? IntersectionType::Inside : (d > 0 ? IntersectionType::Outside : IntersectionType::Intersects);
    }
    Vec3 normal;
    float distance;
};
typedef std::shared_ptr<Plane> PlanePtr;
} // namespace chisel
#endif
