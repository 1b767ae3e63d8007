// open_chisel/geometry/Geometry.h -- drop-in facade over libchisel_b200 (C ABI: include/chisel_b200.h).
// Same typedef names as the reference header (OC/include/open_chisel/geometry/Geometry.h:31-50).
#ifndef CHISEL_B200_GEOMETRY_H_
#define CHISEL_B200_GEOMETRY_H_

#include <vector>
#include <Eigen/Core>
#include <Eigen/Geometry>
#include <Eigen/StdVector>

namespace chisel
{
typedef Eigen::Vector2i Point2;
typedef Eigen::Vector3i Point3;
typedef Eigen::Vector2f Vec2;
typedef Eigen::Vector3f Vec3;
typedef Eigen::Vector4f Vec4;
typedef Eigen::Matrix3f Mat3x3;
typedef Eigen::Matrix4f Mat4x4;
typedef Eigen::Affine3f Transform;
typedef Eigen::Quaternionf Quaternion;

typedef std::vector<Point3, Eigen::aligned_allocator<Point3>> Point3List;
typedef std::vector<Vec2, Eigen::aligned_allocator<Vec2>> Vec2List;
typedef std::vector<Vec3, Eigen::aligned_allocator<Vec3>> Vec3List;
typedef std::vector<Vec4, Eigen::aligned_allocator<Vec4>> Vec4List;
typedef std::vector<Transform, Eigen::aligned_allocator<Transform>> TransformList;

namespace b200
{
// row-major 3x4 [R|t] of a camera->world transform, the pose layout of the C ABI
inline void PoseToArray(const Transform &t, float out[12])
{
    for (int r = 0; r < 3; r++)
    {
        for (int c = 0; c < 3; c++)
            out[r * 4 + c] = t.linear()(r, c);
        out[r * 4 + 3] = t.translation()(r);
    }
}
} // namespace b200
} // namespace chisel
#endif
