// open_chisel/pointcloud/PointCloud.h -- container only (cf. OC/include/open_chisel/pointcloud/PointCloud.h:31-76). The
// point-cloud fusion mode is outside the hot path; Chisel::IntegratePointCloud reports it as unsupported.
#ifndef CHISEL_B200_POINTCLOUD_H_
#define CHISEL_B200_POINTCLOUD_H_
#include <memory>
#include <open_chisel/geometry/Geometry.h>
namespace chisel
{
class PointCloud
{
  public:
    bool HasColor() const { return !colors.empty(); }
    void Clear()
    {
        points.clear();
        colors.clear();
    }
    const Vec3List &GetPoints() const { return points; }
    Vec3List &GetMutablePoints() { return points; }
    const Vec3List &GetColors() const { return colors; }
    Vec3List &GetMutableColors() { return colors; }
    void AddPoint(const Vec3 &p) { points.push_back(p); }
    void AddColor(const Vec3 &c) { colors.push_back(c); }

  protected:
    Vec3List points, colors;
};
typedef std::shared_ptr<PointCloud> PointCloudPtr;
typedef std::shared_ptr<const PointCloud> PointCloudConstPtr;
} // namespace chisel
#endif
