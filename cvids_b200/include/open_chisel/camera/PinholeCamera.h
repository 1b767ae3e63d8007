// open_chisel/camera/PinholeCamera.h -- facade; cf. OC/include/open_chisel/camera/PinholeCamera.h:32-69 and
// OC/src/camera/PinholeCamera.cpp:38-64.
#ifndef CHISEL_B200_PINHOLECAMERA_H_
#define CHISEL_B200_PINHOLECAMERA_H_
#include <memory>
#include <chisel_b200.h>
#include <open_chisel/camera/Intrinsics.h>
#include <open_chisel/geometry/Frustum.h>
#include <open_chisel/geometry/Geometry.h>

namespace chisel
{
class PinholeCamera
{
  public:
    PinholeCamera() : width(0), height(0), nearPlane(0.05f), farPlane(5.0f) {}
    const Intrinsics &GetIntrinsics() const { return intrinsics; }
    Intrinsics &GetMutableIntrinsics() { return intrinsics; }
    void SetIntrinsics(const Intrinsics &v) { intrinsics = v; }
    int GetWidth() const { return width; }
    int GetHeight() const { return height; }
    void SetWidth(int v) { width = v; }
    void SetHeight(int v) { height = v; }
    float GetNearPlane() const { return nearPlane; }
    float GetFarPlane() const { return farPlane; }
    void SetNearPlane(float v) { nearPlane = v; }
    void SetFarPlane(float v) { farPlane = v; }

    chs_camera ToC() const
    {
        chs_camera c;
        c.fx = intrinsics.GetFx();
        c.fy = intrinsics.GetFy();
        c.cx = intrinsics.GetCx();
        c.cy = intrinsics.GetCy();
        c.width = width;
        c.height = height;
        c.near_plane = nearPlane;
        c.far_plane = farPlane;
        return c;
    }
    void SetupFrustum(const Transform &view, Frustum *frustum) const
    {
        float pose[12], corners[24], lines[72], planes[24];
        b200::PoseToArray(view, pose);
        const chs_camera c = ToC();
        chs_frustum(pose, &c, corners, lines, planes);
        frustum->SetFromArrays(corners, lines, planes);
    }
    Vec3 ProjectPoint(const Vec3 &p) const
    {
        const float invZ = 1.0f / p(2);
        return Vec3(intrinsics.GetFx() * p(0) * invZ + intrinsics.GetCx(), intrinsics.GetFy() * p(1) * invZ + intrinsics.GetCy(), p(2));
    }
    Vec3 UnprojectPoint(const Vec3 &p) const
    {
        return Vec3(p(2) * ((p(0) - intrinsics.GetCx()) / intrinsics.GetFx()), p(2) * ((p(1) - intrinsics.GetCy()) / intrinsics.GetFy()), p(2));
    }
    bool IsPointOnImage(const Vec3 &p) const { return p(0) >= 0 && p(1) >= 0 && p(0) < width && p(1) < height; }

  protected:
    Intrinsics intrinsics;
    int width, height;
    float nearPlane, farPlane;
};
typedef std::shared_ptr<PinholeCamera> PinholeCameraPtr;
} // namespace chisel
#endif
