// open_chisel/camera/Intrinsics.h -- facade; cf. OC/include/open_chisel/camera/Intrinsics.h:31-54.
#ifndef CHISEL_B200_INTRINSICS_H_
#define CHISEL_B200_INTRINSICS_H_
#include <memory>
#include <open_chisel/geometry/Geometry.h>

namespace chisel
{
class Intrinsics
{
  public:
    Intrinsics() : matrix(Mat3x3::Identity()) {}
    float GetFx() const { return matrix(0, 0); }
    void SetFx(float v) { matrix(0, 0) = v; }
    float GetFy() const { return matrix(1, 1); }
    void SetFy(float v) { matrix(1, 1) = v; }
    float GetCx() const { return matrix(0, 2); }
    void SetCx(float v) { matrix(0, 2) = v; }
    float GetCy() const { return matrix(1, 2); }
    void SetCy(float v) { matrix(1, 2) = v; }
    const Mat3x3 &GetMatrix() const { return matrix; }
    void SetMatrix(const Mat3x3 &m) { matrix = m; }

  protected:
    Mat3x3 matrix;
};
typedef std::shared_ptr<Intrinsics> IntrinsicsPtr;
} // namespace chisel
#endif
