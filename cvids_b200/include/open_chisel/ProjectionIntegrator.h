// open_chisel/ProjectionIntegrator.h -- facade: the integrator's STATE (truncator, weighter, carving), with the reference's
// setters (OC/include/open_chisel/ProjectionIntegrator.h:185-222). The per-voxel work of Integrate / IntegrateColor
// (:51-183) runs in the CUDA kernels behind chs_integrate_depth[_color]; Chisel turns this object into a chs_integrator.
#ifndef CHISEL_B200_PROJECTIONINTEGRATOR_H_
#define CHISEL_B200_PROJECTIONINTEGRATOR_H_
#include <stdexcept>
#include <vector>
#include <chisel_b200.h>
#include <open_chisel/camera/ColorImage.h>
#include <open_chisel/camera/DepthImage.h>
#include <open_chisel/camera/PinholeCamera.h>
#include <open_chisel/geometry/Geometry.h>
#include <open_chisel/truncation/Truncator.h>
#include <open_chisel/weighting/Weighter.h>

namespace chisel
{
class ProjectionIntegrator
{
  public:
    ProjectionIntegrator() : carvingDist(0), enableVoxelCarving(false) {}
    ProjectionIntegrator(const TruncatorPtr &t, const WeighterPtr &w, float carvingDist_, bool enableCarving, const Vec3List &centroids_)
        : truncator(t), weighter(w), carvingDist(carvingDist_), enableVoxelCarving(enableCarving), centroids(centroids_)
    {
    }
    const TruncatorPtr &GetTruncator() const { return truncator; }
    void SetTruncator(const TruncatorPtr &v) { truncator = v; }
    const WeighterPtr &GetWeighter() const { return weighter; }
    void SetWeighter(const WeighterPtr &v) { weighter = v; }
    float GetCarvingDist() const { return carvingDist; }
    bool IsCarvingEnabled() const { return enableVoxelCarving; }
    void SetCarvingDist(float d) { carvingDist = d; }
    void SetCarvingEnabled(bool e) { enableVoxelCarving = e; }
    void SetCentroids(const Vec3List &c) { centroids = c; }     // kept for source compatibility; the device derives them

    // C-ABI view. A truncator other than the three shipped ones is evaluated here, once per pixel, into `scratch`.
    template <class DataType>
    chs_integrator ToC(const DepthImage<DataType> &depth, std::vector<float> *scratch) const
    {
        if (!truncator)
            throw std::runtime_error("ProjectionIntegrator: no truncator set");
        chs_integrator c;
        c.trunc_kind = truncator->b200_kind();
        c.trunc_param = truncator->b200_param();
        c.trunc_per_pixel = nullptr;
        if (c.trunc_kind == CHS_TRUNC_PER_PIXEL)
        {
            const size_t n = static_cast<size_t>(depth.GetWidth()) * depth.GetHeight();
            scratch->resize(n);
            for (size_t i = 0; i < n; i++)
                (*scratch)[i] = truncator->GetTruncationDistance(static_cast<float>(depth.GetData()[i]));
            c.trunc_per_pixel = scratch->data();
        }
        c.weight = 1.0f;
        if (weighter && !weighter->b200_constant(&c.weight))
            throw std::runtime_error("ProjectionIntegrator: only ConstantWeighter is supported by the device path");
        c.carving_enabled = enableVoxelCarving ? 1 : 0;
        c.carving_dist = carvingDist;
        return c;
    }

  protected:
    TruncatorPtr truncator;
    WeighterPtr weighter;
    float carvingDist;
    bool enableVoxelCarving;
    Vec3List centroids;
};
} // namespace chisel
#endif
