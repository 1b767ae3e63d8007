// open_chisel/Chisel.h -- drop-in facade of chisel::Chisel (OC/include/open_chisel/Chisel.h:37-229, OC/src/Chisel.cpp) over
// libchisel_b200. Source compatible with the calls chisel_ros makes (CR/src/ChiselServer.cpp:189-200, 495-525, 346, 554,
// 721, 729-737): swap the include path, link -lchisel_b200, nothing else changes.
//
// Semantics kept: UpdateMeshes re-meshes on every 10th call only (Chisel.cpp:50-59; counter per instance instead of
// process-global, quirk Q4); every new-and-untouched chunk is garbage collected (the deterministic reading of quirk Q2).
// Not kept: the reference's printf chatter and its per-frame PrintMemoryStatistics scan (Chisel.h:62,109-111).
//
// Frame batching (extension, off by default; SetFrameBatching(n) or the environment variable CHISEL_B200_BATCH=n, n <= 16):
// IntegrateDepthScan[Color] copies the frame into a queue and returns; the queue goes to the device as ONE
// chs_integrate_batch call (fused multi-frame kernels, bit-identical to frame-by-frame integration) when it holds n frames,
// when the integrator / camera settings change, or when anything reads the device map -- UpdateMeshes on a call that actually
// re-meshes (every 10th), GetMeshesToUpdate, HasChunk / GetChunk / GetChunks, Reset. chisel_ros calls UpdateMeshes after
// every frame but re-meshes on every 10th call only, so with n = 10 a live stream runs entirely through the fused path.
#ifndef CHISEL_B200_CHISEL_H_
#define CHISEL_B200_CHISEL_H_
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include <chisel_b200.h>
#include <open_chisel/ChunkManager.h>
#include <open_chisel/ProjectionIntegrator.h>
#include <open_chisel/camera/ColorImage.h>
#include <open_chisel/camera/DepthImage.h>
#include <open_chisel/camera/PinholeCamera.h>
#include <open_chisel/geometry/Frustum.h>
#include <open_chisel/geometry/Geometry.h>
#include <open_chisel/io/PLY.h>
#include <open_chisel/pointcloud/PointCloud.h>

namespace chisel
{
class Chisel
{
  public:
    Chisel() : updateCalls(0), dirtyVersion(-1), batchFrames(1) {}
    Chisel(const Eigen::Vector3i &chunkSize, float voxelResolution, bool useColor) : chunkManager(chunkSize, voxelResolution, useColor), updateCalls(0), dirtyVersion(-1), batchFrames(1)
    {
        InitBatching();
    }
    // multi-GPU / explicit-stream variant (not in the reference): this instance keeps the chunk IDs it owns
    Chisel(const Eigen::Vector3i &chunkSize, float voxelResolution, bool useColor, int device, int rank, int world, void *stream = nullptr)
        : chunkManager(chunkSize, voxelResolution, useColor, device, rank, world, stream), updateCalls(0), dirtyVersion(-1), batchFrames(1)
    {
        InitBatching();
    }
    Chisel(const Chisel &) = delete;                 // the flush hook installed in the ChunkManager points at this object
    Chisel &operator=(const Chisel &) = delete;
    virtual ~Chisel()
    {
        chunkManager.SetBeforeDeviceRead(std::function<void()>());
        chs_host_free(ring);
    }

    const ChunkManager &GetChunkManager() const { return chunkManager; }
    ChunkManager &GetMutableChunkManager() { return chunkManager; }
    void SetChunkManager(const ChunkManager &manager)
    {
        Flush();
        chunkManager = manager;
        chunkManager.SetBeforeDeviceRead([this]() { Flush(); });
    }

    // Queue up to n frames (1 = off: every frame goes to the device at once). See the header comment.
    void SetFrameBatching(int n)
    {
        Flush();
        batchFrames = n < 1 ? 1 : (n > 16 ? 16 : n);
    }
    int GetFrameBatching() const { return batchFrames; }
    // Send the queued frames to the device now.
    void Flush() const
    {
        if (queued == 0)
            return;
        const int n = queued;
        queued = 0;                                   // re-entrancy: the queue is empty while the call below runs
        std::vector<chs_frame> fr(n);
        for (int i = 0; i < n; i++)
        {
            const unsigned char *slot = ring + slotBytes * i;
            std::memset(&fr[i], 0, sizeof(chs_frame));
            fr[i].depth = reinterpret_cast<const float *>(slot);
            fr[i].color = qColorPath ? slot + colorOff : nullptr;
            fr[i].trunc_per_pixel = qInteg.trunc_kind == CHS_TRUNC_PER_PIXEL ? reinterpret_cast<const float *>(slot + truncOff) : nullptr;
            std::memcpy(fr[i].pose, &poses[24 * i], 12 * sizeof(float));
            std::memcpy(fr[i].color_pose, &poses[24 * i + 12], 12 * sizeof(float));
        }
        chs_integrator integ = qInteg;
        integ.trunc_per_pixel = nullptr;              // per frame, in chs_frame
        // the slots are page-locked and equally spaced: the whole queue crosses PCIe as one strided copy per image kind
        b200::Check(chs_integrate_batch(chunkManager.Handle(), &integ, n, fr.data(), CHS_MEM_HOST, &qCam, qChannels, qColorPath ? &qCcam : nullptr),
                    "chs_integrate_batch");
    }

    template <class DataType>
    void IntegrateDepthScan(const ProjectionIntegrator &integrator, const std::shared_ptr<const DepthImage<DataType>> &depthImage, const Transform &extrinsic,
                            const PinholeCamera &camera)
    {
        const chs_integrator integ = integrator.ToC(*depthImage, &truncScratch);
        const chs_camera cam = camera.ToC();
        float pose[12];
        b200::PoseToArray(extrinsic, pose);
        chunkManager.Touch();
        if (batchFrames > 1)
        {
            Enqueue(integ, cam, cam, false, 0, DepthAsFloat(*depthImage), nullptr, pose, pose);
            return;
        }
        b200::Check(chs_integrate_depth(chunkManager.Handle(), &integ, DepthAsFloat(*depthImage), CHS_MEM_HOST, pose, &cam), "chs_integrate_depth");
    }

    template <class DataType, class ColorType>
    void IntegrateDepthScanColor(const ProjectionIntegrator &integrator, const std::shared_ptr<const DepthImage<DataType>> &depthImage, const Transform &depthExtrinsic,
                                 const PinholeCamera &depthCamera, const std::shared_ptr<const ColorImage<ColorType>> &colorImage, const Transform &colorExtrinsic,
                                 const PinholeCamera &colorCamera)
    {
        static_assert(sizeof(ColorType) == 1, "8-bit colour images only (the reference instantiates <float, uint8_t>, CR ChiselServer.h:50-51)");
        const chs_integrator integ = integrator.ToC(*depthImage, &truncScratch);
        const chs_camera cam = depthCamera.ToC(), ccam = colorCamera.ToC();
        float pose[12], cpose[12];
        b200::PoseToArray(depthExtrinsic, pose);
        b200::PoseToArray(colorExtrinsic, cpose);
        chunkManager.Touch();
        if (batchFrames > 1)
        {
            Enqueue(integ, cam, ccam, true, static_cast<int>(colorImage->GetNumChannels()), DepthAsFloat(*depthImage),
                    reinterpret_cast<const uint8_t *>(colorImage->GetData()), pose, cpose);
            return;
        }
        b200::Check(chs_integrate_depth_color(chunkManager.Handle(), &integ, DepthAsFloat(*depthImage), CHS_MEM_HOST, pose, &cam,
                                              reinterpret_cast<const uint8_t *>(colorImage->GetData()), static_cast<int>(colorImage->GetNumChannels()), cpose, &ccam),
                    "chs_integrate_depth_color");
    }

    // The point-cloud fusion mode (Chisel.cpp:107-157) is outside the hot path this library accelerates.
    void IntegratePointCloud(const ProjectionIntegrator &, const PointCloud &, const Transform &, float, float)
    {
        std::fprintf(stderr, "chisel_b200: IntegratePointCloud is not implemented (fusion_mode=DepthImage is the supported mode)\n");
    }

    void GarbageCollect(const ChunkIDList &) {}     // untouched new chunks are never materialised on the device

    // Chisel.cpp:50-59: re-mesh the dirty set on calls 1, 11, 21, ...; dirty IDs accumulate in between.
    void UpdateMeshes()
    {
        if (updateCalls++ % 10 == 0)
        {
            chunkManager.RecomputeDirtyMeshes();
            meshesToUpdate.clear();
            dirtyVersion = -1;
        }
    }

    bool SaveAllMeshesToPLY(const std::string &filename) { return SaveAllMeshes(filename, false); }
    // extension: the same mesh as binary_little_endian PLY (io/PLY.h)
    bool SaveAllMeshesToPLYBinary(const std::string &filename) { return SaveAllMeshes(filename, true); }

  protected:
    bool SaveAllMeshes(const std::string &filename, bool binary)
    {
        // Chisel.cpp:69-105: concatenate every chunk mesh, indices 0..n-1
        MeshPtr full = std::make_shared<Mesh>();
        size_t v = 0;
        for (const std::pair<const ChunkID, MeshPtr> &it : chunkManager.GetAllMeshes())
        {
            for (const Vec3 &vert : it.second->vertices)
            {
                full->vertices.push_back(vert);
                full->indices.push_back(v++);
            }
            for (const Vec3 &c : it.second->colors)
                full->colors.push_back(c);
            for (const Vec3 &n : it.second->normals)
                full->normals.push_back(n);
        }
        return binary ? SaveMeshPLYBinary(filename, full) : SaveMeshPLYASCII(filename, full);
    }

  public:

    void Reset()
    {
        queued = 0;                                  // queued frames would be integrated and then dropped with the map
        chunkManager.Reset();
        meshesToUpdate.clear();
        dirtyVersion = -1;
    }

    // Host mirror of the device dirty set (Chisel.h:221-224), refreshed when a frame was integrated since the last call.
    const ChunkSet &GetMeshesToUpdate() const
    {
        Flush();
        int64_t n = 0;
        b200::Check(chs_num_dirty(chunkManager.Handle(), &n), "chs_num_dirty");
        if (dirtyVersion != n || n == 0)
        {
            meshesToUpdate.clear();
            std::vector<int32_t> ids(3 * n);
            if (n)
                b200::Check(chs_dirty_ids(chunkManager.Handle(), ids.data(), n), "chs_dirty_ids");
            for (int64_t i = 0; i < n; i++)
                meshesToUpdate[ChunkID(ids[3 * i], ids[3 * i + 1], ids[3 * i + 2])] = true;
            dirtyVersion = n;                        // the set only grows between re-meshes, so its size identifies it
        }
        return meshesToUpdate;
    }

  protected:
    template <class DataType>
    const float *DepthAsFloat(const DepthImage<DataType> &img)
    {
        return ConvertDepth(img.GetData(), static_cast<size_t>(img.GetWidth()) * img.GetHeight());
    }
    const float *ConvertDepth(const float *p, size_t) { return p; }
    template <class DataType>
    const float *ConvertDepth(const DataType *p, size_t n)
    {
        depthScratch.resize(n);
        for (size_t i = 0; i < n; i++)
            depthScratch[i] = static_cast<float>(p[i]);     // DepthImage::DepthAt (DepthImage.h:66-70)
        return depthScratch.data();
    }

    void InitBatching()
    {
        chunkManager.SetBeforeDeviceRead([this]() { Flush(); });
        if (const char *e = std::getenv("CHISEL_B200_BATCH"))
            SetFrameBatching(std::atoi(e));
    }

    static bool SameCamera(const chs_camera &a, const chs_camera &b) { return std::memcmp(&a, &b, sizeof(chs_camera)) == 0; }
    static bool SameIntegrator(const chs_integrator &a, const chs_integrator &b)
    {
        return a.trunc_kind == b.trunc_kind && a.trunc_param == b.trunc_param && a.weight == b.weight && a.carving_enabled == b.carving_enabled &&
               a.carving_dist == b.carving_dist;
    }
    // The caller reuses its image buffers (CR ChiselServer.cpp:285-295): the queue owns copies, in page-locked slots.
    void Enqueue(const chs_integrator &integ, const chs_camera &cam, const chs_camera &ccam, bool colorPath, int channels, const float *depth, const uint8_t *color,
                 const float *pose, const float *cpose)
    {
        if (queued > 0 && (!SameIntegrator(integ, qInteg) || !SameCamera(cam, qCam) || !SameCamera(ccam, qCcam) || colorPath != qColorPath || channels != qChannels))
            Flush();
        qInteg = integ;
        qCam = cam;
        qCcam = ccam;
        qColorPath = colorPath;
        qChannels = channels;
        const size_t npx = static_cast<size_t>(cam.width) * cam.height;
        const size_t depthBytes = (npx * sizeof(float) + 255) & ~static_cast<size_t>(255);
        const size_t colorBytes = colorPath ? ((static_cast<size_t>(ccam.width) * ccam.height * channels + 255) & ~static_cast<size_t>(255)) : 0;
        const size_t truncBytes = integ.trunc_kind == CHS_TRUNC_PER_PIXEL ? depthBytes : 0;
        const size_t need = depthBytes + colorBytes + truncBytes;
        if (need != slotBytes || !ring)
        {
            Flush();                                  // nothing is queued with another layout
            chs_host_free(ring);
            slotBytes = need;
            colorOff = depthBytes;
            truncOff = depthBytes + colorBytes;
            ring = static_cast<unsigned char *>(chs_host_alloc(slotBytes * 16));
            if (!ring)
                throw std::runtime_error("chisel_b200: cannot allocate the page-locked frame queue");
            poses.resize(24 * 16);
        }
        unsigned char *slot = ring + slotBytes * queued;
        std::memcpy(slot, depth, npx * sizeof(float));
        if (colorPath)
            std::memcpy(slot + colorOff, color, static_cast<size_t>(ccam.width) * ccam.height * channels);
        if (truncBytes)
            std::memcpy(slot + truncOff, integ.trunc_per_pixel, npx * sizeof(float));
        std::memcpy(&poses[24 * queued], pose, 12 * sizeof(float));
        std::memcpy(&poses[24 * queued + 12], cpose, 12 * sizeof(float));
        if (++queued >= batchFrames)
            Flush();
    }

    ChunkManager chunkManager;
    mutable ChunkSet meshesToUpdate;
    int updateCalls;
    mutable int64_t dirtyVersion;
    std::vector<float> truncScratch, depthScratch;
    int batchFrames;
    mutable int queued = 0;
    unsigned char *ring = nullptr;                   // 16 page-locked slots [depth | colour | per-pixel truncation]
    size_t slotBytes = 0, colorOff = 0, truncOff = 0;
    std::vector<float> poses;
    chs_integrator qInteg;
    chs_camera qCam, qCcam;
    bool qColorPath;
    int qChannels;
};
typedef std::shared_ptr<Chisel> ChiselPtr;
typedef std::shared_ptr<const Chisel> ChiselConstPtr;
} // namespace chisel
#endif
