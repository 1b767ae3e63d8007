// open_chisel/Chisel.h -- drop-in facade of chisel::Chisel (OC/include/open_chisel/Chisel.h:37-229, OC/src/Chisel.cpp) over
// libchisel_b200. Source compatible with the calls chisel_ros makes (CR/src/ChiselServer.cpp:189-200, 495-525, 346, 554,
// 721, 729-737): swap the include path, link -lchisel_b200, nothing else changes.
//
// Semantics kept: UpdateMeshes re-meshes on every 10th call only (Chisel.cpp:50-59; counter per instance instead of
// process-global, quirk Q4); every new-and-untouched chunk is garbage collected (the deterministic reading of quirk Q2).
// Not kept: the reference's printf chatter and its per-frame PrintMemoryStatistics scan (Chisel.h:62,109-111).
//
// Frame batching (extension, off by default; SetFrameBatching(n) or the environment variable CHISEL_B200_BATCH=n, n <= 16):
// IntegrateDepthScan[Color] uploads the frame into a device-side queue (one PCIe copy straight from the caller's image -- the
// facade's DepthImage / ColorImage live in page-locked memory -- and the caller may reuse its buffers at once, as chisel_ros does,
// CR ChiselServer.cpp:285-295) and returns; the queue is integrated as ONE chs_integrate_batch call (fused multi-frame kernels,
// bit-identical to frame-by-frame integration) when it holds n frames, when the integrator / camera settings change, or when
// something needs the device map up to date: UpdateMeshes on a call that actually re-meshes (every 10th), GetChunks, Reset,
// SaveAllMeshesToPLY.
// Reads between flushes. chisel_ros looks at the map after EVERY frame (PublishLatestChunkBoxes: GetMeshesToUpdate, then HasChunk /
// GetChunk()->ComputeBoundingBox() per dirty ID, CR ChiselServer.cpp:533-567). By default those reads flush the queue first --
// exactly the reference's view at every instant, but then every frame travels alone. With SetRelaxedReads(true) (or
// CHISEL_B200_RELAXED_READS=1) they are served from host mirrors of the map AS OF THE LAST FLUSH: the marker boxes lag by at most
// n - 1 frames, the map, its meshes and everything read after a flush are unchanged, and a live stream runs on the fused path.
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
        FreeRings();
    }

    const ChunkManager &GetChunkManager() const { return chunkManager; }
    ChunkManager &GetMutableChunkManager() { return chunkManager; }
    void SetChunkManager(const ChunkManager &manager)
    {
        Flush();
        chunkManager = manager;
        InstallReadHook();
    }

    // Queue up to n frames (1 = off: every frame goes to the device at once). See the header comment.
    void SetFrameBatching(int n)
    {
        Flush();
        batchFrames = n < 1 ? 1 : (n > 16 ? 16 : n);
    }
    int GetFrameBatching() const { return batchFrames; }
    // Reads between flushes are served from the host mirrors of the last flushed state (see the header comment).
    void SetRelaxedReads(bool on)
    {
        relaxedReads = on;
        InstallReadHook();
    }
    bool GetRelaxedReads() const { return relaxedReads; }
    // Integrate the queued frames now.
    void Flush() const
    {
        if (queued == 0)
            return;
        const int n = queued;
        queued = 0;                                   // re-entrancy: the queue is empty while the call below runs
        std::vector<chs_frame> fr(n);
        const unsigned char *base = ringDev[curRing];
        for (int i = 0; i < n; i++)
        {
            const unsigned char *slot = base + slotBytes * i;
            std::memset(&fr[i], 0, sizeof(chs_frame));
            fr[i].depth = reinterpret_cast<const float *>(slot);
            fr[i].color = qColorPath ? slot + colorOff : nullptr;
            fr[i].trunc_per_pixel = qInteg.trunc_kind == CHS_TRUNC_PER_PIXEL ? reinterpret_cast<const float *>(slot + truncOff) : nullptr;
            std::memcpy(fr[i].pose, &poses[24 * i], 12 * sizeof(float));
            std::memcpy(fr[i].color_pose, &poses[24 * i + 12], 12 * sizeof(float));
        }
        chs_integrator integ = qInteg;
        integ.trunc_per_pixel = nullptr;              // per frame, in chs_frame
        // the frames are complete in device memory (chs_upload returned): Hi-Z and colour packing of this batch run on the
        // library's copy stream beside the kernels of the previous one
        b200::Check(chs_integrate_batch(chunkManager.Handle(), &integ, n, fr.data(), CHS_MEM_DEVICE_ASYNC, &qCam, qChannels, qColorPath ? &qCcam : nullptr),
                    "chs_integrate_batch");
        b200::Check(chs_last_batch_ticket(chunkManager.Handle(), &ringTicket[curRing]), "chs_last_batch_ticket");
        curRing ^= 1;                                 // the other ring is refilled while this batch is in flight
        const_cast<ChunkManager &>(chunkManager).Touch();
    }

    template <class DataType>
    void IntegrateDepthScan(const ProjectionIntegrator &integrator, const std::shared_ptr<const DepthImage<DataType>> &depthImage, const Transform &extrinsic,
                            const PinholeCamera &camera)
    {
        const chs_integrator integ = integrator.ToC(*depthImage, &truncScratch);
        const chs_camera cam = camera.ToC();
        float pose[12];
        b200::PoseToArray(extrinsic, pose);
        if (batchFrames > 1)
        {
            Enqueue(integ, cam, cam, false, 0, DepthAsFloat(*depthImage), nullptr, pose, pose);
            return;
        }
        chunkManager.Touch();
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
        if (batchFrames > 1)
        {
            Enqueue(integ, cam, ccam, true, static_cast<int>(colorImage->GetNumChannels()), DepthAsFloat(*depthImage),
                    reinterpret_cast<const uint8_t *>(colorImage->GetData()), pose, cpose);
            return;
        }
        chunkManager.Touch();
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
            Flush();
            chunkManager.RecomputeDirtyMeshes();
            meshesToUpdate.clear();
            dirtyVersion = -1;
            dirtyMapVersion = -1;
        }
    }

    bool SaveAllMeshesToPLY(const std::string &filename) { return SaveAllMeshes(filename, false); }
    // extension: the same mesh as binary_little_endian PLY (io/PLY.h)
    bool SaveAllMeshesToPLYBinary(const std::string &filename) { return SaveAllMeshes(filename, true); }

  protected:
    bool SaveAllMeshes(const std::string &filename, bool binary)
    {
        Flush();
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
        dirtyMapVersion = -1;
    }

    // Host mirror of the device dirty set (Chisel.h:221-224), refreshed when a frame was integrated since the last call.
    const ChunkSet &GetMeshesToUpdate() const
    {
        if (!relaxedReads)
            Flush();
        if (dirtyMapVersion == chunkManager.Version())
            return meshesToUpdate;                   // nothing reached the device since the mirror was read
        dirtyMapVersion = chunkManager.Version();
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

    void InstallReadHook()
    {
        // what ChunkManager runs before HasChunk / GetChunk / RecomputeDirtyMeshes; GetChunks (whole-map export) always flushes
        if (relaxedReads)
            chunkManager.SetBeforeDeviceRead(std::function<void()>());
        else
            chunkManager.SetBeforeDeviceRead([this]() { Flush(); });
        chunkManager.SetBeforeBulkRead([this]() { Flush(); });
    }
    void InitBatching()
    {
        if (const char *e = std::getenv("CHISEL_B200_RELAXED_READS"))
            relaxedReads = std::atoi(e) != 0;
        InstallReadHook();
        if (const char *e = std::getenv("CHISEL_B200_BATCH"))
            SetFrameBatching(std::atoi(e));
    }
    void FreeRings()
    {
        for (int r = 0; r < 2; r++)
        {
            if (ringDev[r])
                chs_device_free(chunkManager.Handle(), ringDev[r]);
            ringDev[r] = nullptr;
            ringTicket[r] = 0;
        }
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
        if (need != slotBytes || !ringDev[0])
        {
            Flush();                                  // nothing is queued with another layout
            FreeRings();
            slotBytes = need;
            colorOff = depthBytes;
            truncOff = depthBytes + colorBytes;
            for (int r = 0; r < 2; r++)
            {
                ringDev[r] = static_cast<unsigned char *>(chs_device_alloc(chunkManager.Handle(), slotBytes * 16));
                if (!ringDev[r])
                    throw std::runtime_error("chisel_b200: cannot allocate the device frame queue");
            }
            curRing = 0;
            poses.resize(24 * 16);
        }
        if (queued == 0 && ringTicket[curRing])
        {
            // the batch that read this ring last (two flushes ago) must be done before its slots are overwritten
            int n = 0;
            b200::Check(chs_wait_batch(chunkManager.Handle(), ringTicket[curRing], nullptr, 0, &n), "chs_wait_batch");
            ringTicket[curRing] = 0;
        }
        unsigned char *slot = ringDev[curRing] + slotBytes * queued;
        b200::Check(chs_upload(chunkManager.Handle(), slot, depth, npx * sizeof(float)), "chs_upload");
        if (colorPath)
            b200::Check(chs_upload(chunkManager.Handle(), slot + colorOff, color, static_cast<size_t>(ccam.width) * ccam.height * channels), "chs_upload");
        if (truncBytes)
            b200::Check(chs_upload(chunkManager.Handle(), slot + truncOff, integ.trunc_per_pixel, npx * sizeof(float)), "chs_upload");
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
    unsigned char *ringDev[2] = {nullptr, nullptr};  // two device rings of 16 slots [depth | colour | per-pixel truncation], used alternately
    mutable int64_t ringTicket[2] = {0, 0};          // the batch that read the ring last
    mutable int curRing = 0;
    bool relaxedReads = false;
    mutable long dirtyMapVersion = -1;
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
