// chisel_client.cpp -- a minimal caller of the open_chisel C++ API, written the way chisel_ros uses it
// (CR/src/ChiselServer.cpp:189-200, 266-295, 480-513, 534-567, 621-644). It uses ONLY the reference's public API, so the
// same file compiles against (a) the reference's own headers + sources and (b) the drop-in facade under
// cvids_b200/include + libchisel_b200.so. tests/test_facade.py builds both and compares their dumps.
//
//   chisel_client <stream.bin> <dump.bin> [time]
// With "time" the frames are read into memory first, the per-frame loop is exactly ChiselServer::IntegrateLastDepthImage
// (integrate, PublishLatestChunkBoxes, PublishDepthFrustum, then UpdateMeshes -- the every-10th gate decides) and its wall time
// is printed (tools/dropin_demo.py, tests/test_facade.py::test_dropin_full_size).
#include <open_chisel/Chisel.h>
#include <open_chisel/truncation/ConstantTruncator.h>
#include <open_chisel/truncation/InverseTruncator.h>
#include <open_chisel/truncation/QuadraticTruncator.h>
#include <open_chisel/weighting/ConstantWeighter.h>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

struct Header
{
    int32_t width, height, channels, frames, chunk, useColor, truncKind, carve, updateEvery;
    float resolution, truncParam, weight, carveDist, nearPlane, farPlane, fx, fy, cx, cy;
};

static bool IdLess(const chisel::ChunkID &a, const chisel::ChunkID &b)
{
    if (a(0) != b(0)) return a(0) < b(0);
    if (a(1) != b(1)) return a(1) < b(1);
    return a(2) < b(2);
}

template <class T> static void Put(FILE *f, const T &v) { fwrite(&v, sizeof(T), 1, f); }
static void PutVecs(FILE *f, const chisel::Vec3List &l)
{
    Put<int64_t>(f, (int64_t)l.size());
    for (size_t i = 0; i < l.size(); i++)
        for (int k = 0; k < 3; k++)
            Put<float>(f, l[i](k));
}

int main(int argc, char **argv)
{
    if (argc < 3)
        return 2;
    FILE *in = fopen(argv[1], "rb");
    if (!in)
        return 3;
    Header h;
    if (fread(&h, sizeof(h), 1, in) != 1)
        return 4;

    chisel::ChiselPtr chiselMap(new chisel::Chisel(Eigen::Vector3i(h.chunk, h.chunk, h.chunk), h.resolution, h.useColor != 0));

    // ChiselServer::SetupProjectionIntegrator (CR/src/ChiselServer.cpp:480-487)
    chisel::ProjectionIntegrator projectionIntegrator;
    chisel::TruncatorPtr truncator;
    if (h.truncKind == 0) truncator.reset(new chisel::ConstantTruncator(h.truncParam));
    else if (h.truncKind == 1) truncator.reset(new chisel::QuadraticTruncator(h.truncParam));
    else truncator.reset(new chisel::InverseTruncator(h.truncParam));
    projectionIntegrator.SetCentroids(chiselMap->GetChunkManager().GetCentroids());
    projectionIntegrator.SetTruncator(truncator);
    projectionIntegrator.SetWeighter(chisel::WeighterPtr(new chisel::ConstantWeighter(h.weight)));
    projectionIntegrator.SetCarvingDist(h.carveDist);
    projectionIntegrator.SetCarvingEnabled(h.carve != 0);

    chisel::PinholeCamera cameraModel;
    chisel::Intrinsics intrinsics;
    intrinsics.SetFx(h.fx);
    intrinsics.SetFy(h.fy);
    intrinsics.SetCx(h.cx);
    intrinsics.SetCy(h.cy);
    cameraModel.SetIntrinsics(intrinsics);
    cameraModel.SetWidth(h.width);
    cameraModel.SetHeight(h.height);
    cameraModel.SetNearPlane(h.nearPlane);
    cameraModel.SetFarPlane(h.farPlane);

    // one depth and one colour buffer reused across frames, like ChiselServer::SetDepthImage / SetColorImage
    std::shared_ptr<chisel::DepthImage<float>> lastDepthImage(new chisel::DepthImage<float>(h.width, h.height));
    std::shared_ptr<chisel::ColorImage<uint8_t>> lastColorImage(new chisel::ColorImage<uint8_t>(h.width, h.height, h.channels > 0 ? h.channels : 1));
    const size_t npx = (size_t)h.width * h.height;
    chisel::Frustum frustum;
    int remeshes = 0;
    const bool timed = argc > 3 && std::strcmp(argv[3], "time") == 0;
    std::vector<float> allPoses, allDepth;
    std::vector<uint8_t> allColor;
    if (timed)
    {
        allPoses.resize((size_t)h.frames * 12);
        allDepth.resize((size_t)h.frames * npx);
        allColor.resize((size_t)h.frames * npx * (h.channels > 0 ? h.channels : 0));
        for (int f = 0; f < h.frames; f++)
        {
            if (fread(&allPoses[(size_t)f * 12], sizeof(float), 12, in) != 12) return 5;
            if (fread(&allDepth[(size_t)f * npx], sizeof(float), npx, in) != npx) return 5;
            if (h.channels > 0 && fread(&allColor[(size_t)f * npx * h.channels], 1, npx * h.channels, in) != npx * h.channels) return 5;
        }
    }
    double boxSum = 0.0;
    long boxes = 0;
    const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    for (int f = 0; f < h.frames; f++)
    {
        float pose[12];
        if (timed)
        {
            // what the ROS callbacks do: copy the message into the one reused image buffer (CR ChiselServer.cpp:266-295)
            std::memcpy(pose, &allPoses[(size_t)f * 12], sizeof(pose));
            std::memcpy(lastDepthImage->GetMutableData(), &allDepth[(size_t)f * npx], npx * sizeof(float));
            if (h.channels > 0)
                std::memcpy(lastColorImage->GetMutableData(), &allColor[(size_t)f * npx * h.channels], npx * h.channels);
        }
        else
        {
            if (fread(pose, sizeof(float), 12, in) != 12) return 5;
            if (fread(lastDepthImage->GetMutableData(), sizeof(float), npx, in) != npx) return 5;
            if (h.channels > 0 && fread(lastColorImage->GetMutableData(), 1, npx * h.channels, in) != npx * h.channels) return 5;
        }
        chisel::Transform lastPose;
        for (int r = 0; r < 3; r++)
        {
            for (int c = 0; c < 3; c++)
                lastPose.linear()(r, c) = pose[r * 4 + c];
            lastPose.translation()(r) = pose[r * 4 + 3];
        }
        std::shared_ptr<const chisel::DepthImage<float>> depth = lastDepthImage;
        std::shared_ptr<const chisel::ColorImage<uint8_t>> color = lastColorImage;
        if (h.channels > 0)
            chiselMap->IntegrateDepthScanColor<float, uint8_t>(projectionIntegrator, depth, lastPose, cameraModel, color, lastPose, cameraModel);
        else
            chiselMap->IntegrateDepthScan<float>(projectionIntegrator, depth, lastPose, cameraModel);
        // ChiselServer::PublishLatestChunkBoxes (CR/src/ChiselServer.cpp:533-567), run after EVERY frame: the centre of every dirty
        // chunk that exists goes into a marker message
        {
            const chisel::ChunkManager &chunkManager = chiselMap->GetChunkManager();
            const chisel::ChunkSet &latest = chiselMap->GetMeshesToUpdate();
            for (const std::pair<const chisel::ChunkID, bool> &id : latest)
                if (chunkManager.HasChunk(id.first))
                {
                    const chisel::Vec3 center = chunkManager.GetChunk(id.first)->ComputeBoundingBox().GetCenter();
                    boxSum += center.x() + center.y() + center.z();
                    boxes++;
                }
        }
        // ChiselServer::PublishDepthFrustum
        cameraModel.SetupFrustum(lastPose, &frustum);
        // the reference gates re-meshing on a process-global call counter (Chisel.cpp:50-59): call it every frame like
        // ChiselServer::IntegrateLastDepthImage does and let the gate decide
        if (timed)
        {
            chiselMap->UpdateMeshes();
            continue;
        }
        const size_t before = chiselMap->GetMeshesToUpdate().size();
        chiselMap->UpdateMeshes();
        if (before > 0 && chiselMap->GetMeshesToUpdate().size() == 0)
            remeshes++;
    }
    if (timed)
    {
        (void)chiselMap->GetMeshesToUpdate().size();             // everything the loop queued has reached the map
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        fprintf(stderr, "TIMING frames %d seconds %.6f fps %.3f boxes %ld (%.3f)\n", h.frames, sec, h.frames / sec, boxes, boxSum);
    }
    fclose(in);

    FILE *out = fopen(argv[2], "wb");
    if (!out)
        return 6;
    const chisel::ChunkManager &cm = chiselMap->GetChunkManager();
    // chunks, sorted by ID (ChiselServer::GetAllChunks / Serialization.h walk the same accessors)
    std::vector<chisel::ChunkID, Eigen::aligned_allocator<chisel::ChunkID>> ids;
    for (const std::pair<const chisel::ChunkID, chisel::ChunkPtr> &it : cm.GetChunks())
        ids.push_back(it.first);
    std::sort(ids.begin(), ids.end(), IdLess);
    Put<int64_t>(out, (int64_t)ids.size());
    Put<int64_t>(out, (int64_t)remeshes);
    for (size_t i = 0; i < ids.size(); i++)
    {
        if (!cm.HasChunk(ids[i]))
            return 7;
        chisel::ChunkPtr c = cm.GetChunk(ids[i]);
        const chisel::Vec3 center = c->ComputeBoundingBox().GetCenter();
        for (int k = 0; k < 3; k++) Put<int32_t>(out, c->GetID()(k));
        for (int k = 0; k < 3; k++) Put<float>(out, center(k));
        Put<int32_t>(out, (int32_t)c->GetTotalNumVoxels());
        Put<int32_t>(out, c->HasColors() ? 1 : 0);
        for (const chisel::DistVoxel &v : c->GetVoxels())
        {
            Put<float>(out, v.GetSDF());
            Put<float>(out, v.GetWeight());
        }
        if (c->HasColors())
            for (const chisel::ColorVoxel &v : c->GetColorVoxels())
            {
                Put<uint8_t>(out, v.GetRed());
                Put<uint8_t>(out, v.GetGreen());
                Put<uint8_t>(out, v.GetBlue());
                Put<uint8_t>(out, v.GetWeight());
            }
    }
    // dirty set
    ids.clear();
    for (const std::pair<const chisel::ChunkID, bool> &it : chiselMap->GetMeshesToUpdate())
        ids.push_back(it.first);
    std::sort(ids.begin(), ids.end(), IdLess);
    Put<int64_t>(out, (int64_t)ids.size());
    for (size_t i = 0; i < ids.size(); i++)
        for (int k = 0; k < 3; k++) Put<int32_t>(out, ids[i](k));
    // meshes (ChiselServer::FillMarkerTopicWithMeshes reads grids, vertices, colors / normals)
    ids.clear();
    for (const std::pair<const chisel::ChunkID, chisel::MeshPtr> &it : cm.GetAllMeshes())
        ids.push_back(it.first);
    std::sort(ids.begin(), ids.end(), IdLess);
    Put<int64_t>(out, (int64_t)ids.size());
    for (size_t i = 0; i < ids.size(); i++)
    {
        const chisel::MeshPtr &m = cm.GetAllMeshes().at(ids[i]);
        for (int k = 0; k < 3; k++) Put<int32_t>(out, ids[i](k));
        PutVecs(out, m->vertices);
        PutVecs(out, m->normals);
        PutVecs(out, m->colors);
        PutVecs(out, m->grids);
        Put<int32_t>(out, m->HasColors() ? 1 : 0);
        Put<int32_t>(out, m->HasNormals() ? 1 : 0);
    }
    // last frustum (ChiselServer::PublishDepthFrustum reads GetLines)
    for (int i = 0; i < 24; i++)
        for (int k = 0; k < 3; k++) Put<float>(out, frustum.GetLines()[i](k));
    fclose(out);
    if (!chiselMap->SaveAllMeshesToPLY(std::string(argv[2]) + ".ply"))
        return 8;
    chiselMap->Reset();
    if (chiselMap->GetChunkManager().GetChunks().size() != 0 || chiselMap->GetMeshesToUpdate().size() != 0)
        return 9;
    return 0;
}
