// oracle/_ref driver: a C-ABI shim over the REFERENCE's own classes, compiled together with the
// unmodified reference sources from /root/reference/OpenChisel/open_chisel (see oracle/Makefile).
// TEST INFRASTRUCTURE, not a product component: only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load the library this builds.
//
// What is the reference's and what is restated here:
//  * depth path: chisel::Chisel::IntegrateDepthScan<float> is called as-is (Chisel.h:59-112; serial).
//  * colour path, mode 0 (oracle): a SERIAL restatement of Chisel.h:115-213 that calls the reference's
//    own public ChunkManager / ProjectionIntegrator::IntegrateColor; needed because the as-is function
//    writes a std::vector<bool> from 16 threads (Chisel.h:130-132,170-173) and is nondeterministic
//    (SURVEY.md Q2). Mode 1 calls the as-is threaded function (CPU-baseline timing only).
//  * meshing: ChunkManager::RecomputeMesh (ChunkManager.cpp:91-128) is called serially over the dirty
//    set (mode 0; Q3), or ChunkManager::RecomputeMeshes as-is (mode 1, timing only). The every-10th
//    gate of Chisel::UpdateMeshes (Chisel.cpp:50-59, process-global static) is NOT applied here.
#include <open_chisel/Chisel.h>
#include <open_chisel/truncation/ConstantTruncator.h>
#include <open_chisel/truncation/QuadraticTruncator.h>
#include <open_chisel/truncation/InverseTruncator.h>
#include <open_chisel/weighting/ConstantWeighter.h>
#include <algorithm>
#include <cstring>
#include <cstdint>

namespace
{
using namespace chisel;

struct OpenChisel : public Chisel
{
    OpenChisel(const Eigen::Vector3i &cs, float res, bool color) : Chisel(cs, res, color) {}
    ChunkSet &Dirty() { return meshesToUpdate; }
};

struct Ref
{
    OpenChisel *map;
    ProjectionIntegrator integrator;
    std::shared_ptr<DepthImage<float>> depth;
    std::shared_ptr<ColorImage<uint8_t>> color;
    long lastCandidates;
    long lastNew;
    long lastGarbage;
};

Transform MakePose(const float *p) // row-major 3x4 [R|t], camera -> world
{
    Transform t;
    for (int r = 0; r < 3; r++)
    {
        for (int c = 0; c < 3; c++)
            t.linear()(r, c) = p[r * 4 + c];
        t.translation()(r) = p[r * 4 + 3];
    }
    return t;
}

PinholeCamera MakeCamera(const float *c) // fx fy cx cy W H near far
{
    PinholeCamera cam;
    Intrinsics k;
    k.SetFx(c[0]);
    k.SetFy(c[1]);
    k.SetCx(c[2]);
    k.SetCy(c[3]);
    cam.SetIntrinsics(k);
    cam.SetWidth((int)c[4]);
    cam.SetHeight((int)c[5]);
    cam.SetNearPlane(c[6]);
    cam.SetFarPlane(c[7]);
    return cam;
}

void SetDepth(Ref *r, const float *d, int W, int H)
{
    if (!r->depth || r->depth->GetWidth() != W || r->depth->GetHeight() != H)
        r->depth.reset(new DepthImage<float>(W, H));
    std::memcpy(r->depth->GetMutableData(), d, sizeof(float) * (size_t)W * H);
}

void SetColor(Ref *r, const uint8_t *c, int W, int H, int ch)
{
    if (!r->color || r->color->GetWidth() != W || r->color->GetHeight() != H || (int)r->color->GetNumChannels() != ch)
        r->color.reset(new ColorImage<uint8_t>(W, H, ch));
    std::memcpy(r->color->GetMutableData(), c, (size_t)W * H * ch);
}

bool IdLess(const ChunkID &a, const ChunkID &b)
{
    if (a(0) != b(0)) return a(0) < b(0);
    if (a(1) != b(1)) return a(1) < b(1);
    return a(2) < b(2);
}
} // namespace

extern "C"
{

void *ref_create(int cs, float res, int useColor)
{
    Ref *r = new Ref();
    r->map = new OpenChisel(Eigen::Vector3i(cs, cs, cs), res, useColor != 0);
    r->integrator.SetCentroids(r->map->GetChunkManager().GetCentroids());
    r->lastCandidates = r->lastNew = r->lastGarbage = 0;
    return r;
}

void ref_destroy(void *h)
{
    Ref *r = (Ref *)h;
    delete r->map;
    delete r;
}

void ref_reset(void *h) { ((Ref *)h)->map->Reset(); }

// truncKind: 0 constant(param = metres), 1 quadratic(param = scale), 2 inverse(param = scale)
void ref_setup_integrator(void *h, int truncKind, float truncParam, float weight, int carve, float carveDist)
{
    Ref *r = (Ref *)h;
    TruncatorPtr t;
    if (truncKind == 0) t.reset(new ConstantTruncator(truncParam));
    else if (truncKind == 1) t.reset(new QuadraticTruncator(truncParam));
    else t.reset(new InverseTruncator(truncParam));
    r->integrator.SetTruncator(t);
    r->integrator.SetWeighter(WeighterPtr(new ConstantWeighter(weight)));
    r->integrator.SetCarvingDist(carveDist);
    r->integrator.SetCarvingEnabled(carve != 0);
}

float ref_truncation(int truncKind, float truncParam, float depth)
{
    if (truncKind == 0) return ConstantTruncator(truncParam).GetTruncationDistance(depth);
    if (truncKind == 1) return QuadraticTruncator(truncParam).GetTruncationDistance(depth);
    return InverseTruncator(truncParam).GetTruncationDistance(depth);
}

void ref_integrate_depth(void *h, const float *depth, int W, int H, const float *pose, const float *cam)
{
    Ref *r = (Ref *)h;
    SetDepth(r, depth, W, H);
    std::shared_ptr<const DepthImage<float>> d = r->depth;
    r->map->IntegrateDepthScan<float>(r->integrator, d, MakePose(pose), MakeCamera(cam));
}

void ref_integrate_color(void *h, const float *depth, int W, int H, const float *pose, const float *cam,
                         const uint8_t *color, int cW, int cH, int channels, const float *cpose, const float *ccam,
                         int asIs)
{
    Ref *r = (Ref *)h;
    SetDepth(r, depth, W, H);
    SetColor(r, color, cW, cH, channels);
    std::shared_ptr<const DepthImage<float>> d = r->depth;
    std::shared_ptr<const ColorImage<uint8_t>> c = r->color;
    const Transform depthPose = MakePose(pose), colorPose = MakePose(cpose);
    const PinholeCamera depthCam = MakeCamera(cam), colorCam = MakeCamera(ccam);
    if (asIs)
    {
        r->map->IntegrateDepthScanColor<float, uint8_t>(r->integrator, d, depthPose, depthCam, c, colorPose, colorCam);
        return;
    }
    // Serial restatement of Chisel.h:115-213 over the reference's own components.
    ChunkManager &cm = r->map->GetMutableChunkManager();
    Frustum frustum;
    depthCam.SetupFrustum(depthPose, &frustum);                 // Chisel.h:119-120
    ChunkIDList ids;
    cm.GetChunkIDsIntersecting(frustum, &ids);                  // :122-123
    const int n = (int)ids.size();
    std::vector<char> isNew(n, 0), isGarbage(n, 0);
    for (int i = 0; i < n; i++)                                 // :133-143
        if (!cm.HasChunk(ids[i]))
        {
            isNew[i] = 1;
            cm.CreateChunk(ids[i]);
        }
    long nNew = 0, nGarbage = 0;
    for (int k = 0; k < n; k++)                                 // :160-190, one thread
    {
        ChunkPtr chunk = cm.GetChunk(ids[k]);
        bool needsUpdate = r->integrator.IntegrateColor(d, depthCam, depthPose, c, colorCam, colorPose, chunk.get());
        if (!needsUpdate && isNew[k])
            isGarbage[k] = 1;
        if (needsUpdate)
            for (int dx = -1; dx <= 1; dx++)
                for (int dy = -1; dy <= 1; dy++)
                    for (int dz = -1; dz <= 1; dz++)
                        r->map->Dirty()[ids[k] + ChunkID(dx, dy, dz)] = true;
    }
    for (int i = 0; i < n; i++)                                 // :202-207 (every garbage flag honoured: Q2)
    {
        nNew += isNew[i];
        if (isGarbage[i])
        {
            cm.RemoveChunk(ids[i]);
            nGarbage++;
        }
    }
    r->lastCandidates = n;
    r->lastNew = nNew;
    r->lastGarbage = nGarbage;
}

// Candidate IDs for a frame exactly as the reference enumerates them (ChunkManager.cpp:182-212).
int ref_candidate_ids(void *h, const float *pose, const float *cam, int *out, int cap)
{
    Ref *r = (Ref *)h;
    Frustum frustum;
    MakeCamera(cam).SetupFrustum(MakePose(pose), &frustum);
    ChunkIDList ids;
    r->map->GetMutableChunkManager().GetChunkIDsIntersecting(frustum, &ids);
    const int n = (int)ids.size();
    for (int i = 0; i < n && i < cap; i++)
        for (int k = 0; k < 3; k++)
            out[3 * i + k] = ids[i](k);
    return n;
}

// corners[8*3], lines[24*3], planes[6*4] in the order far, near, top, bottom, left, right (Frustum.cpp:43).
void ref_frustum(const float *pose, const float *cam, float *corners, float *lines, float *planes)
{
    Frustum f;
    MakeCamera(cam).SetupFrustum(MakePose(pose), &f);
    for (int i = 0; i < 8; i++)
        for (int k = 0; k < 3; k++)
            corners[3 * i + k] = f.GetCorners()[i](k);
    for (int i = 0; i < 24; i++)
        for (int k = 0; k < 3; k++)
            lines[3 * i + k] = f.GetLines()[i](k);
    const Plane *p[6] = {&f.GetFarPlane(), &f.GetNearPlane(), &f.GetTopPlane(), &f.GetBottomPlane(), &f.GetLeftPlane(), &f.GetRightPlane()};
    for (int i = 0; i < 6; i++)
    {
        for (int k = 0; k < 3; k++)
            planes[4 * i + k] = p[i]->normal(k);
        planes[4 * i + 3] = p[i]->distance;
    }
}

void ref_update_meshes(void *h, int asIs)
{
    Ref *r = (Ref *)h;
    ChunkManager &cm = r->map->GetMutableChunkManager();
    if (asIs)
        cm.RecomputeMeshes(r->map->Dirty());                    // ChunkManager.cpp:130-169, 16 threads
    else
    {
        std::mutex m;
        for (const std::pair<const ChunkID, bool> &it : r->map->Dirty())
            if (it.second)
                cm.RecomputeMesh(it.first, m);                  // ChunkManager.cpp:91-128
    }
    r->map->Dirty().clear();                                    // Chisel.cpp:57
}

int ref_num_chunks(void *h) { return (int)((Ref *)h)->map->GetChunkManager().GetChunks().size(); }

void ref_chunk_ids(void *h, int *out) // sorted lexicographically
{
    Ref *r = (Ref *)h;
    std::vector<ChunkID, Eigen::aligned_allocator<ChunkID>> ids;
    for (const auto &it : r->map->GetChunkManager().GetChunks())
        ids.push_back(it.first);
    std::sort(ids.begin(), ids.end(), IdLess);
    for (size_t i = 0; i < ids.size(); i++)
        for (int k = 0; k < 3; k++)
            out[3 * i + k] = ids[i](k);
}

// sdf[V], weight[V], rgba[4V] (r,g,b,colour weight). Returns 0 if the chunk does not exist.
int ref_chunk_voxels(void *h, const int *id, float *sdf, float *weight, uint8_t *rgbw)
{
    Ref *r = (Ref *)h;
    const ChunkManager &cm = r->map->GetChunkManager();
    const ChunkID cid(id[0], id[1], id[2]);
    if (!cm.HasChunk(cid))
        return 0;
    ChunkPtr c = cm.GetChunk(cid);
    const size_t V = c->GetTotalNumVoxels();
    for (size_t i = 0; i < V; i++)
    {
        sdf[i] = c->GetVoxels()[i].GetSDF();
        weight[i] = c->GetVoxels()[i].GetWeight();
    }
    if (rgbw && c->HasColors())
        for (size_t i = 0; i < V; i++)
        {
            const ColorVoxel &v = c->GetColorVoxels()[i];
            rgbw[4 * i + 0] = v.GetRed();
            rgbw[4 * i + 1] = v.GetGreen();
            rgbw[4 * i + 2] = v.GetBlue();
            rgbw[4 * i + 3] = v.GetWeight();
        }
    return 1;
}

int ref_num_dirty(void *h) { return (int)((Ref *)h)->map->GetMeshesToUpdate().size(); }

void ref_dirty_ids(void *h, int *out)
{
    Ref *r = (Ref *)h;
    std::vector<ChunkID, Eigen::aligned_allocator<ChunkID>> ids;
    for (const auto &it : r->map->GetMeshesToUpdate())
        ids.push_back(it.first);
    std::sort(ids.begin(), ids.end(), IdLess);
    for (size_t i = 0; i < ids.size(); i++)
        for (int k = 0; k < 3; k++)
            out[3 * i + k] = ids[i](k);
}

int ref_num_meshes(void *h) { return (int)((Ref *)h)->map->GetChunkManager().GetAllMeshes().size(); }

void ref_mesh_ids(void *h, int *out)
{
    Ref *r = (Ref *)h;
    std::vector<ChunkID, Eigen::aligned_allocator<ChunkID>> ids;
    for (const auto &it : r->map->GetChunkManager().GetAllMeshes())
        ids.push_back(it.first);
    std::sort(ids.begin(), ids.end(), IdLess);
    for (size_t i = 0; i < ids.size(); i++)
        for (int k = 0; k < 3; k++)
            out[3 * i + k] = ids[i](k);
}

// sizes[0] = vertices, [1] = normals, [2] = colors, [3] = grids, [4] = indices; returns 0 if no mesh.
int ref_mesh_sizes(void *h, const int *id, long *sizes)
{
    Ref *r = (Ref *)h;
    const ChunkManager &cm = r->map->GetChunkManager();
    const ChunkID cid(id[0], id[1], id[2]);
    if (!cm.HasMesh(cid))
        return 0;
    const MeshPtr &m = cm.GetMesh(cid);
    sizes[0] = (long)m->vertices.size();
    sizes[1] = (long)m->normals.size();
    sizes[2] = (long)m->colors.size();
    sizes[3] = (long)m->grids.size();
    sizes[4] = (long)m->indices.size();
    return 1;
}

int ref_mesh_data(void *h, const int *id, float *verts, float *normals, float *colors, float *grids, long *indices)
{
    Ref *r = (Ref *)h;
    const ChunkManager &cm = r->map->GetChunkManager();
    const ChunkID cid(id[0], id[1], id[2]);
    if (!cm.HasMesh(cid))
        return 0;
    const MeshPtr &m = cm.GetMesh(cid);
    for (size_t i = 0; i < m->vertices.size(); i++)
        for (int k = 0; k < 3; k++)
            verts[3 * i + k] = m->vertices[i](k);
    for (size_t i = 0; i < m->normals.size(); i++)
        for (int k = 0; k < 3; k++)
            normals[3 * i + k] = m->normals[i](k);
    if (colors)
        for (size_t i = 0; i < m->colors.size(); i++)
            for (int k = 0; k < 3; k++)
                colors[3 * i + k] = m->colors[i](k);
    for (size_t i = 0; i < m->grids.size(); i++)
        for (int k = 0; k < 3; k++)
            grids[3 * i + k] = m->grids[i](k);
    if (indices)
        for (size_t i = 0; i < m->indices.size(); i++)
            indices[i] = (long)m->indices[i];
    return 1;
}

int ref_save_ply(void *h, const char *path) { return ((Ref *)h)->map->SaveAllMeshesToPLY(path) ? 1 : 0; }

void ref_last_counts(void *h, long *out)
{
    Ref *r = (Ref *)h;
    out[0] = r->lastCandidates;
    out[1] = r->lastNew;
    out[2] = r->lastGarbage;
}

void ref_sizeofs(int *out)
{
    out[0] = (int)sizeof(DistVoxel);
    out[1] = (int)sizeof(ColorVoxel);
    out[2] = (int)sizeof(Chunk);
}

} // extern "C"
