// open_chisel/ChunkManager.h -- facade over the device-resident map (C ABI: chisel_b200.h).
//
// The reference's ChunkManager IS the map (an unordered_map of heap chunks, OC/include/open_chisel/ChunkManager.h:58-223).
// Here the map lives in HBM; this class owns the chs_map handle and serves the reference's read API from host mirrors
// that are refreshed lazily, off the timed path: GetChunks() / GetChunk() download voxels on demand, GetAllMeshes() is
// the MeshMap that Chisel::UpdateMeshes maintains with the reference's publication rule (ChunkManager.cpp:101-127).
#ifndef CHISEL_B200_CHUNKMANAGER_H_
#define CHISEL_B200_CHUNKMANAGER_H_
#include <cstring>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>
#include <chisel_b200.h>
#include <open_chisel/Chunk.h>
#include <open_chisel/b200/PinnedBuffer.h>
#include <open_chisel/geometry/Geometry.h>
#include <open_chisel/mesh/Mesh.h>

namespace chisel
{
// Teschner's three-prime spatial hash, as in the reference (ChunkManager.h:40-51)
struct ChunkHasher
{
    static constexpr size_t p1 = 73856093;
    static constexpr size_t p2 = 19349663;
    static constexpr size_t p3 = 8349279;
    std::size_t operator()(const ChunkID &key) const { return (key(0) * p1 ^ key(1) * p2 ^ key(2) * p3); }
};
typedef std::unordered_map<ChunkID, ChunkPtr, ChunkHasher> ChunkMap;
typedef std::unordered_map<ChunkID, bool, ChunkHasher> ChunkSet;
typedef std::unordered_map<ChunkID, MeshPtr, ChunkHasher> MeshMap;

namespace b200
{
inline void Check(int rc, const char *what)
{
    if (rc != CHS_OK)
        throw std::runtime_error(std::string(what) + ": " + chs_last_error_string());
}
} // namespace b200

class ChunkManager
{
  public:
    ChunkManager() : chunkSize(16, 16, 16), voxelResolutionMeters(0.03f), useColor(false), version(0), chunksVersion(-1), indexVersion(-1), indexCount(0) {}   // Q15: useColor initialised
    ChunkManager(const Eigen::Vector3i &size, float res, bool color, int device = -1, int rank = 0, int world = 1, void *stream = nullptr)
        : chunkSize(size), voxelResolutionMeters(res), useColor(color), version(0), chunksVersion(-1), indexVersion(-1), indexCount(0)
    {
        if (size(0) != size(1) || size(0) != size(2))
            throw std::invalid_argument("chisel_b200: cubic chunks only (Chunk::GetVoxelID is only correct for them, quirk Q12)");
        chs_config cfg;
        cfg.chunk_size = size(0);
        cfg.resolution = res;
        cfg.use_color = color ? 1 : 0;
        cfg.device = device;
        cfg.rank = rank;
        cfg.world = world;
        cfg.initial_chunks = 0;
        cfg.stream = stream;
        chs_map *h = nullptr;
        b200::Check(chs_create(&cfg, &h), "chs_create");
        handle.reset(h, [](chs_map *p) { chs_destroy(p); });
        CacheCentroids();
    }

    chs_map *Handle() const { return handle.get(); }
    void Touch() { version++; }                                  // the device map changed: host mirrors are stale
    // Frame batching (Chisel::SetFrameBatching): frames queued in the facade must reach the device before anything reads it.
    void SetBeforeDeviceRead(const std::function<void()> &f) { beforeDeviceRead = f; }
    void SetBeforeBulkRead(const std::function<void()> &f) { beforeBulkRead = f; }
    long Version() const { return version; }
    void Sync() const
    {
        if (beforeDeviceRead)
            beforeDeviceRead();
    }

    const Eigen::Vector3i &GetChunkSize() const { return chunkSize; }
    float GetResolution() const { return voxelResolutionMeters; }
    bool GetUseColor() const { return useColor; }
    const Vec3List &GetCentroids() const { return centroids; }

    // Host index of the chunk IDs of the device map, extended lazily (pool slots are append-only between resets): HasChunk is a
    // hash lookup, not a device round trip -- chisel_ros asks it for every dirty ID after every frame (CR ChiselServer.cpp:554-559).
    void RefreshIndex() const
    {
        if (indexVersion == version)
            return;
        int64_t n = 0;
        b200::Check(chs_num_chunks(handle.get(), &n), "chs_num_chunks");
        if (n < indexCount)
        {
            idIndex.clear();
            indexCount = 0;
        }
        if (n > indexCount)
        {
            std::vector<int32_t> ids(3 * n);
            b200::Check(chs_chunk_ids(handle.get(), ids.data(), n), "chs_chunk_ids");
            for (int64_t i = indexCount; i < n; i++)
                idIndex[ChunkID(ids[3 * i], ids[3 * i + 1], ids[3 * i + 2])] = true;
            indexCount = n;
        }
        indexVersion = version;
    }

    bool HasChunk(const ChunkID &id) const
    {
        Sync();
        RefreshIndex();
        return idIndex.find(id) != idIndex.end();
    }
    bool HasChunk(int x, int y, int z) const { return HasChunk(ChunkID(x, y, z)); }

    // unordered_map::at semantics: throws std::out_of_range for a missing chunk (ChunkManager.h:84-87). The mirror object is
    // cached; its voxels are downloaded on first access and again after the device map has changed (Chunk::Invalidate).
    ChunkPtr GetChunk(const ChunkID &id) const
    {
        Sync();
        RefreshIndex();
        if (idIndex.find(id) == idIndex.end())
            throw std::out_of_range("ChunkManager::GetChunk: no such chunk");
        ChunkPtr &c = chunks[id];
        if (!c)
        {
            c = std::make_shared<Chunk>(id, chunkSize, voxelResolutionMeters, useColor, true);
            std::shared_ptr<chs_map> h = handle;                 // not `this`: the manager may be copied (Chisel::SetChunkManager)
            const bool color = useColor;
            c->SetLoader([h, color](Chunk *ch)
                         {
                             const size_t V = ch->GetTotalNumVoxels();
                             std::vector<float> sdf(V), w(V);
                             std::vector<uint8_t> rgbw(color ? 4 * V : 0);
                             const int32_t cid[3] = {ch->GetID()(0), ch->GetID()(1), ch->GetID()(2)};
                             b200::Check(chs_download_chunk(h.get(), cid, sdf.data(), w.data(), color ? rgbw.data() : nullptr), "chs_download_chunk");
                             FillChunk(ch, sdf.data(), w.data(), color ? rgbw.data() : nullptr);
                         });
        }
        c->Invalidate(version);
        return c;
    }
    ChunkPtr GetChunk(int x, int y, int z) const { return GetChunk(ChunkID(x, y, z)); }
    ChunkID GetIDAt(const Vec3 &pos) const
    {
        // ChunkManager.h:136-145 with per-instance factors (identical to the reference's first instance, quirk Q1)
        const float rx = 1.0f / (chunkSize(0) * voxelResolutionMeters), ry = 1.0f / (chunkSize(1) * voxelResolutionMeters), rz = 1.0f / (chunkSize(2) * voxelResolutionMeters);
        return ChunkID(static_cast<int>(std::floor(pos(0) * rx)), static_cast<int>(std::floor(pos(1) * ry)), static_cast<int>(std::floor(pos(2) * rz)));
    }

    // Whole-map host mirror (one bulk transfer), refreshed only if the device map changed since the last call.
    const ChunkMap &GetChunks() const
    {
        if (beforeBulkRead)
            beforeBulkRead();
        Sync();
        if (chunksVersion != version)
        {
            chunks.clear();
            int64_t n = 0;
            b200::Check(chs_num_chunks(handle.get(), &n), "chs_num_chunks");
            const size_t V = static_cast<size_t>(chunkSize(0)) * chunkSize(1) * chunkSize(2);
            std::vector<int32_t> ids(3 * n);
            std::vector<float> sdf(n * V), w(n * V);
            std::vector<uint8_t> rgbw(useColor ? n * V * 4 : 0);
            if (n)
                b200::Check(chs_download_all(handle.get(), n, ids.data(), sdf.data(), w.data(), useColor ? rgbw.data() : nullptr), "chs_download_all");
            for (int64_t i = 0; i < n; i++)
            {
                const ChunkID id(ids[3 * i], ids[3 * i + 1], ids[3 * i + 2]);
                ChunkPtr c = std::make_shared<Chunk>(id, chunkSize, voxelResolutionMeters, useColor);
                FillChunk(c.get(), sdf.data() + i * V, w.data() + i * V, useColor ? rgbw.data() + i * V * 4 : nullptr);
                chunks[id] = c;
            }
            chunksVersion = version;
        }
        return chunks;
    }

    const MeshMap &GetAllMeshes() const { return allMeshes; }
    MeshMap &GetAllMutableMeshes() { return allMeshes; }
    const MeshPtr &GetMesh(const ChunkID &id) const { return allMeshes.at(id); }
    bool HasMesh(const ChunkID &id) const { return allMeshes.find(id) != allMeshes.end(); }

    // ChunkManager::RecomputeMeshes (ChunkManager.cpp:130-169): the set that is re-meshed is the DEVICE dirty set, which
    // is what Chisel passes (meshesToUpdate); it is cleared afterwards like Chisel.cpp:57. New meshes are published only
    // if they have occupied cells; meshes that already exist are replaced even if they became empty (quirk Q10).
    void RecomputeMeshes(const ChunkSet & /*dirty: the device holds the authoritative set*/) { RecomputeDirtyMeshes(); }
    void RecomputeDirtyMeshes()
    {
        Sync();
        b200::Check(chs_update_meshes(handle.get()), "chs_update_meshes");
        chs_mesh_counts mc;
        b200::Check(chs_mesh_counts_last(handle.get(), &mc), "chs_mesh_counts_last");
        std::vector<int32_t> ids(3 * mc.n_chunks);
        std::vector<int64_t> voff(mc.n_chunks + 1), goff(mc.n_chunks + 1);
        // vertex arrays land in page-locked staging (kept between re-meshes): the download runs at full PCIe speed
        const size_t nvf = 3 * static_cast<size_t>(mc.n_vertices), ngf = 3 * static_cast<size_t>(mc.n_grids);
        const size_t need = nvf * (mc.has_colors ? 3 : 2) + ngf + 16;
        if (!meshStage || meshStage->size() < need)
            meshStage.reset(new b200::PinnedBuffer<float>(need + need / 4));
        float *v = meshStage->data(), *nr = v + nvf, *col = nr + nvf, *g = col + (mc.has_colors ? nvf : 0);
        b200::Check(chs_download_meshes(handle.get(), ids.data(), voff.data(), goff.data(), v, nr, mc.has_colors ? col : nullptr, g), "chs_download_meshes");
        // publication rule first (serial: it touches the MeshMap), then the vertex arrays are copied by a few threads -- a re-mesh of a
        // room-sized dirty set is tens of MB of host copies, the largest host cost of the facade
        std::vector<std::pair<int64_t, Mesh *>> jobs;
        jobs.reserve(static_cast<size_t>(mc.n_chunks));
        for (int64_t i = 0; i < mc.n_chunks; i++)
        {
            const ChunkID id(ids[3 * i], ids[3 * i + 1], ids[3 * i + 2]);
            MeshMap::iterator it = allMeshes.find(id);
            const bool had = it != allMeshes.end();
            if (!had && goff[i + 1] == goff[i])
                continue;
            MeshPtr m = had ? it->second : std::make_shared<Mesh>();
            if (!had)
                allMeshes[id] = m;
            jobs.push_back(std::make_pair(i, m.get()));
        }
        const bool colors = mc.has_colors != 0;
        auto fill = [&](size_t lo, size_t hi)
        {
            for (size_t j = lo; j < hi; j++)
            {
                const int64_t i = jobs[j].first;
                Mesh *m = jobs[j].second;
                const size_t nv = static_cast<size_t>(voff[i + 1] - voff[i]), ng = static_cast<size_t>(goff[i + 1] - goff[i]);
                m->Clear();
                if (sizeof(Vec3) == 3 * sizeof(float))
                {
                    // Vec3 is three packed floats: whole arrays at once, one pass over the memory
                    const Vec3 *pv = reinterpret_cast<const Vec3 *>(v + 3 * voff[i]), *pn = reinterpret_cast<const Vec3 *>(nr + 3 * voff[i]);
                    m->vertices.assign(pv, pv + nv);
                    m->normals.assign(pn, pn + nv);
                    if (colors)
                    {
                        const Vec3 *pc = reinterpret_cast<const Vec3 *>(col + 3 * voff[i]);
                        m->colors.assign(pc, pc + nv);
                    }
                    const Vec3 *pg = reinterpret_cast<const Vec3 *>(g + 3 * goff[i]);
                    m->grids.assign(pg, pg + ng);
                    m->indices.resize(nv);
                    for (size_t k = 0; k < nv; k++)
                        m->indices[k] = k;
                }
                else
                {
                    m->Resize(nv, ng, colors);
                    for (size_t k = 0; k < nv; k++)
                    {
                        const int64_t q = voff[i] + static_cast<int64_t>(k);
                        m->vertices[k] = Vec3(v[3 * q], v[3 * q + 1], v[3 * q + 2]);
                        m->normals[k] = Vec3(nr[3 * q], nr[3 * q + 1], nr[3 * q + 2]);
                        if (colors)
                            m->colors[k] = Vec3(col[3 * q], col[3 * q + 1], col[3 * q + 2]);
                    }
                    for (size_t k = 0; k < ng; k++)
                        m->grids[k] = Vec3(g[3 * (goff[i] + k)], g[3 * (goff[i] + k) + 1], g[3 * (goff[i] + k) + 2]);
                }
            }
        };
        const size_t nThreads = mc.n_vertices > 200000 ? 4 : 1;
        if (nThreads == 1)
            fill(0, jobs.size());
        else
        {
            std::vector<std::thread> pool;
            for (size_t t = 0; t < nThreads; t++)
                pool.emplace_back(fill, jobs.size() * t / nThreads, jobs.size() * (t + 1) / nThreads);
            for (std::thread &th : pool)
                th.join();
        }
    }

    void Reset()
    {
        b200::Check(chs_reset(handle.get()), "chs_reset");
        allMeshes.clear();
        chunks.clear();
        idIndex.clear();
        indexCount = 0;
        Touch();
    }

    // ChunkManager::CacheCentroids (ChunkManager.cpp:50-65)
    void CacheCentroids()
    {
        const Vec3 half = Vec3(voxelResolutionMeters, voxelResolutionMeters, voxelResolutionMeters) * 0.5f;
        centroids.resize(static_cast<size_t>(chunkSize(0)) * chunkSize(1) * chunkSize(2));
        size_t i = 0;
        for (int z = 0; z < chunkSize(2); z++)
            for (int y = 0; y < chunkSize(1); y++)
                for (int x = 0; x < chunkSize(0); x++)
                    centroids[i++] = Vec3(x, y, z) * voxelResolutionMeters + half;
    }

  protected:
    static void FillChunk(Chunk *c, const float *sdf, const float *w, const uint8_t *rgbw)
    {
        std::vector<DistVoxel> &dv = c->GetMutableVoxels();
        for (size_t i = 0; i < dv.size(); i++)
            dv[i] = DistVoxel(sdf[i], w[i]);
        if (rgbw)
        {
            std::vector<ColorVoxel> &cv = c->GetMutableColorVoxels();
            for (size_t i = 0; i < cv.size(); i++)
                cv[i] = ColorVoxel(rgbw[4 * i], rgbw[4 * i + 1], rgbw[4 * i + 2], rgbw[4 * i + 3]);
        }
    }

    std::shared_ptr<chs_map> handle;
    Eigen::Vector3i chunkSize;
    float voxelResolutionMeters;
    bool useColor;
    Vec3List centroids;
    MeshMap allMeshes;
    long version;
    mutable ChunkMap chunks;
    mutable long chunksVersion;
    mutable ChunkSet idIndex;                 // IDs of the chunks of the device map (RefreshIndex)
    mutable long indexVersion;
    mutable int64_t indexCount;
    std::function<void()> beforeDeviceRead, beforeBulkRead;
    std::shared_ptr<b200::PinnedBuffer<float>> meshStage;
};
typedef std::shared_ptr<ChunkManager> ChunkManagerPtr;
typedef std::shared_ptr<const ChunkManager> ChunkManagerConstPtr;
} // namespace chisel
#endif
