// open_chisel/mesh/Mesh.h -- facade. One chunk's triangle soup as chisel_ros reads it (CR/src/ChiselServer.cpp:658-714): positions,
// per-vertex normals and colours, trivial indices 0..n-1, and the corner-0 position of every occupied cell (`grids`). Same member
// and method names as OC/include/open_chisel/mesh/Mesh.h:33-60; the arrays are filled in bulk from the device download
// (ChunkManager::RecomputeDirtyMeshes), hence the sizing helpers.
#ifndef CHISEL_B200_MESH_H_
#define CHISEL_B200_MESH_H_
#include <cstddef>
#include <memory>
#include <numeric>
#include <vector>
#include <open_chisel/geometry/Geometry.h>

namespace chisel
{
typedef size_t VertIndex;
typedef std::vector<VertIndex> VertIndexList;

class Mesh
{
  public:
    // data, in the order the exporters walk it
    Vec3List vertices;
    VertIndexList indices;
    Vec3List normals;
    Vec3List colors;
    Vec3List grids;

    bool HasVertices() const { return vertices.size() != 0; }
    bool HasNormals() const { return normals.size() != 0; }
    bool HasColors() const { return colors.size() != 0; }
    bool HasIndices() const { return indices.size() != 0; }
    size_t NumTriangles() const { return vertices.size() / 3; }

    // Size every per-vertex array for `nVertices` (colours only if the map has them), `nGrids` occupied cells, and write the
    // trivial index list MarchingCubes::MeshCube produces (MarchingCubes.h:91-93).
    void Resize(size_t nVertices, size_t nGrids, bool withColors)
    {
        vertices.resize(nVertices);
        normals.resize(nVertices);
        colors.resize(withColors ? nVertices : 0);
        grids.resize(nGrids);
        indices.resize(nVertices);
        std::iota(indices.begin(), indices.end(), static_cast<VertIndex>(0));
    }

    void Clear()
    {
        for (Vec3List *l : {&vertices, &normals, &colors, &grids})
            l->clear();
        indices.clear();
    }
};
typedef std::shared_ptr<Mesh> MeshPtr;
typedef std::shared_ptr<const Mesh> MeshConstPtr;
} // namespace chisel
#endif
