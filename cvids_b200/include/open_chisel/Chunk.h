// open_chisel/Chunk.h -- host mirror of one chunk of the device pool, filled on demand by ChunkManager.
// Same read API as the reference (OC/include/open_chisel/Chunk.h:47-140) for the consumers in chisel_ros
// (CR/include/chisel_ros/Serialization.h:31-84, CR/src/ChiselServer.cpp:534-603).
#ifndef CHISEL_B200_CHUNK_H_
#define CHISEL_B200_CHUNK_H_
#include <functional>
#include <memory>
#include <vector>
#include <open_chisel/ColorVoxel.h>
#include <open_chisel/DistVoxel.h>
#include <open_chisel/geometry/AABB.h>
#include <open_chisel/geometry/Geometry.h>

namespace chisel
{
typedef Eigen::Vector3i ChunkID;
typedef int VoxelID;
typedef std::vector<ChunkID, Eigen::aligned_allocator<ChunkID>> ChunkIDList;

class Chunk
{
  public:
    // The voxel payload of a mirror may be LAZY: geometry (ID, origin, bounding box -- what CR ChiselServer.cpp:554-567 asks of every
    // dirty chunk after every frame) is there at once, the 48 KB of voxels are downloaded on the first access to them.
    typedef std::function<void(Chunk *)> Loader;
    Chunk() : ID(0, 0, 0), numVoxels(0, 0, 0), voxelResolutionMeters(0), origin(Vec3::Zero()), hasColor(false), loaded(true), seenVersion(-1) {}
    Chunk(const ChunkID &id, const Eigen::Vector3i &nv, float r, bool useColor, bool lazy = false)
        : ID(id), numVoxels(nv), voxelResolutionMeters(r), hasColor(useColor), loaded(!lazy), seenVersion(-1)
    {
        if (!lazy)
            Allocate();
        // Chunk.cpp:43: integer product first, then one multiplication per axis
        origin = Vec3(numVoxels(0) * ID(0) * voxelResolutionMeters, numVoxels(1) * ID(1) * voxelResolutionMeters, numVoxels(2) * ID(2) * voxelResolutionMeters);
    }
    void SetLoader(const Loader &l) { loader = l; }
    void Invalidate(long version)
    {
        if (loader && version != seenVersion)
            loaded = false;
        seenVersion = version;
    }
    void Allocate()
    {
        voxels.resize(GetTotalNumVoxels());
        if (hasColor)
            colors.resize(GetTotalNumVoxels());
    }
    const ChunkID &GetID() const { return ID; }
    bool HasColors() const { return hasColor; }
    bool HasVoxels() const { return true; }
    const std::vector<DistVoxel> &GetVoxels() const { Ensure(); return voxels; }
    std::vector<DistVoxel> &GetMutableVoxels() { Ensure(); return voxels; }
    const std::vector<ColorVoxel> &GetColorVoxels() const { Ensure(); return colors; }
    std::vector<ColorVoxel> &GetMutableColorVoxels() { Ensure(); return colors; }
    const Eigen::Vector3i &GetNumVoxels() const { return numVoxels; }
    float GetVoxelResolutionMeters() const { return voxelResolutionMeters; }
    size_t GetTotalNumVoxels() const { return static_cast<size_t>(numVoxels(0)) * numVoxels(1) * numVoxels(2); }
    VoxelID GetVoxelID(int x, int y, int z) const { return (z * numVoxels(2) + y) * numVoxels(0) + x; }   // Chunk.h:81-84 (Q12)
    const DistVoxel &GetDistVoxel(const VoxelID &id) const { Ensure(); return voxels.at(id); }
    const DistVoxel &GetDistVoxel(int x, int y, int z) const { Ensure(); return voxels.at(GetVoxelID(x, y, z)); }
    const ColorVoxel &GetColorVoxel(const VoxelID &id) const { Ensure(); return colors.at(id); }
    const ColorVoxel &GetColorVoxel(int x, int y, int z) const { Ensure(); return colors.at(GetVoxelID(x, y, z)); }
    bool IsCoordValid(int x, int y, int z) const { return x >= 0 && x < numVoxels(0) && y >= 0 && y < numVoxels(1) && z >= 0 && z < numVoxels(2); }
    const Vec3 &GetOrigin() const { return origin; }
    AABB ComputeBoundingBox() const { return AABB(origin, origin + numVoxels.cast<float>() * voxelResolutionMeters); }

  protected:
    void Ensure() const
    {
        if (!loaded)
        {
            loaded = true;
            Chunk *self = const_cast<Chunk *>(this);
            self->Allocate();
            if (loader)
                loader(self);
        }
    }
    ChunkID ID;
    Eigen::Vector3i numVoxels;
    float voxelResolutionMeters;
    mutable std::vector<DistVoxel> voxels;
    mutable std::vector<ColorVoxel> colors;
    Vec3 origin;
    bool hasColor;
    mutable bool loaded;
    long seenVersion;
    Loader loader;
};
typedef std::shared_ptr<Chunk> ChunkPtr;
typedef std::shared_ptr<const Chunk> ChunkConstPtr;
} // namespace chisel
#endif
