// open_chisel/Chunk.h -- host mirror of one chunk of the device pool, filled on demand by ChunkManager.
// Same read API as the reference (OC/include/open_chisel/Chunk.h:47-140) for the consumers in chisel_ros
// (CR/include/chisel_ros/Serialization.h:31-84, CR/src/ChiselServer.cpp:534-603).
#ifndef CHISEL_B200_CHUNK_H_
#define CHISEL_B200_CHUNK_H_
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
    Chunk() : ID(0, 0, 0), numVoxels(0, 0, 0), voxelResolutionMeters(0), origin(Vec3::Zero()) {}
    Chunk(const ChunkID &id, const Eigen::Vector3i &nv, float r, bool useColor) : ID(id), numVoxels(nv), voxelResolutionMeters(r)
    {
        voxels.resize(GetTotalNumVoxels());
        if (useColor)
            colors.resize(GetTotalNumVoxels());
        // Chunk.cpp:43: integer product first, then one multiplication per axis
        origin = Vec3(numVoxels(0) * ID(0) * voxelResolutionMeters, numVoxels(1) * ID(1) * voxelResolutionMeters, numVoxels(2) * ID(2) * voxelResolutionMeters);
    }
    const ChunkID &GetID() const { return ID; }
    bool HasColors() const { return !colors.empty(); }
    bool HasVoxels() const { return !voxels.empty(); }
    const std::vector<DistVoxel> &GetVoxels() const { return voxels; }
    std::vector<DistVoxel> &GetMutableVoxels() { return voxels; }
    const std::vector<ColorVoxel> &GetColorVoxels() const { return colors; }
    std::vector<ColorVoxel> &GetMutableColorVoxels() { return colors; }
    const Eigen::Vector3i &GetNumVoxels() const { return numVoxels; }
    float GetVoxelResolutionMeters() const { return voxelResolutionMeters; }
    size_t GetTotalNumVoxels() const { return static_cast<size_t>(numVoxels(0)) * numVoxels(1) * numVoxels(2); }
    VoxelID GetVoxelID(int x, int y, int z) const { return (z * numVoxels(2) + y) * numVoxels(0) + x; }   // Chunk.h:81-84 (Q12)
    const DistVoxel &GetDistVoxel(const VoxelID &id) const { return voxels.at(id); }
    const DistVoxel &GetDistVoxel(int x, int y, int z) const { return voxels.at(GetVoxelID(x, y, z)); }
    const ColorVoxel &GetColorVoxel(const VoxelID &id) const { return colors.at(id); }
    const ColorVoxel &GetColorVoxel(int x, int y, int z) const { return colors.at(GetVoxelID(x, y, z)); }
    bool IsCoordValid(int x, int y, int z) const { return x >= 0 && x < numVoxels(0) && y >= 0 && y < numVoxels(1) && z >= 0 && z < numVoxels(2); }
    const Vec3 &GetOrigin() const { return origin; }
    AABB ComputeBoundingBox() const { return AABB(origin, origin + numVoxels.cast<float>() * voxelResolutionMeters); }

  protected:
    ChunkID ID;
    Eigen::Vector3i numVoxels;
    float voxelResolutionMeters;
    std::vector<DistVoxel> voxels;
    std::vector<ColorVoxel> colors;
    Vec3 origin;
};
typedef std::shared_ptr<Chunk> ChunkPtr;
typedef std::shared_ptr<const Chunk> ChunkConstPtr;
} // namespace chisel
#endif
