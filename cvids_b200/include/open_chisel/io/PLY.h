// open_chisel/io/PLY.h -- ASCII PLY export with the reference's layout (OC/src/io/PLY.cpp:29-88): vertices (+ uchar colours),
// then one triangle per three vertices.
#ifndef CHISEL_B200_PLY_H_
#define CHISEL_B200_PLY_H_
#include <fstream>
#include <string>
#include <open_chisel/mesh/Mesh.h>
namespace chisel
{
inline bool SaveMeshPLYASCII(const std::string &fileName, const MeshConstPtr &mesh)
{
    std::ofstream out(fileName.c_str());
    if (!out)
        return false;
    const size_t n = mesh->vertices.size();
    const bool colored = mesh->HasColors();
    out << "ply\nformat ascii 1.0\nelement vertex " << n << "\nproperty float x\nproperty float y\nproperty float z\n";
    if (colored)
        out << "property uchar red\nproperty uchar green\nproperty uchar blue\n";
    out << "element face " << n / 3 << "\nproperty list uchar int vertex_index\nend_header\n";
    for (size_t i = 0; i < n; i++)
    {
        const Vec3 &v = mesh->vertices[i];
        out << v(0) << " " << v(1) << " " << v(2);
        if (colored)
        {
            const Vec3 &c = mesh->colors[i];
            out << " " << static_cast<int>(c(0) * 255.0f) << " " << static_cast<int>(c(1) * 255.0f) << " " << static_cast<int>(c(2) * 255.0f);
        }
        out << "\n";
    }
    for (size_t i = 0; i + 2 < mesh->indices.size(); i += 3)
        out << "3 " << mesh->indices[i] << " " << mesh->indices[i + 1] << " " << mesh->indices[i + 2] << " \n";
    return true;
}
} // namespace chisel
#endif
