// open_chisel/io/PLY.h -- ASCII PLY export with the reference's layout (OC/src/io/PLY.cpp:29-88): vertices (+ uchar colours),
// then one triangle per three vertices. SaveMeshPLYBinary (not in the reference) writes the same elements
// as binary_little_endian: the export of a large map is then bounded by the disk, not by number formatting.
#ifndef CHISEL_B200_PLY_H_
#define CHISEL_B200_PLY_H_
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>
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
// Same header fields and element order as SaveMeshPLYASCII; colours are quantised the same way (static_cast<int>(c * 255.0f)).
inline bool SaveMeshPLYBinary(const std::string &fileName, const MeshConstPtr &mesh)
{
    FILE *out = std::fopen(fileName.c_str(), "wb");
    if (!out)
        return false;
    const size_t n = mesh->vertices.size();
    const bool colored = mesh->HasColors();
    std::fprintf(out, "ply\nformat binary_little_endian 1.0\nelement vertex %zu\nproperty float x\nproperty float y\nproperty float z\n", n);
    if (colored)
        std::fprintf(out, "property uchar red\nproperty uchar green\nproperty uchar blue\n");
    std::fprintf(out, "element face %zu\nproperty list uchar int vertex_index\nend_header\n", n / 3);
    const size_t vstride = 12 + (colored ? 3 : 0);
    std::vector<unsigned char> buf(n * vstride);
    for (size_t i = 0; i < n; i++)
    {
        const float xyz[3] = {mesh->vertices[i](0), mesh->vertices[i](1), mesh->vertices[i](2)};
        std::memcpy(&buf[i * vstride], xyz, 12);
        if (colored)
            for (int k = 0; k < 3; k++)
                buf[i * vstride + 12 + k] = static_cast<unsigned char>(static_cast<int>(mesh->colors[i](k) * 255.0f));
    }
    bool ok = buf.empty() || std::fwrite(buf.data(), 1, buf.size(), out) == buf.size();
    const size_t nf = mesh->indices.size() / 3;
    std::vector<unsigned char> faces(nf * 13);
    for (size_t f = 0; f < nf; f++)
    {
        faces[f * 13] = 3;
        const int32_t idx[3] = {static_cast<int32_t>(mesh->indices[3 * f]), static_cast<int32_t>(mesh->indices[3 * f + 1]), static_cast<int32_t>(mesh->indices[3 * f + 2])};
        std::memcpy(&faces[f * 13 + 1], idx, 12);
    }
    ok = ok && (faces.empty() || std::fwrite(faces.data(), 1, faces.size(), out) == faces.size());
    return std::fclose(out) == 0 && ok;
}
} // namespace chisel
#endif
