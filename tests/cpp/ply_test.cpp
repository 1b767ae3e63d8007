// ply_test.cpp -- host-only check of the facade's PLY writers (no device): the same mesh as ASCII (reference layout) and as
// binary_little_endian. tests/test_facade.py parses both and compares them.   ply_test <out.ascii.ply> <out.binary.ply>
#include <open_chisel/io/PLY.h>
#include <open_chisel/mesh/Mesh.h>

#include <memory>

int main(int argc, char **argv)
{
    if (argc < 3)
        return 2;
    chisel::MeshPtr m = std::make_shared<chisel::Mesh>();
    for (int t = 0; t < 7; t++)
        for (int k = 0; k < 3; k++)
        {
            m->vertices.push_back(chisel::Vec3(0.125f * t + 0.5f * k, -1.75f + 0.25f * t, 0.0625f * (t * 3 + k)));
            m->colors.push_back(chisel::Vec3(0.1f * k + 0.05f * t, 1.0f - 0.1f * t, 0.5f));
            m->normals.push_back(chisel::Vec3(0.0f, 0.0f, 1.0f));
            m->indices.push_back(static_cast<chisel::VertIndex>(3 * t + k));
        }
    return (chisel::SaveMeshPLYASCII(argv[1], m) && chisel::SaveMeshPLYBinary(argv[2], m)) ? 0 : 1;
}
