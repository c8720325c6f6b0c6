// ply_convert — host-only helper: reads any PLY the tools accept and rewrites it in the PCL
// binary layout (savePLYFileBinary).  No GPU involved; used by the CPU tests of the PLY
// reader/writer and handy for normalising capture-tool ASCII files.
#include <iostream>

#include "ply_io.hpp"

int main(int argc, char* argv[]) {
  if (argc != 3) {
    std::cerr << "usage: " << argv[0] << " in.ply out.ply" << std::endl;
    return -1;
  }
  lc3d_tools::Cloud c;
  std::string err;
  if (lc3d_tools::load_ply(argv[1], c, &err) != 0) {
    std::cerr << "Couldn't load input point cloud: " << argv[1] << " (" << err << ")" << std::endl;
    return -1;
  }
  std::cout << "points " << c.size() << " normals " << c.has_normals << " color " << c.has_color << " curvature "
            << c.has_curvature << " dense " << c.is_dense << std::endl;
  if (lc3d_tools::save_ply_binary(argv[2], c) != 0) {
    std::cerr << "Couldn't write " << argv[2] << std::endl;
    return -1;
  }
  return 0;
}
