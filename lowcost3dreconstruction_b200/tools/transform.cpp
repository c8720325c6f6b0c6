// transform — drop-in for pcl_tools/transform.cpp (scripts/alignment.sh:107-112 applies each
// pair's 4x4 to all later views with it): reads a 4x4 text matrix, applies it to the points and
// rotates the normals by its 3x3 block (pcl::transformPointCloudWithNormals, transform.cpp:84-90)
// on the GPU via lc3d_transform.  With this tool the scripted chain of INTEGRATION.md runs on a
// box that has no PCL at all.
#include <fstream>

#include "cli_common.hpp"

using namespace lc3d_tools;

int main(int argc, char* argv[]) {
  return run_tool([&]() -> int {
    Options opt("Options");
    opt.flag("help", 'h', "Print help message")
        .value("input", 'i', "Input file (.ply)")
        .value("output", 'o', "Output file (.ply)")
        .value("transform", 't', "File containing a 4x4 transformation matrix");
    opt.parse(argc, argv);
    if (opt.count("help")) {
      std::cout << "Transforms Point cloud." << std::endl << std::endl;
      opt.print(std::cout);
      std::cout << std::endl;
      return 0;
    }
    if (!(opt.count("input") && opt.count("output") && opt.count("transform")))
      throw std::logic_error("Correct mode of use: " + std::string(argv[0]) +
                             " -i input.ply -o output.ply -t transform_file.txt");
    Cloud cloud;
    if (load_ply(opt.str("input"), cloud) == -1)
      throw std::runtime_error("Couldn't load input point cloud: " + opt.str("input"));
    // the matrix file: 16 whitespace-separated numbers, row-major (transform.cpp:68-81)
    const std::string tf = opt.str("transform");
    double m[16];
    std::ifstream file(tf);
    if (!file.is_open()) throw std::runtime_error("Unable to open file: " + tf);
    for (int k = 0; k < 16; ++k)
      if (!(file >> m[k])) throw std::runtime_error("Error on read transform file: " + tf);
    float mf[16];
    for (int k = 0; k < 16; ++k) mf[k] = (float)m[k];  // Eigen::Matrix4f(i, j) = matrix[i][j]
    const size_t n = cloud.size();
    if (n > 0) {
      Ctx ctx;
      const lc3d_cloud c = as_lc3d(cloud);
      std::vector<float> xyz(3 * n), nrm(3 * n);
      ctx.check(lc3d_transform(ctx.h, &c, mf, xyz.data(), nrm.data()));
      for (size_t k = 0; k < n; ++k) {
        Point& p = cloud.points[k];
        p.x = xyz[3 * k];
        p.y = xyz[3 * k + 1];
        p.z = xyz[3 * k + 2];
        p.nx = nrm[3 * k];
        p.ny = nrm[3 * k + 1];
        p.nz = nrm[3 * k + 2];
      }
    }
    if (save_ply_binary(opt.str("output"), cloud) != 0) throw std::runtime_error("Couldn't write " + opt.str("output"));
    return 0;
  });
}
