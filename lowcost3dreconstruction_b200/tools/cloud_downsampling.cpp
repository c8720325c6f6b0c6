// cloud_downsampling — drop-in for pcl_tools/cloud_downsampling.cpp: pcl::VoxelGrid with a
// cubic leaf on the GPU via lc3d_voxel_grid.
#include "cli_common.hpp"

using namespace lc3d_tools;

int main(int argc, char* argv[]) {
  return run_tool([&]() -> int {
    Options opt("Options");
    opt.flag("help", 'h', "Print help message")
        .value("input", 'i', "Input file (.ply)")
        .value("output", 'o', "Output file (.ply)")
        .value("leaf_size", 's', "Leaf size for pcl::VoxelGrid filter in meters", "1");
    opt.parse(argc, argv);
    if (opt.count("help")) {
      std::cout << "Reduce the number of points of a point cloud, using a voxelized grid approach." << std::endl << std::endl;
      opt.print(std::cout);
      std::cout << std::endl;
      return 0;
    }
    if (!(opt.count("input") && opt.count("output")))
      throw std::logic_error("Correct mode of use: " + std::string(argv[0]) + " -i input.ply -o output.ply [opts]");
    const float leaf = opt.as<float>("leaf_size");
    Cloud cloud;
    if (load_ply(opt.str("input"), cloud) == -1)
      throw std::runtime_error("Couldn't load input point cloud: " + opt.str("input"));
    std::cout << "Cloud before filtering: " << std::endl;
    print_cloud_summary(std::cout, cloud);
    std::cout << std::endl;

    Cloud filtered;
    const size_t n = cloud.size();
    if (n > 0) {
      Ctx ctx;
      const lc3d_cloud c = as_lc3d(cloud);
      const float lf[3] = {leaf, leaf, leaf};
      std::vector<float> xyz(3 * n), nrm(3 * n), curv(n);
      std::vector<uint32_t> rgba(n);
      int64_t m = 0;
      ctx.check(lc3d_voxel_grid(ctx.h, &c, lf, xyz.data(), nrm.data(), rgba.data(), curv.data(), nullptr, &m));
      filtered.points.resize((size_t)m);
      for (size_t i = 0; i < (size_t)m; ++i) {
        Point& p = filtered.points[i];
        p = Point{};
        p.x = xyz[3 * i];
        p.y = xyz[3 * i + 1];
        p.z = xyz[3 * i + 2];
        p.w = 1.0f;
        p.nx = nrm[3 * i];
        p.ny = nrm[3 * i + 1];
        p.nz = nrm[3 * i + 2];
        p.rgba = rgba[i];
        p.curvature = curv[i];
      }
    }
    filtered.width = (uint32_t)filtered.points.size();
    filtered.height = 1;
    filtered.is_dense = true;
    std::cout << "Cloud after filtering: " << std::endl;
    print_cloud_summary(std::cout, filtered);
    std::cout << std::endl;
    if (save_ply_binary(opt.str("output"), filtered) != 0) throw std::runtime_error("Couldn't write " + opt.str("output"));
    return 0;
  });
}
