// normal_estimation — drop-in for pcl_tools/normal_estimation.cpp: k-NN PCA normals
// (pcl::NormalEstimation semantics) on the GPU via lc3d_normals / lc3d_centroid.
#include "cli_common.hpp"

using namespace lc3d_tools;

int main(int argc, char* argv[]) {
  return run_tool([&]() -> int {
    Options opt("Options");
    opt.flag("help", 'h', "Print help message")
        .value("input", 'i', "Input file (.ply)")
        .value("output", 'o', "Output file (.ply)")
        .value("neighbors", 'n', "N. of neighbors to analyze for each point", "50")
        .flag("reverse_normals", 'r', "Reverse normals' direction")
        .flag("centroid", 'c', "Use the centroid as defined view point")
        .flag("origin", 'z', "Use the origin as defined view point");
    opt.parse(argc, argv);
    if (opt.count("help")) {
      std::cout << "Estimate a set of normals for all the points in the input dataset." << std::endl << std::endl;
      opt.print(std::cout);
      std::cout << std::endl;
      return 0;
    }
    if (!(opt.count("input") && opt.count("output")))
      throw std::logic_error("Correct mode of use: " + std::string(argv[0]) + " -i input.ply -o output.ply [opts]");
    const unsigned neighbors = opt.as_uint("neighbors");
    const bool reverse = opt.count("reverse_normals"), centroid = opt.count("centroid"), origin = opt.count("origin");
    if (origin && centroid)
      throw std::runtime_error("It is not possible to use the centroid and origin as viewpoint at the same time");
    Cloud cloud;
    if (load_ply(opt.str("input"), cloud) == -1)
      throw std::runtime_error("Couldn't load input point cloud: " + opt.str("input"));

    Ctx ctx;
    const lc3d_cloud c = as_lc3d(cloud, false);
    float vp[3] = {0.f, 0.f, 0.f};  // default view point, also the sensor origin of a plain PLY (--origin)
    if (centroid) {
      float cen[4];
      ctx.check(lc3d_centroid(ctx.h, &c, cen));
      vp[0] = cen[0];
      vp[1] = cen[1];
      vp[2] = cen[2];
    }
    std::vector<float> nrm(3 * cloud.size() + 3), curv(cloud.size() + 1);
    if (cloud.size() > 0) ctx.check(lc3d_normals(ctx.h, &c, (int32_t)neighbors, vp, nrm.data(), curv.data()));
    // the centroid view point flips normals inward; the tool turns them outward again
    const float sign = ((centroid && !reverse) || (!centroid && reverse)) ? -1.0f : 1.0f;
    for (size_t i = 0; i < cloud.size(); ++i) {
      Point& p = cloud.points[i];
      p.nx = nrm[3 * i] * sign;
      p.ny = nrm[3 * i + 1] * sign;
      p.nz = nrm[3 * i + 2] * sign;
      p.curvature = curv[i];
    }
    if (save_ply_binary(opt.str("output"), cloud) != 0) throw std::runtime_error("Couldn't write " + opt.str("output"));
    return 0;
  });
}
