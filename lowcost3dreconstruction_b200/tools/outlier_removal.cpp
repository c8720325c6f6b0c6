// outlier_removal — drop-in for pcl_tools/outlier_removal.cpp: pcl::StatisticalOutlierRemoval
// on the GPU via lc3d_sor (second pass with negative=1 for --outliers_file, like the reference).
#include "cli_common.hpp"

using namespace lc3d_tools;

namespace {
Cloud select(const Cloud& in, const std::vector<int32_t>& idx, int64_t count) {
  Cloud out;
  out.points.reserve((size_t)count);
  for (int64_t i = 0; i < count; ++i) out.points.push_back(in.points[(size_t)idx[(size_t)i]]);
  out.width = (uint32_t)out.points.size();
  out.height = 1;
  out.is_dense = true;
  return out;
}
}  // namespace

int main(int argc, char* argv[]) {
  return run_tool([&]() -> int {
    Options opt("Options");
    opt.flag("help", 'h', "Print help message")
        .value("input", 'i', "Input file (.ply)")
        .value("output", 'o', "Output file (.ply)")
        .flag("outliers_file", 'f', "Saves the outliers in a ply file")
        .value("neighbors", 'n', "Neighbors to analyze for each point", "50")
        .value("dev_mult", 'd', "Standard deviation multiplier", "1");
    opt.parse(argc, argv);
    if (opt.count("help")) {
      std::cout << "Remove noisy measurements from a point cloud dataset using statistical analysis techniques." << std::endl << std::endl;
      opt.print(std::cout);
      std::cout << std::endl;
      return 0;
    }
    if (!(opt.count("input") && opt.count("output")))
      throw std::logic_error("Correct mode of use: " + std::string(argv[0]) + " -i input.ply -o output.ply [opts]");
    const unsigned k = opt.as_uint("neighbors");
    const double mult = opt.as<double>("dev_mult");
    std::string out_name = opt.str("output");
    Cloud cloud;
    if (load_ply(opt.str("input"), cloud) == -1)
      throw std::runtime_error("Couldn't load input point cloud: " + opt.str("input"));
    std::cout << "Cloud before filtering: " << std::endl;
    print_cloud_summary(std::cout, cloud);
    std::cout << std::endl;

    Ctx ctx;
    const lc3d_cloud c = as_lc3d(cloud, false);
    std::vector<int32_t> kept(cloud.size() + 1);
    int64_t count = 0;
    if (cloud.size() > 0) ctx.check(lc3d_sor(ctx.h, &c, (int32_t)k, mult, 0, kept.data(), &count, nullptr, nullptr));
    Cloud filtered = select(cloud, kept, count);
    std::cout << "Cloud after filtering: " << std::endl;
    print_cloud_summary(std::cout, filtered);
    std::cout << std::endl;
    if (save_ply_binary(out_name, filtered) != 0) throw std::runtime_error("Couldn't write " + out_name);

    if (opt.count("outliers_file")) {
      count = 0;
      if (cloud.size() > 0) ctx.check(lc3d_sor(ctx.h, &c, (int32_t)k, mult, 1, kept.data(), &count, nullptr, nullptr));
      const size_t pos = out_name.rfind(".ply");
      if (pos != std::string::npos) out_name.erase(pos, 4);
      if (save_ply_binary(out_name + "_outliers.ply", select(cloud, kept, count)) != 0)
        throw std::runtime_error("Couldn't write " + out_name + "_outliers.ply");
    }
    return 0;
  });
}
