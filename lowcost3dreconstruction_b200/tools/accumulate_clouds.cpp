// accumulate_clouds — drop-in for pcl_tools/accumulate_clouds.cpp: merges a source cloud into a
// target cloud; unless --all, source points redundant with the target (inside some target
// point's +/- radius box: the reference's per-target-point pcl::CropBox loop, O(N*M), here one
// O(N) grid query via lc3d_box_dedup) are dropped and the remainder is cleaned with
// StatisticalOutlierRemoval (lc3d_sor) first.  Output = target points followed by the kept
// source points, like `*accumulated += *cloud_src`.
#include "cli_common.hpp"

using namespace lc3d_tools;

int main(int argc, char* argv[]) {
  return run_tool([&]() -> int {
    Options opt("Options");
    opt.flag("help", 'h', "Print help message")
        .value("input", 'i', "Input cloud file (.ply)")
        .value("target", 't', "Input target file (.ply)")
        .value("output", 'o', "Output file (.ply)")
        .value("radius", 'r', "Standard radius to select point neighbours", "0.050000000000000003")
        .value("clean_neighbors", 'c', "N. of neighbors to analyze for each point to clean", "50")
        .value("dev_mult", 'd', "Standard deviation multiplier to clean", "1")
        .flag("negative", 'n', "Saves the not redundant points in a .ply file")
        .flag("all", 'a', "Saves accumulated cloud with all points in a .ply file");
    opt.parse(argc, argv);
    if (opt.count("help")) {
      std::cout << "Accumulate points clouds removing redundant points" << std::endl << std::endl;
      opt.print(std::cout);
      std::cout << std::endl;
      return 0;
    }
    if (!(opt.count("input") && opt.count("target") && opt.count("output")))
      throw std::logic_error("Correct mode of use: " + std::string(argv[0]) +
                             " -i input.ply -t target.ply -o output.ply [opts]");
    const unsigned k = opt.as_uint("clean_neighbors");
    const double mult = opt.as<double>("dev_mult"), radius = opt.as<double>("radius");
    const bool negative = opt.count("negative"), with_redundant = opt.count("all");
    std::string out_name = opt.str("output");

    Cloud src, tgt;
    if (load_ply(opt.str("input"), src) == -1)
      throw std::runtime_error("Couldn't load input point cloud: " + opt.str("input"));
    std::cout << "Loaded " << src.size() << " data points from " << opt.str("input") << std::endl;
    if (load_ply(opt.str("target"), tgt) == -1)
      throw std::runtime_error("Couldn't load target point cloud: " + opt.str("target"));
    std::cout << "Loaded " << tgt.size() << " data points from " << opt.str("target") << std::endl;

    std::cout << "Cloud before accumulate: " << std::endl;
    print_cloud_summary(std::cout, tgt);
    std::cout << std::endl;

    if (!with_redundant) {
      // boost::progress_display's banner + bar (the reference prints one tick per target point)
      std::cout << "\n0%   10   20   30   40   50   60   70   80   90   100%\n"
                   "|----|----|----|----|----|----|----|----|----|----|\n"
                << std::string(51, '*') << std::endl;
      Ctx ctx;
      auto select = [](const Cloud& in, const std::vector<int32_t>& idx, int64_t count) {
        Cloud out;
        out.points.reserve((size_t)count);
        for (int64_t i = 0; i < count; ++i) out.points.push_back(in.points[(size_t)idx[(size_t)i]]);
        out.width = (uint32_t)out.points.size();
        out.height = 1;
        out.is_dense = true;
        return out;
      };
      std::vector<int32_t> kept(src.size() + 1);
      int64_t count = 0;
      if (src.size() > 0) {
        const lc3d_cloud s = as_lc3d(src, false), t = as_lc3d(tgt, false);
        ctx.check(lc3d_box_dedup(ctx.h, &s, &t, radius, kept.data(), &count));
      }
      src = select(src, kept, count);
      count = 0;
      if (src.size() > 0) {
        const lc3d_cloud s = as_lc3d(src, false);
        ctx.check(lc3d_sor(ctx.h, &s, (int32_t)k, mult, 0, kept.data(), &count, nullptr, nullptr));
      }
      src = select(src, kept, count);
    }

    Cloud acc = tgt;
    acc.points.insert(acc.points.end(), src.points.begin(), src.points.end());
    acc.width = (uint32_t)acc.points.size();
    acc.height = 1;
    acc.is_dense = tgt.is_dense && src.is_dense;
    std::cout << "Cloud after accumulate: " << std::endl;
    print_cloud_summary(std::cout, acc);
    std::cout << std::endl;
    if (save_ply_binary(out_name, acc) != 0) throw std::runtime_error("Couldn't write " + out_name);
    if (negative) {
      const size_t pos = out_name.rfind(".ply");
      if (pos != std::string::npos) out_name.erase(pos, 4);
      if (save_ply_binary(out_name + "_negative.ply", src) != 0)
        throw std::runtime_error("Couldn't write " + out_name + "_negative.ply");
    }
    return 0;
  });
}
