// cluster_extraction — drop-in for pcl_tools/cluster_extraction.cpp: keeps the points of every
// Euclidean cluster that holds at least cluster_percentage of the cloud
// (pcl::EuclideanClusterExtraction on the GPU via lc3d_euclidean_clusters).  Output order as the
// reference produces it: cluster by cluster (largest first), ascending point index inside a
// cluster; the outliers file holds the remaining points in input order.
#include "cli_common.hpp"

using namespace lc3d_tools;

int main(int argc, char* argv[]) {
  return run_tool([&]() -> int {
    Options opt("Options");
    opt.flag("help", 'h', "Print help message")
        .value("input", 'i', "Input file (.ply)")
        .value("output", 'o', "Output file (.ply)")
        .flag("outliers_file", 'f', "Saves the removed points in a .ply file")
        .value("cluster_percentage", 'p',
               "Percentage (0 to 1) of points that a cluster needs to contain in order to be considered valid", "0.25")
        .value("tolerance", 't', "Spatial cluster tolerance as a measure in the L2 Euclidean space", "0.02");
    opt.parse(argc, argv);
    if (opt.count("help")) {
      std::cout << "Euclidean cluster extraction." << std::endl << std::endl;
      opt.print(std::cout);
      std::cout << std::endl;
      return 0;
    }
    if (!(opt.count("input") && opt.count("output")))
      throw std::logic_error("Correct mode of use: " + std::string(argv[0]) + " -i input.ply -o output.ply");
    const double cluster_percentage = opt.as<double>("cluster_percentage");
    const double tolerance = opt.as<double>("tolerance");
    if (cluster_percentage < 0 || cluster_percentage > 1)
      throw std::logic_error("cluster_percentage must be a value between 0 and 1");
    std::string out_name = opt.str("output");
    Cloud cloud;
    if (load_ply(opt.str("input"), cloud) == -1)
      throw std::runtime_error("Couldn't load input point cloud: " + opt.str("input"));
    std::cout << "Cloud before filtering: " << std::endl;
    print_cloud_summary(std::cout, cloud);
    std::cout << std::endl;

    const size_t n = cloud.size();
    std::vector<int32_t> labels(n + 1, -1);
    int64_t count = 0;
    if (n > 0) {
      Ctx ctx;
      const lc3d_cloud c = as_lc3d(cloud, false);
      // setMinClusterSize(int) <- cloud->size() * cluster_percentage (double truncated to int)
      const int64_t min_size = (int64_t)(int)((double)n * cluster_percentage);
      ctx.check(lc3d_euclidean_clusters(ctx.h, &c, tolerance, min_size, (int64_t)n, labels.data(), nullptr, 0, &count));
    }
    if (count == 0) throw std::runtime_error("Could not extact clusters for the given dataset");
    std::cout << count << " cluster(s) extracted." << std::endl << std::endl;

    // inliers: clusters in rank order, ascending index inside each (counting sort by label)
    std::vector<size_t> start((size_t)count + 1, 0);
    for (size_t i = 0; i < n; ++i)
      if (labels[i] >= 0) ++start[(size_t)labels[i] + 1];
    for (size_t c = 0; c < (size_t)count; ++c) start[c + 1] += start[c];
    Cloud filtered, removed;
    filtered.points.resize(start[(size_t)count]);
    for (size_t i = 0; i < n; ++i) {
      if (labels[i] >= 0)
        filtered.points[start[(size_t)labels[i]]++] = cloud.points[i];
      else
        removed.points.push_back(cloud.points[i]);
    }
    for (Cloud* c : {&filtered, &removed}) {
      c->width = (uint32_t)c->points.size();
      c->height = 1;
      c->is_dense = true;
    }
    std::cout << "Cloud after filtering: " << std::endl;
    print_cloud_summary(std::cout, filtered);
    std::cout << std::endl;
    if (save_ply_binary(out_name, filtered) != 0) throw std::runtime_error("Couldn't write " + out_name);

    if (opt.count("outliers_file")) {
      const size_t pos = out_name.rfind(".ply");
      if (pos != std::string::npos) out_name.erase(pos, 4);
      if (save_ply_binary(out_name + "_outliers.ply", removed) != 0)
        throw std::runtime_error("Couldn't write " + out_name + "_outliers.ply");
    }
    return 0;
  });
}
