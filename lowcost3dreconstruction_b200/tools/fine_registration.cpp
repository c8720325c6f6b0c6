// fine_registration — drop-in for pcl_tools/fine_registration.cpp (same argv, stdout lines,
// exit codes, PLY outputs), with pcl::IterativeClosestPoint replaced by lc3d_icp_align on the
// GPU.  Opt-in extensions that do not alter the default behaviour: --point_to_plane
// (pcl::IterativeClosestPointWithNormals semantics, target normals from the PLY) and
// --matrix_file F (the 4x4 as a `transform -t`-readable text file).
#include "cli_common.hpp"

using namespace lc3d_tools;

namespace {
struct Usage : std::runtime_error {
  using std::runtime_error::runtime_error;
};
}  // namespace

int main(int argc, char* argv[]) {
  return run_tool([&]() -> int {
    Options opt("Options");
    opt.flag("help", 'h', "Print help message")
        .value("input", 'i', "Input cloud file (.ply)")
        .value("target", 't', "Input target file (.ply)")
        .value("output", 'o', "Output file (.ply)")
        .value("accumulated", 'a', "Save the accumulated point cloud")
        .value("distance_threshold", 0, "The maximum distance threshold between two correspondent points",
               "0.10000000000000001")
        .value("max_iterations", 0, "The maximum number of iterations the internal optimization should run for", "50")
        .value("transformation_epsilon", 0,
               "Maximum allowable difference between two consecutive transformations to be considered as having "
               "converged",
               "1.0000000000000001e-09")
        .value("euclidean_fitness_epsilon", 0,
               "Maximum allowed Euclidean error between two consecutive steps in the ICP loop, before the algorithm "
               "is considered to have converged",
               "0.001")
        .flag("point_to_plane", 0, "[extension] point-to-plane ICP using the target's normals")
        .value("matrix_file", 0, "[extension] also write the 4x4 matrix as a text file (transform -t format)");
    opt.parse(argc, argv);
    if (opt.count("help")) {
      opt.print(std::cout);
      std::cout << std::endl;
      return 0;
    }
    if (!(opt.count("input") && opt.count("target") && opt.count("output")))
      throw Usage{"Correct mode of use: " + std::string(argv[0]) + " -i input.ply -t target.ply -o output.ply [opts]"};
    const std::string src_name = opt.str("input"), tgt_name = opt.str("target"), out_name = opt.str("output");
    const int max_iterations = opt.as<int>("max_iterations");
    if (max_iterations <= 0) throw Usage{"max_iterations needs to be greater than zero."};
    lc3d_icp_params prm{};
    prm.max_correspondence_distance = opt.as<double>("distance_threshold");
    prm.transformation_epsilon = opt.as<double>("transformation_epsilon");
    prm.euclidean_fitness_epsilon = opt.as<double>("euclidean_fitness_epsilon");
    prm.max_iterations = max_iterations;
    prm.mode = opt.count("point_to_plane") ? LC3D_ICP_POINT_TO_PLANE : LC3D_ICP_POINT_TO_POINT;
    prm.compute_fitness = 1;
    prm.dump_iteration = -1;

    Cloud src, tgt;
    if (load_ply(src_name, src) == -1) throw Usage{"Couldn't load input cloud file"};
    std::cout << "Loaded " << src.size() << " data points from " << src_name << std::endl;
    if (load_ply(tgt_name, tgt) == -1) throw Usage{"Couldn't load input target file"};
    std::cout << "Loaded " << tgt.size() << " data points from " << tgt_name << std::endl;

    Ctx ctx;
    const lc3d_cloud s = as_lc3d(src), t = as_lc3d(tgt);
    std::vector<float> rx(3 * src.size() + 3), rn(3 * src.size() + 3);
    lc3d_icp_outputs out{};
    out.registered_xyz = rx.data();
    out.registered_normal = rn.data();
    lc3d_icp_result res{};
    ctx.check(lc3d_icp_align(ctx.h, &s, &t, &prm, &res, &out));

    std::cout << "Has converged: " << (res.converged ? "True" : "False") << std::endl
              << "Score: " << res.fitness << std::endl;
    print_matrix4(std::cout, res.transformation);
    std::cout << std::endl;

    Cloud registered = src;  // colour / curvature carried over, geometry from the device
    for (size_t i = 0; i < registered.size(); ++i) {
      Point& p = registered.points[i];
      p.x = rx[3 * i];
      p.y = rx[3 * i + 1];
      p.z = rx[3 * i + 2];
      p.nx = rn[3 * i];
      p.ny = rn[3 * i + 1];
      p.nz = rn[3 * i + 2];
    }
    if (save_ply_binary(out_name, registered) != 0) throw Usage{"Couldn't write " + out_name};
    if (opt.count("accumulated")) {
      Cloud acc = registered;
      acc.points.insert(acc.points.end(), tgt.points.begin(), tgt.points.end());
      acc.width = (uint32_t)acc.points.size();
      acc.is_dense = registered.is_dense && tgt.is_dense;
      if (save_ply_binary(opt.str("accumulated"), acc) != 0) throw Usage{"Couldn't write " + opt.str("accumulated")};
    }
    if (opt.count("matrix_file")) {
      FILE* f = std::fopen(opt.str("matrix_file").c_str(), "w");
      if (!f) throw Usage{"Couldn't write " + opt.str("matrix_file")};
      for (int r = 0; r < 4; ++r)
        std::fprintf(f, "%.9g %.9g %.9g %.9g\n", res.transformation[4 * r], res.transformation[4 * r + 1],
                     res.transformation[4 * r + 2], res.transformation[4 * r + 3]);
      std::fclose(f);
    }
    return 0;
  }, true);
}
