// chain_registration — fills the "Fine alignment - TODO" of scripts/alignment.sh:123-126 in ONE
// process: loads the turntable views 0.ply .. (N-1).ply, registers view i onto view i-1 with ICP
// for every i (the pairs are independent: ICP is equivariant under a rigid motion applied to
// both clouds), composes G_i = G_{i-1} * T_i in double precision — the same
// cumulative pattern as the coarse chain at alignment.sh:106-113 — applies G_i to view i and
// writes the views back (plus `transform -t`-readable matrix files).  Pairs are split into
// contiguous blocks over the visible GPUs, one thread + one lc3d_ctx per device; only the 4x4
// results meet on the host.  Not part of the reference; built on the same C ABI as the four
// drop-in tools.
//
// In-process view pipeline: with --leaf_size / --neighbors / --normals the
// per-view stages that scripts/alignment.sh:99-100 runs as separate processes with a PLY file
// between each (outlier_removal, normal estimation; VoxelGrid for the turntable chain) run on the
// device right after the load — lc3d_prepare_view — and the pairwise ICP works on the resident
// results (lc3d_icp_align_resident): one PLY read and one PLY write per view, nothing in between.
// The loaded views are page-locked (lc3d_host_register) because each one crosses PCIe more than
// once (as source, as target, and for the final transform).
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

#include "cli_common.hpp"

using namespace lc3d_tools;

namespace {
struct PairResult {
  lc3d_icp_result r{};
  std::string error;
};

void mat_mul(const double* a, const double* b, double* o) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[k * 4 + j];
      o[i * 4 + j] = s;
    }
}
}  // namespace

int main(int argc, char* argv[]) {
  return run_tool([&]() -> int {
    Options opt("Options");
    opt.flag("help", 'h', "Print help message")
        .value("num_captures", 'n', "Number of views: <dir>/0.ply .. <dir>/(n-1).ply")
        .value("dir", 'd', "Directory holding the views", ".")
        .value("output_dir", 'o', "Where to write the registered views (default: in place)")
        .value("distance_threshold", 0, "The maximum distance threshold between two correspondent points",
               "0.10000000000000001")
        .value("max_iterations", 0, "The maximum number of ICP iterations", "50")
        .value("transformation_epsilon", 0, "Transformation epsilon", "1.0000000000000001e-09")
        .value("euclidean_fitness_epsilon", 0, "Euclidean fitness epsilon", "0.001")
        .flag("point_to_plane", 0, "Point-to-plane ICP using the target views' normals")
        .value("leaf_size", 's', "In-process pcl::VoxelGrid leaf size per view (0 = off)", "0")
        .value("neighbors", 0, "In-process StatisticalOutlierRemoval: number of neighbours (0 = off)", "0")
        .value("dev_mult", 0, "In-process StatisticalOutlierRemoval: standard deviation multiplier", "1")
        .value("normals", 0, "In-process normal estimation: number of neighbours (0 = keep the PLY's normals)", "0")
        .value("gpus", 'g', "Number of GPUs to use", "1");
    opt.parse(argc, argv);
    if (opt.count("help")) {
      std::cout << "Pairwise ICP over a turntable view chain, poses composed and applied in one process."
                << std::endl << std::endl;
      opt.print(std::cout);
      std::cout << std::endl;
      return 0;
    }
    if (!opt.count("num_captures"))
      throw std::logic_error("Correct mode of use: " + std::string(argv[0]) + " -n num_captures [-d dir] [opts]");
    const int n = opt.as<int>("num_captures");
    if (n < 2) throw std::logic_error("num_captures needs to be at least 2.");
    const std::string dir = opt.str("dir"), out_dir = opt.count("output_dir") ? opt.str("output_dir") : dir;
    lc3d_icp_params prm{};
    prm.max_correspondence_distance = opt.as<double>("distance_threshold");
    prm.transformation_epsilon = opt.as<double>("transformation_epsilon");
    prm.euclidean_fitness_epsilon = opt.as<double>("euclidean_fitness_epsilon");
    prm.max_iterations = opt.as<int>("max_iterations");
    if (prm.max_iterations <= 0) throw std::logic_error("max_iterations needs to be greater than zero.");
    prm.mode = opt.count("point_to_plane") ? LC3D_ICP_POINT_TO_PLANE : LC3D_ICP_POINT_TO_POINT;
    prm.compute_fitness = 1;
    prm.dump_iteration = -1;
    const int gpus = std::max(1, opt.as<int>("gpus"));
    lc3d_prepare_params prep{};
    prep.leaf_size = opt.as<float>("leaf_size");
    prep.sor_mean_k = opt.as<int>("neighbors");
    prep.sor_stddev_mul = opt.as<double>("dev_mult");
    prep.normals_k = opt.as<int>("normals");
    const bool pipeline = prep.leaf_size > 0.0f || prep.sor_mean_k > 0 || prep.normals_k > 0;
    if (prm.mode == LC3D_ICP_POINT_TO_PLANE && pipeline && prep.normals_k <= 0)
      throw std::logic_error("--point_to_plane after --leaf_size/--neighbors needs --normals (the filtered views "
                             "carry no normals).");

    std::vector<Cloud> views((size_t)n);
    for (int i = 0; i < n; ++i) {
      const std::string f = dir + "/" + std::to_string(i) + ".ply";
      if (load_ply(f, views[(size_t)i]) == -1) throw std::runtime_error("Couldn't load input point cloud: " + f);
      std::cout << "Loaded " << views[(size_t)i].size() << " data points from " << f << std::endl;
    }

    // page-lock the loaded views: each crosses PCIe two or three times (source, target, transform)
    for (auto& v : views)
      if (v.size() > 0) lc3d_host_register(v.points.data(), v.size() * sizeof(Point));
    std::vector<Cloud> filtered(pipeline ? (size_t)n : 0);
    // ---- pairs i -> i-1, contiguous blocks per device -----------------------------------------
    const int n_pairs = n - 1;
    std::vector<PairResult> res((size_t)n_pairs);
    auto worker = [&](int dev, int first, int count) {
      lc3d_ctx* ctx = nullptr;
      if (lc3d_create(dev, nullptr, &ctx) != LC3D_OK) {
        for (int p = first; p < first + count; ++p) res[(size_t)p].error = lc3d_last_error(nullptr);
        return;
      }
      if (!pipeline) {
        for (int p = first; p < first + count; ++p) {  // pair index p registers view p+1 onto view p
          const lc3d_cloud s = as_lc3d(views[(size_t)p + 1]), t = as_lc3d(views[(size_t)p]);
          if (lc3d_icp_align(ctx, &s, &t, &prm, &res[(size_t)p].r, nullptr) != LC3D_OK)
            res[(size_t)p].error = lc3d_last_error(ctx);
        }
      } else {
        // views first .. first+count prepared on the device (each once) by a SECOND context on its own
        // thread, up to two views ahead of the pair being aligned: the per-view passes (short kernels
        // separated by host round trips) overlap the ICP loop of the previous pair on the same GPU.
        // Pairs run on the resident results; the filtered views come back to the host for the final
        // transform + write.
        struct Prepared {
          lc3d_dcloud* d = nullptr;
          int64_t cnt[3] = {0, 0, 0};
          std::string err;
        };
        std::mutex mu;
        std::condition_variable cv;
        std::deque<Prepared> ready;
        bool stop = false;
        lc3d_ctx* pctx = nullptr;
        std::string perr;
        if (lc3d_create(dev, nullptr, &pctx) != LC3D_OK) perr = lc3d_last_error(nullptr);
        std::thread preparer([&] {
          for (int v = first; v <= first + count; ++v) {
            Prepared pr;
            if (!perr.empty()) {
              pr.err = perr;
            } else {
              const lc3d_cloud c = as_lc3d(views[(size_t)v], false);
              if (lc3d_prepare_view(pctx, &c, &prep, &pr.d, pr.cnt) != LC3D_OK) pr.err = lc3d_last_error(pctx);
            }
            std::unique_lock<std::mutex> lock(mu);
            cv.wait(lock, [&] { return ready.size() < 2 || stop; });
            if (stop) {
              lock.unlock();
              if (pr.d) lc3d_cloud_free(pctx, pr.d);
              return;
            }
            const bool failed = !pr.err.empty();
            ready.push_back(pr);
            cv.notify_all();
            if (failed) return;
          }
        });
        lc3d_dcloud* prev = nullptr;
        for (int v = first; v <= first + count; ++v) {
          Prepared pr;
          {
            std::unique_lock<std::mutex> lock(mu);
            cv.wait(lock, [&] { return !ready.empty(); });
            pr = ready.front();
            ready.pop_front();
            cv.notify_all();
          }
          lc3d_dcloud* cur = pr.d;
          std::string err = pr.err;
          if (err.empty() && v > first &&
              lc3d_icp_align_resident(ctx, cur, prev, &prm, &res[(size_t)v - 1].r, nullptr) != LC3D_OK)
            err = lc3d_last_error(ctx);
          // the block's first view is downloaded by its own block only when it is view 0 or this is
          // the block that registers it; shared border views are written by the block that sources them
          if (err.empty() && (v > first || v == 0)) {
            Cloud& dst = filtered[(size_t)v];
            const size_t m = (size_t)pr.cnt[2];
            std::vector<float> xyz(3 * m + 3), nrm(3 * m + 3), curv(m + 1);
            if (m > 0 && lc3d_cloud_download(ctx, cur, xyz.data(), nrm.data(), curv.data()) != LC3D_OK)
              err = lc3d_last_error(ctx);
            dst.points.assign(m, Point{});
            for (size_t k = 0; k < m; ++k) {
              Point& p = dst.points[k];
              p.x = xyz[3 * k]; p.y = xyz[3 * k + 1]; p.z = xyz[3 * k + 2]; p.w = 1.0f;
              if (prep.normals_k > 0) { p.nx = nrm[3 * k]; p.ny = nrm[3 * k + 1]; p.nz = nrm[3 * k + 2]; p.curvature = curv[k]; }
              p.rgba = 0xff808080u;
            }
            dst.width = (uint32_t)m; dst.height = 1; dst.is_dense = true;
          }
          if (!err.empty())
            for (int p = std::max(v - 1, first); p < first + count; ++p) res[(size_t)p].error = err;
          if (prev) lc3d_cloud_free(pctx, prev);  // back to the pool of the context that made it (thread-safe)
          prev = cur;
          if (!err.empty()) break;
        }
        {
          std::lock_guard<std::mutex> lock(mu);
          stop = true;
          cv.notify_all();
        }
        preparer.join();
        if (prev) lc3d_cloud_free(pctx, prev);
        for (auto& pr : ready)
          if (pr.d) lc3d_cloud_free(pctx, pr.d);
        if (pctx) lc3d_destroy(pctx);
      }
      lc3d_destroy(ctx);
    };
    {
      std::vector<std::thread> th;
      const int base = n_pairs / gpus, extra = n_pairs % gpus;
      int first = 0;
      for (int d = 0; d < gpus; ++d) {
        const int count = base + (d < extra ? 1 : 0);
        if (count > 0) th.emplace_back(worker, d, first, count);
        first += count;
      }
      for (auto& t : th) t.join();
    }
    for (int p = 0; p < n_pairs; ++p)
      if (!res[(size_t)p].error.empty()) throw std::runtime_error("lc3d: " + res[(size_t)p].error);

    for (auto& v : views)
      if (v.size() > 0) lc3d_host_unregister(v.points.data());
    if (pipeline) views.swap(filtered);  // what gets transformed and written are the filtered views
    // ---- compose, apply, write ---------------------------------------------------------------
    Ctx ctx;
    double G[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    if (out_dir != dir && save_ply_binary(out_dir + "/0.ply", views[0]) != 0)
      throw std::runtime_error("Couldn't write " + out_dir + "/0.ply");
    for (int i = 1; i < n; ++i) {
      const lc3d_icp_result& r = res[(size_t)i - 1].r;
      std::cout << "Pair " << i << " -> " << i - 1 << "\nHas converged: " << (r.converged ? "True" : "False")
                << "\nScore: " << r.fitness << "\nIterations: " << r.iterations << std::endl;
      print_matrix4(std::cout, r.transformation);
      std::cout << std::endl;
      double T[16], Gn[16];
      for (int k = 0; k < 16; ++k) T[k] = r.transformation[k];
      mat_mul(G, T, Gn);
      std::memcpy(G, Gn, sizeof G);
      float Gf[16];
      for (int k = 0; k < 16; ++k) Gf[k] = (float)G[k];
      Cloud& v = views[(size_t)i];
      const lc3d_cloud c = as_lc3d(v);
      std::vector<float> xyz(3 * v.size() + 3), nrm(3 * v.size() + 3);
      if (v.size() > 0) ctx.check(lc3d_transform(ctx.h, &c, Gf, xyz.data(), nrm.data()));
      for (size_t k = 0; k < v.size(); ++k) {
        Point& p = v.points[k];
        p.x = xyz[3 * k];
        p.y = xyz[3 * k + 1];
        p.z = xyz[3 * k + 2];
        p.nx = nrm[3 * k];
        p.ny = nrm[3 * k + 1];
        p.nz = nrm[3 * k + 2];
      }
      const std::string f = out_dir + "/" + std::to_string(i) + ".ply";
      if (save_ply_binary(f, v) != 0) throw std::runtime_error("Couldn't write " + f);
      FILE* m = std::fopen((out_dir + "/fine_" + std::to_string(i) + ".txt").c_str(), "w");
      if (m) {
        for (int rr = 0; rr < 4; ++rr)
          std::fprintf(m, "%.9g %.9g %.9g %.9g\n", G[4 * rr], G[4 * rr + 1], G[4 * rr + 2], G[4 * rr + 3]);
        std::fclose(m);
      }
    }
    return 0;
  }, true);
}
