// cli_common.hpp — what the four tools share: a boost::program_options-compatible argv
// parser (long/short names, --name=value, unambiguous long prefixes, defaults; the reference
// never calls po::notify, so "required" is only enforced by the tools' own checks —
// SURVEY.md Appendix C), Eigen-style matrix printing and PCL's PointCloud operator<<.
#pragma once
#include <cstdint>
#include <cstdio>
#include <iomanip>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/lc3d.h"
#include "ply_io.hpp"

namespace lc3d_tools {

// Thrown for command-line syntax errors (the role of boost::program_options::error).
struct OptionError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

class Options {
 public:
  struct Spec {
    std::string long_name;
    char short_name;
    bool takes_value;
    std::string default_text;  // as shown by --help, empty = none
    std::string help;
  };
  explicit Options(std::string caption) : caption_(std::move(caption)) {}
  Options& flag(const std::string& l, char s, const std::string& help) {
    specs_.push_back({l, s, false, "", help});
    return *this;
  }
  Options& value(const std::string& l, char s, const std::string& help, const std::string& def = "") {
    specs_.push_back({l, s, true, def, help});
    if (!def.empty()) values_[l] = def;
    return *this;
  }
  void parse(int argc, char** argv) {
    for (int i = 1; i < argc; ++i) {
      std::string a = argv[i];
      const Spec* sp = nullptr;
      std::string val;
      bool have_val = false;
      if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
        std::string name = a.substr(2);
        size_t eq = name.find('=');
        if (eq != std::string::npos) {
          val = name.substr(eq + 1);
          name = name.substr(0, eq);
          have_val = true;
        }
        sp = find_long(name, a);
      } else if (a.size() >= 2 && a[0] == '-' && a[1] != '-') {
        sp = find_short(a[1], a);
        if (a.size() > 2) {
          if (!sp->takes_value) throw OptionError("option '--" + sp->long_name + "' does not take any arguments");
          val = a.substr(2);
          have_val = true;
        }
      } else {
        throw OptionError("too many positional options have been specified on the command line");
      }
      if (sp->takes_value) {
        if (!have_val) {
          if (i + 1 >= argc) throw OptionError("the required argument for option '--" + sp->long_name + "' is missing");
          val = argv[++i];
        }
        if (seen_.count(sp->long_name))
          throw OptionError("option '--" + sp->long_name + "' cannot be specified more than once");
        values_[sp->long_name] = val;
      } else if (have_val) {
        throw OptionError("option '--" + sp->long_name + "' does not take any arguments");
      }
      seen_[sp->long_name] = true;
    }
  }
  // vm.count(name): 1 if given on the command line or defaulted
  int count(const std::string& n) const { return (seen_.count(n) || values_.count(n)) ? 1 : 0; }
  const std::string& str(const std::string& n) const { return values_.at(n); }
  template <typename T>
  T as(const std::string& n) const {
    std::istringstream is(values_.at(n));
    T v;
    is >> v;
    if (is.fail() || !(is >> std::ws).eof())
      throw OptionError("the argument ('" + values_.at(n) + "') for option '--" + n + "' is invalid");
    return v;
  }
  unsigned as_uint(const std::string& n) const {
    const std::string& s = values_.at(n);
    if (!s.empty() && s[0] == '-') throw OptionError("the argument ('" + s + "') for option '--" + n + "' is invalid");
    return as<unsigned>(n);
  }
  void print(std::ostream& os) const {
    os << caption_ << ":\n";
    for (const Spec& s : specs_) {
      std::ostringstream l;
      l << "  ";
      if (s.short_name) l << "-" << s.short_name << " [ --" << s.long_name << " ]";
      else l << "--" << s.long_name;
      if (s.takes_value) {
        l << " arg";
        if (!s.default_text.empty()) l << " (=" << s.default_text << ")";
      }
      std::string left = l.str();
      if (left.size() < 38) left.resize(38, ' ');
      else left += ' ';
      os << left << s.help << "\n";
    }
  }

 private:
  const Spec* find_long(const std::string& name, const std::string& shown) const {
    const Spec* hit = nullptr;
    int n = 0;
    for (const Spec& s : specs_) {
      if (s.long_name == name) return &s;
      if (s.long_name.compare(0, name.size(), name) == 0) {
        hit = &s;
        ++n;
      }
    }
    if (n == 1) return hit;  // boost's default style guesses unambiguous prefixes
    if (n > 1) throw OptionError("option '" + shown + "' is ambiguous");
    throw OptionError("unrecognised option '" + shown + "'");
  }
  const Spec* find_short(char c, const std::string& shown) const {
    for (const Spec& s : specs_)
      if (s.short_name == c) return &s;
    throw OptionError("unrecognised option '" + shown + "'");
  }
  std::string caption_;
  std::vector<Spec> specs_;
  std::map<std::string, std::string> values_;
  std::map<std::string, bool> seen_;
};

// std::cout << Eigen::Matrix4f : default IOFormat (stream precision, columns right-aligned to
// the widest coefficient, single space between columns).
inline void print_matrix4(std::ostream& os, const float* m /* row-major */) {
  std::string cell[16];
  size_t width = 0;
  for (int i = 0; i < 16; ++i) {
    std::ostringstream s;
    s.precision(os.precision());
    s << m[i];
    cell[i] = s.str();
    width = std::max(width, cell[i].size());
  }
  for (int r = 0; r < 4; ++r) {
    for (int c = 0; c < 4; ++c) {
      if (c) os << " ";
      os << std::setw((int)width) << cell[r * 4 + c];
    }
    if (r < 3) os << "\n";
  }
}

// std::cout << pcl::PointCloud<PointT>
inline void print_cloud_summary(std::ostream& os, const Cloud& c) {
  os << "header: seq: 0 stamp: 0 frame_id: " << std::endl << std::endl;
  os << "points[]: " << c.points.size() << std::endl;
  os << "width: " << c.width << std::endl;
  os << "height: " << c.height << std::endl;
  os << "is_dense: " << c.is_dense << std::endl;
  os << "sensor origin (xyz): [0, 0, 0] / orientation (xyzw): [0, 0, 0, 1]" << std::endl;
}

// lc3d_cloud over the 48-byte points (PCL's own layout: no repacking on the host).
inline lc3d_cloud as_lc3d(const Cloud& c, bool normals = true) {
  lc3d_cloud v{};
  v.n = (int64_t)c.points.size();
  if (v.n == 0) return v;
  const Point* p = c.points.data();
  v.xyz = &p->x;
  v.xyz_stride = sizeof(Point);
  if (normals) {
    v.normal = &p->nx;
    v.normal_stride = sizeof(Point);
  }
  v.rgba = &p->rgba;
  v.rgba_stride = sizeof(Point);
  v.curvature = &p->curvature;
  v.curvature_stride = sizeof(Point);
  return v;
}

// main() scaffold shared by the tools.  Like the reference mains: whatever goes wrong is one
// line on stderr and exit status -1; option-parser errors get the "ERROR: " prefix where the
// reference tool catches them separately (fine_registration).
template <typename Body>
int run_tool(Body&& body, bool prefix_option_errors = false) {
  try {
    return body();
  } catch (const OptionError& e) {
    std::cerr << (prefix_option_errors ? "ERROR: " : "") << e.what() << std::endl;
  } catch (const std::exception& e) {
    std::cerr << e.what() << std::endl;
  } catch (...) {
    std::cerr << "An unknown error has occurred." << std::endl;
  }
  return -1;
}

struct Ctx {
  lc3d_ctx* h = nullptr;
  Ctx() {
    int dev = 0;
    if (const char* e = std::getenv("LC3D_DEVICE")) dev = std::atoi(e);
    if (lc3d_create(dev, nullptr, &h) != LC3D_OK)
      throw std::runtime_error(std::string("lc3d: ") + lc3d_last_error(nullptr));
  }
  ~Ctx() { lc3d_destroy(h); }
  void check(int rc) const {
    if (rc != LC3D_OK) throw std::runtime_error(std::string("lc3d: ") + lc3d_last_error(h));
  }
};

}  // namespace lc3d_tools
