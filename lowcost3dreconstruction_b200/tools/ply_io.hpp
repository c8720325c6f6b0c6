// ply_io.hpp — PLY reader / PCL-compatible binary writer for the CLI tools.
//
// Replaces pcl::io::loadPLYFile / savePLYFileBinary (every tool, e.g.
// pcl_tools/fine_registration.cpp:94-102,129).  Reads ASCII and binary_little_endian PLY with
// arbitrary property order / types and extra elements (capture-tool ASCII files,
// capture/depth_capture/depth_capture.cpp:281-308; MeshLab output with alpha and empty face
// elements; PCL binary output).  Writes the layout PCL's writer produces for
// PointXYZRGBNormal, camera element included.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

namespace lc3d_tools {

// In-memory point, same 48-byte layout as pcl::PointXYZRGBNormal.
struct alignas(16) Point {
  float x, y, z, w;
  float nx, ny, nz, nw;
  uint32_t rgba;
  float curvature;
  float pad[2];
};
static_assert(sizeof(Point) == 48, "PointXYZRGBNormal layout");

struct Cloud {
  std::vector<Point> points;
  uint32_t width = 0, height = 1;
  bool is_dense = true;
  bool has_normals = false, has_color = false, has_curvature = false;
  size_t size() const { return points.size(); }
};

namespace detail {

enum Type { I8, U8, I16, U16, I32, U32, F32, F64, BAD };
inline Type parse_type(const std::string& t) {
  if (t == "char" || t == "int8") return I8;
  if (t == "uchar" || t == "uint8") return U8;
  if (t == "short" || t == "int16") return I16;
  if (t == "ushort" || t == "uint16") return U16;
  if (t == "int" || t == "int32") return I32;
  if (t == "uint" || t == "uint32") return U32;
  if (t == "float" || t == "float32") return F32;
  if (t == "double" || t == "float64") return F64;
  return BAD;
}
inline int type_size(Type t) {
  switch (t) {
    case I8: case U8: return 1;
    case I16: case U16: return 2;
    case I32: case U32: case F32: return 4;
    case F64: return 8;
    default: return 0;
  }
}
inline double read_bin(const unsigned char* p, Type t) {
  switch (t) {
    case I8: return (double)*reinterpret_cast<const int8_t*>(p);
    case U8: return (double)*p;
    case I16: { int16_t v; std::memcpy(&v, p, 2); return v; }
    case U16: { uint16_t v; std::memcpy(&v, p, 2); return v; }
    case I32: { int32_t v; std::memcpy(&v, p, 4); return v; }
    case U32: { uint32_t v; std::memcpy(&v, p, 4); return v; }
    case F32: { float v; std::memcpy(&v, p, 4); return v; }
    case F64: { double v; std::memcpy(&v, p, 8); return v; }
    default: return 0;
  }
}

// Fast decimal -> double for the ASCII body: exact (correctly rounded) on the Clinger fast path
// (<= 15 significant digits, |exp10| <= 22: mantissa and power of ten are exact doubles, one
// rounded multiply/divide), strtod for everything else (long mantissas, nan/inf, hex...).
// Returns the position after the token, or nullptr at end of input / on a malformed token.
inline const char* parse_number(const char* p, const char* end, double* out) {
  while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p;
  if (p >= end) return nullptr;
  const char* start = p;
  bool neg = false;
  if (*p == '-' || *p == '+') {
    neg = *p == '-';
    ++p;
  }
  uint64_t mant = 0;
  int digits = 0, frac = 0;
  bool any = false;
  while (p < end && *p >= '0' && *p <= '9') {
    if (digits < 18) {
      mant = mant * 10 + (uint64_t)(*p - '0');
      if (mant) ++digits;
    } else {
      --frac;  // dropped digit: fall back below
      digits = 99;
    }
    any = true;
    ++p;
  }
  if (p < end && *p == '.') {
    ++p;
    while (p < end && *p >= '0' && *p <= '9') {
      if (digits < 18) {
        mant = mant * 10 + (uint64_t)(*p - '0');
        if (mant) ++digits;
        ++frac;
      } else {
        digits = 99;
      }
      any = true;
      ++p;
    }
  }
  int exp10 = 0;
  bool fast = any && digits <= 15;
  if (any && p < end && (*p == 'e' || *p == 'E')) {
    const char* q = p + 1;
    bool eneg = false;
    if (q < end && (*q == '-' || *q == '+')) {
      eneg = *q == '-';
      ++q;
    }
    if (q < end && *q >= '0' && *q <= '9') {
      int e = 0;
      while (q < end && *q >= '0' && *q <= '9') {
        if (e < 10000) e = e * 10 + (*q - '0');
        ++q;
      }
      exp10 = eneg ? -e : e;
      p = q;
    }
  }
  const bool token_end = p >= end || *p == ' ' || *p == '\t' || *p == '\n' || *p == '\r';
  static const double kPow10[] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                  1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
  const int e = exp10 - frac;
  if (fast && token_end && e >= -22 && e <= 22) {
    double v = (double)mant;
    v = e < 0 ? v / kPow10[-e] : v * kPow10[e];
    *out = neg ? -v : v;
    return p;
  }
  // slow path: let strtod decide (also handles nan / inf / malformed tokens)
  const char* tok_end = start;
  while (tok_end < end && !(*tok_end == ' ' || *tok_end == '\t' || *tok_end == '\n' || *tok_end == '\r')) ++tok_end;
  char buf[128];
  size_t len = (size_t)(tok_end - start);
  if (len == 0) return nullptr;
  if (len >= sizeof buf) len = sizeof buf - 1;
  std::memcpy(buf, start, len);
  buf[len] = 0;
  char* ep = nullptr;
  double v = std::strtod(buf, &ep);
  if (ep == buf) v = std::nan("");  // non-numeric token
  *out = v;
  return tok_end;
}

struct Property {
  std::string name;
  Type type = BAD;
  bool is_list = false;
  Type count_type = BAD;
};
struct Element {
  std::string name;
  size_t count = 0;
  std::vector<Property> props;
};

// which Point field a vertex property feeds
enum Slot { S_NONE, S_X, S_Y, S_Z, S_NX, S_NY, S_NZ, S_R, S_G, S_B, S_A, S_RGB, S_CURV };
inline Slot slot_of(const std::string& n) {
  if (n == "x") return S_X;
  if (n == "y") return S_Y;
  if (n == "z") return S_Z;
  if (n == "nx" || n == "normal_x") return S_NX;
  if (n == "ny" || n == "normal_y") return S_NY;
  if (n == "nz" || n == "normal_z") return S_NZ;
  if (n == "red" || n == "r" || n == "diffuse_red") return S_R;
  if (n == "green" || n == "g" || n == "diffuse_green") return S_G;
  if (n == "blue" || n == "b" || n == "diffuse_blue") return S_B;
  if (n == "alpha") return S_A;
  if (n == "rgb" || n == "rgba") return S_RGB;
  if (n == "curvature") return S_CURV;
  return S_NONE;
}
inline void assign(Point& p, Slot s, double v, Type t, const unsigned char* raw) {
  // NaN colour channels become 0 (a NaN -> unsigned conversion would be undefined behaviour)
  auto byte = [](double d) { return (uint32_t)(!(d > 0) ? 0 : d > 255 ? 255 : d); };
  switch (s) {
    case S_X: p.x = (float)v; break;
    case S_Y: p.y = (float)v; break;
    case S_Z: p.z = (float)v; break;
    case S_NX: p.nx = (float)v; break;
    case S_NY: p.ny = (float)v; break;
    case S_NZ: p.nz = (float)v; break;
    case S_R: p.rgba = (p.rgba & 0xff00ffffu) | (byte(v) << 16); break;
    case S_G: p.rgba = (p.rgba & 0xffff00ffu) | (byte(v) << 8); break;
    case S_B: p.rgba = (p.rgba & 0xffffff00u) | byte(v); break;
    case S_A: p.rgba = (p.rgba & 0x00ffffffu) | (byte(v) << 24); break;
    case S_RGB:
      if (raw && type_size(t) == 4) std::memcpy(&p.rgba, raw, 4);  // packed rgb(a) bits
      else p.rgba = (uint32_t)v;
      break;
    case S_CURV: p.curvature = (float)v; break;
    default: break;
  }
}

}  // namespace detail

// Returns 0 on success, -1 on failure (like pcl::io::loadPLYFile).
inline int load_ply(const std::string& path, Cloud& cloud, std::string* err = nullptr) {
  using namespace detail;
  auto fail = [&](const std::string& m) {
    if (err) *err = m;
    return -1;
  };
  std::ifstream in(path, std::ios::binary);
  if (!in) return fail("cannot open " + path);
  std::string line;
  if (!std::getline(in, line) || line.substr(0, 3) != "ply") return fail("not a PLY file");
  int format = -1;  // 0 ascii, 1 binary LE
  std::vector<Element> elems;
  bool header_done = false;
  while (std::getline(in, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    std::istringstream ls(line);
    std::string tok;
    if (!(ls >> tok)) continue;
    if (tok == "format") {
      std::string f;
      ls >> f;
      if (f == "ascii") format = 0;
      else if (f == "binary_little_endian") format = 1;
      else return fail("unsupported PLY format " + f);
    } else if (tok == "element") {
      Element e;
      ls >> e.name >> e.count;
      elems.push_back(e);
    } else if (tok == "property") {
      if (elems.empty()) return fail("property before element");
      Property p;
      std::string t;
      ls >> t;
      if (t == "list") {
        std::string ct, it;
        ls >> ct >> it >> p.name;
        p.is_list = true;
        p.count_type = parse_type(ct);
        p.type = parse_type(it);
      } else {
        p.type = parse_type(t);
        ls >> p.name;
      }
      if (p.type == BAD || (p.is_list && p.count_type == BAD)) return fail("bad property type in header");
      elems.back().props.push_back(p);
    } else if (tok == "end_header") {
      header_done = true;
      break;
    }  // comment / obj_info: ignored
  }
  if (!header_done || format < 0) return fail("truncated PLY header");
  cloud = Cloud();
  std::vector<char> ascii;
  const char* cur = nullptr;
  const char* ascii_end = nullptr;
  bool ascii_loaded = false;
  for (const Element& e : elems) {
    const bool is_vertex = e.name == "vertex";
    std::vector<Slot> slots;
    for (const Property& p : e.props) {
      Slot s = is_vertex && !p.is_list ? slot_of(p.name) : S_NONE;
      slots.push_back(s);
      if (s == S_NX) cloud.has_normals = true;
      if (s == S_R || s == S_RGB) cloud.has_color = true;
      if (s == S_CURV) cloud.has_curvature = true;
    }
    if (is_vertex) cloud.points.reserve(e.count);
    if (format == 0) {
      if (!ascii_loaded) {  // slurp the rest of the file once; tokens are parsed in memory
        const std::streampos here = in.tellg();
        in.seekg(0, std::ios::end);
        const std::streampos fin = in.tellg();
        in.seekg(here);
        ascii.resize((size_t)(fin - here));
        in.read(ascii.data(), (std::streamsize)ascii.size());
        cur = ascii.data();
        ascii_end = ascii.data() + ascii.size();
        ascii_loaded = true;
      }
      // Large vertex elements without list properties are parsed by several threads: the token
      // starts of each byte chunk are counted first, so every thread knows which vertex its
      // chunk begins in and parses a contiguous, vertex-aligned range straight into the output.
      bool has_list = false;
      for (const Property& p : e.props) has_list = has_list || p.is_list;
      const size_t np = e.props.size();
      unsigned nthreads = std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
      if (const char* env = std::getenv("LC3D_PLY_THREADS")) nthreads = (unsigned)std::max(1, std::atoi(env));
      if (is_vertex && !has_list && np > 0 && nthreads > 1 && e.count >= (size_t)100000) {
        const size_t bytes = (size_t)(ascii_end - cur);
        const size_t chunk = (bytes + nthreads - 1) / nthreads;
        auto is_ws = [](char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; };
        std::vector<size_t> tok_before(nthreads + 1, 0);
        {
          std::vector<std::thread> th;
          for (unsigned k = 0; k < nthreads; ++k)
            th.emplace_back([&, k] {
              const char* b = cur + std::min(bytes, (size_t)k * chunk);
              const char* f = cur + std::min(bytes, (size_t)(k + 1) * chunk);
              size_t n = 0;
              bool prev_ws = b == cur ? true : is_ws(b[-1]);
              for (const char* q = b; q < f; ++q) {
                const bool ws = is_ws(*q);
                n += (!ws && prev_ws) ? 1 : 0;
                prev_ws = ws;
              }
              tok_before[k + 1] = n;
            });
          for (auto& t : th) t.join();
        }
        for (unsigned k = 0; k < nthreads; ++k) tok_before[k + 1] += tok_before[k];
        if (tok_before[nthreads] < e.count * np) return fail("truncated PLY data");
        cloud.points.resize(e.count);
        std::vector<const char*> end_ptr(nthreads, nullptr);
        std::vector<char> bad(nthreads, 0);
        {
          std::vector<std::thread> th;
          for (unsigned k = 0; k < nthreads; ++k)
            th.emplace_back([&, k] {
              const size_t v0 = std::min(e.count, (tok_before[k] + np - 1) / np);
              const size_t v1 = k + 1 == nthreads ? e.count : std::min(e.count, (tok_before[k + 1] + np - 1) / np);
              if (v0 >= v1) return;
              const char* q = cur + std::min(bytes, (size_t)k * chunk);
              // a token straddling the chunk start belongs to the previous chunk
              if (q != cur && !is_ws(q[-1]))
                while (q < ascii_end && !is_ws(*q)) ++q;
              double v;
              for (size_t skip = v0 * np - tok_before[k]; skip > 0; --skip)
                if (!(q = parse_number(q, ascii_end, &v))) {
                  bad[k] = 1;
                  return;
                }
              for (size_t i = v0; i < v1; ++i) {
                Point pt{};
                pt.w = 1.0f;
                pt.rgba = 0xff000000u;
                for (size_t j = 0; j < np; ++j) {
                  if (!(q = parse_number(q, ascii_end, &v))) {
                    bad[k] = 1;
                    return;
                  }
                  assign(pt, slots[j], v, e.props[j].type, nullptr);
                }
                cloud.points[i] = pt;
              }
              end_ptr[k] = q;
            });
          for (auto& t : th) t.join();
        }
        for (unsigned k = 0; k < nthreads; ++k) {
          if (bad[k]) return fail("truncated PLY data");
          if (end_ptr[k]) cur = end_ptr[k];  // ranges are ascending: the last one ends the element
        }
        continue;
      }
      for (size_t i = 0; i < e.count; ++i) {
        Point pt{};
        pt.w = 1.0f;
        pt.rgba = 0xff000000u;
        for (size_t k = 0; k < e.props.size(); ++k) {
          const Property& p = e.props[k];
          double v;
          if (p.is_list) {
            if (!(cur = parse_number(cur, ascii_end, &v))) return fail("truncated PLY data");
            const long cnt = (long)v;
            for (long c = 0; c < cnt; ++c)
              if (!(cur = parse_number(cur, ascii_end, &v))) return fail("truncated PLY data");
          } else {
            if (!(cur = parse_number(cur, ascii_end, &v))) return fail("truncated PLY data");
            if (is_vertex) assign(pt, slots[k], v, p.type, nullptr);
          }
        }
        if (is_vertex) cloud.points.push_back(pt);
      }
    } else {
      bool fixed = true;
      size_t rec = 0;
      for (const Property& p : e.props) {
        if (p.is_list) fixed = false;
        rec += type_size(p.type);
      }
      if (fixed) {
        std::vector<unsigned char> buf(rec * std::min<size_t>(e.count, 1 << 16));
        size_t done = 0;
        while (done < e.count) {
          size_t n = std::min<size_t>(e.count - done, 1 << 16);
          in.read(reinterpret_cast<char*>(buf.data()), (std::streamsize)(rec * n));
          if ((size_t)in.gcount() != rec * n) return fail("truncated PLY data");
          if (is_vertex) {
            for (size_t i = 0; i < n; ++i) {
              const unsigned char* r = buf.data() + i * rec;
              Point pt{};
              pt.w = 1.0f;
              pt.rgba = 0xff000000u;
              size_t off = 0;
              for (size_t k = 0; k < e.props.size(); ++k) {
                const Property& p = e.props[k];
                if (slots[k] != S_NONE) assign(pt, slots[k], read_bin(r + off, p.type), p.type, r + off);
                off += type_size(p.type);
              }
              cloud.points.push_back(pt);
            }
          }
          done += n;
        }
      } else {
        // an element with a list property has no fixed record size: walk it property by property.
        // A VERTEX element of this kind still contributes its scalar properties (as the ASCII
        // path does); list entries are skipped.
        unsigned char tmp[8];
        for (size_t i = 0; i < e.count; ++i) {
          Point pt{};
          pt.w = 1.0f;
          pt.rgba = 0xff000000u;
          for (size_t k = 0; k < e.props.size(); ++k) {
            const Property& p = e.props[k];
            if (p.is_list) {
              in.read(reinterpret_cast<char*>(tmp), type_size(p.count_type));
              if (!in) return fail("truncated PLY data");
              long cnt = (long)read_bin(tmp, p.count_type);
              if (cnt < 0) return fail("negative list length in PLY data");
              in.ignore((std::streamsize)cnt * type_size(p.type));
            } else if (is_vertex && slots[k] != S_NONE) {
              in.read(reinterpret_cast<char*>(tmp), type_size(p.type));
              if (!in) return fail("truncated PLY data");
              assign(pt, slots[k], read_bin(tmp, p.type), p.type, tmp);
            } else {
              in.ignore(type_size(p.type));
            }
            if (!in) return fail("truncated PLY data");
          }
          if (is_vertex) cloud.points.push_back(pt);
        }
      }
    }
  }
  cloud.width = (uint32_t)cloud.points.size();
  cloud.height = 1;
  cloud.is_dense = true;
  for (const Point& p : cloud.points)
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) {
      cloud.is_dense = false;
      break;
    }
  return 0;
}

// pcl::io::savePLYFileBinary layout for PointXYZRGBNormal.
inline int save_ply_binary(const std::string& path, const Cloud& cloud) {
  // write to a temporary and rename, so that `-i F -o F` (scripts/alignment.sh:99) stays safe
  const std::string tmp = path + ".tmp~";
  FILE* f = std::fopen(tmp.c_str(), "wb");
  if (!f) return -1;
  const size_t n = cloud.points.size();
  std::fprintf(f,
               "ply\nformat binary_little_endian 1.0\ncomment PCL generated\nelement vertex %zu\n"
               "property float x\nproperty float y\nproperty float z\n"
               "property uchar red\nproperty uchar green\nproperty uchar blue\n"
               "property float nx\nproperty float ny\nproperty float nz\nproperty float curvature\n"
               "element camera 1\n"
               "property float view_px\nproperty float view_py\nproperty float view_pz\n"
               "property float x_axisx\nproperty float x_axisy\nproperty float x_axisz\n"
               "property float y_axisx\nproperty float y_axisy\nproperty float y_axisz\n"
               "property float z_axisx\nproperty float z_axisy\nproperty float z_axisz\n"
               "property float focal\nproperty float scalex\nproperty float scaley\n"
               "property float centerx\nproperty float centery\n"
               "property int viewportx\nproperty int viewporty\n"
               "property float k1\nproperty float k2\nend_header\n",
               n);
  std::vector<unsigned char> buf;
  buf.reserve(31 * std::min<size_t>(n, 1 << 16));
  for (size_t i = 0; i < n; ++i) {
    const Point& p = cloud.points[i];
    unsigned char r[31];
    std::memcpy(r + 0, &p.x, 12);
    r[12] = (unsigned char)((p.rgba >> 16) & 0xff);
    r[13] = (unsigned char)((p.rgba >> 8) & 0xff);
    r[14] = (unsigned char)(p.rgba & 0xff);
    std::memcpy(r + 15, &p.nx, 12);
    std::memcpy(r + 27, &p.curvature, 4);
    buf.insert(buf.end(), r, r + 31);
    if (buf.size() >= 31u * (1 << 16) || i + 1 == n) {
      if (std::fwrite(buf.data(), 1, buf.size(), f) != buf.size()) {
        std::fclose(f);
        return -1;
      }
      buf.clear();
    }
  }
  const float w = (float)cloud.width, h = (float)cloud.height;
  float cam[19] = {0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, w / 2.0f, h / 2.0f, 0, 0};
  int32_t vp[2] = {(int32_t)cloud.width, (int32_t)cloud.height};
  std::fwrite(cam, 4, 17, f);
  std::fwrite(vp, 4, 2, f);
  std::fwrite(cam + 17, 4, 2, f);
  if (std::fclose(f) != 0) return -1;
  if (std::rename(tmp.c_str(), path.c_str()) != 0) return -1;
  return 0;
}

}  // namespace lc3d_tools
