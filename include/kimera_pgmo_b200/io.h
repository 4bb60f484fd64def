// Disk formats on either side of the mesh path, host side, C++ (SURVEY 8f "next" row 4):
//   IOData::load / save                 kimera_pgmo/src/utils/mesh_io.cpp:14-101 (PLY through happly)
//   ReadMeshWithStampsFromPly / Write.. kimera_pgmo/src/utils/mesh_io.cpp:103-149
//   loadGraphFile                       the share of DeformationGraph::load the mesh path needs
//                                       (kimera_pgmo/src/deformation_graph_io.cpp:194-349: NODE and
//                                       VERTEX lines; factors stay with the solver upstream)
// so that a recorded session (mesh .ply + graph .dgrf) can drive the shim and scripts/bench_shim.cpp
// without Python.  The PLY reader is written from the format description, not from happly: ASCII
// and binary_little_endian bodies, scalar properties of any PLY type, one list property per element.
#pragma once

#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "kimera_pgmo_b200/deformation_graph.h"

namespace kimera_pgmo_b200 {

struct IOData {
  using Ptr = std::shared_ptr<IOData>;
  std::vector<float> x, y, z;
  std::vector<uint8_t> r, g, b, a;
  std::vector<Timestamp> stamps;
  std::vector<uint32_t> labels;
  std::vector<std::vector<uint32_t>> faces;

  static Ptr load(const std::string& filename);
  void save(const std::string& filename) const;
};

namespace detail {

struct PlyProperty {
  std::string name;
  int type = 0;       // index into kPlyTypes
  bool is_list = false;
  int count_type = 0;
};
struct PlyElement {
  std::string name;
  size_t count = 0;
  std::vector<PlyProperty> props;
};
struct PlyType {
  const char* a;
  const char* b;
  int bytes;
  char kind;  // i, u, f
};
inline const PlyType* plyTypes(int* n) {
  static const PlyType t[] = {{"char", "int8", 1, 'i'},   {"uchar", "uint8", 1, 'u'},
                              {"short", "int16", 2, 'i'}, {"ushort", "uint16", 2, 'u'},
                              {"int", "int32", 4, 'i'},   {"uint", "uint32", 4, 'u'},
                              {"float", "float32", 4, 'f'}, {"double", "float64", 8, 'f'}};
  *n = 8;
  return t;
}
inline int plyTypeIndex(const std::string& s) {
  int n = 0;
  const PlyType* t = plyTypes(&n);
  for (int i = 0; i < n; ++i) {
    if (s == t[i].a || s == t[i].b) return i;
  }
  throw std::runtime_error("PLY: unknown property type '" + s + "'");
}
// one scalar of PLY type `type` from a little-endian byte stream, as double (exact for every
// PLY type but 64-bit doubles that are not needed here)
inline double plyScalar(const unsigned char* p, int type) {
  switch (type) {
    case 0: return static_cast<double>(*reinterpret_cast<const int8_t*>(p));
    case 1: return static_cast<double>(*p);
    case 2: { int16_t v; std::memcpy(&v, p, 2); return v; }
    case 3: { uint16_t v; std::memcpy(&v, p, 2); return v; }
    case 4: { int32_t v; std::memcpy(&v, p, 4); return v; }
    case 5: { uint32_t v; std::memcpy(&v, p, 4); return v; }
    case 6: { float v; std::memcpy(&v, p, 4); return v; }
    default: { double v; std::memcpy(&v, p, 8); return v; }
  }
}

struct PlyColumns {
  // per element name: scalar columns by property name (as double) and list columns
  struct Element {
    std::vector<std::string> names;
    std::vector<std::vector<double>> scalars;            // same order as names, empty for lists
    std::vector<std::vector<std::vector<double>>> lists; // same order as names, empty for scalars
    const std::vector<double>* scalar(const std::string& n) const {
      for (size_t i = 0; i < names.size(); ++i) {
        if (names[i] == n && !scalars[i].empty()) return &scalars[i];
      }
      return nullptr;
    }
    const std::vector<std::vector<double>>* list(const std::string& n) const {
      for (size_t i = 0; i < names.size(); ++i) {
        if (names[i] == n && lists[i].size()) return &lists[i];
      }
      return nullptr;
    }
  };
  std::vector<std::string> element_names;
  std::vector<Element> elements;
  std::vector<size_t> counts;
  const Element* element(const std::string& n) const {
    for (size_t i = 0; i < element_names.size(); ++i) {
      if (element_names[i] == n) return &elements[i];
    }
    return nullptr;
  }
  size_t count(const std::string& n) const {
    for (size_t i = 0; i < element_names.size(); ++i) {
      if (element_names[i] == n) return counts[i];
    }
    return 0;
  }
};

inline PlyColumns readPly(const std::string& filename) {
  std::ifstream in(filename, std::ios::binary);
  if (!in) throw std::runtime_error("PLY: failed to open '" + filename + "'");
  std::string body((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  const size_t eh = body.find("end_header");
  if (eh == std::string::npos || body.compare(0, 3, "ply") != 0) {
    throw std::runtime_error("PLY: '" + filename + "' has no PLY header");
  }
  const size_t data_at = body.find('\n', eh) + 1;
  std::vector<PlyElement> elements;
  std::string format;
  {
    std::istringstream header(body.substr(0, eh));
    std::string line;
    while (std::getline(header, line)) {
      std::istringstream ls(line);
      std::string word;
      if (!(ls >> word)) continue;
      if (word == "format") {
        ls >> format;
      } else if (word == "element") {
        PlyElement e;
        ls >> e.name >> e.count;
        elements.push_back(e);
      } else if (word == "property") {
        if (elements.empty()) throw std::runtime_error("PLY: property before any element");
        PlyProperty p;
        std::string t;
        ls >> t;
        if (t == "list") {
          std::string ct, it;
          ls >> ct >> it >> p.name;
          p.is_list = true;
          p.count_type = plyTypeIndex(ct);
          p.type = plyTypeIndex(it);
        } else {
          p.type = plyTypeIndex(t);
          ls >> p.name;
        }
        elements.back().props.push_back(p);
      }
    }
  }
  int n_types = 0;
  const PlyType* types = plyTypes(&n_types);
  PlyColumns out;
  const bool ascii = format == "ascii";
  if (!ascii && format != "binary_little_endian") {
    throw std::runtime_error("PLY: format '" + format + "' is not supported (ascii, binary_little_endian)");
  }
  const unsigned char* p = reinterpret_cast<const unsigned char*>(body.data()) + data_at;
  const unsigned char* end = reinterpret_cast<const unsigned char*>(body.data()) + body.size();
  const char* text = body.c_str() + data_at;  // NUL-terminated: strtod stops at the end
  auto nextAscii = [&]() {
    char* stop = nullptr;
    const double v = std::strtod(text, &stop);
    if (stop == text) throw std::runtime_error("PLY: '" + filename + "' ends inside an element");
    text = stop;
    return v;
  };
  auto nextBinary = [&](int type) {
    if (p + types[type].bytes > end) throw std::runtime_error("PLY: '" + filename + "' ends inside an element");
    const double v = plyScalar(p, type);
    p += types[type].bytes;
    return v;
  };
  for (const PlyElement& e : elements) {
    PlyColumns::Element cols;
    for (const PlyProperty& pr : e.props) {
      cols.names.push_back(pr.name);
      cols.scalars.emplace_back();
      cols.lists.emplace_back();
      if (pr.is_list) cols.lists.back().reserve(e.count);
      else cols.scalars.back().reserve(e.count);
    }
    for (size_t i = 0; i < e.count; ++i) {
      for (size_t k = 0; k < e.props.size(); ++k) {
        const PlyProperty& pr = e.props[k];
        if (pr.is_list) {
          const size_t n = static_cast<size_t>(ascii ? nextAscii() : nextBinary(pr.count_type));
          std::vector<double> items(n);
          for (size_t j = 0; j < n; ++j) items[j] = ascii ? nextAscii() : nextBinary(pr.type);
          cols.lists[k].push_back(std::move(items));
        } else {
          cols.scalars[k].push_back(ascii ? nextAscii() : nextBinary(pr.type));
        }
      }
    }
    out.element_names.push_back(e.name);
    out.counts.push_back(e.count);
    out.elements.push_back(std::move(cols));
  }
  return out;
}

template <typename T>
std::vector<T> castColumn(const std::vector<double>& c) {
  std::vector<T> out(c.size());
  for (size_t i = 0; i < c.size(); ++i) out[i] = static_cast<T>(c[i]);
  return out;
}

}  // namespace detail

// mesh_io.cpp:65-101.  Colours are all or nothing (the reference reads red, green, blue and alpha
// inside one try block), stamps need both columns, stamp = stampFromSec(secs) + nsecs.
inline IOData::Ptr IOData::load(const std::string& filename) {
  const detail::PlyColumns ply = detail::readPly(filename);
  const auto* vertex = ply.element("vertex");
  if (!vertex) throw std::runtime_error("PLY: '" + filename + "' has no vertex element");
  auto need = [&](const char* name) {
    const auto* c = vertex->scalar(name);
    if (!c && ply.count("vertex")) throw std::runtime_error(std::string("PLY: vertex property '") + name + "' is missing");
    return c ? detail::castColumn<float>(*c) : std::vector<float>();
  };
  auto data = std::make_shared<IOData>();
  data->x = need("x");
  data->y = need("y");
  data->z = need("z");
  const auto *cr = vertex->scalar("red"), *cg = vertex->scalar("green"), *cb = vertex->scalar("blue"),
             *ca = vertex->scalar("alpha");
  if (cr && cg && cb && ca) {
    data->r = detail::castColumn<uint8_t>(*cr);
    data->g = detail::castColumn<uint8_t>(*cg);
    data->b = detail::castColumn<uint8_t>(*cb);
    data->a = detail::castColumn<uint8_t>(*ca);
  }
  const auto *secs = vertex->scalar("secs"), *nsecs = vertex->scalar("nsecs");
  if (secs && nsecs) {
    data->stamps.resize(secs->size());
    for (size_t i = 0; i < secs->size(); ++i) {
      data->stamps[i] = pgmo_stamp_from_sec((*secs)[i]) + static_cast<Timestamp>((*nsecs)[i]);
    }
  }
  if (const auto* labels = vertex->scalar("label")) data->labels = detail::castColumn<uint32_t>(*labels);
  if (const auto* face = ply.element("face")) {
    if (const auto* lists = face->list("vertex_indices")) {
      data->faces.reserve(lists->size());
      for (const auto& f : *lists) data->faces.push_back(detail::castColumn<uint32_t>(f));
    }
  }
  return data;
}

// mesh_io.cpp:14-63: ASCII, colours only when all three (four) columns match the vertex count,
// stamps split into uint secs / nsecs.
inline void IOData::save(const std::string& filename) const {
  std::FILE* f = std::fopen(filename.c_str(), "w");
  if (!f) throw std::runtime_error("failed to open '" + filename + "'");
  const size_t n = x.size();
  const bool rgb = r.size() == n && g.size() == n && b.size() == n;
  const bool alpha = a.size() == n;
  const bool with_stamps = !stamps.empty();
  const bool with_labels = !labels.empty();
  std::fprintf(f, "ply\nformat ascii 1.0\nelement vertex %zu\nproperty float x\nproperty float y\nproperty float z\n", n);
  if (rgb) std::fprintf(f, "property uchar red\nproperty uchar green\nproperty uchar blue\n");
  if (alpha) std::fprintf(f, "property uchar alpha\n");
  if (with_stamps) std::fprintf(f, "property uint secs\nproperty uint nsecs\n");
  if (with_labels) std::fprintf(f, "property uint label\n");
  std::fprintf(f, "element face %zu\nproperty list uchar uint vertex_indices\nend_header\n", faces.size());
  const int digits = std::numeric_limits<float>::max_digits10;
  for (size_t i = 0; i < n; ++i) {
    std::fprintf(f, "%.*g %.*g %.*g", digits, static_cast<double>(x[i]), digits,
                 static_cast<double>(y[i]), digits, static_cast<double>(z[i]));
    if (rgb) std::fprintf(f, " %u %u %u", unsigned(r[i]), unsigned(g[i]), unsigned(b[i]));
    if (alpha) std::fprintf(f, " %u", unsigned(a[i]));
    if (with_stamps) {
      const Timestamp ns = i < stamps.size() ? stamps[i] : 0;
      std::fprintf(f, " %u %u", static_cast<uint32_t>(ns / 1000000000ull), static_cast<uint32_t>(ns % 1000000000ull));
    }
    if (with_labels) std::fprintf(f, " %u", i < labels.size() ? labels[i] : 0u);
    std::fputc('\n', f);
  }
  for (const auto& face : faces) {
    std::fprintf(f, "%zu", face.size());
    for (uint32_t v : face) std::fprintf(f, " %u", v);
    std::fputc('\n', f);
  }
  std::fclose(f);
}

// mesh_io.cpp:137-149 (+ ReadMesh of mesh_io.h): vertices as PointXYZRGBA (colour only when the
// file has one), faces as pcl::Vertices, stamps when asked for and present.
inline void ReadMeshWithStampsFromPly(const std::string& filename, pcl::PolygonMesh& mesh,
                                      std::vector<Timestamp>* vertex_stamps = nullptr) {
  const IOData::Ptr data = IOData::load(filename);
  pcl::PointCloud<pcl::PointXYZRGBA> vertices;
  vertices.resize(data->x.size());
  const bool colour = data->r.size() == data->x.size();
  for (size_t i = 0; i < data->x.size(); ++i) {
    auto& p = vertices.at(i);
    p.x = data->x[i];
    p.y = data->y[i];
    p.z = data->z[i];
    if (colour) {
      p.r = data->r[i];
      p.g = data->g[i];
      p.b = data->b[i];
      p.a = data->a[i];
    }
  }
  mesh.polygons.clear();
  mesh.polygons.reserve(data->faces.size());
  for (const auto& f : data->faces) {
    pcl::Vertices tri;
    tri.vertices.assign(f.begin(), f.end());
    mesh.polygons.push_back(tri);
  }
  if (vertex_stamps) *vertex_stamps = data->stamps;
  pcl::toPCLPointCloud2(vertices, mesh.cloud);
}

// mesh_io.cpp:124-135
inline void WriteMeshWithStampsToPly(const std::string& filename, const pcl::PolygonMesh& mesh,
                                     const std::vector<Timestamp>& stamps) {
  pcl::PointCloud<pcl::PointXYZRGBA> cloud;
  pcl::fromPCLPointCloud2(mesh.cloud, cloud);
  IOData data;
  const size_t n = cloud.size();
  data.x.resize(n), data.y.resize(n), data.z.resize(n);
  data.r.resize(n), data.g.resize(n), data.b.resize(n), data.a.resize(n);
  for (size_t i = 0; i < n; ++i) {
    const auto& p = cloud.at(i);
    data.x[i] = p.x, data.y[i] = p.y, data.z[i] = p.z;
    data.r[i] = p.r, data.g[i] = p.g, data.b[i] = p.b, data.a[i] = p.a;
  }
  data.stamps = stamps;
  for (const auto& poly : mesh.polygons) data.faces.emplace_back(poly.vertices.begin(), poly.vertices.end());
  data.save(filename);
}
inline void WriteMeshToPly(const std::string& filename, const pcl::PolygonMesh& mesh) {
  WriteMeshWithStampsToPly(filename, mesh, {});
}

// What DeformationGraph::load (deformation_graph_io.cpp:194-349) gives the mesh path:
//   NODE key x y z qx qy qz qw        -> values.insert(key, Pose3(Rot3(qw, qx, qy, qz), (x, y, z)))
//   VERTEX key stamp x y z            -> vertex_positions_[prefix] / vertex_stamps_[prefix] in index
//                                        order (index 0 restarts the prefix, :314-317)
// with the same re-keying when set_robot_id is given (:205-207, :309-312).  NODE_TEMP, BETWEEN,
// DEDGE and PRIOR lines belong to the solver and are counted, not interpreted.
struct GraphFileSummary {
  size_t nodes = 0, vertices = 0, solver_lines = 0;
};
inline GraphFileSummary loadGraphFile(const std::string& filename, DeformationGraphMeshPath* graph,
                                      gtsam::Values* values, bool set_robot_id = false,
                                      size_t new_robot_id = 0) {
  std::ifstream in(filename);
  if (!in) throw std::runtime_error("failed to open '" + filename + "'");
  GraphFileSummary summary;
  std::map<char, std::vector<gtsam::Point3>> positions;
  std::map<char, std::vector<Timestamp>> stamps;
  std::string line;
  while (std::getline(in, line)) {
    std::istringstream ls(line);
    std::string tag;
    if (!(ls >> tag)) continue;
    if (tag == "NODE") {
      uint64_t key = 0;
      double x, y, z, qx, qy, qz, qw;
      ls >> key >> x >> y >> z >> qx >> qy >> qz >> qw;
      unsigned char chr = static_cast<unsigned char>(key >> 56);
      const uint64_t index = key & ((uint64_t(1) << 56) - 1);
      if (set_robot_id) {  // rekey (:178-191): pose prefixes and vertex prefixes move to the new robot
        if (GetRobotIdFromPrefix(chr) >= 0) chr = GetRobotPrefix(new_robot_id);
        else if (GetRobotIdFromVertexPrefix(chr) >= 0) chr = GetVertexPrefix(new_robot_id);
      }
      if (values) {
        values->insert(gtsam::Symbol(chr, index),
                       gtsam::Pose3(gtsam::Rot3(qw, qx, qy, qz), gtsam::Point3(x, y, z)));
      }
      ++summary.nodes;
    } else if (tag == "VERTEX") {
      uint64_t key = 0;
      Timestamp stamp = 0;
      double x, y, z;
      ls >> key >> stamp >> x >> y >> z;
      char prefix = static_cast<char>(key >> 56);
      const uint64_t index = key & ((uint64_t(1) << 56) - 1);
      if (set_robot_id && GetRobotIdFromVertexPrefix(prefix) >= 0) prefix = GetVertexPrefix(new_robot_id);
      if (index == 0) {
        positions[prefix].clear();
        stamps[prefix].clear();
      }
      positions[prefix].push_back(gtsam::Point3(x, y, z));
      stamps[prefix].push_back(stamp);
      ++summary.vertices;
    } else {
      ++summary.solver_lines;
    }
  }
  if (graph) {
    for (const auto& kv : positions) graph->setControlPoints(kv.first, kv.second, stamps[kv.first]);
  }
  return summary;
}

}  // namespace kimera_pgmo_b200
