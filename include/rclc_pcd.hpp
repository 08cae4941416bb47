// rclc_pcd.hpp — the on-disk hand-off of zhijianglu/RCLC between stage 1 and stage 2 (SURVEY.md §8f "next #4"), header only.
//
//   reference call (file:line)                                              here
//   -------------------------------------------------------------------------------------------------
//   pcl::io::loadPCDFile<PoinT>(path, cloud)      apps/BoardSegmentation.cpp:38, apps/Calibrate.cpp:98    rclc_b200::io::loadPCDFile
//   pcl::io::savePCDFileBinary(path, cloud)       apps/BoardSegmentation.cpp:171-172                      rclc_b200::io::savePCDFileBinary
//   T_base (lidar axes -> camera axis order)      src/global.cpp:116-122                                  rclc_b200::io::T_base()
//
// PCD v0.7 (.5/.6 headers parse the same way): DATA ascii and DATA binary; the fields x, y, z, intensity are picked by NAME from
// whatever FIELDS / SIZE / TYPE / COUNT the file declares (extra fields such as Livox's tag / line / timestamp are skipped;
// a missing intensity reads as 0). DATA binary_compressed is refused (-1): the reference never writes it. The writer emits
// exactly what PCL writes for a PointXYZI cloud: FIELDS x y z intensity, four F4 per point, no padding.
// Both functions return 0 on success and -1 on failure, like their PCL namesakes.
#ifndef RCLC_PCD_HPP
#define RCLC_PCD_HPP

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "rclc_ops.hpp"

namespace rclc_b200 {
namespace io {

// src/global.cpp:116-122: rows (0 -1 0), (0 0 -1), (1 0 0): lidar x-forward / z-up becomes camera-like z-forward / y-down.
inline Matrix4f T_base()
{
    Matrix4f T;
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) T(r, c) = 0.f;
    T(0, 1) = -1.f; T(1, 2) = -1.f; T(2, 0) = 1.f; T(3, 3) = 1.f;
    return T;
}

namespace detail {
struct Field { std::string name; int size = 4; char type = 'F'; int count = 1; int offset = 0; };

inline double read_scalar(const unsigned char* p, char type, int size)
{
    switch (type) {
    case 'F': if (size == 4) { float v; std::memcpy(&v, p, 4); return v; } if (size == 8) { double v; std::memcpy(&v, p, 8); return v; } break;
    case 'U': if (size == 1) return *p; if (size == 2) { uint16_t v; std::memcpy(&v, p, 2); return v; } if (size == 4) { uint32_t v; std::memcpy(&v, p, 4); return v; }
              if (size == 8) { uint64_t v; std::memcpy(&v, p, 8); return (double)v; } break;
    case 'I': if (size == 1) return *reinterpret_cast<const int8_t*>(p); if (size == 2) { int16_t v; std::memcpy(&v, p, 2); return v; } if (size == 4) { int32_t v; std::memcpy(&v, p, 4); return v; }
              if (size == 8) { int64_t v; std::memcpy(&v, p, 8); return (double)v; } break;
    }
    return 0.0;
}
}  // namespace detail

inline int loadPCDFile(const std::string& path, PointCloud<PoinT>& cloud)
{
    using detail::Field;
    std::ifstream in(path.c_str(), std::ios::binary);
    if (!in) { std::fprintf(stderr, "[rclc_b200] loadPCDFile: cannot open %s\n", path.c_str()); return -1; }
    std::vector<Field> fields;
    long long width = -1, height = 1, points = -1;
    std::string data_kind, line;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty() || line[0] == '#') continue;
        std::istringstream ls(line);
        std::string key;
        ls >> key;
        if (key == "FIELDS" || key == "COLUMNS") { std::string f; while (ls >> f) { Field fd; fd.name = f; fields.push_back(fd); } }
        else if (key == "SIZE") { for (auto& f : fields) ls >> f.size; }
        else if (key == "TYPE") { for (auto& f : fields) ls >> f.type; }
        else if (key == "COUNT") { for (auto& f : fields) ls >> f.count; }
        else if (key == "WIDTH") ls >> width;
        else if (key == "HEIGHT") ls >> height;
        else if (key == "POINTS") ls >> points;
        else if (key == "DATA") { ls >> data_kind; break; }
    }
    if (fields.empty() || data_kind.empty()) { std::fprintf(stderr, "[rclc_b200] loadPCDFile: %s has no PCD header\n", path.c_str()); return -1; }
    if (points < 0) points = width * height;
    if (points < 0) return -1;
    int rec = 0, ix = -1, iy = -1, iz = -1, ii = -1;
    for (size_t k = 0; k < fields.size(); ++k) {
        fields[k].offset = rec;
        rec += fields[k].size * fields[k].count;
        if (fields[k].name == "x") ix = (int)k; else if (fields[k].name == "y") iy = (int)k; else if (fields[k].name == "z") iz = (int)k;
        else if (fields[k].name == "intensity") ii = (int)k;
    }
    if (ix < 0 || iy < 0 || iz < 0) { std::fprintf(stderr, "[rclc_b200] loadPCDFile: %s has no x / y / z fields\n", path.c_str()); return -1; }
    cloud.points.assign((size_t)points, PoinT());
    if (data_kind == "binary") {
        std::vector<unsigned char> buf((size_t)points * (size_t)rec);
        in.read(reinterpret_cast<char*>(buf.data()), (std::streamsize)buf.size());
        if ((size_t)in.gcount() != buf.size()) { std::fprintf(stderr, "[rclc_b200] loadPCDFile: %s is truncated\n", path.c_str()); return -1; }
        for (long long k = 0; k < points; ++k) {
            const unsigned char* p = buf.data() + (size_t)k * (size_t)rec;
            PoinT& q = cloud.points[(size_t)k];
            q.x = (float)detail::read_scalar(p + fields[(size_t)ix].offset, fields[(size_t)ix].type, fields[(size_t)ix].size);
            q.y = (float)detail::read_scalar(p + fields[(size_t)iy].offset, fields[(size_t)iy].type, fields[(size_t)iy].size);
            q.z = (float)detail::read_scalar(p + fields[(size_t)iz].offset, fields[(size_t)iz].type, fields[(size_t)iz].size);
            q.intensity = ii >= 0 ? (float)detail::read_scalar(p + fields[(size_t)ii].offset, fields[(size_t)ii].type, fields[(size_t)ii].size) : 0.f;
        }
        return 0;
    }
    if (data_kind == "ascii") {
        for (long long k = 0; k < points; ++k) {
            if (!std::getline(in, line)) { std::fprintf(stderr, "[rclc_b200] loadPCDFile: %s is truncated\n", path.c_str()); return -1; }
            std::istringstream ls(line);
            PoinT& q = cloud.points[(size_t)k];
            for (size_t f = 0; f < fields.size(); ++f)
                for (int c = 0; c < fields[f].count; ++c) {
                    std::string tok;
                    if (!(ls >> tok)) return -1;
                    if (c > 0) continue;
                    const float v = (tok == "nan" || tok == "NaN") ? std::nanf("") : (float)std::strtod(tok.c_str(), nullptr);
                    if ((int)f == ix) q.x = v; else if ((int)f == iy) q.y = v; else if ((int)f == iz) q.z = v; else if ((int)f == ii) q.intensity = v;
                }
        }
        return 0;
    }
    std::fprintf(stderr, "[rclc_b200] loadPCDFile: DATA %s is not supported (%s)\n", data_kind.c_str(), path.c_str());
    return -1;
}

inline int savePCDFileBinary(const std::string& path, const PointCloud<PoinT>& cloud)
{
    std::ofstream out(path.c_str(), std::ios::binary);
    if (!out) { std::fprintf(stderr, "[rclc_b200] savePCDFileBinary: cannot open %s\n", path.c_str()); return -1; }
    const size_t n = cloud.points.size();
    std::ostringstream h;
    h << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\n"
      << "WIDTH " << n << "\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << n << "\nDATA binary\n";
    const std::string hs = h.str();
    out.write(hs.data(), (std::streamsize)hs.size());
    std::vector<float> rec(n * 4);
    for (size_t k = 0; k < n; ++k) { const PoinT& p = cloud.points[k]; rec[4 * k] = p.x; rec[4 * k + 1] = p.y; rec[4 * k + 2] = p.z; rec[4 * k + 3] = p.intensity; }
    out.write(reinterpret_cast<const char*>(rec.data()), (std::streamsize)(rec.size() * 4));
    return out.good() ? 0 : -1;
}

}  // namespace io
}  // namespace rclc_b200
#endif  // RCLC_PCD_HPP
