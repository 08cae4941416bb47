// PCD hand-off (include/rclc_pcd.hpp): no device needed. Exit 0 = all checks passed.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <random>

#include "rclc_pcd.hpp"

using namespace rclc_b200;
static int g_fail = 0;
#define CHECK(cond, ...) do { if (!(cond)) { std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); ++g_fail; } } while (0)

int main(int argc, char** argv)
{
    const std::string golden = argc > 1 ? argv[1] : "tests/golden/livox_ascii.pcd", tmp = argc > 2 ? argv[2] : "/tmp/rclc_pcd_test";
    // ---- ascii file with extra fields and a NaN row ----
    PointCloud<PoinT> a;
    CHECK(io::loadPCDFile(golden, a) == 0 && a.size() == 4, "ascii load, %zu points", a.size());
    if (a.size() == 4) {
        CHECK(a.points[0].x == 4.25f && a.points[0].y == -0.5f && a.points[0].z == 0.125f && a.points[0].intensity == 101.f, "row 0");
        CHECK(a.points[1].intensity == 14.5f && a.points[3].y == 1e-3f && a.points[3].z == 0.25f && a.points[3].intensity == 255.f, "rows 1, 3");
        CHECK(std::isnan(a.points[2].x) && std::isnan(a.points[2].z), "NaN row");
    }
    // ---- binary round trip, bit for bit; the header is PCL's for PointXYZI ----
    PointCloud<PoinT> c;
    std::mt19937 rng(1); std::uniform_real_distribution<float> u(-50.f, 50.f);
    for (int i = 0; i < 10007; ++i) { PoinT p; p.x = u(rng); p.y = u(rng); p.z = u(rng); p.intensity = std::fabs(u(rng)); c.points.push_back(p); }
    CHECK(io::savePCDFileBinary(tmp + "_a.pcd", c) == 0, "save");
    PointCloud<PoinT> d;
    CHECK(io::loadPCDFile(tmp + "_a.pcd", d) == 0 && d.size() == c.size(), "reload");
    bool same = d.size() == c.size();
    for (size_t i = 0; same && i < c.size(); ++i)
        same = std::memcmp(&c.points[i].x, &d.points[i].x, 12) == 0 && std::memcmp(&c.points[i].intensity, &d.points[i].intensity, 4) == 0;
    CHECK(same, "round trip is bit exact");
    {
        std::ifstream in((tmp + "_a.pcd").c_str(), std::ios::binary);
        std::string head(400, '\0');
        in.read(&head[0], 400);
        const char* want = "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\nWIDTH 10007\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS 10007\nDATA binary\n";
        CHECK(head.compare(0, std::strlen(want), want) == 0, "header text");
        in.seekg(0, std::ios::end);
        CHECK((size_t)in.tellg() == std::strlen(want) + 10007u * 16u, "file size: 16 bytes per point after the header");
    }
    // ---- binary file with a different field order, a double field and a trailing uint16 ----
    {
        std::ofstream out((tmp + "_b.pcd").c_str(), std::ios::binary);
        out << "VERSION .7\nFIELDS intensity z y x ring\nSIZE 4 8 4 4 2\nTYPE F F F F U\nCOUNT 1 1 1 1 1\nWIDTH 2\nHEIGHT 1\nPOINTS 2\nDATA binary\n";
        for (int k = 0; k < 2; ++k) {
            const float I = 7.f + k, y = 2.f + k, x = 1.f + k; const double z = 3.0 + k; const uint16_t ring = (uint16_t)(9 + k);
            out.write((const char*)&I, 4); out.write((const char*)&z, 8); out.write((const char*)&y, 4); out.write((const char*)&x, 4); out.write((const char*)&ring, 2);
        }
    }
    PointCloud<PoinT> e;
    CHECK(io::loadPCDFile(tmp + "_b.pcd", e) == 0 && e.size() == 2, "reordered binary load");
    if (e.size() == 2) CHECK(e.points[1].x == 2.f && e.points[1].y == 3.f && e.points[1].z == 4.f && e.points[1].intensity == 8.f, "reordered fields");
    // ---- failures are -1, like PCL ----
    PointCloud<PoinT> f;
    CHECK(io::loadPCDFile(tmp + "_missing.pcd", f) == -1, "missing file");
    { std::ofstream out((tmp + "_c.pcd").c_str(), std::ios::binary); out << "VERSION .7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 5\nHEIGHT 1\nPOINTS 5\nDATA binary\nxx"; }
    CHECK(io::loadPCDFile(tmp + "_c.pcd", f) == -1, "truncated file");
    // ---- T_base (src/global.cpp:116-122): (x, y, z)_lidar -> (-y, -z, x) ----
    const Matrix4f T = io::T_base();
    const float p[3] = {1.f, 2.f, 3.f};
    float q[3];
    for (int r = 0; r < 3; ++r) q[r] = T(r, 0) * p[0] + T(r, 1) * p[1] + T(r, 2) * p[2] + T(r, 3);
    CHECK(q[0] == -2.f && q[1] == -3.f && q[2] == 1.f && T(3, 3) == 1.f, "T_base");
    if (g_fail) { std::printf("%d check(s) failed\n", g_fail); return 1; }
    std::printf("all PCD checks passed\n");
    return 0;
}
