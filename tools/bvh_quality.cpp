// Development aid: SAH-style quality of the 4-wide tree build_bvh4() produces for a set of boxes.
//   g++ -O2 -std=c++17 -I vk_cinematic_b200/csrc tools/bvh_quality.cpp vk_cinematic_b200/csrc/spb_bvh.cpp -o /tmp/bq/bq
//   /tmp/bq/bq boxes.bin      (u32 n, n x 3 f32 min, n x 3 f32 max)
// Prints the expected number of 4-wide nodes a random line through the root box visits
// (sum of node areas / root area) and the expected number of primitive boxes it enters.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "spb_bvh.h"

static double area(const float *mn, const float *mx)
{
    double dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
    return dx * dy + dy * dz + dz * dx;
}

int main(int argc, char **argv)
{
    FILE *f = fopen(argv[1], "rb");
    uint32_t n = 0;
    if (!f || fread(&n, 4, 1, f) != 1) return 1;
    std::vector<float> mn((size_t)n * 3), mx((size_t)n * 3);
    if (fread(mn.data(), 4, mn.size(), f) != mn.size() || fread(mx.data(), 4, mx.size(), f) != mx.size()) return 1;
    fclose(f);
    auto t0 = std::chrono::steady_clock::now();
    spb::Bvh4 b = spb::build_bvh4(mn.data(), mx.data(), n);
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    double root = area(b.rootMin, b.rootMax), nodes = 1.0, leaves = 0.0;
    for (const spb::Node4 &nd : b.nodes)
        for (int k = 0; k < 4; ++k)
        {
            if (nd.ref[k] == 0xFFFFFFFFu) continue;
            float cmn[3] = {nd.bmin[0][k], nd.bmin[1][k], nd.bmin[2][k]}, cmx[3] = {nd.bmax[0][k], nd.bmax[1][k], nd.bmax[2][k]};
            double a = area(cmn, cmx) / root;
            if (nd.ref[k] & 0x80000000u) leaves += a; else nodes += a;
        }
    printf("%s: prims %u nodes4 %zu depth %u stackNeed %u  E[node visits] %.3f  E[leaf boxes entered] %.3f  build %.3f s\n",
           argv[1], n, b.nodes.size(), b.maxDepth, b.stackNeed, nodes, leaves, secs);
    return 0;
}
