// tools/san_lbvh.cpp -- the host half of the device BVH builder (lbvh_build_binary_host, bvh4_from_binary,
// build_bvh4_lbvh_host; the device collapse's emulation and bvh4_adopt_device_tree, the gate between what the GPU
// returns and what the kernels walk) on random, degenerate and corrupted inputs under ASan / UBSan.  Not product code.
//   g++ -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=all -std=c++17 -ffp-contract=off \
//       -Ivk_cinematic_b200/csrc tools/san_lbvh.cpp vk_cinematic_b200/csrc/spb_bvh.cpp -o /tmp/san_lbvh && /tmp/san_lbvh
#include <cmath>
#include <cstdio>
#include <vector>
#include "spb_bvh.h"
using namespace spb;
static uint32_t s = 777;
static float rnd() { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return (float)(s >> 8) / 16777216.0f; }
// independent check of an accepted tree: every primitive reached exactly once from node 0, every reference in range,
// no node visited twice, no more stack than the tree claims
static bool walk_ok(const Bvh4 &t, uint32_t n)
{
    if (t.nodes.empty() || t.slotPrim.size() != n) return false;
    std::vector<char> node(t.nodes.size(), 0), prim(n, 0);
    std::vector<uint32_t> todo(1, 0u);
    node[0] = 1;
    uint32_t reached = 0;
    while (!todo.empty())
    {
        uint32_t i = todo.back();
        todo.pop_back();
        for (int k = 0; k < 4; ++k)
        {
            uint32_t ref = t.nodes[i].ref[k];
            if (ref == 0xFFFFFFFFu) continue;
            if (ref & 0x80000000u)
            {
                uint32_t slot = ref & 0x7FFFFFFFu;
                if (slot >= n || t.slotPrim[slot] >= n || prim[t.slotPrim[slot]]) return false;
                prim[t.slotPrim[slot]] = 1;
                reached++;
            }
            else
            {
                if (ref >= t.nodes.size() || node[ref]) return false;
                node[ref] = 1;
                todo.push_back(ref);
            }
        }
    }
    return reached == n;
}

int main()
{
    int built = 0, fell = 0, adopted = 0, refused = 0, survived = 0;
    for (int round = 0; round < 400; ++round)
    {
        uint32_t n = 1 + (uint32_t)(rnd() * (round % 7 == 0 ? 5000 : 200));
        std::vector<float> mn(n * 3), mx(n * 3);
        int kind = round % 6;
        for (uint32_t i = 0; i < n; ++i)
            for (int a = 0; a < 3; ++a)
            {
                float c = kind == 0 ? rnd() : kind == 1 ? 0.5f : kind == 2 ? ldexpf(1.0f, -(int)(i % 60)) : kind == 3 ? (a == 1 ? 2.0f : rnd()) : rnd() * 1e30f;
                float r = kind == 4 ? -rnd() : rnd() * 0.01f; // kind 4: inverted boxes
                mn[i * 3 + a] = c - r;
                mx[i * 3 + a] = c + r;
            }
        if (kind == 5 && n > 3) { mn[4] = NAN; mx[7] = INFINITY; mn[9] = -INFINITY; }
        Bvh4 t = build_bvh4_lbvh_host(mn.data(), mx.data(), n);
        if (t.slotPrim.size() != n) { printf("leaf count %zu != %u (kind %d)\n", t.slotPrim.size(), n, kind); return 1; }
        std::vector<char> seen(n, 0);
        for (uint32_t p : t.slotPrim) { if (p >= n || seen[p]) { printf("bad slotPrim\n"); return 1; } seen[p] = 1; }
        BinaryTree bt = lbvh_build_binary_host(mn.data(), mx.data(), n);
        Bvh4 out;
        if (n >= 2 && bvh4_from_binary(mn.data(), mx.data(), n, bt, &out)) built++; else fell++;
        // corrupt the binary tree: must be refused or still produce a complete tree, never crash
        if (n >= 4)
        {
            BinaryTree bad = bt;
            bad.children[(size_t)(rnd() * (bad.children.size() - 1))] = (uint32_t)(rnd() * 4e9f);
            Bvh4 o2;
            if (bvh4_from_binary(mn.data(), mx.data(), n, bad, &o2) && o2.slotPrim.size() != n) { printf("corrupt tree accepted\n"); return 1; }
        }
        // the device collapse (emulated) and its gate: accepted trees must walk; damaged ones -- random words of the
        // node array and of the slot order overwritten -- must be refused or still walk, never crash
        DeviceTree4 dt;
        if (n >= 2 && lbvh_collapse_host_emulation(mn.data(), mx.data(), n, bt, &dt))
        {
            Bvh4 a;
            if (bvh4_adopt_device_tree(mn.data(), mx.data(), n, dt, &a))
            {
                adopted++;
                if (!walk_ok(a, n)) { printf("adopted tree does not walk (kind %d, n %u)\n", kind, n); return 1; }
            }
            else refused++;
            for (int damage = 0; damage < 8; ++damage)
            {
                DeviceTree4 bad = dt;
                int hits = 1 + (int)(rnd() * 3);
                for (int h = 0; h < hits; ++h)
                {
                    if (rnd() < 0.8f) bad.nodes[(size_t)(rnd() * (bad.nodes.size() - 1))] = rnd() < 0.5f ? (uint32_t)(rnd() * 4e9f) : (uint32_t)(rnd() * 2.0f * n);
                    else bad.slotPrim[(size_t)(rnd() * (n - 1))] = (uint32_t)(rnd() * 1.5f * n);
                }
                if (damage == 7) bad.nodes.resize(bad.nodes.size() - 32 * (size_t)(bad.nodes.size() > 64));
                Bvh4 b;
                if (bvh4_adopt_device_tree(mn.data(), mx.data(), n, bad, &b))
                {
                    survived++;
                    if (!walk_ok(b, n)) { printf("damaged tree accepted and does not walk (kind %d, n %u)\n", kind, n); return 1; }
                }
            }
        }
    }
    printf("400 rounds: %d built by the LBVH path, %d handed to the SAH builder; device-collapse trees: %d adopted, %d refused, "
           "%d of %d damaged copies accepted (all of them still complete trees); no sanitizer report\n",
           built, fell, adopted, refused, survived, 8 * (adopted + refused));
}
