// tools/fuzz_obj.cpp -- mutation fuzzing of sp_b200_LoadObj under ASan / UBSan.  Not product code.
//
//   g++ -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=all -std=c++17 -pthread -Iinclude \
//       tools/fuzz_obj.cpp vk_cinematic_b200/csrc/spb_assets.cpp -o /tmp/fuzz_obj
//   /tmp/fuzz_obj 20000 tests/golden/quad_mix.obj [more.obj ...]
//
// Every mutated file must load (indices in range, three per triangle) or be refused (return 0).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "sp_b200.h"

static uint32_t g_state = 0x45BA12F3u;
static uint32_t rnd()
{
    g_state ^= g_state << 13;
    g_state ^= g_state >> 17;
    g_state ^= g_state << 5;
    return g_state;
}

int main(int argc, char **argv)
{
    int iterations = argc > 1 ? atoi(argv[1]) : 1000;
    std::vector<std::string> seeds;
    for (int i = 2; i < argc; ++i)
    {
        FILE *f = fopen(argv[i], "rb");
        if (!f) continue;
        std::string s;
        char buf[65536];
        size_t got;
        while ((got = fread(buf, 1, sizeof(buf), f)) > 0) s.append(buf, got);
        fclose(f);
        if (s.size() > 200000) s.resize(200000); // keep iterations fast: a prefix of a big mesh is still an OBJ
        seeds.push_back(s);
    }
    if (seeds.empty()) return 1;
    static const char *tokens[] = {"f ", "v ", "vn ", "vt ", "/", "//", "-1", "0", "999999999", "1e40", "nan", "\n", " ",
                                   "f 1 2 3 4 5 6 7 8\n", "f -1/-1/-1 -2/-2/-2 -3/-3/-3\n", "v 1 2\n", "f 1/ 2/ 3/\n", "#", "\r\n", "-"};
    size_t loaded = 0, refused = 0;
    const char *path = "/tmp/fuzz_obj_case.obj";
    for (int it = 0; it < iterations; ++it)
    {
        std::string data = seeds[rnd() % seeds.size()];
        int edits = 1 + (int)(rnd() % 6);
        for (int e = 0; e < edits && !data.empty(); ++e)
        {
            size_t at = rnd() % data.size();
            switch (rnd() % 5)
            {
            case 0: data[at] = (char)rnd(); break;
            case 1: data.insert(at, tokens[rnd() % (sizeof(tokens) / sizeof(tokens[0]))]); break;
            case 2: data.erase(at, 1 + rnd() % 16); break;
            case 3: data.resize(at); break;
            default: data[at] = "0123456789-/. \n"[rnd() % 15]; break;
            }
        }
        FILE *f = fopen(path, "wb");
        fwrite(data.data(), 1, data.size(), f);
        fclose(f);
        sp_b200_MeshData mesh = {nullptr, nullptr, 0, 0};
        if (sp_b200_LoadObj(path, &mesh) == 1)
        {
            loaded++;
            if (mesh.indexCount % 3 != 0) { fprintf(stderr, "index count %u not a multiple of 3\n", mesh.indexCount); return 2; }
            for (u32 i = 0; i < mesh.indexCount; ++i)
                if (mesh.indices[i] >= mesh.vertexCount) { fprintf(stderr, "index out of range\n"); return 2; }
            volatile float sum = 0;
            for (u32 i = 0; i < mesh.vertexCount; ++i) sum = sum + mesh.vertices[i].position.x;
            sp_b200_FreeMeshData(&mesh);
        }
        else refused++;
    }
    printf("%d mutated files: %zu loaded, %zu refused, no sanitizer report\n", iterations, loaded, refused);
    return 0;
}
