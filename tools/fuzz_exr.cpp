// tools/fuzz_exr.cpp -- mutation fuzzing of the library's LoadExrImage under ASan / UBSan.
// Not product code.  Seeds: the fixture files given on the command line plus tiled / scanline files
// written by sp_b200_SaveExrImage*.  Every mutated file must either load or be refused (return 1);
// any sanitizer report aborts.
//
//   g++ -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=all -std=c++17 -pthread \
//       -Iinclude tools/fuzz_exr.cpp vk_cinematic_b200/csrc/spb_assets.cpp \
//       vk_cinematic_b200/csrc/spb_output.cpp -o /tmp/fuzz_exr
//   /tmp/fuzz_exr 20000 tests/golden/exr/*.exr
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "sp_b200.h"

static uint32_t g_state = 0x1A34C249u;
static uint32_t rnd()
{
    g_state ^= g_state << 13;
    g_state ^= g_state >> 17;
    g_state ^= g_state << 5;
    return g_state;
}

static std::vector<uint8_t> read_file(const char *path)
{
    std::vector<uint8_t> out;
    FILE *f = fopen(path, "rb");
    if (!f) return out;
    uint8_t buf[65536];
    size_t got;
    while ((got = fread(buf, 1, sizeof(buf), f)) > 0) out.insert(out.end(), buf, buf + got);
    fclose(f);
    return out;
}

int main(int argc, char **argv)
{
    int iterations = argc > 1 ? atoi(argv[1]) : 1000;
    std::vector<std::vector<uint8_t>> seeds;
    for (int i = 2; i < argc; ++i) seeds.push_back(read_file(argv[i]));
    // own seeds: tiled and scanline, both pixel types
    std::vector<float> px(37 * 29 * 4);
    for (size_t i = 0; i < px.size(); ++i) px[i] = (float)(rnd() % 1000) * 0.01f;
    HdrImage img = {px.data(), 37, 29};
    const char *tmp = "/tmp/fuzz_exr_seed.exr";
    for (u32 type = 1; type <= 2; ++type)
        for (u32 comp : {0u, 2u, 3u})
        {
            if (sp_b200_SaveExrImage(&img, tmp, type, comp) == 0) seeds.push_back(read_file(tmp));
            if (sp_b200_SaveExrImageTiled(&img, tmp, type, comp, 16, 8) == 0) seeds.push_back(read_file(tmp));
        }
    size_t loaded = 0, refused = 0;
    const char *path = "/tmp/fuzz_exr_case.exr";
    for (int it = 0; it < iterations; ++it)
    {
        std::vector<uint8_t> data = seeds[rnd() % seeds.size()];
        if (data.empty()) continue;
        int edits = 1 + (int)(rnd() % 4);
        for (int e = 0; e < edits; ++e)
        {
            size_t at = rnd() % data.size();
            switch (rnd() % 5)
            {
            case 0: data[at] ^= (uint8_t)(1u << (rnd() % 8)); break;
            case 1: data[at] = (uint8_t)rnd(); break;
            case 2: data.resize(at + 1); break;                                        // truncate
            case 3: if (at + 4 <= data.size()) { uint32_t v = rnd(); memcpy(&data[at], &v, 4); } break;
            default: if (at + 4 <= data.size()) { uint32_t v = rnd() % 3 ? 0xFFFFFFFFu : 0x7FFFFFFFu; memcpy(&data[at], &v, 4); } break;
            }
        }
        // bias towards the header and the offset table, where the structure lives
        if (rnd() % 3 == 0 && data.size() > 400) data[rnd() % 400] = (uint8_t)rnd();
        FILE *f = fopen(path, "wb");
        fwrite(data.data(), 1, data.size(), f);
        fclose(f);
        HdrImage out = {nullptr, 0, 0};
        if (LoadExrImage(&out, path) == 0)
        {
            loaded++;
            // touch every pixel: the buffer must really be width * height * 4 floats
            volatile float sum = 0;
            for (size_t i = 0; i < (size_t)out.width * out.height * 4; i += 97) sum = sum + out.pixels[i];
            free(out.pixels);
        }
        else refused++;
    }
    printf("%d mutated files: %zu loaded, %zu refused, no sanitizer report\n", iterations, loaded, refused);
    return 0;
}
