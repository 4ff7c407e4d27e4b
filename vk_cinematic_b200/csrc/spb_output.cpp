// spb_output.cpp -- image files out of the path ("8-bit / EXR writer"):
//   sp_b200_SaveExrImage  the radiance image as an OpenEXR file (the format LoadExrImage,
//                         src/asset_loader/asset_loader.h:11-22, reads back: a rendered frame or a
//                         baked cube face can be fed to the reference as an HdrImage again);
//   sp_b200_SavePpm       the tone-mapped RGBA8 image (sp_b200_ToneMap) as a binary PPM.
// The reference has no writer of its own (its frames go to the swap chain, main.cpp:1607
// VulkanCopyImageFromCPU); both functions follow the loader's house convention, 0 = success.
//
// Written from the OpenEXR file-layout description: single-part scanline file, channels A B G R
// (alphabetical, as the format requires), HALF or FLOAT, compression NONE, ZIPS (one line per
// chunk) or ZIP (16 lines per chunk); sp_b200_SaveExrImageTiled writes the single-part tiled layout
// (one level, a chunk per tile) instead.  ZIP chunks: bytes split into even / odd halves, delta
// predictor, then a zlib stream (RFC 1950) holding one DEFLATE block with the fixed Huffman code
// (RFC 1951 §3.2.6) over a hash-chain LZ77 match search; a chunk that does not shrink is stored raw,
// which every reader takes by its size.  Host code only.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "sp_b200.h"

namespace {

struct BitWriter
{
    std::vector<uint8_t> &out;
    uint64_t acc = 0;
    int n = 0;
    explicit BitWriter(std::vector<uint8_t> &o) : out(o) {}
    void put(uint32_t v, int bits) // LSB first
    {
        acc |= (uint64_t)v << n;
        n += bits;
        while (n >= 8)
        {
            out.push_back((uint8_t)(acc & 0xFF));
            acc >>= 8;
            n -= 8;
        }
    }
    void huff(uint32_t code, int bits) // Huffman codes go in MSB first
    {
        uint32_t r = 0;
        for (int i = 0; i < bits; ++i) r |= ((code >> i) & 1u) << (bits - 1 - i);
        put(r, bits);
    }
    void flush()
    {
        if (n > 0) out.push_back((uint8_t)(acc & 0xFF));
        acc = 0;
        n = 0;
    }
};

// fixed literal/length code, RFC 1951 §3.2.6
void put_symbol(BitWriter &w, uint32_t sym)
{
    if (sym < 144) w.huff(0x30 + sym, 8);
    else if (sym < 256) w.huff(0x190 + (sym - 144), 9);
    else if (sym < 280) w.huff(sym - 256, 7);
    else w.huff(0xC0 + (sym - 280), 8);
}

const uint16_t kLengthBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLengthExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

void put_match(BitWriter &w, uint32_t length, uint32_t distance)
{
    int lc = 28;
    while (kLengthBase[lc] > length) --lc;
    put_symbol(w, 257 + (uint32_t)lc);
    if (kLengthExtra[lc]) w.put(length - kLengthBase[lc], kLengthExtra[lc]);
    int dc = 29;
    while (kDistBase[dc] > distance) --dc;
    w.huff((uint32_t)dc, 5);
    if (kDistExtra[dc]) w.put(distance - kDistBase[dc], kDistExtra[dc]);
}

// zlib stream: header 78 9C, one final fixed-Huffman block, Adler-32 of the input (big endian)
void zlib_compress(const uint8_t *src, size_t n, std::vector<uint8_t> &out)
{
    out.clear();
    out.push_back(0x78);
    out.push_back(0x9C);
    BitWriter w(out);
    w.put(1, 1); // BFINAL
    w.put(1, 2); // BTYPE = 01, fixed Huffman
    const int HASH_BITS = 15, CHAIN = 48;
    const uint32_t WINDOW = 32768, MAX_MATCH = 258, MIN_MATCH = 3;
    std::vector<int32_t> head((size_t)1 << HASH_BITS, -1), prev(n ? n : 1, -1);
    auto hash3 = [&](size_t i) -> uint32_t {
        uint32_t v = (uint32_t)src[i] | ((uint32_t)src[i + 1] << 8) | ((uint32_t)src[i + 2] << 16);
        return (v * 2654435761u) >> (32 - HASH_BITS);
    };
    auto insert = [&](size_t i) {
        if (i + MIN_MATCH > n) return;
        uint32_t h = hash3(i);
        prev[i] = head[h];
        head[h] = (int32_t)i;
    };
    size_t i = 0;
    while (i < n)
    {
        uint32_t bestLen = 0, bestDist = 0;
        if (i + MIN_MATCH <= n)
        {
            int32_t cand = head[hash3(i)];
            const uint32_t limit = (uint32_t)(n - i < MAX_MATCH ? n - i : MAX_MATCH);
            for (int chain = 0; cand >= 0 && chain < CHAIN; ++chain)
            {
                uint32_t dist = (uint32_t)(i - (size_t)cand);
                if (dist > WINDOW) break;
                uint32_t len = 0;
                while (len < limit && src[(size_t)cand + len] == src[i + len]) ++len;
                if (len > bestLen)
                {
                    bestLen = len;
                    bestDist = dist;
                    if (len == limit) break;
                }
                cand = prev[(size_t)cand];
            }
        }
        if (bestLen >= MIN_MATCH)
        {
            put_match(w, bestLen, bestDist);
            for (uint32_t k = 0; k < bestLen; ++k) insert(i + k);
            i += bestLen;
        }
        else
        {
            put_symbol(w, src[i]);
            insert(i);
            ++i;
        }
    }
    put_symbol(w, 256);
    w.flush();
    uint32_t a = 1, b = 0;
    for (size_t k = 0; k < n; ++k)
    {
        a = (a + src[k]) % 65521u;
        b = (b + a) % 65521u;
    }
    uint32_t adler = (b << 16) | a;
    out.push_back((uint8_t)(adler >> 24));
    out.push_back((uint8_t)(adler >> 16));
    out.push_back((uint8_t)(adler >> 8));
    out.push_back((uint8_t)adler);
}

// float -> half, round to nearest even; overflow -> infinity; NaN stays NaN
uint16_t float_to_half(float f)
{
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u, mag = x & 0x7FFFFFFFu;
    if (mag >= 0x7F800000u) return (uint16_t)(sign | 0x7C00u | (mag > 0x7F800000u ? 0x200u | ((mag >> 13) & 0x3FFu) : 0u));
    if (mag >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u); // rounds to >= 65520: infinity
    if (mag >= 0x38800000u)                                     // normal half
    {
        uint32_t v = mag - 0x38000000u; // rebias exponent 127 -> 15
        uint32_t r = v >> 13, rest = v & 0x1FFFu;
        if (rest > 0x1000u || (rest == 0x1000u && (r & 1u))) ++r;
        return (uint16_t)(sign | r);
    }
    if (mag < 0x33000000u) return (uint16_t)sign; // below half of the smallest subnormal
    // subnormal half: value = m * 2^-24
    uint32_t e = mag >> 23, m = (mag & 0x7FFFFFu) | 0x800000u;
    uint32_t shift = 126u - e; // 14..24 here
    uint32_t r = m >> shift, rest = m & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
    if (rest > halfway || (rest == halfway && (r & 1u))) ++r;
    return (uint16_t)(sign | r);
}

struct Header
{
    std::vector<uint8_t> b;
    void raw(const void *p, size_t n) { b.insert(b.end(), (const uint8_t *)p, (const uint8_t *)p + n); }
    void str(const char *s) { raw(s, strlen(s) + 1); }
    void i32(int32_t v) { uint8_t t[4] = {(uint8_t)v, (uint8_t)(v >> 8), (uint8_t)(v >> 16), (uint8_t)(v >> 24)}; raw(t, 4); }
    void f32(float v) { int32_t t; memcpy(&t, &v, 4); i32(t); }
    void attr(const char *name, const char *type, const std::vector<uint8_t> &value)
    {
        str(name);
        str(type);
        i32((int32_t)value.size());
        raw(value.data(), value.size());
    }
};

} // namespace

static int save_exr(const HdrImage *image, const char *path, u32 pixelType, u32 compression, u32 tileW, u32 tileH);

extern "C" int sp_b200_SaveExrImage(const HdrImage *image, const char *path, u32 pixelType, u32 compression)
{
    return save_exr(image, path, pixelType, compression, 0, 0);
}

extern "C" int sp_b200_SaveExrImageTiled(const HdrImage *image, const char *path, u32 pixelType, u32 compression,
                                         u32 tileWidth, u32 tileHeight)
{
    if (tileWidth == 0 || tileHeight == 0 || tileWidth > 65536 || tileHeight > 65536) return 1;
    return save_exr(image, path, pixelType, compression, tileWidth, tileHeight);
}

// tileW == 0: scanline file; else single-part tiled file, one level, tiles in row-major order
static int save_exr(const HdrImage *image, const char *path, u32 pixelType, u32 compression, u32 tileW, u32 tileH)
{
    if (!image || !image->pixels || !path || image->width == 0 || image->height == 0) return 1;
    if (image->width > 65536 || image->height > 65536) return 1;
    if (pixelType != SP_B200_EXR_HALF && pixelType != SP_B200_EXR_FLOAT) return 1;
    if (compression != SP_B200_EXR_NONE && compression != SP_B200_EXR_ZIPS && compression != SP_B200_EXR_ZIP) return 1;
    try
    {
        const uint32_t width = image->width, height = image->height;
        const bool tiled = tileW != 0;
        const uint32_t linesPerBlock = tiled ? tileH : (compression == SP_B200_EXR_ZIP ? 16u : 1u);
        const uint32_t blockWidth = tiled ? tileW : width;
        const size_t sampleBytes = pixelType == SP_B200_EXR_HALF ? 2 : 4;
        const uint32_t blocksX = (width + blockWidth - 1) / blockWidth;
        const uint32_t blocksY = (height + linesPerBlock - 1) / linesPerBlock;
        const uint32_t chunks = blocksX * blocksY;

        Header h;
        h.i32(20000630); // magic
        h.i32(tiled ? 2 | 0x200 : 2); // version 2, single part; bit 9: tiled
        {
            Header v;
            const char *names[4] = {"A", "B", "G", "R"};
            for (int c = 0; c < 4; ++c)
            {
                v.str(names[c]);
                v.i32((int32_t)pixelType);
                v.i32(0); // pLinear + 3 reserved bytes
                v.i32(1); // xSampling
                v.i32(1); // ySampling
            }
            v.b.push_back(0);
            h.attr("channels", "chlist", v.b);
        }
        { Header v; v.b.push_back((uint8_t)compression); h.attr("compression", "compression", v.b); }
        { Header v; v.i32(0); v.i32(0); v.i32((int32_t)width - 1); v.i32((int32_t)height - 1); h.attr("dataWindow", "box2i", v.b); h.attr("displayWindow", "box2i", v.b); }
        { Header v; v.b.push_back(0); h.attr("lineOrder", "lineOrder", v.b); } // increasing y
        { Header v; v.f32(1.0f); h.attr("pixelAspectRatio", "float", v.b); }
        { Header v; v.f32(0.0f); v.f32(0.0f); h.attr("screenWindowCenter", "v2f", v.b); }
        { Header v; v.f32(1.0f); h.attr("screenWindowWidth", "float", v.b); }
        if (tiled)
        {
            Header v;
            v.i32((int32_t)tileW);
            v.i32((int32_t)tileH);
            v.b.push_back(0); // ONE_LEVEL, ROUND_DOWN
            h.attr("tiles", "tiledesc", v.b);
        }
        h.b.push_back(0);

        std::vector<uint8_t> file(h.b);
        const size_t tableAt = file.size();
        file.resize(tableAt + (size_t)chunks * 8);

        static const int kSource[4] = {3, 2, 1, 0}; // A B G R from RGBA
        std::vector<uint8_t> raw, filtered, packed;
        for (uint32_t chunk = 0; chunk < chunks; ++chunk)
        {
            const uint32_t blockX = chunk % blocksX, blockY = chunk / blocksX;
            const uint32_t x0 = blockX * blockWidth, y0 = blockY * linesPerBlock;
            const uint32_t cols = width - x0 < blockWidth ? width - x0 : blockWidth;
            const uint32_t lines = height - y0 < linesPerBlock ? height - y0 : linesPerBlock;
            const size_t rowBytes = (size_t)cols * sampleBytes * 4;
            raw.resize(rowBytes * lines);
            for (uint32_t l = 0; l < lines; ++l)
            {
                const float *src = image->pixels + ((size_t)(y0 + l) * width + x0) * 4;
                uint8_t *dst = raw.data() + rowBytes * l;
                for (int c = 0; c < 4; ++c)
                    for (uint32_t x = 0; x < cols; ++x)
                    {
                        float f = src[(size_t)x * 4 + kSource[c]];
                        if (pixelType == SP_B200_EXR_HALF)
                        {
                            uint16_t v = float_to_half(f);
                            *dst++ = (uint8_t)v;
                            *dst++ = (uint8_t)(v >> 8);
                        }
                        else
                        {
                            uint32_t v;
                            memcpy(&v, &f, 4);
                            *dst++ = (uint8_t)v;
                            *dst++ = (uint8_t)(v >> 8);
                            *dst++ = (uint8_t)(v >> 16);
                            *dst++ = (uint8_t)(v >> 24);
                        }
                    }
            }
            const std::vector<uint8_t> *payload = &raw;
            if (compression != SP_B200_EXR_NONE)
            {
                const size_t n = raw.size(), half = (n + 1) / 2;
                filtered.resize(n);
                for (size_t k = 0; k < n; ++k) filtered[(k & 1) ? half + k / 2 : k / 2] = raw[k];
                for (size_t k = n; k-- > 1;) filtered[k] = (uint8_t)(filtered[k] - filtered[k - 1] + 128);
                zlib_compress(filtered.data(), n, packed);
                if (packed.size() < n) payload = &packed;
            }
            const uint64_t at = file.size();
            for (int k = 0; k < 8; ++k) file[tableAt + (size_t)chunk * 8 + k] = (uint8_t)(at >> (8 * k));
            Header c;
            if (tiled)
            {
                c.i32((int32_t)blockX);
                c.i32((int32_t)blockY);
                c.i32(0); // level x
                c.i32(0); // level y
            }
            else c.i32((int32_t)y0);
            c.i32((int32_t)payload->size());
            file.insert(file.end(), c.b.begin(), c.b.end());
            file.insert(file.end(), payload->begin(), payload->end());
        }
        FILE *f = fopen(path, "wb");
        if (!f) return 1;
        bool ok = fwrite(file.data(), 1, file.size(), f) == file.size();
        ok = (fclose(f) == 0) && ok;
        return ok ? 0 : 1;
    }
    catch (...)
    {
        return 1;
    }
}

extern "C" int sp_b200_SavePpm(const u32 *rgba8, u32 width, u32 height, const char *path)
{
    if (!rgba8 || !path || width == 0 || height == 0) return 1;
    FILE *f = fopen(path, "wb");
    if (!f) return 1;
    bool ok = fprintf(f, "P6\n%u %u\n255\n", width, height) > 0;
    std::vector<uint8_t> row((size_t)width * 3);
    for (u32 y = 0; y < height && ok; ++y)
    {
        for (u32 x = 0; x < width; ++x)
        {
            u32 p = rgba8[(size_t)y * width + x]; // r in the low byte (ToColor, math_lib.h:523-532)
            row[(size_t)x * 3 + 0] = (uint8_t)p;
            row[(size_t)x * 3 + 1] = (uint8_t)(p >> 8);
            row[(size_t)x * 3 + 2] = (uint8_t)(p >> 16);
        }
        ok = fwrite(row.data(), 1, row.size(), f) == row.size();
    }
    ok = (fclose(f) == 0) && ok;
    return ok ? 0 : 1;
}
