// spb_assets.cpp -- host-side asset input for the sp_ path: Wavefront OBJ.
//
// The reference loads meshes through assimp (src/mesh.cpp:5-62: aiProcess_Triangulate |
// aiProcess_JoinIdenticalVertices | ..., first mesh of the file, positions + normals into
// VertexPNT, three indices per face).  assimp is a third-party library that is neither vendored
// in the reference checkout nor installed here, so its output order cannot be reproduced; what
// the path needs from it is a VertexPNT[] / u32[] pair, and this loader defines that pair the way
// tools/make_mesh_fixtures.py (which made assets/*.npz, the meshes every parity test uses) does:
//   * one output vertex per distinct (v, vt, vn) index triple, in first-seen order -- what
//     JoinIdenticalVertices yields for files whose corners share indices;
//   * polygons fan-triangulated in file order, so triangle i is the i-th triangle of the file;
//   * numbers parsed as double and rounded once to float (Python's float() -> numpy float32);
//   * missing vn -> normal (0,0,0); missing vt -> textureCoord (0,0); negative (relative) indices
//     and the four corner syntaxes a, a/t, a//n, a/t/n are accepted; objects/groups are merged.
// No CUDA in this file.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <thread>
#include <tuple>
#include <vector>

#include "../../include/sp_b200.h"

namespace {

// parses "a", "a/t", "a//n" or "a/t/n" starting at p; returns the end of the token or null
const char *parse_corner(const char *p, long counts[3], long out[3])
{
    for (int k = 0; k < 3; ++k) out[k] = -1; // -1 = absent
    for (int k = 0; k < 3; ++k)
    {
        if (*p != '/' || k == 0)
        {
            char *end = nullptr;
            long v = strtol(p, &end, 10);
            if (end != p)
            {
                if (v > 0) out[k] = v - 1;
                else if (v < 0) out[k] = counts[k] + v; // relative to the records read so far
                else return nullptr;                     // index 0 is not valid OBJ
                if (out[k] < 0 || out[k] >= counts[k]) return nullptr;
                p = end;
            }
            else if (k == 0) return nullptr;
        }
        if (*p != '/') break;
        ++p; // the separator before field k + 1
    }
    return p;
}

} // namespace

extern "C" int sp_b200_LoadObj(const char *path, sp_b200_MeshData *out)
{
    if (!out) return 0;
    memset(out, 0, sizeof(*out));
    FILE *f = path ? fopen(path, "rb") : nullptr;
    if (!f) return 0;
    std::vector<float> pos, nrm, tex;
    std::vector<VertexPNT> vertices;
    std::vector<u32> indices;
    std::map<std::tuple<long, long, long>, u32> indexOf;
    std::vector<char> line(1 << 16);
    bool ok = true;
    while (ok && fgets(line.data(), (int)line.size(), f))
    {
        const char *p = line.data();
        while (*p == ' ' || *p == '\t') ++p;
        if (p[0] == 'v' && (p[1] == ' ' || p[1] == '\t' || ((p[1] == 'n' || p[1] == 't') && (p[2] == ' ' || p[2] == '\t'))))
        {
            std::vector<float> &dst = p[1] == 'n' ? nrm : (p[1] == 't' ? tex : pos);
            const int want = p[1] == 't' ? 2 : 3;
            p += p[1] == ' ' || p[1] == '\t' ? 1 : 2;
            for (int k = 0; k < want; ++k)
            {
                char *end = nullptr;
                double v = strtod(p, &end);
                if (end == p) { if (want == 2 && k == 1) v = 0.0; else { ok = false; break; } }
                dst.push_back((float)v);
                p = end;
            }
        }
        else if (p[0] == 'f' && (p[1] == ' ' || p[1] == '\t'))
        {
            ++p;
            long counts[3] = {(long)(pos.size() / 3), (long)(tex.size() / 2), (long)(nrm.size() / 3)};
            std::vector<u32> corners;
            for (;;)
            {
                while (*p == ' ' || *p == '\t') ++p;
                if (*p == '\0' || *p == '\n' || *p == '\r' || *p == '#') break;
                long idx[3];
                p = parse_corner(p, counts, idx);
                if (!p) { ok = false; break; }
                auto key = std::make_tuple(idx[0], idx[1], idx[2]);
                auto it = indexOf.find(key);
                if (it == indexOf.end())
                {
                    VertexPNT v;
                    memset(&v, 0, sizeof(v));
                    v.position.x = pos[(size_t)idx[0] * 3 + 0];
                    v.position.y = pos[(size_t)idx[0] * 3 + 1];
                    v.position.z = pos[(size_t)idx[0] * 3 + 2];
                    if (idx[2] >= 0)
                    {
                        v.normal.x = nrm[(size_t)idx[2] * 3 + 0];
                        v.normal.y = nrm[(size_t)idx[2] * 3 + 1];
                        v.normal.z = nrm[(size_t)idx[2] * 3 + 2];
                    }
                    if (idx[1] >= 0)
                    {
                        v.textureCoord.x = tex[(size_t)idx[1] * 2 + 0];
                        v.textureCoord.y = tex[(size_t)idx[1] * 2 + 1];
                    }
                    it = indexOf.emplace(key, (u32)vertices.size()).first;
                    vertices.push_back(v);
                }
                corners.push_back(it->second);
            }
            if (ok && corners.size() < 3) ok = false;
            for (size_t k = 1; ok && k + 1 < corners.size(); ++k)
            {
                indices.push_back(corners[0]);
                indices.push_back(corners[k]);
                indices.push_back(corners[k + 1]);
            }
        }
    }
    fclose(f);
    if (!ok || indices.empty()) return 0;
    // caller frees with sp_b200_FreeMeshData (malloc/free, the convention of LoadExrImage,
    // src/asset_loader/asset_loader.h:11-22)
    out->vertices = (VertexPNT *)malloc(vertices.size() * sizeof(VertexPNT));
    out->indices = (u32 *)malloc(indices.size() * sizeof(u32));
    if (!out->vertices || !out->indices)
    {
        free(out->vertices);
        free(out->indices);
        memset(out, 0, sizeof(*out));
        return 0;
    }
    memcpy(out->vertices, vertices.data(), vertices.size() * sizeof(VertexPNT));
    memcpy(out->indices, indices.data(), indices.size() * sizeof(u32));
    out->vertexCount = (u32)vertices.size();
    out->indexCount = (u32)indices.size();
    return 1;
}

extern "C" void sp_b200_FreeMeshData(sp_b200_MeshData *mesh)
{
    if (!mesh) return;
    free(mesh->vertices);
    free(mesh->indices);
    memset(mesh, 0, sizeof(*mesh));
}

// =============================================================================================
// OpenEXR input behind the reference's own C ABI (src/asset_loader/asset_loader.h:11-22:
// `int LoadExrImage(HdrImage *image, const char *path)`, 0 on success, 1 on failure, pixels
// malloc'ed RGBA f32 that the caller free()s).  The reference implements it with the vendored
// tinyexr (thirdparty/tinyexr, LoadEXR); this is an independent reader written from the OpenEXR
// file-layout specification, for the part of the format an equirect environment map uses:
// single-part scanline images, channels R, G, B (+ A) or one luminance channel, HALF or FLOAT
// pixels, compression NONE / RLE / ZIPS / ZIP / PIZ, any line order, data window != display window.
// Output convention of LoadEXR: rows top to bottom over the data window, RGBA, alpha 1 when the
// file has none, a single channel replicated into all four.  Single-part tiled files are read too
// (level 0 of any level mode, like LoadEXR).  Multi-part, deep, UINT and PXR24 / B44 / DWA files
// are refused (return 1) -- tinyexr refuses those compressions as well -- never misread.
namespace {

struct Bytes
{
    const uint8_t *p = nullptr;
    size_t n = 0, at = 0;
    bool ok = true;
    uint8_t u8() { if (at + 1 > n) { ok = false; return 0; } return p[at++]; }
    uint32_t u32()
    {
        if (at + 4 > n) { ok = false; return 0; }
        uint32_t v = (uint32_t)p[at] | ((uint32_t)p[at + 1] << 8) | ((uint32_t)p[at + 2] << 16) | ((uint32_t)p[at + 3] << 24);
        at += 4;
        return v;
    }
    uint64_t u64() { uint64_t lo = u32(), hi = u32(); return lo | (hi << 32); }
    bool str(char *dst, size_t cap) // NUL-terminated, at most cap - 1 characters
    {
        size_t k = 0;
        for (;;)
        {
            uint8_t c = u8();
            if (!ok) return false;
            if (c == 0) break;
            if (k + 1 >= cap) { ok = false; return false; }
            dst[k++] = (char)c;
        }
        dst[k] = 0;
        return true;
    }
};

// ---- DEFLATE (RFC 1951) inside a zlib stream (RFC 1950); the Adler-32 trailer is not checked
struct Inflate
{
    const uint8_t *in;
    size_t inLen, inAt = 0;
    uint32_t bitBuf = 0;
    int bitCnt = 0;
    bool ok = true;
    int bits(int need)
    {
        uint32_t v = bitBuf;
        while (bitCnt < need)
        {
            if (inAt >= inLen) { ok = false; return 0; }
            v |= (uint32_t)in[inAt++] << bitCnt;
            bitCnt += 8;
        }
        bitBuf = need < 32 ? v >> need : 0;
        bitCnt -= need;
        return (int)(v & ((need < 32 ? (1u << need) : 0u) - 1u));
    }
    struct Huff { uint16_t count[16]; uint16_t symbol[288]; };
    static bool build(Huff &h, const uint8_t *lengths, int n)
    {
        memset(h.count, 0, sizeof(h.count));
        for (int i = 0; i < n; ++i) h.count[lengths[i]]++;
        if (h.count[0] == n) return true; // no codes
        int left = 1;
        for (int len = 1; len < 16; ++len)
        {
            left <<= 1;
            left -= h.count[len];
            if (left < 0) return false; // over-subscribed
        }
        uint16_t offs[16];
        offs[1] = 0;
        for (int len = 1; len < 15; ++len) offs[len + 1] = offs[len] + h.count[len];
        for (int i = 0; i < n; ++i)
            if (lengths[i]) h.symbol[offs[lengths[i]]++] = (uint16_t)i;
        return true;
    }
    int decode(const Huff &h)
    {
        int code = 0, first = 0, index = 0;
        for (int len = 1; len < 16; ++len)
        {
            code |= bits(1);
            if (!ok) return -1;
            int count = h.count[len];
            if (code - count < first) return h.symbol[index + (code - first)];
            index += count;
            first += count;
            first <<= 1;
            code <<= 1;
        }
        ok = false;
        return -1;
    }
    bool run(std::vector<uint8_t> &out, size_t expect)
    {
        static const uint16_t lenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
        static const uint16_t lenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
        static const uint16_t distBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
        static const uint16_t distExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
        out.clear();
        out.reserve(expect);
        if (inLen < 2 || (in[0] & 0x0F) != 8 || (((unsigned)in[0] << 8) | in[1]) % 31 != 0 || (in[1] & 0x20)) return false;
        inAt = 2;
        int last;
        do
        {
            last = bits(1);
            int type = bits(2);
            if (!ok) return false;
            if (type == 0)
            {
                bitBuf = 0; bitCnt = 0;
                if (inAt + 4 > inLen) return false;
                unsigned len = in[inAt] | (in[inAt + 1] << 8), nlen = in[inAt + 2] | (in[inAt + 3] << 8);
                inAt += 4;
                if ((len ^ 0xFFFFu) != nlen || inAt + len > inLen) return false;
                out.insert(out.end(), in + inAt, in + inAt + len);
                inAt += len;
            }
            else if (type == 1 || type == 2)
            {
                Huff lit, dist;
                uint8_t lengths[320];
                if (type == 1)
                {
                    int i = 0;
                    for (; i < 144; ++i) lengths[i] = 8;
                    for (; i < 256; ++i) lengths[i] = 9;
                    for (; i < 280; ++i) lengths[i] = 7;
                    for (; i < 288; ++i) lengths[i] = 8;
                    build(lit, lengths, 288);
                    for (i = 0; i < 30; ++i) lengths[i] = 5;
                    build(dist, lengths, 30);
                }
                else
                {
                    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                    int nlen = bits(5) + 257, ndist = bits(5) + 1, ncode = bits(4) + 4;
                    if (!ok || nlen > 286 || ndist > 30) return false;
                    uint8_t cl[19] = {0};
                    for (int i = 0; i < ncode; ++i) cl[order[i]] = (uint8_t)bits(3);
                    Huff lencode;
                    if (!ok || !build(lencode, cl, 19)) return false;
                    int idx = 0;
                    while (idx < nlen + ndist)
                    {
                        int sym = decode(lencode);
                        if (!ok) return false;
                        if (sym < 16) lengths[idx++] = (uint8_t)sym;
                        else
                        {
                            int prev = 0, rep;
                            if (sym == 16)
                            {
                                if (idx == 0) return false;
                                prev = lengths[idx - 1];
                                rep = 3 + bits(2);
                            }
                            else if (sym == 17) rep = 3 + bits(3);
                            else rep = 11 + bits(7);
                            if (!ok || idx + rep > nlen + ndist) return false;
                            while (rep--) lengths[idx++] = (uint8_t)prev;
                        }
                    }
                    if (lengths[256] == 0) return false;
                    if (!build(lit, lengths, nlen) || !build(dist, lengths + nlen, ndist)) return false;
                }
                for (;;)
                {
                    int sym = decode(lit);
                    if (!ok) return false;
                    if (sym < 256) out.push_back((uint8_t)sym);
                    else if (sym == 256) break;
                    else
                    {
                        sym -= 257;
                        if (sym >= 29) return false;
                        int len = lenBase[sym] + bits(lenExtra[sym]);
                        int ds = decode(dist);
                        if (!ok || ds < 0 || ds >= 30) return false;
                        size_t d = (size_t)distBase[ds] + (size_t)bits(distExtra[ds]);
                        if (!ok || d > out.size()) return false;
                        size_t from = out.size() - d;
                        for (int k = 0; k < len; ++k) out.push_back(out[from + k]);
                    }
                    if (out.size() > expect) return false;
                }
            }
            else return false;
        } while (!last);
        return ok;
    }
};

// the byte predictor and the even/odd interleave every ZIP / RLE chunk is stored with
void exr_unfilter(std::vector<uint8_t> &tmp, uint8_t *dst)
{
    const size_t n = tmp.size();
    for (size_t i = 1; i < n; ++i) tmp[i] = (uint8_t)(tmp[i - 1] + tmp[i] - 128);
    const size_t half = (n + 1) / 2;
    for (size_t k = 0; k < half; ++k)
    {
        dst[2 * k] = tmp[k];
        if (2 * k + 1 < n) dst[2 * k + 1] = tmp[half + k];
    }
}

float half_to_float(uint16_t h)
{
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16, expo = (h >> 10) & 0x1Fu, mant = h & 0x3FFu, bits;
    if (expo == 0)
    {
        if (mant == 0) bits = sign;
        else
        {
            int e = -1;
            do { e++; mant <<= 1; } while ((mant & 0x400u) == 0);
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((mant & 0x3FFu) << 13);
        }
    }
    else if (expo == 31) bits = sign | 0x7F800000u | (mant << 13);
    else bits = sign | ((expo + 112u) << 23) | (mant << 13);
    float f;
    memcpy(&f, &bits, 4);
    return f;
}


// ---- PIZ: a bitmap of the 16-bit values in use (-> lookup table), Huffman-coded 16-bit symbols,
// a 2-D Haar-like wavelet per channel plane.  Written from the published description of the
// format (OpenEXR technical introduction; layout constants as in ILM's reference codec).
struct PizBits
{
    const uint8_t *p, *end;
    uint64_t c = 0;
    int lc = 0;
    bool ok = true;
    uint64_t get(int n) // MSB first
    {
        while (lc < n)
        {
            if (p >= end) { ok = false; return 0; }
            c = (c << 8) | *p++;
            lc += 8;
        }
        lc -= n;
        return (c >> lc) & ((n < 64 ? (1ull << n) : 0ull) - 1ull);
    }
};

bool piz_huffman(const uint8_t *data, size_t size, std::vector<uint16_t> &out, size_t count)
{
    const uint32_t ENC = 65537;
    if (size < 20) return false;
    auto rd32 = [&](size_t at) { return (uint32_t)data[at] | ((uint32_t)data[at + 1] << 8) | ((uint32_t)data[at + 2] << 16) | ((uint32_t)data[at + 3] << 24); };
    uint32_t im = rd32(0), iM = rd32(4), nBits = rd32(12);
    if (im >= ENC || iM >= ENC || im > iM) return false;
    std::vector<uint8_t> len(ENC, 0);
    PizBits b;
    b.p = data + 20;
    b.end = data + size;
    // code lengths, 6 bits each; 59..62 = a run of 2..5 zero lengths, 63 = a run of 6 + next 8 bits
    for (uint32_t i = im; i <= iM; ++i)
    {
        uint32_t l = (uint32_t)b.get(6);
        if (!b.ok) return false;
        if (l == 63 || l >= 59)
        {
            uint32_t run = l == 63 ? (uint32_t)b.get(8) + 6 : l - 59 + 2;
            if (!b.ok || i + run > iM + 1) return false;
            i += run - 1; // the lengths stay 0
        }
        else len[i] = (uint8_t)l;
    }
    const uint8_t *bitsAt = b.p; // the code bits start at the next whole byte
    if ((size_t)(data + size - bitsAt) * 8 < nBits) return false;
    // canonical codes: within a length, codes count up in symbol order; base per length from the
    // longest length down
    uint64_t n[59] = {0};
    for (uint32_t i = 0; i < ENC; ++i) n[len[i]]++;
    uint64_t base[59], c = 0;
    for (int l = 58; l > 0; --l)
    {
        uint64_t nc = (c + n[l]) >> 1;
        base[l] = c;
        c = nc;
    }
    std::vector<uint32_t> firstIndex(60, 0), symbols;
    symbols.reserve(ENC);
    for (int l = 1; l <= 58; ++l)
    {
        firstIndex[l] = (uint32_t)symbols.size();
        if (!n[l]) continue;
        for (uint32_t i = im; i <= iM; ++i)
            if (len[i] == l) symbols.push_back(i);
    }
    firstIndex[59] = (uint32_t)symbols.size();
    PizBits d;
    d.p = bitsAt;
    d.end = data + size;
    out.clear();
    out.reserve(count);
    uint64_t used = 0;
    while (out.size() < count)
    {
        uint64_t code = 0;
        int l = 0;
        uint32_t sym = 0xFFFFFFFFu;
        while (l < 58)
        {
            code = (code << 1) | d.get(1);
            ++l;
            ++used;
            if (!d.ok || used > nBits) return false;
            if (n[l] && code >= base[l] && code - base[l] < n[l])
            {
                sym = symbols[firstIndex[l] + (uint32_t)(code - base[l])];
                break;
            }
        }
        if (sym == 0xFFFFFFFFu) return false;
        if (sym == iM)
        {
            // run-length symbol: repeat the previous value
            uint32_t run = (uint32_t)d.get(8);
            used += 8;
            if (!d.ok || used > nBits || out.empty() || out.size() + run > count) return false;
            out.insert(out.end(), run, out.back());
        }
        else out.push_back((uint16_t)sym);
    }
    return true;
}

inline void piz_wdec14(uint16_t l, uint16_t h, uint16_t &a, uint16_t &b)
{
    int16_t ls = (int16_t)l, hs = (int16_t)h;
    int hi = hs, ai = ls + (hi & 1) + (hi >> 1);
    a = (uint16_t)(int16_t)ai;
    b = (uint16_t)(int16_t)(ai - hi);
}
inline void piz_wdec16(uint16_t l, uint16_t h, uint16_t &a, uint16_t &b)
{
    int m = l, d = h;
    int bb = (m - (d >> 1)) & 0xFFFF;
    int aa = (d + bb - 0x8000) & 0xFFFF;
    b = (uint16_t)bb;
    a = (uint16_t)aa;
}

void piz_wavelet_decode(uint16_t *in, int nx, int ox, int ny, int oy, uint16_t mx)
{
    const bool w14 = mx < (1 << 14);
    int n = nx > ny ? ny : nx, p = 1, p2;
    while (p <= n) p <<= 1;
    p >>= 1;
    p2 = p;
    p >>= 1;
    auto dec = [&](uint16_t l, uint16_t h, uint16_t &a, uint16_t &b) { if (w14) piz_wdec14(l, h, a, b); else piz_wdec16(l, h, a, b); };
    while (p >= 1)
    {
        uint16_t *py = in, *ey = in + (ptrdiff_t)oy * (ny - p2);
        const ptrdiff_t oy1 = (ptrdiff_t)oy * p, oy2 = (ptrdiff_t)oy * p2, ox1 = (ptrdiff_t)ox * p, ox2 = (ptrdiff_t)ox * p2;
        uint16_t i00, i01, i10, i11;
        for (; py <= ey; py += oy2)
        {
            uint16_t *px = py, *ex = py + (ptrdiff_t)ox * (nx - p2);
            for (; px <= ex; px += ox2)
            {
                uint16_t *p01 = px + ox1, *p10 = px + oy1, *p11 = p10 + ox1;
                dec(*px, *p10, i00, i10);
                dec(*p01, *p11, i01, i11);
                dec(i00, i01, *px, *p01);
                dec(i10, i11, *p10, *p11);
            }
            if (nx & p)
            {
                uint16_t *p10 = px + oy1;
                dec(*px, *p10, i00, *p10);
                *px = i00;
            }
        }
        if (ny & p)
        {
            uint16_t *px = py, *ex = py + (ptrdiff_t)ox * (nx - p2);
            for (; px <= ex; px += ox2)
            {
                uint16_t *p01 = px + ox1;
                dec(*px, *p01, i00, *p01);
                *px = i00;
            }
        }
        p2 = p;
        p >>= 1;
    }
}

// one PIZ chunk -> the raw scanline bytes (lines x [channel rows]); sizes[c] = 1 (HALF) or 2 (FLOAT)
bool piz_decode(const uint8_t *src, size_t size, uint8_t *raw, int width, int lines, const std::vector<int> &sizes)
{
    size_t total = 0;
    for (int sz : sizes) total += (size_t)width * lines * sz;
    if (size < 4) return false;
    uint16_t minNZ = (uint16_t)(src[0] | (src[1] << 8)), maxNZ = (uint16_t)(src[2] | (src[3] << 8));
    size_t at = 4;
    std::vector<uint8_t> bitmap(8192, 0);
    if (minNZ <= maxNZ)
    {
        if (maxNZ >= 8192 || at + (size_t)(maxNZ - minNZ + 1) > size) return false;
        memcpy(bitmap.data() + minNZ, src + at, (size_t)(maxNZ - minNZ + 1));
        at += (size_t)(maxNZ - minNZ + 1);
    }
    std::vector<uint16_t> lut(65536, 0);
    uint32_t k = 0;
    for (uint32_t i = 0; i < 65536; ++i)
        if (i == 0 || (bitmap[i >> 3] & (1u << (i & 7)))) lut[k++] = (uint16_t)i;
    const uint16_t maxValue = (uint16_t)(k - 1);
    if (at + 4 > size) return false;
    uint32_t length = (uint32_t)src[at] | ((uint32_t)src[at + 1] << 8) | ((uint32_t)src[at + 2] << 16) | ((uint32_t)src[at + 3] << 24);
    at += 4;
    if (at + length > size) return false;
    std::vector<uint16_t> tmp;
    if (!piz_huffman(src + at, length, tmp, total) || tmp.size() != total) return false;
    size_t start = 0;
    std::vector<size_t> starts;
    for (int sz : sizes)
    {
        starts.push_back(start);
        for (int j = 0; j < sz; ++j) piz_wavelet_decode(tmp.data() + start + j, width, sz, lines, width * sz, maxValue);
        start += (size_t)width * lines * sz;
    }
    for (uint16_t &v : tmp) v = lut[v];
    // channel-planar -> scanline-major
    uint8_t *dst = raw;
    std::vector<size_t> cursor = starts;
    for (int l = 0; l < lines; ++l)
        for (size_t c = 0; c < sizes.size(); ++c)
        {
            size_t nvals = (size_t)width * sizes[c];
            for (size_t i = 0; i < nvals; ++i)
            {
                uint16_t v = tmp[cursor[c] + i];
                dst[2 * i] = (uint8_t)(v & 0xFF);
                dst[2 * i + 1] = (uint8_t)(v >> 8);
            }
            cursor[c] += nvals;
            dst += nvals * 2;
        }
    return true;
}

} // namespace

static int load_exr_image(HdrImage *image, const char *path);

extern "C" int LoadExrImage(HdrImage *image, const char *path)
{
    // a hostile header can ask for more memory than there is: no exception leaves the C ABI
    try
    {
        return load_exr_image(image, path);
    }
    catch (...)
    {
        if (image)
        {
            image->pixels = nullptr;
            image->width = image->height = 0;
        }
        return 1;
    }
}

static int load_exr_image(HdrImage *image, const char *path)
{
    if (!image) return 1;
    image->pixels = nullptr;
    image->width = image->height = 0;
    FILE *f = path ? fopen(path, "rb") : nullptr;
    if (!f) return 1;
    std::vector<uint8_t> file;
    {
        uint8_t buf[1 << 16];
        size_t got;
        while ((got = fread(buf, 1, sizeof(buf), f)) > 0) file.insert(file.end(), buf, buf + got);
        fclose(f);
    }
    Bytes b;
    b.p = file.data();
    b.n = file.size();
    if (b.u32() != 20000630u) return 1;           // magic 76 2f 31 01
    uint32_t version = b.u32();
    if ((version & 0xFFu) != 2 || (version & 0x1800u)) return 1; // deep / multi-part: refused
    const bool tiled = (version & 0x200u) != 0;                  // single-part tiled file

    struct Channel { char name[64]; uint32_t type; };
    std::vector<Channel> channels;
    int compression = -1, lineOrder = 0;
    uint32_t tileW = 0, tileH = 0, tileMode = 0;
    int32_t dw[4] = {0, 0, -1, -1};
    for (;;)
    {
        char name[256], type[64];
        if (!b.str(name, sizeof(name))) return 1;
        if (name[0] == 0) break;
        if (!b.str(type, sizeof(type))) return 1;
        uint32_t size = b.u32();
        if (!b.ok || b.at + size > b.n) return 1;
        Bytes a;
        a.p = b.p + b.at;
        a.n = size;
        b.at += size;
        if (!strcmp(name, "channels") && !strcmp(type, "chlist"))
        {
            for (;;)
            {
                Channel c;
                if (!a.str(c.name, sizeof(c.name))) return 1;
                if (c.name[0] == 0) break;
                c.type = a.u32();
                a.u8(); a.u8(); a.u8(); a.u8();       // pLinear + reserved
                uint32_t xs = a.u32(), ys = a.u32();
                if (!a.ok || xs != 1 || ys != 1) return 1; // sub-sampled channels: refused
                channels.push_back(c);
            }
        }
        else if (!strcmp(name, "compression")) compression = a.u8();
        else if (!strcmp(name, "dataWindow"))
            for (int k = 0; k < 4; ++k) dw[k] = (int32_t)a.u32();
        else if (!strcmp(name, "lineOrder")) lineOrder = a.u8();
        else if (!strcmp(name, "tiles") && !strcmp(type, "tiledesc"))
        {
            tileW = a.u32();
            tileH = a.u32();
            tileMode = a.u8();
        }
        if (!a.ok) return 1;
    }
    (void)lineOrder; // chunks carry their own y
    if (channels.empty() || dw[2] < dw[0] || dw[3] < dw[1]) return 1;
    const int64_t width = (int64_t)dw[2] - dw[0] + 1, height = (int64_t)dw[3] - dw[1] + 1;
    if (width <= 0 || height <= 0 || width > 65536 || height > 65536) return 1;
    int linesPerBlock;
    switch (compression)
    {
    case 0: case 1: case 2: linesPerBlock = 1; break;  // NONE, RLE, ZIPS
    case 3: linesPerBlock = 16; break;                 // ZIP
    case 4: linesPerBlock = 32; break;                 // PIZ
    default: return 1;                                 // PXR24, B44, DWA: refused
    }
    // channel roles (LoadEXR: one channel -> grey; else R, G, B required, A optional)
    int idx[4] = {-1, -1, -1, -1};
    size_t bytesPerPixel = 0;            // all channels of one pixel
    for (size_t c = 0; c < channels.size(); ++c)
    {
        if (channels[c].type != 1 && channels[c].type != 2) return 1; // UINT: refused
        bytesPerPixel += channels[c].type == 1 ? 2 : 4;
        const char *n = channels[c].name;
        if (!strcmp(n, "R")) idx[0] = (int)c;
        else if (!strcmp(n, "G")) idx[1] = (int)c;
        else if (!strcmp(n, "B")) idx[2] = (int)c;
        else if (!strcmp(n, "A")) idx[3] = (int)c;
    }
    const bool grey = channels.size() == 1;
    if (!grey && (idx[0] < 0 || idx[1] < 0 || idx[2] < 0)) return 1;

    // Blocks.  Scanline files: `linesPerBlock` full-width rows per chunk, chunk = {y, size, data}.
    // Tiled files: tileW x tileH rectangles (clipped at the right / bottom edge), chunk = {tileX,
    // tileY, levelX, levelY, size, data}; only level (0, 0) is the image (LoadEXR reads nothing
    // else), and its tiles are the first nx * ny entries of the offset table in every level mode.
    size_t tilesX = 0, tilesY = 0;
    if (tiled)
    {
        if (tileW == 0 || tileH == 0 || tileW > 65536 || tileH > 65536) return 1;
        if ((tileMode & 0xFu) > 2) return 1;
        tilesX = (size_t)((width + tileW - 1) / tileW);
        tilesY = (size_t)((height + tileH - 1) / tileH);
    }
    const size_t chunks = tiled ? tilesX * tilesY : (size_t)((height + linesPerBlock - 1) / linesPerBlock);
    // The header alone must not size anything: the offset table has to fit in what is left of the file, and the
    // image in what a file of this size could hold at all (RLE, the densest coding read here, expands a run of
    // 2 bytes to at most 128; a crafted 65536 x 65536 header would otherwise ask for 64 GiB before a byte is read).
    if (b.at > file.size() || chunks > (file.size() - b.at) / 8) return 1;
    if ((uint64_t)width * (uint64_t)height * bytesPerPixel / 64 > (uint64_t)file.size()) return 1;
    std::vector<uint64_t> offsets(chunks);
    for (size_t i = 0; i < chunks; ++i) offsets[i] = b.u64();
    if (!b.ok) return 1;
    float *out = (float *)calloc((size_t)width * (size_t)height * 4, sizeof(float)); // free()d by the caller
    if (!out) return 1;
    std::vector<std::atomic<uint8_t>> claimed(chunks); // one owner per block of rows / tile
    for (auto &c : claimed) c.store(0);
    // chunks are independent (own pixels of the output): decoded by up to 8 threads
    auto decode_chunk = [&](size_t i, std::vector<uint8_t> &raw, std::vector<uint8_t> &tmp) -> bool
    {
        bool ok = true;
        Bytes c;
        c.p = file.data();
        c.n = file.size();
        c.at = (size_t)offsets[i];
        if (offsets[i] >= file.size()) return false;
        // the block's rectangle inside the data window: columns [bx, bx + bw), rows [by, by + lines)
        int64_t bx = 0, by, bw = width, lines;
        if (tiled)
        {
            uint32_t tx = c.u32(), ty = c.u32(), lx = c.u32(), ly = c.u32();
            if (!c.ok || lx != 0 || ly != 0 || tx >= tilesX || ty >= tilesY) return false;
            if (claimed[(size_t)ty * tilesX + tx].exchange(1) != 0) return false;
            bx = (int64_t)tx * tileW;
            by = (int64_t)ty * tileH;
            bw = std::min<int64_t>(tileW, width - bx);
            lines = std::min<int64_t>(tileH, height - by);
        }
        else
        {
            int32_t y = (int32_t)c.u32();
            if (!c.ok || y < dw[1] || y > dw[3]) return false;
            by = (int64_t)y - dw[1];
            // (blocks start on multiples of linesPerBlock: two chunks of a crafted file must not cover the same
            // rows -- the decoding threads own disjoint rows of `out`)
            if (by % linesPerBlock != 0 || claimed[(size_t)(by / linesPerBlock)].exchange(1) != 0) return false;
            lines = std::min<int64_t>(linesPerBlock, height - by);
        }
        uint32_t dataSize = c.u32();
        if (!c.ok || c.at + dataSize > c.n) return false;
        const size_t bytesPerPixelRow = bytesPerPixel * (size_t)bw;
        const size_t rawSize = bytesPerPixelRow * (size_t)lines;
        raw.resize(rawSize);
        const uint8_t *src = c.p + c.at;
        if (compression == 0 || dataSize >= rawSize)
        {
            // stored: NONE, or a chunk the writer could not shrink
            if (dataSize != rawSize) return false;
            memcpy(raw.data(), src, rawSize);
        }
        else if (compression == 1)
        {
            // RLE: a count byte n >= 0 repeats the next byte n + 1 times, n < 0 copies -n bytes
            tmp.clear();
            size_t at = 0;
            while (at < dataSize && ok)
            {
                int n = (int8_t)src[at++];
                if (n < 0)
                {
                    size_t len = (size_t)(-n);
                    if (at + len > dataSize) return false;
                    tmp.insert(tmp.end(), src + at, src + at + len);
                    at += len;
                }
                else
                {
                    if (at >= dataSize) return false;
                    tmp.insert(tmp.end(), (size_t)n + 1, src[at++]);
                }
                if (tmp.size() > rawSize) return false;
            }
            if (!ok || tmp.size() != rawSize) return false;
            exr_unfilter(tmp, raw.data());
        }
        else if (compression == 4)
        {
            std::vector<int> sizes;
            for (const Channel &ch : channels) sizes.push_back(ch.type == 1 ? 1 : 2);
            if (!piz_decode(src, dataSize, raw.data(), (int)bw, (int)lines, sizes)) return false;
        }
        else
        {
            Inflate z;
            z.in = src;
            z.inLen = dataSize;
            if (!z.run(tmp, rawSize) || tmp.size() != rawSize) return false;
            exr_unfilter(tmp, raw.data());
        }
        // within a row the channels follow each other (file order), bw samples each
        std::vector<size_t> chOffset(channels.size());
        {
            size_t at = 0;
            for (size_t k = 0; k < channels.size(); ++k)
            {
                chOffset[k] = at;
                at += (size_t)bw * (channels[k].type == 1 ? 2 : 4);
            }
        }
        for (int64_t l = 0; l < lines; ++l)
        {
            const uint8_t *row = raw.data() + bytesPerPixelRow * (size_t)l;
            float *dst = out + ((size_t)(by + l) * (size_t)width + (size_t)bx) * 4;
            auto sample = [&](int ch, int64_t x) -> float {
                const uint8_t *p = row + chOffset[(size_t)ch];
                if (channels[(size_t)ch].type == 1)
                {
                    uint16_t h = (uint16_t)(p[2 * x] | (p[2 * x + 1] << 8));
                    return half_to_float(h);
                }
                uint32_t u = (uint32_t)p[4 * x] | ((uint32_t)p[4 * x + 1] << 8) | ((uint32_t)p[4 * x + 2] << 16) | ((uint32_t)p[4 * x + 3] << 24);
                float v;
                memcpy(&v, &u, 4);
                return v;
            };
            for (int64_t x = 0; x < bw; ++x)
            {
                if (grey)
                {
                    float v = sample(0, x);
                    dst[4 * x + 0] = dst[4 * x + 1] = dst[4 * x + 2] = dst[4 * x + 3] = v;
                }
                else
                {
                    dst[4 * x + 0] = sample(idx[0], x);
                    dst[4 * x + 1] = sample(idx[1], x);
                    dst[4 * x + 2] = sample(idx[2], x);
                    dst[4 * x + 3] = idx[3] >= 0 ? sample(idx[3], x) : 1.0f;
                }
            }
        }
        return ok;
    };
    std::atomic<size_t> next(0);
    std::atomic<bool> failed(false);
    auto worker = [&]() {
        std::vector<uint8_t> raw, tmp;
        for (;;)
        {
            size_t i = next.fetch_add(1);
            if (i >= chunks || failed.load()) break;
            bool good = false;
            try { good = decode_chunk(i, raw, tmp); } catch (...) { good = false; }
            if (!good) failed.store(true);
        }
    };
    unsigned threads = std::thread::hardware_concurrency();
    threads = threads < 1 ? 1 : (threads > 8 ? 8 : threads);
    if (chunks < 8) threads = 1;
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < threads; ++t) pool.emplace_back(worker);
    worker();
    for (std::thread &t : pool) t.join();
    const bool ok = !failed.load();
    if (!ok)
    {
        free(out);
        return 1;
    }
    image->pixels = out;
    image->width = (uint32_t)width;
    image->height = (uint32_t)height;
    return 0;
}
