// PNG decode / encode on zlib (no libpng in the image), enough for the reference's I/O surface:
// imread(path) -> 8-bit BGR (alpha dropped, grey / palette expanded; NCT/main.cu:483,491) and
// imwrite(path, bgr) (NCT/main.cu:538).  Non-interlaced and Adam7-free files only, bit depths 8 and 16
// (16 is reduced to the high byte), colour types 0, 2, 3, 4, 6.
#include "nct_internal.h"
#include <zlib.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
void put32(std::vector<uint8_t> &v, uint32_t x)
{
    v.push_back((uint8_t)(x >> 24));
    v.push_back((uint8_t)(x >> 16));
    v.push_back((uint8_t)(x >> 8));
    v.push_back((uint8_t)x);
}

int paeth(int a, int b, int c)
{
    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    if (pa <= pb && pa <= pc) return a;
    return pb <= pc ? b : c;
}

void write_chunk(std::vector<uint8_t> &out, const char *type, const uint8_t *data, size_t n)
{
    put32(out, (uint32_t)n);
    const size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    if (n) out.insert(out.end(), data, data + n);
    put32(out, (uint32_t)crc32(0L, out.data() + start, (uInt)(n + 4)));
}

}  // namespace

extern "C" {

int nct_png_read(const char *path, uint8_t **bgr_out, int *h_out, int *w_out)
{
    if (!path || !bgr_out || !h_out || !w_out) return NCT_ERR_ARG;
    *bgr_out = nullptr;
    FILE *fp = fopen(path, "rb");
    if (!fp) return NCT_ERR_IO;
    fseek(fp, 0, SEEK_END);
    const long sz = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    std::vector<uint8_t> buf((size_t)(sz > 0 ? sz : 0));
    const size_t got = sz > 0 ? fread(buf.data(), 1, buf.size(), fp) : 0;
    fclose(fp);
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (got < 8 + 25 || memcmp(buf.data(), sig, 8) != 0) return NCT_ERR_IO;
    size_t pos = 8;
    uint32_t W = 0, H = 0;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte;
    bool have_ihdr = false, done = false;
    while (!done && pos + 12 <= buf.size()) {
        const uint32_t len = be32(&buf[pos]);
        const uint8_t *type = &buf[pos + 4];
        if (pos + 12 + (size_t)len > buf.size()) return NCT_ERR_IO;
        const uint8_t *data = &buf[pos + 8];
        if (!memcmp(type, "IHDR", 4) && len >= 13) {
            W = be32(data);
            H = be32(data + 4);
            depth = data[8];
            ctype = data[9];
            interlace = data[12];
            have_ihdr = true;
        } else if (!memcmp(type, "PLTE", 4)) {
            plte.assign(data, data + len);
        } else if (!memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!memcmp(type, "IEND", 4)) {
            done = true;
        }
        pos += 12 + (size_t)len;
    }
    if (!have_ihdr || W == 0 || H == 0 || W > 32768 || H > 32768 || interlace != 0) return NCT_ERR_IO;
    if (!(depth == 8 || depth == 16) && !((ctype == 3 || ctype == 0) && (depth == 1 || depth == 2 || depth == 4))) return NCT_ERR_IO;
    int channels;
    switch (ctype) {
    case 0: channels = 1; break;
    case 2: channels = 3; break;
    case 3: channels = 1; break;
    case 4: channels = 2; break;
    case 6: channels = 4; break;
    default: return NCT_ERR_IO;
    }
    const size_t bpp_bits = (size_t)channels * depth;
    const size_t bpp = (bpp_bits + 7) / 8;                 // bytes per complete pixel (>= 1) for filtering
    const size_t stride = ((size_t)W * bpp_bits + 7) / 8;
    std::vector<uint8_t> raw((stride + 1) * (size_t)H);
    uLongf rawlen = (uLongf)raw.size();
    if (uncompress(raw.data(), &rawlen, idat.data(), (uLong)idat.size()) != Z_OK || rawlen != raw.size()) return NCT_ERR_IO;
    // undo the scanline filters in place
    std::vector<uint8_t> prev(stride, 0);
    for (uint32_t y = 0; y < H; ++y) {
        uint8_t *line = &raw[(stride + 1) * (size_t)y];
        const int ft = line[0];
        uint8_t *cur = line + 1;
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
            int v = cur[i];
            switch (ft) {
            case 0: break;
            case 1: v += a; break;
            case 2: v += b; break;
            case 3: v += (a + b) >> 1; break;
            case 4: v += paeth(a, b, c); break;
            default: return NCT_ERR_IO;
            }
            cur[i] = (uint8_t)v;
        }
        memcpy(prev.data(), cur, stride);
    }
    uint8_t *out = (uint8_t *)malloc((size_t)W * H * 3);
    if (!out) return NCT_ERR_NOMEM;
    const size_t step = depth == 16 ? 2 : 1;  // 16-bit samples: keep the high byte
    for (uint32_t y = 0; y < H; ++y) {
        const uint8_t *cur = &raw[(stride + 1) * (size_t)y + 1];
        uint8_t *o = out + (size_t)y * W * 3;
        for (uint32_t x = 0; x < W; ++x) {
            uint8_t r, g, b;
            if (ctype == 3) {
                int idx;
                if (depth == 8) idx = cur[x];
                else {
                    const int per = 8 / depth;
                    idx = (cur[x / per] >> ((per - 1 - x % per) * depth)) & ((1 << depth) - 1);
                }
                if ((size_t)idx * 3 + 2 < plte.size()) { r = plte[idx * 3]; g = plte[idx * 3 + 1]; b = plte[idx * 3 + 2]; }
                else r = g = b = 0;
            } else if (ctype == 0 && depth < 8) {   // packed greyscale samples, scaled to 8 bits
                const int per = 8 / depth, maxv = (1 << depth) - 1;
                const int v = (cur[x / per] >> ((per - 1 - x % per) * depth)) & maxv;
                r = g = b = (uint8_t)(v * 255 / maxv);
            } else if (channels <= 2) {
                r = g = b = cur[(size_t)x * channels * step];
            } else {
                const uint8_t *px = cur + (size_t)x * channels * step;
                r = px[0]; g = px[step]; b = px[2 * step];
            }
            o[x * 3] = b; o[x * 3 + 1] = g; o[x * 3 + 2] = r;
        }
    }
    *bgr_out = out;
    *h_out = (int)H;
    *w_out = (int)W;
    return NCT_OK;
}

void nct_png_free(uint8_t *p) { free(p); }

int nct_png_write(const char *path, const uint8_t *bgr, int h, int w)
{
    if (!path || !bgr || h <= 0 || w <= 0) return NCT_ERR_ARG;
    const size_t stride = (size_t)w * 3;
    std::vector<uint8_t> raw((stride + 1) * (size_t)h);
    for (int y = 0; y < h; ++y) {
        uint8_t *line = &raw[(stride + 1) * (size_t)y];
        const uint8_t *src = bgr + (size_t)y * stride;
        const uint8_t *up = y > 0 ? bgr + (size_t)(y - 1) * stride : nullptr;
        line[0] = up ? 2 : 0;  // "Up" filter compresses smooth images well; first row unfiltered
        for (int x = 0; x < w; ++x)
            for (int c = 0; c < 3; ++c) {
                const int v = src[x * 3 + (2 - c)];                 // BGR -> RGB
                const int pv = up ? up[x * 3 + (2 - c)] : 0;
                line[1 + x * 3 + c] = (uint8_t)(v - pv);
            }
    }
    uLongf clen = compressBound((uLong)raw.size());
    std::vector<uint8_t> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return NCT_ERR_IO;
    std::vector<uint8_t> out;
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    out.insert(out.end(), sig, sig + 8);
    uint8_t ihdr[13];
    ihdr[0] = (uint8_t)(w >> 24); ihdr[1] = (uint8_t)(w >> 16); ihdr[2] = (uint8_t)(w >> 8); ihdr[3] = (uint8_t)w;
    ihdr[4] = (uint8_t)(h >> 24); ihdr[5] = (uint8_t)(h >> 16); ihdr[6] = (uint8_t)(h >> 8); ihdr[7] = (uint8_t)h;
    ihdr[8] = 8; ihdr[9] = 2; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;
    write_chunk(out, "IHDR", ihdr, 13);
    write_chunk(out, "IDAT", comp.data(), clen);
    write_chunk(out, "IEND", nullptr, 0);
    FILE *fp = fopen(path, "wb");
    if (!fp) return NCT_ERR_IO;
    const size_t wr = fwrite(out.data(), 1, out.size(), fp);
    fclose(fp);
    return wr == out.size() ? NCT_OK : NCT_ERR_IO;
}

}  // extern "C"
