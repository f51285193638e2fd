// Minimal .caffemodel reader: pulls the conv1_1 .. conv5_1 weight / bias blobs out of a serialized
// caffe::NetParameter without linking protobuf.
//
// Replaces Net::CopyTrainedLayersFrom / ReadNetParamsFromBinaryFileOrDie (caffe/net.cpp:798-815,
// caffe/util/upgrade_proto.cpp) for the one network the reference loads (NCT/main.cu:575-582).
// Wire format facts used (caffe/proto/caffe.proto):
//   NetParameter    : layers = 2 (V1LayerParameter, legacy files such as VGG_ILSVRC_19_layers.caffemodel),
//                     layer  = 100 (LayerParameter)
//   V1LayerParameter: name = 4, blobs = 6            LayerParameter: name = 1, blobs = 7
//   BlobProto       : num/channels/height/width = 1..4, data = 5 (repeated float, packed or not),
//                     shape = 7 { dim = 1 (repeated int64, packed or not) }
#include "nct_internal.h"
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <map>
#include <memory>
#include <mutex>
#include <sys/stat.h>

namespace {

struct Reader {
    const uint8_t *p, *end;
    bool ok = true;
    bool eof() const { return p >= end; }
    uint64_t varint()
    {
        uint64_t v = 0;
        int shift = 0;
        while (p < end && shift < 64) {
            const uint8_t b = *p++;
            v |= (uint64_t)(b & 0x7f) << shift;
            if (!(b & 0x80)) return v;
            shift += 7;
        }
        ok = false;
        return 0;
    }
    // reads a tag; returns false at the end
    bool tag(uint32_t &field, uint32_t &wire)
    {
        if (eof()) return false;
        const uint64_t t = varint();
        field = (uint32_t)(t >> 3);
        wire = (uint32_t)(t & 7);
        return ok;
    }
    Reader sub()
    {
        const uint64_t n = varint();
        Reader r{p, p + n};
        if (!ok || n > (uint64_t)(end - p)) { ok = false; r.end = r.p; r.ok = false; return r; }
        p += n;
        return r;
    }
    void skip(uint32_t wire)
    {
        switch (wire) {
        case 0: varint(); break;
        case 1: p += 8; break;
        case 2: { const uint64_t n = varint(); if (n > (uint64_t)(end - p)) ok = false; else p += n; break; }
        case 5: p += 4; break;
        default: ok = false;
        }
        if (p > end) ok = false;
    }
};

struct Blob {
    std::vector<int64_t> shape;   // from BlobShape, or {num, channels, height, width}
    std::vector<float> data;
};

bool parse_blob(Reader r, Blob &b)
{
    int64_t legacy[4] = {0, 0, 0, 0};
    bool has_legacy = false;
    uint32_t f, w;
    while (r.tag(f, w)) {
        if (f >= 1 && f <= 4 && w == 0) {
            legacy[f - 1] = (int64_t)r.varint();
            has_legacy = true;
        } else if (f == 5 && w == 2) {  // packed floats
            Reader s = r.sub();
            if (!r.ok) return false;
            const size_t n = (size_t)(s.end - s.p) / 4;
            const size_t old = b.data.size();
            b.data.resize(old + n);
            memcpy(b.data.data() + old, s.p, n * 4);
        } else if (f == 5 && w == 5) {  // unpacked float
            float v;
            if (r.end - r.p < 4) return false;
            memcpy(&v, r.p, 4);
            r.p += 4;
            b.data.push_back(v);
        } else if (f == 7 && w == 2) {  // BlobShape
            Reader s = r.sub();
            if (!r.ok) return false;
            uint32_t sf, sw;
            while (s.tag(sf, sw)) {
                if (sf == 1 && sw == 0) b.shape.push_back((int64_t)s.varint());
                else if (sf == 1 && sw == 2) {
                    Reader d = s.sub();
                    while (!d.eof() && d.ok) b.shape.push_back((int64_t)d.varint());
                } else s.skip(sw);
                if (!s.ok) return false;
            }
        } else {
            r.skip(w);
        }
        if (!r.ok) return false;
    }
    if (b.shape.empty() && has_legacy) b.shape.assign(legacy, legacy + 4);
    return r.ok;
}

struct Layer {
    std::string name;
    std::vector<Blob> blobs;
};

bool parse_layer(Reader r, bool v1, Layer &L)
{
    const uint32_t f_name = v1 ? 4 : 1, f_blobs = v1 ? 6 : 7;
    uint32_t f, w;
    while (r.tag(f, w)) {
        if (f == f_name && w == 2) {
            Reader s = r.sub();
            if (!r.ok) return false;
            L.name.assign((const char *)s.p, (size_t)(s.end - s.p));
        } else if (f == f_blobs && w == 2) {
            Reader s = r.sub();
            if (!r.ok) return false;
            L.blobs.emplace_back();
            if (!parse_blob(s, L.blobs.back())) return false;
        } else {
            r.skip(w);
        }
        if (!r.ok) return false;
    }
    return r.ok;
}

}  // namespace

extern "C" {

// parsed trunk weights of one file: layer i -> (weights O x I x 3 x 3, bias O)
struct ParsedModel {
    std::vector<std::vector<float>> W, B;
};

// The ~550 MB file is read and parsed ONCE per process: the CLI creates ngpu x inflight contexts that all load the same
// model (every later context only uploads).  Keyed by path, size and modification time.
static int parse_caffemodel_cached(nct_ctx *ctx, const char *path, std::shared_ptr<const ParsedModel> &out)
{
    static std::mutex mu;
    static std::map<std::string, std::shared_ptr<const ParsedModel>> cache;
    struct stat st;
    if (stat(path, &st) != 0) return nct_fail(ctx, NCT_ERR_IO, "cannot open caffemodel '%s'", path);
    char keybuf[64];
    snprintf(keybuf, sizeof(keybuf), "|%lld|%lld", (long long)st.st_size, (long long)st.st_mtime);
    const std::string key = std::string(path) + keybuf;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { out = it->second; return NCT_OK; }

    FILE *fp = fopen(path, "rb");
    if (!fp) return nct_fail(ctx, NCT_ERR_IO, "cannot open caffemodel '%s'", path);
    fseek(fp, 0, SEEK_END);
    const long sz = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    std::vector<uint8_t> buf((size_t)(sz > 0 ? sz : 0));
    const size_t got = sz > 0 ? fread(buf.data(), 1, buf.size(), fp) : 0;
    fclose(fp);
    if (sz <= 0 || got != buf.size()) return nct_fail(ctx, NCT_ERR_IO, "cannot read caffemodel '%s'", path);

    const int nlayers = nct_vgg19_num_layers();
    auto model = std::make_shared<ParsedModel>();
    model->W.resize((size_t)nlayers);
    model->B.resize((size_t)nlayers);
    Reader r{buf.data(), buf.data() + buf.size()};
    uint32_t f, w;
    while (r.tag(f, w)) {
        if ((f == 2 || f == 100) && w == 2) {
            Reader s = r.sub();
            if (!r.ok) break;
            Layer L;
            if (!parse_layer(s, f == 2, L)) return nct_fail(ctx, NCT_ERR_IO, "'%s': malformed layer record", path);
            for (int i = 0; i < nlayers; ++i) {
                if (L.name != nct_vgg19_layer_name(i) || L.blobs.empty()) continue;
                int cin = 0, cout = 0;
                nct_vgg19_layer_shape(i, &cin, &cout);
                if (L.blobs.size() < 2) return nct_fail(ctx, NCT_ERR_IO, "layer %s has %zu blobs, expected weights + bias", L.name.c_str(), L.blobs.size());
                Blob &W = L.blobs[0], &B = L.blobs[1];
                if (W.data.size() != (size_t)cout * cin * 9 || B.data.size() != (size_t)cout)
                    return nct_fail(ctx, NCT_ERR_IO, "layer %s: blob sizes %zu / %zu do not match %d x %d x 3 x 3 / %d", L.name.c_str(),
                                    W.data.size(), B.data.size(), cout, cin, cout);
                if (W.shape.size() == 4 && (W.shape[0] != cout || W.shape[1] != cin || W.shape[2] != 3 || W.shape[3] != 3))
                    return nct_fail(ctx, NCT_ERR_IO, "layer %s: unexpected weight shape", L.name.c_str());
                model->W[(size_t)i] = std::move(W.data);
                model->B[(size_t)i] = std::move(B.data);
            }
        } else {
            r.skip(w);
        }
        if (!r.ok) break;
    }
    if (!r.ok) return nct_fail(ctx, NCT_ERR_IO, "'%s' is not a valid serialized NetParameter", path);
    for (int i = 0; i < nlayers; ++i)
        if (model->W[(size_t)i].empty()) return nct_fail(ctx, NCT_ERR_IO, "'%s' has no weights for layer %s", path, nct_vgg19_layer_name(i));
    cache[key] = model;
    out = model;
    return NCT_OK;
}

int nct_vgg19_load_caffemodel(nct_ctx *ctx, const char *path)
{
    if (!ctx || !path) return NCT_ERR_ARG;
    cudaSetDevice(ctx->device);
    std::shared_ptr<const ParsedModel> model;
    int rc = parse_caffemodel_cached(ctx, path, model);
    if (rc) return rc;
    for (int i = 0; i < nct_vgg19_num_layers(); ++i) {
        rc = nct_vgg19_set_weights(ctx, i, model->W[(size_t)i].data(), model->B[(size_t)i].data());
        if (rc) return rc;
    }
    return NCT_OK;
}

}  // extern "C"
