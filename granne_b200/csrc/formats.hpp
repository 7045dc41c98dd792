// formats.hpp — host-side readers and writers for granne's on-disk formats.  Readers decode straight into the HBM
// staging layout; the writer (encode_index, bottom of this file) produces granne index files from staged rows.
//
// Read side: the index file (src/index/io.rs:72-113), the compressed layer blobs
// (src/slice_vector/offsets.rs:127-139,148-218,249-259; src/slice_vector/set_vector.rs:91-115,158-162), dense vector
// files (src/slice_vector/mod.rs:213-221) and SumEmbeddings element files (src/slice_vector/mod.rs:660-676,
// src/odd_byte_int.rs:3-36).  Nothing here is shared with oracle/ (the oracle is test infrastructure).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace granne_b200 {

// Splits [0, n) into contiguous ranges, one per worker thread (large layers: 100M nodes decode/encode in seconds
// instead of minutes).  fn(begin, end, worker).  Small inputs run inline.
template <class F>
inline void parallel_ranges(uint64_t n, F&& fn, uint64_t min_per_thread = 1u << 16) {
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 1;
    uint64_t t = std::min<uint64_t>(std::min<unsigned>(hw, 32u), n / min_per_thread);
    if (t <= 1) {
        fn((uint64_t)0, n, 0u);
        return;
    }
    std::vector<std::thread> th;
    const uint64_t per = (n + t - 1) / t;
    for (uint64_t w = 0; w < t; ++w) {
        const uint64_t b = w * per, e = std::min(n, b + per);
        if (b >= e) break;
        th.emplace_back([&fn, b, e, w] { fn(b, e, (unsigned)w); });
    }
    for (auto& x : th) x.join();
}
inline unsigned parallel_workers(uint64_t n, uint64_t min_per_thread = 1u << 16) {
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 1;
    const uint64_t t = std::min<uint64_t>(std::min<unsigned>(hw, 32u), n / min_per_thread);
    return t <= 1 ? 1u : (unsigned)t;
}

constexpr uint32_t kUnused = 0xFFFFFFFFu;   // NeighborId::max_value(), src/index/mod.rs:27-28
constexpr size_t kMetadataLen = 1024;       // src/index/io.rs:7
constexpr size_t kOffsetsPerChunk = 60;     // src/slice_vector/offsets.rs:7
constexpr size_t kChunkBytes = 128;         // repr(C) Chunk { usize initial; u16 deltas[60] }, offsets.rs:148-153

inline uint64_t load_le(const uint8_t* p, int nbytes) {
    uint64_t v = 0;
    for (int i = 0; i < nbytes; ++i) v |= static_cast<uint64_t>(p[i]) << (8 * i);
    return v;
}

// One HNSW layer decoded to fixed-width rows: row i = ascending neighbour ids of node i, padded with kUnused.
// (Equivalent to the reference's FixedWidthSliceVector<NeighborId> graph view, src/index/mod.rs:540-552.)
struct HostLayer {
    uint64_t num_nodes = 0;
    uint32_t width = 0;   // row stride in u32 (multiple of 8 -> 32-byte aligned rows), >= max degree
    uint32_t max_degree = 0;
    std::vector<uint32_t> rows;
};

struct HostGraph {
    std::vector<HostLayer> layers;
};

// ---- tiny JSON scanner for the 1 KiB header: "granne" + serde_json object (src/index/io.rs:46-67,89-113) ----------
class HeaderJson {
public:
    explicit HeaderJson(std::string text) : s_(std::move(text)) {}

    bool int_array(const char* key, std::vector<uint64_t>* out) const {
        size_t p;
        if (!value_pos(key, &p)) return false;
        skip_ws(&p);
        if (p >= s_.size() || s_[p] != '[') return false;
        ++p;
        out->clear();
        for (;;) {
            skip_ws(&p);
            if (p >= s_.size()) return false;
            if (s_[p] == ']') return true;
            if (s_[p] == ',') {
                ++p;
                continue;
            }
            uint64_t v;
            if (!parse_uint(&p, &v)) return false;
            out->push_back(v);
        }
    }

    bool uint_value(const char* key, uint64_t* out) const {
        size_t p;
        if (!value_pos(key, &p)) return false;
        skip_ws(&p);
        return parse_uint(&p, out);
    }

private:
    // Finds `"key"` at object nesting depth 1 outside of strings and returns the position after the colon.
    bool value_pos(const char* key, size_t* pos) const {
        const std::string quoted = std::string("\"") + key + "\"";
        int depth = 0;
        bool in_str = false;
        for (size_t i = 0; i < s_.size(); ++i) {
            char c = s_[i];
            if (in_str) {
                if (c == '\\')
                    ++i;
                else if (c == '"')
                    in_str = false;
                continue;
            }
            if (c == '{' || c == '[') {
                ++depth;
            } else if (c == '}' || c == ']') {
                --depth;
            } else if (c == '"') {
                if (depth == 1 && s_.compare(i, quoted.size(), quoted) == 0) {
                    size_t p = i + quoted.size();
                    skip_ws(&p);
                    if (p < s_.size() && s_[p] == ':') {
                        *pos = p + 1;
                        return true;
                    }
                }
                in_str = true;
            }
        }
        return false;
    }
    void skip_ws(size_t* p) const {
        while (*p < s_.size() && (s_[*p] == ' ' || s_[*p] == '\n' || s_[*p] == '\t' || s_[*p] == '\r')) ++*p;
    }
    bool parse_uint(size_t* p, uint64_t* out) const {
        if (*p >= s_.size() || s_[*p] < '0' || s_[*p] > '9') return false;
        uint64_t v = 0;
        while (*p < s_.size() && s_[*p] >= '0' && s_[*p] <= '9') v = v * 10 + static_cast<uint64_t>(s_[(*p)++] - '0');
        *out = v;
        return true;
    }
    std::string s_;
};

// ---- compressed neighbour lists -----------------------------------------------------------------------------------

// Decodes one MultiSetVector entry (set_vector.rs:91-115): count byte, then either `count` raw LE u32 (when the byte
// length says so) or a Stream VByte block of max(4,count) numbers (stream-vbyte 0.3.2 scalar layout: control bytes
// first, 2-bit length-1 codes, first number in the low bits); then a prefix sum (delta_decode, :158-162).
// Returns false on a truncated/garbled entry.  `out` receives exactly `count` ids.
inline bool decode_neighbor_list(const uint8_t* enc, size_t len, uint32_t* out, uint32_t* count_out) {
    if (len < 1) return false;
    const uint32_t count = enc[0];
    const uint8_t* body = enc + 1;
    const size_t body_len = len - 1;
    if (body_len == static_cast<size_t>(count) * 4) {
        for (uint32_t i = 0; i < count; ++i) out[i] = static_cast<uint32_t>(load_le(body + 4 * i, 4));
    } else {
        const uint32_t n = count < 4 ? 4 : count;
        const size_t nctrl = (n + 3) / 4;
        if (body_len < nctrl) return false;
        const uint8_t* data = body + nctrl;
        const uint8_t* end = body + body_len;
        for (uint32_t i = 0; i < n; ++i) {
            const int nb = ((body[i >> 2] >> ((i & 3) * 2)) & 3) + 1;
            if (data + nb > end) return false;
            const uint32_t v = static_cast<uint32_t>(load_le(data, nb));
            data += nb;
            if (i < count) out[i] = v;
        }
    }
    for (uint32_t i = 1; i < count; ++i) out[i] += out[i - 1];
    *count_out = count;
    return true;
}

// Decodes a whole layer blob: u64 offset_bytes | Chunk[offset_bytes/128] | encoded lists
// (CompressedVariableWidthSliceVector::load_mmap, offsets.rs:127-139).  Offsets are materialised with one running
// prefix sum per chunk instead of the reference's per-lookup O(60) sum (offsets.rs:171-176).
inline bool decode_layer(const uint8_t* blob, size_t len, HostLayer* layer, std::string* err) {
    if (len < 8) {
        *err = "layer blob shorter than its length prefix";
        return false;
    }
    const uint64_t offset_bytes = load_le(blob, 8);
    if (offset_bytes % kChunkBytes != 0 || 8 + offset_bytes > len) {
        *err = "layer offset table exceeds the blob";
        return false;
    }
    const uint8_t* chunks = blob + 8;
    const size_t num_chunks = offset_bytes / kChunkBytes;
    const uint8_t* data = blob + 8 + offset_bytes;
    const size_t data_len = len - 8 - offset_bytes;

    std::vector<uint64_t> offsets;
    offsets.reserve(num_chunks * kOffsetsPerChunk);
    for (size_t c = 0; c < num_chunks; ++c) {
        const uint8_t* ch = chunks + c * kChunkBytes;
        uint64_t cur = load_le(ch, 8);
        for (size_t i = 0; i < kOffsetsPerChunk; ++i) {
            const uint32_t delta = static_cast<uint32_t>(load_le(ch + 8 + 2 * i, 2));
            if (delta == 0xFFFFu) {  // UNUSED delta terminates the (last) chunk, offsets.rs:155,163-165
                if (c + 1 != num_chunks) {
                    *err = "unterminated offset chunk";
                    return false;
                }
                break;
            }
            cur += delta;
            offsets.push_back(cur);
        }
    }
    const uint64_t n = offsets.empty() ? 0 : offsets.size() - 1;

    // pass 1: degrees (count byte of each entry)
    const unsigned workers = parallel_workers(n);
    std::vector<uint32_t> wmax(workers, 0);
    std::vector<int> wbad(workers, 0);
    parallel_ranges(n, [&](uint64_t b, uint64_t e, unsigned w) {
        uint32_t mx = 0;
        for (uint64_t i = b; i < e; ++i) {
            if (offsets[i] >= offsets[i + 1] || offsets[i + 1] > data_len) {
                wbad[w] = 1;
                return;
            }
            const uint32_t cnt = data[offsets[i]];
            if (cnt > mx) mx = cnt;
        }
        wmax[w] = mx;
    });
    uint32_t max_deg = 0;
    for (unsigned w = 0; w < workers; ++w) {
        if (wbad[w]) {
            *err = "neighbour list offsets out of range";
            return false;
        }
        max_deg = std::max(max_deg, wmax[w]);
    }
    layer->num_nodes = n;
    layer->max_degree = max_deg;
    layer->width = ((max_deg < 1 ? 1 : max_deg) + 7u) & ~7u;
    layer->rows.resize(static_cast<size_t>(n) * layer->width);
    const uint32_t width = layer->width;
    uint32_t* all_rows = layer->rows.data();
    // pass 2: decode every list into its fixed-width row (padding written in the same pass)
    parallel_ranges(n, [&](uint64_t b, uint64_t e, unsigned w) {
        uint32_t tmp[256];
        for (uint64_t i = b; i < e; ++i) {
            uint32_t cnt = 0;
            if (!decode_neighbor_list(data + offsets[i], static_cast<size_t>(offsets[i + 1] - offsets[i]), tmp, &cnt)) {
                wbad[w] = 1;
                return;
            }
            uint32_t* row = all_rows + static_cast<size_t>(i) * width;
            for (uint32_t k = 0; k < cnt; ++k) {
                if (tmp[k] == kUnused) {
                    wbad[w] = 2;
                    return;
                }
                row[k] = tmp[k];
            }
            for (uint32_t k = cnt; k < width; ++k) row[k] = kUnused;
        }
    });
    for (unsigned w = 0; w < workers; ++w)
        if (wbad[w]) {
            *err = wbad[w] == 2 ? "neighbour id 0xFFFFFFFF is reserved" : "garbled neighbour list";
            return false;
        }
    return true;
}

// io::load_layers (src/index/io.rs:72-87): magic, JSON header, one blob per layer.
inline bool parse_index(const uint8_t* buf, size_t len, HostGraph* graph, std::string* err) {
    static const char kMagic[] = "granne";
    if (len < kMetadataLen || std::memcmp(buf, kMagic, 6) != 0) {
        *err = "Library string missing (not a granne index)";
        return false;
    }
    HeaderJson meta(std::string(reinterpret_cast<const char*>(buf) + 6, kMetadataLen - 6));
    std::vector<uint64_t> layer_sizes, layer_counts;
    uint64_t num_layers = 0;
    if (!meta.int_array("layer_sizes", &layer_sizes) || !meta.int_array("layer_counts", &layer_counts) ||
        !meta.uint_value("num_layers", &num_layers)) {
        *err = "Could not read metadata";
        return false;
    }
    if (num_layers != layer_counts.size() || layer_sizes.size() != layer_counts.size()) {
        *err = "metadata num_layers / layer_counts / layer_sizes disagree";
        return false;
    }
    size_t start = kMetadataLen;
    graph->layers.clear();
    graph->layers.resize(layer_sizes.size());
    for (size_t l = 0; l < layer_sizes.size(); ++l) {
        if (layer_sizes[l] > len - start) {
            *err = "layer " + std::to_string(l) + " exceeds the index file";
            return false;
        }
        if (!decode_layer(buf + start, static_cast<size_t>(layer_sizes[l]), &graph->layers[l], err)) {
            *err = "layer " + std::to_string(l) + ": " + *err;
            return false;
        }
        if (graph->layers[l].num_nodes != layer_counts[l]) {
            *err = "layer " + std::to_string(l) + " holds a different number of nodes than layer_counts says";
            return false;
        }
        start += static_cast<size_t>(layer_sizes[l]);
    }
    // Layer k is an id-prefix of layer k+1 (build_partial clones the previous layer, src/index/mod.rs:392-401);
    // every neighbour id must address a node of its own layer.
    for (size_t l = 0; l < graph->layers.size(); ++l) {
        const HostLayer& L = graph->layers[l];
        if (l > 0 && L.num_nodes < graph->layers[l - 1].num_nodes) {
            *err = "layers must grow monotonically";
            return false;
        }
        const uint64_t total = L.rows.size();
        const unsigned workers = parallel_workers(total, 1u << 20);
        std::vector<int> bad(workers, 0);
        const uint32_t* rp = L.rows.data();
        const uint64_t nn = L.num_nodes;
        parallel_ranges(total, [&](uint64_t b, uint64_t e, unsigned w) {
            for (uint64_t i = b; i < e; ++i)
                if (rp[i] != kUnused && rp[i] >= nn) {
                    bad[w] = 1;
                    return;
                }
        }, 1u << 20);
        for (unsigned w = 0; w < workers; ++w)
            if (bad[w]) {
                *err = "neighbour id out of range in layer " + std::to_string(l);
                return false;
            }
    }
    return true;
}

// ---- element files ------------------------------------------------------------------------------------------------

// FixedWidthSliceVector<T>::load_mmap (src/slice_vector/mod.rs:213-221): u64 width, then raw rows.
struct DenseView {
    const uint8_t* data = nullptr;
    uint64_t dim = 0;
    uint64_t num = 0;
};

inline bool parse_dense(const uint8_t* buf, size_t len, size_t scalar_bytes, DenseView* out, std::string* err) {
    if (len < 8) {
        *err = "elements file shorter than its width prefix";
        return false;
    }
    const uint64_t width = load_le(buf, 8);
    const size_t payload = len - 8;
    // assert!(width > 0 && data.len() % width == 0); a width beyond the payload is rejected BEFORE multiplying
    // (width * scalar_bytes must not wrap: 2^62 * 4 == 0 would divide by zero)
    if (width == 0 || width > payload / scalar_bytes + 1 || width > (1ull << 40) ||
        payload % (width * scalar_bytes) != 0) {
        *err = "elements file: width must be > 0 and divide the payload";
        return false;
    }
    out->data = buf + 8;
    out->dim = width;
    out->num = payload / (width * scalar_bytes);
    return true;
}

// VariableWidthSliceVector<ThreeByteInt, FiveByteInt>::load_mmap (src/slice_vector/mod.rs:660-676):
// u64 num_slices | (num_slices+1) x 5-byte offsets | 3-byte embedding ids.
struct SumElements {
    std::vector<uint64_t> offsets;  // num + 1, in units of terms
    std::vector<uint32_t> terms;
};

inline bool parse_sum_elements(const uint8_t* buf, size_t len, SumElements* out, std::string* err) {
    if (len < 8) {
        *err = "embeddings elements file too short";
        return false;
    }
    const uint64_t n = load_le(buf, 8);
    if (n > (len - 8) / 5 || 8 + (n + 1) * 5 > len) {
        *err = "embeddings elements file: offset table exceeds the file";
        return false;
    }
    const uint8_t* off = buf + 8;
    const uint8_t* data = off + (n + 1) * 5;
    const uint64_t nterms = (len - 8 - (n + 1) * 5) / 3;
    out->offsets.resize(n + 1);
    for (uint64_t i = 0; i <= n; ++i) {
        out->offsets[i] = load_le(off + 5 * i, 5);
        if (out->offsets[i] > nterms || (i > 0 && out->offsets[i] < out->offsets[i - 1])) {
            *err = "embeddings elements file: offsets must be monotone and inside the data";
            return false;
        }
    }
    out->terms.resize(nterms);
    for (uint64_t i = 0; i < nterms; ++i) out->terms[i] = static_cast<uint32_t>(load_le(data + 3 * i, 3));
    return true;
}

// ---- writers (Index::write_index, src/index/io.rs:11-70; set_vector.rs:117-148,169-221; offsets.rs:233-241) ----------

// Stream VByte (scalar) encoder for one list of max(4, count) numbers; layout as read by decode_neighbor_list.
inline void encode_vbyte(const uint32_t* nums, uint32_t n, std::vector<uint8_t>* out) {
    const size_t nctrl = (n + 3) / 4;
    const size_t base = out->size();
    out->resize(base + nctrl, 0);
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t v = nums[i];
        const int nb = v < (1u << 8) ? 1 : (v < (1u << 16) ? 2 : (v < (1u << 24) ? 3 : 4));
        (*out)[base + (i >> 2)] |= static_cast<uint8_t>((nb - 1) << ((i & 3) * 2));
        for (int b = 0; b < nb; ++b) out->push_back(static_cast<uint8_t>(v >> (8 * b)));
    }
}

// set_encode (set_vector.rs:117-148): `ids` ascending; count byte, then vbyte deltas or raw u32 when not smaller.
inline void encode_neighbor_list(const uint32_t* ids, uint32_t count, std::vector<uint8_t>* out) {
    if (count > 255) count = 255;
    uint32_t delta[256];
    for (uint32_t i = 0; i < count; ++i) delta[i] = i ? ids[i] - ids[i - 1] : ids[i];
    uint32_t n = count;
    while (n < 4) delta[n++] = 0;
    std::vector<uint8_t> enc;
    encode_vbyte(delta, n, &enc);
    out->push_back(static_cast<uint8_t>(count));
    if (enc.size() >= 4u * count) {
        for (uint32_t i = 0; i < count; ++i)
            for (int b = 0; b < 4; ++b) out->push_back(static_cast<uint8_t>(delta[i] >> (8 * b)));
    } else {
        out->insert(out->end(), enc.begin(), enc.end());
    }
}

// One layer blob from fixed-width rows (FixedWidthSliceVector::write_as_multi_set_vector, set_vector.rs:169-221).
inline bool encode_layer(const uint32_t* rows, uint64_t num_nodes, uint32_t stride, std::vector<uint8_t>* out,
                         std::string* err) {
    const size_t bytes_for_offsets = (1 + num_nodes / kOffsetsPerChunk) * kChunkBytes;
    const size_t base = out->size();
    out->resize(base + 8 + bytes_for_offsets, 0xFF);
    for (int b = 0; b < 8; ++b) (*out)[base + b] = static_cast<uint8_t>(static_cast<uint64_t>(bytes_for_offsets) >> (8 * b));
    // every worker encodes a contiguous node range into its own buffer; the buffers are concatenated in order
    std::vector<uint64_t> offsets(num_nodes + 1, 0);  // offsets[i + 1] = encoded size of node i, prefix-summed below
    const unsigned workers = parallel_workers(num_nodes);
    std::vector<std::vector<uint8_t>> parts(workers);
    parallel_ranges(num_nodes, [&](uint64_t b, uint64_t e, unsigned w) {
        std::vector<uint8_t>& buf = parts[w];
        buf.reserve((e - b) * (stride + 8));
        std::vector<uint32_t> tmp;
        for (uint64_t i = b; i < e; ++i) {
            tmp.clear();
            const uint32_t* row = rows + i * stride;
            for (uint32_t k = 0; k < stride; ++k)
                if (row[k] != kUnused) tmp.push_back(row[k]);  // predicate |&x| x != UNUSED (io.rs:33)
            std::sort(tmp.begin(), tmp.end());
            const size_t before = buf.size();
            encode_neighbor_list(tmp.data(), static_cast<uint32_t>(tmp.size()), &buf);
            offsets[i + 1] = buf.size() - before;
        }
    });
    for (uint64_t i = 0; i < num_nodes; ++i) offsets[i + 1] += offsets[i];
    for (unsigned w = 0; w < workers; ++w) {
        out->insert(out->end(), parts[w].begin(), parts[w].end());
        std::vector<uint8_t>().swap(parts[w]);
    }
    // Offsets::push: chunks of 60 u16 deltas with a u64 `initial` (offsets.rs:148-241)
    uint8_t* chunks = out->data() + base + 8;
    const size_t num_chunks = bytes_for_offsets / kChunkBytes;
    for (size_t c = 0; c < num_chunks; ++c) {
        uint8_t* ch = chunks + c * kChunkBytes;
        const size_t first = c * kOffsetsPerChunk;
        const uint64_t initial = first < offsets.size() ? offsets[first] : 0;
        for (int b = 0; b < 8; ++b) ch[b] = static_cast<uint8_t>(initial >> (8 * b));
        uint64_t prev = initial;
        for (size_t k = 0; k < kOffsetsPerChunk && first + k < offsets.size(); ++k) {
            const uint64_t d = offsets[first + k] - prev;
            if (d >= 0xFFFF) {
                *err = "neighbour list too long for a u16 offset delta";
                return false;
            }
            ch[8 + 2 * k] = static_cast<uint8_t>(d);
            ch[8 + 2 * k + 1] = static_cast<uint8_t>(d >> 8);
            prev = offsets[first + k];
        }
    }
    return true;
}

struct LayerView {
    const uint32_t* rows;
    uint64_t num_nodes;
    uint32_t stride;
};

// write_index (io.rs:11-70): 1 KiB "granne" + JSON header (serde_json's default map orders keys alphabetically),
// then one MultiSetVector blob per layer.
inline bool encode_index(const std::vector<LayerView>& layers, std::vector<uint8_t>* out, std::string* err) {
    out->assign(kMetadataLen, static_cast<uint8_t>(' '));
    std::vector<uint64_t> sizes, counts;
    for (const LayerView& L : layers) {
        const size_t before = out->size();
        if (!encode_layer(L.rows, L.num_nodes, L.stride, out, err)) return false;
        sizes.push_back(out->size() - before);
        counts.push_back(L.num_nodes);
    }
    uint64_t num_neighbors = 0;  // degree of node 0 in the last layer (io.rs:20-24)
    if (!layers.empty() && layers.back().num_nodes > 0)
        for (uint32_t k = 0; k < layers.back().stride; ++k)
            if (layers.back().rows[k] != kUnused) ++num_neighbors;
    auto arr = [](const std::vector<uint64_t>& v) {
        std::string s = "[";
        for (size_t i = 0; i < v.size(); ++i) s += (i ? "," : "") + std::to_string(v[i]);
        return s + "]";
    };
    const std::string meta = std::string("granne") + "{\"compressed\":true,\"granne_version\":\"0.5.2\",\"layer_counts\":" +
                             arr(counts) + ",\"layer_sizes\":" + arr(sizes) + ",\"num_elements\":" +
                             std::to_string(counts.empty() ? 0 : counts.back()) + ",\"num_layers\":" +
                             std::to_string(layers.size()) + ",\"num_neighbors\":" + std::to_string(num_neighbors) +
                             ",\"version\":2}";
    if (meta.size() > kMetadataLen) {
        *err = "metadata does not fit the 1024-byte header";
        return false;
    }
    std::memcpy(out->data(), meta.data(), meta.size());
    return true;
}

// compute_num_elements_in_layer (src/index/mod.rs:634-643)
inline uint64_t num_elements_in_layer(uint64_t total, float layer_multiplier, uint64_t layer_idx) {
    const double m = static_cast<double>(layer_multiplier);
    const double e = std::floor(std::log(static_cast<double>(total)) / std::log(m)) - static_cast<double>(layer_idx);
    const double v = std::ceil(static_cast<double>(total) / std::pow(m, e));
    const uint64_t r = v <= 0 ? 0 : (v >= 1.8e19 ? ~0ull : static_cast<uint64_t>(v));
    return r < total ? r : total;
}

}  // namespace granne_b200
