// reorder.hpp — host side of Granne::reorder / reorder_by_keys (src/index/reorder.rs:59-292): applying an order to
// the graph and to the element files, the layer-preserving sort by keys, and the SumEmbeddings reordering keys
// (src/elements/embeddings/reorder.rs:31-58).  The order itself (compute_order, reorder.rs:126-174) is a batch of
// max_search = 1 layer searches and runs on the GPU (granne_b200_compute_order in granne_b200.cu).
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "formats.hpp"

namespace granne_b200 {

constexpr size_t kTrailLayers = 8;  // NUM_LAYERS, reorder.rs:177

// `order[i] == j`: the node / element at old index j moves to index i (reorder.rs:66-67).  Valid orders are
// permutations that keep every layer's nodes inside that layer's prefix (reorder_layer indexes `layer` with
// mapping[..layer.len()], reorder.rs:241-249 — an id outside the layer panics in the reference).
inline bool check_order(const HostGraph& g, const uint64_t* order, uint64_t n, std::string* err) {
    const uint64_t len = g.layers.empty() ? 0 : g.layers.back().num_nodes;
    if (n != len) {
        *err = "order length must equal Index::len";
        return false;
    }
    std::vector<uint8_t> seen(n, 0);
    for (uint64_t i = 0; i < n; ++i) {
        if (order[i] >= n || seen[order[i]]) {
            *err = "order is not a permutation";
            return false;
        }
        seen[order[i]] = 1;
    }
    for (const HostLayer& L : g.layers)
        for (uint64_t i = 0; i < L.num_nodes; ++i)
            if (order[i] >= L.num_nodes) {
                *err = "order moves a node out of its layer";
                return false;
            }
    return true;
}

// reorder_layers (reorder.rs:209-292): new node i takes the neighbours of old node order[i], renamed through the
// reverse mapping; MultiSetVector::push sorts every list (set_vector.rs:41-47).  Output: the index file image.
inline bool reorder_graph(const HostGraph& g, const uint64_t* order, uint64_t n, std::vector<uint8_t>* image,
                          std::string* err) {
    if (!check_order(g, order, n, err)) return false;
    std::vector<uint32_t> rev(n);
    for (uint64_t i = 0; i < n; ++i) rev[order[i]] = static_cast<uint32_t>(i);
    std::vector<std::vector<uint32_t>> rows(g.layers.size());
    std::vector<LayerView> views;
    for (size_t l = 0; l < g.layers.size(); ++l) {
        const HostLayer& L = g.layers[l];
        rows[l].assign(static_cast<size_t>(L.num_nodes) * L.width, kUnused);
        for (uint64_t i = 0; i < L.num_nodes; ++i) {
            const uint32_t* src = L.rows.data() + order[i] * L.width;
            uint32_t* dst = rows[l].data() + i * L.width;
            uint32_t k = 0;
            for (; k < L.width && src[k] != kUnused; ++k) dst[k] = rev[src[k]];
            std::sort(dst, dst + k);
        }
        views.push_back({rows[l].data(), L.num_nodes, L.width});
    }
    return encode_index(views, image, err);
}

// Permutable::permute for FixedWidthSliceVector (src/slice_vector/mod.rs:437-458): new row i = old row order[i];
// written like FixedWidthSliceVector::write (:460-466).
inline bool permute_dense(const uint8_t* buf, size_t len, size_t scalar_bytes, const uint64_t* order, uint64_t n,
                          std::vector<uint8_t>* out, std::string* err) {
    DenseView v;
    if (!parse_dense(buf, len, scalar_bytes, &v, err)) return false;
    if (v.num != n) {  // assert_eq!(self.len(), permutation.len())
        *err = "permutation length must equal the number of elements";
        return false;
    }
    const size_t row = static_cast<size_t>(v.dim) * scalar_bytes;
    out->resize(8 + static_cast<size_t>(n) * row);
    std::memcpy(out->data(), buf, 8);
    for (uint64_t i = 0; i < n; ++i) {
        if (order[i] >= n) {
            *err = "order is not a permutation";
            return false;
        }
        std::memcpy(out->data() + 8 + i * row, v.data + order[i] * row, row);
    }
    return true;
}

inline void store_le(uint8_t* p, uint64_t v, int nbytes) {
    for (int b = 0; b < nbytes; ++b) p[b] = static_cast<uint8_t>(v >> (8 * b));
}

// Permutable::permute for SumEmbeddings (src/elements/embeddings/mod.rs:191-217): the element -> embedding-id lists
// move, the embedding table stays.  Written like VariableWidthSliceVector::write (src/slice_vector/mod.rs:623-634):
// u64 count | (count + 1) five-byte offsets | three-byte ids.
inline bool permute_sum_elements(const uint8_t* buf, size_t len, const uint64_t* order, uint64_t n,
                                 std::vector<uint8_t>* out, std::string* err) {
    SumElements s;
    if (!parse_sum_elements(buf, len, &s, err)) return false;
    if (s.offsets.size() != n + 1) {
        *err = "permutation length must equal the number of elements";
        return false;
    }
    out->resize(8 + (static_cast<size_t>(n) + 1) * 5 + s.terms.size() * 3);
    store_le(out->data(), n, 8);
    uint8_t* off = out->data() + 8;
    uint8_t* data = off + (n + 1) * 5;
    uint64_t pos = 0;
    store_le(off, 0, 5);
    for (uint64_t i = 0; i < n; ++i) {
        if (order[i] >= n) {
            *err = "order is not a permutation";
            return false;
        }
        for (uint64_t t = s.offsets[order[i]]; t < s.offsets[order[i] + 1]; ++t) {
            if (pos >= s.terms.size()) {
                *err = "order is not a permutation";
                return false;
            }
            store_le(data + 3 * pos++, s.terms[t], 3);
        }
        store_le(off + 5 * (i + 1), pos, 5);
    }
    out->resize(8 + (static_cast<size_t>(n) + 1) * 5 + pos * 3);
    return true;
}

// The ordering part of reorder_by_keys (reorder.rs:96-108): inside every layer's id range, sort by (key, idx).
// Keys are rows of `kw` u64 compared lexicographically.
inline void order_by_keys(const std::vector<uint64_t>& layer_lens, const uint64_t* keys, size_t kw, uint64_t* order) {
    uint64_t begin = 0;
    for (uint64_t end : layer_lens) {
        for (uint64_t i = begin; i < end; ++i) order[i] = i;
        std::sort(order + begin, order + end, [&](uint64_t a, uint64_t b) {
            const uint64_t *ka = keys + a * kw, *kb = keys + b * kw;
            for (size_t j = 0; j < kw; ++j)
                if (ka[j] != kb[j]) return ka[j] < kb[j];
            return a < b;
        });
        begin = std::max(begin, end);
    }
}

// compute_order (reorder.rs:126-174) given the trails: eps[j][idx - layer_lens[j]] = the max_search = 1 result of
// layer j searched from node 0 for element idx (find_entrypoint_trail, reorder.rs:180-207 — the reference seeds
// layer j with eps[j], which is still 0 there), for idx >= layer_lens[j] and j < min(8, num_layers - 1).
// order_inv is filled only for ids >= layer_lens[0] (reorder.rs:161-165); first-layer ids map to 0 as in the
// reference.
inline void order_from_trails(const std::vector<uint64_t>& layer_lens, const std::vector<std::vector<uint32_t>>& eps,
                              uint64_t* order) {
    const size_t nl = layer_lens.size();
    if (nl == 0) return;
    for (uint64_t i = 0; i < layer_lens[0]; ++i) order[i] = i;
    std::vector<uint64_t> order_inv(nl >= 2 ? layer_lens[nl - 2] : 0, 0);
    using Trail = std::array<uint32_t, kTrailLayers>;
    for (size_t layer = 1; layer < nl; ++layer) {
        const uint64_t begin = layer_lens[layer - 1], end = layer_lens[layer];
        if (end <= begin) continue;
        std::vector<std::pair<Trail, uint64_t>> keyed(end - begin);
        const size_t take = std::min(kTrailLayers, layer);
        for (uint64_t idx = begin; idx < end; ++idx) {
            Trail t{};
            for (size_t j = 0; j < take; ++j) t[j] = static_cast<uint32_t>(order_inv[eps[j][idx - layer_lens[j]]]);
            for (size_t j = take; j < kTrailLayers; ++j) t[j] = order_inv.empty() ? 0u : static_cast<uint32_t>(order_inv[0]);
            keyed[idx - begin] = {t, idx};
        }
        std::sort(keyed.begin(), keyed.end());
        for (uint64_t i = begin; i < end; ++i) order[i] = keyed[i - begin].second;
        if (layer < nl - 1)
            for (uint64_t i = begin; i < end; ++i) order_inv[order[i]] = i;
    }
}

// embeddings::compute_keys_for_reordering (src/elements/embeddings/reorder.rs:31-58): per element its embedding ids by
// decreasing norm (stable sort by norm, reversed), the first 8, zero padded.  The norm is the plain left-to-right f32
// `iter().map(|x| x * x).sum().sqrt()` — separate multiply and add, not the 32-lane dot product.
inline bool embedding_reorder_keys(const uint8_t* elements, size_t elements_len, const uint8_t* embeddings,
                                   size_t embeddings_len, uint64_t* keys, std::string* err) {
    SumElements s;
    DenseView e;
    if (!parse_sum_elements(elements, elements_len, &s, err)) return false;
    if (!parse_dense(embeddings, embeddings_len, 4, &e, err)) return false;
    std::vector<float> norms(e.num);
    for (uint64_t w = 0; w < e.num; ++w) {
        volatile float acc = 0.0f;  // volatile: one rounding per multiply and per add, whatever the host flags
        for (uint64_t j = 0; j < e.dim; ++j) {
            float x;
            std::memcpy(&x, e.data + (w * e.dim + j) * 4, 4);
            volatile float sq = x * x;
            acc = acc + sq;
        }
        norms[w] = std::sqrt(static_cast<float>(acc));
    }
    const uint64_t n = s.offsets.size() - 1;
    std::vector<uint32_t> ids;
    for (uint64_t q = 0; q < n; ++q) {
        ids.assign(s.terms.begin() + s.offsets[q], s.terms.begin() + s.offsets[q + 1]);
        for (uint32_t id : ids)
            if (id >= e.num) {
                *err = "element refers to an embedding id outside the table";
                return false;
            }
        std::stable_sort(ids.begin(), ids.end(), [&](uint32_t a, uint32_t b) { return norms[a] < norms[b]; });
        std::reverse(ids.begin(), ids.end());
        for (size_t j = 0; j < kTrailLayers; ++j) keys[q * kTrailLayers + j] = j < ids.size() ? ids[j] : 0;
    }
    return true;
}

}  // namespace granne_b200
