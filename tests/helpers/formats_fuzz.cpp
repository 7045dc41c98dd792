// Sanitizer harness for the host-side file-format code (granne_b200/csrc/formats.hpp, reorder.hpp): parses mutated
// index / elements / embeddings images; any out-of-bounds access aborts under -fsanitize=address,undefined.
// usage: formats_fuzz <index> <elements> <kind 0|1|2> <embeddings|-> <iterations> <seed>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>

#include "formats.hpp"
#include "reorder.hpp"

namespace gb = granne_b200;
using Bytes = std::vector<uint8_t>;

static Bytes slurp(const char* path) {
    std::ifstream f(path, std::ios::binary);
    return Bytes(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
}

static uint64_t state = 88172645463325252ull;
static uint64_t rnd() {
    state ^= state << 13;
    state ^= state >> 7;
    state ^= state << 17;
    return state;
}

static Bytes mutate(const Bytes& src) {
    Bytes b = src;
    if (b.empty()) return b;
    switch (rnd() % 5) {
        case 0: b.resize(rnd() % b.size()); break;
        case 1:
            for (int i = 0, n = 1 + rnd() % 8; i < n; ++i) b[rnd() % b.size()] = (uint8_t)rnd();
            break;
        case 2: {
            size_t at = rnd() % b.size();
            for (size_t i = at; i < b.size() && i < at + 8; ++i) b[i] = (uint8_t)rnd();
            break;
        }
        case 3: b[rnd() % std::min<size_t>(b.size(), 1100)] = (uint8_t)rnd(); break;  // header / first chunk
        default: {
            size_t at = rnd() % b.size();
            b[at] = (rnd() & 1) ? 0xFF : 0x00;
        }
    }
    return b;
}

int main(int argc, char** argv) {
    if (argc < 7) return 2;
    const Bytes index = slurp(argv[1]), elements = slurp(argv[2]);
    const int kind = std::atoi(argv[3]);
    const Bytes embeddings = argv[4][0] == '-' ? Bytes() : slurp(argv[4]);
    const long iters = std::atol(argv[5]);
    state ^= (uint64_t)std::atoll(argv[6]) * 0x9E3779B97F4A7C15ull;
    long parsed = 0, rejected = 0;
    for (long it = 0; it < iters; ++it) {
        const Bytes ib = (it % 3 == 1) ? index : mutate(index);
        const Bytes eb = (it % 3 == 0) ? elements : mutate(elements);
        const Bytes mb = (it % 4 == 3) ? mutate(embeddings) : embeddings;
        std::string err;
        gb::HostGraph graph;
        const bool ok = gb::parse_index(ib.data(), ib.size(), &graph, &err);
        ok ? ++parsed : ++rejected;
        if (ok) {
            std::vector<gb::LayerView> views;
            for (const gb::HostLayer& L : graph.layers) views.push_back({L.rows.data(), L.num_nodes, L.width});
            Bytes image;
            gb::encode_index(views, &image, &err);
            const uint64_t n = graph.layers.empty() ? 0 : graph.layers.back().num_nodes;
            std::vector<uint64_t> order(n);
            for (uint64_t i = 0; i < n; ++i) order[i] = (it % 5 == 4) ? rnd() % (n + 2) : i;
            Bytes out;
            gb::reorder_graph(graph, order.data(), n, &out, &err);
            Bytes pe;
            if (kind == 2)
                gb::permute_sum_elements(eb.data(), eb.size(), order.data(), n, &pe, &err);
            else
                gb::permute_dense(eb.data(), eb.size(), kind == 0 ? 4 : 1, order.data(), n, &pe, &err);
        }
        if (kind == 2) {
            gb::SumElements s;
            if (gb::parse_sum_elements(eb.data(), eb.size(), &s, &err)) {
                std::vector<uint64_t> keys((s.offsets.size() - 1) * gb::kTrailLayers);
                gb::embedding_reorder_keys(eb.data(), eb.size(), mb.data(), mb.size(), keys.data(), &err);
            }
        } else {
            gb::DenseView v;
            gb::parse_dense(eb.data(), eb.size(), kind == 0 ? 4 : 1, &v, &err);
        }
    }
    // hostile width prefixes: width * scalar_bytes wraps in 64 bits (2^62 * 4 == 0 used to divide by zero, 2^62 + 1
    // wrapped to 4 and was accepted with dim ~ 2^62) — all must be rejected with a status, never crash
    for (uint64_t w : {1ull << 62, (1ull << 62) + 1, 1ull << 63, ~0ull, (1ull << 61) + 3, 1ull << 41}) {
        Bytes img(8 + 64, 0);
        for (int b = 0; b < 8; ++b) img[b] = (uint8_t)(w >> (8 * b));
        for (size_t sb : {size_t(4), size_t(1)}) {
            gb::DenseView v;
            std::string err;
            if (gb::parse_dense(img.data(), img.size(), sb, &v, &err)) {
                std::printf("hostile width %llu accepted\n", (unsigned long long)w);
                return 3;
            }
        }
    }
    std::printf("parsed %ld rejected %ld\n", parsed, rejected);
    return 0;
}
