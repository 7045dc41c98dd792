// C++ client of include/granne_b200.hpp, written like the reference's own tests (src/index/tests.rs).
//   cxx_client nodevice   host-only checks; every staging call must fail with GRANNE_B200_ERR_NO_DEVICE
//   cxx_client gpu        build_and_search_float, write_and_load, append_elements, reorder_index on the GPU
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include "granne_b200.hpp"

namespace gb = granne_b200;

#define REQUIRE(cond)                                                      \
    do {                                                                   \
        if (!(cond)) {                                                     \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);  \
            return 1;                                                      \
        }                                                                  \
    } while (0)

// test_helper::random_vector (src/test_helper.rs:3-18): components uniform in [-0.5, 0.5)
static std::vector<float> random_rows(size_t n, size_t dim, uint64_t seed) {
    std::vector<float> v(n * dim);
    uint64_t s = seed * 0x9E3779B97F4A7C15ull + 0x2545F4914F6CDD1Dull;
    for (float& x : v) {
        s ^= s << 13;
        s ^= s >> 7;
        s ^= s << 17;
        x = static_cast<float>((s >> 40) / 16777216.0) - 0.5f;
    }
    return v;
}

// verify_search (tests.rs:50-62): fraction of elements that find themselves first
static double self_recall(const gb::Granne& index, size_t max_search) {
    size_t found = 0;
    for (size_t i = 0; i < index.len(); ++i) {
        const gb::SearchResult r = index.search(index.get_element(i), max_search, 1);
        if (!r.empty() && r[0].first == i) ++found;
    }
    return index.len() ? static_cast<double>(found) / static_cast<double>(index.len()) : 1.0;
}

static bool findable(const gb::Granne& index, const float* raw, size_t dim, size_t idx, size_t max_search) {
    const gb::SearchResult r = index.search_raw(std::vector<float>(raw, raw + dim), max_search, 1);
    return std::any_of(r.begin(), r.end(), [&](const std::pair<size_t, float>& e) { return e.first == idx; });
}

template <class F>
static int error_code_of(F&& f) {
    try {
        f();
    } catch (const gb::Error& e) {
        return e.code();
    }
    return GRANNE_B200_OK;
}

static int nodevice() {
    const gb::BuildConfig d;  // BuildConfig::default() (mod.rs:220-231)
    REQUIRE(d.raw().num_neighbors == 30 && d.raw().max_search == 200 && d.raw().layer_multiplier == 15.0f);
    REQUIRE(d.raw().reinsert_elements == 1 && d.raw().expected_num_elements < 0);
    const gb::BuildConfig c = gb::BuildConfig().num_neighbors(20).max_search(5).layer_multiplier(10.0f)
                                  .expected_num_elements(1000).reinsert_elements(false);
    REQUIRE(c.raw().num_neighbors == 20 && c.raw().max_search == 5 && c.raw().layer_multiplier == 10.0f);
    REQUIRE(c.raw().expected_num_elements == 1000 && c.raw().reinsert_elements == 0 && d.raw().num_neighbors == 30);
    gb::Elements e;  // a hand-made angular::Vectors image: width 3, two rows
    e.bytes.assign(8 + 2 * 3 * 4, 0);
    e.bytes[0] = 3;
    REQUIRE(e.dim() == 3 && e.len() == 2);
    const std::vector<float> raw = random_rows(4, 8, 1);
    // no CPU fallback anywhere above the C ABI
    REQUIRE(error_code_of([&] { gb::Elements::from_raw(gb::ElementKind::Angular, raw.data(), 4, 8); }) == GRANNE_B200_ERR_NO_DEVICE);
    REQUIRE(error_code_of([&] { gb::GranneBuilder b(gb::BuildConfig(), e); }) == GRANNE_B200_ERR_NO_DEVICE);
    const std::vector<uint8_t> junk(2048, 'x');
    REQUIRE(error_code_of([&] { gb::Granne::from_bytes(junk, e); }) == GRANNE_B200_ERR_NO_DEVICE);
    REQUIRE(error_code_of([&] { gb::compute_distance(gb::ElementKind::Angular, {1.f, 0.f}, {0.f, 1.f}); }) == GRANNE_B200_ERR_NO_DEVICE);
    REQUIRE(error_code_of([&] { gb::read_file("/nonexistent/granne"); }) == GRANNE_B200_ERR_IO);
    std::printf("nodevice ok\n");
    return 0;
}

static int gpu() {
    using K = gb::ElementKind;
    {  // build_and_search_float (tests.rs:41-48): num_neighbors 20, max_search 20; verify_search(index, 0.95, 10)
        const size_t n = 1500, dim = 28;
        const std::vector<float> raw = random_rows(n, dim, 2);
        const gb::Elements elements = gb::Elements::from_raw(K::Angular, raw.data(), n, dim);
        REQUIRE(elements.len() == n && elements.dim() == dim);
        gb::GranneBuilder builder(gb::BuildConfig().num_neighbors(20).max_search(20), elements);
        REQUIRE(builder.num_elements() == n && builder.len() == 0);
        builder.build();
        REQUIRE(builder.len() == n);
        const gb::Granne index = builder.get_index();
        REQUIRE(index.len() == n && index.num_layers() == builder.num_layers());
        REQUIRE(self_recall(index, 10) > 0.95);

        // write_and_load (tests.rs:336-372): the written file loads into an identical graph
        const std::vector<uint8_t> image = builder.write_index();
        const gb::Granne loaded = gb::Granne::from_bytes(image, elements);
        REQUIRE(loaded.num_layers() == builder.num_layers() && loaded.len() == builder.len());
        for (size_t layer = 0; layer < loaded.num_layers(); ++layer) {
            REQUIRE(loaded.layer_len(layer) == builder.layer_len(layer));
            for (size_t i = 0; i < loaded.layer_len(layer); i += 7) {
                std::vector<size_t> a = builder.get_neighbors(i, layer), b = loaded.get_neighbors(i, layer);
                std::sort(a.begin(), a.end());
                std::sort(b.begin(), b.end());
                REQUIRE(a == b);
            }
        }
        REQUIRE(loaded.write_index() == image);
        size_t samples = 0, exact = 0;  // dist_to_element(i, get_element(i)) < DIST_EPSILON (tests.rs:365-367)
        for (size_t i = 0; i < n; i += 97, ++samples) {
            const gb::SearchResult r = loaded.search(loaded.get_element(i), 100, 1);
            exact += (!r.empty() && r[0].first == i && r[0].second < 10.0f * 1.1920929e-07f) ? 1 : 0;
        }
        REQUIRE(exact + 1 >= samples);
        // max_search == 0 panics in the reference (mod.rs:1019)
        REQUIRE(error_code_of([&] { loaded.search(loaded.get_element(0), 0, 1); }) == GRANNE_B200_ERR_INVALID_ARGUMENT);

        // reorder_index (reorder.rs:298-323): results map through the returned permutation
        gb::Granne reordered = gb::Granne::from_bytes(image, elements);
        const std::vector<size_t> permutation = reordered.reorder(false);
        REQUIRE(permutation.size() == n);
        for (size_t idx : {size_t(0), size_t(10), size_t(123), size_t(99), size_t(499)}) {
            const std::vector<float> element = loaded.get_element(idx);
            const gb::SearchResult exp = loaded.search(element, 10, 10), res = reordered.search(element, 10, 10);
            REQUIRE(exp.size() == res.size());
            for (size_t i = 0; i < exp.size(); ++i) REQUIRE(exp[i].first == permutation[res[i].first]);
        }
    }
    {  // append_elements (tests.rs:503-567)
        const size_t dim = 50;
        const std::vector<float> first = random_rows(500, dim, 3), second = random_rows(500, dim, 4);
        gb::GranneBuilder builder(
            gb::BuildConfig().expected_num_elements(1000).layer_multiplier(10.0f).num_neighbors(20).max_search(50),
            gb::Elements::from_raw(K::Angular, first.data(), 500, dim));
        builder.build();
        REQUIRE(builder.num_layers() == 3 && builder.layer_len(2) == 500);
        REQUIRE(findable(builder.get_index(), &first[123 * dim], dim, 123, 50));
        builder.push(gb::Elements::from_raw(K::Angular, second.data(), 500, dim));
        REQUIRE(builder.num_elements() == 1000 && builder.len() == 500);
        builder.build();
        REQUIRE(builder.num_layers() == 3 && builder.layer_len(2) == 1000);
        const gb::Granne index = builder.get_index();
        REQUIRE(findable(index, &first[123 * dim], dim, 123, 50));
        REQUIRE(findable(index, &second[123 * dim], dim, 500 + 123, 50));
        // "Cannot index fewer elements than already in index." (mod.rs:379-382)
        REQUIRE(error_code_of([&] { builder.build_partial(10); }) == GRANNE_B200_ERR_INVALID_ARGUMENT);
    }
    {  // build_and_search_int8 (tests.rs:114-132) + compute_distance
        const size_t n = 500, dim = 32;
        const std::vector<float> raw = random_rows(n, dim, 5);
        const gb::Elements elements = gb::Elements::from_raw(K::AngularInt, raw.data(), n, dim);
        gb::GranneBuilder builder(gb::BuildConfig().num_neighbors(20).max_search(20), elements);
        builder.build();
        const gb::Granne index = builder.get_index();
        size_t found = 0;
        for (size_t i = 0; i < n; ++i) {
            const gb::SearchResult r = index.search(index.get_element_i8(i), 10, 1);
            found += (!r.empty() && r[0].first == i) ? 1 : 0;
        }
        REQUIRE(found > 0.95 * n);
        const std::vector<float> a(raw.begin(), raw.begin() + dim), b(raw.begin() + dim, raw.begin() + 2 * dim);
        REQUIRE(std::fabs(gb::compute_distance(K::Angular, a, a)) < 1e-5f);
        const float dab = gb::compute_distance(K::Angular, a, b);
        REQUIRE(dab > 0.0f && dab < 2.0f && dab == gb::compute_distance(K::Angular, b, a));
    }
    {  // several GPUs (here: cuda:0 named twice) behind one handle: same answers as the single-device index
        const size_t n = 1200, dim = 24;
        const std::vector<float> raw = random_rows(n, dim, 9);
        const gb::Elements elements = gb::Elements::from_raw(K::Angular, raw.data(), n, dim);
        gb::GranneBuilder builder(gb::BuildConfig().num_neighbors(12).max_search(30), elements);
        builder.build();
        const gb::Granne index = builder.get_index();
        const std::vector<uint8_t> image = index.write_index();
        const gb::MultiGranne multi = gb::MultiGranne::replicated(image, elements, {0, 0});
        REQUIRE(multi.len() == n && multi.num_parts() == 2);
        const std::vector<float> q = random_rows(37, dim, 10);
        const std::vector<gb::MultiGranne::Result> got = multi.search_batch_raw(q, 30, 5);
        REQUIRE(got.size() == 37);
        for (size_t i = 0; i < got.size(); ++i) {
            const gb::SearchResult want = index.search_raw(std::vector<float>(q.begin() + i * dim, q.begin() + (i + 1) * dim), 30, 5);
            REQUIRE(want.size() == got[i].size());
            for (size_t j = 0; j < want.size(); ++j)
                REQUIRE(want[j].first == got[i][j].first && want[j].second == got[i][j].second);
        }
        // two shards = the same index twice: ids of the second copy are offset by n, every hit appears twice
        const gb::MultiGranne parts = gb::MultiGranne::partitioned({image, image}, {elements, elements}, {0});
        REQUIRE(parts.len() == 2 * n && parts.shard_base(1) == n);
        const std::vector<gb::MultiGranne::Result> two = parts.search_batch_raw(q, 30, 4);
        for (size_t i = 0; i < two.size(); ++i) {
            REQUIRE(two[i].size() == 4);
            REQUIRE(two[i][1].first == two[i][0].first + n && two[i][1].second == two[i][0].second);
        }
    }
    std::printf("gpu ok\n");
    return 0;
}

int main(int argc, char** argv) {
    try {
        if (argc > 1 && std::strcmp(argv[1], "gpu") == 0) return gpu();
        return nodevice();
    } catch (const gb::Error& e) {
        std::printf("granne_b200::Error %d: %s\n", e.code(), e.what());
        return 2;
    }
}
