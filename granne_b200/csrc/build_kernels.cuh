// build_kernels.cuh — GPU-side GranneBuilder: batched `index_element`
// (reference: src/index/mod.rs:645-960).  Candidate generation reuses the search kernel on the layer under
// construction (the reference calls the very same search_for_neighbors, :819-820); this file holds what follows it:
// select_neighbors (:849-883), initialize_node (:886-895), connect_nodes (:899-921), add_and_limit_neighbors
// (:923-959) and the final pruning pass (:794-797), plus element construction (angular.rs:55-61,
// angular_int.rs:28-45).
//
// Parallel semantics: the reference inserts with rayon's par_iter over per-node RwLocks (:755-783), i.e. a batch of
// in-flight insertions that do not see each other's links until they are written.  Here a batch is explicit: all
// elements of a batch search the same snapshot of the layer, then one warp per element links it in under per-node
// spin locks (a warp holds at most one lock at a time).  Like the reference's default build, the resulting graph
// depends on the interleaving; every individual step uses the reference's exact arithmetic.
#pragma once

#include "search_kernels.cuh"

namespace granne_b200 {

struct BuildArgs {
    const uint32_t* ids;        // batch: element ids to index
    uint32_t n_batch;
    const uint32_t* cand_ids;   // [n_batch][cand_stride] search results on the layer under construction (ascending)
    const float* cand_dists;
    const uint32_t* cand_counts;
    uint32_t cand_stride;
    uint32_t* rows;             // the layer under construction, `stride` u32 per node, kUnusedId padded
    uint32_t stride;
    uint32_t node_width;        // node.len() == the builder's num_neighbors (row width of every layer, :394)
    uint32_t max_neighbors;     // config.num_neighbors of THIS layer (halved on upper layers, :665-668)
    int* locks;                 // one spin lock per node
    uint32_t stg_rows;
    uint32_t stg_row_bytes;
    uint32_t tile_rows;
    unsigned int* work_counter;
    uint32_t num_nodes;         // prune pass: nodes in the layer
};

constexpr float kHundredEps = 100.0f * 1.1920929e-07f;  // NotNan::new(100.0 * f32::EPSILON) (:813,829)

// ElementContainer::get(id) -> the query slot (c.qs / q registers) of the distance engine.
template <class Dist>
__device__ __forceinline__ void set_query_from_element(const DeviceIndex& ix, WarpCtx& c, Dist& dist, uint32_t id) {
    load_element_to_qs(ix, c, id);
    dist.load_query(ix, c);
}

__device__ __forceinline__ void node_lock(int* locks, uint32_t i, int lane) {
    if (lane == 0) {
        while (atomicCAS(locks + i, 0, 1) != 0) {
        }
        __threadfence();
    }
    __syncwarp();
}
__device__ __forceinline__ void node_unlock(int* locks, uint32_t i, int lane) {
    __syncwarp();
    if (lane == 0) {
        __threadfence();
        atomicExch(locks + i, 0);
    }
    __syncwarp();
}

// per-warp scratch for linking (shared memory)
struct LinkScratch {
    uint32_t* cid;   // candidate ids   (capacity >= max(cand_stride, node_width + 8, 40))
    float* cd;       // candidate dists
    uint32_t* sid;   // selected ids    (node slots = node_width rounded up to 32)
    float* sd;       // selected dists
    uint32_t* oid;   // the element's own selection (kept while neighbours are pruned)
    float* od;
    uint32_t* tid;   // add_and_limit: the node's current neighbours (+ the extra one) before sorting (node slots + 8)
    float* td;
};
__host__ __device__ constexpr uint32_t link_node_slots(uint32_t node_width) { return (node_width + 31u) & ~31u; }
// candidate scratch: the search results of one element, or a node's neighbours + 1 (whichever is larger)
__host__ __device__ constexpr uint32_t link_cand_cap(uint32_t cand_stride, uint32_t node_width) {
    return cand_stride > node_width + 8 ? (cand_stride > 40 ? cand_stride : 40)
                                        : (node_width + 8 > 40 ? node_width + 8 : 40);
}
// bytes of shared memory the link scratch needs behind the search context
__host__ __device__ constexpr size_t link_scratch_bytes(uint32_t cand_cap, uint32_t node_width) {
    return (size_t)cand_cap * 8 + (size_t)link_node_slots(node_width) * 16 + (size_t)(link_node_slots(node_width) + 8) * 8;
}

// select_neighbors (:849-883) over candidates cid/cd[0..nc) sorted by distance; result in sid/sd, returns count.
template <class Dist>
__device__ __forceinline__ int select_neighbors(const DeviceIndex& ix, WarpCtx& c, Dist& dist, LinkScratch& s, int nc,
                                                int max_neighbors) {
    const int lane = c.lane;
    if (nc <= max_neighbors) {
        for (int t = lane; t < nc; t += 32) {
            s.sid[t] = s.cid[t];
            s.sd[t] = s.cd[t];
        }
        __syncwarp();
        return nc;
    }
    int ns = 0;
    for (int ci = 0; ci < nc && ns < max_neighbors; ++ci) {
        const uint32_t j = s.cid[ci];
        const float d = s.cd[ci];
        bool ok = true;
        if (ns > 0) {
            // add j if it is closer to idx than to every neighbour selected so far: d <= dist(n, element j)
            // (32 selected neighbours per distance batch: num_neighbors is not limited to a warp's width)
            set_query_from_element(ix, c, dist, j);
            for (int b = 0; b < ns && ok; b += 32) {
                const int nb = (ns - b) < 32 ? (ns - b) : 32;
                const uint32_t my = s.sid[b + (lane < nb ? lane : 0)];
                const float dn = dist.dists(ix, c, my, nb);
                ok = !__any_sync(kFullMask, (lane < nb) && !(d <= dn));
            }
        }
        if (ok) {
            if (lane == 0) {
                s.sid[ns] = j;
                s.sd[ns] = d;
            }
            ns += 1;
            __syncwarp();
        }
    }
    return ns;
}

// add_and_limit_neighbors (:923-959) for node i (lock held or exclusive access); extra = (ej, ed) if has_extra.
// Rows wider than a warp are handled 32 slots at a time.
template <class Dist>
__device__ __forceinline__ void add_and_limit(const DeviceIndex& ix, WarpCtx& c, Dist& dist, LinkScratch& s,
                                              uint32_t* rows, uint32_t stride, uint32_t node_width, uint32_t i,
                                              bool has_extra, uint32_t ej, float ed, int limit) {
    const int lane = c.lane;
    uint32_t* row = rows + (size_t)i * stride;
    // take_while(!= UNUSED): rows are always written compactly; distances elements.dists(node_id, &neighbors) (:938)
    int cn = 0;
    for (uint32_t base = 0; base < node_width; base += 32) {
        const uint32_t v = (base + lane < node_width) ? __ldcg(row + base + lane) : kUnusedId;
        const unsigned vm = __ballot_sync(kFullMask, v != kUnusedId);
        const int run = __popc(vm & (((vm + 1u) & ~vm) - 1u));  // length of the leading run of valid slots
        if (run > 0) {
            if (cn == 0) set_query_from_element(ix, c, dist, i);
            const float dt = dist.dists(ix, c, v, run);
            if (lane < run) {
                s.tid[cn + lane] = v;
                s.td[cn + lane] = dt;
            }
        }
        cn += run;
        if (run < 32) break;
    }
    if (has_extra && lane == 0) {
        s.tid[cn] = ej;
        s.td[cn] = ed;
    }
    const int ct = cn + (has_extra ? 1 : 0);
    __syncwarp();
    // candidates.sort_unstable_by_key(|&(_, d)| d) (:945); ties keep their position (the reference's order among
    // equal distances is unspecified)
    for (int t = lane; t < ct; t += 32) {
        const float my_d = s.td[t];
        int rank = 0;
        for (int u = 0; u < ct; ++u) {
            const float du = s.td[u];
            rank += (du < my_d || (du == my_d && u < t)) ? 1 : 0;
        }
        s.cid[rank] = s.tid[t];
        s.cd[rank] = my_d;
    }
    __syncwarp();
    const int ns = select_neighbors(ix, c, dist, s, ct, limit);
    // set new neighbors and mark the remaining positions as unused (:950-958)
    for (uint32_t t = lane; t < node_width; t += 32) __stcg(row + t, (int)t < ns ? s.sid[t] : kUnusedId);
    __syncwarp();
}

// connect_nodes (:899-921): tries to add j as a neighbour of i (lock of i held).
template <class Dist>
__device__ __forceinline__ void connect_nodes(const DeviceIndex& ix, WarpCtx& c, Dist& dist, LinkScratch& s,
                                              uint32_t* rows, uint32_t stride, uint32_t node_width, uint32_t i,
                                              uint32_t j, float d) {
    if (i == j) return;
    const int lane = c.lane;
    uint32_t* row = rows + (size_t)i * stride;
    for (uint32_t base = 0; base < node_width; base += 32) {
        const uint32_t v = (base + lane < node_width) ? __ldcg(row + base + lane) : 0u;
        const unsigned hit = __ballot_sync(kFullMask, (base + lane < node_width) && (v == kUnusedId || v == j));
        if (hit) {  // the first free slot, or j is already a neighbour
            if (lane == 0) __stcg(row + base + (__ffs(hit) - 1), j);
            __syncwarp();
            return;
        }
    }
    add_and_limit(ix, c, dist, s, rows, stride, node_width, i, true, j, d, (int)node_width);
}

template <class Dist>
__device__ __forceinline__ void setup_ctx(const DeviceIndex& ix, WarpCtx& c, LinkScratch& s, unsigned char* smem_raw,
                                          uint32_t stg_rows, uint32_t stg_row_bytes, uint32_t tile_rows,
                                          uint32_t cand_cap, uint32_t node_width = 32) {
    c.lane = threadIdx.x;
    unsigned char* sp = smem_raw;
    c.tile = reinterpret_cast<float*>(sp);
    c.ids = reinterpret_cast<uint32_t*>(sp + tile_bytes_for_rows(tile_rows) - kIdScratchBytes);
    sp += tile_bytes_for_rows(tile_rows);
    c.bar = smem_u32(sp);
    c.phase = 0;
    c.pol_stream = make_policy_evict_first();
    c.pol_keep = make_policy_evict_last();
    sp += 16;
    const uint32_t qbytes = (ix.kind == kAngularI8) ? ix.row_stride : ((ix.dim + 3u) & ~3u) * 4u;
    c.qs = reinterpret_cast<float*>(sp);
    sp += (qbytes + 15u) & ~15u;
    c.xs = reinterpret_cast<float*>(sp);
    if (ix.kind == kSumEmbeddings) sp += (qbytes + 15u) & ~15u;
    c.stg = nullptr;
    c.stg_rows = stg_rows;
    c.stg_row_bytes = stg_row_bytes;
    if (stg_rows) {
        sp = smem_raw + (((size_t)(sp - smem_raw) + 127u) & ~(size_t)127u);
        c.stg = sp;
        sp += (size_t)stg_rows * stg_row_bytes;
    }
    if (Dist::kMbar) {
        if (c.lane == 0) mbar_init(c.bar, 1);
        __syncwarp();
    }
    c.status = 0;
    c.q_norm_i8 = 0;
    c.n_dist = c.n_expand = c.n_nbr = 0;
    c.list = nullptr;
    c.visited = nullptr;
    s.cid = reinterpret_cast<uint32_t*>(sp);
    sp += (size_t)cand_cap * 4;
    s.cd = reinterpret_cast<float*>(sp);
    sp += (size_t)cand_cap * 4;
    const uint32_t slots = link_node_slots(node_width);
    s.sid = reinterpret_cast<uint32_t*>(sp);
    sp += slots * 4;
    s.sd = reinterpret_cast<float*>(sp);
    sp += slots * 4;
    s.oid = reinterpret_cast<uint32_t*>(sp);
    sp += slots * 4;
    s.od = reinterpret_cast<float*>(sp);
    sp += slots * 4;
    s.tid = reinterpret_cast<uint32_t*>(sp);
    sp += (slots + 8) * 4;
    s.td = reinterpret_cast<float*>(sp);
}

// index_element (:805-846) after the candidate search, one warp per element of the batch.
template <class Dist>
__global__ void __launch_bounds__(32) build_link_kernel(const DeviceIndex ix, const BuildArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpCtx c;
    LinkScratch s;
    const uint32_t cand_cap = link_cand_cap(a.cand_stride, a.node_width);
    setup_ctx<Dist>(ix, c, s, smem_raw, a.stg_rows, a.stg_row_bytes, a.tile_rows, cand_cap, a.node_width);
    Dist dist;
    const int lane = c.lane;
    while (true) {
        unsigned int w0 = 0;
        if (lane == 0) w0 = atomicAdd(a.work_counter, 1u);
        const uint32_t w = __shfl_sync(kFullMask, w0, 0);
        if (w >= a.n_batch) break;
        const uint32_t idx = a.ids[w];
        // do not index elements that are zero: elements.dist(idx, idx) > 100 eps (:813)
        set_query_from_element(ix, c, dist, idx);
        const float dself = __shfl_sync(kFullMask, dist.dists(ix, c, idx, 1), 0);
        if (dself > kHundredEps) continue;
        // candidates.into_iter().filter(|&(id, _)| id != idx) (:822)
        const uint32_t cnt = a.cand_counts[w];
        int nc = 0;
        for (uint32_t base = 0; base < cnt; base += 32) {
            const uint32_t t = base + lane;
            uint32_t id = kUnusedId;
            float d = 0.0f;
            if (t < cnt) {
                id = a.cand_ids[(size_t)w * a.cand_stride + t];
                d = a.cand_dists[(size_t)w * a.cand_stride + t];
            }
            const bool keep = (t < cnt) && id != idx;
            const unsigned km = __ballot_sync(kFullMask, keep);
            if (keep) {
                const int p = nc + __popc(km & lanemask_lt());
                s.cid[p] = id;
                s.cd[p] = d;
            }
            nc += __popc(km);
        }
        __syncwarp();
        const int ns = select_neighbors(ix, c, dist, s, nc, (int)a.max_neighbors);
        // a duplicate of too many of its potential neighbours stays unconnected (:826-832)
        if ((uint32_t)ns > a.max_neighbors / 2 && s.sd[a.max_neighbors / 2] < kHundredEps) continue;
        for (int t = lane; t < ns; t += 32) {
            s.oid[t] = s.sid[t];
            s.od[t] = s.sd[t];
        }
        __syncwarp();
        // own node: initialize_node if empty, else connect_nodes for every neighbour (:834-841)
        node_lock(a.locks, idx, lane);
        {
            uint32_t* row = a.rows + (size_t)idx * a.stride;
            const uint32_t first = __ldcg(row);
            if (first == kUnusedId) {
                for (int t = lane; t < ns && (uint32_t)t < a.node_width; t += 32) __stcg(row + t, s.oid[t]);
                __syncwarp();
            } else {
                for (int t = 0; t < ns; ++t)
                    connect_nodes(ix, c, dist, s, a.rows, a.stride, a.node_width, idx, s.oid[t], s.od[t]);
            }
        }
        node_unlock(a.locks, idx, lane);
        // reverse links (:843-845)
        for (int t = 0; t < ns; ++t) {
            const uint32_t nb = s.oid[t];
            node_lock(a.locks, nb, lane);
            connect_nodes(ix, c, dist, s, a.rows, a.stride, a.node_width, nb, idx, s.od[t]);
            node_unlock(a.locks, nb, lane);
        }
    }
}

// limit number of neighbors after a pass (:794-797): add_and_limit_neighbors(node, i, &[], config.num_neighbors)
template <class Dist>
__global__ void __launch_bounds__(32) build_prune_kernel(const DeviceIndex ix, const BuildArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpCtx c;
    LinkScratch s;
    setup_ctx<Dist>(ix, c, s, smem_raw, a.stg_rows, a.stg_row_bytes, a.tile_rows, link_cand_cap(0, a.node_width),
                    a.node_width);
    Dist dist;
    while (true) {
        unsigned int w0 = 0;
        if (c.lane == 0) w0 = atomicAdd(a.work_counter, 1u);
        const uint32_t i = __shfl_sync(kFullMask, w0, 0);
        if (i >= a.num_nodes) break;
        add_and_limit(ix, c, dist, s, a.rows, a.stride, a.node_width, i, false, 0u, 0.0f, (int)a.max_neighbors);
    }
}

// Dist::dist between caller-supplied vectors (py compute_distance, py/src/lib.rs:71-89: `Vector::from(a).dist(
// &Vector::from(b))`): the b vectors are staged as the element container, a_i arrives as a raw query (normalised /
// quantised by prepare_query exactly like Vector::from) and pair i is dist_to_element(i, a_i).  One warp per pair.
struct PairArgs {
    const void* queries;  // n x dim raw f32
    uint32_t n;
    float* out;           // n distances
    int* error_flag;
    uint32_t stg_rows;
    uint32_t stg_row_bytes;
    uint32_t tile_rows;
};

template <class Dist>
__global__ void __launch_bounds__(32) pair_distance_kernel(const DeviceIndex ix, const PairArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpCtx c;
    LinkScratch s;
    setup_ctx<Dist>(ix, c, s, smem_raw, a.stg_rows, a.stg_row_bytes, a.tile_rows, 8);
    Dist dist;
    SearchArgs q{};
    q.queries = a.queries;
    q.query_format = kQueryRawF32;
    for (uint32_t i = blockIdx.x; i < a.n; i += gridDim.x) {
        c.status = 0;
        c.q_norm_i8 = 0;
        prepare_query(ix, q, c, i);
        dist.load_query(ix, c);
        const float d = dist.dists(ix, c, i, 1);
        if (__any_sync(kFullMask, c.status & kStatusNotFinite)) {
            if (c.lane == 0) atomicOr(a.error_flag, kStatusNotFinite);
        }
        if (c.lane == 0) a.out[i] = d;
        __syncwarp();
    }
}

// ids[i] = start + i * step  (step = +1 ascending insert pass, -1 for the reinsertion pass: reverse order, :776-782)
__global__ void iota_kernel(uint32_t* ids, uint32_t n, uint32_t start, int step) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ids[i] = (uint32_t)((long long)start + (long long)i * step);
}

__global__ void fill_u32_kernel(uint32_t* p, unsigned long long n, uint32_t v) {
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
        p[i] = v;
}

// Element construction from caller vectors: angular::Vector::from(Vec<f32>) = normalize_f32 (angular.rs:55-61,
// math.rs:124-150) or angular_int::Vector::quantize (angular_int.rs:28-45).  One warp per row, exact arithmetic.
__global__ void __launch_bounds__(32) make_elements_kernel(const float* __restrict__ raw, unsigned long long n,
                                                           uint32_t dim, int kind, void* __restrict__ out) {
    extern __shared__ float row_s[];
    const int lane = threadIdx.x;
    for (unsigned long long r = blockIdx.x; r < n; r += gridDim.x) {
        const float* src = raw + r * dim;
        __syncwarp();
        for (uint32_t i = lane; i < dim; i += 32) row_s[i] = src[i];
        __syncwarp();
        if (kind == kAngularI8) {
            float mx = 0.0f;
            for (uint32_t i = lane; i < dim; i += 32) mx = fmaxf(mx, fabsf(row_s[i]));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFullMask, mx, o));
            int8_t* dst = static_cast<int8_t*>(out) + r * dim;
            for (uint32_t i = lane; i < dim; i += 32) {
                const float vi = __fdiv_rn(__fmul_rn(row_s[i], 127.0f), mx);
                int q;
                if (vi != vi)
                    q = 0;
                else if (vi >= 127.0f)
                    q = 127;
                else if (vi <= -128.0f)
                    q = -128;
                else
                    q = (int)vi;
                dst[i] = (int8_t)q;
            }
        } else {
            const uint32_t full = dim / 32;
            float p = 0.0f;
            for (uint32_t ch = 0; ch < full; ++ch) {
                const float v = row_s[ch * 32 + lane];
                p = __fmaf_rn(v, v, p);
            }
            float acc = ordered_lane_sum_bcast(p);
            for (uint32_t t = full * 32; t < dim; ++t) {
                const float v = row_s[t];
                acc = __fmaf_rn(v, v, acc);
            }
            const float norm = __fsqrt_rn(acc);
            float* dst = static_cast<float*>(out) + r * dim;
            for (uint32_t i = lane; i < dim; i += 32) dst[i] = norm > 0.0f ? __fdiv_rn(row_s[i], norm) : row_s[i];
        }
    }
}

}  // namespace granne_b200
