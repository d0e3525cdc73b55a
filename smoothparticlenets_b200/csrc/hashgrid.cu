// Hash-grid neighbour search for sm_100a: grid bounds, cell keys, a batched stable onesweep radix
// sort of (cell key, particle index), the ReorderData permutation, the cell table and the
// fixed-width neighbour lists.
//
// Replaces kernel_compute_cellIDs / CUB SortPairs / kernel_fill_cells / kernel_compute_collisions /
// kernel_reorder_data and their launchers (reference src/gpu_kernels.cu:238-548) and the torch
// bounds code of ParticleCollision.forward (ParticleCollision.py:174-181).  Semantics follow
// loc2grid / partial_grid_hash / compute_collisions (src/common_funcs.h:96-119, 875-948).
//
// Design notes (DESIGN.md has the byte counts):
//  * everything is stream-ordered and batched over scenes; the only host-side decisions are grid
//    sizes that depend on (B, N, D, max_grid_dim);
//  * the sort is an LSD onesweep: one histogram kernel (fused with key generation) for all digit
//    places, then one kernel per digit place that ranks a 2048-key tile with warp match/ballot,
//    obtains its global offsets by decoupled look-back over single-word {flag,count} statuses, and
//    scatters.  Tiles take tickets from an atomic counter so every predecessor is already running.
//    Only ceil(log2(ncells+1)) key bits are sorted, split evenly over the passes on the device;
//  * neighbour lists: one warp per 8 consecutive queries; queries in the same cell share a candidate
//    set that is resolved once and staged in shared memory, hits are compacted with ballot/popc so
//    rows are written coalesced, including the -1 padding.
#include "spnb_common.cuh"
#include "tile_lists.cuh"

namespace spnb {

constexpr int kMaxPasses = 4;
constexpr int kRadix = 256;
constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortThreads * kSortItems;
constexpr uint32_t kFlagAgg = 1u << 30;
constexpr uint32_t kFlagIncl = 2u << 30;
constexpr uint32_t kValMask = (1u << 30) - 1;

struct SceneCtl {
    int ncells;  // prod(grid_dims); keys are 0..ncells-1, ncells marks "outside every cell"
    int bits;    // key bits to sort
    int shift[kMaxPasses];
    int width[kMaxPasses];
};

struct WsLayout {
    size_t minmax_off, minmax_bytes;  // uint32 [B][D][2]      (zeroed by spnb_grid_bounds)
    size_t ctl_off;                   // SceneCtl [B]           (start of the sort control block)
    size_t err_off;                   // int
    size_t ticket_off;                // uint32 [kMaxPasses]
    size_t hist_off;                  // uint32 [B][P][256]
    size_t status_off;                // uint32 [P][B*tiles][256]
    size_t ctl_end;                   // end of the block zeroed by spnb_hashgrid_order
    size_t keys_a_off, keys_b_off, vals_a_off, vals_b_off;  // uint32 [B*N] each
    size_t total;
    int passes, tiles;
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static WsLayout ws_layout(int B, int N, int D, int G)
{
    WsLayout w;
    double cells = 1.0;
    for (int k = 0; k < D; ++k) cells *= (double)G;
    int bits = 1;
    while (bits < 32 && (double)(1ull << bits) <= cells) ++bits;  // keys 0..cells inclusive
    w.passes = (bits + 7) / 8;
    if (w.passes < 1) w.passes = 1;
    if (w.passes > kMaxPasses) w.passes = kMaxPasses;
    w.tiles = cdiv(N, kSortTile);
    size_t off = 0;
    w.minmax_off = off;
    w.minmax_bytes = sizeof(uint32_t) * (size_t)B * D * 2;
    off = align_up(off + w.minmax_bytes, 256);
    w.ctl_off = off;
    off = align_up(off + sizeof(SceneCtl) * (size_t)B, 256);
    w.err_off = off;
    off += 64;
    w.ticket_off = off;
    off = align_up(off + sizeof(uint32_t) * kMaxPasses, 256);
    w.hist_off = off;
    off = align_up(off + sizeof(uint32_t) * (size_t)B * w.passes * kRadix, 256);
    w.status_off = off;
    off = align_up(off + sizeof(uint32_t) * (size_t)w.passes * B * w.tiles * kRadix, 256);
    w.ctl_end = off;
    size_t arr = align_up(sizeof(uint32_t) * (size_t)B * N, 256);
    w.keys_a_off = off; off += arr;
    w.keys_b_off = off; off += arr;
    w.vals_a_off = off; off += arr;
    w.vals_b_off = off; off += arr;
    w.total = off;
    return w;
}

// ---- grid bounds -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t enc_ordered(float x)
{
    uint32_t u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_ordered(uint32_t e)
{
    return __uint_as_float((e & 0x80000000u) ? (e & 0x7fffffffu) : ~e);
}

// acc[b][k][0] = max over particles of ~enc(x) (i.e. the minimum), acc[b][k][1] = max of enc(x).
__global__ void __launch_bounds__(256) k_bounds_partial(const float* __restrict__ locs, uint32_t* acc,
                                                        int N, int D)
{
    __shared__ uint32_t s_acc[2 * SPNB_MAXD];
    const int b = blockIdx.y;
    if (threadIdx.x < 2 * D) s_acc[threadIdx.x] = 0;
    __syncthreads();
    const long long total = (long long)N * D;
    const long long T = (long long)gridDim.x * blockDim.x;
    const long long S = (T / D) * D;  // stride is a multiple of D so each thread keeps one coordinate
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < S) {
        const float* p = locs + (size_t)b * total;
        uint32_t mn = 0, mx = 0;
        for (long long e = t; e < total; e += S) {
            uint32_t v = enc_ordered(p[e]);
            mx = max(mx, v);
            mn = max(mn, ~v);
        }
        const int k = (int)(t % D);
        if (mx | mn) {
            atomicMax(&s_acc[2 * k], mn);
            atomicMax(&s_acc[2 * k + 1], mx);
        }
    }
    __syncthreads();
    if (threadIdx.x < 2 * D && s_acc[threadIdx.x])
        atomicMax(&acc[(size_t)b * 2 * D + threadIdx.x], s_acc[threadIdx.x]);
}

__global__ void k_bounds_final(const uint32_t* acc, float* low, float* grid_dims, int B, int D,
                               float radius, float max_grid_dim)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * D) return;
    const float lo = dec_ordered(~acc[2 * i]);
    const float hi = dec_ordered(acc[2 * i + 1]);
    float ext = (hi - lo) / radius;
    ext = fminf(fmaxf(ext, 0.0f), max_grid_dim);
    const float gd = ceilf(ext);
    const float center = (lo + hi) / 2;
    grid_dims[i] = gd;
    low[i] = center - gd * radius / 2;
}

// ---- scene control: key range and digit split -------------------------------------------------
__device__ __forceinline__ int scene_ncells(const float* gd, int D)
{
    long long n = 1;
    for (int k = 0; k < D; ++k) {
        long long g = (long long)gd[k];
        if (g <= 0) return 0;
        n *= g;
        if (n > 0x3fffffff) return 0x3fffffff;
    }
    return (int)n;
}

__global__ void k_scene_setup(const float* grid_dims, SceneCtl* ctl, int B, int D, int passes)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    SceneCtl c;
    c.ncells = scene_ncells(grid_dims + b * D, D);
    int bits = 1;
    while (bits < 31 && (1u << bits) <= (uint32_t)c.ncells) ++bits;
    c.bits = bits;
    const int base = bits / passes, rem = bits % passes;
    int sh = 0;
    for (int p = 0; p < kMaxPasses; ++p) {
        int w = p < passes ? base + (p < rem ? 1 : 0) : 0;
        c.shift[p] = sh;
        c.width[p] = w;
        sh += w;
    }
    ctl[b] = c;
}

// ---- keys + digit histograms for every pass -----------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(kSortThreads)
k_keys_hist(const float* __restrict__ locs, const float* __restrict__ low,
            const float* __restrict__ grid_dims, const SceneCtl* __restrict__ ctl,
            uint32_t* __restrict__ keys, uint32_t* hist, int N, int ndims, float edge, int passes)
{
    const int D = DT > 0 ? DT : ndims;
    __shared__ uint32_t s_hist[kMaxPasses][kRadix];
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < kMaxPasses * kRadix; i += blockDim.x) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const SceneCtl c = ctl[b];
    float lo[DT > 0 ? DT : SPNB_MAXD], gd[DT > 0 ? DT : SPNB_MAXD];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        lo[k] = low[b * D + k];
        gd[k] = grid_dims[b * D + k];
    }
    const int base = blockIdx.x * kSortTile;
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const int i = base + r * kSortThreads + threadIdx.x;
        if (i < N) {
            const float* x = locs + ((size_t)b * N + i) * D;
            int h = 0;
#pragma unroll
            for (int k = 0; k < D; ++k) h += hash_term(grid_coord_of(x[k], lo[k], edge), gd, k, D);
            const uint32_t key = (h < 0 || h >= c.ncells) ? (uint32_t)c.ncells : (uint32_t)h;
            keys[(size_t)b * N + i] = key;
            for (int p = 0; p < passes; ++p)
                atomicAdd(&s_hist[p][(key >> c.shift[p]) & ((1u << c.width[p]) - 1)], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * kRadix; i += blockDim.x) {
        const uint32_t v = (&s_hist[0][0])[i];
        if (v) atomicAdd(&hist[(size_t)b * passes * kRadix + i], v);
    }
}

// ---- one onesweep pass --------------------------------------------------------------------------
// keys_in/vals_in -> keys_out/vals_out, stable by digit `pass`.  vals_in == NULL means "value =
// position" (first pass); idxs_out != NULL means the values are written as float (last pass).
__global__ void __launch_bounds__(kSortThreads)
k_onesweep(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
           uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
           float* __restrict__ idxs_out, const SceneCtl* __restrict__ ctl,
           const uint32_t* __restrict__ hist, uint32_t* status, uint32_t* ticket, int* err, int B,
           int N, int tiles, int passes, int pass)
{
    constexpr int kWarps = kSortThreads / 32;
    __shared__ uint32_t s_warp[kWarps][kRadix];  // per-warp digit counts, then exclusive over warps
    __shared__ uint32_t s_base[kRadix];          // first output slot of each digit for this tile
    __shared__ uint32_t s_wsum[kWarps];
    __shared__ uint32_t s_tile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(&ticket[pass], 1u);
    for (int i = tid; i < kWarps * kRadix; i += kSortThreads) (&s_warp[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile_g = s_tile;
    const int b = tile_g / tiles, t = tile_g % tiles;
    const int shift = ctl[b].shift[pass];
    const uint32_t mask = (1u << ctl[b].width[pass]) - 1;
    const size_t sb = (size_t)b * N;

    uint32_t key[kSortItems], rank[kSortItems];
    const int first = t * kSortTile + warp * 32 * kSortItems + lane;
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const int i = first + r * 32;
        key[r] = i < N ? keys_in[sb + i] : 0xffffffffu;
    }
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const bool valid = first + r * 32 < N;
        const uint32_t digit = valid ? ((key[r] >> shift) & mask) : (uint32_t)kRadix;
        const uint32_t peers = __match_any_sync(0xffffffffu, digit);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (valid && lane == leader) {
            old = s_warp[warp][digit];
            s_warp[warp][digit] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[r] = old + __popc(peers & lanemask_lt());
        __syncwarp();
    }
    __syncthreads();

    // thread `tid` owns digit `tid`
    uint32_t count = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
        const uint32_t c = s_warp[w][tid];
        s_warp[w][tid] = count;
        count += c;
    }
    // decoupled look-back over the earlier tiles of this scene
    uint32_t excl = 0;
    uint32_t* st = status + ((size_t)pass * B * tiles + tile_g) * kRadix + tid;
    if (t == 0) {
        st_relaxed(st, kFlagIncl | count);
    } else {
        st_relaxed(st, kFlagAgg | count);
        const uint32_t* sp = st - kRadix;
        for (int tt = t - 1; tt >= 0; --tt, sp -= kRadix) {
            uint32_t s;
            unsigned spins = 0;
            do {
                s = ld_relaxed(sp);
            } while ((s >> 30) == 0 && ++spins < (1u << 27));
            if ((s >> 30) == 0) {  // never happens with ticket ordering; do not hang if it does
                *err = 1;
                break;
            }
            excl += s & kValMask;
            if ((s >> 30) == 2) break;
        }
        st_relaxed(st, kFlagIncl | (excl + count));
    }
    // exclusive scan of the scene's digit histogram -> first slot of each digit in the scene
    const uint32_t h = hist[((size_t)b * passes + pass) * kRadix + tid];
    uint32_t incl = h;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    uint32_t wpre = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w)
        if (w < warp) wpre += s_wsum[w];
    s_base[tid] = wpre + incl - h + excl;
    __syncthreads();

#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const int i = first + r * 32;
        if (i < N) {
            const uint32_t digit = (key[r] >> shift) & mask;
            const uint32_t pos = s_base[digit] + s_warp[warp][digit] + rank[r];
            const uint32_t val = vals_in ? vals_in[sb + i] : (uint32_t)i;
            keys_out[sb + pos] = key[r];
            if (idxs_out) idxs_out[sb + pos] = (float)val;
            else vals_out[sb + pos] = val;
        }
    }
}

// ---- reorder -------------------------------------------------------------------------------------
// One thread per output float of one tensor (z = 0: locs, z = 1: data), blockIdx.y = scene, 32-bit
// index math with a compile-time row width where it is small.  reverse == 0: out[i,:] = in[idxs[i],:]
// (coalesced writes, row gathers); else out[idxs[i],:] = in[i,:].
template <int WT>
__global__ void __launch_bounds__(256)
k_reorder(const float* __restrict__ locs, const float* __restrict__ data,
          const float* __restrict__ idxs, float* __restrict__ nlocs, float* __restrict__ ndata, int N,
          int D, int C, int reverse)
{
    const bool is_loc = blockIdx.z == 0;
    const int W = WT > 0 ? WT : (is_loc ? D : C);
    const float* __restrict__ in = (is_loc ? locs : data) + (size_t)blockIdx.y * N * W;
    float* __restrict__ out = (is_loc ? nlocs : ndata) + (size_t)blockIdx.y * N * W;
    const float* __restrict__ ix = idxs + (size_t)blockIdx.y * N;
    const unsigned total = (unsigned)N * (unsigned)W;
    for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const unsigned row = e / (unsigned)W, col = e - row * (unsigned)W;
        const unsigned other = (unsigned)(int)ix[row];
        if (reverse) out[other * W + col] = in[e];
        else out[e] = in[other * W + col];
    }
}

// ---- cell table ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_table_clear(const float* __restrict__ grid_dims, float* __restrict__ starts,
              float* __restrict__ ends, int D, int ncells)
{
    const int b = blockIdx.y;
    int n = scene_ncells(grid_dims + b * D, D);
    if (n > ncells) n = ncells;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
        starts[(size_t)b * ncells + c] = 0.0f;
        ends[(size_t)b * ncells + c] = 0.0f;
    }
}

__global__ void __launch_bounds__(256)
k_table_fill(const uint32_t* __restrict__ keys, const float* __restrict__ grid_dims,
             float* __restrict__ starts, float* __restrict__ ends, int N, int D, int ncells)
{
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int n = scene_ncells(grid_dims + b * D, D);
    if (n > ncells) n = ncells;
    const uint32_t c = keys[(size_t)b * N + i];
    const uint32_t p = i > 0 ? keys[(size_t)b * N + i - 1] : 0xffffffffu;
    if (c != p) {
        if (c < (uint32_t)n) starts[(size_t)b * ncells + c] = (float)i;
        if (i > 0 && p < (uint32_t)n) ends[(size_t)b * ncells + p] = (float)i;
    }
    if (i == N - 1 && c < (uint32_t)n) ends[(size_t)b * ncells + c] = (float)N;
}

// ---- tile descriptors (tile_lists.cuh) -------------------------------------------------------------
// One warp per tile block of kTileQ consecutive sorted queries: the block's cells span the keys
// [cf, cl]; for every offset o of the leading D-1 grid dimensions its neighbours lie in the cells
// [cf+o-1, cl+o+1] (a superset: cells that wrap around a grid border only add candidates that fail the
// distance test or are never referenced).  Each lane turns one such cell interval into a range of the
// sorted order with two binary searches over the sorted keys; lane 0 sorts and merges the ranges.
__device__ __forceinline__ int key_lower_bound(const uint32_t* __restrict__ k, int N, long long v)
{
    int lo = 0, hi = N;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((long long)k[mid] < v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256)
k_tile_ranges(const uint32_t* __restrict__ keys, const float* __restrict__ grid_dims,
              const float* __restrict__ starts, const float* __restrict__ ends, int N, int D,
              int ncells, int ntb, TileDesc* __restrict__ descs, int* flag)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tb = blockIdx.x * 8 + warp, b = blockIdx.y;
    if (tb >= ntb) return;
    const uint32_t* k = keys + (size_t)b * N;
    const float* gd = grid_dims + b * D;
    long long used = 1;
    for (int d = 0; d < D; ++d) used *= (long long)gd[d];
    if (used > ncells) used = ncells;
    const int q0 = tb * kTileQ, q1 = min(q0 + kTileQ, N) - 1;
    const long long cf = k[q0], cl = k[q1];
    const int sy = (int)gd[D - 1];
    const int sx = D >= 3 ? sy * (int)gd[D - 2] : 0;
    int nrange = 1;
    for (int d = 1; d < D; ++d) nrange *= 3;
    int s = 0, e = 0;
    if (lane < nrange) {
        long long o = 0;
        if (D == 2) o = (long long)(lane - 1) * sy;
        if (D == 3) o = (long long)(lane % 3 - 1) * sy + (long long)(lane / 3 - 1) * sx;
        long long lo = cf + o - 1, hi = cl + o + 1;
        if (lo < 0) lo = 0;
        if (hi > used - 1) hi = used - 1;
        if (lo <= hi) {
            if (starts != nullptr && hi - lo < 64) {
                // the cell table answers both bounds with one load each unless border cells are empty
                // (empty cells read start == end == 0)
                const float* st = starts + (size_t)b * ncells;
                const float* en = ends + (size_t)b * ncells;
                long long c0 = lo, c1 = hi;
                while (c0 <= hi && !(en[c0] > st[c0])) ++c0;
                while (c1 >= c0 && !(en[c1] > st[c1])) --c1;
                if (c0 <= c1) {
                    s = (int)st[c0];
                    e = (int)en[c1];
                }
            } else {
                s = key_lower_bound(k, N, lo);
                e = key_lower_bound(k, N, hi + 1);
            }
        }
    }
    int rs[kTileMaxRanges], re[kTileMaxRanges];
#pragma unroll
    for (int r = 0; r < kTileMaxRanges; ++r) {
        rs[r] = __shfl_sync(0xffffffffu, s, r);
        re[r] = __shfl_sync(0xffffffffu, e, r);
    }
    if (lane != 0) return;
    // insertion sort by start (empty ranges last), then merge overlaps
    int n = 0;
    int ss[kTileMaxRanges], ee[kTileMaxRanges];
    for (int r = 0; r < nrange; ++r) {
        if (re[r] <= rs[r]) continue;
        int p = n++;
        while (p > 0 && ss[p - 1] > rs[r]) {
            ss[p] = ss[p - 1];
            ee[p] = ee[p - 1];
            --p;
        }
        ss[p] = rs[r];
        ee[p] = re[r];
    }
    TileDesc d;
    d.nr = 0;
    int total = 0;
    for (int r = 0; r < kTileMaxRanges; ++r) d.start[r] = d.prefix[r] = 0;
    int cur_s = 0, cur_e = -1;
    for (int r = 0; r <= n; ++r) {
        if (r < n && cur_e >= 0 && ss[r] <= cur_e) {
            if (ee[r] > cur_e) cur_e = ee[r];
            continue;
        }
        if (cur_e >= 0) {
            d.start[d.nr] = cur_s;
            d.prefix[d.nr] = total;
            total += cur_e - cur_s;
            ++d.nr;
        }
        if (r < n) {
            cur_s = ss[r];
            cur_e = ee[r];
        }
    }
    d.prefix[d.nr] = total;
    for (int r = d.nr + 1; r <= kTileMaxRanges; ++r) d.prefix[r] = total;
    d.total = total;
    for (int i = 0; i < 11; ++i) d.pad[i] = 0;
    descs[(size_t)b * ntb + tb] = d;
}

// ---- neighbour lists -----------------------------------------------------------------------------
// One warp per kQPW consecutive queries.  Consecutive queries that fall in the same grid cell (all of
// them, when the queries are the cell-sorted particles themselves) share one candidate set: the
// 3^D neighbour cells are resolved ONCE per run (cell table lookups, warp prefix sum), the
// candidates' indices and coordinates are staged in shared memory, and every query of the run then
// streams over the staged candidates 32 at a time -- distance test, ballot/popc compaction,
// coalesced row writes, -1 padding.  Cells are visited in the reference's odometer order over
// {-1,0,1}^D with dimension 0 fastest and candidates inside a cell in sorted order
// (common_funcs.h:906-943), so rows are bit-identical to the reference's, truncation included.
#ifndef SPNB_COLLIDE_QPW
#define SPNB_COLLIDE_QPW 16  // measured: 8 -> 318 us, 16 -> 304 us, 32 -> 306 us at c2 (longer same-cell runs share one candidate staging)
#endif
constexpr int kQPW = SPNB_COLLIDE_QPW;  // queries per warp (<= 32)
constexpr int kCollideWarps = 8;  // warps per block

template <int DT>
__global__ void __launch_bounds__(kCollideWarps * 32)
k_collide(const float* __restrict__ qlocs, const float* __restrict__ locs,
          const float* __restrict__ low, const float* __restrict__ grid_dims,
          const float* __restrict__ starts, const float* __restrict__ ends,
          float* __restrict__ coll, int M, int N, int ndims, int K, int ncells, float edge, float r2,
          int include_self, int* trunc_flag)
{
    constexpr int MD = DT > 0 ? DT : SPNB_MAXD;
    constexpr int CM = DT > 0 ? 256 : 64;  // staged candidates per window
    const int D = DT > 0 ? DT : ndims;
    __shared__ int s_off[kCollideWarps][33];
    __shared__ int s_start[kCollideWarps][32];
    __shared__ int s_idx[kCollideWarps][CM];
    __shared__ float s_y[kCollideWarps][MD][CM];
    __shared__ int s_found[kCollideWarps][kQPW];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int q0 = (blockIdx.x * kCollideWarps + warp) * kQPW;
    if (q0 >= M) return;
    const int nq = min(kQPW, M - q0);
    const float* gd = grid_dims + b * D;
    const float* lo = low + b * D;
    const float* sl = locs + (size_t)b * N * D;
    const float* sq = qlocs + ((size_t)b * M + q0) * D;
    float* rows = coll + ((size_t)b * M + q0) * K;
    const float* st = starts + (size_t)b * ncells;
    const float* en = ends + (size_t)b * ncells;
    int total_cells = 1;
#pragma unroll
    for (int k = 0; k < D; ++k) total_cells *= 3;
    bool truncated = false;  // a row was cut at K, or a query lies beyond the clamped grid: relation not symmetric

    int qi = 0;
    while (qi < nq) {
        // ---- the run of queries sharing the cell of query qi
        int gc[MD];
#pragma unroll
        for (int k = 0; k < D; ++k) gc[k] = grid_coord_of(sq[qi * D + k], lo[k], edge);
        bool same = lane < nq - qi;
        if (same) {
#pragma unroll
            for (int k = 0; k < D; ++k)
                same = same && grid_coord_of(sq[(qi + lane) * D + k], lo[k], edge) == gc[k];
        }
        const unsigned sm = __ballot_sync(0xffffffffu, same);
        const int run = __ffs(~sm) - 1;  // leading ones (lane 0 always matches itself)
        // A query two or more cells past the upper border of a clamped grid sees no cell at all, while the
        // border cell it was hashed into (partial_grid_hash clamps, loc2grid does not: common_funcs.h:96-119,
        // 913-914) is still scanned by its neighbours: the relation is then not symmetric.
#pragma unroll
        for (int k = 0; k < D; ++k)
            if (gd[k] > 0.0f && (float)gc[k] >= gd[k] + 1.0f) truncated = true;
        if (lane < run) s_found[warp][lane] = 0;
        __syncwarp();

        for (int cell0 = 0; cell0 < total_cells; cell0 += 32) {
            // ---- lane -> one neighbour cell of this chunk: range of sorted particles in it
            const int ci = cell0 + lane;
            int cnt = 0, cstart = 0;
            if (ci < total_cells) {
                int rem = ci, id = 0;
                bool ok = true;
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const int c = gc[k] + (rem % 3) - 1;
                    rem /= 3;
                    if (c < 0 || (float)c >= gd[k]) ok = false;
                    else id += hash_term(c, gd, k, D);
                }
                if (ok && id >= 0 && id < ncells) {
                    cstart = (int)st[id];
                    cnt = (int)en[id] - cstart;
                    if (cnt < 0) cnt = 0;
                }
            }
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += n;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            __syncwarp();
            s_off[warp][lane + 1] = incl;
            if (lane == 0) s_off[warp][0] = 0;
            s_start[warp][lane] = cstart;
            __syncwarp();

            for (int w0 = 0; w0 < total; w0 += CM) {
                // ---- stage a window of candidates: index + coordinates
                const int wn = min(CM, total - w0);
                for (int t = lane; t < wn; t += 32) {
                    const int tt = w0 + t;
                    int c = 0;  // largest c with s_off[c] <= tt
#pragma unroll
                    for (int s = 16; s > 0; s >>= 1)
                        if (s_off[warp][c + s] <= tt) c += s;
                    const int idx = s_start[warp][c] + (tt - s_off[warp][c]);
                    s_idx[warp][t] = idx;
#pragma unroll
                    for (int k = 0; k < D; ++k) s_y[warp][k][t] = sl[(size_t)idx * D + k];
                }
                __syncwarp();
                // ---- every query of the run scans the window
                for (int r = 0; r < run; ++r) {
                    int found = s_found[warp][r];
                    if (found >= K) continue;
                    float x[MD];
#pragma unroll
                    for (int k = 0; k < D; ++k) x[k] = sq[(qi + r) * D + k];
                    float* row = rows + (size_t)(qi + r) * K;
                    for (int base = 0; base < wn && found < K; base += 32) {
                        const int t = base + lane;
                        bool hit = false;
                        int idx = 0;
                        if (t < wn) {
                            float d = 0.0f;
#pragma unroll
                            for (int k = 0; k < D; ++k) {
                                const float nr = x[k] - s_y[warp][k][t];
                                d += nr * nr;
                            }
                            hit = d < r2 && (d > 0.0f || include_self);
                            idx = s_idx[warp][t];
                        }
                        const unsigned m = __ballot_sync(0xffffffffu, hit);
                        const int pos = found + __popc(m & lanemask_lt());
                        if (hit && pos < K) row[pos] = (float)idx;
                        found += __popc(m);
                    }
                    if (lane == 0) s_found[warp][r] = found;
                }
                __syncwarp();
            }
        }
        // ---- terminate / pad the rows of the run
        for (int r = 0; r < run; ++r) {
            int found = s_found[warp][r];
            if (found >= K) {
                truncated = true;
                found = K;
            }
            float* row = rows + (size_t)(qi + r) * K;
            for (int p = found + lane; p < K; p += 32) row[p] = -1.0f;
        }
        __syncwarp();
        qi += run;
    }
    if (truncated && trunc_flag && lane == 0) atomicOr(trunc_flag, 1);
}

// ---- tile lists (tile_lists.cuh) ---------------------------------------------------------------------
// One block per tile block of 64 queries, run after k_collide on the float rows it wrote: every warp
// reads the rows of 8 queries (four rows at a time, 128 bytes of each per step, up to the terminator),
// maps each neighbour index to its slot in the block's candidate ranges (TileDesc: ascending range
// starts kept in registers) and writes the 16-bit entry, the sentinel padding of the last unit and the
// list length.
constexpr int kBuildThreads = 256;

__global__ void __launch_bounds__(kBuildThreads)
k_tile_build(const float* __restrict__ coll, const TileDesc* __restrict__ descs, int N, int K, int ntb,
             int* __restrict__ tile_flag, int* __restrict__ tcounts, unsigned short* __restrict__ tlists)
{
    __shared__ TileDesc s_desc;
    __shared__ int s_fwd[kTileMaxRanges], s_len[kTileMaxRanges];  // idx -> slot: idx + s_fwd[r]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tb = blockIdx.x, b = blockIdx.y;
    const size_t tile = (size_t)b * ntb + tb;
    if (tid < (int)(sizeof(TileDesc) / sizeof(int)))
        reinterpret_cast<int*>(&s_desc)[tid] = reinterpret_cast<const int*>(descs + tile)[tid];
    __syncthreads();
    int st[kTileMaxRanges];
#pragma unroll
    for (int g = 0; g < kTileMaxRanges; ++g) st[g] = g < s_desc.nr ? s_desc.start[g] : 0x7fffffff;
    if (tid < kTileMaxRanges) {
        s_fwd[tid] = 1 + s_desc.prefix[tid] - s_desc.start[tid];
        s_len[tid] = tid < s_desc.nr ? s_desc.prefix[tid + 1] - s_desc.prefix[tid] : 0;
    }
    __syncthreads();

    bool bad = false, cut = false;
    constexpr int RPW = kTileQ / (kBuildThreads / 32);  // rows per warp
    static_assert(RPW % 4 == 0, "rows are processed four at a time");
    // The kernel is issue-bound (ncu: 86 % issue active), so a warp works on FOUR rows at once: 8 lanes per
    // row, each lane one float4 = 4 entries, i.e. 32 entries per row and step with a quarter of the
    // per-row control instructions.  The row ends at its first negative entry (common_funcs.h:476).
    const int grp = lane >> 3, sub = lane & 7;
    const bool wide = (K & 3) == 0 && (reinterpret_cast<size_t>(coll) & 15) == 0;
#pragma unroll 1
    for (int r0 = 0; r0 < RPW; r0 += 4) {
        const int ql = warp * RPW + r0 + grp;
        const int m = tb * kTileQ + ql;
        const bool live_row = m < N;
        const float* row = coll + ((size_t)b * N + (live_row ? m : 0)) * K;
        unsigned short* trow = tlists + tile_entry_off(ntb, K, b, tb, ql, 0) / 2;
        int cnt = 0;
        bool open = live_row;  // terminator not seen yet
        for (int base = 0; base < K; base += 32) {
            if (!__any_sync(0xffffffffu, open)) break;
            const int k0 = base + sub * 4;
            float f[4] = {-1.0f, -1.0f, -1.0f, -1.0f};
            if (open) {
                if (wide && k0 + 3 < K) {
                    const float4 v = *reinterpret_cast<const float4*>(row + k0);
                    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (k0 + j < K) f[j] = row[k0 + j];
                }
            }
            int nv = 0;  // length of the run of non-negative entries that starts at f[0]
#pragma unroll
            for (int j = 3; j >= 0; --j) nv = f[j] >= 0.0f ? nv + 1 : 0;
            const unsigned fullm = (__ballot_sync(0xffffffffu, nv == 4) >> (grp * 8)) & 0xffu;
            const int l0 = fullm == 0xffu ? 8 : __ffs(~fullm) - 1;         // first lane of the group with a terminator
            const int nv0 = __shfl_sync(0xffffffffu, nv, grp * 8 + (l0 & 7));
            const int take = open ? (l0 == 8 ? 32 : l0 * 4 + nv0) : 0;    // entries before the terminator
            const int padded = (take + kTileUnit - 1) & ~(kTileUnit - 1);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int e = sub * 4 + j;  // entry within this step's 32
                unsigned loc = 0;           // entries past the terminator: sentinel padding of the last unit
                if (e < take) {
                    const int idx = (int)f[j];
                    int g = 0;
#pragma unroll
                    for (int t = 1; t < kTileMaxRanges; ++t) g += idx >= st[t];
                    loc = (unsigned)(idx + s_fwd[g]);
                    if ((unsigned)(idx - st[g]) >= (unsigned)s_len[g] || loc > 0xfffu) {
                        bad = true;  // not in the block's ranges, or slot * 16 does not fit 16 bits
                        loc = 0;
                    }
                    loc <<= 4;  // entries are byte offsets into a plane of 16-byte record quarters
                }
                const int k = base + e;
                if (open && e < padded && k < K)
                    trow[(k >> 4) * 128 + ((k & 3) * 4 + ((k >> 2) & 3))] = (unsigned short)loc;  // 4x4-transposed unit
            }
            cnt += take;
            if (take < 32) open = false;
        }
        if (live_row && cnt >= K) cut = true;
        if (live_row && sub == 0) tcounts[(size_t)b * N + m] = cnt;
    }
    if (cut) atomicOr(tile_flag, 1);
    if (bad) atomicOr(tile_flag, 2);
}

// ---- launch helpers --------------------------------------------------------------------------------
static bool valid_common(int B, int N, int D, const char* fn)
{
    if (B <= 0 || N <= 0 || D <= 0 || D > SPNB_MAX_NDIM) {
        set_error("%s: bad sizes batch_size=%d N=%d ndims=%d (need >0, ndims<=%d)", fn, B, N, D,
                  SPNB_MAX_NDIM);
        return false;
    }
    if (N > (1 << 24)) {
        set_error("%s: N=%d exceeds 2^24, the float32 index limit of the API", fn, N);
        return false;
    }
    return true;
}

}  // namespace spnb

using namespace spnb;

extern "C" {

size_t spnb_hashgrid_workspace_bytes(int batch_size, int N, int ndims, int max_grid_dim)
{
    if (batch_size <= 0 || N <= 0 || ndims <= 0 || max_grid_dim <= 0) return 0;
    return ws_layout(batch_size, N, ndims, max_grid_dim).total;
}

int spnb_grid_bounds(const float* locs, int B, int N, int D, float radius, int max_grid_dim,
                     float* low, float* grid_dims, void* workspace, size_t workspace_bytes,
                     void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!valid_common(B, N, D, "spnb_grid_bounds")) return 0;
    if (!locs || !low || !grid_dims || !workspace) {
        set_error("spnb_grid_bounds: null pointer");
        return 0;
    }
    const WsLayout w = ws_layout(B, N, D, max_grid_dim);
    if (workspace_bytes < w.total) {
        set_error("spnb_grid_bounds: workspace too small (%zu < %zu)", workspace_bytes, w.total);
        return 0;
    }
    uint32_t* acc = (uint32_t*)((char*)workspace + w.minmax_off);
    cudaMemsetAsync(acc, 0, w.minmax_bytes, stream);
    int bx = cdiv((long long)N * D, 256 * 8);
    if (bx > 148 * 4) bx = 148 * 4;
    if (bx < 1) bx = 1;
    while ((long long)bx * 256 < D) ++bx;
    k_bounds_partial<<<dim3(bx, B), 256, 0, stream>>>(locs, acc, N, D);
    k_bounds_final<<<cdiv(B * D, 128), 128, 0, stream>>>(acc, low, grid_dims, B, D, radius,
                                                         (float)max_grid_dim);
    count_launches(2);
    return check_launch("spnb_grid_bounds") ? 1 : 0;
}

int spnb_hashgrid_order(const float* locs, const float* low, const float* grid_dims, float* cellIDs,
                        float* idxs, void* workspace, size_t workspace_bytes, int B, int N, int D,
                        float cellEdge, int max_grid_dim, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!valid_common(B, N, D, "spnb_hashgrid_order")) return 0;
    if (!locs || !low || !grid_dims || !cellIDs || !idxs || !workspace) {
        set_error("spnb_hashgrid_order: null pointer");
        return 0;
    }
    const WsLayout w = ws_layout(B, N, D, max_grid_dim);
    if (workspace_bytes < w.total) {
        set_error("spnb_hashgrid_order: workspace too small (%zu < %zu)", workspace_bytes, w.total);
        return 0;
    }
    char* ws = (char*)workspace;
    SceneCtl* ctl = (SceneCtl*)(ws + w.ctl_off);
    int* err = (int*)(ws + w.err_off);
    uint32_t* ticket = (uint32_t*)(ws + w.ticket_off);
    uint32_t* hist = (uint32_t*)(ws + w.hist_off);
    uint32_t* status = (uint32_t*)(ws + w.status_off);
    uint32_t* kbuf[2] = {(uint32_t*)(ws + w.keys_a_off), (uint32_t*)(ws + w.keys_b_off)};
    uint32_t* vbuf[2] = {(uint32_t*)(ws + w.vals_a_off), (uint32_t*)(ws + w.vals_b_off)};

    cudaMemsetAsync(ws + w.ctl_off, 0, w.ctl_end - w.ctl_off, stream);
    k_scene_setup<<<cdiv(B, 128), 128, 0, stream>>>(grid_dims, ctl, B, D, w.passes);
    const dim3 tg(w.tiles, B);
    switch (D) {
    case 1: k_keys_hist<1><<<tg, kSortThreads, 0, stream>>>(locs, low, grid_dims, ctl, kbuf[0], hist, N, D, cellEdge, w.passes); break;
    case 2: k_keys_hist<2><<<tg, kSortThreads, 0, stream>>>(locs, low, grid_dims, ctl, kbuf[0], hist, N, D, cellEdge, w.passes); break;
    case 3: k_keys_hist<3><<<tg, kSortThreads, 0, stream>>>(locs, low, grid_dims, ctl, kbuf[0], hist, N, D, cellEdge, w.passes); break;
    default: k_keys_hist<0><<<tg, kSortThreads, 0, stream>>>(locs, low, grid_dims, ctl, kbuf[0], hist, N, D, cellEdge, w.passes); break;
    }
    for (int p = 0; p < w.passes; ++p) {
        const bool last = p == w.passes - 1;
        const uint32_t* kin = kbuf[p & 1];
        const uint32_t* vin = p == 0 ? nullptr : vbuf[p & 1];
        uint32_t* kout = last ? (uint32_t*)cellIDs : kbuf[(p + 1) & 1];
        uint32_t* vout = last ? nullptr : vbuf[(p + 1) & 1];
        k_onesweep<<<w.tiles * B, kSortThreads, 0, stream>>>(kin, vin, kout, vout,
                                                             last ? idxs : nullptr, ctl, hist, status,
                                                             ticket, err, B, N, w.tiles, w.passes, p);
    }
    count_launches(2 + w.passes);
    return check_launch("spnb_hashgrid_order") ? 1 : 0;
}

int spnb_reorder_data(const float* locs, const float* data, const float* idxs, float* nlocs,
                      float* ndata, int B, int N, int D, int C, int reverse, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (B <= 0 || N <= 0 || D <= 0) {
        set_error("spnb_reorder_data: bad sizes");
        return 0;
    }
    if (!locs || !idxs || !nlocs || (data && !ndata)) {
        set_error("spnb_reorder_data: null pointer");
        return 0;
    }
    if (!data) C = 0;
    if ((long long)N * (D > C ? D : C) >= (1ll << 32)) {
        set_error("spnb_reorder_data: N * row width exceeds 2^32");
        return 0;
    }
    // one tensor per launch (used when the row widths of locs and data differ): the kernel reads its
    // tensor from the `locs` slot (blockIdx.z == 0)
    auto launch = [&](int W, const float* in, float* out) {
        int blocks = cdiv((long long)N * W, 256 * 4);
        if (blocks < 1) blocks = 1;
        const dim3 grid(blocks, B, 1);
        switch (W) {
        case 1: k_reorder<1><<<grid, 256, 0, stream>>>(in, nullptr, idxs, out, nullptr, N, W, 0, reverse); break;
        case 2: k_reorder<2><<<grid, 256, 0, stream>>>(in, nullptr, idxs, out, nullptr, N, W, 0, reverse); break;
        case 3: k_reorder<3><<<grid, 256, 0, stream>>>(in, nullptr, idxs, out, nullptr, N, W, 0, reverse); break;
        case 4: k_reorder<4><<<grid, 256, 0, stream>>>(in, nullptr, idxs, out, nullptr, N, W, 0, reverse); break;
        default: k_reorder<0><<<grid, 256, 0, stream>>>(in, nullptr, idxs, out, nullptr, N, W, 0, reverse); break;
        }
    };
    if (C == D && D <= 4) {
        // both tensors in one launch (gridDim.z = 2)
        int blocks = cdiv((long long)N * D, 256 * 4);
        if (blocks < 1) blocks = 1;
        const dim3 grid(blocks, B, 2);
        switch (D) {
        case 1: k_reorder<1><<<grid, 256, 0, stream>>>(locs, data, idxs, nlocs, ndata, N, D, C, reverse); break;
        case 2: k_reorder<2><<<grid, 256, 0, stream>>>(locs, data, idxs, nlocs, ndata, N, D, C, reverse); break;
        case 3: k_reorder<3><<<grid, 256, 0, stream>>>(locs, data, idxs, nlocs, ndata, N, D, C, reverse); break;
        default: k_reorder<4><<<grid, 256, 0, stream>>>(locs, data, idxs, nlocs, ndata, N, D, C, reverse); break;
        }
    } else {
        launch(D, locs, nlocs);
        if (C > 0) {
            launch(C, data, ndata);
            count_launches(1);
        }
    }
    count_launches(1);
    return check_launch("spnb_reorder_data") ? 1 : 0;
}

int spnb_compute_collisions(const float* qlocs, const float* locs, const float* low,
                            const float* grid_dims, const float* cellIDs, float* cellStarts,
                            float* cellEnds, float* collisions, int B, int M, int N, int D, int K,
                            int ncells, float cellEdge, float radius, int include_self,
                            int* trunc_flag, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!valid_common(B, N, D, "spnb_compute_collisions")) return 0;
    if (M <= 0 || K <= 0 || ncells <= 0) {
        set_error("spnb_compute_collisions: bad sizes M=%d max_collisions=%d ncells=%d", M, K, ncells);
        return 0;
    }
    if (!qlocs || !locs || !low || !grid_dims || !cellIDs || !cellStarts || !cellEnds || !collisions) {
        set_error("spnb_compute_collisions: null pointer");
        return 0;
    }
    k_table_clear<<<dim3(148, B), 256, 0, stream>>>(grid_dims, cellStarts, cellEnds, D, ncells);
    k_table_fill<<<dim3(cdiv(N, 256), B), 256, 0, stream>>>((const uint32_t*)cellIDs, grid_dims,
                                                            cellStarts, cellEnds, N, D, ncells);
    const float r2 = radius * radius;
    const dim3 blocks(cdiv(M, kCollideWarps * kQPW), B);
#define SPNB_COLLIDE(DT)                                                                          \
    k_collide<DT><<<blocks, kCollideWarps * 32, 0, stream>>>(                                      \
        qlocs, locs, low, grid_dims, cellStarts, cellEnds, collisions, M, N, D, K, ncells, cellEdge, \
        r2, include_self, trunc_flag)
    switch (D) {
    case 1: SPNB_COLLIDE(1); break;
    case 2: SPNB_COLLIDE(2); break;
    case 3: SPNB_COLLIDE(3); break;
    default: SPNB_COLLIDE(0); break;
    }
#undef SPNB_COLLIDE
    count_launches(3);
    return check_launch("spnb_compute_collisions") ? 1 : 0;
}

size_t spnb_tile_lists_bytes(int batch_size, int N, int ndims, int max_collisions)
{
    if (batch_size <= 0 || !tile_lists_supported(N, ndims, max_collisions)) return 0;
    return tile_layout(batch_size, N, max_collisions).total;
}

int spnb_build_tile_lists(const float* cellIDs, const float* grid_dims, const float* cellStarts,
                          const float* cellEnds, const float* collisions, int B, int N, int D, int K,
                          int ncells, void* tile_lists, size_t tile_lists_bytes, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!valid_common(B, N, D, "spnb_build_tile_lists")) return 0;
    if (!cellIDs || !grid_dims || !collisions || !tile_lists) {
        set_error("spnb_build_tile_lists: null pointer");
        return 0;
    }
    if (!tile_lists_supported(N, D, K)) {
        set_error("spnb_build_tile_lists: needs ndims <= %d and max_collisions a multiple of %d",
                  kTileMaxNdim, kTileUnit);
        return 0;
    }
    const TileLayout tl = tile_layout(B, N, K);
    if (tile_lists_bytes < tl.total) {
        set_error("spnb_build_tile_lists: tile buffer too small (%zu < %zu)", tile_lists_bytes, tl.total);
        return 0;
    }
    char* tb = (char*)tile_lists;
    int* tflag = (int*)tb;
    TileDesc* descs = (TileDesc*)(tb + tl.desc_off);
    cudaMemsetAsync(tflag, 0, 128, stream);
    k_tile_ranges<<<dim3(cdiv(tl.ntb, 8), B), 256, 0, stream>>>((const uint32_t*)cellIDs, grid_dims,
                                                                cellEnds ? cellStarts : nullptr, cellEnds, N, D,
                                                                ncells, tl.ntb, descs, tflag);
    k_tile_build<<<dim3(tl.ntb, B), kBuildThreads, 0, stream>>>(
        collisions, descs, N, K, tl.ntb, tflag, (int*)(tb + tl.cnt_off), (unsigned short*)(tb + tl.list_off));
    count_launches(2);
    return check_launch("spnb_build_tile_lists") ? 1 : 0;
}

}  // extern "C"
