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
#include "list_walk.cuh"
#include "spnb_common.cuh"
#include "tile_lists.cuh"

namespace spnb {

constexpr int kMaxPasses = 4;
constexpr int kRadix = 256;
constexpr int kSortThreads = 256;
#ifndef SPNB_SORT_ITEMS
#define SPNB_SORT_ITEMS 8
#endif
constexpr int kSortItems = SPNB_SORT_ITEMS;
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
    // the tile's keys and values in digit order: the scatter then writes runs of consecutive addresses per digit
    // (a key-by-key scatter from registers costs one L2 sector transaction per 4-byte element)
    __shared__ uint32_t s_loc[kRadix];  // first tile-local slot of each digit
    __shared__ uint32_t s_key[kSortTile], s_val[kSortTile];

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
        // look back kLook predecessors per round: their status words are requested together, then consumed in order
        // (a serial walk pays one L2 round trip per predecessor: 41 % of this kernel's stall samples at 2^24 keys)
        constexpr int kLook = 4;
        const uint32_t* sp = st - kRadix;
        int tt = t - 1;
        unsigned spins = 0;
        while (tt >= 0) {
            uint32_t sv[kLook];
#pragma unroll
            for (int u = 0; u < kLook; ++u) sv[u] = tt - u >= 0 ? ld_relaxed(sp - (size_t)u * kRadix) : 0u;
            int used = 0;
            bool found = false;
#pragma unroll
            for (int u = 0; u < kLook; ++u) {
                if (!found && used == u && tt - u >= 0 && (sv[u] >> 30) != 0) {
                    excl += sv[u] & kValMask;
                    ++used;
                    if ((sv[u] >> 30) == 2) found = true;
                }
            }
            if (found) break;
            tt -= used;
            sp -= (size_t)used * kRadix;
            if (used == 0 && ++spins >= (1u << 27)) {  // never happens with ticket ordering; do not hang if it does
                *err = 1;
                break;
            }
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
    // tile-local exclusive scan of the tile's digit counts
    uint32_t linc = count;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, linc, o);
        if (lane >= o) linc += n;
    }
    __syncthreads();  // s_wsum (the histogram scan) has been read by everybody
    if (lane == 31) s_wsum[warp] = linc;
    __syncthreads();
    uint32_t lpre = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w)
        if (w < warp) lpre += s_wsum[w];
    s_loc[tid] = lpre + linc - count;
    __syncthreads();

    // keys and values to their tile-local sorted slots
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const int i = first + r * 32;
        if (i < N) {
            const uint32_t digit = (key[r] >> shift) & mask;
            const uint32_t lp = s_loc[digit] + s_warp[warp][digit] + rank[r];
            s_key[lp] = key[r];
            s_val[lp] = vals_in ? vals_in[sb + i] : (uint32_t)i;
        }
    }
    __syncthreads();
    // out: consecutive threads take consecutive slots; equal digits go to consecutive addresses
    const int nvalid = min(kSortTile, N - t * kSortTile);
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const int l = r * kSortThreads + tid;
        if (l < nvalid) {
            const uint32_t k = s_key[l];
            const uint32_t digit = (k >> shift) & mask;
            const uint32_t pos = s_base[digit] + ((uint32_t)l - s_loc[digit]);
            keys_out[sb + pos] = k;
            if (idxs_out) idxs_out[sb + pos] = (float)s_val[l];
            else vals_out[sb + pos] = s_val[l];
        }
    }
}

// ---- reorder -------------------------------------------------------------------------------------
// reverse == 0: out[i,:] = in[idxs[i],:]   (row gathers, coalesced writes)
// reverse != 0: out[idxs[i],:] = in[i,:]   (coalesced reads, row scatters)
// blockIdx.y = scene, blockIdx.z = tensor (0: locs, 1: data).
//
// Rows of 1..4 floats (positions, velocities, scalars): one thread per row; the contiguous side of the copy
// goes through a per-warp shared-memory transpose so that it moves as 128-bit accesses of a whole warp (32
// rows of W floats are 8*W float4), the permuted side is W scalar accesses per row.  With `pos4` the
// gathered position rows are also written as a float4 plane (x, y, z, 0): the layout TMA bulk copies and
// LDS.128 gathers want (k_collide_tiles and the ConvSP tile kernels stage it).
template <int W>
__global__ void __launch_bounds__(256)
k_reorder_rows(const float* __restrict__ locs, const float* __restrict__ data, const float* __restrict__ idxs,
               float* __restrict__ nlocs, float* __restrict__ ndata, float4* __restrict__ pos4, int N, int reverse)
{
    __shared__ __align__(16) float s_t[8][32 * W];
    const bool is_loc = blockIdx.z == 0;
    const size_t sb = (size_t)blockIdx.y * N;
    const float* __restrict__ in = (is_loc ? locs : data) + sb * W;
    float* __restrict__ out = (is_loc ? nlocs : ndata) + sb * W;
    const float* __restrict__ ix = idxs + sb;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i0 = (blockIdx.x * 8 + warp) * 32;  // first row of this warp
    if (i0 >= N) return;
    const int i = i0 + lane;
    const bool live = i < N;
    const int other = live ? (int)ix[i] : 0;
    // the contiguous side can move as float4 when the warp's 32 rows are all there and 16-byte aligned
    const bool vec = i0 + 32 <= N && ((reinterpret_cast<size_t>(in) | reinterpret_cast<size_t>(out)) & 15) == 0 &&
                     (((size_t)N * W) & 3) == 0;
    float* st = s_t[warp];
    float r[W];
    if (!reverse) {
#pragma unroll
        for (int k = 0; k < W; ++k) r[k] = live ? in[(size_t)other * W + k] : 0.0f;
        if (pos4 != nullptr && is_loc && live)
            pos4[sb + i] = make_float4(r[0], W > 1 ? r[1] : 0.0f, W > 2 ? r[2] : 0.0f, 0.0f);
        if (vec) {
#pragma unroll
            for (int k = 0; k < W; ++k) st[lane * W + k] = r[k];
            __syncwarp();
#pragma unroll
            for (int t = lane; t < 8 * W; t += 32)
                reinterpret_cast<float4*>(out + (size_t)i0 * W)[t] = reinterpret_cast<const float4*>(st)[t];
        } else if (live) {
#pragma unroll
            for (int k = 0; k < W; ++k) out[(size_t)i * W + k] = r[k];
        }
    } else {
        if (vec) {
#pragma unroll
            for (int t = lane; t < 8 * W; t += 32)
                reinterpret_cast<float4*>(st)[t] = reinterpret_cast<const float4*>(in + (size_t)i0 * W)[t];
            __syncwarp();
#pragma unroll
            for (int k = 0; k < W; ++k) r[k] = st[lane * W + k];
        } else {
#pragma unroll
            for (int k = 0; k < W; ++k) r[k] = live ? in[(size_t)i * W + k] : 0.0f;
        }
        if (live) {
#pragma unroll
            for (int k = 0; k < W; ++k) out[(size_t)other * W + k] = r[k];
        }
    }
}

// Wider rows: one thread per 16-byte piece of a row when W is a multiple of 4 (VEC), else per float.
template <bool VEC>
__global__ void __launch_bounds__(256)
k_reorder_wide(const float* __restrict__ in_, const float* __restrict__ idxs, float* __restrict__ out_, int N,
               int W, int reverse)
{
    const int P = VEC ? W / 4 : W;  // pieces per row
    const size_t sb = (size_t)blockIdx.y * N;
    const float* __restrict__ ix = idxs + sb;
    const unsigned total = (unsigned)N * (unsigned)P;
    for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const unsigned row = e / (unsigned)P, col = e - row * (unsigned)P;
        const unsigned other = (unsigned)(int)ix[row];
        const size_t a = (sb + (reverse ? row : other)) * P + col;   // source piece
        const size_t d = (sb + (reverse ? other : row)) * P + col;   // destination piece
        if (VEC) reinterpret_cast<float4*>(out_)[d] = reinterpret_cast<const float4*>(in_)[a];
        else out_[d] = in_[a];
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

// ---- neighbour lists -----------------------------------------------------------------------------
// One warp per kQPW consecutive queries.  Consecutive queries that fall in the same grid cell (all of
// them, when the queries are the cell-sorted particles themselves) share one candidate set: the
// 3^D neighbour cells are resolved ONCE per run (cell table lookups, warp prefix sum), the
// candidates' indices and coordinates are staged in shared memory, and every query of the run then
// streams over the staged candidates 32 at a time -- distance test, ballot/popc compaction,
// coalesced row writes, -1 padding.  Cells are visited in the reference's odometer order over
// {-1,0,1}^D with dimension 0 fastest and candidates inside a cell in sorted order
// (common_funcs.h:906-943), so rows are bit-identical to the reference's, truncation included.
//
// This is the general routine (separate query locations, any ndims, any K).  When the particles are their
// own queries k_collide_tiles below produces the same rows from shared-memory tiles, together with the
// compact tile lists.
#ifndef SPNB_COLLIDE_QPW
#define SPNB_COLLIDE_QPW 16  // measured: 8 -> 318 us, 16 -> 304 us, 32 -> 306 us at c2 (longer same-cell runs share one candidate staging)
#endif
constexpr int kQPW = SPNB_COLLIDE_QPW;  // queries per warp (<= 32)
constexpr int kCollideWarps = 8;  // warps per block

template <int MD, int CM>
struct CollideSmem {
    int off[kCollideWarps][33];
    int start[kCollideWarps][32];
    int idx[kCollideWarps][CM];
    float y[kCollideWarps][MD][CM];
    int found[kCollideWarps][kQPW];
};

// Rows of the nq queries starting at sq (their rows start at `rows`), for one warp.  Returns true when the
// neighbour relation may not be symmetric (a row was cut at K, or a query lies beyond a clamped grid).
template <int DT, int MD, int CM>
__device__ __forceinline__ bool collide_rows_warp(CollideSmem<MD, CM>& sm, const float* __restrict__ sq,
                                                  const float* __restrict__ sl, float* __restrict__ rows, int nq,
                                                  const float* __restrict__ lo, const float* __restrict__ gd,
                                                  const float* __restrict__ st, const float* __restrict__ en, int ndims,
                                                  int K, int ncells, float edge, float r2, int include_self)
{
    const int D = DT > 0 ? DT : ndims;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
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
        const unsigned smk = __ballot_sync(0xffffffffu, same);
        const int run = __ffs(~smk) - 1;  // leading ones (lane 0 always matches itself)
        // A query two or more cells past the upper border of a clamped grid sees no cell at all, while the
        // border cell it was hashed into (partial_grid_hash clamps, loc2grid does not: common_funcs.h:96-119,
        // 913-914) is still scanned by its neighbours: the relation is then not symmetric.
#pragma unroll
        for (int k = 0; k < D; ++k)
            if (gd[k] > 0.0f && (float)gc[k] >= gd[k] + 1.0f) truncated = true;
        if (lane < run) sm.found[warp][lane] = 0;
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
            sm.off[warp][lane + 1] = incl;
            if (lane == 0) sm.off[warp][0] = 0;
            sm.start[warp][lane] = cstart;
            __syncwarp();

            for (int w0 = 0; w0 < total; w0 += CM) {
                // ---- stage a window of candidates: index + coordinates
                const int wn = min(CM, total - w0);
                for (int t = lane; t < wn; t += 32) {
                    const int tt = w0 + t;
                    int c = 0;  // largest c with off[c] <= tt
#pragma unroll
                    for (int s_ = 16; s_ > 0; s_ >>= 1)
                        if (sm.off[warp][c + s_] <= tt) c += s_;
                    const int idx = sm.start[warp][c] + (tt - sm.off[warp][c]);
                    sm.idx[warp][t] = idx;
#pragma unroll
                    for (int k = 0; k < D; ++k) sm.y[warp][k][t] = sl[(size_t)idx * D + k];
                }
                __syncwarp();
                // ---- every query of the run scans the window
                for (int r = 0; r < run; ++r) {
                    int found = sm.found[warp][r];
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
                                const float nr = x[k] - sm.y[warp][k][t];
                                d += nr * nr;
                            }
                            hit = d < r2 && (d > 0.0f || include_self);
                            idx = sm.idx[warp][t];
                        }
                        const unsigned m = __ballot_sync(0xffffffffu, hit);
                        const int pos = found + __popc(m & lanemask_lt());
                        if (hit && pos < K) row[pos] = (float)idx;
                        found += __popc(m);
                    }
                    if (lane == 0) sm.found[warp][r] = found;
                }
                __syncwarp();
            }
        }
        // ---- terminate / pad the rows of the run
        for (int r = 0; r < run; ++r) {
            int found = sm.found[warp][r];
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
    return truncated;
}

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
    __shared__ CollideSmem<MD, CM> sm;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int q0 = (blockIdx.x * kCollideWarps + warp) * kQPW;
    if (q0 >= M) return;
    const bool truncated = collide_rows_warp<DT, MD, CM>(
        sm, qlocs + ((size_t)b * M + q0) * D, locs + (size_t)b * N * D, coll + ((size_t)b * M + q0) * K,
        min(kQPW, M - q0), low + b * D, grid_dims + b * D, starts + (size_t)b * ncells, ends + (size_t)b * ncells,
        ndims, K, ncells, edge, r2, include_self);
    if (truncated && trunc_flag && lane == 0) atomicOr(trunc_flag, 1);
}

// ---- neighbour lists + tile lists in one kernel (particles are their own queries, ndims <= 3) -------------
// One CTA per tile block of kTileQ = 64 consecutive sorted queries (tile_lists.cuh):
//  1. warp 0 derives the block's candidate ranges (TileDesc): the block's cells span the keys [cf, cl]; for every
//     offset o of the leading D-1 grid dimensions its neighbours lie in the cells [cf+o-1, cl+o+1] (a superset:
//     cells that wrap around a grid border only add candidates that fail the distance test).  Each lane turns one
//     such cell interval into a range of the sorted order through the cell table; the ranges are ranked and
//     merged with shuffles;
//  2. the float4 position plane of those ranges is staged into shared memory with TMA bulk copies (one
//     cp.async.bulk per range, completion on an mbarrier), so every candidate is an LDS.128 away;
//  3. each warp takes 8 queries.  Per run of queries sharing a cell the 3^D neighbour cells are resolved once
//     (lane per cell), turned into tile slots and loaded -- 256 candidates at a time -- into REGISTERS (slot and
//     coordinates, 8 per lane); every query of the run then tests them with the reference's exact predicate
//     and order (odometer over the cells with dimension 0 fastest, ascending index inside a cell:
//     common_funcs.h:906-943) and appends the hits' slots to its row in shared memory (ballot / popc);
//  4. the block then ranks its queries by list length and writes, from the same hits, the float rows
//     (coalesced 128-byte lines incl. the -1 padding) and the 16-bit tile rows of every octile.
// A block whose ranges do not fit the staged tile reads its candidates from global memory; only a block with more
// than kTileMaxSlots candidates (a dense clump) computes its rows with collide_rows_warp and raises the tile flag.
constexpr int kCtThreads = 256;
#ifndef SPNB_CT_MINB
#define SPNB_CT_MINB 3  // CTAs per SM the register allocation aims at (80 registers)
#endif
constexpr int kCtWin = 256;  // candidates held in registers per window (8 per lane)

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

// Executed by one full warp; fills nr / total / start / prefix of *d (shared memory).
__device__ __forceinline__ void tile_desc_warp(const uint32_t* __restrict__ k, const float* __restrict__ gd,
                                               const float* __restrict__ st, const float* __restrict__ en, int N, int D,
                                               int ncells, int q0, int nq, TileDesc* d)
{
    const int lane = threadIdx.x & 31;
    long long used = 1;
    for (int i = 0; i < D; ++i) used *= (long long)gd[i];
    if (used > ncells) used = ncells;
    const long long cf = k[q0], cl = k[q0 + nq - 1];
    const int sy = (int)gd[D - 1];
    const int sx = D >= 3 ? sy * (int)gd[D - 2] : 0;
    int nrange = 1;
    for (int i = 1; i < D; ++i) nrange *= 3;
    int s = 0, e = 0;
    if (lane < nrange) {
        long long o = 0;
        if (D == 2) o = (long long)(lane - 1) * sy;
        if (D == 3) o = (long long)(lane % 3 - 1) * sy + (long long)(lane / 3 - 1) * sx;
        long long lo = cf + o - 1, hi = cl + o + 1;
        if (lo < 0) lo = 0;
        if (hi > used - 1) hi = used - 1;
        if (lo <= hi) {
            if (hi - lo < 64) {
                // the cell table answers both bounds with one load each unless border cells are empty
                // (empty cells read start == end == 0)
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
    if (e > N) e = N;
    const bool have = lane < nrange && e > s;
    // rank of my range among the non-empty ones, by start
    int rank = 0;
#pragma unroll
    for (int r = 0; r < kTileMaxRanges; ++r) {
        const int os = __shfl_sync(0xffffffffu, s, r);
        const bool oh = __shfl_sync(0xffffffffu, (int)have, r) != 0;
        if (oh && (os < s || (os == s && r < lane))) ++rank;
    }
    const unsigned hm = __ballot_sync(0xffffffffu, have);
    const int n = __popc(hm);
    // gather the sorted ranges into lanes 0..n-1
    int ss = 0, ee = 0;
#pragma unroll
    for (int r = 0; r < kTileMaxRanges; ++r) {
        const int os = __shfl_sync(0xffffffffu, s, r), oe = __shfl_sync(0xffffffffu, e, r);
        const int orank = __shfl_sync(0xffffffffu, rank, r);
        const bool oh = (hm >> r) & 1u;
        if (oh && orank == lane) {
            ss = os;
            ee = oe;
        }
    }
    // merge overlapping / touching ranges: running maximum of the ends, a range starts a new group when its
    // start lies beyond every earlier end
    int me = ee;  // inclusive prefix maximum of ends over lanes 0..lane
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, me, o);
        if (lane >= o && lane < n) me = max(me, v);
    }
    const int prev_max = __shfl_up_sync(0xffffffffu, me, 1);
    const bool head = lane < n && (lane == 0 || ss > prev_max);
    const unsigned heads = __ballot_sync(0xffffffffu, head);
    const int nr = __popc(heads);
    const int grp = __popc(heads & (0xffffffffu >> (31 - lane))) - 1;  // group of my range (lane < n)
    // end of a group = prefix maximum at its last member = value at the lane before the next head (or n-1)
    int gs = 0, ge = 0;  // lane g < nr: merged range g
    {
        // lane g finds the g-th head
        unsigned hmask = heads;
        int hl = 0;
        for (int i = 0; i <= lane && i < nr; ++i) {
            hl = __ffs(hmask) - 1;
            hmask &= hmask - 1;
        }
        const int next = (lane < nr && hmask) ? __ffs(hmask) - 1 : n;  // first lane of the next group
        const int s_h = __shfl_sync(0xffffffffu, ss, lane < nr ? hl : 0);
        const int e_l = __shfl_sync(0xffffffffu, me, lane < nr ? max(next - 1, 0) : 0);
        if (lane < nr) {
            gs = s_h;
            ge = e_l;
        }
    }
    (void)grp;
    int len = lane < nr ? ge - gs : 0;
    int incl = len;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 15);
    if (lane < kTileMaxRanges) {
        d->start[lane] = lane < nr ? gs : 0;
        d->prefix[lane] = lane < nr ? incl - len : total;
    }
    if (lane == 0) {
        d->nr = nr;
        d->total = total;
        d->prefix[kTileMaxRanges] = total;
    }
}

template <int DT>
__global__ void __launch_bounds__(kCtThreads, SPNB_CT_MINB)
k_collide_tiles(const float4* __restrict__ pos4, const float* __restrict__ locs, const float* __restrict__ low,
                const float* __restrict__ grid_dims, const uint32_t* __restrict__ keys,
                const float* __restrict__ starts, const float* __restrict__ ends, float* __restrict__ coll, int N,
                int K, int ncells, float edge, float r2, int include_self, int* sym_flag, int* tile_flag,
                TileDesc* __restrict__ descs, unsigned char* __restrict__ blobs, size_t blob_stride, int ntb)
{
    constexpr int D = DT;
    extern __shared__ __align__(128) unsigned char s_raw[];
    // A candidate / hit is a 16-bit code: tile slot (12 bits) | range of the block it lies in (4 bits); the range
    // gives the sorted particle index back with one addition (s_shift), the slot is what the tile rows store.
    float4* s_pos = reinterpret_cast<float4*>(s_raw);                                  // [kTileCap]
    unsigned short* s_cand = reinterpret_cast<unsigned short*>(s_raw + kTileCap * 16); // [8][kCtWin]
    unsigned short* s_hits = s_cand + 8 * kCtWin;                                      // [kTileQ][K]
    __shared__ int s_shift[kTileMaxRanges];  // sorted index = slot + s_shift[range]
    __shared__ TileDesc s_desc;
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ int s_cnt[kTileQ];
    __shared__ unsigned char s_perm[kTileQ];
    __shared__ int s_flags;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tb = blockIdx.x, b = blockIdx.y;
    const int q0 = tb * kTileQ, nq = min(kTileQ, N - q0);
    const size_t sb = (size_t)b * N;
    const float* gd = grid_dims + b * D;
    const float* lo = low + b * D;
    const float* st = starts + (size_t)b * ncells;
    const float* en = ends + (size_t)b * ncells;
    TileDesc* gdesc = descs + (size_t)b * ntb + tb;
    unsigned char* blob = blobs + ((size_t)b * ntb + tb) * blob_stride;

    if (warp == 0) tile_desc_warp(keys + sb, gd, st, en, N, D, ncells, q0, nq, &s_desc);
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        s_flags = 0;
    }
    __syncthreads();
    const int total = s_desc.total;
    if (total + 1 > kTileMaxSlots) {
        // ---- more candidates than 16-bit entries can address (a dense clump): rows from global memory with the
        // general routine, no tile rows for this block, tile lists unusable for this call
        typedef CollideSmem<DT, 256> Fallback;
        Fallback& fsm = *reinterpret_cast<Fallback*>(s_raw);
        const int w0 = warp * 8;
        bool asym = false;
        if (w0 < nq)
            asym = collide_rows_warp<DT, DT, 256>(fsm, locs + (sb + q0 + w0) * D, locs + sb * D,
                                                  coll + (sb + q0 + w0) * K, min(8, nq - w0), lo, gd, st, en, D, K,
                                                  ncells, edge, r2, include_self);
        if (asym && lane == 0) {
            if (sym_flag) atomicOr(sym_flag, 1);
            atomicOr(tile_flag, 1);
        }
        if (tid < kTileOctiles + 1) s_desc.goff[tid] = 0;
        if (tid == 0) {
            s_desc.maxcnt = 0;
            s_desc.sumcnt = 0;
        }
        __syncthreads();
        if (tid < 32) reinterpret_cast<int*>(gdesc)[tid] = reinterpret_cast<const int*>(&s_desc)[tid];
        if (tid == 0) atomicOr(tile_flag, 2);  // no tile rows were written for this block
        return;
    }
    // A block whose ranges exceed the staged tile (the 64 queries span many sparse cells; a few per 10^4 blocks
    // of a uniform cloud) reads its candidates' positions from global memory instead; everything else is the same.
    const bool staged = total + 1 <= kTileCap;

    // ---- stage the positions of the block's ranges (TMA), build slot -> sorted index
    if (warp == 0 && staged) {
        if (lane == 0) mbar_expect_tx(&s_bar, (unsigned)total * 16u);
        __syncwarp();
        if (lane < s_desc.nr) {
            const int len = s_desc.prefix[lane + 1] - s_desc.prefix[lane];
            if (len > 0)
                bulk_copy_g2s(s_pos + 1 + s_desc.prefix[lane], pos4 + sb + s_desc.start[lane], (unsigned)len * 16u, &s_bar);
        }
    }
    if (tid == 0) s_pos[0] = make_float4(1e18f, 1e18f, 1e18f, 0.0f);
    if (tid < kTileMaxRanges) s_shift[tid] = s_desc.start[tid] - s_desc.prefix[tid] - 1;
    auto code_index = [&](unsigned code) { return (int)(code & 0xfffu) + s_shift[code >> 12]; };
    // ---- my warp's 8 queries: lanes 0..7 hold one each
    const int w0 = warp * 8;
    const int nqw = max(0, min(8, nq - w0));
    float myx[D];
    int mygc[D];
    {
        const float4 p = pos4[sb + q0 + min(w0 + (lane & 7), nq - 1)];
        const float t[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
        for (int k = 0; k < D; ++k) {
            myx[k] = t[k];
            mygc[k] = grid_coord_of(t[k], lo[k], edge);
        }
    }
    int myfound = 0;
    bool asym = false, bad = false;
    int total_cells = 1;
#pragma unroll
    for (int k = 0; k < D; ++k) total_cells *= 3;
    unsigned short* cand = s_cand + warp * kCtWin;
    if (staged) mbar_wait(&s_bar, 0);

    int qi = 0;
    while (qi < nqw) {
        int gc[D];
#pragma unroll
        for (int k = 0; k < D; ++k) gc[k] = __shfl_sync(0xffffffffu, mygc[k], qi);
        bool same = lane >= qi && lane < nqw;
#pragma unroll
        for (int k = 0; k < D; ++k) same = same && mygc[k] == gc[k];
        const unsigned smk = __ballot_sync(0xffffffffu, same) >> qi;
        const int run = __ffs(~smk) - 1;
#pragma unroll
        for (int k = 0; k < D; ++k)
            if (gd[k] > 0.0f && (float)gc[k] >= gd[k] + 1.0f) asym = true;

        // ---- lane -> one neighbour cell: its particles as a run of tile slots
        int cnt = 0, slot0 = 0;
        if (lane < total_cells) {
            int rem = lane, id = 0;
            bool ok = true;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const int c = gc[k] + (rem % 3) - 1;
                rem /= 3;
                if (c < 0 || (float)c >= gd[k]) ok = false;
                else id += hash_term(c, gd, k, D);
            }
            if (ok && id >= 0 && id < ncells) {
                const int cstart = (int)st[id];
                cnt = (int)en[id] - cstart;
                if (cnt < 0) cnt = 0;
                if (cnt > 0) {
                    // the range of the block that holds this cell (ascending starts; slot = index + shift)
                    int g = 0;
                    for (int t = 1; t < s_desc.nr; ++t) g += cstart >= s_desc.start[t];
                    const int s0 = s_desc.start[g], p0 = s_desc.prefix[g];
                    if (cstart < s0 || cstart + cnt - s0 > s_desc.prefix[g + 1] - p0) {
                        bad = true;  // never expected: the cell is not inside the block's ranges
                        cnt = 0;
                    }
                    slot0 = (cstart - s0 + p0 + 1) | (g << 12);
                }
            }
        }
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        const int ctotal = __shfl_sync(0xffffffffu, incl, 31);
        const int excl = incl - cnt;

        for (int wb = 0; wb < ctotal; wb += kCtWin) {
            const int wn = min(kCtWin, ctotal - wb);
            // every cell lane writes the slots of its particles that fall into this window
            __syncwarp();
            {
                const int a = max(excl, wb), z = min(incl, wb + kCtWin);
                for (int t = a; t < z; ++t) cand[t - wb] = (unsigned short)(slot0 + (t - excl));
            }
            __syncwarp();
            // the window in registers: slot and coordinates of 8 candidates per lane
            unsigned cs[kCtWin / 32];
            float px[kCtWin / 32][D];
#pragma unroll
            for (int i = 0; i < kCtWin / 32; ++i) {
                const int t = i * 32 + lane;
                cs[i] = t < wn ? cand[t] : 0u;  // code 0: the sentinel, far outside every radius
                float4 p;
                if (staged) p = s_pos[cs[i] & 0xfffu];
                else p = cs[i] ? pos4[sb + code_index(cs[i])] : make_float4(1e18f, 1e18f, 1e18f, 0.0f);
                const float tt[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
                for (int k = 0; k < D; ++k) px[i][k] = tt[k];
            }
            for (int r = 0; r < run; ++r) {
                int found = __shfl_sync(0xffffffffu, myfound, qi + r);
                if (found >= K) continue;
                float x[D];
#pragma unroll
                for (int k = 0; k < D; ++k) x[k] = __shfl_sync(0xffffffffu, myx[k], qi + r);
                unsigned short* hrow = s_hits + (size_t)(w0 + qi + r) * K;
                // all distance tests first (no dependence between chunks), then the ordered compaction
                unsigned m[kCtWin / 32], mine = 0;
#pragma unroll
                for (int i = 0; i < kCtWin / 32; ++i) {
                    m[i] = 0;
                    if (i * 32 < wn) {
                        float d = 0.0f;
#pragma unroll
                        for (int k = 0; k < D; ++k) {
                            const float nr = x[k] - px[i][k];
                            d += nr * nr;
                        }
                        const bool hit = d < r2 && (d > 0.0f || include_self);
                        m[i] = __ballot_sync(0xffffffffu, hit);
                        mine |= (hit ? 1u : 0u) << i;
                    }
                }
#pragma unroll
                for (int i = 0; i < kCtWin / 32; ++i) {
                    const int pos = found + __popc(m[i] & lanemask_lt());
                    if (((mine >> i) & 1u) && pos < K) hrow[pos] = (unsigned short)cs[i];
                    found += __popc(m[i]);
                }
                if (lane == qi + r) myfound = found;
            }
        }
        qi += run;
    }
    if (lane < 8) {
        int c = lane < nqw ? myfound : 0;
        if (c >= K) {
            asym = true;  // the row is full: it may have been cut
            c = K;
        }
        s_cnt[w0 + lane] = c;
    }
    if (__any_sync(0xffffffffu, asym) && lane == 0) atomicOr(&s_flags, 1);
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(&s_flags, 4);
    __syncthreads();

    // ---- rank the queries by list length (longest first, ties by position)
    if (tid < kTileQ) {
        const int c = s_cnt[tid];
        int rank = 0;
        for (int j = 0; j < kTileQ; ++j) {
            const int cj = s_cnt[j];
            rank += (cj > c || (cj == c && j < tid)) ? 1 : 0;
        }
        s_perm[rank] = (unsigned char)tid;
    }
    __syncthreads();
    if (tid == 0) {
        int rows = 0, sum = 0;
        for (int g = 0; g < kTileOctiles; ++g) {
            s_desc.goff[g] = (unsigned short)rows;
            rows += (s_cnt[s_perm[8 * g]] + 3) >> 2;  // the first rank of an octile has its longest list
        }
        s_desc.goff[kTileOctiles] = (unsigned short)rows;
        for (int j = 0; j < kTileQ; ++j) sum += s_cnt[j];
        s_desc.maxcnt = (unsigned short)s_cnt[s_perm[0]];
        s_desc.sumcnt = sum;
#pragma unroll
        for (int i = 0; i < 5; ++i) s_desc.pad[i] = 0;
        if (s_flags & 1) {
            if (sym_flag) atomicOr(sym_flag, 1);
            atomicOr(tile_flag, 1);
        }
        if (s_flags & 4) atomicOr(tile_flag, 4);
    }
    __syncthreads();

    // ---- float rows: accepted particle indices in the reference's order, then -1 up to the end of the row
    const bool vec_rows = (K & 3) == 0 && (reinterpret_cast<size_t>(coll) & 15) == 0;
    for (int r = 0; r < 8; ++r) {
        const int ql = w0 + r;
        if (ql >= nq) break;
        const int c = s_cnt[ql];
        float* row = coll + (sb + q0 + ql) * K;
        const unsigned short* hrow = s_hits + (size_t)ql * K;
        if (vec_rows) {
            for (int k4 = lane; k4 * 4 < K; k4 += 32) {
                const uint2 h = *reinterpret_cast<const uint2*>(hrow + 4 * k4);
                const int k = 4 * k4;
                float4 v;
                v.x = k + 0 < c ? (float)code_index(h.x & 0xffffu) : -1.0f;
                v.y = k + 1 < c ? (float)code_index(h.x >> 16) : -1.0f;
                v.z = k + 2 < c ? (float)code_index(h.y & 0xffffu) : -1.0f;
                v.w = k + 3 < c ? (float)code_index(h.y >> 16) : -1.0f;
                *reinterpret_cast<float4*>(row + k) = v;
            }
        } else {
            for (int k = lane; k < K; k += 32) row[k] = k < c ? (float)code_index(hrow[k]) : -1.0f;
        }
    }
    // ---- tile rows of octile `warp`: entry 4s+e of rank 8g+r at position 4r+e of row goff[g]+s
    {
        const int g = warp;
        const int ql = s_perm[8 * g + (lane >> 2)];
        const int c = s_cnt[ql];
        const unsigned short* hrow = s_hits + (size_t)ql * K;
        const int S = s_desc.goff[g + 1] - s_desc.goff[g];
        unsigned short* rows = reinterpret_cast<unsigned short*>(blob + kTileHeaderBytes) + (size_t)s_desc.goff[g] * 32;
        for (int s_ = 0; s_ < S; ++s_) {
            const int e = 4 * s_ + (lane & 3);
            rows[s_ * 32 + lane] = e < c ? (unsigned short)((hrow[e] & 0xfffu) << 4) : (unsigned short)0;
        }
    }
    if (tid < kTileQ) blob[tid] = s_perm[tid];
    if (tid < 32) reinterpret_cast<int*>(gdesc)[tid] = reinterpret_cast<const int*>(&s_desc)[tid];
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

static int reorder_launch(const float* locs, const float* data, const float* idxs, float* nlocs, float* ndata,
                          float* pos4, int B, int N, int D, int C, int reverse, cudaStream_t stream)
{
    if (B <= 0 || N <= 0 || D <= 0) {
        set_error("spnb_reorder_data: bad sizes");
        return 0;
    }
    if (!locs || !idxs || !nlocs || (data && !ndata)) {
        set_error("spnb_reorder_data: null pointer");
        return 0;
    }
    if (!data) C = 0;
    if (pos4 && (reverse || D > 4)) {
        set_error("spnb_reorder_data: the float4 position plane needs reverse == 0 and ndims <= 4");
        return 0;
    }
    if ((long long)N * (D > C ? D : C) >= (1ll << 32)) {
        set_error("spnb_reorder_data: N * row width exceeds 2^32");
        return 0;
    }
    int launches = 0;
    // rows of up to 4 floats: thread per row; both tensors in one launch when their widths agree
    auto rows = [&](int W, const float* a, const float* b_, float* oa, float* ob, float4* p4) {
        const dim3 grid(cdiv(N, 256), B, b_ ? 2 : 1);
        switch (W) {
        case 1: k_reorder_rows<1><<<grid, 256, 0, stream>>>(a, b_, idxs, oa, ob, p4, N, reverse); break;
        case 2: k_reorder_rows<2><<<grid, 256, 0, stream>>>(a, b_, idxs, oa, ob, p4, N, reverse); break;
        case 3: k_reorder_rows<3><<<grid, 256, 0, stream>>>(a, b_, idxs, oa, ob, p4, N, reverse); break;
        case 4: k_reorder_rows<4><<<grid, 256, 0, stream>>>(a, b_, idxs, oa, ob, p4, N, reverse); break;
        case 5: k_reorder_rows<5><<<grid, 256, 0, stream>>>(a, b_, idxs, oa, ob, p4, N, reverse); break;
        case 6: k_reorder_rows<6><<<grid, 256, 0, stream>>>(a, b_, idxs, oa, ob, p4, N, reverse); break;
        case 7: k_reorder_rows<7><<<grid, 256, 0, stream>>>(a, b_, idxs, oa, ob, p4, N, reverse); break;
        default: k_reorder_rows<8><<<grid, 256, 0, stream>>>(a, b_, idxs, oa, ob, p4, N, reverse); break;
        }
        ++launches;
    };
    auto wide = [&](int W, const float* a, float* oa) {
        const bool vec = (W & 3) == 0 && ((reinterpret_cast<size_t>(a) | reinterpret_cast<size_t>(oa)) & 15) == 0;
        int blocks = cdiv((long long)N * (vec ? W / 4 : W), 256 * 4);
        if (blocks < 1) blocks = 1;
        const dim3 grid(blocks, B);
        if (vec) k_reorder_wide<true><<<grid, 256, 0, stream>>>(a, idxs, oa, N, W, reverse);
        else k_reorder_wide<false><<<grid, 256, 0, stream>>>(a, idxs, oa, N, W, reverse);
        ++launches;
    };
    if (C == D && D <= 4) {
        rows(D, locs, data, nlocs, ndata, (float4*)pos4);
    } else {
        if (D <= 8) rows(D, locs, nullptr, nlocs, nullptr, (float4*)pos4);
        else wide(D, locs, nlocs);
        if (C > 0) {
            if (C <= 8) rows(C, data, nullptr, ndata, nullptr, nullptr);
            else wide(C, data, ndata);
        }
    }
    count_launches(launches);
    return check_launch("spnb_reorder_data") ? 1 : 0;
}

int spnb_reorder_data(const float* locs, const float* data, const float* idxs, float* nlocs,
                      float* ndata, int B, int N, int D, int C, int reverse, void* stream_)
{
    return reorder_launch(locs, data, idxs, nlocs, ndata, nullptr, B, N, D, C, reverse, (cudaStream_t)stream_);
}

int spnb_reorder_data_pos4(const float* locs, const float* data, const float* idxs, float* nlocs,
                           float* ndata, float* pos4, int B, int N, int D, int C, void* stream_)
{
    return reorder_launch(locs, data, idxs, nlocs, ndata, pos4, B, N, D, C, 0, (cudaStream_t)stream_);
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

int spnb_compute_collisions_tiled(const float* pos4, const float* locs, const float* low, const float* grid_dims,
                                  const float* cellIDs, float* cellStarts, float* cellEnds, float* collisions,
                                  int B, int N, int D, int K, int ncells, float cellEdge, float radius,
                                  int include_self, int* sym_flag, void* tile_lists, size_t tile_lists_bytes,
                                  void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!valid_common(B, N, D, "spnb_compute_collisions_tiled")) return 0;
    if (K <= 0 || ncells <= 0) {
        set_error("spnb_compute_collisions_tiled: bad sizes max_collisions=%d ncells=%d", K, ncells);
        return 0;
    }
    if (!pos4 || !locs || !low || !grid_dims || !cellIDs || !cellStarts || !cellEnds || !collisions || !tile_lists) {
        set_error("spnb_compute_collisions_tiled: null pointer");
        return 0;
    }
    if (!tile_lists_supported(N, D, K)) {
        set_error("spnb_compute_collisions_tiled: needs ndims <= %d and max_collisions <= %d", kTileMaxNdim, kTileMaxK);
        return 0;
    }
    if ((reinterpret_cast<size_t>(pos4) & 15) != 0) {
        set_error("spnb_compute_collisions_tiled: the position plane must be 16-byte aligned");
        return 0;
    }
    const TileLayout tl = tile_layout(B, N, K);
    if (tile_lists_bytes < tl.total) {
        set_error("spnb_compute_collisions_tiled: tile buffer too small (%zu < %zu)", tile_lists_bytes, tl.total);
        return 0;
    }
    char* tb = (char*)tile_lists;
    int* tflag = (int*)tb;
    TileDesc* descs = (TileDesc*)(tb + tl.desc_off);
    cudaMemsetAsync(tflag, 0, 128, stream);
    k_table_clear<<<dim3(148, B), 256, 0, stream>>>(grid_dims, cellStarts, cellEnds, D, ncells);
    k_table_fill<<<dim3(cdiv(N, 256), B), 256, 0, stream>>>((const uint32_t*)cellIDs, grid_dims, cellStarts,
                                                            cellEnds, N, D, ncells);
    const float r2 = radius * radius;
    size_t smem = (size_t)kTileCap * 16 + 8 * kCtWin * 2 + (size_t)kTileQ * K * 2;
    const dim3 grid(tl.ntb, B);
#define SPNB_CT(DT)                                                                                              \
    do {                                                                                                         \
        if (smem < sizeof(CollideSmem<DT, 256>)) smem = sizeof(CollideSmem<DT, 256>);                            \
        if (smem > 48 * 1024 &&                                                                                  \
            cudaFuncSetAttribute(k_collide_tiles<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != \
                cudaSuccess) {                                                                                   \
            set_error("spnb_compute_collisions_tiled: %zu bytes of shared memory not available", smem);         \
            return 0;                                                                                            \
        }                                                                                                        \
        k_collide_tiles<DT><<<grid, kCtThreads, smem, stream>>>(                                                 \
            (const float4*)pos4, locs, low, grid_dims, (const uint32_t*)cellIDs, cellStarts, cellEnds,           \
            collisions, N, K, ncells, cellEdge, r2, include_self, sym_flag, tflag, descs,                        \
            (unsigned char*)(tb + tl.list_off), tl.blob_stride, tl.ntb);                                         \
    } while (0)
    switch (D) {
    case 1: SPNB_CT(1); break;
    case 2: SPNB_CT(2); break;
    default: SPNB_CT(3); break;
    }
#undef SPNB_CT
    count_launches(3);
    return check_launch("spnb_compute_collisions_tiled") ? 1 : 0;
}

}  // extern "C"
