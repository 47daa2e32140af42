// N5-N9: NavMap::generateVoronoiMap (nav_map.cpp:541-590) on the down-sampled grid.
//
//   N5 markObs/floodFillObs (:384-432)  4-connected components of occupied cells, label = rank of the component's
//                                       first cell in raster order           -> union-find (min-index root) + rank
//   N6 calObsDist (:434-473)            L1 distance to the nearest in-map occupied cell, cap 999 (else 1000), and
//                                       the region of a nearest obstacle      -> separable (dist,region) min-plus
//   N7 findEdge (:475-507)              cells where the nearest-obstacle region changes
//   N8 calEdgeDist (:509-539)           L1 distance to the nearest edge cell through ALL cells, cap 999
//   N9 cost loop (:565-582)             float Voronoi-field cost
//
// Order-dependent parts of the reference (broken heap) are replaced by canonical, order-free
// rules: region_ of a free cell = the SMALLEST region id among all nearest obstacle cells (the reference always
// holds *a* nearest region); obstacle-side edge flags follow the "first differing neighbour" rule without the
// processing-order condition. obs_dist_ is order-independent and bit-exact.
#include "dist.cuh"

namespace mpros
{

constexpr unsigned long long kOne = 1ull << 32;          // +1 distance in a packed (dist<<32 | region) key
constexpr unsigned long long kInfKey = (1ull << 24) << 32; // "no source seen yet"
constexpr int kScanWords = (kHugeInt + 31) / 32 + 1;      // sources further than 999 cells can never matter

// ---- N5 ------------------------------------------------------------------------------------------------
// start column of the horizontal run of set bits that contains (i, j); bit (i, j) must be set
__device__ __forceinline__ int run_start(const uint32_t *__restrict__ row, int j)
{
    int w = j >> 5, b = j & 31;
    uint32_t z = ~row[w] & ((b == 31) ? 0x7fffffffu : ((1u << b) - 1u)); // zero bits strictly below b
    z &= (1u << b) - 1u;
    if (z)
        return (w << 5) + (32 - __clz(z));
    for (--w; w >= 0; --w)
    {
        uint32_t zz = ~row[w];
        if (zz)
            return (w << 5) + (32 - __clz(zz));
    }
    return 0;
}

__device__ __forceinline__ int uf_find(const int *parent, int a)
{
    int p = parent[a];
    while (p != a)
    {
        a = p;
        p = parent[a];
    }
    return a;
}

__device__ __forceinline__ void uf_union(int *parent, int a, int b)
{
    while (true)
    {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b)
            return;
        if (a < b)
        {
            int t = a;
            a = b;
            b = t;
        }
        int old = atomicMin(&parent[a], b); // hook the larger root under the smaller one
        if (old == a)
            return;
        a = old;
    }
}

// parent[k] = linear index of the start of k's horizontal run (occupied cells), -1 for free cells
__global__ void __launch_bounds__(256) ccl_init_kernel(const uint32_t *__restrict__ occ, int H, int W, int pitch,
                                                       int *__restrict__ parent)
{
    size_t total = (size_t)H * W;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x)
    {
        int i = (int)(k / W), j = (int)(k % W);
        const uint32_t *row = occ + (size_t)i * pitch;
        bool o = (row[j >> 5] >> (j & 31)) & 1u;
        parent[k] = o ? i * W + run_start(row, j) : -1;
    }
}

// merge vertically adjacent runs; one union per contact segment (its left-most column)
__global__ void __launch_bounds__(256) ccl_merge_kernel(const uint32_t *__restrict__ occ, int H, int W, int pitch,
                                                        int *__restrict__ parent)
{
    size_t total = (size_t)H * W;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x)
    {
        int i = (int)(k / W), j = (int)(k % W);
        if (i == 0)
            continue;
        const uint32_t *row = occ + (size_t)i * pitch, *up = row - pitch;
        bool o = (row[j >> 5] >> (j & 31)) & 1u, u = (up[j >> 5] >> (j & 31)) & 1u;
        if (!(o && u))
            continue;
        if (j > 0)
        {
            bool ol = (row[(j - 1) >> 5] >> ((j - 1) & 31)) & 1u, ul = (up[(j - 1) >> 5] >> ((j - 1) & 31)) & 1u;
            if (ol && ul)
                continue; // same contact segment as the cell to the left
        }
        uf_union(parent, (int)k, (int)k - W);
    }
}

// flatten + count roots per 1024-cell tile; local rank of every root inside its tile
// groot (banded form only, else nullptr): for a local root that touches the band's first or last row, the GLOBAL root of
// its component after the seam union (global cell index); a component whose global root lies in another band is not
// counted here.
__global__ void __launch_bounds__(1024) ccl_flatten_rank_kernel(int *__restrict__ parent, size_t total, int *__restrict__ local_rank,
                                                                int *__restrict__ tile_count, const int *__restrict__ groot, int base)
{
    __shared__ int warp_sum[32];
    size_t k = (size_t)blockIdx.x * 1024 + threadIdx.x;
    bool root = false;
    if (k < total)
    {
        int p = parent[k];
        if (p >= 0)
        {
            p = uf_find(parent, (int)k);
            root = ((size_t)p == k);
            if (root && groot)
            {
                const int gr = groot[k];
                root = gr < 0 || gr == (int)k + base;
            }
        }
        // flattening is finished by ccl_region_kernel (roots are stable here, chains still may exist)
    }
    unsigned bal = __ballot_sync(0xffffffffu, root);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0)
        warp_sum[wid] = __popc(bal);
    __syncthreads();
    if (wid == 0)
    {
        int v = warp_sum[lane];
        int incl = v;
        for (int o = 1; o < 32; o <<= 1)
        {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o)
                incl += t;
        }
        warp_sum[lane] = incl - v; // exclusive
        if (lane == 31)
            tile_count[blockIdx.x] = incl;
    }
    __syncthreads();
    if (root)
        local_rank[k] = warp_sum[wid] + __popc(bal & ((1u << lane) - 1u));
}

// exclusive scan of tile counts by one block; total written to *n_regions
__global__ void __launch_bounds__(1024) scan_tiles_kernel(int *__restrict__ tile_count, int n_tiles, int *__restrict__ n_regions)
{
    __shared__ int part[1024];
    int t = threadIdx.x;
    int chunk = (n_tiles + 1023) / 1024;
    int b = t * chunk, e = min(n_tiles, b + chunk);
    int s = 0;
    for (int k = b; k < e; ++k)
        s += tile_count[k];
    part[t] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1)
    {
        int v = (t >= o) ? part[t - o] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    int run = part[t] - s;
    for (int k = b; k < e; ++k)
    {
        int v = tile_count[k];
        tile_count[k] = run;
        run += v;
    }
    if (t == 1023)
        *n_regions = part[1023];
}

// root_region (banded form only): final region id of the local roots that touch a band boundary (-1 elsewhere);
// id_offset = number of components whose first cell lies in an earlier band.
__global__ void __launch_bounds__(256) ccl_region_kernel(const int *__restrict__ parent, size_t total, const int *__restrict__ local_rank,
                                                         const int *__restrict__ tile_offset, int *__restrict__ region,
                                                         const int *__restrict__ root_region, const int *__restrict__ id_offset)
{
    const int add = id_offset ? *id_offset : 0;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x)
    {
        int p = parent[k];
        int r = -1;
        if (p >= 0)
        {
            int root = uf_find(parent, (int)k);
            const int rr = root_region ? root_region[root] : -1;
            r = rr >= 0 ? rr : add + tile_offset[root >> 10] + local_rank[root];
        }
        region[k] = r;
    }
}

// ---- N6 / N8: separable L1 distance transform carrying the source label -------------------------------------
// Row pass: for every cell the nearest source in its own row (left or right), as key = dist<<32 | label; ties take
// the smaller label. label == nullptr -> label 0 (plain distance).
__global__ void __launch_bounds__(256) dt_row_kernel(const uint32_t *__restrict__ src, int H, int W, int pitch,
                                                     const int *__restrict__ label, unsigned long long *__restrict__ key)
{
    size_t total = (size_t)H * W;
    const int wpr = (W + 31) >> 5;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x)
    {
        int i = (int)(k / W), j = (int)(k % W);
        const uint32_t *row = src + (size_t)i * pitch;
        int w = j >> 5, b = j & 31;
        uint32_t cur = row[w];
        unsigned long long best = kInfKey;
        if ((cur >> b) & 1u)
            best = label ? (unsigned long long)(unsigned)label[k] : 0ull;
        else
        {
            // nearest set bit to the left
            int jl = -1;
            uint32_t m = cur & ((1u << b) - 1u);
            if (m)
                jl = (w << 5) + 31 - __clz(m);
            else
                for (int ww = w - 1; ww >= 0 && ww >= w - kScanWords; --ww)
                {
                    uint32_t v = row[ww];
                    if (v)
                    {
                        jl = (ww << 5) + 31 - __clz(v);
                        break;
                    }
                }
            // nearest set bit to the right
            int jr = -1;
            m = (b == 31) ? 0u : (cur & ~((2u << b) - 1u));
            if (m)
                jr = (w << 5) + __ffs(m) - 1;
            else
                for (int ww = w + 1; ww < wpr && ww <= w + kScanWords; ++ww)
                {
                    uint32_t v = row[ww];
                    if (v)
                    {
                        jr = (ww << 5) + __ffs(v) - 1;
                        break;
                    }
                }
            if (jl >= 0)
            {
                unsigned long long lab = label ? (unsigned long long)(unsigned)label[(size_t)i * W + jl] : 0ull;
                best = ((unsigned long long)(j - jl) << 32) | lab;
            }
            if (jr >= 0)
            {
                unsigned long long lab = label ? (unsigned long long)(unsigned)label[(size_t)i * W + jr] : 0ull;
                unsigned long long kr = ((unsigned long long)(jr - j) << 32) | lab;
                best = kr < best ? kr : best;
            }
        }
        key[k] = best;
    }
}

// Column pass: key[i] = min over i' of key[i'] + |i - i'| (in the packed key, distance in the high word) = the classic
// forward / backward sweep pair. A sweep is a prefix minimum in disguise:
//   fwd[i] = min_{i' <= i} key[i'] + (i - i')  =  prefixmin_i( key[i'] + (H - i') ) - (H - i),
// so a column is cut into 32 chunks that are scanned in parallel: per chunk aggregate -> exclusive prefix over the 32
// aggregates in shared memory -> sweep of the chunk seeded with it. Block = 32 columns x 32 chunks; a warp reads 32
// consecutive keys of one row (256 B). The old one-thread-per-column sweep (1024 dependent steps, 8 CTAs) took 613 us
// per call at 1024^2 and was 72 % of the whole map preprocessing.
// Banded form (rows [0, H) of `key` are rows [row0, row0 + H) of a grid tiled across `world` ranks): band_agg holds, per
// rank g and column, the two aggregates of that band's ROW-pass keys, [g][0][j] = min key + (Hg - i) and [g][1][j] = min key + i
// (i global). The forward sweep is seeded with the minimum of the earlier bands' first aggregate, the backward sweep with
// the later bands' second one: a source in a later band reaches row i at key + (i' - i) whether or not its own forward
// value was smaller, so the aggregates of the row-pass keys are enough and ONE exchange serves both sweeps.
__global__ void __launch_bounds__(1024) dt_col_kernel(unsigned long long *__restrict__ key, int H, int W, int row0, int Hg,
                                                      const unsigned long long *__restrict__ band_agg, int rank, int world)
{
    __shared__ unsigned long long agg[32][33];
    const int x = threadIdx.x, y = threadIdx.y;
    const int j = blockIdx.x * 32 + x;
    const int R = (H + 31) / 32;
    const int i0 = min(H, y * R), i1 = min(H, i0 + R);
    const bool live = j < W;
    unsigned long long seed_f = ~0ull, seed_b = ~0ull;
    if (band_agg && live)
    {
        for (int g = 0; g < rank; ++g)
        {
            const unsigned long long v = band_agg[((size_t)g * 2 + 0) * W + j];
            seed_f = v < seed_f ? v : seed_f;
        }
        for (int g = rank + 1; g < world; ++g)
        {
            const unsigned long long v = band_agg[((size_t)g * 2 + 1) * W + j];
            seed_b = v < seed_b ? v : seed_b;
        }
    }
    // forward: aggregate of key[i] + (H - i)
    unsigned long long a = ~0ull;
    if (live)
        for (int i = i0; i < i1; ++i)
        {
            unsigned long long v = key[(size_t)i * W + j] + (unsigned long long)(Hg - (row0 + i)) * kOne;
            a = v < a ? v : a;
        }
    agg[y][x] = a;
    __syncthreads();
    unsigned long long run = seed_f;
    for (int c = 0; c < y; ++c)
    {
        unsigned long long v = agg[c][x];
        run = v < run ? v : run;
    }
    __syncthreads();
    // forward sweep of the chunk (in place) + aggregate of fwd[i] + i for the backward pass
    a = ~0ull;
    if (live)
        for (int i = i0; i < i1; ++i)
        {
            const unsigned long long off = (unsigned long long)(Hg - (row0 + i)) * kOne;
            unsigned long long v = key[(size_t)i * W + j] + off;
            run = v < run ? v : run;
            const unsigned long long f = run - off;
            key[(size_t)i * W + j] = f;
            const unsigned long long b = f + (unsigned long long)(row0 + i) * kOne;
            a = b < a ? b : a;
        }
    agg[y][x] = a;
    __syncthreads();
    run = seed_b;
    for (int c = y + 1; c < 32; ++c)
    {
        unsigned long long v = agg[c][x];
        run = v < run ? v : run;
    }
    // backward sweep: out[i] = min_{i' >= i} fwd[i'] + (i' - i)
    if (live)
        for (int i = i1 - 1; i >= i0; --i)
        {
            const unsigned long long off = (unsigned long long)(row0 + i) * kOne;
            unsigned long long v = key[(size_t)i * W + j] + off;
            run = v < run ? v : run;
            key[(size_t)i * W + j] = run - off;
        }
}

// The two aggregates of one band's row-pass keys per column (see dt_col_kernel): out[0][j], out[1][j]. ~0 for an empty band.
__global__ void __launch_bounds__(1024) dt_col_agg_kernel(const unsigned long long *__restrict__ key, int H, int W, int row0, int Hg,
                                                          unsigned long long *__restrict__ out)
{
    __shared__ unsigned long long sf[32][33], sb[32][33];
    const int x = threadIdx.x, y = threadIdx.y;
    const int j = blockIdx.x * 32 + x;
    const int R = (H + 31) / 32;
    const int i0 = min(H, y * R), i1 = min(H, i0 + R);
    unsigned long long f = ~0ull, b = ~0ull;
    if (j < W)
        for (int i = i0; i < i1; ++i)
        {
            const unsigned long long k = key[(size_t)i * W + j];
            const unsigned long long vf = k + (unsigned long long)(Hg - (row0 + i)) * kOne, vb = k + (unsigned long long)(row0 + i) * kOne;
            f = vf < f ? vf : f;
            b = vb < b ? vb : b;
        }
    sf[y][x] = f;
    sb[y][x] = b;
    __syncthreads();
    if (y == 0 && j < W)
    {
        for (int c = 1; c < 32; ++c)
        {
            f = sf[c][x] < f ? sf[c][x] : f;
            b = sb[c][x] < b ? sb[c][x] : b;
        }
        out[j] = f;
        out[(size_t)W + j] = b;
    }
}

// N6 epilogue: cap (>= 1000 stays 1000 / region -1), keep obstacle cells' own labels
__global__ void __launch_bounds__(256) obs_finish_kernel(const unsigned long long *__restrict__ key, size_t total,
                                                         uint16_t *__restrict__ obs_dist, int *__restrict__ region)
{
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x)
    {
        unsigned long long v = key[k];
        unsigned d = (unsigned)(v >> 32);
        if (d >= (unsigned)kHugeInt)
        {
            obs_dist[k] = (uint16_t)kHugeInt;
            region[k] = -1;
        }
        else
        {
            obs_dist[k] = (uint16_t)d;
            region[k] = (int)(unsigned)(v & 0xffffffffull);
        }
    }
}
__global__ void __launch_bounds__(256) edge_finish_kernel(const unsigned long long *__restrict__ key, size_t total,
                                                          uint16_t *__restrict__ edge_dist)
{
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x)
    {
        unsigned d = (unsigned)(key[k] >> 32);
        edge_dist[k] = (uint16_t)(d >= (unsigned)kHugeInt ? kHugeInt : d);
    }
}

// ---- N7 -------------------------------------------------------------------------------------------------
// A reached free cell (obs_dist < 1000, not occupied) is an edge iff some in-map 4-neighbour has a different
// region. Any other cell (obstacle, unreached) is an edge iff it is the FIRST differing neighbour, in the order
// up (i-1), right (j+1), down (i+1), left (j-1) of nav_map.cpp:351-382, of some reached free neighbour.
__device__ __forceinline__ int first_differing(const int *__restrict__ region, int H, int W, int i, int j, int r)
{
    if (i - 1 >= 0 && region[(size_t)(i - 1) * W + j] != r)
        return 0;
    if (j + 1 < W && region[(size_t)i * W + j + 1] != r)
        return 1;
    if (i + 1 < H && region[(size_t)(i + 1) * W + j] != r)
        return 2;
    if (j - 1 >= 0 && region[(size_t)i * W + j - 1] != r)
        return 3;
    return -1;
}

__global__ void __launch_bounds__(256) edge_kernel(const uint32_t *__restrict__ occ, const uint16_t *__restrict__ obs_dist,
                                                   const int *__restrict__ region, int H, int W, int pitch,
                                                   uint32_t *__restrict__ edge_bits, int row0, int nrows)
{
    const int wpr = (W + 31) >> 5;
    const size_t first_word = (size_t)row0 * wpr, total_words = (size_t)(row0 + nrows) * wpr;
    const int lane = threadIdx.x & 31;
    size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t wd = first_word + warp; wd < total_words; wd += nwarps)
    {
        int i = (int)(wd / wpr), w = (int)(wd % wpr);
        int j = (w << 5) + lane;
        bool e = false;
        if (j < W)
        {
            size_t k = (size_t)i * W + j;
            bool o = (occ[(size_t)i * pitch + w] >> lane) & 1u;
            bool reached_free = !o && obs_dist[k] < kHugeInt;
            int r = region[k];
            if (reached_free)
                e = first_differing(region, H, W, i, j, r) >= 0;
            else
            {
                // am I the first differing neighbour of a reached free neighbour? (my direction seen from it)
                const int di[4] = {1, 0, -1, 0}, dj[4] = {0, -1, 0, 1}; // neighbour n such that I am n's up/right/down/left
#pragma unroll
                for (int d = 0; d < 4; ++d)
                {
                    int ni = i + di[d], nj = j + dj[d];
                    if (ni < 0 || ni >= H || nj < 0 || nj >= W)
                        continue;
                    size_t nk = (size_t)ni * W + nj;
                    bool no = (occ[(size_t)ni * pitch + (nj >> 5)] >> (nj & 31)) & 1u;
                    if (no || obs_dist[nk] >= kHugeInt)
                        continue;
                    if (first_differing(region, H, W, ni, nj, region[nk]) == d)
                        e = true;
                }
            }
        }
        uint32_t word = __ballot_sync(0xffffffffu, e);
        if (lane == 0)
            edge_bits[(size_t)i * pitch + w] = word;
    }
}

// ---- N9 -------------------------------------------------------------------------------------------------
// occupied -> 1; edge -> 0; else (a/(a+do)) * (de/(de+do)) * (do-dmax) * (do-dmax) / dmax / dmax, evaluated left
// to right in float exactly as nav_map.cpp:577-580 (library is compiled with --fmad=false).
__global__ void __launch_bounds__(256) voronoi_cost_kernel(const uint32_t *__restrict__ occ, const uint32_t *__restrict__ edge_bits,
                                                           const uint16_t *__restrict__ obs_dist, const uint16_t *__restrict__ edge_dist,
                                                           int H, int W, int pitch, float alpha, float dmax, float *__restrict__ cost)
{
    size_t total = (size_t)H * W;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x)
    {
        int i = (int)(k / W), j = (int)(k % W);
        size_t wi = (size_t)i * pitch + (j >> 5);
        float c;
        if ((occ[wi] >> (j & 31)) & 1u)
            c = 1.0f;
        else if ((edge_bits[wi] >> (j & 31)) & 1u)
            c = 0.0f;
        else
        {
            int dob = obs_dist[k], de = edge_dist[k];
            float fo = (float)dob;
            float a = alpha / (alpha + fo);
            float b = (float)de / (float)(de + dob);
            float t = a * b;
            t = t * (fo - dmax);
            t = t * (fo - dmax);
            t = t / dmax;
            t = t / dmax;
            c = t;
        }
        cost[k] = c;
    }
}

int map_build_voronoi(Ctx *c)
{
    cudaStream_t st = c->stream;
    const int H = c->H, W = c->W, pitch = c->pitch;
    const size_t cells = (size_t)H * W;
    const int n_tiles = (int)((cells + 1023) / 1024);
    // scratch: key u64[cells] | parent i32[cells] | local_rank i32[cells] | tile_count i32[n_tiles] | n_regions
    size_t need = cells * 8 + cells * 4 * 2 + (size_t)n_tiles * 4 + 64;
    int rc = ensure_scratch(c, need);
    if (rc)
        return rc;
    unsigned long long *key = static_cast<unsigned long long *>(c->scratch);
    int *parent = reinterpret_cast<int *>(key + cells);
    int *local_rank = parent + cells;
    int *tile_count = local_rank + cells;
    int *d_nreg = tile_count + n_tiles;

    // N5
    ccl_init_kernel<<<grid_for(c, cells), 256, 0, st>>>(c->ds_bits, H, W, pitch, parent);
    MPROS_LAUNCH_CHECK(c);
    ccl_merge_kernel<<<grid_for(c, cells), 256, 0, st>>>(c->ds_bits, H, W, pitch, parent);
    MPROS_LAUNCH_CHECK(c);
    ccl_flatten_rank_kernel<<<n_tiles, 1024, 0, st>>>(parent, cells, local_rank, tile_count, nullptr, 0);
    MPROS_LAUNCH_CHECK(c);
    scan_tiles_kernel<<<1, 1024, 0, st>>>(tile_count, n_tiles, d_nreg);
    MPROS_LAUNCH_CHECK(c);
    ccl_region_kernel<<<grid_for(c, cells), 256, 0, st>>>(parent, cells, local_rank, tile_count, c->region, nullptr, nullptr);
    MPROS_LAUNCH_CHECK(c);
    MPROS_CUDA(cudaMemcpyAsync(&c->num_regions, d_nreg, sizeof(int), cudaMemcpyDeviceToHost, st));

    // N6
    dt_row_kernel<<<grid_for(c, cells), 256, 0, st>>>(c->ds_bits, H, W, pitch, c->region, key);
    MPROS_LAUNCH_CHECK(c);
    dt_col_kernel<<<cdiv(W, 32), dim3(32, 32), 0, st>>>(key, H, W, 0, H, nullptr, 0, 1);
    MPROS_LAUNCH_CHECK(c);
    obs_finish_kernel<<<grid_for(c, cells), 256, 0, st>>>(key, cells, c->obs_dist, c->region);
    MPROS_LAUNCH_CHECK(c);

    // N7
    size_t words = (size_t)H * ((W + 31) / 32);
    MPROS_CUDA(cudaMemsetAsync(c->edge_bits, 0, (size_t)H * pitch * 4, st));
    edge_kernel<<<grid_for(c, words * 32), 256, 0, st>>>(c->ds_bits, c->obs_dist, c->region, H, W, pitch, c->edge_bits, 0, H);
    MPROS_LAUNCH_CHECK(c);

    // N8
    dt_row_kernel<<<grid_for(c, cells), 256, 0, st>>>(c->edge_bits, H, W, pitch, nullptr, key);
    MPROS_LAUNCH_CHECK(c);
    dt_col_kernel<<<cdiv(W, 32), dim3(32, 32), 0, st>>>(key, H, W, 0, H, nullptr, 0, 1);
    MPROS_LAUNCH_CHECK(c);
    edge_finish_kernel<<<grid_for(c, cells), 256, 0, st>>>(key, cells, c->edge_dist);
    MPROS_LAUNCH_CHECK(c);

    // N9
    voronoi_cost_kernel<<<grid_for(c, cells), 256, 0, st>>>(c->ds_bits, c->edge_bits, c->obs_dist, c->edge_dist, H, W, pitch,
                                                        c->mp.voro_fallout_rate, c->mp.voro_max_region, c->voronoi);
    MPROS_LAUNCH_CHECK(c);
    return MPROS_OK;
}

// ---- N5-N9 on ROW BANDS of the down-sampled grid, one band per rank -----------------------------------
// The exchanges (all in-place all-gathers over the communicator, NCCL over NVLink on the box):
//   N5  (1) labels + same-component representatives of every band's first and last row   4 W int32 per rank
//       (2) number of components that start in the band + region ids of the boundary cells (2 W + 4) int32 per rank
//   N6  (3) the two column aggregates of the band's row-pass keys (the carry rows of the two sweeps)   2 W uint64 per rank
//       (4, 5) the band's rows of obs_dist_ and region_ (N7 looks two rows beyond the band)
//   N8  (6) as (3) for the edge distance
//   N9  (7-9) the band's rows of isedge_, edge_dist_ and the Voronoi cost: every rank ends with complete layers
// The result is bit-identical to map_build_voronoi on one GPU (tests/test_gpu_banded_voronoi.py).
//
// Seam union of the component labelling: node = a boundary cell of some band, id (g * 2 + side) * W + j. Every rank solves
// the same small union-find over the gathered nodes: a node starts under the smallest node of its own band with the same
// local root (computed by the owning band, exchange 1), vertical contacts across a seam are united, and the merged
// component's global root is the smallest local root (= its first cell in raster order).
__global__ void __launch_bounds__(256) ccl_boundary_mark_kernel(const uint32_t *__restrict__ occ, int n, int W, int pitch,
                                                                const int *__restrict__ parent, int node_base, int *__restrict__ bnode)
{
    const int total = 2 * W;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x)
    {
        const int s = t / W, j = t % W;
        const int i = s ? n - 1 : 0;
        if (n <= 0 || !((occ[(size_t)i * pitch + (j >> 5)] >> (j & 31)) & 1u))
            continue;
        const int root = uf_find(parent, i * W + j);
        atomicMin(&bnode[root], node_base + t);
    }
}
// pack = [label top | label bottom | representative top | representative bottom], W entries each; label = global cell index
// of the local root, -1 for free cells and for an empty band
__global__ void __launch_bounds__(256) ccl_boundary_pack_kernel(const uint32_t *__restrict__ occ, int n, int W, int pitch,
                                                                const int *__restrict__ parent, const int *__restrict__ bnode, int base,
                                                                int *__restrict__ pack)
{
    const int total = 2 * W;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x)
    {
        const int s = t / W, j = t % W;
        const int i = s ? n - 1 : 0;
        int lab = -1, rep = -1;
        if (n > 0 && ((occ[(size_t)i * pitch + (j >> 5)] >> (j & 31)) & 1u))
        {
            const int root = uf_find(parent, i * W + j);
            lab = root + base;
            rep = bnode[root];
        }
        pack[t] = lab;
        pack[2 * W + t] = rep;
    }
}
__global__ void __launch_bounds__(256) seam_init_kernel(const int *__restrict__ pack_all, int world, int W, int *__restrict__ sp)
{
    const int total = world * 2 * W;
    for (int nd = blockIdx.x * blockDim.x + threadIdx.x; nd < total; nd += gridDim.x * blockDim.x)
    {
        const int g = nd / (2 * W), t = nd % (2 * W);
        const int *pk = pack_all + (size_t)g * 4 * W;
        sp[nd] = pk[t] >= 0 ? pk[2 * W + t] : -1;
    }
}
__global__ void __launch_bounds__(256) seam_union_kernel(const int *__restrict__ pack_all, int world, int W, int *__restrict__ sp)
{
    const int total = (world - 1) * W;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x)
    {
        const int g = t / W, j = t % W;
        const int below = pack_all[(size_t)g * 4 * W + W + j];       // last row of band g
        const int above = pack_all[(size_t)(g + 1) * 4 * W + j];     // first row of band g + 1
        if (below >= 0 && above >= 0)
            uf_union(sp, (g * 2 + 1) * W + j, ((g + 1) * 2) * W + j);
    }
}
__global__ void __launch_bounds__(256) seam_glabel_kernel(const int *__restrict__ pack_all, int world, int W, const int *__restrict__ sp,
                                                          int *__restrict__ glabel)
{
    const int total = world * 2 * W;
    for (int nd = blockIdx.x * blockDim.x + threadIdx.x; nd < total; nd += gridDim.x * blockDim.x)
    {
        const int g = nd / (2 * W), t = nd % (2 * W);
        const int lab = pack_all[(size_t)g * 4 * W + t];
        if (lab >= 0)
            atomicMin(&glabel[uf_find(sp, nd)], lab);
    }
}
// my boundary cells: global root of the local root (same value from every cell of the root: benign race)
__global__ void __launch_bounds__(256) seam_apply_kernel(const int *__restrict__ pack_mine, int W, int node_base, int base,
                                                         const int *__restrict__ sp, const int *__restrict__ glabel, int *__restrict__ groot)
{
    const int total = 2 * W;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x)
    {
        const int lab = pack_mine[t];
        if (lab >= 0)
            groot[lab - base] = glabel[uf_find(sp, node_base + t)];
    }
}
// gid = [components that start in this band, 0, 0, 0 | id (without the earlier bands' count) of the boundary cells whose
// component starts in this band, -1 otherwise]
__global__ void __launch_bounds__(256) seam_gid_kernel(const int *__restrict__ pack_mine, int W, int base, const int *__restrict__ groot,
                                                       const int *__restrict__ local_rank, const int *__restrict__ tile_offset,
                                                       const int *__restrict__ n_local, int *__restrict__ gid)
{
    const int total = 2 * W;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x)
    {
        if (t < 4)
            gid[t] = t == 0 ? *n_local : 0;
        const int lab = pack_mine[t];
        int v = -1;
        if (lab >= 0)
        {
            const int root = lab - base;
            if (groot[root] == lab)
                v = tile_offset[root >> 10] + local_rank[root];
        }
        gid[4 + t] = v;
    }
}
__global__ void __launch_bounds__(256) seam_regid_kernel(const int *__restrict__ pack_all, const int *__restrict__ gid_all, int world, int W,
                                                         const int *__restrict__ sp, int *__restrict__ regid)
{
    const int total = world * 2 * W;
    for (int nd = blockIdx.x * blockDim.x + threadIdx.x; nd < total; nd += gridDim.x * blockDim.x)
    {
        const int g = nd / (2 * W), t = nd % (2 * W);
        if (pack_all[(size_t)g * 4 * W + t] < 0)
            continue;
        const int v = gid_all[(size_t)g * (2 * W + 4) + 4 + t];
        if (v < 0)
            continue;
        int before = 0;
        for (int e = 0; e < g; ++e)
            before += gid_all[(size_t)e * (2 * W + 4)];
        regid[uf_find(sp, nd)] = before + v;
    }
}
// final region id of my boundary-touching local roots; also the id offset of this band and the total number of regions
__global__ void __launch_bounds__(256) seam_root_region_kernel(const int *__restrict__ pack_mine, const int *__restrict__ gid_all, int world,
                                                               int rank, int W, int node_base, int base, const int *__restrict__ sp,
                                                               const int *__restrict__ regid, int *__restrict__ root_region,
                                                               int *__restrict__ id_offset /* [0] offset, [1] total */)
{
    const int total = 2 * W;
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        int before = 0, all = 0;
        for (int e = 0; e < world; ++e)
        {
            const int v = gid_all[(size_t)e * (2 * W + 4)];
            if (e < rank)
                before += v;
            all += v;
        }
        id_offset[0] = before;
        id_offset[1] = all;
    }
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x)
    {
        const int lab = pack_mine[t];
        if (lab >= 0)
            root_region[lab - base] = regid[uf_find(sp, node_base + t)];
    }
}

int map_build_voronoi_banded(Ctx *c, mpros_comm *cm, int rows_per_rank)
{
    cudaStream_t st = c->stream;
    const int H = c->H, W = c->W, pitch = c->pitch;
    const int world = cm->world, rank = cm->rank;
    if ((unsigned long long)H * (unsigned long long)W >= 0x7f000000ull) // labels are int32 cell indices below the 0x7f7f7f7f fill
    {
        set_error("banded map: the down-sampled grid must have fewer than 2^31 - 2^24 cells");
        return MPROS_ERR_INVALID;
    }
    int lo, hi;
    band_of(H, rows_per_rank, rank, &lo, &hi);
    const int n = hi - lo;
    const size_t cells = (size_t)n * W;
    const int base = lo * W;
    const int n_tiles = (int)((cells + 1023) / 1024);
    const int nodes = world * 2 * W, node_base = rank * 2 * W;
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    // scratch: key u64[cells] | parent, local_rank, bnode (later root_region), groot i32[cells] | tile_count | counters |
    //          pack_all i32[world * 4 W] | gid_all i32[world * (2 W + 4)] | sp, glabel, regid i32[nodes] | agg u64[world * 2 W]
    const size_t b_key = up(cells * 8), b_cell = up(cells * 4), b_tiles = up((size_t)n_tiles * 4 + 4), b_cnt = 256,
                 b_pack = up((size_t)world * 4 * W * 4), b_gid = up((size_t)world * (2 * W + 4) * 4), b_node = up((size_t)nodes * 4),
                 b_agg = up((size_t)world * 2 * W * 8);
    const size_t need = b_key + 4 * b_cell + b_tiles + b_cnt + b_pack + b_gid + 3 * b_node + b_agg;
    int rc = ensure_scratch(c, need);
    if (rc)
        return rc;
    char *p = static_cast<char *>(c->scratch);
    unsigned long long *key = (unsigned long long *)p;
    int *parent = (int *)(p += b_key);
    int *local_rank = (int *)(p += b_cell);
    int *bnode = (int *)(p += b_cell);
    int *groot = (int *)(p += b_cell);
    int *tile_count = (int *)(p += b_cell);
    int *counters = (int *)(p += b_tiles); // [0] components that start in this band, [1] id offset, [2] total regions
    int *pack_all = (int *)(p += b_cnt);
    int *gid_all = (int *)(p += b_pack);
    int *sp = (int *)(p += b_gid);
    int *glabel = (int *)(p += b_node);
    int *regid = (int *)(p += b_node);
    unsigned long long *agg = (unsigned long long *)(p += b_node);
    int *root_region = bnode;
    int *pack_mine = pack_all + (size_t)rank * 4 * W;
    int *gid_mine = gid_all + (size_t)rank * (2 * W + 4);

    const uint32_t *occ = c->ds_bits + (size_t)lo * pitch;
    int *region = c->region + (size_t)lo * W;
    const int g2w = grid_for(c, (size_t)2 * W), gnodes = grid_for(c, (size_t)nodes);

    // N5: local components of the band
    if (n > 0)
    {
        ccl_init_kernel<<<grid_for(c, cells), 256, 0, st>>>(occ, n, W, pitch, parent);
        MPROS_LAUNCH_CHECK(c);
        ccl_merge_kernel<<<grid_for(c, cells), 256, 0, st>>>(occ, n, W, pitch, parent);
        MPROS_LAUNCH_CHECK(c);
    }
    MPROS_CUDA(cudaMemsetAsync(bnode, 0x7f, b_cell, st));
    MPROS_CUDA(cudaMemsetAsync(groot, 0xff, b_cell, st));
    ccl_boundary_mark_kernel<<<g2w, 256, 0, st>>>(occ, n, W, pitch, parent, node_base, bnode);
    MPROS_LAUNCH_CHECK(c);
    ccl_boundary_pack_kernel<<<g2w, 256, 0, st>>>(occ, n, W, pitch, parent, bnode, base, pack_mine);
    MPROS_LAUNCH_CHECK(c);
    rc = comm_allgather(c, cm, pack_all, (size_t)4 * W * 4); // exchange 1
    if (rc)
        return rc;
    seam_init_kernel<<<gnodes, 256, 0, st>>>(pack_all, world, W, sp);
    MPROS_LAUNCH_CHECK(c);
    if (world > 1)
    {
        seam_union_kernel<<<grid_for(c, (size_t)(world - 1) * W), 256, 0, st>>>(pack_all, world, W, sp);
        MPROS_LAUNCH_CHECK(c);
    }
    MPROS_CUDA(cudaMemsetAsync(glabel, 0x7f, b_node, st));
    seam_glabel_kernel<<<gnodes, 256, 0, st>>>(pack_all, world, W, sp, glabel);
    MPROS_LAUNCH_CHECK(c);
    seam_apply_kernel<<<g2w, 256, 0, st>>>(pack_mine, W, node_base, base, sp, glabel, groot);
    MPROS_LAUNCH_CHECK(c);
    // components whose first cell (raster order) lies in this band, ranked in raster order
    MPROS_CUDA(cudaMemsetAsync(counters, 0, b_cnt, st));
    if (n > 0)
    {
        ccl_flatten_rank_kernel<<<n_tiles, 1024, 0, st>>>(parent, cells, local_rank, tile_count, groot, base);
        MPROS_LAUNCH_CHECK(c);
        scan_tiles_kernel<<<1, 1024, 0, st>>>(tile_count, n_tiles, counters);
        MPROS_LAUNCH_CHECK(c);
    }
    seam_gid_kernel<<<g2w, 256, 0, st>>>(pack_mine, W, base, groot, local_rank, tile_count, counters, gid_mine);
    MPROS_LAUNCH_CHECK(c);
    rc = comm_allgather(c, cm, gid_all, (size_t)(2 * W + 4) * 4); // exchange 2
    if (rc)
        return rc;
    MPROS_CUDA(cudaMemsetAsync(regid, 0xff, b_node, st));
    seam_regid_kernel<<<gnodes, 256, 0, st>>>(pack_all, gid_all, world, W, sp, regid);
    MPROS_LAUNCH_CHECK(c);
    MPROS_CUDA(cudaMemsetAsync(root_region, 0xff, b_cell, st)); // bnode is no longer needed
    seam_root_region_kernel<<<g2w, 256, 0, st>>>(pack_mine, gid_all, world, rank, W, node_base, base, sp, regid, root_region, counters + 1);
    MPROS_LAUNCH_CHECK(c);
    if (n > 0)
    {
        ccl_region_kernel<<<grid_for(c, cells), 256, 0, st>>>(parent, cells, local_rank, tile_count, region, root_region, counters + 1);
        MPROS_LAUNCH_CHECK(c);
    }
    MPROS_CUDA(cudaMemcpyAsync(&c->num_regions, counters + 2, sizeof(int), cudaMemcpyDeviceToHost, st));

    // N6
    unsigned long long *agg_mine = agg + (size_t)rank * 2 * W;
    if (n > 0)
    {
        dt_row_kernel<<<grid_for(c, cells), 256, 0, st>>>(occ, n, W, pitch, region, key);
        MPROS_LAUNCH_CHECK(c);
    }
    dt_col_agg_kernel<<<cdiv(W, 32), dim3(32, 32), 0, st>>>(key, n, W, lo, H, agg_mine);
    MPROS_LAUNCH_CHECK(c);
    rc = comm_allgather(c, cm, agg, (size_t)2 * W * 8); // exchange 3: the carry rows of both sweeps
    if (rc)
        return rc;
    if (n > 0)
    {
        dt_col_kernel<<<cdiv(W, 32), dim3(32, 32), 0, st>>>(key, n, W, lo, H, agg, rank, world);
        MPROS_LAUNCH_CHECK(c);
        obs_finish_kernel<<<grid_for(c, cells), 256, 0, st>>>(key, cells, c->obs_dist + (size_t)lo * W, region);
        MPROS_LAUNCH_CHECK(c);
    }
    rc = comm_allgather_rows(c, cm, c->obs_dist, (size_t)W * 2, H, rows_per_rank); // exchange 4
    if (!rc)
        rc = comm_allgather_rows(c, cm, c->region, (size_t)W * 4, H, rows_per_rank); // exchange 5
    if (rc)
        return rc;

    // N7 (reads obs_dist_ / region_ up to two rows beyond the band)
    uint32_t *edge = c->edge_bits + (size_t)lo * pitch;
    if (n > 0)
    {
        MPROS_CUDA(cudaMemsetAsync(edge, 0, (size_t)n * pitch * 4, st));
        size_t words = (size_t)n * ((W + 31) / 32);
        edge_kernel<<<grid_for(c, words * 32), 256, 0, st>>>(c->ds_bits, c->obs_dist, c->region, H, W, pitch, c->edge_bits, lo, n);
        MPROS_LAUNCH_CHECK(c);
        // N8
        dt_row_kernel<<<grid_for(c, cells), 256, 0, st>>>(edge, n, W, pitch, nullptr, key);
        MPROS_LAUNCH_CHECK(c);
    }
    dt_col_agg_kernel<<<cdiv(W, 32), dim3(32, 32), 0, st>>>(key, n, W, lo, H, agg_mine);
    MPROS_LAUNCH_CHECK(c);
    rc = comm_allgather(c, cm, agg, (size_t)2 * W * 8); // exchange 6
    if (rc)
        return rc;
    if (n > 0)
    {
        dt_col_kernel<<<cdiv(W, 32), dim3(32, 32), 0, st>>>(key, n, W, lo, H, agg, rank, world);
        MPROS_LAUNCH_CHECK(c);
        edge_finish_kernel<<<grid_for(c, cells), 256, 0, st>>>(key, cells, c->edge_dist + (size_t)lo * W);
        MPROS_LAUNCH_CHECK(c);
        // N9
        voronoi_cost_kernel<<<grid_for(c, cells), 256, 0, st>>>(occ, edge, c->obs_dist + (size_t)lo * W, c->edge_dist + (size_t)lo * W, n, W,
                                                            pitch, c->mp.voro_fallout_rate, c->mp.voro_max_region, c->voronoi + (size_t)lo * W);
        MPROS_LAUNCH_CHECK(c);
    }
    rc = comm_allgather_rows(c, cm, c->edge_bits, (size_t)pitch * 4, H, rows_per_rank); // exchanges 7-9
    if (!rc)
        rc = comm_allgather_rows(c, cm, c->edge_dist, (size_t)W * 2, H, rows_per_rank);
    if (!rc)
        rc = comm_allgather_rows(c, cm, c->voronoi, (size_t)W * 4, H, rows_per_rank);
    return rc;
}

} // namespace mpros
