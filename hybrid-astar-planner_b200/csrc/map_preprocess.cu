// NavMap preprocessing on bit-packed layers: threshold (N1), down-sampling (N2), disk inflation (N3),
// plus the NavMap getters that move a layer back to the host in the reference's element types.
//
// All three kernels are HBM/L2-streaming integer work. The only full-size HBM read is the int8 occupancy
// grid itself (1 B/cell, read once with 128-bit loads); everything downstream works on 1 bit/cell.
#include <cmath>
#include <cstring>

#include "dist.cuh"

namespace mpros
{

// ---- N1: NavMap::setOccupancyMap threshold loop (nav_map.cpp:114-131) ------------------------------------
// new = (d > free_thres || d < 0) ? 1 : (d < free_thres ? 0 : old). With integer d this is
// d >= occ_min || d < 0, resp. d <= free_max (thresholds pre-rounded on the host, exact).
__device__ __forceinline__ uint32_t classify16(uint4 v, int occ_min, int free_max, uint32_t &free_mask)
{
    uint32_t occ = 0, fre = 0;
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int b = 0; b < 4; ++b)
        {
            int d = (int)(int8_t)((w[q] >> (8 * b)) & 0xff);
            bool o = (d >= occ_min) || (d < 0);
            bool f = !o && (d <= free_max);
            occ |= (uint32_t)o << (4 * q + b);
            fre |= (uint32_t)f << (4 * q + b);
        }
    free_mask = fre;
    return occ;
}

// Fast path: W % 32 == 0 and 16-byte aligned base. One thread = one output word = 32 cells = two 128-bit loads;
// a warp covers 1 KiB of contiguous input.
__global__ void __launch_bounds__(256) threshold_pack_vec_kernel(const int8_t *__restrict__ data, uint32_t *__restrict__ bits,
                                                                 int H, int W, int pitch, int occ_min, int free_max,
                                                                 int have_old)
{
    const int words_per_row = W >> 5;
    const size_t total = (size_t)H * words_per_row;
    for (size_t wd = (size_t)blockIdx.x * blockDim.x + threadIdx.x; wd < total; wd += (size_t)gridDim.x * blockDim.x)
    {
        int i = (int)(wd / words_per_row), w = (int)(wd % words_per_row);
        const uint4 *src = reinterpret_cast<const uint4 *>(data + (size_t)i * W) + 2 * w;
        uint4 v0 = __ldg(src), v1 = __ldg(src + 1);
        uint32_t f0, f1;
        uint32_t occ = classify16(v0, occ_min, free_max, f0);
        occ |= classify16(v1, occ_min, free_max, f1) << 16;
        uint32_t fre = f0 | (f1 << 16);
        size_t widx = (size_t)i * pitch + w;
        uint32_t old = have_old ? bits[widx] : 0u;
        bits[widx] = occ | (old & ~fre);
    }
}

// General path: one lane per cell, warp ballot forms the word.
__global__ void __launch_bounds__(256) threshold_pack_kernel(const int8_t *__restrict__ data, uint32_t *__restrict__ bits, int H,
                                                             int W, int pitch, int occ_min, int free_max, int have_old)
{
    const int words_per_row = (W + 31) >> 5;
    const size_t total_words = (size_t)H * words_per_row;
    const int lane = threadIdx.x & 31;
    size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t wd = warp; wd < total_words; wd += nwarps)
    {
        int i = (int)(wd / words_per_row);
        int w = (int)(wd % words_per_row);
        int j = (w << 5) + lane;
        bool o = false, f = false;
        if (j < W)
        {
            int d = (int)__ldg(data + (size_t)i * W + j);
            o = (d >= occ_min) || (d < 0);
            f = !o && (d <= free_max);
        }
        uint32_t occ = __ballot_sync(0xffffffffu, o);
        uint32_t fre = __ballot_sync(0xffffffffu, f);
        if (lane == 0)
        {
            size_t widx = (size_t)i * pitch + w;
            uint32_t old = have_old ? bits[widx] : 0u;
            bits[widx] = occ | (old & ~fre);
        }
    }
}

// std::vector::resize keeps the linear prefix when the map geometry changes (nav_map.cpp:107): cell (i,j) of
// the new grid starts from the old value at linear index i*W_new + j.
__global__ void relinearize_kernel(const uint32_t *__restrict__ old_bits, int oldH, int oldW, int old_pitch,
                                   uint32_t *__restrict__ new_bits, int H, int W, int pitch)
{
    const int words_per_row = (W + 31) >> 5;
    size_t total = (size_t)H * words_per_row;
    for (size_t wd = (size_t)blockIdx.x * blockDim.x + threadIdx.x; wd < total; wd += (size_t)gridDim.x * blockDim.x)
    {
        int i = (int)(wd / words_per_row), w = (int)(wd % words_per_row);
        uint32_t out = 0;
        for (int b = 0; b < 32; ++b)
        {
            int j = (w << 5) + b;
            if (j >= W)
                break;
            size_t k = (size_t)i * W + j;
            if (k < (size_t)oldH * oldW)
            {
                int oi = (int)(k / oldW), oj = (int)(k % oldW);
                out |= ((old_bits[(size_t)oi * old_pitch + (oj >> 5)] >> (oj & 31)) & 1u) << b;
            }
        }
        new_bits[(size_t)i * pitch + w] = out;
    }
}

// ---- N2: NavMap::downsampleMap / downsampleCellValue (nav_map.cpp:592-652) ------------------------------
// max-pool ds x ds. Off-by-one quirks kept: old_j == W is allowed and reads linear index i*W + W (first cell
// of the next row); old_i == H is allowed but every index > size is skipped. Index == size is undefined
// behaviour in the reference (one past the end); it contributes 0 here.
__global__ void __launch_bounds__(256) downsample_kernel(const uint32_t *__restrict__ raw, int H0, int W0, int pitch0,
                                                         uint32_t *__restrict__ out, int H, int W, int pitch, int ds)
{
    const int words_per_row = (W + 31) >> 5;
    const size_t total_words = (size_t)H * words_per_row;
    const int lane = threadIdx.x & 31;
    size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const size_t size0 = (size_t)H0 * W0;
    for (size_t wd = warp; wd < total_words; wd += nwarps)
    {
        int i = (int)(wd / words_per_row);
        int w = (int)(wd % words_per_row);
        int j = (w << 5) + lane;
        bool v = false;
        if (j < W)
        {
            int i_off = i * ds, j_off = j * ds;
            for (int a = 0; a < ds && !v; ++a)
            {
                int oi = a + i_off;
                if (oi > H0)
                    break;
                for (int b = 0; b < ds; ++b)
                {
                    int oj = b + j_off;
                    if (oj > W0)
                        break;
                    size_t k = (size_t)oi * W0 + oj;
                    if (k >= size0)
                        continue; // > size skipped by the reference; == size is UB there, 0 here
                    int ri = (int)(k / W0), rj = (int)(k % W0);
                    if ((raw[(size_t)ri * pitch0 + (rj >> 5)] >> (rj & 31)) & 1u)
                    {
                        v = true;
                        break;
                    }
                }
            }
        }
        uint32_t word = __ballot_sync(0xffffffffu, v);
        if (lane == 0)
            out[(size_t)i * pitch + w] = word;
    }
}

// Fast path of N2 when ds is a power of two <= 32 that divides both map dimensions (then none of the off-by-one quirks
// above can trigger): one thread per OUTPUT word = ds input rows x ds input words; groups of ds bits are OR-folded in
// the word and their representatives gathered. Reads every input word exactly once, coalesced.
__global__ void __launch_bounds__(256) downsample_pow2_kernel(const uint32_t *__restrict__ raw, int pitch0, uint32_t *__restrict__ out,
                                                              int H, int W, int pitch, int ds, int log2ds)
{
    const int words_per_row = (W + 31) >> 5;
    const size_t total_words = (size_t)H * words_per_row;
    const int per = 32 >> log2ds; // output bits produced by one input word
    for (size_t wd = (size_t)blockIdx.x * blockDim.x + threadIdx.x; wd < total_words; wd += (size_t)gridDim.x * blockDim.x)
    {
        const int i = (int)(wd / words_per_row), w = (int)(wd % words_per_row);
        const int in_words = min(ds, (W - (w << 5) + per - 1) / per); // input words that carry cells of this output word
        uint32_t word = 0;
        for (int q = 0; q < in_words; ++q)
        {
            uint32_t v = 0;
            for (int a = 0; a < ds; ++a)
                v |= __ldg(raw + (size_t)(i * ds + a) * pitch0 + (size_t)w * ds + q);
            for (int sh = 1; sh < ds; sh <<= 1) // bit 0 of every group of ds bits = OR of the group
                v |= v >> sh;
            uint32_t g = 0;
            for (int t = 0; t < per; ++t)
                g |= ((v >> (t << log2ds)) & 1u) << t;
            word |= g << (q * per);
        }
        out[(size_t)i * pitch + w] = word;
    }
}

// ---- N3: NavMap::inflateObstacle (nav_map.cpp:164-214) ----------------------------------------------------
// Binary dilation by SE = {(di,dj): (|di|-1/2)^2 + (|dj|-1/2)^2 <= r^2}, clipped to the map. The SE is
// monotone in |di| and |dj|, so SE = union over column offsets s of the vertical run |di| <= a(|s|).
// Per output word: V_a = OR of the input rows i-a..i+a (word-aligned, no shifting), grown incrementally
// while |s| walks from the widest offset to 0, and out |= funnel-shift(V_a(|s|), +-s).
// ext_tab[k] = a(k) (or -1 when column offset k is outside the SE), in constant memory.
constexpr int kMaxInflRadius = 255;
__constant__ int16_t c_ext_tab[kMaxInflRadius + 1];

template <int K, bool SMALL> // K = words of halo on each side; SMALL: smax <= 31, all shifts stay within +-1 word
__global__ void __launch_bounds__(256) inflate_kernel(const uint32_t *__restrict__ raw, uint32_t *__restrict__ out, int H,
                                                      int W, int pitch, int smax, int row0, int nrows)
{
    // output rows [row0, row0 + nrows) (a row band when the map is tiled across GPUs); input rows are clipped to the map
    const int words_per_row = (W + 31) >> 5;
    const size_t total = (size_t)nrows * words_per_row;
    for (size_t wd = (size_t)blockIdx.x * blockDim.x + threadIdx.x; wd < total; wd += (size_t)gridDim.x * blockDim.x)
    {
        int i = row0 + (int)(wd / words_per_row), w = (int)(wd % words_per_row);
        uint32_t V[2 * K + 2]; // V[k] = OR over current rows of word w - K + k; one spare so V[q+1] is always valid
#pragma unroll
        for (int k = 0; k < 2 * K + 2; ++k)
            V[k] = 0;
        uint32_t acc = 0;
        int a_cur = -1; // rows |di| <= a_cur are folded into V
        for (int s = smax; s >= 0; --s)
        {
            int a = c_ext_tab[s];
            if (a < 0)
                continue;
            for (int di = a_cur + 1; di <= a; ++di)
            {
#pragma unroll
                for (int sign = 0; sign < 2; ++sign)
                {
                    if (di == 0 && sign == 1)
                        continue;
                    int r = sign ? i - di : i + di;
                    if (r < 0 || r >= H)
                        continue;
                    const uint32_t *row = raw + (size_t)r * pitch;
#pragma unroll
                    for (int k = 0; k < 2 * K + 1; ++k)
                    {
                        int ww = w - K + k;
                        if (ww >= 0 && ww < words_per_row)
                            V[k] |= __ldg(row + ww);
                    }
                }
            }
            a_cur = a > a_cur ? a : a_cur;
            // out[j] |= V[j - s] and V[j + s]
            int q = (K == 1 && SMALL) ? 0 : (s >> 5), t = s & 31;
            // left shift by s: bits come from words w-q (high part) and w-q-1 (low part)
            uint32_t hi = V[K - q];
            uint32_t lo = (K - q - 1 >= 0) ? V[K - q - 1] : 0u;
            acc |= __funnelshift_l(lo, hi, t);
            // right shift by s: words w+q (low part) and w+q+1 (high part)
            uint32_t lo2 = V[K + q];
            uint32_t hi2 = (K + q + 1 <= 2 * K) ? V[K + q + 1] : 0u;
            acc |= __funnelshift_r(lo2, hi2, t);
        }
        if (w == words_per_row - 1 && (W & 31))
            acc &= (1u << (W & 31)) - 1u;
        out[(size_t)i * pitch + w] = acc;
    }
}

// ---- getters: unpack a bit layer to int8, widen u16 -> int32 ------------------------------------------------
__global__ void unpack_bits_kernel(const uint32_t *__restrict__ bits, int H, int W, int pitch, int8_t *__restrict__ out)
{
    size_t total = (size_t)H * W;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x)
    {
        int i = (int)(k / W), j = (int)(k % W);
        out[k] = (int8_t)((bits[(size_t)i * pitch + (j >> 5)] >> (j & 31)) & 1u);
    }
}
__global__ void widen_u16_kernel(const uint16_t *__restrict__ in, size_t n, int32_t *__restrict__ out)
{
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
        out[k] = (int32_t)in[k];
}
__global__ void point_lookup_kernel(const uint32_t *__restrict__ bits, int H, int W, int pitch, const int32_t *__restrict__ ij,
                                    size_t n, uint8_t *__restrict__ out)
{
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
    {
        int i = ij[2 * k], j = ij[2 * k + 1];
        bool hit = true; // outside the map counts as collision (nav_map.cpp:249-277)
        if (i >= 0 && i < H && j >= 0 && j < W)
            hit = (bits[(size_t)i * pitch + (j >> 5)] >> (j & 31)) & 1u;
        out[k] = hit;
    }
}

// Padded tile-word copy of a bit-packed collision layer for the fixed-point footprint walk (layout: MapView::raw_tw).
// Outside the map reads 1 (collision) except index -1, which repeats cell 0 (truncation toward zero).
__global__ void __launch_bounds__(256) build_tilewords_kernel(const uint32_t *__restrict__ bits, int H, int W, int pitch,
                                                              uint32_t *__restrict__ tw, int tw_pitch, int tw_rows)
{
    const size_t total = (size_t)tw_pitch * tw_rows;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x)
    {
#if MPROS_TW_BLOCK16
        const size_t blk = t >> 3;
        const int sub = (int)(t & 7);
        const int wi = (int)(blk / tw_pitch) * 4 + (sub >> 1), wj = (int)(blk % tw_pitch) * 2 + (sub & 1);
#else
        const int wi = (int)(t / tw_pitch), wj = (int)(t % tw_pitch);
#endif
        const int j0 = wj * 8 - kTwMargin; // multiple of 8
        uint32_t v = 0;
#pragma unroll
        for (int r = 0; r < 4; ++r)
        {
            int i = wi * 4 + r - kTwMargin;
            if (i == -1)
                i = 0;
            uint32_t row = 0xffu;
            if (i >= 0 && i < H)
            {
                if (j0 >= 0 && j0 + 8 <= W)
                    row = (bits[(size_t)i * pitch + (j0 >> 5)] >> (j0 & 31)) & 0xffu;
                else
                {
                    row = 0;
                    for (int cc = 0; cc < 8; ++cc)
                    {
                        int j = j0 + cc;
                        if (j == -1)
                            j = 0;
                        uint32_t b = 1;
                        if (j >= 0 && j < W)
                            b = (bits[(size_t)i * pitch + (j >> 5)] >> (j & 31)) & 1u;
                        row |= b << cc;
                    }
                }
            }
            v |= row << (8 * r);
        }
        tw[t] = v;
    }
}

// ---- NavMap::vecToMsg (nav_map.cpp:654-693): a layer as the int8 payload of a nav_msgs/OccupancyGrid ------------------
// data[i] = floor(float(layer[i]) * 100 / maxval), maxval = min(100, max(layer)), stored into int8 the way the x86 build of
// the reference does (float -> int32 truncation, low byte): goal-map values above 127 wrap. A 0/1 layer becomes 0/100.
__global__ void __launch_bounds__(256) msg_bits_kernel(const uint32_t *__restrict__ bits, int H, int W, int pitch, int8_t *__restrict__ out)
{
    size_t total = (size_t)H * W;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x)
    {
        int i = (int)(k / W), j = (int)(k % W);
        out[k] = ((bits[(size_t)i * pitch + (j >> 5)] >> (j & 31)) & 1u) ? (int8_t)100 : (int8_t)0;
    }
}
__global__ void __launch_bounds__(256) msg_max_u16_kernel(const uint16_t *__restrict__ in, size_t n, int *__restrict__ maxv)
{
    int m = 0;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
        m = max(m, (int)in[k]);
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0)
        atomicMax(maxv, m);
}
__global__ void __launch_bounds__(256) msg_max_f32_kernel(const float *__restrict__ in, size_t n, int *__restrict__ maxv)
{
    // the Voronoi field is non-negative, so the IEEE bit patterns order like the values
    int m = 0;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
        m = max(m, __float_as_int(fmaxf(in[k], 0.f)));
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0)
        atomicMax(maxv, m);
}
__device__ __forceinline__ int8_t msg_store(float f)
{
    // x86 cvttss2si: NaN / out of range -> INT_MIN (low byte 0); then the low byte is kept
    int v = (f != f || f >= 2147483648.f || f < -2147483648.f) ? (int)0x80000000 : (int)f;
    return (int8_t)(v & 0xff);
}
__global__ void __launch_bounds__(256) msg_scale_u16_kernel(const uint16_t *__restrict__ in, size_t n, const int *__restrict__ maxv,
                                                            int8_t *__restrict__ out)
{
    const int maxval = min(100, *maxv);
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
        out[k] = msg_store(floorf(__fdiv_rn(__fmul_rn((float)in[k], 100.f), (float)maxval)));
}
__global__ void __launch_bounds__(256) msg_scale_f32_kernel(const float *__restrict__ in, size_t n, const int *__restrict__ maxv,
                                                            int8_t *__restrict__ out)
{
    const float maxval = fminf(100.f, __int_as_float(*maxv));
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
        out[k] = msg_store(floorf(__fdiv_rn(__fmul_rn(in[k], 100.f), maxval)));
}

static int realloc_layers(Ctx *c)
{
    size_t words0 = (size_t)c->H0 * c->pitch0, words = (size_t)c->H * c->pitch, cells = (size_t)c->H * c->W;
    if (words0 > c->cap0_words)
    {
        if (c->infl_bits)
            MPROS_CUDA(cudaFree(c->infl_bits));
        c->infl_bits = nullptr;
        MPROS_CUDA(cudaMalloc(&c->infl_bits, words0 * 4));
        // raw_bits is re-created by the caller (it needs the old content for relinearisation)
        c->cap0_words = words0;
    }
    {
#if MPROS_TW_BLOCK16
        c->tw_pitch = (c->W0 + 2 * kTwMargin + 15) / 16;                 // blocks per block row
        const int tw_rows = ((c->H0 + 2 * kTwMargin + 15) / 16) * 8;      // so that tw_pitch * tw_rows = words
#else
        c->tw_pitch = (c->W0 + 2 * kTwMargin + 7) / 8;
        const int tw_rows = (c->H0 + 2 * kTwMargin + 3) / 4;
#endif
        size_t tw_words = (size_t)c->tw_pitch * tw_rows;
        if (tw_words > c->cap_tw)
        {
            if (c->raw_tw)
                MPROS_CUDA(cudaFree(c->raw_tw));
            if (c->infl_tw)
                MPROS_CUDA(cudaFree(c->infl_tw));
            c->raw_tw = c->infl_tw = nullptr;
            MPROS_CUDA(cudaMalloc(&c->raw_tw, tw_words * 4));
            MPROS_CUDA(cudaMalloc(&c->infl_tw, tw_words * 4));
            c->cap_tw = tw_words;
        }
    }
    if (words > c->cap_words)
    {
        if (c->ds_bits)
            MPROS_CUDA(cudaFree(c->ds_bits));
        if (c->edge_bits)
            MPROS_CUDA(cudaFree(c->edge_bits));
        c->ds_bits = c->edge_bits = nullptr;
        MPROS_CUDA(cudaMalloc(&c->ds_bits, words * 4));
        MPROS_CUDA(cudaMalloc(&c->edge_bits, words * 4));
        c->cap_words = words;
    }
    if (cells > c->cap_cells)
    {
        void **bufs[] = {(void **)&c->goal, (void **)&c->obs_dist, (void **)&c->edge_dist, (void **)&c->region,
                         (void **)&c->voronoi};
        size_t esz[] = {2, 2, 2, 4, 4};
        for (int k = 0; k < 5; ++k)
        {
            if (*bufs[k])
                MPROS_CUDA(cudaFree(*bufs[k]));
            *bufs[k] = nullptr;
            MPROS_CUDA(cudaMalloc(bufs[k], cells * esz[k]));
        }
        c->cap_cells = cells;
    }
    return MPROS_OK;
}

// ---- NavMap::setOccupancyMap in stages, so that a row band can be processed per GPU (mpros_map_band_*) -------------
// stage 0: geometry + allocation (nav_map.cpp:54-112). Returns whether an earlier map's cells must be merged (the
// reference only overwrites cells that are definitely occupied or definitely free, :114-131).
static int map_begin(Ctx *c, int width, int height, double resolution, double ox, double oy, double oyaw, int *have_old_out)
{
    cudaStream_t st = c->stream;
    const bool first = !c->get_map;
    const bool geom_change = c->get_map && (c->H0 != height || c->W0 != width || c->res0 != resolution || c->ox != ox ||
                                            c->oy != oy || c->oyaw != oyaw);
    int have_old = 0;
    if (first || geom_change)
    {
        int oldH = c->H0, oldW = c->W0, old_pitch = c->pitch0;
        uint32_t *old_bits = c->raw_bits;
        c->H0 = height;
        c->W0 = width;
        c->pitch0 = pitch_words(width);
        c->res0 = resolution;
        c->ox = ox;
        c->oy = oy;
        c->oyaw = oyaw;
        c->osin = std::sin(oyaw);
        c->ocos = std::cos(oyaw);
        const int ds = c->mp.downsampling_factor;
        c->H = (int)std::ceil(static_cast<double>(height) / ds);
        c->W = (int)std::ceil(static_cast<double>(width) / ds);
        c->pitch = pitch_words(c->W);
        c->res = resolution * ds;
        int rc = realloc_layers(c);
        if (rc)
            return rc;
        uint32_t *nb = nullptr;
        MPROS_CUDA(cudaMalloc(&nb, (size_t)c->H0 * c->pitch0 * 4));
        if (geom_change && old_bits)
        {
            relinearize_kernel<<<grid_for(c, (size_t)c->H0 * c->pitch0), 256, 0, st>>>(old_bits, oldH, oldW, old_pitch, nb, c->H0,
                                                                                  c->W0, c->pitch0);
            MPROS_LAUNCH_CHECK(c);
            MPROS_CUDA(cudaStreamSynchronize(st));
            have_old = 1;
        }
        if (old_bits)
            MPROS_CUDA(cudaFree(old_bits));
        c->raw_bits = nb;
    }
    else
        have_old = 1;
    // padding words/bits must stay zero: clear the layer when there is no old state to merge
    if (!have_old)
        MPROS_CUDA(cudaMemsetAsync(c->raw_bits, 0, (size_t)c->H0 * c->pitch0 * 4, st));
    *have_old_out = have_old;
    return MPROS_OK;
}

// stage 1 (N1) for rows [row0, row0 + nrows): d_rows points at the first of these rows
static int map_threshold_rows(Ctx *c, const int8_t *d_rows, int row0, int nrows, int have_old)
{
    cudaStream_t st = c->stream;
    const int width = c->W0;
    // exact integer form of the double comparisons against free_thres_
    const double t = c->mp.free_thresh;
    double occ_min_d = std::floor(t) + 1.0, free_max_d = std::ceil(t) - 1.0;
    int occ_min = occ_min_d > 200 ? 200 : (occ_min_d < -200 ? -200 : (int)occ_min_d);
    int free_max = free_max_d > 200 ? 200 : (free_max_d < -200 ? -200 : (int)free_max_d);
    uint32_t *bits = c->raw_bits + (size_t)row0 * c->pitch0;
    if ((width % 32) == 0 && (reinterpret_cast<uintptr_t>(d_rows) % 16) == 0)
    {
        size_t words = (size_t)nrows * (width >> 5);
        threshold_pack_vec_kernel<<<grid_for(c, words), 256, 0, st>>>(d_rows, bits, nrows, width, c->pitch0, occ_min, free_max, have_old);
    }
    else
    {
        size_t words = (size_t)nrows * ((width + 31) / 32);
        threshold_pack_kernel<<<grid_for(c, words * 32), 256, 0, st>>>(d_rows, bits, nrows, width, c->pitch0, occ_min, free_max, have_old);
    }
    MPROS_LAUNCH_CHECK(c);
    return MPROS_OK;
}

// stage 2a (N3 set-up): structuring-element table exactly as the reference evaluates it (quarter-integers, exact in
// double). smax = widest column offset of the SE (-1: empty SE, radius 0 inflates nothing); *radius_px = rows of halo.
static int map_inflate_prepare(Ctx *c, int *smax_out, int *radius_px)
{
    int r = (int)std::ceil(c->mp.obstacle_inflation_radius / c->res0);
    if (r < 0)
        r = 0;
    if (r > kMaxInflRadius)
    {
        set_error("obstacle_inflation_radius / resolution exceeds 255 pixels (unsupported)");
        return MPROS_ERR_INVALID;
    }
    double r_sq = r * r;
    int16_t tab[kMaxInflRadius + 1];
    int smax = -1;
    for (int s = 0; s <= kMaxInflRadius; ++s)
    {
        tab[s] = -1;
        if (s > r)
            continue;
        double j_dist = std::abs(s) - 0.5;
        for (int a = 0; a <= r; ++a)
        {
            double i_dist = std::abs(a) - 0.5;
            if (i_dist * i_dist + j_dist * j_dist <= r_sq)
                tab[s] = (int16_t)a;
        }
        if (tab[s] >= 0)
            smax = s;
    }
    MPROS_CUDA(cudaMemcpyToSymbolAsync(c_ext_tab, tab, sizeof(tab), 0, cudaMemcpyHostToDevice, c->stream));
    *smax_out = smax;
    if (radius_px)
        *radius_px = r;
    return MPROS_OK;
}

// stage 2b (N3) for output rows [row0, row0 + nrows); reads raw rows up to the SE radius beyond the band
static int map_inflate_rows(Ctx *c, int smax, int row0, int nrows)
{
    cudaStream_t st = c->stream;
    MPROS_CUDA(cudaMemsetAsync(c->infl_bits + (size_t)row0 * c->pitch0, 0, (size_t)nrows * c->pitch0 * 4, st));
    if (smax < 0 || nrows <= 0)
        return MPROS_OK;
    size_t words = (size_t)nrows * ((c->W0 + 31) / 32);
    int K = (smax + 31) / 32;
    if (K < 1)
        K = 1;
    int g = grid_for(c, words);
    switch (K)
    {
    case 1:
        if (smax <= 31)
            inflate_kernel<1, true><<<g, 256, 0, st>>>(c->raw_bits, c->infl_bits, c->H0, c->W0, c->pitch0, smax, row0, nrows);
        else
            inflate_kernel<1, false><<<g, 256, 0, st>>>(c->raw_bits, c->infl_bits, c->H0, c->W0, c->pitch0, smax, row0, nrows);
        break;
    case 2:
        inflate_kernel<2, false><<<g, 256, 0, st>>>(c->raw_bits, c->infl_bits, c->H0, c->W0, c->pitch0, smax, row0, nrows);
        break;
    case 3:
    case 4:
        inflate_kernel<4, false><<<g, 256, 0, st>>>(c->raw_bits, c->infl_bits, c->H0, c->W0, c->pitch0, smax, row0, nrows);
        break;
    default:
        inflate_kernel<8, false><<<g, 256, 0, st>>>(c->raw_bits, c->infl_bits, c->H0, c->W0, c->pitch0, smax, row0, nrows);
        break;
    }
    MPROS_LAUNCH_CHECK(c);
    return MPROS_OK;
}

// stage 3: everything that needs the complete original-resolution layers: N2, the padded tile-word copies, N4 (if a goal is
// set), N5-N9
static int map_finish(Ctx *c, mpros_comm *cm = nullptr, int ds_rows_per_rank = 0)
{
    cudaStream_t st = c->stream;
    // N2
    const int ds = c->mp.downsampling_factor;
    if (ds == 1)
        MPROS_CUDA(cudaMemcpyAsync(c->ds_bits, c->raw_bits, (size_t)c->H0 * c->pitch0 * 4, cudaMemcpyDeviceToDevice, st));
    else
    {
        MPROS_CUDA(cudaMemsetAsync(c->ds_bits, 0, (size_t)c->H * c->pitch * 4, st));
        size_t words = (size_t)c->H * ((c->W + 31) / 32);
        const bool pow2 = ds <= 32 && (ds & (ds - 1)) == 0 && c->H0 % ds == 0 && c->W0 % ds == 0;
        if (pow2)
        {
            int log2ds = 0;
            while ((1 << log2ds) < ds)
                ++log2ds;
            downsample_pow2_kernel<<<grid_for(c, words), 256, 0, st>>>(c->raw_bits, c->pitch0, c->ds_bits, c->H, c->W, c->pitch, ds, log2ds);
        }
        else
            downsample_kernel<<<grid_for(c, words * 32), 256, 0, st>>>(c->raw_bits, c->H0, c->W0, c->pitch0, c->ds_bits, c->H, c->W,
                                                                  c->pitch, ds);
        MPROS_LAUNCH_CHECK(c);
    }
    // padded tile-word copies of the two collision layers
    {
#if MPROS_TW_BLOCK16
        const int tw_rows = ((c->H0 + 2 * kTwMargin + 15) / 16) * 8;
#else
        const int tw_rows = (c->H0 + 2 * kTwMargin + 3) / 4;
#endif
        const size_t tw_words = (size_t)c->tw_pitch * tw_rows;
        build_tilewords_kernel<<<grid_for(c, tw_words), 256, 0, st>>>(c->raw_bits, c->H0, c->W0, c->pitch0, c->raw_tw, c->tw_pitch, tw_rows);
        MPROS_LAUNCH_CHECK(c);
        build_tilewords_kernel<<<grid_for(c, tw_words), 256, 0, st>>>(c->infl_bits, c->H0, c->W0, c->pitch0, c->infl_tw, c->tw_pitch, tw_rows);
        MPROS_LAUNCH_CHECK(c);
    }
    c->unsafe_valid = false; // derived from the two layers just rebuilt
    // the reference sets get_map_ at the very end but the sub-steps below only need the geometry
    c->get_map = true;
    c->band_open = false;
    if (c->planner_configured)
    {
        const mpros_planner_params keep = c->pp;
        int rc = mpros_planner_config(static_cast<mpros_ctx *>(c), &keep); // n_check depends on the resolution
        if (rc)
        {
            c->planner_configured = false; // the old footprint constants do not fit the new resolution
            return rc;
        }
    }
    // nav_map.cpp:150-156
    if (c->get_goal)
    {
        int rc = map_build_goal(c);
        if (rc)
            return rc;
    }
    int rc = (cm && cm->world > 1) ? map_build_voronoi_banded(c, cm, ds_rows_per_rank) : map_build_voronoi(c);
    if (rc)
        return rc;
    MPROS_CUDA(cudaStreamSynchronize(st));
    return MPROS_OK;
}

// ---- pre-classification layer of the batched footprint test (collision.cu) ---------------------------------------------
// A pose is SURELY FREE when every cell any probe of footPrintInCollision4 can fall into is free and inside the map: the
// axis probes lie within max|footprint_longi_| of the pose, the four corners within the half diagonal. unsafe(i, j) = 1 unless
// all cells of the inflated layer within Chebyshev distance r_axis and all cells of the raw layer within r_corner of (i, j)
// are free, and (i, j) is at least r_corner cells inside the map. Square (Chebyshev) neighbourhoods contain the discs, the
// radii carry one extra cell for the truncation of the pose and of the probes: conservative, never wrong. Binary
// dilation of the bit-packed layers: a row pass (funnel shifts) and a column pass.
__global__ void __launch_bounds__(256) dilate_rows_kernel(const uint32_t *__restrict__ in, int H, int pitch, int R, uint32_t *__restrict__ out)
{
    const size_t total = (size_t)H * pitch;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x)
    {
        const int w = (int)(t % pitch);
        const uint32_t *row = in + (t - w);
        auto word = [&](int k) -> uint32_t { return (k >= 0 && k < pitch) ? row[k] : 0u; };
        uint32_t acc = 0;
        // shift by s cells toward higher columns: result word w = (in[w - q] << r) | (in[w - q - 1] >> (32 - r)), s = 32 q + r
        const int Q = (R + 31) >> 5;
        for (int q = -Q - 1; q <= Q; ++q)
        {
            const uint32_t hi = word(w - q), lo = word(w - q - 1);
            if ((hi | lo) == 0)
                continue;
            const int s0 = max(-R, 32 * q), s1 = min(R, 32 * q + 31);
            for (int sft = s0; sft <= s1; ++sft)
                acc |= __funnelshift_l(lo, hi, sft - 32 * q); // (hi << r) | (lo >> (32 - r)), r in 0..31
        }
        out[t] = acc;
    }
}

__global__ void __launch_bounds__(256) unsafe_cols_kernel(const uint32_t *__restrict__ rows_axis, int r_axis, const uint32_t *__restrict__ rows_corner,
                                                          int r_corner, int H, int W, int pitch, uint32_t *__restrict__ out)
{
    const size_t total = (size_t)H * pitch;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x)
    {
        const int i = (int)(t / pitch), w = (int)(t % pitch);
        uint32_t acc = 0;
        for (int d = max(0, i - r_axis); d <= min(H - 1, i + r_axis); ++d)
            acc |= rows_axis[(size_t)d * pitch + w];
        for (int d = max(0, i - r_corner); d <= min(H - 1, i + r_corner); ++d)
            acc |= rows_corner[(size_t)d * pitch + w];
        // at least r_corner cells inside the map on every side
        uint32_t interior = 0;
        if (i >= r_corner && i < H - r_corner)
        {
            const int j0 = w * 32;
            const int a = max(r_corner - j0, 0), b = min(W - r_corner - j0, 32); // bits [a, b) are interior columns
            if (b > a)
                interior = (b - a >= 32) ? 0xffffffffu : (((1u << (b - a)) - 1u) << a);
        }
        const int nbits = W - w * 32; // padding bits and words of the bit-packed layers stay zero
        const uint32_t valid = nbits >= 32 ? 0xffffffffu : (nbits > 0 ? ((1u << nbits) - 1u) : 0u);
        out[t] = (acc | ~interior) & valid;
    }
}

__global__ void __launch_bounds__(256) interleave_tw_kernel(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b, uint2 *__restrict__ out,
                                                            size_t n)
{
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
        out[k] = make_uint2(a[k], b[k]);
}

int map_build_unsafe(Ctx *c)
{
    if (c->unsafe_valid)
        return MPROS_OK;
    if (!c->get_map || !c->planner_configured)
    {
        set_error("map_build_unsafe: needs a map and a planner configuration");
        return MPROS_ERR_INVALID;
    }
    const PlannerView &v = c->pv;
    double reach_axis = std::fmax(std::fabs(v.longi_x0), std::fabs(v.longi_x1)), reach_corner = 0;
    for (int k = 0; k < 4; ++k)
        reach_corner = std::fmax(reach_corner, std::hypot(v.corner_x[k], v.corner_y[k]));
    // A probe at distance d cells from the pose lies at most floor(frac + d) <= floor(d) + 1 cells from the pose's cell in each
    // axis (frac < 1 is the pose's offset inside its cell); 1e-6 cells cover every rounding on either side.
    const double ra = std::floor(reach_axis / c->res0 + 1e-6) + 1, rc_ = std::floor(reach_corner / c->res0 + 1e-6) + 1;
    cudaStream_t st = c->stream;
    const size_t words0 = (size_t)c->H0 * c->pitch0;
#if MPROS_TW_BLOCK16
    const int tw_rows = ((c->H0 + 2 * kTwMargin + 15) / 16) * 8;
#else
    const int tw_rows = (c->H0 + 2 * kTwMargin + 3) / 4;
#endif
    const size_t tw_words = (size_t)c->tw_pitch * tw_rows;
    if (tw_words > c->cap_unsafe_tw)
    {
        if (c->unsafe_tw)
            MPROS_CUDA(cudaFree(c->unsafe_tw));
        if (c->cls_tw)
            MPROS_CUDA(cudaFree(c->cls_tw));
        c->unsafe_tw = nullptr;
        c->cls_tw = nullptr;
        c->cap_unsafe_tw = 0;
        MPROS_CUDA(cudaMalloc(&c->unsafe_tw, tw_words * 4));
        MPROS_CUDA(cudaMalloc(&c->cls_tw, tw_words * 8));
        c->cap_unsafe_tw = tw_words;
    }
    int rc = ensure_scratch(c, words0 * 4 * 3 + 1024);
    if (rc)
        return rc;
    uint32_t *rows_axis = static_cast<uint32_t *>(c->scratch), *rows_corner = rows_axis + words0, *unsafe_bits = rows_corner + words0;
    if (!(ra < 4096) || !(rc_ < 4096) || ra >= c->H0 || rc_ >= c->W0)
    {
        // degenerate footprint / tiny map: nothing is surely free
        MPROS_CUDA(cudaMemsetAsync(c->unsafe_tw, 0xff, tw_words * 4, st));
    }
    else
    {
        const int r_axis = (int)ra, r_corner = (int)rc_;
        dilate_rows_kernel<<<grid_for(c, words0), 256, 0, st>>>(c->infl_bits, c->H0, c->pitch0, r_axis, rows_axis);
        MPROS_LAUNCH_CHECK(c);
        dilate_rows_kernel<<<grid_for(c, words0), 256, 0, st>>>(c->raw_bits, c->H0, c->pitch0, r_corner, rows_corner);
        MPROS_LAUNCH_CHECK(c);
        unsafe_cols_kernel<<<grid_for(c, words0), 256, 0, st>>>(rows_axis, r_axis, rows_corner, r_corner, c->H0, c->W0, c->pitch0, unsafe_bits);
        MPROS_LAUNCH_CHECK(c);
        build_tilewords_kernel<<<grid_for(c, tw_words), 256, 0, st>>>(unsafe_bits, c->H0, c->W0, c->pitch0, c->unsafe_tw, c->tw_pitch, tw_rows);
        MPROS_LAUNCH_CHECK(c);
    }
    // classification layer of the batched footprint kernel: the "not surely free" word next to the inflated layer's word
    interleave_tw_kernel<<<grid_for(c, tw_words), 256, 0, st>>>(c->unsafe_tw, c->infl_tw, c->cls_tw, tw_words);
    MPROS_LAUNCH_CHECK(c);
    // the axis line has a probe exactly at the pose when n_check is even and the axis is symmetric about it
    // (footprint_longi_ = +-(L - W) / 2, hybrid_a_star.cpp:198-199, with center_to_rotate_center_longitudinal = 0)
    c->unsafe_center_probe = (v.n_check % 2 == 0 && v.longi_x0 == -v.longi_x1 && std::isfinite(v.fx_extent)) ? 1 : 0;
    c->unsafe_valid = true;
    return MPROS_OK;
}

static int set_occupancy_impl(Ctx *c, const int8_t *d_data, int width, int height, double resolution, double ox, double oy,
                              double oyaw)
{
    int have_old = 0, smax = -1;
    int rc = map_begin(c, width, height, resolution, ox, oy, oyaw, &have_old);
    if (rc)
        return rc;
    rc = map_threshold_rows(c, d_data, 0, height, have_old);
    if (rc)
        return rc;
    rc = map_inflate_prepare(c, &smax, nullptr);
    if (rc)
        return rc;
    rc = map_inflate_rows(c, smax, 0, height);
    if (rc)
        return rc;
    return map_finish(c);
}

} // namespace mpros

using namespace mpros;

extern "C" int mpros_map_set_occupancy_dev(mpros_ctx *c, const int8_t *d_data, int width, int height, double resolution,
                                           double ox, double oy, double oyaw)
{
    if (!c || !d_data || width <= 0 || height <= 0 || !(resolution > 0))
    {
        set_error("mpros_map_set_occupancy: invalid argument");
        return MPROS_ERR_INVALID;
    }
    MPROS_CUDA(cudaSetDevice(c->device));
    return set_occupancy_impl(c, d_data, width, height, resolution, ox, oy, oyaw);
}

extern "C" int mpros_map_set_occupancy(mpros_ctx *c, const int8_t *data, int width, int height, double resolution, double ox,
                                       double oy, double oyaw)
{
    if (!c || !data || width <= 0 || height <= 0 || !(resolution > 0))
    {
        set_error("mpros_map_set_occupancy: invalid argument");
        return MPROS_ERR_INVALID;
    }
    MPROS_CUDA(cudaSetDevice(c->device));
    size_t bytes = (size_t)width * height;
    int rc = ensure_stage(c, bytes);
    if (rc)
        return rc;
    MPROS_CUDA(cudaMemcpyAsync(c->stage, data, bytes, cudaMemcpyHostToDevice, c->stream));
    return set_occupancy_impl(c, static_cast<const int8_t *>(c->stage), width, height, resolution, ox, oy, oyaw);
}

// ---- row-band tiling of the original-resolution stages across GPUs (include/mpros_gpu.h: mpros_map_band_*) ----------
static int band_args_ok(mpros_ctx *c, const char *who, int row0, int nrows)
{
    if (!c)
    {
        set_error(std::string(who) + ": NULL context");
        return MPROS_ERR_INVALID;
    }
    if (!c->band_open)
    {
        set_error(std::string(who) + ": call mpros_map_band_begin first");
        return MPROS_ERR_INVALID;
    }
    if (row0 < 0 || nrows < 0 || row0 + nrows > c->H0)
    {
        set_error(std::string(who) + ": row range outside the map");
        return MPROS_ERR_INVALID;
    }
    return MPROS_OK;
}

extern "C" int mpros_map_band_begin(mpros_ctx *c, int width, int height, double resolution, double origin_x, double origin_y,
                                    double origin_yaw, int *halo_rows)
{
    if (!c || width <= 0 || height <= 0 || !(resolution > 0))
    {
        set_error("mpros_map_band_begin: invalid argument");
        return MPROS_ERR_INVALID;
    }
    MPROS_CUDA(cudaSetDevice(c->device));
    int have_old = 0;
    int rc = map_begin(c, width, height, resolution, origin_x, origin_y, origin_yaw, &have_old);
    if (rc)
        return rc;
    c->band_have_old = have_old;
    int r = 0;
    rc = map_inflate_prepare(c, &c->band_smax, &r);
    if (rc)
        return rc;
    if (halo_rows)
        *halo_rows = r;
    c->band_open = true;
    return MPROS_OK;
}

extern "C" int mpros_map_band_threshold(mpros_ctx *c, const int8_t *rows, int row0, int nrows)
{
    int rc = band_args_ok(c, "mpros_map_band_threshold", row0, nrows);
    if (rc)
        return rc;
    if (nrows == 0)
        return MPROS_OK;
    if (!rows)
    {
        set_error("mpros_map_band_threshold: NULL data");
        return MPROS_ERR_INVALID;
    }
    MPROS_CUDA(cudaSetDevice(c->device));
    size_t bytes = (size_t)nrows * c->W0;
    rc = ensure_stage(c, bytes);
    if (rc)
        return rc;
    MPROS_CUDA(cudaMemcpyAsync(c->stage, rows, bytes, cudaMemcpyHostToDevice, c->stream));
    return map_threshold_rows(c, static_cast<const int8_t *>(c->stage), row0, nrows, c->band_have_old);
}

extern "C" int mpros_map_band_threshold_dev(mpros_ctx *c, const int8_t *d_rows, int row0, int nrows)
{
    int rc = band_args_ok(c, "mpros_map_band_threshold_dev", row0, nrows);
    if (rc)
        return rc;
    if (nrows == 0)
        return MPROS_OK;
    if (!d_rows)
    {
        set_error("mpros_map_band_threshold_dev: NULL data");
        return MPROS_ERR_INVALID;
    }
    MPROS_CUDA(cudaSetDevice(c->device));
    return map_threshold_rows(c, d_rows, row0, nrows, c->band_have_old);
}

extern "C" int mpros_map_band_layer_ptr(mpros_ctx *c, int layer, void **d_ptr, size_t *row_pitch_bytes)
{
    if (!c || !d_ptr || !row_pitch_bytes || !c->band_open || (layer != MPROS_LAYER_STATIC_ORIG && layer != MPROS_LAYER_INFLATE_ORIG))
    {
        set_error("mpros_map_band_layer_ptr: needs an open band session and MPROS_LAYER_STATIC_ORIG / _INFLATE_ORIG");
        return MPROS_ERR_INVALID;
    }
    *d_ptr = layer == MPROS_LAYER_STATIC_ORIG ? (void *)c->raw_bits : (void *)c->infl_bits;
    *row_pitch_bytes = (size_t)c->pitch0 * 4;
    return MPROS_OK;
}

extern "C" int mpros_map_band_inflate(mpros_ctx *c, int row0, int nrows)
{
    int rc = band_args_ok(c, "mpros_map_band_inflate", row0, nrows);
    if (rc)
        return rc;
    MPROS_CUDA(cudaSetDevice(c->device));
    return map_inflate_rows(c, c->band_smax, row0, nrows);
}

extern "C" int mpros_map_band_finish(mpros_ctx *c)
{
    if (!c || !c->band_open)
    {
        set_error("mpros_map_band_finish: no open band session");
        return MPROS_ERR_INVALID;
    }
    MPROS_CUDA(cudaSetDevice(c->device));
    return map_finish(c);
}

// ---- every stage on row bands, exchanges inside (include/mpros_gpu.h: mpros_map_set_occupancy_banded) -----------------
// Bands are cut on the DOWN-SAMPLED grid (ceil(H / world) rows each, the last ranks may own fewer or none) so that the
// same partition serves N1/N3 (ds rows of the original grid per down-sampled row) and N5-N9.
static void band_partition(const Ctx *c, const mpros_comm *cm, int height, int *ds_rows_per_rank, int *row0, int *nrows)
{
    const int ds = c->mp.downsampling_factor;
    const int Hd = (int)std::ceil(static_cast<double>(height) / ds);
    const int world = cm ? cm->world : 1, rank = cm ? cm->rank : 0;
    const int per = (Hd + world - 1) / world;
    int lo, hi;
    band_of(Hd, per, rank, &lo, &hi);
    *ds_rows_per_rank = per;
    const long long a = (long long)lo * ds, b = (long long)hi * ds;
    *row0 = (int)std::min<long long>(a, height);
    *nrows = (int)std::min<long long>(b, height) - *row0;
}

extern "C" int mpros_map_band_rows(mpros_ctx *c, const mpros_comm *cm, int height, int *row0, int *nrows)
{
    if (!c || !row0 || !nrows || height <= 0)
    {
        set_error("mpros_map_band_rows: invalid argument");
        return MPROS_ERR_INVALID;
    }
    int per;
    band_partition(c, cm, height, &per, row0, nrows);
    return MPROS_OK;
}

static int set_occupancy_banded_impl(mpros_ctx *c, mpros_comm *cm, const int8_t *rows, bool on_device, int width, int height,
                                     double resolution, double origin_x, double origin_y, double origin_yaw)
{
    if (!c || !rows || width <= 0 || height <= 0 || !(resolution > 0))
    {
        set_error("mpros_map_set_occupancy_banded: invalid argument");
        return MPROS_ERR_INVALID;
    }
    MPROS_CUDA(cudaSetDevice(c->device));
    int have_old = 0;
    int rc = map_begin(c, width, height, resolution, origin_x, origin_y, origin_yaw, &have_old);
    if (rc)
        return rc;
    int smax = -1;
    rc = map_inflate_prepare(c, &smax, nullptr);
    if (rc)
        return rc;
    int per_ds, row0, nrows;
    band_partition(c, cm, height, &per_ds, &row0, &nrows);
    const int per0 = per_ds * c->mp.downsampling_factor;
    // N1 on the band: a rank uploads and reads only its own rows of the int8 grid
    if (nrows > 0)
    {
        const int8_t *d_rows = rows;
        if (!on_device)
        {
            const size_t bytes = (size_t)nrows * width;
            rc = ensure_stage(c, bytes);
            if (rc)
                return rc;
            MPROS_CUDA(cudaMemcpyAsync(c->stage, rows, bytes, cudaMemcpyHostToDevice, c->stream));
            d_rows = static_cast<const int8_t *>(c->stage);
        }
        rc = map_threshold_rows(c, d_rows, row0, nrows, have_old);
        if (rc)
            return rc;
    }
    // every rank needs the complete raw layer anyway (collision layer of the planner): gathering it BEFORE N3 also brings
    // the rows of the neighbouring bands the structuring element reaches into, so no separate halo exchange is needed
    rc = comm_allgather_rows(c, cm, c->raw_bits, (size_t)c->pitch0 * 4, c->H0, per0);
    if (rc)
        return rc;
    rc = map_inflate_rows(c, smax, row0, nrows); // N3 on the band
    if (rc)
        return rc;
    rc = comm_allgather_rows(c, cm, c->infl_bits, (size_t)c->pitch0 * 4, c->H0, per0);
    if (rc)
        return rc;
    return map_finish(c, cm, per_ds); // N2 + probe layouts (replicated, one pass over the bits), N4, N5-N9 banded
}

extern "C" int mpros_map_set_occupancy_banded(mpros_ctx *c, mpros_comm *cm, const int8_t *rows, int width, int height, double resolution,
                                              double origin_x, double origin_y, double origin_yaw)
{
    return set_occupancy_banded_impl(c, cm, rows, false, width, height, resolution, origin_x, origin_y, origin_yaw);
}
extern "C" int mpros_map_set_occupancy_banded_dev(mpros_ctx *c, mpros_comm *cm, const int8_t *d_rows, int width, int height,
                                                  double resolution, double origin_x, double origin_y, double origin_yaw)
{
    return set_occupancy_banded_impl(c, cm, d_rows, true, width, height, resolution, origin_x, origin_y, origin_yaw);
}

extern "C" int mpros_map_get_info(mpros_ctx *c, mpros_map_info *o)
{
    if (!c || !o)
    {
        set_error("mpros_map_get_info: NULL argument");
        return MPROS_ERR_INVALID;
    }
    std::memset(o, 0, sizeof(*o));
    o->get_map = c->get_map;
    o->get_goal = c->get_goal;
    o->map_height = c->H;
    o->map_width = c->W;
    o->map_height_orig = c->H0;
    o->map_width_orig = c->W0;
    o->map_resolution = c->res;
    o->map_resolution_orig = c->res0;
    o->map_origin_x = c->ox;
    o->map_origin_y = c->oy;
    o->map_origin_yaw = c->oyaw;
    o->map_origin_sin = c->osin;
    o->map_origin_cos = c->ocos;
    // nav_map.cpp:74-77
    o->x_max = c->ox + c->res0 * c->W0;
    o->x_min = c->ox;
    o->y_max = c->oy + c->res0 * c->H0;
    o->y_min = c->oy;
    o->num_regions = c->num_regions;
    return MPROS_OK;
}

extern "C" int mpros_map_download(mpros_ctx *c, int layer, void *dst, size_t dst_bytes)
{
    if (!c || !dst)
    {
        set_error("mpros_map_download: NULL argument");
        return MPROS_ERR_INVALID;
    }
    if (!c->get_map)
    {
        set_error("mpros_map_download: no map set");
        return MPROS_ERR_INVALID;
    }
    MPROS_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const size_t cells0 = (size_t)c->H0 * c->W0, cells = (size_t)c->H * c->W;
    size_t need = 0;
    const void *src = nullptr;
    switch (layer)
    {
    case MPROS_LAYER_STATIC_ORIG:
    case MPROS_LAYER_INFLATE_ORIG:
    case MPROS_LAYER_STATIC:
    case MPROS_LAYER_ISEDGE:
    {
        bool orig = (layer == MPROS_LAYER_STATIC_ORIG || layer == MPROS_LAYER_INFLATE_ORIG);
        need = orig ? cells0 : cells;
        if (dst_bytes < need)
            break;
        int rc = ensure_scratch(c, need);
        if (rc)
            return rc;
        const uint32_t *bits = layer == MPROS_LAYER_STATIC_ORIG ? c->raw_bits
                               : layer == MPROS_LAYER_INFLATE_ORIG ? c->infl_bits
                               : layer == MPROS_LAYER_STATIC ? c->ds_bits
                                                             : c->edge_bits;
        unpack_bits_kernel<<<grid_for(c, need), 256, 0, st>>>(bits, orig ? c->H0 : c->H, orig ? c->W0 : c->W,
                                                          orig ? c->pitch0 : c->pitch, static_cast<int8_t *>(c->scratch));
        MPROS_LAUNCH_CHECK(c);
        src = c->scratch;
        break;
    }
    case MPROS_LAYER_GOAL:
    case MPROS_LAYER_OBS_DIST:
    case MPROS_LAYER_EDGE_DIST:
    {
        need = cells * 4;
        if (dst_bytes < need)
            break;
        if (layer == MPROS_LAYER_GOAL && !c->get_goal)
        {
            set_error("mpros_map_download: goal map not generated (no goal set or goal outside the map)");
            return MPROS_ERR_INVALID;
        }
        int rc = ensure_scratch(c, need);
        if (rc)
            return rc;
        const uint16_t *in = layer == MPROS_LAYER_GOAL ? c->goal : layer == MPROS_LAYER_OBS_DIST ? c->obs_dist : c->edge_dist;
        widen_u16_kernel<<<grid_for(c, cells), 256, 0, st>>>(in, cells, static_cast<int32_t *>(c->scratch));
        MPROS_LAUNCH_CHECK(c);
        src = c->scratch;
        break;
    }
    case MPROS_LAYER_REGION:
        need = cells * 4;
        src = c->region;
        break;
    case MPROS_LAYER_VORONOI:
        need = cells * 4;
        src = c->voronoi;
        break;
    default:
        set_error("mpros_map_download: unknown layer");
        return MPROS_ERR_INVALID;
    }
    if (dst_bytes < need)
    {
        set_error("mpros_map_download: destination buffer too small");
        return MPROS_ERR_CAPACITY;
    }
    MPROS_CUDA(cudaMemcpyAsync(dst, src, need, cudaMemcpyDeviceToHost, st));
    MPROS_CUDA(cudaStreamSynchronize(st));
    return MPROS_OK;
}

extern "C" int mpros_map_layer_msg(mpros_ctx *c, int layer, int8_t *dst, size_t dst_bytes, size_t *n_out)
{
    if (!c || !n_out)
    {
        set_error("mpros_map_layer_msg: NULL argument");
        return MPROS_ERR_INVALID;
    }
    *n_out = 0;
    const bool bit_layer = layer == MPROS_LAYER_STATIC || layer == MPROS_LAYER_STATIC_ORIG || layer == MPROS_LAYER_INFLATE_ORIG;
    if (!bit_layer && layer != MPROS_LAYER_GOAL && layer != MPROS_LAYER_VORONOI)
    {
        set_error("mpros_map_layer_msg: the reference publishes only the static, goal, Voronoi, static_orig and inflate_orig layers");
        return MPROS_ERR_INVALID;
    }
    // vecToMsg returns an empty message while there is no map / no such layer yet (nav_map.cpp:659-662)
    if (!c->get_map || (layer == MPROS_LAYER_GOAL && !c->get_goal))
        return MPROS_OK;
    const bool orig = layer == MPROS_LAYER_STATIC_ORIG || layer == MPROS_LAYER_INFLATE_ORIG;
    const size_t n = orig ? (size_t)c->H0 * c->W0 : (size_t)c->H * c->W;
    *n_out = n;
    if (!dst)
        return MPROS_OK; // size query
    if (dst_bytes < n)
    {
        set_error("mpros_map_layer_msg: destination buffer too small");
        return MPROS_ERR_CAPACITY;
    }
    MPROS_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    int rc = ensure_scratch(c, n + 256);
    if (rc)
        return rc;
    int8_t *out = static_cast<int8_t *>(c->scratch);
    int *maxv = reinterpret_cast<int *>(static_cast<char *>(c->scratch) + ((n + 255) / 256) * 256);
    if (bit_layer)
    {
        const uint32_t *bits = layer == MPROS_LAYER_STATIC_ORIG ? c->raw_bits : layer == MPROS_LAYER_INFLATE_ORIG ? c->infl_bits : c->ds_bits;
        msg_bits_kernel<<<grid_for(c, n), 256, 0, st>>>(bits, orig ? c->H0 : c->H, orig ? c->W0 : c->W, orig ? c->pitch0 : c->pitch, out);
        MPROS_LAUNCH_CHECK(c);
    }
    else
    {
        MPROS_CUDA(cudaMemsetAsync(maxv, 0, 4, st));
        if (layer == MPROS_LAYER_GOAL)
        {
            msg_max_u16_kernel<<<grid_for(c, n), 256, 0, st>>>(c->goal, n, maxv);
            MPROS_LAUNCH_CHECK(c);
            msg_scale_u16_kernel<<<grid_for(c, n), 256, 0, st>>>(c->goal, n, maxv, out);
        }
        else
        {
            msg_max_f32_kernel<<<grid_for(c, n), 256, 0, st>>>(c->voronoi, n, maxv);
            MPROS_LAUNCH_CHECK(c);
            msg_scale_f32_kernel<<<grid_for(c, n), 256, 0, st>>>(c->voronoi, n, maxv, out);
        }
        MPROS_LAUNCH_CHECK(c);
    }
    MPROS_CUDA(cudaMemcpyAsync(dst, out, n, cudaMemcpyDeviceToHost, st));
    MPROS_CUDA(cudaStreamSynchronize(st));
    return MPROS_OK;
}

extern "C" int mpros_map_upload(mpros_ctx *c, int layer, const void *src, size_t src_bytes)
{
    if (!c || !src)
    {
        set_error("mpros_map_upload: NULL argument");
        return MPROS_ERR_INVALID;
    }
    if (!c->get_map || layer != MPROS_LAYER_VORONOI)
    {
        set_error("mpros_map_upload: needs a map; only MPROS_LAYER_VORONOI can be uploaded");
        return MPROS_ERR_INVALID;
    }
    size_t need = (size_t)c->H * c->W * sizeof(float);
    if (src_bytes != need)
    {
        set_error("mpros_map_upload: size mismatch");
        return MPROS_ERR_CAPACITY;
    }
    MPROS_CUDA(cudaSetDevice(c->device));
    MPROS_CUDA(cudaMemcpyAsync(c->voronoi, src, need, cudaMemcpyHostToDevice, c->stream));
    MPROS_CUDA(cudaStreamSynchronize(c->stream));
    return MPROS_OK;
}

extern "C" int mpros_map_point_in_collision(mpros_ctx *c, int layer, const int32_t *ij, size_t n, uint8_t *out)
{
    if (!c || (!ij && n) || (!out && n))
    {
        set_error("mpros_map_point_in_collision: NULL argument");
        return MPROS_ERR_INVALID;
    }
    if (!c->get_map)
    {
        set_error("mpros_map_point_in_collision: no map set");
        return MPROS_ERR_INVALID;
    }
    if (n == 0)
        return MPROS_OK;
    const uint32_t *bits;
    int H, W, pitch;
    if (layer == MPROS_LAYER_STATIC)
        bits = c->ds_bits, H = c->H, W = c->W, pitch = c->pitch;
    else if (layer == MPROS_LAYER_STATIC_ORIG)
        bits = c->raw_bits, H = c->H0, W = c->W0, pitch = c->pitch0;
    else if (layer == MPROS_LAYER_INFLATE_ORIG)
        bits = c->infl_bits, H = c->H0, W = c->W0, pitch = c->pitch0;
    else
    {
        set_error("mpros_map_point_in_collision: layer must be STATIC, STATIC_ORIG or INFLATE_ORIG");
        return MPROS_ERR_INVALID;
    }
    MPROS_CUDA(cudaSetDevice(c->device));
    size_t in_bytes = n * 8, total = in_bytes + n;
    int rc = ensure_stage(c, total);
    if (rc)
        return rc;
    int32_t *d_ij = static_cast<int32_t *>(c->stage);
    uint8_t *d_out = reinterpret_cast<uint8_t *>(c->stage) + in_bytes;
    MPROS_CUDA(cudaMemcpyAsync(d_ij, ij, in_bytes, cudaMemcpyHostToDevice, c->stream));
    point_lookup_kernel<<<grid_for(c, n), 256, 0, c->stream>>>(bits, H, W, pitch, d_ij, n, d_out);
    MPROS_LAUNCH_CHECK(c);
    MPROS_CUDA(cudaMemcpyAsync(out, d_out, n, cudaMemcpyDeviceToHost, c->stream));
    MPROS_CUDA(cudaStreamSynchronize(c->stream));
    return MPROS_OK;
}
