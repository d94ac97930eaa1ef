// K2, tile-centric formulation for sm_100a: right-matrix row segments staged through TMA into shared memory.
//
// Replaces the block loop of StringGrouper._build_matches
// (/root/reference/string_grouper/string_grouper.py:734-750: `Bs` = row blocks of the right matrix, one
// sp_matmul_topn per (left block, right block) pair, :737-743) for L2-normalised non-negative matrices
// (the K1 output).  Where the reference slices ~4000-row right blocks so that the accumulators of
// sparse_dot_topn stay cache-resident (:387-389), this kernel takes 256-row right blocks ("column tiles"
// of the product) whose whole inverted index fits in shared memory:
//
//   tiles_build      right matrix -> per tile one contiguous blob: the tile's postings sorted by (feature,
//                    column) as 4-byte {fixed-point weight, column}, a bitmap over the features present, its
//                    rank table and the bucket offsets; plus the fp16 block maxima the filter streams
//   pack_left        pruned left rows (sg_prune_rows) -> {feature, fixed-point weight} pairs and one 16-byte
//                    record {start, kept features, threshold, pruned norm} per row in processing order
//   tile_filter      block-max test of every (left row, tile) pair: no column of tile t can collect more than
//                    ub = sum_f |a_f| max|w_(f,t)|; pairs that cannot reach their candidate threshold are
//                    dropped.  Output: one bit per pair, transposed ([tile word][left rank]) so that a tile
//                    reads the ranks that survive for it with coalesced loads
//   tile_candidates  one CTA per (tile, segment of left ranks): the tile's blob is copied into shared memory
//                    by cp.async.bulk (TMA, completion on an mbarrier); every warp then takes surviving left
//                    rows, finds the buckets of the row's kept features through the bitmap (two shared-memory
//                    loads), walks their concatenation 32 postings per step and adds integer products
//                    a_q * w_q (2^-30 units) into a 256-column accumulator tile with native shared-memory
//                    atomics.  All weights are non-negative, so a column's partial score only grows: the step
//                    in which it crosses the candidate threshold reports the pair — no sweep of the tile.
//
// Every candidate is re-scored exactly (sg_rescore), so the result is identical to the plain traversal.
#include <cub/cub.cuh>
#include <cuda_fp16.h>

#include "sg_common.cuh"

namespace sg {

struct __align__(16) TileDesc {
    long long blob_off;     // byte offset of the tile's blob (16-byte aligned)
    int n_post;             // postings of the tile
    int n_dist;             // distinct features of the tile = buckets
};

constexpr int TL_W = 256;            // columns per tile (accumulator: 256 x u32 = 1 KB per warp)
constexpr int TL_CBUF = 96;          // candidate buffer entries per warp
constexpr float TL_FIX = 32768.f;    // weights in 2^-15 units, products in 2^-30 units
constexpr int TL_HEAD_BYTES = 128;   // mbarrier, item broadcast, survivor count

__host__ __device__ __forceinline__ int a16(int x) { return (x + 15) & ~15; }
// bitmap words per tile: one bit per feature, rounded up to 16 bytes
__host__ __device__ __forceinline__ int bitmap_words(int64_t n_cols) { return (int)(((n_cols + 31) / 32 + 3) & ~(int64_t)3); }
__host__ __device__ __forceinline__ int blob_bytes(int n_post, int n_dist, int bw) {
    return a16(4 * n_post) + 4 * bw + a16(2 * bw) + a16(2 * (n_dist + 1));
}

// ---------------------------------------------------------------------------
// mbarrier / bulk-copy (TMA) primitives
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "SG_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra SG_DONE_%=;\n"
        "bra SG_WAIT_%=;\n"
        "SG_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ---------------------------------------------------------------------------
// build: right matrix -> tile blobs
// ---------------------------------------------------------------------------
// key = ((tile * V + feature) << 16) | column inside the tile; value = weight bits
__global__ void tiles_keys_kernel(int64_t n_rows, const int64_t *__restrict__ indptr,
                                  const int32_t *__restrict__ indices, const float *__restrict__ val,
                                  const int32_t *__restrict__ rank, int W, int64_t V, int64_t base, float w_scale,
                                  uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int64_t pos = rank ? rank[row] : row;
    const int64_t t = pos / W;
    const uint64_t local = (uint64_t)(pos - t * W);
    const int64_t p1 = indptr[row + 1];
    for (int64_t p = indptr[row] + lane_id(); p < p1; p += 32) {
        keys[p - base] = ((uint64_t)(t * V + indices[p]) << 16) | local;
        vals[p - base] = __float_as_uint(val[p] * w_scale);
    }
}

__global__ void tiles_ptr_kernel(int64_t T, int64_t V, int64_t nnz, const uint64_t *__restrict__ keys,
                                 int32_t *__restrict__ tile_ptr) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > T) return;
    const uint64_t k = ((uint64_t)(t * V)) << 16;
    int64_t lo = 0, hi = nnz;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (keys[mid] < k) lo = mid + 1; else hi = mid;
    }
    tile_ptr[t] = (int32_t)lo;
}

// one CTA per tile: distinct features -> blob size; also the largest blob / posting count over all tiles
__global__ void tiles_count_kernel(int64_t T, const int32_t *__restrict__ tile_ptr, const uint64_t *__restrict__ keys,
                                   int bw, int32_t *__restrict__ n_dist, int64_t *__restrict__ bytes,
                                   int32_t *__restrict__ maxima) {
    const int64_t t = blockIdx.x;
    const int b = tile_ptr[t], e = tile_ptr[t + 1];
    int c = 0;
    for (int p = b + threadIdx.x; p < e; p += blockDim.x)
        c += (p == b || (keys[p] >> 16) != (keys[p - 1] >> 16)) ? 1 : 0;
    typedef cub::BlockReduce<int, 256> Red;
    __shared__ typename Red::TempStorage tmp;
    const int total = Red(tmp).Sum(c);
    if (threadIdx.x == 0) {
        n_dist[t] = total;
        const int bb = blob_bytes(e - b, total, bw);
        bytes[t] = bb;
        atomicMax(maxima, bb);
        atomicMax(maxima + 1, e - b);
    }
}

// one CTA (256 threads) per tile
__global__ void __launch_bounds__(256)
tiles_fill_kernel(int64_t T, int64_t V, int64_t Tp, const int32_t *__restrict__ tile_ptr,
                  const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals,
                  const int32_t *__restrict__ n_dist, const int64_t *__restrict__ blob_off, int bw,
                  unsigned char *__restrict__ blob, TileDesc *__restrict__ desc,
                  unsigned short *__restrict__ maxw_rows) {
    extern __shared__ uint32_t s_bitmap[];          // bw words
    typedef cub::BlockScan<int, 256> Scan;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ int s_running;
    const int64_t t = blockIdx.x;
    const int b = tile_ptr[t], e = tile_ptr[t + 1];
    const int n_post = e - b, nd = n_dist[t];
    const long long boff = blob_off[t];
    uint32_t *post = reinterpret_cast<uint32_t *>(blob + boff);
    uint32_t *bm_out = reinterpret_cast<uint32_t *>(blob + boff + a16(4 * n_post));
    unsigned short *prefix = reinterpret_cast<unsigned short *>(blob + boff + a16(4 * n_post) + 4 * bw);
    unsigned short *off = reinterpret_cast<unsigned short *>(blob + boff + a16(4 * n_post) + 4 * bw + a16(2 * bw));
    for (int i = threadIdx.x; i < bw; i += 256) s_bitmap[i] = 0u;
    if (threadIdx.x == 0) {
        s_running = 0;
        desc[t].blob_off = boff;
        desc[t].n_post = n_post;
        desc[t].n_dist = nd;
    }
    __syncthreads();
    const uint64_t fbase = (uint64_t)(t * V);
    for (int base = 0; base < n_post; base += 256) {
        const int p = b + base + threadIdx.x;
        int head = 0;
        uint64_t kf = 0;
        if (p < e) {
            const uint64_t key = keys[p];
            kf = key >> 16;
            head = (p == b || (keys[p - 1] >> 16) != kf) ? 1 : 0;
            float w = __uint_as_float(vals[p]);
            w = w < 0.f ? 0.f : w;
            unsigned wq = (unsigned)__float2int_rn(w * TL_FIX);
            wq = wq > 65535u ? 65535u : wq;
            post[p - b] = (wq << 16) | ((unsigned)(key & 0xffffu) << 2);
        }
        int excl = 0;
        Scan(tmp).ExclusiveSum(head, excl);
        const int running = s_running;
        if (head) {
            const unsigned f = (unsigned)(kf - fbase);
            atomicOr(&s_bitmap[f >> 5], 1u << (f & 31));
            off[running + excl] = (unsigned short)(p - b);
            // largest fixed-point weight of the bucket (the run of equal features starting here)
            unsigned mq = 0;
            for (int q = p; q < e && (keys[q] >> 16) == kf; ++q) {
                float w = __uint_as_float(vals[q]);
                w = w < 0.f ? 0.f : w;
                unsigned wq = (unsigned)__float2int_rn(w * TL_FIX);
                mq = wq > mq ? wq : mq;
            }
            mq = mq > 65535u ? 65535u : mq;
            // what the kernel accumulates is w_q / 32768: the fp16 bound must not fall short of it
            maxw_rows[(int64_t)f * Tp + t] = __half_as_ushort(__float2half_ru((float)mq * (1.f / TL_FIX)));
        }
        __syncthreads();
        if (threadIdx.x == 255) s_running = running + excl + head;
        __syncthreads();
    }
    if (threadIdx.x == 0) off[nd] = (unsigned short)n_post;
    // zero padding of the posting section (never walked; keeps the blob deterministic)
    for (int i = n_post + threadIdx.x; i < (a16(4 * n_post) >> 2); i += 256) post[i] = 0u;
    __syncthreads();
    // rank table: prefix[w] = set bits before word w
    if (threadIdx.x == 0) s_running = 0;
    __syncthreads();
    for (int base = 0; base < bw; base += 256) {
        const int i = base + threadIdx.x;
        const uint32_t word = i < bw ? s_bitmap[i] : 0u;
        int excl = 0;
        Scan(tmp).ExclusiveSum(__popc(word), excl);
        const int running = s_running;
        if (i < bw) {
            bm_out[i] = word;
            prefix[i] = (unsigned short)(running + excl);
        }
        __syncthreads();
        if (threadIdx.x == 255) s_running = running + excl + __popc(word);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// left operand: pruned rows -> packed {feature, fixed-point weight}, 16-byte row records in processing order
// ---------------------------------------------------------------------------
__global__ void pack_left_kernel(int64_t n_ranks, const int32_t *__restrict__ perm, int64_t row_begin,
                                 const int64_t *__restrict__ indptr, const int32_t *__restrict__ p_len,
                                 const int32_t *__restrict__ p_idx, const float *__restrict__ p_val,
                                 const float *__restrict__ p_thr, const float *__restrict__ p_xp, float a_scale,
                                 int2 *__restrict__ lpack, int4 *__restrict__ rowinfo) {
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n_ranks) return;
    const int64_t row = perm ? perm[r] : row_begin + r;
    const int64_t p0 = indptr[row];
    const int nf = p_len ? p_len[row] : (int)(indptr[row + 1] - p0);
    for (int k = lane_id(); k < nf; k += 32) {
        float a = p_val[p0 + k] * a_scale;
        a = a < 0.f ? 0.f : a;
        lpack[p0 + k] = make_int2(p_idx[p0 + k], __float2int_rn(a * TL_FIX));
    }
    if (lane_id() == 0)
        rowinfo[r] = make_int4((int)p0, nf, __float_as_int(p_thr[row]), __float_as_int(p_xp ? p_xp[row] : 0.f));
}

// ---------------------------------------------------------------------------
// block-max filter: one bit per (left rank, tile), transposed
// ---------------------------------------------------------------------------
// Word layout: the 64 tiles of batch b = [64b, 64b+64) occupy words 2b (even tiles) and 2b+1 (odd tiles), bit
// (tile & 63) >> 1 — the bounds of two neighbouring tiles are evaluated in one packed fp16 lane.
constexpr int FL_WARPS = 8;
constexpr int FL_RANKS = 32;      // ranks per CTA (4 per warp)
constexpr int FL_WORDS = 128;     // mask words per pass through shared memory (4096 tiles)

__global__ void __launch_bounds__(FL_WARPS * 32)
tile_filter_kernel(int64_t n_ranks, const int4 *__restrict__ rowinfo, const int2 *__restrict__ lpack,
                   const uint32_t *__restrict__ maxw_h, int Tp, int64_t T, const float *__restrict__ tile_bound,
                   uint32_t *__restrict__ mask, int64_t mask_stride) {
    __shared__ uint32_t buf[FL_WORDS][FL_RANKS + 1];
    constexpr int RPW = FL_RANKS / FL_WARPS;          // ranks per warp
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t rank0 = (int64_t)blockIdx.x * FL_RANKS;
    const int n_words = Tp >> 5;
    const int half_tp = Tp >> 1;
    // the warp's ranks: kept features in registers (one per lane)
    int nf[RPW], f0[RPW];
    float thr_r[RPW], xp[RPW], slack[RPW];
    __half2 a2[RPW];
#pragma unroll
    for (int ri = 0; ri < RPW; ++ri) {
        const int64_t r = rank0 + warp * RPW + ri;
        nf[ri] = 0; f0[ri] = 0; thr_r[ri] = 0.f; xp[ri] = 0.f;
        a2[ri] = __float2half2_rn(0.f);
        if (r < n_ranks) {
            const int4 info = rowinfo[r];
            nf[ri] = info.y;
            thr_r[ri] = __int_as_float(info.z);
            xp[ri] = __int_as_float(info.w);
            if (lane < info.y) {
                const int2 fa = lpack[(int64_t)info.x + lane];
                f0[ri] = fa.x;
                // rounded up: the bound must not fall short
                a2[ri] = __half2half2(__float2half_ru((float)fa.y * (1.f / TL_FIX)));
            }
        }
        const int nk = nf[ri] < 32 ? nf[ri] : 32;
        slack[ri] = 5e-4f * (float)nk + 1e-4f;        // fp16 arithmetic of the bound
    }
    for (int w0 = 0; w0 < n_words; w0 += FL_WORDS) {
        const int w1 = w0 + FL_WORDS < n_words ? w0 + FL_WORDS : n_words;
        // batches outermost: the 32 neighbouring ranks of the CTA (similar rows: mostly the same features) read the same
        // block-maxima lines within a short time, so most of these loads hit L1 instead of L2
        for (int wd = w0; wd < w1; wd += 2) {
            const int tb = wd << 5;                       // first tile of the batch
            const int t0 = tb + 2 * lane;
            const uint32_t *mrow = maxw_h + (tb >> 1) + lane;
            const float2 tb2 = reinterpret_cast<const float2 *>(tile_bound)[(tb >> 1) + lane];
#pragma unroll
            for (int ri = 0; ri < RPW; ++ri) {
                unsigned m_even = 0, m_odd = 0;
                if (nf[ri] > 32) {                        // more kept features than lanes: every tile is walked
                    m_even = __ballot_sync(FULL, t0 < T);
                    m_odd = __ballot_sync(FULL, t0 + 1 < T);
                } else if (nf[ri] > 0) {
                    const int nk = nf[ri];
                    __half2 ub2 = __float2half2_rn(0.f);
                    for (int k0 = 0; k0 < nk; k0 += 8) {      // eight loads in flight per lane
                        uint32_t m[8];
                        __half2 ak[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int kk = k0 + j;
                            const int fk = __shfl_sync(FULL, f0[ri], kk & 31);
                            ak[j] = __shfl_sync(FULL, a2[ri], kk & 31);
                            m[j] = kk < nk ? mrow[(int64_t)fk * half_tp] : 0u;
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) ub2 = __hfma2(ak[j], *reinterpret_cast<const __half2 *>(&m[j]), ub2);
                    }
                    const float2 ub = __half22float2(ub2);
                    const float thr0 = xp[ri] > 0.f ? fmaxf(fmaf(-xp[ri], tb2.x, thr_r[ri]), 0.f) : thr_r[ri];
                    const float thr1 = xp[ri] > 0.f ? fmaxf(fmaf(-xp[ri], tb2.y, thr_r[ri]), 0.f) : thr_r[ri];
                    m_even = __ballot_sync(FULL, t0 < T && ub.x + slack[ri] > thr0);
                    m_odd = __ballot_sync(FULL, t0 + 1 < T && ub.y + slack[ri] > thr1);
                }
                if (lane == 0) {
                    buf[wd - w0][warp * RPW + ri] = m_even;
                    buf[wd - w0 + 1][warp * RPW + ri] = m_odd;
                }
            }
        }
        __syncthreads();
        for (int wd = w0 + warp; wd < w1; wd += FL_WARPS)
            if (rank0 + lane < mask_stride) mask[(int64_t)wd * mask_stride + rank0 + lane] = buf[wd - w0][lane];
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// candidates
// ---------------------------------------------------------------------------
// shared-memory accesses by 32-bit shared address (no generic-pointer arithmetic in the hot loop)
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t a) {
    uint32_t v;
    asm volatile("{ .reg .u16 h; ld.shared.u16 h, [%1]; cvt.u32.u16 %0, h; }" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts64(uint32_t a, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void sts128z(uint32_t a) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(a), "r"(0u) : "memory");
}
__device__ __forceinline__ void reds_or(uint32_t a, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
// old = (pred ? atomicAdd(shared a, v) : 0), without a branch
__device__ __forceinline__ uint32_t atoms_add_if(uint32_t a, uint32_t v, bool pred) {
    uint32_t old;
    asm volatile(
        "{ .reg .pred p; setp.ne.u32 p, %3, 0; mov.u32 %0, 0; @p atom.shared.add.u32 %0, [%1], %2; }"
        : "=r"(old)
        : "r"(a), "r"(v), "r"((unsigned)pred)
        : "memory");
    return old;
}
__device__ __forceinline__ uint32_t lds32_if(uint32_t a, bool pred) {
    uint32_t v;
    asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; mov.u32 %0, 0; @p ld.shared.u32 %0, [%1]; }"
                 : "=r"(v)
                 : "r"(a), "r"((unsigned)pred));
    return v;
}

__device__ __forceinline__ void sts16(uint32_t a, uint32_t v) {
    asm volatile("{ .reg .u16 h; cvt.u16.u32 h, %1; st.shared.u16 [%0], h; }" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// the shared-window base as a value the compiler cannot re-derive (it would recompute it from the CTA id per use)
__device__ __forceinline__ uint32_t opaque(uint32_t x) {
    asm volatile("mov.u32 %0, %0;" : "+r"(x));
    return x;
}

// Thread-per-pair layout.  A (left row, tile) pair is tiny — about 13 kept features, 70 postings, 2 columns over the
// threshold — so warp-collective processing (prefix sums, owner search, votes per 32 postings) costs more than the work.
// Here every LANE owns one pair: its 256 partial scores are 16-bit fixed point (2^-15 units) in a lane-private column
// of the warp's accumulator block (acc[col][lane]: conflict-free, no atomics), its features arrive by cp.async into a
// lane-private row of a staging block, its buckets are found and walked serially.  32 pairs advance per warp step.
constexpr int TP_FS = 16;                           // features staged per pair and round (rows with more: more rounds)
constexpr int TP_FSTRIDE = TP_FS + 1;               // 8-byte entries per pair row, padded: lane = pair access is conflict-free
constexpr int TP_ACC_BYTES = TL_W * 32 * 2;         // 16 KB: 256 columns x 32 pairs x u16
constexpr int TL_WARP_BYTES2 = TP_ACC_BYTES + 32 * TP_FSTRIDE * 8 + TL_CBUF * 8;
constexpr int tl_min_ctas(int) { return 1; }
__host__ __device__ constexpr int tl_list(int nw) { return nw * 256 + nw * 32; }     // survivor ranks held per round
constexpr float TP_FIX = 32768.f;

template <int NW>
__global__ void __launch_bounds__(NW * 32, tl_min_ctas(NW))
tile_candidates_kernel(const int32_t *__restrict__ perm_a, int64_t n_ranks, int64_t row_begin,
                       const int4 *__restrict__ rowinfo, const int2 *__restrict__ lpack,
                       const uint32_t *__restrict__ mask, int64_t mask_stride, const TileDesc *__restrict__ tdesc,
                       const unsigned char *__restrict__ blob, int64_t T, int bw,
                       const float *__restrict__ tile_bound, const int32_t *__restrict__ perm_b,
                       int64_t seg_ranks, int64_t n_seg, int32_t *__restrict__ cand_row,
                       int32_t *__restrict__ cand_col, unsigned long long cap,
                       unsigned long long *__restrict__ cand_count, unsigned long long *__restrict__ queue,
                       unsigned long long *__restrict__ walk_stats, int stage_bytes) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem);
    volatile long long *s_item = reinterpret_cast<volatile long long *>(smem + 16);
    int *s_count = reinterpret_cast<int *>(smem + 32);
    uint32_t *s_list = reinterpret_cast<uint32_t *>(smem + TL_HEAD_BYTES);
    unsigned char *stage = smem + TL_HEAD_BYTES + tl_list(NW) * 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint32_t smem_s = opaque(smem_u32(smem));
    const uint32_t stage_s = smem_s + TL_HEAD_BYTES + tl_list(NW) * 4;
    const uint32_t warp_s = stage_s + stage_bytes + warp * TL_WARP_BYTES2;
    const uint32_t acc_s = warp_s + lane * 2;                               // this lane's column of partial scores
    const uint32_t feat_s = warp_s + TP_ACC_BYTES + lane * (TP_FSTRIDE * 8);  // this lane's feature / bucket row
    const uint32_t cbuf_s = warp_s + TP_ACC_BYTES + 32 * TP_FSTRIDE * 8;
    int ccount = 0;
    for (int c = lane; c < TP_ACC_BYTES / 16; c += 32) sts128z(warp_s + c * 16);
    if (threadIdx.x == 0) {
        mbar_init(mbar, 1);
        *s_count = 0;
    }
    __syncthreads();

    const unsigned long long n_items = (unsigned long long)T * (unsigned long long)n_seg;
    unsigned parity = 0;
    unsigned long long n_pairs = 0, n_walked = 0;       // (row, tile) pairs taken / postings added by this lane

    // buffered candidates -> global list (one atomic per flush)
    auto flush = [&]() {
        __syncwarp();
        if (ccount > 0) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(cand_count, (unsigned long long)ccount);
            base = __shfl_sync(FULL, base, 0);
            for (int i = lane; i < ccount; i += 32) {
                const uint2 c = lds64(cbuf_s + i * 8);
                if (base + i < cap) {
                    cand_row[base + i] = (int32_t)(perm_a ? perm_a[c.x] : row_begin + c.x);
                    cand_col[base + i] = perm_b ? perm_b[c.y] : (int32_t)c.y;
                }
            }
            ccount = 0;
        }
        __syncwarp();
    };

    for (;;) {
        // ---- next (tile, rank segment); its blob goes into shared memory by bulk copies (TMA)
        if (threadIdx.x == 0) {
            const unsigned long long it = atomicAdd(queue, 1ull);
            *s_item = (long long)it;
            if (it < n_items) {
                const TileDesc d = tdesc[it / (unsigned long long)n_seg];
                const unsigned bytes = (unsigned)blob_bytes(d.n_post, d.n_dist, bw);
                mbar_expect_tx(mbar, bytes);
                for (unsigned o = 0; o < bytes; o += 32768u) {      // pieces of at most 32 KB
                    const unsigned n = bytes - o < 32768u ? bytes - o : 32768u;
                    bulk_g2s(stage + o, blob + d.blob_off + o, n, mbar);
                }
            }
        }
        __syncthreads();
        const unsigned long long it = (unsigned long long)*s_item;
        if (it >= n_items) break;
        const int t = (int)(it / (unsigned long long)n_seg);
        const int64_t seg = (int64_t)(it % (unsigned long long)n_seg);
        const TileDesc d = tdesc[t];
        const uint32_t post_s = stage_s;
        const uint32_t bitmap_s = stage_s + a16(4 * d.n_post);
        const uint32_t prefix_s = bitmap_s + 4 * bw;
        const uint32_t off_s = prefix_s + a16(2 * bw);
        const float tbound = tile_bound[t];
        const int col0 = t * TL_W;
        const int64_t rank_lo = seg * seg_ranks;
        const int64_t rank_hi = rank_lo + seg_ranks < n_ranks ? rank_lo + seg_ranks : n_ranks;
        const uint32_t *mrow = mask + (int64_t)(((t >> 6) << 1) | (t & 1)) * mask_stride;
        const int bit = (t & 63) >> 1;
        bool staged = false;

        int64_t base = rank_lo;
        while (base < rank_hi) {
            // ---- scan rounds: every warp tests 256 ranks per round (eight coalesced mask words in flight) and appends
            // the survivors to the CTA's list, until the list holds a full 32-pair chunk for every warp
            int count = 0;
            for (;;) {
                const int64_t gbase = base + ((int64_t)warp << 8);
                uint32_t wsv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int64_t r = gbase + j * 32 + lane;
                    wsv[j] = r < rank_hi ? mrow[r] : 0u;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const unsigned sv = __ballot_sync(FULL, (wsv[j] >> bit) & 1u);
                    if (sv) {
                        int pos = 0;
                        if (lane == 0) pos = atomicAdd(s_count, __popc(sv));
                        pos = __shfl_sync(FULL, pos, 0);
                        if ((sv >> lane) & 1u) s_list[pos + __popc(sv & lt_mask)] = (uint32_t)(gbase + j * 32 + lane);
                    }
                }
                base += (int64_t)NW * 256;
                __syncthreads();
                count = *reinterpret_cast<volatile int *>(s_count);
                if (count >= NW * 32 || base >= rank_hi) break;
                __syncthreads();          // everybody has read the count before the next round appends
            }
            if (count > 0) {
                if (!staged) {
                    mbar_wait(mbar, parity);
                    staged = true;
                }
                // ---- chunks of 32 consecutive list entries (neighbouring ranks: similar rows, similar work), dealt
                // round-robin to the warps; lane = pair
                for (int c0 = warp * 32; c0 < count; c0 += NW * 32) {
                    const int idx = c0 + lane;
                    int r = 0;
                    int4 info = make_int4(0, 0, 0, 0);
                    if (idx < count) {
                        r = (int)s_list[idx];
                        info = rowinfo[r];
                    }
                    const int nf = info.y;                      // 0: no pair in this lane
                    const float thr_r = __int_as_float(info.z);
                    const float xp = __int_as_float(info.w);
                    const float thr_f = xp > 0.f ? fmaxf(fmaf(-xp, tbound, thr_r), 0.f) : thr_r;
                    const unsigned thr_c = (unsigned)__float2uint_rd(fminf(thr_f, 1.99f) * TP_FIX);
                    if (nf > 0) ++n_pairs;
                    for (int fb0 = 0; __any_sync(FULL, fb0 < nf); fb0 += TP_FS) {
                        // features [fb0, fb0 + TP_FS) of every pair -> the pair's staging row (cp.async: 32 rows in
                        // flight, no registers); two pairs per step, 16 lanes each
                        for (int jj = 0; jj < 32; jj += 2) {
                            const int j = jj + (lane >> 4), k = lane & 15;
                            const int p0j = __shfl_sync(FULL, info.x, j);
                            const int nfj = __shfl_sync(FULL, nf, j);
                            if (fb0 + k < nfj)
                                cp_async8(warp_s + TP_ACC_BYTES + (j * TP_FSTRIDE + k) * 8, lpack + (int64_t)p0j + fb0 + k);
                        }
                        cp_async_wait_all();
                        __syncwarp();
                        // buckets of this lane's features through the bitmap directory, compacted in place
                        const int myn = nf - fb0 < TP_FS ? nf - fb0 : TP_FS;
                        int nb = 0;
                        for (int k = 0; __any_sync(FULL, k < myn); ++k) {
                            if (k < myn) {
                                const uint2 fa = lds64(feat_s + k * 8);
                                const unsigned f = fa.x;
                                const uint32_t bmw = lds32(bitmap_s + ((f >> 5) << 2));
                                if ((bmw >> (f & 31)) & 1u) {
                                    const uint32_t jb = lds16(prefix_s + ((f >> 5) << 1)) + __popc(bmw & ((1u << (f & 31)) - 1u));
                                    const uint32_t o0 = lds16(off_s + (jb << 1));
                                    const uint32_t o1 = lds16(off_s + (jb << 1) + 2);
                                    sts64(feat_s + nb * 8, o0 | ((o1 - o0) << 16), fa.y);
                                    ++nb;
                                }
                            }
                        }
                        // walk: one posting per lane and step; products a_q * w_q >> 15 into the lane's own column
                        int bi = 0;
                        uint32_t p = 0, pend = 0, aq = 0;
                        bool done = nb == 0;
                        for (;;) {
                            if (!done && p == pend) {
                                if (bi < nb) {
                                    const uint2 b = lds64(feat_s + bi * 8);
                                    ++bi;
                                    p = b.x & 0xffffu;
                                    pend = p + (b.x >> 16);
                                    aq = b.y;
                                    n_walked += (b.x >> 16);
                                } else {
                                    done = true;
                                }
                            }
                            if (__all_sync(FULL, done)) break;
                            bool crossed = false;
                            uint32_t cb = 0;
                            if (!done) {
                                const uint32_t e = lds32(post_s + (p << 2));
                                ++p;
                                const uint32_t x = ((e >> 16) * aq) >> 15;
                                cb = e & 0xffffu;                              // column * 4
                                const uint32_t a_addr = acc_s + (cb << 4);     // column * 64 bytes + lane * 2
                                const uint32_t old = lds16(a_addr);
                                const uint32_t now = old + x;
                                sts16(a_addr, now);
                                crossed = old <= thr_c && now > thr_c;
                            }
                            const unsigned em = __ballot_sync(FULL, crossed);
                            if (em) {
                                if (crossed) sts64(cbuf_s + (ccount + __popc(em & lt_mask)) * 8, (uint32_t)r, (uint32_t)(col0 + (int)(cb >> 2)));
                                ccount += __popc(em);
                                if (ccount > TL_CBUF - 32) flush();
                            }
                        }
                        __syncwarp();
                    }
                    // clear the warp's accumulator block (contiguous 16 KB)
#pragma unroll 4
                    for (int c = 0; c < TP_ACC_BYTES / 16 / 32; ++c) sts128z(warp_s + (c * 32 + lane) * 16);
                    __syncwarp();
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) *s_count = 0;
            __syncthreads();
        }
        if (!staged) mbar_wait(mbar, parity);     // nothing survived: still consume the phase before the stage is reused
        parity ^= 1u;
        __syncthreads();      // every warp is done with the staged tile before the next one is copied over it
    }
    flush();
    if (walk_stats) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            n_pairs += __shfl_xor_sync(FULL, n_pairs, o);
            n_walked += __shfl_xor_sync(FULL, n_walked, o);
        }
        if (lane == 0) {
            atomicAdd(walk_stats, n_pairs);
            atomicAdd(walk_stats + 1, n_walked);
        }
    }
}

static int bits_for64(uint64_t v) {
    int b = 0;
    while (v) { ++b; v >>= 1; }
    return b < 1 ? 1 : b;
}

}  // namespace sg

using namespace sg;

extern "C" {

int sg_tiles_tile_w(void) { return TL_W; }

int64_t sg_tiles_max_cols(void) { return (int64_t)1 << 18; }     // bitmap + rank table: 6 bytes per 32 features

/* upper bound of the blob array in bytes, known without a device read-back */
int64_t sg_tiles_blob_bound(int64_t nnz, int64_t n_rows, int64_t n_cols) {
    const int64_t T = sg_num_tiles(n_rows, TL_W);
    const int bw = bitmap_words(n_cols);
    return 6 * nnz + T * (int64_t)(4 * bw + a16(2 * bw) + 64) + 256;
}

size_t sg_tiles_workspace_bytes(int64_t nnz, int64_t n_rows, int64_t n_cols) {
    const int64_t T = sg_num_tiles(n_rows, TL_W);
    const int64_t n = nnz < 1 ? 1 : nnz;
    size_t b1 = 0, b2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b1, (uint64_t *)nullptr, (uint64_t *)nullptr, (uint32_t *)nullptr,
                                    (uint32_t *)nullptr, n);
    cub::DeviceScan::ExclusiveSum(nullptr, b2, (int64_t *)nullptr, (int64_t *)nullptr, T + 1);
    return 2 * align_up((size_t)n * 8, 256) + 2 * align_up((size_t)n * 4, 256) + 4 * align_up((size_t)(T + 2) * 8, 256) +
           align_up(b1 > b2 ? b1 : b2, 256) + 4096;
}

/*
 * Right matrix -> tile blobs, descriptors, block maxima.  `maxima` [dev, 2 x int32] receives the largest blob in
 * bytes and the largest posting count of a tile (the caller sizes the shared-memory stage with the former).
 */
int sg_tiles_build(int64_t n_rows, int64_t n_cols, int64_t nnz, const int64_t *indptr, const int32_t *indices,
                   const float *val32, const int32_t *rank, int64_t indptr_base, float w_scale, void *tile_desc,
                   void *blob, int64_t blob_cap, void *bucket_maxw, int32_t *maxima, void *ws, size_t ws_bytes,
                   void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    if (n_cols > sg_tiles_max_cols())
        return fail(SG_ERR_UNSUPPORTED, "%lld features exceed the bitmap directory of the tile kernel", (long long)n_cols);
    if (nnz >= (int64_t)0x7fffffff)
        return fail(SG_ERR_OVERFLOW, "right matrix nnz %lld does not fit int32 postings", (long long)nnz);
    const int64_t T = sg_num_tiles(n_rows, TL_W);
    const int64_t Tp = sg_num_tiles_padded(n_rows, TL_W);
    const int bw = bitmap_words(n_cols);
    if (blob_cap < sg_tiles_blob_bound(nnz, n_rows, n_cols)) return fail(SG_ERR_INVALID, "blob buffer too small");
    Arena ar(ws, ws_bytes);
    const size_t n = (size_t)(nnz < 1 ? 1 : nnz);
    uint64_t *keys = ar.take<uint64_t>(n);
    uint64_t *keys_sorted = ar.take<uint64_t>(n);
    uint32_t *vals = ar.take<uint32_t>(n);
    uint32_t *vals_sorted = ar.take<uint32_t>(n);
    int32_t *tile_ptr = ar.take<int32_t>((size_t)T + 2);
    int32_t *n_dist = ar.take<int32_t>((size_t)T + 2);
    int64_t *bytes = ar.take<int64_t>((size_t)T + 2);
    int64_t *blob_off = ar.take<int64_t>((size_t)T + 2);
    size_t b1 = 0, b2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b1, keys, keys_sorted, vals, vals_sorted, (int64_t)n);
    cub::DeviceScan::ExclusiveSum(nullptr, b2, bytes, blob_off, T + 1);
    size_t cub_bytes = b1 > b2 ? b1 : b2;
    char *tmp = ar.take<char>(cub_bytes);
    if (!ar.ok()) return fail(SG_ERR_INVALID, "tiles workspace too small (%zu < %zu)", ws_bytes, ar.off);
    SG_CUDA_TRY(cudaMemsetAsync(maxima, 0, 2 * sizeof(int32_t), st));
    SG_CUDA_TRY(cudaMemsetAsync(bucket_maxw, 0, (size_t)(n_cols + 1) * Tp * 2, st));
    SG_CUDA_TRY(cudaMemsetAsync(bytes, 0, (size_t)(T + 2) * 8, st));
    if (n_rows > 0 && nnz > 0) {
        tiles_keys_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, st>>>(n_rows, indptr, indices, val32, rank, TL_W,
                                                                      n_cols, indptr_base, w_scale, keys, vals);
        SG_LAUNCH_CHECK();
        const int bits = 16 + bits_for64((uint64_t)(T * n_cols));
        SG_CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, cub_bytes, keys, keys_sorted, vals, vals_sorted, nnz, 0,
                                                    bits > 64 ? 64 : bits, st));
    }
    tiles_ptr_kernel<<<(unsigned)((T + 1 + 255) / 256), 256, 0, st>>>(T, n_cols, nnz, keys_sorted, tile_ptr);
    SG_LAUNCH_CHECK();
    tiles_count_kernel<<<(unsigned)T, 256, 0, st>>>(T, tile_ptr, keys_sorted, bw, n_dist, bytes, maxima);
    SG_LAUNCH_CHECK();
    SG_CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp, cub_bytes, bytes, blob_off, T + 1, st));
    const size_t fill_smem = (size_t)bw * 4;
    SG_CUDA_TRY(cudaFuncSetAttribute(tiles_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fill_smem));
    tiles_fill_kernel<<<(unsigned)T, 256, fill_smem, st>>>(T, n_cols, Tp, tile_ptr, keys_sorted, vals_sorted, n_dist,
                                                           blob_off, bw, (unsigned char *)blob, (TileDesc *)tile_desc,
                                                           (unsigned short *)bucket_maxw);
    SG_LAUNCH_CHECK();
    return SG_OK;
}

/* pruned left rows -> packed operand of the tile kernels; `perm` = processing order of the n_ranks rows (NULL:
 * rows row_begin, row_begin+1, ...) */
int sg_tiles_pack_left(int64_t n_ranks, const int32_t *perm, int64_t row_begin, const int64_t *indptr,
                       const int32_t *pruned_len, const int32_t *pruned_indices, const float *pruned_val32,
                       const float *threshold_row, const float *pruned_norm_row, float a_scale, void *lpack,
                       void *rowinfo, void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    if (n_ranks <= 0) return SG_OK;
    if (!threshold_row) return fail(SG_ERR_INVALID, "threshold_row is required");
    pack_left_kernel<<<(unsigned)((n_ranks + 7) / 8), 256, 0, st>>>(n_ranks, perm, row_begin, indptr, pruned_len,
                                                                   pruned_indices, pruned_val32, threshold_row,
                                                                   pruned_norm_row, a_scale, (int2 *)lpack,
                                                                   (int4 *)rowinfo);
    SG_LAUNCH_CHECK();
    return SG_OK;
}

int64_t sg_tiles_mask_words(int64_t n_right) { return sg_num_tiles_padded(n_right, TL_W) / 32; }

/* survivors of the block-max test: mask[word * mask_stride + rank], mask_stride = n_ranks rounded up to 32 */
int sg_tiles_filter(int64_t n_ranks, const void *rowinfo, const void *lpack, const void *bucket_maxw,
                    int64_t n_right, const float *tile_bound, uint32_t *mask, int64_t mask_stride, void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    if (n_ranks <= 0 || n_right <= 0) return SG_OK;
    if (mask_stride < n_ranks || (mask_stride & 31)) return fail(SG_ERR_INVALID, "mask_stride must be n_ranks rounded up to 32");
    const int64_t T = sg_num_tiles(n_right, TL_W);
    const int Tp = (int)sg_num_tiles_padded(n_right, TL_W);
    tile_filter_kernel<<<(unsigned)(mask_stride / FL_RANKS), FL_WARPS * 32, 0, st>>>(
        n_ranks, (const int4 *)rowinfo, (const int2 *)lpack, (const uint32_t *)bucket_maxw, Tp, T, tile_bound, mask,
        mask_stride);
    SG_LAUNCH_CHECK();
    return SG_OK;
}

size_t sg_tiles_smem_bytes(int stage_bytes, int warps_per_cta) {
    return (size_t)TL_HEAD_BYTES + (size_t)tl_list(warps_per_cta) * 4 + (size_t)a16(stage_bytes) + (size_t)warps_per_cta * TL_WARP_BYTES2;
}

int sg_tiles_candidates(const int32_t *perm_a, int64_t n_ranks, int64_t row_begin, const void *rowinfo,
                        const void *lpack, const uint32_t *mask, int64_t mask_stride, const void *tile_desc,
                        const void *blob, int64_t n_right, int64_t n_cols, const float *tile_bound,
                        const int32_t *perm_b, int stage_bytes, int32_t *cand_row, int32_t *cand_col,
                        int64_t cand_cap, unsigned long long *cand_count, unsigned long long *queue,
                        unsigned long long *walk_stats, int warps_per_cta, void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    if (n_ranks <= 0 || n_right <= 0) return SG_OK;
    if (warps_per_cta != 4 && warps_per_cta != 6 && warps_per_cta != 8)
        return fail(SG_ERR_INVALID, "warps_per_cta must be 4, 6 or 8");
    int dev = 0, n_sm = 0, smem_optin = 0;
    SG_CUDA_TRY(cudaGetDevice(&dev));
    SG_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    SG_CUDA_TRY(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    stage_bytes = a16(stage_bytes);
    const size_t smem = sg_tiles_smem_bytes(stage_bytes, warps_per_cta);
    if (smem > (size_t)smem_optin)
        return fail(SG_ERR_UNSUPPORTED, "a tile of the right matrix needs %zu bytes of shared memory (limit %d)", smem,
                    smem_optin);
    const int64_t T = sg_num_tiles(n_right, TL_W);
    const int bw = bitmap_words(n_cols);
    int per_sm = 1;
    int64_t ctas = 1, n_seg = 1, seg_ranks = 0;
#define SG_TL_LAUNCH(NW)                                                                                             \
    do {                                                                                                             \
        SG_CUDA_TRY(cudaFuncSetAttribute(tile_candidates_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                         (int)smem));                                                                \
        SG_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tile_candidates_kernel<NW>, NW * 32, smem)); \
        if (per_sm < 1) per_sm = 1;                                                                                  \
        ctas = (int64_t)n_sm * per_sm;                                                                               \
        /* about 16 work items per resident CTA; a segment holds at least 1024 left ranks */                         \
        n_seg = (16 * ctas + T - 1) / T;                                                                             \
        if (n_seg > (n_ranks + 1023) / 1024) n_seg = (n_ranks + 1023) / 1024;                                        \
        if (n_seg < 1) n_seg = 1;                                                                                    \
        seg_ranks = ((n_ranks + n_seg - 1) / n_seg + 255) / 256 * 256;                                               \
        n_seg = (n_ranks + seg_ranks - 1) / seg_ranks;                                                               \
        if (ctas > T * n_seg) ctas = T * n_seg;                                                                      \
        tile_candidates_kernel<NW><<<(unsigned)ctas, NW * 32, smem, st>>>(                                           \
            perm_a, n_ranks, row_begin, (const int4 *)rowinfo, (const int2 *)lpack, mask, mask_stride,               \
            (const TileDesc *)tile_desc, (const unsigned char *)blob, T, bw, tile_bound, perm_b, seg_ranks, n_seg,   \
            cand_row, cand_col, (unsigned long long)cand_cap, cand_count, queue, walk_stats, stage_bytes);           \
    } while (0)
    if (warps_per_cta == 4) SG_TL_LAUNCH(4); else if (warps_per_cta == 6) SG_TL_LAUNCH(6); else SG_TL_LAUNCH(8);
#undef SG_TL_LAUNCH
    SG_LAUNCH_CHECK();
    return SG_OK;
}

}  // extern "C"
