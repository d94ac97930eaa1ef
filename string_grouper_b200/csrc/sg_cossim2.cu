// K2, second formulation: 2-D register/shared-memory tiling of C = A * B^T with posting reuse.
//
// Same contract as csrc/sg_cossim.cu (replaces StringGrouper._build_matches,
// /root/reference/string_grouper/string_grouper.py:709-752) and the same downstream stages
// (sg_rescore, sg_topn_select).  What changes is the candidate generator:
//
//   * both sides are taken in heavy-feature signature order (csrc/sg_order.cu);
//   * a warp owns a tile of R consecutive (permuted) left rows x Wc (permuted) columns, fp32, in
//     shared memory;
//   * the R rows' stored values are merged into a tile list sorted by feature: for every distinct
//     feature the list of (row, weight) pairs of the tile  (sg_left_tiles_build);
//   * for every feature of the tile, the posting bucket (column tile, feature) is read ONCE
//     (32 postings per warp step, one per lane, kept in registers) and applied to every row of the
//     tile that holds the feature:  acc[row][col] += a_row * w_col;
//   * the docs of a frequent feature are runs of consecutive columns in signature order and buckets
//     are sorted by column, so the 32 lanes of a step touch 32 different banks.
#include <cub/cub.cuh>

#include "sg_common.cuh"

namespace sg {

constexpr int K2_LCAP = 192;   // tile-list entries kept in shared memory (longer lists are read from HBM/L2)
constexpr int K2_LSEG = K2_LCAP + 4;   // segment slots (incl. sentinel), keeps the per-warp block 16-byte aligned
constexpr int K2_SHORT = 4;    // buckets up to this length are walked lane-privately

__host__ __device__ constexpr size_t k2_per_warp_bytes(int R, int Wc) {
    return (size_t)R * Wc * 4 + (size_t)K2_LCAP * 8 + (size_t)K2_LSEG * 8;
}

// ---------------------------------------------------------------------------
// left tile lists
// ---------------------------------------------------------------------------
__global__ void tiles_rowlen_kernel(int64_t n, const int64_t *__restrict__ indptr, const int32_t *__restrict__ perm,
                                    int64_t *__restrict__ len) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int64_t r = perm[i];
        len[i] = indptr[r + 1] - indptr[r];
    } else if (i == n) {
        len[i] = 0;
    }
}

// flip-bitonic sort of (key, payload) pairs by key; positions >= n act as +inf padding
__device__ void warp_sort_pairs(uint32_t *keys, uint32_t *vals, int n, int lane) {
    if (n < 2) return;
    int P = 2;
    while (P < n) P <<= 1;
    const int half = P >> 1;
    for (int k = 2; k <= P; k <<= 1) {
        const int hk = k >> 1;
        for (int i = lane; i < half; i += 32) {
            const int blk = i / hk, o = i - blk * hk;
            const int a = blk * k + o, b = blk * k + (k - 1 - o);
            if (b < n) {
                const uint32_t ka = keys[a], kb = keys[b];
                if (ka > kb) {
                    keys[a] = kb; keys[b] = ka;
                    const uint32_t va = vals[a]; vals[a] = vals[b]; vals[b] = va;
                }
            }
        }
        __syncwarp();
        for (int j = k >> 2; j >= 1; j >>= 1) {
            for (int i = lane; i < half; i += 32) {
                const int a = (i / j) * 2 * j + (i % j), b = a + j;
                if (b < n) {
                    const uint32_t ka = keys[a], kb = keys[b];
                    if (ka > kb) {
                        keys[a] = kb; keys[b] = ka;
                        const uint32_t va = vals[a]; vals[a] = vals[b]; vals[b] = va;
                    }
                }
            }
            __syncwarp();
        }
    }
}

// One warp per tile of R permuted rows: gather (feature << 8 | local row, weight) into scratch, sort by
// feature, write the (row, weight) list and the segment table {feature, start} (+ sentinel start = count).
// Tile t owns list slots [row_pos[t*R], row_pos[t*R + rows)) and segment slots starting at row_pos[t*R] + t
// (one extra slot per tile for the sentinel).
__global__ void __launch_bounds__(256)
tiles_build_kernel(int64_t n_rows, int R, const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                   const float *__restrict__ val, const int32_t *__restrict__ perm,
                   const int64_t *__restrict__ row_pos, int64_t n_tiles, uint32_t *__restrict__ tl_key,
                   uint32_t *__restrict__ tl_w, uint2 *__restrict__ tl_ra, int32_t *__restrict__ seg_f,
                   int32_t *__restrict__ seg_start, int32_t *__restrict__ tile_nseg) {
    const int lane = threadIdx.x & 31;
    const int64_t tile = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tile >= n_tiles) return;
    const int64_t r0 = tile * R;
    const int nr = (int)(n_rows - r0 < R ? n_rows - r0 : R);
    const int64_t base = row_pos[r0];
    const int cnt = (int)(row_pos[r0 + nr] - base);
    uint32_t *keys = tl_key + base;     // scratch, sorted in place (the working set of a warp stays in L1/L2)
    uint32_t *wts = tl_w + base;
    for (int r = 0; r < nr; ++r) {
        const int64_t row = perm[r0 + r];
        const int64_t p0 = indptr[row];
        const int len = (int)(indptr[row + 1] - p0);
        const int o = (int)(row_pos[r0 + r] - base);
        for (int k = lane; k < len; k += 32) {
            keys[o + k] = ((uint32_t)indices[p0 + k] << 8) | (uint32_t)r;
            wts[o + k] = __float_as_uint(val[p0 + k]);
        }
    }
    __syncwarp();
    warp_sort_pairs(keys, wts, cnt, lane);
    int32_t *sf = seg_f + base + tile;
    int32_t *ss = seg_start + base + tile;
    int nseg = 0;
    for (int b = 0; b < cnt; b += 32) {
        const int k = b + lane;
        const bool valid = k < cnt;
        const uint32_t key = valid ? keys[k] : 0u;
        const uint32_t f = valid ? (key >> 8) : 0xffffffffu;
        uint32_t prev = __shfl_up_sync(FULL, f, 1);
        if (lane == 0) prev = b > 0 ? (keys[b - 1] >> 8) : ~f;
        const bool head = valid && f != prev;
        const unsigned hb = __ballot_sync(FULL, head);
        if (head) {
            const int h = nseg + __popc(hb & ((1u << lane) - 1u));
            sf[h] = (int32_t)f;
            ss[h] = k;
        }
        if (valid) tl_ra[base + k] = make_uint2(key & 0xffu, wts[k]);
        nseg += __popc(hb);
    }
    if (lane == 0) {
        ss[nseg] = cnt;
        tile_nseg[tile] = nseg;
    }
}

// ---------------------------------------------------------------------------
// candidate generation, formulation 2
// ---------------------------------------------------------------------------
template <int NW, int R>
__global__ void __launch_bounds__(NW * 32)
cossim2_candidates_kernel(const int64_t *__restrict__ tile_ptr, const uint2 *__restrict__ tl_ra,
                          const int32_t *__restrict__ seg_f, const int32_t *__restrict__ seg_start,
                          const int32_t *__restrict__ tile_nseg, int64_t n_tiles_left, int64_t n_left_rows,
                          const int32_t *__restrict__ perm_a, const int32_t *__restrict__ bptr,
                          const uint2 *__restrict__ post, int64_t V1, int Wc, int64_t T, int64_t tiles_per_group,
                          int64_t n_right, const int32_t *__restrict__ perm_b, float thr_c,
                          int32_t *__restrict__ cand_row, int32_t *__restrict__ cand_col, unsigned long long cap,
                          unsigned long long *__restrict__ cand_count, unsigned long long *__restrict__ queue) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const size_t per_warp = k2_per_warp_bytes(R, Wc);
    unsigned char *mine = smem_raw + (size_t)warp * per_warp;
    float *acc = reinterpret_cast<float *>(mine);
    uint2 *s_ra = reinterpret_cast<uint2 *>(mine + (size_t)R * Wc * 4);
    int32_t *s_segf = reinterpret_cast<int32_t *>(s_ra + K2_LCAP);
    int32_t *s_segs = s_segf + K2_LSEG;
    const int tile_elems = R * Wc;

    for (int c = lane * 4; c < tile_elems; c += 128) *reinterpret_cast<float4 *>(acc + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();

    const int64_t n_groups = (T + tiles_per_group - 1) / tiles_per_group;
    const unsigned long long n_items = (unsigned long long)n_tiles_left * (unsigned long long)n_groups;
    for (;;) {
        unsigned long long item = 0;
        if (lane == 0) item = atomicAdd(queue, 1ull);
        item = __shfl_sync(FULL, item, 0);
        if (item >= n_items) break;
        const int64_t group = (int64_t)(item / (unsigned long long)n_tiles_left);
        const int64_t tile = (int64_t)(item % (unsigned long long)n_tiles_left);
        const int64_t lbase = tile_ptr[tile * R];
        const int64_t r_hi = (tile + 1) * R < n_left_rows ? (tile + 1) * R : n_left_rows;
        const int nr = (int)(r_hi - tile * R);
        const int cnt = (int)(tile_ptr[r_hi] - lbase);
        const int S = tile_nseg[tile];
        if (cnt == 0) continue;
        // tile list -> shared memory (element offset of the row inside the accumulator tile, weight)
        const uint2 *ra;
        const int32_t *segf, *segs;
        if (cnt <= K2_LCAP) {
            for (int k = lane; k < cnt; k += 32) {
                const uint2 e = tl_ra[lbase + k];
                s_ra[k] = make_uint2(e.x * (uint32_t)Wc, e.y);
            }
            for (int k = lane; k <= S; k += 32) {
                s_segs[k] = seg_start[lbase + tile + k];
                if (k < S) s_segf[k] = seg_f[lbase + tile + k];
            }
            __syncwarp();
            ra = s_ra; segf = s_segf; segs = s_segs;
        } else {
            ra = nullptr; segf = seg_f + lbase + tile; segs = seg_start + lbase + tile;
        }
        const uint2 *gra = tl_ra + lbase;   // long lists: rows still need the * Wc, done on the fly

        const int64_t t_begin = group * tiles_per_group;
        const int64_t t_end = t_begin + tiles_per_group < T ? t_begin + tiles_per_group : T;
        for (int64_t t = t_begin; t < t_end; ++t) {
            const int64_t c0 = t * Wc;
            const int32_t *bp = bptr + t * V1;
            for (int sb = 0; sb < S; sb += 32) {
                const int k = sb + lane;
                int b0 = 0, b1 = 0, i0 = 0, i1 = 0;
                if (k < S) {
                    const int f = segf[k];
                    i0 = segs[k];
                    i1 = segs[k + 1];
                    b0 = bp[f];
                    b1 = bp[f + 1];
                }
                const int len = b1 - b0;
                // ---- short buckets: lane-private walk; equal targets inside one step are serialised
                const bool is_short = len > 0 && len <= K2_SHORT;
                const unsigned short_mask = __ballot_sync(FULL, is_short);
                if (short_mask) {
                    const int my_len = is_short ? len : 0, my_cnt = is_short ? i1 - i0 : 0;
                    int max_len = my_len, max_cnt = my_cnt;
#pragma unroll
                    for (int o = 16; o; o >>= 1) {
                        max_len = max(max_len, __shfl_xor_sync(FULL, max_len, o));
                        max_cnt = max(max_cnt, __shfl_xor_sync(FULL, max_cnt, o));
                    }
                    for (int j = 0; j < max_len; ++j) {
                        uint2 e = make_uint2(0u, 0u);
                        if (j < my_len) e = post[b0 + j];
                        for (int ii = 0; ii < max_cnt; ++ii) {
                            const bool act = j < my_len && ii < my_cnt;
                            unsigned target = 0x80000000u | (unsigned)lane;
                            float a = 0.f;
                            if (act) {
                                const uint2 q = ra ? ra[i0 + ii] : make_uint2(gra[i0 + ii].x * (uint32_t)Wc, gra[i0 + ii].y);
                                target = q.x + e.x;
                                a = __uint_as_float(q.y);
                            }
                            const unsigned same = __match_any_sync(FULL, target);
                            const int ord = __popc(same & ((1u << lane) - 1u));
                            int rounds = act ? __popc(same) : 0;
#pragma unroll
                            for (int o = 16; o; o >>= 1) rounds = max(rounds, __shfl_xor_sync(FULL, rounds, o));
                            for (int rr = 0; rr < rounds; ++rr) {
                                if (act && ord == rr) acc[target] += a * __uint_as_float(e.y);
                                __syncwarp();
                            }
                        }
                    }
                }
                // ---- long buckets: one posting per lane in registers, applied to every row of the tile
                unsigned m = __ballot_sync(FULL, len > K2_SHORT);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const int s = __shfl_sync(FULL, b0, src);
                    const int e = __shfl_sync(FULL, b1, src);
                    const int j0 = __shfl_sync(FULL, i0, src);
                    const int j1 = __shfl_sync(FULL, i1, src);
                    int p = s + lane;
                    uint2 cur = p < e ? post[p] : make_uint2(0u, 0u);
                    for (; p - lane < e; p += 32) {
                        const bool act = p < e;
                        const uint2 nxt = (p + 32 < e) ? post[p + 32] : make_uint2(0u, 0u);
                        const float w = __uint_as_float(cur.y);
                        if (ra) {
                            for (int i = j0; i < j1; ++i) {
                                const uint2 q = ra[i];                 // same address in all lanes: broadcast
                                if (act) acc[q.x + cur.x] += __uint_as_float(q.y) * w;
                            }
                        } else {
                            for (int i = j0; i < j1; ++i) {
                                const uint2 q = gra[i];
                                if (act) acc[q.x * (uint32_t)Wc + cur.x] += __uint_as_float(q.y) * w;
                            }
                        }
                        cur = nxt;
                    }
                    __syncwarp();
                }
            }
            // ---- sweep the tile: report scores above the candidate threshold, clear
            const int sweep = nr * Wc;
            for (int c = lane * 4; c < sweep; c += 128) {
                float4 *q = reinterpret_cast<float4 *>(acc + c);
                const float4 v = *q;
                const unsigned nz = (__float_as_uint(v.x) | __float_as_uint(v.y) | __float_as_uint(v.z) |
                                     __float_as_uint(v.w)) << 1;
                if (nz) {
                    *q = make_float4(0.f, 0.f, 0.f, 0.f);
                    const float mx = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
                    if (mx > thr_c) {
                        const int r = c / Wc;
                        const int col = c - r * Wc;
                        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            if (vv[i] > thr_c) {
                                const unsigned long long slot = atomicAdd(cand_count, 1ull);
                                if (slot < cap) {
                                    cand_row[slot] = perm_a[tile * R + r];
                                    cand_col[slot] = perm_b[c0 + col + i];
                                }
                            }
                        }
                    }
                }
            }
            __syncwarp();
        }
    }
}

}  // namespace sg

using namespace sg;

extern "C" {

size_t sg_left_tiles_workspace_bytes(int64_t n_rows, int64_t nnz) {
    size_t b = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, b, (int64_t *)nullptr, (int64_t *)nullptr, n_rows + 1);
    return 2 * align_up((size_t)(n_rows + 2) * 8, 256) + 2 * align_up((size_t)(nnz + 1) * 4, 256) + align_up(b, 256) + 4096;
}

// Tile lists of the left rows perm[0..n_rows) (ids are absolute row numbers of the CSR), R rows per tile.
// Outputs: row_pos[n_rows+1] (entry offset of every permuted row; tile t starts at row_pos[t*R]),
// tl_ra[nnz] {local row, weight}, seg_f / seg_start [nnz + n_tiles + 1] (tile t at row_pos[t*R] + t), tile_nseg.
int sg_left_tiles_build(int64_t n_rows, int64_t nnz, int R, const int64_t *indptr, const int32_t *indices,
                        const float *val32, const int32_t *perm, int64_t *row_pos, void *tl_ra, int32_t *seg_f,
                        int32_t *seg_start, int32_t *tile_nseg, void *ws, size_t ws_bytes, void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    if (R < 1 || R > 255) return fail(SG_ERR_INVALID, "rows per tile must be in [1, 255]");
    if (n_rows <= 0) return SG_OK;
    Arena ar(ws, ws_bytes);
    int64_t *len = ar.take<int64_t>((size_t)n_rows + 2);
    uint32_t *tl_key = ar.take<uint32_t>((size_t)nnz + 1);
    uint32_t *tl_w = ar.take<uint32_t>((size_t)nnz + 1);
    size_t b = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, b, len, row_pos, n_rows + 1);
    char *tmp = ar.take<char>(b);
    if (!ar.ok()) return fail(SG_ERR_INVALID, "left tiles workspace too small (%zu < %zu)", ws_bytes, ar.off);
    tiles_rowlen_kernel<<<(unsigned)((n_rows + 1 + 255) / 256), 256, 0, st>>>(n_rows, indptr, perm, len);
    SG_LAUNCH_CHECK();
    SG_CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp, b, len, row_pos, n_rows + 1, st));
    const int64_t n_tiles = (n_rows + R - 1) / R;
    tiles_build_kernel<<<(unsigned)((n_tiles + 7) / 8), 256, 0, st>>>(
        n_rows, R, indptr, indices, val32, perm, row_pos, n_tiles, tl_key, tl_w, (uint2 *)tl_ra, seg_f, seg_start,
        tile_nseg);
    SG_LAUNCH_CHECK();
    return SG_OK;
}

}  // extern "C"

template <int NW, int R>
static int launch_cossim2(const int64_t *row_pos, const void *tl_ra, const int32_t *seg_f, const int32_t *seg_start,
                          const int32_t *tile_nseg, int64_t n_left_rows, const int32_t *perm_a,
                          const int32_t *bucket_ptr, const void *postings, int64_t n_cols, int tile_w,
                          int64_t tiles_per_group, int64_t n_right, const int32_t *perm_b, float thr_c,
                          int32_t *cand_row, int32_t *cand_col, int64_t cand_cap, unsigned long long *cand_count,
                          unsigned long long *queue, int n_sm, cudaStream_t st) {
    const size_t smem = k2_per_warp_bytes(R, tile_w) * NW;
    SG_CUDA_TRY(cudaFuncSetAttribute(cossim2_candidates_kernel<NW, R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
    int per_sm = 1;
    SG_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cossim2_candidates_kernel<NW, R>, NW * 32, smem));
    if (per_sm < 1) return fail(SG_ERR_INVALID, "cossim2: %zu bytes of shared memory per CTA do not fit", smem);
    const int64_t T = sg_num_tiles(n_right, tile_w);
    const int64_t n_tiles_left = (n_left_rows + R - 1) / R;
    int64_t ctas = (n_tiles_left + NW - 1) / NW;
    if (ctas > (int64_t)n_sm * per_sm) ctas = (int64_t)n_sm * per_sm;
    if (ctas < 1) ctas = 1;
    cossim2_candidates_kernel<NW, R><<<(unsigned)ctas, NW * 32, smem, st>>>(
        row_pos, (const uint2 *)tl_ra, seg_f, seg_start, tile_nseg, n_tiles_left, n_left_rows, perm_a, bucket_ptr,
        (const uint2 *)postings, n_cols + 1, tile_w, T, tiles_per_group < 1 ? 1 : tiles_per_group, n_right, perm_b,
        thr_c, cand_row, cand_col, (unsigned long long)cand_cap, cand_count, queue);
    SG_LAUNCH_CHECK();
    return SG_OK;
}

extern "C" {

size_t sg_cossim2_smem_bytes(int warps_per_cta, int rows_per_tile, int tile_w) {
    return k2_per_warp_bytes(rows_per_tile, tile_w) * warps_per_cta;
}

int sg_cossim2_candidates(const int64_t *row_pos, const void *tl_ra, const int32_t *seg_f, const int32_t *seg_start,
                          const int32_t *tile_nseg, int64_t n_left_rows, const int32_t *perm_a,
                          const int32_t *bucket_ptr, const void *postings, int64_t n_cols, int tile_w,
                          int64_t tiles_per_group, int64_t n_right, const int32_t *perm_b, float cand_threshold,
                          int32_t *cand_row, int32_t *cand_col, int64_t cand_cap, unsigned long long *cand_count,
                          unsigned long long *queue, int warps_per_cta, int rows_per_tile, void *stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    if (n_left_rows <= 0 || n_right <= 0) return SG_OK;
    if (tile_w <= 0 || (tile_w & 31)) return fail(SG_ERR_INVALID, "tile_w must be a positive multiple of 32");
    if (!(cand_threshold >= 0.f)) return fail(SG_ERR_INVALID, "cand_threshold must be >= 0");
    int dev = 0, n_sm = 0;
    SG_CUDA_TRY(cudaGetDevice(&dev));
    SG_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
#define SG_CASE2(NW, R)                                                                                          \
    if (warps_per_cta == NW && rows_per_tile == R)                                                               \
        return launch_cossim2<NW, R>(row_pos, tl_ra, seg_f, seg_start, tile_nseg, n_left_rows, perm_a,           \
                                     bucket_ptr, postings, n_cols, tile_w, tiles_per_group, n_right, perm_b,     \
                                     cand_threshold, cand_row, cand_col, cand_cap, cand_count, queue, n_sm, st);
    SG_CASE2(8, 4) SG_CASE2(8, 8) SG_CASE2(8, 16)
    SG_CASE2(16, 4) SG_CASE2(16, 8) SG_CASE2(16, 16)
    SG_CASE2(24, 4) SG_CASE2(24, 8)
    SG_CASE2(32, 2) SG_CASE2(32, 4) SG_CASE2(32, 8)
#undef SG_CASE2
    return fail(SG_ERR_INVALID, "unsupported (warps_per_cta=%d, rows_per_tile=%d)", warps_per_cta, rows_per_tile);
}

}  // extern "C"
