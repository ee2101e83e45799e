// HBM-bound / CUDA-core kernels of the STROTSS loss path: operand preparation (row norms, column
// sums, bf16 operand emission incl. transposes), the K=3 palette distance, the sparse relaxed-EMD
// backward, the small reductions and the final gradient assembly.  All reductions that feed a
// reported loss are fixed-order (two-stage partials), so results are run-to-run deterministic.
#pragma once
#include "common.cuh"

namespace sb {

constexpr float kL2NEps = 1e-12f;      // tf.nn.l2_normalize epsilon       (nn/losses.py:13-14)
constexpr float kL2DClamp = 1e-6f;     // l2_distance clamp               (nn/losses.py:23)
constexpr float kColsumClamp = 1e-12f; // self_similarity column-sum clamp (nn/losses.py:60,63)
constexpr int kRowsPerBlock = 32;      // rows per block in the column-partial kernels

// scalar slots written by the library (device float array, see include/strotss_b200.h)
enum Scalar {
    S_TOTAL = 0, S_LOSS_C = 1, S_LOSS_S = 2, S_LM = 3, S_LREMD = 4, S_LPAL = 5,
    S_REMD_RX = 6, S_REMD_RY = 7, S_LCOV = 8, S_LMEAN = 9, S_PAL_RX = 10, S_PAL_RY = 11,
    S_REMD_BRANCH = 12, S_PAL_BRANCH = 13, S_COUNT = 16
};

// --------------------------------------------------------------------------------------
// row statistics + per-block column partial sums
//   inv[r]            = rsqrt(max(sum_d x[r][d]^2, 1e-12))
//   part_raw[b][d]    = sum_{r in block b} x[r][d]              (-> column mean)
//   part_hat[b][d]    = sum_{r in block b} x[r][d] * inv[r]     (-> sum of normalised rows)
// --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) row_stats_kernel(const float* __restrict__ x, long long ld, int n, int D,
                                                        float* __restrict__ inv, float* __restrict__ part_raw,
                                                        float* __restrict__ part_hat, int rpb) {
    pdl_wait();
    __shared__ float s_inv[kRowsPerBlock];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = blockIdx.x * rpb;
    for (int rr = warp; rr < rpb; rr += 8) {
        const int r = r0 + rr;
        float ss = 0.f;
        if (r < n) {
            const float* xr = x + static_cast<long long>(r) * ld;
            for (int d = lane; d < D; d += 32) { const float v = xr[d]; ss = fmaf(v, v, ss); }
        }
        ss = warp_sum(ss);
        if (lane == 0) {
            const float iv = rsqrtf(fmaxf(ss, kL2NEps));
            s_inv[rr] = (r < n) ? iv : 0.f;
            if (r < n) inv[r] = iv;
        }
    }
    __syncthreads();
    const int rows = min(rpb, n - r0);
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float a = 0.f, h = 0.f;
        for (int rr = 0; rr < rows; ++rr) {
            const float v = x[static_cast<long long>(r0 + rr) * ld + d];
            a += v;
            h = fmaf(v, s_inv[rr], h);
        }
        if (part_raw) part_raw[static_cast<long long>(blockIdx.x) * D + d] = a;
        if (part_hat) part_hat[static_cast<long long>(blockIdx.x) * D + d] = h;
    }
}

// weighted sum of normalised rows: part[b][d] = sum_{r in block b} coef[r] * inv[r] * x[r][d]
__global__ void __launch_bounds__(256) weighted_colsum_kernel(const float* __restrict__ x, long long ld, int n, int D,
                                                              const float* __restrict__ inv, const float* __restrict__ coef,
                                                              float* __restrict__ part, int rpb) {
    pdl_wait();
    __shared__ float s_w[kRowsPerBlock];
    const int r0 = blockIdx.x * rpb;
    if (threadIdx.x < rpb) {
        const int r = r0 + threadIdx.x;
        s_w[threadIdx.x] = (r < n) ? coef[r] * inv[r] : 0.f;
    }
    __syncthreads();
    const int rows = min(rpb, n - r0);
#pragma unroll 3
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float a = 0.f;
#pragma unroll 8
        for (int rr = 0; rr < rows; ++rr) a = fmaf(x[static_cast<long long>(r0 + rr) * ld + d], s_w[rr], a);
        part[static_cast<long long>(blockIdx.x) * D + d] = a;
    }
}

// the same from the normalised bf16 rows: part[b][d] = sum_{r in block b} coef[r] * xh[r][d].  Four columns per thread (8-byte
// loads, eight rows in flight), blockIdx.y picks a chunk of 1024 columns; the rows hold Dp >= D zero-padded columns (Dp % 4 == 0).
__global__ void __launch_bounds__(256) weighted_colsum_bf16_kernel(const __nv_bfloat16* __restrict__ xh, long long ld, int n, int D,
                                                                   const float* __restrict__ coef, float* __restrict__ part, int rpb) {
    pdl_wait();
    __shared__ float s_w[kRowsPerBlock];
    const int r0 = blockIdx.x * rpb;
    if (threadIdx.x < rpb) {
        const int r = r0 + threadIdx.x;
        s_w[threadIdx.x] = (r < n) ? coef[r] : 0.f;
    }
    __syncthreads();
    const int rows = min(rpb, n - r0);
    const int d = blockIdx.y * 1024 + 4 * threadIdx.x;
    if (d >= D) return;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const __nv_bfloat16* src = xh + static_cast<long long>(r0) * ld + d;
#pragma unroll 8
    for (int rr = 0; rr < rows; ++rr) {
        const uint2 v = *reinterpret_cast<const uint2*>(src + static_cast<long long>(rr) * ld);
        const float w = s_w[rr];
        a0 = fmaf(__uint_as_float(v.x << 16), w, a0);
        a1 = fmaf(__uint_as_float(v.x & 0xffff0000u), w, a1);
        a2 = fmaf(__uint_as_float(v.y << 16), w, a2);
        a3 = fmaf(__uint_as_float(v.y & 0xffff0000u), w, a3);
    }
    float* dst = part + static_cast<long long>(blockIdx.x) * D + d;
    dst[0] = a0;
    if (d + 1 < D) dst[1] = a1;
    if (d + 2 < D) dst[2] = a2;
    if (d + 3 < D) dst[3] = a3;
}

// out[d] = scale * sum_b part[b][d]   (fixed order: 8 interleaved block groups, then a fixed tree)
// block = 32 columns x 8 block-groups
__global__ void __launch_bounds__(256) colsum_finish_kernel(const float* __restrict__ part, int nblocks, int D, float scale,
                                                            float* __restrict__ out) {
    pdl_wait();
    pdl_trigger();
    __shared__ float sh[8][33];
    const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int d = blockIdx.x * 32 + c;
    float a = 0.f;
    if (d < D)
        for (int b = g; b < nblocks; b += 8) a += part[static_cast<long long>(b) * D + d];
    sh[g][c] = a;
    __syncthreads();
    if (g == 0 && d < D) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += sh[k][c];
        out[d] = t * scale;
    }
}

// --------------------------------------------------------------------------------------
// bf16 operand emission, 64x64 tiles through shared memory.
//   xh  [r][d]  = x*inv                 (row-major, ld = Dp, zero in the K padding)
//   cen [r][d]  = x - mean[d]           (row-major, ld = Dp)
//   dlt [r][d]  = x*inv - y*inv_y       (row-major, ld = Dp)          (self-similarity delta form)
//   xhT [d][r]  = x*inv                 (transposed, ld = np, zero for r >= n)
//   cenT[d][r]  = x - mean[d]           (transposed, ld = np)
// Any output pointer may be null.
// --------------------------------------------------------------------------------------
struct EmitArgs {
    const float* x; long long ldx; int n, D, Dp, np;
    const float* inv; const float* mean;
    const float* y; long long ldy; const float* inv_y;      // only for dlt
    __nv_bfloat16* xh; __nv_bfloat16* cen; __nv_bfloat16* dlt;
    __nv_bfloat16* xhT; __nv_bfloat16* cenT;
};

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// Persistent, double-buffered variant for the calls without a delta operand (the fused evaluation): every block walks the
// 64 x 64 tiles with stride gridDim.x and fetches its next tile with cp.async while it converts and stores the current one.
__global__ void __launch_bounds__(256) emit_operands2_kernel(const EmitArgs a, int tiles_x, int tiles_y) {
    pdl_wait();
    pdl_trigger();
    __shared__ float sx[2][64][65];
    __shared__ float s_inv[2][64], s_mean[2][64];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int ntiles = tiles_x * tiles_y;
    auto fetch = [&](int t, int b) {
        const int col0 = (t % tiles_x) * 64, row0 = (t / tiles_x) * 64;
        if (threadIdx.x < 64) {
            const int r = row0 + threadIdx.x;
            s_inv[b][threadIdx.x] = (r < a.n) ? a.inv[r] : 0.f;
            const int c = col0 + threadIdx.x;
            s_mean[b][threadIdx.x] = (a.mean && c < a.D) ? a.mean[c] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int rr = grp * 8 + i, r = row0 + rr;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int cc = lane + 32 * h, c = col0 + cc;
                if (r < a.n && c < a.D) cp_async4(&sx[b][rr][cc], a.x + static_cast<long long>(r) * a.ldx + c);
                else sx[b][rr][cc] = 0.f;
            }
        }
        cp_async_commit();
    };
    int t = blockIdx.x, b = 0;
    if (t < ntiles) fetch(t, 0);
    for (; t < ntiles; t += gridDim.x, b ^= 1) {
        const int tn = t + gridDim.x;
        if (tn < ntiles) { fetch(tn, b ^ 1); cp_async_wait<1>(); } else cp_async_wait<0>();
        __syncthreads();
        const int col0 = (t % tiles_x) * 64, row0 = (t / tiles_x) * 64;
        for (int rr = grp * 8; rr < grp * 8 + 8; ++rr) {
            const int r = row0 + rr;
            if (r >= a.n) continue;
            const int cc = 2 * lane, c = col0 + cc;
            const float x0 = sx[b][rr][cc], x1 = sx[b][rr][cc + 1];
            const float iv = s_inv[b][rr];
            const long long off = static_cast<long long>(r) * a.Dp + c;
            if (a.xh) *reinterpret_cast<uint32_t*>(a.xh + off) = pack_bf16x2(x0 * iv, x1 * iv);
            if (a.cen) {
                const float c0 = (c < a.D) ? x0 - s_mean[b][cc] : 0.f;
                const float c1 = (c + 1 < a.D) ? x1 - s_mean[b][cc + 1] : 0.f;
                *reinterpret_cast<uint32_t*>(a.cen + off) = pack_bf16x2(c0, c1);
            }
        }
        if (a.xhT || a.cenT) {
            for (int cc = grp * 8; cc < grp * 8 + 8; ++cc) {
                const int c = col0 + cc;
                if (c >= a.D) continue;
                const int rr = 2 * lane, r = row0 + rr;
                if (r >= a.np) continue;
                const long long off = static_cast<long long>(c) * a.np + r;
                if (a.xhT)
                    *reinterpret_cast<uint32_t*>(a.xhT + off) = pack_bf16x2(sx[b][rr][cc] * s_inv[b][rr], sx[b][rr + 1][cc] * s_inv[b][rr + 1]);
                if (a.cenT) {
                    const float m = s_mean[b][cc];
                    const float c0 = (r < a.n) ? sx[b][rr][cc] - m : 0.f;
                    const float c1 = (r + 1 < a.n) ? sx[b][rr + 1][cc] - m : 0.f;
                    *reinterpret_cast<uint32_t*>(a.cenT + off) = pack_bf16x2(c0, c1);
                }
            }
        }
        __syncthreads();                    // buffer b may be refilled by the fetch of the next iteration
    }
}

__global__ void __launch_bounds__(256) emit_operands_kernel(const EmitArgs a) {
    pdl_wait();
    __shared__ float sx[64][65];
    __shared__ float sy[64][65];
    __shared__ float s_inv[64], s_invy[64], s_mean[64];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int col0 = blockIdx.x * 64, row0 = blockIdx.y * 64;
    if (threadIdx.x < 64) {
        const int r = row0 + threadIdx.x;
        s_inv[threadIdx.x] = (r < a.n) ? a.inv[r] : 0.f;
        s_invy[threadIdx.x] = (a.dlt && r < a.n) ? a.inv_y[r] : 0.f;
        const int c = col0 + threadIdx.x;
        s_mean[threadIdx.x] = (a.mean && c < a.D) ? a.mean[c] : 0.f;
    }
    for (int rr = grp * 8; rr < grp * 8 + 8; ++rr) {
        const int r = row0 + rr;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int cc = lane + 32 * h, c = col0 + cc;
            const bool ok = (r < a.n) && (c < a.D);
            sx[rr][cc] = ok ? a.x[static_cast<long long>(r) * a.ldx + c] : 0.f;
            sy[rr][cc] = (ok && a.dlt) ? a.y[static_cast<long long>(r) * a.ldy + c] : 0.f;
        }
    }
    __syncthreads();
    // row-major outputs: thread (grp, lane) -> rows grp*8.., columns 2*lane, 2*lane+1
    for (int rr = grp * 8; rr < grp * 8 + 8; ++rr) {
        const int r = row0 + rr;
        if (r >= a.n) continue;
        const int cc = 2 * lane, c = col0 + cc;
        const float x0 = sx[rr][cc], x1 = sx[rr][cc + 1];
        const float iv = s_inv[rr];
        const long long off = static_cast<long long>(r) * a.Dp + c;
        if (a.xh) *reinterpret_cast<uint32_t*>(a.xh + off) = pack_bf16x2(x0 * iv, x1 * iv);
        if (a.cen) {
            const float c0 = (c < a.D) ? x0 - s_mean[cc] : 0.f;
            const float c1 = (c + 1 < a.D) ? x1 - s_mean[cc + 1] : 0.f;
            *reinterpret_cast<uint32_t*>(a.cen + off) = pack_bf16x2(c0, c1);
        }
        if (a.dlt) {
            const float ivy = s_invy[rr];
            *reinterpret_cast<uint32_t*>(a.dlt + off) =
                pack_bf16x2(fmaf(x0, iv, -sy[rr][cc] * ivy), fmaf(x1, iv, -sy[rr][cc + 1] * ivy));
        }
    }
    // transposed outputs: thread (grp, lane) -> columns grp*8.., rows 2*lane, 2*lane+1
    if (a.xhT || a.cenT) {
        for (int cc = grp * 8; cc < grp * 8 + 8; ++cc) {
            const int c = col0 + cc;
            if (c >= a.D) continue;
            const int rr = 2 * lane, r = row0 + rr;
            if (r >= a.np) continue;
            const long long off = static_cast<long long>(c) * a.np + r;
            if (a.xhT)
                *reinterpret_cast<uint32_t*>(a.xhT + off) = pack_bf16x2(sx[rr][cc] * s_inv[rr], sx[rr + 1][cc] * s_inv[rr + 1]);
            if (a.cenT) {
                const float m = s_mean[cc];
                const float c0 = (r < a.n) ? sx[rr][cc] - m : 0.f;
                const float c1 = (r + 1 < a.n) ? sx[rr + 1][cc] - m : 0.f;
                *reinterpret_cast<uint32_t*>(a.cenT + off) = pack_bf16x2(c0, c1);
            }
        }
    }
}

// --------------------------------------------------------------------------------------
// Fused row pass over the prediction x and the content y (one read of each):
//   inv_x[r], inv_y[r]                          row norms
//   xh[r][:], yh[r][:]   = x^, y^               bf16, row-major, zero K padding
//   dlt[r][:]            = x^ - y^              bf16 (fp32 difference, then rounded)
//   part[b][0][d] = sum_r x[r][d]   part[b][1][d] = sum_r x^[r][d]   part[b][2][d] = sum_r y^[r][d]
// over the kPrRowsPerBlock rows of block b, in a fixed order (deterministic).
// Rows are staged in shared memory two at a time (2 x 2 x Dp floats = 36 KB), so six blocks fit on an SM
// and the loads of some blocks overlap the arithmetic/stores of the others.
// --------------------------------------------------------------------------------------
constexpr int kPrGroup = 2;               // rows per staging group
constexpr int kPrRowsPerBlock = 32;      // upper bound; small inputs use fewer rows per block to fill the chip

__global__ void __launch_bounds__(256, 4) prep_pair_rows_kernel(const float* __restrict__ x, long long ldx,
                                                             const float* __restrict__ y, long long ldy, int n, int D, int Dp,
                                                             float* __restrict__ inv_x, float* __restrict__ inv_y,
                                                             __nv_bfloat16* __restrict__ xh, __nv_bfloat16* __restrict__ yh,
                                                             __nv_bfloat16* __restrict__ dlt, float* __restrict__ part,
                                                             int rows_per_block) {
    pdl_wait();
    extern __shared__ float sm[];          // [2][kPrGroup][Dp]
    __shared__ float s_inv[2][kPrGroup];
    float* sx = sm;
    float* sy = sm + kPrGroup * Dp;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row_base = blockIdx.x * rows_per_block;
    constexpr int kMaxCols = 10;           // columns per thread: Dp <= 2560
    float ax[kMaxCols], ahx[kMaxCols], ahy[kMaxCols];
#pragma unroll
    for (int c = 0; c < kMaxCols; ++c) { ax[c] = 0.f; ahx[c] = 0.f; ahy[c] = 0.f; }

    for (int g = 0; g < rows_per_block; g += kPrGroup) {
        __syncthreads();                    // previous group fully consumed
#pragma unroll
        for (int rr = 0; rr < kPrGroup; ++rr) {
            const int r = row_base + g + rr;
            const bool live = r < n;
            const float* xr = x + static_cast<long long>(r) * ldx;
            const float* yr = y + static_cast<long long>(r) * ldy;
            // unrolled so that ~10 independent loads per thread are in flight (the kernel is latency-bound otherwise)
#pragma unroll 5
            for (int d = threadIdx.x; d < Dp; d += 256) {
                const bool ok = live && d < D;
                sx[rr * Dp + d] = ok ? xr[d] : 0.f;
                sy[rr * Dp + d] = ok ? yr[d] : 0.f;
            }
        }
        __syncthreads();
        // warps 0..G-1: norm of prediction row `warp`; warps G..2G-1: norm of content row `warp - G`
        if (warp < 2 * kPrGroup) {
            const float* src = (warp < kPrGroup) ? sx + warp * Dp : sy + (warp - kPrGroup) * Dp;
            float ss = 0.f;
            for (int d = lane; d < Dp; d += 32) { const float v = src[d]; ss = fmaf(v, v, ss); }
            ss = warp_sum(ss);
            if (lane == 0) s_inv[warp / kPrGroup][warp % kPrGroup] = rsqrtf(fmaxf(ss, kL2NEps));
        }
        __syncthreads();
        // emit: 8/G warps per row, 64-column chunks dealt round-robin (bf16x2 per lane = 128 B per warp store)
        {
            constexpr int kWarpsPerRow = 8 / kPrGroup;
            const int rr = warp / kWarpsPerRow, sub = warp % kWarpsPerRow;
            const int r = row_base + g + rr;
            if (r < n) {
                const float ix = s_inv[0][rr], iy = s_inv[1][rr];
                if (lane == 0 && sub == 0) { inv_x[r] = ix; inv_y[r] = iy; }
                const long long off = static_cast<long long>(r) * Dp;
                for (int d = sub * 64 + 2 * lane; d < Dp; d += 64 * kWarpsPerRow) {
                    const float x0 = sx[rr * Dp + d] * ix, x1 = sx[rr * Dp + d + 1] * ix;
                    const float y0 = sy[rr * Dp + d] * iy, y1 = sy[rr * Dp + d + 1] * iy;
                    *reinterpret_cast<uint32_t*>(xh + off + d) = pack_bf16x2(x0, x1);
                    *reinterpret_cast<uint32_t*>(yh + off + d) = pack_bf16x2(y0, y1);
                    *reinterpret_cast<uint32_t*>(dlt + off + d) = pack_bf16x2(x0 - y0, x1 - y1);
                }
            }
        }
        // column sums over the group's rows, thread t owns columns t, t+256, ...
#pragma unroll
        for (int c = 0; c < kMaxCols; ++c) {
            const int d = threadIdx.x + c * 256;
            if (d < D) {
#pragma unroll
                for (int rr = 0; rr < kPrGroup; ++rr) {
                    const float xv = sx[rr * Dp + d], yv = sy[rr * Dp + d];
                    ax[c] += xv;
                    ahx[c] = fmaf(xv, s_inv[0][rr], ahx[c]);
                    ahy[c] = fmaf(yv, s_inv[1][rr], ahy[c]);
                }
            }
        }
    }
    float* pb = part + static_cast<long long>(blockIdx.x) * 3 * D;
#pragma unroll
    for (int c = 0; c < kMaxCols; ++c) {
        const int d = threadIdx.x + c * 256;
        if (d < D) { pb[d] = ax[c]; pb[D + d] = ahx[c]; pb[2 * D + d] = ahy[c]; }
    }
}

// Double-buffered variant: the next group's rows are fetched with cp.async (4-byte LDGSTS, the rows are only 4-byte aligned)
// while the current group is normalised, emitted and column-summed, so every block always has a 36 KB load in flight and the
// three barriers per group no longer serialise memory latency with arithmetic.  Same outputs, same summation order.

__global__ void __launch_bounds__(256, 3) prep_pair_rows2_kernel(const float* __restrict__ x, long long ldx,
                                                                 const float* __restrict__ y, long long ldy, int n, int D, int Dp,
                                                                 float* __restrict__ inv_x, float* __restrict__ inv_y,
                                                                 __nv_bfloat16* __restrict__ xh, __nv_bfloat16* __restrict__ yh,
                                                                 __nv_bfloat16* __restrict__ dlt, float* __restrict__ part,
                                                                 int rows_per_block) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ float sm[];          // [2 buffers][x | y][kPrGroup][Dp]
    __shared__ float s_inv[2][kPrGroup];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row_base = blockIdx.x * rows_per_block;
    const int buf_floats = 2 * kPrGroup * Dp;
    constexpr int kMaxCols = 10;
    float ax[kMaxCols], ahx[kMaxCols], ahy[kMaxCols];
#pragma unroll
    for (int c = 0; c < kMaxCols; ++c) { ax[c] = 0.f; ahx[c] = 0.f; ahy[c] = 0.f; }
    // K padding columns stay zero in both buffers for the whole kernel
    for (int e = threadIdx.x; e < 2 * 2 * kPrGroup * (Dp - D); e += 256) {
        const int vec = e / (Dp - D), d = D + e % (Dp - D);
        sm[vec * Dp + d] = 0.f;
    }
    auto fetch = [&](int g, int b) {
        float* sx = sm + b * buf_floats;
        float* sy = sx + kPrGroup * Dp;
#pragma unroll
        for (int rr = 0; rr < kPrGroup; ++rr) {
            const int r = row_base + g + rr;
            if (r < n) {
                const float* xr = x + static_cast<long long>(r) * ldx;
                const float* yr = y + static_cast<long long>(r) * ldy;
#pragma unroll 3
                for (int d = threadIdx.x; d < D; d += 256) {
                    cp_async4(sx + rr * Dp + d, xr + d);
                    cp_async4(sy + rr * Dp + d, yr + d);
                }
            } else {
                for (int d = threadIdx.x; d < D; d += 256) { sx[rr * Dp + d] = 0.f; sy[rr * Dp + d] = 0.f; }
            }
        }
        cp_async_commit();
    };
    fetch(0, 0);
    int b = 0;
    for (int g = 0; g < rows_per_block; g += kPrGroup, b ^= 1) {
        if (g + kPrGroup < rows_per_block) { fetch(g + kPrGroup, b ^ 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();                    // group g landed (and the padding zeros of the first iteration)
        const float* sx = sm + b * buf_floats;
        const float* sy = sx + kPrGroup * Dp;
        if (warp < 2 * kPrGroup) {
            const float* src = (warp < kPrGroup) ? sx + warp * Dp : sy + (warp - kPrGroup) * Dp;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            int d = lane;
            for (; d + 96 < Dp; d += 128) {
                const float v0 = src[d], v1 = src[d + 32], v2 = src[d + 64], v3 = src[d + 96];
                s0 = fmaf(v0, v0, s0); s1 = fmaf(v1, v1, s1); s2 = fmaf(v2, v2, s2); s3 = fmaf(v3, v3, s3);
            }
            for (; d < Dp; d += 32) { const float v = src[d]; s0 = fmaf(v, v, s0); }
            const float ss = warp_sum((s0 + s1) + (s2 + s3));
            if (lane == 0) s_inv[warp / kPrGroup][warp % kPrGroup] = rsqrtf(fmaxf(ss, kL2NEps));
        }
        __syncthreads();
        {
            constexpr int kWarpsPerRow = 8 / kPrGroup;
            const int rr = warp / kWarpsPerRow, sub = warp % kWarpsPerRow;
            const int r = row_base + g + rr;
            if (r < n) {
                const float ix = s_inv[0][rr], iy = s_inv[1][rr];
                if (lane == 0 && sub == 0) { inv_x[r] = ix; inv_y[r] = iy; }
                const long long off = static_cast<long long>(r) * Dp;
                for (int d = sub * 64 + 2 * lane; d < Dp; d += 64 * kWarpsPerRow) {
                    const float x0 = sx[rr * Dp + d] * ix, x1 = sx[rr * Dp + d + 1] * ix;
                    const float y0 = sy[rr * Dp + d] * iy, y1 = sy[rr * Dp + d + 1] * iy;
                    *reinterpret_cast<uint32_t*>(xh + off + d) = pack_bf16x2(x0, x1);
                    *reinterpret_cast<uint32_t*>(yh + off + d) = pack_bf16x2(y0, y1);
                    *reinterpret_cast<uint32_t*>(dlt + off + d) = pack_bf16x2(x0 - y0, x1 - y1);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < kMaxCols; ++c) {
            const int d = threadIdx.x + c * 256;
            if (d < D) {
#pragma unroll
                for (int rr = 0; rr < kPrGroup; ++rr) {
                    const float xv = sx[rr * Dp + d], yv = sy[rr * Dp + d];
                    ax[c] += xv;
                    ahx[c] = fmaf(xv, s_inv[0][rr], ahx[c]);
                    ahy[c] = fmaf(yv, s_inv[1][rr], ahy[c]);
                }
            }
        }
        __syncthreads();                    // buffer b is free for the fetch issued in the next iteration
    }
    float* pb = part + static_cast<long long>(blockIdx.x) * 3 * D;
#pragma unroll
    for (int c = 0; c < kMaxCols; ++c) {
        const int d = threadIdx.x + c * 256;
        if (d < D) { pb[d] = ax[c]; pb[D + d] = ahx[c]; pb[2 * D + d] = ahy[c]; }
    }
}

// --------------------------------------------------------------------------------------
// Streaming row passes of the fused evaluation (prediction x and content y, both n x D fp32 with CONTIGUOUS, 16-byte
// aligned rows).  Groups of kRpGroup = 4 rows -- 4 * D floats are contiguous and a multiple of 16 bytes for any D -- are
// fetched by ONE elected thread with 1-D bulk copies (cp.async.bulk, completion on an mbarrier) into a two-stage
// shared-memory ring: the memory system always has a 2 x 35 KB request in flight per block and no thread spends issue
// slots on loads.  The row pass needs the column sums of ALL rows before it can emit the centred operand and the
// self-similarity vectors, so it is two passes over x and y -- the same number of reads as one row pass plus
// ss_vectors_kernel, without the third read and the transposed operands of emit_operands2_kernel (the GEMMs read the
// row-major operands through MN-major descriptors instead):
//   pass 1 (rows_stats3_kernel):  inv_x, inv_y, column partials  part[b][0] = sum x, [1] = sum x^, [2] = sum y^
//   pass 2 (rows_emit3_kernel):   x^, y^, delta = x^ - y^, cen = x - mean (bf16, row-major, zero K padding) and
//                                 u, w, sclamp (exactly what ss_vectors_kernel computes, from the rows already in smem)
// A ragged last group (n % 4 rows) is loaded by all threads with ordinary loads.
// --------------------------------------------------------------------------------------
constexpr int kRpGroup = 4;
constexpr int kRpThreads = 256;
constexpr int kRpMaxCols = 10;            // columns per thread in the column sums: Dp <= 2560

struct RowsArgs {
    const float* x; const float* y; int n, D, Dp;
    int rows_per_block;                   // multiple of kRpGroup
    float* inv_x; float* inv_y;
    float* part;                          // pass 1: [blocks][3][D]
    const float* mean; const float* sumhx; const float* sumhy;          // pass 2
    __nv_bfloat16* xh; __nv_bfloat16* yh; __nv_bfloat16* dlt; __nv_bfloat16* cen;
    int cen_r0, cen_r1;                   // rows whose centred operand is written (all rows, or this rank's shard)
    float* u; float* w; float* sclamp;
};

__host__ __device__ constexpr int rows_stage_floats(int D) { return 2 * kRpGroup * D; }

// Stage the row group starting at row r into (sx | sy): bulk copies for a full group (elected thread), ordinary loads
// by every thread for the ragged last group (zero fill for the missing rows).  Returns true if the group went by bulk copy.
__device__ __forceinline__ bool rows_group_is_bulk(const RowsArgs& a, int r) { return r + kRpGroup <= a.n; }

__device__ __forceinline__ void rows_issue_bulk(const RowsArgs& a, int r, float* stage, uint64_t* bar) {
    const uint32_t bytes = static_cast<uint32_t>(kRpGroup) * a.D * 4u;
    fence_proxy_async();                  // the stage was last read through the generic proxy
    mbar_arrive_expect_tx(bar, 2 * bytes);
    bulk_load_1d(stage, a.x + static_cast<long long>(r) * a.D, bytes, bar);
    bulk_load_1d(stage + kRpGroup * a.D, a.y + static_cast<long long>(r) * a.D, bytes, bar);
}

__device__ __forceinline__ void rows_load_ragged(const RowsArgs& a, int r, float* stage) {
    const int live = (a.n - r) * a.D;     // floats that exist (fewer than a full group)
    const float* xs = a.x + static_cast<long long>(r) * a.D;
    const float* ys = a.y + static_cast<long long>(r) * a.D;
    for (int e = threadIdx.x; e < kRpGroup * a.D; e += kRpThreads) {
        stage[e] = e < live ? xs[e] : 0.f;
        stage[kRpGroup * a.D + e] = e < live ? ys[e] : 0.f;
    }
}

__global__ void __launch_bounds__(kRpThreads, 1) rows_stats3_kernel(const RowsArgs a) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ __align__(16) float rp_smem[];      // [2 stages][x | y][kRpGroup][D]
    __shared__ uint64_t full[2];
    __shared__ float s_inv[2][kRpGroup];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = a.D;
    const int row_base = blockIdx.x * a.rows_per_block;
    const int row_end = min(a.n, row_base + a.rows_per_block);
    const int stage_floats = rows_stage_floats(D);
    if (threadIdx.x == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); fence_barrier_init(); }
    __syncthreads();
    float ax[kRpMaxCols], ahx[kRpMaxCols], ahy[kRpMaxCols];
#pragma unroll
    for (int c = 0; c < kRpMaxCols; ++c) { ax[c] = 0.f; ahx[c] = 0.f; ahy[c] = 0.f; }
    if (threadIdx.x == 0 && row_base < row_end && rows_group_is_bulk(a, row_base)) rows_issue_bulk(a, row_base, rp_smem, &full[0]);
    int g = 0;
    for (int r = row_base; r < row_end; r += kRpGroup, ++g) {
        const int s = g & 1;
        float* stage = rp_smem + s * stage_floats;
        const int rn = r + kRpGroup;
        if (threadIdx.x == 0 && rn < row_end && rows_group_is_bulk(a, rn)) rows_issue_bulk(a, rn, rp_smem + (s ^ 1) * stage_floats, &full[s ^ 1]);
        if (rows_group_is_bulk(a, r)) mbar_wait(&full[s], (g >> 1) & 1);
        else { rows_load_ragged(a, r, stage); __syncthreads(); }
        const float* sx = stage;
        const float* sy = stage + kRpGroup * D;
        {   // warp w: squared norm of row (w & 3) of x (w < 4) or y
            const float* src = (warp < kRpGroup ? sx : sy) + (warp & (kRpGroup - 1)) * D;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            int d = lane;
            for (; d + 96 < D; d += 128) {
                const float v0 = src[d], v1 = src[d + 32], v2 = src[d + 64], v3 = src[d + 96];
                s0 = fmaf(v0, v0, s0); s1 = fmaf(v1, v1, s1); s2 = fmaf(v2, v2, s2); s3 = fmaf(v3, v3, s3);
            }
            for (; d < D; d += 32) { const float v = src[d]; s0 = fmaf(v, v, s0); }
            const float ss = warp_sum((s0 + s1) + (s2 + s3));
            if (lane == 0) {
                const int row = r + (warp & (kRpGroup - 1));
                const float iv = rsqrtf(fmaxf(ss, kL2NEps));
                s_inv[warp / kRpGroup][warp & (kRpGroup - 1)] = iv;
                if (row < a.n) { if (warp < kRpGroup) a.inv_x[row] = iv; else a.inv_y[row] = iv; }
            }
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < kRpMaxCols; ++c) {
            const int d = threadIdx.x + c * kRpThreads;
            if (d < D) {
#pragma unroll
                for (int rr = 0; rr < kRpGroup; ++rr) {
                    const float xv = sx[rr * D + d], yv = sy[rr * D + d];
                    ax[c] += xv;
                    ahx[c] = fmaf(xv, s_inv[0][rr], ahx[c]);
                    ahy[c] = fmaf(yv, s_inv[1][rr], ahy[c]);
                }
            }
        }
        __syncthreads();                  // stage s (and s_inv) may be overwritten from here on
    }
    float* pb = a.part + static_cast<long long>(blockIdx.x) * 3 * D;
#pragma unroll
    for (int c = 0; c < kRpMaxCols; ++c) {
        const int d = threadIdx.x + c * kRpThreads;
        if (d < D) { pb[d] = ax[c]; pb[D + d] = ahx[c]; pb[2 * D + d] = ahy[c]; }
    }
}

__global__ void __launch_bounds__(kRpThreads, 1) rows_emit3_kernel(const RowsArgs a) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ __align__(16) float rp_smem[];      // [2 stages][x | y][kRpGroup][D] | mean[Dp] | sumhx[Dp] | sumhy[Dp]
    __shared__ uint64_t full[2];
    __shared__ float s_dot[2][kRpGroup][2][2];           // [stage][row][half][x | y]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = a.D, Dp = a.Dp;
    const int row_base = blockIdx.x * a.rows_per_block;
    const int row_end = min(a.n, row_base + a.rows_per_block);
    const int stage_floats = rows_stage_floats(D);
    float* s_mean = rp_smem + 2 * stage_floats;
    float* s_shx = s_mean + Dp;
    float* s_shy = s_shx + Dp;
    if (threadIdx.x == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x == 0 && row_base < row_end && rows_group_is_bulk(a, row_base)) rows_issue_bulk(a, row_base, rp_smem, &full[0]);
    for (int d = threadIdx.x; d < Dp; d += kRpThreads) {
        const bool ok = d < D;
        s_mean[d] = (ok && a.mean) ? a.mean[d] : 0.f;
        s_shx[d] = ok ? a.sumhx[d] : 0.f;
        s_shy[d] = ok ? a.sumhy[d] : 0.f;
    }
    __syncthreads();
    const int rr = warp >> 1, half = warp & 1;            // two warps per row, 64-column chunks dealt alternately
    const float fN = static_cast<float>(a.n);
    int g = 0;
    for (int r = row_base; r < row_end; r += kRpGroup, ++g) {
        const int s = g & 1;
        float* stage = rp_smem + s * stage_floats;
        const int rn = r + kRpGroup;
        if (threadIdx.x == 0 && rn < row_end && rows_group_is_bulk(a, rn)) rows_issue_bulk(a, rn, rp_smem + (s ^ 1) * stage_floats, &full[s ^ 1]);
        const int row = r + rr;
        const bool live = row < a.n;
        const float ix = live ? a.inv_x[row] : 0.f, iy = live ? a.inv_y[row] : 0.f;
        if (rows_group_is_bulk(a, r)) mbar_wait(&full[s], (g >> 1) & 1);
        else { rows_load_ragged(a, r, stage); __syncthreads(); }
        const float* sx = stage + rr * D;
        const float* sy = stage + (kRpGroup + rr) * D;
        const bool wcen = a.cen && row >= a.cen_r0 && row < a.cen_r1;
        const long long off = static_cast<long long>(row) * Dp;
        float dx0 = 0.f, dx1 = 0.f, dy0 = 0.f, dy1 = 0.f;
        if (live) {
#pragma unroll 2
            for (int d = half * 64 + 2 * lane; d < Dp; d += 128) {
                const bool ok0 = d < D, ok1 = d + 1 < D;
                const float x0 = ok0 ? sx[d] : 0.f, x1 = ok1 ? sx[d + 1] : 0.f;
                const float y0 = ok0 ? sy[d] : 0.f, y1 = ok1 ? sy[d + 1] : 0.f;
                const float hx0 = x0 * ix, hx1 = x1 * ix, hy0 = y0 * iy, hy1 = y1 * iy;
                *reinterpret_cast<uint32_t*>(a.xh + off + d) = pack_bf16x2(hx0, hx1);
                *reinterpret_cast<uint32_t*>(a.yh + off + d) = pack_bf16x2(hy0, hy1);
                *reinterpret_cast<uint32_t*>(a.dlt + off + d) = pack_bf16x2(hx0 - hy0, hx1 - hy1);
                if (wcen)
                    *reinterpret_cast<uint32_t*>(a.cen + off + d) = pack_bf16x2(ok0 ? x0 - s_mean[d] : 0.f, ok1 ? x1 - s_mean[d + 1] : 0.f);
                dx0 = fmaf(x0, s_shx[d], dx0); dx1 = fmaf(x1, s_shx[d + 1 < Dp ? d + 1 : d], dx1);
                dy0 = fmaf(y0, s_shy[d], dy0); dy1 = fmaf(y1, s_shy[d + 1 < Dp ? d + 1 : d], dy1);
            }
        }
        const float dx = warp_sum(dx0 + dx1), dy = warp_sum(dy0 + dy1);
        if (lane == 0) { s_dot[s][rr][half][0] = dx; s_dot[s][rr][half][1] = dy; }
        __syncthreads();                  // stage s may be overwritten; the dot-product halves are visible
        if (threadIdx.x < kRpGroup) {
            const int j = r + threadIdx.x;
            if (j < a.n) {
                // s_j = N - x^_j . sum_i x^_i,  t_j likewise (column sums of Xd, Yd); see ss_vectors_kernel
                const float ddx = (s_dot[s][threadIdx.x][0][0] + s_dot[s][threadIdx.x][1][0]) * a.inv_x[j];
                const float ddy = (s_dot[s][threadIdx.x][0][1] + s_dot[s][threadIdx.x][1][1]) * a.inv_y[j];
                const float sv = fN - ddx, tv = fN - ddy;
                const float sc = fmaxf(sv, kColsumClamp), tcl = fmaxf(tv, kColsumClamp);
                const bool plain = (sv >= kColsumClamp) && (tv >= kColsumClamp);
                const float tms = plain ? (ddx - ddy) : (tcl - sc);
                a.u[j] = 1.f / sc;
                a.w[j] = tms / (sc * tcl);
                a.sclamp[j] = (sv >= kColsumClamp) ? 1.f : 0.f;
            }
        }
    }
}

// out[q][d] = scale_q * sum_b part[b][q][d] for the three sums of prep_pair_rows_kernel (fixed order)
__global__ void __launch_bounds__(256) colsum3_finish_kernel(const float* __restrict__ part, int nblocks, int D,
                                                             float scale0, float* __restrict__ out0,
                                                             float* __restrict__ out1, float* __restrict__ out2) {
    pdl_wait();
    pdl_trigger();
    __shared__ float sh[8][33];
    const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int d = blockIdx.x * 32 + c;
    const int q = blockIdx.y;
    float a = 0.f;
    if (d < D)
        for (int b = g; b < nblocks; b += 8) a += part[(static_cast<long long>(b) * 3 + q) * D + d];
    sh[g][c] = a;
    __syncthreads();
    if (g == 0 && d < D) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += sh[k][c];
        if (q == 0) out0[d] = t * scale0;
        else if (q == 1) out1[d] = t;
        else out2[d] = t;
    }
}

// --------------------------------------------------------------------------------------
// self-similarity per-sample vectors (one warp per sample j):
//   s_j = N - x^_j . sum_i x^_i,  t_j = N - y^_j . sum_i y^_i      (column sums of Xd, Yd)
//   u_j = 1/max(s_j,1e-12),  w_j = 1/max(s_j,..) - 1/max(t_j,..)
// --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ss_vectors_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ invx,
                                                         const float* __restrict__ sumhx,
                                                         const float* __restrict__ y, long long ldy, const float* __restrict__ invy,
                                                         const float* __restrict__ sumhy,
                                                         int N, int D, float* __restrict__ u, float* __restrict__ w,
                                                         float* __restrict__ sclamp) {
    pdl_wait();
    const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (j >= N) return;
    const float* xr = x + static_cast<long long>(j) * ldx;
    const float* yr = y + static_cast<long long>(j) * ldy;
    // four independent partial sums per dot product keep 16 loads in flight per lane (the loop is latency-bound otherwise)
    float dx0 = 0.f, dx1 = 0.f, dx2 = 0.f, dx3 = 0.f, dy0 = 0.f, dy1 = 0.f, dy2 = 0.f, dy3 = 0.f;
    int d = lane;
    for (; d + 96 < D; d += 128) {
        dx0 = fmaf(xr[d], sumhx[d], dx0); dx1 = fmaf(xr[d + 32], sumhx[d + 32], dx1);
        dx2 = fmaf(xr[d + 64], sumhx[d + 64], dx2); dx3 = fmaf(xr[d + 96], sumhx[d + 96], dx3);
        dy0 = fmaf(yr[d], sumhy[d], dy0); dy1 = fmaf(yr[d + 32], sumhy[d + 32], dy1);
        dy2 = fmaf(yr[d + 64], sumhy[d + 64], dy2); dy3 = fmaf(yr[d + 96], sumhy[d + 96], dy3);
    }
    for (; d < D; d += 32) { dx0 = fmaf(xr[d], sumhx[d], dx0); dy0 = fmaf(yr[d], sumhy[d], dy0); }
    float dx = (dx0 + dx1) + (dx2 + dx3), dy = (dy0 + dy1) + (dy2 + dy3);
    dx = warp_sum(dx) * invx[j];
    dy = warp_sum(dy) * invy[j];
    if (lane == 0) {
        const float s = static_cast<float>(N) - dx, t = static_cast<float>(N) - dy;
        const float sc = fmaxf(s, kColsumClamp), tcl = fmaxf(t, kColsumClamp);
        const bool plain = (s >= kColsumClamp) && (t >= kColsumClamp);
        const float tms = plain ? (dx - dy) : (tcl - sc);       // t - s without the N - N cancellation
        u[j] = 1.f / sc;
        w[j] = tms / (sc * tcl);
        sclamp[j] = (s >= kColsumClamp) ? 1.f : 0.f;             // tf.maximum passes the gradient iff s >= clamp
    }
}

// r_i = (1/N) sum of the row-form partials (+ in symmetric mode the column-form partials written by the
// panels left of row i's panel);  coef_i = r_i u_i^2 [s_i >= clamp];  rowloss_i = sum of loss partials.
// In symmetric mode a panel only wrote the slots of column tiles >= its first row.
// Block = 32 rows x 8 partial-sum groups (fixed interleaved order, then a fixed tree: deterministic).
__global__ void __launch_bounds__(256) ss_rows_kernel(const float* __restrict__ loss_part, const float* __restrict__ r_part, int nslots, int N,
                               int r0, int r1, const float* __restrict__ u, const float* __restrict__ sclamp,
                               float* __restrict__ coef, float* __restrict__ rowloss,
                               int sym, int panel_rows, int slots_per_panel, const float* __restrict__ rcol_part,
                               int colparts_per_panel) {
    pdl_wait();
    __shared__ float sl[8][33], sr[8][33];
    const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int i = r0 + blockIdx.x * 32 + c;
    float l = 0.f, r = 0.f;
    if (i < r1) {
        const int pi = sym ? i / panel_rows : 0;
        for (int t = pi * slots_per_panel * sym + g; t < nslots; t += 8) {
            l += loss_part[static_cast<long long>(t) * N + i];
            r += r_part[static_cast<long long>(t) * N + i];
        }
        if (sym) {
            const int nb = pi * colparts_per_panel;
            for (int b = g; b < nb; b += 8) r += rcol_part[static_cast<long long>(b) * N + i];
        }
    }
    sl[g][c] = l; sr[g][c] = r;
    __syncthreads();
    if (g == 0 && i < r1) {
        float lt = 0.f, rt = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { lt += sl[k][c]; rt += sr[k][c]; }
        rowloss[i] = lt;
        const float ui = u[i];
        coef[i] = (rt / static_cast<float>(N)) * ui * ui * sclamp[i];
    }
}

// --------------------------------------------------------------------------------------
// Row-sharded symmetric self-similarity: a rank computes a list of rectangular "jobs" (own rows x a column range) of the
// symmetric matrix, every tile accounting for its mirror image as well (see self_sim_sharded_sym in api.cu).  After stage 1
//   rowloss[i]  = sum of the loss partials of row i over the column tiles of the jobs that contain row i
//   r_full[j]   = row-form partials (j an own row) + column-form partials (j a column of a job whose tiles were mirrored)
// r_full is then summed over ranks: every r_j has contributions on exactly the ranks that computed a tile of column/row j.
// The enumeration order is fixed (8 interleaved groups, then a fixed tree): deterministic.
// --------------------------------------------------------------------------------------
constexpr int kMaxSsJobs = 8;
struct SsJobList {
    int n;
    int r0[kMaxSsJobs], r1[kMaxSsJobs], c0[kMaxSsJobs], c1[kMaxSsJobs], diag[kMaxSsJobs];
};

__global__ void __launch_bounds__(256) ss_rows_jobs_kernel(const SsJobList jobs, const float* __restrict__ loss_part,
                                                           const float* __restrict__ r_part, const float* __restrict__ rcol_part,
                                                           int N, int slots_per_tile, float* __restrict__ rowloss,
                                                           float* __restrict__ r_full) {
    pdl_wait();
    __shared__ float sl[8][33], sr[8][33];
    const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + c;
    float l = 0.f, r = 0.f;
    if (i < N) {
        int e = 0;                                     // running entry index: entry e is summed by group e % 8
        for (int k = 0; k < jobs.n; ++k) {
            if (i >= jobs.r0[k] && i < jobs.r1[k]) {   // row-form partials: one slot per (column tile, column half)
                const int t0 = (jobs.diag[k] ? i / 256 : jobs.c0[k] / 256) * slots_per_tile;
                const int t1 = (jobs.c1[k] / 256) * slots_per_tile;
                for (int t = t0; t < t1; ++t, ++e)
                    if ((e & 7) == g) {
                        l += loss_part[static_cast<long long>(t) * N + i];
                        r += r_part[static_cast<long long>(t) * N + i];
                    }
            }
            if (i >= jobs.c0[k] && i < jobs.c1[k]) {   // column-form partials: one per (128-row block, epilogue warp quadrant)
                const int b0 = (jobs.r0[k] / 128) * 4;
                // trapezoid job: only the row tiles above the column's own diagonal tile wrote a column-form partial
                const int b1 = (jobs.diag[k] ? min((i / 256) * 2, jobs.r1[k] / 128) : jobs.r1[k] / 128) * 4;
                for (int b = b0; b < b1; ++b, ++e)
                    if ((e & 7) == g) r += rcol_part[static_cast<long long>(b) * N + i];
            }
        }
    }
    sl[g][c] = l; sr[g][c] = r;
    __syncthreads();
    if (g == 0 && i < N) {
        float lt = 0.f, rt = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { lt += sl[k][c]; rt += sr[k][c]; }
        rowloss[i] = lt;
        r_full[i] = rt;
    }
}

// coef_i = (r_i / N) u_i^2 [s_i >= clamp] for the rows [r0, r1) of this rank, r already summed over ranks
__global__ void ss_coef_kernel(const float* __restrict__ r_full, const float* __restrict__ u, const float* __restrict__ sclamp, int N,
                               int r0, int r1, float* __restrict__ coef) {
    pdl_wait();
    const int i = r0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < r1) { const float ui = u[i]; coef[i] = (r_full[i] / static_cast<float>(N)) * ui * ui * sclamp[i]; }
}

// Row-sharded symmetric self-similarity, after the exchange of the mirrored stage-2 products:
//   ss2[i][:] += full[i][:] (this rank's own products that landed in its own rows, if any) + sum_r recv_r[i - off_r][:]
constexpr int kMaxSsRecv = 8;
struct Ss2AddArgs {
    float* ss2; const float* full;            // both start at the rank's first row; full may be null
    int full_row0;                            // first row (relative to the rank's first row) that `full` holds products for
    long long row_floats;                     // floats per row (Dp)
    int rows;
    int nrecv;
    const float* recv[kMaxSsRecv]; int off[kMaxSsRecv], cnt[kMaxSsRecv];
};
__global__ void __launch_bounds__(256) ss2_add_kernel(const Ss2AddArgs a) {
    pdl_wait();
    const long long per_row4 = a.row_floats / 4;
    const long long total4 = per_row4 * a.rows;
    for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total4;
         t += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int i = static_cast<int>(t / per_row4);
        float4 v = reinterpret_cast<const float4*>(a.ss2)[t];
        if (a.full && i >= a.full_row0) {
            const float4 f = reinterpret_cast<const float4*>(a.full)[t];
            v.x += f.x; v.y += f.y; v.z += f.z; v.w += f.w;
        }
        for (int r = 0; r < a.nrecv; ++r) {
            if (i >= a.off[r] && i < a.off[r] + a.cnt[r]) {
                const float4 f = reinterpret_cast<const float4*>(a.recv[r])[t - per_row4 * a.off[r]];
                v.x += f.x; v.y += f.y; v.z += f.z; v.w += f.w;
            }
        }
        reinterpret_cast<float4*>(a.ss2)[t] = v;
    }
}

// --------------------------------------------------------------------------------------
// Row-sharded covariance, owner side: tile number lt = j * world + rank of the upper block triangle belongs to this rank.
// Its g partial sums (one per sender, written by EpiGramScatter into this rank's window) are added in rank order, turned into
// V - Vx, |.| partial sums (off-diagonal tiles count twice) and the sign matrix, which is stored -- the tile and, right of the
// diagonal, its mirror image -- into the Sg buffer of EVERY rank (peer-mapped pointers).  One block per 64 x 64 sub-tile.
// --------------------------------------------------------------------------------------
constexpr int kCovMaxRanks = 16;
struct CovOwnArgs {
    const float* slots;                    // [world][nslots][256 * 256]
    int world, rank, nslots, tiles;        // tiles: 256-wide tiles per side
    const float* Vx; long long ldv; float inv_n; int D;
    __nv_bfloat16* sg[kCovMaxRanks]; long long lds;
    float* part;                           // [gridDim.x]
};
__global__ void __launch_bounds__(256) cov_owner_kernel(const CovOwnArgs a) {
    pdl_wait();
    __shared__ __nv_bfloat16 sgn[64][72];
    __shared__ float red[8];
    const int j = blockIdx.x >> 4, sub = blockIdx.x & 15;
    int lt = j * a.world + a.rank;
    const int ntri = a.tiles * (a.tiles + 1) / 2;
    if (lt >= ntri) { if (threadIdx.x == 0) a.part[blockIdx.x] = 0.f; return; }
    // upper-triangle tile (ut, ut + lt); its slot holds the TRANSPOSED tile, i.e. the lower-triangle tile (tm, tn) = (ut + lt, ut)
    // of the symmetric matrix, row-major -- that is the tile this block works on; `mirror` below then writes the upper one
    int ut = 0;
    while (lt >= a.tiles - ut) { lt -= a.tiles - ut; ++ut; }
    const int tm = ut + lt, tn = ut;
    const int sr = sub >> 2, sc = sub & 3;
    const int c4 = (threadIdx.x & 15) * 4;
    float l1 = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = (threadIdx.x >> 4) + 16 * i;
        const long long off = static_cast<long long>(j) * 65536 + (sr * 64 + r) * 256 + sc * 64 + c4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < a.world; ++s) {
            const float4 v = *reinterpret_cast<const float4*>(a.slots + static_cast<long long>(s) * a.nslots * 65536 + off);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        const int grow = tm * 256 + sr * 64 + r, gcol = tn * 256 + sc * 64 + c4;
        const float av[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float sv = 0.f;
            if (grow < a.D && gcol + e < a.D) {
                const float d = fmaf(av[e], a.inv_n, -a.Vx[static_cast<long long>(grow) * a.ldv + gcol + e]);
                l1 += fabsf(d);
                sv = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
            }
            sgn[r][c4 + e] = __float2bfloat16(sv);
        }
    }
    __syncthreads();
    const bool mirror = tn != tm;
    for (int q = 0; q < a.world; ++q) {
        __nv_bfloat16* S = a.sg[q];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = (threadIdx.x >> 4) + 16 * i;
            {   // the tile itself: row grow, columns gcol .. gcol + 3
                const int grow = tm * 256 + sr * 64 + r, gcol = tn * 256 + sc * 64 + c4;
                if (grow < a.D) {
                    __nv_bfloat16* dst = S + static_cast<long long>(grow) * a.lds + gcol;
                    if (gcol + 3 < a.D) *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(&sgn[r][c4]);
                    else for (int e = 0; e < 4; ++e) if (gcol + e < a.D) dst[e] = sgn[r][c4 + e];
                }
            }
            if (mirror) {   // mirror image: row = a column of the tile, 4 consecutive tile rows along it
                const int mrow = tn * 256 + sc * 64 + r, mcol = tm * 256 + sr * 64 + c4;
                if (mrow < a.D) {
                    __nv_bfloat16* dst = S + static_cast<long long>(mrow) * a.lds + mcol;
                    if (mcol + 3 < a.D) {
                        __nv_bfloat16 v[4] = {sgn[c4][r], sgn[c4 + 1][r], sgn[c4 + 2][r], sgn[c4 + 3][r]};
                        *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(v);
                    } else {
                        for (int e = 0; e < 4; ++e) if (mcol + e < a.D) dst[e] = sgn[c4 + e][r];
                    }
                }
            }
        }
    }
    if (mirror) l1 *= 2.f;
    l1 = warp_sum(l1);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = l1;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k];
        a.part[blockIdx.x] = t;
    }
}

// --------------------------------------------------------------------------------------
// One-shot allreduce over the peer windows (row-sharded evaluation; replaces the small NCCL allreduces).  Every rank pushes its
// payload -- nf floats to be summed, nu packed u64 keys to be maximised -- into ITS slot of every rank's window (plain 16-byte
// stores to CUDA-IPC mapped pointers: NVLink for remote ranks), publishes a per-sender flag = epoch in every window
// (st.release.sys after a system-scope fence), waits until all senders' flags in its own window have reached the epoch
// (ld.acquire.sys, bounded spin -> trap, never a hang) and reduces the world slots in rank order -- the sum is bit-identical on
// all ranks and from run to run.  Slots are double-buffered by epoch parity: a sender can only be one collective ahead of the
// slowest rank, because finishing collective e needs every rank's flag e.
// --------------------------------------------------------------------------------------
constexpr int kArMaxRanks = 16;
struct PeerArArgs {
    unsigned char* push[kArMaxRanks];             // this rank's slot in rank q's window (current parity)
    unsigned long long* flag_out[kArMaxRanks];    // this rank's flag in rank q's window
    const unsigned char* slots;                   // own window: [world] slots of the current parity
    const unsigned long long* flags_in;           // own window: [world] flags
    long long slot_bytes;
    int world, rank;
    unsigned long long epoch;
    float* f; long long nf;                       // summed in place
    unsigned long long* u; long long nu;          // maximised in place
    unsigned int* counter;                        // zero between calls
};
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__global__ void __launch_bounds__(256) peer_allreduce_kernel(const PeerArArgs a) {
    pdl_wait();
    __shared__ bool last;
    const long long fbytes = (a.nf * 4 + 15) / 16 * 16;              // the u64 part starts 16-byte aligned
    const long long nf4 = a.nf / 4, nu2 = a.nu / 2;
    const long long gtid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x, gsz = static_cast<long long>(gridDim.x) * blockDim.x;
    // ---- push
    for (int q = 0; q < a.world; ++q) {
        float4* df = reinterpret_cast<float4*>(a.push[q]);
        const float4* sf = reinterpret_cast<const float4*>(a.f);
        for (long long i = gtid; i < nf4; i += gsz) df[i] = sf[i];
        for (long long i = nf4 * 4 + gtid; i < a.nf; i += gsz) reinterpret_cast<float*>(a.push[q])[i] = a.f[i];
        ulonglong2* du = reinterpret_cast<ulonglong2*>(a.push[q] + fbytes);
        const ulonglong2* su = reinterpret_cast<const ulonglong2*>(a.u);
        for (long long i = gtid; i < nu2; i += gsz) du[i] = su[i];
        for (long long i = nu2 * 2 + gtid; i < a.nu; i += gsz) reinterpret_cast<unsigned long long*>(a.push[q] + fbytes)[i] = a.u[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(a.counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (last) {
        __threadfence_system();
        if (threadIdx.x < a.world) st_release_sys(a.flag_out[threadIdx.x], a.epoch);
        if (threadIdx.x == 0) *a.counter = 0;
    }
    // ---- wait for every sender
    if (threadIdx.x < a.world) {
        unsigned int spins = 0;
        while (ld_acquire_sys(a.flags_in + threadIdx.x) < a.epoch) {
            __nanosleep(64);
            if (++spins > (1u << 24)) __trap();                      // ~ seconds: a rank died or the protocol is broken
        }
    }
    __syncthreads();
    // ---- reduce in rank order
    for (long long i = gtid; i < a.nf; i += gsz) {
        float acc = 0.f;
        for (int q = 0; q < a.world; ++q) acc += reinterpret_cast<const float*>(a.slots + q * a.slot_bytes)[i];
        a.f[i] = acc;
    }
    for (long long i = gtid; i < a.nu; i += gsz) {
        unsigned long long m = 0;
        for (int q = 0; q < a.world; ++q) {
            const unsigned long long v = reinterpret_cast<const unsigned long long*>(a.slots + q * a.slot_bytes + fbytes)[i];
            m = v > m ? v : m;
        }
        a.u[i] = m;
    }
}

// out[slot] = scale * sum_i in[i]      (single block, fixed order)
__global__ void __launch_bounds__(1024) reduce_sum_kernel(const float* __restrict__ in, int n, float scale, float* __restrict__ out) {
    pdl_wait();
    pdl_trigger();
    __shared__ float sh[32];
    float a = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) a += in[i];
    a = warp_sum(a);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x < 32) {
        float b = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.f;
        b = warp_sum(b);
        if (threadIdx.x == 0) *out = b * scale;
    }
}

// --------------------------------------------------------------------------------------
// relaxed EMD reductions.  cost = offset - value  (cosine: offset 1, value = max dot;
// palette: offset 0, value = -min cost).
//   best_partial:  *out = sum_{j in [0,n)} cost(best[j])        (this rank's prediction rows)
//   remd_finish:   R_X = mean_i cost(rowbest_i) over all M target rows (rowbest is already global),
//                  R_Y = *ry_sum / N (already summed over ranks), L = max(R_X, R_Y).
// tf.maximum sends the gradient to its FIRST argument (R_X) on ties -> branch = (R_X >= R_Y).
// --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) best_partial_kernel(const unsigned long long* __restrict__ best, int n, float offset,
                                                            float* __restrict__ out) {
    pdl_wait();
    pdl_trigger();
    __shared__ float sh[32];
    float a = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) a += offset - best_val(best[i]);
    a = warp_sum(a);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x < 32) {
        float b = warp_sum(sh[threadIdx.x]);
        if (threadIdx.x == 0) *out = b;
    }
}

__global__ void __launch_bounds__(1024) remd_finish_kernel(const unsigned long long* __restrict__ rowbest, int M,
                                                           const float* __restrict__ ry_sum, int N,
                                                           float offset, float* __restrict__ scalars, int slot_loss,
                                                           int slot_rx, int slot_ry, int slot_branch,
                                                           int* __restrict__ row_arg,
                                                           const unsigned long long* __restrict__ colbest, int r0, int r1,
                                                           int* __restrict__ col_arg) {
    pdl_wait();
    pdl_trigger();
    // ry_sum == nullptr (single GPU): the column sum is taken here as well, over colbest[r0, r1) == all N columns
    __shared__ float sh[32], shb[32];
    float a = 0.f, b = 0.f;
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
        const unsigned long long k = rowbest[i];
        a += offset - best_val(k);
        if (row_arg) row_arg[i] = static_cast<int>(best_idx(k));
    }
    if (col_arg || !ry_sum)
        for (int j = r0 + threadIdx.x; j < r1; j += blockDim.x) {
            const unsigned long long k = colbest[j];
            if (!ry_sum) b += offset - best_val(k);
            if (col_arg) col_arg[j] = static_cast<int>(best_idx(k));
        }
    a = warp_sum(a); b = warp_sum(b);
    if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5] = a; shb[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x < 32) {
        const float aa = warp_sum(sh[threadIdx.x]), bb = warp_sum(shb[threadIdx.x]);
        if (threadIdx.x == 0) {
            const float rx = aa / static_cast<float>(M), ry = (ry_sum ? *ry_sum : bb) / static_cast<float>(N);
            scalars[slot_rx] = rx; scalars[slot_ry] = ry;
            scalars[slot_loss] = fmaxf(rx, ry);
            scalars[slot_branch] = (rx >= ry) ? 1.f : 0.f;
        }
    }
}

// Sparse backward of the cosine relaxed EMD (gradient w.r.t. the NORMALISED prediction rows r0..r1 owned by
// this rank).  The branch flag lives in device memory (no host sync):
//   branch Y (R_Y > R_X): g^_j = -(1/N) x^_{argmin_j}: a pure gather, done inside finalize_grad_kernel.
//   branch X (R_X >= R_Y): g[argmin_i][:] += -(1/M) x^_i for every target row i whose match is owned here:
//            a scatter with atomics into a zeroed buffer; both kernels below return at once in branch Y.
__global__ void cond_zero_kernel(float* __restrict__ g, long long n, const float* __restrict__ scalars, int slot_branch) {
    pdl_wait();
    if (scalars[slot_branch] == 0.f) return;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    float4* g4 = reinterpret_cast<float4*>(g);
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n / 4; i += stride)
        g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) g[(n & ~3ll) + threadIdx.x] = 0.f;
}

__global__ void __launch_bounds__(256) remd_backward_kernel(const unsigned long long* __restrict__ rowbest, int M,
                                                            int r0, int r1,
                                                            const float* __restrict__ xs, long long ldxs,
                                                            const float* __restrict__ inv_s, int D,
                                                            const float* __restrict__ scalars, int slot_branch,
                                                            float* __restrict__ g, long long ldg) {
    pdl_wait();
    if (scalars[slot_branch] == 0.f) return;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const int j = static_cast<int>(best_idx(rowbest[row]));
    if (j < r0 || j >= r1) return;
    const float sc = -inv_s[row] / static_cast<float>(M);
    const float* src = xs + static_cast<long long>(row) * ldxs;
    float* dst = g + static_cast<long long>(j - r0) * ldg;
    for (int d = lane; d < D; d += 32) atomicAdd(dst + d, src[d] * sc);
}

// --------------------------------------------------------------------------------------
// palette (nn/strotss_utils.py:166-167 + nn/losses.py:12-28 'both' on 3 channels)
// rec = { y,u,v, y^,u^,v^, |yuv|^2, inv_norm }
// --------------------------------------------------------------------------------------
__constant__ float c_rgb2yuv[9] = {0.299f, -0.14714119f, 0.61497538f,
                                   0.587f, -0.28886916f, -0.51496512f,
                                   0.114f, 0.43601035f, -0.10001026f};

// srec = { y^,u^,v^, c*y,c*u,c*v, |yuv|^2/3, 0 } with c = sqrt(2/3): the candidate search evaluates
//   cost = 1 - a^.b^ + sqrt(max(|a|^2/3 + |b|^2/3 - (c a).(c b), 1e-6/3))      (= nn/losses.py:12-28 'both', D = 3)
// in 13 instructions per pair.
__global__ void pal_prep_kernel(const float* __restrict__ x, long long ld, int n, int convert, float* __restrict__ rec,
                                float* __restrict__ srec) {
    pdl_wait();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const float* xr = x + static_cast<long long>(r) * ld;
    const float R = xr[0], G = xr[1], B = xr[2];
    float y, u, v;
    if (convert) {
        y = R * c_rgb2yuv[0] + G * c_rgb2yuv[3] + B * c_rgb2yuv[6];
        u = R * c_rgb2yuv[1] + G * c_rgb2yuv[4] + B * c_rgb2yuv[7];
        v = R * c_rgb2yuv[2] + G * c_rgb2yuv[5] + B * c_rgb2yuv[8];
    } else { y = R; u = G; v = B; }
    const float sq = y * y + u * u + v * v;
    const float iv = rsqrtf(fmaxf(sq, kL2NEps));
    float* o = rec + static_cast<long long>(r) * 8;
    o[0] = y; o[1] = u; o[2] = v; o[3] = y * iv; o[4] = u * iv; o[5] = v * iv; o[6] = sq; o[7] = iv;
    if (srec) {
        const float c = 0.81649658092772603f;          // sqrt(2/3)
        float* s = srec + static_cast<long long>(r) * 8;
        s[0] = y * iv; s[1] = u * iv; s[2] = v * iv; s[3] = c * y; s[4] = c * u; s[5] = c * v; s[6] = sq * (1.f / 3.f); s[7] = 0.f;
    }
}

// best[q] = max over keys of -cost(q, key)  (packed; ties -> lowest key index).
// Each thread owns kPalQT queries (register blocking), keys stream through shared memory as float4
// pairs, and the key range is split over blockIdx.y so the grid fills the chip.
// sqrt.approx (MUFU, <= 1 ulp-class error): the candidate search needs ordering, not the last bit;
// remd_finish recomputes nothing from it beyond the selected minimum (relative error ~1e-7).
__device__ __forceinline__ float fast_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
constexpr int kPalQT = 4;
constexpr int kPalThreads = 128;
constexpr int kPalKeyTile = 256;

__global__ void __launch_bounds__(kPalThreads) pal_min_kernel(const float* __restrict__ qrec, int nq, const float* __restrict__ krec, int nk,
                                                              int kchunk, int mode, int kidx_base,
                                                              unsigned long long* __restrict__ best) {
    pdl_wait();
    pdl_trigger();
    __shared__ float4 sk[kPalKeyTile * 2];
    float a[kPalQT][8];
    float bv[kPalQT];
    int bi[kPalQT];
    const int qbase = blockIdx.x * (kPalThreads * kPalQT) + threadIdx.x;
#pragma unroll
    for (int t = 0; t < kPalQT; ++t) {
        const int q = qbase + t * kPalThreads;
        const float4* src = reinterpret_cast<const float4*>(qrec + static_cast<long long>(q < nq ? q : 0) * 8);
        const float4 v0 = src[0], v1 = src[1];
        a[t][0] = v0.x; a[t][1] = v0.y; a[t][2] = v0.z; a[t][3] = v0.w;
        a[t][4] = v1.x; a[t][5] = v1.y; a[t][6] = v1.z; a[t][7] = v1.w;
        bv[t] = INFINITY; bi[t] = 0;
    }
    const int k0 = blockIdx.y * kchunk, k1 = min(nk, k0 + kchunk);
    for (int kb = k0; kb < k1; kb += kPalKeyTile) {
        const int cnt = min(kPalKeyTile, k1 - kb);
        __syncthreads();
        const float4* ksrc = reinterpret_cast<const float4*>(krec + static_cast<long long>(kb) * 8);
        for (int e = threadIdx.x; e < cnt * 2; e += kPalThreads) sk[e] = ksrc[e];
        __syncthreads();
#pragma unroll 2
        for (int k = 0; k < cnt; ++k) {
            const float4 b0 = sk[2 * k], b1 = sk[2 * k + 1];       // (y^,u^,v^,c*y) (c*u,c*v,|.|^2/3,0)
#pragma unroll
            for (int t = 0; t < kPalQT; ++t) {
                float c = 0.f;
                if (mode != 1) c = fmaf(-a[t][2], b0.z, fmaf(-a[t][1], b0.y, fmaf(-a[t][0], b0.x, 1.f)));
                if (mode != 0) {
                    const float m = fmaf(-a[t][5], b1.y, fmaf(-a[t][4], b1.x, fmaf(-a[t][3], b0.w, a[t][6] + b1.z)));
                    c += fast_sqrt(fmaxf(m, kL2DClamp * (1.f / 3.f)));
                }
                if (c < bv[t]) { bv[t] = c; bi[t] = kb + k; }
            }
        }
    }
    if (k1 > k0) {
#pragma unroll
        for (int t = 0; t < kPalQT; ++t) {
            const int q = qbase + t * kPalThreads;
            if (q < nq) atomicMax(best + q, pack_best(-bv[t], static_cast<uint32_t>(kidx_base + bi[t])));
        }
    }
}

// Both directions of the same cost matrix in ONE pass: every (query, key) cost is evaluated once and feeds the
// per-thread minimum over keys (qbest, as above) and the minimum over queries of each key (kbest): minimum over the
// thread's own queries, then redux.sync over the warp, one packed atomicMax per (warp, key).  Ties -> lowest index,
// exactly as two separate passes would resolve them.
template <int mode>
__global__ void __launch_bounds__(kPalThreads) pal_min2_kernel(const float* __restrict__ qrec, int nq, const float* __restrict__ krec, int nk,
                                                               int kchunk, int kidx_base,
                                                               unsigned long long* __restrict__ qbest,
                                                               unsigned long long* __restrict__ kbest) {
    pdl_wait();
    pdl_trigger();
    __shared__ float4 sk[kPalKeyTile * 2];
    float a[kPalQT][7];
    float one[kPalQT];
    float bv[kPalQT];
    int bi[kPalQT];
    const int qbase = blockIdx.x * (kPalThreads * kPalQT) + threadIdx.x;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int t = 0; t < kPalQT; ++t) {
        const int q = qbase + t * kPalThreads;
        const bool valid = q < nq;
        const float4* src = reinterpret_cast<const float4*>(qrec + static_cast<long long>(valid ? q : 0) * 8);
        const float4 v0 = src[0], v1 = src[1];
        a[t][0] = v0.x; a[t][1] = v0.y; a[t][2] = v0.z; a[t][3] = v0.w;
        a[t][4] = v1.x; a[t][5] = v1.y;
        // a query beyond nq must never win a key's minimum: its costs are +inf in every mode
        a[t][6] = valid ? v1.z : INFINITY;
        one[t] = valid ? 1.f : INFINITY;
        bv[t] = INFINITY; bi[t] = 0;
    }
    const int k0 = blockIdx.y * kchunk, k1 = min(nk, k0 + kchunk);
    for (int kb = k0; kb < k1; kb += kPalKeyTile) {
        const int cnt = min(kPalKeyTile, k1 - kb);
        __syncthreads();
        const float4* ksrc = reinterpret_cast<const float4*>(krec + static_cast<long long>(kb) * 8);
        for (int e = threadIdx.x; e < cnt * 2; e += kPalThreads) sk[e] = ksrc[e];
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < cnt; ++k) {
            const float4 b0 = sk[2 * k], b1 = sk[2 * k + 1];
            float cm = INFINITY; int tm = 0;
#pragma unroll
            for (int t = 0; t < kPalQT; ++t) {
                float c = 0.f;
                if (mode != 1) c = fmaf(-a[t][2], b0.z, fmaf(-a[t][1], b0.y, fmaf(-a[t][0], b0.x, one[t])));
                if (mode != 0) {
                    const float m = fmaf(-a[t][5], b1.y, fmaf(-a[t][4], b1.x, fmaf(-a[t][3], b0.w, a[t][6] + b1.z)));
                    c += fast_sqrt(fmaxf(m, kL2DClamp * (1.f / 3.f)));
                }
                if (c < bv[t]) { bv[t] = c; bi[t] = kb + k; }
                if (c < cm) { cm = c; tm = t; }               // ties keep the lower t == the lower query index
            }
            const uint32_t ord = f2ord(-cm);
            const uint32_t wmax = __reduce_max_sync(0xffffffffu, ord);
            const uint32_t cand = (ord == wmax) ? static_cast<uint32_t>(qbase + tm * kPalThreads) : 0xffffffffu;
            const uint32_t qmin = __reduce_min_sync(0xffffffffu, cand);
            if (lane == 0 && qmin < static_cast<uint32_t>(nq))
                atomicMax(kbest + kb + k, (static_cast<unsigned long long>(wmax) << 32) | static_cast<unsigned long long>(~qmin));
        }
    }
    if (k1 > k0) {
#pragma unroll
        for (int t = 0; t < kPalQT; ++t) {
            const int q = qbase + t * kPalThreads;
            if (q < nq) atomicMax(qbest + q, pack_best(-bv[t], static_cast<uint32_t>(kidx_base + bi[t])));
        }
    }
}

// Sparse backward of the 3-channel relaxed EMD w.r.t. the prediction's RGB (through the YUV matrix).
// gpal[j - r0][0..2] += gradient for prediction rows owned by this rank; one thread per selected pair.
__global__ void pal_backward_kernel(const unsigned long long* __restrict__ rowbest, int M,
                                    const unsigned long long* __restrict__ colbest, int N, int r0, int r1,
                                    const float* __restrict__ arec, const float* __restrict__ brec, int mode, int convert,
                                    const float* __restrict__ scalars, int slot_branch, float* __restrict__ gpal) {
    pdl_wait();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool bx = scalars[slot_branch] != 0.f;
    int i, j; float wgt;
    if (bx) {
        if (t >= M) return;
        i = t; j = static_cast<int>(best_idx(rowbest[t])); wgt = 1.f / static_cast<float>(M);
        if (j < r0 || j >= r1) return;
    } else {
        j = r0 + t;
        if (j >= r1) return;
        i = static_cast<int>(best_idx(colbest[j])); wgt = 1.f / static_cast<float>(N);
    }
    const float* a = arec + static_cast<long long>(i) * 8;
    const float* b = brec + static_cast<long long>(j) * 8;
    float g[3] = {0.f, 0.f, 0.f};
    if (mode != 1) {
        // d(1 - a^.b^)/db = -(a^ - b^ (a^.b^)) / |b|   (no projection when |b|^2 < 1e-12)
        const float dot = a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
        const bool live = b[6] >= kL2NEps;
#pragma unroll
        for (int e = 0; e < 3; ++e) g[e] += -(a[3 + e] - (live ? b[3 + e] * dot : 0.f)) * b[7];
    }
    if (mode != 0) {
        const float m = a[6] + b[6] - 2.f * (a[0] * b[0] + a[1] * b[1] + a[2] * b[2]);
        if (m >= kL2DClamp) {
            const float l2 = sqrtf(m / 3.f);
#pragma unroll
            for (int e = 0; e < 3; ++e) g[e] += (b[e] - a[e]) / (3.f * l2);
        }
    }
    float o[3];
    if (convert) {
#pragma unroll
        for (int c = 0; c < 3; ++c) o[c] = g[0] * c_rgb2yuv[3 * c] + g[1] * c_rgb2yuv[3 * c + 1] + g[2] * c_rgb2yuv[3 * c + 2];
    } else { o[0] = g[0]; o[1] = g[1]; o[2] = g[2]; }
#pragma unroll
    for (int c = 0; c < 3; ++c) atomicAdd(gpal + static_cast<long long>(j - r0) * 4 + c, o[c] * wgt);
}

// --------------------------------------------------------------------------------------
// Hypercolumn sampler (SURVEY 8f "next #1"; Sampling._sample, nn/strotss_utils.py:25-81):
// out[i][:] = concat_k gather_k(indices_i), feature maps NHWC, nearest or 4-tap bilinear.  One block per
// sample; the index rescaling (`indices /= y`, :36-37) is replayed per map in fp32 in the reference's order,
// and products/sums are left unfused so the result is bit-identical to an fp32 NumPy/TF evaluation.
// --------------------------------------------------------------------------------------
constexpr int kMaxSamplerMaps = 16;
struct SamplerMaps {
    const float* ptr[kMaxSamplerMaps];
    float* gptr[kMaxSamplerMaps];          // gradient buffers (backward; may be null per map)
    int h[kMaxSamplerMaps], w[kMaxSamplerMaps], c[kMaxSamplerMaps], off[kMaxSamplerMaps];
    float div[kMaxSamplerMaps];            // divisor applied to the running indices at this map (1 = none)
    int nmaps;
};

struct SamplerTaps { int a, b, c, d; float wa, wb, wc, wd; };

__device__ __forceinline__ SamplerTaps sampler_taps(float gx, float gy, int h, int w, int bilinear) {
    SamplerTaps t;
    if (bilinear) {
        const float gxf = floorf(gx), gyf = floorf(gy);
        const float dx = __fsub_rn(gx, gxf), dy = __fsub_rn(gy, gyf);
        t.wa = __fmul_rn(__fsub_rn(1.f, dx), __fsub_rn(1.f, dy));
        t.wb = __fmul_rn(__fsub_rn(1.f, dx), dy);
        t.wc = __fmul_rn(dx, __fsub_rn(1.f, dy));
        t.wd = __fmul_rn(dx, dy);
        const int xi = static_cast<int>(fminf(fmaxf(gxf, 0.f), static_cast<float>(h - 1)));
        const int yi = static_cast<int>(fminf(fmaxf(gyf, 0.f), static_cast<float>(w - 1)));
        const int xb = min(max(xi + 1, 0), h - 1), yb = min(max(yi + 1, 0), w - 1);
        t.a = xi * w + yi; t.b = xi * w + yb; t.c = xb * w + yi; t.d = xb * w + yb;
    } else {
        const int xi = static_cast<int>(fminf(fmaxf(gx, 0.f), static_cast<float>(h - 1)));
        const int yi = static_cast<int>(fminf(fmaxf(gy, 0.f), static_cast<float>(w - 1)));
        t.a = t.b = t.c = t.d = xi * w + yi;
        t.wa = 1.f; t.wb = t.wc = t.wd = 0.f;
    }
    return t;
}

__global__ void __launch_bounds__(256) sampler_fwd_kernel(const SamplerMaps m, const float* __restrict__ idx, int n, int bilinear,
                                                          float* __restrict__ out, long long ld) {
    const int i = blockIdx.x;
    float gx = idx[2 * i], gy = idx[2 * i + 1];
    float* o = out + static_cast<long long>(i) * ld;
    for (int k = 0; k < m.nmaps; ++k) {
        if (m.div[k] != 1.f) { gx = __fdiv_rn(gx, m.div[k]); gy = __fdiv_rn(gy, m.div[k]); }
        const int c = m.c[k];
        const SamplerTaps t = sampler_taps(gx, gy, m.h[k], m.w[k], bilinear);
        const float* pa = m.ptr[k] + static_cast<long long>(t.a) * c;
        if (bilinear) {
            const float* pb = m.ptr[k] + static_cast<long long>(t.b) * c;
            const float* pc = m.ptr[k] + static_cast<long long>(t.c) * c;
            const float* pd = m.ptr[k] + static_cast<long long>(t.d) * c;
            for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
                float v = __fmul_rn(pa[ch], t.wa);
                v = __fadd_rn(v, __fmul_rn(pb[ch], t.wb));
                v = __fadd_rn(v, __fmul_rn(pc[ch], t.wc));
                v = __fadd_rn(v, __fmul_rn(pd[ch], t.wd));
                o[m.off[k] + ch] = v;
            }
        } else {
            for (int ch = threadIdx.x; ch < c; ch += blockDim.x) o[m.off[k] + ch] = pa[ch];
        }
    }
}

// Backward: scatter-add of grad_out through the same taps into the feature-map gradients (atomics).
__global__ void __launch_bounds__(256) sampler_bwd_kernel(const SamplerMaps m, const float* __restrict__ idx, int n, int bilinear,
                                                          const float* __restrict__ gout, long long ld) {
    const int i = blockIdx.x;
    float gx = idx[2 * i], gy = idx[2 * i + 1];
    const float* g = gout + static_cast<long long>(i) * ld;
    for (int k = 0; k < m.nmaps; ++k) {
        if (m.div[k] != 1.f) { gx = __fdiv_rn(gx, m.div[k]); gy = __fdiv_rn(gy, m.div[k]); }
        float* gm = m.gptr[k];
        if (!gm) continue;
        const int c = m.c[k];
        const SamplerTaps t = sampler_taps(gx, gy, m.h[k], m.w[k], bilinear);
        for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
            const float v = g[m.off[k] + ch];
            atomicAdd(gm + static_cast<long long>(t.a) * c + ch, v * t.wa);
            if (bilinear) {
                atomicAdd(gm + static_cast<long long>(t.b) * c + ch, v * t.wb);
                atomicAdd(gm + static_cast<long long>(t.c) * c + ch, v * t.wc);
                atomicAdd(gm + static_cast<long long>(t.d) * c + ch, v * t.wd);
            }
        }
    }
}

// --------------------------------------------------------------------------------------
// moment matching, mean part + covariance partial sum (single block):
//   l_mean = mean_d |mu_y - mu_x|,  gmu[d] = sign(mu_y - mu_x) / D
//   l_cov  = sum(part) / D^2 ;  l_m = l_cov + l_mean
// --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) moment_finish_kernel(const float* __restrict__ mu_y, const float* __restrict__ mu_x, int D,
                                                             const float* __restrict__ part, int npart,
                                                             float* __restrict__ gmu, float* __restrict__ scalars) {
    pdl_wait();
    pdl_trigger();
    __shared__ float sh[2][32];
    float a = 0.f, b = 0.f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        const float df = mu_y[d] - mu_x[d];
        a += fabsf(df);
        gmu[d] = ((df > 0.f) ? 1.f : ((df < 0.f) ? -1.f : 0.f)) / static_cast<float>(D);
    }
    for (int i = threadIdx.x; i < npart; i += blockDim.x) b += part[i];
    a = warp_sum(a); b = warp_sum(b);
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = a; sh[1][threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x < 32) {
        float aa = warp_sum(sh[0][threadIdx.x]), bb = warp_sum(sh[1][threadIdx.x]);
        if (threadIdx.x == 0) {
            const float lmean = aa / static_cast<float>(D);
            const float lcov = bb / (static_cast<float>(D) * static_cast<float>(D));
            scalars[S_LMEAN] = lmean; scalars[S_LCOV] = lcov; scalars[S_LM] = lmean + lcov;
        }
    }
}

// loss_s = l_m + l_remd + inv_alpha*l_pal ; total = (alpha*loss_c + loss_s)/denom   (run_strotss.py:40,140)
__global__ void combine_scalars_kernel(float* __restrict__ s, float alpha, float inv_alpha, float denom,
                                       const float* __restrict__ ss_loss_sum, float inv_n) {
    pdl_wait();
    pdl_trigger();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        if (ss_loss_sum) s[S_LOSS_C] = *ss_loss_sum * inv_n;
        const float ls = s[S_LM] + s[S_LREMD] + inv_alpha * s[S_LPAL];
        s[S_LOSS_S] = ls;
        s[S_TOTAL] = (alpha * s[S_LOSS_C] + ls) / denom;
    }
}

// --------------------------------------------------------------------------------------
// final gradient assembly, one block per prediction row i:
//   g^[d]  = w_ss * ( -ss2[i][d]/N + v[d] + coef_i * sumhat[d] )  +  w_remd * gremd[i][d]
//   grad   = (g^ - x^ (x^ . g^)) * inv_i                     (through l2_normalize; no projection if clamped)
//          + w_mom * ( q_scale * Q[i][d] + gmu[d]/N )
//          + w_pal * gpal[i][d]  for d < 3
// Any term whose pointer is null is skipped.  All weights already include 1/loss_denom.
// --------------------------------------------------------------------------------------
struct FinalizeArgs {
    const float* x; long long ldx; const float* inv; int N, D;
    int r0;                                   // first prediction row owned by this rank; local buffers start there
    const float* ss2; long long ld_ss2; const float* v; const float* coef; const float* sumhat; float w_ss;
    const float* gremd; long long ld_gremd; float w_remd;
    // relaxed-EMD gather branch (R_Y > R_X): g^_j = -(1/N) x^_{argmin_j}, read straight from the target rows
    const unsigned long long* remd_colbest; const float* remd_xs; long long remd_ldxs; const float* remd_inv_s;
    const float* scalars; int slot_branch;
    const float* Q; long long ldq; float q_scale; const float* gmu; float w_mom;
    const float* gpal; float w_pal;
    float* grad; long long ldg;
};

// The row is processed in batches of kFinU x 256 elements whose loads are all issued before the first use:
// with one load group per loop iteration the kernel is latency-bound (ncu: 80 % long-scoreboard stalls).
constexpr int kFinU = 9;                    // 9 x 256 = 2304 >= 2179: one batch per row at the reference's width

// Single-batch variant (D <= kFinU * 256, i.e. every width up to 2304 incl. the reference's 2179): each thread keeps its nine
// elements of the row in registers across the block reduction and fetches the moment-matching row Q together with the other
// operands, so a block pays one global-memory latency instead of two.
__global__ void __launch_bounds__(256) finalize_grad_1b_kernel(const FinalizeArgs a) {
    pdl_wait();
    __shared__ float sh[8];
    __shared__ float sh2[8];
    __shared__ float s_dot;
    const int li = blockIdx.x;
    const int i = a.r0 + li;
    const float* xr = a.x + static_cast<long long>(i) * a.ldx;
    const float iv = a.inv[i];
    const float invN = 1.f / static_cast<float>(a.N);
    const float ci = a.ss2 ? a.coef[i] : 0.f;
    const bool remd_gather = a.gremd && a.scalars[a.slot_branch] == 0.f;
    const float* g2 = nullptr; float g2s = 0.f;
    if (a.gremd) {
        if (remd_gather) {
            const int it = static_cast<int>(best_idx(a.remd_colbest[i]));
            g2 = a.remd_xs + static_cast<long long>(it) * a.remd_ldxs;
            g2s = -a.remd_inv_s[it] * invN * a.w_remd;
        } else {
            g2 = a.gremd + static_cast<long long>(li) * a.ld_gremd;
            g2s = a.w_remd;
        }
    }
    const float* s2 = a.ss2 ? a.ss2 + static_cast<long long>(li) * a.ld_ss2 : nullptr;
    const float* qr = a.Q ? a.Q + static_cast<long long>(li) * a.ldq : nullptr;
    const float wss = a.w_ss;
    float v_x[kFinU], v_g[kFinU], v_q[kFinU];
    {
        float v_s2[kFinU], v_g2[kFinU];
#pragma unroll
        for (int k = 0; k < kFinU; ++k) {
            const int d = threadIdx.x + 256 * k;
            const bool ok = d < a.D;
            v_x[k] = ok ? xr[d] : 0.f;
            v_s2[k] = (ok && s2) ? s2[d] : 0.f;
            v_g2[k] = (ok && g2) ? g2[d] : 0.f;
            v_q[k] = (ok && qr) ? qr[d] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < kFinU; ++k) {
            const int d = threadIdx.x + 256 * k;
            const bool ok = d < a.D;
            const float vv = (ok && s2) ? a.v[d] : 0.f;
            const float vs = (ok && s2) ? a.sumhat[d] : 0.f;
            v_g[k] = ok ? wss * (-v_s2[k] * invN + vv + ci * vs) + g2s * v_g2[k] : 0.f;
        }
    }
    float dot = 0.f, ssq = 0.f;
#pragma unroll
    for (int k = 0; k < kFinU; ++k) { dot = fmaf(v_g[k], v_x[k], dot); ssq = fmaf(v_x[k], v_x[k], ssq); }
    dot = warp_sum(dot); ssq = warp_sum(ssq);
    if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5] = dot; sh2[threadIdx.x >> 5] = ssq; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f, q = 0.f;
        for (int k = 0; k < 8; ++k) { t += sh[k]; q += sh2[k]; }
        s_dot = (q >= kL2NEps) ? t * iv * iv * iv : 0.f;
    }
    __syncthreads();
    const float pd = s_dot;
    float* gr = a.grad + static_cast<long long>(i) * a.ldg;
#pragma unroll
    for (int k = 0; k < kFinU; ++k) {
        const int d = threadIdx.x + 256 * k;
        if (d < a.D) {
            float o = v_g[k] * iv - v_x[k] * pd;
            if (qr) o += a.w_mom * (a.q_scale * v_q[k] + a.gmu[d] * invN);
            if (a.gpal && d < 3) o += a.w_pal * a.gpal[static_cast<long long>(li) * 4 + d];
            gr[d] = o;
        }
    }
}

__global__ void __launch_bounds__(256) finalize_grad_kernel(const FinalizeArgs a) {
    pdl_wait();
    extern __shared__ float sg[];       // D floats: g^
    __shared__ float sh[8];
    __shared__ float sh2[8];
    __shared__ float s_dot;
    const int li = blockIdx.x;                // local row (ss2, gremd, Q, gpal)
    const int i = a.r0 + li;                  // global row (x, inv, coef, grad)
    const float* xr = a.x + static_cast<long long>(i) * a.ldx;
    const float iv = a.inv[i];
    const float invN = 1.f / static_cast<float>(a.N);
    const float ci = a.ss2 ? a.coef[i] : 0.f;
    const bool remd_gather = a.gremd && a.scalars[a.slot_branch] == 0.f;
    // second gradient source of the relaxed EMD: either the scatter buffer row or the matched target row
    const float* g2 = nullptr; float g2s = 0.f;
    if (a.gremd) {
        if (remd_gather) {
            const int it = static_cast<int>(best_idx(a.remd_colbest[i]));
            g2 = a.remd_xs + static_cast<long long>(it) * a.remd_ldxs;
            g2s = -a.remd_inv_s[it] * invN * a.w_remd;
        } else {
            g2 = a.gremd + static_cast<long long>(li) * a.ld_gremd;
            g2s = a.w_remd;
        }
    }
    const float* s2 = a.ss2 ? a.ss2 + static_cast<long long>(li) * a.ld_ss2 : nullptr;
    const float wss = a.w_ss;
    float dot = 0.f, ssq = 0.f;
    for (int base = 0; base < a.D; base += kFinU * 256) {
        float v_s2[kFinU], v_g2[kFinU], v_x[kFinU], v_v[kFinU], v_sh[kFinU];
#pragma unroll
        for (int k = 0; k < kFinU; ++k) {
            const int d = base + threadIdx.x + 256 * k;
            const bool ok = d < a.D;
            v_x[k] = ok ? xr[d] : 0.f;
            v_s2[k] = (ok && s2) ? s2[d] : 0.f;
            v_v[k] = (ok && s2) ? a.v[d] : 0.f;
            v_sh[k] = (ok && s2) ? a.sumhat[d] : 0.f;
            v_g2[k] = (ok && g2) ? g2[d] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < kFinU; ++k) {
            const int d = base + threadIdx.x + 256 * k;
            if (d < a.D) {
                const float g = wss * (-v_s2[k] * invN + v_v[k] + ci * v_sh[k]) + g2s * v_g2[k];
                sg[d] = g;
                dot = fmaf(g, v_x[k], dot);
                ssq = fmaf(v_x[k], v_x[k], ssq);
            }
        }
    }
    dot = warp_sum(dot); ssq = warp_sum(ssq);
    if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5] = dot; sh2[threadIdx.x >> 5] = ssq; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f, q = 0.f;
        for (int k = 0; k < 8; ++k) { t += sh[k]; q += sh2[k]; }
        // g^.x^ = (g^.x) * inv ; the projection term is x^ (x^.g^) * inv = x * (g^.x) * inv^3
        s_dot = (q >= kL2NEps) ? t * iv * iv * iv : 0.f;
    }
    __syncthreads();
    const float pd = s_dot;
    float* gr = a.grad + static_cast<long long>(i) * a.ldg;
    const float* qr = a.Q ? a.Q + static_cast<long long>(li) * a.ldq : nullptr;
    for (int base = 0; base < a.D; base += kFinU * 256) {
        float v_q[kFinU], v_x[kFinU], v_m[kFinU];
#pragma unroll
        for (int k = 0; k < kFinU; ++k) {
            const int d = base + threadIdx.x + 256 * k;
            const bool ok = d < a.D;
            v_x[k] = ok ? xr[d] : 0.f;
            v_q[k] = (ok && qr) ? qr[d] : 0.f;
            v_m[k] = (ok && qr) ? a.gmu[d] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < kFinU; ++k) {
            const int d = base + threadIdx.x + 256 * k;
            if (d < a.D) {
                float o = sg[d] * iv - v_x[k] * pd;
                if (qr) o += a.w_mom * (a.q_scale * v_q[k] + v_m[k] * invN);
                if (a.gpal && d < 3) o += a.w_pal * a.gpal[static_cast<long long>(li) * 4 + d];
                gr[d] = o;
            }
        }
    }
}

}  // namespace sb
