// Persistent, warp-specialised tcgen05 GEMM core for sm_100a:  acc[a] (+)= A_seg * B_seg^T
//
//   * operands are bf16, K-major (row-major with K contiguous), staged by TMA with 128-byte swizzle
//   * a launch is a list of up to three "segments"; each segment is one (A, B) tensor-map pair with
//     its own K extent and the index of the TMEM accumulator it adds into (that is how the
//     self-similarity kernel forms  delta.x^T + y.delta^T  in one accumulator and  y.y^T  in a second)
//   * tile = 128 (rows of A) x BN (rows of B), fp32 accumulators in TMEM, double-buffered when
//     2 * BN * NACC <= 512 columns so the epilogue of tile t overlaps the MMAs of tile t+1
//   * warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warp 2 = TMEM allocator,
//     warps 4..7 = epilogue (warp q owns TMEM lanes 32q..32q+31, one accumulator row per thread)
//   * the epilogue is a policy class; no tile of the product is written to HBM unless the policy
//     stores it.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace sb {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kNonEpiThreads = 128;           // warps 0..3: TMA, MMA, TMEM alloc, spare
constexpr int kMaxSeg = 10;          // the row-sharded stage 2 sums up to world + 2 operand blocks into one accumulator
constexpr int kSmemBudget = 232448 - 1024;   // 227 KB minus alignment slack

template <int BN, int NACC>
struct TileCfg {
    static constexpr int ACC_COLS = BN * NACC;
    static constexpr int ACC_STAGES = (2 * ACC_COLS <= 512) ? 2 : 1;
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static_assert(ACC_COLS <= 512, "accumulators exceed TMEM");
};

template <class Epi>
struct GemmParams {
    CUtensorMap tmA[kMaxSeg];
    CUtensorMap tmB[kMaxSeg];
    int nseg;
    int seg_kblocks[kMaxSeg];
    int seg_acc[kMaxSeg];
    int tiles_m, tiles_n;
    int group_n;                 // raster: column tiles are swept in groups of group_n (L2 blocking)
    int tri;                     // 1: square tile grid of a symmetric product, only tiles with tn >= tm are visited
    int a_row0, b_row0;          // element row where tile (0,0) starts in A / B
    // Optional tile-dependent K range per segment (CTA-pair kernel only; all zero = full range): for column tile tn, segment s
    // runs K blocks [kb_lo_mul[s] * tn, kb_hi_mul[s] ? min(seg_kblocks[s], kb_hi_mul[s] * tn) : seg_kblocks[s]).
    // Stage 2 of the self-similarity uses it to walk only the upper block triangle of the symmetric P panel.
    int kb_lo_mul[kMaxSeg], kb_hi_mul[kMaxSeg];
    // B_MODE == 2 of the pair kernel: segment s reads its B operand MN-major (transposed) iff seg_bmn[s]
    int seg_bmn[kMaxSeg];
    // CTA-pair kernel: number of 16-wide MMA steps that carry data in the LAST K block of a segment (0 = all four).  K is
    // zero-padded to a multiple of 64 (D = 2179 -> 2240): the last block holds 3 real columns, so 3 of its 4 steps multiply zeros.
    int k_tail_steps;
    // gemm2s_kernel: K blocks by which the two 256-wide halves of a tile couple are skewed against each other
    int skew;
    typename Epi::Params epi;
};

struct TileInfo {
    int tm, tn;                  // tile indices
    int row0, col0;              // a_row0 + tm*BM, b_row0 + tn*BN  (global element coordinates)
    int q, lane;                 // TMEM lane quadrant (= warp % 4) and lane
    int w, nw;                   // epilogue warp index and count (4 or 8)
    int csplit, nsplit;          // with 8 epilogue warps the BN columns are split in two halves
    int c0, c1;                  // range of 32-column chunks this warp handles
    int tid;                     // thread index within the epilogue group
    uint32_t taddr;              // TMEM address of accumulator 0, this warp's lane quadrant
    int tile_seq;                // running count of tiles processed by this CTA
};

// Tile order: for each group of group_n column tiles, sweep all row tiles, column tile fastest.
// CTAs that run concurrently then share a few A row blocks and one group of B column blocks, so
// both stay L2-resident instead of B being re-streamed from HBM for every row block.
template <class P>
__host__ __device__ __forceinline__ int num_tiles_of(const P& p) {
    return p.tri ? p.tiles_n * (p.tiles_n + 1) / 2 : p.tiles_m * p.tiles_n;
}

template <class P>
__host__ __device__ __forceinline__ void decode_tile(const P& p, int t, int& tm, int& tn) {
    if (p.tri) {                 // row-major walk over the upper triangle (tiles_m == tiles_n)
        int row = 0, len = p.tiles_n;
        while (t >= len) { t -= len; ++row; --len; }
        tm = row; tn = row + t;
        return;
    }
    const int gsz = p.tiles_m * p.group_n;
    const int tg = t / gsz;
    const int r = t - tg * gsz;
    const int gn = min(p.group_n, p.tiles_n - tg * p.group_n);
    tm = r / gn;
    tn = tg * p.group_n + (r - tm * gn);
}

template <int NT>
__device__ __forceinline__ void epi_bar_sync() {
    asm volatile("bar.sync 1, %0;" :: "n"(NT) : "memory");
}

// A_MN: the A operand is read from a matrix stored K x M (M contiguous), i.e. A = (stored)^T, through
// MN-major shared-memory descriptors; its tensor map has dims {M, K} and box {64, 64}.
template <int BN, int NACC, int STAGES, int EPI_WARPS, class Epi, bool A_MN = false>
__global__ void __launch_bounds__(kNonEpiThreads + 32 * EPI_WARPS, 1)
gemm_kernel(const __grid_constant__ GemmParams<Epi> p) {
    static_assert(EPI_WARPS == 4 || EPI_WARPS == 8, "4 or 8 epilogue warps");
    using Cfg = TileCfg<BN, NACC>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* epi_smem = smem + STAGES * Cfg::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + Epi::SMEM_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tfull = bars + 2 * STAGES;
    uint64_t* tempty = tfull + Cfg::ACC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + Cfg::ACC_STAGES);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.nseg; ++s) { tma_prefetch_desc(&p.tmA[s]); tma_prefetch_desc(&p.tmB[s]); }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < Cfg::ACC_STAGES; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();                      // persistent single-wave grid: dependents may be scheduled as CTAs exit
    pdl_wait();                         // the previous kernel of the stream has completed; its writes are visible

    const int num_tiles = num_tiles_of(p);

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                int tm, tn;
                decode_tile(p, t, tm, tn);
                const int arow = p.a_row0 + tm * BM, brow = p.b_row0 + tn * BN;
                for (int s = 0; s < p.nseg; ++s) {
                    for (int kb = 0; kb < p.seg_kblocks[s]; ++kb) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
                        uint8_t* sB = sA + Cfg::A_BYTES;
                        mbar_arrive_expect_tx(&full[stage], Cfg::STAGE_BYTES);
                        if constexpr (A_MN) {
                            tma_load_2d(sA, &p.tmA[s], &full[stage], arow, kb * BK);
                            tma_load_2d(sA + Cfg::A_BYTES / 2, &p.tmA[s], &full[stage], arow + 64, kb * BK);
                        } else {
                            tma_load_2d(sA, &p.tmA[s], &full[stage], kb * BK, arow);
                        }
                        tma_load_2d(sB, &p.tmB[s], &full[stage], kb * BK, brow);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN);
            int stage = 0; uint32_t phase = 0;
            int as = 0; uint32_t aphase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                mbar_wait(&tempty[as], aphase ^ 1);
                tc_fence_after();
                uint32_t touched = 0;
                for (int s = 0; s < p.nseg; ++s) {
                    const int acc = p.seg_acc[s];
                    const uint32_t d_addr = tmem_base + as * Cfg::ACC_COLS + acc * BN;
                    for (int kb = 0; kb < p.seg_kblocks[s]; ++kb) {
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                        const uint64_t adesc = A_MN ? make_mnmajor_sw128_desc(a_addr, Cfg::A_BYTES / 2)
                                                    : make_kmajor_sw128_desc(a_addr);
                        const uint64_t bdesc = make_kmajor_sw128_desc(a_addr + Cfg::A_BYTES);
                        // K-major: 16 bf16 = 32 bytes along K inside the swizzle atom; MN-major: 16 k-rows = 2048 bytes
                        constexpr uint32_t a_step = A_MN ? (16 * 128) >> 4 : 2;
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            umma_bf16(d_addr, adesc + a_step * k, bdesc + 2 * k, idesc,
                                      ((touched >> acc) & 1u) | (k > 0 ? 1u : 0u));
                        }
                        touched |= (1u << acc);
                        umma_commit(&empty[stage]);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
                umma_commit(&tfull[as]);
                if (++as == Cfg::ACC_STAGES) { as = 0; aphase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        int as = 0; uint32_t aphase = 0;
        int seq = 0;
        typename Epi::State st;
        Epi::init(st, p.epi, q, lane);
        constexpr int kSplit = EPI_WARPS / 4;
        constexpr int kChunks = BN / 32 / kSplit;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++seq) {
            TileInfo ti;
            decode_tile(p, t, ti.tm, ti.tn);
            ti.row0 = p.a_row0 + ti.tm * BM; ti.col0 = p.b_row0 + ti.tn * BN;
            ti.q = q; ti.lane = lane; ti.tile_seq = seq;
            ti.w = warp - 4; ti.nw = EPI_WARPS; ti.csplit = (warp - 4) >> 2; ti.nsplit = kSplit;
            ti.c0 = ti.csplit * kChunks; ti.c1 = ti.c0 + kChunks;
            ti.tid = threadIdx.x - kNonEpiThreads;
            ti.taddr = tmem_base + as * Cfg::ACC_COLS + (static_cast<uint32_t>(q * 32) << 16);
            Epi::prologue(st, p.epi, ti, epi_smem);      // may overlap the MMAs of this tile
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            Epi::run(st, p.epi, ti, epi_smem);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[as]);
            if (++as == Cfg::ACC_STAGES) { as = 0; aphase ^= 1; }
        }
        Epi::finish(st, p.epi, q, lane);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ======================================================================================
// Epilogue policies
// ======================================================================================

// ---- plain store:  C[row][col] = alpha * acc   (fp32, row stride ldc) -----------------
struct EpiStore {
    static constexpr int SMEM_BYTES = 0;
    struct Params { float* C; long long ldc; int rows, cols; float alpha; int row_off; int accumulate; };
    struct State {};
    __device__ static void init(State&, const Params&, int, int) {}
    __device__ static void prologue(State&, const Params&, const TileInfo&, uint8_t*) {}
    __device__ static void finish(State&, const Params&, int, int) {}
    template <int BN>
    __device__ static void run_bn(State&, const Params& P, const TileInfo& ti, uint8_t*) {
        const int row = ti.row0 + ti.q * 32 + ti.lane;
        const bool rvalid = row < P.rows;
        float* crow = P.C + static_cast<long long>(row - P.row_off) * P.ldc;
#pragma unroll 1
        for (int c = ti.c0; c < ti.c1; ++c) {
            uint32_t r[32];
            tmem_ld32(ti.taddr + c * 32, r);
            tmem_ld_wait();
            const int col = ti.col0 + c * 32;
            if (rvalid) {
                if (col + 32 <= P.cols && (P.ldc & 3) == 0) {
#pragma unroll
                    float4 o[8];
                    if (P.accumulate) {          // loads first, then stores (see EpiStoreTr)
#pragma unroll
                        for (int e = 0; e < 8; ++e) o[e] = *reinterpret_cast<const float4*>(crow + col + 4 * e);
                    }
#pragma unroll
                    for (int e = 0; e < 32; e += 4) {
                        float4 v = make_float4(__uint_as_float(r[e]) * P.alpha, __uint_as_float(r[e + 1]) * P.alpha,
                                               __uint_as_float(r[e + 2]) * P.alpha, __uint_as_float(r[e + 3]) * P.alpha);
                        if (P.accumulate) { v.x += o[e >> 2].x; v.y += o[e >> 2].y; v.z += o[e >> 2].z; v.w += o[e >> 2].w; }
                        *reinterpret_cast<float4*>(crow + col + e) = v;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 32; ++e)
                        if (col + e < P.cols)
                            crow[col + e] = __uint_as_float(r[e]) * P.alpha + (P.accumulate ? crow[col + e] : 0.f);
                }
            }
        }
    }
};
template <int BN> struct EpiStoreT : EpiStore {
    __device__ static void run(State& s, const Params& P, const TileInfo& ti, uint8_t* sm) { run_bn<BN>(s, P, ti, sm); }
};

// ---- transposed store:  C[col][row] (+)= alpha * acc  (fp32, row stride ldc along the B index) ----------
// Used with the operand roles swapped (A = the matrix whose rows are output COLUMNS): thread = accumulator row,
// so for a fixed register (one B row = one output row) the 32 lanes of a warp write 32 consecutive floats --
// a coalesced 128-byte line per instruction instead of 32 lines.
template <int BN>
struct EpiStoreTr {
    static constexpr int SMEM_BYTES = 0;
    // accumulate: add to C everywhere; acc_cols_below: add to C for output rows (B index) below this bound only
    struct Params { float* C; long long ldc; int rows, cols; float alpha; int col_off; int accumulate; int acc_cols_below; };
    struct State {};
    __device__ static void init(State&, const Params&, int, int) {}
    __device__ static void prologue(State&, const Params&, const TileInfo&, uint8_t*) {}
    __device__ static void finish(State&, const Params&, int, int) {}
    __device__ static void run(State&, const Params& P, const TileInfo& ti, uint8_t*) {
        const int row = ti.row0 + ti.q * 32 + ti.lane;            // index along the contiguous output dimension
        const bool rvalid = row < P.rows;
        const bool accumulate = P.accumulate || (ti.col0 < P.acc_cols_below);      // tile-uniform (bounds are tile-aligned)
#pragma unroll 1
        for (int c = ti.c0; c < ti.c1; ++c) {
            uint32_t r[32];
            tmem_ld32(ti.taddr + c * 32, r);
            tmem_ld_wait();
            const int colbase = ti.col0 + c * 32;
            if (rvalid && colbase < P.cols) {
                float* dst = P.C + static_cast<long long>(colbase - P.col_off) * P.ldc + row;
                if (accumulate) {
                    // all 32 loads are issued before the first store: a load-add-store per element would be
                    // serialised by the compiler (it cannot prove the 32 addresses distinct) -- 32 round trips
                    float old[32];
#pragma unroll
                    for (int e = 0; e < 32; ++e) old[e] = (colbase + e < P.cols) ? dst[e * P.ldc] : 0.f;
#pragma unroll
                    for (int e = 0; e < 32; ++e)
                        if (colbase + e < P.cols) dst[e * P.ldc] = old[e] + __uint_as_float(r[e]) * P.alpha;
                } else {
#pragma unroll
                    for (int e = 0; e < 32; ++e)
                        if (colbase + e < P.cols) dst[e * P.ldc] = __uint_as_float(r[e]) * P.alpha;
                }
            }
        }
    }
};

// ---- partial Gram tile -> the window slot of the rank that owns the tile (row-sharded covariance) ----------
// Symmetric product over THIS rank's rows only (triangular tile walk): tile number lt of the upper block triangle belongs to
// rank lt % world and is stored whole -- 256 x 256 fp32, transposed, phantom rows/columns are zeros -- as slot lt / world of
// this sender's slot array in the owner's peer window (base[owner]; plain stores to a CUDA-IPC mapped pointer, NVLink for a
// remote owner).
template <int BN>
struct EpiGramScatter {
    static constexpr int SMEM_BYTES = 0;
    static constexpr int kMaxRanks = 16;
    struct Params { float* base[kMaxRanks]; int world; int tiles; };      // tiles: 256-wide tiles per side of the matrix
    struct State {};
    __device__ static void init(State&, const Params&, int, int) {}
    __device__ static void prologue(State&, const Params&, const TileInfo&, uint8_t*) {}
    __device__ static void finish(State&, const Params&, int, int) {}
    __device__ static void run(State&, const Params& P, const TileInfo& ti, uint8_t*) {
        static_assert(BN == 256, "256 x 256 tiles");
        const int tm2 = ti.tm >> 1;                                    // ti.tm counts 128-row blocks
        const int lt = tm2 * P.tiles - tm2 * (tm2 - 1) / 2 + (ti.tn - tm2);
        const int owner = lt % P.world, slot = lt / P.world;
        // the slot holds the tile TRANSPOSED (slot[col][row]; the matrix is symmetric, so that is its mirror tile): for a fixed
        // accumulator column the 32 lanes of a warp -- 32 consecutive rows -- then write one full 128-byte line, which is what a
        // remote (NVLink) store wants; row-major slots would send 32 partial sectors per instruction
        float* dst = P.base[owner] + static_cast<long long>(slot) * (256 * 256) + ((ti.tm & 1) * 128 + ti.q * 32 + ti.lane);
#pragma unroll 1
        for (int c = ti.c0; c < ti.c1; ++c) {
            uint32_t r[32];
            tmem_ld32(ti.taddr + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) dst[(c * 32 + e) * 256] = __uint_as_float(r[e]);
        }
    }
};

// ---- relaxed EMD: best (max dot = min cosine distance) per A row and per B row ---------
// A rows = target/style samples i (M of them), B rows = prediction samples j (N of them).
// rowbest[i] = max_j (dot_ij, lowest j on ties);  colbest[j] = max_i (dot_ij, lowest i on ties).
// Phantom rows/columns produced by TMA zero fill are masked out (SURVEY "OOB / ragged tiles").
template <int BN>
struct EpiRemd {
    static constexpr int SMEM_BYTES = 0;
    struct Params { unsigned long long* rowbest; unsigned long long* colbest; int M, N; };
    struct State {};
    __device__ static void init(State&, const Params&, int, int) {}
    __device__ static void prologue(State&, const Params&, const TileInfo&, uint8_t*) {}
    __device__ static void finish(State&, const Params&, int, int) {}
    __device__ static void run(State&, const Params& P, const TileInfo& ti, uint8_t*) {
        const int row = ti.row0 + ti.q * 32 + ti.lane;
        const bool rvalid = row < P.M;
        float best_v = -INFINITY;
        int best_j = 0;
#pragma unroll 1
        for (int c = ti.c0; c < ti.c1; ++c) {
            uint32_t r[32];
            tmem_ld32(ti.taddr + c * 32, r);
            tmem_ld_wait();
            const int colbase = ti.col0 + c * 32;
            uint32_t my_key = 0; int my_src = 0;
#pragma unroll
            for (int e = 0; e < 32; ++e) {
                const float v = __uint_as_float(r[e]);
                const int col = colbase + e;
                if (col < P.N && v > best_v) { best_v = v; best_j = col; }
                const uint32_t key = rvalid ? f2ord(v) : 0u;
                const uint32_t m = __reduce_max_sync(0xffffffffu, key);
                const uint32_t ball = __ballot_sync(0xffffffffu, key == m);
                if (ti.lane == e) { my_key = m; my_src = __ffs(ball) - 1; }
            }
            const int mycol = colbase + ti.lane;
            if (mycol < P.N && my_key != 0u) {
                const uint32_t src_row = static_cast<uint32_t>(ti.row0 + ti.q * 32 + my_src);
                const unsigned long long packed = (static_cast<unsigned long long>(my_key) << 32) |
                                                  static_cast<unsigned long long>(~src_row);
                atomicMax(P.colbest + mycol, packed);
            }
        }
        if (rvalid && best_v > -INFINITY) atomicMax(P.rowbest + row, pack_best(best_v, static_cast<uint32_t>(best_j)));
    }
};

// ---- self-similarity, stage 1 ----------------------------------------------------------
// acc0 = delta_i.x^_j + y^_i.delta_j = -(Xd_ij - Yd_ij)      (delta = x^ - y^; x^ pred, y^ content)
// acc1 = y^_i.y^_j                   = 1 - Yd_ij
// column form  term_ij  = Xd_ij/s_j - Yd_ij/t_j = diff*u_j + Yd*w_j      (u = 1/s, w = 1/s - 1/t)
// row form     term'_ij = term_ji               = diff*u_i + Yd*w_i      (Xd, Yd symmetric)
// loss_i += |term'_ij|,  r_i += sign(term'_ij) * Xd_ij,  P_ij = sign(term_ij)*u_j + sign(term'_ij)*u_i
// P (bf16) goes to the L2-resident row panel that stage 2 multiplies with x^.
template <int BN_, int EPI_WARPS>
struct EpiSS1 {
    static constexpr int BN = BN_;
    static constexpr int SMEM_BYTES = 2 * 2 * BN * sizeof(float);
    struct Params {
        const float* u; const float* w;       // per sample, length N
        __nv_bfloat16* P; long long ldp;      // panel, row stride (elements), panel starts at row panel_row0
        int panel_row0;
        float* loss_part; float* r_part;      // [tiles_n * nsplit][N]
        int N;
        int row_end;                          // rows >= row_end belong to another rank (or do not exist)
        int write_p;
        // symmetric mode (single GPU): only column tiles >= the panel's first row are computed; tiles
        // strictly right of the panel (col0 >= panel_end) also account for their mirror images:
        int sym, panel_end;
        float* rcol_part;                     // [row_blocks * 4][N]: per-warp column sums of sign_col * Xd
        int p_col0;                           // column of the matrix stored in column 0 of P (row-sharded jobs keep one
                                              // contiguous [rows][cols] block per column range; 0 for a full-width panel)
    };
    struct State {};
    __device__ static void init(State&, const Params&, int, int) {}
    __device__ static void finish(State&, const Params&, int, int) {}
    __device__ static void prologue(State&, const Params& P, const TileInfo& ti, uint8_t* sm) {
        float* buf = reinterpret_cast<float*>(sm) + (ti.tile_seq & 1) * 2 * BN;
        for (int t = ti.tid; t < BN; t += 32 * EPI_WARPS) {
            const int col = ti.col0 + t;
            buf[t] = (col < P.N) ? P.u[col] : 0.f;
            buf[BN + t] = (col < P.N) ? P.w[col] : 0.f;
        }
        epi_bar_sync<32 * EPI_WARPS>();
    }
    __device__ static void run(State&, const Params& P, const TileInfo& ti, uint8_t* sm) {
        const float* su = reinterpret_cast<const float*>(sm) + (ti.tile_seq & 1) * 2 * BN;
        const float* sw = su + BN;
        const int row = ti.row0 + ti.q * 32 + ti.lane;
        const bool rvalid = row < P.row_end;
        const float ui = rvalid ? P.u[row] : 0.f;
        const float wi = rvalid ? P.w[row] : 0.f;
        float loss = 0.f, racc = 0.f;
        const bool both = P.sym && (ti.col0 >= P.panel_end);
        __nv_bfloat16* prow = P.P + static_cast<long long>(row - P.panel_row0) * P.ldp;
#pragma unroll 1
        for (int c = ti.c0; c < ti.c1; ++c) {
            uint32_t a0[32], a1[32];
            tmem_ld32(ti.taddr + c * 32, a0);
            tmem_ld32(ti.taddr + BN + c * 32, a1);
            tmem_ld_wait();
            const int colbase = ti.col0 + c * 32;
            uint32_t packed[16];
            float cv[32];                     // sign_col * Xd, reduced over the warp's rows when `both`
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
                float pv[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int col = colbase + e + h;
                    const bool live = rvalid && (col < P.N) && (col != row);
                    const float diff = -__uint_as_float(a0[e + h]);
                    const float yd = 1.f - __uint_as_float(a1[e + h]);
                    const float uj = su[c * 32 + e + h], wj = sw[c * 32 + e + h];
                    const float tc = fmaf(diff, uj, yd * wj);
                    const float tr = fmaf(diff, ui, yd * wi);
                    const float sc = live ? ((tc > 0.f) ? 1.f : ((tc < 0.f) ? -1.f : 0.f)) : 0.f;
                    const float sr = live ? ((tr > 0.f) ? 1.f : ((tr < 0.f) ? -1.f : 0.f)) : 0.f;
                    loss += live ? fabsf(tr) : 0.f;
                    racc = fmaf(sr, yd + diff, racc);
                    if (both) { loss += live ? fabsf(tc) : 0.f; cv[e + h] = sc * (yd + diff); }
                    pv[h] = fmaf(sc, uj, sr * ui);
                }
                packed[e >> 1] = pack_bf16x2(pv[0], pv[1]);
            }
            if (P.write_p && rvalid && colbase - P.p_col0 < P.ldp) {
                // ldp is a multiple of 64 and colbase a multiple of 32, so a 32-wide chunk is
                // either fully inside the padded row or fully outside it.
                uint4* dst = reinterpret_cast<uint4*>(prow + (colbase - P.p_col0));
#pragma unroll
                for (int v = 0; v < 4; ++v)
                    dst[v] = make_uint4(packed[4 * v], packed[4 * v + 1], packed[4 * v + 2], packed[4 * v + 3]);
            }
            if (both) {
                // transpose-reduce: after 5 exchange steps lane l holds the sum over the warp's 32 rows of column l
#pragma unroll
                for (int sft = 16; sft >= 1; sft >>= 1) {
                    const bool up = (ti.lane & sft) != 0;
#pragma unroll
                    for (int e = 0; e < sft; ++e) {
                        const float send = up ? cv[e] : cv[e + sft];
                        const float keep = up ? cv[e + sft] : cv[e];
                        cv[e] = keep + __shfl_xor_sync(0xffffffffu, send, sft);
                    }
                }
                const int col = colbase + ti.lane;
                if (col < P.N)
                    P.rcol_part[(static_cast<long long>(ti.row0 / BM) * 4 + ti.q) * P.N + col] = cv[0];
            }
        }
        if (rvalid) {
            const long long slot = static_cast<long long>(ti.col0 / BN) * ti.nsplit + ti.csplit;
            P.loss_part[slot * P.N + row] = loss;
            P.r_part[slot * P.N + row] = racc;
        }
    }
};

// ---- covariance forward:  V = acc/n ; loss partial = sum |V - Vx| ; Sg = sign(V - Vx) (bf16) ---
template <int BN>
struct EpiCovFwd {
    static constexpr int SMEM_BYTES = 0;
    struct Params {
        const float* Vx; long long ldv;        // target covariance (fp32)
        __nv_bfloat16* Sg; long long lds;      // sign matrix out (bf16, zero-padded by the owner)
        float* part;                           // [num_tiles * 4] partial sums of |V - Vx|
        float inv_n; int D;
        int tiles_n;
        int sym;                               // tiles right of the diagonal also stand for their mirror images
        int diag_cols;                         // width of a diagonal block in columns (256)
    };
    struct State {};
    __device__ static void init(State&, const Params&, int, int) {}
    __device__ static void prologue(State&, const Params&, const TileInfo&, uint8_t*) {}
    __device__ static void finish(State&, const Params&, int, int) {}
    __device__ static void run(State&, const Params& P, const TileInfo& ti, uint8_t*) {
        const int row = ti.row0 + ti.q * 32 + ti.lane;
        const bool rvalid = row < P.D;
        const float* vrow = P.Vx + static_cast<long long>(row) * P.ldv;
        __nv_bfloat16* srow = P.Sg + static_cast<long long>(row) * P.lds;
        // symmetric mode: a tile strictly right of the diagonal block counts twice and also writes Sg^T
        const bool mirror = P.sym && (ti.col0 >= (ti.row0 / P.diag_cols + 1) * P.diag_cols);
        float part = 0.f;
#pragma unroll 1
        for (int c = ti.c0; c < ti.c1; ++c) {
            uint32_t r[32];
            tmem_ld32(ti.taddr + c * 32, r);
            tmem_ld_wait();
            const int colbase = ti.col0 + c * 32;
            if (rvalid && colbase < P.D) {
                if (colbase + 32 <= P.D) {
                    uint32_t packed[16];
#pragma unroll
                    for (int e = 0; e < 32; e += 4) {
                        const float4 vx = *reinterpret_cast<const float4*>(vrow + colbase + e);
                        const float d0 = fmaf(__uint_as_float(r[e]), P.inv_n, -vx.x);
                        const float d1 = fmaf(__uint_as_float(r[e + 1]), P.inv_n, -vx.y);
                        const float d2 = fmaf(__uint_as_float(r[e + 2]), P.inv_n, -vx.z);
                        const float d3 = fmaf(__uint_as_float(r[e + 3]), P.inv_n, -vx.w);
                        part += fabsf(d0) + fabsf(d1) + fabsf(d2) + fabsf(d3);
                        auto sg = [](float d) { return (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f); };
                        packed[e >> 1] = pack_bf16x2(sg(d0), sg(d1));
                        packed[(e >> 1) + 1] = pack_bf16x2(sg(d2), sg(d3));
                        if (mirror) {          // Sg[col][row]: 32 lanes write 32 consecutive bf16 of one row
                            __nv_bfloat16* t = P.Sg + static_cast<long long>(colbase + e) * P.lds + row;
                            t[0] = __float2bfloat16(sg(d0)); t[P.lds] = __float2bfloat16(sg(d1));
                            t[2 * P.lds] = __float2bfloat16(sg(d2)); t[3 * P.lds] = __float2bfloat16(sg(d3));
                        }
                    }
                    uint4* dst = reinterpret_cast<uint4*>(srow + colbase);
#pragma unroll
                    for (int v = 0; v < 4; ++v)
                        dst[v] = make_uint4(packed[4 * v], packed[4 * v + 1], packed[4 * v + 2], packed[4 * v + 3]);
                } else {
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        if (colbase + e < P.D) {
                            const float d = fmaf(__uint_as_float(r[e]), P.inv_n, -vrow[colbase + e]);
                            part += fabsf(d);
                            const __nv_bfloat16 sv = __float2bfloat16((d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f));
                            srow[colbase + e] = sv;
                            if (mirror) P.Sg[static_cast<long long>(colbase + e) * P.lds + row] = sv;
                        }
                    }
                }
            }
        }
        part = warp_sum(part);
        if (mirror) part *= 2.f;
        if (ti.lane == 0) P.part[(static_cast<long long>(ti.tm) * P.tiles_n + ti.tn) * ti.nw + ti.w] = part;
    }
};

}  // namespace sb
