// CTA-pair (cta_group::2) variant of the GEMM core: a cluster of two CTAs (one SM pair) owns one
// 256 x 256 output tile.  Each CTA stages its own 128 rows of A and its own 128 rows of B per K block
// (32 KB instead of 48 KB), the leader CTA issues 256 x 256 x 16 MMAs that read both CTAs' shared
// memory, and each CTA's TMEM receives its 128 accumulator rows.
//
// Why: with one CTA per tile a K block moves 48 KB into shared memory (TMA) and 48 KB out of it (MMA
// operand reads) per 512 tensor cycles = 192 B/clk against the SM's 128 B/clk shared-memory bandwidth,
// which caps the tensor pipe at ~68 % (measured, profiles/).  The pair halves the B traffic per SM:
// 32 KB in + 32 KB out per 512 cycles = 128 B/clk.
//
// Protocol (per pipeline stage / accumulator stage):
//   full[s]   lives in the LEADER: its producer arms it with the bytes of BOTH CTAs; both CTAs' TMA
//             loads credit it (cp.async.bulk.tensor ... cta_group::2 with the leader's barrier address)
//   empty[s]  one per CTA: tcgen05.commit.cta_group::2 multicast arrives in both after the MMAs that
//             read the stage completed
//   tfull[a]  one per CTA, same multicast commit after the last K block of a tile
//   tempty[a] lives in the LEADER: the epilogue warps of both CTAs arrive on it (remote arrive via mapa)
#pragma once
#include "gemm_core.cuh"

namespace sb {

constexpr int BM2 = 256;             // rows of A per CTA pair

template <int NACC>
struct PairCfg {
    static constexpr int BN = 256;
    static constexpr int ACC_COLS = BN * NACC;
    static constexpr int ACC_STAGES = (2 * ACC_COLS <= 512) ? 2 : 1;
    static constexpr int A_BYTES = 128 * BK * 2;
    static constexpr int B_BYTES = 128 * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;     // per CTA
};

// tiles_m of GemmParams counts 256-row tiles here.
// B_MODE 1: the B operand is read from a matrix stored K x N (N contiguous) through MN-major descriptors; its tensor
// map has dims {N, K} and box {64, 64} (each CTA stages its 128 columns of B as two 64-wide chunks).
// B_MODE 2: chosen per segment at run time (p.seg_bmn[s]); both layouts stage the same 16 KB per CTA and K block.
// A_MODE 1: the A operand of EVERY segment is read from a matrix stored K x M (M contiguous) the same way (tensor map dims
// {M, K}, box {64, 64}, each CTA stages its 128 rows of A as two 64-wide chunks): the row-major bf16 operands x^ / cen
// then serve as the "transposed" operands of stage 2 and of the covariance without a transposed copy in HBM.
template <int NACC, int STAGES, int EPI_WARPS, class Epi, int B_MODE = 0, int A_MODE = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kNonEpiThreads + 32 * EPI_WARPS, 1)
gemm2_kernel(const __grid_constant__ GemmParams<Epi> p) {
    static_assert(EPI_WARPS == 4 || EPI_WARPS == 8, "4 or 8 epilogue warps");
    using Cfg = PairCfg<NACC>;
    constexpr int BN = Cfg::BN;
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
    const uint32_t rank = cluster_ctarank();
    const bool leader = (rank == 0);
    const int pair = blockIdx.x >> 1;
    const int npairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.nseg; ++s) { tma_prefetch_desc(&p.tmA[s]); tma_prefetch_desc(&p.tmB[s]); }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < Cfg::ACC_STAGES; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 2 * EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc2(tmem_slot, 512); tmem_relinquish2(); }
    tc_fence_before();
    cluster_sync_all();                 // barriers of BOTH CTAs are initialised before anyone signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();                      // persistent single-wave grid: dependents may be scheduled as CTAs exit
    pdl_wait();                         // the previous kernel of the stream has completed; its writes are visible

    const int num_tiles = num_tiles_of(p);

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int t = pair; t < num_tiles; t += npairs) {
                int tm, tn;
                decode_tile(p, t, tm, tn);
                const int arow = p.a_row0 + tm * BM2 + static_cast<int>(rank) * 128;
                const int brow = p.b_row0 + tn * BN + static_cast<int>(rank) * 128;
                for (int s = 0; s < p.nseg; ++s) {
                    const int kb0 = p.kb_lo_mul[s] * tn;
                    const int kb1 = p.kb_hi_mul[s] ? min(p.seg_kblocks[s], p.kb_hi_mul[s] * tn) : p.seg_kblocks[s];
                    const bool bmn = (B_MODE == 1) || (B_MODE == 2 && p.seg_bmn[s]);
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
                        uint8_t* sB = sA + Cfg::A_BYTES;
                        const uint32_t lead_full = mapa_u32(smem_u32(&full[stage]), 0);
                        if (leader) mbar_arrive_expect_tx(&full[stage], 2 * Cfg::STAGE_BYTES);
                        if constexpr (A_MODE == 1) {
                            tma_load_2d_2cta(sA, &p.tmA[s], lead_full, arow, kb * BK);
                            tma_load_2d_2cta(sA + Cfg::A_BYTES / 2, &p.tmA[s], lead_full, arow + 64, kb * BK);
                        } else {
                            tma_load_2d_2cta(sA, &p.tmA[s], lead_full, kb * BK, arow);
                        }
                        if (bmn) {
                            tma_load_2d_2cta(sB, &p.tmB[s], lead_full, brow, kb * BK);
                            tma_load_2d_2cta(sB + Cfg::B_BYTES / 2, &p.tmB[s], lead_full, brow + 64, kb * BK);
                        } else {
                            tma_load_2d_2cta(sB, &p.tmB[s], lead_full, kb * BK, brow);
                        }
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            constexpr uint32_t idesc_k = make_idesc_bf16(BM2, BN, A_MODE == 1, false);
            constexpr uint32_t idesc_mn = make_idesc_bf16(BM2, BN, A_MODE == 1, true);
            constexpr uint32_t a_step = (A_MODE == 1) ? (16 * 128) >> 4 : 2;
            int stage = 0; uint32_t phase = 0;
            int as = 0; uint32_t aphase = 0;
            for (int t = pair; t < num_tiles; t += npairs) {
                int tm_, tn_ = 0;
                if constexpr (B_MODE == 2) decode_tile(p, t, tm_, tn_);
                else if (p.kb_lo_mul[0] | p.kb_hi_mul[0]) decode_tile(p, t, tm_, tn_);
                mbar_wait(&tempty[as], aphase ^ 1);
                tc_fence_after();
                uint32_t touched = 0;
                for (int s = 0; s < p.nseg; ++s) {
                    const int acc = p.seg_acc[s];
                    const uint32_t d_addr = tmem_base + as * Cfg::ACC_COLS + acc * BN;
                    const int kb0 = p.kb_lo_mul[s] * tn_;
                    const int kb1 = p.kb_hi_mul[s] ? min(p.seg_kblocks[s], p.kb_hi_mul[s] * tn_) : p.seg_kblocks[s];
                    const bool bmn = (B_MODE == 1) || (B_MODE == 2 && p.seg_bmn[s]);
                    const uint32_t idesc = bmn ? idesc_mn : idesc_k;
                    const uint32_t b_step = bmn ? (16 * 128) >> 4 : 2;
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                        const uint64_t adesc = (A_MODE == 1) ? make_mnmajor_sw128_desc(a_addr, Cfg::A_BYTES / 2) : make_kmajor_sw128_desc(a_addr);
                        const uint64_t bdesc = bmn ? make_mnmajor_sw128_desc(a_addr + Cfg::A_BYTES, Cfg::B_BYTES / 2)
                                                   : make_kmajor_sw128_desc(a_addr + Cfg::A_BYTES);
                        const int ksteps = (p.k_tail_steps && kb == p.seg_kblocks[s] - 1) ? p.k_tail_steps : BK / 16;
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            if (k < ksteps)
                                umma_bf16_2cta(d_addr, adesc + a_step * k, bdesc + b_step * k, idesc,
                                               ((touched >> acc) & 1u) | (k > 0 ? 1u : 0u));
                        touched |= (1u << acc);
                        umma_commit_2cta(&empty[stage], 3);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
                umma_commit_2cta(&tfull[as], 3);
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
        for (int t = pair; t < num_tiles; t += npairs, ++seq) {
            TileInfo ti;
            int tm2;
            decode_tile(p, t, tm2, ti.tn);
            ti.tm = tm2 * 2 + static_cast<int>(rank);            // row block in units of 128 rows
            ti.row0 = p.a_row0 + tm2 * BM2 + static_cast<int>(rank) * 128;
            ti.col0 = p.b_row0 + ti.tn * BN;
            ti.q = q; ti.lane = lane; ti.tile_seq = seq;
            ti.w = warp - 4; ti.nw = EPI_WARPS; ti.csplit = (warp - 4) >> 2; ti.nsplit = kSplit;
            ti.c0 = ti.csplit * kChunks; ti.c1 = ti.c0 + kChunks;
            ti.tid = threadIdx.x - kNonEpiThreads;
            ti.taddr = tmem_base + as * Cfg::ACC_COLS + (static_cast<uint32_t>(q * 32) << 16);
            Epi::prologue(st, p.epi, ti, epi_smem);
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            Epi::run(st, p.epi, ti, epi_smem);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[as]), 0));
            if (++as == Cfg::ACC_STAGES) { as = 0; aphase ^= 1; }
        }
        Epi::finish(st, p.epi, q, lane);
    }

    tc_fence_before();
    cluster_sync_all();                 // the peer may still be reading this CTA's shared memory / signalling its barriers
    if (warp == 2) { tc_fence_after(); tmem_dealloc2(tmem_base, 512); }
}

// ---- wide variant: one CTA pair owns a 256 x 512 tile (two 256-wide B sub-tiles per K block, one shared A tile) -------------
// Why: every pair kernel delivers ~11-11.5 TB/s from L2 to shared memory (32 KB per CTA and K block) whatever its epilogue -- the
// L2 -> SM throughput cap (~6300 B/clk chip-wide), not the tensor pipe, sets the 0.43 us per K block they all show.  Bytes per
// FLOP scale with 1/BM + 1/BN: a 256 x 512 tile needs 48 KB per CTA and K block for twice the MMAs (-25 %).  The price: its two
// accumulator halves fill TMEM, so the epilogue of a tile is not hidden behind the next tile's MMAs -- worth it where K is long
// (stage 2 of the self-similarity: 32..256 K blocks per tile).  Sub-tile `sub` covers B rows tn*512 + sub*256 ..; the optional
// tile-dependent K ranges of GemmParams apply per 256-row sub-tile index 2*tn + sub (an MMA is skipped outside its range).
template <int STAGES, int EPI_WARPS, class Epi, int B_MODE = 0, int A_MODE = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kNonEpiThreads + 32 * EPI_WARPS, 1)
gemm2w_kernel(const __grid_constant__ GemmParams<Epi> p) {
    static_assert(EPI_WARPS == 8, "8 epilogue warps");
    constexpr int BNW = 512;
    constexpr int A_BYTES = 128 * BK * 2, B_BYTES = 128 * BK * 2;
    constexpr int STAGE_BYTES = A_BYTES + 2 * B_BYTES;           // per CTA
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* epi_smem = smem + STAGES * STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + Epi::SMEM_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tfull = bars + 2 * STAGES;
    uint64_t* tempty = tfull + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = (rank == 0);
    const int pair = blockIdx.x >> 1;
    const int npairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.nseg; ++s) { tma_prefetch_desc(&p.tmA[s]); tma_prefetch_desc(&p.tmB[s]); }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(&tfull[0], 1); mbar_init(&tempty[0], 2 * EPI_WARPS);
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc2(tmem_slot, 512); tmem_relinquish2(); }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();                      // persistent single-wave grid: dependents may be scheduled as CTAs exit
    pdl_wait();                         // the previous kernel of the stream has completed; its writes are visible
    const int num_tiles = num_tiles_of(p);

    // K-block range of segment s for the whole tile (union over the two sub-tiles) and per sub-tile
    auto lo_of = [&](int s, int t256) { return p.kb_lo_mul[s] * t256; };
    auto hi_of = [&](int s, int t256) { return p.kb_hi_mul[s] ? min(p.seg_kblocks[s], p.kb_hi_mul[s] * t256) : p.seg_kblocks[s]; };

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int t = pair; t < num_tiles; t += npairs) {
                int tm, tn;
                decode_tile(p, t, tm, tn);
                const int arow = p.a_row0 + tm * BM2 + static_cast<int>(rank) * 128;
                const int brow = p.b_row0 + tn * BNW + static_cast<int>(rank) * 128;
                for (int s = 0; s < p.nseg; ++s) {
                    const int kb0 = min(lo_of(s, 2 * tn), lo_of(s, 2 * tn + 1));
                    const int kb1 = max(hi_of(s, 2 * tn), hi_of(s, 2 * tn + 1));
                    const bool bmn = (B_MODE == 1) || (B_MODE == 2 && p.seg_bmn[s]);
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        uint8_t* sA = smem + stage * STAGE_BYTES;
                        const uint32_t lead_full = mapa_u32(smem_u32(&full[stage]), 0);
                        if (leader) mbar_arrive_expect_tx(&full[stage], 2 * STAGE_BYTES);
                        if constexpr (A_MODE == 1) {
                            tma_load_2d_2cta(sA, &p.tmA[s], lead_full, arow, kb * BK);
                            tma_load_2d_2cta(sA + A_BYTES / 2, &p.tmA[s], lead_full, arow + 64, kb * BK);
                        } else {
                            tma_load_2d_2cta(sA, &p.tmA[s], lead_full, kb * BK, arow);
                        }
#pragma unroll
                        for (int sub = 0; sub < 2; ++sub) {
                            uint8_t* sB = sA + A_BYTES + sub * B_BYTES;
                            if (bmn) {
                                tma_load_2d_2cta(sB, &p.tmB[s], lead_full, brow + sub * 256, kb * BK);
                                tma_load_2d_2cta(sB + B_BYTES / 2, &p.tmB[s], lead_full, brow + sub * 256 + 64, kb * BK);
                            } else {
                                tma_load_2d_2cta(sB, &p.tmB[s], lead_full, kb * BK, brow + sub * 256);
                            }
                        }
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            constexpr uint32_t idesc_k = make_idesc_bf16(BM2, 256, A_MODE == 1, false);
            constexpr uint32_t idesc_mn = make_idesc_bf16(BM2, 256, A_MODE == 1, true);
            constexpr uint32_t a_step = (A_MODE == 1) ? (16 * 128) >> 4 : 2;
            int stage = 0; uint32_t phase = 0;
            uint32_t aphase = 0;
            for (int t = pair; t < num_tiles; t += npairs) {
                int tm_, tn_;
                decode_tile(p, t, tm_, tn_);
                mbar_wait(&tempty[0], aphase ^ 1);
                tc_fence_after();
                uint32_t touched = 0;                          // bit sub: accumulator half already written in this tile
                for (int s = 0; s < p.nseg; ++s) {
                    const int lo0 = lo_of(s, 2 * tn_), lo1 = lo_of(s, 2 * tn_ + 1);
                    const int hi0 = hi_of(s, 2 * tn_), hi1 = hi_of(s, 2 * tn_ + 1);
                    const int kb0 = min(lo0, lo1), kb1 = max(hi0, hi1);
                    const bool bmn = (B_MODE == 1) || (B_MODE == 2 && p.seg_bmn[s]);
                    const uint32_t idesc = bmn ? idesc_mn : idesc_k;
                    const uint32_t b_step = bmn ? (16 * 128) >> 4 : 2;
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
                        const uint64_t adesc = (A_MODE == 1) ? make_mnmajor_sw128_desc(a_addr, A_BYTES / 2) : make_kmajor_sw128_desc(a_addr);
                        const int ksteps = (p.k_tail_steps && kb == p.seg_kblocks[s] - 1) ? p.k_tail_steps : BK / 16;
#pragma unroll
                        for (int sub = 0; sub < 2; ++sub) {
                            const bool active = sub == 0 ? (kb >= lo0 && kb < hi0) : (kb >= lo1 && kb < hi1);
                            if (!active) continue;
                            const uint32_t b_addr = a_addr + A_BYTES + sub * B_BYTES;
                            const uint64_t bdesc = bmn ? make_mnmajor_sw128_desc(b_addr, B_BYTES / 2) : make_kmajor_sw128_desc(b_addr);
                            const uint32_t d_addr = tmem_base + sub * 256;
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k)
                                if (k < ksteps)
                                    umma_bf16_2cta(d_addr, adesc + a_step * k, bdesc + b_step * k, idesc,
                                                   ((touched >> sub) & 1u) | (k > 0 ? 1u : 0u));
                            touched |= (1u << sub);
                        }
                        umma_commit_2cta(&empty[stage], 3);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
                umma_commit_2cta(&tfull[0], 3);
                aphase ^= 1;
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        uint32_t aphase = 0;
        int seq = 0;
        typename Epi::State st;
        Epi::init(st, p.epi, q, lane);
        constexpr int kSplit = EPI_WARPS / 4;
        constexpr int kChunks = BNW / 32 / kSplit;
        for (int t = pair; t < num_tiles; t += npairs, ++seq) {
            TileInfo ti;
            int tm2;
            decode_tile(p, t, tm2, ti.tn);
            ti.tm = tm2 * 2 + static_cast<int>(rank);
            ti.row0 = p.a_row0 + tm2 * BM2 + static_cast<int>(rank) * 128;
            ti.col0 = p.b_row0 + ti.tn * BNW;
            ti.q = q; ti.lane = lane; ti.tile_seq = seq;
            ti.w = warp - 4; ti.nw = EPI_WARPS; ti.csplit = (warp - 4) >> 2; ti.nsplit = kSplit;
            ti.c0 = ti.csplit * kChunks; ti.c1 = ti.c0 + kChunks;
            ti.tid = threadIdx.x - kNonEpiThreads;
            ti.taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
            Epi::prologue(st, p.epi, ti, epi_smem);
            mbar_wait(&tfull[0], aphase);
            tc_fence_after();
            Epi::run(st, p.epi, ti, epi_smem);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[0]), 0));
            aphase ^= 1;
        }
        Epi::finish(st, p.epi, q, lane);
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) { tc_fence_after(); tmem_dealloc2(tmem_base, 512); }
}

// ---- skewed couple: one CTA pair owns TWO neighbouring 256 x 256 tiles that share their A tile ---------------------------------
// The wide kernel above saves a quarter of the L2 -> shared-memory traffic but leaves its epilogue exposed (both accumulator
// halves complete together and fill TMEM) -- too expensive where K is short (relaxed EMD: 35 K blocks per tile).  Here the two
// halves are skewed by `skew` K blocks so that each half's epilogue hides behind MMAs that do not need its TMEM columns:
//
//   MMA warp, per couple:  wait half 0 free -> K blocks [0, skew) of half 0 alone          (A + B0: 32 KB stages)
//                          wait half 1 free -> K blocks [skew, K) of BOTH halves           (A + B0 + B1: 48 KB, two MMAs per stage)
//                          commit tfull[0]  -> K blocks [0, skew) of half 1 alone          (A + B1)   -> commit tfull[1]
//   epilogue, per couple:  wait tfull[0] -> half 0 (while half 1 finishes)   -> release half 0
//                          wait tfull[1] -> half 1 (while the next couple's half 0 starts) -> release half 1
//
// Per couple and CTA 2*skew*32 + (K - skew)*48 KB instead of 2*K*32 KB (K = 35, skew = 12: -17 %).  The epilogue policy sees two
// ordinary 256-wide tiles (tn = 2 * couple + half).  One segment, K-major B, no tile-dependent K ranges.  p.tiles_n counts couples;
// a ragged last couple computes a phantom half (TMA zero fill or columns the policy masks).
template <int STAGES, int EPI_WARPS, class Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kNonEpiThreads + 32 * EPI_WARPS, 1)
gemm2s_kernel(const __grid_constant__ GemmParams<Epi> p) {
    static_assert(EPI_WARPS == 4 || EPI_WARPS == 8, "4 or 8 epilogue warps");
    constexpr int BN = 256;
    constexpr int A_BYTES = 128 * BK * 2, B_BYTES = 128 * BK * 2;
    constexpr int STAGE_BYTES = A_BYTES + 2 * B_BYTES;           // per CTA
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* epi_smem = smem + STAGES * STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + Epi::SMEM_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tfull = bars + 2 * STAGES;      // [half]
    uint64_t* tempty = tfull + 2;             // [half]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = (rank == 0);
    const int pair = blockIdx.x >> 1;
    const int npairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&p.tmA[0]); tma_prefetch_desc(&p.tmB[0]); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 2 * EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc2(tmem_slot, 512); tmem_relinquish2(); }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();                      // persistent single-wave grid: dependents may be scheduled as CTAs exit
    pdl_wait();                         // the previous kernel of the stream has completed; its writes are visible
    const int num_tiles = num_tiles_of(p);
    const int K = p.seg_kblocks[0];
    const int skew = p.skew < K ? (p.skew > 0 ? p.skew : 0) : K;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int t = pair; t < num_tiles; t += npairs) {
                int tm, tn;
                decode_tile(p, t, tm, tn);
                const int arow = p.a_row0 + tm * BM2 + static_cast<int>(rank) * 128;
                const int brow = p.b_row0 + tn * 2 * BN + static_cast<int>(rank) * 128;
                for (int part = 0; part < 3; ++part) {
                    const int kb0 = (part == 1) ? skew : 0;
                    const int kb1 = (part == 1) ? K : skew;
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        uint8_t* sA = smem + stage * STAGE_BYTES;
                        const uint32_t lead_full = mapa_u32(smem_u32(&full[stage]), 0);
                        if (leader) mbar_arrive_expect_tx(&full[stage], 2 * (A_BYTES + (part == 1 ? 2 : 1) * B_BYTES));
                        tma_load_2d_2cta(sA, &p.tmA[0], lead_full, kb * BK, arow);
                        // a lone half always sits in the first B slot
                        tma_load_2d_2cta(sA + A_BYTES, &p.tmB[0], lead_full, kb * BK, brow + (part == 2 ? BN : 0));
                        if (part == 1) tma_load_2d_2cta(sA + A_BYTES + B_BYTES, &p.tmB[0], lead_full, kb * BK, brow + BN);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            constexpr uint32_t idesc = make_idesc_bf16(BM2, BN, false, false);
            int stage = 0; uint32_t phase = 0;
            uint32_t tph = 0;
            for (int t = pair; t < num_tiles; t += npairs) {
                uint32_t touched = 0;                            // bit h: half h already written in this couple
                for (int part = 0; part < 3; ++part) {
                    const int kb0 = (part == 1) ? skew : 0;
                    const int kb1 = (part == 1) ? K : skew;
                    if (part < 2) {                              // half 0 before part 0, half 1 before part 1
                        mbar_wait(&tempty[part], tph ^ 1);
                        tc_fence_after();
                    }
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
                        const uint64_t adesc = make_kmajor_sw128_desc(a_addr);
                        const int ksteps = (p.k_tail_steps && kb == K - 1) ? p.k_tail_steps : BK / 16;
#pragma unroll
                        for (int slot = 0; slot < 2; ++slot) {
                            if (slot == 1 && part != 1) continue;
                            const int half = (part == 2) ? 1 : slot;
                            const uint64_t bdesc = make_kmajor_sw128_desc(a_addr + A_BYTES + slot * B_BYTES);
                            const uint32_t d_addr = tmem_base + half * BN;
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k)
                                if (k < ksteps)
                                    umma_bf16_2cta(d_addr, adesc + 2 * k, bdesc + 2 * k, idesc, ((touched >> half) & 1u) | (k > 0 ? 1u : 0u));
                            touched |= (1u << half);
                        }
                        umma_commit_2cta(&empty[stage], 3);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    if (part == 1) umma_commit_2cta(&tfull[0], 3);
                }
                umma_commit_2cta(&tfull[1], 3);
                tph ^= 1;
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        uint32_t tph = 0;
        int seq = 0;
        typename Epi::State st;
        Epi::init(st, p.epi, q, lane);
        constexpr int kSplit = EPI_WARPS / 4;
        constexpr int kChunks = BN / 32 / kSplit;
        for (int t = pair; t < num_tiles; t += npairs, ++seq) {
            int tm2, tnc;
            decode_tile(p, t, tm2, tnc);
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                TileInfo ti;
                ti.tn = 2 * tnc + half;
                ti.tm = tm2 * 2 + static_cast<int>(rank);            // row block in units of 128 rows
                ti.row0 = p.a_row0 + tm2 * BM2 + static_cast<int>(rank) * 128;
                ti.col0 = p.b_row0 + ti.tn * BN;
                ti.q = q; ti.lane = lane; ti.tile_seq = 2 * seq + half;
                ti.w = warp - 4; ti.nw = EPI_WARPS; ti.csplit = (warp - 4) >> 2; ti.nsplit = kSplit;
                ti.c0 = ti.csplit * kChunks; ti.c1 = ti.c0 + kChunks;
                ti.tid = threadIdx.x - kNonEpiThreads;
                ti.taddr = tmem_base + half * BN + (static_cast<uint32_t>(q * 32) << 16);
                Epi::prologue(st, p.epi, ti, epi_smem);
                mbar_wait(&tfull[half], tph);
                tc_fence_after();
                Epi::run(st, p.epi, ti, epi_smem);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[half]), 0));
            }
            tph ^= 1;
        }
        Epi::finish(st, p.epi, q, lane);
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) { tc_fence_after(); tmem_dealloc2(tmem_base, 512); }
}

}  // namespace sb
