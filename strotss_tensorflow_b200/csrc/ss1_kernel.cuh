// Self-similarity stage 1, specialised kernel (sm_100a): same math as EpiSS1 in gemm_core.cuh, but the
// two TMEM accumulators are released separately so the epilogue of tile t hides behind the MMAs of
// tile t+1 although both accumulators together fill all 512 TMEM columns:
//
//   MMA warp, per tile:   wait acc1 free -> segment 0 (y^.y^T -> acc1) -> commit tfull[1]
//                         wait acc0 free -> segments 1,2 (delta.x^T + y^.delta^T -> acc0) -> commit tfull[0]
//   epilogue, per tile:   wait tfull[1] -> copy its 128 columns of acc1 (Yd = 1 - acc1) into registers
//                         -> release acc1 (the MMA warp may start the NEXT tile's segment 0 right away)
//                         wait tfull[0] -> stream acc0 in 32-column chunks, combine with the stashed Yd
//                         -> release acc0 (needed only after the next tile's ~35 k-blocks of segment 0)
//
// 128 fp32 of stash per epilogue thread need more registers than a 384-thread CTA grants uniformly,
// so the non-epilogue warpgroup shrinks to 56 registers and the two epilogue warpgroups grow to 224
// (setmaxnreg), as warp-specialised Hopper/Blackwell GEMMs do.
//
// PAIR = true runs the same protocol on a CTA pair (cta_group::2, see gemm2_core.cuh): 256 x 256 tiles,
// each CTA stages 128 rows of A and 128 rows of B, barriers full/tempty live in the leader CTA.
//
// MERGED = true (CTA pair only): segments 0 and 2 share their A operand (y^ rows of the tile), so they run as ONE K loop whose
// stages hold A = y^_I and two B tiles (delta_J -> accumulator "D", y^_J -> accumulator "Y"): 48 KB per CTA and K block for two
// MMAs instead of 2 x 32 KB -- the pair kernels sit at the L2 -> shared-memory throughput cap, and this removes a sixth of the
// stage-1 operand traffic (35 x (32 + 48) KB instead of 105 x 32 KB per tile and CTA).  Both accumulators are then busy until the
// merged loop ends, so to keep the epilogue hidden
//   * the delta.x^T segment is split around the merged loop: its first kblocks - tail_blocks K blocks run BEFORE it (they need only
//     the D columns), its last tail_blocks after it (while they run the epilogue stashes Y and releases those columns), and
//   * the two 256-column TMEM regions swap roles every tile: the region whose Y was stashed is free almost at once and takes the
//     next tile's D, whose leading delta.x^T blocks cover the epilogue streaming the previous D out of the other region.
//   tempty[] is indexed by REGION (each region is released once per tile), tfull[] by role (0 = D, 1 = Y).
#pragma once
#include "gemm2_core.cuh"

namespace sb {

constexpr int kSs1BN = 256;
constexpr int kSs1EpiWarps = 8;
constexpr int kSs1Stages = 4;
constexpr int kSs1Threads = kNonEpiThreads + 32 * kSs1EpiWarps;
using Ss1Epi = EpiSS1<kSs1BN, kSs1EpiWarps>;

struct Ss1Params {
    CUtensorMap tmA[3];
    CUtensorMap tmB[3];          // segment 0: (y^, y^) -> acc1 ; 1: (delta, x^) -> acc0 ; 2: (y^, delta) -> acc0
    int kblocks;                 // K blocks per segment
    int k_tail_steps;            // 16-wide MMA steps that carry data in the last K block (0 = all four; see GemmParams)
    int tiles_m, tiles_n, group_n;
    int tri;                     // always 0 here (decode_tile's triangular walk is used by the covariance kernel)
    int a_row0, b_row0;
    // trap = 1 (CTA-pair kernel, symmetric mode, a_row0 == b_row0): only tiles with tn >= tm are visited -- the block
    // trapezoid of a row panel of a symmetric matrix -- and every tile right of its own diagonal tile accounts for its
    // mirror image.  Same L2 raster as decode_tile: groups of group_n column tiles, row tiles swept inside a group.
    int trap;
    int tail_blocks;             // MERGED kernel: K blocks of the delta.x^T segment issued after the merged loop
    Ss1Epi::Params epi;
};

__host__ __device__ __forceinline__ int ss1_num_tiles(const Ss1Params& p) {
    if (!p.trap) return num_tiles_of(p);
    const int tmx = min(p.tiles_m, p.tiles_n);
    return tmx * p.tiles_n - tmx * (tmx - 1) / 2;
}

__host__ __device__ __forceinline__ void ss1_decode(const Ss1Params& p, int t, int& tm, int& tn) {
    if (!p.trap) { decode_tile(p, t, tm, tn); return; }
    for (int c_lo = 0; c_lo < p.tiles_n; c_lo += p.group_n) {
        const int c_hi = min(p.tiles_n, c_lo + p.group_n);
        const int rows = min(p.tiles_m, c_hi);              // row tiles that own a tile in this column group
        for (int r = 0; r < rows; ++r) {
            const int first = max(c_lo, r);
            const int len = c_hi - first;
            if (t < len) { tm = r; tn = first + t; return; }
            t -= len;
        }
    }
    tm = 0; tn = 0;                                         // unreachable for t < ss1_num_tiles(p)
}

constexpr int kSs1PairStages = 6;
constexpr int kSs1MergedStages = 4;                       // 48 KB stages
constexpr int kSs1SmemBytes = kSs1Stages * TileCfg<kSs1BN, 2>::STAGE_BYTES + Ss1Epi::SMEM_BYTES + (2 * kSs1Stages + 4) * 8 + 16 + 1024;
constexpr int kSs1PairSmemBytes = kSs1PairStages * PairCfg<2>::STAGE_BYTES + Ss1Epi::SMEM_BYTES + (2 * kSs1PairStages + 4) * 8 + 16 + 1024;
constexpr int kSs1MergedSmemBytes = kSs1MergedStages * 3 * 128 * BK * 2 + Ss1Epi::SMEM_BYTES + (2 * kSs1MergedStages + 4) * 8 + 16 + 1024;
static_assert(kSs1MergedSmemBytes <= 232448, "shared memory budget exceeded");

template <bool PAIR, bool MERGED = false>
__device__ __forceinline__ void ss1_body(const Ss1Params& p) {
    static_assert(PAIR || !MERGED, "the merged K loop exists for the CTA-pair kernel only");
    constexpr int BN = kSs1BN;
    constexpr int STAGES = MERGED ? kSs1MergedStages : (PAIR ? kSs1PairStages : kSs1Stages);
    constexpr int A_BYTES = 128 * BK * 2;
    constexpr int B_BYTES = (PAIR ? 128 : BN) * BK * 2;
    constexpr int STAGE_BYTES = A_BYTES + (MERGED ? 2 : 1) * B_BYTES;
    constexpr int TILE_M = PAIR ? BM2 : BM;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* epi_smem = smem + STAGES * STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + Ss1Epi::SMEM_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tfull = bars + 2 * STAGES;      // [acc]
    uint64_t* tempty = tfull + 2;             // [acc]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int rank = PAIR ? static_cast<int>(cluster_ctarank()) : 0;
    const bool leader = (rank == 0);
    const int first = PAIR ? (blockIdx.x >> 1) : blockIdx.x;       // first tile and tile stride of this CTA (pair)
    const int stride = PAIR ? (gridDim.x >> 1) : gridDim.x;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < 3; ++s) { tma_prefetch_desc(&p.tmA[s]); tma_prefetch_desc(&p.tmB[s]); }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], (PAIR ? 2 : 1) * kSs1EpiWarps); }
        fence_barrier_init();
    }
    if (warp == 2) {
        if constexpr (PAIR) { tmem_alloc2(tmem_slot, 512); tmem_relinquish2(); }
        else { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    }
    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();                      // persistent single-wave grid: dependents may be scheduled as CTAs exit
    pdl_wait();                         // the previous kernel of the stream has completed; its writes are visible
    const int num_tiles = ss1_num_tiles(p);
    const int ksplit = (p.kblocks > p.tail_blocks) ? p.kblocks - p.tail_blocks : 0;      // MERGED only
    (void)ksplit;

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (warp == 0 && lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int t = first; t < num_tiles; t += stride) {
                int tm, tn;
                ss1_decode(p, t, tm, tn);
                const int arow = p.a_row0 + tm * TILE_M + rank * 128, brow = p.b_row0 + tn * BN + rank * 128;
                if constexpr (MERGED) {
                    // part 0: delta.x^T blocks [0, ksplit) ; part 1: merged y^.(delta | y^)^T blocks ; part 2: delta.x^T blocks [ksplit, kblocks)
                    for (int part = 0; part < 3; ++part) {
                        const bool merged = (part == 1);
                        const int kb0 = (part == 2) ? ksplit : 0;
                        const int kb1 = (part == 0) ? ksplit : p.kblocks;
                        for (int kb = kb0; kb < kb1; ++kb) {
                            mbar_wait(&empty[stage], phase ^ 1);
                            uint8_t* sA = smem + stage * STAGE_BYTES;
                            const uint32_t lead_full = mapa_u32(smem_u32(&full[stage]), 0);
                            if (merged) {
                                if (leader) mbar_arrive_expect_tx(&full[stage], 2 * (A_BYTES + 2 * B_BYTES));
                                tma_load_2d_2cta(sA, &p.tmA[2], lead_full, kb * BK, arow);
                                tma_load_2d_2cta(sA + A_BYTES, &p.tmB[2], lead_full, kb * BK, brow);
                                tma_load_2d_2cta(sA + A_BYTES + B_BYTES, &p.tmB[0], lead_full, kb * BK, brow);
                            } else {
                                if (leader) mbar_arrive_expect_tx(&full[stage], 2 * (A_BYTES + B_BYTES));
                                tma_load_2d_2cta(sA, &p.tmA[1], lead_full, kb * BK, arow);
                                tma_load_2d_2cta(sA + A_BYTES, &p.tmB[1], lead_full, kb * BK, brow);
                            }
                            if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                    continue;
                }
                for (int s = 0; s < 3; ++s) {
                    for (int kb = 0; kb < p.kblocks; ++kb) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        uint8_t* sA = smem + stage * STAGE_BYTES;
                        if constexpr (PAIR) {
                            const uint32_t lead_full = mapa_u32(smem_u32(&full[stage]), 0);
                            if (leader) mbar_arrive_expect_tx(&full[stage], 2 * STAGE_BYTES);
                            tma_load_2d_2cta(sA, &p.tmA[s], lead_full, kb * BK, arow);
                            tma_load_2d_2cta(sA + A_BYTES, &p.tmB[s], lead_full, kb * BK, brow);
                        } else {
                            mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
                            tma_load_2d(sA, &p.tmA[s], &full[stage], kb * BK, arow);
                            tma_load_2d(sA + A_BYTES, &p.tmB[s], &full[stage], kb * BK, brow);
                        }
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        } else if (warp == 1 && lane == 0 && leader) {
            constexpr uint32_t idesc = make_idesc_bf16(TILE_M, BN);
            int stage = 0; uint32_t phase = 0;
            uint32_t tph = 0;
            int seq = 0;
            for (int t = first; t < num_tiles; t += stride, ++seq) {
                if constexpr (MERGED) {
                    const int dreg = seq & 1, yreg = dreg ^ 1;       // TMEM region of the D / Y accumulator of this tile
                    const uint32_t d_addr = tmem_base + dreg * BN, y_addr = tmem_base + yreg * BN;
                    uint32_t dacc = 0;                               // D already written in this tile
                    for (int part = 0; part < 3; ++part) {
                        const bool merged = (part == 1);
                        const int kb0 = (part == 2) ? ksplit : 0;
                        const int kb1 = (part == 0) ? ksplit : p.kblocks;
                        if (part < 2) {                              // D region before part 0, Y region before the merged loop
                            mbar_wait(&tempty[part == 0 ? dreg : yreg], tph ^ 1);
                            tc_fence_after();
                        }
                        for (int kb = kb0; kb < kb1; ++kb) {
                            mbar_wait(&full[stage], phase);
                            tc_fence_after();
                            const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
                            const uint64_t adesc = make_kmajor_sw128_desc(a_addr);
                            const uint64_t bdesc = make_kmajor_sw128_desc(a_addr + A_BYTES);
                            const int ksteps = (p.k_tail_steps && kb == p.kblocks - 1) ? p.k_tail_steps : BK / 16;
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k) {
                                if (k >= ksteps) break;
                                umma_bf16_2cta(d_addr, adesc + 2 * k, bdesc + 2 * k, idesc, dacc | (k > 0 ? 1u : 0u));
                            }
                            dacc = 1u;
                            if (merged) {
                                const uint64_t ydesc = make_kmajor_sw128_desc(a_addr + A_BYTES + B_BYTES);
#pragma unroll
                                for (int k = 0; k < BK / 16; ++k) {
                                    if (k >= ksteps) break;
                                    umma_bf16_2cta(y_addr, adesc + 2 * k, ydesc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                                }
                            }
                            umma_commit_2cta(&empty[stage], 3);
                            if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        }
                        if (merged) umma_commit_2cta(&tfull[1], 3);
                    }
                    umma_commit_2cta(&tfull[0], 3);
                    tph ^= 1;
                    continue;
                }
                for (int s = 0; s < 3; ++s) {
                    const int acc = (s == 0) ? 1 : 0;
                    if (s < 2) {                       // first segment of each accumulator: wait until it is free
                        mbar_wait(&tempty[acc], tph ^ 1);
                        tc_fence_after();
                    }
                    const uint32_t d_addr = tmem_base + acc * BN;
                    for (int kb = 0; kb < p.kblocks; ++kb) {
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
                        const uint64_t adesc = make_kmajor_sw128_desc(a_addr);
                        const uint64_t bdesc = make_kmajor_sw128_desc(a_addr + A_BYTES);
                        const int ksteps = (p.k_tail_steps && kb == p.kblocks - 1) ? p.k_tail_steps : BK / 16;
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            if (k >= ksteps) break;
                            const uint32_t accum = ((s == 2) || kb > 0 || k > 0) ? 1u : 0u;
                            if constexpr (PAIR) umma_bf16_2cta(d_addr, adesc + 2 * k, bdesc + 2 * k, idesc, accum);
                            else umma_bf16(d_addr, adesc + 2 * k, bdesc + 2 * k, idesc, accum);
                        }
                        if constexpr (PAIR) umma_commit_2cta(&empty[stage], 3); else umma_commit(&empty[stage]);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    if (s == 0) { if constexpr (PAIR) umma_commit_2cta(&tfull[1], 3); else umma_commit(&tfull[1]); }
                }
                if constexpr (PAIR) umma_commit_2cta(&tfull[0], 3); else umma_commit(&tfull[0]);
                tph ^= 1;
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        const Ss1Epi::Params& P = p.epi;
        const int q = warp & 3;
        const int csplit = (warp - 4) >> 2;                  // column half handled by this warp
        constexpr int kChunks = BN / 32 / 2;                 // 4 chunks of 32 columns per warp
        const int tid = threadIdx.x - kNonEpiThreads;
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        uint32_t tph = 0;
        int seq = 0;
        for (int t = first; t < num_tiles; t += stride, ++seq) {
            int tm, tn;
            ss1_decode(p, t, tm, tn);
            const int row0 = p.a_row0 + tm * TILE_M + rank * 128, col0 = p.b_row0 + tn * BN;
            // per-column vectors u, w of this tile -> shared memory (double-buffered across tiles)
            float* buf = reinterpret_cast<float*>(epi_smem) + (seq & 1) * 2 * BN;
            for (int i = tid; i < BN; i += 32 * kSs1EpiWarps) {
                const int col = col0 + i;
                buf[i] = (col < P.N) ? P.u[col] : 0.f;
                buf[BN + i] = (col < P.N) ? P.w[col] : 0.f;
            }
            epi_bar_sync<32 * kSs1EpiWarps>();
            const float* su = buf;
            const float* sw = buf + BN;
            const int row = row0 + q * 32 + lane;
            const bool rvalid = row < P.row_end;
            const float ui = rvalid ? P.u[row] : 0.f;
            const float wi = rvalid ? P.w[row] : 0.f;
            const bool both = P.sym && (p.trap ? (tn > tm) : (col0 >= P.panel_end));
            const int dreg = MERGED ? (seq & 1) : 0, yreg = dreg ^ 1;     // TMEM regions (MERGED: roles swap every tile)

            // ---- accumulator 1 (y^.y^T): stash Yd = 1 - acc1, then hand the columns back to the MMA warp
            float yd[kChunks * 32];
            mbar_wait(&tfull[1], tph);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
                uint32_t a1[32];
                tmem_ld32(lane_addr + yreg * BN + (csplit * kChunks + c) * 32, a1);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 32; ++e) yd[c * 32 + e] = 1.f - __uint_as_float(a1[e]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[yreg]), 0)); else mbar_arrive(&tempty[1]);
            }

            // ---- accumulator 0 (delta form of Xd - Yd), 32 columns at a time
            mbar_wait(&tfull[0], tph);
            tc_fence_after();
            float loss = 0.f, racc = 0.f;
            __nv_bfloat16* prow = P.P + static_cast<long long>(row - P.panel_row0) * P.ldp;
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
                const int cc = csplit * kChunks + c;
                uint32_t a0[32];
                tmem_ld32(lane_addr + dreg * BN + cc * 32, a0);
                tmem_ld_wait();
                const int colbase = col0 + cc * 32;
                const bool pwrite = P.write_p && rvalid && colbase - P.p_col0 < P.ldp;
                // 8 columns at a time: the bf16 P values leave as one 16-byte store, and sign_col * Xd
                // overwrites the accumulator registers it was computed from (a0[e] is dead by then)
#pragma unroll
                for (int e8 = 0; e8 < 32; e8 += 8) {
                    uint32_t packed[4];
#pragma unroll
                    for (int e2 = 0; e2 < 8; e2 += 2) {
                        float pv[2];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int e = e8 + e2 + h;
                            const int col = colbase + e;
                            const bool live = rvalid && (col < P.N) && (col != row);
                            const float diff = -__uint_as_float(a0[e]);
                            const float y = yd[c * 32 + e];
                            const float uj = su[cc * 32 + e], wj = sw[cc * 32 + e];
                            const float tc = fmaf(diff, uj, y * wj);
                            const float tr = fmaf(diff, ui, y * wi);
                            const float sc = live ? ((tc > 0.f) ? 1.f : ((tc < 0.f) ? -1.f : 0.f)) : 0.f;
                            const float sr = live ? ((tr > 0.f) ? 1.f : ((tr < 0.f) ? -1.f : 0.f)) : 0.f;
                            loss += live ? fabsf(tr) : 0.f;
                            racc = fmaf(sr, y + diff, racc);
                            if (both) loss += live ? fabsf(tc) : 0.f;
                            a0[e] = __float_as_uint(sc * (y + diff));
                            pv[h] = fmaf(sc, uj, sr * ui);
                        }
                        packed[e2 >> 1] = pack_bf16x2(pv[0], pv[1]);
                    }
                    if (pwrite) {
                        uint4* pdst = reinterpret_cast<uint4*>(prow + (colbase - P.p_col0) + e8);
                        const uint4 pval = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                        if (P.write_p == 2) __stcs(pdst, pval);      // experiment: streaming (evict-first) stores for the panel
                        else *pdst = pval;
                    }
                }
                if (both) {
                    // transpose-reduce: lane l ends with the sum over the warp's 32 rows of column l
#pragma unroll
                    for (int sft = 16; sft >= 1; sft >>= 1) {
                        const bool up = (lane & sft) != 0;
#pragma unroll
                        for (int e = 0; e < sft; ++e) {
                            const float send = __uint_as_float(up ? a0[e] : a0[e + sft]);
                            const float keep = __uint_as_float(up ? a0[e + sft] : a0[e]);
                            a0[e] = __float_as_uint(keep + __shfl_xor_sync(0xffffffffu, send, sft));
                        }
                    }
                    const int col = colbase + lane;
                    if (col < P.N)
                        P.rcol_part[(static_cast<long long>(row0 / BM) * 4 + q) * P.N + col] = __uint_as_float(a0[0]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[dreg]), 0)); else mbar_arrive(&tempty[0]);
            }
            if (rvalid) {
                const long long slot = static_cast<long long>(col0 / BN) * 2 + csplit;
                P.loss_part[slot * P.N + row] = loss;
                P.r_part[slot * P.N + row] = racc;
            }
            tph ^= 1;
        }
    }

    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        if constexpr (PAIR) tmem_dealloc2(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
    }
}

__global__ void __launch_bounds__(kSs1Threads, 1) ss1_kernel(const __grid_constant__ Ss1Params p) { ss1_body<false>(p); }

// tiles_m of Ss1Params counts 256-row tiles for the pair kernel
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kSs1Threads, 1)
ss1_pair_kernel(const __grid_constant__ Ss1Params p) { ss1_body<true>(p); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kSs1Threads, 1)
ss1_pair_merged_kernel(const __grid_constant__ Ss1Params p) { ss1_body<true, true>(p); }

}  // namespace sb
