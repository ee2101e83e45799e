// Work split of the row-sharded SYMMETRIC self-similarity (host side; pure integer logic, replayed by tests/test_ss_jobs_cpu.py
// through strotss_debug_ss_jobs).
//
// Xd, Yd and the sign matrix P of nn/losses.py:56-68 are symmetric, so only the upper block triangle of 256 x 256 tiles has to be
// computed -- a single GPU does that (ss1_decode's trapezoid walk).  Row-sharded over g ranks, each rank owning a block of
// `per` = N / g consecutive rows, the triangle is dealt out circulantly so that every rank computes the same number of tiles
// and every tile (I, J), I != J, is computed by exactly ONE rank (which accounts for its mirror image (J, I) as well):
//
//   rank k computes   its diagonal block (k, k)                                       [upper triangle of tiles]
//                     the h = (g - 1) / 2 blocks (k, k+1 .. k+h), indices mod g       [whole blocks]
//                     for even g one half of the block pair {k, k + g/2}:
//                         k <  g/2 : all own rows      x the first half of the columns of block k + g/2
//                         k >= g/2 : second half of own rows x all columns of block k - g/2
//                     (mirrored, the two halves tile the block: rows x first-half columns | second-half columns x rows)
//
// A block with a wrapped index (k + d >= g) lies left of the diagonal; the rank computes it as the rectangle (own rows) x (those
// columns), which is the mirror image of the tile the triangle asks for -- the same numbers, Xd_ij = Xd_ji.
//
// The rank's tiles are grouped into at most kMaxSsJobs rectangular "jobs" = one stage-1 launch each (rows <= panel):
//   diag = 1 : rows [r0, r1) x columns [r0, c1), tiles at or right of each row tile's diagonal tile (a trapezoid)
//   diag = 0 : rows [r0, r1) x columns [c0, c1), every tile, every tile mirrored
// The gradient products of the mirrored tiles, ss2[J] += P[I,J]^T x^[I], belong to the rank that owns rows J: ss_exchange_plan
// lists the row ranges a rank sends / receives after stage 2.
#pragma once

namespace sb {

constexpr int kSsJobsMax = 8;          // == kMaxSsJobs of kernels.cuh
struct SsJob { int r0, r1, c0, c1, diag, kind; };      // kind: 0 trapezoid right of the diagonal, 1 wrapped blocks, 2 half block
struct SsPlan {
    int njobs;
    SsJob job[kSsJobsMax];
    // stage-2 products this rank computes for rows of OTHER ranks / receives for its own rows
    int nsend, nrecv;
    int send_peer[kSsJobsMax], send_r0[kSsJobsMax], send_r1[kSsJobsMax];
    int recv_peer[kSsJobsMax], recv_r0[kSsJobsMax], recv_r1[kSsJobsMax];
    // rows of the SENDER whose tiles stand behind an entry: the P block of an entry is (src rows) x (r0..r1 of the receiver)
    int send_src_r0[kSsJobsMax], send_src_r1[kSsJobsMax];
    int recv_src_r0[kSsJobsMax], recv_src_r1[kSsJobsMax];
};

// false: this (N, world, panel) cannot use the scheme (ragged blocks, too many jobs) -- the caller falls back to rectangular
// row sharding, where every rank computes all columns of its rows.
inline bool ss_make_plan(int N, int world, int rank, int panel, SsPlan& pl) {
    pl.njobs = pl.nsend = pl.nrecv = 0;
    if (world < 2 || rank < 0 || rank >= world || panel < 256 || panel % 256) return false;
    if (N % (256 * world)) return false;
    const int per = N / world;
    const bool even = (world % 2) == 0;
    if (even && per % 512) return false;
    const int h = (world - 1) / 2;
    if (h + (even ? 1 : 0) > kSsJobsMax) return false;
    const int a = rank * per, b = a + per;
    // contiguous column range right of the diagonal: own block + the blocks ahead that do not wrap (+ half a block)
    int ahead = h;
    if (rank + ahead > world - 1) ahead = world - 1 - rank;
    int cend = (rank + ahead + 1) * per;
    if (even && rank < world / 2) cend += per / 2;
    // wrapped whole blocks: columns [0, wend)
    const int wrapped = h - ahead;
    const int wend = wrapped * per;
    // half block of a rank in the upper half of an even world
    const bool half = even && rank >= world / 2;
    const int hr0 = a + per / 2, hc0 = (rank - world / 2) * per, hc1 = hc0 + per;
    for (int r0 = a; r0 < b; r0 += panel) {
        const int r1 = (r0 + panel < b) ? r0 + panel : b;
        if (pl.njobs + 3 > kSsJobsMax) return false;
        pl.job[pl.njobs++] = SsJob{r0, r1, r0, cend, 1, 0};
        if (wend > 0) pl.job[pl.njobs++] = SsJob{r0, r1, 0, wend, 0, 1};
        if (half && r1 > hr0) pl.job[pl.njobs++] = SsJob{r0 > hr0 ? r0 : hr0, r1, hc0, hc1, 0, 2};
    }
    for (int d = 1; d <= h; ++d) {
        const int to = (rank + d) % world, from = (rank - d + world) % world;
        pl.send_src_r0[pl.nsend] = a; pl.send_src_r1[pl.nsend] = b;
        pl.recv_src_r0[pl.nrecv] = from * per; pl.recv_src_r1[pl.nrecv] = (from + 1) * per;
        pl.send_peer[pl.nsend] = to; pl.send_r0[pl.nsend] = to * per; pl.send_r1[pl.nsend] = (to + 1) * per; ++pl.nsend;
        pl.recv_peer[pl.nrecv] = from; pl.recv_r0[pl.nrecv] = a; pl.recv_r1[pl.nrecv] = b; ++pl.nrecv;
    }
    if (even) {
        const int peer = (rank + world / 2) % world;
        const bool low = rank < world / 2;
        // low rank: computed (own rows) x (first half of the peer's rows) -> sends those rows' products, receives whole own block
        pl.send_src_r0[pl.nsend] = low ? a : a + per / 2; pl.send_src_r1[pl.nsend] = b;
        pl.recv_src_r0[pl.nrecv] = low ? peer * per + per / 2 : peer * per; pl.recv_src_r1[pl.nrecv] = (peer + 1) * per;
        pl.send_peer[pl.nsend] = peer; pl.send_r0[pl.nsend] = peer * per; pl.send_r1[pl.nsend] = peer * per + (low ? per / 2 : per); ++pl.nsend;
        pl.recv_peer[pl.nrecv] = peer; pl.recv_r0[pl.nrecv] = a; pl.recv_r1[pl.nrecv] = low ? b : a + per / 2; ++pl.nrecv;
    }
    return true;
}

// Element offset of the P block that `from` sends inside the receive buffer of the rank whose plan is `pl` (blocks are stored one
// after another in the order of the receive list, each as [source rows][receiver rows]); -1 if `from` sends nothing there.
inline long long ss_recv_offset(const SsPlan& pl, int from, long long* total = nullptr) {
    long long off = 0, found = -1;
    for (int k = 0; k < pl.nrecv; ++k) {
        if (pl.recv_peer[k] == from && found < 0) found = off;
        off += static_cast<long long>(pl.recv_src_r1[k] - pl.recv_src_r0[k]) * (pl.recv_r1[k] - pl.recv_r0[k]);
    }
    if (total) *total = off;
    return found;
}

// Sign-block copies a rank issues after stage 1 of job `k` when the blocks travel instead of the products: the part of the job
// that mirrors into rows of rank `peer` -- rows [i0, i1) x columns [j0, j1) of the matrix -- goes to element offset `dst_off`
// of that rank's window, stored with row length `ld` (= the number of receiver rows of the block).
struct SsCopy { int peer; long long dst_off; int ld; int i0, i1, j0, j1; };
inline int ss_job_copies(int N, int world, int rank, int panel, const SsPlan& pl, int k, SsCopy* out /* [kSsJobsMax] */) {
    const SsJob& jb = pl.job[k];
    int n = 0;
    for (int e = 0; e < pl.nsend; ++e) {
        const int i0 = jb.r0 > pl.send_src_r0[e] ? jb.r0 : pl.send_src_r0[e], i1 = jb.r1 < pl.send_src_r1[e] ? jb.r1 : pl.send_src_r1[e];
        const int j0 = jb.c0 > pl.send_r0[e] ? jb.c0 : pl.send_r0[e], j1 = jb.c1 < pl.send_r1[e] ? jb.c1 : pl.send_r1[e];
        if (i0 >= i1 || j0 >= j1) continue;
        SsPlan pp;
        if (!ss_make_plan(N, world, pl.send_peer[e], panel, pp)) return -1;
        const long long off = ss_recv_offset(pp, rank);
        if (off < 0) return -1;
        const int ld = pl.send_r1[e] - pl.send_r0[e];
        out[n++] = SsCopy{pl.send_peer[e], off + static_cast<long long>(i0 - pl.send_src_r0[e]) * ld + (j0 - pl.send_r0[e]), ld, i0, i1, j0, j1};
    }
    return n;
}

}  // namespace sb
