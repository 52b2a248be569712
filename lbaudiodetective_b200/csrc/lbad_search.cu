/*
 * lbad_search.cu — fingerprint database search kernels for sm_100a.
 *
 * Replaces, on the GPU, the reference's matcher (file:line into /root/reference/LBAudioDetective/LBAudioDetectiveFingerprint.m):
 *   CompareSubfingerprints   FP.m:151-176   masked 2-bit-pair agreement ratio hits/possible
 *   CompareToFingerprint     FP.m:119-149   swap so fp1 is the longer, time-offset search, f32 mean, running max
 * evaluated for every (query, database clip) pair as  score = CompareToFingerprint(clip, query, range)  — the
 * argument order of LBAudioDetectiveTests.m:68 — followed by a per-query top-k ordered (score desc, clip index asc).
 *
 * Bit-exactness: hits and possible are integer popcounts (identity in SURVEY.md §2.3); hits/possible is produced
 * by multiply + two FMAs against a correctly rounded reciprocal (Markstein), verified exhaustively to equal the
 * IEEE quotient for all 0 <= hits <= possible <= 256 (tests/test_host_logic.py); the per-offset sum adds the
 * ratios in increasing subfingerprint order in f32; the mean uses an IEEE divide; the running max reproduces
 * Apple's MAX macro (NaN from an empty fingerprint leaves the match at 0).
 *
 * Mapping: one warp = 32 queries (one per lane, query words in registers) x a contiguous chunk of clips; the
 * database words are warp-uniform broadcast loads, so the ~600 MB database streams through L2 once per 32 queries.
 * The kernel is integer-pipe bound (LOP3 + POPC), not HBM bound.
 */
#include "lbad_common.cuh"
#include "lbad_math.cuh"
#include <vector>
#include <algorithm>
#include <string.h>
#include <stdlib.h>

namespace lbad {

constexpr int SEARCH_WARPS = 4;
constexpr uint32_t EMPTY_IDX = 0xffffffffu;

__constant__ float c_rcp[257];            /* c_rcp[p] = RN(1/p), c_rcp[0] = 0 */

template <int W> struct PairMask { uint32_t w[W]; };
template <int W> __host__ __device__ inline PairMask<W> make_mask(uint32_t pairs) {
    PairMask<W> m;
    for (int i = 0; i < W; i++) m.w[i] = pairs >= 32u * (i + 1) ? 0xffffffffu : (pairs > 32u * i ? ((1u << (pairs - 32u * i)) - 1u) : 0u);
    return m;
}

/* per-lane top-k kept in shared memory, lane-strided: slot r of lane l at [r*32 + l] */
struct TopK {
    float* sc; uint32_t* id; int k, lane;
    __device__ __forceinline__ void init() { for (int r = 0; r < k; r++) { sc[r * 32 + lane] = -1.0f; id[r * 32 + lane] = EMPTY_IDX; } }
    __device__ __forceinline__ float worst() const { return sc[(k - 1) * 32 + lane]; }
    /* clips arrive in ascending index order, so a tie never displaces an earlier clip */
    __device__ __forceinline__ void insert(float s, uint32_t c) {
        int pos = k - 1;
        while (pos > 0 && sc[(pos - 1) * 32 + lane] < s) { sc[pos * 32 + lane] = sc[(pos - 1) * 32 + lane]; id[pos * 32 + lane] = id[(pos - 1) * 32 + lane]; pos--; }
        sc[pos * 32 + lane] = s; id[pos * 32 + lane] = c;
    }
};

/* small non-negative integer -> float without the (quarter-rate, XU-pipe) I2F: 0x4B000000 is 2^23 */
__device__ __forceinline__ float small_uint_to_float(uint32_t v) { return __int_as_float(0x4B000000u | v) - 8388608.0f; }

/* hits/possible as f32, exactly the IEEE quotient (FP.m:171-175); possible == 0 -> 0 because rcp == 0 and hits == 0 */
__device__ __forceinline__ float ratio_exact(uint32_t hits, float fposs, float rcp) {
    const float fh = small_uint_to_float(hits);
    const float q0 = __fmul_rn(fh, rcp);
    const float rem = fmaf(-q0, fposs, fh);
    return fmaf(rem, rcp, q0);
}

/* sum / count as the IEEE quotient for count <= 16 and sum in {0} U [2^-8, 16] (verified exhaustively over every
 * float in [0, 64]: the multiply + two-FMA form only misrounds below 1e-37); larger counts use the divide */
template <int CQ>
__device__ __forceinline__ float mean_exact(float sum) {
    if (CQ > 16) return __fdiv_rn(sum, (float)CQ);
    constexpr float b = (float)CQ, r = 1.0f / (float)CQ;
    const float q0 = __fmul_rn(sum, r);
    const float rem = fmaf(-q0, b, sum);
    return fmaf(rem, r, q0);
}

template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return r;
}
/* hit_word() of lbad_math.cuh as exactly two LOP3: u = (p1|m1) & ~(p1^p2) [0xA4], h = u & ~(m1^m2) [0x90] */
__device__ __forceinline__ uint32_t hit_word2(uint32_t p1, uint32_t m1, uint32_t p2, uint32_t m2) {
    return lop3<0x90>(lop3<0xA4>(p1, m1, p2), m1, m2);
}

/* popcount of the W hit words with a carry-save adder per group of three words (x ^ y ^ z has the ones, maj(x, y, z) the
 * twos): POPC is a quarter-rate XU-pipe instruction and the search kernel is bound by it, LOP3 is not. */
template <int W>
__device__ __forceinline__ uint32_t popc_words(const uint32_t (&h)[W]) {
    uint32_t ones = 0, twos = 0;
#pragma unroll
    for (int w = 0; w + 3 <= W; w += 3) {
        ones += __popc(lop3<0x96>(h[w], h[w + 1], h[w + 2]));
        twos += __popc(lop3<0xE8>(h[w], h[w + 1], h[w + 2]));
    }
#pragma unroll
    for (int w = (W / 3) * 3; w < W; w++) ones += __popc(h[w]);
    return ones + 2 * twos;
}

/* per database subfingerprint: (float)possible and RN(1/possible) over the FULL length — what every unmasked compare needs */
template <int W>
__global__ void meta_kernel(const uint32_t* __restrict__ db, float2* __restrict__ meta, const uint64_t first, const uint64_t count,
                            const uint32_t pairs_full, uint32_t* __restrict__ irregular) {
    uint32_t bad = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t* src = db + (first + i) * 2 * W;
        uint32_t possible = 0, both = 0;
#pragma unroll
        for (int w = 0; w < W; w++) { possible += __popc(src[w] | src[W + w]); both |= src[w] & src[W + w]; }
        meta[first + i] = make_float2((float)possible, c_rcp[possible]);
        bad += (possible != pairs_full || both != 0) ? 1u : 0u;
    }
    bad = __reduce_add_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(irregular, bad);
}

constexpr int STAGE_SUBFPS = 128;          /* database subfingerprints staged in shared memory per tile (whole clips only) */

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

/* Fast path: every clip in the database has between CQ and STAGE_SUBFPS subfingerprints, so the clip is always fp1
 * (FP.m:123-131: no swap when counts are equal).  Offset-outer / query-subfingerprint-inner, exactly the reference's loop
 * nest.  The four warps of a CTA work on the SAME clips with different query groups: tiles of whole clips (words and
 * (possible, 1/possible)) are brought into shared memory by cp.async, double buffered, and read back as warp-uniform
 * broadcast loads; the CQ query subfingerprints live in registers. */
#ifndef LBAD_SEARCH_MIN_CTAS
#define LBAD_SEARCH_MIN_CTAS 6          /* at most 85 registers: six CTAs (24 warps) per SM for the longest queries; measured 7.61 ms on a 125,000-clip shard against 8.19 with 7 (spills) and 9.38 with 1 */
#endif
template <int W, int CQ, bool MASKED>
__global__ void __launch_bounds__(SEARCH_WARPS * 32, CQ >= 4 ? LBAD_SEARCH_MIN_CTAS : 0)      /* (0: no constraint — the short-query variants are far below it and compile best left alone) */
search_fast_kernel(const uint32_t* __restrict__ db, const float2* __restrict__ meta, const uint32_t* __restrict__ offsets, const uint32_t n_clips,
                   const uint32_t clip_base, const uint32_t* __restrict__ clip_ids, const uint32_t* __restrict__ qwords, const uint32_t n_q, const uint32_t pairs, const int k,
                   const uint32_t n_qgroups, const uint32_t clips_per_chunk, float* __restrict__ part_sc, uint32_t* __restrict__ part_id,
                   float* __restrict__ all_scores, const uint32_t groups_per_chunk, const int db_regular, const uint32_t rep,
                   const uint32_t all_stride, const float* __restrict__ floor_sc, const uint32_t uniform) {
    /* uniform: every clip has that many subfingerprints (0: ragged) — clip c then starts at subfingerprint c * uniform and neither the
     * tile bounds nor the clips need the offsets array */
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    /* rep > 1 (one or two query groups, i.e. at most 64 queries): the warps a CTA would leave idle take a share of the tile's clips —
     * warp wid serves query group wid % n_qgroups on the clips c0 + sub, c0 + sub + rep, .. (sub = wid / n_qgroups) and keeps its own list */
    const uint32_t chunk = blockIdx.x / groups_per_chunk;
    const uint32_t qg = rep > 1 ? (uint32_t)wid % n_qgroups : (blockIdx.x % groups_per_chunk) * SEARCH_WARPS + wid, sub = rep > 1 ? (uint32_t)wid / n_qgroups : 0u;
    const uint32_t q = qg * 32 + lane;
    const bool qvalid = qg < n_qgroups && q < n_q;
    TopK top{reinterpret_cast<float*>(smem_raw) + (size_t)wid * 2 * k * 32, reinterpret_cast<uint32_t*>(smem_raw) + (size_t)wid * 2 * k * 32 + (size_t)k * 32, k, lane};
    top.init();
    /* staging buffers after the top-k lists: [2][STAGE_SUBFPS][2W] words, then [2][STAGE_SUBFPS] float2 */
    uint32_t* st_words = reinterpret_cast<uint32_t*>(smem_raw) + (size_t)SEARCH_WARPS * 2 * k * 32;
    float2* st_meta = reinterpret_cast<float2*>(st_words + 2 * STAGE_SUBFPS * 2 * W);
    const PairMask<W> mask = make_mask<W>(pairs);
    uint32_t qp[CQ][W], qm[CQ][W];
#pragma unroll
    for (int i = 0; i < CQ; i++)
#pragma unroll
        for (int w = 0; w < W; w++) {
            qp[i][w] = qvalid ? qwords[((size_t)q * CQ + i) * 2 * W + w] : 0u;
            qm[i][w] = qvalid ? qwords[((size_t)q * CQ + i) * 2 * W + W + w] : 0u;
        }
    const uint32_t c_begin = chunk * clips_per_chunk;
    const uint32_t c_end = min(n_clips, c_begin + clips_per_chunk);
    /* floor_sc (optional): per query the k-th best score of a SAMPLE of the database (the threshold pass of lbadcu_db_search_device):
     * at least k clips score that much, so a clip scoring strictly less is in nobody's top k and never enters a list.  Folded into the
     * running `worst` as the largest float below the floor, the test stays the one compare per clip it was — and the per-lane list
     * insertion (which the whole warp walks through whenever ANY lane inserts) becomes rare instead of happening for nearly every clip. */
    float floor_below = -1.0f;
    if (floor_sc && qvalid) { const float f = floor_sc[(size_t)q * k + (k - 1)]; if (f > 0.0f) floor_below = __int_as_float(__float_as_int(f) - 1); }
    float worst = fmaxf(top.worst(), floor_below);
    /* "Regular" codes: every one of the first `pairs` ranks carries exactly one sign bit, which is what extraction produces whenever
     * the selected coefficients are non-zero.  Then M = ~P on those ranks, possible = pairs, and a pair hits iff the P bits agree:
     * one LOP3 per word instead of two, no M plane, no per-subfingerprint (possible, 1/possible).  The query side is checked here (a
     * warp takes the short form only if all of its queries qualify); the database side per TILE: a database found regular when it was
     * appended (db_regular) needs no check, in any other the subfingerprints of a landed tile are inspected — so the few clips with
     * empty ranks (digital silence) cost their own tiles the general form, not the whole database. */
    bool q_regular = false;
    if (!MASKED) {
        bool mine = true;
#pragma unroll
        for (int i = 0; i < CQ; i++) {
            uint32_t both = 0, cover = 0;
#pragma unroll
            for (int w = 0; w < W; w++) { both |= qp[i][w] & qm[i][w]; cover += __popc(qp[i][w] | qm[i][w]); }
            mine = mine && both == 0 && cover == pairs;
        }
        q_regular = __all_sync(0xffffffffu, mine || !qvalid);
    }
    uint32_t qnib = 0;                                                          /* nibble i: the fourth-word bits of query subfingerprint i */
    if constexpr (W == 4) {
#pragma unroll
        for (int i = 0; i < CQ; i++) qnib |= (qp[i][3] & 0xFu) << (4 * i);
    }
    float* ratio_tab = reinterpret_cast<float*>(st_meta + 2 * STAGE_SUBFPS);        /* [pairs + 1]: (float)h / (float)pairs */
    for (uint32_t h = tid; h <= pairs; h += SEARCH_WARPS * 32) ratio_tab[h] = pairs ? __fdiv_rn((float)h, (float)pairs) : 0.0f;
    /* 100 ranks (the reference's default subfingerprint) are three full words and FOUR bits of the fourth: the fourth word's bits of eight
     * consecutive subfingerprints are gathered into one "window" word per subfingerprint when a tile lands, so that a query's leftover
     * bits against ALL of its offset's subfingerprints are one XOR and one nibble-wise population count in plain integer arithmetic —
     * and a compare needs TWO POPC (carry-save over the three full words) instead of three.  The kernel was bound by exactly that
     * quarter-rate instruction.  miss_tab[m] = (100 - m) / 100 for m mismatching ranks. */
    constexpr bool CAN100 = W == 4 && !MASKED && CQ >= 4;      /* (shorter queries: the per-offset window arithmetic is not amortised — measured 0.83 against 0.64 ms at CQ = 1) */
    /* CTA-uniform, and only when EVERY warp of the CTA takes the short form: the landed tile is then rewritten in place (below), which
     * the general form of another warp could not read */
    bool cta100 = false;
    if constexpr (CAN100) cta100 = __syncthreads_and(q_regular && pairs == 100) != 0;
    if constexpr (CAN100) {
        /* The carry-save adder over the three full words needs ones = x0 ^ x1 ^ x2 and twos = maj(x0, x1, x2) with x_w = a_w ^ q_w.  The
         * ones word is (a0 ^ a1 ^ a2) ^ (q0 ^ q1 ^ q2): both XORs of three are formed once — the database's when its tile lands, in place
         * of word 2, the query's here — and x2 = ones ^ x0 ^ x1 makes the majority a function of (x0, x1, ones): four integer-pipe
         * instructions per compare instead of five (the kernel is bound by that pipe: 59 of 108 instructions per offset). */
        if (cta100) {
#pragma unroll
            for (int i = 0; i < CQ; i++) qp[i][2] ^= qp[i][0] ^ qp[i][1];
        }
    }
    float* miss_tab = ratio_tab + 260;                                          /* [101] */
    uint32_t* st_win = reinterpret_cast<uint32_t*>(miss_tab + 104);             /* [2][STAGE_SUBFPS] */
    if (cta100) for (uint32_t m = tid; m <= 100; m += SEARCH_WARPS * 32) miss_tab[m] = __fdiv_rn((float)(100 - m), 100.0f);
    /* (visible to every warp after the first __syncthreads of the tile loop) */

    /* tile = clips [c0, c1) whose subfingerprints [s_lo, s_hi) fit the staging buffer; every thread computes the same bounds */
    const uint32_t uniform_per_tile = uniform ? (uint32_t)STAGE_SUBFPS / uniform : 0u;
    auto first_subfp = [&](uint32_t c) -> uint32_t { return uniform ? c * uniform : offsets[c]; };
    auto tile_end = [&](uint32_t c0) -> uint32_t {
        if (uniform) return min(c_end, c0 + uniform_per_tile);
        const uint32_t s_lo = offsets[c0];
        uint32_t c1 = c0;
        while (c1 < c_end && offsets[c1 + 1] - s_lo <= (uint32_t)STAGE_SUBFPS) c1++;
        return c1;
    };
    auto issue_tile = [&](uint32_t c0, uint32_t c1, int buf) {
        const uint32_t s_lo = first_subfp(c0), n_sub = first_subfp(c1) - s_lo;
        const uint4* src = reinterpret_cast<const uint4*>(db + (size_t)s_lo * 2 * W);
        uint4* dst = reinterpret_cast<uint4*>(st_words + (size_t)buf * STAGE_SUBFPS * 2 * W);
        for (uint32_t i = tid; i < n_sub * (2 * W / 4); i += SEARCH_WARPS * 32) cp_async16(dst + i, src + i);
        if (!MASKED) for (uint32_t i = tid; i < n_sub; i += SEARCH_WARPS * 32) cp_async8(st_meta + (size_t)buf * STAGE_SUBFPS + i, meta + (size_t)s_lo + i);
        cp_async_commit();
    };

    uint32_t c0 = c_begin, c1 = c_begin < c_end ? tile_end(c_begin) : c_begin;
    int buf = 0;
    if (c0 < c_end) issue_tile(c0, c1, 0);
    while (c0 < c_end) {
        cp_async_wait_all();
        __syncthreads();                                                       /* tile `buf` has landed; everyone is done with the other buffer */
        const uint32_t n0 = c1, n1 = n0 < c_end ? tile_end(n0) : n0;
        if (n0 < c_end) issue_tile(n0, n1, buf ^ 1);                           /* prefetch the next tile while this one is compared */
        const uint32_t s_lo = first_subfp(c0);
        const uint32_t* tw = st_words + (size_t)buf * STAGE_SUBFPS * 2 * W;
        const float2* tm = st_meta + (size_t)buf * STAGE_SUBFPS;
        bool tile_reg = db_regular != 0;
        if (!MASKED && !db_regular) {                                          /* (CTA-uniform) is every subfingerprint of this tile regular?  P ^ M = the rank mask, word by word */
            const uint32_t n_sub = first_subfp(c1) - s_lo;
            bool ok = true;
            for (uint32_t j = tid; j < n_sub; j += SEARCH_WARPS * 32) {
#pragma unroll
                for (int w = 0; w < W; w++) ok = ok && ((tw[(size_t)j * 2 * W + w] ^ tw[(size_t)j * 2 * W + W + w]) == mask.w[w]);
            }
            tile_reg = __syncthreads_and(ok) != 0;
        }
        const bool regular = q_regular && tile_reg;
        const bool tile100 = cta100 && tile_reg;                               /* (cta100 implies q_regular in every warp) */
        if (tile100) {                                                         /* windows of this tile: win[j] nibble t = fourth-word bits of subfingerprint j + t */
            const uint32_t n_sub = first_subfp(c1) - s_lo;
            for (uint32_t j = tid; j < n_sub; j += SEARCH_WARPS * 32) {
                uint32_t wv = 0;
#pragma unroll
                for (uint32_t t = 0; t < 8; t++) if (j + t < n_sub) wv |= (tw[(size_t)(j + t) * 2 * W + 3] & 0xFu) << (4 * t);
                st_win[buf * STAGE_SUBFPS + j] = wv;
                uint32_t* own = const_cast<uint32_t*>(tw) + (size_t)j * 2 * W;      /* word 2 <- a0 ^ a1 ^ a2 (nobody else touches this subfingerprint's words 0..2) */
                own[2] ^= own[0] ^ own[1];
            }
            __syncthreads();
        }
        if (qg < n_qgroups) for (uint32_t c = c0 + sub; c < c1; c += rep) {
            uint32_t s0, cnt;                                                 /* warp-uniform */
            if (uniform) { s0 = (c - c0) * uniform; cnt = uniform; }
            else { const uint32_t a = offsets[c]; s0 = a - s_lo; cnt = offsets[c + 1] - a; }
            float best = 0.0f;                                                /* FP.m:133 */
            if (CAN100 && tile100) {
                const uint32_t* src0 = tw + (size_t)s0 * 2 * W;
                const uint32_t* wn = st_win + buf * STAGE_SUBFPS + s0;
                const uint32_t miss_tab_addr = smem_u32(miss_tab);
                for (uint32_t n_off = cnt - CQ + 1; n_off; n_off--, src0 += 2 * W, wn++) {      /* FP.m:136 */
                    uint32_t x = *wn ^ qnib;                                   /* mismatching leftover bits, nibble i belongs to compare i (nibbles >= CQ: never read) */
                    x = x - ((x >> 1) & 0x55555555u);
                    x = (x & 0x33333333u) + ((x >> 2) & 0x33333333u);          /* nibble i: how many of them (0..4) */
                    /* ... spread to bytes, times four: byte t of xe / xo = 4 x the count of compare 2t / 2t + 1 — a compare then takes its
                     * byte with ONE permute, already scaled as the table's byte offset (a shift and a mask per compare before) */
                    const uint32_t xe = (x << 2) & 0x3C3C3C3Cu, xo = (x >> 2) & 0x3C3C3C3Cu;
                    float sum = 0.0f;
#pragma unroll
                    for (int i = 0; i < CQ; i++) {                            /* FP.m:139-142 */
                        const uint4 a = *reinterpret_cast<const uint4*>(src0 + i * 2 * W);      /* a.z = a0 ^ a1 ^ a2 (rewritten when the tile landed) */
                        const uint32_t x0 = a.x ^ qp[i][0], x1 = a.y ^ qp[i][1], ones = a.z ^ qp[i][2];
                        const uint32_t twos = lop3<0xD4>(x0, x1, ones);        /* maj(x0, x1, x2) with x2 = ones ^ x0 ^ x1: (x0 & x1) | ((x0 ^ x1) & ~ones) */
                        const uint32_t nib4 = __byte_perm((i & 1) ? xo : xe, 0u, 0x4440u + (uint32_t)(i >> 1));
                        /* address of miss_tab[popc(ones) + 2 popc(twos) + nibble], every step a multiply-add (FMA pipe) instead of a shift or
                         * a three-input add (integer pipe, the binding one) */
                        uint32_t addr; float r;
                        asm("mad.lo.u32 %0, %1, 1, %2;" : "=r"(addr) : "r"(nib4), "r"(miss_tab_addr));
                        asm("mad.lo.u32 %0, %1, 4, %2;" : "=r"(addr) : "r"(__popc(ones)), "r"(addr));
                        asm("mad.lo.u32 %0, %1, 8, %2;" : "=r"(addr) : "r"(__popc(twos)), "r"(addr));
                        asm("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(addr));      /* (100 - miss) / 100, tabulated with the IEEE divide (filled before the tile loop's first barrier) */
                        sum = i == 0 ? r : __fadd_rn(sum, r);
                    }
                    const float mean = CQ == 1 ? sum : mean_exact<CQ>(sum);
                    best = fmaxf(best, mean);
                }
            }
            else if (regular) {
                const uint32_t* src0 = tw + (size_t)s0 * 2 * W;               /* the clip's subfingerprint o; one pointer step per offset */
                for (uint32_t n_off = cnt - CQ + 1; n_off; n_off--, src0 += 2 * W) {      /* FP.m:136, short form (cnt >= CQ on this path) */
                float sum = 0.0f;
#pragma unroll
                for (int i = 0; i < CQ; i++) {
                    const uint32_t* src = src0 + i * 2 * W;
                    uint32_t h[W];
                    if (W % 4 == 0) {
#pragma unroll
                        for (int w = 0; w < W; w += 4) {
                            const uint4 a = *reinterpret_cast<const uint4*>(src + w);
                            h[w] = lop3<0x82>(a.x, qp[i][w], mask.w[w]); h[w + 1] = lop3<0x82>(a.y, qp[i][w + 1], mask.w[w + 1]);      /* ~(a ^ b) & mask */
                            h[w + 2] = lop3<0x82>(a.z, qp[i][w + 2], mask.w[w + 2]); h[w + 3] = lop3<0x82>(a.w, qp[i][w + 3], mask.w[w + 3]);
                        }
                    } else {
#pragma unroll
                        for (int w = 0; w < W; w += 2) {
                            const uint2 a = *reinterpret_cast<const uint2*>(src + w);
                            h[w] = lop3<0x82>(a.x, qp[i][w], mask.w[w]); h[w + 1] = lop3<0x82>(a.y, qp[i][w + 1], mask.w[w + 1]);
                        }
                    }
                    const float r = ratio_tab[popc_words<W>(h)];               /* hits / pairs, tabulated with the IEEE divide */
                    sum = i == 0 ? r : __fadd_rn(sum, r);                      /* 0 + r = r */
                }
                const float mean = CQ == 1 ? sum : mean_exact<CQ>(sum);
                best = fmaxf(best, mean);                                      /* Apple MAX: no NaN on this path (cnt >= CQ >= 1) */
                }
            }
            else for (uint32_t o = 0; o + CQ <= cnt; o++) {                   /* FP.m:136 */
                float sum = 0.0f;
#pragma unroll
                for (int i = 0; i < CQ; i++) {                                /* FP.m:139-142 */
                    const uint32_t* src = tw + (size_t)(s0 + o + i) * 2 * W;
                    uint32_t p1[W], m1[W];
                    if (W % 4 == 0) {
#pragma unroll
                        for (int w = 0; w < W; w += 4) {
                            const uint4 a = *reinterpret_cast<const uint4*>(src + w), b = *reinterpret_cast<const uint4*>(src + W + w);
                            p1[w] = a.x; p1[w + 1] = a.y; p1[w + 2] = a.z; p1[w + 3] = a.w; m1[w] = b.x; m1[w + 1] = b.y; m1[w + 2] = b.z; m1[w + 3] = b.w;
                        }
                    } else {
#pragma unroll
                        for (int w = 0; w < W; w += 2) {
                            const uint2 a = *reinterpret_cast<const uint2*>(src + w), b = *reinterpret_cast<const uint2*>(src + W + w);
                            p1[w] = a.x; p1[w + 1] = a.y; m1[w] = b.x; m1[w + 1] = b.y;
                        }
                    }
                    float fposs, rcp;
                    if (MASKED) {
                        uint32_t possible = 0;
#pragma unroll
                        for (int w = 0; w < W; w++) { p1[w] &= mask.w[w]; m1[w] &= mask.w[w]; possible += __popc(p1[w] | m1[w]); }   /* FP.m:159-160 */
                        fposs = (float)possible; rcp = c_rcp[possible];
                    } else {
                        const float2 mt = tm[s0 + o + i];
                        fposs = mt.x; rcp = mt.y;
                    }
                    uint32_t h[W];
#pragma unroll
                    for (int w = 0; w < W; w++) {
                        uint32_t qw = qp[i][w];
                        if (CAN100 && w == 2 && cta100) qw ^= qp[i][0] ^ qp[i][1];                          /* (the register holds q0 ^ q1 ^ q2 for the 100-rank form) */
                        h[w] = hit_word2(p1[w], m1[w], qw, qm[i][w]);                                       /* FP.m:162-167 */
                    }
                    sum = __fadd_rn(sum, ratio_exact(popc_words<W>(h), fposs, rcp));
                }
                const float mean = mean_exact<CQ>(sum);                        /* FP.m:144 */
                best = (best < mean) ? mean : best;                            /* Apple MAX */
            }
            if (qvalid) {
                if (all_scores) all_scores[(size_t)q * all_stride + c] = best;
                if (best > worst) { top.insert(best, clip_ids ? clip_ids[c] : clip_base + c); worst = fmaxf(top.worst(), floor_below); }
            }
        }
        c0 = n0; c1 = n1; buf ^= 1;
    }
    if (qvalid) for (int r = 0; r < k; r++) {
        part_sc[(((size_t)chunk * rep + sub) * n_q + q) * k + r] = top.sc[r * 32 + lane];
        part_id[(((size_t)chunk * rep + sub) * n_q + q) * k + r] = top.id[r * 32 + lane];
    }
}

/* Generic path: any query length (uniform over the launch, words staged in shared memory), any clip length,
 * including clips SHORTER than the query, where the reference swaps and the query becomes fp1 (FP.m:123-131). */
template <int W>
__global__ void __launch_bounds__(SEARCH_WARPS * 32)
search_generic_kernel(const uint32_t* __restrict__ db, const uint32_t* __restrict__ offsets, const uint32_t n_clips, const uint32_t clip_base,
                      const uint32_t* __restrict__ clip_ids, const uint32_t* __restrict__ qwords, const uint32_t n_q, const uint32_t cq, const uint32_t pairs, const int k,
                      const uint32_t n_qgroups, const uint32_t clips_per_chunk, float* __restrict__ part_sc, uint32_t* __restrict__ part_id,
                      float* __restrict__ all_scores, const uint32_t total_warps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t gw = blockIdx.x * SEARCH_WARPS + wid;
    if (gw >= total_warps) return;
    const uint32_t qg = gw % n_qgroups, chunk = gw / n_qgroups;
    const uint32_t q = qg * 32 + lane;
    const bool qvalid = q < n_q;
    const size_t per_warp = (size_t)2 * k * 32 + (size_t)cq * 2 * W * 32;      /* 4-byte words */
    uint32_t* base = reinterpret_cast<uint32_t*>(smem_raw) + (size_t)wid * per_warp;
    TopK top{reinterpret_cast<float*>(base), base + (size_t)k * 32, k, lane};
    top.init();
    uint32_t* qs = base + (size_t)2 * k * 32;                                  /* [cq][2W][32] */
    const PairMask<W> mask = make_mask<W>(pairs);
    for (uint32_t i = 0; i < cq; i++)
        for (int w = 0; w < 2 * W; w++) qs[(i * 2 * W + w) * 32 + lane] = qvalid ? (qwords[((size_t)q * cq + i) * 2 * W + w] & mask.w[w % W]) : 0u;
    __syncwarp();
    const uint32_t c_begin = chunk * clips_per_chunk;
    const uint32_t c_end = min(n_clips, c_begin + clips_per_chunk);
    for (uint32_t c = c_begin; c < c_end; c++) {
        const uint32_t s0 = offsets[c], cd = offsets[c + 1] - s0;
        const bool db_first = cd >= cq;                                        /* FP.m:123: swap only if count1 < count2 */
        const uint32_t c1 = db_first ? cd : cq, c2 = db_first ? cq : cd;
        float best = 0.0f;
        for (uint32_t o = 0; o + c2 <= c1; o++) {                              /* FP.m:136 */
            float sum = 0.0f;
            for (uint32_t i = 0; i < c2; i++) {                                /* FP.m:139-142 */
                const uint32_t jd = db_first ? o + i : i, jq = db_first ? i : o + i;
                const uint32_t* src = db + ((size_t)s0 + jd) * 2 * W;
                uint32_t hits = 0, possible = 0;
#pragma unroll
                for (int w = 0; w < W; w++) {
                    const uint32_t dp = __ldg(src + w) & mask.w[w], dm = __ldg(src + W + w) & mask.w[w];
                    const uint32_t xp = qs[(jq * 2 * W + w) * 32 + lane], xm = qs[(jq * 2 * W + W + w) * 32 + lane];
                    if (db_first) { possible += __popc(dp | dm); hits += __popc(hit_word(dp, dm, xp, xm)); }
                    else          { possible += __popc(xp | xm); hits += __popc(hit_word(xp, xm, dp, dm)); }
                }
                const float r = possible ? __fdiv_rn((float)hits, (float)possible) : 0.0f;   /* FP.m:171-175 */
                sum = __fadd_rn(sum, r);
            }
            const float mean = __fdiv_rn(sum, (float)c2);                      /* FP.m:144; c2 == 0 -> NaN */
            best = (best < mean) ? mean : best;                                /* Apple MAX keeps `best` on NaN */
        }
        if (qvalid) {
            if (all_scores) all_scores[(size_t)q * n_clips + c] = best;
            if (best > top.worst()) top.insert(best, clip_ids ? clip_ids[c] : clip_base + c);
        }
    }
    if (qvalid) for (int r = 0; r < k; r++) {
        part_sc[((size_t)chunk * n_q + q) * k + r] = top.sc[r * 32 + lane];
        part_id[((size_t)chunk * n_q + q) * k + r] = top.id[r * 32 + lane];
    }
}

/* One pair of fingerprints, LBAudioDetectiveFingerprintCompareToFingerprint (FP.m:119-149) as written: swap so that fp1 has more
 * subfingerprints, then for every time offset the f32 mean of CompareSubfingerprints over fp2's subfingerprints, and the maximum.
 * The offsets are spread over the threads of one CTA; the maximum of non-NaN means does not depend on the order, and the only NaN
 * (0/0 when fp2 is empty) leaves the match at 0, as Apple's MAX does. */
template <int W>
__global__ void __launch_bounds__(128)
compare_pair_kernel(const uint32_t* __restrict__ wa, const uint32_t ca, const uint32_t* __restrict__ wb, const uint32_t cb, const uint32_t pairs,
                    float* __restrict__ out) {
    const bool swap = ca < cb;                                                  /* FP.m:123-131 */
    const uint32_t* w1 = swap ? wb : wa; const uint32_t* w2 = swap ? wa : wb;
    const uint32_t c1 = swap ? cb : ca, c2 = swap ? ca : cb;
    const PairMask<W> mask = make_mask<W>(pairs);
    float best = 0.0f;                                                          /* FP.m:133 */
    if (c2 > 0) for (uint32_t o = threadIdx.x; o + c2 <= c1; o += blockDim.x) {  /* FP.m:136 */
        float sum = 0.0f;
        for (uint32_t i = 0; i < c2; i++) {                                     /* FP.m:139-142 */
            const uint32_t* s1 = w1 + (size_t)(o + i) * 2 * W; const uint32_t* s2 = w2 + (size_t)i * 2 * W;
            uint32_t hits = 0, possible = 0;
#pragma unroll
            for (int w = 0; w < W; w++) {
                const uint32_t p1 = s1[w] & mask.w[w], m1 = s1[W + w] & mask.w[w];
                possible += __popc(p1 | m1);                                    /* FP.m:159-160 */
                hits += __popc(hit_word(p1, m1, s2[w] & mask.w[w], s2[W + w] & mask.w[w]));   /* FP.m:162-167 */
            }
            sum = __fadd_rn(sum, possible ? __fdiv_rn((float)hits, (float)possible) : 0.0f);   /* FP.m:171-175 */
        }
        const float mean = __fdiv_rn(sum, (float)c2);                           /* FP.m:144 */
        best = (best < mean) ? mean : best;
    }
    __shared__ float red[4];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, d));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) *out = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
}

/* (score desc, clip index asc) strict order; a is "better" than b */
__device__ __forceinline__ bool better(float sa, uint32_t ia, float sb, uint32_t ib) { return sa > sb || (sa == sb && ia < ib); }

/* Few queries (at most FEW_MAX_Q — the server-style "identify one recording" call): lane = database clip instead of lane = query, so
 * that every lane works when there is one query.  A lane walks its clip once; subfingerprint j of the clip meets query subfingerprints
 * i = 0 .. cq - 1 and feeds the running sums of the offsets o = j - i, which therefore receive their terms in the order i = 0, 1, ..
 * of FP.m:139-142 (a ring of cq partial sums; offset j - cq + 1 completes at step j).  Same arithmetic as the other search kernels:
 * IEEE hits / possible, sequential f32 sum, IEEE mean, Apple MAX.  The warp keeps ONE top-k list per query (entry r in lane r): a batch
 * of 32 scores is tested against the list's last entry with one ballot and the few that pass are inserted by shuffles. */
constexpr uint32_t FEW_MAX_Q = 8, FEW_MAX_CQ = 6;       /* beyond 8 queries the query-per-lane kernel (32 queries at the price of one) is the faster one */
template <int W>
__global__ void __launch_bounds__(SEARCH_WARPS * 32)
search_few_kernel(const uint32_t* __restrict__ db, const uint32_t* __restrict__ offsets, const uint32_t n_clips, const uint32_t clip_base,
                  const uint32_t* __restrict__ clip_ids, const uint32_t* __restrict__ qwords, const uint32_t n_q, const uint32_t cq, const uint32_t pairs, const int k,
                  const uint32_t clips_per_warp, float* __restrict__ part_sc, uint32_t* __restrict__ part_id, float* __restrict__ all_scores,
                  const uint32_t total_warps, const int db_regular) {
    const int lane = threadIdx.x & 31;
    const uint32_t gw = blockIdx.x * SEARCH_WARPS + (threadIdx.x >> 5);
    if (gw >= total_warps) return;
    const uint32_t c_begin = gw * clips_per_warp, c_end = min(n_clips, c_begin + clips_per_warp);      /* clips_per_warp is a multiple of 32 */
    const PairMask<W> mask = make_mask<W>(pairs);
    const float fcq = (float)cq;
    __shared__ __align__(16) uint32_t q_smem[SEARCH_WARPS][FEW_MAX_CQ][2 * W];
    uint32_t (*qw)[2 * W] = q_smem[threadIdx.x >> 5];
    for (uint32_t q = 0; q < n_q; q++) {
        /* the query's words sit in shared memory (one copy per warp, read back as broadcasts): in registers they would cost the kernel,
         * which waits on memory, a third of its resident warps */
        __syncwarp();
        for (uint32_t t = lane; t < FEW_MAX_CQ * 2 * W; t += 32) {
            const uint32_t i = t / (2 * W), w = t % (2 * W);
            qw[i][w] = i < cq ? (__ldg(qwords + ((size_t)q * cq + i) * 2 * W + w) & mask.w[w % W]) : 0u;
        }
        __syncwarp();
        /* Regular codes on both sides (one sign bit on every one of the first `pairs` ranks: what extraction produces; the database was
         * checked when it was appended, the query is checked here): M = ~P on those ranks, possible = pairs, and a rank hits iff the P bits
         * agree — the M plane is neither loaded nor compared, `possible` and its reciprocal are constants.  Same scores, bit for bit. */
        bool short_form = db_regular != 0;
        if (short_form) {
            for (uint32_t i = 0; i < cq; i++)
#pragma unroll
                for (int w = 0; w < W; w++) short_form = short_form && ((qw[i][w] ^ qw[i][W + w]) == mask.w[w]);      /* (warp-uniform: every lane reads the same words) */
        }
        const float fpairs = small_uint_to_float(pairs), rcp_pairs = pairs ? __frcp_rn(fpairs) : 0.0f;
        float tsc = -1.0f; uint32_t tid = EMPTY_IDX;                           /* lane r < k: entry r of the warp's list, best first */
        for (uint32_t c0 = c_begin; c0 < c_end; c0 += 32) {
            const uint32_t c = c0 + lane;
            const bool valid = c < c_end;
            float best = 0.0f;                                                  /* FP.m:133 */
            if (valid) {
                const uint32_t s0 = offsets[c], cd = offsets[c + 1] - s0;     /* cd >= cq: the database side is fp1 (FP.m:123) */
                float ring[FEW_MAX_CQ];
#pragma unroll
                for (uint32_t i = 0; i < FEW_MAX_CQ; i++) ring[i] = 0.0f;
                /* the words of subfingerprint j + 1 are requested before subfingerprint j is compared: the kernel waits on memory, not on arithmetic */
                auto load_words = [&](const uint32_t j, uint32_t (&p)[W], uint32_t (&m)[W]) {
                    const uint32_t* src = db + ((size_t)s0 + j) * 2 * W;
                    if constexpr (W >= 4) {                                     /* subfingerprints are 8 W bytes apart: 16-byte aligned planes for W = 4, 8 */
#pragma unroll
                        for (int w = 0; w < W; w += 4) {
                            const uint4 a = __ldg(reinterpret_cast<const uint4*>(src + w));
                            p[w] = a.x; p[w + 1] = a.y; p[w + 2] = a.z; p[w + 3] = a.w;
                            if (!short_form) { const uint4 b = __ldg(reinterpret_cast<const uint4*>(src + W + w)); m[w] = b.x; m[w + 1] = b.y; m[w + 2] = b.z; m[w + 3] = b.w; }
                        }
                    } else {
                        const uint2 a = __ldg(reinterpret_cast<const uint2*>(src));
                        p[0] = a.x; p[1] = a.y;
                        if (!short_form) { const uint2 b = __ldg(reinterpret_cast<const uint2*>(src + W)); m[0] = b.x; m[1] = b.y; }
                    }
                };
                uint32_t np[W], nm[W];
#pragma unroll
                for (int w = 0; w < W; w++) nm[w] = 0u;
                if (cd) load_words(0, np, nm);
                for (uint32_t j = 0; j < cd; j++) {
                    uint32_t dp[W], dm[W], cover[W];
#pragma unroll
                    for (int w = 0; w < W; w++) { dp[w] = np[w]; dm[w] = nm[w]; }
                    if (j + 1 < cd) load_words(j + 1, np, nm);
                    float r[FEW_MAX_CQ];
                    /* subfingerprint j meets query subfingerprint i only where offset j - i exists (0 <= j - i <= cd - cq): at the two ends of a
                     * clip the other pairs belong to no offset — their sums are never read — and are skipped (30 of 114 pairs for 19 against 6) */
                    const uint32_t last_off = cd - cq;
                    if (short_form) {                                           /* warp-uniform */
#pragma unroll
                        for (uint32_t i = 0; i < FEW_MAX_CQ; i++) {
                            r[i] = 0.0f;
                            if (i < cq && j >= i && j - i <= last_off) {
                                uint32_t x[W];
#pragma unroll
                                for (int w = 0; w < W; w++) x[w] = (dp[w] ^ qw[i][w]) & mask.w[w];               /* ranks whose P bits differ */
                                r[i] = ratio_exact(pairs - popc_words<W>(x), fpairs, rcp_pairs);                  /* hits / possible with possible = pairs */
                            }
                        }
                    } else {
#pragma unroll
                    for (int w = 0; w < W; w++) { dp[w] &= mask.w[w]; dm[w] &= mask.w[w]; cover[w] = dp[w] | dm[w]; }
                    const uint32_t possible = popc_words<W>(cover);            /* FP.m:159-160 */
                    const float fposs = small_uint_to_float(possible), rcp = possible ? __frcp_rn(fposs) : 0.0f;
#pragma unroll
                    for (uint32_t i = 0; i < FEW_MAX_CQ; i++) {
                        r[i] = 0.0f;
                        if (i < cq && j >= i && j - i <= last_off) {           /* (uniform over the warp when the clips are equally long) */
                            uint32_t h[W];
#pragma unroll
                            for (int w = 0; w < W; w++) h[w] = hit_word2(dp[w], dm[w], qw[i][w], qw[i][W + w]);  /* FP.m:162-167 */
                            r[i] = ratio_exact(popc_words<W>(h), fposs, rcp);   /* FP.m:171-175 */
                        }
                    }
                    }
                    /* offset j - i gets its term number i: in decreasing i, so that ring[i - 1] is still the sum of terms 0 .. i - 1 */
#pragma unroll
                    for (int i = (int)FEW_MAX_CQ - 1; i >= 1; i--) ring[i] = __fadd_rn(ring[i - 1], r[i]);
                    ring[0] = __fadd_rn(0.0f, r[0]);
                    if (j + 1 >= cq) {                                          /* offset j - cq + 1 is complete (FP.m:144) */
                        float sum = ring[0];
#pragma unroll
                        for (uint32_t i = 1; i < FEW_MAX_CQ; i++) sum = (i + 1 == cq) ? ring[i] : sum;
                        const float mean = __fdiv_rn(sum, fcq);
                        best = (best < mean) ? mean : best;
                    }
                }
                if (all_scores) all_scores[(size_t)q * n_clips + c] = best;
            }
            /* into the warp's list: first those that beat its last entry, one at a time */
            const uint32_t cid = (clip_ids && valid) ? clip_ids[c] : clip_base + c;
            const float last_sc = __shfl_sync(0xffffffffu, tsc, k - 1); const uint32_t last_id = __shfl_sync(0xffffffffu, tid, k - 1);
            uint32_t pass = __ballot_sync(0xffffffffu, valid && (last_id == EMPTY_IDX || better(best, cid, last_sc, last_id)));
            while (pass) {
                const int src = __ffs(pass) - 1; pass &= pass - 1;
                const float cs = __shfl_sync(0xffffffffu, best, src); const uint32_t ci = __shfl_sync(0xffffffffu, cid, src);
                const bool ahead = lane < k && tid != EMPTY_IDX && better(tsc, tid, cs, ci);      /* entries that stay in front of the newcomer: a prefix */
                const int pos = __popc(__ballot_sync(0xffffffffu, ahead));
                const float up_sc = __shfl_up_sync(0xffffffffu, tsc, 1); const uint32_t up_id = __shfl_up_sync(0xffffffffu, tid, 1);
                if (lane < k) { if (lane > pos) { tsc = up_sc; tid = up_id; } else if (lane == pos) { tsc = cs; tid = ci; } }
            }
        }
        if (lane < k) { part_sc[((size_t)gw * n_q + q) * k + lane] = tsc; part_id[((size_t)gw * n_q + q) * k + lane] = tid; }
    }
}

/* One warp per (group of lists, query): k-way merge of the partial top-k lists [list][q][k] of the group into out[group][q][k].  Every
 * list is sorted best first (unused slots last), so the merge is a tournament of list heads: lane l owns lists l, l + 32, .. (at most
 * MERGE_LPL of them) and keeps their heads in registers; a round finds the best head of the warp by shuffles and only the lane that
 * owned it fetches a successor — k rounds of a few dozen instructions, whatever the number of lists (the earlier form rescanned every
 * candidate in every round).  group_size >= n_lists is the plain merge (one group); more than 32 x MERGE_LPL lists (a search with few
 * queries cuts the database into thousands of chunks to fill the device) are merged in two levels.  The order (score desc, clip asc) is
 * strict and every clip sits in exactly one list, so the top k of the groups' top k are the top k of everything. */
constexpr uint32_t MERGE_LPL = 8, MERGE_GROUP = 32 * MERGE_LPL;
__global__ void __launch_bounds__(128)
merge_topk_kernel(const float* __restrict__ part_sc, const uint32_t* __restrict__ part_id, const uint32_t n_lists, const uint32_t n_q, const int k,
                  float* __restrict__ out_sc, uint32_t* __restrict__ out_id, const uint32_t group_size, const size_t list_stride) {
    const int lane = threadIdx.x & 31;
    const uint32_t wq = blockIdx.x * 4 + (threadIdx.x >> 5);
    const uint32_t n_groups = (n_lists + group_size - 1) / group_size;
    if (wq >= n_q * n_groups) return;
    const uint32_t grp = wq / n_q, q = wq % n_q;
    const uint32_t list0 = grp * group_size, lists = (n_lists - list0 < group_size) ? n_lists - list0 : group_size;      /* <= MERGE_GROUP */
    float* o_sc = out_sc + ((size_t)grp * n_q + q) * k; uint32_t* o_id = out_id + ((size_t)grp * n_q + q) * k;
    float hs[MERGE_LPL]; uint32_t hi[MERGE_LPL], pos[MERGE_LPL];                /* head of each of my lists: score, clip (EMPTY_IDX: exhausted), slot */
    auto head = [&](uint32_t j, uint32_t slot, float& sc, uint32_t& id) {
        const size_t at = (size_t)(list0 + lane + 32u * j) * list_stride + (size_t)q * k + slot;
        sc = part_sc[at]; id = part_id[at];
    };
#pragma unroll
    for (uint32_t j = 0; j < MERGE_LPL; j++) {
        hs[j] = -2.0f; hi[j] = EMPTY_IDX; pos[j] = 0;
        if (lane + 32u * j < lists) head(j, 0, hs[j], hi[j]);
    }
    for (int r = 0; r < k; r++) {
        float bs = -2.0f; uint32_t bi = EMPTY_IDX;                            /* my best head */
#pragma unroll
        for (uint32_t j = 0; j < MERGE_LPL; j++) if (hi[j] != EMPTY_IDX && (bi == EMPTY_IDX || better(hs[j], hi[j], bs, bi))) { bs = hs[j]; bi = hi[j]; }
        float ws = bs; uint32_t wi = bi;                                     /* the warp's */
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            const float os = __shfl_xor_sync(0xffffffffu, ws, d); const uint32_t oi = __shfl_xor_sync(0xffffffffu, wi, d);
            if (oi != EMPTY_IDX && (wi == EMPTY_IDX || better(os, oi, ws, wi))) { ws = os; wi = oi; }
        }
        if (lane == 0) { o_sc[r] = wi != EMPTY_IDX ? ws : -1.0f; o_id[r] = wi; }
        if (wi == EMPTY_IDX) { for (int rr = r + 1; rr < k; rr++) if (lane == 0) { o_sc[rr] = -1.0f; o_id[rr] = EMPTY_IDX; } break; }      /* every list exhausted (warp-uniform) */
        if (bi == wi) {                                                      /* mine won (clip ids are unique): its list moves on */
#pragma unroll
            for (uint32_t j = 0; j < MERGE_LPL; j++) if (hi[j] == wi) {
                pos[j]++;
                if (pos[j] < (uint32_t)k) head(j, pos[j], hs[j], hi[j]); else hi[j] = EMPTY_IDX;
            }
        }
    }
}

}  // namespace lbad

/* ================================================================================================== host ==== */

using namespace lbad;

struct lbadcu_db {
    int device = 0; uint32_t W = 4;
    cudaStream_t stream = nullptr;
    uint32_t* d_words = nullptr; float2* d_meta = nullptr; size_t cap_subfps = 0, n_subfps = 0; uint32_t pairs_full = 0;
    std::vector<uint32_t> h_offsets{0};
    uint32_t* d_offsets = nullptr; size_t d_offsets_cap = 0; bool offsets_dirty = true;
    uint32_t min_count = 0xffffffffu, max_count = 0, base = 0;
    bool regular = true; uint32_t* d_irregular = nullptr;    /* device counter of subfingerprints that are not "one sign bit per rank" */
    float* d_part_sc = nullptr; uint32_t* d_part_id = nullptr; size_t part_cap = 0;
    float* d_floor_sc = nullptr; uint32_t* d_floor_id = nullptr; size_t floor_cap = 0;      /* [q][k] top-k of the sample (threshold pass) */
    /* the partial-list buffers are shared by every search of this database: a search on another stream than the previous one waits
     * for that one's merge (searches on one database are serialised on the device, whatever streams they are enqueued on) */
    cudaEvent_t part_done = nullptr; cudaStream_t part_stream = nullptr; bool part_used = false;
    /* global clip ids (sharded databases whose clips are not one contiguous id range): ids[c] for every local clip c, ascending */
    std::vector<uint32_t> h_ids; uint32_t* d_ids = nullptr; size_t d_ids_cap = 0; bool use_ids = false;
    const uint32_t* ids_ptr() const { return use_ids ? d_ids : nullptr; }
    int sm_count = 0; size_t smem_optin = 0;
    uint64_t launches = 0;
    LaunchTimer timer;
};

static int upload_rcp() {
    float h[257]; h[0] = 0.0f;
    for (int p = 1; p <= 256; p++) h[p] = 1.0f / (float)p;
    LBAD_CUDA_TRY(cudaMemcpyToSymbol(c_rcp, h, sizeof h));
    return LBAD_OK;
}

extern "C" int lbadcu_db_create(uint32_t W, uint32_t pairs_full, lbadcu_db** out) {
    *out = nullptr;
    if ((W != 2 && W != 4 && W != 8) || pairs_full == 0 || pairs_full > 32 * W) return LBAD_ERR_ARG;
    if (lbadcu_device_available() != LBAD_OK) { set_error("no CUDA device available (this library has no CPU fallback)"); return LBAD_ERR_NODEVICE; }
    lbadcu_db* db = new lbadcu_db(); db->W = W; db->pairs_full = pairs_full;
    Guard<lbadcu_db> guard(db, lbadcu_db_destroy);
    LBAD_CUDA_TRY(cudaGetDevice(&db->device));
    cudaDeviceProp prop; LBAD_CUDA_TRY(cudaGetDeviceProperties(&prop, db->device));
    db->sm_count = prop.multiProcessorCount; db->smem_optin = prop.sharedMemPerBlockOptin;
    LBAD_CUDA_TRY(cudaStreamCreateWithFlags(&db->stream, cudaStreamNonBlocking));
    LBAD_CUDA_TRY(cudaEventCreateWithFlags(&db->part_done, cudaEventDisableTiming));
    int e = upload_rcp(); if (e != LBAD_OK) return e;       /* per device; cheap enough to repeat per database */
    *out = guard.release();
    return LBAD_OK;
}

extern "C" void lbadcu_db_destroy(lbadcu_db* db) {
    if (!db) return;
    DeviceScope _device_scope(db->device);
    if (db->stream) cudaStreamSynchronize(db->stream);
    db->timer.clear();
    if (db->part_done) { cudaEventSynchronize(db->part_done); cudaEventDestroy(db->part_done); }
    cudaFree(db->d_words); cudaFree(db->d_meta); cudaFree(db->d_irregular); cudaFree(db->d_offsets); cudaFree(db->d_part_sc); cudaFree(db->d_part_id); cudaFree(db->d_ids); cudaFree(db->d_floor_sc); cudaFree(db->d_floor_id);
    if (db->stream) cudaStreamDestroy(db->stream);
    delete db;
}

extern "C" uint32_t lbadcu_db_clips(const lbadcu_db* db) { return (uint32_t)db->h_offsets.size() - 1; }
extern "C" uint64_t lbadcu_db_subfps(const lbadcu_db* db) { return db->n_subfps; }
extern "C" uint32_t lbadcu_db_min_count(const lbadcu_db* db) { return db->min_count; }
extern "C" uint32_t lbadcu_db_max_count(const lbadcu_db* db) { return db->max_count; }
extern "C" void lbadcu_db_set_base(lbadcu_db* db, uint32_t base) { db->base = base; }
extern "C" void* lbadcu_db_stream(lbadcu_db* db) { return db->stream; }
extern "C" uint64_t lbadcu_db_launches(const lbadcu_db* db) { return db->launches; }
extern "C" uint32_t lbadcu_db_timing(lbadcu_db* db, int enable, int reset, double* total_ms) {
    uint32_t n = db->timer.collect(total_ms, reset != 0);
    db->timer.enabled = enable != 0;
    return n;
}

extern "C" int lbadcu_db_append(lbadcu_db* db, const uint32_t* words, int on_device, uint32_t n_clips, const uint32_t* counts, uint32_t uniform,
                                void* producer_stream, int64_t first_global_id) {
    if (!db || (!words && n_clips)) return LBAD_ERR_ARG;
    LBAD_ON_DEVICE(db->device);
    /* global ids: either every clip of the database has one (ascending, so that ties keep the lower id in front) or none has */
    const uint32_t have = lbadcu_db_clips(db);
    if (first_global_id >= 0) {
        if ((have && !db->use_ids) || (uint64_t)first_global_id + n_clips > 0xfffffff0ull) return LBAD_ERR_ARG;
        if (have && (uint64_t)first_global_id <= db->h_ids.back()) return LBAD_ERR_ARG;
    } else if (db->use_ids && n_clips) return LBAD_ERR_ARG;
    if (on_device && producer_stream) {
        /* the words may still be in the making on the caller's stream (an asynchronous extraction): order the copy after it */
        cudaEvent_t ready;
        LBAD_CUDA_TRY(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
        cudaError_t e1 = cudaEventRecord(ready, (cudaStream_t)producer_stream), e2 = e1 == cudaSuccess ? cudaStreamWaitEvent(db->stream, ready, 0) : e1;
        cudaEventDestroy(ready);
        LBAD_CUDA_TRY(e2);
    }
    uint64_t add = 0;
    for (uint32_t c = 0; c < n_clips; c++) add += counts ? counts[c] : uniform;
    if (db->n_subfps + add > 0xfffffff0ull) return LBAD_ERR_ARG;
    if (db->n_subfps + add > db->cap_subfps) {
        size_t cap = std::max<size_t>(db->cap_subfps * 2, db->n_subfps + add);
        uint32_t* nw = nullptr; float2* nm = nullptr;
        LBAD_CUDA_TRY(cudaMalloc(&nw, cap * 2 * db->W * sizeof(uint32_t))); LBAD_CUDA_TRY(cudaMalloc(&nm, cap * sizeof(float2)));
        if (db->n_subfps) {
            LBAD_CUDA_TRY(cudaMemcpyAsync(nw, db->d_words, db->n_subfps * 2 * db->W * sizeof(uint32_t), cudaMemcpyDeviceToDevice, db->stream));
            LBAD_CUDA_TRY(cudaMemcpyAsync(nm, db->d_meta, db->n_subfps * sizeof(float2), cudaMemcpyDeviceToDevice, db->stream));
        }
        LBAD_CUDA_TRY(cudaStreamSynchronize(db->stream));
        cudaFree(db->d_words); cudaFree(db->d_meta); db->d_words = nw; db->d_meta = nm; db->cap_subfps = cap;
    }
    if (add) LBAD_CUDA_TRY(cudaMemcpyAsync(db->d_words + db->n_subfps * 2 * db->W, words, add * 2 * db->W * sizeof(uint32_t),
                                           on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, db->stream));
    if (add) {
        const unsigned blocks = (unsigned)std::min<uint64_t>((add + 255) / 256, 4096);
        if (!db->d_irregular) { LBAD_CUDA_TRY(cudaMalloc(&db->d_irregular, sizeof(uint32_t))); LBAD_CUDA_TRY(cudaMemsetAsync(db->d_irregular, 0, sizeof(uint32_t), db->stream)); }
        if (db->W == 2) meta_kernel<2><<<blocks, 256, 0, db->stream>>>(db->d_words, db->d_meta, db->n_subfps, add, db->pairs_full, db->d_irregular);
        else if (db->W == 4) meta_kernel<4><<<blocks, 256, 0, db->stream>>>(db->d_words, db->d_meta, db->n_subfps, add, db->pairs_full, db->d_irregular);
        else meta_kernel<8><<<blocks, 256, 0, db->stream>>>(db->d_words, db->d_meta, db->n_subfps, add, db->pairs_full, db->d_irregular);
        db->launches++;
        LBAD_CUDA_TRY(cudaGetLastError());
        uint32_t irregular = 0;
        LBAD_CUDA_TRY(cudaMemcpyAsync(&irregular, db->d_irregular, sizeof(uint32_t), cudaMemcpyDeviceToHost, db->stream));
        LBAD_CUDA_TRY(cudaStreamSynchronize(db->stream));
        db->regular = irregular == 0;
    }
    LBAD_CUDA_TRY(cudaStreamSynchronize(db->stream));
    db->h_offsets.reserve(db->h_offsets.size() + n_clips);
    for (uint32_t c = 0; c < n_clips; c++) {
        const uint32_t n = counts ? counts[c] : uniform;
        db->h_offsets.push_back(db->h_offsets.back() + n);
        db->min_count = std::min(db->min_count, n); db->max_count = std::max(db->max_count, n);
    }
    if (first_global_id >= 0 && n_clips) {
        db->use_ids = true;
        for (uint32_t c = 0; c < n_clips; c++) db->h_ids.push_back((uint32_t)first_global_id + c);
    }
    db->n_subfps += add; db->offsets_dirty = true;
    return LBAD_OK;
}

/* copies the packed words and the per-clip counts back to the host (for persisting a database) */
extern "C" int lbadcu_db_download(lbadcu_db* db, uint32_t* h_words, uint32_t* h_counts) {
    if (!db) return LBAD_ERR_ARG;
    LBAD_ON_DEVICE(db->device);
    if (h_words && db->n_subfps) {
        LBAD_CUDA_TRY(cudaMemcpyAsync(h_words, db->d_words, db->n_subfps * 2 * db->W * sizeof(uint32_t), cudaMemcpyDeviceToHost, db->stream));
        LBAD_CUDA_TRY(cudaStreamSynchronize(db->stream));
    }
    if (h_counts) for (size_t c = 0; c + 1 < db->h_offsets.size(); c++) h_counts[c] = db->h_offsets[c + 1] - db->h_offsets[c];
    return LBAD_OK;
}

extern "C" uint64_t lbadcu_db_compares_per_query(const lbadcu_db* db, uint32_t cq) {
    uint64_t total = 0;
    for (size_t c = 0; c + 1 < db->h_offsets.size(); c++) {
        const uint64_t cd = db->h_offsets[c + 1] - db->h_offsets[c];
        const uint64_t c1 = std::max<uint64_t>(cd, cq), c2 = std::min<uint64_t>(cd, cq);
        total += (c1 - c2 + 1) * c2;                                            /* FP.m:136-142 */
    }
    return total;
}

/* n_search: the first n_search clips are searched (the whole database, or the sample of the threshold pass); floor_sc: see the kernel */
template <int W, int CQ>
static void launch_fast(lbadcu_db* db, bool masked, uint32_t n_chunks, size_t smem_topk, cudaStream_t s, const uint32_t* d_q, uint32_t n_q, uint32_t pairs, int k,
                        uint32_t n_qgroups, uint32_t cpc, float* d_all, uint32_t rep, uint32_t n_search, const float* floor_sc) {
    const uint32_t n_clips = n_search, all_stride = lbadcu_db_clips(db);
    const uint32_t total_warps = (n_qgroups + SEARCH_WARPS - 1) / SEARCH_WARPS;      /* = CTAs per clip chunk (passed in the last kernel argument) */
    const uint32_t blocks = n_chunks * total_warps;
    const size_t smem = smem_topk + (size_t)2 * STAGE_SUBFPS * (2 * W * sizeof(uint32_t) + sizeof(float2)) + (260 + 104) * sizeof(float) + 2 * STAGE_SUBFPS * sizeof(uint32_t);
    if (masked) {
        cudaFuncSetAttribute(search_fast_kernel<W, CQ, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        search_fast_kernel<W, CQ, true><<<blocks, SEARCH_WARPS * 32, smem, s>>>(db->d_words, db->d_meta, db->d_offsets, n_clips, db->base, db->ids_ptr(), d_q, n_q, pairs, k, n_qgroups, cpc,
                                                                                db->d_part_sc, db->d_part_id, d_all, total_warps, db->regular ? 1 : 0, rep, all_stride, floor_sc, db->min_count == db->max_count ? db->min_count : 0u);
    } else {
        cudaFuncSetAttribute(search_fast_kernel<W, CQ, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        search_fast_kernel<W, CQ, false><<<blocks, SEARCH_WARPS * 32, smem, s>>>(db->d_words, db->d_meta, db->d_offsets, n_clips, db->base, db->ids_ptr(), d_q, n_q, pairs, k, n_qgroups, cpc,
                                                                                 db->d_part_sc, db->d_part_id, d_all, total_warps, db->regular ? 1 : 0, rep, all_stride, floor_sc, db->min_count == db->max_count ? db->min_count : 0u);
    }
}

template <int W>
static void launch_generic(lbadcu_db* db, uint32_t blocks, size_t smem, cudaStream_t s, const uint32_t* d_q, uint32_t n_q, uint32_t cq, uint32_t pairs, int k,
                           uint32_t n_qgroups, uint32_t cpc, float* d_all, uint32_t total_warps) {
    cudaFuncSetAttribute(search_generic_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    search_generic_kernel<W><<<blocks, SEARCH_WARPS * 32, smem, s>>>(db->d_words, db->d_offsets, lbadcu_db_clips(db), db->base, db->ids_ptr(), d_q, n_q, cq, pairs, k, n_qgroups, cpc,
                                                                     db->d_part_sc, db->d_part_id, d_all, total_warps);
}

extern "C" int lbadcu_db_search_device(lbadcu_db* db, const uint32_t* d_q, uint32_t n_q, uint32_t cq, uint32_t pairs, uint32_t k,
                                       float* d_scores, uint32_t* d_idx, float* d_all, void* stream) {
    if (!db || !d_q || !d_scores || !d_idx || n_q == 0 || k == 0 || k > 64) return LBAD_ERR_ARG;
    LBAD_ON_DEVICE(db->device);
    cudaStream_t s = stream ? (cudaStream_t)stream : db->stream;
    const uint32_t W = db->W, n_clips = lbadcu_db_clips(db);
    if (pairs > 32 * W) pairs = 32 * W;
    if (db->offsets_dirty) {
        if (db->d_offsets_cap < db->h_offsets.size()) {
            cudaFree(db->d_offsets); db->d_offsets = nullptr;
            LBAD_CUDA_TRY(cudaMalloc(&db->d_offsets, db->h_offsets.size() * sizeof(uint32_t)));
            db->d_offsets_cap = db->h_offsets.size();
        }
        LBAD_CUDA_TRY(cudaMemcpyAsync(db->d_offsets, db->h_offsets.data(), db->h_offsets.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
        if (db->use_ids) {
            if (db->d_ids_cap < db->h_ids.size()) {
                cudaFree(db->d_ids); db->d_ids = nullptr;
                LBAD_CUDA_TRY(cudaMalloc(&db->d_ids, db->h_ids.size() * sizeof(uint32_t)));
                db->d_ids_cap = db->h_ids.size();
            }
            LBAD_CUDA_TRY(cudaMemcpyAsync(db->d_ids, db->h_ids.data(), db->h_ids.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
        }
        LBAD_CUDA_TRY(cudaStreamSynchronize(s));
        db->offsets_dirty = false;
    }
    const uint32_t n_qgroups = (n_q + 31) / 32;
    const bool fast = n_clips > 0 && db->min_count >= cq && db->max_count <= (uint32_t)STAGE_SUBFPS && cq >= 1 && cq <= 6;
    const bool few = fast && n_q <= FEW_MAX_Q && k <= 32;                 /* lane = clip: every lane works even for ONE query */
    static const uint32_t warps_per_sm = [] { const char* e = getenv("LBAD_SEARCH_WARPS_PER_SM"); const int v = e ? atoi(e) : 0; return (uint32_t)(v >= 4 && v <= 512 ? v : 128); }();
    const uint32_t rep = (fast && !few && n_qgroups <= 2) ? (uint32_t)SEARCH_WARPS / n_qgroups : 1u;      /* search_fast_kernel: lists per chunk */
    const bool masked = pairs < db->pairs_full;       /* words beyond L are zero already; a mask is only needed for a shorter range */
    const size_t smem_fast = (size_t)SEARCH_WARPS * 2 * k * 32 * 4;
    const size_t smem_gen = (size_t)SEARCH_WARPS * ((size_t)2 * k * 32 + (size_t)cq * 2 * W * 32) * 4;
    if (!fast && smem_gen > db->smem_optin) return LBAD_ERR_ARG;
    /* how the first n_search clips are cut into chunks: ~128 warps per SM, several times what is resident (20 warps of the 95-register
     * CQ = 6 kernel) — the block scheduler hands the next chunk to whichever SM frees a slot, which evens out the SMs (measured on a
     * 125,000-clip shard: 9.33 ms with 32, 8.93 with 128, 9.03 with 384 where the chunk lists start to cost; LBAD_SEARCH_WARPS_PER_SM) */
    struct Cut { uint32_t n_chunks, cpc, n_lists; };
    auto cut = [&](uint32_t n_search) {
        uint32_t n_chunks = ((uint32_t)db->sm_count * warps_per_sm + n_qgroups - 1) / n_qgroups;
        if (few) n_chunks = std::min<uint32_t>((n_search + 31) / 32, (uint32_t)db->sm_count * 32);      /* one list per warp, 32 clips per step */
        if (n_chunks > n_search) n_chunks = n_search;
        if (n_chunks < 1) n_chunks = 1;
        uint32_t cpc = n_search ? (n_search + n_chunks - 1) / n_chunks : 1;
        if (few) cpc = (cpc + 31) & ~31u;
        n_chunks = n_search ? (n_search + cpc - 1) / cpc : 1;
        return Cut{n_chunks, cpc, n_chunks * rep};
    };
    /* threshold pass (large databases, the query-per-lane kernel): the top k of a SAMPLE — the first n_sample clips — give, per query, a
     * score that at least k clips reach; the main pass then lets only clips at or above it into its lists.  Exact: the main pass still
     * visits every clip, the sample included.  Costs n_sample / n_clips more compares; saves the list insertions, which the whole warp
     * walks through whenever any of its 32 queries inserts — with short chunks that was nearly every clip (a third of all stall samples
     * at one subfingerprint per query). */
    const bool no_floor = getenv("LBAD_SEARCH_NO_FLOOR") != nullptr;       /* (tests: the same search with and without the threshold pass) */
    /* sample size: 1/24 of the clips for one or two subfingerprints per query (few compares per clip: the insertions dominate), 1/96 for
     * longer queries; measured with 1,000 queries: one subfingerprint per query 1.11 -> 0.64 ms on 100,000 clips of 5 (with the leaner
     * inner loop), six per query 8.01 -> 7.96 ms on 125,000 clips of 19 */
    const uint32_t n_sample = std::min<uint32_t>(8192u, std::max<uint32_t>(1024u, n_clips / (cq <= 2 ? 24u : 96u)));
    const bool use_floor = fast && !few && !no_floor && n_clips >= 65536u;
    const Cut main_cut = cut(n_clips), pre_cut = use_floor ? cut(n_sample) : Cut{0, 0, 0};
    const uint32_t max_lists = std::max(main_cut.n_lists, pre_cut.n_lists);
    const size_t need = ((size_t)max_lists + (max_lists + MERGE_GROUP - 1) / MERGE_GROUP) * n_q * k;      /* chunk lists + (two-level merge) group lists */
    if (db->part_used && db->part_stream != s) LBAD_CUDA_TRY(cudaStreamWaitEvent(s, db->part_done, 0));      /* the previous search, on another stream, still owns the lists */
    if (db->part_cap < need || (use_floor && db->floor_cap < (size_t)n_q * k)) {
        if (db->part_used) LBAD_CUDA_TRY(cudaEventSynchronize(db->part_done));
        LBAD_CUDA_TRY(cudaStreamSynchronize(s));
        if (db->part_cap < need) {
            cudaFree(db->d_part_sc); cudaFree(db->d_part_id); db->d_part_sc = nullptr; db->d_part_id = nullptr; db->part_cap = 0;
            LBAD_CUDA_TRY(cudaMalloc(&db->d_part_sc, need * sizeof(float))); LBAD_CUDA_TRY(cudaMalloc(&db->d_part_id, need * sizeof(uint32_t)));
            db->part_cap = need;
        }
        if (use_floor && db->floor_cap < (size_t)n_q * k) {
            cudaFree(db->d_floor_sc); cudaFree(db->d_floor_id); db->d_floor_sc = nullptr; db->d_floor_id = nullptr; db->floor_cap = 0;
            LBAD_CUDA_TRY(cudaMalloc(&db->d_floor_sc, (size_t)n_q * k * sizeof(float))); LBAD_CUDA_TRY(cudaMalloc(&db->d_floor_id, (size_t)n_q * k * sizeof(uint32_t)));
            db->floor_cap = (size_t)n_q * k;
        }
    }
    /* one pass: the search kernel over the first n_search clips, then the merge of its chunk lists into o_sc / o_id */
    auto pass = [&](const Cut& c, uint32_t n_search, float* all, const float* floor_sc, float* o_sc, uint32_t* o_id) -> int {
        const uint32_t n_chunks = c.n_chunks, cpc = c.cpc, n_lists = c.n_lists;
        const uint32_t total_warps = n_chunks * n_qgroups;
        const uint32_t blocks = (total_warps + SEARCH_WARPS - 1) / SEARCH_WARPS;
#define LBAD_FAST(WW, CC) launch_fast<WW, CC>(db, masked, n_chunks, smem_fast, s, d_q, n_q, pairs, (int)k, n_qgroups, cpc, all, rep, n_search, floor_sc)
#define LBAD_GEN(WW) launch_generic<WW>(db, blocks, smem_gen, s, d_q, n_q, cq, pairs, (int)k, n_qgroups, cpc, all, total_warps)
        if (few) {
            const uint32_t fblocks = (n_chunks + SEARCH_WARPS - 1) / SEARCH_WARPS;
            if (W == 2) search_few_kernel<2><<<fblocks, SEARCH_WARPS * 32, 0, s>>>(db->d_words, db->d_offsets, n_clips, db->base, db->ids_ptr(), d_q, n_q, cq, pairs, (int)k, cpc, db->d_part_sc, db->d_part_id, all, n_chunks, db->regular ? 1 : 0);
            else if (W == 4) search_few_kernel<4><<<fblocks, SEARCH_WARPS * 32, 0, s>>>(db->d_words, db->d_offsets, n_clips, db->base, db->ids_ptr(), d_q, n_q, cq, pairs, (int)k, cpc, db->d_part_sc, db->d_part_id, all, n_chunks, db->regular ? 1 : 0);
            else search_few_kernel<8><<<fblocks, SEARCH_WARPS * 32, 0, s>>>(db->d_words, db->d_offsets, n_clips, db->base, db->ids_ptr(), d_q, n_q, cq, pairs, (int)k, cpc, db->d_part_sc, db->d_part_id, all, n_chunks, db->regular ? 1 : 0);
        } else if (fast) {
#define LBAD_FAST_W(WW) switch (cq) { case 1: LBAD_FAST(WW, 1); break; case 2: LBAD_FAST(WW, 2); break; case 3: LBAD_FAST(WW, 3); break; \
                                      case 4: LBAD_FAST(WW, 4); break; case 5: LBAD_FAST(WW, 5); break; default: LBAD_FAST(WW, 6); break; }
            if (W == 2) { LBAD_FAST_W(2) } else if (W == 4) { LBAD_FAST_W(4) } else { LBAD_FAST_W(8) }
#undef LBAD_FAST_W
        } else {
            if (W == 2) LBAD_GEN(2); else if (W == 4) LBAD_GEN(4); else LBAD_GEN(8);
        }
#undef LBAD_FAST
#undef LBAD_GEN
        db->launches++;
        LBAD_CUDA_TRY(cudaGetLastError());
        if (n_lists > MERGE_GROUP) {
            /* many lists (few queries): merged group by group, then the groups — behind the chunk lists in the same buffer */
            const uint32_t n_groups = (n_lists + MERGE_GROUP - 1) / MERGE_GROUP;
            float* g_sc = db->d_part_sc + (size_t)n_lists * n_q * k; uint32_t* g_id = db->d_part_id + (size_t)n_lists * n_q * k;
            merge_topk_kernel<<<(n_q * n_groups + 3) / 4, 128, 0, s>>>(db->d_part_sc, db->d_part_id, n_lists, n_q, (int)k, g_sc, g_id, MERGE_GROUP, (size_t)n_q * k);
            merge_topk_kernel<<<(n_q + 3) / 4, 128, 0, s>>>(g_sc, g_id, n_groups, n_q, (int)k, o_sc, o_id, n_groups, (size_t)n_q * k);
            db->launches += 2;
        } else {
            merge_topk_kernel<<<(n_q + 3) / 4, 128, 0, s>>>(db->d_part_sc, db->d_part_id, n_lists, n_q, (int)k, o_sc, o_id, n_lists, (size_t)n_q * k);
            db->launches++;
        }
        LBAD_CUDA_TRY(cudaGetLastError());
        return LBAD_OK;
    };
    db->timer.begin(s);                                                   /* the timed span: both passes with their merges */
    if (use_floor) { const int e = pass(pre_cut, n_sample, nullptr, nullptr, db->d_floor_sc, db->d_floor_id); if (e != LBAD_OK) return e; }
    { const int e = pass(main_cut, n_clips, d_all, use_floor ? db->d_floor_sc : nullptr, d_scores, d_idx); if (e != LBAD_OK) return e; }
    db->timer.end(s);
    LBAD_CUDA_TRY(cudaEventRecord(db->part_done, s));
    db->part_stream = s; db->part_used = true;
    return LBAD_OK;
}

extern "C" int lbadcu_db_search_host(lbadcu_db* db, const uint32_t* h_q, uint32_t n_q, uint32_t cq, uint32_t pairs, uint32_t k,
                                     float* h_scores, uint32_t* h_idx, float* h_all) {
    if (!db || !h_scores || !h_idx || n_q == 0 || k == 0) return LBAD_ERR_ARG;
    LBAD_ON_DEVICE(db->device);
    const uint32_t W = db->W, n_clips = lbadcu_db_clips(db);
    const size_t qn = (size_t)n_q * cq * 2 * W;
    DevBuf<uint32_t> d_q, d_id; DevBuf<float> d_sc, d_all;
    LBAD_CUDA_TRY(d_q.alloc(qn)); LBAD_CUDA_TRY(d_sc.alloc((size_t)n_q * k)); LBAD_CUDA_TRY(d_id.alloc((size_t)n_q * k));
    if (h_all && n_clips) LBAD_CUDA_TRY(d_all.alloc((size_t)n_q * n_clips));
    if (qn) LBAD_CUDA_TRY(cudaMemcpyAsync(d_q, h_q, qn * sizeof(uint32_t), cudaMemcpyHostToDevice, db->stream));
    int e = lbadcu_db_search_device(db, d_q, n_q, cq, pairs, k, d_sc, d_id, d_all, db->stream);
    if (e == LBAD_OK) {
        LBAD_CUDA_TRY(cudaMemcpyAsync(h_scores, d_sc, (size_t)n_q * k * sizeof(float), cudaMemcpyDeviceToHost, db->stream));
        LBAD_CUDA_TRY(cudaMemcpyAsync(h_idx, d_id, (size_t)n_q * k * sizeof(uint32_t), cudaMemcpyDeviceToHost, db->stream));
        if (d_all.p) LBAD_CUDA_TRY(cudaMemcpyAsync(h_all, d_all, (size_t)n_q * n_clips * sizeof(float), cudaMemcpyDeviceToHost, db->stream));
    }
    LBAD_CUDA_TRY(cudaStreamSynchronize(db->stream));
    return e;
}

/* k-way merge of gathered per-shard lists on the device (same kernel as the per-chunk merge) */
extern "C" int lbadcu_merge_topk_host(const float* h_sc, const uint32_t* h_id, uint32_t n_lists, uint32_t n_q, uint32_t k, float* o_sc, uint32_t* o_id) {
    if (!h_sc || !h_id || !o_sc || !o_id || n_lists == 0 || n_q == 0 || k == 0) return LBAD_ERR_ARG;
    if (lbadcu_device_available() != LBAD_OK) { set_error("no CUDA device available (this library has no CPU fallback)"); return LBAD_ERR_NODEVICE; }
    const size_t n = (size_t)n_lists * n_q * k;
    DevBuf<float> d_sc, d_o; DevBuf<uint32_t> d_id, d_oi;
    LBAD_CUDA_TRY(d_sc.alloc(n)); LBAD_CUDA_TRY(d_id.alloc(n)); LBAD_CUDA_TRY(d_o.alloc((size_t)n_q * k)); LBAD_CUDA_TRY(d_oi.alloc((size_t)n_q * k));
    LBAD_CUDA_TRY(cudaMemcpy(d_sc, h_sc, n * 4, cudaMemcpyHostToDevice)); LBAD_CUDA_TRY(cudaMemcpy(d_id, h_id, n * 4, cudaMemcpyHostToDevice));
    merge_topk_kernel<<<(n_q + 3) / 4, 128>>>(d_sc, d_id, n_lists, n_q, (int)k, d_o, d_oi, n_lists, (size_t)n_q * k);
    LBAD_CUDA_TRY(cudaGetLastError());
    LBAD_CUDA_TRY(cudaMemcpy(o_sc, d_o, (size_t)n_q * k * 4, cudaMemcpyDeviceToHost)); LBAD_CUDA_TRY(cudaMemcpy(o_id, d_oi, (size_t)n_q * k * 4, cudaMemcpyDeviceToHost));
    return LBAD_OK;
}

/* same merge on lists that already live on the device (e.g. the output of an NCCL all-gather), enqueued on the caller's stream */
extern "C" int lbadcu_merge_topk_device(const float* d_sc, const uint32_t* d_id, uint32_t n_lists, uint64_t list_stride, uint32_t n_q, uint32_t k, float* d_o_sc, uint32_t* d_o_id, void* stream) {
    if (!d_sc || !d_id || !d_o_sc || !d_o_id || n_lists == 0 || n_q == 0 || k == 0) return LBAD_ERR_ARG;
    if (list_stride == 0) list_stride = (uint64_t)n_q * k;
    merge_topk_kernel<<<(n_q + 3) / 4, 128, 0, (cudaStream_t)stream>>>(d_sc, d_id, n_lists, n_q, (int)k, d_o_sc, d_o_id, n_lists, (size_t)list_stride);
    LBAD_CUDA_TRY(cudaGetLastError());
    return LBAD_OK;
}

/* Pairwise compare with a cached per-thread context (stream, pinned staging, device buffers): the drop-in
 * LBAudioDetectiveFingerprintCompareToFingerprint is called in tight loops by the reference's own tests, so it must not pay for
 * allocations.  One H2D copy, one kernel, one 4-byte D2H copy. */
namespace {
struct PairCtx {
    int device = -1; cudaStream_t stream = nullptr;
    uint32_t* h_stage = nullptr; uint32_t* d_words = nullptr; size_t cap_words = 0;
    float* d_score = nullptr; float* h_score = nullptr;
    void release() {
        if (device < 0) return;
        cudaSetDevice(device);
        if (stream) cudaStreamDestroy(stream);
        cudaFreeHost(h_stage); cudaFree(d_words); cudaFree(d_score); cudaFreeHost(h_score);
        *this = PairCtx();
    }
    ~PairCtx() { /* process teardown: the CUDA context may already be gone; leave the buffers to the driver */ }
};
thread_local PairCtx g_pair;
}

extern "C" int lbadcu_compare_pair_device(uint32_t W, uint32_t pairs, const uint32_t* d_w1, uint32_t c1, const uint32_t* d_w2, uint32_t c2, float* d_out, void* stream) {
    if ((W != 2 && W != 4 && W != 8) || !d_out || (c1 && !d_w1) || (c2 && !d_w2)) return LBAD_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (pairs > 32 * W) pairs = 32 * W;
    if (W == 2) compare_pair_kernel<2><<<1, 128, 0, s>>>(d_w1, c1, d_w2, c2, pairs, d_out);
    else if (W == 4) compare_pair_kernel<4><<<1, 128, 0, s>>>(d_w1, c1, d_w2, c2, pairs, d_out);
    else compare_pair_kernel<8><<<1, 128, 0, s>>>(d_w1, c1, d_w2, c2, pairs, d_out);
    LBAD_CUDA_TRY(cudaGetLastError());
    return LBAD_OK;
}

extern "C" int lbadcu_compare_pair(uint32_t W, uint32_t pairs, const uint32_t* w1, uint32_t c1, const uint32_t* w2, uint32_t c2, float* out) {
    if ((W != 2 && W != 4 && W != 8) || !out || (c1 && !w1) || (c2 && !w2)) return LBAD_ERR_ARG;
    if (lbadcu_device_available() != LBAD_OK) { set_error("no CUDA device available (this library has no CPU fallback)"); return LBAD_ERR_NODEVICE; }
    int dev = 0; LBAD_CUDA_TRY(cudaGetDevice(&dev));
    PairCtx& c = g_pair;
    if (c.device != dev) {
        c.release();
        c.device = dev;
        LBAD_CUDA_TRY(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
        LBAD_CUDA_TRY(cudaMalloc(&c.d_score, sizeof(float))); LBAD_CUDA_TRY(cudaHostAlloc(&c.h_score, sizeof(float), cudaHostAllocDefault));
    }
    const size_t n1 = (size_t)c1 * 2 * W, n2 = (size_t)c2 * 2 * W, need = n1 + n2 ? n1 + n2 : 1;
    if (c.cap_words < need) {
        LBAD_CUDA_TRY(cudaStreamSynchronize(c.stream));
        cudaFreeHost(c.h_stage); cudaFree(c.d_words); c.h_stage = nullptr; c.d_words = nullptr; c.cap_words = 0;
        const size_t cap = need < 4096 ? 4096 : need * 2;
        LBAD_CUDA_TRY(cudaHostAlloc(&c.h_stage, cap * sizeof(uint32_t), cudaHostAllocDefault)); LBAD_CUDA_TRY(cudaMalloc(&c.d_words, cap * sizeof(uint32_t)));
        c.cap_words = cap;
    }
    if (n1) memcpy(c.h_stage, w1, n1 * sizeof(uint32_t));
    if (n2) memcpy(c.h_stage + n1, w2, n2 * sizeof(uint32_t));
    if (n1 + n2) LBAD_CUDA_TRY(cudaMemcpyAsync(c.d_words, c.h_stage, (n1 + n2) * sizeof(uint32_t), cudaMemcpyHostToDevice, c.stream));
    if (pairs > 32 * W) pairs = 32 * W;
    if (W == 2) compare_pair_kernel<2><<<1, 128, 0, c.stream>>>(c.d_words, c1, c.d_words + n1, c2, pairs, c.d_score);
    else if (W == 4) compare_pair_kernel<4><<<1, 128, 0, c.stream>>>(c.d_words, c1, c.d_words + n1, c2, pairs, c.d_score);
    else compare_pair_kernel<8><<<1, 128, 0, c.stream>>>(c.d_words, c1, c.d_words + n1, c2, pairs, c.d_score);
    LBAD_CUDA_TRY(cudaGetLastError());
    LBAD_CUDA_TRY(cudaMemcpyAsync(c.h_score, c.d_score, sizeof(float), cudaMemcpyDeviceToHost, c.stream));
    LBAD_CUDA_TRY(cudaStreamSynchronize(c.stream));
    *out = *c.h_score;
    return LBAD_OK;
}

/* ================================================================================================ group ==== */
/* A database sharded by clip over several GPUs, driven by ONE process (include/LBAudioDetectiveDatabase.h,
 * LBAudioDetectiveDatabaseGroup*).  Every shard is an ordinary lbadcu_db on its own device whose clips carry global ids; a search
 * uploads the query batch to every shard, runs the per-shard top-k kernels concurrently (each on its shard's stream), brings the
 * [query][k] lists to the first shard's device with peer copies over NVLink and merges them there with merge_topk_kernel — the result
 * is what ONE database holding all the clips returns, bit for bit (order: score descending, global clip id ascending).  No NCCL: the
 * exchange is 8 x 80 KB, latency-bound, and a peer copy enqueued behind each shard's kernel is the shortest path. */
struct lbadcu_group {
    std::vector<lbadcu_db*> shards;
    uint32_t W = 4, pairs_full = 0;
    uint64_t next_id = 0;
    /* per shard: query words, [query][k] result lists, an event marking "this shard's lists are on the root device" */
    std::vector<uint32_t*> d_q; std::vector<float*> d_sc; std::vector<uint32_t*> d_id; std::vector<cudaEvent_t> landed;
    size_t q_cap = 0, r_cap = 0;
    float* d_gather_sc = nullptr; uint32_t* d_gather_id = nullptr; float* d_out_sc = nullptr; uint32_t* d_out_id = nullptr;     /* on the root device */
    uint32_t* h_q = nullptr; float* h_sc = nullptr; uint32_t* h_id = nullptr; size_t hq_cap = 0, hr_cap = 0;                    /* pinned staging */
    uint64_t launches = 0;
    double last_ms = 0.0; cudaEvent_t t0 = nullptr, t1 = nullptr;
};

extern "C" void lbadcu_group_destroy(lbadcu_group* g) {
    if (!g) return;
    for (size_t i = 0; i < g->shards.size(); i++) {
        if (!g->shards[i]) continue;
        DeviceScope ds(g->shards[i]->device);
        cudaStreamSynchronize(g->shards[i]->stream);
        if (i < g->d_q.size()) { cudaFree(g->d_q[i]); cudaFree(g->d_sc[i]); cudaFree(g->d_id[i]); if (g->landed[i]) cudaEventDestroy(g->landed[i]); }
        if (i == 0) { cudaFree(g->d_gather_sc); cudaFree(g->d_gather_id); cudaFree(g->d_out_sc); cudaFree(g->d_out_id); if (g->t0) cudaEventDestroy(g->t0); if (g->t1) cudaEventDestroy(g->t1); }
    }
    cudaFreeHost(g->h_q); cudaFreeHost(g->h_sc); cudaFreeHost(g->h_id);
    for (auto* db : g->shards) lbadcu_db_destroy(db);
    delete g;
}

extern "C" int lbadcu_group_create(uint32_t W, uint32_t pairs_full, const int* devices, uint32_t n_shards, lbadcu_group** out) {
    *out = nullptr;
    if (!devices || n_shards == 0 || n_shards > 64) return LBAD_ERR_ARG;
    if (lbadcu_device_available() != LBAD_OK) { set_error("no CUDA device available (this library has no CPU fallback)"); return LBAD_ERR_NODEVICE; }
    int n_dev = 0; LBAD_CUDA_TRY(cudaGetDeviceCount(&n_dev));
    for (uint32_t i = 0; i < n_shards; i++) if (devices[i] < 0 || devices[i] >= n_dev) return LBAD_ERR_ARG;
    lbadcu_group* g = new lbadcu_group(); g->W = W; g->pairs_full = pairs_full;
    Guard<lbadcu_group> guard(g, lbadcu_group_destroy);
    g->shards.assign(n_shards, nullptr); g->d_q.assign(n_shards, nullptr); g->d_sc.assign(n_shards, nullptr); g->d_id.assign(n_shards, nullptr); g->landed.assign(n_shards, nullptr);
    for (uint32_t i = 0; i < n_shards; i++) {
        LBAD_ON_DEVICE(devices[i]);
        int e = lbadcu_db_create(W, pairs_full, &g->shards[i]); if (e != LBAD_OK) return e;
        LBAD_CUDA_TRY(cudaEventCreateWithFlags(&g->landed[i], cudaEventDisableTiming));
        if (i > 0 && devices[i] != devices[0]) {              /* direct peer copies between the shard and the root where the topology allows it */
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[i], devices[0]) == cudaSuccess && can) { cudaError_t pe = cudaDeviceEnablePeerAccess(devices[0], 0); if (pe != cudaSuccess) cudaGetLastError(); }
        }
    }
    { LBAD_ON_DEVICE(devices[0]); LBAD_CUDA_TRY(cudaEventCreate(&g->t0)); LBAD_CUDA_TRY(cudaEventCreate(&g->t1)); }
    *out = guard.release();
    return LBAD_OK;
}

extern "C" uint32_t lbadcu_group_shards(const lbadcu_group* g) { return (uint32_t)g->shards.size(); }
extern "C" lbadcu_db* lbadcu_group_shard(lbadcu_group* g, uint32_t i) { return i < g->shards.size() ? g->shards[i] : nullptr; }
extern "C" int lbadcu_group_shard_device(const lbadcu_group* g, uint32_t i) { return i < g->shards.size() ? g->shards[i]->device : -1; }
extern "C" uint64_t lbadcu_group_clips(const lbadcu_group* g) { uint64_t n = 0; for (auto* db : g->shards) n += lbadcu_db_clips(db); return n; }
extern "C" uint64_t lbadcu_group_next_id(const lbadcu_group* g) { return g->next_id; }
extern "C" uint64_t lbadcu_group_launches(const lbadcu_group* g) { uint64_t n = g->launches; for (auto* db : g->shards) n += db->launches; return n; }
extern "C" double lbadcu_group_last_search_ms(const lbadcu_group* g) { return g->last_ms; }

/* Appends n_clips clips with the next global ids, in contiguous blocks over the shards (block i on shard i), host memory. */
extern "C" int lbadcu_group_append(lbadcu_group* g, const uint32_t* h_words, uint32_t n_clips, const uint32_t* counts, uint32_t uniform) {
    if (!g || (!h_words && n_clips)) return LBAD_ERR_ARG;
    const uint32_t S = (uint32_t)g->shards.size();
    if (g->next_id + n_clips > 0xfffffff0ull) return LBAD_ERR_ARG;
    const uint32_t* w = h_words;
    for (uint32_t i = 0; i < S; i++) {
        const uint32_t lo = (uint32_t)((uint64_t)n_clips * i / S), hi = (uint32_t)((uint64_t)n_clips * (i + 1) / S);
        if (hi == lo) continue;
        uint64_t sub = 0;
        for (uint32_t c = lo; c < hi; c++) sub += counts ? counts[c] : uniform;
        int e = lbadcu_db_append(g->shards[i], w, 0, hi - lo, counts ? counts + lo : nullptr, uniform, nullptr, (int64_t)(g->next_id + lo));
        if (e != LBAD_OK) return e;
        w += sub * 2 * g->W;
    }
    g->next_id += n_clips;
    return LBAD_OK;
}

/* One clip (host memory) with the next global id, to the shard that holds the fewest clips: appended one at a time the shards stay balanced. */
extern "C" int lbadcu_group_append_one(lbadcu_group* g, const uint32_t* h_words, uint32_t count, uint64_t* out_id) {
    if (!g || (!h_words && count)) return LBAD_ERR_ARG;
    if (g->next_id + 1 > 0xfffffff0ull) return LBAD_ERR_ARG;
    size_t best = 0;
    for (size_t i = 1; i < g->shards.size(); i++) if (lbadcu_db_clips(g->shards[i]) < lbadcu_db_clips(g->shards[best])) best = i;
    static const uint32_t none[16] = {0};
    int e = lbadcu_db_append(g->shards[best], count ? h_words : none, 0, 1, &count, 0, nullptr, (int64_t)g->next_id);
    if (e != LBAD_OK) return e;
    if (out_id) *out_id = g->next_id;
    g->next_id++;
    return LBAD_OK;
}

/* Appends clips whose packed words already sit on shard `shard`'s device, with explicit global ids [first_global_id, +n_clips). */
extern "C" int lbadcu_group_append_device(lbadcu_group* g, uint32_t shard, const uint32_t* d_words, uint32_t n_clips, uint32_t uniform, uint64_t first_global_id, void* producer_stream) {
    if (!g || shard >= g->shards.size() || !d_words) return LBAD_ERR_ARG;
    int e = lbadcu_db_append(g->shards[shard], d_words, 1, n_clips, nullptr, uniform, producer_stream, (int64_t)first_global_id);
    if (e == LBAD_OK && first_global_id + n_clips > g->next_id) g->next_id = first_global_id + n_clips;
    return e;
}

extern "C" int lbadcu_group_search_host(lbadcu_group* g, const uint32_t* h_qwords, uint32_t n_q, uint32_t cq, uint32_t pairs, uint32_t k, float* h_scores, uint32_t* h_idx) {
    if (!g || !h_scores || !h_idx || n_q == 0 || k == 0 || k > 64 || (!h_qwords && cq)) return LBAD_ERR_ARG;
    const uint32_t S = (uint32_t)g->shards.size();
    const size_t qn = (size_t)n_q * cq * 2 * g->W, rn = (size_t)n_q * k;
    lbadcu_db* root = g->shards[0];
    /* buffers: grown on demand, kept between calls */
    if (g->hq_cap < qn || g->hr_cap < rn) {
        cudaFreeHost(g->h_q); cudaFreeHost(g->h_sc); cudaFreeHost(g->h_id); g->h_q = nullptr; g->h_sc = nullptr; g->h_id = nullptr; g->hq_cap = g->hr_cap = 0;
        LBAD_CUDA_TRY(cudaHostAlloc(&g->h_q, (qn ? qn : 1) * sizeof(uint32_t), cudaHostAllocPortable));
        LBAD_CUDA_TRY(cudaHostAlloc(&g->h_sc, rn * sizeof(float), cudaHostAllocPortable)); LBAD_CUDA_TRY(cudaHostAlloc(&g->h_id, rn * sizeof(uint32_t), cudaHostAllocPortable));
        g->hq_cap = qn; g->hr_cap = rn;
    }
    if (g->q_cap < qn || g->r_cap < rn) {
        for (uint32_t i = 0; i < S; i++) {
            LBAD_ON_DEVICE(g->shards[i]->device);
            LBAD_CUDA_TRY(cudaStreamSynchronize(g->shards[i]->stream));
            cudaFree(g->d_q[i]); cudaFree(g->d_sc[i]); cudaFree(g->d_id[i]); g->d_q[i] = nullptr; g->d_sc[i] = nullptr; g->d_id[i] = nullptr;
            LBAD_CUDA_TRY(cudaMalloc(&g->d_q[i], (qn ? qn : 1) * sizeof(uint32_t))); LBAD_CUDA_TRY(cudaMalloc(&g->d_sc[i], rn * sizeof(float))); LBAD_CUDA_TRY(cudaMalloc(&g->d_id[i], rn * sizeof(uint32_t)));
            if (i == 0) {
                cudaFree(g->d_gather_sc); cudaFree(g->d_gather_id); cudaFree(g->d_out_sc); cudaFree(g->d_out_id); g->d_gather_sc = nullptr; g->d_gather_id = nullptr; g->d_out_sc = nullptr; g->d_out_id = nullptr;
                LBAD_CUDA_TRY(cudaMalloc(&g->d_gather_sc, S * rn * sizeof(float))); LBAD_CUDA_TRY(cudaMalloc(&g->d_gather_id, S * rn * sizeof(uint32_t)));
                LBAD_CUDA_TRY(cudaMalloc(&g->d_out_sc, rn * sizeof(float))); LBAD_CUDA_TRY(cudaMalloc(&g->d_out_id, rn * sizeof(uint32_t)));
            }
        }
        g->q_cap = qn; g->r_cap = rn;
    }
    if (qn) memcpy(g->h_q, h_qwords, qn * sizeof(uint32_t));
    { LBAD_ON_DEVICE(root->device); LBAD_CUDA_TRY(cudaEventRecord(g->t0, root->stream)); }
    /* 1: the query batch to every shard; 2: every shard's search (kernels of different shards run concurrently); 3: every shard's lists to
     * the root by a peer copy enqueued behind its kernels.  Three passes, so that no shard's launch waits for another shard's copies. */
    for (uint32_t i = 0; i < S; i++) {
        LBAD_ON_DEVICE(g->shards[i]->device);
        if (qn) LBAD_CUDA_TRY(cudaMemcpyAsync(g->d_q[i], g->h_q, qn * sizeof(uint32_t), cudaMemcpyHostToDevice, g->shards[i]->stream));
    }
    for (uint32_t i = 0; i < S; i++) {
        int e = lbadcu_db_search_device(g->shards[i], g->d_q[i], n_q, cq, pairs, k, i == 0 ? g->d_gather_sc : g->d_sc[i], i == 0 ? g->d_gather_id : g->d_id[i], nullptr, g->shards[i]->stream);
        if (e != LBAD_OK) return e;
    }
    for (uint32_t i = 1; i < S; i++) {
        LBAD_ON_DEVICE(g->shards[i]->device);
        LBAD_CUDA_TRY(cudaMemcpyPeerAsync(g->d_gather_sc + (size_t)i * rn, root->device, g->d_sc[i], g->shards[i]->device, rn * sizeof(float), g->shards[i]->stream));
        LBAD_CUDA_TRY(cudaMemcpyPeerAsync(g->d_gather_id + (size_t)i * rn, root->device, g->d_id[i], g->shards[i]->device, rn * sizeof(uint32_t), g->shards[i]->stream));
        LBAD_CUDA_TRY(cudaEventRecord(g->landed[i], g->shards[i]->stream));
    }
    {
        LBAD_ON_DEVICE(root->device);
        for (uint32_t i = 1; i < S; i++) LBAD_CUDA_TRY(cudaStreamWaitEvent(root->stream, g->landed[i], 0));
        const float* src_sc = g->d_gather_sc; const uint32_t* src_id = g->d_gather_id;
        if (S > 1) {
            merge_topk_kernel<<<(n_q + 3) / 4, 128, 0, root->stream>>>(g->d_gather_sc, g->d_gather_id, S, n_q, (int)k, g->d_out_sc, g->d_out_id, S, rn);
            g->launches++;
            LBAD_CUDA_TRY(cudaGetLastError());
            src_sc = g->d_out_sc; src_id = g->d_out_id;
        }
        LBAD_CUDA_TRY(cudaMemcpyAsync(g->h_sc, src_sc, rn * sizeof(float), cudaMemcpyDeviceToHost, root->stream));
        LBAD_CUDA_TRY(cudaMemcpyAsync(g->h_id, src_id, rn * sizeof(uint32_t), cudaMemcpyDeviceToHost, root->stream));
        LBAD_CUDA_TRY(cudaEventRecord(g->t1, root->stream));
        LBAD_CUDA_TRY(cudaStreamSynchronize(root->stream));
        float ms = 0; if (cudaEventElapsedTime(&ms, g->t0, g->t1) == cudaSuccess) g->last_ms = ms;
    }
    memcpy(h_scores, g->h_sc, rn * sizeof(float)); memcpy(h_idx, g->h_id, rn * sizeof(uint32_t));
    return LBAD_OK;
}
