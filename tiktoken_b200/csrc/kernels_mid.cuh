// kernels_mid.cuh -- pieces of 17..256 bytes (CJK runs, indentation, separators, long identifiers): a GROUP OF LANES
// per piece.  The literal loop of `_byte_pair_merge` (src/lib.rs:140-196): take the smallest rank (strict `<` =>
// leftmost on ties), merge, re-rank the two neighbouring pairs -- one merge per round.
//
// What bounds this stage is the length of a round (a chain of dependent shared-memory accesses + one L2 round trip
// for the pair-table probes), not bandwidth.  mid_thread_kernel gives a whole piece to ONE lane: its round scans the
// piece's group minima serially (up to 32 dependent loads) and a 129..256-byte class leaves 7 of a block's 8 warps
// without column space.  Here a piece of up to 16*G bytes is spread over G = 2 / 4 / 8 / 16 lanes, 16 parts per lane:
//   * every lane keeps the minimum of its own 16 keys (16 independent shared loads), the group minimum is a
//     log2(G)-step shuffle butterfly -- keys are (rank << 8 | position), so the minimum IS the leftmost smallest rank;
//   * the neighbours of the merged pair come from per-lane 16-bit live masks (register shuffles, no dead-byte walks);
//   * the two pair-table probes of a round run on two different lanes at the same time;
//   * the state of a warp is 4 KiB of shared memory whatever the class (key[16][32] + id[16][32]), so every class runs
//     at the same occupancy (48 warps per SM) and a warp steps through the classes on its own, without block barriers.
// 32 / G pieces share one instruction stream per warp (convergent code: a group that has finished idles along).
// Keys pack the rank into 24 bits: vocabularies with ranks of 2^24 and above keep using mid_thread_kernel.
#pragma once
#include "dev_common.cuh"

using namespace b2bpe;

static const int MIDG_WARPS = 4;                 // warps per block (independent of each other)
static const int MIDG_E = 16;                    // parts per lane
static const uint32_t MIDG_NONE = 0xFFFFFFFFu;   // "no mergeable pair here" (above every key)
static const uint32_t MIDG_MAX_RANK = 1u << 24;

struct MidGSmem {
    uint32_t key[MIDG_E * 32];                   // [slot][lane]: rank of (this part, next part) << 8 | position, or MIDG_NONE
    uint32_t id[MIDG_E * 32];                    // [slot][lane]: token id of the part that starts here
};

template <int G>
__device__ void midg_class(const uint8_t *__restrict__ text, const DevTables &T, const LongQ &q, int cls, uint32_t *ltok,
                           Counters *ctr, MidGSmem &S) {
    const int lane = threadIdx.x & 31;
    const int gl = lane & (G - 1);               // lane within the group
    const int gb = lane & ~(G - 1);              // first lane of the group
    const int P = 32 / G;                        // pieces per warp pass
    const uint32_t gmask_bits = G == 32 ? 0xFFFFFFFFu : ((1u << G) - 1u);
    const unsigned int n_items = ctr->n_cls[cls];
    const unsigned int *list = q.cls[cls];
    uint32_t *const keyc = S.key + lane, *const idc = S.id + lane;      // this lane's columns: [slot * 32]
    for (;;) {
        unsigned int k0 = 0;
        if (lane == 0) k0 = atomicAdd(&ctr->cls_head[cls], (unsigned int)P);
        k0 = __shfl_sync(0xFFFFFFFFu, k0, 0);
        if (k0 >= n_items) break;
        const unsigned int item = k0 + (unsigned int)(lane / G);
        const bool have = item < n_items;
        unsigned int qi = 0; unsigned long long st = 0; int n = 0;
        if (have) { qi = list[item]; st = q.start[qi]; n = (int)q.len[qi]; }
        const uint8_t *piece = text + st;
        uint32_t *out = ltok + st;
        // ---- this lane's 16 bytes (+ the first byte of the next lane's), packed in two 64-bit words -----------
        const int j0 = gl * MIDG_E;
        uint64_t w0 = 0, w1 = 0; uint32_t b16 = 0;
#pragma unroll
        for (int s = 0; s < 8; s++) {
            if (j0 + s < n) w0 |= (uint64_t)__ldg(piece + j0 + s) << (8 * s);
            if (j0 + 8 + s < n) w1 |= (uint64_t)__ldg(piece + j0 + 8 + s) << (8 * s);
        }
        if (j0 + MIDG_E < n) b16 = (uint32_t)__ldg(piece + j0 + MIDG_E);
        auto byte_of = [&](int s) -> uint32_t {               // s is a compile-time constant after unrolling
            return s < 8 ? (uint32_t)(w0 >> (8 * s)) & 0xFFu : s < 16 ? (uint32_t)(w1 >> (8 * (s - 8))) & 0xFFu : b16;
        };
        // ---- whole-piece probe (src/lib.rs:367-368): only a token of exactly this length can match --------------
        if (T.n_long_tokens) {
            uint64_t hsh = 0;
            if ((uint32_t)n <= T.max_token_len) {
                if (j0 < n) hsh ^= long_hash_word(w0, (uint32_t)(2 * gl));
                if (j0 + 8 < n) hsh ^= long_hash_word(w1, (uint32_t)(2 * gl + 1));
            }
#pragma unroll
            for (int o = 1; o < G; o <<= 1) hsh ^= __shfl_xor_sync(0xFFFFFFFFu, hsh, o);
            uint32_t r = RANK_MAX;
            if (have && gl == 0 && (uint32_t)n <= T.max_token_len)
                r = piece_lookup_long(T, hsh ^ long_hash_init((uint64_t)n), (uint32_t)n, [&](uint32_t i) { return piece[i]; });
            r = __shfl_sync(0xFFFFFFFFu, r, gb);
            if (r != RANK_MAX) { if (gl == 0) { out[0] = r; q.ntok[qi] = 1; } n = 0; }
        }
        // ---- initial parts: one per byte (8 at a time: the table loads of a block are issued before its stores) --
        uint32_t live = 0;
#pragma unroll
        for (int hh = 0; hh < MIDG_E; hh += 8) {
            uint32_t iv[8], kv[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int s = hh + k;
                iv[k] = 0; kv[k] = MIDG_NONE;
                if (j0 + s < n) {
                    iv[k] = __ldg(T.byte_id + byte_of(s));
                    if (j0 + s + 1 < n) {
                        const uint32_t r = __ldg(T.pair2 + (byte_of(s) << 8 | byte_of(s + 1)));
                        if (r != RANK_MAX) kv[k] = (r << 8) | (uint32_t)(j0 + s);
                    }
                    live |= 1u << s;
                }
            }
#pragma unroll
            for (int k = 0; k < 8; k++) { idc[(hh + k) * 32] = iv[k]; keyc[(hh + k) * 32] = kv[k]; }
        }
        __syncwarp();
        // ---- merge rounds -----------------------------------------------------------------------------------
        for (;;) {
            uint32_t lmin = MIDG_NONE;
#pragma unroll
            for (int s = 0; s < MIDG_E; s++) lmin = min(lmin, keyc[s * 32]);
            uint32_t gmin = lmin;
#pragma unroll
            for (int o = 1; o < G; o <<= 1) gmin = min(gmin, __shfl_xor_sync(0xFFFFFFFFu, gmin, o));
            const bool act = gmin != MIDG_NONE;
            if (!__any_sync(0xFFFFFFFFu, act)) break;
            const uint32_t best = gmin >> 8;                          // rank of the merged token == its id
            const int bj = (int)(gmin & 0xFFu), wl = bj >> 4, ws = bj & 15;
            const uint32_t ne = (__ballot_sync(0xFFFFFFFFu, live != 0) >> gb) & gmask_bits;      // lanes of the group with live parts
            const uint32_t lw = __shfl_sync(0xFFFFFFFFu, live, gb + wl);
            // the part right of bj (it exists: key[bj] is the rank of (bj, next)), the one after it, the one before bj
            int l2 = wl, s2 = 0, l3 = -1, s3 = 0, lp = -1, sp = 0;
            {
                uint32_t m = lw & ~((2u << ws) - 1u) & 0xFFFFu;
                if (!m) { const uint32_t nm = ne & ~((2u << wl) - 1u); l2 = nm ? __ffs((int)nm) - 1 : wl; }
                const uint32_t lv2 = __shfl_sync(0xFFFFFFFFu, live, gb + l2);
                if (!m) m = lv2;
                s2 = m ? __ffs((int)m) - 1 : 0;
                uint32_t m3 = lv2 & ~((2u << s2) - 1u) & 0xFFFFu;
                l3 = l2;
                if (!m3) { const uint32_t nm = ne & ~((2u << l2) - 1u); l3 = nm ? __ffs((int)nm) - 1 : -1; }
                const uint32_t lv3 = __shfl_sync(0xFFFFFFFFu, live, gb + (l3 < 0 ? 0 : l3));
                if (!m3 && l3 >= 0) m3 = lv3;
                if (l3 >= 0) s3 = __ffs((int)m3) - 1;
                uint32_t mp = lw & ((1u << ws) - 1u);
                lp = wl;
                if (!mp) { const uint32_t nm = ne & ((1u << wl) - 1u); lp = nm ? 31 - __clz((int)nm) : -1; }
                const uint32_t lvp = __shfl_sync(0xFFFFFFFFu, live, gb + (lp < 0 ? 0 : lp));
                if (!mp && lp >= 0) mp = lvp;
                if (lp >= 0) sp = 31 - __clz((int)mp);
            }
            const bool need_r = act && l3 >= 0, need_l = act && lp >= 0;
            // the two neighbour probes (src/lib.rs:178-194) on two lanes of the group at once
            uint32_t a = 0, b = 0; bool pr = false;
            if (gl == 0 && need_r) { a = best; b = S.id[s3 * 32 + gb + l3]; pr = true; }
            if (gl == 1 && need_l) { a = S.id[sp * 32 + gb + lp]; b = best; pr = true; }
            uint32_t r = RANK_MAX;
            {
                uint32_t sidx = pair_hash(a, b) & T.pair_mask;
                while (__any_sync(0xFFFFFFFFu, pr)) {
                    if (pr) {
                        const U4 e0 = B2_LDG_U4(T.pair_tab + 2 * sidx), e1 = B2_LDG_U4(T.pair_tab + 2 * sidx + 1);
                        if (e0.x == a && e0.y == b) { r = e0.z; pr = false; }
                        else if (e1.x == a && e1.y == b) { r = e1.z; pr = false; }
                        else if (e1.x == 0xFFFFFFFFu) pr = false;
                        else sidx = (sidx + 1) & T.pair_mask;
                    }
                }
            }
            const uint32_t r_left = __shfl_sync(0xFFFFFFFFu, r, gb + (G > 1 ? 1 : 0));
            if (act) {
                if (gl == 0) {                                         // one lane of the group commits the merge
                    S.id[ws * 32 + gb + wl] = best;
                    S.key[ws * 32 + gb + wl] = (need_r && r != RANK_MAX) ? ((r << 8) | (uint32_t)bj) : MIDG_NONE;
                    S.key[s2 * 32 + gb + l2] = MIDG_NONE;              // the right part disappears
                    if (need_l) S.key[sp * 32 + gb + lp] = r_left != RANK_MAX ? ((r_left << 8) | (uint32_t)(lp * MIDG_E + sp)) : MIDG_NONE;
                }
                if (gl == l2) live &= ~(1u << s2);
            }
            __syncwarp();
        }
        // ---- tokens: the live parts, left to right ------------------------------------------------------------
        {
            const uint32_t c = (uint32_t)__popc(live);
            uint32_t inc = c;
#pragma unroll
            for (int o = 1; o < G; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (gl >= o) inc += y; }
            const uint32_t total = __shfl_sync(0xFFFFFFFFu, inc, gb + G - 1);
            uint32_t o = inc - c; bool bad = false;
            for (uint32_t m = live; m;) {
                const int s = __ffs((int)m) - 1; m &= m - 1;
                const uint32_t x = idc[s * 32];
                out[o++] = x; bad |= x >= PSEUDO_BASE;
            }
            if (bad) atomicOr(&ctr->err, ERR_NOBYTE);
            if (have && n && gl == 0) q.ntok[qi] = total;
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(MIDG_WARPS * 32, 12) mid_group_kernel(const uint8_t *__restrict__ text, DevTables T, LongQ q,
                                                                       uint32_t *ltok, Counters *ctr) {
    __shared__ MidGSmem smem[MIDG_WARPS];
    MidGSmem &S = smem[threadIdx.x >> 5];
    midg_class<16>(text, T, q, 3, ltok, ctr, S);
    midg_class<8>(text, T, q, 2, ltok, ctr, S);
    midg_class<4>(text, T, q, 1, ltok, ctr, S);
    midg_class<2>(text, T, q, 0, ltok, ctr, S);
}
