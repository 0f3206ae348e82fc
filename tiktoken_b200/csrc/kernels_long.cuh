// kernels_long.cuh -- pieces longer than SHORT_MAX bytes: find them, then merge them per length class
// (byte_pair_encode dispatch src/lib.rs:198-211; _byte_pair_merge :140-196; _byte_pair_merge_large :47-138).
#pragma once
#include <cooperative_groups.h>

#include "dev_common.cuh"

using namespace b2bpe;
namespace cg = cooperative_groups;

// --------------------------------------------------------------------------------------------
// kernel 2: find pieces longer than SHORT_MAX bytes
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) find_long_kernel(const uint32_t *__restrict__ pbits,
                                                        const uint32_t *__restrict__ psum, long long n_bytes,
                                                        long long n_words, LongQ q, uint32_t *lidx,
                                                        const uint32_t *__restrict__ sbits, Counters *ctr) {
    long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    const uint32_t pw = pbits[w];
    if (pw == 0) return;
    // bit-parallel filter: a start at bit j is "long" iff the 16 bits after it are all zero
    uint32_t cand;
    {
        const uint64_t x = ((uint64_t)pbits[w + 1] << 32) | pw;      // pbits has zeroed tail words
        uint64_t z = ~x;
        z &= z >> 1; z &= z >> 2; z &= z >> 4; z &= z >> 8;           // bit i: x[i .. i+15] are all zero
        cand = pw & (uint32_t)(z >> 1);
        if (sbits) cand &= ~sbits[w];                       // a special token is one piece with a known id, whatever its length
    }
    for (uint32_t mm = cand; mm;) {
        const int j = __ffs(mm) - 1; mm &= mm - 1;
        const long long s = w * 32 + j;
        if (s >= n_bytes) break;
        const uint32_t m = (j == 31) ? 0u : (pw & ~((2u << j) - 1u));     // piece starts after j in this word
        long long nxt;
        if (m) nxt = w * 32 + (__ffs(m) - 1);
        else {
            long long w2 = w + 1;
            uint32_t x = pbits[w2];
            if (x == 0) {
                // finish the current group of 32 words, then hop over whole groups via the summary bitmap
                // (the sentinel bit at n_bytes guarantees termination)
                uint32_t sm = ((w2 & 31) == 31) ? 0u : (psum[w2 >> 5] & ~((2u << (w2 & 31)) - 1u));
                long long grp = w2 >> 5;
                while (sm == 0) { grp++; sm = psum[grp]; }
                w2 = grp * 32 + (__ffs(sm) - 1);
                x = pbits[w2];
            }
            nxt = w2 * 32 + (__ffs(x) - 1);
        }
        const long long len = nxt - s;
        if (len > SHORT_MAX) {
            // one queue-slot atomic per warp iteration (the lanes that found a long piece together)
            const uint32_t peers = __activemask();
            unsigned int i = 0;
            if ((threadIdx.x & 31) == __ffs(peers) - 1) i = atomicAdd(&ctr->n_long, (unsigned int)__popc(peers));
            i = __shfl_sync(peers, i, __ffs(peers) - 1) + __popc(peers & ((1u << (threadIdx.x & 31)) - 1u));
            // tokens land in ltok at the piece's own byte offset (tokens <= bytes, pieces are disjoint);
            // only pieces beyond the shared-memory path need a region of the global merge scratch.  The scratch
            // is sized from experience, not for the worst case: a piece that does not fit raises ERR_LONGCAP and is
            // left out of its work list (long_bytes keeps counting, so the host re-runs the batch with the exact size).
            unsigned long long off = 0; bool fits = true;
            if (len > LONG_SCRATCH_MIN) {
                off = atomicAdd(&ctr->long_bytes, (unsigned long long)len);
                if (off + (unsigned long long)len > q.scratch_cap) { fits = false; atomicOr(&ctr->err, ERR_LONGCAP); }
            }
            q.start[i] = (unsigned long long)s; q.len[i] = (unsigned int)len; q.off[i] = off; q.ntok[i] = 0;
            lidx[s >> 4] = i;
            // per-class work list, one atomic per (warp iteration, class)
            const int c = !fits ? N_CLS : len > CLUSTER_MIN ? CLS_CLUSTER : len > BLOCK_MIN ? CLS_BLOCK : len > GROUP_MAX ? CLS_WARP
                          : len > 256 ? CLS_G1024 : len > 128 ? 3 : len > 64 ? 2 : len > 32 ? 1 : 0;
            const uint32_t same = __match_any_sync(peers, c);
            unsigned int k = 0;
            if ((threadIdx.x & 31) == __ffs(same) - 1) k = atomicAdd(&ctr->n_cls[c], (unsigned int)__popc(same));
            k = __shfl_sync(peers, k, __ffs(same) - 1) + __popc(same & ((1u << (threadIdx.x & 31)) - 1u));
            if (c == N_CLS) continue;
            // (constant indices: a dynamically indexed kernel parameter would be copied to local memory by every thread)
            unsigned int *lst = c == 0 ? q.cls[0] : c == 1 ? q.cls[1] : c == 2 ? q.cls[2] : c == 3 ? q.cls[3] : c == 4 ? q.cls[4]
                                : c == 5 ? q.cls[5] : c == 6 ? q.cls[6] : q.cls[7];
            lst[k] = i;
        }
    }
}

// --------------------------------------------------------------------------------------------
// kernel 3: long pieces, one warp per piece.
//
// Exact parallel form of the reference's merge loop (src/lib.rs:47-138 / :140-196).  A round takes
// the current global minimum rank g.  Sequentially the reference would merge the g-pairs left to
// right (ties break leftmost; a merged pair destroys an overlapping g-pair to its right, hence the
// alternating selection inside a chain of overlapping candidates).  All of them are merged in ONE
// round, except that the sequential order is only guaranteed while no merge creates a new pair of
// rank < g; the round therefore commits the selected merges up to and including the first one
// that does ("violation"), and the next round continues from the exact sequential state.
// State lives in global scratch (L2 resident): parts as dense arrays id[], rk[] (rank of the pair
// starting at that part), double buffered for the per-round compaction.
// --------------------------------------------------------------------------------------------
__device__ uint32_t long_piece_warp(const DevTables &T, const uint8_t *__restrict__ piece, uint32_t n,
                                    LongScratch S, uint32_t *__restrict__ out, uint32_t *err) {
    const int lane = threadIdx.x & 31;
    // whole-piece probe (src/lib.rs:367-368)
    if (n <= T.max_token_len) {
        uint32_t r = RANK_MAX;
        if (lane == 0) {
            if (n <= (uint32_t)SHORT_MAX) {
                uint64_t k0 = 0, k1 = 0;
                for (uint32_t i = 0; i < n; i++) {
                    if (i < 8) k0 |= (uint64_t)piece[i] << (8 * i); else k1 |= (uint64_t)piece[i] << (8 * (i - 8));
                }
                r = piece_lookup16(T, k0, k1, n);
            } else {
                uint64_t h = long_hash_init(n);
                for (uint32_t i = 0; i < n; i += 8) {
                    uint64_t w = 0;
                    for (uint32_t k = 0; k < 8 && i + k < n; k++) w |= (uint64_t)piece[i + k] << (8 * k);
                    h = long_hash_step(h, w, i / 8);
                }
                r = piece_lookup_long(T, h, n, [&](uint32_t i) { return piece[i]; });
            }
        }
        r = __shfl_sync(0xFFFFFFFFu, r, 0);
        if (r != RANK_MAX) { if (lane == 0) out[0] = r; return 1; }
    }
    if (n == 1) {
        uint32_t id = T.byte_id[piece[0]];
        if (lane == 0) { out[0] = id; if (id >= PSEUDO_BASE) atomicOr(err, ERR_NOBYTE); }
        return 1;
    }
    uint32_t *id = S.idA, *rk = S.rkA, *id2 = S.idB, *rk2 = S.rkB;
    for (uint32_t i = lane; i < n; i += 32) {
        uint32_t b = piece[i];
        id[i] = __ldg(T.byte_id + b);
        rk[i] = (i + 1 < n) ? __ldg(T.pair2 + ((b << 8) | piece[i + 1])) : RANK_MAX;
    }
    __syncwarp();
    uint32_t m = n;
    for (;;) {
        // A. global minimum rank
        uint32_t g = RANK_MAX;
        for (uint32_t i = lane; i < m; i += 32) g = min(g, rk[i]);
        g = warp_min_u32(g);
        if (g == RANK_MAX) break;
        // B. select: odd positions (1st, 3rd, ...) inside each chain of consecutive candidates
        uint32_t carry_par = 0;                            // parity of the candidate run ending before this tile
        for (uint32_t base = 0; base < m; base += 32) {
            uint32_t i = base + lane;
            bool cand = i < m && rk[i] == g;
            uint32_t c = __ballot_sync(0xFFFFFFFFu, cand);
            uint32_t zeros_below = ~c & ((1u << lane) - 1u);
            uint32_t before;                               // candidates immediately before lane, mod 2
            if (zeros_below == 0) before = (uint32_t)lane + carry_par;
            else before = (uint32_t)lane - (32u - (uint32_t)__clz((int)zeros_below));
            bool sel = cand && ((before & 1u) == 0);
            if (i < m) S.flag[i] = sel ? 1 : 0;
            if (c == 0xFFFFFFFFu) carry_par = carry_par;   // 32 more candidates: parity unchanged
            else carry_par = (uint32_t)__clz((int)~c) & 1u;
        }
        __syncwarp();
        // C. new neighbour ranks of every selected merge, first violation
        uint32_t vmin = RANK_MAX;
        for (uint32_t i = lane; i < m; i += 32) {
            if (!S.flag[i]) continue;
            uint32_t nl = RANK_MAX, nr = RANK_MAX;
            if (i >= 1 && i + 2 < m) {                      // both neighbour probes in flight together
                const uint32_t lid = (i >= 2 && S.flag[i - 2]) ? g : id[i - 1];
                pair_lookup2(T, lid, g, g, id[i + 2], nl, nr);
            } else if (i >= 1) {
                nl = pair_lookup(T, (i >= 2 && S.flag[i - 2]) ? g : id[i - 1], g);
            } else if (i + 2 < m) {
                nr = pair_lookup(T, g, id[i + 2]);
            }
            S.aux1[i] = nl; S.aux2[i] = nr;
            if (nl < g || nr < g) vmin = min(vmin, i);
        }
        uint32_t v = warp_min_u32(vmin);
        __syncwarp();
        // D. commit merges at positions <= v, compact into the other buffer
        uint32_t outn = 0;
        for (uint32_t base = 0; base < m; base += 32) {
            uint32_t i = base + lane;
            bool in = i < m;
            bool com = in && S.flag[i] && i <= v;
            bool absorbed = in && i >= 1 && S.flag[i - 1] && (i - 1) <= v;
            bool survive = in && !absorbed;
            uint32_t sb = __ballot_sync(0xFFFFFFFFu, survive);
            if (survive) {
                uint32_t nid, nrk;
                if (com) {
                    nid = g;
                    bool com2 = (i + 2 < m) && S.flag[i + 2] && (i + 2) <= v;
                    nrk = com2 ? S.aux1[i + 2] : S.aux2[i];
                } else {
                    nid = id[i];
                    bool com1 = (i + 1 < m) && S.flag[i + 1] && (i + 1) <= v;
                    nrk = com1 ? S.aux1[i + 1] : rk[i];
                }
                uint32_t o = outn + __popc(sb & ((1u << lane) - 1u));
                id2[o] = nid; rk2[o] = nrk;
            }
            outn += __popc(sb);
        }
        __syncwarp();
        m = outn;
        uint32_t *t1 = id; id = id2; id2 = t1;
        uint32_t *t2 = rk; rk = rk2; rk2 = t2;
    }
    bool bad = false;
    for (uint32_t i = lane; i < m; i += 32) {
        uint32_t x = id[i];
        out[i] = x;
        bad |= x >= PSEUDO_BASE;
    }
    if (__any_sync(0xFFFFFFFFu, bad) && lane == 0) atomicOr(err, ERR_NOBYTE);
    return m;
}

static const int LONG_WARPS = 8;               // warps per block of long_piece_kernel

__global__ void __launch_bounds__(LONG_WARPS * 32) long_piece_kernel(const uint8_t *__restrict__ text, DevTables T, LongQ q,
                                                                    LongScratch S, uint32_t *ltok, Counters *ctr, int cls) {
    const int lane = threadIdx.x & 31;
    const unsigned int n_long = ctr->n_cls[cls];
    const unsigned int *list = cls == CLS_WARP ? q.cls[CLS_WARP] : q.cls[CLS_G1024];
    for (;;) {
        unsigned int k = 0;
        if (lane == 0) k = atomicAdd(&ctr->cls_head[cls], 1u);
        k = __shfl_sync(0xFFFFFFFFu, k, 0);
        if (k >= n_long) break;
        const unsigned int i = list[k];
        const unsigned long long off = q.off[i], st0 = q.start[i];
        LongScratch P = S;
        P.idA += off; P.rkA += off; P.idB += off; P.rkB += off; P.aux1 += off; P.aux2 += off; P.flag += off;
        const uint32_t nt = long_piece_warp(T, text + st0, q.len[i], P, ltok + st0, &ctr->err);
        if (lane == 0) long_piece_done(q, i, nt);
        __syncwarp();
    }
}

// --------------------------------------------------------------------------------------------
// kernel 3b: giant pieces (> BLOCK_MIN bytes: "x"*1_000_000, long whitespace / separator runs).
// Same round-synchronous algorithm as long_piece_warp, executed by a whole 1024-thread block:
// every phase walks the parts in tiles of 1024 with warp ballots and a small cross-warp carry.
// --------------------------------------------------------------------------------------------
static const int GIANT_THREADS = 1024;

__device__ __forceinline__ uint32_t block_min_u32(uint32_t v, uint32_t *s_red) {
    v = warp_min_u32(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t r = s_red[threadIdx.x & 31];
    r = warp_min_u32(r);
    return r;
}

__device__ uint32_t long_piece_block(const DevTables &T, const uint8_t *__restrict__ piece, uint32_t n,
                                     LongScratch S, uint32_t *__restrict__ out, uint32_t *err) {
    __shared__ uint32_t s_red[32];
    __shared__ uint32_t s_wmask[32];     // per-warp candidate / survivor ballots of the current tile
    __shared__ uint32_t s_carry;         // parity carry (select) or running output offset (compaction)
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (n <= T.max_token_len) {          // whole-piece probe (only a token of that length could match)
        uint32_t r = RANK_MAX;
        if (tid == 0) {
            uint64_t h = long_hash_init(n);
            for (uint32_t i = 0; i < n; i += 8) {
                uint64_t w = 0;
                for (uint32_t k = 0; k < 8 && i + k < n; k++) w |= (uint64_t)piece[i + k] << (8 * k);
                h = long_hash_step(h, w, i / 8);
            }
            r = piece_lookup_long(T, h, n, [&](uint32_t i) { return piece[i]; });
            s_red[0] = r;
        }
        __syncthreads();
        r = s_red[0];
        __syncthreads();
        if (r != RANK_MAX) { if (tid == 0) out[0] = r; return 1; }
    }
    uint32_t *id = S.idA, *rk = S.rkA, *id2 = S.idB, *rk2 = S.rkB;
    for (uint32_t i = tid; i < n; i += GIANT_THREADS) {
        uint32_t b = piece[i];
        id[i] = __ldg(T.byte_id + b);
        rk[i] = (i + 1 < n) ? __ldg(T.pair2 + ((b << 8) | piece[i + 1])) : RANK_MAX;
    }
    __syncthreads();
    uint32_t m = n;
    for (;;) {
        // A. global minimum rank
        uint32_t g = RANK_MAX;
        for (uint32_t i = tid; i < m; i += GIANT_THREADS) g = min(g, rk[i]);
        g = block_min_u32(g, s_red);
        if (g == RANK_MAX) break;
        // B. select alternate members of every chain of consecutive candidates
        if (tid == 0) s_carry = 0;
        __syncthreads();
        for (uint32_t base = 0; base < m; base += GIANT_THREADS) {
            const uint32_t i = base + tid;
            const bool cand = i < m && rk[i] == g;
            const uint32_t c = __ballot_sync(0xFFFFFFFFu, cand);
            if (lane == 0) s_wmask[wid] = c;
            __syncthreads();
            // parity of the candidate run that ends right before this warp's first lane
            uint32_t par = 0; bool open = true;
            for (int w = wid - 1; w >= 0 && open; w--) {
                const uint32_t cw = s_wmask[w];
                if (cw == 0xFFFFFFFFu) continue;               // 32 more candidates: parity unchanged
                par = (uint32_t)__clz((int)~cw) & 1u; open = false;
            }
            if (open) par = s_carry;                            // run reaches back into the previous tile
            const uint32_t zeros_below = ~c & ((1u << lane) - 1u);
            uint32_t before;
            if (zeros_below == 0) before = (uint32_t)lane + par;
            else before = (uint32_t)lane - (32u - (uint32_t)__clz((int)zeros_below));
            if (i < m) S.flag[i] = (cand && ((before & 1u) == 0)) ? 1 : 0;
            __syncthreads();
            if (tid == GIANT_THREADS - 1) {                     // carry for the next tile
                uint32_t par2 = s_carry; bool open2 = true;
                for (int w = 31; w >= 0 && open2; w--) {
                    const uint32_t cw = s_wmask[w];
                    if (cw == 0xFFFFFFFFu) continue;
                    par2 = (uint32_t)__clz((int)~cw) & 1u; open2 = false;
                }
                s_carry = par2;
            }
            __syncthreads();
        }
        // C. new neighbour ranks of the selected merges, first violation
        uint32_t vmin = RANK_MAX;
        for (uint32_t i = tid; i < m; i += GIANT_THREADS) {
            if (!S.flag[i]) continue;
            uint32_t nl = RANK_MAX, nr = RANK_MAX;
            if (i >= 1) nl = pair_lookup(T, (i >= 2 && S.flag[i - 2]) ? g : id[i - 1], g);
            if (i + 2 < m) nr = pair_lookup(T, g, id[i + 2]);
            S.aux1[i] = nl; S.aux2[i] = nr;
            if (nl < g || nr < g) vmin = min(vmin, i);
        }
        const uint32_t v = block_min_u32(vmin, s_red);
        __syncthreads();
        // D. commit merges at positions <= v, compact into the other buffer
        if (tid == 0) s_carry = 0;
        __syncthreads();
        for (uint32_t base = 0; base < m; base += GIANT_THREADS) {
            const uint32_t i = base + tid;
            const bool in = i < m;
            const bool com = in && S.flag[i] && i <= v;
            const bool absorbed = in && i >= 1 && S.flag[i - 1] && (i - 1) <= v;
            const bool survive = in && !absorbed;
            const uint32_t sb = __ballot_sync(0xFFFFFFFFu, survive);
            if (lane == 0) s_wmask[wid] = (uint32_t)__popc(sb);
            __syncthreads();
            uint32_t wbase = s_carry;
            for (int w = 0; w < wid; w++) wbase += s_wmask[w];
            if (survive) {
                uint32_t nid, nrk;
                if (com) {
                    nid = g;
                    const bool com2 = (i + 2 < m) && S.flag[i + 2] && (i + 2) <= v;
                    nrk = com2 ? S.aux1[i + 2] : S.aux2[i];
                } else {
                    nid = id[i];
                    const bool com1 = (i + 1 < m) && S.flag[i + 1] && (i + 1) <= v;
                    nrk = com1 ? S.aux1[i + 1] : rk[i];
                }
                const uint32_t o = wbase + __popc(sb & ((1u << lane) - 1u));
                id2[o] = nid; rk2[o] = nrk;
            }
            __syncthreads();
            if (tid == 0) { uint32_t t = s_carry; for (int w = 0; w < 32; w++) t += s_wmask[w]; s_carry = t; }
            __syncthreads();
        }
        m = s_carry;
        __syncthreads();
        uint32_t *t1 = id; id = id2; id2 = t1;
        uint32_t *t2 = rk; rk = rk2; rk2 = t2;
    }
    bool bad = false;
    for (uint32_t i = tid; i < m; i += GIANT_THREADS) {
        uint32_t x = id[i];
        out[i] = x;
        bad |= x >= PSEUDO_BASE;
    }
    if (bad) atomicOr(err, ERR_NOBYTE);
    return m;
}

__global__ void __launch_bounds__(GIANT_THREADS) giant_piece_kernel(const uint8_t *__restrict__ text, DevTables T, LongQ q,
                                                                    LongScratch S, uint32_t *ltok, Counters *ctr) {
    __shared__ unsigned int s_i;
    const unsigned int n_giant = ctr->n_cls[CLS_BLOCK];
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_i = atomicAdd(&ctr->cls_head[CLS_BLOCK], 1u);
        __syncthreads();
        if (s_i >= n_giant) break;
        const unsigned int i = q.cls[CLS_BLOCK][s_i];
        const unsigned long long off = q.off[i];
        LongScratch P = S;
        P.idA += off; P.rkA += off; P.idB += off; P.rkB += off; P.aux1 += off; P.aux2 += off; P.flag += off;
        const uint32_t nt = long_piece_block(T, text + q.start[i], q.len[i], P, ltok + q.start[i], &ctr->err);
        if (threadIdx.x == 0) long_piece_done(q, i, nt);
    }
}

// --------------------------------------------------------------------------------------------
// kernel 3a: mid-size pieces (17..256 bytes: CJK runs, indentation, separators, long words), ONE
// PIECE PER LANE.  find_long_kernel sorts the pieces into length classes of capacity 32 / 64 / 128 /
// 256 parts; a warp merges 32 pieces of one class at a time, walking one convergent instruction
// stream (merge_short_conv for 32, merge_mid_conv above that): the cost of a merge round is shared
// by 32 pieces instead of being paid per piece as in the warp-per-piece kernels.
// State: two [CAP][32] shared-memory columns per warp (id, rank) + the two-level minimum, conflict
// free for any per-lane index.  One launch serves the four classes, longest first: a block owns
// MID_SMEM_BYTES of columns, enough for 256 / CAP warps of a class, and moves to the next class (a
// block-local barrier, no kernel boundary) when the class's work list is drained.
// --------------------------------------------------------------------------------------------
static const int MID_WARPS = 8;                                  // 8 x 32 = 256 parts x 32 lanes per block
static const size_t MID_SMEM_BYTES = (size_t)2 * (256 + 256 / MID_G) * 32 * sizeof(uint32_t);

template <int CAP>
__device__ void mid_class(const uint8_t *__restrict__ text, const DevTables &T, const LongQ &q, int cls, uint32_t *ltok,
                          Counters *ctr, uint32_t *s_cols) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp >= 256 / CAP) return;                             // the block's columns hold 256 / CAP warps of this class
    const int per_warp = CAP == 32 ? 2 * CAP * 32 : 2 * (CAP + CAP / MID_G) * 32;
    uint32_t *base = s_cols + (size_t)warp * per_warp;
    SmemCol32 id{base + lane}, rk{base + CAP * 32 + lane};
    SmemCol32 gmin{base + 2 * CAP * 32 + lane}, gpos{base + (2 * CAP + CAP / MID_G) * 32 + lane};
    const unsigned int n_items = ctr->n_cls[cls];
    const unsigned int *list = q.cls[cls];
    for (;;) {
        unsigned int k0 = 0;
        if (lane == 0) k0 = atomicAdd(&ctr->cls_head[cls], 32u);
        k0 = __shfl_sync(0xFFFFFFFFu, k0, 0);
        if (k0 >= n_items) break;
        const bool have = k0 + lane < n_items;
        unsigned int qi = 0; unsigned long long st = 0; int n = 0;
        if (have) { qi = list[k0 + lane]; st = q.start[qi]; n = (int)q.len[qi]; }
        const uint8_t *piece = text + st;
        uint32_t *out = ltok + st;
        int n_max = (int)__reduce_max_sync(0xFFFFFFFFu, (unsigned)n);
        // stage the bytes (column rk doubles as the byte buffer until the merge initialises it)
        for (int j = 0; j < n_max; j++) rk[j] = j < n ? (uint32_t)piece[j] : 0u;
        // whole-piece probe (src/lib.rs:367-368): only a token of exactly this length can match
        if (have && (uint32_t)n <= T.max_token_len && T.n_long_tokens) {
            uint64_t h = long_hash_init((uint64_t)n);
            for (int i = 0; i < n; i += 8) {
                uint64_t w = 0;
                for (int k = 0; k < 8 && i + k < n; k++) w |= (uint64_t)rk[i + k] << (8 * k);
                h = long_hash_step(h, w, (uint32_t)(i >> 3));
            }
            const uint32_t r = piece_lookup_long(T, h, (uint32_t)n, [&](uint32_t i) { return piece[i]; });
            if (r != RANK_MAX) { out[0] = r; long_piece_done(q, qi, 1); n = 0; }
        }
        n_max = (int)__reduce_max_sync(0xFFFFFFFFu, (unsigned)n);
        if (n_max) {
            uint32_t c = 0; bool bad = false;
            if (CAP == 32) {
                const uint32_t mask = merge_short_conv<32>(T, [&](int j) { return rk[j]; }, n, n_max, 0xFFFFFFFFu, id, rk);
                for (uint32_t mm = mask; mm;) {
                    const int j = __ffs(mm) - 1; mm &= mm - 1;
                    const uint32_t x = id[j];
                    out[c++] = x; bad |= x >= PSEUDO_BASE;
                }
            } else {
                merge_mid_conv(T, n, n_max, 0xFFFFFFFFu, id, rk, gmin, gpos);
                for (int j = 0; j < n; j++) {
                    const uint32_t x = id[j];
                    if (x != ID_DEAD) { out[c++] = x; bad |= x >= PSEUDO_BASE; }
                }
            }
            if (n) long_piece_done(q, qi, c);
            if (bad) atomicOr(&ctr->err, ERR_NOBYTE);
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(MID_WARPS * 32) mid_thread_kernel(const uint8_t *__restrict__ text, DevTables T, LongQ q,
                                                                   uint32_t *ltok, Counters *ctr) {
    extern __shared__ uint32_t s_cols[];
    mid_class<256>(text, T, q, 3, ltok, ctr, s_cols);
    __syncthreads();
    mid_class<128>(text, T, q, 2, ltok, ctr, s_cols);
    __syncthreads();
    mid_class<64>(text, T, q, 1, ltok, ctr, s_cols);
    __syncthreads();
    mid_class<32>(text, T, q, 0, ltok, ctr, s_cols);
}

// --------------------------------------------------------------------------------------------
// kernel 3c: pieces beyond CLUSTER_MIN bytes ("x"*1_000_000, a 256 KiB run of spaces): one THREAD-BLOCK CLUSTER
// (8 CTAs x 1024 threads on 8 SMs of one GPC) per piece.  Same round-synchronous exact merge as long_piece_block;
// CTA r owns a contiguous range of the parts, the carries that the block version passes from tile to tile
// (candidate-run parity of the selection, output offset of the compaction) and the two reductions (minimum rank,
// first violation) cross CTA boundaries through distributed shared memory: every CTA publishes a small summary in
// its own shared memory, cluster.sync(), and reads its peers' summaries with ld.shared::cluster.  The part arrays
// themselves stay in global scratch (L2); cluster.sync() orders those writes between the CTAs.
// --------------------------------------------------------------------------------------------
static const int CLUSTER_CTAS = 8;

struct ClusterShared {
    uint32_t gmin;            // A: minimum rank of the range
    uint32_t sel_par;         // B: parity of the range's trailing candidate run (computed with carry-in 0)
    uint32_t sel_all;         //    the whole range consists of candidates
    uint32_t vmin;            // C: first violation in the range
    uint32_t n_surv;          // D: survivors of the range
    uint32_t next_item;       // work queue index (CTA 0)
    uint32_t probe;           // whole-piece probe result (CTA 0)
};

// Every CTA owns a contiguous range of whole 1024-part tiles and every WARP a contiguous 1/32 of it (a multiple of 32
// parts), which it walks 32 parts at a time with ballots and a carry in a register -- no block barrier inside a pass.
// The carries (run parity of the selection, output offset of the compaction) enter a warp from its predecessors through
// one exchange per pass: warp summaries in shared memory, CTA summaries in the peers' shared memory (DSMEM).
__device__ uint32_t long_piece_cluster(cg::cluster_group &cluster, ClusterShared &sh, const DevTables &T,
                                       const uint8_t *__restrict__ piece, uint32_t n, LongScratch S,
                                       uint32_t *__restrict__ out, uint32_t *err) {
    __shared__ uint32_t s_red[32];
    __shared__ uint32_t s_wpar[32], s_wall[32], s_wcnt[32];      // per-warp summaries of the current pass
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t rank = cluster.block_rank();
    const uint32_t gtid = rank * GIANT_THREADS + tid, gthreads = CLUSTER_CTAS * GIANT_THREADS;
    if (n <= T.max_token_len) {          // whole-piece probe (only a token of that length could match): CTA 0, thread 0
        if (rank == 0 && tid == 0) {
            uint64_t h = long_hash_init(n);
            for (uint32_t i = 0; i < n; i += 8) {
                uint64_t w = 0;
                for (uint32_t k = 0; k < 8 && i + k < n; k++) w |= (uint64_t)piece[i + k] << (8 * k);
                h = long_hash_step(h, w, i / 8);
            }
            sh.probe = piece_lookup_long(T, h, n, [&](uint32_t i) { return piece[i]; });
        }
        cluster.sync();
        const uint32_t r = *cluster.map_shared_rank(&sh.probe, 0);
        cluster.sync();
        if (r != RANK_MAX) { if (rank == 0 && tid == 0) out[0] = r; return 1; }
    }
    uint32_t *id = S.idA, *rk = S.rkA, *id2 = S.idB, *rk2 = S.rkB;
    for (uint32_t i = gtid; i < n; i += gthreads) {
        const uint32_t b = piece[i];
        id[i] = __ldg(T.byte_id + b);
        rk[i] = (i + 1 < n) ? __ldg(T.pair2 + ((b << 8) | piece[i + 1])) : RANK_MAX;
    }
    cluster.sync();
    uint32_t m = n;
    for (;;) {
        const uint32_t chunk = (((m + CLUSTER_CTAS - 1) / CLUSTER_CTAS) + (GIANT_THREADS - 1)) & ~(uint32_t)(GIANT_THREADS - 1);
        const uint32_t lo = min(rank * chunk, m), hi = min(lo + chunk, m);
        const uint32_t wchunk = chunk / 32;                              // a multiple of 32 parts per warp
        const uint32_t wlo = min(lo + (uint32_t)wid * wchunk, hi), whi = min(wlo + wchunk, hi);
        // A. global minimum rank
        uint32_t g = RANK_MAX;
        for (uint32_t i = lo + tid; i < hi; i += GIANT_THREADS) g = min(g, rk[i]);
        g = block_min_u32(g, s_red);
        if (tid == 0) sh.gmin = g;
        cluster.sync();
        g = RANK_MAX;
#pragma unroll
        for (int c = 0; c < CLUSTER_CTAS; c++) g = min(g, *cluster.map_shared_rank(&sh.gmin, c));
        if (g == RANK_MAX) break;                        // the same value in every CTA
        // B. select alternate members of every chain of consecutive candidates.
        //    B1: every warp summarises its sub-range: parity of its trailing candidate run (carry-in 0), all candidates?
        {
            uint32_t par = 0; bool all = (whi - wlo) == wchunk;
            for (uint32_t base = wlo; base < whi; base += 32) {
                const uint32_t i = base + lane;
                const uint32_t c = __ballot_sync(0xFFFFFFFFu, i < whi && rk[i] == g);
                if (c != 0xFFFFFFFFu) { par = (uint32_t)__clz((int)~c) & 1u; all = false; }   // 32 more candidates: parity unchanged
            }
            if (lane == 0) { s_wpar[wid] = par; s_wall[wid] = all ? 1u : 0u; }
            __syncthreads();
            if (tid == 0) {
                uint32_t cp = 0, ca = 1;
                for (int w = 31; w >= 0; w--) { if (!s_wall[w]) { cp = s_wpar[w]; ca = 0; break; } }
                sh.sel_par = cp; sh.sel_all = ca;
            }
            cluster.sync();
        }
        //    B2: carry-in = parity of the candidate run that ends right before the sub-range, then the selection itself
        {
            uint32_t carry = 0; bool found = false;
            for (int w = wid - 1; w >= 0 && !found; w--) if (!s_wall[w]) { carry = s_wpar[w]; found = true; }   // full sub-ranges are even
            for (int c = (int)rank - 1; c >= 0 && !found; c--)
                if (!*cluster.map_shared_rank(&sh.sel_all, c)) { carry = *cluster.map_shared_rank(&sh.sel_par, c); found = true; }
            for (uint32_t base = wlo; base < whi; base += 32) {
                const uint32_t i = base + lane;
                const bool cand = i < whi && rk[i] == g;
                const uint32_t c = __ballot_sync(0xFFFFFFFFu, cand);
                const uint32_t zeros_below = ~c & ((1u << lane) - 1u);
                uint32_t before;
                if (zeros_below == 0) before = (uint32_t)lane + carry;
                else before = (uint32_t)lane - (32u - (uint32_t)__clz((int)zeros_below));
                if (i < whi) S.flag[i] = (cand && ((before & 1u) == 0)) ? 1 : 0;
                if (c != 0xFFFFFFFFu) carry = (uint32_t)__clz((int)~c) & 1u;
            }
            cluster.sync();
        }
        // C. new neighbour ranks of the selected merges, first violation
        uint32_t vmin = RANK_MAX;
        for (uint32_t i = lo + tid; i < hi; i += GIANT_THREADS) {
            if (!S.flag[i]) continue;
            uint32_t nl = RANK_MAX, nr = RANK_MAX;
            if (i >= 1) nl = pair_lookup(T, (i >= 2 && S.flag[i - 2]) ? g : id[i - 1], g);
            if (i + 2 < m) nr = pair_lookup(T, g, id[i + 2]);
            S.aux1[i] = nl; S.aux2[i] = nr;
            if (nl < g || nr < g) vmin = min(vmin, i);
        }
        vmin = block_min_u32(vmin, s_red);
        if (tid == 0) sh.vmin = vmin;
        cluster.sync();
        uint32_t v = RANK_MAX;
#pragma unroll
        for (int c = 0; c < CLUSTER_CTAS; c++) v = min(v, *cluster.map_shared_rank(&sh.vmin, c));
        // D. commit merges at positions <= v, compact into the other buffer.
        //    D1: survivors per warp sub-range -> per CTA -> exchanged
        {
            uint32_t cnt = 0;
            for (uint32_t base = wlo; base < whi; base += 32) {
                const uint32_t i = base + lane;
                const bool absorbed = i < whi && i >= 1 && S.flag[i - 1] && (i - 1) <= v;
                cnt += (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, i < whi && !absorbed));
            }
            if (lane == 0) s_wcnt[wid] = cnt;
            __syncthreads();
            if (tid == 0) { uint32_t t = 0; for (int w = 0; w < 32; w++) t += s_wcnt[w]; sh.n_surv = t; }
            cluster.sync();
        }
        //    D2: every warp writes its survivors at its scanned offset
        uint32_t total = 0;
        {
            uint32_t o = 0;
#pragma unroll
            for (int c = 0; c < CLUSTER_CTAS; c++) {
                const uint32_t x = *cluster.map_shared_rank(&sh.n_surv, c);
                if (c < (int)rank) o += x;
                total += x;
            }
            for (int w = 0; w < wid; w++) o += s_wcnt[w];
            for (uint32_t base = wlo; base < whi; base += 32) {
                const uint32_t i = base + lane;
                const bool in = i < whi;
                const bool com = in && S.flag[i] && i <= v;
                const bool absorbed = in && i >= 1 && S.flag[i - 1] && (i - 1) <= v;
                const bool survive = in && !absorbed;
                const uint32_t sb = __ballot_sync(0xFFFFFFFFu, survive);
                if (survive) {
                    uint32_t nid, nrk;
                    if (com) {
                        nid = g;
                        const bool com2 = (i + 2 < m) && S.flag[i + 2] && (i + 2) <= v;
                        nrk = com2 ? S.aux1[i + 2] : S.aux2[i];
                    } else {
                        nid = id[i];
                        const bool com1 = (i + 1 < m) && S.flag[i + 1] && (i + 1) <= v;
                        nrk = com1 ? S.aux1[i + 1] : rk[i];
                    }
                    const uint32_t oo = o + __popc(sb & ((1u << lane) - 1u));
                    id2[oo] = nid; rk2[oo] = nrk;
                }
                o += (uint32_t)__popc(sb);
            }
        }
        cluster.sync();                                  // the compacted arrays are complete and visible
        m = total;
        uint32_t *t1 = id; id = id2; id2 = t1;
        uint32_t *t2 = rk; rk = rk2; rk2 = t2;
    }
    bool bad = false;
    for (uint32_t i = gtid; i < m; i += gthreads) {
        const uint32_t x = id[i];
        out[i] = x;
        bad |= x >= PSEUDO_BASE;
    }
    if (bad) atomicOr(err, ERR_NOBYTE);
    return m;
}

__global__ void __cluster_dims__(CLUSTER_CTAS, 1, 1) __launch_bounds__(GIANT_THREADS)
    cluster_piece_kernel(const uint8_t *__restrict__ text, DevTables T, LongQ q, LongScratch S, uint32_t *ltok, Counters *ctr) {
    __shared__ ClusterShared sh;
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t rank = cluster.block_rank();
    const unsigned int n_items = ctr->n_cls[CLS_CLUSTER];
    for (;;) {
        if (rank == 0 && threadIdx.x == 0) sh.next_item = atomicAdd(&ctr->cls_head[CLS_CLUSTER], 1u);
        cluster.sync();
        const unsigned int k = *cluster.map_shared_rank(&sh.next_item, 0);
        cluster.sync();                                  // nobody reads CTA 0's slot after this point
        if (k >= n_items) break;
        const unsigned int i = q.cls[CLS_CLUSTER][k];
        const unsigned long long off = q.off[i];
        LongScratch P = S;
        P.idA += off; P.rkA += off; P.idB += off; P.rkB += off; P.aux1 += off; P.aux2 += off; P.flag += off;
        const uint32_t nt = long_piece_cluster(cluster, sh, T, text + q.start[i], q.len[i], P, ltok + q.start[i], &ctr->err);
        if (rank == 0 && threadIdx.x == 0) long_piece_done(q, i, nt);
        cluster.sync();
    }
}
