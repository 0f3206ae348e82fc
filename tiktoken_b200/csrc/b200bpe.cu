// b200bpe.cu -- sm_100a kernels and the C ABI of libb200bpe.so (see include/b200bpe.h).
//
// Path replaced: CoreBPE::encode_ordinary / CoreBPE::encode (src/lib.rs:360-442) and the per-call
// thread pool that fans documents out to it (tiktoken/core.py:164-206).  One call encodes the
// whole batch:
//
//   mark_docs_kernel      doc_off[] -> doc-start bitmask D + first-doc-per-span index
//   pretok_kernel<PAT>    UTF-8 bytes + D -> piece-start bitmask P   (bit-parallel regex rules)
//   find_long_kernel      P -> queue of pieces longer than 16 bytes + one work list per length class
//   mid_thread_kernel     17..256 bytes: one piece per lane, 32 pieces per warp in one convergent
//                         instruction stream, merge state in shared-memory columns
//   long_piece_kernel     257..4096 bytes: a warp per piece;  giant_piece_kernel  > 4096 bytes: a block per
//                         piece -- the round-synchronous exact merge in global scratch
//   probe_kernel          one warp per 1 KiB sub-tile: whole-piece table probe of every short piece
//                         (one 32 B sector each), one slot per piece, misses -> global queue
//   miss_{hist,base,scatter}, miss_kernel   the ~5 % misses, sorted by length, one piece per lane,
//                         warp-convergent exact min-rank merge
//   scan_{partial,top,final}, gather_kernel, big_copy_kernel   token counts -> offsets -> tokens and
//                         per-document offsets at their final place
//
// No tensor cores: nothing here is a contraction.  The work is byte/integer, bound by HBM reads
// of the text, L2 probes of the rank tables and instruction issue.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b200bpe.h"
#include "bpe_tables.h"
#include "text_access.cuh"
#include "pretok_fast.cuh"
#include "unicode_classes.inc"

using namespace b2bpe;

// --------------------------------------------------------------------------------------------
// error plumbing
// --------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
static int fail(int code, const std::string &msg) { g_last_error = msg; return code; }
#define CUDA_TRY(expr)                                                                             \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return fail(B200BPE_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));        \
    } while (0)

// --------------------------------------------------------------------------------------------
// device-side parameter blocks
// --------------------------------------------------------------------------------------------
static const uint32_t ERR_NOBYTE = 1u, ERR_DOCOFF = 2u;

struct UcTables { const uint16_t *stage1; const uint8_t *stage2; const uint8_t *ascii; };

struct Counters {            // device-resident, zeroed per call
    unsigned long long long_bytes;
    unsigned int n_long;
    unsigned int long_head;
    unsigned int n_cls[6];           // long pieces per length class (see LongQ::cls)
    unsigned int cls_head[6];        // work-queue heads of the per-class kernels
    unsigned int n_big;
    unsigned int n_miss;
    unsigned long long miss_bytes;
    unsigned int miss_hist[20];      // misses per piece length (2..16)
    unsigned int miss_fill[20];      // running fill of each length bucket (miss_sort_kernel)
    unsigned int ticket;
    unsigned int err;
    unsigned long long total_tokens;
};

static const uint32_t GIANT_MIN = 4096;      // pieces longer than this get a whole block (kernel 3b)
static const uint32_t LONG_SCRATCH_MIN = 256; // pieces longer than this merge in global scratch (warp / block per piece)

struct LongQ {               // queue of pieces longer than SHORT_MAX bytes
    unsigned long long *start;   // byte offset of the piece
    unsigned int *len;
    unsigned long long *off;     // offset of its region in the global merge scratch (pieces > LONG_SCRATCH_MIN only)
    unsigned int *ntok;
    // indices (into this queue) per length class: 0: 17..32, 1: 33..64, 2: 65..128, 3: 129..256 bytes
    // (thread-per-piece kernels), 4: 257..GIANT_MIN (warp per piece), 5: longer (block per piece)
    unsigned int *cls[6];
};
static const int N_CLS = 6, CLS_WARP = 4, CLS_GIANT = 5;

struct SmemCol32 {           // per-lane column of a [k][32] shared-memory array: bank == lane whatever k is
    uint32_t *base;
    __device__ __forceinline__ uint32_t &operator[](int j) const { return base[j * 32]; }
};

// --------------------------------------------------------------------------------------------
// kernel 0: documents -> doc-start bitmask, first document index per 32-byte span
// --------------------------------------------------------------------------------------------
__global__ void mark_docs_kernel(const unsigned long long *__restrict__ doc_off, unsigned long long n_docs,
                                 unsigned long long n_bytes, uint32_t *dbits, uint32_t *span_first_doc,
                                 Counters *ctr) {
    unsigned long long d = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (d > n_docs) return;                           // index n_docs is the end sentinel (== n_bytes)
    unsigned long long pos = doc_off[d];
    bool bad = pos > n_bytes || (d < n_docs && doc_off[d + 1] < pos) || (d == 0 && pos != 0) ||
               (d == n_docs && pos != n_bytes);
    if (bad) { atomicOr(&ctr->err, ERR_DOCOFF); return; }
    atomicOr(&dbits[pos >> 5], 1u << (pos & 31));
    atomicMin(&span_first_doc[pos >> 5], (uint32_t)d);
}

// --------------------------------------------------------------------------------------------
// kernel 1: pre-tokeniser.  One thread per 32-byte span = one word of the piece-start bitmask.
// The positions the bit-parallel rules cannot decide locally go through the general rule function,
// which is long and branchy: when a lane holds more than one of them, the warp pools its undecided
// positions and deals them out one per lane, so that the function runs once per ~32 positions instead
// of once per (busiest lane's) position with most lanes idle.
// --------------------------------------------------------------------------------------------
static const int PRETOK_WARPS = 8;

// o200k's rule function is long (case / mark chains): one out-of-line copy serves both call sites; the
// two shorter ones are cheaper inlined (measured both ways per pattern).
__device__ __noinline__ bool slow_boundary_o200k(const TextAccess &t, long long pos) { return boundary_before<PAT_O200K>(t, pos); }

template <int PAT>
__device__ __forceinline__ bool slow_boundary(const TextAccess &t, long long pos) {
    if (PAT == PAT_O200K) return slow_boundary_o200k(t, pos);
    return boundary_before<PAT>(t, pos);
}

template <int PAT>
__global__ void __launch_bounds__(PRETOK_WARPS * 32) pretok_kernel(const uint8_t *__restrict__ text, long long n_bytes,
                                                                  const uint32_t *__restrict__ dbits, UcTables uc,
                                                                  uint32_t *__restrict__ pbits, uint32_t *__restrict__ psum,
                                                                  long long n_words) {
    __shared__ uint16_t s_list[PRETOK_WARPS][1024];     // (owner lane << 5 | bit) of the pooled positions
    __shared__ uint32_t s_res[PRETOK_WARPS][32];
    const long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const TextAccess t{text, n_bytes, dbits, uc.stage1, uc.stage2, uc.ascii};
    uint64_t b = 0, slow = 0;
    if (w < n_words) b = span_fast<PAT>(t, w, slow);
    const uint32_t sm = (uint32_t)(slow >> 8);          // own positions only
    const int cnt = __popc(sm);
    const int mx = (int)__reduce_max_sync(0xFFFFFFFFu, (unsigned)cnt);
    if (mx == 1) {
        if (cnt && slow_boundary<PAT>(t, w * 32 + (__ffs(sm) - 1))) b |= (uint64_t)sm << 8;
    } else if (mx > 1) {
        int pre = cnt;                                  // inclusive scan over the lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xFFFFFFFFu, pre, o); if (lane >= o) pre += v; }
        const int total = __shfl_sync(0xFFFFFFFFu, pre, 31);
        pre -= cnt;
        s_res[warp][lane] = 0;
        for (uint32_t mm = sm; mm; mm &= mm - 1) s_list[warp][pre++] = (uint16_t)((lane << 5) | (__ffs(mm) - 1));
        __syncwarp();
        const long long w0 = w - lane;
        for (int i = lane; i < total; i += 32) {
            const unsigned e = s_list[warp][i];
            if (slow_boundary<PAT>(t, (w0 + (e >> 5)) * 32 + (e & 31))) atomicOr(&s_res[warp][e >> 5], 1u << (e & 31));
        }
        __syncwarp();
        b |= (uint64_t)s_res[warp][lane] << 8;
    }
    uint32_t word = 0;
    if (w < n_words) { word = span_word(t, w, b); pbits[w] = word; }
    // summary bitmap: bit = "this word of pbits has a piece start" (lets find_long skip long runs 32x faster)
    const uint32_t nz = __ballot_sync(0xFFFFFFFFu, word != 0);
    if (lane == 0 && (w >> 5) <= ((n_words - 1) >> 5)) psum[w >> 5] = nz;
}

// single-piece mode (encode_single_piece): P = {0, n_bytes}
__global__ void single_piece_bits_kernel(uint32_t *pbits, uint32_t *psum, long long n_bytes, long long n_words) {
    long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    uint32_t word = 0;
    if (w < n_words) {
        if (w == 0) word |= 1u;
        if ((n_bytes >> 5) == w) word |= 1u << (n_bytes & 31);
        pbits[w] = word;
    }
    const uint32_t nz = __ballot_sync(0xFFFFFFFFu, word != 0);
    if ((threadIdx.x & 31) == 0 && (w >> 5) <= ((n_words - 1) >> 5)) psum[w >> 5] = nz;
}

// --------------------------------------------------------------------------------------------
// kernel 2: find pieces longer than SHORT_MAX bytes
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) find_long_kernel(const uint32_t *__restrict__ pbits,
                                                        const uint32_t *__restrict__ psum, long long n_bytes,
                                                        long long n_words, LongQ q, uint32_t *lidx,
                                                        Counters *ctr) {
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
            // only pieces beyond the shared-memory path need a region of the global merge scratch
            unsigned long long off = 0;
            if (len > LONG_SCRATCH_MIN) off = atomicAdd(&ctr->long_bytes, (unsigned long long)len);
            q.start[i] = (unsigned long long)s; q.len[i] = (unsigned int)len; q.off[i] = off;
            lidx[s >> 4] = i;
            // per-class work list, one atomic per (warp iteration, class)
            const int c = len > GIANT_MIN ? CLS_GIANT : len > 256 ? CLS_WARP : len > 128 ? 3 : len > 64 ? 2 : len > 32 ? 1 : 0;
            const uint32_t same = __match_any_sync(peers, c);
            unsigned int k = 0;
            if ((threadIdx.x & 31) == __ffs(same) - 1) k = atomicAdd(&ctr->n_cls[c], (unsigned int)__popc(same));
            k = __shfl_sync(peers, k, __ffs(same) - 1) + __popc(same & ((1u << (threadIdx.x & 31)) - 1u));
            // (constant indices: a dynamically indexed kernel parameter would be copied to local memory by every thread)
            unsigned int *lst = c == 0 ? q.cls[0] : c == 1 ? q.cls[1] : c == 2 ? q.cls[2] : c == 3 ? q.cls[3] : c == 4 ? q.cls[4] : q.cls[5];
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
struct LongScratch {
    uint32_t *idA, *rkA, *idB, *rkB, *aux1, *aux2;
    uint8_t *flag;
};

__device__ __forceinline__ uint32_t warp_min_u32(uint32_t v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    return v;
}

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
                                                                    LongScratch S, uint32_t *ltok, Counters *ctr) {
    const int lane = threadIdx.x & 31;
    const unsigned int n_long = ctr->n_cls[CLS_WARP];
    const unsigned int *list = q.cls[CLS_WARP];
    for (;;) {
        unsigned int k = 0;
        if (lane == 0) k = atomicAdd(&ctr->cls_head[CLS_WARP], 1u);
        k = __shfl_sync(0xFFFFFFFFu, k, 0);
        if (k >= n_long) break;
        const unsigned int i = list[k];
        const unsigned long long off = q.off[i], st0 = q.start[i];
        LongScratch P = S;
        P.idA += off; P.rkA += off; P.idB += off; P.rkB += off; P.aux1 += off; P.aux2 += off; P.flag += off;
        const uint32_t nt = long_piece_warp(T, text + st0, q.len[i], P, ltok + st0, &ctr->err);
        if (lane == 0) q.ntok[i] = nt;
        __syncwarp();
    }
}

// --------------------------------------------------------------------------------------------
// kernel 3b: giant pieces (> GIANT_MIN bytes: "x"*1_000_000, long whitespace / separator runs).
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
    const unsigned int n_giant = ctr->n_cls[CLS_GIANT];
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_i = atomicAdd(&ctr->cls_head[CLS_GIANT], 1u);
        __syncthreads();
        if (s_i >= n_giant) break;
        const unsigned int i = q.cls[CLS_GIANT][s_i];
        const unsigned long long off = q.off[i];
        LongScratch P = S;
        P.idA += off; P.rkA += off; P.idB += off; P.rkB += off; P.aux1 += off; P.aux2 += off; P.flag += off;
        const uint32_t nt = long_piece_block(T, text + q.start[i], q.len[i], P, ltok + q.start[i], &ctr->err);
        if (threadIdx.x == 0) q.ntok[i] = nt;
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
            if (r != RANK_MAX) { out[0] = r; q.ntok[qi] = 1; n = 0; }
        }
        n_max = (int)__reduce_max_sync(0xFFFFFFFFu, (unsigned)n);
        if (n_max) {
            uint32_t c = 0; bool bad = false;
            if (CAP == 32) {
                const uint32_t mask = merge_short_conv(T, [&](int j) { return rk[j]; }, n, n_max, 0xFFFFFFFFu, id, rk);
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
            if (n) q.ntok[qi] = c;
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
// kernels 4-6: encode = probe -> merge the misses -> (scan) -> gather.
//
// probe_kernel   one WARP per 1 KiB sub-tile: stage text + piece starts in shared memory, compact
//                the piece starts into a list (all 32 lanes busy whatever the distribution), probe
//                every piece of <= 16 bytes in the piece table (src/lib.rs:367-368), two probes in
//                flight per lane.  One 32-bit slot per PIECE goes to `ptok` (token id, or a tagged
//                reference to the miss queue / the long-piece queue); misses are appended to a
//                global queue.  Small footprint (4.4 KB smem/warp) => many warps hide the L2 latency.
// miss_kernel    the ~5 % of pieces that are not tokens, DENSE: one piece per lane, 32 per warp,
//                all lanes walking one convergent instruction stream (merge_short_conv), the
//                literal min-rank loop of _byte_pair_merge (src/lib.rs:140-196).
// gather_kernel  one warp per sub-tile: per-piece token counts -> warp scan -> tokens and
//                per-document offsets written at their final position.
// Kernel boundaries do the ordering; there is no look-back chain and no ticket counter.
// --------------------------------------------------------------------------------------------
static const uint32_t PT_MISS = 0x40000000u, PT_LONG = 0x80000000u, PT_KIND = 0xC0000000u, PT_PAYLOAD = 0x3FFFFFFFu;
static const uint32_t PT_EMPTY = 0xFFFFFFFFu;             // zero-token slot (only on error paths)

struct MissQ {                // queue of pieces (2..16 bytes) that are not tokens themselves
    uint32_t *pos;            // byte offset of the piece
    uint32_t *roff;           // offset of its result region in mres (sum of lengths => tokens always fit)
    uint8_t *len;
    uint8_t *cnt;             // tokens produced (written by miss_kernel)
    uint32_t *order;          // queue indices sorted by piece length (so that a warp merges pieces of one length)
};

struct TileParams {
    const uint8_t *text; long long n_bytes; long long n_words; long long n_sub;
    const uint32_t *pbits; const uint32_t *dbits; const uint32_t *span_first_doc;
    const unsigned long long *doc_off; unsigned long long n_docs;
    LongQ q; const uint32_t *lidx; const uint32_t *ltok;
    uint32_t *ptok;               // [n_sub][SUB_BYTES] one slot per piece, in piece order
    MissQ mq; uint32_t *mres;     // miss queue and its token results
    uint32_t *sub_count;          // [n_sub] tokens emitted by the sub-tile
    unsigned long long *sub_base; // [n_sub+1] exclusive prefix of sub_count
    uint32_t *out; unsigned long long *tok_off;
    unsigned long long *big_dst, *big_src; uint32_t *big_n;   // token copies too large for one warp (kernel 7)
    Counters *ctr;
};

static const int ENC_WARPS = 4;                          // warps per block
static const int SUB_BYTES = 1024;                       // bytes per warp sub-tile (also: max pieces per sub-tile)

struct ProbeSmem {
    __align__(16) uint8_t text[SUB_BYTES + 32];
    uint32_t p[34];
    uint32_t nmiss;
    uint16_t plist[SUB_BYTES + 2];     // piece start offsets of the sub-tile, in order, + end sentinel
    uint16_t miss[SUB_BYTES / 2];      // piece indices (into plist) of the misses
};

__global__ void __launch_bounds__(ENC_WARPS * 32) probe_kernel(TileParams p, DevTables T) {
    __shared__ ProbeSmem smem[ENC_WARPS];
    ProbeSmem &S = smem[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const long long safe_end = p.n_bytes & ~15ll;
    const long long sub = (long long)blockIdx.x * ENC_WARPS + (threadIdx.x >> 5);
    if (sub >= p.n_sub) return;
    const long long sub_byte = sub * SUB_BYTES;
    const long long gw = sub * 32 + lane;              // this lane's bitmask word
    uint32_t *const slot = p.ptok + sub * SUB_BYTES;

    // ---- stage text (1 KiB + 32 B tail) and piece-start words ----------------------------
    for (int v = lane; v < (SUB_BYTES + 32) / 16; v += 32) {
        long long gpos = sub_byte + (long long)v * 16;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (gpos + 16 <= safe_end) val = __ldg(reinterpret_cast<const uint4 *>(p.text + gpos));
        else if (gpos < p.n_bytes) {
            uint32_t tmp[4] = {0, 0, 0, 0};
            for (int k = 0; k < 16; k++)
                if (gpos + k < p.n_bytes) tmp[k >> 2] |= (uint32_t)p.text[gpos + k] << (8 * (k & 3));
            val = make_uint4(tmp[0], tmp[1], tmp[2], tmp[3]);
        }
        *reinterpret_cast<uint4 *>(S.text + v * 16) = val;
    }
    {
        S.p[lane] = (gw < p.n_words) ? __ldg(p.pbits + gw) : 0u;
        if (lane < 2) { long long w2 = sub * 32 + 32 + lane; S.p[32 + lane] = (w2 < p.n_words) ? __ldg(p.pbits + w2) : 0u; }
        if (lane == 0) S.nmiss = 0;
    }
    __syncwarp();

    // ---- piece list ------------------------------------------------------------------------
    uint32_t pv = S.p[lane];
    {
        const long long span0 = sub_byte + lane * 32;
        if (span0 + 32 > p.n_bytes) {                      // drop the end sentinel / bits beyond the text
            const long long keep = p.n_bytes - span0;
            pv = keep <= 0 ? 0u : (pv & ((1u << keep) - 1u));
        }
    }
    uint32_t np;
    {
        const uint32_t c = __popc(pv);
        uint32_t inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= o) inc += y;
        }
        np = __shfl_sync(0xFFFFFFFFu, inc, 31);
        uint16_t *dst = S.plist + (inc - c);
        for (uint32_t m = pv; m;) { const int j = __ffs(m) - 1; m &= m - 1; *dst++ = (uint16_t)(lane * 32 + j); }
        if (lane == 0) {                                   // end sentinel: next piece start (or "far away")
            const uint32_t nx = S.p[32];
            const long long tail = p.n_bytes - sub_byte;           // text ends inside this sub-tile?
            S.plist[np] = (uint16_t)(tail <= SUB_BYTES ? tail : (nx ? SUB_BYTES + __ffs(nx) - 1 : SUB_BYTES + 32));
        }
    }
    __syncwarp();

    // ---- whole-piece probe, two pieces per lane per iteration --------------------------------
    uint32_t cnt = 0;                                      // tokens known so far (hits, single bytes, long pieces)
    auto prep = [&](uint32_t i, int &off, int &len, uint64_t &k0, uint64_t &k1) -> int {
        if (i >= np) return 0;
        off = S.plist[i];
        len = (int)S.plist[i + 1] - off;
        if (len > SHORT_MAX) {                             // long path: precomputed by long_piece / giant_piece
            const uint32_t qi = p.lidx[(sub_byte + off) >> 4];
            cnt += p.q.ntok[qi];
            slot[i] = PT_LONG | qi;
            return 0;
        }
        if (len == 1) {
            const uint32_t id = __ldg(T.byte_id + S.text[off]);
            if (id >= PSEUDO_BASE) { atomicOr(&p.ctr->err, ERR_NOBYTE); slot[i] = PT_EMPTY; }
            else { slot[i] = id; cnt++; }
            return 0;
        }
        const uint32_t *wp = reinterpret_cast<const uint32_t *>(S.text) + (off >> 2);
        const int sh = (off & 3) * 8;
        const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3], w4 = wp[4];
        const uint32_t a0 = __funnelshift_r(w0, w1, sh), a1 = __funnelshift_r(w1, w2, sh);
        const uint32_t a2 = __funnelshift_r(w2, w3, sh), a3 = __funnelshift_r(w3, w4, sh);
        const int nb0 = len >= 8 ? 64 : len * 8, nb1 = len <= 8 ? 0 : (len - 8) * 8;
        const uint64_t mk0 = nb0 >= 64 ? ~0ull : ((1ull << nb0) - 1ull);
        const uint64_t mk1 = nb1 >= 64 ? ~0ull : ((1ull << nb1) - 1ull);
        k0 = (((uint64_t)a1 << 32) | a0) & mk0; k1 = (((uint64_t)a3 << 32) | a2) & mk1;
        return 1;
    };
    auto finish = [&](uint32_t i, int len, uint64_t k0, uint64_t k1, uint32_t s, U4 m, U4 k) {
        uint32_t r = RANK_MAX;
        for (;;) {                                         // continue the linear probe from the prefetched slot
            if (m.x == 0) break;
            if (m.x == (uint32_t)len && k.x == (uint32_t)k0 && k.y == (uint32_t)(k0 >> 32) && k.z == (uint32_t)k1 &&
                k.w == (uint32_t)(k1 >> 32)) { r = m.y; break; }
            s = (s + 1) & T.piece_mask;
            m = B2_LDG_U4(T.piece_tab + 2 * s + 1); k = B2_LDG_U4(T.piece_tab + 2 * s);
        }
        if (r != RANK_MAX) { slot[i] = r; cnt++; }
        else S.miss[atomicAdd(&S.nmiss, 1u)] = (uint16_t)i;
    };
    for (uint32_t i = lane; i < np; i += 64) {
        int offA = 0, lenA = 0, offB = 0, lenB = 0;
        uint64_t a0 = 0, a1 = 0, b0 = 0, b1 = 0;
        const int needA = prep(i, offA, lenA, a0, a1);
        const int needB = prep(i + 32, offB, lenB, b0, b1);
        uint32_t sA = 0, sB = 0;
        U4 mA = {0, 0, 0, 0}, kA = {0, 0, 0, 0}, mB = {0, 0, 0, 0}, kB = {0, 0, 0, 0};
        if (needA) { sA = (uint32_t)piece_hash(a0, a1, (uint32_t)lenA) & T.piece_mask; mA = B2_LDG_U4(T.piece_tab + 2 * sA + 1); kA = B2_LDG_U4(T.piece_tab + 2 * sA); }
        if (needB) { sB = (uint32_t)piece_hash(b0, b1, (uint32_t)lenB) & T.piece_mask; mB = B2_LDG_U4(T.piece_tab + 2 * sB + 1); kB = B2_LDG_U4(T.piece_tab + 2 * sB); }
        if (needA) finish(i, lenA, a0, a1, sA, mA, kA);
        if (needB) finish(i + 32, lenB, b0, b1, sB, mB, kB);
    }
    __syncwarp();

    // ---- misses -> global queue (one atomic per sub-tile), result space = sum of their lengths ----
    const uint32_t nmiss = S.nmiss;
    if (nmiss) {
        uint32_t qbase = 0, rbase = 0, run = 0;
        for (uint32_t k0 = 0; k0 < nmiss; k0 += 32) {      // total length first
            const uint32_t k = k0 + lane;
            uint32_t len = 0;
            if (k < nmiss) { const uint32_t i = S.miss[k]; len = (uint32_t)S.plist[i + 1] - S.plist[i]; }
#pragma unroll
            for (int o = 16; o; o >>= 1) len += __shfl_xor_sync(0xFFFFFFFFu, len, o);
            run += len;
        }
        if (lane == 0) {
            qbase = atomicAdd(&p.ctr->n_miss, nmiss);
            rbase = (uint32_t)atomicAdd(&p.ctr->miss_bytes, (unsigned long long)run);
        }
        qbase = __shfl_sync(0xFFFFFFFFu, qbase, 0); rbase = __shfl_sync(0xFFFFFFFFu, rbase, 0);
        run = 0;
        for (uint32_t k0 = 0; k0 < nmiss; k0 += 32) {
            const uint32_t k = k0 + lane;
            uint32_t i = 0, off = 0, len = 0;
            if (k < nmiss) { i = S.miss[k]; off = S.plist[i]; len = (uint32_t)S.plist[i + 1] - off; }
            uint32_t inc = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += y; }
            if (k < nmiss) {
                const uint32_t qi = qbase + k;
                p.mq.pos[qi] = (uint32_t)(sub_byte + off); p.mq.len[qi] = (uint8_t)len; p.mq.roff[qi] = rbase + run + inc - len;
                slot[i] = PT_MISS | qi;
            }
            run += __shfl_sync(0xFFFFFFFFu, inc, 31);
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
    if (lane == 0) p.sub_count[sub] = cnt;                 // miss_kernel adds the tokens of the misses
}

// --------------------------------------------------------------------------------------------
// kernel 5: the misses, one piece per lane
// --------------------------------------------------------------------------------------------
static const int MISS_WARPS = 4;

struct MissSmem {
    uint32_t id[SHORT_MAX * 32];       // [part][lane]
    uint32_t rk[SHORT_MAX * 32];
    uint32_t bytes[4 * 32];            // [word][lane]: the piece bytes, little-endian
};

// counting sort of the miss queue by piece length (so that a warp merges pieces of one length):
// per-block histograms -> bucket bases -> scatter.  Only block-local shared-memory atomics and
// 17 values per block in global memory; no hot global counters.
static const int SORT_BLOCKS = 148 * 2;

__global__ void __launch_bounds__(256) miss_hist_kernel(TileParams p, unsigned int *block_hist /* [SORT_BLOCKS][17] */) {
    __shared__ unsigned int s_h[17];
    if (threadIdx.x < 17) s_h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t n_miss = p.ctr->n_miss;
    const uint32_t chunk = (n_miss + gridDim.x - 1) / gridDim.x;
    const uint32_t lo = blockIdx.x * chunk, hi = min(n_miss, lo + chunk);
    for (uint32_t qi = lo + threadIdx.x; qi < hi; qi += 256) atomicAdd(&s_h[p.mq.len[qi]], 1u);
    __syncthreads();
    if (threadIdx.x < 17) block_hist[blockIdx.x * 17 + threadIdx.x] = s_h[threadIdx.x];
}

// bucket bases: warp l turns column l of block_hist into exclusive offsets (bucket-major, then block order)
__global__ void __launch_bounds__(17 * 32) miss_base_kernel(unsigned int *block_hist, int n_blocks) {
    __shared__ unsigned int s_tot[17];
    const int l = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned int run = 0;
    for (int b0 = 0; b0 < n_blocks; b0 += 32) {
        const int b = b0 + lane;
        const unsigned int c = b < n_blocks ? block_hist[b * 17 + l] : 0u;
        unsigned int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned int y = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += y; }
        if (b < n_blocks) block_hist[b * 17 + l] = run + inc - c;
        run += __shfl_sync(0xFFFFFFFFu, inc, 31);
    }
    if (lane == 0) s_tot[l] = run;
    __syncthreads();
    unsigned int base = 0;
    for (int k = 0; k < l; k++) base += s_tot[k];
    for (int b = lane; b < n_blocks; b += 32) block_hist[b * 17 + l] += base;
}

__global__ void __launch_bounds__(256) miss_scatter_kernel(TileParams p, const unsigned int *block_base) {
    __shared__ unsigned int s_b[17];
    if (threadIdx.x < 17) s_b[threadIdx.x] = block_base[blockIdx.x * 17 + threadIdx.x];
    __syncthreads();
    const uint32_t n_miss = p.ctr->n_miss;
    const uint32_t chunk = (n_miss + gridDim.x - 1) / gridDim.x;
    const uint32_t lo = blockIdx.x * chunk, hi = min(n_miss, lo + chunk);
    for (uint32_t qi = lo + threadIdx.x; qi < hi; qi += 256) p.mq.order[atomicAdd(&s_b[p.mq.len[qi]], 1u)] = qi;
}

__global__ void __launch_bounds__(MISS_WARPS * 32) miss_kernel(TileParams p, DevTables T) {
    __shared__ MissSmem smem[MISS_WARPS];
    MissSmem &S = smem[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const uint32_t n_miss = p.ctr->n_miss;
    const uint32_t stride = gridDim.x * MISS_WARPS * 32;
    for (uint32_t q0 = (blockIdx.x * MISS_WARPS + (threadIdx.x >> 5)) * 32; q0 < n_miss; q0 += stride) {
        const bool have = q0 + lane < n_miss;
        const uint32_t qi = have ? p.mq.order[q0 + lane] : 0u;
        uint32_t pos = 0; int len = 0;
        if (have) { pos = p.mq.pos[qi]; len = p.mq.len[qi]; }
        {   // 16 bytes at an arbitrary offset: five aligned words + funnel shifts (text is padded)
            const uint32_t *wp = reinterpret_cast<const uint32_t *>(p.text + (pos & ~3u));
            const int sh = (pos & 3) * 8;
            uint32_t w[5] = {0, 0, 0, 0, 0};
            if (have) {
#pragma unroll
                for (int k = 0; k < 5; k++) w[k] = ((long long)(pos & ~3u) + 4 * k < p.n_bytes) ? __ldg(wp + k) : 0u;
            }
#pragma unroll
            for (int k = 0; k < 4; k++) S.bytes[k * 32 + lane] = __funnelshift_r(w[k], w[k + 1], sh);
        }
        int n_max = len;
#pragma unroll
        for (int o = 16; o; o >>= 1) n_max = max(n_max, __shfl_xor_sync(0xFFFFFFFFu, n_max, o));
        SmemCol32 id{S.id + lane}, rk{S.rk + lane};
        const uint32_t *bw = S.bytes + lane;
        const uint32_t mask = merge_short_conv(
            T, [&](int j) { return (bw[(j >> 2) * 32] >> (8 * (j & 3))) & 0xFFu; }, len, n_max, 0xFFFFFFFFu, id, rk);
        if (have) {
            uint32_t *dst = p.mres + p.mq.roff[qi];
            uint32_t c = 0; bool bad = false;
            for (uint32_t mm = mask; mm;) {
                const int j = __ffs(mm) - 1; mm &= mm - 1;
                const uint32_t x = id[j];
                bad |= x >= PSEUDO_BASE;
                dst[c++] = x;
            }
            if (bad) atomicOr(&p.ctr->err, ERR_NOBYTE);
            p.mq.cnt[qi] = (uint8_t)c;
            atomicAdd(&p.sub_count[pos >> 10], c);
        }
        __syncwarp();
    }
}

static const int SCAN_ITEMS = 4096;                 // counts per block of the two-level scan

__global__ void __launch_bounds__(256) scan_partial_kernel(const uint32_t *__restrict__ cnt, long long n,
                                                          unsigned long long *__restrict__ part) {
    __shared__ unsigned long long s_w[8];
    const long long lo = (long long)blockIdx.x * SCAN_ITEMS;
    unsigned long long sum = 0;
    for (int k = threadIdx.x; k < SCAN_ITEMS; k += 256) { long long i = lo + k; if (i < n) sum += cnt[i]; }
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) { unsigned long long t = 0; for (int i = 0; i < 8; i++) t += s_w[i]; part[blockIdx.x] = t; }
}

// single block: exclusive scan of the per-block partial sums (n_blocks <= a few thousand)
__global__ void __launch_bounds__(1024) scan_top_kernel(unsigned long long *part, long long n_blocks, Counters *ctr) {
    __shared__ unsigned long long s_part[1024];
    const int tid = threadIdx.x;
    const long long per = (n_blocks + 1023) / 1024;
    const long long lo = tid * per, hi = (lo + per < n_blocks) ? lo + per : n_blocks;
    unsigned long long sum = 0;
    for (long long i = lo; i < hi; i++) sum += part[i];
    s_part[tid] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        unsigned long long v = (tid >= o) ? s_part[tid - o] : 0ull;
        __syncthreads();
        s_part[tid] += v;
        __syncthreads();
    }
    unsigned long long run = s_part[tid] - sum;
    for (long long i = lo; i < hi; i++) { unsigned long long c = part[i]; part[i] = run; run += c; }
    if (tid == 1023) ctr->total_tokens = s_part[1023];
}

__global__ void __launch_bounds__(256) scan_final_kernel(const uint32_t *__restrict__ cnt, long long n,
                                                        const unsigned long long *__restrict__ part,
                                                        unsigned long long *__restrict__ base, const Counters *ctr) {
    __shared__ unsigned long long s_w[8];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long long lo = (long long)blockIdx.x * SCAN_ITEMS + (long long)tid * 16;   // 16 consecutive counts per thread
    uint32_t c[16]; unsigned long long sum = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) { long long i = lo + k; c[k] = (i < n) ? cnt[i] : 0u; sum += c[k]; }
    unsigned long long inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += y; }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    unsigned long long run = part[blockIdx.x] + inc - sum;
    for (int i = 0; i < wid; i++) run += s_w[i];
#pragma unroll
    for (int k = 0; k < 16; k++) { long long i = lo + k; if (i < n) base[i] = run; run += c[k]; }
    if (blockIdx.x == 0 && tid == 0) base[n] = ctr->total_tokens;
}

// chunked host path: rebase a slice of the caller's document offsets / globalise token offsets
__global__ void add_offset_kernel(unsigned long long *a, unsigned long long n, long long delta) {
    unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i < n) a[i] = (unsigned long long)((long long)a[i] + delta);
}

// --------------------------------------------------------------------------------------------
// kernel 6: gather.  One warp per sub-tile; each lane walks the pieces that start in its 32-byte
// span (they are consecutive slots of ptok): count, warp scan, then write tokens and the
// per-document token offsets at their final position.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_kernel(TileParams p) {
    const int lane = threadIdx.x & 31;
    const long long sub = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (sub >= p.n_sub) return;
    const unsigned long long base = p.sub_base[sub];
    const uint32_t *slot = p.ptok + sub * SUB_BYTES;
    const long long sub_byte = sub * SUB_BYTES;
    const long long gw = sub * 32 + lane;
    const bool in = gw < p.n_words;
    const uint32_t dm = in ? __ldg(p.dbits + gw) : 0u;
    uint32_t pv = in ? __ldg(p.pbits + gw) : 0u;
    {
        const long long span0 = sub_byte + lane * 32;
        if (span0 + 32 > p.n_bytes) {
            const long long keep = p.n_bytes - span0;
            pv = keep <= 0 ? 0u : (pv & ((1u << keep) - 1u));
        }
    }
    const uint32_t c = __popc(pv);
    uint32_t pinc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xFFFFFFFFu, pinc, o); if (lane >= o) pinc += y; }
    const uint32_t pi0 = pinc - c;                           // index of this lane's first piece
    auto count_of = [&](uint32_t v) -> uint32_t {
        if (v == PT_EMPTY) return 0u;
        const uint32_t kind = v & PT_KIND;
        if (kind == 0) return 1u;
        if (kind == PT_MISS) return p.mq.cnt[v & PT_PAYLOAD];
        return p.q.ntok[v & PT_PAYLOAD];
    };
    uint32_t t = 0;
    for (uint32_t k = 0; k < c; k++) t += count_of(slot[pi0 + k]);
    uint32_t tinc = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xFFFFFFFFu, tinc, o); if (lane >= o) tinc += y; }
    unsigned long long k = base + (tinc - t);
    unsigned long long big_dst = 0, big_src = 0; uint32_t big_n = 0;
    unsigned long long d = dm ? (unsigned long long)__ldg(p.span_first_doc + gw) : 0ull;
    uint32_t pi = pi0;
    uint32_t walk = pv | dm;
    while (walk) {
        const int j = __ffs(walk) - 1; walk &= walk - 1;
        const long long pos = sub_byte + lane * 32 + j;
        if ((dm >> j) & 1u) {
            while (d <= p.n_docs && p.doc_off[d] == (unsigned long long)pos) { p.tok_off[d] = k; d++; }
        }
        if ((pv >> j) & 1u) {
            const uint32_t v = slot[pi++];
            if (v == PT_EMPTY) continue;
            const uint32_t kind = v & PT_KIND;
            if (kind == 0) { p.out[k++] = v; }
            else if (kind == PT_MISS) {
                const uint32_t qi = v & PT_PAYLOAD;
                const uint32_t n = p.mq.cnt[qi];
                const uint32_t *src = p.mres + p.mq.roff[qi];
                for (uint32_t x = 0; x < n; x++) p.out[k + x] = src[x];
                k += n;
            } else {
                const uint32_t qi = v & PT_PAYLOAD;
                const uint32_t nt = p.q.ntok[qi];
                const unsigned long long lsrc = p.q.start[qi];
                if (nt <= 32) { for (uint32_t x = 0; x < nt; x++) p.out[k + x] = p.ltok[lsrc + x]; }
                else { big_dst = k; big_src = lsrc; big_n = nt; }   // > 32 tokens => > 32 bytes: at most one per span
                k += nt;
            }
        }
    }
    for (uint32_t pending = __ballot_sync(0xFFFFFFFFu, big_n != 0); pending; pending &= pending - 1) {
        const int src_lane = __ffs(pending) - 1;
        const unsigned long long dst = __shfl_sync(0xFFFFFFFFu, big_dst, src_lane);
        const unsigned long long bs = __shfl_sync(0xFFFFFFFFu, big_src, src_lane);
        const uint32_t nt = __shfl_sync(0xFFFFFFFFu, big_n, src_lane);
        if (nt > 4096) {                                   // giant piece: leave it to the whole grid (kernel 7)
            if (lane == 0) { const uint32_t e = atomicAdd(&p.ctr->n_big, 1u); p.big_dst[e] = dst; p.big_src[e] = bs; p.big_n[e] = nt; }
            continue;
        }
        for (uint32_t x = lane; x < nt; x += 32) p.out[dst + x] = p.ltok[bs + x];
    }
}

// kernel 7: token copies of giant pieces, spread over the whole grid
__global__ void __launch_bounds__(256) big_copy_kernel(TileParams p) {
    const unsigned int nb = p.ctr->n_big;
    for (unsigned int e = 0; e < nb; e++) {
        const unsigned long long dst = p.big_dst[e], src = p.big_src[e];
        const uint32_t n = p.big_n[e];
        for (unsigned long long x = blockIdx.x * 256ull + threadIdx.x; x < n; x += (unsigned long long)gridDim.x * 256ull)
            p.out[dst + x] = p.ltok[src + x];
    }
}

// --------------------------------------------------------------------------------------------
// decode ("next" row): CoreBPE::decode_bytes (src/lib.rs:345-358) for a whole batch -- token id ->
// byte string gather.  Length look-up, two-level scan (the same scan kernels), copy.
// --------------------------------------------------------------------------------------------
static const uint32_t ERR_BADTOKEN = 4u;

__global__ void __launch_bounds__(256) decode_len_kernel(const uint32_t *__restrict__ tokens, unsigned long long n,
                                                        const uint32_t *__restrict__ tok_boff, uint32_t n_ids,
                                                        uint32_t *__restrict__ len, Counters *ctr) {
    unsigned long long i = blockIdx.x * 256ull + threadIdx.x;
    if (i >= n) return;
    const uint32_t t = tokens[i];
    uint32_t l = 0;
    if (t < n_ids) l = __ldg(tok_boff + t + 1) - __ldg(tok_boff + t);
    if (l == 0) {                                           // unknown id (every real token has >= 1 byte)
        if (atomicOr(&ctr->err, ERR_BADTOKEN) == 0 || true) atomicMin(&ctr->ticket, (unsigned int)min(i, 0xFFFFFFFFull));
    }
    len[i] = l;
}

__global__ void __launch_bounds__(256) decode_copy_kernel(const uint32_t *__restrict__ tokens, unsigned long long n,
                                                         const uint32_t *__restrict__ tok_boff, uint32_t n_ids,
                                                         const uint8_t *__restrict__ blob,
                                                         const unsigned long long *__restrict__ base,
                                                         uint8_t *__restrict__ out) {
    unsigned long long i = blockIdx.x * 256ull + threadIdx.x;
    if (i >= n) return;
    const uint32_t t = tokens[i];
    if (t >= n_ids) return;
    const uint32_t b0 = __ldg(tok_boff + t), b1 = __ldg(tok_boff + t + 1);
    uint8_t *dst = out + base[i];
    for (uint32_t k = b0; k < b1; k++) *dst++ = __ldg(blob + k);
}

__global__ void decode_doc_off_kernel(const unsigned long long *__restrict__ tok_off, unsigned long long n_docs,
                                      const unsigned long long *__restrict__ base, unsigned long long *byte_off) {
    unsigned long long d = blockIdx.x * 256ull + threadIdx.x;
    if (d <= n_docs) byte_off[d] = base[tok_off[d]];
}

// --------------------------------------------------------------------------------------------
// host side: engine
// --------------------------------------------------------------------------------------------
namespace {

// pat_strs of tiktoken_ext/openai_public.py:12-14, :89, :104-114
const char *R50K_PAT = R"('(?:[sdmt]|ll|ve|re)| ?\p{L}++| ?\p{N}++| ?[^\s\p{L}\p{N}]++|\s++$|\s+(?!\S)|\s)";
const char *CL100K_PAT =
    R"('(?i:[sdmt]|ll|ve|re)|[^\r\n\p{L}\p{N}]?+\p{L}++|\p{N}{1,3}+| ?[^\s\p{L}\p{N}]++[\r\n]*+|\s++$|\s*[\r\n]|\s+(?!\S)|\s)";
const char *O200K_PAT =
    R"([^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]*[\p{Ll}\p{Lm}\p{Lo}\p{M}]+(?i:'s|'t|'re|'ve|'m|'ll|'d)?|)"
    R"([^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]+[\p{Ll}\p{Lm}\p{Lo}\p{M}]*(?i:'s|'t|'re|'ve|'m|'ll|'d)?|)"
    R"(\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n/]*|\s*[\r\n]+|\s+(?!\S)|\s+)";

template <class Tp>
struct DevBuf {
    Tp *p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 256;
        cudaError_t e = cudaMalloc((void **)&p, want * sizeof(Tp));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct PinnedBuf {
    void *p = nullptr; size_t cap = 0;
};

}  // namespace

// One pipeline slot: a stream, its events and a grow-only workspace.  The host path keeps
// three slots in flight (H2D of chunk c+1, kernels of chunk c, D2H of chunk c-1).
struct Slot {
    DevBuf<uint8_t> w_text; DevBuf<unsigned long long> w_docoff, w_tokoff, w_sub_base;
    DevBuf<uint32_t> w_ptok, w_mres, w_mq_pos, w_mq_roff, w_mq_order, w_sub_count; DevBuf<uint8_t> w_mq_len, w_mq_cnt;
    DevBuf<uint32_t> w_dbits, w_pbits, w_psum, w_sfd, w_lidx, w_out, w_ltok;
    DevBuf<unsigned long long> w_lq_start, w_lq_off; DevBuf<unsigned int> w_lq_len, w_lq_ntok, w_lq_cls, w_big_n, w_sort_hist; DevBuf<unsigned long long> w_big_dst, w_big_src, w_scan_part;
    DevBuf<uint32_t> w_idA, w_rkA, w_idB, w_rkB, w_aux1, w_aux2; DevBuf<uint8_t> w_flag;
    Counters *d_ctr = nullptr; Counters *h_ctr = nullptr;   // h_ctr pinned
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[9];
    float last_ms[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t last_launches = 0;
    bool ok = false;

    cudaError_t init() {
        cudaError_t e = cudaMalloc((void **)&d_ctr, sizeof(Counters));
        if (e == cudaSuccess) e = cudaHostAlloc((void **)&h_ctr, sizeof(Counters), cudaHostAllocDefault);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
        for (int i = 0; i < 9 && e == cudaSuccess; i++) e = cudaEventCreate(&ev[i]);
        ok = (e == cudaSuccess);
        return e;
    }
    void destroy() {
        w_text.release(); w_docoff.release(); w_tokoff.release(); w_sub_base.release();
        w_ptok.release(); w_mres.release(); w_mq_pos.release(); w_mq_roff.release(); w_mq_order.release(); w_sub_count.release();
        w_mq_len.release(); w_mq_cnt.release();
        w_dbits.release(); w_pbits.release(); w_psum.release(); w_sfd.release(); w_lidx.release(); w_out.release(); w_ltok.release();
        w_lq_start.release(); w_lq_off.release(); w_lq_len.release(); w_lq_ntok.release(); w_lq_cls.release(); w_big_n.release(); w_sort_hist.release(); w_big_dst.release(); w_big_src.release(); w_scan_part.release();
        w_idA.release(); w_rkA.release(); w_idB.release(); w_rkB.release(); w_aux1.release(); w_aux2.release();
        w_flag.release();
        if (d_ctr) cudaFree(d_ctr);
        if (h_ctr) cudaFreeHost(h_ctr);
        if (ok) { for (int i = 0; i < 9; i++) cudaEventDestroy(ev[i]); }
        if (stream) cudaStreamDestroy(stream);
    }
};

struct b200bpe_result {
    b200bpe *owner = nullptr;
    PinnedBuf tok, off;                 // pinned when produced by the engine
    std::vector<uint32_t> vtok;         // used when the result is assembled on the host (specials)
    std::vector<uint64_t> voff;
    bool on_host_vec = false;
    uint64_t n_tokens = 0, n_docs = 0;
};

struct b200bpe {
    int device = 0;
    int pattern = 0;
    HostTables H;
    std::vector<std::string> specials; std::vector<uint32_t> special_rank;
    std::unordered_map<uint32_t, std::string> special_decoder;
    // device tables
    uint32_t *d_byte_id = nullptr, *d_pair2 = nullptr;
    U4 *d_pair_tab = nullptr, *d_piece_tab = nullptr, *d_long_tab = nullptr;
    uint8_t *d_long_blob = nullptr;
    uint16_t *d_uc1 = nullptr; uint8_t *d_uc2 = nullptr, *d_ascii = nullptr;
    uint32_t *d_tok_boff = nullptr; uint8_t *d_tok_blob = nullptr; uint32_t n_ids = 0;   // decode: id -> bytes
    DevTables T; UcTables uc;
    uint64_t table_bytes[4] = {0, 0, 0, 0};
    static const int N_SLOTS = 3;
    Slot slots[3];
    float last_ms[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t last_launches = 0;
    size_t chunk_bytes = 64u << 20;
    std::mutex mu;
    std::vector<PinnedBuf> pinned_pool;
    // results keep the engine alive: b200bpe_destroy with results outstanding only marks the handle dead, the last
    // b200bpe_result_free tears it down (the reference's TiktokenBuffer owns its Vec, src/py.rs:186-189)
    int live_results = 0;
    bool dead = false;

    PinnedBuf take_pinned(size_t bytes) {
        for (size_t i = 0; i < pinned_pool.size(); i++)
            if (pinned_pool[i].cap >= bytes) { PinnedBuf b = pinned_pool[i]; pinned_pool.erase(pinned_pool.begin() + i); return b; }
        PinnedBuf b; size_t want = bytes + bytes / 8 + 4096;
        if (cudaHostAlloc(&b.p, want, cudaHostAllocDefault) != cudaSuccess) { b.p = nullptr; b.cap = 0; return b; }
        b.cap = want; return b;
    }
    void give_pinned(PinnedBuf b) {
        if (!b.p) return;
        if (pinned_pool.size() >= 4) { cudaFreeHost(b.p); return; }
        pinned_pool.push_back(b);
    }
};

template <class Tp>
static cudaError_t upload(Tp **dst, const void *src, size_t bytes) {
    cudaError_t e = cudaMalloc((void **)dst, bytes ? bytes : 16);
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
}

extern "C" const char *b200bpe_last_error(void) { return g_last_error.c_str(); }
extern "C" const char *b200bpe_version(void) { return "b200bpe 0.2 (sm_100a)"; }
extern "C" int b200bpe_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int b200bpe_create(const uint8_t *tok_bytes, const uint64_t *tok_off, const uint32_t *tok_rank,
                              uint32_t n_tok, const uint8_t *sp_bytes, const uint64_t *sp_off,
                              const uint32_t *sp_rank, uint32_t n_sp, const char *pat_str, int device,
                              b200bpe_t **out) {
    if (!out || !pat_str || (n_tok && (!tok_bytes || !tok_off || !tok_rank))) return fail(B200BPE_EINVAL, "null argument");
    int pattern;
    if (strcmp(pat_str, R50K_PAT) == 0) pattern = PAT_R50K;
    else if (strcmp(pat_str, CL100K_PAT) == 0) pattern = PAT_CL100K;
    else if (strcmp(pat_str, O200K_PAT) == 0) pattern = PAT_O200K;
    else return fail(B200BPE_EPATTERN,
                     "unsupported pat_str: the B200 pre-tokeniser implements exactly the r50k/p50k, cl100k and "
                     "o200k patterns of tiktoken_ext/openai_public.py (there is no CPU regex fallback)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(B200BPE_ECUDA, "no CUDA device: libb200bpe has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(B200BPE_EINVAL, "bad device index");
    b200bpe *h = new b200bpe();
    h->device = device; h->pattern = pattern;
    int rc = build_tables(tok_bytes, tok_off, tok_rank, n_tok, h->H);
    if (rc) { std::string m = h->H.error; delete h; return fail(rc == -3 ? B200BPE_EDUPRANK : B200BPE_EINVAL, m); }
    for (uint32_t i = 0; i < n_sp; i++) {
        std::string s((const char *)sp_bytes + sp_off[i], (size_t)(sp_off[i + 1] - sp_off[i]));
        h->specials.push_back(s); h->special_rank.push_back(sp_rank[i]);
        h->special_decoder[sp_rank[i]] = s;
    }
    cudaError_t e = cudaSetDevice(device);
    const HostTables &H = h->H;
    uint8_t ascii[128];
    for (int i = 0; i < 128; i++) ascii[i] = UC_STAGE2[(uint32_t)UC_STAGE1[0] * 256 + i];
    if (e == cudaSuccess) e = upload(&h->d_byte_id, H.byte_id.data(), 256 * 4);
    if (e == cudaSuccess) e = upload(&h->d_pair2, H.pair2.data(), 65536 * 4);
    if (e == cudaSuccess) e = upload(&h->d_pair_tab, H.pair_tab.data(), H.pair_tab.size() * sizeof(U4));
    if (e == cudaSuccess) e = upload(&h->d_piece_tab, H.piece_tab.data(), H.piece_tab.size() * sizeof(U4));
    if (e == cudaSuccess) e = upload(&h->d_long_tab, H.long_tab.data(), H.long_tab.size() * sizeof(U4));
    if (e == cudaSuccess) e = upload(&h->d_long_blob, H.long_blob.data(), H.long_blob.size());
    if (e == cudaSuccess) e = upload(&h->d_uc1, UC_STAGE1, sizeof(UC_STAGE1));
    if (e == cudaSuccess) e = upload(&h->d_uc2, UC_STAGE2, sizeof(UC_STAGE2));
    if (e == cudaSuccess) e = upload(&h->d_ascii, ascii, 128);
    {   // decode tables: byte offsets by token id (mergeable ranks and specials); ids above 2^24 are left out
        uint32_t max_id = 0; bool any = false;
        for (auto &kv : H.decoder) if (kv.first < (1u << 24)) { if (kv.first > max_id) max_id = kv.first; any = true; }
        for (auto &kv : h->special_decoder) if (kv.first < (1u << 24)) { if (kv.first > max_id) max_id = kv.first; any = true; }
        h->n_ids = any ? max_id + 1 : 0;
        std::vector<uint32_t> boff((size_t)h->n_ids + 2, 0);
        std::vector<uint8_t> blob;
        for (uint32_t id = 0; id < h->n_ids; id++) {
            boff[id] = (uint32_t)blob.size();
            const std::string *sp = nullptr;
            auto it = H.decoder.find(id);
            if (it != H.decoder.end()) sp = &it->second;
            else { auto it2 = h->special_decoder.find(id); if (it2 != h->special_decoder.end()) sp = &it2->second; }
            if (sp) blob.insert(blob.end(), sp->begin(), sp->end());
        }
        boff[h->n_ids] = (uint32_t)blob.size(); boff[h->n_ids + 1] = (uint32_t)blob.size();
        if (blob.empty()) blob.push_back(0);
        if (e == cudaSuccess) e = upload(&h->d_tok_boff, boff.data(), boff.size() * 4);
        if (e == cudaSuccess) e = upload(&h->d_tok_blob, blob.data(), blob.size());
    }
    for (int i = 0; i < b200bpe::N_SLOTS && e == cudaSuccess; i++) e = h->slots[i].init();
    CUDA_TRY(cudaFuncSetAttribute(mid_thread_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MID_SMEM_BYTES));
    if (const char *cm = getenv("B200BPE_CHUNK_MB")) { long v = atol(cm); if (v >= 1 && v <= 2048) h->chunk_bytes = (size_t)v << 20; }
    if (e != cudaSuccess) { std::string m = cudaGetErrorString(e); delete h; return fail(B200BPE_ECUDA, "table upload: " + m); }
    h->T.byte_id = h->d_byte_id; h->T.pair2 = h->d_pair2;
    h->T.pair_tab = h->d_pair_tab; h->T.pair_mask = H.pair_mask;
    h->T.piece_tab = h->d_piece_tab; h->T.piece_mask = H.piece_mask;
    h->T.long_tab = h->d_long_tab; h->T.long_mask = H.long_mask;
    h->T.long_blob = h->d_long_blob; h->T.max_token_len = H.max_token_len; h->T.n_long_tokens = H.n_long_tokens;
    h->uc.stage1 = h->d_uc1; h->uc.stage2 = h->d_uc2; h->uc.ascii = h->d_ascii;
    h->table_bytes[0] = H.piece_tab.size() * sizeof(U4);
    h->table_bytes[1] = H.pair_tab.size() * sizeof(U4) + 65536 * 4 + 1024;
    h->table_bytes[2] = H.long_tab.size() * sizeof(U4) + H.long_blob.size();
    h->table_bytes[3] = sizeof(UC_STAGE1) + sizeof(UC_STAGE2) + 128;
    *out = h;
    return B200BPE_OK;
}

static void engine_teardown(b200bpe *h);

extern "C" void b200bpe_destroy(b200bpe_t *h) {
    if (!h) return;
    {
        std::lock_guard<std::mutex> lk(h->mu);
        if (h->live_results > 0) { h->dead = true; return; }
    }
    engine_teardown(h);
}

static void engine_teardown(b200bpe *h) {
    int prev = 0; cudaGetDevice(&prev);
    cudaSetDevice(h->device);
    cudaFree(h->d_byte_id); cudaFree(h->d_pair2); cudaFree(h->d_pair_tab); cudaFree(h->d_piece_tab);
    cudaFree(h->d_long_tab); cudaFree(h->d_long_blob); cudaFree(h->d_tok_boff); cudaFree(h->d_tok_blob); cudaFree(h->d_uc1); cudaFree(h->d_uc2); cudaFree(h->d_ascii);
    for (int i = 0; i < b200bpe::N_SLOTS; i++) h->slots[i].destroy();
    for (auto &b : h->pinned_pool) cudaFreeHost(b.p);
    delete h;
    cudaSetDevice(prev);
}

// The device pipeline.  All pointers are device pointers on h->device; d_text must be 16-byte
// aligned.  Caller holds h->mu.  `single_piece`: treat the whole buffer as one piece (no regex).
static int finish_pipeline(b200bpe *h, Slot &S, cudaStream_t st);

// Batches up to this size get their long-piece scratch sized for the worst case up front, so the
// whole pipeline is enqueued without a host round trip in the middle.
static const uint64_t PRESIZE_LIMIT = 256ull << 20;

static int run_pipeline(b200bpe *h, Slot &S, const uint8_t *d_text, uint64_t n_bytes, const unsigned long long *d_doc_off,
                        uint64_t n_docs, uint32_t *d_out, unsigned long long *d_tok_off, cudaStream_t st,
                        bool single_piece, bool defer_finish = false) {
    if (n_bytes >= (1ull << 32) - 4096) return fail(B200BPE_EINVAL, "batch too large for one call (>= 4 GiB)");
    if (n_docs >= 0xFFFFFFFEull) return fail(B200BPE_EINVAL, "too many documents in one call");
    const long long n_words = (long long)((n_bytes + 1 + 31) / 32);
    const long long n_tiles = (n_words + 31) / 32;                  // 1 KiB sub-tiles, one warp each
    CUDA_TRY(S.w_dbits.ensure((size_t)n_words + 4));
    CUDA_TRY(S.w_pbits.ensure((size_t)n_words + 4));
    CUDA_TRY(S.w_psum.ensure((size_t)(n_words >> 5) + 4));
    CUDA_TRY(S.w_sfd.ensure((size_t)n_words + 4));
    CUDA_TRY(S.w_sub_base.ensure((size_t)n_tiles + 2));
    CUDA_TRY(S.w_scan_part.ensure((size_t)(n_tiles / SCAN_ITEMS) + 4));
    CUDA_TRY(S.w_sub_count.ensure((size_t)n_tiles + 2));
    CUDA_TRY(S.w_ptok.ensure((size_t)n_tiles * SUB_BYTES + 64));
    {   // a miss has >= 2 bytes; its tokens never outnumber its bytes
        const size_t mcap = (size_t)(n_bytes / 2) + 64;
        CUDA_TRY(S.w_mq_pos.ensure(mcap)); CUDA_TRY(S.w_mq_roff.ensure(mcap)); CUDA_TRY(S.w_mq_order.ensure(mcap));
        CUDA_TRY(S.w_mq_len.ensure(mcap)); CUDA_TRY(S.w_mq_cnt.ensure(mcap));
        CUDA_TRY(S.w_mres.ensure((size_t)n_bytes + 64));
        CUDA_TRY(S.w_sort_hist.ensure((size_t)SORT_BLOCKS * 17 + 32));
    }
    CUDA_TRY(S.w_lidx.ensure((size_t)(n_bytes >> 4) + 4));
    const size_t qcap = (size_t)(n_bytes / (SHORT_MAX + 1)) + 4;
    CUDA_TRY(S.w_lq_start.ensure(qcap)); CUDA_TRY(S.w_lq_off.ensure(qcap));
    CUDA_TRY(S.w_lq_len.ensure(qcap)); CUDA_TRY(S.w_lq_ntok.ensure(qcap));
    size_t cls_cap[N_CLS], cls_total = 0;                          // per-class index lists, back to back
    {
        const size_t min_len[N_CLS] = {SHORT_MAX + 1, 33, 65, 129, 257, GIANT_MIN + 1};
        for (int c = 0; c < N_CLS; c++) {
            cls_cap[c] = (size_t)(n_bytes / min_len[c]) + 4;
            cls_total += cls_cap[c];
        }
    }
    CUDA_TRY(S.w_lq_cls.ensure(cls_total));
    CUDA_TRY(S.w_big_n.ensure((size_t)(n_bytes / (GIANT_MIN + 1)) + 4));
    CUDA_TRY(S.w_big_dst.ensure((size_t)(n_bytes / (GIANT_MIN + 1)) + 4)); CUDA_TRY(S.w_big_src.ensure((size_t)(n_bytes / (GIANT_MIN + 1)) + 4));
    LongQ q{S.w_lq_start.p, S.w_lq_len.p, S.w_lq_off.p, S.w_lq_ntok.p, {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}};
    { size_t o = 0; for (int c = 0; c < N_CLS; c++) { q.cls[c] = S.w_lq_cls.p + o; o += cls_cap[c]; } }
    uint32_t launches = 0;

    CUDA_TRY(cudaEventRecord(S.ev[0], st));
    CUDA_TRY(cudaMemsetAsync(S.d_ctr, 0, sizeof(Counters), st));
    CUDA_TRY(cudaMemsetAsync(S.w_dbits.p, 0, ((size_t)n_words + 4) * 4, st));
    CUDA_TRY(cudaMemsetAsync(S.w_sfd.p, 0xFF, ((size_t)n_words + 4) * 4, st));
    CUDA_TRY(cudaMemsetAsync(S.w_pbits.p + n_words, 0, 4 * 4, st));
    {
        unsigned long long nd1 = n_docs + 1;
        mark_docs_kernel<<<(unsigned)((nd1 + 255) / 256), 256, 0, st>>>(d_doc_off, n_docs, n_bytes, S.w_dbits.p,
                                                                       S.w_sfd.p, S.d_ctr);
        launches++;
    }
    CUDA_TRY(cudaEventRecord(S.ev[1], st));
    {
        unsigned grid = (unsigned)((n_words + 255) / 256);
        if (single_piece) single_piece_bits_kernel<<<grid, 256, 0, st>>>(S.w_pbits.p, S.w_psum.p, (long long)n_bytes, n_words);
        else if (h->pattern == PAT_R50K)
            pretok_kernel<PAT_R50K><<<grid, 256, 0, st>>>(d_text, (long long)n_bytes, S.w_dbits.p, h->uc, S.w_pbits.p, S.w_psum.p, n_words);
        else if (h->pattern == PAT_CL100K)
            pretok_kernel<PAT_CL100K><<<grid, 256, 0, st>>>(d_text, (long long)n_bytes, S.w_dbits.p, h->uc, S.w_pbits.p, S.w_psum.p, n_words);
        else
            pretok_kernel<PAT_O200K><<<grid, 256, 0, st>>>(d_text, (long long)n_bytes, S.w_dbits.p, h->uc, S.w_pbits.p, S.w_psum.p, n_words);
        launches++;
    }
    CUDA_TRY(cudaEventRecord(S.ev[2], st));
    {
        unsigned grid = (unsigned)((n_words + 255) / 256);
        find_long_kernel<<<grid, 256, 0, st>>>(S.w_pbits.p, S.w_psum.p, (long long)n_bytes, n_words, q, S.w_lidx.p, S.d_ctr);
        launches++;
    }
    const bool presized = n_bytes <= PRESIZE_LIMIT;
    size_t long_cap = (size_t)n_bytes + 4;
    unsigned long_blocks = 148 * 8;
    if (!presized) {
        CUDA_TRY(cudaMemcpyAsync(S.h_ctr, S.d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (S.h_ctr->err & ERR_DOCOFF) return fail(B200BPE_EINVAL, "document offsets must start at 0, be non-decreasing and end at n_bytes");
        long_cap = (size_t)S.h_ctr->long_bytes + 4;
        long_blocks = (S.h_ctr->n_long + LONG_WARPS - 1) / LONG_WARPS;
        if (long_blocks > 148 * 8) long_blocks = 148 * 8;
    }
    CUDA_TRY(S.w_ltok.ensure((size_t)n_bytes + 4));
    if (long_blocks) {
        // 72 KiB of columns per block: 3 blocks per SM
        mid_thread_kernel<<<148 * 3, MID_WARPS * 32, MID_SMEM_BYTES, st>>>(d_text, h->T, q, S.w_ltok.p, S.d_ctr);
        launches++;
    }
    if (long_blocks) {
        CUDA_TRY(S.w_idA.ensure(long_cap)); CUDA_TRY(S.w_rkA.ensure(long_cap));
        CUDA_TRY(S.w_idB.ensure(long_cap)); CUDA_TRY(S.w_rkB.ensure(long_cap));
        CUDA_TRY(S.w_aux1.ensure(long_cap)); CUDA_TRY(S.w_aux2.ensure(long_cap));
        CUDA_TRY(S.w_flag.ensure(long_cap));
        LongScratch LS{S.w_idA.p, S.w_rkA.p, S.w_idB.p, S.w_rkB.p, S.w_aux1.p, S.w_aux2.p, S.w_flag.p};
        long_piece_kernel<<<long_blocks, LONG_WARPS * 32, 0, st>>>(d_text, h->T, q, LS, S.w_ltok.p, S.d_ctr);
        giant_piece_kernel<<<148, GIANT_THREADS, 0, st>>>(d_text, h->T, q, LS, S.w_ltok.p, S.d_ctr);
        launches += 2;
    }
    CUDA_TRY(cudaEventRecord(S.ev[3], st));
    {
        TileParams p;
        p.text = d_text; p.n_bytes = (long long)n_bytes; p.n_words = n_words; p.n_sub = n_tiles;
        p.pbits = S.w_pbits.p; p.dbits = S.w_dbits.p; p.span_first_doc = S.w_sfd.p;
        p.doc_off = d_doc_off; p.n_docs = n_docs; p.q = q; p.lidx = S.w_lidx.p; p.ltok = S.w_ltok.p;
        p.ptok = S.w_ptok.p; p.mres = S.w_mres.p;
        p.mq = MissQ{S.w_mq_pos.p, S.w_mq_roff.p, S.w_mq_len.p, S.w_mq_cnt.p, S.w_mq_order.p};
        p.sub_count = S.w_sub_count.p; p.sub_base = S.w_sub_base.p;
        p.out = d_out; p.tok_off = d_tok_off; p.ctr = S.d_ctr;
        p.big_dst = S.w_big_dst.p; p.big_src = S.w_big_src.p; p.big_n = S.w_big_n.p;
        probe_kernel<<<(unsigned)((n_tiles + ENC_WARPS - 1) / ENC_WARPS), ENC_WARPS * 32, 0, st>>>(p, h->T);
        CUDA_TRY(cudaEventRecord(S.ev[8], st));
        miss_hist_kernel<<<SORT_BLOCKS, 256, 0, st>>>(p, S.w_sort_hist.p);
        miss_base_kernel<<<1, 17 * 32, 0, st>>>(S.w_sort_hist.p, SORT_BLOCKS);
        miss_scatter_kernel<<<SORT_BLOCKS, 256, 0, st>>>(p, S.w_sort_hist.p);
        miss_kernel<<<148 * 16, MISS_WARPS * 32, 0, st>>>(p, h->T);
        CUDA_TRY(cudaEventRecord(S.ev[7], st));
        {
            const long long nb = (n_tiles + SCAN_ITEMS - 1) / SCAN_ITEMS;
            scan_partial_kernel<<<(unsigned)nb, 256, 0, st>>>(S.w_sub_count.p, n_tiles, S.w_scan_part.p);
            scan_top_kernel<<<1, 1024, 0, st>>>(S.w_scan_part.p, nb, S.d_ctr);
            scan_final_kernel<<<(unsigned)nb, 256, 0, st>>>(S.w_sub_count.p, n_tiles, S.w_scan_part.p, S.w_sub_base.p, S.d_ctr);
        }
        gather_kernel<<<(unsigned)((n_tiles + 7) / 8), 256, 0, st>>>(p);
        big_copy_kernel<<<148 * 2, 256, 0, st>>>(p);
        launches += 10;
    }
    CUDA_TRY(cudaEventRecord(S.ev[4], st));
    CUDA_TRY(cudaMemcpyAsync(S.h_ctr, S.d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, st));
    S.last_launches = launches;
    if (defer_finish) return B200BPE_OK;
    return finish_pipeline(h, S, st);
}

// wait for an enqueued pipeline, collect stage timings and device-side error flags
static int finish_pipeline(b200bpe *h, Slot &S, cudaStream_t st) {
    (void)h;
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    cudaEventElapsedTime(&S.last_ms[0], S.ev[0], S.ev[1]);
    cudaEventElapsedTime(&S.last_ms[1], S.ev[1], S.ev[2]);
    cudaEventElapsedTime(&S.last_ms[2], S.ev[2], S.ev[3]);
    cudaEventElapsedTime(&S.last_ms[3], S.ev[3], S.ev[7]);
    cudaEventElapsedTime(&S.last_ms[7], S.ev[7], S.ev[4]);
    cudaEventElapsedTime(&S.last_ms[8], S.ev[3], S.ev[8]);
    cudaEventElapsedTime(&S.last_ms[4], S.ev[0], S.ev[4]);
    if (S.h_ctr->err & ERR_DOCOFF)
        return fail(B200BPE_EINVAL, "document offsets must start at 0, be non-decreasing and end at n_bytes");
    if (S.h_ctr->err & ERR_NOBYTE)
        return fail(B200BPE_ENOBYTE, "a piece needs a single-byte token that mergeable_ranks does not contain");
    return B200BPE_OK;
}

extern "C" int b200bpe_encode_device(b200bpe_t *h, const uint8_t *d_text, uint64_t n_bytes, const uint64_t *d_doc_off,
                                     uint64_t n_docs, uint32_t *d_tokens, uint64_t *d_tok_off, uint64_t *n_tokens,
                                     void *stream) {
    if (!h || !d_doc_off || !d_tokens || !d_tok_off || (n_bytes && !d_text)) return fail(B200BPE_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    CUDA_TRY(cudaSetDevice(h->device));
    Slot &S = h->slots[0];
    cudaStream_t st = stream ? (cudaStream_t)stream : S.stream;
    const uint8_t *txt = d_text;
    if (((uintptr_t)d_text & 15u) != 0) {                       // vector loads need 16-byte alignment
        CUDA_TRY(S.w_text.ensure((size_t)n_bytes + 64));
        CUDA_TRY(cudaMemcpyAsync(S.w_text.p, d_text, n_bytes, cudaMemcpyDeviceToDevice, st));
        txt = S.w_text.p;
    }
    int rc = run_pipeline(h, S, txt, n_bytes, (const unsigned long long *)d_doc_off, n_docs, d_tokens,
                          (unsigned long long *)d_tok_off, st, false);
    memcpy(h->last_ms, S.last_ms, sizeof(h->last_ms));
    h->last_ms[5] = h->last_ms[6] = 0.f;
    h->last_launches = S.last_launches;
    if (rc) return rc;
    if (n_tokens) *n_tokens = S.h_ctr->total_tokens;
    return B200BPE_OK;
}

// Host buffers in, pinned host buffers out.  Large batches are cut at document boundaries into
// chunks of ~chunk_bytes and pipelined over three slots: while chunk c is in the kernels, chunk
// c+1 is already crossing PCIe host->device and the tokens of chunk c-1 are crossing device->host.
static int encode_host(b200bpe *h, const uint8_t *text, const uint64_t *doc_off, uint64_t n_docs, bool single_piece,
                       b200bpe_result **out) {
    const uint64_t n_bytes = doc_off[n_docs];
    CUDA_TRY(cudaSetDevice(h->device));
    if (doc_off[0] != 0) return fail(B200BPE_EINVAL, "document offsets must start at 0");
    // ---- chunk plan: document ranges [lo, hi) ------------------------------------------------
    std::vector<uint64_t> cut; cut.push_back(0);
    if (!single_piece && n_bytes > h->chunk_bytes + h->chunk_bytes / 2) {
        uint64_t lo = 0;
        while (lo < n_docs) {
            const uint64_t want = doc_off[lo] + h->chunk_bytes;
            uint64_t hi = (uint64_t)(std::upper_bound(doc_off + lo, doc_off + n_docs + 1, want) - doc_off);   // first off > want
            if (hi > 0) hi--;                                    // last doc boundary <= want
            if (hi <= lo) hi = lo + 1;                           // a single document larger than a chunk
            if (hi > n_docs) hi = n_docs;
            cut.push_back(hi); lo = hi;
        }
    } else cut.push_back(n_docs);
    const size_t n_chunks = cut.size() - 1;
    for (size_t c = 0; c < n_chunks; c++) {
        const uint64_t cb = doc_off[cut[c + 1]] - doc_off[cut[c]];
        if (doc_off[cut[c + 1]] < doc_off[cut[c]] || cb >= (1ull << 32) - 4096)
            return fail(B200BPE_EINVAL, cb >= (1ull << 32) - 4096 ? "a single document of >= 4 GiB is not supported"
                                                                  : "document offsets must be non-decreasing");
    }
    b200bpe_result *r = new b200bpe_result();
    r->owner = h; r->n_docs = n_docs;
    r->off = h->take_pinned((size_t)(n_docs + 1) * 8);
    size_t tok_cap = (size_t)(n_bytes / 3) + 4096;               // tokens <= bytes; typical bytes/4; grows if needed
    r->tok = h->take_pinned(tok_cap * 4);
    auto cleanup = [&](int rc) { h->give_pinned(r->tok); h->give_pinned(r->off); delete r; return rc; };
    if (!r->off.p || !r->tok.p) return cleanup(fail(B200BPE_ECUDA, "pinned allocation failed"));
    tok_cap = r->tok.cap / 4;

    auto enqueue_h2d = [&](size_t c) -> int {
        Slot &S = h->slots[c % b200bpe::N_SLOTS];
        CUDA_TRY(cudaStreamSynchronize(S.stream));               // slot free again (its chunk c-3 fully drained)
        const uint64_t lo = cut[c], hi = cut[c + 1], b0 = doc_off[lo], nb = doc_off[hi] - b0, nd = hi - lo;
        CUDA_TRY(S.w_text.ensure((size_t)nb + 64)); CUDA_TRY(S.w_docoff.ensure((size_t)nd + 2));
        CUDA_TRY(S.w_tokoff.ensure((size_t)nd + 2)); CUDA_TRY(S.w_out.ensure((size_t)nb + 64));
        CUDA_TRY(cudaEventRecord(S.ev[5], S.stream));
        if (nb) CUDA_TRY(cudaMemcpyAsync(S.w_text.p, text + b0, nb, cudaMemcpyHostToDevice, S.stream));
        CUDA_TRY(cudaMemcpyAsync(S.w_docoff.p, doc_off + lo, (nd + 1) * 8, cudaMemcpyHostToDevice, S.stream));
        if (b0) add_offset_kernel<<<(unsigned)((nd + 1 + 255) / 256), 256, 0, S.stream>>>(S.w_docoff.p, nd + 1, -(long long)b0);
        CUDA_TRY(cudaEventRecord(S.ev[6], S.stream));
        return B200BPE_OK;
    };

    float sum_ms[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}; uint32_t launches = 0;
    uint64_t token_base = 0;
    // finalise chunk c: wait for its kernels, then send its offsets + tokens home (async)
    auto drain = [&](size_t c) -> int {
        Slot &S = h->slots[c % b200bpe::N_SLOTS];
        const uint64_t lo = cut[c], hi = cut[c + 1], nd = hi - lo;
        int rc2 = finish_pipeline(h, S, S.stream);
        if (rc2) return rc2;
        float h2d = 0; cudaEventElapsedTime(&h2d, S.ev[5], S.ev[6]);
        const uint64_t nt = S.h_ctr->total_tokens;
        if (token_base + nt > tok_cap) {                         // rare: grow the pinned token buffer
            for (int i = 0; i < b200bpe::N_SLOTS; i++) CUDA_TRY(cudaStreamSynchronize(h->slots[i].stream));
            size_t want = (size_t)((token_base + nt) * 1.5) + 4096;
            PinnedBuf nbuf = h->take_pinned(want * 4);
            if (!nbuf.p) return fail(B200BPE_ECUDA, "pinned allocation failed");
            memcpy(nbuf.p, r->tok.p, (size_t)token_base * 4);
            h->give_pinned(r->tok); r->tok = nbuf; tok_cap = nbuf.cap / 4;
        }
        if (token_base) add_offset_kernel<<<(unsigned)((nd + 1 + 255) / 256), 256, 0, S.stream>>>(S.w_tokoff.p, nd + 1, (long long)token_base);
        CUDA_TRY(cudaEventRecord(S.ev[5], S.stream));
        CUDA_TRY(cudaMemcpyAsync((uint64_t *)r->off.p + lo, S.w_tokoff.p, (nd + 1) * 8, cudaMemcpyDeviceToHost, S.stream));
        if (nt) CUDA_TRY(cudaMemcpyAsync((uint32_t *)r->tok.p + token_base, S.w_out.p, nt * 4, cudaMemcpyDeviceToHost, S.stream));
        CUDA_TRY(cudaEventRecord(S.ev[6], S.stream));
        token_base += nt;
        for (int i = 0; i < 5; i++) sum_ms[i] += S.last_ms[i];
        sum_ms[7] += S.last_ms[7]; sum_ms[8] += S.last_ms[8]; sum_ms[5] += h2d; launches += S.last_launches;
        return B200BPE_OK;
    };
    auto fail_all = [&](int rc2) {
        for (int i = 0; i < b200bpe::N_SLOTS; i++) cudaStreamSynchronize(h->slots[i].stream);
        return cleanup(rc2);
    };
    int rc = enqueue_h2d(0);
    if (rc) return fail_all(rc);
    for (size_t c = 0; c < n_chunks; c++) {
        // slot (c+1)%3 last served chunk c-2, which was drained in the previous iteration
        if (c + 1 < n_chunks) { rc = enqueue_h2d(c + 1); if (rc) return fail_all(rc); }
        Slot &S = h->slots[c % b200bpe::N_SLOTS];
        const uint64_t lo = cut[c], hi = cut[c + 1], nb = doc_off[hi] - doc_off[lo], nd = hi - lo;
        rc = run_pipeline(h, S, S.w_text.p, nb, S.w_docoff.p, nd, S.w_out.p, S.w_tokoff.p, S.stream, single_piece, true);
        if (rc) return fail_all(rc);
        if (c >= 1) { rc = drain(c - 1); if (rc) return fail_all(rc); }   // overlaps with chunk c's kernels
    }
    rc = drain(n_chunks - 1);
    if (rc) return fail_all(rc);
    for (int i = 0; i < b200bpe::N_SLOTS; i++) CUDA_TRY(cudaStreamSynchronize(h->slots[i].stream));
    {   // D2H time of the last chunk only (the others overlap with later chunks' kernels)
        Slot &S = h->slots[(n_chunks - 1) % b200bpe::N_SLOTS];
        float d2h = 0; cudaEventElapsedTime(&d2h, S.ev[5], S.ev[6]); sum_ms[6] = d2h;
    }
    memcpy(h->last_ms, sum_ms, sizeof(sum_ms)); h->last_launches = launches;
    r->n_tokens = token_base;
    h->live_results++;                                            // caller holds h->mu
    *out = r;
    return B200BPE_OK;
}

extern "C" int b200bpe_encode_ordinary_batch(b200bpe_t *h, const uint8_t *text, const uint64_t *doc_off,
                                             uint64_t n_docs, b200bpe_result_t **out) {
    if (!h || !doc_off || !out) return fail(B200BPE_EINVAL, "null argument");
    if (doc_off[n_docs] && !text) return fail(B200BPE_EINVAL, "null text");
    std::lock_guard<std::mutex> lk(h->mu);
    return encode_host(h, text, doc_off, n_docs, false, out);
}

extern "C" int b200bpe_encode_single_piece(b200bpe_t *h, const uint8_t *piece, uint64_t len, b200bpe_result_t **out) {
    if (!h || !out || (len && !piece)) return fail(B200BPE_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    uint64_t off[2] = {0, len};
    return encode_host(h, piece, off, 1, true, out);
}

// CoreBPE::encode (src/lib.rs:375-442): allowed specials cut each document into haystacks on the
// host (specials are rare; a device multi-pattern scan is a "next" row), the haystacks are encoded
// as one batch on the device, and the special ids are spliced in while unpacking.
extern "C" int b200bpe_encode_batch(b200bpe_t *h, const uint8_t *text, const uint64_t *doc_off, uint64_t n_docs,
                                    const uint8_t *allowed, b200bpe_result_t **out) {
    if (!h || !doc_off || !out) return fail(B200BPE_EINVAL, "null argument");
    bool any = false;
    if (allowed) for (size_t i = 0; i < h->specials.size(); i++) any |= allowed[i] != 0;
    if (!any) return b200bpe_encode_ordinary_batch(h, text, doc_off, n_docs, out);
    std::lock_guard<std::mutex> lk(h->mu);
    // segment list: pseudo-documents = text between allowed specials; special bytes get
    // zero-length treatment by being skipped (they are excluded through a compacted copy).
    std::vector<uint8_t> ctext; ctext.reserve((size_t)doc_off[n_docs]);
    std::vector<uint64_t> seg_off; seg_off.push_back(0);
    struct Cut { uint64_t seg_index; uint32_t rank; };           // special emitted after segment seg_index
    std::vector<Cut> cuts; std::vector<uint64_t> doc_first_seg(n_docs + 1);
    std::vector<std::pair<std::string, uint32_t>> act;
    for (size_t i = 0; i < h->specials.size(); i++) if (allowed[i] && !h->specials[i].empty()) act.push_back({h->specials[i], h->special_rank[i]});
    std::vector<uint64_t> nxt(act.size());                       // next occurrence of each allowed special in the document
    for (uint64_t d = 0; d < n_docs; d++) {
        doc_first_seg[d] = seg_off.size() - 1;
        uint64_t s = doc_off[d], e = doc_off[d + 1], pos = s;
        // every special is searched once per stretch of text: its next occurrence is kept until the cursor
        // passes it (a special that does not occur any more is never searched again in this document)
        auto find_from = [&](size_t a, uint64_t from) -> uint64_t {
            const std::string &sp = act[a].first;
            if (from >= e || sp.size() > e - from) return e;
            const void *f = memmem(text + from, (size_t)(e - from), sp.data(), sp.size());
            return f ? (uint64_t)((const uint8_t *)f - text) : e;
        };
        for (size_t a = 0; a < act.size(); a++) nxt[a] = find_from(a, s);
        while (pos < e) {
            // leftmost allowed special at or after pos (longest on ties)
            uint64_t best_pos = e; size_t best = (size_t)-1;
            for (size_t a = 0; a < act.size(); a++) {
                if (nxt[a] < pos) nxt[a] = find_from(a, pos);
                if (nxt[a] >= e) continue;
                if (nxt[a] < best_pos || (nxt[a] == best_pos && act[a].first.size() > act[best].first.size())) { best_pos = nxt[a]; best = a; }
            }
            ctext.insert(ctext.end(), text + pos, text + best_pos);
            seg_off.push_back(ctext.size());
            if (best == (size_t)-1) { pos = e; break; }
            cuts.push_back({seg_off.size() - 2, act[best].second});
            pos = best_pos + act[best].first.size();
            if (pos >= e) { seg_off.push_back(ctext.size()); }    // trailing empty haystack after a final special
        }
        if (s == e) seg_off.push_back(ctext.size());
    }
    doc_first_seg[n_docs] = seg_off.size() - 1;
    b200bpe_result *seg = nullptr;
    uint8_t dummy = 0;
    int rc = encode_host(h, ctext.empty() ? &dummy : ctext.data(), seg_off.data(), seg_off.size() - 1, false, &seg);
    if (rc) return rc;
    const uint32_t *stok = (const uint32_t *)seg->tok.p; const uint64_t *soff = (const uint64_t *)seg->off.p;
    b200bpe_result *r = new b200bpe_result();
    r->owner = h; r->on_host_vec = true; r->n_docs = n_docs;
    r->vtok.reserve((size_t)seg->n_tokens + cuts.size()); r->voff.resize(n_docs + 1);
    size_t ci = 0;
    for (uint64_t d = 0; d < n_docs; d++) {
        r->voff[d] = r->vtok.size();
        for (uint64_t sgi = doc_first_seg[d]; sgi < doc_first_seg[d + 1]; sgi++) {
            r->vtok.insert(r->vtok.end(), stok + soff[sgi], stok + soff[sgi + 1]);
            if (ci < cuts.size() && cuts[ci].seg_index == sgi) { r->vtok.push_back(cuts[ci].rank); ci++; }
        }
    }
    r->voff[n_docs] = r->vtok.size(); r->n_tokens = r->vtok.size();
    h->give_pinned(seg->tok); h->give_pinned(seg->off); delete seg;
    h->live_results--;                                            // the internal segment result
    h->live_results++;
    *out = r;
    return B200BPE_OK;
}

extern "C" const uint32_t *b200bpe_result_tokens(const b200bpe_result_t *r) {
    return r->on_host_vec ? r->vtok.data() : (const uint32_t *)r->tok.p;
}
extern "C" const uint64_t *b200bpe_result_offsets(const b200bpe_result_t *r) {
    return r->on_host_vec ? r->voff.data() : (const uint64_t *)r->off.p;
}
extern "C" uint64_t b200bpe_result_n_tokens(const b200bpe_result_t *r) { return r->n_tokens; }
extern "C" uint64_t b200bpe_result_n_docs(const b200bpe_result_t *r) { return r->n_docs; }
extern "C" void b200bpe_result_free(b200bpe_result_t *r) {
    if (!r) return;
    b200bpe *h = r->owner;
    bool last = false;
    if (h) {
        std::lock_guard<std::mutex> lk(h->mu);
        if (!r->on_host_vec) { h->give_pinned(r->tok); h->give_pinned(r->off); }
        h->live_results--;
        last = h->dead && h->live_results == 0;
    }
    delete r;
    if (last) engine_teardown(h);
}

extern "C" int b200bpe_decode_bytes(b200bpe_t *h, const uint32_t *tokens, uint64_t n_tokens, uint8_t *out,
                                    uint64_t out_cap, uint64_t *out_len, uint32_t *bad_token) {
    if (!h || (n_tokens && !tokens) || !out_len) return fail(B200BPE_EINVAL, "null argument");
    uint64_t k = 0;
    for (uint64_t i = 0; i < n_tokens; i++) {
        const std::string *s;
        auto it = h->H.decoder.find(tokens[i]);
        if (it != h->H.decoder.end()) s = &it->second;
        else {
            auto it2 = h->special_decoder.find(tokens[i]);
            if (it2 == h->special_decoder.end()) {
                if (bad_token) *bad_token = tokens[i];
                return fail(B200BPE_EKEY, "Invalid token for decoding: " + std::to_string(tokens[i]));
            }
            s = &it2->second;
        }
        if (out && k + s->size() <= out_cap) memcpy(out + k, s->data(), s->size());
        k += s->size();
    }
    *out_len = k;
    return B200BPE_OK;
}

// Batched CoreBPE::decode_bytes (src/lib.rs:345-358) on the device: tokens of all documents
// concatenated + per-document token offsets (HOST buffers) -> bytes of all documents concatenated
// + per-document byte offsets.  The result object reuses b200bpe_result: "tokens" holds the bytes
// (n_tokens = byte count), "offsets" the byte offsets.
extern "C" int b200bpe_decode_batch(b200bpe_t *h, const uint32_t *tokens, const uint64_t *tok_off, uint64_t n_docs,
                                    b200bpe_result_t **out, uint32_t *bad_token) {
    if (!h || !tok_off || !out) return fail(B200BPE_EINVAL, "null argument");
    const uint64_t n = tok_off[n_docs];
    if (n && !tokens) return fail(B200BPE_EINVAL, "null tokens");
    std::lock_guard<std::mutex> lk(h->mu);
    CUDA_TRY(cudaSetDevice(h->device));
    Slot &S = h->slots[0];
    cudaStream_t st = S.stream;
    // reuse slot-0 workspace: w_out = tokens, w_sub_count = lengths, w_sub_base = byte base, w_text = bytes out
    CUDA_TRY(S.w_out.ensure((size_t)n + 4)); CUDA_TRY(S.w_sub_count.ensure((size_t)n + 4));
    CUDA_TRY(S.w_sub_base.ensure((size_t)n + 4)); CUDA_TRY(S.w_scan_part.ensure((size_t)(n / SCAN_ITEMS) + 4));
    CUDA_TRY(S.w_docoff.ensure((size_t)n_docs + 2)); CUDA_TRY(S.w_tokoff.ensure((size_t)n_docs + 2));
    CUDA_TRY(cudaMemsetAsync(S.d_ctr, 0, sizeof(Counters), st));
    CUDA_TRY(cudaMemsetAsync(&S.d_ctr->ticket, 0xFF, sizeof(unsigned int), st));
    CUDA_TRY(cudaEventRecord(S.ev[0], st));
    if (n) CUDA_TRY(cudaMemcpyAsync(S.w_out.p, tokens, n * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(S.w_docoff.p, tok_off, (n_docs + 1) * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaEventRecord(S.ev[1], st));
    const long long nn = (long long)n;
    const long long nb = (nn + SCAN_ITEMS - 1) / SCAN_ITEMS;
    if (n) decode_len_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(S.w_out.p, n, h->d_tok_boff, h->n_ids, S.w_sub_count.p, S.d_ctr);
    if (nb) scan_partial_kernel<<<(unsigned)nb, 256, 0, st>>>(S.w_sub_count.p, nn, S.w_scan_part.p);
    scan_top_kernel<<<1, 1024, 0, st>>>(S.w_scan_part.p, nb, S.d_ctr);
    if (nb) scan_final_kernel<<<(unsigned)nb, 256, 0, st>>>(S.w_sub_count.p, nn, S.w_scan_part.p, S.w_sub_base.p, S.d_ctr);
    else CUDA_TRY(cudaMemsetAsync(S.w_sub_base.p, 0, 8, st));
    CUDA_TRY(cudaMemcpyAsync(S.h_ctr, S.d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (S.h_ctr->err & ERR_BADTOKEN) {
        const uint32_t t = tokens[S.h_ctr->ticket];
        if (bad_token) *bad_token = t;
        return fail(B200BPE_EKEY, "Invalid token for decoding: " + std::to_string(t));
    }
    const uint64_t n_out = S.h_ctr->total_tokens;
    CUDA_TRY(S.w_text.ensure((size_t)n_out + 64));
    if (n) decode_copy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(S.w_out.p, n, h->d_tok_boff, h->n_ids, h->d_tok_blob,
                                                                        S.w_sub_base.p, S.w_text.p);
    decode_doc_off_kernel<<<(unsigned)((n_docs + 1 + 255) / 256), 256, 0, st>>>(S.w_docoff.p, n_docs, S.w_sub_base.p, S.w_tokoff.p);
    CUDA_TRY(cudaEventRecord(S.ev[2], st));
    b200bpe_result *r = new b200bpe_result();
    r->owner = h; r->n_docs = n_docs; r->n_tokens = n_out;
    r->off = h->take_pinned((size_t)(n_docs + 1) * 8);
    r->tok = h->take_pinned((size_t)n_out + 16);
    if (!r->off.p || !r->tok.p) { h->give_pinned(r->tok); h->give_pinned(r->off); delete r; return fail(B200BPE_ECUDA, "pinned allocation failed"); }
    CUDA_TRY(cudaMemcpyAsync(r->off.p, S.w_tokoff.p, (n_docs + 1) * 8, cudaMemcpyDeviceToHost, st));
    if (n_out) CUDA_TRY(cudaMemcpyAsync(r->tok.p, S.w_text.p, n_out, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaEventRecord(S.ev[3], st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    memset(h->last_ms, 0, sizeof(h->last_ms));
    cudaEventElapsedTime(&h->last_ms[5], S.ev[0], S.ev[1]);
    cudaEventElapsedTime(&h->last_ms[4], S.ev[1], S.ev[2]);
    cudaEventElapsedTime(&h->last_ms[6], S.ev[2], S.ev[3]);
    h->last_launches = 6;
    h->live_results++;
    *out = r;
    return B200BPE_OK;
}

extern "C" int b200bpe_last_timings(b200bpe_t *h, float *ms9, uint32_t *n_launches) {
    if (!h) return fail(B200BPE_EINVAL, "null handle");
    if (ms9) memcpy(ms9, h->last_ms, sizeof(h->last_ms));
    if (n_launches) *n_launches = h->last_launches;
    return B200BPE_OK;
}

extern "C" int b200bpe_table_bytes(b200bpe_t *h, uint64_t *bytes4) {
    if (!h || !bytes4) return fail(B200BPE_EINVAL, "null argument");
    memcpy(bytes4, h->table_bytes, sizeof(h->table_bytes));
    return B200BPE_OK;
}
