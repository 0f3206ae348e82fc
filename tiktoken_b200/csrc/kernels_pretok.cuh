// kernels_pretok.cuh -- documents -> doc-start bitmask; UTF-8 bytes -> piece-start bitmask (the regex pre-tokeniser,
// src/lib.rs:363-366 `find_iter` over the pat_strs of tiktoken_ext/openai_public.py:12-14, :89, :104-114).
#pragma once
#include "dev_common.cuh"
#include "pretok_fast.cuh"

using namespace b2bpe;

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
__global__ void __launch_bounds__(PRETOK_WARPS * 32, PAT == PAT_O200K ? 3 : 5) pretok_kernel(const uint8_t *__restrict__ text, long long n_bytes,
                                                                  const uint32_t *__restrict__ dbits, UcTables uc,
                                                                  uint32_t *__restrict__ pbits, uint32_t *__restrict__ psum,
                                                                  long long n_words, const uint32_t *__restrict__ ibits) {
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
    if (w < n_words) {
        word = span_word(t, w, b);
        if (ibits) word &= ~ibits[w];                       // no piece starts inside an accepted special token
        pbits[w] = word;
    }
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

