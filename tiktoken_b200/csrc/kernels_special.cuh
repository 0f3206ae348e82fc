// kernels_special.cuh -- special tokens on the device.
//
// Reference: CoreBPE::encode (src/lib.rs:375-442) looks for the next allowed special token with a regex over the
// escaped special strings (:389-401, built :625-631), encodes the text before it as its own haystack (:405-424) and
// pushes the special's id (:426-436); Encoding.encode first searches the text for any DISALLOWED special and raises
// (tiktoken/core.py:120-124, :431-438).  Here both are one multi-pattern scan over the packed batch:
//
//   special_mark_kernel     every position whose byte can start a special is hashed for each distinct special length
//                           and looked up in a small hash table of the specials (verified byte by byte, never across a
//                           document boundary).  Disallowed match -> ERR_SPECIAL with the leftmost position; allowed
//                           match -> candidate bit.
//   special_resolve_kernel  a match that starts inside an earlier accepted match does not count (the reference resumes
//                           its search at the END of a match): candidates are resolved left to right along chains of
//                           overlapping candidates.  Accepted specials become: two haystack boundaries (start, end) for
//                           the pre-tokeniser, an interior mask (no piece start inside), a bit in the special-piece mask
//                           and their id at ltok[start] -- the probe kernel emits that id for the piece.
#pragma once
#include <string>
#include <vector>

#include "dev_common.cuh"

using namespace b2bpe;

static const int SP_MAX_LENS = 32;

struct SpecialTables {            // device view
    const uint8_t *blob; const uint32_t *off; const uint32_t *rank; const uint32_t *table;
    uint32_t table_mask, n, max_len, n_lens, n_first;
    uint32_t lens[SP_MAX_LENS];   // distinct lengths, longest first (0 entries: linear scan over all specials)
    uint32_t first_mask[8];       // bytes that start a special
    uint32_t first[4];            // the distinct first bytes when there are at most four
};

struct SpecialHost {
    std::vector<uint8_t> blob; std::vector<uint32_t> off, rank, table;
    SpecialTables view;           // pointers filled per device by special_upload
};

B2_HD uint32_t special_hash(const uint8_t *p, uint32_t len) {
    uint32_t h = 0x811C9DC5u ^ (len * 0x9E3779B1u);
    for (uint32_t i = 0; i < len; i++) { h ^= p[i]; h *= 0x01000193u; }
    h ^= h >> 15;
    return h;
}

static void special_build(const std::vector<std::string> &names, const std::vector<uint32_t> &ranks, SpecialHost &S) {
    SpecialTables &v = S.view;
    memset(&v, 0, sizeof(v));
    S.off.assign(1, 0);
    std::vector<uint32_t> lens;
    bool firsts[256] = {false};
    for (size_t i = 0; i < names.size(); i++) {
        S.blob.insert(S.blob.end(), names[i].begin(), names[i].end());
        S.off.push_back((uint32_t)S.blob.size());
        S.rank.push_back(ranks[i]);
        const uint32_t len = (uint32_t)names[i].size();
        if (len == 0) continue;
        v.max_len = std::max(v.max_len, len);
        if (std::find(lens.begin(), lens.end(), len) == lens.end()) lens.push_back(len);
        const uint8_t b = (uint8_t)names[i][0];
        v.first_mask[b >> 5] |= 1u << (b & 31);
        firsts[b] = true;
    }
    std::sort(lens.begin(), lens.end(), [](uint32_t a, uint32_t b) { return a > b; });
    if (lens.size() <= (size_t)SP_MAX_LENS) { v.n_lens = (uint32_t)lens.size(); for (size_t i = 0; i < lens.size(); i++) v.lens[i] = lens[i]; }
    uint32_t nf = 0;
    for (int b = 0; b < 256; b++) if (firsts[b]) { if (nf < 4) v.first[nf] = (uint32_t)b; nf++; }
    v.n_first = nf;
    v.n = (uint32_t)names.size();
    uint32_t cap = 16;
    while (cap < 4 * (uint32_t)names.size() + 4) cap <<= 1;
    v.table_mask = cap - 1;
    S.table.assign(cap, 0);
    for (size_t i = 0; i < names.size(); i++) {
        if (names[i].empty()) continue;
        uint32_t s = special_hash((const uint8_t *)names[i].data(), (uint32_t)names[i].size()) & v.table_mask;
        while (S.table[s]) s = (s + 1) & v.table_mask;
        S.table[s] = (uint32_t)i + 1;
    }
    if (S.blob.empty()) S.blob.push_back(0);
    if (S.rank.empty()) S.rank.push_back(0);
}

static cudaError_t special_upload(const SpecialHost &S, uint8_t **arena, SpecialTables *out) {
    const size_t b0 = (S.blob.size() + 15) & ~(size_t)15, b1 = S.off.size() * 4, b2 = S.rank.size() * 4, b3 = S.table.size() * 4;
    const size_t o1 = b0, o2 = (o1 + b1 + 15) & ~(size_t)15, o3 = (o2 + b2 + 15) & ~(size_t)15, total = o3 + b3 + 16;
    cudaError_t e = cudaMalloc((void **)arena, total);
    if (e == cudaSuccess) e = cudaMemcpy(*arena, S.blob.data(), S.blob.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(*arena + o1, S.off.data(), b1, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(*arena + o2, S.rank.data(), b2, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(*arena + o3, S.table.data(), b3, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
    *out = S.view;
    out->blob = *arena; out->off = (const uint32_t *)(*arena + o1); out->rank = (const uint32_t *)(*arena + o2);
    out->table = (const uint32_t *)(*arena + o3);
    return cudaSuccess;
}

// bits (pos, pos + len) of the doc-start mask: does a document start strictly inside [pos, pos + len)?
__device__ __forceinline__ bool doc_start_inside(const uint32_t *__restrict__ dbits, long long pos, uint32_t len) {
    for (long long q = pos + 1; q < pos + (long long)len;) {
        const uint32_t w = dbits[q >> 5] >> (q & 31);
        const int avail = 32 - (int)(q & 31);
        const long long left = pos + (long long)len - q;
        const uint32_t m = left >= avail ? w : (w & ((1u << left) - 1u));
        if (m) return true;
        q += avail;
    }
    return false;
}

// Specials matching at `pos` (whole match inside one document).  Returns the longest ALLOWED one through idx / len
// (len 0: none) and reports whether any DISALLOWED one matches (dis_idx >= 0).
__device__ void special_match_at(const SpecialTables &sp, const uint8_t *__restrict__ flags, const uint8_t *__restrict__ text,
                                 long long n_bytes, const uint32_t *__restrict__ dbits, long long pos, int &idx, uint32_t &len,
                                 int &dis_idx) {
    idx = -1; len = 0; dis_idx = -1;
    auto consider = [&](uint32_t i, uint32_t l) {
        const uint8_t f = flags[i];
        if (f == 1 && l > len) { idx = (int)i; len = l; }
        else if (f == 2 && dis_idx < 0) dis_idx = (int)i;
    };
    if (sp.n_lens) {
        for (uint32_t k = 0; k < sp.n_lens; k++) {
            const uint32_t l = sp.lens[k];
            if (pos + (long long)l > n_bytes) continue;
            if (doc_start_inside(dbits, pos, l)) continue;
            uint32_t s = special_hash(text + pos, l) & sp.table_mask;
            for (;;) {
                const uint32_t e = sp.table[s];
                if (!e) break;
                const uint32_t i = e - 1, o = sp.off[i];
                if (sp.off[i + 1] - o == l) {
                    bool same = true;
                    for (uint32_t b = 0; b < l; b++) if (sp.blob[o + b] != text[pos + b]) { same = false; break; }
                    if (same) { consider(i, l); break; }
                }
                s = (s + 1) & sp.table_mask;
            }
        }
    } else {
        for (uint32_t i = 0; i < sp.n; i++) {
            const uint32_t o = sp.off[i], l = sp.off[i + 1] - o;
            if (l == 0 || pos + (long long)l > n_bytes) continue;
            bool same = true;
            for (uint32_t b = 0; b < l; b++) if (sp.blob[o + b] != text[pos + b]) { same = false; break; }
            if (same && !doc_start_inside(dbits, pos, l)) consider(i, l);
        }
    }
}

// positions of the span (32 bytes at w * 32) whose byte can start a special
__device__ __forceinline__ uint32_t special_first_bytes(const SpecialTables &sp, const uint8_t *__restrict__ text, long long n_bytes,
                                                        long long w) {
    const long long base = w * 32;
    if (base >= n_bytes) return 0;
    uint32_t m = 0;
    if (base + 32 <= (n_bytes & ~15ll) && sp.n_first <= 4) {
        const uint4 a = ld_stream_u4(text + base), b = ld_stream_u4(text + base + 16);
        const uint32_t W[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int k = 0; k < 8; k++) {
            uint32_t hit = 0;
            for (uint32_t f = 0; f < sp.n_first; f++) {
                const uint32_t x = W[k] ^ (sp.first[f] * 0x01010101u);
                hit |= (x - 0x01010101u) & ~x & 0x80808080u;              // zero byte of x (exact: borrows only through zero bytes matter below)
            }
            // the classic zero-byte test can flag the byte ABOVE a true zero byte: re-check exactly
            for (uint32_t hm = hit; hm;) {
                const int bit = __ffs(hm) - 1; hm &= hm - 1;
                const int byte = bit >> 3;
                const uint32_t v = (W[k] >> (8 * byte)) & 0xFFu;
                if ((sp.first_mask[v >> 5] >> (v & 31)) & 1u) m |= 1u << (4 * k + byte);
            }
        }
    } else {
        for (int j = 0; j < 32 && base + j < n_bytes; j++) {
            const uint32_t v = text[base + j];
            if ((sp.first_mask[v >> 5] >> (v & 31)) & 1u) m |= 1u << j;
        }
    }
    return m;
}

__global__ void __launch_bounds__(256) special_mark_kernel(const uint8_t *__restrict__ text, long long n_bytes,
                                                          const uint32_t *__restrict__ dbits, SpecialTables sp,
                                                          const uint8_t *__restrict__ flags, uint32_t *__restrict__ cbits,
                                                          long long n_words, Counters *ctr) {
    const long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    uint32_t cand = 0;
    for (uint32_t m = special_first_bytes(sp, text, n_bytes, w); m;) {
        const int j = __ffs(m) - 1; m &= m - 1;
        const long long pos = w * 32 + j;
        int idx, dis; uint32_t len;
        special_match_at(sp, flags, text, n_bytes, dbits, pos, idx, len, dis);
        if (dis >= 0) {                                          // leftmost disallowed special wins (core.py:120-124)
            atomicOr(&ctr->err, ERR_SPECIAL);
            atomicMax(&ctr->special_pos, ~(((unsigned long long)pos << 16) | (unsigned long long)(dis & 0xFFFF)));
        }
        if (len) cand |= 1u << j;
    }
    cbits[w] = cand;
}

__device__ __forceinline__ uint32_t special_len_at(const SpecialTables &sp, const uint8_t *flags, const uint8_t *text, long long n_bytes,
                                                   const uint32_t *dbits, long long pos, int *idx_out = nullptr) {
    int idx, dis; uint32_t len;
    special_match_at(sp, flags, text, n_bytes, dbits, pos, idx, len, dis);
    if (idx_out) *idx_out = idx;
    return len;
}

// nearest candidate position in [lo, hi), searching downwards from hi - 1; -1 if none
__device__ __forceinline__ long long prev_candidate(const uint32_t *__restrict__ cbits, long long lo, long long hi) {
    if (lo < 0) lo = 0;
    for (long long q = hi - 1; q >= lo;) {
        const uint32_t w = cbits[q >> 5];
        const int top = (int)(q & 31);
        const uint32_t m = top == 31 ? w : (w & ((2u << top) - 1u));
        if (m) { const long long p = (q & ~31ll) + (31 - __clz((int)m)); return p >= lo ? p : -1; }
        q = (q & ~31ll) - 1;
    }
    return -1;
}
// nearest candidate position in (lo, hi]; -1 if none
__device__ __forceinline__ long long next_candidate(const uint32_t *__restrict__ cbits, long long lo, long long hi) {
    for (long long q = lo + 1; q <= hi;) {
        const uint32_t w = cbits[q >> 5] >> (q & 31);
        if (w) { const long long p = q + (__ffs((int)w) - 1); return p <= hi ? p : -1; }
        q = (q & ~31ll) + 32;
    }
    return -1;
}

__global__ void __launch_bounds__(256) special_resolve_kernel(const uint8_t *__restrict__ text, long long n_bytes,
                                                             const uint32_t *__restrict__ dbits, SpecialTables sp,
                                                             const uint8_t *__restrict__ flags, const uint32_t *__restrict__ cbits,
                                                             uint32_t *hbits, uint32_t *ibits, uint32_t *__restrict__ sbits,
                                                             uint32_t *ltok, long long n_words, Counters *ctr) {
    const long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    uint32_t keep = 0;
    for (uint32_t m = cbits[w]; m;) {
        const int j = __ffs(m) - 1; m &= m - 1;
        const long long s = w * 32 + j;
        // head of the chain of overlapping candidates that reaches s
        long long head = s;
        for (;;) {
            long long c = prev_candidate(cbits, head - (long long)sp.max_len + 1, head);
            bool moved = false;
            while (c >= 0) {
                if (c + (long long)special_len_at(sp, flags, text, n_bytes, dbits, c) > head) { head = c; moved = true; break; }
                c = prev_candidate(cbits, head - (long long)sp.max_len + 1, c);
            }
            if (!moved) break;
        }
        // left to right from the head: a candidate counts iff it starts at or after the end of the last accepted one
        long long cur = head, last_end = 0; bool acc = true;
        for (;;) {
            acc = cur >= last_end;
            if (acc) last_end = cur + (long long)special_len_at(sp, flags, text, n_bytes, dbits, cur);
            if (cur == s) break;
            cur = next_candidate(cbits, cur, s);
        }
        if (!acc) continue;
        int idx = -1;
        const uint32_t len = special_len_at(sp, flags, text, n_bytes, dbits, s, &idx);
        keep |= 1u << j;
        ltok[s] = sp.rank[idx];
        atomicAdd(&ctr->n_cut, 1u);
        const long long e = s + (long long)len;                   // haystack boundaries at both ends
        atomicOr(&hbits[s >> 5], 1u << (s & 31));
        atomicOr(&hbits[e >> 5], 1u << (e & 31));
        for (long long q = s + 1; q < e;) {                       // no piece starts inside the special
            const int lo = (int)(q & 31);
            const long long left = e - q;
            const uint32_t bits = left >= 32 - lo ? (0xFFFFFFFFu << lo) : (((1u << left) - 1u) << lo);
            atomicOr(&ibits[q >> 5], bits);
            q += 32 - lo;
        }
    }
    sbits[w] = keep;
}

// last kernel of a pipeline: keep the error bits for b200bpe_device_wait, publish the counts for a count exchange
__global__ void finalize_kernel(const Counters *ctr, unsigned int *sticky, unsigned long long *d_counts, unsigned long long n_docs) {
    if (threadIdx.x == 0) {
        if (ctr->err) atomicOr(sticky, ctr->err);
        if (d_counts) { d_counts[0] = ctr->total_tokens; d_counts[1] = n_docs; }
    }
}
