// b200bpe.cu -- the engine and the C ABI of libb200bpe.so (see include/b200bpe.h).  sm_100a only.
//
// Path replaced: CoreBPE::encode_ordinary / CoreBPE::encode (src/lib.rs:360-442) and the per-call
// thread pool that fans documents out to it (tiktoken/core.py:164-206).  One call encodes the
// whole batch (kernels in kernels_*.cuh):
//
//   mark_docs_kernel      doc_off[] -> doc-start bitmask D + first-doc-per-span index
//   special_*_kernel      (CoreBPE::encode only) multi-pattern scan: disallowed specials -> error; allowed specials ->
//                         haystack boundaries, interior mask, special-piece mask + their ids
//   pretok_kernel<PAT>    UTF-8 bytes + D -> piece-start bitmask P   (bit-parallel regex rules)
//   find_long_kernel      P -> queue of pieces longer than 16 bytes + one work list per length class
//   mid_thread_kernel     17..256 bytes: one piece per lane, 32 pieces per warp in one convergent
//                         instruction stream, merge state in shared-memory columns
//   long_piece_kernel     257..4096 bytes: a warp per piece;  giant_piece_kernel ..32768 bytes: a block per piece;
//   cluster_piece_kernel  beyond: a thread-block cluster (8 x 1024 threads, DSMEM carries) per piece
//                         -- the round-synchronous exact merge in global scratch
//   probe_kernel          persistent warps, TMA-staged 1 KiB sub-tiles: whole-piece table probe of every short piece
//                         (one 32 B sector each), one slot per piece, misses -> global queue with their key bytes
//   miss_{hist,base,scatter}, miss_kernel   the ~5 % misses, sorted by length, one piece per lane,
//                         warp-convergent exact min-rank merge on dense records
//   scan_{partial,top,final}, gather_kernel, big_copy_kernel   token counts -> offsets -> tokens and
//                         per-document offsets at their final place
//
// No tensor cores: nothing here is a contraction.  The work is byte/integer, bound by HBM reads
// of the text, L2 probes of the rank tables and instruction issue.
//
// Host side: one DevCtx per CUDA device (tables replicated, three pipeline slots each); batches are cut at
// document boundaries into chunks that go round-robin over the devices; H2D of chunk c+1, kernels of chunk c and
// D2H of chunk c-1 overlap on every device, and every chunk's tokens land at their final offset of ONE pinned
// result buffer (the "gather" of SURVEY 8(e) is a host prefix sum over per-chunk counts).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200bpe.h"
#include "dev_common.cuh"
#include "kernels_pretok.cuh"
#include "kernels_long.cuh"
#include "kernels_mid.cuh"
#include "kernels_pmerge.cuh"
#include "kernels_encode.cuh"
#include "kernels_special.cuh"
#include "kernels_decode.cuh"
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

namespace {

// pat_strs of tiktoken_ext/openai_public.py:12-14, :89, :104-114
const char *R50K_PAT = R"('(?:[sdmt]|ll|ve|re)| ?\p{L}++| ?\p{N}++| ?[^\s\p{L}\p{N}]++|\s++$|\s+(?!\S)|\s)";
const char *CL100K_PAT =
    R"('(?i:[sdmt]|ll|ve|re)|[^\r\n\p{L}\p{N}]?+\p{L}++|\p{N}{1,3}+| ?[^\s\p{L}\p{N}]++[\r\n]*+|\s++$|\s*[\r\n]|\s+(?!\S)|\s)";
const char *O200K_PAT =
    R"([^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]*[\p{Ll}\p{Lm}\p{Lo}\p{M}]+(?i:'s|'t|'re|'ve|'m|'ll|'d)?|)"
    R"([^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]+[\p{Ll}\p{Lm}\p{Lo}\p{M}]*(?i:'s|'t|'re|'ve|'m|'ll|'d)?|)"
    R"(\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n/]*|\s*[\r\n]+|\s+(?!\S)|\s+)";

// the calling thread's current device is restored when an ABI call returns (torch and friends keep per-thread state)
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); } }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

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

long env_long(const char *name, long dflt, long lo, long hi) {
    const char *v = getenv(name);
    if (!v || !*v) return dflt;
    long x = atol(v);
    return x < lo ? lo : x > hi ? hi : x;
}

// Persistent helper threads of one engine: they widen the bit-packed tokens that come back over PCIe and stage pageable
// caller memory into pinned blocks, next to the device pipeline (the reference's counterpart is its rayon / thread pool).
struct TaskPool {
    std::vector<std::thread> th; std::mutex mu; std::condition_variable cv; std::deque<std::function<void()>> q; bool stop = false;
    void ensure(int n) {
        std::lock_guard<std::mutex> lk(mu);
        while ((int)th.size() < n) th.emplace_back([this] {
            for (;;) {
                std::function<void()> f;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [this] { return stop || !q.empty(); });
                    if (q.empty()) return;
                    f = std::move(q.front()); q.pop_front();
                }
                f();
            }
        });
    }
    void push(std::function<void()> f) { { std::lock_guard<std::mutex> lk(mu); q.push_back(std::move(f)); } cv.notify_one(); }
    ~TaskPool() {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv.notify_all();
        for (auto &t : th) t.join();
    }
};
struct Latch {
    std::mutex mu; std::condition_variable cv; int left;
    explicit Latch(int n) : left(n) {}
    void done() { std::lock_guard<std::mutex> lk(mu); if (--left == 0) cv.notify_all(); }
    void wait() { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [this] { return left == 0; }); }
};
// fn(lo, hi) over [0, n) in blocks, on the pool; the caller waits
void pool_for(TaskPool *pool, size_t n, size_t block, const std::function<void(size_t, size_t)> &fn) {
    const size_t nb = (n + block - 1) / block;
    if (!pool || nb <= 1) { if (n) fn(0, n); return; }
    Latch latch((int)nb);
    for (size_t b = 0; b < nb; b++) pool->push([&, b] { fn(b * block, std::min(n, (b + 1) * block)); latch.done(); });
    latch.wait();
}
// tokens come back from the device as `bits`-bit fields, little-endian bit order (pack_tokens_kernel); src is readable 8 bytes past the end
void unpack_tokens(const uint8_t *src, uint32_t *dst, size_t lo, size_t hi, int bits) {
    const uint64_t mask = (1ull << bits) - 1ull;
    size_t bitpos = lo * (size_t)bits;
    for (size_t i = lo; i < hi; i++, bitpos += (size_t)bits) {
        uint64_t x; memcpy(&x, src + (bitpos >> 3), 8);
        dst[i] = (uint32_t)((x >> (bitpos & 7)) & mask);
    }
}

}  // namespace

// One pipeline slot: a stream, its events and a grow-only workspace.  The host path keeps
// three slots in flight per device (H2D of chunk c+1, kernels of chunk c, D2H of chunk c-1).
struct Slot {
    DevBuf<uint8_t> w_text; DevBuf<unsigned long long> w_docoff, w_tokoff, w_sub_base;
    DevBuf<uint32_t> w_ptok, w_mres, w_mq_pos, w_mq_roff, w_doc_tiles, w_sub_count; DevBuf<uint8_t> w_mq_len;
    DevBuf<uint4> w_mq_key, w_mq_skey, w_mq_smeta, w_mq_rec;
    DevBuf<uint32_t> w_dbits, w_pbits, w_psum, w_sfd, w_lidx, w_out, w_ltok, w_hbits, w_ibits, w_sbits, w_cbits, w_slow;
    DevBuf<unsigned long long> w_lq_start, w_lq_off; DevBuf<unsigned int> w_lq_len, w_lq_ntok, w_lq_cls, w_big_n, w_sort_hist;
    DevBuf<unsigned long long> w_big_dst, w_big_src, w_scan_part;
    DevBuf<uint32_t> w_idA, w_rkA, w_idB, w_rkB, w_aux1, w_aux2; DevBuf<uint8_t> w_flag, w_spflags;
    size_t long_cap = 0;            // entries of the global merge scratch (pieces > 256 bytes), grown on ERR_LONGCAP
    size_t miss_cap = 0, mres_cap = 0;   // miss queue entries / miss result tokens, grown on ERR_MISSCAP
    size_t slow_cap = 0;            // positions left to the general pre-tokeniser rule function, grown on ERR_SLOWCAP
    Counters *d_ctr = nullptr; Counters *h_ctr = nullptr;   // h_ctr pinned
    unsigned int *d_sticky = nullptr;                       // error bits of every pipeline since the last wait
    cudaStream_t stream = nullptr;
    cudaStream_t side = nullptr;     // the group kernels (17..1024 bytes) run here, next to probe + miss on the main stream,
    cudaStream_t side2 = nullptr;    // and the scratch kernels (warp / block / cluster per piece: a few SMs each) here
    cudaStream_t up = nullptr;       // host path: uploads of the slot's NEXT chunk (ordered behind the kernels, not the download, of its last one)
    static const int N_EV = 15;       // [10] fork, [11] side start, [12] side end (join), [13] side2 end (join), [14] packed tokens on the host
    cudaEvent_t ev[N_EV];
    PinnedBuf stage;                // pinned staging for callers whose text is pageable memory
    DevBuf<uint32_t> w_pack;        // bit-packed tokens of the chunk (host path)
    PinnedBuf stage_out;            // ... and where they land on the host before the helper threads widen them
    std::atomic<int> stage_out_busy{0};
    float last_ms[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t last_launches = 0;
    bool ok = false;

    cudaError_t init() {
        cudaError_t e = cudaMalloc((void **)&d_ctr, sizeof(Counters));
        if (e == cudaSuccess) e = cudaMalloc((void **)&d_sticky, 16);
        if (e == cudaSuccess) e = cudaMemset(d_sticky, 0, 16);
        if (e == cudaSuccess) e = cudaHostAlloc((void **)&h_ctr, sizeof(Counters), cudaHostAllocPortable);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&side2, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking);
        for (int i = 0; i < N_EV && e == cudaSuccess; i++) e = cudaEventCreate(&ev[i]);
        ok = (e == cudaSuccess);
        return e;
    }
    void release_workspace() {
        w_text.release(); w_docoff.release(); w_tokoff.release(); w_sub_base.release();
        w_ptok.release(); w_mres.release(); w_mq_pos.release(); w_mq_roff.release(); w_doc_tiles.release(); w_sub_count.release();
        w_mq_len.release(); w_mq_rec.release(); w_mq_key.release(); w_mq_skey.release(); w_mq_smeta.release();
        w_dbits.release(); w_pbits.release(); w_psum.release(); w_sfd.release(); w_lidx.release(); w_out.release(); w_ltok.release();
        w_hbits.release(); w_ibits.release(); w_sbits.release(); w_cbits.release(); w_slow.release(); w_spflags.release();
        w_lq_start.release(); w_lq_off.release(); w_lq_len.release(); w_lq_ntok.release(); w_lq_cls.release(); w_big_n.release();
        w_sort_hist.release(); w_big_dst.release(); w_big_src.release(); w_scan_part.release();
        w_idA.release(); w_rkA.release(); w_idB.release(); w_rkB.release(); w_aux1.release(); w_aux2.release();
        w_flag.release(); w_pack.release();
        if (stage.p) cudaFreeHost(stage.p);
        stage.p = nullptr; stage.cap = 0;
        if (stage_out.p) cudaFreeHost(stage_out.p);
        stage_out.p = nullptr; stage_out.cap = 0;
        long_cap = miss_cap = mres_cap = slow_cap = 0;
    }
    void destroy() {
        release_workspace();
        if (d_ctr) cudaFree(d_ctr);
        if (d_sticky) cudaFree(d_sticky);
        if (h_ctr) cudaFreeHost(h_ctr);
        if (ok) { for (int i = 0; i < N_EV; i++) cudaEventDestroy(ev[i]); }
        if (stream) cudaStreamDestroy(stream);
        if (side) cudaStreamDestroy(side);
        if (side2) cudaStreamDestroy(side2);
        if (up) cudaStreamDestroy(up);
    }
};

#ifndef B2_N_SLOTS
#define B2_N_SLOTS 3          // pipeline slots per device: uploads run B2_N_SLOTS - 2 chunks ahead of the kernels
#endif
// Everything that lives on one CUDA device: the replicated tables and the pipeline slots.
struct DevCtx {
    int device = 0;
    uint8_t *arena = nullptr; size_t arena_bytes = 0, hot_bytes = 0;   // all tables in one allocation (one L2 window)
    DevTables T; UcTables uc;
    uint32_t *d_tok_boff = nullptr; uint8_t *d_tok_blob = nullptr;     // decode: id -> bytes
    SpecialTables sp;                                                  // device copy of the special-token patterns
    uint8_t *d_sp_arena = nullptr;
    static const int N_SLOTS = B2_N_SLOTS;
    Slot slots[B2_N_SLOTS];
    size_t l2_window = 0; float l2_ratio = 0.f;
    int probe_blocks_per_sm = 10;
};

struct b200bpe_result {
    b200bpe *owner = nullptr;
    PinnedBuf tok, off;                 // pinned when produced by the engine
    std::vector<uint32_t> vtok;         // used when the result is assembled on the host (fallback special path)
    std::vector<uint64_t> voff;
    bool on_host_vec = false;
    uint64_t n_tokens = 0, n_docs = 0;
};

struct PendingDeviceCall {             // the most recent b200bpe_encode_device_async call (for b200bpe_device_wait)
    bool active = false; int queued = 0;
    const uint8_t *d_text = nullptr; uint64_t n_bytes = 0; const unsigned long long *d_doc_off = nullptr; uint64_t n_docs = 0;
    uint32_t *d_tokens = nullptr; unsigned long long *d_tok_off = nullptr; unsigned long long *d_counts = nullptr;
    cudaStream_t st = nullptr;
};

struct b200bpe {
    int pattern = 0;
    HostTables H;
    std::vector<std::string> specials; std::vector<uint32_t> special_rank;
    std::unordered_map<uint32_t, std::string> special_decoder;
    SpecialHost sp_host;                // hash table of the specials (built once), uploaded to every device
    uint32_t n_ids = 0; bool decode_on_device = true;
    std::vector<DevCtx *> devs;
    uint64_t table_bytes[4] = {0, 0, 0, 0};
    float last_ms[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t last_launches = 0;
    size_t chunk_bytes = 64u << 20; bool chunk_forced = false;
    int copy_threads = 4;
    TaskPool *pool = nullptr;        // helper threads (created with the first host-path call)
    int pack_bits = 0;               // host path: tokens cross PCIe as fields of this many bits (0: plain u32)
    bool mid_group = true;           // 17..1024-byte pieces: group-of-lanes kernels (need ranks < 2^22)
    bool pmerge = true;              // 129..1024-byte pieces: segmented parallel merge (needs ranks < 2^22); off: the group-of-lanes kernels
    int pmerge_min_cls = 3;          // shortest length class the parallel merge takes (3: 129..256 bytes; 0: everything from 17 bytes)
    std::mutex mu;
    std::vector<PinnedBuf> pinned_pool;
    // results keep the engine alive: b200bpe_destroy with results outstanding only marks the handle dead, the last
    // b200bpe_result_free tears it down (the reference's TiktokenBuffer owns its Vec, src/py.rs:186-189)
    int live_results = 0;
    bool dead = false;
    PendingDeviceCall pending;

    PinnedBuf take_pinned(size_t bytes) {
        for (size_t i = 0; i < pinned_pool.size(); i++)
            if (pinned_pool[i].cap >= bytes && pinned_pool[i].cap <= 2 * bytes + (64u << 20)) {
                PinnedBuf b = pinned_pool[i]; pinned_pool.erase(pinned_pool.begin() + i); return b;
            }
        PinnedBuf b; size_t want = bytes + bytes / 8 + 4096;
        if (cudaHostAlloc(&b.p, want, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); b.p = nullptr; b.cap = 0; return b; }
        b.cap = want; return b;
    }
    void give_pinned(PinnedBuf b) {
        if (!b.p) return;
        if (pinned_pool.size() >= 4) {                     // keep the largest blocks
            size_t mn = 0;
            for (size_t i = 1; i < pinned_pool.size(); i++) if (pinned_pool[i].cap < pinned_pool[mn].cap) mn = i;
            if (pinned_pool[mn].cap >= b.cap) { cudaFreeHost(b.p); return; }
            cudaFreeHost(pinned_pool[mn].p); pinned_pool[mn] = b; return;
        }
        pinned_pool.push_back(b);
    }
};

extern "C" const char *b200bpe_last_error(void) { return g_last_error.c_str(); }
extern "C" const char *b200bpe_version(void) { return "b200bpe 0.2 (sm_100a)"; }
extern "C" int b200bpe_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// --------------------------------------------------------------------------------------------
// construction
// --------------------------------------------------------------------------------------------
static void devctx_destroy(DevCtx *D) {
    if (!D) return;
    cudaSetDevice(D->device);
    for (int i = 0; i < DevCtx::N_SLOTS; i++) D->slots[i].destroy();
    if (D->arena) cudaFree(D->arena);
    if (D->d_tok_boff) cudaFree(D->d_tok_boff);
    if (D->d_tok_blob) cudaFree(D->d_tok_blob);
    if (D->d_sp_arena) cudaFree(D->d_sp_arena);
    delete D;
}

static void apply_l2_window(DevCtx *D, cudaStream_t st) {
    if (!D->l2_window) return;
    cudaStreamAttrValue v;
    memset(&v, 0, sizeof(v));
    v.accessPolicyWindow.base_ptr = D->arena;
    v.accessPolicyWindow.num_bytes = D->l2_window;
    v.accessPolicyWindow.hitRatio = D->l2_ratio;
    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) cudaGetLastError();
}

static int devctx_create(b200bpe *h, int device, const std::vector<uint32_t> &boff, const std::vector<uint8_t> &blob,
                         DevCtx **out) {
    DevCtx *D = new DevCtx();
    D->device = device;
    const HostTables &H = h->H;
    cudaError_t e = cudaSetDevice(device);
    uint8_t ascii[128];
    for (int i = 0; i < 128; i++) ascii[i] = UC_STAGE2[(uint32_t)UC_STAGE1[0] * 256 + i];
    // one arena, hottest tables first: the L2 persistence window covers a prefix of it
    struct Part { const void *src; size_t bytes; size_t off; };
    Part parts[9] = {{H.piece_tab.data(), H.piece_tab.size() * sizeof(U4), 0}, {H.pair2.data(), 65536 * 4, 0},
                     {H.byte_id.data(), 256 * 4, 0}, {H.pair_tab.data(), H.pair_tab.size() * sizeof(U4), 0},
                     {ascii, 128, 0}, {UC_STAGE1, sizeof(UC_STAGE1), 0}, {UC_STAGE2, sizeof(UC_STAGE2), 0},
                     {H.long_tab.data(), H.long_tab.size() * sizeof(U4), 0}, {H.long_blob.data(), H.long_blob.size(), 0}};
    size_t total = 0;
    for (auto &p : parts) { p.off = total; total += (p.bytes + 255) & ~(size_t)255; }
    D->arena_bytes = total; D->hot_bytes = parts[4].off;
    if (e == cudaSuccess) e = cudaMalloc((void **)&D->arena, total);
    for (auto &p : parts) if (e == cudaSuccess && p.bytes) e = cudaMemcpy(D->arena + p.off, p.src, p.bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc((void **)&D->d_tok_boff, boff.size() * 4);
    if (e == cudaSuccess) e = cudaMemcpy(D->d_tok_boff, boff.data(), boff.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc((void **)&D->d_tok_blob, blob.size());
    if (e == cudaSuccess) e = cudaMemcpy(D->d_tok_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = special_upload(h->sp_host, &D->d_sp_arena, &D->sp);
    for (int i = 0; i < DevCtx::N_SLOTS && e == cudaSuccess; i++) e = D->slots[i].init();
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mid_thread_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MID_SMEM_BYTES);
    if (e == cudaSuccess) {
        const long carve = env_long("B200BPE_PROBE_CARVEOUT", -1, -1, 100);
        if (carve >= 0) e = cudaFuncSetAttribute(probe_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)carve);
        const long mcarve = env_long("B200BPE_MISS_CARVEOUT", -1, -1, 100);
        if (e == cudaSuccess && mcarve >= 0) e = cudaFuncSetAttribute(miss_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)mcarve);
        D->probe_blocks_per_sm = (int)env_long("B200BPE_PROBE_BLOCKS", 10, 1, 16);
    }
    if (e != cudaSuccess) { std::string m = cudaGetErrorString(e); devctx_destroy(D); return fail(B200BPE_ECUDA, "device " + std::to_string(device) + " setup: " + m); }
    D->T.piece_tab = (const U4 *)(D->arena + parts[0].off); D->T.piece_mask = H.piece_mask;
    D->T.pair2 = (const uint32_t *)(D->arena + parts[1].off);
    D->T.byte_id = (const uint32_t *)(D->arena + parts[2].off);
    D->T.pair_tab = (const U4 *)(D->arena + parts[3].off); D->T.pair_mask = H.pair_mask;
    D->uc.ascii = D->arena + parts[4].off;
    D->uc.stage1 = (const uint16_t *)(D->arena + parts[5].off);
    D->uc.stage2 = D->arena + parts[6].off;
    D->uc.one = 1u;
    D->T.long_tab = (const U4 *)(D->arena + parts[7].off); D->T.long_mask = H.long_mask;
    D->T.long_blob = D->arena + parts[8].off;
    D->T.max_token_len = H.max_token_len; D->T.n_long_tokens = H.n_long_tokens;
    // L2 persistence (off by default): a persisting access-policy window over the rank tables was measured on B200 at
    // 1 GiB of English text -- probe_kernel 2.78 ms with the window vs 2.60 ms without, whole step 8.64 vs 8.43 ms
    // (profiles/r02_b_knobs.txt): the tables already stay in the 126 MB L2 (they are re-touched every few microseconds)
    // and the set-aside only shrinks what the streams can use.  B200BPE_L2_PERSIST=1 turns it on for experiments.
    if (env_long("B200BPE_L2_PERSIST", 0, 0, 1)) {
        int max_win = 0, max_persist = 0;
        cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, device);
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, device);
        if (max_win > 0 && max_persist > 0) {
            const size_t want = std::min((size_t)max_persist, D->hot_bytes);
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) {
                D->l2_window = std::min(D->hot_bytes, (size_t)max_win);
                D->l2_ratio = (float)std::min(1.0, (double)want / (double)D->l2_window);
            } else cudaGetLastError();
        }
    }
    for (int i = 0; i < DevCtx::N_SLOTS; i++) apply_l2_window(D, D->slots[i].stream);
    *out = D;
    return B200BPE_OK;
}

extern "C" int b200bpe_create_multi(const uint8_t *tok_bytes, const uint64_t *tok_off, const uint32_t *tok_rank,
                                    uint32_t n_tok, const uint8_t *sp_bytes, const uint64_t *sp_off,
                                    const uint32_t *sp_rank, uint32_t n_sp, const char *pat_str, const int *devices,
                                    int n_dev, b200bpe_t **out) {
    if (!out || !pat_str || (n_tok && (!tok_bytes || !tok_off || !tok_rank)) || !devices || n_dev < 1)
        return fail(B200BPE_EINVAL, "null argument");
    int pattern;
    if (strcmp(pat_str, R50K_PAT) == 0) pattern = PAT_R50K;
    else if (strcmp(pat_str, CL100K_PAT) == 0) pattern = PAT_CL100K;
    else if (strcmp(pat_str, O200K_PAT) == 0) pattern = PAT_O200K;
    else return fail(B200BPE_EPATTERN,
                     "unsupported pat_str: the B200 pre-tokeniser implements exactly the r50k/p50k, cl100k and "
                     "o200k patterns of tiktoken_ext/openai_public.py (there is no CPU regex fallback)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(B200BPE_ECUDA, "no CUDA device: libb200bpe has no CPU fallback");
    }
    for (int i = 0; i < n_dev; i++) {
        if (devices[i] < 0 || devices[i] >= ndev) return fail(B200BPE_EINVAL, "bad device index");
        for (int j = 0; j < i; j++) if (devices[j] == devices[i]) return fail(B200BPE_EINVAL, "duplicate device index");
    }
    DeviceGuard guard;
    b200bpe *h = new b200bpe();
    h->pattern = pattern;
    int rc = build_tables(tok_bytes, tok_off, tok_rank, n_tok, h->H, (uint32_t)env_long("B200BPE_PAIR_SLACK", 3, 2, 16));
    if (rc) { std::string m = h->H.error; delete h; return fail(rc == -3 ? B200BPE_EDUPRANK : B200BPE_EINVAL, m); }
    for (uint32_t i = 0; i < n_sp; i++) {
        std::string s((const char *)sp_bytes + sp_off[i], (size_t)(sp_off[i + 1] - sp_off[i]));
        h->specials.push_back(s); h->special_rank.push_back(sp_rank[i]);
        h->special_decoder[sp_rank[i]] = s;
    }
    special_build(h->specials, h->special_rank, h->sp_host);
    const HostTables &H = h->H;
    // decode tables: byte offsets by token id (mergeable ranks and specials).  Ids of 2^24 and above do not get a
    // dense table: with any such id the batched decode runs on the host maps (same results, KeyError included).
    uint32_t max_id = 0; bool any = false;
    for (auto &kv : H.decoder) { if (kv.first >= (1u << 24)) h->decode_on_device = false; else { max_id = std::max(max_id, kv.first); any = true; } }
    for (auto &kv : h->special_decoder) { if (kv.first >= (1u << 24)) h->decode_on_device = false; else { max_id = std::max(max_id, kv.first); any = true; } }
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
    for (int i = 0; i < n_dev; i++) {
        DevCtx *D = nullptr;
        rc = devctx_create(h, devices[i], boff, blob, &D);
        if (rc) { for (auto *d : h->devs) devctx_destroy(d); delete h; return rc; }
        h->devs.push_back(D);
    }
    if (const char *cm = getenv("B200BPE_CHUNK_MB")) { long v = atol(cm); if (v >= 1 && v <= 2048) { h->chunk_bytes = (size_t)v << 20; h->chunk_forced = true; } }
    h->mid_group = H.max_rank < MIDG_MAX_RANK && env_long("B200BPE_MID_GROUP", 1, 0, 1) != 0;
    h->pmerge = h->mid_group && env_long("B200BPE_PMERGE", 1, 0, 1) != 0;
    h->pmerge_min_cls = (int)env_long("B200BPE_PMERGE_MIN_CLS", 3, 1, 3);
    h->copy_threads = (int)env_long("B200BPE_COPY_THREADS", std::max(1u, std::min(16u, std::thread::hardware_concurrency() / 4)), 1, 64);
    {   // tokens return over PCIe as bit fields just wide enough for the largest id (17 bits for cl100k, 18 for o200k, 16 for
        // r50k / p50k instead of 32): the return traffic shares the link with the text going up
        uint32_t max_id = H.max_rank;
        for (uint32_t r : h->special_rank) max_id = std::max(max_id, r);
        int bits = 1;
        while (bits < 32 && (max_id >> bits)) bits++;
        bits = std::max(bits, 8);
        h->pack_bits = (bits <= 24 && env_long("B200BPE_PACK", 0, 0, 1)) ? bits : 0;   // opt-in: see DESIGN 4 (measured: no gain)
    }
    h->table_bytes[0] = H.piece_tab.size() * sizeof(U4);
    h->table_bytes[1] = H.pair_tab.size() * sizeof(U4) + 65536 * 4 + 1024;
    h->table_bytes[2] = H.long_tab.size() * sizeof(U4) + H.long_blob.size();
    h->table_bytes[3] = sizeof(UC_STAGE1) + sizeof(UC_STAGE2) + 128;
    *out = h;
    return B200BPE_OK;
}

extern "C" int b200bpe_create(const uint8_t *tok_bytes, const uint64_t *tok_off, const uint32_t *tok_rank,
                              uint32_t n_tok, const uint8_t *sp_bytes, const uint64_t *sp_off,
                              const uint32_t *sp_rank, uint32_t n_sp, const char *pat_str, int device,
                              b200bpe_t **out) {
    return b200bpe_create_multi(tok_bytes, tok_off, tok_rank, n_tok, sp_bytes, sp_off, sp_rank, n_sp, pat_str, &device, 1, out);
}

static void engine_teardown(b200bpe *h) {
    DeviceGuard guard;
    for (auto *d : h->devs) devctx_destroy(d);
    for (auto &b : h->pinned_pool) cudaFreeHost(b.p);
    delete h->pool;
    delete h;
}

extern "C" void b200bpe_destroy(b200bpe_t *h) {
    if (!h) return;
    {
        std::lock_guard<std::mutex> lk(h->mu);
        if (h->live_results > 0) { h->dead = true; return; }
    }
    engine_teardown(h);
}

extern "C" int b200bpe_n_devices(b200bpe_t *h) { return h ? (int)h->devs.size() : 0; }

// Give the grow-only work-spaces (about 25 bytes of device memory per input byte of the largest batch seen, per pipeline
// slot) and the pooled pinned result blocks back; the tables stay.  The next call re-allocates what it needs.
extern "C" int b200bpe_trim(b200bpe_t *h) {
    if (!h) return fail(B200BPE_EINVAL, "null handle");
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->pending.active) return fail(B200BPE_EINVAL, "a device call is in flight: b200bpe_device_wait first");
    DeviceGuard guard;
    for (auto *D : h->devs) {
        CUDA_TRY(cudaSetDevice(D->device));
        for (int i = 0; i < DevCtx::N_SLOTS; i++) {
            CUDA_TRY(cudaStreamSynchronize(D->slots[i].stream));
            CUDA_TRY(cudaStreamSynchronize(D->slots[i].side));
            CUDA_TRY(cudaStreamSynchronize(D->slots[i].side2));
            D->slots[i].release_workspace();
        }
    }
    for (auto &b : h->pinned_pool) cudaFreeHost(b.p);
    h->pinned_pool.clear();
    return B200BPE_OK;
}

// --------------------------------------------------------------------------------------------
// the device pipeline
// --------------------------------------------------------------------------------------------
struct PipeArgs {
    const uint8_t *d_text = nullptr; uint64_t n_bytes = 0;
    const unsigned long long *d_doc_off = nullptr; uint64_t n_docs = 0;
    uint32_t *d_out = nullptr; unsigned long long *d_tok_off = nullptr;
    unsigned long long *d_counts = nullptr;      // optional device u64[2]: {n_tokens, n_docs} (for a count exchange)
    cudaStream_t st = nullptr;
    bool single_piece = false;
    const uint8_t *sp_flags = nullptr;           // host: per special 1 = allowed, 2 = disallowed (NULL: no special handling)
};

static const unsigned long long MISS_CAP_MIN = 1u << 16, LONG_CAP_MIN = 1u << 20;

// Enqueue the whole pipeline on args.st.  All pointers are device pointers on D->device; d_text must be 16-byte
// aligned and readable up to n_bytes + 16.  No host synchronisation.
static int enqueue_pipeline(b200bpe *h, DevCtx *D, Slot &S, const PipeArgs &a) {
    const uint64_t n_bytes = a.n_bytes, n_docs = a.n_docs;
    cudaStream_t st = a.st;
    if (n_bytes >= (1ull << 32) - 4096) return fail(B200BPE_EINVAL, "batch too large for one call (>= 4 GiB)");
    if (n_docs >= 0xFFFFFFFEull) return fail(B200BPE_EINVAL, "too many documents in one call");
    const long long n_words = (long long)((n_bytes + 1 + 31) / 32);
    const long long n_tiles = (n_words + 31) / 32;                  // 1 KiB sub-tiles
    const size_t pb_words = (size_t)n_tiles * 32 + STAGE_PW + 8;    // the probe stages 36 words per sub-tile
    CUDA_TRY(S.w_dbits.ensure((size_t)n_words + 8));
    CUDA_TRY(S.w_pbits.ensure(pb_words));
    CUDA_TRY(S.w_psum.ensure((size_t)(n_words >> 5) + 4));
    CUDA_TRY(S.w_sfd.ensure((size_t)n_words + 4));
    CUDA_TRY(S.w_sub_base.ensure((size_t)n_tiles + 2));
    CUDA_TRY(S.w_scan_part.ensure((size_t)(n_tiles / SCAN_ITEMS) + 4));
    CUDA_TRY(S.w_sub_count.ensure((size_t)n_tiles + 2));
    CUDA_TRY(S.w_ptok.ensure((size_t)n_tiles * SUB_BYTES + 64));
    {   // miss queue: sized from experience (one miss per 8 bytes, two result tokens per 4 bytes), grown on ERR_MISSCAP
        const size_t want_q = std::max<size_t>((size_t)(n_bytes / 8) + 4096, MISS_CAP_MIN);
        const size_t want_r = std::max<size_t>((size_t)(n_bytes / 2) + 4096, MISS_CAP_MIN);
        if (S.miss_cap < want_q) S.miss_cap = want_q;
        if (S.mres_cap < want_r) S.mres_cap = want_r;
        const size_t mcap = S.miss_cap;
        CUDA_TRY(S.w_mq_pos.ensure(mcap)); CUDA_TRY(S.w_mq_roff.ensure(mcap)); CUDA_TRY(S.w_mq_rec.ensure(mcap));
        CUDA_TRY(S.w_mq_len.ensure(mcap));
        CUDA_TRY(S.w_mq_key.ensure(mcap)); CUDA_TRY(S.w_mq_skey.ensure(mcap)); CUDA_TRY(S.w_mq_smeta.ensure(mcap));
        CUDA_TRY(S.w_mres.ensure(S.mres_cap + 64));
        CUDA_TRY(S.w_sort_hist.ensure((size_t)SORT_BLOCKS * 17 + 32));
    }
    {   // undecided pre-tokeniser positions: under 2 % on the worst corpus seen; sized for 12 %, grown on ERR_SLOWCAP
        const size_t want = std::max<size_t>((size_t)(n_bytes / 8) + 4096, MISS_CAP_MIN);
        if (S.slow_cap < want) S.slow_cap = want;
        CUDA_TRY(S.w_slow.ensure(S.slow_cap));
    }
    CUDA_TRY(S.w_lidx.ensure((size_t)(n_bytes >> 4) + 4));
    CUDA_TRY(S.w_ltok.ensure((size_t)n_bytes + 4));
    const size_t qcap = (size_t)(n_bytes / (SHORT_MAX + 1)) + 4;
    CUDA_TRY(S.w_lq_start.ensure(qcap)); CUDA_TRY(S.w_lq_off.ensure(qcap));
    CUDA_TRY(S.w_lq_len.ensure(qcap)); CUDA_TRY(S.w_lq_ntok.ensure(qcap));
    size_t cls_cap[N_CLS], cls_total = 0;                          // per-class index lists, back to back
    {
        const size_t min_len[N_CLS] = {SHORT_MAX + 1, 33, 65, 129, 257, GROUP_MAX + 1, BLOCK_MIN + 1, CLUSTER_MIN + 1};
        for (int c = 0; c < N_CLS; c++) { cls_cap[c] = (size_t)(n_bytes / min_len[c]) + 4; cls_total += cls_cap[c]; }
    }
    CUDA_TRY(S.w_lq_cls.ensure(cls_total));
    CUDA_TRY(S.w_big_n.ensure((size_t)(n_bytes / 4097) + 4));
    CUDA_TRY(S.w_big_dst.ensure((size_t)(n_bytes / 4097) + 4)); CUDA_TRY(S.w_big_src.ensure((size_t)(n_bytes / 4097) + 4));
    {   // global merge scratch of the pieces > 256 bytes: sized from experience, grown on ERR_LONGCAP
        const size_t want = std::max<size_t>((size_t)(n_bytes / 8) + 4096, LONG_CAP_MIN);
        if (S.long_cap < want) S.long_cap = want;
        const size_t lc = S.long_cap + 4;
        CUDA_TRY(S.w_idA.ensure(lc)); CUDA_TRY(S.w_rkA.ensure(lc)); CUDA_TRY(S.w_idB.ensure(lc)); CUDA_TRY(S.w_rkB.ensure(lc));
        CUDA_TRY(S.w_aux1.ensure(lc)); CUDA_TRY(S.w_aux2.ensure(lc)); CUDA_TRY(S.w_flag.ensure(lc));
    }
    LongQ q;
    q.start = S.w_lq_start.p; q.len = S.w_lq_len.p; q.off = S.w_lq_off.p; q.ntok = S.w_lq_ntok.p; q.scratch_cap = S.long_cap;
    q.sub_count = S.w_sub_count.p;
    { size_t o = 0; for (int c = 0; c < N_CLS; c++) { q.cls[c] = S.w_lq_cls.p + o; o += cls_cap[c]; } }
    // documents are sparse when fewer than one sub-tile in four can hold a document start: the gather then takes the
    // doc-start sub-tiles from a list instead of giving every sub-tile the shared memory their bookkeeping needs
    const bool sparse_docs = (n_docs + 1) * 4 < (uint64_t)n_tiles;
    if (sparse_docs) CUDA_TRY(S.w_doc_tiles.ensure((size_t)n_docs + 8));
    uint32_t launches = 0;

    CUDA_TRY(cudaEventRecord(S.ev[0], st));
    CUDA_TRY(cudaMemsetAsync(S.d_ctr, 0, sizeof(Counters), st));
    CUDA_TRY(cudaMemsetAsync(S.w_dbits.p, 0, ((size_t)n_words + 8) * 4, st));
    CUDA_TRY(cudaMemsetAsync(S.w_sfd.p, 0xFF, ((size_t)n_words + 4) * 4, st));
    CUDA_TRY(cudaMemsetAsync(S.w_sub_count.p, 0, ((size_t)n_tiles + 2) * 4, st));
    CUDA_TRY(cudaMemsetAsync(S.w_pbits.p + n_words, 0, (pb_words - (size_t)n_words) * 4, st));
    {
        unsigned long long nd1 = n_docs + 1;
        mark_docs_kernel<<<(unsigned)((nd1 + 255) / 256), 256, 0, st>>>(a.d_doc_off, n_docs, n_bytes, S.w_dbits.p, S.w_sfd.p,
                                                                       sparse_docs ? S.w_doc_tiles.p : nullptr, S.d_ctr);
        launches++;
    }
    // ---- special tokens (CoreBPE::encode, lib.rs:375-442; disallowed check, core.py:120-124) ------------------
    const uint32_t *hbits = S.w_dbits.p;         // haystack starts seen by the pre-tokeniser (= documents unless cut)
    const uint32_t *ibits = nullptr, *sbits = nullptr;
    if (a.sp_flags && !h->specials.empty() && !a.single_piece) {
        const size_t nsp = h->specials.size();
        CUDA_TRY(S.w_spflags.ensure(nsp + 16));
        CUDA_TRY(S.w_hbits.ensure((size_t)n_words + 8)); CUDA_TRY(S.w_ibits.ensure((size_t)n_words + 8));
        CUDA_TRY(S.w_sbits.ensure((size_t)n_words + 8)); CUDA_TRY(S.w_cbits.ensure((size_t)n_words + 8));
        CUDA_TRY(cudaMemcpyAsync(S.w_spflags.p, a.sp_flags, nsp, cudaMemcpyHostToDevice, st));   // tiny, from the caller's array
        CUDA_TRY(cudaMemsetAsync(S.w_ibits.p, 0, ((size_t)n_words + 8) * 4, st));
        CUDA_TRY(cudaMemsetAsync(S.w_cbits.p + n_words, 0, 8 * 4, st));
        const unsigned grid = (unsigned)((n_words + 255) / 256);
        special_mark_kernel<<<grid, 256, 0, st>>>(a.d_text, (long long)n_bytes, S.w_dbits.p, D->sp, S.w_spflags.p, S.w_cbits.p, n_words, S.d_ctr);
        CUDA_TRY(cudaMemcpyAsync(S.w_hbits.p, S.w_dbits.p, ((size_t)n_words + 8) * 4, cudaMemcpyDeviceToDevice, st));
        special_resolve_kernel<<<grid, 256, 0, st>>>(a.d_text, (long long)n_bytes, S.w_dbits.p, D->sp, S.w_spflags.p, S.w_cbits.p,
                                                     S.w_hbits.p, S.w_ibits.p, S.w_sbits.p, S.w_ltok.p, n_words, S.d_ctr);
        launches += 2;
        hbits = S.w_hbits.p; ibits = S.w_ibits.p; sbits = S.w_sbits.p;
    }
    CUDA_TRY(cudaEventRecord(S.ev[1], st));
    {
        unsigned grid = (unsigned)((n_words + 255) / 256);
        if (a.single_piece) single_piece_bits_kernel<<<grid, 256, 0, st>>>(S.w_pbits.p, S.w_psum.p, (long long)n_bytes, n_words);
        else {
            const uint32_t scap = (uint32_t)std::min<size_t>(S.slow_cap, 0xFFFFFFF0u);
#define B2_PRETOK(P)                                                                                                                   \
    pretok_kernel<P><<<grid, 256, 0, st>>>(a.d_text, (long long)n_bytes, hbits, D->uc, S.w_pbits.p, S.w_psum.p, n_words, ibits,          \
                                          S.w_slow.p, scap, S.d_ctr);                                                                  \
    pretok_slow_kernel<P><<<148 * 8, 256, 0, st>>>(a.d_text, (long long)n_bytes, hbits, D->uc, S.w_pbits.p, S.w_psum.p, S.w_slow.p, scap, S.d_ctr)
            if (h->pattern == PAT_R50K) { B2_PRETOK(PAT_R50K); }
            else if (h->pattern == PAT_CL100K) { B2_PRETOK(PAT_CL100K); }
            else { B2_PRETOK(PAT_O200K); }
#undef B2_PRETOK
            launches++;
        }
        launches++;
    }
    CUDA_TRY(cudaEventRecord(S.ev[2], st));
    {
        unsigned grid = (unsigned)((n_words + 255) / 256);
        find_long_kernel<<<grid, 256, 0, st>>>(S.w_pbits.p, S.w_psum.p, (long long)n_bytes, n_words, q, S.w_lidx.p, sbits, S.d_ctr);
        launches++;
        // The long-piece kernels only depend on find_long and nothing before the scan depends on them: they run on a
        // side stream next to probe + miss sort + miss (few SMs are busy with a giant piece, the rest probe).
        cudaStream_t ls = S.side, ls2 = S.side2;
        CUDA_TRY(cudaEventRecord(S.ev[10], st));
        CUDA_TRY(cudaStreamWaitEvent(ls, S.ev[10], 0));
        CUDA_TRY(cudaStreamWaitEvent(ls2, S.ev[10], 0));
        CUDA_TRY(cudaEventRecord(S.ev[11], ls));
        LongScratch LS{S.w_idA.p, S.w_rkA.p, S.w_idB.p, S.w_rkB.p, S.w_aux1.p, S.w_aux2.p, S.w_flag.p};
        if (h->pmerge) {         // 129..1024 bytes: segmented parallel merge (rounds, not merges, are sequential); shorter: a group of lanes per piece
            pmerge_kernel<256, 3><<<148 * 10, PM_WARPS * 32, 0, ls>>>(a.d_text, D->T, q, S.w_ltok.p, S.d_ctr);
            pmerge_long_kernel<<<148 * 4, PM_WARPS_L * 32, 0, ls>>>(a.d_text, D->T, q, S.w_ltok.p, S.d_ctr);
            if (h->pmerge_min_cls <= 2) pmerge_kernel<128, 2><<<148 * 16, PM_WARPS * 32, 0, ls>>>(a.d_text, D->T, q, S.w_ltok.p, S.d_ctr);
            else mid_group32_kernel<<<148 * 4, MIDG_WARPS * 32, 0, ls>>>(a.d_text, D->T, q, S.w_ltok.p, S.d_ctr, 2);
            if (h->pmerge_min_cls <= 1) pmerge_kernel<64, 1><<<148 * 16, PM_WARPS * 32, 0, ls>>>(a.d_text, D->T, q, S.w_ltok.p, S.d_ctr);
            mid_group16_kernel<<<148 * 9, MIDG_WARPS * 32, 0, ls>>>(a.d_text, D->T, q, S.w_ltok.p, S.d_ctr, h->pmerge_min_cls <= 1 ? 0 : 1);
        } else if (h->mid_group) {      // 17..1024 bytes: a group of lanes per piece, state in shared memory
            mid_group32_kernel<<<148 * 4, MIDG_WARPS * 32, 0, ls>>>(a.d_text, D->T, q, S.w_ltok.p, S.d_ctr, 4);
            mid_group16_kernel<<<148 * 9, MIDG_WARPS * 32, 0, ls>>>(a.d_text, D->T, q, S.w_ltok.p, S.d_ctr, 1);
        } else {                 // ranks of 2^22 and above: one piece per lane (72 KiB of columns per block) / warp per piece
            mid_thread_kernel<<<148 * 3, MID_WARPS * 32, MID_SMEM_BYTES, ls>>>(a.d_text, D->T, q, S.w_ltok.p, S.d_ctr);
            long_piece_kernel<<<148 * 8, LONG_WARPS * 32, 0, ls>>>(a.d_text, D->T, q, LS, S.w_ltok.p, S.d_ctr, CLS_G1024);
        }
        long_piece_kernel<<<148 * 8, LONG_WARPS * 32, 0, ls2>>>(a.d_text, D->T, q, LS, S.w_ltok.p, S.d_ctr, CLS_WARP);
        giant_piece_kernel<<<148, GIANT_THREADS, 0, ls2>>>(a.d_text, D->T, q, LS, S.w_ltok.p, S.d_ctr);
        cluster_piece_kernel<<<(148 / CLUSTER_CTAS) * CLUSTER_CTAS, GIANT_THREADS, 0, ls2>>>(a.d_text, D->T, q, LS, S.w_ltok.p, S.d_ctr);
        CUDA_TRY(cudaEventRecord(S.ev[13], ls2));
        CUDA_TRY(cudaEventRecord(S.ev[12], ls));
        launches += 5;
    }
    CUDA_TRY(cudaEventRecord(S.ev[3], st));
    {
        TileParams p;
        p.text = a.d_text; p.n_bytes = (long long)n_bytes; p.n_words = n_words; p.n_sub = n_tiles;
        p.pbits = S.w_pbits.p; p.dbits = S.w_dbits.p; p.span_first_doc = S.w_sfd.p;
        p.doc_off = a.d_doc_off; p.n_docs = n_docs; p.q = q; p.lidx = S.w_lidx.p; p.ltok = S.w_ltok.p;
        p.ptok = S.w_ptok.p; p.mres = S.w_mres.p; p.sbits = sbits;
        p.mq.key = S.w_mq_key.p; p.mq.pos = S.w_mq_pos.p; p.mq.roff = S.w_mq_roff.p; p.mq.len = S.w_mq_len.p; p.mq.rec = S.w_mq_rec.p;
        p.doc_tiles = S.w_doc_tiles.p;
        p.mq.skey = S.w_mq_skey.p; p.mq.smeta = S.w_mq_smeta.p; p.mq.cap = (uint32_t)std::min<size_t>(S.miss_cap, 0xFFFFFFF0u);
        p.mq.mres_cap = S.mres_cap;
        p.sub_count = S.w_sub_count.p; p.sub_base = S.w_sub_base.p;
        p.out = a.d_out; p.tok_off = a.d_tok_off; p.ctr = S.d_ctr;
        p.big_dst = S.w_big_dst.p; p.big_src = S.w_big_src.p; p.big_n = S.w_big_n.p;
        const long long want_blocks = (n_tiles + ENC_WARPS - 1) / ENC_WARPS;
        const unsigned probe_grid = (unsigned)std::min<long long>(want_blocks, 148ll * D->probe_blocks_per_sm);
        probe_kernel<<<probe_grid, ENC_WARPS * 32, 0, st>>>(p, D->T);
        CUDA_TRY(cudaEventRecord(S.ev[8], st));
        miss_hist_kernel<<<SORT_BLOCKS, 256, 0, st>>>(p, S.w_sort_hist.p);
        miss_base_kernel<<<1, 17 * 32, 0, st>>>(S.w_sort_hist.p, SORT_BLOCKS);
        miss_scatter_kernel<<<SORT_BLOCKS, 256, 0, st>>>(p, S.w_sort_hist.p);
        miss_kernel<<<148 * 16, MISS_WARPS * 32, 0, st>>>(p, D->T);
        CUDA_TRY(cudaEventRecord(S.ev[7], st));
        CUDA_TRY(cudaStreamWaitEvent(st, S.ev[12], 0));          // join: the long pieces' tokens and counts are needed from here on
        CUDA_TRY(cudaStreamWaitEvent(st, S.ev[13], 0));
        {
            const long long nb = (n_tiles + SCAN_ITEMS - 1) / SCAN_ITEMS;
            scan_partial_kernel<<<(unsigned)nb, 256, 0, st>>>(S.w_sub_count.p, n_tiles, S.w_scan_part.p);
            scan_top_kernel<<<1, 1024, 0, st>>>(S.w_scan_part.p, nb, S.d_ctr);
            scan_final_kernel<<<(unsigned)nb, 256, 0, st>>>(S.w_sub_count.p, n_tiles, S.w_scan_part.p, S.w_sub_base.p, S.d_ctr);
        }
        const unsigned gather_grid = (unsigned)((n_tiles + GATHER_WARPS - 1) / GATHER_WARPS);
        if (sparse_docs) {
            gather_kernel<0><<<gather_grid, GATHER_WARPS * 32, 0, st>>>(p);
            gather_kernel<1><<<(unsigned)((n_docs + 1 + GATHER_WARPS - 1) / GATHER_WARPS), GATHER_WARPS * 32, 0, st>>>(p);
        } else gather_kernel<2><<<gather_grid, GATHER_WARPS * 32, 0, st>>>(p);
        big_copy_kernel<<<148 * 2, 256, 0, st>>>(p);
        finalize_kernel<<<1, 32, 0, st>>>(S.d_ctr, S.d_sticky, a.d_counts, n_docs);
        launches += sparse_docs ? 12 : 11;
    }
    CUDA_TRY(cudaEventRecord(S.ev[4], st));
    CUDA_TRY(cudaMemcpyAsync(S.h_ctr, S.d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaEventRecord(S.ev[9], st));
    S.last_launches = launches;
    return B200BPE_OK;
}

static const int B200BPE_RETRY = 1;      // internal: capacities were grown, run the pipeline again

// Wait for an enqueued pipeline, collect stage timings and device-side flags.  Returns B200BPE_RETRY when a
// work-space that is sized from experience was too small for this batch (it has been grown to the exact need).
static int collect_pipeline(b200bpe *h, Slot &S, int *special_idx, uint64_t *special_pos) {
    (void)h;
    CUDA_TRY(cudaEventSynchronize(S.ev[9]));
    CUDA_TRY(cudaGetLastError());
    cudaEventElapsedTime(&S.last_ms[0], S.ev[0], S.ev[1]);
    cudaEventElapsedTime(&S.last_ms[1], S.ev[1], S.ev[2]);
    cudaEventElapsedTime(&S.last_ms[2], S.ev[11], S.ev[12]);     // long-piece kernels (side stream, overlapped with [3])
    cudaEventElapsedTime(&S.last_ms[3], S.ev[3], S.ev[7]);
    cudaEventElapsedTime(&S.last_ms[7], S.ev[7], S.ev[4]);
    cudaEventElapsedTime(&S.last_ms[8], S.ev[3], S.ev[8]);
    cudaEventElapsedTime(&S.last_ms[4], S.ev[0], S.ev[4]);
    const Counters &c = *S.h_ctr;
    if (c.err & ERR_DOCOFF)
        return fail(B200BPE_EINVAL, "document offsets must start at 0, be non-decreasing and end at n_bytes");
    if (c.err & ERR_SPECIAL) {
        const unsigned long long packed = ~c.special_pos;       // the scan keeps max(~(pos << 16 | index)) = the leftmost match
        if (special_idx) *special_idx = (int)(packed & 0xFFFFu);
        if (special_pos) *special_pos = packed >> 16;
        return fail(B200BPE_ESPECIAL, "text contains a disallowed special token");
    }
    if (c.err & (ERR_LONGCAP | ERR_MISSCAP | ERR_SLOWCAP)) {
        if (c.err & ERR_SLOWCAP) S.slow_cap = (size_t)c.n_slow + (size_t)(c.n_slow / 8) + 4096;
        if (c.err & ERR_LONGCAP) S.long_cap = (size_t)c.long_bytes + (size_t)(c.long_bytes / 8) + 4096;
        if (c.err & ERR_MISSCAP) {
            S.miss_cap = std::max(S.miss_cap, (size_t)c.n_miss + (size_t)(c.n_miss / 8) + 4096);
            S.mres_cap = std::max(S.mres_cap, (size_t)c.miss_bytes + (size_t)(c.miss_bytes / 8) + 4096);
        }
        return B200BPE_RETRY;
    }
    if (c.err & ERR_INTERNAL) return fail(B200BPE_ECUDA, "internal error: a merge kernel did not converge");
    if (c.err & ERR_NOBYTE)
        return fail(B200BPE_ENOBYTE, "a piece needs a single-byte token that mergeable_ranks does not contain");
    return B200BPE_OK;
}

static int device_args_check(b200bpe *h, const uint8_t *d_text, uint64_t n_bytes, const uint64_t *d_doc_off, uint32_t *d_tokens,
                             uint64_t *d_tok_off) {
    if (!h || !d_doc_off || !d_tokens || !d_tok_off || (n_bytes && !d_text)) return fail(B200BPE_EINVAL, "null argument");
    return B200BPE_OK;
}

static int device_enqueue_locked(b200bpe *h, const PendingDeviceCall &c) {
    DevCtx *D = h->devs[0];
    Slot &S = D->slots[0];
    cudaStream_t st = c.st ? c.st : S.stream;
    if (c.st) apply_l2_window(D, st);
    const uint8_t *txt = c.d_text;
    if (((uintptr_t)c.d_text & 15u) != 0) {                       // TMA staging and vector loads need 16-byte alignment
        CUDA_TRY(S.w_text.ensure((size_t)c.n_bytes + 64));
        CUDA_TRY(cudaMemcpyAsync(S.w_text.p, c.d_text, c.n_bytes, cudaMemcpyDeviceToDevice, st));
        txt = S.w_text.p;
    }
    PipeArgs a;
    a.d_text = txt; a.n_bytes = c.n_bytes; a.d_doc_off = c.d_doc_off; a.n_docs = c.n_docs; a.d_out = c.d_tokens;
    a.d_tok_off = c.d_tok_off; a.d_counts = c.d_counts; a.st = st;
    return enqueue_pipeline(h, D, S, a);
}

extern "C" int b200bpe_encode_device_async(b200bpe_t *h, const uint8_t *d_text, uint64_t n_bytes, const uint64_t *d_doc_off,
                                           uint64_t n_docs, uint32_t *d_tokens, uint64_t *d_tok_off, uint64_t *d_counts,
                                           void *stream) {
    int rc = device_args_check(h, d_text, n_bytes, d_doc_off, d_tokens, d_tok_off);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard;
    CUDA_TRY(cudaSetDevice(h->devs[0]->device));
    PendingDeviceCall c;
    c.active = true; c.queued = h->pending.active ? h->pending.queued + 1 : 1;
    c.d_text = d_text; c.n_bytes = n_bytes; c.d_doc_off = (const unsigned long long *)d_doc_off; c.n_docs = n_docs;
    c.d_tokens = d_tokens; c.d_tok_off = (unsigned long long *)d_tok_off; c.d_counts = (unsigned long long *)d_counts;
    c.st = (cudaStream_t)stream;
    rc = device_enqueue_locked(h, c);
    if (rc) return rc;
    h->pending = c;
    return B200BPE_OK;
}

extern "C" int b200bpe_device_wait(b200bpe_t *h, uint64_t *n_tokens) {
    if (!h) return fail(B200BPE_EINVAL, "null handle");
    std::lock_guard<std::mutex> lk(h->mu);
    if (!h->pending.active) return fail(B200BPE_EINVAL, "no device call in flight");
    DeviceGuard guard;
    DevCtx *D = h->devs[0];
    CUDA_TRY(cudaSetDevice(D->device));
    Slot &S = D->slots[0];
    PendingDeviceCall c = h->pending;
    h->pending.active = false;
    cudaStream_t st = c.st ? c.st : S.stream;
    int rc = collect_pipeline(h, S, nullptr, nullptr);
    const uint64_t total = S.h_ctr->total_tokens;
    // error bits of the earlier calls of a queued series (the counters only describe the last one)
    unsigned int sticky = 0;
    CUDA_TRY(cudaMemcpyAsync(&sticky, S.d_sticky, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemsetAsync(S.d_sticky, 0, sizeof(unsigned int), st));
    CUDA_TRY(cudaStreamSynchronize(st));
    uint64_t total2 = total;
    bool reran = false;
    for (int attempt = 0; rc == B200BPE_RETRY && attempt < 4; attempt++) {
        rc = device_enqueue_locked(h, c);
        if (rc) break;
        rc = collect_pipeline(h, S, nullptr, nullptr);
        total2 = S.h_ctr->total_tokens;
        CUDA_TRY(cudaMemsetAsync(S.d_sticky, 0, sizeof(unsigned int), st));
        reran = true;
    }
    memcpy(h->last_ms, S.last_ms, sizeof(h->last_ms));
    h->last_ms[5] = h->last_ms[6] = 0.f;
    h->last_launches = S.last_launches;
    if (rc == B200BPE_RETRY) return fail(B200BPE_ECUDA, "work-space sizing did not converge");
    if (rc) return rc;
    if (c.queued > 1) {
        if (reran || (sticky & (ERR_LONGCAP | ERR_MISSCAP | ERR_SLOWCAP)))
            return fail(B200BPE_ECAPACITY, "a work-space had to grow while several device calls were queued: only the last "
                                           "one was re-run, re-issue the others");
        if (sticky & ERR_NOBYTE) return fail(B200BPE_ENOBYTE, "a piece needs a single-byte token that mergeable_ranks does not contain");
        if (sticky & ERR_DOCOFF) return fail(B200BPE_EINVAL, "malformed document offsets in a queued device call");
    }
    if (n_tokens) *n_tokens = total2;
    return B200BPE_OK;
}

extern "C" int b200bpe_encode_device(b200bpe_t *h, const uint8_t *d_text, uint64_t n_bytes, const uint64_t *d_doc_off,
                                     uint64_t n_docs, uint32_t *d_tokens, uint64_t *d_tok_off, uint64_t *n_tokens,
                                     void *stream) {
    int rc = b200bpe_encode_device_async(h, d_text, n_bytes, d_doc_off, n_docs, d_tokens, d_tok_off, nullptr, stream);
    if (rc) return rc;
    return b200bpe_device_wait(h, n_tokens);
}

// --------------------------------------------------------------------------------------------
// host path: host buffers in, ONE pinned result out; chunks round-robin over the devices
// --------------------------------------------------------------------------------------------
struct HostJob {
    b200bpe *h = nullptr;
    const uint8_t *text = nullptr; const uint64_t *doc_off = nullptr; uint64_t n_docs = 0;
    bool single_piece = false, pageable = false;
    const uint8_t *sp_flags = nullptr;
    std::vector<uint64_t> cut;                    // chunk c = documents [cut[c], cut[c+1])
    std::vector<std::atomic<long long>> count;    // tokens of chunk c, -1 until its kernels are done
    b200bpe_result *r = nullptr; size_t tok_cap = 0;
    std::atomic<int> error{0}; std::atomic<bool> overflow{false};
    std::string error_msg; std::mutex err_mu;
    int special_idx = -1; uint64_t special_pos = 0;
    float sum_ms[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}; uint32_t launches = 0; std::mutex stat_mu;

    void set_error(int rc) {
        std::lock_guard<std::mutex> lk(err_mu);
        if (!error.load()) { error_msg = g_last_error; error.store(rc); }
    }
};

// the chunks of one device: c = first, first + step, ...
static void host_worker(HostJob *J, int dev_index, size_t first, size_t step) {
    b200bpe *h = J->h;
    DevCtx *D = h->devs[dev_index];
    const size_t n_chunks = J->cut.size() - 1;
    std::vector<size_t> mine;
    for (size_t c = first; c < n_chunks; c += step) mine.push_back(c);
    // ---- packed return path: a thread of its own waits for the packed tokens of a chunk to land in pinned staging and has
    //      the pool widen them into the result, so that neither the device pipeline nor this worker waits for host memory
    struct UnpackJob { Slot *S; cudaEvent_t ev; const uint8_t *src; uint32_t *dst; uint64_t n; };
    std::mutex uq_mu; std::condition_variable uq_cv; std::deque<UnpackJob> uq; bool uq_stop = false;
    const int pack_bits = (h->pack_bits && !J->single_piece) ? h->pack_bits : 0;
    std::thread unpacker;
    auto stop_unpacker = [&] {
        if (!unpacker.joinable()) return;
        { std::lock_guard<std::mutex> lk(uq_mu); uq_stop = true; }
        uq_cv.notify_all();
        unpacker.join();                                         // drains the queue first
    };
    auto bail = [&](int rc) {
        if (rc) J->set_error(rc);
        for (size_t k = 0; k < mine.size(); k++) { long long m1 = -1; J->count[mine[k]].compare_exchange_strong(m1, 0); }   // never leave a waiter spinning
        for (int i = 0; i < DevCtx::N_SLOTS; i++) { cudaStreamSynchronize(D->slots[i].up); cudaStreamSynchronize(D->slots[i].stream); }
        stop_unpacker();                                         // after the copies: everything queued is widened before we return
    };
    if (mine.empty()) return;
    if (cudaSetDevice(D->device) != cudaSuccess) { fail(B200BPE_ECUDA, "cudaSetDevice failed"); return bail(B200BPE_ECUDA); }
    if (pack_bits) unpacker = std::thread([&] {
        cudaSetDevice(D->device);
        for (;;) {
            UnpackJob job;
            {
                std::unique_lock<std::mutex> lk(uq_mu);
                uq_cv.wait(lk, [&] { return uq_stop || !uq.empty(); });
                if (uq.empty()) return;
                job = uq.front(); uq.pop_front();
            }
            if (cudaEventSynchronize(job.ev) == cudaSuccess)
                pool_for(h->pool, (size_t)job.n, (size_t)1 << 19, [&](size_t lo, size_t hi) { unpack_tokens(job.src, job.dst, lo, hi, pack_bits); });
            else J->set_error(fail(B200BPE_ECUDA, "cudaEventSynchronize failed in the token unpacker"));
            job.S->stage_out_busy.store(0);
        }
    });
    const uint64_t *doc_off = J->doc_off;
    auto slot_of = [&](size_t k) -> Slot & { return D->slots[k % DevCtx::N_SLOTS]; };

    // Upload of chunk k into its slot.  Pinned caller memory: on the slot's own stream, i.e. behind the download of the
    // slot's previous chunk -- measured better than letting uploads run free (free-running uploads and downloads fight for
    // the link: 38.8 vs 41.4 GB/s end to end).  Pageable caller memory: helper threads fill a pinned block quarter by
    // quarter and the quarters go up on the slot's UPLOAD stream, which only waits for the kernels (ev[4]) of the slot's
    // previous chunk, so that the host never sits behind a download (18 -> 22 GB/s).
    auto enqueue_h2d = [&](size_t k) -> int {
        Slot &S = slot_of(k);
        const size_t c = mine[k];
        const uint64_t lo = J->cut[c], hi = J->cut[c + 1], b0 = doc_off[lo], nb = doc_off[hi] - b0, nd = hi - lo;
        const bool staged = J->pageable && nb;
        cudaStream_t us = staged ? S.up : S.stream;
        if (!staged || S.w_text.cap < (size_t)nb + 64 || S.w_docoff.cap < (size_t)nd + 2 || S.w_tokoff.cap < (size_t)nd + 2 ||
            S.w_out.cap < (size_t)nb + 64)
            CUDA_TRY(cudaStreamSynchronize(S.stream));           // slot free again (or a buffer has to grow: nothing may be in flight)
        else if (k >= (size_t)DevCtx::N_SLOTS) CUDA_TRY(cudaStreamWaitEvent(S.up, S.ev[4], 0));
        CUDA_TRY(S.w_text.ensure((size_t)nb + 64)); CUDA_TRY(S.w_docoff.ensure((size_t)nd + 2));
        CUDA_TRY(S.w_tokoff.ensure((size_t)nd + 2)); CUDA_TRY(S.w_out.ensure((size_t)nb + 64));
        const uint8_t *src = J->text + b0;
        CUDA_TRY(cudaEventRecord(S.ev[5], us));
        if (staged) {
            CUDA_TRY(cudaStreamSynchronize(S.up));               // the block's previous upload has left it
            if (S.stage.cap < nb) {
                if (S.stage.p) cudaFreeHost(S.stage.p);
                S.stage.p = nullptr; S.stage.cap = 0;
                const size_t want = (size_t)nb + (size_t)(nb / 8) + 4096;
                CUDA_TRY(cudaHostAlloc(&S.stage.p, want, cudaHostAllocPortable));
                S.stage.cap = want;
            }
            uint8_t *stg = (uint8_t *)S.stage.p;
            const size_t group = std::max<size_t>((((size_t)nb / 4) + 4095) & ~(size_t)4095, (size_t)4 << 20);
            for (size_t g0 = 0; g0 < (size_t)nb; g0 += group) {
                const size_t len = std::min(group, (size_t)nb - g0);
                pool_for(h->pool, len, (size_t)1 << 20, [&](size_t lo, size_t hi) { memcpy(stg + g0 + lo, src + g0 + lo, hi - lo); });
                CUDA_TRY(cudaMemcpyAsync(S.w_text.p + g0, stg + g0, len, cudaMemcpyHostToDevice, us));
            }
        } else if (nb) CUDA_TRY(cudaMemcpyAsync(S.w_text.p, src, nb, cudaMemcpyHostToDevice, us));
        CUDA_TRY(cudaMemcpyAsync(S.w_docoff.p, doc_off + lo, (nd + 1) * 8, cudaMemcpyHostToDevice, us));
        if (b0) add_offset_kernel<<<(unsigned)((nd + 1 + 255) / 256), 256, 0, us>>>(S.w_docoff.p, nd + 1, -(long long)b0);
        CUDA_TRY(cudaEventRecord(S.ev[6], us));
        if (staged) CUDA_TRY(cudaStreamWaitEvent(S.stream, S.ev[6], 0));     // the slot's pipeline starts when its text has arrived
        return B200BPE_OK;
    };
    auto args_of = [&](size_t k) {
        Slot &S = slot_of(k);
        const size_t c = mine[k];
        const uint64_t lo = J->cut[c], hi = J->cut[c + 1];
        PipeArgs a;
        a.d_text = S.w_text.p; a.n_bytes = doc_off[hi] - doc_off[lo]; a.d_doc_off = S.w_docoff.p; a.n_docs = hi - lo;
        a.d_out = S.w_out.p; a.d_tok_off = S.w_tokoff.p; a.st = S.stream; a.single_piece = J->single_piece; a.sp_flags = J->sp_flags;
        return a;
    };
    size_t known = 0; uint64_t known_sum = 0;                    // prefix of the per-chunk token counts seen so far
    float sum_ms[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}; uint32_t launches = 0; float last_d2h = 0;
    // finalise chunk k: wait for its kernels, then send its offsets + tokens home (async) at their final place
    auto drain = [&](size_t k) -> int {
        Slot &S = slot_of(k);
        const size_t c = mine[k];
        const uint64_t lo = J->cut[c], hi = J->cut[c + 1], nd = hi - lo;
        int sidx = -1; uint64_t spos = 0;
        int rc = collect_pipeline(h, S, &sidx, &spos);
        for (int attempt = 0; rc == B200BPE_RETRY && attempt < 4; attempt++) {   // a work-space grew: same chunk again
            rc = enqueue_pipeline(h, D, S, args_of(k));
            if (!rc) rc = collect_pipeline(h, S, &sidx, &spos);
        }
        if (rc == B200BPE_RETRY) rc = fail(B200BPE_ECUDA, "work-space sizing did not converge");
        if (rc == B200BPE_ESPECIAL) {
            std::lock_guard<std::mutex> lk(J->err_mu);
            const uint64_t gpos = spos + doc_off[lo];
            if (J->special_idx < 0 || gpos < J->special_pos) { J->special_idx = sidx; J->special_pos = gpos; }
        }
        if (rc) return rc;
        float h2d = 0; cudaEventElapsedTime(&h2d, S.ev[5], S.ev[6]);
        const uint64_t nt = S.h_ctr->total_tokens;
        J->count[c].store((long long)nt);
        while (known < c) {                                      // token base = counts of all earlier chunks (other devices)
            long long v = J->count[known].load();
            if (v < 0) {
                if (J->error.load() || J->overflow.load()) return B200BPE_OK;
                std::this_thread::yield();
                continue;
            }
            known_sum += (uint64_t)v; known++;
        }
        const uint64_t token_base = known_sum;
        if (token_base + nt > J->tok_cap) { J->overflow.store(true); return B200BPE_OK; }
        if (token_base) add_offset_kernel<<<(unsigned)((nd + 1 + 255) / 256), 256, 0, S.stream>>>(S.w_tokoff.p, nd + 1, (long long)token_base);
        CUDA_TRY(cudaEventRecord(S.ev[5], S.stream));
        CUDA_TRY(cudaMemcpyAsync((uint64_t *)J->r->off.p + lo, S.w_tokoff.p, (nd + 1) * 8, cudaMemcpyDeviceToHost, S.stream));
        if (nt && pack_bits) {
            const unsigned long long n_words = (nt * (uint64_t)pack_bits + 31) / 32;
            const size_t bytes = (size_t)n_words * 4;
            CUDA_TRY(S.w_pack.ensure((size_t)n_words + 4));
            while (S.stage_out_busy.load()) std::this_thread::yield();        // the slot's previous chunk is still being widened
            if (S.stage_out.cap < bytes + 16) {
                if (S.stage_out.p) cudaFreeHost(S.stage_out.p);
                S.stage_out.p = nullptr; S.stage_out.cap = 0;
                const size_t want = bytes + bytes / 4 + 4096;
                CUDA_TRY(cudaHostAlloc(&S.stage_out.p, want, cudaHostAllocPortable));
                S.stage_out.cap = want;
            }
            pack_tokens_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, S.stream>>>(S.w_out.p, nt, pack_bits, S.w_pack.p, n_words);
            CUDA_TRY(cudaMemcpyAsync(S.stage_out.p, S.w_pack.p, bytes, cudaMemcpyDeviceToHost, S.stream));
            CUDA_TRY(cudaEventRecord(S.ev[14], S.stream));
            S.stage_out_busy.store(1);
            { std::lock_guard<std::mutex> lk(uq_mu); uq.push_back({&S, S.ev[14], (const uint8_t *)S.stage_out.p, (uint32_t *)J->r->tok.p + token_base, nt}); }
            uq_cv.notify_one();
        } else if (nt) CUDA_TRY(cudaMemcpyAsync((uint32_t *)J->r->tok.p + token_base, S.w_out.p, nt * 4, cudaMemcpyDeviceToHost, S.stream));
        CUDA_TRY(cudaEventRecord(S.ev[6], S.stream));
        for (int i = 0; i < 5; i++) sum_ms[i] += S.last_ms[i];
        sum_ms[7] += S.last_ms[7]; sum_ms[8] += S.last_ms[8]; sum_ms[5] += h2d; launches += S.last_launches;
        return B200BPE_OK;
    };
    auto stop = [&]() { return J->error.load() != 0 || J->overflow.load(); };
    // uploads run AHEAD chunks in front of the kernels; the slot of chunk k + AHEAD last served chunk k + AHEAD - N_SLOTS
    // <= k - 2, which was drained (its download issued) in an earlier iteration
    const size_t AHEAD = (size_t)DevCtx::N_SLOTS - 2;
    int rc = B200BPE_OK;
    for (size_t k = 0; k < AHEAD && k < mine.size(); k++) { rc = enqueue_h2d(k); if (rc) return bail(rc); }
    for (size_t k = 0; k < mine.size() && !stop(); k++) {
        if (k + AHEAD < mine.size()) { rc = enqueue_h2d(k + AHEAD); if (rc) return bail(rc); }
        rc = enqueue_pipeline(h, D, slot_of(k), args_of(k));
        if (rc) return bail(rc);
        if (k >= 1) { rc = drain(k - 1); if (rc) return bail(rc); }   // overlaps with chunk k's kernels
    }
    if (!stop()) { rc = drain(mine.size() - 1); if (rc) return bail(rc); }
    bail(0);
    if (!stop()) { Slot &S = slot_of(mine.size() - 1); cudaEventElapsedTime(&last_d2h, S.ev[5], S.ev[6]); }
    std::lock_guard<std::mutex> lk(J->stat_mu);
    for (int i = 0; i < 9; i++) J->sum_ms[i] = std::max(J->sum_ms[i], sum_ms[i]);       // per device; report the slowest
    J->sum_ms[6] = std::max(J->sum_ms[6], last_d2h);
    J->launches += launches;
}

static int encode_host(b200bpe *h, const uint8_t *text, const uint64_t *doc_off, uint64_t n_docs, bool single_piece,
                       const uint8_t *sp_flags, b200bpe_result **out, int *special_idx) {
    const uint64_t n_bytes = doc_off[n_docs];
    if (doc_off[0] != 0) return fail(B200BPE_EINVAL, "document offsets must start at 0");
    const int n_dev = (int)h->devs.size();
    if (!h->pool) h->pool = new TaskPool();
    h->pool->ensure(h->copy_threads);
    // ---- chunk plan: document ranges [lo, hi) ------------------------------------------------
    size_t chunk = h->chunk_bytes;
    if (!h->chunk_forced && n_dev > 1) chunk = std::min<size_t>(64u << 20, std::max<size_t>(8u << 20, (size_t)(n_bytes / (4 * (uint64_t)n_dev))));
    std::vector<uint64_t> cut; cut.push_back(0);
    if (!single_piece && n_bytes > chunk + chunk / 2) {
        uint64_t lo = 0;
        while (lo < n_docs) {
            const uint64_t want = doc_off[lo] + chunk;
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
    bool pageable = false;
    if (n_bytes) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, text) != cudaSuccess) { cudaGetLastError(); pageable = true; }
        else pageable = (at.type == cudaMemoryTypeUnregistered);
    }
    // tokens <= bytes; bytes/2 covers everything but pathological text, which gets a second pass with the worst case
    for (int pass = 0; pass < 2; pass++) {
        HostJob J;
        J.h = h; J.text = text; J.doc_off = doc_off; J.n_docs = n_docs; J.single_piece = single_piece; J.pageable = pageable;
        J.sp_flags = sp_flags; J.cut = cut;
        J.count = std::vector<std::atomic<long long>>(n_chunks);
        for (auto &c : J.count) c.store(-1);
        b200bpe_result *r = new b200bpe_result();
        r->owner = h; r->n_docs = n_docs;
        r->off = h->take_pinned((size_t)(n_docs + 1) * 8);
        const size_t want_tok = pass == 0 ? (size_t)(n_bytes / 2) + 4096 : (size_t)n_bytes + 4096;
        r->tok = h->take_pinned(want_tok * 4);
        auto cleanup = [&](int rc) { h->give_pinned(r->tok); h->give_pinned(r->off); delete r; return rc; };
        if (!r->off.p || !r->tok.p) return cleanup(fail(B200BPE_ECUDA, "pinned allocation failed"));
        J.r = r; J.tok_cap = r->tok.cap / 4;
        const int workers = (int)std::min<size_t>((size_t)n_dev, n_chunks);
        if (workers <= 1) host_worker(&J, 0, 0, 1);
        else {
            std::vector<std::thread> th;
            for (int d = 0; d < workers; d++) th.emplace_back(host_worker, &J, d, (size_t)d, (size_t)workers);
            for (auto &t : th) t.join();
        }
        if (J.error.load()) {
            g_last_error = J.error_msg;
            if (J.error.load() == B200BPE_ESPECIAL && special_idx) *special_idx = J.special_idx;
            return cleanup(J.error.load());
        }
        if (J.overflow.load()) { cleanup(0); continue; }
        uint64_t total = 0;
        for (auto &c : J.count) total += (uint64_t)std::max<long long>(0, c.load());
        memcpy(h->last_ms, J.sum_ms, sizeof(J.sum_ms)); h->last_launches = J.launches;
        r->n_tokens = total;
        h->live_results++;                                        // caller holds h->mu
        *out = r;
        return B200BPE_OK;
    }
    return fail(B200BPE_ECUDA, "token buffer sizing did not converge");
}

extern "C" int b200bpe_encode_ordinary_batch(b200bpe_t *h, const uint8_t *text, const uint64_t *doc_off,
                                             uint64_t n_docs, b200bpe_result_t **out) {
    if (!h || !doc_off || !out) return fail(B200BPE_EINVAL, "null argument");
    if (doc_off[n_docs] && !text) return fail(B200BPE_EINVAL, "null text");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard;
    return encode_host(h, text, doc_off, n_docs, false, nullptr, out, nullptr);
}

extern "C" int b200bpe_encode_single_piece(b200bpe_t *h, const uint8_t *piece, uint64_t len, b200bpe_result_t **out) {
    if (!h || !out || (len && !piece)) return fail(B200BPE_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard;
    uint64_t off[2] = {0, len};
    return encode_host(h, piece, off, 1, true, nullptr, out, nullptr);
}

// CoreBPE::encode (src/lib.rs:375-442) + the disallowed-special check of Encoding.encode (tiktoken/core.py:120-124),
// both on the device: one multi-pattern scan finds every occurrence of a special token; a disallowed one is an
// error (the leftmost is reported), allowed ones cut their document into haystacks for the pre-tokeniser and are
// emitted as their own ids.  flags[i]: 1 = allowed, 2 = disallowed, 0 = ordinary text.
extern "C" int b200bpe_encode_batch_special(b200bpe_t *h, const uint8_t *text, const uint64_t *doc_off, uint64_t n_docs,
                                            const uint8_t *flags, b200bpe_result_t **out, int32_t *special_index) {
    if (!h || !doc_off || !out) return fail(B200BPE_EINVAL, "null argument");
    if (doc_off[n_docs] && !text) return fail(B200BPE_EINVAL, "null text");
    bool any = false;
    if (flags) for (size_t i = 0; i < h->specials.size(); i++) any |= flags[i] != 0;
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard;
    int sidx = -1;
    int rc = encode_host(h, text, doc_off, n_docs, false, any ? flags : nullptr, out, &sidx);
    if (rc == B200BPE_ESPECIAL && special_index) *special_index = sidx;
    return rc;
}

extern "C" int b200bpe_encode_batch(b200bpe_t *h, const uint8_t *text, const uint64_t *doc_off, uint64_t n_docs,
                                    const uint8_t *allowed, b200bpe_result_t **out) {
    if (!h) return fail(B200BPE_EINVAL, "null argument");
    std::vector<uint8_t> flags(h->specials.size() + 1, 0);
    if (allowed) for (size_t i = 0; i < h->specials.size(); i++) flags[i] = allowed[i] ? 1 : 0;
    return b200bpe_encode_batch_special(h, text, doc_off, n_docs, allowed ? flags.data() : nullptr, out, nullptr);
}

extern "C" const char *b200bpe_special_name(b200bpe_t *h, int32_t index) {
    if (!h || index < 0 || (size_t)index >= h->specials.size()) return nullptr;
    return h->specials[(size_t)index].c_str();
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

// --------------------------------------------------------------------------------------------
// decode
// --------------------------------------------------------------------------------------------
static const std::string *decode_one(b200bpe *h, uint32_t t) {
    auto it = h->H.decoder.find(t);
    if (it != h->H.decoder.end()) return &it->second;
    auto it2 = h->special_decoder.find(t);
    if (it2 != h->special_decoder.end()) return &it2->second;
    return nullptr;
}

extern "C" int b200bpe_decode_bytes(b200bpe_t *h, const uint32_t *tokens, uint64_t n_tokens, uint8_t *out,
                                    uint64_t out_cap, uint64_t *out_len, uint32_t *bad_token) {
    if (!h || (n_tokens && !tokens) || !out_len) return fail(B200BPE_EINVAL, "null argument");
    uint64_t k = 0;
    for (uint64_t i = 0; i < n_tokens; i++) {
        const std::string *s = decode_one(h, tokens[i]);
        if (!s) {
            if (bad_token) *bad_token = tokens[i];
            return fail(B200BPE_EKEY, "Invalid token for decoding: " + std::to_string(tokens[i]));
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
    if (tok_off[0] != 0) return fail(B200BPE_EINVAL, "token offsets must start at 0");
    for (uint64_t d = 0; d < n_docs; d++)
        if (tok_off[d + 1] < tok_off[d]) return fail(B200BPE_EINVAL, "token offsets must be non-decreasing");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard;
    struct ResultGuard {               // no leak on the error paths below
        b200bpe *h; b200bpe_result *r = nullptr; bool keep = false;
        ~ResultGuard() { if (r && !keep) { h->give_pinned(r->tok); h->give_pinned(r->off); delete r; } }
    } rg{h};
    if (!h->decode_on_device) {        // ids of 2^24 and above exist: host maps (same results)
        b200bpe_result *r = new b200bpe_result();
        rg.r = r;
        r->owner = h; r->n_docs = n_docs;
        uint64_t total = 0;
        for (uint64_t i = 0; i < n; i++) {
            const std::string *s = decode_one(h, tokens[i]);
            if (!s) { if (bad_token) *bad_token = tokens[i]; return fail(B200BPE_EKEY, "Invalid token for decoding: " + std::to_string(tokens[i])); }
            total += s->size();
        }
        r->off = h->take_pinned((size_t)(n_docs + 1) * 8);
        r->tok = h->take_pinned((size_t)total + 16);
        if (!r->off.p || !r->tok.p) return fail(B200BPE_ECUDA, "pinned allocation failed");
        uint8_t *dst = (uint8_t *)r->tok.p; uint64_t k = 0;
        for (uint64_t d = 0; d < n_docs; d++) {
            ((uint64_t *)r->off.p)[d] = k;
            for (uint64_t i = tok_off[d]; i < tok_off[d + 1]; i++) { const std::string *s = decode_one(h, tokens[i]); memcpy(dst + k, s->data(), s->size()); k += s->size(); }
        }
        ((uint64_t *)r->off.p)[n_docs] = k;
        r->n_tokens = k; rg.keep = true; h->live_results++;
        memset(h->last_ms, 0, sizeof(h->last_ms)); h->last_launches = 0;
        *out = r;
        return B200BPE_OK;
    }
    DevCtx *D = h->devs[0];
    CUDA_TRY(cudaSetDevice(D->device));
    Slot &S = D->slots[0];
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
    if (n) decode_len_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(S.w_out.p, n, D->d_tok_boff, h->n_ids, S.w_sub_count.p, S.d_ctr);
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
    if (n) decode_copy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(S.w_out.p, n, D->d_tok_boff, h->n_ids, D->d_tok_blob,
                                                                        S.w_sub_base.p, S.w_text.p);
    decode_doc_off_kernel<<<(unsigned)((n_docs + 1 + 255) / 256), 256, 0, st>>>(S.w_docoff.p, n_docs, S.w_sub_base.p, S.w_tokoff.p);
    CUDA_TRY(cudaEventRecord(S.ev[2], st));
    b200bpe_result *r = new b200bpe_result();
    rg.r = r;
    r->owner = h; r->n_docs = n_docs; r->n_tokens = n_out;
    r->off = h->take_pinned((size_t)(n_docs + 1) * 8);
    r->tok = h->take_pinned((size_t)n_out + 16);
    if (!r->off.p || !r->tok.p) return fail(B200BPE_ECUDA, "pinned allocation failed");
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
    rg.keep = true;
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
