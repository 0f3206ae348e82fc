// bpe_tables.h -- host-side construction of the engine's lookup tables from mergeable_ranks.
// Replaces CoreBPE::new_internal's map building (src/lib.rs:618-663): runs once per Encoding.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "bpe_device.cuh"

namespace b2bpe {

struct HostTables {
    std::vector<uint32_t> byte_id;    // 256
    std::vector<uint32_t> pair2;      // 65536
    std::vector<U4> pair_tab;  uint32_t pair_mask = 0;
    std::vector<U4> piece_tab; uint32_t piece_mask = 0;
    std::vector<U4> long_tab;  uint32_t long_mask = 0;
    std::vector<uint8_t> long_blob;
    uint32_t max_token_len = 0, n_long_tokens = 0, max_rank = 0;
    uint64_t n_pairs = 0;
    // decode side (host): rank -> bytes
    std::unordered_map<uint32_t, std::string> decoder;
    std::string error;

    DevTables view() const {
        DevTables T;
        T.byte_id = byte_id.data(); T.pair2 = pair2.data();
        T.pair_tab = pair_tab.data(); T.pair_mask = pair_mask;
        T.piece_tab = piece_tab.data(); T.piece_mask = piece_mask;
        T.long_tab = long_tab.data(); T.long_mask = long_mask;
        T.long_blob = long_blob.data();
        T.max_token_len = max_token_len; T.n_long_tokens = n_long_tokens;
        return T;
    }
};

inline uint32_t pow2_at_least(uint64_t n) { uint32_t c = 16; while (c < n) c <<= 1; return c; }

inline void pack16(const uint8_t *p, uint32_t len, uint64_t &k0, uint64_t &k1) {
    uint8_t b[16] = {0};
    memcpy(b, p, len);
    memcpy(&k0, b, 8); memcpy(&k1, b + 8, 8);       // little-endian hosts only (x86-64, aarch64)
}

inline uint64_t long_hash_bytes(const uint8_t *p, uint32_t len) {
    uint64_t h = long_hash_init(len);
    for (uint32_t i = 0; i < len; i += 8) {
        uint8_t b[8] = {0};
        memcpy(b, p + i, len - i < 8 ? len - i : 8);
        uint64_t w; memcpy(&w, b, 8);
        h = long_hash_step(h, w, i / 8);
    }
    return h;
}

// returns 0 or a negative B200BPE_E* code (values mirrored from include/b200bpe.h)
inline int build_tables(const uint8_t *tok_bytes, const uint64_t *tok_off, const uint32_t *tok_rank,
                        uint32_t n, HostTables &H, uint32_t pair_slack = 3) {
    std::unordered_map<std::string, uint32_t> enc;
    enc.reserve((size_t)n * 2 + 16);
    H.decoder.reserve((size_t)n * 2 + 16);
    for (uint32_t i = 0; i < n; i++) {
        uint64_t len = tok_off[i + 1] - tok_off[i];
        if (len == 0) { H.error = "empty token in mergeable_ranks"; return -1; }
        if (tok_rank[i] >= (1u << 30)) { H.error = "rank too large (token ids must be < 2^30)"; return -1; }
        std::string s((const char *)tok_bytes + tok_off[i], (size_t)len);
        if (!enc.emplace(s, tok_rank[i]).second) { H.error = "duplicate token bytes"; return -1; }
        if (!H.decoder.emplace(tok_rank[i], s).second) {
            // reference: assert!(encoder.len() == decoder.len(), ...) src/lib.rs:636-641
            H.error = "Encoder and decoder must be of equal length. Maybe you had duplicate token indices in your encoder?";
            return -3;
        }
        if (len > H.max_token_len) H.max_token_len = (uint32_t)len;
        if (tok_rank[i] > H.max_rank) H.max_rank = tok_rank[i];
    }
    H.byte_id.assign(256, 0);
    for (int b = 0; b < 256; b++) {
        auto it = enc.find(std::string(1, (char)b));
        H.byte_id[b] = it != enc.end() ? it->second : PSEUDO_BASE + (uint32_t)b;
    }
    H.pair2.assign(65536, RANK_MAX);
    auto id_of = [&](const std::string &s, uint32_t &id) -> bool {
        auto it = enc.find(s);
        if (it != enc.end()) { id = it->second; return true; }
        if (s.size() == 1) { id = PSEUDO_BASE + (uint8_t)s[0]; return true; }
        return false;
    };
    struct Pair { uint32_t a, b, r; };
    std::vector<Pair> pairs;
    uint32_t n_short = 0, n_long = 0;
    for (auto &kv : enc) {
        const std::string &t = kv.first;
        if (t.size() <= (size_t)SHORT_MAX) n_short++; else n_long++;
        if (t.size() == 2) H.pair2[((uint8_t)t[0] << 8) | (uint8_t)t[1]] = kv.second;
        for (size_t k = 1; k < t.size(); k++) {
            uint32_t a, b;
            if (id_of(t.substr(0, k), a) && id_of(t.substr(k), b)) pairs.push_back({a, b, kv.second});
        }
    }
    H.n_pairs = pairs.size();
    // buckets of two slots; capacity >= 3x the entries
    uint32_t nbuckets = pow2_at_least(((uint64_t)pairs.size() * pair_slack + 2) / 2 + 1);
    H.pair_mask = nbuckets - 1;
    H.pair_tab.assign((size_t)nbuckets * 2, U4{0xFFFFFFFFu, 0xFFFFFFFFu, RANK_MAX, 0});
    for (auto &p : pairs) {
        uint32_t s = pair_hash(p.a, p.b) & H.pair_mask;
        for (;;) {
            if (H.pair_tab[2 * s].x == 0xFFFFFFFFu) { H.pair_tab[2 * s] = U4{p.a, p.b, p.r, 0}; break; }
            if (H.pair_tab[2 * s + 1].x == 0xFFFFFFFFu) { H.pair_tab[2 * s + 1] = U4{p.a, p.b, p.r, 0}; break; }
            s = (s + 1) & H.pair_mask;
        }
    }
    uint32_t sc = pow2_at_least((uint64_t)n_short * 3 + 2);
    H.piece_mask = sc - 1;
    H.piece_tab.assign((size_t)sc * 2, U4{0, 0, 0, 0});
    uint32_t lc = pow2_at_least((uint64_t)n_long * 2 + 2);
    H.long_mask = lc - 1;
    H.long_tab.assign((size_t)lc * 2, U4{0, 0, 0, 0});
    H.n_long_tokens = n_long;
    for (auto &kv : enc) {
        const std::string &t = kv.first;
        uint32_t len = (uint32_t)t.size();
        if (len <= (uint32_t)SHORT_MAX) {
            uint64_t k0, k1; pack16((const uint8_t *)t.data(), len, k0, k1);
            uint32_t s = (uint32_t)piece_hash(k0, k1, len) & H.piece_mask;
            while (H.piece_tab[2 * s + 1].x != 0) s = (s + 1) & H.piece_mask;
            H.piece_tab[2 * s] = U4{(uint32_t)k0, (uint32_t)(k0 >> 32), (uint32_t)k1, (uint32_t)(k1 >> 32)};
            H.piece_tab[2 * s + 1] = U4{len, kv.second, 0, 0};
        } else {
            uint64_t h = long_hash_bytes((const uint8_t *)t.data(), len);
            uint32_t s = (uint32_t)(h ^ (h >> 32)) & H.long_mask;
            while (H.long_tab[2 * s].w != 0) s = (s + 1) & H.long_mask;
            uint32_t off = (uint32_t)H.long_blob.size();
            H.long_blob.insert(H.long_blob.end(), t.begin(), t.end());
            H.long_tab[2 * s] = U4{(uint32_t)h, (uint32_t)(h >> 32), off, len};
            H.long_tab[2 * s + 1] = U4{kv.second, 0, 0, 0};
        }
    }
    if (H.long_blob.empty()) H.long_blob.push_back(0);
    return 0;
}

}  // namespace b2bpe
