// tools/bpe_train.cpp -- a small, fast BPE trainer used ONLY to synthesise vocabularies
// (the public tiktoken vocabulary files are not available offline).
// Development tooling, not product: the output is a mergeable_ranks table in which
// rank = merge order, with the 256 single bytes at ranks 0..255, i.e. the same shape as a
// .tiktoken file (tiktoken/load.py:147-171).
//
// Classic word-frequency BPE: count distinct pieces, keep pair counts and pair -> word
// occurrence lists, pop the most frequent pair from a lazily-invalidated heap.
#include <cstdint>
#include <cstring>
#include <queue>
#include <string>
#include <unordered_map>
#include <vector>
#include <algorithm>

namespace {
struct Word { std::vector<int32_t> sym; uint64_t count; };
struct HeapItem {
    uint64_t count; uint64_t key;
    bool operator<(const HeapItem &o) const { return count != o.count ? count < o.count : key > o.key; }
};
inline uint64_t pkey(int32_t a, int32_t b) { return (uint64_t(uint32_t(a)) << 32) | uint32_t(b); }
}

extern "C" int64_t bpe_train(const uint8_t *text, const uint64_t *starts, const uint64_t *ends,
                             uint64_t n_pieces, uint32_t target_vocab, uint64_t min_count,
                             uint8_t *out_bytes, uint64_t out_cap, uint64_t *out_off /* target_vocab+1 */) {
    std::unordered_map<std::string, uint64_t> wc;
    wc.reserve(1 << 20);
    for (uint64_t i = 0; i < n_pieces; i++)
        wc[std::string((const char *)text + starts[i], ends[i] - starts[i])]++;
    std::vector<std::pair<std::string, uint64_t>> sorted(wc.begin(), wc.end());
    std::sort(sorted.begin(), sorted.end());          // deterministic order
    std::vector<Word> words(sorted.size());
    for (size_t w = 0; w < sorted.size(); w++) {
        words[w].count = sorted[w].second;
        for (unsigned char c : sorted[w].first) words[w].sym.push_back(c);
    }
    std::vector<std::string> tok(256);
    for (int i = 0; i < 256; i++) tok[i] = std::string(1, char(i));
    std::unordered_map<std::string, int32_t> known;
    for (int i = 0; i < 256; i++) known[tok[i]] = i;

    std::unordered_map<uint64_t, uint64_t> pc;                      // pair -> count
    std::unordered_map<uint64_t, std::vector<uint32_t>> occ;        // pair -> words (may be stale)
    pc.reserve(1 << 20); occ.reserve(1 << 20);
    for (uint32_t w = 0; w < words.size(); w++) {
        auto &s = words[w].sym;
        for (size_t i = 0; i + 1 < s.size(); i++) {
            uint64_t k = pkey(s[i], s[i + 1]);
            pc[k] += words[w].count;
            auto &v = occ[k];
            if (v.empty() || v.back() != w) v.push_back(w);
        }
    }
    std::priority_queue<HeapItem> heap;
    for (auto &kv : pc) heap.push({kv.second, kv.first});

    while (tok.size() < target_vocab && !heap.empty()) {
        HeapItem top = heap.top(); heap.pop();
        auto it = pc.find(top.key);
        if (it == pc.end() || it->second != top.count) continue;    // stale
        if (top.count < min_count) break;
        int32_t a = int32_t(top.key >> 32), b = int32_t(top.key & 0xFFFFFFFFu);
        std::string nb = tok[a] + tok[b];
        if (known.count(nb)) { pc.erase(it); continue; }            // same bytes already a token
        int32_t id = int32_t(tok.size());
        tok.push_back(nb); known[nb] = id;
        std::vector<uint32_t> ws; ws.swap(occ[top.key]);
        pc.erase(top.key);
        std::unordered_map<uint64_t, int64_t> delta;
        for (uint32_t w : ws) {
            auto &s = words[w].sym; uint64_t c = words[w].count;
            std::vector<int32_t> ns; ns.reserve(s.size());
            bool any = false;
            for (size_t i = 0; i < s.size();) {
                if (i + 1 < s.size() && s[i] == a && s[i + 1] == b) { ns.push_back(id); i += 2; any = true; }
                else ns.push_back(s[i++]);
            }
            if (!any) continue;
            for (size_t i = 0; i + 1 < s.size(); i++) delta[pkey(s[i], s[i + 1])] -= int64_t(c);
            for (size_t i = 0; i + 1 < ns.size(); i++) {
                uint64_t k = pkey(ns[i], ns[i + 1]);
                delta[k] += int64_t(c);
                if (ns[i] == id || ns[i + 1] == id) {
                    auto &v = occ[k];
                    if (v.empty() || v.back() != w) v.push_back(w);
                }
            }
            s.swap(ns);
        }
        for (auto &d : delta) {
            if (d.second == 0 || d.first == top.key) continue;
            uint64_t &c = pc[d.first];
            c = uint64_t(int64_t(c) + d.second);
            if (c == 0) pc.erase(d.first); else heap.push({c, d.first});
        }
    }
    uint64_t o = 0;
    for (size_t i = 0; i < tok.size(); i++) {
        out_off[i] = o;
        if (o + tok[i].size() > out_cap) return -1;
        memcpy(out_bytes + o, tok[i].data(), tok[i].size());
        o += tok[i].size();
    }
    out_off[tok.size()] = o;
    return int64_t(tok.size());
}
