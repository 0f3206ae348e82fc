/*
 * oracle/bpe_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded-per-document CPU restatement of tiktoken's encode hot path,
 * used as the checker for the CUDA engine (tests/, __graft_entry__.smoke(), and the
 * cpu_baseline / --impl reference leg of bench.py).  Nothing under tiktoken_b200/ may
 * import, link or call this file: the product path has no CPU fallback.
 *
 * What is restated (reference = openai/tiktoken v0.14.0, /root/reference):
 *   - byte_pair_encode / _byte_pair_merge / _byte_pair_merge_large   src/lib.rs:47-211
 *   - CoreBPE::encode_ordinary                                        src/lib.rs:360-373
 *   - CoreBPE::encode (special-token slicing)                         src/lib.rs:375-442
 *   - the regex pre-tokeniser: the reference delegates to the third-party crates
 *     fancy-regex 0.19 / regex 1.13 (Cargo.toml:23-24, not vendored).  Their published
 *     semantics (Perl-style leftmost-first alternation, greedy quantifiers with
 *     backtracking, possessive quantifiers, negative look-ahead, `$` = end of haystack,
 *     Unicode classes, (?i) simple case folding) are restated here LITERALLY for the three
 *     pat_strs of tiktoken_ext/openai_public.py:12-14 (r50k family), :89 (cl100k),
 *     :104-114 (o200k): every alternative is tried in order at each match start and
 *     backtracking loops are written out, no "simplified" closed forms.
 *
 * Pinning: see oracle/README.md.  The restatement is checked against the reference's own
 * Rust unit tests (src/lib.rs:689-701), and -- because no vocabulary file exists offline --
 * against outputs of the real Rust engine (the tiktoken 0.12.0 wheel installed in this
 * image) on exhaustive class-strings and on synthetic vocabularies; the fixtures and the
 * generating script live in tests/golden/.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "unicode_classes.inc"

enum { C_O = 0, C_LU = 1, C_LL = 2, C_LB = 3, C_M = 4, C_N = 5, C_SP = 6, C_WS = 7, C_NL = 8,
       C_APOS = 9, C_SLASH = 10 };

enum { PAT_R50K = 0, PAT_CL100K = 1, PAT_O200K = 2 };

#define RANK_MAX 0xFFFFFFFFu

/* ------------------------------------------------------------------------------------ */
/* byte-string -> rank map (stands in for FxHashMap<Vec<u8>, Rank>, src/lib.rs:321)       */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    const uint8_t *key;
    uint32_t len;
    uint32_t rank;
} slot_t;

typedef struct {
    slot_t *slots;
    uint64_t mask;
    uint8_t *blob;        /* owned copy of all token bytes */
    uint64_t *off;        /* n+1 offsets into blob */
    uint32_t *rank;       /* rank per entry */
    uint32_t n;
    int pattern;
    /* special tokens */
    uint8_t *sp_blob;
    uint64_t *sp_off;
    uint32_t *sp_rank;
    uint32_t n_sp;
} orc_t;

static uint64_t fnv1a(const uint8_t *p, uint64_t n) {
    uint64_t h = 1469598103934665603ull;
    for (uint64_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; }
    return h ^ (h >> 29);
}

static uint32_t map_get(const orc_t *o, const uint8_t *p, uint64_t n) {
    uint64_t i = fnv1a(p, n) & o->mask;
    for (;;) {
        const slot_t *s = &o->slots[i];
        if (!s->key) return RANK_MAX;
        if (s->len == n && memcmp(s->key, p, n) == 0) return s->rank;
        i = (i + 1) & o->mask;
    }
}

/* ------------------------------------------------------------------------------------ */
/* _byte_pair_merge -- src/lib.rs:140-196 (linear scan, pieces < 100 bytes)               */
/* returns number of parts entries written (boundaries incl. the two sentinels)          */
/* ------------------------------------------------------------------------------------ */
typedef struct { uint64_t start; uint32_t rank; } part_t;

static uint64_t merge_small(const orc_t *o, const uint8_t *piece, uint64_t len, part_t *parts) {
    uint64_t np = 0;
    uint32_t min_rank = RANK_MAX; uint64_t min_i = (uint64_t)-1;
    for (uint64_t i = 0; i + 1 < len; i++) {                   /* lib.rs:149-155 */
        uint32_t r = map_get(o, piece + i, 2);
        if (r < min_rank) { min_rank = r; min_i = i; }
        parts[np].start = i; parts[np].rank = r; np++;
    }
    parts[np].start = len - 1; parts[np].rank = RANK_MAX; np++; /* lib.rs:156 */
    parts[np].start = len;     parts[np].rank = RANK_MAX; np++; /* lib.rs:157 */

    while (min_rank != RANK_MAX) {                              /* lib.rs:178 */
        uint64_t i = min_i;
        /* get_rank(parts, j): lib.rs:159-172 */
        if (i > 0) {
            uint64_t j = i - 1;
            parts[j].rank = (j + 3 < np)
                ? map_get(o, piece + parts[j].start, parts[j + 3].start - parts[j].start) : RANK_MAX;
        }
        parts[i].rank = (i + 3 < np)
            ? map_get(o, piece + parts[i].start, parts[i + 3].start - parts[i].start) : RANK_MAX;
        memmove(&parts[i + 1], &parts[i + 2], (np - i - 2) * sizeof(part_t)); /* parts.remove(i+1) */
        np--;
        min_rank = RANK_MAX; min_i = (uint64_t)-1;
        for (uint64_t k = 0; k + 1 < np; k++)                    /* lib.rs:188-193 */
            if (parts[k].rank < min_rank) { min_rank = parts[k].rank; min_i = k; }
    }
    return np;
}

/* ------------------------------------------------------------------------------------ */
/* _byte_pair_merge_large -- src/lib.rs:47-138 (heap + linked state, pieces >= 100)       */
/* ------------------------------------------------------------------------------------ */
typedef struct { uint64_t prev, end, next_end; uint32_t next_rank, cur_rank; } state_t;
typedef struct { uint64_t start; uint32_t rank; } hmerge_t;

/* BinaryHeap<Merge> with Ord = (smaller rank, then smaller start) is "greater": lib.rs:23-31 */
static int hm_before(hmerge_t a, hmerge_t b) {
    if (a.rank != b.rank) return a.rank < b.rank;
    return a.start < b.start;
}
typedef struct { hmerge_t *a; uint64_t n, cap; } heap_t;
static void heap_push(heap_t *h, hmerge_t m) {
    if (h->n == h->cap) { h->cap = h->cap ? h->cap * 2 : 64; h->a = realloc(h->a, h->cap * sizeof(hmerge_t)); }
    uint64_t i = h->n++;
    while (i > 0) {
        uint64_t p = (i - 1) / 2;
        if (!hm_before(m, h->a[p])) break;
        h->a[i] = h->a[p]; i = p;
    }
    h->a[i] = m;
}
static hmerge_t heap_pop(heap_t *h) {
    hmerge_t top = h->a[0];
    hmerge_t m = h->a[--h->n];
    uint64_t i = 0;
    for (;;) {
        uint64_t c = 2 * i + 1;
        if (c >= h->n) break;
        if (c + 1 < h->n && hm_before(h->a[c + 1], h->a[c])) c++;
        if (!hm_before(h->a[c], m)) break;
        h->a[i] = h->a[c]; i = c;
    }
    if (h->n) h->a[i] = m;
    return top;
}

static void potential_merge(const orc_t *o, const uint8_t *piece, uint64_t len, state_t *st,
                            heap_t *heap, uint64_t start, uint64_t next_end_item) {
    st[start].next_end = next_end_item;                          /* lib.rs:85-93 */
    st[start].next_rank = RANK_MAX;
    if (next_end_item <= len) {
        uint32_t r = map_get(o, piece + start, next_end_item - start);
        if (r != RANK_MAX) {
            hmerge_t m = { start, r };
            heap_push(heap, m);
            st[start].next_rank = r;
        }
    }
}

static uint64_t merge_large(const orc_t *o, const uint8_t *piece, uint64_t len, uint32_t *out) {
    state_t *st = malloc(len * sizeof(state_t));
    heap_t heap = {0};
    st[0].prev = (uint64_t)-1; st[0].end = 1; st[0].next_end = 2;
    st[0].next_rank = RANK_MAX; st[0].cur_rank = RANK_MAX;
    for (uint64_t i = 0; i + 1 < len; i++) {                     /* lib.rs:58-71 */
        uint32_t r = map_get(o, piece + i, 2);
        if (r != RANK_MAX) { hmerge_t m = { i, r }; heap_push(&heap, m); st[i].next_rank = r; }
        st[i + 1].prev = i; st[i + 1].end = i + 2; st[i + 1].next_end = i + 3;
        st[i + 1].next_rank = RANK_MAX; st[i + 1].cur_rank = RANK_MAX;
    }
    while (heap.n) {                                             /* lib.rs:97-125 */
        hmerge_t left = heap_pop(&heap);
        if (left.rank == RANK_MAX) break;
        if (left.rank != st[left.start].next_rank) continue;
        uint64_t left_start = left.start;
        uint64_t right_start = st[left_start].end;
        uint64_t right_end = st[left_start].next_end;
        uint64_t right_next_end = st[right_start].next_end;
        st[left_start].cur_rank = st[left_start].next_rank;
        st[left_start].end = right_end;
        potential_merge(o, piece, len, st, &heap, left_start, right_next_end);
        if (right_end < len) st[right_end].prev = left_start;
        if (left_start > 0) {
            uint64_t prev_start = st[left_start].prev;
            potential_merge(o, piece, len, st, &heap, prev_start, right_end);
        }
        st[right_start].next_rank = RANK_MAX;
    }
    uint64_t k = 0, i = 0;
    while (i < len) {                                            /* lib.rs:127-136 */
        if (st[i].cur_rank != RANK_MAX) out[k++] = st[i].cur_rank;
        else out[k++] = map_get(o, piece + i, st[i].end - i);    /* ranks[...] panics if absent */
        i = st[i].end;
    }
    free(st); free(heap.a);
    return k;
}

/* byte_pair_encode -- src/lib.rs:198-211.  `force` selects an algorithm for cross-checks:
 * 0 = reference dispatch, 1 = always linear, 2 = always heap.  A RANK_MAX in the output
 * marks the place where the reference would panic (`ranks[...]` on a missing key). */
static uint64_t byte_pair_encode(const orc_t *o, const uint8_t *piece, uint64_t len, uint32_t *out,
                                 int force) {
    if (len == 1) { out[0] = map_get(o, piece, 1); return 1; }
    if ((force == 0 && len < 100) || force == 1) {
        part_t *parts = malloc((len + 2) * sizeof(part_t));
        uint64_t np = merge_small(o, piece, len, parts);
        uint64_t k = 0;
        for (uint64_t i = 0; i + 1 < np; i++)                    /* windows(2), lib.rs:206-208 */
            out[k++] = map_get(o, piece + parts[i].start, parts[i + 1].start - parts[i].start);
        free(parts);
        return k;
    }
    return merge_large(o, piece, len, out);
}

/* whole-piece probe then byte_pair_encode: src/lib.rs:367-370 (== encode_single_piece, py.rs:145-150) */
static uint64_t encode_piece(const orc_t *o, const uint8_t *piece, uint64_t len, uint32_t *out, int force) {
    uint32_t r = map_get(o, piece, len);
    if (r != RANK_MAX) { out[0] = r; return 1; }
    return byte_pair_encode(o, piece, len, out, force);
}

/* ------------------------------------------------------------------------------------ */
/* UTF-8 decode of one haystack into code points + classes                               */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    uint32_t *cp; uint8_t *cls; uint64_t *boff; /* byte offset of each scalar, +1 sentinel */
    int64_t n;
} hay_t;

static uint8_t class_of(uint32_t cp) {
    if (cp >= 0x110000) return C_O;
    return UC_STAGE2[(uint32_t)UC_STAGE1[cp >> 8] * 256 + (cp & 255)];
}

static void hay_build(hay_t *h, const uint8_t *s, uint64_t len) {
    h->cp = malloc((len + 1) * sizeof(uint32_t));
    h->cls = malloc(len + 1);
    h->boff = malloc((len + 1) * sizeof(uint64_t));
    int64_t n = 0; uint64_t i = 0;
    while (i < len) {
        uint8_t b = s[i]; uint32_t cp; int k;
        if (b < 0x80) { cp = b; k = 1; }
        else if (b < 0xE0) { cp = b & 0x1F; k = 2; }
        else if (b < 0xF0) { cp = b & 0x0F; k = 3; }
        else { cp = b & 0x07; k = 4; }
        for (int j = 1; j < k && i + j < len; j++) cp = (cp << 6) | (s[i + j] & 0x3F);
        h->cp[n] = cp; h->cls[n] = class_of(cp); h->boff[n] = i; n++;
        i += k;
    }
    h->boff[n] = len; h->n = n;
}
static void hay_free(hay_t *h) { free(h->cp); free(h->cls); free(h->boff); }

static int isL(uint8_t c) { return c == C_LU || c == C_LL || c == C_LB; }
static int isN(uint8_t c) { return c == C_N; }
static int isS(uint8_t c) { return c == C_SP || c == C_WS || c == C_NL; }
static int isOther(uint8_t c) { return !isS(c) && !isL(c) && !isN(c); }     /* [^\s\p{L}\p{N}] */
static int isPrefix(uint8_t c) { return c != C_NL && !isL(c) && !isN(c); }  /* [^\r\n\p{L}\p{N}] */
static int isU200(uint8_t c) { return c == C_LU || c == C_LB || c == C_M; } /* [\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}] */
static int isL200(uint8_t c) { return c == C_LL || c == C_LB || c == C_M; } /* [\p{Ll}\p{Lm}\p{Lo}\p{M}] */

/* (?i:x) for an ASCII letter x under Unicode simple case folding; the only non-ASCII member
 * for the letters sdmtlver is U+017F -> s (swept from the reference engine, see
 * tools/gen_unicode_tables.py). */
static int ci_eq(uint32_t cp, char x) {
    if (cp == (uint32_t)x || cp == (uint32_t)(x - 32)) return 1;
    if (x == 's' && cp == 0x17F) return 1;
    return 0;
}

/* `'(?:[sdmt]|ll|ve|re)` at p (r50k: case-sensitive; cl100k: (?i:...)).  Returns length in
 * scalars including the apostrophe, or 0.  Alternatives tried in the written order. */
static int64_t contraction_at(const hay_t *h, int64_t p, int ci) {
    if (p >= h->n || h->cp[p] != '\'') return 0;
    if (p + 1 < h->n) {
        uint32_t a = h->cp[p + 1];
        const char *one = "sdmt";
        for (int i = 0; i < 4; i++) if (ci ? ci_eq(a, one[i]) : a == (uint32_t)one[i]) return 2;
    }
    if (p + 2 < h->n) {
        uint32_t a = h->cp[p + 1], b = h->cp[p + 2];
        const char *two[3] = { "ll", "ve", "re" };
        for (int i = 0; i < 3; i++) {
            int ok = ci ? (ci_eq(a, two[i][0]) && ci_eq(b, two[i][1]))
                        : (a == (uint32_t)two[i][0] && b == (uint32_t)two[i][1]);
            if (ok) return 3;
        }
    }
    return 0;
}

/* o200k suffix `(?i:'s|'t|'re|'ve|'m|'ll|'d)` in the written order */
static int64_t contraction200_at(const hay_t *h, int64_t p) {
    if (p >= h->n || h->cp[p] != '\'') return 0;
    uint32_t a = p + 1 < h->n ? h->cp[p + 1] : 0xFFFFFFFF;
    uint32_t b = p + 2 < h->n ? h->cp[p + 2] : 0xFFFFFFFF;
    if (ci_eq(a, 's')) return 2;
    if (ci_eq(a, 't')) return 2;
    if (ci_eq(a, 'r') && ci_eq(b, 'e')) return 3;
    if (ci_eq(a, 'v') && ci_eq(b, 'e')) return 3;
    if (ci_eq(a, 'm')) return 2;
    if (ci_eq(a, 'l') && ci_eq(b, 'l')) return 3;
    if (ci_eq(a, 'd')) return 2;
    return 0;
}

/* `\s+(?!\S)`: greedy, backtrack one scalar at a time until the look-ahead holds */
static int64_t ws_not_before_nonspace(const hay_t *h, int64_t p) {
    int64_t q = p;
    while (q < h->n && isS(h->cls[q])) q++;
    for (int64_t e = q; e > p; e--)
        if (!(e < h->n && !isS(h->cls[e]))) return e;
    return -1;
}

/* ---- r50k family: openai_public.py:12-14 ------------------------------------------- */
static int64_t match_r50k(const hay_t *h, int64_t p) {
    const uint8_t *c = h->cls; int64_t n = h->n, q;
    /* 1: '(?:[sdmt]|ll|ve|re) */
    q = contraction_at(h, p, 0);
    if (q) return p + q;
    /* 2..4: ` ?X++` with X = \p{L}, \p{N}, [^\s\p{L}\p{N}] */
    for (int alt = 0; alt < 3; alt++) {
        for (int sp = 1; sp >= 0; sp--) {             /* greedy ` ?`: with the space first */
            if (sp && h->cp[p] != ' ') continue;
            q = p + sp;
            int64_t s = q;
            while (q < n && (alt == 0 ? isL(c[q]) : alt == 1 ? isN(c[q]) : isOther(c[q]))) q++;
            if (q > s) return q;
        }
    }
    /* 5: \s++$ */
    q = p; while (q < n && isS(c[q])) q++;
    if (q > p && q == n) return q;
    /* 6: \s+(?!\S) */
    q = ws_not_before_nonspace(h, p);
    if (q > 0) return q;
    /* 7: \s */
    if (isS(c[p])) return p + 1;
    return -1;
}

/* ---- cl100k: openai_public.py:89 ----------------------------------------------------- */
static int64_t match_cl100k(const hay_t *h, int64_t p) {
    const uint8_t *c = h->cls; int64_t n = h->n, q;
    /* 1: '(?i:[sdmt]|ll|ve|re) */
    q = contraction_at(h, p, 1);
    if (q) return p + q;
    /* 2: [^\r\n\p{L}\p{N}]?+\p{L}++  (possessive optional: taken whenever it matches) */
    q = p;
    if (isPrefix(c[q])) q++;
    { int64_t s = q; while (q < n && isL(c[q])) q++; if (q > s) return q; }
    /* 3: \p{N}{1,3}+ */
    q = p; while (q < n && q < p + 3 && isN(c[q])) q++;
    if (q > p) return q;
    /* 4:  ?[^\s\p{L}\p{N}]++[\r\n]*+ */
    for (int sp = 1; sp >= 0; sp--) {
        if (sp && h->cp[p] != ' ') continue;
        q = p + sp;
        int64_t s = q;
        while (q < n && isOther(c[q])) q++;
        if (q > s) { while (q < n && c[q] == C_NL) q++; return q; }
    }
    /* 5: \s++$ */
    q = p; while (q < n && isS(c[q])) q++;
    if (q > p && q == n) return q;
    /* 6: \s*[\r\n]  (greedy \s*, give back until a CR/LF follows) */
    q = p; while (q < n && isS(c[q])) q++;
    for (int64_t e = q; e >= p; e--)
        if (e < n && c[e] == C_NL) return e + 1;
    /* 7: \s+(?!\S) */
    q = ws_not_before_nonspace(h, p);
    if (q > 0) return q;
    /* 8: \s */
    if (isS(c[p])) return p + 1;
    return -1;
}

/* ---- o200k: openai_public.py:104-114 ------------------------------------------------- */
static int64_t match_o200k(const hay_t *h, int64_t p) {
    const uint8_t *c = h->cls; int64_t n = h->n, q;
    /* 1: [^\r\n\p{L}\p{N}]?[U]*[L]+(?i:'s|'t|'re|'ve|'m|'ll|'d)?   -- full backtracking */
    for (int pre = 1; pre >= 0; pre--) {
        if (pre && !isPrefix(c[p])) continue;
        int64_t s = p + pre, umax = s;
        while (umax < n && isU200(c[umax])) umax++;
        for (int64_t u = umax; u >= s; u--) {                 /* U* greedy, give back */
            int64_t l = u;
            while (l < n && isL200(c[l])) l++;               /* L+ greedy */
            if (l > u) return l + contraction200_at(h, l);    /* optional suffix, greedy */
        }
    }
    /* 2: [^\r\n\p{L}\p{N}]?[U]+[L]*(?i:...)? */
    for (int pre = 1; pre >= 0; pre--) {
        if (pre && !isPrefix(c[p])) continue;
        int64_t s = p + pre, u = s;
        while (u < n && isU200(c[u])) u++;
        if (u > s) {
            int64_t l = u;
            while (l < n && isL200(c[l])) l++;
            return l + contraction200_at(h, l);
        }
    }
    /* 3: \p{N}{1,3} */
    q = p; while (q < n && q < p + 3 && isN(c[q])) q++;
    if (q > p) return q;
    /* 4:  ?[^\s\p{L}\p{N}]+[\r\n/]* */
    for (int sp = 1; sp >= 0; sp--) {
        if (sp && h->cp[p] != ' ') continue;
        q = p + sp;
        int64_t s = q;
        while (q < n && isOther(c[q])) q++;
        if (q > s) { while (q < n && (c[q] == C_NL || c[q] == C_SLASH)) q++; return q; }
    }
    /* 5: \s*[\r\n]+ */
    q = p; while (q < n && isS(c[q])) q++;
    for (int64_t e = q; e >= p; e--)
        if (e < n && c[e] == C_NL) { int64_t t = e; while (t < n && c[t] == C_NL) t++; return t; }
    /* 6: \s+(?!\S) */
    q = ws_not_before_nonspace(h, p);
    if (q > 0) return q;
    /* 7: \s+ */
    q = p; while (q < n && isS(c[q])) q++;
    if (q > p) return q;
    return -1;
}

static int64_t match_at(const orc_t *o, const hay_t *h, int64_t p) {
    switch (o->pattern) {
        case PAT_R50K: return match_r50k(h, p);
        case PAT_CL100K: return match_cl100k(h, p);
        default: return match_o200k(h, p);
    }
}

/* ------------------------------------------------------------------------------------ */
/* public C entry points (ctypes)                                                        */
/* ------------------------------------------------------------------------------------ */
orc_t *orc_new(const uint8_t *tok_bytes, const uint64_t *tok_off, const uint32_t *tok_rank, uint32_t n,
               const uint8_t *sp_bytes, const uint64_t *sp_off, const uint32_t *sp_rank, uint32_t n_sp,
               int pattern) {
    orc_t *o = calloc(1, sizeof(orc_t));
    o->n = n; o->pattern = pattern;
    uint64_t total = tok_off[n];
    o->blob = malloc(total ? total : 1); memcpy(o->blob, tok_bytes, total);
    o->off = malloc((n + 1) * sizeof(uint64_t)); memcpy(o->off, tok_off, (n + 1) * sizeof(uint64_t));
    o->rank = malloc((n ? n : 1) * sizeof(uint32_t)); memcpy(o->rank, tok_rank, n * sizeof(uint32_t));
    uint64_t cap = 16; while (cap < 2ull * n + 2) cap <<= 1;
    o->slots = calloc(cap, sizeof(slot_t)); o->mask = cap - 1;
    for (uint32_t i = 0; i < n; i++) {
        const uint8_t *k = o->blob + o->off[i]; uint64_t len = o->off[i + 1] - o->off[i];
        uint64_t s = fnv1a(k, len) & o->mask;
        while (o->slots[s].key) s = (s + 1) & o->mask;
        o->slots[s].key = k; o->slots[s].len = (uint32_t)len; o->slots[s].rank = tok_rank[i];
    }
    o->n_sp = n_sp;
    if (n_sp) {
        uint64_t st = sp_off[n_sp];
        o->sp_blob = malloc(st ? st : 1); memcpy(o->sp_blob, sp_bytes, st);
        o->sp_off = malloc((n_sp + 1) * sizeof(uint64_t)); memcpy(o->sp_off, sp_off, (n_sp + 1) * sizeof(uint64_t));
        o->sp_rank = malloc(n_sp * sizeof(uint32_t)); memcpy(o->sp_rank, sp_rank, n_sp * sizeof(uint32_t));
    }
    return o;
}

void orc_free(orc_t *o) {
    if (!o) return;
    free(o->slots); free(o->blob); free(o->off); free(o->rank);
    free(o->sp_blob); free(o->sp_off); free(o->sp_rank); free(o);
}

/* regex.find_iter(text): piece END byte offsets (pieces are contiguous unless text is
 * skipped; a skipped span is reported as start of next piece via out_starts). */
int64_t orc_split(const orc_t *o, const uint8_t *text, uint64_t len, uint64_t *out_starts,
                  uint64_t *out_ends, uint64_t cap) {
    hay_t h; hay_build(&h, text, len);
    int64_t p = 0, k = 0;
    while (p < h.n) {
        int64_t e = match_at(o, &h, p);
        if (e < 0) { p++; continue; }                 /* find_iter skips unmatched text */
        if ((uint64_t)k < cap) { out_starts[k] = h.boff[p]; out_ends[k] = h.boff[e]; }
        k++; p = e;
    }
    hay_free(&h);
    return k;
}

/* CoreBPE::encode_ordinary on one haystack: src/lib.rs:360-373 */
static int64_t encode_ordinary_into(const orc_t *o, const uint8_t *text, uint64_t len, uint32_t *out) {
    hay_t h; hay_build(&h, text, len);
    int64_t p = 0, k = 0;
    while (p < h.n) {
        int64_t e = match_at(o, &h, p);
        if (e < 0) { p++; continue; }
        k += (int64_t)encode_piece(o, text + h.boff[p], h.boff[e] - h.boff[p], out + k, 0);
        p = e;
    }
    hay_free(&h);
    return k;
}

int64_t orc_encode_ordinary(const orc_t *o, const uint8_t *text, uint64_t len, uint32_t *out) {
    return encode_ordinary_into(o, text, len, out);    /* out must hold len entries */
}

int64_t orc_encode_piece(const orc_t *o, const uint8_t *piece, uint64_t len, uint32_t *out, int force) {
    if (len == 0) return 0;
    return (int64_t)encode_piece(o, piece, len, out, force);
}

/* byte_pair_split (src/lib.rs:213-219): boundaries of _byte_pair_merge, for the Rust unit tests */
int64_t orc_byte_pair_split(const orc_t *o, const uint8_t *piece, uint64_t len, uint64_t *bounds) {
    part_t *parts = malloc((len + 2) * sizeof(part_t));
    uint64_t np = merge_small(o, piece, len, parts);
    for (uint64_t i = 0; i < np; i++) bounds[i] = parts[i].start;
    free(parts);
    return (int64_t)np;
}

/* CoreBPE::encode: src/lib.rs:375-442.  allowed[i] != 0 marks special i as allowed.  The
 * reference's special regex is an alternation in HashMap iteration order (unspecified);
 * at a given position we take the longest special, which only differs when one special is
 * a prefix of another (never the case in openai_public.py). */
int64_t orc_encode(const orc_t *o, const uint8_t *text, uint64_t len, const uint8_t *allowed, uint32_t *out) {
    uint64_t start = 0; int64_t k = 0;
    for (;;) {
        uint64_t best_pos = len; int64_t best = -1; uint64_t best_len = 0;
        uint64_t start_find = start;
        for (;;) {                                               /* lib.rs:389-401 */
            best = -1; best_pos = len; best_len = 0;
            for (uint64_t pos = start_find; pos < len && best < 0; pos++) {
                for (uint32_t s = 0; s < o->n_sp; s++) {
                    uint64_t sl = o->sp_off[s + 1] - o->sp_off[s];
                    if (sl && pos + sl <= len && memcmp(text + pos, o->sp_blob + o->sp_off[s], sl) == 0
                        && sl > best_len) { best = s; best_len = sl; best_pos = pos; }
                }
            }
            if (best < 0) break;
            if (allowed && allowed[best]) break;
            start_find = best_pos + 1;
            /* stay on a scalar boundary like find_from_pos over &str would */
            while (start_find < len && (text[start_find] & 0xC0) == 0x80) start_find++;
        }
        uint64_t end = best >= 0 ? best_pos : len;
        k += encode_ordinary_into(o, text + start, end - start, out + k);   /* lib.rs:405-424 */
        if (best < 0) break;
        out[k++] = o->sp_rank[best];                             /* lib.rs:426-436 */
        start = best_pos + best_len;
    }
    return k;
}

/* ------------------------------------------------------------------------------------ */
/* batch driver with host threads -- the "port" CPU baseline of bench.py                 */
/* (stands in for ThreadPoolExecutor over per-document calls, tiktoken/core.py:164-176)  */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    const orc_t *o; const uint8_t *text; const uint64_t *doc_off; uint64_t n_docs;
    uint32_t *scratch;   /* token scratch, same indexing as text bytes */
    uint64_t *counts;    /* tokens per doc */
    uint64_t next; pthread_mutex_t mu;
} batch_t;

static void *batch_worker(void *arg) {
    batch_t *b = arg;
    for (;;) {
        pthread_mutex_lock(&b->mu);
        uint64_t lo = b->next; b->next += 16; pthread_mutex_unlock(&b->mu);
        if (lo >= b->n_docs) break;
        uint64_t hi = lo + 16 < b->n_docs ? lo + 16 : b->n_docs;
        for (uint64_t d = lo; d < hi; d++) {
            uint64_t s = b->doc_off[d], e = b->doc_off[d + 1];
            b->counts[d] = (uint64_t)encode_ordinary_into(b->o, b->text + s, e - s, b->scratch + s);
        }
    }
    return NULL;
}

/* tokens_out must hold doc_off[n_docs] entries; tok_off_out n_docs+1 entries. */
int64_t orc_encode_ordinary_batch(const orc_t *o, const uint8_t *text, const uint64_t *doc_off,
                                  uint64_t n_docs, int n_threads, uint32_t *tokens_out,
                                  uint64_t *tok_off_out) {
    batch_t b; memset(&b, 0, sizeof b);
    b.o = o; b.text = text; b.doc_off = doc_off; b.n_docs = n_docs;
    uint64_t total = doc_off[n_docs];
    b.scratch = malloc((total ? total : 1) * sizeof(uint32_t));
    b.counts = calloc(n_docs ? n_docs : 1, sizeof(uint64_t));
    pthread_mutex_init(&b.mu, NULL);
    if (n_threads < 1) n_threads = 1;
    pthread_t *th = malloc(n_threads * sizeof(pthread_t));
    for (int i = 0; i < n_threads; i++) pthread_create(&th[i], NULL, batch_worker, &b);
    for (int i = 0; i < n_threads; i++) pthread_join(th[i], NULL);
    uint64_t k = 0;
    for (uint64_t d = 0; d < n_docs; d++) {
        tok_off_out[d] = k;
        memcpy(tokens_out + k, b.scratch + doc_off[d], b.counts[d] * sizeof(uint32_t));
        k += b.counts[d];
    }
    tok_off_out[n_docs] = k;
    free(th); free(b.scratch); free(b.counts); pthread_mutex_destroy(&b.mu);
    return (int64_t)k;
}
