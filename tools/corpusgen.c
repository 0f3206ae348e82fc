/*
 * tools/corpusgen.c -- seeded synthetic corpora for the parity tests and bench.py
 * (SURVEY.md section 8(d) workloads).  Development / measurement tooling, not product.
 *
 * All generators write EXACTLY nbytes of valid UTF-8 into out (the tail is padded with ASCII
 * spaces when the next unit does not fit).  Deterministic for a given (kind, seed, nbytes).
 *
 *   kind 0  english   words Zipf-drawn from a syllable-built lexicon, punctuation, numbers,
 *                     contractions, capitalisation, paragraphs, a few out-of-lexicon words
 *   kind 1  mixed     spans of english / CJK / kana+hangul / emoji / whitespace runs / digits+punct
 *   kind 2  code      indented code-like lines plus long-piece stressors (runs, blobs)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef struct { uint64_t s; } rng_t;
static uint64_t rnd(rng_t *r) {             /* splitmix64 */
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static uint32_t rbelow(rng_t *r, uint32_t n) { return (uint32_t)((rnd(r) >> 32) * (uint64_t)n >> 32); }
static double runit(rng_t *r) { return (double)(rnd(r) >> 11) * (1.0 / 9007199254740992.0); }

typedef struct { uint8_t *p; uint64_t n, cap; } buf_t;
static int put(buf_t *b, const void *s, uint64_t k) {
    if (b->n + k > b->cap) return 0;
    memcpy(b->p + b->n, s, k); b->n += k; return 1;
}
static int putc_(buf_t *b, char c) { return put(b, &c, 1); }
static int put_utf8(buf_t *b, uint32_t cp) {
    uint8_t t[4]; int k;
    if (cp < 0x80) { t[0] = (uint8_t)cp; k = 1; }
    else if (cp < 0x800) { t[0] = 0xC0 | (cp >> 6); t[1] = 0x80 | (cp & 63); k = 2; }
    else if (cp < 0x10000) { t[0] = 0xE0 | (cp >> 12); t[1] = 0x80 | ((cp >> 6) & 63); t[2] = 0x80 | (cp & 63); k = 3; }
    else { t[0] = 0xF0 | (cp >> 18); t[1] = 0x80 | ((cp >> 12) & 63); t[2] = 0x80 | ((cp >> 6) & 63); t[3] = 0x80 | (cp & 63); k = 4; }
    return put(b, t, k);
}

/* ---- lexicon ----------------------------------------------------------------------- */
#define LEX_N 50000
typedef struct {
    char *blob; uint32_t off[LEX_N + 1];
    double *cdf;   /* Zipf(s=1.1) */
} lex_t;

static const char *ONSET[] = { "", "b", "c", "d", "f", "g", "h", "j", "k", "l", "m", "n", "p", "r", "s", "t", "v", "w",
    "st", "tr", "ch", "sh", "th", "pl", "pr", "br", "gr", "cl", "sp", "wh", "qu", "fl", "cr", "dr", "sl" };
static const char *NUCLEUS[] = { "a", "e", "i", "o", "u", "ea", "ou", "ai", "ee", "oo", "ie", "io", "au", "y" };
static const char *CODA[] = { "", "", "", "n", "r", "s", "t", "l", "d", "m", "ng", "st", "nt", "ck", "ll", "rs", "nd", "ct", "ss", "p", "x" };

static int gen_word(rng_t *r, char *w) {
    int syl = 1 + (int)(rbelow(r, 100) < 62) + (int)(rbelow(r, 100) < 30) + (int)(rbelow(r, 100) < 12);
    int n = 0;
    for (int i = 0; i < syl && n < 14; i++) {
        const char *a = ONSET[rbelow(r, sizeof ONSET / sizeof *ONSET)];
        const char *b = NUCLEUS[rbelow(r, sizeof NUCLEUS / sizeof *NUCLEUS)];
        const char *c = CODA[rbelow(r, sizeof CODA / sizeof *CODA)];
        for (const char *s = a; *s && n < 14; s++) w[n++] = *s;
        for (const char *s = b; *s && n < 14; s++) w[n++] = *s;
        for (const char *s = c; *s && n < 14; s++) w[n++] = *s;
    }
    return n;
}

static lex_t *lex_new(void) {
    lex_t *L = malloc(sizeof *L);
    L->blob = malloc(LEX_N * 15);
    L->cdf = malloc(LEX_N * sizeof(double));
    rng_t r = { 0x1234ABCDull };              /* the lexicon itself is fixed across seeds */
    uint32_t o = 0;
    static const char *COMMON[] = { "the", "of", "and", "to", "a", "in", "is", "that", "it", "was", "for", "on",
        "are", "as", "with", "his", "they", "at", "be", "this", "from", "or", "had", "by", "not", "but", "what",
        "all", "were", "we", "when", "your", "can", "said", "there", "use", "an", "each", "which", "she", "do",
        "how", "their", "if", "will", "up", "other", "about", "out", "many", "then", "them", "these", "so",
        "some", "her", "would", "make", "like", "him", "into", "time", "has", "look", "two", "more", "write",
        "go", "see", "number", "no", "way", "could", "people", "my", "than", "first", "water", "been", "call",
        "who", "its", "now", "find", "long", "down", "day", "did", "get", "come", "made", "may", "part" };
    uint32_t nc = sizeof COMMON / sizeof *COMMON;
    for (uint32_t i = 0; i < LEX_N; i++) {
        L->off[i] = o;
        if (i < nc) { size_t k = strlen(COMMON[i]); memcpy(L->blob + o, COMMON[i], k); o += (uint32_t)k; }
        else o += (uint32_t)gen_word(&r, L->blob + o);
    }
    L->off[LEX_N] = o;
    double z = 0;
    for (uint32_t i = 0; i < LEX_N; i++) { z += 1.0 / pow((double)(i + 1), 1.1); L->cdf[i] = z; }
    for (uint32_t i = 0; i < LEX_N; i++) L->cdf[i] /= z;
    return L;
}
static void lex_free(lex_t *L) { free(L->blob); free(L->cdf); free(L); }
static uint32_t lex_draw(const lex_t *L, rng_t *r) {
    double u = runit(r); uint32_t lo = 0, hi = LEX_N - 1;
    while (lo < hi) { uint32_t mid = (lo + hi) / 2; if (L->cdf[mid] < u) lo = mid + 1; else hi = mid; }
    return lo;
}

/* ---- english ----------------------------------------------------------------------- */
static int english_word(const lex_t *L, rng_t *r, buf_t *b) {
    char w[64]; int n;
    uint32_t k = rbelow(r, 1000);
    if (k < 30) {                                   /* number 1..6 digits */
        n = 1 + (int)rbelow(r, 6);
        for (int i = 0; i < n; i++) w[i] = (char)('0' + rbelow(r, 10));
    } else if (k < 70) {                            /* out-of-lexicon word */
        n = gen_word(r, w);
    } else {
        uint32_t id = lex_draw(L, r);
        n = (int)(L->off[id + 1] - L->off[id]);
        memcpy(w, L->blob + L->off[id], (size_t)n);
    }
    if (k >= 30 && rbelow(r, 100) < 10 && w[0] >= 'a' && w[0] <= 'z') w[0] = (char)(w[0] - 32);
    if (k >= 30 && rbelow(r, 1000) < 8)             /* ALL CAPS */
        for (int i = 0; i < n; i++) if (w[i] >= 'a' && w[i] <= 'z') w[i] = (char)(w[i] - 32);
    if (rbelow(r, 100) < 2) {                       /* contraction */
        static const char *CT[] = { "'s", "'t", "'re", "'ve", "'m", "'ll", "'d" };
        const char *c = CT[rbelow(r, 7)];
        for (; *c; c++) w[n++] = *c;
    }
    return put(b, w, (uint64_t)n);
}

static void gen_english_into(const lex_t *L, rng_t *r, buf_t *b, uint64_t limit, int punct) {
    uint64_t since_nl = 0, since_par = 0; int to_period = 12 + (int)rbelow(r, 14);
    buf_t lim = *b; lim.cap = limit < b->cap ? limit : b->cap;
    for (;;) {
        uint64_t before = lim.n;
        int open = 0;
        if (punct && rbelow(r, 100) < 2) { if (!putc_(&lim, rbelow(r, 2) ? '(' : '"')) break; open = 1; }
        if (!english_word(L, r, &lim)) break;
        if (open && !putc_(&lim, rbelow(r, 2) ? ')' : '"')) break;
        if (punct && rbelow(r, 100) < 10) {
            static const char P[] = ",;:-?!";
            if (!putc_(&lim, P[rbelow(r, 100) < 60 ? 0 : rbelow(r, 6)])) break;
        }
        if (--to_period <= 0) { if (!putc_(&lim, '.')) break; to_period = 12 + (int)rbelow(r, 14); }
        since_nl += lim.n - before; since_par += lim.n - before;
        if (since_par > 600) { if (!put(&lim, "\n\n", 2)) break; since_par = 0; since_nl = 0; }
        else if (since_nl > 80) { if (!putc_(&lim, '\n')) break; since_nl = 0; }
        else if (!putc_(&lim, ' ')) break;
    }
    b->n = lim.n;
}

/* ---- mixed UTF-8 ------------------------------------------------------------------- */
static void gen_mixed_span(const lex_t *L, rng_t *r, buf_t *b) {
    uint64_t span = 64 + rbelow(r, 449);
    uint64_t limit = b->n + span;
    uint32_t k = rbelow(r, 100);
    if (k < 40) { gen_english_into(L, r, b, limit, 1); return; }
    buf_t lim = *b; lim.cap = limit < b->cap ? limit : b->cap;
    if (k < 75) {                                   /* CJK ideographs, Zipf-ish over 3500 */
        static const uint32_t SEP[] = { 0xFF0C, 0x3002, 0xFF01, 0xFF1F, 0x3001 };
        for (;;) {
            int run = 5 + (int)rbelow(r, 56), ok = 1;
            for (int i = 0; i < run && ok; i++) {
                double u = runit(r); uint32_t idx = (uint32_t)(3500.0 * u * u * u);
                ok = put_utf8(&lim, 0x4E00 + idx * 5 % 0x51A6);
            }
            if (!ok || !put_utf8(&lim, SEP[rbelow(r, 5)])) break;
        }
    } else if (k < 80) {                            /* kana / hangul */
        for (;;) {
            uint32_t cp = rbelow(r, 2) ? 0x3041 + rbelow(r, 86) : 0xAC00 + rbelow(r, 11172);
            if (!put_utf8(&lim, cp)) break;
            if (rbelow(r, 12) == 0 && !putc_(&lim, ' ')) break;
        }
    } else if (k < 85) {                            /* emoji incl. VS16, ZWJ, skin tones */
        for (;;) {
            if (!put_utf8(&lim, 0x1F300 + rbelow(r, 0x1FAFF - 0x1F300))) break;
            uint32_t m = rbelow(r, 10);
            if (m == 0 && !put_utf8(&lim, 0xFE0F)) break;
            if (m == 1 && !(put_utf8(&lim, 0x200D) && put_utf8(&lim, 0x1F300 + rbelow(r, 512)))) break;
            if (m == 2 && !put_utf8(&lim, 0x1F3FB + rbelow(r, 5))) break;
            if (m >= 7 && !putc_(&lim, ' ')) break;
        }
    } else if (k < 95) {                            /* whitespace runs */
        for (;;) {
            uint32_t m = rbelow(r, 8); int ok = 1;
            if (m < 3) { int n = 1 + (int)rbelow(r, 64); for (int i = 0; i < n && ok; i++) ok = putc_(&lim, ' '); }
            else if (m == 3) { int n = 1 + (int)rbelow(r, 8); for (int i = 0; i < n && ok; i++) ok = putc_(&lim, '\t'); }
            else if (m == 4) { int n = 1 + (int)rbelow(r, 8); for (int i = 0; i < n && ok; i++) ok = putc_(&lim, '\n'); }
            else if (m == 5) ok = put(&lim, "\r\n", 2);
            else if (m == 6) ok = put_utf8(&lim, rbelow(r, 2) ? 0xA0 : 0x3000);
            else { char w[16]; int n = gen_word(r, w); ok = put(&lim, w, (uint64_t)n); }
            if (!ok) break;
        }
    } else {                                        /* digit / punctuation runs */
        static const char P[] = "0123456789.,:;/-+*=%$#@&()[]{}<>!?'\"|\\~^_`";
        for (;;) {
            int n = 1 + (int)rbelow(r, 12), ok = 1;
            if (rbelow(r, 2)) for (int i = 0; i < n && ok; i++) ok = putc_(&lim, (char)('0' + rbelow(r, 10)));
            else for (int i = 0; i < n && ok; i++) ok = putc_(&lim, P[rbelow(r, sizeof P - 1)]);
            if (!ok || !putc_(&lim, ' ')) break;
        }
    }
    b->n = lim.n;
}

/* ---- code -------------------------------------------------------------------------- */
static int code_ident(rng_t *r, buf_t *b) {
    char w[64]; int n = 0, parts = 1 + (int)rbelow(r, 3), camel = (int)rbelow(r, 2);
    for (int p = 0; p < parts && n < 40; p++) {
        char t[16]; int k = gen_word(r, t);
        if (p && !camel) w[n++] = '_';
        if (p && camel && t[0] >= 'a') t[0] = (char)(t[0] - 32);
        memcpy(w + n, t, (size_t)k); n += k;
    }
    return put(b, w, (uint64_t)n);
}
static int code_run(buf_t *b, char c, uint64_t n) {
    if (b->n + n > b->cap) return 0;
    memset(b->p + b->n, c, n); b->n += n; return 1;
}
static void gen_code(rng_t *r, buf_t *b, uint64_t nbytes) {
    static const char *KW[] = { "if", "else", "for", "while", "return", "def", "class", "import", "from", "int",
        "void", "const", "static", "struct", "let", "var", "fn", "pub", "match", "None", "True", "self" };
    static const char *OPS[] = { " = ", " == ", " != ", " + ", " - ", " * ", " / ", " -> ", " += ", " && ", " || ",
        "(", ")", "[", "]", "{", "}", ", ", ": ", ".", "::", ";", " < ", " >= " };
    static const char B64[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    /* big stressors placed at fixed fractions when the document is large enough */
    uint64_t big_at[3] = { nbytes / 5, nbytes / 2, nbytes * 4 / 5 };
    uint64_t big_len[3] = { 1u << 20, 256u << 10, 64u << 10 };
    char big_ch[3] = { 'x', ' ', '\n' };
    int big_done[3] = { 0, 0, 0 };
    for (;;) {
        for (int i = 0; i < 3; i++)
            if (!big_done[i] && nbytes >= (64u << 20) && b->n >= big_at[i]) {
                big_done[i] = 1;
                code_run(b, big_ch[i], big_len[i]); putc_(b, '\n');
            }
        uint32_t k = rbelow(r, 1000);
        int ok = 1;
        if (k < 6) {                                 /* separator runs */
            static const char RC[] = "=-#*";
            ok = code_run(b, RC[rbelow(r, 4)], 20 + rbelow(r, 181)) && putc_(b, '\n');
        } else if (k < 8) {                          /* hex / base64 blob 1..64 KiB */
            uint64_t n = 1024 + rbelow(r, 63 * 1024); int hex = (int)rbelow(r, 2);
            ok = putc_(b, '"');
            for (uint64_t i = 0; i < n && ok; i++)
                ok = putc_(b, hex ? "0123456789abcdef"[rbelow(r, 16)] : B64[rbelow(r, 64)]);
            ok = ok && put(b, "\"\n", 2);
        } else {
            uint32_t ind = rbelow(r, 9);
            if (rbelow(r, 20) == 0) ok = code_run(b, '\t', 1 + ind / 2); else ok = code_run(b, ' ', 4 * ind);
            int toks = 2 + (int)rbelow(r, 9);
            for (int t = 0; t < toks && ok; t++) {
                uint32_t m = rbelow(r, 100);
                if (m < 15) { const char *kw = KW[rbelow(r, sizeof KW / sizeof *KW)]; ok = put(b, kw, strlen(kw)) && putc_(b, ' '); }
                else if (m < 55) ok = code_ident(r, b);
                else if (m < 85) { const char *o = OPS[rbelow(r, sizeof OPS / sizeof *OPS)]; ok = put(b, o, strlen(o)); }
                else if (m < 93) { char w[24]; int n = 1 + (int)rbelow(r, 8); for (int i = 0; i < n; i++) w[i] = (char)('0' + rbelow(r, 10)); ok = put(b, w, (uint64_t)n); }
                else { ok = putc_(b, '"') && code_ident(r, b) && putc_(b, ' ') && code_ident(r, b) && putc_(b, '"'); }
            }
            ok = ok && putc_(b, '\n');
        }
        if (!ok) break;
    }
}

/* ---- entry point ------------------------------------------------------------------- */
int corpus_generate(int kind, uint64_t seed, uint64_t nbytes, uint8_t *out) {
    buf_t b = { out, 0, nbytes };
    rng_t r = { seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull };
    lex_t *L = lex_new();
    if (kind == 0) gen_english_into(L, &r, &b, nbytes, 1);
    else if (kind == 1) { while (b.n + 8 < nbytes) { uint64_t before = b.n; gen_mixed_span(L, &r, &b); if (b.n == before) break; } }
    else if (kind == 2) gen_code(&r, &b, nbytes);
    else { lex_free(L); return -1; }
    while (b.n < nbytes) out[b.n++] = ' ';
    lex_free(L);
    return 0;
}
