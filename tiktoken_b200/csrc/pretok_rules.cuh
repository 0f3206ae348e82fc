// pretok_rules.cuh -- position-parallel restatement of tiktoken's three pre-tokeniser regexes.
//
// The reference walks the text sequentially with a backtracking regex engine
// (`regex.find_iter(text)`, src/lib.rs:365 and :405; patterns in
// tiktoken_ext/openai_public.py:12-14 [r50k family], :89 [cl100k], :104-114 [o200k]).
// On the GPU every scalar value must decide ON ITS OWN whether a piece starts at it.  For these
// three patterns that decision is a function of the character classes in a small window
// around the position plus a few "scan along a homogeneous run" look-ups (how many digits
// precede, does a CR/LF follow inside this whitespace run, ...).  `boundary_before<PAT>(t, pos)`
// is that function; it is exact (not a heuristic): tests/test_pretok_rules.py checks it against
// the literal backtracking matcher in oracle/bpe_oracle.c on every string over per-pattern
// alphabets up to length 7-8 plus random long strings, and the oracle itself is pinned to the
// reference engine.
//
// The functions are written against an abstract text accessor T so that the SAME code runs in
// the CUDA kernel (device accessor over HBM) and in the CPU-only rule tests (host accessor):
//   int      T::cls(int64 pos)   class of the scalar whose lead byte is at pos
//   int64    T::prev(int64 pos)  lead byte of the previous scalar of the same document, or -1
//   int64    T::next(int64 pos)  lead byte of the next scalar of the same document, or -1
//   unsigned T::byte(int64 pos)  raw byte
// Documents are separate haystacks: prev/next never cross a document start, which is what
// makes `$` / `\s++$` / `(?!\S)` see the document end (src/lib.rs:405 slices per haystack).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define B2_HD __host__ __device__ __forceinline__
#define B2_HDN __host__ __device__ __noinline__
#else
#define B2_HD inline
#define B2_HDN inline
#endif

namespace b2bpe {

enum : int {
    C_O = 0, C_LU = 1, C_LL = 2, C_LB = 3, C_M = 4, C_N = 5, C_SP = 6, C_WS = 7, C_NL = 8,
    C_APOS = 9, C_SLASH = 10, C_CONT = 15
};
enum : int { PAT_R50K = 0, PAT_CL100K = 1, PAT_O200K = 2 };

B2_HD bool is_letter(int c) { return c >= C_LU && c <= C_LB; }                 // \p{L}
B2_HD bool is_ws(int c) { return c >= C_SP && c <= C_NL; }                     // \s
B2_HD bool is_other(int c) { return c == C_O || c == C_M || c >= C_APOS; }     // [^\s\p{L}\p{N}]
B2_HD bool is_xo(int c) { return c == C_O || c >= C_APOS; }                    // o200k "X": other minus \p{M}
B2_HD bool is_letterish(int c) { return c >= C_LU && c <= C_M; }               // o200k word alphabet

// ASCII letter that the scalar at pos folds onto (lower case), or 0.  With ci the match is
// (?i:...) under Unicode simple case folding: upper case ASCII and U+017F (-> s) also match.
template <class T>
B2_HD int fold_letter(const T &t, int64_t pos, bool ci) {
    unsigned b = t.byte(pos);
    if (b >= 'a' && b <= 'z') return (int)b;
    if (ci && b >= 'A' && b <= 'Z') return (int)(b | 0x20);
    if (ci && b == 0xC5 && t.byte(pos + 1) == 0xBF) return 's';
    return 0;
}

// `'(?:[sdmt]|ll|ve|re)` (r50k, cl100k with (?i)) == `(?i:'s|'t|'re|'ve|'m|'ll|'d)` (o200k):
// number of scalars matched at the apostrophe q (2 or 3), or 0.
template <class T>
B2_HD int contraction_len(const T &t, int64_t q, bool ci) {
    int64_t a = t.next(q);
    if (a < 0) return 0;
    int x = fold_letter(t, a, ci);
    if (x == 's' || x == 'd' || x == 'm' || x == 't') return 2;
    if (x != 'l' && x != 'v' && x != 'r') return 0;
    int64_t b = t.next(a);
    if (b < 0) return 0;
    int y = fold_letter(t, b, ci);
    if ((x == 'l' && y == 'l') || (x == 'v' && y == 'e') || (x == 'r' && y == 'e')) return 3;
    return 0;
}

// ------------------------------------------------------------------------------------------
// r50k family and cl100k
// ------------------------------------------------------------------------------------------
// An apostrophe is tried against alternative 1 only if a match starts at it.  A match starts
// at q unless the previous scalar drags q into its own piece: another "other" char (the
// possessive run ` ?[^\s\p{L}\p{N}]++` continues) or a space (which always joins a following
// "other" char, because the whitespace alternatives leave the last space of a run unmatched).
template <class T>
B2_HD bool apos_is_match_start(const T &t, int64_t q) {
    int64_t pq = t.prev(q);
    if (pq < 0) return true;
    int pc = t.cls(pq);
    return !(is_other(pc) || pc == C_SP);
}

// Does a contraction piece end exactly before pos?  (pos's predecessor is a letter.)
template <class T>
B2_HD bool contraction_ends_before(const T &t, int64_t p1, bool ci) {
    int64_t p2 = t.prev(p1);
    if (p2 < 0) return false;
    if (t.cls(p2) == C_APOS) return contraction_len(t, p2, ci) == 2 && apos_is_match_start(t, p2);
    int64_t p3 = t.prev(p2);
    if (p3 < 0 || t.cls(p3) != C_APOS) return false;
    return contraction_len(t, p3, ci) == 3 && apos_is_match_start(t, p3);
}

template <class T>
B2_HD bool boundary_r50k(const T &t, int64_t pos) {
    const int c = t.cls(pos);
    const int64_t pp = t.prev(pos);
    const int p = t.cls(pp);
    if (is_ws(c)) {
        if (!is_ws(p)) return true;
        // inside a whitespace run: `\s+(?!\S)` gives back exactly the last char when a
        // non-space follows; `\s++$` keeps a run that reaches the end whole.
        int64_t nn = t.next(pos);
        return nn >= 0 && !is_ws(t.cls(nn));
    }
    if (is_letter(c)) {
        if (p == C_APOS && apos_is_match_start(t, pp) && contraction_len(t, pp, false) >= 2) return false;
        if (is_letter(p)) return contraction_ends_before(t, pp, false);
        return p != C_SP;
    }
    if (c == C_N) return !(p == C_N || p == C_SP);
    return !(is_other(p) || p == C_SP);          // "other", incl. an apostrophe starting alt. 1
}

B2_HD int b2_ctz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
B2_HD int b2_ctz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)x) - 1;
#else
    return __builtin_ctzll(x);
#endif
}
// What lies ahead of the whitespace scalar at `pos` in its run of whitespace: 1 = a CR/LF, 2 = the end of the document,
// 0 = a non-whitespace scalar.  Runs of ASCII whitespace are walked eight bytes at a time (SWAR on one aligned 64-bit word
// + the doc-start bits of its eight positions); anything else takes the scalar-by-scalar accessor.  (The mixed-script
// workload has whitespace runs of hundreds of bytes behind line ends: the per-scalar walk made the listed positions as
// expensive as the whole bit-parallel kernel.)
template <class T>
B2_HD int ws_run_ahead(const T &t, int64_t pos) {
    int64_t q = t.next(pos);
    for (;;) {
        if (q < 0) return 2;
        const int64_t a = q & ~(int64_t)7;
        const uint64_t w = *reinterpret_cast<const uint64_t *>(t.text + a);
        const uint64_t H = 0x8080808080808080ull, L = 0x0101010101010101ull;
        const uint64_t y = w & ~H, hi = w & H;
        const uint64_t ge9 = (y + (0x80 - 0x09) * L) & H, geE = (y + (0x80 - 0x0E) * L) & H;
        const uint64_t eq20 = ~(((y ^ (0x20 * L)) + (0x7F * L))) & H;
        const uint64_t eqA = ~(((y ^ (0x0A * L)) + (0x7F * L))) & H, eqD = ~(((y ^ (0x0D * L)) + (0x7F * L))) & H;
        const uint64_t nl = (eqA | eqD) & ~hi;                               // bit 7 of the byte lanes that hold CR / LF
        const uint64_t ws = ((ge9 & ~geE) | eq20) & ~hi;                     // ... ASCII whitespace (CR / LF included)
        uint32_t d = (t.dbits[a >> 5] >> (a & 31)) & 0xFFu;                  // documents that start at a .. a+7
        const int k0 = (int)(q - a);
        d &= ~((2u << k0) - 1u);                                             // only AFTER q: q itself is inside the document
        uint64_t stop = ~ws & H;                                             // lanes that are not ASCII whitespace
        for (uint32_t dd = d; dd; dd &= dd - 1) stop |= 0x80ull << (8 * (b2_ctz32(dd)));
        if (a + 8 > t.n) for (int k = (int)(t.n - a); k < 8; k++) stop |= 0x80ull << (8 * (k < 0 ? 0 : k));
        const uint64_t from = ~0ull << (8 * k0);
        const uint64_t ev = (stop | nl) & from;
        if (!ev) { q = a + 8; if (q >= t.n) return 2; if (t.doc_start(q)) return 2; continue; }
        const int k = b2_ctz64(ev) >> 3;
        const uint64_t lane = 0x80ull << (8 * k);
        if (!(stop & lane)) return 1;                                        // CR / LF first
        if (a + k >= t.n || ((d >> k) & 1u)) return 2;                       // the document ends first
        if (!(hi & lane)) return 0;                                          // an ASCII non-whitespace byte
        q = a + k;                                                           // non-ASCII: one scalar through the accessor
        const int cq = t.cls(q);
        if (cq == C_NL) return 1;
        if (!is_ws(cq)) return 0;
        q = t.next(q);
    }
}

template <class T>
B2_HD bool boundary_cl100k(const T &t, int64_t pos) {
    const int c = t.cls(pos);
    const int64_t pp = t.prev(pos);
    const int p = t.cls(pp);
    if (is_letter(c)) {
        if (p == C_APOS && apos_is_match_start(t, pp) && contraction_len(t, pp, true) >= 2) return false;
        if (is_letter(p)) return contraction_ends_before(t, pp, true);
        if (p == C_N || p == C_NL) return true;
        if (p == C_SP || p == C_WS) return false;        // last char of a ws run is always free
        // p is "other": it is the optional one-char prefix iff a match starts at it
        int64_t p2 = t.prev(pp);
        if (p2 < 0) return false;
        int c2 = t.cls(p2);
        return is_other(c2) || c2 == C_SP;
    }
    if (c == C_N) {
        if (p != C_N) return true;
        int k = 1;                                       // digits before pos in this run, mod 3
        for (int64_t q = t.prev(pp); q >= 0 && t.cls(q) == C_N; q = t.prev(q)) k++;
        return k % 3 == 0;
    }
    if (c == C_NL) return is_letter(p) || p == C_N;      // joins "other"[\r\n]* and any ws run
    if (c == C_SP || c == C_WS) {
        if (!is_ws(p)) return true;
        if (p != C_NL) {                                 // k-1 rule of `\s+(?!\S)`
            int64_t nn = t.next(pos);
            return nn >= 0 && !is_ws(t.cls(nn));
        }
        // previous scalar is CR/LF: `\s++$` / `\s*[\r\n]` swallow pos iff the run reaches the
        // document end or has another CR/LF ahead -- unless the CR/LFs before pos were taken by
        // a preceding `[^\s\p{L}\p{N}]++[\r\n]*+` piece, in which case a fresh match starts here.
        if (ws_run_ahead(t, pos) == 0) return true;
        int64_t q = pp;
        while (q >= 0 && t.cls(q) == C_NL) q = t.prev(q);
        return q >= 0 && is_other(t.cls(q));
    }
    return !(is_other(p) || p == C_SP);                  // "other"
}

// ------------------------------------------------------------------------------------------
// o200k
// ------------------------------------------------------------------------------------------
// States of a scalar of class X (other minus \p{M}), \p{M} or CR/LF inside its piece.
enum : int {
    XS_WORD = 0,      // \p{M} acting as a word character
    XS_PREFIX = 1,    // X taken as the optional one-char prefix of a word
    XS_XRUN = 2,      // inside ` ?[^\s\p{L}\p{N}]+`
    XS_TRAIL = 3,     // CR/LF or '/' absorbed by the trailing `[\r\n/]*`
    XS_WSNL = 4,      // CR/LF that belongs to a whitespace piece
    XS_SUFFIX = 5     // apostrophe of a contraction suffix attached to a word
};
struct XmState { int st; bool start; };

template <class T> B2_HDN bool o200k_suffix_at(const T &t, int64_t q);
template <class T> B2_HDN XmState o200k_xm_state(const T &t, int64_t q);

// Is the scalar before an apostrophe a word character (letter, or \p{M} in word mode)?
template <class T>
B2_HD bool o200k_wordchar(const T &t, int64_t p) {
    int c = t.cls(p);
    if (is_letter(c)) return true;
    if (c != C_M) return false;
    return o200k_xm_state(t, p).st == XS_WORD;
}

// Does a contraction SUFFIX `(?i:'s|'t|'re|'ve|'m|'ll|'d)?` start at apostrophe q?  It does iff
// a word piece ends right before q.  A word char right before q could itself be the last letter
// of an earlier suffix (then the piece ended there and q starts a new match: "a's's" is
// a's | 's), so walk the chain of back-to-back candidates and take its parity.
template <class T>
B2_HDN bool o200k_suffix_at(const T &t, int64_t q) {
    int count = 0;
    int64_t cur = q;
    for (;;) {
        if (t.cls(cur) != C_APOS || contraction_len(t, cur, true) == 0) break;
        int64_t b1 = t.prev(cur);
        if (b1 < 0 || !o200k_wordchar(t, b1)) break;
        count++;
        int64_t b2 = t.prev(b1);
        if (b2 < 0) break;
        if (t.cls(b2) == C_APOS) {
            if (contraction_len(t, b2, true) == 2) { cur = b2; continue; }
            break;
        }
        int64_t b3 = t.prev(b2);
        if (b3 >= 0 && t.cls(b3) == C_APOS && contraction_len(t, b3, true) == 3) { cur = b3; continue; }
        break;
    }
    return (count & 1) != 0;
}

template <class T>
B2_HD bool o200k_contraction_ends_before(const T &t, int64_t p1) {   // p1 = letter right before pos
    int64_t p2 = t.prev(p1);
    if (p2 < 0) return false;
    if (t.cls(p2) == C_APOS) return contraction_len(t, p2, true) == 2 && o200k_suffix_at(t, p2);
    int64_t p3 = t.prev(p2);
    if (p3 < 0 || t.cls(p3) != C_APOS) return false;
    return contraction_len(t, p3, true) == 3 && o200k_suffix_at(t, p3);
}

// Resolve the role of scalar q (class X, \p{M} or CR/LF).  Walk back to the start of the
// maximal chain of such scalars, then replay the chain forwards with the tiny automaton that
// the alternatives 1, 2, 4, 5 induce on it.  Chains are 1-3 scalars long in ordinary text.
template <class T>
B2_HDN XmState o200k_xm_state(const T &t, int64_t q) {
    int64_t r = q;
    for (;;) {
        int64_t pr = t.prev(r);
        if (pr < 0) break;
        int pc = t.cls(pr);
        if (!(is_xo(pc) || pc == C_M || pc == C_NL)) break;
        r = pr;
    }
    const int64_t ctx_pos = t.prev(r);
    const int ctx = ctx_pos < 0 ? -1 : t.cls(ctx_pos);     // -1: document start
    int st = -1;                                            // -1: a match starts at the next scalar
    const bool joined_sp = (ctx == C_SP);                   // a free space joins a following X (` ?`)
    const bool joined_ws = (ctx == C_SP || ctx == C_WS);    // any free non-CR/LF ws prefixes a word
    XmState out = { XS_XRUN, false };
    for (int64_t e = r;; e = t.next(e)) {
        // On well-formed UTF-8 the forward walk lands exactly on q.  On malformed input next() can step over
        // it: stop there (the answer for such bytes is unspecified, the walk must still end).
        if (e < 0 || e > q) { out.st = st < 0 ? XS_XRUN : st; break; }
        const int ce = t.cls(e);
        bool start = false;
        if (ce == C_NL) {
            st = (st == XS_XRUN || st == XS_TRAIL) ? XS_TRAIL : XS_WSNL;
        } else if (ce == C_M) {
            if (st == XS_XRUN) st = XS_XRUN;
            else if (st == XS_WORD || st == XS_PREFIX) st = XS_WORD;
            else {
                // first word character of a piece -- unless the chain starts right after a letter,
                // in which case \p{M} simply continues that word (or starts one after a suffix)
                if (e == r && ctx >= 0 && is_letter(ctx)) start = o200k_contraction_ends_before(t, ctx_pos);
                else start = !(e == r && joined_ws);
                st = XS_WORD;
            }
        } else {                                            // X: O, apostrophe, slash
            if (st == XS_XRUN) st = XS_XRUN;
            else if (st == XS_TRAIL && ce == C_SLASH) st = XS_TRAIL;
            else {
                bool word_before = (st == XS_WORD) || (e == r && ctx >= 0 && is_letter(ctx));
                if (word_before && ce == C_APOS && o200k_suffix_at(t, e)) st = XS_SUFFIX;
                else if (e == r && joined_sp) st = XS_XRUN;                   // " !" joined
                else {
                    start = true;
                    int64_t ne = t.next(e);
                    st = (ne >= 0 && is_letterish(t.cls(ne))) ? XS_PREFIX : XS_XRUN;
                }
            }
        }
        if (e == q) { out.st = st; out.start = start; break; }
    }
    return out;
}

template <class T>
B2_HD bool boundary_o200k(const T &t, int64_t pos) {
    const int c = t.cls(pos);
    const int64_t pp = t.prev(pos);
    const int p = t.cls(pp);
    if (is_letter(c)) {
        if (p == C_N || p == C_NL) return true;
        if (p == C_SP || p == C_WS) return false;
        bool p_word = is_letter(p);
        if (!p_word) {                                   // p is X or \p{M}
            XmState s = o200k_xm_state(t, pp);
            if (s.st == XS_PREFIX || s.st == XS_SUFFIX) return false;
            if (s.st != XS_WORD) return true;
        }
        // pos continues a run of word characters
        if (is_letter(p) && o200k_contraction_ends_before(t, pp)) return true;
        if (c != C_LU) return false;                     // lower / other letters extend any word
        if (p == C_LL) {
            // lower -> upper ends `[L]+` ... unless both sit inside a 3-scalar suffix ('lL etc.)
            int64_t p2 = t.prev(pp);
            if (p2 >= 0 && t.cls(p2) == C_APOS && contraction_len(t, p2, true) == 3 && o200k_suffix_at(t, p2))
                return false;
            return true;
        }
        if (p == C_LU) return false;
        // p is Lm/Lo/M (in both [U] and [L]).  Phase: are we already inside `[L]+`?
        bool phase_b = false;
        for (int64_t k = pp;;) {
            int64_t pk = t.prev(k);
            if (pk < 0) break;
            int ck = t.cls(pk);
            if (ck == C_LL) { phase_b = !o200k_contraction_ends_before(t, pk); break; }
            if (ck == C_LB) { k = pk; continue; }
            if (ck == C_M && o200k_xm_state(t, pk).st == XS_WORD) { k = pk; continue; }
            break;
        }
        if (phase_b) return true;
        // still inside `[U]*`: alternative 1 ends at the LAST [L]-capable char of the upper run
        // when no lower-case letter follows it; that is pp iff only Lu/Lt follow up to the run end.
        for (int64_t k = pos;;) {
            int64_t nk = t.next(k);
            if (nk < 0) return true;
            int cn = t.cls(nk);
            if (cn == C_LU) { k = nk; continue; }
            return !(cn == C_LL || cn == C_LB || cn == C_M);
        }
    }
    if (c == C_N) {
        if (p != C_N) return true;
        int k = 1;
        for (int64_t q = t.prev(pp); q >= 0 && t.cls(q) == C_N; q = t.prev(q)) k++;
        return k % 3 == 0;
    }
    if (c == C_SP || c == C_WS) {
        if (!is_ws(p)) return true;
        if (p != C_NL) {
            int64_t nn = t.next(pos);
            return nn >= 0 && !is_ws(t.cls(nn));
        }
        // o200k has no `\s++$`: `\s*[\r\n]+` is tried first, so only a CR/LF further on in the run
        // (not the document end) keeps pos inside the piece of the preceding CR/LF.
        if (ws_run_ahead(t, pos) != 1) return true;
        return o200k_xm_state(t, pp).st == XS_TRAIL;
    }
    if (c == C_NL) {
        if (p == C_NL || p == C_SP || p == C_WS) {
            // CR/LF after whitespace stays in that piece; after an absorbed CR/LF it stays absorbed
            return false;
        }
        if (is_letter(p) || p == C_N) return true;
        return o200k_xm_state(t, pos).st != XS_TRAIL;    // p is X or \p{M}
    }
    // X or \p{M}
    if (c == C_M && is_letter(p)) return o200k_contraction_ends_before(t, pp);
    return o200k_xm_state(t, pos).start;
}

template <int PAT, class T>
B2_HD bool boundary_before(const T &t, int64_t pos) {
    if (PAT == PAT_R50K) return boundary_r50k(t, pos);
    if (PAT == PAT_CL100K) return boundary_cl100k(t, pos);
    return boundary_o200k(t, pos);
}

}  // namespace b2bpe
