// text_access.cuh -- the accessor the pre-tokeniser rules run against (see pretok_rules.cuh).
// UTF-8 text in one flat buffer, documents delimited by a doc-start bitmask (bit i set = a
// document begins at byte i).  The buffer must be readable up to n+3 (padding) so that a
// truncated multi-byte sequence at the very end cannot fault.
#pragma once
#include "pretok_rules.cuh"

namespace b2bpe {

struct TextAccess {
    const uint8_t *text;
    int64_t n;
    const uint32_t *dbits;          // doc-start bitmask, (n+31)/32 words
    const uint16_t *stage1;         // Unicode class table, 256-code-point blocks
    const uint8_t *stage2;
    const uint8_t *ascii;           // 128 class bytes for U+0000..U+007F
    uint32_t one = 1u;              // the constant 1, opaque to the device compiler (see classify_word)

    B2_HD unsigned byte(int64_t pos) const { return text[pos]; }
    B2_HD bool doc_start(int64_t pos) const { return (dbits[pos >> 5] >> (pos & 31)) & 1u; }
    B2_HD int cls(int64_t pos) const {
        unsigned b = text[pos];
        if (b < 0x80) return ascii[b];
        uint32_t cp;
        if (b < 0xE0) cp = ((b & 0x1Fu) << 6) | (text[pos + 1] & 0x3Fu);
        else if (b < 0xF0) cp = ((b & 0x0Fu) << 12) | ((text[pos + 1] & 0x3Fu) << 6) | (text[pos + 2] & 0x3Fu);
        else cp = ((b & 0x07u) << 18) | ((text[pos + 1] & 0x3Fu) << 12) | ((text[pos + 2] & 0x3Fu) << 6) | (text[pos + 3] & 0x3Fu);
        if (cp >= 0x110000u) return C_O;
        return stage2[(uint32_t)stage1[cp >> 8] * 256u + (cp & 255u)];
    }
    B2_HD int64_t prev(int64_t pos) const {
        if (pos <= 0 || doc_start(pos)) return -1;
        int64_t q = pos - 1;
        while (q > 0 && (text[q] & 0xC0u) == 0x80u && !doc_start(q)) q--;
        return q;
    }
    B2_HD int64_t next(int64_t pos) const {
        unsigned b = text[pos];
        int64_t q = pos + (b < 0x80 ? 1 : b < 0xE0 ? 2 : b < 0xF0 ? 3 : 4);
        if (q >= n || doc_start(q)) return -1;
        return q;
    }
};

}  // namespace b2bpe
