// hostcheck.cpp -- TEST-ONLY host build of the engine's __host__ __device__ building blocks.
// Compiled with plain g++ into libb200bpe_hostcheck.so and loaded only by tests (-m "not gpu"):
// it lets the CPU-only suite exercise exactly the rule / merge code the CUDA kernels execute.
// It is NOT part of libb200bpe.so and is never reachable from the product path.
#include <cstdint>
#include <cstring>
#include <vector>
#include "text_access.cuh"
#include "unicode_classes.inc"

using namespace b200bpe;

extern "C" int hc_piece_starts(int pattern, const uint8_t *text, int64_t n, const uint64_t *doc_off,
                               int64_t n_docs, uint8_t *is_start /* n bytes, 0/1 */) {
    std::vector<uint8_t> padded((size_t)n + 8, 0);
    memcpy(padded.data(), text, (size_t)n);
    std::vector<uint32_t> dbits((size_t)(n + 63) / 32 + 1, 0);
    for (int64_t d = 0; d < n_docs; d++) {
        uint64_t o = doc_off[d];
        if ((int64_t)o < n) dbits[o >> 5] |= 1u << (o & 31);
    }
    uint8_t ascii[128];
    for (int i = 0; i < 128; i++) ascii[i] = UC_STAGE2[(uint32_t)UC_STAGE1[0] * 256 + i];
    TextAccess t{padded.data(), n, dbits.data(), UC_STAGE1, UC_STAGE2, ascii};
    for (int64_t pos = 0; pos < n; pos++) {
        uint8_t b = padded[pos];
        bool lead = (b & 0xC0u) != 0x80u;
        bool s;
        if (t.doc_start(pos)) s = true;
        else if (!lead) s = false;
        else if (pattern == PAT_R50K) s = boundary_before<PAT_R50K>(t, pos);
        else if (pattern == PAT_CL100K) s = boundary_before<PAT_CL100K>(t, pos);
        else if (pattern == PAT_O200K) s = boundary_before<PAT_O200K>(t, pos);
        else return -1;
        is_start[pos] = s ? 1 : 0;
    }
    return 0;
}
