// kernels_decode.cuh -- CoreBPE::decode_bytes (src/lib.rs:345-358) for a whole batch.
#pragma once
#include "dev_common.cuh"

using namespace b2bpe;

// --------------------------------------------------------------------------------------------
// decode ("next" row): CoreBPE::decode_bytes (src/lib.rs:345-358) for a whole batch -- token id ->
// byte string gather.  Length look-up, two-level scan (the same scan kernels), copy.
// --------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) decode_len_kernel(const uint32_t *__restrict__ tokens, unsigned long long n,
                                                        const uint32_t *__restrict__ tok_boff, uint32_t n_ids,
                                                        uint32_t *__restrict__ len, Counters *ctr) {
    unsigned long long i = blockIdx.x * 256ull + threadIdx.x;
    if (i >= n) return;
    const uint32_t t = tokens[i];
    uint32_t l = 0;
    if (t < n_ids) l = __ldg(tok_boff + t + 1) - __ldg(tok_boff + t);
    if (l == 0) {                                           // unknown id (every real token has >= 1 byte)
        if (atomicOr(&ctr->err, ERR_BADTOKEN) == 0 || true) atomicMin(&ctr->ticket, (unsigned int)min(i, 0xFFFFFFFFull));
    }
    len[i] = l;
}

__global__ void __launch_bounds__(256) decode_copy_kernel(const uint32_t *__restrict__ tokens, unsigned long long n,
                                                         const uint32_t *__restrict__ tok_boff, uint32_t n_ids,
                                                         const uint8_t *__restrict__ blob,
                                                         const unsigned long long *__restrict__ base,
                                                         uint8_t *__restrict__ out) {
    unsigned long long i = blockIdx.x * 256ull + threadIdx.x;
    if (i >= n) return;
    const uint32_t t = tokens[i];
    if (t >= n_ids) return;
    const uint32_t b0 = __ldg(tok_boff + t), b1 = __ldg(tok_boff + t + 1);
    uint8_t *dst = out + base[i];
    for (uint32_t k = b0; k < b1; k++) *dst++ = __ldg(blob + k);
}

__global__ void decode_doc_off_kernel(const unsigned long long *__restrict__ tok_off, unsigned long long n_docs,
                                      const unsigned long long *__restrict__ base, unsigned long long *byte_off) {
    unsigned long long d = blockIdx.x * 256ull + threadIdx.x;
    if (d <= n_docs) byte_off[d] = base[tok_off[d]];
}

