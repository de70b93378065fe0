/* ORACLE (test infrastructure, not product code).
 * C restatement of the reference LFSR PRN generator:
 *   amaranth_twstft/common.py:15-21  unary_xor  (parity of the masked state)
 *   amaranth_twstft/common.py:23-30  nextstate  (shift right, feed parity in at the top)
 *   amaranth_twstft/common.py:59-73  write_prn_seq (emit state%2 BEFORE stepping, seed 1,
 *                                    QPSK = bytes of generator A and B interleaved)
 *   tools/mseq_calculator.c:9-18     same step in C.
 */
#include <stdint.h>
#include <stddef.h>

static inline uint32_t oracle_parity32(uint32_t x)
{
    x ^= x >> 16; x ^= x >> 8; x ^= x >> 4; x ^= x >> 2; x ^= x >> 1;
    return x & 1u;
}

uint32_t oracle_lfsr_next(uint32_t state, uint32_t taps, int bitlen)
{
    uint32_t bit = oracle_parity32(state & taps);
    return (state >> 1) | (bit << (bitlen - 1));
}

/* out[i] = LSB of the state before the i-th step; returns the final state. */
uint32_t oracle_lfsr_bits(int bitlen, uint32_t taps, uint32_t seed, size_t n, uint8_t *out)
{
    uint32_t s = seed;
    for (size_t i = 0; i < n; i++) {
        out[i] = (uint8_t)(s & 1u);
        s = oracle_lfsr_next(s, taps, bitlen);
    }
    return s;
}
